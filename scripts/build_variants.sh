# Builds edge-kernel variants HERE (no GPU needed): $VARIANTS = ';'-separated "name:nvcc -D switches".  Each variant gets its own
# edge_ws.o linked with the default build's other objects into dfmdock_b200/lib_variants/<name>.so (git-ignored, travels with gpurun);
# scripts/gpu_r02_ab4.sh swaps them in on the GPU box.
set -e
cd "$(dirname "$0")/.."
python -m dfmdock_b200.build > /dev/null
mkdir -p dfmdock_b200/lib_variants
IFS=';' read -ra VS <<< "$VARIANTS"
for v in "${VS[@]}"; do
  name="${v%%:*}"; flags="${v#*:}"
  (
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr $flags \
      -c dfmdock_b200/csrc/edge_ws.cu -o dfmdock_b200/lib_variants/$name.edge_ws.o
    objs=$(ls dfmdock_b200/lib/*.o | grep -v edge_ws.o)
    nvcc -shared -o dfmdock_b200/lib_variants/$name.so $objs dfmdock_b200/lib_variants/$name.edge_ws.o -lcudart
    echo "built $name [$flags]"
  ) &
done
wait
