# A/B of programmatic dependent launch in the bench (DFM_PDL=0 / 1, interleaved, two runs each)
mkdir -p gpurun_out
B="--steps 20 --warmup 5 --no-cpu-baseline --no-full-job"
for i in 1 2; do for pdl in 1 0; do
  DFM_PDL=$pdl timeout 600 python bench.py $B > gpurun_out/bench_pdl${pdl}_$i.json 2> gpurun_out/bench_pdl${pdl}_$i.err
  python - $pdl gpurun_out/bench_pdl${pdl}_$i.json <<'PY'
import json, sys
d = json.load(open(sys.argv[2]))
oc = d.get("other_configs", {})
print("PDL=%s value %.0f ms/step %.3f edge ms %.4f e2e %.0f | c2 us/step %.0f c4 %.0f" % (sys.argv[1], d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"], d["e2e"]["value"],
      oc.get("c2", {}).get("us_per_lockstep_step", 0), oc.get("c4", {}).get("poses_per_s", 0)))
PY
done; done
