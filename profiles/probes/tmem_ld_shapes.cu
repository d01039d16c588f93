// Probe: which (lane, column) does each register of tcgen05.ld.16x256b.x2 hold?  (layout check before using the shape
// in the edge kernel's epilogue).  TMEM is filled with tcgen05.st.32x32b (lane = row, register k = column k) with the
// value row * 1000 + col, then read back with the 16x256b shape at lane offsets 0 and 16 of every quarter.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tmem_probe profiles/probes/tmem_ld_shapes.cu && /tmp/tmem_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(128, 1) k_probe(uint32_t* out) {
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(32u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot;
  const uint32_t row = warp * 32 + lane;
  // st 32x32b.x16: thread's lane <- 16 columns
  uint32_t v[16];
  for (int c = 0; c < 16; ++c) v[c] = row * 1000 + c;
  const uint32_t taddr = base + ((uint32_t)(warp * 32) << 16);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int hh = 0; hh < 2; ++hh) {
    uint32_t r[8];
    const uint32_t ta = base + ((uint32_t)(warp * 32 + hh * 16) << 16);
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(ta));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int k = 0; k < 8; ++k) out[((warp * 2 + hh) * 32 + lane) * 8 + k] = r[k];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(32u));
}

int main() {
  uint32_t* d;
  cudaMalloc(&d, 4 * 2 * 32 * 8 * 4);
  cudaMemset(d, 0xff, 4 * 2 * 32 * 8 * 4);
  k_probe<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  static uint32_t h[4 * 2 * 32 * 8];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  for (int w = 0; w < 4; w += 3)
    for (int hh = 0; hh < 2; ++hh)
      for (int l = 0; l < 32; ++l) {
        printf("warp %d half %d lane %2d:", w, hh, l);
        for (int k = 0; k < 8; ++k) printf(" (%3u,%2u)", h[((w * 2 + hh) * 32 + l) * 8 + k] / 1000, h[((w * 2 + hh) * 32 + l) * 8 + k] % 1000);
        printf("\n");
      }
  return 0;
}
