// Tile hand-over inside the fused last-layer kernel (edge_ws.cu k_last_fused): the CTAs of one launch are split into an
// edge role (P CTAs: edge MLP + gate of the ligand residues' tiles, as in every other layer) and a coordinate-head role
// (the remaining CTAs: coord_mlp on the gated messages, src/models/egnn.py:118-137).  A tile of gated messages (128 rows x
// 256 fp16 = 64 KB) travels through a ring of NR slots in global memory that is small enough to stay in the 126 MB L2, so
// the 1.26 GB per step that the two-kernel version wrote to and read back from HBM never leave the chip.
//
// Order: producer p handles its logical tiles p * chunk + tau, tau = 0, 1, ..; the hand-over sequence number of a tile is
// s = tau * P + p (what all producers finish at about the same time is adjacent in s), ring slot s % NR, epoch s / NR.
// Consumer c takes s = c, c + C, c + 2 C, ...  Flags (zeroed by the host before the launch):
//   ready[slot]  += 1 by each of the 8 epilogue warps once its rows of the slot are written (release, gpu scope)
//   done[slot]   += 1 by the consumer once its TMA loads of the slot have landed in shared memory
// A producer may write epoch e of a slot when done[slot] >= e, a consumer may read it when ready[slot] >= 8 (e + 1).
// Every wait points to a strictly smaller sequence number or to the producer of the same one, and all CTAs of the launch are
// co-resident (grid <= number of SMs, one CTA per SM), so there is no cycle; a lost signal traps instead of hanging.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

struct LastRing {
  int P;                 // CTAs in the edge role; the coordinate-head role has gridDim.x - P
  int chunk;             // logical tiles per producer
  int ntiles;            // logical tiles of the ligand-only walk (B * tpt)
  int tpt, N, R;         // tiles per trajectory, residues, receptor residues
  int total_nodes;       // B * N
  int NR;                // ring slots
  unsigned int* ready;   // [NR]
  unsigned int* done;    // [NR]
  __half* ring;          // [NR * 128, 256] fp16 (m* x 2^-6, columns in the epilogue's fragment order)
};

#ifdef __CUDACC__
// logical tile of the ligand-only walk -> tile of the [B*N/2] node-pair grid (same rule as ews::k_edge_ws)
__device__ __forceinline__ int ring_phys(const LastRing& r, int lt) {
  const int b = lt / r.tpt, k = lt - b * r.tpt;
  const int base = b * r.N;
  return min(((base + r.R) >> 1) + k, (base + r.N - 1) >> 1);
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void ring_wait_ge(const unsigned int* p, unsigned int v) {
  unsigned int tries = 0;
  while (ld_acquire_gpu(p) < v) {
    __nanosleep(64);
    if (++tries > (1u << 22)) __trap();     // ~0.5 s: a lost hand-over must fail loudly, never hang the GPU
  }
}
// generic-proxy global writes of other threads (made visible by an acquire) -> async-proxy (TMA) reads, and the reverse
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
#endif
