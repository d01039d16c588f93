// Node-side tcgen05 kernels, transposed formulation: D^T[feature, row] = W[feature, K] * X[row, K]^T.
// The weight image is the A operand (M = 128 output features per MMA), the activation tile is the B operand
// (N = rows of the tile), so in the epilogue a TMEM lane is an output FEATURE and the 32 lanes of a warp write 32
// consecutive features of one row: every global load / store of the epilogue is a full coalesced line (the
// row-per-lane formulation of node_tc.cu spent 8x the LSU wavefronts on 16-byte pieces of 32 different rows).
//
//   MODE_AB  Ah = fp16((W1s h + b1)/2), Bm = fp16((W1d h)/2)   node halves of edge_mlp.0   (src/models/egnn.py:95-104)
//   MODE_Z   z = W3h h + W3a agg + b3                           node_mlp.0 on [h, agg]      (src/models/egnn.py:106-116)
//   MODE_H   h += W4 SiLU(GraphNorm(z)) + b4  (+ fp16 copy)     node_mlp.1-3 + residual     (src/models/egnn.py:74,106-116)
//
// One persistent CTA per SM, 19 warps: 16 workers (epilogue; MODE_H also builds its operand), 1 loader warp whose elected
// thread issues one TMA tensor-map load (cp.async.bulk.tensor.2d, 64 columns x ROWS rows, SWIZZLE_128B, rows past M
// zero-filled by the hardware) per K block, 1 MMA issuer.  Tiles are 64 rows (AB, H: two 128-feature halves -> 2 x 64 accumulator columns) or 128 rows (Z: one
// 128-feature half of the output per grid half, K = 512); the operand ring holds 64 KB = 8 or 4 K blocks, the
// accumulator (128 TMEM columns per tile) is quadruple buffered.
#include <stdio.h>
#include <string.h>

#include "common.cuh"
#include "tma.cuh"

namespace ntt {

constexpr uint32_t W_BYTES = 256 * 256 * 2;          // 128 KB in every mode
constexpr uint32_t OFF_W = 0;
constexpr uint32_t OFF_S = W_BYTES;                  // operand ring, 64 KB
#ifndef NTT_RING_KB
#define NTT_RING_KB 64
#endif
constexpr uint32_t RING_BYTES = NTT_RING_KB * 1024;
constexpr uint32_t OFF_VEC = OFF_S + RING_BYTES;     // 256 floats bias
constexpr uint32_t OFF_BAR = OFF_VEC + 1024;         // full[16], empty[16], accf[4], acce[4], tmem base
constexpr uint32_t SMEM_BYTES = OFF_BAR + 384;
constexpr uint32_t SMEM_ALLOC = SMEM_BYTES + 1024;
constexpr int NWORK = 16;
constexpr int NT = (NWORK + 3) * 32;                 // 608
constexpr int NACC = 4;                              // accumulator buffers (128 TMEM columns each)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok, tries = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  while (!ok) {
    __nanosleep(40);
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (++tries > (1u << 24)) __trap();   // a lost arrival must fail loudly, never hang the GPU
  }
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ uint4 pack8(const float* x) {
  __half2 a = __floats2half2_rn(x[0], x[1]), b = __floats2half2_rn(x[2], x[3]);
  __half2 c = __floats2half2_rn(x[4], x[5]), d = __floats2half2_rn(x[6], x[7]);
  uint4 o;
  o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
  o.z = *reinterpret_cast<uint32_t*>(&c); o.w = *reinterpret_cast<uint32_t*>(&d);
  return o;
}
__device__ __forceinline__ float silu_tanh(float x) {
  float t;
  const float h = 0.5f * x;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

enum Mode { MODE_AB = 0, MODE_Z = 1, MODE_H = 2 };

#ifndef NTT_Z_ROWS
#define NTT_Z_ROWS 128    // rows per MODE_Z tile (128 or 256)
#endif
#ifndef NTT_AB_ROWS
#define NTT_AB_ROWS 128   // rows per MODE_AB tile (64 or 128)
#endif
#ifndef NTT_AB_ISSUERS
#define NTT_AB_ISSUERS 2
#endif
#ifndef NTT_BULK_W
#define NTT_BULK_W 1   // weight image by cp.async.bulk (TMA 1-D) overlapped with the set-up instead of 14 rounds of LDG + STS
#endif
#ifndef NTT_TIMING
#define NTT_TIMING 0   // 1: per-role wait cycles (diagnostic builds), printed every 16 launches of a mode
#endif
#if NTT_TIMING
#define NTWAIT(acc, call) do { const long long _t0 = clock64(); call; acc += (unsigned long long)(clock64() - _t0); } while (0)
#else
#define NTWAIT(acc, call) do { call; } while (0)
#endif

struct Params {
  unsigned long long* timing;   // NTT_TIMING: [8] {prologue, mma wait full, mma wait acce, mma total, loader wait empty, loader total, worker wait accf, worker total}
  int M, ntiles, N;
  // MODE_AB: X = h16; CTAs [0, grid/2) use W0/bias0/out0, the rest W1/(no bias)/out1; out = fp16(0.5 * (acc + bias))
  // MODE_Z : K blocks 0-3 from X = h16, 4-7 from X2 = agg16; CTAs [0, grid/2) use W0 (output features 0-127), the
  //          rest W1 (features 128-255); out32[:, half] = acc + bias0[half]
  // MODE_H : operand = fp16(SiLU(z * gscale[b] + gshift[b])) with W0; h = h + acc + bias0; also h16
  const __half* X;
  const __half* X2;
  const __half* W0;
  const __half* W1;
  const float* bias0;
  __half* out0;
  __half* out1;
  float* out32;
  const float* z;
  const float* gscale;   // [B, 256]
  const float* gshift;   // [B, 256]
  float* h;
  __half* h16;
};

template <int MODE>
__global__ void __launch_bounds__(NT, 1) k_nodeT(const Params p, const __grid_constant__ CUtensorMap tmX,
                                                 const __grid_constant__ CUtensorMap tmX2) {
  // rows per tile = N of the MMA.  The issue path costs ~130 cycles per tcgen05.mma whatever its N (wait counters,
  // NTT_TIMING: with N = 64 the issuer was busy 87 % of MODE_AB while the tensor pipe was 16 % active), so MODE_AB uses
  // 128-row tiles like MODE_Z; MODE_H stays at 64 rows (its workers build the operand: not issue-bound)
  constexpr int ROWS = (MODE == MODE_Z) ? NTT_Z_ROWS : (MODE == MODE_AB && NTT_AB_ROWS == 128) ? 128 : 64;
  constexpr int KB = (MODE == MODE_Z) ? 8 : 4;             // K blocks (64 wide) per tile
  constexpr int NHALF = (MODE == MODE_Z) ? 1 : 2;          // 128-feature halves computed per tile
  constexpr uint32_t SLOT_BYTES = ROWS * 128;              // one K block of the activation tile
  constexpr int NSLOT = RING_BYTES / SLOT_BYTES;           // 8 or 4
  constexpr uint32_t W_KBLK = (MODE == MODE_Z ? 128 : 256) * 128;   // bytes per K block of the weight image
  constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(ROWS >> 3) << 17) | ((128u >> 4) << 24);
  constexpr bool SPLIT = (MODE == MODE_AB || MODE == MODE_Z);
  constexpr int TCOLS = NHALF * ROWS;                      // TMEM columns of one accumulator buffer: 128 or 256
  constexpr int NACCM = 512 / TCOLS < NACC ? 512 / TCOLS : NACC;
  // MMA issuers: the issue path costs ~130-160 cycles per tcgen05.mma and MODE_AB's issuer was busy 88 % of the launch
  // (NTT_TIMING), so its two feature halves are issued by two threads (warps NWORK + 2 and NWORK + 1), each into its own
  // accumulator columns; an operand slot is free / a tile is ready when both have committed
  constexpr int NISSUE = (MODE == MODE_AB && NTT_AB_ISSUERS == 2) ? 2 : 1;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  float* vbias = reinterpret_cast<float*>(smem + OFF_VEC);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 336);
  const uint32_t bar_full = sbase + OFF_BAR, bar_empty = sbase + OFF_BAR + 128;
  const uint32_t bar_accf = sbase + OFF_BAR + 256, bar_acce = sbase + OFF_BAR + 288;
  const uint32_t bar_w = sbase + OFF_BAR + 352;
  static_assert(NSLOT <= 16, "barrier block holds 16 ring slots");
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long t_kernel = clock64();
  pdl_trigger();            // programmatic dependent launch, see common.cuh

  const int half_grid = (int)gridDim.x >> 1;
  const int side = (SPLIT && (int)blockIdx.x >= half_grid) ? 1 : 0;
  const int cta = SPLIT ? ((int)blockIdx.x - side * half_grid) : (int)blockIdx.x;
  const int ncta = SPLIT ? half_grid : (int)gridDim.x;

  {
#if !NTT_BULK_W
    const uint4* src = reinterpret_cast<const uint4*>(side ? p.W1 : p.W0);
    uint4* dst = reinterpret_cast<uint4*>(smem + OFF_W);
    for (int i = tid; i < (int)(W_BYTES / 16); i += NT) dst[i] = __ldg(src + i);
#endif
    if (tid < 256) {
      float b = p.bias0 ? p.bias0[tid] : 0.f;
      if (MODE == MODE_AB && side) b = 0.f;
      vbias[tid] = b;
    }
  }
  if (tid == 0) {
    for (int i = 0; i < NSLOT; ++i) { mbar_init(bar_full + 8 * i, MODE == MODE_H ? NWORK * 32 : 1); mbar_init(bar_empty + 8 * i, NISSUE); }
    if (MODE != MODE_H) { tma_prefetch_desc(&tmX); if (MODE == MODE_Z) tma_prefetch_desc(&tmX2); }
    for (int i = 0; i < NACCM; ++i) { mbar_init(bar_accf + 8 * i, NISSUE); mbar_init(bar_acce + 8 * i, NWORK * 32); }
#if NTT_BULK_W
    mbar_init(bar_w, 1);
#endif
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#if NTT_BULK_W
    // the 128 KB weight image arrives by four bulk copies (async proxy) while the CTA finishes its set-up and the loaders
    // already fetch the first tiles; only the MMA issuer waits for it
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_w), "r"(W_BYTES) : "memory");
    const char* wsrc = reinterpret_cast<const char*>(side ? p.W1 : p.W0);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(sbase + OFF_W + (uint32_t)i * (W_BYTES / 4)), "l"(wsrc + (size_t)i * (W_BYTES / 4)), "r"(W_BYTES / 4), "r"(bar_w) : "memory");
#endif
  }
  if (warp == NWORK + 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();               // activations written by the preceding kernels are complete and visible from here on
  const long long t_start = clock64();
  if (NTT_TIMING && tid == 0) atomicAdd(p.timing + 0, (unsigned long long)(t_start - t_kernel));
  unsigned long long tw0 = 0, tw1 = 0;

  if (warp == NWORK + 2 || (NISSUE == 2 && warp == NWORK + 1)) {
    // =================================== MMA ISSUER(S) ================================================
    const int hf0 = (NISSUE == 2) ? (warp == NWORK + 2 ? 0 : 1) : 0;          // first feature half of this issuer
    const int hf1 = (NISSUE == 2) ? hf0 + 1 : NHALF;
    if (lane == 0) {
      const uint64_t dW = make_desc(sbase + OFF_W);
      const uint64_t dS = make_desc(sbase + OFF_S);
      int it = 0;
      uint32_t c = 0;                                        // running K-block count -> ring slot / phase
#if NTT_BULK_W
      mbar_wait(bar_w, 0u);                                  // weight image landed
#endif
      for (int tile = cta; tile < p.ntiles; tile += ncta, ++it) {
        const int buf = it % NACCM;
        const int use = it / NACCM;
        if (use >= 1) NTWAIT(tw1, mbar_wait(bar_acce + 8 * buf, (uint32_t)((use - 1) & 1)));
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * TCOLS);
#pragma unroll 1
        for (int kb = 0; kb < KB; ++kb, ++c) {
          const uint32_t slot = c % NSLOT;
          NTWAIT(tw0, mbar_wait(bar_full + 8 * slot, (c / NSLOT) & 1u));
          tc_fence_after();
#pragma unroll
          for (int hf = 0; hf < NHALF; ++hf) {
            if (hf < hf0 || hf >= hf1) continue;
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              // A = weight rows [hf*128, +128) of K block kb; B = activation rows of ring slot `slot`
              const uint64_t da = dW + (uint64_t)(((uint32_t)kb * W_KBLK + (uint32_t)hf * 16384u + k4 * 32) >> 4);
              const uint64_t db = dS + (uint64_t)((slot * SLOT_BYTES + k4 * 32) >> 4);
              mma_f16(d_tmem + (uint32_t)(hf * ROWS), da, db, IDESC, (kb | k4) ? 1u : 0u);
            }
          }
          mma_commit(bar_empty + 8 * slot);
        }
        mma_commit(bar_accf + 8 * buf);
      }
      if (NTT_TIMING) { atomicAdd(p.timing + 1, tw0); atomicAdd(p.timing + 2, tw1); atomicAdd(p.timing + 3, (unsigned long long)(clock64() - t_start)); }
    }
    __syncwarp();
  } else if (warp >= NWORK) {
    // =================================== LOADER (TMA) =================================================
    if (MODE != MODE_H && warp == NWORK && lane == 0) {
      uint32_t c = 0;
      for (int tile = cta; tile < p.ntiles; tile += ncta) {
        const int row0 = tile * ROWS;
#pragma unroll 1
        for (int kb = 0; kb < KB; ++kb, ++c) {
          const uint32_t slot = c % NSLOT;
          if (c >= (uint32_t)NSLOT) NTWAIT(tw0, mbar_wait(bar_empty + 8 * slot, ((c / NSLOT) - 1) & 1u));
          mbar_expect_tx(bar_full + 8 * slot, SLOT_BYTES);
          tma_load_2d(sbase + OFF_S + slot * SLOT_BYTES, (MODE == MODE_Z && kb >= 4) ? &tmX2 : &tmX, (kb & 3) * 64, row0,
                      bar_full + 8 * slot);
        }
      }
      if (NTT_TIMING) { atomicAdd(p.timing + 4, tw0); atomicAdd(p.timing + 5, (unsigned long long)(clock64() - t_start)); }
    }
    __syncwarp();
  } else {
    // =================================== WORKERS ======================================================
    // warp (q, g): TMEM lanes q*32.. (features), accumulator columns g*32.. of the tile's 128:
    //   AB/H: g>>1 = feature half, (g&1)*32 = first tile row;  Z: g*32 = first tile row
    const int q = warp & 3, g = warp >> 2;
    const int fhalf = (MODE == MODE_Z) ? side : (g >> 1);
    const int feat = fhalf * 128 + q * 32 + lane;            // output feature of this thread (column of [M, 256])
    // row groups (32 rows) per warp: Z: one (g); AB/H: ROWS / 64 consecutive groups starting at (g & 1) * (ROWS / 64)
    constexpr int RGW = (MODE == MODE_Z) ? ROWS / 128 : ROWS / 64;
    const int rbase = (MODE == MODE_Z) ? g * RGW * 32 : (g & 1) * RGW * 32;
    const float bias = vbias[feat];
    uint32_t cb = 0;

    // MODE_H operand: y = SiLU(z * gscale[b] + gshift[b]) -> fp16, one K block (64 columns x 64 rows) at a time;
    // 8 lanes per row (8 columns each), 4 rows per warp: row = warp*4 + (lane >> 3)
    auto build_h = [&](int tile) {
      const int c8 = lane & 7, r = warp * 4 + (lane >> 3);
      const int m = tile * ROWS + r;
#pragma unroll 1
      for (int kb = 0; kb < 4; ++kb, ++cb) {
        const uint32_t slot = cb % NSLOT;
        const int col = kb * 64 + c8 * 8;
        float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (m < p.M) {
          const float4* zp = reinterpret_cast<const float4*>(p.z + (size_t)m * H + col);
          const float4 z0 = __ldg(zp), z1 = __ldg(zp + 1);
          const int b = m / p.N;
          const float4* sc = reinterpret_cast<const float4*>(p.gscale + (size_t)b * H + col);
          const float4* sh = reinterpret_cast<const float4*>(p.gshift + (size_t)b * H + col);
          const float4 s0 = __ldg(sc), s1 = __ldg(sc + 1), h0 = __ldg(sh), h1 = __ldg(sh + 1);
          x[0] = silu_tanh(fmaf(z0.x, s0.x, h0.x)); x[1] = silu_tanh(fmaf(z0.y, s0.y, h0.y));
          x[2] = silu_tanh(fmaf(z0.z, s0.z, h0.z)); x[3] = silu_tanh(fmaf(z0.w, s0.w, h0.w));
          x[4] = silu_tanh(fmaf(z1.x, s1.x, h1.x)); x[5] = silu_tanh(fmaf(z1.y, s1.y, h1.y));
          x[6] = silu_tanh(fmaf(z1.z, s1.z, h1.z)); x[7] = silu_tanh(fmaf(z1.w, s1.w, h1.w));
        }
        if (cb >= (uint32_t)NSLOT) mbar_wait(bar_empty + 8 * slot, ((cb / NSLOT) - 1) & 1u);
        *reinterpret_cast<uint4*>(smem + OFF_S + slot * SLOT_BYTES + (uint32_t)r * 128u + (uint32_t)((c8 ^ (r & 7)) << 4)) = pack8(x);
        fence_async_smem();
        mbar_arrive(bar_full + 8 * slot);
      }
    };

    auto epilogue = [&](int tile, int it) {
      const int buf = it % NACCM;
      NTWAIT(tw0, mbar_wait(bar_accf + 8 * buf, (uint32_t)((it / NACCM) & 1)));
      tc_fence_after();
      // accumulator column of (feature half hf, tile row r): buf * TCOLS + hf * ROWS + r
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) +
                             (uint32_t)(buf * TCOLS + (MODE == MODE_Z ? 0 : fhalf * ROWS) + rbase);
      float vv[RGW][32];
#pragma unroll
      for (int rr = 0; rr < RGW; ++rr) tmem_ld32_issue(taddr + (uint32_t)(rr * 32), vv[rr]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(bar_acce + 8 * buf);
#pragma unroll
      for (int rr = 0; rr < RGW; ++rr) {
      const float* v = vv[rr];
      const int row0 = tile * ROWS + rbase + rr * 32;
      if (MODE == MODE_AB) {
        __half* out = (side ? p.out1 : p.out0) + feat;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int m = row0 + e;
          if (m < p.M) out[(size_t)m * H] = __float2half_rn(0.5f * (v[e] + bias));
        }
      } else if (MODE == MODE_Z) {
        float* out = p.out32 + feat;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int m = row0 + e;
          if (m < p.M) out[(size_t)m * H] = v[e] + bias;
        }
      } else {
        float* hp = p.h + feat;
        __half* h16 = p.h16 + feat;
        float hv[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int m = row0 + e;
          hv[e] = (m < p.M) ? hp[(size_t)m * H] : 0.f;
        }
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int m = row0 + e;
          if (m < p.M) {
            const float o = hv[e] + v[e] + bias;
            hp[(size_t)m * H] = o;
            h16[(size_t)m * H] = __float2half_rn(o);
          }
        }
      }
      }
    };

    int it = 0;
    if (MODE == MODE_H) {
      // build(t) -> epilogue(t-1) -> build(t+1) ...: the MMA of tile t runs under the epilogue of tile t-1
      int prev = -1;
      for (int tile = cta; tile < p.ntiles; tile += ncta, ++it) {
        build_h(tile);
        if (prev >= 0) epilogue(prev, it - 1);
        prev = tile;
      }
      if (prev >= 0) epilogue(prev, it - 1);
    } else {
      for (int tile = cta; tile < p.ntiles; tile += ncta, ++it) epilogue(tile, it);
    }
    if (NTT_TIMING && tid == 0) { atomicAdd(p.timing + 6, tw0); atomicAdd(p.timing + 7, (unsigned long long)(clock64() - t_start)); }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NWORK + 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

template <int MODE>
static int launch(dfm_ctx* ctx, const Params& p, int grid, cudaStream_t s) {
  static unsigned long long attr_devices = 0;
  if (dfm_once_per_device(attr_devices, ctx->device)) {
    CUDA_TRY(cudaFuncSetAttribute(k_nodeT<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_ALLOC));
  }
  if (grid <= 0) return 0;
  // tensor maps of the fp16 activation matrices [M, 256] (MODE_H builds its operand from fp32 z: no map needed)
  constexpr int ROWS = (MODE == MODE_Z) ? NTT_Z_ROWS : (MODE == MODE_AB && NTT_AB_ROWS == 128) ? 128 : 64;
  CUtensorMap tmX, tmX2;
  memset(&tmX, 0, sizeof(tmX));
  memset(&tmX2, 0, sizeof(tmX2));
  if (MODE != MODE_H) {
    int rc = dfm_make_tmap_f16(&tmX, p.X, (uint64_t)p.M, H, ROWS);
    if (rc) return rc;
    if (MODE == MODE_Z && (rc = dfm_make_tmap_f16(&tmX2, p.X2, (uint64_t)p.M, H, ROWS))) return rc;
  }
#if NTT_TIMING
  static unsigned long long* tbuf = nullptr;
  static int calls = 0;
  if (!tbuf) { CUDA_TRY(cudaMalloc(&tbuf, 64)); CUDA_TRY(cudaMemset(tbuf, 0, 64)); }
  Params q = p; q.timing = tbuf;
  CUDA_TRY(dfm_launch_pdl(k_nodeT<MODE>, dim3(grid), dim3(NT), SMEM_ALLOC, s, q, tmX, tmX2));
  if (++calls % 16 == 0) {
    unsigned long long hb[8];
    cudaMemcpy(hb, tbuf, 64, cudaMemcpyDeviceToHost);
    fprintf(stderr, "[ntt timing mode %d, sums over %d CTAs x 16 launches, kcycles] prologue %.0f | mma wait full %.0f acce %.0f of %.0f | loader wait empty %.0f of %.0f | worker0 wait accf %.0f of %.0f\n",
            MODE, grid, hb[0] * 1e-3, hb[1] * 1e-3, hb[2] * 1e-3, hb[3] * 1e-3, hb[4] * 1e-3, hb[5] * 1e-3, hb[6] * 1e-3, hb[7] * 1e-3);
    cudaMemset(tbuf, 0, 64);
  }
  LAUNCH_CHECK(ctx);
  return 0;
#endif
  CUDA_TRY(dfm_launch_pdl(k_nodeT<MODE>, dim3(grid), dim3(NT), SMEM_ALLOC, s, p, tmX, tmX2));
  LAUNCH_CHECK(ctx);
  return 0;
}

}  // namespace ntt

// Ah = fp16((W1s h + b1eff)/2) and Bm = fp16((W1d h)/2) in one launch (two halves of the grid)
int launch_node_ab(dfm_ctx* ctx, int layer, int M, const __half* h16, __half* Ah, __half* Bm, cudaStream_t s) {
  const LayerW& w = ctx->layer[layer];
  ntt::Params p{};
  p.M = M; p.ntiles = (M + NTT_AB_ROWS - 1) / NTT_AB_ROWS; p.N = ctx->N;
  p.X = h16; p.W0 = w.img_W1s; p.W1 = w.img_W1d; p.bias0 = w.b1eff; p.out0 = Ah; p.out1 = Bm;
  int half = ctx->num_sms / 2;
  if (half > p.ntiles) half = p.ntiles;
  return ntt::launch<ntt::MODE_AB>(ctx, p, 2 * half, s);
}

// z = W3h h + W3a agg + b3 (agg16 carries agg x 2^-6, the image carries W3a x 2^6)
int launch_node_z(dfm_ctx* ctx, int layer, int M, const __half* h16, const __half* agg16, float* z, cudaStream_t s) {
  const LayerW& w = ctx->layer[layer];
  ntt::Params p{};
  p.M = M; p.ntiles = (M + NTT_Z_ROWS - 1) / NTT_Z_ROWS; p.N = ctx->N;
  p.X = h16; p.X2 = agg16; p.W0 = w.img_W3z0; p.W1 = w.img_W3z1; p.bias0 = w.b3; p.out32 = z;
  int half = ctx->num_sms / 2;
  if (half > p.ntiles) half = p.ntiles;
  return ntt::launch<ntt::MODE_Z>(ctx, p, 2 * half, s);
}

// h += W4 SiLU(z * gscale + gshift) + b4; h16 = fp16(h)
int launch_node_h(dfm_ctx* ctx, int layer, int M, const float* z, const float* gscale, const float* gshift, float* h,
                  __half* h16, cudaStream_t s) {
  const LayerW& w = ctx->layer[layer];
  ntt::Params p{};
  p.M = M; p.ntiles = (M + 63) / 64; p.N = ctx->N;
  p.W0 = w.img_W4; p.bias0 = w.b4; p.z = z; p.gscale = gscale; p.gshift = gshift; p.h = h; p.h16 = h16;
  return ntt::launch<ntt::MODE_H>(ctx, p, p.ntiles < ctx->num_sms ? p.ntiles : ctx->num_sms, s);
}
