// Warp-specialised tcgen05 edge kernel (throughput path of the fused edge MLP):
//   m = SiLU(W2 SiLU(u) + b2), gate, per-residue segment sum              (src/models/egnn.py:95-116, 139-148)
//   u = A_i + B_j + radial * w1r + T[d, relpos] + T[omega, theta, phi]     (SURVEY App. A.5 / A.7 decomposition)
//
// One persistent CTA per SM, 28 warps (25 active):
//   warps 0-7   producers: gather B_j / table rows (fp16, L2), form u/2 in packed half2, SiLU via tanh.approx.f16x2,
//               write the 128 x 256 fp16 operand tile S into shared memory one 64-column K block at a time
//   warp  24    MMA issuer: D[128 x 256] (TMEM, fp32) = S * (W2/2)^T, one tcgen05.commit per K block (frees that
//               block for the next tile's build) and one per tile (accumulator ready)
//   warps 8-23  epilogue: TMEM -> registers, + b2/2, SiLU in half2, gate logit, gate, 32-row column sums by
//               shuffle transposition, segment sum of the two residues of the tile -> agg (fp32)
// The fp16 weight image (128 KB) stays resident in shared memory; accumulators are double buffered in TMEM
// (2 x 256 columns) so that build(t+1), MMA(t) and epilogue(t-1) overlap.
//
// All operands are pre-halved at production (A, B, tables, w1r, W2, b2 carry a factor 1/2), because
// SiLU(x) = h + h * tanh(h) with h = x/2: one MUFU and one HFMA2 per element pair.
#include <stdlib.h>

#include "common.cuh"

namespace ews {

constexpr int TILE_M = 128;
constexpr uint32_t W_BYTES = 256 * 256 * 2;          // 131072
constexpr uint32_t S_BYTES = TILE_M * 256 * 2;       // 65536
constexpr uint32_t W_KBLK = 256 * 128;               // bytes per 64-wide K block of the weight image
constexpr uint32_t S_KBLK = TILE_M * 128;
constexpr uint32_t OFF_W = 0;
constexpr uint32_t OFF_S = W_BYTES;
constexpr uint32_t OFF_VEC = OFF_S + S_BYTES;        // b2/2 [256] half, wa [256] half, w1r' [256] half, (spare 512 B)
constexpr uint32_t OFF_PART = OFF_VEC + 2048;        // [4 column quarters][128 rows] float gate partials
constexpr uint32_t OFF_AGG = OFF_PART + 4 * 128 * 4; // [2 tile parities][4 lane quarters][256] float column sums
constexpr uint32_t OFF_META = OFF_AGG + 2 * 4 * 256 * 4; // [8 producer warps][2 slots][16 rows] int4 edge metadata
constexpr uint32_t OFF_BAR = OFF_META + 8 * 2 * 16 * 16;  // 12 mbarriers + tmem base
constexpr uint32_t SMEM_BYTES = OFF_BAR + 128;
constexpr uint32_t SMEM_ALLOC = SMEM_BYTES + 1024;   // slack for the manual 1024-byte alignment

constexpr int NPROD = 8;                 // producer warps
constexpr int NEPI = 16;                 // epilogue warps
// 28 warps: 8 producers, 16 epilogue, 1 MMA issuer + 3 idle warps that only complete its warpgroup (setmaxnreg is a
// warpgroup-wide operation: a lone 17th warp never finishes it and the epilogue's .inc then blocks forever).
constexpr int NT = (NPROD + NEPI + 4) * 32;   // 896 -> 72 registers/thread at launch, pool 28*32*72 = 64512
constexpr int PROD_REGS = 104;           // setmaxnreg: 8*32*104 + 16*32*64 + 4*32*24 = 62464 <= 64512
constexpr int EPI_REGS = 64;
constexpr int MMA_REGS = 24;
#ifndef EWS_USE_ALO
#define EWS_USE_ALO 0     // carry A_i as fp16 hi + lo (1) or a single fp16 (0)
#endif

// kind::f16 instruction descriptor: D=f32 (bit 4), A=B=f16 (0), both K-major, N=256 (>>3 at bit 17), M=128 (>>4 at bit 24)
constexpr uint32_t IDESC = (1u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  // K-major, SWIZZLE_128B: start>>4 | LBO(ignored)=1 | SBO = 1024 B (8 rows x 128 B) | version 1 | layout 2
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t gtimer() {
  uint64_t t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  uint64_t t0 = 0;
  do {
    // the suspend-time hint parks the warp inside try_wait instead of spinning through the issue slots
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
    if (!ok) {   // a lost arrival must fail loudly, never hang the GPU
      const uint64_t t = gtimer();
      if (t0 == 0) t0 = t;
      else if (t - t0 > 4000000000ull) __trap();
    }
  } while (!ok);
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t* r) {
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
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- packed half2 helpers on raw 32-bit registers ---------------------------------------------------
__device__ __forceinline__ uint32_t h2add(uint32_t a, uint32_t b) {
  uint32_t d; asm("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d;
}
__device__ __forceinline__ uint32_t h2mul(uint32_t a, uint32_t b) {
  uint32_t d; asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d;
}
__device__ __forceinline__ uint32_t h2fma(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d; asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}
__device__ __forceinline__ uint32_t h2tanh(uint32_t a) {
  uint32_t d; asm("tanh.approx.f16x2 %0, %1;" : "=r"(d) : "r"(a)); return d;
}
// (lo, hi) fp32 -> packed f16x2, round to nearest, saturating to the finite range
__device__ __forceinline__ uint32_t f2h2(float lo, float hi) {
  uint32_t d; asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo)); return d;
}
__device__ __forceinline__ float2 h2f2(uint32_t a) {
  return __half22float2(*reinterpret_cast<const __half2*>(&a));
}
// SiLU(2h) = h + h * tanh(h)
__device__ __forceinline__ uint32_t h2silu(uint32_t h) { return h2fma(h, h2tanh(h), h); }

struct Params {
  int ntiles;
  int total_nodes;          // B * N
  int N, R, K;
  int last;                 // spill gated messages of ligand residues (coordinate head input)
  const __half* Wimg;       // (W2 / 2) fp16 SW128 image
  const int4* emeta;        // [B*N, 64] {global row of j, Tdrp row, Totp row or -1, radial bits}
  const __half* Ahi;        // [B*N, 256] fp16((W1s h_i + b1)/2)
  const __half* Alo;        // [B*N, 256] residual of the above
  const __half* Bm;         // [B*N, 256] fp16((W1d h_j)/2)
  const __half* Tdrp;       // pre-halved merged tables
  const __half* Totp;
  const float* w1r;         // [256] fp32 (unhalved)
  const float* b2;          // [256]
  const float* wa;          // [256]
  const float* ba;          // [1]
  float* agg;               // [B*N, 256] fp32 out
  __half* mstar;            // [B*L, 64, 256] fp16 (m* x 2^-6), last layer only
};

constexpr float RAD_SCALE = 0.03125f;     // radial is carried as fp16(radial / 32); w1r' = 32 * w1r / 2
constexpr float MSTAR_SCALE = 0.015625f;  // gated messages are carried x 2^-6 in fp16 (column sums stay < 65504)

// 32 lanes x NR packed registers -> lane l ends with the lane-sums of registers 2l and 2l+1 in v[0], v[1]
template <int NR>
__device__ __forceinline__ void lane_transpose_sum_h2(uint32_t* v, int lane) {
#pragma unroll
  for (int o = 16, n = NR; o >= 1; o >>= 1, n >>= 1) {
    const bool up = (lane & o) != 0;
    const int half = n >> 1;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const uint32_t send = up ? v[i] : v[i + half];
      const uint32_t keep = up ? v[i + half] : v[i];
      v[i] = h2add(keep, __shfl_xor_sync(0xffffffffu, send, o));
    }
  }
}

__global__ void __launch_bounds__(NT, 1) k_edge_ws(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  __half* vb2 = reinterpret_cast<__half*>(smem + OFF_VEC);         // b2/2
  __half* vwa = vb2 + 256;
  __half* vwr = vb2 + 512;                                          // 16 * w1r
  float* part = reinterpret_cast<float*>(smem + OFF_PART);
  float* aggp = reinterpret_cast<float*>(smem + OFF_AGG);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 112);
  const uint32_t bar_full = sbase + OFF_BAR;            // [4]
  const uint32_t bar_empty = sbase + OFF_BAR + 32;      // [4]
  const uint32_t bar_accf = sbase + OFF_BAR + 64;       // [2]
  const uint32_t bar_acce = sbase + OFF_BAR + 80;       // [2]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time setup ------------------------------------------------------------------------------
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.Wimg);
    uint4* dst = reinterpret_cast<uint4*>(smem + OFF_W);
    for (int i = tid; i < (int)(W_BYTES / 16); i += NT) dst[i] = __ldg(src + i);
    if (tid < 256) {
      vb2[tid] = __float2half_rn(0.5f * p.b2[tid]);
      vwa[tid] = __float2half_rn(p.wa[tid]);
      vwr[tid] = __float2half_rn(p.w1r[tid] * (0.5f / RAD_SCALE));
    }
  }
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) { mbar_init(bar_full + 8 * i, NPROD * 32); mbar_init(bar_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_accf + 8 * i, 1); mbar_init(bar_acce + 8 * i, NEPI * 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == NPROD + NEPI) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < NPROD) {
    // =================================== PRODUCERS ===================================================
    // rows of this lane: r_i = warp*16 + 4 i + (lane >> 3), i = 0..3 (all in residue `warp >> 2` of the tile);
    // 16-byte chunk c8 = lane & 7 of each 128-byte K-block row.  Work item = (K block, pair of rows); the gathers of
    // item n+1 are in flight while item n is computed (two register buffers), across K blocks and across tiles.
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(PROD_REGS));
    const int c8 = lane & 7, rsub = lane >> 3;
    const int r0 = warp * 16 + rsub;
    int4* mring = reinterpret_cast<int4*>(smem + OFF_META) + warp * 32;   // [2 slots][16 rows] private to this warp
    const int4 pad_meta = make_int4(0, 40 * 66 + 32, -1, 0);
    struct GBuf { uint4 hb[2], td[2], to[2]; uint32_t rad[2]; };
    auto issue = [&](GBuf& g, const int4* mslot, int kb, int beta) {
      const int colh = kb * 64 + c8 * 8;
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
        const int4 mt = mslot[4 * (2 * beta + s2) + rsub];
        g.hb[s2] = __ldg(reinterpret_cast<const uint4*>(p.Bm + (size_t)mt.x * H + colh));
        g.td[s2] = __ldg(reinterpret_cast<const uint4*>(p.Tdrp + (size_t)mt.y * H + colh));
        g.to[s2] = make_uint4(0, 0, 0, 0);
        if (mt.z >= 0) g.to[s2] = __ldg(reinterpret_cast<const uint4*>(p.Totp + (size_t)mt.z * H + colh));
        const float rs = fminf(__int_as_float(mt.w) * RAD_SCALE, 65000.f);
        g.rad[s2] = f2h2(rs, rs);
      }
    };
    auto compute = [&](const GBuf& g, const uint4& ahi, const uint4& alo, int kb, int beta) {
      const uint4 wr = *reinterpret_cast<const uint4*>(vwr + kb * 64 + c8 * 8);
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
        const int r = r0 + 4 * (2 * beta + s2);
        uint4 o;
        uint32_t q;
        q = h2fma(g.rad[s2], wr.x, h2add(g.td[s2].x, g.to[s2].x)); o.x = h2silu(EWS_USE_ALO ? h2add(h2add(h2add(ahi.x, g.hb[s2].x), q), alo.x) : h2add(h2add(ahi.x, g.hb[s2].x), q));
        q = h2fma(g.rad[s2], wr.y, h2add(g.td[s2].y, g.to[s2].y)); o.y = h2silu(EWS_USE_ALO ? h2add(h2add(h2add(ahi.y, g.hb[s2].y), q), alo.y) : h2add(h2add(ahi.y, g.hb[s2].y), q));
        q = h2fma(g.rad[s2], wr.z, h2add(g.td[s2].z, g.to[s2].z)); o.z = h2silu(EWS_USE_ALO ? h2add(h2add(h2add(ahi.z, g.hb[s2].z), q), alo.z) : h2add(h2add(ahi.z, g.hb[s2].z), q));
        q = h2fma(g.rad[s2], wr.w, h2add(g.td[s2].w, g.to[s2].w)); o.w = h2silu(EWS_USE_ALO ? h2add(h2add(h2add(ahi.w, g.hb[s2].w), q), alo.w) : h2add(h2add(ahi.w, g.hb[s2].w), q));
        *reinterpret_cast<uint4*>(smem + OFF_S + (uint32_t)kb * S_KBLK + (uint32_t)r * 128u + (uint32_t)((c8 ^ (r & 7)) << 4)) = o;
      }
    };
    auto load_meta = [&](int tile) -> int4 {     // lanes 0..15: edge metadata of row warp*16 + lane of `tile`
      int4 mt = pad_meta;
      const int r = warp * 16 + (lane & 15);
      const int node = tile * 2 + (r >> 6);
      if (tile < p.ntiles && node < p.total_nodes) mt = __ldg(p.emeta + (size_t)node * SLOTS + (r & 63));
      return mt;
    };
    auto a_node = [&](int tile) -> size_t {
      int node = tile * 2 + (warp >> 2);
      if (node >= p.total_nodes) node = p.total_nodes - 1;   // odd tail: those rows are masked in the epilogue
      return (size_t)node * H + c8 * 8;
    };
    GBuf g0, g1;
    uint4 ahi, alo = make_uint4(0, 0, 0, 0), ahn, aln = make_uint4(0, 0, 0, 0);
    if ((int)blockIdx.x < p.ntiles) {
      const int4 m0 = load_meta(blockIdx.x);
      if (lane < 16) mring[lane] = m0;
      __syncwarp();
      issue(g0, mring, 0, 0);
      const size_t ao = a_node(blockIdx.x);
      ahi = __ldg(reinterpret_cast<const uint4*>(p.Ahi + ao));
      if (EWS_USE_ALO) alo = __ldg(reinterpret_cast<const uint4*>(p.Alo + ao));
    }
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int4* mcur = mring + (it & 1) * 16;
      int4* mnext = mring + ((it & 1) ^ 1) * 16;
      const int ntile = tile + (int)gridDim.x;
      const bool has_next = ntile < p.ntiles;
      const int4 nm = load_meta(ntile);
      const size_t ao = a_node(tile), aon = a_node(has_next ? ntile : tile);
#pragma unroll 1
      for (int kb = 0; kb < 4; ++kb) {
        issue(g1, mcur, kb, 1);
        if (it > 0) mbar_wait(bar_empty + 8 * kb, (uint32_t)((it - 1) & 1));   // MMA of the previous tile has read this block
        compute(g0, ahi, alo, kb, 0);
        if (kb == 1) {
          if (lane < 16) mnext[lane] = nm;
          __syncwarp();
        }
        if (kb < 3) {
          issue(g0, mcur, kb + 1, 0);
          ahn = __ldg(reinterpret_cast<const uint4*>(p.Ahi + ao + (kb + 1) * 64));
          if (EWS_USE_ALO) aln = __ldg(reinterpret_cast<const uint4*>(p.Alo + ao + (kb + 1) * 64));
        } else if (has_next) {
          issue(g0, mnext, 0, 0);
          ahn = __ldg(reinterpret_cast<const uint4*>(p.Ahi + aon));
          if (EWS_USE_ALO) aln = __ldg(reinterpret_cast<const uint4*>(p.Alo + aon));
        }
        compute(g1, ahi, alo, kb, 1);
        fence_async_smem();
        mbar_arrive(bar_full + 8 * kb);
        ahi = ahn; alo = aln;
      }
    }
  } else if (warp >= NPROD + NEPI) {
    // =================================== MMA ISSUER ===================================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(MMA_REGS));
    if (warp == NPROD + NEPI && lane == 0) {
      const uint64_t dW = make_desc(sbase + OFF_W);
      const uint64_t dS = make_desc(sbase + OFF_S);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        if (it >= 2) mbar_wait(bar_acce + 8 * buf, (uint32_t)(((it >> 1) - 1) & 1));   // epilogue drained this buffer
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 256);
#pragma unroll 1
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(bar_full + 8 * kb, (uint32_t)(it & 1));
          tc_fence_after();
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const uint64_t da = dS + (uint64_t)((kb * S_KBLK + k4 * 32) >> 4);
            const uint64_t db = dW + (uint64_t)((kb * W_KBLK + k4 * 32) >> 4);
            mma_f16(d_tmem, da, db, (kb | k4) ? 1u : 0u);
          }
          mma_commit(bar_empty + 8 * kb);
        }
        mma_commit(bar_accf + 8 * buf);
      }
    }
    __syncwarp();
  } else {
    // =================================== EPILOGUE =====================================================
    // 16 warps = 4 TMEM lane quarters (rows) x 4 column quarters; warp (q, cq) owns rows q*32.. and columns cq*64..
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(EPI_REGS));
    const int e = warp - NPROD;
    const int q = warp & 3;            // TMEM lane quarter this warp may access (warp id % 4)
    const int cq = e >> 2;             // column quarter
    const int erow = q * 32 + lane;
    const int hn = q >> 1;             // residue of the tile this warp's rows belong to
    const int ecol = (cq * 2 + (q & 1)) * 32 + lane;   // column this thread writes in the final combine (0..255)
    const float ba = p.ba[0];
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const int node = tile * 2 + hn, k = erow & 63;
      const bool valid = node < p.total_nodes && k < p.K;
      mbar_wait(bar_accf + 8 * buf, (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 256 + cq * 64);
      uint32_t m[32];
      float dot = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t acc[16];
        tmem_ld16_issue(taddr + c * 16, acc);
        tmem_ld_wait();
        uint32_t d0 = 0, d1 = 0, d2 = 0, d3 = 0;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const uint4 bb = *reinterpret_cast<const uint4*>(vb2 + cq * 64 + c * 16 + g * 8);
          const uint4 ww = *reinterpret_cast<const uint4*>(vwa + cq * 64 + c * 16 + g * 8);
          const uint32_t x0 = h2silu(h2add(f2h2(__uint_as_float(acc[g * 8 + 0]), __uint_as_float(acc[g * 8 + 1])), bb.x));
          const uint32_t x1 = h2silu(h2add(f2h2(__uint_as_float(acc[g * 8 + 2]), __uint_as_float(acc[g * 8 + 3])), bb.y));
          const uint32_t x2 = h2silu(h2add(f2h2(__uint_as_float(acc[g * 8 + 4]), __uint_as_float(acc[g * 8 + 5])), bb.z));
          const uint32_t x3 = h2silu(h2add(f2h2(__uint_as_float(acc[g * 8 + 6]), __uint_as_float(acc[g * 8 + 7])), bb.w));
          m[c * 8 + g * 4 + 0] = x0; m[c * 8 + g * 4 + 1] = x1; m[c * 8 + g * 4 + 2] = x2; m[c * 8 + g * 4 + 3] = x3;
          d0 = h2fma(x0, ww.x, d0); d1 = h2fma(x1, ww.y, d1); d2 = h2fma(x2, ww.z, d2); d3 = h2fma(x3, ww.w, d3);
        }
        const float2 f0 = h2f2(d0), f1 = h2f2(d1), f2 = h2f2(d2), f3 = h2f2(d3);
        dot += ((f0.x + f0.y) + (f1.x + f1.y)) + ((f2.x + f2.y) + (f3.x + f3.y));
      }
      tc_fence_before();
      mbar_arrive(bar_acce + 8 * buf);          // accumulator buffer may be overwritten by tile it + 2
      part[cq * 128 + erow] = dot;
      asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");      // the 4 column quarters of this row quarter
      const float tot = (part[erow] + part[128 + erow]) + (part[256 + erow] + part[384 + erow]) + ba;
      const float g = valid ? MSTAR_SCALE * __fdividef(1.f, 1.f + __expf(-tot)) : 0.f;
      const uint32_t g2 = f2h2(g, g);
#pragma unroll
      for (int i = 0; i < 32; ++i) m[i] = h2mul(m[i], g2);
      if (p.last && valid) {
        const int b = node / p.N, i = node - b * p.N;
        if (i >= p.R) {
          uint4* dst = reinterpret_cast<uint4*>(p.mstar + (((size_t)b * (p.N - p.R) + (i - p.R)) * SLOTS + k) * H + cq * 64);
#pragma unroll
          for (int v4 = 0; v4 < 8; ++v4) dst[v4] = make_uint4(m[v4 * 4], m[v4 * 4 + 1], m[v4 * 4 + 2], m[v4 * 4 + 3]);
        }
      }
      lane_transpose_sum_h2<32>(m, lane);       // lane l: columns cq*64 + 2l, 2l+1 summed over this warp's 32 rows
      float* ag = aggp + buf * 1024;
      *reinterpret_cast<float2*>(ag + q * 256 + cq * 64 + lane * 2) = h2f2(m[0]);
      asm volatile("bar.sync %0, 256;" ::"r"(5 + hn) : "memory");     // the 8 warps that hold this residue's rows
      if (node < p.total_nodes)
        p.agg[(size_t)node * H + ecol] = (ag[(2 * hn) * 256 + ecol] + ag[(2 * hn + 1) * 256 + ecol]) * (1.f / MSTAR_SCALE);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NPROD + NEPI) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

}  // namespace ews

int launch_edge_ws(dfm_ctx* ctx, const EdgeArgs& a, const int4* emeta, const __half* Ahi, const __half* Alo,
                   cudaStream_t s) {
  const LayerW& w = ctx->layer[a.layer];
  ews::Params p{};
  p.total_nodes = a.B * a.N;
  p.ntiles = (p.total_nodes + 1) / 2;
  p.N = a.N; p.R = a.R; p.K = a.K;
  p.last = a.last ? 1 : 0;
  p.Wimg = w.img_W2h;
  p.emeta = emeta;
  p.Ahi = Ahi; p.Alo = Alo;
  p.Bm = reinterpret_cast<const __half*>(a.Bm);
  p.Tdrp = w.Tdrp16h; p.Totp = w.Totp16h;
  p.w1r = w.w1r; p.b2 = w.b2; p.wa = w.wa; p.ba = w.ba;
  p.agg = a.agg; p.mstar = a.mstar;
  static bool attr = false;
  if (!attr) {
    CUDA_TRY(cudaFuncSetAttribute(ews::k_edge_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ews::SMEM_ALLOC));
    attr = true;
  }
  const int grid = p.ntiles < ctx->num_sms ? p.ntiles : ctx->num_sms;
  if (grid <= 0) return 0;
  ews::k_edge_ws<<<grid, ews::NT, ews::SMEM_ALLOC, s>>>(p);
  LAUNCH_CHECK(ctx);
  return 0;
}
