// tcgen05 (5th-gen tensor core) kernels for the three 256x256 contractions on the hot path:
//   EDGE   m = SiLU(W2 SiLU(u) + b2), gate, per-residue segment sum        (src/models/egnn.py:95-116, 139-148)
//   COORD  w = clamp(wc2 . SiLU(Wc1 m* + bc1)), coordinate displacement     (src/models/egnn.py:118-137)
//   LINEAR out = add + A W^T + bias                                         (node_mlp, edge_mlp.0 node halves, to_energy)
//
// One persistent CTA per SM, 256 threads.  The fp16 weight image (128 KB, K-major SWIZZLE_128B) stays resident in
// shared memory; a 128-row activation tile (64 KB, same layout) is rebuilt per tile; accumulators are double
// buffered in TMEM (2 x 256 columns) so the MMA of tile t overlaps the epilogue of tile t-1.  fp16 operands,
// fp32 accumulation; operand scaling by exact powers of two keeps fp16 in range (common.cuh).
//
//   D[128 x 256] (TMEM, lane = row, column = feature) = S[128 x 256] (smem) * W[256 x 256]^T (smem)
#include <stdlib.h>

#include "common.cuh"

namespace tc {

constexpr int TILE_M = 128;
constexpr uint32_t W_BYTES = 256 * 256 * 2;          // 131072
constexpr uint32_t S_BYTES = TILE_M * 256 * 2;       // 65536
constexpr uint32_t W_KBLK = 256 * 128;               // bytes per 64-wide K block of the weight image
constexpr uint32_t S_KBLK = TILE_M * 128;
constexpr uint32_t OFF_W = 0;
constexpr uint32_t OFF_S = W_BYTES;
constexpr uint32_t OFF_VEC = OFF_S + S_BYTES;        // 4 x 256 floats of per-column parameters
constexpr uint32_t OFF_PART = OFF_VEC + 4 * 256 * 4; // [4][128] gate partials
constexpr uint32_t OFF_AGG = OFF_PART + 4 * 128 * 4; // [4][256] column partial sums
constexpr uint32_t OFF_META = OFF_AGG + 4 * 256 * 4; // [2 buffers][3][128] per-row edge metadata (neighbour, bins, radial)
constexpr uint32_t OFF_BAR = OFF_META + 2 * 3 * 128 * 4;  // 2 mbarriers + tmem base
constexpr uint32_t SMEM_BYTES = OFF_BAR + 64;
constexpr uint32_t SMEM_ALLOC = SMEM_BYTES + 1024;   // slack for the manual 1024-byte alignment

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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  uint32_t spins = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok && ++spins > (1u << 24)) __trap();   // a lost MMA completion must fail loudly, never hang the GPU
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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// issue only (no wait): 32 consecutive columns of this thread's TMEM lane into v[0..31]
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

// SiLU(x) * 2^-4 = h (1 + tanh(x/2)) with h = x/32: one MUFU (tanh.approx, rel. error 2^-11 -- below the fp16 operand rounding)
__device__ __forceinline__ float silu_scaled_tanh(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  const float hs = x * (0.5f * S_SCALE);
  return fmaf(hs, t, hs);
}
// packed variant: two SiLU(x) * 2^-4 results straight into the fp16 operand (tanh.approx.f16x2: one MUFU per pair)
__device__ __forceinline__ uint32_t silu_scaled_tanh_h2(float x0, float x1) {
  const __half2 xh = __floats2half2_rn(0.5f * x0, 0.5f * x1);
  uint32_t xi = *reinterpret_cast<const uint32_t*>(&xh), ti;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(ti) : "r"(xi));
  const __half2 t = *reinterpret_cast<const __half2*>(&ti);
  const __half2 hs = __hmul2(xh, __float2half2_rn(S_SCALE));   // (x/2) * 2^-4 = x/32
  const __half2 o = __hfma2(hs, t, hs);
  return *reinterpret_cast<const uint32_t*>(&o);
}
__device__ __forceinline__ float silu_tanh(float x) {
  float t;
  const float h = 0.5f * x;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
__device__ __forceinline__ float silu_fast(float x) { return __fdividef(x, 1.f + __expf(-x)); }

// 16-byte chunk `c16` (0..31) of tile row r -> byte offset inside the S tile (K-major SWIZZLE_128B)
__device__ __forceinline__ uint32_t s_off(int r, int c16) {
  return (uint32_t)(c16 >> 3) * S_KBLK + (uint32_t)r * 128u + (uint32_t)(((c16 & 7) ^ (r & 7)) << 4);
}

__device__ __forceinline__ uint4 pack8(const float (&x)[8]) {
  __half2 a = __floats2half2_rn(x[0], x[1]), b = __floats2half2_rn(x[2], x[3]);
  __half2 c = __floats2half2_rn(x[4], x[5]), d = __floats2half2_rn(x[6], x[7]);
  uint4 o;
  o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
  o.z = *reinterpret_cast<uint32_t*>(&c); o.w = *reinterpret_cast<uint32_t*>(&d);
  return o;
}
__device__ __forceinline__ void add_half8(float (&u)[8], uint4 h) {
  const __half2* p = reinterpret_cast<const __half2*>(&h);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float2 f = __half22float2(p[q]);
    u[2 * q] += f.x;
    u[2 * q + 1] += f.y;
  }
}

// 32 lanes x 32 values -> lane l ends with sum over lanes of v[l]   (31 shuffles)
__device__ __forceinline__ float lane_transpose_sum(float* v, int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = up ? v[i] : v[i + o];
      const float keep = up ? v[i + o] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}

enum Mode { LINEAR = 0, EDGE = 1, COORD = 2 };

struct Params {
  int mode;
  int ntiles;
  const __half* Wimg;
  // LINEAR
  LinearArgs lin;
  // EDGE / COORD
  EdgeArgs ed;
  const __half* Tdrp;   // [(z*40 + d)*66 + rp][256]: T_d + T_relpos (+ the three zero-angle rows when z = 1)
  const __half* Totp;   // [(o*24 + t)*12 + p][256]: T_omega + T_theta + T_phi
  const float* w1r;
  const float* v0;   // EDGE: b2    COORD: bc1   LINEAR: bias (or null)
  const float* v1;   // EDGE: wa    COORD: wc2
  const float* ba;   // EDGE: att bias
};

constexpr int NT = 512;            // threads per CTA: 16 warps = 4 TMEM lane quarters x 4 column quarters
constexpr int NWARP = NT / 32;
constexpr int CW = 256 / (NWARP / 4);   // accumulator columns per thread in the epilogue (64)

// VAR (EDGE only): bit 0 = build SiLU in packed half2 (less accurate, kept for experiments)
template <int MODE, int VAR = 0>
__global__ void __launch_bounds__(NT, 1) k_tc(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  float* vec0 = reinterpret_cast<float*>(smem + OFF_VEC);
  float* vec1 = vec0 + 256;
  float* vec2 = vec0 + 512;                                // EDGE: w1r
  float* part = reinterpret_cast<float*>(smem + OFF_PART); // [4][128] gate partials (one per column quarter)
  float* aggp = reinterpret_cast<float*>(smem + OFF_AGG);  // [4][256] column sums per lane quarter
  int* meta_j = reinterpret_cast<int*>(smem + OFF_META);
  uint32_t* meta_ft = reinterpret_cast<uint32_t*>(smem + OFF_META + 2 * 128 * 4);
  float* meta_rad = reinterpret_cast<float*>(smem + OFF_META + 4 * 128 * 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 16);
  const uint32_t bar0 = sbase + OFF_BAR, bar1 = sbase + OFF_BAR + 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time setup: weight image -> smem, parameters, barriers, TMEM
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.Wimg);
    uint4* dst = reinterpret_cast<uint4*>(smem + OFF_W);
#pragma unroll 4
    for (int i = tid; i < (int)(W_BYTES / 16); i += NT) dst[i] = __ldg(src + i);
    if (tid < 256) {
      vec0[tid] = p.v0 ? p.v0[tid] : 0.f;
      vec1[tid] = p.v1 ? p.v1[tid] : 0.f;
      vec2[tid] = (MODE == EDGE) ? p.w1r[tid] : 0.f;
    }
    if (MODE == EDGE && tid < 128) {   // metadata of this CTA's first tile
      const int node0 = (int)blockIdx.x * 2 + (tid >> 6);
      int j0 = 0; uint32_t f0 = 0; float r0 = 0.f;
      if (node0 < p.ed.B * p.ed.N) {
        const size_t eo = (size_t)node0 * SLOTS + (tid & 63);
        j0 = p.ed.nbr[eo]; f0 = p.ed.feat[eo]; r0 = p.ed.radial[eo];
      }
      meta_j[tid] = j0; meta_ft[tid] = f0; meta_rad[tid] = r0;
    }
  }
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int q = warp & 3, cq = warp >> 2;       // epilogue: TMEM lane quarter / column quarter
  const int erow = q * 32 + lane;               // tile row owned in the epilogue

  const EdgeArgs& ed = p.ed;
  const int total_nodes = (MODE == EDGE) ? ed.B * ed.N : (MODE == COORD ? ed.B * (ed.N - ed.R) : 0);

  // ---------------------------------------------------------------------------------------------
  auto build = [&](int tile, int it) {
    if (MODE == LINEAR) {
#pragma unroll 4
      for (int r = warp; r < TILE_M; r += NWARP) {
        const int m = tile * TILE_M + r;
        float x[8];
        if (m < p.lin.M) {
          const float4* a = reinterpret_cast<const float4*>(p.lin.A + (size_t)m * H + lane * 8);
          float4 a0 = __ldg(a), a1 = __ldg(a + 1);
          x[0] = a0.x; x[1] = a0.y; x[2] = a0.z; x[3] = a0.w; x[4] = a1.x; x[5] = a1.y; x[6] = a1.z; x[7] = a1.w;
#pragma unroll
          for (int e = 0; e < 8; ++e) x[e] *= p.lin.a_scale;
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) x[e] = 0.f;
        }
        *reinterpret_cast<uint4*>(smem + OFF_S + s_off(r, lane)) = pack8(x);
      }
    } else if (MODE == COORD) {
      // rows are contiguous fp16 in mstar: node pair `tile`, 64 slots each
#pragma unroll 4
      for (int r = warp; r < TILE_M; r += NWARP) {
        const int node = tile * 2 + (r >> 6);
        uint4 v = make_uint4(0, 0, 0, 0);
        if (node < total_nodes)
          v = __ldg(reinterpret_cast<const uint4*>(ed.mstar + ((size_t)node * SLOTS + (r & 63)) * H) + lane);
        *reinterpret_cast<uint4*>(smem + OFF_S + s_off(r, lane)) = v;
      }
    } else {
      // EDGE: S = SiLU(A_i + B_j + radial*w1r + T_drp[d,rp] (+ T_otp[o,t,p])) * 2^-4, one warp per row, 8 columns per lane.
      // Rows of this warp: r = warp + 16 q, q = 0..7 (q >> 2 = residue inside the tile).  The gathers of four rows are
      // issued together (12 independent 16-byte loads per lane) before any of them is consumed.
      const int mb = it & 1;
      const int* mj = meta_j + mb * 128;
      const uint32_t* mft = meta_ft + mb * 128;
      const float* mrad = meta_rad + mb * 128;
      // prefetch the next tile's metadata (registers now, shared memory at the end of this build)
      int nj = 0; uint32_t nft = 0; float nrad = 0.f;
      {
        const int ntile = tile + (int)gridDim.x;
        const int nnode = ntile * 2 + (tid >> 6);
        if (tid < 128 && ntile < p.ntiles && nnode < total_nodes) {
          const size_t eo = (size_t)nnode * SLOTS + (tid & 63);
          nj = __ldg(ed.nbr + eo); nft = __ldg(ed.feat + eo); nrad = __ldg(ed.radial + eo);
        }
      }
      float wr[8];
      {
        const float4 w0 = *reinterpret_cast<const float4*>(vec2 + lane * 8), w1 = *reinterpret_cast<const float4*>(vec2 + lane * 8 + 4);
        wr[0] = w0.x; wr[1] = w0.y; wr[2] = w0.z; wr[3] = w0.w; wr[4] = w1.x; wr[5] = w1.y; wr[6] = w1.z; wr[7] = w1.w;
      }
      const __half* Bm = reinterpret_cast<const __half*>(ed.Bm);
#pragma unroll
      for (int hn = 0; hn < 2; ++hn) {
        const int node = tile * 2 + hn;
        const bool nvalid = node < total_nodes;
        const size_t brow = nvalid ? (size_t)(node / ed.N) * ed.N : 0;
        float ai[8];
        if (nvalid) {
          const float4* a = reinterpret_cast<const float4*>(ed.A + (size_t)node * H + lane * 8);
          const float4 a0 = __ldg(a), a1 = __ldg(a + 1);
          ai[0] = a0.x; ai[1] = a0.y; ai[2] = a0.z; ai[3] = a0.w; ai[4] = a1.x; ai[5] = a1.y; ai[6] = a1.z; ai[7] = a1.w;
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) ai[e] = 0.f;
        }
        uint4 hb[4], td[4], to[4];
        float rad[4];
        bool val[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = warp + NWARP * (hn * 4 + i);
          const bool v = nvalid && (r & 63) < ed.K;
          const int j = v ? mj[r] : 0;
          const uint32_t ft = v ? mft[r] : 0u;
          const uint32_t otp = (ft >> 6) & 0x3FFFu;
          const uint32_t drp = ((otp == 0 ? 40u : 0u) + (ft & 63u)) * 66u + ((ft >> 20) & 127u);
          const uint32_t oidx = (((ft >> 6) & 31u) * 24u + ((ft >> 11) & 31u)) * 12u + ((ft >> 16) & 15u);
          val[i] = v;
          rad[i] = mrad[r];
          hb[i] = __ldg(reinterpret_cast<const uint4*>(Bm + (brow + j) * H) + lane);
          td[i] = __ldg(reinterpret_cast<const uint4*>(p.Tdrp + (size_t)drp * H) + lane);
          to[i] = make_uint4(0, 0, 0, 0);
          if (otp != 0) to[i] = __ldg(reinterpret_cast<const uint4*>(p.Totp + (size_t)oidx * H) + lane);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = warp + NWARP * (hn * 4 + i);
          float u[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) u[e] = fmaf(rad[i], wr[e], ai[e]);
          if (VAR & 1) {
            const __half2* hbp = reinterpret_cast<const __half2*>(&hb[i]);
            const __half2* tdp = reinterpret_cast<const __half2*>(&td[i]);
            const __half2* top = reinterpret_cast<const __half2*>(&to[i]);
            uint32_t o4[4];
#pragma unroll
            for (int q2 = 0; q2 < 4; ++q2) {
              const float2 g = __half22float2(__hadd2(__hadd2(hbp[q2], tdp[q2]), top[q2]));
              o4[q2] = val[i] ? silu_scaled_tanh_h2(u[2 * q2] + g.x, u[2 * q2 + 1] + g.y) : 0u;
            }
            *reinterpret_cast<uint4*>(smem + OFF_S + s_off(r, lane)) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
          } else {
            add_half8(u, hb[i]);
            add_half8(u, td[i]);
            add_half8(u, to[i]);
#pragma unroll
            for (int e = 0; e < 8; ++e) u[e] = val[i] ? silu_scaled_tanh(u[e]) : 0.f;
            *reinterpret_cast<uint4*>(smem + OFF_S + s_off(r, lane)) = pack8(u);
          }
        }
      }
      if (tid < 128) {
        meta_j[(mb ^ 1) * 128 + tid] = nj;
        meta_ft[(mb ^ 1) * 128 + tid] = nft;
        meta_rad[(mb ^ 1) * 128 + tid] = nrad;
      }
    }
  };

  // ---------------------------------------------------------------------------------------------
  auto epilogue = [&](int tile, int buf) {
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 256 + cq * CW);
    float m[CW];
#pragma unroll
    for (int c = 0; c < CW / 32; ++c) tmem_ld32_issue(taddr + c * 32, m + c * 32);
    tmem_ld_wait();
    tc_fence_before();
    if (MODE == LINEAR) {
      const int mrow = tile * TILE_M + erow;
      if (mrow < p.lin.M) {
        const int col0 = cq * CW;
        const size_t o = (size_t)mrow * H + col0;
#pragma unroll
        for (int e = 0; e < CW; ++e) m[e] += vec0[col0 + e];
        if (p.lin.add) {
#pragma unroll
          for (int e4 = 0; e4 < CW / 4; ++e4) {
            const float4 ad = *reinterpret_cast<const float4*>(p.lin.add + o + e4 * 4);
            m[e4 * 4] += ad.x; m[e4 * 4 + 1] += ad.y; m[e4 * 4 + 2] += ad.z; m[e4 * 4 + 3] += ad.w;
          }
        }
        if (p.lin.out) {
#pragma unroll
          for (int e4 = 0; e4 < CW / 4; ++e4)
            *reinterpret_cast<float4*>(p.lin.out + o + e4 * 4) = make_float4(m[e4 * 4], m[e4 * 4 + 1], m[e4 * 4 + 2], m[e4 * 4 + 3]);
        }
        if (p.lin.out16) {
          const float osc = p.lin.out_scale != 0.f ? p.lin.out_scale : 1.f;
#pragma unroll
          for (int e8 = 0; e8 < CW / 8; ++e8) {
            float x8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) x8[e] = m[e8 * 8 + e] * osc;
            const uint4 hi = pack8(x8);
            *reinterpret_cast<uint4*>(p.lin.out16 + o + e8 * 8) = hi;
            if (p.lin.out16_lo) {
              const __half2* hp = reinterpret_cast<const __half2*>(&hi);
#pragma unroll
              for (int q2 = 0; q2 < 4; ++q2) {
                const float2 f = __half22float2(hp[q2]);
                x8[2 * q2] -= f.x;
                x8[2 * q2 + 1] -= f.y;
              }
              *reinterpret_cast<uint4*>(p.lin.out16_lo + o + e8 * 8) = pack8(x8);
            }
          }
        }
      }
      return;
    }
    const int hn = erow >> 6, k = erow & 63;
    const int node = tile * 2 + hn;
    const bool valid = node < total_nodes && k < ed.K;
    float dotp = 0.f;
#pragma unroll
    for (int e = 0; e < CW; ++e) {
      const float x = silu_tanh(m[e] + vec0[cq * CW + e]);
      m[e] = x;
      dotp = fmaf(x, vec1[cq * CW + e], dotp);
    }
    part[cq * 128 + erow] = dotp;
    __syncthreads();
    const float tot = part[erow] + part[128 + erow] + part[256 + erow] + part[384 + erow];
    if (MODE == COORD) {
      // coordinate displacement of ligand residue `node` (index over B*L): mean_k diffn_k * clamp(w_k, +-2)
      float fx = 0.f, fy = 0.f, fz = 0.f;
      if (valid && cq == 0) {
        const int L = ed.N - ed.R;
        const int b = node / L, i = ed.R + node % L;
        const size_t gi = (size_t)b * ed.N + i;
        const int j = __ldg(ed.nbr + gi * SLOTS + k);
        const float* pi = ed.pos + gi * 9 + 3;
        const float* pj = ed.pos + ((size_t)b * ed.N + j) * 9 + 3;
        const float dx = pi[0] - pj[0], dy = pi[1] - pj[1], dz = pi[2] - pj[2];
        const float rad = dx * dx + dy * dy + dz * dz;
        const float sc = fminf(fmaxf(tot, -2.f), 2.f) / (sqrtf(rad + 1e-8f) + 1.0f);
        fx = dx * sc; fy = dy * sc; fz = dz * sc;
      }
      fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz);
      if (cq == 0 && lane == 0) { aggp[q * 4] = fx; aggp[q * 4 + 1] = fy; aggp[q * 4 + 2] = fz; }
      __syncthreads();
      if (tid < 2) {
        const int nd = tile * 2 + tid;
        if (nd < total_nodes) {
          const float inv = 1.f / (float)ed.K;
          float* fo = ed.fbuf + (size_t)nd * 4;
          fo[0] = (aggp[(2 * tid) * 4] + aggp[(2 * tid + 1) * 4]) * inv;
          fo[1] = (aggp[(2 * tid) * 4 + 1] + aggp[(2 * tid + 1) * 4 + 1]) * inv;
          fo[2] = (aggp[(2 * tid) * 4 + 2] + aggp[(2 * tid + 1) * 4 + 2]) * inv;
          fo[3] = 0.f;
        }
      }
      return;
    }
    // EDGE: gate, optional m* spill for the coordinate head, segment sum over the residue's rows
    const float g = valid ? __fdividef(1.f, 1.f + __expf(-(tot + p.ba[0]))) : 0.f;
#pragma unroll
    for (int e = 0; e < CW; ++e) m[e] *= g;
    if (ed.last && valid) {
      const int b = node / ed.N, i = node % ed.N;
      if (i >= ed.R) {
        __half* dst = ed.mstar + (((size_t)b * (ed.N - ed.R) + (i - ed.R)) * SLOTS + k) * H + cq * CW;
#pragma unroll
        for (int e8 = 0; e8 < CW / 8; ++e8) {
          float x8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) x8[e] = m[e8 * 8 + e] * S_SCALE;
          *reinterpret_cast<uint4*>(dst + e8 * 8) = pack8(x8);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < CW / 32; ++c) {
      const float cs = lane_transpose_sum(m + c * 32, lane);
      aggp[q * 256 + cq * CW + c * 32 + lane] = cs;
    }
    __syncthreads();
    if (tid < 256) {
#pragma unroll
      for (int hn2 = 0; hn2 < 2; ++hn2) {
        const int nd = tile * 2 + hn2;
        if (nd < total_nodes) ed.agg[(size_t)nd * H + tid] = aggp[(2 * hn2) * 256 + tid] + aggp[(2 * hn2 + 1) * 256 + tid];
      }
    }
  };

  // ---------------------------------------------------------------------------------------------
  const uint64_t dW = make_desc(sbase + OFF_W);
  const uint64_t dS = make_desc(sbase + OFF_S);
  int it = 0, prev_tile = -1;
  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
    const int buf = it & 1;
    build(tile, it);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 256);
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
        // K step kk: K-block kk/4 (64 elements), 32 bytes per step inside the 128-byte swizzle row
        const uint64_t da = dS + (uint64_t)(((kk >> 2) * S_KBLK + (kk & 3) * 32) >> 4);
        const uint64_t db = dW + (uint64_t)(((kk >> 2) * W_KBLK + (kk & 3) * 32) >> 4);
        mma_f16(d_tmem, da, db, kk > 0 ? 1u : 0u);
      }
      mma_commit(buf ? bar1 : bar0);
    }
    if (prev_tile >= 0) {
      const int pb = buf ^ 1;
      mbar_wait(pb ? bar1 : bar0, (uint32_t)(((it - 1) >> 1) & 1));
      tc_fence_after();
      epilogue(prev_tile, pb);
    }
    // the S tile may only be rebuilt once the MMA that reads it has retired
    mbar_wait(buf ? bar1 : bar0, (uint32_t)((it >> 1) & 1));
    tc_fence_after();
    prev_tile = tile;
  }
  if (prev_tile >= 0) {
    const int pb = (it - 1) & 1;
    epilogue(prev_tile, pb);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

template <int MODE, int VAR = 0>
static int launch(dfm_ctx* ctx, const Params& p, cudaStream_t s) {
  static unsigned long long attr_devices = 0;
  if (dfm_once_per_device(attr_devices, ctx->device)) {
    CUDA_TRY(cudaFuncSetAttribute(k_tc<MODE, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_ALLOC));
  }
  const int grid = p.ntiles < ctx->num_sms ? p.ntiles : ctx->num_sms;
  if (grid <= 0) return 0;
  k_tc<MODE, VAR><<<grid, NT, SMEM_ALLOC, s>>>(p);
  LAUNCH_CHECK(ctx);
  return 0;
}

}  // namespace tc

int launch_linear_tc(dfm_ctx* ctx, const LinearArgs& a, cudaStream_t s) {
  tc::Params p{};
  p.mode = tc::LINEAR;
  p.ntiles = (a.M + tc::TILE_M - 1) / tc::TILE_M;
  p.Wimg = a.Wimg;
  p.lin = a;
  p.v0 = a.bias;
  return tc::launch<tc::LINEAR>(ctx, p, s);
}

int launch_edge_tc(dfm_ctx* ctx, const EdgeArgs& a, cudaStream_t s) {
  const LayerW& w = ctx->layer[a.layer];
  tc::Params p{};
  p.mode = tc::EDGE;
  p.ntiles = (a.B * a.N + 1) / 2;
  p.Wimg = w.img_W2;
  p.ed = a;
  p.Tdrp = w.Tdrp16;
  p.Totp = w.Totp16;
  p.w1r = w.w1r;
  p.v0 = w.b2;
  p.v1 = w.wa;
  p.ba = w.ba;
  static int variant = -1;
  if (variant < 0) {
    const char* e = getenv("DFM_EDGE_VARIANT");
    variant = e ? atoi(e) : 0;
  }
  if (variant == 1) return tc::launch<tc::EDGE, 1>(ctx, p, s);
  return tc::launch<tc::EDGE, 0>(ctx, p, s);
}

int launch_coord_tc(dfm_ctx* ctx, const EdgeArgs& a, cudaStream_t s) {
  const LayerW& w = ctx->layer[a.layer];
  tc::Params p{};
  p.mode = tc::COORD;
  p.ntiles = (a.B * (a.N - a.R) + 1) / 2;
  p.Wimg = a.coord_img ? a.coord_img : w.img_Wc1;
  p.ed = a;
  p.v0 = w.bc1;
  p.v1 = w.wc2;
  return tc::launch<tc::COORD>(ctx, p, s);
}
