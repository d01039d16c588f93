// tcgen05 (5th-gen tensor core) kernel for a plain 256-wide contraction on fp32 activations:
//   LINEAR out = add + A W^T + bias        (the two halves of to_energy.0 on the final forward, score_net_mlsb.py:385-390)
//
// One persistent CTA per SM, 512 threads.  The fp16 weight image (128 KB, K-major SWIZZLE_128B) stays resident in
// shared memory; a 128-row activation tile (64 KB, same layout) is converted to fp16 and rebuilt per tile; accumulators
// are double buffered in TMEM (2 x 256 columns) so the MMA of tile t overlaps the epilogue of tile t-1.
//   D[128 x 256] (TMEM, lane = row, column = feature) = S[128 x 256] (smem) * W[256 x 256]^T (smem)
// (The first-generation fused edge / coordinate kernels that used to live here were superseded by edge_ws.cu and
// node_tc.cu and have been removed.)
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
constexpr uint32_t OFF_VEC = OFF_S + S_BYTES;        // 256 floats: bias
constexpr uint32_t OFF_BAR = OFF_VEC + 256 * 4;      // 2 mbarriers + tmem base
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
struct Params {
  int ntiles;
  const __half* Wimg;
  LinearArgs lin;
};

constexpr int NT = 512;            // threads per CTA: 16 warps = 4 TMEM lane quarters x 4 column quarters
constexpr int NWARP = NT / 32;
constexpr int CW = 256 / (NWARP / 4);   // accumulator columns per thread in the epilogue (64)

__global__ void __launch_bounds__(NT, 1) k_linear(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  float* vec0 = reinterpret_cast<float*>(smem + OFF_VEC);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 16);
  const uint32_t bar0 = sbase + OFF_BAR, bar1 = sbase + OFF_BAR + 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time setup: weight image -> smem, bias, barriers, TMEM
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.Wimg);
    uint4* dst = reinterpret_cast<uint4*>(smem + OFF_W);
#pragma unroll 4
    for (int i = tid; i < (int)(W_BYTES / 16); i += NT) dst[i] = __ldg(src + i);
    if (tid < 256) vec0[tid] = p.lin.bias ? p.lin.bias[tid] : 0.f;
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

  auto build = [&](int tile) {
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
  };

  auto epilogue = [&](int tile, int buf) {
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 256 + cq * CW);
    float m[CW];
#pragma unroll
    for (int c = 0; c < CW / 32; ++c) tmem_ld32_issue(taddr + c * 32, m + c * 32);
    tmem_ld_wait();
    tc_fence_before();
    const int mrow = tile * TILE_M + erow;
    if (mrow >= p.lin.M) return;
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
        *reinterpret_cast<uint4*>(p.lin.out16 + o + e8 * 8) = pack8(x8);
      }
    }
  };

  const uint64_t dW = make_desc(sbase + OFF_W);
  const uint64_t dS = make_desc(sbase + OFF_S);
  int it = 0, prev_tile = -1;
  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
    const int buf = it & 1;
    build(tile);
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

static int launch(dfm_ctx* ctx, const Params& p, cudaStream_t s) {
  static unsigned long long attr_devices = 0;
  if (dfm_once_per_device(attr_devices, ctx->device)) {
    CUDA_TRY(cudaFuncSetAttribute(k_linear, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_ALLOC));
  }
  const int grid = p.ntiles < ctx->num_sms ? p.ntiles : ctx->num_sms;
  if (grid <= 0) return 0;
  k_linear<<<grid, NT, SMEM_ALLOC, s>>>(p);
  LAUNCH_CHECK(ctx);
  return 0;
}

}  // namespace tc

int launch_linear_tc(dfm_ctx* ctx, const LinearArgs& a, cudaStream_t s) {
  tc::Params p{};
  p.ntiles = (a.M + tc::TILE_M - 1) / tc::TILE_M;
  p.Wimg = a.Wimg;
  p.lin = a;
  return tc::launch(ctx, p, s);
}
