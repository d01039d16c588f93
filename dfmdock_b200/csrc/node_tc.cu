// Coordinate head of the last E_GCL layer on the tensor cores (ligand rows only):
//   w = clamp(wc2 . SiLU(Wc1 m* + bc1), +-2),  f_i = mean_k (x_i - x_j) / (|x_i - x_j| + 1) w      (src/models/egnn.py:118-148)
// m* = the gated messages edge_ws.cu spills for the ligand residues, [B*L, 64 slots, 256] fp16 (x 2^-6).
//
// One persistent CTA per SM, 19 warps:
//   warps 0-15  workers: epilogue from TMEM (SiLU, dot with wc2, clamp, displacement, mean over the 60 slots)
//   warps 16-17 loaders: the tile's K blocks HBM -> shared memory with cp.async, straight into the K-major SWIZZLE_128B
//               operand layout, one 64-column K block (16 KB) at a time into a ring of four blocks
//   warp  18    MMA issuer: one tcgen05.commit per K block (frees the ring slot) and one per tile (accumulator ready)
// The 128 KB fp16 image of Wc1 (x 2^6) is resident in shared memory; accumulators are double buffered in TMEM.
// (The other node-side contractions live in node_t.cu.)
#include "common.cuh"

namespace ntc {

constexpr int TILE_M = 128;
constexpr uint32_t W_BYTES = 256 * 256 * 2;          // 128 KB in every mode (256 x 256 or 128 x 512 fp16)
constexpr uint32_t S_KBLK = TILE_M * 128;            // one K block of the operand tile: 128 rows x 64 fp16
constexpr uint32_t OFF_W = 0;
constexpr uint32_t OFF_S = W_BYTES;                  // ring of 4 K blocks
constexpr uint32_t OFF_VEC = OFF_S + 4 * S_KBLK;     // 256 floats bias, 256 floats wc2
constexpr uint32_t OFF_PART = OFF_VEC + 2048;        // MODE_C: [4][128] dot partials, [4][4] force partials
constexpr uint32_t OFF_BAR = OFF_PART + 2048 + 64;   // full[4], empty[4], accf[2], acce[2], tmem base
constexpr uint32_t SMEM_BYTES = OFF_BAR + 128;
constexpr uint32_t SMEM_ALLOC = SMEM_BYTES + 1024;
constexpr int NWORK = 16;
constexpr int NT = (NWORK + 3) * 32;                 // 608

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

struct Params {
  int M, ntiles, N;
  const __half* X;       // gated messages of the ligand residues [B*L, 64, 256] fp16 (x 2^-6)
  const __half* W0;      // Wc1 x 2^6 image
  const float* bias0;    // bc1
  const float* wc2;      // [256]
  const int32_t* nbr;    // [B*N, 64]
  const float* pos;      // [B*N, 3, 3] centred backbone
  float* fbuf;           // [B*L, 4] out
  int R, K;
};

__global__ void __launch_bounds__(NT, 1) k_coord(const Params p) {
  constexpr int KB = 4;                                   // K blocks per tile
  constexpr int NCOL = 256;                               // accumulator columns per tile
  constexpr uint32_t W_KBLK = NCOL * 128;                 // bytes per K block of the weight image
  constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(NCOL >> 3) << 17) | ((128u >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  float* vbias = reinterpret_cast<float*>(smem + OFF_VEC);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 112);
  const uint32_t bar_full = sbase + OFF_BAR, bar_empty = sbase + OFF_BAR + 32;
  const uint32_t bar_accf = sbase + OFF_BAR + 64, bar_acce = sbase + OFF_BAR + 80;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = (int)blockIdx.x, ncta = (int)gridDim.x;

  // ---- setup: weight image (four bulk copies, async proxy; only the MMA issuer waits for them), vectors, barriers, TMEM
  const uint32_t bar_w = sbase + OFF_BAR + 96;
  if (tid < 256) {
    vbias[tid] = p.bias0[tid];
    vbias[256 + tid] = p.wc2[tid];
  }
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) { mbar_init(bar_full + 8 * i, 32); mbar_init(bar_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_accf + 8 * i, 1); mbar_init(bar_acce + 8 * i, NWORK * 32); }
    mbar_init(bar_w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_w), "r"(W_BYTES) : "memory");
    const char* wsrc = reinterpret_cast<const char*>(p.W0);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(sbase + OFF_W + (uint32_t)i * (W_BYTES / 4)), "l"(wsrc + (size_t)i * (W_BYTES / 4)), "r"(W_BYTES / 4), "r"(bar_w) : "memory");
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

  if (warp == NWORK + 2) {
    // =================================== MMA ISSUER ===================================================
    if (lane == 0) {
      const uint64_t dW = make_desc(sbase + OFF_W);
      const uint64_t dS = make_desc(sbase + OFF_S);
      int it = 0;
      uint32_t c = 0;                                        // running K-block count -> ring slot / phase
      mbar_wait(bar_w, 0u);                                  // weight image landed
      for (int tile = cta; tile < p.ntiles; tile += ncta, ++it) {
        const int buf = it & 1;
        if (it >= 2) mbar_wait(bar_acce + 8 * buf, (uint32_t)(((it >> 1) - 1) & 1));
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * NCOL);
#pragma unroll 1
        for (int kb = 0; kb < KB; ++kb, ++c) {
          const uint32_t slot = c & 3u;
          mbar_wait(bar_full + 8 * slot, (c >> 2) & 1u);
          tc_fence_after();
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const uint64_t da = dS + (uint64_t)((slot * S_KBLK + k4 * 32) >> 4);
            const uint64_t db = dW + (uint64_t)(((uint32_t)kb * W_KBLK + k4 * 32) >> 4);
            mma_f16(d_tmem, da, db, IDESC, (kb | k4) ? 1u : 0u);
          }
          mma_commit(bar_empty + 8 * slot);
        }
        mma_commit(bar_accf + 8 * buf);
      }
    }
    __syncwarp();
  } else if (warp >= NWORK) {
    // =================================== LOADERS ======================================================
    const int lw = warp - NWORK;                           // loader 0 takes even K blocks, loader 1 odd ones
    const int c8 = lane & 7, rsub = lane >> 3;
    uint32_t c = 0;
    for (int tile = cta; tile < p.ntiles; tile += ncta) {
      const int row0 = tile * TILE_M;
#pragma unroll 1
      for (int kb = 0; kb < KB; ++kb, ++c) {
        if ((kb & 1) != lw) continue;
        const uint32_t slot = c & 3u;
        if (c >= 4) mbar_wait(bar_empty + 8 * slot, ((c >> 2) - 1) & 1u);
        const int kcol = kb * 64 + c8 * 8;
        const uint32_t dst0 = sbase + OFF_S + slot * S_KBLK;
#pragma unroll 8
        for (int i = 0; i < 32; ++i) {
          const int r = rsub + 4 * i;
          const int m = row0 + r;
          // pad slots (K..63) are never written by the edge kernel: zero-fill them instead of reading stale memory
          const bool ok = m < p.M && (r & 63) < p.K;
          cp_async16(dst0 + (uint32_t)r * 128u + (uint32_t)((c8 ^ (r & 7)) << 4), p.X + (size_t)(ok ? m : 0) * H + kcol, ok ? 16u : 0u);
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar_full + 8 * slot) : "memory");
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
  } else {
    // =================================== WORKERS ======================================================
    const int q = warp & 3, cq = warp >> 2;                  // TMEM lane quarter / column quarter
    const int erow = q * 32 + lane;
    constexpr int CW = NCOL / 4;                             // accumulator columns per thread (64)

    auto epilogue = [&](int tile, int buf) {
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * NCOL + cq * CW);
      float dotp = 0.f;
#pragma unroll
      for (int c = 0; c < CW / 32; ++c) {
        float v[32];
        tmem_ld32_issue(taddr + c * 32, v);
        tmem_ld_wait();
        if (c == CW / 32 - 1) {               // all TMEM reads of this thread are done: release the accumulator buffer
          tc_fence_before();
          mbar_arrive(bar_acce + 8 * buf);
        }
        const int col0 = cq * CW + c * 32;
#pragma unroll
        for (int e = 0; e < 32; ++e) dotp = fmaf(silu_tanh(v[e] + vbias[col0 + e]), vbias[256 + col0 + e], dotp);
      }
      // tile = 2 ligand residues x 64 slots: clamp(dot) -> displacement along x_i - x_j -> mean over the K slots
      float* part = reinterpret_cast<float*>(smem + OFF_PART);
      float* fpart = part + 512;
      const int total = p.M / SLOTS;                       // B * L residues
      const int node = tile * 2 + (erow >> 6), k = erow & 63;
      const bool valid = node < total && k < p.K;
      part[cq * 128 + erow] = dotp;
      asm volatile("bar.sync 1, 512;" ::: "memory");
      float fx = 0.f, fy = 0.f, fz = 0.f;
      if (cq == 0 && valid) {
        const float tot = (part[erow] + part[128 + erow]) + (part[256 + erow] + part[384 + erow]);
        const int L = p.N - p.R;
        const int b = node / L, i = p.R + node % L;
        const size_t gi = (size_t)b * p.N + i;
        const int j = __ldg(p.nbr + gi * SLOTS + k);
        const float* pi = p.pos + gi * 9 + 3;
        const float* pj = p.pos + ((size_t)b * p.N + j) * 9 + 3;
        const float dx = pi[0] - pj[0], dy = pi[1] - pj[1], dz = pi[2] - pj[2];
        const float rad = dx * dx + dy * dy + dz * dz;
        const float sc = fminf(fmaxf(tot, -2.f), 2.f) / (sqrtf(rad + 1e-8f) + 1.0f);
        fx = dx * sc; fy = dy * sc; fz = dz * sc;
      }
      if (cq == 0) {
        fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz);
        if (lane == 0) { fpart[q * 4] = fx; fpart[q * 4 + 1] = fy; fpart[q * 4 + 2] = fz; }
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");
      if (tid < 2) {
        const int nd = tile * 2 + tid;
        if (nd < total) {
          const float inv = 1.f / (float)p.K;
          float* fo = p.fbuf + (size_t)nd * 4;
          fo[0] = (fpart[(2 * tid) * 4] + fpart[(2 * tid + 1) * 4]) * inv;
          fo[1] = (fpart[(2 * tid) * 4 + 1] + fpart[(2 * tid + 1) * 4 + 1]) * inv;
          fo[2] = (fpart[(2 * tid) * 4 + 2] + fpart[(2 * tid + 1) * 4 + 2]) * inv;
          fo[3] = 0.f;
        }
      }
    };

    int it = 0;
    for (int tile = cta; tile < p.ntiles; tile += ncta, ++it) {
      const int buf = it & 1;
      mbar_wait(bar_accf + 8 * buf, (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      epilogue(tile, buf);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NWORK + 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

static int launch(dfm_ctx* ctx, const Params& p, int grid, cudaStream_t s) {
  static unsigned long long attr_devices = 0;
  if (dfm_once_per_device(attr_devices, ctx->device)) {
    CUDA_TRY(cudaFuncSetAttribute(k_coord, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_ALLOC));
  }
  if (grid <= 0) return 0;
  k_coord<<<grid, NT, SMEM_ALLOC, s>>>(p);
  LAUNCH_CHECK(ctx);
  return 0;
}

}  // namespace ntc

// fp16 image of W3 = [W3h | W3a x 2^6] (K = 512) for one half of the output columns, K-major SWIZZLE_128B:
// 8 K blocks of [128 rows x 128 B]; 16-byte chunk c of row n sits at chunk c ^ (n & 7).
__global__ void k_image_pack_z(const float* __restrict__ W3, int half, float scale_hi, __half* __restrict__ img) {
  const int n = blockIdx.x, k = threadIdx.x + blockIdx.y * 256;      // n in [0,128), k in [0,512)
  const int kb = k >> 6, c = (k & 63) >> 3, e = k & 7;
  const float v = W3[(size_t)(half * 128 + n) * 512 + k] * (k >= 256 ? scale_hi : 1.f);
  img[(((size_t)kb * 128 + n) * 8 + (c ^ (n & 7))) * 8 + e] = __float2half_rn(v);
}
int launch_image_pack_z(dfm_ctx* ctx, const float* W3, float scale_hi, __half* img0, __half* img1, cudaStream_t s) {
  k_image_pack_z<<<dim3(128, 2), 256, 0, s>>>(W3, 0, scale_hi, img0);
  LAUNCH_CHECK(ctx);
  k_image_pack_z<<<dim3(128, 2), 256, 0, s>>>(W3, 1, scale_hi, img1);
  LAUNCH_CHECK(ctx);
  return 0;
}

// per-residue force of the ligand (last layer)
int launch_node_coord(dfm_ctx* ctx, const EdgeArgs& a, cudaStream_t s) {
  const LayerW& w = ctx->layer[a.layer];
  ntc::Params p{};
  const int L = a.N - a.R;
  p.M = a.B * L * SLOTS; p.ntiles = (a.B * L + 1) / 2; p.N = a.N; p.R = a.R; p.K = a.K;
  p.X = a.mstar; p.W0 = w.img_Wc1s; p.bias0 = w.bc1; p.wc2 = w.wc2; p.nbr = a.nbr; p.pos = a.pos; p.fbuf = a.fbuf;
  return ntc::launch(ctx, p, p.ntiles < ctx->num_sms ? p.ntiles : ctx->num_sms, s);
}
