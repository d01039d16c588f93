// Coordinate head of the last E_GCL layer on the tensor cores (ligand rows only):
//   w = clamp(wc2 . SiLU(Wc1 m* + bc1), +-2),  f_i = mean_k (x_i - x_j) / (|x_i - x_j| + 1) w      (src/models/egnn.py:118-148)
// m* = the gated messages edge_ws.cu spills for the ligand residues, [B*L, 64 slots, 256] fp16 (x 2^-6).
//
// The spill is stored with the columns of each 128-column half in the epilogue's fragment order (edge_ws.cu, position
// cq*32 + 2j + e <-> column 8j + 2cq + e, so that a thread stores 64 contiguous bytes); the image of Wc1 carries the same
// permutation along K (launch_image_pack_perm), so the contraction is unchanged.
//
// One persistent CTA per SM, 19 warps:
//   warps 0-15  workers: epilogue from TMEM (SiLU, dot with wc2, clamp, displacement, mean over the 60 slots)
//   warp  16    loader: one elected thread issues one TMA tensor-map load (cp.async.bulk.tensor.2d, 64 columns x 128 rows,
//               SWIZZLE_128B) per K block into a ring of five 16 KB slots -- 80 KB in flight per SM, what HBM latency needs
//   warp  18    MMA issuer: one tcgen05.commit per K block (frees the ring slot) and one per tile (accumulator ready)
// The 128 KB fp16 image of Wc1 (x 2^6) is resident in shared memory; accumulators are double buffered in TMEM.
// (The other node-side contractions live in node_t.cu.)
#include "common.cuh"
#include "tma.cuh"

namespace ntc {

constexpr int TILE_M = 128;
constexpr uint32_t W_BYTES = 256 * 256 * 2;          // 128 KB in every mode (256 x 256 or 128 x 512 fp16)
constexpr uint32_t S_KBLK = TILE_M * 128;            // one K block of the operand tile: 128 rows x 64 fp16
constexpr uint32_t OFF_W = 0;
constexpr int NSLOT = 5;                             // ring slots (K blocks of 16 KB)
constexpr uint32_t OFF_S = W_BYTES;                  // ring of NSLOT K blocks
constexpr uint32_t OFF_VEC = OFF_S + NSLOT * S_KBLK; // 256 floats bias, 256 floats wc2
constexpr uint32_t OFF_PART = OFF_VEC + 2048;        // [2][4][128] dot partials, [2][4][4] force partials
constexpr uint32_t OFF_BAR = OFF_PART + 4096 + 128;   // full[8], empty[8], accf[2], acce[2], w, tmem base
constexpr uint32_t SMEM_BYTES = OFF_BAR + 192;
static_assert(SMEM_BYTES + 1024 <= 232448, "shared memory budget");
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

__global__ void __launch_bounds__(NT, 1) k_coord(const Params p, const __grid_constant__ CUtensorMap tmX) {
  constexpr int KB = 4;                                   // K blocks per tile
  constexpr int NCOL = 256;                               // accumulator columns per tile
  constexpr uint32_t W_KBLK = NCOL * 128;                 // bytes per K block of the weight image
  constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(NCOL >> 3) << 17) | ((128u >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  float* vbias = reinterpret_cast<float*>(smem + OFF_VEC);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 176);
  const uint32_t bar_full = sbase + OFF_BAR, bar_empty = sbase + OFF_BAR + 64;
  const uint32_t bar_accf = sbase + OFF_BAR + 128, bar_acce = sbase + OFF_BAR + 144;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = (int)blockIdx.x, ncta = (int)gridDim.x;
  pdl_trigger();            // programmatic dependent launch, see common.cuh

  // ---- setup: weight image (four bulk copies, async proxy; only the MMA issuer waits for them), vectors, barriers, TMEM
  const uint32_t bar_w = sbase + OFF_BAR + 160;
  if (tid < 256) {
    vbias[tid] = p.bias0[tid];
    vbias[256 + tid] = p.wc2[tid];
  }
  if (tid == 0) {
    for (int i = 0; i < NSLOT; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_accf + 8 * i, 1); mbar_init(bar_acce + 8 * i, NWORK * 32); }
    mbar_init(bar_w, 1);
    tma_prefetch_desc(&tmX);
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
  pdl_wait();               // the edge kernel's spill is complete and visible from here on

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
          const uint32_t slot = c % NSLOT;
          mbar_wait(bar_full + 8 * slot, (c / NSLOT) & 1u);
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
    // =================================== LOADER (TMA) =================================================
    if (warp == NWORK && lane == 0) {
      uint32_t c = 0;
      for (int tile = cta; tile < p.ntiles; tile += ncta) {
        const int row0 = tile * TILE_M;
#pragma unroll 1
        for (int kb = 0; kb < KB; ++kb, ++c) {
          const uint32_t slot = c % NSLOT;
          if (c >= (uint32_t)NSLOT) mbar_wait(bar_empty + 8 * slot, ((c / NSLOT) - 1) & 1u);
          mbar_expect_tx(bar_full + 8 * slot, S_KBLK);
          tma_load_2d(sbase + OFF_S + slot * S_KBLK, &tmX, kb * 64, row0, bar_full + 8 * slot);
        }
      }
    }
    __syncwarp();
  } else {
    // =================================== WORKERS ======================================================
    const int q = warp & 3, cq = warp >> 2;                  // TMEM lane quarter / column quarter
    const int erow = q * 32 + lane;
    constexpr int CW = NCOL / 4;                             // accumulator columns per thread (64)

    const uint32_t vb_s = sbase + OFF_VEC + (uint32_t)(cq * CW) * 4u;      // bc1 / wc2 of this thread's 64 columns (shared space)
    const int total = p.M / SLOTS;                                          // B * L ligand residues
    const int L = p.N - p.R;
    float* part = reinterpret_cast<float*>(smem + OFF_PART);                // [2 tile parities][4 column quarters][128 rows]
    float* fpart = part + 1024;                                             // [2][4 row quarters][4]

    int it = 0;
    for (int tile = cta; tile < p.ntiles; tile += ncta, ++it) {
      const int buf = it & 1;
      // geometry of this row's edge, fetched BEFORE the accumulator is waited for (it does not depend on the MMA; the
      // nbr -> pos dependent loads used to sit between two CTA-wide barriers and cost ~2 k cycles per tile):
      // g = (x_i - x_j) / ((|x_i - x_j| + 1) K), row `erow` of the tile = slot erow & 63 of ligand residue tile*2 + (erow >> 6)
      float gx = 0.f, gy = 0.f, gz = 0.f;
      if (cq == 0) {
        const int node = tile * 2 + (erow >> 6), k = erow & 63;
        if (node < total && k < p.K) {
          const int b = node / L, i = p.R + node % L;
          const size_t gi = (size_t)b * p.N + i;
          const int j = __ldg(p.nbr + gi * SLOTS + k);
          const float* pi = p.pos + gi * 9 + 3;
          const float* pj = p.pos + ((size_t)b * p.N + j) * 9 + 3;
          const float dx = __ldg(pi) - __ldg(pj), dy = __ldg(pi + 1) - __ldg(pj + 1), dz = __ldg(pi + 2) - __ldg(pj + 2);
          const float rad = dx * dx + dy * dy + dz * dz;
          const float sc = 1.f / ((sqrtf(rad + 1e-8f) + 1.0f) * (float)p.K);
          gx = dx * sc; gy = dy * sc; gz = dz * sc;
        }
      }
      mbar_wait(bar_accf + 8 * buf, (uint32_t)((it >> 1) & 1));
      tc_fence_after();
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
#pragma unroll
        for (int e4 = 0; e4 < 8; ++e4) {
          float4 bb, ww;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bb.x), "=f"(bb.y), "=f"(bb.z), "=f"(bb.w) : "r"(vb_s + (uint32_t)(c * 32 + e4 * 4) * 4u));
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(ww.x), "=f"(ww.y), "=f"(ww.z), "=f"(ww.w) : "r"(vb_s + 1024u + (uint32_t)(c * 32 + e4 * 4) * 4u));
          dotp = fmaf(silu_tanh(v[e4 * 4] + bb.x), ww.x, dotp);
          dotp = fmaf(silu_tanh(v[e4 * 4 + 1] + bb.y), ww.y, dotp);
          dotp = fmaf(silu_tanh(v[e4 * 4 + 2] + bb.z), ww.z, dotp);
          dotp = fmaf(silu_tanh(v[e4 * 4 + 3] + bb.w), ww.w, dotp);
        }
      }
      // the four column quarters of a row meet in shared memory (double buffered by tile parity: only the cq == 0 warps
      // go on to the reduction, the other twelve move straight to the next tile)
      part[buf * 512 + cq * 128 + erow] = dotp;
      asm volatile("bar.sync 1, 512;" ::: "memory");
      if (cq == 0) {
        const float* pp = part + buf * 512;
        const float tot = (pp[erow] + pp[128 + erow]) + (pp[256 + erow] + pp[384 + erow]);
        const float w = fminf(fmaxf(tot, -2.f), 2.f);
        const float fx = warp_sum(gx * w), fy = warp_sum(gy * w), fz = warp_sum(gz * w);
        float* fp = fpart + buf * 16;
        if (lane == 0) { fp[q * 4] = fx; fp[q * 4 + 1] = fy; fp[q * 4 + 2] = fz; }
        asm volatile("bar.sync 2, 128;" ::: "memory");          // the four row-quarter warps
        if (tid < 2) {
          const int nd = tile * 2 + tid;
          if (nd < total) {
            float* fo = p.fbuf + (size_t)nd * 4;
            fo[0] = fp[(2 * tid) * 4] + fp[(2 * tid + 1) * 4];
            fo[1] = fp[(2 * tid) * 4 + 1] + fp[(2 * tid + 1) * 4 + 1];
            fo[2] = fp[(2 * tid) * 4 + 2] + fp[(2 * tid + 1) * 4 + 2];
            fo[3] = 0.f;
          }
        }
      }
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
  CUtensorMap tmX;
  int rc = dfm_make_tmap_f16(&tmX, p.X, (uint64_t)p.M, H, TILE_M);
  if (rc) return rc;
  CUDA_TRY(dfm_launch_pdl(k_coord, dim3(grid), dim3(NT), SMEM_ALLOC, s, p, tmX));
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
