// Node-side tcgen05 kernels of the throughput path (per-residue 256x256 contractions, HBM-bound):
//   MODE_AB  Ah = fp16((W1s h + b1)/2), Bm = fp16((W1d h)/2)   node halves of edge_mlp.0   (src/models/egnn.py:95-104)
//   MODE_Z   z = W3h h + W3a agg + b3                           node_mlp.0 on [h, agg]      (src/models/egnn.py:106-116)
//   MODE_H   h += W4 SiLU(GraphNorm(z)) + b4  (+ fp16 copy)     node_mlp.1-3 + residual     (src/models/egnn.py:74,106-116)
//   MODE_C   w = clamp(wc2 . SiLU(Wc1 m* + bc1), +-2), f_i = mean_k (x_i - x_j)/(|x_i - x_j| + 1) w   coord_model of the
//            last layer, ligand rows only                                                   (src/models/egnn.py:118-148)
//
// fp16 activations (h16, agg16) go HBM -> shared memory with cp.async straight into the K-major SWIZZLE_128B operand
// layout (no registers, no ALU); the 128 KB fp16 weight image is resident in shared memory; accumulators are
// double buffered in TMEM so that the epilogue of tile t overlaps the loads of tile t+1.  MODE_Z swaps the weight
// image (W3h -> W3a) between the two accumulation passes of a pair of tiles instead of writing z twice.
#include "common.cuh"

namespace ntc {

constexpr int TILE_M = 128;
constexpr uint32_t W_BYTES = 256 * 256 * 2;
constexpr uint32_t S_BYTES = TILE_M * 256 * 2;
constexpr uint32_t W_KBLK = 256 * 128;
constexpr uint32_t S_KBLK = TILE_M * 128;
constexpr uint32_t OFF_W = 0;
constexpr uint32_t OFF_S = W_BYTES;
constexpr uint32_t OFF_VEC = OFF_S + S_BYTES;        // 256 floats bias, 256 floats wc2
constexpr uint32_t OFF_PART = OFF_VEC + 2048;        // MODE_C: [4][128] dot partials, [4][4] force partials
constexpr uint32_t OFF_BAR = OFF_PART + 2048 + 64;   // 2 mbarriers + tmem base
constexpr uint32_t SMEM_BYTES = OFF_BAR + 64;
constexpr uint32_t SMEM_ALLOC = SMEM_BYTES + 1024;
constexpr int NT = 512;
constexpr uint32_t IDESC = (1u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok, tries = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok && ++tries > (1u << 24)) __trap();
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
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ uint32_t s_off(int r, int c16) {
  return (uint32_t)(c16 >> 3) * S_KBLK + (uint32_t)r * 128u + (uint32_t)(((c16 & 7) ^ (r & 7)) << 4);
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

enum Mode { MODE_AB = 0, MODE_Z = 1, MODE_H = 2, MODE_C = 3 };

struct Params {
  int M, ntiles, N;
  // MODE_AB: X = h16; CTAs [0, grid/2) use W0/bias0/out0, the rest W1/(no bias)/out1; out = fp16(0.5 * (acc + bias))
  // MODE_Z : pass 1 X = h16 with W0, pass 2 X2 = agg16 with W1; out32 = acc + bias0
  // MODE_H : operand = fp16(SiLU(z * gscale[b] + gshift[b])) with W0; h = h + acc + bias0; also h16
  const __half* X;
  const __half* X2;
  const __half* W0;
  const __half* W1;
  const float* bias0;
  __half* out0;
  __half* out1;
  float* out32;
  const float* z;        // MODE_H
  const float* gscale;   // [B, 256]
  const float* gshift;   // [B, 256]
  float* h;              // MODE_H in/out
  __half* h16;           // MODE_H out
  // MODE_C: X = gated messages of the ligand residues [B*L, 64, 256] fp16 (x 2^-6; W0 = Wc1 x 2^6), bias0 = bc1
  const float* wc2;      // [256]
  const int32_t* nbr;    // [B*N, 64]
  const float* pos;      // [B*N, 3, 3] centred backbone
  float* fbuf;           // [B*L, 4] out
  int R, K;
};

template <int MODE>
__global__ void __launch_bounds__(NT, 1) k_node(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  float* vbias = reinterpret_cast<float*>(smem + OFF_VEC);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 16);
  const uint32_t bar0 = sbase + OFF_BAR, bar1 = sbase + OFF_BAR + 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // which half of the grid (MODE_AB only)
  const int half_grid = (int)gridDim.x >> 1;
  const int side = (MODE == MODE_AB && (int)blockIdx.x >= half_grid) ? 1 : 0;
  const int cta = (MODE == MODE_AB) ? ((int)blockIdx.x - side * half_grid) : (int)blockIdx.x;
  const int ncta = (MODE == MODE_AB) ? half_grid : (int)gridDim.x;

  auto load_weights = [&](const __half* img) {
    const char* src = reinterpret_cast<const char*>(img);
#pragma unroll 4
    for (int i = tid; i < (int)(W_BYTES / 16); i += NT) cp_async16(sbase + OFF_W + (uint32_t)i * 16u, src + (size_t)i * 16, 16u);
  };
  auto load_tile16 = [&](const __half* X, int tile) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = tid + NT * i;
      const int r = idx >> 5, c16 = idx & 31;
      const int m = tile * TILE_M + r;
      const bool ok = m < p.M;
      cp_async16(sbase + OFF_S + s_off(r, c16), X + (size_t)(ok ? m : 0) * H + c16 * 8, ok ? 16u : 0u);
    }
  };
  // MODE_H operand: y = SiLU(z * scale + shift) -> fp16; one warp per row, 8 columns per lane, 4 rows in flight
  auto build_h = [&](int tile) {
#pragma unroll 1
    for (int r4 = 0; r4 < 8; r4 += 4) {   // 8 rows per warp (128 rows / 16 warps), 4 at a time
      float4 z0[4], z1[4];
      int mrow[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = warp + 16 * (r4 + i);
        const int m = tile * TILE_M + r;
        mrow[i] = m;
        z0[i] = make_float4(0.f, 0.f, 0.f, 0.f); z1[i] = z0[i];
        if (m < p.M) {
          const float4* zp = reinterpret_cast<const float4*>(p.z + (size_t)m * H + lane * 8);
          z0[i] = __ldg(zp); z1[i] = __ldg(zp + 1);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = warp + 16 * (r4 + i);
        float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (mrow[i] < p.M) {
          const int b = mrow[i] / p.N;
          const float4* sc = reinterpret_cast<const float4*>(p.gscale + (size_t)b * H + lane * 8);
          const float4* sh = reinterpret_cast<const float4*>(p.gshift + (size_t)b * H + lane * 8);
          const float4 s0 = __ldg(sc), s1 = __ldg(sc + 1), h0 = __ldg(sh), h1 = __ldg(sh + 1);
          x[0] = silu_tanh(fmaf(z0[i].x, s0.x, h0.x)); x[1] = silu_tanh(fmaf(z0[i].y, s0.y, h0.y));
          x[2] = silu_tanh(fmaf(z0[i].z, s0.z, h0.z)); x[3] = silu_tanh(fmaf(z0[i].w, s0.w, h0.w));
          x[4] = silu_tanh(fmaf(z1[i].x, s1.x, h1.x)); x[5] = silu_tanh(fmaf(z1[i].y, s1.y, h1.y));
          x[6] = silu_tanh(fmaf(z1[i].z, s1.z, h1.z)); x[7] = silu_tanh(fmaf(z1[i].w, s1.w, h1.w));
        }
        *reinterpret_cast<uint4*>(smem + OFF_S + s_off(r, lane)) = pack8(x);
      }
    }
  };

  // ---- setup
  const __half* wimg0 = (MODE == MODE_AB && side) ? p.W1 : p.W0;
  load_weights(wimg0);
  if (tid < 256) {
    vbias[tid] = (MODE == MODE_AB && side) ? 0.f : (p.bias0 ? p.bias0[tid] : 0.f);
    if (MODE == MODE_C) vbias[256 + tid] = p.wc2[tid];
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
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint64_t dW = make_desc(sbase + OFF_W);
  const uint64_t dS = make_desc(sbase + OFF_S);
  const int q = warp & 3, cq = warp >> 2;
  const int erow = q * 32 + lane;
  uint32_t ph0 = 0, ph1 = 0;     // parities of the next completion of bar0 / bar1

  auto issue_mma = [&](int buf, uint32_t accumulate) {   // call from one thread after the operands are visible
    const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 256);
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const uint64_t da = dS + (uint64_t)(((kk >> 2) * S_KBLK + (kk & 3) * 32) >> 4);
      const uint64_t db = dW + (uint64_t)(((kk >> 2) * W_KBLK + (kk & 3) * 32) >> 4);
      mma_f16(d_tmem, da, db, (accumulate | (uint32_t)kk) ? 1u : 0u);
    }
    mma_commit(buf ? bar1 : bar0);
  };
  auto wait_mma = [&](int buf) {
    if (buf) { mbar_wait(bar1, ph1); ph1 ^= 1; } else { mbar_wait(bar0, ph0); ph0 ^= 1; }
    tc_fence_after();
  };
  // operands written by cp.async / st.shared -> visible to the tensor core, then one thread issues
  auto publish_and_mma = [&](int buf, uint32_t accumulate) {
    cp_async_wait_all();
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      issue_mma(buf, accumulate);
    }
  };
  auto epilogue = [&](int tile, int buf) {
    const int mrow = tile * TILE_M + erow;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 256 + cq * 64);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float v[32];
      tmem_ld32_issue(taddr + c * 32, v);
      tmem_ld_wait();
      if (mrow < p.M) {
        const int col0 = cq * 64 + c * 32;
        const size_t o = (size_t)mrow * H + col0;
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] += vbias[col0 + e];
        if (MODE == MODE_AB) {
          __half* out = side ? p.out1 : p.out0;
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] *= 0.5f;
#pragma unroll
          for (int e8 = 0; e8 < 4; ++e8) *reinterpret_cast<uint4*>(out + o + e8 * 8) = pack8(v + e8 * 8);
        } else if (MODE == MODE_Z) {
#pragma unroll
          for (int e4 = 0; e4 < 8; ++e4)
            *reinterpret_cast<float4*>(p.out32 + o + e4 * 4) = make_float4(v[e4 * 4], v[e4 * 4 + 1], v[e4 * 4 + 2], v[e4 * 4 + 3]);
        } else {
#pragma unroll
          for (int e4 = 0; e4 < 8; ++e4) {
            const float4 hv = *reinterpret_cast<const float4*>(p.h + o + e4 * 4);
            v[e4 * 4] += hv.x; v[e4 * 4 + 1] += hv.y; v[e4 * 4 + 2] += hv.z; v[e4 * 4 + 3] += hv.w;
            *reinterpret_cast<float4*>(p.h + o + e4 * 4) = make_float4(v[e4 * 4], v[e4 * 4 + 1], v[e4 * 4 + 2], v[e4 * 4 + 3]);
          }
#pragma unroll
          for (int e8 = 0; e8 < 4; ++e8) *reinterpret_cast<uint4*>(p.h16 + o + e8 * 8) = pack8(v + e8 * 8);
        }
      }
    }
    tc_fence_before();
  };

  // MODE_C epilogue: tile = 2 ligand residues x 64 slots (all threads take part: it synchronises the CTA)
  auto epilogue_c = [&](int tile, int buf) {
    float* part = reinterpret_cast<float*>(smem + OFF_PART);
    float* fpart = part + 512;
    const int total = p.M / SLOTS;                       // B * L residues
    const int node = tile * 2 + (erow >> 6), k = erow & 63;
    const bool valid = node < total && k < p.K;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 256 + cq * 64);
    float dotp = 0.f;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float v[32];
      tmem_ld32_issue(taddr + c * 32, v);
      tmem_ld_wait();
      const int col0 = cq * 64 + c * 32;
#pragma unroll
      for (int e = 0; e < 32; ++e) dotp = fmaf(silu_tanh(v[e] + vbias[col0 + e]), vbias[256 + col0 + e], dotp);
    }
    tc_fence_before();
    part[cq * 128 + erow] = dotp;
    __syncthreads();
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
    __syncthreads();
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

  if (MODE == MODE_Z) {
    // pairs of tiles: both accumulators take the h16 pass with W3h, then the weight image is swapped to W3a for the
    // agg16 pass; z is written once.
    bool w_is_0 = true;
    for (int t0 = cta * 2; t0 < p.ntiles; t0 += ncta * 2) {
      const int nt = (t0 + 1 < p.ntiles) ? 2 : 1;
      for (int pass = 0; pass < 2; ++pass) {
        if ((pass == 0) != w_is_0) {          // previous MMAs that read the old image have completed (waited below)
          load_weights(pass == 0 ? p.W0 : p.W1);
          w_is_0 = (pass == 0);
        }
        for (int u = 0; u < nt; ++u) {
          load_tile16(pass == 0 ? p.X : p.X2, t0 + u);
          publish_and_mma(u, pass ? 1u : 0u);
          wait_mma(u);                        // S (and W) may be overwritten
        }
      }
      for (int u = 0; u < nt; ++u) epilogue(t0 + u, u);
      __syncthreads();                        // accumulators drained before the next pair's first MMA overwrites them
    }
  } else {
    int it = 0;
    int tile = cta;
    if (tile < p.ntiles) {
      if (MODE == MODE_H) build_h(tile); else load_tile16(p.X, tile);
    }
    for (; tile < p.ntiles; tile += ncta, ++it) {
      const int buf = it & 1;
      publish_and_mma(buf, 0u);
      wait_mma(buf);
      const int ntile = tile + ncta;
      if (ntile < p.ntiles) {
        if (MODE == MODE_H) build_h(ntile); else load_tile16(p.X, ntile);   // in flight during the epilogue below
      }
      if (MODE == MODE_C) epilogue_c(tile, buf); else epilogue(tile, buf);
    }
  }
  cp_async_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

template <int MODE>
static int launch(dfm_ctx* ctx, const Params& p, int grid, cudaStream_t s) {
  static bool attr = false;
  if (!attr) {
    CUDA_TRY(cudaFuncSetAttribute(k_node<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_ALLOC));
    attr = true;
  }
  if (grid <= 0) return 0;
  k_node<MODE><<<grid, NT, SMEM_ALLOC, s>>>(p);
  LAUNCH_CHECK(ctx);
  return 0;
}

}  // namespace ntc

// Ah = fp16((W1s h + b1eff)/2) and Bm = fp16((W1d h)/2) in one launch (two halves of the grid)
int launch_node_ab(dfm_ctx* ctx, int layer, int M, const __half* h16, __half* Ah, __half* Bm, cudaStream_t s) {
  const LayerW& w = ctx->layer[layer];
  ntc::Params p{};
  p.M = M; p.ntiles = (M + ntc::TILE_M - 1) / ntc::TILE_M; p.N = ctx->N;
  p.X = h16; p.W0 = w.img_W1s; p.W1 = w.img_W1d; p.bias0 = w.b1eff; p.out0 = Ah; p.out1 = Bm;
  int half = ctx->num_sms / 2;
  if (half > p.ntiles) half = p.ntiles;
  return ntc::launch<ntc::MODE_AB>(ctx, p, 2 * half, s);
}

// z = W3h h + W3a agg + b3 (agg16 carries agg x 2^-6, img_W3a carries W3a x 2^6)
int launch_node_z(dfm_ctx* ctx, int layer, int M, const __half* h16, const __half* agg16, float* z, cudaStream_t s) {
  const LayerW& w = ctx->layer[layer];
  ntc::Params p{};
  p.M = M; p.ntiles = (M + ntc::TILE_M - 1) / ntc::TILE_M; p.N = ctx->N;
  p.X = h16; p.X2 = agg16; p.W0 = w.img_W3h; p.W1 = w.img_W3a; p.bias0 = w.b3; p.out32 = z;
  const int pairs = (p.ntiles + 1) / 2;
  return ntc::launch<ntc::MODE_Z>(ctx, p, pairs < ctx->num_sms ? pairs : ctx->num_sms, s);
}

// h += W4 SiLU(z * gscale + gshift) + b4; h16 = fp16(h)
int launch_node_h(dfm_ctx* ctx, int layer, int M, const float* z, const float* gscale, const float* gshift, float* h,
                  __half* h16, cudaStream_t s) {
  const LayerW& w = ctx->layer[layer];
  ntc::Params p{};
  p.M = M; p.ntiles = (M + ntc::TILE_M - 1) / ntc::TILE_M; p.N = ctx->N;
  p.W0 = w.img_W4; p.bias0 = w.b4; p.z = z; p.gscale = gscale; p.gshift = gshift; p.h = h; p.h16 = h16;
  return ntc::launch<ntc::MODE_H>(ctx, p, p.ntiles < ctx->num_sms ? p.ntiles : ctx->num_sms, s);
}

// per-residue force of the ligand (last layer): replaces tc.cu k_tc<COORD> on the throughput path
int launch_node_coord(dfm_ctx* ctx, const EdgeArgs& a, cudaStream_t s) {
  const LayerW& w = ctx->layer[a.layer];
  ntc::Params p{};
  const int L = a.N - a.R;
  p.M = a.B * L * SLOTS; p.ntiles = (a.B * L + 1) / 2; p.N = a.N; p.R = a.R; p.K = a.K;
  p.X = a.mstar; p.W0 = w.img_Wc1s; p.bias0 = w.bc1; p.wc2 = w.wc2; p.nbr = a.nbr; p.pos = a.pos; p.fbuf = a.fbuf;
  return ntc::launch<ntc::MODE_C>(ctx, p, p.ntiles < ctx->num_sms ? p.ntiles : ctx->num_sms, s);
}
