// Body of the coordinate head (shared by the stand-alone kernel ntc::k_coord, node_tc.cu, and the fused last-layer kernel
// ews::k_last_fused, edge_ws.cu).  See node_tc.cu for the description of the roles.
#pragma once
#include "common.cuh"
#include "tma.cuh"
#include "last_ring.cuh"

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

// FUSED = false: the stand-alone kernel (k_coord): tiles of the spill buffer X, two ligand residues each, strided over the
// CTAs.  FUSED = true: the coordinate-head role inside ews::k_last_fused (896-thread CTAs; warps 19-27 idle): tiles arrive
// through the ring of last_ring.cuh in hand-over order; a tile holds two consecutive residues of the node-pair grid, of
// which only ligand residues produce a force.
template <bool FUSED>
__device__ __forceinline__ void coord_body(const Params& p, const CUtensorMap& tmX, const LastRing& ring, uint8_t* smem,
                                           const int cta, const int ncta) {
  constexpr int KB = 4;                                   // K blocks per tile
  constexpr int NCOL = 256;                               // accumulator columns per tile
  constexpr uint32_t W_KBLK = NCOL * 128;                 // bytes per K block of the weight image
  constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(NCOL >> 3) << 17) | ((128u >> 4) << 24);

  const uint32_t sbase = smem_u32(smem);
  float* vbias = reinterpret_cast<float*>(smem + OFF_VEC);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 176);
  const uint32_t bar_full = sbase + OFF_BAR, bar_empty = sbase + OFF_BAR + 64;
  const uint32_t bar_accf = sbase + OFF_BAR + 128, bar_acce = sbase + OFF_BAR + 144;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // work list: stand-alone: tile = cta, cta + ncta, .. < ntiles;  fused: sequence number s = cta, cta + ncta, .. <
  // chunk * P, of which those with p * chunk + tau < ntiles exist (tau = s / P, p = s % P)
  const int n_items = FUSED ? ring.chunk * ring.P : p.ntiles;
  auto item_tile = [&](int s) -> int {       // stand-alone: the tile itself; fused: logical tile or -1
    if (!FUSED) return s;
    const int tau = s / ring.P, pp = s - tau * ring.P;
    const int lt = pp * ring.chunk + tau;
    return (tau < ring.chunk && lt < ring.ntiles) ? lt : -1;
  };
  if (FUSED) {
    if (warp < NWORK) asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  }

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
      for (int item = cta; item < n_items; item += ncta) {
        if (item_tile(item) < 0) continue;
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
        // all four K blocks of the tile have landed in shared memory (observed through bar_full): the ring slot may be
        // refilled by its next producer
        if (FUSED) red_release_gpu_add(ring.done + (unsigned)item % (unsigned)ring.NR, 1u);
        ++it;
      }
    }
    __syncwarp();
  } else if (warp >= NWORK) {
    // =================================== LOADER (TMA) =================================================
    if (warp == NWORK && lane == 0) {
      uint32_t c = 0;
      for (int item = cta; item < n_items; item += ncta) {
        const int tile = item_tile(item);
        if (tile < 0) continue;
        int row0 = tile * TILE_M;
        if (FUSED) {
          const unsigned rs = (unsigned)item % (unsigned)ring.NR, epoch = (unsigned)item / (unsigned)ring.NR;
          ring_wait_ge(ring.ready + rs, 8u * (epoch + 1u));      // the eight epilogue warps of the producer have written it
          fence_proxy_async_all();                                // their generic-proxy writes -> this thread's TMA reads
          row0 = (int)rs * TILE_M;
        }
#pragma unroll 1
        for (int kb = 0; kb < KB; ++kb, ++c) {
          const uint32_t slot = c % NSLOT;
          if (c >= (uint32_t)NSLOT) mbar_wait(bar_empty + 8 * slot, ((c / NSLOT) - 1) & 1u);
          mbar_expect_tx(bar_full + 8 * slot, S_KBLK);
          tma_load_2d(sbase + OFF_S + slot * S_KBLK, &tmX, 0, row0 * 4 + kb * TILE_M, bar_full + 8 * slot);   // [tile][K block][128][64]
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

    // residue of tile row `r64` (0 / 1): stand-alone: ligand residue tile * 2 + r64 of B * L; fused: residue ptile * 2 + r64
    // of the node-pair grid, a ligand residue or not.  -> global row gi (or -1) and index into fbuf
    // (32-bit arithmetic only: this runs once per tile in the four cq == 0 warps while the other twelve wait at the barrier)
    auto residue = [&](int tile, int r64, int& lig_index, int& b) -> int {
      if (!FUSED) {
        const int node = tile * 2 + r64;
        lig_index = node;
        b = node / L;
        return node < total ? b * p.N + p.R + (node - b * L) : -1;
      }
      const int node = ring_phys(ring, tile) * 2 + r64;
      b = node / p.N;
      const int i = node - b * p.N;
      lig_index = b * L + i - p.R;
      return (node < ring.total_nodes && i >= p.R) ? node : -1;
    };
    int it = 0;
    for (int item = cta; item < n_items; item += ncta) {
      const int tile = item_tile(item);
      if (tile < 0) continue;
      const int buf = it & 1;
      // geometry of this row's edge, fetched BEFORE the accumulator is waited for (it does not depend on the MMA; the
      // nbr -> pos dependent loads used to sit between two CTA-wide barriers and cost ~2 k cycles per tile):
      // g = (x_i - x_j) / ((|x_i - x_j| + 1) K), row `erow` of the tile = slot erow & 63 of ligand residue tile*2 + (erow >> 6)
      float gx = 0.f, gy = 0.f, gz = 0.f;
      if (cq == 0) {
        const int k = erow & 63;
        int li, b;
        const int gl = residue(tile, erow >> 6, li, b);
        if (gl >= 0 && k < p.K) {
          const size_t gi = (size_t)gl;
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
      asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");     // only the four warps that share this row quarter meet
      if (cq == 0) {
        const float* pp = part + buf * 512;
        const float tot = (pp[erow] + pp[128 + erow]) + (pp[256 + erow] + pp[384 + erow]);
        const float w = fminf(fmaxf(tot, -2.f), 2.f);
        const float fx = warp_sum(gx * w), fy = warp_sum(gy * w), fz = warp_sum(gz * w);
        float* fp = fpart + buf * 16;
        if (lane == 0) { fp[q * 4] = fx; fp[q * 4 + 1] = fy; fp[q * 4 + 2] = fz; }
        asm volatile("bar.sync 5, 128;" ::: "memory");          // the four row-quarter warps
        if (tid < 2) {
          int nd, bb;
          if (residue(tile, tid, nd, bb) >= 0) {
            float* fo = p.fbuf + (size_t)nd * 4;
            fo[0] = fp[(2 * tid) * 4] + fp[(2 * tid + 1) * 4];
            fo[1] = fp[(2 * tid) * 4 + 1] + fp[(2 * tid + 1) * 4 + 1];
            fo[2] = fp[(2 * tid) * 4 + 2] + fp[(2 * tid + 1) * 4 + 2];
            fo[3] = 0.f;
          }
        }
      }
      ++it;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NWORK + 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

}  // namespace ntc
