// Warp-specialised tcgen05 edge kernel (throughput path of the fused edge MLP):
//   m = SiLU(W2 SiLU(u) + b2), gate, per-residue segment sum              (src/models/egnn.py:95-116, 139-148)
//   u = A_i + B_j + radial * w1r + T[d, relpos] + T[omega, theta, phi]     (SURVEY App. A.5 / A.7 decomposition)
//
// One persistent CTA per SM over a contiguous range of 128-row tiles (2 residues x 64 edge slots), 28 warps whose
// registers are re-partitioned with setmaxnreg (72 / 88 / 40):
//   warps 0-15  producers: table rows of every edge gathered into registers one K block ahead (L2), + A_i + the B_j row
//               the loaders staged + radial * w1r, SiLU via tanh.approx.f16x2 in packed half2, written in place into
//               the 128 x 256 fp16 operand tile S (SWIZZLE_128B, K-major), one 64-column K block at a time
//   warps 25-26 loaders: TMA tile::gather4 of the B_j rows of a K block straight into the S tile as soon as the previous
//               tile's MMA has consumed that block; warp 27 relays the completion to the producers (EWS_RELAY)
//   warp  24    MMA issuer: one K = 16 MMA that sets the accumulator to b2/2 (constant 1/16 tile x bias tile, no-swizzle
//               descriptors), then D[128 x 256] (TMEM, fp32) += S * (W2/2)^T: four tcgen05.mma per K block, one
//               tcgen05.commit per K block (frees it for the next tile) and one per tile (accumulator ready)
//   warps 16-23 epilogue: tcgen05.ld.16x256b fragments (a thread owns 4 rows x 16 column pairs of its warp's 32 x 128
//               block), SiLU in half2, gate logit (16-product half2 chains -> fp32, shuffles over the 4 lanes of a row,
//               two-half combine through shared memory), sigmoid gate, gate-weighted segment sum as in-thread HFMA2 +
//               a 14-step shuffle reduction, two-quarter combine through shared memory -> agg16; the last layer also
//               spills the gated messages of the ligand rows for the coordinate head and, when no energy is wanted,
//               walks only the tiles that hold a ligand residue (Params::lig_only)
// The fp16 weight image (128 KB, cp.async.bulk at start-up) stays resident in shared memory; accumulators are double
// buffered in TMEM (2 x 256 columns) so that build(t+1), MMA(t) and epilogue(t-1) overlap.
//
// All operands are pre-halved at production (A, B, tables, w1r, W2, b2 carry a factor 1/2), because
// SiLU(x) = h + h * tanh(h) with h = x/2: two MUFU (tanh.approx.f16x2 has no packed form) and one HFMA2 per element pair.
// The last layer without the energy head can also run as ONE launch with the coordinate head (k_last_fused, last_ring.cuh).
// Build switches (-D...): EWS_TIMING (per-role wait cycles), EWS_EXP (diagnostic, wrong results), EWS_*_REGS, EWS_SLEEP_*,
// EWS_FOLD, EWS_BIAS_MMA, EWS_PIN_ADDR, EWS_BULK_W (profiles/r01/README.md lists what each measured); EWS_RELAY, EWS_NLOAD,
// EWS_LOAD_BYKB, EWS_WAIT_HINT (round 2, profiles/r02/README.md and DESIGN.md section 4).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "common.cuh"
#include "tma.cuh"
#include "last_ring.cuh"
#include "coord_body.cuh"

namespace ews {

constexpr int TILE_M = 128;
constexpr uint32_t W_BYTES = 256 * 256 * 2;          // 131072
constexpr uint32_t S_BYTES = TILE_M * 256 * 2;       // 65536
constexpr uint32_t W_KBLK = 256 * 128;               // bytes per 64-wide K block of the weight image
constexpr uint32_t S_KBLK = TILE_M * 128;
constexpr uint32_t OFF_W = 0;
constexpr uint32_t OFF_S = W_BYTES;
constexpr uint32_t OFF_VEC = OFF_S + S_BYTES;        // (1 KB spare), w1r' [256] half at + 1024, (512 B spare)
constexpr uint32_t OFF_PART = OFF_VEC + 2048;        // [<=4 column groups][128 rows] float gate partials
constexpr uint32_t OFF_AGG = OFF_PART + 4 * 128 * 4; // [2 tile parities][4 lane quarters][256] float column sums
constexpr uint32_t OFF_META = OFF_AGG + 2 * 4 * 256 * 4; // [16 producer warps][2 slots][8 rows] int4 edge metadata
constexpr uint32_t OFF_VEC32 = OFF_META + 16 * 2 * 8 * 16; // fragment-ordered b2/2 and wa as half2: [2 halves][4 t%4][20-word rows] each
constexpr uint32_t OFF_BAR = OFF_VEC32 + 2048;            // 16 mbarriers + tmem base
// b2 folded into the accumulator by one extra K = 16 MMA per tile: A = [128 x 16] of 1/16, B = [256 x 16] with row n = b2[n]/2
// (both K-major, no swizzle: 8-row x 16-byte core matrices, K chunks 128 B apart, 8-row groups 256 B apart)
constexpr uint32_t OFF_BIASA = (OFF_BAR + 192 + 127) & ~127u;
constexpr uint32_t OFF_BIASB = OFF_BIASA + 128 * 32;
constexpr uint32_t SMEM_BYTES = OFF_BIASB + 256 * 32;
static_assert(SMEM_BYTES + 1024 <= 232448, "shared memory budget");
constexpr uint32_t SMEM_ALLOC = SMEM_BYTES + 1024;   // slack for the manual 1024-byte alignment

constexpr int NPROD = 16;                // producer warps
constexpr int NEPI = 8;                  // epilogue warps
// 28 warps: 16 producers, 8 epilogue, 1 MMA issuer + 2 loaders + 1 relay warp; the last four form one warpgroup (setmaxnreg is a
// warpgroup-wide operation: a lone 17th warp never finishes it and the epilogue's .inc then blocks forever).
constexpr int NT = (NPROD + NEPI + 4) * 32;   // 896 -> 72 registers/thread at launch, pool 28*32*72 = 64512
#ifndef EWS_PROD_REGS
#define EWS_PROD_REGS 72
#endif
#ifndef EWS_EPI_REGS
#define EWS_EPI_REGS 88
#endif
#ifndef EWS_MMA_REGS
#define EWS_MMA_REGS 40
#endif
constexpr int PROD_REGS = EWS_PROD_REGS;  // setmaxnreg: 16*32*72 + 8*32*88 + 4*32*40 = 64512 = the launch pool (28 warps x 32 x 72)
constexpr int EPI_REGS = EWS_EPI_REGS;
constexpr int MMA_REGS = EWS_MMA_REGS;
static_assert(NPROD * PROD_REGS + NEPI * EPI_REGS + 4 * MMA_REGS <= (NPROD + NEPI + 4) * 72, "register pool exceeded");
#ifndef EWS_EXP
#define EWS_EXP 0      // diagnostic experiments (wrong results): 2 = table rows from row 0, 4 = no tanh, 8 = half of the B_j gathers, 16 = no spill stores
#endif
#ifndef EWS_TIMING
#define EWS_TIMING 0   // 1: accumulate cycles spent in barrier waits per role into Params::timing (diagnostic builds)
#endif
#ifndef EWS_PIN_ADDR
#define EWS_PIN_ADDR 1
#endif
#ifndef EWS_BIAS_MMA
#define EWS_BIAS_MMA 1    // b2 enters the accumulator through the tensor core instead of 64 HADD2 per epilogue thread and tile
#endif
#ifndef EWS_BIAS_DESC
#define EWS_BIAS_DESC 0   // 0: LBO = 128 B (K chunk), SBO = 256 B (8-row group); 1: swapped (descriptor bring-up switch)
#endif
// back-off (ns) between mbarrier polls per role: producers, MMA issuer, B_j loaders, epilogue
#ifndef EWS_SLEEP_P
#define EWS_SLEEP_P 100
#endif
#ifndef EWS_SLEEP_M
#define EWS_SLEEP_M 0
#endif
#ifndef EWS_SLEEP_L
#define EWS_SLEEP_L 100
#endif
#ifndef EWS_SLEEP_E
#define EWS_SLEEP_E 200
#endif
#ifndef EWS_LOAD_BYKB
#define EWS_LOAD_BYKB 0    // 1: loader warp lw takes whole K blocks lw, lw + EWS_NLOAD, ..; 0: every loader warp takes a share of the rows of every K block
#endif
#ifndef EWS_RELAY
#define EWS_RELAY 1        // 1: two loader warps issue the gathers, the third waits for their transaction bytes and hands every K block to
#endif                     // the 16 producer warps through a one-arrival barrier (the producers then wake once per K block instead of at
                           // every partial completion of the 32 gathers)
#ifndef EWS_NLOAD
#define EWS_NLOAD (EWS_RELAY ? 2 : 3)   // B_j loader warps that issue gathers: the 32 gather units of a K block are split over them
#endif
#ifndef EWS_BULK_W
#define EWS_BULK_W 1      // weight image by cp.async.bulk (TMA 1-D), overlapped with the rest of the set-up and the first tile's build
#endif
#ifndef EWS_SKIP_PAD
#define EWS_SKIP_PAD 1    // a producer warp whose second 4-row group lies entirely in the pad slots K..63 (warps 7 and 15 at K = 60)
#endif                    // skips it; those rows keep the raw B_j the loaders staged (finite; their gate is 0)
#ifndef EWS_TABLE_NOALLOC
#define EWS_TABLE_NOALLOC 0   // 1: the producers' table-row gathers bypass L1 allocation (ld.global.nc.L1::no_allocate)
#endif
#ifndef EWS_WARP_ARRIVE
#define EWS_WARP_ARRIVE 0   // 1: one mbarrier arrival per producer / epilogue WARP (fence or tcgen05 fence by every lane, __syncwarp, lane 0 arrives)
#endif                      // instead of one per thread: 16 + 8 instead of 512 + 256 updates of the same barrier word per K block / tile
#ifndef EWS_FOLD
#define EWS_FOLD 16       // gate-logit products accumulated in half2 before they are folded to fp32: 4 (every column group), 8, 16 or 32
#endif

// kind::f16 instruction descriptor: D=f32 (bit 4), A=B=f16 (0), both K-major, N=256 (>>3 at bit 17), M=128 (>>4 at bit 24)
constexpr uint32_t IDESC = (1u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  // K-major, SWIZZLE_128B: start>>4 | LBO(ignored)=1 | SBO = 1024 B (8 rows x 128 B) | version 1 | layout 2
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint64_t make_desc_noswz(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // K-major, no swizzle (layout type 0): start>>4 | LBO>>4 at bit 16 | SBO>>4 at bit 32 | version 1
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrival of a whole warp whose lanes have each fenced their own writes / TMEM reads
__device__ __forceinline__ void role_arrive(uint32_t bar, int lane) {
#if EWS_WARP_ARRIVE
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
#else
  mbar_arrive(bar);
#endif
}
// try_wait parks the thread inside the instruction for a hardware-bounded time; between retries back off so that
// waiting warps do not eat the issue slots of the working warps.  A lost arrival traps instead of hanging the GPU.
#if EWS_TIMING
#define TWAIT(acc, call) do { const long long _t0 = clock64(); call; acc += (unsigned long long)(clock64() - _t0); } while (0)
#else
#define TWAIT(acc, call) do { call; } while (0)
#endif
#ifndef EWS_WAIT_HINT
#define EWS_WAIT_HINT 1   // 1: retries park inside try_wait with a suspend-time hint (no poll instructions while parked); 0: nanosleep back-off
#endif
template <int SLEEP_NS = 40>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok, tries = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  while (!ok) {
#if EWS_WAIT_HINT
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
    if (++tries > (1u << 20)) __trap();
#else
    if (SLEEP_NS > 0) __nanosleep(SLEEP_NS);
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (++tries > (1u << 24)) __trap();
#endif
  }
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 16 lanes x 256 bit x 2: lane t gets, for the two 8-column blocks b = 0, 1 at [taddr]:
//   r[4b + 0..1] = (row t/4,     columns 8b + 2(t%4) + {0,1}),  r[4b + 2..3] = (row t/4 + 8, same columns)
// (layout measured with profiles/probes/tmem_ld_shapes.cu)
__device__ __forceinline__ void tmem_ld16x256_x2_issue(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
// wait for all outstanding tcgen05.ld; the registers of the chunk about to be consumed are tied to the wait so that no
// use of them can be scheduled above it
__device__ __forceinline__ void tmem_ld_wait8(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
               :: "memory");
}

// ---- shared-memory accessors on 32-bit shared-space addresses (LDS/STS instead of generic LD/ST) ----------
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v; asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v;
}
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
  uint32_t v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a)); return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ float ldsf(uint32_t a) { return __uint_as_float(lds32(a)); }
__device__ __forceinline__ void stsf(uint32_t a, float v) { sts32(a, __float_as_uint(v)); }

__device__ __forceinline__ uint4 ldg_table(const void* p) {
#if EWS_TABLE_NOALLOC
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
#else
  return __ldg(reinterpret_cast<const uint4*>(p));
#endif
}
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
__device__ __forceinline__ uint32_t h2silu(uint32_t h) { return (EWS_EXP & 4) ? h2fma(h, h, h) : h2fma(h, h2tanh(h), h); }
// diagnostic split of EWS_EXP 4: 32 = no tanh in the producers only, 64 = no tanh in the epilogue only
__device__ __forceinline__ uint32_t h2silu_p(uint32_t h) { return (EWS_EXP & 32) ? h2fma(h, h, h) : h2silu(h); }
__device__ __forceinline__ uint32_t h2silu_e(uint32_t h) { return (EWS_EXP & 64) ? h2fma(h, h, h) : h2silu(h); }

struct Params {
  int ntiles;
  int chunk;                // tiles per CTA: CTA c owns the contiguous tiles [c * chunk, (c + 1) * chunk)
  int total_nodes;          // B * N
  int N, R, K;
  int last;                 // spill gated messages of ligand residues (coordinate head input)
  int lig_only;             // last layer without the energy head: only tiles that hold a ligand residue are processed (the
  int tpt;                  // receptor rows' segment sums would feed a node update that is never run); tpt = such tiles per trajectory
  const __half* Wimg;       // (W2 / 2) fp16 SW128 image
  const int4* emeta;        // [B*N, 64] {global row of j, Tdrp row, Totp row or -1, radial bits}
  const __half* Ahi;        // [B*N, 256] fp16((W1s h_i + b1)/2)
  const __half* Bm;         // [B*N, 256] fp16((W1d h_j)/2)
  const __half* Tdrp;       // pre-halved merged tables
  const __half* Totp;
  const float* w1r;         // [256] fp32 (unhalved)
  const float* b2;          // [256]
  const float* wa;          // [256]
  const float* ba;          // [1]
  __half* agg16;            // [B*N, 256] fp16(agg x 2^-6) out
  __half* mstar;            // [B*L, 64, 256] fp16 (m* x 2^-6), last layer only
  unsigned long long* timing;   // EWS_TIMING: [8] cycles {prod wait bfull, prod total, loader wait empty, loader total,
                                //                        mma wait full, mma wait acce, epi wait accf, epi total}
};

constexpr float RAD_SCALE = 0.03125f;     // radial is carried as fp16(radial / 32); w1r' = 32 * w1r / 2
constexpr float MSTAR_SCALE = 0.015625f;  // gated messages are carried x 2^-6 in fp16 (column sums stay < 65504)


// NR packed registers summed over the lanes that differ in the bits OHI .. OLO (powers of two, OHI >= OLO), halving the
// register count at every step: with NR = 16, OHI = 16, OLO = 4 lane l ends with v[0], v[1] = sums of registers
// 2 (l >> 2) and 2 (l >> 2) + 1 over the 8 lanes that share l & 3
template <int NR, int OHI, int OLO>
__device__ __forceinline__ void lane_group_sum_h2(uint32_t* v, int lane) {
#pragma unroll
  for (int o = OHI, n = NR; o >= OLO; o >>= 1, n >>= 1) {
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

// LAST: the last E_GCL layer (spill of the ligand rows' gated messages for the coordinate head, optional ligand-only tile walk);
// a separate instantiation so that the five other launches carry neither its code nor its registers
// FUSED: the edge role of k_last_fused (LAST, ligand-only walk): the gated messages of a tile go to the ring of
// last_ring.cuh instead of the spill buffer, `cta` is the rank among the ring.P producer CTAs
template <bool LAST, bool FUSED>
__device__ __forceinline__ void edge_body(const Params& p, const CUtensorMap& tmB, const LastRing& ring, uint8_t* smem, const int cta) {
  const uint32_t sbase = smem_u32(smem);
  __half* vwr = reinterpret_cast<__half*>(smem + OFF_VEC) + 512;   // 16 * w1r
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 144);
  const uint32_t bar_bfull = sbase + OFF_BAR + 96;      // [4] B_j rows of a K block have landed in the S tile
  const uint32_t bar_full = sbase + OFF_BAR;            // [4]
  const uint32_t bar_empty = sbase + OFF_BAR + 32;      // [4]
  const uint32_t bar_accf = sbase + OFF_BAR + 64;       // [2]
  const uint32_t bar_acce = sbase + OFF_BAR + 80;       // [2]
  const uint32_t bar_w = sbase + OFF_BAR + 128;         // weight image landed (bulk copy)
  const uint32_t bar_bready = sbase + OFF_BAR + 160;    // [4] EWS_RELAY: one arrival per K block once all its B_j rows have landed
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time setup (constants only: nothing a predecessor kernel wrote is read before pdl_wait) -----
  {
#if !EWS_BULK_W
    const uint4* src = reinterpret_cast<const uint4*>(p.Wimg);
    uint4* dst = reinterpret_cast<uint4*>(smem + OFF_W);
    for (int i = tid; i < (int)(W_BYTES / 16); i += NT) dst[i] = __ldg(src + i);
#endif
    if (tid < 256) {
      vwr[tid] = __float2half_rn(p.w1r[tid] * (0.5f / RAD_SCALE));
      if (tid < 128) {
        // fragment order of the 16x256b epilogue: [column half][t % 4][8-column block j] -> column pair ch*64 + 4 j + (t%4)
        const int pr = (tid >> 6) * 64 + 4 * (tid & 15) + ((tid >> 4) & 3);
        const __half2 b = __floats2half2_rn(0.5f * p.b2[2 * pr], 0.5f * p.b2[2 * pr + 1]);
        const __half2 w = __floats2half2_rn(p.wa[2 * pr], p.wa[2 * pr + 1]);
        // 20-word stride per (ch, t%4) row: the four rows a quarter-warp reads land in disjoint bank groups
        const int slot = (tid >> 4) * 20 + (tid & 15);
        reinterpret_cast<__half2*>(smem + OFF_VEC32)[slot] = b;
        reinterpret_cast<__half2*>(smem + OFF_VEC32)[160 + slot] = w;
      }
    }
  }
#if EWS_BIAS_MMA
  {
    uint32_t* ones = reinterpret_cast<uint32_t*>(smem + OFF_BIASA);
    for (int i = tid; i < 128 * 32 / 4; i += NT) ones[i] = 0x2C002C00u;             // half2(1/16, 1/16)
    // row n, K chunk c (8 halfs): byte offset (n >> 3) * 256 + c * 128 + (n & 7) * 16
    for (int i = tid; i < 256 * 2; i += NT) {
      const int n = i >> 1, c = i & 1;
      const __half2 b = __half2half2(__float2half_rn(0.5f * p.b2[n]));
      const uint32_t bits = *reinterpret_cast<const uint32_t*>(&b);
      *reinterpret_cast<uint4*>(smem + OFF_BIASB + (n >> 3) * 256 + c * 128 + (n & 7) * 16) = make_uint4(bits, bits, bits, bits);
    }
  }
#endif
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) { mbar_init(bar_full + 8 * i, EWS_WARP_ARRIVE ? NPROD : NPROD * 32); mbar_init(bar_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_accf + 8 * i, 1); mbar_init(bar_acce + 8 * i, EWS_WARP_ARRIVE ? NEPI : NEPI * 32); }
    for (int i = 0; i < 4; ++i) { mbar_init(bar_bfull + 8 * i, EWS_LOAD_BYKB ? 1 : EWS_NLOAD); mbar_init(bar_bready + 8 * i, 1); }
    tma_prefetch_desc(&tmB);
#if EWS_BULK_W
    mbar_init(bar_w, 1);
#endif
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#if EWS_BULK_W
    // the 128 KB weight image arrives by four bulk copies (async proxy) while the CTA finishes its set-up and the
    // producers / loaders already work on the first tile; only the MMA issuer waits for it
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_w), "r"(W_BYTES) : "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(sbase + OFF_W + (uint32_t)i * (W_BYTES / 4)), "l"(reinterpret_cast<const char*>(p.Wimg) + (size_t)i * (W_BYTES / 4)),
                     "r"(W_BYTES / 4), "r"(bar_w) : "memory");
#endif
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
  pdl_wait();               // Ah / Bm / emeta of the preceding kernels are complete and visible from here on
  // contiguous tile ranges: concurrently running CTAs work on different trajectories, so the gathered B_j rows are
  // not hot lines shared by all SMs (strided assignment had every SM hammer the same 300 rows at the same time)
  const int t_begin = cta * p.chunk;
  const int t_end = min(p.ntiles, t_begin + p.chunk);
  // logical tile index -> tile of the [B*N/2] node-pair grid.  Ligand-only launches walk, per trajectory, the tiles
  // (b N + R) / 2 .. (b N + N - 1) / 2; when that count is one short of tpt (odd N) the last tile is simply done twice.
  auto phys = [&](int lt) -> int {
    if (!LAST || !p.lig_only) return lt;
    const int b = lt / p.tpt, r = lt - b * p.tpt;
    const int base = b * p.N;
    return min(((base + p.R) >> 1) + r, (base + p.N - 1) >> 1);
  };

  if (warp < NPROD) {
    // =================================== PRODUCERS ===================================================
    // rows of this lane: ring index idx = (lane >> 3) + 4 i, i = 0..1 -> tile row prow(idx), all in residue `warp >> 3`;
    // 16-byte chunk c8 = lane & 7 of each 128-byte K-block row.  Work item = (K block, pair of rows); the gathers of
    // item n+1 are in flight while item n is computed (two register buffers), across K blocks and across tiles.
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PROD_REGS));
    const int c8 = lane & 7, rsub = lane >> 3;
    auto prow = [&](int idx) -> int { return warp * 8 + idx; };
    const int4 pad_meta = make_int4(0, 40 * 66 + 32, -1, 0);
    // B_j is already in the S tile (loader warps, cp.async); this role gathers the two table rows of each edge into
    // registers one K block ahead (two buffers), forms u/2, applies SiLU and overwrites the 16-byte chunk in place.
    struct GBuf { uint4 td[2], to[2]; uint4 a; };
    uint32_t mring_s = sbase + OFF_META + (uint32_t)warp * 256u;      // shared-space address of this warp's ring
    uint32_t vwr_s = sbase + OFF_VEC + 1024u + (uint32_t)c8 * 16u;
    // swizzled byte offset of this lane's chunk in row r0 + 4 i of a K block: (r0 + 4 i) * 128 + ((c8 ^ (r & 7)) << 4)
    uint32_t soff[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = prow(rsub + 4 * i);
      soff[i] = sbase + OFF_S + (uint32_t)r * 128u + (uint32_t)((c8 ^ (r & 7)) << 4);
    }
    uint32_t pbar = sbase + OFF_BAR;          // this role's copy of the barrier block address (full at +0, bfull at +96)
#if EWS_PIN_ADDR
    // opaque to the compiler: keeps these addresses in registers instead of re-deriving them from %tid and the
    // shared window base in every K block (ptxas otherwise rematerialises ~25 instructions per K block)
    asm volatile("" : "+r"(soff[0]), "+r"(soff[1]), "+r"(mring_s), "+r"(vwr_s), "+r"(pbar));
#endif
    // per-lane views: this lane's rows of a metadata slot start at + rsub * 16; table / A pointers already at column c8 * 8
    uint32_t mlane = (uint32_t)rsub * 16u;
    const __half* tdrp_l = p.Tdrp + c8 * 8;
    const __half* totp_l = p.Totp + c8 * 8;
#if EWS_PIN_ADDR
    asm volatile("" : "+r"(mlane), "+l"(tdrp_l), "+l"(totp_l));
#endif
    const bool full = !EWS_SKIP_PAD || (warp & 7) * 8 + 4 < p.K;     // second 4-row group holds real edge slots
    auto issue = [&](GBuf& g, uint32_t mslot, size_t aoff, int kb) {
      g.a = __ldg(reinterpret_cast<const uint4*>(p.Ahi + aoff + kb * 64));
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (i == 1 && !full) break;
        uint4 mt = lds128(mslot + mlane + (uint32_t)(4 * i) * 16u);
        if (EWS_EXP & 2) { mt.y = 0; mt.z = ((int)mt.z >= 0) ? 0u : mt.z; }
        g.td[i] = ldg_table(tdrp_l + (size_t)mt.y * H + kb * 64);
        g.to[i] = make_uint4(0, 0, 0, 0);
        if ((int)mt.z >= 0) g.to[i] = ldg_table(totp_l + (size_t)mt.z * H + kb * 64);
      }
    };
    auto compute = [&](const GBuf& g, uint32_t mslot, int kb) {
      const uint4 wr = lds128(vwr_s + (uint32_t)kb * 128u);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (i == 1 && !full) break;
        const uint32_t rad = lds32(mslot + mlane + (uint32_t)(4 * i) * 16u + 12u);
        const uint32_t sa = soff[i] + (uint32_t)kb * S_KBLK;
        const uint4 hb = lds128(sa);
        uint4 o;
        o.x = h2silu_p(h2add(h2add(g.a.x, hb.x), h2fma(rad, wr.x, h2add(g.td[i].x, g.to[i].x))));
        o.y = h2silu_p(h2add(h2add(g.a.y, hb.y), h2fma(rad, wr.y, h2add(g.td[i].y, g.to[i].y))));
        o.z = h2silu_p(h2add(h2add(g.a.z, hb.z), h2fma(rad, wr.z, h2add(g.td[i].z, g.to[i].z))));
        o.w = h2silu_p(h2add(h2add(g.a.w, hb.w), h2fma(rad, wr.w, h2add(g.td[i].w, g.to[i].w))));
        sts128(sa, o);
      }
    };
    auto load_meta = [&](int lt) -> int4 {       // lanes 0..7: edge metadata of row warp*8 + lane of logical tile `lt`
      int4 mt = pad_meta;
      const int r = prow(lane & 7);
      const int node = phys(lt) * 2 + (r >> 6);
      if (lt < t_end && node < p.total_nodes) mt = __ldg(p.emeta + (size_t)node * SLOTS + (r & 63));
      return mt;
    };
    auto a_node = [&](int lt) -> size_t {
      int node = phys(lt) * 2 + (warp >> 3);
      if (node >= p.total_nodes) node = p.total_nodes - 1;   // odd tail: those rows are masked in the epilogue
      return (size_t)node * H + c8 * 8;
    };
    GBuf g0, g1;
    unsigned long long tw0 = 0;
    const long long tstart = clock64();
    if (t_begin < t_end) {
      const int4 m0 = load_meta(t_begin);
      if (lane < 8) sts128(mring_s + (uint32_t)lane * 16u, make_uint4(m0.x, m0.y, m0.z, m0.w));
      __syncwarp();
      issue(g0, mring_s, a_node(t_begin), 0);
    }
    int it = 0;
    for (int tile = t_begin; tile < t_end; ++tile, ++it) {
      const uint32_t mcur = mring_s + (uint32_t)(it & 1) * 128u;
      const uint32_t mnext = mring_s + (uint32_t)((it & 1) ^ 1) * 128u;
      const int ntile = tile + 1;
      const bool has_next = ntile < t_end;
      const int4 nm = load_meta(ntile);
      const size_t ao = a_node(tile), aon = a_node(has_next ? ntile : tile);
      const uint32_t par = (uint32_t)(it & 1);
      // kb 0 (buffer 0); prefetch kb 1
      issue(g1, mcur, ao, 1);
      TWAIT(tw0, mbar_wait<EWS_SLEEP_P>(pbar + (EWS_RELAY ? 160u : 96u), par));
      compute(g0, mcur, 0);
      fence_async_smem(); role_arrive(pbar + 0u, lane);
      // kb 1 (buffer 1); prefetch kb 2
      issue(g0, mcur, ao, 2);
      TWAIT(tw0, mbar_wait<EWS_SLEEP_P>(pbar + (EWS_RELAY ? 168u : 104u), par));
      compute(g1, mcur, 1);
      fence_async_smem(); role_arrive(pbar + 8u, lane);
      // kb 2 (buffer 0); prefetch kb 3
      issue(g1, mcur, ao, 3);
      TWAIT(tw0, mbar_wait<EWS_SLEEP_P>(pbar + (EWS_RELAY ? 176u : 112u), par));
      compute(g0, mcur, 2);
      fence_async_smem(); role_arrive(pbar + 16u, lane);
      // kb 3 (buffer 1); stage the next tile's metadata (its load has had three K blocks to land), prefetch its kb 0
      if (lane < 8) sts128(mnext + (uint32_t)lane * 16u, make_uint4(nm.x, nm.y, nm.z, nm.w));
      __syncwarp();
      if (has_next) issue(g0, mnext, aon, 0);
      TWAIT(tw0, mbar_wait<EWS_SLEEP_P>(pbar + (EWS_RELAY ? 184u : 120u), par));
      compute(g1, mcur, 3);
      fence_async_smem(); role_arrive(pbar + 24u, lane);
    }
    if (EWS_TIMING && warp == 0 && lane == 0) { atomicAdd(p.timing + 0, tw0); atomicAdd(p.timing + 1, (unsigned long long)(clock64() - tstart)); }
  } else if (warp >= NPROD + NEPI) {
    // =================================== MMA ISSUER + B_j LOADERS ======================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(MMA_REGS));
    if (warp == NPROD + NEPI) {
      if (lane == 0) {
        const uint64_t dW = make_desc(sbase + OFF_W);
        const uint64_t dS = make_desc(sbase + OFF_S);
#if EWS_BIAS_MMA
        const uint64_t dOnes = make_desc_noswz(sbase + OFF_BIASA, EWS_BIAS_DESC ? 256u : 128u, EWS_BIAS_DESC ? 128u : 256u);
        const uint64_t dBias = make_desc_noswz(sbase + OFF_BIASB, EWS_BIAS_DESC ? 256u : 128u, EWS_BIAS_DESC ? 128u : 256u);
#endif
        unsigned long long tw0 = 0, tw1 = 0;
        int it = 0;
#if EWS_BULK_W
        mbar_wait<20>(bar_w, 0u);
#endif
        for (int tile = t_begin; tile < t_end; ++tile, ++it) {
          const int buf = it & 1;
          if (it >= 2) TWAIT(tw1, mbar_wait<EWS_SLEEP_M>(bar_acce + 8 * buf, (uint32_t)(((it >> 1) - 1) & 1)));   // epilogue drained this buffer
          const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 256);
#if EWS_BIAS_MMA
          tc_fence_after();
          mma_f16(d_tmem, dOnes, dBias, 0u);      // D = 1 * (b2/2)^T: needs no producer, so it goes first
#endif
#pragma unroll 1
          for (int kb = 0; kb < 4; ++kb) {
            TWAIT(tw0, mbar_wait<EWS_SLEEP_M>(bar_full + 8 * kb, (uint32_t)(it & 1)));
            tc_fence_after();
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const uint64_t da = dS + (uint64_t)((kb * S_KBLK + k4 * 32) >> 4);
              const uint64_t db = dW + (uint64_t)((kb * W_KBLK + k4 * 32) >> 4);
              mma_f16(d_tmem, da, db, (EWS_BIAS_MMA || (kb | k4)) ? 1u : 0u);
            }
            mma_commit(bar_empty + 8 * kb);
          }
          mma_commit(bar_accf + 8 * buf);
        }
        if (EWS_TIMING) { atomicAdd(p.timing + 4, tw0); atomicAdd(p.timing + 5, tw1); }
      }
    } else if (warp <= NPROD + NEPI + EWS_NLOAD) {
      // Loader warps bring the B_j rows of every K block of every tile straight into the S tile with TMA: one
      // cp.async.bulk.tensor.2d tile::gather4 per four tile rows (four neighbour rows of Bm, 64 columns each,
      // SWIZZLE_128B), completion by transaction bytes on bar_bfull.  No registers are held while the rows are in flight and
      // nothing passes through the LSU / L1 pipe; the producers add the rest in place.  A block may be refilled as soon as
      // the MMA of the previous tile has consumed it (bar_empty).  The 32 four-row units of a K block are split over the
      // EWS_NLOAD loader warps (the per-lane TMA issue is serialised by the hardware's uniform-operand rule, so the split
      // shortens the refill latency of a block, which sits on the build -> MMA -> refill -> build cycle of the S tile).
      const int lw = warp - (NPROD + NEPI + 1);
      constexpr int UPW = EWS_LOAD_BYKB ? 32 : (32 + EWS_NLOAD - 1) / EWS_NLOAD;          // units per loader warp
      const int unit = EWS_LOAD_BYKB ? lane : lw * UPW + lane;                              // this lane's four-row unit (tile rows 4 unit ..)
      const bool act = lane < UPW && unit < 32 && !((EWS_EXP & 8) && (unit & 1));
      const uint32_t nact = (EWS_EXP & 8) ? (uint32_t)__popc(__ballot_sync(0xffffffffu, act)) : (uint32_t)(EWS_LOAD_BYKB ? 32 : min(UPW, 32 - lw * UPW));   // active lanes of this warp
      auto load_j4 = [&](int ptile, bool inrange) -> int4 {   // global rows of the neighbours of tile rows 4 unit .. 4 unit + 3
        const int node = ptile * 2 + (unit >> 4);
        int4 j = make_int4(0, 0, 0, 0);
        if (act && inrange && node < p.total_nodes) {
          const int* e = reinterpret_cast<const int*>(p.emeta + (size_t)node * SLOTS + ((4 * unit) & 63));
          j.x = __ldg(e); j.y = __ldg(e + 4); j.z = __ldg(e + 8); j.w = __ldg(e + 12);
        }
        return j;
      };
      int4 jc = make_int4(0, 0, 0, 0);
      if (t_begin < t_end) jc = load_j4(phys(t_begin), true);
      const uint32_t dst_lane = sbase + OFF_S + (uint32_t)unit * 512u;
      unsigned long long tw0 = 0;
      const long long tstart = clock64();
      int it = 0;
      for (int tile = t_begin; tile < t_end; ++tile, ++it) {
        const bool nin = tile + 1 < t_end;
        const int4 jn = load_j4(nin ? phys(tile + 1) : 0, nin);
#pragma unroll 1
        for (int kb = EWS_LOAD_BYKB ? lw : 0; kb < 4; kb += EWS_LOAD_BYKB ? EWS_NLOAD : 1) {
          if (it > 0) TWAIT(tw0, mbar_wait<EWS_SLEEP_L>(bar_empty + 8 * kb, (uint32_t)((it - 1) & 1)));
          if (lane == 0) mbar_expect_tx(bar_bfull + 8 * kb, nact * 512u);
          __syncwarp();
          if (act) tma_gather4_2d(dst_lane + (uint32_t)kb * S_KBLK, &tmB, kb * 64, jc.x, jc.y, jc.z, jc.w, bar_bfull + 8 * kb);
        }
        jc = jn;
      }
      if (EWS_TIMING && lw == 0 && lane == 0) { atomicAdd(p.timing + 2, tw0); atomicAdd(p.timing + 3, (unsigned long long)(clock64() - tstart)); }
    }
#if EWS_RELAY
    else if (warp == NPROD + NEPI + EWS_NLOAD + 1) {
      // relay: bar_bfull completes by transaction bytes, and every partial completion wakes the threads parked on it; this
      // warp absorbs those wake-ups and arrives once on bar_bready, the barrier the 512 producer threads wait on
      int it = 0;
      for (int tile = t_begin; tile < t_end; ++tile, ++it) {
#pragma unroll 1
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait<0>(bar_bfull + 8 * kb, (uint32_t)(it & 1));
          if (lane == 0) mbar_arrive(bar_bready + 8 * kb);
        }
      }
    }
#endif
    __syncwarp();
  } else {
    // =================================== EPILOGUE =====================================================
    // 8 warps = 4 TMEM lane quarters (rows) x 2 column halves; warp (q, ch) owns rows q*32.. and columns ch*128..
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(EPI_REGS));
    const int e = warp - NPROD;
    const int q = warp & 3;            // TMEM lane quarter this warp may access (warp id % 4)
    const int ch = e >> 2;             // column half
    const int hn = q >> 1;             // residue of the tile this warp's rows belong to
    const int ecol = ((ch * 2 + (q & 1)) * 32 + lane) * 2;   // first of the 2 columns this thread writes in the final combine
    const float ba = p.ba[0];
    unsigned long long tw0 = 0;
    const long long tstart = clock64();
    int it = 0;
    // Fragment layout: lane t = (rg = t >> 2, cq = t & 3) holds rows rr(k) = q*32 + rg + 8 k (k = 0..3) and, for every
    // 8-column block j = 0..15 of this warp's column half, the column pair 8 j + 2 cq + {0, 1}: m[k * 16 + j].
    const int rg = lane >> 2, cq = lane & 3;
    const uint32_t vx_s = sbase + OFF_VEC32 + (uint32_t)((ch * 4 + cq) * 20) * 4u;
    const uint32_t part_row_s = sbase + OFF_PART + (uint32_t)(q * 32 + rg) * 4u;      // + 32 k bytes for row k
    for (int tile = t_begin; tile < t_end; ++tile, ++it) {
      const int buf = it & 1;
      const int node = phys(tile) * 2 + hn;
      TWAIT(tw0, mbar_wait<EWS_SLEEP_E>(bar_accf + 8 * buf, (uint32_t)((it >> 1) & 1)));
      tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 256 + ch * 128);
      uint32_t m[64];
      float dot[4] = {0.f, 0.f, 0.f, 0.f};
      {
        uint32_t accA[8], accB[8];
        uint32_t dA[2] = {0, 0}, dB[2] = {0, 0};
        uint4 bb = make_uint4(0, 0, 0, 0), ww = make_uint4(0, 0, 0, 0);
        // iteration n = g * 4 + hh * 2 + half: column group g (blocks 4 g .. 4 g + 3: one LDS.128 of b2 / wa serves both
        // row halves), row half hh (rows k = 2 hh, 2 hh + 1), blocks 2 jj, 2 jj + 1 with jj = 2 g + half
        auto frag_addr = [&](int n) -> uint32_t {
          const int g = n >> 2, hh = (n >> 1) & 1, jj = 2 * g + (n & 1);
          return tbase + ((uint32_t)(hh * 16) << 16) + (uint32_t)(jj * 16);
        };
        tmem_ld16x256_x2_issue(frag_addr(0), accA);
#pragma unroll
        for (int n = 0; n < 16; ++n) {
          const int g = n >> 2, hh = (n >> 1) & 1, half = n & 1, jj = 2 * g + half;
          uint32_t* cur = (n & 1) ? accB : accA;
          uint32_t* nxt = (n & 1) ? accA : accB;
          tmem_ld_wait8(cur);
          if (n + 1 < 16) tmem_ld16x256_x2_issue(frag_addr(n + 1), nxt);
          if ((n & 3) == 0) {
            if (!EWS_BIAS_MMA) bb = lds128(vx_s + (uint32_t)g * 16u);
            ww = lds128(vx_s + 640u + (uint32_t)g * 16u);
          }
          const uint32_t w0 = half ? ww.z : ww.x, w1 = half ? ww.w : ww.y;
#if EWS_BIAS_MMA
          const uint32_t x00 = h2silu_e(f2h2(__uint_as_float(cur[0]), __uint_as_float(cur[1])));
          const uint32_t x10 = h2silu_e(f2h2(__uint_as_float(cur[2]), __uint_as_float(cur[3])));
          const uint32_t x01 = h2silu_e(f2h2(__uint_as_float(cur[4]), __uint_as_float(cur[5])));
          const uint32_t x11 = h2silu_e(f2h2(__uint_as_float(cur[6]), __uint_as_float(cur[7])));
#else
          const uint32_t b0 = half ? bb.z : bb.x, b1 = half ? bb.w : bb.y;
          const uint32_t x00 = h2silu_e(h2add(f2h2(__uint_as_float(cur[0]), __uint_as_float(cur[1])), b0));
          const uint32_t x10 = h2silu_e(h2add(f2h2(__uint_as_float(cur[2]), __uint_as_float(cur[3])), b0));
          const uint32_t x01 = h2silu_e(h2add(f2h2(__uint_as_float(cur[4]), __uint_as_float(cur[5])), b1));
          const uint32_t x11 = h2silu_e(h2add(f2h2(__uint_as_float(cur[6]), __uint_as_float(cur[7])), b1));
#endif
          m[(2 * hh) * 16 + 2 * jj] = x00; m[(2 * hh) * 16 + 2 * jj + 1] = x01;
          m[(2 * hh + 1) * 16 + 2 * jj] = x10; m[(2 * hh + 1) * 16 + 2 * jj + 1] = x11;
          dA[hh] = h2fma(x00, w0, dA[hh]); dA[hh] = h2fma(x01, w1, dA[hh]);
          dB[hh] = h2fma(x10, w0, dB[hh]); dB[hh] = h2fma(x11, w1, dB[hh]);
          if ((((g & (EWS_FOLD / 4 - 1)) == EWS_FOLD / 4 - 1) || g == 3) && half) {   // EWS_FOLD products per half2 lane, then out to fp32 (short fp16 chains keep the gate logit accurate)
            const float2 fa = h2f2(dA[hh]), fb = h2f2(dB[hh]);
            dot[2 * hh] += fa.x + fa.y; dot[2 * hh + 1] += fb.x + fb.y;
            dA[hh] = dB[hh] = 0;
          }
        }
      }
      tc_fence_before();
      role_arrive(bar_acce + 8 * buf, lane);    // accumulator buffer may be overwritten by tile it + 2
      // gate logit: sum over the 4 lanes of a row, then over the two column halves (shared memory)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        dot[k] += __shfl_xor_sync(0xffffffffu, dot[k], 1);
        dot[k] += __shfl_xor_sync(0xffffffffu, dot[k], 2);
      }
      if (cq == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) stsf(part_row_s + (uint32_t)ch * 512u + (uint32_t)k * 32u, dot[k]);
      }
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");       // the 2 column halves of this row quarter
      uint32_t gk[4];
      {
        // lane (rg, cq) evaluates the gate of row k = cq; the 4 lanes of a row group then exchange them
        const float mine = cq == 0 ? dot[0] : cq == 1 ? dot[1] : cq == 2 ? dot[2] : dot[3];
        const float other = ldsf(part_row_s + (uint32_t)(ch ^ 1) * 512u + (uint32_t)cq * 32u);
        const float tot = (ch == 0 ? mine + other : other + mine) + ba;
        const int slot = (q * 32 + rg + 8 * cq) & 63;
        const bool valid = node < p.total_nodes && slot < p.K;
        const float g = valid ? MSTAR_SCALE * __fdividef(1.f, 1.f + __expf(-tot)) : 0.f;
        const uint32_t g2 = f2h2(g, g);
#pragma unroll
        for (int k = 0; k < 4; ++k) gk[k] = __shfl_sync(0xffffffffu, g2, (lane & ~3) | k);
      }
      uint32_t sacc[16];
      if (LAST) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
          for (int j = 0; j < 16; ++j) m[k * 16 + j] = h2mul(m[k * 16 + j], gk[k]);
        }
        if (FUSED) {
          // hand the tile to the coordinate-head role through the L2-resident ring (last_ring.cuh); all 128 rows are
          // written (rows of receptor residues and pad slots are ignored / zero-gated on the other side)
          const unsigned sq = (unsigned)it * (unsigned)ring.P + (unsigned)cta;
          const unsigned rs = sq % (unsigned)ring.NR, epoch = sq / (unsigned)ring.NR;
          if (epoch > 0) {
            if (lane == 0) ring_wait_ge(ring.done + rs, epoch);
            __syncwarp();
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int rr = q * 32 + rg + 8 * k;
            __half* dst = ring.ring + (((size_t)rs * 4 + ch * 2) * TILE_M + rr) * 64 + cq * 16;      // [slot][K block][row][64]
#pragma unroll
            for (int v = 0; v < 2; ++v)
              asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + (size_t)v * (TILE_M * 64)),
                           "r"(m[k * 16 + 8 * v]), "r"(m[k * 16 + 8 * v + 1]), "r"(m[k * 16 + 8 * v + 2]), "r"(m[k * 16 + 8 * v + 3]),
                           "r"(m[k * 16 + 8 * v + 4]), "r"(m[k * 16 + 8 * v + 5]), "r"(m[k * 16 + 8 * v + 6]), "r"(m[k * 16 + 8 * v + 7]) : "memory");
          }
          fence_proxy_async_all();            // these generic-proxy writes are read by TMA on the other side
          __syncwarp();
          if (lane == 0) { __threadfence(); red_release_gpu_add(ring.ready + rs, 1u); }
          continue;
        }
        if (!(EWS_EXP & 16) && node < p.total_nodes) {
          const int b = node / p.N, i = node - b * p.N;
          if (i >= p.R) {
            // spill in fragment order and K-block-major per tile: [tile][K block ch*2+v][tile row][64], this lane's column
            // pairs 8 v .. 8 v + 7 of row k are the 32 contiguous bytes at position cq*16 of the row's K block, so the four
            // lanes of a row write one full 128-byte line per 256-bit store and the coordinate head fetches a K block of
            // a tile as ONE contiguous 16 KB box (the Wc1 image is K-permuted to match, node.cu k_image_pack_perm).  All
            // 64 slots are written -- the pad slots' gate is 0.
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int slot = (q * 32 + rg + 8 * k) & 63;
              const size_t ln = (size_t)b * (p.N - p.R) + (i - p.R);          // ligand residue -> tile ln / 2, tile row (ln & 1) * 64 + slot
              __half* dst = p.mstar + ((((ln >> 1) * 4 + ch * 2) * TILE_M) + (ln & 1) * 64 + slot) * 64 + cq * 16;
#pragma unroll
              for (int v = 0; v < 2; ++v)
                asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + (size_t)v * (TILE_M * 64)),
                             "r"(m[k * 16 + 8 * v]), "r"(m[k * 16 + 8 * v + 1]), "r"(m[k * 16 + 8 * v + 2]), "r"(m[k * 16 + 8 * v + 3]),
                             "r"(m[k * 16 + 8 * v + 4]), "r"(m[k * 16 + 8 * v + 5]), "r"(m[k * 16 + 8 * v + 6]), "r"(m[k * 16 + 8 * v + 7]) : "memory");
            }
          }
        }
        if (p.lig_only) continue;     // no node update follows this launch: the segment sum would never be read
#pragma unroll
        for (int j = 0; j < 16; ++j) sacc[j] = h2add(h2add(m[j], m[16 + j]), h2add(m[32 + j], m[48 + j]));
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          sacc[j] = h2fma(m[48 + j], gk[3], h2fma(m[32 + j], gk[2], h2fma(m[16 + j], gk[1], h2mul(m[j], gk[0]))));
      }
      lane_group_sum_h2<16, 16, 4>(sacc, lane);  // lane (rg, cq): blocks 2 rg, 2 rg + 1, column pair cq, summed over the 32 rows
      const uint32_t ag = sbase + OFF_AGG + (uint32_t)buf * 4096u;
      {
        const float2 s0 = h2f2(sacc[0]), s1 = h2f2(sacc[1]);
        const uint32_t a0 = ag + (uint32_t)(q * 256 + ch * 128 + 16 * rg + 2 * cq) * 4u;
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a0), "f"(s0.x), "f"(s0.y) : "memory");
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a0 + 32u), "f"(s1.x), "f"(s1.y) : "memory");
      }
      asm volatile("bar.sync %0, 128;" ::"r"(5 + hn) : "memory");     // the 4 warps that hold this residue's rows
      if (node < p.total_nodes) {
        const uint32_t a0 = ag + (uint32_t)((2 * hn) * 256 + ecol) * 4u;
        // agg stays x 2^-6 in fp16 (the W3a image carries the 2^6): operand of node_tc.cu MODE_Z
        *reinterpret_cast<uint32_t*>(p.agg16 + (size_t)node * H + ecol) = f2h2(ldsf(a0) + ldsf(a0 + 1024u), ldsf(a0 + 4u) + ldsf(a0 + 1028u));
      }
    }
    if (EWS_TIMING && e == 0 && lane == 0) { atomicAdd(p.timing + 6, tw0); atomicAdd(p.timing + 7, (unsigned long long)(clock64() - tstart)); }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NPROD + NEPI) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

template <bool LAST>
__global__ void __launch_bounds__(NT, 1) k_edge_ws(const Params p, const __grid_constant__ CUtensorMap tmB) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  pdl_trigger();            // the next kernel of the stream may start its prologue while this one drains (common.cuh)
  LastRing none{};
  edge_body<LAST, false>(p, tmB, none, smem, (int)blockIdx.x);
}

// Last layer without the energy head, edge MLP + coordinate head in ONE launch: CTAs [0, ring.P) run the edge role over the
// ligand residues' tiles, the others the coordinate head (coord_body.cuh) on the tiles the ring hands over.
__global__ void __launch_bounds__(NT, 1) k_last_fused(const Params p, const __grid_constant__ CUtensorMap tmB, const ntc::Params pc,
                                                      const __grid_constant__ CUtensorMap tmX, const LastRing ring) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  pdl_trigger();
  if ((int)blockIdx.x < ring.P) edge_body<true, true>(p, tmB, ring, smem, (int)blockIdx.x);
  else ntc::coord_body<true>(pc, tmX, ring, smem, (int)blockIdx.x - ring.P, (int)gridDim.x - ring.P);
}

}  // namespace ews

int launch_edge_ws(dfm_ctx* ctx, const EdgeArgs& a, const int4* emeta, const __half* Ahi, __half* agg16, cudaStream_t s) {
  const LayerW& w = ctx->layer[a.layer];
  ews::Params p{};
  p.total_nodes = a.B * a.N;
  p.ntiles = (p.total_nodes + 1) / 2;
  p.N = a.N; p.R = a.R; p.K = a.K;
  p.last = a.last ? 1 : 0;
  p.lig_only = (a.last && a.lig_only && a.N > a.R && a.R > 1) ? 1 : 0;
  if (p.lig_only) {
    const int L2a = ((a.N - 1) >> 1) - (a.R >> 1) + 1;                  // b N even
    const int L2b = (a.N >> 1) - ((a.R + 1) >> 1) + 1;                  // b N odd (only when N is odd)
    p.tpt = (a.N & 1) ? (L2a > L2b ? L2a : L2b) : L2a;
    p.ntiles = a.B * p.tpt;
  }
  p.Wimg = w.img_W2h;
  p.emeta = emeta;
  p.Ahi = Ahi;
  p.Bm = reinterpret_cast<const __half*>(a.Bm);
  p.Tdrp = w.Tdrp16h; p.Totp = w.Totp16h;
  p.w1r = w.w1r; p.b2 = w.b2; p.wa = w.wa; p.ba = w.ba;
  p.agg16 = agg16; p.mstar = a.mstar;
#if EWS_TIMING
  static unsigned long long* tbuf = nullptr;
  if (!tbuf) { CUDA_TRY(cudaMalloc(&tbuf, 64)); CUDA_TRY(cudaMemset(tbuf, 0, 64)); }
  p.timing = tbuf;
  {
    static int calls = 0;
    if (++calls % 24 == 0) {
      unsigned long long hbuf[8];
      cudaMemcpy(hbuf, tbuf, 64, cudaMemcpyDeviceToHost);
      fprintf(stderr, "[ews timing, sums over CTAs, Mcycles] prod wait bfull %.1f of %.1f | loader wait empty %.1f of %.1f | mma wait full %.1f acce %.1f | epi wait accf %.1f of %.1f\n",
              hbuf[0] * 1e-6, hbuf[1] * 1e-6, hbuf[2] * 1e-6, hbuf[3] * 1e-6, hbuf[4] * 1e-6, hbuf[5] * 1e-6, hbuf[6] * 1e-6, hbuf[7] * 1e-6);
      cudaMemset(tbuf, 0, 64);
    }
  }
#endif
  static unsigned long long attr_devices = 0;
  if (dfm_once_per_device(attr_devices, ctx->device)) {
    CUDA_TRY(cudaFuncSetAttribute(ews::k_edge_ws<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ews::SMEM_ALLOC));
    CUDA_TRY(cudaFuncSetAttribute(ews::k_edge_ws<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ews::SMEM_ALLOC));
  }
  int grid = p.ntiles < ctx->num_sms ? p.ntiles : ctx->num_sms;
  if (grid <= 0) return 0;
  p.chunk = (p.ntiles + grid - 1) / grid;
  grid = (p.ntiles + p.chunk - 1) / p.chunk;
  // Bm as a rank-2 tensor {256 columns, B*N rows}; box = 64 columns x 1 row, four rows per tile::gather4 instruction
  CUtensorMap tmB;
  int rc = dfm_make_tmap_f16(&tmB, p.Bm, (uint64_t)p.total_nodes, H, 1);
  if (rc) return rc;
  if (p.last) CUDA_TRY(dfm_launch_pdl(ews::k_edge_ws<true>, dim3(grid), dim3(ews::NT), ews::SMEM_ALLOC, s, p, tmB));
  else CUDA_TRY(dfm_launch_pdl(ews::k_edge_ws<false>, dim3(grid), dim3(ews::NT), ews::SMEM_ALLOC, s, p, tmB));
  LAUNCH_CHECK(ctx);
  return 0;
}

// Last layer without the energy head as one launch (see k_last_fused / last_ring.cuh).  Returns DFM_OK with *used = 0 when the
// fused form does not apply (not requested by flag or DFM_LAST_FUSED=1, small batches, a device that cannot hold the whole grid at once): the caller
// then runs launch_edge_ws + launch_node_coord.
int launch_last_fused(dfm_ctx* ctx, const EdgeArgs& a, const int4* emeta, const __half* Ahi, unsigned int* ring_flags,
                      int ring_flag_words, bool requested, int* used, cudaStream_t s) {
  *used = 0;
  static int enabled = -1, split_pct = 0;
  if (enabled < 0) {
    const char* e = getenv("DFM_LAST_FUSED"); enabled = e ? atoi(e) : 0;
    const char* f = getenv("DFM_LAST_FUSED_SPLIT"); split_pct = f ? atoi(f) : 68;     // % of the CTAs in the edge role
  }
  if (!(enabled || requested) || !a.last || !a.lig_only || a.N <= a.R || a.R <= 1) return 0;
  const LayerW& w = ctx->layer[a.layer];
  const int L = a.N - a.R;
  ews::Params p{};
  p.total_nodes = a.B * a.N;
  p.N = a.N; p.R = a.R; p.K = a.K; p.last = 1; p.lig_only = 1;
  {
    const int L2a = ((a.N - 1) >> 1) - (a.R >> 1) + 1;
    const int L2b = (a.N >> 1) - ((a.R + 1) >> 1) + 1;
    p.tpt = (a.N & 1) ? (L2a > L2b ? L2a : L2b) : L2a;
  }
  p.ntiles = a.B * p.tpt;
  const int grid = ctx->num_sms;
  if (p.ntiles < 4 * grid) return 0;                       // too little work to split the SMs into two roles
  int P = (grid * split_pct + 50) / 100;
  if (P < 1) P = 1;
  if (P > grid - 1) P = grid - 1;
  p.chunk = (p.ntiles + P - 1) / P;
  p.Wimg = w.img_W2h; p.emeta = emeta; p.Ahi = Ahi; p.Bm = reinterpret_cast<const __half*>(a.Bm);
  p.Tdrp = w.Tdrp16h; p.Totp = w.Totp16h;
  p.w1r = w.w1r; p.b2 = w.b2; p.wa = w.wa; p.ba = w.ba;
  p.agg16 = nullptr; p.mstar = a.mstar;

  LastRing ring{};
  ring.P = P; ring.chunk = p.chunk; ring.ntiles = p.ntiles; ring.tpt = p.tpt; ring.N = a.N; ring.R = a.R;
  ring.total_nodes = p.total_nodes;
  // ring slots: four per producer, bounded by the spill buffer it lives in ([B, L, 64, 256] fp16) and by the flag words
  const size_t spill_tiles = ((size_t)a.B * L * SLOTS) / ews::TILE_M;
  size_t NR = (size_t)4 * P;
  if (NR > spill_tiles) NR = spill_tiles;
  if (NR > (size_t)ring_flag_words / 2) NR = (size_t)ring_flag_words / 2;
  if (NR < 8) return 0;
  ring.NR = (int)NR;
  ring.ready = ring_flags; ring.done = ring_flags + NR;
  ring.ring = a.mstar;

  ntc::Params pc{};
  pc.M = a.B * L * SLOTS; pc.ntiles = 0; pc.N = a.N; pc.R = a.R; pc.K = a.K;
  pc.X = a.mstar; pc.W0 = w.img_Wc1s; pc.bias0 = w.bc1; pc.wc2 = w.wc2; pc.nbr = a.nbr; pc.pos = a.pos; pc.fbuf = a.fbuf;

  static unsigned long long attr_devices = 0, ok_devices = 0;
  constexpr int SMEM = (int)(ews::SMEM_ALLOC > ntc::SMEM_ALLOC ? ews::SMEM_ALLOC : ntc::SMEM_ALLOC);
  if (dfm_once_per_device(attr_devices, ctx->device)) {
    CUDA_TRY(cudaFuncSetAttribute(ews::k_last_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ews::k_last_fused, ews::NT, SMEM));
    if (per_sm >= 1 && ctx->device < 64) ok_devices |= 1ull << ctx->device;      // the whole grid is co-resident
  }
  if (ctx->device >= 64 || !(ok_devices & (1ull << ctx->device))) return 0;

  CUtensorMap tmB, tmX;
  int rc = dfm_make_tmap_f16(&tmB, p.Bm, (uint64_t)p.total_nodes, H, 1);
  if (rc) return rc;
  if ((rc = dfm_make_tmap_f16(&tmX, ring.ring, (uint64_t)NR * 4 * ews::TILE_M, 64, ntc::TILE_M))) return rc;
  CUDA_TRY(cudaMemsetAsync(ring_flags, 0, sizeof(unsigned int) * 2 * NR, s));
  CUDA_TRY(dfm_launch_pdl(ews::k_last_fused, dim3(grid), dim3(ews::NT), (size_t)SMEM, s, p, tmB, pc, tmX, ring));
  LAUNCH_CHECK(ctx);
  *used = 1;
  return 0;
}
