// Shared declarations for the dfmdock_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <utility>
#include <vector>

#include "../../include/dfmdock_b200.h"

#define H DFM_NODE_DIM          // 256 node features
#define ED DFM_EDGE_DIM         // 128 pair-embedding features
#define SLOTS DFM_EDGE_SLOTS    // 64 edge slots per node (60 used)
#define NSPATIAL 100            // 40 dist + 24 omega + 24 theta + 12 phi one-hot rows
#define NRELPOS 66

// fp16 operand scaling (exact powers of two, undone in the weight images): agg and the spilled gated messages travel as
// value * 2^-6 (observed |agg| up to 4.4e4)
#define AGG_SCALE 0.015625f
#define AGG_UNSCALE 64.0f

void dfm_set_error(const char* fmt, ...);

#define CUDA_TRY(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      dfm_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));  \
      return DFM_ECUDA;                                                                     \
    }                                                                                       \
  } while (0)

// cudaFuncSetAttribute is per device: remember per kernel (one static mask per call site) which devices are done, so
// that several contexts on different GPUs of one process all get their dynamic shared memory limit raised
__host__ inline bool dfm_once_per_device(unsigned long long& mask, int device) {
  const unsigned long long bit = 1ull << (device & 63);
  if (mask & bit) return false;
  mask |= bit;
  return true;
}

#define LAUNCH_CHECK(ctx)                                                                   \
  do {                                                                                      \
    (ctx)->launches++;                                                                      \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess) {                                                                \
      dfm_set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return DFM_ECUDA;                                                                     \
    }                                                                                       \
  } while (0)

// Programmatic dependent launch (PDL): a kernel launched through dfm_launch_pdl may be scheduled while its predecessor in the
// stream is still draining, runs its prologue (barrier init, TMEM allocation, the bulk copy of its constant weight image)
// and blocks in pdl_wait() until the predecessor has completed and its writes are visible; pdl_trigger() at the top of a
// kernel allows ITS successor to be scheduled as soon as SM resources free up.  Every kernel on the chain executes
// pdl_wait() before it touches anything a predecessor wrote (or still reads), so completion order stays the stream order.
// Without the launch attribute both instructions are no-ops.  The attribute is set for batches of at most DFM_PDL_MAX_ROWS
// residue rows (api.cu: dfm_pdl_enabled); DFM_PDL=0 / 1 in the environment forces it off / on.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
constexpr long long DFM_PDL_MAX_ROWS = 28000;   // crossover between 64 x 300 (PDL faster) and 128 x 300 rows (PDL slower)
bool dfm_pdl_enabled();
void dfm_pdl_set_rows(long long rows);
template <typename... KArgs, typename... Args>
inline cudaError_t dfm_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = dfm_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

struct WTensor {
  float* d = nullptr;
  std::vector<int64_t> shape;
  int64_t numel = 0;
};

// Per-layer device pointers (fp32 originals are views into the WTensor copies).
struct LayerW {
  const float* W1;      // edge_mlp.0.weight [256, 641] = [W1s | W1d | w1r | W1e]
  const float* b1;      // [256]
  const float* W2;      // edge_mlp.2.weight [256,256]
  const float* b2;
  const float* W3;      // node_mlp.0.weight [256,512] = [W3h | W3a]
  const float* b3;
  const float* gn_w;    // GraphNorm weight / bias / mean_scale
  const float* gn_b;
  const float* gn_ms;
  const float* W4;      // node_mlp.3.weight
  const float* b4;
  const float* wa;      // att_mlp.0.weight [256]
  const float* ba;      // [1]
  const float* Wc1;     // coord_mlp.0 (last layer only)
  const float* bc1;
  const float* wc2;     // coord_mlp.2.weight [256]
  // derived (owned)
  float* T32;           // [100+P, 256] pair table, fp32
  float* w1r;           // [256] column 512 of W1, contiguous
  float* b1eff;         // [256] b1 + sym * T[166] (set per complex)
  __half* img_W1s;      // fp16 SW128 images [4 kblk][256 rows][64]
  __half* img_W1d;
  __half* img_W4;
  // operands of the warp-specialised edge kernel (edge_ws.cu): everything pre-halved (SiLU(x) = h + h tanh(h), h = x/2)
  __half* Tdrp16h;      // [(z*40+d)*66+rp][256] merged dist + relpos rows (z=1: + the three zero-angle rows) / 2, fp16 of the fp32 sum
  __half* Totp16h;      // [(o*24+t)*12+p][256] merged omega + theta + phi rows / 2
  __half* img_W2h;      // W2 / 2
  __half* img_Wc1s;     // Wc1 x 2^6 (gated messages are spilled x 2^-6), K in the spill's fragment order (node_tc.cu)
  __half* img_W3z0;     // [W3h | W3a x 2^6] rows 0-127, K = 512 (node_tc.cu MODE_Z)
  __half* img_W3z1;     // rows 128-255
};

struct dfm_ctx {
  int device = 0;
  std::map<std::string, WTensor> w;
  bool finalized = false;
  int P = 0;            // positional_embed width (66 or 67)
  int x_dim = 0;
  float cut_off = 20.f;
  LayerW layer[DFM_DEPTH];
  // heads
  const float* We = nullptr;      // to_energy.0.weight [256,512]
  const float* e_ln_w = nullptr;
  const float* e_ln_b = nullptr;
  const float* e_w = nullptr;     // to_energy.3.weight [256]
  __half* img_WeR = nullptr;
  __half* img_WeL = nullptr;
  const float* t_W = nullptr;     // t_embed.0.W [64]
  const float* t_lin = nullptr;   // t_embed.1.weight transposed: [in 128, out 128]
  const float* sc_W1[2] = {nullptr, nullptr};  // tr_scale / rot_scale .0.weight transposed: [in 129, out 128]
  const float* sc_lnw[2] = {nullptr, nullptr};
  const float* sc_lnb[2] = {nullptr, nullptr};
  const float* sc_w2[2] = {nullptr, nullptr};  // .4.weight [128]
  // to_ires (optional: dead at inference, only for the reference-shaped output dict)
  float* ires_W1t = nullptr;      // [256, 512] = to_ires.0.weight^T
  float* ires_W3t = nullptr;      // [512, 512] = to_ires.2.weight^T
  const float* ires_b1 = nullptr; const float* ires_b3 = nullptr; const float* ires_w5 = nullptr; const float* ires_b5 = nullptr;
  // VE-SDE schedules of the checkpoint (hyper_parameters.diffuser): so3 logarithmic, r3 geometric
  double so3_min_sigma = 0.1, so3_max_sigma = 1.5, r3_min_sigma = 0.1, r3_max_sigma = 30.0;
  // complex
  bool has_complex = false;
  int R = 0, L = 0, N = 0, K = 0, knn = 0, ns = 0;
  float sym = 0.f;
  float* h0 = nullptr;        // [N,256]
  float* rec_pos = nullptr;   // [R,3,3]
  size_t h0_cap = 0, rec_cap = 0;
  float* clash_partial = nullptr;   // [CLASH_BATCH, parts, 4] partial clash forces (pose.cu), grow-only, sized by dfm_set_complex
  size_t clash_cap = 0;
  uint64_t launches = 0;
  int num_sms = 148;
  // optional CUDA-event timing of the dominant (edge) kernel, for bench.py's roofline line
  bool profile = false;
  std::vector<cudaEvent_t> prof_events;   // pairs (start, stop)
  size_t prof_used = 0;
  std::vector<void*> owned;   // cudaMalloc'd derived buffers
};

// Workspace carve-up for B trajectories (all offsets 256-byte aligned).
constexpr int RING_FLAG_WORDS = 2048;
constexpr int CLASH_BATCH = 1024;    // trajectories per clash-force launch pair (pose.cu)
size_t clash_scratch_floats(int R, int L);
struct Workspace {
  float* centre;     // [B,4]
  float* pos;        // [B,N,3,3] centred N/CA/C
  float* cb;         // [B,N,4]   virtual CB (padded)
  int32_t* nbr;      // [B,N,64]
  uint32_t* feat;    // [B,N,64]  packed bins: d | o<<6 | t<<11 | p<<16 | rp<<20
  float* radial;     // [B,N,64]
  float* h;          // [B,N,256]
  float* A;          // [B,N,256]
  float* Bm;         // [B,N,256] fp32 (simt) or fp16 in the first half (tc)
  float* agg;        // [B,N,256]
  float* z;          // [B,N,256]
  float* y;          // [B,N,256]
  float* gstat;      // [B,2,256] GraphNorm mean / rstd
  __half* mstar;     // [B,L,64,256] fp16 * 2^-4 (tc path, last layer)
  float* fbuf;       // [B,L,4]
  float* esum;       // [B,R,2]
  float* tsc;        // [B,8] tr_score(3) rot_score(3) scratch when the caller passes NULL
  int4* emeta;       // [B,N,64] {global row of j, Tdrp row, Totp row or -1, packed half2 radial/32} (edge_ws.cu)
  __half* h16;       // [B,N,256] fp16 copy of h (operand of the node-side GEMMs, node_tc.cu)
  __half* agg16;     // [B,N,256] fp16(agg * 2^-6) written by edge_ws.cu
  float* gscale;     // [B,256] GraphNorm weight * rstd
  float* gshift;     // [B,256] GraphNorm bias - mean*mean_scale*gscale
  unsigned int* ring_flags;   // [RING_FLAG_WORDS] ready / done counters of the fused last-layer kernel (last_ring.cuh)
  size_t bytes;
};

Workspace carve_workspace(const dfm_ctx* ctx, int B, void* base);

// ---- kernels (launchers) -----------------------------------------------------------------------
struct LinearArgs {
  const float* A;       // [M, 256] fp32 activations (row stride 256)
  float a_scale;        // applied before the fp16 conversion in the tc path (1 = none)
  const float* W32;     // fp32 weight, row-major [256, ldw], first column = w_col0
  int ldw;
  int w_col0;
  const __half* Wimg;   // fp16 image (already multiplied by 1/a_scale)
  const float* bias;    // [256] or null
  const float* add;     // [M,256] or null; out = add + A*W^T + bias  (may alias out)
  float* out;           // [M,256] fp32 or null
  __half* out16;        // [M,256] fp16 or null: fp16(out_scale * value)
  float out_scale;      // applied to out16 only (0 = 1)
  int M;
};
int launch_linear_simt(dfm_ctx* ctx, const LinearArgs& a, cudaStream_t s);
int launch_linear_tc(dfm_ctx* ctx, const LinearArgs& a, cudaStream_t s);

struct EdgeArgs {
  int B, N, R, K;
  int layer;
  bool last;            // coordinate update for ligand rows
  bool lig_only;        // last layer, no energy head wanted: only the tiles that hold a ligand residue are needed
  const int32_t* nbr;
  const uint32_t* feat;
  const float* radial;
  const float* A;       // [B,N,256] W1s h_i + b1eff
  const void* Bm;       // [B,N,256] W1d h_j  (fp32 simt / fp16 tc)
  const float* pos;     // [B,N,3,3] (for the coordinate update)
  float* agg;           // [B,N,256]
  __half* mstar;        // tc path, last layer
  float* fbuf;          // [B,L,4] out (last layer)
};
int launch_edge_simt(dfm_ctx* ctx, const EdgeArgs& a, cudaStream_t s);
int launch_edge_ws(dfm_ctx* ctx, const EdgeArgs& a, const int4* emeta, const __half* Ahi, __half* agg16, cudaStream_t s);
int launch_last_fused(dfm_ctx* ctx, const EdgeArgs& a, const int4* emeta, const __half* Ahi, unsigned int* ring_flags,
                      int ring_flag_words, bool requested, int* used, cudaStream_t s);

int launch_prepare(dfm_ctx* ctx, int B, const float* lig_pos, Workspace& ws, cudaStream_t s);
int launch_graph(dfm_ctx* ctx, int B, bool generic, const int32_t* edges, const float* exp_noise, uint64_t seed,
                 uint64_t stream_base, uint32_t fwd_index, Workspace& ws, cudaStream_t s);
int launch_broadcast_h0(dfm_ctx* ctx, int B, Workspace& ws, cudaStream_t s);
int launch_graphnorm_silu(dfm_ctx* ctx, int B, int layer, Workspace& ws, cudaStream_t s);
int launch_graphnorm_stats(dfm_ctx* ctx, int B, int layer, const float* z, float* gscale, float* gshift, cudaStream_t s);
int launch_node_ab(dfm_ctx* ctx, int layer, int M, const __half* h16, __half* Ah, __half* Bm, cudaStream_t s);
int launch_image_pack_z(dfm_ctx* ctx, const float* W3, float scale_hi, __half* img0, __half* img1, cudaStream_t s);
int launch_node_coord(dfm_ctx* ctx, const EdgeArgs& a, cudaStream_t s);
int launch_node_z(dfm_ctx* ctx, int layer, int M, const __half* h16, const __half* agg16, float* z, cudaStream_t s);
int launch_node_h(dfm_ctx* ctx, int layer, int M, const float* z, const float* gscale, const float* gshift, float* h,
                  __half* h16, cudaStream_t s);
int launch_force_head(dfm_ctx* ctx, int B, const float* t, Workspace& ws, float* tr_score, float* rot_score,
                      float* f_out, cudaStream_t s);
int launch_energy(dfm_ctx* ctx, int B, bool fp32_path, Workspace& ws, float* energy, int32_t* clashes,
                  cudaStream_t s);
int launch_image_pack(dfm_ctx* ctx, const float* W, int ldw, int col0, float scale, __half* img, cudaStream_t s);
int launch_image_pack_perm(dfm_ctx* ctx, const float* W, int ldw, float scale, __half* img, cudaStream_t s);
int launch_pair_table(dfm_ctx* ctx, int l, cudaStream_t s);
int launch_single_embed(dfm_ctx* ctx, const float* rec_x, const float* lig_x, cudaStream_t s);
int launch_transpose(dfm_ctx* ctx, const float* W, int rows, int cols, float* Wt, cudaStream_t s);
int launch_ires(dfm_ctx* ctx, int rows, const float* h, float* out, cudaStream_t s);

// ---- small device helpers ----------------------------------------------------------------------
__device__ __forceinline__ float silu_f(float x) { return x / (1.f + __expf(-x)); }
__device__ __forceinline__ float silu_acc(float x) { return x / (1.f + expf(-x)); }
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Philox4x32-10 (Salmon et al. 2011), counter-based: independent of launch geometry and sharding.
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
// uniform on the OPEN interval (0, 1): 23 random bits + 1/2, every value exactly representable, so neither 0 nor 1 can be
// returned (with 24 bits the top value rounds to 1.0f and -logf gives -0.0f, which the bit-pattern selection of
// k_graph_sel orders last while the float comparison of k_graph orders it first -- found by the kernel-equivalence test)
__device__ __forceinline__ float u01_open(uint32_t x) { return ((float)(x >> 9) + 0.5f) * (1.0f / 8388608.0f); }
// kinds of draws (third counter word, high byte)
#define RNG_EDGE 0u
#define RNG_STEP 1u
#define RNG_INIT 2u
__device__ __forceinline__ uint4 dfm_rng(uint64_t seed, uint64_t stream, uint32_t kind, uint32_t index,
                                         uint32_t a, uint32_t b) {
  uint2 key = make_uint2((uint32_t)seed ^ (uint32_t)(stream >> 32), (uint32_t)(seed >> 32) ^ 0x5bd1e995u);
  uint4 ctr = make_uint4(a, b, (kind << 24) | (index & 0xffffffu), (uint32_t)stream);
  return philox4x32(ctr, key);
}
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  float r = sqrtf(-2.f * logf(u01_open(a)));
  float th = 6.283185307179586f * u01_open(b);
  return make_float2(r * cosf(th), r * sinf(th));
}
