// Pose preparation, stochastic residue graph (kNN-20 + 40 samples ~ 1/d^3) and 6D pair-feature bins.
//
// Reference behaviour restated here (never its code):
//   centring                    src/models/score_net_mlsb.py:352-355
//   get_knn_and_sample          src/models/score_net_mlsb.py:85-133   (torch.multinomial w/o replacement
//                               == top-k of p / Exp(1): the exponential race)
//   get_coords6d / get_bins     src/utils/coords6d.py:62-103, src/models/score_net_mlsb.py:30-70
//   relpos                      src/inference_base.py:246-292
// Unlike the reference, the N x N x 100 one-hot tensor is never built: only the 60 selected pairs of each
// residue get their five bin indices, packed into one uint32.
#include <float.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

__constant__ float c_edge_d[39];
__constant__ float c_edge_a[23];
__constant__ float c_edge_p[11];

static void fill_linspace(float* out, float lo, float hi, int steps) {
  // torch.linspace (CPU, fp32): fused multiply-add from both ends.
  float step = (hi - lo) / (float)(steps - 1);
  for (int i = 0; i < steps; ++i)
    out[i] = (i < steps / 2) ? fmaf(step, (float)i, lo) : fmaf(-step, (float)(steps - 1 - i), hi);
}

int dfm_upload_bin_edges() {
  float d[39], a[23], p[11];
  fill_linspace(d, 3.25f, 50.75f, 39);
  fill_linspace(a, -180.f, 180.f, 23);
  fill_linspace(p, 0.f, 180.f, 11);
  if (cudaMemcpyToSymbol(c_edge_d, d, sizeof(d)) != cudaSuccess) return -1;
  if (cudaMemcpyToSymbol(c_edge_a, a, sizeof(a)) != cudaSuccess) return -1;
  if (cudaMemcpyToSymbol(c_edge_p, p, sizeof(p)) != cudaSuccess) return -1;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// k_prepare: per trajectory, centre everything on the ligand CA centroid and build the virtual CB.
__global__ void __launch_bounds__(256) k_prepare(int N, int R, const float* __restrict__ rec_pos,
                                                 const float* __restrict__ lig_pos, float* __restrict__ centre,
                                                 float* __restrict__ pos, float* __restrict__ cb) {
  const int b = blockIdx.x, L = N - R, tid = threadIdx.x;
  const float* lp = lig_pos + (size_t)b * L * 9;
  __shared__ float red[3][256];
  float sx = 0.f, sy = 0.f, sz = 0.f;
  for (int l = tid; l < L; l += 256) {
    sx += lp[l * 9 + 3];
    sy += lp[l * 9 + 4];
    sz += lp[l * 9 + 5];
  }
  red[0][tid] = sx; red[1][tid] = sy; red[2][tid] = sz;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) {
      red[0][tid] += red[0][tid + o];
      red[1][tid] += red[1][tid + o];
      red[2][tid] += red[2][tid + o];
    }
    __syncthreads();
  }
  const float cx = red[0][0] / (float)L, cy = red[1][0] / (float)L, cz = red[2][0] / (float)L;
  if (tid == 0) {
    centre[b * 4 + 0] = cx; centre[b * 4 + 1] = cy; centre[b * 4 + 2] = cz; centre[b * 4 + 3] = 0.f;
  }
  for (int n = tid; n < N; n += 256) {
    const float* src = (n < R) ? rec_pos + (size_t)n * 9 : lp + (size_t)(n - R) * 9;
    float v[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) v[q] = src[q] - ((q % 3 == 0) ? cx : (q % 3 == 1) ? cy : cz);
    float* dst = pos + ((size_t)b * N + n) * 9;
#pragma unroll
    for (int q = 0; q < 9; ++q) dst[q] = v[q];
    // virtual CB (coords6d.py:72-76)
    float bx = v[3] - v[0], by = v[4] - v[1], bz = v[5] - v[2];
    float cx2 = v[6] - v[3], cy2 = v[7] - v[4], cz2 = v[8] - v[5];
    float ax = by * cz2 - bz * cy2, ay = bz * cx2 - bx * cz2, az = bx * cy2 - by * cx2;
    float* cbo = cb + ((size_t)b * N + n) * 4;
    cbo[0] = -0.58273431f * ax + 0.56802827f * bx - 0.54067466f * cx2 + v[3];
    cbo[1] = -0.58273431f * ay + 0.56802827f * by - 0.54067466f * cy2 + v[4];
    cbo[2] = -0.58273431f * az + 0.56802827f * bz - 0.54067466f * cz2 + v[5];
    cbo[3] = 0.f;
  }
}

int launch_prepare(dfm_ctx* ctx, int B, const float* lig_pos, Workspace& ws, cudaStream_t s) {
  k_prepare<<<B, 256, 0, s>>>(ctx->N, ctx->R, ctx->rec_pos, lig_pos, ws.centre, ws.pos, ws.cb);
  LAUNCH_CHECK(ctx);
  return 0;
}

// ------------------------------------------------------------------------------------------------
struct V3 { float x, y, z; };
__device__ __forceinline__ V3 sub3(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 cross3(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float norm3(V3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
__device__ __forceinline__ V3 unit3(V3 a) { float n = norm3(a); return {a.x / n, a.y / n, a.z / n}; }

__device__ __forceinline__ float dihedral_deg(V3 p0, V3 p1, V3 p2, V3 p3) {
  V3 b1 = sub3(p0, p1), b2 = sub3(p1, p2), b3 = sub3(p2, p3);
  V3 n1 = unit3(cross3(b1, b2));
  V3 n2 = unit3(cross3(b2, b3));
  V3 m1 = cross3(n1, unit3(b2));
  float ang = atan2f(dot3(m1, n2), dot3(n1, n2));
  return ang * 180.f / 3.14159265358979323846f;
}
__device__ __forceinline__ float planar_deg(V3 p0, V3 p1, V3 p2) {
  V3 v1 = sub3(p0, p1), v2 = sub3(p2, p1);
  float ang = acosf(dot3(v1, v2) / (norm3(v1) * norm3(v2)));
  return ang * 180.f / 3.14159265358979323846f;
}
template <int NE>
__device__ __forceinline__ int count_above(const float* edges, float x) {
  int c = 0;
#pragma unroll
  for (int k = 0; k < NE; ++k) c += (x > edges[k]) ? 1 : 0;   // NaN compares false -> bin 0
  return c;
}

__device__ __forceinline__ uint32_t pair_bins(const float* __restrict__ pos, const float* __restrict__ cb, int gi,
                                              int gj, int i, int j, int R, float* radial_out) {
  const float* pi = pos + (size_t)gi * 9;
  const float* pj = pos + (size_t)gj * 9;
  V3 Ni = {pi[0], pi[1], pi[2]}, CAi = {pi[3], pi[4], pi[5]};
  V3 CAj = {pj[3], pj[4], pj[5]};
  V3 CBi = {cb[(size_t)gi * 4], cb[(size_t)gi * 4 + 1], cb[(size_t)gi * 4 + 2]};
  V3 CBj = {cb[(size_t)gj * 4], cb[(size_t)gj * 4 + 1], cb[(size_t)gj * 4 + 2]};
  V3 d = sub3(CAi, CAj);
  float r2 = d.x * d.x + d.y * d.y + d.z * d.z;
  *radial_out = r2;
  float dist = sqrtf(r2);
  uint32_t db = count_above<39>(c_edge_d, dist);
  uint32_t ob = 0, tb = 0, pb = 0;
  if (dist < 22.0f && i != j) {
    ob = count_above<23>(c_edge_a, dihedral_deg(CAi, CBi, CBj, CAj));
    tb = count_above<23>(c_edge_a, dihedral_deg(Ni, CAi, CBi, CBj));
    pb = count_above<11>(c_edge_p, planar_deg(CAi, CBi, CBj));
  }
  uint32_t rp;
  if ((i < R) == (j < R)) {
    int o = i - j + 32;
    rp = (uint32_t)min(max(o, 0), 64);
  } else {
    rp = 65u;
  }
  return db | (ob << 6) | (tb << 11) | (pb << 16) | (rp << 20);
}

// warp-wide argmin over (value, index); ties -> smaller index
__device__ __forceinline__ void warp_argmin(float& v, int& j) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, v, o);
    int oj = __shfl_xor_sync(0xffffffffu, j, o);
    if (ov < v || (ov == v && oj < j)) { v = ov; j = oj; }
  }
}

// One warp per (trajectory, residue) row.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k_graph(int B, int N, int R, int K, int knn, int ns, const float* __restrict__ pos, const float* __restrict__ cb,
        const int32_t* __restrict__ edges_in, const float* __restrict__ exp_noise, uint64_t seed,
        uint64_t stream_base, uint32_t fwd, int32_t* __restrict__ nbr, uint32_t* __restrict__ feat,
        float* __restrict__ radial, int4* __restrict__ emeta) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int npad = (N + 31) & ~31;
  float* key = smem + (size_t)warp * (npad + 32);
  int* knn_list = reinterpret_cast<int*>(key + npad);   // the kNN indices, for the compacted noise index
  const long row = (long)blockIdx.x * WARPS + warp;
  if (row >= (long)B * N) return;
  const int b = (int)(row / N), i = (int)(row % N);
  const size_t gbase = (size_t)b * N;
  int sel0 = i, sel1 = i;   // neighbour for slot lane / lane+32

  if (edges_in != nullptr) {
    const int32_t* e = edges_in + (gbase + i) * K;
    if (lane < K) sel0 = e[lane];
    if (lane + 32 < K) sel1 = e[lane + 32];
  } else {
    const float* pi = pos + (gbase + i) * 9;
    const float xi = pi[3], yi = pi[4], zi = pi[5];
    for (int j = lane; j < N; j += 32) {
      const float* pj = pos + (gbase + j) * 9;
      float dx = xi - pj[3], dy = yi - pj[4], dz = zi - pj[5];
      key[j] = sqrtf(dx * dx + dy * dy + dz * dz);
    }
    __syncwarp();
    // ---- kNN: knn smallest distances (the residue itself first, d = 0)
    for (int k = 0; k < knn; ++k) {
      float bv = FLT_MAX;
      int bj = 0x7fffffff;
      for (int j = lane; j < N; j += 32) {
        float v = key[j];
        if (v < bv) { bv = v; bj = j; }
      }
      warp_argmin(bv, bj);
      if (lane == (k & 31)) { if (k < 32) sel0 = bj; else sel1 = bj; }
      if (lane == 0) { key[bj] = FLT_MAX; knn_list[k] = bj; }
      __syncwarp();
    }
    // ---- exponential race over the remaining residues: key = Exp(1) * d^3, keep the ns smallest
    if (ns > 0) {
      const uint64_t strm = stream_base + (uint64_t)b;
      for (int j = lane; j < N; j += 32) {
        float d = key[j];
        if (d != FLT_MAX) {
          float e;
          if (exp_noise != nullptr) {
            int below = 0;
            for (int k = 0; k < knn; ++k) below += (knn_list[k] < j) ? 1 : 0;
            e = exp_noise[(gbase + i) * (size_t)(N - knn) + (j - below)];
          } else {
            uint4 r4 = dfm_rng(seed, strm, RNG_EDGE, fwd, (uint32_t)(j >> 2), (uint32_t)i);
            uint32_t u = (j & 3) == 0 ? r4.x : (j & 3) == 1 ? r4.y : (j & 3) == 2 ? r4.z : r4.w;
            e = -logf(u01_open(u));
          }
          d = fmaxf(d, 1e-10f);
          key[j] = e * (d * d * d);
        }
      }
      __syncwarp();
      for (int k = knn; k < knn + ns; ++k) {
        float bv = FLT_MAX;
        int bj = 0x7fffffff;
        for (int j = lane; j < N; j += 32) {
          float v = key[j];
          if (v < bv) { bv = v; bj = j; }
        }
        warp_argmin(bv, bj);
        if (lane == (k & 31)) { if (k < 32) sel0 = bj; else sel1 = bj; }
        if (lane == 0 && bj < N) key[bj] = FLT_MAX;
        __syncwarp();
      }
    }
  }
  // ---- pair features for the K selected edges; pad slots point at the residue itself
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int k = lane + 32 * half;
    int j = half == 0 ? sel0 : sel1;
    uint32_t ft = 0;
    float r2 = 0.f;
    if (k < K) {
      j = min(max(j, 0), N - 1);
      ft = pair_bins(pos, cb, (int)(gbase + i), (int)(gbase + j), i, j, R, &r2);
    } else {
      j = i;
    }
    const size_t o = (gbase + i) * SLOTS + k;
    nbr[o] = j;
    feat[o] = ft;
    radial[o] = r2;
    {  // gather indices for the warp-specialised edge kernel (same decode as tc.cu k_tc<EDGE>)
      const uint32_t otp = (ft >> 6) & 0x3FFFu;
      const int drp = (int)(((otp == 0 ? 40u : 0u) + (ft & 63u)) * 66u + ((k < K) ? ((ft >> 20) & 127u) : 32u));
      const int oidx = (int)((((ft >> 6) & 31u) * 24u + ((ft >> 11) & 31u)) * 12u + ((ft >> 16) & 15u));
      // radial travels as packed half2 of radial / 32 (edge_ws.cu RAD_SCALE), clamped to the fp16 range
      const __half2 rh = __float2half2_rn(fminf(r2 * 0.03125f, 65000.f));
      emeta[o] = make_int4((int)(gbase + j), drp, otp == 0 ? -1 : oidx, *reinterpret_cast<const int*>(&rh));
    }
  }
}


// ------------------------------------------------------------------------------------------------
// k_graph_sel: same result as k_graph (as sets), selection by bit-wise search instead of k argmin passes.
// One warp per (trajectory, residue) row; the lane holds the keys of residues j = lane + 32 s in registers as
// order-preserving unsigned bit patterns of non-negative floats.  The k-th smallest key is found most-significant
// bit first (one compare per element and one REDUX per bit, starting below the common prefix of all keys), ties at
// the threshold go to the smaller residue index (like k_graph); members are compacted to slots with warp ballots.
template <int WARPS, int S, bool STAGE>
__global__ void __launch_bounds__(WARPS * 32)
k_graph_sel(int B, int N, int R, int K, int knn, int ns, int rpc, int cpt, const float* __restrict__ pos,
            const float* __restrict__ cb, const float* __restrict__ exp_noise, uint64_t seed, uint64_t stream_base,
            uint32_t fwd, int32_t* __restrict__ nbr, uint32_t* __restrict__ feat, float* __restrict__ radial,
            int4* __restrict__ emeta) {
  __shared__ float s_e[WARPS][S * 32];        // Exp(1) draws in residue order (Philox blocks cover 4 residues each)
  __shared__ int s_slot[WARPS][SLOTS];
  __shared__ __align__(8) unsigned long long s_bar;
  extern __shared__ __align__(16) unsigned char s_tile[];   // the trajectory's backbone [N,3,3] and virtual CB [N,4], staged by TMA
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // CTA = rows [c * rpc, (c + 1) * rpc) of ONE trajectory b (cpt CTAs per trajectory).
  // STAGE: the trajectory's coordinates are brought into shared memory once per CTA by two bulk copies (cp.async.bulk,
  // completion by transaction bytes on an mbarrier) and every distance / pair feature of the CTA's rows is computed from
  // there; rpc is chosen by the launcher so that the copy is amortised over several rows per warp while the grid still
  // fills the GPU several times over.  The [N,9] float block starts on a 4-byte boundary only: the copy starts at the
  // 16-byte boundary below it.  Measured on B200 at 2x150 residues x 256 trajectories: 421 us staged vs 386 us reading the
  // same 15.6 KB per trajectory through L1 (same 325 M warp instructions; the staged form loses 4 % to its coarser grid
  // and 4 % of issue rate to shared-memory bank conflicts of the random-neighbour reads), so the direct form is the
  // default and DFM_GRAPH_STAGE=1 selects this one (tests/test_gpu_parity.py checks that both give the same graph).
  const int b = (int)blockIdx.x / cpt, c = (int)blockIdx.x - b * cpt;
  const size_t gbase = (size_t)b * N;
  const float* P = pos + gbase * 9;             // trajectory-local views (residue j at P + 9 j, CB + 4 j)
  const float* CB = cb + gbase * 4;
  if (STAGE) {
    const uintptr_t psrc = reinterpret_cast<uintptr_t>(pos + gbase * 9);
    const uint32_t lead = (uint32_t)(psrc & 15u);
    const uint32_t pbytes = (lead + (uint32_t)N * 36u + 15u) & ~15u, cbytes = (uint32_t)N * 16u;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
    if (threadIdx.x == 0) {
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_tile);
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(pbytes + cbytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst), "l"(psrc - lead), "r"(pbytes), "r"(bar) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst + pbytes), "l"(cb + gbase * 4), "r"(cbytes), "r"(bar) : "memory");
    }
    __syncthreads();                             // the barrier is initialised before anyone waits on it
    uint32_t ok = 0, tries = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(bar) : "memory");
      if (!ok && ++tries > (1u << 24)) __trap();
    }
    P = reinterpret_cast<const float*>(s_tile + lead);
    CB = reinterpret_cast<const float*>(s_tile + pbytes);
  }
  const uint32_t lt_mask = (1u << lane) - 1u;
  auto do_row = [&](const int i) {
    // ---- distances (float bits; absent -> 0xFFFFFFFF)
    uint32_t kd[S];
    {
      const float* pi = P + (size_t)i * 9;
      const float xi = pi[3], yi = pi[4], zi = pi[5];
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const int j = lane + 32 * s;
        kd[s] = 0xFFFFFFFFu;
        if (j < N) {
          const float* pj = P + (size_t)j * 9;
          const float dx = xi - pj[3], dy = yi - pj[4], dz = zi - pj[5];
          kd[s] = __float_as_uint(sqrtf(dx * dx + dy * dy + dz * dz));
        }
      }
    }
    // k-th smallest of the present keys, then membership flags with ties to the smaller index
    auto select = [&](const uint32_t (&key)[S], int k, uint32_t& member) {
      // common prefix: skip the bits on which min and max agree
      uint32_t mn = 0xFFFFFFFFu, mx = 0u;
#pragma unroll
      for (int s = 0; s < S; ++s) {
        if (key[s] != 0xFFFFFFFFu) { mn = min(mn, key[s]); mx = max(mx, key[s]); }
      }
      mn = __reduce_min_sync(0xffffffffu, mn);
      mx = __reduce_max_sync(0xffffffffu, mx);
      uint32_t v = mn;
      bool exact = false;            // true: {key < v} has exactly k members
      if (mn != mx) {
        const int top = 31 - __clz(mn ^ mx);            // highest differing bit
        v = (top == 31) ? 0u : (mn >> (top + 1)) << (top + 1);
        for (int bit = top; bit >= 0; --bit) {
          const uint32_t test = v | (1u << bit);
          int cnt = 0;
#pragma unroll
          for (int s = 0; s < S; ++s) cnt += (key[s] < test) ? 1 : 0;
          cnt = __reduce_add_sync(0xffffffffu, cnt);
          if (cnt == k) { v = test; exact = true; break; }
          if (cnt < k) v = test;
        }
      }
      member = 0u;                    // bit s: element s of this lane is selected
      if (exact) {
#pragma unroll
        for (int s = 0; s < S; ++s) member |= (key[s] < v) ? (1u << s) : 0u;
      } else {
        int less = 0;
#pragma unroll
        for (int s = 0; s < S; ++s) less += (key[s] < v) ? 1 : 0;
        less = __reduce_add_sync(0xffffffffu, less);
        int need = k - less;          // how many of the keys equal to v to take, smallest index first
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const bool eq = key[s] == v && key[s] != 0xFFFFFFFFu;
          const uint32_t bal = __ballot_sync(0xffffffffu, eq);
          const bool take = eq && (int)__popc(bal & lt_mask) < need;
          member |= ((key[s] < v) || take) ? (1u << s) : 0u;
          need -= min(need, (int)__popc(bal));
        }
      }
    };

    uint32_t mem1;
    select(kd, knn, mem1);
    // slots of the kNN members (index order) and, for injected noise, the rank of every residue among the non-members
    int base = 0;
    int nonmember_before[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const bool m = (mem1 >> s) & 1u;
      const uint32_t bal = __ballot_sync(0xffffffffu, m);
      const int below = base + (int)__popc(bal & lt_mask);      // members with a smaller index
      if (m) s_slot[warp][below] = lane + 32 * s;
      nonmember_before[s] = (lane + 32 * s) - below;            // compacted column of this residue in the noise row
      base += (int)__popc(bal);
    }
    // ---- exponential race over the remaining residues: key = Exp(1) * d^3, keep the ns smallest
    if (ns > 0) {
      if (exp_noise == nullptr) {
        const uint64_t strm = stream_base + (uint64_t)b;
        for (int q = lane; q * 4 < N; q += 32) {               // Philox block q covers residues 4q .. 4q+3
          const uint4 r4 = dfm_rng(seed, strm, RNG_EDGE, fwd, (uint32_t)q, (uint32_t)i);
          *reinterpret_cast<float4*>(&s_e[warp][q * 4]) =
              make_float4(-logf(u01_open(r4.x)), -logf(u01_open(r4.y)), -logf(u01_open(r4.z)), -logf(u01_open(r4.w)));
        }
        __syncwarp();
      }
      uint32_t k2[S];
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const int j = lane + 32 * s;
        k2[s] = 0xFFFFFFFFu;
        if (j < N && !((mem1 >> s) & 1u)) {
          const float e = exp_noise ? exp_noise[(gbase + i) * (size_t)(N - knn) + nonmember_before[s]] : s_e[warp][j];
          const float d = fmaxf(__uint_as_float(kd[s]), 1e-10f);
          k2[s] = min(__float_as_uint(e * (d * d * d)), 0xFFFFFFFEu);
        }
      }
      uint32_t mem2;
      select(k2, ns, mem2);
      int base2 = knn;
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const bool m = (mem2 >> s) & 1u;
        const uint32_t bal = __ballot_sync(0xffffffffu, m);
        if (m) s_slot[warp][base2 + (int)__popc(bal & lt_mask)] = lane + 32 * s;
        base2 += (int)__popc(bal);
      }
    }
    __syncwarp();
    // ---- pair features for the K selected edges; pad slots point at the residue itself
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int k = lane + 32 * half;
      int j = i;
      uint32_t ft = 0;
      float r2 = 0.f;
      if (k < K) {
        j = min(max(s_slot[warp][k], 0), N - 1);
        ft = pair_bins(P, CB, i, j, i, j, R, &r2);
      }
      const size_t o = (gbase + i) * SLOTS + k;
      nbr[o] = j;
      feat[o] = ft;
      radial[o] = r2;
      {
        const uint32_t otp = (ft >> 6) & 0x3FFFu;
        const int drp = (int)(((otp == 0 ? 40u : 0u) + (ft & 63u)) * 66u + ((k < K) ? ((ft >> 20) & 127u) : 32u));
        const int oidx = (int)((((ft >> 6) & 31u) * 24u + ((ft >> 11) & 31u)) * 12u + ((ft >> 16) & 15u));
        const __half2 rh = __float2half2_rn(fminf(r2 * 0.03125f, 65000.f));
        emeta[o] = make_int4((int)(gbase + j), drp, otp == 0 ? -1 : oidx, *reinterpret_cast<const int*>(&rh));
      }
    }
  };
  if (STAGE) {
    const int i_end = min(N, (c + 1) * rpc);
    for (int i = c * rpc + warp; i < i_end; i += WARPS) {
      do_row(i);
      __syncwarp();      // s_e / s_slot of this warp are reused by its next row
    }
  } else {               // one row per warp (rpc == WARPS)
    const int i = c * WARPS + warp;
    if (i < N) do_row(i);
  }
}

// ------------------------------------------------------------------------------------------------
// k_graph_big: complexes too large for the register-resident kernel (N > 1024).  Same result as k_graph (as sets); the keys
// of a row live in shared memory as order-preserving bit patterns and the k smallest are found in two phases:
//   A  bit-wise search from the highest differing bit over all N keys, only until the keys that share the current prefix
//      (the candidates among which the k-th smallest lies) are at most 64 -- for distances the count falls eightfold per
//      bit, for the race keys twofold, instead of the 25-30 full passes of a search carried to the last bit (or the 60
//      arg-min passes of k_graph);
//   B  one pass that emits every key below the prefix (selected for sure) and gathers the candidates, which are then
//      ranked against each other (ties to the smaller residue index, like k_graph) in registers.
// One warp per (trajectory, residue) row.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k_graph_big(int B, int N, int R, int K, int knn, int ns, const float* __restrict__ pos, const float* __restrict__ cb,
            const float* __restrict__ exp_noise, uint64_t seed, uint64_t stream_base, uint32_t fwd,
            int32_t* __restrict__ nbr, uint32_t* __restrict__ feat, float* __restrict__ radial, int4* __restrict__ emeta) {
  extern __shared__ uint32_t smem_u[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int npad = (N + 31) & ~31, S = npad >> 5;
  uint32_t* key = smem_u + (size_t)warp * (npad + 192);
  uint32_t* cand_key = key + npad;                          // [64]
  int* cand_idx = reinterpret_cast<int*>(cand_key + 64);    // [64]
  int* slot = cand_idx + 64;                                // [64] selected residues
  const long row = (long)blockIdx.x * WARPS + warp;
  if (row >= (long)B * N) return;
  const int b = (int)(row / N), i = (int)(row % N);
  const size_t gbase = (size_t)b * N;
  const uint32_t lt_mask = (1u << lane) - 1u;
  constexpr uint32_t ABSENT = 0xFFFFFFFFu;
  uint32_t* kl = key + lane;                                // this lane's keys: kl[32 s] <-> residue lane + 32 s

  {
    const float* pi = pos + (gbase + i) * 9;
    const float xi = pi[3], yi = pi[4], zi = pi[5];
    const float* pl = pos + (gbase + lane) * 9 + 3;
#pragma unroll 4
    for (int s = 0; s < S; ++s) {
      uint32_t kd = ABSENT;
      if (lane + 32 * s < N) {
        const float* pj = pl + (size_t)s * (32 * 9);
        const float dx = xi - __ldg(pj), dy = yi - __ldg(pj + 1), dz = zi - __ldg(pj + 2);
        kd = __float_as_uint(sqrtf(dx * dx + dy * dy + dz * dz));
      }
      kl[32 * s] = kd;
    }
  }
  __syncwarp();

  // the k smallest present keys -> slot[base .. base + k); selected keys are overwritten with ABSENT
  auto select = [&](int k, int base) {
    uint32_t mn = ABSENT, mx = 0u;
    int present = 0;
#pragma unroll 4
    for (int s = 0; s < S; ++s) {
      const uint32_t v = kl[32 * s];
      const bool p = v != ABSENT;
      mn = min(mn, v); mx = max(mx, p ? v : 0u); present += p ? 1 : 0;
    }
    mn = __reduce_min_sync(0xffffffffu, mn);
    mx = __reduce_max_sync(0xffffffffu, mx);
    present = __reduce_add_sync(0xffffffffu, present);
    if (k > present) k = present;
    if (k <= 0) return;
    // invariant: lo = #(key < v) < k <= hi = #(key < vhi), the k-th smallest key lies in [v, vhi)
    uint32_t v = mn, vhi = ABSENT;             // vhi = ABSENT: every present key is below it
    int lo = 0, hi = present;
    if (mn != mx) {
      const int top = 31 - __clz(mn ^ mx);
      v = (top == 31) ? 0u : (mn >> (top + 1)) << (top + 1);
      for (int bit = top; bit >= 0 && hi - lo > 64; --bit) {
        const uint32_t test = v | (1u << bit);
        int cnt = 0;
#pragma unroll 8
        for (int s = 0; s < S; ++s) cnt += (kl[32 * s] < test) ? 1 : 0;
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (cnt < k) { v = test; lo = cnt; } else { vhi = test; hi = cnt; }
      }
    }
    const int need = k - lo;                   // how many of the candidates [v, vhi) to take
    const int ncand = hi - lo;
    if (ncand <= 64) {
      // phase B: emit the sure members, gather the candidates (both in residue order)
      int nm = 0, nc = 0;
#pragma unroll 2
      for (int s = 0; s < S; ++s) {
        const uint32_t kv = kl[32 * s];
        const bool m = kv < v;
        const bool c = !m && kv < vhi;         // ABSENT is never below vhi
        const uint32_t bm = __ballot_sync(0xffffffffu, m), bc = __ballot_sync(0xffffffffu, c);
        if (m) { slot[base + nm + (int)__popc(bm & lt_mask)] = lane + 32 * s; kl[32 * s] = ABSENT; }
        if (c) { const int q = nc + (int)__popc(bc & lt_mask); cand_key[q] = kv; cand_idx[q] = lane + 32 * s; }
        nm += (int)__popc(bm); nc += (int)__popc(bc);
      }
      __syncwarp();
      // rank the candidates among themselves (key, then residue index); candidate q = lane and lane + 32
      const uint32_t k0 = lane < nc ? cand_key[lane] : ABSENT, k1 = lane + 32 < nc ? cand_key[lane + 32] : ABSENT;
      int r0 = 0, r1 = 0;
      for (int q = 0; q < nc; ++q) {
        const uint32_t kq = cand_key[q];
        r0 += (kq < k0 || (kq == k0 && q < lane)) ? 1 : 0;           // the list is in residue order: q orders the ties
        r1 += (kq < k1 || (kq == k1 && q < lane + 32)) ? 1 : 0;
      }
      const bool t0 = lane < nc && r0 < need, t1 = lane + 32 < nc && r1 < need;
      const uint32_t b0 = __ballot_sync(0xffffffffu, t0), b1 = __ballot_sync(0xffffffffu, t1);
      if (t0) { const int j = cand_idx[lane]; slot[base + nm + (int)__popc(b0 & lt_mask)] = j; key[j] = ABSENT; }
      if (t1) { const int j = cand_idx[lane + 32]; slot[base + nm + (int)__popc(b0) + (int)__popc(b1 & lt_mask)] = j; key[j] = ABSENT; }
    } else {
      // more than 64 keys equal to v (the search ran out of bits): sure members, then the first `need` of the ties
      int nm = 0, left = need;
      for (int s = 0; s < S; ++s) {
        const uint32_t kv = kl[32 * s];
        const bool m = kv < v, e = kv == v && kv != ABSENT;
        const uint32_t be = __ballot_sync(0xffffffffu, e);
        const bool take = m || (e && (int)__popc(be & lt_mask) < left);
        const uint32_t bt = __ballot_sync(0xffffffffu, take);
        if (take) { slot[base + nm + (int)__popc(bt & lt_mask)] = lane + 32 * s; kl[32 * s] = ABSENT; }
        nm += (int)__popc(bt);
        left -= min(left, (int)__popc(be));
      }
    }
    __syncwarp();
  };

  // the residue itself (d = 0) is always the first of its nearest neighbours: taking it out first lets the search start at
  // the bits in which real distances differ instead of walking down from the exponent's top bit
  int knn_rest = knn;
  if (knn > 0) {
    if (lane == 0) { slot[0] = i; key[i] = ABSENT; }
    knn_rest = knn - 1;
    __syncwarp();
  }
  select(knn_rest, knn - knn_rest);
  // ---- exponential race over the remaining residues: key = Exp(1) * d^3, keep the ns smallest
  if (ns > 0) {
    if (exp_noise != nullptr) {
      int members_below = 0;                   // kNN members with a smaller index (compacted column of the injected noise)
      for (int s = 0; s < S; ++s) {
        const int j = lane + 32 * s;
        const uint32_t kv = kl[32 * s];
        const bool member = j < N && kv == ABSENT;
        const uint32_t bal = __ballot_sync(0xffffffffu, member);
        if (j < N && !member) {
          const float e = exp_noise[(gbase + i) * (size_t)(N - knn) + (j - members_below - (int)__popc(bal & lt_mask))];
          const float d = fmaxf(__uint_as_float(kv), 1e-10f);
          kl[32 * s] = min(__float_as_uint(e * (d * d * d)), 0xFFFFFFFEu);
        }
        members_below += (int)__popc(bal);
      }
    } else {
      // one Philox block per four residues (the same counters as k_graph / k_graph_sel): lane takes blocks lane, lane + 32, ..
      const uint64_t strm = stream_base + (uint64_t)b;
      uint4* k4 = reinterpret_cast<uint4*>(key);
      for (int q = lane; q * 4 < npad; q += 32) {
        const uint4 r4 = dfm_rng(seed, strm, RNG_EDGE, fwd, (uint32_t)q, (uint32_t)i);
        uint4 kv = k4[q];
        auto race = [&](uint32_t dbits, uint32_t u) -> uint32_t {
          if (dbits == ABSENT) return ABSENT;
          const float d = fmaxf(__uint_as_float(dbits), 1e-10f);
          return min(__float_as_uint(-logf(u01_open(u)) * (d * d * d)), 0xFFFFFFFEu);
        };
        kv.x = race(kv.x, r4.x); kv.y = race(kv.y, r4.y); kv.z = race(kv.z, r4.z); kv.w = race(kv.w, r4.w);
        k4[q] = kv;
      }
    }
    __syncwarp();
    select(ns, knn);
  }
  __syncwarp();
  // ---- pair features for the K selected edges; pad slots point at the residue itself
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int k = lane + 32 * half;
    int j = i;
    uint32_t ft = 0;
    float r2 = 0.f;
    if (k < K) {
      j = min(max(slot[k], 0), N - 1);
      ft = pair_bins(pos, cb, (int)(gbase + i), (int)(gbase + j), i, j, R, &r2);
    }
    const size_t o = (gbase + i) * SLOTS + k;
    nbr[o] = j;
    feat[o] = ft;
    radial[o] = r2;
    {
      const uint32_t otp = (ft >> 6) & 0x3FFFu;
      const int drp = (int)(((otp == 0 ? 40u : 0u) + (ft & 63u)) * 66u + ((k < K) ? ((ft >> 20) & 127u) : 32u));
      const int oidx = (int)((((ft >> 6) & 31u) * 24u + ((ft >> 11) & 31u)) * 12u + ((ft >> 16) & 15u));
      const __half2 rh = __float2half2_rn(fminf(r2 * 0.03125f, 65000.f));
      emeta[o] = make_int4((int)(gbase + j), drp, otp == 0 ? -1 : oidx, *reinterpret_cast<const int*>(&rh));
    }
  }
}

template <int WARPS, int S>
static int launch_graph_sel_w(dfm_ctx* ctx, int B, const float* exp_noise, uint64_t seed, uint64_t stream_base,
                              uint32_t fwd_index, Workspace& ws, cudaStream_t s) {
  const long rows = (long)B * ctx->N;
  static int stage = -1;
  if (stage < 0) { const char* e = getenv("DFM_GRAPH_STAGE"); stage = e ? atoi(e) : 0; }
  if (!stage) {       // one row per warp, coordinates read through L1
    const int cpt = (ctx->N + WARPS - 1) / WARPS;
    k_graph_sel<WARPS, S, false><<<B * cpt, WARPS * 32, 0, s>>>(B, ctx->N, ctx->R, ctx->K, ctx->knn, ctx->ns, WARPS, cpt, ws.pos, ws.cb,
                                                               exp_noise, seed, stream_base, fwd_index, ws.nbr, ws.feat, ws.radial, ws.emeta);
    LAUNCH_CHECK(ctx);
    return 0;
  }
  const size_t dyn = (((size_t)ctx->N * 36 + 15 + 15) & ~(size_t)15) + (size_t)ctx->N * 16;   // staged coordinates of one trajectory
  static unsigned long long attr_devices = 0;
  if (dfm_once_per_device(attr_devices, ctx->device))
    CUDA_TRY(cudaFuncSetAttribute(k_graph_sel<WARPS, S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  // rows per CTA: enough CTAs for ~12 waves of 3 CTAs per SM, at least one row per warp, at most 8 rows per warp
  long rpc = rows / ((long)ctx->num_sms * 36);
  rpc = (rpc + WARPS - 1) / WARPS * WARPS;
  if (rpc < WARPS) rpc = WARPS;
  if (rpc > 8 * WARPS) rpc = 8 * WARPS;
  const int cpt = (int)((ctx->N + rpc - 1) / rpc);
  k_graph_sel<WARPS, S, true><<<B * cpt, WARPS * 32, dyn, s>>>(B, ctx->N, ctx->R, ctx->K, ctx->knn, ctx->ns, (int)rpc, cpt, ws.pos, ws.cb,
                                                              exp_noise, seed, stream_base, fwd_index, ws.nbr, ws.feat, ws.radial, ws.emeta);
  LAUNCH_CHECK(ctx);
  return 0;
}
// rows (warps) per CTA of the pair scan; DFM_GRAPH_WARPS = 2 / 4 / 8 overrides the default for the pair-tile sweep of
// BASELINE config #4 (profiles/c4_sweep.py -> profiles/r02/c4_pair_tile_sweep.txt)
template <int S>
static int launch_graph_sel(dfm_ctx* ctx, int B, const float* exp_noise, uint64_t seed, uint64_t stream_base,
                            uint32_t fwd_index, Workspace& ws, cudaStream_t s) {
  // default: 8 rows per CTA up to 16 keys per lane, 4 at 32 keys per lane (measured at 2x400 residues: 0.98 vs 1.22 ms,
  // the 8-warp CTA of the 32-key instantiation is limited to one CTA per scheduler by its registers)
  static int warps_env = -1;
  if (warps_env < 0) { const char* e = getenv("DFM_GRAPH_WARPS"); warps_env = e ? atoi(e) : 0; }
  const int warps = warps_env ? warps_env : (S >= 32 ? 4 : 8);
  if (warps == 4) return launch_graph_sel_w<4, S>(ctx, B, exp_noise, seed, stream_base, fwd_index, ws, s);
  if (warps == 2) return launch_graph_sel_w<2, S>(ctx, B, exp_noise, seed, stream_base, fwd_index, ws, s);
  return launch_graph_sel_w<8, S>(ctx, B, exp_noise, seed, stream_base, fwd_index, ws, s);
}

int launch_graph(dfm_ctx* ctx, int B, bool generic, const int32_t* edges, const float* exp_noise, uint64_t seed,
                 uint64_t stream_base, uint32_t fwd_index, Workspace& ws, cudaStream_t s) {
  constexpr int WARPS = 8;
  const int N = ctx->N;
  // register-resident selection for the common sizes (no injected edge table); larger complexes use the smem kernel
  static int use_sel = -1;
  if (use_sel < 0) { const char* e = getenv("DFM_GRAPH_KERNEL"); use_sel = e ? atoi(e) : 1; }
  if (use_sel && !generic && edges == nullptr && N >= 2) {
    if (N <= 320) return launch_graph_sel<10>(ctx, B, exp_noise, seed, stream_base, fwd_index, ws, s);
    if (N <= 512) return launch_graph_sel<16>(ctx, B, exp_noise, seed, stream_base, fwd_index, ws, s);
    if (N <= 1024) return launch_graph_sel<32>(ctx, B, exp_noise, seed, stream_base, fwd_index, ws, s);
  }
  const int npad = (N + 31) & ~31;
  static int use_big = -1;
  if (use_big < 0) { const char* e = getenv("DFM_GRAPH_BIG"); use_big = e ? atoi(e) : 1; }
  if (use_big && use_sel && !generic && edges == nullptr && N >= 2) {
    // large complexes: two-phase selection with the keys in shared memory (4 rows per CTA above 6 k residues)
    const size_t per_warp = (size_t)(npad + 192) * sizeof(uint32_t);
    const long rows = (long)B * N;
    static unsigned long long big_devices = 0;
    if (dfm_once_per_device(big_devices, ctx->device)) {
      CUDA_TRY(cudaFuncSetAttribute(k_graph_big<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      CUDA_TRY(cudaFuncSetAttribute(k_graph_big<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    if (8 * per_warp <= 200 * 1024) {
      k_graph_big<8><<<(int)((rows + 7) / 8), 256, 8 * per_warp, s>>>(B, N, ctx->R, ctx->K, ctx->knn, ctx->ns, ws.pos, ws.cb, exp_noise,
                                                                    seed, stream_base, fwd_index, ws.nbr, ws.feat, ws.radial, ws.emeta);
      LAUNCH_CHECK(ctx);
      return 0;
    }
    if (2 * per_warp <= 200 * 1024) {
      k_graph_big<2><<<(int)((rows + 1) / 2), 64, 2 * per_warp, s>>>(B, N, ctx->R, ctx->K, ctx->knn, ctx->ns, ws.pos, ws.cb, exp_noise,
                                                                   seed, stream_base, fwd_index, ws.nbr, ws.feat, ws.radial, ws.emeta);
      LAUNCH_CHECK(ctx);
      return 0;
    }
  }
  const size_t smem = (size_t)WARPS * (npad + 32) * sizeof(float);
  if (smem > 200 * 1024) {
    dfm_set_error("complex too large for the graph kernel (N=%d)", N);
    return DFM_EINVAL;
  }
  static unsigned long long attr_devices = 0;
  if (dfm_once_per_device(attr_devices, ctx->device)) {
    CUDA_TRY(cudaFuncSetAttribute(k_graph<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  const long rows = (long)B * N;
  const int grid = (int)((rows + WARPS - 1) / WARPS);
  k_graph<WARPS><<<grid, WARPS * 32, smem, s>>>(B, N, ctx->R, ctx->K, ctx->knn, ctx->ns, ws.pos, ws.cb, edges,
                                                exp_noise, seed, stream_base, fwd_index, ws.nbr, ws.feat,
                                                ws.radial, ws.emeta);
  LAUNCH_CHECK(ctx);
  return 0;
}
