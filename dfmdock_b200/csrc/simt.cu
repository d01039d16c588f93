// fp32 FFMA ("parity mode", DFM_PRECISION_FP32) versions of the two GEMM-shaped operators.
// They exist to (a) give a <=1e-4 comparison against the fp32 reference and (b) cross-check the tcgen05
// kernels on the GPU; the throughput path is tc.cu.
//
// Restates: E_GCL.edge_model / coord_model / segment sums   src/models/egnn.py:95-148, 11-26
//           nn.Linear                                      (node_mlp, to_energy halves)
#include "common.cuh"

#define LDS 260   // padded row stride of the activation tile in shared memory (floats)

// acc[4][16] += S[64 x 256] (smem) * W[256 x 256]^T, W row-major with leading dim ldw starting at column col0.
// 256 threads; thread (ty = tid/16, tx = tid%16) owns rows 4ty..4ty+3 and columns tx + 16c.
__device__ __forceinline__ void tile_gemm_64x256(const float* __restrict__ S, float* __restrict__ Wc,
                                                 const float* __restrict__ W, int ldw, int col0,
                                                 float (&acc)[4][16]) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[r][c] = 0.f;
  for (int k0 = 0; k0 < H; k0 += 16) {
    __syncthreads();
    {  // thread n stages W[n, col0+k0 .. +16) as Wc[k][n]
      const float* src = W + (size_t)tid * ldw + col0 + k0;
#pragma unroll
      for (int k = 0; k < 16; ++k) Wc[k * 256 + tid] = __ldg(src + k);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) a[r] = S[(ty * 4 + r) * LDS + k0 + k];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const float w = Wc[k * 256 + tx + 16 * c];
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r][c] = fmaf(a[r], w, acc[r][c]);
      }
    }
  }
}

// sum over the 16 tx lanes that share a row (they are 16 consecutive lanes of one warp)
__device__ __forceinline__ float row_sum16(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_linear_simt(LinearArgs a) {
  extern __shared__ float sm[];
  float* S = sm;
  float* Wc = sm + 64 * LDS;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int row0 = blockIdx.x * 64;
  for (int r = 0; r < 64; ++r) {
    const int m = row0 + r;
    S[r * LDS + tid] = (m < a.M) ? a.A[(size_t)m * H + tid] : 0.f;
  }
  float acc[4][16];
  tile_gemm_64x256(S, Wc, a.W32, a.ldw, a.w_col0, acc);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int m = row0 + ty * 4 + r;
    if (m >= a.M) continue;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const int col = tx + 16 * c;
      float v = acc[r][c];
      if (a.bias) v += a.bias[col];
      if (a.add) v += a.add[(size_t)m * H + col];
      if (a.out) a.out[(size_t)m * H + col] = v;
      if (a.out16) a.out16[(size_t)m * H + col] = __float2half_rn(v);
    }
  }
}

int launch_linear_simt(dfm_ctx* ctx, const LinearArgs& a, cudaStream_t s) {
  const size_t smem = (64 * LDS + 16 * 256) * sizeof(float);
  static unsigned long long attr_devices = 0;
  if (dfm_once_per_device(attr_devices, ctx->device)) {
    CUDA_TRY(cudaFuncSetAttribute(k_linear_simt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  k_linear_simt<<<(a.M + 63) / 64, 256, smem, s>>>(a);
  LAUNCH_CHECK(ctx);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// One block per (trajectory, residue): its <=60 edges are rows of a 64 x 256 tile.
__global__ void __launch_bounds__(256)
k_edge_simt(EdgeArgs a, const float* __restrict__ T, const float* __restrict__ w1r, const float* __restrict__ W2,
            const float* __restrict__ b2, const float* __restrict__ wa, const float* __restrict__ ba,
            const float* __restrict__ Wc1, const float* __restrict__ bc1, const float* __restrict__ wc2) {
  extern __shared__ float sm[];
  float* S = sm;                      // [64][LDS]
  float* Wc = sm + 64 * LDS;          // [16][256], reused as the column-reduction scratch
  __shared__ int s_j[64];
  __shared__ uint32_t s_ft[64];
  __shared__ float s_rad[64];
  __shared__ float s_cw[64];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int node = blockIdx.x;                 // b*N + i
  const int b = node / a.N, i = node % a.N;
  const float* Bm = reinterpret_cast<const float*>(a.Bm);
  if (tid < 64) {
    s_j[tid] = a.nbr[(size_t)node * SLOTS + tid];
    s_ft[tid] = a.feat[(size_t)node * SLOTS + tid];
    s_rad[tid] = a.radial[(size_t)node * SLOTS + tid];
  }
  __syncthreads();
  {  // u = A_i + B_j + radial*w1r + sum of five table rows, S = SiLU(u)
    const float ai = a.A[(size_t)node * H + tid];
    const float wr = w1r[tid];
    for (int k = 0; k < 64; ++k) {
      float s = 0.f;
      if (k < a.K) {
        const uint32_t ft = s_ft[k];
        float u = ai + Bm[((size_t)b * a.N + s_j[k]) * H + tid];
        u = fmaf(s_rad[k], wr, u);
        u += T[(ft & 63u) * H + tid];
        u += T[(40u + ((ft >> 6) & 31u)) * H + tid];
        u += T[(64u + ((ft >> 11) & 31u)) * H + tid];
        u += T[(88u + ((ft >> 16) & 15u)) * H + tid];
        u += T[(100u + ((ft >> 20) & 127u)) * H + tid];
        s = silu_acc(u);
      }
      S[k * LDS + tid] = s;
    }
  }
  float acc[4][16];
  tile_gemm_64x256(S, Wc, W2, H, 0, acc);
  // m = SiLU(. + b2); gate = sigmoid(wa.m + ba); m* = gate*m  (egnn.py:95-104)
  float colsum[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) colsum[c] = 0.f;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    float dotp = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const int col = tx + 16 * c;
      acc[r][c] = silu_acc(acc[r][c] + b2[col]);
      dotp = fmaf(acc[r][c], wa[col], dotp);
    }
    dotp = row_sum16(dotp);
    const float g = (ty * 4 + r < a.K) ? sigmoid_acc(dotp + ba[0]) : 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      acc[r][c] *= g;
      colsum[c] += acc[r][c];
    }
  }
  __syncthreads();   // everyone is done reading S / Wc
#pragma unroll
  for (int c = 0; c < 16; ++c) Wc[ty * 256 + tx + 16 * c] = colsum[c];
  const bool do_coord = a.last && i >= a.R;
  if (do_coord) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 16; ++c) S[(ty * 4 + r) * LDS + tx + 16 * c] = acc[r][c];
  }
  __syncthreads();
  {
    float sacc = 0.f;
#pragma unroll
    for (int q = 0; q < 16; ++q) sacc += Wc[q * 256 + tid];
    a.agg[(size_t)node * H + tid] = sacc;
  }
  if (!do_coord) return;
  // coordinate head (egnn.py:118-137): w = clamp(wc2 . SiLU(Wc1 m* + bc1), +-2); x_i += mean_k diffn_k w_k
  tile_gemm_64x256(S, Wc, Wc1, H, 0, acc);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    float dotp = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const int col = tx + 16 * c;
      dotp = fmaf(silu_acc(acc[r][c] + bc1[col]), wc2[col], dotp);
    }
    dotp = row_sum16(dotp);
    if (tx == 0) s_cw[ty * 4 + r] = fminf(fmaxf(dotp, -2.f), 2.f);
  }
  __syncthreads();
  if (tid < 32) {
    float fx = 0.f, fy = 0.f, fz = 0.f;
    const float* pi = a.pos + (size_t)node * 9 + 3;
    for (int k = tid; k < a.K; k += 32) {
      const float* pj = a.pos + ((size_t)b * a.N + s_j[k]) * 9 + 3;
      const float dx = pi[0] - pj[0], dy = pi[1] - pj[1], dz = pi[2] - pj[2];
      const float rad = dx * dx + dy * dy + dz * dz;
      const float sc = s_cw[k] / (sqrtf(rad + 1e-8f) + 1.0f);
      fx += dx * sc; fy += dy * sc; fz += dz * sc;
    }
    fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz);
    if (tid == 0) {
      float* fo = a.fbuf + ((size_t)b * (a.N - a.R) + (i - a.R)) * 4;
      const float inv = 1.f / (float)a.K;
      fo[0] = fx * inv; fo[1] = fy * inv; fo[2] = fz * inv; fo[3] = 0.f;
    }
  }
}

int launch_edge_simt(dfm_ctx* ctx, const EdgeArgs& a, cudaStream_t s) {
  const LayerW& w = ctx->layer[a.layer];
  const size_t smem = (64 * LDS + 16 * 256) * sizeof(float);
  static unsigned long long attr_devices = 0;
  if (dfm_once_per_device(attr_devices, ctx->device)) {
    CUDA_TRY(cudaFuncSetAttribute(k_edge_simt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  k_edge_simt<<<a.B * a.N, 256, smem, s>>>(a, w.T32, w.w1r, w.W2, w.b2, w.wa, w.ba, w.Wc1, w.bc1, w.wc2);
  LAUNCH_CHECK(ctx);
  return 0;
}
