// Per-node / per-trajectory small kernels: weight repacking, GraphNorm, force-and-torque head, energy head.
//
// Restates: GraphNorm (torch_geometric 2.6.0, call site src/models/egnn.py:74)
//           heads of Score_Net.forward            src/models/score_net_mlsb.py:382-411
//           GaussianFourierProjection             src/models/score_net_mlsb.py:162-172
#include "common.cuh"

// ---- weight repacking ----------------------------------------------------------------------------
// fp16 image of W[256, 256] (rows = output features, K contiguous) in the tcgen05 K-major SWIZZLE_128B
// shared-memory layout: 4 K-blocks of [256 rows x 128 B]; 16-byte chunk c of row n sits at chunk c ^ (n & 7).
__global__ void k_image_pack(const float* __restrict__ W, int ldw, int col0, float scale, __half* __restrict__ img) {
  const int n = blockIdx.x, k = threadIdx.x;
  const int kb = k >> 6, c = (k & 63) >> 3, e = k & 7;
  img[(((size_t)kb * 256 + n) * 8 + (c ^ (n & 7))) * 8 + e] = __float2half_rn(W[(size_t)n * ldw + col0 + k] * scale);
}
// Same image with K permuted into the order in which edge_ws.cu spills the gated messages: position
// p = ch*128 + v*64 + cq*16 + 2 jl + e holds column c = ch*128 + 8 (8 v + jl) + 2 cq + e (v = 0..1, cq = 0..3, jl = 0..7,
// e = 0..1): the eight registers of one 256-bit store are contiguous and the four lanes of a row write one full 128-byte line.
__global__ void k_image_pack_perm(const float* __restrict__ W, int ldw, float scale, __half* __restrict__ img) {
  const int n = blockIdx.x, p = threadIdx.x;
  const int q = p & 127, src = (p & 128) + 8 * (8 * (q >> 6) + ((q & 15) >> 1)) + 2 * ((q >> 4) & 3) + (q & 1);
  const int kb = p >> 6, c = (p & 63) >> 3, e = p & 7;
  img[(((size_t)kb * 256 + n) * 8 + (c ^ (n & 7))) * 8 + e] = __float2half_rn(W[(size_t)n * ldw + src] * scale);
}
int launch_image_pack_perm(dfm_ctx* ctx, const float* W, int ldw, float scale, __half* img, cudaStream_t s) {
  k_image_pack_perm<<<256, 256, 0, s>>>(W, ldw, scale, img);
  LAUNCH_CHECK(ctx);
  return 0;
}
int launch_image_pack(dfm_ctx* ctx, const float* W, int ldw, int col0, float scale, __half* img, cudaStream_t s) {
  k_image_pack<<<256, 256, 0, s>>>(W, ldw, col0, scale, img);
  LAUNCH_CHECK(ctx);
  return 0;
}

// T_l[r][c] = sum_e W1e[c][e] * [Ws | Wp][e][r]   (SURVEY App. A.5), plus the radial column w1r.
__global__ void k_pair_table(const float* __restrict__ W1, const float* __restrict__ Ws, const float* __restrict__ Wp,
                             int P, float* __restrict__ T32, float* __restrict__ w1r) {
  const int r = blockIdx.x, c = threadIdx.x;
  float acc = 0.f;
  for (int e = 0; e < ED; ++e) {
    const float emb = (r < NSPATIAL) ? Ws[e * NSPATIAL + r] : Wp[e * P + (r - NSPATIAL)];
    acc = fmaf(W1[(size_t)c * 641 + 513 + e], emb, acc);
  }
  T32[(size_t)r * H + c] = acc;
  if (r == 0) w1r[c] = W1[(size_t)c * 641 + 512];
}
// Merged gather tables for the tensor-core edge kernel (stored pre-halved, see common.cuh): one row per (dist bin, relpos bin) pair -- with the three
// zero-angle rows folded in for pairs whose angle bins are all 0 (masked: dist >= 22 A or self) -- and one row per
// (omega, theta, phi) bin triple.  Sums are formed in fp32 and rounded to fp16 once.
__global__ void k_merge_tables(const float* __restrict__ T32, __half* __restrict__ Tdrph, __half* __restrict__ Totph) {
  const int row = blockIdx.x, c = threadIdx.x;
  if (row < 2 * 40 * 66) {
    const int z = row / (40 * 66), d = (row / 66) % 40, rp = row % 66;
    float v = T32[(size_t)d * H + c] + T32[(size_t)(NSPATIAL + rp) * H + c];
    if (z) v += T32[(size_t)40 * H + c] + T32[(size_t)64 * H + c] + T32[(size_t)88 * H + c];
    Tdrph[(size_t)row * H + c] = __float2half_rn(0.5f * v);
  } else {
    const int q = row - 2 * 40 * 66;
    const int o = q / (24 * 12), t = (q / 12) % 24, ph = q % 12;
    const float v = T32[(size_t)(40 + o) * H + c] + T32[(size_t)(64 + t) * H + c] + T32[(size_t)(88 + ph) * H + c];
    Totph[(size_t)q * H + c] = __float2half_rn(0.5f * v);
  }
}
int launch_pair_table(dfm_ctx* ctx, int l, cudaStream_t s) {
  LayerW& w = ctx->layer[l];
  k_pair_table<<<NSPATIAL + ctx->P, 256, 0, s>>>(w.W1, ctx->w["spatial_embed.weight"].d,
                                                ctx->w["positional_embed.weight"].d, ctx->P, w.T32, w.w1r);
  LAUNCH_CHECK(ctx);
  k_merge_tables<<<2 * 40 * 66 + 24 * 24 * 12, 256, 0, s>>>(w.T32, w.Tdrp16h, w.Totp16h);
  LAUNCH_CHECK(ctx);
  return 0;
}

// h0 = single_embed([rec_x; lig_x])  (score_net_mlsb.py:365-366), once per complex.
__global__ void __launch_bounds__(256) k_single_embed(int R, int x_dim, const float* __restrict__ rec_x,
                                                     const float* __restrict__ lig_x, const float* __restrict__ W,
                                                     float* __restrict__ h0) {
  const int n = blockIdx.x, c = threadIdx.x;
  const float* x = (n < R) ? rec_x + (size_t)n * x_dim : lig_x + (size_t)(n - R) * x_dim;
  __shared__ float xs[256];
  float acc = 0.f;
  for (int k0 = 0; k0 < x_dim; k0 += 256) {
    __syncthreads();
    xs[c] = (k0 + c < x_dim) ? x[k0 + c] : 0.f;
    __syncthreads();
    const int kn = min(256, x_dim - k0);
    const float* wrow = W + (size_t)c * x_dim + k0;
    for (int k = 0; k < kn; ++k) acc = fmaf(xs[k], __ldg(wrow + k), acc);
  }
  h0[(size_t)n * H + c] = acc;
}
int launch_single_embed(dfm_ctx* ctx, const float* rec_x, const float* lig_x, cudaStream_t s) {
  k_single_embed<<<ctx->N, 256, 0, s>>>(ctx->R, ctx->x_dim, rec_x, lig_x, ctx->w["single_embed.weight"].d, ctx->h0);
  LAUNCH_CHECK(ctx);
  return 0;
}

__global__ void k_broadcast_h0(size_t per, const float4* __restrict__ h0, float4* __restrict__ h, uint2* __restrict__ h16) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < per) {
    const float4 v = h0[idx];
    h[(size_t)blockIdx.y * per + idx] = v;
    if (h16) {
      const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
      h16[(size_t)blockIdx.y * per + idx] = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
    }
  }
}
int launch_broadcast_h0(dfm_ctx* ctx, int B, Workspace& ws, cudaStream_t s) {
  const size_t per = (size_t)ctx->N * H / 4;
  dim3 grid((unsigned)((per + 255) / 256), B);
  k_broadcast_h0<<<grid, 256, 0, s>>>(per, reinterpret_cast<const float4*>(ctx->h0), reinterpret_cast<float4*>(ws.h),
                                      reinterpret_cast<uint2*>(ws.h16));
  LAUNCH_CHECK(ctx);
  return 0;
}

// ---- GraphNorm + SiLU ------------------------------------------------------------------------------
// y = SiLU(weight * (z - mean*mean_scale) / sqrt(mean((z - mean*mean_scale)^2) + 1e-5) + bias), statistics over the
// N residues of one trajectory, per feature.  grid (B, 8): 32 columns per block, 8 row lanes.
__global__ void __launch_bounds__(256) k_graphnorm_silu(int N, const float* __restrict__ z, const float* __restrict__ gw,
                                                       const float* __restrict__ gb, const float* __restrict__ gms,
                                                       float* __restrict__ y) {
  const int b = blockIdx.x, c = blockIdx.y * 32 + (threadIdx.x & 31), ry = threadIdx.x >> 5;
  const float* zb = z + (size_t)b * N * H;
  float* yb = y + (size_t)b * N * H;
  __shared__ float red[8][32];
  float s = 0.f;
  for (int n = ry; n < N; n += 8) s += zb[(size_t)n * H + c];
  red[ry][threadIdx.x & 31] = s;
  __syncthreads();
  float mean = 0.f;
#pragma unroll
  for (int q = 0; q < 8; ++q) mean += red[q][threadIdx.x & 31];
  mean /= (float)N;
  const float shift = mean * gms[c];
  __syncthreads();
  float v = 0.f;
  for (int n = ry; n < N; n += 8) {
    const float o = zb[(size_t)n * H + c] - shift;
    v = fmaf(o, o, v);
  }
  red[ry][threadIdx.x & 31] = v;
  __syncthreads();
  float var = 0.f;
#pragma unroll
  for (int q = 0; q < 8; ++q) var += red[q][threadIdx.x & 31];
  var /= (float)N;
  const float sd = sqrtf(var + 1e-5f);
  const float w = gw[c], bi = gb[c];
  for (int n = ry; n < N; n += 8) {
    const float o = zb[(size_t)n * H + c] - shift;
    yb[(size_t)n * H + c] = silu_acc(w * o / sd + bi);
  }
}
// Statistics only (throughput path): gscale = weight * rstd, gshift = bias - mean*mean_scale*gscale, so that the
// consumer (node_tc.cu MODE_H) forms SiLU(z * gscale + gshift) while it builds its operand tile.
__global__ void __launch_bounds__(256) k_graphnorm_stats(int N, const float* __restrict__ z, const float* __restrict__ gw,
                                                        const float* __restrict__ gb, const float* __restrict__ gms,
                                                        float* __restrict__ gscale, float* __restrict__ gshift) {
  pdl_trigger();            // programmatic dependent launch, see common.cuh
  pdl_wait();
  const int b = blockIdx.x, c = blockIdx.y * 32 + (threadIdx.x & 31), ry = threadIdx.x >> 5;
  const float* zb = z + (size_t)b * N * H;
  __shared__ float red[8][32];
  float s = 0.f;
  for (int n = ry; n < N; n += 8) s += zb[(size_t)n * H + c];
  red[ry][threadIdx.x & 31] = s;
  __syncthreads();
  float mean = 0.f;
#pragma unroll
  for (int q = 0; q < 8; ++q) mean += red[q][threadIdx.x & 31];
  mean /= (float)N;
  const float shift = mean * gms[c];
  __syncthreads();
  float v = 0.f;
  for (int n = ry; n < N; n += 8) {
    const float o = zb[(size_t)n * H + c] - shift;
    v = fmaf(o, o, v);
  }
  red[ry][threadIdx.x & 31] = v;
  __syncthreads();
  if (ry == 0) {
    float var = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) var += red[q][threadIdx.x & 31];
    var /= (float)N;
    const float sc = gw[c] / sqrtf(var + 1e-5f);
    gscale[(size_t)b * H + c] = sc;
    gshift[(size_t)b * H + c] = gb[c] - shift * sc;
  }
}
int launch_graphnorm_stats(dfm_ctx* ctx, int B, int layer, const float* z, float* gscale, float* gshift, cudaStream_t s) {
  const LayerW& w = ctx->layer[layer];
  dim3 grid(B, 8);
  CUDA_TRY(dfm_launch_pdl(k_graphnorm_stats, grid, dim3(256), 0, s, ctx->N, z, w.gn_w, w.gn_b, w.gn_ms, gscale, gshift));
  LAUNCH_CHECK(ctx);
  return 0;
}

int launch_graphnorm_silu(dfm_ctx* ctx, int B, int layer, Workspace& ws, cudaStream_t s) {
  const LayerW& w = ctx->layer[layer];
  dim3 grid(B, 8);
  k_graphnorm_silu<<<grid, 256, 0, s>>>(ctx->N, ws.z, w.gn_w, w.gn_b, w.gn_ms, ws.y);
  LAUNCH_CHECK(ctx);
  return 0;
}

// ---- force / torque head -----------------------------------------------------------------------------
__device__ __forceinline__ float block_sum128(float v, float* red) {
  const int tid = threadIdx.x;
  v = warp_sum(v);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  return red[0] + red[1] + red[2] + red[3];
}

struct HeadW {
  const float* t_W; const float* t_lin;
  const float* W1[2]; const float* lnw[2]; const float* lnb[2]; const float* w2[2];
};

__global__ void __launch_bounds__(128) k_force_head(int N, int R, const float* __restrict__ t, const float* __restrict__ pos,
                                                   const float* __restrict__ fbuf, HeadW hw, float* __restrict__ tr_score,
                                                   float* __restrict__ rot_score, float* __restrict__ f_out) {
  const int b = blockIdx.x, tid = threadIdx.x, L = N - R;
  __shared__ float red[4];
  __shared__ float four[128];
  __shared__ float temb[128];
  // tr_pred = mean f ; rot_pred = mean r x f   (score_net_mlsb.py:393-404)
  float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int l = tid; l < L; l += 128) {
    const float* f = fbuf + ((size_t)b * L + l) * 4;
    const float* r = pos + ((size_t)b * N + R + l) * 9 + 3;
    const float fx = f[0], fy = f[1], fz = f[2];
    a[0] += fx; a[1] += fy; a[2] += fz;
    a[3] += r[1] * fz - r[2] * fy;
    a[4] += r[2] * fx - r[0] * fz;
    a[5] += r[0] * fy - r[1] * fx;
    if (f_out) {
      float* fo = f_out + ((size_t)b * L + l) * 3;
      fo[0] = fx; fo[1] = fy; fo[2] = fz;
    }
  }
  float pred[6];
#pragma unroll
  for (int q = 0; q < 6; ++q) pred[q] = block_sum128(a[q], red) / (float)L;
  // time embedding (score_net_mlsb.py:162-172, 305-309)
  {
    const float tt = t[b];
    const int k = tid & 63;
    const float proj = tt * hw.t_W[k] * 2.f * 3.14159265358979323846f;
    four[tid] = (tid < 64) ? sinf(proj) : cosf(proj);
  }
  __syncthreads();
  {
    float acc = 0.f;
    const float* wc = hw.t_lin + tid;                      // [in, out]: coalesced over the CTA
#pragma unroll 32
    for (int k = 0; k < 128; ++k) acc = fmaf(__ldg(wc + k * 128), four[k], acc);
    temb[tid] = sigmoid_acc(acc);
  }
  __syncthreads();
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    const float vx = pred[which * 3], vy = pred[which * 3 + 1], vz = pred[which * 3 + 2];
    const float nrm = sqrtf(vx * vx + vy * vy + vz * vz);
    const float* wc = hw.W1[which] + tid;                  // [in 129, out 128]
    float yv = __ldg(wc) * nrm;
#pragma unroll 32
    for (int k = 0; k < 128; ++k) yv = fmaf(__ldg(wc + (1 + k) * 128), temb[k], yv);
    const float mean = block_sum128(yv, red) / 128.f;
    const float dv = yv - mean;
    const float var = block_sum128(dv * dv, red) / 128.f;
    const float ln = dv / sqrtf(var + 1e-5f) * hw.lnw[which][tid] + hw.lnb[which][tid];
    const float pre = block_sum128(silu_acc(ln) * hw.w2[which][tid], red);
    const float sp = pre > 20.f ? pre : log1pf(expf(pre));
    if (tid == 0) {
      float* o = (which == 0 ? tr_score : rot_score) + (size_t)b * 3;
      const float sc = sp / (nrm + 1e-6f);
      o[0] = vx * sc; o[1] = vy * sc; o[2] = vz * sc;
    }
  }
}
int launch_force_head(dfm_ctx* ctx, int B, const float* t, Workspace& ws, float* tr_score, float* rot_score,
                      float* f_out, cudaStream_t s) {
  HeadW hw;
  hw.t_W = ctx->t_W; hw.t_lin = ctx->t_lin;
  for (int q = 0; q < 2; ++q) { hw.W1[q] = ctx->sc_W1[q]; hw.lnw[q] = ctx->sc_lnw[q]; hw.lnb[q] = ctx->sc_lnb[q]; hw.w2[q] = ctx->sc_w2[q]; }
  k_force_head<<<B, 128, 0, s>>>(ctx->N, ctx->R, t, ws.pos, ws.fbuf, hw, tr_score, rot_score, f_out);
  LAUNCH_CHECK(ctx);
  return 0;
}

// ---- energy head ---------------------------------------------------------------------------------------
// E_rl = we . SiLU(LN(Wer h_r + Wel h_l)), energy = sum_{D_rl < cut} E_rl / (count + 1e-6)  (score_net_mlsb.py:385-390).
// U = h Wer^T and V = h Wel^T are [B,N,256] GEMMs done by the caller; one warp per (b, r) walks the ligand.
__global__ void __launch_bounds__(256) k_energy_pairs(int B, int N, int R, float cut, const float* __restrict__ U,
                                                     const float* __restrict__ V, const float* __restrict__ pos,
                                                     const float* __restrict__ lnw, const float* __restrict__ lnb,
                                                     const float* __restrict__ we, float* __restrict__ esum) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long row = (long)blockIdx.x * 8 + warp;
  if (row >= (long)B * R) return;
  const int b = (int)(row / R), r = (int)(row % R), L = N - R;
  const float* u = U + ((size_t)b * N + r) * H + lane * 8;
  float ur[8], gw[8], gb[8], ew[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) { ur[q] = u[q]; gw[q] = lnw[lane * 8 + q]; gb[q] = lnb[lane * 8 + q]; ew[q] = we[lane * 8 + q]; }
  const float* pr = pos + ((size_t)b * N + r) * 9 + 3;
  const float rx = pr[0], ry = pr[1], rz = pr[2];
  float sum = 0.f, cnt = 0.f, clash = 0.f;
  for (int l = 0; l < L; ++l) {
    const float* pl = pos + ((size_t)b * N + R + l) * 9 + 3;
    const float dx = rx - pl[0], dy = ry - pl[1], dz = rz - pl[2];
    const float d = sqrtf(dx * dx + dy * dy + dz * dz);
    if (d <= 3.0f) clash += 1.f;
    if (!(d < cut)) continue;
    const float* v = V + ((size_t)b * N + R + l) * H + lane * 8;
    float x[8], s = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) { x[q] = ur[q] + v[q]; s += x[q]; }
    const float mean = warp_sum(s) / 256.f;
    float vs = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) { x[q] -= mean; vs = fmaf(x[q], x[q], vs); }
    const float rstd = 1.f / sqrtf(warp_sum(vs) / 256.f + 1e-5f);
    float e = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) e = fmaf(silu_acc(x[q] * rstd * gw[q] + gb[q]), ew[q], e);
    sum += warp_sum(e);
    cnt += 1.f;
  }
  if (lane == 0) {
    float* o = esum + row * 4;
    o[0] = sum; o[1] = cnt; o[2] = clash; o[3] = 0.f;
  }
}
__global__ void k_energy_finish(int R, const float* __restrict__ esum, float* __restrict__ energy, int32_t* __restrict__ clashes) {
  const int b = blockIdx.x, lane = threadIdx.x;
  float s = 0.f, c = 0.f, k = 0.f;
  for (int r = lane; r < R; r += 32) {
    const float* o = esum + ((size_t)b * R + r) * 4;
    s += o[0]; c += o[1]; k += o[2];
  }
  s = warp_sum(s); c = warp_sum(c); k = warp_sum(k);
  if (lane == 0) {
    if (energy) energy[b] = s / (c + 1e-6f);
    if (clashes) clashes[b] = (int32_t)(k + 0.5f);
  }
}
int launch_energy(dfm_ctx* ctx, int B, bool fp32_path, Workspace& ws, float* energy, int32_t* clashes, cudaStream_t s) {
  // U -> ws.A, V -> ws.z (both free after the last layer)
  LinearArgs la{};
  la.A = ws.h; la.a_scale = 1.f; la.W32 = ctx->We; la.ldw = 512; la.bias = nullptr; la.add = nullptr;
  la.out16 = nullptr; la.M = B * ctx->N;
  la.w_col0 = 0; la.Wimg = ctx->img_WeR; la.out = ws.A;
  int rc = fp32_path ? launch_linear_simt(ctx, la, s) : launch_linear_tc(ctx, la, s);
  if (rc) return rc;
  la.w_col0 = 256; la.Wimg = ctx->img_WeL; la.out = ws.z;
  rc = fp32_path ? launch_linear_simt(ctx, la, s) : launch_linear_tc(ctx, la, s);
  if (rc) return rc;
  const long rows = (long)B * ctx->R;
  k_energy_pairs<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(B, ctx->N, ctx->R, ctx->cut_off, ws.A, ws.z, ws.pos,
                                                          ctx->e_ln_w, ctx->e_ln_b, ctx->e_w, ws.esum);
  LAUNCH_CHECK(ctx);
  k_energy_finish<<<B, 32, 0, s>>>(ctx->R, ws.esum, energy, clashes);
  LAUNCH_CHECK(ctx);
  return 0;
}

// ---- interface-residue head ---------------------------------------------------------------------------
// ires = W5 SiLU(W3 SiLU(W1 h + b1) + b3) + b5, to_ires = Linear(256,512) SiLU Linear(512,512) SiLU Linear(512,1)
// (src/models/score_net_mlsb.py:296-302, applied at :383).  Never read at inference (inference_base.py:494-500), so it is
// only computed on request (dfm_interface_logits): one CTA per residue, weights pre-transposed to [in, out] so that the
// 512 threads of a CTA read consecutive addresses.
__global__ void k_transpose(const float* __restrict__ W, int rows, int cols, float* __restrict__ Wt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows * cols) { const int r = i / cols, c = i % cols; Wt[(size_t)c * rows + r] = W[i]; }
}
int launch_transpose(dfm_ctx* ctx, const float* W, int rows, int cols, float* Wt, cudaStream_t s) {
  k_transpose<<<(rows * cols + 255) / 256, 256, 0, s>>>(W, rows, cols, Wt);
  LAUNCH_CHECK(ctx);
  return 0;
}
__global__ void __launch_bounds__(512) k_ires(const float* __restrict__ h, const float* __restrict__ W1t,
                                              const float* __restrict__ b1, const float* __restrict__ W3t,
                                              const float* __restrict__ b3, const float* __restrict__ w5,
                                              const float* __restrict__ b5, float* __restrict__ out) {
  __shared__ float x[512];
  __shared__ float red[16];
  const int o = threadIdx.x;
  const size_t row = blockIdx.x;
  if (o < H) x[o] = h[row * H + o];
  __syncthreads();
  float a = b1[o];
  for (int k = 0; k < H; ++k) a = fmaf(x[k], W1t[(size_t)k * 512 + o], a);
  a = silu_acc(a);
  __syncthreads();
  x[o] = a;
  __syncthreads();
  float c = b3[o];
  for (int k = 0; k < 512; ++k) c = fmaf(x[k], W3t[(size_t)k * 512 + o], c);
  float v = warp_sum(silu_acc(c) * w5[o]);
  if ((o & 31) == 0) red[o >> 5] = v;
  __syncthreads();
  if (o < 16) {
    v = red[o];
#pragma unroll
    for (int d = 8; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffu, v, d);
    if (o == 0) out[row] = v + b5[0];
  }
}
int launch_ires(dfm_ctx* ctx, int rows, const float* h, float* out, cudaStream_t s) {
  k_ires<<<rows, 512, 0, s>>>(h, ctx->ires_W1t, ctx->ires_b1, ctx->ires_W3t, ctx->ires_b3, ctx->ires_w5, ctx->ires_b5, out);
  LAUNCH_CHECK(ctx);
  return 0;
}
