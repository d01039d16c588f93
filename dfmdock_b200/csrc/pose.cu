// SO(3) x R^3 rigid-body pose kernels: random initial pose, Euler-Maruyama reverse step, soft-clash force.
//
// Restates: randomize_pose / modify_coords / rot_compose / get_clash_force  src/inference_base.py:311-384
//           (all-atom-centroid variants                                     src/inference.py:213-286)
//           SO3Diffuser.torch_reverse / R3Diffuser.torch_reverse            src/utils/so3_diffuser.py:344-369,
//                                                                           src/utils/r3_diffuser.py:40-55
//           axis-angle <-> quaternion <-> matrix                            src/utils/geometry.py:18-200
#include "common.cuh"

struct M3 { float m[9]; };

__device__ __forceinline__ void aa_to_quat(const float* aa, float* q) {
  const float ang = sqrtf(aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2]);
  const float half = 0.5f * ang;
  const float k = (fabsf(ang) < 1e-6f) ? (0.5f - ang * ang / 48.f) : (sinf(half) / ang);
  q[0] = cosf(half); q[1] = aa[0] * k; q[2] = aa[1] * k; q[3] = aa[2] * k;
}
__device__ __forceinline__ M3 quat_to_mat(const float* q) {
  const float r = q[0], i = q[1], j = q[2], k = q[3];
  const float s = 2.f / (r * r + i * i + j * j + k * k);
  M3 o;
  o.m[0] = 1.f - s * (j * j + k * k); o.m[1] = s * (i * j - k * r); o.m[2] = s * (i * k + j * r);
  o.m[3] = s * (i * j + k * r); o.m[4] = 1.f - s * (i * i + k * k); o.m[5] = s * (j * k - i * r);
  o.m[6] = s * (i * k - j * r); o.m[7] = s * (j * k + i * r); o.m[8] = 1.f - s * (i * i + j * j);
  return o;
}
__device__ __forceinline__ M3 aa_to_mat(const float* aa) {
  float q[4];
  aa_to_quat(aa, q);
  return quat_to_mat(q);
}
// geometry.py:64-123: four candidate quaternions, take the one with the largest |q_abs| (floor 0.1)
__device__ __forceinline__ void mat_to_quat(const M3& R, float* q) {
  const float m00 = R.m[0], m01 = R.m[1], m02 = R.m[2], m10 = R.m[3], m11 = R.m[4], m12 = R.m[5], m20 = R.m[6],
              m21 = R.m[7], m22 = R.m[8];
  float qa[4] = {1.f + m00 + m11 + m22, 1.f + m00 - m11 - m22, 1.f - m00 + m11 - m22, 1.f - m00 - m11 + m22};
#pragma unroll
  for (int c = 0; c < 4; ++c) qa[c] = qa[c] > 0.f ? sqrtf(qa[c]) : 0.f;
  int best = 0;
#pragma unroll
  for (int c = 1; c < 4; ++c) if (qa[c] > qa[best]) best = c;
  float cand[4];
  if (best == 0)      { cand[0] = qa[0] * qa[0]; cand[1] = m21 - m12; cand[2] = m02 - m20; cand[3] = m10 - m01; }
  else if (best == 1) { cand[0] = m21 - m12; cand[1] = qa[1] * qa[1]; cand[2] = m10 + m01; cand[3] = m02 + m20; }
  else if (best == 2) { cand[0] = m02 - m20; cand[1] = m10 + m01; cand[2] = qa[2] * qa[2]; cand[3] = m12 + m21; }
  else                { cand[0] = m10 - m01; cand[1] = m20 + m02; cand[2] = m21 + m12; cand[3] = qa[3] * qa[3]; }
  const float den = 2.f * fmaxf(qa[best], 0.1f);
#pragma unroll
  for (int c = 0; c < 4; ++c) q[c] = cand[c] / den;
}
__device__ __forceinline__ void quat_to_aa(const float* q, float* aa) {
  const float nrm = sqrtf(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const float half = atan2f(nrm, q[0]);
  const float ang = 2.f * half;
  const float k = (fabsf(ang) < 1e-6f) ? (0.5f - ang * ang / 48.f) : (sinf(half) / ang);
  aa[0] = q[1] / k; aa[1] = q[2] / k; aa[2] = q[3] / k;
}
__device__ __forceinline__ void mat_to_aa(const M3& R, float* aa) {
  float q[4];
  mat_to_quat(R, q);
  quat_to_aa(q, aa);
}

// block-wide (256 threads) deterministic sum of three values
__device__ __forceinline__ void block_sum3(float& x, float& y, float& z, float (*red)[8]) {
  const int tid = threadIdx.x;
  x = warp_sum(x); y = warp_sum(y); z = warp_sum(z);
  __syncthreads();
  if ((tid & 31) == 0) { red[0][tid >> 5] = x; red[1][tid >> 5] = y; red[2][tid >> 5] = z; }
  __syncthreads();
  x = y = z = 0.f;
#pragma unroll
  for (int q = 0; q < 8; ++q) { x += red[0][q]; y += red[1][q]; z += red[2][q]; }
}

// centroid of the CA atoms (mode 0) or of all N/CA/C atoms (mode 1) of x[n,3,3]
__device__ __forceinline__ void centroid(const float* x, int n, int all_atoms, float& cx, float& cy, float& cz,
                                         float (*red)[8]) {
  float sx = 0.f, sy = 0.f, sz = 0.f;
  if (all_atoms) {
    for (int a = threadIdx.x; a < n * 3; a += 256) { sx += x[a * 3]; sy += x[a * 3 + 1]; sz += x[a * 3 + 2]; }
  } else {
    for (int l = threadIdx.x; l < n; l += 256) { sx += x[l * 9 + 3]; sy += x[l * 9 + 4]; sz += x[l * 9 + 5]; }
  }
  block_sum3(sx, sy, sz, red);
  const float inv = 1.f / (float)(all_atoms ? n * 3 : n);
  cx = sx * inv; cy = sy * inv; cz = sz * inv;
}

__global__ void __launch_bounds__(256)
k_randomize_pose(int R, int L, int all_atoms, const float* __restrict__ rec_pos, const float* __restrict__ lig0,
                 const float* __restrict__ rot0, const float* __restrict__ tr0, uint64_t seed, uint64_t stream_base,
                 float* __restrict__ lig_pos, float* __restrict__ rot_update, float* __restrict__ tr_update) {
  __shared__ float red[3][8];
  const int b = blockIdx.x;
  float c1x, c1y, c1z, c2x, c2y, c2z;
  centroid(rec_pos, R, all_atoms, c1x, c1y, c1z, red);
  centroid(lig0, L, all_atoms, c2x, c2y, c2z, red);
  M3 Rm;
  float tn[3];
  if (rot0 != nullptr) {
#pragma unroll
    for (int q = 0; q < 9; ++q) Rm.m[q] = rot0[(size_t)b * 9 + q];
  } else {
    // scipy Rotation.random(): normalised 4-vector of N(0,1) draws, scalar-last
    uint4 r4a = dfm_rng(seed, stream_base + b, RNG_INIT, 0, 0, 0);
    float2 g0 = box_muller(r4a.x, r4a.y), g1 = box_muller(r4a.z, r4a.w);
    float q[4] = {g1.y, g0.x, g0.y, g1.x};   // (w, x, y, z)
    Rm = quat_to_mat(q);
  }
  if (tr0 != nullptr) {
    tn[0] = tr0[b * 3]; tn[1] = tr0[b * 3 + 1]; tn[2] = tr0[b * 3 + 2];
  } else {
    uint4 r4b = dfm_rng(seed, stream_base + b, RNG_INIT, 0, 1, 0);
    float2 g0 = box_muller(r4b.x, r4b.y), g1 = box_muller(r4b.z, r4b.w);
    tn[0] = 30.f * g0.x; tn[1] = 30.f * g0.y; tn[2] = 30.f * g1.x;
  }
  const float tx = tn[0] - c2x + c1x, ty = tn[1] - c2y + c1y, tz = tn[2] - c2z + c1z;
  float* out = lig_pos + (size_t)b * L * 9;
  for (int a = threadIdx.x; a < L * 3; a += 256) {
    const float x = lig0[a * 3] - c2x, y = lig0[a * 3 + 1] - c2y, z = lig0[a * 3 + 2] - c2z;
    out[a * 3 + 0] = Rm.m[0] * x + Rm.m[1] * y + Rm.m[2] * z + c2x + tx;
    out[a * 3 + 1] = Rm.m[3] * x + Rm.m[4] * y + Rm.m[5] * z + c2y + ty;
    out[a * 3 + 2] = Rm.m[6] * x + Rm.m[7] * y + Rm.m[8] * z + c2z + tz;
  }
  if (threadIdx.x == 0) {
    float aa[3];
    mat_to_aa(Rm, aa);
    rot_update[b * 3] = aa[0]; rot_update[b * 3 + 1] = aa[1]; rot_update[b * 3 + 2] = aa[2];
    tr_update[b * 3] = tx; tr_update[b * 3 + 1] = ty; tr_update[b * 3 + 2] = tz;
  }
}

__global__ void __launch_bounds__(256)
k_reverse_step(int L, int all_atoms, int ode, float* __restrict__ lig_pos, float* __restrict__ rot_update,
               float* __restrict__ tr_update, const float* __restrict__ tr_score, const float* __restrict__ rot_score,
               float g_rot, float g_tr, float dt, float ns_rot, float ns_tr, const float* __restrict__ z,
               uint64_t seed, uint64_t stream_base, uint32_t step) {
  __shared__ float red[3][8];
  const int b = blockIdx.x;
  float* x = lig_pos + (size_t)b * L * 9;
  float zr[3] = {0.f, 0.f, 0.f}, zt[3] = {0.f, 0.f, 0.f};
  if (!ode) {
    if (z != nullptr) {
#pragma unroll
      for (int q = 0; q < 3; ++q) { zr[q] = z[(size_t)b * 6 + q]; zt[q] = z[(size_t)b * 6 + 3 + q]; }
    } else {
      uint4 ra = dfm_rng(seed, stream_base + b, RNG_STEP, step, 0, 0);
      uint4 rb = dfm_rng(seed, stream_base + b, RNG_STEP, step, 1, 0);
      float2 a0 = box_muller(ra.x, ra.y), a1 = box_muller(ra.z, ra.w);
      float2 b0 = box_muller(rb.x, rb.y), b1 = box_muller(rb.z, rb.w);
      zr[0] = a0.x; zr[1] = a0.y; zr[2] = a1.x;
      zt[0] = b0.x; zt[1] = b0.y; zt[2] = b1.x;
    }
  }
  float rot[3], tr[3];
  const float sq = sqrtf(dt);
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    if (ode) {
      rot[q] = 0.5f * (g_rot * g_rot) * rot_score[b * 3 + q] * dt;
      tr[q] = 0.5f * (g_tr * g_tr) * tr_score[b * 3 + q] * dt;
    } else {
      rot[q] = (g_rot * g_rot) * rot_score[b * 3 + q] * dt + g_rot * sq * (ns_rot * zr[q]);
      tr[q] = (g_tr * g_tr) * tr_score[b * 3 + q] * dt + g_tr * sq * (ns_tr * zt[q]);
    }
  }
  float cx, cy, cz;
  centroid(x, L, all_atoms, cx, cy, cz, red);
  const M3 Rm = aa_to_mat(rot);
  for (int a = threadIdx.x; a < L * 3; a += 256) {
    const float px = x[a * 3] - cx, py = x[a * 3 + 1] - cy, pz = x[a * 3 + 2] - cz;
    x[a * 3 + 0] = Rm.m[0] * px + Rm.m[1] * py + Rm.m[2] * pz + cx + tr[0];
    x[a * 3 + 1] = Rm.m[3] * px + Rm.m[4] * py + Rm.m[5] * pz + cy + tr[1];
    x[a * 3 + 2] = Rm.m[6] * px + Rm.m[7] * py + Rm.m[8] * pz + cz + tr[2];
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < 3; ++q) tr_update[b * 3 + q] += tr[q];
    // rot_compose(rot_update, rot): R = R(rot) * R(rot_update)
    const M3 R1 = aa_to_mat(rot_update + b * 3);
    M3 Rc;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int k = 0; k < 3; ++k)
        Rc.m[i * 3 + k] = Rm.m[i * 3] * R1.m[k] + Rm.m[i * 3 + 1] * R1.m[3 + k] + Rm.m[i * 3 + 2] * R1.m[6 + k];
    float aa[3];
    mat_to_aa(Rc, aa);
    rot_update[b * 3] = aa[0]; rot_update[b * 3 + 1] = aa[1]; rot_update[b * 3 + 2] = aa[2];
  }
}

// U = -5 sum_{d<4} (4-d)^1.5 / (0.75 d) over all backbone atom pairs; F = mean over ligand atoms of dU/dx.
// Analytic gradient of the reference's autograd formulation.
// Residue-pair pruning: the atoms of a residue lie within rho = max(|N-CA|, |C-CA|) of its CA, so a residue pair whose
// CA-CA distance exceeds 4 + rho_l + rho_r has no atom pair inside the 4 A cut-off -- exact, not an approximation.
// Work split: one CTA per (trajectory, tile of 64 receptor residues, slice of 32 ligand residues); a warp takes four ligand
// residues of the slice, its lanes the receptor residues of the tile.  The partial forces of a trajectory's CTAs are summed in
// a fixed order by k_clash_apply (bit-reproducible), which also moves the ligand.  (One CTA per trajectory, the first form of
// this kernel, left 108 of 148 SMs idle at 40 trajectories and took 74 us of a 920 us step at BASELINE config #2.)
constexpr int CLASH_RT = 64, CLASH_LS = 32;
__global__ void __launch_bounds__(256)
k_clash_partial(int R, int L, int nrt, const float* __restrict__ rec_pos, const float* __restrict__ lig_pos,
                float* __restrict__ partial) {
  __shared__ float red[3][8];
  __shared__ float rs[CLASH_RT * 10];       // per receptor residue: N, CA, C coordinates + rho
  const int b = blockIdx.y, part = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rt = part % nrt, ls = part / nrt;
  const int r0 = rt * CLASH_RT, cnt = min(CLASH_RT, R - r0);
  const float* x = lig_pos + (size_t)b * L * 9;
  for (int i = tid; i < cnt * 9; i += 256) rs[(i / 9) * 10 + (i % 9)] = rec_pos[(size_t)r0 * 9 + i];
  __syncthreads();
  if (tid < cnt) {
    const float* q = rs + tid * 10;
    const float ax = q[0] - q[3], ay = q[1] - q[4], az = q[2] - q[5];
    const float cx = q[6] - q[3], cy = q[7] - q[4], cz = q[8] - q[5];
    rs[tid * 10 + 9] = sqrtf(fmaxf(ax * ax + ay * ay + az * az, cx * cx + cy * cy + cz * cz));
  }
  __syncthreads();
  float fx = 0.f, fy = 0.f, fz = 0.f;
  for (int li = warp; li < CLASH_LS; li += 8) {
    const int l = ls * CLASH_LS + li;
    if (l >= L) break;
    float p[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) p[k] = x[l * 9 + k];
    const float ax = p[0] - p[3], ay = p[1] - p[4], az = p[2] - p[5];
    const float cx = p[6] - p[3], cy = p[7] - p[4], cz = p[8] - p[5];
    const float reach = 4.f + sqrtf(fmaxf(ax * ax + ay * ay + az * az, cx * cx + cy * cy + cz * cz)) + 1e-3f;
    for (int r = lane; r < cnt; r += 32) {
      const float* q = rs + r * 10;
      const float ex = p[3] - q[3], ey = p[4] - q[4], ez = p[5] - q[5];
      const float lim = reach + q[9];
      if (ex * ex + ey * ey + ez * ez > lim * lim) continue;
#pragma unroll
      for (int la = 0; la < 3; ++la) {
#pragma unroll
        for (int ra = 0; ra < 3; ++ra) {
          const float dx = p[la * 3] - q[ra * 3], dy = p[la * 3 + 1] - q[ra * 3 + 1], dz = p[la * 3 + 2] - q[ra * 3 + 2];
          const float d2 = dx * dx + dy * dy + dz * dz;
          if (d2 < 16.f) {
            const float d = sqrtf(d2);
            const float g = 4.f - d;
            const float sg = sqrtf(g);
            const float dphi = -(1.5f * sg * d + g * sg) / (0.75f * d * d);
            const float coef = -5.f * dphi / d;
            fx = fmaf(coef, dx, fx); fy = fmaf(coef, dy, fy); fz = fmaf(coef, dz, fz);
          }
        }
      }
    }
  }
  block_sum3(fx, fy, fz, red);
  if (tid == 0) {
    float* o = partial + ((size_t)b * gridDim.x + part) * 4;
    o[0] = fx; o[1] = fy; o[2] = fz; o[3] = 0.f;
  }
}
__global__ void __launch_bounds__(256)
k_clash_apply(int L, int nparts, const float* __restrict__ partial, float* __restrict__ lig_pos, float* __restrict__ tr_update) {
  __shared__ float red[3][8];
  const int b = blockIdx.x, tid = threadIdx.x;
  float fx = 0.f, fy = 0.f, fz = 0.f;
  for (int q = tid; q < nparts; q += 256) {      // fixed assignment and order: the same bits every run
    const float* o = partial + ((size_t)b * nparts + q) * 4;
    fx += o[0]; fy += o[1]; fz += o[2];
  }
  block_sum3(fx, fy, fz, red);
  const float inv = 1.f / (float)(L * 3);
  fx *= inv; fy *= inv; fz *= inv;
  float* x = lig_pos + (size_t)b * L * 9;
  for (int l = tid; l < L * 3; l += 256) { x[l * 3] += fx; x[l * 3 + 1] += fy; x[l * 3 + 2] += fz; }
  if (tid == 0) { tr_update[b * 3] += fx; tr_update[b * 3 + 1] += fy; tr_update[b * 3 + 2] += fz; }
}

// floats of clash-force scratch a complex of R + L residues needs (dfm_set_complex allocates it, grow-only)
size_t clash_scratch_floats(int R, int L) {
  return (size_t)CLASH_BATCH * ((R + CLASH_RT - 1) / CLASH_RT) * ((L + CLASH_LS - 1) / CLASH_LS) * 4;
}

int launch_randomize_pose(dfm_ctx* ctx, int B, const float* lig0, const float* rot0, const float* tr0, uint64_t seed,
                          uint64_t stream_base, uint32_t flags, float* lig_pos, float* rot_update, float* tr_update,
                          cudaStream_t s) {
  k_randomize_pose<<<B, 256, 0, s>>>(ctx->R, ctx->L, (flags & DFM_CENTRE_ALL_ATOMS) ? 1 : 0, ctx->rec_pos, lig0, rot0,
                                     tr0, seed, stream_base, lig_pos, rot_update, tr_update);
  LAUNCH_CHECK(ctx);
  return 0;
}

int launch_reverse_step(dfm_ctx* ctx, int B, float* lig_pos, float* rot_update, float* tr_update, const float* tr_score,
                        const float* rot_score, float g_rot, float g_tr, float dt, float ns_rot, float ns_tr,
                        const float* z, uint64_t seed, uint64_t stream_base, uint32_t step, uint32_t flags,
                        cudaStream_t s) {
  k_reverse_step<<<B, 256, 0, s>>>(ctx->L, (flags & DFM_CENTRE_ALL_ATOMS) ? 1 : 0, (flags & DFM_ODE) ? 1 : 0, lig_pos,
                                   rot_update, tr_update, tr_score, rot_score, g_rot, g_tr, dt, ns_rot, ns_tr, z, seed,
                                   stream_base, step);
  LAUNCH_CHECK(ctx);
  if (flags & DFM_CLASH_FORCE) {
    const int nrt = (ctx->R + CLASH_RT - 1) / CLASH_RT, nls = (ctx->L + CLASH_LS - 1) / CLASH_LS;
    const int nparts = nrt * nls;
    // scratch of CLASH_BATCH trajectories' partial forces, allocated by dfm_set_complex (nothing is allocated here)
    for (int b0 = 0; b0 < B; b0 += CLASH_BATCH) {
      const int nb = min(CLASH_BATCH, B - b0);
      k_clash_partial<<<dim3(nparts, nb), 256, 0, s>>>(ctx->R, ctx->L, nrt, ctx->rec_pos, lig_pos + (size_t)b0 * ctx->L * 9,
                                                     ctx->clash_partial);
      LAUNCH_CHECK(ctx);
      k_clash_apply<<<nb, 256, 0, s>>>(ctx->L, nparts, ctx->clash_partial, lig_pos + (size_t)b0 * ctx->L * 9, tr_update + (size_t)b0 * 3);
      LAUNCH_CHECK(ctx);
    }
  }
  return 0;
}

// ---- all-atom rigid transform of the docked ligand -----------------------------------------
// Restates modify_aa_coords: x <- (x - c) R^T + c + tr with R = axis_angle_to_matrix(rot_update);
//   centre_mode 0: c = centroid of the ligand's backbone CA atoms      (src/inference_base.py:354-364)
//   centre_mode 1: c = centroid of the all-atom coordinates themselves (src/inference.py:256-266)
// One CTA per pose; atoms [A,3] are shared by all poses, out is [T,A,3].
__global__ void __launch_bounds__(256)
k_transform_atoms(int A, int L, int centre_mode, const float* __restrict__ atoms, const float* __restrict__ lig_bb,
                  const float* __restrict__ rot_update, const float* __restrict__ tr_update, float* __restrict__ out) {
  __shared__ float red[3][8];
  const int b = blockIdx.x;
  float cx, cy, cz;
  if (centre_mode == 0) {
    centroid(lig_bb, L, 0, cx, cy, cz, red);
  } else {
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (int a = threadIdx.x; a < A; a += 256) { sx += atoms[a * 3]; sy += atoms[a * 3 + 1]; sz += atoms[a * 3 + 2]; }
    block_sum3(sx, sy, sz, red);
    const float inv = 1.f / (float)A;
    cx = sx * inv; cy = sy * inv; cz = sz * inv;
  }
  const float aa[3] = {rot_update[b * 3], rot_update[b * 3 + 1], rot_update[b * 3 + 2]};
  const M3 Rm = aa_to_mat(aa);
  const float tx = tr_update[b * 3] + cx, ty = tr_update[b * 3 + 1] + cy, tz = tr_update[b * 3 + 2] + cz;
  float* o = out + (size_t)b * A * 3;
  for (int a = threadIdx.x; a < A; a += 256) {
    const float x = atoms[a * 3] - cx, y = atoms[a * 3 + 1] - cy, z = atoms[a * 3 + 2] - cz;
    o[a * 3 + 0] = Rm.m[0] * x + Rm.m[1] * y + Rm.m[2] * z + tx;
    o[a * 3 + 1] = Rm.m[3] * x + Rm.m[4] * y + Rm.m[5] * z + ty;
    o[a * 3 + 2] = Rm.m[6] * x + Rm.m[7] * y + Rm.m[8] * z + tz;
  }
}

extern "C" int dfm_transform_atoms(int device, int T, int A, int L, int centre_mode, const float* atoms,
                                   const float* lig_bb, const float* rot_update, const float* tr_update, float* out,
                                   void* stream) {
  if (T <= 0 || A <= 0 || !atoms || !rot_update || !tr_update || !out || (centre_mode == 0 && (L <= 0 || !lig_bb)) ||
      (centre_mode != 0 && centre_mode != 1)) {
    dfm_set_error("dfm_transform_atoms: bad argument");
    return DFM_EINVAL;
  }
  CUDA_TRY(cudaSetDevice(device));
  k_transform_atoms<<<T, 256, 0, (cudaStream_t)stream>>>(A, L, centre_mode, atoms, lig_bb, rot_update, tr_update, out);
  if (cudaGetLastError() != cudaSuccess) { dfm_set_error("dfm_transform_atoms: launch failed"); return DFM_ECUDA; }
  return DFM_OK;
}
