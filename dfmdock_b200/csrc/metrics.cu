// Docking quality metrics of T docked poses of one complex, batched on the GPU.
//
// Restates: compute_metrics / get_c_rmsd / get_i_rmsd / get_l_rmsd / get_fnat / get_DockQ / get_interface_res /
//           get_dist / find_rigid_alignment                                  src/utils/metrics.py:3-121
// All point sets are backbone atoms (N, CA, C) of whole residues, like the reference ([n,3,3] flattened to [3n,3]).
// The optimal proper rotation (the reference's Kabsch SVD with its reflection fix) is obtained from Horn's 4x4
// quaternion matrix by Jacobi iteration in fp64; RMSDs are then evaluated on the transformed points (no cancellation).
#include <math.h>

#include "common.cuh"

namespace {

constexpr float IFACE_CUT = 10.0f;   // get_interface_res cutoff (metrics.py:18, :43)
constexpr float FNAT_CUT = 5.5f;     // get_fnat cutoff (metrics.py:59)

__device__ __forceinline__ float min_res_dist(const float* a, const float* b) {
  // minimum over the 3 x 3 backbone-atom pairs of two residues (metrics.py:77-85)
  float m = 3.0e38f;
#pragma unroll
  for (int p = 0; p < 3; ++p)
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const float dx = a[p * 3] - b[q * 3], dy = a[p * 3 + 1] - b[q * 3 + 1], dz = a[p * 3 + 2] - b[q * 3 + 2];
      m = fminf(m, sqrtf(dx * dx + dy * dy + dz * dz));
    }
  return m;
}

// per complex: native contacts (< 5.5 A) as one byte per residue pair, interface residues (< 10 A) as flags
__global__ void k_metrics_native(int R, int L, const float* __restrict__ nrec, const float* __restrict__ nlig,
                                 uint8_t* __restrict__ contact, int* __restrict__ rec_if, int* __restrict__ lig_if) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * L) return;
  const int r = idx / L, l = idx % L;
  const float d = min_res_dist(nrec + (size_t)r * 9, nlig + (size_t)l * 9);
  contact[idx] = d < FNAT_CUT ? 1 : 0;
  if (d < IFACE_CUT) { rec_if[r] = 1; lig_if[l] = 1; }
}

__device__ double block_sum_d(double v, double* red) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
  return s;
}

// largest-eigenvalue eigenvector of the symmetric 4x4 matrix K (cyclic Jacobi, fp64) -> unit quaternion (w, x, y, z)
__device__ void jacobi4_max(double K[4][4], double q[4]) {
  double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0;
    for (int i = 0; i < 4; ++i)
      for (int j = i + 1; j < 4; ++j) off += K[i][j] * K[i][j];
    if (off < 1e-30) break;
    for (int p = 0; p < 4; ++p)
      for (int r = p + 1; r < 4; ++r) {
        if (fabs(K[p][r]) < 1e-300) continue;
        const double theta = (K[r][r] - K[p][p]) / (2.0 * K[p][r]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 4; ++k) {
          const double kp = K[k][p], kr = K[k][r];
          K[k][p] = c * kp - s * kr;
          K[k][r] = s * kp + c * kr;
        }
        for (int k = 0; k < 4; ++k) {
          const double kp = K[p][k], kr = K[r][k];
          K[p][k] = c * kp - s * kr;
          K[r][k] = s * kp + c * kr;
        }
        for (int k = 0; k < 4; ++k) {
          const double vp = V[k][p], vr = V[k][r];
          V[k][p] = c * vp - s * vr;
          V[k][r] = s * vp + c * vr;
        }
      }
  }
  int best = 0;
  for (int i = 1; i < 4; ++i) if (K[i][i] > K[best][best]) best = i;
  double n = 0.0;
  for (int k = 0; k < 4; ++k) { q[k] = V[k][best]; n += q[k] * q[k]; }
  n = 1.0 / sqrt(n);
  for (int k = 0; k < 4; ++k) q[k] *= n;
}

// Rigid alignment of the selected source atoms onto the target atoms (find_rigid_alignment, metrics.py:91-121).
// Atom a of the virtual concatenation [rec (3R atoms); lig (3L atoms)]; sel(a) says whether it takes part.
// Returns R (row-major) and t in shared memory `xf[12]` (valid for all threads after the call).
template <typename Sel>
__device__ void align(int R, int L, const float* srec, const float* slig, const float* trec, const float* tlig, Sel sel,
                      double* red, double* xf, int* cnt_out) {
  const int na = 3 * (R + L);
  auto src = [&](int a) { return a < 3 * R ? srec + (size_t)a * 3 : slig + (size_t)(a - 3 * R) * 3; };
  auto tgt = [&](int a) { return a < 3 * R ? trec + (size_t)a * 3 : tlig + (size_t)(a - 3 * R) * 3; };
  double s[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int a = threadIdx.x; a < na; a += blockDim.x) {
    if (!sel(a)) continue;
    const float* p = src(a); const float* t = tgt(a);
    s[0] += p[0]; s[1] += p[1]; s[2] += p[2]; s[3] += t[0]; s[4] += t[1]; s[5] += t[2]; s[6] += 1.0;
  }
  double m[7];
  for (int k = 0; k < 7; ++k) m[k] = block_sum_d(s[k], red);
  const double n = m[6];
  const double inv = n > 0 ? 1.0 / n : 0.0;
  const double ax = m[0] * inv, ay = m[1] * inv, az = m[2] * inv, bx = m[3] * inv, by = m[4] * inv, bz = m[5] * inv;
  double h[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int a = threadIdx.x; a < na; a += blockDim.x) {
    if (!sel(a)) continue;
    const float* p = src(a); const float* t = tgt(a);
    const double px = p[0] - ax, py = p[1] - ay, pz = p[2] - az, tx = t[0] - bx, ty = t[1] - by, tz = t[2] - bz;
    h[0] += px * tx; h[1] += px * ty; h[2] += px * tz;
    h[3] += py * tx; h[4] += py * ty; h[5] += py * tz;
    h[6] += pz * tx; h[7] += pz * ty; h[8] += pz * tz;
  }
  double S[9];
  for (int k = 0; k < 9; ++k) S[k] = block_sum_d(h[k], red);
  __syncthreads();
  if (threadIdx.x == 0) {
    const double Sxx = S[0], Sxy = S[1], Sxz = S[2], Syx = S[3], Syy = S[4], Syz = S[5], Szx = S[6], Szy = S[7], Szz = S[8];
    double K[4][4] = {{Sxx + Syy + Szz, Syz - Szy, Szx - Sxz, Sxy - Syx},
                      {Syz - Szy, Sxx - Syy - Szz, Sxy + Syx, Szx + Sxz},
                      {Szx - Sxz, Sxy + Syx, -Sxx + Syy - Szz, Syz + Szy},
                      {Sxy - Syx, Szx + Sxz, Syz + Szy, -Sxx - Syy + Szz}};
    double q[4];
    jacobi4_max(K, q);
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    double Rm[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                    2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                    2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)};
    for (int k = 0; k < 9; ++k) xf[k] = Rm[k];
    xf[9] = bx - (Rm[0] * ax + Rm[1] * ay + Rm[2] * az);
    xf[10] = by - (Rm[3] * ax + Rm[4] * ay + Rm[5] * az);
    xf[11] = bz - (Rm[6] * ax + Rm[7] * ay + Rm[8] * az);
    *cnt_out = (int)(n + 0.5);
  }
  __syncthreads();
}

// sum over selected atoms of |R p + t - target|^2 (and the count)
template <typename Sel>
__device__ double sq_dev(int R, int L, const float* srec, const float* slig, const float* trec, const float* tlig, Sel sel,
                         const double* xf, double* red) {
  const int na = 3 * (R + L);
  double acc = 0.0;
  for (int a = threadIdx.x; a < na; a += blockDim.x) {
    if (!sel(a)) continue;
    const float* p = a < 3 * R ? srec + (size_t)a * 3 : slig + (size_t)(a - 3 * R) * 3;
    const float* t = a < 3 * R ? trec + (size_t)a * 3 : tlig + (size_t)(a - 3 * R) * 3;
    const double x = xf[0] * p[0] + xf[1] * p[1] + xf[2] * p[2] + xf[9] - t[0];
    const double y = xf[3] * p[0] + xf[4] * p[1] + xf[5] * p[2] + xf[10] - t[1];
    const double z = xf[6] * p[0] + xf[7] * p[1] + xf[8] * p[2] + xf[11] - t[2];
    acc += x * x + y * y + z * z;
  }
  return block_sum_d(acc, red);
}

__global__ void __launch_bounds__(256)
k_metrics_pose(int R, int L, const float* __restrict__ mrec, long rec_stride, const float* __restrict__ mlig,
               const float* __restrict__ nrec, const float* __restrict__ nlig, const uint8_t* __restrict__ contact,
               const int* __restrict__ rec_if, const int* __restrict__ lig_if, float* __restrict__ out) {
  __shared__ double red[8];
  __shared__ double xf[12];
  __shared__ int cnt;
  const int b = blockIdx.x;
  const float* srec = mrec + (size_t)b * rec_stride;
  const float* slig = mlig + (size_t)b * L * 9;
  // c_rmsd: all atoms (metrics.py:33-38)
  auto all = [&](int) { return true; };
  align(R, L, srec, slig, nrec, nlig, all, red, xf, &cnt);
  const double c_sq = sq_dev(R, L, srec, slig, nrec, nlig, all, xf, red);
  const double c_rmsd = sqrt(c_sq / (double)(3 * (R + L)));
  // i_rmsd: atoms of the native interface residues (metrics.py:40-46)
  auto iface = [&](int a) { return a < 3 * R ? rec_if[a / 3] != 0 : lig_if[(a - 3 * R) / 3] != 0; };
  align(R, L, srec, slig, nrec, nlig, iface, red, xf, &cnt);
  const int n_if = cnt;
  const double i_sq = sq_dev(R, L, srec, slig, nrec, nlig, iface, xf, red);
  const double i_rmsd = n_if > 0 ? sqrt(i_sq / (double)n_if) : nan("");
  // l_rmsd: superimpose the receptors, measure the ligand (metrics.py:48-56)
  auto recs = [&](int a) { return a < 3 * R; };
  auto ligs = [&](int a) { return a >= 3 * R; };
  align(R, L, srec, slig, nrec, nlig, recs, red, xf, &cnt);
  const double l_sq = sq_dev(R, L, srec, slig, nrec, nlig, ligs, xf, red);
  const double l_rmsd = sqrt(l_sq / (double)(3 * L));
  // fnat: native contacts kept in the model (metrics.py:58-69)
  double kept = 0.0, total = 0.0;
  for (int idx = threadIdx.x; idx < R * L; idx += blockDim.x) {
    if (!contact[idx]) continue;
    total += 1.0;
    const int r = idx / L, l = idx % L;
    if (min_res_dist(srec + (size_t)r * 9, slig + (size_t)l * 9) < FNAT_CUT) kept += 1.0;
  }
  kept = block_sum_d(kept, red);
  total = block_sum_d(total, red);
  if (threadIdx.x == 0) {
    const double fnat = round(kept / (total + 1e-6) * 1e6) / 1e6;
    const double is = 1.0 / (1.0 + (i_rmsd / 1.5) * (i_rmsd / 1.5));
    const double ls = 1.0 / (1.0 + (l_rmsd / 8.5) * (l_rmsd / 8.5));
    float* o = out + (size_t)b * 5;
    o[0] = (float)c_rmsd; o[1] = (float)i_rmsd; o[2] = (float)l_rmsd; o[3] = (float)fnat;
    o[4] = (float)((fnat + is + ls) / 3.0);
  }
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

extern "C" size_t dfm_metrics_workspace_bytes(int R, int L) {
  if (R <= 0 || L <= 0) return 0;
  return align256((size_t)R * L) + align256((size_t)R * 4) + align256((size_t)L * 4);
}

extern "C" int dfm_compute_metrics(int device, int T, int R, int L, const float* model_rec, int rec_is_shared,
                                   const float* model_lig, const float* native_rec, const float* native_lig, float* out,
                                   void* workspace, size_t workspace_bytes, void* stream) {
  if (T <= 0 || R <= 0 || L <= 0 || !model_rec || !model_lig || !native_rec || !native_lig || !out || !workspace) {
    dfm_set_error("dfm_compute_metrics: bad argument");
    return DFM_EINVAL;
  }
  if (workspace_bytes < dfm_metrics_workspace_bytes(R, L)) {
    dfm_set_error("dfm_compute_metrics: workspace too small: %zu < %zu bytes", workspace_bytes, dfm_metrics_workspace_bytes(R, L));
    return DFM_ENOMEM;
  }
  cudaStream_t s = (cudaStream_t)stream;
  CUDA_TRY(cudaSetDevice(device));
  uint8_t* base = reinterpret_cast<uint8_t*>(workspace);
  uint8_t* contact = base;
  int* rec_if = reinterpret_cast<int*>(base + align256((size_t)R * L));
  int* lig_if = reinterpret_cast<int*>(base + align256((size_t)R * L) + align256((size_t)R * 4));
  CUDA_TRY(cudaMemsetAsync(rec_if, 0, align256((size_t)R * 4) + align256((size_t)L * 4), s));
  k_metrics_native<<<(R * L + 255) / 256, 256, 0, s>>>(R, L, native_rec, native_lig, contact, rec_if, lig_if);
  if (cudaGetLastError() != cudaSuccess) { dfm_set_error("dfm_compute_metrics: launch failed"); return DFM_ECUDA; }
  k_metrics_pose<<<T, 256, 0, s>>>(R, L, model_rec, rec_is_shared ? 0L : (long)R * 9, model_lig, native_rec, native_lig,
                                   contact, rec_if, lig_if, out);
  if (cudaGetLastError() != cudaSuccess) { dfm_set_error("dfm_compute_metrics: launch failed"); return DFM_ECUDA; }
  return DFM_OK;
}
