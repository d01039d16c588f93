// Coordinate head of the last E_GCL layer on the tensor cores (ligand rows only):
//   w = clamp(wc2 . SiLU(Wc1 m* + bc1), +-2),  f_i = mean_k (x_i - x_j) / (|x_i - x_j| + 1) w      (src/models/egnn.py:118-148)
// m* = the gated messages edge_ws.cu spills for the ligand residues, fp16 (x 2^-6), [tile = 2 residues][4 K blocks][128 rows][64].
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
#include "coord_body.cuh"

namespace ntc {

__global__ void __launch_bounds__(NT, 1) k_coord(const Params p, const __grid_constant__ CUtensorMap tmX) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  pdl_trigger();            // programmatic dependent launch, see common.cuh
  LastRing none{};
  coord_body<false>(p, tmX, none, smem, (int)blockIdx.x, (int)gridDim.x);
}

static int launch(dfm_ctx* ctx, const Params& p, int grid, cudaStream_t s) {
  static unsigned long long attr_devices = 0;
  if (dfm_once_per_device(attr_devices, ctx->device)) {
    CUDA_TRY(cudaFuncSetAttribute(k_coord, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_ALLOC));
  }
  if (grid <= 0) return 0;
  CUtensorMap tmX;
  // the spill as [tiles * 4 K blocks * 128 rows, 64 columns]: one K block of a tile is one contiguous 16 KB box
  int rc = dfm_make_tmap_f16(&tmX, p.X, (uint64_t)p.ntiles * 4 * TILE_M, 64, TILE_M);
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
  if ((a.B * L) & 1) {      // odd number of ligand residues in the batch: the second half of the last tile is never written
    for (int kb = 0; kb < 4; ++kb)
      CUDA_TRY(cudaMemsetAsync(a.mstar + (((size_t)(p.ntiles - 1) * 4 + kb) * 128 + 64) * 64, 0, 64 * 64 * sizeof(__half), s));
  }
  p.X = a.mstar; p.W0 = w.img_Wc1s; p.bias0 = w.bc1; p.wc2 = w.wc2; p.nbr = a.nbr; p.pos = a.pos; p.fbuf = a.fbuf;
  return ntc::launch(ctx, p, p.ntiles < ctx->num_sms ? p.ntiles : ctx->num_sms, s);
}
