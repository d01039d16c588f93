// C ABI (include/dfmdock_b200.h): context, weight repacking, the forward pass schedule and the sampler loop.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"

int dfm_upload_bin_edges();
int launch_randomize_pose(dfm_ctx*, int, const float*, const float*, const float*, uint64_t, uint64_t, uint32_t, float*,
                          float*, float*, cudaStream_t);
int launch_reverse_step(dfm_ctx*, int, float*, float*, float*, const float*, const float*, float, float, float, float,
                        float, const float*, uint64_t, uint64_t, uint32_t, uint32_t, cudaStream_t);

static thread_local char g_err[512] = "";
void dfm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* dfm_last_error(void) { return g_err; }
// Programmatic dependent launch pays where a step is latency-bound and costs a little where every kernel fills the GPU
// for long (measured on B200 at 2x150 residues, profiles/r02/pdl_by_batch.txt: 32 trajectories -6.5 %, 64 -1.7 %, 128 +0.7 %,
// 256 +1.4 % step time with the attribute set), so by default it follows the size of the batch in flight (rows = B * N of
// the last forward on this thread).  DFM_PDL=0 / 1 in the environment forces it off / on.
static thread_local long long g_pdl_rows = 0;
void dfm_pdl_set_rows(long long rows) { g_pdl_rows = rows; }
bool dfm_pdl_enabled() {
  static int mode = -2;                  // -1: by batch size, 0 / 1: forced
  if (mode == -2) { const char* e = getenv("DFM_PDL"); mode = e ? (atoi(e) != 0) : -1; }
  if (mode >= 0) return mode != 0;
  return g_pdl_rows <= DFM_PDL_MAX_ROWS;
}
extern "C" const char* dfm_version(void) { return "dfmdock_b200 0.1 (sm_100a)"; }

// ---------------------------------------------------------------------------------------------------
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

Workspace carve_workspace(const dfm_ctx* ctx, int B, void* base) {
  Workspace w{};
  const size_t N = ctx->N, L = ctx->L, R = ctx->R, b = B;
  size_t off = 0;
  uint8_t* p = reinterpret_cast<uint8_t*>(base);
  auto take = [&](size_t bytes) {
    uint8_t* r = p ? p + off : nullptr;
    off += align256(bytes);
    return r;
  };
  w.centre = (float*)take(b * 4 * 4);
  w.pos = (float*)take(b * N * 9 * 4);
  w.cb = (float*)take(b * N * 4 * 4);
  w.nbr = (int32_t*)take(b * N * SLOTS * 4);
  w.feat = (uint32_t*)take(b * N * SLOTS * 4);
  w.radial = (float*)take(b * N * SLOTS * 4);
  w.h = (float*)take(b * N * H * 4);
  w.A = (float*)take(b * N * H * 4);
  w.Bm = (float*)take(b * N * H * 4);
  w.agg = (float*)take(b * N * H * 4);
  w.z = (float*)take(b * N * H * 4);
  w.y = (float*)take(b * N * H * 4);
  w.gstat = (float*)take(b * 2 * H * 4);
  w.mstar = (__half*)take(((b * L + 1) & ~(size_t)1) * SLOTS * H * 2);   // whole 128-row tiles (two ligand residues each)
  w.fbuf = (float*)take(b * L * 4 * 4);
  w.esum = (float*)take(b * R * 4 * 4);
  w.tsc = (float*)take(b * 8 * 4);
  w.emeta = (int4*)take(b * N * SLOTS * 16);
  w.h16 = (__half*)take(b * N * H * 2);
  w.agg16 = (__half*)take(b * N * H * 2);
  w.gscale = (float*)take(b * H * 4);
  w.gshift = (float*)take(b * H * 4);
  w.ring_flags = (unsigned int*)take(RING_FLAG_WORDS * 4);
  w.bytes = off;
  return w;
}

extern "C" size_t dfm_workspace_bytes(const dfm_ctx* ctx, int B) {
  if (!ctx || !ctx->has_complex || B <= 0) return 0;
  return carve_workspace(ctx, B, nullptr).bytes;
}
extern "C" int dfm_edges_per_node(const dfm_ctx* ctx) { return ctx ? ctx->K : 0; }
extern "C" uint64_t dfm_launch_count(const dfm_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---------------------------------------------------------------------------------------------------
extern "C" int dfm_create(dfm_ctx** out, int device) {
  if (!out) { dfm_set_error("dfm_create: null out"); return DFM_EINVAL; }
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    dfm_set_error("dfm_create: no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
    return DFM_ECUDA;
  }
  if (device < 0 || device >= count) { dfm_set_error("dfm_create: bad device %d", device); return DFM_EINVAL; }
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    dfm_set_error("dfm_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    return DFM_ECUDA;
  }
  dfm_ctx* c = new dfm_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  if (dfm_upload_bin_edges() != 0) {
    delete c;
    dfm_set_error("dfm_create: cannot upload bin edges");
    return DFM_ECUDA;
  }
  *out = c;
  return DFM_OK;
}

extern "C" void dfm_destroy(dfm_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (auto& kv : ctx->w) cudaFree(kv.second.d);
  for (void* p : ctx->owned) cudaFree(p);
  for (cudaEvent_t e : ctx->prof_events) cudaEventDestroy(e);
  cudaFree(ctx->h0);
  cudaFree(ctx->rec_pos);
  cudaFree(ctx->clash_partial);
  delete ctx;
}

extern "C" int dfm_set_weight(dfm_ctx* ctx, const char* name, const float* data, const int64_t* shape, int ndim) {
  if (!ctx || !name || !data || ndim < 0 || ndim > 4) { dfm_set_error("dfm_set_weight: bad argument"); return DFM_EINVAL; }
  CUDA_TRY(cudaSetDevice(ctx->device));
  WTensor t;
  t.numel = 1;
  for (int i = 0; i < ndim; ++i) { t.shape.push_back(shape[i]); t.numel *= shape[i]; }
  CUDA_TRY(cudaMalloc(&t.d, sizeof(float) * (size_t)(t.numel > 0 ? t.numel : 1)));
  CUDA_TRY(cudaMemcpy(t.d, data, sizeof(float) * (size_t)t.numel, cudaMemcpyDeviceToDevice));
  auto it = ctx->w.find(name);
  if (it != ctx->w.end()) { cudaFree(it->second.d); ctx->w.erase(it); }
  ctx->w[name] = t;
  ctx->finalized = false;
  return DFM_OK;
}

static int need(dfm_ctx* ctx, const std::string& name, std::vector<int64_t> shape, const float** out) {
  auto it = ctx->w.find(name);
  if (it == ctx->w.end()) { dfm_set_error("missing weight '%s'", name.c_str()); return DFM_EMISSING; }
  if (it->second.shape != shape) {
    std::string got, want;
    for (auto v : it->second.shape) got += std::to_string(v) + ",";
    for (auto v : shape) want += std::to_string(v) + ",";
    dfm_set_error("weight '%s' has shape [%s], expected [%s]", name.c_str(), got.c_str(), want.c_str());
    return DFM_EINVAL;
  }
  *out = it->second.d;
  return 0;
}

template <typename T>
static int dev_alloc(dfm_ctx* ctx, T** p, size_t n) {
  CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T)));
  ctx->owned.push_back(*p);
  return 0;
}

extern "C" int dfm_finalize_weights(dfm_ctx* ctx, float cut_off, void* stream) {
  if (!ctx) { dfm_set_error("null ctx"); return DFM_EINVAL; }
  cudaStream_t s = (cudaStream_t)stream;
  CUDA_TRY(cudaSetDevice(ctx->device));
  int rc;
#define NEED(name, ...) if ((rc = need(ctx, name, __VA_ARGS__, &tmp)) != 0) return rc
  const float* tmp = nullptr;
  auto pe = ctx->w.find("positional_embed.weight");
  if (pe == ctx->w.end() || pe->second.shape.size() != 2) { dfm_set_error("missing weight 'positional_embed.weight'"); return DFM_EMISSING; }
  ctx->P = (int)pe->second.shape[1];
  if (ctx->P != 66 && ctx->P != 67) { dfm_set_error("positional_embed width %d not supported (66 or 67)", ctx->P); return DFM_EINVAL; }
  auto se = ctx->w.find("single_embed.weight");
  if (se == ctx->w.end() || se->second.shape.size() != 2 || se->second.shape[0] != H) { dfm_set_error("missing/invalid 'single_embed.weight'"); return DFM_EMISSING; }
  ctx->x_dim = (int)se->second.shape[1];
  NEED("spatial_embed.weight", {ED, NSPATIAL});
  NEED("positional_embed.weight", {ED, ctx->P});
  for (void* p : ctx->owned) cudaFree(p);
  ctx->owned.clear();
  for (int l = 0; l < DFM_DEPTH; ++l) {
    LayerW& w = ctx->layer[l];
    const std::string pre = "network.EGNN_" + std::to_string(l) + ".egcl.";
    NEED(pre + "edge_mlp.0.weight", {H, 2 * H + 1 + ED}); w.W1 = tmp;
    NEED(pre + "edge_mlp.0.bias", {H}); w.b1 = tmp;
    NEED(pre + "edge_mlp.2.weight", {H, H}); w.W2 = tmp;
    NEED(pre + "edge_mlp.2.bias", {H}); w.b2 = tmp;
    NEED(pre + "node_mlp.0.weight", {H, 2 * H}); w.W3 = tmp;
    NEED(pre + "node_mlp.0.bias", {H}); w.b3 = tmp;
    NEED(pre + "node_mlp.1.weight", {H}); w.gn_w = tmp;
    NEED(pre + "node_mlp.1.bias", {H}); w.gn_b = tmp;
    NEED(pre + "node_mlp.1.mean_scale", {H}); w.gn_ms = tmp;
    NEED(pre + "node_mlp.3.weight", {H, H}); w.W4 = tmp;
    NEED(pre + "node_mlp.3.bias", {H}); w.b4 = tmp;
    NEED(pre + "att_mlp.0.weight", {1, H}); w.wa = tmp;
    NEED(pre + "att_mlp.0.bias", {1}); w.ba = tmp;
    w.Wc1 = w.bc1 = w.wc2 = nullptr;
    if (l == DFM_DEPTH - 1) {
      NEED(pre + "coord_mlp.0.weight", {H, H}); w.Wc1 = tmp;
      NEED(pre + "coord_mlp.0.bias", {H}); w.bc1 = tmp;
      NEED(pre + "coord_mlp.2.weight", {1, H}); w.wc2 = tmp;
    }
    const size_t trows = NSPATIAL + ctx->P;
    if ((rc = dev_alloc(ctx, &w.T32, trows * H))) return rc;
    if ((rc = dev_alloc(ctx, &w.Tdrp16h, (size_t)2 * 40 * 66 * H))) return rc;
    if ((rc = dev_alloc(ctx, &w.Totp16h, (size_t)24 * 24 * 12 * H))) return rc;
    if ((rc = dev_alloc(ctx, &w.w1r, H))) return rc;
    if ((rc = dev_alloc(ctx, &w.b1eff, H))) return rc;
    __half** imgs[] = {&w.img_W1s, &w.img_W1d, &w.img_W4, &w.img_W2h, &w.img_Wc1s, &w.img_W3z0, &w.img_W3z1};
    for (auto pp : imgs) if ((rc = dev_alloc(ctx, pp, (size_t)H * H))) return rc;
    if ((rc = launch_pair_table(ctx, l, s))) return rc;
    if ((rc = launch_image_pack(ctx, w.W1, 641, 0, 1.f, w.img_W1s, s))) return rc;
    if ((rc = launch_image_pack(ctx, w.W1, 641, 256, 1.f, w.img_W1d, s))) return rc;
    if ((rc = launch_image_pack(ctx, w.W4, 256, 0, 1.f, w.img_W4, s))) return rc;
    if ((rc = launch_image_pack(ctx, w.W2, 256, 0, 0.5f, w.img_W2h, s))) return rc;
    if ((rc = launch_image_pack_z(ctx, w.W3, AGG_UNSCALE, w.img_W3z0, w.img_W3z1, s))) return rc;
    if (w.Wc1 && (rc = launch_image_pack_perm(ctx, w.Wc1, 256, 64.f, w.img_Wc1s, s))) return rc;
  }
  NEED("to_energy.0.weight", {H, 2 * H}); ctx->We = tmp;
  NEED("to_energy.1.weight", {H}); ctx->e_ln_w = tmp;
  NEED("to_energy.1.bias", {H}); ctx->e_ln_b = tmp;
  NEED("to_energy.3.weight", {1, H}); ctx->e_w = tmp;
  if ((rc = dev_alloc(ctx, &ctx->img_WeR, (size_t)H * H))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->img_WeL, (size_t)H * H))) return rc;
  if ((rc = launch_image_pack(ctx, ctx->We, 512, 0, 1.f, ctx->img_WeR, s))) return rc;
  if ((rc = launch_image_pack(ctx, ctx->We, 512, 256, 1.f, ctx->img_WeL, s))) return rc;
  NEED("t_embed.0.W", {DFM_INNER_DIM / 2}); ctx->t_W = tmp;
  // the head's small Linear layers are applied one output per thread: stored transposed ([in, out]) the 128 threads of a CTA
  // read consecutive addresses
  {
    float* tl;
    NEED("t_embed.1.weight", {DFM_INNER_DIM, DFM_INNER_DIM});
    if ((rc = dev_alloc(ctx, &tl, (size_t)DFM_INNER_DIM * DFM_INNER_DIM))) return rc;
    if ((rc = launch_transpose(ctx, tmp, DFM_INNER_DIM, DFM_INNER_DIM, tl, s))) return rc;
    ctx->t_lin = tl;
  }
  const char* sc[2] = {"tr_scale", "rot_scale"};
  for (int q = 0; q < 2; ++q) {
    float* w1t;
    NEED(std::string(sc[q]) + ".0.weight", {DFM_INNER_DIM, DFM_INNER_DIM + 1});
    if ((rc = dev_alloc(ctx, &w1t, (size_t)DFM_INNER_DIM * (DFM_INNER_DIM + 1)))) return rc;
    if ((rc = launch_transpose(ctx, tmp, DFM_INNER_DIM, DFM_INNER_DIM + 1, w1t, s))) return rc;
    ctx->sc_W1[q] = w1t;
    NEED(std::string(sc[q]) + ".1.weight", {DFM_INNER_DIM}); ctx->sc_lnw[q] = tmp;
    NEED(std::string(sc[q]) + ".1.bias", {DFM_INNER_DIM}); ctx->sc_lnb[q] = tmp;
    NEED(std::string(sc[q]) + ".4.weight", {1, DFM_INNER_DIM}); ctx->sc_w2[q] = tmp;
  }
  // to_ires.* is optional (dead at inference): present -> dfm_interface_logits works
  ctx->ires_W1t = ctx->ires_W3t = nullptr;
  if (ctx->w.count("to_ires.0.weight")) {
    const float *w1, *w3;
    NEED("to_ires.0.weight", {512, H}); w1 = tmp;
    NEED("to_ires.0.bias", {512}); ctx->ires_b1 = tmp;
    NEED("to_ires.2.weight", {512, 512}); w3 = tmp;
    NEED("to_ires.2.bias", {512}); ctx->ires_b3 = tmp;
    NEED("to_ires.4.weight", {1, 512}); ctx->ires_w5 = tmp;
    NEED("to_ires.4.bias", {1}); ctx->ires_b5 = tmp;
    if ((rc = dev_alloc(ctx, &ctx->ires_W1t, (size_t)512 * H))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->ires_W3t, (size_t)512 * 512))) return rc;
    if ((rc = launch_transpose(ctx, w1, 512, H, ctx->ires_W1t, s))) return rc;
    if ((rc = launch_transpose(ctx, w3, 512, 512, ctx->ires_W3t, s))) return rc;
  }
#undef NEED
  ctx->cut_off = cut_off;
  CUDA_TRY(cudaStreamSynchronize(s));
  ctx->finalized = true;
  ctx->has_complex = false;
  return DFM_OK;
}

__global__ void k_b1eff(const float* __restrict__ b1, const float* __restrict__ T32, int P, float sym, float* __restrict__ out) {
  const int c = threadIdx.x;
  float v = b1[c];
  if (P == 67) v = fmaf(sym, T32[(size_t)(NSPATIAL + 66) * H + c], v);
  out[c] = v;
}

extern "C" int dfm_set_complex(dfm_ctx* ctx, int R, int L, int x_dim, const float* rec_x, const float* lig_x,
                               const float* rec_pos, float sym, void* stream) {
  if (!ctx || !ctx->finalized) { dfm_set_error("dfm_set_complex: weights not finalised"); return DFM_ESTATE; }
  if (R <= 0 || L <= 0 || !rec_x || !lig_x || !rec_pos) { dfm_set_error("dfm_set_complex: bad argument"); return DFM_EINVAL; }
  if (x_dim != ctx->x_dim) { dfm_set_error("dfm_set_complex: x_dim %d != single_embed width %d", x_dim, ctx->x_dim); return DFM_EINVAL; }
  cudaStream_t s = (cudaStream_t)stream;
  CUDA_TRY(cudaSetDevice(ctx->device));
  ctx->R = R; ctx->L = L; ctx->N = R + L;
  const int N = ctx->N;
  // score_net_mlsb.py:89-94
  ctx->knn = N < DFM_KNN ? N : DFM_KNN;
  ctx->ns = N < DFM_KNN ? 0 : (N < DFM_KNN + DFM_NSAMPLE ? N - DFM_KNN : DFM_NSAMPLE);
  ctx->K = ctx->knn + ctx->ns;
  ctx->sym = sym;
  // grow-only arena (4096 residues up front, doubling): the stream is only synchronised when a complex is larger than
  // anything this context has seen -- never in a db5-sized sweep (largest complex 2548 residues)
  if ((size_t)N * H > ctx->h0_cap) {
    size_t cap = ctx->h0_cap ? ctx->h0_cap : (size_t)4096 * H;
    while (cap < (size_t)N * H) cap *= 2;
    CUDA_TRY(cudaStreamSynchronize(s));
    cudaFree(ctx->h0);
    ctx->h0 = nullptr; ctx->h0_cap = 0;
    CUDA_TRY(cudaMalloc(&ctx->h0, sizeof(float) * cap));
    ctx->h0_cap = cap;
  }
  if ((size_t)R * 9 > ctx->rec_cap) {
    size_t cap = ctx->rec_cap ? ctx->rec_cap : (size_t)4096 * 9;
    while (cap < (size_t)R * 9) cap *= 2;
    CUDA_TRY(cudaStreamSynchronize(s));
    cudaFree(ctx->rec_pos);
    ctx->rec_pos = nullptr; ctx->rec_cap = 0;
    CUDA_TRY(cudaMalloc(&ctx->rec_pos, sizeof(float) * cap));
    ctx->rec_cap = cap;
  }
  if (clash_scratch_floats(R, L) > ctx->clash_cap) {
    size_t cap = ctx->clash_cap ? ctx->clash_cap : clash_scratch_floats(256, 256);
    while (cap < clash_scratch_floats(R, L)) cap *= 2;
    CUDA_TRY(cudaStreamSynchronize(s));
    cudaFree(ctx->clash_partial);
    ctx->clash_partial = nullptr; ctx->clash_cap = 0;
    CUDA_TRY(cudaMalloc(&ctx->clash_partial, sizeof(float) * cap));
    ctx->clash_cap = cap;
  }
  CUDA_TRY(cudaMemcpyAsync(ctx->rec_pos, rec_pos, sizeof(float) * (size_t)R * 9, cudaMemcpyDeviceToDevice, s));
  int rc = launch_single_embed(ctx, rec_x, lig_x, s);
  if (rc) return rc;
  for (int l = 0; l < DFM_DEPTH; ++l) {
    k_b1eff<<<1, 256, 0, s>>>(ctx->layer[l].b1, ctx->layer[l].T32, ctx->P, sym, ctx->layer[l].b1eff);
    LAUNCH_CHECK(ctx);
  }
  ctx->has_complex = true;
  return DFM_OK;
}

extern "C" int dfm_set_receptor_pose(dfm_ctx* ctx, const float* rec_pos, void* stream) {
  if (!ctx || !ctx->has_complex) { dfm_set_error("no complex set"); return DFM_ESTATE; }
  if (!rec_pos) { dfm_set_error("dfm_set_receptor_pose: null rec_pos"); return DFM_EINVAL; }
  CUDA_TRY(cudaSetDevice(ctx->device));
  CUDA_TRY(cudaMemcpyAsync(ctx->rec_pos, rec_pos, sizeof(float) * (size_t)ctx->R * 9, cudaMemcpyDeviceToDevice,
                           (cudaStream_t)stream));
  return DFM_OK;
}

// ---------------------------------------------------------------------------------------------------
static int forward_impl(dfm_ctx* ctx, int B, const float* lig_pos, const float* t, const int32_t* edges,
                        const float* exp_noise, uint64_t seed, uint64_t stream_base, uint32_t fwd_index, uint32_t flags,
                        float* tr_score, float* rot_score, float* f, float* energy, int32_t* clashes, int32_t* edges_out,
                        Workspace& ws, cudaStream_t s) {
  const bool fp32 = (flags & DFM_PRECISION_FP32) != 0;
  const bool want_energy = (flags & DFM_WANT_ENERGY) != 0;
  const int N = ctx->N, M = B * N;
  int rc;
  dfm_pdl_set_rows(M);
  if ((rc = launch_prepare(ctx, B, lig_pos, ws, s))) return rc;
  if ((rc = launch_graph(ctx, B, (flags & DFM_GRAPH_GENERIC) != 0, edges, exp_noise, seed, stream_base, fwd_index, ws, s))) return rc;
  if (edges_out) {
    CUDA_TRY(cudaMemcpy2DAsync(edges_out, sizeof(int32_t) * ctx->K, ws.nbr, sizeof(int32_t) * SLOTS,
                               sizeof(int32_t) * ctx->K, (size_t)M, cudaMemcpyDeviceToDevice, s));
  }
  if ((rc = launch_broadcast_h0(ctx, B, ws, s))) return rc;
  for (int l = 0; !fp32 && l < DFM_DEPTH; ++l) {
    // throughput path: fp16 activations between kernels, warp-specialised edge kernel, fused node-side GEMMs
    const bool last = l == DFM_DEPTH - 1;
    __half* Ah = reinterpret_cast<__half*>(ws.A);
    __half* Bm = reinterpret_cast<__half*>(ws.Bm);
    if ((rc = launch_node_ab(ctx, l, M, ws.h16, Ah, Bm, s))) return rc;
    EdgeArgs ea{};
    ea.B = B; ea.N = N; ea.R = ctx->R; ea.K = ctx->K; ea.layer = l; ea.last = last;
    ea.lig_only = last && !want_energy;   // the receptor rows' layer-5 messages only feed the node update the energy head needs
    ea.nbr = ws.nbr; ea.feat = ws.feat; ea.radial = ws.radial; ea.A = ws.A; ea.Bm = ws.Bm; ea.pos = ws.pos;
    ea.agg = ws.agg; ea.mstar = ws.mstar; ea.fbuf = ws.fbuf;
    // timed launches (dfm_profile_*): the five full-size launches of a forward; the last layer's launch walks only the
    // ligand tiles and may carry the coordinate head (k_last_fused), so it is not an edge-kernel sample
    const bool prof = ctx->profile && !last && ctx->prof_used + 2 <= ctx->prof_events.size();
    if (prof) CUDA_TRY(cudaEventRecord(ctx->prof_events[ctx->prof_used], s));
    int fused = 0;      // last layer without the energy head: edge MLP + coordinate head in one launch when it applies
    if (last && ea.lig_only && (rc = launch_last_fused(ctx, ea, ws.emeta, Ah, ws.ring_flags, RING_FLAG_WORDS, (flags & DFM_LAST_FUSED) != 0, &fused, s))) return rc;
    if (!fused && (rc = launch_edge_ws(ctx, ea, ws.emeta, Ah, ws.agg16, s))) return rc;
    if (prof) {
      CUDA_TRY(cudaEventRecord(ctx->prof_events[ctx->prof_used + 1], s));
      ctx->prof_used += 2;
    }
    if (last && !fused && (rc = launch_node_coord(ctx, ea, s))) return rc;
    if (last && !want_energy) break;   // layer-5 node update only feeds the energy head
    if ((rc = launch_node_z(ctx, l, M, ws.h16, ws.agg16, ws.z, s))) return rc;
    if ((rc = launch_graphnorm_stats(ctx, B, l, ws.z, ws.gscale, ws.gshift, s))) return rc;
    if ((rc = launch_node_h(ctx, l, M, ws.z, ws.gscale, ws.gshift, ws.h, ws.h16, s))) return rc;
  }
  for (int l = 0; fp32 && l < DFM_DEPTH; ++l) {
    // parity path (DFM_PRECISION_FP32): fp32 FFMA kernels of simt.cu, one Linear at a time
    const LayerW& w = ctx->layer[l];
    const bool last = l == DFM_DEPTH - 1;
    LinearArgs la{};
    la.A = ws.h; la.a_scale = 1.f; la.M = M;
    // A = W1s h + b1
    la.W32 = w.W1; la.ldw = 641; la.w_col0 = 0; la.bias = w.b1eff; la.add = nullptr; la.out = ws.A;
    if ((rc = launch_linear_simt(ctx, la, s))) return rc;
    // Bm = W1d h
    la.w_col0 = 256; la.bias = nullptr; la.out = ws.Bm;
    if ((rc = launch_linear_simt(ctx, la, s))) return rc;
    EdgeArgs ea{};
    ea.B = B; ea.N = N; ea.R = ctx->R; ea.K = ctx->K; ea.layer = l; ea.last = last;
    ea.nbr = ws.nbr; ea.feat = ws.feat; ea.radial = ws.radial; ea.A = ws.A; ea.Bm = ws.Bm; ea.pos = ws.pos;
    ea.agg = ws.agg; ea.mstar = ws.mstar; ea.fbuf = ws.fbuf;
    if ((rc = launch_edge_simt(ctx, ea, s))) return rc;
    if (last && !want_energy) break;   // layer-5 node update only feeds the energy head
    // z = W3h h + b3 + W3a agg
    la.A = ws.h; la.W32 = w.W3; la.ldw = 512; la.w_col0 = 0; la.bias = w.b3; la.add = nullptr; la.out = ws.z;
    if ((rc = launch_linear_simt(ctx, la, s))) return rc;
    la.A = ws.agg; la.w_col0 = 256; la.bias = nullptr; la.add = ws.z;
    if ((rc = launch_linear_simt(ctx, la, s))) return rc;
    if ((rc = launch_graphnorm_silu(ctx, B, l, ws, s))) return rc;
    // h += W4 y + b4
    la.A = ws.y; la.W32 = w.W4; la.ldw = 256; la.w_col0 = 0; la.bias = w.b4; la.add = ws.h; la.out = ws.h;
    if ((rc = launch_linear_simt(ctx, la, s))) return rc;
  }
  float* trs = tr_score ? tr_score : ws.tsc;
  float* rots = rot_score ? rot_score : ws.tsc + (size_t)B * 4;
  if ((rc = launch_force_head(ctx, B, t, ws, trs, rots, f, s))) return rc;
  if (want_energy && (rc = launch_energy(ctx, B, fp32, ws, energy, clashes, s))) return rc;
  return 0;
}

static int check_ws(dfm_ctx* ctx, int B, void* workspace, size_t bytes, Workspace* ws) {
  if (!ctx || !ctx->has_complex) { dfm_set_error("no complex set"); return DFM_ESTATE; }
  if (B <= 0) { dfm_set_error("B must be positive"); return DFM_EINVAL; }
  *ws = carve_workspace(ctx, B, workspace);
  if (!workspace || bytes < ws->bytes) {
    dfm_set_error("workspace too small: %zu < %zu bytes", bytes, ws->bytes);
    return DFM_ENOMEM;
  }
  if (reinterpret_cast<uintptr_t>(workspace) & 255) { dfm_set_error("workspace must be 256-byte aligned"); return DFM_EINVAL; }
  return 0;
}

extern "C" int dfm_score_forward(dfm_ctx* ctx, int B, const float* lig_pos, const float* t, const int32_t* edges,
                                 const float* exp_noise, uint64_t seed, uint64_t stream_base, uint32_t forward_index,
                                 uint32_t flags, float* tr_score, float* rot_score, float* f, float* energy,
                                 int32_t* num_clashes, int32_t* edges_out, void* workspace, size_t workspace_bytes,
                                 void* stream) {
  Workspace ws;
  int rc = check_ws(ctx, B, workspace, workspace_bytes, &ws);
  if (rc) return rc;
  if (!lig_pos || !t) { dfm_set_error("dfm_score_forward: lig_pos and t are required"); return DFM_EINVAL; }
  CUDA_TRY(cudaSetDevice(ctx->device));
  return forward_impl(ctx, B, lig_pos, t, edges, exp_noise, seed, stream_base, forward_index, flags, tr_score, rot_score,
                      f, energy, num_clashes, edges_out, ws, (cudaStream_t)stream);
}

extern "C" int dfm_reverse_step(dfm_ctx* ctx, int B, float* lig_pos, float* rot_update, float* tr_update,
                                const float* tr_score, const float* rot_score, float g_rot, float g_tr, float dt,
                                float ns_rot, float ns_tr, const float* z, uint64_t seed, uint64_t stream_base,
                                uint32_t step_index, uint32_t flags, void* stream) {
  if (!ctx || !ctx->has_complex) { dfm_set_error("no complex set"); return DFM_ESTATE; }
  if (B <= 0 || !lig_pos || !rot_update || !tr_update || !tr_score || !rot_score) { dfm_set_error("dfm_reverse_step: bad argument"); return DFM_EINVAL; }
  CUDA_TRY(cudaSetDevice(ctx->device));
  return launch_reverse_step(ctx, B, lig_pos, rot_update, tr_update, tr_score, rot_score, g_rot, g_tr, dt, ns_rot, ns_tr, z,
                             seed, stream_base, step_index, flags, (cudaStream_t)stream);
}

extern "C" int dfm_randomize_pose(dfm_ctx* ctx, int B, const float* lig_pos0, const float* rot0, const float* tr0,
                                  uint64_t seed, uint64_t stream_base, uint32_t flags, float* lig_pos, float* rot_update,
                                  float* tr_update, void* stream) {
  if (!ctx || !ctx->has_complex) { dfm_set_error("no complex set"); return DFM_ESTATE; }
  if (B <= 0 || !lig_pos0 || !lig_pos || !rot_update || !tr_update) { dfm_set_error("dfm_randomize_pose: bad argument"); return DFM_EINVAL; }
  CUDA_TRY(cudaSetDevice(ctx->device));
  return launch_randomize_pose(ctx, B, lig_pos0, rot0, tr0, seed, stream_base, flags, lig_pos, rot_update, tr_update,
                               (cudaStream_t)stream);
}

// host-side schedules, fp64 like the reference's numpy (so3_diffuser.py:210-227, r3_diffuser.py:20-24); the sigmas are the
// checkpoint's hyper_parameters.diffuser values (dfm_set_schedule; defaults = both shipped checkpoints)
static double so3_g(const dfm_ctx* c, double t) {
  const double lo = c->so3_min_sigma, hi = c->so3_max_sigma;
  const double sg = log(t * exp(hi) + (1 - t) * exp(lo));
  return sqrt(2 * (exp(hi) - exp(lo)) * sg / exp(sg));
}
static double r3_g(const dfm_ctx* c, double t) {
  const double lo = c->r3_min_sigma, hi = c->r3_max_sigma;
  return lo * pow(hi / lo, t) * sqrt(2 * (log(hi) - log(lo)));
}

extern "C" int dfm_set_schedule(dfm_ctx* ctx, double so3_min_sigma, double so3_max_sigma, double r3_min_sigma,
                                double r3_max_sigma) {
  if (!ctx) { dfm_set_error("null ctx"); return DFM_EINVAL; }
  if (!(so3_min_sigma > 0) || !(so3_max_sigma > so3_min_sigma) || !(r3_min_sigma > 0) || !(r3_max_sigma > r3_min_sigma)) {
    dfm_set_error("dfm_set_schedule: need 0 < min_sigma < max_sigma for both diffusers");
    return DFM_EINVAL;
  }
  ctx->so3_min_sigma = so3_min_sigma; ctx->so3_max_sigma = so3_max_sigma;
  ctx->r3_min_sigma = r3_min_sigma; ctx->r3_max_sigma = r3_max_sigma;
  return DFM_OK;
}

extern "C" int dfm_interface_logits(dfm_ctx* ctx, int B, float* ires, void* workspace, size_t workspace_bytes, void* stream) {
  Workspace ws;
  int rc = check_ws(ctx, B, workspace, workspace_bytes, &ws);
  if (rc) return rc;
  if (!ires) { dfm_set_error("dfm_interface_logits: null output"); return DFM_EINVAL; }
  if (!ctx->ires_W1t) { dfm_set_error("dfm_interface_logits: the to_ires.* weights were not supplied"); return DFM_EMISSING; }
  CUDA_TRY(cudaSetDevice(ctx->device));
  return launch_ires(ctx, B * ctx->N, ws.h, ires, (cudaStream_t)stream);
}

__global__ void k_fill(float* p, float v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

extern "C" int dfm_sample(dfm_ctx* ctx, int B, const float* lig_pos0, int num_steps, float eps, float tr_noise_scale,
                          float rot_noise_scale, uint32_t flags, uint64_t seed, uint64_t stream_base, float* lig_pos,
                          float* rot_update, float* tr_update, float* energy, int32_t* num_clashes, void* workspace,
                          size_t workspace_bytes, void* stream) {
  Workspace ws;
  int rc = check_ws(ctx, B, workspace, workspace_bytes, &ws);
  if (rc) return rc;
  if (!lig_pos0 || !lig_pos || !rot_update || !tr_update || num_steps < 2) { dfm_set_error("dfm_sample: bad argument"); return DFM_EINVAL; }
  cudaStream_t s = (cudaStream_t)stream;
  CUDA_TRY(cudaSetDevice(ctx->device));
  if ((rc = launch_randomize_pose(ctx, B, lig_pos0, nullptr, nullptr, seed, stream_base, flags, lig_pos, rot_update, tr_update, s))) return rc;
  // torch.linspace(1, eps, S) in fp32 and dt = ts[0] - ts[1]  (inference_base.py:404-405)
  std::vector<float> ts(num_steps);
  {
    const float step = (eps - 1.0f) / (float)(num_steps - 1);
    for (int i = 0; i < num_steps; ++i)
      ts[i] = (i < num_steps / 2) ? fmaf(step, (float)i, 1.0f) : fmaf(-step, (float)(num_steps - 1 - i), eps);
  }
  const float dt = ts[0] - ts[1];
  float* tbuf = ws.tsc + (size_t)B * 7;   // [B] time values live in the tail of the scratch block
  float* trs = ws.tsc;
  float* rots = ws.tsc + (size_t)B * 4;
  const uint32_t fflags = flags & (DFM_PRECISION_FP32 | DFM_GRAPH_GENERIC);
  for (int i = 0; i < num_steps; ++i) {
    const bool last = i == num_steps - 1;
    k_fill<<<(B + 255) / 256, 256, 0, s>>>(tbuf, ts[i], B);
    LAUNCH_CHECK(ctx);
    if ((rc = forward_impl(ctx, B, lig_pos, tbuf, nullptr, nullptr, seed, stream_base, (uint32_t)i, fflags, trs, rots,
                           nullptr, nullptr, nullptr, nullptr, ws, s))) return rc;
    float ns_tr, ns_rot;
    if (flags & DFM_NOISE_ANNEAL) ns_tr = ns_rot = ts[i];
    else if (last) ns_tr = ns_rot = 0.f;
    else { ns_tr = tr_noise_scale; ns_rot = rot_noise_scale; }
    if ((rc = launch_reverse_step(ctx, B, lig_pos, rot_update, tr_update, trs, rots, (float)so3_g(ctx, (double)ts[i]),
                                  (float)r3_g(ctx, (double)ts[i]), dt, ns_rot, ns_tr, nullptr, seed, stream_base, (uint32_t)i,
                                  flags, s))) return rc;
    if (last) {
      if ((rc = forward_impl(ctx, B, lig_pos, tbuf, nullptr, nullptr, seed, stream_base, (uint32_t)num_steps,
                             fflags | DFM_WANT_ENERGY, trs, rots, nullptr, energy, num_clashes, nullptr, ws, s))) return rc;
    }
  }
  return DFM_OK;
}

extern "C" int dfm_profile_enable(dfm_ctx* ctx, int max_launches) {
  if (!ctx) { dfm_set_error("null ctx"); return DFM_EINVAL; }
  CUDA_TRY(cudaSetDevice(ctx->device));
  for (cudaEvent_t e : ctx->prof_events) cudaEventDestroy(e);
  ctx->prof_events.clear();
  ctx->prof_used = 0;
  ctx->profile = max_launches > 0;
  for (int i = 0; i < 2 * max_launches; ++i) {
    cudaEvent_t e;
    CUDA_TRY(cudaEventCreate(&e));
    ctx->prof_events.push_back(e);
  }
  return DFM_OK;
}

extern "C" int dfm_profile_read(dfm_ctx* ctx, double* edge_kernel_ms, int* launches) {
  if (!ctx || !edge_kernel_ms || !launches) { dfm_set_error("dfm_profile_read: bad argument"); return DFM_EINVAL; }
  CUDA_TRY(cudaSetDevice(ctx->device));
  double total = 0.0;
  for (size_t i = 0; i + 1 < ctx->prof_used; i += 2) {
    CUDA_TRY(cudaEventSynchronize(ctx->prof_events[i + 1]));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, ctx->prof_events[i], ctx->prof_events[i + 1]));
    total += ms;
  }
  *edge_kernel_ms = total;
  *launches = (int)(ctx->prof_used / 2);
  ctx->prof_used = 0;
  return DFM_OK;
}

extern "C" int64_t dfm_debug_read(dfm_ctx* ctx, int B, int which, void* out, size_t out_bytes, void* workspace,
                                  void* stream) {
  if (!ctx || !ctx->has_complex || !out || !workspace || B <= 0) { dfm_set_error("dfm_debug_read: bad argument"); return DFM_EINVAL; }
  Workspace ws = carve_workspace(ctx, B, workspace);
  const void* src = nullptr;
  size_t n = 0, esz = 4;
  switch (which) {
    case 0: src = ws.h; n = (size_t)B * ctx->N * H; break;
    case 1: src = ws.feat; n = (size_t)B * ctx->N * SLOTS; break;
    case 2: src = ws.radial; n = (size_t)B * ctx->N * SLOTS; break;
    case 3: src = ws.nbr; n = (size_t)B * ctx->N * SLOTS; break;
    case 4: src = ws.agg; n = (size_t)B * ctx->N * H; break;
    case 5: src = ws.A; n = (size_t)B * ctx->N * H; break;
    case 6: src = ws.fbuf; n = (size_t)B * ctx->L * 4; break;
    default: dfm_set_error("dfm_debug_read: unknown buffer %d", which); return DFM_EINVAL;
  }
  if (out_bytes < n * esz) { dfm_set_error("dfm_debug_read: out too small"); return DFM_ENOMEM; }
  if (cudaMemcpyAsync(out, src, n * esz, cudaMemcpyDeviceToDevice, (cudaStream_t)stream) != cudaSuccess) {
    dfm_set_error("dfm_debug_read: copy failed");
    return DFM_ECUDA;
  }
  return (int64_t)n;
}
