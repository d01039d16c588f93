/*
 * dfmdock_b200 -- C ABI of the B200-native DFMDock reverse-diffusion docking sampler.
 *
 * The reference (Graylab/DFMDock) is pure Python/PyTorch and has no FFI layer; its seam for this
 * path is Python duck typing.  Each entry point below names the reference
 * interface it replaces (paths relative to the reference repository root).  INTEGRATION.md shows
 * the ctypes binding a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success or a negative DFM_E* code; nothing throws across the ABI;
 *     dfm_last_error() returns a thread-local message for the last failure.
 *   - all tensor arguments are DEVICE pointers (fp32 / int32, row-major, contiguous) unless a
 *     parameter name ends in _host.  The caller owns every buffer passed in; the library only
 *     owns what dfm_create / dfm_set_weight / dfm_set_complex allocate inside the context and
 *     allocates nothing in dfm_score_forward / dfm_reverse_step / dfm_sample (the workspace is
 *     handed in by the caller, sized by dfm_workspace_bytes).
 *   - all work is stream-ordered on the cudaStream_t passed as `stream` (a void* here so that the
 *     header needs no CUDA include).  Entry points that synchronise: dfm_create, dfm_finalize_weights,
 *     dfm_destroy (device / stream), dfm_profile_read (its own events), and dfm_set_complex ONLY when the
 *     complex is larger than the context's grow-only arena (4096 residues up front, doubled on demand; the
 *     arena also holds the clash-force scratch of dfm_reverse_step / dfm_sample): then it synchronises
 *     `stream` once and re-allocates.  Nothing else synchronises or allocates.
 *   - one context per (device, stream); contexts are independent; a context is not thread-safe.
 *   - there is NO CPU fallback: without a CUDA device every entry point fails with DFM_ECUDA.
 */
#ifndef DFMDOCK_B200_H_
#define DFMDOCK_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dfm_ctx dfm_ctx;

enum {
  DFM_OK = 0,
  DFM_EINVAL = -1,    /* bad argument / shape mismatch */
  DFM_ECUDA = -2,     /* CUDA runtime error (message has the cudaError string) */
  DFM_ESTATE = -3,    /* call out of order (weights not finalised, no complex set, ...) */
  DFM_ENOMEM = -4,    /* workspace too small */
  DFM_EMISSING = -5   /* a required weight tensor was never supplied */
};

/* dfm_score_forward / dfm_sample flags */
enum {
  DFM_WANT_ENERGY = 1u << 0,   /* run the pair-energy head + clash count (final forward only in the sampler); without it the
                                  last layer only processes the ligand residues' edges: the receptor rows' layer-5 messages
                                  feed a node update nothing but the energy head consumes */
  DFM_PRECISION_FP32 = 1u << 1,/* fp32 FFMA kernels (parity mode); default = fp16-operand tcgen05 kernels, fp32 accumulate */
  DFM_CLASH_FORCE = 1u << 2,   /* inference.py:358-361 soft-clash translation after every step */
  DFM_NOISE_ANNEAL = 1u << 3,  /* noise_scale = t (inference_base.py:428-430) */
  DFM_CENTRE_ALL_ATOMS = 1u << 4, /* rotate about the N/CA/C centroid (inference.py:224-245) instead of the CA centroid (inference_base.py:322-343) */
  DFM_ODE = 1u << 5,           /* probability-flow ODE branch of torch_reverse (so3_diffuser.py:366-367) */
  DFM_GRAPH_GENERIC = 1u << 6, /* build the stochastic graph with the generic (shared-memory) kernel that complexes of more than
                                  1024 residues use, whatever the size: same neighbours as the register-resident kernels (test knob) */
  DFM_LAST_FUSED = 1u << 7     /* last layer without the energy head as ONE launch: the SMs are split into an edge-MLP role and a
                                  coordinate-head role, the gated messages travel through an L2-resident ring instead of HBM
                                  (same results bit for bit; measured slower than the two-kernel form on B200, so it is off by
                                  default -- DESIGN.md section 6; the environment variable DFM_LAST_FUSED=1 turns it on globally) */
};

/* Fixed architecture of the shipped checkpoints (configs/model/score_model_mlsb.yaml). */
#define DFM_NODE_DIM 256
#define DFM_EDGE_DIM 128
#define DFM_INNER_DIM 128
#define DFM_DEPTH 6
#define DFM_KNN 20
#define DFM_NSAMPLE 40
#define DFM_EDGE_SLOTS 64 /* per-node stride of every [B, N, slot] edge array (60 used) */

/* Replaces: Score_Model.load_from_checkpoint(...) -> Score_Net.__init__ (src/models/score_model_mlsb.py:22-59,
 * src/models/score_net_mlsb.py:251-330).  Creates an empty context on CUDA device `device`. */
int dfm_create(dfm_ctx** out, int device);

/* Replaces: load_state_dict of one tensor.  `name` is the checkpoint key without the leading "net."
 * (e.g. "network.EGNN_3.egcl.edge_mlp.0.weight"); `data` is a DEVICE fp32 pointer in nn.Linear [out, in]
 * row-major layout; the tensor is copied.  Unknown names are stored and ignored. */
int dfm_set_weight(dfm_ctx* ctx, const char* name, const float* data, const int64_t* shape, int ndim);

/* Validates that all 104-10 hot-path tensors are present with the expected shapes, then builds the
 * derived device tables: per-layer pair tables T_l = W1e_l * [spatial_embed | positional_embed]
 * (SURVEY App. A.5), fp16 weight images in the tcgen05 shared-memory layout, folded biases.
 * cut_off = hyper_parameters.model.cut_off (score_net_mlsb.py:264). Synchronises. */
int dfm_finalize_weights(dfm_ctx* ctx, float cut_off, void* stream);

/* Replaces: the per-complex, pose-invariant part of Score_Net.forward: single_embed(x)
 * (score_net_mlsb.py:365-366) and get_position_matrix (inference_base.py:230-244).
 * rec_x [R, x_dim], lig_x [L, x_dim] (x_dim = 1301), rec_pos [R,3,3] (N, CA, C; Angstrom).
 * sym = value of positional channel 66 for 67-wide checkpoints, ignored otherwise. */
int dfm_set_complex(dfm_ctx* ctx, int R, int L, int x_dim, const float* rec_x, const float* lig_x,
                    const float* rec_pos, float sym, void* stream);

/* Replaces: the sigmas SO3Diffuser / R3Diffuser take from hyper_parameters.diffuser (src/utils/so3_diffuser.py:22-31,
 * src/utils/r3_diffuser.py:9-18), used by dfm_sample for g(t).  Defaults 0.1 / 1.5 / 0.1 / 30.0 (both shipped checkpoints). */
int dfm_set_schedule(dfm_ctx* ctx, double so3_min_sigma, double so3_max_sigma, double r3_min_sigma, double r3_max_sigma);

/* Replaces only the receptor backbone [R,3,3] of the current complex (the reference sampler re-sends
 * batch["rec_pos"] every step, inference_base.py:422); R must equal the current complex's. */
int dfm_set_receptor_pose(dfm_ctx* ctx, const float* rec_pos, void* stream);

/* Bytes of scratch dfm_score_forward / dfm_sample need for B simultaneous trajectories of the current complex. */
size_t dfm_workspace_bytes(const dfm_ctx* ctx, int B);

/* Number of edges per node for the current complex: min(N, 60) (score_net_mlsb.py:89-94). */
int dfm_edges_per_node(const dfm_ctx* ctx);

/* Replaces: Score_Model.forward(batch) (src/models/score_model_mlsb.py:61-63 -> score_net_mlsb.py:343-425)
 * for B independent ligand poses of the current complex.
 *   lig_pos  [B, L, 3, 3]   current ligand backbone poses
 *   t        [B]            diffusion times
 *   edges    [B, N, K] int32 or NULL  injected neighbour table (parity mode), K = dfm_edges_per_node
 *   exp_noise[B, N, N-20] or NULL     injected Exp(1) draws in torch.multinomial's compacted order (parity mode)
 *   seed, stream_base, forward_index  Philox key/counter when neither is injected: trajectory b uses
 *                                     subsequence stream_base + b, so results do not depend on sharding
 * Outputs (any may be NULL): tr_score [B,3], rot_score [B,3], f [B,L,3], energy [B], num_clashes [B] int32
 * (energy / num_clashes are only written with DFM_WANT_ENERGY), edges_out [B,N,K] int32. */
int dfm_score_forward(dfm_ctx* ctx, int B, const float* lig_pos, const float* t, const int32_t* edges,
                      const float* exp_noise, uint64_t seed, uint64_t stream_base, uint32_t forward_index,
                      uint32_t flags, float* tr_score, float* rot_score, float* f, float* energy,
                      int32_t* num_clashes, int32_t* edges_out, void* workspace, size_t workspace_bytes,
                      void* stream);

/* Replaces: one body of the reverse loop after the forward (inference_base.py:439-461):
 * so3/r3 torch_reverse (so3_diffuser.py:344-369, r3_diffuser.py:40-55), modify_coords (:342-352),
 * tr_update/rot_compose accumulation (:455-456, :311-316) and, with DFM_CLASH_FORCE, get_clash_force (:366-384).
 *   g_rot, g_tr     diffusion coefficients g(t) (computed by the host in fp64 like the reference, passed as float)
 *   dt              step size (time_steps[0]-time_steps[1])
 *   ns_rot, ns_tr   noise scales for this step
 *   z [B, 2, 3] or NULL   injected N(0,1) draws (z[b,0]=rotation, z[b,1]=translation); NULL -> Philox
 * In/out: lig_pos [B,L,3,3], rot_update [B,3] (axis-angle), tr_update [B,3]. */
int dfm_reverse_step(dfm_ctx* ctx, int B, float* lig_pos, float* rot_update, float* tr_update,
                     const float* tr_score, const float* rot_score, float g_rot, float g_tr, float dt,
                     float ns_rot, float ns_tr, const float* z, uint64_t seed, uint64_t stream_base,
                     uint32_t step_index, uint32_t flags, void* stream);

/* Replaces: randomize_pose (inference_base.py:318-340 / inference.py:220-242) for B trajectories.
 *   lig_pos0 [L,3,3]  input ligand pose;  rot0 [B,3,3] / tr0 [B,3] or NULL: injected initial rotation matrices
 *   (scipy Rotation.random().as_matrix()) and N(0, 30^2) translation draws; NULL -> Philox.
 * Out: lig_pos [B,L,3,3], rot_update [B,3], tr_update [B,3]. */
int dfm_randomize_pose(dfm_ctx* ctx, int B, const float* lig_pos0, const float* rot0, const float* tr0,
                       uint64_t seed, uint64_t stream_base, uint32_t flags, float* lig_pos,
                       float* rot_update, float* tr_update, void* stream);

/* Replaces: Euler_Maruyama_sampler (inference_base.py:390-468) for B trajectories in lock step plus the
 * serial per-trajectory driver loop (inference_base.py:644-657): randomize_pose, num_steps x
 * {forward, reverse step}, final forward with energy.  All noise from Philox (seed, stream_base + b).
 * Out: lig_pos [B,L,3,3], rot_update [B,3], tr_update [B,3], energy [B], num_clashes [B] int32. */
int dfm_sample(dfm_ctx* ctx, int B, const float* lig_pos0, int num_steps, float eps, float tr_noise_scale,
               float rot_noise_scale, uint32_t flags, uint64_t seed, uint64_t stream_base, float* lig_pos,
               float* rot_update, float* tr_update, float* energy, int32_t* num_clashes, void* workspace,
               size_t workspace_bytes, void* stream);

/* Replaces: ires = to_ires(h) (src/models/score_net_mlsb.py:296-302, :383), the sixth entry of the reference's output
 * dict.  Nothing reads it at inference (inference_base.py:494-500), so it is computed on request only: call after a
 * dfm_score_forward with DFM_WANT_ENERGY (which runs the last node update) on the same B / workspace.  ires [B, N].
 * Needs the to_ires.* tensors (dfm_set_weight); DFM_EMISSING otherwise. */
int dfm_interface_logits(dfm_ctx* ctx, int B, float* ires, void* workspace, size_t workspace_bytes, void* stream);

/* Number of kernels the library launched on behalf of this context since creation (bench.py "gpu_launches"). */
uint64_t dfm_launch_count(const dfm_ctx* ctx);

/* Measurement hooks (bench.py "roofline"): when enabled, every launch of the dominant kernel (the fused edge
 * kernel) is bracketed by CUDA events on the caller's stream.  dfm_profile_read synchronises on those events and
 * returns the summed kernel time and the number of launches since the last read. */
int dfm_profile_enable(dfm_ctx* ctx, int max_launches);
int dfm_profile_read(dfm_ctx* ctx, double* edge_kernel_ms, int* launches);

/* Debug / parity taps: copy an internal 4-byte-element buffer of the last dfm_score_forward (same B, same
 * workspace) into the DEVICE buffer `out`.  which: 0 = node features h after the last node update [B,N,256],
 * 1 = packed pair-feature bins [B,N,64] (uint32: d | omega<<6 | theta<<11 | phi<<16 | relpos<<20),
 * 2 = radial [B,N,64], 3 = neighbour table [B,N,64] int32, 4 = last agg [B,N,256], 5 = last A [B,N,256],
 * 6 = per-residue force [B,L,4].  Returns the element count or a negative error. */
int64_t dfm_debug_read(dfm_ctx* ctx, int B, int which, void* out, size_t out_bytes, void* workspace, void* stream);

/* ---- the consumer of the sampler's output ------------------------------------------------
 * Replaces: compute_metrics (src/utils/metrics.py:3-16: get_c_rmsd :33-38, get_i_rmsd :40-46, get_l_rmsd :48-56,
 * get_fnat :58-69, get_DockQ :71-74, find_rigid_alignment :91-121) for T docked poses of one complex, batched.
 *   model_rec  [R,3,3] when rec_is_shared != 0 (the sampler never moves the receptor) else [T,R,3,3]
 *   model_lig  [T,L,3,3];  native_rec [R,3,3];  native_lig [L,3,3]   (backbone N, CA, C; Angstrom; DEVICE pointers)
 *   out        [T,5] = c_rmsd, i_rmsd, l_rmsd, fnat, DockQ  (i_rmsd / DockQ are NaN when the native has no interface)
 *   workspace  dfm_metrics_workspace_bytes(R, L) bytes of device scratch
 * Needs no context; stream-ordered on `stream`; returns 0 or a negative DFM_E* code. */
size_t dfm_metrics_workspace_bytes(int R, int L);
int dfm_compute_metrics(int device, int T, int R, int L, const float* model_rec, int rec_is_shared,
                        const float* model_lig, const float* native_rec, const float* native_lig, float* out,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ---- all-atom output ---------------------------------------------------------------------
 * Replaces: modify_aa_coords -- the rigid transform the sampler accumulated (rot_update, tr_update) applied to the
 * ligand's all-atom coordinates before the structure is written:
 *   centre_mode 0: rotate about the ligand's backbone CA centroid   (src/inference_base.py:354-364; lig_bb [L,3,3])
 *   centre_mode 1: rotate about the all-atom centroid of `atoms`    (src/inference.py:256-266; lig_bb may be NULL)
 *   atoms [A,3] (shared by all poses), rot_update [T,3] axis-angle, tr_update [T,3]  ->  out [T,A,3]; DEVICE pointers.
 * Needs no context; stream-ordered on `stream`; returns 0 or a negative DFM_E* code. */
int dfm_transform_atoms(int device, int T, int A, int L, int centre_mode, const float* atoms, const float* lig_bb,
                        const float* rot_update, const float* tr_update, float* out, void* stream);

const char* dfm_last_error(void);
const char* dfm_version(void);
void dfm_destroy(dfm_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* DFMDOCK_B200_H_ */
