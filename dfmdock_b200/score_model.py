"""Score_Model: host-side mirror of the reference's model wrapper over the CUDA library.

Reference interface mirrored (src/models/score_model_mlsb.py:22-63):
    model = Score_Model.load_from_checkpoint(path, map_location=device); model.to(device).eval()
    out = model(batch)   # dict: tr_score[1,3] rot_score[1,3] energy[] f[L,3] num_clashes[] ires[N,1]
    model.so3_diffuser.torch_reverse(...), model.r3_diffuser.torch_reverse(...)
plus the batched calls the reference does not have (many trajectories of one complex per launch):
    model.set_complex(batch); model.score(lig_pos[B,L,3,3], t[B], ...); model.reverse_step(...); model.sample(...)

All compute happens in libdfmdock_b200.so (hand-written sm_100a kernels); torch is only used for device memory,
streams and RNG draws.  There is no CPU path: .to("cpu") raises.
"""
import ctypes
from ctypes import c_int64, c_void_p

import torch

from . import _lib
from .checkpoint import load_checkpoint
from .diffusers import R3Diffuser, SO3Diffuser

HOT_PREFIXES = ("single_embed", "spatial_embed", "positional_embed", "network.", "to_energy", "to_ires", "t_embed",
                "tr_scale", "rot_scale")

# hyper_parameters.model values the kernels are built for (configs/model/score_model_mlsb.yaml; both shipped checkpoints)
SUPPORTED_MODEL_HPARAMS = {"node_dim": 256, "edge_dim": 128, "inner_dim": 128, "depth": 6, "spatial_embed_dim": 100,
                           "normalize": True}


def validate_hparams(hparams):
    """The CUDA kernels hard-wire the shipped architecture: 6 E_GCL layers, 256 / 128 / 128 widths, normalised coordinate
    differences (egnn.py:144-146) and GraphNorm.  A checkpoint trained with anything else would load (same tensor shapes
    for e.g. normalize=False) and give silently different scores -- refuse it instead."""
    model = hparams.get("model", {}) if hasattr(hparams, "get") else {}
    for key, want in SUPPORTED_MODEL_HPARAMS.items():
        if key in model and model[key] != want:
            raise ValueError("dfmdock_b200: hyper_parameters.model.%s = %r is not supported (the kernels are built for %r)"
                             % (key, model[key], want))
    so3 = hparams.get("diffuser", {}).get("so3", {}) if hasattr(hparams, "get") else {}
    if so3.get("schedule", "logarithmic") != "logarithmic":
        raise ValueError("dfmdock_b200: only the logarithmic SO(3) schedule is supported, got %r" % (so3.get("schedule"),))


class Score_Model:
    def __init__(self, state_dict, hparams, precision="fp16"):
        validate_hparams(hparams)
        self.hparams = hparams
        self.state_dict_cpu = {k: v.detach().float().contiguous() for k, v in state_dict.items()}
        self.so3_diffuser = SO3Diffuser(hparams["diffuser"]["so3"])
        self.r3_diffuser = R3Diffuser(hparams["diffuser"]["r3"])
        self.cut_off = float(hparams["model"].get("cut_off", 20.0))
        self.pos_width = int(self.state_dict_cpu["positional_embed.weight"].shape[1])
        self.precision = precision          # "fp16" (tcgen05, default) or "fp32" (FFMA parity mode)
        self.graph_generic = False          # True: always build the graph with the generic kernel (the N > 1024 path)
        self.last_fused = False             # True: last layer + coordinate head as one launch (DFM_LAST_FUSED, test knob)
        self.edge_rng = "torch"             # forward(batch): "torch" = reference-style global RNG, "philox" = in-kernel
        self.device = None
        self._ctx = None
        self._ws = {}
        self._complex_key = None
        self._complex = None
        self._fwd_counter = 0
        self.training = False

    # ---- construction -------------------------------------------------------------------------------
    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, **kwargs):
        sd, hp = load_checkpoint(str(checkpoint_path), map_location="cpu")
        model = cls(sd, hp, **kwargs)
        if map_location is not None and torch.device(map_location).type == "cuda":
            model.to(map_location)
        return model

    def state_dict(self):
        return {"net." + k: v for k, v in self.state_dict_cpu.items()}

    def eval(self):
        return self

    def train(self, mode=True):
        if mode:
            raise NotImplementedError("dfmdock_b200 is inference-only (training is out of scope, SURVEY.md section 8)")
        return self

    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("dfmdock_b200.Score_Model has no CPU path; use a CUDA (sm_100a) device")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if self._ctx is not None and device == self.device:
            return self
        self._free()
        lib = _lib.load()
        ctx = c_void_p()
        _lib.check(lib.dfm_create(ctypes.byref(ctx), device.index), "dfm_create")
        self._ctx, self.device = ctx, device
        with torch.cuda.device(device):
            for name, t in self.state_dict_cpu.items():
                if not name.startswith(HOT_PREFIXES):
                    continue      # to_ires.* is dead at inference
                d = t.to(device)
                shape = (c_int64 * max(d.dim(), 1))(*d.shape)
                _lib.check(lib.dfm_set_weight(ctx, name.encode(), _lib.ptr(d), shape, d.dim()), "dfm_set_weight(%s)" % name)
            _lib.check(lib.dfm_finalize_weights(ctx, self.cut_off, self._stream()), "dfm_finalize_weights")
        # dfm_sample computes g(t) on the host from the checkpoint's own sigmas (hyper_parameters.diffuser)
        _lib.check(lib.dfm_set_schedule(ctx, float(self.so3_diffuser.min_sigma), float(self.so3_diffuser.max_sigma),
                                        float(self.r3_diffuser.min_sigma), float(self.r3_diffuser.max_sigma)), "dfm_set_schedule")
        return self

    def cuda(self, index=None):
        return self.to(torch.device("cuda", index if index is not None else torch.cuda.current_device()))

    def _free(self):
        if self._ctx is not None:
            _lib.load().dfm_destroy(self._ctx)
            self._ctx = None
            self._ws = {}
            self._complex_key = None

    def __del__(self):
        try:
            self._free()
        except Exception:
            pass

    def _stream(self):
        return c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _need_ctx(self):
        if self._ctx is None:
            raise RuntimeError("Score_Model is not on a CUDA device yet: call .to('cuda')")

    # ---- per-complex state ---------------------------------------------------------------------------
    def set_complex(self, batch, check_position_matrix=True):
        """Pose-invariant per-complex work (single_embed, relpos).  `batch` as built by get_batch_from_inputs."""
        self._need_ctx()
        rec_x, lig_x, rec_pos = batch["rec_x"], batch["lig_x"], batch["rec_pos"]
        key = (rec_x.data_ptr(), lig_x.data_ptr(), tuple(rec_x.shape), tuple(lig_x.shape))
        lib = _lib.load()
        dev = self.device
        rp = rec_pos.to(dev, torch.float32).contiguous()
        if key == self._complex_key:
            _lib.check(lib.dfm_set_receptor_pose(self._ctx, _lib.ptr(rp), self._stream()), "dfm_set_receptor_pose")
            return
        rx = rec_x.to(dev, torch.float32).contiguous()
        lx = lig_x.to(dev, torch.float32).contiguous()
        R, L = rx.shape[0], lx.shape[0]
        sym = float(batch.get("sym", 0.0)) if self.pos_width == 67 else 0.0
        pm = batch.get("position_matrix")
        if pm is not None:
            if pm.shape[-1] == 67:
                sym = float(pm[0, 0, 66])
            if check_position_matrix:
                # the kernels rebuild relpos from (i, j, R); make sure that is what the caller's one-hot encodes
                idx = torch.arange(R + L, device=pm.device)
                same = (idx[:, None] < R) == (idx[None, :] < R)
                want = torch.where(same, (idx[:, None] - idx[None, :] + 32).clamp(0, 64), torch.full_like(same, 65, dtype=torch.long))
                if not torch.equal(pm[..., :66].argmax(-1), want):
                    raise ValueError("position_matrix is not relpos(arange(N), chain id): unsupported by the fused pair-feature kernel")
        with torch.cuda.device(dev):
            _lib.check(lib.dfm_set_complex(self._ctx, R, L, rx.shape[1], _lib.ptr(rx), _lib.ptr(lx), _lib.ptr(rp), sym,
                                           self._stream()), "dfm_set_complex")
        self._complex_key = key
        self._complex = (R, L)
        self._keepalive = (rec_x, lig_x)   # keeps data_ptr-based key valid
        self._ws = {}

    def _workspace(self, B):
        ws = self._ws.get(B)
        if ws is None:
            n = _lib.load().dfm_workspace_bytes(self._ctx, B)
            ws = torch.empty(n + 256, dtype=torch.uint8, device=self.device)
            self._ws = {B: ws}     # keep one size resident
        off = (-ws.data_ptr()) % 256
        return c_void_p(ws.data_ptr() + off), ws.numel() - off

    @property
    def edges_per_node(self):
        return _lib.load().dfm_edges_per_node(self._ctx)

    def _flags(self, want_energy=False, precision=None, **kw):
        f = 0
        if want_energy:
            f |= _lib.WANT_ENERGY
        if (precision or self.precision) == "fp32":
            f |= _lib.PRECISION_FP32
        if kw.get("use_clash_force"):
            f |= _lib.CLASH_FORCE
        if kw.get("noise_annealing"):
            f |= _lib.NOISE_ANNEAL
        if kw.get("centre_mode", 0) == 1:
            f |= _lib.CENTRE_ALL_ATOMS
        if kw.get("ode"):
            f |= _lib.ODE
        if kw.get("graph_generic") or self.graph_generic:
            f |= _lib.GRAPH_GENERIC
        if kw.get("last_fused") or self.last_fused:
            f |= _lib.LAST_FUSED
        return f

    # ---- batched operators ---------------------------------------------------------------------------
    def score(self, lig_pos, t, edges=None, exp_noise=None, want_energy=False, seed=0, stream_base=0, forward_index=0,
              precision=None, return_edges=False):
        """Score-network forward for B ligand poses of the current complex -> dict of [B, ...] tensors."""
        self._need_ctx()
        R, L = self._complex
        dev = self.device
        lig_pos = lig_pos.to(dev, torch.float32).contiguous().view(-1, L, 3, 3)
        B = lig_pos.shape[0]
        t = t.to(dev, torch.float32).contiguous().view(-1)
        if t.numel() == 1 and B > 1:
            t = t.expand(B).contiguous()
        K = self.edges_per_node
        if edges is not None:
            edges = edges.to(dev, torch.int32).contiguous().view(B, R + L, K)
        if exp_noise is not None:
            exp_noise = exp_noise.to(dev, torch.float32).contiguous().view(B, R + L, -1)
        out = {
            "tr_score": torch.empty(B, 3, device=dev), "rot_score": torch.empty(B, 3, device=dev),
            "f": torch.empty(B, L, 3, device=dev),
        }
        if want_energy:
            out["energy"] = torch.empty(B, device=dev)
            out["num_clashes"] = torch.empty(B, dtype=torch.int32, device=dev)
        if return_edges:
            out["edges"] = torch.empty(B, R + L, K, dtype=torch.int32, device=dev)
        ws, nws = self._workspace(B)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().dfm_score_forward(
                self._ctx, B, _lib.ptr(lig_pos), _lib.ptr(t), _lib.ptr(edges), _lib.ptr(exp_noise), seed, stream_base,
                forward_index, self._flags(want_energy, precision), _lib.ptr(out["tr_score"]), _lib.ptr(out["rot_score"]),
                _lib.ptr(out["f"]), _lib.ptr(out.get("energy")), _lib.ptr(out.get("num_clashes")), _lib.ptr(out.get("edges")),
                ws, nws, self._stream()), "dfm_score_forward")
        return out

    def interface_logits(self, B):
        """ires = to_ires(h) [B, N] of the last score(..., want_energy=True) call with the same B (score_net_mlsb.py:383)."""
        self._need_ctx()
        R, L = self._complex
        out = torch.empty(B, R + L, device=self.device)
        ws, nws = self._workspace(B)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().dfm_interface_logits(self._ctx, B, _lib.ptr(out), ws, nws, self._stream()), "dfm_interface_logits")
        return out

    def debug_read(self, B, which, shape, dtype=torch.float32):
        out = torch.empty(shape, dtype=dtype, device=self.device)
        ws, _ = self._workspace(B)
        n = _lib.load().dfm_debug_read(self._ctx, B, which, _lib.ptr(out), out.numel() * 4, ws, self._stream())
        if n < 0:
            _lib.check(int(n), "dfm_debug_read")
        return out

    def randomize_pose(self, lig_pos0, B, rot0=None, tr0=None, seed=0, stream_base=0, centre_mode=0):
        self._need_ctx()
        R, L = self._complex
        dev = self.device
        lig0 = lig_pos0.to(dev, torch.float32).contiguous()
        lig = torch.empty(B, L, 3, 3, device=dev)
        rot_u = torch.empty(B, 3, device=dev)
        tr_u = torch.empty(B, 3, device=dev)
        if rot0 is not None:
            rot0 = rot0.to(dev, torch.float32).contiguous().view(B, 3, 3)
        if tr0 is not None:
            tr0 = tr0.to(dev, torch.float32).contiguous().view(B, 3)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().dfm_randomize_pose(self._ctx, B, _lib.ptr(lig0), _lib.ptr(rot0), _lib.ptr(tr0), seed,
                                                      stream_base, self._flags(centre_mode=centre_mode), _lib.ptr(lig),
                                                      _lib.ptr(rot_u), _lib.ptr(tr_u), self._stream()), "dfm_randomize_pose")
        return lig, tr_u, rot_u

    def reverse_step(self, lig_pos, rot_update, tr_update, tr_score, rot_score, t, dt, ns_rot, ns_tr, z=None, seed=0,
                     stream_base=0, step_index=0, use_clash_force=False, centre_mode=0, ode=False):
        """In-place Euler-Maruyama update of B poses (lig_pos [B,L,3,3], rot_update [B,3], tr_update [B,3])."""
        self._need_ctx()
        B = lig_pos.shape[0]
        if z is not None:
            z = z.to(self.device, torch.float32).contiguous().view(B, 2, 3)
        g_rot = float(self.so3_diffuser.diffusion_coef(float(t)))
        g_tr = float(self.r3_diffuser.diffusion_coef(float(t)))
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().dfm_reverse_step(
                self._ctx, B, _lib.ptr(lig_pos), _lib.ptr(rot_update), _lib.ptr(tr_update), _lib.ptr(tr_score),
                _lib.ptr(rot_score), g_rot, g_tr, float(dt), float(ns_rot), float(ns_tr), _lib.ptr(z), seed, stream_base,
                step_index, self._flags(use_clash_force=use_clash_force, centre_mode=centre_mode, ode=ode), self._stream()),
                "dfm_reverse_step")

    def sample(self, lig_pos0, num_traj, num_steps=40, eps=1e-3, tr_noise_scale=0.5, rot_noise_scale=0.5,
               use_clash_force=False, noise_annealing=False, centre_mode=0, seed=0, stream_base=0, precision=None,
               ode=False, record=False):
        """num_traj independent reverse-diffusion trajectories of the current complex, in lock step, on this GPU.

        Equivalent of the reference's serial driver loop around Euler_Maruyama_sampler (inference_base.py:644-657);
        trajectory k draws from Philox subsequence (seed, stream_base + k).  ode=True takes the probability-flow branch of
        torch_reverse (so3_diffuser.py:366-367; Sampler.Euler_Maruyama_sampler(ode=...), inference_mlsb.py:264-350).
        record=True additionally returns "frames" [num_steps + 1, B, L, 3, 3]: the initial pose and the pose after every
        step (rec_trj / lig_trj of inference_mlsb.py:273-348); it runs the same kernels step by step from the host.
        """
        if record:
            return self._sample_recorded(lig_pos0, num_traj, num_steps, eps, tr_noise_scale, rot_noise_scale, use_clash_force,
                                         noise_annealing, centre_mode, seed, stream_base, precision, ode)
        self._need_ctx()
        R, L = self._complex
        dev = self.device
        B = int(num_traj)
        lig0 = lig_pos0.to(dev, torch.float32).contiguous()
        out = {
            "lig_pos": torch.empty(B, L, 3, 3, device=dev), "rot_update": torch.empty(B, 3, device=dev),
            "tr_update": torch.empty(B, 3, device=dev), "energy": torch.empty(B, device=dev),
            "num_clashes": torch.empty(B, dtype=torch.int32, device=dev),
        }
        ws, nws = self._workspace(B)
        flags = self._flags(False, precision, use_clash_force=use_clash_force, noise_annealing=noise_annealing,
                            centre_mode=centre_mode, ode=ode)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().dfm_sample(
                self._ctx, B, _lib.ptr(lig0), int(num_steps), float(eps), float(tr_noise_scale), float(rot_noise_scale),
                flags, seed, stream_base, _lib.ptr(out["lig_pos"]), _lib.ptr(out["rot_update"]), _lib.ptr(out["tr_update"]),
                _lib.ptr(out["energy"]), _lib.ptr(out["num_clashes"]), ws, nws, self._stream()), "dfm_sample")
        return out

    def _sample_recorded(self, lig_pos0, num_traj, num_steps, eps, tr_noise_scale, rot_noise_scale, use_clash_force,
                         noise_annealing, centre_mode, seed, stream_base, precision, ode):
        """dfm_sample unrolled on the host (same kernels, same Philox keys -> same poses) with every frame kept."""
        if num_steps < 2:
            raise ValueError("num_steps must be >= 2")
        B = int(num_traj)
        lig, tr_u, rot_u = self.randomize_pose(lig_pos0, B, seed=seed, stream_base=stream_base, centre_mode=centre_mode)
        frames = [lig.clone()]
        ts = torch.linspace(1.0, eps, num_steps)            # fp32, like the reference (inference_base.py:404)
        dt = float(ts[0] - ts[1])
        t_dev = torch.empty(B, device=self.device)
        o = None
        for i in range(num_steps):
            t = float(ts[i])
            last = i == num_steps - 1
            t_dev.fill_(t)
            o = self.score(lig, t_dev, seed=seed, stream_base=stream_base, forward_index=i, precision=precision)
            if noise_annealing:
                ns_tr = ns_rot = t
            elif last:
                ns_tr = ns_rot = 0.0
            else:
                ns_tr, ns_rot = tr_noise_scale, rot_noise_scale
            self.reverse_step(lig, rot_u, tr_u, o["tr_score"], o["rot_score"], t, dt, ns_rot, ns_tr, seed=seed,
                              stream_base=stream_base, step_index=i, use_clash_force=use_clash_force,
                              centre_mode=centre_mode, ode=ode)
            frames.append(lig.clone())
            if last:
                o = self.score(lig, t_dev, seed=seed, stream_base=stream_base, forward_index=num_steps, precision=precision,
                               want_energy=True)
        return {"lig_pos": lig, "rot_update": rot_u, "tr_update": tr_u, "energy": o["energy"], "num_clashes": o["num_clashes"],
                "frames": torch.stack(frames, dim=0)}

    def gt_energy(self, batch):
        """Energy / clash count of the pose in `batch` itself at t = 1e-5 (Sampler.run_sampling with get_gt_energy,
        src/inference_mlsb.py:190-199) -> (energy float, num_clashes int)."""
        b = dict(batch)
        b["t"] = torch.zeros(1) + 1e-5
        o = self.forward(b)
        return float(o["energy"]), int(o["num_clashes"])

    def profile_enable(self, max_launches):
        _lib.check(_lib.load().dfm_profile_enable(self._ctx, int(max_launches)), "dfm_profile_enable")

    def profile_read(self):
        """-> (summed edge-kernel milliseconds, launches) since the last read; synchronises on the recorded events."""
        ms, n = ctypes.c_double(), ctypes.c_int()
        _lib.check(_lib.load().dfm_profile_read(self._ctx, ctypes.byref(ms), ctypes.byref(n)), "dfm_profile_read")
        return ms.value, n.value

    @property
    def launch_count(self):
        return int(_lib.load().dfm_launch_count(self._ctx)) if self._ctx is not None else 0

    # ---- reference-compatible single-pose forward ------------------------------------------------------
    def forward(self, batch):
        """Score_Model.forward(batch) (src/models/score_model_mlsb.py:61-63): one pose, reference output dict."""
        self.set_complex(batch)
        R, L = self._complex
        N = R + L
        lig_pos = batch["lig_pos"].to(self.device, torch.float32).view(1, L, 3, 3)
        exp_noise = None
        K = self.edges_per_node
        if self.edge_rng == "torch" and K == 60:
            # same draw torch.multinomial(replacement=False) makes inside the reference (score_net_mlsb.py:130)
            exp_noise = torch.empty(1, N, N - 20, device=self.device).exponential_(1)
        self._fwd_counter += 1
        o = self.score(lig_pos, batch["t"], exp_noise=exp_noise, want_energy=True, forward_index=self._fwd_counter)
        return {
            "tr_score": o["tr_score"].view(1, 3),
            "rot_score": o["rot_score"].view(1, 3),
            "energy": o["energy"][0],
            "f": o["f"][0],
            "num_clashes": o["num_clashes"][0].long(),
            # to_ires (score_net_mlsb.py:383): never read at inference (inference_base.py:494-500), computed here only so
            # that the returned dict is the reference's; checkpoints stripped of to_ires.* get NaN so misuse is visible
            "ires": (self.interface_logits(1).view(N, 1) if "to_ires.0.weight" in self.state_dict_cpu
                     else torch.full((N, 1), float("nan"), device=self.device)),
        }

    __call__ = forward
