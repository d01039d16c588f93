#!/usr/bin/env python
"""bench.py -- docked poses/sec of the reverse-diffusion sampler hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # the sm_100a CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host CPU cores (oracle port)

Workload (config #3 of BASELINE.json, the one the metric is quoted on): synthetic 2x150-residue complex, 256
trajectories per GPU advancing in lock step, weights/pinder_0.ckpt (oracle/_ref/pinder_0.pt travels with the repo
snapshot; seeded random weights of the same architecture only if it is absent -- config.weights says which ran).
A "step" = one reverse-diffusion step of all 256 trajectories (score-network forward + SO(3)xR^3 Euler-Maruyama
update) = 256 pose-steps, timed on poses IN CONTACT in the second half of the schedule (t <= 0.5: every inter-chain
pair inside 22 A takes the angle-table gathers -- the expensive regime; far starts are cheaper and are part of
full_job).  value = pose-steps/s over all GPUs.  Default scaling is weak (256 trajectories per GPU);
--scaling strong splits 256 trajectories over the GPUs, and every multi-GPU line carries the strong
figure next to the weak one under "strong_scaling".

Keys beyond the base contract:
  e2e          same metric through the public Python API with HOST buffers (pinned H2D of poses/times, D2H of poses/scores per step)
  roofline     dominant kernel (warp-specialised fused tcgen05 edge kernel, edge_ws.cu): algorithmic FLOP per launch / CUDA-event kernel time, vs the
               measured dense fp16/bf16 tensor peak in MEASURED_PEAKS.json (sustained figure: kernel timed inside a long step)
  cpu_baseline oracle port (the reference's algorithm, torch CPU ops as the reference writes them) on a bounded sample
  full_job     one complete config-#3 job (256 trajectories x 100 steps incl. random init and the final energy forward)
  other_configs  (N=1) complete jobs of BASELINE configs #2 (1QA9, 40 x 40, pinder_0, clash force) and #4 (2x400, 64 x 100)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_REC, N_LIG = 150, 150
TRAJ_PER_GPU = 256
FULL_STEPS = 100
H = 256
METRIC = "docked poses/sec (traj x steps/s), synthetic 2x150-residue complex"
UNIT = "pose-steps/s"


def algorithmic_flops_per_pose_step(n):
    """depth*[2EH^2 + 2EH + 4NH^2 + 6NH^2] + [2EH^2 + 2EH], E = 60 N."""
    e = 60 * n
    return 6 * (2 * e * H * H + 2 * e * H + 4 * n * H * H + 6 * n * H * H) + (2 * e * H * H + 2 * e * H)


def edge_kernel_flops_per_launch(b, n):
    e = 60 * n * b
    return 2 * e * H * H + 2 * e * H


def last_layer_row_fraction(n, r):
    """Share of the node-pair tiles the LAST edge launch of a forward walks when no energy is wanted: only tiles that hold
    a ligand residue (csrc/edge_ws.cu, Params::lig_only) -- the receptor rows' layer-5 messages feed a node update that
    is never run (the reference computes and discards it)."""
    a = ((n - 1) >> 1) - (r >> 1) + 1
    b = (n >> 1) - ((r + 1) >> 1) + 1
    tpt = max(a, b) if n & 1 else a
    return min(1.0, 2.0 * tpt / n)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"tflops": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))), "source": "measured (MEASURED_PEAKS.json, sustained)"}
    return {"tflops": 1400.0, "source": "fallback (B200_PROFILING.md, sustained)"}


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: NVML every 20 ms (nvidia-smi every 200 ms if NVML
    cannot be opened)."""
    FIELDS = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index, uuid=None):
        self.index, self.uuid, self.samples, self.stop, self.source = index, uuid, [], False, "nvidia-smi"
        self.thread = threading.Thread(target=self.run, daemon=True)

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        if self.uuid:
            try:
                return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + self.uuid).encode())
            except Exception:
                pass
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = self.index
        if vis:
            try:
                idx = int(vis.split(",")[self.index])
            except Exception:
                pass
        return pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)

    def run(self):
        try:
            nv, h = self._nvml_handle()
            mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            bits = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                    ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
            self.source = "nvml"
            while not self.stop:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                self.samples.append([str(sm), str(mx)] + ["Active" if r & b else "Not Active" for _, b in bits])
                time.sleep(0.02)
            return
        except Exception:
            self.source = "nvidia-smi"
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.thread.start()
        time.sleep(0.05)       # let the sampler open NVML before the timed region starts
        return self

    def __exit__(self, *a):
        self.stop = True
        self.thread.join(timeout=6)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples), "source": self.source}


_JSON_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on the first
    communicator), so everything that is not the result line is sent to stderr: fd 1 is re-pointed at fd 2 and the line
    is written to a private duplicate of the original stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def load_weights():
    """-> (state_dict, hparams, positional width, description).  The shipped weights/pinder_0.ckpt when oracle/_ref holds
    its re-serialised copy (oracle/build_ref.py; git-ignored, travels with the gpurun snapshot), else seeded random weights."""
    import torch
    from dfmdock_b200.synthetic import synthetic_hparams, synthetic_state_dict
    p = os.path.join(REF_DIR, "pinder_0.pt")
    if os.path.exists(p) and not os.environ.get("DFM_BENCH_SYNTHETIC_WEIGHTS"):
        ck = torch.load(p, weights_only=False)
        return ck["state_dict"], ck["hparams"], 67, "weights/pinder_0.ckpt (trained, 3.57 M parameters)"
    return synthetic_state_dict(0, 66), synthetic_hparams(66), 66, "seeded random weights of the shipped architecture (oracle/_ref/pinder_0.pt absent)"


def make_workload(n_rec=N_REC, n_lig=N_LIG):
    from dfmdock_b200.features import synthetic_complex
    sd, hp, width, desc = load_weights()
    batch = synthetic_complex(n_rec, n_lig, seed=0, pos_width=width)
    return sd, hp, batch, desc


def contact_poses(lig0, n, seed=0, max_angle=0.6, tr_std=3.0):
    """n rigid copies of the generator's ligand pose (chains in contact, SURVEY 8d: +25 A along x): rotation about the CA
    centroid by a random axis-angle (<= max_angle rad) + N(0, tr_std^2) translation per trajectory."""
    import torch
    g = torch.Generator().manual_seed(seed)
    aa = torch.randn(n, 3, generator=g)
    ang = torch.rand(n, 1, generator=g) * max_angle
    ax = aa / aa.norm(dim=-1, keepdim=True)
    K = torch.zeros(n, 3, 3)
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0], K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -ax[:, 2], ax[:, 1], ax[:, 2], -ax[:, 0], -ax[:, 1], ax[:, 0]
    Rm = torch.eye(3)[None] + torch.sin(ang)[:, :, None] * K + (1 - torch.cos(ang))[:, :, None] * (K @ K)   # Rodrigues
    c = lig0[:, 1].mean(0)
    tr = torch.randn(n, 1, 1, 3, generator=g) * tr_std
    return torch.einsum("lac,ndc->nlad", lig0 - c, Rm) + c + tr


def cpu_reference_throughput(num_traj, num_steps, threads=None):
    """The reference's algorithm on the host cores: oracle port, serial trajectories at batch 1 like the reference."""
    import torch
    from oracle import dfmdock_oracle as orc
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core it can
    torch.set_num_threads(threads or os.cpu_count() or 1)
    sd, hp, batch, _ = make_workload()
    net = orc.OracleNet(sd, cut_off=hp["model"]["cut_off"])
    import numpy as np
    np.random.seed(0)
    torch.manual_seed(0)
    with torch.no_grad():
        orc.euler_maruyama_sampler(net, batch, num_steps=2)        # warm-up (thread pools, allocator)
        t0 = time.perf_counter()
        for _ in range(num_traj):
            orc.euler_maruyama_sampler(net, batch, num_steps=num_steps)
        dt = time.perf_counter() - t0
    return num_traj * num_steps / dt, dt, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(2, min(args.steps, 10))
    traj = 2
    value, wall, cores = cpu_reference_throughput(traj, steps)
    desc = load_weights()[3]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "synthetic 2x150-residue complex, reference algorithm on host CPU (oracle port), serial trajectories at batch 1",
                   "n_res": N_REC + N_LIG, "weights": desc,
                   "sample": "%d trajectories x %d steps (+ final forward each)" % (traj, steps)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d trajectories x %d reverse steps, N=300, %.1f s wall" % (traj, steps, wall)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


DTYPE = ("f16: tcgen05 operands, inter-kernel activations and the packed-half2 SIMT math of the edge kernel (both SiLUs, gate "
         "logit partial sums, gated 60-edge segment sum) are fp16; MMA accumulators, GraphNorm statistics, the residual "
         "stream h and all geometry are f32")
TOLERANCE = "1e-2 relative (L2) on forces / scores vs the fp32 oracle at this configuration (tests/test_gpu_configs.py; measured worst printed there)"


def time_full_job(model, batch, T, S, seed, stream_base=0, **kw):
    """One complete dfm_sample job with CUDA events -> (milliseconds, result dict)."""
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = model.sample(batch["lig_pos"], T, num_steps=S, seed=seed, stream_base=stream_base, **kw)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), res


def other_configs(dev, sd, hp):
    """BASELINE configs #2 and #4 as complete jobs on one GPU (parity-test sizes, reported beside the bench line)."""
    import torch
    from dfmdock_b200 import Score_Model
    from dfmdock_b200.features import batch_from_record, synthetic_complex
    out = {}
    model = Score_Model(sd, hp, precision="fp16").to(dev)
    rec_path = os.path.join(REF_DIR, "db5_1QA9.pt")
    if os.path.exists(rec_path):
        batch = batch_from_record(torch.load(rec_path, weights_only=False), pos_width=model.pos_width)
        model.set_complex(batch)
        kw = dict(use_clash_force=True, centre_mode=1)
        time_full_job(model, batch, 40, 3, 1, **kw)
        ms, res = time_full_job(model, batch, 40, 40, 2, **kw)
        out["c2"] = {"workload": "db5 1QA9 (N=197), 40 trajectories x 40 steps, clash force (src/inference.py defaults)",
                     "wall_ms": ms, "poses_per_s": 40 * 40 / (ms * 1e-3), "us_per_lockstep_step": ms * 1e3 / 41,
                     "best_energy": float(res["energy"].min())}
    batch = synthetic_complex(400, 400, seed=0, pos_width=model.pos_width)
    model.set_complex(batch)
    time_full_job(model, batch, 64, 3, 1)
    ms, res = time_full_job(model, batch, 64, 100, 2)
    out["c4"] = {"workload": "synthetic 2x400 residues, 64 trajectories x 100 steps", "wall_ms": ms,
                 "poses_per_s": 64 * 100 / (ms * 1e-3), "best_energy": float(res["energy"].min())}
    return out


def run_cuda(args):
    import torch
    import torch.distributed as dist
    from dfmdock_b200 import Score_Model

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    sd, hp, batch, wdesc = make_workload()
    model = Score_Model(sd, hp, precision="fp16").to(dev)
    model.set_complex(batch)
    L, N = N_LIG, N_REC + N_LIG
    S = FULL_STEPS
    ts = torch.linspace(1.0, 1e-3, S)
    dt = float(ts[0] - ts[1])
    half = S // 2

    try:
        gpu_uuid = str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:
        gpu_uuid = None

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def make_state(B, base):
        """B trajectories of this rank in contact (global trajectory index base + b -> pose seed and Philox subsequence)."""
        lig = contact_poses(batch["lig_pos"], B, seed=1000 + base).to(dev)
        return {"B": B, "base": base, "lig": lig, "tr_u": torch.zeros(B, 3, device=dev), "rot_u": torch.zeros(B, 3, device=dev),
                "t": torch.empty(B, device=dev), "i": 0}

    def one_step(st):
        i = half + st["i"] % (S - 1 - half)      # second half of the schedule, never the noise-free last step
        t = float(ts[i])
        st["t"].fill_(t)
        o = model.score(st["lig"], st["t"], seed=args.seed, stream_base=st["base"], forward_index=st["i"])
        model.reverse_step(st["lig"], st["rot_u"], st["tr_u"], o["tr_score"], o["rot_score"], t, dt, 0.5, 0.5, seed=args.seed,
                           stream_base=st["base"], step_index=st["i"])
        st["i"] += 1

    def timed_region(st, steps, profile):
        """steps lock-step steps + the path's only collective, CUDA events, max over ranks -> (ms, launches, clocks, edge ms, edge n)."""
        for _ in range(max(args.warmup, 3)):
            one_step(st)
        sync_all()
        if profile:
            model.profile_enable(steps * 8)
        launches0 = model.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local_rank, gpu_uuid) as clocks:
            sync_all()
            e0.record()
            for _ in range(steps):
                one_step(st)
            if world > 1:
                table = torch.cat([st["rot_u"], st["tr_u"], torch.zeros(st["B"], 2, device=dev)], dim=1)
                full = torch.empty(world * st["B"], 8, device=dev)
                dist.all_gather_into_tensor(full, table)         # the path's only collective: [T_local, 8] result rows
            e1.record()
            sync_all()
        ms = e0.elapsed_time(e1)
        launches = model.launch_count - launches0
        edge_ms, edge_n = (0.0, 0)
        if profile:
            edge_ms, edge_n = model.profile_read()
            model.profile_enable(0)
        tmax = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        return float(tmax.item()), launches, clocks.summary(), edge_ms, edge_n

    strong = args.scaling == "strong"
    B_weak = TRAJ_PER_GPU
    lo_s = rank * TRAJ_PER_GPU // world
    B_strong = (rank + 1) * TRAJ_PER_GPU // world - lo_s
    # ---- primary timed region (device resident) ------------------------------------------------------
    B = B_strong if strong else B_weak
    base = lo_s if strong else rank * B_weak
    st = make_state(B, base)
    elapsed_ms, launches, clk, edge_ms, edge_n = timed_region(st, args.steps, True)
    total_traj = TRAJ_PER_GPU if strong else world * B_weak
    value = total_traj * args.steps / (elapsed_ms * 1e-3)
    # ---- the other scaling mode beside it (multi-GPU lines only; at N = 1 the two coincide) ----------------
    strong_line = {"trajectories_total": TRAJ_PER_GPU, "trajectories_per_gpu": B_strong, "value": value if (strong or world == 1) else None,
                   "unit": UNIT, "ms_per_step": elapsed_ms / args.steps if (strong or world == 1) else None}
    if world > 1 and not strong:
        st2 = make_state(B_strong, lo_s)
        ms2, _, _, _, _ = timed_region(st2, args.steps, False)
        strong_line.update(value=TRAJ_PER_GPU * args.steps / (ms2 * 1e-3), ms_per_step=ms2 / args.steps)
        del st2

    # ---- end-to-end through the public API with host buffers -----------------------------------------
    lig, tr_u, rot_u = st["lig"], st["tr_u"], st["rot_u"]
    # two pinned pose buffers used in turn: the pose a step returns to the host is the pose the next step sends to the device
    pose_host = [torch.empty(B, L, 3, 3).pin_memory(), torch.empty(B, L, 3, 3).pin_memory()]
    t_host = torch.empty(B).pin_memory()
    sc_host = torch.empty(B, 6).pin_memory()
    pose_host[1].copy_(lig.cpu())
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_step(k):
        t = float(ts[half + k % (S - 1 - half)])
        t_host.fill_(t)
        src, dst = pose_host[(k + 1) & 1], pose_host[k & 1]
        lig_d = src.to(dev, non_blocking=True)
        t_d = t_host.to(dev, non_blocking=True)
        o = model.score(lig_d, t_d, seed=args.seed, stream_base=base, forward_index=1000 + k)
        model.reverse_step(lig_d, rot_u, tr_u, o["tr_score"], o["rot_score"], t, dt, 0.5, 0.5, seed=args.seed,
                           stream_base=base, step_index=1000 + k)
        dst.copy_(lig_d, non_blocking=True)
        sc_host.copy_(torch.cat([o["tr_score"], o["rot_score"]], dim=1), non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()      # the caller reads the result of every step

    e2e_step(0)
    sync_all()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        e2e_step(1 + k)
    sync_all()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = total_traj * e2e_steps / float(te.item())

    # ---- one complete config-#3 job (random init far apart, 100 steps, final energy forward) ----------------
    full_job = None
    if not args.no_full_job:
        sync_all()
        ms, res = time_full_job(model, batch, B, FULL_STEPS, args.seed, stream_base=base)
        fj = torch.tensor([ms], device=dev)
        emin = res["energy"].min().reshape(1)
        if world > 1:
            dist.all_reduce(fj, op=dist.ReduceOp.MAX)
            dist.all_reduce(emin, op=dist.ReduceOp.MIN)
        full_job = {"trajectories": total_traj, "steps": FULL_STEPS, "wall_s": float(fj.item()) * 1e-3,
                    "poses_per_s": total_traj * FULL_STEPS / (float(fj.item()) * 1e-3),
                    "best_energy": float(emin.item())}

    others = None
    if world == 1 and not args.no_other_configs:
        others = other_configs(dev, sd, hp)

    if rank == 0:
        peaks = measured_peaks()
        # timed: the five launches of a forward that walk every edge (the sixth walks only the ligand residues' tiles and,
        # fused with the coordinate head, is not an edge-kernel sample)
        executed = 1.0
        flops_launch = edge_kernel_flops_per_launch(B, N) * executed
        edge_avg_ms = edge_ms / max(edge_n, 1)
        achieved = flops_launch / (edge_avg_ms * 1e-3) / 1e12 if edge_n else None
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, wall, cores = cpu_reference_throughput(2, 10)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "oracle port, 2 trajectories x 10 reverse steps of the same 2x150 workload (%.1f s)" % wall}
        traffic = None
        tp = os.path.join(ROOT, "profiles", "edge_kernel_traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        step_flops = algorithmic_flops_per_pose_step(N) * B
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": "synthetic 2x150-residue complex (BASELINE config #3), %d trajectories %s in lock step, poses in contact, "
                                   "second half of the 100-step schedule" % (TRAJ_PER_GPU, "in total" if strong else "per GPU"),
                       "n_res": N, "trajectories_per_gpu": B, "edges_per_step_per_gpu": 60 * N * B,
                       "weights": wdesc, "dtype_detail": DTYPE, "tolerance": TOLERANCE,
                       "l2": "per-step working set ~%.1f GB per GPU, far larger than the 126 MB L2; no explicit flush" % (1.9 * B / 256.0),
                       "parallelism": "trajectory-sharded x%d, one all-gather of [T,8] at the end" % world},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * L * 36 + B * 4, "d2h_bytes_per_step": B * L * 36 + B * 24,
                    "steps": e2e_steps},
            "gpu_launches": launches,
            "clocks": clk,
            "roofline": {"bound": "tensor", "kernel": "ews::k_edge_ws (warp-specialised fused edge MLP, tcgen05)", "achieved": achieved,
                         "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": (achieved / peaks["tflops"]) if achieved else None,
                         "traffic": traffic, "peak_source": peaks["source"], "kernel_ms_per_launch": edge_avg_ms,
                         "launches_timed": edge_n, "flops_per_launch": flops_launch,
                         "flops_note": "the 5 launches of a forward that walk every edge (B x N x 60 edges x (2 x 256^2 + 2 x 256) flop); the "
                                       "last layer's ligand-only launch (fused with the coordinate head) is not in the sample",
                         "kernel_share_of_step": (edge_ms / elapsed_ms) if edge_n else None,
                         "whole_step_tflops": step_flops / (elapsed_ms / args.steps * 1e-3) / 1e12},
            "strong_scaling": strong_line,
            "cpu_baseline": cpu,
            "full_job": full_job,
            "other_configs": others,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-full-job", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 256 trajectories per GPU (default); strong: 256 trajectories split over the GPUs (SURVEY 8e)")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
