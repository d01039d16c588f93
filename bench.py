#!/usr/bin/env python
"""bench.py -- docked poses/sec of the reverse-diffusion sampler hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # the sm_100a CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host CPU cores (oracle port)

Workload (config #3 of BASELINE.json, the one the metric is quoted on): synthetic 2x150-residue complex, 256
trajectories per GPU advancing in lock step, random-init weights of the shipped architecture.  A "step" = one
reverse-diffusion step of all 256 trajectories (score-network forward + SO(3)xR^3 Euler-Maruyama update) =
256 pose-steps.  value = pose-steps/s over all GPUs (weak scaling: 256 trajectories per GPU).

Keys beyond the base contract:
  e2e          same metric through the public Python API with HOST buffers (pinned H2D of poses/times, D2H of poses/scores per step)
  roofline     dominant kernel (warp-specialised fused tcgen05 edge kernel, edge_ws.cu): algorithmic FLOP per launch / CUDA-event kernel time, vs the
               measured dense fp16/bf16 tensor peak in MEASURED_PEAKS.json (sustained figure: kernel timed inside a long step)
  cpu_baseline oracle port (the reference's algorithm, torch CPU ops as the reference writes them) on a bounded sample
  full_job     one complete config-#3 job (256 trajectories x 100 steps incl. random init and the final energy forward)
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


def make_workload():
    from dfmdock_b200.features import synthetic_complex
    from dfmdock_b200.synthetic import synthetic_hparams, synthetic_state_dict
    batch = synthetic_complex(N_REC, N_LIG, seed=0, pos_width=66)
    return synthetic_state_dict(0, 66), synthetic_hparams(66), batch


def cpu_reference_throughput(num_traj, num_steps, threads=None):
    """The reference's algorithm on the host cores: oracle port, serial trajectories at batch 1 like the reference."""
    import torch
    from oracle import dfmdock_oracle as orc
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core it can
    torch.set_num_threads(threads or os.cpu_count() or 1)
    sd, hp, batch = make_workload()
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
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "synthetic 2x150-residue complex, reference algorithm on host CPU (oracle port), serial trajectories at batch 1",
                   "n_res": N_REC + N_LIG, "sample": "%d trajectories x %d steps (+ final forward each)" % (traj, steps)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d trajectories x %d reverse steps, N=300, %.1f s wall" % (traj, steps, wall)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


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

    sd, hp, batch = make_workload()
    model = Score_Model(sd, hp, precision="fp16").to(dev)
    model.set_complex(batch)
    B, L, N = TRAJ_PER_GPU, N_LIG, N_REC + N_LIG
    S = FULL_STEPS
    ts = torch.linspace(1.0, 1e-3, S)
    dt = float(ts[0] - ts[1])
    base = rank * B          # global trajectory index -> Philox subsequence (results independent of the number of GPUs)

    lig, tr_u, rot_u = model.randomize_pose(batch["lig_pos"], B, seed=args.seed, stream_base=base)
    t_dev = torch.empty(B, device=dev)
    state = {"i": 0}

    def one_step():
        i = state["i"] % (S - 1)            # never the noise-free last step: steady-state steps only
        t = float(ts[i])
        t_dev.fill_(t)
        o = model.score(lig, t_dev, seed=args.seed, stream_base=base, forward_index=state["i"])
        model.reverse_step(lig, rot_u, tr_u, o["tr_score"], o["rot_score"], t, dt, 0.5, 0.5, seed=args.seed,
                           stream_base=base, step_index=state["i"])
        state["i"] += 1

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        one_step()
    sync_all()

    # ---- device-resident timed region ------------------------------------------------------------
    try:
        gpu_uuid = str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:
        gpu_uuid = None
    model.profile_enable(args.steps * 8)
    launches0 = model.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank, gpu_uuid) as clocks:
        sync_all()
        e0.record()
        for _ in range(args.steps):
            one_step()
        if world > 1:
            table = torch.cat([rot_u, tr_u, torch.zeros(B, 2, device=dev)], dim=1)
            full = torch.empty(world * B, 8, device=dev)
            dist.all_gather_into_tensor(full, table)         # the path's only collective: [T_local, 8] result rows
        e1.record()
        sync_all()
    elapsed_ms = e0.elapsed_time(e1)
    launches = model.launch_count - launches0
    edge_ms, edge_n = model.profile_read()
    model.profile_enable(0)
    tmax = torch.tensor([elapsed_ms], device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    elapsed_ms = float(tmax.item())
    value = world * B * args.steps / (elapsed_ms * 1e-3)

    # ---- end-to-end through the public API with host buffers -----------------------------------------
    lig_host = torch.empty(B, L, 3, 3).pin_memory()
    t_host = torch.empty(B).pin_memory()
    out_host = torch.empty(B, L, 3, 3).pin_memory()
    sc_host = torch.empty(B, 6).pin_memory()
    lig_host.copy_(lig.cpu())
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_step(k):
        t = float(ts[k % (S - 1)])
        t_host.fill_(t)
        lig_d = lig_host.to(dev, non_blocking=True)
        t_d = t_host.to(dev, non_blocking=True)
        o = model.score(lig_d, t_d, seed=args.seed, stream_base=base, forward_index=1000 + k)
        model.reverse_step(lig_d, rot_u, tr_u, o["tr_score"], o["rot_score"], t, dt, 0.5, 0.5, seed=args.seed,
                           stream_base=base, step_index=1000 + k)
        out_host.copy_(lig_d, non_blocking=True)
        sc_host.copy_(torch.cat([o["tr_score"], o["rot_score"]], dim=1), non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()      # the caller reads the result of every step
        lig_host.copy_(out_host)

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
    e2e_value = world * B * e2e_steps / float(te.item())

    # ---- one complete config-#3 job (256 x 100 with init + final energy forward) ------------------------
    full_job = None
    if not args.no_full_job:
        sync_all()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        res = model.sample(batch["lig_pos"], B, num_steps=FULL_STEPS, seed=args.seed, stream_base=base)
        f1.record()
        sync_all()
        fj = torch.tensor([f0.elapsed_time(f1)], device=dev)
        if world > 1:
            dist.all_reduce(fj, op=dist.ReduceOp.MAX)
        full_job = {"trajectories": world * B, "steps": FULL_STEPS, "wall_s": float(fj.item()) * 1e-3,
                    "poses_per_s": world * B * FULL_STEPS / (float(fj.item()) * 1e-3),
                    "best_energy": float(res["energy"].min().item())}

    if rank == 0:
        peaks = measured_peaks()
        # five launches per forward walk every edge, the sixth only the ligand residues' edges: average work per launch
        executed = (5.0 + last_layer_row_fraction(N, N_REC)) / 6.0
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
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 operands / f32 accumulate", "data": "synthetic",
            "config": {"workload": "synthetic 2x150-residue complex (BASELINE config #3), %d trajectories per GPU in lock step" % B,
                       "n_res": N, "trajectories_per_gpu": B, "edges_per_step_per_gpu": 60 * N * B,
                       "weights": "random-init, shipped architecture (H=256, depth 6); the scores are not physical, so the poses drift "
                                  "apart and the final energy head sees no pair inside its 20 A cut-off (full_job.best_energy = 0)",
                       "l2": "per-step working set ~%.1f GB per GPU, far larger than the 126 MB L2; no explicit flush" % (1.9),
                       "parallelism": "trajectory-sharded x%d, one all-gather of [T,8] at the end" % world},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * L * 36 + B * 4, "d2h_bytes_per_step": B * L * 36 + B * 24,
                    "steps": e2e_steps},
            "gpu_launches": launches,
            "clocks": clocks.summary(),
            "roofline": {"bound": "tensor", "kernel": "ews::k_edge_ws (warp-specialised fused edge MLP, tcgen05)", "achieved": achieved,
                         "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": (achieved / peaks["tflops"]) if achieved else None,
                         "traffic": traffic, "peak_source": peaks["source"], "kernel_ms_per_launch": edge_avg_ms,
                         "launches_timed": edge_n, "flops_per_launch": flops_launch,
                         "flops_note": "average over the 6 launches of a forward of the edges each launch actually processes "
                                       "(5 x all edges + 1 x the ligand residues' edges = %.4f of 6 full launches)" % executed,
                         "kernel_share_of_step": (edge_ms / elapsed_ms) if edge_n else None,
                         "whole_step_tflops": step_flops / (elapsed_ms / args.steps * 1e-3) / 1e12},
            "cpu_baseline": cpu,
            "full_job": full_job,
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
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
