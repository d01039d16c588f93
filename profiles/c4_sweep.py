"""Pair-tile sweep of BASELINE config #4 (synthetic 2x400, 64 trajectories): the graph kernel (O(N^2) pair scan + selection +
6D features of the selected pairs) timed alone with CUDA events for every variant the library has -- register-resident
selection (keys per lane fixed by N: k_graph_sel<W, 32>) with W = 2 / 4 / 8 rows per CTA, and the generic shared-memory
kernel -- plus the complete lock-step step.  One process per variant (the switches are read once)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import torch
    from dfmdock_b200 import Score_Model
    from dfmdock_b200.features import synthetic_complex
    from dfmdock_b200.synthetic import synthetic_hparams, synthetic_state_dict
    n, T = int(sys.argv[2]), int(sys.argv[3])
    sd, hp = synthetic_state_dict(0, 66), synthetic_hparams(66)
    ref = os.path.join(ROOT, "oracle", "_ref", "pinder_0.pt")
    if os.path.exists(ref):
        ck = torch.load(ref, weights_only=False); sd, hp = ck["state_dict"], ck["hparams"]
    model = Score_Model(sd, hp, precision="fp16").to("cuda")
    batch = synthetic_complex(n, n, seed=0, pos_width=model.pos_width)
    model.set_complex(batch)
    lig = batch["lig_pos"][None].repeat(T, 1, 1, 1).cuda().contiguous()
    t = torch.full((T,), 0.3, device="cuda")
    def step(i):
        o = model.score(lig, t, seed=0, forward_index=i)
    for i in range(3): step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10): step(3 + i)
    e1.record(); torch.cuda.synchronize()
    fwd_ms = e0.elapsed_time(e1) / 10
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(5): step(20 + i)
        torch.cuda.synchronize()
    g = [ev for ev in prof.key_averages() if "k_graph" in ev.key]
    g_ms = sum(ev.device_time_total for ev in g) / 5 / 1e3 if g else float("nan")
    print("RESULT %.4f %.4f" % (fwd_ms, g_ms))
    sys.exit(0)
rows = []
for n, T in ((400, 64), (150, 256), (250, 128), (1000, 16)):
    for tag, env in (("sel W=8", {"DFM_GRAPH_WARPS": "8"}), ("sel W=4", {"DFM_GRAPH_WARPS": "4"}), ("sel W=2", {"DFM_GRAPH_WARPS": "2"}), ("default", {}),
                     ("generic (smem, k argmin passes)", {"DFM_GRAPH_KERNEL": "0"})):
        e = dict(os.environ); e.update(env)
        out = subprocess.run([sys.executable, __file__, "child", str(n), str(T)], env=e, capture_output=True, text=True)
        r = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
        if not r:
            print("failed", tag, out.stderr[-300:]); continue
        fwd, g = map(float, r[0].split()[1:])
        print("N=2x%d T=%d  %-34s forward %.3f ms  graph kernel %.3f ms (%.1f %% of the forward)" % (n, T, tag, fwd, g, 100 * g / fwd))
