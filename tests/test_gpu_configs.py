"""GPU: the CUDA path against the oracle AT THE CONFIGURATIONS THE NUMBERS ARE QUOTED ON (BASELINE configs #3 and #4),
with the shipped weights/pinder_0.ckpt (from oracle/_ref) when present, at near-contact poses and at randomize_pose
starts (t = 1.0, chains 50-150 A apart, radial up to 1e5 A^2); the free-running same-noise trajectory (T4); and the
live-reference goldens at far poses.

Tolerances (SURVEY 8c protocol):
  fp32 mode (FFMA kernels):                  1e-4 relative (L2) on f / tr_score / rot_score (measured worst 5.5e-5), 2e-3 absolute on energy
  fp16 mode (fp16 operands + fp16 SIMT, fp32 MMA accumulate): 1e-2 relative, 5e-2 absolute on energy per 10 units of |energy|
  T4 final pose: CA-RMSD <= 0.05 A in fp32 mode (the graph is injected, so no neighbour flips are possible)
"""
import os

import pytest
import torch

from util import load_golden, rel_err

pytestmark = pytest.mark.gpu

REAL = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
TOL = {"fp32": dict(rel=1e-4, energy=2e-3), "fp16": dict(rel=1e-2, energy=5e-2)}


def _weights(kind):
    """("pinder", state_dict, hparams, pos_width) from oracle/_ref, or the seeded synthetic weights."""
    from dfmdock_b200.synthetic import synthetic_hparams, synthetic_state_dict
    if kind == "pinder":
        p = os.path.join(REAL, "pinder_0.pt")
        if not os.path.exists(p):
            pytest.skip("oracle/_ref/pinder_0.pt not present (python oracle/build_ref.py in the build container)")
        ck = torch.load(p, weights_only=False)
        return ck["state_dict"], ck["hparams"], 67
    return synthetic_state_dict(0, 66), synthetic_hparams(66), 66


def _rigid_perturb(lig, n, gen, max_angle, tr_std):
    """n rigid copies of lig [L,3,3]: rotation about the CA centroid by a random axis-angle (|angle| <= max_angle) + N(0, tr_std^2)."""
    from oracle.dfmdock_oracle import aa_to_mat
    aa = torch.randn(n, 3, generator=gen)
    aa = aa / aa.norm(dim=-1, keepdim=True) * (torch.rand(n, 1, generator=gen) * max_angle)
    Rm = aa_to_mat(aa)                                            # [n,3,3]
    c = lig[:, 1].mean(0)
    tr = torch.randn(n, 1, 1, 3, generator=gen) * tr_std
    return torch.einsum("lac,ndc->nlad", lig - c, Rm) + c + tr


def _compare_with_oracle(model_fp32, model_fp16, net, batch, lig, t, picks, label):
    """Forward of the whole batch in both precisions on the SAME graph (Philox edges of the fp32 run, then injected);
    trajectories `picks` are recomputed by the oracle on those edges."""
    from oracle import dfmdock_oracle as orc  # noqa: F401
    o32 = model_fp32.score(lig, t, seed=11, forward_index=3, want_energy=True, return_edges=True)
    o32 = {k: v.clone() for k, v in o32.items()}
    o16 = model_fp16.score(lig, t, edges=o32["edges"], want_energy=True)
    torch.cuda.synchronize()
    worst = {"fp32": {}, "fp16": {}}
    for b in picks:
        bb = dict(batch)
        bb["lig_pos"] = lig[b].cpu()
        bb["t"] = t[b:b + 1].cpu()
        ref = net.forward(bb, edges=o32["edges"][b].cpu().long())
        for prec, o in (("fp32", o32), ("fp16", o16)):
            tol = TOL[prec]
            for k in ("f", "tr_score", "rot_score"):
                e = rel_err(o[k][b].cpu(), ref[k].reshape(o[k].shape[1:]))
                worst[prec][k] = max(worst[prec].get(k, 0.0), e)
                assert e <= tol["rel"], (label, prec, b, k, e)
            de = abs(float(o["energy"][b]) - float(ref["energy"]))
            worst[prec]["energy"] = max(worst[prec].get("energy", 0.0), de)
            assert de <= tol["energy"] * max(1.0, abs(float(ref["energy"])) / 10.0), (label, prec, b, de)
            assert int(o["num_clashes"][b]) == int(ref["num_clashes"]), (label, prec, b)
    print("worst errors vs oracle", label, worst)
    return worst


@pytest.mark.parametrize("weights", ["pinder", "synthetic"])
@pytest.mark.parametrize("n_res,n_traj,picks", [(150, 256, (0, 85, 170, 255)), (400, 64, (0, 63))])
def test_benchmark_configs_vs_oracle_near_and_far(weights, n_res, n_traj, picks):
    """BASELINE config #3 (2x150, 256 trajectories) and #4 (2x400, 64 trajectories) at full batch size."""
    from dfmdock_b200 import Score_Model
    from dfmdock_b200.features import synthetic_complex
    from oracle import dfmdock_oracle as orc
    sd, hp, width = _weights(weights)
    batch = synthetic_complex(n_res, n_res, seed=0, pos_width=width)
    m32 = Score_Model(sd, hp, precision="fp32").to("cuda")
    m16 = Score_Model(sd, hp, precision="fp16").to("cuda")
    m32.set_complex(batch)
    m16.set_complex(batch)
    net = orc.OracleNet(sd, cut_off=hp["model"]["cut_off"])
    gen = torch.Generator().manual_seed(5)
    # near contact: the generator puts the ligand 25 A along x; pull it in and jitter every trajectory rigidly
    near0 = batch["lig_pos"] - torch.tensor([12.0, 0.0, 0.0])
    near = _rigid_perturb(near0, n_traj, gen, max_angle=0.6, tr_std=3.0)
    t_near = torch.linspace(0.9, 0.05, n_traj)
    _compare_with_oracle(m32, m16, net, batch, near.cuda(), t_near.cuda(), picks, "c%d near %s" % (n_res, weights))
    # far: what every trajectory starts from (randomize_pose, Philox; t = 1.0)
    far, _, _ = m32.randomize_pose(batch["lig_pos"], n_traj, seed=17)
    sep = (far[:, :, 1].mean(1).cpu() - batch["rec_pos"][:, 1].mean(0)).norm(dim=-1)
    assert float(sep.max()) > 60.0
    _compare_with_oracle(m32, m16, net, batch, far, torch.ones(n_traj, device="cuda"), picks, "c%d far %s" % (n_res, weights))


needs_real = pytest.mark.skipif(not os.path.exists(os.path.join(REAL, "golden_real.pt")),
                                reason="oracle/_ref not built (python oracle/build_ref.py in the build container)")


@needs_real
@pytest.mark.parametrize("precision", ["fp32", "fp16"])
@pytest.mark.parametrize("name,ckpt,centre_mode", [("t4_1QA9_dips_s10.pt", "dips_model_0", 0),
                                                   ("t4_1QA9_pinder_s10_clash.pt", "pinder_0", 1)])
def test_free_running_trajectory_same_noise_vs_reference(name, ckpt, centre_mode, precision):
    """T4 (north_star "final Calpha-RMSD"): the CUDA sampler runs 10 reverse steps FREE (its own poses feed its own
    next forward) on the reference's recorded noise -- initial rotation / translation, the neighbour table of every forward,
    z of every step -- and is compared with the reference's trajectory.

    A free trajectory is a chaotic map in fp32 (g(t)^2 dt ~ 1100 at t = 1; binned pair features): the golden stores how far
    the reference's own final pose moves when its inputs are perturbed by a relative 1e-6 ("sensitivity_rmsd", 0.004 - 0.09 A
    on these two trajectories; tests/golden/make_t4_golden.py documents that no seed of 40 stays below 0.01 A).  Bounds:
      * fp32 mode, for as long as no pair-feature bin of the injected graph differs from the reference's (integer work,
        compared every step): every atom within 5e-3 A of the reference's pose;
      * final pose: CA-RMSD <= max(0.05 A, 3 x the trajectory's sensitivity floor) in fp32 mode (0.05 A alone when no bin
        flipped), max(0.5 A, 3 x floor) in fp16 mode; final energy within 2e-2 when no bin flipped."""
    from dfmdock_b200 import Score_Model
    from dfmdock_b200.features import batch_from_record
    from oracle import dfmdock_oracle as orc
    g = load_golden(name)
    ck = torch.load(os.path.join(REAL, ckpt + ".pt"), weights_only=False)
    model = Score_Model(ck["state_dict"], ck["hparams"], precision=precision).to("cuda")
    batch = batch_from_record(torch.load(os.path.join(REAL, "db5_1QA9.pt"), weights_only=False), pos_width=model.pos_width)
    model.set_complex(batch)
    R, L = batch["rec_pos"].shape[0], batch["lig_pos"].shape[0]
    N = R + L
    S = int(g["num_steps"])
    ts = torch.linspace(1.0, 1e-3, S)
    dt = float(ts[0] - ts[1])
    rows = torch.arange(N)[:, None].expand(N, 60)

    def bins_of(i):          # bins of the injected edges on the REFERENCE's pose at forward i
        pos = torch.cat([batch["rec_pos"], g["fwd_lig_pos"][i]], 0)
        pos = pos - g["fwd_lig_pos"][i][:, 1].mean(0)
        nbr = g["nbr"][i].long()
        return [x[rows, nbr] for x in orc.spatial_bins(pos)]

    def flips(i):
        ft = model.debug_read(1, 1, (N, 64), dtype=torch.int32).cpu()[:, :60].long()
        mine = (ft & 63, (ft >> 6) & 31, (ft >> 11) & 31, (ft >> 16) & 15)
        return sum(int((a != b).sum()) for a, b in zip(mine, bins_of(i)))

    lig, tr_u, rot_u = model.randomize_pose(batch["lig_pos"], 1, rot0=g["rot0"][None], tr0=g["tr0"], centre_mode=centre_mode)
    first_flip, dev = None, []
    for i in range(S + 1):
        dev.append(float((lig[0].cpu() - (g["fwd_lig_pos"][i] if i < S + 1 else g["lig_pos"])).norm(dim=-1).max()))
        if precision == "fp32" and first_flip is None:
            assert dev[-1] <= 5e-3, (i, dev)
        last = i == S
        o = model.score(lig, ts[min(i, S - 1)][None], edges=g["nbr"][i][None].int(), want_energy=last)
        if first_flip is None and flips(i) > 0:
            first_flip = i
        if last:
            break
        ns = 0.0 if i == S - 1 else 0.5
        model.reverse_step(lig, rot_u, tr_u, o["tr_score"], o["rot_score"], float(ts[i]), dt, ns, ns, z=g["z"][i][None],
                           use_clash_force=bool(g["use_clash_force"]), centre_mode=centre_mode)
    torch.cuda.synchronize()
    rmsd = float(((lig[0, :, 1].cpu() - g["lig_pos"][:, 1]) ** 2).sum(-1).mean().sqrt())
    de = abs(float(o["energy"][0]) - float(g["energy"]))
    floor = max(g["sensitivity_rmsd"])
    print("T4 %s %s: final CA-RMSD %.2e A (reference's own 1e-6 sensitivity %.2e A), first bin flip at forward %s, "
          "max atom deviation per forward %s, energy diff %.2e" % (name, precision, rmsd, floor, first_flip, ["%.1e" % d for d in dev], de))
    if precision == "fp32":
        assert rmsd <= (0.05 if first_flip is None else max(0.05, 3 * floor)), (rmsd, floor, first_flip)
        if first_flip is None:
            assert de <= 2e-2, de
            assert int(o["num_clashes"][0]) == int(g["num_clashes"])
    else:
        assert rmsd <= max(0.5, 3 * floor), (rmsd, floor)
    assert torch.isfinite(o["energy"]).all()


@needs_real
@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_real_checkpoints_far_poses_vs_live_reference_golden(precision):
    """Both shipped checkpoints on 1QA9 / 7CEI / 4POU at randomize_pose starts 60 / 140 A away, t = 1.0 (outputs of the
    unmodified reference, oracle/build_ref.py): the regime where radial * w1r dominates the edge pre-activation."""
    from dfmdock_b200 import Score_Model
    from dfmdock_b200.features import batch_from_record
    golden = [g for g in torch.load(os.path.join(REAL, "golden_real.pt"), weights_only=False) if g["pose"] == "far"]
    assert len(golden) == 12
    tol = TOL[precision]
    models, worst = {}, {}
    for g in golden:
        if g["ckpt"] not in models:
            ck = torch.load(os.path.join(REAL, g["ckpt"] + ".pt"), weights_only=False)
            models[g["ckpt"]] = Score_Model(ck["state_dict"], ck["hparams"], precision=precision).to("cuda")
        model = models[g["ckpt"]]
        batch = batch_from_record(torch.load(os.path.join(REAL, "db5_%s.pt" % g["complex"]), weights_only=False), pos_width=model.pos_width)
        model.set_complex(batch)
        out = model.score(g["lig_pos"][None], torch.tensor([g["t"]]), edges=g["nbr"][None].int(), want_energy=True)
        # tr = mean f and rot = mean r x f cancel at these separations (near-uniform force field, sum r = 0): a relative
        # error eps in f can appear as kappa * eps in the direction of the reduced vector, kappa = mean|v_l| / |mean v_l|
        # (3 - 300 here, ~10 at the bound pose).  f is held to the plain tolerance; the scores to tol * max(1, kappa / 10).
        r = g["lig_pos"][:, 1] - g["lig_pos"][:, 1].mean(0)
        cr = torch.cross(r, g["f"], dim=-1)
        kappa = {"tr_score": float(g["f"].norm(dim=-1).mean() / g["f"].mean(0).norm()),
                 "rot_score": float(cr.norm(dim=-1).mean() / cr.mean(0).norm())}
        for k in ("f", "tr_score", "rot_score"):
            e = rel_err(out[k].cpu()[0], g[k].reshape(out[k].shape[1:]))
            bound = tol["rel"] * max(1.0, kappa.get(k, 0.0) / 10.0)
            worst[k] = max(worst.get(k, 0.0), e / bound)
            assert e <= bound, (g["ckpt"], g["complex"], g["sep"], k, e, kappa.get(k))
        assert abs(float(out["energy"][0]) - float(g["energy"])) <= tol["energy"]
        assert int(out["num_clashes"][0]) == int(g["num_clashes"])
    print("worst error / bound at far poses", precision, worst)


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_reference_shaped_forward_returns_interface_logits(precision):
    """Score_Model.forward(batch) returns the reference's six entries; `ires` = to_ires(h) (score_net_mlsb.py:383)."""
    from dfmdock_b200 import Score_Model
    from oracle import dfmdock_oracle as orc
    from util import case_small
    sd, hp, batch = case_small()
    model = Score_Model(sd, hp, precision=precision).to("cuda")
    model.edge_rng = "philox"
    b = dict(batch)
    b["t"] = torch.tensor([0.4])
    out = model(b)
    assert set(out) == {"tr_score", "rot_score", "energy", "f", "num_clashes", "ires"}
    N = batch["rec_pos"].shape[0] + batch["lig_pos"].shape[0]
    assert out["ires"].shape == (N, 1) and out["tr_score"].shape == (1, 3) and out["f"].shape == (batch["lig_pos"].shape[0], 3)
    edges = model.debug_read(1, 3, (N, 64), dtype=torch.int32)[:, :model.edges_per_node].cpu().long()
    ref = orc.OracleNet(sd).forward(b, edges=edges)
    e = rel_err(out["ires"].cpu(), ref["ires"])
    assert e <= TOL[precision]["rel"], e
    # a checkpoint stripped of the (dead at inference) to_ires.* tensors still loads; ires is NaN then
    slim = {k: v for k, v in sd.items() if not k.startswith("to_ires")}
    m2 = Score_Model(slim, hp, precision=precision).to("cuda")
    m2.edge_rng = "philox"
    o2 = m2(b)
    assert torch.isnan(o2["ires"]).all() and torch.isfinite(o2["energy"])


def test_batched_sampler_uses_the_checkpoints_own_sigmas():
    """dfm_sample takes g(t) from hyper_parameters.diffuser (dfm_set_schedule), like the step-wise path that uses
    model.so3_diffuser / r3_diffuser: the two must agree bit for bit for NON-default sigmas as well."""
    import copy
    from dfmdock_b200 import Score_Model
    from util import case_small
    sd, hp, batch = case_small()
    hp2 = copy.deepcopy(hp)
    hp2["diffuser"]["so3"].update(min_sigma=0.05, max_sigma=1.0)
    hp2["diffuser"]["r3"].update(min_sigma=0.2, max_sigma=12.0)
    outs = {}
    for tag, h in (("default", hp), ("custom", hp2)):
        model = Score_Model(sd, h, precision="fp16").to("cuda")
        model.set_complex(batch)
        a = model.sample(batch["lig_pos"], 3, num_steps=4, seed=6)
        a = {k: v.clone() for k, v in a.items()}
        b = model.sample(batch["lig_pos"], 3, num_steps=4, seed=6, record=True)
        for k in ("lig_pos", "rot_update", "tr_update", "energy"):
            assert torch.equal(a[k], b[k]), (tag, k)
        outs[tag] = a
    assert not torch.equal(outs["default"]["tr_update"], outs["custom"]["tr_update"])


@pytest.mark.gpu
@pytest.mark.parametrize("n_rec,n_lig,B", [(150, 150, 64), (151, 148, 48), (120, 97, 96)])
def test_last_layer_fused_launch_is_bit_identical(n_rec, n_lig, B):
    """DFM_LAST_FUSED: the last layer's edge MLP and the coordinate head (src/models/egnn.py:118-148) as one launch, tiles of
    gated messages handed from an edge role to a coordinate-head role through an L2-resident ring (csrc/last_ring.cuh).  Same
    arithmetic on the same rows in the same order as the two-kernel form -> forces and scores equal bit for bit; even / odd
    receptor and ligand sizes cover the tiles that straddle the chains and the duplicated last tile of an odd complex."""
    import torch
    from dfmdock_b200 import Score_Model
    from dfmdock_b200.features import synthetic_complex
    from dfmdock_b200.synthetic import synthetic_hparams, synthetic_state_dict
    sd, hp = synthetic_state_dict(1, 66), synthetic_hparams(66)
    batch = synthetic_complex(n_rec, n_lig, seed=5)
    batch["lig_pos"] = batch["lig_pos"] - torch.tensor([15.0, 0.0, 0.0])
    model = Score_Model(sd, hp, precision="fp16").to("cuda")
    model.set_complex(batch)
    g = torch.Generator().manual_seed(0)
    lig = (batch["lig_pos"][None] + torch.randn(B, 1, 1, 3, generator=g)).contiguous()
    t = torch.full((B,), 0.35)
    a = model.score(lig, t, seed=11, stream_base=3, forward_index=2)
    before = model.launch_count
    model.last_fused = True
    b = model.score(lig, t, seed=11, stream_base=3, forward_index=2)
    fused_launches = model.launch_count - before
    model.last_fused = False
    c = model.score(lig, t, seed=11, stream_base=3, forward_index=2)
    unfused_launches = model.launch_count - before - fused_launches
    assert fused_launches == unfused_launches - 1, (fused_launches, unfused_launches)      # the fused form really ran
    for k in ("f", "tr_score", "rot_score"):
        assert torch.equal(a[k], c[k]), k
        assert torch.equal(a[k], b[k]), (k, float((a[k] - b[k]).abs().max()))
