"""Docking metrics: oracle vs the reference goldens on CPU, CUDA (through the C ABI) vs the
goldens and size-independent properties on the GPU.  Tolerances: RMSDs 2e-3 A absolute (fp32 inputs; the reference
itself runs a float32 SVD), fnat exact to its 6-decimal rounding, DockQ 1e-3."""
import math
import os

import pytest
import torch

from util import GOLDEN

KEYS = ("c_rmsd", "i_rmsd", "l_rmsd", "fnat", "DockQ")
TOL = {"c_rmsd": 2e-3, "i_rmsd": 2e-3, "l_rmsd": 2e-3, "fnat": 2e-6, "DockQ": 1e-3}


def _cases():
    return torch.load(os.path.join(GOLDEN, "metrics_cases.pt"), map_location="cpu", weights_only=False)


def test_metrics_oracle_matches_reference_golden():
    from oracle import metrics_oracle as mo
    for c in _cases():
        got = mo.compute_metrics((c["model_rec"], c["model_lig"]), (c["native_rec"], c["native_lig"]))
        for k in KEYS:
            assert abs(got[k] - c["out"][k]) <= TOL[k], (c["set"], k, got[k], c["out"][k])


def _rot(axis, ang):
    axis = torch.tensor(axis, dtype=torch.float64)
    axis = axis / axis.norm()
    K = torch.tensor([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]], dtype=torch.float64)
    return (torch.eye(3, dtype=torch.float64) + math.sin(ang) * K + (1 - math.cos(ang)) * (K @ K)).float()


@pytest.mark.gpu
def test_metrics_cuda_matches_reference_golden():
    from dfmdock_b200.metrics import compute_metrics, compute_metrics_batch
    cases = _cases()
    for c in cases:
        got = compute_metrics((c["model_rec"], c["model_lig"]), (c["native_rec"], c["native_lig"]))
        for k in KEYS:
            assert abs(got[k] - c["out"][k]) <= TOL[k], (c["set"], k, got[k], c["out"][k])
    # batched call, shared receptor: the cases of one set whose receptor did not move
    for name in ("1QA9", "synth"):
        sel = [c for c in cases if c["set"] == name and torch.equal(c["model_rec"], c["native_rec"])]
        out = compute_metrics_batch(sel[0]["native_rec"], torch.stack([c["model_lig"] for c in sel]), sel[0]["native_rec"],
                                    sel[0]["native_lig"]).cpu()
        for row, c in zip(out, sel):
            for i, k in enumerate(KEYS):
                assert abs(float(row[i]) - c["out"][k]) <= TOL[k], (name, k)


@pytest.mark.gpu
def test_metrics_properties_full_size():
    """256 poses of the 2x150-residue benchmark complex: native scores perfectly; C-/I-RMSD and Fnat do not change when
    the whole model is moved rigidly (only L-RMSD's receptor superposition absorbs it); batched == one by one."""
    from dfmdock_b200.features import synthetic_complex
    from dfmdock_b200.metrics import compute_metrics, compute_metrics_batch
    b = synthetic_complex(150, 150, seed=0)
    nr, nl = b["rec_pos"], b["lig_pos"] - torch.tensor([12.0, 0.0, 0.0])
    g = torch.Generator().manual_seed(3)
    T = 256
    ligs = []
    for t in range(T):
        Rm = _rot(torch.randn(3, generator=g).tolist(), 0.0 if t == 0 else float(torch.rand(1, generator=g)) * (0.1 if t % 2 else 3.0))
        tr = torch.zeros(3) if t == 0 else torch.randn(3, generator=g) * (0.5 if t % 2 else 10.0)
        c = nl.reshape(-1, 3).mean(0)
        ligs.append((nl - c) @ Rm.T + c + tr)
    ligs = torch.stack(ligs)
    out = compute_metrics_batch(nr, ligs, nr, nl).cpu()
    assert torch.isfinite(out).all()
    assert out[0, 0] < 1e-3 and out[0, 1] < 1e-3 and out[0, 2] < 1e-3 and abs(float(out[0, 3]) - 1.0) < 1e-6 and abs(float(out[0, 4]) - 1.0) < 1e-4
    assert (out[:, 3] >= 0).all() and (out[:, 3] <= 1).all() and (out[:, 4] > 0).all() and (out[:, 4] <= 1.0 + 1e-6).all()
    # DockQ formula (metrics.py:71-74) on the kernel's own components
    dq = (out[:, 3] + 1 / (1 + (out[:, 1] / 1.5) ** 2) + 1 / (1 + (out[:, 2] / 8.5) ** 2)) / 3
    assert torch.allclose(dq, out[:, 4], atol=1e-5)
    # global rigid motion of the whole model
    G, gt = _rot([0.3, -1.0, 0.5], 1.3), torch.tensor([40.0, -25.0, 13.0])
    mr2 = nr @ G.T + gt
    ligs2 = ligs @ G.T + gt
    out2 = compute_metrics_batch(mr2[None].expand(T, -1, -1, -1).contiguous(), ligs2, nr, nl).cpu()
    assert torch.allclose(out2, out, atol=3e-3)
    one = compute_metrics((nr, ligs[17]), (nr, nl))
    assert all(abs(one[k] - float(out[17, i])) < 1e-5 for i, k in enumerate(KEYS))


@pytest.mark.gpu
def test_metrics_errors_are_loud():
    from dfmdock_b200.metrics import compute_metrics_batch
    x = torch.zeros(4, 3, 3)
    with pytest.raises(ValueError):
        compute_metrics_batch(x, torch.zeros(2, 5, 3, 3), x, torch.zeros(6, 3, 3))
    with pytest.raises(RuntimeError):
        compute_metrics_batch(x, torch.zeros(2, 5, 3, 3), x, torch.zeros(5, 3, 3), device="cpu")
