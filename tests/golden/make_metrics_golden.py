"""TEST INFRASTRUCTURE.  Golden vectors for the docking metrics from the UNMODIFIED reference (src/utils/metrics.py
compute_metrics, imported through oracle/ref_shims.py): rigid perturbations of a real complex (db5 1QA9) and of a small
synthetic complex, including moved receptors, near-native and far poses.  Writes tests/golden/metrics_cases.pt."""
import math
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def rot(axis, ang):
    axis = torch.tensor(axis, dtype=torch.float64)
    axis = axis / axis.norm()
    K = torch.tensor([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]], dtype=torch.float64)
    return (torch.eye(3, dtype=torch.float64) + math.sin(ang) * K + (1 - math.cos(ang)) * (K @ K)).float()


def move(x, Rm, t):
    c = x.reshape(-1, 3).mean(0)
    return (x - c) @ Rm.T + c + torch.tensor(t)


def main():
    from oracle import ref_shims
    from dfmdock_b200.features import synthetic_complex
    ref_shims.install()
    from utils.metrics import compute_metrics      # the reference's own function
    cases = []
    rec = ref_shims.load_db5_record(os.path.join(ref_shims.REFERENCE_ROOT, "data", "db5_test", "1QA9.pt"))
    syn = synthetic_complex(24, 18, seed=11)
    syn["lig_pos"] = syn["lig_pos"] - torch.tensor([17.0, 0.0, 0.0])
    sets = {"1QA9": (rec["receptor"]["pos"].float(), rec["ligand"]["pos"].float()), "synth": (syn["rec_pos"], syn["lig_pos"])}
    g = torch.Generator().manual_seed(5)
    for name, (nr, nl) in sets.items():
        poses = [(rot([0, 0, 1], 0.0), [0.0, 0.0, 0.0], None),                     # native
                 (rot([1, 2, 3], 0.05), [0.5, -0.3, 0.2], None),                   # near native
                 (rot([0, 1, 0], 0.4), [3.0, 1.0, -2.0], None),
                 (rot([1, 0, 1], 2.5), [25.0, -10.0, 8.0], None),                  # far
                 (rot([3, 1, 2], 1.0), [6.0, 6.0, 0.0], (rot([1, 1, 0], 0.7), [4.0, -9.0, 2.0]))]   # receptor moved too
        for k in range(3):
            a = torch.randn(3, generator=g).tolist()
            poses.append((rot(a, float(torch.rand(1, generator=g)) * 3.0), (torch.randn(3, generator=g) * 8).tolist(), None))
        for Rm, t, recmove in poses:
            ml = move(nl, Rm, t)
            mr = nr.clone() if recmove is None else move(nr, recmove[0], recmove[1])
            out = compute_metrics((mr, ml), (nr, nl))
            cases.append({"set": name, "model_rec": mr, "model_lig": ml, "native_rec": nr, "native_lig": nl,
                          "out": {k: float(v) for k, v in out.items()}})
            print(name, {k: round(float(v), 4) for k, v in out.items()})
    torch.save(cases, os.path.join(HERE, "metrics_cases.pt"))


if __name__ == "__main__":
    main()
