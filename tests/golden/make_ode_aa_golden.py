"""Generates tests/golden/ode_aa_cases.pt from the UNMODIFIED reference (build container only):
  * torch_reverse(ode=True) of both diffusers (src/utils/so3_diffuser.py:344-369, src/utils/r3_diffuser.py:40-55) followed by
    the reference's modify_coords / rot_compose (src/inference_base.py:311-352) on seeded poses and scores;
  * modify_aa_coords, both variants (src/inference_base.py:354-364 -- re-stated inline because that module imports biotite
    and esm at load time -- and src/inference.py:256-266 likewise), built on the reference's own axis_angle_to_matrix.
    python tests/golden/make_ode_aa_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_shims  # noqa: E402
from util import case_small  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ode_aa_cases.pt")


def main():
    ref_shims.install()
    from utils.geometry import axis_angle_to_matrix, matrix_to_axis_angle
    from utils.r3_diffuser import R3Diffuser
    from utils.so3_diffuser import SO3Diffuser
    from dfmdock_b200.synthetic import synthetic_hparams

    hp = ref_shims.AttrDict(synthetic_hparams(66))
    hp["diffuser"]["so3"]["cache_dir"] = "/tmp/dfmdock_so3_cache"
    so3, r3 = SO3Diffuser(hp.diffuser.so3), R3Diffuser(hp.diffuser.r3)

    def modify_coords(x, rot, tr):                         # src/inference_base.py:342-352, verbatim semantics
        center = torch.mean(x[..., 1, :], dim=0, keepdim=True)
        rot = axis_angle_to_matrix(rot).squeeze()
        x = (x - center) @ rot.T + center
        return x + tr

    def rot_compose(r1, r2):                               # src/inference_base.py:311-316
        R = torch.einsum("b i j, b j k -> b i k", axis_angle_to_matrix(r2), axis_angle_to_matrix(r1))
        return matrix_to_axis_angle(R)

    _, _, batch = case_small()
    g = torch.Generator().manual_seed(11)
    ts = torch.linspace(1.0, 1e-3, 10)
    dt = ts[0] - ts[1]
    lig = batch["lig_pos"].clone()
    rot_u, tr_u = torch.zeros(1, 3), torch.zeros(1, 3)
    ode = []
    for i in (0, 3, 6, 9):
        t = float(ts[i])
        rs, trs = 0.3 * torch.randn(1, 3, generator=g), 0.05 * torch.randn(1, 3, generator=g)
        rot = so3.torch_reverse(score_t=rs, t=t, dt=dt, noise_scale=0.5, ode=True)
        tr = r3.torch_reverse(score_t=trs, t=t, dt=dt, noise_scale=0.5, ode=True)
        before = lig.clone()
        lig = modify_coords(lig, rot, tr)
        tr_u = tr_u + tr
        rot_u = rot_compose(rot_u, rot)
        ode.append({"t": t, "dt": float(dt), "rot_score": rs, "tr_score": trs, "rot": rot, "tr": tr, "lig_before": before,
                    "lig_after": lig.clone(), "rot_update": rot_u.clone(), "tr_update": tr_u.clone()})

    # all-atom transform: 7 pseudo-atoms per ligand residue scattered around CA
    bb = batch["lig_pos"].double().numpy()
    aa = (batch["lig_pos"][:, 1:2, :] + 1.5 * torch.randn(bb.shape[0], 7, 3, generator=g)).reshape(-1, 3).numpy().astype(np.float32)
    cases = []
    for k in range(4):
        rot = torch.randn(1, 3, generator=g)
        rot = rot / rot.norm() * (0.2 + 0.9 * k)
        tr = 10.0 * torch.randn(1, 3, generator=g)
        Rm = axis_angle_to_matrix(rot).squeeze().cpu().numpy()
        c0 = bb[:, 1].mean(axis=0)                          # src/inference_base.py:355
        out0 = (aa - c0) @ Rm.T + c0 + tr.cpu().numpy()
        c1 = aa.mean(axis=0)                                # src/inference.py:257
        out1 = (aa - c1) @ Rm.T + c1 + tr.cpu().numpy()
        cases.append({"rot": rot, "tr": tr, "out_ca_centre": torch.from_numpy(np.asarray(out0, dtype=np.float64)),
                      "out_aa_centre": torch.from_numpy(np.asarray(out1, dtype=np.float64))})
    torch.save({"ode": ode, "aa": torch.from_numpy(aa), "aa_cases": cases}, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
