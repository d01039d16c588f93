"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's docking metrics (src/utils/metrics.py:3-121), written
as the reference writes it (torch SVD Kabsch with the reflection fix, dense residue-pair minimum distances).
Pinned against the live reference by tests/golden/make_metrics_golden.py -> tests/golden/metrics_*.pt."""
import torch


def find_rigid_alignment(A, B):
    """src/utils/metrics.py:91-121"""
    a_mean, b_mean = A.mean(0), B.mean(0)
    H = (A - a_mean).T @ (B - b_mean)
    U, S, Vt = torch.linalg.svd(H)
    R = Vt.T @ U.T
    if torch.linalg.det(R) < 0:
        R = (Vt.T @ torch.diag(torch.tensor([1.0, 1.0, -1.0], dtype=R.dtype))) @ U.T
    t = b_mean - R @ a_mean
    return R, t


def rmsd(pred, label):
    return torch.sqrt(((pred - label) ** 2).sum(-1).mean())


def res_min_dist(x1, x2):
    """src/utils/metrics.py:77-85: [n1,3,3], [n2,3,3] -> [n1,n2] minimum over the 3x3 atom pairs"""
    d = x1[:, None, :, None, :] - x2[None, :, None, :, :]
    return (d ** 2).sum(-1).sqrt().flatten(start_dim=-2).min(dim=-1).values


def compute_metrics(model, native):
    mr, ml, nr, nl = (x.squeeze().double() for x in (model[0], model[1], native[0], native[1]))
    fl = lambda x: x.flatten(end_dim=1)
    pred, label = torch.cat([fl(mr), fl(ml)]), torch.cat([fl(nr), fl(nl)])
    R, t = find_rigid_alignment(pred, label)
    c = float(rmsd(pred @ R.T + t, label))
    nd = res_min_dist(nr, nl)
    idx = torch.where(nd < 10.0)
    r1, r2 = torch.unique(idx[0]), torch.unique(idx[1])
    pred, label = torch.cat([fl(mr[r1]), fl(ml[r2])]), torch.cat([fl(nr[r1]), fl(nl[r2])])
    R, t = find_rigid_alignment(pred, label)
    i = float(rmsd(pred @ R.T + t, label))
    R, t = find_rigid_alignment(fl(mr), fl(nr))
    l = float(rmsd(fl(ml) @ R.T + t, fl(nl)))
    act = torch.where(nd < 5.5)
    md = res_min_dist(mr, ml)
    fnat = round(int((md[act[0], act[1]] < 5.5).sum()) / (len(act[0]) + 1e-6), 6)
    dockq = (fnat + 1.0 / (1.0 + (i / 1.5) ** 2) + 1.0 / (1.0 + (l / 8.5) ** 2)) / 3
    return {"c_rmsd": c, "i_rmsd": i, "l_rmsd": l, "fnat": fnat, "DockQ": dockq}
