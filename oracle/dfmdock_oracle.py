"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of DFMDock's reverse-diffusion hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this file.  The product path (dfmdock_b200/) never does: it fails loudly when the CUDA
library is missing.

What is restated (reference = Graylab/DFMDock @ e2fd4991, paths relative to /root/reference):
  * Score_Net.forward                      src/models/score_net_mlsb.py:343-425
  * get_spatial_matrix / get_bins          src/models/score_net_mlsb.py:30-70
  * get_knn_and_sample(_graph)             src/models/score_net_mlsb.py:85-157
  * get_coords6d, calc_dihedral/planar     src/utils/coords6d.py:10-103
  * E_GCL.forward and its sub-models       src/models/egnn.py:95-159
  * GraphNorm (third party, torch_geometric==2.6.0, call site src/models/egnn.py:74)
  * relpos / get_position_matrix           src/inference_base.py:230-292 (= src/utils/crop.py:3-49)
  * SO3Diffuser / R3Diffuser schedules     src/utils/so3_diffuser.py:210-227,344-369; src/utils/r3_diffuser.py:20-55
  * axis-angle <-> quaternion <-> matrix   src/utils/geometry.py:7-200
  * randomize_pose / modify_coords / rot_compose / get_clash_force / Euler_Maruyama_sampler
                                           src/inference_base.py:311-468 (centre_mode 0)
                                           src/inference.py:213-370       (centre_mode 1)
  * torch_reverse(ode=True) branch         src/utils/so3_diffuser.py:366-367, src/utils/r3_diffuser.py:53-54
  * modify_aa_coords (all-atom output)     src/inference_base.py:354-364, src/inference.py:256-266

The arithmetic is written "as the reference writes it" (dense N x N x 100 one-hot pair features,
N^2-row embedding GEMMs, [R, L, 512] energy tensor) because this file is also the CPU baseline
that bench.py times: it must cost what the reference costs.  It is fp32 throughout, like the
reference.

Parity pin: the reference has no tests or golden vectors for this path, so
the oracle is pinned against outputs of the reference itself, run in the build container through
oracle/ref_shims.py: tests/golden/make_goldens.py writes tests/golden/*.pt and
tests/test_oracle_vs_golden.py checks this file against them (and against the live reference
when /root/reference is present).

Determinism hooks (not in the reference): `edges` injects the neighbour table, `exp_noise`
injects the Exp(1) draws that torch.multinomial(replacement=False) consumes
(multinomial(p, k) == topk(p / q, k) with q = empty_like(p).exponential_(1)), and the sampler
takes `noise` = dict of pre-drawn tensors.
"""
import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

KNN = 20
NSAMPLE = 40


# ----------------------------------------------------------------------------------------------
# 6D pair features  (src/utils/coords6d.py:10-103)

def _unit(v):
    return v / v.norm(dim=-1, keepdim=True)


def dihedral_deg(p0, p1, p2, p3):
    """src/utils/coords6d.py:23-43"""
    b1 = p0 - p1
    b2 = p1 - p2
    b3 = p2 - p3
    n1 = _unit(torch.linalg.cross(b1, b2))
    n2 = _unit(torch.linalg.cross(b2, b3))
    m1 = torch.linalg.cross(n1, _unit(b2))
    ang = torch.atan2((m1 * n2).sum(-1), (n1 * n2).sum(-1))
    return ang * 180 / math.pi


def planar_deg(p0, p1, p2):
    """src/utils/coords6d.py:46-58"""
    v1 = p0 - p1
    v2 = p2 - p1
    ang = torch.acos((v1 * v2).sum(-1) / (v1.norm(dim=-1) * v2.norm(dim=-1)))
    return ang * 180 / math.pi


def virtual_cb(pos):
    """src/utils/coords6d.py:72-76; pos [n,3,3] (N, CA, C)."""
    n_at, ca, c_at = pos[:, 0], pos[:, 1], pos[:, 2]
    b = ca - n_at
    c = c_at - ca
    a = torch.cross(b, c, dim=-1)
    return -0.58273431 * a + 0.56802827 * b - 0.54067466 * c + ca


def coords6d(pos):
    """src/utils/coords6d.py:62-103 -> dist, omega, theta, phi each [n,n] (degrees)."""
    n = pos.shape[0]
    n_at, ca = pos[:, 0], pos[:, 1]
    cb = virtual_cb(pos)
    row = lambda v: v[:, None, :].expand(n, n, 3)
    col = lambda v: v[None, :, :].expand(n, n, 3)
    dist = (row(ca) - col(ca)).norm(dim=-1)
    omega = dihedral_deg(row(ca), row(cb), col(cb), col(ca))
    theta = dihedral_deg(row(n_at), row(ca), row(cb), col(cb))
    phi = planar_deg(row(ca), row(cb), col(cb))
    return dist, omega, theta, phi


def bin_index(x, lo, hi, nbins):
    """src/models/score_net_mlsb.py:61-70: number of linspace(lo,hi,nbins-1) edges strictly below x."""
    edges = torch.linspace(lo, hi, nbins - 1, device=x.device)
    return torch.sum(x.unsqueeze(-1) > edges, dim=-1)


def spatial_bins(pos):
    """src/models/score_net_mlsb.py:30-51 -> int64 bins (d, omega, theta, phi), each [n,n]."""
    dist, omega, theta, phi = coords6d(pos)
    near = dist < 22.0
    d_bin = bin_index(dist, 3.25, 50.75, 40)
    out = [d_bin]
    for x, lo, hi, nb in ((omega, -180.0, 180.0, 24), (theta, -180.0, 180.0, 24), (phi, 0.0, 180.0, 12)):
        b = bin_index(x, lo, hi, nb)
        b[~near] = 0
        b.fill_diagonal_(0)
        out.append(b)
    return out


def spatial_matrix(pos):
    """src/models/score_net_mlsb.py:53-59 -> one-hot [n,n,100] fp32."""
    d, o, t, p = spatial_bins(pos)
    return torch.cat([F.one_hot(d, 40).float(), F.one_hot(o, 24).float(),
                      F.one_hot(t, 24).float(), F.one_hot(p, 12).float()], dim=-1)


# ----------------------------------------------------------------------------------------------
# relative-position features (src/inference_base.py:230-292)

def relpos_bins(n_rec, n_lig):
    """clamp(i-j+32, 0, 64) within a chain, 65 across chains; int64 [N,N]."""
    n = n_rec + n_lig
    idx = torch.arange(n)
    chain = (idx >= n_rec).long()
    off = torch.clamp(idx[:, None] - idx[None, :] + 32, 0, 64)
    return torch.where(chain[:, None] == chain[None, :], off, torch.full_like(off, 65))


def position_matrix(n_rec, n_lig, width=66, sym=0.0):
    """One-hot(66) of relpos_bins; `width` 67 appends the homomer channel."""
    pm = F.one_hot(relpos_bins(n_rec, n_lig), 66).float()
    if width == 67:
        pm = torch.cat([pm, torch.full(pm.shape[:-1] + (1,), float(sym))], dim=-1)
    return pm


# ----------------------------------------------------------------------------------------------
# stochastic graph (src/models/score_net_mlsb.py:85-157)

def knn_and_sample(ca, exp_noise=None, generator=None, return_noise=False):
    """Returns nbr [n, K] int64 (K = knn + samples) and, optionally, the Exp(1) draws used.

    exp_noise: optional [n, n-knn] tensor replacing empty_like(p).exponential_(1).
    """
    n = ca.shape[0]
    knn, ns = KNN, NSAMPLE
    if n < knn:
        knn, ns = n, 0
    if n < knn + ns:
        ns = n - knn
    dmat = torch.cdist(ca, ca)
    knn_idx = torch.topk(dmat, k=knn, largest=False).indices
    used = None
    if ns > 0:
        keep = torch.ones(n, n, dtype=torch.bool)
        keep.scatter_(1, knn_idx, False)
        d_rest = dmat[keep].view(n, -1)
        d_rest = torch.where(d_rest < 1e-10, torch.tensor(1e-10), d_rest)
        w = 1 / torch.pow(d_rest, 3)
        p = w / w.sum(dim=1, keepdim=True)
        p = torch.clamp(torch.nan_to_num(p, nan=0.0, posinf=0.0, neginf=0.0), min=0)
        p = p / p.sum(dim=1, keepdim=True)
        rest_idx = torch.arange(n).expand(n, n)[keep].view(n, -1)
        if exp_noise is None:
            q = torch.empty_like(p).exponential_(1, generator=generator)
        else:
            q = exp_noise.to(p.dtype)
        used = q
        pick = torch.topk(p / q, ns).indices
        nbr = torch.cat([knn_idx, rest_idx.gather(1, pick)], dim=-1)
    else:
        nbr = knn_idx
    return (nbr, used) if return_noise else nbr


# ----------------------------------------------------------------------------------------------
# network

def graph_norm(x, weight, bias, mean_scale, eps=1e-5):
    """torch_geometric 2.6.0 GraphNorm, single graph (call site src/models/egnn.py:74)."""
    o = x - x.mean(dim=0, keepdim=True) * mean_scale
    var = o.pow(2).mean(dim=0, keepdim=True)
    return weight * o / (var + eps).sqrt() + bias


def layer_norm(x, w, b, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


class OracleNet:
    """Score_Net (src/models/score_net_mlsb.py:251-425) over a plain {name: tensor} dict.

    `weights` keys are the checkpoint's state_dict keys without the leading "net.".
    """

    def __init__(self, weights: Dict[str, torch.Tensor], depth=6, cut_off=20.0):
        self.w = {k: v.detach().float().cpu() for k, v in weights.items()}
        self.depth = depth
        self.cut_off = float(cut_off)
        self.pos_width = self.w["positional_embed.weight"].shape[1]

    @classmethod
    def from_state_dict(cls, sd, depth=6, cut_off=20.0):
        return cls({k[4:] if k.startswith("net.") else k: v for k, v in sd.items()}, depth, cut_off)

    # -- E_GCL (src/models/egnn.py:95-159) ------------------------------------------------------
    def egcl(self, li, h, x, row, col, e_attr, lig_mask, last, keep=None):
        w = self.w
        p = "network.EGNN_%d.egcl." % li
        diff = x[row] - x[col]
        radial = (diff ** 2).sum(1, keepdim=True)
        diff = diff / (torch.sqrt(radial + 1e-8) + 1.0)          # normalize=True (egnn.py:144-146)
        u = F.linear(torch.cat([h[row], h[col], radial, e_attr], dim=1),
                     w[p + "edge_mlp.0.weight"], w[p + "edge_mlp.0.bias"])
        m = F.silu(F.linear(F.silu(u), w[p + "edge_mlp.2.weight"], w[p + "edge_mlp.2.bias"]))
        gate = torch.sigmoid(F.linear(m, w[p + "att_mlp.0.weight"], w[p + "att_mlp.0.bias"]))
        m = m * gate
        n = h.shape[0]
        if last:                                                  # update_coords (egnn.py:118-137)
            cw = F.linear(F.silu(F.linear(m, w[p + "coord_mlp.0.weight"], w[p + "coord_mlp.0.bias"])),
                          w[p + "coord_mlp.2.weight"])
            cw = cw.clamp(-2.0, 2.0)
            tr = diff * cw
            s = torch.zeros(n, 3).index_add_(0, row, tr)
            cnt = torch.zeros(n, 3).index_add_(0, row, torch.ones_like(tr))
            x = x + (s / cnt.clamp(min=1)) * lig_mask[:, None]
        agg = torch.zeros(n, m.shape[1]).index_add_(0, row, m)
        z = F.linear(torch.cat([h, agg], dim=1), w[p + "node_mlp.0.weight"], w[p + "node_mlp.0.bias"])
        z = graph_norm(z, w[p + "node_mlp.1.weight"], w[p + "node_mlp.1.bias"], w[p + "node_mlp.1.mean_scale"])
        h_new = h + F.linear(F.silu(z), w[p + "node_mlp.3.weight"], w[p + "node_mlp.3.bias"])
        if keep is not None:
            keep["u%d" % li] = u
            keep["agg%d" % li] = agg
            keep["h%d" % li] = h_new
        return h_new, x

    # -- Score_Net.forward (src/models/score_net_mlsb.py:343-425) -------------------------------
    def forward(self, batch, edges: Optional[torch.Tensor] = None, exp_noise=None, generator=None,
                keep: Optional[dict] = None):
        w = self.w
        rec_x, lig_x = batch["rec_x"].float(), batch["lig_x"].float()
        rec_pos, lig_pos = batch["rec_pos"].float(), batch["lig_pos"].float()
        t = batch["t"].float().reshape(-1)
        pm = batch["position_matrix"].float()
        n_rec, n_lig = rec_pos.shape[0], lig_pos.shape[0]

        centre = lig_pos[:, 1, :].mean(dim=0)
        rec_pos = rec_pos - centre
        lig_pos = lig_pos - centre
        pos = torch.cat([rec_pos, lig_pos], dim=0)
        ca = pos[:, 1, :]
        dcross = (rec_pos[:, None, 1, :] - lig_pos[None, :, 1, :]).norm(dim=-1)

        h = F.linear(torch.cat([rec_x, lig_x], dim=0), w["single_embed.weight"])
        pair = F.linear(spatial_matrix(pos), w["spatial_embed.weight"]) + F.linear(pm, w["positional_embed.weight"])

        if edges is None:
            nbr = knn_and_sample(ca, exp_noise=exp_noise, generator=generator)
        else:
            nbr = edges.long()
        n, k = nbr.shape
        row = torch.arange(n)[:, None].repeat(1, k).reshape(-1)
        col = nbr.reshape(-1)
        e_attr = pair[row, col]
        lig_mask = torch.zeros(n)
        lig_mask[n_rec:] = 1.0
        if keep is not None:
            keep["nbr"] = nbr
            keep["h_in"] = h

        x = ca
        for li in range(self.depth):
            h, x = self.egcl(li, h, x, row, col, e_attr, lig_mask, li == self.depth - 1, keep)

        ires = F.linear(F.silu(F.linear(F.silu(F.linear(h, w["to_ires.0.weight"], w["to_ires.0.bias"])),
                                        w["to_ires.2.weight"], w["to_ires.2.bias"])),
                        w["to_ires.4.weight"], w["to_ires.4.bias"])

        hr = h[:n_rec, None, :].expand(n_rec, n_lig, -1)
        hl = h[None, n_rec:, :].expand(n_rec, n_lig, -1)
        e = F.linear(torch.cat([hr, hl], dim=-1), w["to_energy.0.weight"])
        e = F.linear(F.silu(layer_norm(e, w["to_energy.1.weight"], w["to_energy.1.bias"])),
                     w["to_energy.3.weight"]).squeeze(-1)
        msk = (dcross < self.cut_off).float()
        energy = (e * msk).sum() / (msk.sum() + 1e-6)

        r = lig_pos[:, 1, :]
        f = x[n_rec:] - r
        tr_pred = f.mean(dim=0, keepdim=True)
        rot_pred = torch.cross(r, f, dim=-1).mean(dim=0, keepdim=True)

        proj = t[:, None] * w["t_embed.0.W"][None, :] * 2 * np.pi
        temb = torch.sigmoid(F.linear(torch.cat([torch.sin(proj), torch.cos(proj)], dim=-1), w["t_embed.1.weight"]))

        def scale(prefix, v):
            nrm = torch.linalg.vector_norm(v, keepdim=True)
            y = F.linear(torch.cat([nrm, temb], dim=-1), w[prefix + ".0.weight"])
            y = F.silu(layer_norm(y, w[prefix + ".1.weight"], w[prefix + ".1.bias"]))
            y = F.softplus(F.linear(y, w[prefix + ".4.weight"]))
            return v / (nrm + 1e-6) * y

        return {
            "tr_score": scale("tr_scale", tr_pred),
            "rot_score": scale("rot_scale", rot_pred),
            "energy": energy,
            "f": f,
            "num_clashes": torch.sum(dcross <= 3.0),
            "ires": ires,
        }

    __call__ = forward


# ----------------------------------------------------------------------------------------------
# SDE schedules (src/utils/so3_diffuser.py:210-227, src/utils/r3_diffuser.py:20-24); fp64 like numpy

def so3_sigma(t, lo=0.1, hi=1.5):
    return np.log(t * np.exp(hi) + (1 - t) * np.exp(lo))


def so3_g(t, lo=0.1, hi=1.5):
    s = so3_sigma(t, lo, hi)
    return np.sqrt(2 * (np.exp(hi) - np.exp(lo)) * s / np.exp(s))


def r3_g(t, lo=0.1, hi=30.0):
    return lo * (hi / lo) ** t * np.sqrt(2 * (np.log(hi) - np.log(lo)))


def reverse_increment(g_t, score, dt, z_scaled, ode=False):
    """torch_reverse (so3_diffuser.py:363-368 = r3_diffuser.py:50-55); dt a 0-d fp32 tensor.  ode=True is the
    probability-flow branch: half the drift, no noise."""
    if ode:
        return (0.5 * (g_t ** 2) * score * dt).float()
    return ((g_t ** 2) * score * dt + g_t * torch.sqrt(dt) * z_scaled).float()


# ----------------------------------------------------------------------------------------------
# rotations (src/utils/geometry.py:18-200), batch of one

def aa_to_quat(aa):
    ang = aa.norm(dim=-1, keepdim=True)
    half = 0.5 * ang
    small = ang.abs() < 1e-6
    k = torch.where(small, 0.5 - ang * ang / 48, torch.sin(half) / torch.where(small, torch.ones_like(ang), ang))
    return torch.cat([torch.cos(half), aa * k], dim=-1)


def quat_to_mat(q):
    r, i, j, k = torch.unbind(q, -1)
    s = 2.0 / (q * q).sum(-1)
    m = torch.stack((1 - s * (j * j + k * k), s * (i * j - k * r), s * (i * k + j * r),
                     s * (i * j + k * r), 1 - s * (i * i + k * k), s * (j * k - i * r),
                     s * (i * k - j * r), s * (j * k + i * r), 1 - s * (i * i + j * j)), -1)
    return m.reshape(q.shape[:-1] + (3, 3))


def aa_to_mat(aa):
    return quat_to_mat(aa_to_quat(aa))


def mat_to_quat(m):
    """geometry.py:64-123: four candidates, pick the best conditioned (denominator floored at 0.1)."""
    b = m.shape[:-2]
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(m.reshape(b + (9,)), dim=-1)
    qa = torch.stack([1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22,
                      1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22], dim=-1)
    qa = torch.sqrt(torch.clamp(qa, min=0))
    cand = torch.stack([
        torch.stack([qa[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=-1),
        torch.stack([m21 - m12, qa[..., 1] ** 2, m10 + m01, m02 + m20], dim=-1),
        torch.stack([m02 - m20, m10 + m01, qa[..., 2] ** 2, m12 + m21], dim=-1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, qa[..., 3] ** 2], dim=-1)], dim=-2)
    cand = cand / (2.0 * qa[..., None].clamp(min=0.1))
    best = qa.argmax(dim=-1)
    return torch.gather(cand, -2, best[..., None, None].expand(b + (1, 4))).squeeze(-2)


def quat_to_aa(q):
    nrm = q[..., 1:].norm(dim=-1, keepdim=True)
    half = torch.atan2(nrm, q[..., :1])
    ang = 2 * half
    small = ang.abs() < 1e-6
    k = torch.where(small, 0.5 - ang * ang / 48, torch.sin(half) / torch.where(small, torch.ones_like(ang), ang))
    return q[..., 1:] / k


def mat_to_aa(m):
    return quat_to_aa(mat_to_quat(m))


def rot_compose(r1, r2):
    """src/inference_base.py:311-316: R = R(r2) @ R(r1)."""
    return mat_to_aa(aa_to_mat(r2) @ aa_to_mat(r1))


# ----------------------------------------------------------------------------------------------
# pose helpers (src/inference_base.py:318-384; src/inference.py:220-286 for centre_mode=1)

def _centroid(x, centre_mode):
    return x[:, 1, :].mean(dim=0) if centre_mode == 0 else x.mean(dim=(0, 1))


def randomize_pose(rec_pos, lig_pos, rot0_matrix, tr0_normal, centre_mode=0):
    """rot0_matrix [3,3] (scipy Rotation.random().as_matrix()), tr0_normal [1,3] ~ N(0, 30^2)."""
    c1 = _centroid(rec_pos, centre_mode)
    c2 = _centroid(lig_pos, centre_mode)
    rot = rot0_matrix.float()
    tr = tr0_normal - c2 + c1
    x = (lig_pos - c2) @ rot.T + c2 + tr
    return x, tr, mat_to_aa(rot.unsqueeze(0))


def modify_coords(x, rot_aa, tr, centre_mode=0):
    c = _centroid(x, centre_mode)
    if centre_mode == 0:
        c = c[None, :]
    rot = aa_to_mat(rot_aa).squeeze()
    return (x - c) @ rot.T + c + tr


def clash_force(rec_pos, lig_pos):
    """Analytic gradient of the reference's autograd soft-clash term (inference_base.py:366-384).

    U = -5 * sum_{d<4} (4-d)^1.5 / (0.75 d);  returns mean over ligand atoms of dU/dx_lig.
    """
    a = rec_pos.reshape(-1, 3)
    b = lig_pos.reshape(-1, 3)
    diff = b[None, :, :] - a[:, None, :]
    d = diff.norm(dim=-1)
    g = 4.0 - d
    inside = d < 4.0
    gp = torch.where(inside, g, torch.zeros_like(g))
    dphi = -(1.5 * torch.sqrt(gp) * d + gp ** 1.5) / (0.75 * d * d)
    coef = torch.where(inside, -5.0 * dphi / d, torch.zeros_like(d))
    return (coef[..., None] * diff).sum(dim=0).mean(dim=0)


def clash_force_autograd(rec_pos, lig_pos):
    """The reference's own formulation (autograd), used to pin clash_force()."""
    a = rec_pos.reshape(-1, 3)
    b = lig_pos.reshape(-1, 3).clone().requires_grad_(True)
    with torch.enable_grad():
        d = (a[:, None, :] - b[None, :, :]).norm(dim=-1)
        rep = torch.where(d < 4, (torch.abs(4 - d) ** 1.5) / (1.5 * d * 0.5), torch.tensor(0.0))
        u = -5 * rep.sum()
        (grad,) = torch.autograd.grad(u, b)
    return grad.mean(dim=0).detach()


# ----------------------------------------------------------------------------------------------
# sampler (src/inference_base.py:390-468)

def euler_maruyama_sampler(net: OracleNet, batch, num_steps=40, eps=1e-3, use_clash_force=False,
                           noise_annealing=False, tr_noise_scale=0.5, rot_noise_scale=0.5,
                           centre_mode=0, noise: Optional[dict] = None, record: Optional[list] = None, ode=False):
    """Restatement of Euler_Maruyama_sampler.

    noise (all optional): {"rot0": [3,3], "tr0": [1,3], "edges": [S+1,N,K] or "exp": [S+1,N,N-20],
    "z_rot": [S,1,3], "z_tr": [S,1,3]}; anything absent is drawn from numpy/torch global RNGs in the
    reference's order (numpy normal(4) -> torch normal(1,3) -> per step {Exp, randn, randn}).
    """
    noise = noise or {}
    ts = torch.linspace(1.0, eps, num_steps)
    dt = ts[0] - ts[1]
    rec_pos = batch["rec_pos"].clone().float()
    lig_pos = batch["lig_pos"].clone().float()

    if "rot0" in noise:
        rot0 = noise["rot0"]
    else:
        from scipy.spatial.transform import Rotation
        rot0 = torch.from_numpy(Rotation.random().as_matrix()).float()
    tr0 = noise["tr0"] if "tr0" in noise else torch.normal(0.0, 30.0, size=(1, 3))
    lig_pos, tr_update, rot_update = randomize_pose(rec_pos, lig_pos, rot0, tr0, centre_mode)

    def fwd(i, t):
        b = dict(batch)
        b["t"] = t
        b["rec_pos"] = rec_pos
        b["lig_pos"] = lig_pos
        kw = {}
        if "edges" in noise:
            kw["edges"] = noise["edges"][i]
        elif "exp" in noise:
            kw["exp_noise"] = noise["exp"][i]
        return net.forward(b, **kw)

    out = None
    for i in range(num_steps):
        t = ts[i]
        last = i == num_steps - 1
        out = fwd(i, torch.ones(1) * t)
        if noise_annealing:
            ns_tr = ns_rot = float(t)
        elif last:
            ns_tr = ns_rot = 0.0
        else:
            ns_tr, ns_rot = tr_noise_scale, rot_noise_scale
        z_rot = noise["z_rot"][i] if "z_rot" in noise else torch.randn(1, 3)
        rot = reverse_increment(so3_g(float(t)), out["rot_score"], dt, ns_rot * z_rot, ode)
        z_tr = noise["z_tr"][i] if "z_tr" in noise else torch.randn(1, 3)
        tr = reverse_increment(r3_g(float(t)), out["tr_score"], dt, ns_tr * z_tr, ode)
        lig_pos = modify_coords(lig_pos, rot, tr, centre_mode)
        tr_update = tr_update + tr
        rot_update = rot_compose(rot_update, rot)
        if use_clash_force:
            cf = clash_force(rec_pos, lig_pos)
            lig_pos = lig_pos + cf
            tr_update = tr_update + cf
        if record is not None:
            record.append({"lig_pos": lig_pos.clone(), "rot": rot.clone(), "tr": tr.clone(),
                           "tr_score": out["tr_score"].clone(), "rot_score": out["rot_score"].clone()})
        if last:
            out = fwd(num_steps, torch.ones(1) * t)
    return rec_pos, lig_pos, rot_update, tr_update, out


# ----------------------------------------------------------------------------------------------
# all-atom output

def modify_aa_coords(x, bb_coords, rot_aa, tr, centre_mode=0):
    """modify_aa_coords: centre_mode 0 = src/inference_base.py:354-364 (CA centroid of the backbone),
    1 = src/inference.py:256-266 (centroid of the all-atom coordinates).  numpy float64 like the reference."""
    x = np.asarray(x)
    center = np.asarray(bb_coords, dtype=np.float64)[:, 1].mean(axis=0) if centre_mode == 0 else x.mean(axis=0)
    rot = aa_to_mat(torch.as_tensor(rot_aa).float().view(1, 3)).squeeze().cpu().numpy()
    return (x - center) @ rot.T + center + torch.as_tensor(tr).float().view(1, 3).cpu().numpy()
