"""Input records -> the batch dict Score_Model consumes (the part of get_batch_from_inputs that needs no ESM model).

Reference: src/inference_base.py:190-244 (one-hot + ESM concat, get_position_matrix, relpos :246-292) and
src/utils/residue_constants.py:855-929 (restype_order_with_x, sequence_to_onehot).  ESM-2 embeddings are taken
pre-computed from the record (data/db5_test/*.pt ship them); running ESM-2 itself is a "next" row.
"""
import torch

RESTYPES = "ARNDCQEGHILKMFPSTWYV"   # AlphaFold order; index 20 = unknown 'X'


def sequence_to_onehot(seq):
    idx = torch.tensor([RESTYPES.find(c) if c in RESTYPES else 20 for c in seq], dtype=torch.long)
    return torch.nn.functional.one_hot(idx, 21).float()


def relpos_bins(n_rec, n_lig):
    n = n_rec + n_lig
    idx = torch.arange(n)
    same = (idx[:, None] < n_rec) == (idx[None, :] < n_rec)
    return torch.where(same, (idx[:, None] - idx[None, :] + 32).clamp(0, 64), torch.full((n, n), 65))


def get_position_matrix(n_rec, n_lig, width=66, sym=0.0):
    pm = torch.nn.functional.one_hot(relpos_bins(n_rec, n_lig), 66).float()
    if width == 67:   # homomer channel the pinder checkpoint expects
        pm = torch.cat([pm, torch.full((n_rec + n_lig, n_rec + n_lig, 1), float(sym))], dim=-1)
    return pm


def batch_from_record(rec, pos_width=66, with_position_matrix=True):
    """rec = {"receptor": {x[n,1280], pos[n,3,3], seq}, "ligand": {...}} -> batch dict (CPU tensors).

    with_position_matrix=False leaves out the dense [N,N,P] one-hot (1.7 GB for the 2548-residue db5 complex): the CUDA
    library rebuilds relpos from (i, j, R) anyway and only needs the homomer flag, passed as batch["sym"]."""
    r, l = rec["receptor"], rec["ligand"]
    rec_x = torch.cat([r["x"].float(), sequence_to_onehot(r["seq"])], dim=-1)
    lig_x = torch.cat([l["x"].float(), sequence_to_onehot(l["seq"])], dim=-1)
    sym = 1.0 if r["seq"] == l["seq"] else 0.0
    batch = {"rec_x": rec_x, "lig_x": lig_x, "rec_pos": r["pos"].float(), "lig_pos": l["pos"].float(), "sym": sym}
    if with_position_matrix:
        batch["position_matrix"] = get_position_matrix(rec_x.shape[0], lig_x.shape[0], pos_width, sym)
    return batch


def synthetic_complex(n_rec, n_lig, seed=0, x_dim=1301, pos_width=66):
    """Synthetic two-chain complex: CA random walk with 3.8 A steps, N/C at 1.46/1.52 A, ligand +25 A in x."""
    g = torch.Generator().manual_seed(seed)

    def chain(n):
        steps = torch.randn(n, 3, generator=g)
        steps = 3.8 * steps / steps.norm(dim=-1, keepdim=True)
        ca = torch.cumsum(steps, dim=0)
        ca = ca - ca.mean(dim=0)
        u = torch.randn(n, 3, generator=g)
        v = torch.randn(n, 3, generator=g)
        n_at = ca + 1.46 * u / u.norm(dim=-1, keepdim=True)
        c_at = ca + 1.52 * v / v.norm(dim=-1, keepdim=True)
        return torch.stack([n_at, ca, c_at], dim=1)

    rec_pos, lig_pos = chain(n_rec), chain(n_lig)
    lig_pos = lig_pos + torch.tensor([25.0, 0.0, 0.0])
    return {
        "rec_x": torch.randn(n_rec, x_dim, generator=g), "lig_x": torch.randn(n_lig, x_dim, generator=g),
        "rec_pos": rec_pos, "lig_pos": lig_pos,
        "position_matrix": get_position_matrix(n_rec, n_lig, pos_width, 0.0),
    }
