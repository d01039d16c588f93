"""Entry points of the docking sampler, mirroring the reference's CLI surface on top of the CUDA library.

Reference entry points mirrored:
    src/inference.py:569-596           argparse surface (--paths | --csv, --ckpt, --out_dir, --out_csv_dir, --out_csv,
                                       --num_samples, --num_steps, --tr_noise_scale, --rot_noise_scale,
                                       --use_clash_force, --noise_annealing, --seed)
    src/inference.py:375-413, 418-495  run(args, model, inputs, batch, device) / main(args): per-sample rows of the
                                       metrics CSV (id, index, ..., energy, num_clashes), one structure file per sample
    src/inference.py:500-567           inference(in_1, in_2): pinder_0.ckpt, 40 samples x 40 steps, clash force,
                                       all-atom-centroid rotation convention, keep the lowest-energy pose -> output.pdb
    src/inference_base.py:601-670      inference(in_1, in_2): dips ckpt, 120 x 40, CA-centroid convention (variant="base")
    src/inference_single.py:1-12       `python -m dfmdock_b200.inference_single a b`

What differs, and why:
  * inputs are either the reference's own pre-embedded records (data/db5_test/<id>.pt: backbone + ESM-2 embeddings +
    sequence) or raw PDB files.  PDB files go through dfmdock_b200.pdbio (biotite-free get_info_from_pdb) and need ESM-2
    650M embeddings: `--esm_dir` points at a LOCAL Hugging Face export of esm2_t33_650M_UR50D (fair-esm is absent and
    there is no network here); without it a `.pdb` argument fails loudly -- nothing is faked.
  * all trajectories of a complex advance in lock step on the GPU (sample_trajectories) instead of the reference's
    serial loop; `--reference_rng` restores the serial loop with the reference's RNG consumption order.
  * the metric columns (c_rmsd, i_rmsd, l_rmsd, fnat, DockQ: src/utils/metrics.py) come from one batched CUDA launch
    over all poses (dfmdock_b200.metrics) instead of one torch SVD per sample.
  * PDB inputs are written back all-atom: modify_aa_coords for every sample in one launch (dfm_transform_atoms) +
    combine_atom_arrays + a fixed-column writer (pdbio.write_pdb).  Pre-embedded records carry no side chains, so they are
    written backbone-only (N, CA, C of both chains); rot_update / tr_update are stored in the CSV either way.
  * under torchrun (WORLD_SIZE > 1) the trajectories of every complex are sharded over the ranks (one all-gather of the
    [T,8] result rows per complex); rank 0 writes the files.
  * extras of the class-form sampler (src/inference_mlsb.py): --ode, --out_trj_dir (multi-MODEL trajectory dumps),
    --get_gt_energy (score the input pose at t = 1e-5 instead of sampling).
  * `confidence_logits` (src/inference.py:397) does not exist in the reference's own network output.
"""
import argparse
import csv
import os
import random

import numpy as np
import torch

from . import pdbio
from .checkpoint import load_db5_record
from .features import batch_from_record
from .metrics import KEYS as METRIC_KEYS, compute_metrics_batch
from .sampler import Euler_Maruyama_sampler, sample_complex_set, sample_trajectories
from .score_model import Score_Model

THREE = {"A": "ALA", "R": "ARG", "N": "ASN", "D": "ASP", "C": "CYS", "Q": "GLN", "E": "GLU", "G": "GLY", "H": "HIS",
         "I": "ILE", "L": "LEU", "K": "LYS", "M": "MET", "F": "PHE", "P": "PRO", "S": "SER", "T": "THR", "W": "TRP",
         "Y": "TYR", "V": "VAL"}



def complex_seed(seed, complex_id):
    """Philox key of one complex: the run's seed mixed with a stable hash of the complex id, so that trajectory k of
    different complexes does not reuse the same initial pose and noise (the reference advances ONE global RNG through
    all complexes, src/inference_base.py:644-657); trajectory k of a complex is still Philox subsequence k."""
    import zlib
    return (int(seed) ^ (zlib.crc32(str(complex_id).encode()) << 20)) & 0xFFFFFFFFFFFFFFFF


def set_seed(seed):
    """src/inference_base.py:24-32"""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


STRUCTURE_EXT = (".pdb", ".ent")


def load_inputs(path_1, path_2=None, id=None, embedder=None, parse_only=False):
    """-> inputs dict {"id", "receptor": {x, pos, seq[, structure, aa_coords, bb_coords]}, "ligand": {...}}.

    path_1 alone: a two-chain record (data/db5_test/<id>.pt layout).  path_1 + path_2: one single-chain record each
    ({"x", "pos", "seq"}), or two raw PDB files (src/inference_base.py:563-575) -- those need `embedder`, a callable
    seq -> [len(seq), 1280] ESM-2 representation (pdbio.EsmEmbedder).
    """
    is_pdb = [p is not None and str(p).lower().endswith(STRUCTURE_EXT) for p in (path_1, path_2)]
    if any(is_pdb):
        if not all(is_pdb):
            raise ValueError("give either two PDB files or pre-embedded records, not a mix: %s, %s" % (path_1, path_2))
        if parse_only:
            return pdbio.record_from_pdbs(path_1, path_2, None, id=id)
        if embedder is None:
            raise RuntimeError(
                "%s: raw structure files need ESM-2 650M embeddings (src/inference_base.py:294-306).  Pass --esm_dir with a local "
                "Hugging Face export of esm2_t33_650M_UR50D, or use the reference's pre-embedded records (.pt with x / pos / "
                "seq per chain)." % path_1)
        return pdbio.record_from_pdbs(path_1, path_2, embedder, id=id)
    if path_2 is None or path_2 == path_1:
        rec = load_db5_record(path_1)
    else:
        a = torch.load(path_1, map_location="cpu", weights_only=False)
        b = torch.load(path_2, map_location="cpu", weights_only=False)
        rec = {"receptor": a.get("receptor", a), "ligand": b.get("ligand", b)}
    rec["id"] = id or rec.get("name") or os.path.splitext(os.path.basename(str(path_1)))[0]
    return rec


def write_backbone_pdb(path, rec_pos, lig_pos, rec_seq, lig_seq):
    """Backbone-only PDB (chain A = receptor, chain B = ligand; atoms N, CA, C)."""
    lines = []
    serial = 1
    for chain, pos, seq in (("A", rec_pos, rec_seq), ("B", lig_pos, lig_seq)):
        pos = torch.as_tensor(pos).detach().cpu().double()
        for r in range(pos.shape[0]):
            res = THREE.get(seq[r] if r < len(seq) else "X", "UNK")
            for a, name in enumerate(("N", "CA", "C")):
                x, y, z = (float(v) for v in pos[r, a])
                lines.append("ATOM  %5d  %-3s %3s %1s%4d    %8.3f%8.3f%8.3f  1.00  0.00           %1s" %
                             (serial % 100000, name, res, chain, (r + 1) % 10000, x, y, z, name[0]))
                serial += 1
        lines.append("TER")
    lines.append("END")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


def _rmsd(a, b):
    return float(((a - b) ** 2).sum(-1).mean().sqrt())


def ligand_rmsd(lig_pos, native_lig_pos):
    """CA RMSD of the docked ligand to its pose in the input record with the receptor frame fixed (the receptor never
    moves in the sampler); a quick CA-only figure -- the reference's metric set is dfmdock_b200.metrics."""
    return _rmsd(torch.as_tensor(lig_pos)[:, 1].double().cpu(), torch.as_tensor(native_lig_pos)[:, 1].double().cpu())


def _rank():
    return torch.distributed.get_rank() if torch.distributed.is_available() and torch.distributed.is_initialized() else 0


def write_samples(out_dir, inputs, batch, poses, centre_mode, device):
    """One structure file per sample (src/inference.py:401-412): all-atom when the inputs came from PDB files."""
    if "structure" in inputs["ligand"]:
        rot = torch.stack([p[1] for p in poses]).to(device)
        tr = torch.stack([p[2] for p in poses]).to(device)
        aa = pdbio.modify_aa_coords(inputs["ligand"]["aa_coords"], inputs["ligand"]["bb_coords"], rot, tr,
                                    centre_mode=centre_mode, device=device).cpu()
        for i in range(len(poses)):
            pdbio.write_complex_pdb(os.path.join(out_dir, "%s_%d.pdb" % (inputs["id"], i)), inputs["receptor"], inputs["ligand"], aa[i])
    else:
        for i, p in enumerate(poses):
            write_backbone_pdb(os.path.join(out_dir, "%s_%d.pdb" % (inputs["id"], i)), batch["rec_pos"], p[0],
                               inputs["receptor"]["seq"], inputs["ligand"]["seq"])


def run(args, model, inputs, batch, device):
    """Per-sample metric rows + one structure file per sample (src/inference.py:375-413)."""
    centre_mode = int(getattr(args, "centre_mode", 1))
    ode = bool(getattr(args, "ode", False))
    trj_dir = getattr(args, "out_trj_dir", None)
    rows = []
    if getattr(args, "native_dir", None):
        nat = pdbio.get_native(os.path.join(args.native_dir, "%s.pdb" % inputs["id"]))     # src/inference_base.py:478-479
        native_rec, native_lig = nat[0], nat[1]
    else:
        native_rec, native_lig = batch["rec_pos"], batch["lig_pos"].clone()
    if getattr(args, "get_gt_energy", False):
        # Sampler.run_sampling, get_gt_energy branch (src/inference_mlsb.py:190-199): score the input pose, no sampling
        energy, clashes = model.gt_energy(batch)
        met = compute_metrics_batch(batch["rec_pos"], batch["lig_pos"][None], native_rec, native_lig, device=device).cpu()
        row = {"id": inputs["id"]}
        row.update({k: float(met[0, j]) for j, k in enumerate(METRIC_KEYS)})
        row.update({"energy": energy, "num_clashes": clashes})
        return [row]
    frames = None
    if getattr(args, "reference_rng", False):
        poses, frames = [], []
        for i in range(args.num_samples):
            trj = [] if trj_dir else None
            rec_pos, lig_pos, rot_update, tr_update, output = Euler_Maruyama_sampler(
                model=model, batch=dict(batch), num_steps=args.num_steps, device=device,
                use_clash_force=args.use_clash_force, noise_annealing=args.noise_annealing,
                tr_noise_scale=args.tr_noise_scale, rot_noise_scale=args.rot_noise_scale, centre_mode=centre_mode, ode=ode,
                trajectory=trj)
            poses.append((lig_pos.cpu(), rot_update.view(3).cpu(), tr_update.view(3).cpu(), float(output["energy"]),
                          int(output["num_clashes"])))
            frames.append(trj)
    elif trj_dir:
        model.set_complex(batch)
        res = model.sample(batch["lig_pos"], args.num_samples, num_steps=args.num_steps, tr_noise_scale=args.tr_noise_scale,
                           rot_noise_scale=args.rot_noise_scale, use_clash_force=args.use_clash_force,
                           noise_annealing=args.noise_annealing, centre_mode=centre_mode, seed=complex_seed(args.seed, inputs["id"]), ode=ode, record=True)
        poses = [(res["lig_pos"][i].cpu(), res["rot_update"][i].cpu(), res["tr_update"][i].cpu(), float(res["energy"][i]),
                  int(res["num_clashes"][i])) for i in range(args.num_samples)]
        frames = [list(res["frames"][:, i].cpu()) for i in range(args.num_samples)]
    else:
        res = sample_trajectories(model, batch, args.num_samples, num_steps=args.num_steps,
                                  use_clash_force=args.use_clash_force, noise_annealing=args.noise_annealing,
                                  tr_noise_scale=args.tr_noise_scale, rot_noise_scale=args.rot_noise_scale,
                                  centre_mode=centre_mode, seed=complex_seed(args.seed, inputs["id"]), gather_poses=True, ode=ode)
        poses = [(res["lig_pos"][i].cpu(), res["rot_update"][i].cpu(), res["tr_update"][i].cpu(), float(res["energy"][i]),
                  int(res["num_clashes"][i])) for i in range(args.num_samples)]
    return finish_samples(args, inputs, batch["rec_pos"], native_rec, native_lig, poses, frames, centre_mode, device)


def finish_samples(args, inputs, rec_pos, native_rec, native_lig, poses, frames, centre_mode, device):
    """Metric rows of every sample + (rank 0) the structure / trajectory files."""
    trj_dir = getattr(args, "out_trj_dir", None)
    rows = []
    # metrics of every pose in one launch (src/inference.py:393 calls compute_metrics per sample); the native pose is the
    # one in the input record unless --native_dir is given, like the reference's `native = inputs[...]['bb_coords']`
    met = compute_metrics_batch(rec_pos, torch.stack([p[0] for p in poses]), native_rec, native_lig, device=device).cpu()
    for i, (lig_pos, rot_u, tr_u, energy, clashes) in enumerate(poses):
        row = {"id": inputs["id"], "index": str(i)}
        row.update({k: (round(float(met[i, j]), 6) if k == "fnat" else float(met[i, j])) for j, k in enumerate(METRIC_KEYS)})
        row.update({"energy": energy, "num_clashes": clashes, "rot_update": " ".join("%.6f" % float(v) for v in rot_u),
                    "tr_update": " ".join("%.6f" % float(v) for v in tr_u)})
        rows.append(row)
    if _rank() == 0:
        if getattr(args, "out_dir", None):
            write_samples(args.out_dir, inputs, {"rec_pos": rec_pos}, poses, centre_mode, device)
        if trj_dir and frames:
            os.makedirs(trj_dir, exist_ok=True)
            for i, trj in enumerate(frames):
                pdbio.write_trajectory_pdb(os.path.join(trj_dir, "%s_p%d.pdb" % (inputs["id"], i)), rec_pos, trj,
                                           inputs["receptor"]["seq"], inputs["ligand"]["seq"])
    return rows


def run_planned(args, model, paths_list, embedder, device):
    """Many complexes under torchrun (BASELINE config #5): (complex, trajectory) work items planned over the ranks
    (dfmdock_b200.distributed.plan_work), one object all-gather at the end, rank 0 computes the metrics and writes."""
    centre_mode = int(getattr(args, "centre_mode", 1))
    light = [load_inputs(p1, p2, id=id, parse_only=True) for id, p1, p2 in paths_list]
    sizes = [(len(r["receptor"]["seq"]), len(r["ligand"]["seq"])) for r in light]      # (R, L): the planner's cost model uses both

    def loader(c):
        def load():
            rec = light[c]
            if rec["receptor"].get("x") is None or rec["ligand"].get("x") is None:
                if embedder is None:
                    raise RuntimeError("%s: raw structure files need ESM-2 embeddings: pass --esm_dir" % paths_list[c][1])
                pdbio.embed_record(rec, embedder)
            return batch_from_record(rec, pos_width=model.pos_width, with_position_matrix=False)
        return load

    results, plan = sample_complex_set(
        model, [loader(c) for c in range(len(light))], sizes, args.num_samples, num_steps=args.num_steps,
        use_clash_force=args.use_clash_force, noise_annealing=args.noise_annealing, tr_noise_scale=args.tr_noise_scale,
        rot_noise_scale=args.rot_noise_scale, centre_mode=centre_mode, seed=args.seed, ode=bool(getattr(args, "ode", False)),
        seeds=[complex_seed(args.seed, id) for id, _, _ in paths_list])
    rows = []
    if _rank() == 0:
        for c, res in enumerate(results):
            inputs = light[c]
            rec_pos, lig0 = inputs["receptor"]["pos"].float(), inputs["ligand"]["pos"].float()
            if getattr(args, "native_dir", None):
                nat = pdbio.get_native(os.path.join(args.native_dir, "%s.pdb" % inputs["id"]))
                native_rec, native_lig = nat[0], nat[1]
            else:
                native_rec, native_lig = rec_pos, lig0
            poses = [(res["lig_pos"][i], res["rot_update"][i], res["tr_update"][i], float(res["energy"][i]),
                      int(res["num_clashes"][i])) for i in range(args.num_samples)]
            rows.extend(finish_samples(args, inputs, rec_pos, native_rec, native_lig, poses, None, centre_mode, device))
    return rows


def init_distributed():
    """One process per GPU under torchrun: NCCL group + device from LOCAL_RANK.  No-op for a plain `python` launch."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and torch.distributed.is_available() and not torch.distributed.is_initialized():
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world


def main(args, embedder=None):
    """src/inference.py:418-495.  `embedder` overrides the ESM-2 front end (seq -> [n,1280]); default: --esm_dir."""
    paths_list = []
    if args.paths:
        id, p1, p2 = args.paths
        if os.path.exists(p1) and os.path.exists(p2):
            paths_list.append((id, p1, p2))
        else:
            print("One or both paths do not exist.")
    elif args.csv:
        if os.path.exists(args.csv):
            with open(args.csv, "r") as f:
                for row in csv.reader(f):
                    if row:
                        paths_list.append((row[0], row[1], row[2] if len(row) > 2 else row[1]))
        else:
            print("CSV file does not exist.")
    os.makedirs(args.out_dir, exist_ok=True)
    if not torch.cuda.is_available():
        raise RuntimeError("dfmdock_b200.inference needs a CUDA (sm_100a) device; there is no CPU path")
    init_distributed()
    device = torch.device("cuda", torch.cuda.current_device())
    model = Score_Model.load_from_checkpoint(args.ckpt, map_location=device)
    model.to(device).eval()
    if embedder is None and getattr(args, "esm_dir", None):
        embedder = pdbio.EsmEmbedder(args.esm_dir, device=device)
    world = torch.distributed.get_world_size() if torch.distributed.is_available() and torch.distributed.is_initialized() else 1
    planned = (world > 1 and len(paths_list) > 1 and not getattr(args, "reference_rng", False)
               and not getattr(args, "out_trj_dir", None) and not getattr(args, "get_gt_energy", False))
    results = run_planned(args, model, paths_list, embedder, device) if planned else []
    for id, p1, p2 in ([] if planned else paths_list):
        inputs = load_inputs(p1, p2, id=id, embedder=embedder)
        batch = batch_from_record(inputs, pos_width=model.pos_width)
        results.extend(run(args, model, inputs, batch, device))
    os.makedirs(args.out_csv_dir, exist_ok=True)
    out = os.path.join(args.out_csv_dir, args.out_csv)
    if results and _rank() == 0:
        with open(out, "w", newline="") as f:
            w = csv.DictWriter(f, fieldnames=list(results[0].keys()))
            w.writeheader()
            for row in results:
                w.writerow(row)
    return results


def inference(in_1, in_2=None, ckpt=None, variant="pinder", num_samples=None, num_steps=40, out="output.pdb", seed=None,
              embedder=None, esm_dir=None):
    """Dock one complex and write the lowest-energy pose (src/inference.py:500-567; variant="base":
    src/inference_base.py:601-670).  Returns {"energy", "rot_update", "tr_update", "lig_pos", "index"}."""
    if variant == "pinder":
        ckpt = ckpt or "./weights/pinder_0.ckpt"
        num_samples = num_samples or 40
        use_clash_force, centre_mode = True, 1
    else:
        ckpt = ckpt or "./checkpoints/dips/model_0.ckpt"
        num_samples = num_samples or 120
        use_clash_force, centre_mode = False, 0
    if not torch.cuda.is_available():
        raise RuntimeError("dfmdock_b200.inference needs a CUDA (sm_100a) device; there is no CPU path")
    device = torch.device("cuda", torch.cuda.current_device())
    model = Score_Model.load_from_checkpoint(ckpt, map_location=device)
    model.to(device).eval()
    if embedder is None and esm_dir:
        embedder = pdbio.EsmEmbedder(esm_dir, device=device)
    inputs = load_inputs(in_1, in_2, embedder=embedder)
    batch = batch_from_record(inputs, pos_width=model.pos_width)
    res = sample_trajectories(model, batch, num_samples, num_steps=num_steps, use_clash_force=use_clash_force,
                              centre_mode=centre_mode, seed=0 if seed is None else seed, gather_poses=True)
    best = res["best"]
    lig = res["lig_pos"][best].cpu()
    if out and "structure" in inputs["ligand"]:
        # src/inference.py:558-567 / src/inference_base.py:661-670: best pose, all-atom
        aa = pdbio.modify_aa_coords(inputs["ligand"]["aa_coords"], inputs["ligand"]["bb_coords"], res["rot_update"][best][None],
                                    res["tr_update"][best][None], centre_mode=centre_mode, device=device)[0]
        pdbio.write_complex_pdb(out, inputs["receptor"], inputs["ligand"], aa)
    elif out:
        write_backbone_pdb(out, batch["rec_pos"], lig, inputs["receptor"]["seq"], inputs["ligand"]["seq"])
    return {"energy": float(res["energy"][best]), "rot_update": res["rot_update"][best].cpu(),
            "tr_update": res["tr_update"][best].cpu(), "lig_pos": lig, "index": best}


def build_parser():
    """Same flags as src/inference.py:569-589 (+ --reference_rng, --centre_mode)."""
    parser = argparse.ArgumentParser(description="DFMDock reverse-diffusion docking sampler (B200 CUDA path)")
    group = parser.add_mutually_exclusive_group(required=True)
    group.add_argument("--paths", nargs=3, metavar=("id", "in_1", "in_2"), help="id and two input records")
    group.add_argument("--csv", type=str, help="CSV file with rows: id, in_1, in_2")
    parser.add_argument("--ckpt", type=str, default="../checkpoints/dips/model_0.ckpt")
    parser.add_argument("--out_dir", type=str, default="./pdbs")
    parser.add_argument("--out_csv_dir", type=str, default="./csv_files")
    parser.add_argument("--out_csv", type=str, default="./test.csv")
    parser.add_argument("--num_samples", type=int, default=1)
    parser.add_argument("--num_steps", type=int, default=40)
    parser.add_argument("--tr_noise_scale", type=float, default=0.5)
    parser.add_argument("--rot_noise_scale", type=float, default=0.5)   # the reference declares type=int here (App. D.5)
    parser.add_argument("--use_clash_force", action="store_true")
    parser.add_argument("--noise_annealing", action="store_true")
    parser.add_argument("--seed", type=int, default=42)
    parser.add_argument("--reference_rng", action="store_true",
                        help="serial trajectories with the reference's RNG consumption order instead of batched Philox")
    parser.add_argument("--centre_mode", type=int, default=1, choices=[0, 1],
                        help="1: rotate about the N/CA/C centroid (src/inference.py), 0: CA centroid (src/inference_base.py)")
    parser.add_argument("--native_dir", type=str, default=None, help="<native_dir>/<id>.pdb = native complex (src/inference_base.py:478)")
    parser.add_argument("--esm_dir", type=str, default=None,
                        help="local Hugging Face export of esm2_t33_650M_UR50D; required for raw PDB inputs")
    parser.add_argument("--ode", action="store_true", help="probability-flow ODE instead of the reverse SDE (inference_mlsb.py)")
    parser.add_argument("--out_trj_dir", type=str, default=None, help="write every sample's trajectory as a multi-MODEL PDB")
    parser.add_argument("--get_gt_energy", action="store_true", help="score the input pose at t=1e-5 instead of sampling")
    return parser


if __name__ == "__main__":
    _args = build_parser().parse_args()
    set_seed(_args.seed)
    main(_args)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()
