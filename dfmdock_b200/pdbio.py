"""Structure-file side of the entry points, without biotite.

Reference functions mirrored (argument meaning, return layout and quirks kept; biotite replaced by a small
fixed-column PDB reader/writer):
    get_info_from_pdb        src/inference_base.py:72-126   -> {"structure", "seq", "aa_coords", "bb_coords"}
    get_native               src/inference_base.py:128-188  -> [bb_coords of chain 1, bb_coords of chain 2, ...]
    combine_atom_arrays      src/inference_base.py:37-66
    modify_aa_coords         src/inference_base.py:354-364 (CA-centroid) / src/inference.py:256-266 (all-atom centroid)
                             -> dfm_transform_atoms (csrc/pose.cu), all poses of a complex in one launch
    PDBFile().set_structure(...).write(path)  (biotite)     -> write_pdb
    save_PDB / get_full_coords / place_fourth_atom  src/utils/pdb.py:32-88, src/inference_mlsb.py:70-89
                             -> write_trajectory_pdb (multi-MODEL N, CA, C, O, CB dump of a sampling trajectory)
    get_esm_rep              src/inference_base.py:294-306  -> EsmEmbedder (Hugging Face `transformers` EsmModel from a
                             LOCAL directory; fair-esm is not installed here and there is no network for weights)
"""
import os

import numpy as np
import torch

RESTYPE_3TO1 = {"ALA": "A", "ARG": "R", "ASN": "N", "ASP": "D", "CYS": "C", "GLN": "Q", "GLU": "E", "GLY": "G",
                "HIS": "H", "ILE": "I", "LEU": "L", "LYS": "K", "MET": "M", "PHE": "F", "PRO": "P", "SER": "S",
                "THR": "T", "TRP": "W", "TYR": "Y", "VAL": "V"}     # src/utils/residue_constants.py:932-959
AA_1TO3 = {v: k for k, v in RESTYPE_3TO1.items()}
AA_1TO3.update({"-": "GAP", "X": "URI"})                            # src/utils/pdb.py:4-27 (sic)


class Structure:
    """Minimal atom table (the subset of biotite's AtomArray the reference touches)."""
    FIELDS = ("atom_name", "res_name", "chain_id", "res_id", "ins_code", "element", "hetero")

    def __init__(self, coord, atom_name, res_name, chain_id, res_id, ins_code=None, element=None, hetero=None):
        n = len(atom_name)
        self.coord = np.asarray(coord, dtype=np.float32).reshape(n, 3)
        self.atom_name = np.asarray(atom_name, dtype="U6")
        self.res_name = np.asarray(res_name, dtype="U5")
        self.chain_id = np.asarray(chain_id, dtype="U4")
        self.res_id = np.asarray(res_id, dtype=np.int64)
        self.ins_code = np.asarray(ins_code if ins_code is not None else [""] * n, dtype="U1")
        self.element = np.asarray(element if element is not None else [guess_element(a) for a in atom_name], dtype="U2")
        self.hetero = np.asarray(hetero if hetero is not None else [False] * n, dtype=bool)

    def __len__(self):
        return len(self.atom_name)

    def __getitem__(self, mask):
        return Structure(self.coord[mask], self.atom_name[mask], self.res_name[mask], self.chain_id[mask],
                         self.res_id[mask], self.ins_code[mask], self.element[mask], self.hetero[mask])

    def copy(self):
        return self[np.ones(len(self), dtype=bool)]


def guess_element(atom_name):
    s = "".join(c for c in atom_name if c.isalpha())
    return s[:1].upper() if s else ""


def load_structure(path):
    """First model of a PDB file; alternate locations: the first altloc of each residue is kept (biotite's default)."""
    coord, name, resn, chain, resi, ins, elem, het = [], [], [], [], [], [], [], []
    first_alt = {}
    with open(path) as f:
        for line in f:
            rec = line[:6]
            if rec == "ENDMDL":
                break
            if rec not in ("ATOM  ", "HETATM"):
                continue
            alt = line[16]
            key = (line[21], line[22:26], line[26])
            if alt != " ":
                if first_alt.setdefault(key, alt) != alt:
                    continue
            coord.append((float(line[30:38]), float(line[38:46]), float(line[46:54])))
            an = line[12:16].strip()
            name.append(an)
            resn.append(line[17:20].strip())
            chain.append(line[21].strip())
            resi.append(int(line[22:26]))
            ins.append(line[26].strip())
            e = line[76:78].strip() if len(line) >= 78 else ""
            elem.append(e.upper() if e else guess_element(an))
            het.append(rec == "HETATM")
    if not name:
        raise ValueError("%s: no ATOM records" % path)
    return Structure(coord, name, resn, chain, resi, ins, elem, het)


def residue_starts(s):
    """Indices where a new residue begins (change of chain, residue number, insertion code or residue name)."""
    n = len(s)
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    change = (s.chain_id[1:] != s.chain_id[:-1]) | (s.res_id[1:] != s.res_id[:-1]) | (s.ins_code[1:] != s.ins_code[:-1]) | \
             (s.res_name[1:] != s.res_name[:-1])
    return np.concatenate([[0], np.nonzero(change)[0] + 1])


def _valid_backbone_mask(s):
    """The reference's residue filter, quirk included: atoms are grouped by residue NUMBER over the whole file
    (src/inference_base.py:88-100), a group is kept when it holds N, CA and C."""
    valid = np.zeros(len(s), dtype=bool)
    for rid in np.unique(s.res_id):
        m = s.res_id == rid
        if {"N", "CA", "C"}.issubset(set(s.atom_name[m].tolist())):
            valid[m] = True
    return valid


def _backbone(s, what):
    n, ca, c = s.coord[s.atom_name == "N"], s.coord[s.atom_name == "CA"], s.coord[s.atom_name == "C"]
    if not (len(n) == len(ca) == len(c)):
        raise ValueError("%s: N/CA/C atom counts differ (%d/%d/%d)" % (what, len(n), len(ca), len(c)))
    return np.stack([n, ca, c], axis=1).astype(np.float64)


def get_info_from_pdb(pdb_path):
    """src/inference_base.py:72-126.  "structure"/"aa_coords" keep every ATOM record, "seq"/"bb_coords" only the residues
    with a complete backbone -- exactly the reference's (inconsistent) pairing."""
    structure = load_structure(pdb_path)
    structure = structure[~structure.hetero]
    filtered = structure[_valid_backbone_mask(structure)]
    starts = residue_starts(filtered)
    seq = "".join(RESTYPE_3TO1.get(r, "X") for r in filtered.res_name[starts])
    bb = _backbone(filtered, pdb_path)
    if bb.shape[0] != len(seq):
        raise ValueError("%s: %d residues but %d backbone triplets" % (pdb_path, len(seq), bb.shape[0]))
    return {"structure": structure, "seq": seq, "aa_coords": structure.coord, "bb_coords": bb}


def get_native(pdb_path):
    """src/inference_base.py:128-188: backbone [n,3,3] float tensors per chain, chains in sorted order."""
    structure = load_structure(pdb_path)
    structure = structure[~structure.hetero]
    filtered = structure[_valid_backbone_mask(structure)]
    return [torch.from_numpy(_backbone(filtered[filtered.chain_id == c], pdb_path)).float() for c in np.unique(filtered.chain_id)]


def combine_atom_arrays(a, b):
    """src/inference_base.py:37-66: concatenation that keeps coordinates, element, atom/residue names, ids and chain ids."""
    if a.coord.shape[1] != 3 or b.coord.shape[1] != 3:
        raise ValueError("Both AtomArray objects must have 3D coordinates (Nx3 arrays)")
    return Structure(np.concatenate([a.coord, b.coord]), np.concatenate([a.atom_name, b.atom_name]),
                     np.concatenate([a.res_name, b.res_name]), np.concatenate([a.chain_id, b.chain_id]),
                     np.concatenate([a.res_id, b.res_id]), None, np.concatenate([a.element, b.element]), None)


def _pdb_atom_name(name, element):
    return name if (len(name) >= 4 or len(element) > 1) else " " + name


def write_pdb(path, s):
    """Fixed-column ATOM records, serial renumbered from 1, occupancy 1.00, B-factor 0.00 (what biotite's PDBFile writes
    for an AtomArray without those annotations)."""
    lines = []
    for i in range(len(s)):
        x, y, z = (float(v) for v in s.coord[i])
        lines.append("%-6s%5d %-4s %3s %1s%4d%1s   %8.3f%8.3f%8.3f%6.2f%6.2f          %2s  " % (
            "HETATM" if s.hetero[i] else "ATOM", i % 99999 + 1, _pdb_atom_name(str(s.atom_name[i]), str(s.element[i])),
            str(s.res_name[i])[:3], str(s.chain_id[i])[:1], (int(s.res_id[i]) - 1) % 9999 + 1 if s.res_id[i] > 0 else int(s.res_id[i]),
            str(s.ins_code[i])[:1], x, y, z, 1.0, 0.0, str(s.element[i]).rjust(2)))
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


def modify_aa_coords(x, bb_coords, rot, tr, centre_mode=0, device=None):
    """All-atom coordinates of the docked ligand for T poses in one launch.

    x [A,3], bb_coords [L,3,3] (used when centre_mode == 0), rot [T,3] or [1,3] axis-angle, tr likewise -> [T,A,3]
    (torch, on `device`).  centre_mode 0 = src/inference_base.py:354-364, 1 = src/inference.py:256-266.
    """
    from . import _lib
    if device is None:
        device = rot.device if isinstance(rot, torch.Tensor) and rot.is_cuda else torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("dfmdock_b200.pdbio.modify_aa_coords has no CPU path; use a CUDA (sm_100a) device")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    xa = torch.as_tensor(np.asarray(x) if not isinstance(x, torch.Tensor) else x).to(device, torch.float32).contiguous().view(-1, 3)
    rot = torch.as_tensor(rot).to(device, torch.float32).contiguous().view(-1, 3)
    tr = torch.as_tensor(tr).to(device, torch.float32).contiguous().view(-1, 3)
    if rot.shape != tr.shape:
        raise ValueError("modify_aa_coords: rot %s and tr %s differ" % (tuple(rot.shape), tuple(tr.shape)))
    bb = None
    L = 0
    if centre_mode == 0:
        bb = torch.as_tensor(np.asarray(bb_coords) if not isinstance(bb_coords, torch.Tensor) else bb_coords)
        bb = bb.to(device, torch.float32).contiguous().view(-1, 3, 3)
        L = bb.shape[0]
    T, A = rot.shape[0], xa.shape[0]
    out = torch.empty(T, A, 3, device=device)
    with torch.cuda.device(device):
        _lib.check(_lib.load().dfm_transform_atoms(device.index, T, A, L, int(centre_mode), _lib.ptr(xa), _lib.ptr(bb),
                                                   _lib.ptr(rot), _lib.ptr(tr), _lib.ptr(out),
                                                   torch.cuda.current_stream(device).cuda_stream), "dfm_transform_atoms")
    return out


def write_complex_pdb(path, receptor, ligand, lig_aa_coords):
    """src/inference_base.py:502-513: receptor structure + ligand structure with the transformed coordinates."""
    lig = ligand["structure"].copy()
    lig.coord = np.asarray(torch.as_tensor(lig_aa_coords).detach().cpu(), dtype=np.float32).reshape(-1, 3)
    write_pdb(path, combine_atom_arrays(receptor["structure"], lig))


# ---- trajectory dumps (src/inference_mlsb.py:70-89, 126-149; src/utils/pdb.py:32-88) -------------------------------
def place_fourth_atom(a, b, c, length, planar, dihedral):
    bc = b - c
    bc = bc / bc.norm(dim=-1, keepdim=True)
    n = torch.linalg.cross((b - a).expand(bc.shape), bc, dim=-1)
    n = n / n.norm(dim=-1, keepdim=True)
    m = [bc, torch.linalg.cross(n, bc, dim=-1), n]
    d = [length * torch.cos(planar), length * torch.sin(planar) * torch.cos(dihedral), -length * torch.sin(planar) * torch.sin(dihedral)]
    return c + sum(mi * di for mi, di in zip(m, d))


def get_full_coords(coords):
    """[n,3,3] (N, CA, C) -> [n,5,3] (N, CA, C, O, CB) with ideal O / CB placement (src/inference_mlsb.py:70-89)."""
    coords = torch.as_tensor(coords).detach().cpu().float()
    N, CA, C = coords[:, 0], coords[:, 1], coords[:, 2]
    b, c = CA - N, C - CA
    a = torch.linalg.cross(b, c, dim=-1)
    CB = -0.58273431 * a + 0.56802827 * b - 0.54067466 * c + CA
    O = place_fourth_atom(torch.roll(N, -1, 0), CA, C, torch.tensor(1.231), torch.tensor(2.108), torch.tensor(-3.142))
    return torch.stack([N, CA, C, O, CB], dim=1)


def save_PDB(out_pdb, coords, seq, b_factors=None, delim=None):
    """src/utils/pdb.py:52-88 (appends; chain A up to residue index `delim`, chain B after)."""
    if delim is None:
        delim = -1
    if b_factors is None:
        b_factors = torch.zeros(coords.shape[0])
    atoms = ["N", "CA", "C", "O", "CB"]
    with open(out_pdb, "a") as f:
        k = 0
        for r, residue in enumerate(coords):
            AA = AA_1TO3[seq[r]]
            for a, atom in enumerate(residue):
                if AA == "GLY" and atoms[a] == "CB":
                    continue
                x, y, z = (float(v) for v in atom)
                f.write("ATOM  %5d  %-2s  %3s %s%4d    %8.3f%8.3f%8.3f  %4.2f %4.2f\n"
                        % (k + 1, atoms[a], AA, "A" if r <= delim else "B", r + 1, x, y, z, 1, float(b_factors[r])))
                k += 1


def write_trajectory_pdb(out_pdb, rec_pos, lig_frames, rec_seq, lig_seq):
    """Sampler.save_trj (src/inference_mlsb.py:126-149): one MODEL per recorded step, receptor + ligand frame."""
    if os.path.exists(out_pdb):
        os.remove(out_pdb)
    rec = torch.as_tensor(rec_pos).detach().cpu().float()
    for i, lig in enumerate(lig_frames):
        coords = get_full_coords(torch.cat([rec, torch.as_tensor(lig).detach().cpu().float()], dim=0))
        assert len(rec_seq) + len(lig_seq) == coords.shape[0]
        with open(out_pdb, "a") as f:
            f.write("MODEL        " + str(i) + "\n")
        save_PDB(out_pdb, coords, rec_seq + lig_seq, delim=len(rec_seq) - 1)


# ---- ESM-2 embedding front end (src/inference_base.py:294-306, 547-549) ----------------------------------------------
class EsmEmbedder:
    """get_esm_rep with Hugging Face `transformers` EsmModel weights from a local directory (the HF export of
    esm2_t33_650M_UR50D; representations[33][0, 1:-1] == last_hidden_state[0, 1:-1]).  Library code by design: the
    language model runs once per chain, outside the sampler hot path."""

    def __init__(self, model_dir, device="cuda"):
        if not model_dir or not os.path.isdir(model_dir):
            raise FileNotFoundError(
                "ESM-2 weights directory %r not found.  Raw PDB input needs the esm2_t33_650M_UR50D weights in Hugging Face "
                "format on local disk (--esm_dir); there is no network here.  Alternatively pass the reference's "
                "pre-embedded records (.pt with x / pos / seq per chain)." % (model_dir,))
        from transformers import AutoTokenizer, EsmModel
        self.tokenizer = AutoTokenizer.from_pretrained(model_dir)
        self.model = EsmModel.from_pretrained(model_dir, add_pooling_layer=False).to(device).eval()
        self.device = device

    def __call__(self, seq):
        tok = self.tokenizer([seq], return_tensors="pt", add_special_tokens=True)
        with torch.no_grad():
            out = self.model(**{k: v.to(self.device) for k, v in tok.items()}).last_hidden_state
        return out[0, 1:-1, :].float().cpu()


def embed_record(rec, embedder):
    """Adds the ESM-2 rows "x" to the chains of a record that does not have them yet (in place)."""
    for key in ("receptor", "ligand"):
        info = rec[key]
        if info.get("x") is None:
            x = embedder(info["seq"])
            if x.shape[0] != len(info["seq"]):
                raise ValueError("%s: embedder returned %d rows for %d residues" % (key, x.shape[0], len(info["seq"])))
            info["x"] = x
    return rec


def record_from_pdbs(pdb_1, pdb_2, embedder, id=None):
    """Two raw PDB files -> the input record batch_from_record consumes (+ the structures for all-atom output).
    embedder=None parses only (sizes, sequences, structures); embed_record() completes the record later."""
    out = {"id": id or os.path.splitext(os.path.basename(str(pdb_1)))[0]}
    for key, path in (("receptor", pdb_1), ("ligand", pdb_2)):
        info = get_info_from_pdb(path)
        info.update({"x": None, "pos": torch.from_numpy(info["bb_coords"]).float()})
        out[key] = info
    return embed_record(out, embedder) if embedder is not None else out
