"""`python -m dfmdock_b200.inference_single in_1 [in_2]` -- mirrors src/inference_single.py:1-12 (which calls
inference_base.inference: dips checkpoint, 120 trajectories x 40 steps, lowest-energy pose -> output.pdb)."""
import argparse

from .inference import inference


def parse_args():
    parser = argparse.ArgumentParser(description="Dock two chains: two PDB files (+ --esm_dir) or the reference's pre-embedded records.")
    parser.add_argument("pdb_1", type=str, help="receptor record (or a two-chain record)")
    parser.add_argument("pdb_2", type=str, nargs="?", default=None, help="ligand record")
    parser.add_argument("--ckpt", type=str, default=None)
    parser.add_argument("--esm_dir", type=str, default=None,
                        help="local Hugging Face export of esm2_t33_650M_UR50D; required when the inputs are raw PDB files")
    return parser.parse_args()


if __name__ == "__main__":
    args = parse_args()
    r = inference(args.pdb_1, args.pdb_2, ckpt=args.ckpt, variant="base", esm_dir=args.esm_dir)
    print("lowest energy %.4f (trajectory %d) -> output.pdb" % (r["energy"], r["index"]))
