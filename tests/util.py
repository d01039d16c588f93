"""Shared helpers for the tests: regenerate the seeded inputs the goldens were made from."""
import os

import torch

from dfmdock_b200.features import synthetic_complex
from dfmdock_b200.synthetic import synthetic_hparams, synthetic_state_dict

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)


# must mirror tests/golden/make_goldens.py main()
def case_small():
    b = synthetic_complex(40, 30, seed=3, pos_width=66)
    b["lig_pos"] = b["lig_pos"] - torch.tensor([17.0, 0.0, 0.0])
    return synthetic_state_dict(1, 66), synthetic_hparams(66), b


def case_mid_p67():
    b = synthetic_complex(75, 53, seed=4, pos_width=67)
    b["lig_pos"] = b["lig_pos"] - torch.tensor([15.0, 0.0, 0.0])
    return synthetic_state_dict(2, 67), synthetic_hparams(67), b


def case_tiny():
    b = synthetic_complex(25, 20, seed=5, pos_width=66)
    b["lig_pos"] = b["lig_pos"] - torch.tensor([18.0, 0.0, 0.0])
    return synthetic_state_dict(1, 66), synthetic_hparams(66), b


FWD_CASES = {
    "fwd_synth_n70.pt": case_small,
    "fwd_synth_n128_p67.pt": case_mid_p67,
    "fwd_synth_n45.pt": case_tiny,
}


def rel_err(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def max_abs(a, b):
    return float((a.double() - b.double()).abs().max())
