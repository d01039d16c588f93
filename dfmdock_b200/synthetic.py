"""Seeded random weights of the shipped architecture (bench.py runs on them: no checkpoint travels to the GPU box).

The reference initialises with std 0.02 (score_net_mlsb.py:332-341), which makes every activation vanish; these
weights use fan-in scaling so that all stages of the network carry O(1) signal, which is what parity tests need.
Deterministic given (seed, pos_width): torch's CPU generator is bit-reproducible across machines.
"""
import math

import torch

HPARAMS = {
    "model": {"lm_embed_dim": 1301, "positional_embed_dim": 66, "spatial_embed_dim": 100, "node_dim": 256,
              "edge_dim": 128, "inner_dim": 128, "depth": 6, "dropout": 0.1, "cut_off": 20.0, "normalize": True},
    "diffuser": {"r3": {"min_sigma": 0.1, "max_sigma": 30.0, "schedule": "VE"},
                 "so3": {"num_omega": 1000, "num_sigma": 1000, "min_sigma": 0.1, "max_sigma": 1.5,
                         "schedule": "logarithmic", "cache_dir": ".cache/", "use_cached_score": False}},
    "experiment": {"lr": 1e-4, "weight_decay": 0.0, "perturb_tr": True, "perturb_rot": True,
                   "separate_energy_loss": True, "separate_tr_loss": True, "separate_rot_loss": True,
                   "use_interface_loss": True, "grad_energy": False, "use_contrastive_loss": False},
}


def synthetic_hparams(pos_width=66):
    import copy
    hp = copy.deepcopy(HPARAMS)
    hp["model"]["positional_embed_dim"] = pos_width
    return hp


def synthetic_state_dict(seed=0, pos_width=66, x_dim=1301):
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def lin(name, out_f, in_f, bias=True, gain=1.0):
        sd[name + ".weight"] = torch.randn(out_f, in_f, generator=g) * (gain / math.sqrt(in_f))
        if bias:
            sd[name + ".bias"] = torch.randn(out_f, generator=g) * 0.1

    def norm(name, n, extra=()):
        sd[name + ".weight"] = 1.0 + 0.1 * torch.randn(n, generator=g)
        sd[name + ".bias"] = 0.1 * torch.randn(n, generator=g)
        for e in extra:
            sd[name + "." + e] = 1.0 + 0.1 * torch.randn(n, generator=g)

    lin("single_embed", 256, x_dim, bias=False)
    lin("spatial_embed", 128, 100, bias=False, gain=5.0)
    lin("positional_embed", 128, pos_width, bias=False, gain=4.0)
    for l in range(6):
        p = "network.EGNN_%d.egcl." % l
        lin(p + "edge_mlp.0", 256, 641)
        sd[p + "edge_mlp.0.weight"][:, 512] = 0.01 * torch.randn(256, generator=g)   # radial column (radial ~ 1e2..1e3 A^2)
        lin(p + "edge_mlp.2", 256, 256)
        lin(p + "node_mlp.0", 256, 512, gain=0.5)
        norm(p + "node_mlp.1", 256, extra=("mean_scale",))
        lin(p + "node_mlp.3", 256, 256, gain=0.7)
        if l == 5:
            lin(p + "coord_mlp.0", 256, 256)
            lin(p + "coord_mlp.2", 1, 256, bias=False, gain=2.0)
        lin(p + "att_mlp.0", 1, 256)
    lin("to_energy.0", 256, 512, bias=False)
    norm("to_energy.1", 256)
    lin("to_energy.3", 1, 256, bias=False, gain=3.0)
    lin("to_ires.0", 512, 256)
    lin("to_ires.2", 512, 512)
    lin("to_ires.4", 1, 512)
    sd["t_embed.0.W"] = torch.randn(64, generator=g)
    lin("t_embed.1", 128, 128, bias=False)
    for pre in ("tr_scale", "rot_scale"):
        lin(pre + ".0", 128, 129, bias=False)
        norm(pre + ".1", 128)
        lin(pre + ".4", 1, 128, bias=False)
    return sd


def write_lightning_ckpt(path, state_dict, hparams):
    """Minimal Lightning-layout checkpoint (plain-dict hyper_parameters), readable by checkpoint.load_checkpoint."""
    torch.save({"epoch": 0, "global_step": 0, "pytorch-lightning_version": "2.4.0",
                "state_dict": {"net." + k: v for k, v in state_dict.items()},
                "hyper_parameters": hparams}, path)
