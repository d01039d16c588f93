"""Reads the reference's Lightning checkpoints (weights/pinder_0.ckpt, checkpoints/dips/model_0.ckpt) unchanged.

File layout: torch zip-pickle with `state_dict` (104 fp32 tensors, prefix "net.") and
`hyper_parameters` = {model, diffuser, experiment} pickled as omegaconf DictConfig objects.  omegaconf /
pytorch_lightning are not needed: the unpickler below maps the omegaconf classes to inert shells and the
nested configs are unwrapped into plain dicts.  Replaces Score_Model.load_from_checkpoint
(reference src/inference_base.py:552-557, 611-616).
"""
import importlib.machinery
import sys
import types

import torch

DEFAULT_HPARAMS = {
    "model": {"lm_embed_dim": 1301, "positional_embed_dim": 66, "spatial_embed_dim": 100, "node_dim": 256,
              "edge_dim": 128, "inner_dim": 128, "depth": 6, "dropout": 0.1, "cut_off": 20.0, "normalize": True},
    "diffuser": {"r3": {"min_sigma": 0.1, "max_sigma": 30.0, "schedule": "VE"},
                 "so3": {"num_omega": 1000, "num_sigma": 1000, "min_sigma": 0.1, "max_sigma": 1.5,
                         "schedule": "logarithmic", "cache_dir": ".cache/", "use_cached_score": False}},
    "experiment": {},
}


class _Shell:
    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self.__dict__["_state"] = state


_STUBS = {
    "omegaconf": ["DictConfig", "ListConfig", "OmegaConf"],
    "omegaconf.dictconfig": ["DictConfig"],
    "omegaconf.listconfig": ["ListConfig"],
    "omegaconf.base": ["ContainerMetadata", "Metadata"],
    "omegaconf.nodes": ["AnyNode", "IntegerNode", "FloatNode", "StringNode", "BooleanNode"],
    # db5_test records are torch_geometric HeteroData pickles
    "torch_geometric": [],
    "torch_geometric.data": ["HeteroData", "Data"],
    "torch_geometric.data.hetero_data": ["HeteroData"],
    "torch_geometric.data.data": ["Data", "DataEdgeAttr", "DataTensorAttr"],
    "torch_geometric.data.storage": ["NodeStorage", "BaseStorage", "EdgeStorage", "GlobalStorage"],
}


class _stub_modules:
    """Temporarily provide the modules the pickles reference, unless the real ones are importable."""

    def __enter__(self):
        self.added = []
        for name, classes in _STUBS.items():
            if name in sys.modules:
                continue
            m = types.ModuleType(name)
            m.__spec__ = importlib.machinery.ModuleSpec(name, None)
            m.__path__ = []
            for c in classes:
                setattr(m, c, type(c, (_Shell,), {}))
            sys.modules[name] = m
            self.added.append(name)
        return self

    def __exit__(self, *exc):
        for name in self.added:
            sys.modules.pop(name, None)


def _unwrap(obj):
    d = getattr(obj, "__dict__", None)
    if d is not None and "_content" in d:
        c = d["_content"]
        if isinstance(c, dict):
            return {k: _unwrap(v) for k, v in c.items()}
        if isinstance(c, (list, tuple)):
            return [_unwrap(v) for v in c]
        return c
    if d is not None and "_val" in d:
        return d["_val"]
    if isinstance(obj, dict):
        return {k: _unwrap(v) for k, v in obj.items()}
    return obj


def load_checkpoint(path, map_location="cpu"):
    """-> (state_dict without the "net." prefix, hparams dict {model, diffuser, experiment})."""
    with _stub_modules():
        ck = torch.load(path, map_location=map_location, weights_only=False)
    if "state_dict" not in ck:
        raise KeyError("%s: not a Lightning checkpoint (no 'state_dict')" % path)
    sd = {}
    for k, v in ck["state_dict"].items():
        sd[k[4:] if k.startswith("net.") else k] = v
    hp = {k: _unwrap(v) for k, v in ck.get("hyper_parameters", {}).items()}
    for k, v in DEFAULT_HPARAMS.items():
        hp.setdefault(k, v)
    return sd, hp


def load_db5_record(path):
    """data/db5_test/<id>.pt -> {"receptor": {x, pos, seq}, "ligand": {...}, "name"} (reference datasets/ppi_dataset.py layout)."""
    with _stub_modules():
        obj = torch.load(path, map_location="cpu", weights_only=False)
    if isinstance(obj, dict) and "receptor" in obj:
        return obj
    stores = obj.__dict__["_node_store_dict"]
    out = {}
    for key in ("receptor", "ligand"):
        mp = stores[key].__dict__["_mapping"]
        out[key] = {"x": mp["x"], "pos": mp["pos"], "seq": mp["seq"]}
    g = obj.__dict__.get("_global_store")
    out["name"] = g.__dict__.get("_mapping", {}).get("name") if g is not None else None
    return out
