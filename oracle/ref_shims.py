"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (dfmdock_b200/).

Import shims that let the *unmodified* reference (Graylab/DFMDock, mounted read-only
at /root/reference) be imported in this container, which lacks pytorch_lightning,
torch_geometric, omegaconf, hydra, fair-esm, biotite and dm-tree.

Used by:
  * tests/golden/make_goldens.py  -- generates the committed golden vectors
  * oracle/build_ref.py           -- extracts weights/fixtures into oracle/_ref/
  * tests (only when /root/reference exists) -- pins oracle/ against the live reference

/root/reference does not exist on the GPU box, so nothing that runs there may call
`install()`; callers must check `reference_available()` first.

Third-party arithmetic restated here (not under /root/reference):
  torch_geometric==2.6.0 (requirements.txt:50) `torch_geometric.nn.norm.GraphNorm`,
  call site src/models/egnn.py:6,74.  Published algorithm (Cai et al. 2021, "GraphNorm"),
  single-graph case:  o = x - mean_0(x) * mean_scale ;  y = weight * o / sqrt(mean_0(o^2) + eps) + bias
  with eps = 1e-5 and no running statistics (so eval == train).
"""
import importlib.machinery
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("DFMDOCK_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "models"))


def _mod(name):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__path__ = []
    sys.modules[name] = m
    return m


class GraphNorm(nn.Module):
    """Restatement of torch_geometric 2.6.0 GraphNorm for a single graph (batch=None)."""

    def __init__(self, in_channels, eps=1e-5):
        super().__init__()
        self.in_channels = in_channels
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(in_channels))
        self.bias = nn.Parameter(torch.zeros(in_channels))
        self.mean_scale = nn.Parameter(torch.ones(in_channels))

    def forward(self, x, batch=None, batch_size=None):
        mean = x.mean(dim=0, keepdim=True)
        out = x - mean * self.mean_scale
        var = out.pow(2).mean(dim=0, keepdim=True)
        std = (var + self.eps).sqrt()
        return self.weight * out / std + self.bias


class _State:
    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self.__dict__["_state"] = state


def unwrap_omegaconf(obj):
    """DictConfig pickle -> nested plain dict (values are AnyNode._val)."""
    d = getattr(obj, "__dict__", None)
    if d is not None and "_content" in d:
        c = d["_content"]
        if isinstance(c, dict):
            return {k: unwrap_omegaconf(v) for k, v in c.items()}
        if isinstance(c, list):
            return [unwrap_omegaconf(v) for v in c]
        return c
    if d is not None and "_val" in d:
        return d["_val"]
    if isinstance(obj, dict):
        return {k: unwrap_omegaconf(v) for k, v in obj.items()}
    return obj


class AttrDict(dict):
    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return AttrDict(v) if isinstance(v, dict) and not isinstance(v, AttrDict) else v

    def __setattr__(self, k, v):
        self[k] = v


_installed = False


def install():
    """Register the stub modules and put the reference's src/ on sys.path."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)

    oc = _mod("omegaconf")
    for sub, names in {
        "omegaconf.dictconfig": ["DictConfig"],
        "omegaconf.listconfig": ["ListConfig"],
        "omegaconf.base": ["ContainerMetadata", "Metadata"],
        "omegaconf.nodes": ["AnyNode", "IntegerNode", "FloatNode", "StringNode", "BooleanNode"],
    }.items():
        m = _mod(sub)
        for n in names:
            setattr(m, n, type(n, (_State,), {}))
    oc.DictConfig = sys.modules["omegaconf.dictconfig"].DictConfig
    oc.OmegaConf = type("OmegaConf", (), {})

    hy = _mod("hydra")
    hy.main = lambda *a, **k: (lambda f: f)
    hyu = _mod("hydra.utils")
    hyu.instantiate = lambda *a, **k: None

    pl = _mod("pytorch_lightning")

    class LightningModule(nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, *a, **k):
            pass

        @property
        def device(self):
            return next(self.parameters()).device

    pl.LightningModule = LightningModule
    pl.LightningDataModule = type("LightningDataModule", (), {})
    pl.Trainer = type("Trainer", (), {})
    pl.seed_everything = lambda *a, **k: None

    _mod("torch_geometric")
    _mod("torch_geometric.nn")
    tgn = _mod("torch_geometric.nn.norm")
    tgn.GraphNorm = GraphNorm
    tgl = _mod("torch_geometric.loader")
    tgl.DataLoader = type("DataLoader", (), {})
    tgd = _mod("torch_geometric.data")
    tgd.HeteroData = type("HeteroData", (_State,), {})
    tgd.Data = type("Data", (_State,), {})
    tgd.Dataset = type("Dataset", (), {})
    tgh = _mod("torch_geometric.data.hetero_data")
    tgh.HeteroData = tgd.HeteroData
    tgs = _mod("torch_geometric.data.storage")
    for n in ["NodeStorage", "BaseStorage", "EdgeStorage", "GlobalStorage"]:
        setattr(tgs, n, type(n, (_State,), {}))

    esm = _mod("esm")
    esm.pretrained = types.SimpleNamespace()
    _mod("biotite")
    _mod("biotite.structure")
    bio_io = _mod("biotite.structure.io")
    bio_pdb = _mod("biotite.structure.io.pdb")
    bio_pdb.PDBFile = type("PDBFile", (), {})
    sys.modules["biotite"].structure = sys.modules["biotite.structure"]
    sys.modules["biotite.structure"].io = bio_io
    bio_io.pdb = bio_pdb

    tree = _mod("tree")

    def map_structure(fn, *structs):
        s0 = structs[0]
        if isinstance(s0, (list, tuple)):
            return type(s0)(map_structure(fn, *xs) for xs in zip(*structs))
        if isinstance(s0, dict):
            return {k: map_structure(fn, *[s[k] for s in structs]) for k in s0}
        return fn(*structs)

    tree.map_structure = map_structure

    src = os.path.join(REFERENCE_ROOT, "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    _installed = True


def load_reference_ckpt(path):
    """torch.load a Lightning ckpt with the omegaconf stubs; returns (state_dict, hparams AttrDict)."""
    install()
    ck = torch.load(path, map_location="cpu", weights_only=False)
    hp = AttrDict({k: unwrap_omegaconf(v) for k, v in ck["hyper_parameters"].items()})
    return ck["state_dict"], hp


def build_reference_model(path, cache_dir="/tmp/dfmdock_so3_cache"):
    """The reference's own Score_Model with the checkpoint loaded strict=True."""
    install()
    from models.score_model_mlsb import Score_Model

    sd, hp = load_reference_ckpt(path)
    hp["diffuser"]["so3"]["cache_dir"] = cache_dir
    model = Score_Model(hp.model, hp.diffuser, hp.experiment)
    model.load_state_dict(sd, strict=True)
    model.eval()
    return model, hp


def load_db5_record(path):
    """data/db5_test/<id>.pt (PyG HeteroData pickle) -> dict of plain tensors."""
    install()
    obj = torch.load(path, map_location="cpu", weights_only=False)
    stores = obj.__dict__["_node_store_dict"]
    out = {}
    for key in ("receptor", "ligand"):
        mp = stores[key].__dict__["_mapping"]
        out[key] = {"x": mp["x"], "pos": mp["pos"], "seq": mp["seq"]}
    g = obj.__dict__.get("_global_store")
    if g is not None:
        out["name"] = g.__dict__.get("_mapping", {}).get("name")
    return out
