"""Docking quality metrics on the GPU: host-side mirror of the reference's src/utils/metrics.py over the C ABI.

    compute_metrics(model, native) -> {"c_rmsd", "i_rmsd", "l_rmsd", "fnat", "DockQ"}     (reference signature, :3-16)
    compute_metrics_batch(model_rec, model_lig[T], native_rec, native_lig) -> [T, 5]      (all poses of a complex at once)

All compute is dfm_compute_metrics (csrc/metrics.cu); there is no CPU path.
"""
import ctypes

import torch

from . import _lib

KEYS = ("c_rmsd", "i_rmsd", "l_rmsd", "fnat", "DockQ")


def compute_metrics_batch(model_rec, model_lig, native_rec, native_lig, device=None):
    """model_rec [R,3,3] (shared) or [T,R,3,3]; model_lig [T,L,3,3] -> float32 tensor [T,5] on the device."""
    if device is None:
        device = model_lig.device if model_lig.is_cuda else torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("dfmdock_b200.metrics has no CPU path; use a CUDA (sm_100a) device")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    f = lambda t: t.to(device, torch.float32).contiguous()
    model_lig = f(model_lig)
    if model_lig.dim() == 3:
        model_lig = model_lig[None]
    T, L = model_lig.shape[0], model_lig.shape[1]
    native_rec, native_lig, model_rec = f(native_rec), f(native_lig), f(model_rec)
    R = native_rec.shape[0]
    shared = model_rec.dim() == 3
    if tuple(native_lig.shape) != (L, 3, 3) or tuple(native_rec.shape) != (R, 3, 3) or \
            tuple(model_rec.shape) != ((R, 3, 3) if shared else (T, R, 3, 3)):
        raise ValueError("compute_metrics: inconsistent shapes %s %s %s %s" % (tuple(model_rec.shape), tuple(model_lig.shape),
                                                                              tuple(native_rec.shape), tuple(native_lig.shape)))
    lib = _lib.load()
    nws = lib.dfm_metrics_workspace_bytes(R, L)
    ws = torch.empty(nws, dtype=torch.uint8, device=device)
    out = torch.empty(T, 5, device=device)
    with torch.cuda.device(device):
        stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        _lib.check(lib.dfm_compute_metrics(device.index, T, R, L, _lib.ptr(model_rec), 1 if shared else 0, _lib.ptr(model_lig),
                                           _lib.ptr(native_rec), _lib.ptr(native_lig), _lib.ptr(out), _lib.ptr(ws), nws, stream),
                   "dfm_compute_metrics")
    return out


def compute_metrics(model, native):
    """Reference signature (src/utils/metrics.py:3): model = (rec [R,3,3], lig [L,3,3]), native likewise -> dict of floats."""
    o = compute_metrics_batch(model[0].squeeze(), model[1].squeeze()[None], native[0].squeeze(), native[1].squeeze())[0].cpu()
    d = {k: float(o[i]) for i, k in enumerate(KEYS)}
    d["fnat"] = round(d["fnat"], 6)
    return d
