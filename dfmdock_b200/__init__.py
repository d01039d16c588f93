"""dfmdock_b200: B200-native (sm_100a) implementation of DFMDock's reverse-diffusion docking sampler.

Public surface mirrors the reference's hot path:
    Score_Model.load_from_checkpoint / .to / .eval / __call__(batch) / .so3_diffuser / .r3_diffuser
    Euler_Maruyama_sampler(model, batch, ...)
plus batched / sharded sampling (sample_trajectories).  All compute is in libdfmdock_b200.so; no CPU fallback.
"""
from .score_model import Score_Model
from .sampler import Euler_Maruyama_sampler, sample_trajectories
from .features import batch_from_record, get_position_matrix, synthetic_complex
from .metrics import compute_metrics, compute_metrics_batch

__all__ = ["Score_Model", "Euler_Maruyama_sampler", "sample_trajectories", "batch_from_record",
           "get_position_matrix", "synthetic_complex", "compute_metrics", "compute_metrics_batch"]
