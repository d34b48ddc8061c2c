"""dpde_b200 -- B200-native physics-guided diffusion sampler step (drop-in for the reference's sampling API).

Importing the package does not need a GPU; calling any op does (there is no CPU fallback).
"""
from .ops import GuidanceEngine, LLGConstants, heat_loss2, laplacian, llg_loss2, llg_residual_loss  # noqa: F401
from .sampler import JointSampler, Sampler, UnconditionalSampler, X_and_dXdt, X_and_dXdt_dummy, X_and_dXdt_fd, X_and_dXdt_fd_batched, sampling_context  # noqa: F401

from .training import EDMHeatLoss, heat_residual_sq  # noqa: F401,E402

__all__ = ["EDMHeatLoss", "heat_residual_sq", "GuidanceEngine", "LLGConstants", "heat_loss2", "laplacian", "llg_loss2", "llg_residual_loss",
           "JointSampler", "Sampler", "UnconditionalSampler", "X_and_dXdt", "X_and_dXdt_dummy", "X_and_dXdt_fd", "X_and_dXdt_fd_batched", "sampling_context"]
