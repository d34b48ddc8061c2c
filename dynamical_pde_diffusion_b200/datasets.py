"""Validation data from the reference's HDF5 files -> the loader ``evaluation.test_loop`` consumes.

Mirrors ``ValidationDataset`` and ``get_validation_dataloader`` of the reference (``src/diffusion_pde/datasets/dataset.py:169-238,
309-339``): the file holds ``U`` (N, C, H, W, T), ``t_steps`` (T,) and optionally ``labels`` (N,) / (N, label_dim) (layout written by
``pdes/utils.py:70-127``); every (sample, target time) pair becomes one observation ``{"A": U[..., 0], "U": U[..., t], "labels"}``,
yielded with a leading batch dimension of 1 like the reference's ``DataLoader(batch_size=1, collate_fn=collate_optional)``.
The file is decoded by :mod:`.hdf5` (h5py is not a dependency).
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

from .hdf5 import H5File

__all__ = ["ValidationDataset", "get_validation_dataloader", "read_validation_file"]


class ValidationDataset(torch.utils.data.Dataset):
    """``dataset.py:169-238``: items ``{"A": (C,H,W), "U": (C,H,W), "labels": (label_dim,) | None}``; item ``n * T' + k`` pairs
    sample n's initial state with its state at target time k (``T' = T`` with ``include_t0_as_target`` else ``T - 1``)."""

    def __init__(self, data, t_steps, labels=None, time_as_label: bool = False, include_t0_as_target: bool = False):
        data = torch.as_tensor(np.asarray(data)).float()             # (N, C, H, W, T)
        t_steps = torch.as_tensor(np.asarray(t_steps)).float()       # (T,)
        labels = torch.as_tensor(np.asarray(labels)).float() if labels is not None else None
        N, C, H, W, T = data.shape
        if len(t_steps) != T:
            raise ValueError(f"Length of t_steps ({len(t_steps)}) must match the last dimension of data ({T})")
        if len(t_steps) < 2:
            raise ValueError(f"t_steps must contain at least 2 time steps, but got {len(t_steps)}")
        first = 0 if include_t0_as_target else 1
        Tt = T - first
        A = data[..., 0].repeat_interleave(Tt, dim=0)                # (N*T', C, H, W)
        U = data[..., first:].permute(0, 4, 1, 2, 3).reshape(N * Tt, C, H, W)
        self.data = torch.cat((A, U), dim=1)                         # (N*T', 2C, H, W)
        self.labels = None
        if labels is not None:
            if labels.ndim == 1:
                labels = labels.reshape((-1, 1))
            labels = labels.repeat_interleave(Tt, dim=0)
            self.labels = torch.cat((t_steps[first:].repeat(N).unsqueeze(1), labels), dim=1) if time_as_label else labels
        self.N, self.C = N * Tt, C

    def __len__(self) -> int:
        return self.N

    def __getitem__(self, idx):
        return {"A": self.data[idx, :self.C], "U": self.data[idx, self.C:], "labels": self.labels[idx] if self.labels is not None else None}


def collate_optional(batch):
    """``dataset.py:240-248``: stack every key, keeping ``None`` entries ``None``."""
    return {k: (torch.stack([item[k] for item in batch], dim=0) if batch[0][k] is not None else None) for k in batch[0]}


def read_validation_file(path):
    """``(U, t_steps, labels | None, attrs)`` of one data file (``dataset.py:331-334``); ``attrs`` carries ``dx``, ``T`` ... ."""
    with H5File(Path(path)) as f:
        data, t_steps = f["U"][:], f["t_steps"][:]
        labels = f["labels"][:] if "labels" in f else None
        return data, t_steps, labels, dict(f.attrs)


def get_validation_dataloader(data_path, time_as_label: bool, include_t0_as_target: bool) -> torch.utils.data.DataLoader:
    """``dataset.py:309-339`` (without the repository-root lookup: ``data_path`` is used as given)."""
    data, t_steps, labels, _ = read_validation_file(data_path)
    valset = ValidationDataset(data, t_steps, labels=labels, time_as_label=time_as_label, include_t0_as_target=include_t0_as_target)
    return torch.utils.data.DataLoader(valset, batch_size=1, shuffle=False, collate_fn=collate_optional)
