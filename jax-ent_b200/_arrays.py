"""Array plumbing shared by the facade: accept numpy / torch, keep things on the CUDA device."""
from __future__ import annotations

import numpy as np
import torch


def to_numpy(x, dtype=np.float32) -> np.ndarray:
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy().astype(dtype, copy=False)
    return np.asarray(x, dtype)


def scalar(x) -> float:
    if isinstance(x, torch.Tensor):
        return float(x.detach().reshape(-1)[0].item())
    return float(np.asarray(x).reshape(-1)[0])


def to_device(x, device, dtype=torch.float32) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype)
    return torch.as_tensor(np.asarray(x), dtype=dtype, device=device)
