"""Gradient masks (reference jaxent/src/opt/gradients.py:8-130).

On the CUDA path the masks are two scalars of the device control block (`Ctl.mask_frame`, `Ctl.mask_model`: frame-only
before `initial_steps`, the configured set afterwards -- `opt/optimiser.py:365-381`) and are applied inside the pass /
tail kernels; nothing in a step calls these functions.  They exist for code that builds or inspects mask pytrees the way
the reference does (same names, return types and error behaviour), on host arrays or torch tensors."""
from __future__ import annotations

from typing import Any, Optional

import numpy as np

from ..custom_types.config import Optimisable_Parameters
from ..interfaces.simulation import Simulation_Parameters


def _full_like(x, value: float):
    try:
        import torch

        if isinstance(x, torch.Tensor):
            return torch.full_like(x, value, dtype=torch.float32)
    except ImportError:  # pragma: no cover
        pass
    return np.full_like(np.asarray(x), value, dtype=np.float32)


def _vdot(a, b) -> float:
    try:
        import torch

        if isinstance(a, torch.Tensor) or isinstance(b, torch.Tensor):
            return float((torch.as_tensor(a).double().flatten().cpu() * torch.as_tensor(b).double().flatten().cpu()).sum())
    except ImportError:  # pragma: no cover
        pass
    return float(np.vdot(np.asarray(a, np.float64), np.asarray(b, np.float64)))


def _leaves(params: Simulation_Parameters):
    out = [params.frame_weights, params.frame_mask]
    for mp in params.model_parameters:
        out.extend(mp.tree_flatten()[0])
    out.extend([params.forward_model_weights, params.forward_model_scaling, params.normalise_loss_functions])
    return out


def get_previous_grads(opt_state: Any) -> Any:
    """:8-13 -- zeros shaped like the parameters when no gradient has been stored yet"""
    if opt_state.gradients is not None:
        return opt_state.gradients
    p = opt_state.params
    return Simulation_Parameters(
        frame_weights=_full_like(p.frame_weights, 0.0), frame_mask=_full_like(p.frame_mask, 0.0),
        model_parameters=[mp._map(lambda s, v: _full_like(v, 0.0)) for mp in p.model_parameters],
        forward_model_weights=_full_like(p.forward_model_weights, 0.0), normalise_loss_functions=_full_like(p.normalise_loss_functions, 0.0),
        forward_model_scaling=_full_like(p.forward_model_scaling, 0.0))


def check_gradient_oscillation(prev_grads: Any, current_grads: Any) -> float:
    """:15-26 -- sum over all leaves of vdot(previous, current); 0.0 without current gradients"""
    if current_grads is None:
        return 0.0
    return sum(_vdot(a, b) for a, b in zip(_leaves(prev_grads), _leaves(current_grads)))


def create_gradient_masks(parameter_partition_masks, params: Simulation_Parameters, optimisable_funcs: Optional[Any] = None) -> Simulation_Parameters:
    """:28-93 -- float 0/1 masks: frame weights, model parameters; the loss-weight leaves are always zero"""
    frame = 1.0 if Optimisable_Parameters.frame_weights in parameter_partition_masks else 0.0
    model = 1.0 if Optimisable_Parameters.model_parameters in parameter_partition_masks else 0.0
    if Optimisable_Parameters.frame_mask in parameter_partition_masks:
        raise NotImplementedError(
            "Frame mask optimization not fully implemented - while gradients can flow, "
            "frame masking is not applied during weights normalisation before the forward step")
    return Simulation_Parameters(
        frame_weights=_full_like(params.frame_weights, frame), frame_mask=_full_like(params.frame_mask, 0.0),
        model_parameters=[mp._map(lambda s, v: _full_like(v, model)) for mp in params.model_parameters],
        normalise_loss_functions=_full_like(params.normalise_loss_functions, 0.0),
        forward_model_weights=_full_like(params.forward_model_weights, 0.0), forward_model_scaling=_full_like(params.forward_model_scaling, 0.0))


def mask_gradients(grads: Simulation_Parameters, masks: Simulation_Parameters) -> Simulation_Parameters:
    """:96-130 -- leaf-wise product; the loss-weight leaves are REPLACED by the (zero) masks"""
    return Simulation_Parameters(
        frame_weights=grads.frame_weights * masks.frame_weights, frame_mask=grads.frame_mask * masks.frame_mask,
        model_parameters=[g._map(lambda s, v, m=m: v * getattr(m, s)) for g, m in zip(grads.model_parameters, masks.model_parameters)],
        normalise_loss_functions=masks.normalise_loss_functions, forward_model_weights=masks.forward_model_weights,
        forward_model_scaling=masks.forward_model_scaling)
