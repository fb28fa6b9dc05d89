"""Simulation_Parameters (reference jaxent/src/interfaces/simulation.py:17-226).

Same six fields, same arithmetic and the same two normalisers.  Arrays are torch tensors (or numpy
arrays); `frame_weights` holds logits inside the optimiser and softmax weights in `save_state`,
exactly like the reference (SURVEY.md section 8a, quirk iii)."""
from __future__ import annotations

import operator
from dataclasses import dataclass
from typing import Any, Sequence

import numpy as np
import torch

from .._arrays import to_numpy


@dataclass(frozen=True, slots=True)
class Simulation_Parameters:
    frame_weights: Any
    frame_mask: Any
    model_parameters: Sequence[Any]
    forward_model_weights: Any
    normalise_loss_functions: Any
    forward_model_scaling: Any

    @staticmethod
    def normalize_masked_loss_scalingweights(params: "Simulation_Parameters") -> "Simulation_Parameters":
        """fw <- fw*(1-m) + (fw*m / sum(fw*m))*m, m = clip(normalise_loss_functions, 0, 1)   (:45-76).
        L-sized host arithmetic, done once at Simulation.initialise."""
        w = to_numpy(params.forward_model_weights)
        m = np.clip(to_numpy(params.normalise_loss_functions), 0, 1).astype(np.float32)
        masked = w * m
        result = w * (np.float32(1.0) - m) + (masked / masked.sum(dtype=np.float32)) * m
        return _replace(params, forward_model_weights=result.astype(np.float32))

    @staticmethod
    def normalize_weights(params: "Simulation_Parameters") -> "Simulation_Parameters":
        """frame_weights <- softmax(frame_weights); frame_mask <- clip(sigmoid(10 (mask - 0.5)), 0, 1);
        model parameters pass through unclamped (:103-142).  On CUDA tensors this is two torch
        elementwise ops used outside the optimise step (initialise / user code); the step itself gets its
        softmax from the fused kernels."""
        fw = params.frame_weights
        if isinstance(fw, torch.Tensor):
            if fw.dim() != 1 or tuple(fw.shape) != tuple(params.frame_mask.shape):
                raise ValueError("frame_weights and frame_mask must be rank-1 arrays of equal shape")
            w = torch.softmax(fw.float(), 0)
            mask = torch.clamp(torch.sigmoid(10 * (torch.as_tensor(params.frame_mask, device=fw.device).float() - 0.5)), 0, 1)
        else:
            z = to_numpy(fw)
            if z.ndim != 1 or z.shape != to_numpy(params.frame_mask).shape:
                raise ValueError("frame_weights and frame_mask must be rank-1 arrays of equal shape")
            e = np.exp(z - z.max())
            w = (e / e.sum(dtype=np.float32)).astype(np.float32)
            mask = np.clip(1.0 / (1.0 + np.exp(-10 * (to_numpy(params.frame_mask) - 0.5))), 0, 1).astype(np.float32)
        return _replace(params, frame_weights=w, frame_mask=mask)

    @staticmethod
    def propagate_model_parameters(params: "Simulation_Parameters", model_index: int = 0):
        mp = params.model_parameters[model_index]
        return _replace(params, model_parameters=[mp for _ in params.model_parameters])

    # ---- arithmetic over the six fields (:144-179)
    def _apply(self, op, other, reverse=False):
        if isinstance(other, Simulation_Parameters):
            pick = lambda name: getattr(other, name)
            mps = [op(a, b) if not reverse else op(b, a) for a, b in zip(self.model_parameters, other.model_parameters)]
        elif isinstance(other, (float, int)):
            pick = lambda name: other
            mps = [op(a, other) if not reverse else op(other, a) for a in self.model_parameters]
        else:
            raise TypeError("Unsupported type for operation")
        f = (lambda a, b: op(b, a)) if reverse else op
        return Simulation_Parameters(
            frame_weights=f(self.frame_weights, pick("frame_weights")),
            frame_mask=f(self.frame_mask, pick("frame_mask")),
            model_parameters=mps,
            forward_model_weights=f(self.forward_model_weights, pick("forward_model_weights")),
            normalise_loss_functions=f(self.normalise_loss_functions, pick("normalise_loss_functions")),
            forward_model_scaling=f(self.forward_model_scaling, pick("forward_model_scaling")),
        )

    def __add__(self, other):
        return self._apply(operator.add, other)

    def __sub__(self, other):
        return self._apply(operator.sub, other)

    def __mul__(self, other):
        return self._apply(operator.mul, other)

    def __truediv__(self, other):
        return self._apply(operator.truediv, other)

    __radd__ = __add__
    __rmul__ = __mul__

    def __rsub__(self, other):
        return self._apply(operator.sub, other, reverse=True)

    def __rtruediv__(self, other):
        return self._apply(operator.truediv, other, reverse=True)

    def tree_flatten(self):
        return (self.frame_weights, self.frame_mask, list(self.model_parameters), self.forward_model_weights,
                self.forward_model_scaling, self.normalise_loss_functions), ()

    @classmethod
    def tree_unflatten(cls, static, arrays):
        fw, fm, mp, fmw, fms, nlf = arrays
        return cls(fw, fm, mp, fmw, nlf, fms)


def _replace(p: Simulation_Parameters, **kw) -> Simulation_Parameters:
    d = {k: getattr(p, k) for k in ("frame_weights", "frame_mask", "model_parameters", "forward_model_weights",
                                    "normalise_loss_functions", "forward_model_scaling")}
    d.update(kw)
    return Simulation_Parameters(**d)
