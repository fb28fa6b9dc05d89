"""Convergence tracking (reference jaxent/src/opt/track.py): LR-scaled descending thresholds (:12-21), the functional
tracker of the pure loop (:24-125: initialise_convergence_carry / update_convergence / check_and_advance_threshold, what
`_optimise_pure` and therefore batch_optimise use) and the ConvergenceTracker class of the Python loop (:128-197).  Host-side
scalars only: on the B200 path the same rules run inside the tail kernels (csrc/bv_tail.cu derive_rest, csrc/bg_step.cu
bt_decide); these mirrors drive the host loops and state what the device code has to reproduce."""
from __future__ import annotations

from typing import Any, Optional

import numpy as np


def create_convergence_thresholds(convergence, learning_rate) -> list:
    thr = np.atleast_1d(np.asarray(convergence, np.float32))
    return [float(t) for t in (np.sort(thr)[::-1] * np.float32(learning_rate))]


def initialise_convergence_carry(initial_params):
    """reference opt/track.py:24-32"""
    from .base import ConvergenceCarry

    return ConvergenceCarry(np.float32(0.0), initial_params, 0, 0, False)


def get_relative_convergence(carry, current_loss):
    current_loss = np.float32(current_loss)
    return np.float32(carry.ema_loss_delta / current_loss) if current_loss > 0 else np.float32(0.0)


def update_convergence(carry, previous_loss, current_loss, current_params, ema_alpha):
    """reference opt/track.py:40-78: the EMAs RESTART whenever steps_since_threshold_start == 0.  `current_params` may be
    None (EMA parameters not tracked by the caller) or anything supporting a * x + b * y."""
    raw = np.abs(np.float32(previous_loss) - np.float32(current_loss))
    first = carry.steps_since_threshold_start == 0
    a = np.float32(ema_alpha)
    ema = raw if first else np.float32(a * raw + (np.float32(1.0) - a) * carry.ema_loss_delta)
    if current_params is None or carry.ema_params is None:
        ema_p = current_params
    else:
        ema_p = current_params if first else float(ema_alpha) * current_params + (1.0 - float(ema_alpha)) * carry.ema_params
    return carry._replace(ema_loss_delta=ema, ema_params=ema_p, steps_since_threshold_start=carry.steps_since_threshold_start + 1), raw


def check_and_advance_threshold(carry, current_loss, step, thresholds, min_steps, initial_steps):
    """reference opt/track.py:81-125; `step` is the 1-based state.step after the update (opt/optimiser.py:545-551)"""
    n = len(thresholds)
    thr = np.float32(thresholds[min(carry.current_threshold_idx, n - 1)])
    met = carry.steps_since_threshold_start >= min_steps and get_relative_convergence(carry, current_loss) < thr and step > initial_steps
    more = carry.current_threshold_idx + 1 < n
    if met and more:
        return carry._replace(current_threshold_idx=carry.current_threshold_idx + 1, steps_since_threshold_start=0)
    return carry._replace(converged=bool(carry.converged or met))


class ConvergenceTracker:
    def __init__(self, convergence, learning_rate: float, ema_alpha: float, min_steps_per_threshold: int):
        self.ema_alpha = ema_alpha
        self.min_steps_per_threshold = min_steps_per_threshold
        self.convergence_thresholds = create_convergence_thresholds(convergence, learning_rate)
        self.current_threshold_idx = 0
        self.current_threshold = self.convergence_thresholds[0]
        self.ema_loss_delta: Optional[float] = None
        self.ema_params: Optional[Any] = None
        self.steps_since_threshold_start = 0

    def update(self, previous_loss, current_loss, current_params):
        if previous_loss is not None:
            raw = abs(previous_loss - current_loss)
            if self.ema_loss_delta is None or self.ema_params is None:
                self.ema_loss_delta, self.ema_params = raw, current_params
            else:
                self.ema_loss_delta = self.ema_alpha * raw + (1 - self.ema_alpha) * self.ema_loss_delta
                self.ema_params = self.ema_alpha * current_params + (1 - self.ema_alpha) * self.ema_params
        else:
            raw = 0.0
        self.steps_since_threshold_start += 1
        return raw

    def get_relative_convergence(self, current_loss) -> float:
        if self.ema_loss_delta is not None and current_loss > 0:
            return float(self.ema_loss_delta / current_loss)
        return 0.0

    def is_threshold_met(self, current_loss, step: int, initial_steps: int) -> bool:
        return bool(
            self.steps_since_threshold_start >= self.min_steps_per_threshold
            and self.ema_loss_delta is not None
            and self.get_relative_convergence(current_loss) < self.current_threshold
            and step > initial_steps
        )

    def advance_threshold(self) -> bool:
        self.current_threshold_idx += 1
        self.steps_since_threshold_start = 0
        if self.current_threshold_idx >= len(self.convergence_thresholds):
            return False
        self.current_threshold = self.convergence_thresholds[self.current_threshold_idx]
        return True

    def reset_threshold_steps(self) -> None:
        self.steps_since_threshold_start = 0
