"""ConvergenceTracker (reference jaxent/src/opt/track.py:12-21,128-197): LR-scaled descending thresholds,
EMA of |delta loss| and of the normalised parameters."""
from __future__ import annotations

from typing import Any, Optional

import numpy as np


def create_convergence_thresholds(convergence, learning_rate) -> list:
    thr = np.atleast_1d(np.asarray(convergence, np.float32))
    return [float(t) for t in (np.sort(thr)[::-1] * np.float32(learning_rate))]


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
