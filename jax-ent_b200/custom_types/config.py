"""Optimiser settings and the enumeration of optimisable parameter groups -- the two configuration types the optimise path
reads (names, fields, defaults and enum values as in the reference's `custom_types/config.py:28-51`; its featuriser and
experiment `Settings` classes belong to parts of the package that stay on stock JAX-ENT)."""
from __future__ import annotations

import collections
import enum
from dataclasses import dataclass, field
from typing import List, Union

# the four legacy loss constants travel with the settings but are not read by any loss on this path
LossConstants = collections.namedtuple("LossConstants", ("GAMMA", "LAMBDA", "PHI", "PSI"))

# values are part of the on-disk / CLI vocabulary of the reference: keep them stable
Optimisable_Parameters = enum.Enum(
    "Optimisable_Parameters", {"frame_weights": 0, "model_parameters": 1, "forward_model_weights": 2, "frame_mask": 3}, module=__name__
)


@dataclass
class OptimiserSettings:
    name: str
    n_steps: int = 100  # upper bound of the loop (`_optimise`, opt/run.py:272)
    tolerance: float = 1e-2  # stop when the total train loss falls below it
    convergence: Union[float, List[float]] = 1e-5  # threshold(s), multiplied by the learning rate (opt/track.py:12-21)
    learning_rate: float = 1e-4
    optimiser_type: str = "adam"
    loss_constants: LossConstants = field(default_factory=lambda: LossConstants(0.1, 0.1, 0.1, 0.1))
    ema_alpha: float = 0.5  # EMA factor of the loss delta and of the parameters
    min_steps_per_threshold: int = 2
