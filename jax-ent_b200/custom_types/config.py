"""OptimiserSettings / Optimisable_Parameters (reference jaxent/src/custom_types/config.py:35-51)."""
from dataclasses import dataclass
from enum import Enum
from typing import NamedTuple


class LossConstants(NamedTuple):
    GAMMA: float
    LAMBDA: float
    PHI: float
    PSI: float


@dataclass
class OptimiserSettings:
    name: str
    n_steps: int = 100
    tolerance: float = 1e-2
    convergence: "float | list[float]" = 1e-5
    learning_rate: float = 1e-4
    optimiser_type: str = "adam"
    loss_constants: LossConstants = LossConstants(GAMMA=0.1, LAMBDA=0.1, PHI=0.1, PSI=0.1)
    ema_alpha: float = 0.5
    min_steps_per_threshold: int = 2


class Optimisable_Parameters(Enum):
    frame_weights = 0
    model_parameters = 1
    forward_model_weights = 2
    frame_mask = 3
