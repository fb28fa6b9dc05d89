"""LossComponents / OptimizationState / OptimizationHistory (reference jaxent/src/opt/base.py:61-196)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, NamedTuple, Optional, Protocol, runtime_checkable

from ..interfaces.simulation import Simulation_Parameters


@runtime_checkable
class JaxEnt_Loss(Protocol):
    def __call__(self, model, dataset, prediction_index) -> tuple: ...


class LossComponents(NamedTuple):
    train_losses: Any
    val_losses: Any
    scaled_train_losses: Any
    scaled_val_losses: Any
    total_train_loss: Any
    total_val_loss: Any


class ConvergenceCarry(NamedTuple):
    """carry of the functional tracker (reference opt/base.py: ConvergenceCarry), scalars held as Python / NumPy fp32 values"""

    ema_loss_delta: Any
    ema_params: Any
    steps_since_threshold_start: Any
    current_threshold_idx: Any
    converged: Any


class OptimizationState(NamedTuple):
    params: Simulation_Parameters
    opt_state: Any
    step: Any = 0
    losses: Optional[LossComponents] = None
    gradients: Optional[Simulation_Parameters] = None

    def update(self, new_params, new_opt_state, new_losses, new_gradients=None, step=None):
        return OptimizationState(
            params=new_params,
            opt_state=new_opt_state,
            step=(self.step + 1) if step is None else step,
            losses=new_losses if new_losses is not None else self.losses,
            gradients=new_gradients if new_gradients is not None else self.gradients,
        )


@dataclass
class OptimizationHistory:
    states: list = field(default_factory=list)
    best_state: Optional[OptimizationState] = None

    def add_state(self, state: OptimizationState):
        self.states.append(state)

    @staticmethod
    def _pick_best_state(states):
        best = states[-1]
        for st in states:
            if st.losses.val_losses[0] < best.losses.val_losses[0]:  # first validation component (:180-189)
                best = st
        return best

    def get_best_state(self):
        self.best_state = self._pick_best_state(self.states)
        return self.best_state


class HParamBatch(NamedTuple):
    """Batch of hyper-parameter values for `batch_optimise` (reference opt/base.py:212-217)."""

    forward_model_weights: Any  # (n_hparams, n_losses)
    forward_model_scaling: Any  # (n_hparams, n_losses)
    learning_rate: Optional[Any] = None  # (n_hparams,)


class BatchOptimisationResult(NamedTuple):
    """Structured output of `batch_optimise` (reference opt/base.py:220-227)."""

    histories: tuple
    best_states: tuple
    convergence_steps: Any
    hparam_batch: HParamBatch
