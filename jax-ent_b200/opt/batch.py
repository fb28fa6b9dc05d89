"""batch_optimise (reference jaxent/src/opt/batch.py:51-192): sweeps over (forward_model_weights,
forward_model_scaling, learning_rate) that share the features, the data and the starting point.

The reference vmaps `_optimise_pure` over the hyper-parameter batch.  Here every run uses the on-device loop
(`_optimise_device`) on an engine that shares ONE packed copy of the features, so the sweep costs no extra
feature memory and no host synchronisation inside a run.  Runs are issued back to back on the device; the
tensor-core formulation that advances all replicas in one (R x F)(F x B) GEMM per step (BASELINE config 5) is
the planned replacement for this loop and is described in DESIGN.md section 8."""
from __future__ import annotations

from typing import NamedTuple, Optional, Sequence

import numpy as np
import torch

from .._arrays import to_numpy
from ..custom_types.config import OptimiserSettings
from ..interfaces.simulation import Simulation_Parameters, _replace
from ..models.core import Simulation
from .base import OptimizationHistory, OptimizationState
from .optimiser import OptaxOptimizer
from .run import _optimise_device


class HParamBatch(NamedTuple):
    forward_model_weights: object  # (n_hparams, n_losses)
    forward_model_scaling: object  # (n_hparams, n_losses)
    learning_rate: Optional[object] = None  # (n_hparams,)


class BatchOptimisationResult(NamedTuple):
    histories: tuple
    best_states: tuple
    convergence_steps: object
    hparam_batch: HParamBatch


def batch_optimise(simulation: Simulation, hparam_batch: HParamBatch, batch_size: int, data_to_fit: Sequence,
                   config: OptimiserSettings, indexes: Sequence[int], loss_functions: Sequence,
                   optimizer: Optional[OptaxOptimizer] = None, optimisable_funcs=None) -> BatchOptimisationResult:
    if batch_size <= 0:
        raise ValueError("batch_size must be > 0")
    fw = to_numpy(hparam_batch.forward_model_weights)
    fs = to_numpy(hparam_batch.forward_model_scaling)
    if fw.shape[0] != fs.shape[0]:
        raise ValueError("forward_model_weights and forward_model_scaling must have the same n_hparams")
    lrs = None if hparam_batch.learning_rate is None else to_numpy(hparam_batch.learning_rate).reshape(-1)
    if lrs is not None and lrs.shape[0] != fw.shape[0]:
        raise ValueError("learning_rate must match n_hparams when provided")
    if simulation._features is None:
        if not simulation.initialise():
            raise ValueError("Failed to initialise simulation before batch optimisation")
    if optimizer is None:
        optimizer = OptaxOptimizer(learning_rate=config.learning_rate, optimizer=config.optimiser_type)

    histories, best_states, conv_steps = [], [], []
    base_params = simulation.params
    for i in range(fw.shape[0]):
        run_opt = OptaxOptimizer(
            learning_rate=float(lrs[i]) if lrs is not None else config.learning_rate,
            optimizer="adam",
            parameter_partition_masks=set(optimizer.parameter_partition_masks),
            clip_value=optimizer.clip_value,
            plateau_denominator=optimizer.plateau_denominator,
            save_ema_history=False,
            initial_learning_rate=optimizer.initial_learning_rate,
            initial_steps=optimizer.initial_steps,
            model_parameters_lr_scale=optimizer.model_parameters_lr_scale,
        )
        # hyper-parameters replace the simulation's weights / scaling as given (reference _replace_hparams, batch.py:36-48)
        run_sim = simulation._shallow_copy()
        run_sim.params = _replace(base_params, forward_model_weights=fw[i].astype(np.float32), forward_model_scaling=fs[i].astype(np.float32))
        state = run_opt.initialise(run_sim, optimisable_funcs)
        _, run_opt = _optimise_device(run_sim, data_to_fit, config.n_steps, config.tolerance, config.convergence, indexes,
                                      loss_functions, state, run_opt, ema_alpha=config.ema_alpha,
                                      min_steps_per_threshold=config.min_steps_per_threshold)
        hist = run_opt.history
        histories.append(hist)
        best_states.append(hist.best_state if hist.states else state)
        conv_steps.append(int(run_opt._device_status.steps_done))
    return BatchOptimisationResult(tuple(histories), tuple(best_states), np.asarray(conv_steps, np.int32),
                                   HParamBatch(hparam_batch.forward_model_weights, hparam_batch.forward_model_scaling, hparam_batch.learning_rate))
