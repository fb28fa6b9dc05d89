"""run_optimise / _optimise (reference jaxent/src/opt/run.py:235-461): the Python optimisation loop with
relative-convergence thresholds, history snapshots and best-state selection, over the fused CUDA step.

Same semantics as the reference loop; the differences are mechanical: the per-step scalars (loss, grad-dot)
come from the device-side loss record in ONE small device->host read per step (the reference does 5-8
separate host syncs), and frame-sized arrays are only copied when a threshold snapshot is taken."""
from __future__ import annotations

import logging
import math
import time
from typing import Callable, Optional, Sequence

import torch

from ..custom_types.config import OptimiserSettings
from ..interfaces.simulation import Simulation_Parameters, _replace
from ..models.core import Simulation
from .base import OptimizationHistory, OptimizationState
from .optimiser import OptaxOptimizer
from .track import (ConvergenceTracker, check_and_advance_threshold, create_convergence_thresholds, initialise_convergence_carry,
                    update_convergence)

LOGGER = logging.getLogger("jaxent.opt")


def _optimise(_simulation: Simulation, data_to_fit: Sequence, n_steps: int, tolerance: float, convergence, indexes: Sequence[int],
              loss_functions: Sequence, opt_state: OptimizationState, optimizer: OptaxOptimizer, ema_alpha: float = 0.5,
              min_steps_per_threshold: int = 2, logger: Optional[logging.Logger] = None, silent: bool = False):
    _logger = logger if logger is not None else LOGGER
    if isinstance(convergence, float):
        convergence = [convergence]
    tracker = ConvergenceTracker(convergence, optimizer.learning_rate, ema_alpha, min_steps_per_threshold)
    previous_loss = None
    save_state = None
    optimizer.history = OptimizationHistory()
    t_start = time.time()
    step = 0
    data_targets, losses_t, idx_t = tuple(data_to_fit), tuple(loss_functions), tuple(indexes)
    try:
        for step in range(n_steps):
            opt_state, current_loss, save_state, _simulation = optimizer.step(
                optimizer=optimizer, state=opt_state, simulation=_simulation, data_targets=data_targets,
                loss_functions=losses_t, indexes=idx_t)
            eng = optimizer._engine
            # one device->host read: the record of the step just taken (loss) and of the next (grad-dot, lr)
            rec, nxt = eng.records(step, 2)
            loss_val = float(rec.total_train)
            optimizer._current_lr = float(nxt.lr)
            optimizer._current_model_lr = float(nxt.lr) * optimizer.model_parameters_lr_scale
            optimizer._current_gradient_mask_idx = 1 if step >= optimizer.initial_steps else 0
            tracker.update(previous_loss, loss_val, save_state.params)
            previous_loss = loss_val
            grad_dot = float(nxt.grad_dot) if step > 0 else 0.0  # check_gradient_oscillation: zeros before the first step
            if not silent:
                _logger.debug("step %d/%d loss %.6g rel-conv %.3e thr %.3e grad-dot %.3e lr %.4g", step, n_steps, loss_val,
                              tracker.get_relative_convergence(loss_val), tracker.current_threshold, grad_dot, nxt.lr)
            if grad_dot < 0:
                tracker.reset_threshold_steps()
            if loss_val < tolerance or math.isnan(loss_val) or math.isinf(loss_val):
                _logger.info("Stopping optimisation at step %s due to tolerance/non-finite loss.", step)
                break
            if step == 0 or tracker.is_threshold_met(loss_val, step, optimizer.initial_steps):
                optimizer = optimizer.update_history_compute_ema_loss(
                    optimizer=optimizer, simulation=_simulation, data_targets=data_targets, indexes=idx_t,
                    loss_functions=losses_t, state=save_state, ema_params=tracker.ema_params)
                if step > 0:
                    if tracker.advance_threshold():
                        _logger.info("Moving to threshold %s/%s: %.2e", tracker.current_threshold_idx + 1,
                                     len(tracker.convergence_thresholds), tracker.current_threshold)
                    else:
                        _logger.info("All relative thresholds completed at step %s", step)
                        break
    except Exception as e:  # reference opt/run.py:350-353
        raise RuntimeError(f"Optimisation failed at step {step}: {e}") from e
    if optimizer.history.states:
        best = optimizer.history.get_best_state()
        if best is not None:
            _simulation.params = optimizer.history.best_state.params
    if not silent:
        dt = time.time() - t_start
        _logger.info("Optimisation finished: %d steps in %.3f s (%.1f it/s)", step + 1, dt, (step + 1) / max(dt, 1e-9))
    return _simulation, optimizer


def _optimise_pure(_simulation: Simulation, data_to_fit: Sequence, n_steps: int, tolerance: float, convergence, indexes: Sequence[int],
                   loss_functions: Sequence, opt_state: OptimizationState, optimizer: OptaxOptimizer, ema_alpha: float = 0.5,
                   min_steps_per_threshold: int = 2, learning_rate: Optional[float] = None, **_unused):
    """The loop of the reference's `_optimise_pure` (opt/run.py:141-232: lax.while_loop over OptaxOptimizer._pure_step,
    opt/optimiser.py:494-579) driven from the host over the fused CUDA step -- the semantics batch_optimise vmaps
    (opt/batch.py:107-146), which differ from the Python loop `_optimise`:
      * the state enters dense: previous_loss of the first iteration is the loss at the initial parameters (first delta 0);
      * update_convergence / check_and_advance_threshold (opt/track.py:40-125): EMAs restart with every threshold, no
        gradient-oscillation reset, the threshold test uses the 1-based step;
      * cond_fn: stop when converged, after n_steps, or when the loss stored by the previous iteration is non-finite / below
        the tolerance (checked on the initial loss before the first iteration);
      * every iteration is a history entry: the parameters AFTER the update with the losses evaluated BEFORE it, step = the
        history index; best state = smallest val_losses[0], the last entry on ties with it (opt/base.py:180-189).
    Frame weights are kept for the best and the last entry (the others carry frame_weights=None): the reference's dense
    [n_steps, n_frames] history does not fit at the ensemble sizes this path is built for.  Returns (simulation, optimizer);
    optimizer.history holds the entries, optimizer._pure_carry the final ConvergenceCarry and write_idx."""
    from .optimiser import compute_loss

    if isinstance(convergence, float):
        convergence = [convergence]
    lr = optimizer.learning_rate if learning_rate is None else float(learning_rate)
    thresholds = create_convergence_thresholds(convergence, lr)
    data_targets, losses_t, idx_t = tuple(data_to_fit), tuple(loss_functions), tuple(indexes)
    optimizer.history = OptimizationHistory()
    carry = initialise_convergence_carry(None)  # EMA parameters are not part of the returned history (opt/batch.py:158-180)
    initial_losses, _ = compute_loss(_simulation, opt_state.params, data_targets, idx_t, losses_t)  # _ensure_dense_state, run.py:89-118
    prev_loss = float(initial_losses.total_train_loss)
    best_i, best_val = None, None
    write_idx = 0
    states = optimizer.history.states

    def _strip(i):  # drop the frame weights of entry i (idempotent)
        states[i] = states[i]._replace(params=_replace(states[i].params, frame_weights=None))

    while not (carry.converged or write_idx >= n_steps or not math.isfinite(prev_loss) or prev_loss < tolerance):
        opt_state, _, save_state, _simulation = optimizer.step(
            optimizer=optimizer, state=opt_state, simulation=_simulation, data_targets=data_targets, loss_functions=losses_t, indexes=idx_t)
        rec = optimizer._engine.records(write_idx, 1)[0]  # one device->host read per iteration
        loss_val, val0 = float(rec.total_train), float(rec.val[0])
        carry, _ = update_convergence(carry, prev_loss, loss_val, None, ema_alpha)
        carry = check_and_advance_threshold(carry, loss_val, write_idx + 1, thresholds, min_steps_per_threshold, optimizer.initial_steps)
        prev_loss = loss_val
        entry = OptimizationState(save_state.params, save_state.opt_state, write_idx, save_state.losses, None)
        if best_val is None or val0 < best_val:  # first strict minimum so far: this entry keeps its weights, the old best drops them
            if best_i is not None:
                _strip(best_i)
            best_i, best_val = write_idx, val0
        if states and (write_idx - 1) != best_i:  # the previous entry is no longer the last one
            _strip(write_idx - 1)
        states.append(entry)
        write_idx += 1
    if states:
        last_val0 = float(states[-1].losses.val_losses[0])
        pick = write_idx - 1 if not (best_val < last_val0) else best_i
        optimizer.history.best_state = states[pick]
        _simulation.params = optimizer.history.best_state.params
    optimizer._pure_carry = (carry, write_idx)
    return _simulation, optimizer


def run_optimise(simulation: Simulation, data_to_fit: Sequence, config: OptimiserSettings, forward_models: Sequence,
                 indexes: Sequence[int], loss_functions: list, optimizer: Optional[OptaxOptimizer] = None,
                 optimisable_funcs=None, initialise: Optional[bool] = False, _opt_fn: Optional[Callable] = None,
                 jit_update_step: bool = False, logger: Optional[logging.Logger] = None, silent: bool = False):
    """Run the optimisation (reference opt/run.py:376-461) -> (simulation, OptimizationHistory).

    `_opt_fn` (reference default: `_optimise`, the Python loop): left at None, the loop runs on the device
    (`_optimise_device`: tracker, EMA parameters, snapshots and the stop decision in the tail kernel, one status read per 32
    steps) -- it returns the same history, EMA history and best state as the Python loop, bit for bit
    (tests/test_gpu_api.py::test_on_device_loop_matches_host_loop), without its per-step host synchronisation.  The
    Python loop is used when the logger is at DEBUG level (it logs every step) or when `_opt_fn=_optimise` is passed."""
    _logger = logger if logger is not None else LOGGER
    if _opt_fn is None:
        _opt_fn = _optimise if (not silent and _logger.isEnabledFor(logging.DEBUG)) else _optimise_device
    if initialise:
        if not simulation.initialise():
            raise ValueError("Failed to initialise simulation")
    if len(data_to_fit) != len(loss_functions):
        raise ValueError("Number of data targets and loss functions must match")
    if optimizer is None:
        optimizer = OptaxOptimizer(learning_rate=config.learning_rate, optimizer=config.optimiser_type)
    jit_test_args = (data_to_fit, loss_functions, indexes) if jit_update_step else None
    opt_state = optimizer.initialise(simulation, optimisable_funcs, _jit_test_args=jit_test_args)
    kw = dict(ema_alpha=config.ema_alpha, min_steps_per_threshold=config.min_steps_per_threshold)
    if _opt_fn is _optimise:
        kw.update(logger=_logger, silent=silent)
    simulation, optimizer = _opt_fn(simulation, data_to_fit, config.n_steps, config.tolerance, config.convergence, indexes,
                                    loss_functions, opt_state, optimizer, **kw)
    if optimizer:
        return simulation, optimizer.history
    raise ValueError("Optimisation failed")


def _optimise_device(_simulation: Simulation, data_to_fit: Sequence, n_steps: int, tolerance: float, convergence, indexes: Sequence[int],
                     loss_functions: Sequence, opt_state: OptimizationState, optimizer: OptaxOptimizer, ema_alpha: float = 0.5,
                     min_steps_per_threshold: int = 2, chunk: int = 32, **_unused):
    """Drop-in replacement for `_optimise` (pass it as `run_optimise(..., _opt_fn=_optimise_device)`) that keeps the
    whole loop on the device: ConvergenceTracker, EMA parameters, history snapshots and the stop decision are evaluated
    by the tail kernel every step; the host reads one 32-byte status per `chunk` steps instead of synchronising 5-8
    times per step.  Produces the same OptimizationHistory / ema_history / best state."""
    from .optimiser import B200OptState, _build_engine, _losses_from_record, compute_loss
    from .._arrays import scalar, to_device, to_numpy
    from .. import _lib as L
    import ctypes as C
    import numpy as np

    if isinstance(convergence, float):
        convergence = [convergence]
    data_targets, losses_t, idx_t = tuple(data_to_fit), tuple(loss_functions), tuple(indexes)
    params = opt_state.params
    eng = _build_engine(_simulation, data_targets, losses_t, idx_t, optimizer._settings())
    eng.enable_tracking(convergence, optimizer.learning_rate, ema_alpha, min_steps_per_threshold, tolerance)
    mp = params.model_parameters[0]
    eng.init_state(to_device(params.frame_weights, _simulation.device).reshape(-1), scalar(mp.bv_bc), scalar(mp.bv_bh),
                   to_numpy(params.forward_model_weights), to_numpy(params.forward_model_scaling))
    optimizer._engine, optimizer._engine_key, optimizer._n_losses = eng, None, len(losses_t)
    optimizer.history = OptimizationHistory()
    if optimizer.save_ema_history:
        optimizer.ema_history = OptimizationHistory()
    t0 = time.time()
    st = eng.run(n_steps, chunk=chunk)
    n = len(losses_t)
    for sn, w, ema_w in eng.track_snapshots(st.n_snapshots):
        mp_s = type(mp)(bv_bc=torch.tensor([sn.bc], device=eng.device), bv_bh=torch.tensor([sn.bh], device=eng.device),
                        temperature=mp.temperature, timepoints=mp.timepoints)
        sp = _replace(Simulation_Parameters.normalize_weights(params), frame_weights=w.clone(), model_parameters=[mp_s])
        rec = torch.frombuffer(bytearray(bytes(sn.losses)), dtype=torch.float32).to(eng.device)
        state = OptimizationState(sp, B200OptState(eng, sn.state_step), sn.state_step, _losses_from_record(rec, n), None)
        optimizer.history.add_state(state)
        if optimizer.save_ema_history and optimizer.ema_history is not None and ema_w is not None:
            mp_e = type(mp)(bv_bc=torch.tensor([sn.ema_bc], device=eng.device), bv_bh=torch.tensor([sn.ema_bh], device=eng.device),
                            temperature=mp.temperature, timepoints=mp.timepoints)
            ep = _replace(sp, frame_weights=ema_w.clone(), model_parameters=[mp_e])
            el, _ = compute_loss(_simulation, ep, data_targets, idx_t, losses_t)
            optimizer.ema_history.add_state(OptimizationState(ep, state.opt_state, sn.state_step, el, None))
    if optimizer.history.states:
        best = optimizer.history.get_best_state()
        if best is not None:
            _simulation.params = best.params
    optimizer._device_status = st
    LOGGER.info("On-device optimisation: %d updates, %d snapshots, stop=%d (reason %d) in %.3f s", st.steps_done, st.n_snapshots,
                st.stop, st.reason, time.time() - t0)
    return _simulation, optimizer
