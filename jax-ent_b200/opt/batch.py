"""batch_optimise (reference jaxent/src/opt/batch.py:51-192): sweeps over (forward_model_weights,
forward_model_scaling, learning_rate) that share the features, the data and the starting point.

The reference vmaps `_optimise_pure` over the hyper-parameter batch.  Here a batch of replicas is advanced by ONE fused
step on the tensor cores (BatchedBvEngine: the two feature sweeps are tcgen05 GEMMs over all replicas, DESIGN.md
section 6) with the per-replica tracker of `_optimise_pure` (update_convergence / check_and_advance_threshold,
opt/track.py:40-125), the best-state selection and the stop decision evaluated on the device; replicas whose loop has
ended are frozen while the others continue -- the semantics of a vmapped `lax.while_loop`.  Every step is a history entry
with its losses, as in the reference; frame weights are kept for the best and the last entry only.  The
tensor-core path needs features that are exact in bf16 (hard-cutoff contact counts); otherwise (`engine="sequential"`,
or "auto" on other features) every run goes through `_optimise_pure` (opt/run.py: the same loop semantics, driven from the
host over the single-ensemble fused step) on shared packed features, one replica after the other."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from .._arrays import to_numpy
from ..custom_types.config import OptimiserSettings
from ..interfaces.simulation import Simulation_Parameters, _replace
from ..models.core import Simulation
from .base import BatchOptimisationResult, HParamBatch, OptimizationHistory, OptimizationState  # noqa: F401 (re-exported)
from .optimiser import OptaxOptimizer
from .run import _optimise_pure


def _batch_optimise_gemm(simulation: Simulation, fw, fs, lrs, batch_size: int, data_to_fit, config: OptimiserSettings, indexes,
                         loss_functions, optimizer: OptaxOptimizer, optimisable_funcs):
    """the tensor-core path: batches of up to 256 replicas on one BatchedBvEngine"""
    from .. import _lib as L
    from .._arrays import scalar, to_device
    from ..batch_engine import MAX_REPLICAS, BatchedBvEngine, pack_gemm_features
    from ..interfaces.simulation import Simulation_Parameters
    from .optimiser import B200OptState, _losses_from_record, _slot_specs

    feats = simulation.__dict__.get("_gemm_features")
    if feats is None:
        f0 = simulation.input_features[0]
        feats = pack_gemm_features(f0.heavy_contacts, f0.acceptor_contacts, simulation.device)  # raises ValueError if not bf16-exact
        simulation.__dict__["_gemm_features"] = feats
    data_targets, losses_t, idx_t = tuple(data_to_fit), tuple(loss_functions), tuple(indexes)
    specs, prior = _slot_specs(simulation, data_targets, losses_t, idx_t)
    state0 = optimizer.initialise(simulation, optimisable_funcs)
    params = state0.params
    mp = params.model_parameters[0]
    z0 = to_device(params.frame_weights, simulation.device).reshape(-1)
    n_h, n_l = fw.shape[0], len(losses_t)
    convergence = [config.convergence] if isinstance(config.convergence, float) else list(config.convergence)
    histories, best_states, conv_steps = [], [], []
    per = max(1, min(int(batch_size), MAX_REPLICAS))
    for lo in range(0, n_h, per):
        hi = min(lo + per, n_h)
        eng = BatchedBvEngine(feats, None, simulation._k_ints, simulation._timepoints, specs, simulation._uptake, optimizer._settings(),
                              hi - lo, prior=prior, device=simulation.device)
        eng.enable_tracking(convergence, config.ema_alpha, config.min_steps_per_threshold, config.tolerance, loop="pure")
        lr = lrs[lo:hi] if lrs is not None else np.full(hi - lo, config.learning_rate, np.float32)
        eng.init_state(z0, scalar(mp.bv_bc), scalar(mp.bv_bh), fw[lo:hi], fs[lo:hi], lr)
        status = eng.run(config.n_steps, keep_records=True)
        w_last = eng.weights_all()
        norm = Simulation_Parameters.normalize_weights(params)

        # the per-step loss records of all replicas, uploaded once: history entries are views into this tensor
        recs = torch.from_numpy(eng.record_history.copy()).view(torch.float32).to(eng.device)  # (steps + 1, Bn, record floats)
        val0_all = eng.record_history.view(np.float32)[:, :, 14]  # LossRecord.val[0]

        def _state(b, i, weights):
            # history entry i of the reference (opt/optimiser.py:555-567, opt/batch.py:158-172): the parameters AFTER update i
            # (the BV parameters step i + 1 was evaluated at) paired with the losses evaluated BEFORE it; step = the history index
            nxt = recs[i + 1, b]
            mp_s = type(mp)(bv_bc=nxt[4:5], bv_bh=nxt[5:6], temperature=mp.temperature, timepoints=mp.timepoints)
            sp = _replace(norm, frame_weights=weights, model_parameters=[mp_s],
                          forward_model_weights=fw[lo + b].astype(np.float32), forward_model_scaling=fs[lo + b].astype(np.float32))
            return OptimizationState(sp, B200OptState(None, i + 1), i, _losses_from_record(recs[i, b], n_l, copy=False), None)

        for b in range(hi - lo):
            st = status[b]
            n_written = int(st.steps_done)  # write_idx of the reference carry
            hist = OptimizationHistory()
            if n_written:
                sn, w_best, _ = eng.track_snapshots(b, 1)[0]
                val0 = val0_all[:n_written, b]
                # _pick_best_state (opt/base.py:180-189): the last entry unless an earlier one is strictly better, then the first
                # entry that reaches the minimum -- which is the one the device followed
                best_i = n_written - 1 if not (val0.min() < val0[-1]) else int(sn.loop_step)
                for i in range(n_written):
                    if i == n_written - 1:
                        weights = w_last[:, b].clone()
                    elif i == best_i:
                        weights = w_best.clone()
                    else:
                        weights = None  # a dense [n_steps, F] history per replica is not kept on this path
                    hist.states.append(_state(b, i, weights))
                hist.best_state = hist.states[best_i]
            histories.append(hist)
            best_states.append(hist.best_state if hist.states else state0)
            conv_steps.append(n_written)
        del eng
    return histories, best_states, conv_steps


def batch_optimise(simulation: Simulation, hparam_batch: HParamBatch, batch_size: int, data_to_fit: Sequence,
                   config: OptimiserSettings, indexes: Sequence[int], loss_functions: Sequence,
                   optimizer: Optional[OptaxOptimizer] = None, optimisable_funcs=None, engine: str = "auto") -> BatchOptimisationResult:
    if engine not in ("auto", "gemm", "sequential"):
        raise ValueError("engine must be 'auto', 'gemm' or 'sequential'")
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

    result_hparams = HParamBatch(hparam_batch.forward_model_weights, hparam_batch.forward_model_scaling, hparam_batch.learning_rate)
    if not getattr(simulation, "_average_first", True):
        if engine == "gemm":
            raise NotImplementedError("the tensor-core sweep covers average_first=True forward passes")
        engine = "sequential"
    if engine in ("auto", "gemm"):
        try:
            histories, best_states, conv_steps = _batch_optimise_gemm(simulation, fw, fs, lrs, batch_size, data_to_fit, config, indexes,
                                                                      loss_functions, optimizer, optimisable_funcs)
            return BatchOptimisationResult(tuple(histories), tuple(best_states), np.asarray(conv_steps, np.int32), result_hparams)
        except ValueError as e:
            if engine == "gemm" or "bf16" not in str(e):
                raise

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
        # the same loop semantics as the tensor-core path: _optimise_pure (the loop the reference vmaps), host-driven
        _, run_opt = _optimise_pure(run_sim, data_to_fit, config.n_steps, config.tolerance, config.convergence, indexes,
                                    loss_functions, state, run_opt, ema_alpha=config.ema_alpha,
                                    min_steps_per_threshold=config.min_steps_per_threshold)
        hist = run_opt.history
        histories.append(hist)
        best_states.append(hist.best_state if hist.states else state)
        conv_steps.append(int(run_opt._pure_carry[1]))
    return BatchOptimisationResult(tuple(histories), tuple(best_states), np.asarray(conv_steps, np.int32),
                                   HParamBatch(hparam_batch.forward_model_weights, hparam_batch.forward_model_scaling, hparam_batch.learning_rate))
