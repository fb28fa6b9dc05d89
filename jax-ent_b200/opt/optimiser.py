"""OptaxOptimizer and compute_loss (reference jaxent/src/opt/optimiser.py:46-644) over the fused CUDA step.

`OptaxOptimizer.step(optimizer=, state=, simulation=, data_targets=, loss_functions=, indexes=)` returns
`(new_state, loss_value, save_state, updated_sim)` like the reference (opt/run.py:276-283), with
`OptimizationState` / `LossComponents` / `Simulation_Parameters` results holding CUDA tensors.  One call
= one pass + reduce + tail launch on the device; nothing is synchronised unless the caller reads a value.

Only `optimizer="adam"` is on the fused path (the reference's default); the optimiser state (logits,
Adam moments, plateau-LR bookkeeping) lives on the device inside the engine, `state.opt_state` is a
handle to it."""
from __future__ import annotations

import logging
from typing import Any, Optional, Sequence

import numpy as np
import torch

from .. import _lib as L
from .._arrays import scalar, to_device, to_numpy
from ..custom_types.config import Optimisable_Parameters
from ..data.loader import ExpD_Dataloader
from ..engine import BvEngine, OptSettings, SlotSpec
from ..interfaces.simulation import Simulation_Parameters, _replace
from ..models.core import Simulation
from ..models.HDX.BV.features import BV_output_features, uptake_BV_output_features
from .base import LossComponents, OptimizationHistory, OptimizationState
from .losses import loss_kind

LOGGER = logging.getLogger("jaxent.opt")
_REC_FLOATS = 36  # sizeof(JaxentLossRecord) / 4


class B200OptState:
    """`OptimizationState.opt_state`: a handle to the device-resident Adam state of one engine."""

    def __init__(self, engine: Optional[BvEngine], step: int):
        self.engine, self.step = engine, step

    def __repr__(self):
        return f"B200OptState(step={self.step})"


def _slot_specs(simulation: Simulation, data_targets, loss_functions, indexes):
    """(loss function, data target) pairs -> engine slots (+ the MaxEnt prior)."""
    if len(data_targets) != len(loss_functions):
        raise ValueError("Number of data targets and loss functions must match")
    specs, prior = [], None
    for fn, tgt, idx in zip(loss_functions, data_targets, indexes):
        kind = loss_kind(fn)
        if idx not in (0, None):
            raise NotImplementedError("prediction_index must be 0: the fused path has one BV forward model")
        if kind <= L.LOSS_PF_L2:
            if not hasattr(tgt, "train") or not hasattr(tgt, "val"):
                raise TypeError(f"{fn.__name__} needs an ExpD_Dataloader with train / val datasets")
            tr, va = tgt.train, tgt.val
            dense = lambda ds: ds.data_mapping.dense() if hasattr(ds.data_mapping, "dense") else to_numpy(ds.residue_feature_ouput_mapping)
            y = lambda ds: to_numpy(ds.y_true).reshape(np.shape(ds.y_true)[0], -1) if kind != L.LOSS_PF_L2 else to_numpy(ds.y_true).reshape(-1)
            cov = lambda ds: None if ds.covariance_matrix is None else to_numpy(ds.covariance_matrix)
            specs.append(SlotSpec(kind, dense(tr), y(tr), dense(va), y(va), cov(tr), cov(va)))
        elif kind == L.LOSS_MAXENT_CONVEX_KL:
            if not isinstance(tgt, Simulation_Parameters):
                raise TypeError("maxent_convexKL_loss needs the prior Simulation_Parameters as its data target")
            prior = tgt.frame_weights
            specs.append(SlotSpec(kind))
        else:
            if not isinstance(tgt, Simulation_Parameters):
                raise TypeError(f"{fn.__name__} needs the prior Simulation_Parameters as its data target")
            mp = tgt.model_parameters[0]
            specs.append(SlotSpec(kind, prior_bc=scalar(mp.bv_bc), prior_bh=scalar(mp.bv_bh)))
    return specs, prior


def _build_engine(simulation: Simulation, data_targets, loss_functions, indexes, settings: OptSettings) -> BvEngine:
    if simulation._features is None:
        raise RuntimeError("Simulation must be initialised first")
    specs, prior = _slot_specs(simulation, data_targets, loss_functions, indexes)
    return BvEngine(simulation._features, None, simulation._k_ints, simulation._timepoints, specs, simulation._uptake,
                    settings, prior=prior, device=simulation.device, average_first=getattr(simulation, "_average_first", True))


_EVAL_CACHE_MAX = 4  # scratch engines kept per Simulation (each pins one per-frame workspace on the GPU)


def _target_fingerprint(t):
    """what the engine copied to the device at build time, as far as it can be told cheaply (this runs on EVERY optimiser
    step, so nothing frame-sized is touched): shapes, storage addresses and the small target arrays' sums.  A new object at a
    recycled id() or a re-assigned array does not hit a stale engine; in-place edits of a frame-sized prior are not detected."""
    def addr(x):
        if isinstance(x, torch.Tensor):
            return x.data_ptr()
        if isinstance(x, np.ndarray):
            return x.__array_interface__["data"][0]
        return id(x)

    parts = []
    for ds in (getattr(t, "train", None), getattr(t, "val", None)):
        if ds is not None:
            y = getattr(ds, "y_true", None)
            if y is None:
                parts.append(None)
            else:
                small = int(np.prod(np.shape(y))) <= 65536
                parts.append((tuple(np.shape(y)), addr(y), float(np.asarray(to_numpy(y), np.float64).sum()) if small else 0.0))
    fwts = getattr(t, "frame_weights", None)
    if fwts is not None:
        parts.append((tuple(np.shape(fwts)), addr(fwts)))
    mps = getattr(t, "model_parameters", None)
    if mps:
        parts.append((scalar(mps[0].bv_bc), scalar(mps[0].bv_bh)))
    return tuple(parts)


def _cached_engine(cache: list, simulation, data_targets, loss_functions, indexes, settings, max_entries: int) -> BvEngine:
    """Engine for (targets, losses, indexes): entries hold STRONG references to their targets and are matched by identity
    (an id() can be recycled once a dataloader is freed, e.g. across cross-validation folds on one Simulation) plus a
    content fingerprint; least recently used entries beyond `max_entries` are closed."""
    names = tuple(getattr(f, "__name__", repr(f)) for f in loss_functions)
    idx = tuple(indexes)
    fps = tuple(_target_fingerprint(t) for t in data_targets)
    for i, (tg, nm, ix, fp, eng) in enumerate(cache):
        if nm == names and ix == idx and len(tg) == len(data_targets) and all(a is b for a, b in zip(tg, data_targets)) and fp == fps:
            cache.append(cache.pop(i))
            return eng
    eng = _build_engine(simulation, data_targets, loss_functions, indexes, settings)
    cache.append((tuple(data_targets), names, idx, fps, eng))
    while len(cache) > max_entries:
        old = cache.pop(0)
        old[4].close(barrier=False)
    return eng


def _record_view(eng: BvEngine, step: int) -> torch.Tensor:
    """float32 view (on the device, no sync) of the loss record of `step`."""
    recs = eng._records.view(torch.float32)
    i = (step % L.RECORD_RING) * _REC_FLOATS
    return recs[i : i + _REC_FLOATS]


def _losses_from_record(rec: torch.Tensor, n: int, copy: bool = True) -> LossComponents:
    if copy:
        rec = rec.clone()  # the ring slot is reused after JAXENT_RECORD_RING steps
    return LossComponents(
        train_losses=rec[8 : 8 + n],
        val_losses=rec[14 : 14 + n],
        scaled_train_losses=rec[20 : 20 + n],
        scaled_val_losses=rec[26 : 26 + n],
        total_train_loss=rec[6],
        total_val_loss=rec[7],
    )


def _outputs_of(eng: BvEngine):
    if eng.uptake:
        return (uptake_BV_output_features(uptake=eng.uptake_out.clone()),)
    return (BV_output_features(log_Pf=eng.log_pf.clone(), k_ints=None),)


def compute_loss(simulation: Simulation, params: Simulation_Parameters, data_targets, indexes, loss_functions):
    """Forward + all losses at `params` (reference opt/optimiser.py:608-644) -> (LossComponents, updated sim).
    Runs the forward-only pass and the tail kernel on a scratch engine that shares the packed features."""
    eng = _cached_engine(simulation.__dict__.setdefault("_eval_engines", []), simulation, data_targets, loss_functions, indexes,
                         OptSettings(), _EVAL_CACHE_MAX)
    mp = params.model_parameters[0]
    eng.init_state(to_device(params.frame_weights, simulation.device).reshape(-1), scalar(mp.bv_bc), scalar(mp.bv_bh),
                   to_numpy(params.forward_model_weights), to_numpy(params.forward_model_scaling))
    eng.finish_record()
    losses = _losses_from_record(_record_view(eng, 0), len(loss_functions))
    new_sim = simulation._shallow_copy()
    new_sim.params = _replace(Simulation_Parameters.normalize_weights(params), frame_weights=eng.weights.clone())
    new_sim.outputs = _outputs_of(eng)
    return losses, new_sim


class OptaxOptimizer:
    def __init__(
        self,
        learning_rate: float = 1e-4,
        optimizer: str = "adam",
        parameter_partition_masks: set = frozenset({Optimisable_Parameters.frame_weights}),
        clip_value: Optional[float] = 1.0,
        force_simplex: Optional[bool] = None,
        plateau_denominator: float = 1.005,
        save_ema_history: bool = True,
        initial_learning_rate: float = 1e0,
        initial_steps: int = 0,
        model_parameters_lr_scale: float = 1.0,
    ):
        if optimizer.lower() not in ("adam", "sgd", "adagrad", "adamw", "rmsprop", "lbfgs"):
            raise ValueError(f"Unsupported optimizer: {optimizer}")
        if optimizer.lower() != "adam":
            raise NotImplementedError(f"optimizer={optimizer!r}: only 'adam' is on the fused B200 path")
        if force_simplex:
            raise NotImplementedError("force_simplex is only used by the reference's sgd path")
        if Optimisable_Parameters.frame_mask in parameter_partition_masks:
            raise NotImplementedError(  # reference opt/gradients.py:54-58
                "Frame mask optimization not fully implemented - while gradients can flow, "
                "frame masking is not applied during weights normalisation before the forward step"
            )
        self.parameter_partition_masks = set(parameter_partition_masks)
        self.clip_value = clip_value
        self.history = OptimizationHistory()
        self.save_ema_history = save_ema_history
        self.ema_history = OptimizationHistory() if save_ema_history else None
        self.model_parameters_lr_scale = model_parameters_lr_scale
        self.plateau_denominator = plateau_denominator
        self.initial_learning_rate = initial_learning_rate
        self.initial_steps = initial_steps
        self.learning_rate = learning_rate
        self.force_logit_simplex = False
        self._current_lr = float(initial_learning_rate)
        self._current_model_lr = float(initial_learning_rate * model_parameters_lr_scale)
        self._current_gradient_mask_idx = 0
        self._engine: Optional[BvEngine] = None
        self._engine_key = None
        self._init_params: Optional[Simulation_Parameters] = None
        self.step = self._step

    @property
    def current_learning_rate(self) -> float:
        return self._current_lr

    @property
    def current_model_learning_rate(self) -> float:
        return self._current_model_lr

    def _settings(self) -> OptSettings:
        return OptSettings(
            learning_rate=self.learning_rate,
            initial_learning_rate=self.initial_learning_rate,
            initial_steps=self.initial_steps,
            plateau_denominator=self.plateau_denominator,
            model_parameters_lr_scale=self.model_parameters_lr_scale,
            clip_value=self.clip_value,
            optimise_frame_weights=Optimisable_Parameters.frame_weights in self.parameter_partition_masks,
            optimise_model_parameters=Optimisable_Parameters.model_parameters in self.parameter_partition_masks,
        )

    # ------------------------------------------------------------------ initialise
    def initialise(self, model: Simulation, optimisable_funcs=None, _jit_test_args=None) -> OptimizationState:
        """z0 = w_init * n_frames, fresh Adam state, histories reset (reference :221-304).  With
        `_jit_test_args=(data_targets, loss_functions, indexes)` the engine is built right away."""
        p = model.params
        n = int(np.shape(p.frame_weights)[0])
        params = _replace(p, frame_weights=to_device(p.frame_weights, model.device) * n)
        self.history = OptimizationHistory()
        if self.save_ema_history:
            self.ema_history = OptimizationHistory()
        self._current_lr = float(self.initial_learning_rate)
        self._current_model_lr = float(self.initial_learning_rate * self.model_parameters_lr_scale)
        self._current_gradient_mask_idx = 0
        self._engine, self._engine_key = None, None
        self._init_params = params
        self.step = self._step
        if _jit_test_args is not None:
            data_targets, loss_functions, indexes = _jit_test_args
            self._ensure_engine(model, tuple(data_targets), tuple(loss_functions), tuple(indexes), params)
        return OptimizationState(params=params, opt_state=B200OptState(self._engine, 0))

    def _ensure_engine(self, simulation, data_targets, loss_functions, indexes, params: Simulation_Parameters) -> BvEngine:
        key = (tuple(data_targets), tuple(getattr(f, "__name__", repr(f)) for f in loss_functions), tuple(indexes),
               tuple(_target_fingerprint(t) for t in data_targets))  # strong references: compared by identity below
        if self._engine is not None and self._engine_key is not None and self._engine_key[1:] == key[1:] and \
                len(self._engine_key[0]) == len(key[0]) and all(a is b for a, b in zip(self._engine_key[0], key[0])):
            return self._engine
        if self._engine is not None:
            self._engine.close(barrier=False)
        eng = _build_engine(simulation, data_targets, loss_functions, indexes, self._settings())
        mp = params.model_parameters[0]
        eng.init_state(to_device(params.frame_weights, simulation.device).reshape(-1), scalar(mp.bv_bc), scalar(mp.bv_bh),
                       to_numpy(params.forward_model_weights), to_numpy(params.forward_model_scaling))
        self._engine, self._engine_key = eng, key
        self._n_losses = len(loss_functions)
        return eng

    # ------------------------------------------------------------------ step
    @staticmethod
    def _step(optimizer: "OptaxOptimizer", state: OptimizationState, simulation: Simulation, data_targets: tuple,
              loss_functions: tuple, indexes: tuple):
        eng = optimizer._ensure_engine(simulation, tuple(data_targets), tuple(loss_functions), tuple(indexes), state.params)
        k = eng.steps_done
        if int(state.step) != k or (isinstance(state.opt_state, B200OptState) and state.opt_state.engine not in (None, eng)):
            raise RuntimeError(
                f"state is at step {int(state.step)} but the device-resident optimiser state is at step {k}: "
                "pass the state returned by the previous step (the Adam moments live on the device)"
            )
        eng.step(1)
        return optimizer._materialise(eng, state, simulation, k)

    def _materialise(self, eng: BvEngine, state: OptimizationState, simulation: Simulation, k: int):
        """The reference's return tuple for the step that evaluated the loss at state.step == k."""
        n = self._n_losses
        losses = _losses_from_record(_record_view(eng, k), n)
        bufs, _ = _export_device(eng)
        mp_old = state.params.model_parameters[0]
        nxt = _record_view(eng, k + 1)
        bc = nxt[4:5].clone()
        bh = nxt[5:6].clone()
        mp_new = mp_old.update_parameters(type(mp_old)(bv_bc=bc, bv_bh=bh, temperature=mp_old.temperature, timepoints=mp_old.timepoints))
        rec = _record_view(eng, k)
        g_model = type(mp_old)(bv_bc=rec[32:33].clone(), bv_bh=rec[33:34].clone(), temperature=mp_old.temperature, timepoints=mp_old.timepoints)
        zeros_l = torch.zeros(n, device=eng.device)
        grads = Simulation_Parameters(
            frame_weights=bufs["g"],
            frame_mask=torch.zeros_like(bufs["g"]),
            model_parameters=[g_model],
            forward_model_weights=zeros_l,
            normalise_loss_functions=zeros_l,
            forward_model_scaling=zeros_l,
        )
        new_params = _replace(state.params, frame_weights=bufs["z"], model_parameters=[mp_new])
        opt_state = B200OptState(eng, k + 1)
        new_state = OptimizationState(new_params, opt_state, k + 1, losses, grads)
        save_params = _replace(Simulation_Parameters.normalize_weights(new_params), frame_weights=eng.weights.clone())
        save_state = OptimizationState(save_params, opt_state, k + 1, losses, grads)
        # outputs at the parameters the loss was evaluated at are not kept by the fused step (the engine already
        # holds the forward of the NEW parameters); the updated simulation carries the new forward.
        updated_sim = simulation._shallow_copy()
        updated_sim.params = save_params
        updated_sim.outputs = _outputs_of(eng)
        self._sync_host_scalars = (nxt[2:3], nxt[1:2])  # lr, branch of the update just applied (device views)
        return new_state, losses.total_train_loss, save_state, updated_sim

    @staticmethod
    def update_history_compute_ema_loss(optimizer, simulation, data_targets, loss_functions, indexes, state, ema_params=None):
        """history.add_state(state) and, when EMA parameters are tracked, their loss (reference :581-605)."""
        optimizer.history.add_state(state)
        if optimizer.save_ema_history and optimizer.ema_history is not None and ema_params is not None:
            losses, _ = compute_loss(simulation, ema_params, data_targets, indexes, loss_functions)
            optimizer.ema_history.add_state(OptimizationState(ema_params, state.opt_state, state.step, losses, state.gradients))
        return optimizer


def _export_device(eng: BvEngine):
    bufs, sc = {}, None
    with torch.cuda.device(eng.device):
        bufs = {k: torch.empty(eng.F, dtype=torch.float32, device=eng.device) for k in ("z", "g")}
        L.check(eng.lib.jaxent_bv_export_state(eng.plan, bufs["z"].data_ptr(), None, None, bufs["g"].data_ptr(), None,
                                               int(torch.cuda.current_stream().cuda_stream)))
    return bufs, sc
