"""HDF5 persistence of optimisation histories (reference `jaxent/src/utils/hdf.py:29-427`).

Same file layout as the reference, so histories written here are read by the reference's `load_optimization_history_from_file`
/ analysis scripts and vice versa:

    optimization_history/                      attrs: has_best_state
        states/<i>/                            attrs: step, has_losses
            params/{frame_weights, frame_mask, forward_model_weights, normalise_loss_functions, forward_model_scaling}
            params/model_parameters/<j>/       attrs: class_info, key; one dataset per dynamic slot (bv_bc, bv_bh)
            gradients/...                      (only when the state carries gradients)
            losses/{train_losses, val_losses, scaled_train_losses, scaled_val_losses, total_train_loss, total_val_loss}
        best_state/...

`class_info` is stored with the reference's module path (`jaxent.src.…`) and mapped back to this package on load.
h5py is used when it can be imported; otherwise the built-in writer / reader of `_h5mini.py` (real HDF5, 'earliest' file
format, contiguous datasets -- `compress` is then accepted and ignored)."""
from __future__ import annotations

import importlib
from typing import Optional

import numpy as np

from .._arrays import to_numpy
from ..interfaces.model import Model_Parameters
from ..interfaces.simulation import Simulation_Parameters
from ..opt.base import LossComponents, OptimizationHistory, OptimizationState
from . import _h5mini

try:  # pragma: no cover - not installed in the build image
    import h5py as _h5py
except Exception:  # noqa: BLE001
    _h5py = None

_OURS, _THEIRS = "jaxent_b200.", "jaxent.src."


def _open(filename: str, mode: str):
    return _h5py.File(filename, mode) if _h5py is not None else _h5mini.File(filename, mode)


def backend() -> str:
    return "h5py" if _h5py is not None else "builtin"


def save_array_to_hdf5(h5file, path: str, array, **kwargs) -> None:
    """reference :29-50 -- scalars are stored without chunking / compression"""
    a = to_numpy(array)
    if a.shape == () or a.size <= 1:  # h5py cannot chunk scalars (nor empty arrays)
        kwargs = {k: v for k, v in kwargs.items() if k not in ("compression", "compression_opts", "chunks")}
    h5file.create_dataset(path, data=a, **kwargs)


def load_array_from_hdf5(h5file, path: str):
    return np.asarray(h5file[path][()])


def save_model_parameters_to_hdf5(h5file, path: str, model_params: Model_Parameters, **kwargs) -> None:
    """reference :66-92"""
    group = h5file.create_group(path)
    cls = model_params.__class__
    module = cls.__module__
    if module.startswith(_OURS):
        module = _THEIRS + module[len(_OURS):]
    group.attrs["class_info"] = f"{module}.{cls.__name__}"
    dynamic_slots, _static = model_params._get_grouped_slots()
    for slot in dynamic_slots:
        save_array_to_hdf5(group, slot, getattr(model_params, slot), **kwargs)
    group.attrs["key"] = str([str(k) for k in model_params.key])


def _resolve_class(class_info: str):
    if isinstance(class_info, bytes):
        class_info = class_info.decode("utf-8")
    module_name, class_name = class_info.rsplit(".", 1)
    if module_name.startswith(_THEIRS):
        module_name = _OURS + module_name[len(_THEIRS):]
    return getattr(importlib.import_module(module_name), class_name)


def load_model_parameters_from_hdf5(h5file, path: str, default_cls: Optional[type] = None) -> Model_Parameters:
    """reference :95-131 -- static slots keep the class defaults"""
    group = h5file[path]
    cls = default_cls if default_cls is not None else _resolve_class(group.attrs["class_info"])
    dynamic_slots, _static = cls._get_grouped_slots()
    return cls(**{slot: load_array_from_hdf5(group, slot) for slot in dynamic_slots})


_PARAM_ARRAYS = ("frame_weights", "frame_mask", "forward_model_weights", "normalise_loss_functions", "forward_model_scaling")


def save_simulation_parameters_to_hdf5(h5file, path: str, sim_params: Simulation_Parameters, **kwargs) -> None:
    """reference :134-160"""
    group = h5file.create_group(path)
    for name in _PARAM_ARRAYS:
        save_array_to_hdf5(group, name, getattr(sim_params, name), **kwargs)
    mp_group = group.create_group("model_parameters")
    for i, mp in enumerate(sim_params.model_parameters):
        save_model_parameters_to_hdf5(mp_group, f"{i}", mp, **kwargs)


def load_simulation_parameters_from_hdf5(h5file, path: str, default_model_params_cls: Optional[type] = None) -> Simulation_Parameters:
    """reference :163-203"""
    group = h5file[path]
    arrays = {name: load_array_from_hdf5(group, name) for name in _PARAM_ARRAYS}
    mp_group = group["model_parameters"]
    model_params = [load_model_parameters_from_hdf5(mp_group, f"{i}", default_model_params_cls) for i in range(len(mp_group))]
    return Simulation_Parameters(model_parameters=model_params, **arrays)


_LOSS_FIELDS = LossComponents._fields


def save_loss_components_to_hdf5(h5file, path: str, loss_comps: LossComponents, **kwargs) -> None:
    """reference :206-225"""
    group = h5file.create_group(path)
    for name in _LOSS_FIELDS:
        save_array_to_hdf5(group, name, getattr(loss_comps, name), **kwargs)


def load_loss_components_from_hdf5(h5file, path: str) -> LossComponents:
    group = h5file[path]
    return LossComponents(**{name: load_array_from_hdf5(group, name) for name in _LOSS_FIELDS})


def save_optimization_state_to_hdf5(h5file, path: str, state: OptimizationState, **kwargs) -> None:
    """reference :261-294 -- the optimiser state itself is not persisted"""
    group = h5file.create_group(path)
    save_simulation_parameters_to_hdf5(group, "params", state.params, **kwargs)
    if state.opt_state is not None and state.gradients is not None:
        save_simulation_parameters_to_hdf5(group, "gradients", state.gradients, **kwargs)
    group.attrs["step"] = int(state.step)
    if state.losses is not None:
        save_loss_components_to_hdf5(group, "losses", state.losses, **kwargs)
        group.attrs["has_losses"] = True
    else:
        group.attrs["has_losses"] = False


def load_optimization_state_from_hdf5(h5file, path: str, default_model_params_cls: Optional[type] = None) -> OptimizationState:
    """reference :297-331"""
    group = h5file[path]
    params = load_simulation_parameters_from_hdf5(group, "params", default_model_params_cls)
    losses = load_loss_components_from_hdf5(group, "losses") if group.attrs["has_losses"] else None
    return OptimizationState(params=params, opt_state=None, step=int(group.attrs["step"]), losses=losses)


def save_optimization_history_to_hdf5(h5file, path: str, history: OptimizationHistory, **kwargs) -> None:
    """reference :334-360"""
    group = h5file.create_group(path)
    states_group = group.create_group("states")
    for i, state in enumerate(history.states):
        save_optimization_state_to_hdf5(states_group, f"{i}", state, **kwargs)
    if history.best_state is not None:
        save_optimization_state_to_hdf5(group, "best_state", history.best_state, **kwargs)
        group.attrs["has_best_state"] = True
    else:
        group.attrs["has_best_state"] = False


def load_optimization_history_from_hdf5(h5file, path: str, default_model_params_cls: Optional[type] = None) -> OptimizationHistory:
    """reference :363-393"""
    group = h5file[path]
    states_group = group["states"]
    states = [load_optimization_state_from_hdf5(states_group, f"{i}", default_model_params_cls) for i in range(len(states_group))]
    best = load_optimization_state_from_hdf5(group, "best_state", default_model_params_cls) if group.attrs["has_best_state"] else None
    return OptimizationHistory(states=states, best_state=best)


def save_optimization_history_to_file(filename: str, history: OptimizationHistory, compress: bool = True) -> None:
    """reference :396-414"""
    kwargs = {"compression": "gzip", "compression_opts": 9, "chunks": True} if compress else {}
    with _open(filename, "w") as f:
        save_optimization_history_to_hdf5(f, "optimization_history", history, **kwargs)


def load_optimization_history_from_file(filename: str, default_model_params_cls: Optional[type] = None) -> OptimizationHistory:
    """reference :417-427"""
    with _open(filename, "r") as f:
        return load_optimization_history_from_hdf5(f, "optimization_history", default_model_params_cls)
