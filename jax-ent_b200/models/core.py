"""Simulation (reference jaxent/src/models/core.py:21-315) over the fused CUDA engine.

Same constructor, `initialise()`, static `forward(sim, params, mutate)`, `.params`, `.outputs`,
`.outputs_by_key`.  The path covers ONE Best-Vendruscolo forward model per Simulation
(`BV_ForwardPass` or `BV_uptake_ForwardPass`, `average_first=True`); other forward models stay on the
reference's JAX path.  `initialise()` packs the (R,F) contact arrays into the tile layout once
(this replaces `cast_to_jax`, custom_types/features.py:82-89)."""
from __future__ import annotations

import logging
from typing import Any, Optional, Sequence

import numpy as np
import torch

from .. import _lib as L
from .._arrays import scalar, to_device, to_numpy
from ..engine import BvEngine, OptSettings, PackedFeatures, SlotSpec
from ..interfaces.simulation import Simulation_Parameters, _replace
from .HDX.BV.features import BV_input_features, BV_output_features, uptake_BV_output_features

LOGGER = logging.getLogger("jaxent.models")


def _forwardpass_of(model):
    return getattr(model, "forwardpass", model)


class Simulation:
    """The core object used during optimisation."""

    def __init__(self, input_features: Sequence[Any], forward_models: Sequence[Any], params: Optional[Simulation_Parameters],
                 raise_jit_failure: bool = False, device: Optional[str] = None) -> None:
        self.input_features = input_features
        self.forward_models = forward_models
        self.params = params
        self.forwardpass = tuple(_forwardpass_of(m) for m in forward_models)
        self.outputs: tuple = tuple()
        self.raise_jit_failure = raise_jit_failure
        self.device = torch.device(device) if device is not None else None
        self._features: Optional[PackedFeatures] = None
        self._fwd_engine: Optional[BvEngine] = None

    def __repr__(self) -> str:
        return f"Simulation(raise_jit_failure={self.raise_jit_failure})"

    # ------------------------------------------------------------------ initialise
    def initialise(self) -> bool:
        lengths = [f.features_shape[-1] for f in self.input_features]
        if len(set(lengths)) != 1:
            raise AssertionError("Input features have different frame counts")
        self.length = lengths[0]
        if self.params is None:
            raise ValueError("No simulation parameters were provided. Exiting.")
        if len(self.forward_models) != len(self.params.model_parameters):
            raise AssertionError("Number of forward models must equal number of model parameters")
        if len(self.forward_models) != 1 or not isinstance(self.input_features[0], BV_input_features):
            raise NotImplementedError("the fused B200 path covers exactly one BV forward model per Simulation")
        fp = self.forwardpass[0]
        if not hasattr(fp, "uptake"):
            raise NotImplementedError(f"forward pass {type(fp).__name__} is not a BV forward pass of the fused path")
        if not torch.cuda.is_available():
            raise RuntimeError("jaxent_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if self.device is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        feat = self.input_features[0]
        if tuple(np.shape(feat.heavy_contacts)) != tuple(np.shape(feat.acceptor_contacts)) or len(np.shape(feat.heavy_contacts)) != 2:
            raise ValueError("BV features must be two (n_residues, n_frames) arrays of equal shape")
        self._uptake = bool(fp.uptake)
        # average_first=False: the uptake is evaluated per frame and the outputs are frame-averaged (reference :302-304);
        # for the linear log_Pf forward pass the two orders coincide, so only the uptake pass has a per-frame kernel
        self._average_first = bool(getattr(fp, "average_first", True)) or not self._uptake
        mp = self.params.model_parameters[0]
        self._timepoints = to_numpy(mp.timepoints).reshape(-1) if self._uptake else None
        self._k_ints = to_numpy(feat.k_ints).reshape(-1) if (self._uptake and feat.k_ints is not None) else None
        if self._uptake and self._k_ints is None:
            raise ValueError("Failed to apply forward models without JIT: BV_uptake_ForwardPass needs k_ints")
        try:
            self._fwd_engine = BvEngine(feat.heavy_contacts, feat.acceptor_contacts, self._k_ints, self._timepoints,
                                        [SlotSpec(L.LOSS_PARAMS_L2)], self._uptake, OptSettings(optimise_model_parameters=False),
                                        device=self.device, average_first=self._average_first)
        except Exception as e:
            raise ValueError(f"Failed to apply forward models without JIT: {e}")
        self._features = self._fwd_engine.features
        # the engine has started the host -> device transfer of the features (the one large input of a job); the host-side
        # parameter normalisation runs while it is in flight
        self.params = Simulation_Parameters.normalize_weights(self.params)
        self.params = Simulation_Parameters.normalize_masked_loss_scalingweights(self.params)
        # the reference re-applies the softmax here (models/core.py:72,97-98): params end up softmax(softmax(w0))
        _sim, _outputs = Simulation.forward(self, self.params, mutate=False)
        self.params = _sim.params
        self.outputs = _outputs
        LOGGER.info("Simulation initialised successfully.")
        return True

    # ------------------------------------------------------------------ forward
    def _forward_logits(self, logits, norm_params: Simulation_Parameters):
        eng = self._fwd_engine
        mp = norm_params.model_parameters[0]
        eng.init_state(to_device(logits, self.device).reshape(-1), scalar(mp.bv_bc), scalar(mp.bv_bh), [0.0], [0.0])
        if self._uptake:
            out = uptake_BV_output_features(uptake=eng.uptake_out.clone())
        else:
            out = BV_output_features(log_Pf=eng.log_pf.clone(), k_ints=None)
        new_sim = self._shallow_copy()
        new_sim.params = _replace(norm_params, frame_weights=eng.weights.clone())
        new_sim.outputs = (out,)
        return new_sim, (out,)

    @staticmethod
    def forward(sim: "Simulation", params: Simulation_Parameters, mutate: bool = True):
        """normalize_weights(params) (softmax of params.frame_weights) -> frame average -> BV forward pass.
        Returns (new_sim, outputs) when mutate=False, else updates and returns `sim` (:131-163)."""
        if sim._fwd_engine is None:
            raise RuntimeError("Simulation must be initialised before calling forward")
        norm = Simulation_Parameters.normalize_weights(params)
        new_sim, outputs = sim._forward_logits(params.frame_weights, norm)
        if mutate:
            sim.params, sim.outputs = new_sim.params, new_sim.outputs
            return sim
        return new_sim, outputs

    def predict(self, params):
        """Forward passes on the un-averaged (frame-wise) features: every output keeps the frame axis last
        (reference models/core.py:181-208).  log_Pf (R, F) or uptake (T, R, F)."""
        if self._fwd_engine is None:
            raise RuntimeError("Simulation must be initialized before calling predict")
        model_parameters = params.model_parameters if isinstance(params, Simulation_Parameters) else params
        if len(model_parameters) != len(self.forwardpass):
            raise ValueError(f"Number of model parameters ({len(model_parameters)}) must match number of forward models ({len(self.forwardpass)})")
        feats = self._features
        if getattr(feats, "format", 0) != 0:
            raise NotImplementedError("Simulation.predict reads the fp32 feature layout")
        mp = model_parameters[0]
        eng = self._fwd_engine
        R, F, T = eng.R, eng.F, (eng.T if self._uptake else 0)
        with torch.cuda.device(self.device):
            out = torch.empty((T, R, F) if T else (R, F), dtype=torch.float32, device=self.device)
            L.check(eng.lib.jaxent_bv_predict_frames(feats.packed.data_ptr(), R, F, T,
                                                     eng.k_ints.data_ptr() if T else None, eng.timepoints.data_ptr() if T else None,
                                                     scalar(mp.bv_bc), scalar(mp.bv_bh), out.data_ptr(),
                                                     int(torch.cuda.current_stream().cuda_stream)))
        if T:
            return [uptake_BV_output_features(uptake=out)]
        return [BV_output_features(log_Pf=out, k_ints=None)]

    @property
    def outputs_by_key(self) -> dict:
        result = {}
        for o in self.outputs:
            if o.key in result:
                raise KeyError(f"Duplicate output key '{o.key}' — each forward model must produce outputs with unique m_keys.")
            result[o.key] = o
        return result

    def _shallow_copy(self) -> "Simulation":
        new = Simulation.__new__(Simulation)
        new.__dict__.update(self.__dict__)
        return new
