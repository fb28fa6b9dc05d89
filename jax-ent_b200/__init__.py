"""jaxent_b200 -- B200-native (sm_100a) BV MaxEnt optimise step behind JAX-ENT's API.

The directory is named ``jax-ent_b200`` (the framework name); ``import jaxent_b200`` resolves
to it through the shim package next to it.  Only the hot path of SURVEY.md section 8 lives here.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401

# reference-API mirror of the hot path (same names as jaxent.src.*)
from .custom_types.config import Optimisable_Parameters, OptimiserSettings  # noqa: E402,F401
from .data.loader import Dataset, ExpD_Dataloader, SparseFragmentMapping  # noqa: E402,F401
from .interfaces.simulation import Simulation_Parameters  # noqa: E402,F401
from .models.core import Simulation  # noqa: E402,F401
from .models.HDX.BV.features import BV_input_features, BV_output_features, uptake_BV_output_features  # noqa: E402,F401
from .models.HDX.BV.parameters import BV_Model_Parameters  # noqa: E402,F401
from .models.HDX.forward import BV_ForwardPass, BV_uptake_ForwardPass  # noqa: E402,F401
from .opt.base import LossComponents, OptimizationHistory, OptimizationState  # noqa: E402,F401
from .opt.losses import (  # noqa: E402,F401
    hdx_pf_l2_loss,
    hdx_uptake_eye_MSE_loss,
    hdx_uptake_mean_centred_eye_MSE_loss,
    hdx_uptake_sigma_MSE_loss,
    maxent_convexKL_loss,
    model_params_L1_loss,
    model_params_L2_loss,
)
from .opt.optimiser import OptaxOptimizer, compute_loss  # noqa: E402,F401
from .opt.batch import BatchOptimisationResult, HParamBatch, batch_optimise  # noqa: E402,F401
from .opt.run import _optimise, _optimise_device, _optimise_pure, run_optimise  # noqa: E402,F401
