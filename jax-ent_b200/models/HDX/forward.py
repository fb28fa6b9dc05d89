"""BV forward passes (reference jaxent/src/models/HDX/forward.py:15-76).

Inside ``Simulation`` these objects are *tokens*: ``average_first = True`` selects the fused CUDA
path (softmax-weighted frame average, then ln PF / uptake).  They are not callable on the host --
this package has no CPU implementation of the model."""
from __future__ import annotations

from ...custom_types.key import m_key


class _FusedForwardPass:
    average_first: bool = True
    uptake: bool = False
    key = None

    def __call__(self, input_features, parameters):
        raise NotImplementedError(
            f"{type(self).__name__} is evaluated by the fused CUDA step (Simulation.forward / OptaxOptimizer.step); "
            "jaxent_b200 has no host implementation of the forward model"
        )


class BV_ForwardPass(_FusedForwardPass):
    """ln PF_r = bc * <heavy>_r + bh * <acceptor>_r  -> BV_output_features (models/HDX/forward.py:15-37)."""

    uptake = False
    key = m_key("HDX_resPF")


class BV_uptake_ForwardPass(_FusedForwardPass):
    """uptake[t,r] = 1 - exp(-k_r t / PF_r) -> uptake_BV_output_features (models/HDX/forward.py:40-76)."""

    uptake = True
    key = m_key("HDX_peptide")
