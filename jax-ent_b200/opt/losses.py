"""Loss functions of the path (reference jaxent/src/opt/losses.py), as tokens.

The reference evaluates these Python callables inside jax.value_and_grad.  Here each name is a token
that ``OptaxOptimizer`` / ``compute_loss`` map to the fused CUDA implementation (forward *and* the
hand-written backward); the token's position in ``loss_functions`` selects its forward_model_weights /
forward_model_scaling entry exactly as in the reference (opt/optimiser.py:621-635).  Calling a token on
the host raises: there is no CPU implementation.  Losses outside SURVEY.md section 8(a) are not defined
here and stay on the reference's stock JAX path."""
from __future__ import annotations

from .. import _lib as L


class _LossToken:
    def __init__(self, name: str, kind: int, doc: str):
        self.__name__ = name
        self.kind = kind
        self.__doc__ = doc

    def __call__(self, model, dataset, prediction_index):
        raise NotImplementedError(
            f"{self.__name__} is evaluated inside the fused CUDA step (use compute_loss / OptaxOptimizer.step); "
            "jaxent_b200 has no host implementation of the losses"
        )

    def __repr__(self):
        return f"<jaxent_b200 loss {self.__name__}>"


hdx_uptake_eye_MSE_loss = _LossToken("hdx_uptake_eye_MSE_loss", L.LOSS_UPTAKE_EYE_MSE, "opt/losses.py:1601-1657")
hdx_uptake_mean_centred_eye_MSE_loss = _LossToken("hdx_uptake_mean_centred_eye_MSE_loss", L.LOSS_UPTAKE_MC_EYE_MSE, "opt/losses.py:1790-1841")
hdx_uptake_sigma_MSE_loss = _LossToken("hdx_uptake_sigma_MSE_loss", L.LOSS_UPTAKE_SIGMA_MSE, "opt/losses.py:1541-1598")
hdx_pf_l2_loss = _LossToken("hdx_pf_l2_loss", L.LOSS_PF_L2, "opt/losses.py:17-50")
maxent_convexKL_loss = _LossToken("maxent_convexKL_loss", L.LOSS_MAXENT_CONVEX_KL, "opt/losses.py:304-322")
model_params_L1_loss = _LossToken("model_params_L1_loss", L.LOSS_PARAMS_L1, "opt/losses.py:495-513")
model_params_L2_loss = _LossToken("model_params_L2_loss", L.LOSS_PARAMS_L2, "opt/losses.py:476-492")

SUPPORTED = {t.__name__: t for t in (hdx_uptake_eye_MSE_loss, hdx_uptake_mean_centred_eye_MSE_loss, hdx_uptake_sigma_MSE_loss,
                                     hdx_pf_l2_loss, maxent_convexKL_loss, model_params_L1_loss, model_params_L2_loss)}


def loss_kind(fn) -> int:
    """token (or a reference function of the same name) -> C-ABI loss kind; ValueError otherwise."""
    k = getattr(fn, "kind", None)
    if isinstance(k, int):
        return k
    name = getattr(fn, "__name__", None)
    if name in SUPPORTED:
        return SUPPORTED[name].kind
    raise ValueError(
        f"loss function {name or fn!r} is not on the fused B200 path (supported: {', '.join(sorted(SUPPORTED))}); "
        "run it through the reference's stock JAX optimiser instead"
    )
