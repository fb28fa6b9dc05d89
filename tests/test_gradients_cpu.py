"""opt/gradients.py mirrors reference jaxent/src/opt/gradients.py:8-130 (host-side mask pytrees)."""
import numpy as np
import pytest

from jaxent_b200 import Optimisable_Parameters
from jaxent_b200.interfaces.simulation import Simulation_Parameters
from jaxent_b200.models.HDX.BV.parameters import BV_Model_Parameters
from jaxent_b200.opt.base import OptimizationState
from jaxent_b200.opt.gradients import check_gradient_oscillation, create_gradient_masks, get_previous_grads, mask_gradients


def _params(scale=1.0):
    return Simulation_Parameters(
        frame_weights=np.arange(1, 6, dtype=np.float32) * scale, frame_mask=np.ones(5, np.float32),
        model_parameters=[BV_Model_Parameters(bv_bc=np.array([0.35 * scale], np.float32), bv_bh=np.array([2.0 * scale], np.float32))],
        forward_model_weights=np.array([1.0, 2.0], np.float32), normalise_loss_functions=np.array([1.0, 0.0], np.float32),
        forward_model_scaling=np.array([3.0, 4.0], np.float32))


def test_masks_follow_the_partition_set():
    p = _params()
    m = create_gradient_masks({Optimisable_Parameters.frame_weights}, p, None)
    assert m.frame_weights.tolist() == [1.0] * 5 and m.frame_mask.tolist() == [0.0] * 5
    assert float(m.model_parameters[0].bv_bc[0]) == 0.0 and m.model_parameters[0].temperature == p.model_parameters[0].temperature
    assert not m.forward_model_weights.any() and not m.forward_model_scaling.any() and not m.normalise_loss_functions.any()
    m = create_gradient_masks({Optimisable_Parameters.frame_weights, Optimisable_Parameters.model_parameters}, p, None)
    assert float(m.model_parameters[0].bv_bc[0]) == 1.0 and float(m.model_parameters[0].bv_bh[0]) == 1.0
    with pytest.raises(NotImplementedError, match="Frame mask optimization"):
        create_gradient_masks({Optimisable_Parameters.frame_mask}, p, None)


def test_mask_gradients_and_oscillation():
    g = _params(-2.0)
    m = create_gradient_masks({Optimisable_Parameters.model_parameters}, g, None)
    mg = mask_gradients(g, m)
    assert not mg.frame_weights.any() and float(mg.model_parameters[0].bv_bc[0]) == pytest.approx(-0.7)
    assert mg.forward_model_weights is m.forward_model_weights  # loss-weight leaves are replaced by the zero masks (:126-128)
    a, b = _params(1.0), _params(-1.0)
    want = -(np.arange(1, 6) ** 2).sum() + 5 - 0.35 ** 2 - 4.0 + (1 + 4) + (9 + 16) + 1  # only frame weights and BV parameters are scaled
    assert check_gradient_oscillation(a, b) == pytest.approx(want, rel=1e-6)
    assert check_gradient_oscillation(a, None) == 0.0
    st = OptimizationState(a, None, 0, None, None)
    z = get_previous_grads(st)
    assert not z.frame_weights.any() and float(z.model_parameters[0].bv_bh[0]) == 0.0
    assert get_previous_grads(st._replace(gradients=b)) is b


def test_masks_on_torch_tensors():
    import torch

    p = _params()
    p = Simulation_Parameters(frame_weights=torch.tensor(p.frame_weights), frame_mask=torch.tensor(p.frame_mask), model_parameters=p.model_parameters,
                              forward_model_weights=p.forward_model_weights, normalise_loss_functions=p.normalise_loss_functions,
                              forward_model_scaling=p.forward_model_scaling)
    m = create_gradient_masks({Optimisable_Parameters.frame_weights}, p, None)
    assert isinstance(m.frame_weights, torch.Tensor) and m.frame_weights.dtype == torch.float32 and bool((m.frame_weights == 1).all())
    assert check_gradient_oscillation(p, p) > 0
