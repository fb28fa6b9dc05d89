"""Per-frame BV uptake mode (ForwardPass.average_first = False): uptake per (time point, residue, frame),
then the frame average -- reference models/core.py:302-304, models/HDX/forward.py:57-74.

The pass evaluates F*R*(T+1) exponentials per step on the special-function unit (ex2.approx).  Per-step tolerance against the
fp64 oracle: 1e-5 on every loss component and on the gradients (north_star's bound, the same as the average-first path;
measured 1.5e-7 / 9.5e-7), frame weights after a long run 1e-4.  BV-parameter optimisation in this mode: a sweep of its own
for d loss / d (bc, bh) + the forward half of the pass at the updated parameters."""
import numpy as np
import pytest
import torch

import helpers as H
from oracle import bv_oracle as O
from test_gpu_parity import _check_steps

pytestmark = pytest.mark.gpu

PF_RTOL = 1e-5


def _case(F, R, T, kind=O.UPTAKE_EYE_MSE, seed=4, **kw):
    case = H.make_case(F, R, T, kind, seed=seed, **kw)
    case["problem"].average_first = False
    return case


@pytest.mark.parametrize("kind,F,R,T,clip", [
    (O.UPTAKE_EYE_MSE, 777, 97, 5, None),
    (O.UPTAKE_MC_EYE_MSE, 1000, 53, 3, 1.0),
    (O.UPTAKE_SIGMA_MSE, 515, 64, 8, None),
    (O.UPTAKE_EYE_MSE, 4099, 500, 8, None),
])
def test_per_frame_step_parity(kind, F, R, T, clip):
    case = _case(F, R, T, kind, gamma=2.0, prior="random")
    cfg = O.OptConfig(learning_rate=1.0, initial_learning_rate=1.0, initial_steps=0, clip_value=clip, optimise_model_parameters=False)
    _, wl, wg = _check_steps(case, cfg, 6, STEP_RTOL=PF_RTOL)
    assert wl < 1e-6 and wg < 5e-6, (wl, wg)


def test_per_frame_collapsed_weights():
    """logit spread 10: 70 % of the frames below eps = 1e-10/F, the MaxEnt part of hbar is O(lambda)"""
    from test_gpu_parity import _collapsed_case

    case, frac = _collapsed_case(2000, 40, 4, O.UPTAKE_EYE_MSE, 10.0)
    case["problem"].average_first = False
    assert frac > 0.5
    cfg = O.OptConfig(learning_rate=1.0, clip_value=None, optimise_model_parameters=False)
    _, wl, wg = _check_steps(case, cfg, 6, STEP_RTOL=PF_RTOL, mirror_slack=True)
    assert wl < 1e-5, wl


def test_per_frame_forward_outputs_and_difference_to_average_first():
    case = _case(1500, 64, 4)
    cfg = O.OptConfig(learning_rate=0.5, clip_value=None, optimise_model_parameters=False)
    eng = H.make_engine(case, cfg)
    H.init_engine(eng, case)
    torch.cuda.synchronize()
    st = O.optimizer_initialise(case["params"], cfg, np.float64)
    _, out = O.forward(case["problem"], st.params, np.float64)
    np.testing.assert_allclose(eng.uptake_out.cpu().numpy(), out["uptake"], rtol=5e-6, atol=1e-7)
    case["problem"].average_first = True
    _, out_af = O.forward(case["problem"], st.params, np.float64)
    assert np.abs(out_af["uptake"] - out["uptake"]).max() > 1e-4  # the two orders differ for a non-linear forward pass


def test_per_frame_weights_after_500_steps():
    case = _case(600, 53, 3, seed=9)
    cfg = O.OptConfig(learning_rate=0.05, initial_learning_rate=1.0, initial_steps=2, clip_value=None, optimise_model_parameters=False)
    eng = H.make_engine(case, cfg)
    H.init_engine(eng, case)
    eng.step(500)
    torch.cuda.synchronize()
    st = O.optimizer_initialise(case["params"], cfg, np.float64)
    for _ in range(500):
        st, _, _ = O.step_with_rates(case["problem"], st, cfg, np.float64)
    w_ref = O.softmax(st.params.z, np.float64)
    np.testing.assert_allclose(eng.weights.cpu().numpy(), w_ref, atol=1e-4)


@pytest.mark.parametrize("kind,extra,F,R,T,clip,init_steps", [
    (O.UPTAKE_EYE_MSE, (O.PARAMS_L2,), 777, 97, 5, None, 0),
    (O.UPTAKE_MC_EYE_MSE, (O.PARAMS_L1,), 1000, 53, 3, 1.0, 2),
    (O.UPTAKE_SIGMA_MSE, (), 515, 64, 8, None, 0),
    (O.UPTAKE_EYE_MSE, (), 4099, 500, 8, None, 0),
])
def test_per_frame_step_parity_model_parameters(kind, extra, F, R, T, clip, init_steps):
    """BV-parameter optimisation with average_first = False (reference models/core.py:298-305 places no restriction;
    examples/common/optimization.py:267-296): d loss / d (bc, bh) = sum_f w_f sum_{t,r} G[t,r] du[t,r,f]/d(bc, bh) comes from a
    frame sweep of its own (pf_dbeta_kernel), and the forward half of the pass runs at the UPDATED parameters of each plateau
    branch.  Teacher-forced against the fp64 oracle: losses (evaluated at the updated bc, bh), frame gradients, grad_bc / grad_bh."""
    case = _case(F, R, T, kind, gamma=2.0, prior="random", extra=extra)
    cfg = O.OptConfig(learning_rate=0.2, initial_learning_rate=1.0, initial_steps=init_steps, clip_value=clip, optimise_model_parameters=True)
    eng, wl, wg = _check_steps(case, cfg, 7, STEP_RTOL=PF_RTOL)
    assert wl < 1e-6 and wg < 5e-6, (wl, wg)
    recs = eng.records(0, 8)
    assert recs[init_steps].grad_bc != 0.0 and recs[7].bc != recs[0].bc  # the BV parameters really move
    if init_steps:
        assert recs[0].grad_bc == 0.0 and recs[1].bc == recs[0].bc  # masked during the initial steps


def test_per_frame_model_parameters_trajectory():
    """free-running: 200 steps with the BV parameters optimised, against the fp64 oracle loop"""
    case = _case(600, 53, 3, seed=9, extra=(O.PARAMS_L2,))
    cfg = O.OptConfig(learning_rate=0.05, initial_learning_rate=1.0, initial_steps=2, clip_value=None, optimise_model_parameters=True)
    eng = H.make_engine(case, cfg)
    H.init_engine(eng, case)
    eng.step(200)
    torch.cuda.synchronize()
    st = O.optimizer_initialise(case["params"], cfg, np.float64)
    for _ in range(200):
        st, _, _ = O.step_with_rates(case["problem"], st, cfg, np.float64)
    np.testing.assert_allclose(eng.weights.cpu().numpy(), O.softmax(st.params.z, np.float64), atol=1e-4)
    sc = eng.export_state(())[1]
    np.testing.assert_allclose([sc["bc"], sc["bh"]], [st.params.bc, st.params.bh], rtol=1e-3)
    assert abs(sc["bc"] - case["params"].bc) > 1e-3  # moved away from the start
