"""CPU tests of the host-side mirror of the reference interface (no compute calls)."""
import numpy as np
import pytest
import torch

import jaxent_b200 as J
from jaxent_b200.opt.losses import loss_kind
from jaxent_b200.opt.track import ConvergenceTracker, create_convergence_thresholds
from oracle import bv_oracle as O


def _params(F=16, L=3):
    return J.Simulation_Parameters(
        frame_weights=np.linspace(0.1, 2.0, F).astype(np.float32),
        frame_mask=np.ones(F, np.float32),
        model_parameters=[J.BV_Model_Parameters()],
        forward_model_weights=np.array([2.0, 1.0, 5.0], np.float32)[:L],
        normalise_loss_functions=np.array([1.0, 1.0, 0.0], np.float32)[:L],
        forward_model_scaling=np.full(L, 100.0, np.float32),
    )


def test_simulation_parameters_normalisers_match_reference_semantics():
    p = _params()
    q = J.Simulation_Parameters.normalize_masked_loss_scalingweights(p)
    np.testing.assert_allclose(q.forward_model_weights, [2 / 3, 1 / 3, 5.0], rtol=1e-6)  # simulation.py:45-76
    np.testing.assert_allclose(q.forward_model_weights, O.normalize_masked_loss_scalingweights(p.forward_model_weights, p.normalise_loss_functions))
    n = J.Simulation_Parameters.normalize_weights(p)
    np.testing.assert_allclose(n.frame_weights, O.softmax(p.frame_weights), rtol=1e-6)
    np.testing.assert_allclose(n.frame_mask, 1 / (1 + np.exp(-5.0)), rtol=1e-6)
    assert n.model_parameters[0] is p.model_parameters[0]  # passed through unclamped
    nt = J.Simulation_Parameters.normalize_weights(J.Simulation_Parameters(
        torch.tensor(p.frame_weights), torch.ones(16), p.model_parameters, p.forward_model_weights, p.normalise_loss_functions, p.forward_model_scaling))
    np.testing.assert_allclose(nt.frame_weights.numpy(), n.frame_weights, rtol=1e-6)
    with pytest.raises(ValueError):
        J.Simulation_Parameters.normalize_weights(J.Simulation_Parameters(np.ones((2, 2)), np.ones((2, 2)), [], None, None, None))


def test_simulation_parameters_arithmetic():
    p = _params()
    s = 0.5 * p + 0.5 * p
    np.testing.assert_allclose(s.frame_weights, p.frame_weights, rtol=1e-6)
    np.testing.assert_allclose(s.model_parameters[0].bv_bc, p.model_parameters[0].bv_bc)
    d = p - p
    assert float(np.abs(d.frame_weights).max()) == 0.0 and float(d.model_parameters[0].bv_bh[0]) == 0.0
    with pytest.raises(TypeError):
        p + "x"


def test_bv_model_parameters_pytree_and_update():
    mp = J.BV_Model_Parameters(bv_bc=np.array([0.4]), bv_bh=np.array([1.9]), timepoints=np.array([1.0, 2.0]))
    arrays, static = mp.tree_flatten()
    assert len(arrays) == 2 and set(J.BV_Model_Parameters.static_params) == {"temperature", "key", "timepoints"}
    mp2 = J.BV_Model_Parameters.tree_unflatten(static, arrays)
    assert float(mp2.bv_bc[0]) == 0.4 and np.array_equal(mp2.timepoints, mp.timepoints)
    mp3 = mp.update_parameters(J.BV_Model_Parameters(bv_bc=np.array([1.0]), bv_bh=np.array([3.0]), timepoints=np.array([9.0])))
    assert float(mp3.bv_bc[0]) == 1.0 and np.array_equal(mp3.timepoints, mp.timepoints)  # static fields preserved


def test_feature_containers():
    f = J.BV_input_features(np.zeros((5, 7)), np.zeros((5, 7)), np.ones(5))
    assert f.features_shape == (5, 7) and f.__features__ == {"heavy_contacts", "acceptor_contacts", "k_ints"}
    assert J.BV_input_features(np.zeros(5), np.zeros(5)).features_shape == (5, 1)
    with pytest.raises(TypeError):
        J.BV_input_features(np.zeros((5, 7)), torch.zeros(5, 7)).features_shape
    assert J.uptake_BV_output_features(np.zeros((3, 5))).output_shape == (3, 5)
    assert J.BV_output_features(np.zeros(5)).output_shape == (1, 5)
    assert J.BV_ForwardPass.average_first and J.BV_uptake_ForwardPass.average_first


def test_loss_tokens():
    assert J.hdx_uptake_eye_MSE_loss.__name__ == "hdx_uptake_eye_MSE_loss"
    assert loss_kind(J.maxent_convexKL_loss) == 4

    def hdx_pf_l2_loss(model, dataset, idx):  # a reference function of the same name maps to the same kernel
        return 0, 0

    assert loss_kind(hdx_pf_l2_loss) == 3

    def maxent_JSD_loss(model, dataset, idx):
        return 0, 0

    with pytest.raises(ValueError, match="not on the fused B200 path"):
        loss_kind(maxent_JSD_loss)
    with pytest.raises(NotImplementedError, match="no host implementation"):
        J.hdx_pf_l2_loss(None, None, 0)
    with pytest.raises(NotImplementedError, match="no host implementation"):
        J.BV_ForwardPass()(None, None)


def test_optimizer_argument_validation():
    with pytest.raises(ValueError, match="Unsupported optimizer"):
        J.OptaxOptimizer(optimizer="nadam")
    with pytest.raises(NotImplementedError, match="only 'adam'"):
        J.OptaxOptimizer(optimizer="sgd")
    with pytest.raises(NotImplementedError, match="Frame mask optimization"):
        J.OptaxOptimizer(parameter_partition_masks={J.Optimisable_Parameters.frame_mask})
    o = J.OptaxOptimizer(learning_rate=0.3, initial_learning_rate=2.0, model_parameters_lr_scale=0.5)
    assert o.current_learning_rate == 2.0 and o.current_model_learning_rate == 1.0
    s = o._settings()
    assert s.optimise_frame_weights and not s.optimise_model_parameters and s.clip_value == 1.0


def test_convergence_tracker_matches_oracle_tracker():
    assert create_convergence_thresholds([1e-3, 1e-1, 1e-2], 0.5) == pytest.approx([0.05, 0.005, 0.0005])
    a = ConvergenceTracker([1e-2, 1e-1], 1.0, 0.5, 2)
    b = O.ConvergenceTracker([1e-2, 1e-1], 1.0, 0.5, 2)
    rng = np.random.default_rng(0)
    prev = None
    loss = 10.0
    for step in range(40):
        loss = loss * (0.5 + 0.5 * rng.random())
        w = rng.random(4)
        a.update(prev, loss, w)
        b.update(prev, loss, w)
        prev = loss
        assert a.is_threshold_met(loss, step, 0) == b.met(loss, step, 0)
        assert a.get_relative_convergence(loss) == pytest.approx(b.rel(loss))
        if a.is_threshold_met(loss, step, 0):
            assert a.advance_threshold() == b.advance()
    np.testing.assert_allclose(a.ema_params, b.ema_params)


def test_functional_tracker_differs_from_the_class_tracker_where_the_reference_does():
    """opt/track.py:24-125 (the pure loop's carry) against the ConvergenceTracker class (:128-197) on one loss sequence: the
    functional tracker starts its EMA at the FIRST iteration (raw delta 0 against the dense initial loss), restarts it after
    every threshold advance, holds the EMA of the parameters the same way, and ends `converged` without advancing the index."""
    from jaxent_b200.opt import track as T

    losses = np.asarray([10.0, 6.0, 4.0, 3.5, 3.4, 3.39, 3.389, 3.3889], np.float32)
    thr = T.create_convergence_thresholds([0.5, 0.05], 1.0)
    carry = T.initialise_convergence_carry(np.zeros(3))
    prev, trace = losses[0], []
    for i, l in enumerate(losses):
        before = carry
        carry, raw = T.update_convergence(carry, prev, l, np.full(3, float(i)), 0.5)
        if before.steps_since_threshold_start == 0:
            assert carry.ema_loss_delta == raw and np.all(carry.ema_params == float(i))  # restart
        else:
            assert carry.ema_loss_delta == pytest.approx(0.5 * raw + 0.5 * before.ema_loss_delta)
            np.testing.assert_allclose(carry.ema_params, 0.5 * i + 0.5 * before.ema_params)
        carry = T.check_and_advance_threshold(carry, l, i + 1, thr, 2, 0)
        trace.append((carry.current_threshold_idx, carry.steps_since_threshold_start, bool(carry.converged)))
        prev = l
        if carry.converged:
            break
    assert trace[0] == (0, 1, False)            # first iteration: raw = 0 -> relative convergence 0 < thr, but only 1 step
    assert trace[1][0] == 1 and trace[1][1] == 0  # second iteration: min_steps reached with ema = 0.5*4 + 0.5*0 = 2, 2/6 < 0.5 -> advance
    assert carry.converged and carry.current_threshold_idx == 1  # the last threshold ends the loop without advancing the index
    # the class tracker on the same sequence has no EMA at all after its first update (previous_loss is None there)
    c = ConvergenceTracker([0.5, 0.05], 1.0, 0.5, 2)
    c.update(None, float(losses[0]), np.zeros(3))
    assert c.ema_loss_delta is None and not c.is_threshold_met(float(losses[0]), 0, 0)


def test_simulation_requires_cuda_and_valid_inputs():
    f = J.BV_input_features(np.zeros((5, 8), np.float32), np.zeros((5, 8), np.float32), np.ones(5, np.float32))
    p = J.Simulation_Parameters(np.full(8, 1 / 8, np.float32), np.ones(8, np.float32), [J.BV_Model_Parameters()],
                                np.ones(1, np.float32), np.ones(1, np.float32), np.ones(1, np.float32))
    sim = J.Simulation([f], [J.BV_uptake_ForwardPass()], None)
    with pytest.raises(ValueError, match="No simulation parameters"):
        sim.initialise()
    with pytest.raises(AssertionError, match="must equal number of model parameters"):
        J.Simulation([f], [J.BV_ForwardPass(), J.BV_ForwardPass()], p).initialise()
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            J.Simulation([f], [J.BV_uptake_ForwardPass()], p).initialise()
    with pytest.raises(RuntimeError, match="initialised"):
        J.Simulation.forward(J.Simulation([f], [J.BV_ForwardPass()], p), p)


def test_dataloader_containers():
    S = np.eye(3, 5, dtype=np.float32)
    ds = J.Dataset([], np.zeros((3, 2, 1)), J.SparseFragmentMapping(S))
    assert ds.residue_feature_ouput_mapping is S and ds.covariance_matrix is None
    np.testing.assert_array_equal(J.SparseFragmentMapping(torch.tensor(S)).dense(), S)
    dl = J.ExpD_Dataloader(train=ds, val=ds)
    assert dl.train is ds and dl.key == "unknown"
    with pytest.raises(NotImplementedError):
        J.ExpD_Dataloader(data=[1])


def test_optimise_pure_host_loop_on_a_scripted_optimizer(monkeypatch):
    """opt/run.py::_optimise_pure (the loop of the reference's _optimise_pure, opt/run.py:141-232, driven from the host) with the
    fused step replaced by a scripted one: where the loop ends, which entry is returned (_pick_best_state: the last entry unless
    an earlier one is strictly better, then the first that reaches the minimum), and that only the best and the last entry keep
    their frame weights."""
    import types

    from jaxent_b200.opt import optimiser as OPT
    from jaxent_b200.opt import run as RUN
    from jaxent_b200.opt.base import LossComponents, OptimizationState

    class P:  # stands in for Simulation_Parameters: only _replace(frame_weights=...) is used on it
        def __init__(self, w):
            self.frame_weights = w

    monkeypatch.setattr(RUN, "_replace", lambda p, **kw: P(kw.get("frame_weights", p.frame_weights)))

    def scripted(losses, val0):
        def lc(i):
            return LossComponents(np.array([losses[i]]), np.array([val0[i]]), np.array([losses[i]]), np.array([val0[i]]), losses[i], val0[i])

        monkeypatch.setattr(OPT, "compute_loss", lambda sim, params, *a: (lc(0), sim))
        opt = types.SimpleNamespace(learning_rate=1.0, initial_steps=0, history=None, calls=0)
        opt._engine = types.SimpleNamespace(records=lambda k, n: [types.SimpleNamespace(total_train=losses[k], val=[val0[k]])])

        def step(optimizer, state, simulation, data_targets, loss_functions, indexes):
            k = optimizer.calls
            optimizer.calls += 1
            save = OptimizationState(P(("w", k)), None, k + 1, lc(k), None)
            return OptimizationState(P(("z", k + 1)), None, k + 1, lc(k), None), losses[k], save, simulation

        opt.step = step
        sim = types.SimpleNamespace(params=None)
        return sim, opt

    # (1) converges: thresholds 0.5 / 0.05 (x lr = 1), min 2 steps per threshold
    losses = [10.0, 6.0, 4.0, 3.5, 3.4, 3.39, 3.389, 3.3889, 3.3888, 3.3887]
    val0 = [5.0, 4.0, 3.0, 2.5, 2.6, 2.7, 2.8, 2.9, 3.0, 3.1]
    sim, opt = scripted(losses, val0)
    sim2, opt2 = RUN._optimise_pure(sim, (None,), 10, 1e-10, [0.5, 0.05], (0,), (None,), OptimizationState(P("z0"), None), opt)
    carry, n = opt2._pure_carry
    ref_carry = initialise_and_replay(losses, [0.5, 0.05])
    assert n == ref_carry[1] and carry.converged == ref_carry[0].converged and n < 10
    h = opt2.history
    assert [s.step for s in h.states] == list(range(n))
    assert h.best_state is h.states[3] and sim2.params is h.best_state.params  # val0 is smallest at entry 3
    kept = [i for i, s in enumerate(h.states) if s.params.frame_weights is not None]
    assert kept == sorted({3, n - 1}) and h.states[3].params.frame_weights == ("w", 3)
    # (2) n_steps ends the loop; the last entry ties with the minimum -> the last entry is returned
    sim, opt = scripted([5.0, 4.0, 3.0, 2.0], [2.0, 1.0, 1.5, 1.0])
    _, opt2 = RUN._optimise_pure(sim, (None,), 4, 1e-10, [1e-9], (0,), (None,), OptimizationState(P("z0"), None), opt)
    assert opt2._pure_carry[1] == 4 and opt2.history.best_state is opt2.history.states[3]
    assert [i for i, s in enumerate(opt2.history.states) if s.params.frame_weights is not None] == [1, 3]
    # (3) cond_fn looks at the initial loss before the first iteration: below the tolerance -> no iteration at all
    sim, opt = scripted([1e-12, 1.0], [1.0, 1.0])
    _, opt2 = RUN._optimise_pure(sim, (None,), 5, 1e-10, [1e-3], (0,), (None,), OptimizationState(P("z0"), None), opt)
    assert opt2._pure_carry[1] == 0 and opt2.history.states == [] and opt.calls == 0


def initialise_and_replay(losses, convergence, lr=1.0, alpha=0.5, min_steps=2, initial_steps=0):
    from jaxent_b200.opt import track as T

    thr = T.create_convergence_thresholds(convergence, lr)
    carry, prev, n = T.initialise_convergence_carry(None), losses[0], 0
    while not carry.converged and n < len(losses):
        carry, _ = T.update_convergence(carry, prev, losses[n], None, alpha)
        carry = T.check_and_advance_threshold(carry, losses[n], n + 1, thr, min_steps, initial_steps)
        prev = losses[n]
        n += 1
    return carry, n
