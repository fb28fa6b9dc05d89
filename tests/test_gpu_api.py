"""GPU tests of the reference-API mirror: Simulation / OptaxOptimizer.step / compute_loss / run_optimise, used
the way examples/common/optimization.py:85-261 uses the reference, checked against the oracle."""
import numpy as np
import pytest
import torch

import helpers as H
import jaxent_b200 as J
from oracle import bv_oracle as O

pytestmark = pytest.mark.gpu


class _Model:  # a ForwardModel stand-in: only `.forwardpass` is read on this path (models/core.py:42-44)
    def __init__(self, fp):
        self.forwardpass = fp


def _setup(F=600, R=41, T=4, kind=O.UPTAKE_EYE_MSE, opt_model=False, gamma=2.0, scaling=100.0, lr=0.5):
    case = H.make_case(F, R, T, kind, extra=(O.PARAMS_L2,), seed=13, gamma=gamma, scaling=scaling)
    ens, pb = case["ens"], case["problem"]
    uptake = kind != O.PF_L2
    feats = J.BV_input_features(ens.heavy, ens.acceptor, ens.k_ints)
    mp = J.BV_Model_Parameters(bv_bc=np.array([0.35], np.float32), bv_bh=np.array([2.0], np.float32), timepoints=ens.timepoints)
    yt = lambda y: y[:, :, None] if uptake else y  # the loader's (P,T,1) layout (data/loader.py:219-221)
    loader = J.ExpD_Dataloader(
        train=J.Dataset([], yt(ens.y_train), J.SparseFragmentMapping(ens.S_train)),
        val=J.Dataset([], yt(ens.y_val), J.SparseFragmentMapping(ens.S_val)))
    n = 3
    params = J.Simulation_Parameters(
        frame_weights=np.full(F, 1.0 / F, np.float32), frame_mask=np.ones(F, np.float32), model_parameters=[mp],
        forward_model_weights=np.array([gamma, 1.0, 1.0], np.float32), normalise_loss_functions=np.ones(n, np.float32),
        forward_model_scaling=np.full(n, scaling, np.float32))
    prior_mp = J.BV_Model_Parameters(bv_bc=np.array([0.3], np.float32), bv_bh=np.array([2.2], np.float32), timepoints=ens.timepoints)
    prior = J.Simulation_Parameters(params.frame_weights, params.frame_mask, [prior_mp], params.forward_model_weights,
                                    params.normalise_loss_functions, params.forward_model_scaling)
    fp = J.BV_uptake_ForwardPass() if uptake else J.BV_ForwardPass()
    sim = J.Simulation(input_features=(feats,), forward_models=(_Model(fp),), params=params)
    data = (loader, prior, prior)
    primary = J.hdx_uptake_eye_MSE_loss if uptake else J.hdx_pf_l2_loss
    losses = (primary, J.maxent_convexKL_loss, J.model_params_L2_loss)
    masks = {J.Optimisable_Parameters.frame_weights}
    if opt_model:
        masks.add(J.Optimisable_Parameters.model_parameters)
    opt = J.OptaxOptimizer(learning_rate=lr, parameter_partition_masks=masks, clip_value=None, optimizer="adam",
                           initial_learning_rate=1.0, initial_steps=2)
    cfg = O.OptConfig(learning_rate=lr, initial_learning_rate=1.0, initial_steps=2, clip_value=None,
                      optimise_model_parameters=opt_model)
    return case, sim, data, losses, opt, cfg


def test_simulation_initialise_and_forward():
    case, sim, *_ = _setup()
    assert sim.initialise() is True
    F = 600
    w = sim.params.frame_weights
    assert isinstance(w, torch.Tensor) and w.is_cuda and abs(float(w.sum()) - 1) < 1e-5
    np.testing.assert_allclose(sim.params.forward_model_weights, [0.5, 0.25, 0.25], rtol=1e-6)
    assert isinstance(sim.outputs[0], J.uptake_BV_output_features) and tuple(sim.outputs[0].uptake.shape) == (4, 41)
    assert sim.outputs_by_key["HDX_peptide"] is sim.outputs[0]
    # forward at non-uniform logits vs the oracle
    z = np.random.default_rng(0).normal(size=F).astype(np.float32)
    p2 = J.Simulation_Parameters(z, sim.params.frame_mask, sim.params.model_parameters, sim.params.forward_model_weights,
                                 sim.params.normalise_loss_functions, sim.params.forward_model_scaling)
    new_sim, outs = J.Simulation.forward(sim, p2, mutate=False)
    par = case["params"]
    _, npar, out = O.compute_loss(case["problem"], O.Params(z, par.frame_mask, 0.35, 2.0, par.fw, par.fs, par.nlf), np.float64)
    np.testing.assert_allclose(outs[0].uptake.cpu().numpy(), out["uptake"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(new_sim.params.frame_weights.cpu().numpy(), npar.z, rtol=3e-6)
    assert new_sim is not sim and sim.outputs[0] is not outs[0]


@pytest.mark.parametrize("kind,opt_model", [(O.UPTAKE_EYE_MSE, True), (O.PF_L2, False)])
def test_step_api_tracks_oracle(kind, opt_model):
    case, sim, data, losses, opt, cfg = _setup(kind=kind, T=4 if kind != O.PF_L2 else 1, opt_model=opt_model)
    sim.initialise()
    state = opt.initialise(sim, _jit_test_args=(data, losses, (0, 0, 0)))
    assert float(state.params.frame_weights[0]) == pytest.approx(1.0)  # z0 = w * n_frames (optimiser.py:243)
    par = case["params"]
    par = O.Params(par.z, par.frame_mask, par.bc, par.bh, O.normalize_masked_loss_scalingweights([2.0, 1.0, 1.0], np.ones(3)), par.fs, par.nlf)
    ost = O.optimizer_initialise(par, cfg, np.float64)
    for k in range(6):
        # teacher-forced part: the oracle evaluated at the parameters the API state holds BEFORE the step (north_star: 1e-5)
        z_pre = state.params.frame_weights.cpu().numpy().astype(np.float64)
        mp_pre = state.params.model_parameters[0]
        tf_par = O.Params(z_pre, par.frame_mask, float(mp_pre.bv_bc[0]), float(mp_pre.bv_bh[0]), par.fw, par.fs, par.nlf)
        tf_l, tf_g, _ = O.loss_and_grads(case["problem"], tf_par, np.float64)
        state, loss, save, usim = opt.step(optimizer=opt, state=state, simulation=sim, data_targets=data,
                                           loss_functions=losses, indexes=(0, 0, 0))
        np.testing.assert_allclose(float(loss), tf_l.total_train_loss, rtol=1e-5)
        np.testing.assert_allclose(state.losses.train_losses.cpu().numpy(), tf_l.train_losses, rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(state.losses.val_losses.cpu().numpy(), tf_l.val_losses, rtol=1e-5, atol=1e-9)
        g_tf = state.gradients.frame_weights.cpu().numpy()
        assert float(np.abs(g_tf - tf_g.z).max() / np.abs(tf_g.z).max()) < 1e-5
        ost, ol, ow = O.step_with_rates(case["problem"], ost, cfg, np.float64)
        assert int(state.step) == k + 1 and isinstance(state, J.OptimizationState)
        # step 0 starts from identical states (tight); later steps are free-running, not teacher-forced (loose)
        rt = 2e-5 if k == 0 else 3e-3
        np.testing.assert_allclose(float(loss), ol, rtol=rt)
        np.testing.assert_allclose(state.losses.train_losses.cpu().numpy(), ost.losses.train_losses, rtol=rt, atol=1e-6)
        np.testing.assert_allclose(state.losses.val_losses.cpu().numpy(), ost.losses.val_losses, rtol=rt, atol=1e-6)
        np.testing.assert_allclose(save.params.frame_weights.cpu().numpy(), ow, rtol=1e-3, atol=1e-7)
        g = state.gradients.frame_weights.cpu().numpy()
        # free-running comparison (not teacher-forced): trajectories separate slowly, so 1e-3 of the gradient scale
        gerr = float(np.abs(g - ost.gradients.z).max() / np.abs(ost.gradients.z).max())
        assert gerr <= 1e-3, f"step {k}: free-running gradient error {gerr:.2e}"
        np.testing.assert_allclose(float(state.params.model_parameters[0].bv_bc[0]), ost.params.bc, rtol=1e-4)
        assert abs(float(save.params.frame_weights.sum()) - 1.0) < 1e-5
    with pytest.raises(RuntimeError, match="device-resident optimiser state"):
        opt.step(optimizer=opt, state=state._replace(step=2), simulation=sim, data_targets=data, loss_functions=losses, indexes=(0, 0, 0))


def test_compute_loss_matches_oracle():
    case, sim, data, losses, opt, cfg = _setup()
    sim.initialise()
    z = np.random.default_rng(1).normal(size=600).astype(np.float32)
    p = sim.params
    q = J.Simulation_Parameters(z, p.frame_mask, p.model_parameters, p.forward_model_weights, p.normalise_loss_functions, p.forward_model_scaling)
    lc, new_sim = J.compute_loss(sim, q, data, (0, 0, 0), losses)
    par = case["params"]
    fw = O.normalize_masked_loss_scalingweights([2.0, 1.0, 1.0], np.ones(3))
    ol, _, _ = O.compute_loss(case["problem"], O.Params(z, par.frame_mask, 0.35, 2.0, fw, par.fs, par.nlf), np.float64)
    assert isinstance(lc, J.LossComponents)
    np.testing.assert_allclose(lc.train_losses.cpu().numpy(), ol.train_losses, rtol=1e-5)
    np.testing.assert_allclose(lc.val_losses.cpu().numpy(), ol.val_losses, rtol=1e-5)
    np.testing.assert_allclose(float(lc.total_train_loss), ol.total_train_loss, rtol=1e-5)
    np.testing.assert_allclose(float(lc.total_val_loss), ol.total_val_loss, rtol=1e-5)


def test_run_optimise_loop_semantics():
    case, sim, data, losses, opt, cfg = _setup(lr=0.1, scaling=100.0)
    settings = J.OptimiserSettings(name="t", n_steps=80, tolerance=1e-10, convergence=[1e-2, 1e-3], learning_rate=0.1)
    sim2, history = J.run_optimise(sim, data, settings, forward_models=[], indexes=[0, 0, 0], loss_functions=list(losses),
                                   optimizer=opt, initialise=True, silent=True)
    par = case["params"]
    par = O.Params(par.z, par.frame_mask, par.bc, par.bh, O.normalize_masked_loss_scalingweights([2.0, 1.0, 1.0], np.ones(3)), par.fs, par.nlf)
    res = O.optimise(case["problem"], par, cfg.__class__(**{**cfg.__dict__, "learning_rate": 0.1}), 80, tolerance=1e-10,
                     convergence=[1e-2, 1e-3], dtype=np.float32)
    assert len(history.states) == len(res["history"]) >= 1
    assert [int(s.step) for s in history.states] == [h["step"] for h in res["history"]]
    for s, h in zip(history.states, res["history"]):
        np.testing.assert_allclose(s.params.frame_weights.cpu().numpy(), h["weights"], atol=1e-4)
    assert history.best_state is not None
    assert opt.ema_history is not None and len(opt.ema_history.states) == len(history.states) - 1  # no EMA yet at step 0
    np.testing.assert_allclose(sim2.params.frame_weights.cpu().numpy(), res["best"]["weights"], atol=1e-4)


def test_on_device_loop_matches_host_loop():
    """_optimise_device (tracker / EMA / snapshots / stop flag in the tail kernel) == _optimise (host loop)."""
    hist = {}
    for name, fn in (("host", J._optimise), ("device", J._optimise_device)):
        case, sim, data, losses, opt, cfg = _setup(lr=0.1, scaling=100.0)
        settings = J.OptimiserSettings(name="t", n_steps=120, tolerance=1e-10, convergence=[1e-2, 1e-3, 1e-4], learning_rate=0.1)
        sim2, history = J.run_optimise(sim, data, settings, forward_models=[], indexes=[0, 0, 0], loss_functions=list(losses),
                                       optimizer=opt, initialise=True, silent=True, _opt_fn=fn)
        hist[name] = (history, opt, sim2)
    h, d = hist["host"][0], hist["device"][0]
    assert len(d.states) == len(h.states) >= 2
    assert [int(s.step) for s in d.states] == [int(s.step) for s in h.states]
    for a, b in zip(d.states, h.states):
        assert torch.equal(a.params.frame_weights, b.params.frame_weights)  # same kernels, same order: bitwise
        np.testing.assert_allclose(a.losses.train_losses.cpu().numpy(), b.losses.train_losses.cpu().numpy(), rtol=0, atol=0)
        np.testing.assert_allclose(float(a.losses.total_val_loss), float(b.losses.total_val_loss), rtol=0, atol=0)
    eh, ed = hist["host"][1].ema_history, hist["device"][1].ema_history
    assert len(ed.states) == len(eh.states)
    for a, b in zip(ed.states, eh.states):
        np.testing.assert_allclose(a.params.frame_weights.cpu().numpy(), b.params.frame_weights.cpu().numpy(), rtol=2e-6, atol=1e-9)
        np.testing.assert_allclose(float(a.losses.total_train_loss), float(b.losses.total_train_loss), rtol=1e-5)
    assert torch.equal(hist["device"][2].params.frame_weights, hist["host"][2].params.frame_weights)
    st = hist["device"][1]._device_status
    assert st.n_snapshots == len(d.states)
    # after the stop, further launches are no-ops
    eng = hist["device"][1]._engine
    if st.stop:
        z0 = eng.export_state(("z",))[0]["z"].clone()
        eng.step(3)
        torch.cuda.synchronize()
        assert torch.equal(eng.export_state(("z",))[0]["z"], z0) and eng.track_status().steps_done == st.steps_done


@pytest.mark.parametrize("engine", ["gemm", "sequential"])
def test_batch_optimise_follows_the_pure_loop(engine):
    """batch_optimise (reference opt/batch.py:51-192) vmaps _optimise_pure: every step is a history entry, the functional tracker
    decides where each run ends, best state = smallest val_losses[0].  Both engines against oracle.optimise_pure (pinned on the
    reference-executed fixtures), at the tolerances of the reference's own batch test (modules/optimise/
    test_module_batch_optimise.py:163-229: loss rtol 1e-4, frame weights atol 1e-4)."""
    case, sim, data, losses, opt, cfg = _setup(lr=0.1, scaling=100.0)
    sim.initialise()
    conv = [1.0, 0.3]  # the three runs end after 59 / 14 / 60 iterations with best entries 30 / 0 / 22 (oracle, fp32 and fp64 agree)
    settings = J.OptimiserSettings(name="b", n_steps=60, tolerance=1e-10, convergence=conv, learning_rate=0.1)
    fw = np.array([[0.5, 0.25, 0.25], [0.8, 0.1, 0.1], [0.2, 0.6, 0.2]], np.float32)
    fs = np.full((3, 3), 100.0, np.float32)
    lr = np.array([0.1, 0.05, 0.2], np.float32)
    res = J.batch_optimise(sim, J.HParamBatch(fw, fs, lr), batch_size=2, data_to_fit=data, config=settings, indexes=[0, 0, 0],
                           loss_functions=list(losses), optimizer=opt, engine=engine)
    assert len(res.histories) == 3 and res.convergence_steps.shape == (3,)
    p = case["params"]
    for i in range(3):
        par = O.Params(p.z, p.frame_mask, 0.35, 2.0, fw[i], fs[i], p.nlf)
        ref = O.optimise_pure(case["problem"], par, cfg, 60, tolerance=1e-10, convergence=conv, ema_alpha=settings.ema_alpha,
                              min_steps_per_threshold=settings.min_steps_per_threshold, dtype=np.float64, learning_rate=float(lr[i]))
        n = ref["write_idx"]
        hist = res.histories[i]
        assert int(res.convergence_steps[i]) == n == len(hist.states)
        assert [int(s.step) for s in hist.states] == list(range(n))
        np.testing.assert_allclose([float(s.losses.total_train_loss) for s in hist.states], ref["history_total_train"][:n], rtol=1e-4)
        np.testing.assert_allclose([float(s.losses.val_losses[0]) for s in hist.states], ref["history_val0"][:n], rtol=1e-4)
        best = res.best_states[i]
        assert int(best.step) == ref["best_index"] and hist.best_state is best
        np.testing.assert_allclose(best.params.frame_weights.cpu().numpy(), ref["history_w"][ref["best_index"]], atol=1e-4)
        np.testing.assert_allclose(hist.states[-1].params.frame_weights.cpu().numpy(), ref["history_w"][n - 1], atol=1e-4)
        np.testing.assert_allclose(float(best.losses.total_train_loss), ref["history_total_train"][ref["best_index"]], rtol=1e-4)
        kept = [k for k, s in enumerate(hist.states) if s.params.frame_weights is not None]
        assert sorted(set(kept)) == sorted({ref["best_index"], n - 1})  # weights for the best and the last entry only
    with pytest.raises(ValueError, match="batch_size"):
        J.batch_optimise(sim, J.HParamBatch(fw, fs), 0, data, settings, [0, 0, 0], list(losses))
    with pytest.raises(ValueError, match="same n_hparams"):
        J.batch_optimise(sim, J.HParamBatch(fw, fs[:2]), 2, data, settings, [0, 0, 0], list(losses))


@pytest.mark.parametrize("kind", [O.UPTAKE_EYE_MSE, O.PF_L2])
def test_simulation_predict_keeps_the_frame_axis(kind):
    """Simulation.predict (reference models/core.py:181-208): forward pass on the un-averaged features."""
    case, sim, *_ = _setup(F=333, R=41, T=4 if kind != O.PF_L2 else 1, kind=kind)
    with pytest.raises(RuntimeError, match="initiali"):
        sim.predict(sim.params)
    sim.initialise()
    pb = case["problem"]
    out = sim.predict(sim.params)[0]
    lp = O.bv_log_pf(pb.heavy, pb.acceptor, 0.35, 2.0, np.float64)
    if kind == O.PF_L2:
        assert tuple(out.log_Pf.shape) == (41, 333)
        np.testing.assert_allclose(out.log_Pf.cpu().numpy(), lp, rtol=1e-6)
    else:
        assert tuple(out.uptake.shape) == (4, 41, 333)  # reference unit/models/test_bv_forward_pass.py:179-196 shape (T,R,F)
        np.testing.assert_allclose(out.uptake.cpu().numpy(), O.bv_uptake(lp, pb.k_ints, pb.timepoints, np.float64), rtol=2e-5, atol=1e-7)
    out2 = sim.predict(sim.params.model_parameters)[0]  # a bare list of model parameters is accepted as well
    a, b = (out.log_Pf, out2.log_Pf) if kind == O.PF_L2 else (out.uptake, out2.uptake)
    assert torch.equal(a, b)
    with pytest.raises(ValueError, match="must match"):
        sim.predict([])


def test_per_frame_forward_pass_through_the_api():
    """A BV_uptake_ForwardPass with average_first=False (BV_uptake_ForwardPass_frames of the examples) runs the per-frame
    kernel: Simulation.forward == frame average of Simulation.predict, and run_optimise lowers the loss."""
    case, sim, data, losses, opt, cfg = _setup(F=500, R=41, T=4)
    fp = J.BV_uptake_ForwardPass()
    fp.average_first = False
    sim = J.Simulation(input_features=sim.input_features, forward_models=(_Model(fp),), params=sim.params)
    assert sim.initialise() is True
    w = sim.params.frame_weights
    per_frame = sim.predict(sim.params)[0].uptake  # (T,R,F)
    np.testing.assert_allclose(sim.outputs[0].uptake.cpu().numpy(), (per_frame * w[None, None, :]).sum(-1).cpu().numpy(), rtol=2e-5, atol=1e-7)
    opt = J.OptaxOptimizer(learning_rate=0.02, parameter_partition_masks={J.Optimisable_Parameters.frame_weights}, clip_value=None,
                           initial_learning_rate=0.02, initial_steps=0)
    state = opt.initialise(sim)
    n_steps = 25
    sim2, opt2 = J._optimise_device(sim, data, n_steps, 1e-10, [1e-9], [0, 0, 0], list(losses), state, opt)
    recs = opt2._engine.records(0, n_steps + 1)
    assert recs[n_steps].step == n_steps
    # the same trajectory on the oracle (per-frame mode), started from the parameters the Simulation was initialised to
    pb = case["problem"]
    pb.average_first = False
    ocfg = O.OptConfig(learning_rate=0.02, initial_learning_rate=0.02, initial_steps=0, clip_value=None, optimise_model_parameters=False)
    par = case["params"]
    par = O.Params(par.z, par.frame_mask, par.bc, par.bh, O.normalize_masked_loss_scalingweights([2.0, 1.0, 1.0], np.ones(3)), par.fs, par.nlf)
    st = O.optimizer_initialise(par, ocfg, np.float64)
    for k in range(n_steps + 1):
        l, _, _ = O.compute_loss(pb, st.params, np.float64)
        np.testing.assert_allclose(recs[k].total_train, l.total_train_loss, rtol=2e-5)
        st, _, _ = O.step_with_rates(pb, st, ocfg, np.float64)
    # model-parameter optimisation in this mode (reference models/core.py:298-305 has no restriction): same oracle comparison
    opt_m = J.OptaxOptimizer(learning_rate=0.02, parameter_partition_masks={J.Optimisable_Parameters.frame_weights, J.Optimisable_Parameters.model_parameters},
                             clip_value=None, initial_learning_rate=0.02, initial_steps=0)
    _, sim_b, *_ = _setup(F=500, R=41, T=4)  # the first run left its best parameters in `sim`: start again from the initial ones
    sim_m = J.Simulation(input_features=sim_b.input_features, forward_models=(_Model(fp),), params=sim_b.params)
    assert sim_m.initialise() is True
    st_m = opt_m.initialise(sim_m)
    sim3, opt3 = J._optimise_device(sim_m, data, n_steps, 1e-10, [1e-9], [0, 0, 0], list(losses), st_m, opt_m)
    recs = opt3._engine.records(0, n_steps + 1)
    ocfg_m = O.OptConfig(learning_rate=0.02, initial_learning_rate=0.02, initial_steps=0, clip_value=None, optimise_model_parameters=True)
    st = O.optimizer_initialise(par, ocfg_m, np.float64)
    for k in range(n_steps + 1):
        l, _, _ = O.compute_loss(pb, st.params, np.float64)
        np.testing.assert_allclose(recs[k].total_train, l.total_train_loss, rtol=1e-4)  # free-running, BV parameters moving: observed 3e-5
        np.testing.assert_allclose([recs[k].bc, recs[k].bh], [st.params.bc, st.params.bh], rtol=1e-4)
        st, _, _ = O.step_with_rates(pb, st, ocfg_m, np.float64)
    assert abs(recs[n_steps].bc - recs[0].bc) > 1e-3
