"""CPU tests that pin the oracle: the reference's own known-answer tests (tests/golden/
reference_known_answers.json), an independent torch.autograd fp64 gradient check, finite differences
in the style of the reference's unit/gradients/test_gradient_flow.py, and the optimiser rules."""
import json
import os

import numpy as np
import pytest
import torch

import helpers as H
from oracle import bv_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
KA = json.load(open(os.path.join(GOLD, "reference_known_answers.json")))


@pytest.mark.parametrize("case", KA["bv_log_pf"], ids=lambda c: c["src"].split(":")[-1])
def test_bv_log_pf_known_answers(case):
    out = O.bv_log_pf(np.array(case["heavy"]), np.array(case["acceptor"]), case["bc"], case["bh"], np.float32)
    np.testing.assert_allclose(out, np.array(case["expected"], np.float32), rtol=case["rtol"], atol=1e-7)


def test_bv_uptake_known_answers():
    c0, c1, c2, c3 = KA["bv_uptake"]
    up = lambda c: O.bv_uptake(O.bv_log_pf(np.array(c["heavy"]), np.array(c["acceptor"]), c["bc"], c["bh"]),
                               np.array(c["k_ints"]), np.array(c["timepoints"]), np.float32)
    np.testing.assert_allclose(up(c0)[0, 0], c0["expected_00"], rtol=c0["rtol"])
    assert up(c1)[0, 0] < c1["upper_bound_00"]
    assert list(up(c2).shape) == c2["expected_shape"]  # per-frame input -> (T,R,F)
    u = up(c3)
    assert np.all(np.diff(u[:, :, 0], axis=0) >= -1e-6)


def test_loss_known_answers():
    l2, pf = KA["losses"]
    # L2 = 5/3 through an identity peptide map (unit/opt/test_loss_math.py:47-58, test_generalised_loss.py:240-274)
    slot = O.LossSlot(O.PF_L2, np.eye(3), np.array(l2["target"]), np.eye(3), np.array(l2["target"]))
    tr, va = O.slot_losses(None, slot, None, {"log_Pf": np.array(l2["pred"], np.float32)}, np.float32)
    np.testing.assert_allclose(tr, l2["expected"], rtol=l2["rtol"])
    slot = O.LossSlot(O.PF_L2, np.eye(3), np.array(pf["target"]), np.eye(3), np.array(pf["target"]))
    tr, _ = O.slot_losses(None, slot, None, {"log_Pf": np.array(pf["log_pf"], np.float32)}, np.float32)
    np.testing.assert_allclose(tr, pf["expected"], atol=pf["atol"])


def test_loss_weight_normalisation():
    # interfaces/simulation.py:45-76: masked slots renormalised to sum 1, unmasked pass through
    fw = O.normalize_masked_loss_scalingweights(np.array([2.0, 1.0, 5.0]), np.array([1.0, 1.0, 0.0]))
    np.testing.assert_allclose(fw, [2 / 3, 1 / 3, 5.0], rtol=1e-6)


def test_softmax_and_frame_average_shapes():
    # modules/model/test_core_real_inputs_random_data.py:79-118: forward output (R,), deterministic
    rng = np.random.default_rng(42)
    N = rng.random((20, 100)).astype(np.float32)
    w = O.softmax(rng.normal(size=100))
    a1, a2 = O.frame_average(N, w), O.frame_average(N, w)
    assert a1.shape == (20,) and np.array_equal(a1, a2)
    np.testing.assert_allclose(a1, (N * w[None, :]).sum(-1), rtol=2e-6)
    assert O.frame_average(np.arange(5.0), w).shape == (5,)  # 1-D leaves (k_ints) untouched


def _torch_loss(case, z, bc, bh):
    pb, par = case["problem"], case["params"]
    f64 = torch.float64
    Nc, Nh = torch.tensor(pb.heavy, dtype=f64), torch.tensor(pb.acceptor, dtype=f64)
    w = torch.softmax(z, 0)
    a, b = (Nc * w[None, :]).sum(-1), (Nh * w[None, :]).sum(-1)
    lp = bc * a + bh * b
    total = 0.0
    for i, s in enumerate(pb.slots):
        sc = float(par.fw[i]) * float(par.fs[i])
        if s.kind in O.UPTAKE_KINDS:
            k, tp = torch.tensor(pb.k_ints, dtype=f64), torch.tensor(pb.timepoints, dtype=f64)
            if pb.average_first:
                u = 1 - torch.exp(-k[None, :] * tp[:, None] / torch.exp(lp)[None, :])
            else:  # per-frame forward pass, then the frame average of the outputs (reference models/core.py:302-304)
                lpf = bc * Nc + bh * Nh
                uf = 1 - torch.exp(-k[None, :, None] * tp[:, None, None] / torch.exp(lpf)[None, :, :])
                u = (uf * w[None, None, :]).sum(-1)
            S, y = torch.tensor(s.S_train, dtype=f64), torch.tensor(s.y_train, dtype=f64)
            W = torch.tensor(s.W_train, dtype=f64) if s.kind == O.UPTAKE_SIGMA_MSE else torch.eye(y.shape[0], dtype=f64)
            tot = 0.0
            for t in range(y.shape[1]):
                tr, pr = y[:, t], S @ u[t]
                if s.kind == O.UPTAKE_MC_EYE_MSE:
                    tr, pr = tr - tr.mean(), pr - pr.mean()
                r = pr - tr
                tot = tot + 0.5 * torch.dot(r, W @ r)
            li = tot / (y.shape[1] * y.shape[0])
        elif s.kind == O.PF_L2:
            S, y = torch.tensor(s.S_train, dtype=f64), torch.tensor(s.y_train, dtype=f64)
            li = ((S @ lp - y) ** 2).mean()
        elif s.kind == O.MAXENT_CONVEX_KL:
            F = z.shape[0]
            eps = 1e-10 / F
            sw = w.abs() + eps
            sw = sw / sw.sum()
            pw = torch.tensor(s.prior, dtype=f64).abs() + eps
            pw = pw / pw.sum()
            li = (pw * (pw.log() - sw.log())).sum() + (sw - pw).sum()
        elif s.kind == O.PARAMS_L1:
            li = (bc - s.prior_bc).abs() + (bh - s.prior_bh).abs()
        else:
            li = (bc - s.prior_bc) ** 2 + (bh - s.prior_bh) ** 2
        total = total + sc * li
    return total


@pytest.mark.parametrize("kind", [O.UPTAKE_EYE_MSE, O.UPTAKE_MC_EYE_MSE, O.UPTAKE_SIGMA_MSE, O.PF_L2])
def test_closed_form_gradients_match_autograd(kind):
    case = H.make_case(200, 17, 4, kind, extra=(O.PARAMS_L2, O.PARAMS_L1), seed=5, n_pep=9, prior="random", gamma=3.0, scaling=50.0)
    rng = np.random.default_rng(1)
    z = rng.normal(size=200) * 2
    par = O.Params(z, np.ones(200), 0.4, 1.9, case["params"].fw, case["params"].fs, case["params"].nlf)
    L, g, _ = O.loss_and_grads(case["problem"], par, np.float64)
    tz = torch.tensor(z, dtype=torch.float64, requires_grad=True)
    tbc = torch.tensor(0.4, dtype=torch.float64, requires_grad=True)
    tbh = torch.tensor(1.9, dtype=torch.float64, requires_grad=True)
    tot = _torch_loss(case, tz, tbc, tbh)
    tot.backward()
    np.testing.assert_allclose(float(tot.detach()), float(L.total_train_loss), rtol=1e-12)
    np.testing.assert_allclose(g.z, tz.grad.numpy(), rtol=0, atol=1e-12 * np.abs(g.z).max())
    np.testing.assert_allclose([g.bc, g.bh], [float(tbc.grad), float(tbh.grad)], rtol=1e-10)


@pytest.mark.parametrize("kind", [O.UPTAKE_EYE_MSE, O.UPTAKE_MC_EYE_MSE, O.UPTAKE_SIGMA_MSE])
def test_per_frame_uptake_gradients_match_autograd(kind):
    """average_first=False: closed-form backward of the per-frame uptake mode vs torch.autograd fp64"""
    case = H.make_case(150, 13, 4, kind, extra=(O.PARAMS_L2,), seed=8, n_pep=7, prior="random", gamma=2.0, scaling=30.0)
    case["problem"].average_first = False
    rng = np.random.default_rng(2)
    z = rng.normal(size=150) * 1.5
    par = O.Params(z, np.ones(150), 0.1, 0.6, case["params"].fw, case["params"].fs, case["params"].nlf)
    L, g, aux = O.loss_and_grads(case["problem"], par, np.float64)
    tz = torch.tensor(z, dtype=torch.float64, requires_grad=True)
    tbc = torch.tensor(0.1, dtype=torch.float64, requires_grad=True)
    tbh = torch.tensor(0.6, dtype=torch.float64, requires_grad=True)
    tot = _torch_loss(case, tz, tbc, tbh)
    tot.backward()
    np.testing.assert_allclose(float(tot.detach()), float(L.total_train_loss), rtol=1e-12)
    np.testing.assert_allclose(g.z, tz.grad.numpy(), rtol=0, atol=1e-11 * np.abs(g.z).max())
    np.testing.assert_allclose([g.bc, g.bh], [float(tbc.grad), float(tbh.grad)], rtol=1e-9)
    # the two modes differ (the average of a non-linear function is not the function of the average)
    case["problem"].average_first = True
    L2, _, _ = O.loss_and_grads(case["problem"], par, np.float64)
    assert abs(float(L2.total_train_loss) - float(L.total_train_loss)) > 1e-6 * abs(float(L.total_train_loss))


def test_finite_difference_bv_parameters():
    # reference style: unit/gradients/test_gradient_flow.py:103-140,181-223 (h=1e-5, rtol 2e-3)
    for kind in (O.PF_L2, O.UPTAKE_EYE_MSE):
        case = H.make_case(64, 11, 3, kind, seed=3, n_pep=6)
        par = case["params"]
        par = O.Params(np.zeros(64), par.frame_mask, 0.35, 2.0, par.fw, par.fs, par.nlf)
        _, g, _ = O.loss_and_grads(case["problem"], par, np.float64)
        h = 1e-5
        for name, gv in (("bc", g.bc), ("bh", g.bh)):
            from dataclasses import replace

            lp, _, _ = O.compute_loss(case["problem"], replace(par, **{name: getattr(par, name) + h}), np.float64)
            lm, _, _ = O.compute_loss(case["problem"], replace(par, **{name: getattr(par, name) - h}), np.float64)
            fd = (lp.total_train_loss - lm.total_train_loss) / (2 * h)
            np.testing.assert_allclose(gv, fd, rtol=2e-3)


def test_fp32_mirror_tracks_fp64():
    case = H.make_case(500, 53, 3, O.UPTAKE_EYE_MSE, seed=0)
    cfg = O.OptConfig(learning_rate=1.0, clip_value=None)
    s32 = O.optimizer_initialise(case["params"], cfg, np.float32)
    s64 = O.optimizer_initialise(case["params"], cfg, np.float64)
    for _ in range(3):
        s32, l32, _ = O.step_with_rates(case["problem"], s32, cfg, np.float32)
        s64, l64, _ = O.step_with_rates(case["problem"], s64, cfg, np.float64)
        assert abs(l32 - l64) <= 1e-4 * abs(l64)


def test_adam_first_step_is_unit_sign_step():
    # optax adam, t=1: u = -g/(|g|+eps)  -> a step of ~lr-independent unit size
    cfg = O.OptConfig()
    g = np.array([1e-3, -2.0, 0.0], np.float32)
    u, m, v, c = O._adam_update(g, np.zeros(3, np.float32), np.zeros(3, np.float32), 0, cfg, np.float32)
    np.testing.assert_allclose(u, [-1.0, 1.0, 0.0], atol=2e-5)
    assert c == 1


def test_plateau_rule_and_initial_steps():
    # opt/optimiser.py:365-381,403-413: LR only reduced when dot<0 and step>1; frame-only mask before initial_steps
    case = H.make_case(128, 9, 3, O.UPTAKE_EYE_MSE, seed=2, n_pep=5)
    cfg = O.OptConfig(learning_rate=0.5, initial_learning_rate=1.0, initial_steps=2, clip_value=None, optimise_model_parameters=True)
    st = O.optimizer_initialise(case["params"], cfg, np.float64)
    lrs, bcs = [], []
    for k in range(8):
        st, _, _ = O.step_with_rates(case["problem"], st, cfg, np.float64)
        lrs.append(st.new_lr)
        bcs.append(st.params.bc)
        base = 1.0 if k < 2 else 0.5
        if k < 2:
            assert bcs[-1] == 0.35  # model parameters masked during the initial steps
        expect = base / 1.005 if (st.grad_dot < 0 and k > 1) else base
        if k < 2:
            expect = lrs[-1]  # carried LR during initial steps
        np.testing.assert_allclose(st.new_lr, expect, rtol=1e-6)
    assert bcs[-1] != 0.35


def test_keep_params_nonnegative():
    case = H.make_case(64, 7, 3, O.UPTAKE_EYE_MSE, seed=4, n_pep=4, bc=1e-4, bh=1e-4)
    cfg = O.OptConfig(learning_rate=1.0, clip_value=None, optimise_model_parameters=True)
    st = O.optimizer_initialise(case["params"], cfg, np.float64)
    for _ in range(5):
        st, _, _ = O.step_with_rates(case["problem"], st, cfg, np.float64)
        assert st.params.bc >= 0 and st.params.bh >= 0


def test_bpti_fixture_optimise_runs_and_recovers():
    # BASELINE config 1 shape: real BPTI features (53 x 500), synthetic targets; loss must fall (reference asserts
    # only "finite / decreased": modules/optimise/test_module_optimise_convergence.py:186-191)
    d = np.load(os.path.join(GOLD, "bpti_bv_features.npz"))
    heavy, acc = d["heavy"].astype(np.float32), d["acceptor"].astype(np.float32)
    assert heavy.shape == (53, 500)
    rng = np.random.default_rng(0)
    resids = d["resids"]
    S = np.zeros((len(d["segs"]), 53), np.float32)
    for p, (a, b) in enumerate(d["segs"]):
        m = (resids >= a) & (resids <= b)
        if m.any():
            S[p, m] = 1.0 / m.sum()
    S = S[S.sum(1) > 0]
    k = (10 ** rng.uniform(-2, 2, 53)).astype(np.float32)
    tp = np.array([0.167, 1.0, 10.0], np.float32)
    wstar = rng.dirichlet(np.full(500, 0.3))
    lp = 0.35 * heavy @ wstar + 2.0 * acc @ wstar
    y = (S @ (1 - np.exp(-k[None, :] * tp[:, None] * np.exp(-lp)[None, :])).T).astype(np.float32)
    n = S.shape[0]
    slots = [O.LossSlot(O.UPTAKE_EYE_MSE, S[: n - 5], y[: n - 5], S[n - 5 :], y[n - 5 :]),
             O.LossSlot(O.MAXENT_CONVEX_KL, prior=np.full(500, 1 / 500, np.float32))]
    pb = O.Problem(heavy, acc, k, tp, slots, True)
    fw = O.normalize_masked_loss_scalingweights(np.array([1.0, 1.0]), np.ones(2))
    par = O.Params(np.full(500, 1 / 500, np.float32), np.ones(500, np.float32), 0.35, 2.0, fw, np.full(2, 100.0, np.float32), np.ones(2))
    res = O.optimise(pb, par, O.OptConfig(learning_rate=0.1, clip_value=None), 60, convergence=[1e-3, 1e-4])
    tr = res["loss_trace"]
    # Adam with unit LR on pre-scaled gradients takes a +-1 sign step first, so the loss jumps at
    # step 1 and then has to come back down
    assert np.all(np.isfinite(tr)) and min(tr[20:]) < tr[1]
    assert len(res["history"]) >= 1 and res["best"] is not None
    np.testing.assert_allclose(res["history"][-1]["weights"].sum(), 1.0, rtol=1e-5)


@pytest.mark.parametrize("clip", [None, 0.05])
def test_trajectory_matches_torch_optim_adam(clip):
    """The optax pieces are not importable here (oracle header: parity unpinned).  This pins the oracle's update rule against an
    INDEPENDENT implementation of the same published algorithm: torch.optim.Adam(lr=1) fed the LR-pre-scaled, clipped
    gradients of a torch.autograd fp64 loss (opt/optimiser.py:124-148,403-424), with the plateau rule and
    keep_params_nonnegative applied by hand."""
    case = H.make_case(160, 11, 4, O.UPTAKE_EYE_MSE, extra=(O.PARAMS_L2,), seed=12, n_pep=6, prior="random", gamma=2.0, scaling=40.0)
    cfg = O.OptConfig(learning_rate=0.3, initial_learning_rate=0.3, clip_value=clip, optimise_model_parameters=True, model_parameters_lr_scale=0.5)
    st = O.optimizer_initialise(case["params"], cfg, np.float64)
    tz = torch.tensor(st.params.z, dtype=torch.float64, requires_grad=True)
    tb = torch.tensor([st.params.bc, st.params.bh], dtype=torch.float64, requires_grad=True)
    opt = torch.optim.Adam([tz, tb], lr=1.0, betas=(cfg.b1, cfg.b2), eps=cfg.eps)
    prev = None
    reduced = 0
    for k in range(25):
        st, loss, _ = O.step_with_rates(case["problem"], st, cfg, np.float64)
        opt.zero_grad()
        tot = _torch_loss(case, tz, tb[0], tb[1])
        tot.backward()
        np.testing.assert_allclose(float(tot.detach()), float(loss), rtol=1e-10)
        g = [tz.grad.clone(), tb.grad.clone()]
        dot = float(sum((a * b).sum() for a, b in zip(prev if prev is not None else g, g)))
        prev = g
        red = dot < 0 and k > 1
        reduced += red
        lr = cfg.learning_rate / (cfg.plateau_denominator if red else 1.0)
        tz.grad.mul_(lr)
        tb.grad.mul_(lr * cfg.model_parameters_lr_scale)
        if clip is not None:
            tz.grad.clamp_(-clip, clip)
            tb.grad.clamp_(-clip, clip)
        opt.step()
        with torch.no_grad():
            tb.clamp_(min=0.0)  # where(p + u < 0, -p, u)
        np.testing.assert_allclose(st.params.z, tz.detach().numpy(), rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose([st.params.bc, st.params.bh], tb.detach().numpy(), rtol=1e-9)
    assert reduced > 0  # the plateau branch was exercised


def test_convex_kl_matches_torch_kl_div():
    """maxent_convexKL_loss (opt/losses.py:304-322) = KL(p || q) + sum(q - p): against torch.nn.functional.kl_div"""
    rng = np.random.default_rng(3)
    F = 500
    z = rng.normal(size=F) * 3
    prior = rng.random(F) + 0.01
    prior /= prior.sum()
    w = O.softmax(z, np.float64)
    eps = 1e-10 / F
    q = torch.tensor((np.abs(w) + eps) / (np.abs(w) + eps).sum())
    p = torch.tensor((np.abs(prior) + eps) / (np.abs(prior) + eps).sum())
    want = float(torch.nn.functional.kl_div(q.log(), p, reduction="sum") + (q - p).sum())
    case = H.make_case(F, 5, 3, O.UPTAKE_EYE_MSE, seed=1, n_pep=3)
    slot = O.LossSlot(O.MAXENT_CONVEX_KL, prior=prior)
    pb = O.Problem(case["problem"].heavy, case["problem"].acceptor, case["problem"].k_ints, case["problem"].timepoints, [slot], True)
    par = O.Params(z, np.ones(F), 0.35, 2.0, np.ones(1), np.ones(1), np.zeros(1))
    L, _, _ = O.loss_and_grads(pb, par, np.float64)
    np.testing.assert_allclose(float(L.train_losses[0]), want, rtol=1e-10)
