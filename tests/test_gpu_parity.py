"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded
inputs, against the committed golden fixtures, and -- at BASELINE config-3 size -- against a plain
torch fp64 autograd restatement of the same op.

Tolerances (BASELINE.json north_star): fp32 loss and gradients within 1e-5 relative per step
(gradients: max-abs error relative to the max-abs gradient); frame weights within 1e-4 (absolute)
after 1000 steps.  Per-step checks are teacher-forced: the oracle is evaluated at the state the GPU
exported, so trajectory divergence of the lr=1 Adam dynamics does not leak into the per-step bound.
"""
import os

import numpy as np
import pytest
import torch

import helpers as H
from oracle import bv_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")

STEP_RTOL = 1e-5


def _sync():
    torch.cuda.synchronize()


def _oracle_params(case, z, sc):
    p = case["params"]
    return O.Params(np.asarray(z, np.float64), p.frame_mask, sc["bc"], sc["bh"], p.fw, p.fs, p.nlf)


def _check_steps(case, cfg, n_steps, check_adam=True, STEP_RTOL=STEP_RTOL, mirror_slack=False):
    eng = H.make_engine(case, cfg)
    H.init_engine(eng, case)
    pb = case["problem"]
    worst_loss = worst_g = 0.0
    for k in range(n_steps):
        pre, sc = eng.export_state(("z", "m", "v"))
        z, m, v = (pre[x].cpu().numpy() for x in ("z", "m", "v"))
        eng.step(1)
        _sync()
        rec = eng.records(k, 1)[0]
        post, sc2 = eng.export_state(("z", "g"))
        assert rec.step == k and rec.complete == 1 and sc2["step"] == k + 1
        losses, g, aux = O.loss_and_grads(pb, _oracle_params(case, z, sc), np.float64)
        # loss components and total
        np.testing.assert_allclose(rec.total_train, losses.total_train_loss, rtol=STEP_RTOL)
        np.testing.assert_allclose(rec.total_val, losses.total_val_loss, rtol=STEP_RTOL)
        n = len(pb.slots)
        np.testing.assert_allclose(np.array(rec.train[:n]), losses.train_losses, rtol=STEP_RTOL, atol=1e-9)
        np.testing.assert_allclose(np.array(rec.val[:n]), losses.val_losses, rtol=STEP_RTOL, atol=1e-9)
        # masked gradients
        mask_model = 1.0 if (cfg.optimise_model_parameters and k >= cfg.initial_steps) else 0.0
        mask_frame = 1.0 if (cfg.optimise_frame_weights or k < cfg.initial_steps) else 0.0
        gg = post["g"].cpu().numpy()
        scale = max(np.abs(g.z).max(), 1e-30)
        gerr = np.abs(gg - g.z * mask_frame).max() / scale
        tol = STEP_RTOL
        if mirror_slack:
            # ill-conditioned corner (one frame carries ~all the weight, g -> 0 against h = O(lambda)): where the fp32
            # arithmetic of the reference itself is more than 3e-5 from the fp64 truth, require at most a third of THAT
            _, g32, _ = O.loss_and_grads(pb, _oracle_params(case, z, sc), np.float32)
            tol = max(tol, np.abs(g32.z - g.z).max() / scale / 3.0)
        assert gerr < tol, f"step {k}: grad error {gerr:.2e} (tolerance {tol:.1e})"
        np.testing.assert_allclose(rec.grad_bc, g.bc * mask_model, rtol=2 * STEP_RTOL, atol=1e-6 * abs(g.bc))
        np.testing.assert_allclose(rec.grad_bh, g.bh * mask_model, rtol=2 * STEP_RTOL, atol=1e-6 * abs(g.bh))
        worst_loss = max(worst_loss, abs(rec.total_train - losses.total_train_loss) / abs(losses.total_train_loss))
        worst_g = max(worst_g, gerr)
        if check_adam:
            # the frame Adam update, restated from the exported moments in fp32 (optax formula)
            nxt = eng.records(k + 1, 1)[0]  # carries lr / branch of the update that led to step k+1
            sg = (gg * np.float32(nxt.lr)).astype(np.float32)
            if cfg.clip_value is not None:
                sg = np.clip(sg, -np.float32(cfg.clip_value), np.float32(cfg.clip_value))
            u, _, _, _ = O._adam_update(sg, m, v, sc["count_frame"], cfg, np.float32)
            np.testing.assert_allclose(post["z"].cpu().numpy(), z + u, rtol=0, atol=2e-6 * max(1.0, np.abs(z).max()))
            assert nxt.branch == int(nxt.grad_dot < 0 and k > 1)
    return eng, worst_loss, worst_g


@pytest.mark.parametrize(
    "kind,extra,opt_model,clip",
    [
        (O.UPTAKE_EYE_MSE, (), False, None),
        (O.UPTAKE_MC_EYE_MSE, (O.PARAMS_L1,), True, 1.0),
        (O.UPTAKE_SIGMA_MSE, (), False, None),
        (O.PF_L2, (O.PARAMS_L2,), True, None),
    ],
)
def test_step_parity_small(kind, extra, opt_model, clip):
    case = H.make_case(777, 97, 5 if kind != O.PF_L2 else 1, kind, extra=extra, seed=11, gamma=2.0)
    cfg = O.OptConfig(learning_rate=1.0, initial_learning_rate=1.0, initial_steps=0, clip_value=clip,
                      optimise_model_parameters=opt_model)
    _check_steps(case, cfg, 8)


def test_step_parity_initial_steps_and_random_prior():
    case = H.make_case(1000, 130, 1, O.PF_L2, extra=(O.PARAMS_L2,), seed=3, prior="random")
    cfg = O.OptConfig(learning_rate=0.1, initial_learning_rate=1.0, initial_steps=3, clip_value=None, optimise_model_parameters=True)
    eng, _, _ = _check_steps(case, cfg, 7)
    recs = eng.records(0, 7)
    assert all(r.grad_bc == 0.0 for r in recs[:3]) and recs[4].grad_bc != 0.0  # model params masked during initial steps


def _collapsed_case(F, R, T, kind, spread, seed=13, **kw):
    """logits ~ N(0, spread^2): most frames fall below eps = 1e-10/F, where the MaxEnt part of hbar = sum_j w_j kappa_j
    is O(lambda) instead of O(1e-10 lambda) (reference opt/losses.py:304-322 differentiated exactly by jax.value_and_grad,
    opt/optimiser.py:391-393)."""
    rng = np.random.default_rng(seed + 4)
    z0 = rng.normal(0.0, spread, F)
    case = H.make_case(F, R, T, kind, seed=seed, **kw)
    case["params"].z = (z0 / F).astype(np.float32)  # OptaxOptimizer.initialise multiplies by F (opt/optimiser.py:243)
    w = O.softmax(z0, np.float64)
    return case, float((w < 1e-10 / F).mean())


@pytest.mark.parametrize("spread,min_collapsed", [(10.0, 0.5), (30.0, 0.8)])
@pytest.mark.parametrize("kind,opt_model", [(O.UPTAKE_EYE_MSE, False), (O.PF_L2, True)])
def test_step_parity_collapsed_weights(spread, min_collapsed, kind, opt_model):
    """the exact softmax Jacobian (hbar includes the MaxEnt sum) at north_star's 1e-5."""
    case, frac = _collapsed_case(2000, 40, 4 if kind != O.PF_L2 else 1, kind, spread, extra=(O.PARAMS_L2,) if opt_model else ())
    assert frac >= min_collapsed, frac
    cfg = O.OptConfig(learning_rate=1.0, clip_value=None, optimise_model_parameters=opt_model)
    # the oracle's own MaxEnt part of hbar is O(lambda) here, i.e. the case exercises the term
    st = O.optimizer_initialise(case["params"], cfg, np.float64)
    _, _, aux = O.loss_and_grads(case["problem"], st.params, np.float64)
    assert abs(aux["hbar_kl"]) > 0.3 * float(st.params.fw[1] * st.params.fs[1])
    _check_steps(case, cfg, 8, mirror_slack=spread > 20)


def test_step_parity_collapsed_weights_config3_rows():
    """same regime at R = 500 rows, T = 8 and a frame count that is not a multiple of the tile"""
    case, frac = _collapsed_case(20003, 500, 8, O.UPTAKE_EYE_MSE, 30.0, seed=5)
    assert frac > 0.8
    cfg = O.OptConfig(learning_rate=1.0, clip_value=None)
    _check_steps(case, cfg, 4)


def test_bpti_fixture_shape_config1():
    """BASELINE config 1: the real BPTI BV features (53 x 500), T=3."""
    d = np.load(os.path.join(GOLD, "bpti_bv_features.npz"))
    case = H.make_case(500, 53, 3, O.UPTAKE_EYE_MSE, seed=0)
    case["problem"].heavy = d["heavy"].astype(np.float32)
    case["problem"].acceptor = d["acceptor"].astype(np.float32)
    cfg = O.OptConfig(learning_rate=1.0, clip_value=None)
    _check_steps(case, cfg, 6)


@pytest.mark.parametrize("F,R", [(5, 3), (8, 32), (9, 33), (1001, 513), (300, 1000)])
def test_ragged_shapes(F, R):
    """frames not a multiple of the 8-frame tile, rows not a multiple of a warp, R > 512 (2 rows/thread)."""
    case = H.make_case(F, R, 3, O.UPTAKE_EYE_MSE, seed=F + R, n_pep=max(4, R // 3))
    cfg = O.OptConfig(learning_rate=1.0, clip_value=None, optimise_model_parameters=True)
    _check_steps(case, cfg, 4)


def test_forward_outputs_match_oracle():
    case = H.make_case(3001, 500, 8, O.UPTAKE_EYE_MSE, seed=1)
    cfg = O.OptConfig(learning_rate=1.0, clip_value=None)
    eng = H.make_engine(case, cfg)
    rng = np.random.default_rng(0)
    case["params"].z = rng.dirichlet(np.full(3001, 0.5)).astype(np.float32)  # non-uniform start
    H.init_engine(eng, case)
    _sync()
    st = O.optimizer_initialise(case["params"], cfg, np.float64)
    _, npar, out = O.compute_loss(case["problem"], st.params, np.float64)
    np.testing.assert_allclose(eng.weights.cpu().numpy(), npar.z, rtol=3e-6, atol=1e-12)
    np.testing.assert_allclose(eng.avg.cpu().numpy()[:500], out["avg_heavy"], rtol=2e-6)
    np.testing.assert_allclose(eng.avg.cpu().numpy()[500:], out["avg_acceptor"], rtol=2e-6)
    np.testing.assert_allclose(eng.log_pf.cpu().numpy(), out["log_Pf"], rtol=2e-6)
    np.testing.assert_allclose(eng.uptake_out.cpu().numpy(), out["uptake"], rtol=1e-5, atol=1e-7)
    assert abs(float(eng.weights.sum()) - 1.0) < 1e-5


def test_frame_weights_frozen_when_not_optimised():
    case = H.make_case(256, 20, 3, O.UPTAKE_EYE_MSE, seed=5, n_pep=8)
    cfg = O.OptConfig(learning_rate=1.0, clip_value=None, optimise_frame_weights=False, optimise_model_parameters=True)
    eng = H.make_engine(case, cfg)
    H.init_engine(eng, case)
    z0 = eng.export_state(("z",))[0]["z"].clone()
    eng.step(5)
    _sync()
    bufs, sc = eng.export_state(("z",))
    assert torch.equal(bufs["z"], z0)
    assert sc["bc"] != 0.35


def test_determinism_and_graph_replay_bitwise():
    case = H.make_case(4099, 130, 5, O.UPTAKE_EYE_MSE, seed=8)
    cfg = O.OptConfig(learning_rate=1.0, clip_value=None, optimise_model_parameters=True)
    outs = []
    for mode in ("eager", "eager", "graph"):
        eng = H.make_engine(case, cfg)
        H.init_engine(eng, case)
        if mode == "graph":
            eng.capture(2)
            eng.step_graph(5)
        else:
            eng.step(10)
        _sync()
        bufs, sc = eng.export_state()
        outs.append((bufs["z"].clone(), bufs["m"].clone(), bufs["g"].clone(), sc["bc"], eng.records(9, 1)[0].total_train))
    for o in outs[1:]:
        assert torch.equal(o[0], outs[0][0]) and torch.equal(o[1], outs[0][1]) and torch.equal(o[2], outs[0][2])
        assert o[3] == outs[0][3] and o[4] == outs[0][4]


def test_weights_after_1000_steps_config1_and_2():
    """north_star: frame weights within 1e-4 after 1000 steps (vs the fp32 mirror of the reference)."""
    for F, R, T, kind in ((500, 53, 3, O.UPTAKE_EYE_MSE), (1000, 130, 1, O.PF_L2)):
        case = H.make_case(F, R, T, kind, seed=21, scaling=100.0)
        cfg = O.OptConfig(learning_rate=0.05, initial_learning_rate=1.0, initial_steps=2, clip_value=None)
        eng = H.make_engine(case, cfg)
        H.init_engine(eng, case)
        eng.step(1000)
        _sync()
        st = O.optimizer_initialise(case["params"], cfg, np.float32)
        for _ in range(1000):
            st, loss, w = O.step_with_rates(case["problem"], st, cfg, np.float32)
        wg = eng.weights.cpu().numpy()
        assert np.abs(wg - w).max() < 1e-4
        rec = eng.records(999, 1)[0]
        np.testing.assert_allclose(rec.total_train, loss, rtol=1e-3)
        assert abs(wg.sum() - 1) < 1e-5


def test_weights_after_1000_steps_example_settings():
    """The examples' settings (run_maxent_parallel_BV_aSyn_conditions.sh:24-31: lr = 1.0, forward_model_scaling = 1000,
    Adam, no clipping) for 1000 free-running steps: frame weights within north_star's 1e-4 of the fp64 oracle AND of the
    fp32 mirror of the reference's arithmetic (measured drift on B200: < 1e-6, tools/explore_r2.py)."""
    for F, R, T, kind in ((500, 53, 3, O.UPTAKE_EYE_MSE), (1000, 130, 1, O.PF_L2)):
        case = H.make_case(F, R, T, kind, seed=21, scaling=1000.0)
        cfg = O.OptConfig(learning_rate=1.0, initial_learning_rate=1.0, initial_steps=0, clip_value=None)
        eng = H.make_engine(case, cfg)
        H.init_engine(eng, case)
        eng.step(1000)
        _sync()
        wg = eng.weights.cpu().numpy()
        for dt in (np.float32, np.float64):
            st = O.optimizer_initialise(case["params"], cfg, dt)
            for _ in range(1000):
                st, loss, w = O.step_with_rates(case["problem"], st, cfg, dt)
            assert np.abs(wg - w).max() < 1e-4
            np.testing.assert_allclose(eng.records(999, 1)[0].total_train, loss, rtol=1e-3 if dt == np.float32 else 1e-5)
        assert abs(wg.sum() - 1) < 1e-5


def _torch_reference_step(heavy, acc, k, tp, S, y, z, bc, bh, fw, fs, prior):
    """Plain torch fp64 restatement of compute_loss + grad for the config-3 size check (runs on the GPU)."""
    f64 = torch.float64
    z = z.to(f64).requires_grad_(True)
    w = torch.softmax(z, 0)
    a, b = heavy.to(f64) @ w, acc.to(f64) @ w
    lp = bc * a + bh * b
    u = 1 - torch.exp(-k.to(f64)[None, :] * tp.to(f64)[:, None] / torch.exp(lp)[None, :])
    r = S.to(f64) @ u.T - y.to(f64)
    l0 = 0.5 * (r * r).sum() / (y.shape[0] * y.shape[1])
    F = z.shape[0]
    eps = 1e-10 / F
    sw = w.abs() + eps
    sw = sw / sw.sum()
    pw = prior.to(f64).abs() + eps
    pw = pw / pw.sum()
    l1 = (pw * (pw.log() - sw.log())).sum() + (sw - pw).sum()
    tot = l0 * fw[0] * fs[0] + l1 * fw[1] * fs[1]
    (g,) = torch.autograd.grad(tot, z)
    return float(tot), g, w.detach()


def test_config3_size_against_torch_fp64():
    """BASELINE config 3 (100k frames x 500 residues x 8 timepoints): loss, gradient and Sum(w)=1 at full size."""
    from jaxent_b200 import _lib as L
    from jaxent_b200 import synthetic
    from jaxent_b200.engine import BvEngine, OptSettings, SlotSpec

    F, R, T = 100_000, 500, 8
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev)
    gen.manual_seed(0)
    heavy = torch.randint(0, 40, (R, F), device=dev, generator=gen, dtype=torch.int32).float()
    acc = torch.randint(0, 3, (R, F), device=dev, generator=gen, dtype=torch.int32).float()
    small = synthetic.make_ensemble(2048, R, T, seed=0)
    slots = [SlotSpec(L.LOSS_UPTAKE_EYE_MSE, small.S_train, small.y_train, small.S_val, small.y_val), SlotSpec(L.LOSS_MAXENT_CONVEX_KL)]
    prior = torch.full((F,), 1.0 / F, device=dev)
    eng = BvEngine(heavy, acc, small.k_ints, small.timepoints, slots, True, OptSettings(learning_rate=1.0, clip_value=None), prior=prior)
    fw, fs = [0.5, 0.5], [1000.0, 1000.0]
    eng.init_state(torch.ones(F, device=dev), 0.35, 2.0, fw, fs)
    k, tp = torch.tensor(small.k_ints, device=dev), torch.tensor(small.timepoints, device=dev)
    S, y = torch.tensor(small.S_train, device=dev), torch.tensor(small.y_train, device=dev)
    for step in range(3):
        z = eng.export_state(("z",))[0]["z"]
        eng.step(1)
        _sync()
        tot, g, w = _torch_reference_step(heavy, acc, k, tp, S, y, z, 0.35, 2.0, fw, fs, prior)
        rec = eng.records(step, 1)[0]
        np.testing.assert_allclose(rec.total_train, tot, rtol=STEP_RTOL)
        gg = eng.export_state(("g",))[0]["g"].to(torch.float64)
        assert float((gg - g).abs().max() / g.abs().max()) < STEP_RTOL
    assert abs(float(eng.weights.double().sum()) - 1.0) < 1e-6


def test_errors_surface_as_exceptions():
    from jaxent_b200 import _lib as L
    from jaxent_b200.engine import BvEngine, OptSettings, SlotSpec

    z = np.zeros((4, 16), np.float32)
    with pytest.raises(ValueError, match="prior"):
        BvEngine(z, z, None, None, [SlotSpec(L.LOSS_MAXENT_CONVEX_KL)], False, OptSettings())
    with pytest.raises(L.JaxentError, match="uptake loss needs"):
        S = np.eye(4, dtype=np.float32)
        BvEngine(z, z, None, None, [SlotSpec(L.LOSS_UPTAKE_EYE_MSE, S, np.zeros((4, 0)), S, np.zeros((4, 0)))], False, OptSettings())
    with pytest.raises(ValueError, match="different shapes"):
        BvEngine(z, np.zeros((4, 8), np.float32), None, None, [SlotSpec(L.LOSS_PARAMS_L2)], False, OptSettings())


@pytest.mark.parametrize("F,R", [(3001, 500), (77, 53), (1001, 513)])
def test_u8_features_match_fp32_storage(F, R):
    """lossless u8 storage of the integer contact counts: same per-element arithmetic, a quarter of the feature bytes.
    Only the grouping of the fp64 partial sums differs (32-frame instead of 8-frame work units), so results agree to
    fp32 rounding rather than bit for bit."""
    case = H.make_case(F, R, 5, O.UPTAKE_EYE_MSE, extra=(O.PARAMS_L2,), seed=F, n_pep=max(4, R // 3))
    cfg = O.OptConfig(learning_rate=1.0, clip_value=None, optimise_model_parameters=True)
    outs = []
    for fd in ("f32", "u8", "auto"):
        eng = H.make_engine(case, cfg, feature_dtype=fd)
        assert eng.features.format == (0 if fd == "f32" else 1)
        H.init_engine(eng, case)
        eng.step(3)  # few steps: lr=1 Adam dynamics amplify last-bit differences quickly
        eng.finish_record()
        _sync()
        bufs, sc = eng.export_state()
        outs.append((bufs["z"].clone(), bufs["m"].clone(), bufs["v"].clone(), bufs["g"].clone(), eng.weights.clone(), sc["bc"],
                     eng.records(3, 1)[0].total_train, eng.actual_bytes_per_step()))
    for o in outs[1:]:
        for a, b in zip(o[:5], outs[0][:5]):
            scale = float(b.abs().max())
            assert float((a - b).abs().max()) <= 2e-5 * scale
        np.testing.assert_allclose(o[5], outs[0][5], rtol=1e-5)
        np.testing.assert_allclose(o[6], outs[0][6], rtol=1e-5)
    assert torch.equal(outs[1][0], outs[2][0])  # "u8" and "auto" take the same path: bitwise
    if F >= 1000:  # tiny ensembles are dominated by padding and the 28 B/frame of optimiser state
        assert outs[1][7] < 0.3 * outs[0][7]


def test_u8_rejected_for_fractional_features():
    case = H.make_case(64, 16, 3, O.UPTAKE_EYE_MSE, seed=2, n_pep=6)
    case["problem"].heavy = case["problem"].heavy + 0.5
    cfg = O.OptConfig(learning_rate=1.0, clip_value=None)
    with pytest.raises(ValueError, match="integer-valued"):
        H.make_engine(case, cfg, feature_dtype="u8")
    eng = H.make_engine(case, cfg, feature_dtype="auto")  # falls back to fp32 storage
    assert eng.features.format == 0
