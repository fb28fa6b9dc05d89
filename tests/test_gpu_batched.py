"""Batched (tensor-core) optimise step against the sequential CUDA engine and the oracle.

Reference behaviour being checked: batch_optimise == sequential runs (modules/optimise/test_module_batch_optimise.py:163-229,
loss rtol 1e-4, frame weights atol 1e-4)."""
import numpy as np
import pytest
import torch

import helpers as H
from jaxent_b200 import _lib as L
from oracle import bv_oracle as O

pytestmark = pytest.mark.gpu


def _batched(case, cfg, B, n_pieces=3):
    from jaxent_b200.batch_engine import BatchedBvEngine
    from jaxent_b200.engine import OptSettings, SlotSpec

    pb = case["problem"]
    specs = [SlotSpec(H.KIND_TO_C[s.kind], s.S_train, s.y_train, s.S_val, s.y_val, s.W_train, s.W_val, s.prior_bc, s.prior_bh) for s in pb.slots]
    st = OptSettings(cfg.learning_rate, cfg.initial_learning_rate, cfg.initial_steps, cfg.plateau_denominator, cfg.model_parameters_lr_scale,
                     cfg.clip_value, cfg.optimise_frame_weights, cfg.optimise_model_parameters, cfg.b1, cfg.b2, cfg.eps)
    return BatchedBvEngine(pb.heavy, pb.acceptor, pb.k_ints, pb.timepoints, specs, pb.uptake, st, B, prior=case["prior"], device="cuda:0",
                           n_pieces=n_pieces)


def _hparams(B, n, seed):
    rng = np.random.default_rng(seed)
    fw = rng.uniform(0.1, 1.0, (B, n)).astype(np.float32)
    fw /= fw.sum(1, keepdims=True)
    fs = np.full((B, n), 100.0, np.float32) * rng.uniform(0.5, 2.0, (B, 1)).astype(np.float32)
    lr = rng.choice([0.05, 0.1, 0.2, 0.5], B).astype(np.float32)
    return fw, fs, lr


@pytest.mark.parametrize("kind,extra,opt_model,F,R,T,B", [
    (O.UPTAKE_EYE_MSE, (), False, 777, 53, 3, 5),
    (O.UPTAKE_EYE_MSE, (O.PARAMS_L2,), True, 2000, 130, 5, 20),
    (O.PF_L2, (), True, 1500, 64, 0, 16),
    (O.UPTAKE_MC_EYE_MSE, (), False, 900, 40, 4, 3),
])
def test_batched_step_matches_sequential_engine(kind, extra, opt_model, F, R, T, B):
    case = H.make_case(F, R, max(T, 1), kind, extra=extra, seed=11, prior="random")
    n = len(case["problem"].slots)
    fw, fs, lr = _hparams(B, n, 5)
    cfg = O.OptConfig(learning_rate=0.1, initial_learning_rate=1.0, initial_steps=2, clip_value=None, optimise_model_parameters=opt_model)
    eng = _batched(case, cfg, B)
    p = case["params"]
    z0 = np.asarray(p.z, np.float32) * np.float32(F)
    eng.init_state(z0, p.bc, p.bh, fw, fs, lr)
    n_steps = 12
    eng.step(n_steps)
    torch.cuda.synchronize()
    recs = [eng.records_at(k) for k in range(n_steps + 1)]
    W = eng.weights_all().cpu().numpy()
    bcs = eng.ctl_field(1, np.float32)
    for b in range(B):
        cfg_b = O.OptConfig(learning_rate=float(lr[b]), initial_learning_rate=1.0, initial_steps=2, clip_value=None, optimise_model_parameters=opt_model)
        ref = H.make_engine(case, cfg_b, device="cuda:0")
        ref.init_state(z0, p.bc, p.bh, fw[b], fs[b])
        ref.step(n_steps)
        ref.finish_record()
        torch.cuda.synchronize()
        rrecs = ref.records(0, n_steps + 1)
        for k in range(n_steps + 1):
            a, r = recs[k][b], rrecs[k]
            assert a.step == r.step == k and a.complete == 1
            assert a.branch == r.branch, (b, k)
            np.testing.assert_allclose(a.total_train, r.total_train, rtol=1e-4)
            np.testing.assert_allclose(a.total_val, r.total_val, rtol=1e-4)
            np.testing.assert_allclose(a.lr, r.lr, rtol=1e-6)
        np.testing.assert_allclose(W[:, b], ref.weights.cpu().numpy(), atol=1e-4 / F, rtol=2e-3)
        np.testing.assert_allclose(bcs[b], ref.export_state(())[1]["bc"], rtol=1e-4)
        assert abs(W[:, b].sum() - 1.0) < 1e-5


def test_batched_first_step_matches_oracle():
    F, R, T, B = 1200, 53, 3, 4
    case = H.make_case(F, R, T, O.UPTAKE_EYE_MSE, extra=(O.PARAMS_L2,), seed=3, prior="random")
    n = len(case["problem"].slots)
    fw, fs, lr = _hparams(B, n, 9)
    cfg = O.OptConfig(learning_rate=0.2, clip_value=None, optimise_model_parameters=True)
    eng = _batched(case, cfg, B)
    p = case["params"]
    eng.init_state(np.asarray(p.z, np.float32) * np.float32(F), p.bc, p.bh, fw, fs, lr)
    torch.cuda.synchronize()
    recs = eng.records_at(0)
    for b in range(B):
        par = O.Params(p.z, p.frame_mask, p.bc, p.bh, fw[b], fs[b], p.nlf)
        st = O.optimizer_initialise(par, cfg, np.float64)
        losses, _, _ = O.compute_loss(case["problem"], st.params, np.float64)
        np.testing.assert_allclose(recs[b].total_train, losses.total_train_loss, rtol=1e-5)
        np.testing.assert_allclose(recs[b].total_val, losses.total_val_loss, rtol=1e-5)


@pytest.mark.parametrize("spread", [10.0, 30.0])
def test_batched_gradient_collapsed_weights(spread):
    """VERDICT r1 #1 for the tensor-core path: masked gradients of every replica against the fp64 oracle when most
    frames sit below eps = 1e-10/F (hbar = sum_j w_j (H_j + kappa_j) with its MaxEnt part)."""
    from test_gpu_parity import _collapsed_case

    F, R, T, B = 2000, 40, 4, 16
    case, frac = _collapsed_case(F, R, T, O.UPTAKE_EYE_MSE, spread)
    assert frac > 0.5
    n = len(case["problem"].slots)
    fw, fs, lr = _hparams(B, n, 3)
    fs *= 10.0  # forward_model_scaling ~ 1000 as in the examples
    cfg = O.OptConfig(learning_rate=1.0, clip_value=None)
    eng = _batched(case, cfg, B)
    p = case["params"]
    z0 = np.asarray(p.z, np.float32) * np.float32(F)
    eng.init_state(z0, p.bc, p.bh, fw, fs, lr)
    eng.step(1)
    torch.cuda.synchronize()
    G = eng.gradients_all().cpu().numpy()  # [F][B] masked gradients of the update just applied (state.gradients)
    recs = eng.records_at(0)
    for b in range(B):
        par = O.Params(z0.astype(np.float64), p.frame_mask, p.bc, p.bh, fw[b], fs[b], p.nlf)
        losses, g, aux = O.loss_and_grads(case["problem"], par, np.float64)
        assert abs(aux["hbar_kl"]) > 0.3 * float(fw[b, 1] * fs[b, 1])
        np.testing.assert_allclose(recs[b].total_train, losses.total_train_loss, rtol=1e-5)
        gerr = np.abs(G[:, b] - g.z).max() / np.abs(g.z).max()
        assert gerr < 2e-5, (b, gerr)  # three-piece bf16 split of c: 6e-6 of the column scale (test_gpu_bgemm.py)


def test_batched_rejects_fractional_features():
    from jaxent_b200.batch_engine import pack_gemm_features

    x = torch.rand(8, 64, device="cuda") * 3.0 + 0.123456
    with pytest.raises(ValueError, match="bf16"):
        pack_gemm_features(x, x, "cuda:0")


@pytest.mark.parametrize("B,F", [(128, 1100), (256, 2500)])
def test_fused_gradient_kernel(B, F, monkeypatch):
    """Bn a multiple of 128: the backward product and the gradient are ONE tensor-core kernel (H^T = c^T N with a replica per
    TMEM lane, bt_fused_grad_kernel).  Against the two-kernel path (gemm1 + bt_grad) over several steps, and against the fp64
    oracle's gradient for a few replicas.  F is chosen so that the number of 128-frame blocks is odd (a half-empty tile)."""
    R, T = 53, 3
    case = H.make_case(F, R, T, O.UPTAKE_EYE_MSE, extra=(O.PARAMS_L2,), seed=21, prior="random")
    n = len(case["problem"].slots)
    fw, fs, lr = _hparams(B, n, 13)
    fs *= 10.0
    cfg = O.OptConfig(learning_rate=1.0, clip_value=None, optimise_model_parameters=True)
    p = case["params"]
    z0 = np.asarray(p.z, np.float32) * np.float32(F)
    out = {}
    for fused in ("1", "0"):  # the fused kernel is opt-in (JAXENT_BG_FUSED=1)
        monkeypatch.setenv("JAXENT_BG_FUSED", fused)
        eng = _batched(case, cfg, B)
        eng.init_state(z0, p.bc, p.bh, fw, fs, lr)
        eng.step(1)
        torch.cuda.synchronize()
        G1 = eng.gradients_all().cpu().numpy().copy()
        eng.step(7)
        torch.cuda.synchronize()
        out[fused] = (G1, eng.gradients_all().cpu().numpy().copy(), [eng.records_at(k) for k in range(9)], eng.weights_all().cpu().numpy().copy())
        eng.close() if hasattr(eng, "close") else None
    (G1a, G8a, ra, Wa), (G1b, G8b, rb, Wb) = out["1"], out["0"]
    scale = np.abs(G1b).max(axis=0, keepdims=True)
    assert (np.abs(G1a - G1b) / scale).max() < 2e-6  # same arithmetic per element; only the tensor-core summation order differs
    for k in range(9):
        for b in range(B):
            assert ra[k][b].branch == rb[k][b].branch and ra[k][b].complete == 1, (k, b)
            np.testing.assert_allclose(ra[k][b].total_train, rb[k][b].total_train, rtol=2e-5)
            np.testing.assert_allclose(ra[k][b].grad_dot, rb[k][b].grad_dot, rtol=1e-4, atol=1e-30)
    np.testing.assert_allclose(Wa, Wb, atol=1e-4 / F, rtol=1e-3)
    for b in (0, B // 2 - 1, B // 2, B - 1):  # both replica halves
        par = O.Params(z0.astype(np.float64), p.frame_mask, p.bc, p.bh, fw[b], fs[b], p.nlf)
        _, g, _ = O.loss_and_grads(case["problem"], par, np.float64)
        gerr = np.abs(G1a[:, b] - g.z).max() / np.abs(g.z).max()
        assert gerr < 2e-5, (b, gerr)


def test_config5_size_against_oracle():
    """BASELINE config 5 at full size -- 256 replicas (32 MaxEnt strengths x 8 ... here: one sweep of 256 strengths) on 100 000
    frames x 500 residues x 8 time points, three bf16 pieces -- against the NumPy fp64 oracle: three teacher-forced steps of four
    replicas (both ends and the middle of the sweep): loss of the state, masked gradient of the update, sum of the weights."""
    F, R, T, B = 100_000, 500, 8, 256
    case = H.make_case(F, R, T, O.UPTAKE_EYE_MSE, seed=5)
    n = len(case["problem"].slots)
    gam = np.logspace(0, 3, B).astype(np.float32)
    fw = np.stack([gam / (gam + 1), 1 / (gam + 1)], 1).astype(np.float32)
    fs = np.full((B, n), 1000.0, np.float32)
    lr = np.full(B, 1.0, np.float32)
    cfg = O.OptConfig(learning_rate=1.0, clip_value=None)
    eng = _batched(case, cfg, B)
    p = case["params"]
    z0 = np.asarray(p.z, np.float32) * np.float32(F)
    eng.init_state(z0, p.bc, p.bh, fw, fs, lr)
    which = (0, 85, 170, 255)
    errs = []
    for k in range(3):
        torch.cuda.synchronize()
        zk = eng.z[:, list(which)].cpu().numpy().astype(np.float64)
        eng.step(1)
        torch.cuda.synchronize()
        recs = eng.records_at(k)
        G = eng.gradients_all()[:, list(which)].cpu().numpy()
        for j, b in enumerate(which):
            par = O.Params(zk[:, j], p.frame_mask, p.bc, p.bh, fw[b], fs[b], p.nlf)
            losses, g, _ = O.loss_and_grads(case["problem"], par, np.float64)
            np.testing.assert_allclose(recs[b].train[0], losses.train_losses[0], rtol=1e-5)
            # The MaxEnt value is ~0 here (weights within a few steps of the uniform prior) and is multiplied by
            # forward_model_scaling = 1000.  The batched path forms sum_f p_f log w_f from the logits (no logarithm per
            # (frame, replica)): log w_f = z_f - lse - log S, exact up to the common relative error of the fp32 exponentials in
            # S (a few 1e-8 when all logits are equal).  So: absolute 1e-7 on the KL value, i.e. 1e-7 * fw * fs on the total.
            np.testing.assert_allclose(recs[b].train[1], losses.train_losses[1], rtol=1e-5, atol=1e-7)
            np.testing.assert_allclose(recs[b].total_train, losses.total_train_loss, rtol=1e-5, atol=1e-7 * float(fw[b, 1] * fs[b, 1]))
            gerr = np.abs(G[:, j] - g.z).max() / np.abs(g.z).max()
            # three-piece bf16 split of c (6e-6 of the column scale, test_gpu_bgemm.py) on CENTRED features: without the centring
            # the fp32 column sums H_f ~ 1e4 lose the cancellation H_f - hbar and this error is 3e-4 at R = 500 (measured)
            errs.append((k, b, float(abs(recs[b].train[0] / losses.train_losses[0] - 1)), float(gerr)))
            assert gerr < 2e-5, (k, b, gerr)
    print("config-5 size: (step, replica, data-loss rel err, gradient rel err)", errs)
    W = eng.weights_all()
    assert float((W.double().sum(0) - 1.0).abs().max()) < 1e-5


@pytest.mark.parametrize("kind,opt_model", [(O.UPTAKE_EYE_MSE, False), (O.PF_L2, True)])
def test_batched_pure_loop_tracker_matches_oracle(kind, opt_model):
    """The loop batch_optimise vmaps is _optimise_pure (opt/run.py:141-232) with update_convergence / check_and_advance_threshold
    (opt/track.py:40-125), not the Python loop: EMAs restart with every threshold, no oscillation reset, 1-based step test,
    dense history, best state = smallest val_losses[0] over all steps.  Device tracker (jaxent_bg_track_mode 1) against
    oracle.optimise_pure, which is pinned on the reference-executed fixtures (tests/test_oracle_ref.py)."""
    F, R, T, B = 600, 40, 3, 5
    case = H.make_case(F, R, T if kind != O.PF_L2 else 1, kind, seed=21, prior="random", scaling=100.0)
    n = len(case["problem"].slots)
    fw, fs, lr = _hparams(B, n, 13)
    cfg = O.OptConfig(learning_rate=0.1, initial_learning_rate=1.0, initial_steps=2, clip_value=None, optimise_model_parameters=opt_model)
    conv, n_steps = [3.0, 1.0, 0.3], 80  # replicas end after 7 ... 80 steps (oracle, fp32 and fp64 agree)
    eng = _batched(case, cfg, B)
    eng.enable_tracking(conv, 0.5, 2, 1e-10, loop="pure")
    p = case["params"]
    eng.init_state(np.asarray(p.z, np.float32) * np.float32(F), p.bc, p.bh, fw, fs, lr)
    status = eng.run(n_steps, chunk=7, keep_records=True)
    torch.cuda.synchronize()
    W = eng.weights_all().cpu().numpy()
    steps_since = eng.ctl_field(5, np.int32)
    tot = eng.record_history.view(np.float32)[:, :, 6]
    different_lengths = set()
    for b in range(B):
        par = O.Params(p.z, p.frame_mask, p.bc, p.bh, fw[b], fs[b], p.nlf)
        ref = O.optimise_pure(case["problem"], par, cfg, n_steps, tolerance=1e-10, convergence=conv, ema_alpha=0.5, min_steps_per_threshold=2,
                              dtype=np.float64, learning_rate=float(lr[b]))
        st = status[b]
        nw = ref["write_idx"]
        different_lengths.add(nw)
        assert st.steps_done == nw, (b, st.steps_done, nw)
        assert (st.reason == 1) == ref["converged"] and bool(st.stop) == ref["converged"]
        assert st.threshold_idx == ref["threshold_idx"] and steps_since[b] == ref["steps_since"]
        np.testing.assert_allclose(st.ema_loss_delta, ref["ema_loss_delta"], rtol=2e-3)
        # free-running for up to 80 steps (the BV-parameter case at lr = 0.5 swings the loss over three decades): the reference's own
        # batch test allows rtol 1e-4 on losses (test_module_batch_optimise.py:163-229); observed 1e-4 here
        np.testing.assert_allclose(tot[:nw, b], ref["history_total_train"][:nw], rtol=5e-4)
        np.testing.assert_allclose(W[:, b], ref["history_w"][nw - 1], rtol=2e-3, atol=1e-4 / F)
        sn, w_best, _ = eng.track_snapshots(b, 1)[0]
        assert st.n_snapshots == 1
        # the device follows the first strict minimum of val_losses[0]; _pick_best_state prefers the last entry on a tie with it
        v = ref["history_val0"][:nw]
        assert sn.loop_step == int(np.argmin(v)) or np.isclose(v[sn.loop_step], v.min(), rtol=1e-6)
        if v.min() < v[-1]:
            assert sn.loop_step == ref["best_index"]
        np.testing.assert_allclose(w_best.cpu().numpy(), ref["history_w"][sn.loop_step], rtol=2e-3, atol=1e-4 / F)
        np.testing.assert_allclose(sn.losses.total_train, ref["history_total_train"][sn.loop_step], rtol=5e-4)
        np.testing.assert_allclose(eng._trk_ema[:, b].cpu().numpy(), ref["ema_w"], rtol=2e-3, atol=1e-4 / F)
    assert len(different_lengths) >= 3  # the replicas really ended at different steps
