"""Shared builders: one synthetic ensemble -> (oracle Problem/Params/OptConfig, BvEngine)."""
from __future__ import annotations

import numpy as np

from jaxent_b200 import _lib as L
from jaxent_b200 import synthetic
from oracle import bv_oracle as O

KIND_TO_C = {
    O.UPTAKE_EYE_MSE: L.LOSS_UPTAKE_EYE_MSE,
    O.UPTAKE_MC_EYE_MSE: L.LOSS_UPTAKE_MC_EYE_MSE,
    O.UPTAKE_SIGMA_MSE: L.LOSS_UPTAKE_SIGMA_MSE,
    O.PF_L2: L.LOSS_PF_L2,
    O.MAXENT_CONVEX_KL: L.LOSS_MAXENT_CONVEX_KL,
    O.PARAMS_L1: L.LOSS_PARAMS_L1,
    O.PARAMS_L2: L.LOSS_PARAMS_L2,
}


def spd(n, seed):
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(n, n))
    W = A @ A.T + n * np.eye(n)
    return (W * (n / np.trace(W))).astype(np.float32)


def make_case(F, R, T, data_kind=O.UPTAKE_EYE_MSE, extra=(), seed=0, n_pep=None, gamma=1.0, scaling=1000.0,
              prior="uniform", z0=None, bc=0.35, bh=2.0):
    """Returns dict(problem, params, ens) -- the oracle-side description of a run.
    Loss slots: [data_kind, maxent_convexKL, *extra]; forward_model_weights = [gamma, 1, 1...]
    normalised over all slots (examples/common/optimization.py:169-188)."""
    uptake = data_kind != O.PF_L2
    ens = synthetic.make_ensemble(F, R, T, n_pep=n_pep, uptake=uptake, seed=seed)
    rng = np.random.default_rng(seed + 100)
    if prior == "uniform":
        pr = ens.prior
    else:
        pr = rng.dirichlet(np.ones(F)).astype(np.float32)
    Wt = spd(ens.S_train.shape[0], seed + 7) if data_kind == O.UPTAKE_SIGMA_MSE else None
    Wv = spd(ens.S_val.shape[0], seed + 8) if data_kind == O.UPTAKE_SIGMA_MSE else None
    slots = [O.LossSlot(data_kind, ens.S_train, ens.y_train, ens.S_val, ens.y_val, Wt, Wv), O.LossSlot(O.MAXENT_CONVEX_KL, prior=pr)]
    for k in extra:
        slots.append(O.LossSlot(k, prior_bc=0.3, prior_bh=2.2))
    n = len(slots)
    fw = O.normalize_masked_loss_scalingweights(np.array([gamma] + [1.0] * (n - 1)), np.ones(n))
    fs = np.full(n, scaling, np.float32)
    w0 = np.full(F, 1.0 / F, np.float32) if z0 is None else np.asarray(z0, np.float32)
    params = O.Params(w0, np.ones(F, np.float32), bc, bh, fw, fs, np.ones(n, np.float32))
    problem = O.Problem(ens.heavy, ens.acceptor, ens.k_ints if uptake else None, ens.timepoints if uptake else None, slots, uptake)
    return {"problem": problem, "params": params, "ens": ens, "prior": pr}


def make_engine(case, cfg: O.OptConfig, device=None, frames=None, F_total=None, group=None, feature_dtype="f32", exchange="auto"):
    """BvEngine over the same arrays (optionally a frame shard `frames` = slice)."""
    from jaxent_b200.engine import BvEngine, OptSettings, SlotSpec

    pb = case["problem"]
    sl = slice(None) if frames is None else frames
    specs = []
    for s in pb.slots:
        specs.append(SlotSpec(KIND_TO_C[s.kind], s.S_train, s.y_train, s.S_val, s.y_val, s.W_train, s.W_val, s.prior_bc, s.prior_bh))
    st = OptSettings(cfg.learning_rate, cfg.initial_learning_rate, cfg.initial_steps, cfg.plateau_denominator,
                     cfg.model_parameters_lr_scale, cfg.clip_value, cfg.optimise_frame_weights, cfg.optimise_model_parameters,
                     cfg.b1, cfg.b2, cfg.eps)
    eng = BvEngine(np.ascontiguousarray(pb.heavy[:, sl]), np.ascontiguousarray(pb.acceptor[:, sl]), pb.k_ints, pb.timepoints, specs,
                   pb.uptake, st, prior=case["prior"][sl], device=device, F_total=F_total, group=group, feature_dtype=feature_dtype,
                   exchange=exchange, average_first=getattr(pb, "average_first", True))
    return eng


def init_engine(eng, case, frames=None):
    p = case["params"]
    sl = slice(None) if frames is None else frames
    F_total = p.z.shape[0]
    z0 = (np.asarray(p.z, np.float32) * np.float32(F_total))[sl]  # opt/optimiser.py:243
    eng.init_state(z0, p.bc, p.bh, p.fw, p.fs)
