"""NumPy restatement of JAX-ENT's BV MaxEnt optimise step (CPU oracle).

TEST INFRASTRUCTURE ONLY -- never imported by the product path (see oracle/__init__.py).

Every function cites the reference file:line it follows (paths relative to
/root/reference/jaxent/src).  The oracle runs in a chosen dtype:
``np.float32`` mirrors the reference's arithmetic type (x64 is never enabled there),
``np.float64`` is the "truth" used to bound both the reference-style fp32 noise and the
CUDA path at large F.

Parity pinning status
---------------------
* Against the reference's own code, executed: tests/golden/ref_*.npz are produced by running the UNMODIFIED reference
  sources (/root/reference/jaxent/src) on the torch-backed jax / optax / chex stand-ins of oracle/refstub
  (tests/golden/make_ref_fixtures.py); tests/test_oracle_ref.py compares this file against them: losses and masked
  gradients teacher-forced at 1e-9 (fp64) on six cases incl. collapsed weights and the per-frame branch
  (average_first = False) with optimised BV parameters, free-running trajectories, the Python-loop history and the
  pure-loop carry (_optimise_pure).
* Forward model, losses, loss-weight normalisation: also pinned against the reference's own
  known-answer tests (tests/golden/reference_known_answers.json, transcribed from
  jaxent/tests/unit/... -- see tests/golden/make_golden.py) and, for gradients, against
  torch.autograd in fp64 (tests/test_oracle.py).
* optax 0.2.4 pieces (uv.lock:1125) -- adam / clip / keep_params_nonnegative /
  convex_kl_divergence -- are third-party code that is NOT vendored in the reference and
  not installable here (no jax/optax in this image).  The reference's tests assert no
  numbers for them, so for those pieces **parity is unpinned**: we restate optax's
  published formulas (cited inline).  Independent cross-checks (tests/test_oracle.py): the
  whole update trajectory -- LR pre-scaling, clip, Adam, plateau rule, non-negativity --
  equals torch.optim.Adam(lr=1) driven by torch.autograd fp64 gradients to 1e-9 over 25
  steps, and the convex KL equals torch.nn.functional.kl_div + sum(q - p) to 1e-10.  That
  pins the published algorithms, not optax's own code.
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace
from typing import Optional, Sequence

import numpy as np

# ----------------------------------------------------------------------------- kinds
UPTAKE_EYE_MSE = "hdx_uptake_eye_MSE_loss"  # opt/losses.py:1601-1657
UPTAKE_MC_EYE_MSE = "hdx_uptake_mean_centred_eye_MSE_loss"  # opt/losses.py:1790-1841
UPTAKE_SIGMA_MSE = "hdx_uptake_sigma_MSE_loss"  # opt/losses.py:1541-1598
PF_L2 = "hdx_pf_l2_loss"  # opt/losses.py:17-50
MAXENT_CONVEX_KL = "maxent_convexKL_loss"  # opt/losses.py:304-322
PARAMS_L1 = "model_params_L1_loss"  # opt/losses.py:495-513
PARAMS_L2 = "model_params_L2_loss"  # opt/losses.py:476-492

DATA_KINDS = (UPTAKE_EYE_MSE, UPTAKE_MC_EYE_MSE, UPTAKE_SIGMA_MSE, PF_L2)
UPTAKE_KINDS = (UPTAKE_EYE_MSE, UPTAKE_MC_EYE_MSE, UPTAKE_SIGMA_MSE)


@dataclass
class LossSlot:
    """One (loss function, data target) pair of ``run_optimise`` (opt/run.py:376-406)."""

    kind: str
    # data losses: dense peptide<-residue maps (P,R) and targets.
    # uptake kinds: y (P,T) [the loader's (P,T,1) squeezed, data/loader.py:219-221]; PF: y (P,)
    S_train: Optional[np.ndarray] = None
    y_train: Optional[np.ndarray] = None
    S_val: Optional[np.ndarray] = None
    y_val: Optional[np.ndarray] = None
    W_train: Optional[np.ndarray] = None  # sigma kinds: trace-normalised precision (P,P)
    W_val: Optional[np.ndarray] = None
    # maxent: prior frame weights (F,)
    prior: Optional[np.ndarray] = None
    # params_L1/L2: prior bc / bh
    prior_bc: float = 0.35
    prior_bh: float = 2.0


@dataclass
class Problem:
    heavy: np.ndarray  # (R,F) heavy_contacts   models/HDX/BV/features.py:24
    acceptor: np.ndarray  # (R,F) acceptor_contacts models/HDX/BV/features.py:25
    k_ints: Optional[np.ndarray]  # (R,)
    timepoints: Optional[np.ndarray]  # (T,) static field of BV_Model_Parameters
    slots: Sequence[LossSlot]
    uptake: bool = True  # BV_uptake_ForwardPass (True) or BV_ForwardPass (False)
    # ForwardPass.average_first (models/core.py:299): True = frame-average the features, then the forward pass;
    # False = forward pass per frame, then frame-average the outputs (uptake only; BV_uptake_ForwardPass_frames)
    average_first: bool = True


@dataclass
class Params:
    """Simulation_Parameters (interfaces/simulation.py:17-24), BV model parameters inlined."""

    z: np.ndarray  # frame_weights (logits inside the optimiser, opt/optimiser.py:243)
    frame_mask: np.ndarray
    bc: float
    bh: float
    fw: np.ndarray  # forward_model_weights (L,)
    fs: np.ndarray  # forward_model_scaling (L,)
    nlf: np.ndarray  # normalise_loss_functions (L,)


@dataclass
class Losses:
    """LossComponents (opt/base.py:61-69)."""

    train_losses: np.ndarray
    val_losses: np.ndarray
    scaled_train_losses: np.ndarray
    scaled_val_losses: np.ndarray
    total_train_loss: float
    total_val_loss: float


@dataclass
class Grads:
    z: np.ndarray
    bc: float
    bh: float


@dataclass
class OptConfig:
    """OptaxOptimizer.__init__ arguments (opt/optimiser.py:68-82)."""

    learning_rate: float = 1e-4
    initial_learning_rate: float = 1.0
    initial_steps: int = 0
    plateau_denominator: float = 1.005
    model_parameters_lr_scale: float = 1.0
    clip_value: Optional[float] = 1.0
    optimise_frame_weights: bool = True
    optimise_model_parameters: bool = False
    b1: float = 0.9
    b2: float = 0.999
    eps: float = 1e-8


@dataclass
class AdamState:
    """optax ScaleByAdamState per multi_transform label (frame / model)."""

    count_frame: int = 0
    m_z: Optional[np.ndarray] = None
    v_z: Optional[np.ndarray] = None
    count_model: int = 0
    m_b: np.ndarray = field(default_factory=lambda: np.zeros(2))
    v_b: np.ndarray = field(default_factory=lambda: np.zeros(2))


@dataclass
class OptState:
    """OptimizationState (opt/base.py:81-125) + the dynamic LR / mask carried on the optimiser
    object (opt/optimiser.py:486-491)."""

    params: Params
    adam: AdamState
    step: int = 0
    losses: Optional[Losses] = None
    gradients: Optional[Grads] = None
    current_lr: float = 1.0
    current_model_lr: float = 1.0
    mask_idx: int = 0
    grad_dot: float = 0.0
    new_lr: float = 1.0


# ----------------------------------------------------------------------------- helpers
def normalize_masked_loss_scalingweights(fw, nlf, dtype=np.float32):
    """interfaces/simulation.py:45-76."""
    fw = np.asarray(fw, dtype)
    m = np.clip(np.asarray(nlf, dtype), 0, 1).astype(dtype)
    masked = fw * m
    total = masked.sum(dtype=dtype)
    normalized = masked / total
    return (fw * (dtype(1.0) - m) + normalized * m).astype(dtype)


def softmax(z, dtype=np.float32):
    """jax.nn.softmax (interfaces/simulation.py:112): exp(x - max) / sum."""
    z = np.asarray(z, dtype)
    e = np.exp(z - z.max())
    return (e / e.sum(dtype=dtype)).astype(dtype)


def normalize_weights(p: Params, dtype=np.float32) -> Params:
    """interfaces/simulation.py:103-142.  Model parameters pass through unclamped (:131,:138)."""
    w = softmax(p.z, dtype)
    fm = 1.0 / (1.0 + np.exp(-(10 * (np.asarray(p.frame_mask, dtype) - dtype(0.5)))))
    fm = np.clip(fm, 0, 1).astype(dtype)
    return replace(p, z=w, frame_mask=fm)


def simulation_initialise(p: Params, dtype=np.float32) -> Params:
    """Simulation.initialise models/core.py:59-129 as far as the parameters go: normalize_weights (:72), the loss-weight
    normalisation (:73), then the eager forward whose returned simulation -- with normalize_weights applied AGAIN by
    Simulation.forward (:131-163) -- replaces self.params (:97-98).  The stored frame weights are therefore
    softmax(softmax(w_in)); OptaxOptimizer.initialise multiplies those by F (opt/optimiser.py:243)."""
    q = normalize_weights(p, dtype)
    q = replace(q, fw=normalize_masked_loss_scalingweights(np.asarray(q.fw, dtype), np.asarray(q.nlf, dtype)).astype(dtype))
    return normalize_weights(q, dtype)


def frame_average(x, w, dtype=np.float32):
    """utils/jax_fn.py:15-38: sum(x * w.reshape(1,-1), axis=-1) for ndim>1 leaves."""
    x = np.asarray(x)
    if x.ndim <= 1:
        return x
    # x @ w is the same contraction; BLAS keeps the CPU baseline honest (multi-threaded).
    return (x.astype(dtype, copy=False) @ np.asarray(w, dtype)).astype(dtype)


def bv_log_pf(avg_heavy, avg_acceptor, bc, bh, dtype=np.float32):
    """BV_ForwardPass.__call__ models/HDX/forward.py:18-37."""
    return (dtype(bc) * np.asarray(avg_heavy, dtype) + dtype(bh) * np.asarray(avg_acceptor, dtype)).astype(dtype)


def bv_uptake(log_pf, k_ints, timepoints, dtype=np.float32):
    """BV_uptake_ForwardPass.__call__ models/HDX/forward.py:45-76 -> (T, R[, F])."""
    log_pf = np.asarray(log_pf, dtype)
    pf = np.exp(log_pf)
    k = np.asarray(k_ints, dtype)
    if pf.ndim > 1:
        k = k[:, None]
    t = np.asarray(timepoints, dtype).reshape(-1)
    t = t[(slice(None),) + (None,) * pf.ndim]
    return (dtype(1) - np.exp(-k * t / pf)).astype(dtype)


def apply_sparse_mapping(S, x, dtype=np.float32):
    """data/splitting/sparse_map.py:220-235: squeeze(S.todense() @ x)."""
    return np.squeeze(np.asarray(S, dtype) @ np.asarray(x, dtype)).astype(dtype)


def convex_kl_divergence(log_predictions, targets, dtype=np.float32):
    """optax.losses.convex_kl_divergence (optax 0.2.4, losses/_classification.py):
    kl_divergence(log_q, p) + sum(exp(log_q) - p), with
    kl_divergence = sum(p * (where(p == 0, 0, log p) - log_q)).  PARITY UNPINNED (see header)."""
    p = np.asarray(targets, dtype)
    lq = np.asarray(log_predictions, dtype)
    with np.errstate(divide="ignore"):
        lp = np.where(p == 0, dtype(0), np.log(np.where(p == 0, dtype(1), p)))
    kl = (p * (lp - lq)).sum(dtype=dtype)
    return dtype(kl + (np.exp(lq) - p).sum(dtype=dtype))


# ----------------------------------------------------------------------------- forward + losses
def forward(problem: Problem, params: Params, dtype=np.float32):
    """Simulation.forward + forward_pure (models/core.py:131-163,270-307), both branches of `average_first`.
    Returns (normalised params, outputs dict)."""
    npar = normalize_weights(params, dtype)
    a = frame_average(problem.heavy, npar.z, dtype)
    b = frame_average(problem.acceptor, npar.z, dtype)
    log_pf = bv_log_pf(a, b, npar.bc, npar.bh, dtype)
    out = {"avg_heavy": a, "avg_acceptor": b, "log_Pf": log_pf}
    if problem.uptake and problem.average_first:
        out["uptake"] = bv_uptake(log_pf, problem.k_ints, problem.timepoints, dtype)
    elif problem.uptake:
        # models/core.py:302-304: output = single_pass(fp, feat, param); output = frame_average_features(output, w)
        lp_f = bv_log_pf(problem.heavy, problem.acceptor, npar.bc, npar.bh, dtype)  # (R,F)
        u_f = bv_uptake(lp_f, problem.k_ints, problem.timepoints, dtype)  # (T,R,F)
        out["log_Pf_frames"] = lp_f
        out["uptake_frames"] = u_f
        out["uptake"] = (u_f @ np.asarray(npar.z, dtype)).astype(dtype)  # sum(x * w, axis=-1), utils/jax_fn.py:30-38
    return npar, out


def _uptake_mse(S, y, W, uptake, centre, dtype):
    """compute_loss closures of opt/losses.py:1561-1584 / 1612-1642 / 1811-1828."""
    S = np.asarray(S, dtype)
    y = np.asarray(y, dtype)
    P, T = y.shape
    total = dtype(0)
    for t in range(T):
        true_t = y[:, t]
        pred = apply_sparse_mapping(S, uptake[t], dtype).reshape(-1)
        if centre:
            true_t = true_t - true_t.mean(dtype=dtype)
            pred = pred - pred.mean(dtype=dtype)
        r = (pred - true_t).astype(dtype)
        wr = r if W is None else np.asarray(W, dtype) @ r
        total = dtype(total + dtype(0.5) * np.dot(r, wr))
    return dtype(total / dtype(T * P))


def slot_losses(problem: Problem, slot: LossSlot, npar: Params, out, dtype=np.float32):
    """(train_loss, val_loss) of one slot."""
    k = slot.kind
    if k in UPTAKE_KINDS:
        centre = k == UPTAKE_MC_EYE_MSE
        Wt = slot.W_train if k == UPTAKE_SIGMA_MSE else None
        Wv = slot.W_val if k == UPTAKE_SIGMA_MSE else None
        tr = _uptake_mse(slot.S_train, slot.y_train, Wt, out["uptake"], centre, dtype)
        va = _uptake_mse(slot.S_val, slot.y_val, Wv, out["uptake"], centre, dtype)
        return tr, va
    if k == PF_L2:
        lp = out["log_Pf"].reshape(-1)
        tr = ((apply_sparse_mapping(slot.S_train, lp, dtype) - np.asarray(slot.y_train, dtype).reshape(-1)) ** 2).mean(dtype=dtype)
        va = ((apply_sparse_mapping(slot.S_val, lp, dtype) - np.asarray(slot.y_val, dtype).reshape(-1)) ** 2).mean(dtype=dtype)
        return dtype(tr), dtype(va)
    if k == MAXENT_CONVEX_KL:
        F = slot.prior.shape[0]
        eps = dtype(1e-10 / F)
        sw = np.abs(npar.z) + eps
        sw = sw / sw.sum(dtype=dtype)
        pw = np.abs(np.asarray(slot.prior, dtype)) + eps
        pw = pw / pw.sum(dtype=dtype)
        loss = convex_kl_divergence(np.log(sw), pw, dtype)
        return loss, loss
    if k == PARAMS_L1:
        loss = dtype(abs(dtype(npar.bc) - dtype(slot.prior_bc)) + abs(dtype(npar.bh) - dtype(slot.prior_bh)))
        return loss, loss
    if k == PARAMS_L2:
        loss = dtype((dtype(npar.bc) - dtype(slot.prior_bc)) ** 2 + (dtype(npar.bh) - dtype(slot.prior_bh)) ** 2)
        return loss, loss
    raise ValueError(f"unsupported loss kind {k!r}")


def compute_loss(problem: Problem, params: Params, dtype=np.float32):
    """compute_loss opt/optimiser.py:608-644."""
    npar, out = forward(problem, params, dtype)
    pairs = [slot_losses(problem, s, npar, out, dtype) for s in problem.slots]
    train = np.array([p[0] for p in pairs], dtype)
    val = np.array([p[1] for p in pairs], dtype)
    fw = np.asarray(npar.fw, dtype)
    fs = np.asarray(npar.fs, dtype)
    st = train * fw * fs
    sv = val * fw * fs
    losses = Losses(train, val, st, sv, dtype(st.sum(dtype=dtype)), dtype(sv.sum(dtype=dtype)))
    return losses, npar, out


# ----------------------------------------------------------------------------- closed-form backward
def _dpred_uptake(S, y, W, uptake, centre, scale, dtype):
    """d(scale * uptake-MSE)/d uptake[t, r]  -> (T,R)."""
    S = np.asarray(S, dtype)
    y = np.asarray(y, dtype)
    P, T = y.shape
    g = np.zeros_like(uptake, dtype=dtype)
    for t in range(T):
        true_t = y[:, t]
        pred = S @ uptake[t]
        if centre:
            true_t = true_t - true_t.mean()
            pred = pred - pred.mean()
        r = pred - true_t
        if W is None:
            gr = r
        else:
            Wd = np.asarray(W, dtype)
            gr = dtype(0.5) * (Wd @ r + Wd.T @ r)
        gr = gr * dtype(scale / (T * P))
        if centre:
            gr = gr - gr.mean()
        g[t] = S.T @ gr
    return g


def loss_and_grads(problem: Problem, params: Params, dtype=np.float32):
    """value_and_grad(loss_fn)(state.params) of opt/optimiser.py:383-394, written in closed form
    (SURVEY.md section 8a "Maths the kernels must reproduce").  Returns (Losses, Grads (unmasked),
    aux dict).  Validated against torch.autograd fp64 in tests/test_oracle.py."""
    losses, npar, out = compute_loss(problem, params, dtype)
    w = npar.z
    a, b, lp = out["avg_heavy"], out["avg_acceptor"], out["log_Pf"]
    fw = np.asarray(npar.fw, dtype)
    fs = np.asarray(npar.fs, dtype)
    R = a.shape[0]
    c = np.zeros(R, dtype)  # dL/d log_Pf_r
    hw = np.zeros_like(w)  # explicit dL/dw_f terms (KL)
    per_frame = problem.uptake and not problem.average_first
    hpf = np.zeros_like(w)  # per-frame mode: sum_{t,r} G[t,r] u[t,r,f]
    Gtot = None
    gbc = dtype(0)
    gbh = dtype(0)
    for i, s in enumerate(problem.slots):
        sc = dtype(fw[i] * fs[i])
        k = s.kind
        if k in UPTAKE_KINDS:
            centre = k == UPTAKE_MC_EYE_MSE
            W = s.W_train if k == UPTAKE_SIGMA_MSE else None
            du = _dpred_uptake(s.S_train, s.y_train, W, out["uptake"], centre, sc, dtype)
            if per_frame:
                # uptake[t,r] = sum_f w_f u[t,r,f]: dL/dw_f = sum_{t,r} G[t,r] u[t,r,f]; d u/d lnPF[r,f] = -x exp(-x)
                u_f, lp_f = out["uptake_frames"], out["log_Pf_frames"]
                hpf += np.einsum("tr,trf->f", du, u_f).astype(dtype)
                kk3 = np.asarray(problem.k_ints, dtype)[None, :, None]
                tt3 = np.asarray(problem.timepoints, dtype).reshape(-1, 1, 1)
                x3 = kk3 * tt3 * np.exp(-lp_f)[None, :, :]
                dl = np.einsum("tr,trf->rf", du, -x3 * np.exp(-x3)) * w[None, :]  # dL/d lnPF[r,f]
                gbc += dtype((dl * problem.heavy).sum())
                gbh += dtype((dl * problem.acceptor).sum())
                Gtot = du if Gtot is None else Gtot + du
                continue
            kk = np.asarray(problem.k_ints, dtype)[None, :]
            tt = np.asarray(problem.timepoints, dtype).reshape(-1, 1)
            x = kk * tt * np.exp(-lp)[None, :]
            c += (du * (-x * np.exp(-x))).sum(axis=0).astype(dtype)
        elif k == PF_L2:
            S = np.asarray(s.S_train, dtype)
            y = np.asarray(s.y_train, dtype).reshape(-1)
            r = S @ lp - y
            c += (sc * dtype(2.0 / y.shape[0]) * (S.T @ r)).astype(dtype)
        elif k == MAXENT_CONVEX_KL:
            F = s.prior.shape[0]
            eps = dtype(1e-10 / F)
            swu = np.abs(w) + eps
            Dn = swu.sum(dtype=dtype)
            q = swu / Dn
            pw = np.abs(np.asarray(s.prior, dtype)) + eps
            p = pw / pw.sum(dtype=dtype)
            G = dtype(1) - p / q
            corr = (q - p).sum(dtype=dtype)  # == sum_j G_j q_j, 0 in exact arithmetic
            hw += (sc * (G - corr) / Dn * np.sign(w)).astype(dtype)
        elif k == PARAMS_L1:
            gbc += sc * np.sign(dtype(npar.bc) - dtype(s.prior_bc))
            gbh += sc * np.sign(dtype(npar.bh) - dtype(s.prior_bh))
        elif k == PARAMS_L2:
            gbc += sc * 2 * (dtype(npar.bc) - dtype(s.prior_bc))
            gbh += sc * 2 * (dtype(npar.bh) - dtype(s.prior_bh))
        else:
            raise ValueError(k)
    heavy = problem.heavy.astype(dtype, copy=False)
    acc = problem.acceptor.astype(dtype, copy=False)
    h = heavy.T @ (c * dtype(npar.bc)) + acc.T @ (c * dtype(npar.bh)) + hw + hpf
    hbar = np.dot(w, h)
    gz = (w * (h - hbar)).astype(dtype)
    gbc = dtype(gbc + np.dot(c, a))
    gbh = dtype(gbh + np.dot(c, b))
    aux = {"w": w, "c": c, "h": h, "hbar": hbar, "hbar_kl": float(np.dot(w, hw)) if np.ndim(hw) else 0.0, "out": out, "G": Gtot}
    return losses, Grads(gz, gbc, gbh), aux


# ----------------------------------------------------------------------------- optimiser
def optimizer_initialise(params: Params, cfg: OptConfig, dtype=np.float32) -> OptState:
    """OptaxOptimizer.initialise opt/optimiser.py:221-304: logits z0 = w_init * F (:243)."""
    F = params.z.shape[0]
    p = replace(params, z=(np.asarray(params.z, dtype) * dtype(F)).astype(dtype))
    adam = AdamState(0, np.zeros(F, dtype), np.zeros(F, dtype), 0, np.zeros(2, dtype), np.zeros(2, dtype))
    lr0 = float(cfg.initial_learning_rate)
    return OptState(p, adam, 0, None, None, lr0, lr0 * cfg.model_parameters_lr_scale, 0)


def _adam_update(g, m, v, count, cfg: OptConfig, dtype):
    """optax.scale_by_adam + scale(-1) (optax 0.2.4 _src/transform.py): PARITY UNPINNED.
    mu = (1-b1) g + b1 mu ; nu = (1-b2) g^2 + b2 nu ; count+1 ;
    mu_hat = mu/(1-b1^count) ; nu_hat = nu/(1-b2^count) ; u = -mu_hat/(sqrt(nu_hat)+eps)."""
    g = np.asarray(g, dtype)
    m = (dtype(1 - cfg.b1) * g + dtype(cfg.b1) * m).astype(dtype)
    v = (dtype(1 - cfg.b2) * (g * g) + dtype(cfg.b2) * v).astype(dtype)
    count = count + 1
    bc1 = dtype(1) - np.power(dtype(cfg.b1), dtype(count))
    bc2 = dtype(1) - np.power(dtype(cfg.b2), dtype(count))
    mh = m / bc1
    vh = v / bc2
    u = -(mh / (np.sqrt(vh) + dtype(cfg.eps)))
    return u.astype(dtype), m, v, count


def step_with_rates(problem: Problem, state: OptState, cfg: OptConfig, dtype=np.float32):
    """OptaxOptimizer._step_with_rates + _step (opt/optimiser.py:337-492).
    Returns (new_state, loss_value, save_weights) where save_weights = softmax(new z) (:429-435)."""
    step = int(state.step)
    switched = step >= cfg.initial_steps
    mask_idx = 1 if switched else state.mask_idx
    base_lr = dtype(cfg.learning_rate) if switched else dtype(state.current_lr)
    base_model_lr = (
        dtype(cfg.learning_rate * cfg.model_parameters_lr_scale) if switched else dtype(state.current_model_lr)
    )
    # gradient masks: initial = frame weights only; final = configured set (:253-262, gradients.py:28-93)
    if mask_idx == 0:
        mf, mm = dtype(1), dtype(0)
    else:
        mf = dtype(1 if cfg.optimise_frame_weights else 0)
        mm = dtype(1 if cfg.optimise_model_parameters else 0)

    losses, g, _ = loss_and_grads(problem, state.params, dtype)
    mg = Grads((g.z * mf).astype(dtype), dtype(g.bc * mm), dtype(g.bh * mm))
    prev = state.gradients if state.gradients is not None else mg
    dot = dtype(np.vdot(prev.z, mg.z)) + dtype(prev.bc * mg.bc) + dtype(prev.bh * mg.bh)

    reduce_lr = bool(dot < 0) and step > 1
    new_lr = dtype(base_lr / dtype(cfg.plateau_denominator)) if reduce_lr else base_lr
    new_model_lr = dtype(base_model_lr / dtype(cfg.plateau_denominator)) if reduce_lr else base_model_lr

    sz = (mg.z * new_lr).astype(dtype)
    sb = np.array([mg.bc * new_model_lr, mg.bh * new_model_lr], dtype)
    if cfg.clip_value is not None:  # optax.clip: clamp to [-c, c]
        sz = np.clip(sz, -dtype(cfg.clip_value), dtype(cfg.clip_value))
        sb = np.clip(sb, -dtype(cfg.clip_value), dtype(cfg.clip_value))
    a = state.adam
    uz, m_z, v_z, cf = _adam_update(sz, a.m_z, a.v_z, a.count_frame, cfg, dtype)
    ub, m_b, v_b, cm = _adam_update(sb, a.m_b, a.v_b, a.count_model, cfg, dtype)
    # optax.keep_params_nonnegative: where(p + u < 0, -p, u)   (model chain only, :134)
    pb = np.array([state.params.bc, state.params.bh], dtype)
    ub = np.where(pb + ub < 0, -pb, ub).astype(dtype)
    nz = (np.asarray(state.params.z, dtype) + uz).astype(dtype)
    nb = (pb + ub).astype(dtype)
    new_params = replace(state.params, z=nz, bc=dtype(nb[0]), bh=dtype(nb[1]))
    new_state = OptState(
        new_params,
        AdamState(cf, m_z, v_z, cm, m_b, v_b),
        step + 1,
        losses,
        mg,
        float(new_lr),
        float(new_model_lr),
        mask_idx,
        float(dot),
        float(new_lr),
    )
    return new_state, losses.total_train_loss, softmax(nz, dtype)


# ----------------------------------------------------------------------------- loop
class ConvergenceTracker:
    """opt/track.py:128-197."""

    def __init__(self, convergence, learning_rate, ema_alpha, min_steps_per_threshold):
        self.ema_alpha = ema_alpha
        self.min_steps = min_steps_per_threshold
        thr = np.sort(np.atleast_1d(np.asarray(convergence, np.float32)))[::-1] * np.float32(learning_rate)
        self.thresholds = [float(t) for t in thr]  # track.py:12-21
        self.idx = 0
        self.current = self.thresholds[0]
        self.ema_loss_delta = None
        self.ema_params = None
        self.steps_since = 0

    def update(self, prev_loss, cur_loss, cur_w):
        if prev_loss is not None:
            raw = abs(prev_loss - cur_loss)
            if self.ema_loss_delta is None:
                self.ema_loss_delta, self.ema_params = raw, cur_w
            else:
                self.ema_loss_delta = self.ema_alpha * raw + (1 - self.ema_alpha) * self.ema_loss_delta
                self.ema_params = self.ema_alpha * cur_w + (1 - self.ema_alpha) * self.ema_params
        else:
            raw = 0.0
        self.steps_since += 1
        return raw

    def rel(self, cur_loss):
        if self.ema_loss_delta is not None and cur_loss > 0:
            return float(self.ema_loss_delta / cur_loss)
        return 0.0

    def met(self, cur_loss, step, initial_steps):
        return (
            self.steps_since >= self.min_steps
            and self.ema_loss_delta is not None
            and self.rel(cur_loss) < self.current
            and step > initial_steps
        )

    def advance(self):
        self.idx += 1
        self.steps_since = 0
        if self.idx >= len(self.thresholds):
            return False
        self.current = self.thresholds[self.idx]
        return True


def optimise(problem, params, cfg: OptConfig, n_steps, tolerance=1e-10, convergence=(1e-3,), ema_alpha=0.5,
             min_steps_per_threshold=2, dtype=np.float32):
    """_optimise opt/run.py:235-373 (Python loop).  Returns dict(history, losses, state, best)."""
    state = optimizer_initialise(params, cfg, dtype)
    tr = ConvergenceTracker(list(np.atleast_1d(convergence)), cfg.learning_rate, ema_alpha, min_steps_per_threshold)
    prev_loss = None
    history = []
    loss_trace = []
    for step in range(n_steps):
        prev_g = state.gradients
        state, cur, save_w = step_with_rates(problem, state, cfg, dtype)
        loss_trace.append(float(cur))
        tr.update(prev_loss, float(cur), save_w)
        prev_loss = float(cur)
        # check_gradient_oscillation (opt/gradients.py:15-26): zeros for the first step
        if prev_g is None:
            gdot = 0.0
        else:
            gdot = float(np.vdot(prev_g.z, state.gradients.z) + prev_g.bc * state.gradients.bc + prev_g.bh * state.gradients.bh)
        if gdot < 0:
            tr.steps_since = 0
        if cur < tolerance or not np.isfinite(cur):
            break
        if step == 0 or tr.met(float(cur), step, cfg.initial_steps):
            history.append({"step": state.step, "weights": save_w, "losses": state.losses, "bc": state.params.bc,
                            "bh": state.params.bh, "ema_weights": tr.ema_params})
            if step > 0 and not tr.advance():
                break
    best = None
    if history:
        best = history[-1]
        for h in history:  # OptimizationHistory._pick_best_state opt/base.py:180-189
            if h["losses"].val_losses[0] < best["losses"].val_losses[0]:
                best = h
    return {"history": history, "loss_trace": loss_trace, "state": state, "best": best}


def optimise_pure(problem, params, cfg: OptConfig, n_steps, tolerance=1e-10, convergence=(1e-3,), ema_alpha=0.5,
                  min_steps_per_threshold=2, dtype=np.float32, learning_rate=None):
    """_optimise_pure opt/run.py:141-232 = lax.while_loop over OptaxOptimizer._pure_step (opt/optimiser.py:494-579) with the
    functional tracker update_convergence / check_and_advance_threshold (opt/track.py:40-125) -- the loop batch_optimise
    vmaps (opt/batch.py:107-146).  Differences to the Python loop `optimise` above, all reproduced here:
      * the state enters dense (_ensure_dense_state, run.py:89-118): losses evaluated at the initial parameters, gradients
        = zeros (so the first grad-dot is 0, not |g|^2);
      * previous_loss of step k is the loss stored by step k-1 (the initial loss for k = 0: first raw delta = 0);
      * the EMA of the loss delta and of the parameters RESTARTS whenever steps_since_threshold_start == 0 (at the start and
        after every threshold advance), track.py:53-68;  there is no gradient-oscillation reset;
      * the threshold test uses the 1-based state.step: step + 1 > initial_steps (optimiser.py:545-551);
      * cond_fn (run.py:202-211) stops on converged / step >= n_steps / non-finite or < tolerance loss OF THE PREVIOUS STEP;
      * every step is written to the history: save_state.params (softmax weights AFTER the update) paired with the losses
        evaluated BEFORE it (optimiser.py:555-567);  `learning_rate` overrides cfg.learning_rate (the vmapped hyper-parameter)
        for the thresholds and the target rates."""
    if learning_rate is not None:
        cfg = replace(cfg, learning_rate=float(learning_rate))
    state = optimizer_initialise(params, cfg, dtype)
    losses0, _, _ = loss_and_grads(problem, state.params, dtype)
    zero = Grads(np.zeros_like(np.asarray(state.params.z, dtype)), dtype(0), dtype(0))
    state = replace(state, losses=losses0, gradients=zero)
    thr = np.sort(np.atleast_1d(np.asarray(convergence, np.float32)))[::-1] * np.float32(cfg.learning_rate)  # track.py:12-21
    ema_delta = np.float32(0.0)
    ema_w = np.asarray(state.params.z, dtype).copy()  # initialise_convergence_carry(initial_params), track.py:24-32
    steps_since, idx, converged = 0, 0, False
    F = np.asarray(state.params.z).shape[0]
    hist_w = np.zeros((n_steps, F), dtype)
    hist_total = np.zeros(n_steps, dtype)
    hist_val0 = np.zeros(n_steps, dtype)
    hist_losses = []
    write_idx = 0
    alpha = np.float32(ema_alpha)
    while True:
        cur_prev = state.losses.total_train_loss
        if converged or state.step >= n_steps or not np.isfinite(cur_prev) or cur_prev < np.float32(tolerance):
            break
        prev_loss = np.float32(state.losses.total_train_loss)
        state, loss, save_w = step_with_rates(problem, state, cfg, dtype)
        cur = np.float32(loss)
        raw = np.abs(prev_loss - cur)
        first = steps_since == 0
        ema_delta = raw if first else np.float32(alpha * raw + (np.float32(1.0) - alpha) * ema_delta)
        ema_w = save_w if first else (dtype(ema_alpha) * save_w + (dtype(1.0) - dtype(ema_alpha)) * ema_w).astype(dtype)
        steps_since += 1
        # check_and_advance_threshold
        ti = min(idx, len(thr) - 1)
        rel = np.float32(ema_delta / cur) if cur > 0 else np.float32(0.0)
        met = steps_since >= min_steps_per_threshold and rel < thr[ti] and state.step > cfg.initial_steps
        more = idx + 1 < len(thr)
        if met and more:
            idx += 1
            steps_since = 0
        elif met:
            converged = True
        hist_w[write_idx] = save_w
        hist_total[write_idx] = state.losses.total_train_loss
        hist_val0[write_idx] = state.losses.val_losses[0]
        hist_losses.append(state.losses)
        write_idx += 1
    best_index = None
    if write_idx:
        best_index = write_idx - 1
        for i in range(write_idx):  # OptimizationHistory._pick_best_state opt/base.py:180-189 over ALL steps (batch.py:158-180)
            if hist_val0[i] < hist_val0[best_index]:
                best_index = i
    return {"state": state, "write_idx": write_idx, "converged": converged, "threshold_idx": idx, "steps_since": steps_since,
            "ema_loss_delta": float(ema_delta), "ema_w": ema_w, "history_w": hist_w, "history_total_train": hist_total,
            "history_val0": hist_val0, "history_losses": hist_losses, "best_index": best_index}
