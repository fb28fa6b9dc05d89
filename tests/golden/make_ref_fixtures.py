"""Generates tests/golden/ref_*.npz by EXECUTING THE REFERENCE'S OWN SOURCE FILES (/root/reference/jaxent/src, unmodified)
on seeded cases.  jax / optax / chex are not installable in this image, so the reference runs on the torch-backed stand-ins
of oracle/refstub (what is substituted and what stays unpinned: oracle/refstub/README.md).  Run in the build container:

    python tests/golden/make_ref_fixtures.py

The inputs of every case are rebuilt by tests/helpers.make_case from the seeds stored in the file, so the fixtures stay small
(outputs only).  tests/test_oracle_ref.py compares oracle/bv_oracle.py against them.

Reference entry points executed (file:line in /root/reference/jaxent/src):
  Simulation.initialise / forward / forward_pure      models/core.py:59-163,270-307
  frame_average_features                               utils/jax_fn.py:15-38
  BV_ForwardPass / BV_uptake_ForwardPass               models/HDX/forward.py:15-76
  Simulation_Parameters.normalize_*                    interfaces/simulation.py:45-142
  hdx_uptake_eye / mean_centred_eye / sigma MSE, hdx_pf_l2, maxent_convexKL, model_params_L1/L2   opt/losses.py
  compute_loss, OptaxOptimizer.initialise / _step_with_rates / _step / _pure_step                  opt/optimiser.py
  create_gradient_masks / mask_gradients                opt/gradients.py
  _optimise, _optimise_pure, ConvergenceTracker, update_convergence, check_and_advance_threshold  opt/run.py, opt/track.py
  batch_optimise (vmap of _optimise_pure over the hyper-parameter batch), _pick_best_state        opt/batch.py, opt/base.py
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.refstub import load as refload  # noqa: E402

CASES = {
    # name: dict(F, R, T, kind, extra, seed, gamma, scaling, prior, opt_model, clip, lr, init_lr, init_steps, spread)
    "uptake_eye": dict(F=300, R=24, T=3, kind="hdx_uptake_eye_MSE_loss", extra=(), seed=31, gamma=2.0, scaling=1000.0, prior="uniform",
                       opt_model=False, clip=None, lr=1.0, init_lr=1.0, init_steps=0, spread=0.0),
    "uptake_mc_l1": dict(F=257, R=19, T=4, kind="hdx_uptake_mean_centred_eye_MSE_loss", extra=("model_params_L1_loss",), seed=32, gamma=1.0,
                         scaling=100.0, prior="random", opt_model=True, clip=1.0, lr=0.5, init_lr=1.0, init_steps=2, spread=0.0),
    "uptake_sigma": dict(F=200, R=16, T=3, kind="hdx_uptake_sigma_MSE_loss", extra=(), seed=33, gamma=1.0, scaling=1000.0, prior="uniform",
                         opt_model=False, clip=None, lr=1.0, init_lr=1.0, init_steps=0, spread=0.0),
    "pf_l2": dict(F=400, R=30, T=1, kind="hdx_pf_l2_loss", extra=("model_params_L2_loss",), seed=34, gamma=1.0, scaling=1000.0, prior="random",
                  opt_model=True, clip=None, lr=1.0, init_lr=1.0, init_steps=0, spread=0.0),
    "collapsed": dict(F=500, R=20, T=4, kind="hdx_uptake_eye_MSE_loss", extra=(), seed=35, gamma=1.0, scaling=1000.0, prior="uniform",
                      opt_model=False, clip=None, lr=1.0, init_lr=1.0, init_steps=0, spread=10.0),
    # ForwardPass.average_first = False (models/core.py:302-304 + the 2-D branch of BV_uptake_ForwardPass, models/HDX/forward.py:57-74)
    # with the BV parameters optimised: pins the per-frame closed forms of the oracle, d loss / d (bc, bh) included
    "uptake_perframe": dict(F=220, R=18, T=3, kind="hdx_uptake_eye_MSE_loss", extra=("model_params_L2_loss",), seed=36, gamma=2.0, scaling=100.0,
                            prior="random", opt_model=True, clip=None, lr=0.3, init_lr=1.0, init_steps=1, spread=0.0, per_frame=True),
}
N_STEPS = 12


def build_case(spec):
    import helpers as H

    case = H.make_case(spec["F"], spec["R"], max(spec["T"], 1), spec["kind"], extra=tuple(spec["extra"]), seed=spec["seed"], gamma=spec["gamma"],
                       scaling=spec["scaling"], prior=spec["prior"])
    if spec.get("per_frame"):
        case["problem"].average_first = False
    return case


def case_from_arrays(heavy, acc, small):
    """the bench workload (bench.py: features + the small peptide problem) as an oracle-style case dict"""
    from oracle import bv_oracle as O

    F = heavy.shape[1]
    prior = np.full(F, 1.0 / F, np.float32)
    slots = [O.LossSlot(O.UPTAKE_EYE_MSE, small.S_train, small.y_train, small.S_val, small.y_val), O.LossSlot(O.MAXENT_CONVEX_KL, prior=prior)]
    pb = O.Problem(heavy, acc, small.k_ints, small.timepoints, slots, True)
    fw = O.normalize_masked_loss_scalingweights(np.array([1.0, 1.0]), np.ones(2))
    par = O.Params(prior.copy(), np.ones(F, np.float32), 0.35, 2.0, fw, np.full(2, 1000.0, np.float32), np.ones(2, np.float32))
    return {"problem": pb, "params": par, "prior": prior}


def to_np(x):
    import torch

    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


class RefRun:
    """the reference's objects for one case"""

    def __init__(self, case, spec, x64, real_jax=False):
        if real_jax:  # bench.py --impl reference on a box that has jax + optax and the reference package on sys.path
            import jax
        else:
            jax = refload.load(x64=x64)
        import jax.numpy as jnp
        from jax.experimental import sparse

        from jaxent.src.custom_types.config import Optimisable_Parameters
        from jaxent.src.data.loader import Dataset, ExpD_Dataloader
        from jaxent.src.data.splitting.mapping import SparseFragmentMapping
        from jaxent.src.interfaces.simulation import Simulation_Parameters
        from jaxent.src.models.config import BV_model_Config
        from jaxent.src.models.core import Simulation
        from jaxent.src.models.HDX.BV.features import BV_input_features
        from jaxent.src.models.HDX.BV.forwardmodel import BV_model
        from jaxent.src.models.HDX.BV.parameters import BV_Model_Parameters
        from jaxent.src.models.HDX.forward import BV_ForwardPass, BV_uptake_ForwardPass
        from jaxent.src.opt import losses as RL
        from jaxent.src.opt.optimiser import OptaxOptimizer

        self.jax, self.jnp, self.SP = jax, jnp, Simulation_Parameters
        pb, par = case["problem"], case["params"]
        fdt = jnp.float64 if x64 else jnp.float32
        feats = BV_input_features(heavy_contacts=jnp.asarray(pb.heavy, dtype=fdt), acceptor_contacts=jnp.asarray(pb.acceptor, dtype=fdt),
                                  k_ints=jnp.asarray(pb.k_ints if pb.k_ints is not None else np.ones(pb.heavy.shape[0]), dtype=fdt))
        tp = jnp.asarray(pb.timepoints if pb.uptake else np.array([0.167, 1.0, 10.0]), dtype=fdt)
        cfg = BV_model_Config(num_timepoints=int(tp.shape[0]), timepoints=tp) if pb.uptake else BV_model_Config()
        model = BV_model(cfg)  # .forwardpass follows the config key: BV_uptake_ForwardPass with time points, else BV_ForwardPass
        if not getattr(pb, "average_first", True):
            # what the examples' BV_uptake_ForwardPass_frames selects (examples/common/optimization.py:267-296): the forward pass on
            # the un-averaged features, outputs frame-averaged afterwards -- the reference's own class with its flag flipped
            model.forwardpass.average_first = False
        mp = BV_Model_Parameters(bv_bc=jnp.asarray([par.bc], dtype=fdt), bv_bh=jnp.asarray([par.bh], dtype=fdt), timepoints=tp)
        self.loss_fns, self.targets = [], []
        name2fn = {n: getattr(RL, n) for n in ("hdx_uptake_eye_MSE_loss", "hdx_uptake_mean_centred_eye_MSE_loss", "hdx_uptake_sigma_MSE_loss",
                                               "hdx_pf_l2_loss", "maxent_convexKL_loss", "model_params_L1_loss", "model_params_L2_loss")}
        for s in pb.slots:
            self.loss_fns.append(name2fn[s.kind])
            if s.S_train is not None:
                def ds(S, y, W):
                    yy = np.asarray(y, np.float64)
                    yy = yy.reshape(yy.shape[0], -1, 1) if pb.uptake else yy.reshape(-1)
                    return Dataset(data=[], y_true=jnp.asarray(yy, dtype=fdt),
                                   data_mapping=SparseFragmentMapping(sparse_map=sparse.BCOO.fromdense(jnp.asarray(S, dtype=fdt))),
                                   covariance_matrix=None if W is None else jnp.asarray(W, dtype=fdt))
                self.targets.append(ExpD_Dataloader(train=ds(s.S_train, s.y_train, s.W_train), val=ds(s.S_val, s.y_val, s.W_val)))
            elif s.kind == "maxent_convexKL_loss":
                F = pb.heavy.shape[1]
                self.targets.append(Simulation_Parameters(
                    frame_weights=jnp.asarray(s.prior, dtype=fdt), frame_mask=jnp.ones(F, dtype=fdt), model_parameters=[mp],
                    forward_model_weights=jnp.asarray(par.fw, dtype=fdt), forward_model_scaling=jnp.asarray(par.fs, dtype=fdt),
                    normalise_loss_functions=jnp.asarray(par.nlf, dtype=fdt)))
            else:  # parameter regularisers: a Simulation_Parameters holding the prior BV parameters is the data target
                F = pb.heavy.shape[1]
                prior_mp = BV_Model_Parameters(bv_bc=jnp.asarray([s.prior_bc], dtype=fdt), bv_bh=jnp.asarray([s.prior_bh], dtype=fdt), timepoints=tp)
                self.targets.append(Simulation_Parameters(
                    frame_weights=jnp.ones(F, dtype=fdt) / F, frame_mask=jnp.ones(F, dtype=fdt), model_parameters=[prior_mp],
                    forward_model_weights=jnp.asarray(par.fw, dtype=fdt), forward_model_scaling=jnp.asarray(par.fs, dtype=fdt),
                    normalise_loss_functions=jnp.asarray(par.nlf, dtype=fdt)))
        F = pb.heavy.shape[1]
        # un-normalised loss weights go in; Simulation.initialise normalises them (models/core.py:73)
        n = len(pb.slots)
        fw_raw = np.array([spec["gamma"]] + [1.0] * (n - 1), np.float64)
        params = Simulation_Parameters(
            frame_weights=jnp.asarray(par.z, dtype=fdt), frame_mask=jnp.asarray(par.frame_mask, dtype=fdt), model_parameters=[mp],
            forward_model_weights=jnp.asarray(fw_raw, dtype=fdt), forward_model_scaling=jnp.asarray(par.fs, dtype=fdt),
            normalise_loss_functions=jnp.asarray(par.nlf, dtype=fdt))
        self.sim = Simulation(input_features=[feats], forward_models=[model], params=params)
        assert isinstance(self.sim.forwardpass[0], BV_uptake_ForwardPass if pb.uptake else BV_ForwardPass)
        self.sim.initialise()
        masks = {Optimisable_Parameters.frame_weights}
        if spec["opt_model"]:
            masks = masks | {Optimisable_Parameters.model_parameters}
        self.optimizer = OptaxOptimizer(learning_rate=spec["lr"], optimizer="adam", parameter_partition_masks=masks, clip_value=spec["clip"],
                                        initial_learning_rate=spec["init_lr"], initial_steps=spec["init_steps"])
        self.indexes = tuple(0 for _ in pb.slots)
        self.state = self.optimizer.initialise(self.sim)
        if spec["spread"] > 0:
            # collapsed-weights case: the logits are set directly (Simulation.initialise would softmax them away);
            # everything else of the state is the reference's own
            from jaxent.src.opt.base import OptimizationState

            z0 = np.random.default_rng(spec["seed"] + 4).normal(0.0, spec["spread"], spec["F"]).astype(np.float32)
            p0 = self.state.params
            self.state = OptimizationState(
                params=Simulation_Parameters(frame_weights=jnp.asarray(z0, dtype=fdt), frame_mask=p0.frame_mask, model_parameters=p0.model_parameters,
                                             forward_model_weights=p0.forward_model_weights, forward_model_scaling=p0.forward_model_scaling,
                                             normalise_loss_functions=p0.normalise_loss_functions),
                opt_state=self.state.opt_state)


def run_case(name, spec, x64):
    case = build_case(spec)
    rr = RefRun(case, spec, x64)
    from jaxent.src.opt.optimiser import compute_loss

    out = {}
    sim, opt, state = rr.sim, rr.optimizer, rr.state
    out["fw_normalised"] = to_np(sim.params.forward_model_weights)
    # ---- forward + losses at the initial state
    losses, sim2 = compute_loss(sim, state.params, tuple(rr.targets), rr.indexes, tuple(rr.loss_fns))
    o = sim2.outputs[0]
    out["log_Pf0"] = to_np(o.log_Pf) if hasattr(o, "log_Pf") and getattr(o, "log_Pf", None) is not None else np.zeros(0)
    if hasattr(o, "uptake"):
        out["uptake0"] = to_np(o.uptake)
    # ---- free-running steps through the reference's own _step
    rec = {k: [] for k in ("train", "val", "scaled_train", "total_train", "total_val", "lr", "mask_idx", "grad_dot", "bc", "bh", "g_bc", "g_bh")}
    z_hist, g_hist, w_hist = [], [], []
    for k in range(N_STEPS):
        z_hist.append(to_np(state.params.frame_weights).copy())
        (new_state, loss_value, save_state, sim, new_lr, new_model_lr, new_mask_idx, grad_dot) = opt._step_with_rates(
            optimizer=opt, state=state, simulation=sim, data_targets=tuple(rr.targets), loss_functions=tuple(rr.loss_fns), indexes=rr.indexes,
            lr=rr.jnp.asarray(opt._current_lr, dtype=rr.jnp.float32), model_lr=rr.jnp.asarray(opt._current_model_lr, dtype=rr.jnp.float32),
            target_lr=rr.jnp.asarray(opt.learning_rate, dtype=rr.jnp.float32),
            target_model_lr=rr.jnp.asarray(opt.learning_rate * opt.model_parameters_lr_scale, dtype=rr.jnp.float32),
            gradient_mask_idx=rr.jnp.asarray(opt._current_gradient_mask_idx, dtype=rr.jnp.int32))
        opt._current_lr, opt._current_model_lr, opt._current_gradient_mask_idx = float(new_lr), float(new_model_lr), int(new_mask_idx)
        L = new_state.losses
        rec["train"].append(to_np(L.train_losses))
        rec["val"].append(to_np(L.val_losses))
        rec["scaled_train"].append(to_np(L.scaled_train_losses))
        rec["total_train"].append(float(L.total_train_loss))
        rec["total_val"].append(float(L.total_val_loss))
        rec["lr"].append(float(new_lr))
        rec["mask_idx"].append(int(new_mask_idx))
        rec["grad_dot"].append(float(grad_dot))
        g = new_state.gradients
        g_hist.append(to_np(g.frame_weights).copy())
        mpg = g.model_parameters[0]
        rec["g_bc"].append(float(to_np(mpg.bv_bc).reshape(-1)[0]))
        rec["g_bh"].append(float(to_np(mpg.bv_bh).reshape(-1)[0]))
        mpn = new_state.params.model_parameters[0]
        rec["bc"].append(float(to_np(mpn.bv_bc).reshape(-1)[0]))
        rec["bh"].append(float(to_np(mpn.bv_bh).reshape(-1)[0]))
        w_hist.append(to_np(save_state.params.frame_weights).copy())
        state = new_state
    for k, v in rec.items():
        out["step_" + k] = np.asarray(v)
    out["step_z_in"] = np.asarray(z_hist)
    out["step_g"] = np.asarray(g_hist)
    out["step_w_out"] = np.asarray(w_hist)
    out["z_final"] = to_np(state.params.frame_weights)
    return out


def run_loops(name, spec, x64):
    """_optimise (Python loop) and _optimise_pure (the batch path's while_loop body) on the same case"""
    from jaxent.src.opt.run import _optimise, _optimise_pure

    out = {}
    for which in ("loop", "pure"):
        case = build_case(spec)
        rr = RefRun(case, spec, x64)
        n_steps, conv = 60, [1e-1, 1e-2, 1e-3]
        if which == "loop":
            sim, opt = _optimise(rr.sim, rr.targets, n_steps, 1e-10, conv, rr.indexes, rr.loss_fns, rr.state, rr.optimizer, ema_alpha=0.5,
                                 min_steps_per_threshold=2, silent=True)
            hist = opt.history
            out["loop_n_states"] = np.asarray(len(hist.states))
            out["loop_state_steps"] = np.asarray([int(s.step) for s in hist.states])
            out["loop_state_total_train"] = np.asarray([float(s.losses.total_train_loss) for s in hist.states])
            out["loop_state_w"] = np.asarray([to_np(s.params.frame_weights) for s in hist.states])
            best = hist.get_best_state()
            out["loop_best_step"] = np.asarray(int(best.step))
            out["loop_final_w"] = to_np(sim.params.frame_weights)
        else:
            res = _optimise_pure(rr.sim, tuple(rr.targets), n_steps, 1e-10, conv, rr.indexes, tuple(rr.loss_fns), rr.state, rr.optimizer,
                                 ema_alpha=0.5, min_steps_per_threshold=2)
            n = int(res.write_idx)
            out["pure_write_idx"] = np.asarray(n)
            out["pure_converged"] = np.asarray(bool(res.convergence.converged))
            out["pure_threshold_idx"] = np.asarray(int(res.convergence.current_threshold_idx))
            out["pure_steps_since"] = np.asarray(int(res.convergence.steps_since_threshold_start))
            out["pure_ema_loss_delta"] = np.asarray(float(res.convergence.ema_loss_delta))
            out["pure_ema_w"] = to_np(res.convergence.ema_params.frame_weights)
            out["pure_history_total_train"] = to_np(res.history_losses.total_train_loss)
            out["pure_history_val0"] = to_np(res.history_losses.val_losses)[:, 0]
            out["pure_history_w"] = to_np(res.history_params.frame_weights)
            out["pure_z"] = to_np(res.opt_state.params.frame_weights)
            out["pure_step"] = np.asarray(int(res.opt_state.step))
            # the selection batch_optimise makes from this carry (opt/batch.py:158-180 -> OptimizationHistory._pick_best_state)
            import jax
            from jaxent.src.opt.base import OptimizationHistory, OptimizationState

            states = []
            for i in range(n):
                states.append(OptimizationState(params=jax.tree_util.tree_map(lambda x: x[i], res.history_params), opt_state=res.opt_state.opt_state,
                                                step=i, losses=jax.tree_util.tree_map(lambda x: x[i], res.history_losses),
                                                gradients=res.opt_state.gradients))
            best = OptimizationHistory(states=states).get_best_state()
            out["pure_best_index"] = np.asarray(int(best.step))
    return out


BATCH = dict(fw=[[0.5, 0.5, 0.3], [0.8, 0.2, 0.1], [0.3, 0.7, 0.5]], fs=[[1000.0, 1000.0, 100.0], [500.0, 1000.0, 1000.0], [1000.0, 200.0, 10.0]],
             lr=[1.0, 0.5, 0.2], n_steps=40, convergence=[1.0, 0.3], batch_size=2)  # the first n_slots columns of fw / fs are used


def run_batch(name, spec, x64):
    """the reference's batch_optimise (opt/batch.py:51-192: jax.vmap of _optimise_pure over (fw, fs, lr), lax.map over the
    batches, padding of the last batch) on one case: what every replica ends with, and the selection it returns"""
    case = build_case(spec)
    rr = RefRun(case, spec, x64)  # installs the stand-ins: the reference imports below need them
    from jaxent.src.custom_types.config import OptimiserSettings
    from jaxent.src.opt.base import HParamBatch
    from jaxent.src.opt.batch import batch_optimise

    jnp = rr.jnp
    fdt = jnp.float64 if x64 else jnp.float32
    cfg = OptimiserSettings(name="batch", n_steps=BATCH["n_steps"], tolerance=1e-10, convergence=list(BATCH["convergence"]), learning_rate=spec["lr"])
    n_slots = len(case["problem"].slots)
    fw_b, fs_b = np.asarray(BATCH["fw"])[:, :n_slots], np.asarray(BATCH["fs"])[:, :n_slots]
    hp = HParamBatch(jnp.asarray(fw_b, dtype=fdt), jnp.asarray(fs_b, dtype=fdt), jnp.asarray(BATCH["lr"], dtype=jnp.float32))
    res = batch_optimise(rr.sim, hp, BATCH["batch_size"], tuple(rr.targets), cfg, list(rr.indexes), list(rr.loss_fns), optimizer=rr.optimizer)
    n_h, n_steps = len(BATCH["lr"]), BATCH["n_steps"]
    tot = np.full((n_h, n_steps), np.nan)
    val0 = np.full((n_h, n_steps), np.nan)
    for b, h in enumerate(res.histories):
        for i, st in enumerate(h.states):
            tot[b, i] = float(st.losses.total_train_loss)
            val0[b, i] = float(to_np(st.losses.val_losses).reshape(-1)[0])
    return {
        "batch_conv_steps": to_np(res.convergence_steps).astype(np.int64),
        "batch_n_states": np.asarray([len(h.states) for h in res.histories]),
        "batch_best_step": np.asarray([int(b.step) for b in res.best_states]),
        "batch_best_total_train": np.asarray([float(b.losses.total_train_loss) for b in res.best_states]),
        "batch_best_w": np.asarray([to_np(b.params.frame_weights) for b in res.best_states]),
        "batch_last_w": np.asarray([to_np(h.states[-1].params.frame_weights) for h in res.histories]),
        "batch_best_bc": np.asarray([float(to_np(b.params.model_parameters[0].bv_bc).reshape(-1)[0]) for b in res.best_states]),
        "batch_hist_total_train": tot, "batch_hist_val0": val0,
        "batch_fw": fw_b, "batch_fs": fs_b, "batch_lr": np.asarray(BATCH["lr"]),
    }


def main():
    if not refload.available():
        raise SystemExit("reference sources not available: fixtures can only be generated in the build container")
    only = set(sys.argv[1:])  # optional: names of the cases to (re)generate ("batch": the batch_optimise fixtures)
    if not only or "batch" in only:
        for name in ("uptake_eye", "pf_l2"):
            for x64 in (False, True):
                out = run_batch(name, CASES[name], x64)
                path = os.path.join(HERE, f"ref_batch_{name}_{'f64' if x64 else 'f32'}.npz")
                np.savez_compressed(path, **out)
                print(f"{path}: iterations {out['batch_conv_steps']}, best entries {out['batch_best_step']}")
    for name, spec in CASES.items():
        if only and name not in only:
            continue
        for x64 in (False, True):
            out = run_case(name, spec, x64)
            if name in ("uptake_eye", "pf_l2"):
                out.update(run_loops(name, spec, x64))
            tag = "f64" if x64 else "f32"
            path = os.path.join(HERE, f"ref_{name}_{tag}.npz")
            spec_arr = {"spec_" + k: np.asarray(v if v is not None else -1.0) for k, v in spec.items() if not isinstance(v, (tuple, str))}
            np.savez_compressed(path, **out, **spec_arr)
            print(f"{path}: {len(out)} arrays, total_train[0..2] = {out['step_total_train'][:3]}")


if __name__ == "__main__":
    main()
