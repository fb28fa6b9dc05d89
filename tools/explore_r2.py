"""Round-2 exploration on one B200 (not a test): (a) how far lr = 1 / scaling = 1000 trajectories of the GPU, the fp32
mirror and the fp64 oracle drift apart over 1000 steps; (b) whether the fused peer exchange can be driven by two
processes that share ONE GPU (gloo for the set-up collectives, CUDA IPC for the inboxes)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
from oracle import bv_oracle as O  # noqa: E402


def drift(F, R, T, kind, lr, scaling, n=1000, init_steps=0):
    case = H.make_case(F, R, T, kind, seed=21, scaling=scaling)
    cfg = O.OptConfig(learning_rate=lr, initial_learning_rate=1.0, initial_steps=init_steps, clip_value=None)
    eng = H.make_engine(case, cfg)
    H.init_engine(eng, case)
    eng.step(n)
    torch.cuda.synchronize()
    wg = eng.weights.cpu().numpy().astype(np.float64)
    res = {}
    for name, dt in (("f32", np.float32), ("f64", np.float64)):
        st = O.optimizer_initialise(case["params"], cfg, dt)
        for _ in range(n):
            st, loss, w = O.step_with_rates(case["problem"], st, cfg, dt)
        res[name] = (np.asarray(w, np.float64), float(loss))
    rec = eng.records(n - 1, 1)[0]
    print(f"drift F={F} R={R} T={T} kind={kind} lr={lr} scaling={scaling}: |gpu-f64|={np.abs(wg - res['f64'][0]).max():.3e} "
          f"|gpu-f32|={np.abs(wg - res['f32'][0]).max():.3e} |f32-f64|={np.abs(res['f32'][0] - res['f64'][0]).max():.3e} "
          f"wmax={wg.max():.3e} loss gpu={rec.total_train:.6g} f32={res['f32'][1]:.6g} f64={res['f64'][1]:.6g}", flush=True)


def _same_gpu_worker(rank, world, port, out):
    import torch.distributed as dist

    from jaxent_b200.sharding import frame_shard

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        F, R, T, n_steps = 5003, 97, 5, 6
        case = H.make_case(F, R, T, O.UPTAKE_EYE_MSE, extra=(O.PARAMS_L2,), seed=6, prior="random")
        cfg = O.OptConfig(learning_rate=1.0, clip_value=None, optimise_model_parameters=True)
        sh = frame_shard(F, rank, world)
        t0 = time.time()
        eng = H.make_engine(case, cfg, device="cuda:0", frames=sh.slice(), F_total=F, group=dist.group.WORLD, exchange="nvlink")
        H.init_engine(eng, case, frames=sh.slice())
        eng.step(n_steps)
        eng.finish_record()
        torch.cuda.synchronize()
        dt = time.time() - t0
        err = eng.exchange_error()
        recs = eng.records(0, n_steps + 1)
        print(f"rank {rank}: exchange={eng.exchange} err={err} {dt:.2f}s losses={[r.total_train for r in recs]}", flush=True)
        if rank == 0:
            ref = H.make_engine(case, cfg, device="cuda:0")
            H.init_engine(ref, case)
            ref.step(n_steps)
            ref.finish_record()
            torch.cuda.synchronize()
            print("single   :", [r.total_train for r in ref.records(0, n_steps + 1)], flush=True)
        dist.barrier()
        eng.close()
        out["ok"] = 1
    finally:
        dist.destroy_process_group()


def same_gpu():
    import socket

    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = mp.get_context("spawn").Manager().dict()
    mp.spawn(_same_gpu_worker, args=(2, port, out), nprocs=2, join=True)
    print("same-gpu sharded run:", dict(out))


def collapsed():
    from test_gpu_parity import _collapsed_case

    for F, R, T, kind, spread, om in ((2000, 40, 4, O.UPTAKE_EYE_MSE, 10.0, False), (2000, 40, 4, O.UPTAKE_EYE_MSE, 30.0, False),
                                      (2000, 40, 1, O.PF_L2, 10.0, True), (2000, 40, 1, O.PF_L2, 30.0, True),
                                      (20003, 500, 8, O.UPTAKE_EYE_MSE, 30.0, False), (20003, 500, 8, O.UPTAKE_EYE_MSE, 10.0, False)):
        case, frac = _collapsed_case(F, R, T, kind, spread, seed=13 if F == 2000 else 5, extra=(O.PARAMS_L2,) if om else ())
        cfg = O.OptConfig(learning_rate=1.0, clip_value=None, optimise_model_parameters=om)
        eng = H.make_engine(case, cfg)
        H.init_engine(eng, case)
        p = case["params"]
        for k in range(6):
            pre, sc = eng.export_state(("z",))
            z = pre["z"].cpu().numpy()
            eng.step(1)
            torch.cuda.synchronize()
            rec = eng.records(k, 1)[0]
            gg = eng.export_state(("g",))[0]["g"].cpu().numpy()
            par = O.Params(z.astype(np.float64), p.frame_mask, sc["bc"], sc["bh"], p.fw, p.fs, p.nlf)
            l, g, aux = O.loss_and_grads(case["problem"], par, np.float64)
            l32, g32, _ = O.loss_and_grads(case["problem"], par, np.float32)
            sc_ = np.abs(g.z).max()
            print(f"F={F} R={R} {kind} spread={spread} step {k}: wmax={aux['w'].max():.4f} collapsed={float((aux['w'] < 1e-10 / F).mean()):.2f} "
                  f"gerr={np.abs(gg - g.z).max() / sc_:.2e} mirror32={np.abs(g32.z - g.z).max() / sc_:.2e} "
                  f"lerr={abs(rec.total_train - l.total_train_loss) / abs(l.total_train_loss):.1e} hbar_kl={aux['hbar_kl']:.1f}", flush=True)


def batched_dbg():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_batched import _batched, _hparams
    from test_gpu_parity import _collapsed_case

    F, R, T, B = 2000, 40, 4, 16
    for spread in (10.0, 30.0):
        case, frac = _collapsed_case(F, R, T, O.UPTAKE_EYE_MSE, spread)
        n = len(case["problem"].slots)
        fw, fs, lr = _hparams(B, n, 3)
        fs *= 10.0
        cfg = O.OptConfig(learning_rate=1.0, clip_value=None)
        eng = _batched(case, cfg, B)
        p = case["params"]
        z0 = np.asarray(p.z, np.float32) * np.float32(F)
        eng.init_state(z0, p.bc, p.bh, fw, fs, lr)
        eng.step(1)
        torch.cuda.synchronize()
        G = eng.gradients_all().cpu().numpy()
        recs = eng.records_at(0)
        for b in range(0, B, 5):
            par = O.Params(z0.astype(np.float64), p.frame_mask, p.bc, p.bh, fw[b], fs[b], p.nlf)
            l, g, aux = O.loss_and_grads(case["problem"], par, np.float64)
            print(f"batched spread={spread} b={b}: train gpu={list(recs[b].train[:n])} oracle={list(l.train_losses)} "
                  f"gerr={np.abs(G[:, b] - g.z).max() / np.abs(g.z).max():.2e}", flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "drift"
    if what == "drift":
        for lr, sc in ((1.0, 1000.0), (1.0, 100.0), (0.05, 100.0)):
            drift(500, 53, 3, O.UPTAKE_EYE_MSE, lr, sc)
            drift(1000, 130, 1, O.PF_L2, lr, sc)
    elif what == "batched":
        batched_dbg()
    elif what == "collapsed":
        collapsed()
    else:
        same_gpu()
