"""Frame-sharded run against the single-GPU run of the same problem and against the oracle.

Two set-ups: (a) >= 2 GPUs, one rank per GPU over NCCL (skipped on boxes with one GPU); (b) two ranks that SHARE cuda:0
(gloo for the set-up collectives, CUDA IPC for the inboxes of the fused exchange) -- the same kernels, stamped-word protocol
and rank-ordered sums as (a), runnable on a one-GPU box; the two contexts are time-sliced, so a waiting kernel spins until
the peer's slice arrives (slow, but bounded).  The protocol itself is also covered on CPU by test_sharding_gloo.py."""
import os
import socket

import numpy as np
import pytest
import torch

import helpers as H
from oracle import bv_oracle as O

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, F, R, T, n_steps, exchange, out, per_frame=False, one_gpu=False, spread=0.0):
    import torch.distributed as dist

    from jaxent_b200.sharding import frame_shard

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = 0 if one_gpu else rank
    torch.cuda.set_device(dev)
    if one_gpu:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        case = H.make_case(F, R, T, O.UPTAKE_EYE_MSE, extra=(O.PARAMS_L2,), seed=6, prior="random")
        if spread > 0:  # collapsed weights: the MaxEnt part of hbar is O(lambda) and crosses the ranks (kl words)
            z0 = np.random.default_rng(5).normal(0.0, spread, F)
            case["params"].z = (z0 / F).astype(np.float32)
        case["problem"].average_first = not per_frame
        cfg = O.OptConfig(learning_rate=1.0, clip_value=None, optimise_model_parameters=not per_frame)
        sh = frame_shard(F, rank, world)
        eng = H.make_engine(case, cfg, device=f"cuda:{dev}", frames=sh.slice(), F_total=F, group=dist.group.WORLD, exchange=exchange)
        assert eng.exchange == exchange
        H.init_engine(eng, case, frames=sh.slice())
        eng.step(n_steps)
        eng.finish_record()
        torch.cuda.synchronize()
        assert eng.exchange_error() == 0
        recs = eng.records(0, n_steps + 1)
        w = eng.weights.cpu().numpy()
        if exchange == "nvlink":
            # a second run on the same plan (flags are re-armed by state_init) and CUDA-graph replay of the fused step
            H.init_engine(eng, case, frames=sh.slice())
            eng.capture(2)
            eng.step_graph(n_steps // 2)
            eng.finish_record()
            torch.cuda.synchronize()
            assert eng.exchange_error() == 0
            recs2 = eng.records(0, n_steps + 1)
            for a, b in zip(recs, recs2):
                assert a.step == b.step and a.total_train == b.total_train and a.bc == b.bc  # bitwise
            np.testing.assert_array_equal(w, eng.weights.cpu().numpy())
        g1 = None
        if spread > 0:  # masked gradients of the FIRST update against the oracle (teacher-forced: same initial state)
            H.init_engine(eng, case, frames=sh.slice())
            eng.step(1)
            torch.cuda.synchronize()
            g1 = eng.export_state(("g",))[0]["g"].cpu().numpy()
        gathered = [None] * world
        dist.all_gather_object(gathered, (sh.start, sh.stop, w, g1))
        if rank == 0:
            wfull = np.concatenate([g[2] for g in sorted(gathered, key=lambda g: g[0])])
            # single-GPU run of the same problem on this device
            ref = H.make_engine(case, cfg, device="cuda:0")
            H.init_engine(ref, case)
            ref.step(n_steps)
            ref.finish_record()
            torch.cuda.synchronize()
            rrecs = ref.records(0, n_steps + 1)
            for a, b in zip(recs, rrecs):
                assert a.step == b.step and a.branch == b.branch
                np.testing.assert_allclose(a.total_train, b.total_train, rtol=2e-6)
                np.testing.assert_allclose(a.bc, b.bc, rtol=1e-6)
            np.testing.assert_allclose(wfull, ref.weights.cpu().numpy(), rtol=2e-4, atol=1e-9)
            assert abs(wfull.sum() - 1.0) < 1e-5
            if spread > 0:
                st0 = O.optimizer_initialise(case["params"], cfg, np.float64)
                _, g0, _ = O.loss_and_grads(case["problem"], st0.params, np.float64)
                gfull = np.concatenate([g[3] for g in sorted(gathered, key=lambda g: g[0])])
                assert np.abs(gfull - g0.z).max() / np.abs(g0.z).max() < 1e-5
            # and the oracle, step 0 (same state on both sides)
            st = O.optimizer_initialise(case["params"], cfg, np.float64)
            losses, g, aux = O.loss_and_grads(case["problem"], st.params, np.float64)
            np.testing.assert_allclose(recs[0].total_train, losses.total_train_loss, rtol=1e-5 if not per_frame else 5e-5)
            if spread > 0:
                assert abs(aux["hbar_kl"]) > 0.3 * float(st.params.fw[1] * st.params.fs[1])
            out["ok"] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("exchange", ["nvlink", "nccl"])
@pytest.mark.parametrize("world", [2])
def test_sharded_matches_single_gpu(world, exchange):
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < world:
        pytest.skip("not enough GPUs")
    out = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), 5003, 97, 5, 6, exchange, out), nprocs=world, join=True)
    assert out.get("ok") == 1


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_per_frame_uptake_matches_single_gpu():
    """the per-frame uptake mode exchanges 2*T*R + 4 doubles per step through the same fused NVLink inbox"""
    import torch.multiprocessing as mp

    out = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(2, _free_port(), 3001, 64, 4, 6, "nvlink", out, True), nprocs=2, join=True)
    assert out.get("ok") == 1


@pytest.mark.timeout(600)
@pytest.mark.parametrize("exchange,spread", [("nvlink", 0.0), ("nvlink", 10.0), ("nccl", 10.0)])
def test_sharded_two_ranks_on_one_gpu(exchange, spread):
    """Two ranks sharing cuda:0 (runs on a one-GPU box): fused peer exchange over CUDA IPC, and the collective fallback
    driven by gloo; with collapsed weights the MaxEnt part of hbar has to cross the ranks before any Adam update."""
    import torch.multiprocessing as mp

    out = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(2, _free_port(), 5003, 97, 5, 6, exchange, out, False, True, spread), nprocs=2, join=True)
    assert out.get("ok") == 1
