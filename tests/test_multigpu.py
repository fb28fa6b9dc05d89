"""Frame-sharded run on >= 2 GPUs (NCCL) against the single-GPU run of the same problem and against the
oracle.  Skipped on boxes with one GPU; the protocol itself is covered on CPU by test_sharding_gloo.py."""
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


def _worker(rank, world, port, F, R, T, n_steps, exchange, out, per_frame=False):
    import torch.distributed as dist

    from jaxent_b200.sharding import frame_shard

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        case = H.make_case(F, R, T, O.UPTAKE_EYE_MSE, extra=(O.PARAMS_L2,), seed=6, prior="random")
        case["problem"].average_first = not per_frame
        cfg = O.OptConfig(learning_rate=1.0, clip_value=None, optimise_model_parameters=not per_frame)
        sh = frame_shard(F, rank, world)
        eng = H.make_engine(case, cfg, device=f"cuda:{rank}", frames=sh.slice(), F_total=F, group=dist.group.WORLD, exchange=exchange)
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
        gathered = [None] * world
        dist.all_gather_object(gathered, (sh.start, sh.stop, w))
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
            # and the oracle, step 0 (same state on both sides)
            st = O.optimizer_initialise(case["params"], cfg, np.float64)
            losses, _, _ = O.compute_loss(case["problem"], st.params, np.float64)
            np.testing.assert_allclose(recs[0].total_train, losses.total_train_loss, rtol=1e-5)
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
