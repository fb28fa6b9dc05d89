"""CPU tests of the frame-sharded (N>1) host path: shard planning and the collective protocol,
run with world_size 2 and 3 over gloo on 127.0.0.1.  Each rank forms the partial sums the kernels
form on its shard (restated in NumPy fp64), pushes them through the product's own ShardComm, and the
result must equal the unsharded oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H
from jaxent_b200.sharding import FRAME_ALIGN, ShardComm, frame_shard
from oracle import bv_oracle as O


def test_frame_shard_partition_properties():
    for F, W in [(1_000_000, 8), (1000, 3), (17, 2), (8, 1), (1001, 4), (64, 8)]:
        shards = [frame_shard(F, r, W) for r in range(W)]
        assert shards[0].start == 0 and shards[-1].stop == F
        for a, b in zip(shards, shards[1:]):
            assert a.stop == b.start and a.start % FRAME_ALIGN == 0 and b.start % FRAME_ALIGN == 0
        assert sum(s.size for s in shards) == F and all(s.size > 0 for s in shards)
    assert [frame_shard(1_000_000, r, 8).size for r in range(8)] == [125_000] * 8
    with pytest.raises(ValueError):
        frame_shard(3, 0, 8)
    with pytest.raises(ValueError):
        frame_shard(100, 5, 4)
    with pytest.raises(ValueError):
        frame_shard(9, 0, 2, align=16)  # one 16-frame tile cannot be split: rank 0 would own nothing


def test_shardcomm_is_noop_without_process_group():
    c = ShardComm(None)
    t = torch.tensor([1.0, 2.0])
    assert not c.active and c.world == 1 and c.sum_(t) is t and c.max_(t) is t and c.calls == 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, F, R, T, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = ShardComm(dist.group.WORLD)
        assert comm.active and comm.world == world and comm.rank == rank
        case = H.make_case(F, R, T, O.UPTAKE_EYE_MSE, seed=4, prior="random")
        pb, par = case["problem"], case["params"]
        sh = frame_shard(F, rank, world)
        rng = np.random.default_rng(9)
        z = rng.normal(size=F) * 2.0  # same logits on every rank, each uses its slice
        zl = z[sh.slice()]
        Nc, Nh = pb.heavy[:, sh.slice()].astype(np.float64), pb.acceptor[:, sh.slice()].astype(np.float64)

        # initialisation collectives: global max (softmax shift) and prior normaliser
        zmax = comm.max_(torch.tensor([zl.max()], dtype=torch.float64))
        eps = 1e-10 / sh.total  # eps uses the GLOBAL frame count (opt/losses.py:308)
        prior = case["prior"].astype(np.float64)
        pnorm = comm.sum_(torch.tensor([np.sum(np.abs(prior[sh.slice()]) + eps)], dtype=torch.float64))

        # all-reduce #1: the (4R+4) partial-sum vector of the pass (both plateau branches, here identical)
        e = np.exp(zl - float(zmax))
        red = np.concatenate([Nc @ e, Nh @ e, Nc @ e, Nh @ e, [e.sum(), e.sum(), 0.0, 0.0]])
        assert red.shape[0] == 4 * R + 4
        red = comm.sum_(torch.from_numpy(red)).numpy()
        S = red[4 * R]
        a, b = red[:R] / S, red[R : 2 * R] / S

        # all-reduce #2: the MaxEnt sums of the tail (needs the completed S)
        w = e / S
        q = np.abs(w) + eps
        p = (np.abs(prior[sh.slice()]) + eps) / float(pnorm)
        kl = np.array([0.0, np.sum(p * np.log(p / q)), np.sum(q - p), w.sum()])
        kl = comm.sum_(torch.from_numpy(kl)).numpy()
        assert comm.calls == 4

        # every rank must now hold exactly what the unsharded oracle computes
        full = O.Params(z, par.frame_mask, par.bc, par.bh, par.fw, par.fs, par.nlf)
        losses, npar, outp = O.compute_loss(pb, full, np.float64)
        np.testing.assert_allclose(a, outp["avg_heavy"], rtol=1e-12)
        np.testing.assert_allclose(b, outp["avg_acceptor"], rtol=1e-12)
        np.testing.assert_allclose(kl[1] + kl[2], losses.train_losses[1], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(kl[3], 1.0, rtol=1e-12)
        np.testing.assert_allclose(w, npar.z[sh.slice()], rtol=1e-12)
        out[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,F", [(2, 1000), (3, 530)])
def test_sharded_protocol_matches_unsharded_oracle(world, F):
    port = _free_port()
    out = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, port, F, 23, 3, out), nprocs=world, join=True)
    assert sorted(out.keys()) == list(range(world))


def _worker_per_frame(rank, world, port, F, R, T, out):
    """per-frame uptake mode (average_first=False): the reduced vector is 2*T*R + 4 doubles holding sum_f e_f exp(-k tau / PF_f)
    for both branches; every rank forms <u> = 1 - acc / S from it (tail of the per-frame mode)"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = ShardComm(dist.group.WORLD)
        from dataclasses import replace

        case = H.make_case(F, R, T, O.UPTAKE_EYE_MSE, seed=6, prior="random")
        pb, par = replace(case["problem"], average_first=False), case["params"]
        sh = frame_shard(F, rank, world)
        z = np.random.default_rng(11).normal(size=F) * 2.0
        zl = z[sh.slice()]
        Nc, Nh = pb.heavy[:, sh.slice()].astype(np.float64), pb.acceptor[:, sh.slice()].astype(np.float64)
        zmax = comm.max_(torch.tensor([zl.max()], dtype=torch.float64))
        e = np.exp(zl - float(zmax))
        pinv = np.exp(-(par.bc * Nc + par.bh * Nh))  # (R, Fl)
        k, tp = pb.k_ints.astype(np.float64), pb.timepoints.astype(np.float64)
        E = np.exp(-k[None, :, None] * tp[:, None, None] * pinv[None, :, :])  # (T, R, Fl)
        acc = (E * e[None, None, :]).sum(-1).reshape(-1)
        red = np.concatenate([acc, acc, [e.sum(), e.sum(), 0.0, 0.0]])
        assert red.shape[0] == 2 * T * R + 4
        red = comm.sum_(torch.from_numpy(red)).numpy()
        S = red[2 * T * R]
        u = 1.0 - red[: T * R].reshape(T, R) / S
        full = O.Params(z, par.frame_mask, par.bc, par.bh, par.fw, par.fs, par.nlf)
        _, _, outp = O.compute_loss(pb, full, np.float64)
        np.testing.assert_allclose(u, outp["uptake"], rtol=1e-11, atol=1e-13)
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_sharded_per_frame_protocol_matches_unsharded_oracle():
    port = _free_port()
    out = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker_per_frame, args=(2, port, 600, 19, 4, out), nprocs=2, join=True)
    assert sorted(out.keys()) == [0, 1]
