#!/usr/bin/env python
"""bench.py -- throughput of the BV MaxEnt optimise step (BASELINE.json metric) on B200.

  python bench.py --gpus 1 --steps K --warmup W            our arm (CUDA path through the C ABI)
  torchrun ... bench.py --gpus N ...                        frame-sharded over N GPUs (NCCL)
  python bench.py --impl reference ...                      the CPU port of the reference path (oracle/)

A "step" is one optimise step (forward + backward + Adam) over the whole ensemble.  Default
workload: BASELINE config 4, synthetic 1M frames x 500 residues x 8 timepoints (4 GB of fp32
features, far larger than the 126 MB L2, so no L2 flush is needed between steps); with N GPUs
the same 1M frames are sharded (strong scaling), the per-step exchange fused into the kernels over
NVLink peer memory (NCCL all-reduces as the fallback).  The K timed steps are replayed from a CUDA
graph (two launches per step, no host synchronisation); the same K steps are then launched eagerly
with CUDA events around sampled pass kernels for the roofline.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def _pin_cpu_threads():
    """The CPU arm must use every host core it may run on.  torchrun exports OMP_NUM_THREADS=1 to its workers, which
    would silently make NumPy's BLAS single-threaded; the thread counts are therefore set explicitly BEFORE numpy /
    torch are imported, and the number actually used is read back from the BLAS library (cpu_threads_used)."""
    n = str(host_threads())
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS", "VECLIB_MAXIMUM_THREADS"):
        os.environ[v] = n


_ARGV = " ".join(sys.argv[1:])
if "--impl reference" in _ARGV or "--impl=reference" in _ARGV:
    _pin_cpu_threads()
elif int(os.environ.get("WORLD_SIZE", "1")) == 1 and os.environ.get("OMP_NUM_THREADS") in (None, "1"):
    _pin_cpu_threads()  # the cpu_baseline leg of a single-GPU run

import numpy as np


def cpu_threads_used():
    """threads the BLAS behind numpy really uses (threadpoolctl), else what the environment asks for"""
    try:
        from threadpoolctl import threadpool_info

        n = [int(i["num_threads"]) for i in threadpool_info() if i.get("user_api") == "blas"]
        if n:
            return max(n)
    except Exception:
        pass
    try:
        return int(os.environ.get("OMP_NUM_THREADS", host_threads()))
    except ValueError:
        return host_threads()

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "cfg3": dict(F=100_000, R=500, T=8, name="synthetic 100k frames x 500 residues x 8 timepoints (BASELINE config 3)"),
    "cfg4": dict(F=1_000_000, R=500, T=8, name="synthetic 1M frames x 500 residues x 8 timepoints (BASELINE config 4)"),
    "cfg1": dict(F=500, R=53, T=3, name="BPTI-shaped 500 frames x 53 residues x 3 timepoints (BASELINE config 1 shape)"),
    "cfg2": dict(F=1000, R=130, T=1, name="aSyn-shaped 1000 frames x 130 residues (BASELINE config 2 shape)"),
    "cfg5": dict(F=100_000, R=500, T=8, B=256, name="batched sweep: 256 replicas (MaxEnt strengths) on 100k frames x 500 residues x 8 timepoints, tensor-core GEMM path (BASELINE config 5)"),
}
N_BLOCKS = 8  # features are generated in 8 frame blocks so that shards are identical for N = 1, 2, 4, 8


def tensor_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["bf16_tflops_sustained"]), "measured sustained bf16 (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 1500.0, "fallback (B200_PROFILING.md)"


def run_batched(args, wl):
    """BASELINE config 5: B replicas advanced together; the two feature sweeps are tcgen05 GEMMs (single GPU)."""
    if int(os.environ.get("WORLD_SIZE", 1)) > 1:
        if int(os.environ.get("RANK", 0)) == 0:
            sys.stderr.write("config 5 is a single-GPU workload (replicas only): run with --gpus 1\n")
        return
    print(json.dumps(batched_measure(args, wl, args.steps, args.warmup, full=True)), flush=True)


def batched_summary(args, dev):
    """compact config-5 result for the `extra` block of the default run"""
    line = batched_measure(args, WORKLOADS["cfg5"], 10, 3, full=False)
    r = line["roofline"]
    return {"value": line["value"], "unit": line["unit"], "ms_per_step": line["ms_per_step"], "replicas": WORKLOADS["cfg5"]["B"],
            "tensor_roofline_frac": r["frac"], "tensor_executed_frac": r["executed_frac"], "gemm1_ms": r["gemm1_ms"], "gemm2_ms": r["gemm2_ms"],
            "hbm_floor_ms": line["hbm_floor_ms"], "check": line["check"]}


def batched_measure(args, wl, K, W, full=True):
    import torch

    from jaxent_b200 import _lib as L
    from jaxent_b200.batch_engine import BatchedBvEngine
    from jaxent_b200.engine import BvEngine, OptSettings, SlotSpec

    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    F, R, T, B = wl["F"], wl["R"], wl["T"], wl["B"]
    NS = args.pieces
    small = small_problem(R, T)
    heavy, acc, _, _ = gen_features(R, F, 0, 1, dev)
    slots = [SlotSpec(L.LOSS_UPTAKE_EYE_MSE, small.S_train, small.y_train, small.S_val, small.y_val), SlotSpec(L.LOSS_MAXENT_CONVEX_KL)]
    settings = OptSettings(learning_rate=1.0, initial_learning_rate=1.0, initial_steps=0, clip_value=None)
    prior = torch.full((F,), 1.0 / F, device=dev)
    gam = np.logspace(0, 3, B).astype(np.float32)  # MaxEnt strengths: fw = [gamma, 1] normalised (examples/common/optimization.py:169-188)
    fw = np.stack([gam / (gam + 1), 1 / (gam + 1)], 1).astype(np.float32)
    fs = np.full((B, 2), 1000.0, np.float32)
    host_h = host_a = None
    if full:
        host_h = torch.empty((R, F), dtype=torch.float32, pin_memory=True)
        host_a = torch.empty((R, F), dtype=torch.float32, pin_memory=True)
        host_h.copy_(heavy)
        host_a.copy_(acc)
    eng = BatchedBvEngine(heavy, acc, small.k_ints, small.timepoints, slots, True, settings, B, prior=prior, device=dev, n_pieces=NS)
    eng.init_state(torch.ones(F, device=dev), 0.35, 2.0, fw, fs, None)
    eng.step(max(3, W))
    torch.cuda.synchronize()
    sp = torch.cuda.current_stream().cuda_stream
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(0)
    sampler.start()
    torch.cuda.synchronize()
    e0.record()
    for k in range(K):
        ev[k][0].record()
        L.check(eng.lib.jaxent_bg_step_phase(eng.plan, 0, sp))
        ev[k][1].record()
        L.check(eng.lib.jaxent_bg_step_phase(eng.plan, 1, sp))
        ev[k][2].record()
        L.check(eng.lib.jaxent_bg_step_phase(eng.plan, 2, sp))
        ev[k][3].record()
        L.check(eng.lib.jaxent_bg_step_phase(eng.plan, 3, sp))
    e1.record()
    torch.cuda.synchronize()
    sampler.stop_flag = True
    sampler.join()
    eng.steps_done += K
    ms_step = e0.elapsed_time(e1) / K
    g1_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    g2_ms = float(np.mean([e[2].elapsed_time(e[3]) for e in ev]))
    rec = eng.records_at(eng.steps_done)
    wsum = float(eng.weights_all().double().sum(0).mean())

    # e2e: whole sweep from pinned host features through the engine API, per-step loss records read back
    import ctypes

    del eng
    torch.cuda.empty_cache()
    e2e = seq = None
    if full:
        e2e, seq = _batched_e2e_and_seq(args, wl, dev, host_h, host_a, small, slots, settings, prior, fw, fs, K, NS)
    peak, peak_src = tensor_peak()
    alg_flops = 4 * R * F * B  # per GEMM launch: 2 flops x (2 arrays x R) x F x B
    gemm_ms = 0.5 * (g1_ms + g2_ms)
    achieved = alg_flops / (gemm_ms * 1e-3) / 1e12
    hbm_peak, _ = peaks()
    line = {
        "metric": "optimise steps/sec", "value": B * 1e3 / ms_step, "unit": "replica-steps/s", "n_gpus": 1, "steps": K, "warmup": max(3, W),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16 x%d pieces, fp32 accumulate" % NS,
        "data": "synthetic", "batched_steps_per_s": 1e3 / ms_step, "frame_residue_timepoint_per_s": B * 1e3 / ms_step * F * R * T,
        "config": {"workload": wl["name"], "frames": F, "residues": R, "timepoints": T, "replicas": B, "bf16_pieces": NS,
                   "parallelism": "single GPU, replicas only", "l2": "state arrays 6 x 102 MB + features 205 MB per step >> 126 MB L2, no flush needed",
                   "losses": ["hdx_uptake_eye_MSE_loss", "maxent_convexKL_loss"], "optimizer": "adam lr=1.0, plateau 1.005"},
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                     "kernel": "bg_gemm_kernel<1>/<2> (mean of the backward and forward product)", "algorithmic_flops_per_launch": alg_flops,
                     "kernel_ms": gemm_ms, "gemm1_ms": g1_ms, "gemm2_ms": g2_ms, "executed_tflops": NS * achieved,
                     "executed_frac": NS * achieved / peak, "peak_source": peak_src, "step_frac": 2 * gemm_ms / ms_step,
                     "note": "achieved counts the algorithmic 2*(2R)*F*B flops per product; the fp32 operand is split into bf16 pieces, so the tensor "
                             "pipe executes `bf16_pieces` times that (executed_tflops)"},
        "hbm_floor_ms": (F * (8 * R + 28 * B)) / hbm_peak / 1e6,
        "cpu_baseline": None, "e2e": e2e, "gpu_launches": 9 * K, "clocks": sampler.result(),
        "sequential_replica_steps_per_s": seq, "speedup_vs_sequential_engine": (B * 1e3 / ms_step / seq) if seq else None,
        "check": {"mean_sum_weights": wsum, "final_loss_replica0": float(rec[0].total_train), "final_step": int(rec[0].step)},
    }
    return line


def _batched_e2e_and_seq(args, wl, dev, host_h, host_a, small, slots, settings, prior, fw, fs, K, NS):
    import ctypes

    import torch

    from jaxent_b200 import _lib as L
    from jaxent_b200.batch_engine import BatchedBvEngine
    from jaxent_b200.engine import BvEngine

    F, R, T, B = wl["F"], wl["R"], wl["T"], wl["B"]
    t0 = time.perf_counter()
    eng2 = BatchedBvEngine(host_h, host_a, small.k_ints, small.timepoints, slots, True, settings, B, prior=prior, device=dev, n_pieces=NS)
    eng2.init_state(torch.ones(F, device=dev), 0.35, 2.0, fw, fs, None)
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t0
    t1 = time.perf_counter()
    sz = ctypes.sizeof(L.LossRecord)
    pinned = torch.empty(B * sz, dtype=torch.uint8, pin_memory=True)
    for k in range(K):
        eng2.step(1)
        row = ((k + 1) % 256) * eng2.Bn * sz
        pinned.copy_(eng2._brecords[row : row + B * sz], non_blocking=True)  # every replica's loss record of this step
        torch.cuda.current_stream().synchronize()
    last = [L.LossRecord.from_buffer_copy(pinned.numpy()[:sz].tobytes())]
    w_host = eng2.weights_all().cpu()
    t_steps = time.perf_counter() - t1
    rec_bytes = ctypes.sizeof(L.LossRecord) * B
    e2e = {"value": B * K / (t_setup + t_steps), "unit": "replica-steps/s", "h2d_bytes_per_step": (2 * R * F * 4 + F * 4) / K,
           "d2h_bytes_per_step": rec_bytes + F * B * 4 / K, "steady_value": B * K / t_steps, "setup_s": t_setup,
           "note": "whole K-step sweep of all replicas from pinned host features (upload + bf16 pack + init inside the timed region), every replica's loss "
                   "record read back and host-synchronised every step, final weights of all replicas read back",
           "last_loss_replica0": float(last[0].total_train)}
    del eng2, w_host

    # sequential baseline on the same GPU: the single-ensemble engine, one replica after the other (measured on 4 replicas)
    seq = None
    try:
        heavy, acc, _, _ = gen_features(R, F, 0, 1, dev)
        one = BvEngine(heavy, acc, small.k_ints, small.timepoints, slots, True, settings, prior=prior, device=dev)
        del heavy, acc
        one.init_state(torch.ones(F, device=dev), 0.35, 2.0, fw[0], fs[0])
        one.step(5)
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        one.step(50)
        s1.record()
        torch.cuda.synchronize()
        seq = 50 * 1e3 / s0.elapsed_time(s1)
        del one
    except Exception as ex:  # pragma: no cover
        sys.stderr.write(f"sequential comparison skipped: {ex}\n")

    return e2e, seq


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def small_problem(R, T, seed=0):
    """peptide maps / targets / rates (tiny; shared by all ranks and by the CPU arm)."""
    from jaxent_b200 import synthetic

    return synthetic.make_ensemble(2048, R, T, seed=seed)


def row_hi(R, seed=0):
    rng = np.random.default_rng(seed)
    lam = rng.uniform(5, 40, R)
    mu = rng.uniform(0.1, 1.5, R)
    return np.rint(2 * lam).astype(np.int64), np.rint(2 * mu).astype(np.int64)


def gen_block_torch(R, f0, f1, blk, device, seed=0):
    """integer contact counts U{0..hi_r} for frames [f0,f1) of block `blk` (seeded per block)."""
    import torch

    hi_h, hi_a = row_hi(R, seed)
    g = torch.Generator(device=device)
    g.manual_seed(1000 * seed + blk)
    n = f1 - f0
    hh = torch.tensor(hi_h + 1, device=device, dtype=torch.float32)[:, None]
    ha = torch.tensor(hi_a + 1, device=device, dtype=torch.float32)[:, None]
    heavy = torch.floor(torch.rand((R, n), device=device, generator=g) * hh).clamp_(max=255)
    acc = torch.floor(torch.rand((R, n), device=device, generator=g) * ha).clamp_(max=255)
    return heavy, acc


def gen_features(R, F, rank, world, device, seed=0):
    """this rank's (R, F/world) shard, assembled from the global frame blocks."""
    import torch

    edges = [F * b // N_BLOCKS for b in range(N_BLOCKS + 1)]
    per = N_BLOCKS // world
    blocks = range(rank * per, (rank + 1) * per)
    hs, as_ = [], []
    for b in blocks:
        h, a = gen_block_torch(R, edges[b], edges[b + 1], b, device, seed)
        hs.append(h)
        as_.append(a)
    lo, hi = edges[rank * per], edges[(rank + 1) * per]
    if len(hs) == 1:
        return hs[0], as_[0], lo, hi
    return torch.cat(hs, 1), torch.cat(as_, 1), lo, hi


class ClockSampler(threading.Thread):
    """samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.stop_flag = [], set(), False
        self.max_mhz = None
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = str(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def result(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------- CPU arm
PF_CPU_FRAMES = 50_000  # per-frame mode: the oracle materialises (T, R, F) arrays, so the CPU arms time a frame sample


def cpu_steps(F, R, T, heavy, acc, small, budget_s, max_steps, warm=1, average_first=True):
    """times the oracle port (NumPy fp32, multi-threaded BLAS) on the full-size workload; per-frame mode (average_first False):
    on the first PF_CPU_FRAMES frames, steps/s scaled by the frame ratio (the cost is linear in the number of frames)."""
    from oracle import bv_oracle as O

    scale = 1.0
    if not average_first and F > PF_CPU_FRAMES:
        scale = PF_CPU_FRAMES / F
        heavy, acc, F = np.ascontiguousarray(heavy[:, :PF_CPU_FRAMES]), np.ascontiguousarray(acc[:, :PF_CPU_FRAMES]), PF_CPU_FRAMES

    slots = [O.LossSlot(O.UPTAKE_EYE_MSE, small.S_train, small.y_train, small.S_val, small.y_val),
             O.LossSlot(O.MAXENT_CONVEX_KL, prior=np.full(F, 1.0 / F, np.float32))]
    pb = O.Problem(heavy, acc, small.k_ints, small.timepoints, slots, True, average_first)
    fw = O.normalize_masked_loss_scalingweights(np.array([1.0, 1.0]), np.ones(2))
    par = O.Params(np.full(F, 1.0 / F, np.float32), np.ones(F, np.float32), 0.35, 2.0, fw, np.full(2, 1000.0, np.float32), np.ones(2))
    cfg = O.OptConfig(learning_rate=1.0, clip_value=None)
    st = O.optimizer_initialise(par, cfg, np.float32)
    for _ in range(warm):
        st, _, _ = O.step_with_rates(pb, st, cfg, np.float32)
    t0 = time.perf_counter()
    n = 0
    while n < max_steps and (n < 2 or time.perf_counter() - t0 < budget_s):
        st, loss, _ = O.step_with_rates(pb, st, cfg, np.float32)
        n += 1
    dt = time.perf_counter() - t0
    return scale * n / dt, n, dt, float(loss)


def host_features(F, R, seed=0):
    """full (R,F) fp32 arrays on the host: from the GPU generator when a GPU is there (same data as our arm)."""
    try:
        import torch

        if torch.cuda.is_available():
            dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
            h, a, _, _ = gen_features(R, F, 0, 1, dev, seed)
            return h.cpu().numpy(), a.cpu().numpy()
    except Exception:
        pass
    from jaxent_b200 import synthetic

    hi_h, hi_a = row_hi(R, seed)
    return (synthetic.make_counts(R, F, hi_h / 2.0, seed * 7919 + 11, host_threads()),
            synthetic.make_counts(R, F, hi_a / 2.0, seed * 7919 + 13, host_threads()))


def workload_config(args, wl):
    """`config` of the JSON line: the workload only, so that both arms (ours / --impl reference) print the SAME dict for the
    same command line; everything specific to how an arm ran goes into `run_details`."""
    F, R, T = wl["F"], wl["R"], wl["T"]
    n = args.gpus
    return {
        "workload": wl["name"], "frames": F, "residues": R, "timepoints": T,
        "parallelism": f"frame-sharded x{n} (strong scaling: the same {F} frames)" if n > 1 else "single GPU",
        "forward_mode": "per-frame uptake (average_first=False)" if args.per_frame else "average_first=True",
        "losses": ["hdx_uptake_eye_MSE_loss", "maxent_convexKL_loss"],
        "optimizer": "adam lr=1.0, forward_model_scaling=1000, plateau 1.005, no clipping",
        "l2": (f"inputs {8 * R * F / n / 1e9:.2f} GB per GPU per step >> 126 MB L2, no flush needed" if 8 * R * F / n > 1.26e8
               else "inputs fit in L2 (latency-bound config; steps/s only)"),
    }


def _try_real_reference(args, wl, heavy, acc, small, W, K):
    """if a real jax + optax AND the reference package are importable on this box, time the
    reference's own loop (`OptaxOptimizer.step` jitted, JAX_PLATFORMS=cpu).  Returns (steps/s, note) or (None, why not)."""
    os.environ.setdefault("JAX_PLATFORMS", "cpu")
    try:
        import jax  # noqa: F401
        import optax  # noqa: F401

        if str(getattr(jax, "__version__", "")).endswith("refstub"):
            return None, "only the test stand-in for jax is importable"
    except Exception as ex:
        return None, f"import jax, optax failed ({type(ex).__name__})"
    try:
        for cand in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
            if os.path.isdir(os.path.join(cand, "jaxent")) and cand not in sys.path:
                sys.path.insert(0, cand)
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        import make_ref_fixtures as M  # builds the reference's Simulation / dataloaders / OptaxOptimizer for a case

        case = M.case_from_arrays(heavy, acc, small)
        rr = M.RefRun(case, dict(gamma=1.0, opt_model=False, clip=None, lr=1.0, init_lr=1.0, init_steps=0, spread=0.0), x64=False, real_jax=True)
        state, sim, opt = rr.state, rr.sim, rr.optimizer
        step = jax.jit(opt._step, static_argnames=("loss_functions", "indexes"))
        for _ in range(W):
            state, loss, _, sim = step(opt, state, sim, tuple(rr.targets), tuple(rr.loss_fns), rr.indexes)
        jax.block_until_ready(loss)
        t0 = time.perf_counter()
        for _ in range(K):
            state, loss, _, sim = step(opt, state, sim, tuple(rr.targets), tuple(rr.loss_fns), rr.indexes)
        jax.block_until_ready(loss)
        return K / (time.perf_counter() - t0), "reference OptaxOptimizer._step under jax.jit, JAX CPU backend"
    except Exception as ex:
        return None, f"jax is importable but the reference loop did not run ({type(ex).__name__}: {ex})"


def run_reference(args, wl):
    """The CPU arm: rank 0 alone (other ranks exit 0 without work), every host thread this process may use, exactly W warm-up
    and K timed full-size steps of the same workload / config / metric as our arm."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    F, R, T = wl["F"], wl["R"], wl["T"]
    W, K = max(3, args.warmup), max(1, args.steps)
    small = small_problem(R, T)
    heavy, acc = host_features(F, R)
    sps, how = _try_real_reference(args, wl, heavy, acc, small, W, K)
    kind = "reference"
    frames_note = "full-size optimise steps"
    if sps is None:
        kind = "port"
        sps, n, dt, loss = cpu_steps(F, R, T, heavy, acc, small, budget_s=240.0, max_steps=K, warm=W, average_first=not args.per_frame)
        K = n  # (only differs from --steps when K full-size CPU steps would take longer than four minutes)
        if args.per_frame and F > PF_CPU_FRAMES:
            frames_note = f"optimise steps on the first {PF_CPU_FRAMES} frames (per-frame mode; steps/s scaled by {PF_CPU_FRAMES}/{F})"
        how = (f"NumPy fp32 oracle port (oracle/bv_oracle.py), {dt:.1f} s; the reference's JAX loop was tried first: {how}")
    else:
        loss = float("nan")
    threads = cpu_threads_used()
    line = {
        "impl": "reference",
        "metric": "optimise steps/sec",
        "value": sps,
        "unit": "steps/s",
        "n_gpus": args.gpus,
        "steps": K,
        "warmup": W,
        "ms_per_step": 1e3 / sps,
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "frame_residue_timepoint_per_s": sps * F * R * T,
        "config": workload_config(args, wl),
        "cpu_baseline": {"value": sps, "unit": "steps/s", "cores": threads, "kind": kind, "sample": f"{K} {frames_note}: {how}",
                         "host_threads_available": host_threads(), "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS")},
        "e2e": {"value": sps, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "final_loss": loss,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- our arm
def eager_rate(eng, K, W, EV=8):
    """W warm-up + K timed eager steps of a single-GPU engine (two PDL-chained launches per step); CUDA events around
    every EV-th pass kernel.  -> (ms per step, ms per pass kernel)"""
    import torch

    from jaxent_b200 import _lib as L

    sp = torch.cuda.current_stream().cuda_stream
    eng.step(W)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range((K + EV - 1) // EV)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(K):
        if k % EV == 0:
            ev[k // EV][0].record()
        L.check(eng.lib.jaxent_bv_pass(eng.plan, 1, sp))
        if k % EV == 0:
            ev[k // EV][1].record()
        L.check(eng.lib.jaxent_bv_tail(eng.plan, sp))
    e1.record()
    torch.cuda.synchronize()
    eng.steps_done += K
    eng.launches += 2 * K
    return e0.elapsed_time(e1) / K, float(np.mean([a.elapsed_time(b) for a, b in ev]))


def graph_rate(eng, K):
    import torch

    eng.capture(2)
    eng.step_graph(2)
    torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    eng.step_graph(max(1, K // 2))
    g1.record()
    torch.cuda.synchronize()
    return g0.elapsed_time(g1) / (2 * max(1, K // 2))


def extra_results(args, dev, main_eng, small, slots, settings, peak):
    """The other BASELINE configs measured in the same process (N = 1 only), so that they are driver-visible: config 3,
    config 1, the per-frame uptake mode at config 4 and the batched tensor-core sweep of config 5.  Compact dicts."""
    import torch

    from jaxent_b200 import _lib as L
    from jaxent_b200.engine import BvEngine, OptSettings, SlotSpec

    out = {}
    # ---- per-frame uptake mode on the SAME packed features as the headline run
    try:
        F, R, T = main_eng.F, main_eng.R, main_eng.T
        prior = torch.full((F,), 1.0 / F, device=dev)
        pf = BvEngine(main_eng.features, None, small.k_ints, small.timepoints, slots, True, settings, prior=prior, device=dev, average_first=False)
        pf.init_state(torch.ones(F, device=dev), 0.35, 2.0, [0.5, 0.5], [1000.0, 1000.0])
        ms, pms = eager_rate(pf, 20, 3)
        clk = 1965.0
        out["cfg4_per_frame"] = {"value": 1e3 / ms, "unit": "steps/s", "ms_per_step": ms, "pass_kernel_ms": pms,
                                 "special_function_frac": F * R * (T + 1) / (pms * 1e-3) / (16 * 148 * clk * 1e6),
                                 "note": "average_first=False: F*R*(T+1) ex2 per step against 16/clk/SM at 1965 MHz"}
        pf.close()
        del pf
    except Exception as ex:  # pragma: no cover
        out["cfg4_per_frame"] = {"error": str(ex)[:200]}
    # ---- the same mode with the BV parameters optimised too: a sweep of its own for d loss / d (bc, bh) and the forward half of
    # the pass at the updated parameters of both plateau branches (4 exponential sweeps per step instead of 1)
    try:
        import dataclasses

        st_m = dataclasses.replace(settings, optimise_model_parameters=True)
        pf = BvEngine(main_eng.features, None, small.k_ints, small.timepoints, slots, True, st_m, prior=prior, device=dev, average_first=False)
        pf.init_state(torch.ones(F, device=dev), 0.35, 2.0, [0.5, 0.5], [1000.0, 1000.0])
        ms, pms = eager_rate(pf, 10, 3)
        out["cfg4_per_frame_bv_params"] = {"value": 1e3 / ms, "unit": "steps/s", "ms_per_step": ms, "pass_kernel_ms": pms,
                                           "special_function_frac": 4 * F * R * (T + 1) / (ms * 1e-3) / (16 * 148 * clk * 1e6),
                                           "note": "average_first=False, optimise_model_parameters: 4*F*R*(T+1) ex2 per step (gradient sweep + "
                                                   "pass with the forward half re-evaluated for both plateau branches)"}
        pf.close()
        del pf
    except Exception as ex:  # pragma: no cover
        out["cfg4_per_frame_bv_params"] = {"error": str(ex)[:200]}
    return out


def extra_small_configs(args, dev, peak):
    import torch

    from jaxent_b200 import _lib as L
    from jaxent_b200.engine import BvEngine, OptSettings, SlotSpec

    out = {}
    settings = OptSettings(learning_rate=1.0, initial_learning_rate=1.0, initial_steps=0, clip_value=None)
    for key in ("cfg3", "cfg1"):
        try:
            w = WORKLOADS[key]
            F, R, T = w["F"], w["R"], w["T"]
            small = small_problem(R, T)
            heavy, acc, _, _ = gen_features(R, F, 0, 1, dev)
            slots = [SlotSpec(L.LOSS_UPTAKE_EYE_MSE, small.S_train, small.y_train, small.S_val, small.y_val), SlotSpec(L.LOSS_MAXENT_CONVEX_KL)]
            eng = BvEngine(heavy, acc, small.k_ints, small.timepoints, slots, True, settings, prior=torch.full((F,), 1.0 / F, device=dev), device=dev)
            del heavy, acc
            eng.init_state(torch.ones(F, device=dev), 0.35, 2.0, [0.5, 0.5], [1000.0, 1000.0])
            ms, pms = eager_rate(eng, 200, 10)
            gms = graph_rate(eng, 200)
            best = min(ms, gms)
            alg = 8 * R * F + 28 * F
            out[key] = {"value": 1e3 / best, "unit": "steps/s", "ms_per_step_eager": ms, "ms_per_step_graph": gms, "pass_kernel_ms": pms,
                        "pass_roofline_frac": alg / (pms * 1e-3) / 1e9 / peak, "step_roofline_frac": alg / (best * 1e-3) / 1e9 / peak,
                        "l2": "features fit in the 126 MB L2 (latency-bound: steps/s only)" if 8 * R * F < 1.2e8 else "features 400 MB > L2"}
            eng.close()
            del eng
        except Exception as ex:  # pragma: no cover
            out[key] = {"error": str(ex)[:200]}
    return out


def oracle_check(eng, heavy_host, acc_host, small, F, R, fw, fs):
    """parity inside the bench run: one more optimise step of the timed engine, teacher-forced against the fp64 oracle at
    full size (loss and masked gradients at the logits the device had before that step)"""
    import torch

    from oracle import bv_oracle as O

    z = eng.export_state(("z",))[0]["z"].cpu().numpy()
    k = eng.steps_done
    eng.step(1)
    eng.finish_record()
    torch.cuda.synchronize()
    rec = eng.records(k, 1)[0]
    g = eng.export_state(("g",))[0]["g"].cpu().numpy()
    slots = [O.LossSlot(O.UPTAKE_EYE_MSE, small.S_train, small.y_train, small.S_val, small.y_val),
             O.LossSlot(O.MAXENT_CONVEX_KL, prior=np.full(F, 1.0 / F, np.float32))]
    pb = O.Problem(heavy_host, acc_host, small.k_ints, small.timepoints, slots, True, eng.average_first)
    par = O.Params(z.astype(np.float64), np.ones(F, np.float32), 0.35, 2.0, np.asarray(fw, np.float64), np.asarray(fs, np.float64), np.ones(2))
    t0 = time.perf_counter()
    losses, grads, _ = O.loss_and_grads(pb, par, np.float64)
    return {"step": int(rec.step), "loss_rel_err": abs(rec.total_train - losses.total_train_loss) / abs(losses.total_train_loss),
            "grad_max_rel_err": float(np.abs(g - grads.z).max() / np.abs(grads.z).max()),
            "oracle": "oracle/bv_oracle.py fp64 at full size, teacher-forced", "oracle_s": time.perf_counter() - t0}


def e2e_api_single(args, wl, dev, host_h, host_a, small, K):
    """End to end through the reference-named API, one GPU: BV features in pinned host memory -> Simulation.initialise ->
    OptaxOptimizer -> run_optimise (the reference's Python loop: one optimise step per iteration, the step's loss record read
    back to the host EVERY step, exactly what opt/run.py:276-316 does) -> history.  Also times the same job with the on-device
    loop (`_opt_fn=_optimise_device`, status read every 32 steps)."""
    import torch

    import jaxent_b200 as J
    from jaxent_b200.opt.run import _optimise, _optimise_device, run_optimise

    F, R, T = wl["F"], wl["R"], wl["T"]

    class _Model:
        def __init__(self, fp):
            self.forwardpass = fp

    def job(opt_fn):
        t0 = time.perf_counter()
        feats = J.BV_input_features(host_h, host_a, small.k_ints)
        mp = J.BV_Model_Parameters(bv_bc=np.array([0.35], np.float32), bv_bh=np.array([2.0], np.float32), timepoints=small.timepoints)
        yt = lambda y: y[:, :, None]
        loader = J.ExpD_Dataloader(train=J.Dataset([], yt(small.y_train), J.SparseFragmentMapping(small.S_train)),
                                   val=J.Dataset([], yt(small.y_val), J.SparseFragmentMapping(small.S_val)))
        w0 = np.full(F, 1.0 / F, np.float32)
        params = J.Simulation_Parameters(frame_weights=w0, frame_mask=np.ones(F, np.float32), model_parameters=[mp],
                                         forward_model_weights=np.array([1.0, 1.0], np.float32), normalise_loss_functions=np.ones(2, np.float32),
                                         forward_model_scaling=np.full(2, 1000.0, np.float32))
        fp = J.BV_uptake_ForwardPass()
        if args.per_frame:
            fp.average_first = False
        sim = J.Simulation(input_features=(feats,), forward_models=(_Model(fp),), params=params, device=dev)
        sim.initialise()
        opt = J.OptaxOptimizer(learning_rate=1.0, parameter_partition_masks={J.Optimisable_Parameters.frame_weights}, clip_value=None,
                               optimizer="adam", initial_learning_rate=1.0, initial_steps=0)
        cfg = J.OptimiserSettings(name="bench", n_steps=K, tolerance=1e-30, convergence=1e-30, learning_rate=1.0)
        torch.cuda.synchronize()
        t_setup = time.perf_counter() - t0
        t1 = time.perf_counter()
        sim2, hist = run_optimise(sim, (loader, params), cfg, forward_models=(), indexes=(0, 0),
                                  loss_functions=[J.hdx_uptake_eye_MSE_loss, J.maxent_convexKL_loss], optimizer=opt, _opt_fn=opt_fn, silent=True)
        w = sim2.params.frame_weights
        w_host = w.cpu() if hasattr(w, "cpu") else w
        last = float(hist.states[-1].losses.total_train_loss) if hist.states else float("nan")
        torch.cuda.synchronize()
        t_steps = time.perf_counter() - t1
        del sim, sim2, opt, w_host
        torch.cuda.empty_cache()
        return t_setup, t_steps, last

    ps, pk, _ = job(_optimise)   # the reference-style Python loop (explicit)
    ts, tk, last = job(None)      # the default call: run_optimise picks the on-device loop
    import ctypes

    from jaxent_b200 import _lib as L

    rec_bytes = ctypes.sizeof(L.LossRecord)
    return {
        "value": K / (ts + tk), "unit": "steps/s",
        "h2d_bytes_per_step": (2 * R * F * 4 + 3 * F * 4) / K,
        "d2h_bytes_per_step": 2 * rec_bytes + F * 4 / K,
        "steady_value": K / tk, "setup_s": ts,
        "python_loop_value": K / (ps + pk), "python_loop_steady_value": K / pk, "python_loop_us_per_step": 1e6 * pk / K,
        "api": "jaxent_b200.Simulation.initialise + OptaxOptimizer + run_optimise (reference opt/run.py:376-461), default arguments",
        "note": "whole K-step job, inputs start in pinned host memory: features uploaded + packed + forward + optimiser set-up inside the timed region "
                "(the features are the only per-job input; they are constant over the steps, so their upload is amortised over K), "
                "run_optimise keeps the loop on the device by default (convergence tracker / snapshots / stop in the tail kernel, one "
                "status read per 32 steps), history + final frame weights read back; steady_value excludes the one-off upload; "
                "python_loop_* = same job with _opt_fn=_optimise (the reference-style Python loop: the loss record of every step is "
                "read back to the host, host-synchronised each step)",
        "last_loss": last,
    }


class gpu_local_affinity:
    """context manager: run the enclosed allocation on the CPUs local to GPU `index` (sysfs local_cpulist of its PCI function),
    then restore the previous affinity.  Yields a short description (or why nothing was changed)."""

    def __init__(self, index):
        self.index, self.saved, self.note = index, None, "unchanged"

    def __enter__(self):
        try:
            import torch

            pr = torch.cuda.get_device_properties(self.index)
            bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
            base = f"/sys/bus/pci/devices/{bdf}"
            cpus = set()
            for part in open(f"{base}/local_cpulist").read().strip().split(","):
                if part:
                    a, _, b = part.partition("-")
                    cpus.update(range(int(a), int(b or a) + 1))
            node = open(f"{base}/numa_node").read().strip()
            allowed = os.sched_getaffinity(0)
            use = cpus & allowed
            if use and use != allowed:
                self.saved = allowed
                os.sched_setaffinity(0, use)
                self.note = f"pinned buffers allocated on {len(use)} CPUs local to GPU {bdf} (NUMA node {node})"
            else:
                self.note = f"GPU {bdf} NUMA node {node}: " + ("all allowed CPUs are local" if use else "no allowed CPU is local, affinity unchanged")
        except Exception as ex:  # pragma: no cover
            self.note = f"affinity unchanged ({type(ex).__name__}: {str(ex)[:80]})"
        return self.note

    def __exit__(self, *exc):
        if self.saved is not None:
            os.sched_setaffinity(0, self.saved)
        return False


def run_ours(args, wl):
    import torch
    import torch.distributed as dist

    from jaxent_b200 import _lib as L
    from jaxent_b200.engine import BvEngine, OptSettings, SlotSpec

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torchrun (one process per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    if N_BLOCKS % world:
        raise SystemExit("--gpus must divide 8")

    F, R, T = wl["F"], wl["R"], wl["T"]
    small = small_problem(R, T)
    heavy, acc, lo, hi = gen_features(R, F, rank, world, dev)
    Fl = hi - lo
    slots = [SlotSpec(L.LOSS_UPTAKE_EYE_MSE, small.S_train, small.y_train, small.S_val, small.y_val), SlotSpec(L.LOSS_MAXENT_CONVEX_KL)]
    settings = OptSettings(learning_rate=1.0, initial_learning_rate=1.0, initial_steps=0, clip_value=None)
    prior = torch.full((Fl,), 1.0 / F, device=dev)
    fw, fs = [0.5, 0.5], [1000.0, 1000.0]

    # ---- end-to-end job through the host-facing API: every rank's shard of the features starts in pinned host memory
    e2e = None
    host_h = host_a = None
    numa_note = None
    if not args.no_e2e or not args.no_cpu:
        # the job's input buffers: pinned host memory on the NUMA node the GPU hangs off (pages are placed where the allocating
        # thread runs; a remote node halves the upload rate: 0.20 s instead of 0.09 s for the 4 GB of config 4 was observed).  The
        # affinity is narrowed for the allocation only and restored right after, so the CPU arms keep all their host threads.
        with gpu_local_affinity(local) as numa_note:
            host_h = torch.empty((R, Fl), dtype=torch.float32, pin_memory=True)
            host_a = torch.empty((R, Fl), dtype=torch.float32, pin_memory=True)
            host_h.copy_(heavy)
            host_a.copy_(acc)
            torch.cuda.synchronize()

    eng = BvEngine(heavy, acc, small.k_ints, small.timepoints, slots, True, settings, prior=prior, device=dev, F_total=F, group=group,
                   exchange=args.exchange, average_first=not args.per_frame)
    del heavy, acc
    eng.init_state(torch.ones(Fl, device=dev), 0.35, 2.0, fw, fs)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up
    W = max(3, args.warmup)
    eng.step(W)
    barrier()

    # ---- timed region: exactly K steps.  The step is two kernel launches with no host synchronisation, so the loop is replayed
    # from a CUDA graph (BvEngine.capture / step_graph, two steps per graph + eager launches for an odd remainder): driving
    # 2 launches per ~120 us step through Python + ctypes from 8 processes on one host costs 4-5 % at 8 GPUs (0.4 % at one).
    # Where a step contains NCCL collectives (exchange="nccl") the loop stays eager.
    K = args.steps
    EV = max(1, args.event_every)
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream
    use_graph = not (world > 1 and not eng.fused)
    if use_graph:
        try:
            eng.capture(2)
            eng.step_graph(1)  # the replay path itself warmed up
            barrier()
        except Exception as ex:  # pragma: no cover
            use_graph = False
            sys.stderr.write(f"graph capture failed, eager launches: {ex}\n")

    def eager_step():
        L.check(eng.lib.jaxent_bv_pass(eng.plan, 1, sp))
        if eng.sharded and not eng.fused:
            dist.all_reduce(eng.red, group=group)
        L.check(eng.lib.jaxent_bv_tail(eng.plan, sp))
        if eng.sharded and not eng.fused:
            dist.all_reduce(eng.klsum, group=group)

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    e0.record()
    if use_graph:
        eng.step_graph(K // 2)
        for _ in range(K % 2):
            eager_step()
            eng.steps_done += 1
            eng.launches += 2
    else:
        for _ in range(K):
            eager_step()
        eng.steps_done += K
        eng.launches += 2 * K
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)

    # ---- the same K steps launched eagerly, with CUDA events around every EV-th pass kernel (an event between two kernels
    # breaks their PDL overlap, so only a sample is bracketed): the live duration of the dominant kernel for the roofline,
    # and the eager step time for comparison
    ev_pass = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range((K + EV - 1) // EV)]
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    g0.record()
    for k in range(K):
        timed = (k % EV) == 0
        if timed:
            ev_pass[k // EV][0].record()
        L.check(eng.lib.jaxent_bv_pass(eng.plan, 1, sp))
        if timed:
            ev_pass[k // EV][1].record()
        if eng.sharded and not eng.fused:
            dist.all_reduce(eng.red, group=group)
        L.check(eng.lib.jaxent_bv_tail(eng.plan, sp))
        if eng.sharded and not eng.fused:
            dist.all_reduce(eng.klsum, group=group)
    g1.record()
    barrier()
    sampler.stop_flag = True
    sampler.join()
    eng.steps_done += K
    eng.launches += 2 * K
    eager_ms_total = g0.elapsed_time(g1)
    pass_ms = float(np.mean([a.elapsed_time(b) for a, b in ev_pass]))
    t = torch.tensor([ms_total, pass_ms, eager_ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, pass_ms, eager_ms_total = float(t[0]), float(t[1]), float(t[2])
    ms_step = ms_total / K
    eager_ms = eager_ms_total / K
    sps = 1e3 / ms_step

    eng.finish_record()
    torch.cuda.synchronize()
    n_done = eng.steps_done
    rec = eng.records(n_done, 1)[0]
    wloc = eng.weights.double()
    wsum = wloc.sum()
    # position-weighted checksum of the frame weights (global frame index), to compare sharded and single-GPU runs
    idx = torch.arange(lo, hi, device=dev, dtype=torch.float64)
    wchk = (wloc * torch.cos(idx * 1e-3)).sum()
    if world > 1:
        dist.all_reduce(wsum)
        dist.all_reduce(wchk)
    wsum, wchk = float(wsum), float(wchk)
    exchange_kind = eng.exchange
    xerr = eng.exchange_error()
    check = {"sum_weights": wsum, "final_loss": rec.total_train, "final_step": rec.step, "exchange_error": xerr, "weights_checksum": wchk}

    peak, peak_src = peaks()
    extra = None
    if world == 1 and args.workload == "cfg4" and not args.no_extra and not args.per_frame:
        extra = extra_results(args, dev, eng, small, slots, settings, peak)

    # ---- parity carried by the bench run itself
    if rank == 0 and world == 1 and not args.no_cpu and host_h is not None:
        try:
            check["oracle"] = oracle_check(eng, host_h.numpy(), host_a.numpy(), small, F, R, fw, fs)
        except Exception as ex:  # pragma: no cover
            check["oracle"] = {"error": str(ex)[:200]}
    eng.close()
    del eng
    torch.cuda.empty_cache()
    if world > 1:
        # the driver's one-GPU pytest cannot run the multi-GPU tests: rank 0 repeats the SAME number of steps on the whole
        # ensemble on its own GPU and compares final loss and weights checksum with the sharded result
        try:
            if rank == 0:
                hh, aa, _, _ = gen_features(R, F, 0, 1, dev)
                one = BvEngine(hh, aa, small.k_ints, small.timepoints, slots, True, settings, prior=torch.full((F,), 1.0 / F, device=dev), device=dev,
                               average_first=not args.per_frame)
                del hh, aa
                one.init_state(torch.ones(F, device=dev), 0.35, 2.0, fw, fs)
                one.step(n_done)
                one.finish_record()
                torch.cuda.synchronize()
                r1 = one.records(n_done, 1)[0]
                w1 = one.weights.double()
                c1 = float((w1 * torch.cos(torch.arange(F, device=dev, dtype=torch.float64) * 1e-3)).sum())
                check["single_gpu"] = {"final_loss": r1.total_train, "final_step": r1.step, "weights_checksum": c1,
                                       "loss_rel_diff": abs(r1.total_train - rec.total_train) / abs(r1.total_train),
                                       "checksum_abs_diff": abs(c1 - wchk),
                                       "note": f"the same {n_done} steps on all {F} frames on rank 0's GPU alone vs the sharded run"}
                one.close()
                del one, w1
                torch.cuda.empty_cache()
        except Exception as ex:  # pragma: no cover
            check["single_gpu"] = {"error": str(ex)[:200]}
        barrier()

    # ---- e2e
    if not args.no_e2e and host_h is not None:
        import ctypes

        rec_bytes = ctypes.sizeof(L.LossRecord)
        if world == 1:
            e2e = e2e_api_single(args, wl, dev, host_h, host_a, small, K)
        else:
            # frame-sharded: the engine API (the reference has no multi-GPU API to mirror); whole job from each rank's pinned host shard
            barrier()
            t0 = time.perf_counter()
            eng2 = BvEngine(host_h, host_a, small.k_ints, small.timepoints, slots, True, settings, prior=prior, device=dev, F_total=F, group=group,
                            exchange=args.exchange, average_first=not args.per_frame)
            eng2.init_state(torch.ones(Fl, device=dev), 0.35, 2.0, fw, fs)
            torch.cuda.synchronize()
            t_setup = time.perf_counter() - t0
            rec_dev = eng2._records
            pinned_rec = torch.empty(rec_bytes, dtype=torch.uint8, pin_memory=True)
            t1 = time.perf_counter()
            last = 0.0
            for k in range(K):
                eng2.step(1)
                i = (k % L.RECORD_RING) * rec_bytes
                pinned_rec.copy_(rec_dev[i : i + rec_bytes], non_blocking=True)  # loss record of step k (complete)
                torch.cuda.current_stream().synchronize()
                last = float(np.frombuffer(pinned_rec.numpy().tobytes(), dtype=np.float32)[6])
            w_host = eng2.weights.cpu()
            t_steps = time.perf_counter() - t1
            tt = torch.tensor([t_setup, t_steps], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t_setup, t_steps = float(tt[0]), float(tt[1])
            e2e = {
                "value": K / (t_setup + t_steps), "unit": "steps/s",
                "h2d_bytes_per_step": (2 * R * F * 4 + F * 4 * 2) / K,
                "d2h_bytes_per_step": rec_bytes * world + F * 4 / K,
                "steady_value": K / t_steps, "setup_s": t_setup,
                "api": "jaxent_b200.engine.BvEngine (frame-sharded engine API; the reference has no multi-GPU API)",
                "note": "whole job, max over ranks, byte counts summed over ranks; value = whole K-step job from pinned host features (upload+pack+exchange "
                        "set-up+init inside the timed region, loss record read back and host-synchronised every step, final weights read back); "
                        "steady_value excludes the one-off upload",
                "last_loss": last,
            }
            eng2.close()
            del eng2, w_host
        if e2e is not None:
            e2e["host_buffers"] = numa_note
            e2e["upload_GBps"] = (2 * R * Fl * 4) / max(e2e["setup_s"], 1e-9) / 1e9  # lower bound: set-up does more than the copy

    if world == 1 and rank == 0 and args.workload == "cfg4" and not args.no_extra and not args.per_frame:
        extra.update(extra_small_configs(args, dev, peak))
        try:
            extra["cfg5"] = batched_summary(args, dev)
        except Exception as ex:  # pragma: no cover
            extra["cfg5"] = {"error": str(ex)[:200]}

    traffic = None
    try:  # dram__bytes_read+write per bv_pass_kernel launch from the committed ncu --set full capture of this workload
        for tag in ("r2", "r1"):
            fn = os.path.join(ROOT, "profiles", f"ncu_pass_{tag}_{args.workload}.json")
            if world == 1 and os.path.exists(fn):
                traffic = json.load(open(fn))["dram_bytes_per_launch"]
                break
    except Exception:
        pass
    alg_bytes = 8 * R * Fl + 28 * Fl
    achieved = alg_bytes / (pass_ms * 1e-3) / 1e9

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            hh, aa = host_h.numpy(), host_a.numpy()
            c_sps, c_n, c_dt, _ = cpu_steps(F, R, T, hh, aa, small, budget_s=15.0, max_steps=50, average_first=not args.per_frame)
            what = (f"optimise steps on the first {PF_CPU_FRAMES} frames (per-frame mode; steps/s scaled by {PF_CPU_FRAMES}/{F})"
                    if args.per_frame and F > PF_CPU_FRAMES else "full-size optimise steps")
            cpu = {"value": c_sps, "unit": "steps/s", "cores": cpu_threads_used(), "kind": "port",
                   "sample": f"{c_n} {what} of the NumPy fp32 oracle port in {c_dt:.1f} s (oracle restatement, not the reference's JAX loop)"}
        sfu = None
        if args.per_frame:  # the bound that matters in this mode: 16 special-function results per clock per SM
            clk = sampler.result().get("sm_mhz") or 1900.0
            sf_peak = 16 * 148 * clk * 1e6 / 1e9
            sf_ach = Fl * R * (T + 1) / (pass_ms * 1e-3) / 1e9
            sfu = {"achieved": sf_ach, "peak": sf_peak, "unit": "G ex2/s", "frac": sf_ach / sf_peak,
                   "note": "F*R*(T+1) ex2.approx per launch against 16/clk/SM at the sampled SM clock"}
        line = {
            "metric": "optimise steps/sec",
            "value": sps,
            "unit": "steps/s",
            "n_gpus": world,
            "steps": K,
            "warmup": W,
            "ms_per_step": ms_step,
            "higher_is_better": True,
            "scaling": "strong",
            "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "frame_residue_timepoint_per_s": sps * F * R * T,
            "config": workload_config(args, wl),
            "run_details": {"frames_per_gpu": Fl,
                            "exchange": {"nvlink": "fused in-kernel all-reduce over NVLink peer memory (no collective launches)",
                                         "nccl": "two NCCL all-reduces per step", "none": "none (single GPU)"}[exchange_kind],
                            "launch": ("CUDA-graph replay, two steps per graph (BvEngine.capture / step_graph)" if use_graph
                                       else "eager launches (NCCL collectives inside the step)"),
                            "pass_events": f"CUDA events around every {EV}th pass kernel of {K} eagerly launched steps that follow the timed region",
                            "roofline_numerator": "8*R*F + 28*F bytes per launch (SURVEY 8d)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "bv_pass_kernel", "algorithmic_bytes_per_launch": alg_bytes,
                         "kernel_ms": pass_ms, "peak_source": peak_src, "step_frac": alg_bytes / (ms_step * 1e-3) / 1e9 / peak,
                         **({"special_function": sfu} if sfu else {})},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": 2 * K,
            "clocks": sampler.result(),
            "eager_ms_per_step": eager_ms,
            "check": check,
        }
        if extra is not None:
            line["extra"] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--exchange", default="auto", choices=["auto", "nvlink", "nccl"])
    ap.add_argument("--event-every", type=int, default=8)
    ap.add_argument("--pieces", type=int, default=3, help="cfg5: bf16 pieces of the fp32 operand (2 or 3)")
    ap.add_argument("--per-frame", action="store_true", help="BV uptake per frame, then frame-averaged (average_first=False)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the config 3 / 1 / 5 / per-frame summaries of the default run")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    elif "B" in wl:
        run_batched(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
