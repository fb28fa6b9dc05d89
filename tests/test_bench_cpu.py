"""bench.py contract checks that need no GPU: the reference arm's JSON line, and that our arm fails loudly (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run("--impl", "reference", "--workload", "cfg1", "--steps", "3", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "optimise steps/sec" and d["unit"] == "steps/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["value"] > 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"]["frames"] == 500 and d["config"]["residues"] == 53 and d["config"]["timepoints"] == 3
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "oracle" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1", "--gpus", "2", "--steps", "2"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_our_arm_has_no_cpu_fallback():
    r = _run("--workload", "cfg1", "--steps", "2", "--warmup", "1")
    assert r.returncode != 0 and not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_reference_arm_uses_all_host_threads_under_torchrun_env_and_prints_our_config():
    """torchrun exports OMP_NUM_THREADS=1 to its workers (the CPU arm ran single-threaded while reporting 32
    cores): the arm sets the thread count itself and reports what the BLAS really uses; `config` is the same dict as our arm's"""
    sys.path.insert(0, ROOT)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg2", "--gpus", "2", "--steps", "4", "--warmup", "5"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][0])
    avail = len(os.sched_getaffinity(0))
    assert d["cpu_baseline"]["cores"] == avail and d["cpu_baseline"]["omp_num_threads_env"] == str(avail)
    assert d["steps"] == 4 and d["warmup"] == 5 and d["n_gpus"] == 2
    import argparse

    import bench

    ours = bench.workload_config(argparse.Namespace(gpus=2, per_frame=False), bench.WORKLOADS["cfg2"])
    assert d["config"] == ours


def test_gpu_local_affinity_always_restores_the_affinity():
    """the pinned input buffers are allocated under a narrowed CPU affinity (GPU-local NUMA node); whatever happens inside --
    no GPU here, so the lookup fails -- the process affinity is what it was, and the note says what was done"""
    sys.path.insert(0, ROOT)
    import bench

    before = os.sched_getaffinity(0)
    with bench.gpu_local_affinity(0) as note:
        assert isinstance(note, str) and note
    assert os.sched_getaffinity(0) == before
    # narrowed-and-restored path, driven by hand
    ctx = bench.gpu_local_affinity(0)
    ctx.saved = before
    os.sched_setaffinity(0, {min(before)})
    ctx.__exit__(None, None, None)
    assert os.sched_getaffinity(0) == before
