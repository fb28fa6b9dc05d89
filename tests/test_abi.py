"""CPU tests of the drop-in boundary: the C-ABI library builds/loads, exports every symbol that
include/jaxent_b200.h declares, its structs match the ctypes mirror, argument validation raises
the documented errors, and host-side logic (CSR conversion, sizes) is right.  No compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from jaxent_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "jaxent_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(jaxent_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/jaxent_b200.h but not exported"
        assert name in L.SYMBOLS, f"{name} has no ctypes prototype"
    assert sorted(L.SYMBOLS) == declared
    assert lib.jaxent_version() >= 100


def test_struct_layouts_match():
    lib = L.load()
    for i, st in enumerate((L.BvProblem, L.OptConfig, L.LossRecord, L.LossSlot, L.PeptideMap, L.TrackSnapshot, L.TrackStatus)):
        assert lib.jaxent_abi_sizeof(i) == C.sizeof(st)
    assert lib.jaxent_abi_sizeof(99) == -1


def test_packed_and_workspace_sizes():
    lib = L.load()
    assert lib.jaxent_bv_packed_bytes(500, 1_000_000) == 8 * 500 * 1_000_000
    assert lib.jaxent_bv_packed_bytes(53, 500) == 8 * 53 * 504  # frames padded to a multiple of 8
    assert lib.jaxent_bv_packed_bytes(0, 10) == 0
    assert lib.jaxent_bv_packed_bytes_fmt(500, 1_000_000, 1) == 2 * 500 * 1_000_000  # u8: a quarter of the fp32 bytes
    assert lib.jaxent_bv_packed_bytes_fmt(53, 500, 1) == 2 * 53 * 512  # frames padded to a multiple of 32
    assert lib.jaxent_bv_packed_bytes_fmt(53, 500, 0) == lib.jaxent_bv_packed_bytes(53, 500)
    pb = L.BvProblem()
    pb.R, pb.T, pb.uptake, pb.n_slots, pb.F, pb.F_total = 500, 8, 1, 2, 100_000, 100_000
    n = lib.jaxent_bv_workspace_bytes(C.byref(pb))
    assert n > 16 * 4 * 100_000
    assert n < 64 * 1024 * 1024


def _valid_problem():
    pb = L.BvProblem()
    pb.R, pb.T, pb.uptake, pb.n_slots, pb.F, pb.F_total = 16, 3, 1, 1, 64, 64
    pb.packed = 0x1000  # never dereferenced by plan_create
    pb.k_ints = 0x2000
    pb.timepoints = 0x3000
    pb.slots[0].kind = L.LOSS_PARAMS_L2
    return pb


def _create(pb, cfg=None, ws=0x100000, nbytes=1 << 30):
    lib = L.load()
    cfg = cfg or L.OptConfig(1.0, 1.0, 0, 1.005, 1.0, -1.0, 1, 0, 0.9, 0.999, 1e-8)
    plan = C.c_void_p()
    rc = lib.jaxent_bv_plan_create(C.byref(pb), C.byref(cfg), ws, nbytes, 148, C.byref(plan))
    return rc, plan, lib.jaxent_last_error().decode()


def test_plan_create_validation_errors():
    lib = L.load()
    rc, plan, msg = _create(_valid_problem())
    assert rc == 0 and plan.value
    lib.jaxent_bv_plan_destroy(plan)

    pb = _valid_problem(); pb.R = 0
    assert _create(pb)[0] == L.E_INVALID
    pb = _valid_problem(); pb.packed = 0x1004
    rc, _, msg = _create(pb); assert rc == L.E_INVALID and "aligned" in msg
    pb = _valid_problem(); pb.n_slots = 7
    assert _create(pb)[0] == L.E_INVALID
    pb = _valid_problem(); pb.slots[0].kind = 42
    rc, _, msg = _create(pb); assert rc == L.E_UNSUPPORTED and "unsupported loss kind" in msg
    pb = _valid_problem(); pb.slots[0].kind = L.LOSS_PF_L2  # PF loss on the uptake forward pass
    assert _create(pb)[0] == L.E_UNSUPPORTED
    pb = _valid_problem(); pb.uptake = 0; pb.slots[0].kind = L.LOSS_UPTAKE_EYE_MSE
    assert _create(pb)[0] == L.E_UNSUPPORTED
    pb = _valid_problem(); pb.slots[0].kind = L.LOSS_MAXENT_CONVEX_KL  # no prior
    assert _create(pb)[0] == L.E_INVALID
    pb = _valid_problem(); pb.R = 1025
    assert _create(pb)[0] == L.E_UNSUPPORTED
    rc, _, msg = _create(_valid_problem(), nbytes=1024); assert rc == L.E_WORKSPACE and "too small" in msg
    rc, _, msg = _create(_valid_problem(), ws=0x100010); assert rc == L.E_INVALID
    bad = L.OptConfig(1.0, 1.0, 0, 0.0, 1.0, -1.0, 1, 0, 0.9, 0.999, 1e-8)
    assert _create(_valid_problem(), cfg=bad)[0] == L.E_INVALID


def test_calls_before_state_init_fail_loudly():
    lib = L.load()
    rc, plan, _ = _create(_valid_problem())
    assert rc == 0
    assert lib.jaxent_bv_pass(plan, 1, None) == L.E_INVALID
    assert "state_init" in lib.jaxent_last_error().decode()
    assert lib.jaxent_bv_step(plan, None) == L.E_INVALID
    assert lib.jaxent_bv_pass(None, 1, None) == L.E_INVALID
    with pytest.raises(L.JaxentError):
        L.check(lib.jaxent_bv_tail(plan, None))
    lib.jaxent_bv_plan_destroy(plan)


def test_pack_argument_validation():
    lib = L.load()
    assert lib.jaxent_bv_pack_features_f32(None, None, 4, 8, 8, 0, 8, None, None) == L.E_INVALID
    assert lib.jaxent_bv_pack_features_f32(0x1000, 0x2000, 4, 8, 4, 0, 8, 0x3000, None) == L.E_INVALID  # ld < F
    assert lib.jaxent_bv_pack_features_f32(0x1000, 0x2000, 4, 8, 8, 2, 16, 0x3000, None) == L.E_INVALID  # offset % 4
    assert lib.jaxent_bv_pack_features_f32(0x1000, 0x2000, 4, 8, 8, 0, 8, 0x3004, None) == L.E_INVALID  # alignment


def test_engine_refuses_to_run_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from jaxent_b200.engine import BvEngine, OptSettings, SlotSpec

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        BvEngine(np.zeros((4, 8), np.float32), np.zeros((4, 8), np.float32), None, None,
                 [SlotSpec(L.LOSS_PARAMS_L2)], False, OptSettings())


def test_csr_conversion_round_trip():
    from jaxent_b200.engine import _csr

    rng = np.random.default_rng(0)
    S = (rng.random((7, 11)) < 0.3) * rng.random((7, 11))
    ptr, col, val = _csr(S.astype(np.float32))
    D = np.zeros_like(S, dtype=np.float32)
    for p in range(7):
        D[p, col[ptr[p]:ptr[p + 1]]] = val[ptr[p]:ptr[p + 1]]
        assert np.all(np.diff(col[ptr[p]:ptr[p + 1]]) > 0)
    np.testing.assert_array_equal(D, S.astype(np.float32))
    tptr, trow, tval = _csr(S.T.astype(np.float32))
    assert tptr[-1] == ptr[-1]
    ptr0, _, _ = _csr(np.zeros((3, 5), np.float32))  # empty (ragged) map
    assert list(ptr0) == [0, 0, 0, 0]


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "jax-ent_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn


# ---------------------------------------------------------------------------------------- round-1 additions
def test_per_frame_mode_validation():
    """uptake = 2 (ForwardPass.average_first = False): fp32 features, <= 512 residues, <= 8 time points; optimised BV parameters on
    an unsharded ensemble only (their gradient is a frame sum of its own that the exchange does not carry)"""
    lib = L.load()
    pb = _valid_problem(); pb.uptake = 2; pb.slots[0].kind = L.LOSS_PARAMS_L2
    rc, plan, _ = _create(pb)
    assert rc == 0
    lib.jaxent_bv_plan_destroy(plan)
    opt_model = L.OptConfig(1.0, 1.0, 0, 1.005, 1.0, -1.0, 1, 1, 0.9, 0.999, 1e-8)
    rc, plan, msg = _create(pb, cfg=opt_model)
    assert rc == 0, msg
    lib.jaxent_bv_plan_destroy(plan)
    pb.F_total = 4 * pb.F  # a frame shard
    rc, _, msg = _create(pb, cfg=opt_model)
    assert rc == L.E_UNSUPPORTED and "one GPU" in msg
    pb = _valid_problem(); pb.uptake = 2; pb.T = 9
    assert _create(pb)[0] == L.E_UNSUPPORTED
    pb = _valid_problem(); pb.uptake = 2; pb.feature_format = 1
    assert _create(pb)[0] == L.E_UNSUPPORTED
    pb = _valid_problem(); pb.uptake = 3
    assert _create(pb)[0] == L.E_INVALID
    # the reduced vector grows from 4R+4 to 2TR+4 doubles: so do the exchange buffers
    pb = _valid_problem(); pb.R, pb.T = 500, 8
    pb.uptake = 1
    a = lib.jaxent_bv_exchange_bytes_for(C.byref(pb), 8)
    pb.uptake = 2
    b = lib.jaxent_bv_exchange_bytes_for(C.byref(pb), 8)
    assert a == lib.jaxent_bv_exchange_bytes(500, 8) and b > 3 * a


def test_exchange_sizes_and_validation():
    lib = L.load()
    n1, n8 = lib.jaxent_bv_exchange_bytes(500, 1), lib.jaxent_bv_exchange_bytes(500, 8)
    assert n1 >= (4 * 500 + 4 + 4) * 8 and n1 % 256 == 0
    # 2 parities x 8 ranks x (4R+4 + 4) stamped 16-byte words + the plain copy + the error word
    assert n8 >= 2 * 8 * (4 * 500 + 4 + 4) * 16 + n1 - 256 and n8 % 256 == 0
    assert lib.jaxent_bv_exchange_bytes(500, 9) == 0 and lib.jaxent_bv_exchange_bytes(0, 2) == 0
    rc, plan, _ = _create(_valid_problem())
    assert rc == 0
    peers = (C.c_void_p * 2)(0x10000, 0x20000)
    assert lib.jaxent_bv_plan_set_exchange(plan, 0, 2, peers) == 0
    assert lib.jaxent_bv_plan_set_exchange(plan, 2, 2, peers) == L.E_INVALID  # rank outside the world
    assert lib.jaxent_bv_plan_set_exchange(plan, 0, 9, peers) == L.E_INVALID
    bad = (C.c_void_p * 2)(0x10000, 0x20008)
    assert lib.jaxent_bv_plan_set_exchange(plan, 0, 2, bad) == L.E_INVALID  # misaligned peer buffer
    assert lib.jaxent_bv_plan_set_exchange(plan, 0, 2, (C.c_void_p * 2)(0x10000, None)) == L.E_INVALID
    assert lib.jaxent_bv_pass(plan, 1, None) == L.E_INVALID  # state_init must follow set_exchange
    lib.jaxent_bv_plan_destroy(plan)


def _batched_problem():
    pb = _valid_problem()
    pb.feature_format = 2
    return pb


def test_batched_plan_validation():
    lib = L.load()
    cfg = L.OptConfig(1.0, 1.0, 0, 1.005, 1.0, -1.0, 1, 0, 0.9, 0.999, 1e-8)

    def create(pb, B=16, pieces=3, ws=0x100000, nbytes=1 << 34):
        plan = C.c_void_p()
        rc = lib.jaxent_bg_plan_create(C.byref(pb), C.byref(cfg), B, pieces, ws, nbytes, 148, C.byref(plan))
        return rc, plan, lib.jaxent_last_error().decode()

    rc, plan, _ = create(_batched_problem())
    assert rc == 0 and plan.value
    assert lib.jaxent_bg_step(plan, None) == L.E_INVALID  # state_init has not run
    # tracker mode: 0 Python loop (ConvergenceTracker), 1 pure loop (_optimise_pure, what batch_optimise vmaps)
    assert lib.jaxent_bg_track_mode(plan, 1) == 0 and lib.jaxent_bg_track_mode(plan, 0) == 0
    assert lib.jaxent_bg_track_mode(plan, 2) == L.E_INVALID and "mode" in lib.jaxent_last_error().decode()
    assert lib.jaxent_bg_track_mode(None, 1) == L.E_INVALID
    lib.jaxent_bg_plan_destroy(plan)
    assert create(_batched_problem(), B=15)[0] == L.E_INVALID  # replicas are padded to a multiple of 16 by the caller
    assert create(_batched_problem(), B=272)[0] == L.E_INVALID
    assert create(_batched_problem(), pieces=4)[0] == L.E_INVALID
    pb = _batched_problem(); pb.feature_format = 0
    rc, _, msg = create(pb); assert rc == L.E_INVALID and "bf16" in msg
    pb = _batched_problem(); pb.slots[0].kind = L.LOSS_MAXENT_CONVEX_KL
    assert create(pb)[0] == L.E_INVALID  # no prior
    rc, _, msg = create(_batched_problem(), nbytes=4096); assert rc == L.E_WORKSPACE
    pb = _batched_problem()
    n16 = lib.jaxent_bg_workspace_bytes(C.byref(pb), 16, 3, 148)
    n256 = lib.jaxent_bg_workspace_bytes(C.byref(pb), 256, 3, 148)
    assert 0 < n16 < n256
    assert lib.jaxent_bg_workspace_bytes(C.byref(pb), 17, 3, 148) == 0
    # feature layout: residues and frames padded to 128, two arrays, bf16
    assert lib.jaxent_bg_features_bytes(500, 100_000) == 512 * 100_096 * 2 * 2
    assert lib.jaxent_bg_features_bytes(0, 5) == 0
    # GEMM entry points validate before launching
    assert lib.jaxent_bg_gemm1(0x1000, 0x2000, 0x3000, 4, 60, 64, 3, 148, None) == L.E_INVALID  # residues not padded to 128
    assert lib.jaxent_bg_gemm1(0x1000, 0x2000, 0x3000, 4, 64, 40, 3, 148, None) == L.E_INVALID  # Bn not a multiple of 16
    assert lib.jaxent_bg_gemm2(0x1000, 0x2000, 0x3000, 4, 64, 64, 3, 99, None) == L.E_INVALID  # more k-splits than half blocks
    assert lib.jaxent_bg_ctl_sizeof() % 16 == 0 and lib.jaxent_bg_ctl_offset(0) % 8 == 0 and lib.jaxent_bg_ctl_offset(7) > 0 and lib.jaxent_bg_ctl_offset(9) == -1


def test_contacts_and_predict_argument_validation():
    lib = L.load()
    ok = dict(coords=0x1000, F=4, A=10, ti=0x2000, tr=0x3000, Nt=2, ci=0x4000, cr=0x5000, Nc=5, lo=-2, hi=2, radius=6.5, sw=0, box=None, out=0x6000)

    def call(**kw):
        a = dict(ok, **kw)
        return lib.jaxent_bv_contacts(a["coords"], a["F"], a["A"], a["ti"], a["tr"], a["Nt"], a["ci"], a["cr"], a["Nc"], a["lo"], a["hi"],
                                      a["radius"], a["sw"], a["box"], a["out"], None)

    assert call(coords=None) == L.E_INVALID
    assert call(Nt=0) == L.E_INVALID
    assert call(radius=0.0) == L.E_INVALID
    assert call(lo=3, hi=2) == L.E_INVALID
    assert "bv_contacts" in lib.jaxent_last_error().decode()
    assert lib.jaxent_bv_contacts_cell(0x1000, 4, 10, 0x2000, 0x3000, 2, 0x4000, 0x5000, 5, -2, 2, 6.5, 0, 0x7000, 4, 0x6000, None) == L.E_INVALID
    assert "box_stride" in lib.jaxent_last_error().decode()
    assert lib.jaxent_bv_predict_frames(None, 4, 8, 0, None, None, 0.35, 2.0, 0x1000, None) == L.E_INVALID
    assert lib.jaxent_bv_predict_frames(0x1000, 4, 8, 3, None, None, 0.35, 2.0, 0x2000, None) == L.E_INVALID  # uptake needs k_ints / time points
    assert lib.jaxent_bv_predict_frames(0x1004, 4, 8, 0, None, None, 0.35, 2.0, 0x2000, None) == L.E_INVALID  # alignment


def test_build_is_hash_based_and_current():
    from jaxent_b200 import _build

    assert os.path.exists(_build.LIB) and os.path.exists(_build.HASH)
    assert _build.needs_build() is False  # the in-tree .so matches the sources: nothing is rebuilt at import time
    assert len(_build.source_hash()) == 64
