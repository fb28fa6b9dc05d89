"""The jax.ffi handlers (csrc/xla_ffi.cc, built against tests/ffi_stub) doing real work on the GPU: the same optimise steps
through JaxentBvForwardInit / JaxentBvStep as through the C ABI, a workspace that XLA 'moved' between two calls, and the
pack / predict handlers against the direct calls."""
import ctypes as C

import numpy as np
import pytest
import torch

import helpers as H
from jaxent_b200 import _build, _lib as L
from oracle import bv_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ffi():
    L.load()
    lib = C.CDLL(_build.build_ffi_test())
    lib.ffi_drv_workspace_call.restype = C.c_int
    lib.ffi_drv_workspace_call.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_char_p, C.c_int]
    lib.ffi_drv_pack_call.restype = C.c_int
    lib.ffi_drv_pack_call.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_longlong, C.c_void_p,
                                      C.c_longlong, C.c_char_p, C.c_int]
    lib.ffi_drv_predict_call.restype = C.c_int
    lib.ffi_drv_predict_call.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong,
                                         C.c_longlong, C.c_float, C.c_float, C.c_void_p, C.c_longlong, C.c_char_p, C.c_int]
    return lib


def _ws_call(ffi, name, ptr, nbytes, handle):
    err = C.create_string_buffer(512)
    rc = ffi.ffi_drv_workspace_call(C.cast(getattr(ffi, name), C.c_void_p), torch.cuda.current_stream().cuda_stream, ptr, ptr, nbytes, handle,
                                    err, 512)
    assert rc == 0, err.value.decode()


def _records(eng, holder, n):
    """loss records read through the plan's (re-based) buffer table; `holder` is the tensor the workspace now lives in"""
    ptr, nb = C.c_void_p(), C.c_size_t()
    L.check(eng.lib.jaxent_bv_plan_buffer(eng.plan, L.BUF_RECORDS, C.byref(ptr), C.byref(nb)))
    off = ptr.value - holder.data_ptr()
    assert 0 <= off and off + nb.value <= holder.numel()
    torch.cuda.synchronize()
    raw = bytes(holder[off:off + nb.value].cpu().numpy())
    recs = (L.LossRecord * (nb.value // C.sizeof(L.LossRecord))).from_buffer_copy(raw)
    return [recs[i] for i in range(n)]


def test_steps_through_the_ffi_handlers_equal_the_c_abi(ffi):
    case = H.make_case(3000, 40, 4, O.UPTAKE_EYE_MSE, extra=(O.PARAMS_L2,), seed=3)
    cfg = O.OptConfig(learning_rate=1.0, clip_value=None, optimise_model_parameters=True)
    ref = H.make_engine(case, cfg)
    H.init_engine(ref, case)
    ref.step(5)
    ref.finish_record()
    torch.cuda.synchronize()
    want = [(r.total_train, r.bc, r.bh, r.lr) for r in ref.records(0, 6)]

    eng = H.make_engine(case, cfg)
    H.init_engine(eng, case)  # state_init + forward through the C ABI: the engine owns plan and workspace
    lib = eng.lib
    h = lib.jaxent_handle_register(eng.plan, 1)
    assert h > 0
    base, nb = C.c_void_p(), C.c_size_t()
    L.check(lib.jaxent_bv_plan_workspace(eng.plan, C.byref(base), C.byref(nb)))
    for _ in range(2):
        _ws_call(ffi, "JaxentBvStep", base.value, nb.value, h)
    # XLA could not donate the operand: the next call sees a COPY of the workspace at another address
    torch.cuda.synchronize()
    moved = torch.empty(nb.value + 256, dtype=torch.uint8, device="cuda")
    off = (-moved.data_ptr()) % 256
    src_off = base.value - eng.ws.data_ptr()
    moved[off:off + nb.value].copy_(eng.ws[src_off:src_off + nb.value])
    eng.ws.zero_()  # the old allocation is dead: a handler that kept writing there would be caught
    torch.cuda.synchronize()
    new_base = moved.data_ptr() + off
    _ws_call(ffi, "JaxentBvPass", new_base, nb.value, h)
    _ws_call(ffi, "JaxentBvTail", new_base, nb.value, h)
    for _ in range(2):
        _ws_call(ffi, "JaxentBvStep", new_base, nb.value, h)
    L.check(lib.jaxent_bv_plan_workspace(eng.plan, C.byref(base), C.byref(nb)))
    assert base.value == new_base
    L.check(lib.jaxent_bv_finish_record(eng.plan, torch.cuda.current_stream().cuda_stream))
    got = [(r.total_train, r.bc, r.bh, r.lr) for r in _records(eng, moved, 6)]
    assert got == want  # bitwise: the same kernels on the same state
    # hand the plan back to its original allocation before the engine tears it down
    L.check(lib.jaxent_bv_plan_rebind_workspace(eng.plan, eng.ws.data_ptr() + src_off, nb.value))


def test_pack_and_predict_handlers(ffi):
    rng = np.random.default_rng(0)
    R, F, T = 37, 501, 3
    heavy = torch.tensor(rng.integers(0, 30, (R, F)), dtype=torch.float32, device="cuda")
    acc = torch.tensor(rng.integers(0, 4, (R, F)), dtype=torch.float32, device="cuda")
    lib = L.load()
    n = lib.jaxent_bv_packed_bytes(R, F) // 4
    direct = torch.zeros(n, device="cuda")
    L.check(lib.jaxent_bv_pack_features_f32(heavy.data_ptr(), acc.data_ptr(), R, F, F, 0, F, direct.data_ptr(), torch.cuda.current_stream().cuda_stream))
    via = torch.zeros(n, device="cuda")
    err = C.create_string_buffer(512)
    s = torch.cuda.current_stream().cuda_stream
    rc = ffi.ffi_drv_pack_call(C.cast(ffi.JaxentBvPack, C.c_void_p), s, heavy.data_ptr(), acc.data_ptr(), R, F, F, via.data_ptr(), n, err, 512)
    assert rc == 0, err.value
    torch.cuda.synchronize()
    assert torch.equal(direct, via)
    k = torch.tensor(10.0 ** rng.uniform(-2, 2, R), dtype=torch.float32, device="cuda")
    tp = torch.tensor([0.167, 1.0, 10.0], dtype=torch.float32, device="cuda")
    out = torch.zeros(T * R * F, device="cuda")
    rc = ffi.ffi_drv_predict_call(C.cast(ffi.JaxentBvPredict, C.c_void_p), s, via.data_ptr(), n, k.data_ptr(), tp.data_ptr(), R, F, T, 0.35, 2.0,
                                  out.data_ptr(), T * R * F, err, 512)
    assert rc == 0, err.value
    torch.cuda.synchronize()
    lp = 0.35 * heavy.double() + 2.0 * acc.double()
    want = 1.0 - torch.exp(-k.double()[None, :, None] * tp.double()[:, None, None] * torch.exp(-lp)[None])
    assert torch.allclose(out.view(T, R, F).double(), want, rtol=2e-5, atol=2e-6)
