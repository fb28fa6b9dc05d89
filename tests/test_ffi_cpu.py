"""The jax.ffi layer (csrc/xla_ffi.cc) without jax: it is compiled against the stand-in header of tests/ffi_stub (so type
errors in the handlers surface in this container) and its handlers are driven through the stand-in call frame.  CPU part:
everything a handler decides BEFORE it enqueues work -- plan handles, workspace identity, aliasing, malformed frames."""
import ctypes as C

import pytest

from jaxent_b200 import _build, _lib as L
from test_abi import _create, _valid_problem

HANDLERS = ["JaxentBvStep", "JaxentBvForwardInit", "JaxentBvPass", "JaxentBvTail", "JaxentBvPack", "JaxentBgStep", "JaxentBvPredict"]
K_INVALID, K_INTERNAL = 3, 13


@pytest.fixture(scope="module")
def ffi():
    L.load()
    lib = C.CDLL(_build.build_ffi_test())
    for name in ("ffi_drv_workspace_call", "ffi_drv_pack_call", "ffi_drv_predict_call", "ffi_drv_malformed_call"):
        getattr(lib, name).restype = C.c_int
    lib.ffi_drv_workspace_call.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_char_p, C.c_int]
    lib.ffi_drv_pack_call.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_longlong, C.c_void_p,
                                      C.c_longlong, C.c_char_p, C.c_int]
    lib.ffi_drv_malformed_call.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_char_p, C.c_int]
    return lib


def _fn(ffi, name):
    return C.cast(getattr(ffi, name), C.c_void_p)


def _call(ffi, name, ws_in, ws_out, nbytes, handle):
    err = C.create_string_buffer(512)
    rc = ffi.ffi_drv_workspace_call(_fn(ffi, name), None, ws_in, ws_out, nbytes, handle, err, 512)
    return rc, err.value.decode()


def test_shim_builds_against_the_stub_and_exports_its_handlers(ffi):
    for name in HANDLERS:
        assert hasattr(ffi, name), name


def test_plan_handles_are_generation_checked():
    lib = L.load()
    rc, plan, _ = _create(_valid_problem())
    assert rc == 0
    h = lib.jaxent_handle_register(plan, 1)
    assert h > 0 and lib.jaxent_handle_register(plan, 1) == h  # idempotent
    assert lib.jaxent_handle_lookup(h, 1) == plan.value
    assert lib.jaxent_handle_lookup(h, 2) is None  # a batched-plan call must not accept a single-ensemble handle
    assert lib.jaxent_handle_lookup(h + (1 << 16), 1) is None and lib.jaxent_handle_lookup(0, 1) is None
    assert lib.jaxent_handle_lookup(plan.value, 1) is None  # a raw pointer is not a handle
    lib.jaxent_bv_plan_destroy(plan)
    assert lib.jaxent_handle_lookup(h, 1) is None and "stale" in lib.jaxent_last_error().decode()
    rc, plan2, _ = _create(_valid_problem())
    h2 = lib.jaxent_handle_register(plan2, 1)
    assert h2 != h and lib.jaxent_handle_lookup(h, 1) is None  # the slot is reused, the old handle stays dead
    lib.jaxent_bv_plan_destroy(plan2)
    assert lib.jaxent_handle_register(None, 1) == 0 and lib.jaxent_handle_register(plan2, 7) == 0


def test_handlers_reject_stale_handles_and_unaliased_workspaces(ffi):
    lib = L.load()
    rc, plan, _ = _create(_valid_problem(), ws=0x100000)
    h = lib.jaxent_handle_register(plan, 1)
    nb = 1 << 30
    for name in ("JaxentBvStep", "JaxentBvForwardInit", "JaxentBvPass", "JaxentBvTail"):
        rc, msg = _call(ffi, name, 0x100000, 0x100000, nb, 987654321)
        assert rc == K_INVALID and "handle" in msg, (name, msg)
        rc, msg = _call(ffi, name, 0x100000, 0x200000, nb, h)  # result is not the operand: no in-place alias
        assert rc == K_INVALID and "alias" in msg, (name, msg)
        rc, msg = _call(ffi, name, 0x100000, 0x100000, 1024, h)
        assert rc == K_INVALID and "smaller" in msg, (name, msg)
        rc, msg = _call(ffi, name, 0x100000, 0x100000, nb, h)  # resolves; the C ABI then refuses the un-initialised plan
        assert rc == K_INVALID and "state_init" in msg, (name, msg)
    rc, msg = _call(ffi, "JaxentBgStep", 0x100000, 0x100000, nb, h)  # single-ensemble handle on the batched call
    assert rc == K_INVALID and "handle" in msg
    lib.jaxent_bv_plan_destroy(plan)
    rc, msg = _call(ffi, "JaxentBvStep", 0x100000, 0x100000, nb, h)
    assert rc == K_INVALID and "stale" in msg


def test_handler_rebinds_a_moved_workspace(ffi):
    """XLA passes a copy of an operand it could not donate: the plan must follow the buffer instead of updating the old one"""
    lib = L.load()
    rc, plan, _ = _create(_valid_problem(), ws=0x100000)
    h = lib.jaxent_handle_register(plan, 1)
    base, nb = C.c_void_p(), C.c_size_t()
    assert lib.jaxent_bv_plan_workspace(plan, C.byref(base), C.byref(nb)) == 0 and base.value == 0x100000 and nb.value > 0
    red0, n0 = C.c_void_p(), C.c_size_t()
    lib.jaxent_bv_plan_buffer(plan, L.BUF_RED, C.byref(red0), C.byref(n0))
    rc, msg = _call(ffi, "JaxentBvStep", 0x40000000, 0x40000000, 1 << 30, h)
    assert rc == K_INVALID and "state_init" in msg  # (the plan was never initialised: nothing is launched)
    lib.jaxent_bv_plan_workspace(plan, C.byref(base), C.byref(nb))
    assert base.value == 0x40000000
    red1 = C.c_void_p()
    lib.jaxent_bv_plan_buffer(plan, L.BUF_RED, C.byref(red1), C.byref(n0))
    assert red1.value - 0x40000000 == red0.value - 0x100000  # derived pointers moved with the base
    assert lib.jaxent_bv_plan_rebind_workspace(plan, 0x40000010, 1 << 30) == L.E_INVALID  # misaligned
    assert lib.jaxent_bv_plan_rebind_workspace(plan, 0x40000000, 16) == L.E_WORKSPACE
    lib.jaxent_bv_plan_destroy(plan)


def test_binding_layer_refuses_malformed_frames(ffi):
    err = C.create_string_buffer(512)
    fn = _fn(ffi, "JaxentBvStep")
    assert ffi.ffi_drv_malformed_call(fn, None, 0x100000, 1 << 20, 0, err, 512) == K_INVALID and b"attribute" in err.value  # no `plan`
    assert ffi.ffi_drv_malformed_call(fn, None, 0x100000, 1 << 20, 1, err, 512) == K_INVALID and b"dtype" in err.value
    assert ffi.ffi_drv_malformed_call(fn, None, 0x100000, 1 << 20, 2, err, 512) == K_INVALID and b"attribute" in err.value


def test_pack_handler_validates_shapes(ffi):
    err = C.create_string_buffer(512)
    rc = ffi.ffi_drv_pack_call(_fn(ffi, "JaxentBvPack"), None, 0x1000, 0x2000, 16, 64, 32, 0x3000, 16 * 64 * 2, err, 512)
    assert rc == K_INVALID and b"n_residues" in err.value  # heavy (16, 64) vs acceptor (16, 32)
