"""ctypes binding of the C ABI declared in include/jaxent_b200.h.

This is the stand-in for the jax.ffi registration of the target deployment:
the same entry points, driven with raw device pointers and a CUDA stream handle.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _build

MAX_SLOTS = 6
RECORD_RING = 4096

LOSS_UPTAKE_EYE_MSE = 0
LOSS_UPTAKE_MC_EYE_MSE = 1
LOSS_UPTAKE_SIGMA_MSE = 2
LOSS_PF_L2 = 3
LOSS_MAXENT_CONVEX_KL = 4
LOSS_PARAMS_L1 = 5
LOSS_PARAMS_L2 = 6

BUF_RED, BUF_KLSUM, BUF_WEIGHTS, BUF_RECORDS, BUF_LOGPF, BUF_UPTAKE, BUF_AVG = 0, 1, 2, 4, 5, 6, 7
BUF_XCHG_ERR = 9
MAX_RANKS = 8
IPC_HANDLE_BYTES = 64

E_INVALID, E_UNSUPPORTED, E_CUDA, E_WORKSPACE = -1, -2, -3, -4


class PeptideMap(C.Structure):
    _fields_ = [
        ("n_pep", C.c_int32),
        ("nnz", C.c_int32),
        ("row_ptr", C.c_void_p),
        ("col", C.c_void_p),
        ("val", C.c_void_p),
        ("t_ptr", C.c_void_p),
        ("t_row", C.c_void_p),
        ("t_val", C.c_void_p),
        ("y", C.c_void_p),
        ("W", C.c_void_p),
    ]


class LossSlot(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("train", PeptideMap),
        ("val", PeptideMap),
        ("prior_bc", C.c_float),
        ("prior_bh", C.c_float),
    ]


class BvProblem(C.Structure):
    _fields_ = [
        ("R", C.c_int32),
        ("T", C.c_int32),
        ("uptake", C.c_int32),
        ("n_slots", C.c_int32),
        ("F", C.c_int64),
        ("F_total", C.c_int64),
        ("packed", C.c_void_p),
        ("k_ints", C.c_void_p),
        ("timepoints", C.c_void_p),
        ("prior", C.c_void_p),
        ("prior_norm", C.c_double),
        ("slots", LossSlot * MAX_SLOTS),
        ("feature_format", C.c_int32),
        ("pad_", C.c_int32),
    ]


class OptConfig(C.Structure):
    _fields_ = [
        ("learning_rate", C.c_float),
        ("initial_learning_rate", C.c_float),
        ("initial_steps", C.c_int32),
        ("plateau_denominator", C.c_float),
        ("model_parameters_lr_scale", C.c_float),
        ("clip_value", C.c_float),
        ("optimise_frame_weights", C.c_int32),
        ("optimise_model_parameters", C.c_int32),
        ("b1", C.c_double),
        ("b2", C.c_double),
        ("eps", C.c_double),
    ]


class LossRecord(C.Structure):
    _fields_ = [
        ("step", C.c_int32),
        ("branch", C.c_int32),
        ("lr", C.c_float),
        ("grad_dot", C.c_float),
        ("bc", C.c_float),
        ("bh", C.c_float),
        ("total_train", C.c_float),
        ("total_val", C.c_float),
        ("train", C.c_float * MAX_SLOTS),
        ("val", C.c_float * MAX_SLOTS),
        ("scaled_train", C.c_float * MAX_SLOTS),
        ("scaled_val", C.c_float * MAX_SLOTS),
        ("grad_bc", C.c_float),
        ("grad_bh", C.c_float),
        ("complete", C.c_int32),
        ("pad", C.c_int32),
    ]


class TrackSnapshot(C.Structure):
    _fields_ = [
        ("loop_step", C.c_int32),
        ("state_step", C.c_int32),
        ("have_ema", C.c_int32),
        ("threshold_idx", C.c_int32),
        ("bc", C.c_float),
        ("bh", C.c_float),
        ("ema_bc", C.c_float),
        ("ema_bh", C.c_float),
        ("losses", LossRecord),
    ]


class TrackStatus(C.Structure):
    _fields_ = [
        ("stop", C.c_int32),
        ("reason", C.c_int32),
        ("stop_step", C.c_int32),
        ("n_snapshots", C.c_int32),
        ("steps_done", C.c_int32),
        ("threshold_idx", C.c_int32),
        ("last_loss", C.c_float),
        ("ema_loss_delta", C.c_float),
    ]


TRACK_MAX_THRESHOLDS = 8

# every symbol include/jaxent_b200.h declares: name -> (restype, argtypes)
_VP, _I32, _I64, _SZ, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t, C.c_float
SYMBOLS = {
    "jaxent_last_error": (C.c_char_p, []),
    "jaxent_version": (C.c_int, []),
    "jaxent_abi_sizeof": (C.c_int, [C.c_int]),
    "jaxent_bv_packed_bytes": (_SZ, [_I32, _I64]),
    "jaxent_bv_pack_features_f32": (C.c_int, [_VP, _VP, _I32, _I64, _I64, _I64, _I64, _VP, _VP]),
    "jaxent_bv_packed_bytes_fmt": (_SZ, [_I32, _I64, _I32]),
    "jaxent_bv_features_fit_u8": (C.c_int, [_VP, _VP, _I32, _I64, _I64, _VP, _VP]),
    "jaxent_bv_pack_features_u8": (C.c_int, [_VP, _VP, _I32, _I64, _I64, _VP, _VP]),
    "jaxent_bv_prior_partial": (C.c_int, [_VP, _I64, C.c_double, _VP, _VP]),
    "jaxent_bv_max": (C.c_int, [_VP, _I64, _VP, _VP]),
    "jaxent_bv_workspace_bytes": (_SZ, [C.POINTER(BvProblem)]),
    "jaxent_bv_plan_create": (C.c_int, [C.POINTER(BvProblem), C.POINTER(OptConfig), _VP, _SZ, _I32, C.POINTER(_VP)]),
    "jaxent_bv_plan_destroy": (None, [_VP]),
    "jaxent_bv_plan_buffer": (C.c_int, [_VP, _I32, C.POINTER(_VP), C.POINTER(_SZ)]),
    "jaxent_bv_exchange_bytes": (_SZ, [_I32, _I32]),
    "jaxent_bv_exchange_bytes_for": (_SZ, [C.POINTER(BvProblem), _I32]),
    "jaxent_peer_alloc": (C.c_int, [_SZ, C.POINTER(_VP), C.c_char_p]),
    "jaxent_peer_open": (C.c_int, [C.c_char_p, C.POINTER(_VP)]),
    "jaxent_peer_close": (C.c_int, [_VP]),
    "jaxent_peer_free": (C.c_int, [_VP]),
    "jaxent_bv_plan_set_exchange": (C.c_int, [_VP, _I32, _I32, C.POINTER(_VP)]),
    "jaxent_bv_exchange_error": (C.c_int, [_VP, _VP, _VP]),
    "jaxent_bv_contacts": (C.c_int, [_VP, _I64, _I32, _VP, _VP, _I32, _VP, _VP, _I32, _I32, _I32, _F, _I32, _VP, _VP, _VP]),
    "jaxent_bv_contacts_cell": (C.c_int, [_VP, _I64, _I32, _VP, _VP, _I32, _VP, _VP, _I32, _I32, _I32, _F, _I32, _VP, _I32, _VP, _VP]),
    "jaxent_bv_predict_frames": (C.c_int, [_VP, _I32, _I64, _I32, _VP, _VP, _F, _F, _VP, _VP]),
    "jaxent_bg_features_bytes": (_SZ, [_I32, _I64]),
    "jaxent_bg_pack_features": (C.c_int, [_VP, _VP, _I32, _I64, _I64, _VP, _VP, _VP]),
    "jaxent_bg_pack_features_centred": (C.c_int, [_VP, _VP, _VP, _I32, _I64, _I64, _VP, _VP, _VP]),
    "jaxent_bg_plan_set_row_offsets": (C.c_int, [_VP, _VP]),
    "jaxent_bg_gemm1": (C.c_int, [_VP, _VP, _VP, _I32, _I32, _I32, _I32, _I32, _VP]),
    "jaxent_bg_gemm2": (C.c_int, [_VP, _VP, _VP, _I32, _I32, _I32, _I32, _I32, _VP]),
    "jaxent_bg_workspace_bytes": (_SZ, [C.POINTER(BvProblem), _I32, _I32, _I32]),
    "jaxent_bg_plan_create": (C.c_int, [C.POINTER(BvProblem), C.POINTER(OptConfig), _I32, _I32, _VP, _SZ, _I32, C.POINTER(_VP)]),
    "jaxent_bg_plan_destroy": (None, [_VP]),
    "jaxent_bg_state_init": (C.c_int, [_VP, _VP, _F, _F, _F, C.POINTER(_F), C.POINTER(_F), C.POINTER(_F), _VP]),
    "jaxent_bg_step": (C.c_int, [_VP, _VP]),
    "jaxent_bg_steps": (C.c_int, [_VP, _I32, _VP]),
    "jaxent_bg_step_phase": (C.c_int, [_VP, _I32, _VP]),
    "jaxent_bg_plan_buffer": (C.c_int, [_VP, _I32, C.POINTER(_VP), C.POINTER(_SZ)]),
    "jaxent_bg_track_init": (C.c_int, [_VP, C.POINTER(_F), _I32, _F, _I32, _F, _VP, _VP, _VP, _VP]),
    "jaxent_bg_track_mode": (C.c_int, [_VP, _I32]),
    "jaxent_bg_track_status": (C.c_int, [_VP, _VP, _VP]),
    "jaxent_bg_ctl_sizeof": (C.c_int, []),
    "jaxent_bg_ctl_offset": (C.c_int, [_I32]),
    "jaxent_handle_register": (_I64, [_VP, _I32]),
    "jaxent_handle_lookup": (_VP, [_I64, _I32]),
    "jaxent_handle_release_ptr": (None, [_VP]),
    "jaxent_bv_plan_workspace": (C.c_int, [_VP, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "jaxent_bv_plan_rebind_workspace": (C.c_int, [_VP, _VP, _SZ]),
    "jaxent_bg_plan_workspace": (C.c_int, [_VP, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "jaxent_bv_state_init": (C.c_int, [_VP, _VP, _F, _F, _F, C.POINTER(_F), C.POINTER(_F), _VP]),
    "jaxent_bv_pass": (C.c_int, [_VP, _I32, _VP]),
    "jaxent_bv_reduce": (C.c_int, [_VP, _VP]),
    "jaxent_bv_tail": (C.c_int, [_VP, _VP]),
    "jaxent_bv_finish_record": (C.c_int, [_VP, _VP]),
    "jaxent_bv_forward_init": (C.c_int, [_VP, _VP]),
    "jaxent_bv_step": (C.c_int, [_VP, _VP]),
    "jaxent_bv_track_init": (C.c_int, [_VP, C.POINTER(_F), _I32, _F, _I32, _F, _VP, _VP, _VP, _VP]),
    "jaxent_bv_track_status": (C.c_int, [_VP, _VP, _VP]),
    "jaxent_bv_export_state": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, _VP]),
}


class JaxentError(RuntimeError):
    """Raised for any non-zero return of the C ABI (mirrors the reference's ValueError /
    RuntimeError conventions: models/core.py:65-66,102,122; opt/run.py:350-353,405-406)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"[jaxent_b200 {code}] {msg}")
        self.code = code


_lib = None


def lib_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load libjaxent_b200.so (building it in-tree with nvcc if needed).  There is no fallback:
    if the library cannot be built or loaded this raises."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    override = os.environ.get("JAXENT_B200_LIB")  # A/B experiments against another build of the same ABI
    if override:
        path, build_if_missing = override, False
    if build_if_missing and _build.needs_build():
        try:
            _build.build()
        except Exception:
            if not os.path.exists(path):
                raise
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing and could not be built; jaxent_b200 has no CPU fallback")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        if override and not hasattr(lib, name):
            continue
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise JaxentError(rc, load().jaxent_last_error().decode("utf-8", "replace"))
