"""BatchedBvEngine: B replicas of the BV MaxEnt optimise step on shared features (BASELINE config 5).

What the reference expresses as ``jax.vmap`` over ``_optimise_pure`` (opt/batch.py:113-146) -- runs that share the
features, the data and the starting point and differ in (forward_model_weights, forward_model_scaling,
learning_rate) -- is advanced here by one fused step for all replicas: the two feature sweeps are tcgen05 GEMMs
(csrc/bg_gemm.cu), the rest is elementwise over (frame, replica) plus one tail CTA per replica (csrc/bg_step.cu).
PyTorch only owns the device memory and the stream.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib as L
from .engine import BvEngine, OptSettings, SlotSpec, _stream_ptr

REPLICA_ALIGN = 16
MAX_REPLICAS = 256
RECORD_RING = 256


class GemmFeatures:
    """bf16 copy of the BV features in the UMMA operand layout (jaxent_bg_pack_features)."""

    format = 2  # JaxentBvProblem.feature_format of this layout

    def __init__(self, packed: torch.Tensor, n_residues: int, n_frames: int, exact: bool, row_off: Optional[torch.Tensor] = None):
        self.packed, self.R, self.F, self.exact = packed, int(n_residues), int(n_frames), bool(exact)
        self.row_off = row_off  # float32 [2][Rp]: the integer row constants the stored values were centred with

    @property
    def nbytes(self) -> int:
        return self.packed.numel()

    @property
    def shape(self):
        return (self.R, self.F)


def pack_gemm_features(heavy, acceptor, device, require_exact: bool = True) -> GemmFeatures:
    lib = L.load()
    dev = torch.device(device)
    with torch.cuda.device(dev):
        h = torch.as_tensor(np.asarray(heavy) if not isinstance(heavy, torch.Tensor) else heavy).to(dev, torch.float32).contiguous()
        a = torch.as_tensor(np.asarray(acceptor) if not isinstance(acceptor, torch.Tensor) else acceptor).to(dev, torch.float32).contiguous()
        R, F = int(h.shape[0]), int(h.shape[1])
        if tuple(a.shape) != (R, F):
            raise ValueError("Input features have different shapes")  # reference models/core.py:61-62
        out = torch.empty(lib.jaxent_bg_features_bytes(R, F), dtype=torch.uint8, device=dev)
        flag = torch.ones(1, dtype=torch.int32, device=dev)
        # centred storage N - round(row mean): still exact in bf16 for integer counts, and the backward column sums shrink from
        # R x mean to the fluctuation scale (jaxent_b200.h, jaxent_bg_pack_features_centred)
        Rp = (R + 127) // 128 * 128
        row_off = torch.zeros(2, Rp, dtype=torch.float32, device=dev)
        row_off[0, :R] = torch.round(h.mean(dim=1))
        row_off[1, :R] = torch.round(a.mean(dim=1))
        L.check(lib.jaxent_bg_pack_features_centred(h.data_ptr(), a.data_ptr(), row_off.data_ptr(), R, F, F, out.data_ptr(), flag.data_ptr(),
                                                    _stream_ptr()))
        exact = bool(flag.item())
    if require_exact and not exact:
        raise ValueError("the tensor-core sweep stores the features in bf16: values must be exactly representable "
                         "(hard-cutoff BV contact counts <= 256 are); use the sequential engine otherwise")
    return GemmFeatures(out, R, F, exact, row_off)


class BatchedBvEngine(BvEngine):
    """Same construction arguments as BvEngine plus the number of replicas; single GPU (config 5 is not sharded)."""

    def __init__(self, heavy, acceptor, k_ints, timepoints, slots: Sequence[SlotSpec], uptake: bool, settings: OptSettings,
                 n_replicas: int, prior=None, device: Optional[torch.device] = None, n_pieces: int = 3):
        if not torch.cuda.is_available():
            raise RuntimeError("jaxent_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if n_replicas < 1 or n_replicas > MAX_REPLICAS:
            raise ValueError(f"between 1 and {MAX_REPLICAS} replicas per batch")
        self.B = int(n_replicas)
        self.Bn = (self.B + REPLICA_ALIGN - 1) // REPLICA_ALIGN * REPLICA_ALIGN
        self.n_pieces = int(n_pieces)
        self._batched = True
        super().__init__(heavy, acceptor, k_ints, timepoints, slots, uptake, settings, prior=prior, device=device)

    # BvEngine.__init__ calls these two hooks ------------------------------------------------------------
    def _pack(self, heavy, acceptor, feature_dtype: str = "f32"):
        return pack_gemm_features(heavy, acceptor, self.device)

    def _create_plan(self, prob, cfg, sm):
        nbytes = self.lib.jaxent_bg_workspace_bytes(C.byref(prob), self.Bn, self.n_pieces, sm)
        if nbytes == 0:
            raise L.JaxentError(L.E_INVALID, self.lib.jaxent_last_error().decode() or "batched workspace size query failed")
        self.ws = torch.zeros(nbytes + 256, dtype=torch.uint8, device=self.device)
        self._ws_off = (-self.ws.data_ptr()) % 256
        plan = C.c_void_p()
        L.check(self.lib.jaxent_bg_plan_create(C.byref(prob), C.byref(cfg), self.Bn, self.n_pieces, self.ws.data_ptr() + self._ws_off,
                                               nbytes, sm, C.byref(plan)))
        feats = getattr(self, "features", None)
        if feats is not None and getattr(feats, "row_off", None) is not None:
            L.check(self.lib.jaxent_bg_plan_set_row_offsets(plan, feats.row_off.data_ptr()))
        return plan

    def __del__(self):
        try:
            if getattr(self, "plan", None):
                self.lib.jaxent_bg_plan_destroy(self.plan)
                self.plan = None
        except Exception:
            pass

    def _bview(self, which: int, dtype):
        ptr, nb = C.c_void_p(), C.c_size_t()
        L.check(self.lib.jaxent_bg_plan_buffer(self.plan, which, C.byref(ptr), C.byref(nb)))
        off = ptr.value - self.ws.data_ptr()
        return self.ws[off : off + nb.value].view(dtype)

    # ------------------------------------------------------------------------------------------------------
    def init_state(self, z0, bc: float, bh: float, fw, fs, lr=None) -> None:
        """fw, fs: (B, n_losses); lr: (B,) or None.  z0 / bc / bh are shared by all replicas (reference batch.py:93-126)."""
        n = len(self.kinds)
        fw = np.asarray(fw, np.float32).reshape(self.B, n)
        fs = np.asarray(fs, np.float32).reshape(self.B, n)
        pad = self.Bn - self.B
        fwp = np.zeros((self.Bn, L.MAX_SLOTS), np.float32)
        fsp = np.zeros((self.Bn, L.MAX_SLOTS), np.float32)
        fwp[: self.B, :n], fsp[: self.B, :n] = fw, fs
        if pad:  # padded replicas repeat the last one (reference _pad_array, batch.py:25-29)
            fwp[self.B :], fsp[self.B :] = fwp[self.B - 1], fsp[self.B - 1]
        lrp = None
        if lr is not None:
            lr = np.asarray(lr, np.float32).reshape(self.B)
            lrp = np.concatenate([lr, np.repeat(lr[-1:], pad)]).astype(np.float32)
        with torch.cuda.device(self.device):
            z0 = self._dev_f32(z0).reshape(-1)
            if z0.numel() != self.F:
                raise ValueError("frame_weights must have one entry per frame")
            zmax = float(z0.max().item())
            L.check(self.lib.jaxent_bg_state_init(
                self.plan, z0.data_ptr(), zmax, float(bc), float(bh),
                fwp.ctypes.data_as(C.POINTER(C.c_float)), fsp.ctypes.data_as(C.POINTER(C.c_float)),
                lrp.ctypes.data_as(C.POINTER(C.c_float)) if lrp is not None else None, _stream_ptr()))
        self.steps_done = 0
        self.z = self._bview(0, torch.float32).view(self.F, self.Bn)
        self.e = self._bview(1, torch.float32).view(self.F, self.Bn)
        self._brecords = self._bview(2, torch.uint8)
        self._bctl = self._bview(3, torch.uint8)
        self.log_pf_b = self._bview(4, torch.float32).view(self.Bn, self.R)
        self.uptake_b = self._bview(5, torch.float32).view(self.Bn, max(self.T, 1), self.R) if self.uptake else None

    # ------------------------------------------------------------------ on-device loop of every replica
    def enable_tracking(self, convergence, ema_alpha: float = 0.5, min_steps_per_threshold: int = 2, tolerance: float = 1e-10,
                        loop: str = "python") -> None:
        """Arm the per-replica tracker; call before init_state().  The thresholds are sorted descending here and scaled by
        each replica's learning rate on the device side.  loop="python": ConvergenceTracker of `_optimise`
        (reference opt/track.py:128-197, opt/run.py:285-348), a snapshot per completed threshold.  loop="pure": the functional
        tracker of `_optimise_pure` (opt/track.py:40-125, opt/run.py:141-232), the loop the reference's batch_optimise vmaps --
        every step is a history entry, snapshot slot 0 follows the best one (jaxent_bg_track_mode, jaxent_b200.h)."""
        if loop not in ("python", "pure"):
            raise ValueError("loop must be 'python' or 'pure'")
        conv = np.sort(np.atleast_1d(np.asarray(convergence, np.float32)))[::-1].copy()
        n = int(conv.shape[0])
        if n < 1 or n > L.TRACK_MAX_THRESHOLDS:
            raise ValueError(f"between 1 and {L.TRACK_MAX_THRESHOLDS} convergence thresholds are supported on the device")
        with torch.cuda.device(self.device):
            self._trk_ema = torch.zeros((self.F, self.Bn), dtype=torch.float32, device=self.device)
            self._trk_snap_w = torch.zeros((n + 1, self.F, self.Bn), dtype=torch.float32, device=self.device)
            self._trk_snap_ema = torch.zeros((n + 1, self.F, self.Bn), dtype=torch.float32, device=self.device)
            self._trk_snaps = torch.zeros((n + 1) * self.Bn * C.sizeof(L.TrackSnapshot), dtype=torch.uint8, device=self.device)
            self._trk_status = torch.zeros(self.Bn * C.sizeof(L.TrackStatus), dtype=torch.uint8, device=self.device)
        arr = (C.c_float * n)(*[float(t) for t in conv])
        L.check(self.lib.jaxent_bg_track_init(self.plan, arr, n, float(ema_alpha), int(min_steps_per_threshold), float(tolerance),
                                              self._trk_ema.data_ptr(), self._trk_snap_w.data_ptr(), self._trk_snap_ema.data_ptr(),
                                              self._trk_snaps.data_ptr()))
        L.check(self.lib.jaxent_bg_track_mode(self.plan, 1 if loop == "pure" else 0))
        self._trk_n = n
        self._trk_loop = loop
        self.tracking = True

    def track_status(self):
        """[TrackStatus] per replica (device -> host read: the only host synchronisation of the loop)."""
        with torch.cuda.device(self.device):
            L.check(self.lib.jaxent_bg_track_status(self.plan, self._trk_status.data_ptr(), _stream_ptr()))
            host = self._trk_status.cpu().numpy().tobytes()
        sz = C.sizeof(L.TrackStatus)
        return [L.TrackStatus.from_buffer_copy(host[i * sz : (i + 1) * sz]) for i in range(self.B)]

    def track_snapshots(self, b: int, n: int):
        """[(TrackSnapshot, weights[F] view, ema_weights[F] view or None)] for the first n snapshots of replica b."""
        sz = C.sizeof(L.TrackSnapshot)
        if getattr(self, "_trk_snaps_host", None) is None:
            self._trk_snaps_host = self._trk_snaps.cpu().numpy()
        out = []
        for i in range(n):
            off = (i * self.Bn + b) * sz
            sn = L.TrackSnapshot.from_buffer_copy(self._trk_snaps_host[off : off + sz].tobytes())
            out.append((sn, self._trk_snap_w[i, :, b], self._trk_snap_ema[i, :, b] if sn.have_ema else None))
        return out

    def run(self, n_steps: int, chunk: int = 32, keep_records: bool = False):
        """Up to n_steps steps; replicas whose loop has ended are frozen on the device, the host looks at the status every
        `chunk` steps and stops when every replica has ended.  Returns [TrackStatus].  keep_records: the loss records of
        every step are drained from the device ring with each status read (one small copy per chunk) and kept in
        `self.record_history`, a (steps + 1, Bn) array of LossRecord bytes: row k = the losses evaluated at optimiser step k
        (the reference's per-step history_losses, opt/optimiser.py:555-567), valid for k <= the replica's steps_done."""
        if not getattr(self, "tracking", False):
            raise RuntimeError("enable_tracking() has not been called")
        chunk = max(1, min(int(chunk), RECORD_RING - 1))
        done, st = 0, None
        self._trk_snaps_host = None
        rows = [self._record_rows(0, 1)] if keep_records else []  # row 0: the initial forward (jaxent_bg_state_init)
        while done < n_steps:
            m = min(chunk, n_steps - done)
            self.step(m)
            if keep_records:  # row k is written by the tail of step k - 1; rows beyond a frozen replica's steps_done are stale
                rows.append(self._record_rows(done + 1, m))
            done += m
            st = self.track_status()
            if all(s.stop for s in st):
                break
        if keep_records:
            sz = C.sizeof(L.LossRecord)
            self.record_history = np.concatenate(rows, axis=0) if rows else np.zeros((0, self.Bn, sz), np.uint8)
        return st if st is not None else self.track_status()

    def _record_rows(self, first: int, n: int) -> np.ndarray:
        """(n, Bn, sizeof(LossRecord)) bytes: ring rows of optimiser steps first .. first + n - 1"""
        sz = C.sizeof(L.LossRecord)
        ring = self._brecords.view(RECORD_RING, self.Bn * sz)
        idx = torch.arange(first, first + n, device=self.device) % RECORD_RING
        return ring[idx].cpu().numpy().reshape(n, self.Bn, sz)

    def record_of(self, step: int, b: int):
        """LossRecord of replica b at optimiser step `step` from the history kept by run(keep_records=True)"""
        return L.LossRecord.from_buffer_copy(self.record_history[step, b].tobytes())

    def step(self, n: int = 1) -> None:
        with torch.cuda.device(self.device):
            L.check(self.lib.jaxent_bg_steps(self.plan, int(n), _stream_ptr()))  # one call: consecutive steps overlap (jaxent_b200.h)
        self.steps_done += n
        self.launches += 9 * n

    def capture(self, steps_per_graph: int = 2) -> None:
        raise NotImplementedError("the batched step is launched eagerly")

    def records_at(self, step: int):
        """LossRecord of every replica at optimiser step `step` (must be within the last 256 steps)."""
        sz = C.sizeof(L.LossRecord)
        row = (step % RECORD_RING) * self.Bn * sz
        host = self._brecords[row : row + self.B * sz].cpu().numpy().tobytes()
        return [L.LossRecord.from_buffer_copy(host[i * sz : (i + 1) * sz]) for i in range(self.B)]

    def ctl_field(self, field: int, dtype) -> np.ndarray:
        sz, off = self.lib.jaxent_bg_ctl_sizeof(), self.lib.jaxent_bg_ctl_offset(field)
        host = self._bctl.cpu().numpy().reshape(self.Bn, sz)
        width = np.dtype(dtype).itemsize
        return host[: self.B, off : off + width].copy().view(dtype).reshape(-1)

    def weights_all(self) -> torch.Tensor:
        """softmax frame weights of every replica, (F, B) float32 (save_state.params.frame_weights)."""
        S = torch.from_numpy(self.ctl_field(0, np.float64)).to(self.device)
        return self.e[:, : self.B] / S.to(torch.float32)[None, :]

    def gradients_all(self) -> torch.Tensor:
        """masked gradients d loss / d logits of the update applied last, (F, B) float32 (state.gradients)."""
        sel = int(self._bview(10, torch.int32).cpu().item())
        return self._bview(8 + sel, torch.float32).view(self.F, self.Bn)[:, : self.B]

    def algorithmic_flops_per_step(self) -> int:
        """2*(2*R*F*B) forward + 2*(2*R*F*B) backward, times the number of bf16 pieces."""
        return 8 * self.R * self.F * self.B * self.n_pieces
