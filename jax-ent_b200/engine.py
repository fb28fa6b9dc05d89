"""BvEngine: owns the device buffers of one rank's shard and drives the C ABI.

PyTorch is used only as plumbing here (device memory, streams, CUDA graphs, torch.distributed);
every arithmetic kernel on the path lives in libjaxent_b200.so.  Frame sharding: each rank
holds a contiguous block of frames; per step there are two small all-reduces (NCCL, sum):
the (4R+4)-double partial-sum vector after the pass and the 4-double MaxEnt vector after the
tail (include/jaxent_b200.h, "the step, in the three phases").
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from . import _lib as L
from .sharding import PeerExchange, ShardComm


@dataclass
class SlotSpec:
    """One (loss function, data target) pair; dense maps are converted to CSR/CSC on upload."""

    kind: int
    S_train: Optional[np.ndarray] = None
    y_train: Optional[np.ndarray] = None
    S_val: Optional[np.ndarray] = None
    y_val: Optional[np.ndarray] = None
    W_train: Optional[np.ndarray] = None
    W_val: Optional[np.ndarray] = None
    prior_bc: float = 0.35
    prior_bh: float = 2.0


@dataclass
class OptSettings:
    """OptaxOptimizer.__init__ arguments (reference opt/optimiser.py:68-82)."""

    learning_rate: float = 1e-4
    initial_learning_rate: float = 1.0
    initial_steps: int = 0
    plateau_denominator: float = 1.005
    model_parameters_lr_scale: float = 1.0
    clip_value: Optional[float] = 1.0
    optimise_frame_weights: bool = True
    optimise_model_parameters: bool = False
    b1: float = 0.9
    b2: float = 0.999
    eps: float = 1e-8

    def to_c(self) -> L.OptConfig:
        return L.OptConfig(
            self.learning_rate,
            self.initial_learning_rate,
            int(self.initial_steps),
            self.plateau_denominator,
            self.model_parameters_lr_scale,
            -1.0 if self.clip_value is None else float(self.clip_value),
            int(self.optimise_frame_weights),
            int(self.optimise_model_parameters),
            self.b1,
            self.b2,
            self.eps,
        )


def _csr(S: np.ndarray):
    S = np.asarray(S, np.float32)
    rows, cols = np.nonzero(S)
    order = np.lexsort((cols, rows))
    rows, cols = rows[order], cols[order]
    ptr = np.zeros(S.shape[0] + 1, np.int32)
    np.add.at(ptr, rows + 1, 1)
    return np.cumsum(ptr).astype(np.int32), cols.astype(np.int32), S[rows, cols].astype(np.float32)


def _stream_ptr(stream: Optional[torch.cuda.Stream] = None) -> int:
    s = stream if stream is not None else torch.cuda.current_stream()
    return int(s.cuda_stream)


class PackedFeatures:
    """The packed (tile layout) copy of one rank's BV features; can be shared by several engines
    (e.g. the optimiser's engine and the forward-only engine behind ``Simulation.forward``)."""

    def __init__(self, packed: torch.Tensor, n_residues: int, n_frames: int, fmt: int = 0):
        self.packed, self.R, self.F, self.format = packed, int(n_residues), int(n_frames), int(fmt)

    @property
    def nbytes(self) -> int:
        return self.packed.numel() * self.packed.element_size()

    @property
    def shape(self):
        return (self.R, self.F)


class BvEngine:
    def __init__(
        self,
        heavy,
        acceptor,
        k_ints,
        timepoints,
        slots: Sequence[SlotSpec],
        uptake: bool,
        settings: OptSettings,
        prior=None,
        device: Optional[torch.device] = None,
        F_total: Optional[int] = None,
        group=None,
        feature_dtype: str = "f32",
        exchange: str = "auto",
        average_first: bool = True,
    ):
        if not torch.cuda.is_available():
            raise RuntimeError("jaxent_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = L.load()
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.group = group
        self.comm = ShardComm(group)
        self.sharded = self.comm.active
        self.settings = settings
        self.uptake = bool(uptake)
        # ForwardPass.average_first (reference models/core.py:299): False = uptake per frame, then frame-averaged
        self.average_first = bool(average_first) or not self.uptake
        self._keep = []  # device tensors referenced by raw pointer from the plan

        R, F = int(heavy.shape[0]), int(heavy.shape[1])
        if not hasattr(heavy, "packed") and tuple(acceptor.shape) != (R, F):
            raise ValueError("Input features have different shapes")  # reference models/core.py:61-62
        self.R, self.F = R, F
        self.F_total = int(F_total) if F_total is not None else F
        self.T = int(np.asarray(timepoints).reshape(-1).shape[0]) if self.uptake else 0

        with torch.cuda.device(self.device):
            staged = None
            if isinstance(heavy, PackedFeatures) or hasattr(heavy, "packed"):
                self.features = heavy
            elif getattr(self, "_batched", False) or feature_dtype != "f32":
                self.features = self._pack(heavy, acceptor, feature_dtype)
            else:
                # host -> device copies of the feature arrays start now, on a side stream; the set-up work that does not need
                # them (CSR conversion and upload of the peptide maps, prior normalisation, the peer-exchange handshake of a
                # frame-sharded engine) overlaps the transfer, the pack kernel waits for it
                staged = self._upload(heavy, acceptor)
                self.features = self._alloc_packed(R, F, feature_dtype)
            self.packed = self.features.packed
            prob = L.BvProblem()
            prob.R, prob.T, prob.n_slots = R, self.T, len(slots)
            prob.uptake = (1 if self.average_first else 2) if self.uptake else 0
            prob.F, prob.F_total = F, self.F_total
            prob.packed = self.packed.data_ptr()
            prob.feature_format = self.features.format
            if self.uptake:
                if k_ints is None:
                    raise ValueError("BV_uptake_ForwardPass needs k_ints")
                self.k_ints = self._dev(np.asarray(k_ints, np.float32).reshape(-1))
                self.timepoints = self._dev(np.asarray(timepoints, np.float32).reshape(-1))
                if self.k_ints.numel() != R:
                    raise ValueError("k_ints must have one entry per residue")
                prob.k_ints, prob.timepoints = self.k_ints.data_ptr(), self.timepoints.data_ptr()
            if len(slots) > L.MAX_SLOTS:
                raise ValueError(f"at most {L.MAX_SLOTS} loss slots are supported")
            self.kinds = [int(s.kind) for s in slots]
            has_kl = L.LOSS_MAXENT_CONVEX_KL in self.kinds
            if has_kl:
                if prior is None:
                    raise ValueError("maxent_convexKL_loss needs the prior frame weights")
                self.prior = self._dev_f32(prior).reshape(-1)
                if self.prior.numel() != F:
                    raise ValueError("prior frame weights must match n_frames")
                eps = 1e-10 / self.F_total
                tmp = torch.zeros(1, dtype=torch.float64, device=self.device)
                L.check(self.lib.jaxent_bv_prior_partial(self.prior.data_ptr(), F, eps, tmp.data_ptr(), _stream_ptr()))
                self.comm.sum_(tmp)
                prob.prior = self.prior.data_ptr()
                prob.prior_norm = float(tmp.item())
            for i, s in enumerate(slots):
                prob.slots[i].kind = int(s.kind)
                prob.slots[i].prior_bc, prob.slots[i].prior_bh = float(s.prior_bc), float(s.prior_bh)
                if s.kind <= L.LOSS_PF_L2:
                    self._fill_map(prob.slots[i].train, s.S_train, s.y_train, s.W_train, s.kind)
                    self._fill_map(prob.slots[i].val, s.S_val, s.y_val, s.W_val, s.kind)
            self.prob = prob
            cfg = settings.to_c()
            sm = torch.cuda.get_device_properties(self.device).multi_processor_count
            if getattr(self, "_batched", False):  # BatchedBvEngine (batch_engine.py): same problem description, batched plan
                self.plan = self._create_plan(prob, cfg, sm)
                self.steps_done = 0
                self.launches = 0
                self.peer, self.exchange, self.fused = None, "none", False
                return
            nbytes = self.lib.jaxent_bv_workspace_bytes(C.byref(prob))
            if nbytes == 0:
                raise L.JaxentError(L.E_INVALID, "workspace size query failed")
            self.ws = torch.zeros(nbytes + 256, dtype=torch.uint8, device=self.device)
            self._ws_off = (-self.ws.data_ptr()) % 256
            plan = C.c_void_p()
            L.check(
                self.lib.jaxent_bv_plan_create(
                    C.byref(prob), C.byref(cfg), self.ws.data_ptr() + self._ws_off, nbytes, sm, C.byref(plan)
                )
            )
            self.plan = plan
            # frame-sharded: fuse the two per-step exchanges into the kernels over NVLink peer memory when the
            # exchange buffers can be mapped ("nvlink"); otherwise NCCL all-reduces between the phases ("nccl")
            if exchange not in ("auto", "nvlink", "nccl"):
                raise ValueError("exchange must be 'auto', 'nvlink' or 'nccl'")
            self.peer = None
            self.exchange = "none"
            if self.sharded:
                self.exchange = "nccl"
                if exchange in ("auto", "nvlink"):
                    try:
                        self.peer = PeerExchange.acquire(self.lib, group, prob, self.device)
                        self.peer.attach(self.plan)
                        self.exchange = "nvlink"
                    except RuntimeError:
                        if exchange == "nvlink":
                            raise
            if staged is not None:
                self._pack_staged(staged, feature_dtype)
            self.fused = self.exchange == "nvlink"
            self.red = self._view(L.BUF_RED, torch.float64)
            self.klsum = self._view(L.BUF_KLSUM, torch.float64)
            self.weights = self._view(L.BUF_WEIGHTS, torch.float32)
            self.log_pf = self._view(L.BUF_LOGPF, torch.float32)
            self.avg = self._view(L.BUF_AVG, torch.float32)
            self.uptake_out = self._view(L.BUF_UPTAKE, torch.float32).view(self.T, R) if self.uptake else None
            self._records = self._view(L.BUF_RECORDS, torch.uint8)
        self.steps_done = 0
        self._graph = None
        self._graph_steps = 0
        self.launches = 0

    # ------------------------------------------------------------------ plumbing
    def close(self, barrier: bool = True) -> None:
        """Releases the plan and the peer-mapped exchange buffers.  Frame-sharded engines: call this on EVERY rank (it is a
        collective when `barrier` is true) -- a rank must not free its inbox while a slower peer's last tail kernel may still
        store its MaxEnt words into it, so the ranks synchronise their devices and meet at a barrier first."""
        peer = getattr(self, "peer", None)
        if peer is not None:
            torch.cuda.synchronize(self.device)
            if barrier and self.group is not None and dist.is_available() and dist.is_initialized():
                dist.barrier(group=self.group)
        if getattr(self, "plan", None):
            self.lib.jaxent_bv_plan_destroy(self.plan)
            self.plan = None
        if peer is not None:
            if barrier:
                peer.release()  # the ranks have synchronised: the mapping goes back to the pool (PeerExchange.acquire)
            else:
                peer.in_use = True  # cannot prove the peers are done with it: never reuse, freed at process exit
            self.peer = None

    def __del__(self):
        try:  # best effort, never a collective: an explicit close() is the safe path for sharded engines
            self.close(barrier=False)
        except Exception:
            pass

    def _dev(self, a: np.ndarray) -> torch.Tensor:
        t = torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
        self._keep.append(t)
        return t

    def _dev_f32(self, a) -> torch.Tensor:
        if isinstance(a, torch.Tensor):
            t = a.to(self.device, torch.float32).contiguous()
        else:
            t = torch.from_numpy(np.ascontiguousarray(np.asarray(a, np.float32))).to(self.device)
        self._keep.append(t)
        return t

    def _view(self, which: int, dtype) -> torch.Tensor:
        ptr, nb = C.c_void_p(), C.c_size_t()
        L.check(self.lib.jaxent_bv_plan_buffer(self.plan, which, C.byref(ptr), C.byref(nb)))
        off = ptr.value - self.ws.data_ptr()
        if off < 0 or off + nb.value > self.ws.numel():
            return None  # lives in the peer-mapped exchange buffer, not in the workspace (fused exchange)
        return self.ws[off : off + nb.value].view(dtype)

    def _pack(self, heavy, acceptor, feature_dtype: str = "f32") -> PackedFeatures:
        """(R,F) fp32 -> packed tile layout (csrc/bv_pass.cu pack kernels); replaces cast_to_jax
        (reference custom_types/features.py:82-89).  feature_dtype: "f32" keeps fp32 values; "u8" stores the
        features as bytes (lossless for integer contact counts <= 255, raises otherwise); "auto" picks u8 when
        it is lossless."""
        if feature_dtype not in ("f32", "u8", "auto"):
            raise ValueError("feature_dtype must be 'f32', 'u8' or 'auto'")
        R, F = int(heavy.shape[0]), int(heavy.shape[1])
        h = self._to_dev_matrix(heavy)
        a = self._to_dev_matrix(acceptor)
        fmt = 0
        if feature_dtype in ("u8", "auto"):
            flag = torch.ones(1, dtype=torch.int32, device=self.device)
            L.check(self.lib.jaxent_bv_features_fit_u8(h.data_ptr(), a.data_ptr(), R, F, F, flag.data_ptr(), _stream_ptr()))
            fits = bool(flag.item())
            if feature_dtype == "u8" and not fits:
                raise ValueError("feature_dtype='u8' needs integer-valued features in [0, 255] (hard-cutoff BV contacts)")
            fmt = 1 if fits else 0
        if fmt == 1:
            nbytes = self.lib.jaxent_bv_packed_bytes_fmt(R, F, 1)
            packed = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            L.check(self.lib.jaxent_bv_pack_features_u8(h.data_ptr(), a.data_ptr(), R, F, F, packed.data_ptr(), _stream_ptr()))
        else:
            nbytes = self.lib.jaxent_bv_packed_bytes(R, F)
            packed = torch.empty(nbytes // 4, dtype=torch.float32, device=self.device)
            L.check(self.lib.jaxent_bv_pack_features_f32(h.data_ptr(), a.data_ptr(), R, F, F, 0, F, packed.data_ptr(), _stream_ptr()))
        torch.cuda.current_stream().synchronize()
        del h, a
        return PackedFeatures(packed, R, F, fmt)

    def _upload(self, heavy, acceptor):
        """starts the host -> device transfer of both (R,F) arrays on a side stream; -> (heavy_dev, acceptor_dev, event)"""
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            h = self._to_dev_matrix(heavy)
            a = self._to_dev_matrix(acceptor)
            ev = torch.cuda.Event()
            ev.record(side)
        return h, a, ev, side

    def _alloc_packed(self, R: int, F: int, feature_dtype: str) -> PackedFeatures:
        nbytes = self.lib.jaxent_bv_packed_bytes(R, F)
        return PackedFeatures(torch.empty(nbytes // 4, dtype=torch.float32, device=self.device), R, F, 0)

    def _pack_staged(self, staged, feature_dtype: str) -> None:
        h, a, ev, side = staged
        torch.cuda.current_stream().wait_event(ev)
        h.record_stream(torch.cuda.current_stream())
        a.record_stream(torch.cuda.current_stream())
        R, F = self.R, self.F
        L.check(self.lib.jaxent_bv_pack_features_f32(h.data_ptr(), a.data_ptr(), R, F, F, 0, F, self.packed.data_ptr(), _stream_ptr()))
        # no host synchronisation here: the pack kernel is stream-ordered before everything that reads the packed features, the
        # staging buffers are kept alive by record_stream, and the caller's host-side set-up keeps overlapping the transfer
        del h, a

    def _to_dev_matrix(self, x) -> torch.Tensor:
        if isinstance(x, torch.Tensor):
            return x.to(self.device, torch.float32, non_blocking=True).contiguous()
        x = np.asarray(x)
        if x.dtype != np.float32 or not x.flags.c_contiguous:
            x = np.ascontiguousarray(x, np.float32)
        return torch.from_numpy(x).to(self.device, non_blocking=True)

    def _fill_map(self, m: L.PeptideMap, S, y, W, kind: int) -> None:
        if S is None or y is None:
            raise ValueError("data loss slot needs a peptide map and targets for train and val")
        S = np.asarray(S, np.float32)
        if S.ndim != 2 or S.shape[1] != self.R:
            raise ValueError(f"peptide map must be (n_peptides, {self.R})")
        P = S.shape[0]
        y = np.asarray(y, np.float32)
        if kind == L.LOSS_PF_L2:
            y = y.reshape(-1)
            if y.shape[0] != P:
                raise ValueError("y_true must have one entry per peptide")
        else:
            y = y.reshape(P, -1)  # Dataset.y_true is (P,T,1) in the reference loader
            if y.shape[1] != self.T:
                raise ValueError(f"y_true must be (n_peptides, {self.T})")
        ptr, col, val = _csr(S)
        tptr, trow, tval = _csr(S.T)
        m.n_pep = P
        m.nnz = int(ptr[-1])
        m.row_ptr, m.col, m.val = self._dev(ptr).data_ptr(), self._dev(col).data_ptr(), self._dev(val).data_ptr()
        m.t_ptr, m.t_row, m.t_val = self._dev(tptr).data_ptr(), self._dev(trow).data_ptr(), self._dev(tval).data_ptr()
        m.y = self._dev(y).data_ptr()
        if kind == L.LOSS_UPTAKE_SIGMA_MSE:
            if W is None:
                raise ValueError("hdx_uptake_sigma_MSE_loss needs Dataset.covariance_matrix")
            W = np.asarray(W, np.float32)
            if W.shape != (P, P):
                raise ValueError("covariance_matrix must be (n_peptides, n_peptides)")
            m.W = self._dev(W).data_ptr()

    # ------------------------------------------------------------------ state
    def init_state(self, z0, bc: float, bh: float, fw, fs) -> None:
        """OptaxOptimizer.initialise (reference opt/optimiser.py:221-304) + the step-0 forward."""
        with torch.cuda.device(self.device):
            z0 = self._dev_f32(z0).reshape(-1)
            if z0.numel() != self.F:
                raise ValueError("frame_weights must have one entry per frame")
            zmax = torch.empty(1, dtype=torch.float32, device=self.device)
            L.check(self.lib.jaxent_bv_max(z0.data_ptr(), self.F, zmax.data_ptr(), _stream_ptr()))
            self.comm.max_(zmax)
            n = len(self.kinds)
            fw = np.asarray(fw, np.float32).reshape(-1)
            fs = np.asarray(fs, np.float32).reshape(-1)
            if fw.shape[0] != n or fs.shape[0] != n:
                raise ValueError("forward_model_weights / forward_model_scaling must match the number of losses")
            fwc = (C.c_float * L.MAX_SLOTS)(*fw.tolist())
            fsc = (C.c_float * L.MAX_SLOTS)(*fs.tolist())
            L.check(
                self.lib.jaxent_bv_state_init(
                    self.plan, z0.data_ptr(), float(zmax.item()), float(bc), float(bh), fwc, fsc, _stream_ptr()
                )
            )
            if self.fused:
                # every rank's state_init (which clears its local flags) must complete before any peer stores to it
                self.comm.sum_(torch.zeros(1, device=self.device))
            self._phase(0)
            self.steps_done = 0
            self._graph = None

    def _phase(self, mode: int) -> None:
        s = _stream_ptr()
        L.check(self.lib.jaxent_bv_pass(self.plan, mode, s))
        if self.sharded and not self.fused:
            # NCCL exchange: complete the sums into JAXENT_BUF_RED, all-reduce them, the tail reads them from there
            L.check(self.lib.jaxent_bv_reduce(self.plan, s))
            self.comm.sum_(self.red)
            self.launches += 1
        L.check(self.lib.jaxent_bv_tail(self.plan, s))
        if self.sharded and not self.fused:
            self.comm.sum_(self.klsum)
        self.launches += 2

    def step(self, n: int = 1) -> None:
        """n optimise steps, enqueued on the current stream without host synchronisation."""
        with torch.cuda.device(self.device):
            for _ in range(n):
                self._phase(1)
            self.steps_done += n

    def capture(self, steps_per_graph: int = 2) -> None:
        """Capture `steps_per_graph` steps (kernels + collectives) into one CUDA graph."""
        with torch.cuda.device(self.device):
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                g = torch.cuda.CUDAGraph()
                launches = self.launches
                with torch.cuda.graph(g, stream=side):
                    for _ in range(steps_per_graph):
                        self._phase(1)
                self.launches = launches
            torch.cuda.current_stream().wait_stream(side)
            self._graph, self._graph_steps = g, steps_per_graph

    def step_graph(self, n_graphs: int = 1) -> None:
        if self._graph is None:
            raise RuntimeError("capture() has not been called")
        for _ in range(n_graphs):
            self._graph.replay()
        self.steps_done += n_graphs * self._graph_steps
        self.launches += (3 if self.sharded and not self.fused else 2) * n_graphs * self._graph_steps

    def exchange_error(self) -> int:
        """non-zero once a peer-flag wait of the fused exchange timed out (a peer never arrived)"""
        if not self.fused:
            return 0
        out = torch.zeros(1, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            L.check(self.lib.jaxent_bv_exchange_error(self.plan, out.data_ptr(), _stream_ptr()))
        return int(out.item())

    def finish_record(self) -> None:
        L.check(self.lib.jaxent_bv_finish_record(self.plan, _stream_ptr()))

    # ------------------------------------------------------------------ results
    def records(self, first: int, count: int):
        """Loss records of steps [first, first+count) as ctypes structs (device -> host copy)."""
        if count <= 0:
            return []
        sz = C.sizeof(L.LossRecord)
        host = self._records.cpu().numpy()
        out = []
        for s in range(first, first + count):
            i = s % L.RECORD_RING
            out.append(L.LossRecord.from_buffer_copy(host[i * sz : (i + 1) * sz].tobytes()))
        return out

    def export_state(self, want=("z", "m", "v", "g")):
        with torch.cuda.device(self.device):
            bufs = {k: torch.empty(self.F, dtype=torch.float32, device=self.device) for k in want}
            sc = torch.empty(8, dtype=torch.float32, device=self.device)
            ptr = lambda k: bufs[k].data_ptr() if k in bufs else None
            L.check(
                self.lib.jaxent_bv_export_state(self.plan, ptr("z"), ptr("m"), ptr("v"), ptr("g"), sc.data_ptr(), _stream_ptr())
            )
            s = sc.cpu().numpy()
        scal = dict(step=int(s[0]), lr=float(s[1]), model_lr=float(s[2]), mask_idx=int(s[3]), bc=float(s[4]), bh=float(s[5]),
                    count_frame=int(s[6]), count_model=int(s[7]))
        return bufs, scal

    def algorithmic_bytes_per_step(self) -> int:
        """8*R*F (features once) + 28*F (z,m,v,prior read; z,m,v written)."""
        return 8 * self.R * self.F + 28 * self.F

    def actual_bytes_per_step(self) -> int:
        """bytes the pass really streams: the packed features (fp32 or u8) once + the 28 B/frame of optimiser state."""
        return self.features.nbytes + 28 * self.F

    # ------------------------------------------------------------------ on-device loop
    def enable_tracking(self, convergence, learning_rate: float, ema_alpha: float = 0.5, min_steps_per_threshold: int = 2,
                        tolerance: float = 1e-10) -> None:
        """Arm the on-device ConvergenceTracker (reference opt/track.py:128-197, opt/run.py:272-348).  Must be
        called before init_state().  Thresholds are sorted descending and scaled by the learning rate exactly
        like create_convergence_thresholds (opt/track.py:12-21)."""
        thr = np.sort(np.atleast_1d(np.asarray(convergence, np.float32)))[::-1] * np.float32(learning_rate)
        n = int(thr.shape[0])
        if n < 1 or n > L.TRACK_MAX_THRESHOLDS:
            raise ValueError(f"between 1 and {L.TRACK_MAX_THRESHOLDS} convergence thresholds are supported on the device")
        with torch.cuda.device(self.device):
            self._trk_ema = torch.zeros(self.F, dtype=torch.float32, device=self.device)
            self._trk_snap_w = torch.zeros((n + 1, self.F), dtype=torch.float32, device=self.device)
            self._trk_snap_ema = torch.zeros((n + 1, self.F), dtype=torch.float32, device=self.device)
            self._trk_snaps = torch.zeros((n + 1) * C.sizeof(L.TrackSnapshot), dtype=torch.uint8, device=self.device)
            self._trk_status = torch.zeros(C.sizeof(L.TrackStatus), dtype=torch.uint8, device=self.device)
        arr = (C.c_float * n)(*[float(t) for t in thr])
        L.check(self.lib.jaxent_bv_track_init(self.plan, arr, n, float(ema_alpha), int(min_steps_per_threshold), float(tolerance),
                                              self._trk_ema.data_ptr(), self._trk_snap_w.data_ptr(), self._trk_snap_ema.data_ptr(),
                                              self._trk_snaps.data_ptr()))
        self._trk_n = n
        self.tracking = True

    def track_status(self) -> "L.TrackStatus":
        """Device -> host read of the loop status (the only host synchronisation of the on-device loop)."""
        with torch.cuda.device(self.device):
            L.check(self.lib.jaxent_bv_track_status(self.plan, self._trk_status.data_ptr(), _stream_ptr()))
            return L.TrackStatus.from_buffer_copy(self._trk_status.cpu().numpy().tobytes())

    def track_snapshots(self, n: int):
        """[(TrackSnapshot, weights[F] view, ema_weights[F] view or None)] for the first n snapshots."""
        sz = C.sizeof(L.TrackSnapshot)
        host = self._trk_snaps.cpu().numpy()
        out = []
        for i in range(n):
            sn = L.TrackSnapshot.from_buffer_copy(host[i * sz : (i + 1) * sz].tobytes())
            out.append((sn, self._trk_snap_w[i], self._trk_snap_ema[i] if sn.have_ema else None))
        return out

    def run(self, n_steps: int, chunk: int = 32):
        """Up to n_steps optimise steps with the stop decision taken on the device; the host only looks at
        the status every `chunk` steps (kernels launched after the stop are no-ops).  Returns TrackStatus."""
        if not getattr(self, "tracking", False):
            raise RuntimeError("enable_tracking() has not been called")
        done = 0
        st = None
        first = self.steps_done
        while done < n_steps:
            m = min(chunk, n_steps - done)
            self.step(m)
            done += m
            st = self.track_status()
            # launches after the device-side stop are no-ops: the number of updates applied is the device's count
            self.steps_done = int(st.steps_done)
            if st.stop:
                break
        self.finish_record()
        if st is None:
            st = self.track_status()
            self.steps_done = int(st.steps_done)
        assert self.steps_done >= first
        return st
