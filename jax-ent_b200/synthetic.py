"""Seeded synthetic BV ensembles of the BASELINE.json shapes.

Product-side helper (NumPy only, no oracle import): bench.py, the tests and smoke() all build
their inputs here so that the CUDA path and the CPU oracle see identical arrays.

Contact counts are integer valued and stored as fp32, like the reference's hard-cutoff BV
features (jaxent/src/models/func/contacts.py:236-240; BPTI text fixtures hold 0..21).
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from typing import Optional

import numpy as np


@dataclass
class SyntheticEnsemble:
    heavy: np.ndarray  # (R,F) fp32, integer valued
    acceptor: np.ndarray  # (R,F) fp32, integer valued
    k_ints: np.ndarray  # (R,) fp32
    timepoints: np.ndarray  # (T,) fp32
    S_train: np.ndarray  # (P_tr,R) fp32 dense peptide<-residue map
    y_train: np.ndarray  # (P_tr,T) fp32 (uptake) or (P_tr,) (PF)
    S_val: np.ndarray
    y_val: np.ndarray
    prior: np.ndarray  # (F,) fp32 uniform
    bc: float = 0.35
    bh: float = 2.0


def _uniform_counts(rng_seed, hi, n_frames, out, r0, r1):
    """rows r0:r1 of ``out`` <- integers U{0..hi_r}; one child generator per row block."""
    rng = np.random.default_rng(rng_seed)
    step = 1 << 20
    for f0 in range(0, n_frames, step):
        f1 = min(n_frames, f0 + step)
        blk = rng.integers(0, hi[r0:r1, None] + 1, size=(r1 - r0, f1 - f0), dtype=np.uint8)
        out[r0:r1, f0:f1] = blk


def make_counts(n_res: int, n_frames: int, mean: np.ndarray, seed: int, threads: int = 8) -> np.ndarray:
    """(R,F) fp32 integer counts with per-residue mean ``mean`` (uniform on 0..2*mean)."""
    hi = np.clip(np.rint(2 * mean), 0, 255).astype(np.int64)
    out = np.empty((n_res, n_frames), np.float32)
    nblk = max(1, min(threads, n_res))
    edges = np.linspace(0, n_res, nblk + 1).astype(int)
    seeds = np.random.SeedSequence(seed).spawn(nblk)
    with ThreadPoolExecutor(nblk) as ex:
        list(ex.map(lambda i: _uniform_counts(seeds[i], hi, n_frames, out, edges[i], edges[i + 1]), range(nblk)))
    return out


def make_peptides(n_res: int, n_pep: int, seed: int):
    """Contiguous residue windows of length 5..15; S[p,r] = 1/len inside the window
    (overlap/peptide_len, jaxent/src/data/splitting/sparse_map.py:173,209-213)."""
    rng = np.random.default_rng(seed)
    S = np.zeros((n_pep, n_res), np.float32)
    for p in range(n_pep):
        ln = int(rng.integers(5, 16))
        ln = min(ln, n_res)
        s = int(rng.integers(0, n_res - ln + 1))
        S[p, s : s + ln] = 1.0 / ln
    return S


def make_ensemble(
    n_frames: int,
    n_res: int,
    n_timepoints: int = 8,
    n_pep: Optional[int] = None,
    uptake: bool = True,
    seed: int = 0,
    threads: int = 8,
) -> SyntheticEnsemble:
    """Synthetic ensemble: features seed, peptides seed+1, hidden weights seed+2, noise seed+3."""
    rng = np.random.default_rng(seed)
    lam = rng.uniform(5, 40, n_res)
    mu = rng.uniform(0.1, 1.5, n_res)
    heavy = make_counts(n_res, n_frames, lam, seed * 7919 + 11, threads)
    acceptor = make_counts(n_res, n_frames, mu, seed * 7919 + 13, threads)
    k_ints = (10.0 ** rng.uniform(-2, 2, n_res)).astype(np.float32)
    if n_timepoints == 3:
        tp = np.array([0.167, 1.0, 10.0], np.float32)
    elif n_timepoints == 5:
        tp = np.array([0.167, 1.0, 10.0, 60.0, 120.0], np.float32)
    else:
        tp = np.geomspace(0.167, 1440.0, n_timepoints).astype(np.float32)
    n_pep = n_pep if n_pep is not None else max(4, n_res // 2)
    S = make_peptides(n_res, n_pep, seed + 1)

    # hidden target ensemble: Dirichlet(0.3) weights on a bounded subsample of frames (cheap at F=1e6)
    rng_w = np.random.default_rng(seed + 2)
    n_sub = min(n_frames, 4096)
    idx = rng_w.choice(n_frames, n_sub, replace=False) if n_sub < n_frames else np.arange(n_frames)
    wsub = rng_w.dirichlet(np.full(n_sub, 0.3))
    a = heavy[:, idx].astype(np.float64) @ wsub
    b = acceptor[:, idx].astype(np.float64) @ wsub
    lp = 0.35 * a + 2.0 * b
    rng_n = np.random.default_rng(seed + 3)
    if uptake:
        u = 1.0 - np.exp(-k_ints[None, :].astype(np.float64) * tp[:, None].astype(np.float64) * np.exp(-lp)[None, :])
        y = (S.astype(np.float64) @ u.T) + rng_n.normal(0, 0.02, (n_pep, len(tp)))
    else:
        y = S.astype(np.float64) @ lp + rng_n.normal(0, 0.05, n_pep)
    y = y.astype(np.float32)

    perm = np.random.default_rng(seed + 1).permutation(n_pep)
    n_val = max(1, n_pep // 5)
    va, tr = np.sort(perm[:n_val]), np.sort(perm[n_val:])
    prior = np.full(n_frames, 1.0 / n_frames, np.float32)
    return SyntheticEnsemble(heavy, acceptor, k_ints, tp if uptake else tp, S[tr], y[tr], S[va], y[va], prior)
