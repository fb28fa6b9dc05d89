"""Data formats either side of the hot path.

* features `.npz` in the key layout of the reference's `AbstractFeatures.save / load / save_features / load_features`
  (custom_types/features.py:91-286), so files written by either side load on the other;
* HDXer text files: per-residue contact files `Contacts_chain_<c>_res_<i>.tmp` / `Hbonds_chain_<c>_res_<i>.tmp`
  (one line per frame; what the reference's BPTI / MoPrP fixtures are stored as), intrinsic rates (`load_HDXer_kints`,
  utils/HDXer/load_data.py:23-47), segment files and dfrac / segment-fraction tables (`load_HDXer_dfrac`, `:50-122`;
  `save_dfrac_file` `:125-157`).
"""
from __future__ import annotations

import glob
import os
import re
from typing import Optional, Sequence

import numpy as np

from .._arrays import to_numpy
from ..models.HDX.BV.features import BV_input_features

REF_CLASS_MODULE = "jaxent.src.models.HDX.BV.features"
_DYNAMIC_SLOTS = ("heavy_contacts", "acceptor_contacts", "k_ints")


# ------------------------------------------------------------------------------------------ features .npz
def save_features_npz(features: BV_input_features, filepath: str) -> str:
    """`Input_Features.save` (custom_types/features.py:91-126): class metadata + one array per dynamic slot."""
    if not filepath.endswith(".npz"):
        filepath += ".npz"
    d = os.path.dirname(filepath)
    if d:
        os.makedirs(d, exist_ok=True)
    save = {
        "__class_module__": REF_CLASS_MODULE,
        "__class_name__": "BV_input_features",
        "__static_data__": np.asarray((), dtype=object),
        "__dynamic_slots__": np.asarray(_DYNAMIC_SLOTS),
    }
    for slot in _DYNAMIC_SLOTS:
        v = getattr(features, slot)
        save[slot] = np.asarray(None, dtype=object) if v is None else to_numpy(v).astype(np.float32)
    np.savez(filepath, **save)
    return filepath


def _maybe_none(a):
    return None if (isinstance(a, np.ndarray) and a.shape == () and a.dtype == object and a.item() is None) else a


def load_features_npz(filepath: str) -> BV_input_features:
    """Reads both layouts the reference writes: `save` (with class metadata, `:128-196`) and `save_features`
    (`:198-217`, plain slot -> array).  Raises FileNotFoundError / KeyError / TypeError like the reference."""
    if not str(filepath).endswith(".npz") and not os.path.exists(filepath):
        filepath = str(filepath) + ".npz"
    if not os.path.exists(filepath):
        raise FileNotFoundError(f"Features file not found: {filepath}")
    data = np.load(filepath, allow_pickle=True)
    if "__class_name__" in data:
        name = str(data["__class_name__"].item())
        if name != "BV_input_features":
            raise TypeError(f"Loaded class {name} is not a subclass of BV_input_features")
        slots = tuple(str(s) for s in data["__dynamic_slots__"])
        missing = [s for s in slots if s not in data]
        if missing:
            raise KeyError(f"File {filepath} is missing required metadata: {missing}")
    else:
        slots = tuple(s for s in _DYNAMIC_SLOTS if s in data)
    kw = {s: _maybe_none(data[s]) for s in slots}
    for req in ("heavy_contacts", "acceptor_contacts"):
        if kw.get(req) is None:
            raise TypeError(f"BV_input_features.__init__() missing required features: {req}")
    kw = {k: (None if v is None else np.ascontiguousarray(v, np.float32)) for k, v in kw.items()}
    return BV_input_features(heavy_contacts=kw["heavy_contacts"], acceptor_contacts=kw["acceptor_contacts"], k_ints=kw.get("k_ints"))


# ------------------------------------------------------------------------------------------ HDXer text files
def load_HDXer_contacts(directory: str, chain: int = 0, resids: Optional[Sequence[int]] = None):
    """Per-residue HDXer contact files -> (BV_input_features without k_ints, resids).  One file per residue, one line per
    frame: `Contacts_chain_<c>_res_<i>.tmp` (heavy-atom contacts) and `Hbonds_chain_<c>_res_<i>.tmp` (acceptor contacts)."""
    if resids is None:
        files = glob.glob(os.path.join(directory, f"Contacts_chain_{chain}_res_*.tmp"))
        resids = sorted(int(re.search(r"res_(\d+)\.tmp$", f).group(1)) for f in files)
    if not resids:
        raise FileNotFoundError(f"no Contacts_chain_{chain}_res_*.tmp files in {directory}")
    heavy, acc = [], []
    for r in resids:
        hp = os.path.join(directory, f"Contacts_chain_{chain}_res_{r}.tmp")
        ap = os.path.join(directory, f"Hbonds_chain_{chain}_res_{r}.tmp")
        if not (os.path.exists(hp) and os.path.exists(ap)):
            raise FileNotFoundError(f"missing contact files of residue {r} in {directory}")
        heavy.append(np.loadtxt(hp, dtype=np.float32, ndmin=1))
        acc.append(np.loadtxt(ap, dtype=np.float32, ndmin=1))
    n = {len(x) for x in heavy} | {len(x) for x in acc}
    if len(n) != 1:
        raise ValueError("Input features have different shapes")  # reference models/core.py:61-62
    return BV_input_features(np.stack(heavy), np.stack(acc), None), list(resids)


def save_HDXer_contacts(features: BV_input_features, directory: str, resids: Sequence[int], chain: int = 0) -> None:
    os.makedirs(directory, exist_ok=True)
    h, a = to_numpy(features.heavy_contacts), to_numpy(features.acceptor_contacts)
    for i, r in enumerate(resids):
        np.savetxt(os.path.join(directory, f"Contacts_chain_{chain}_res_{r}.tmp"), h[i], fmt="%g")
        np.savetxt(os.path.join(directory, f"Hbonds_chain_{chain}_res_{r}.tmp"), a[i], fmt="%g")


def load_HDXer_kints(kint_path: str):
    """`resid rate` per line, `#` comments (utils/HDXer/load_data.py:23-47) -> (rates float32, resids)."""
    rates, resids = [], []
    with open(kint_path) as f:
        for line in f:
            line = line.strip()
            if not line or line.startswith("#"):
                continue
            parts = line.split()
            if len(parts) != 2:
                continue
            resids.append(int(parts[0]))
            rates.append(float(parts[1]))
    return np.asarray(rates, np.float32), resids


def load_segs_file(filename: str):
    """`start end` per line (save_segs_file, utils/HDXer/load_data.py:160-174)."""
    segs = np.loadtxt(filename, dtype=np.int64, ndmin=2)
    return [(int(a), int(b)) for a, b in segs[:, :2]]


def _parse_header(header_line: str):
    """time points from a `# ... Times / min: t1 t2 ...` style header; numbers after the last ':' (or all numbers)"""
    txt = header_line.strip().lstrip("#")
    tail = txt.split(":")[-1] if ":" in txt else txt
    nums = re.findall(r"[-+]?\d*\.?\d+(?:[eE][-+]?\d+)?", tail)
    return [float(x) for x in nums]


def load_HDXer_dfrac(filename: str, expected_timepoints: Optional[Sequence[float]] = None):
    """`res_start res_end frac_t1 frac_t2 ...` rows under a time-point header (utils/HDXer/load_data.py:50-122)
    -> (fractions (P,T) float32, segments [(start,end)], timepoints)."""
    if not os.path.exists(filename):
        raise FileNotFoundError(f"File not found: {filename}")
    with open(filename) as f:
        lines = f.readlines()
    if not lines:
        raise ValueError("File is empty")
    header_is_comment = lines[0].lstrip().startswith("#")
    timepoints = _parse_header(lines[0]) if header_is_comment else []
    rows, segs = [], []
    for line in lines[1 if header_is_comment else 0 :]:
        line = line.strip()
        if not line or line.startswith("#"):
            continue
        parts = re.split(r"\s+", line)
        if len(parts) < 3:
            continue
        try:
            fr = [float(v) for v in parts[2:] if v]
            a, b = int(float(parts[0])), int(float(parts[1]))
        except ValueError:
            continue
        if timepoints and len(fr) != len(timepoints):
            continue
        segs.append((a, b))
        rows.append(fr)
    if not timepoints and rows:
        timepoints = list(range(len(rows[0])))
    if expected_timepoints is not None and len(timepoints) != len(expected_timepoints):
        raise ValueError(f"Expected {len(expected_timepoints)} timepoints, found {len(timepoints)}")
    return np.asarray(rows, np.float32), segs, timepoints


def save_dfrac_file(filename: str, fractions, segments, timepoints) -> None:
    fr = np.asarray(fractions, np.float32)
    with open(filename, "w") as f:
        f.write("# ResStr, ResEnd, Deuterated fractions: Times / min: " + " ".join(f"{t:g}" for t in timepoints) + "\n")
        for (a, b), row in zip(segments, fr):
            f.write(f"{a} {b} " + " ".join(f"{v:.5f}" for v in row) + "\n")


def segments_to_map(segments, resids, exclude_first: int = 1) -> np.ndarray:
    """Peptide <- residue averaging map (P, R): entry 1/len over the residues of a segment that are present in `resids`,
    skipping the first `exclude_first` residues of each peptide (fast back-exchange; the reference's peptide trimming)."""
    idx = {r: i for i, r in enumerate(resids)}
    S = np.zeros((len(segments), len(resids)), np.float32)
    for p, (a, b) in enumerate(segments):
        cols = [idx[r] for r in range(a + exclude_first, b + 1) if r in idx]
        if cols:
            S[p, cols] = 1.0 / len(cols)
    return S
