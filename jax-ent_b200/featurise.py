"""BV featuriser on the GPU: contact counts from raw coordinates.

The reference computes the two BV feature arrays with MDAnalysis on the CPU (`calc_BV_contacts_universe`,
jaxent/src/models/func/contacts.py:125-243, driven by `BV_model.featurise`, models/HDX/BV/forwardmodel.py:214-300).
Trajectory / topology parsing stays with the caller (MDAnalysis is not part of this path): the boundary is the coordinate
array of the trajectory plus the atom index / residue id lists the reference's selections produce.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib as L
from .engine import _stream_ptr
from .models.HDX.BV.features import BV_input_features

HEAVY_RADIUS = 6.5  # BV_model_Config.heavy_radius (models/config.py:20), Angstrom
O_RADIUS = 2.4  # BV_model_Config.o_radius (models/config.py:21)


def _i32(x, dev):
    return torch.as_tensor(np.asarray(x), dtype=torch.int32).to(dev).contiguous()


def _cell_rows(dimensions: np.ndarray) -> np.ndarray:
    """(n, 6) [lx, ly, lz, alpha, beta, gamma] in degrees -> (n, 6) [a_x, b_x, b_y, c_x, c_y, c_z], the non-zero elements of the
    lower-triangular cell matrix (rows a, b, c) in MDAnalysis' convention for `universe.dimensions`."""
    lx, ly, lz = dimensions[:, 0], dimensions[:, 1], dimensions[:, 2]
    ca, cb, cg = (np.cos(np.deg2rad(dimensions[:, k])) for k in (3, 4, 5))
    sg = np.sin(np.deg2rad(dimensions[:, 5]))
    cx, cy = lz * cb, lz * (ca - cb * cg) / sg
    cz2 = lz * lz - cx * cx - cy * cy
    if np.any(cz2 <= 0) or np.any(sg <= 0):
        raise ValueError("box angles do not describe a cell")
    return np.stack([lx, ly * cg, ly * sg, cx, cy, np.sqrt(cz2)], axis=1)


def calc_BV_contacts(coords, target_idx: Sequence[int], target_resid: Sequence[int], contact_idx: Sequence[int],
                     contact_resid: Sequence[int], radius: float, residue_ignore=(-2, 2), switch: bool = False, box=None,
                     device: Optional[str] = None) -> torch.Tensor:
    """coords: (n_frames, n_atoms, 3) float32 positions; returns (n_targets, n_frames) float32 contacts on the device.
    Same arguments as the reference function, with the AtomGroups replaced by (atom index, residue id) lists.  box: None,
    orthorhombic lengths (3,) / (n_frames, 3), or MDAnalysis `dimensions` (6,) / (n_frames, 6) for a general (triclinic) cell."""
    if not torch.cuda.is_available():
        raise RuntimeError("jaxent_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    lib = L.load()
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    xyz = (coords if isinstance(coords, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(coords, np.float32))).to(dev, torch.float32).contiguous()
    if xyz.ndim != 3 or xyz.shape[2] != 3:
        raise ValueError("coords must be (n_frames, n_atoms, 3)")
    F, A = int(xyz.shape[0]), int(xyz.shape[1])
    ti, tr, ci, cr = _i32(target_idx, dev), _i32(target_resid, dev), _i32(contact_idx, dev), _i32(contact_resid, dev)
    if ti.numel() != tr.numel() or ci.numel() != cr.numel():
        raise ValueError("one residue id per target / contact atom")
    if ti.numel() == 0 or ci.numel() == 0:
        raise ValueError("empty atom selection")
    if int(ti.max()) >= A or int(ci.max()) >= A or int(ti.min()) < 0 or int(ci.min()) < 0:
        raise ValueError("atom index outside the coordinate array")
    bx, stride = None, 3
    if box is not None:
        b = np.asarray(box, np.float64)
        if b.shape[-1] == 6:  # MDAnalysis `dimensions`: [lx, ly, lz, alpha, beta, gamma] (what the reference passes, contacts.py:210-215)
            b = b.reshape(-1, 6)
            if np.allclose(b[:, 3:], 90.0):
                b = b[:, :3]
            else:
                b = _cell_rows(b)
                stride = 6
                if float(radius) >= 0.5 * float(b[:, [0, 2, 5]].min()):
                    raise ValueError("radius must be below half the smallest diagonal element of the triclinic cell")
        else:
            b = b.reshape(-1, 3)
        bx = torch.as_tensor(b.astype(np.float32)).to(dev)
        if bx.shape[0] == 1:
            bx = bx.expand(F, stride)
        bx = bx.contiguous()
        if bx.shape[0] != F:
            raise ValueError("box must be (3,) / (n_frames, 3) orthorhombic lengths or (6,) / (n_frames, 6) MDAnalysis dimensions")
    out = torch.empty((ti.numel(), F), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.jaxent_bv_contacts_cell(xyz.data_ptr(), F, A, ti.data_ptr(), tr.data_ptr(), ti.numel(), ci.data_ptr(), cr.data_ptr(),
                                            ci.numel(), int(residue_ignore[0]), int(residue_ignore[1]), float(radius), int(bool(switch)),
                                            bx.data_ptr() if bx is not None else None, stride, out.data_ptr(), _stream_ptr()))
    return out


def featurise_BV(coords, amide_n, amide_h, heavy_atoms, oxygen_atoms, k_ints=None, heavy_radius: float = HEAVY_RADIUS,
                 o_radius: float = O_RADIUS, switch: bool = False, box=None, device: Optional[str] = None) -> BV_input_features:
    """BV_model.featurise (models/HDX/BV/forwardmodel.py:214-300) for one trajectory.  Every selection is a pair
    (atom indices, residue ids): amide N (targets of the heavy-atom contacts), amide H (targets of the acceptor contacts),
    all heavy atoms (`not type H`) and all oxygens (`type O`).  Returns BV_input_features with (n_residues, n_frames) arrays."""
    heavy = calc_BV_contacts(coords, amide_n[0], amide_n[1], heavy_atoms[0], heavy_atoms[1], heavy_radius, switch=switch, box=box, device=device)
    acc = calc_BV_contacts(coords, amide_h[0], amide_h[1], oxygen_atoms[0], oxygen_atoms[1], o_radius, switch=switch, box=box, device=device)
    if heavy.shape != acc.shape:
        raise ValueError("Input features have different shapes")  # one amide N and one amide H per residue
    return BV_input_features(heavy, acc, k_ints)
