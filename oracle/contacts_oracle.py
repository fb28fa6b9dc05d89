"""NumPy restatement of calc_BV_contacts_universe (jaxent/src/models/func/contacts.py:125-243).  TEST INFRASTRUCTURE ONLY.

The MDAnalysis pieces (atom selections, `distance_array`) are replaced by index lists and an explicit double-precision
distance matrix with the minimum-image convention (orthorhombic lengths, or a general cell given as MDAnalysis `dimensions`); the masking, cut-off and switching logic follows the
reference line by line.  Parity pinning: the reference has no numeric test for this function (and MDAnalysis is not
installable here), so parity is unpinned beyond the hand-checkable cases in tests/test_contacts_cpu.py."""
import numpy as np


def rational_switch(r, r0):
    """contacts.py:191-199: (1 - x) / (1 - x^2) with x = (r / r0)^6, NaN at r = r0 replaced by 0.5."""
    with np.errstate(divide="ignore", invalid="ignore"):
        x = (r / r0) ** 6
        val = (1 - x) / (1 - x * x)
    val[np.isnan(val)] = 0.5
    return val


def triclinic_vectors(dimensions):
    """[lx, ly, lz, alpha, beta, gamma] (degrees) -> 3 x 3 lower-triangular box matrix with rows a, b, c -- MDAnalysis'
    lib.mdamath.triclinic_vectors, the convention `universe.dimensions` follows (contacts.py:210-215 passes it as `box=`)."""
    lx, ly, lz, al, be, ga = [float(x) for x in dimensions]
    al, be, ga = np.deg2rad([al, be, ga])
    B = np.zeros((3, 3))
    B[0, 0] = lx
    B[1, 0] = ly * np.cos(ga)
    B[1, 1] = ly * np.sin(ga)
    B[2, 0] = lz * np.cos(be)
    B[2, 1] = lz * (np.cos(al) - np.cos(be) * np.cos(ga)) / np.sin(ga)
    B[2, 2] = np.sqrt(lz * lz - B[2, 0] ** 2 - B[2, 1] ** 2)
    return B


def minimum_image_triclinic(d, B):
    """true minimum image of the separation vectors d (..., 3) in the cell B: fractional rounding, then the 27 neighbouring
    images are searched (what MDAnalysis' calc_distances.h does after wrapping both points into the cell)."""
    s = d @ np.linalg.inv(B)
    d0 = (s - np.rint(s)) @ B
    best = d0.copy()
    best_r2 = (d0 * d0).sum(-1)
    for i in (-1, 0, 1):
        for j in (-1, 0, 1):
            for k in (-1, 0, 1):
                cand = d0 + i * B[0] + j * B[1] + k * B[2]
                r2 = (cand * cand).sum(-1)
                better = r2 < best_r2
                best[better] = cand[better]
                best_r2[better] = r2[better]
    return best


def calc_bv_contacts(coords, target_idx, target_resid, contact_idx, contact_resid, radius, residue_ignore=(-2, 2), switch=False, box=None):
    coords = np.asarray(coords, np.float32).astype(np.float64)
    F = coords.shape[0]
    target_idx, contact_idx = np.asarray(target_idx), np.asarray(contact_idx)
    resid_diff = np.asarray(contact_resid)[None, :] - np.asarray(target_resid)[:, None]  # :178-185
    ignore_mask = (resid_diff >= residue_ignore[0]) & (resid_diff <= residue_ignore[1])  # :189, inclusive
    out = np.zeros((len(target_idx), F))
    for f in range(F):
        d = coords[f, contact_idx][None, :, :] - coords[f, target_idx][:, None, :]
        if box is not None and np.asarray(box).shape[-1] == 6:  # [lx, ly, lz, alpha, beta, gamma]: general (triclinic) cell
            b = np.asarray(box, np.float64).reshape(-1, 6)
            d = minimum_image_triclinic(d, triclinic_vectors(b[f if b.shape[0] > 1 else 0]))
        elif box is not None:
            b = np.asarray(box, np.float32).astype(np.float64).reshape(-1, 3)
            b = b[f if b.shape[0] > 1 else 0]
            d -= b * np.rint(d / b)
        dists = np.sqrt((d * d).sum(-1))  # distance_array :210-215
        dists[ignore_mask] = np.inf  # :218
        if switch:
            within = dists <= radius  # :222
            vals = rational_switch(dists, radius)  # :226
            vals[~within] = 0.0  # :229
            out[:, f] = vals.sum(1)  # :232
        else:
            out[:, f] = (dists <= radius).sum(1)  # :236-240
    return out
