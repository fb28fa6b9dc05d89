"""The NumPy restatement of calc_BV_contacts_universe (oracle/contacts_oracle.py) on hand-checkable cases.

The reference has no numeric test for its featuriser kernel; these cases pin the semantics its code states
(jaxent/src/models/func/contacts.py:125-243): inclusive residue-ignore window, `<= radius` cut-off, the rational
switch with its r = r0 limit, orthorhombic minimum image."""
import numpy as np

from oracle import contacts_oracle as CO


def _line(n, spacing=1.0):
    xyz = np.zeros((1, n, 3), np.float32)
    xyz[0, :, 0] = np.arange(n) * spacing
    return xyz


def test_hard_cutoff_counts_and_ignore_window():
    xyz = _line(11)  # atoms at x = 0..10, residue id = atom index
    idx = np.arange(11)
    out = CO.calc_bv_contacts(xyz, [5], [5], idx, idx, radius=3.0, residue_ignore=(-2, 2))
    # within 3.0 of atom 5: atoms 2..8; residues 3..7 ignored (inclusive window) -> atoms 2 and 8
    assert out.shape == (1, 1) and out[0, 0] == 2
    out0 = CO.calc_bv_contacts(xyz, [5], [5], idx, idx, radius=3.0, residue_ignore=(0, 0))
    assert out0[0, 0] == 6  # only the atom's own residue is skipped
    out_edge = CO.calc_bv_contacts(xyz, [0], [0], idx, idx, radius=3.0, residue_ignore=(-2, 2))
    assert out_edge[0, 0] == 1  # atom 3 (distance exactly 3.0 counts: `<= radius`)


def test_rational_switch_values():
    r = np.array([0.0, 1.0, 2.0])
    v = CO.rational_switch(r, 2.0)
    np.testing.assert_allclose(v, [1.0, 1.0 / (1.0 + (0.5) ** 6), 0.5], rtol=1e-12)  # limit 1/2 at r = r0
    xyz = _line(6)
    idx = np.arange(6)
    out = CO.calc_bv_contacts(xyz, [0], [0], idx, idx, radius=4.0, residue_ignore=(0, 0), switch=True)
    expect = sum(1.0 / (1.0 + (d / 4.0) ** 6) for d in (1.0, 2.0, 3.0, 4.0))
    np.testing.assert_allclose(out[0, 0], expect, rtol=1e-12)


def test_minimum_image():
    xyz = np.zeros((1, 2, 3), np.float32)
    xyz[0, 1, 0] = 9.5  # 0.5 away through the periodic boundary of a 10 A box
    a = CO.calc_bv_contacts(xyz, [0], [0], [1], [10], radius=1.0)
    b = CO.calc_bv_contacts(xyz, [0], [0], [1], [10], radius=1.0, box=[10.0, 10.0, 10.0])
    assert a[0, 0] == 0 and b[0, 0] == 1


def test_triclinic_cell_matrix_and_minimum_image():
    # MDAnalysis' convention: a along x, b in the xy plane; a cubic cell gives the diagonal matrix
    np.testing.assert_allclose(CO.triclinic_vectors([10, 10, 10, 90, 90, 90]), np.diag([10.0, 10.0, 10.0]), atol=1e-12)
    B = CO.triclinic_vectors([10, 12, 9, 70, 80, 65])
    np.testing.assert_allclose(np.linalg.norm(B, axis=1), [10, 12, 9], rtol=1e-12)
    cosang = lambda u, v: u @ v / np.linalg.norm(u) / np.linalg.norm(v)
    np.testing.assert_allclose(np.rad2deg(np.arccos([cosang(B[1], B[2]), cosang(B[0], B[2]), cosang(B[0], B[1])])), [70, 80, 65], rtol=1e-12)
    # a separation of (b + c) + (0.3, 0.2, 0.1) is 0.374 long through the periodic images
    d = (B[1] + B[2] + np.array([0.3, 0.2, 0.1]))[None, :]
    np.testing.assert_allclose(CO.minimum_image_triclinic(d, B), [[0.3, 0.2, 0.1]], atol=1e-12)
    xyz = np.zeros((1, 2, 3), np.float32)
    xyz[0, 1] = d[0]
    a = CO.calc_bv_contacts(xyz, [0], [0], [1], [10], radius=1.0)
    b = CO.calc_bv_contacts(xyz, [0], [0], [1], [10], radius=1.0, box=[10, 12, 9, 70, 80, 65])
    assert a[0, 0] == 0 and b[0, 0] == 1
    # brute force over a wide image range agrees with the 27-image search
    rng = np.random.default_rng(0)
    dd = rng.uniform(-30, 30, size=(200, 3))
    best = None
    for i in range(-5, 6):
        for j in range(-5, 6):
            for k in range(-5, 6):
                r2 = ((dd + i * B[0] + j * B[1] + k * B[2]) ** 2).sum(-1)
                best = r2 if best is None else np.minimum(best, r2)
    np.testing.assert_allclose((CO.minimum_image_triclinic(dd, B) ** 2).sum(-1), best, rtol=1e-12)
