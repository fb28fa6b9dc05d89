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
