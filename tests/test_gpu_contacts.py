"""GPU featuriser (jaxent_bv_contacts) against the NumPy restatement of calc_BV_contacts_universe."""
import numpy as np
import pytest
import torch

from oracle import contacts_oracle as CO

pytestmark = pytest.mark.gpu


def _protein_like(F, n_res, seed):
    """a random-walk chain with N, H, CA, C, O (+ a few side-chain heavy atoms / hydrogens) per residue"""
    rng = np.random.default_rng(seed)
    per = 9
    names = ["N", "H", "CA", "HA", "C", "O", "CB", "HB", "OG"]
    ca = np.cumsum(rng.normal(scale=2.2, size=(F, n_res, 3)), axis=1)
    xyz = (ca[:, :, None, :] + rng.normal(scale=1.2, size=(F, n_res, per, 3))).reshape(F, n_res * per, 3).astype(np.float32)
    resid = np.repeat(np.arange(1, n_res + 1), per)
    name = np.tile(names, n_res)
    idx = np.arange(n_res * per)
    sel = lambda m: (idx[m], resid[m])
    return xyz, sel(name == "N"), sel(name == "H"), sel(~np.char.startswith(name, "H")), sel(np.char.startswith(name, "O"))


@pytest.mark.parametrize("F,n_res,switch,box", [(7, 60, False, None), (5, 150, False, 38.0), (4, 300, True, None), (3, 40, True, 25.0)])
def test_contacts_match_oracle(F, n_res, switch, box):
    from jaxent_b200.featurise import calc_BV_contacts

    xyz, n_sel, h_sel, heavy, oxy = _protein_like(F, n_res, 5)
    bx = None if box is None else [box, box * 1.1, box * 0.9]
    for (tgt, con, radius) in ((n_sel, heavy, 6.5), (h_sel, oxy, 2.4)):
        ref = CO.calc_bv_contacts(xyz, tgt[0], tgt[1], con[0], con[1], radius, switch=switch, box=bx)
        got = calc_BV_contacts(xyz, tgt[0], tgt[1], con[0], con[1], radius, switch=switch, box=bx, device="cuda:0").cpu().numpy()
        assert got.shape == ref.shape == (n_res, F)
        if switch:
            np.testing.assert_allclose(got, ref, rtol=2e-5, atol=1e-5)
        else:
            np.testing.assert_array_equal(got, ref)  # counts are exact (borderline pairs are re-evaluated in fp64)
        assert ref.max() > 0


@pytest.mark.parametrize("dims,switch", [((40.0, 44.0, 36.0, 70.0, 80.0, 65.0), False), ((30.0, 30.0, 30.0, 60.0, 60.0, 90.0), False),
                                         ((38.0, 41.0, 35.0, 100.0, 95.0, 110.0), True), ((33.0, 35.0, 31.0, 90.0, 90.0, 90.0), False)])
def test_contacts_triclinic_cell(dims, switch):
    """general cells (`box=universe.dimensions`, contacts.py:210-215): truncated-octahedron-like, rhombic-dodecahedron-like
    and obtuse angles; the chain spans several cell lengths so most pairs are seen through an image."""
    from jaxent_b200.featurise import calc_BV_contacts

    F, n_res = 4, 120
    xyz, n_sel, h_sel, heavy, oxy = _protein_like(F, n_res, 11)
    per_frame = np.tile(np.asarray(dims, np.float64), (F, 1))
    per_frame[:, :3] *= (1.0 + 0.01 * np.arange(F))[:, None]  # NPT-like: the cell changes from frame to frame
    for (tgt, con, radius) in ((n_sel, heavy, 6.5), (h_sel, oxy, 2.4)):
        ref = CO.calc_bv_contacts(xyz, tgt[0], tgt[1], con[0], con[1], radius, switch=switch, box=per_frame)
        free = CO.calc_bv_contacts(xyz, tgt[0], tgt[1], con[0], con[1], radius, switch=switch)
        got = calc_BV_contacts(xyz, tgt[0], tgt[1], con[0], con[1], radius, switch=switch, box=per_frame, device="cuda:0").cpu().numpy()
        if switch:
            np.testing.assert_allclose(got, ref, rtol=2e-5, atol=1e-5)
        else:
            # the cell is fp32 on the device, fp64 in the oracle: a pair within 1e-5 A of the cut-off may fall either way
            assert np.abs(got - ref).max() <= 1 and (got != ref).mean() < 2e-3
        if radius > 3:
            assert (ref != free).any()  # the images matter in this case
    if dims[3:] != (90.0, 90.0, 90.0):  # orthorhombic `dimensions` take the per-axis path, which has no radius limit
        with pytest.raises(ValueError, match="half"):
            calc_BV_contacts(xyz, n_sel[0], n_sel[1], heavy[0], heavy[1], 20.0, box=dims, device="cuda:0")


def test_featurise_feeds_the_optimise_path():
    from jaxent_b200.featurise import featurise_BV
    from jaxent_b200.batch_engine import pack_gemm_features

    xyz, n_sel, h_sel, heavy, oxy = _protein_like(64, 48, 9)
    feats = featurise_BV(xyz, n_sel, h_sel, heavy, oxy, device="cuda:0")
    assert feats.features_shape == (48, 64)
    h = feats.heavy_contacts
    assert isinstance(h, torch.Tensor) and h.is_cuda and float(h.max()) <= 255 and torch.equal(h, torch.round(h))
    assert pack_gemm_features(feats.heavy_contacts, feats.acceptor_contacts, "cuda:0").exact  # integer counts: exact in bf16 / u8


def test_contacts_argument_errors():
    from jaxent_b200.featurise import calc_BV_contacts

    xyz = np.zeros((2, 10, 3), np.float32)
    with pytest.raises(ValueError, match="outside"):
        calc_BV_contacts(xyz, [11], [1], [0], [5], 3.0, device="cuda:0")
    with pytest.raises(ValueError, match="n_frames, n_atoms, 3"):
        calc_BV_contacts(xyz[0], [1], [1], [0], [5], 3.0, device="cuda:0")
    with pytest.raises(ValueError, match="empty"):
        calc_BV_contacts(xyz, [], [], [0], [5], 3.0, device="cuda:0")
