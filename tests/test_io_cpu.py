"""I/O formats at the boundary: features .npz in the reference's key layout, HDXer text files."""
import os

import numpy as np
import pytest

from jaxent_b200.models.HDX.BV.features import BV_input_features
from jaxent_b200.utils import io as IO

GOLD = os.path.join(os.path.dirname(__file__), "golden", "bpti_bv_features.npz")


def _bpti():
    g = np.load(GOLD)
    return g["heavy"].astype(np.float32), g["acceptor"].astype(np.float32), g["resids"].tolist(), g["segs"]


def test_features_npz_round_trip_reference_keys(tmp_path):
    h, a, _, _ = _bpti()
    k = np.logspace(-2, 2, h.shape[0]).astype(np.float32)
    f = BV_input_features(h, a, k)
    path = str(tmp_path / "sub" / "features")
    f.save(path)
    raw = np.load(path + ".npz", allow_pickle=True)
    # the metadata the reference's AbstractFeatures.load needs (custom_types/features.py:148-181)
    assert str(raw["__class_module__"].item()) == "jaxent.src.models.HDX.BV.features"
    assert str(raw["__class_name__"].item()) == "BV_input_features"
    assert tuple(raw["__dynamic_slots__"]) == ("heavy_contacts", "acceptor_contacts", "k_ints")
    assert tuple(raw["__static_data__"]) == ()
    g = BV_input_features.load(path + ".npz")
    np.testing.assert_array_equal(g.heavy_contacts, h)
    np.testing.assert_array_equal(g.acceptor_contacts, a)
    np.testing.assert_array_equal(g.k_ints, k)
    assert g.features_shape == (53, 500)


def test_features_npz_none_kints_and_plain_layout(tmp_path):
    h, a, _, _ = _bpti()
    BV_input_features(h, a, None).save(str(tmp_path / "f.npz"))
    g = BV_input_features.load(str(tmp_path / "f.npz"))
    assert g.k_ints is None
    # `save_features` layout of the reference: plain slot -> array, no metadata
    np.savez(str(tmp_path / "plain.npz"), heavy_contacts=h, acceptor_contacts=a, k_ints=np.asarray(None, dtype=object))
    g2 = IO.load_features_npz(str(tmp_path / "plain.npz"))
    assert g2.k_ints is None and g2.heavy_contacts.shape == (53, 500)
    with pytest.raises(FileNotFoundError):
        IO.load_features_npz(str(tmp_path / "missing.npz"))
    np.savez(str(tmp_path / "bad.npz"), heavy_contacts=h)
    with pytest.raises(TypeError, match="acceptor_contacts"):
        IO.load_features_npz(str(tmp_path / "bad.npz"))


def test_hdxer_contact_files_round_trip(tmp_path):
    h, a, resids, _ = _bpti()
    d = str(tmp_path / "hdxer")
    IO.save_HDXer_contacts(BV_input_features(h, a, None), d, resids)
    assert os.path.exists(os.path.join(d, f"Contacts_chain_0_res_{resids[0]}.tmp"))
    f, r = IO.load_HDXer_contacts(d)
    assert r == resids
    np.testing.assert_array_equal(f.heavy_contacts, h)
    np.testing.assert_array_equal(f.acceptor_contacts, a)
    os.remove(os.path.join(d, f"Hbonds_chain_0_res_{resids[3]}.tmp"))
    with pytest.raises(FileNotFoundError):
        IO.load_HDXer_contacts(d)


def test_kints_segs_dfrac(tmp_path):
    p = tmp_path / "k.dat"
    p.write_text("# resid rate\n2 0.5\n3 12.25\n\nbad line here\n5 1e-3\n")
    k, r = IO.load_HDXer_kints(str(p))
    assert r == [2, 3, 5] and np.allclose(k, [0.5, 12.25, 1e-3])
    _, _, resids, segs = _bpti()
    seg_list = [(int(a), int(b)) for a, b in segs]
    S = IO.segments_to_map(seg_list, resids)
    assert S.shape == (len(seg_list), len(resids))
    nz = S.sum(1) > 0
    np.testing.assert_allclose(S[nz].sum(1), 1.0, rtol=1e-6)
    fr = np.random.default_rng(0).uniform(0, 1, (len(seg_list), 3)).astype(np.float32)
    IO.save_dfrac_file(str(tmp_path / "d.dat"), fr, seg_list, [0.167, 1.0, 10.0])
    fr2, segs2, tp = IO.load_HDXer_dfrac(str(tmp_path / "d.dat"), expected_timepoints=[0.167, 1.0, 10.0])
    assert segs2 == seg_list and tp == [0.167, 1.0, 10.0]
    np.testing.assert_allclose(fr2, fr, atol=1e-5)
    with pytest.raises(ValueError, match="Expected 2 timepoints"):
        IO.load_HDXer_dfrac(str(tmp_path / "d.dat"), expected_timepoints=[1.0, 2.0])


REF_BPTI = "/root/reference/notebooks_DEPRECATED/CrossValidation/BPTI/HDXer/BPTI_TFES_RW_bench_R3_k_sequence/train_BPTI_TFES_1"


@pytest.mark.skipif(not os.path.isdir(REF_BPTI), reason="reference tree only exists in the builder container")
def test_loader_reads_the_reference_bpti_fixture():
    """the HDXer files the reference itself ships -> the committed golden fixture (tests/golden/make_golden.py)"""
    h, a, resids, _ = _bpti()
    f, r = IO.load_HDXer_contacts(REF_BPTI)
    assert r == resids
    np.testing.assert_array_equal(f.heavy_contacts, h)
    np.testing.assert_array_equal(f.acceptor_contacts, a)
