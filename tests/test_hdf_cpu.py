"""HDF5 optimisation histories (utils/hdf.py, reference jaxent/src/utils/hdf.py) and the built-in HDF5 writer / reader.

h5py is absent from the build image, so the byte-level checks are: (1) the reader parses a file written by the real HDF5
library (scipy ships one MATLAB v7.3 file: version-0 superblock, symbol-table group, local heap, B-tree, attributes),
(2) the writer's files are read back by that reader, (3) the on-disk structures follow the conventions seen in (1)."""
import glob
import os
import struct

import numpy as np
import pytest

from jaxent_b200.interfaces.simulation import Simulation_Parameters
from jaxent_b200.models.HDX.BV.parameters import BV_Model_Parameters
from jaxent_b200.opt.base import LossComponents, OptimizationHistory, OptimizationState
from jaxent_b200.utils import _h5mini, hdf


def _real_hdf5_file():
    try:
        import scipy.io.matlab as m
    except Exception:  # noqa: BLE001
        return None
    hits = glob.glob(os.path.join(os.path.dirname(m.__file__), "tests", "data", "testhdf5_7.4_GLNX86.mat"))
    return hits[0] if hits else None


def test_reader_parses_a_file_written_by_the_hdf5_library():
    path = _real_hdf5_file()
    if path is None:
        pytest.skip("scipy's HDF5 test file is not installed")
    f = _h5mini.File(path, "r")
    assert f.keys() == ["testdouble"]
    ds = f["testdouble"]
    assert ds.shape == (9, 1) and ds.dtype == np.float64
    np.testing.assert_allclose(ds[()][:, 0], np.arange(9) * np.pi / 4, rtol=1e-15)
    assert ds.attrs["MATLAB_class"] == "double"  # fixed-length string attribute


def test_writer_reader_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    path = str(tmp_path / "t.h5")
    with _h5mini.File(path, "w") as f:
        g = f.create_group("a/b")
        g.attrs["text"] = "héllo"
        g.attrs["flag"] = True
        g.attrs["off"] = np.bool_(False)
        g.attrs["step"] = 12345678901
        g.attrs["vec"] = np.arange(3, dtype=np.float32)
        f.create_dataset("a/x", data=rng.standard_normal((5, 7)).astype(np.float32), compression="gzip", chunks=True)
        f.create_dataset("a/b/scalar", data=np.float32(2.5))
        f.create_dataset("a/b/i64", data=np.arange(-3, 4, dtype=np.int64))
        f.create_dataset("a/b/u8", data=np.arange(5, dtype=np.uint8))
        f.create_dataset("a/b/f64", data=rng.standard_normal(4))
        f.create_dataset("a/b/empty", data=np.zeros((0,), np.float32))
        f.create_dataset("a/b/mask", data=np.array([True, False, True]))
        many = f.create_group("many")
        for i in range(300):  # > 256 children: two B-tree levels
            many.create_dataset(f"{i}", data=np.float32(i))
        with pytest.raises(ValueError):
            f.create_group("a/b")
    r = _h5mini.File(path, "r")
    g = r["a/b"]
    assert g.attrs["text"] == "héllo" and g.attrs["flag"] is np.True_ and not g.attrs["off"] and int(g.attrs["step"]) == 12345678901
    np.testing.assert_array_equal(g.attrs["vec"], np.arange(3, dtype=np.float32))
    rng = np.random.default_rng(0)
    np.testing.assert_array_equal(r["a/x"][()], rng.standard_normal((5, 7)).astype(np.float32))
    assert r["a/b/scalar"][()] == np.float32(2.5) and r["a/b/scalar"].shape == ()
    np.testing.assert_array_equal(r["a/b/i64"][()], np.arange(-3, 4))
    assert r["a/b/u8"].dtype == np.uint8 and r["a/b/f64"].dtype == np.float64 and r["a/b/empty"].shape == (0,)
    np.testing.assert_array_equal(r["a/b/mask"][()], [True, False, True])
    assert len(r["many"]) == 300 and all(r["many"][f"{i}"][()] == i for i in (0, 9, 10, 99, 255, 256, 299))
    assert "a/b/nope" not in r and "a/b/u8" in r
    with pytest.raises(KeyError):
        r["a/zzz"]


def test_written_file_structures(tmp_path):
    """superblock / heap / B-tree conventions of the real library (cf. the MATLAB file: free list ends with 1, sizes multiple of 8)"""
    path = str(tmp_path / "s.h5")
    with _h5mini.File(path, "w") as f:
        f.create_dataset("only", data=np.arange(4, dtype=np.float32))
        f.attrs["note"] = "x"
    d = open(path, "rb").read()
    assert d[:8] == b"\x89HDF\r\n\x1a\n" and len(d) % 8 == 0
    v_sb, v_fs, v_root, _, v_sh, so, sl, _, leaf_k, int_k, flags = struct.unpack_from("<BBBBBBBBHHI", d, 8)
    assert (v_sb, v_fs, v_root, v_sh, so, sl, leaf_k, int_k, flags) == (0, 0, 0, 0, 8, 8, 4, 16, 0)
    base, free, eof, drv = struct.unpack_from("<QQQQ", d, 24)
    assert base == 0 and free == _h5mini.UNDEF and drv == _h5mini.UNDEF and eof == len(d)
    name_off, hdr, cache, _res, btree, heap = struct.unpack_from("<QQIIQQ", d, 56)
    assert name_off == 0 and cache == 1
    assert d[btree:btree + 4] == b"TREE" and d[heap:heap + 4] == b"HEAP"
    node_type, level, used, left, right = struct.unpack_from("<BBHQQ", d, btree + 4)
    assert (node_type, level, used, left, right) == (0, 0, 1, _h5mini.UNDEF, _h5mini.UNDEF)
    _ver, seg_size, free_off, seg = struct.unpack_from("<B3xQQQ", d, heap + 4)
    assert d[seg:seg + 8] == b"\0" * 8 and d[seg + 8:seg + 13] == b"only\0"
    nxt, fsize = struct.unpack_from("<QQ", d, seg + free_off)
    assert nxt == 1 and free_off + fsize == seg_size and fsize >= 16
    key0, snod, key1 = struct.unpack_from("<QQQ", d, btree + 24)
    assert key0 == 0 and key1 == 8 and d[snod:snod + 4] == b"SNOD" and struct.unpack_from("<BBH", d, snod + 4) == (1, 0, 1)
    ver, _r, nmsg, refs, size = struct.unpack_from("<BBHII", d, hdr)
    assert (ver, refs) == (1, 1) and nmsg == 2 and size % 8 == 0  # symbol table message + one attribute
    assert b"GCOL" in d  # the string attribute lives in a global heap collection (variable-length UTF-8, as h5py writes it)
    col = d.index(b"GCOL")
    assert struct.unpack_from("<B3xQ", d, col + 4)[1] >= 4096


def test_reader_chunked_deflate_shuffle_dataset(tmp_path):
    """the reference writes gzip-compressed chunked datasets; parity unpinned (no h5py here to produce one): the bytes below
    are assembled by hand from the format specification"""
    import zlib

    a = np.arange(35, dtype=np.float32).reshape(5, 7)
    w = _h5mini._Writer()
    cdims = (2, 4)
    entries = []
    for i in range(0, 5, 2):
        for j in range(0, 7, 4):
            chunk = np.zeros(cdims, np.float32)
            blk = a[i:i + 2, j:j + 4]
            chunk[:blk.shape[0], :blk.shape[1]] = blk
            raw = np.frombuffer(chunk.tobytes(), np.uint8).reshape(-1, 4).T.tobytes()  # shuffle
            comp = zlib.compress(raw, 9)
            entries.append((len(comp), (i, j, 0), w.alloc(comp)))
    node = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(entries), _h5mini.UNDEF, _h5mini.UNDEF)
    for size, off, addr in entries:
        node += struct.pack("<IIQQQ", size, 0, *off) + struct.pack("<Q", addr)
    node += struct.pack("<IIQQQ", 0, 0, 6, 8, 0)
    bt = w.alloc(node)
    pipeline = struct.pack("<BB6x", 1, 2) + struct.pack("<HHHHI4x", 2, 0, 0, 1, 4) + struct.pack("<HHHHI4x", 1, 0, 0, 1, 9)
    layout = struct.pack("<BBBQIII", 3, 2, 3, bt, 2, 4, 4)
    hdr = w._object_header([(0x0001, _h5mini._dataspace_message(a.shape)), (0x0003, _h5mini._dtype_message(a.dtype)),
                            (0x000B, pipeline), (0x0008, layout)])
    rd = _h5mini._Reader(bytes(w.buf))
    np.testing.assert_array_equal(rd._object(hdr)[()], a)


def _history(n_states=3, with_best=True):
    rng = np.random.default_rng(1)
    states = []
    for i in range(n_states):
        mp = BV_Model_Parameters(bv_bc=np.array([0.35 + i], np.float32), bv_bh=np.array([2.0 - i], np.float32))
        w = rng.random(11).astype(np.float32)
        sp = Simulation_Parameters(frame_weights=w / w.sum(), frame_mask=np.ones(11, np.float32), model_parameters=[mp],
                                   forward_model_weights=np.array([0.5, 0.5], np.float32), normalise_loss_functions=np.array([1.0, 0.0], np.float32),
                                   forward_model_scaling=np.array([1000.0, 1.0], np.float32))
        lc = LossComponents(rng.random(2).astype(np.float32), rng.random(2).astype(np.float32), rng.random(2).astype(np.float32),
                            rng.random(2).astype(np.float32), np.float32(rng.random()), np.float32(rng.random()))
        states.append(OptimizationState(sp, None, 10 * i, lc, None))
    h = OptimizationHistory(states=states)
    if with_best:
        h.get_best_state()
    return h


def test_history_file_round_trip(tmp_path):
    h = _history()
    path = str(tmp_path / "hist.h5")
    hdf.save_optimization_history_to_file(path, h, compress=True)
    g = hdf.load_optimization_history_from_file(path)
    assert len(g.states) == 3 and g.best_state is not None
    for a, b in zip(h.states + [h.best_state], g.states + [g.best_state]):
        assert int(a.step) == b.step and b.opt_state is None
        for name in hdf._PARAM_ARRAYS:
            np.testing.assert_array_equal(getattr(a.params, name), getattr(b.params, name))
        ma, mb = a.params.model_parameters[0], b.params.model_parameters[0]
        assert isinstance(mb, BV_Model_Parameters)
        np.testing.assert_array_equal(ma.bv_bc, mb.bv_bc)
        np.testing.assert_array_equal(ma.bv_bh, mb.bv_bh)
        for name in LossComponents._fields:
            np.testing.assert_array_equal(np.asarray(getattr(a.losses, name)), np.asarray(getattr(b.losses, name)))
    # same best-state rule after a reload (first validation component, reference opt/base.py:180-189)
    assert g.get_best_state().step == h.best_state.step
    # explicit class, as the reference's analysis scripts pass it
    g2 = hdf.load_optimization_history_from_file(path, default_model_params_cls=BV_Model_Parameters)
    assert len(g2.states) == 3


def test_history_layout_matches_the_reference(tmp_path):
    """group / dataset / attribute names of reference utils/hdf.py:134-360, class_info under the reference's module path"""
    path = str(tmp_path / "hist.h5")
    hdf.save_optimization_history_to_file(path, _history(2, with_best=False), compress=False)
    f = _h5mini.File(path, "r") if hdf.backend() == "builtin" else __import__("h5py").File(path, "r")
    top = f["optimization_history"]
    assert not top.attrs["has_best_state"] and sorted(top.keys()) == ["states"]
    st = top["states/1"]
    assert int(st.attrs["step"]) == 10 and st.attrs["has_losses"]
    assert sorted(st.keys()) == ["losses", "params"]
    assert sorted(st["params"].keys()) == sorted(list(hdf._PARAM_ARRAYS) + ["model_parameters"])
    mp = st["params/model_parameters/0"]
    assert mp.attrs["class_info"] == "jaxent.src.models.HDX.BV.parameters.BV_Model_Parameters"
    assert sorted(mp.keys()) == ["bv_bc", "bv_bh"] and "HDX" in mp.attrs["key"]
    assert sorted(st["losses"].keys()) == sorted(LossComponents._fields)
    assert st["losses/total_train_loss"].shape in ((), (1,))


def test_empty_history_and_missing_losses(tmp_path):
    path = str(tmp_path / "e.h5")
    hdf.save_optimization_history_to_file(path, OptimizationHistory())
    g = hdf.load_optimization_history_from_file(path)
    assert g.states == [] and g.best_state is None
    h = _history(1, with_best=False)
    h.states[0] = h.states[0]._replace(losses=None)
    hdf.save_optimization_history_to_file(path, h)
    assert hdf.load_optimization_history_from_file(path).states[0].losses is None


def test_random_trees_round_trip(tmp_path):
    """seeded random group trees: names that sort differently as bytes / numbers, all supported dtypes, attributes everywhere"""
    rng = np.random.default_rng(7)
    dtypes = [np.float32, np.float64, np.int8, np.int32, np.int64, np.uint16, np.uint64, np.bool_]

    def fill(g, depth):
        spec = {}
        for k in range(int(rng.integers(0, 12))):
            name = f"{'n' if rng.random() < 0.5 else ''}{int(rng.integers(0, 1000))}_{k}"
            if depth < 3 and rng.random() < 0.3:
                spec[name] = fill(g.create_group(name), depth + 1)
            else:
                dt = dtypes[int(rng.integers(len(dtypes)))]
                shape = tuple(int(x) for x in rng.integers(0, 5, size=int(rng.integers(0, 4))))
                a = (rng.random(shape) * 100).astype(dt)
                g.create_dataset(name, data=a)
                spec[name] = a
        g.attrs["depth"] = depth
        g.attrs["label"] = f"level-{depth}-é"
        g.attrs["vals"] = rng.random(3)
        return spec

    def check(g, spec, depth):
        assert sorted(g.keys()) == sorted(spec) and int(g.attrs["depth"]) == depth and g.attrs["label"] == f"level-{depth}-é"
        assert g.attrs["vals"].shape == (3,)
        for name, want in spec.items():
            if isinstance(want, dict):
                check(g[name], want, depth + 1)
            else:
                got = g[name][()]
                assert np.asarray(got).dtype == want.dtype and np.asarray(got).shape == want.shape
                np.testing.assert_array_equal(got, want)

    for trial in range(5):
        path = str(tmp_path / f"r{trial}.h5")
        with _h5mini.File(path, "w") as f:
            spec = fill(f, 0)
        check(_h5mini.File(path, "r"), spec, 0)
