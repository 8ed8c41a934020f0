"""Host-only tests: the h5py-like store and the facade's argument validation / error behaviour
(everything the reference checks before it computes, nabo/_mapping.py:296-311, 495-524, 581-612)."""
import os

import numpy as np
import pytest

from nabo_b200 import store


def test_store_roundtrip(tmp_path):
    fn = str(tmp_path / "a.h5")
    h = store.File(fn, "w")
    g = h.create_group("grp")
    g.create_dataset("b", data=np.arange(5))
    g.create_dataset("a", data=[b"x", b"yy"])
    h.create_dataset("deep/er/ds", shape=(3,), dtype=np.float64)
    h.create_row_group("rows", ["c2", "c1", "c3"], np.arange(12).reshape(3, 4))
    dt = np.dtype([("idx", np.uint32), ("val", np.int64)])
    s = np.zeros(2, dtype=dt)
    s["idx"], s["val"] = [1, 4], [7, 9]
    h.create_dataset("cell_data/C1", data=s)
    h.close()
    assert os.path.getsize(fn) > 0
    h = store.File(fn, "r")
    assert list(h["grp"]) == ["a", "b"]                      # name-sorted iteration, as HDF5
    assert h["grp/a"][1] == b"yy" and h["grp"]["b"][:].tolist() == [0, 1, 2, 3, 4]
    assert "deep/er/ds" in h and "deep/xx" not in h and h["deep/er/ds"].shape == (3,)
    assert list(h["rows"]) == ["c1", "c2", "c3"]
    assert h["rows"]["c1"][:2].tolist() == [4, 5] and h["rows/c3"][:].tolist() == [8, 9, 10, 11]
    assert h["cell_data"]["C1"]["idx"].tolist() == [1, 4]
    with pytest.raises(KeyError):
        h["nope"]
    h.close()
    h = store.File(fn, "a")
    del h["grp"]
    h["deep/er/ds"][[0, 2]] = [1.5, 2.5]
    h.close()
    h = store.File(fn, "r")
    assert "grp" not in h and h["deep/er/ds"][:].tolist() == [1.5, 0.0, 2.5]
    h.close()
    with pytest.raises(OSError):
        store.File(str(tmp_path / "missing.h5"), "r")


def _pca_file(fn, names, mat):
    h = store.File(fn, "w")
    h.create_row_group("data", names, mat)
    h.close()


def test_mapping_validation(tmp_path):
    from nabo_b200.mapping import Mapping
    ref_fn, map_fn = str(tmp_path / "ref.h5"), str(tmp_path / "map.h5")
    names = ["R%03d" % i for i in range(20)]
    _pca_file(ref_fn, names[::-1], np.random.default_rng(0).normal(size=(20, 6)))
    with pytest.raises(ValueError, match="Underscores"):
        Mapping(map_fn, "a__b", ref_fn, "data")
    with pytest.raises(ValueError, match="cannot be same"):
        Mapping(ref_fn, "REF", ref_fn, "data")
    with pytest.raises(ValueError, match="doesn't exist"):
        Mapping(map_fn, "REF", str(tmp_path / "none.h5"), "data")
    with pytest.raises(ValueError, match="does not exist"):
        Mapping(map_fn, "REF", ref_fn, "nogroup")
    m = Mapping(map_fn, "REF", ref_fn, "data", overwrite=True)
    assert m.refName == "REF" and m.refCells == names          # bytewise name order, not insertion order
    with pytest.raises(ValueError, match="set the parameters"):
        m.calc_dist(ref_fn, "data", "d", "s", [])
    with pytest.raises(ValueError, match="Set parameters"):
        m.calc_snn("x", "REF", "g")
    for bad in (0, -1.0, "abc"):
        with pytest.raises(ValueError, match="dist_factor"):
            m.set_parameters(5, 3, bad, 100)
    m.set_parameters(5, 3, 0.25, 100)
    with pytest.raises(ValueError, match="same as that of reference"):
        m.map_target("T", ref_fn, "data")
    with pytest.raises(ValueError, match="cannot be same"):
        m.map_target("T", map_fn, "data")
    with pytest.raises(ValueError, match="same as reference name"):
        m.map_target("REF", str(tmp_path / "t.h5"), "data")
    with pytest.raises(ValueError, match="Underscores"):
        m.map_target("T__1", str(tmp_path / "t.h5"), "data")
    with pytest.raises(KeyError):
        m.calc_snn("missing_sortedDist", "T", "g")
    # re-attach: same ref name OK, different one refused, changed cell set refused
    m2 = Mapping(map_fn, "REF", ref_fn, "data")
    assert m2._nameStash["REF"] == m._nameStash["REF"]
    with pytest.raises(ValueError, match="different ref_name"):
        Mapping(map_fn, "OTHER", ref_fn, "data")
    other = str(tmp_path / "ref2.h5")
    _pca_file(other, names[:-1] + ["ZZZ"], np.zeros((20, 6)))
    with pytest.raises(ValueError, match="does not match"):
        Mapping(map_fn, "REF", other, "data")


def test_graph_validation(tmp_path):
    from nabo_b200.graph import Graph
    g = Graph()
    with pytest.raises(IOError):
        g.load_from_h5(str(tmp_path / "none.h5"), "REF", "reference")
    fn = str(tmp_path / "m.h5")
    store.File(fn, "w").close()
    with pytest.raises(ValueError, match="Kind"):
        g.load_from_h5(fn, "REF", "bogus")
    with pytest.raises(ValueError, match="load reference kind first"):
        g.load_from_h5(fn, "T", "target")
    with pytest.raises(KeyError, match="stashed names"):
        g.load_from_h5(fn, "REF", "reference")
    with pytest.raises(ValueError, match="not present"):
        g.get_mapping_score("T")


def test_dataset_stats_restatement_matches_reference(golden):
    """Oracle restatement of set_sf / set_gene_stats (NumPy float32 reductions) == the unmodified reference."""
    import scipy.sparse as sp
    from oracle import nabo_oracle as O
    g = golden("dataset_small")
    counts = g["counts_ref"].astype(np.float32)
    csr = sp.csr_matrix(counts)
    csr.sort_indices()
    n_cells, n_genes = counts.shape
    tot = O.size_factor_sums(csr.indptr, csr.indices, csr.data, n_genes, np.arange(n_genes))
    tot[tot == 0] = 1
    sf = (1000.0 / tot).astype(np.float32)
    assert np.array_equal(sf, g["sf_ref"])
    csc = csr.tocsc()
    csc.sort_indices()
    m, nzm, var, nc = O.gene_stats(csc.indptr, csc.indices, csc.data, n_cells, np.arange(n_cells), sf)
    valid = nc > 0
    gm = g["gene_m_ref"]
    assert np.array_equal(m[valid].astype(np.float64), gm[valid])
    gi = g["gene_idx"]
    assert np.array_equal(m[gi].astype(np.float64), g["mu"])
    assert np.array_equal(np.sqrt(var[gi].astype(np.float64)), g["sigma"])


def test_reopen_uses_cached_content_and_notices_foreign_writes(tmp_path):
    """Reopening a file this process wrote is served from the content cache; a file replaced behind the
    cache's back (different mtime / size) is parsed again."""
    import os
    import shutil
    from nabo_b200 import store
    fn, other = str(tmp_path / "a.h5"), str(tmp_path / "b.h5")
    with store.File(fn, "w") as h:
        h.create_row_group("data", ["c1", "c2"], np.arange(6.0).reshape(2, 3))
        h.create_dataset("meta/k", data=np.array([7]))
    with store.File(fn, "a") as h:                         # cache hit, then a write through it
        assert h["data"]["c2"][:].tolist() == [3.0, 4.0, 5.0]
        h.create_dataset("meta/extra", data=np.array([1, 2, 3]))
    with store.File(fn, "r") as h:
        assert h["meta/extra"][:].tolist() == [1, 2, 3] and h["meta/k"][0] == 7
    with store.File(other, "w") as h:
        h.create_dataset("only", data=np.array([42]))
    shutil.copyfile(other, fn)                             # foreign writer
    os.utime(fn, ns=(1, 1))
    with store.File(fn, "r") as h:
        assert "only" in h and "data" not in h


def test_batched_writes_defer_and_replay(tmp_path):
    """Inside batched_writes every close only records the content (re-opens see it), the file is written once
    at the end; nested blocks join the outer one; a block without writes leaves the file alone."""
    import os
    from nabo_b200 import store
    fn = str(tmp_path / "m.h5")
    with store.File(fn, "w") as h:
        h.create_dataset("a", data=np.array([1]))
    before = os.stat(fn).st_mtime_ns
    with store.batched_writes(fn):
        with store.File(fn, "a") as h:
            h.create_dataset("b", data=np.array([2]))
        with store.batched_writes(fn):                        # nested: same batch
            with store.File(fn, "a") as h:
                assert h["b"][0] == 2                         # pending content is visible
                h.create_row_group("rows", ["x", "y"], np.eye(2))
        assert os.stat(fn).st_mtime_ns == before              # nothing written yet
        with store.File(fn, "r") as h:
            assert "rows" in h and h["rows"]["y"][:].tolist() == [0.0, 1.0]
    assert os.stat(fn).st_mtime_ns != before
    store._CONTENT_CACHE.clear()                              # force a real parse of what was written
    with store.File(fn, "r") as h:
        assert h["a"][0] == 1 and h["b"][0] == 2 and h["rows"]["x"][:].tolist() == [1.0, 0.0]
    new = str(tmp_path / "n.h5")
    with store.batched_writes(new):                           # a file born inside a batch
        with store.File(new, "w") as h:
            h.create_dataset("z", data=np.array([9]))
        with store.File(new, "a") as h:
            assert h["z"][0] == 9
    store._CONTENT_CACHE.clear()
    with store.File(new, "r") as h:
        assert h["z"][0] == 9
    stamp = os.stat(fn).st_mtime_ns
    with store.batched_writes(fn):
        with store.File(fn, "r") as h:
            assert "a" in h
    assert os.stat(fn).st_mtime_ns == stamp


def test_in_place_dataset_writes_persist_and_read_only_files_refuse_them(tmp_path):
    """h5['g/x'][i] = v on a file opened 'a' / 'r+' reaches the disk at the next flush (h5py behaviour); the same on a
    file opened 'r' raises instead of leaking into the content cache."""
    from nabo_b200 import store
    fn = str(tmp_path / "w.h5")
    h = store.File(fn, "w")
    h.create_group("g").create_dataset("x", data=np.zeros(4))
    h.create_row_group("rows", ["a", "b"], np.zeros((2, 3)))
    h.close()
    h = store.File(fn, "r+")
    h["g/x"][1] = 5.0
    h["rows"]["b"][2] = 7.0
    h.close()
    store._CONTENT_CACHE.clear()                     # force a real read of the container
    h = store.File(fn, "r")
    assert h["g/x"][:].tolist() == [0.0, 5.0, 0.0, 0.0] and h["rows"]["b"][:].tolist() == [0.0, 0.0, 7.0]
    with pytest.raises(OSError, match="read-only"):
        h["g/x"][0] = 1.0
    with pytest.raises(OSError, match="read-only"):
        h["rows"]["a"][0] = 1.0
    h.close()
    assert store.File(fn, "r")["g/x"][0] == 0.0


def test_content_cache_is_bounded(tmp_path):
    from nabo_b200 import store
    store._CONTENT_CACHE.clear()
    for i in range(store._CONTENT_CACHE_MAX + 5):
        h = store.File(str(tmp_path / ("f%d.h5" % i)), "w")
        h.create_dataset("x", data=np.arange(3))
        h.close()
    assert len(store._CONTENT_CACHE) <= store._CONTENT_CACHE_MAX


def test_weight_table_for_k2_is_lazy():
    """k = 2: upstream divides by zero only if a pair really shares both neighbours (nabo/_mapping.py:194)."""
    from nabo_b200 import core
    lut = core.snn_weight_lut(2, strict=False)
    assert lut[0] == 0.0 and lut[1] == 1.0 and np.isnan(lut[2])
    with pytest.raises(ZeroDivisionError):
        core.snn_weight_lut(2)
    assert core.snn_int_weights(2).tolist() == [0, 100, 0]
    assert core.snn_weight_lut(1).tolist() == [0.0, -1.0]


# ----------------------------------------------------------------------------- h5py backend of store.open_file
class _FakeH5:
    """Minimal in-memory stand-in with h5py's object model (Group with .attrs / create_group / create_dataset /
    [] with a/b paths / del / keys, Dataset with slicing) so that the H5File adapter is exercised in this image,
    where h5py itself is absent.  The real-HDF5 round trip below runs wherever h5py is importable."""
    files = {}

    class Dataset:
        def __init__(self, data, name):
            self._a, self.name = np.array(data), name
        shape = property(lambda s: s._a.shape)
        dtype = property(lambda s: s._a.dtype)
        def __len__(self): return len(self._a)
        def __getitem__(self, k): return self._a[k]
        def __setitem__(self, k, v): self._a[k] = v
        def __iter__(self): return iter(self._a)
        def __array__(self, dtype=None, copy=None): return self._a if dtype is None else self._a.astype(dtype)

    class Group:
        def __init__(self, name="/"):
            self.name, self.attrs, self._c = name, {}, {}
        def _walk(self, path, create=False):
            node = self
            parts = [p for p in path.split("/") if p]
            for p in parts[:-1]:
                if p not in node._c:
                    if not create:
                        raise KeyError(path)
                    node._c[p] = _FakeH5.Group(node.name.rstrip("/") + "/" + p)
                node = node._c[p]
            return node, parts[-1]
        def __contains__(self, path):
            try:
                node, leaf = self._walk(path)
            except KeyError:
                return False
            return leaf in node._c
        def __getitem__(self, path):
            node, leaf = self._walk(path)
            return node._c[leaf]
        def __delitem__(self, path):
            node, leaf = self._walk(path)
            del node._c[leaf]
        def __len__(self): return len(self._c)
        def keys(self): return list(self._c.keys())
        def create_group(self, path):
            node, leaf = self._walk(path, True)
            if leaf in node._c:
                raise ValueError("exists")
            node._c[leaf] = _FakeH5.Group(node.name.rstrip("/") + "/" + leaf)
            return node._c[leaf]
        def create_dataset(self, path, shape=None, dtype=None, data=None, **kw):
            node, leaf = self._walk(path, True)
            if leaf in node._c:
                raise ValueError("exists")
            arr = np.zeros(shape, dtype) if data is None else np.array(data, dtype=dtype)
            assert arr.dtype.kind != "U", "h5py cannot store fixed-width unicode"
            node._c[leaf] = _FakeH5.Dataset(arr, node.name.rstrip("/") + "/" + leaf)
            return node._c[leaf]

    class File(Group):
        def __init__(self, fn, mode="r"):
            super().__init__("/")
            if mode in ("r", "r+") and fn not in _FakeH5.files:
                raise OSError("no such file")
            if mode == "w" or fn not in _FakeH5.files:
                _FakeH5.files[fn] = ({}, {})
                open(fn, "ab").close()
            self._c, self.attrs = _FakeH5.files[fn]
        def flush(self): pass
        def close(self): pass


def _h5_roundtrip(tmp_path, monkeypatch, module):
    from nabo_b200 import store
    monkeypatch.setattr(store, "_h5py", lambda: module)
    fn = str(tmp_path / "real.h5")
    h = store.open_file(fn, "w")
    assert isinstance(h, store.H5File)
    h.create_group("name_stash").create_dataset("ref_name", data=[b"REF", b"abc"])
    h["name_stash"].create_dataset("target_names", data=[[b"T1", b"u1"], [b"T2", b"u2"]])
    h.create_dataset("g/x", data=np.arange(6.0).reshape(2, 3))
    h.create_dataset("g/names", data=np.array(["b", "a"]))                 # unicode in -> bytes stored
    h.create_row_group("uid_sortedDist", ["c2", "c1"], np.array([[3, 4], [5, 6]], dtype=np.int64))
    node = h.create_group("uid_graph")                                      # the reference's per-node layout
    node.create_dataset("c1_T", data=np.array([(b"c9_REF", 0.25), (b"c8_REF", 0.5)]))
    h.flush()
    h.close()
    h = store.open_file(fn, "r")
    assert h["name_stash/ref_name"][0] == b"REF" and h["name_stash/target_names"][:][1][1] == b"u2"
    assert h["g/x"][:].tolist() == [[0, 1, 2], [3, 4, 5]] and h["g/names"][:].tolist() == [b"b", b"a"]
    assert "uid_sortedDist" in h and "nope" not in h and list(h["g"]) == ["names", "x"]
    rg = h["uid_sortedDist"]
    assert isinstance(rg, store.RowGroup) and list(rg) == ["c1", "c2"] and rg["c2"][:2].tolist() == [3, 4]
    row = h["uid_graph/c1_T"][:]
    assert row.dtype.kind == "S" and row[0][0] == b"c9_REF" and float(row[1][1]) == 0.5
    h.close()
    h = store.open_file(fn, "a")
    del h["g/x"]
    assert "g/x" not in h
    h.close()


def test_h5py_backend_adapter_with_stand_in(tmp_path, monkeypatch):
    _FakeH5.files.clear()
    _h5_roundtrip(tmp_path, monkeypatch, _FakeH5)
    # a file that already is a zip/npy container stays one even where h5py exists
    from nabo_b200 import store
    monkeypatch.setattr(store, "_h5py", lambda: None)
    fn = str(tmp_path / "container.h5")
    h = store.open_file(fn, "w")
    h.create_dataset("x", data=np.arange(3))
    h.close()
    monkeypatch.setattr(store, "_h5py", lambda: _FakeH5)
    assert isinstance(store.open_file(fn, "r"), store.File)


def test_h5py_backend_real_hdf5(tmp_path, monkeypatch):
    h5py = pytest.importorskip("h5py")
    _h5_roundtrip(tmp_path, monkeypatch, h5py)
