"""Hierarchical array store with the h5py call surface the nabo path uses.

The reference keeps every inter-layer hand-off in HDF5 files (dataset file, PCA
file, mapping file).  ``h5py``/libhdf5 are absent from the target
image, so the facade goes through this module: ``open_file`` returns ``File``
below, which exposes the same group/dataset names and the subset of the h5py API
the path touches (``in``, ``[]`` with ``a/b`` paths, ``del``, ``create_group``,
``create_dataset``, name-sorted iteration, ``flush``/``close``, dataset slicing,
field access, fancy-index assignment).

Two additions make the layout scale beyond one-dataset-per-cell:

* ``RowGroup`` - a group whose members ``grp[name]`` are the rows of ONE 2-D
  array (``create_row_group``).  It reads like the reference's per-cell groups
  (``<uid>_sortedDist/<cell>[:k]``, nabo/_mapping.py:190) but is stored columnar.
* a file is a single ``.npz``-style zip container holding one ``.npy`` per
  dataset plus a JSON manifest; no pickling.

Iteration order is bytewise name order, as in HDF5 (nabo/_mapping.py:79,
403-405 depend on it: neighbour indices are positions in that order).
"""
from __future__ import annotations

import io
import json
import os
import zipfile
from typing import Dict, Iterator, List, Optional

import numpy as np

__all__ = ["File", "Group", "Dataset", "RowGroup", "open_file", "MEMORY_FILES",
           "set_memory_only"]

# fn -> root Group for files held only in memory (used by the golden generator so
# that the unmodified reference can run against this module as its "h5py").
MEMORY_FILES: Dict[str, "Group"] = {}
_MEMORY_ONLY = False


def set_memory_only(flag: bool) -> None:
    """Keep files in process memory (a zero-byte placeholder is still touched on
    disk so ``os.path.exists`` checks, nabo/_mapping.py:331, keep working)."""
    global _MEMORY_ONLY
    _MEMORY_ONLY = bool(flag)


class Dataset:
    """A named array.  Mirrors the slice of ``h5py.Dataset`` the path uses."""

    def __init__(self, data, name: str = "", parent: Optional["Group"] = None):
        self._a = data if isinstance(data, np.ndarray) else np.asarray(data)
        self.name = name
        self._parent = parent          # the group holding it: its _root is the file to mark dirty on writes

    # -- numpy-ish protocol
    @property
    def shape(self):
        return self._a.shape

    @property
    def dtype(self):
        return self._a.dtype

    @property
    def size(self):
        return self._a.size

    def __len__(self):
        return len(self._a)

    def __array__(self, dtype=None, copy=None):
        return self._a if dtype is None else self._a.astype(dtype)

    def __iter__(self):
        return iter(self._a)

    def __getitem__(self, key):
        if isinstance(key, tuple) and len(key) == 0:
            return self._a[()]
        return self._a[key]

    def __setitem__(self, key, value):
        """In-place write (``h5['g/x'][i] = v``): refused on a file opened 'r', persisted by the next flush."""
        root = self._parent._root if self._parent is not None else None
        if root is not None and getattr(root, "mode", "a") == "r":
            raise OSError("Can't write data (file %r was opened read-only)" % getattr(root, "filename", ""))
        self._a[key] = value
        if root is not None:
            root._dirty = True

    def __repr__(self):
        return "<store.Dataset %r shape %s dtype %s>" % (self.name, self.shape, self.dtype)


class Group:
    """Name -> Group | Dataset mapping with h5py path semantics."""

    def __init__(self, name: str = "/", root: Optional["Group"] = None):
        self.name = name
        self._c: Dict[str, object] = {}
        self._root = root if root is not None else self

    # -- path helpers
    def _walk(self, path: str, create: bool = False):
        parts = [p for p in path.split("/") if p]
        node = self
        for p in parts[:-1]:
            if not isinstance(node, Group) or p not in node._children():
                if create and isinstance(node, Group):
                    node = node.create_group(p)
                    continue
                raise KeyError("Unable to open object (component %r of %r not found)" % (p, path))
            node = node._children()[p]
        return node, (parts[-1] if parts else "")

    def _children(self) -> Dict[str, object]:
        return self._c

    def _touch(self):
        self._root._dirty = True

    # -- mapping protocol
    def __contains__(self, path) -> bool:
        if not isinstance(path, str):
            return False
        try:
            node, leaf = self._walk(path)
        except KeyError:
            return False
        return isinstance(node, Group) and (leaf == "" or leaf in node._children())

    def __getitem__(self, path: str):
        node, leaf = self._walk(path)
        if leaf == "":
            return node
        try:
            return node._children()[leaf]
        except (KeyError, AttributeError):
            raise KeyError("Unable to open object (object %r doesn't exist)" % path)

    def __delitem__(self, path: str):
        node, leaf = self._walk(path)
        if leaf not in node._children():
            raise KeyError("Couldn't delete link (name %r doesn't exist)" % path)
        node._delete(leaf)
        self._touch()

    def _delete(self, leaf: str):
        del self._c[leaf]

    def __iter__(self) -> Iterator[str]:
        return iter(sorted(self._children().keys()))

    def __len__(self):
        return len(self._children())

    def keys(self):
        return sorted(self._children().keys())

    def values(self):
        c = self._children()
        return [c[k] for k in sorted(c.keys())]

    def items(self):
        c = self._children()
        return [(k, c[k]) for k in sorted(c.keys())]

    # -- creation
    def create_group(self, path: str) -> "Group":
        node, leaf = self._walk(path, create=True)
        if leaf in node._children():
            raise ValueError("Unable to create group (name already exists)")
        g = Group(node.name.rstrip("/") + "/" + leaf, self._root)
        node._c[leaf] = g
        self._touch()
        return g

    def require_group(self, path: str) -> "Group":
        return self[path] if path in self else self.create_group(path)

    def create_dataset(self, path: str, shape=None, dtype=None, data=None, **_ignored) -> Dataset:
        node, leaf = self._walk(path, create=True)
        if leaf in node._children():
            raise ValueError("Unable to create dataset (name already exists)")
        if data is None:
            arr = np.zeros(shape, dtype=dtype if dtype is not None else np.float32)
        else:
            arr = np.array(data, dtype=dtype) if not isinstance(data, np.ndarray) else \
                (data.astype(dtype) if dtype is not None and data.dtype != dtype else np.array(data))
            if arr.dtype.kind == "U":          # h5py stores str as bytes on this path
                arr = np.char.encode(arr, "utf-8")
            if shape is not None and tuple(np.atleast_1d(shape)) != arr.shape:
                arr = arr.reshape(shape)
        ds = Dataset(arr, node.name.rstrip("/") + "/" + leaf, node)
        node._c[leaf] = ds
        self._touch()
        return ds

    def create_row_group(self, path: str, names: List[str], data: np.ndarray) -> "RowGroup":
        """Columnar group: ``grp[names[i]]`` is ``data[i]``."""
        node, leaf = self._walk(path, create=True)
        if leaf in node._children():
            raise ValueError("Unable to create group (name already exists)")
        g = RowGroup(node.name.rstrip("/") + "/" + leaf, self._root, names, np.asarray(data))
        node._c[leaf] = g
        self._touch()
        return g

    def __repr__(self):
        return "<store.Group %r (%d members)>" % (self.name, len(self))


class RowGroup(Group):
    """Group view over the rows of a 2-D array (read-only membership)."""

    def __init__(self, name, root, names, data):
        super().__init__(name, root)
        if len(names) != data.shape[0]:
            raise ValueError("row group: %d names for %d rows" % (len(names), data.shape[0]))
        self.row_names = [n.decode("utf-8") if isinstance(n, bytes) else str(n) for n in names]
        self.data = data
        self._index: Optional[Dict[str, int]] = None
        self._sorted: Optional[List[str]] = None
        self._names_enc: Optional[np.ndarray] = None      # utf-8 byte-string array of row_names (file form)

    def names_encoded(self) -> np.ndarray:
        if self._names_enc is None:
            self._names_enc = (np.char.encode(np.asarray(self.row_names, dtype=str), "utf-8")
                               if len(self.row_names) else np.zeros(0, "S1"))
        return self._names_enc

    def _idx(self) -> Dict[str, int]:
        if self._index is None:
            self._index = {n: i for i, n in enumerate(self.row_names)}
        return self._index

    class _Rows(dict):
        """Lazy child dict: builds Dataset views on demand."""

        def __init__(self, owner):
            super().__init__()
            self.o = owner

        def __contains__(self, k):
            return k in self.o._idx()

        def __getitem__(self, k):
            return Dataset(self.o.data[self.o._idx()[k]], self.o.name + "/" + k, self.o)

        def keys(self):
            return self.o._idx().keys()

        def __len__(self):
            return len(self.o.row_names)

    def _children(self):
        return RowGroup._Rows(self)

    def _delete(self, leaf):
        raise TypeError("rows of a RowGroup cannot be deleted individually")

    def __iter__(self):
        if self._sorted is None:
            self._sorted = sorted(self.row_names)
        return iter(self._sorted)

    def keys(self):
        return list(iter(self))

    def __len__(self):
        return len(self.row_names)

    def create_group(self, path):
        raise TypeError("RowGroup has fixed membership")

    def create_dataset(self, path, **kw):
        raise TypeError("RowGroup has fixed membership")


# Content of the files this process last read or wrote, keyed by absolute path and valid while the file's
# (mtime_ns, size) is unchanged: the facade opens the same mapping file once per step (as the reference does
# with h5py), and re-parsing a container of (cells x k) tables each time would dominate a 20 ms GPU job.
_CONTENT_CACHE: Dict[str, tuple] = {}
_CONTENT_CACHE_MAX = 8          # files; least recently stored goes first


def _cache_put(key: str, value: tuple) -> None:
    _CONTENT_CACHE.pop(key, None)
    _CONTENT_CACHE[key] = value
    while len(_CONTENT_CACHE) > _CONTENT_CACHE_MAX:
        _CONTENT_CACHE.pop(next(iter(_CONTENT_CACHE)))


def _stat_key(fn: str):
    st = os.stat(fn)
    return (st.st_mtime_ns, st.st_size)


# Files whose writes are being batched (see `batched_writes`): abs path -> content tree waiting to be written
_DEFERRED: Dict[str, Optional[dict]] = {}


class batched_writes:
    """``with batched_writes(fn): ...`` - every flush / close of ``fn`` inside the block only records the
    content; the container is written once when the block ends.  The facade wraps its multi-step calls
    (``Mapping.map_target`` = stash name + distances + graph) in it: three rewrites of a growing file become one.
    Re-opening the file inside the block sees the pending content."""

    def __init__(self, fn: str):
        self.key = os.path.abspath(str(fn))
        self.fn = str(fn)
        self.outer = False

    def __enter__(self):
        self.outer = self.key in _DEFERRED
        if not self.outer:
            _DEFERRED[self.key] = None
        return self

    def __exit__(self, *exc):
        if self.outer:
            return False
        pending = _DEFERRED.pop(self.key, None)
        if pending is not None:
            f = File.__new__(File)
            Group.__init__(f, "/", None)
            f._root, f.filename, f.mode, f._dirty, f._open = f, self.fn, "a", True, True
            f._c = pending
            _reroot(f, f)
            f._persist()
            f._open = False
        return False


class File(Group):
    """``File(fn, mode)`` with h5py's mode letters: r, r+, a, w, w-/x."""

    def __init__(self, fn: str, mode: str = "r", **_ignored):
        super().__init__("/", None)
        self._root = self
        self.filename = str(fn)
        self.mode = mode
        self._dirty = False
        self._open = True
        exists = (self.filename in MEMORY_FILES) if _MEMORY_ONLY else \
            ((os.path.exists(self.filename) and os.path.getsize(self.filename) > 0) or
             _DEFERRED.get(os.path.abspath(self.filename)) is not None)
        if mode in ("r", "r+") and not exists:
            raise OSError("Unable to open file (unable to open file: name = %r)" % self.filename)
        if mode in ("w-", "x") and exists:
            raise OSError("Unable to create file (file exists)")
        if mode in ("r", "r+", "a") and exists:
            self._load()
        else:
            self._dirty = True
            if mode in ("w", "w-", "x", "a"):
                self._persist()        # the file exists on disk as soon as it is created

    # -- persistence
    def _load(self):
        pending = _DEFERRED.get(os.path.abspath(self.filename))
        if pending is not None:                     # inside batched_writes: the newest content is not on disk yet
            self._c = pending
            _reroot(self, self)
            self._dirty = False
            return
        if _MEMORY_ONLY:
            self._c = MEMORY_FILES[self.filename]._c
            _reroot(self, self)
            return
        key = os.path.abspath(self.filename)
        hit = _CONTENT_CACHE.get(key)
        if hit is not None and hit[0] == _stat_key(self.filename):
            self._c = hit[1]
            _reroot(self, self)
            self._dirty = False
            return
        with zipfile.ZipFile(self.filename, "r") as z:
            manifest = json.loads(z.read("__manifest__.json").decode("utf-8"))
            for path in manifest["groups"]:
                if path not in self:
                    Group.create_group(self, path)
            for path, key in manifest["datasets"].items():
                arr = np.load(io.BytesIO(z.read(key)), allow_pickle=False)
                Group.create_dataset(self, path, data=arr)
            for path, (nkey, dkey) in manifest["rowgroups"].items():
                names = np.load(io.BytesIO(z.read(nkey)), allow_pickle=False)
                data = np.load(io.BytesIO(z.read(dkey)), allow_pickle=False)
                Group.create_row_group(self, path, list(names), data)._names_enc = names
        _cache_put(key, (_stat_key(self.filename), self._c))
        self._dirty = False

    def _persist(self):
        key = os.path.abspath(self.filename)
        if key in _DEFERRED and not _MEMORY_ONLY:
            _DEFERRED[key] = self._c                # written once, when the batched_writes block ends
            if not os.path.exists(self.filename):
                open(self.filename, "ab").close()
            self._dirty = False
            return
        if _MEMORY_ONLY:
            holder = MEMORY_FILES.setdefault(self.filename, Group("/"))
            holder._c = self._c
            if not os.path.exists(self.filename):
                open(self.filename, "ab").close()
            self._dirty = False
            return
        manifest = {"groups": [], "datasets": {}, "rowgroups": {}}
        tmp = self.filename + ".tmp%d" % os.getpid()
        with zipfile.ZipFile(tmp, "w", zipfile.ZIP_STORED, allowZip64=True) as z:
            counter = [0]

            def put(arr) -> str:
                key = "a%d.npy" % counter[0]
                counter[0] += 1
                buf = io.BytesIO()
                np.save(buf, np.ascontiguousarray(arr), allow_pickle=False)
                z.writestr(key, buf.getvalue())
                return key

            def rec(g: Group, prefix: str):
                for k, v in g._c.items():
                    p = prefix + "/" + k if prefix else k
                    if isinstance(v, RowGroup):
                        manifest["rowgroups"][p] = [put(v.names_encoded()), put(v.data)]
                    elif isinstance(v, Group):
                        manifest["groups"].append(p)
                        rec(v, p)
                    else:
                        manifest["datasets"][p] = put(v._a)
            rec(self, "")
            z.writestr("__manifest__.json", json.dumps(manifest))
        os.replace(tmp, self.filename)
        _cache_put(os.path.abspath(self.filename), (_stat_key(self.filename), self._c))
        self._dirty = False

    def flush(self):
        if self._open and self.mode != "r" and self._dirty:
            self._persist()

    def close(self):
        if self._open:
            self.flush()
            self._open = False

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _reroot(g: Group, root: Group):
    for v in g._c.values():
        if isinstance(v, Group):
            v._root = root
            _reroot(v, root)


# ----------------------------------------------------------------------------- real HDF5 where h5py exists
ROWGROUP_ATTR = "nabo_b200_layout"          # attribute marking an HDF5 group that stores a RowGroup
ROWGROUP_NAMES, ROWGROUP_ROWS = "names", "rows"


def _h5py():
    """The h5py module, or None (absent from the target image).  NABO_B200_STORE=container forces the zip/npy
    container even where h5py is importable."""
    if os.environ.get("NABO_B200_STORE", "") == "container":
        return None
    try:
        import h5py
        return h5py
    except Exception:
        return None


class H5Group:
    """The same call surface as ``Group`` on top of an ``h5py`` group: files are real HDF5 with the reference's
    group / dataset names (nabo/_mapping.py:252-273, 340-355; nabo/_dataset.py:1028).  A ``RowGroup`` is stored as
    an HDF5 group with the attribute ``nabo_b200_layout = 'rowgroup'`` holding ``rows`` (one 2-D dataset) and
    ``names`` (byte strings); per-node graph groups written with ``graph_layout='reference'`` are plain HDF5
    datasets that upstream nabo opens directly."""

    def __init__(self, node, h5):
        self._g, self._h5 = node, h5
        self.name = node.name

    def _wrap(self, obj):
        if isinstance(obj, self._h5.Group):
            kind = obj.attrs.get(ROWGROUP_ATTR, "")
            if isinstance(kind, bytes):
                kind = kind.decode("utf-8")
            if kind == "rowgroup":
                return RowGroup(obj.name, None, list(obj[ROWGROUP_NAMES][:]), obj[ROWGROUP_ROWS][:])
            return H5Group(obj, self._h5)
        return obj                                      # h5py.Dataset: slicing, fields, shape, dtype as upstream

    def __contains__(self, path) -> bool:
        return isinstance(path, str) and path in self._g

    def __getitem__(self, path: str):
        return self._wrap(self._g[path])

    def __delitem__(self, path: str):
        del self._g[path]

    def __iter__(self) -> Iterator[str]:
        return iter(sorted(self._g.keys()))

    def __len__(self):
        return len(self._g)

    def keys(self):
        return sorted(self._g.keys())

    def values(self):
        return [self[k] for k in self.keys()]

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def create_group(self, path: str) -> "H5Group":
        return H5Group(self._g.create_group(path), self._h5)

    def require_group(self, path: str) -> "H5Group":
        return self[path] if path in self else self.create_group(path)

    def create_dataset(self, path: str, shape=None, dtype=None, data=None, **kw):
        if data is not None:
            data = np.asarray(data) if dtype is None else np.asarray(data, dtype=dtype)
            if data.dtype.kind == "U":                  # h5py has no fixed-width unicode: bytes, as on this path upstream
                data = np.char.encode(data, "utf-8")
            return self._g.create_dataset(path, data=data, **kw)
        return self._g.create_dataset(path, shape=shape, dtype=dtype if dtype is not None else np.float32, **kw)

    def create_row_group(self, path: str, names: List[str], data: np.ndarray) -> "RowGroup":
        g = self._g.create_group(path)
        g.attrs[ROWGROUP_ATTR] = "rowgroup"
        data = np.asarray(data)
        enc = np.char.encode(np.asarray([n.decode("utf-8") if isinstance(n, bytes) else str(n) for n in names], dtype=str),
                             "utf-8") if len(names) else np.zeros(0, "S1")
        g.create_dataset(ROWGROUP_NAMES, data=enc)
        g.create_dataset(ROWGROUP_ROWS, data=data)
        return RowGroup(g.name, None, list(names), data)

    def __repr__(self):
        return "<store.H5Group %r (%d members)>" % (self.name, len(self))


class H5File(H5Group):
    """``h5py.File`` behind the ``File`` call surface (mode letters r, r+, a, w, w-/x)."""

    def __init__(self, fn: str, mode: str, h5):
        self._f = h5.File(str(fn), mode)
        super().__init__(self._f, h5)
        self.filename, self.mode = str(fn), mode

    def flush(self):
        if self.mode != "r":
            self._f.flush()

    def close(self):
        try:
            self._f.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def _is_container(fn: str) -> bool:
    try:
        return os.path.getsize(fn) > 0 and zipfile.is_zipfile(fn)
    except OSError:
        return False


def open_file(fn: str, mode: str = "r", **kw):
    """Single entry point the facade uses to open dataset / PCA / mapping files ("h5py when
    importable"): real HDF5 through ``H5File`` where h5py exists, otherwise the zip/npy container ``File`` with
    the same group / dataset names.  A file that already is a container stays one (it is never re-read as HDF5)."""
    h5 = None if (_MEMORY_ONLY or str(fn) in MEMORY_FILES) else _h5py()
    if h5 is not None and not _is_container(str(fn)) and _DEFERRED.get(os.path.abspath(str(fn))) is None:
        return H5File(fn, mode, h5)
    return File(fn, mode=mode)
