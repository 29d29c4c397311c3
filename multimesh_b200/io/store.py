"""
Model-file access used by the readers and drivers.

Two backends with one small interface:
  * HDF5 (Salvus mesh files, `MODEL/coordinates`, `MODEL/data`, `MODEL/element_data`, the
    `DIMENSION_LABELS` attribute; salvus_mesh_reader.py:38-79, utils.py:137-168,206-217) through
    h5py, imported lazily -- h5py is not installed in the build/test image;
  * NPZ (`*.npz`): the same dataset names in a numpy archive, so that the file-based drivers can be
    exercised end to end without h5py.  Labels are stored as `<dataset>@labels`.
"""
import json
import os

import numpy as np


def _parse_labels(raw):
    """'[ A | B | C ]' -> ['A', 'B', 'C'] (salvus_mesh_reader.py:67-72)."""
    if isinstance(raw, bytes):
        raw = raw.decode()
    return [s for s in raw.replace(" ", "").strip("[]").split("|") if s]


def _format_labels(names):
    return "[ " + " | ".join(names) + " ]"  # utils.py:165


def _pinned_empty(shape, dtype):
    """A pinned (page-locked) host tensor and its numpy view: the staging buffer of an asynchronous H2D copy."""
    import torch

    t = torch.empty(tuple(int(v) for v in shape), dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
    return t, t.numpy()


class _LazyArrays(dict):
    """name -> array of an .npz file, each member read on first access (np.load's NpzFile is lazy per key)."""

    def __init__(self, path):
        super().__init__()
        self._z = np.load(path, allow_pickle=False)
        self._names = set(self._z.files)

    def __contains__(self, k):
        return dict.__contains__(self, k) or k in self._names

    def __missing__(self, k):
        if k not in self._names:
            raise KeyError(k)
        v = self._z[k]
        self[k] = v
        return v

    def load_all(self):
        for k in sorted(self._names):
            self[k]
        return dict(self)

    def close(self):
        self._z.close()


class NpzStore:
    def __init__(self, path, mode="r"):
        self.path, self.mode = str(path), mode
        self._dirty = False
        self._arrays = {}
        if os.path.exists(self.path):
            self._arrays = _LazyArrays(self.path)
        elif mode == "r":
            raise FileNotFoundError(self.path)

    def __contains__(self, name):
        return name in self._arrays

    def read(self, name):
        return self._arrays[name]

    def _member_header(self, name):
        """(zip member, shape, dtype, fortran_order) of an array without reading its data."""
        import zipfile

        zf = zipfile.ZipFile(self.path)
        fh = zf.open(name + ".npy")
        major, minor = np.lib.format.read_magic(fh)
        shape, fortran, dtype = (np.lib.format.read_array_header_1_0(fh) if major == 1
                                 else np.lib.format.read_array_header_2_0(fh))
        return zf, fh, shape, dtype, fortran

    def shape(self, name):
        if isinstance(self._arrays, _LazyArrays) and not dict.__contains__(self._arrays, name):
            zf, fh, shape, _, _ = self._member_header(name)
            fh.close()
            zf.close()
            return tuple(shape)
        return self._arrays[name].shape

    def read_pinned(self, name):
        """The array as a PINNED host tensor, read from the file straight into the page-locked buffer (no
        intermediate pageable copy), ready for an asynchronous H2D copy (io/staging.py)."""
        zf, fh, shape, dtype, fortran = self._member_header(name)
        try:
            if fortran or dtype.hasobject:
                raise ValueError(f"{name}: only C-ordered plain arrays can be staged")
            t, view = _pinned_empty(shape, dtype)
            buf = memoryview(view.reshape(-1).view(np.uint8))
            got = 0
            while got < len(buf):
                n = fh.readinto(buf[got:got + (64 << 20)])
                if not n:
                    raise IOError(f"{self.path}:{name}: short read")
                got += n
            return t
        finally:
            fh.close()
            zf.close()

    def write(self, name, array):
        assert self.mode != "r", "store opened read-only"
        self._arrays[name] = np.ascontiguousarray(array)
        self._dirty = True

    def labels(self, name):
        return _parse_labels(str(np.asarray(self._arrays[name + "@labels"]).reshape(-1)[0]))

    def set_labels(self, name, names):
        self.write(name + "@labels", np.array(_format_labels(names)))

    def attrs(self, group="MODEL"):
        key = group + "@attrs"
        return json.loads(str(np.asarray(self._arrays[key]).reshape(-1)[0])) if key in self._arrays else {}

    def set_attr(self, group, name, value):
        a = self.attrs(group)
        a[name] = value
        self.write(group + "@attrs", np.array(json.dumps(a)))

    def close(self):
        lazy = isinstance(self._arrays, _LazyArrays)
        if self._dirty and self.mode != "r":
            arrays = self._arrays.load_all() if lazy else dict(self._arrays)
            if lazy:
                self._arrays.close()
                self._arrays = arrays
            np.savez(self.path if self.path.endswith(".npz") else self.path + ".npz", **arrays)
            self._dirty = False
        elif lazy:
            self._arrays.close()
            self._arrays = dict(self._arrays)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class H5Store:
    def __init__(self, path, mode="r"):
        import h5py  # lazy: absent in the test image

        self.f = h5py.File(path, mode)

    def __contains__(self, name):
        return name in self.f

    def read(self, name):
        return self.f[name][()]

    def shape(self, name):
        return self.f[name].shape

    def read_pinned(self, name):
        """The dataset as a PINNED host tensor: HDF5 reads straight into the page-locked buffer (read_direct)."""
        ds = self.f[name]
        t, view = _pinned_empty(ds.shape, ds.dtype)
        ds.read_direct(view)
        return t

    def write(self, name, array):
        if name in self.f:
            if self.f[name].shape == array.shape:
                self.f[name][...] = array
                return
            del self.f[name]
        self.f.create_dataset(name, data=array)

    def labels(self, name):
        return _parse_labels(self.f[name].attrs.get("DIMENSION_LABELS")[1])

    def set_labels(self, name, names):
        ds = self.f[name]
        ds.dims[0].label = "element"
        ds.dims[1].label = _format_labels(names)
        if len(ds.shape) > 2:
            ds.dims[2].label = "point"

    def attrs(self, group="MODEL"):
        out = {}
        for k, v in self.f[group].attrs.items():
            if isinstance(v, (bytes, np.bytes_)):
                out[k] = v.decode()
        return out

    def set_attr(self, group, name, value):
        self.f[group].attrs[name] = np.bytes_(value)

    def close(self):
        self.f.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def open_store(path, mode="r"):
    p = str(path)
    if p.endswith(".npz"):
        return NpzStore(p, mode)
    return H5Store(p, mode)


def write_gll_model(path, coordinates, data, names, element_data=None, element_names=None,
                    global_strings=None):
    """Create a model file in the Salvus layout (used by tests and examples)."""
    with open_store(path, "w") as st:
        st.write("MODEL/coordinates", np.asarray(coordinates, dtype=np.float64))
        st.write("MODEL/data", np.asarray(data, dtype=np.float64))
        st.set_labels("MODEL/data", list(names))
        if element_data is not None:
            st.write("MODEL/element_data", np.asarray(element_data, dtype=np.float64))
            st.set_labels("MODEL/element_data", list(element_names))
        for k, v in (global_strings or {}).items():
            st.set_attr("MODEL", k, str(v))
