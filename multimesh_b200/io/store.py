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


class NpzStore:
    def __init__(self, path, mode="r"):
        self.path, self.mode = str(path), mode
        self._dirty = False
        self._arrays = {}
        if os.path.exists(self.path):
            with np.load(self.path, allow_pickle=False) as z:
                self._arrays = {k: z[k] for k in z.files}
        elif mode == "r":
            raise FileNotFoundError(self.path)

    def __contains__(self, name):
        return name in self._arrays

    def read(self, name):
        return self._arrays[name]

    def shape(self, name):
        return self._arrays[name].shape

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
        if self._dirty and self.mode != "r":
            np.savez(self.path if self.path.endswith(".npz") else self.path + ".npz", **self._arrays)
            self._dirty = False

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
