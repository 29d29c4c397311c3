"""
SalvusMesh -- reader/writer of the fields of a Salvus mesh the interpolation path needs.
Same class, attributes and method names as multi_mesh/components/salvus_mesh_reader.py:7-178;
additionally constructible from arrays (`from_arrays`) because neither Salvus files nor h5py
exist in this environment.
"""
import pathlib
from typing import Union

import numpy as np

from ..gll import order_from_npoints
from ..io.store import open_store


class SalvusMesh(object):
    def __init__(self, filename: Union[str, pathlib.Path], fast_mode: bool = True):
        self.filename = filename
        with open_store(filename, "r") as st:
            self.points = np.asarray(st.read("MODEL/coordinates"), dtype=np.float64)
            self._data = np.asarray(st.read("MODEL/data"), dtype=np.float64)
            self.nodal_parameter_indices = st.labels("MODEL/data")
            if "MODEL/element_data" in st:
                self._element_data = np.asarray(st.read("MODEL/element_data"))
                self.elemental_parameter_indices = st.labels("MODEL/element_data")
            else:
                self._element_data = np.zeros((self.points.shape[0], 0))
                self.elemental_parameter_indices = []
            self.global_strings = st.attrs("MODEL")
        self._finish(fast_mode)

    @classmethod
    def from_arrays(cls, points, data, nodal_names, element_data=None, elemental_names=(),
                    global_strings=None, fast_mode=False):
        self = cls.__new__(cls)
        self.filename = None
        self.points = np.ascontiguousarray(points, dtype=np.float64)
        self._data = np.ascontiguousarray(data, dtype=np.float64)
        self.nodal_parameter_indices = list(nodal_names)
        self._element_data = (np.zeros((self.points.shape[0], 0)) if element_data is None
                              else np.asarray(element_data, dtype=np.float64))
        self.elemental_parameter_indices = list(elemental_names)
        self.global_strings = dict(global_strings or {})
        self._finish(fast_mode)
        return self

    def _finish(self, fast_mode):
        self.nelem = self.get_nelem()
        self.n_gll_points = self.get_n_gll_points()
        self.dimensions = self.get_dimensions()
        self.shape_order = self.get_shape_order()
        if not fast_mode:
            self.elemental_fields = self.get_elemental_fields()
            self.element_nodal_fields = self.get_element_nodal_fields()

    # -- getters (names as in the reference) ----------------------------------------------------
    def get_points(self):
        return self.points

    def get_n_gll_points(self):
        return self.points.shape[1]

    def get_dimensions(self):
        return self.points.shape[2]

    def get_shape_order(self):
        return order_from_npoints(self.n_gll_points, self.dimensions)

    def get_nelem(self):
        return self.points.shape[0]

    def get_global_strings(self):
        return self.global_strings

    def get_nodal_parameter_indices(self):
        return self.nodal_parameter_indices

    def get_elemental_parameter_indices(self):
        return self.elemental_parameter_indices

    def get_elemental_fields(self):
        if not hasattr(self, "elemental_fields"):
            self.elemental_fields = {p: self._element_data[:, i].copy()
                                     for i, p in enumerate(self.elemental_parameter_indices)}
        return self.elemental_fields

    def get_element_nodal_fields(self):
        if not hasattr(self, "element_nodal_fields"):
            self.element_nodal_fields = {p: self._data[:, i, :].copy()
                                         for i, p in enumerate(self.nodal_parameter_indices)}
        return self.element_nodal_fields

    def get_element_centroids(self):
        # sequential mean over the nodes; the CUDA K0 kernel is bit-equal to this
        return np.mean(self.points, axis=1)

    def get_element_nodes(self):
        return self.points

    def get_element_nodal_field(self, param):
        return self._data[:, self.nodal_parameter_indices.index(param), :]

    def get_elemental_field(self, param):
        return self._element_data[:, self.elemental_parameter_indices.index(param)]

    def set_global_string(self, name: str, value: str):
        assert isinstance(value, str), "Value needs to be a string"
        assert isinstance(name, str), "Name needs to be a string"
        self.global_strings[name] = value
        if self.filename is not None:
            with open_store(self.filename, "r+") as st:
                st.set_attr("MODEL", name, value)

    def attach_field(self, name: str, data: np.ndarray):
        """Attach an elemental field [nelem] or an element-nodal field [nelem, n_gll_points];
        only existing fields can be (re)attached, as in the reference (:165-178)."""
        assert isinstance(data, np.ndarray), "Data needs to be a numpy array"
        if data.shape == (self.nelem, self.n_gll_points):
            if name not in self.nodal_parameter_indices:
                raise ValueError("Currently we only attach existing fields")
            self._data[:, self.nodal_parameter_indices.index(name), :] = data
            if hasattr(self, "element_nodal_fields"):
                self.element_nodal_fields[name] = np.array(data)
            target = ("MODEL/data", self._data)
        elif data.shape == (self.nelem,):
            if name not in self.elemental_parameter_indices:
                raise ValueError("Currently we only attach existing fields")
            self._element_data[:, self.elemental_parameter_indices.index(name)] = data
            if hasattr(self, "elemental_fields"):
                self.elemental_fields[name] = np.array(data)
            target = ("MODEL/element_data", self._element_data)
        else:
            raise ValueError("We can only attach elemental_nodal_field or elemental_fields")
        if self.filename is not None:
            with open_store(self.filename, "r+") as st:
                st.write(*target)
        print(f"Attached field {name} to mesh")
