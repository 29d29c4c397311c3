"""
Exodus -- nodal HEX8 mesh container; same attributes/methods as multi_mesh/io/exodus.py:9-142.
File access needs pyexodus (imported lazily, absent here); `from_arrays` builds one in memory.
`get_element_centroid` runs the CUDA twin of centroid.c through the library's legacy `centroid`
symbol (host pointers), which is what the reference's `lib.centroid` call (:56-63) binds to.
"""
import ctypes as C

import numpy as np


class Exodus(object):
    def __init__(self, filename, mode="r"):
        self._filename = filename
        assert mode in ["a", "r"], "Only mode 'a', 'r' is supported"
        self.mode = mode
        self._nodal = {}
        self._elemental = {}
        self._read()

    @classmethod
    def from_arrays(cls, points, connectivity, nodal_fields=None, mode="a"):
        self = cls.__new__(cls)
        self._filename = None
        self.mode = mode
        self.points = np.ascontiguousarray(points, dtype=np.float64)
        self.connectivity = np.ascontiguousarray(connectivity, dtype=np.int64)
        self.ndim = self.points.shape[1]
        self.nelem, self.nodes_per_element = self.connectivity.shape
        self._nodal = {k: np.asarray(v, dtype=np.float64) for k, v in (nodal_fields or {}).items()}
        self._elemental = {}
        self.nodal_parameters = list(self._nodal)
        self.elem_var_names = []
        return self

    def _read(self):
        from pyexodus import exodus  # lazy

        with exodus(self._filename, self.mode) as e:
            self.ndim = e.num_dims
            conn, self.nelem, self.nodes_per_element = e.get_elem_connectivity(id=1)
            self.connectivity = np.array(conn, dtype="int64") - 1  # exodus is 1-based
            self.elem_var_names = e.get_element_variable_names()
            self.points = np.array((e.get_coords())).T.astype(np.float64)
            self.nodal_parameters = e.get_node_variable_names()

    def get_element_centroid(self):
        from .. import _lib

        lib = _lib.load_lib()
        centroid = np.zeros((self.nelem, self.ndim))
        pts = np.ascontiguousarray(self.points)
        conn = np.ascontiguousarray(self.connectivity, dtype=np.int64)
        lib.centroid(self.ndim, self.nelem, self.nodes_per_element, conn.ctypes.data_as(C.c_void_p),
                     pts.ctypes.data_as(C.c_void_p), centroid.ctypes.data_as(C.c_void_p))
        return centroid

    def attach_field(self, name, values):
        assert self.mode in ["a"], "Attach field option only available in mode 'a'"
        values = np.asarray(values)
        if self._filename is None:
            if values.size == self.nelem and values.size != self.npoint:
                self._elemental[name] = values.copy()
            elif values.size == self.npoint:
                self._nodal[name] = values.copy()
                if name not in self.nodal_parameters:
                    self.nodal_parameters.append(name)
            else:
                raise ValueError("Shape matches neither the nodes nor the elements")
            return
        from pyexodus import exodus

        with exodus(self._filename, self.mode) as e:
            if values.size == self.nelem:
                e.put_element_variable_values(blockId=1, name=name, step=1, values=values)
            elif values.size == self.npoint:
                idx = e.get_node_variable_names().index(name) + 1
                e.put_node_variable_name(name, index=idx)
                e.put_node_variable_values(name, 1, values)
            else:
                raise ValueError("Shape matches neither the nodes nor the elements")

    def get_element_field(self, name):
        if self._filename is None:
            return self._elemental[name]
        from pyexodus import exodus

        assert name in self.elem_var_names, "Could not find the requested field"
        with exodus(self._filename, self.mode) as e:
            return e.get_element_variable_values(blockId=1, name=name, step=1)

    def get_nodal_field(self, name):
        if self._filename is None:
            assert name in self._nodal, "Could not find the requested field"
            return self._nodal[name]
        from pyexodus import exodus

        with exodus(self._filename, self.mode) as e:
            assert name in e.get_node_variable_names(), "Could not find the requested field"
            return e.get_node_variable_values(name=name, step=1)

    @property
    def npoint(self):
        return self.points.shape[0]
