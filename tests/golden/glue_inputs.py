"""Seeded inputs shared by the fixture generator (make_golden_glue.py) and the tests that replay them."""
import numpy as np

from multimesh_b200 import meshgen

NAMES = ["QKAPPA", "QMU", "RHO", "VP", "VS"]
# outer-core shell (fluid, lowest id) + mantle + two crustal layers: what `layers="nocore"` needs (utils.py:426-436)
CORE_LAYERS = [(3000e3, 3480e3, 1, 0, 1), (3480e3, 6291e3, 3, 3, 0), (6291e3, 6346e3, 1, 2, 0),
               (6346e3, 6371e3, 1, 1, 0)]


def shell_pair(order, seed, layers=None):
    """Cubed-sphere shell pair (BASELINE config 3, scaled down): source n_lat = 4, target n_lat = 3.
    Returns {"from"/"to": (coords [E,P,3], data [E,6,P], element_data [E,2])}; fields = ISO + z_node_1D,
    element_data = (fluid, layer); the target's ISO fields are random (what must be overwritten / kept)."""
    layers = layers or meshgen.default_shell_layers()
    names = NAMES + ["z_node_1D"]
    rng = np.random.default_rng(seed)
    out = {}
    for key, n_lat in (("from", 4), ("to", 3)):
        coords, el, z1d = meshgen.shell_mesh(n_lat, layers, order)
        data = meshgen.analytic_fields(coords, names)
        data[:, 5, :] = z1d
        if key == "to":
            data[:, :5, :] = rng.uniform(1.0, 2.0, data[:, :5, :].shape)
        out[key] = (coords, data, np.stack([el["fluid"], el["layer"]], axis=1))
    return out
