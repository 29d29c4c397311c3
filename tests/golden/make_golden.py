"""
Generates the golden fixtures of tests/golden/ from the reference itself.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
  * trilinear_ref.npz : inputs and outputs of the reference's own C routines
                        (multi_mesh/src/centroid.c, trilinearinterpolator.c), compiled by
                        oracle/build.py into oracle/_ref, on a seeded warped HEX8 mesh including
                        points outside the mesh (failure branch).
  * hex8_weights.json : known-answer vectors recorded from the compiled reference at survey time
                         -- re-derived here from the same routine through one-element
                        meshes and checked against the recorded values.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from multimesh_b200 import meshgen  # noqa: E402
from oracle import capi  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    assert capi.ref_lib() is not None, "needs the compiled reference (oracle/_ref)"
    rng = np.random.default_rng(20261018)
    points, conn = meshgen.hex8_mesh((5, 4, 6), warp=0.04)
    connC = np.ascontiguousarray(conn[:, np.argsort([0, 3, 2, 1, 4, 5, 6, 7])])
    cent = capi.ref_centroid(conn, points)
    q = np.concatenate([rng.uniform(-0.06, 1.06, (600, 3)), points[::7]])
    nn = capi.knn_bruteforce(cent, q, 20).astype(np.int64)
    nf, enc, w = capi.ref_trilinear_interpolator(20, nn, connC, points, q)
    np.savez_compressed(os.path.join(HERE, "trilinear_ref.npz"), k=20, nearest=nn, connectivity=connC,
                        connectivity_exodus=conn, nodes=points, points=q, nfailed=nf, enclosing=enc,
                        weights=w, centroids=cent)
    # unit cube, C vertex order: weights of the reference at two survey-time points
    R = [-1, -1, +1, +1, -1, +1, +1, -1]
    S = [-1, +1, +1, -1, -1, -1, +1, +1]
    T = [-1, -1, -1, -1, +1, +1, +1, +1]
    cube = np.array([[(r + 1) / 2, (s + 1) / 2, (t + 1) / 2] for r, s, t in zip(R, S, T)])
    recorded = {
        (0.25, 0.5, 0.75): [0.09375, 0.09375, 0.03125, 0.03125, 0.28125, 0.09375, 0.09375, 0.28125],
        (0.9, 0.9, 0.1): [0.009, 0.081, 0.729, 0.081, 0.001, 0.009, 0.081, 0.009],
    }
    cases = []
    connI = np.arange(8, dtype=np.int64)[None, :]
    for p, wrec in recorded.items():
        nf1, _, w1 = capi.ref_trilinear_interpolator(1, np.zeros((1, 1), dtype=np.int64), connI, cube,
                                                      np.array([p]))
        assert nf1 == 0
        local = [2 * c - 1 for c in p]
        cases.append({"point_in_unit_cube": list(p), "local": local, "weights": w1[0].tolist(),
                      "recorded_at_survey": wrec})
    json.dump({"source": "compiled reference trilinearinterpolator.c:interpolateAtPoint via triLinearInterpolator",
               "cases": cases}, open(os.path.join(HERE, "hex8_weights.json"), "w"), indent=1)
    print("wrote fixtures:", nf, "failed points in trilinear_ref.npz")
    for c in cases:
        print(c["point_in_unit_cube"], np.max(np.abs(np.array(c["weights"]) - np.array(c["recorded_at_survey"]))))


if __name__ == "__main__":
    main()
