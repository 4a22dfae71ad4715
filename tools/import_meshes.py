"""Convert the reference's pickled 67P meshes into the bundled .npz fixtures.

Run ONCE in the build container (reads /root/reference/3dmeshes/*.pk, which does
not exist on the GPU box).  The output keeps the raw, un-scaled vertex
coordinates exactly as pickled (float64) and the triangle indices (int32); the
scaling of `read_pk_file` / `create_mesh` (reference toss/mesh/mesh_utility.py:11-34,60)
is applied at load time by toss_b200.mesh.mesh_utility.

Data provenance: rasmusmarak/TOSS `3dmeshes/` (GPL-3.0), derived from the ESA
Rosetta shape model of 67P/Churyumov-Gerasimenko.  See NOTICE.
"""
import pickle
import sys
from pathlib import Path

import numpy as np

SRC = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/3dmeshes")
DST = Path(__file__).resolve().parent.parent / "toss_b200" / "resources" / "meshes"
NAMES = {
    "churyumov-gerasimenko_lllp": "67p_lllp",
    "churyumov-gerasimenko_llp": "67p_llp",
    "churyumov-gerasimenko_lp": "67p_lp",
    "churyumov-gerasimenko": "67p_full",
}

def main():
    DST.mkdir(parents=True, exist_ok=True)
    for src, dst in NAMES.items():
        with open(SRC / f"{src}.pk", "rb") as f:
            points, triangles = pickle.load(f)
        points = np.asarray(points, dtype=np.float64)
        triangles = np.asarray(triangles, dtype=np.int32)
        assert points.ndim == 2 and points.shape[1] == 3
        assert triangles.ndim == 2 and triangles.shape[1] == 3
        np.savez_compressed(DST / f"{dst}.npz", points=points, triangles=triangles)
        print(dst, points.shape, triangles.shape)

if __name__ == "__main__":
    main()
