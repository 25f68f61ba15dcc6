"""Minimal stand-in for trimesh 3.10.8: only what the reference's planners call (see tests/shims/README.md)."""
import numpy as np

from . import collision, convex, util  # noqa: F401
from .base import Scene, Trimesh  # noqa: F401


def load_mesh(path, process=False, maintain_order=True, **kwargs):
    """Wavefront OBJ -> Trimesh (vertex order and face order of the file are kept)."""
    V, F = [], []
    with open(path) as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            if t[0] == "v":
                V.append([float(t[1]), float(t[2]), float(t[3])])
            elif t[0] == "f":
                idx = [int(s.split("/")[0]) for s in t[1:]]
                idx = [i - 1 if i > 0 else len(V) + i for i in idx]
                for k in range(1, len(idx) - 1):   # fan triangulation of polygons
                    F.append([idx[0], idx[k], idx[k + 1]])
    return Trimesh(vertices=np.array(V, dtype=np.float64).reshape(-1, 3), faces=np.array(F, dtype=np.int64).reshape(-1, 3))


load = load_mesh
