import numpy as np
from scipy.spatial import ConvexHull


def convex_hull(obj, qhull_options=None):
    """Convex hull of a point cloud (or of a mesh's vertices) as a Trimesh whose vertices are the hull vertices."""
    from .base import Trimesh
    pts = np.asarray(getattr(obj, "vertices", obj), dtype=np.float64).reshape(-1, 3)
    h = ConvexHull(pts)
    remap = -np.ones(len(pts), dtype=np.int64)
    remap[h.vertices] = np.arange(len(h.vertices))
    faces = remap[h.simplices]
    V = pts[h.vertices]
    # outward orientation
    c = V.mean(axis=0)
    n = np.cross(V[faces[:, 1]] - V[faces[:, 0]], V[faces[:, 2]] - V[faces[:, 0]])
    flip = np.einsum("ij,ij->i", n, V[faces[:, 0]] - c) < 0
    faces[flip] = faces[flip][:, ::-1]
    return Trimesh(V, faces)
