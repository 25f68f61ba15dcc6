"""CollisionManager for CONVEX objects (the planners only register convex hulls, run_joint_plan.py:232-235)."""
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)


class CollisionManager:
    def __init__(self):
        self._objs = {}

    def add_object(self, name, mesh, transform=None):
        V = np.asarray(mesh.vertices, dtype=np.float64)
        if transform is not None:
            T = np.asarray(transform, dtype=np.float64)
            V = V @ T[:3, :3].T + T[:3, 3]
        self._objs[name] = V

    def remove_object(self, name):
        del self._objs[name]

    def in_collision_single(self, mesh, transform=None, return_names=False, return_data=False):
        """True iff the SURFACE of `mesh` (a convex hull) intersects the surface of a registered hull: what FCL's
        triangle-mesh query reports (one hull strictly inside the other is not a collision)."""
        from assemble_them_all_b200.planner import HullPair
        V = np.asarray(mesh.vertices, dtype=np.float64)
        T = np.eye(4) if transform is None else np.asarray(transform, dtype=np.float64)
        hit, names = False, set()
        for name, W in self._objs.items():
            if bool(HullPair(V, W).surfaces_intersect(T)[0][0]):
                hit = True
                names.add(name)
        return (hit, names) if return_names else hit
