import numpy as np


def concatenate(meshes, *more):
    from .base import Trimesh
    if not isinstance(meshes, (list, tuple)):
        meshes = [meshes]
    meshes = list(meshes) + list(more)
    V, F, off = [], [], 0
    for m in meshes:
        V.append(m.vertices)
        F.append(m.faces + off)
        off += len(m.vertices)
    return Trimesh(np.vstack(V), np.vstack(F))
