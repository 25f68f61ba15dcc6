import numpy as np


class _Visual:
    def __init__(self):
        self.face_colors = None
        self.vertex_colors = None


class Trimesh:
    def __init__(self, vertices=None, faces=None, **kwargs):
        self.vertices = np.array(vertices, dtype=np.float64).reshape(-1, 3)
        self.faces = np.array(faces if faces is not None else np.zeros((0, 3)), dtype=np.int64).reshape(-1, 3)
        self.visual = _Visual()

    def copy(self):
        m = Trimesh(self.vertices.copy(), self.faces.copy())
        m.visual.face_colors = self.visual.face_colors
        return m

    def apply_transform(self, matrix):
        M = np.asarray(matrix, dtype=np.float64)
        self.vertices = self.vertices @ M[:3, :3].T + M[:3, 3]
        return self

    @property
    def bounds(self):
        return np.array([self.vertices.min(axis=0), self.vertices.max(axis=0)])

    def export(self, path, **kwargs):
        with open(path, "w") as f:
            for v in self.vertices:
                f.write("v %.9f %.9f %.9f\n" % tuple(v))
            for t in self.faces:
                f.write("f %d %d %d\n" % tuple(t + 1))


class Scene:
    def __init__(self, geometry=None):
        if geometry is None:
            self.geometry = {}
        elif isinstance(geometry, dict):
            self.geometry = dict(geometry)
        else:
            self.geometry = {f"geometry_{i}": g for i, g in enumerate(geometry)}

    def show(self, **kwargs):
        raise RuntimeError("tests/shims/trimesh: no viewer")
