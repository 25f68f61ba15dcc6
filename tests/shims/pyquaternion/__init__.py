"""Minimal stand-in for pyquaternion 0.9.9: construction from (w, x, y, z) and Quaternion.distance."""
import numpy as np


class Quaternion:
    def __init__(self, *args, **kwargs):
        if len(args) == 1:
            a = args[0]
            self.q = np.array(a.q if isinstance(a, Quaternion) else a, dtype=np.float64).reshape(4)
        elif len(args) == 4:
            self.q = np.array(args, dtype=np.float64)
        elif "array" in kwargs:
            self.q = np.array(kwargs["array"], dtype=np.float64).reshape(4)
        else:
            self.q = np.array([1.0, 0.0, 0.0, 0.0])

    @property
    def norm(self):
        return float(np.sqrt(np.dot(self.q, self.q)))

    @property
    def inverse(self):
        w, x, y, z = self.q
        return Quaternion(np.array([w, -x, -y, -z]) / np.dot(self.q, self.q))

    def __mul__(self, o):
        a, b = self.q, o.q
        return Quaternion(np.array([
            a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
            a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
            a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
            a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]]))

    @classmethod
    def log(cls, q):
        v = q.q[1:]
        vn = float(np.linalg.norm(v))
        n = q.norm
        tol = 1e-14
        if vn < tol:
            if q.q[0] < 0:
                return Quaternion(np.array([np.log(n), np.pi, 0.0, 0.0]))
            return Quaternion(np.array([np.log(n), 0.0, 0.0, 0.0]))
        return Quaternion(np.concatenate([[np.log(n)], np.arccos(np.clip(q.q[0] / n, -1.0, 1.0)) * v / vn]))

    @classmethod
    def log_map(cls, q, p):
        return cls.log(q.inverse * p)

    @classmethod
    def distance(cls, q0, q1):
        """Norm of the log map of q0^-1 q1 (pyquaternion.Quaternion.distance)."""
        return cls.log_map(q0, q1).norm
