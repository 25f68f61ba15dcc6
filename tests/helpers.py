"""Shared test helpers: golden fixtures (outputs of the unmodified reference,
see oracle/make_golden.py) and tolerance checks."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# north_star: fp64 values agree to 1e-9 relative. Jacobian blocks contain
# entries that cancel to ~0, so "relative" is taken against the largest
# magnitude of the array being compared.
RTOL = 1e-9


def rel_err(ref, got):
    ref, got = np.asarray(ref, dtype=np.float64), np.asarray(got, dtype=np.float64)
    scale = np.max(np.abs(ref)) if ref.size else 0.0
    if scale == 0.0:
        return float(np.max(np.abs(got))) if got.size else 0.0
    return float(np.max(np.abs(ref - got)) / scale)


def assert_close(ref, got, what, rtol=RTOL):
    e = rel_err(ref, got)
    assert e <= rtol, f"{what}: relative error {e:.3e} > {rtol:.1e}"


Q_RTOL = 1e-9      # generalized coordinates (north_star)
QDOT_RTOL = 1e-6   # see assert_state_close


def assert_state_close(q_ref, qd_ref, q, qd, h, what):
    """q within 1e-9 relative. qdot is the derived quantity (q1 - q0) / h (Simulation.cpp:1259):
    a 1e-9 relative perturbation of q moves it by 1e-9 |q| / h, and the Newton tolerance
    (|g| < 1e-8, Simulation.cpp:1008) is amplified by 1/h as well, so qdot is held to
    1e-6 relative plus that absolute floor."""
    assert rel_err(q_ref, q) <= Q_RTOL, f"{what}: q relative error {rel_err(q_ref, q):.3e}"
    q_ref, qd_ref = np.asarray(q_ref), np.asarray(qd_ref)
    tol = QDOT_RTOL * float(np.max(np.abs(qd_ref))) + Q_RTOL * float(np.max(np.abs(q_ref))) / h
    err = float(np.max(np.abs(qd_ref - np.asarray(qd))))
    assert err <= tol, f"{what}: qdot error {err:.3e} > {tol:.3e}"


class GoldenContact:
    def __init__(self, z):
        self.z = z
        self.names = [str(n) for n in z["names"]]
        self.n_forces = int(z["n_forces"])
        self.n_cases = int(z["n_cases"])
        self.ndof_m = int(z["ndof_m"])

    def grid_dict(self, name, col_th=None):
        z = self.z
        ni, nj, nk = (int(v) for v in z[f"{name}_dims"])
        dx, dy, dz = (float(v) for v in z[f"{name}_spacing"])
        return {"ni": ni, "nj": nj, "nk": nk, "min_box": z[f"{name}_min_box"], "max_box": z[f"{name}_max_box"],
                "dx": dx, "dy": dy, "dz": dz, "col_th": float(z[f"{name}_col_th"]) if col_th is None else col_th,
                "grid": z[f"{name}_grid_f32"].astype(np.float64)}

    def points(self, name):
        return self.z[f"{name}_points"]

    def index(self, name):
        return int(self.z[f"{name}_index"])

    def force_bodies(self, fi):
        a, b = self.z[f"f{fi}_bodies"]
        return str(a), str(b)

    def case_params(self, c):
        kn, kt, mu, damp, scale, cscale = (float(v) for v in self.z[f"c{c}_params"])
        return {"kn": kn, "kt": kt, "mu": mu, "damping": damp, "scale": scale, "contact_scale": cscale}

    def case_pose(self, c, name):
        return self.z[f"c{c}_{name}_E"], self.z[f"c{c}_{name}_phi"]

    def out(self, c, fi, key):
        return self.z[f"c{c}_f{fi}_{key}"]


def load_contact_golden():
    return GoldenContact(np.load(os.path.join(GOLDEN, "contact_small.npz")))


def load_traj_golden():
    return np.load(os.path.join(GOLDEN, "traj_small.npz"))


# ---------------------------------------------------------------- stepper fixtures
JOINT_KEYS = {"translational": "translational", "free3d-exp": "free3d_exp"}


def traj_bodies(T, joint_type):
    """Body descriptors of the traj_small scene (part0 moving, part1 fixed): the
    reference's own mass properties / frames, dumped by oracle/make_golden.py."""
    key = JOINT_KEYS[joint_type]
    out = []
    for name, jt in (("part0", joint_type), ("part1", "fixed")):
        E_ij = T[f"{key}_{name}_E_ij"]
        E_ji = np.eye(4)
        E_ji[:3, :3] = E_ij[:3, :3].T
        E_ji[:3, 3] = -E_ij[:3, :3].T @ E_ij[:3, 3]
        out.append({"joint_type": jt, "E_pj_0": np.eye(4), "E_ji": E_ji, "inertia": T[f"{key}_{name}_inertia"],
                    "mass": float(T[f"{key}_{name}_mass"])})
    return out


def oracle_sim(G, T, joint_type, kn=1e6, kt=1e3, mu=0.0, damping=0.0):
    """oracle.sim_oracle.Sim of the traj_small scene."""
    from oracle import sdfc_oracle as O, sim_oracle as S
    O.build()
    jt = {"fixed": S.FIXED, "translational": S.TRANSLATIONAL, "free3d-exp": S.FREE3D_EXP}
    bodies = [S.BodyDesc(jt[b["joint_type"]], b["E_pj_0"], b["E_ji"], np.asarray(b["inertia"], dtype=np.float64), b["mass"])
              for b in traj_bodies(T, joint_type)]
    grids = {n: O.Grid.from_dict(G.grid_dict(n)) for n in G.names}
    pairs = [S.PairDesc(0, 1, grids["part1"], G.points("part0"), kn, kt, mu, damping),
             S.PairDesc(1, 0, grids["part0"], G.points("part1"), kn, kt, mu, damping)]
    key = JOINT_KEYS[joint_type]
    return S.Sim(bodies, pairs, float(T[f"{key}_h"]), np.array([0.0, 0.0, 1e-12]))


def load_freefall_golden():
    """tests/golden/freefall_small.npz -> (npz, scene dict in the layout of
    contact.build_simulation / oracle.ref_scene.extract_scene)."""
    z = np.load(os.path.join(GOLDEN, "freefall_small.npz"))
    names = [str(n) for n in z["names"]]
    bodies = []
    for n in names:
        ni, nj, nk = (int(v) for v in z[f"{n}_dims"])
        sp = z[f"{n}_spacing"]
        grid = {"ni": ni, "nj": nj, "nk": nk, "min_box": z[f"{n}_min_box"], "max_box": z[f"{n}_max_box"],
                "dx": float(sp[0]), "dy": float(sp[1]), "dz": float(sp[2]), "col_th": float(z[f"{n}_col_th"]),
                "grid": z[f"{n}_grid_f32"].astype(np.float64)}
        bodies.append({"name": n, "points": z[f"{n}_points"], "grid": grid, "joint_type": str(z[f"{n}_joint_type"]),
                       "E_pj_0": z[f"{n}_E_pj_0"], "E_ji": z[f"{n}_E_ji"], "inertia": z[f"{n}_inertia"],
                       "mass": float(z[f"{n}_mass"])})
    forces = []
    for i in range(int(z["n_forces"])):
        a, b, kn, kt, mu, damping = z[f"f{i}"]
        forces.append({"a": int(a), "b": int(b), "kn": kn, "kt": kt, "mu": mu, "damping": damping})
    grounds = []
    for i in range(int(z["n_grounds"])):
        kn, kt, mu, damping = z[f"g{i}_params"]
        grounds.append({"body": int(z[f"g{i}_body"]), "E_g": z[f"g{i}_E_g"], "kn": kn, "kt": kt, "mu": mu, "damping": damping})
    return z, {"bodies": bodies, "forces": forces, "grounds": grounds, "h": float(z["h"]), "gravity": z["gravity"]}


def oracle_sim_from_scene(sc):
    from oracle import sdfc_oracle as O, sim_oracle as S
    O.build()
    jt = {"fixed": S.FIXED, "translational": S.TRANSLATIONAL, "free3d-exp": S.FREE3D_EXP}
    grids = [O.Grid.from_dict(b["grid"]) for b in sc["bodies"]]
    bodies = [S.BodyDesc(jt[b["joint_type"]], b["E_pj_0"], b["E_ji"], b["inertia"], b["mass"]) for b in sc["bodies"]]
    pairs = [S.PairDesc(f["a"], f["b"], grids[f["b"]], sc["bodies"][f["a"]]["points"], f["kn"], f["kt"], f["mu"],
                        f["damping"]) for f in sc["forces"]]
    grounds = [S.GroundDesc(g["body"], sc["bodies"][g["body"]]["points"], g["E_g"], g["kn"], g["kt"], g["mu"],
                            g["damping"]) for g in sc.get("grounds", [])]
    return S.Sim(bodies, pairs, sc["h"], sc["gravity"], grounds=grounds)
