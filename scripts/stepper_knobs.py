"""forward(frame_skip) of the batched stepper on the C2 scenes for lane counts x round-kernel variants
(SDFC_SIM_LANES, SDFC_SIM_TAIL: 0 = warp per environment always, 100000 = CTA per environment always).
Usage (GPU box): python scripts/stepper_knobs.py [n_envs]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(lanes, tail, n_envs, joint_type, n_vertices=10000):
    os.environ["SDFC_SIM_LANES"] = lanes
    if tail is None:
        os.environ.pop("SDFC_SIM_TAIL", None)
    else:
        os.environ["SDFC_SIM_TAIL"] = tail
    from assemble_them_all_b200 import contact, scenes
    ctx = contact.Context(0)
    rs = scenes.c2_planner_rollouts(n_envs=n_envs, joint_type=joint_type, n_vertices=n_vertices)
    sim, batch, *keep = rs.upload(ctx)
    best = 1e9
    for rep in range(3):
        sim.reset(rs.q0, rs.qdot0)
        for b in range(len(rs.bodies)):
            sim.set_body_external_force(b, rs.f_ext[:, b])
        ctx.synchronize()
        t0 = time.perf_counter()
        sim.forward(rs.frame_skip)
        ctx.synchronize()
        if rep:
            best = min(best, 1e3 * (time.perf_counter() - t0))
    q = sim.get_state()[0]
    sim.close()
    ctx.close()
    return best, float(abs(q).sum())


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    if len(sys.argv) > 2 and sys.argv[2] == "gv":   # rounds per item of the value-only force list
        for jt in ("translational", "free3d-exp"):
            for gv in ("1", "2", "4", "8"):
                os.environ["SDFC_FORCE_GV"] = gv
                ms, chk = run("2", None, n, jt)
                print(f"{jt:14s} SDFC_FORCE_GV={gv}: {ms:8.2f} ms  ({n * 100 / ms * 1e3:,.0f} env-steps/s) checksum {chk:.12e}", flush=True)
        sys.exit(0)
    for jt in ("translational", "free3d-exp"):
        for tail in ("0", None, "100000"):
            for lanes in ("1", "2", "3", "4"):
                ms, chk = run(lanes, tail, n, jt)
                print(f"{jt:14s} tail={tail} lanes={lanes}: {ms:8.2f} ms  ({n * 100 / ms * 1e3:,.0f} env-steps/s) checksum {chk:.12e}", flush=True)
