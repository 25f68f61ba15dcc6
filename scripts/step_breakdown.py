"""Where a forward(frame_skip) of the batched stepper spends its time: per-slot CUDA-event totals of
the contact kernels (Context.enable_timing) against the wall time of the call, for SDFC_SKIN=0/1 and
1 / 2 lanes. Usage (GPU box): python scripts/step_breakdown.py [n_envs]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(skin, lanes, n_envs, joint_type="translational"):
    os.environ["SDFC_SKIN"] = skin
    os.environ["SDFC_SIM_LANES"] = lanes
    from assemble_them_all_b200 import contact, scenes
    ctx = contact.Context(0)
    rs = scenes.c2_planner_rollouts(n_envs=n_envs, joint_type=joint_type)
    sim, batch, *keep = rs.upload(ctx)
    out = {}
    for timing in (False, True):
        sim.reset(rs.q0, rs.qdot0)
        for b in range(len(rs.bodies)):
            sim.set_body_external_force(b, rs.f_ext[:, b])
        sim.forward(5)
        ctx.enable_timing(timing)
        ctx.read_timing(reset=True)
        ctx.synchronize()
        t0 = time.perf_counter()
        sim.forward(rs.frame_skip - 5)
        out["timed" if timing else "plain"] = 1e3 * (time.perf_counter() - t0)
        if timing:
            tm = ctx.read_timing()
            out["slots_ms"] = {k: round(v[0], 2) for k, v in tm.items()}
            out["slot_launches"] = {k: v[1] for k, v in tm.items()}
        ctx.enable_timing(False)
    out["rounds"] = sim.stats()["rounds"]
    out["skin"] = batch.skin_stats()
    sim.close()
    ctx.close()
    return out


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    for jt in ("translational", "free3d-exp"):
        for skin in ("0", "1"):
            for lanes in ("1", "2"):
                print(jt, "skin", skin, "lanes", lanes, run(skin, lanes, n, jt), flush=True)
