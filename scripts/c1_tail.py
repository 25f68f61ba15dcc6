"""C1 free-fall batch: one forward(steps) after a warm-up forward, wall time and rounds. Run under
`ncu --metrics gpu__time_duration.sum --launch-skip N --launch-count M` to list the kernels of the straggler
rounds (a few environments whose Newton iteration does not converge keep the round loop alive).
Usage (GPU box): python scripts/c1_tail.py [n_envs] [steps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if __name__ == "__main__":
    from assemble_them_all_b200 import contact, scenes
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    ctx = contact.Context(0)
    rs = scenes.c1_freefall_rollouts(n_envs=n, steps=steps)
    sim, *keep = rs.upload(ctx)
    for rep in range(2):
        sim.reset(rs.q0, rs.qdot0)
        ctx.synchronize()
        r0 = sim.stats()["rounds"]
        l0 = ctx.launch_count
        t0 = time.perf_counter()
        sim.forward(steps)
        ctx.synchronize()
        dt = time.perf_counter() - t0
        st = sim.stats()
        print(f"rep {rep}: {1e3 * dt:.1f} ms, rounds {st['rounds'] - r0}, launches {ctx.launch_count - l0}, "
              f"{n * steps / dt:.0f} env-steps/s, converged {int(sim.is_converged().sum())}/{n}", flush=True)
    sim.close()
    ctx.close()
