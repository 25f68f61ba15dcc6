"""Experiment: one 1024-env stepper vs G groups of 1024/G envs driven by G host threads on G streams."""
import sys, time, threading, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from assemble_them_all_b200 import contact, scenes
import dataclasses

rs = scenes.c2_planner_rollouts(n_envs=1024)
for G in (1, 2, 4):
    n = 1024 // G
    groups = []
    for g in range(G):
        ctx = contact.Context(0)
        sub = dataclasses.replace(rs, q0=rs.q0[g * n:(g + 1) * n], qdot0=rs.qdot0[g * n:(g + 1) * n], f_ext=rs.f_ext[g * n:(g + 1) * n])
        sim, *keep = sub.upload(ctx)
        groups.append((ctx, sub, sim, keep))
    def prepare():
        for ctx, sub, sim, _ in groups:
            sim.reset(sub.q0, sub.qdot0)
            for b in range(len(sub.bodies)):
                sim.set_body_external_force(b, sub.f_ext[:, b])
    def run():
        th = [threading.Thread(target=sim.forward, args=(100,)) for _, _, sim, _ in groups]
        for t in th: t.start()
        for t in th: t.join()
    prepare(); run()
    ts = []
    for _ in range(4):
        prepare()
        for ctx, *_ in groups: ctx.synchronize()
        t0 = time.perf_counter(); run(); ts.append(time.perf_counter() - t0)
    q = np.concatenate([sim.get_q() for _, _, sim, _ in groups])
    print(G, "groups:", [round(1e3 * t, 2) for t in ts], "ms per forward(100); checksum", float(np.abs(q).sum()))
