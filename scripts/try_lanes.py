import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from assemble_them_all_b200 import contact, scenes
rs = scenes.c2_planner_rollouts(n_envs=640)
res = {}
for L, sync, split in ((1, 0, None), (2, 1, 32), (2, 1, 64), (2, 1, 33), (2, 1, 512)):
    os.environ["SDFC_SIM_LANES"] = str(L)
    for k, v in (("SDFC_SIM_LANE_SYNC", sync and "1"), ("SDFC_SIM_LANE_SPLIT", split is not None and str(split))):
        if v: os.environ[k] = v
        else: os.environ.pop(k, None)
    ctx = contact.Context(0)
    sim, *keep = rs.upload(ctx)
    sim.reset(rs.q0, rs.qdot0)
    for b in range(len(rs.bodies)):
        sim.set_body_external_force(b, rs.f_ext[:, b])
    s0 = sim.stats()
    sim.forward(100)
    s1 = sim.stats()
    res[(L, sync, split)] = sim.get_q()
    d = np.abs(res[(1, 0, None)] - res[(L, sync, split)]).max(axis=1)
    print(L, sync, split, {k: s1[k] - s0[k] for k in s0}, "differing envs:", int((d > 0).sum()), np.nonzero(d)[0][:8])
