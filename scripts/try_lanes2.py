import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from assemble_them_all_b200 import contact, scenes
rs = scenes.c2_planner_rollouts(n_envs=256)
sims = []
for L in (1, 2):
    os.environ["SDFC_SIM_LANES"] = str(L)
    ctx = contact.Context(0)
    sim, *keep = rs.upload(ctx)
    sim.reset(rs.q0, rs.qdot0)
    for b in range(len(rs.bodies)):
        sim.set_body_external_force(b, rs.f_ext[:, b])
    sims.append((sim, keep, ctx))
os.environ["SDFC_SIM_LANE_SYNC"] = "1"
os.environ["SDFC_SIM_LANE_SPLIT"] = "40"
for step in range(60):
    st = []
    for sim, *_ in sims:
        s0 = sim.stats(); sim.forward(1); s1 = sim.stats()
        st.append({k: s1[k] - s0[k] for k in s0})
    qa, qda = sims[0][0].get_state(); qb, qdb = sims[1][0].get_state()
    d = np.abs(qa - qb).max(axis=1)
    if (d > 0).any():
        e = int(np.nonzero(d)[0][0])
        print("first difference at step", step, "envs", np.nonzero(d)[0][:10], "max", d.max(), st)
        print("env", e, "q", qa[e], qb[e], "touched", sims[0][0].get_contact_history()[e], sims[1][0].get_contact_history()[e])
        break
else:
    print("no difference in 60 steps")
