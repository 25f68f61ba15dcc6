import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from assemble_them_all_b200 import contact, scenes
sc = scenes.c2_bfs_rollouts(n_envs=256)
ctx = contact.Context(0)
batch, *keep = sc.upload(ctx)
os.environ.pop("SDFC_DEBUG_SPLIT_EVAL", None)
ref = batch.eval(sc.E, sc.phi, deriv=True)
refv = batch.eval(sc.E, sc.phi, deriv=False)
for c0 in (128, 40, 250):
    os.environ["SDFC_DEBUG_SPLIT_EVAL"] = str(c0)
    out = batch.eval(sc.E, sc.phi, deriv=True)
    outv = batch.eval(sc.E, sc.phi, deriv=False)
    for name, a, b in zip(("fm", "Km", "Dm", "nact"), ref, out):
        d = np.abs(np.asarray(a, dtype=np.float64) - b).reshape(len(a), -1).max(axis=1)
        print(c0, name, "envs differing:", int((d > 0).sum()), np.nonzero(d)[0][:6], d.max())
    d = np.abs(refv[0] - outv[0]).max(axis=1); print(c0, "fm(value)", int((d > 0).sum()), d.max())
