"""Development diagnostic: where does the GPU EnKF window of a golden case leave the reference bit for bit?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import helpers as H
import enkf_oracle as EO
from ravenhydroframework_b200 import api, enkf
case = sys.argv[1] if len(sys.argv) > 1 else "enkf_hbv_n64"
ref = dict(np.load(os.path.join(H.GOLDEN, case, "reference_enkf.npz")))
model = api.RavenModel(os.path.join(H.GOLDEN, case, "model"))
sim = enkf.EnKFSimulation(model, device=0)
out = sim.forecast()
def cmp(name, a, b):
    a, b = np.asarray(a), np.asarray(b)
    ne = (a != b)
    print("%-8s equal=%s  n_diff=%d of %d  max rel=%.3e" % (name, np.array_equal(a, b), int(ne.sum()), a.size, H.rel_err(a, b)))
    return ne
cmp("out", out, ref["out"])
st = sim.sim.download_state()
ne = cmp("state", st, ref["state"])
if ne.any():
    names = model.sv_names()
    idx = np.argwhere(ne)
    print("  differing SVs:", sorted(set(names[i] for i in idx[:, 2])), " members:", len(set(idx[:, 0])), " hrus:", len(set(idx[:, 1])))
    e, k, i = idx[0]
    print("  first: member %d hru %d sv %s gpu %r ref %r" % (e, k, names[i], st[e, k, i], ref["state"][e, k, i]))
q = sim.sim.hydrographs(0, sim.nsteps)
cmp("qlast", q[-1], ref["qout_last"])
pre, post = sim.analysis()
cmp("Xpre", pre, ref["Xpre"])
cmp("Xpost", post, ref["Xpost"])
Z = model.enkf_compute_Z(out)
Zr = model.enkf_compute_Z(ref["out"])
cmp("Z", Z, Zr)
Xa, Zo = EO.assimilation_calcs(pre, out, ref["obs"], ref["noise"])
cmp("dev-upd", post, Xa)
