"""Shared test utilities: the CPU oracle (oracle/liboracle.so) and the reference dump driver (oracle/_ref).

The oracle and the reference build are TEST INFRASTRUCTURE: they are only ever used here as checkers.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from ravenhydroframework_b200 import api  # noqa: E402

REF_DUMP = os.path.join(ROOT, "oracle", "_ref", "ref_dump")
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE_INPUTS = "/root/reference/benchmarking/_InputFiles"
_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        lib = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
        vp, i = C.c_void_p, C.c_int
        lib.rvo_run_member.argtypes = [vp, i] + [vp] * 6 + [i, i] + [vp] * 4
        lib.rvo_run_member_res.argtypes = [vp, i] + [vp] * 6 + [i, i] + [vp] * 5
        _oracle = lib
    return _oracle


def have_reference():
    return os.path.exists(REF_DUMP)


def oracle_run(model, nsteps, member=0, state=None, record_state=True, record_forcings=False):
    """Run the C oracle for one member of a flattened model. Returns a dict of arrays."""
    lib = oracle_lib()
    st = (model.initial_state() if state is None else np.array(state, dtype=np.float64)).copy()
    qin, qlat, qout, scal = [a.copy() for a in model.basin_state()]
    cumul = np.zeros(2)
    qrec = np.zeros((nsteps, model.nSB))
    wrec = np.zeros((nsteps, 4))
    srec = np.zeros((nsteps, model.nHRU, model.NS)) if record_state else None
    frec = np.zeros((nsteps, model.nHRU, api.F_COUNT)) if record_forcings else None
    res = model.reservoir_state()[0].copy()   # [nRes][8]: stage, stage_last, Qout, Qout_last, precip, losses, AET (empty without reservoirs)
    rc = lib.rvo_run_member_res(model.desc, member, st.ctypes.data, qin.ctypes.data, qlat.ctypes.data, qout.ctypes.data, scal.ctypes.data,
                                cumul.ctypes.data, 0, nsteps, qrec.ctypes.data, wrec.ctypes.data,
                                srec.ctypes.data if record_state else None, frec.ctypes.data if record_forcings else None,
                                res.ctypes.data if res.size else None)
    if rc != 0:
        raise RuntimeError("oracle does not implement one of the model's processes (rc=%d)" % rc)
    return dict(state=st, qin=qin, qlat=qlat, qout=qout, scal=scal, cumul=cumul, qout_rec=qrec, wshed_rec=wrec, state_rec=srec, forc_rec=frec, res=res)


def oracle_flux_balance(model, nsteps, member=0):
    """Cumulative per-connection fluxes [nConn][nHRU] of an oracle run (reference CModel::IncrementBalance)."""
    lib = oracle_lib()
    f, _ = model.connections()
    buf = np.zeros((f.size, model.nHRU))
    lib.rvo_set_flux_buffer.argtypes = [C.c_void_p]
    lib.rvo_set_flux_buffer(buf.ctypes.data)
    try:
        oracle_run(model, nsteps, member=member, record_state=False)
    finally:
        lib.rvo_set_flux_buffer(None)
    return buf


def flux_by_state_variable(model, cum):
    """[nHRU][NS][2]: what CModel::GetCumulativeFlux(k, i, to=True / False) sums (src/Model.cpp:902-923) from per-connection arrays cum[nConn][nHRU]."""
    f, t = model.connections()
    out = np.zeros((model.nHRU, model.NS, 2))
    for q in range(f.size):   # process-list order, as the reference's loop
        out[:, t[q], 0] += cum[q]
        out[:, f[q], 1] += cum[q]
    return out


def reference_run(base, outdir, nsteps=-1):
    """Run the unmodified reference on <base>.rv* and load its in-memory FP64 dumps."""
    os.makedirs(outdir, exist_ok=True)
    if not outdir.endswith("/"):
        outdir += "/"
    r = subprocess.run([REF_DUMP, os.path.basename(base), outdir, str(nsteps)], cwd=os.path.dirname(base) or ".",
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0 or not os.path.exists(outdir + "state.f64"):
        raise RuntimeError("reference run failed:\n" + r.stdout[-3000:])
    return load_reference_dump(outdir)


def load_reference_dump(outdir):
    if not outdir.endswith("/"):
        outdir += "/"
    meta = open(outdir + "meta.txt").read().split("\n")
    kv = {l.split()[0]: l.split()[1] for l in meta if l and l.split()[0] in ("nHRU", "NS", "NB", "NF", "nsteps", "dt")}
    nH, NS, NB, NF = int(kv["nHRU"]), int(kv["NS"]), int(kv["NB"]), int(kv["NF"])
    out = dict(meta=meta, nHRU=nH, NS=NS, NB=NB)
    out["state"] = np.fromfile(outdir + "state.f64").reshape(-1, nH, NS)
    out["forc"] = np.fromfile(outdir + "forc.f64").reshape(-1, nH, NF)
    out["basin"] = np.fromfile(outdir + "basin.f64").reshape(-1, NB, 6)
    out["wshed"] = np.fromfile(outdir + "wshed.f64").reshape(-1, 3)
    if os.path.exists(outdir + "res.f64"):   # reservoir stage, outflow, losses, storage per basin (zeros without a reservoir)
        out["res"] = np.fromfile(outdir + "res.f64").reshape(-1, NB, 4)
    if os.path.exists(outdir + "cumflux.f64"):   # CModel::GetCumulativeFlux(k, i, to / from) at the end of the run
        out["cumflux"] = np.fromfile(outdir + "cumflux.f64").reshape(nH, NS, 2)
    out["sv"] = [(int(l.split()[2]), int(l.split()[3])) for l in meta if l.startswith("SV ")]
    out["order"] = [int(l.split()[2]) for l in meta if l.startswith("ORDER ")]
    out["proc"] = [l.split()[1:] for l in meta if l.startswith("PROC ")]
    out["param"] = [l.split()[1:] for l in meta if l.startswith("PARAM ")]
    return out


def rel_err(a, b, floor=1e-3):
    """max |a-b| / max(|b|, floor).  With the 1e-9 tolerance of the north_star this is "1e-9 relative, with an
    absolute floor of 1e-12 [mm or m3/s] for stores that pass through zero" (SURVEY App. E)."""
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0
