"""The reference's own shipped input sets (benchmarking/_InputFiles: the four Salmon River templates -- Salmon_GR4J is BASELINE
config 1 -- and Irondequoit), read unmodified by the host layer and run by the oracle (CPU) and by the CUDA path (GPU), against the
outputs of the unmodified reference executable on the same files (tests/golden/reference_inputs/*.npz, written by
tools/stage_reference_inputs.py).  The input files themselves are not part of this repository: they are read from
tests/_ref_inputs/ (staged, git-ignored, travels to the GPU box) or from /root/reference, and the tests skip when neither exists."""
import glob
import os

import numpy as np
import pytest

import helpers as H
from ravenhydroframework_b200 import api

TOL = 1e-9   # north_star: 1e-9 relative, FP64 throughout
OUT = os.path.join(H.GOLDEN, "reference_inputs")
SETS = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(OUT, "*.npz")))


def _base(name):
    roots = [os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref_inputs")]
    if not name.startswith("York"):   # York is pinned on a copy with one line dropped (tools/stage_reference_inputs.py): staged copy only
        roots.append("/root/reference/benchmarking/_InputFiles")
    folder, _, stem = name.partition("__")   # "dir__basename": several set-ups in one folder (York_nc)
    for root in roots:
        rvi = glob.glob(os.path.join(root, folder, (stem or "*") + ".rvi"))
        if rvi:
            return rvi[0][:-4]
    pytest.skip("reference input set %s not staged here (tools/stage_reference_inputs.py)" % name)


def _duration_days(base, n):
    """n model steps in days, from the :TimeStep of the .rvi (daily unless hh:mm:ss)."""
    for line in open(base + ".rvi"):
        t = line.split()
        if len(t) >= 2 and t[0] == ":TimeStep" and ":" in t[1]:
            h, m, sec = (float(x) for x in t[1].split(":"))
            return n * (h / 24.0 + m / 1440.0 + sec / 86400.0)
    return n


def test_sets_are_pinned():
    assert {"Salmon_GR4J", "Salmon_HMETS", "Salmon_HBV", "Salmon_MOHYSE", "Irondequoit", "York_nongridded_m_daily_i_daily", "LOTW", "Nith",
            "York_nc__York_nongridded_m_subdaily_i_subdaily", "York_nc__York_nongridded_m_subdaily_i_daily", "York_nc__York_nongridded_m_daily_i_subdaily", "York_nc__York"} <= set(SETS)


@pytest.mark.parametrize("name", SETS)
def test_oracle_on_reference_inputs(name):
    g = np.load(os.path.join(OUT, name + ".npz"))
    n = g["qout"].shape[0]
    base = _base(name)
    model = api.RavenModel(base, duration=_duration_days(base, n))
    model.flatten(1)
    import ctypes as C
    model.lib.rvh_desc_sv_table.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    t = np.zeros(model.NS, dtype=np.int32); l = np.zeros(model.NS, dtype=np.int32)
    model.lib.rvh_desc_sv_table(model.desc, t.ctypes.data, l.ctypes.data)
    assert np.array_equal(t, g["sv_type"]) and np.array_equal(l, g["sv_layer"])   # state-variable catalogue in the reference's order
    o = H.oracle_run(model, n)
    assert H.rel_err(o["state_rec"][g["state_steps"] - 1], g["state"]) < TOL
    assert H.rel_err(o["qout_rec"], g["qout"]) < TOL
    assert H.rel_err(o["wshed_rec"][:, :3], g["wshed"]) < TOL
    if "res_stage" in g:   # reservoir stages at the end of the run (the last recorded step is the last step)
        _, sb = model.reservoir_state()
        assert g["state_steps"][-1] == n and H.rel_err(o["res"][:, 0], g["res_stage"][-1][sb]) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("name", SETS)
def test_gpu_on_reference_inputs(name):
    g = np.load(os.path.join(OUT, name + ".npz"))
    n = g["qout"].shape[0]
    base = _base(name)
    model = api.RavenModel(base, duration=_duration_days(base, n))
    model.flatten(2)
    sim = api.GpuSimulation(model, n_members=2)
    sim.set_recording(n)
    got = []
    for s in g["state_steps"]:
        sim.run(int(s) - sim.step)
        got.append(sim.download_state()[1])
    assert H.rel_err(np.array(got), g["state"]) < TOL
    assert H.rel_err(sim.hydrographs(0, n)[:, 1, :], g["qout"]) < TOL
    assert H.rel_err(sim.watershed(0, n)[:, 1, :3], g["wshed"]) < TOL
    if "res_stage" in g:
        _, sb = model.reservoir_state()
        assert H.rel_err(sim.download_reservoir_state()[1][:, 0], g["res_stage"][-1][sb]) < TOL
    sim.close()
