"""Raven BMI surface (reference src/Raven_BMI.h: extern "C" bmi_model_create/bmi_model_destroy + class CRavenBMI) over the GPU model."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers as H
from ravenhydroframework_b200 import api, synth


def _lib():
    lib = api.host_lib()
    lib.bmi_model_create.restype = C.c_void_p
    lib.bmi_model_destroy.argtypes = [C.c_void_p]
    lib.bmi_last_error.restype = C.c_char_p
    vp = C.c_void_p
    lib.bmi_initialize.argtypes = [vp, C.c_char_p]
    lib.bmi_update.argtypes = [vp]
    lib.bmi_update_until.argtypes = [vp, C.c_double]
    lib.bmi_finalize.argtypes = [vp]
    lib.bmi_get_value.argtypes = [vp, C.c_char_p, vp]
    lib.bmi_set_value.argtypes = [vp, C.c_char_p, vp]
    lib.bmi_get_value_at_indices.argtypes = [vp, C.c_char_p, vp, vp, C.c_int]
    lib.bmi_get_var_grid.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int)]
    lib.bmi_get_grid_size.argtypes = [vp, C.c_int, C.POINTER(C.c_int)]
    lib.bmi_get_current_time.restype = C.c_double
    lib.bmi_get_current_time.argtypes = [vp]
    lib.bmi_get_end_time.restype = C.c_double
    lib.bmi_get_end_time.argtypes = [vp]
    return lib


def _config(tmp_path, base, duration, extra=""):
    cfg = os.path.join(str(tmp_path), "bmi_config.txt")
    with open(cfg, "w") as f:
        f.write("# Raven BMI configuration\ncli_args: %s -o out/\nduration: %d\ninput_vars:\n- SOIL[0]: hru_state\noutput_vars:\n"
                "- SOIL[0]: hru_state\n- SNOW: hru_state\n- STREAMFLOW: subbasin_state\n- PRECIP: forcing\n%s" % (os.path.basename(base), duration, extra))
    return cfg


def test_bmi_config_errors_are_logic_errors(tmp_path):
    """Error behaviour of the reference (src/Raven_BMI.cpp:113-270): std::logic_error with these messages, before any device work."""
    lib = _lib()
    p = lib.bmi_model_create()
    assert lib.bmi_initialize(p, b"/nonexistent/config.txt") == -1
    assert b"Cannot find configuration file" in lib.bmi_last_error()
    bad = os.path.join(str(tmp_path), "bad.txt")
    open(bad, "w").write("cli_args: model\n")
    assert lib.bmi_initialize(p, bad.encode()) == -1 and b"missing duration" in lib.bmi_last_error()
    open(bad, "w").write("duration: 10\n")
    assert lib.bmi_initialize(p, bad.encode()) == -1 and b"missing cli_args" in lib.bmi_last_error()
    open(bad, "w").write("cli_args: model\nduration: 10\noutput_vars:\n- OUTFLOW: subbasin_state\n")
    assert lib.bmi_initialize(p, bad.encode()) == -1 and b"Invalid subbasin state variable" in lib.bmi_last_error()
    lib.bmi_model_destroy(p)


@pytest.mark.gpu
def test_bmi_driven_run_equals_batch_run(tmp_path):
    lib = _lib()
    n = 60
    base = synth.write_hbv(str(tmp_path), n_sb=5, hru_per_sb=6, n_days=90, seed=8, routing="ROUTE_MUSKINGUM")
    cfg = _config(tmp_path, base, n)
    p = lib.bmi_model_create()
    assert lib.bmi_initialize(p, cfg.encode()) == 0, lib.bmi_last_error()
    g, nh, nb = C.c_int(), C.c_int(), C.c_int()
    assert lib.bmi_get_var_grid(p, b"STREAMFLOW", C.byref(g)) == 0 and g.value == 1
    assert lib.bmi_get_var_grid(p, b"SNOW", C.byref(g)) == 0 and g.value == 0
    lib.bmi_get_grid_size(p, 0, C.byref(nh)); lib.bmi_get_grid_size(p, 1, C.byref(nb))
    assert (nh.value, nb.value) == (30, 5)
    assert lib.bmi_get_end_time(p) == float(n)
    model = api.RavenModel(base, duration=n)
    model.flatten(1)
    sim = api.GpuSimulation(model)
    sim.set_recording(n)
    sim.enable_forcing_output(True)
    soil0 = model.sv_names().index("SOIL[0]")
    q = np.zeros(nb.value); s = np.zeros(nh.value); pr = np.zeros(nh.value)
    for step in range(20):
        assert lib.bmi_update(p) == 0, lib.bmi_last_error()
        sim.run(1)
        assert lib.bmi_get_value(p, b"STREAMFLOW", q.ctypes.data) == 0
        assert lib.bmi_get_value(p, b"SOIL[0]", s.ctypes.data) == 0
        assert lib.bmi_get_value(p, b"PRECIP", pr.ctypes.data) == 0
        assert np.array_equal(q, sim.hydrographs(step, 1)[0, 0, :])
        assert np.array_equal(s, sim.download_state()[0][:, soil0])
        assert np.array_equal(pr, sim.forcings()[0][:, api.FORCING_NAMES.index("PRECIP")])
    assert lib.bmi_get_current_time(p) == 20.0
    # SetValue on an hru_state input, then UpdateUntil: the batch run with the same state replacement must agree
    s2 = s * 0.5
    assert lib.bmi_set_value(p, b"SOIL[0]", s2.ctypes.data) == 0
    st = sim.download_state()
    st[0][:, soil0] = s2
    sim.upload_state(st, 0, 1)
    assert lib.bmi_update_until(p, 39.0) == 0      # steps at t = 20 .. 39
    sim.run(20)
    assert lib.bmi_get_current_time(p) == 40.0
    assert lib.bmi_get_value(p, b"SOIL[0]", s.ctypes.data) == 0
    assert np.array_equal(s, sim.download_state()[0][:, soil0])
    idx = np.array([3, 0], dtype=np.int32); two = np.zeros(2)
    assert lib.bmi_get_value_at_indices(p, b"STREAMFLOW", two.ctypes.data, idx.ctypes.data, 2) == 0
    lib.bmi_get_value(p, b"STREAMFLOW", q.ctypes.data)
    assert two[0] == q[3] and two[1] == q[0]
    assert lib.bmi_get_value(p, b"CANOPY", s.ctypes.data) == -1 and b"not included in config file" in lib.bmi_last_error()
    assert lib.bmi_finalize(p) == 0
    lib.bmi_model_destroy(p)
