"""DDS calibration on the GPU (reference CDDSEnsemble, src/DDS.cpp).  Trials are sequential by construction (trial e+1 perturbs the
best vector found so far), so each trial is one full device run of a single member; what the GPU parallelises is the HRUs.
The host layer (CDDSEnsemble, host/rvh_ensemble.cpp) draws from the reference's rand() stream and evaluates the objective."""
import ctypes as C

import numpy as np

from . import api


def _decl(lib):
    lib.rvh_dds_finish_run.restype = C.c_double
    lib.rvh_dds_finish_run.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_double)]
    lib.rvh_dds_best.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.rvh_desc_nseg.argtypes = [C.c_void_p, C.c_void_p]


def initial_outflows(model):
    """CSubBasin::GetOutflowRate of every basin before the first step."""
    nseg = (C.c_int32 * model.nSB)()
    model.lib.rvh_desc_nseg(model.desc, nseg)
    qout = model.basin_state()[2]
    return np.array([qout[p, nseg[p] - 1] for p in range(model.nSB)])


def finish_trial(model, e, qout_rec, qout_initial):
    """CDDSEnsemble::FinishEnsembleRun(e): objective of the trial from its hydrographs [nsteps][nSB]; returns (Ftest, Fbest)."""
    _decl(model.lib)
    q = np.ascontiguousarray(qout_rec, dtype=np.float64)
    q0 = np.ascontiguousarray(qout_initial, dtype=np.float64)
    fb = C.c_double()
    F = model.lib.rvh_dds_finish_run(model.h, int(e), q.ctypes.data, q0.ctypes.data, q.shape[0], C.byref(fb))
    if F <= -1e299:
        raise api.RavenError(model.lib.rvh_last_error().decode())
    return F, fb.value


def objective_inputs(model):
    """(basin index, diagnostic type, nnstart, nnend, observations sampled on the model step) of the DDS calibration target."""
    lib = model.lib
    lib.rvh_dds_objective_inputs.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4 + [C.c_void_p, C.c_int]
    b, t, a, e = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    n = lib.rvh_dds_objective_inputs(model.h, C.byref(b), C.byref(t), C.byref(a), C.byref(e), None, 0)
    if n < 0:
        raise api.RavenError(lib.rvh_last_error().decode())
    obs = np.zeros(n)
    lib.rvh_dds_objective_inputs(model.h, C.byref(b), C.byref(t), C.byref(a), C.byref(e), obs.ctypes.data, n)
    return b.value, t.value, a.value, e.value, obs


def finish_trial_value(model, e, diagnostic):
    """CDDSEnsemble::FinishEnsembleRun(e) for a trial whose diagnostic was evaluated on the device; returns (Ftest, Fbest)."""
    lib = model.lib
    lib.rvh_dds_finish_value.restype = C.c_double
    lib.rvh_dds_finish_value.argtypes = [C.c_void_p, C.c_int, C.c_double, C.POINTER(C.c_double)]
    fb = C.c_double()
    F = lib.rvh_dds_finish_value(model.h, int(e), float(diagnostic), C.byref(fb))
    if F <= -1e299:
        raise api.RavenError(lib.rvh_last_error().decode())
    return F, fb.value


def best_parameters(model, n):
    _decl(model.lib)
    out = np.zeros(n)
    k = model.lib.rvh_dds_best(model.h, out.ctypes.data, n)
    if k < 0:
        raise api.RavenError(model.lib.rvh_last_error().decode())
    return out[:k]


def run_dds(model, device=0, device_objective=True):
    """The reference's member loop for ENSEMBLE_DDS (src/RavenMain.cpp:115-196): returns dict(trials=[(e+1, Ftest, Fbest)],
    best=parameter vector, improvements=[(trial, Fbest)] = the rows of DDSOutput.csv).  device_objective: the trial's diagnostic is
    evaluated on the device (rvn_gpu_diagnostic, same sums in the same order) and only one double comes back per trial."""
    model.flatten(1)
    n_trials = model.ensemble_begin()
    sim = api.GpuSimulation(model, n_members=1, device=device)
    n = model.nSteps
    trials, improvements, nP = [], [], 0
    for e in range(n_trials):
        vals = model.ensemble_update_model(e)          # CDDSEnsemble::UpdateModel: perturb + CModel::UpdateParameter + ParseInitialConditions
        nP = len(vals)
        sim.upload_member_params(0)
        # every trial restarts from the .rvc state as re-read under the trial's parameters (the reference re-runs
        # ParseInitialConditions after UpdateParameter, src/DDS.cpp, so stores are clipped to the trial's capacities)
        st0 = model.initial_state()
        basin0 = model.basin_state()
        q0 = initial_outflows(model)
        sim.upload_state(st0, 0, 1)
        sim._ck(sim.lib.rvn_gpu_upload_basin_state(sim.h, basin0[0].ctypes.data, basin0[1].ctypes.data, basin0[2].ctypes.data, basin0[3].ctypes.data, 0, 1))
        sim.set_recording(n, hydrographs=True, watershed=False)   # buffers are allocated once and reused (rvn_gpu_set_recording)
        sim.step = 0
        sim.run(n)
        if device_objective:
            basin, kind, nn0, nn1, obs = objective_inputs(model)
            F, Fbest = finish_trial_value(model, e, sim.diagnostic(kind, basin, obs, nn0, nn1, q0[basin], 0, 1)[0])
        else:
            F, Fbest = finish_trial(model, e, sim.hydrographs(0, n)[:, 0, :], q0)
        if F <= Fbest:                                 # accepted: this trial is the new best (a row of DDSOutput.csv)
            improvements.append((e + 1, Fbest))
        trials.append((e + 1, F, Fbest))
    sim.close()
    return dict(trials=trials, improvements=improvements, best=best_parameters(model, nP))
