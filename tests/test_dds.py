"""DDS calibration (reference CDDSEnsemble, src/DDS.cpp): the host layer's perturbation sequence (rand() replay, reflecting bounds,
greedy acceptance) and objective function (NSE of the averaged hydrograph, negated) against the reference EXECUTABLE's own run of the
same 25 trials.  The fixture holds what the reference prints: 6 significant digits, hence the tolerance."""
import json
import os

import numpy as np
import pytest

import helpers as H
from ravenhydroframework_b200 import api, dds

CASE = os.path.join(H.GOLDEN, "dds_hmets")


def _check(result, ref):
    F = np.array([t[1] for t in result["trials"]])
    Fb = np.array([t[2] for t in result["trials"]])
    rF = np.array([t[0] for t in ref["trials"]])
    rFb = np.array([t[1] for t in ref["trials"]])
    assert len(F) == len(rF) == 25
    assert np.allclose(F, rF, rtol=2e-5) and np.allclose(Fb, rFb, rtol=2e-5)          # every trial: same parameters drawn, same objective
    assert [int(r[0]) for r in ref["improvements"]] == [t for t, _ in result["improvements"]]   # same accept / reject decisions
    assert np.allclose(result["best"], ref["best"], rtol=2e-5)


def test_dds_host_logic_matches_reference_executable():
    """CPU: the trial runs come from the oracle (checker), everything else is the product's host layer."""
    ref = json.load(open(os.path.join(CASE, "reference_dds.json")))
    model = api.RavenModel(os.path.join(CASE, "model"))
    model.flatten(1)
    n_trials = model.ensemble_begin()
    assert n_trials == 25
    q0 = dds.initial_outflows(model)
    trials, improvements, nP = [], [], 0
    for e in range(n_trials):
        vals = model.ensemble_update_model(e)
        nP = len(vals)
        model.flatten_member(0)
        r = H.oracle_run(model, model.nSteps, record_state=False)
        F, Fbest = dds.finish_trial(model, e, r["qout_rec"], q0)
        if F <= Fbest:
            improvements.append((e + 1, Fbest))
        trials.append((e + 1, F, Fbest))
    _check(dict(trials=trials, improvements=improvements, best=dds.best_parameters(model, nP)), ref)


@pytest.mark.gpu
def test_dds_on_gpu_matches_reference_executable():
    """objective evaluated on the device (rvn_gpu_diagnostic): only one double per trial comes back"""
    ref = json.load(open(os.path.join(CASE, "reference_dds.json")))
    model = api.RavenModel(os.path.join(CASE, "model"))
    on_device = dds.run_dds(model, device_objective=True)
    _check(on_device, ref)
    model2 = api.RavenModel(os.path.join(CASE, "model"))
    on_host = dds.run_dds(model2, device_objective=False)       # hydrographs copied back, host CalculateDiagnostic
    assert on_device["trials"] == on_host["trials"]             # same sums in the same order: identical bits


@pytest.mark.gpu
@pytest.mark.parametrize("kind", [0, 1, 2])
def test_device_diagnostics_equal_host_formulas(kind, tmp_path):
    """NSE / RMSE / KGE of every member's hydrograph on the device against a numpy restatement of src/Diagnostics.cpp, with blanks."""
    from ravenhydroframework_b200 import synth
    base = synth.write_hbv(str(tmp_path), n_sb=5, hru_per_sb=4, n_days=120, seed=5, routing="ROUTE_MUSKINGUM", season_offset=90)
    model = api.RavenModel(base)
    model.flatten(4)
    sim = api.GpuSimulation(model, n_members=4)
    sim.set_recording(120)
    sim.run(120)
    q = sim.hydrographs(0, 120)[:, :, 2]                        # [step][member]
    q0 = dds.initial_outflows(model)[2]
    rng = np.random.RandomState(kind)
    obs = np.abs(q[:, 0].mean() * (1 + 0.3 * rng.normal(size=121)))
    obs[rng.choice(121, 9, replace=False)] = -1.2345            # RAV_BLANK_DATA
    got = sim.diagnostic(kind, 2, obs, 1, 121, q0)
    for e in range(4):
        mod = np.empty(121)
        mod[0] = q0
        for nn in range(1, 121):
            ql = q[nn - 2, e] if nn >= 2 else q0
            mod[nn] = (0.5 * (q[nn - 1, e] + ql) * 86400.0) / 86400.0
        w = (obs != -1.2345).astype(float)
        w[0] = 0.0
        N = w.sum()
        if kind == 0:
            avg = (w * obs).sum() / N
            want = 1.0 - (w * (obs - mod) ** 2).sum() / (w * (obs - avg) ** 2).sum()
        elif kind == 1:
            want = np.sqrt((w * (obs - mod) ** 2).sum() / N)
        else:
            oa, ma = (obs * w).sum() / N, (mod * w).sum() / N
            os_, ms_ = np.sqrt(((obs - oa) ** 2 * w).sum() / N), np.sqrt(((mod - ma) ** 2 * w).sum() / N)
            r = ((obs - oa) * (mod - ma) * w).sum() / N / os_ / ms_
            want = 1.0 - np.sqrt((r - 1) ** 2 + (ms_ / os_ - 1) ** 2 + (ma / oa - 1) ** 2)
        assert abs(got[e] - want) <= 1e-12 * max(1.0, abs(want))
