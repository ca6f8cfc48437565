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
    ref = json.load(open(os.path.join(CASE, "reference_dds.json")))
    model = api.RavenModel(os.path.join(CASE, "model"))
    _check(dds.run_dds(model), ref)
