"""Checkpoint / warm start (reference solution.rvc: CModel::WriteMajorOutput, src/StandardOutput.cpp:1388-1451; reader
src/ParseInitialConditionFile.cpp:312-660).  The reader is pinned on a file the reference executable wrote (tests/golden/hbv_restart,
covered by the golden tests); here the WRITER: a run interrupted by write -> reload -> continue equals the uninterrupted run bit for
bit when the file is written at full precision, and to ~1e-5 in the reference's own 5-decimal format.  Cumulative accounting
variables (ATMOSPHERE, ATMOS_PRECIP, GLACIER_ICE and the GLACIER_MB sum) are not restored by the reference's reader and are skipped."""
import datetime
import os
import shutil

import numpy as np
import pytest

import helpers as H
from ravenhydroframework_b200 import api, synth

N1, N2 = 90, 60


RESERVOIRS = (":Reservoir LakeA\n  :SubBasinID 2\n  :HRUID 6\n  :WeirCoefficient 0.6\n  :CrestWidth 4.5\n  :MaxDepth 6\n  :LakeArea 1.4e6\n:EndReservoir\n"
              ":Reservoir Out\n  :SubBasinID 1\n  :MaxDepth 8\n  :VolumeStageRelation POWER_LAW\n    2.4e6 1.6\n  :EndVolumeStageRelation\n"
              "  :OutflowStageRelation POWER_LAW\n    7.5 1.5\n  :EndOutflowStageRelation\n:EndReservoir\n")


def _inputs(tmp_path, reservoirs=False):
    base = synth.write_hbv(str(tmp_path / "full"), n_sb=7, hru_per_sb=4, n_days=N1 + N2, seed=31, routing="ROUTE_MUSKINGUM", name="model", season_offset=40)
    if reservoirs:   # :ResFlow / :ResStage lines of the checkpoint (CReservoir::WriteToSolutionFile)
        open(base + ".rvh", "a").write(RESERVOIRS)
    return base


def _restart_inputs(tmp_path, base, rvc):
    d = str(tmp_path / "warm")
    shutil.copytree(os.path.dirname(base), d)
    rvi = open(os.path.join(d, "model.rvi")).read()
    start = datetime.date(2000, 1, 1) + datetime.timedelta(days=N1)
    rvi = rvi.replace(":StartDate             2000-01-01 00:00:00", ":StartDate             %s 00:00:00" % start.isoformat())
    rvi = rvi.replace(":Duration              %d" % (N1 + N2), ":Duration              %d" % N2)
    open(os.path.join(d, "model.rvi"), "w").write(rvi)
    shutil.copy(rvc, os.path.join(d, "model.rvc"))
    return os.path.join(d, "model")


def _comparable(model):
    skip = ("ATMOSPHERE", "ATMOS_PRECIP", "GLACIER_ICE", "GLACIER_MB")
    return np.array([n not in skip for n in model.sv_names()])


@pytest.mark.parametrize("reservoirs", [False, True])
def test_oracle_run_survives_write_and_reload(tmp_path, reservoirs):
    base = _inputs(tmp_path, reservoirs)
    model = api.RavenModel(base)
    model.flatten(1)
    full = H.oracle_run(model, N1 + N2)
    part = H.oracle_run(model, N1)
    rvc = str(tmp_path / "ckpt.rvc")
    model.write_solution(rvc, N1 * model.timestep, part["state"], part["qin"], part["qlat"], part["qout"], part["scal"], full_precision=True, res=part["res"])
    warm = api.RavenModel(_restart_inputs(tmp_path, base, rvc))
    warm.flatten(1)
    cont = H.oracle_run(warm, N2)
    keep = _comparable(model)
    assert np.array_equal(cont["state"][:, keep], full["state"][:, keep])
    assert np.array_equal(cont["qout_rec"], full["qout_rec"][N1:])
    assert np.array_equal(cont["res"][:, :4], full["res"][:, :4]) and (cont["res"].shape[0] == (2 if reservoirs else 0))
    # the reference's own precision: 5 decimals
    rvc5 = str(tmp_path / "ckpt5.rvc")
    model.write_solution(rvc5, N1 * model.timestep, part["state"], part["qin"], part["qlat"], part["qout"], part["scal"], full_precision=False, res=part["res"])
    warm5 = api.RavenModel(_restart_inputs(tmp_path / "b", base, rvc5))
    warm5.flatten(1)
    cont5 = H.oracle_run(warm5, N2)
    assert H.rel_err(cont5["qout_rec"], full["qout_rec"][N1:]) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("reservoirs", [False, True])
def test_gpu_run_survives_write_and_reload(tmp_path, reservoirs):
    base = _inputs(tmp_path, reservoirs)
    model = api.RavenModel(base)
    model.flatten(2)
    sim = api.GpuSimulation(model, n_members=2)
    sim.set_recording(N1 + N2)
    sim.run(N1)
    rvc = str(tmp_path / "ckpt.rvc")
    sim.write_solution(rvc, member=1)
    sim.run(N2)
    full_state, full_q = sim.download_state()[1], sim.hydrographs(N1, N2)[:, 1, :]
    warm = api.RavenModel(_restart_inputs(tmp_path, base, rvc))
    warm.flatten(1)
    b = api.GpuSimulation(warm)
    b.set_recording(N2)
    b.run(N2)
    keep = _comparable(model)
    assert np.array_equal(b.download_state()[0][:, keep], full_state[:, keep])
    assert np.array_equal(b.hydrographs(0, N2)[:, 0, :], full_q)
    assert np.array_equal(b.download_reservoir_state()[0][:, :4], sim.download_reservoir_state()[1][:, :4])
