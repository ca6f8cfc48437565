"""Parity at the BASELINE.json sizes (GPU only).  The C oracle is fast enough to check ONE member of the full-size watershed
completely; the other members are covered by size-independent properties: members with identical parameters are bit-identical,
members with sampled parameters differ, and every member closes its water balance (the MB Error column of the reference's
WatershedStorage.csv: (S - S0) + (cumulative out - cumulative in) = 0)."""
import numpy as np
import pytest

import helpers as H
from ravenhydroframework_b200 import api, synth

pytestmark = pytest.mark.gpu
TOL = 1e-9


def test_config3_share_hmets_512_members_x_10k_hrus(tmp_path):
    """BASELINE config 3, the per-GPU share: HMETS, 10,000 HRUs / 500 subbasins, 512 members, Monte-Carlo parameters."""
    n = 12
    base = synth.write_hmets(str(tmp_path), n_sb=500, hru_per_sb=20, n_days=n + 1, seed=1, single_class=True, ensemble_members=512, season_offset=105)
    model = api.RavenModel(base)
    E = 512
    model.flatten(E)
    sim = api.GpuSimulation(model, n_members=E)
    assert sim.kernel_variant() == "static:HMETS"
    samples = sim.apply_ensemble(0)                      # CMonteCarloEnsemble::UpdateModel for all 512 members (rand() replay)
    assert samples.shape[0] == E and len(np.unique(samples[:, 0])) > 500
    st0 = model.initial_state()
    sim.set_recording(n)
    sim.run(n)
    # one complete member against the oracle (member 37: its own sampled parameters)
    ref = H.oracle_run(model, n, member=37)
    st = sim.download_state(37, 1)[0]
    assert H.rel_err(st, ref["state"]) < TOL
    assert H.rel_err(sim.hydrographs(0, n, 37, 1)[:, 0, :], ref["qout_rec"]) < TOL
    assert H.rel_err(sim.watershed(0, n, 37, 1)[:, 0, :3], ref["wshed_rec"][:, :3]) < TOL
    # every member: finite, different from its neighbours, balance closed
    w = sim.watershed(0, n)                               # [n][E][4]
    assert np.all(np.isfinite(w))
    s_init = ref["wshed_rec"][0, 3] - (ref["wshed_rec"][0, 0] - ref["wshed_rec"][0, 1])   # S0 of a member = S1 - (in1 - out1); same initial state for all
    mb = (w[:, :, 3] - s_init) + (w[:, :, 1] - w[:, :, 0])
    assert np.abs(mb).max() < 1e-7 * max(1.0, np.abs(w[:, :, 3]).max()), np.abs(mb).max()
    q = sim.hydrographs(0, n)[:, :, 0]
    assert len(np.unique(q[-1])) > 500
    sim.close()
    assert st0.shape == (10000, model.NS)


def test_config2_hbvec_10k_hrus_500_subbasins_muskingum(tmp_path):
    """BASELINE config 2: HBV-EC semi-distributed, 10,000 HRUs / 500 subbasins, Muskingum routing -- complete comparison with the
    oracle (every HRU storage, every subbasin hydrograph), plus identical replicas."""
    n = 60
    base = synth.write_hbv(str(tmp_path), n_sb=500, hru_per_sb=20, n_days=n, seed=2, routing="ROUTE_MUSKINGUM", season_offset=100)
    model = api.RavenModel(base)
    model.flatten(3)
    sim = api.GpuSimulation(model, n_members=3)
    assert sim.kernel_variant() == "static:HBVEC"
    sim.set_recording(n)
    sim.run(n)
    ref = H.oracle_run(model, n)
    st = sim.download_state()
    assert H.rel_err(st[0], ref["state"]) < TOL
    assert np.array_equal(st[0], st[1]) and np.array_equal(st[0], st[2])
    assert H.rel_err(sim.hydrographs(0, n)[:, 0, :], ref["qout_rec"]) < TOL
    assert H.rel_err(sim.watershed(0, n)[:, 0, :3], ref["wshed_rec"][:, :3]) < TOL
    assert H.rel_err(sim.cumulative()[0], ref["cumul"]) < TOL
    sim.close()


def test_config4_blended_distributed_muskingum(tmp_path):
    """BASELINE config 4 shape at a size the oracle finishes in seconds: Blended process groups, 40,000 HRUs / 2,000 subbasins,
    9-station interpolation (the sparse gridded-forcing gather), hourly step, Muskingum routing."""
    n = 72
    base = synth.write_blended(str(tmp_path), n_sb=2000, hru_per_sb=20, n_days=3, seed=3, routing="ROUTE_MUSKINGUM", catchment_route="TRIANGULAR_UH",
                               season_offset=124, n_gauges=9)
    rvi = open(base + ".rvi").read()
    open(base + ".rvi", "w").write(rvi.replace(":TimeStep              1.0", ":TimeStep              01:00:00"))
    model = api.RavenModel(base)
    model.flatten(1)
    assert model.nSteps == n
    sim = api.GpuSimulation(model)
    assert sim.kernel_variant() == "static:BLENDED"
    sim.set_recording(n)
    sim.run(n)
    ref = H.oracle_run(model, n, record_state=False)
    st = sim.download_state()[0]
    # strict: the device exp / log / pow are the host libm's bit for bit (csrc/rvn_glibc_math.h), so the melt-out knife edge of the
    # blended snow balance (`SWE <= 0` on a 1e-17 residual, src/SnowBalance.cpp:507) falls on the reference's side in every HRU
    assert H.rel_err(st, ref["state"]) < TOL
    assert H.rel_err(sim.hydrographs(0, n)[:, 0, :], ref["qout_rec"]) < TOL
    assert H.rel_err(sim.watershed(0, n)[:, 0, :3], ref["wshed_rec"][:, :3]) < TOL
    sim.close()
