"""Parity at the BASELINE.json sizes (GPU only).  The C oracle is fast enough to check ONE member of the full-size watershed
completely; the other members are covered by size-independent properties: members with identical parameters are bit-identical,
members with sampled parameters differ, and every member closes its water balance (the MB Error column of the reference's
WatershedStorage.csv: (S - S0) + (cumulative out - cumulative in) = 0)."""
import numpy as np
import pytest

import helpers as H
from ravenhydroframework_b200 import api, enkf, synth

pytestmark = pytest.mark.gpu
TOL = 1e-9


def test_config3_share_hmets_512_members_x_10k_hrus(tmp_path):
    """BASELINE config 3, the per-GPU share: HMETS, 10,000 HRUs / 500 subbasins, 512 members, Monte-Carlo parameters."""
    n = 1461   # four years of the configuration's daily steps: long-run drift of the convolution stores and the snow pack is part of the check
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
    # complete members against the oracle over the whole run (their own sampled parameters): every storage of every HRU at the end, every
    # hydrograph value and the watershed balance of every step
    for mem in (37, 0, 511):
        ref = H.oracle_run(model, n, member=mem, record_state=False)
        st = sim.download_state(mem, 1)[0]
        assert H.rel_err(st, ref["state"]) < TOL
        assert H.rel_err(sim.hydrographs(0, n, mem, 1)[:, 0, :], ref["qout_rec"]) < TOL
        assert H.rel_err(sim.watershed(0, n, mem, 1)[:, 0, :3], ref["wshed_rec"][:, :3]) < TOL
    # every member: finite, different from its neighbours, balance closed
    w = sim.watershed(0, n)                               # [n][E][4]
    assert np.all(np.isfinite(w))
    s_init = ref["wshed_rec"][0, 3] - (ref["wshed_rec"][0, 0] - ref["wshed_rec"][0, 1])   # S0 of a member = S1 - (in1 - out1); same initial state for all
    mb = (w[:, :, 3] - s_init) + (w[:, :, 1] - w[:, :, 0])
    assert np.abs(mb).max() < 1e-7 * max(1.0, np.abs(w[:, :, 3]).max()), np.abs(mb).max()
    q = sim.hydrographs(n - 1, 1)[:, :, 0]
    assert len(np.unique(q[-1])) > 500
    sim.close()
    assert st0.shape == (10000, model.NS)


def test_config2_hbvec_10k_hrus_500_subbasins_muskingum(tmp_path):
    """BASELINE config 2: HBV-EC semi-distributed, 10,000 HRUs / 500 subbasins, Muskingum routing -- complete comparison with the
    oracle (every HRU storage, every subbasin hydrograph), plus identical replicas."""
    n = 7305   # the configuration's full 20 years of daily steps
    base = synth.write_hbv(str(tmp_path), n_sb=500, hru_per_sb=20, n_days=n, seed=2, routing="ROUTE_MUSKINGUM", season_offset=100)
    model = api.RavenModel(base)
    model.flatten(3)
    sim = api.GpuSimulation(model, n_members=3)
    assert sim.kernel_variant() == "static:HBVEC"
    sim.set_recording(n)
    sim.run(n)
    ref = H.oracle_run(model, n, record_state=False)
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


def test_config4_full_size_1m_hrus_50k_subbasins_gridded_hourly(tmp_path):
    """BASELINE config 4 at its full size: Blended process groups, 1,000,000 HRUs / 50,000 subbasins, forcing on a 200 x 200 grid
    (40,000 cells, 4 cells per HRU, CSR gather), hourly step, Muskingum routing.  Every HRU storage, every subbasin hydrograph and
    the watershed balance against the oracle under the strict tolerance - no exception list."""
    n = 24
    base = synth.write_blended(str(tmp_path), n_sb=50000, hru_per_sb=20, n_days=1, seed=4, routing="ROUTE_MUSKINGUM", catchment_route="TRIANGULAR_UH",
                               season_offset=124)
    rvi = open(base + ".rvi").read()
    open(base + ".rvi", "w").write(rvi.replace(":TimeStep              1.0", ":TimeStep              01:00:00"))
    model = api.RavenModel(base)
    model.flatten(1)
    assert (model.nHRU, model.nSB, model.nSteps) == (1000000, 50000, n)
    assert synth.apply_cell_grid(model, base, 200, 200) == 40000
    model.flatten(1)
    sim = api.GpuSimulation(model)
    assert sim.kernel_variant() == "static:BLENDED"
    sim.set_recording(n)
    sim.run(n)
    ref = H.oracle_run(model, n, record_state=False)
    assert H.rel_err(sim.download_state()[0], ref["state"]) < TOL
    assert H.rel_err(sim.hydrographs(0, n)[:, 0, :], ref["qout_rec"]) < TOL
    assert H.rel_err(sim.watershed(0, n)[:, 0, :3], ref["wshed_rec"][:, :3]) < TOL
    sim.close()


def test_config5_enkf_512_members_x_100k_hrus(tmp_path):
    """BASELINE config 5 at its full size on one GPU: HBV-EC, 100,000 HRUs / 5,000 subbasins, 512 members with forcing perturbations,
    SOIL[0] + SNOW of every HRU and STREAMFLOW of every subbasin assimilated at 10 gauged outlets.  Forecast: two complete members against
    the oracle.  Analysis: Z from the host layer equals the oracle's (same Golub-Reinsch SVD), and the device update of a 4,096-row
    sample of the 310,000-row state matrix equals the oracle's analysis of those rows bit for bit (rows are independent given Z)."""
    import sys
    sys.path.insert(0, H.ROOT + "/oracle")
    import enkf_oracle as EO
    n_sb, hps, days, N = 5000, 20, 6, 512
    base = synth.write_hbv(str(tmp_path), n_sb=n_sb, hru_per_sb=hps, n_days=days, seed=31, routing="ROUTE_MUSKINGUM", name="model", season_offset=130)
    synth.add_enkf(base, n_members=N, n_sb=n_sb, n_hru=n_sb * hps, n_days=days, obs_basins=tuple(range(1, 11)), window=1)
    model = api.RavenModel(base)
    sim = enkf.EnKFSimulation(model, device=0)
    assert (sim.N, model.nHRU, model.nSB) == (N, 100000, 5000)
    out = sim.forecast()
    for e in (0, 377):
        r = H.oracle_run(model, sim.nsteps, member=e, record_state=False)
        assert H.rel_err(sim.sim.download_state(e, 1)[0], r["state"]) < TOL
        assert H.rel_err(sim.sim.hydrographs(0, sim.nsteps, e, 1)[:, 0, :], r["qout_rec"]) < TOL
    obs, noise, _, _ = model.enkf_observations()
    pre, post = sim.analysis()
    M = pre.shape[1]
    assert M == sim.info["M"] and M >= 300000
    rows = np.random.RandomState(5).choice(M, size=4096, replace=False)
    Xa, Zo = EO.assimilation_calcs(pre[:, rows], out, obs, noise)
    assert np.array_equal(model.enkf_compute_Z(out), Zo)
    assert np.array_equal(post[:, rows], Xa)
    assert np.array_equal(sim.sim.enkf_pack(), np.maximum(post, 0.0))   # UpdateFromStateMatrix
    sim.sim.close()
