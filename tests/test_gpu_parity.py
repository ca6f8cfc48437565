"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.
Tolerance: 1e-9 relative (absolute floor 1e-9 mm / m3/s for stores that pass through zero), FP64 (north_star)."""
import numpy as np
import pytest

import helpers as H
from ravenhydroframework_b200 import api, synth

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _run_pair(tmp_path, n_sb, hru_per_sb, n_days, seed, members=1, **kw):
    base = synth.write_hmets(str(tmp_path), n_sb=n_sb, hru_per_sb=hru_per_sb, n_days=n_days, seed=seed, **kw)
    model = api.RavenModel(base)
    model.flatten(members)
    sim = api.GpuSimulation(model, n_members=members)
    sim.set_recording(n_days)
    sim.enable_forcing_output(True)
    sim.run(n_days)
    return model, sim


def test_hmets_single_hru_two_years(tmp_path):
    model, sim = _run_pair(tmp_path, 1, 1, 730, 5, single_class=True)
    ref = H.oracle_run(model, 730, record_forcings=True)
    assert H.rel_err(sim.download_state()[0], ref["state"]) < TOL
    assert H.rel_err(sim.hydrographs(0, 730)[:, 0, :], ref["qout_rec"]) < TOL
    assert H.rel_err(sim.forcings()[0], ref["forc_rec"][-1]) < TOL
    w = sim.watershed(0, 730)[:, 0, :]
    assert H.rel_err(w[:, :3], ref["wshed_rec"][:, :3]) < TOL
    assert H.rel_err(w[:, 3], ref["wshed_rec"][:, 3]) < 1e-9


def test_hmets_tree_ragged_block(tmp_path):
    # 31 subbasins x 13 HRUs = 403 HRUs: not a multiple of the 128-thread CTA nor of the 32-HRU padding
    model, sim = _run_pair(tmp_path, 31, 13, 400, 7)
    ref = H.oracle_run(model, 400)
    assert H.rel_err(sim.download_state()[0], ref["state"]) < TOL
    assert H.rel_err(sim.hydrographs(0, 400)[:, 0, :], ref["qout_rec"]) < TOL
    qin, qlat, qout, scal = sim.download_basin_state()
    assert H.rel_err(scal[0], ref["scal"]) < TOL
    assert H.rel_err(sim.cumulative()[0], ref["cumul"]) < TOL
    # mass balance closure (MB Error column of WatershedStorage.csv): (S - S0) + (out - in) ~ 0
    w = sim.watershed(0, 400)[:, 0, :]
    assert np.all(np.isfinite(w))


def test_members_are_independent_and_identical(tmp_path):
    model, sim = _run_pair(tmp_path, 7, 9, 120, 3, members=4)
    st = sim.download_state()
    for e in range(1, 4):
        assert np.array_equal(st[0], st[e])
    ref = H.oracle_run(model, 120)
    assert H.rel_err(st[3], ref["state"]) < TOL


def test_member_parameters_differ(tmp_path):
    base = synth.write_hmets(str(tmp_path), n_sb=3, hru_per_sb=5, n_days=200, seed=9)
    model = api.RavenModel(base)
    model.flatten(3)
    sim = api.GpuSimulation(model, n_members=3)
    # member 1: different gamma shape + melt factor, member 2: different soil parameters (CModel::UpdateParameter)
    model.update_parameter(2, "GAMMA_SHAPE", "FOREST", 7.7)
    model.update_parameter(2, "MAX_MELT_FACTOR", "OPEN", 5.5)
    sim.upload_member_params(1)
    model.update_parameter(0, "BASEFLOW_COEFF", "TOPSOIL", 0.031)
    model.update_parameter(4, "SNOW_SWI_MAX", "", 0.21)
    sim.upload_member_params(2)
    sim.run(200)
    st = sim.download_state()
    assert not np.array_equal(st[0], st[1]) and not np.array_equal(st[1], st[2])
    for e in (1, 2):
        ref = H.oracle_run(model, 200, member=e)
        assert H.rel_err(st[e], ref["state"]) < TOL


def test_stepwise_equals_batched(tmp_path):
    base = synth.write_hmets(str(tmp_path), n_sb=3, hru_per_sb=4, n_days=60, seed=2)
    model = api.RavenModel(base)
    model.flatten(1)
    a = api.GpuSimulation(model)
    a.run(60)
    b = api.GpuSimulation(model)
    for _ in range(60):
        b.run(1)
    assert np.array_equal(a.download_state(), b.download_state())


def test_streaming_steps_equal_batched(tmp_path):
    """rvn_gpu_step_async (forcing slab in, outflows + watershed record out, nothing waits) gives the bits of the batched run."""
    import torch
    model, sim = _run_pair(tmp_path, 7, 9, 60, 3, members=3)
    q_ref, w_ref, st_ref = sim.hydrographs(0, 60), sim.watershed(0, 60), sim.download_state()
    import ctypes as C
    model.lib.rvh_desc_gauge_series.restype = C.c_void_p
    model.lib.rvh_desc_gauge_series.argtypes = [C.c_void_p]
    nG = model.nGauges
    series = np.ctypeslib.as_array((C.c_double * (9 * nG * model.nSteps)).from_address(model.lib.rvh_desc_gauge_series(model.desc))).reshape(9, nG, model.nSteps)
    slabs = torch.from_numpy(np.ascontiguousarray(series.transpose(2, 1, 0))).pin_memory()   # per step: [nGauges][RVN_GF_COUNT]
    b = api.GpuSimulation(model, n_members=3)
    b.upload_forcings(np.zeros((9, nG, model.nSteps)), 0, model.nSteps)        # the device copy is wiped: every step must bring its own slab
    q = torch.zeros((60, 3, model.nSB), dtype=torch.float64).pin_memory()
    w = torch.zeros((60, 3, 4), dtype=torch.float64).pin_memory()
    for n in range(60):
        b.step_async(slabs.data_ptr() + n * 9 * nG * 8, q.data_ptr() + n * 3 * model.nSB * 8, w.data_ptr() + n * 3 * 4 * 8)
    b.sync()
    assert np.array_equal(q.numpy(), q_ref) and np.array_equal(w.numpy(), w_ref)
    assert np.array_equal(b.download_state(), st_ref)


def test_state_roundtrip(tmp_path):
    base = synth.write_hmets(str(tmp_path), n_sb=2, hru_per_sb=3, n_days=10, seed=2)
    model = api.RavenModel(base)
    model.flatten(2)
    sim = api.GpuSimulation(model, n_members=2, upload_initial=False)
    rng = np.random.RandomState(0)
    st = rng.uniform(0, 100, (2, model.nHRU, model.NS))
    sim.upload_state(st, 0, 2)
    assert np.array_equal(sim.download_state(), st)


# ---- the CUDA path against the committed golden fixtures (outputs of the unmodified reference) ---------------------
import glob
import os

GOLDEN_CASES = sorted(os.path.basename(os.path.dirname(p)) for p in glob.glob(os.path.join(H.GOLDEN, "*", "reference.npz")))


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_gpu_matches_reference_golden(case):
    g = np.load(os.path.join(H.GOLDEN, case, "reference.npz"))
    model = api.RavenModel(os.path.join(H.GOLDEN, case, "model"))
    model.flatten(2)
    n = g["qout"].shape[0]
    sim = api.GpuSimulation(model, n_members=2)
    sim.set_recording(n)
    sim.enable_forcing_output(True)
    steps = list(g["state_steps"])
    got = []
    res = []
    for s in steps:                      # stop at every recorded step and read the state back
        sim.run(int(s) - sim.step)
        got.append(sim.download_state()[1])
        res.append(sim.download_reservoir_state()[1])
    assert H.rel_err(np.array(got), g["state"]) < TOL
    if "res_stage" in g:                 # reservoirs: stage and mass-balance losses of every reservoir at the recorded steps
        _, sb = model.reservoir_state()
        assert H.rel_err(np.array(res)[:, :, 0], g["res_stage"][:, sb]) < TOL
        assert H.rel_err(np.array(res)[:, :, 5], g["res_losses"][:, sb]) < TOL
    if "_atm_" in case:                  # derived atmospheric forcings of the last step (reference force_struct columns)
        f = sim.forcings()[1]
        cols = {"AIR_PRES": 18, "REL_HUMIDITY": 19, "SW_RADIA": 24, "WIND_VEL": 33, "PET": 34, "OW_PET": 35}
        if case in ("hbv_atm_priestley", "hbv_atm_penmancomb", "gr4j_atm_potmelt_eb", "hbv_atm_potmelt_restricted"):   # net radiation and the melt on it
            cols.update({"SW_RADIA_NET": 25, "LW_RADIA_NET": 28, "POTENTIAL_MELT": 32})
        for name, j in cols.items():
            assert H.rel_err(f[:, api.FORCING_NAMES.index(name)], g["forc"][-1][:, j]) < TOL, name
    assert H.rel_err(sim.hydrographs(0, n)[:, 1, :], g["qout"]) < TOL
    assert H.rel_err(sim.watershed(0, n)[:, 1, :3], g["wshed"]) < TOL


@pytest.mark.parametrize("case", ["hbv_res_lake", "hbv_res_tables", "hbv_assim", "hbv_assim_late", "hbv_diffusive_vary", "hbv_chanprofile_hydrologic"])
def test_level_routing_path_matches_reference_golden(case):
    """Ensembles of more than 16 members route with one launch per network level (route_level_kernel, da_propagate_kernel, ...) instead
    of the one-CTA-per-member kernel the two-member golden runs take: the same fixtures through that path."""
    g = np.load(os.path.join(H.GOLDEN, case, "reference.npz"))
    model = api.RavenModel(os.path.join(H.GOLDEN, case, "model"))
    E = 20
    model.flatten(E)
    n = g["qout"].shape[0]
    sim = api.GpuSimulation(model, n_members=E)
    sim.set_recording(n)
    sim.run(n)
    assert H.rel_err(sim.download_state()[E - 1], g["state"][-1]) < TOL
    q = sim.hydrographs(0, n)
    assert H.rel_err(q[:, E - 1, :], g["qout"]) < TOL and H.rel_err(q[:, 7, :], g["qout"]) < TOL
    assert H.rel_err(sim.watershed(0, n)[:, E - 1, :3], g["wshed"]) < TOL
    if "res_stage" in g:
        _, sb = model.reservoir_state()
        assert H.rel_err(sim.download_reservoir_state()[E - 1][:, 0], g["res_stage"][-1][sb]) < TOL


@pytest.mark.parametrize("case", sorted(os.path.basename(os.path.dirname(p)) for p in glob.glob(os.path.join(H.GOLDEN, "*", "reference_Hydrographs.csv"))))
def test_gpu_hydrographs_csv_equals_the_reference_executables(case, tmp_path):
    """Hydrographs.csv formatted from the device's per-step records = the text file of the reference executable."""
    model = api.RavenModel(os.path.join(H.GOLDEN, case, "model"))
    model.flatten(1)
    ref = open(os.path.join(H.GOLDEN, case, "reference_Hydrographs.csv")).read().split("\n")
    n = len([l for l in ref if l.strip()]) - 2
    sim = api.GpuSimulation(model)
    sim.set_recording(n)
    sim.run(n)
    out = str(tmp_path / "Hydrographs.csv")
    model.write_hydrographs(out, sim.hydrographs(0, n)[:, 0, :], sim.watershed(0, n)[:, 0, 2])
    assert open(out).read().split("\n") == ref


def test_state_upload_mid_run_invalidates_cached_sums(tmp_path):
    """Replacing the state from the host (EnKF update, restart) must give the same result as a fresh simulation."""
    base = synth.write_hmets(str(tmp_path), n_sb=3, hru_per_sb=5, n_days=260, seed=12, season_offset=100)
    model = api.RavenModel(base)
    model.flatten(1)
    a = api.GpuSimulation(model)
    a.run(200)
    st = a.download_state()
    st[:, :, 13:40] *= 0.5               # perturb convolution stores behind the engine's back
    a.upload_state(st, 0, 1)
    a.run(60)
    ref = H.oracle_run(model, 200)
    # continue the oracle from the perturbed state: rerun 60 steps starting at step 200
    lib = H.oracle_lib()
    s2 = st[0].copy(); cumul = ref["cumul"].copy()
    qin, qlat, qout, scal = ref["qin"].copy(), ref["qlat"].copy(), ref["qout"].copy(), ref["scal"].copy()
    rc = lib.rvo_run_member(model.desc, 0, s2.ctypes.data, qin.ctypes.data, qlat.ctypes.data, qout.ctypes.data, scal.ctypes.data, cumul.ctypes.data,
                            200, 60, None, None, None, None)
    assert rc == 0
    assert H.rel_err(a.download_state()[0], s2) < TOL


# ---- HBV-EC / GR4J templates, channel routing methods, kernel variants ---------------------------------------------
WRITERS = {"hmets": synth.write_hmets, "hbv": synth.write_hbv, "gr4j": synth.write_gr4j, "blended": synth.write_blended}


def _gpu_vs_oracle(tmp_path, kind, n_days, members=1, **kw):
    base = WRITERS[kind](str(tmp_path), n_days=n_days, **kw)
    model = api.RavenModel(base)
    model.flatten(members)
    sim = api.GpuSimulation(model, n_members=members)
    sim.set_recording(n_days)
    sim.enable_forcing_output(True)
    sim.run(n_days)
    ref = H.oracle_run(model, n_days, record_forcings=True)
    e = members - 1
    assert H.rel_err(sim.download_state()[e], ref["state"]) < TOL
    assert H.rel_err(sim.hydrographs(0, n_days)[:, e, :], ref["qout_rec"]) < TOL
    assert H.rel_err(sim.forcings()[e], ref["forc_rec"][-1]) < TOL
    w = sim.watershed(0, n_days)[:, e, :]
    assert H.rel_err(w[:, :3], ref["wshed_rec"][:, :3]) < TOL
    assert H.rel_err(w[:, 3], ref["wshed_rec"][:, 3]) < TOL
    qin, qlat, qout, scal = sim.download_basin_state()
    assert H.rel_err(scal[e], ref["scal"]) < TOL
    return sim


@pytest.mark.parametrize("routing,catch", [("ROUTE_MUSKINGUM", "TRIANGULAR_UH"), ("ROUTE_MUSKINGUM_CUNGE", "ROUTE_DUMP"),
                                           ("ROUTE_STORAGECOEFF", "ROUTE_GAMMA_CONVOLUTION"), ("ROUTE_PLUG_FLOW", "TRIANGULAR_UH"),
                                           ("ROUTE_NONE", "TRIANGULAR_UH")])
def test_hbv_routing_methods(tmp_path, routing, catch):
    """HBV-EC multi-class watershed with glacier HRUs (BASELINE config 2 process list) over every channel routing method."""
    sim = _gpu_vs_oracle(tmp_path, "hbv", 300, n_sb=15, hru_per_sb=9, seed=21, routing=routing, catchment_route=catch)
    assert sim.kernel_variant() == "static:HBVEC"


def test_hbv_ragged_members(tmp_path):
    _gpu_vs_oracle(tmp_path, "hbv", 200, members=3, n_sb=31, hru_per_sb=13, seed=5, routing="ROUTE_MUSKINGUM")


def test_gr4j_tree(tmp_path):
    sim = _gpu_vs_oracle(tmp_path, "gr4j", 400, members=2, n_sb=7, hru_per_sb=19, seed=13)
    assert sim.kernel_variant() == "static:GR4J"


@pytest.mark.parametrize("kind", ["hmets", "hbv", "gr4j", "blended"])
def test_specialised_kernel_equals_interpreter(tmp_path, kind):
    """The compile-time specialisation and the interpreter run the same process functions: results must be bit-identical."""
    base = WRITERS[kind](str(tmp_path), n_sb=5, hru_per_sb=31, n_days=150, seed=17)
    model = api.RavenModel(base)
    model.flatten(2)
    a = api.GpuSimulation(model, n_members=2)
    assert a.kernel_variant().startswith("static:")
    a.set_recording(150)
    a.run(150)
    b = api.GpuSimulation(model, n_members=2, force_interpreter=True)
    assert b.kernel_variant() == "interpreter"
    b.set_recording(150)
    b.run(150)
    assert np.array_equal(a.download_state(), b.download_state())
    assert np.array_equal(a.hydrographs(0, 150), b.hydrographs(0, 150))
    assert np.array_equal(a.watershed(0, 150), b.watershed(0, 150))


def test_iterated_heun_many_steps_equals_oracle(tmp_path):
    """:Method ITER_HEUN over a whole season: the reference executable aborts after its first step (tests/golden/hbv_heun_* pin that
    step on it), the oracle restates src/Solvers.cpp:304-397 and the interpreter sweep must follow it bit for bit."""
    base = synth.write_hbv(str(tmp_path), n_sb=5, hru_per_sb=13, n_days=200, seed=21, routing="ROUTE_MUSKINGUM", season_offset=60)
    rvi = open(base + ".rvi").read()
    assert ":Method                ORDERED_SERIES" in rvi
    open(base + ".rvi", "w").write(rvi.replace(":Method                ORDERED_SERIES", ":Method                ITER_HEUN 0.002 12"))
    model = api.RavenModel(base)
    model.flatten(2)
    sim = api.GpuSimulation(model, n_members=2)
    assert sim.kernel_variant() == "interpreter"
    sim.set_recording(200)
    sim.run(200)
    ref = H.oracle_run(model, 200)
    st = sim.download_state()
    assert H.rel_err(st[0], ref["state"]) < TOL and np.array_equal(st[0], st[1])
    assert H.rel_err(sim.hydrographs(0, 200)[:, 0, :], ref["qout_rec"]) < TOL
    assert H.rel_err(sim.watershed(0, 200)[:, 0, :3], ref["wshed_rec"][:, :3]) < TOL


def test_blended_weights_are_runtime_data(tmp_path):
    """Different :EndProcessGroup weights keep the specialised Blended kernel (weights are not part of the compiled program) and
    change the result; both runs equal the oracle."""
    res = []
    for i, r in enumerate([(0.35, 0.6, 0.2, 0.7, 0.45, 0.8, 0.3, 0.55), (0.9, 0.1, 0.5, 0.5, 0.05, 0.3, 0.75, 0.2)]):
        base = synth.write_blended(str(tmp_path / ("w%d" % i)), n_sb=3, hru_per_sb=9, n_days=200, seed=9, season_offset=80, r=r)
        model = api.RavenModel(base)
        model.flatten(1)
        sim = api.GpuSimulation(model)
        assert sim.kernel_variant() == "static:BLENDED"
        sim.run(200)
        ref = H.oracle_run(model, 200)
        st = sim.download_state()[0]
        assert H.rel_err(st, ref["state"]) < TOL
        res.append(st)
    assert not np.array_equal(res[0], res[1])


def test_non_template_process_list_runs_on_interpreter(tmp_path):
    """A process list that is not one of the compiled templates (HMETS without the second baseflow) must still run, on the interpreter."""
    base = synth.write_hmets(str(tmp_path), n_sb=3, hru_per_sb=7, n_days=120, seed=4)
    rvi = open(base + ".rvi").read().replace("  :Baseflow        BASE_LINEAR     SOIL[1]        SURFACE_WATER\n", "")
    open(base + ".rvi", "w").write(rvi)
    model = api.RavenModel(base)
    model.flatten(1)
    sim = api.GpuSimulation(model)
    assert sim.kernel_variant() == "interpreter"
    sim.run(120)
    ref = H.oracle_run(model, 120)
    assert H.rel_err(sim.download_state()[0], ref["state"]) < TOL


def test_monte_carlo_ensemble_matches_reference_member_and_shards():
    """GPU ensemble (all members at once) vs member 2 of the reference's sequential member loop; a shard that starts at
    global member 2 (what rank 1 of 2 would hold) reproduces members 2..3 of the full run bit for bit."""
    d = os.path.join(H.GOLDEN, "hmets_mc")
    g = np.load(os.path.join(d, "reference_member2.npz"))
    n = g["qout"].shape[0]
    model = api.RavenModel(os.path.join(d, "model"))
    model.flatten(4)
    full = api.GpuSimulation(model, n_members=4)
    full.apply_ensemble(0)
    full.set_recording(n)
    full.run(n)
    assert H.rel_err(full.hydrographs(0, n)[:, 2, :], g["qout"]) < TOL
    st = full.download_state()
    model2 = api.RavenModel(os.path.join(d, "model"))
    model2.flatten(2)
    shard = api.GpuSimulation(model2, n_members=2)
    shard.apply_ensemble(2)
    shard.run(n)
    assert np.array_equal(shard.download_state(), st[2:4])


def test_sparse_station_interpolation(tmp_path):
    """K3: every HRU gathers its 4 nearest of 25 stations (CSR weights from an INTERP_FROM_FILE table) -- the sparse weighted
    sums of a gridded forcing -- with elevation-dependent corrections from the weighted station elevations."""
    _gpu_vs_oracle(tmp_path, "hbv", 120, members=2, n_sb=9, hru_per_sb=17, seed=31, routing="ROUTE_MUSKINGUM", n_gauges=25)
    _gpu_vs_oracle(tmp_path / "b", "hmets", 120, n_sb=3, hru_per_sb=50, seed=32, n_gauges=16)


FLUX_CASES = sorted(os.path.basename(os.path.dirname(p)) for p in glob.glob(os.path.join(H.GOLDEN, "*", "reference_cumflux.npz")))


@pytest.mark.gpu
@pytest.mark.parametrize("case", FLUX_CASES)
def test_gpu_flux_balance_matches_oracle_and_reference(case):
    """Optional per-connection cumulative flux arrays (rvn_options::keep_flux_balance, reference CModel::_aCumulativeBal): every connection of
    every HRU against the oracle, and summed by state variable against the reference's GetCumulativeFlux; a model created without the option
    refuses the download."""
    g = np.load(os.path.join(H.GOLDEN, case, "reference_cumflux.npz"))
    n = int(g["nsteps"])
    model = api.RavenModel(os.path.join(H.GOLDEN, case, "model"))
    model.set_flux_balance(True)
    model.flatten(2)
    want = H.oracle_flux_balance(model, n)
    sim = api.GpuSimulation(model, n_members=2)
    assert sim.kernel_variant() == "interpreter"
    sim.run(n)
    got = sim.download_flux_balance()
    assert H.rel_err(got[0], want) < TOL and np.array_equal(got[0], got[1])
    assert H.rel_err(H.flux_by_state_variable(model, got[1]), g["cumflux"]) < TOL
    sim.close()
    model.set_flux_balance(False)
    model.flatten(1)
    sim = api.GpuSimulation(model, n_members=1)
    sim.run(2)
    with pytest.raises(api.RavenError):
        sim.download_flux_balance()
    sim.close()
