"""CPU tests: the C oracle + the C++ host layer against (a) the committed golden fixtures (outputs of the unmodified
reference, tests/golden/make_golden.py) and (b), when the reference tree is present, a live run of the reference on
its own Salmon inputs.  Integer tables (state-variable catalogue, routing order, connections) must match exactly;
FP64 results within 1e-9 relative (they are in fact bit-identical on the build host)."""
import glob
import os

import numpy as np
import pytest

import helpers as H
from ravenhydroframework_b200 import api

TOL = 1e-9
CASES = sorted(os.path.basename(os.path.dirname(p)) for p in glob.glob(os.path.join(H.GOLDEN, "*", "reference.npz")))


def _load(case):
    g = np.load(os.path.join(H.GOLDEN, case, "reference.npz"))
    model = api.RavenModel(os.path.join(H.GOLDEN, case, "model"))
    model.flatten(1)
    return g, model


@pytest.mark.parametrize("case", CASES)
def test_integer_tables_match_reference(case):
    g, model = _load(case)
    import ctypes as C
    lib = model.lib
    # state-variable catalogue: types (reference enum numbering) and layers
    names = model.sv_names()
    assert model.NS == len(g["sv_type"]) == len(names)
    info = (C.c_int * 8)()
    lib.rvh_desc_info(model.desc, info)
    lib.rvh_desc_sv_table.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    t = np.zeros(model.NS, dtype=np.int32); l = np.zeros(model.NS, dtype=np.int32)
    lib.rvh_desc_sv_table(model.desc, t.ctypes.data, l.ctypes.data)
    assert np.array_equal(t, g["sv_type"])
    assert np.array_equal(l, g["sv_layer"])
    # routing order
    lib.rvh_desc_order.argtypes = [C.c_void_p, C.c_void_p]
    o = np.zeros(model.nSB, dtype=np.int32)
    lib.rvh_desc_order(model.desc, o.ctypes.data)
    assert np.array_equal(o, g["order"])
    # process connection tables (reference dump lists nconn and the first <= 8 connections of every process)
    procs = model.processes()
    assert len(procs) == len(g["proc"])
    for (name, nconn, frm, to), ref in zip(procs, g["proc"]):
        f = str(ref).split()
        assert int(f[2]) == nconn, (name, f)
        conns = ["%d>%d" % (a, b) for a, b in zip(frm, to)]
        assert conns == f[3:3 + len(conns)], (name, conns, f)


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_golden(case):
    g, model = _load(case)
    n = g["qout"].shape[0]
    o = H.oracle_run(model, n, record_forcings=True)
    steps = g["state_steps"]
    assert H.rel_err(o["state_rec"][steps - 1], g["state"]) < TOL
    assert H.rel_err(o["qout_rec"], g["qout"]) < TOL
    assert H.rel_err(o["wshed_rec"][:, :3], g["wshed"]) < TOL
    if "res_stage" in g:   # reservoirs: stage and losses after the last step (the last recorded step)
        _, sb = model.reservoir_state()
        assert steps[-1] == n and H.rel_err(o["res"][:, 0], g["res_stage"][-1][sb]) < TOL and H.rel_err(o["res"][:, 5], g["res_losses"][-1][sb]) < TOL
    # derived forcings the oracle keeps (subset of the reference force_struct; indices from src/RavenInclude.h:1294-1343)
    ref_f = g["forc"]
    idx = {"PRECIP": 0, "SNOW_FRAC": 3, "TEMP_AVE": 7, "TEMP_DAILY_MIN": 8, "TEMP_DAILY_MAX": 9, "TEMP_DAILY_AVE": 10,
           "POTENTIAL_MELT": 32, "PET": 34, "OW_PET": 35}
    if case.startswith("hmets") or case.startswith("blended"):
        idx.pop("OW_PET")   # :OW_Evaporation defaults to PET_HARGREAVES_1985 there; no HMETS process consumes OW_PET (not derived)
    if "_atm_" in case:   # PET from air pressure / relative humidity / shortwave radiation: the atmospheric chain is derived too
        idx.update({"AIR_PRES": 18, "REL_HUMIDITY": 19, "SW_RADIA": 24, "WIND_VEL": 33})
    if case in ("hbv_atm_priestley", "hbv_atm_penmancomb", "gr4j_atm_potmelt_eb", "hbv_atm_potmelt_restricted"):   # net radiation consumed by the PET method
        idx.update({"SW_RADIA_NET": 25, "LW_RADIA_NET": 28})
    for name, j in idx.items():
        mine = o["forc_rec"][steps - 1][:, :, api.FORCING_NAMES.index(name)]
        assert H.rel_err(mine, ref_f[:, :, j]) < TOL, name


@pytest.mark.parametrize("case", CASES)
def test_parameters_match_reference(case):
    """Class parameters as the reference resolved them ([DEFAULT] rows, autogeneration) vs the host layer's tables."""
    g, model = _load(case)
    import ctypes as C
    lib = model.lib
    lib.rvh_model_param_value.restype = C.c_double
    lib.rvh_model_param_value.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_char_p]
    checked = 0
    for row in g["param"]:
        kind, k, name, val = str(row).split()
        val = float(val)
        if val < -60000:      # reference marks parameters the model does not use
            continue
        if name in ("VAR_CAPACITY", "VAR_SNOW_CAPACITY") or (name == "THICKNESS"):
            continue
        mine = lib.rvh_model_param_value(model.h, kind.encode(), int(k), name.encode())
        if mine < -30000 and kind == "VEG" and name in ("MAX_CAPACITY", "MAX_SNOW_CAPACITY"):
            continue
        assert abs(mine - val) <= 1e-12 * max(1.0, abs(val)), (kind, k, name, mine, val)
        checked += 1
    assert checked > 10


@pytest.mark.skipif(not (H.have_reference() and os.path.isdir(H.REFERENCE_INPUTS)), reason="reference tree not present (GPU box)")
def test_oracle_matches_live_reference_salmon_hmets(tmp_path):
    """The reference's own Salmon_HMETS inputs (benchmarking/_InputFiles), first 3000 daily steps, run by the unmodified
    reference here and by host layer + oracle."""
    base = os.path.join(H.REFERENCE_INPUTS, "Salmon_HMETS", "raven-hmets-salmon")
    n = 3000
    ref = H.reference_run(base, str(tmp_path / "out"), n)
    model = api.RavenModel(base, duration=n)
    model.flatten(1)
    o = H.oracle_run(model, n)
    assert H.rel_err(o["state_rec"], ref["state"][1:n + 1]) < TOL
    assert H.rel_err(o["qout_rec"], ref["basin"][1:n + 1, :, 0]) < TOL
    assert H.rel_err(o["wshed_rec"][:, :2], ref["wshed"][1:n + 1, :2]) < TOL


@pytest.mark.skipif(not (H.have_reference() and os.path.isdir(H.REFERENCE_INPUTS)), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("rel", ["Salmon_HBV/raven-hbv-salmon", "Salmon_GR4J/raven-gr4j-salmon"])
def test_oracle_matches_live_reference_salmon_hbv_gr4j(tmp_path, rel):
    """The reference's own Salmon_HBV (HBV-EC) and Salmon_GR4J (BASELINE config 1) inputs as shipped, first 2000 daily steps:
    unmodified reference vs host layer + oracle, HRU state, hydrograph, balance and the derived forcings (incl. the
    Hargreaves-1985 open-water PET from extraterrestrial radiation that GR4J consumes)."""
    base = os.path.join(H.REFERENCE_INPUTS, rel)
    n = 2000
    ref = H.reference_run(base, str(tmp_path / "out"), n)
    model = api.RavenModel(base, duration=n)
    model.flatten(1)
    o = H.oracle_run(model, n, record_forcings=True)
    assert H.rel_err(o["state_rec"], ref["state"][1:n + 1]) < TOL
    assert H.rel_err(o["qout_rec"], ref["basin"][1:n + 1, :, 0]) < TOL
    assert H.rel_err(o["wshed_rec"][:, :2], ref["wshed"][1:n + 1, :2]) < TOL
    for name, j in {"PRECIP": 0, "SNOW_FRAC": 3, "TEMP_AVE": 7, "POTENTIAL_MELT": 32, "PET": 34, "OW_PET": 35}.items():
        assert H.rel_err(o["forc_rec"][:, :, api.FORCING_NAMES.index(name)], ref["forc"][1:n + 1, :, j]) < TOL, name


def test_mass_balance_closes():
    """WatershedStorage.csv 'MB Error' column: (S - S0) + (CumulOutput - CumulInput) stays at round-off."""
    g, model = _load("hmets_tree")
    n = 400
    o = H.oracle_run(model, n)
    st0 = model.initial_state()
    # initial water: area-weighted SOIL stores (everything else starts empty)
    w = o["wshed_rec"]
    names = model.sv_names()
    import ctypes as C
    area = np.zeros(model.nHRU)
    model.lib.rvh_desc_hru_area.argtypes = [C.c_void_p, C.c_void_p]
    model.lib.rvh_desc_hru_area(model.desc, area.ctypes.data)
    water = [i for i, nm in enumerate(names) if nm.split("[")[0] in ("SURFACE_WATER", "ATMOSPHERE", "PONDED_WATER", "SOIL", "SNOW", "SNOW_LIQ", "CONVOLUTION")]
    s0 = float((st0[:, water].sum(axis=1) * area).sum() / area.sum())
    # + initial channel and rivulet storage [m3 -> mm] (CModel::CalculateInitialWaterStorage, src/ModelInitialize.cpp:412-447)
    _, _, _, scal = model.basin_state()
    s0 += float(scal[:, 4].sum() + scal[:, 5].sum()) / (area.sum() * 1e6) * 1000.0
    mb = (w[:, 3] - s0) + (w[:, 1] - w[:, 0])
    assert np.max(np.abs(mb)) < 1e-8


def test_monte_carlo_member_matches_reference_golden():
    """CMonteCarloEnsemble on the host (libc rand() replay, seed offset by the start date, uniform/normal/gamma draws,
    CModel::UpdateParameter) + oracle vs member 2 of the reference's own ensemble loop."""
    d = os.path.join(H.GOLDEN, "hmets_mc")
    g = np.load(os.path.join(d, "reference_member2.npz"))
    model = api.RavenModel(os.path.join(d, "model"))
    assert model.ensemble_mode == 1 and model.num_members == 4
    model.flatten(4)
    assert model.ensemble_begin() == 4
    samples = []
    for e in range(4):
        samples.append(model.ensemble_update_model(e))
        model.flatten_member(e)
    samples = np.array(samples)
    assert samples.shape == (4, 19) and len(np.unique(samples[:, 0])) == 4
    n = g["qout"].shape[0]
    o = H.oracle_run(model, n, member=2)
    steps = g["state_steps"]
    assert H.rel_err(o["state_rec"][steps - 1], g["state"]) < TOL
    assert H.rel_err(o["qout_rec"], g["qout"]) < TOL
    other = H.oracle_run(model, n, member=1)
    assert H.rel_err(other["qout_rec"], g["qout"]) > 1e-3     # members really differ


HYDRO_CASES = sorted(os.path.basename(os.path.dirname(p)) for p in glob.glob(os.path.join(H.GOLDEN, "*", "reference_Hydrographs.csv")))


@pytest.mark.parametrize("case", HYDRO_CASES)
def test_hydrographs_csv_equals_the_reference_executables(case, tmp_path):
    """Output path: Hydrographs.csv written by the host layer from per-step records (here the oracle's) is the same TEXT the reference
    executable writes for the same inputs (step-averaged flows, observed columns with blanks, 6 significant digits, date / hour stamps)."""
    _, model = _load(case)
    ref = open(os.path.join(H.GOLDEN, case, "reference_Hydrographs.csv")).read().split("\n")
    n = len([l for l in ref if l.strip()]) - 2   # header + initial row
    o = H.oracle_run(model, n)
    out = str(tmp_path / "Hydrographs.csv")
    model.write_hydrographs(out, o["qout_rec"], o["wshed_rec"][:, 2])
    assert open(out).read().split("\n") == ref


FLUX_CASES = sorted(os.path.basename(os.path.dirname(p)) for p in glob.glob(os.path.join(H.GOLDEN, "*", "reference_cumflux.npz")))


@pytest.mark.parametrize("case", FLUX_CASES)
def test_cumulative_connection_fluxes_match_reference(case):
    """SURVEY 8a row a11, the optional arrays: per-connection cumulative fluxes (CModel::IncrementBalance, src/Model.cpp:2428-2436) of the oracle,
    summed the way CModel::GetCumulativeFlux does (src/Model.cpp:902-923), against the reference's GetCumulativeFlux for every HRU and state variable."""
    g = np.load(os.path.join(H.GOLDEN, case, "reference_cumflux.npz"))
    model = api.RavenModel(os.path.join(H.GOLDEN, case, "model"))
    model.flatten(1)
    cum = H.oracle_flux_balance(model, int(g["nsteps"]))
    assert np.abs(cum).sum() > 0.0
    assert H.rel_err(H.flux_by_state_variable(model, cum), g["cumflux"]) < TOL
