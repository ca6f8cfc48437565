"""Synthetic Raven input sets (.rvi/.rvh/.rvp/.rvt/.rvc[/.rve]) for the BASELINE.json configurations.

The files are plain Raven inputs: the same set is read by this package's C++ host layer and, in the
parity tests, by the unmodified reference executable.  Watershed = heap-ordered binary tree of
subbasins (downstream of p is p//2, outlet p=1) with `hru_per_sb` HRUs each; HRU attributes and the
daily weather series are drawn from a seeded generator (no reference data is used).
"""
import math
import os

import numpy as np


def weather(n_days, seed=1, season_offset=0):
    """Daily RAINFALL, SNOWFALL [mm/d], TEMP_DAILY_MIN, TEMP_DAILY_MAX [C], PET [mm/d] for a cold-region climate."""
    rng = np.random.RandomState(seed)
    doy = (np.arange(n_days) + season_offset) % 365
    tmean = 2.5 + 13.0 * np.sin(2 * np.pi * (doy - 110) / 365.0)
    noise = np.zeros(n_days)
    eps = rng.normal(0, 3.0, n_days)
    for i in range(1, n_days):
        noise[i] = 0.7 * noise[i - 1] + eps[i]
    tmean = tmean + noise
    tmin = tmean - rng.uniform(2.5, 8.0, n_days)
    tmax = tmean + rng.uniform(2.5, 8.0, n_days)
    wet = rng.uniform(0, 1, n_days) < 0.42
    amount = rng.gamma(0.8, 6.0, n_days) * wet
    snow_frac = np.clip((1.5 - tmean) / 3.0, 0.0, 1.0)
    snow = amount * snow_frac
    rain = amount - snow
    pet = np.clip(0.18 * (tmean + 4.0), 0.0, None) * (0.6 + 0.4 * np.sin(2 * np.pi * (doy - 80) / 365.0) ** 2)
    return np.round(rain, 4), np.round(snow, 4), np.round(tmin, 4), np.round(tmax, 4), np.round(pet, 4)


HMETS_RVI = """# HMETS emulation, synthetic watershed (written by ravenhydroframework_b200.synth)
:StartDate             2000-01-01 00:00:00
:Duration              {duration}
:TimeStep              1.0
:Method                ORDERED_SERIES
:PotentialMeltMethod   POTMELT_HMETS
:RainSnowFraction      RAINSNOW_DATA
:Evaporation           PET_DATA
:CatchmentRoute        {catchment_route}
:Routing               {routing}
:SoilModel             SOIL_TWO_LAYER
:SilentMode
{extra}
:Alias DELAYED_RUNOFF CONVOLUTION[1]
:HydrologicProcesses
  :SnowBalance     SNOBAL_HMETS    MULTIPLE       MULTIPLE
  :Precipitation   RAVEN_DEFAULT   ATMOS_PRECIP   MULTIPLE
  :Infiltration    INF_HMETS       PONDED_WATER   MULTIPLE
    :Overflow      OVERFLOW_RAVEN  SOIL[0]        DELAYED_RUNOFF
  :Baseflow        BASE_LINEAR     SOIL[0]        SURFACE_WATER
  :Percolation     PERC_LINEAR     SOIL[0]        SOIL[1]
    :Overflow      OVERFLOW_RAVEN  SOIL[1]        DELAYED_RUNOFF
  :SoilEvaporation SOILEVAP_ALL    SOIL[0]        ATMOSPHERE
  :Convolve        CONVOL_GAMMA    CONVOLUTION[0] SURFACE_WATER
  :Convolve        CONVOL_GAMMA_2  DELAYED_RUNOFF SURFACE_WATER
  :Baseflow        BASE_LINEAR     SOIL[1]        SURFACE_WATER
:EndHydrologicProcesses
"""

HMETS_RVP = """# HMETS parameters (synthetic)
:SoilClasses
  :Attributes,
  :Units,
  TOPSOIL,
  PHREATIC,
:EndSoilClasses
:LandUseClasses,
  :Attributes,  IMPERM, FOREST_COV,
  :Units,         frac,       frac,
  FOREST,          0.0,        1.0,
  OPEN,           0.05,        0.1,
:EndLandUseClasses
:VegetationClasses,
  :Attributes,  MAX_HT, MAX_LAI, MAX_LEAF_COND,
  :Units,            m,    none,      mm_per_s,
  FOREST,            4,       5,             5,
  GRASS,           0.5,       2,             5,
:EndVegetationClasses
:SoilProfiles
  LAKE, 0
  ROCK, 0
  DEFAULT_P, 2, TOPSOIL, 0.3107, PHREATIC, 0.9162,
  THIN_P,    2, TOPSOIL, 0.2200, PHREATIC, 0.6100,
:EndSoilProfiles
:GlobalParameter     SNOW_SWI_MIN 0.0464
:GlobalParameter     SNOW_SWI_MAX 0.2462
:GlobalParameter SWI_REDUCT_COEFF 0.0222
:GlobalParameter         SNOW_SWI 0.05
:SoilParameterList
  :Parameters, POROSITY, PERC_COEFF, PET_CORRECTION, BASEFLOW_COEFF
  :Units,             -,        1/d,              -,            1/d
  TOPSOIL,          1.0,     0.0114,            1.0,         0.0243
  PHREATIC,         1.0,        0.0,            0.0,         0.0069
:EndSoilParameterList
:LandUseParameterList
  :Parameters, MIN_MELT_FACTOR, MAX_MELT_FACTOR, DD_MELT_TEMP, DD_AGGRADATION, REFREEZE_FACTOR, REFREEZE_EXP, DD_REFREEZE_TEMP, HMETS_RUNOFF_COEFF,
  :Units,               mm/d/C,          mm/d/C,            C,           1/mm,          mm/d/C,            -,                C,                  -,
  [DEFAULT],            1.2875,          6.7009,       2.3641,         0.0973,          2.6851,       0.3740,          -1.0919,             0.4739,
  OPEN,                 1.6000,          7.9000,       1.8000,         0.0700,          2.1000,       0.4500,          -0.5000,             0.3900,
:EndLandUseParameterList
:LandUseParameterList
  :Parameters, GAMMA_SHAPE, GAMMA_SCALE, GAMMA_SHAPE2, GAMMA_SCALE2,
  :Units,                -,         1/d,            -,          1/d,
  [DEFAULT],        9.5019,      0.2774,       6.3942,       0.6884,
  OPEN,             5.1000,      0.6100,       4.2000,       0.9500,
:EndLandUseParameterList
:VegetationParameterList
  :Parameters, RAIN_ICEPT_PCT, SNOW_ICEPT_PCT,
  :Units,                   -,              -,
  [DEFAULT],              0.0,            0.0,
  FOREST,                0.06,           0.04,
:EndVegetationParameterList
{extra}
"""

HMETS_RVC = """# half-full soil stores
:UniformInitialConditions SOIL[0] 155.36055
:UniformInitialConditions SOIL[1] 458.09735
"""

# the 19 HMETS parameters CModel::UpdateParameter can reach, uniform +-20 %
HMETS_DISTS = [
    ("GAMMA_SHAPE", "LANDUSE", "FOREST", 9.5019), ("GAMMA_SCALE", "LANDUSE", "FOREST", 0.2774),
    ("GAMMA_SHAPE2", "LANDUSE", "FOREST", 6.3942), ("GAMMA_SCALE2", "LANDUSE", "FOREST", 0.6884),
    ("MIN_MELT_FACTOR", "LANDUSE", "FOREST", 1.2875), ("MAX_MELT_FACTOR", "LANDUSE", "FOREST", 6.7009),
    ("DD_MELT_TEMP", "LANDUSE", "FOREST", 2.3641), ("DD_AGGRADATION", "LANDUSE", "FOREST", 0.0973),
    ("SNOW_SWI_MIN", "GLOBALS", "ALL", 0.0464), ("SNOW_SWI_MAX", "GLOBALS", "ALL", 0.2462),
    ("SWI_REDUCT_COEFF", "GLOBALS", "ALL", 0.0222), ("DD_REFREEZE_TEMP", "LANDUSE", "FOREST", -1.0919),
    ("REFREEZE_FACTOR", "LANDUSE", "FOREST", 2.6851), ("REFREEZE_EXP", "LANDUSE", "FOREST", 0.3740),
    ("PET_CORRECTION", "SOIL", "TOPSOIL", 1.0), ("HMETS_RUNOFF_COEFF", "LANDUSE", "FOREST", 0.4739),
    ("PERC_COEFF", "SOIL", "TOPSOIL", 0.0114), ("BASEFLOW_COEFF", "SOIL", "TOPSOIL", 0.0243),
    ("BASEFLOW_COEFF", "SOIL", "PHREATIC", 0.0069),
]


def _write(path, text):
    with open(path, "w") as f:
        f.write(text)


TOPOLOGY = "tree"   # "tree": heap-ordered binary tree down(p) = p // 2; "chain": down(p) = p - 1, as deep as the network can be


def write_network(f, n_sb, hru_per_sb, rng, profile, lu_classes, veg_classes, soil_profiles, gauged_all=False, paired_profiles=False,
                  terrain_classes=None):
    f.write(":SubBasins\n  :Attributes NAME DOWNSTREAM_ID PROFILE REACH_LENGTH GAUGED\n  :Units none none none km none\n")
    for p in range(1, n_sb + 1):
        down = -1 if p == 1 else (p // 2 if TOPOLOGY == "tree" else p - 1)
        reach = "_AUTO" if profile == "NONE" else "%.4f" % rng.uniform(5, 15)
        f.write("  %d, sb%d, %d, %s, %s, %d\n" % (p, p, down, profile, reach, 1 if (p == 1 or gauged_all) else 0))
    f.write(":EndSubBasins\n:HRUs\n  :Attributes AREA ELEVATION LATITUDE LONGITUDE BASIN_ID LAND_USE_CLASS VEG_CLASS SOIL_PROFILE AQUIFER_PROFILE TERRAIN_CLASS SLOPE ASPECT\n")
    f.write("  :Units km2 m deg deg none none none none none none deg deg\n")
    k = 1
    for p in range(1, n_sb + 1):
        for _ in range(hru_per_sb):
            c = rng.randint(0, len(lu_classes))
            f.write("  %d, %.4f, %.2f, %.5f, %.5f, %d, %s, %s, %s, [NONE], %s, %.3f, %.2f\n" % (
                k, rng.uniform(0.5, 2.5), rng.uniform(600, 1400), rng.uniform(54, 55), rng.uniform(-123.5, -122.5), p,
                lu_classes[c], veg_classes[c], soil_profiles[c if paired_profiles else rng.randint(0, len(soil_profiles))],
                "[NONE]" if not terrain_classes else terrain_classes[c % len(terrain_classes)],
                rng.uniform(0, 20), rng.uniform(0, 360)))
            k += 1
    f.write(":EndHRUs\n")


def write_gauge(f, rain, snow, tmin, tmax, pet, elev=843.0, lat=54.4848, lon=-123.3659, extra=""):
    n = len(rain)
    f.write(":Gauge met1\n  :Latitude %.4f\n  :Longitude %.4f\n  :Elevation %.1f\n%s" % (lat, lon, elev, extra))
    f.write("  :MultiData\n    2000-01-01 00:00:00 1 %d\n" % n)
    f.write("    :Parameters RAINFALL SNOWFALL TEMP_DAILY_MIN TEMP_DAILY_MAX PET\n    :Units mm/d mm/d C C mm/d\n")
    for i in range(n):
        f.write("    %.4f %.4f %.4f %.4f %.4f\n" % (rain[i], snow[i], tmin[i], tmax[i], pet[i]))
    f.write("  :EndMultiData\n:EndGauge\n")


def write_gauge_grid(base, n_gauges, n_days, seed, season_offset, hru_latlon, extra=""):
    """n_gauges stations on a regular lat/lon grid over the watershed, each with its own series (base weather x a smooth
    spatial factor, different elevations), + an INTERP_FROM_FILE weight table giving every HRU its 4 nearest stations with
    normalised inverse-distance weights: the same sparse weighted sums a gridded forcing (CForcingGrid::GetWeightedValue,
    src/ForcingGrid.cpp:1886-1946) produces, expressed through the gauge path the reference can run without NetCDF.
    Returns the .rvi line that selects the table."""
    rng = np.random.RandomState(seed + 1000)
    gx = int(math.ceil(math.sqrt(n_gauges)))
    rain, snow, tmin, tmax, pet = weather(n_days, seed, season_offset)
    lats, lons = [], []
    with open(base + ".rvt", "w") as f:
        for g in range(n_gauges):
            iy, ix = divmod(g, gx)
            lat = 54.0 + (iy + 0.5) / gx
            lon = -123.5 + (ix + 0.5) / gx
            lats.append(lat); lons.append(lon)
            fac = 1.0 + 0.25 * math.sin(1.3 * ix + 0.4) * math.cos(0.9 * iy - 0.2)
            dT = 2.0 * math.sin(0.7 * ix) - 1.5 * math.cos(1.1 * iy)
            elev = 700.0 + 500.0 * rng.uniform(0, 1)
            f.write(":Gauge met%d\n  :Latitude %.4f\n  :Longitude %.4f\n  :Elevation %.1f\n%s" % (g + 1, lat, lon, elev, extra))
            f.write("  :MultiData\n    2000-01-01 00:00:00 1 %d\n" % n_days)
            f.write("    :Parameters RAINFALL SNOWFALL TEMP_DAILY_MIN TEMP_DAILY_MAX PET\n    :Units mm/d mm/d C C mm/d\n")
            for i in range(n_days):
                f.write("    %.4f %.4f %.4f %.4f %.4f\n" % (rain[i] * fac, snow[i] * fac, tmin[i] + dT, tmax[i] + dT, pet[i] * (2.0 - fac)))
            f.write("  :EndMultiData\n:EndGauge\n")
    wfile = base + "_gauge_weights.txt"
    with open(wfile, "w") as f:
        f.write("%d %d\n" % (n_gauges, len(hru_latlon)))
        for (la, lo) in hru_latlon:
            d2 = [(la - a) ** 2 + (lo - b) ** 2 for a, b in zip(lats, lons)]
            near = sorted(np.argsort(d2)[:min(4, n_gauges)])
            raw = np.array([1.0 / (d2[g] + 1e-6) for g in near])
            w = raw / raw.sum()
            row = [0.0] * n_gauges
            for g, x in zip(near[:-1], w[:-1]):
                row[g] = float("%.12g" % x)
            row[near[-1]] = 1.0 - sum(row[g] for g in near[:-1])
            f.write(" ".join("%.17g" % x for x in row) + "\n")
    return ":InterpolationMethod INTERP_FROM_FILE %s\n" % os.path.basename(wfile)


def apply_cell_grid(model, base, nx=200, ny=200, seed=1):
    """BASELINE config 4's forcing input: a synthetic nx x ny grid over the watershed (values = the single-gauge series of the
    parsed model x a smooth spatial factor, temperatures + a smooth offset, elevations 700-1200 m), every HRU weighted to its 4 nearest
    cells with normalised inverse-distance weights (CSR, nnz = 4 nHRU) - handed to the host layer through RavenModel.set_cell_forcing
    (the arithmetic of CForcingGrid::GetWeightedValue, src/ForcingGrid.cpp:1886-1900).  `model` must be flattened (1 gauge)."""
    base_series = model.gauge_series()                 # [9][1][nSteps]
    assert base_series.shape[1] == 1
    nS = base_series.shape[2]
    ll = np.array(_hru_latlon(base + ".rvh"))          # [nH][2]
    la0, la1, lo0, lo1 = ll[:, 0].min(), ll[:, 0].max(), ll[:, 1].min(), ll[:, 1].max()
    dy, dx = (la1 - la0) / ny + 1e-12, (lo1 - lo0) / nx + 1e-12
    iy, ix = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    fac = (1.0 + 0.25 * np.sin(0.13 * ix + 0.4) * np.cos(0.09 * iy - 0.2)).reshape(-1)
    dT = (2.0 * np.sin(0.07 * ix) - 1.5 * np.cos(0.11 * iy)).reshape(-1)
    rng = np.random.RandomState(seed + 2000)
    elev = 700.0 + 500.0 * rng.uniform(0, 1, nx * ny)
    nC = nx * ny
    series = np.empty((9, nC, nS))
    GF = dict(PRECIP=0, SNOW_FRAC=1, TEMP_AVE=2, TEMP_DAILY_AVE=3, TEMP_DAILY_MIN=4, TEMP_DAILY_MAX=5, PET=6, OW_PET=7, POTENTIAL_MELT=8)
    for name, i in GF.items():
        b = base_series[i, 0][None, :]
        if name in ("PRECIP",):
            series[i] = b * fac[:, None]
        elif name in ("PET", "OW_PET"):
            series[i] = b * (2.0 - fac)[:, None]
        elif name.startswith("TEMP"):
            series[i] = b + dT[:, None]
        else:
            series[i] = np.broadcast_to(b, (nC, nS))
    # 4 nearest cell centres: the cell containing the HRU and its neighbours towards the HRU's side
    fy = (ll[:, 0] - la0) / dy - 0.5
    fx = (ll[:, 1] - lo0) / dx - 0.5
    y0 = np.clip(np.floor(fy).astype(int), 0, ny - 2)
    x0 = np.clip(np.floor(fx).astype(int), 0, nx - 2)
    nH = ll.shape[0]
    idx = np.stack([y0 * nx + x0, y0 * nx + x0 + 1, (y0 + 1) * nx + x0, (y0 + 1) * nx + x0 + 1], axis=1)   # ascending cell index
    cy = np.stack([y0, y0, y0 + 1, y0 + 1], axis=1) + 0.5
    cx = np.stack([x0, x0 + 1, x0, x0 + 1], axis=1) + 0.5
    d2 = (cy - (fy + 0.5)[:, None]) ** 2 + (cx - (fx + 0.5)[:, None]) ** 2
    w = 1.0 / (d2 + 1e-6)
    w = w / w.sum(axis=1, keepdims=True)
    w[:, 3] = 1.0 - w[:, :3].sum(axis=1)
    ptr = np.arange(nH + 1, dtype=np.int32) * 4
    model.set_cell_forcing(elev, series, ptr, idx.reshape(-1).astype(np.int32), w.reshape(-1))
    return nC


def _hru_latlon(rvh_path):
    out, on = [], False
    for line in open(rvh_path):
        t = line.replace(",", " ").split()
        if not t:
            continue
        if t[0] == ":HRUs":
            on = True; continue
        if t[0] == ":EndHRUs":
            break
        if on and not t[0].startswith(":"):
            out.append((float(t[3]), float(t[4])))
    return out


def write_hmets(dirpath, n_sb=1, hru_per_sb=1, n_days=365, seed=1, routing="ROUTE_NONE", catchment_route="ROUTE_DUMP",
                single_class=False, ensemble_members=0, name="hmets", season_offset=0, mixed_distributions=False, n_gauges=1):
    """HMETS template (BASELINE config 3 member): returns the input base path."""
    os.makedirs(dirpath, exist_ok=True)
    base = os.path.join(dirpath, name)
    rng = np.random.RandomState(seed)
    extra_rvi = ""
    if ensemble_members > 0:
        extra_rvi = ":EnsembleMode ENSEMBLE_MONTECARLO %d\n:RandomSeed %d\n" % (ensemble_members, seed)
    hmets_rvi_extra = extra_rvi
    profile = "NONE" if routing == "ROUTE_NONE" else "CHN"
    extra_rvp = ""
    if n_sb > 1 or routing != "ROUTE_NONE":
        extra_rvp += ":AvgAnnualRunoff 400\n"
    if profile != "NONE":
        extra_rvp += ":TrapezoidalChannel CHN 10.0 0.0 2.0 0.035 0.001\n"
    _write(base + ".rvp", HMETS_RVP.format(extra=extra_rvp))
    with open(base + ".rvh", "w") as f:
        if single_class:
            write_network(f, n_sb, hru_per_sb, rng, profile, ["FOREST"], ["FOREST"], ["DEFAULT_P"])
        else:
            write_network(f, n_sb, hru_per_sb, rng, profile, ["FOREST", "OPEN"], ["FOREST", "GRASS"], ["DEFAULT_P", "THIN_P"])
    if n_gauges > 1:
        hmets_rvi_extra += write_gauge_grid(base, n_gauges, n_days, seed, season_offset, _hru_latlon(base + ".rvh"))
    else:
        with open(base + ".rvt", "w") as f:
            write_gauge(f, *weather(n_days, seed, season_offset))
    _write(base + ".rvi", HMETS_RVI.format(duration=n_days, routing=routing, catchment_route=catchment_route, extra=hmets_rvi_extra))
    _write(base + ".rvc", HMETS_RVC)
    if ensemble_members > 0:
        with open(base + ".rve", "w") as f:
            f.write(":OutputDirectoryFormat ./ens_*\n:ParameterDistributions\n  :Attributes NAME CLASS_TYPE CLASS BASE_VALUE DIST PAR1 PAR2\n")
            for i, (pn, ct, cn, v) in enumerate(HMETS_DISTS):
                lo, hi = sorted([0.8 * v, 1.2 * v])
                if mixed_distributions and i % 4 == 1:
                    f.write("  %s %s %s %.8g DIST_NORMAL %.8g %.8g\n" % (pn, ct, cn, v, v, 0.03 * abs(v)))
                elif mixed_distributions and i % 4 == 2:
                    f.write("  %s %s %s %.8g DIST_GAMMA %.8g %.8g\n" % (pn, ct, cn, v, 40.0, abs(v) / 40.0))
                else:
                    f.write("  %s %s %s %.8g DIST_UNIFORM %.8g %.8g\n" % (pn, ct, cn, v, lo, hi))
            f.write(":EndParameterDistributions\n")
    return base


# ------------------------------------------------------------------------------------------------ HBV-EC
HBV_RVI = """# HBV-EC emulation, synthetic watershed (written by ravenhydroframework_b200.synth)
:StartDate             2000-01-01 00:00:00
:Duration              {duration}
:TimeStep              1.0
:Method                ORDERED_SERIES
:Routing               {routing}
:CatchmentRoute        {catchment_route}
:Evaporation           PET_FROMMONTHLY
:OW_Evaporation        PET_FROMMONTHLY
:SWRadiationMethod     SW_RAD_DEFAULT
:SWCloudCorrect        SW_CLOUD_CORR_NONE
:SWCanopyCorrect       SW_CANOPY_CORR_NONE
:LWRadiationMethod     LW_RAD_DEFAULT
:RainSnowFraction      RAINSNOW_HBV
:PotentialMeltMethod   POTMELT_HBV
:OroTempCorrect        OROCORR_HBV
:OroPrecipCorrect      OROCORR_HBV
:OroPETCorrect         OROCORR_HBV
:CloudCoverMethod      CLOUDCOV_NONE
:PrecipIceptFract      PRECIP_ICEPT_USER
:MonthlyInterpolationMethod MONTHINT_LINEAR_21
:SoilModel             SOIL_MULTILAYER 3
:SilentMode
{extra}
:Alias       FAST_RESERVOIR SOIL[1]
:Alias       SLOW_RESERVOIR SOIL[2]
:LakeStorage SLOW_RESERVOIR
:HydrologicProcesses
  :SnowRefreeze      FREEZE_DEGREE_DAY  SNOW_LIQ        SNOW
  :Precipitation     PRECIP_RAVEN       ATMOS_PRECIP    MULTIPLE
  :CanopyEvaporation CANEVP_ALL         CANOPY          ATMOSPHERE
  :CanopySnowEvap    CANEVP_ALL         CANOPY_SNOW     ATMOSPHERE
  :SnowBalance       SNOBAL_SIMPLE_MELT SNOW            SNOW_LIQ
    :-->Overflow     RAVEN_DEFAULT      SNOW_LIQ        PONDED_WATER
  :Flush             RAVEN_DEFAULT      PONDED_WATER    GLACIER
    :-->Conditional HRU_TYPE IS GLACIER
  :GlacierMelt       GMELT_HBV          GLACIER_ICE     GLACIER
  :GlacierRelease    GRELEASE_HBV_EC    GLACIER         SURFACE_WATER
  :Infiltration      INF_HBV            PONDED_WATER    MULTIPLE
  :Flush             RAVEN_DEFAULT      SURFACE_WATER   FAST_RESERVOIR
    :-->Conditional HRU_TYPE IS_NOT GLACIER
  :SoilEvaporation   SOILEVAP_HBV       SOIL[0]         ATMOSPHERE
  :CapillaryRise     RISE_HBV           FAST_RESERVOIR  SOIL[0]
  :LakeEvaporation   LAKE_EVAP_BASIC    SLOW_RESERVOIR  ATMOSPHERE
  :Percolation       PERC_CONSTANT      FAST_RESERVOIR  SLOW_RESERVOIR
  :Baseflow          BASE_POWER_LAW     FAST_RESERVOIR  SURFACE_WATER
  :Baseflow          BASE_LINEAR        SLOW_RESERVOIR  SURFACE_WATER
:EndHydrologicProcesses
"""

HBV_RVP = """# HBV-EC parameters (synthetic)
:SoilClasses
  :Attributes,
  :Units,
  TOPSOIL,
  SLOW_RES,
  FAST_RES,
:EndSoilClasses
:SoilProfiles
  DEFAULT_P,  3, TOPSOIL, 2.036937, FAST_RES, 100.0, SLOW_RES, 100.0
  SHALLOW_P,  3, TOPSOIL, 1.100000, FAST_RES, 100.0, SLOW_RES, 100.0
  GLACIER_P,  3, TOPSOIL, 0.500000, FAST_RES, 100.0, SLOW_RES, 100.0
:EndSoilProfiles
:VegetationClasses
  :Attributes,  MAX_HT, MAX_LAI, MAX_LEAF_COND
  :Units,            m,    none,      mm_per_s
  VEG_ALL,          25,     6.0,           5.3
  VEG_LOW,           3,     2.5,           5.3
:EndVegetationClasses
:VegetationParameterList
  :Parameters,  MAX_CAPACITY, MAX_SNOW_CAPACITY,  TFRAIN,  TFSNOW,
  :Units,                 mm,                mm,    frac,    frac,
  VEG_ALL,             10000,             10000,    0.88,    0.88,
  VEG_LOW,               2.5,               4.0,    0.95,    0.93,
:EndVegetationParameterList
:SeasonalRelativeLAI
  VEG_LOW, 0.3, 0.3, 0.4, 0.6, 0.9, 1.0, 1.0, 1.0, 0.8, 0.5, 0.3, 0.3
:EndSeasonalRelativeLAI
:LandUseClasses
  :Attributes,  IMPERM, FOREST_COV
  :Units,         frac,       frac
  LU_ALL,          0.0,          1
  LU_OPEN,        0.02,       0.25
  LU_ICE,          0.0,        0.0
:EndLandUseClasses
:GlobalParameter ADIABATIC_LAPSE     0.5718813
:GlobalParameter RAINSNOW_TEMP       0.05984519
:GlobalParameter RAINSNOW_DELTA             2.0
:GlobalParameter SNOW_SWI            0.03473693
:GlobalParameter PRECIP_LAPSE          4.873686
:LandUseParameterList
  :Parameters,   MELT_FACTOR, MIN_MELT_FACTOR,   HBV_MELT_FOR_CORR, REFREEZE_FACTOR, HBV_MELT_ASP_CORR
  :Units     ,        mm/d/K,          mm/d/K,                none,          mm/d/K,              none
    [DEFAULT],      4.072232,             2.2,           0.4452843,        2.001574,              0.48
    LU_OPEN,        4.600000,             2.5,            _DEFAULT,        1.700000,          _DEFAULT
:EndLandUseParameterList
:LandUseParameterList
  :Parameters, HBV_MELT_GLACIER_CORR,   HBV_GLACIER_KMIN, GLAC_STORAGE_COEFF, HBV_GLACIER_AG
  :Units     ,                  none,                1/d,                1/d,           1/mm
   [DEFAULT],                  1.64,               0.05,          0.6771759,           0.05
:EndLandUseParameterList
:SoilParameterList
  :Parameters,   POROSITY, FIELD_CAPACITY,   SAT_WILT,  HBV_BETA, MAX_CAP_RISE_RATE, MAX_PERC_RATE, BASEFLOW_COEFF, BASEFLOW_N
  :Units     ,       none,           none,       none,      none,              mm/d,          mm/d,            1/d,       none
    [DEFAULT], 0.09985144,      0.5060520, 0.04505643,  3.438486,          18.94145,           0.0,            0.0,        0.0
     FAST_RES,   _DEFAULT,       _DEFAULT,        0.0,  _DEFAULT,          _DEFAULT,      38.32455,      0.4606565,   1.877607
     SLOW_RES,   _DEFAULT,       _DEFAULT,        0.0,  _DEFAULT,          _DEFAULT,      _DEFAULT,     0.06303738,        1.0
:EndSoilParameterList
{extra}
"""

HBV_RVC = """# initial lower groundwater storage
:UniformInitialConditions SOIL[2] 0.50657
:UniformInitialConditions SOIL[0] 60.0
"""

MONTHLY_EVAP = [0.028, 0.089, 0.315, 1.003, 2.077, 2.959, 3.190, 2.548, 1.382, 0.520, 0.101, 0.023]
MONTHLY_TEMP = [-11.4, -7.2, -2.8, 2.8, 8.1, 12.2, 14.4, 13.5, 8.9, 3.4, -4.0, -9.3]


# Algorithm sets for write_hbv(variant="algo:<key>"): each swaps the infiltration / soil evaporation / percolation / baseflow
# algorithms of the HBV-EC process list for other members of the same process classes.
ALGO_SETS = {
    "a": dict(inf="INF_RATIONAL", sevap="SOILEVAP_LINEAR", perc="PERC_GAWSER", base_fast="BASE_LINEAR_CONSTRAIN", base_slow="BASE_CONSTANT"),
    "b": dict(inf="INF_ALL_INFILTRATES", sevap="SOILEVAP_PDM", perc="PERC_GAWSER_CONSTRAIN", base_fast="BASE_THRESH_POWER", base_slow="BASE_THRESH_STOR"),
    "c": dict(inf="INF_VIC", sevap="SOILEVAP_VIC", perc="PERC_POWER_LAW", base_fast="BASE_VIC"),
    "d": dict(inf="INF_PRMS", sevap="SOILEVAP_SEQUEN", perc="PERC_PRMS"),
    "e": dict(inf="INF_PDM", sevap="SOILEVAP_HYMOD2", perc="PERC_LINEAR_ANALYTIC"),
    "f": dict(sevap="SOILEVAP_ROOT", perc="PERC_SACRAMENTO", perc_top="PERC_SACRAMENTO"),
    "g": dict(sevap="SOILEVAP_ROOT_CONSTRAIN", perc="PERC_ASPEN"),
    "h": dict(sevap="SOILEVAP_AWBM", inf="INF_PDM"),
    "i": dict(sevap="SOILEVAP_UBC"),
    "j": dict(sevap="SOILEVAP_HYPR", depression=True),
    "k": dict(inf="INF_GREEN_AMPT", rvp=""":SoilParameterList
  :Parameters, HYDRAUL_COND, WETTING_FRONT_PSI
  :Units     ,         mm/d,               -mm
    [DEFAULT],          4.5,             110.0
     FAST_RES,         22.0,              75.0
:EndSoilParameterList
"""),
    "l": dict(inf="INF_GA_SIMPLE", rvp=""":SoilParameterList
  :Parameters, HYDRAUL_COND, WETTING_FRONT_PSI
  :Units     ,         mm/d,               -mm
    [DEFAULT],          4.5,             110.0
     FAST_RES,         22.0,              75.0
:EndSoilParameterList
"""),
    "m": dict(inf="INF_SCS", rvp=""":LandUseParameterList
  :Parameters, SCS_CN, SCS_IA_FRACTION
  :Units     ,   none,            none
    [DEFAULT],   71.0,            0.18
    LU_OPEN,     83.0,            0.20
:EndLandUseParameterList
"""),
    "n": dict(inf="INF_SCS_NOABSTRACTION", rvp=""":LandUseParameterList
  :Parameters, SCS_CN
  :Units     ,   none
    [DEFAULT],   64.0
    LU_OPEN,     78.0
:EndLandUseParameterList
"""),
}
ALGO_RVP = """:SoilParameterList
  :Parameters, PERC_N, SAC_PERC_ALPHA, SAC_PERC_EXPON, PERC_ASPEN, BASEFLOW_THRESH, BASEFLOW_COEFF2, STORAGE_THRESHOLD, MAX_BASEFLOW_RATE, PERC_COEFF
  :Units     ,   none,           none,           none,       mm/d,            none,             1/d,                mm,              mm/d,        1/d
    [DEFAULT],    1.7,           35.0,            2.4,        1.9,            0.22,           0.083,              12.5,              3.25,      0.061
     FAST_RES,    2.2,           28.0,            3.1,        2.3,            0.15,           0.120,               9.0,              9.75,      0.143
:EndSoilParameterList
:SoilParameterList
  :Parameters, VIC_ZMIN, VIC_ZMAX, VIC_ALPHA, VIC_EVAP_GAMMA, UBC_EVAP_SOIL_DEF, UBC_INFIL_SOIL_DEF
  :Units     ,       mm,       mm,      none,           none,                mm,                 mm
    [DEFAULT],     0.05,     1.35,       0.8,            1.6,             105.0,               85.0
:EndSoilParameterList
:LandUseParameterList
  :Parameters, PARTITION_COEFF, MAX_SAT_AREA_FRAC, PDM_B, AET_COEFF, MAX_DEP_AREA_FRAC, PONDED_EXP, DEP_MAX, AWBM_AREAFRAC1, AWBM_AREAFRAC2
  :Units     ,            none,              none,  none,       1/d,              none,       none,      mm,           none,           none
    [DEFAULT],            0.37,              0.62,   1.4,     0.031,              0.45,        1.8,    40.0,          0.134,          0.433
    LU_OPEN,              0.52,              0.48,   0.9,     0.044,              0.30,        2.2,    25.0,          0.210,          0.390
:EndLandUseParameterList
:LandUseParameterList
  :Parameters, HYMOD2_G, HYMOD2_KMAX, HYMOD2_EXP, PDMROF_B
  :Units     ,     none,        none,       none,     none
    [DEFAULT],     0.28,        0.91,        1.3,      0.5
:EndLandUseParameterList
"""


def write_hbv(dirpath, n_sb=1, hru_per_sb=1, n_days=365, seed=1, routing="ROUTE_NONE", catchment_route="TRIANGULAR_UH",
              single_class=False, name="hbv", season_offset=0, n_gauges=1, variant=None):
    """HBV-EC template (BASELINE config 2): multi-class, with glacier HRUs unless single_class. Returns the input base path.
    variant="melt_squeeze_interflow": the snow balance is replaced by :SnowMelt + :SnowSqueeze and a PRMS :Interflow drains the
    fast reservoir (process classes outside the three templates)."""
    os.makedirs(dirpath, exist_ok=True)
    base = os.path.join(dirpath, name)
    rng = np.random.RandomState(seed)
    profile = "NONE" if routing == "ROUTE_NONE" else "CHN"
    extra_rvp = ""
    if n_sb > 1 or routing != "ROUTE_NONE":
        extra_rvp += ":AvgAnnualRunoff 400\n"
    if profile != "NONE":
        extra_rvp += ":TrapezoidalChannel CHN 10.0 0.0 2.0 0.035 0.001\n"
    if variant == "melt_squeeze_interflow":
        extra_rvp += ":SoilParameterList\n  :Parameters, MAX_INTERFLOW_RATE\n  :Units, mm/d\n  [DEFAULT], 0.0\n  FAST_RES, 7.25\n:EndSoilParameterList\n"
    algo = ALGO_SETS[variant[5:]] if (variant or "").startswith("algo:") else None
    if algo is not None:
        extra_rvp += ALGO_RVP + algo.get("rvp", "")
    _write(base + ".rvp", HBV_RVP.format(extra=extra_rvp))
    with open(base + ".rvh", "w") as f:
        if single_class:
            write_network(f, n_sb, hru_per_sb, rng, profile, ["LU_ALL"], ["VEG_ALL"], ["DEFAULT_P"])
        else:
            write_network(f, n_sb, hru_per_sb, rng, profile, ["LU_ALL", "LU_OPEN", "LU_ICE"], ["VEG_ALL", "VEG_LOW", "VEG_LOW"],
                          ["DEFAULT_P", "SHALLOW_P", "GLACIER_P"], paired_profiles=True)
        f.write(":SubBasinProperties\n  :Parameters, TIME_CONC, TIME_TO_PEAK\n  :Units, d, d\n")
        for p in range(1, n_sb + 1):
            tc = rng.uniform(1.2, 3.2)
            f.write("  %d, %.6f, %.6f\n" % (p, tc, 0.5 * tc))
        f.write(":EndSubBasinProperties\n")
    gx = "  :RainCorrection 1.141608\n  :SnowCorrection 1.024278\n"
    gx += "  :MonthlyAveEvaporation, " + ", ".join("%.3f" % v for v in MONTHLY_EVAP) + ",\n"
    gx += "  :MonthlyAveTemperature, " + ", ".join("%.1f" % v for v in MONTHLY_TEMP) + ",\n"
    extra_rvi = ""
    if n_gauges > 1:
        extra_rvi = write_gauge_grid(base, n_gauges, n_days, seed, season_offset, _hru_latlon(base + ".rvh"), extra=gx)
    else:
        with open(base + ".rvt", "w") as f:
            write_gauge(f, *weather(n_days, seed, season_offset), extra=gx)
    rvi = HBV_RVI.format(duration=n_days, routing=routing, catchment_route=catchment_route, extra=extra_rvi)
    if variant == "melt_squeeze_interflow":
        a = "  :SnowBalance       SNOBAL_SIMPLE_MELT SNOW            SNOW_LIQ\n    :-->Overflow     RAVEN_DEFAULT      SNOW_LIQ        PONDED_WATER\n"
        b = "  :SnowMelt          MELT_POTMELT       SNOW            SNOW_LIQ\n  :SnowSqueeze       SQUEEZE_RAVEN      SNOW_LIQ        PONDED_WATER\n"
        assert a in rvi
        rvi = rvi.replace(a, b)
        c = "  :LakeEvaporation   LAKE_EVAP_BASIC    SLOW_RESERVOIR  ATMOSPHERE\n"
        assert c in rvi
        rvi = rvi.replace(c, "  :Interflow         INTERFLOW_PRMS     FAST_RESERVOIR  SURFACE_WATER\n" + c)
    if algo is not None:
        def swap(line_old, line_new):
            nonlocal rvi
            assert line_old in rvi, line_old
            rvi = rvi.replace(line_old, line_new)
        if "inf" in algo:
            swap("  :Infiltration      INF_HBV            PONDED_WATER    MULTIPLE\n", "  :Infiltration      %s PONDED_WATER MULTIPLE\n" % algo["inf"])
        if "sevap" in algo:
            swap("  :SoilEvaporation   SOILEVAP_HBV       SOIL[0]         ATMOSPHERE\n", "  :SoilEvaporation   %s SOIL[0] ATMOSPHERE\n" % algo["sevap"])
        if "perc" in algo:
            swap("  :Percolation       PERC_CONSTANT      FAST_RESERVOIR  SLOW_RESERVOIR\n", ("  :Percolation       %s SOIL[0] FAST_RESERVOIR\n" % algo["perc_top"] if "perc_top" in algo else "")
                 + "  :Percolation       %s FAST_RESERVOIR SLOW_RESERVOIR\n" % algo["perc"])
        if "base_fast" in algo:
            swap("  :Baseflow          BASE_POWER_LAW     FAST_RESERVOIR  SURFACE_WATER\n", "  :Baseflow          %s FAST_RESERVOIR SURFACE_WATER\n" % algo["base_fast"])
        if "base_slow" in algo:
            swap("  :Baseflow          BASE_LINEAR        SLOW_RESERVOIR  SURFACE_WATER\n", "  :Baseflow          %s SLOW_RESERVOIR SURFACE_WATER\n" % algo["base_slow"])
    if variant == "hrugroup_cond":
        # process conditions on HRU groups (reference CHydroProcessABC::ShouldApply, BASIS_HRU_GROUP): capillary rise only inside
        # 'Wetted', the slow baseflow everywhere except 'Sealed'
        a = ":HydrologicProcesses\n"
        rvi = rvi.replace(a, ":DefineHRUGroups Wetted Sealed\n" + a)
        x = "  :CapillaryRise     RISE_HBV           FAST_RESERVOIR  SOIL[0]\n"
        assert x in rvi
        rvi = rvi.replace(x, x + "    :-->Conditional HRU_GROUP IS Wetted\n")
        x = "  :Baseflow          BASE_LINEAR        SLOW_RESERVOIR  SURFACE_WATER\n"
        assert x in rvi
        rvi = rvi.replace(x, x + "    :-->Conditional HRU_GROUP IS_NOT Sealed\n")
        n_hru = n_sb * hru_per_sb
        with open(base + ".rvh", "a") as f:
            f.write(":HRUGroup Wetted\n  " + " ".join(str(k) for k in range(1, n_hru + 1) if k % 2 == 0) + "\n:EndHRUGroup\n")
            f.write(":HRUGroup Sealed\n  " + " ".join(str(k) for k in range(1, n_hru + 1) if k % 3 == 0) + "\n:EndHRUGroup\n")
    if variant == "lateral":
        # lateral exchange (reference CmvLatFlush): surface water of the 'UpGroup' HRUs is flushed into the fast reservoir of the
        # 'LowGroup' HRUs of the same subbasin (limited by the room left there); one more flush moves ponded water to one single
        # recipient HRU across basins
        a = ":HydrologicProcesses\n"
        assert a in rvi
        rvi = rvi.replace(a, ":DefineHRUGroups UpGroup LowGroup Sink\n" + a)
        e = ":EndHydrologicProcesses"
        rvi = rvi.replace(e, "  :LateralFlush      RAVEN_DEFAULT      UpGroup SURFACE_WATER To LowGroup FAST_RESERVOIR\n"
                             "  :LateralFlush      RAVEN_DEFAULT      LowGroup SNOW_LIQ To Sink SOIL[0] INTERBASIN\n" + e)
        n_hru = n_sb * hru_per_sb
        with open(base + ".rvh", "a") as f:
            f.write(":HRUGroup UpGroup\n  " + " ".join(str(k) for k in range(1, n_hru + 1) if k % 3 != 0) + "\n:EndHRUGroup\n")
            f.write(":HRUGroup LowGroup\n  " + " ".join(str(k) for k in range(1, n_hru + 1) if k % 3 == 0 or k % 5 == 0) + "\n:EndHRUGroup\n")
            f.write(":HRUGroup Sink\n  2\n:EndHRUGroup\n")
    if variant in ("redistribute", "redistribute_threshold"):
        # gravitational snow redistribution (reference CmvLatRedistribute): within each subbasin HRU j sheds snow to HRUs j+1 and j+2
        # (weights 0.6 / 0.4; 1.0 for the last pair) along the :LateralConnections table.  Slopes are steepened (x3.2, up to 64 deg)
        # so that the threshold method's holding depth is exceeded and its > 60 deg branch is taken.
        e = ":EndHydrologicProcesses"
        rvi = rvi.replace(e, "  :SnowRedistribute  %s SNOW %s\n" % (("THRESHOLD", "500.0") if variant.endswith("threshold") else ("CONTINUOUS", "150.0")) + e)
        lines = open(base + ".rvh").read().split("\n")
        i0 = [i for i, l in enumerate(lines) if l.startswith(":HRUs")][0] + 3
        i1 = [i for i, l in enumerate(lines) if l.startswith(":EndHRUs")][0]
        for i in range(i0, i1):
            t = lines[i].split(",")
            t[11] = " %.3f" % (3.2 * float(t[11]))
            lines[i] = ",".join(t)
        con = [":LateralConnections SNOW_REDISTRIBUTE"]
        for p_ in range(n_sb):
            ks = [p_ * hru_per_sb + j + 1 for j in range(hru_per_sb)]
            for j in range(len(ks) - 1):
                if j + 2 < len(ks):
                    con += ["  %d %d 0.6" % (ks[j], ks[j + 1]), "  %d %d 0.4" % (ks[j], ks[j + 2])]
                else:
                    con += ["  %d %d 1.0" % (ks[j], ks[j + 1])]
        con += [":EndLateralConnections"]
        open(base + ".rvh", "w").write("\n".join(lines + con) + "\n")
    if variant in ("equilibrate", "equilibrate_interbasin"):
        # lateral mixing (reference CmvLatEquilibrate): the fast reservoirs of the 'MixGroup' HRUs of a subbasin are partially
        # mixed each step; with INTERBASIN the snow packs of the 'Wet' HRUs are mixed across the whole model
        a = ":HydrologicProcesses\n"
        assert a in rvi
        rvi = rvi.replace(a, ":DefineHRUGroups MixGroup Wet\n" + a)
        e = ":EndHydrologicProcesses"
        rvi = rvi.replace(e, "  :LateralEquilibrate RAVEN_DEFAULT     MixGroup FAST_RESERVOIR 0.35\n"
                          + ("  :LateralEquilibrate RAVEN_DEFAULT     Wet SNOW 0.6 INTERBASIN\n" if variant == "equilibrate_interbasin" else "") + e)
        n_hru = n_sb * hru_per_sb
        with open(base + ".rvh", "a") as f:
            f.write(":HRUGroup MixGroup\n  " + " ".join(str(k) for k in range(1, n_hru + 1) if k % 4 != 0) + "\n:EndHRUGroup\n")
            f.write(":HRUGroup Wet\n  " + " ".join(str(k) for k in range(1, n_hru + 1) if k % 2 == 0) + "\n:EndHRUGroup\n")
    _write(base + ".rvi", rvi)
    _write(base + ".rvc", HBV_RVC + (":UniformInitialConditions DEPRESSION 18.0\n" if (algo or {}).get("depression") else ""))
    return base


# ------------------------------------------------------------------------------------------------ GR4J
GR4J_RVI = """# GR4J + Cema-Neige emulation, synthetic watershed (written by ravenhydroframework_b200.synth)
:StartDate             2000-01-01 00:00:00
:Duration              {duration}
:TimeStep              1.0
:Method                ORDERED_SERIES
:SoilModel             SOIL_MULTILAYER  4
:Routing               {routing}
:CatchmentRoute        {catchment_route}
:Evaporation           PET_DATA
:RainSnowFraction      RAINSNOW_DINGMAN
:PotentialMeltMethod   POTMELT_DEGREE_DAY
:OroTempCorrect        OROCORR_SIMPLELAPSE
:OroPrecipCorrect      OROCORR_SIMPLELAPSE
:SilentMode
{extra}
:Alias PRODUCT_STORE      SOIL[0]
:Alias ROUTING_STORE      SOIL[1]
:Alias TEMP_STORE         SOIL[2]
:Alias GW_STORE           SOIL[3]
:HydrologicProcesses
 :Precipitation            PRECIP_RAVEN       ATMOS_PRECIP    MULTIPLE
 :SnowTempEvolve           SNOTEMP_NEWTONS    SNOW_TEMP
 :SnowBalance              SNOBAL_CEMA_NIEGE  SNOW            PONDED_WATER
 :OpenWaterEvaporation     OPEN_WATER_EVAP    PONDED_WATER    ATMOSPHERE
 :Infiltration             INF_GR4J           PONDED_WATER    MULTIPLE
 :SoilEvaporation          SOILEVAP_GR4J      PRODUCT_STORE   ATMOSPHERE
 :Percolation              PERC_GR4J          PRODUCT_STORE   TEMP_STORE
 :Flush                    RAVEN_DEFAULT      SURFACE_WATER   TEMP_STORE
 :Split                    RAVEN_DEFAULT      TEMP_STORE      CONVOLUTION[0] CONVOLUTION[1] 0.9
 :Convolve                 CONVOL_GR4J_1      CONVOLUTION[0]  ROUTING_STORE
 :Convolve                 CONVOL_GR4J_2      CONVOLUTION[1]  TEMP_STORE
 :Percolation              PERC_GR4JEXCH      ROUTING_STORE   GW_STORE
 :Percolation              PERC_GR4JEXCH2     TEMP_STORE      GW_STORE
 :Flush                    RAVEN_DEFAULT      TEMP_STORE      SURFACE_WATER
 :Baseflow                 BASE_GR4J          ROUTING_STORE   SURFACE_WATER
:EndHydrologicProcesses
"""

GR4J_RVP = """# GR4J parameters (synthetic)
:SoilClasses
  :Attributes
  :Units
   SOIL_PROD
   SOIL_ROUT
   SOIL_TEMP
   SOIL_GW
:EndSoilClasses
:SoilProfiles
  DEFAULT_P,  4, SOIL_PROD, 0.529, SOIL_ROUT, 0.300, SOIL_TEMP, 1.000, SOIL_GW, 1.000,
  THIN_P,     4, SOIL_PROD, 0.310, SOIL_ROUT, 0.300, SOIL_TEMP, 1.000, SOIL_GW, 1.000,
:EndSoilProfiles
:VegetationClasses
   :Attributes,       MAX_HT,       MAX_LAI,      MAX_LEAF_COND
        :Units,            m,          none,           mm_per_s
       VEG_ALL,           0.0,          0.0,                0.0
:EndVegetationClasses
:LandUseClasses
  :Attributes, IMPERM, FOREST_COV
  :Units     ,   frac,       frac
       LU_ALL,    0.0,        0.0
       LU_TOWN,   0.3,        0.0
:EndLandUseClasses
:GlobalParameter RAINSNOW_TEMP       0.0
:GlobalParameter RAINSNOW_DELTA      1.0
:GlobalParameter AIRSNOW_COEFF     0.053
:GlobalParameter AVG_ANNUAL_SNOW    16.9
:GlobalParameter PRECIP_LAPSE     0.0004
:GlobalParameter ADIABATIC_LAPSE  0.0065
:SoilParameterList
 :Parameters, POROSITY ,  GR4J_X3, GR4J_X2
 :Units     ,     none ,       mm,    mm/d
   [DEFAULT],      1.0 ,   407.29,  -3.396
:EndSoilParameterList
:LandUseParameterList
 :Parameters, GR4J_X4, MELT_FACTOR
 :Units     ,       d,      mm/d/C
   [DEFAULT],   1.072,        7.73
   LU_TOWN,     2.600,        6.10
:EndLandUseParameterList
{extra}
"""

GR4J_RVC = """:UniformInitialConditions SOIL[0] 150.0
:UniformInitialConditions SOIL[1] 15.0
"""


def write_gr4j(dirpath, n_sb=1, hru_per_sb=1, n_days=365, seed=1, routing="ROUTE_NONE", catchment_route="ROUTE_DUMP",
               single_class=False, name="gr4j", season_offset=0):
    """GR4J + Cema-Neige template (BASELINE config 1 process list). :OW_Evaporation is left at the reference's default
    (PET_HARGREAVES_1985), so the open-water PET that OPEN_WATER_EVAP consumes comes from extraterrestrial radiation."""
    os.makedirs(dirpath, exist_ok=True)
    base = os.path.join(dirpath, name)
    rng = np.random.RandomState(seed)
    _write(base + ".rvi", GR4J_RVI.format(duration=n_days, routing=routing, catchment_route=catchment_route, extra=""))
    profile = "NONE" if routing == "ROUTE_NONE" else "CHN"
    extra_rvp = ""
    if n_sb > 1 or routing != "ROUTE_NONE":
        extra_rvp += ":AvgAnnualRunoff 400\n"
    if profile != "NONE":
        extra_rvp += ":TrapezoidalChannel CHN 10.0 0.0 2.0 0.035 0.001\n"
    _write(base + ".rvp", GR4J_RVP.format(extra=extra_rvp))
    with open(base + ".rvh", "w") as f:
        if single_class:
            write_network(f, n_sb, hru_per_sb, rng, profile, ["LU_ALL"], ["VEG_ALL"], ["DEFAULT_P"])
        else:
            write_network(f, n_sb, hru_per_sb, rng, profile, ["LU_ALL", "LU_TOWN"], ["VEG_ALL", "VEG_ALL"], ["DEFAULT_P", "THIN_P"])
    with open(base + ".rvt", "w") as f:
        write_gauge(f, *weather(n_days, seed, season_offset))
    _write(base + ".rvc", GR4J_RVC)
    return base


def add_enkf(base, n_members, n_sb, n_hru, n_days, obs_basins=(1,), seed=3, window=1, mode="ENKF_SPINUP", random_seed=11,
             assimilated=("SOIL[0]", "SNOW"), start="2000-01-01", obs_level=6.0,
             perturbations=("RAINFALL DIST_GAMMA 4.0 0.25 MULTIPLICATIVE", "TEMP_AVE DIST_NORMAL 0 1.5 ADDITIVE"),
             obs_error="STREAMFLOW DIST_NORMAL 1 0.07 MULTIPLICATIVE", hru_group_range=None):
    """Turn the input set at `base` (written by write_hbv / write_hmets / write_gr4j) into an EnKF ensemble run (BASELINE
    config 5): `:EnsembleMode ENSEMBLE_ENKF`, HRU / subbasin groups, synthetic HYDROGRAPH observations at `obs_basins`, and the
    .rve commands of reference src/ParseEnsembleFile.cpp (:EnKFMode, :ForcingPerturbation, :ObservationErrorModel,
    :AssimilatedState, :AssimilateStreamflow)."""
    rng = np.random.RandomState(seed)
    with open(base + ".rvi", "a") as f:
        f.write(":EnsembleMode ENSEMBLE_ENKF %d\n:RandomSeed %d\n" % (n_members, random_seed))
    lo, hi = hru_group_range or (1, n_hru)
    with open(base + ".rvh", "a") as f:
        f.write(":HRUGroup AssimHRUs\n  %d-%d\n:EndHRUGroup\n" % (lo, hi))
        f.write(":SubBasinGroup AssimSBs\n")
        ids = list(range(1, n_sb + 1))
        for i in range(0, len(ids), 20):
            f.write("  " + " ".join(str(x) for x in ids[i:i + 20]) + "\n")
        f.write(":EndSubBasinGroup\n")
    with open(base + ".rvt", "a") as f:
        for sb in obs_basins:
            f.write(":ObservationData HYDROGRAPH %d m3/s\n  %s 00:00:00 1.0 %d\n" % (sb, start, n_days + 1))
            q = obs_level * (1.0 + 0.5 * rng.uniform(size=n_days + 1))
            for v in q:
                f.write("  %.4f\n" % v)
            f.write(":EndObservationData\n")
    with open(base + ".rve", "w") as f:
        f.write(":EnKFMode %s\n:WindowSize %d\n" % (mode, window))
        for p in perturbations:
            f.write(":ForcingPerturbation %s\n" % p)
        if obs_error:
            f.write(":ObservationErrorModel %s\n" % obs_error)
        for a in assimilated:
            f.write(":AssimilatedState %s AssimHRUs\n" % a)
        f.write(":AssimilatedState STREAMFLOW AssimSBs\n")
        for sb in obs_basins:
            f.write(":AssimilateStreamflow %d\n" % sb)
    return base


# ------------------------------------------------------------------------------------------------ Blended (Mai et al. 2020)
BLENDED_RVI = """# Blended model emulation (process groups with weights), synthetic watershed (written by ravenhydroframework_b200.synth)
:StartDate             2000-01-01 00:00:00
:Duration              {duration}
:TimeStep              1.0
:Method                ORDERED_SERIES
:PotentialMeltMethod   POTMELT_HMETS
:RainSnowFraction      RAINSNOW_HBV
:Evaporation           PET_DATA
:CatchmentRoute        {catchment_route}
:Routing               {routing}
:SoilModel             SOIL_MULTILAYER 3
:SilentMode
{extra}
:Alias DELAYED_RUNOFF CONVOLUTION[1]
:Alias DIRECT_RUNOFF  SURFACE_WATER
:HydrologicProcesses
  :Precipitation     RAVEN_DEFAULT         ATMOS_PRECIP   MULTIPLE
  :ProcessGroup
    :Infiltration    INF_HMETS             PONDED_WATER   MULTIPLE
    :Infiltration    INF_VIC_ARNO          PONDED_WATER   MULTIPLE
    :Infiltration    INF_HBV               PONDED_WATER   MULTIPLE
  :EndProcessGroup CALCULATE_WTS {r[0]} {r[1]}
    :Overflow        OVERFLOW_RAVEN        SOIL[0]        DIRECT_RUNOFF
  :ProcessGroup
    :Baseflow        BASE_LINEAR_ANALYTIC  SOIL[0]        SURFACE_WATER
    :Baseflow        BASE_VIC              SOIL[0]        SURFACE_WATER
    :Baseflow        BASE_TOPMODEL         SOIL[0]        SURFACE_WATER
  :EndProcessGroup CALCULATE_WTS {r[2]} {r[3]}
  :Percolation       PERC_LINEAR           SOIL[0]        SOIL[1]
    :Overflow        OVERFLOW_RAVEN        SOIL[1]        DIRECT_RUNOFF
  :Percolation       PERC_LINEAR           SOIL[1]        SOIL[2]
  :ProcessGroup
    :SoilEvaporation SOILEVAP_ALL          SOIL[0]        ATMOSPHERE
    :SoilEvaporation SOILEVAP_TOPMODEL     SOIL[0]        ATMOSPHERE
  :EndProcessGroup CALCULATE_WTS {r[4]}
  :Convolve          CONVOL_GAMMA          CONVOLUTION[0] SURFACE_WATER
  :Convolve          CONVOL_GAMMA_2        DELAYED_RUNOFF SURFACE_WATER
  :ProcessGroup
    :Baseflow        BASE_LINEAR_ANALYTIC  SOIL[1]        SURFACE_WATER
    :Baseflow        BASE_POWER_LAW        SOIL[1]        SURFACE_WATER
  :EndProcessGroup CALCULATE_WTS {r[5]}
  :ProcessGroup
    :SnowBalance     SNOBAL_HMETS          MULTIPLE       MULTIPLE
    :SnowBalance     SNOBAL_SIMPLE_MELT    SNOW           PONDED_WATER
    :SnowBalance     SNOBAL_HBV            MULTIPLE       MULTIPLE
  :EndProcessGroup CALCULATE_WTS {r[6]} {r[7]}
:EndHydrologicProcesses
"""

BLENDED_RVP = """# Blended model parameters (synthetic)
:SoilClasses
  :Attributes,
  :Units,
  TOPSOIL,
  PHREATIC,
  DEEP_GW,
:EndSoilClasses
:LandUseClasses,
  :Attributes,  IMPERM, FOREST_COV,
  :Units,         frac,       frac,
  FOREST,          0.0,        1.0,
  OPEN,           0.05,        0.1,
:EndLandUseClasses
:VegetationClasses,
  :Attributes,  MAX_HT, MAX_LAI, MAX_LEAF_COND,
  :Units,            m,    none,      mm_per_s,
  FOREST,            4,       5,             5,
  GRASS,           0.5,       2,             5,
:EndVegetationClasses
:TerrainClasses
  :Attributes, HILLSLOPE_LENGTH, DRAINAGE_DENSITY, TOPMODEL_LAMBDA
  :Units,                     m,             km/km2,            m
  DEFAULT_T,                1.0,                1.0,       7.0832
  STEEP_T,                  1.0,                1.0,       5.9000
:EndTerrainClasses
:SoilProfiles
  LAKE, 0
  ROCK, 0
  DEFAULT_P, 3, TOPSOIL, 0.4038, PHREATIC, 1.0199, DEEP_GW, 5.0
  THIN_P,    3, TOPSOIL, 0.2700, PHREATIC, 0.7100, DEEP_GW, 5.0
:EndSoilProfiles
:GlobalParameter     SNOW_SWI_MIN 0.0355
:GlobalParameter     SNOW_SWI_MAX 0.1838
:GlobalParameter SWI_REDUCT_COEFF 0.0120
:GlobalParameter         SNOW_SWI 0.0420
:GlobalParameter    RAINSNOW_TEMP -0.1500
:GlobalParameter   RAINSNOW_DELTA 1.6500
:SoilParameterList
  :Parameters, POROSITY, PERC_COEFF, PET_CORRECTION, BASEFLOW_COEFF, VIC_B_EXP, HBV_BETA, MAX_BASEFLOW_RATE, BASEFLOW_N, FIELD_CAPACITY, SAT_WILT
  :Units,             -,        1/d,              -,            1/d,     -,        -,              mm/d,          -,              -,        -
  TOPSOIL,          1.0,     0.0270,         0.8750,         0.0541, 0.1150,   1.9320,           21.200,     2.5400,         0.5314,      0.0
  PHREATIC,         1.0,     0.0042,            0.0,         0.0117,    0.0,      0.0,              0.0,     1.2200,            0.0,      0.0
  DEEP_GW,          1.0,        0.0,            0.0,            0.0,    0.0,      0.0,              0.0,        0.0,            0.0,      0.0
:EndSoilParameterList
:LandUseParameterList
  :Parameters, MIN_MELT_FACTOR, MAX_MELT_FACTOR, DD_MELT_TEMP, DD_AGGRADATION, REFREEZE_FACTOR, REFREEZE_EXP, DD_REFREEZE_TEMP, HMETS_RUNOFF_COEFF,
  :Units,               mm/d/C,          mm/d/C,            C,           1/mm,          mm/d/C,            -,                C,                  -,
  [DEFAULT],            1.2875,          6.7009,       2.3641,         0.0973,          2.6851,       0.3740,          -1.0919,             0.4739,
  OPEN,                 1.6000,          7.9000,       1.8000,         0.0700,          2.1000,       0.4500,          -0.5000,             0.3900,
:EndLandUseParameterList
:LandUseParameterList
  :Parameters, GAMMA_SHAPE, GAMMA_SCALE, GAMMA_SHAPE2, GAMMA_SCALE2,
  :Units,                -,         1/d,            -,          1/d,
  [DEFAULT],        3.5019,      0.8774,       4.3942,       1.1884,
  OPEN,             2.1000,      1.1100,       3.2000,       1.4500,
:EndLandUseParameterList
:VegetationParameterList
  :Parameters, RAIN_ICEPT_PCT, SNOW_ICEPT_PCT,
  :Units,                   -,              -,
  [DEFAULT],              0.0,            0.0,
  FOREST,                0.06,           0.04,
:EndVegetationParameterList
{extra}
"""

BLENDED_RVC = """:UniformInitialConditions SOIL[0] 120.0
:UniformInitialConditions SOIL[1] 300.0
"""


def write_blended(dirpath, n_sb=1, hru_per_sb=1, n_days=365, seed=1, routing="ROUTE_NONE", catchment_route="ROUTE_DUMP",
                  name="blended", season_offset=0, n_gauges=1, r=(0.35, 0.6, 0.2, 0.7, 0.45, 0.8, 0.3, 0.55)):
    """Blended multi-model template (BASELINE config 4): process groups whose members' rates are blended with the weights
    CALCULATE_WTS derives from `r` (reference CProcessGroup).  Returns the input base path."""
    os.makedirs(dirpath, exist_ok=True)
    base = os.path.join(dirpath, name)
    rng = np.random.RandomState(seed)
    profile = "NONE" if routing == "ROUTE_NONE" else "CHN"
    extra_rvp = ""
    if n_sb > 1 or routing != "ROUTE_NONE":
        extra_rvp += ":AvgAnnualRunoff 400\n"
    if profile != "NONE":
        extra_rvp += ":TrapezoidalChannel CHN 10.0 0.0 2.0 0.035 0.001\n"
    _write(base + ".rvp", BLENDED_RVP.format(extra=extra_rvp))
    with open(base + ".rvh", "w") as f:
        write_network(f, n_sb, hru_per_sb, rng, profile, ["FOREST", "OPEN"], ["FOREST", "GRASS"], ["DEFAULT_P", "THIN_P"], terrain_classes=["DEFAULT_T", "STEEP_T"])
    extra_rvi = ""
    if n_gauges > 1:
        extra_rvi += write_gauge_grid(base, n_gauges, n_days, seed, season_offset, _hru_latlon(base + ".rvh"))
    else:
        with open(base + ".rvt", "w") as f:
            write_gauge(f, *weather(n_days, seed, season_offset))
    _write(base + ".rvi", BLENDED_RVI.format(duration=n_days, routing=routing, catchment_route=catchment_route, extra=extra_rvi, r=["%.6f" % x for x in r]))
    _write(base + ".rvc", BLENDED_RVC)
    return base
