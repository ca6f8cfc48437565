#!/usr/bin/env python
"""Generate the golden fixtures: run the UNMODIFIED reference (oracle/_ref/ref_dump, built from /root/reference by
oracle/build_ref.sh) on synthetic input sets and store its in-memory FP64 results.

Only runs where the reference tree is available (the build container).  The input files written next to each
fixture are our own synthetic inputs (ravenhydroframework_b200.synth); the .npz holds reference OUTPUTS:
  state_steps / state   HRU state [nrec][nHRU][NS] at the recorded steps (1-based step numbers; 0 = initial)
  qout                  [nsteps][NB] CSubBasin::GetOutflowRate after every step
  wshed                 [nsteps][3]  _CumulInput, _CumulOutput, GetAveragePrecip
  forc_steps / forc     force_struct rows at the recorded steps (reference layout, 39 doubles)
  sv_type / sv_layer    state-variable catalogue (integers, must match exactly)
  order                 routing order _aOrderedSBind (integers, must match exactly)
  proc                  per process: reference process_type id, nconn, first <=8 connections
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
from ravenhydroframework_b200 import synth  # noqa: E402

CASES = {
    "hmets_single": dict(kind="hmets", n_sb=1, hru_per_sb=1, n_days=730, seed=5, single_class=True),
    "hmets_tree": dict(kind="hmets", n_sb=7, hru_per_sb=4, n_days=400, seed=3),
    "hmets_tree31": dict(kind="hmets", n_sb=31, hru_per_sb=13, n_days=300, seed=7),
}
WRITERS = {"hmets": synth.write_hmets}
if hasattr(synth, "write_hbv"):
    WRITERS["hbv"] = synth.write_hbv
    CASES["hbv_single"] = dict(kind="hbv", n_sb=1, hru_per_sb=1, n_days=730, seed=4)
    CASES["hbv_muskingum"] = dict(kind="hbv", n_sb=15, hru_per_sb=6, n_days=400, seed=6, routing="ROUTE_MUSKINGUM")
    CASES["hbv_storagecoeff"] = dict(kind="hbv", n_sb=7, hru_per_sb=3, n_days=200, seed=8, routing="ROUTE_STORAGECOEFF")
    # sparse multi-station interpolation (INTERP_FROM_FILE: 4 nearest of 9 stations per HRU) = the gridded-forcing gather
    CASES["hbv_gauges9"] = dict(kind="hbv", n_sb=5, hru_per_sb=4, n_days=150, seed=12, routing="ROUTE_MUSKINGUM", n_gauges=9)
    # process classes outside the three templates: :SnowMelt, :SnowSqueeze, :Interflow
    CASES["hbv_meltsqueeze"] = dict(kind="hbv", n_sb=3, hru_per_sb=5, n_days=300, seed=14, routing="ROUTE_MUSKINGUM", variant="melt_squeeze_interflow")
    # lateral exchange between HRUs (CmvLatFlush): within-subbasin flush limited by the recipient's room + an INTERBASIN flush
    for _k in "abcdefghijklmn":   # other members of the infiltration / soil evaporation / percolation / baseflow process classes
        CASES["hbv_algo_" + _k] = dict(kind="hbv", n_sb=2, hru_per_sb=4, n_days=220, seed=30 + ord(_k), routing="ROUTE_MUSKINGUM", season_offset=80,
                                       variant="algo:" + _k)
    # lateral mixing (CmvLatEquilibrate): several subbasins (where the reference's pool bookkeeping never reaches its hand-back step) and
    # one subbasin (where it does), the latter with an INTERBASIN mixing of the snow packs as well
    CASES["hbv_equil"] = dict(kind="hbv", n_sb=4, hru_per_sb=7, n_days=200, seed=41, routing="ROUTE_MUSKINGUM", season_offset=60, variant="equilibrate")
    CASES["hbv_equil_1sb"] = dict(kind="hbv", n_sb=1, hru_per_sb=9, n_days=200, seed=42, season_offset=20, variant="equilibrate_interbasin")
    # gravitational snow redistribution along a :LateralConnections table (CmvLatRedistribute), both methods
    CASES["hbv_redist"] = dict(kind="hbv", n_sb=3, hru_per_sb=6, n_days=240, seed=43, routing="ROUTE_MUSKINGUM", variant="redistribute")
    CASES["hbv_redist_thresh"] = dict(kind="hbv", n_sb=3, hru_per_sb=6, n_days=240, seed=44, routing="ROUTE_MUSKINGUM", variant="redistribute_threshold")
    # PET estimated from temperature / day geometry instead of read or monthly (CModel::EstimatePET)
    # INF_UBC (UBCWM infiltration: topsoil deficit, interflow, upper / lower groundwater shares, flash-flood factor) on the four soil stores of GR4J
    CASES["gr4j_inf_ubc"] = dict(kind="gr4j", n_sb=2, hru_per_sb=4, n_days=400, seed=57, rvi_sub=[(r"INF_GR4J ", "INF_UBC  ")],
                                 rvp_extra=":GlobalParameter UBC_GW_SPLIT 0.35\n:GlobalParameter UBC_FLASH_PONDING 12.0\n"
                                           ":SoilParameterList\n  :Parameters, MAX_PERC_RATE, UBC_INFIL_SOIL_DEF\n  :Units, mm/d, mm\n  [DEFAULT], 6.5, 210.0\n:EndSoilParameterList\n")
    # SNOBAL_COLD_CONTENT (Brook90-style energy balance of the pack: cold content, refreeze, liquid holding capacity) with the cold content of
    # new snow added by the precipitation partition; reads day length and the canopy LAI / SAI
    CASES["hbv_snobal_coldcontent"] = dict(kind="hbv", n_sb=2, hru_per_sb=4, n_days=400, seed=58, routing="ROUTE_MUSKINGUM",
                                           rvi_sub=[(r"  :SnowBalance       SNOBAL_SIMPLE_MELT SNOW            SNOW_LIQ\n    :-->Overflow     RAVEN_DEFAULT      SNOW_LIQ        PONDED_WATER\n",
                                                     "  :SnowBalance       SNOBAL_COLD_CONTENT SNOW           PONDED_WATER\n")])
    # SNOBAL_GAWSER (Hinckley et al. 1996): snowfall through NEW_SNOW, refreeze, density-based compaction of SNOW_DEPTH, fixed sublimation
    CASES["hbv_snobal_gawser"] = dict(kind="hbv", n_sb=2, hru_per_sb=4, n_days=400, seed=59, routing="ROUTE_MUSKINGUM",
                                      rvi_sub=[(r"  :SnowBalance       SNOBAL_SIMPLE_MELT SNOW            SNOW_LIQ\n    :-->Overflow     RAVEN_DEFAULT      SNOW_LIQ        PONDED_WATER\n",
                                                "  :SnowBalance       SNOBAL_GAWSER      SNOW            PONDED_WATER\n")])
    # SNOBAL_CRHM_EBSM (energy-budget melt with a minimum snow energy from the daily minimum temperature)
    CASES["hbv_snobal_crhm"] = dict(kind="hbv", n_sb=2, hru_per_sb=4, n_days=400, seed=60, routing="ROUTE_MUSKINGUM",
                                    rvi_sub=[(r"  :SnowBalance       SNOBAL_SIMPLE_MELT SNOW            SNOW_LIQ\n    :-->Overflow     RAVEN_DEFAULT      SNOW_LIQ        PONDED_WATER\n",
                                              "  :SnowBalance       SNOBAL_CRHM_EBSM   SNOW            PONDED_WATER\n")])
    # SOILEVAP_SACSMA (Sacramento soil accounting, six soil stores) on a six-layer variant of the GR4J set-up
    CASES["gr4j_sevap_sacsma"] = dict(kind="gr4j", n_sb=2, hru_per_sb=4, n_days=400, seed=63,
                                      rvi_sub=[(r"SOIL_MULTILAYER  4", "SOIL_MULTILAYER  6"), (r"SOILEVAP_GR4J      PRODUCT_STORE", "SOILEVAP_SACSMA    PRODUCT_STORE")],
                                      rvp_sub=[(r"   SOIL_GW\n:EndSoilClasses", "   SOIL_GW\n   SOIL_LZFS\n   SOIL_ADIMC\n:EndSoilClasses"),
                                               (r"DEFAULT_P,  4, SOIL_PROD, 0.529, SOIL_ROUT, 0.300, SOIL_TEMP, 1.000, SOIL_GW, 1.000,",
                                                "DEFAULT_P,  6, SOIL_PROD, 0.529, SOIL_ROUT, 0.300, SOIL_TEMP, 1.000, SOIL_GW, 1.000, SOIL_LZFS, 0.400, SOIL_ADIMC, 0.250,"),
                                               (r"THIN_P,     4, SOIL_PROD, 0.310, SOIL_ROUT, 0.300, SOIL_TEMP, 1.000, SOIL_GW, 1.000,",
                                                "THIN_P,     6, SOIL_PROD, 0.310, SOIL_ROUT, 0.300, SOIL_TEMP, 1.000, SOIL_GW, 1.000, SOIL_LZFS, 0.300, SOIL_ADIMC, 0.200,")],
                                      rvp_extra=":SoilParameterList\n  :Parameters, UNAVAIL_FRAC\n  :Units, none\n  [DEFAULT], 0.22\n:EndSoilParameterList\n"
                                                ":LandUseParameterList\n  :Parameters, MAX_SAT_AREA_FRAC\n  :Units, none\n  [DEFAULT], 0.12\n  LU_TOWN, 0.07\n:EndLandUseParameterList\n",
                                      rvc_extra=":UniformInitialConditions SOIL[2] 120.0\n:UniformInitialConditions SOIL[3] 90.0\n:UniformInitialConditions SOIL[4] 55.0\n:UniformInitialConditions SOIL[5] 35.0\n")
    CASES["gr4j_pet_oudin"] = dict(kind="gr4j", n_sb=2, hru_per_sb=4, n_days=400, seed=51, pet="PET_OUDIN")
    CASES["gr4j_pet_hamon"] = dict(kind="gr4j", n_sb=2, hru_per_sb=4, n_days=400, seed=52, pet="PET_HAMON")
    CASES["hbv_pet_lintemp"] = dict(kind="hbv", n_sb=2, hru_per_sb=4, n_days=400, seed=53, routing="ROUTE_MUSKINGUM", pet="PET_LINEAR_TEMP",
                                    rvp_extra=":LandUseParameterList\n  :Parameters, PET_LIN_COEFF\n  :Units, mm/d/K\n  [DEFAULT], 0.21\n  LU_OPEN, 0.27\n:EndLandUseParameterList\n")
    CASES["hbv_pet_mohyse"] = dict(kind="hbv", n_sb=2, hru_per_sb=4, n_days=400, seed=54, routing="ROUTE_MUSKINGUM", pet="PET_MOHYSE",
                                   rvp_extra=":GlobalParameter MOHYSE_PET_COEFF 1.15\n")
    # PET from the atmospheric chain: air pressure, relative humidity, clear-sky shortwave radiation with cloud-cover correction
    # (EstimateAirPressure / EstimateRelativeHumidity / CRadiation::EstimateShortwaveRadiation), every estimator option built
    CASES["gr4j_atm_turc"] = dict(kind="gr4j", n_sb=2, hru_per_sb=4, n_days=400, seed=61, pet="PET_TURC_1961",
                                  rvi_extra=":RelativeHumidityMethod RELHUM_MINDEWPT\n")
    CASES["gr4j_atm_makkink"] = dict(kind="gr4j", n_sb=2, hru_per_sb=4, n_days=400, seed=62, pet="PET_MAKKINK_1957",
                                     rvi_extra=":AirPressureMethod AIRPRESS_UBC\n:SWCloudCorrect SW_CLOUD_CORR_DINGMAN\n")
    CASES["hbv_atm_penman33"] = dict(kind="hbv", n_sb=2, hru_per_sb=4, n_days=400, seed=63, routing="ROUTE_MUSKINGUM", pet="PET_PENMAN_SIMPLE33",
                                     rvi_extra=":RelativeHumidityMethod RELHUM_MINDEWPT\n:SWCloudCorrect SW_CLOUD_CORR_ANNANDALE\n",
                                     rvp_extra=":LandUseParameterList\n  :Parameters, RELHUM_CORR\n  :Units, none\n  [DEFAULT], 0.9\n  LU_OPEN, 1.05\n:EndLandUseParameterList\n")
    CASES["hbv_atm_penman39"] = dict(kind="hbv", n_sb=2, hru_per_sb=4, n_days=400, seed=64, routing="ROUTE_MUSKINGUM", pet="PET_PENMAN_SIMPLE39")
    CASES["hbv_atm_vapdeficit"] = dict(kind="hbv", n_sb=2, hru_per_sb=4, n_days=400, seed=65, routing="ROUTE_MUSKINGUM", pet="PET_VAPDEFICIT",
                                       rvi_extra=":RelativeHumidityMethod RELHUM_MINDEWPT\n",
                                       rvp_extra=":LandUseParameterList\n  :Parameters, PET_VAP_COEFF\n  :Units, mm/d/kPa\n  [DEFAULT], 2.4\n  LU_OPEN, 3.1\n:EndLandUseParameterList\n")
    CASES["gr4j_atm_linacre"] = dict(kind="gr4j", n_sb=2, hru_per_sb=4, n_days=400, seed=66, pet="PET_LINACRE",
                                     rvi_extra=":RelativeHumidityMethod RELHUM_MINDEWPT\n:AirPressureMethod AIRPRESS_CONST\n:OW_Evaporation PET_LINACRE\n")
    CASES["hbv_atm_hargreaves"] = dict(kind="hbv", n_sb=2, hru_per_sb=4, n_days=400, seed=67, routing="ROUTE_MUSKINGUM", pet="PET_HARGREAVES",
                                       rvi_extra=":RelativeHumidityMethod RELHUM_MINDEWPT\n",
                                       gauge_extra="  :MonthlyMinTemperature, -16.9, -13.1, -8.0, -2.6, 2.7, 7.0, 9.3, 8.2, 4.1, -1.0, -7.9, -14.2,\n"
                                                   "  :MonthlyMaxTemperature, -6.2, -1.8, 2.9, 8.6, 14.0, 17.8, 20.1, 19.2, 14.3, 8.1, 0.2, -4.7,\n")
    # net radiation: total albedo above / below the canopy, default longwave, static canopy correction of shortwave, and the energy-balance PET
    # and potential-melt methods on them.  PENMAN_COMBINATION is given every parameter it reads: the reference leaves those missing from its
    # required-parameter list at the "not needed" marker instead of autogenerating them (src/ModelParamCheck.cpp:50-54)
    _rad_rvp = (":VegetationParameterList\n  :Parameters, SAI_HT_RATIO, RELATIVE_LAI, RELATIVE_HT, SVF_EXTINCTION, ALBEDO\n  :Units, none, none, none, none, none\n"
                "  [DEFAULT], 0.3, 0.9, 1.0, 0.45, 0.13\n  VEG_LOW, 0.1, 0.7, 0.8, 0.6, 0.18\n:EndVegetationParameterList\n"
                ":LandUseParameterList\n  :Parameters, ROUGHNESS, FOREST_SPARSENESS, PRIESTLEYTAYLOR_COEFF\n  :Units, m, none, none\n  [DEFAULT], 0.02, 0.1, 1.26\n  LU_OPEN, 0.005, 0.3, 1.31\n:EndLandUseParameterList\n"
                ":SoilParameterList\n  :Parameters, ALBEDO_WET, ALBEDO_DRY\n  :Units, none, none\n  [DEFAULT], 0.12, 0.27\n  TOPSOIL, 0.11, 0.29\n:EndSoilParameterList\n")
    CASES["hbv_atm_priestley"] = dict(kind="hbv", n_sb=2, hru_per_sb=4, n_days=400, seed=101, routing="ROUTE_MUSKINGUM", pet="PET_PRIESTLEY_TAYLOR",
                                      rvi_extra=":RelativeHumidityMethod RELHUM_MINDEWPT\n:OW_Evaporation PET_PRIESTLEY_TAYLOR\n")
    CASES["hbv_atm_penmancomb"] = dict(kind="hbv", n_sb=2, hru_per_sb=4, n_days=400, seed=102, routing="ROUTE_MUSKINGUM", pet="PET_PENMAN_COMBINATION",
                                       rvi_extra=":SWCanopyCorrect SW_CANOPY_CORR_STATIC\n:WindspeedMethod WINDVEL_UBC_MOD\n", rvp_extra=_rad_rvp)
    CASES["gr4j_atm_potmelt_eb"] = dict(kind="gr4j", n_sb=2, hru_per_sb=4, n_days=400, seed=103, potmelt="POTMELT_EB", pet="PET_PRIESTLEY_TAYLOR",
                                        rvi_extra=":SWCanopyCorrect SW_CANOPY_CORR_STATIC\n:RelativeHumidityMethod RELHUM_MINDEWPT\n:WindspeedMethod WINDVEL_UBC_MOD\n",
                                        rvp_extra=":LandUseParameterList\n  :Parameters, ROUGHNESS\n  :Units, m\n  [DEFAULT], 0.003\n:EndLandUseParameterList\n")
    CASES["hbv_atm_potmelt_restricted"] = dict(kind="hbv", n_sb=2, hru_per_sb=4, n_days=400, seed=104, routing="ROUTE_MUSKINGUM", potmelt="POTMELT_RESTRICTED",
                                               pet="PET_PRIESTLEY_TAYLOR", rvi_extra=":LWRadiationMethod LW_RAD_DEFAULT\n")
    # snow sublimation (CmvSublimation), all five algorithms, over the wind-velocity estimators; HBV-EC has no SNOW_TEMP variable (snow
    # temperature = the SNOW_TEMPERATURE global), GR4J carries SNOW_TEMP as state
    _hbv_after = r"(    :-->Overflow     RAVEN_DEFAULT      SNOW_LIQ        PONDED_WATER\n)"
    _gr4j_after = r"( :SnowBalance              SNOBAL_CEMA_NIEGE  SNOW            PONDED_WATER\n)"
    _sub = "  :Sublimation       %-18s SNOW            ATMOSPHERE\n"
    CASES["hbv_atm_sub_sverdrup"] = dict(kind="hbv", n_sb=2, hru_per_sb=4, n_days=400, seed=71, routing="ROUTE_MUSKINGUM", proc_after=(_hbv_after, _sub % "SUBLIM_SVERDRUP"),
                                         rvi_extra=":RelativeHumidityMethod RELHUM_MINDEWPT\n:WindspeedMethod WINDVEL_UBC_MOD\n")
    CASES["gr4j_atm_sub_kuzmin"] = dict(kind="gr4j", n_sb=2, hru_per_sb=4, n_days=400, seed=72, proc_after=(_gr4j_after, _sub % "SUBLIM_KUZMIN"),
                                        rvi_extra=":RelativeHumidityMethod RELHUM_MINDEWPT\n:WindspeedMethod WINDVEL_SQRT\n",
                                        rvp_extra=":GlobalParameter WINDVEL_ICEPT 0.6\n:GlobalParameter WINDVEL_SCALE 0.21\n")
    CASES["hbv_atm_sub_kuchment"] = dict(kind="hbv", n_sb=2, hru_per_sb=4, n_days=400, seed=73, routing="ROUTE_MUSKINGUM", proc_after=(_hbv_after, _sub % "SUBLIM_KUCHMENT_GELFAN"))
    CASES["gr4j_atm_sub_bulkaero"] = dict(kind="gr4j", n_sb=2, hru_per_sb=4, n_days=400, seed=74, proc_after=(_gr4j_after, _sub % "SUBLIM_BULK_AERO"),
                                          rvi_extra=":AirPressureMethod AIRPRESS_UBC\n:RelativeHumidityMethod RELHUM_MINDEWPT\n")
    CASES["hbv_atm_sub_sierra"] = dict(kind="hbv", n_sb=2, hru_per_sb=4, n_days=400, seed=75, routing="ROUTE_MUSKINGUM", proc_after=(_hbv_after, _sub % "SUBLIM_CENTRAL_SIERRA"),
                                       rvi_extra=":RelativeHumidityMethod RELHUM_MINDEWPT\n:WindspeedMethod WINDVEL_UBC_MOD\n",
                                       rvp_extra=":LandUseParameterList\n  :Parameters, MIN_WIND_SPEED, MAX_WIND_SPEED\n  :Units, m/s, m/s\n  [DEFAULT], 0.7, 6.5\n  LU_OPEN, 1.4, 9.0\n:EndLandUseParameterList\n")
    # canopy: interception fractions from LAI (:PrecipIceptFract), evaporation / sublimation at the remaining PET, Rutter; depression abstraction
    _can = r"  :CanopyEvaporation CANEVP_ALL         CANOPY          ATMOSPHERE\n  :CanopySnowEvap    CANEVP_ALL         CANOPY_SNOW     ATMOSPHERE\n"
    _can_max = "  :CanopyEvaporation CANEVP_MAXIMUM     CANOPY          ATMOSPHERE\n  :CanopySnowEvap    SUBLIM_MAXIMUM     CANOPY_SNOW     ATMOSPHERE\n"
    _can_rut = "  :CanopyEvaporation CANEVP_RUTTER      CANOPY          ATMOSPHERE\n  :CanopySnowEvap    CANEVP_MAXIMUM     CANOPY_SNOW     ATMOSPHERE\n"
    _abst = "  :Abstraction       %-18s PONDED_WATER    DEPRESSION\n  :OpenWaterEvaporation OPEN_WATER_EVAP   DEPRESSION      ATMOSPHERE\n"
    _dep = ":LandUseParameterList\n  :Parameters, DEP_MAX, ABST_PERCENT\n  :Units, mm, none\n  [DEFAULT], 14.0, 0.22\n  LU_OPEN, 6.5, 0.4\n:EndLandUseParameterList\n"
    _icf = ":VegetationParameterList\n  :Parameters, RAIN_ICEPT_FACT, SNOW_ICEPT_FACT\n  :Units, none, none\n  [DEFAULT], 0.05, 0.03\n  VEG_LOW, 0.09, 0.07\n:EndVegetationParameterList\n"
    CASES["hbv_can_icept_lai"] = dict(kind="hbv", n_sb=2, hru_per_sb=4, n_days=400, seed=81, routing="ROUTE_MUSKINGUM", rvi_extra=":PrecipIceptFract PRECIP_ICEPT_LAI\n",
                                      rvi_sub=[(_can, _can_max)], rvp_extra=_icf)
    CASES["hbv_can_icept_explai"] = dict(kind="hbv", n_sb=2, hru_per_sb=4, n_days=400, seed=82, routing="ROUTE_MUSKINGUM", rvi_extra=":PrecipIceptFract PRECIP_ICEPT_EXPLAI\n",
                                         rvi_sub=[(_can, _can_rut)])
    CASES["hbv_can_icept_none"] = dict(kind="hbv", n_sb=2, hru_per_sb=4, n_days=300, seed=83, routing="ROUTE_MUSKINGUM", rvi_extra=":PrecipIceptFract PRECIP_ICEPT_NONE\n")
    CASES["hbv_abst_fill"] = dict(kind="hbv", n_sb=2, hru_per_sb=4, n_days=400, seed=84, routing="ROUTE_MUSKINGUM", proc_after=(_hbv_after, _abst % "ABST_FILL"), rvp_extra=_dep)
    CASES["hbv_abst_scs"] = dict(kind="hbv", n_sb=2, hru_per_sb=4, n_days=400, seed=86, routing="ROUTE_MUSKINGUM", proc_after=(_hbv_after, _abst % "ABST_SCS"),
                                 rvp_extra=_dep + ":LandUseParameterList\n  :Parameters, SCS_CN, SCS_IA_FRACTION\n  :Units, none, none\n  [DEFAULT], 74.0, 0.05\n  LU_OPEN, 86.0, 0.08\n:EndLandUseParameterList\n")
    CASES["hbv_abst_pdmrof"] = dict(kind="hbv", n_sb=2, hru_per_sb=4, n_days=400, seed=87, routing="ROUTE_MUSKINGUM", proc_after=(_hbv_after, _abst % "ABST_PDMROF"),
                                    rvp_extra=_dep + ":LandUseParameterList\n  :Parameters, PDMROF_B\n  :Units, none\n  [DEFAULT], 0.7\n  LU_OPEN, 1.6\n:EndLandUseParameterList\n")
    # Ontario crop heat units: CHU_ONTARIO evolves CROP_HEAT_UNITS (-1 outside the growing season), SOILEVAP_CHU scales PET by CHU / CHU_MATURITY
    CASES["hbv_chu_ontario"] = dict(kind="hbv", n_sb=2, hru_per_sb=4, n_days=500, seed=88, routing="ROUTE_MUSKINGUM",
                                    proc_after=(_hbv_after, "  :CropHeatUnitEvolve CHU_ONTARIO\n"),
                                    rvi_sub=[(r"  :SoilEvaporation   SOILEVAP_HBV       SOIL\[0\]         ATMOSPHERE\n", "  :SoilEvaporation   SOILEVAP_CHU       SOIL[0]         ATMOSPHERE\n")],
                                    rvp_extra=":VegetationParameterList\n  :Parameters, CHU_MATURITY\n  :Units, none\n  [DEFAULT], 2600.0\n  VEG_LOW, 1900.0\n:EndVegetationParameterList\n",
                                    rvc_extra=":UniformInitialConditions CROP_HEAT_UNITS -1.0\n")
    # :ExchangeFlow between the fast and the slow reservoir (constant two-way mixing flux)
    CASES["hbv_exchangeflow"] = dict(kind="hbv", n_sb=2, hru_per_sb=4, n_days=300, seed=89, routing="ROUTE_MUSKINGUM",
                                     proc_after=(r"(  :CapillaryRise     RISE_HBV           FAST_RESERVOIR  SOIL\[0\]\n)", "  :ExchangeFlow      RAVEN_DEFAULT      FAST_RESERVOIR  SLOW_RESERVOIR\n"),
                                     rvp_extra=":SoilParameterList\n  :Parameters, EXCHANGE_FLOW\n  :Units, mm/d\n  [DEFAULT], 0.0\n  FAST_RES, 0.35\n:EndSoilParameterList\n")
    # the other glacier algorithms: potential melt passed on unchanged + linear glacier storage; UBC melt (GLACIER_CC, no melt under snow) + analytic release
    CASES["hbv_glacier_simple"] = dict(kind="hbv", n_sb=2, hru_per_sb=4, n_days=400, seed=90, routing="ROUTE_MUSKINGUM",
                                       rvi_sub=[(r"GMELT_HBV          GLACIER_ICE     GLACIER", "GMELT_SIMPLE_MELT  GLACIER_ICE     GLACIER"),
                                                (r"GRELEASE_HBV_EC    GLACIER", "GRELEASE_LINEAR    GLACIER")])
    CASES["hbv_glacier_ubc"] = dict(kind="hbv", n_sb=2, hru_per_sb=4, n_days=400, seed=91, routing="ROUTE_MUSKINGUM",
                                    rvi_sub=[(r"GMELT_HBV          GLACIER_ICE     GLACIER", "GMELT_UBC          GLACIER_ICE     PONDED_WATER"),
                                             (r"GRELEASE_HBV_EC    GLACIER", "GRELEASE_LINEAR_ANALYTIC GLACIER")],
                                    rvc_extra=":UniformInitialConditions GLACIER_CC 6.0\n")
    # what leaves the depressions again: threshold-power / linear / weir overflow to surface water, linear seepage to the fast reservoir
    _dpar = ":LandUseParameterList\n  :Parameters, DEP_THRESHOLD, DEP_N, DEP_MAX_FLOW, DEP_K, DEP_CRESTRATIO, DEP_SEEP_K\n  :Units, mm, none, mm/d, 1/d, none, 1/d\n  [DEFAULT], 4.0, 1.6, 9.0, 0.18, 0.02, 0.035\n  LU_OPEN, 2.0, 2.1, 14.0, 0.3, 0.05, 0.06\n:EndLandUseParameterList\n"
    for _i, _d in enumerate(("DFLOW_THRESHPOW", "DFLOW_LINEAR", "DFLOW_WEIR")):
        CASES["hbv_depflow_" + _d[6:].lower()] = dict(kind="hbv", n_sb=2, hru_per_sb=4, n_days=400, seed=92 + _i, routing="ROUTE_MUSKINGUM",
            proc_after=(_hbv_after, _abst % "ABST_PERCENTAGE" + "  :DepressionOverflow %-16s DEPRESSION      SURFACE_WATER\n" % _d + ("  :Seepage           SEEP_LINEAR        DEPRESSION      FAST_RESERVOIR\n" if _i == 1 else "")),
            rvp_extra=_dep + _dpar)
    # OPEN_WATER_RIPARIAN: evaporation from the stream fraction of the HRU (here from the depressions filled by ABST_FILL)
    CASES["hbv_ow_riparian"] = dict(kind="hbv", n_sb=2, hru_per_sb=4, n_days=400, seed=95, routing="ROUTE_MUSKINGUM",
                                    proc_after=(_hbv_after, (_abst % "ABST_FILL").replace("OPEN_WATER_EVAP  ", "OPEN_WATER_RIPARIAN")),
                                    rvp_extra=_dep + ":LandUseParameterList\n  :Parameters, STREAM_FRACTION\n  :Units, none\n  [DEFAULT], 0.35\n  LU_OPEN, 0.6\n:EndLandUseParameterList\n")
    CASES["hbv_abst_percentage"] = dict(kind="hbv", n_sb=2, hru_per_sb=4, n_days=400, seed=85, routing="ROUTE_MUSKINGUM", proc_after=(_hbv_after, _abst % "ABST_PERCENTAGE"),
                                        rvi_sub=[(_can, _can_max)], rvp_extra=_dep)
    # surveyed channel cross-section (:ChannelProfile with roughness zones -> 20-point rating curves) under the two routing methods that
    # read the channel geometry; _AUTO / RELATIVE_LAI / :SeasonalCanopyLAI / :SnowParameterList spellings of the LOTW set-up
    _prof = (":ChannelProfile CHN\n  :Bedslope 0.0011\n  :SurveyPoints\n    0 103.2\n    6.0 101.1\n    6.0 100.4\n    9.5 100.0\n    16.0 100.15\n    16.001 101.3\n    41.0 102.1\n    60.0 104.0\n  :EndSurveyPoints\n"
             "  :RoughnessZones\n    0 0.08\n    5.0 0.034\n    17.5 0.061\n    38.0 0.1\n  :EndRoughnessZones\n:EndChannelProfile\n")
    _lotw_rvp = (":VegetationParameterList\n  :Parameters, RELATIVE_LAI, MAX_CAPACITY\n  :Units, none, mm\n  [DEFAULT], 0.8, _AUTO\n  VEG_LOW, _AUTO, 2.5\n:EndVegetationParameterList\n"
                 ":SeasonalCanopyLAI\n  VEG_ALL 0.5 0.5 0.5 0.5 0.69 0.88 1.0 1.0 0.75 0.5 0.5 0.5\n:EndSeasonalCanopyLAI\n"
                 ":SnowParameterList\n  :Parameters, REFREEZE_FACTOR\n  :Units, mm/d/K\n  [DEFAULT], 2.1\n:EndSnowParams\n")
    CASES["hbv_chanprofile_hydrologic"] = dict(kind="hbv", n_sb=7, hru_per_sb=3, n_days=250, seed=91, routing="ROUTE_HYDROLOGIC", season_offset=50,
                                               rvp_sub=[(r":TrapezoidalChannel CHN[^\n]*\n", _prof)], rvp_extra=_lotw_rvp)
    CASES["hbv_chanprofile_muskingum"] = dict(kind="hbv", n_sb=7, hru_per_sb=3, n_days=250, seed=92, routing="ROUTE_MUSKINGUM", season_offset=70,
                                              rvp_sub=[(r":TrapezoidalChannel CHN[^\n]*\n", _prof)])
    # potential melt: degree day with refreeze / rain-on-snow variants (CModel::EstimatePotentialMelt)
    CASES["gr4j_potmelt_ddfreeze"] = dict(kind="gr4j", n_sb=2, hru_per_sb=3, n_days=300, seed=55, potmelt="POTMELT_DD_FREEZE",
                                          rvp_extra=":LandUseParameterList\n  :Parameters, REFREEZE_FACTOR\n  :Units, mm/d/K\n  [DEFAULT], 1.3\n:EndLandUseParameterList\n")
    CASES["hbv_potmelt_ros"] = dict(kind="hbv", n_sb=2, hru_per_sb=4, n_days=300, seed=57, routing="ROUTE_MUSKINGUM", potmelt="POTMELT_HBV_ROS",
                                    rvp_extra=":LandUseParameterList\n  :Parameters, RAIN_MELT_MULT\n  :Units, none\n  [DEFAULT], 1.4\n:EndLandUseParameterList\n")
    # ITER_HEUN (src/Solvers.cpp:304-397).  The reference executable only survives the FIRST step of such a run: its rate1 / rate2 scratch
    # arrays are sized with a per-call `maxConns` that is computed on the first call only (src/Solvers.cpp:30,116-118,313-314), so the
    # second step writes past zero-length allocations and glibc aborts.  The cases therefore pin one step each, from different seasons.
    for _i, (_crit, _so) in enumerate([("0.01 30", 150), ("1e-7 6", 130), ("0.001 30", 200), ("0.05 3", 100)]):
        CASES["hbv_heun_%d" % _i] = dict(kind="hbv", n_sb=3, hru_per_sb=6, n_days=1, seed=14 + _i, routing="ROUTE_MUSKINGUM", season_offset=_so,
                                         rvi_sub=[(r":Method\s+ORDERED_SERIES", ":Method                ITER_HEUN " + _crit)])
    # reservoirs at subbasin outlets (CReservoir level-pool routing): lake type behind a weir with a linked HRU (its surface water is
    # precipitation on the lake, its open-water PET the lake evaporation) + an extraction series + an initial stage; lookup-table
    # (:StageRelations), power-law and volume / area-curve-overridden lake reservoirs, one of them at the outlet
    _ext = ":ReservoirExtraction 5\n  2000-01-01 00:00:00 1.0 %d\n" % 260 + "".join("  %.3f\n" % (0.05 + 0.04 * ((i * 7) % 11) / 11.0) for i in range(260)) + ":EndReservoirExtraction\n"
    CASES["hbv_res_lake"] = dict(kind="hbv", n_sb=7, hru_per_sb=3, n_days=260, seed=95, routing="ROUTE_MUSKINGUM", season_offset=60,
                                 rvh_extra=":Reservoir LakeA\n  :SubBasinID 2\n  :HRUID 5\n  :WeirCoefficient 0.6\n  :CrestWidth 4.5\n  :MaxDepth 6\n  :LakeArea 1.4e6\n:EndReservoir\n"
                                           ":Reservoir LakeB\n  :SubBasinID 5\n  :HRUID 13\n  :WeirCoefficient 0.55\n  :CrestWidth 2.0\n  :MaxDepth 4\n  :LakeArea 0.9e6\n  :AbsoluteCrestHeight 312.5\n:EndReservoir\n",
                                 rvt_extra=_ext, rvc_extra=":InitialReservoirStage 5 312.9\n:InitialReservoirStage 2 0.35\n")
    CASES["hbv_res_tables"] = dict(kind="hbv", n_sb=7, hru_per_sb=3, n_days=260, seed=96, routing="ROUTE_DIFFUSIVE_WAVE", season_offset=90,
                                   rvh_extra=":Reservoir Tab\n  :SubBasinID 3\n  :StageRelations\n    6\n    100.0 0.0 0.0 4.0e5\n    101.0 0.0 4.2e5 4.4e5\n    102.0 1.5 8.8e5 4.8e5\n"
                                             "    103.0 6.0 1.38e6 5.2e5 \n    104.0 15.0 1.92e6 5.6e5\n    106.0 50.0 3.1e6 6.2e5\n  :EndStageRelations\n:EndReservoir\n"
                                             ":Reservoir Pow\n  :SubBasinID 1\n  :MaxDepth 8\n  :VolumeStageRelation POWER_LAW\n    2.4e6 1.6\n  :EndVolumeStageRelation\n"
                                             "  :OutflowStageRelation POWER_LAW\n    7.5 1.5\n  :EndOutflowStageRelation\n:EndReservoir\n"
                                             ":Reservoir Over\n  :SubBasinID 6\n  :HRUID 16\n  :WeirCoefficient 0.7\n  :CrestWidth 3.0\n  :MaxDepth 5\n  :LakeArea 1.1e6\n"
                                             "  :VolumeStageRelation LOOKUP_TABLE\n    4\n    -5.0 0.0\n    -2.0 2.0e6\n    0.0 4.1e6\n    5.0 1.02e7\n  :EndVolumeStageRelation\n"
                                             "  :AreaStageRelation LOOKUP_TABLE\n    3\n    -5.0 5.0e5\n    0.0 1.1e6\n    5.0 1.3e6\n  :EndAreaStageRelation\n:EndReservoir\n",
                                   rvc_extra=":InitialReservoirStage 3 101.6\n")
    # ROUTE_DIFFUSIVE_VARY: the routing hydrograph of every reach is rebuilt each step from a history of flow-dependent celerities
    CASES["hbv_diffusive_vary"] = dict(kind="hbv", n_sb=15, hru_per_sb=4, n_days=260, seed=97, routing="ROUTE_DIFFUSIVE_VARY", season_offset=70)
    # direct-insertion streamflow assimilation (src/Assimilate.cpp): gauges at the outlet and at an interior basin, blanks in both series,
    # corrections propagated upstream with a distance decay; the second case starts assimilating mid-run
    def _obs(sb, n, seed, blanks):
        r = np.random.RandomState(seed)
        v = ["-1.2345" if (i in blanks) else "%.4f" % (0.6 + 2.5 * abs(np.sin(i / 17.0 + seed)) + 0.3 * r.rand()) for i in range(n + 1)]
        return ":ObservationData HYDROGRAPH %d m3/s\n  2000-01-01 00:00:00 1.0 %d\n  " % (sb, n + 1) + "\n  ".join(v) + "\n:EndObservationData\n:AssimilateStreamflow %d\n" % sb
    CASES["hbv_assim"] = dict(kind="hbv", n_sb=7, hru_per_sb=3, n_days=200, seed=98, routing="ROUTE_MUSKINGUM", season_offset=60,
                              rvi_extra=":AssimilateStreamflow\n", rvp_extra=":GlobalParameter ASSIM_UPSTREAM_DECAY 0.04\n:GlobalParameter ASSIMILATION_FACT 0.8\n",
                              rvt_extra=_obs(1, 200, 1, set(range(40, 55)) | {120}) + _obs(3, 200, 2, set(range(0, 5)) | set(range(90, 140))))
    CASES["hbv_assim_late"] = dict(kind="hbv", n_sb=15, hru_per_sb=2, n_days=150, seed=99, routing="ROUTE_DIFFUSIVE_WAVE", season_offset=100,
                                   rvi_extra=":AssimilateStreamflow\n:AssimilationStartTime 2000-02-10 00:00:00\n",
                                   rvt_extra=_obs(2, 150, 3, {60, 61}) + _obs(5, 150, 4, set()) + _obs(1, 150, 5, set(range(100, 151))))
    CASES["hbv_hrugroup_cond"] = dict(kind="hbv", n_sb=3, hru_per_sb=5, n_days=250, seed=12, routing="ROUTE_MUSKINGUM", season_offset=40, variant="hrugroup_cond")
    CASES["hbv_lateral"] = dict(kind="hbv", n_sb=4, hru_per_sb=7, n_days=250, seed=8, routing="ROUTE_MUSKINGUM", season_offset=40, variant="lateral")
if hasattr(synth, "write_blended"):
    # Blended multi-model template (BASELINE config 4): weighted process groups, VIC/TOPMODEL/analytic algorithms, SNOBAL_HBV
    WRITERS["blended"] = synth.write_blended
    CASES["blended_tree"] = dict(kind="blended", n_sb=5, hru_per_sb=4, n_days=300, seed=16, routing="ROUTE_MUSKINGUM", catchment_route="TRIANGULAR_UH")
    CASES["blended_single"] = dict(kind="blended", n_sb=1, hru_per_sb=1, n_days=500, seed=17)
    # sub-daily steps (BASELINE config 4 asks for hourly): daily min/max temperatures downscaled by the gauge, daily items
    # carried through the day, Muskingum routing with sub-daily histories
    CASES["blended_hourly"] = dict(kind="blended", n_sb=5, hru_per_sb=4, n_days=25, seed=18, routing="ROUTE_MUSKINGUM", catchment_route="TRIANGULAR_UH",
                                   season_offset=95, n_gauges=9, timestep="01:00:00")
    CASES["hbv_6hourly"] = dict(kind="hbv", n_sb=7, hru_per_sb=3, n_days=120, seed=19, routing="ROUTE_MUSKINGUM", season_offset=60, timestep="06:00:00")
    # explicit Euler (:Method EULER): every process sees the start-of-step state
    CASES["hbv_diffusive"] = dict(kind="hbv", n_sb=7, hru_per_sb=3, n_days=200, seed=22, routing="ROUTE_DIFFUSIVE_WAVE", season_offset=70)
    CASES["hbv_hydrologic"] = dict(kind="hbv", n_sb=7, hru_per_sb=3, n_days=250, seed=24, routing="ROUTE_HYDROLOGIC", season_offset=50)
    CASES["hbv_euler"] = dict(kind="hbv", n_sb=4, hru_per_sb=5, n_days=300, seed=9, routing="ROUTE_MUSKINGUM", season_offset=30, method="EULER")
    # SUBDAILY_SIMPLE: the day's PET (daily estimators) and degree-day melt distributed over the steps by a half-sine between dawn and dusk
    CASES["hbv_3hourly_subdaily"] = dict(kind="hbv", n_sb=3, hru_per_sb=4, n_days=90, seed=21, routing="ROUTE_MUSKINGUM", season_offset=70, timestep="03:00:00",
                                         rvi_extra=":SubdailyMethod SUBDAILY_SIMPLE\n")
    CASES["hmets_3hourly"] = dict(kind="hmets", n_sb=3, hru_per_sb=4, n_days=60, seed=20, season_offset=90, timestep="03:00:00")
if hasattr(synth, "write_gr4j"):
    WRITERS["gr4j"] = synth.write_gr4j
    CASES["gr4j_single"] = dict(kind="gr4j", n_sb=1, hru_per_sb=1, n_days=730, seed=9)
    CASES["gr4j_tree"] = dict(kind="gr4j", n_sb=7, hru_per_sb=5, n_days=300, seed=10)


# Monte-Carlo ensemble (reference CMonteCarloEnsemble): fixture = ensemble member MEMBER as the reference's own member
# loop would run it (ref_dump replays UpdateModel for the members before it), stored as reference_member<k>.npz
MC_CASES = {"hmets_mc": dict(kind="hmets", n_sb=3, hru_per_sb=4, n_days=150, seed=7, ensemble_members=4, mixed_distributions=True, member=2)}


# EnKF ensembles (reference CEnKFEnsemble): one assimilation window run by the UNMODIFIED reference through
# oracle/_ref/ref_enkf_dump; fixture = reference_enkf.npz with the observation / noise / output matrices, the forcing
# perturbation samples, every member's final state and the state matrix before and after the reference's AssimilationCalcs.
ENKF_CASES = {
    "enkf_hbv": dict(kind="hbv", n_sb=7, hru_per_sb=4, n_days=40, seed=21, routing="ROUTE_MUSKINGUM", season_offset=120,
                     enkf=dict(n_members=6, obs_basins=(1, 3), window=2)),
    "enkf_hmets": dict(kind="hmets", n_sb=3, hru_per_sb=5, n_days=30, seed=23, routing="ROUTE_NONE", season_offset=100,
                       enkf=dict(n_members=5, obs_basins=(1,), window=1, assimilated=("SOIL[1]", "SNOW_LIQ"), hru_group_range=(3, 12),
                                 perturbations=("PRECIP DIST_NORMAL 0.5 1.0 ADDITIVE", "SNOWFALL DIST_UNIFORM 0.5 1.5 MULTIPLICATIVE",
                                                "TEMP_AVE DIST_LOGNORMAL 0.0 0.1 MULTIPLICATIVE AssimHRUs"),
                                 obs_error="STREAMFLOW DIST_GAMMA 2.0 0.3 ADDITIVE", obs_level=3.0)),
    # the config-5 shape at a size the reference finishes in seconds: 64 members, 12 gauged outlets x window 3 (the reference's slot
    # quirk leaves 24 + 1 distinct observation slots), 31 subbasins -- pins the truncated-SVD solve of an ill-conditioned P
    "enkf_hbv_n64": dict(kind="hbv", n_sb=31, hru_per_sb=3, n_days=30, seed=29, routing="ROUTE_MUSKINGUM", season_offset=120,
                         enkf=dict(n_members=64, obs_basins=(1, 2, 3, 4, 5, 6, 7, 9, 11, 13, 17, 23), window=3,
                                   obs_error="STREAMFLOW DIST_NORMAL 1 0.0002 MULTIPLICATIVE")),
}


def make_enkf(name, spec):
    import subprocess
    spec = dict(spec)
    kind = spec.pop("kind")
    en = spec.pop("enkf")
    d = os.path.join(HERE, name)
    shutil.rmtree(d, ignore_errors=True)
    base = WRITERS[kind](d, name="model", **spec)
    synth.add_enkf(base, n_sb=spec["n_sb"], n_hru=spec["n_sb"] * spec["hru_per_sb"], n_days=spec["n_days"], **en)
    out = os.path.join(d, "_refout") + "/"
    os.makedirs(out, exist_ok=True)
    r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_enkf_dump"), "model", out], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and os.path.exists(out + "enkf_Xpost.f64"), r.stdout[-3000:]
    meta = dict(l.split()[:2] for l in open(out + "enkf_meta.txt") if not l.startswith("ROW"))
    N, M, Nobs, nH, NS, NB, nsteps, nP = [int(meta[k]) for k in ("N", "M", "Nobs", "nHRU", "NS", "NB", "nsteps", "nPert")]
    ld = lambda f, *shape: np.fromfile(out + f).reshape(*shape)  # noqa: E731
    np.savez_compressed(os.path.join(d, "reference_enkf.npz"), obs=ld("enkf_obs.f64", N, Nobs), noise=ld("enkf_noise.f64", N, Nobs),
                        out=ld("enkf_out.f64", N, Nobs), Xpre=ld("enkf_Xpre.f64", N, M), Xpost=ld("enkf_Xpost.f64", N, M),
                        state=ld("enkf_state.f64", N, nH, NS), qout_member0=ld("enkf_qout.f64", N, nsteps, NB)[0],
                        qout_last=ld("enkf_qout.f64", N, nsteps, NB)[:, -1, :], eps=ld("enkf_eps.f64", N, nsteps, nP),
                        row_names=np.array([l.split()[2] for l in open(out + "enkf_meta.txt") if l.startswith("ROW")]))
    shutil.rmtree(out, ignore_errors=True)
    for junk in ("ens_1", "ens_2"):
        shutil.rmtree(os.path.join(d, junk), ignore_errors=True)
    print("golden %-18s N=%d M=%d Nobs=%d nHRU=%d NB=%d steps=%d nPert=%d" % (name, N, M, Nobs, nH, NB, nsteps, nP))


def make_restart(name="hbv_restart", n1=120, n2=100):
    """Warm start from a checkpoint written by the reference EXECUTABLE: Raven.exe runs n1 days and writes solution.rvc (HRU state
    table + :BasinStateVariables, 5 decimals); the fixture model starts at day n1 from that file and the reference runs n2 more days."""
    import datetime
    import subprocess
    d = os.path.join(HERE, name)
    shutil.rmtree(d, ignore_errors=True)
    base = synth.write_hbv(d, n_sb=7, hru_per_sb=3, n_days=n1 + n2, seed=26, routing="ROUTE_MUSKINGUM", name="model", season_offset=20)
    rvi0 = open(base + ".rvi").read()
    open(base + ".rvi", "w").write(rvi0.replace(":Duration              %d" % (n1 + n2), ":Duration              %d" % n1).replace(":SilentMode", ":SilentMode\n:SuppressOutput"))
    out = os.path.join(d, "_first") + "/"
    # :SuppressOutput also suppresses the solution file: run without it, quietly
    open(base + ".rvi", "w").write(rvi0.replace(":Duration              %d" % (n1 + n2), ":Duration              %d" % n1))
    r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "Raven.exe"), "model", "-o", out], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert os.path.exists(out + "solution.rvc"), r.stdout[-2000:]
    shutil.copy(out + "solution.rvc", base + ".rvc")
    shutil.rmtree(out, ignore_errors=True)
    start = datetime.date(2000, 1, 1) + datetime.timedelta(days=n1)
    rvi = rvi0.replace(":StartDate             2000-01-01 00:00:00", ":StartDate             %s 00:00:00" % start.isoformat())
    rvi = rvi.replace(":Duration              %d" % (n1 + n2), ":Duration              %d" % n2)
    assert rvi != rvi0
    open(base + ".rvi", "w").write(rvi)
    outd = os.path.join(d, "_refout")
    ref = H.reference_run(base, outd, n2)
    rec = sorted(set([1, 2, 3, 10, 50, n2]))
    np.savez_compressed(os.path.join(d, "reference.npz"), state_steps=np.array(rec), state=ref["state"][rec], qout=ref["basin"][1:n2 + 1, :, 0],
                        basin_last=ref["basin"][n2], wshed=ref["wshed"][1:n2 + 1], forc_steps=np.array(rec), forc=ref["forc"][rec],
                        sv_type=np.array([t for t, _ in ref["sv"]]), sv_layer=np.array([l for _, l in ref["sv"]]),
                        order=np.array(ref["order"]), proc=np.array([" ".join(p) for p in ref["proc"]]),
                        param=np.array([" ".join(p) for p in ref["param"]]))
    shutil.rmtree(outd, ignore_errors=True)
    print("golden %-18s warm start at day %d, %d steps" % (name, n1, n2))


def make_dds(name="dds_hmets"):
    """DDS calibration (reference CDDSEnsemble): the reference EXECUTABLE runs all trials; fixture = every trial's objective function
    value as printed (6 digits), the rows of DDSOutput.csv and the best parameter vector."""
    import json
    import re
    import subprocess
    d = os.path.join(HERE, name)
    shutil.rmtree(d, ignore_errors=True)
    base = synth.write_hmets(d, n_sb=3, hru_per_sb=4, n_days=200, seed=6, name="model", season_offset=60)
    rvi = open(base + ".rvi").read().replace(":SilentMode", ":SilentMode\n:EnsembleMode ENSEMBLE_DDS 25\n:RandomSeed 7\n:EvaluationMetrics NASH_SUTCLIFFE")
    open(base + ".rvi", "w").write(rvi)
    rng = np.random.RandomState(2)
    with open(base + ".rvt", "a") as f:
        f.write(":ObservationData HYDROGRAPH 1 m3/s\n  2000-01-01 00:00:00 1.0 201\n")
        for i in range(201):
            f.write("  %.4f\n" % (1.5 + 1.2 * np.sin(i / 20.0) ** 2 + 0.2 * rng.rand()))
        f.write(":EndObservationData\n")
    with open(base + ".rve", "w") as f:
        f.write(":ObjectiveFunction 1 NASH_SUTCLIFFE\n:ParameterDistributions\n")
        f.write("  GAMMA_SHAPE LANDUSE FOREST 9.5019 DIST_UNIFORM 3.0 15.0\n  GAMMA_SCALE LANDUSE FOREST 0.2774 DIST_UNIFORM 0.1 0.9\n")
        f.write("  MAX_MELT_FACTOR LANDUSE FOREST 6.7009 DIST_UNIFORM 4.0 9.0\n  PERC_COEFF SOIL TOPSOIL 0.0114 DIST_UNIFORM 0.002 0.05\n")
        f.write("  BASEFLOW_COEFF SOIL PHREATIC 0.0069 DIST_UNIFORM 0.001 0.03\n:EndParameterDistributions\n")
    out = os.path.join(d, "_refout") + "/"
    r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "Raven.exe"), "model", "-o", out], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    trials = [[float(x) for x in re.findall(r"DDS Obj. Function: ([-0-9.eE+]+) \[best: ([-0-9.eE+]+)\]", l)[0]] for l in r.stdout.split("\n") if "DDS Obj" in l]
    txt = open(out + "DDSOutput.csv").read().split("Best parameter vector:")
    rows = [[float(x) for x in l.strip().strip(",").split(",")] for l in txt[0].strip().split("\n")]
    best = [float(l.split(",")[1]) for l in txt[1].strip().split("\n")]
    json.dump(dict(trials=trials, improvements=rows, best=best), open(os.path.join(d, "reference_dds.json"), "w"), indent=1)
    shutil.rmtree(out, ignore_errors=True)
    print("golden %-18s trials=%d improvements=%d" % (name, len(trials), len(rows)))


HYDROGRAPH_CASES = ("hbv_assim", "hbv_muskingum", "hmets_tree")


def make_hydrographs(name):
    """Hydrographs.csv as the reference EXECUTABLE writes it for an existing fixture's inputs (text, 6 significant digits)."""
    import subprocess
    d = os.path.join(HERE, name)
    out = os.path.join(d, "_exe") + "/"
    r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "Raven.exe"), "model", "-o", out], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert os.path.exists(out + "Hydrographs.csv"), r.stdout[-2000:]
    shutil.copy(out + "Hydrographs.csv", os.path.join(d, "reference_Hydrographs.csv"))
    shutil.rmtree(out, ignore_errors=True)
    print("golden %-18s Hydrographs.csv (%d lines)" % (name, len(open(os.path.join(d, "reference_Hydrographs.csv")).readlines())))


FLUX_CASES = ("hbv_muskingum", "hbv_algo_d", "hbv_hrugroup_cond", "hbv_snobal_crhm")   # no :Convolve


def make_cumflux(name):
    """reference_cumflux.npz: CModel::GetCumulativeFlux of every (HRU, state variable) at the end of the golden case's run."""
    d = os.path.join(HERE, name)
    n = int(np.load(os.path.join(d, "reference.npz"))["qout"].shape[0])
    out = os.path.join(d, "_refout")
    ref = H.reference_run(os.path.join(d, "model"), out, n)
    np.savez_compressed(os.path.join(d, "reference_cumflux.npz"), cumflux=ref["cumflux"], nsteps=np.array(n))
    shutil.rmtree(out, ignore_errors=True)
    print("golden %-18s cumulative fluxes of %d steps" % (name, n))


def make(name, spec):
    spec = dict(spec)
    kind = spec.pop("kind")
    member = spec.pop("member", None)
    timestep = spec.pop("timestep", None)     # e.g. "01:00:00": sub-daily model step on the daily forcing series
    method = spec.pop("method", None)         # :Method (numerical method) other than ORDERED_SERIES
    pet = spec.pop("pet", None)               # :Evaporation method other than the template's
    potmelt = spec.pop("potmelt", None)       # :PotentialMeltMethod other than the template's
    rvp_extra = spec.pop("rvp_extra", None)   # lines appended to the .rvp
    rvi_extra = spec.pop("rvi_extra", None)   # option commands inserted at the top of the .rvi
    rvi_sub = spec.pop("rvi_sub", None)   # [(regex, replacement)] applied to the .rvi
    rvp_sub = spec.pop("rvp_sub", None)   # [(regex, replacement)] applied to the .rvp
    proc_after = spec.pop("proc_after", None)   # (regex of a process line, line inserted after it) in the .rvi process list
    gauge_extra = spec.pop("gauge_extra", None)   # lines added to the gauge block of the .rvt (after :MonthlyAveTemperature)
    rvh_extra = spec.pop("rvh_extra", None)   # lines appended to the .rvh / .rvt / .rvc
    rvt_extra = spec.pop("rvt_extra", None)
    rvc_extra = spec.pop("rvc_extra", None)
    d = os.path.join(HERE, name)
    shutil.rmtree(d, ignore_errors=True)
    base = WRITERS[kind](d, name="model", **spec)
    n = spec["n_days"]
    if method is not None:
        rvi = open(base + ".rvi").read()
        assert ":Method                ORDERED_SERIES" in rvi
        open(base + ".rvi", "w").write(rvi.replace(":Method                ORDERED_SERIES", ":Method                " + method))
    if pet is not None:
        import re
        rvi = open(base + ".rvi").read()
        rvi, cnt = re.subn(r":Evaporation\s+\S+", ":Evaporation           " + pet, rvi)
        assert cnt == 1
        open(base + ".rvi", "w").write(rvi)
    if potmelt is not None:
        import re
        rvi = open(base + ".rvi").read()
        rvi, cnt = re.subn(r":PotentialMeltMethod\s+\S+", ":PotentialMeltMethod   " + potmelt, rvi)
        assert cnt == 1
        open(base + ".rvi", "w").write(rvi)
    if proc_after is not None:
        import re
        rvi = open(base + ".rvi").read()
        rvi, cnt = re.subn(proc_after[0], lambda m: m.group(1) + proc_after[1], rvi)
        assert cnt == 1, proc_after[0]
        open(base + ".rvi", "w").write(rvi)
    for pat, rep in (rvi_sub or []):
        import re
        rvi = open(base + ".rvi").read()
        rvi, cnt = re.subn(pat, lambda m: rep, rvi)
        assert cnt == 1, pat
        open(base + ".rvi", "w").write(rvi)
    for pat, rep in (rvp_sub or []):
        import re
        rvp = open(base + ".rvp").read()
        rvp, cnt = re.subn(pat, lambda m: rep, rvp)
        assert cnt == 1, pat
        open(base + ".rvp", "w").write(rvp)
    if rvi_extra is not None:
        rvi = open(base + ".rvi").read()
        open(base + ".rvi", "w").write(rvi_extra + rvi)
    if gauge_extra is not None:
        import re
        rvt = open(base + ".rvt").read()
        rvt, cnt = re.subn(r"(\s*:MonthlyAveTemperature[^\n]*\n)", lambda m: m.group(1) + gauge_extra, rvt)
        assert cnt == 1
        open(base + ".rvt", "w").write(rvt)
    if rvp_extra is not None:
        open(base + ".rvp", "a").write(rvp_extra)
    for ext, txt in ((".rvh", rvh_extra), (".rvt", rvt_extra), (".rvc", rvc_extra)):
        if txt is not None:
            open(base + ext, "a").write(txt)
    if timestep is not None:
        rvi = open(base + ".rvi").read()
        assert ":TimeStep              1.0" in rvi
        open(base + ".rvi", "w").write(rvi.replace(":TimeStep              1.0", ":TimeStep              " + timestep))
        h, mi, se = [float(x) for x in timestep.split(":")]
        n = int(round(n / ((h + mi / 60.0 + se / 3600.0) / 24.0)))
    out = os.path.join(d, "_refout")
    if member is not None:
        os.environ["REF_DUMP_MEMBER"] = str(member)
    try:
        ref = H.reference_run(base, out, n)
    finally:
        os.environ.pop("REF_DUMP_MEMBER", None)
    rec = sorted(set([1, 2, 3, 10, 50, 100, 150, 200, 250, 300, 365, 400, 500, 600, 700, n]) & set(range(1, n + 1)))
    np.savez_compressed(os.path.join(d, "reference.npz" if member is None else "reference_member%d.npz" % member),
                        state_steps=np.array(rec), state=ref["state"][rec], qout=ref["basin"][1:n + 1, :, 0],
                        basin_last=ref["basin"][n], wshed=ref["wshed"][1:n + 1], forc_steps=np.array(rec), forc=ref["forc"][rec],
                        sv_type=np.array([t for t, _ in ref["sv"]]), sv_layer=np.array([l for _, l in ref["sv"]]),
                        order=np.array(ref["order"]), proc=np.array([" ".join(p) for p in ref["proc"]]),
                        param=np.array([" ".join(p) for p in ref["param"]]),
                        **({"res_stage": ref["res"][rec][:, :, 0], "res_losses": ref["res"][rec][:, :, 2]} if np.any(ref["res"][:, :, 3] != 0.0) else {}))
    shutil.rmtree(out, ignore_errors=True)
    print("golden %-18s nHRU=%d NS=%d NB=%d steps=%d" % (name, ref["nHRU"], ref["NS"], ref["NB"], n))


if __name__ == "__main__":
    assert H.have_reference(), "oracle/_ref/ref_dump missing: run oracle/build_ref.sh where /root/reference exists"
    only = sys.argv[1:]
    for name, spec in list(CASES.items()) + list(MC_CASES.items()):
        if only and name not in only:
            continue
        make(name, spec)
    for name in HYDROGRAPH_CASES:
        if not only or name in only:
            make_hydrographs(name)
    for name in FLUX_CASES:
        if not only or name in only or "cumflux" in only:
            make_cumflux(name)
    if not only or "dds_hmets" in only:
        make_dds()
    if not only or "hbv_restart" in only:
        make_restart()
    for name, spec in ENKF_CASES.items():
        if only and name not in only:
            continue
        make_enkf(name, spec)
