// rvn_processes.cuh -- device half of the hydrological-process plugin contract.
//
// Every process of the reference exposes  GetRatesOfChange(state_vars,pHRU,Options,tt,rates)  and
// ApplyConstraints(...)  (src/HydroProcessABC.h:102-112).  Here each supported (process, algorithm)
// pair is one __device__ function template  rates_<NAME>(ctx, pr)  that performs both calls back to back on the
// calling thread's (member, HRU) pair.
//
//   ctx  plays the role of pHRU/Options/tt: newest state vector SV(slot), rate vector R(q), the lagged start-of-step
//        values the reference reads through pHRU->GetStateVarValue, derived forcings, class
//        parameters and HRU attributes.  Two context types exist (rvn_engine.cu):
//          DynCtx        state/rates in shared-memory columns, catalogue lookups at run time  (interpreter kernel)
//          StatCtx<Prog> state/rates in registers, every index a compile-time constant         (specialised kernels)
//   pr   is the process record: connection slots from(q)/to(q), nconn(), integer/double arguments ia(i)/da(i);
//        run-time (DynProc) or constexpr (StatProc<Prog,J>).
//
// Arithmetic order follows the cited reference lines; min/max are the reference's (first argument wins ties), FMA
// contraction is disabled at compile time (-fmad=false).
#ifndef RVN_PROCESSES_CUH
#define RVN_PROCESSES_CUH
#include "rvn_device.cuh"
#include "rvn_glibc_math.h"   // rvn_exp / rvn_log / rvn_pow / rvn_tanh: the host libm of the reference build, bit for bit

#define D_ALMOST_INF 1e99
#define D_REAL_SMALL 1e-12
#define D_PI 3.1415926535898
#define D_SEC_PER_DAY 86400.0
#define D_MM_PER_METER 1000.0
#define D_M2_PER_KM2 1e6
#define D_FREEZING_TEMP 0.0
#define D_NEGLIGBLE_SNOW 0.1
#define D_TYPICAL_SNOW_DENS 0.250e3
#define D_DENSITY_WATER 1.000e3
#define D_DENSITY_ICE 0.917e3
#define D_HCP_ICE 1.938
#define D_LH_VAPOR 2.501
#define D_SPH_WATER 4.186e-3   // [MJ/kg/K]
#define D_LH_FUSION 0.334      // [MJ/kg]
#define D_AMBIENT_AIR_PRESSURE 101.3
#define D_KPA_PER_ATM 101.325
#define D_ZERO_CELSIUS 273.16
#define D_GLOBAL_ALBEDO 0.3
#define D_SPH_AIR 1.012e-3
#define D_AIR_H20_MW_RAT 0.622
#define D_WINTER_SOLSTICE_ANG 6.111043
#define D_HBV_PET_ELEV_CORR 0.0005
#define D_HBV_PET_TEMP_CORR 0.5
#define D_GLOBAL_PET_CORR 1.0

#define RVN_DI __device__ __forceinline__
#define RVN_HD __host__ __device__ __forceinline__

RVN_DI double dmax(double r, double l) { return (r >= l) ? r : l; }   // src/RavenInclude.h:1456
RVN_DI double dmin(double r, double l) { return (r <= l) ? r : l; }   // src/RavenInclude.h:1464
RVN_DI double rvn_sqr(double x) { return x * x; }   // what g++ -O2 makes of the reference's pow(x, 2) / pow(x, 2.0) (no libm call)

// ---- helpers shared by both context types (CRTP-free: plain function templates over the context) -------------
// CHydroUnit::GetSoilCapacity (src/HydroUnits.cpp:260-267)
template <class C> RVN_DI double SoilCapacity(const C &c, int m) { return c.thick(m) * D_MM_PER_METER * c.SOIL(m, RVN_S_CAP_RATIO); }
// CHydroUnit::GetSoilTensionStorageCapacity (src/HydroUnits.cpp:275-286)
template <class C> RVN_DI double SoilTensionCapacity(const C &c, int m) {
  return c.thick(m) * D_MM_PER_METER * c.SOIL(m, RVN_S_CAP_RATIO) * (c.SOIL(m, RVN_S_FIELD_CAPACITY) - c.SOIL(m, RVN_S_SAT_WILT));
}
// CHydroUnit::GetSnowCover (src/HydroUnits.cpp:679-705), lagged, SNOWCOV_NONE
template <class C> RVN_DI double SnowCover(const C &c) {
  if (c.iSV(RVN_SV_SNOW) < 0) return 0.0;
  if (c.iSV(RVN_SV_SNOW_COVER) >= 0) return c.old_snow_cover;
  return (c.old_snow < D_NEGLIGBLE_SNOW) ? 0.0 : 1.0;
}
// CHydroUnit::GetSnowTemperature (src/HydroUnits.cpp:645-658), lagged
template <class C> RVN_DI double SnowTemperature(const C &c) {
  if (c.iSV(RVN_SV_SNOW_TEMP) < 0) { double t = c.G(RVN_G_SNOW_TEMPERATURE); return (t < -60000.0) ? 0.0 : t; }
  return c.old_snow_temp;
}
// CHydroUnit::GetStateVarMax (src/HydroUnits.cpp:559-606); CalculateSnowLiquidCapacity (src/SnowParams.cpp:79-88)
template <class C> RVN_DI double StateVarMax(const C &c, int slot) {
  switch (c.slot_type(slot)) {
    case RVN_SV_SOIL: return SoilCapacity(c, c.slot_layer(slot));
    case RVN_SV_CANOPY: return c.canopy_cap;
    case RVN_SV_CANOPY_SNOW: return c.canopy_snow_cap;
    case RVN_SV_DEPRESSION: { double dm = c.LU(RVN_LU_DEP_MAX); return (dm >= 0) ? dm : D_ALMOST_INF; }
    case RVN_SV_SNOW_COVER: return 1.0;
    case RVN_SV_SNOW_LIQ: return c.G(RVN_G_SNOW_SWI) * c.SV(c.iSV(RVN_SV_SNOW));
    default: return D_ALMOST_INF;
  }
}

// CmvPrecipitation::GetRatesOfChange  src/PartitionPrecip.cpp:135-355 (models without ICE_THICKNESS)
template <class C, class PR> RVN_DI void rates_PRECIPITATION(C &c, const PR &pr) {
  enum { qPond = 0, qCan, qCanSnow, qDep, qAtmos, qSnow, qSnowLiq, qSnowDef, qNewSnow, qLake, qSW, qColdContent, qSnowDepth };
  const int iCan = c.iSV(RVN_SV_CANOPY), iSCan = c.iSV(RVN_SV_CANOPY_SNOW), iSnLiq = c.iSV(RVN_SV_SNOW_LIQ), iSWE = c.iSV(RVN_SV_SNOW);
  const double tstep = c.tstep;
  double total_precip = c.F[RVN_F_PRECIP], Fsnow = c.F[RVN_F_SNOW_FRAC], air_temp = c.F[RVN_F_TEMP_AVE];
  if (iSWE < 0) Fsnow = 0;
  double snowfall = (Fsnow)*total_precip;
  double rainfall = (1.0 - Fsnow) * total_precip;
  rainfall += 0.0;  // irrigation
  double SWE = 0.0;
  if (iSWE >= 0) SWE = c.SV(iSWE);
  const double Fs = SnowCover(c);
  double rainthru, snowthru;
  const bool is_lake = (c.type == RVN_HRU_LAKE);
  if (!is_lake) {
    const double Fcan = c.LU(RVN_LU_FOREST_COVERAGE);
    double rain_int = rainfall * c.rain_icept_pct;
    if (iCan >= 0) {
      double capacity = c.canopy_cap;
      rain_int = dmin(rain_int, dmax((capacity - c.SV(iCan)) / tstep, 0.0));
      c.R(qCan) = Fcan * rain_int;
    } else c.R(qAtmos) += Fcan * rain_int;
    rainthru = rainfall - Fcan * rain_int;
    double snow_int = snowfall * c.snow_icept_pct;
    if (iSCan >= 0) {
      double snow_capacity = c.canopy_snow_cap;
      snow_int = dmax(dmin(snow_int, dmax((snow_capacity - c.SV(iSCan)) / tstep, 0.0)), 0.0);
      c.R(qCanSnow) = Fcan * snow_int;
    } else c.R(qAtmos) += Fcan * snow_int;
    snowthru = snowfall - Fcan * snow_int;
  } else { snowthru = snowfall; rainthru = rainfall; }
  if (is_lake) { if (c.linked) c.R(qSW) = snowthru + rainthru; else c.R(qLake) = snowthru + rainthru; }   // src/PartitionPrecip.cpp:258-265
  else if (c.type == RVN_HRU_WETLAND) c.R(qDep) = snowthru + rainthru;
  else if (c.type == RVN_HRU_WATER) c.R(qSW) = snowthru + rainthru;
  else {
    if (c.iSV(RVN_SV_NEW_SNOW) < 0) {
      if (iSWE >= 0) {
        c.R(qSnow) = snowthru;
        if (c.iSV(RVN_SV_SNOW_DEPTH) >= 0) {  // CalcFreshSnowDensity, src/SnowParams.cpp:37-47
          double rho = (air_temp <= D_FREEZING_TEMP) ? 67.92 + 51.25 * rvn_exp((air_temp - D_FREEZING_TEMP) / 2.59)
                                                     : dmin((119.17 + 20.0 * (air_temp - D_FREEZING_TEMP)), 200.0);
          c.R(qSnowDepth) = snowthru * D_DENSITY_ICE / rho;
        }
        if (c.iSV(RVN_SV_COLD_CONTENT) >= 0) c.R(qColdContent) = snowthru * dmax(D_FREEZING_TEMP - air_temp, 0.0) * D_HCP_ICE / D_MM_PER_METER;
        if ((iSnLiq >= 0) && (c.iSV(RVN_SV_SNOW_DEFICIT) < 0) && (SWE > 0.0)) {
          double melt_cap = StateVarMax(c, iSnLiq);
          double fill_rate = Fs * dmin(dmax(melt_cap - c.SV(iSnLiq), 0.0) / tstep, rainthru);
          c.R(qSnowLiq) = fill_rate;
          rainthru -= fill_rate;
        }
        if ((c.iSV(RVN_SV_SNOW_DEFICIT) >= 0) && (SWE > 0.0)) c.R(qSnowDef) = c.G(RVN_G_SNOW_SWI) * snowthru;
      }
    } else c.R(qNewSnow) = snowthru;
    c.R(qPond) = rainthru;
  }
  (void)pr;
}

// CmvSnowBalance SNOBAL_HMETS  src/SnowBalance.cpp:472-522
template <class C, class PR> RVN_DI void rates_SNOBAL_HMETS(C &c, const PR &pr) {
  const double tstep = c.tstep;
  double SWE = c.SV(pr.from(0)), SL = c.SV(pr.from(3)), cum_melt = c.SV(pr.from(4));
  const double Kf = c.LU(RVN_LU_REFREEZE_FACTOR), Tbf = c.LU(RVN_LU_DD_REFREEZE_TEMP), exp_fe = c.LU(RVN_LU_REFREEZE_EXP);
  const double fcmin = c.G(RVN_G_SNOW_SWI_MIN), fcmax = c.G(RVN_G_SNOW_SWI_MAX), Ccum = c.G(RVN_G_SWI_REDUCT_COEFF);
  const double potmelt = dmax(c.F[RVN_F_POTENTIAL_MELT], 0.0);
  const double Tdiurnal = 0.5 * (c.F[RVN_F_TEMP_DAILY_AVE] + c.F[RVN_F_TEMP_DAILY_MIN]);
  SL = dmax(SL, 0.0);
  double pot_freeze = Kf * rvn_pow(dmax(Tbf - Tdiurnal, 0.0), exp_fe);
  double freeze = dmin(pot_freeze * tstep, SL);
  SL -= freeze; SWE += freeze;
  double melt = dmin(potmelt * tstep, SWE);
  SWE -= melt; cum_melt += melt;
  if (SWE <= 0.0) cum_melt = 0.0;
  double SWI = dmax(fcmin, fcmax * (1 - Ccum * cum_melt));
  double deficit = dmax(SWI * SWE - SL, 0.0);
  double to_liq = dmin(melt, deficit);
  melt -= to_liq; SL += to_liq;
  double ripe = dmax(SL - SWI * SWE, 0.0);
  SL -= ripe;
  c.R(0) = to_liq / tstep; c.R(1) = melt / tstep; c.R(2) = ripe; c.R(3) = freeze / tstep; c.R(4) = (melt + to_liq) / tstep;
  if (SWE <= 0.0) c.R(4) = -c.SV(pr.from(4)) / tstep;
}

// CmvSnowBalance SNOBAL_SIMPLE_MELT  src/SnowBalance.cpp:393-396; ApplyConstraints :1311-1317 (no FIRN)
template <class C, class PR> RVN_DI void rates_SNOBAL_SIMPLE_MELT(C &c, const PR &pr) {
  c.R(0) = dmax(c.F[RVN_F_POTENTIAL_MELT], 0.0);
  if (c.R(0) < 0.0) c.R(0) = 0.0;
  c.R(0) = dmin(c.R(0), dmax(c.SV(pr.from(0)) / c.tstep, 0.0));
}
// CmvSnowBalance SNOBAL_CEMA_NEIGE  src/SnowBalance.cpp:398-413 (snow temperature is the lagged, stored value)
template <class C, class PR> RVN_DI void rates_SNOBAL_CEMA_NEIGE(C &c, const PR &pr) {
  const double avg_annual_snow = c.G(RVN_G_AVG_ANNUAL_SNOW);
  const double snotemp = SnowTemperature(c);
  double pot_melt = 0.0;
  if (snotemp == D_FREEZING_TEMP) pot_melt = dmax(c.F[RVN_F_POTENTIAL_MELT], 0.0);
  const double SWE = c.SV(pr.from(0));
  const double snow_cov = dmin(SWE / avg_annual_snow, 1.0);
  c.R(0) = (0.9 * snow_cov + 0.1) * dmin(SWE / c.tstep, pot_melt);
  c.R(1) = (snow_cov - c.SV(pr.from(1))) / c.tstep;
}
// CmvSnowRefreeze FREEZE_DEGREE_DAY  src/SnowMeltRefreeze.cpp:324-341; ApplyConstraints :354-365
template <class C, class PR> RVN_DI void rates_SNOWREFREEZE_DD(C &c, const PR &pr) {
  const double Kf = c.LU(RVN_LU_REFREEZE_FACTOR), Ta = c.F[RVN_F_TEMP_DAILY_AVE];
  c.R(0) = dmax(Kf * (D_FREEZING_TEMP - Ta), 0.0);
  if (c.SV(pr.from(0)) <= 0) c.R(0) = 0.0;
  c.R(0) = dmin(c.R(0), c.SV(pr.from(0)) / c.tstep);
  if (c.R(0) < 0) c.R(0) = 0.0;
}
// CmvSnowTempEvolve SNOTEMP_NEWTONS  src/SnowTempEvolve.cpp:42-77
template <class C, class PR> RVN_DI void rates_SNOTEMP_NEWTONS(C &c, const PR &pr) {
  const double Tsnow = c.SV(pr.from(0)), Tair = c.F[RVN_F_TEMP_AVE];
  c.R(0) = c.G(RVN_G_AIRSNOW_COEFF) * (Tair - Tsnow);
  c.R(0) = dmin(c.R(0), (D_FREEZING_TEMP - Tsnow) / c.tstep);
}
// CmvCanopyEvap CANEVP_ALL  src/VegetationMovers.cpp:110-160; ApplyConstraints :176-190
template <class C, class PR> RVN_DI void rates_CANEVP_ALL(C &c, const PR &pr) {
  if ((c.type != RVN_HRU_STANDARD) && (c.type != RVN_HRU_WETLAND)) return;
  const double Fc = c.LU(RVN_LU_FOREST_COVERAGE);
  if (Fc != 0) {
    c.R(0) = c.SV(pr.from(0)) / c.tstep;
    c.R(1) = c.R(0);
  }
  const double oldRates = c.R(0);
  c.R(0) = dmax(c.R(0), 0.0);
  c.R(0) = dmin(c.R(0), c.SV(pr.from(0)) / c.tstep);
  c.R(1) -= (oldRates - c.R(0));
}
// CmvCanopyEvap CANEVP_MAXIMUM / CANEVP_RUTTER and CmvCanopySublimation SUBLIM_MAXIMUM: evaporation at (a share of) the PET left
// by earlier consumers  src/VegetationMovers.cpp:110-190, 291-346.  MODE 0 = MAXIMUM, 1 = RUTTER (Ft = 0: no TRUNK variable), 2 = snow MAXIMUM
template <int MODE, class C, class PR> RVN_DI void rates_CANEVP_PET(C &c, const PR &pr) {
  if ((c.type != RVN_HRU_STANDARD) && (c.type != RVN_HRU_WETLAND)) return;
  const double Fc = c.LU(RVN_LU_FOREST_COVERAGE);
  if (Fc != 0) {
    double PET = dmax(c.F[RVN_F_PET], 0.0);
    if (!c.opt().suppress_competitive_ET) { PET -= (c.SV(c.iSV(RVN_SV_AET)) / c.tstep); PET = dmax(PET, 0.0); }
    if (MODE == 1) {
      const double cap = c.canopy_cap;
      const double stor = dmin(dmax(c.SV(pr.from(0)), 0.0), cap * Fc);
      c.R(0) = (1.0 - 0.0) * Fc * PET * (stor / (cap * Fc));
    } else c.R(0) = Fc * PET;
    c.R(1) = c.R(0);
  }
  const double oldRates = c.R(0);
  if (MODE != 2) c.R(0) = dmax(c.R(0), 0.0);
  c.R(0) = dmin(c.R(0), c.SV(pr.from(0)) / c.tstep);
  c.R(1) -= (oldRates - c.R(0));
}
// CmvAbstraction ABST_PERCENTAGE / ABST_FILL  src/Abstraction.cpp:230-256; ApplyConstraints :514-532
template <class C, class PR> RVN_DI void rates_ABSTRACTION(C &c, const PR &pr) {
  const double ponded = c.SV(pr.from(0)), depression = c.SV(pr.to(0));
  if (c.type != RVN_HRU_STANDARD) return;
  if (pr.ia(0) == RVN_ABST_PERCENTAGE) c.R(0) = c.LU(RVN_LU_ABST_PERCENT) * ponded / c.tstep;
  else if (pr.ia(0) == RVN_ABST_SCS) {   // src/Abstraction.cpp:257-285: initial abstraction of the curve-number method (condition-3 polynomial as written there)
    int condition = 2;
    const double TR = c.precip_5day / 25.4;
    double CN = c.LU(RVN_LU_SCS_CN);
    if ((c.month > 4) && (c.month < 9)) { if (TR < 1.4) condition = 1; else if (TR > 2.1) condition = 3; }
    else { if (TR < 0.5) condition = 1; else if (TR > 1.1) condition = 3; }
    if (condition == 1) CN = 5E-05 * rvn_pow(CN, 3) + 0.0008 * rvn_pow(CN, 2) + 0.4431 * CN;
    else if (condition == 3) CN = 0.0003 * rvn_pow(CN, 3) - 0.0185 * rvn_pow(CN, 2) + 2.1586 * CN;
    const double S = 25.4 * (1000 / CN - 10);
    const double Ia = (c.LU(RVN_LU_SCS_IA_FRACTION)) * S;
    c.R(0) = dmax(Ia, ponded) / c.tstep;
  } else if (pr.ia(0) == RVN_ABST_PDMROF) {   // :287-309 (Mekonnen et al. 2014): Pareto-distributed depression capacity, the rest runs off
    const double dep_max = c.LU(RVN_LU_DEP_MAX), b = c.LU(RVN_LU_PDMROF_B);
    const double c_max = (b + 1) * dep_max;
    const double c_star = c_max * (1.0 - rvn_pow(1.0 - (depression / dep_max), 1.0 / (b + 1.0)));
    double abstracted = dep_max * (rvn_pow(1.0 - (c_star / c_max), b + 1.0) - rvn_pow(1.0 - dmin(c_star + ponded, c_max) / c_max, b + 1.0));
    abstracted = dmin(abstracted, ponded);
    c.R(0) = (abstracted) / c.tstep;
    c.R(1) = (ponded - abstracted) / c.tstep;
  } else {
    const double dep_space = dmax(c.LU(RVN_LU_DEP_MAX) - depression, 0.0);
    c.R(0) = dmin(dep_space, dmax(ponded, 0.0)) / c.tstep;
  }
  const double pond = dmax(c.SV(pr.from(0)), 0.0);
  c.R(0) = dmin(c.R(0), pond / c.tstep);
  const double stor = dmax(c.SV(pr.to(0)), 0.0), max_stor = StateVarMax(c, pr.to(0));
  const double deficit = dmax(max_stor - stor, 0.0);
  c.R(0) = dmin(c.R(0), deficit / c.tstep);
}
// CmvCanopySublimation SUBLIM_ALL (:CanopySnowEvap CANEVP_ALL)  src/VegetationMovers.cpp:283-346
template <class C, class PR> RVN_DI void rates_CANSNOWEVP_ALL(C &c, const PR &pr) {
  if ((c.type != RVN_HRU_STANDARD) && (c.type != RVN_HRU_WETLAND)) return;
  const double Fc = c.LU(RVN_LU_FOREST_COVERAGE);
  if (Fc != 0) {
    c.R(0) = c.SV(pr.from(0)) / c.tstep;
    c.R(1) = c.R(0);
  }
  const double oldRates = c.R(0);
  c.R(0) = dmin(c.R(0), c.SV(pr.from(0)) / c.tstep);
  c.R(1) -= (oldRates - c.R(0));
}
// CmvSublimation  src/Sublimation.cpp:98-243 (SublimationRate, GetRatesOfChange, ApplyConstraints).  Forcings and the snow
// temperature / depth / SWE of the depth-change term are the HRU's lagged values, as in the reference.
RVN_DI double SatVap(const double T) { return (T >= 0) ? 0.61078 * rvn_exp(17.26939 * T / (T + 237.3)) : 0.61078 * rvn_exp(21.87456 * T / (T + 265.5)); }
template <class C, class PR> RVN_DI void rates_SUBLIMATION(C &c, const PR &pr) {
  const int type = pr.ia(0);
  const double Ta = c.F[RVN_F_TEMP_AVE], rel_hum = c.rel_hum, wind_vel = c.wind_vel, Tsnow = SnowTemperature(c);
  double sub = 0.0;
  if (type == RVN_SUBLIM_SVERDRUP) {
    const double roughness = c.G(RVN_G_SNOW_ROUGHNESS), air_pres = c.air_pres;
    const double air_density = (air_pres - 0.378 * SatVap(Ta)) / (287.042 * (Ta + D_ZERO_CELSIUS)) * 1000;   // GetAirDensity, src/CommonFunctions.cpp:1039-1043
    const double es = SatVap(Tsnow), ea = SatVap(Ta) * rel_hum;
    const double C_E = 0.42 * 0.42 / (rvn_log(8.0 / roughness)) / (rvn_log(8.0 / roughness));
    sub = air_density * C_E * wind_vel * 0.623 * (es - ea) / air_pres * D_SEC_PER_DAY / D_DENSITY_WATER * D_MM_PER_METER;
  } else if (type == RVN_SUBLIM_KUZMIN) {
    const double ea = SatVap(c.F[RVN_F_TEMP_DAILY_AVE]) * rel_hum * 10, es = SatVap(Tsnow) * 1.0 * 10;
    sub = ((0.18 + 0.098 * wind_vel) * (es - ea));
  } else if (type == RVN_SUBLIM_KUCHMENT_GELFAN) {
    const double ea = SatVap(Ta) * rel_hum * 10, es = SatVap(Tsnow) * 1.0 * 10;
    const double Q_E = 32.82 * (0.18 + 0.098 * wind_vel) * (es - ea) * 0.0864;
    sub = Q_E / 2.845 / D_DENSITY_WATER * (D_MM_PER_METER);
  } else if (type == RVN_SUBLIM_BULK_AERO) {
    const double air_pres = c.air_pres;
    const double air_dens = (air_pres - 0.378 * SatVap(Ta)) / (287.042 * (Ta + D_ZERO_CELSIUS)) * 1000;
    const double ea = SatVap(Ta) * rel_hum, es = SatVap(Tsnow) * 1.0;
    const double z_ref = 2.0, z0 = 0.00005, z0e = 0.00005;
    const double C_E = (0.42 * 0.42) / (rvn_log(z_ref / z0) * rvn_log(z_ref / z0e));
    const double Q_E = air_dens * 2.845 * C_E * wind_vel * ((D_AIR_H20_MW_RAT * (es - ea)) / air_pres) * D_SEC_PER_DAY;
    sub = Q_E / (2.845 * D_DENSITY_WATER) * D_MM_PER_METER;
  } else if (type == RVN_SUBLIM_CENTRAL_SIERRA) {
    const double sat_vap = SatVap(Tsnow) * 10, vap_pres = SatVap(c.F[RVN_F_TEMP_DAILY_AVE]) * rel_hum * 10;
    const double wind_v = wind_vel * 2.237;
    sub = ((0.0063 * 1.0 * wind_v * (sat_vap - vap_pres)) * 25.4);   // the height factor rvn_pow(h*h, (-1/6)) is rvn_pow(.., 0) in the reference (integer division)
  }
  c.R(0) = sub;
  if (pr.nconn() == 2) {   // SNOW_DEPTH simulated: lagged depth and SWE
    if (c.old_snow > 0) c.R(1) = c.R(0) * c.old_snow_depth / c.old_snow;
  }
  if (c.SV(pr.from(0)) <= 0) c.R(0) = 0.0;
  c.R(0) = dmin(c.R(0), c.SV(pr.from(0)) / c.tstep);   // threshMin(.., .., 0.0)
}
// CmvOWEvaporation OPEN_WATER_EVAP  src/OpenWaterEvap.cpp:113-160; ApplyConstraints :170-186
template <bool RIPARIAN = false, class C, class PR> RVN_DI void rates_OPEN_WATER_EVAP(C &c, const PR &pr) {   // RIPARIAN: OPEN_WATER_RIPARIAN (:137-141), evaporation from the stream fraction of the HRU
  if (c.type == RVN_HRU_MASKED_GLACIER) return;
  if (c.linked) return;   // reservoir-linked HRUs handle ET via the reservoir mass balance (src/OpenWaterEvap.cpp:124)
  double OWPET = c.F[RVN_F_OW_PET];
  if (!c.opt().suppress_competitive_ET) { OWPET -= (c.SV(c.iSV(RVN_SV_AET)) / c.tstep); OWPET = dmax(OWPET, 0.0); }
  if (RIPARIAN) c.R(0) = c.LU(RVN_LU_OW_PET_CORR) * c.LU(RVN_LU_STREAM_FRACTION) * OWPET;
  else c.R(0) = c.LU(RVN_LU_OW_PET_CORR) * OWPET;
  if (c.R(0) < 0) c.R(0) = 0.0;
  if (pr.from(0) != c.iSV(RVN_SV_SURFACE_WATER)) {
    c.R(0) = dmin(c.R(0), c.SV(pr.from(0)) / c.tstep);
    if (c.SV(pr.from(0)) <= 0) c.R(0) = 0.0;
  }
  c.R(pr.nconn() - 1) = c.R(0);
}
// CmvLakeEvaporation LAKE_EVAP_BASIC  src/OpenWaterEvap.cpp:269-290; ApplyConstraints :300-314.  ia(0) = 1 when the
// source is the :LakeStorage variable (then the process applies to every HRU type, not only lakes)
template <class C, class PR> RVN_DI void rates_LAKE_EVAP_BASIC(C &c, const PR &pr) {
  if (((c.type == RVN_HRU_LAKE) || (pr.ia(0) == 1)) && !c.linked) {   // reservoir-linked HRUs: ET via the reservoir mass balance (:279)
    c.R(0) = c.LU(RVN_LU_LAKE_PET_CORR) * c.F[RVN_F_OW_PET];
    c.R(pr.nconn() - 1) = c.R(0);
  }
  if (c.SV(pr.from(0)) <= 0) c.R(0) = 0.0;
  if (c.R(0) < 0) c.R(0) = 0.0;
  c.R(0) = dmin(c.R(0), c.SV(pr.from(0)) / c.tstep);
  c.R(pr.nconn() - 1) = c.R(0);
}
// CmvGlacierMelt GMELT_HBV  src/GlacierProcesses.cpp:117-160; ApplyConstraints :176-190 (glacier_model_on = false)
template <class C, class PR> RVN_DI void rates_GLACIERMELT_HBV(C &c, const PR &pr) {
  if (c.type != RVN_HRU_GLACIER) return;
  if (!(c.SV(c.iSV(RVN_SV_SNOW)) > D_REAL_SMALL)) c.R(0) = c.LU(RVN_LU_HBV_MELT_GLACIER_CORR) * dmax(c.F[RVN_F_POTENTIAL_MELT], 0.0);
  if (c.R(0) < 0.0) c.R(0) = 0.0;
  (void)pr;
}
// CmvGlacierMelt GMELT_SIMPLE_MELT / GMELT_UBC  src/GlacierProcesses.cpp:126-129,139-166; ApplyConstraints :170-186 (glacier_model_on off)
template <class C, class PR> RVN_DI void rates_GLACIERMELT_SIMPLE(C &c, const PR &pr) {
  if (c.type != RVN_HRU_GLACIER) return;
  c.R(0) = c.F[RVN_F_POTENTIAL_MELT];
  if (c.R(0) < 0.0) c.R(0) = 0.0;
  (void)pr;
}
template <class C, class PR> RVN_DI void rates_GLACIERMELT_UBC(C &c, const PR &pr) {
  if (c.type != RVN_HRU_GLACIER) return;
  double CC = c.SV(pr.from(1));
  const double snow_coverage = SnowCover(c);
  double pot_melt = c.F[RVN_F_POTENTIAL_MELT];
  if (pot_melt > 0.0) pot_melt *= (1.0 - snow_coverage);   // no glacier melt under snow
  else pot_melt = 0.0;
  if (CC > pot_melt * c.tstep) { CC -= pot_melt * c.tstep; pot_melt = 0; }
  else { pot_melt -= CC / c.tstep; CC = 0.0; }
  const double P0CTKG = 0.0;
  CC *= (1 - (1 / (1 + P0CTKG)));
  c.R(0) = pot_melt;
  c.R(1) = (CC - c.SV(pr.from(1))) / c.tstep;
  if (c.R(0) < 0.0) c.R(0) = 0.0;
}
// CmvGlacierRelease GRELEASE_LINEAR / GRELEASE_LINEAR_ANALYTIC  src/GlacierProcesses.cpp:317-327; ApplyConstraints :340-350
template <class C, class PR> RVN_DI void rates_GLACIERRELEASE_LINEAR(C &c, const PR &pr) {
  if (c.type != RVN_HRU_GLACIER) return;
  const double glacier_stor = c.SV(c.iSV(RVN_SV_GLACIER)), K = c.LU(RVN_LU_GLAC_STORAGE_COEFF);
  if (pr.ia(0) == 0) c.R(0) = K * glacier_stor;
  else c.R(0) = glacier_stor * (1 - rvn_exp(-K * c.tstep)) / c.tstep;
  if (c.R(0) < 0.0) c.R(0) = 0.0;
}
// CmvGlacierRelease GRELEASE_HBV_EC  src/GlacierProcesses.cpp:292-325; ApplyConstraints :340-350
template <class C, class PR> RVN_DI void rates_GLACIERRELEASE_HBVEC(C &c, const PR &pr) {
  if (c.type != RVN_HRU_GLACIER) return;
  const double glacier_stor = c.SV(c.iSV(RVN_SV_GLACIER)), snow = c.SV(c.iSV(RVN_SV_SNOW)), liq_snow = c.SV(c.iSV(RVN_SV_SNOW_LIQ));
  const double Kmin = c.LU(RVN_LU_HBV_GLACIER_KMIN), Kmax = c.LU(RVN_LU_GLAC_STORAGE_COEFF), Ag = c.LU(RVN_LU_HBV_GLACIER_AG);
  const double K = Kmin + (Kmax - Kmin) * rvn_exp(-Ag * (snow + liq_snow));
  c.R(0) = K * glacier_stor;
  if (c.R(0) < 0.0) c.R(0) = 0.0;
  (void)pr;
}

// LambertN (src/CommonFunctions.cpp:1572-1579), N = 3: the recursion of the reference unrolled (sigma is the same at every depth)
RVN_DI double lambert_n3(const double x) {
  const double lx = rvn_log(-x);
  const double sigma = rvn_pow(-2.0 - 2.0 * lx, 0.5);
  double tmp = -1.0 - 0.5 * sigma * sigma - sigma / (1.0 + sigma / 6.3);
  tmp = (1 + lx - rvn_log(-tmp)) * tmp / (1 + tmp);            // LambertN(x, 2)
  return (1 + lx - rvn_log(-tmp)) * tmp / (1 + tmp);           // LambertN(x, 3)
}
// CmvInfiltration::GreenAmptCumInf  src/GreenAmpt.cpp:26-52
RVN_DI double green_ampt_cum_inf(const double t, const double alpha, const double Ks, const double w) {
  if (Ks == 0) return 0.0;
  if (w == 0) return 0.0;
  if (alpha == 0) return 0.0;
  const double tp = alpha * Ks / w / (w - Ks);
  if ((t < tp) || (w < Ks)) return w * t;
  const double Fp = w * tp;
  const double tstar = Ks / alpha * (t - tp);
  const double x = -(1.0 + Fp / alpha) * rvn_exp(-(1.0 + tstar + Fp / alpha));
  return alpha * (-1.0 - lambert_n3(x));
}
// CmvInfiltration INF_HMETS / INF_HBV / INF_GR4J  src/Infiltration.cpp:304-330, 384-398, 477-489, 491-514; ApplyConstraints :605-626
template <class C, class PR> RVN_DI void rates_INFILTRATION(C &c, const PR &pr, const int op) {
  const double tstep = c.tstep;
  const int type = c.type;
  if (type != RVN_HRU_STANDARD && type != RVN_HRU_ROCK && type != RVN_HRU_MASKED_GLACIER) return;
  const double Fimp = c.LU(RVN_LU_IMPERMEABLE_FRAC);
  const double ponded = dmax(c.SV(c.iSV(RVN_SV_PONDED_WATER)), 0.0);
  const double rainthru = (ponded / tstep);
  if (type == RVN_HRU_ROCK) { c.R(0) = 0.0; c.R(1) = rainthru; }
  else if (op == RVN_OP_INF_HMETS) {
    double stor = c.SV(pr.to(0)), max_stor = SoilCapacity(c, 0), coef = c.LU(RVN_LU_HMETS_RUNOFF_COEFF);
    double sat = dmin(dmax(stor / max_stor, 0.0), 1.0);
    double direct = Fimp * rainthru;
    double runoff = coef * (sat) * (1.0 - Fimp) * rainthru;
    double infil = (1.0 - Fimp) * rainthru - runoff;
    double delayed = coef * rvn_sqr(sat) * infil;
    infil -= delayed;
    c.R(0) = infil; c.R(1) = direct; c.R(2) = runoff; c.R(3) = delayed;
  } else if (op == RVN_OP_INF_HBV) {
    double stor = c.SV(pr.to(0)), max_stor = SoilCapacity(c, 0), beta = c.SOIL(0, RVN_S_HBV_BETA);
    double sat = dmax(dmin(stor / max_stor, 1.0), 0.0);
    double runoff = rvn_pow(sat, beta) * rainthru;
    runoff = (Fimp)*rainthru + (1 - Fimp) * runoff;
    c.R(0) = rainthru - runoff; c.R(1) = runoff;
  } else if (op == RVN_OP_INF_VIC_ARNO) {   // src/Infiltration.cpp:400-416
    double stor = c.SV(pr.to(0)), max_stor = SoilCapacity(c, 0), b = c.SOIL(0, RVN_S_VIC_B_EXP);
    double sat = dmin(stor / max_stor, 1.0);
    double sat_area = 1.0 - rvn_pow(1.0 - sat, b);
    double runoff = sat_area * rainthru;
    runoff = (Fimp)*rainthru + (1 - Fimp) * runoff;
    c.R(0) = rainthru - runoff; c.R(1) = runoff;
  } else if (op == RVN_OP_INF_RATIONAL) {   // src/Infiltration.cpp:333-338 (no impermeable-fraction correction in the reference)
    double runoff = c.LU(RVN_LU_PARTITION_COEFF) * rainthru;
    c.R(0) = rainthru - runoff; c.R(1) = runoff;
  } else if (op == RVN_OP_INF_SCS || op == RVN_OP_INF_SCS_NOABSTRACTION) {   // :338-343 + GetSCSRunoff :638-680
    int condition = 2;   // antecedent moisture condition
    const double TR = c.precip_5day / 25.4;   // [in] (MM_PER_INCH)
    double CN = c.LU(RVN_LU_SCS_CN);
    if ((c.month > 4) && (c.month < 9)) { if (TR < 1.4) condition = 1; else if (TR > 2.1) condition = 3; }
    else { if (TR < 0.5) condition = 1; else if (TR > 1.1) condition = 3; }
    if (condition == 1) CN = 5E-05 * rvn_pow(CN, 3) + 0.0008 * rvn_pow(CN, 2) + 0.4431 * CN;
    else if (condition == 3) CN = 7E-05 * rvn_pow(CN, 3) - 0.0185 * rvn_pow(CN, 2) + 2.1586 * CN;
    CN = dmin(CN, 100.0);
    const double S = 25.4 * (1000.0 / CN - 10.0);
    const double W = rainthru * tstep;
    const double Ia = (op == RVN_OP_INF_SCS) ? (c.LU(RVN_LU_SCS_IA_FRACTION)) * S : 0.0;
    double Weff = rvn_pow(dmax(W - Ia, 0.0), 2) / (W + (S - Ia));
    if (W + (S - Ia) == 0.0) Weff = 0.0;
    const double runoff = Weff / tstep;
    c.R(0) = rainthru - runoff; c.R(1) = runoff;
  } else if (op == RVN_OP_INF_ALL_INFILTRATES) {   // :347-353
    double runoff = 0.0;
    runoff = (Fimp)*rainthru + (1 - Fimp) * runoff;
    c.R(0) = rainthru - runoff; c.R(1) = runoff;
  } else if (op == RVN_OP_INF_VIC) {   // :361-383
    double stor = c.SV(c.soil_slot(0)), max_stor = SoilCapacity(c, 0);
    double zmax = c.SOIL(0, RVN_S_VIC_ZMAX), zmin = c.SOIL(0, RVN_S_VIC_ZMIN), alpha = c.SOIL(0, RVN_S_VIC_ALPHA);
    double sat = dmin(stor / max_stor, 1.0);
    double gamma = 1.0 / (alpha + 1.0);
    double K1 = rvn_pow((zmax - zmin) * alpha * gamma, -gamma);
    double Smax = gamma * (alpha * zmax + zmin);
    double runoff = rainthru * dmax(1.0 - K1 * rvn_pow(Smax - sat, gamma), 1.0);   // threshMax(.,1.0,0.0) as written in the reference
    runoff = (Fimp)*rainthru + (1 - Fimp) * runoff;
    c.R(0) = rainthru - runoff; c.R(1) = runoff;
  } else if (op == RVN_OP_INF_PRMS) {   // :418-431
    double sat_frac = c.LU(RVN_LU_MAX_SAT_AREA_FRAC), tens_stor = SoilTensionCapacity(c, 0), stor = c.SV(c.soil_slot(0));
    double runoff = sat_frac * dmin(stor / tens_stor, 1.0) * rainthru;
    runoff = (Fimp)*rainthru + (1 - Fimp) * runoff;
    c.R(0) = rainthru - runoff; c.R(1) = runoff;
  } else if (op == RVN_OP_INF_PDM) {   // :540-563 (Moore 1985)
    double b = c.LU(RVN_LU_PDM_B), stor = c.SV(c.soil_slot(0)), max_stor = SoilCapacity(c, 0);
    double c_max = (b + 1) * max_stor;
    double c_star = c_max * (1.0 - rvn_pow(1.0 - (stor / max_stor), 1.0 / (b + 1.0)));
    double infil = max_stor * (rvn_pow(1.0 - (c_star / c_max), b + 1.0) - rvn_pow(1.0 - dmin(c_star + ponded, c_max) / c_max, b + 1.0));
    infil = dmin(infil, ponded);
    infil = (1.0 - Fimp) * infil / tstep;
    c.R(0) = infil; c.R(1) = rainthru - infil;
  } else if (op == RVN_OP_INF_GREEN_AMPT || op == RVN_OP_INF_GA_SIMPLE) {   // src/GreenAmpt.cpp:65-175
    const double Ksat = c.SOIL(0, RVN_S_HYDRAUL_COND), psi_wf = c.SOIL(0, RVN_S_WETTING_FRONT_PSI), poro = c.SOIL(0, RVN_S_POROSITY);
    const double stor = c.SV(pr.to(0)), max_stor = SoilCapacity(c, 0);
    const double Keff = Ksat;
    double cumInf = 0, initStor = 0;
    if (op == RVN_OP_INF_GA_SIMPLE) {   // rainfall-event memory: cumulative infiltration and the storage at the start of the event
      const double cum0 = c.SV(pr.from(2)), init0 = c.SV(pr.from(3));
      cumInf = cum0; initStor = init0;
      if (rainthru <= Keff) {   // redistribution during periods of low rainfall
        if (cumInf <= 0) { cumInf = 0; initStor = 0; }
        else {
          const double distribution_rate = dmax(cumInf * 2 / 3, Keff);
          cumInf -= distribution_rate * tstep;
          initStor += distribution_rate * tstep;
          initStor = dmin(initStor, stor);
          if (cumInf <= 0) { cumInf = 0; initStor = 0; }
        }
        c.R(2) = (cumInf - cum0) / tstep;
        c.R(3) = (initStor - init0) / tstep;
      }
    }
    if (rainthru <= D_REAL_SMALL) { c.R(0) = 0; c.R(1) = 0; }
    else {
      double deficit = (1.0 - stor / max_stor) * poro;
      double alpha = psi_wf * deficit;
      if (op == RVN_OP_INF_GREEN_AMPT) cumInf = green_ampt_cum_inf(tstep, alpha, Keff, rainthru);
      else if (cumInf <= 0) { cumInf = green_ampt_cum_inf(tstep, alpha, Keff, rainthru); initStor = stor; }   // first step of a rainfall event
      else { deficit = (1.0 - initStor / max_stor) * poro; alpha = psi_wf * deficit; }
      double finf = 0.0;
      if (cumInf > 0) finf = Keff * (1.0 + alpha / cumInf);
      double runoff = dmax(rainthru - finf, 0.0);
      runoff = (Fimp)*rainthru + (1 - Fimp) * runoff;
      if (op == RVN_OP_INF_GA_SIMPLE) {
        cumInf += (rainthru - runoff) * tstep;
        c.R(2) = (cumInf - c.SV(pr.from(2))) / tstep;
        c.R(3) = (initStor - c.SV(pr.from(3))) / tstep;
      }
      c.R(0) = rainthru - runoff; c.R(1) = runoff;
    }
  } else if (op == RVN_OP_INF_UBC) {   // GetUBCWMRunoff, src/Infiltration.cpp:685-759 (Quick 1995): topsoil deficit, groundwater and interflow shares, flash-flood factor
    const double V0FLAX = 1800.0;
    const double soil_deficit = dmax(0.0, SoilCapacity(c, 0) - c.SV(pr.to(0)));
    const double max_perc_rate = c.SOIL(1, RVN_S_MAX_PERC_RATE), P0AGEN = c.SOIL(0, RVN_S_UBC_INFIL_SOIL_DEF);
    const double P0DSH = c.G(RVN_G_UBC_GW_SPLIT), V0FLAS = c.G(RVN_G_UBC_FLASH_PONDING);
    double b1 = 1.0;
    if (Fimp < 1.0) b1 = Fimp * rvn_pow(10.0, -soil_deficit / P0AGEN);
    double flash_factor = 0.0;
    if (rainthru > V0FLAS) flash_factor = (1.0 + rvn_log(rainthru / V0FLAX) / rvn_log(V0FLAX / V0FLAS));
    if (1.0 < flash_factor) flash_factor = 1.0;
    if (0.0 > flash_factor) flash_factor = 0.0;
    const double b2 = (b1) * (1.0) + (1.0 - b1) * flash_factor;
    double infil = dmin(soil_deficit / tstep, rainthru);
    double to_GW = dmin(max_perc_rate, rainthru - infil);
    double to_interflow = (rainthru - infil - to_GW);
    infil *= 1 - b2; to_GW *= 1 - b2; to_interflow *= 1 - b2;
    const double runoff = rainthru - infil - to_GW - to_interflow;
    c.R(0) = infil; c.R(1) = runoff; c.R(2) = to_interflow; c.R(3) = (1.0 - P0DSH) * to_GW; c.R(4) = (P0DSH)*to_GW;
  } else if (op == RVN_OP_INF_GR4J) {
    double x1 = SoilCapacity(c, 0), stor = c.SV(pr.to(0));
    double sat = stor / x1;
    double tmp = rvn_tanh(ponded / x1);
    double infil = x1 * (1.0 - (sat * sat)) * tmp / (1.0 + sat * tmp);
    infil = (1.0 - Fimp) * infil;
    c.R(0) = infil; c.R(1) = rainthru - infil;
  }
  if (type != RVN_HRU_STANDARD && type != RVN_HRU_MASKED_GLACIER) return;
  c.R(0) = dmin(c.R(0), c.SV(pr.from(0)) / tstep);
  double inf = c.R(0);
  if (!c.opt().allow_soil_overfill) {
    double max_stor = StateVarMax(c, pr.to(0));
    inf = dmin(c.R(0), dmax(max_stor - c.SV(pr.to(0)), 0.0) / tstep);
  }
  c.R(1) += (c.R(0) - inf);
  c.R(0) = inf;
}

// CmvOverflow  src/Flush.cpp:259-279
template <class C, class PR> RVN_DI void rates_OVERFLOW(C &c, const PR &pr) {
  double max_storage = dmax(StateVarMax(c, pr.from(0)), 0.0);
  c.R(0) = dmax(c.SV(pr.from(0)) - max_storage, 0.0) / c.tstep;
}
// CmvExchangeFlow  src/Flush.cpp:358-368 (no constraints, :380-388)
template <class C, class PR> RVN_DI void rates_EXCHANGE_FLOW(C &c, const PR &pr) {
  const double q = c.SOIL(pr.ia(0), RVN_S_EXCHANGE_FLOW);
  c.R(0) = q; c.R(1) = q;
}
// CmvDepressionOverflow  src/DepressionProcesses.cpp:106-165 (DEPRESSION -> SURFACE_WATER)
template <class C, class PR> RVN_DI void rates_DEPRESSION_OVERFLOW(C &c, const PR &pr) {
  const double stor = c.SV(pr.from(0)), thresh_stor = c.LU(RVN_LU_DEP_THRESHOLD);
  if (pr.ia(0) == 0) {   // DFLOW_THRESHPOW
    const double max_flow = c.LU(RVN_LU_DEP_MAX_FLOW), n = c.LU(RVN_LU_DEP_N), max_stor = c.LU(RVN_LU_DEP_MAX);
    if (max_stor < thresh_stor) c.R(0) = max_flow;
    else c.R(0) = max_flow * rvn_pow(dmax((stor - thresh_stor) / (max_stor - thresh_stor), 0.0), n);
  } else if (pr.ia(0) == 1) {   // DFLOW_LINEAR (analytical)
    c.R(0) = dmax(stor - thresh_stor, 0.0) * (1 - rvn_exp(-c.LU(RVN_LU_DEP_K) * c.tstep)) / c.tstep;
  } else {   // DFLOW_WEIR
    const double cw = c.LU(RVN_LU_DEP_CRESTRATIO) * sqrt(c.area);
    c.R(0) = 0.666 * cw / c.area * sqrt(2 * 9.80616) * rvn_pow(dmax(stor - thresh_stor, 0.0), 1.5);
  }
  c.R(0) = dmin(c.R(0), dmax(c.SV(pr.from(0)) / c.tstep, 0.0));
}
// CmvSeepage SEEP_LINEAR  src/DepressionProcesses.cpp:254-298
template <class C, class PR> RVN_DI void rates_SEEPAGE(C &c, const PR &pr) {
  const double stor = c.SV(pr.from(0));
  c.R(0) = dmax(stor, 0.0) * (1 - rvn_exp(-c.LU(RVN_LU_DEP_SEEP_K) * c.tstep)) / c.tstep;
  c.R(0) = dmin(c.R(0), dmax(c.SV(pr.from(0)) / c.tstep, 0.0));
  const double room = dmax(StateVarMax(c, pr.to(0)) - c.SV(pr.to(0)), 0.0);
  c.R(0) = dmin(c.R(0), room / c.tstep);
}
// CmvFlush  src/Flush.cpp:75-86
template <class C, class PR> RVN_DI void rates_FLUSH(C &c, const PR &pr) { c.R(0) = pr.da(0) * c.SV(pr.from(0)) / c.tstep; }
// CmvSplit  src/Flush.cpp:178-186
template <class C, class PR> RVN_DI void rates_SPLIT(C &c, const PR &pr) {
  c.R(0) = (pr.da(0)) * c.SV(pr.from(0)) / c.tstep;
  c.R(1) = (1.0 - pr.da(0)) * c.SV(pr.from(0)) / c.tstep;
}
// CmvBaseflow BASE_LINEAR / BASE_POWER_LAW / BASE_GR4J  src/Baseflow.cpp:152-186,208-215,238-243; ApplyConstraints :297-310
template <class C, class PR> RVN_DI void rates_BASEFLOW(C &c, const PR &pr, const int op) {
  const int type = c.type;
  if (!(type == RVN_HRU_LAKE || type == RVN_HRU_WATER || type == RVN_HRU_ROCK)) {
    const int m = pr.ia(0);
    double stor = c.SV(pr.from(0)), max_stor = (m >= 0) ? SoilCapacity(c, m) : 0.0;
    stor = dmin(dmax(stor, 0.0), max_stor);
    if (op == RVN_OP_BASE_LINEAR) c.R(0) = c.SOIL(m, RVN_S_BASEFLOW_COEFF) * stor;
    else if (op == RVN_OP_BASE_POWER_LAW) c.R(0) = c.SOIL(m, RVN_S_BASEFLOW_COEFF) * rvn_pow(stor, c.SOIL(m, RVN_S_BASEFLOW_N));
    else if (op == RVN_OP_BASE_LINEAR_ANALYTIC) c.R(0) = stor * (1 - rvn_exp(-c.SOIL(m, RVN_S_BASEFLOW_COEFF) * c.tstep)) / c.tstep;
    else if (op == RVN_OP_BASE_VIC) c.R(0) = c.SOIL(m, RVN_S_MAX_BASEFLOW_RATE) * rvn_pow(stor / max_stor, c.SOIL(m, RVN_S_BASEFLOW_N));
    else if (op == RVN_OP_BASE_TOPMODEL) {
      const double K = c.SOIL(m, RVN_S_MAX_BASEFLOW_RATE), n = c.SOIL(m, RVN_S_BASEFLOW_N), lambda = c.TERR(RVN_T_TOPMODEL_LAMBDA);
      c.R(0) = K * (max_stor / n * rvn_pow(lambda, -n)) * rvn_pow(stor / max_stor, n);
    }
    else if (op == RVN_OP_BASE_LINEAR_CONSTRAIN) {   // src/Baseflow.cpp:188-197
      const double sat = stor / max_stor;
      if (sat > c.SOIL(m, RVN_S_FIELD_CAPACITY)) c.R(0) = c.SOIL(m, RVN_S_BASEFLOW_COEFF) * stor;
    }
    else if (op == RVN_OP_BASE_CONSTANT) c.R(0) = c.SOIL(m, RVN_S_MAX_BASEFLOW_RATE);
    else if (op == RVN_OP_BASE_THRESH_POWER) {   // :250-266
      const double K = c.SOIL(m, RVN_S_MAX_BASEFLOW_RATE), n = c.SOIL(m, RVN_S_BASEFLOW_N), tsat = c.SOIL(m, RVN_S_BASEFLOW_THRESH);
      const double sat = stor / max_stor;
      c.R(0) = 0.0;
      if (sat > tsat) {
        c.R(0) = K * rvn_pow((sat - tsat) / (1.0 - tsat), n);
        c.R(0) = dmin(c.R(0), max_stor * (sat - tsat) / c.tstep);
      }
    }
    else if (op == RVN_OP_BASE_THRESH_STOR) {   // :268-279
      const double K = c.SOIL(m, RVN_S_BASEFLOW_COEFF2), tstor = c.SOIL(m, RVN_S_STORAGE_THRESHOLD);
      c.R(0) = 0.0;
      if (stor > tstor) c.R(0) = K * (stor - tstor);
    }
    else if (op == RVN_OP_BASE_GR4J) {
      const double x3 = c.SOIL(m, RVN_S_GR4J_X3);
      c.R(0) = stor * (1.0 - rvn_pow(1.0 + rvn_pow(dmax(stor / x3, 0.0), 4.0), -0.25)) / c.tstep;
    }
  }
  c.R(0) = dmin(c.R(0), dmax(c.SV(pr.from(0)), 0.0) / c.tstep);
  c.R(0) = dmax(c.R(0), 0.0);
}
// CmvPercolation PERC_LINEAR / PERC_CONSTANT / PERC_GR4J / PERC_GR4JEXCH / PERC_GR4JEXCH2
// src/Percolation.cpp:183-202,237-243,293-316; ApplyConstraints :339-360
template <class C, class PR> RVN_DI void rates_PERCOLATION(C &c, const PR &pr, const int op) {
  const int type = c.type;
  if (type == RVN_HRU_LAKE || type == RVN_HRU_WATER || type == RVN_HRU_ROCK) return;
  const int m = pr.ia(0);
  double stor = c.SV(pr.from(0)), max_stor = SoilCapacity(c, m);
  if (max_stor > 0.0) {
    if (op == RVN_OP_PERC_LINEAR) c.R(0) = c.SOIL(m, RVN_S_PERC_COEFF) * stor;
    else if (op == RVN_OP_PERC_CONSTANT) c.R(0) = c.SOIL(m, RVN_S_MAX_PERC_RATE);
    else if (op == RVN_OP_PERC_GR4J) c.R(0) = stor * (1.0 - rvn_pow(1.0 + rvn_pow(4.0 / 9.0 * dmax(stor / max_stor, 0.0), 4.0), -0.25)) / c.tstep;
    else if (op == RVN_OP_PERC_GAWSER) {   // src/Percolation.cpp:204-213
      const double field_cap = c.SOIL(m, RVN_S_FIELD_CAPACITY) * max_stor, max_rate = c.SOIL(m, RVN_S_MAX_PERC_RATE);
      c.R(0) = max_rate * dmax(stor - field_cap, 0.0) / (max_stor - field_cap);
    } else if (op == RVN_OP_PERC_GAWSER_CONSTRAIN) {   // :215-226
      const double field_cap = c.SOIL(m, RVN_S_FIELD_CAPACITY) * max_stor, max_rate = c.SOIL(m, RVN_S_MAX_PERC_RATE);
      c.R(0) = 0.0;
      if (stor > field_cap) {
        c.R(0) = max_rate * dmax(stor - field_cap, 0.0) / (max_stor - field_cap);
        c.R(0) = dmin(c.R(0), (stor - field_cap) / c.tstep);
      }
    } else if (op == RVN_OP_PERC_POWER_LAW) c.R(0) = c.SOIL(m, RVN_S_MAX_PERC_RATE) * rvn_pow(stor / max_stor, c.SOIL(m, RVN_S_PERC_N));
    else if (op == RVN_OP_PERC_LINEAR_ANALYTIC) c.R(0) = stor * (1 - rvn_exp(-c.SOIL(m, RVN_S_PERC_COEFF) * c.tstep)) / c.tstep;
    else if (op == RVN_OP_PERC_PRMS) {   // :253-267
      const double tens_stor = SoilTensionCapacity(c, m), max_rate = c.SOIL(m, RVN_S_MAX_PERC_RATE), n = c.SOIL(m, RVN_S_PERC_N);
      const double free_stor = dmax(0.0, stor - tens_stor), free_stor_max = (max_stor - tens_stor);
      c.R(0) = max_rate * rvn_pow(free_stor / free_stor_max, n);
    } else if (op == RVN_OP_PERC_SACRAMENTO) {   // :269-294
      const int m2 = pr.ia(1);
      const double stor2 = c.SV(pr.to(0)), tens_stor = SoilTensionCapacity(c, m);
      const double psi = c.SOIL(m2, RVN_S_SAC_PERC_EXPON), alpha = c.SOIL(m2, RVN_S_SAC_PERC_ALPHA);
      const double max_stor2 = (m == 0) ? SoilCapacity(c, m2) : D_ALMOST_INF;
      const double free_stor = dmax(0.0, stor - tens_stor), free_stor_max = (max_stor - tens_stor);
      const double max_baseflow = c.SOIL(m, RVN_S_MAX_BASEFLOW_RATE);
      const double lz_perc = 1.0 + alpha * rvn_pow((1 - (stor2 / max_stor2)), psi);
      c.R(0) = lz_perc * max_baseflow * (free_stor / free_stor_max);
    } else if (op == RVN_OP_PERC_ASPEN) c.R(0) = c.SOIL(m, RVN_S_PERC_ASPEN);
    else if (op == RVN_OP_PERC_GR4JEXCH) {
      const double x2 = c.SOIL(m, RVN_S_GR4J_X2), x3 = c.SOIL(m, RVN_S_GR4J_X3);
      c.R(0) = -x2 * rvn_pow(dmax(dmin(stor / x3, 1.0), 0.0), 3.5);
    } else if (op == RVN_OP_PERC_GR4JEXCH2) {
      const double x2 = c.SOIL(1, RVN_S_GR4J_X2), x3 = c.SOIL(1, RVN_S_GR4J_X3);
      stor = c.SV(c.soil_slot(1));   // SOIL[1]
      c.R(0) = -x2 * rvn_pow(dmax(dmin(stor / x3, 1.0), 0.0), 3.5);
    }
  }
  c.R(0) = dmin(c.R(0), dmax(c.SV(pr.from(0)) - 0.0, 0.0) / c.tstep);
  if (!c.opt().allow_soil_overfill) {
    double room = dmax(StateVarMax(c, pr.to(0)) - c.SV(pr.to(0)), 0.0);
    c.R(0) = dmin(c.R(0), room / c.tstep);
  }
}
// CmvCapillaryRise RISE_HBV  src/CapillaryRise.cpp:103-147; ApplyConstraints :160-179
template <class C, class PR> RVN_DI void rates_CAPRISE_HBV(C &c, const PR &pr) {
  const int type = c.type;
  if (type == RVN_HRU_LAKE || type == RVN_HRU_WATER || type == RVN_HRU_ROCK) return;
  const int m = pr.ia(1);   // soil layer of the receiving store
  double stor = c.SV(pr.to(0)), max_stor = SoilCapacity(c, m);
  stor = dmin(dmax(stor, 0.0), max_stor);
  c.R(0) = c.SOIL(m, RVN_S_MAX_CAP_RISE_RATE) * (1.0 - dmin(stor / max_stor, 1.0));
  c.R(0) = dmin(c.R(0), dmax(c.SV(pr.from(0)), 0.0) / c.tstep);
  if (!c.opt().allow_soil_overfill) c.R(0) = dmin(c.R(0), (StateVarMax(c, pr.to(0)) - c.SV(pr.to(0))) / c.tstep);
}
// CmvSoilEvap SOILEVAP_ALL / SOILEVAP_HBV / SOILEVAP_GR4J  src/SoilEvaporation.cpp:325-343,379-383,385-405,640-650,751; ApplyConstraints :767-783
template <class C, class PR> RVN_DI void rates_SOILEVAP(C &c, const PR &pr, const int op) {
  if (c.type != RVN_HRU_STANDARD) return;
  double PET = c.F[RVN_F_PET];
  if (!c.opt().suppress_competitive_ET) { PET -= (c.SV(c.iSV(RVN_SV_AET)) / c.tstep); PET = dmax(PET, 0.0); }
  double PETused = 0.0; bool multi = false;   // several evaporating stores: PETused is their sum (last connection = AET)
  if (op == RVN_OP_SOILEVAP_ALL) c.R(0) = PET;
  else if (op == RVN_OP_SOILEVAP_HBV || op == RVN_OP_SOILEVAP_TOPMODEL) {
    const double stor = dmax(c.SV(pr.from(0)), 0.0), tens_stor = SoilTensionCapacity(c, 0);
    c.R(0) = PET * dmin(stor / tens_stor, 1.0);
    const int iSnow = c.iSV(RVN_SV_SNOW);
    if (op == RVN_OP_SOILEVAP_HBV && iSnow >= 0) { if (c.SV(iSnow) > D_REAL_SMALL) c.R(0) = (c.LU(RVN_LU_FOREST_COVERAGE)) * c.R(0); }
  } else if (op == RVN_OP_SOILEVAP_GR4J) {
    const double max_stor = SoilCapacity(c, 0), stor = c.SV(pr.from(0));
    const double sat = stor / max_stor;
    const double tmp = rvn_tanh(dmax(PET, 0.0) / max_stor);
    c.R(0) = stor * (2.0 - sat) * tmp / (1.0 + (1.0 - sat) * tmp);
  } else if (op == RVN_OP_SOILEVAP_HYPR) {   // src/SoilEvaporation.cpp:385-420 (Ahmed et al. 2020)
    const double stor = dmax(c.SV(pr.from(0)), 0.0), tens_stor = SoilTensionCapacity(c, 0);
    c.R(0) = PET * dmin(stor / tens_stor, 1.0);
    const double maxPondedAreaFrac = c.LU(RVN_LU_MAX_DEP_AREA_FRAC), n = c.LU(RVN_LU_PONDED_EXP), dep_max = c.LU(RVN_LU_DEP_MAX);
    const double area_ponded = maxPondedAreaFrac * rvn_pow(dmin(c.SV(c.iSV(RVN_SV_DEPRESSION)) / dep_max, 1.0), n);
    c.R(0) = (1.0 - area_ponded) * c.R(0);
    c.R(1) = (area_ponded)*c.F[RVN_F_OW_PET];
    PETused = (c.R(0) + c.R(1)); multi = true;
  } else if (op == RVN_OP_SOILEVAP_CHU) {   // :460-468: evaporation scaled by the crop heat units accumulated towards maturity
    const double CHU = dmax(c.SV(c.iSV(RVN_SV_CROP_HEAT_UNITS)), 0.0);
    c.R(0) = PET * dmin(CHU / c.VEG(RVN_V_CHU_MATURITY), 1.0);
  } else if (op == RVN_OP_SOILEVAP_LINEAR) {   // :370-377
    c.R(0) = dmin(c.LU(RVN_LU_AET_COEFF) * c.SV(pr.from(0)), PET);
  } else if (op == RVN_OP_SOILEVAP_PDM || op == RVN_OP_SOILEVAP_HYMOD2) {   // :424-458
    const double stor = c.SV(pr.from(0)), max_stor = SoilCapacity(c, 0), b = c.LU(RVN_LU_PDM_B);
    const double c_max = (b + 1) * max_stor;
    const double c_star = c_max * (1.0 - rvn_pow(1.0 - (stor / max_stor), 1.0 / (b + 1.0)));
    if (op == RVN_OP_SOILEVAP_PDM) c.R(0) = dmin(c_star, (c_star / c_max) * PET);
    else {
      const double gp = c.LU(RVN_LU_HYMOD2_G), Kmax = c.LU(RVN_LU_HYMOD2_KMAX), ce = c.LU(RVN_LU_HYMOD2_EXP);
      const double K = Kmax * (gp + (1 - gp) * rvn_pow(c_star / c_max, ce));
      c.R(0) = K * dmin(c_star, PET);
    }
  } else if (op == RVN_OP_SOILEVAP_UBC) {   // :470-494 (Quick 1995, RFS emulation)
    const double P0AGEN = c.SOIL(0, RVN_S_UBC_INFIL_SOIL_DEF), P0EGEN = c.SOIL(0, RVN_S_UBC_EVAP_SOIL_DEF);
    const double soil_deficit = dmax(SoilCapacity(c, 0) - c.SV(pr.from(0)), 0.0);
    const double Fimp = c.LU(RVN_LU_IMPERMEABLE_FRAC);
    const double AET = (PET)*rvn_pow(10.0, -soil_deficit / P0EGEN);
    const double total_precip = c.SV(c.iSV(RVN_SV_PONDED_WATER));
    const double soil_deficit_est = dmax(soil_deficit - total_precip + AET * c.tstep, 0.0);
    double b1 = 0.0;
    if (Fimp < 1.0) b1 = Fimp * rvn_pow(10.0, -soil_deficit_est / P0AGEN);
    c.R(0) = (PET)*rvn_pow(10.0, -soil_deficit_est / P0EGEN) * (1.0 - b1);
  } else if (op == RVN_OP_SOILEVAP_VIC) {   // :496-515 (Woods et al. 1992)
    const double stor = c.SV(pr.from(0)), stor_max = SoilCapacity(c, 0);
    const double alpha = c.SOIL(0, RVN_S_VIC_ALPHA), zmax = c.SOIL(0, RVN_S_VIC_ZMAX), zmin = c.SOIL(0, RVN_S_VIC_ZMIN), gamma2 = c.SOIL(0, RVN_S_VIC_EVAP_GAMMA);
    const double Smax = 1.0 / (alpha + 1.0) * (alpha * zmax + zmin);
    const double Sat = stor / stor_max;
    c.R(0) = PET * (1.0 - rvn_pow(1.0 - Sat / Smax, gamma2));
  } else if (op == RVN_OP_SOILEVAP_SEQUEN) {   // :553-568
    const double stor_u = dmax(c.SV(pr.from(0)), 0.0), stor_l = dmax(c.SV(pr.from(1)), 0.0);
    const double tens_stor_u = SoilTensionCapacity(c, 0), tens_stor_l = SoilTensionCapacity(c, 1);
    c.R(0) = (PET)*dmin(stor_u / tens_stor_u, 1.0);
    c.R(1) = (PET - c.R(0)) * dmin(stor_l / tens_stor_l, 1.0);
    PETused = c.R(0) + c.R(1); multi = true;
  } else if (op == RVN_OP_SOILEVAP_ROOT) {   // :570-596
    const double rootfrac_u = 0.7, rootfrac_l = 1.0 - rootfrac_u;
    const double stor_u = c.SV(pr.from(0));
    double tens_stor_u = SoilTensionCapacity(c, 0); tens_stor_u = dmax(0.0001, tens_stor_u);
    c.R(0) = PET * rootfrac_u * dmin(stor_u / tens_stor_u, 1.0);
    const double stor_l = c.SV(pr.from(1));
    double tens_stor_l = SoilTensionCapacity(c, 1); tens_stor_l = dmax(0.0001, tens_stor_l);
    c.R(1) = PET * rootfrac_l * dmin(stor_l / tens_stor_l, 1.0);
    PETused = c.R(0) + c.R(1); multi = true;
  } else if (op == RVN_OP_SOILEVAP_ROOT_CONSTRAIN) {   // :598-622
    const double rootfrac_u = 0.7;
    const double stor_u = c.SV(pr.from(0)), tens_stor_u = SoilTensionCapacity(c, 0);
    const double stor_wilt = c.SOIL(0, RVN_S_SAT_WILT) * SoilCapacity(c, 0);
    c.R(0) = PET * rootfrac_u * dmin(stor_u / tens_stor_u, 1.0);
    if (c.R(0) > dmax(stor_u - stor_wilt, 0.0)) c.R(0) = dmax(stor_u - stor_wilt, 0.0);
    const double rootfrac_l = 1.0 - rootfrac_u;
    c.R(1) = PET * rootfrac_l * 1;
    PETused = c.R(0) + c.R(1); multi = true;
  } else if (op == RVN_OP_SOILEVAP_SACSMA) {   // :652-715 (Sacramento soil moisture accounting): UZT, UZF, LZT, LZFP, LZFS, ADIMC = SOIL[0..5], depths over the pervious area
    const double Adimp = c.LU(RVN_LU_MAX_SAT_AREA_FRAC), Aimp = c.LU(RVN_LU_IMPERMEABLE_FRAC), rserv = c.SOIL(3, RVN_S_UNAVAIL_FRAC);
    const double Aperv = 1.0 - Adimp - Aimp;
    const double uzt_stor_max = SoilCapacity(c, 0), uzf_stor_max = SoilCapacity(c, 1), lzt_stor_max = SoilCapacity(c, 2);
    const double lzfp_stor_max = SoilCapacity(c, 3), lzfs_stor_max = SoilCapacity(c, 4);
    double uzt_stor = c.SV(c.soil_slot(0)) / Aperv, uzf_stor = c.SV(c.soil_slot(1)) / Aperv, lzt_stor = c.SV(c.soil_slot(2)) / Aperv;
    const double lzfp_stor = c.SV(c.soil_slot(3)) / Aperv;
    double lzfs_stor = c.SV(c.soil_slot(4)) / Aperv, adimc_stor = c.SV(c.soil_slot(5)) / Adimp;
    if (Adimp == 0) adimc_stor = 0;
    const double e1 = dmin(PET * (uzt_stor / uzt_stor_max), uzt_stor);
    double red = PET - e1;
    uzt_stor -= e1;
    c.R(0) = e1 * Aperv / c.tstep;
    const double e2 = dmin(red, uzf_stor);
    red -= e2;
    uzf_stor -= e2;
    c.R(1) = e2 * Aperv / c.tstep;
    if ((uzt_stor / uzt_stor_max) < (uzf_stor / uzf_stor_max)) {   // equilibrate the storage ratios of the upper zone
      const double delta = uzt_stor_max * (uzt_stor + uzf_stor) / (uzt_stor_max + uzf_stor_max) - uzt_stor;
      uzt_stor += delta;
      uzf_stor -= delta;
      c.R(2) = delta * Aperv / c.tstep;
    }
    const double e3 = dmin(red * (lzt_stor / (uzt_stor_max + lzt_stor_max)), lzt_stor);
    lzt_stor -= e3;
    c.R(3) = e3 * Aperv / c.tstep;
    const double ratlzt = lzt_stor / lzt_stor_max;
    const double saved = rserv * (lzfp_stor_max + lzfs_stor_max);
    const double ratlz = (lzt_stor + lzfp_stor + lzfs_stor - saved) / (lzt_stor_max + lzfp_stor_max + lzfs_stor_max - saved);
    if (ratlzt < ratlz) {   // resupply of the lower-zone tension water: bookkeeping only in the reference (its rates[4] stays 0)
      const double del = dmin((ratlz - ratlzt) * lzt_stor_max, lzfs_stor);
      lzt_stor += del;
      lzfs_stor -= del;
    }
    const double e5 = dmin(e1 + (red + e2) * (adimc_stor - uzt_stor - e1) / (uzt_stor_max + lzt_stor_max), adimc_stor);
    adimc_stor -= e5;
    c.R(5) = e5 * Adimp / c.tstep;
    PETused = (e1 + e2 + e3) * Aperv + e5 * Adimp; multi = true;
  } else if (op == RVN_OP_SOILEVAP_AWBM) {   // :716-727 (Boughton 2004)
    const double a1 = c.LU(RVN_LU_AWBM_AREAFRAC1), a2 = c.LU(RVN_LU_AWBM_AREAFRAC2);
    const double a3 = 1.0 - a1 - a2;
    c.R(0) = dmin(a1 * PET, c.SV(pr.from(0)) / c.tstep);
    c.R(1) = dmin(a2 * PET, c.SV(pr.from(1)) / c.tstep);
    c.R(2) = dmin(a3 * PET, c.SV(pr.from(2)) / c.tstep);
    PETused = c.R(0) + c.R(1) + c.R(2); multi = true;
  }
  const int last = pr.nconn() - 1;
  c.R(last) = multi ? PETused : c.R(0);
  double corr = 0;
#pragma unroll
  for (int q = 0; q < last; q++) {
    double oldrate = c.R(q);
    c.R(q) = dmin(c.R(q), c.SV(pr.from(q)) / c.tstep);
    corr += oldrate - c.R(q);
  }
  c.R(last) -= corr;
}

// CmvSnowBalance SNOBAL_COLD_CONTENT  src/SnowBalance.cpp:732-835 (ColdContentBalance, adapted there from Brook90 SNOENRGY / SNOPACK);
// connections :41-55: COLD_CONTENT -> ENERGY_LOSSES, SNOW_LIQ -> SNOW, ENERGY_LOSSES -> COLD_CONTENT, SNOW -> PONDED, SNOW_LIQ -> PONDED
template <class C, class PR> RVN_DI void rates_SNOBAL_COLD_CONTENT(C &c, const PR &pr) {
  const double CCFAC = 0.3, MELT_FAC = 1.5, LAIMLT = 0.2, SAIMLT = 0.5, SPH_ICE = 2.100e-3;
  const double SWI = c.G(RVN_G_SNOW_SWI), tstep = c.tstep;
  const double Ta = c.F[RVN_F_TEMP_DAILY_AVE], day_length = c.day_length;
  const double CC = c.SV(pr.from(0)), S = c.SV(pr.from(3));
  double SL = c.SV(pr.from(1));
  const double rainthru = 0.0;
  if (S == 0) return;
  if ((CC > 0) && (SL > 0)) {   // instant equilibrium of cold content and liquid water
    const double instant_refreeze = dmin(SL / tstep, CC / D_LH_FUSION / tstep);
    c.R(1) += instant_refreeze;
    c.R(0) -= instant_refreeze * D_LH_FUSION;
    SL -= c.R(1) * tstep;
  }
  const double Tsnow = D_FREEZING_TEMP - CC / S / SPH_ICE;
  double incoming_snow_en;
  if (Ta <= D_FREEZING_TEMP) incoming_snow_en = CCFAC * 2.0 * day_length * (Ta - Tsnow);
  else incoming_snow_en = MELT_FAC * 2.0 * day_length * (Ta - D_FREEZING_TEMP) * rvn_exp(-SAIMLT * c.veg_SAI) * rvn_exp(-LAIMLT * c.veg_LAI);
  incoming_snow_en += rainthru * dmax(Ta - D_FREEZING_TEMP, 0.0) * D_SPH_WATER * D_DENSITY_WATER / D_MM_PER_METER;
  double pot_melt = incoming_snow_en * (1.0 / D_DENSITY_WATER / D_LH_FUSION * D_MM_PER_METER);
  const double CC_air = (D_FREEZING_TEMP - Ta) * SPH_ICE * S;
  const double liq_snow_cap = SWI * S;   // CalculateSnowLiquidCapacity, src/SnowParams.cpp:79-88
  if (pot_melt <= 0) {   // negative energy balance: refreeze, then cool (not below the air temperature)
    const double refreeze = dmin(SL / tstep, -pot_melt);
    c.R(1) += refreeze;
    pot_melt += refreeze;
    double cooling = dmin(-pot_melt, 0.0) * D_LH_FUSION;
    cooling = dmin(cooling, (CC_air - CC) / tstep);
    c.R(2) += cooling;
  } else if ((pot_melt * D_LH_FUSION < CC / tstep) || (Ta < D_FREEZING_TEMP)) {   // energy only reduces the cold content
    double warming = pot_melt * D_LH_FUSION;
    warming = dmin(warming, -(CC - CC_air) / tstep);
    c.R(0) -= warming;
  } else {   // all cold content lost, the rest melts: into the pack's liquid capacity first, then out
    double eq_en_avail = pot_melt;
    const double warming = CC / tstep;
    c.R(0) += warming;
    eq_en_avail -= warming / D_LH_FUSION;
    const double melt_to_liq = dmin(eq_en_avail, (liq_snow_cap - SL) / tstep);
    c.R(1) -= melt_to_liq;
    eq_en_avail -= melt_to_liq;
    const double melt_to_SW = dmin(eq_en_avail, S / tstep - melt_to_liq);
    c.R(3) += melt_to_SW;
    c.R(4) += melt_to_SW * SWI;
  }
}

// CmvCropHeatUnitEvolve CHU_ONTARIO  src/CropGrowth.cpp:91-136: -1 outside the growing season, reset to 0 when the daily mean exceeds 12.8 C, then daily day / night increments
template <class C, class PR> RVN_DI void rates_CHU_ONTARIO(C &c, const PR &pr) {
  if (c.type != RVN_HRU_STANDARD) return;
  double CHU = c.SV(pr.from(0));
  const double old_CHU = CHU;
  const double temp_ave = c.F[RVN_F_TEMP_DAILY_AVE], temp_max = c.F[RVN_F_TEMP_DAILY_MAX], temp_min = c.F[RVN_F_TEMP_DAILY_MIN];
  const bool growing_season = (CHU > -0.5);
  if (!growing_season) { if (temp_ave > 12.8) CHU = 0.0; }
  else {
    const double CHU_day = dmax(3.33 * (temp_max - 10) - 0.084 * rvn_pow(temp_max - 10.0, 2.0), 0.0);
    const double CHU_night = dmax(1.8 * (temp_min - 4.4), 0.0);
    CHU += 0.5 * (CHU_day + CHU_night);
    if (temp_ave < -2.0) CHU = -1.0;
    if (c.month > 9) CHU = -1.0;
  }
  c.R(0) = (CHU - old_CHU) / c.tstep;
}
// CmvSnowBalance SNOBAL_GAWSER  src/SnowBalance.cpp:1098-1201 (GawserBalance); connections :123-145: SNOW -> SNOW_LIQ, PONDED -> SNOW_LIQ, SNOW_LIQ -> SNOW,
// three SNOW_DEPTH self-connections (compaction, melt + sublimation, snowfall), NEW_SNOW -> SNOW, SNOW -> ATMOSPHERE, SNOW_LIQ -> PONDED, SNOW -> PONDED
template <class C, class PR> RVN_DI void rates_SNOBAL_GAWSER(C &c, const PR &pr) {
  const double tstep = c.tstep;
  const double SWC = c.SV(pr.from(0)), rainthru = c.SV(pr.from(1)), snow_liq = c.SV(pr.from(2)), snow_depth = c.SV(pr.from(3)), newSnow = c.SV(pr.from(6));
  const double KF = c.LU(RVN_LU_REFREEZE_FACTOR), KM = c.LU(RVN_LU_MELT_FACTOR), SWI = c.G(RVN_G_SNOW_SWI);
  const double RHOICE = D_DENSITY_ICE / D_DENSITY_WATER, MRHO = 0.35, a = 0.1, b = 96, NEWDEN = 0.1;
  const double TEMPs = c.F[RVN_F_TEMP_AVE], TBAS = 0;
  double MELTP = 0, REFRZP = 0;
  if (TEMPs < TBAS) REFRZP = tstep * KF * (TBAS - TEMPs);
  if (TEMPs > TBAS) MELTP = tstep * KM * (TEMPs - TBAS);
  double RHO = 0;
  if (SWC > 0) RHO = dmin(MRHO, SWC / snow_depth);
  const double POR = 1 - RHO / RHOICE;
  const double snow_liqAP = POR * snow_depth * SWI;
  double RAIN_IN_1 = 0, MELT_1 = 0, MELT_2 = 0, EXCESS_snow_liq = 0, REFRZ = 0;
  if (TBAS > TEMPs) REFRZ = dmin(REFRZP, snow_liq);
  const double SWC2 = SWC + REFRZ;
  const double RHO2 = dmin(MRHO, SWC2 / snow_depth);
  const double snow_liq2 = snow_liq - REFRZ;
  if (snow_liq2 >= snow_liqAP) { EXCESS_snow_liq = snow_liq2 - snow_liqAP; MELT_2 = dmin(MELTP, SWC); }
  else { MELT_1 = dmin(MELTP, SWC); RAIN_IN_1 = rainthru; }
  const double KC = b * rvn_exp(-1 * a * TEMPs);
  const double TRHO = (RHO2 * MRHO) / (RHO2 + (MRHO - RHO2) * rvn_exp(-1 * tstep * 24 / KC));
  double compaction = 0, snowmelt = 0, SUBLIM = 0, SUBLIM_snow_depth = 0;
  if (SWC2 > 0) compaction = snow_depth - SWC2 / TRHO;
  if (SWC2 > 0) snowmelt = (MELT_1 + MELT_2) / TRHO;
  if (TEMPs < 0) SUBLIM = 2.4 * tstep;
  if (SWC2 - MELT_1 - MELT_2 < SUBLIM) SUBLIM = dmax(dmin(SUBLIM, SWC2 - MELT_1 - MELT_2), 0.0);
  if (SWC2 > 0) SUBLIM_snow_depth = SUBLIM / TRHO;
  c.R(0) = MELT_1 / tstep;
  c.R(1) = RAIN_IN_1 / tstep;
  c.R(2) = REFRZ / tstep;
  c.R(3) = -compaction / tstep;
  c.R(4) = -(snowmelt + SUBLIM_snow_depth) / tstep;
  c.R(5) = newSnow / NEWDEN / tstep;
  c.R(6) = newSnow / tstep;
  c.R(7) = SUBLIM / tstep;
  c.R(8) = EXCESS_snow_liq / tstep;
  c.R(9) = MELT_2 / tstep;
}

// CmvSnowBalance SNOBAL_CRHM_EBSM  src/SnowBalance.cpp:1214-1291 (CRHMSnowBalance; amounts per step are handed over as rates exactly as written there);
// connections :146-158: SNOW -> SNOW_LIQ, SNOW -> PONDED, SNOW_LIQ -> PONDED, SNOW_LIQ -> SNOW, COLD_CONTENT self-connection
template <class C, class PR> RVN_DI void rates_SNOBAL_CRHM_EBSM(C &c, const PR &pr) {
  const double snow_liq = c.SV(c.iSV(RVN_SV_SNOW_LIQ)), SWE = c.SV(c.iSV(RVN_SV_SNOW));
  double snow_energy = -c.SV(c.iSV(RVN_SV_COLD_CONTENT));
  const double snow_energy_old = snow_energy;
  if (SWE <= 0.0) return;
  double pot_melt = c.F[RVN_F_POTENTIAL_MELT];
  const double T_min = c.F[RVN_F_TEMP_DAILY_MIN];
  pot_melt *= D_LH_FUSION * D_DENSITY_WATER / D_MM_PER_METER * c.tstep;
  const double snoliq_max = c.G(RVN_G_SNOW_SWI) * SWE;
  const double t_minus = dmin(T_min, 0.0);
  const double Umin = SWE * (2.115 + 0.00779 * t_minus) * t_minus;
  double refreeze = 0.0, snowmelt0 = 0.0, snowmelt1 = 0.0, snowmelt2 = 0.0;
  snow_energy += pot_melt;
  snow_energy = dmax(Umin, snow_energy);
  if (snow_energy > 0.0) {
    const double B = 0.95;
    const double melt = snow_energy / D_LH_FUSION / D_DENSITY_WATER * D_MM_PER_METER / B;
    if (melt > (snoliq_max - snow_liq)) {
      if (SWE > melt) { snowmelt0 = snoliq_max - snow_liq; snowmelt1 = melt - (snoliq_max - snow_liq); snowmelt2 = 0.0; }
      else { snowmelt0 = 0; snowmelt1 = SWE; snowmelt2 = snow_liq; }
    } else { snowmelt0 = melt; snowmelt1 = 0; snowmelt2 = 0; }
    snow_energy = 0.0;
  } else {
    refreeze = -snow_energy / D_LH_FUSION / D_DENSITY_WATER * D_MM_PER_METER;
    if (snow_liq > refreeze) snow_energy = 0.0;
    else { snow_energy += snow_liq * D_LH_FUSION * D_DENSITY_WATER / D_MM_PER_METER; refreeze = snow_liq; }
  }
  c.R(0) = snowmelt0; c.R(1) = snowmelt1; c.R(2) = snowmelt2; c.R(3) = refreeze;
  c.R(4) = -(snow_energy - snow_energy_old) / c.tstep;
  (void)pr;
}

// CmvSnowBalance SNOBAL_HBV  src/SnowBalance.cpp:430-465 (connections :56-70; CalculateSnowLiquidCapacity src/SnowParams.cpp:79-88)
template <class C, class PR> RVN_DI void rates_SNOBAL_HBV(C &c, const PR &pr) {
  const double Ka = c.LU(RVN_LU_REFREEZE_FACTOR), Ta = c.F[RVN_F_TEMP_DAILY_AVE], tstep = c.tstep;
  double SWE = c.SV(pr.from(0)), SL = c.SV(pr.to(0));
  double melt = dmax(c.F[RVN_F_POTENTIAL_MELT], 0.0);
  melt = dmin(melt, SWE / tstep);
  SWE -= melt * tstep;
  double refreeze = Ka * dmax(D_FREEZING_TEMP - Ta, 0.0);
  refreeze = dmax(dmin(SL / tstep, refreeze), 0.0);
  SL -= refreeze * tstep;
  const double liq_cap = c.G(RVN_G_SNOW_SWI) * SWE;
  const double to_liq = dmin(melt, dmax(liq_cap - SL, 0.0) / tstep);
  SL += to_liq * tstep;
  const double overflow = dmax((SL - liq_cap) / tstep, 0.0);
  c.R(0) = -refreeze;
  c.R(0) += to_liq;
  c.R(1) = dmax(melt - to_liq, 0.0);
  c.R(2) = overflow;
  if (c.type != RVN_HRU_STANDARD) { c.R(3) = c.R(1); c.R(1) = 0.0; c.R(4) = c.R(2); c.R(2) = 0.0; }
}
// CmvInterflow INTERFLOW_PRMS  src/Interflow.cpp:98-124; ApplyConstraints :136-148
template <class C, class PR> RVN_DI void rates_INTERFLOW_PRMS(C &c, const PR &pr) {
  const int type = c.type;
  if (type == RVN_HRU_LAKE || type == RVN_HRU_WATER || type == RVN_HRU_ROCK) return;
  const int m = pr.ia(0);
  const double stor = c.SV(pr.from(0)), max_stor = SoilCapacity(c, m), tens_stor = SoilTensionCapacity(c, m);
  c.R(0) = c.SOIL(m, RVN_S_MAX_INTERFLOW_RATE) * dmax(stor - tens_stor, 0.0) / (max_stor - tens_stor);
  c.R(0) = dmin(c.R(0), c.SV(pr.from(0)) / c.tstep);
}
// CmvSnowMelt MELT_POTMELT  src/SnowMeltRefreeze.cpp:87-131
template <class C, class PR> RVN_DI void rates_SNOWMELT_POTMELT(C &c, const PR &pr) {
  c.R(0) = dmax(c.F[RVN_F_POTENTIAL_MELT], 0.0);
  if (c.SV(pr.from(0)) <= 0) c.R(0) = 0.0;
  if (c.R(0) < 0.0) c.R(0) = 0.0;
  c.R(0) = dmin(c.R(0), c.SV(pr.from(0)) / c.tstep);
}
// CmvSnowSqueeze  src/SnowMeltRefreeze.cpp:197-244
template <class C, class PR> RVN_DI void rates_SNOWSQUEEZE(C &c, const PR &pr) {
  const double S = c.SV(c.iSV(RVN_SV_SNOW)), SL = c.SV(c.iSV(RVN_SV_SNOW_LIQ));
  const double liq_cap = c.G(RVN_G_SNOW_SWI) * S;
  c.R(0) = dmax(SL - liq_cap, 0.0) / c.tstep;
  if (c.SV(pr.from(0)) <= 0) c.R(0) = 0.0;
  if (c.R(0) < 0.0) c.R(0) = 0.0;
}

// GetRatesOfChange + ApplyConstraints of process `pr` (opcode `op`: a compile-time constant in the specialised kernels)
template <class C, class PR> RVN_DI void dispatch_rates(C &c, const PR &pr, const int op) {
  switch (op) {   // uniform across the CTA: every thread runs the same process at the same time
    case RVN_OP_PRECIPITATION: rates_PRECIPITATION(c, pr); break;
    case RVN_OP_SNOBAL_HMETS: rates_SNOBAL_HMETS(c, pr); break;
    case RVN_OP_SNOBAL_COLD_CONTENT: rates_SNOBAL_COLD_CONTENT(c, pr); break;
    case RVN_OP_SNOBAL_GAWSER: rates_SNOBAL_GAWSER(c, pr); break;
    case RVN_OP_CHU_ONTARIO: rates_CHU_ONTARIO(c, pr); break;
    case RVN_OP_EXCHANGE_FLOW: rates_EXCHANGE_FLOW(c, pr); break;
    case RVN_OP_DEPRESSION_OVERFLOW: rates_DEPRESSION_OVERFLOW(c, pr); break;
    case RVN_OP_SEEPAGE: rates_SEEPAGE(c, pr); break;
    case RVN_OP_SNOBAL_CRHM_EBSM: rates_SNOBAL_CRHM_EBSM(c, pr); break;
    case RVN_OP_SNOBAL_SIMPLE_MELT: rates_SNOBAL_SIMPLE_MELT(c, pr); break;
    case RVN_OP_SNOBAL_CEMA_NEIGE: rates_SNOBAL_CEMA_NEIGE(c, pr); break;
    case RVN_OP_SNOWREFREEZE_DD: rates_SNOWREFREEZE_DD(c, pr); break;
    case RVN_OP_SNOTEMP_NEWTONS: rates_SNOTEMP_NEWTONS(c, pr); break;
    case RVN_OP_INF_HMETS: case RVN_OP_INF_HBV: case RVN_OP_INF_GR4J: case RVN_OP_INF_VIC_ARNO: case RVN_OP_INF_RATIONAL:
    case RVN_OP_INF_ALL_INFILTRATES: case RVN_OP_INF_VIC: case RVN_OP_INF_PRMS: case RVN_OP_INF_PDM: case RVN_OP_INF_GREEN_AMPT: case RVN_OP_INF_GA_SIMPLE: case RVN_OP_INF_UBC: case RVN_OP_INF_SCS: case RVN_OP_INF_SCS_NOABSTRACTION: rates_INFILTRATION(c, pr, op); break;
    case RVN_OP_SNOBAL_HBV: rates_SNOBAL_HBV(c, pr); break;
    case RVN_OP_INTERFLOW_PRMS: rates_INTERFLOW_PRMS(c, pr); break;
    case RVN_OP_SNOWMELT_POTMELT: rates_SNOWMELT_POTMELT(c, pr); break;
    case RVN_OP_SNOWSQUEEZE: rates_SNOWSQUEEZE(c, pr); break;
    case RVN_OP_OVERFLOW: rates_OVERFLOW(c, pr); break;
    case RVN_OP_FLUSH: rates_FLUSH(c, pr); break;
    case RVN_OP_SPLIT: rates_SPLIT(c, pr); break;
    case RVN_OP_BASE_LINEAR: case RVN_OP_BASE_POWER_LAW: case RVN_OP_BASE_GR4J: case RVN_OP_BASE_LINEAR_ANALYTIC: case RVN_OP_BASE_VIC:
    case RVN_OP_BASE_TOPMODEL: case RVN_OP_BASE_LINEAR_CONSTRAIN: case RVN_OP_BASE_CONSTANT: case RVN_OP_BASE_THRESH_POWER:
    case RVN_OP_BASE_THRESH_STOR: rates_BASEFLOW(c, pr, op); break;
    case RVN_OP_PERC_LINEAR: case RVN_OP_PERC_CONSTANT: case RVN_OP_PERC_GR4J: case RVN_OP_PERC_GR4JEXCH: case RVN_OP_PERC_GR4JEXCH2:
    case RVN_OP_PERC_GAWSER: case RVN_OP_PERC_GAWSER_CONSTRAIN: case RVN_OP_PERC_POWER_LAW: case RVN_OP_PERC_LINEAR_ANALYTIC: case RVN_OP_PERC_PRMS:
    case RVN_OP_PERC_SACRAMENTO: case RVN_OP_PERC_ASPEN: rates_PERCOLATION(c, pr, op); break;
    case RVN_OP_SOILEVAP_ALL: case RVN_OP_SOILEVAP_HBV: case RVN_OP_SOILEVAP_GR4J: case RVN_OP_SOILEVAP_TOPMODEL:
    case RVN_OP_SOILEVAP_LINEAR: case RVN_OP_SOILEVAP_PDM: case RVN_OP_SOILEVAP_HYMOD2: case RVN_OP_SOILEVAP_UBC: case RVN_OP_SOILEVAP_VIC:
    case RVN_OP_SOILEVAP_HYPR: case RVN_OP_SOILEVAP_SEQUEN: case RVN_OP_SOILEVAP_ROOT: case RVN_OP_SOILEVAP_ROOT_CONSTRAIN: case RVN_OP_SOILEVAP_AWBM: case RVN_OP_SOILEVAP_SACSMA: case RVN_OP_SOILEVAP_CHU:
      rates_SOILEVAP(c, pr, op); break;
    case RVN_OP_CANEVP_ALL: rates_CANEVP_ALL(c, pr); break;
    case RVN_OP_CANSNOWEVP_ALL: rates_CANSNOWEVP_ALL(c, pr); break;
    case RVN_OP_OPEN_WATER_EVAP: rates_OPEN_WATER_EVAP(c, pr); break;
    case RVN_OP_OPEN_WATER_RIPARIAN: rates_OPEN_WATER_EVAP<true>(c, pr); break;
    case RVN_OP_SUBLIMATION: rates_SUBLIMATION(c, pr); break;
    case RVN_OP_ABSTRACTION: rates_ABSTRACTION(c, pr); break;
    case RVN_OP_CANEVP_MAXIMUM: rates_CANEVP_PET<0>(c, pr); break;
    case RVN_OP_CANEVP_RUTTER: rates_CANEVP_PET<1>(c, pr); break;
    case RVN_OP_CANSNOWEVP_MAXIMUM: rates_CANEVP_PET<2>(c, pr); break;
    case RVN_OP_LAKE_EVAP_BASIC: rates_LAKE_EVAP_BASIC(c, pr); break;
    case RVN_OP_CAPRISE_HBV: rates_CAPRISE_HBV(c, pr); break;
    case RVN_OP_GLACIERMELT_HBV: rates_GLACIERMELT_HBV(c, pr); break;
    case RVN_OP_GLACIERMELT_SIMPLE: rates_GLACIERMELT_SIMPLE(c, pr); break;
    case RVN_OP_GLACIERMELT_UBC: rates_GLACIERMELT_UBC(c, pr); break;
    case RVN_OP_GLACIERRELEASE_LINEAR: rates_GLACIERRELEASE_LINEAR(c, pr); break;
    case RVN_OP_GLACIERRELEASE_HBVEC: rates_GLACIERRELEASE_HBVEC(c, pr); break;
    default: break;
  }
}
__host__ __device__ inline bool rvn_op_supported(int op) { return op > RVN_OP_NOP && op < RVN_OP_COUNT; }
#endif
