/* oracle/raven_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement, in plain C, of the reference's
 * per-timestep water balance for the supported option/process subset.  It consumes the same flattened
 * rvn_model_desc as the CUDA engine and follows the REFERENCE's loop structure (one HRU after another,
 * one process after another, per-connection state update), so that tests can compare the GPU path
 * against it on seeded inputs of any size.  Nothing under ravenhydroframework_b200/ may link or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg do.
 *
 * Parity pin: tests/test_oracle_golden.py and tests/test_reference_inputs.py check this oracle against in-memory FP64 dumps of the
 * unmodified reference (oracle/_ref/ref_dump, built from /root/reference by oracle/build_ref.sh) and
 * against the committed fixtures under tests/golden/ that the same driver generated.
 *
 * Each function cites the reference file:line it restates (paths relative to the reference tree).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "rvn_gpu.h"

#define O_ALMOST_INF 1e99
#define O_REAL_SMALL 1e-12
#define O_TIME_CORRECTION 0.0001
#define O_PI 3.1415926535898
#define O_SEC_PER_DAY 86400.0
#define O_MM_PER_METER 1000.0
#define O_M2_PER_KM2 1e6
#define O_FREEZING_TEMP 0.0
#define O_NEGLIGBLE_SNOW 0.1
#define O_TYPICAL_SNOW_DENS 0.250e3
#define O_DENSITY_WATER 1.000e3
#define O_DENSITY_ICE 0.917e3
#define O_HCP_ICE 1.938
#define O_WINTER_SOLSTICE_ANG 6.111043
#define O_HBV_PET_ELEV_CORR 0.0005
#define O_HBV_PET_TEMP_CORR 0.5
#define O_GLOBAL_PET_CORR 1.0
#define O_MAXCONN 256

/* src/RavenInclude.h:1456-1464 -- ties return the FIRST argument */
static double *g_cum_flux = 0; /* optional [nConn][nHRU] cumulative amount moved through every process connection (CModel::IncrementBalance, src/Model.cpp:2428-2436); rvo_set_flux_buffer */
static double *g_precip5 = 0; /* [nHRU] force_struct::precip_5day of the current step (update_forcings -> INF_SCS) */
static double omax(double r, double l) { return (r >= l) ? r : l; }
static double omin(double r, double l) { return (r <= l) ? r : l; }

typedef struct {
  const rvn_model_desc *d;
  int member;
  const double *ptab;     /* this member's parameter row */
  int k;                  /* HRU */
  int step;               /* model step (calendar record d->times[step]) */
  const double *F;        /* derived forcings of this HRU [RVN_F_COUNT] */
  const double *phi_old;  /* stored (start-of-step) state: what pHRU->GetStateVarValue returns */
  double canopy_cap, canopy_snow_cap; /* veg_var_struct capacity / snow_capacity for this HRU and step */
  double rain_icept_pct, snow_icept_pct; /* veg_var_struct interception fractions, src/VegetationParams.cpp:143-166 */
  double veg_LAI, veg_SAI, day_length;   /* veg_var_struct LAI / SAI and force_struct day_length of this HRU and step (SNOBAL_COLD_CONTENT) */
  int iSV[RVN_SV_NTYPES]; /* first index of each SV type, or -1 */
} octx;

/* CHydroUnit::IsLinkedToReservoir (src/HydroUnits.cpp:194; set in src/ModelInitialize.cpp:165-175) */
static int hru_linked(const rvn_model_desc *d, int k) {
  for (int r = 0; r < d->nRes; r++) if (d->res_hru[r] == k) return 1;
  return 0;
}
static int sv_index(const rvn_model_desc *d, int type, int layer) {
  for (int i = 0; i < d->NS; i++) if (d->sv_type[i] == type && (layer < 0 || d->sv_layer[i] == layer)) return i;
  return -1;
}
static double P_G(const octx *c, int f) { return c->ptab[c->d->glob_off + f]; }
static double P_LU(const octx *c, int f) { return c->ptab[c->d->lu_off + c->d->hru_lu[c->k] * RVN_LU_COUNT + f]; }
static double P_V(const octx *c, int f) { return c->ptab[c->d->veg_off + c->d->hru_veg[c->k] * RVN_V_COUNT + f]; }
static double P_S(const octx *c, int m, int f) {
  int cls = c->d->prof_class[c->d->hru_profile[c->k] * RVN_MAX_SOIL_LAYERS + m];
  return c->ptab[c->d->soil_off + cls * RVN_S_COUNT + f];
}
/* CHydroUnit::GetSoilCapacity, src/HydroUnits.cpp:260-267 */
/* LambertN, src/CommonFunctions.cpp:1572-1579 */
static double o_lambert_n(double x, int N) {
  double sigma, tmp;
  sigma = pow(-2.0 - 2.0 * log(-x), 0.5);
  if (N <= 2) tmp = -1.0 - 0.5 * sigma * sigma - sigma / (1.0 + sigma / 6.3);
  else tmp = o_lambert_n(x, N - 1);
  return (1 + log(-x) - log(-tmp)) * tmp / (1 + tmp);
}
/* CmvInfiltration::GreenAmptCumInf, src/GreenAmpt.cpp:26-52 */
static double o_green_ampt_cum_inf(double t, double alpha, double Ks, double w) {
  if (Ks == 0) return 0.0;
  if (w == 0) return 0.0;
  if (alpha == 0) return 0.0;
  double tp = alpha * Ks / w / (w - Ks);
  if ((t < tp) || (w < Ks)) return w * t;
  double tstar, x;
  double Fp = w * tp;
  tstar = Ks / alpha * (t - tp);
  x = -(1.0 + Fp / alpha) * exp(-(1.0 + tstar + Fp / alpha));
  return alpha * (-1.0 - o_lambert_n(x, 3));
}
static double soil_capacity(const octx *c, int m) {
  double thick = c->d->prof_thick[c->d->hru_profile[c->k] * RVN_MAX_SOIL_LAYERS + m];
  return thick * O_MM_PER_METER * P_S(c, m, RVN_S_CAP_RATIO);
}
/* CHydroUnit::GetSoilTensionStorageCapacity, src/HydroUnits.cpp:275-286 */
static double soil_tension_capacity(const octx *c, int m) {
  double thick = c->d->prof_thick[c->d->hru_profile[c->k] * RVN_MAX_SOIL_LAYERS + m];
  return thick * O_MM_PER_METER * P_S(c, m, RVN_S_CAP_RATIO) * (P_S(c, m, RVN_S_FIELD_CAPACITY) - P_S(c, m, RVN_S_SAT_WILT));
}
/* CHydroUnit::GetSnowCover, src/HydroUnits.cpp:679-705 (SNOWCOV_NONE), lagged */
static double snow_cover(const octx *c) {
  int iSn = c->iSV[RVN_SV_SNOW], iSC = c->iSV[RVN_SV_SNOW_COVER];
  if (iSn < 0) return 0.0;
  if (iSC >= 0) return c->phi_old[iSC];
  if (c->phi_old[iSn] < O_NEGLIGBLE_SNOW) return 0.0;
  return 1.0;
}
/* CHydroUnit::GetSnowDepth, src/HydroUnits.cpp:626-636, lagged */
static double snow_depth(const octx *c) {
  int iSn = c->iSV[RVN_SV_SNOW], iSD = c->iSV[RVN_SV_SNOW_DEPTH];
  if (iSn < 0) return 0.0;
  if (iSD >= 0) return c->phi_old[iSD];
  return (c->phi_old[iSn] / O_TYPICAL_SNOW_DENS) * O_DENSITY_WATER;
}
/* CHydroUnit::GetSnowTemperature, src/HydroUnits.cpp:645-658, lagged */
static double snow_temperature(const octx *c) {
  int i = c->iSV[RVN_SV_SNOW_TEMP];
  if (i < 0) {
    double t = P_G(c, RVN_G_SNOW_TEMPERATURE);
    if (t < -60000.0) return 0.0; /* NOT_NEEDED_AUTO / unspecified */
    return t;
  }
  return c->phi_old[i];
}
/* CHydroUnit::GetStateVarMax, src/HydroUnits.cpp:559-606 */
static double state_var_max(const octx *c, int i, const double *sv) {
  const rvn_model_desc *d = c->d;
  switch (d->sv_type[i]) {
    case RVN_SV_SOIL: return soil_capacity(c, d->sv_layer[i]);
    case RVN_SV_CANOPY: return c->canopy_cap;
    case RVN_SV_CANOPY_SNOW: return c->canopy_snow_cap;
    case RVN_SV_DEPRESSION: { double dm = P_LU(c, RVN_LU_DEP_MAX); return (dm >= 0) ? dm : O_ALMOST_INF; }
    case RVN_SV_SNOW_COVER: return 1.0;
    case RVN_SV_SNOW_LIQ: { double SWE = sv[c->iSV[RVN_SV_SNOW]]; (void)snow_depth; return P_G(c, RVN_G_SNOW_SWI) * SWE; } /* CalculateSnowLiquidCapacity, src/SnowParams.cpp:79-88 */
    default: return O_ALMOST_INF;
  }
}
/* CStateVariable::IsWaterStorage(typ), src/StateVariables.cpp:612-645 (conv_coverup=true) */
static int is_water_storage(int t) {
  switch (t) {
    case RVN_SV_SURFACE_WATER: case RVN_SV_PONDED_WATER: case RVN_SV_ATMOSPHERE: case RVN_SV_ATMOS_PRECIP: case RVN_SV_SNOW:
    case RVN_SV_GROUNDWATER: case RVN_SV_SOIL: case RVN_SV_CANOPY: case RVN_SV_CANOPY_SNOW: case RVN_SV_ROOT: case RVN_SV_TRUNK:
    case RVN_SV_DEPRESSION: case RVN_SV_SNOW_LIQ: case RVN_SV_WETLAND: case RVN_SV_GLACIER: case RVN_SV_GLACIER_ICE: case RVN_SV_FIRN:
    case RVN_SV_CONVOLUTION: case RVN_SV_NEW_SNOW: case RVN_SV_SNOW_DRIFT: case RVN_SV_LAKE_STORAGE: return 1;
    default: return 0;
  }
}

/* ======================================================================== forcings ============= */
/* CModel::UpdateHRUForcingFunctions, src/UpdateForcings.cpp:30-668, for gauge-based inputs and the
 * option subset {RAINSNOW_DATA|DINGMAN|HBV|UBCWM|THRESHOLD, POTMELT_*, PET_DATA|FROMMONTHLY|CONSTANT|NONE,
 * OROCORR_NONE|SIMPLELAPSE|HBV}, daily-or-longer... (timestep>=1 => subdaily_corr = 1, :949-958). */
/* CModel::ApplyForcingPerturbation, src/Model.cpp:2856-2879, with CHydroUnit::AdjustHRUForcing (src/HydroUnits.cpp:461-509)
 * and AdjustDailyHRUForcings (:513-538).  Daily-or-longer steps: nStepsPerDay = 1, nn = 0, every step starts a day. */
static void apply_forcing_perturbation(const rvn_model_desc *d, int ftype, int member, int k, int step, double *F) {
  for (int i = 0; i < d->nPert; i++) {
    if (d->pert_forcing[i] != ftype) continue;
    if (d->pert_mask != NULL && !d->pert_mask[(size_t)i * d->nHRU + k]) continue; /* (kk==DOESNT_EXIST) || IsInGroup(k) */
    const int nStepsPerDay = 1;
    const double *eps = &d->pert_eps[((size_t)member * d->nSteps + step) * d->nPert + i]; /* eps[n], n < nStepsPerDay */
    const double epsilon = eps[0];
    const int adj = d->pert_adj[i];
    /* AdjustHRUForcing */
    if (ftype == RVN_PF_PRECIP) {
      if (adj == RVN_ADJ_MULTIPLICATIVE) F[RVN_F_PRECIP] *= epsilon;
      else if (adj == RVN_ADJ_ADDITIVE) F[RVN_F_PRECIP] += epsilon;
      else if (adj == RVN_ADJ_REPLACE) F[RVN_F_PRECIP] = epsilon;
      if (0.0 > F[RVN_F_PRECIP]) F[RVN_F_PRECIP] = 0.0;
    } else if (ftype == RVN_PF_RAINFALL) {
      double sf = F[RVN_F_SNOW_FRAC], Po = F[RVN_F_PRECIP];
      if (adj == RVN_ADJ_MULTIPLICATIVE) F[RVN_F_PRECIP] += Po * (1.0 - sf) * (epsilon - 1.0);
      else if (adj == RVN_ADJ_ADDITIVE) F[RVN_F_PRECIP] += epsilon;
      else if (adj == RVN_ADJ_REPLACE) F[RVN_F_PRECIP] += epsilon - (1 - sf) * Po;
      if (0.0 > F[RVN_F_PRECIP]) F[RVN_F_PRECIP] = 0.0;
      if (F[RVN_F_PRECIP] == 0) F[RVN_F_SNOW_FRAC] = 0;
      else F[RVN_F_SNOW_FRAC] = F[RVN_F_SNOW_FRAC] * Po / F[RVN_F_PRECIP];
    } else if (ftype == RVN_PF_SNOWFALL) {
      double sf = F[RVN_F_SNOW_FRAC], Po = F[RVN_F_PRECIP];
      if (adj == RVN_ADJ_MULTIPLICATIVE) F[RVN_F_PRECIP] += Po * (sf) * (epsilon - 1.0);
      else if (adj == RVN_ADJ_ADDITIVE) F[RVN_F_PRECIP] += epsilon;
      else if (adj == RVN_ADJ_REPLACE) F[RVN_F_PRECIP] += epsilon - (sf)*Po;
      if (0.0 > F[RVN_F_PRECIP]) F[RVN_F_PRECIP] = 0.0;
      if (F[RVN_F_PRECIP] == 0) F[RVN_F_SNOW_FRAC] = 0;
      else F[RVN_F_SNOW_FRAC] = 1.0 - (1.0 - F[RVN_F_SNOW_FRAC]) * Po / F[RVN_F_PRECIP];
    } else if (ftype == RVN_PF_TEMP_AVE) {
      if (adj == RVN_ADJ_MULTIPLICATIVE) F[RVN_F_TEMP_AVE] *= epsilon;
      else if (adj == RVN_ADJ_ADDITIVE) F[RVN_F_TEMP_AVE] += epsilon;
      else if (adj == RVN_ADJ_REPLACE) F[RVN_F_TEMP_AVE] = epsilon;
    }
    /* AdjustDailyHRUForcings (start of day) */
    if (ftype == RVN_PF_TEMP_AVE) {
      if (adj == RVN_ADJ_MULTIPLICATIVE) {
        double adjust = 0;
        for (int n = 0; n < nStepsPerDay; n++) adjust += eps[n] / nStepsPerDay;
        F[RVN_F_TEMP_DAILY_MIN] *= adjust; F[RVN_F_TEMP_DAILY_MAX] *= adjust; F[RVN_F_TEMP_DAILY_AVE] *= adjust;
      } else if (adj == RVN_ADJ_ADDITIVE) {
        double adjust = 0;
        for (int n = 0; n < nStepsPerDay; n++) adjust += eps[n] / nStepsPerDay;
        F[RVN_F_TEMP_DAILY_AVE] += adjust; F[RVN_F_TEMP_DAILY_MIN] += adjust; F[RVN_F_TEMP_DAILY_MAX] += adjust;
      }
    }
  }
}

/* GetSaturatedVaporPressure [kPa], src/CommonFunctions.cpp:935-949; GetDewPointTemp [C], :1131-1139 */
static double o_sat_vap(double T) { return (T >= 0) ? 0.61078 * exp(17.26939 * T / (T + 237.3)) : 0.61078 * exp(21.87456 * T / (T + 265.5)); }
static double o_dew_point_vp(double e) { double numer = log(e) + 0.4926, denom = 0.0708 - 0.00421 * log(e); return numer / denom; } /* GetDewPointTemp(e), src/CommonFunctions.cpp:1148-1156 */
static double o_dew_point(double Ta, double rel_hum) { double tmp = (17.27 * Ta / (237.7 + Ta)) + log(rel_hum); return 237.7 * tmp / (17.27 - tmp); }
/* daily[7]: the HRU's daily forcing items of the previous step (temp_daily_min/max/ave, temp_month_max/min/ave, PET_month_ave), i.e. the
 * part of CHydroUnit::_Forcings that CopyDailyForcingItems (src/Forcings.cpp:72-89) carries through the day on sub-daily steps */
static void update_forcings(const rvn_model_desc *d, const double *ptab, int member, int k, int step, const double *phi_old, const int *iSV, double *F, double *daily) {
  const rvn_time *tt = &d->times[step];
  const int nG = d->nGauges, nS = d->nSteps, mo = tt->month;
  octx c; c.d = d; c.ptab = ptab; c.k = k; c.F = F; c.phi_old = phi_old; memcpy(c.iSV, iSV, sizeof(c.iSV));
  const double elev = d->hru_elev[k];
  double ref_elev_temp = 0.0, ref_elev_precip = 0.0;
  double temp_month_min = 0, temp_month_max = 0, temp_month_ave = 0, PET_month_ave = 0;
  for (int f = 0; f < RVN_F_COUNT; f++) F[f] = 0.0;
  if (!d->hru_enabled[k]) return;
#define GS(f, g) d->gauge_series[((size_t)(f) * nG + (g)) * nS + step]
#define GC(g, f) d->gauge_const[(size_t)(g) * RVN_GC_COUNT + (f)]
  const int w0 = d->wt_ptr[k], w1 = d->wt_ptr[k + 1]; /* CSR row of this HRU: the reference's dense loops, zero weights omitted */
  for (int i = w0; i < w1; i++) { /* :184-235 */
    const int g = d->wt_idx[i];
    double wt = d->wt_val[3 * (size_t)i + 0];
    if (wt != 0.0) { F[RVN_F_PRECIP] += wt * GS(RVN_GF_PRECIP, g); F[RVN_F_SNOW_FRAC] += wt * GS(RVN_GF_SNOW_FRAC, g); ref_elev_precip += wt * GC(g, RVN_GC_ELEVATION); }
    wt = d->wt_val[3 * (size_t)i + 1];
    if (wt != 0.0) {
      F[RVN_F_TEMP_AVE] += wt * GS(RVN_GF_TEMP_AVE, g); F[RVN_F_TEMP_DAILY_AVE] += wt * GS(RVN_GF_TEMP_DAILY_AVE, g);
      F[RVN_F_TEMP_DAILY_MIN] += wt * GS(RVN_GF_TEMP_DAILY_MIN, g); F[RVN_F_TEMP_DAILY_MAX] += wt * GS(RVN_GF_TEMP_DAILY_MAX, g);
      temp_month_min += wt * GC(g, RVN_GC_MONTHLY_MIN_TEMP0 + mo - 1); temp_month_max += wt * GC(g, RVN_GC_MONTHLY_MAX_TEMP0 + mo - 1);
      temp_month_ave += wt * GC(g, RVN_GC_MONTHLY_AVE_TEMP0 + mo - 1);
      ref_elev_temp += wt * GC(g, RVN_GC_ELEVATION);
    }
    wt = d->wt_val[3 * (size_t)i + 2];
    if (wt != 0.0) {
      PET_month_ave += wt * GC(g, RVN_GC_MONTHLY_AVE_PET0 + mo - 1);
      F[RVN_F_POTENTIAL_MELT] += wt * GS(RVN_GF_POTENTIAL_MELT, g);
      F[RVN_F_PET] += wt * GS(RVN_GF_PET, g); F[RVN_F_OW_PET] += wt * GS(RVN_GF_OW_PET, g);
    }
  }
  const int p = d->hru_sb[k];
  { /* temperature gauge corrections :441-455 -- sums are recomputed over ALL gauges */
    double tc = d->sb_temp_corr[p];
    F[RVN_F_TEMP_AVE] = F[RVN_F_TEMP_DAILY_AVE] = F[RVN_F_TEMP_DAILY_MAX] = F[RVN_F_TEMP_DAILY_MIN] = 0.0;
    for (int i = w0; i < w1; i++) {
      const int g = d->wt_idx[i];
      double gauge_corr = tc + GC(g, RVN_GC_TEMP_CORR), wt = d->wt_val[3 * (size_t)i + 1];
      F[RVN_F_TEMP_AVE] += wt * (gauge_corr + GS(RVN_GF_TEMP_AVE, g));
      F[RVN_F_TEMP_DAILY_AVE] += wt * (gauge_corr + GS(RVN_GF_TEMP_DAILY_AVE, g));
      F[RVN_F_TEMP_DAILY_MAX] += wt * (gauge_corr + GS(RVN_GF_TEMP_DAILY_MAX, g));
      F[RVN_F_TEMP_DAILY_MIN] += wt * (gauge_corr + GS(RVN_GF_TEMP_DAILY_MIN, g));
    }
  }
  const double temp_ave_unc = F[RVN_F_TEMP_DAILY_AVE]; /* :468 */
  /* CModel::CorrectTemp, src/OrographicCorrections.cpp:24-58 */
  if (d->opt.orocorr_temp == RVN_OROCORR_SIMPLELAPSE || d->opt.orocorr_temp == RVN_OROCORR_HBV) {
    double lapse = P_G(&c, RVN_G_ADIABATIC_LAPSE);
    lapse /= 1000.0;
    F[RVN_F_TEMP_AVE] -= lapse * (elev - ref_elev_temp);
    if (tt->day_changed) {
      F[RVN_F_TEMP_DAILY_AVE] -= lapse * (elev - ref_elev_temp);
      F[RVN_F_TEMP_DAILY_MIN] -= lapse * (elev - ref_elev_temp);
      F[RVN_F_TEMP_DAILY_MAX] -= lapse * (elev - ref_elev_temp);
      temp_month_max -= lapse * (elev - ref_elev_temp);
      temp_month_min -= lapse * (elev - ref_elev_temp);
      if (d->opt.orocorr_temp != RVN_OROCORR_HBV) temp_month_ave -= lapse * (elev - ref_elev_temp);
    }
  }
  if (d->nPert > 0) apply_forcing_perturbation(d, RVN_PF_TEMP_AVE, member, k, step, F); /* src/UpdateForcings.cpp:474 */
  if (!tt->day_changed) { /* src/UpdateForcings.cpp:480-482: CopyDailyForcings */
    F[RVN_F_TEMP_DAILY_MIN] = daily[0]; F[RVN_F_TEMP_DAILY_MAX] = daily[1]; F[RVN_F_TEMP_DAILY_AVE] = daily[2];
    temp_month_max = daily[3]; temp_month_min = daily[4]; temp_month_ave = daily[5];
    PET_month_ave = daily[6];
  }
  daily[0] = F[RVN_F_TEMP_DAILY_MIN]; daily[1] = F[RVN_F_TEMP_DAILY_MAX]; daily[2] = F[RVN_F_TEMP_DAILY_AVE];
  daily[3] = temp_month_max; daily[4] = temp_month_min; daily[5] = temp_month_ave; daily[6] = PET_month_ave;
  F[RVN_F_TEMP_MONTH_AVE] = temp_month_ave; F[RVN_F_PET_MONTH_AVE] = PET_month_ave;
  const double subdaily_corr = d->subdaily_corr ? d->subdaily_corr[(size_t)d->et_row[step] * d->nHRU + k] : 1.0; /* CalculateSubDailyCorrection, host-tabulated */
  /* CModel::EstimateSnowFraction, src/UpdateForcings.cpp:1068-1197 */
  {
    double sf = F[RVN_F_SNOW_FRAC];
    if (d->opt.rainsnow == RVN_RAINSNOW_DINGMAN) {
      double temp = P_G(&c, RVN_G_RAINSNOW_TEMP);
      if (F[RVN_F_TEMP_DAILY_MAX] <= temp) sf = 1.0;
      else if (F[RVN_F_TEMP_DAILY_MIN] >= temp) sf = 0.0;
      else sf = (temp - F[RVN_F_TEMP_DAILY_MIN]) / (F[RVN_F_TEMP_DAILY_MAX] - F[RVN_F_TEMP_DAILY_MIN]);
    } else if (d->opt.rainsnow == RVN_RAINSNOW_THRESHOLD) {
      sf = (F[RVN_F_TEMP_AVE] <= P_G(&c, RVN_G_RAINSNOW_TEMP)) ? 1.0 : 0.0;
    } else if (d->opt.rainsnow == RVN_RAINSNOW_HBV || d->opt.rainsnow == RVN_RAINSNOW_UBCWM) {
      double frac, delta = P_G(&c, RVN_G_RAINSNOW_DELTA), temp = P_G(&c, RVN_G_RAINSNOW_TEMP);
      if (F[RVN_F_TEMP_DAILY_AVE] <= (temp - 0.5 * delta)) frac = 1.0;
      else if (F[RVN_F_TEMP_DAILY_AVE] >= (temp + 0.5 * delta)) frac = 0.0;
      else frac = 0.5 + (temp - F[RVN_F_TEMP_DAILY_AVE]) / delta;
      sf = (d->opt.rainsnow == RVN_RAINSNOW_UBCWM) ? frac : frac * (1.0 - F[RVN_F_SNOW_FRAC]) + F[RVN_F_SNOW_FRAC];
    }
    F[RVN_F_SNOW_FRAC] = sf;
  }
  double p5 = 0.0; /* precip_5day (src/UpdateForcings.cpp:107,525) */
  { /* precip gauge corrections :507-527 */
    double rc = d->sb_rain_corr[p], sc = d->sb_snow_corr[p];
    F[RVN_F_PRECIP] = 0.0;
    for (int i = w0; i < w1; i++) {
      const int g = d->wt_idx[i];
      double gauge_corr = F[RVN_F_SNOW_FRAC] * sc * GC(g, RVN_GC_SNOW_CORR) + (1.0 - F[RVN_F_SNOW_FRAC]) * rc * GC(g, RVN_GC_RAIN_CORR);
      double wt = d->wt_val[3 * (size_t)i + 0];
      F[RVN_F_PRECIP] += wt * gauge_corr * GS(RVN_GF_PRECIP, g);
      if (d->gauge_precip5) p5 += wt * gauge_corr * d->gauge_precip5[(size_t)g * d->nSteps + step];
    }
  }
  /* CModel::CorrectPrecip, src/OrographicCorrections.cpp:213-245 */
  if (d->opt.orocorr_precip == RVN_OROCORR_SIMPLELAPSE) {
    double lapse = P_G(&c, RVN_G_PRECIP_LAPSE);
    lapse /= 1000;
    if (F[RVN_F_PRECIP] > O_REAL_SMALL) { F[RVN_F_PRECIP] = omax(F[RVN_F_PRECIP] + lapse * (elev - ref_elev_precip), 0.0); p5 = omax(p5 + lapse * (elev - ref_elev_precip), 0.0); }
  } else if (d->opt.orocorr_precip == RVN_OROCORR_HBV) {
    double corr_upper = P_G(&c, RVN_G_HBVEC_LAPSE_UPPER) / 1000.0, corr = P_G(&c, RVN_G_HBVEC_LAPSE_RATE) / 1000.0;
    double lapse_elev = P_G(&c, RVN_G_HBVEC_LAPSE_ELEV), add = 0.0;
    if (elev > lapse_elev) add = (corr_upper - corr) * (elev - lapse_elev);
    F[RVN_F_PRECIP] *= omax(1.0 + corr * (elev - ref_elev_precip) + add, 0.0);
    p5 *= omax(1.0 + corr * (elev - ref_elev_precip) + add, 0.0);
  }
  if (d->nPert > 0) { /* src/UpdateForcings.cpp:558-560 */
    apply_forcing_perturbation(d, RVN_PF_PRECIP, member, k, step, F);
    apply_forcing_perturbation(d, RVN_PF_RAINFALL, member, k, step, F);
    apply_forcing_perturbation(d, RVN_PF_SNOWFALL, member, k, step, F);
  }
  if (d->hru_type[k] == RVN_HRU_MASKED_GLACIER) { F[RVN_F_PRECIP] = 0; p5 = 0; }
  if (g_precip5) g_precip5[k] = p5;
  /* air pressure / relative humidity (EstimateAirPressure, EstimateRelativeHumidity, src/UpdateForcings.cpp:677-731) and incoming
   * shortwave (CRadiation::EstimateShortwaveRadiation / ClearSkySolarRadiation / SWCloudCoverCorrection, src/Radiation.cpp:24-57,
   * 992-1038, 333-364); cloud cover is CLOUDCOV_NONE (0).  Date/geometry pieces come from the host tables. */
  double ET_radia = 0.0;
  if (d->opt.need_atmos) {
    const size_t row = (size_t)d->et_row[step] * d->nHRU + k;
    double relhum = 0.5, cc = 1.0;
    if (d->opt.air_pressure == RVN_AIRPRESS_BASIC) F[RVN_F_AIR_PRES] = 101.3 * pow(1.0 - 0.0065 * elev / (273.16 + F[RVN_F_TEMP_AVE]), 5.26);
    else if (d->opt.air_pressure == RVN_AIRPRESS_UBC) F[RVN_F_AIR_PRES] = 101.325 * (1.0 - 0.0001 * elev);
    else F[RVN_F_AIR_PRES] = 101.3;
    if (d->opt.rel_humidity == RVN_RELHUM_MINDEWPT) relhum = omin(o_sat_vap(F[RVN_F_TEMP_DAILY_MIN]) / o_sat_vap(F[RVN_F_TEMP_AVE]), 1.0);
    F[RVN_F_REL_HUMIDITY] = omin(relhum * P_LU(&c, RVN_LU_RELHUM_CORR), 1.0);
    ET_radia = d->et_radia[row];
    if (d->opt.sw_radiation == RVN_SW_RAD_DEFAULT) {
      double dew_pt = o_dew_point(F[RVN_F_TEMP_AVE], F[RVN_F_REL_HUMIDITY]);
      double Mopt = d->opt_air_mass[row], Ketp = ET_radia, Ket = d->et_flat[row];
      double wv = 1.12 * exp(0.0614 * dew_pt);                                                         /* Dingman E-10 */
      double tau = exp((-0.124 - 0.0207 * wv) + (-0.0682 - 0.0248 * wv) * Mopt) - 0.025;               /* E-9, E-11..13 */
      double tau2 = 0.5 * (1.0 - exp((-0.0363 - 0.0084 * wv) + (-0.0572 - 0.0173 * wv) * Mopt) + 0.025); /* E-15..18 */
      F[RVN_F_SW_RADIA] = tau * Ketp + tau2 * Ket + tau2 * 0.3 * (tau2 + tau) * Ketp;
    }
    if (d->opt.sw_cloudcorr == RVN_SW_CLOUD_CORR_DINGMAN) cc = 0.355 + 0.68 * (1 - 0.0);
    else if (d->opt.sw_cloudcorr == RVN_SW_CLOUD_CORR_ANNANDALE)
      cc = (F[RVN_F_TEMP_DAILY_MAX] > F[RVN_F_TEMP_DAILY_MIN]) ? 0.16 * (1.0 + 2.7E-5 * elev) * sqrt(F[RVN_F_TEMP_DAILY_MAX] - F[RVN_F_TEMP_DAILY_MIN]) : 0.0;
    F[RVN_F_SW_RADIA] *= cc;
    F[RVN_F_WIND_VEL] = 2.0; /* CModel::EstimateWindVelocity, src/UpdateForcings.cpp:739-777 */
    if (d->opt.wind_velocity == RVN_WINDVEL_UBC_MOD) {
      double wt = omin(0.04 * (F[RVN_F_TEMP_DAILY_MAX] - F[RVN_F_TEMP_DAILY_MIN]), 1.0);
      F[RVN_F_WIND_VEL] = omax(0.0, (1 - wt) * P_LU(&c, RVN_LU_MIN_WIND_SPEED) + (wt) * P_LU(&c, RVN_LU_MAX_WIND_SPEED));
    } else if (d->opt.wind_velocity == RVN_WINDVEL_SQRT)
      F[RVN_F_WIND_VEL] = omax(0.0, P_G(&c, RVN_G_WINDVEL_SCALE) * (F[RVN_F_TEMP_DAILY_MAX] - F[RVN_F_TEMP_DAILY_MIN]) + P_G(&c, RVN_G_WINDVEL_ICEPT));
  }
  /* net radiation (src/UpdateForcings.cpp:591-606): sub-canopy shortwave (CRadiation::SWCanopyCorrection, src/Radiation.cpp:377-415), total albedo
   * above and below the canopy (CHydroUnit::GetTotalAlbedo, src/HydroUnits.cpp:728-781; snow albedo = DEFAULT_SNOW_ALBEDO), net longwave
   * LW_RAD_DEFAULT with LW_INC_DEFAULT (src/Radiation.cpp:210-231), cloud cover 0.  Canopy variables of this step as
   * CVegetationClass::RecalculateCanopyParams / RecalculateRootParams (src/VegetationParams.cpp:85-142,236-262) derive them. */
  double SW_subcan_net = 0.0, veg_height = 0.0, veg_LAI = 0.0, veg_SAI = 0.0;
  if (d->opt.need_radiation) {
    const int veg = d->hru_veg[k];
    const double rel_ht = d->veg_season[((size_t)step * d->nVeg + veg) * 2 + 0], rel_LAI = d->veg_season[((size_t)step * d->nVeg + veg) * 2 + 1];
    const double max_height = P_V(&c, RVN_V_MAX_HEIGHT), sparseness = P_LU(&c, RVN_LU_FOREST_SPARSENESS), max_LAI = P_V(&c, RVN_V_MAX_LAI);
    const double SAI_ht_ratio = P_V(&c, RVN_V_SAI_HT_RATIO), extinction = P_V(&c, RVN_V_SVF_EXTINCTION), Fc = P_LU(&c, RVN_LU_FOREST_COVERAGE);
    veg_height = max_height * rel_ht;
    veg_LAI = (1.0 - sparseness) * max_LAI * rel_LAI;
    veg_SAI = (1.0 - sparseness) * (veg_height * SAI_ht_ratio);
    double svf = exp(-extinction * (veg_LAI + veg_SAI));
    double canopy_corr = 1.0;
    if (d->opt.sw_canopycorr == 1) canopy_corr = (1.0 - Fc) * 1.0 + Fc * exp(-extinction * (veg_LAI + veg_SAI));
    const double SW_subcan = F[RVN_F_SW_RADIA] * canopy_corr;
    /* GetTotalAlbedo */
    double land_albedo = 0.0;
    const double snow_albedo = 0.8, snow_cov = snow_cover(&c);
    const int type = d->hru_type[k];
    if (type == RVN_HRU_STANDARD) {
      const int iTop = sv_index(d, RVN_SV_SOIL, 0);
      double soil_sat = omax(omin(phi_old[iTop] / soil_capacity(&c, 0), 1.0), 0.0);
      land_albedo = (soil_sat)*P_S(&c, 0, RVN_S_ALBEDO_WET) + (1.0 - soil_sat) * P_S(&c, 0, RVN_S_ALBEDO_DRY);
    } else if (type == RVN_HRU_GLACIER || type == RVN_HRU_MASKED_GLACIER) land_albedo = 0.4;
    else if (type == RVN_HRU_LAKE) land_albedo = 0.1;
    else if (type == RVN_HRU_WATER) land_albedo = 0.1;
    else if (type == RVN_HRU_ROCK) land_albedo = 0.35;
    else if (type == RVN_HRU_WETLAND) land_albedo = 0.15;
    land_albedo = (snow_cov)*snow_albedo + (1.0 - snow_cov) * land_albedo;
    double veg_albedo = P_V(&c, RVN_V_ALBEDO), svf2 = svf, land2 = land_albedo;
    if (veg_albedo < 0) veg_albedo = 0.14;
    if (svf2 > 1.0) svf2 = 0.0;
    if (land2 < 0) land2 = 0.3;
    const double albedo_above = (1.0 - svf2) * (Fc)*veg_albedo + ((svf2) * (Fc) + (1.0 - Fc)) * land2;
    F[RVN_F_SW_RADIA_NET] = F[RVN_F_SW_RADIA] * (1 - albedo_above);
    SW_subcan_net = SW_subcan * (1 - land_albedo);
    if (d->opt.lw_radiation == 1) F[RVN_F_LW_RADIA_NET] = 0.0;
    else {
      const double emissivity = 0.99, Tair = F[RVN_F_TEMP_AVE] + 273.16, Tsurf = F[RVN_F_TEMP_AVE] + 273.16;
      double ea = F[RVN_F_REL_HUMIDITY] * o_sat_vap(F[RVN_F_TEMP_AVE]);
      double emiss_eff = 1.72 * pow(ea / Tair, 1 / 7.0);
      double eps_at = emiss_eff * (1.0 + 0.22 * 0.0 * 0.0);
      eps_at = (1 - Fc) * eps_at + (Fc)*1.0;
      double LW_incoming = 4.90e-9 * eps_at * pow(Tair, 4);
      F[RVN_F_LW_RADIA_NET] = LW_incoming - 4.90e-9 * emissivity * pow(Tsurf, 4);
    }
  }
  /* CModel::EstimatePotentialMelt, src/PotentialMelt.cpp:20-246 */
  {
    double pm = F[RVN_F_POTENTIAL_MELT];
    if (d->opt.pot_melt == RVN_POTMELT_DEGREE_DAY) {
      pm = P_LU(&c, RVN_LU_MELT_FACTOR) * (F[RVN_F_TEMP_DAILY_AVE] - P_LU(&c, RVN_LU_DD_MELT_TEMP)) * subdaily_corr;
    } else if (d->opt.pot_melt == RVN_POTMELT_HBV || d->opt.pot_melt == RVN_POTMELT_HBV_ROS) {
      double Ma = P_LU(&c, RVN_LU_MELT_FACTOR), Ma_min = P_LU(&c, RVN_LU_MIN_MELT_FACTOR), AM = P_LU(&c, RVN_LU_HBV_MELT_ASP_CORR);
      double MRF = P_LU(&c, RVN_LU_HBV_MELT_FOR_CORR), Fc = P_LU(&c, RVN_LU_FOREST_COVERAGE), melt_temp = P_LU(&c, RVN_LU_DD_MELT_TEMP);
      double adj = 0, slope_corr;
      if (d->hru_lat_rad[k] < 0.0) adj = O_PI;
      Ma = Ma_min + (Ma - Ma_min) * 0.5 * (1.0 - cos(tt->day_angle - O_WINTER_SOLSTICE_ANG - adj));
      Ma *= ((1.0 - Fc) * 1.0 + (Fc)*MRF);
      slope_corr = ((1.0 - Fc) * 1.0 + (Fc)*sin(d->hru_slope[k]));
      Ma *= omax(1.0 - AM * slope_corr * cos(d->hru_aspect[k] - adj), 0.0);
      double rain_energy = 0.0;
      if (d->opt.pot_melt == RVN_POTMELT_HBV_ROS) /* :77-81 */
        rain_energy = P_LU(&c, RVN_LU_RAIN_MELT_MULT) * 4.186e-3 / 0.334 * omax(F[RVN_F_TEMP_AVE], 0.0) * (F[RVN_F_PRECIP] * (1.0 - F[RVN_F_SNOW_FRAC]));
      pm = Ma * (F[RVN_F_TEMP_DAILY_AVE] - melt_temp) * subdaily_corr + rain_energy;
    } else if (d->opt.pot_melt == RVN_POTMELT_DD_FREEZE) { /* :40-49 */
      double Ma = P_LU(&c, RVN_LU_MELT_FACTOR), melt_temp = P_LU(&c, RVN_LU_DD_MELT_TEMP), Mf = P_LU(&c, RVN_LU_REFREEZE_FACTOR);
      if ((Mf > 0) && (F[RVN_F_TEMP_DAILY_AVE] - melt_temp) < 0.0) pm = Mf * (F[RVN_F_TEMP_DAILY_AVE] - melt_temp);
      else pm = Ma * (F[RVN_F_TEMP_DAILY_AVE] - melt_temp) * subdaily_corr;
    } else if (d->opt.pot_melt == RVN_POTMELT_HMETS) {
      double Ma_min = P_LU(&c, RVN_LU_MIN_MELT_FACTOR), Ma_max = P_LU(&c, RVN_LU_MAX_MELT_FACTOR);
      double melt_temp = P_LU(&c, RVN_LU_DD_MELT_TEMP), Kcum = P_LU(&c, RVN_LU_DD_AGGRADATION);
      double cum_melt = 0.0, Ma;
      if (iSV[RVN_SV_CUM_SNOWMELT] >= 0) cum_melt = phi_old[iSV[RVN_SV_CUM_SNOWMELT]];
      Ma = omin(Ma_max, Ma_min * (1 + Kcum * cum_melt));
      pm = Ma * (F[RVN_F_TEMP_DAILY_AVE] - melt_temp) * subdaily_corr;
    } else if (d->opt.pot_melt == RVN_POTMELT_EB) { /* :103-124; GetSensibleHeatSnow / GetLatentHeatSnow / GetRainHeatInput, src/SnowParams.cpp:111-190 */
      const double rainfall = F[RVN_F_PRECIP] * (1 - F[RVN_F_SNOW_FRAC]) + 0.0;
      const double air_temp = F[RVN_F_TEMP_AVE], air_pres = F[RVN_F_AIR_PRES], rel_humid = F[RVN_F_REL_HUMIDITY], wind_vel = F[RVN_F_WIND_VEL];
      const double ref_ht = 2.0, roughness = P_LU(&c, RVN_LU_ROUGHNESS), surf_temp = snow_temperature(&c);
      const double K = SW_subcan_net, L = F[RVN_F_LW_RADIA_NET];
      double temp_var = pow(0.42 / log(ref_ht / roughness), 2);
      const double H = 1.2466 * 1.012e-3 * temp_var * wind_vel * O_SEC_PER_DAY * (air_temp - surf_temp);
      double vap_pres = o_sat_vap(air_temp) * rel_humid, surf_pres = o_sat_vap(surf_temp) * 1.0;
      double CE = pow(0.42, 2) / pow(log(ref_ht / roughness), 2);
      temp_var = 0.622 * 1.2466 * CE / air_pres;
      const double LH = 2.845 * temp_var * wind_vel * O_SEC_PER_DAY * (vap_pres - surf_pres);
      double R = 0, rain_temp;
      rain_temp = o_dew_point_vp(o_sat_vap(air_temp) * rel_humid);
      if (surf_temp < O_FREEZING_TEMP) R = O_DENSITY_WATER * (rainfall / O_MM_PER_METER) * (4.186e-3 * (rain_temp - O_FREEZING_TEMP) + 0.334);
      if (surf_temp >= O_FREEZING_TEMP) R = O_DENSITY_WATER * (rainfall / O_MM_PER_METER) * (4.186e-3 * (rain_temp - O_FREEZING_TEMP));
      double S = K + L + H + LH + R;
      S *= (O_MM_PER_METER / O_DENSITY_WATER / 0.334);
      pm = S;
    } else if (d->opt.pot_melt == RVN_POTMELT_RESTRICTED) { /* :127-136 */
      const double melt_temp = P_LU(&c, RVN_LU_DD_MELT_TEMP), Ma = P_LU(&c, RVN_LU_MELT_FACTOR);
      const double tmp_rate = omax(Ma * (F[RVN_F_TEMP_DAILY_AVE] - melt_temp), 0.0);
      const double rad = (SW_subcan_net + F[RVN_F_LW_RADIA_NET]);
      const double convert = O_MM_PER_METER / O_DENSITY_WATER / 0.334;
      pm = tmp_rate * subdaily_corr + omax(rad * convert, 0.0);
    } else if (d->opt.pot_melt == RVN_POTMELT_NONE) {
      pm = 0.0;
    }
    F[RVN_F_POTENTIAL_MELT] = pm;
  }
  /* CModel::EstimatePET, src/Evaporation.cpp:282-540 (PET_DATA / PET_FROMMONTHLY / PET_CONSTANT / PET_NONE) */
  for (int ow = 0; ow < 2; ow++) {
    int method = ow ? d->opt.ow_evaporation : d->opt.evaporation;
    double PET = ow ? F[RVN_F_OW_PET] : F[RVN_F_PET];
    if (method == RVN_PET_CONSTANT) PET = 3.0;
    else if (method == RVN_PET_NONE) PET = 0.0;
    else if (method == RVN_PET_HARGREAVES_1985) { /* Hargreaves1985Evap, src/Evaporation.cpp:215-228; ET radiation hoisted by the host layer */
      if (d->et_radia) {
        double Ra = d->et_radia[(size_t)d->et_row[step] * d->nHRU + k], delT;
        Ra /= (2.501 * O_DENSITY_WATER / O_MM_PER_METER);
        delT = F[RVN_F_TEMP_DAILY_MAX] - F[RVN_F_TEMP_DAILY_MIN];
        delT = omax(delT, 0.0);
        PET = omax(0.0, 0.0023 * Ra * sqrt(delT) * (F[RVN_F_TEMP_DAILY_AVE] + 17.8));
      }
    }
    else if (method == RVN_PET_OUDIN) { /* src/Evaporation.cpp:485-489 */
      if (d->et_radia) {
        double ET_radia = d->et_radia[(size_t)d->et_row[step] * d->nHRU + k];
        PET = omax(ET_radia / O_DENSITY_WATER / 2.501 * O_MM_PER_METER * (F[RVN_F_TEMP_DAILY_AVE] + 5.0) / 100.0, 0.0);
      }
    }
    else if (method == RVN_PET_HAMON) { /* Hamon1961Evap, src/Evaporation.cpp:263-271; GetSaturatedVaporPressure, src/CommonFunctions.cpp:935-949 */
      if (d->day_length) {
        double T = F[RVN_F_TEMP_DAILY_AVE], day_length = d->day_length[(size_t)d->et_row[step] * d->nHRU + k];
        double sat_vap = (T >= 0) ? 0.61078 * exp(17.26939 * T / (T + 237.3)) : 0.61078 * exp(21.87456 * T / (T + 265.5));
        double abs_hum = 216.7 * (sat_vap * 10) / (T + 273.16);
        PET = omax(0.0, 0.0055 * 4.0 * abs_hum * day_length * day_length * 25.4);
      }
    }
    else if (method == RVN_PET_LINEAR_TEMP) { /* src/Evaporation.cpp:306-311 */
      PET = P_LU(&c, RVN_LU_PET_LIN_COEFF) * omax(F[RVN_F_TEMP_DAILY_AVE], 0.0);
    }
    else if (method == RVN_PET_MOHYSE) { /* src/Evaporation.cpp:476-483 */
      if (d->sunset_angle) {
        double cpet = P_G(&c, RVN_G_MOHYSE_PET_COEFF), ws = d->sunset_angle[(size_t)d->et_row[step] * d->nHRU + k];
        PET = cpet / 3.1415926535898 * ws * exp((17.3 * F[RVN_F_TEMP_AVE]) / (238 + F[RVN_F_TEMP_AVE]));
      }
    }
    else if (method == RVN_PET_TURC_1961) { /* TurcEvap, src/Evaporation.cpp:61-68 */
      double t = omax(0.0, F[RVN_F_TEMP_DAILY_AVE]);
      double evp = 0.013 * t / (t + 15) * ((F[RVN_F_SW_RADIA]) * 23.8846 + 50);
      if (F[RVN_F_REL_HUMIDITY] < 0.5) evp *= (1 + (50 - F[RVN_F_REL_HUMIDITY] * 100) / 70);
      PET = omax(0.0, evp);
    }
    else if (method == RVN_PET_MAKKINK_1957) { /* Makkink1957Evap, src/Evaporation.cpp:32-48; GetSatVapSlope / GetLatentHeatVaporization / GetPsychrometricConstant, src/CommonFunctions.cpp:959-998 */
      double T = F[RVN_F_TEMP_AVE], sat_vap = o_sat_vap(T);
      double de_dT = (T > 0) ? 17.26939 * 237.3 / pow(T + 237.3, 2) * sat_vap : 21.87456 * 265.5 / pow(T + 265.5, 2) * sat_vap;
      double LH_vapor = 2.495 - 0.002361 * T;
      double gamma = 1.012e-3 / 0.622 * F[RVN_F_AIR_PRES] / LH_vapor;
      PET = omax(0.0, 0.61 * (de_dT / (de_dT + gamma)) * F[RVN_F_SW_RADIA] * 23.8846 / 58.5 - 0.12);
    }
    else if (method == RVN_PET_PENMAN_SIMPLE33) /* Valiantzas (2006) eqn 33, src/Evaporation.cpp:454-463 */
      PET = 0.047 * F[RVN_F_SW_RADIA] * sqrt(F[RVN_F_TEMP_AVE] + 9.5) - 2.4 * pow(F[RVN_F_SW_RADIA] / ET_radia, 2.0) + 0.09 * (F[RVN_F_TEMP_AVE] + 20) * (1 - F[RVN_F_REL_HUMIDITY]);
    else if (method == RVN_PET_PENMAN_SIMPLE39) /* eqn 39, :465-474 */
      PET = 0.038 * F[RVN_F_SW_RADIA] * sqrt(F[RVN_F_TEMP_AVE] + 9.5) - 2.4 * pow(F[RVN_F_SW_RADIA] / ET_radia, 2.0) + 0.075 * (F[RVN_F_TEMP_AVE] + 20) * (1 - F[RVN_F_REL_HUMIDITY]);
    else if (method == RVN_PET_VAPDEFICIT) { /* src/Evaporation.cpp:506-518 */
      double e_sat = o_sat_vap(F[RVN_F_TEMP_AVE]), ea = F[RVN_F_REL_HUMIDITY] * e_sat;
      PET = P_LU(&c, RVN_LU_PET_VAP_COEFF) * (e_sat - ea);
    }
    else if (method == RVN_PET_LINACRE) { /* src/Evaporation.cpp:491-504 */
      double T = F[RVN_F_TEMP_DAILY_AVE], Tdew = o_dew_point(T, F[RVN_F_REL_HUMIDITY]), latit = d->hru_lat_rad[k] * 57.29578951;
      PET = ((ow ? 700 : 500) * T / (100 - latit) + 15.0 * (T - Tdew)) / (80.0 - T);
    }
    else if (method == RVN_PET_PRIESTLEY_TAYLOR) { /* PriestleyTaylorEvap, src/Evaporation.cpp:155-168 */
      double T = F[RVN_F_TEMP_AVE], sat_vap = o_sat_vap(T);
      double de_dT = (T > 0) ? 17.26939 * 237.3 / pow(T + 237.3, 2) * sat_vap : 21.87456 * 265.5 / pow(T + 265.5, 2) * sat_vap;
      double LH_vapor = 2.495 - 0.002361 * T;
      double gamma = 1.012e-3 / 0.622 * F[RVN_F_AIR_PRES] / LH_vapor;
      PET = P_LU(&c, RVN_LU_PRIESTLEYTAYLOR_COEFF) * (de_dT / (de_dT + gamma)) * omax(F[RVN_F_SW_RADIA_NET] + F[RVN_F_LW_RADIA_NET], 0.0) / LH_vapor / O_DENSITY_WATER * O_MM_PER_METER;
    }
    else if (method == RVN_PET_PENMAN_COMBINATION) { /* :404-416, PenmanCombinationEvap :119-144, GetVerticalTransportEfficiency src/CommonFunctions.cpp:1083-1094 */
      /* canopy roughness of this step (Brook90 ROUGH, src/VegetationParams.cpp:236-262) */
      const double SAI_ht_ratio = P_V(&c, RVN_V_SAI_HT_RATIO), lu_rough = P_LU(&c, RVN_LU_ROUGHNESS);
      double closed_roughness, closed_zerodisp, zero_pl, z0, ratio, xx;
      if (veg_height >= 10.0) closed_roughness = 0.05 * veg_height;
      else if (veg_height <= 1.0) closed_roughness = 0.13 * veg_height;
      else { closed_roughness = 0.13 * 1.0; closed_roughness += (0.05 * 10.0 - 0.13 * 1.0) * (veg_height - 1.0) / (10.0 - 1.0); }
      closed_zerodisp = veg_height - closed_roughness / 0.3;
      if (lu_rough > closed_roughness) closed_roughness = lu_rough;
      ratio = (veg_LAI + veg_SAI) / (4.0 + SAI_ht_ratio * veg_height);
      if (ratio >= 1.0) { zero_pl = closed_zerodisp; z0 = closed_roughness; }
      else {
        xx = 0.0;
        if (veg_height > O_REAL_SMALL) xx = ratio * pow(-1.0 + exp(0.909 - 3.03 * closed_zerodisp / veg_height), 4);
        zero_pl = 1.1 * veg_height * log(1.0 + pow(xx, 0.25));
        z0 = omin(0.3 * (veg_height - zero_pl), lu_rough + 0.3 * veg_height * sqrt(xx));
      }
      double ref_ht = veg_height + 2.0;
      double numer = 0.622 * 1.2466;
      double denom = F[RVN_F_AIR_PRES] * O_DENSITY_WATER * (1.0 / pow(0.42, 2) * (pow((log((ref_ht - zero_pl) / z0)), 2)));
      const double vert_trans = numer / denom;
      double T = F[RVN_F_TEMP_AVE], sat_vap = o_sat_vap(T);
      double de_dT = (T > 0) ? 17.26939 * 237.3 / pow(T + 237.3, 2) * sat_vap : 21.87456 * 265.5 / pow(T + 265.5, 2) * sat_vap;
      double LH_vapor = 2.495 - 0.002361 * T;
      double gamma = 1.012e-3 / 0.622 * F[RVN_F_AIR_PRES] / LH_vapor;
      double vapor_def = sat_vap * (1.0 - (F[RVN_F_REL_HUMIDITY]));
      numer = de_dT * omax((F[RVN_F_SW_RADIA_NET]) + (F[RVN_F_LW_RADIA_NET]), 0.0);
      numer += gamma * vert_trans * O_DENSITY_WATER * LH_vapor * (F[RVN_F_WIND_VEL]) * vapor_def;
      denom = O_DENSITY_WATER * LH_vapor * (de_dT + gamma);
      PET = omax(0.0, (numer / denom)) * O_MM_PER_METER;
    }
    else if (method == RVN_PET_HARGREAVES) { /* HargreavesEvap, src/Evaporation.cpp:179-203 */
      double Ra = ET_radia / (2.501 * O_DENSITY_WATER / O_MM_PER_METER), Ct = 0.125;
      double delT = (1.8 * temp_month_max + 32.0) - (1.8 * temp_month_min + 32.0);
      if (F[RVN_F_REL_HUMIDITY] >= 0.54) Ct = 0.035 * pow(100 * (1.0 - F[RVN_F_REL_HUMIDITY]), 0.333);
      PET = omax(0.0, 0.0075 * Ra * Ct * sqrt(delT) * (1.8 * F[RVN_F_TEMP_DAILY_AVE] + 32.0));
    }
    else if (method == RVN_PET_FROMMONTHLY) { /* src/Evaporation.cpp:349-355 */
      double peRatio = 1.0 + O_HBV_PET_TEMP_CORR * (temp_ave_unc - temp_month_ave);
      peRatio = omax(0.0, omin(2.0, peRatio));
      PET = PET_month_ave * peRatio;
    }
    PET = omax(PET, 0.0);                 /* src/Evaporation.cpp:534-537 */
    PET = PET * P_V(&c, RVN_V_PET_VEG_CORR);
    if (ow) F[RVN_F_OW_PET] = PET; else F[RVN_F_PET] = PET;
  }
  /* CModel::CorrectPET, src/OrographicCorrections.cpp:350-440 */
  if (d->opt.orocorr_pet == RVN_OROCORR_HBV) {
    double factor = O_GLOBAL_PET_CORR * omax(1.0 - O_HBV_PET_ELEV_CORR * (elev - ref_elev_temp), 0.0);
    F[RVN_F_PET] *= factor; F[RVN_F_OW_PET] *= factor;
  }
  if (d->opt.evaporation == RVN_PET_FROMMONTHLY || d->opt.evaporation == RVN_PET_CONSTANT || d->opt.evaporation == RVN_PET_HAMON || d->opt.evaporation == RVN_PET_LINEAR_TEMP ||
      d->opt.evaporation == RVN_PET_TURC_1961 || d->opt.evaporation == RVN_PET_LINACRE) /* IsDailyPETmethod, src/Evaporation.cpp:540-556 */
    F[RVN_F_PET] *= subdaily_corr;
  if (d->opt.snow_suppress_PET) {
    double SWE = 0.0, sc = 1.0;
    if (iSV[RVN_SV_SNOW] >= 0) SWE = phi_old[iSV[RVN_SV_SNOW]];
    if (iSV[RVN_SV_SNOW_COVER] >= 0) sc = phi_old[iSV[RVN_SV_SNOW_COVER]];
    if (SWE > 0.1) { F[RVN_F_PET] *= (1.0 - sc); F[RVN_F_OW_PET] *= (1.0 - sc); }
  }
  if (d->hru_type[k] == RVN_HRU_STANDARD) F[RVN_F_PET] *= P_S(&c, 0, RVN_S_PET_CORRECTION);
  if (d->hru_type[k] == RVN_HRU_MASKED_GLACIER) F[RVN_F_PET] = F[RVN_F_OW_PET] = 0.0;
  if (d->opt.direct_evap) { /* :640-650 */
    double rainfall = F[RVN_F_PRECIP] * (1.0 - F[RVN_F_SNOW_FRAC]);
    double reduce = omin(F[RVN_F_PET], rainfall + 0.0);
    double rainfrac = rainfall / (rainfall + 0.0);
    if ((rainfall + 0.0) == 0) rainfrac = 1.0;
    if (F[RVN_F_PRECIP] + 0.0 - reduce > 0.0) F[RVN_F_SNOW_FRAC] = 1.0 - (rainfall - reduce * rainfrac) / (F[RVN_F_PRECIP] - reduce * rainfrac);
    F[RVN_F_PRECIP] -= reduce * rainfrac;
    F[RVN_F_PET] -= reduce * rainfrac;
  }
#undef GS
#undef GC
}

/* ======================================================================== processes ============ */
/* each op restates GetRatesOfChange followed by ApplyConstraints; rates[] arrives zeroed
 * (CModel::ApplyProcess, src/Model.cpp:3066-3071) */
static void op_rates(const octx *c, const rvn_process *pr, const int *iFrom, const int *iTo, const double *sv, double *rates) {
  const rvn_model_desc *d = c->d;
  const double tstep = d->opt.timestep;
  const int type = d->hru_type[c->k];
  const double *F = c->F;
  switch (pr->op) {
    case RVN_OP_PRECIPITATION: { /* src/PartitionPrecip.cpp:135-355 (no ICE_THICKNESS) */
      enum { qPond = 0, qCan, qCanSnow, qDep, qAtmos, qSnow, qSnowLiq, qSnowDef, qNewSnow, qLake, qSW, qColdContent, qSnowDepth };
      const int iCan = c->iSV[RVN_SV_CANOPY], iSCan = c->iSV[RVN_SV_CANOPY_SNOW], iSnLiq = c->iSV[RVN_SV_SNOW_LIQ], iSWE = c->iSV[RVN_SV_SNOW];
      double total_precip = F[RVN_F_PRECIP], Fsnow = F[RVN_F_SNOW_FRAC], air_temp = F[RVN_F_TEMP_AVE];
      double snowfall, rainfall, rain_int, snow_int, rainthru, snowthru;
      if (iSWE < 0) Fsnow = 0;
      snowfall = (Fsnow)*total_precip;
      rainfall = (1.0 - Fsnow) * total_precip;
      rainfall += 0.0; /* irrigation */
      double SWE = 0.0;
      if (iSWE >= 0) SWE = sv[iSWE];
      double Fs = snow_cover(c);
      const int is_lake = (type == RVN_HRU_LAKE);
      if (!is_lake) {
        double Fcan = P_LU(c, RVN_LU_FOREST_COVERAGE);
        double rain_pct = c->rain_icept_pct;
        rain_int = rainfall * rain_pct;
        if (iCan >= 0) {
          double capacity = c->canopy_cap;
          rain_int = omin(rain_int, omax((capacity - sv[iCan]) / tstep, 0.0));
          rates[qCan] = Fcan * rain_int;
        } else rates[qAtmos] += Fcan * rain_int;
        rainthru = rainfall - Fcan * rain_int;
        double snow_pct = c->snow_icept_pct;
        snow_int = snowfall * snow_pct;
        if (iSCan >= 0) {
          double snow_capacity = c->canopy_snow_cap;
          snow_int = omax(omin(snow_int, omax((snow_capacity - sv[iSCan]) / tstep, 0.0)), 0.0);
          rates[qCanSnow] = Fcan * snow_int;
        } else rates[qAtmos] += Fcan * snow_int;
        snowthru = snowfall - Fcan * snow_int;
      } else { snowthru = snowfall; rainthru = rainfall; }
      if (is_lake) { if (hru_linked(d, c->k)) rates[qSW] = snowthru + rainthru; else rates[qLake] = snowthru + rainthru; } /* src/PartitionPrecip.cpp:258-265 */
      else if (type == RVN_HRU_WETLAND) rates[qDep] = snowthru + rainthru;
      else if (type == RVN_HRU_WATER) rates[qSW] = snowthru + rainthru;
      else {
        if (c->iSV[RVN_SV_NEW_SNOW] < 0) {
          if (iSWE >= 0) {
            rates[qSnow] = snowthru;
            if (c->iSV[RVN_SV_SNOW_DEPTH] >= 0) { /* CalcFreshSnowDensity, src/SnowParams.cpp */
              double rho = (air_temp <= O_FREEZING_TEMP) ? 67.92 + 51.25 * exp((air_temp - O_FREEZING_TEMP) / 2.59)
                                                         : omin((119.17 + 20.0 * (air_temp - O_FREEZING_TEMP)), 200.0); /* src/SnowParams.cpp:37-47 */
              rates[qSnowDepth] = snowthru * O_DENSITY_ICE / rho;
            }
            if (c->iSV[RVN_SV_COLD_CONTENT] >= 0) rates[qColdContent] = snowthru * omax(O_FREEZING_TEMP - air_temp, 0.0) * O_HCP_ICE / O_MM_PER_METER;
            if ((iSnLiq >= 0) && (c->iSV[RVN_SV_SNOW_DEFICIT] < 0) && (SWE > 0.0)) {
              double melt_cap = state_var_max(c, iSnLiq, sv);
              double fill_rate = Fs * omin(omax(melt_cap - sv[iSnLiq], 0.0) / tstep, rainthru);
              rates[qSnowLiq] = fill_rate;
              rainthru -= fill_rate;
            }
            if ((c->iSV[RVN_SV_SNOW_DEFICIT] >= 0) && (SWE > 0.0)) rates[qSnowDef] = P_G(c, RVN_G_SNOW_SWI) * snowthru;
          }
        } else rates[qNewSnow] = snowthru;
        rates[qPond] = rainthru; /* lake/wetland branches unreachable here */
      }
      break;
    }
    case RVN_OP_SNOBAL_HMETS: { /* src/SnowBalance.cpp:472-522 */
      double SWE = sv[iFrom[0]], SL = sv[iFrom[3]], cum_melt = sv[iFrom[4]];
      double Kf = P_LU(c, RVN_LU_REFREEZE_FACTOR), Tbf = P_LU(c, RVN_LU_DD_REFREEZE_TEMP), exp_fe = P_LU(c, RVN_LU_REFREEZE_EXP);
      double fcmin = P_G(c, RVN_G_SNOW_SWI_MIN), fcmax = P_G(c, RVN_G_SNOW_SWI_MAX), Ccum = P_G(c, RVN_G_SWI_REDUCT_COEFF);
      double potmelt = omax(F[RVN_F_POTENTIAL_MELT], 0.0);
      double Tdiurnal = 0.5 * (F[RVN_F_TEMP_DAILY_AVE] + F[RVN_F_TEMP_DAILY_MIN]);
      double pot_freeze, freeze, melt, SWI, deficit, to_liq, ripe;
      SL = omax(SL, 0.0);
      pot_freeze = Kf * pow(omax(Tbf - Tdiurnal, 0.0), exp_fe);
      freeze = omin(pot_freeze * tstep, SL);
      SL -= freeze; SWE += freeze;
      melt = omin(potmelt * tstep, SWE);
      SWE -= melt; cum_melt += melt;
      if (SWE <= 0.0) cum_melt = 0.0;
      SWI = omax(fcmin, fcmax * (1 - Ccum * cum_melt));
      deficit = omax(SWI * SWE - SL, 0.0);
      to_liq = omin(melt, deficit);
      melt -= to_liq; SL += to_liq;
      ripe = omax(SL - SWI * SWE, 0.0);
      SL -= ripe;
      rates[0] = to_liq / tstep; rates[1] = melt / tstep; rates[2] = ripe; rates[3] = freeze / tstep; rates[4] = (melt + to_liq) / tstep;
      if (SWE <= 0.0) rates[4] = -sv[iFrom[4]] / tstep;
      break;
    }
    case RVN_OP_INF_HMETS: case RVN_OP_INF_HBV: case RVN_OP_INF_GR4J: case RVN_OP_INF_VIC_ARNO: case RVN_OP_INF_RATIONAL:
    case RVN_OP_INF_ALL_INFILTRATES: case RVN_OP_INF_VIC: case RVN_OP_INF_PRMS: case RVN_OP_INF_PDM: case RVN_OP_INF_GREEN_AMPT: case RVN_OP_INF_GA_SIMPLE: case RVN_OP_INF_UBC: case RVN_OP_INF_SCS: case RVN_OP_INF_SCS_NOABSTRACTION: { /* src/Infiltration.cpp:304-330,384-398,477-489,491-514 + constraints :605-626 */
      if (type != RVN_HRU_STANDARD && type != RVN_HRU_ROCK && type != RVN_HRU_MASKED_GLACIER) break;
      double Fimp = P_LU(c, RVN_LU_IMPERMEABLE_FRAC);
      double ponded = omax(sv[c->iSV[RVN_SV_PONDED_WATER]], 0.0);
      double rainthru = (ponded / tstep);
      if (type == RVN_HRU_ROCK) { rates[0] = 0.0; rates[1] = rainthru; }
      else if (pr->op == RVN_OP_INF_HMETS) {
        double stor = sv[iTo[0]], max_stor = soil_capacity(c, 0), coef = P_LU(c, RVN_LU_HMETS_RUNOFF_COEFF);
        double sat = omin(omax(stor / max_stor, 0.0), 1.0);
        double direct = Fimp * rainthru;
        double runoff = coef * (sat) * (1.0 - Fimp) * rainthru;
        double infil = (1.0 - Fimp) * rainthru - runoff;
        double delayed = coef * pow(sat, 2.0) * infil;
        infil -= delayed;
        rates[0] = infil; rates[1] = direct; rates[2] = runoff; rates[3] = delayed;
      } else if (pr->op == RVN_OP_INF_HBV) {
        double stor = sv[iTo[0]], max_stor = soil_capacity(c, 0), beta = P_S(c, 0, RVN_S_HBV_BETA);
        double sat = omax(omin(stor / max_stor, 1.0), 0.0);
        double runoff = pow(sat, beta) * rainthru;
        runoff = (Fimp)*rainthru + (1 - Fimp) * runoff;
        rates[0] = rainthru - runoff; rates[1] = runoff;
      } else if (pr->op == RVN_OP_INF_VIC_ARNO) { /* src/Infiltration.cpp:400-416 */
        double stor = sv[iTo[0]], max_stor = soil_capacity(c, 0), b = P_S(c, 0, RVN_S_VIC_B_EXP);
        double sat = omin(stor / max_stor, 1.0);
        double sat_area = 1.0 - pow(1.0 - sat, b);
        double runoff = sat_area * rainthru;
        runoff = (Fimp)*rainthru + (1 - Fimp) * runoff;
        rates[0] = rainthru - runoff; rates[1] = runoff;
      } else if (pr->op == RVN_OP_INF_RATIONAL) { /* :333-338 */
        double runoff = P_LU(c, RVN_LU_PARTITION_COEFF) * rainthru;
        rates[0] = rainthru - runoff; rates[1] = runoff;
      } else if (pr->op == RVN_OP_INF_ALL_INFILTRATES) { /* :347-353 */
        double runoff = 0.0;
        runoff = (Fimp)*rainthru + (1 - Fimp) * runoff;
        rates[0] = rainthru - runoff; rates[1] = runoff;
      } else if (pr->op == RVN_OP_INF_VIC) { /* :361-383 */
        int iTop = sv_index(d, RVN_SV_SOIL, 0);
        double stor = sv[iTop], max_stor = soil_capacity(c, 0);
        double zmax = P_S(c, 0, RVN_S_VIC_ZMAX), zmin = P_S(c, 0, RVN_S_VIC_ZMIN), alpha = P_S(c, 0, RVN_S_VIC_ALPHA);
        double sat = omin(stor / max_stor, 1.0);
        double gamma = 1.0 / (alpha + 1.0);
        double K1 = pow((zmax - zmin) * alpha * gamma, -gamma);
        double Smax = gamma * (alpha * zmax + zmin);
        double runoff = rainthru * omax(1.0 - K1 * pow(Smax - sat, gamma), 1.0); /* threshMax(x,1.0,0.0) */
        runoff = (Fimp)*rainthru + (1 - Fimp) * runoff;
        rates[0] = rainthru - runoff; rates[1] = runoff;
      } else if (pr->op == RVN_OP_INF_PRMS) { /* :418-431 */
        int iTop = sv_index(d, RVN_SV_SOIL, 0);
        double sat_frac = P_LU(c, RVN_LU_MAX_SAT_AREA_FRAC), tens_stor = soil_tension_capacity(c, 0), stor = sv[iTop];
        double runoff = sat_frac * omin(stor / tens_stor, 1.0) * rainthru;
        runoff = (Fimp)*rainthru + (1 - Fimp) * runoff;
        rates[0] = rainthru - runoff; rates[1] = runoff;
      } else if (pr->op == RVN_OP_INF_PDM) { /* :540-563 */
        int iTop = sv_index(d, RVN_SV_SOIL, 0);
        double b = P_LU(c, RVN_LU_PDM_B), stor = sv[iTop], max_stor = soil_capacity(c, 0);
        double c_max = (b + 1) * max_stor;
        double c_star = c_max * (1.0 - pow(1.0 - (stor / max_stor), 1.0 / (b + 1.0)));
        double infil = max_stor * (pow(1.0 - (c_star / c_max), b + 1.0) - pow(1.0 - omin(c_star + ponded, c_max) / c_max, b + 1.0));
        infil = omin(infil, ponded);
        infil = (1.0 - Fimp) * infil / tstep;
        rates[0] = infil; rates[1] = rainthru - infil;
      } else if (pr->op == RVN_OP_INF_GREEN_AMPT || pr->op == RVN_OP_INF_GA_SIMPLE) { /* CmvInfiltration::GetGreenAmptRunoff, src/GreenAmpt.cpp:65-175 */
        double finf, cumInf = 0, Keff, alpha, deficit, runoff;
        double Ksat = P_S(c, 0, RVN_S_HYDRAUL_COND), psi_wf = P_S(c, 0, RVN_S_WETTING_FRONT_PSI), poro = P_S(c, 0, RVN_S_POROSITY);
        double stor = sv[iTo[0]], max_stor = soil_capacity(c, 0), initStor = 0;
        int simple = (pr->op == RVN_OP_INF_GA_SIMPLE);
        Keff = Ksat;
        if (simple) {
          cumInf = sv[iFrom[2]];
          initStor = sv[iFrom[3]];
          if (rainthru <= Keff) {
            if (cumInf <= 0) { cumInf = 0; initStor = 0; }
            else {
              double distribution_rate = omax(cumInf * 2 / 3, Keff);
              cumInf -= distribution_rate * tstep;
              initStor += distribution_rate * tstep;
              initStor = omin(initStor, stor);
              if (cumInf <= 0) { cumInf = 0; initStor = 0; }
            }
            rates[2] = (cumInf - sv[iFrom[2]]) / tstep;
            rates[3] = (initStor - sv[iFrom[3]]) / tstep;
          }
        }
        if (rainthru <= O_REAL_SMALL) { rates[0] = 0; rates[1] = 0; }
        else {
          deficit = (1.0 - stor / max_stor) * poro;
          alpha = psi_wf * deficit;
          if (!simple) cumInf = o_green_ampt_cum_inf(tstep, alpha, Keff, rainthru);
          else {
            if (cumInf <= 0) { cumInf = o_green_ampt_cum_inf(tstep, alpha, Keff, rainthru); initStor = stor; }
            else { deficit = (1.0 - initStor / max_stor) * poro; alpha = psi_wf * deficit; }
          }
          if (cumInf > 0) finf = Keff * (1.0 + alpha / cumInf); else finf = 0.0;
          runoff = omax(rainthru - finf, 0.0);
          runoff = (Fimp)*rainthru + (1 - Fimp) * runoff;
          if (simple) {
            cumInf += (rainthru - runoff) * tstep;
            rates[2] = (cumInf - sv[iFrom[2]]) / tstep;
            rates[3] = (initStor - sv[iFrom[3]]) / tstep;
          }
          rates[0] = rainthru - runoff; rates[1] = runoff;
        }
      } else if (pr->op == RVN_OP_INF_SCS || pr->op == RVN_OP_INF_SCS_NOABSTRACTION) { /* src/Infiltration.cpp:338-343; GetSCSRunoff :638-680 */
        int condition = 2;
        double S, CN, W, Weff, TR, Ia;
        TR = g_precip5[c->k] / 25.4;
        CN = P_LU(c, RVN_LU_SCS_CN);
        int month = d->times[c->step].month;
        if ((month > 4) && (month < 9)) { if (TR < 1.4) { condition = 1; } else if (TR > 2.1) { condition = 3; } }
        else { if (TR < 0.5) { condition = 1; } else if (TR > 1.1) { condition = 3; } }
        if (condition == 1) { CN = 5E-05 * pow(CN, 3) + 0.0008 * pow(CN, 2) + 0.4431 * CN; }
        else if (condition == 3) { CN = 7E-05 * pow(CN, 3) - 0.0185 * pow(CN, 2) + 2.1586 * CN; }
        CN = omin(CN, 100.0);
        S = 25.4 * (1000.0 / CN - 10.0);
        W = rainthru * tstep;
        if (pr->op == RVN_OP_INF_SCS) Ia = (P_LU(c, RVN_LU_SCS_IA_FRACTION)) * S; else Ia = 0.0;
        Weff = pow(omax(W - Ia, 0.0), 2) / (W + (S - Ia));
        if (W + (S - Ia) == 0.0) { Weff = 0.0; }
        double runoff = Weff / tstep;
        rates[0] = rainthru - runoff; rates[1] = runoff;
      } else if (pr->op == RVN_OP_INF_UBC) { /* CmvInfiltration::GetUBCWMRunoff, src/Infiltration.cpp:685-759 */
        const double V0FLAX = 1800.0;
        double infil, to_GW, to_interflow, runoff, flash_factor, b1, b2;
        double soil_deficit = omax(0.0, soil_capacity(c, 0) - sv[iTo[0]]);
        double max_perc_rate = P_S(c, 1, RVN_S_MAX_PERC_RATE), P0AGEN = P_S(c, 0, RVN_S_UBC_INFIL_SOIL_DEF);
        double P0DSH = P_G(c, RVN_G_UBC_GW_SPLIT), V0FLAS = P_G(c, RVN_G_UBC_FLASH_PONDING);
        b1 = 1.0;
        if (Fimp < 1.0) b1 = Fimp * pow(10.0, -soil_deficit / P0AGEN);
        flash_factor = 0.0;
        if (rainthru > V0FLAS) flash_factor = (1.0 + log(rainthru / V0FLAX) / log(V0FLAX / V0FLAS));
        if (1.0 < flash_factor) flash_factor = 1.0; /* lowerswap */
        if (0.0 > flash_factor) flash_factor = 0.0; /* upperswap */
        b2 = (b1) * (1.0) + (1.0 - b1) * flash_factor;
        infil = omin(soil_deficit / tstep, rainthru);
        to_GW = omin(max_perc_rate, rainthru - infil);
        to_interflow = (rainthru - infil - to_GW);
        infil *= 1 - b2; to_GW *= 1 - b2; to_interflow *= 1 - b2;
        runoff = rainthru - infil - to_GW - to_interflow;
        rates[0] = infil; rates[1] = runoff; rates[2] = to_interflow; rates[3] = (1.0 - P0DSH) * to_GW; rates[4] = (P0DSH)*to_GW;
      } else {
        double x1 = soil_capacity(c, 0), stor = sv[iTo[0]];
        double sat = stor / x1;
        double tmp = tanh(ponded / x1);
        double infil = x1 * (1.0 - (sat * sat)) * tmp / (1.0 + sat * tmp);
        infil = (1.0 - Fimp) * infil;
        rates[0] = infil; rates[1] = rainthru - infil;
      }
      if (type != RVN_HRU_STANDARD && type != RVN_HRU_MASKED_GLACIER) break; /* ApplyConstraints guard */
      rates[0] = omin(rates[0], sv[iFrom[0]] / tstep);
      double inf = rates[0];
      if (!d->opt.allow_soil_overfill) {
        double max_stor = state_var_max(c, iTo[0], sv);
        inf = omin(rates[0], omax(max_stor - sv[iTo[0]], 0.0) / tstep);
      }
      rates[1] += (rates[0] - inf);
      rates[0] = inf;
      break;
    }
    case RVN_OP_OVERFLOW: { /* src/Flush.cpp:259-279 */
      double max_storage = omax(state_var_max(c, iFrom[0], sv), 0.0);
      rates[0] = omax(sv[iFrom[0]] - max_storage, 0.0) / tstep;
      break;
    }
    case RVN_OP_FLUSH: rates[0] = pr->da[0] * sv[iFrom[0]] / tstep; break; /* src/Flush.cpp:75-86 */
    case RVN_OP_SPLIT: /* src/Flush.cpp:178-186 */
      rates[0] = (pr->da[0]) * sv[iFrom[0]] / tstep;
      rates[1] = (1.0 - pr->da[0]) * sv[iFrom[0]] / tstep;
      break;
    case RVN_OP_BASE_LINEAR: case RVN_OP_BASE_POWER_LAW: case RVN_OP_BASE_GR4J: case RVN_OP_BASE_LINEAR_ANALYTIC: case RVN_OP_BASE_VIC:
    case RVN_OP_BASE_TOPMODEL: case RVN_OP_BASE_LINEAR_CONSTRAIN: case RVN_OP_BASE_CONSTANT: case RVN_OP_BASE_THRESH_POWER:
    case RVN_OP_BASE_THRESH_STOR: { /* src/Baseflow.cpp:152-186,208-215,238-243 + :297-310 */
      if (type == RVN_HRU_LAKE || type == RVN_HRU_WATER || type == RVN_HRU_ROCK) { /* GetRatesOfChange returns; constraints still run */ }
      else {
        int m = pr->ia[0];
        double stor = sv[iFrom[0]], max_stor = (m >= 0) ? soil_capacity(c, m) : 0.0;
        stor = omin(omax(stor, 0.0), max_stor);
        if (pr->op == RVN_OP_BASE_LINEAR) rates[0] = P_S(c, m, RVN_S_BASEFLOW_COEFF) * stor;
        else if (pr->op == RVN_OP_BASE_POWER_LAW) rates[0] = P_S(c, m, RVN_S_BASEFLOW_COEFF) * pow(stor, P_S(c, m, RVN_S_BASEFLOW_N));
        else if (pr->op == RVN_OP_BASE_LINEAR_ANALYTIC) { /* :199-205 */
          double K = P_S(c, m, RVN_S_BASEFLOW_COEFF);
          rates[0] = stor * (1 - exp(-K * tstep)) / tstep;
        } else if (pr->op == RVN_OP_BASE_VIC) { /* :221-228 */
          double max_rate = P_S(c, m, RVN_S_MAX_BASEFLOW_RATE), n = P_S(c, m, RVN_S_BASEFLOW_N);
          rates[0] = max_rate * pow(stor / max_stor, n);
        } else if (pr->op == RVN_OP_BASE_TOPMODEL) { /* :230-242 */
          double K = P_S(c, m, RVN_S_MAX_BASEFLOW_RATE), n = P_S(c, m, RVN_S_BASEFLOW_N);
          int terr = d->hru_terrain ? d->hru_terrain[c->k] : -1;
          double lambda = c->ptab[d->terr_off + (terr >= 0 ? terr : 0) * RVN_T_COUNT + RVN_T_TOPMODEL_LAMBDA];
          rates[0] = K * (max_stor / n * pow(lambda, -n)) * pow(stor / max_stor, n);
        } else if (pr->op == RVN_OP_BASE_LINEAR_CONSTRAIN) { /* :188-197 */
          double K = P_S(c, m, RVN_S_BASEFLOW_COEFF), sat = stor / max_stor;
          if (sat > P_S(c, m, RVN_S_FIELD_CAPACITY)) rates[0] = K * stor;
        } else if (pr->op == RVN_OP_BASE_CONSTANT) { rates[0] = P_S(c, m, RVN_S_MAX_BASEFLOW_RATE);
        } else if (pr->op == RVN_OP_BASE_THRESH_POWER) { /* :250-266 */
          double K = P_S(c, m, RVN_S_MAX_BASEFLOW_RATE), n = P_S(c, m, RVN_S_BASEFLOW_N), tsat = P_S(c, m, RVN_S_BASEFLOW_THRESH);
          double sat = stor / max_stor;
          rates[0] = 0.0;
          if (sat > tsat) {
            rates[0] = K * pow((sat - tsat) / (1.0 - tsat), n);
            rates[0] = omin(rates[0], max_stor * (sat - tsat) / tstep);
          }
        } else if (pr->op == RVN_OP_BASE_THRESH_STOR) { /* :268-279 */
          double K = P_S(c, m, RVN_S_BASEFLOW_COEFF2), tstor = P_S(c, m, RVN_S_STORAGE_THRESHOLD);
          rates[0] = 0.0;
          if (stor > tstor) rates[0] = K * (stor - tstor);
        } else {
          double x3 = P_S(c, m, RVN_S_GR4J_X3);
          rates[0] = stor * (1.0 - pow(1.0 + pow(omax(stor / x3, 0.0), 4), -0.25)) / tstep;
        }
      }
      rates[0] = omin(rates[0], omax(sv[iFrom[0]], 0.0) / tstep);
      rates[0] = omax(rates[0], 0.0);
      break;
    }
    case RVN_OP_PERC_LINEAR: case RVN_OP_PERC_CONSTANT: case RVN_OP_PERC_GR4J: case RVN_OP_PERC_GR4JEXCH: case RVN_OP_PERC_GR4JEXCH2:
    case RVN_OP_PERC_GAWSER: case RVN_OP_PERC_GAWSER_CONSTRAIN: case RVN_OP_PERC_POWER_LAW: case RVN_OP_PERC_LINEAR_ANALYTIC: case RVN_OP_PERC_PRMS:
    case RVN_OP_PERC_SACRAMENTO: case RVN_OP_PERC_ASPEN: {
      /* src/Percolation.cpp:183-202,237-243,293-316 + :339-360 */
      if (type == RVN_HRU_LAKE || type == RVN_HRU_WATER || type == RVN_HRU_ROCK) break;
      int m = pr->ia[0];
      double stor = sv[iFrom[0]], max_stor = soil_capacity(c, m);
      if (max_stor > 0.0) {
        if (pr->op == RVN_OP_PERC_LINEAR) rates[0] = P_S(c, m, RVN_S_PERC_COEFF) * stor;
        else if (pr->op == RVN_OP_PERC_CONSTANT) rates[0] = P_S(c, m, RVN_S_MAX_PERC_RATE);
        else if (pr->op == RVN_OP_PERC_GR4J) rates[0] = stor * (1.0 - pow(1.0 + pow(4.0 / 9.0 * omax(stor / max_stor, 0.0), 4), -0.25)) / tstep;
        else if (pr->op == RVN_OP_PERC_GAWSER) { /* :204-213 */
          double field_cap = P_S(c, m, RVN_S_FIELD_CAPACITY) * max_stor, max_rate = P_S(c, m, RVN_S_MAX_PERC_RATE);
          rates[0] = max_rate * omax(stor - field_cap, 0.0) / (max_stor - field_cap);
        } else if (pr->op == RVN_OP_PERC_GAWSER_CONSTRAIN) { /* :215-226 */
          double field_cap = P_S(c, m, RVN_S_FIELD_CAPACITY) * max_stor, max_rate = P_S(c, m, RVN_S_MAX_PERC_RATE);
          rates[0] = 0.0;
          if (stor > field_cap) {
            rates[0] = max_rate * omax(stor - field_cap, 0.0) / (max_stor - field_cap);
            rates[0] = omin(rates[0], (stor - field_cap) / tstep);
          }
        } else if (pr->op == RVN_OP_PERC_POWER_LAW) { rates[0] = P_S(c, m, RVN_S_MAX_PERC_RATE) * pow(stor / max_stor, P_S(c, m, RVN_S_PERC_N));
        } else if (pr->op == RVN_OP_PERC_LINEAR_ANALYTIC) { rates[0] = stor * (1 - exp(-P_S(c, m, RVN_S_PERC_COEFF) * tstep)) / tstep;
        } else if (pr->op == RVN_OP_PERC_PRMS) { /* :253-267 */
          double tens_stor = soil_tension_capacity(c, m), max_rate = P_S(c, m, RVN_S_MAX_PERC_RATE), n = P_S(c, m, RVN_S_PERC_N);
          double free_stor = omax(0.0, stor - tens_stor), free_stor_max = (max_stor - tens_stor);
          rates[0] = max_rate * pow(free_stor / free_stor_max, n);
        } else if (pr->op == RVN_OP_PERC_SACRAMENTO) { /* :269-294 */
          int m2 = pr->ia[1];
          double stor2 = sv[iTo[0]], tens_stor = soil_tension_capacity(c, m);
          double psi = P_S(c, m2, RVN_S_SAC_PERC_EXPON), alpha = P_S(c, m2, RVN_S_SAC_PERC_ALPHA);
          double max_stor2 = (m == 0) ? soil_capacity(c, m2) : 1e99;
          double free_stor = omax(0.0, stor - tens_stor), free_stor_max = (max_stor - tens_stor);
          double max_baseflow = P_S(c, m, RVN_S_MAX_BASEFLOW_RATE);
          double lz_perc = 1.0 + alpha * pow((1 - (stor2 / max_stor2)), psi);
          rates[0] = lz_perc * max_baseflow * (free_stor / free_stor_max);
        } else if (pr->op == RVN_OP_PERC_ASPEN) { rates[0] = P_S(c, m, RVN_S_PERC_ASPEN);
        } else if (pr->op == RVN_OP_PERC_GR4JEXCH) {
          double x2 = P_S(c, m, RVN_S_GR4J_X2), x3 = P_S(c, m, RVN_S_GR4J_X3);
          rates[0] = -x2 * pow(omax(omin(stor / x3, 1.0), 0.0), 3.5);
        } else {
          double x2 = P_S(c, 1, RVN_S_GR4J_X2), x3 = P_S(c, 1, RVN_S_GR4J_X3);
          stor = sv[sv_index(d, RVN_SV_SOIL, 1)];
          rates[0] = -x2 * pow(omax(omin(stor / x3, 1.0), 0.0), 3.5);
        }
      }
      rates[0] = omin(rates[0], omax(sv[iFrom[0]] - 0.0, 0.0) / tstep);
      if (!d->opt.allow_soil_overfill) {
        double room = omax(state_var_max(c, iTo[0], sv) - sv[iTo[0]], 0.0);
        rates[0] = omin(rates[0], room / tstep);
      }
      break;
    }
    case RVN_OP_SOILEVAP_ALL: case RVN_OP_SOILEVAP_HBV: case RVN_OP_SOILEVAP_GR4J: case RVN_OP_SOILEVAP_TOPMODEL:
    case RVN_OP_SOILEVAP_LINEAR: case RVN_OP_SOILEVAP_PDM: case RVN_OP_SOILEVAP_HYMOD2: case RVN_OP_SOILEVAP_UBC: case RVN_OP_SOILEVAP_VIC:
    case RVN_OP_SOILEVAP_HYPR: case RVN_OP_SOILEVAP_SEQUEN: case RVN_OP_SOILEVAP_ROOT: case RVN_OP_SOILEVAP_ROOT_CONSTRAIN: case RVN_OP_SOILEVAP_AWBM: case RVN_OP_SOILEVAP_SACSMA: case RVN_OP_SOILEVAP_CHU: { /* src/SoilEvaporation.cpp:325-343,379-405,640-650 + :767-783 */
      if (type != RVN_HRU_STANDARD) break;
      double PET = F[RVN_F_PET];
      if (!d->opt.suppress_competitive_ET) { PET -= (sv[c->iSV[RVN_SV_AET]] / tstep); PET = omax(PET, 0.0); }
      double PETused = 0.0; int multi = 0;
      if (pr->op == RVN_OP_SOILEVAP_ALL) rates[0] = PET;
      else if (pr->op == RVN_OP_SOILEVAP_HBV || pr->op == RVN_OP_SOILEVAP_TOPMODEL) {
        double stor = omax(sv[iFrom[0]], 0.0), tens_stor = soil_tension_capacity(c, 0);
        rates[0] = PET * omin(stor / tens_stor, 1.0);
        if (pr->op == RVN_OP_SOILEVAP_HBV) { /* correction for snow in non-forested areas */
          int iSnow = c->iSV[RVN_SV_SNOW];
          double Fc = P_LU(c, RVN_LU_FOREST_COVERAGE);
          if ((iSnow >= 0) && (sv[iSnow] > O_REAL_SMALL)) rates[0] = (Fc)*rates[0];
        }
      } else if (pr->op == RVN_OP_SOILEVAP_GR4J) {
        double max_stor = soil_capacity(c, 0), stor = sv[iFrom[0]];
        double sat = stor / max_stor;
        double tmp = tanh(omax(PET, 0.0) / max_stor);
        rates[0] = stor * (2.0 - sat) * tmp / (1.0 + (1.0 - sat) * tmp);
      } else if (pr->op == RVN_OP_SOILEVAP_HYPR) { /* :385-420 */
        double stor = omax(sv[iFrom[0]], 0.0), tens_stor = soil_tension_capacity(c, 0);
        rates[0] = PET * omin(stor / tens_stor, 1.0);
        int iDep = c->iSV[RVN_SV_DEPRESSION];
        double maxPondedAreaFrac = P_LU(c, RVN_LU_MAX_DEP_AREA_FRAC), n = P_LU(c, RVN_LU_PONDED_EXP), dep_max = P_LU(c, RVN_LU_DEP_MAX);
        double area_ponded = maxPondedAreaFrac * pow(omin(sv[iDep] / dep_max, 1.0), n);
        rates[0] = (1.0 - area_ponded) * rates[0];
        rates[1] = (area_ponded)*F[RVN_F_OW_PET];
        PETused = (rates[0] + rates[1]); multi = 1;
      } else if (pr->op == RVN_OP_SOILEVAP_CHU) { /* :460-468 */
        double CHU = omax(sv[c->iSV[RVN_SV_CROP_HEAT_UNITS]], 0.0);
        rates[0] = PET * omin(CHU / P_V(c, RVN_V_CHU_MATURITY), 1.0);
      } else if (pr->op == RVN_OP_SOILEVAP_LINEAR) { /* :370-377 */
        double stor = sv[iFrom[0]], alpha = P_LU(c, RVN_LU_AET_COEFF);
        rates[0] = omin(alpha * stor, PET);
      } else if (pr->op == RVN_OP_SOILEVAP_PDM || pr->op == RVN_OP_SOILEVAP_HYMOD2) { /* :424-458 */
        double stor = sv[iFrom[0]], max_stor = soil_capacity(c, 0), b = P_LU(c, RVN_LU_PDM_B);
        double c_max = (b + 1) * max_stor;
        double c_star = c_max * (1.0 - pow(1.0 - (stor / max_stor), 1.0 / (b + 1.0)));
        if (pr->op == RVN_OP_SOILEVAP_PDM) rates[0] = omin(c_star, (c_star / c_max) * PET);
        else {
          double gp = P_LU(c, RVN_LU_HYMOD2_G), Kmax = P_LU(c, RVN_LU_HYMOD2_KMAX), ce = P_LU(c, RVN_LU_HYMOD2_EXP);
          double K = Kmax * (gp + (1 - gp) * pow(c_star / c_max, ce));
          rates[0] = K * omin(c_star, PET);
        }
      } else if (pr->op == RVN_OP_SOILEVAP_UBC) { /* :470-494 */
        double P0AGEN = P_S(c, 0, RVN_S_UBC_INFIL_SOIL_DEF), P0EGEN = P_S(c, 0, RVN_S_UBC_EVAP_SOIL_DEF);
        double soil_deficit = omax(soil_capacity(c, 0) - sv[iFrom[0]], 0.0);
        double Fimp = P_LU(c, RVN_LU_IMPERMEABLE_FRAC);
        double AET = (PET)*pow(10.0, -soil_deficit / P0EGEN);
        double total_precip = sv[c->iSV[RVN_SV_PONDED_WATER]];
        double soil_deficit_est = omax(soil_deficit - total_precip + AET * tstep, 0.0);
        double b1 = 0.0;
        if (Fimp < 1.0) b1 = Fimp * pow(10.0, -soil_deficit_est / P0AGEN);
        rates[0] = (PET)*pow(10.0, -soil_deficit_est / P0EGEN) * (1.0 - b1);
      } else if (pr->op == RVN_OP_SOILEVAP_VIC) { /* :496-515 */
        double stor = sv[iFrom[0]], stor_max = soil_capacity(c, 0);
        double alpha = P_S(c, 0, RVN_S_VIC_ALPHA), zmax = P_S(c, 0, RVN_S_VIC_ZMAX), zmin = P_S(c, 0, RVN_S_VIC_ZMIN), gamma2 = P_S(c, 0, RVN_S_VIC_EVAP_GAMMA);
        double Smax = 1.0 / (alpha + 1.0) * (alpha * zmax + zmin);
        double Sat = stor / stor_max;
        rates[0] = PET * (1.0 - pow(1.0 - Sat / Smax, gamma2));
      } else if (pr->op == RVN_OP_SOILEVAP_SEQUEN) { /* :553-568 */
        double stor_u = omax(sv[iFrom[0]], 0.0), stor_l = omax(sv[iFrom[1]], 0.0);
        double tens_stor_u = soil_tension_capacity(c, 0), tens_stor_l = soil_tension_capacity(c, 1);
        rates[0] = (PET)*omin(stor_u / tens_stor_u, 1.0);
        rates[1] = (PET - rates[0]) * omin(stor_l / tens_stor_l, 1.0);
        PETused = rates[0] + rates[1]; multi = 1;
      } else if (pr->op == RVN_OP_SOILEVAP_ROOT) { /* :570-596 */
        double rootfrac_u = 0.7, rootfrac_l = 1.0 - rootfrac_u;
        double stor_u = sv[iFrom[0]];
        double tens_stor_u = soil_tension_capacity(c, 0); tens_stor_u = omax(0.0001, tens_stor_u);
        rates[0] = PET * rootfrac_u * omin(stor_u / tens_stor_u, 1.0);
        double stor_l = sv[iFrom[1]];
        double tens_stor_l = soil_tension_capacity(c, 1); tens_stor_l = omax(0.0001, tens_stor_l);
        rates[1] = PET * rootfrac_l * omin(stor_l / tens_stor_l, 1.0);
        PETused = rates[0] + rates[1]; multi = 1;
      } else if (pr->op == RVN_OP_SOILEVAP_ROOT_CONSTRAIN) { /* :598-622 */
        double rootfrac_u = 0.7;
        double stor_u = sv[iFrom[0]], tens_stor_u = soil_tension_capacity(c, 0);
        double stor_wilt = P_S(c, 0, RVN_S_SAT_WILT) * soil_capacity(c, 0);
        rates[0] = PET * rootfrac_u * omin(stor_u / tens_stor_u, 1.0);
        if (rates[0] > omax(stor_u - stor_wilt, 0.0)) rates[0] = omax(stor_u - stor_wilt, 0.0);
        double rootfrac_l = 1.0 - rootfrac_u;
        rates[1] = PET * rootfrac_l * 1;
        PETused = rates[0] + rates[1]; multi = 1;
      } else if (pr->op == RVN_OP_SOILEVAP_SACSMA) { /* :652-715 */
        double red, e1, e2, e3, e5;
        double Adimp = P_LU(c, RVN_LU_MAX_SAT_AREA_FRAC), Aimp = P_LU(c, RVN_LU_IMPERMEABLE_FRAC), rserv = P_S(c, 3, RVN_S_UNAVAIL_FRAC);
        double Aperv = 1.0 - Adimp - Aimp;
        double uzt_stor_max = soil_capacity(c, 0), uzf_stor_max = soil_capacity(c, 1), lzt_stor_max = soil_capacity(c, 2);
        double lzfp_stor_max = soil_capacity(c, 3), lzfs_stor_max = soil_capacity(c, 4);
        double uzt_stor = sv[sv_index(d, RVN_SV_SOIL, 0)] / Aperv, uzf_stor = sv[sv_index(d, RVN_SV_SOIL, 1)] / Aperv;
        double lzt_stor = sv[sv_index(d, RVN_SV_SOIL, 2)] / Aperv, lzfp_stor = sv[sv_index(d, RVN_SV_SOIL, 3)] / Aperv;
        double lzfs_stor = sv[sv_index(d, RVN_SV_SOIL, 4)] / Aperv, adimc_stor = sv[sv_index(d, RVN_SV_SOIL, 5)] / Adimp;
        if (Adimp == 0) { adimc_stor = 0; }
        e1 = omin(PET * (uzt_stor / uzt_stor_max), uzt_stor);
        red = PET - e1;
        uzt_stor -= e1;
        rates[0] = e1 * Aperv / tstep;
        e2 = omin(red, uzf_stor);
        red -= e2;
        uzf_stor -= e2;
        rates[1] = e2 * Aperv / tstep;
        if ((uzt_stor / uzt_stor_max) < (uzf_stor / uzf_stor_max)) {
          double delta = uzt_stor_max * (uzt_stor + uzf_stor) / (uzt_stor_max + uzf_stor_max) - uzt_stor;
          uzt_stor += delta;
          uzf_stor -= delta;
          rates[2] = delta * Aperv / tstep;
        }
        e3 = omin(red * (lzt_stor / (uzt_stor_max + lzt_stor_max)), lzt_stor);
        lzt_stor -= e3;
        rates[3] = e3 * Aperv / tstep;
        double ratlzt = lzt_stor / lzt_stor_max;
        double saved = rserv * (lzfp_stor_max + lzfs_stor_max);
        double ratlz = (lzt_stor + lzfp_stor + lzfs_stor - saved) / (lzt_stor_max + lzfp_stor_max + lzfs_stor_max - saved);
        if (ratlzt < ratlz) {
          double del = omin((ratlz - ratlzt) * lzt_stor_max, lzfs_stor);
          lzt_stor += del;
          lzfs_stor -= del;
        }
        e5 = omin(e1 + (red + e2) * (adimc_stor - uzt_stor - e1) / (uzt_stor_max + lzt_stor_max), adimc_stor);
        adimc_stor -= e5;
        rates[5] = e5 * Adimp / tstep;
        PETused = (e1 + e2 + e3) * Aperv + e5 * Adimp; multi = 1;
      } else if (pr->op == RVN_OP_SOILEVAP_AWBM) { /* :716-727 */
        double a1 = P_LU(c, RVN_LU_AWBM_AREAFRAC1), a2 = P_LU(c, RVN_LU_AWBM_AREAFRAC2);
        double a3 = 1.0 - a1 - a2;
        rates[0] = omin(a1 * PET, sv[iFrom[0]] / tstep);
        rates[1] = omin(a2 * PET, sv[iFrom[1]] / tstep);
        rates[2] = omin(a3 * PET, sv[iFrom[2]] / tstep);
        PETused = rates[0] + rates[1] + rates[2]; multi = 1;
      }
      rates[pr->nconn - 1] = multi ? PETused : rates[0]; /* PETused -> AET (src/SoilEvaporation.cpp end of GetRatesOfChange) */
      double corr = 0, oldrate;
      for (int q = 0; q < pr->nconn - 1; q++) {
        oldrate = rates[q];
        rates[q] = omin(rates[q], sv[iFrom[q]] / tstep);
        corr += oldrate - rates[q];
      }
      rates[pr->nconn - 1] -= corr;
      break;
    }
    case RVN_OP_DEPRESSION_OVERFLOW: { /* CmvDepressionOverflow, src/DepressionProcesses.cpp:106-165 */
      double stor = sv[iFrom[0]], thresh_stor = P_LU(c, RVN_LU_DEP_THRESHOLD);
      if (pr->ia[0] == 0) {
        double max_flow = P_LU(c, RVN_LU_DEP_MAX_FLOW), n = P_LU(c, RVN_LU_DEP_N), max_stor = P_LU(c, RVN_LU_DEP_MAX);
        if (max_stor < thresh_stor) rates[0] = max_flow;
        else rates[0] = max_flow * pow(omax((stor - thresh_stor) / (max_stor - thresh_stor), 0.0), n);
      } else if (pr->ia[0] == 1) {
        double dep_k = P_LU(c, RVN_LU_DEP_K);
        rates[0] = omax(stor - thresh_stor, 0.0) * (1 - exp(-dep_k * tstep)) / tstep;
      } else {
        double dep_rat = P_LU(c, RVN_LU_DEP_CRESTRATIO), area = d->hru_area[c->k];
        double cw = dep_rat * sqrt(area);
        rates[0] = 0.666 * cw / area * sqrt(2 * 9.80616) * pow(omax(stor - thresh_stor, 0.0), 1.5);
      }
      rates[0] = omin(rates[0], omax(sv[iFrom[0]] / tstep, 0.0));
      break;
    }
    case RVN_OP_SEEPAGE: { /* CmvSeepage SEEP_LINEAR, src/DepressionProcesses.cpp:254-298 */
      double stor = sv[iFrom[0]], dep_k = P_LU(c, RVN_LU_DEP_SEEP_K), room;
      rates[0] = omax(stor, 0.0) * (1 - exp(-dep_k * tstep)) / tstep;
      rates[0] = omin(rates[0], omax(sv[iFrom[0]] / tstep, 0.0));
      room = omax(state_var_max(c, iTo[0], sv) - sv[iTo[0]], 0.0);
      rates[0] = omin(rates[0], room / tstep);
      break;
    }
    case RVN_OP_EXCHANGE_FLOW: { /* CmvExchangeFlow, src/Flush.cpp:358-368 */
      rates[0] = P_S(c, pr->ia[0], RVN_S_EXCHANGE_FLOW);
      rates[1] = P_S(c, pr->ia[0], RVN_S_EXCHANGE_FLOW);
      break;
    }
    case RVN_OP_CHU_ONTARIO: { /* CmvCropHeatUnitEvolve, src/CropGrowth.cpp:91-136 */
      if (type != RVN_HRU_STANDARD) break;
      double CHU = sv[iFrom[0]], old_CHU = CHU, CHU_day, CHU_night;
      double temp_ave = F[RVN_F_TEMP_DAILY_AVE], temp_max = F[RVN_F_TEMP_DAILY_MAX], temp_min = F[RVN_F_TEMP_DAILY_MIN];
      int growing_season = (CHU > -0.5);
      if (!growing_season) { if (temp_ave > 12.8) { CHU = 0.0; } }
      else {
        CHU_day = omax(3.33 * (temp_max - 10) - 0.084 * pow(temp_max - 10.0, 2.0), 0.0);
        CHU_night = omax(1.8 * (temp_min - 4.4), 0.0);
        CHU += 0.5 * (CHU_day + CHU_night);
        if (temp_ave < -2.0) { CHU = -1.0; }
        if (d->times[c->step].month > 9) { CHU = -1.0; }
      }
      rates[0] = (CHU - old_CHU) / tstep;
      break;
    }
    case RVN_OP_SNOBAL_GAWSER: { /* CmvSnowBalance::GawserBalance, src/SnowBalance.cpp:1098-1201 */
      double SWC = sv[iFrom[0]], rainthru = sv[iFrom[1]], snow_liq = sv[iFrom[2]], snow_depth = sv[iFrom[3]], newSnow = sv[iFrom[6]];
      double KF = P_LU(c, RVN_LU_REFREEZE_FACTOR), KM = P_LU(c, RVN_LU_MELT_FACTOR), SWI = P_G(c, RVN_G_SNOW_SWI);
      double RHOICE = O_DENSITY_ICE / O_DENSITY_WATER, MRHO = 0.35, a = 0.1, b = 96, NEWDEN = 0.1;
      double TEMPs = F[RVN_F_TEMP_AVE], TBAS = 0;
      double MELTP = 0, REFRZP = 0;
      if (TEMPs < TBAS) { REFRZP = tstep * KF * (TBAS - TEMPs); }
      if (TEMPs > TBAS) { MELTP = tstep * KM * (TEMPs - TBAS); }
      double RHO = 0;
      if (SWC > 0) { RHO = omin(MRHO, SWC / snow_depth); }
      double POR = 1 - RHO / RHOICE;
      double snow_liqAP = POR * snow_depth * SWI;
      double RAIN_IN_1 = 0, MELT_1 = 0, MELT_2 = 0, EXCESS_snow_liq = 0;
      double REFRZ = 0;
      if (TBAS > TEMPs) { REFRZ = omin(REFRZP, snow_liq); }
      double SWC2 = SWC + REFRZ;
      double RHO2 = omin(MRHO, SWC2 / snow_depth);
      double snow_liq2 = snow_liq - REFRZ;
      if (snow_liq2 >= snow_liqAP) { EXCESS_snow_liq = snow_liq2 - snow_liqAP; MELT_2 = omin(MELTP, SWC); }
      else { MELT_1 = omin(MELTP, SWC); RAIN_IN_1 = rainthru; }
      double KC = b * exp(-1 * a * TEMPs);
      double TRHO = (RHO2 * MRHO) / (RHO2 + (MRHO - RHO2) * exp(-1 * tstep * 24 / KC));
      double compaction = 0;
      if (SWC2 > 0) { compaction = snow_depth - SWC2 / TRHO; }
      double snowmelt = 0;
      if (SWC2 > 0) { snowmelt = (MELT_1 + MELT_2) / TRHO; }
      double SUBLIM = 0;
      if (TEMPs < 0) { SUBLIM = 2.4 * tstep; }
      if (SWC2 - MELT_1 - MELT_2 < SUBLIM) { SUBLIM = omax(omin(SUBLIM, SWC2 - MELT_1 - MELT_2), 0.0); }
      double SUBLIM_snow_depth = 0;
      if (SWC2 > 0) { SUBLIM_snow_depth = SUBLIM / TRHO; }
      rates[0] = MELT_1 / tstep;
      rates[1] = RAIN_IN_1 / tstep;
      rates[2] = REFRZ / tstep;
      rates[3] = -compaction / tstep;
      rates[4] = -(snowmelt + SUBLIM_snow_depth) / tstep;
      rates[5] = newSnow / NEWDEN / tstep;
      rates[6] = newSnow / tstep;
      rates[7] = SUBLIM / tstep;
      rates[8] = EXCESS_snow_liq / tstep;
      rates[9] = MELT_2 / tstep;
      break;
    }
    case RVN_OP_SNOBAL_CRHM_EBSM: { /* CmvSnowBalance::CRHMSnowBalance, src/SnowBalance.cpp:1214-1291 */
      const double LH_FUSION = 0.334;
      double snow_liq = sv[c->iSV[RVN_SV_SNOW_LIQ]], SWE = sv[c->iSV[RVN_SV_SNOW]], snow_energy = -sv[c->iSV[RVN_SV_COLD_CONTENT]];
      double snow_energy_old = snow_energy;
      if (SWE <= 0.0) break;
      double pot_melt = F[RVN_F_POTENTIAL_MELT], T_min = F[RVN_F_TEMP_DAILY_MIN];
      pot_melt *= LH_FUSION * O_DENSITY_WATER / O_MM_PER_METER * tstep;
      double snoliq_max = P_G(c, RVN_G_SNOW_SWI) * SWE;
      double t_minus = omin(T_min, 0.0);
      double Umin = SWE * (2.115 + 0.00779 * t_minus) * t_minus;
      double refreeze = 0.0, snowmelt0 = 0.0, snowmelt1 = 0.0, snowmelt2 = 0.0;
      snow_energy += pot_melt;
      snow_energy = omax(Umin, snow_energy);
      if (snow_energy > 0.0) {
        double B = 0.95;
        double melt = snow_energy / LH_FUSION / O_DENSITY_WATER * O_MM_PER_METER / B;
        if (melt > (snoliq_max - snow_liq)) {
          if (SWE > melt) { snowmelt0 = snoliq_max - snow_liq; snowmelt1 = melt - (snoliq_max - snow_liq); snowmelt2 = 0.0; }
          else { snowmelt0 = 0; snowmelt1 = SWE; snowmelt2 = snow_liq; }
        } else { snowmelt0 = melt; snowmelt1 = 0; snowmelt2 = 0; }
        snow_energy = 0.0;
      } else {
        snowmelt0 = snowmelt1 = snowmelt2 = 0;
        refreeze = -snow_energy / LH_FUSION / O_DENSITY_WATER * O_MM_PER_METER;
        if (snow_liq > refreeze) { snow_energy = 0.0; }
        else { snow_energy += snow_liq * LH_FUSION * O_DENSITY_WATER / O_MM_PER_METER; refreeze = snow_liq; }
      }
      rates[0] = snowmelt0; rates[1] = snowmelt1; rates[2] = snowmelt2; rates[3] = refreeze;
      rates[4] = -(snow_energy - snow_energy_old) / tstep;
      break;
    }
    case RVN_OP_SNOBAL_COLD_CONTENT: { /* CmvSnowBalance::ColdContentBalance, src/SnowBalance.cpp:732-835 */
      const double CCFAC = 0.3, MELT_FAC = 1.5, LAIMLT = 0.2, SAIMLT = 0.5, LH_FUSION = 0.334, SPH_ICE = 2.100e-3, SPH_WATER = 4.186e-3;
      const double SWI = P_G(c, RVN_G_SNOW_SWI);
      double Ta = F[RVN_F_TEMP_DAILY_AVE], day_length = c->day_length;
      double CC = sv[iFrom[0]], SL = sv[iFrom[1]], S = sv[iFrom[3]];
      double LAI = c->veg_LAI, SAI = c->veg_SAI;
      double incoming_snow_en, pot_melt, CC_air, liq_snow_cap, rainthru = 0.0, Tsnow, instant_refreeze;
      double refreeze, cooling, warming, melt_to_liq, melt_to_SW, shrinking_poro, eq_en_avail;
      if (S == 0) break;
      if ((CC > 0) && (SL > 0)) {
        instant_refreeze = omin(SL / tstep, CC / LH_FUSION / tstep);
        rates[1] += instant_refreeze;
        rates[0] -= instant_refreeze * LH_FUSION;
        SL -= rates[1] * tstep;
      }
      Tsnow = O_FREEZING_TEMP - CC / S / SPH_ICE;
      if (Ta <= O_FREEZING_TEMP) incoming_snow_en = CCFAC * 2.0 * day_length * (Ta - Tsnow);
      else incoming_snow_en = MELT_FAC * 2.0 * day_length * (Ta - O_FREEZING_TEMP) * exp(-SAIMLT * SAI) * exp(-LAIMLT * LAI);
      incoming_snow_en += rainthru * omax(Ta - O_FREEZING_TEMP, 0.0) * SPH_WATER * O_DENSITY_WATER / O_MM_PER_METER;
      pot_melt = incoming_snow_en * (1.0 / O_DENSITY_WATER / LH_FUSION * O_MM_PER_METER);
      CC_air = (O_FREEZING_TEMP - Ta) * SPH_ICE * S;
      liq_snow_cap = SWI * S; /* CalculateSnowLiquidCapacity, src/SnowParams.cpp:79-88 */
      if (pot_melt <= 0) {
        refreeze = omin(SL / tstep, -pot_melt);
        rates[1] += refreeze;
        pot_melt += refreeze;
        cooling = omin(-pot_melt, 0.0) * LH_FUSION;
        cooling = omin(cooling, (CC_air - CC) / tstep);
        rates[2] += cooling;
      } else {
        if ((pot_melt * LH_FUSION < CC / tstep) || (Ta < O_FREEZING_TEMP)) {
          warming = pot_melt * LH_FUSION;
          warming = omin(warming, -(CC - CC_air) / tstep);
          rates[0] -= warming;
        } else {
          eq_en_avail = pot_melt;
          warming = CC / tstep;
          rates[0] += warming;
          eq_en_avail -= warming / LH_FUSION;
          melt_to_liq = omin(eq_en_avail, (liq_snow_cap - SL) / tstep);
          rates[1] -= melt_to_liq;
          eq_en_avail -= melt_to_liq;
          melt_to_SW = omin(eq_en_avail, S / tstep - melt_to_liq);
          rates[3] += melt_to_SW;
          shrinking_poro = melt_to_SW * SWI;
          rates[4] += shrinking_poro;
        }
      }
      break;
    }
    case RVN_OP_SNOBAL_HBV: { /* src/SnowBalance.cpp:430-465; liquid capacity src/SnowParams.cpp:79-88 */
      double Ka = P_LU(c, RVN_LU_REFREEZE_FACTOR), Ta = F[RVN_F_TEMP_DAILY_AVE];
      double melt, refreeze, liq_cap, to_liq, overflow;
      double SWE = sv[iFrom[0]], SL = sv[iTo[0]];
      melt = omax(F[RVN_F_POTENTIAL_MELT], 0.0);
      melt = omin(melt, SWE / tstep);
      SWE -= melt * tstep;
      refreeze = Ka * omax(O_FREEZING_TEMP - Ta, 0.0);
      refreeze = omax(omin(SL / tstep, refreeze), 0.0);
      SL -= refreeze * tstep;
      liq_cap = P_G(c, RVN_G_SNOW_SWI) * SWE;
      to_liq = omin(melt, omax(liq_cap - SL, 0.0) / tstep);
      SL += to_liq * tstep;
      overflow = omax((SL - liq_cap) / tstep, 0.0);
      rates[0] = -refreeze;
      rates[0] += to_liq;
      rates[1] = omax(melt - to_liq, 0.0);
      rates[2] = overflow;
      if (type != RVN_HRU_STANDARD) { rates[3] = rates[1]; rates[1] = 0.0; rates[4] = rates[2]; rates[2] = 0.0; }
      break;
    }
    case RVN_OP_INTERFLOW_PRMS: { /* src/Interflow.cpp:98-124 + :136-148 */
      if (type == RVN_HRU_LAKE || type == RVN_HRU_WATER || type == RVN_HRU_ROCK) break;
      int m = pr->ia[0];
      double stor = sv[iFrom[0]];
      double max_stor = soil_capacity(c, m), tens_stor = soil_tension_capacity(c, m), K = P_S(c, m, RVN_S_MAX_INTERFLOW_RATE);
      rates[0] = K * omax(stor - tens_stor, 0.0) / (max_stor - tens_stor);
      rates[0] = omin(rates[0], sv[iFrom[0]] / tstep);
      break;
    }
    case RVN_OP_SNOWMELT_POTMELT: /* src/SnowMeltRefreeze.cpp:87-131 */
      rates[0] = omax(F[RVN_F_POTENTIAL_MELT], 0.0);
      if (sv[iFrom[0]] <= 0) rates[0] = 0.0;
      if (rates[0] < 0.0) rates[0] = 0.0;
      rates[0] = omin(rates[0], sv[iFrom[0]] / tstep);
      break;
    case RVN_OP_SNOWSQUEEZE: { /* src/SnowMeltRefreeze.cpp:197-244 (no SNOW_DEPTH variable) */
      double S = sv[c->iSV[RVN_SV_SNOW]], SL = sv[c->iSV[RVN_SV_SNOW_LIQ]];
      double liq_cap = P_G(c, RVN_G_SNOW_SWI) * S;
      rates[0] = omax(SL - liq_cap, 0.0) / tstep;
      if (sv[iFrom[0]] <= 0) rates[0] = 0.0;
      if (rates[0] < 0.0) rates[0] = 0.0;
      break;
    }
    case RVN_OP_SNOBAL_SIMPLE_MELT: /* src/SnowBalance.cpp:393-396 + :1311-1317 */
      rates[0] = omax(F[RVN_F_POTENTIAL_MELT], 0.0);
      if (rates[0] < 0.0) rates[0] = 0.0;
      rates[0] = omin(rates[0], omax(sv[iFrom[0]] / tstep, 0.0));
      break;
    case RVN_OP_SNOBAL_CEMA_NEIGE: { /* src/SnowBalance.cpp:398-413 */
      double avg_annual_snow = P_G(c, RVN_G_AVG_ANNUAL_SNOW), snotemp = snow_temperature(c), pot_melt = 0.0, SWE, snow_cov;
      if (snotemp == O_FREEZING_TEMP) pot_melt = omax(F[RVN_F_POTENTIAL_MELT], 0.0);
      SWE = sv[iFrom[0]];
      snow_cov = omin(SWE / avg_annual_snow, 1.0);
      rates[0] = (0.9 * snow_cov + 0.1) * omin(SWE / tstep, pot_melt);
      rates[1] = (snow_cov - sv[iFrom[1]]) / tstep;
      break;
    }
    case RVN_OP_SNOWREFREEZE_DD: { /* src/SnowMeltRefreeze.cpp:324-365 */
      double Kf = P_LU(c, RVN_LU_REFREEZE_FACTOR), Ta = F[RVN_F_TEMP_DAILY_AVE];
      rates[0] = omax(Kf * (O_FREEZING_TEMP - Ta), 0.0);
      if (sv[iFrom[0]] <= 0) rates[0] = 0.0;
      rates[0] = omin(rates[0], sv[iFrom[0]] / tstep);
      if (rates[0] < 0) rates[0] = 0.0;
      break;
    }
    case RVN_OP_SNOTEMP_NEWTONS: { /* src/SnowTempEvolve.cpp:42-77 */
      double Tsnow = sv[iFrom[0]], Tair = F[RVN_F_TEMP_AVE];
      rates[0] = P_G(c, RVN_G_AIRSNOW_COEFF) * (Tair - Tsnow);
      rates[0] = omin(rates[0], (O_FREEZING_TEMP - Tsnow) / tstep);
      break;
    }
    case RVN_OP_CANEVP_ALL: case RVN_OP_CANSNOWEVP_ALL: { /* src/VegetationMovers.cpp:110-190 and :283-346 */
      if ((type != RVN_HRU_STANDARD) && (type != RVN_HRU_WETLAND)) break;
      double Fc = P_LU(c, RVN_LU_FOREST_COVERAGE), oldRates;
      if (Fc != 0) { rates[0] = sv[iFrom[0]] / tstep; rates[1] = rates[0]; }
      oldRates = rates[0];
      if (pr->op == RVN_OP_CANEVP_ALL) rates[0] = omax(rates[0], 0.0);
      rates[0] = omin(rates[0], sv[iFrom[0]] / tstep);
      rates[1] -= (oldRates - rates[0]);
      break;
    }
    case RVN_OP_CANEVP_MAXIMUM: case RVN_OP_CANEVP_RUTTER: case RVN_OP_CANSNOWEVP_MAXIMUM: { /* src/VegetationMovers.cpp:110-190, 291-346 */
      if ((type != RVN_HRU_STANDARD) && (type != RVN_HRU_WETLAND)) break;
      double Fc = P_LU(c, RVN_LU_FOREST_COVERAGE), oldRates;
      if (Fc != 0) {
        double PET = omax(F[RVN_F_PET], 0.0), PETused;
        if (!d->opt.suppress_competitive_ET) { PET -= (sv[c->iSV[RVN_SV_AET]] / tstep); PET = omax(PET, 0.0); }
        if (pr->op == RVN_OP_CANEVP_RUTTER) { /* trunk fraction forced to 0: no TRUNK state variable */
          double cap = c->canopy_cap, stor = omin(omax(sv[iFrom[0]], 0.0), cap * Fc), Ft = 0.0;
          rates[0] = (1.0 - Ft) * Fc * PET * (stor / (cap * Fc));
        } else rates[0] = Fc * PET;
        PETused = rates[0];
        rates[1] = PETused;
      }
      oldRates = rates[0];
      if (pr->op != RVN_OP_CANSNOWEVP_MAXIMUM) rates[0] = omax(rates[0], 0.0);
      rates[0] = omin(rates[0], sv[iFrom[0]] / tstep);
      rates[1] -= (oldRates - rates[0]);
      break;
    }
    case RVN_OP_ABSTRACTION: { /* CmvAbstraction ABST_PERCENTAGE / ABST_FILL, src/Abstraction.cpp:230-256, 514-532 */
      double ponded = sv[iFrom[0]], depression = sv[iTo[0]];
      if (type != RVN_HRU_STANDARD) break;
      if (pr->ia[0] == RVN_ABST_PERCENTAGE) rates[0] = P_LU(c, RVN_LU_ABST_PERCENT) * ponded / tstep;
      else if (pr->ia[0] == RVN_ABST_SCS) { /* src/Abstraction.cpp:257-285 */
        double S, CN, TR, Ia;
        int condition = 2, month = d->times[c->step].month;
        TR = g_precip5[c->k] / 25.4;
        CN = P_LU(c, RVN_LU_SCS_CN);
        if ((month > 4) && (month < 9)) { if (TR < 1.4) { condition = 1; } else if (TR > 2.1) { condition = 3; } }
        else { if (TR < 0.5) { condition = 1; } else if (TR > 1.1) { condition = 3; } }
        if (condition == 1) { CN = 5E-05 * pow(CN, 3) + 0.0008 * pow(CN, 2) + 0.4431 * CN; }
        else if (condition == 3) { CN = 0.0003 * pow(CN, 3) - 0.0185 * pow(CN, 2) + 2.1586 * CN; }
        S = 25.4 * (1000 / CN - 10);
        Ia = (P_LU(c, RVN_LU_SCS_IA_FRACTION)) * S;
        rates[0] = omax(Ia, ponded) / tstep;
      } else if (pr->ia[0] == RVN_ABST_PDMROF) { /* :287-309 */
        double dep_max = P_LU(c, RVN_LU_DEP_MAX), b = P_LU(c, RVN_LU_PDMROF_B);
        double c_max = (b + 1) * dep_max;
        double c_star = c_max * (1.0 - pow(1.0 - (depression / dep_max), 1.0 / (b + 1.0)));
        double abstracted = dep_max * (pow(1.0 - (c_star / c_max), b + 1.0) - pow(1.0 - omin(c_star + ponded, c_max) / c_max, b + 1.0));
        abstracted = omin(abstracted, ponded);
        rates[0] = (abstracted) / tstep;
        rates[1] = (ponded - abstracted) / tstep;
      } else { double dep_space = omax(P_LU(c, RVN_LU_DEP_MAX) - depression, 0.0); rates[0] = omin(dep_space, omax(ponded, 0.0)) / tstep; }
      {
        double pond = omax(sv[iFrom[0]], 0.0), stor = omax(sv[iTo[0]], 0.0), max_stor = state_var_max(c, iTo[0], sv), deficit;
        rates[0] = omin(rates[0], pond / tstep);
        deficit = omax(max_stor - stor, 0.0);
        rates[0] = omin(rates[0], deficit / tstep);
      }
      break;
    }
    case RVN_OP_SUBLIMATION: { /* CmvSublimation, src/Sublimation.cpp:98-243: SublimationRate + depth-change term + ApplyConstraints;
                                * forcings, snow temperature, SNOW_DEPTH and SWE of the depth term are the HRU's lagged values */
      const int styp = pr->ia[0];
      double Ta = F[RVN_F_TEMP_AVE], rel_hum = F[RVN_F_REL_HUMIDITY], wind_vel = F[RVN_F_WIND_VEL], air_pres = F[RVN_F_AIR_PRES];
      double Tsnow = snow_temperature(c), sublim = 0.0;
      double air_dens = (air_pres - 0.378 * o_sat_vap(Ta)) / (287.042 * (Ta + 273.16)) * 1000; /* GetAirDensity, src/CommonFunctions.cpp:1039-1043 */
      if (styp == RVN_SUBLIM_SVERDRUP) {
        double roughness = P_G(c, RVN_G_SNOW_ROUGHNESS), es = o_sat_vap(Tsnow), ea = o_sat_vap(Ta) * rel_hum;
        double C_E = 0.42 * 0.42 / (log(8.0 / roughness)) / (log(8.0 / roughness));
        sublim = air_dens * C_E * wind_vel * 0.623 * (es - ea) / air_pres * O_SEC_PER_DAY / O_DENSITY_WATER * O_MM_PER_METER;
      } else if (styp == RVN_SUBLIM_KUZMIN) {
        double ea = o_sat_vap(F[RVN_F_TEMP_DAILY_AVE]) * rel_hum * 10, es = o_sat_vap(Tsnow) * 1.0 * 10;
        sublim = ((0.18 + 0.098 * wind_vel) * (es - ea));
      } else if (styp == RVN_SUBLIM_KUCHMENT_GELFAN) {
        double ea = o_sat_vap(Ta) * rel_hum * 10, es = o_sat_vap(Tsnow) * 1.0 * 10;
        double Q_E = 32.82 * (0.18 + 0.098 * wind_vel) * (es - ea) * 0.0864;
        sublim = Q_E / 2.845 / O_DENSITY_WATER * (O_MM_PER_METER);
      } else if (styp == RVN_SUBLIM_BULK_AERO) {
        double ea = o_sat_vap(Ta) * rel_hum, es = o_sat_vap(Tsnow) * 1.0;
        double C_E = (0.42 * 0.42) / (log(2.0 / 0.00005) * log(2.0 / 0.00005));
        double Q_E = air_dens * 2.845 * C_E * wind_vel * ((0.622 * (es - ea)) / air_pres) * O_SEC_PER_DAY;
        sublim = Q_E / (2.845 * O_DENSITY_WATER) * O_MM_PER_METER;
      } else if (styp == RVN_SUBLIM_CENTRAL_SIERRA) {
        double sat_vap = o_sat_vap(Tsnow) * 10, vap_pres = o_sat_vap(F[RVN_F_TEMP_DAILY_AVE]) * rel_hum * 10;
        sublim = ((0.0063 * (pow((2.0 * 3.28 * 2.0 * 3.28), (-1 / 6))) * (wind_vel * 2.237) * (sat_vap - vap_pres)) * 25.4); /* (-1/6) == 0 */
      }
      rates[0] = sublim;
      if (pr->nconn == 2) {
        double SD = c->phi_old[c->iSV[RVN_SV_SNOW_DEPTH]], SWE = c->phi_old[iFrom[0]];
        if (SWE > 0) rates[1] = rates[0] * SD / SWE;
      }
      if (sv[iFrom[0]] <= 0) rates[0] = 0.0;
      rates[0] = omin(rates[0], sv[iFrom[0]] / tstep);
      break;
    }
    case RVN_OP_OPEN_WATER_RIPARIAN: /* :137-141 */
    case RVN_OP_OPEN_WATER_EVAP: { /* src/OpenWaterEvap.cpp:113-186 */
      if (type == RVN_HRU_MASKED_GLACIER) break;
      if (hru_linked(d, c->k)) break; /* :124 reservoir-linked HRUs handle ET via the reservoir mass balance */
      double OWPET = F[RVN_F_OW_PET];
      if (!d->opt.suppress_competitive_ET) { OWPET -= (sv[c->iSV[RVN_SV_AET]] / tstep); OWPET = omax(OWPET, 0.0); }
      if (pr->op == RVN_OP_OPEN_WATER_RIPARIAN) rates[0] = P_LU(c, RVN_LU_OW_PET_CORR) * P_LU(c, RVN_LU_STREAM_FRACTION) * OWPET;
      else rates[0] = P_LU(c, RVN_LU_OW_PET_CORR) * OWPET;
      if (rates[0] < 0) rates[0] = 0.0;
      if (iFrom[0] != c->iSV[RVN_SV_SURFACE_WATER]) {
        rates[0] = omin(rates[0], sv[iFrom[0]] / tstep);
        if (sv[iFrom[0]] <= 0) rates[0] = 0.0;
      }
      rates[pr->nconn - 1] = rates[0];
      break;
    }
    case RVN_OP_LAKE_EVAP_BASIC: { /* src/OpenWaterEvap.cpp:269-314 */
      if (((type == RVN_HRU_LAKE) || (d->opt.lake_sv == iFrom[0])) && !hru_linked(d, c->k)) { /* :279 */
        rates[0] = P_LU(c, RVN_LU_LAKE_PET_CORR) * F[RVN_F_OW_PET];
        rates[pr->nconn - 1] = rates[0];
      }
      if (sv[iFrom[0]] <= 0) rates[0] = 0.0;
      if (rates[0] < 0) rates[0] = 0.0;
      rates[0] = omin(rates[0], sv[iFrom[0]] / tstep);
      rates[pr->nconn - 1] = rates[0];
      break;
    }
    case RVN_OP_CAPRISE_HBV: { /* src/CapillaryRise.cpp:103-179 */
      if (type == RVN_HRU_LAKE || type == RVN_HRU_WATER || type == RVN_HRU_ROCK) break;
      int m = d->sv_layer[iTo[0]];
      double stor = sv[iTo[0]], max_stor = soil_capacity(c, m);
      stor = omin(omax(stor, 0.0), max_stor);
      rates[0] = P_S(c, m, RVN_S_MAX_CAP_RISE_RATE) * (1.0 - omin(stor / max_stor, 1.0));
      rates[0] = omin(rates[0], omax(sv[iFrom[0]], 0.0) / tstep);
      if (!d->opt.allow_soil_overfill) rates[0] = omin(rates[0], (state_var_max(c, iTo[0], sv) - sv[iTo[0]]) / tstep);
      break;
    }
    case RVN_OP_GLACIERMELT_HBV: /* src/GlacierProcesses.cpp:117-190 */
      if (type != RVN_HRU_GLACIER) break;
      if (!(sv[c->iSV[RVN_SV_SNOW]] > O_REAL_SMALL)) rates[0] = P_LU(c, RVN_LU_HBV_MELT_GLACIER_CORR) * omax(F[RVN_F_POTENTIAL_MELT], 0.0);
      if (rates[0] < 0.0) rates[0] = 0.0;
      break;
    case RVN_OP_GLACIERMELT_SIMPLE: /* GMELT_SIMPLE_MELT, src/GlacierProcesses.cpp:126-129 + :170-186 */
      if (type != RVN_HRU_GLACIER) break;
      rates[0] = F[RVN_F_POTENTIAL_MELT];
      if (rates[0] < 0.0) rates[0] = 0.0;
      break;
    case RVN_OP_GLACIERMELT_UBC: { /* GMELT_UBC, :139-166 */
      if (type != RVN_HRU_GLACIER) break;
      double CC = sv[iFrom[1]], snow_coverage = snow_cover(c), pot_melt = F[RVN_F_POTENTIAL_MELT];
      if (pot_melt > 0.0) { pot_melt *= (1.0 - snow_coverage); } else { pot_melt = 0.0; }
      if (CC > pot_melt * tstep) { CC -= pot_melt * tstep; pot_melt = 0; }
      else { pot_melt -= CC / tstep; CC = 0.0; }
      double P0CTKG = 0.0;
      CC *= (1 - (1 / (1 + P0CTKG)));
      rates[0] = pot_melt;
      rates[1] = (CC - sv[iFrom[1]]) / tstep;
      if (rates[0] < 0.0) rates[0] = 0.0;
      break;
    }
    case RVN_OP_GLACIERRELEASE_LINEAR: { /* GRELEASE_LINEAR / _ANALYTIC, :317-327 + :340-350 */
      if (type != RVN_HRU_GLACIER) break;
      double glacier_stor = sv[c->iSV[RVN_SV_GLACIER]], K = P_LU(c, RVN_LU_GLAC_STORAGE_COEFF);
      if (pr->ia[0] == 0) rates[0] = K * glacier_stor;
      else rates[0] = glacier_stor * (1 - exp(-K * tstep)) / tstep;
      if (rates[0] < 0.0) rates[0] = 0.0;
      break;
    }
    case RVN_OP_GLACIERRELEASE_HBVEC: { /* src/GlacierProcesses.cpp:292-350 */
      if (type != RVN_HRU_GLACIER) break;
      double glacier_stor = sv[c->iSV[RVN_SV_GLACIER]], snow = sv[c->iSV[RVN_SV_SNOW]], liq_snow = sv[c->iSV[RVN_SV_SNOW_LIQ]];
      double Kmin = P_LU(c, RVN_LU_HBV_GLACIER_KMIN), Kmax = P_LU(c, RVN_LU_GLAC_STORAGE_COEFF), Ag = P_LU(c, RVN_LU_HBV_GLACIER_AG);
      double K = Kmin + (Kmax - Kmin) * exp(-Ag * (snow + liq_snow));
      rates[0] = K * glacier_stor;
      if (rates[0] < 0.0) rates[0] = 0.0;
      break;
    }
    default: break; /* unsupported ops are rejected in rvo_run before the time loop */
  }
}

/* CmvConvolution::GetRatesOfChange, src/Convolution.cpp:245-303, with the hoisted unit hydrograph */
static void convolution_rates(const octx *c, const rvn_process *pr, const double *sv, double *rates) {
  const rvn_model_desc *d = c->d;
  const int NST = RVN_CONV_STORES, i0 = pr->ia[2], iTS = pr->ia[3];
  const double tstep = d->opt.timestep;
  size_t slot = ((size_t)c->member * d->nLU + d->hru_lu[c->k]) * d->nConvProc + pr->ia[0];
  const int N = d->conv_n[slot];
  const double *UH = &d->conv_uh[slot * RVN_CONV_STORES];
  double S[RVN_CONV_STORES];
  double TS_old = sv[iTS], sum = 0.0;
  for (int i = 0; i < N; i++) { S[i] = sv[i0 + i]; sum += S[i]; }
  S[0] += (TS_old - sum);
  rates[0] += (TS_old - sum) / tstep;
  double sumrem = 1.0, orig;
  for (int i = 0; i < N; i++) {
    if (sumrem < O_REAL_SMALL) orig = 0.0; else orig = S[i] / sumrem;
    rates[i + 1] = UH[i] * orig / tstep;
    S[i] -= rates[i + 1] * tstep;
    sumrem -= UH[i];
  }
  for (int i = N - 2; i >= 0; i--) {
    double move = S[i];
    rates[NST + i + 1] = move / tstep;
    S[i] -= move; S[i + 1] += move;
  }
  double TS_new = 0;
  for (int i = 1; i < N; i++) TS_new += S[i];
  rates[2 * NST] = (TS_new - TS_old) / tstep;
}

static int op_supported(int op) { return op > RVN_OP_NOP && op < RVN_OP_COUNT; }

/* CVegetationClass::RecalculateCanopyParams, src/VegetationParams.cpp:85-135 (capacities only; the month interpolation
 * of relative height / LAI arrives hoisted in d->veg_season) */
static void canopy_capacities(octx *c, int step) {
  const rvn_model_desc *d = c->d;
  c->canopy_cap = c->canopy_snow_cap = 0.0;
  c->rain_icept_pct = c->snow_icept_pct = 0.0; /* PRECIP_ICEPT_NONE */
  int veg = d->hru_veg[c->k];
  double rel_ht = d->veg_season[((size_t)step * d->nVeg + veg) * 2 + 0], rel_LAI = d->veg_season[((size_t)step * d->nVeg + veg) * 2 + 1];
  double max_height = P_V(c, RVN_V_MAX_HEIGHT), sparseness = P_LU(c, RVN_LU_FOREST_SPARSENESS), max_LAI = P_V(c, RVN_V_MAX_LAI);
  double SAI_ht_ratio = P_V(c, RVN_V_SAI_HT_RATIO);
  double max_SAI = max_height * SAI_ht_ratio;
  double height = max_height * rel_ht;
  double LAI = (1.0 - sparseness) * max_LAI * rel_LAI;
  double SAI = (1.0 - sparseness) * (height * SAI_ht_ratio);
  c->veg_LAI = LAI; c->veg_SAI = SAI;
  c->day_length = d->day_length ? d->day_length[(size_t)d->et_row[step] * d->nHRU + c->k] : 0.0;
  if (d->opt.interception_factor == RVN_ICEPT_USER) { c->rain_icept_pct = P_V(c, RVN_V_RAIN_ICEPT_PCT); c->snow_icept_pct = P_V(c, RVN_V_SNOW_ICEPT_PCT); }
  else if (d->opt.interception_factor == RVN_ICEPT_LAI) {
    c->rain_icept_pct = omin(P_V(c, RVN_V_RAIN_ICEPT_FACT) * (LAI + SAI), 1.0);
    c->snow_icept_pct = omin(P_V(c, RVN_V_SNOW_ICEPT_FACT) * (LAI + SAI), 1.0);
  } else if (d->opt.interception_factor == RVN_ICEPT_EXPLAI) c->rain_icept_pct = c->snow_icept_pct = (1.0 - exp(-0.5 * (LAI + SAI)));
  if (c->iSV[RVN_SV_CANOPY] < 0 && c->iSV[RVN_SV_CANOPY_SNOW] < 0) return;
  if ((max_LAI + max_SAI) > 0) {
    c->canopy_cap = ((LAI + SAI) / (max_LAI + max_SAI)) * P_V(c, RVN_V_MAX_CAPACITY);
    c->canopy_snow_cap = ((LAI + SAI) / (max_LAI + max_SAI)) * P_V(c, RVN_V_MAX_SNOW_CAPACITY);
  }
}

/* ======================================================================== routing ============== */
typedef struct { double *qin, *qlat, *qout, *scal; } obasin; /* one member; layouts of rvn_gpu_upload_basin_state */

/* CSubBasin::RouteWater, src/SubBasin.cpp:2584-2863 */
/* CChannelXSect::GetArea (src/ChannelXSect.cpp:297-302) -> InterpolateCurve (src/CommonFunctions.cpp:1982-2004), linear scan for the interval */
static double channel_get_area(const rvn_model_desc *d, int p, double Q) {
  const int N = d->nChanPts;
  const double *xx = d->chan_Q + (size_t)d->sb_channel[p] * N, *y = d->chan_A + (size_t)d->sb_channel[p] * N;
  const double x = Q / d->sb_Qmult[p];
  if (x <= xx[0]) return y[0] + (y[1] - y[0]) / (xx[1] - xx[0]) * (x - xx[0]);
  else if (x >= xx[N - 1]) return y[N - 1] + (y[N - 1] - y[N - 2]) / (xx[N - 1] - xx[N - 2]) * (x - xx[N - 1]);
  int i = 0;
  while (!((x >= xx[i]) && (x < xx[i + 1]))) i++;
  if (fabs(xx[i + 1] - xx[i]) < O_REAL_SMALL) return (y[i] + y[i + 1]) / 2;
  return y[i] + (y[i + 1] - y[i]) / (xx[i + 1] - xx[i]) * (x - xx[i]);
}

/* ROUTE_DIFFUSIVE_VARY: routing hydrographs [nSB][max_nqin] and celerity histories [nSB][max_nqin + 1] of the member being run */
static double *g_route_hydro = NULL, *g_c_hist = NULL;
/* rvn_erfc, src/CommonFunctions.cpp:1487-1512 (rational approximation / asymptotic series, not libm's erfc) */
static double o_erfc(double x) {
  double tmp = fabs(x), fun, f1, tmp2, tmp3;
  if (tmp > 3.0) {
    f1 = (1.0 - 1.0 / (2.0 * tmp * tmp) + 3.0 / (4.0 * pow(tmp, 4)) - 5.0 / (6.0 * pow(tmp, 6)));
    fun = f1 * exp(-tmp * tmp) / (tmp * sqrt(O_PI));
  } else {
    tmp2 = 1.0 / (1.0 + (0.3275911 * tmp));
    tmp3 = 0.254829592 * tmp2 - (0.284496736 * tmp2 * tmp2) + (1.421413741 * pow(tmp2, 3)) - (1.453152027 * pow(tmp2, 4)) + (1.061405429 * pow(tmp2, 5));
    fun = tmp3 * exp(-tmp * tmp);
  }
  if (tmp == x) return fun;
  return (2 - fun);
}
/* TimeVaryingADRCumDist, src/CommonFunctions.cpp:1819-1836 */
static double o_tv_adr_cum_dist(double t, double L, const double *v, double D, double dt) {
  int i = (int)(floor((t + O_TIME_CORRECTION) / dt));
  double tstar = v[i] / v[0] * (t - (double)(i)*dt);
  for (int j = 0; j < i; j++) tstar += dt * v[j] / v[0];
  double F = 0;
  if (t <= 0) return 0.0;
  if (L < 500 * (D / v[0])) F = 0.5 * (exp((v[0] * L) / D) * o_erfc((L + v[0] * tstar) / sqrt(4.0 * D * tstar)));
  F += 0.5 * (o_erfc((L - v[0] * tstar) / sqrt(4.0 * D * tstar)));
  return F;
}
/* CChannelXSect::GetCelerity, src/ChannelXSect.cpp (flow -> wave celerity off the rating curve of basin p's channel) */
static double channel_get_celerity(const rvn_model_desc *d, int p, double Q) {
  const int N = d->nChanPts;
  const double *aQ = d->chan_Q + (size_t)d->sb_channel[p] * N, *aA = d->chan_A + (size_t)d->sb_channel[p] * N;
  const double Q_mult = d->sb_Qmult[p];
  if (Q < 0) return 0.0;
  int i = 0;
  while ((Q > Q_mult * aQ[i + 1]) && (i < (N - 2))) i++;
  return Q_mult * (aQ[i + 1] - aQ[i]) / (aA[i + 1] - aA[i]);
}
/* CSubBasin::UpdateRoutingHydro (src/SubBasin.cpp:2521-2572), ROUTE_DIFFUSIVE_VARY: called for every basin before the routing loop
 * (src/Solvers.cpp:552), i.e. _aQinHist[0] is still the inflow of the previous step.  c_hist [max_nqin + 1], rh [max_nqin] of basin p. */
static void update_routing_hydro(const rvn_model_desc *d, int p, const double *QinHist, double *c_hist, double *rh) {
  const int nQ = d->sb_nqin[p];
  const double tstep = d->opt.timestep, L = d->sb_reach_m[p];
  double Q = QinHist[0];
  double cc = channel_get_celerity(d, p, Q) * O_SEC_PER_DAY;
  for (int n = nQ; n > 0; n--) c_hist[n] = c_hist[n - 1];
  c_hist[0] = cc;
  double sum = 0, alpha = d->sb_diff_alpha[p];
  for (int n = 0; n < nQ; n++) {
    rh[n] = omax(o_tv_adr_cum_dist((n)*tstep, L, c_hist, alpha, tstep) - sum, 0.0);
    sum += rh[n];
  }
  for (int n = 0; n < nQ; n++) rh[n] /= sum;
  double travel_time = L / d->sb_c_ref[p];   /* [s] compared with a step in days, as the reference does */
  if (travel_time < tstep) {
    rh[0] = 1.0 - travel_time / tstep;
    rh[1] = travel_time / tstep;
    for (int n = 2; n < nQ; n++) rh[n] = 0.0;
  }
}
static void route_water(const rvn_model_desc *d, int p, const obasin *B, double *aQout_new) {
  const int nSeg = d->sb_nseg[p], nQin = d->sb_nqin[p], nQlat = d->sb_nqlat[p];
  const double tstep = d->opt.timestep;
  const double *QinHist = &B->qin[(size_t)p * d->max_nqin], *QlatHist = &B->qlat[(size_t)p * d->max_nqlat];
  double *aQout = &B->qout[(size_t)p * RVN_MAX_RIVER_SEGS];
  const double *UH = &d->sb_unit_hydro[(size_t)p * d->max_nqlat];
  const double *RH = &d->sb_route_hydro[(size_t)p * d->max_nqin];
  const double seg_fraction = 1.0 / (double)(nSeg);
  double Qlat_new = 0.0;
  for (int n = 0; n < nQlat; n++) Qlat_new += UH[n] * QlatHist[n];
  int method = d->opt.routing;
  if (d->sb_headwater[p]) method = RVN_ROUTE_NONE;
  if (method == RVN_ROUTE_MUSKINGUM || method == RVN_ROUTE_MUSKINGUM_CUNGE) {
    double K = d->sb_musk_K[p], X = d->sb_musk_X[p], c1, c2, c3, c4, denom, cunge, dt, Qin, Qin_new;
    double stored[RVN_MAX_RIVER_SEGS];
    dt = omin(K, tstep);
    for (int s = 0; s < nSeg; s++) stored[s] = aQout[s];
    for (double t = 0; t < tstep; t += dt) {
      if (dt > (tstep - t)) dt = tstep - t;
      cunge = (method == RVN_ROUTE_MUSKINGUM_CUNGE) ? 1 : 0;
      denom = (2 * K * (1.0 - X) + dt);
      c1 = (dt - 2 * K * (X)) / denom; c2 = (dt + 2 * K * (X)) / denom; c3 = (-dt + 2 * K * (1 - X)) / denom; c4 = (dt) / denom;
      Qin = QinHist[1] + ((t) / tstep) * (QinHist[0] - QinHist[1]);
      Qin_new = QinHist[1] + ((t + dt) / tstep) * (QinHist[0] - QinHist[1]);
      for (int s = 0; s < nSeg; s++) {
        aQout_new[s] = c1 * Qin_new + c2 * Qin + c3 * aQout[s] + cunge * c4 * (Qlat_new * seg_fraction);
        Qin = aQout[s]; Qin_new = aQout_new[s];
        aQout[s] = aQout_new[s];
      }
    }
    for (int s = 0; s < nSeg; s++) aQout[s] = stored[s];
  } else if (method == RVN_ROUTE_STORAGECOEFF) {
    double Qin_new = QinHist[0], Qin = QinHist[1], Qlocal = B->scal[(size_t)p * 8 + 2];
    for (int s = 0; s < nSeg; s++) {
      double ttime = d->sb_K_reach[p] * seg_fraction;
      double sc = omin(1.0 / (ttime / tstep + 0.5), 1.0);
      double c1 = sc / 2, c2 = sc / 2, c3 = (1.0 - sc), corr = 1.0;
      aQout_new[s] = c1 * Qin + c2 * Qin_new + c3 * (aQout[s] - corr * Qlocal);
      Qin = aQout[s]; Qin_new = aQout_new[s];
    }
  } else if (method == RVN_ROUTE_HYDROLOGIC) { /* src/SubBasin.cpp:2685-2741 */
    const double ROUTE_MAXITER = 20, ROUTE_TOLERANCE = 0.0001;
    double reach_length = d->sb_reach_m[p];
    double Qout_old = aQout[nSeg - 1] - B->scal[(size_t)p * 8 + 2]; /* _aQout[last] - _Qlocal */
    double Qin_new = QinHist[0], Qin_old = QinHist[1];
    double V_old = channel_get_area(d, p, Qout_old) * reach_length;
    int iter = 0;
    double change = 0;
    double gamma = V_old + (Qin_old + Qin_new - Qout_old) / 2.0 * (tstep * O_SEC_PER_DAY);
    double dQ = 0.1, f, dfdQ, Q_guess = Qout_old, relax = 1.0;
    do {
      f = (channel_get_area(d, p, Q_guess) * reach_length + (Q_guess) / 2.0 * (tstep * O_SEC_PER_DAY));
      dfdQ = (channel_get_area(d, p, Q_guess + dQ) * reach_length + (Q_guess + dQ) / 2.0 * (tstep * O_SEC_PER_DAY) - f) / dQ;
      change = -(f - gamma) / dfdQ;
      if (dfdQ == 0.0) change = 1e-7;
      Q_guess += relax * change;
      if (Q_guess < 0) { Q_guess = 0; change = 0.0; }
      iter++;
      if (iter > 3) relax = 0.9;
      if (iter > 10) relax = 0.7;
    } while ((iter < ROUTE_MAXITER) && (fabs(change) > ROUTE_TOLERANCE));
    for (int s = 0; s < nSeg; s++) aQout_new[s] = aQout[s];
    aQout_new[nSeg - 1] = Q_guess;
  } else if (method == RVN_ROUTE_PLUG_FLOW || method == RVN_ROUTE_DIFFUSIVE_WAVE || method == RVN_ROUTE_DIFFUSIVE_VARY) {
    if (method == RVN_ROUTE_DIFFUSIVE_VARY) RH = g_route_hydro + (size_t)p * d->max_nqin;   /* this step's hydrograph (update_routing_hydro) */
    aQout_new[nSeg - 1] = 0.0;
    for (int n = 0; n < nQin; n++) aQout_new[nSeg - 1] += RH[n] * QinHist[n];
  } else { /* ROUTE_NONE */
    for (int s = 0; s < nSeg; s++) aQout_new[s] = 0.0;
    aQout_new[nSeg - 1] = QinHist[0];
  }
  aQout_new[nSeg - 1] += Qlat_new;
}
/* ---- reservoirs: CReservoir (src/Reservoir.cpp) without control structures, DZTR, seepage, constraint series, stage assimilation */
static double interpolate_curve(double x, const double *xx, const double *y, int N, int extrapbottom) { /* src/CommonFunctions.cpp:1982-2004 */
  if (x <= xx[0]) {
    if (extrapbottom) return y[0] + (y[1] - y[0]) / (xx[1] - xx[0]) * (x - xx[0]);
    return y[0];
  } else if (x >= xx[N - 1]) return y[N - 1] + (y[N - 1] - y[N - 2]) / (xx[N - 1] - xx[N - 2]) * (x - xx[N - 1]);
  int i = 0;
  while (!((x >= xx[i]) && (x < xx[i + 1]))) i++;
  if (fabs(xx[i + 1] - xx[i]) < O_REAL_SMALL) return (y[i] + y[i + 1]) / 2;
  return y[i] + (y[i + 1] - y[i]) / (xx[i + 1] - xx[i]) * (x - xx[i]);
}
static const double *res_curve(const rvn_model_desc *d, int r, int c) { return d->res_curves + ((size_t)r * RVN_RES_CURVES + c) * d->nResPts; }
static double res_volume(const rvn_model_desc *d, int r, double ht) { return interpolate_curve(ht, res_curve(d, r, 0), res_curve(d, r, 4), d->res_np[r], 1); } /* :1871 */
static double res_area(const rvn_model_desc *d, int r, double ht) { return interpolate_curve(ht, res_curve(d, r, 0), res_curve(d, r, 3), d->res_np[r], 0); }   /* :1880 */
static double res_weir_outflow(const rvn_model_desc *d, int r, double ht) { /* :1890-1894, weir adjustment 0 */
  double underflow = interpolate_curve(ht, res_curve(d, r, 0), res_curve(d, r, 2), d->res_np[r], 0);
  return interpolate_curve(ht - 0.0, res_curve(d, r, 0), res_curve(d, r, 1), d->res_np[r], 0) + underflow;
}
/* CReservoir::RouteWater (:1532-1853); rs = {stage, stage_last, Qout, Qout_last, Precip, MB_losses, AET m3, -}; OW_PET / lake_PET_corr of the
 * linked HRU (ignored without one).  Returns the new stage, *res_outflow the outflow at the end of the step. */
static double reservoir_route_water(const rvn_model_desc *d, int r, const double *rs, double Qin_old, double Qin_new, double OW_PET, double lake_PET_corr,
                                    double demand, double *res_outflow) {
  const double RES_TOLERANCE = 0.0001; const int RES_MAXITER = 100;
  const double tstep = d->opt.timestep, _stage = rs[0], _Qout = rs[2];
  double stage_new = 0.0, Qmin = 0.0, Qmax = 1e99;
  double dh = 0.001, V_old = res_volume(d, r, _stage), A_old = res_area(d, r, _stage), h_guess = _stage;
  int iter = 0;
  double change = 0, f, dfdh, out, out2, ET = 0.0, precip = 0.0, ext_old = 0.0, ext_new = 0.0, seep_old = 0.0;
  const double _seepage_const = 0.0, _local_GW_head = 0.0;
  if (d->res_hru[r] >= 0) {
    ET = OW_PET / O_SEC_PER_DAY / O_MM_PER_METER;
    ET *= lake_PET_corr;
    precip = rs[4] / d->opt.timestep / O_SEC_PER_DAY;
  }
  ext_new = ext_old = 0.0;
  ext_old = demand; ext_new = demand; /* one demand series (or none: demand = 0) */
  double gamma = V_old + ((Qin_old + Qin_new) - _Qout + 2.0 * precip - ET * A_old - seep_old - (ext_old + ext_new)) / 2.0 * (tstep * O_SEC_PER_DAY);
  if (gamma < 0) { *res_outflow = 0.0; return d->res_min_stage[r]; }
  double relax = 1.0;
  do {
    out = out2 = 0.0;
    out = res_weir_outflow(d, r, h_guess);
    out2 = res_weir_outflow(d, r, h_guess + dh);
    out += ET * res_area(d, r, h_guess) + _seepage_const * (h_guess - _local_GW_head);
    out2 += ET * res_area(d, r, h_guess + dh) + _seepage_const * (h_guess + dh - _local_GW_head);
    f = (res_volume(d, r, h_guess) + out / 2.0 * (tstep * O_SEC_PER_DAY));
    dfdh = ((res_volume(d, r, h_guess + dh) + out2 / 2.0 * (tstep * O_SEC_PER_DAY)) - f) / dh;
    change = -(f - gamma) / dfdh;
    if (dfdh == 0) change = 1e-7;
    if (iter > 3) relax *= 0.98;
    h_guess += relax * change;
    iter++;
  } while ((iter < RES_MAXITER) && (fabs(change / relax) > RES_TOLERANCE));
  stage_new = h_guess;
  *res_outflow = res_weir_outflow(d, r, stage_new);
  if ((*res_outflow < Qmin) || (*res_outflow > Qmax)) { /* :1775-1803 */
    if (*res_outflow < Qmin) *res_outflow = Qmin;
    if (*res_outflow > Qmax) *res_outflow = Qmax;
    double A_guess = A_old, A_last, V_new, seep_guess = seep_old;
    do {
      V_new = V_old + ((Qin_old + Qin_new) - (_Qout + *res_outflow) + 2.0 * precip - ET * (A_old + A_guess) - (seep_old + seep_guess) - (ext_old + ext_new)) / 2.0 * (tstep * O_SEC_PER_DAY);
      if (V_new < 0) {
        V_new = 0;
        *res_outflow = -2.0 * (V_new - V_old) / (tstep * O_SEC_PER_DAY) + (-_Qout + 2.0 * precip + (Qin_old + Qin_new) - ET * (A_old + 0.0) - (ext_old + ext_new));
      }
      stage_new = interpolate_curve(V_new, res_curve(d, r, 4), res_curve(d, r, 0), d->res_np[r], 0);
      A_last = A_guess;
      A_guess = res_area(d, r, stage_new);
      seep_guess = _seepage_const * (stage_new - _local_GW_head);
    } while (fabs(1.0 - A_guess / A_last) > 0.00001);
  }
  return stage_new;
}
/* CReservoir::GetAET [mm/d] (:1283-1292) */
static double reservoir_aet(const rvn_model_desc *d, int r, const double *rs, double OW_PET, double lake_PET_corr) {
  if (d->res_hru[r] < 0) return 0.0;
  double Evap = OW_PET;
  Evap *= lake_PET_corr;
  return Evap * 0.5 * (res_area(d, r, rs[0]) + res_area(d, r, rs[1])) / (d->hru_area[d->res_hru[r]] * O_M2_PER_KM2);
}

/* CSubBasin::UpdateOutflows, src/SubBasin.cpp:2320-2421 (no irrigation/diversion; the reservoir part is in the caller) */
static void update_outflows(const rvn_model_desc *d, int p, const obasin *B, const double *aQo) {
  const int nSeg = d->sb_nseg[p], nQlat = d->sb_nqlat[p];
  double *aQout = &B->qout[(size_t)p * RVN_MAX_RIVER_SEGS], *sc = &B->scal[(size_t)p * 8];
  const double *QinHist = &B->qin[(size_t)p * d->max_nqin], *QlatHist = &B->qlat[(size_t)p * d->max_nqlat];
  const double *UH = &d->sb_unit_hydro[(size_t)p * d->max_nqlat];
  sc[0] = aQout[nSeg - 1]; /* _QoutLast */
  for (int s = 0; s < nSeg; s++) aQout[s] = aQo[s];
  double irrQ_act = omin(0.0, aQout[nSeg - 1]); aQout[nSeg - 1] -= irrQ_act;
  double divQ_act = omin(0.0, aQout[nSeg - 1]); aQout[nSeg - 1] -= divQ_act;
  double dt = d->opt.timestep * O_SEC_PER_DAY, dV = 0.0, Qlat_new = 0.0;
  for (int n = 0; n < nQlat; n++) Qlat_new += UH[n] * QlatHist[n];
  sc[3] = sc[2]; sc[2] = Qlat_new; /* _QlocLast=_Qlocal; _Qlocal=Qlat_new */
  dV += 0.5 * (QinHist[0] + QinHist[1]) * dt;
  dV -= 0.5 * (aQout[nSeg - 1] + sc[0]) * dt;
  dV += 0.5 * (Qlat_new + sc[1]) * dt;
  dV -= 0.5 * (irrQ_act + 0.0) * dt;
  dV -= 0.5 * (divQ_act + 0.0) * dt;
  sc[4] += dV;
  dV = 0.0;
  dV += QlatHist[0] * dt;
  dV -= 0.5 * (Qlat_new + sc[1]) * dt;
  sc[5] += dV;
  sc[1] = Qlat_new; /* _QlatLast */
}

/* reservoir state of the member being run [nRes][RVN_RES_STATE] (set by rvo_run_member_res; the oracle is single-threaded test infrastructure) */
static double *g_res_state = NULL;
/* ======================================================================== one member, n steps === */
/* routing prep (src/Solvers.cpp:495-511) and diagnostics of the write-back (:697-729) for one HRU's newest state vector */
static void finish_hru(const rvn_model_desc *d, int k, double *phinew, double *aRouted, int iSW, int iRO, int iTotalSWE, int iGlacierMB) {
  const int NS = d->NS;
  int p = d->hru_sb[k];
  double SWvol = (phinew[iSW] / O_MM_PER_METER) * (d->hru_area[k] * O_M2_PER_KM2);
  if (hru_linked(d, k)) { for (int r = 0; r < d->nRes; r++) if (d->res_hru[r] == k) g_res_state[(size_t)r * RVN_RES_STATE + 4] = SWvol; } /* SetPrecip, src/Solvers.cpp:502-503 */
  else aRouted[p] += SWvol;
  phinew[iRO] = phinew[iSW];
  phinew[iSW] = 0.0;
  if (iTotalSWE >= 0) {
    phinew[iTotalSWE] = 0.0;
    for (int i = 0; i < NS; i++) if (d->sv_type[i] == RVN_SV_SNOW || d->sv_type[i] == RVN_SV_SNOW_LIQ) phinew[iTotalSWE] += phinew[i];
  }
  if (iGlacierMB >= 0) {
    phinew[iGlacierMB] = 0.0;
    for (int i = 0; i < NS; i++) {
      int t = d->sv_type[i];
      if (t == RVN_SV_SNOW || t == RVN_SV_SNOW_LIQ || t == RVN_SV_FIRN || t == RVN_SV_GLACIER_ICE || t == RVN_SV_GLACIER) phinew[iGlacierMB] += phinew[i];
    }
  }
}

/* state   [nHRU][NS]   in/out
 * basin   qin[nSB][max_nqin] qlat[nSB][max_nqlat] qout[nSB][50] scal[nSB][8]  in/out
 * cumul   [2] {_CumulInput,_CumulOutput} in/out
 * optional records (NULL to skip): qout_rec[nsteps][nSB], wshed_rec[nsteps][4], state_rec[nsteps][nHRU][NS],
 *                                  forc_rec[nsteps][nHRU][RVN_F_COUNT]
 * returns 0, or -1 if the model uses an op the oracle does not implement. */
int rvo_run_member_res(const rvn_model_desc *d, int member, double *state, double *qin, double *qlat, double *qout, double *scal, double *cumul,
                       int step0, int nsteps, double *qout_rec, double *wshed_rec, double *state_rec, double *forc_rec, double *res_state);
/* the same for a model without reservoirs, or with reservoirs started from the descriptor's initial reservoir state */
/* test hook: cumulative per-connection fluxes of the following rvo_run_member* calls are added into buf[nConn][nHRU] (NULL switches it off) */
void rvo_set_flux_buffer(double *buf) { g_cum_flux = buf; }
int rvo_run_member(const rvn_model_desc *d, int member, double *state, double *qin, double *qlat, double *qout, double *scal, double *cumul,
                   int step0, int nsteps, double *qout_rec, double *wshed_rec, double *state_rec, double *forc_rec) {
  double *rs = NULL;
  if (d->nRes > 0) { rs = (double *)malloc(sizeof(double) * (size_t)d->nRes * RVN_RES_STATE); memcpy(rs, d->res_state0, sizeof(double) * (size_t)d->nRes * RVN_RES_STATE); }
  int rc = rvo_run_member_res(d, member, state, qin, qlat, qout, scal, cumul, step0, nsteps, qout_rec, wshed_rec, state_rec, forc_rec, rs);
  free(rs);
  return rc;
}
/* res_state [nRes][RVN_RES_STATE] in/out (NULL without reservoirs) */
int rvo_run_member_res(const rvn_model_desc *d, int member, double *state, double *qin, double *qlat, double *qout, double *scal, double *cumul,
                       int step0, int nsteps, double *qout_rec, double *wshed_rec, double *state_rec, double *forc_rec, double *res_state) {
  g_res_state = res_state;
  free(g_route_hydro); free(g_c_hist); g_route_hydro = g_c_hist = NULL;
  if (d->opt.routing == RVN_ROUTE_DIFFUSIVE_VARY) {
    /* the celerity history is not part of the model state the reference saves or the C ABI carries: every call starts from the initial
     * history (c_ref everywhere) and the dump hydrograph, like a fresh run of the reference (src/SubBasin.cpp:2138-2148) */
    if (step0 != 0) return -1;
    g_route_hydro = (double *)malloc(sizeof(double) * (size_t)d->nSB * d->max_nqin);
    g_c_hist = (double *)malloc(sizeof(double) * (size_t)d->nSB * (d->max_nqin + 1));
    memcpy(g_route_hydro, d->sb_route_hydro, sizeof(double) * (size_t)d->nSB * d->max_nqin);
    for (int p = 0; p < d->nSB; p++) for (int n = 0; n <= d->max_nqin; n++) g_c_hist[(size_t)p * (d->max_nqin + 1) + n] = d->sb_c_ref[p] * O_SEC_PER_DAY;
  }
  if (d->nRes > 0 && (!res_state || d->opt.sol_method != RVN_SOL_ORDERED_SERIES)) return -1;
  const int nH = d->nHRU, NS = d->NS, NB = d->nSB, NP = d->nProc;
  const double tstep = d->opt.timestep;
  for (int j = 0; j < NP; j++) if (!op_supported(d->proc[j].op)) return -1;
  if (d->opt.sol_method == RVN_SOL_EULER && d->nConvProc > 0) return -1; /* the reference's Euler branch mishandles convolution stores (:279-282) */
  if (d->opt.sol_method == RVN_SOL_ITERATED_HEUN) {   /* restated for plain process lists only */
    if (d->nConvProc > 0) return -1;
    for (int j = 0; j < d->nProc; j++) if (d->proc[j].group != 0) return -1;
  }
  octx c; memset(&c, 0, sizeof(c));
  c.d = d; c.member = member; c.ptab = d->ptab + (size_t)member * d->ptab_stride;
  for (int t = 0; t < RVN_SV_NTYPES; t++) c.iSV[t] = sv_index(d, t, -1);
  const int iSW = c.iSV[RVN_SV_SURFACE_WATER], iAET = c.iSV[RVN_SV_AET], iRO = c.iSV[RVN_SV_RUNOFF], iGW = c.iSV[RVN_SV_GROUNDWATER];
  const int iTotalSWE = c.iSV[RVN_SV_TOTAL_SWE], iGlacierMB = c.iSV[RVN_SV_GLACIER_MB];
  double *F = (double *)malloc(sizeof(double) * (size_t)nH * RVN_F_COUNT);
  g_precip5 = d->gauge_precip5 ? (double *)calloc((size_t)nH, sizeof(double)) : 0;
  double *daily = (double *)calloc((size_t)nH * 7, sizeof(double)); /* valid from the first step of a day on: runs must start at a day boundary */
  double *phinew = (double *)malloc(sizeof(double) * NS), *phiread = (double *)malloc(sizeof(double) * NS);
  double *aRouted = (double *)malloc(sizeof(double) * NB), *aQinnew = (double *)malloc(sizeof(double) * NB);
  double rates[O_MAXCONN]; int cf[O_MAXCONN], ct[O_MAXCONN];
  obasin B; B.qin = qin; B.qlat = qlat; B.qout = qout; B.scal = scal;
  double *DAQadjust = (double *)calloc((size_t)NB, sizeof(double)); /* _aDAQadjust: persists between steps */
  for (int s = 0; s < nsteps; s++) {
    const int step = step0 + s;
    /* ---- UpdateHRUForcingFunctions (all HRUs, on the stored state) */
    for (int k = 0; k < nH; k++) update_forcings(d, c.ptab, member, k, step, state + (size_t)k * NS, c.iSV, F + (size_t)k * RVN_F_COUNT, daily + (size_t)k * 7);
    if (forc_rec) memcpy(forc_rec + (size_t)s * nH * RVN_F_COUNT, F, sizeof(double) * (size_t)nH * RVN_F_COUNT);
    /* ---- MassEnergyBalance, ORDERED_SERIES (src/Solvers.cpp:155-251) */
    for (int p = 0; p < NB; p++) { aRouted[p] = 0.0; aQinnew[p] = 0.0; }
    for (int k = 0; k < nH; k++) {
      double *phi_old = state + (size_t)k * NS;
      memcpy(phinew, phi_old, sizeof(double) * NS);
      if (iAET >= 0) phinew[iAET] = 0.0;
      if (iRO >= 0) phinew[iRO] = 0.0;
      if (iGW >= 0) phinew[iGW] = 0.0;
      /* EULER (src/Solvers.cpp:256-297): every process sees aPhi, the start-of-step vector (with the same reboots), and the
       * changes accumulate in aPhinew; ORDERED_SERIES (:205-251): processes see the newest vector */
      const int euler = (d->opt.sol_method == RVN_SOL_EULER);
      if (euler) memcpy(phiread, phinew, sizeof(double) * NS);
      const double *phisee = euler ? phiread : phinew;
      if (d->hru_enabled[k] && d->opt.sol_method == RVN_SOL_ITERATED_HEUN) {
        /* ITERATED_HEUN (src/Solvers.cpp:304-397): rate1 from aPhi, rate2 from the previous iterate, the iterate rebuilt from aPhi with
         * the mean rate (water from ATMOS_PRECIP keeps rate1) until mean |change| <= convergence_crit or max_iterations */
        double *phiprev = (double *)malloc(sizeof(double) * NS), *phi0 = (double *)malloc(sizeof(double) * NS);
        double rate2[O_MAXCONN];
        c.k = k; c.step = step; c.F = F + (size_t)k * RVN_F_COUNT; c.phi_old = phi_old;
        canopy_capacities(&c, step);
        memcpy(phi0, phinew, sizeof(double) * NS);
        const int iAtm = c.iSV[RVN_SV_ATMOS_PRECIP];
        int iter = 0, converg = 0;
        do {
          iter++;
          for (int i = 0; i < NS; i++) { phiprev[i] = phinew[i]; phinew[i] = phi0[i]; }
          for (int j = 0; j < NP; j++) {
            const rvn_process *pr = &d->proc[j];
            if (!((d->apply_mask[k] >> j) & 1ull)) continue;
            const int nconn = pr->nconn;
            for (int q = 0; q < nconn; q++) { cf[q] = d->conn_from[pr->conn0 + q]; ct[q] = d->conn_to[pr->conn0 + q]; rates[q] = 0.0; rate2[q] = 0.0; }
            op_rates(&c, pr, cf, ct, phi0, rates);
            op_rates(&c, pr, cf, ct, phiprev, rate2);
            for (int q = 0; q < nconn; q++) {
              int typ = d->sv_type[cf[q]];
              double rg = 0.5 * (rates[q] + rate2[q]);
              if (cf[q] == iAtm) rg = rates[q];
              if (ct[q] != cf[q]) { phinew[cf[q]] -= rg * tstep; phinew[ct[q]] += rg * tstep; }
              else if (is_water_storage(typ) && (typ != RVN_SV_CONVOLUTION)) { phinew[ct[q]] += 0.0; }
              else phinew[ct[q]] += rg * tstep;
            }
          }
          double check = 0.0;
          for (int i = 0; i < NS; i++) check += (fabs(phinew[i] - phiprev[i]) / NS);
          converg = (check <= d->opt.convergence_crit) || (iter == d->opt.max_iterations);
        } while (!converg);
        free(phiprev); free(phi0);
        if (d->nLat > 0) { memcpy(phi_old, phinew, sizeof(double) * NS); continue; }
        finish_hru(d, k, phinew, aRouted, iSW, iRO, iTotalSWE, iGlacierMB);
        memcpy(phi_old, phinew, sizeof(double) * NS);
      } else
      if (d->hru_enabled[k]) {
        c.k = k; c.step = step; c.F = F + (size_t)k * RVN_F_COUNT; c.phi_old = phi_old;
        canopy_capacities(&c, step);
        for (int j = 0; j < NP; j++) {
          const rvn_process *pr = &d->proc[j];
          if (!((d->apply_mask[k] >> j) & 1ull)) continue; /* CModel::ApplyProcess gate, src/Model.cpp:3061 */
          int nconn = pr->nconn;
          if (pr->group != 0) {
            /* CProcessGroup::GetRatesOfChange (src/ProcessGroup.cpp:72-99): j is the group's FIRST member (the loop skips the others
             * below); loc_rates is zeroed once and shared by all members; group rates[q1+qq] = weight * loc_rates[qq]. */
            if (!(pr->group & RVN_GRP_FIRST)) continue;
            const int gn = pr->gnconn;
            double loc_rates[O_MAXCONN];
            for (int q = 0; q < gn; q++) { rates[q] = 0.0; loc_rates[q] = 0.0; }
            int q1 = 0, jj = j;
            for (;;) {
              const rvn_process *sp = &d->proc[jj];
              int scf[O_MAXCONN], sct[O_MAXCONN];
              for (int q = 0; q < sp->nconn; q++) { scf[q] = d->conn_from[sp->conn0 + q]; sct[q] = d->conn_to[sp->conn0 + q]; }
              op_rates(&c, sp, scf, sct, phisee, loc_rates);
              for (int qq = 0; qq < sp->nconn; qq++) rates[q1 + qq] = sp->weight * loc_rates[qq];
              q1 += sp->nconn;
              if (sp->group & RVN_GRP_LAST) break;
              jj++;
            }
            nconn = gn;
            for (int q = 0; q < nconn; q++) { cf[q] = d->conn_from[pr->gconn0 + q]; ct[q] = d->conn_to[pr->gconn0 + q]; }
          } else {
          for (int q = 0; q < nconn; q++) rates[q] = 0.0;
          if (pr->op == RVN_OP_CONVOLVE) {
            const int NST = RVN_CONV_STORES;
            cf[0] = ct[0] = pr->ia[2];
            for (int i = 0; i < NST; i++) {
              cf[i + 1] = pr->ia[2] + i; ct[i + 1] = pr->ia[1];
              if (i < NST - 1) { cf[NST + i + 1] = pr->ia[2] + i; ct[NST + i + 1] = pr->ia[2] + i + 1; }
            }
            cf[2 * NST] = ct[2 * NST] = pr->ia[3];
            convolution_rates(&c, pr, phisee, rates);
          } else {
            for (int q = 0; q < nconn; q++) { cf[q] = d->conn_from[pr->conn0 + q]; ct[q] = d->conn_to[pr->conn0 + q]; }
            op_rates(&c, pr, cf, ct, phisee, rates);
          }
          }
          for (int q = 0; q < nconn; q++) { /* src/Solvers.cpp:220-236 */
            int typ = d->sv_type[cf[q]];
            if (ct[q] != cf[q]) { phinew[cf[q]] -= rates[q] * tstep; phinew[ct[q]] += rates[q] * tstep; }
            else if (is_water_storage(typ) && (typ != RVN_SV_CONVOLUTION) && (typ != RVN_SV_CONV_STOR)) { rates[q] = 0.0; phinew[ct[q]] += 0.0; }
            else phinew[ct[q]] += rates[q] * tstep;
          }
          if (g_cum_flux && pr->op != RVN_OP_CONVOLVE) { /* pModel->IncrementBalance(qs, k, rates_of_change[q] * tstep), src/Solvers.cpp:234 */
            const int q0 = (pr->group != 0) ? pr->gconn0 : pr->conn0;
            for (int q = 0; q < nconn; q++) g_cum_flux[(size_t)(q0 + q) * nH + k] += rates[q] * tstep;
          }
        }
        if (d->nLat > 0) { memcpy(phi_old, phinew, sizeof(double) * NS); continue; } /* lateral exchange first: the row now holds aPhinew[k] */
        finish_hru(d, k, phinew, aRouted, iSW, iRO, iTotalSWE, iGlacierMB);
        memcpy(phi_old, phinew, sizeof(double) * NS);
      }
    }
    if (d->nLat > 0) {
      /* ---- lateral exchange processes on the post-vertical state of ALL HRUs (src/Solvers.cpp:404-443; CModel::ApplyLateralProcess,
       * src/Model.cpp:3126-3165; CmvLatFlush::GetLateralExchange, src/LatFlush.cpp:204-234).  state[k] is aPhinew[k] here. */
      for (int l = 0; l < d->nLat; l++) {
        const rvn_lateral *L = &d->lat[l];
        double *ex = (double *)malloc(sizeof(double) * (size_t)(L->nconn > 0 ? L->nconn : 1));
        if (L->kind == RVN_LAT_EQUILIBRATE) {
          /* CmvLatEquilibrate::GetLateralExchange (src/LatEquilibrate.cpp:168-212).  The reference walks all connections with one
           * pool sum[p] per subbasin; the list here is the same connections grouped by subbasin (chunks) in the same relative
           * order, and the p/plast/qstart bookkeeping, which depends on the list only, arrives as lat_flag. */
          double mix = omin(L->mixing_rate * tstep, 1.0);
          const int32_t *cp = d->lat_chunk_ptr + L->chunk0;
          for (int ch = 0; ch < L->nchunks; ch++) {
            double sum = 0.0;
            for (int q = cp[ch]; q < cp[ch + 1]; q++) {
              const int kF = d->lat_kfrom[L->conn0 + q], kT = d->lat_kto[L->conn0 + q], flag = d->lat_flag[L->conn0 + q];
              double stor = state[(size_t)kF * NS + L->sv_from], to_stor = state[(size_t)kT * NS + L->sv_to];
              double Afrom = d->hru_area[kF], Ato = d->hru_area[kT];
              if (flag & 1) sum += mix * omax(to_stor, 0.0) * Ato;
              if (flag & 2) { /* step 1: a fraction of the storage moves to the mixing unit */
                ex[q] = mix * omax(stor, 0.0) / tstep * Afrom;
                sum += ex[q] * tstep;
              } else {        /* step 2: the mixed water goes back by area */
                ex[q] = sum * d->lat_frac[L->conn0 + q] / tstep;
              }
            }
          }
        }
        else if (L->kind == RVN_LAT_REDISTRIBUTE) { /* CmvLatRedistribute::GetLateralExchange, src/SnowRedistribute.cpp:114-215 */
          const double TAN_80_DEGREES = tan(80.0 * O_PI / 180.0), RAD_TO_DEG = 180.0 / O_PI;
          const double SWE_TO_DEPTH_FACTOR = 1000.0 / 250.0;
          for (int q = 0; q < L->nconn; q++) {
            const int kF = d->lat_kfrom[L->conn0 + q];
            double weight = d->lat_frac[L->conn0 + q];
            double areaFrom = d->hru_area[kF], slope_rad = d->hru_slope[kF];
            double slope_deg = slope_rad * RAD_TO_DEG;
            double snowSWE = state[(size_t)kF * NS + L->sv_from], snowMoved = 0.0;
            if (L->method == RVN_REDIST_CONTINUOUS) {
              double snowTransport = omin(0.5, slope_deg / TAN_80_DEGREES) * omin(1.0, snowSWE / L->mixing_rate);
              snowMoved = snowTransport * snowSWE;
            } else {
              double snowDepth = snowSWE * SWE_TO_DEPTH_FACTOR;
              double slope_thres = slope_deg;
              if (slope_thres < 10.0) slope_thres = 10.0;
              double threshold_snow_height_m = 3178.4 * pow(slope_thres, -1.998);
              double threshold_snow_height = threshold_snow_height_m * 1000;
              if (slope_deg > 60.0) threshold_snow_height = 0.0;
              if (snowDepth > threshold_snow_height) {
                double excessSnowDepth = snowDepth - threshold_snow_height;
                snowMoved = excessSnowDepth / SWE_TO_DEPTH_FACTOR;
              }
            }
            ex[q] = (snowMoved * weight * areaFrom) / tstep;
          }
        }
        else for (int q = 0; q < L->nconn; q++) {
          const int kF = d->lat_kfrom[L->conn0 + q], kT = d->lat_kto[L->conn0 + q];
          double stor = state[(size_t)kF * NS + L->sv_from], to_stor = state[(size_t)kT * NS + L->sv_to];
          double Afrom = d->hru_area[kF], Ato = d->hru_area[kT];
          octx cc = c; cc.k = kT;
          double max_to_stor = state_var_max(&cc, L->sv_to, state + (size_t)kT * NS);
          double max_rate = omax(max_to_stor - to_stor, 0.0) / tstep * Ato;
          double mult = 1.0;
          ex[q] = omax(mult * stor, 0.0) / tstep * Afrom * d->lat_frac[L->conn0 + q];
          ex[q] = omin(ex[q], max_rate);
        }
        for (int q = 0; q < L->nconn; q++) {
          const int kF = d->lat_kfrom[L->conn0 + q], kT = d->lat_kto[L->conn0 + q];
          state[(size_t)kF * NS + L->sv_from] -= ex[q] / d->hru_area[kF] * tstep;
          state[(size_t)kT * NS + L->sv_to] += ex[q] / d->hru_area[kT] * tstep;
        }
        free(ex);
      }
      for (int k = 0; k < nH; k++) if (d->hru_enabled[k]) finish_hru(d, k, state + (size_t)k * NS, aRouted, iSW, iRO, iTotalSWE, iGlacierMB);
    }
    /* ---- user-specified inflows to the upstream end of the reach (src/Solvers.cpp:543): value at t + dt */
    for (int i = 0; i < d->nSpec; i++) aQinnew[d->spec_basin[i]] += d->spec_in[(size_t)step * d->nSpec + i];
    if (d->opt.routing == RVN_ROUTE_DIFFUSIVE_VARY) /* CSubBasin::UpdateSubBasin for every basin, before any of them is routed (src/Solvers.cpp:552) */
      for (int p = 0; p < NB; p++) update_routing_hydro(d, p, &qin[(size_t)p * d->max_nqin], g_c_hist + (size_t)p * (d->max_nqin + 1), g_route_hydro + (size_t)p * d->max_nqin);
    /* ---- routing, upstream to downstream (src/Solvers.cpp:565-615) */
    for (int pp = 0; pp < NB; pp++) {
      int p = d->sb_order[pp];
      double aQoutnew[RVN_MAX_RIVER_SEGS];
      double *QinHist = &qin[(size_t)p * d->max_nqin], *QlatHist = &qlat[(size_t)p * d->max_nqlat];
      for (int n = d->sb_nqin[p] - 1; n > 0; n--) QinHist[n] = QinHist[n - 1];   /* UpdateInflow, src/SubBasin.cpp:1531-1537 */
      QinHist[0] = aQinnew[p];
      for (int n = d->sb_nqlat[p] - 1; n > 0; n--) QlatHist[n] = QlatHist[n - 1]; /* UpdateLateralInflow :1547-1553 */
      QlatHist[0] = aRouted[p] / (tstep * O_SEC_PER_DAY);
      route_water(d, p, &B, aQoutnew);
      {
        double down_Q = 0.0;   /* GetDownstreamInflow(t) + GetTotalReturnFlow() (src/Solvers.cpp:586) */
        for (int i = 0; i < d->nSpec; i++) if (d->spec_basin[i] == p) down_Q = d->spec_down[(size_t)step * d->nSpec + i] + 0.0;
        aQoutnew[d->sb_nseg[p] - 1] += down_Q;
      }
      int r = -1;
      for (int i = 0; i < d->nRes; i++) if (d->res_basin[i] == p) r = i;
      double res_ht = 0.0, res_outflow = 0.0, owpet = 0.0, lpc = 0.0;
      if (r >= 0) { /* src/Solvers.cpp:589-596 */
        double *rs = res_state + (size_t)r * RVN_RES_STATE;
        const int kr = d->res_hru[r];
        if (kr >= 0) { owpet = F[(size_t)kr * RVN_F_COUNT + RVN_F_OW_PET]; lpc = c.ptab[d->lu_off + d->hru_lu[kr] * RVN_LU_COUNT + RVN_LU_LAKE_PET_CORR]; }
        double res_inflow_last = qout[(size_t)p * RVN_MAX_RIVER_SEGS + d->sb_nseg[p] - 1];
        double res_inflow = omax((aQoutnew[d->sb_nseg[p] - 1] - 0.0 - 0.0), 0.0);
        res_ht = reservoir_route_water(d, r, rs, res_inflow_last, res_inflow, owpet, lpc, d->res_extract[(size_t)step * d->nRes + r], &res_outflow);
      }
      update_outflows(d, p, &B, aQoutnew);
      if (r >= 0) { /* CReservoir::UpdateStage + UpdateMassBalance (src/Reservoir.cpp:1264-1323), called from UpdateOutflows */
        double *rs = res_state + (size_t)r * RVN_RES_STATE;
        rs[1] = rs[0]; rs[0] = res_ht; rs[3] = rs[2]; rs[2] = res_outflow;
        rs[5] = 0.0; rs[6] = 0.0;
        if (d->res_hru[r] >= 0) rs[6] = reservoir_aet(d, r, rs, owpet, lpc) * (d->hru_area[d->res_hru[r]] * O_M2_PER_KM2) / O_MM_PER_METER * tstep;
        rs[5] += rs[6];
        rs[5] += d->res_extract[(size_t)step * d->nRes + r] * O_SEC_PER_DAY * tstep;
      }
      if (d->assimilate_flow) { /* CModel::AssimilationOverride, src/Assimilate.cpp:67-118 (DA_ECCC) */
        const size_t ix = (size_t)step * NB + p;
        if (d->da_override[ix]) {
          double Qobs = d->da_obsQ[ix], Qobs2 = d->da_obsQ2[ix], alpha = d->assimilation_fact;
          double Qmod = qout[(size_t)p * RVN_MAX_RIVER_SEGS + d->sb_nseg[p] - 1];
          if ((Qmod > 1e-8) && (Qobs != -1.2345)) {
            if (Qobs2 == -1.2345) DAQadjust[p] = alpha * (Qobs - Qmod);
            else DAQadjust[p] = alpha * (0.5 * (Qobs2 + Qobs) - Qmod);
          } else DAQadjust[p] = 0.0;
          /* CSubBasin::AdjustAllFlows(adjust, overriding = true, assimsite = true), src/SubBasin.cpp:1572-1616 */
          double va = 0.0;
          for (int i = 0; i < d->sb_nseg[p]; i++) {
            double *q = &qout[(size_t)p * RVN_MAX_RIVER_SEGS + i];
            *q += DAQadjust[p]; if (*q < 0.0) *q = 0.0;
            va += DAQadjust[p] * tstep * O_SEC_PER_DAY;
          }
          if (va > 0.0) cumul[0] += va / (d->watershed_area * O_M2_PER_KM2) * O_MM_PER_METER;
          else cumul[1] -= va / (d->watershed_area * O_M2_PER_KM2) * O_MM_PER_METER;
        } else {
          cumul[1] -= 0.0 / (d->watershed_area * O_M2_PER_KM2) * O_MM_PER_METER;
        }
      }
      int pTo = d->sb_down[p];
      if (pTo >= 0) aQinnew[pTo] += (r >= 0) ? res_state[(size_t)r * RVN_RES_STATE + 2] : qout[(size_t)p * RVN_MAX_RIVER_SEGS + d->sb_nseg[p] - 1]; /* GetOutflowRate */
      if (r >= 0 && d->res_hru[r] >= 0 && iAET >= 0) /* :608-612 */
        state[(size_t)d->res_hru[r] * NS + iAET] = reservoir_aet(d, r, res_state + (size_t)r * RVN_RES_STATE, owpet, lpc);
    }
    if (d->assimilate_flow) { /* CModel::AssimilationBackPropagate, src/Assimilate.cpp:283-339 */
      for (int pp = NB - 1; pp >= 0; pp--) { /* downstream to upstream */
        int p = d->sb_order[pp], pdown = d->sb_down[p];
        if (!d->da_override[(size_t)step * NB + p]) {
          if (pdown >= 0) DAQadjust[p] = DAQadjust[pdown] * (d->sb_drain_area[p] / d->sb_drain_area[pdown]);
          else DAQadjust[p] = 0.0;
        }
      }
      for (int p = 0; p < NB; p++) {
        const size_t ix = (size_t)step * NB + p;
        if (!d->da_override[ix]) {
          if (d->sb_down[p] >= 0) DAQadjust[p] *= d->da_decay[ix];
          DAQadjust[p] = DAQadjust[p] * d->da_wt[ix];
        }
      }
      for (int p = 0; p < NB; p++) { /* AdjustAllFlows(adjust, overriding = false, assimsite = override) */
        const double adjust = DAQadjust[p];
        double *QinHist = &qin[(size_t)p * d->max_nqin];
        for (int n = 1; n < d->sb_nqin[p]; n++) {
          QinHist[n] += adjust * (d->sb_drain_area[p] - d->sb_area[p]) / d->sb_drain_area[p] * d->sb_da_corr[p];
          if (QinHist[n] < 0.0) QinHist[n] = 0.0;
        }
        if (!d->da_override[(size_t)step * NB + p])
          for (int i = 0; i < d->sb_nseg[p]; i++) { double *q = &qout[(size_t)p * RVN_MAX_RIVER_SEGS + i]; *q += adjust; if (*q < 0.0) *q = 0.0; }
      }
    }
    /* ---- IncrementCumulInput / IncrementCumOutflow (src/Model.cpp:2458-2539) */
    {
      double area = d->watershed_area * O_M2_PER_KM2, sum = 0.0;
      for (int k = 0; k < nH; k++) if (d->hru_enabled[k]) sum += F[(size_t)k * RVN_F_COUNT + RVN_F_PRECIP] * d->hru_area[k]; /* GetAveragePrecip :1048-1060 */
      double avg_precip = sum / d->watershed_area;
      cumul[0] += avg_precip * tstep;
      cumul[0] += 0.0 * tstep; /* recharge */
      for (int i = 0; i < d->nSpec; i++) cumul[0] += d->spec_cum[(size_t)step * d->nSpec + i]; /* GetIntegratedSpecInflow of the basins that have one (the others add 0) */
      for (int p = 0; p < NB; p++) {
        int pd = d->sb_down[p];
        int r = -1;
        for (int i = 0; i < d->nRes; i++) if (d->res_basin[i] == p) r = i;
        if (d->sb_level[p] == 0 || pd < 0) {
          double integ = 0.5 * (qout[(size_t)p * RVN_MAX_RIVER_SEGS + d->sb_nseg[p] - 1] + scal[(size_t)p * 8 + 0]) * (tstep * O_SEC_PER_DAY);
          if (r >= 0) integ = 0.5 * (omax(res_state[(size_t)r * RVN_RES_STATE + 2] + 0.0, 0.0) + omax(res_state[(size_t)r * RVN_RES_STATE + 3] + 0.0, 0.0)) * (tstep * O_SEC_PER_DAY); /* CReservoir::GetIntegratedOutflow */
          cumul[1] += integ / area * O_MM_PER_METER;
        }
        cumul[1] += ((r >= 0) ? res_state[(size_t)r * RVN_RES_STATE + 5] : 0.0) / area * O_MM_PER_METER; /* GetReservoirLosses */
        cumul[1] += 0.0 / area * O_MM_PER_METER; cumul[1] += 0.0 / area * O_MM_PER_METER;
      }
      if (wshed_rec) {
        double *w = wshed_rec + (size_t)s * 4;
        w[0] = cumul[0]; w[1] = cumul[1]; w[2] = avg_precip;
        /* total water storage [mm]: area-weighted water-storage SVs + channel + rivulet (StandardOutput.cpp:745-785) */
        double tot = 0.0;
        for (int i = 0; i < NS; i++) {
          if (!is_water_storage(d->sv_type[i]) || i == c.iSV[RVN_SV_ATMOS_PRECIP]) continue;
          double a = 0.0;
          for (int k = 0; k < nH; k++) if (d->hru_enabled[k]) a += state[(size_t)k * NS + i] * d->hru_area[k];
          tot += a / d->watershed_area;
        }
        double ch = 0, rv = 0;
        for (int p = 0; p < NB; p++) { ch += scal[(size_t)p * 8 + 4]; rv += scal[(size_t)p * 8 + 5]; }
        double rsv = 0.0;
        for (int r = 0; r < d->nRes; r++) rsv += res_volume(d, r, res_state[(size_t)r * RVN_RES_STATE + 0]); /* GetTotalReservoirStorage */
        tot += ch / (d->watershed_area * O_M2_PER_KM2) * O_MM_PER_METER + rv / (d->watershed_area * O_M2_PER_KM2) * O_MM_PER_METER + rsv / (d->watershed_area * O_M2_PER_KM2) * O_MM_PER_METER;
        w[3] = tot;
      }
    }
    if (qout_rec) {
      for (int p = 0; p < NB; p++) qout_rec[(size_t)s * NB + p] = qout[(size_t)p * RVN_MAX_RIVER_SEGS + d->sb_nseg[p] - 1];
      for (int r = 0; r < d->nRes; r++) qout_rec[(size_t)s * NB + d->res_basin[r]] = res_state[(size_t)r * RVN_RES_STATE + 2]; /* GetOutflowRate */
    }
    if (state_rec) memcpy(state_rec + (size_t)s * nH * NS, state, sizeof(double) * (size_t)nH * NS);
  }
  free(g_precip5); g_precip5 = 0;
  free(F); free(daily); free(phinew); free(phiread); free(aRouted); free(aQinnew); free(DAQadjust);
  return 0;
}

/* ---------------------------------------------------------------------------------------------------------------------
 * EnKF analysis support: x = pinv(P) b for nrhs right-hand sides with singular values below tol * max dropped -- restates the
 * reference's SVD() (src/Matrix.cpp:144-390, Numerical Recipes' svdcmp + svbksb: Householder bidiagonalisation, accumulation of
 * the right- and left-hand transformations, implicit-shift QR with at most 30 sweeps per singular value; the reference's
 * cancellation loop starts at l-1, kept).  P [n][n] row-major; B, X [n][nrhs].  Returns 0, or 1 if the QR iteration did not
 * converge.  Used by oracle/enkf_oracle.py; pinned on the reference's own analysis by tests/golden/enkf_*.
 * --------------------------------------------------------------------------------------------------------------------- */
static double o_sign(double a, double b) { return b >= 0.0 ? (a >= 0.0 ? a : -a) : (a >= 0.0 ? -a : a); }
static double o_pythag(double a, double b) {
  double absa = fabs(a), absb = fabs(b);
  if (absa > absb) return absa * sqrt(1.0 + (absb / absa) * (absb / absa));
  return (absb == 0.0 ? 0.0 : absb * sqrt(1.0 + (absa / absb) * (absa / absb)));
}
#define OA(i, j) A[(size_t)(i) * n + (j)]
#define OV(i, j) v[(size_t)(i) * n + (j)]
int rvo_svd_solve(const double *P, int n, double tol, const double *B, int nrhs, double *X) {
  double *A = (double *)malloc(sizeof(double) * (size_t)n * n), *v = (double *)calloc((size_t)n * n, sizeof(double));
  double *w = (double *)calloc((size_t)n, sizeof(double)), *rv1 = (double *)calloc((size_t)n, sizeof(double));
  int i, its, j, jj, k, l = 0, nm = 0, rc = 0;
  double anorm = 0.0, c, f, g = 0.0, h, s, scale = 0.0, x, y, z;
  memcpy(A, P, sizeof(double) * (size_t)n * n);
  for (i = 0; i < n; i++) {
    l = i + 2;
    rv1[i] = scale * g;
    g = s = scale = 0.0;
    for (k = i; k < n; k++) scale += fabs(OA(k, i));
    if (scale != 0.0) {
      for (k = i; k < n; k++) { OA(k, i) /= scale; s += OA(k, i) * OA(k, i); }
      f = OA(i, i);
      g = -o_sign(sqrt(s), f);
      h = f * g - s;
      OA(i, i) = f - g;
      for (j = l - 1; j < n; j++) {
        for (s = 0.0, k = i; k < n; k++) s += OA(k, i) * OA(k, j);
        f = s / h;
        for (k = i; k < n; k++) OA(k, j) += f * OA(k, i);
      }
      for (k = i; k < n; k++) OA(k, i) *= scale;
    }
    w[i] = scale * g;
    g = s = scale = 0.0;
    for (k = l - 1; k < n; k++) scale += fabs(OA(i, k));
    if (scale != 0.0) {
      for (k = l - 1; k < n; k++) { OA(i, k) /= scale; s += OA(i, k) * OA(i, k); }
      f = OA(i, l - 1);
      g = -o_sign(sqrt(s), f);
      h = f * g - s;
      OA(i, l - 1) = f - g;
      for (k = l - 1; k < n; k++) rv1[k] = OA(i, k) / h;
      for (j = l - 1; j < n; j++) {
        for (s = 0.0, k = l - 1; k < n; k++) s += OA(j, k) * OA(i, k);
        for (k = l - 1; k < n; k++) OA(j, k) += s * rv1[k];
      }
      for (k = l - 1; k < n; k++) OA(i, k) *= scale;
    }
    if (fabs(w[i]) + fabs(rv1[i]) > anorm) anorm = fabs(w[i]) + fabs(rv1[i]);
  }
  for (i = n - 1; i >= 0; i--) {
    if (i < n - 1) {
      if (g != 0.0) {
        for (j = l; j < n; j++) OV(j, i) = (OA(i, j) / OA(i, l)) / g;
        for (j = l; j < n; j++) {
          for (s = 0.0, k = l; k < n; k++) s += OA(i, k) * OV(k, j);
          for (k = l; k < n; k++) OV(k, j) += s * OV(k, i);
        }
      }
      for (j = l; j < n; j++) OV(i, j) = OV(j, i) = 0.0;
    }
    OV(i, i) = 1.0;
    g = rv1[i];
    l = i;
  }
  for (i = n - 1; i >= 0; i--) {
    l = i + 1;
    g = w[i];
    for (j = l; j < n; j++) OA(i, j) = 0.0;
    if (g != 0.0) {
      g = 1.0 / g;
      for (j = l; j < n; j++) {
        for (s = 0.0, k = l; k < n; k++) s += OA(k, i) * OA(k, j);
        f = (s / OA(i, i)) * g;
        for (k = i; k < n; k++) OA(k, j) += f * OA(k, i);
      }
      for (j = i; j < n; j++) OA(j, i) *= g;
    } else for (j = i; j < n; j++) OA(j, i) = 0.0;
    OA(i, i) += 1.0;
  }
  for (k = n - 1; k >= 0 && !rc; k--) {
    for (its = 0; its < 30; its++) {
      int flag = 1;
      for (l = k; l >= 0; l--) {
        nm = l - 1;
        if (fabs(rv1[l]) + anorm == anorm) { flag = 0; break; }
        if (fabs(w[nm]) + anorm == anorm) break;
      }
      if (flag) {
        c = 0.0; s = 1.0;
        for (i = l - 1; i < k + 1; i++) {
          f = s * rv1[i];
          rv1[i] *= c;
          if (fabs(f) + anorm == anorm) break;
          g = w[i];
          h = o_pythag(f, g);
          w[i] = h;
          h = 1.0 / h;
          c = g * h;
          s = -f * h;
          for (j = 0; j < n; j++) { y = OA(j, nm); z = OA(j, i); OA(j, nm) = y * c + z * s; OA(j, i) = z * c - y * s; }
        }
      }
      z = w[k];
      if (l == k) {
        if (z < 0.0) { w[k] = -z; for (j = 0; j < n; j++) OV(j, k) = -OV(j, k); }
        break;
      }
      if (its == 29) { rc = 1; break; }
      x = w[l]; nm = k - 1; y = w[nm]; g = rv1[nm]; h = rv1[k];
      f = ((y - z) * (y + z) + (g - h) * (g + h)) / (2.0 * h * y);
      g = o_pythag(f, 1.0);
      f = ((x - z) * (x + z) + h * ((y / (f + o_sign(g, f))) - h)) / x;
      c = s = 1.0;
      for (j = l; j <= nm; j++) {
        i = j + 1;
        g = rv1[i]; y = w[i]; h = s * g; g = c * g;
        z = o_pythag(f, h);
        rv1[j] = z;
        c = f / z; s = h / z;
        f = x * c + g * s; g = g * c - x * s; h = y * s; y *= c;
        for (jj = 0; jj < n; jj++) { x = OV(jj, j); z = OV(jj, i); OV(jj, j) = x * c + z * s; OV(jj, i) = z * c - x * s; }
        z = o_pythag(f, h);
        w[j] = z;
        if (z != 0.0) { z = 1.0 / z; c = f * z; s = h * z; }
        f = c * g + s * y; x = c * y - s * g;
        for (jj = 0; jj < n; jj++) { y = OA(jj, j); z = OA(jj, i); OA(jj, j) = y * c + z * s; OA(jj, i) = z * c - y * s; }
      }
      rv1[l] = 0.0; rv1[k] = f; w[k] = x;
    }
  }
  if (!rc) {
    double maxw = -O_ALMOST_INF;
    for (j = 0; j < n; j++) if (w[j] > maxw) maxw = w[j];
    for (j = 0; j < n; j++) if (fabs(w[j] / maxw) < tol) w[j] = 0.0;
    for (int r = 0; r < nrhs; r++) {
      for (j = 0; j < n; j++) {
        s = 0.0;
        if (w[j] != 0.0) { for (i = 0; i < n; i++) s += OA(i, j) * B[(size_t)i * nrhs + r]; s /= w[j]; }
        rv1[j] = s;
      }
      for (j = 0; j < n; j++) {
        for (s = 0.0, jj = 0; jj < n; jj++) s += OV(j, jj) * rv1[jj];
        X[(size_t)j * nrhs + r] = s;
      }
    }
  }
  free(A); free(v); free(w); free(rv1);
  return rc;
}
#undef OA
#undef OV
