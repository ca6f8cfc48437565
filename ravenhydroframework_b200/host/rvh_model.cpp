// rvh_model.cpp -- model container, time series/gauges, subbasins and initialisation of the host layer.
// Everything here runs once (or once per ensemble member): it prepares the tables the GPU consumes.
#include "rvh_model.h"
#include "rvh_process.h"
#include <fstream>

// ---------------------------------------------------------------------------------- field name tables
static const char *kSoilNames[RVN_S_COUNT] = {"POROSITY", "CAP_RATIO", "PERC_COEFF", "PET_CORRECTION", "BASEFLOW_COEFF", "BASEFLOW_N",
  "FIELD_CAPACITY", "SAT_WILT", "HBV_BETA", "MAX_CAP_RISE_RATE", "MAX_PERC_RATE", "GR4J_X2", "GR4J_X3", "VIC_B_EXP", "MAX_BASEFLOW_RATE", "MAX_INTERFLOW_RATE",
  "PERC_N", "SAC_PERC_ALPHA", "SAC_PERC_EXPON", "PERC_ASPEN", "BASEFLOW_THRESH", "BASEFLOW_COEFF2", "STORAGE_THRESHOLD", "VIC_ZMIN", "VIC_ZMAX",
  "VIC_ALPHA", "VIC_EVAP_GAMMA", "UBC_EVAP_SOIL_DEF", "UBC_INFIL_SOIL_DEF", "ALBEDO_WET", "ALBEDO_DRY",
  "HYDRAUL_COND", "WETTING_FRONT_PSI", "UNAVAIL_FRAC", "EXCHANGE_FLOW"};
static const char *kLUNames[RVN_LU_COUNT] = {"IMPERMEABLE_FRAC", "FOREST_COVERAGE", "FOREST_SPARSENESS", "MELT_FACTOR", "MIN_MELT_FACTOR",
  "MAX_MELT_FACTOR", "DD_MELT_TEMP", "DD_AGGRADATION", "REFREEZE_FACTOR", "DD_REFREEZE_TEMP", "REFREEZE_EXP", "HMETS_RUNOFF_COEFF",
  "GAMMA_SHAPE", "GAMMA_SCALE", "GAMMA_SHAPE2", "GAMMA_SCALE2", "GR4J_X4", "HBV_MELT_FOR_CORR", "HBV_MELT_ASP_CORR",
  "HBV_MELT_GLACIER_CORR", "HBV_GLACIER_KMIN", "GLAC_STORAGE_COEFF", "HBV_GLACIER_AG", "DEP_MAX", "LAKE_PET_CORR", "OW_PET_CORR",
  "PARTITION_COEFF", "CC_DECAY_COEFF", "MAX_SAT_AREA_FRAC", "PDM_B", "AET_COEFF", "MAX_DEP_AREA_FRAC", "PONDED_EXP", "AWBM_AREAFRAC1",
  "AWBM_AREAFRAC2", "HYMOD2_G", "HYMOD2_KMAX", "HYMOD2_EXP", "PET_LIN_COEFF", "RAIN_MELT_MULT", "RELHUM_CORR", "PET_VAP_COEFF",
  "MIN_WIND_SPEED", "MAX_WIND_SPEED", "ABST_PERCENT", "ROUGHNESS", "PRIESTLEYTAYLOR_COEFF", "SCS_CN", "SCS_IA_FRACTION", "PDMROF_B", "DEP_THRESHOLD", "DEP_N", "DEP_MAX_FLOW", "DEP_K", "DEP_CRESTRATIO", "DEP_SEEP_K", "STREAM_FRACTION"};
static const char *kVegNames[RVN_V_COUNT] = {"MAX_HEIGHT", "MAX_LAI", "SAI_HT_RATIO", "MAX_CAPACITY", "MAX_SNOW_CAPACITY", "RAIN_ICEPT_PCT",
  "SNOW_ICEPT_PCT", "PET_VEG_CORR", "RAIN_ICEPT_FACT", "SNOW_ICEPT_FACT", "ALBEDO", "SVF_EXTINCTION", "CHU_MATURITY", "VAR_CAPACITY", "VAR_SNOW_CAPACITY", "VAR_RAIN_ICEPT_PCT", "VAR_SNOW_ICEPT_PCT"};
static const char *kGlobalNames[RVN_G_COUNT] = {"SNOW_SWI", "SNOW_SWI_MIN", "SNOW_SWI_MAX", "SWI_REDUCT_COEFF", "ADIABATIC_LAPSE", "PRECIP_LAPSE",
  "RAINSNOW_TEMP", "RAINSNOW_DELTA", "AVG_ANNUAL_SNOW", "SNOW_TEMPERATURE", "HBVEC_LAPSE_UPPER", "HBVEC_LAPSE_RATE", "HBVEC_LAPSE_ELEV",
  "AIRSNOW_COEFF", "AVG_ANNUAL_RUNOFF", "MAX_SWE_SURFACE", "TOC_MULTIPLIER", "TIME_TO_PEAK_MULTIPLIER", "GAMMA_SHAPE_MULTIPLIER",
  "GAMMA_SCALE_MULTIPLIER", "MOHYSE_PET_COEFF", "SNOW_ROUGHNESS", "WINDVEL_ICEPT", "WINDVEL_SCALE",
  "UBC_GW_SPLIT", "UBC_FLASH_PONDING"};
static int FindName(const char *const *tab, int n, const std::string &name) {
  for (int i = 0; i < n; i++) if (name == tab[i]) return i;
  return -1;
}
int SoilFieldFromName(const std::string &n) { return FindName(kSoilNames, RVN_S_COUNT, n); }
int LandUseFieldFromName(const std::string &n) {
  if (n == "IMPERM") return RVN_LU_IMPERMEABLE_FRAC;
  if (n == "FOREST_COV") return RVN_LU_FOREST_COVERAGE;
  return FindName(kLUNames, RVN_LU_COUNT, n);
}
int VegFieldFromName(const std::string &n) {
  if (n == "MAX_HT") return RVN_V_MAX_HEIGHT;
  int f = FindName(kVegNames, RVN_V_COUNT, n);
  return (f >= RVN_V_VAR_CAPACITY) ? -1 : f;
}
int GlobalFieldFromName(const std::string &n) { return FindName(kGlobalNames, RVN_G_COUNT, n); }
const char *SoilFieldName(int f) { return kSoilNames[f]; }
const char *LandUseFieldName(int f) { return kLUNames[f]; }
const char *VegFieldName(int f) { return kVegNames[f]; }
const char *GlobalFieldName(int f) { return kGlobalNames[f]; }

// ---------------------------------------------------------------------------------- CTimeSeries
int CTimeSeries::GetTimeIndex(double t_loc) const {
  int n = (int)(rvh_max(floor(t_loc / _interval), 0.0));
  int last = (int)_aVal.size() - 1;
  return (n <= last) ? n : last;
}
double CTimeSeries::GetValue(double t) const { return _aVal[GetTimeIndex(t + _t_corr)]; }
double CTimeSeries::GetValueInterp(double t) const {
  const double t_loc = t + _t_corr;
  const int n = GetTimeIndex(t_loc), nP = (int)_aVal.size();
  if (n == nP - 1) return _aVal[n];
  return _aVal[n] + (t_loc - (double)(n)*_interval) / (_interval) * (_aVal[n + 1] - _aVal[n]);
}
double CTimeSeries::GetAvgValue(double t, double tstep) const {
  // pulse-type series only (all forcing series of the supported subset): reference src/TimeSeries.cpp:343-405
  double t_loc = t + _t_corr;
  int n1 = GetTimeIndex(t_loc), n2 = GetTimeIndex(t_loc + tstep);
  const int nP = (int)_aVal.size();
  if (t_loc < -TIME_CORRECTION) return RAV_BLANK_DATA;
  if (t_loc > nP * _interval) return RAV_BLANK_DATA;
  if (n1 == n2) return _aVal[n1];
  double sum = 0, blank = 0, inc;
  inc = ((double)(n1 + 1) * _interval - t_loc);
  if (_aVal[n1] == RAV_BLANK_DATA) blank += inc; else sum += _aVal[n1] * inc;
  for (int n = n1 + 1; n < n2; n++) { if (_aVal[n] == RAV_BLANK_DATA) blank += _interval; else sum += _aVal[n] * _interval; }
  inc = ((t_loc + tstep) - (double)(n2) * _interval);
  if (_aVal[n2] == RAV_BLANK_DATA) blank += inc; else sum += _aVal[n2] * inc;
  if (blank / tstep > 0.001) return RAV_BLANK_DATA;
  return sum / tstep;
}
void CTimeSeries::Initialize(double model_start_day, int model_start_year, double model_duration, double timestep, int calendar, bool is_observation) {
  _t_corr = -TimeDifference(model_start_day, model_start_year, _start_day, _start_year, calendar);
  double duration = (double)_aVal.size() * _interval;
  double lstart = _t_corr, lend = _t_corr + model_duration;
  if (is_observation) {  // observations need not cover the run; one extra sample for the last period-ending value
    int nSampVal = (int)(ceil((model_duration + timestep) / timestep - TIME_CORRECTION));
    _aSampVal.assign(nSampVal, 0.0);
    _sampInterval = timestep;
    double t = 0;
    for (int nn = 0; nn < nSampVal; nn++) { _aSampVal[nn] = GetAvgValue(t, timestep); t += timestep; }
    return;
  }
  ExitGracefullyIf(duration < lstart, "CTimeSeries::Initialize: time series forcing data not available for entire model simulation duration (" + _name + ")");
  ExitGracefullyIf(duration + TIME_CORRECTION < lend, "CTimeSeries::Initialize: time series forcing data not available at end of model simulation (" + _name + ")");
  ExitGracefullyIf((lstart < 0) || (_start_year > model_start_year), "CTimeSeries::Initialize: time series forcing data not available at beginning of model simulation (" + _name + ")");
  int nSampVal = (int)(ceil(model_duration / timestep - TIME_CORRECTION));
  _aSampVal.assign(nSampVal, 0.0);
  _sampInterval = timestep;
  double t = 0;
  for (int nn = 0; nn < nSampVal; nn++) { _aSampVal[nn] = GetAvgValue(t, timestep); t += timestep; }
}
double CTimeSeries::GetSampledValue(int nn) const {
  if (nn >= (int)_aSampVal.size()) return RAV_BLANK_DATA;
  return _aSampVal[nn];
}

// ---------------------------------------------------------------------------------- CGauge
CGauge::CGauge(const std::string &nm) : name(nm), latitude(0), longitude(0), elevation(0), rainfall_corr(1.0), snowfall_corr(1.0), temperature_corr(0.0) {
  for (int i = 0; i < F_NTYPES; i++) { _pTS[i] = NULL; _owned[i] = false; }
  for (int m = 0; m < 12; m++) aAveTemp[m] = aMinTemp[m] = aMaxTemp[m] = aAvePET[m] = NOT_SPECIFIED;
}
CGauge::~CGauge() { for (int i = 0; i < F_NTYPES; i++) if (_owned[i]) delete _pTS[i]; }
void CGauge::AddTimeSeries(CTimeSeries *pTS, forcing_type f) {
  if (_owned[f]) delete _pTS[f];
  _pTS[f] = pTS; _owned[f] = true;
}
double CGauge::GetForcingValue(forcing_type f, int nn) const { return _pTS[f] ? _pTS[f]->GetSampledValue(nn) : 0.0; }
double CGauge::GetAverageSnowFrac(int nn) const {
  double rain = _pTS[F_RAINFALL]->GetSampledValue(nn), snow = _pTS[F_SNOWFALL]->GetSampledValue(nn);
  if ((snow + rain) == 0) return 0.0;
  return snow / (snow + rain);
}
void CGauge::Initialize(const optStruct &O) {
  // temperature series (reference src/Gauge.cpp:78-101, 647-706); daily data on a daily-or-finer model step
  CTimeSeries *Tave = _pTS[F_TEMP_AVE], *Tmin = _pTS[F_TEMP_DAILY_MIN], *Tmax = _pTS[F_TEMP_DAILY_MAX];
  if (Tave || _pTS[F_TEMP_DAILY_AVE] || (Tmin && Tmax)) {
    if (Tave && Tave->IsDaily()) AddTimeSeries(new CTimeSeries("TEMP_DAILY_AVE", Tave->GetStartDay(), Tave->GetStartYear(), 1.0, Tave->Values()), F_TEMP_DAILY_AVE);
    if (Tave && !Tave->IsDaily()) {
      ExitGracefully("CGauge::Initialize: sub-daily temperature input is not supported by the B200 build yet");
    } else if (Tmin && Tmax) {
      Tmin->Initialize(O.julian_start_day, O.julian_start_year, O.duration, O.timestep, O.calendar);
      Tmax->Initialize(O.julian_start_day, O.julian_start_year, O.duration, O.timestep, O.calendar);
      if (!_pTS[F_TEMP_DAILY_AVE]) {
        int nVals = (int)ceil(O.duration);
        std::vector<double> aAvg(nVals);
        double t = 0.0;
        for (int n = 0; n < nVals; n++) { aAvg[n] = 0.5 * (Tmin->GetValue(t + 0.5) + Tmax->GetValue(t + 0.5)); t += 1.0; }
        AddTimeSeries(new CTimeSeries("TEMP_DAILY_AVE", O.julian_start_day, O.julian_start_year, 1.0, aAvg), F_TEMP_DAILY_AVE);
      }
      if (O.timestep < (1.0 - TIME_CORRECTION)) {   // GenerateAveSubdailyTempFromMinMax, sub-daily branch (src/Gauge.cpp:683-699)
        const int nVals = (int)ceil(O.duration / O.timestep);
        std::vector<double> aT(nVals);
        double t = 0.0;
        for (int n = 0; n < nVals; n++) {
          const double Tmax_ = Tmax->GetValue(t + O.timestep / 2.0), Tmin_ = Tmin->GetValue(t + O.timestep / 2.0);
          // CGauge::DailyTempCorrection (src/Gauge.cpp:468-471): -cos(2*PI*(t - PEAK_TEMP_HR/HR_PER_DAY)), PEAK_TEMP_HR = 3
          const double c1 = -cos(2.0 * RVH_PI * (t - 3 / HR_PER_DAY)), c2 = -cos(2.0 * RVH_PI * ((t + O.timestep) - 3 / HR_PER_DAY));
          aT[n] = 0.5 * (Tmax_ + Tmin_) + 0.5 * (Tmax_ - Tmin_) * 0.5 * (c1 + c2);
          t += O.timestep;
        }
        AddTimeSeries(new CTimeSeries("TEMP_AVE", O.julian_start_day, O.julian_start_year, O.timestep, aT), F_TEMP_AVE);
      } else {
        CTimeSeries *T = _pTS[F_TEMP_DAILY_AVE];
        AddTimeSeries(new CTimeSeries("TEMP_AVE", T->GetStartDay(), T->GetStartYear(), T->GetInterval(), T->Values()), F_TEMP_AVE);
      }
    } else if (_pTS[F_TEMP_DAILY_AVE]) {
      ExitGracefully("CGauge::Initialize: min/max generation from daily average temperature is not supported by the B200 build yet");
    }
  }
  bool hasPrecip = _pTS[F_PRECIP] != NULL, hasSnow = _pTS[F_SNOWFALL] != NULL, hasRain = _pTS[F_RAINFALL] != NULL;
  ExitGracefullyIf((hasPrecip || hasRain) && (!hasSnow) && (O.rainsnow == RVN_RAINSNOW_DATA),
                   "CGauge::Initialize: snow autogeneration is off at gauge with precip data, but no snow data has been supplied.");
  if (hasSnow && hasRain) {  // CTimeSeries::Sum
    const CTimeSeries *a = _pTS[F_RAINFALL], *b = _pTS[F_SNOWFALL];
    ExitGracefullyIf(a->Values().size() != b->Values().size() || a->GetInterval() != b->GetInterval() || a->GetStartDay() != b->GetStartDay() ||
                     a->GetStartYear() != b->GetStartYear(), "CTimeSeries::Sum: time series must have same start, interval and length");
    std::vector<double> v(a->Values().size());
    for (size_t i = 0; i < v.size(); i++) {
      double x = a->Values()[i], y = b->Values()[i];
      v[i] = (x == RAV_BLANK_DATA || y == RAV_BLANK_DATA) ? RAV_BLANK_DATA : x + y;
    }
    AddTimeSeries(new CTimeSeries("PRECIP", a->GetStartDay(), a->GetStartYear(), a->GetInterval(), v), F_PRECIP);
  }
  if (!hasRain && !hasSnow && hasPrecip) {
    const CTimeSeries *a = _pTS[F_PRECIP];
    AddTimeSeries(new CTimeSeries("RAINFALL", a->GetStartDay(), a->GetStartYear(), a->GetInterval(), a->Values()), F_RAINFALL);
  }
  if (!hasSnow && (hasRain || hasPrecip)) {
    const CTimeSeries *a = _pTS[F_PRECIP] ? _pTS[F_PRECIP] : _pTS[F_RAINFALL];
    std::vector<double> z(a->Values().size(), 0.0);
    AddTimeSeries(new CTimeSeries("SNOWFALL", a->GetStartDay(), a->GetStartYear(), a->GetInterval(), z), F_SNOWFALL);
  }
  for (int i = 0; i < F_NTYPES; i++) if (_pTS[i]) _pTS[i]->Initialize(O.julian_start_day, O.julian_start_year, O.duration, O.timestep, O.calendar);
  // if OW_PET requests data but only PET is provided, use PET (src/Gauge.cpp:189-195)
  if (_pTS[F_PET] && !_pTS[F_OW_PET] && O.ow_evaporation == RVN_PET_DATA) {
    const CTimeSeries *a = _pTS[F_PET];
    CTimeSeries *c = new CTimeSeries("OW_PET", a->GetStartDay(), a->GetStartYear(), a->GetInterval(), a->Values());
    c->Initialize(O.julian_start_day, O.julian_start_year, O.duration, O.timestep, O.calendar);
    AddTimeSeries(c, F_OW_PET);
  }
  ExitGracefullyIf((aAveTemp[0] == NOT_SPECIFIED) && (O.evaporation == RVN_PET_FROMMONTHLY),
                   "CGauge::Initialize: monthly temperatures for gauge not specified (using :MonthlyAveTemperature command) , but are needed");
  ExitGracefullyIf((aAvePET[0] == NOT_SPECIFIED) && (O.evaporation == RVN_PET_FROMMONTHLY),
                   "CGauge::Initialize: monthly PET values for gauge not specified (using :MonthlyAveEvaporation command), but are needed");
}

// ---------------------------------------------------------------------------------- channel
void CChannelXSect::GenerateTrapezoid(double bottom_w, double bottom_elev, double sidewall_ratio) {
  // 50-point rating curve of the reference trapezoid constructor (src/ChannelXSect.cpp:134-168),
  // including its area formula (bottom_w+sidewall_ratio)*h
  ExitGracefullyIf(bedslope <= 0.0, "CChannelXSect Constructor: channel profile bedslope must be greater than zero");
  ExitGracefullyIf(mannings_n <= 0.0, "CChannelXSect Constructor: Manning's n must be greater than zero");
  min_mannings = mannings_n;
  const int nP = 50;
  aQ.resize(nP); aStage.resize(nP); aTopWidth.resize(nP); aXArea.resize(nP); aPerim.resize(nP);
  double dh = 0.1, h;
  for (int i = 0; i < nP; i++) {
    h = i * dh;
    if (i == nP - 1) h = 20;
    aStage[i] = bottom_elev + h;
    aTopWidth[i] = bottom_w + 2.0 * sidewall_ratio * h;
    aXArea[i] = (bottom_w + sidewall_ratio) * h;
    aPerim[i] = bottom_w + 2 * sqrt(h * h * (1 + sidewall_ratio * sidewall_ratio));
    aQ[i] = sqrt(bedslope) * aXArea[i] * pow(aXArea[i] / aPerim[i], 2.0 / 3.0) / min_mannings;
  }
}
// :ChannelProfile: rating curves at 20 equally spaced stages between the lowest and highest survey point; per stage, every survey
// segment contributes its wetted area / perimeter / top width and its own Manning flow (reference CChannelXSect survey constructor +
// GenerateRatingCurvesFromProfile + GetPropsFromProfile, src/ChannelXSect.cpp:30-76, 430-527)
void CChannelXSect::GenerateFromProfile(const std::vector<double> &X, const std::vector<double> &Z, const std::vector<double> &N) {
  const int nS = (int)X.size();
  ExitGracefullyIf(nS < 2, "CChannelXSect Constructor: a channel profile needs at least two survey points");
  ExitGracefullyIf(bedslope < 0.0, "CChannelXSect Constructor: channel profile bedslope must be greater than zero");
  min_mannings = ALMOST_INF;
  for (int i = 0; i < nS; i++) {
    ExitGracefullyIf(N[i] < 0, "CChannelXSect::Constructor: invalid mannings n");
    if (N[i] < min_mannings) min_mannings = N[i];
  }
  mannings_n = min_mannings;
  double lo = Z[0], hi = Z[0];
  for (int i = 0; i < nS; i++) { if (Z[i] > hi) hi = Z[i]; if (Z[i] < lo) lo = Z[i]; }
  ExitGracefullyIf((hi - lo) < REAL_SMALL, "CChannelXSect::GenerateRatingCurvesFromProfile: profile survey points all have same elevation");
  const int nP = 20;
  aQ.assign(nP, 0.0); aStage.assign(nP, 0.0); aTopWidth.assign(nP, 0.0); aXArea.assign(nP, 0.0); aPerim.assign(nP, 0.0);
  const double dz = (hi - lo) / ((int)(nP)-1.0);
  int r = 0;
  for (double z = lo; z < hi + 0.5 * dz; z += dz) {   // the reference's floating-point stage loop
    ExitGracefullyIf(r >= nP, "CChannelXSect::GenerateRatingCurvesFromProfile: stage loop overran the rating table");
    double A = 0, P = 0, Q = 0, w = 0;
    for (int i = 0; i < nS - 1; i++) {
      const double zl = std::min(Z[i], Z[i + 1]), zu = std::max(Z[i], Z[i + 1]), dx = fabs(X[i + 1] - X[i]);
      double Ai, Pi, wi;
      if ((z <= zl) || (dx == 0)) { Ai = 0; Pi = 0; wi = 0; }            // dry segment
      else if (z >= zu) {                                                 // submerged segment (+ vertical banks next to it)
        wi = dx;
        Ai = ((z - zu) * dx + 0.5 * dx * (zu - zl));
        Pi = sqrt(dx * dx + (zu - zl) * (zu - zl));
        if ((i > 0) && (Z[i - 1] > Z[i]) && (X[i - 1] == X[i])) Pi += (z <= Z[i - 1]) ? (z - Z[i]) : (Z[i - 1] - Z[i]);
        if ((i < (nS - 2)) && (Z[i + 2] > Z[i + 1]) && (X[i + 2] == X[i + 1])) Pi += (z <= Z[i + 2]) ? (z - Z[i + 1]) : (Z[i + 2] - Z[i + 1]);
        if (i == 0) Pi += (z - Z[i]);
        if (i == nS - 2) Pi += (z - Z[i + 1]);
      } else {                                                            // partly wet segment
        double ddx = (z - zl) / (zu - zl) * dx;
        if ((zu - zl) < REAL_SMALL) ddx = 0.0;
        Ai = 0.5 * (z - zl) * ddx;
        Pi = sqrt(ddx * ddx + (z - zl) * (z - zl));
        wi = ddx;
      }
      A += Ai; P += Pi; w += wi;
      if (Ai > 0) Q += pow(bedslope, 0.5) * Ai * pow(Ai / Pi, 2.0 / 3.0) / N[i];
    }
    aStage[r] = z; aQ[r] = Q; aTopWidth[r] = w; aXArea[r] = A; aPerim[r] = P;
    r++;
  }
}
static void FlowCorrections(double bedslope, double min_n, double SB_slope, double SB_n, double &Q_mult) {
  Q_mult = 1.0;
  if (SB_slope != AUTO_COMPUTE) Q_mult = pow(SB_slope / bedslope, 0.5);
  if (SB_n != AUTO_COMPUTE) Q_mult *= (min_n / SB_n);
}
double CChannelXSect::GetCelerity(double Q, double SB_slope, double SB_n) const {
  double Q_mult; FlowCorrections(bedslope, min_mannings, SB_slope, SB_n, Q_mult);
  if (Q < 0) return 0.0;
  const int nP = (int)aQ.size();
  int i = 0; while ((Q > Q_mult * aQ[i + 1]) && (i < (nP - 2))) i++;
  return Q_mult * (aQ[i + 1] - aQ[i]) / (aXArea[i + 1] - aXArea[i]);
}
double CChannelXSect::GetTopWidth(double Qin, double SB_slope, double SB_n) const {
  double Q_mult; FlowCorrections(bedslope, min_mannings, SB_slope, SB_n, Q_mult);
  // InterpolateCurve(x,xx,y,N,extrapbottom=true), src/CommonFunctions.cpp:1982-2004
  const double x = Qin / Q_mult;
  const std::vector<double> &xx = aQ, &y = aTopWidth;
  const int N = (int)xx.size();
  if (x <= xx[0]) return y[0] + (y[1] - y[0]) / (xx[1] - xx[0]) * (x - xx[0]);
  if (x >= xx[N - 1]) return y[N - 1] + (y[N - 1] - y[N - 2]) / (xx[N - 1] - xx[N - 2]) * (x - xx[N - 1]);
  int i = 0; while ((x > xx[i + 1]) && (i < (N - 2))) i++;
  if (fabs(xx[i + 1] - xx[i]) < REAL_SMALL) return (y[i] + y[i + 1]) / 2;
  return y[i] + (y[i + 1] - y[i]) / (xx[i + 1] - xx[i]) * (x - xx[i]);
}

double CChannelXSect::GetFlowMultiplier(double SB_slope, double SB_n) const { double Q_mult; FlowCorrections(bedslope, min_mannings, SB_slope, SB_n, Q_mult); return Q_mult; }
double CChannelXSect::GetArea(double Qin, double SB_slope, double SB_n) const {
  const double x = Qin / GetFlowMultiplier(SB_slope, SB_n);
  const std::vector<double> &xx = aQ, &y = aXArea;
  const int N = (int)xx.size();
  if (x <= xx[0]) return y[0] + (y[1] - y[0]) / (xx[1] - xx[0]) * (x - xx[0]);
  if (x >= xx[N - 1]) return y[N - 1] + (y[N - 1] - y[N - 2]) / (xx[N - 1] - xx[N - 2]) * (x - xx[N - 1]);
  int i = 0; while (!((x >= xx[i]) && (x < xx[i + 1]))) i++;
  if (fabs(xx[i + 1] - xx[i]) < REAL_SMALL) return (y[i] + y[i + 1]) / 2;
  return y[i] + (y[i + 1] - y[i]) / (xx[i + 1] - xx[i]) * (x - xx[i]);
}

// ---------------------------------------------------------------------------------- reservoirs
// InterpolateCurve, src/CommonFunctions.cpp:1982-2004 (the interval is the one with xx[i] <= x < xx[i+1])
static double InterpolateCurve(double x, const double *xx, const double *y, int N, bool extrapbottom) {
  if (x <= xx[0]) {
    if (extrapbottom) return y[0] + (y[1] - y[0]) / (xx[1] - xx[0]) * (x - xx[0]);
    return y[0];
  }
  if (x >= xx[N - 1]) return y[N - 1] + (y[N - 1] - y[N - 2]) / (xx[N - 1] - xx[N - 2]) * (x - xx[N - 1]);
  int i = 0; while ((i < N - 1) && !((x >= xx[i]) && (x < xx[i + 1]))) i++;
  ExitGracefullyIf(i >= N - 1, "InterpolateCurve::mis-ordered list or infinite x");   // e.g. NaN from a diverged Newton iteration (src/CommonFunctions.cpp:1999)
  if (fabs(xx[i + 1] - xx[i]) < REAL_SMALL) return (y[i] + y[i + 1]) / 2;
  return y[i] + (y[i + 1] - y[i]) / (xx[i + 1] - xx[i]) * (x - xx[i]);
}
static const double RES_GRAVITY = 9.80616;   // src/RavenInclude.h:118
void CReservoir::SetPowerLaw(double a_V, double b_V, double a_Q, double b_Q, double a_A, double b_A, double crestht, double depth) {   // src/Reservoir.cpp:122-163
  const int Np = 100;
  _aStage.assign(Np, 0.0); _aQ.assign(Np, 0.0); _aQunder.assign(Np, 0.0); _aArea.assign(Np, 0.0); _aVolume.assign(Np, 0.0);
  double aA = a_A, bA = b_A;
  if (aA == AUTO_COMPUTE) aA = a_V / depth * b_V;
  if (bA == AUTO_COMPUTE) bA = b_V - 1.0;
  _crest_ht = crestht; _min_stage = _crest_ht - depth; _max_stage = _crest_ht + 6.0;
  const double dh = (_max_stage - _min_stage) / (double)(Np - 1);
  for (int i = 0; i < Np; i++) {
    double ht = _min_stage + (double)(i) * dh;
    if (((ht + dh) - _crest_ht) * (ht - _crest_ht) < 0.0) ht = _crest_ht;
    const double ht_above_bottom = ht - _min_stage;
    _aStage[i] = ht;
    _aQ[i] = a_Q * pow(std::max(ht - _crest_ht, 0.0), b_Q);
    _aArea[i] = aA * pow(ht_above_bottom / depth, bA);
    _aVolume[i] = a_V * pow(ht_above_bottom / depth, b_V);
    _aQunder[i] = 0.0;
  }
  _max_capacity = _aVolume[Np - 1];
}
void CReservoir::SetLookup(const std::vector<double> &ht, const std::vector<double> &Q, const std::vector<double> &Qund, const std::vector<double> &A,
                           const std::vector<double> &V) {   // src/Reservoir.cpp:173-226
  const int Np = (int)ht.size();
  ExitGracefullyIf(Np < 2, "CReservoir::constructor: must have more than 1 data point in stage relations");
  _min_stage = ALMOST_INF; _max_stage = -ALMOST_INF;
  _aStage = ht; _aQ = Q; _aArea = A; _aVolume = V; _aQunder = Qund.empty() ? std::vector<double>(Np, 0.0) : Qund;
  for (int i = 0; i < Np; i++) {
    if (_aStage[i] < _min_stage) _min_stage = _aStage[i];
    if (_aStage[i] > _max_stage) _max_stage = _aStage[i];
    if (_aQ[i] == 0.0) _crest_ht = _aStage[i];
    const std::string id = " [bad reservoir: " + _name + " " + std::to_string(_SBID) + "]";
    ExitGracefullyIf((i > 0) && ((_aStage[i] - _aStage[i - 1]) < 0), "CReservoir::constructor: stage relations must be specified in order of increasing stage." + id);
    ExitGracefullyIf((i > 0) && ((_aVolume[i] - _aVolume[i - 1]) <= -REAL_SMALL), "CReservoir::constructor: volume-stage relationships must be monotonically increasing for all stages." + id);
    ExitGracefullyIf((i > 0) && ((_aQ[i] - _aQ[i - 1]) < -REAL_SMALL), "CReservoir::constructor: stage-discharge relationships must be increasing or flat for all stages." + id);
    ExitGracefullyIf((i > 0) && ((_aQunder[i] - _aQunder[i - 1]) < -REAL_SMALL), "CReservoir::constructor: stage-discharge (underflow) relationships must be increasing or flat for all stages." + id);
  }
  _max_capacity = _aVolume[Np - 1];
}
void CReservoir::SetLake(double weircoeff, double crestw, double crestht, double A, double depth) {   // src/Reservoir.cpp:311-358
  _crest_width = crestw; _crest_ht = crestht; _min_stage = -depth + _crest_ht; _max_stage = 5.0 + _crest_ht;
  ExitGracefullyIf(depth <= 0, "CReservoir::Constructor (Lake): cannot have negative maximum lake depth");
  ExitGracefullyIf(A <= 0, "CReservoir::Constructor (Lake): cannot have negative or zero lake area");
  ExitGracefullyIf(crestw <= 0, "CReservoir::Constructor (Lake): cannot have negative or zero crest width");
  ExitGracefullyIf(weircoeff <= 0, "CReservoir::Constructor (Lake): cannot have negative or zero weir discharge coefficient");
  const int Np = 102;
  _aStage.assign(Np, 0.0); _aQ.assign(Np, 0.0); _aQunder.assign(Np, 0.0); _aArea.assign(Np, 0.0); _aVolume.assign(Np, 0.0);
  const double dh = (_max_stage - _crest_ht) / (Np - 2);
  _aStage[0] = _min_stage; _aQ[0] = 0.0; _aQunder[0] = 0.0; _aArea[0] = A; _aVolume[0] = 0.0;
  for (int i = 1; i < Np; i++) {
    _aStage[i] = _crest_ht + (i - 1) * dh;
    _aQ[i] = 2.0 / 3.0 * weircoeff * sqrt(2 * RES_GRAVITY) * crestw * pow((_aStage[i] - _crest_ht), 1.5);
    _aQunder[i] = 0.0;
    _aArea[i] = A;
    _aVolume[i] = A * (_aStage[i] - _min_stage);
  }
  _max_capacity = _aVolume[Np - 1];
}
void CReservoir::SetVolumeStageCurve(const std::vector<double> &a_ht, const std::vector<double> &a_V, double weircoeff, double crestwidth) {   // src/Reservoir.cpp:1068-1116
  const int nPoints = (int)a_ht.size(), Np = (int)_aStage.size();
  const std::string id = " [bad reservoir: " + _name + " " + std::to_string(_SBID) + "]";
  for (int i = 1; i < nPoints; i++)
    ExitGracefullyIf(((a_ht[i] - a_ht[i - 1]) <= REAL_SMALL) || ((a_V[i] - a_V[i - 1]) <= REAL_SMALL), "CReservoir::SetVolumeStageCurve: volume-stage relationships must be monotonically increasing for all stages." + id);
  ExitGracefullyIf(a_V[0] != 0.0, "CReservoir::SetVolumeStageCurve: volume-stage relationships must start with volume of zero." + id);
  _min_stage = a_ht[0]; _max_stage = a_ht[nPoints - 1];
  const double dh = (_max_stage - _min_stage) / (Np - 1);
  _aStage[0] = _min_stage; _aQ[0] = 0.0; _aQunder[0] = 0.0;
  _aArea[0] = (InterpolateCurve(_aStage[0] + dh, a_ht.data(), a_V.data(), nPoints, false)) / dh;
  _aVolume[0] = 0.0;
  for (int i = 1; i < Np; i++) {
    _aStage[i] = _min_stage + i * dh;
    _aQ[i] = 0.0;
    if ((_aStage[i] - _crest_ht) > 0.0) _aQ[i] = 2.0 / 3.0 * weircoeff * sqrt(2 * RES_GRAVITY) * crestwidth * pow((_aStage[i] - _crest_ht), 1.5);
    if (((_aStage[i] - _crest_ht) < dh) && ((_aStage[i] - _crest_ht) >= 0.0)) { _aStage[i] = _crest_ht; _aQ[i] = 0.0; }
    _aQunder[i] = 0.0;
    _aVolume[i] = InterpolateCurve(_aStage[i], a_ht.data(), a_V.data(), nPoints, false);
    _aArea[i] = (InterpolateCurve(_aStage[i] + dh, a_ht.data(), a_V.data(), nPoints, false) - _aVolume[i]) / dh;
  }
  _max_capacity = _aVolume[Np - 1];
}
void CReservoir::SetAreaStageCurve(const std::vector<double> &a_ht, const std::vector<double> &a_A) {   // src/Reservoir.cpp:1124-1134
  const int Np = (int)_aStage.size();
  for (int i = 0; i < Np; i++) {
    _aArea[i] = InterpolateCurve(_aStage[i], a_ht.data(), a_A.data(), (int)a_ht.size(), false);
    ExitGracefullyIf((i > 0) && ((_aArea[i] - _aArea[i - 1]) <= -REAL_SMALL) && (_aArea[i] != 0.0),
                     "CReservoir::SetAreaStageCurve: area-stage relationships must be monotonically increasing for all stages. [bad reservoir: " + _name + " " + std::to_string(_SBID) + "]");
  }
}
double CReservoir::GetVolume(double ht) const { return InterpolateCurve(ht, _aStage.data(), _aVolume.data(), (int)_aStage.size(), true); }
double CReservoir::GetArea(double ht) const { return InterpolateCurve(ht, _aStage.data(), _aArea.data(), (int)_aStage.size(), false); }
double CReservoir::GetWeirOutflow(double ht) const {   // src/Reservoir.cpp:1890-1894, weir adjustment 0
  const double underflow = InterpolateCurve(ht, _aStage.data(), _aQunder.data(), (int)_aStage.size(), false);
  return InterpolateCurve(ht - 0.0, _aStage.data(), _aQ.data(), (int)_aStage.size(), false) + underflow;
}
void CReservoir::SetInitialFlow(double initQ, double initQlast, bool justFlow) {   // src/Reservoir.cpp:1413-1485 (no control structures / DZTR)
  if ((initQ == 0.0) || justFlow) { _Qout = initQ; _Qout_last = initQlast; return; }
  const double RES_TOLERANCE = 0.001; const int RES_MAXITER = 20;
  const double dh = 0.0001;
  double h_guess = 0.1;
  const int Np = (int)_aStage.size();
  for (int i = 0; i < Np; i++)
    if (_aQ[i] > 0.0) { h_guess = _min_stage + (double)(i) / (double)(Np) * (_max_stage - _min_stage); break; }
  int iter = 0; double change = 0, Q, dQdh;
  do {
    Q = GetWeirOutflow(h_guess);
    dQdh = (GetWeirOutflow(h_guess + dh) - Q) / dh;
    change = -(Q - initQ) / dQdh;
    h_guess += change;
    iter++;
  } while ((iter < RES_MAXITER) && (fabs(change) > RES_TOLERANCE));
  if (iter == RES_MAXITER) WriteWarning("CReservoir::SetInitialFlow did not converge after " + std::to_string(RES_MAXITER) + "  iterations for basin " + std::to_string(_SBID));
  _stage = h_guess;
  _Qout = GetWeirOutflow(_stage);
  _stage_last = _stage; _Qout_last = _Qout;
}

// ---------------------------------------------------------------------------------- HRU / subbasin
// the reference converts with the rounded literal DEGREES_TO_RADIANS = 0.017453293 (src/RavenInclude.h:166; src/ParseHRUFile.cpp:327-328,
// src/HydroUnits.cpp:95), not with PI/180
static const double DEGREES_TO_RADIANS = 0.017453293;
double CHydroUnit::GetLatRad() const { return latitude * DEGREES_TO_RADIANS; }
double CHydroUnit::GetSlope() const { return slope * DEGREES_TO_RADIANS; }
double CHydroUnit::GetAspect() const { return aspect * DEGREES_TO_RADIANS; }

CSubBasin::CSubBasin()
    : ID(0), downstream_ID(DOESNT_EXIST), pChannel(NULL), reach_length(AUTO_COMPUTE), gauged(false), is_headwater(true), disabled(false), assimilate(false),
      nSegments(1), basin_area(0), drainage_area(0), t_conc(AUTO_COMPUTE), t_peak(AUTO_COMPUTE), t_lag(AUTO_COMPUTE),
      gamma_shape(AUTO_COMPUTE), gamma_scale(AUTO_COMPUTE), Q_ref(AUTO_COMPUTE), c_ref(AUTO_COMPUTE), w_ref(AUTO_COMPUTE),
      mannings_n(AUTO_COMPUTE), slope(AUTO_COMPUTE), rain_corr(1.0), snow_corr(1.0), temperature_corr(0.0), QoutLast(0), QlatLast(0),
      Qlocal(0), QlocLast(0), channel_storage(0), rivulet_storage(0), pInflowHydro(NULL), pInflowHydro2(NULL), pReservoir(NULL) { aQout.assign(1, AUTO_COMPUTE); diff_alpha = 0.0; }

bool CSubBasin::SetBasinProperties(const std::string &label_in, double value) {
  std::string l = StringToUppercase(label_in);
  if (l == "TIME_CONC") t_conc = value;
  else if (l == "TIME_TO_PEAK") t_peak = value;
  else if (l == "TIME_LAG") t_lag = value;
  else if (l == "GAMMA_SHAPE") gamma_shape = value;
  else if (l == "GAMMA_SCALE") gamma_scale = value;
  else if (l == "Q_REFERENCE") Q_ref = value;
  else if (l == "MANNINGS_N") mannings_n = value;
  else if (l == "SLOPE") slope = value;
  else if (l == "CELERITY") c_ref = value;
  else if (l == "RAIN_CORR") rain_corr = value;
  else if (l == "SNOW_CORR") snow_corr = value;
  else if (l == "TEMP_CORR") temperature_corr = value;
  else if (l == "REACH_LENGTH") reach_length = value * M_PER_KM;
  else return false;
  return true;
}
double CSubBasin::GetMuskingumK(double dx) const { return dx / (c_ref * SEC_PER_DAY); }
double CSubBasin::GetMuskingumX(double dx) const {
  ExitGracefullyIf(!pChannel, "CSubBasin::GetMuskingumX");
  double bedslope = slope;
  if (slope == AUTO_COMPUTE) bedslope = pChannel->bedslope;
  return rvh_max(0.0, 0.5 * (1.0 - Q_ref / bedslope / w_ref / c_ref / dx));
}
void CSubBasin::ResetReferenceFlow(double Qreference) {
  Q_ref = Qreference;
  if ((Q_ref != AUTO_COMPUTE) && (pChannel != NULL)) {
    ExitGracefullyIf((Q_ref <= 0.0) && (!is_headwater), "CSubBasin::ResetReferenceFlow: invalid (negative or zero) reference flow rate in non-headwater basin");
    if (c_ref == AUTO_COMPUTE) {
      c_ref = pChannel->GetCelerity(Q_ref, slope, mannings_n);
      ExitGracefullyIf(c_ref <= 0.0, "CSubBasin::ResetReferenceFlow: negative or zero celerity");
    }
    w_ref = pChannel->GetTopWidth(Q_ref, slope, mannings_n);
  } else {
    c_ref = AUTO_COMPUTE; w_ref = AUTO_COMPUTE;
  }
}
double CSubBasin::CalculateTOC() const {  // TOC_MCDERMOTT_PILGRIM (reference default, src/SubBasin.cpp:1742-1745)
  double A = rvh_max(basin_area, 0.001);
  return 0.76 * pow(A, 0.38) / HR_PER_DAY;
}
void CSubBasin::GenerateCatchmentHydrograph(double Qlat_avg, const optStruct &O) {
  const double tstep = O.timestep;
  int n, nQlat = 0;
  if (O.catchment_routing == ROUTE_TRI_CONVOLUTION) nQlat = (int)(ceil((t_conc) / tstep)) + 3;
  else if (O.catchment_routing == ROUTE_GAMMA_CONVOLUTION) nQlat = (int)(ceil(4.5 * pow(gamma_shape, 0.6) / gamma_scale / tstep)) + 1;
  else if (O.catchment_routing == ROUTE_DELAYED_FIRST_ORDER) nQlat = 2;
  else if (O.catchment_routing == ROUTE_DUMP) nQlat = 3;
  nQlat += (int)(ceil(t_lag / tstep));
  ExitGracefullyIf(nQlat <= 0, "GenerateCatchmentHydrograph: negative or zero inflow history");
  aQlatHist.assign(nQlat, Qlat_avg);
  aUnitHydro.assign(nQlat, 0.0);
  double sum, t;
  if (O.catchment_routing == ROUTE_GAMMA_CONVOLUTION) {
    sum = 0;
    for (n = 0; n < nQlat; n++) { t = n * tstep - t_lag; aUnitHydro[n] = GammaCumDist(t + tstep, gamma_shape, gamma_scale) - sum; sum += aUnitHydro[n]; }
    aUnitHydro[nQlat - 1] = 0.0;
  } else if (O.catchment_routing == ROUTE_TRI_CONVOLUTION) {
    sum = 0.0;
    for (n = 0; n < nQlat; n++) { t = n * tstep - t_lag; aUnitHydro[n] = TriCumDist(t + tstep, t_conc, t_peak) - sum; sum += aUnitHydro[n]; }
  } else if (O.catchment_routing == ROUTE_DUMP) {
    aUnitHydro[0] = 1.0;
  } else {
    ExitGracefully("GenerateCatchmentHydrograph: catchment routing method not implemented in the B200 build");
  }
  sum = 0.0;
  for (n = 0; n < nQlat; n++) sum += aUnitHydro[n];
  ExitGracefullyIf(sum == 0.0, "CSubBasin::GenerateCatchmentHydrograph: bad unit hydrograph constructed");
  for (n = 0; n < nQlat; n++) aUnitHydro[n] /= sum;
}
void CSubBasin::GenerateRoutingHydrograph(double Qin_avg, const optStruct &O) {
  const double tstep = O.timestep;
  double travel_time = reach_length / c_ref / SEC_PER_DAY;
  int nQin;
  if (O.routing == RVN_ROUTE_PLUG_FLOW) nQin = (int)(ceil(travel_time / tstep)) + 2;
  else if (O.routing == RVN_ROUTE_DIFFUSIVE_WAVE) nQin = (int)(ceil(2 * travel_time / tstep)) + 2;
  else if (O.routing == RVN_ROUTE_DIFFUSIVE_VARY) nQin = (int)(ceil(2.5 * travel_time / tstep)) + 2;   // src/SubBasin.cpp:2078-2081
  else if (O.routing == RVN_ROUTE_HYDROLOGIC) nQin = 2;
  else nQin = 20;
  aQinHist.assign(nQin, Qin_avg);
  aRouteHydro.assign(nQin, 0.0);
  if (O.routing == RVN_ROUTE_PLUG_FLOW) {
    for (int n = 0; n < nQin - 1; n++) {
      double ts = (travel_time - n * tstep) / tstep;
      if ((ts >= 0.0) && (ts < 1.0)) { aRouteHydro[n] = 1 - ts; aRouteHydro[n + 1] = ts; }
    }
  } else if (O.routing == RVN_ROUTE_DIFFUSIVE_WAVE) {   // src/SubBasin.cpp:2113-2133
    ExitGracefullyIf(nSegments > 1, "ROUTE_DIFFUSIVE_WAVE only valid for single-segment rivers");
    // CChannelXSect::GetDiffusivity at the reference flow (src/ChannelXSect.cpp:388-397; Roberson et al. 1995)
    ExitGracefullyIf(Q_ref <= 0, "CChannelXSect::GetDiffusivity: Invalid channel flowrate");
    const double slope_mult = (slope != AUTO_COMPUTE) ? (slope / pChannel->bedslope) : 1.0;
    const double diff_ref = 0.5 * (Q_ref) / pChannel->GetTopWidth(Q_ref, slope, mannings_n) / (slope_mult * pChannel->bedslope);
    const double cc = c_ref * SEC_PER_DAY, diffusivity = diff_ref * SEC_PER_DAY;   // [m/d], [m2/d]
    for (int n = 0; n < nQin; n++) aRouteHydro[n] = DiffusiveWaveUH(n, reach_length, cc, diffusivity, tstep);
  } else if (O.routing == RVN_ROUTE_DIFFUSIVE_VARY) {   // src/SubBasin.cpp:2138-2148: a dump hydrograph until the first UpdateRoutingHydro
    ExitGracefullyIf(nSegments > 1, "ROUTE_DIFFUSIVE_VARY only valid for single-segment rivers");
    ExitGracefullyIf(Q_ref <= 0, "CChannelXSect::GetDiffusivity: Invalid channel flowrate");
    aRouteHydro[0] = 1.0;
    const double slope_mult = (slope != AUTO_COMPUTE) ? (slope / pChannel->bedslope) : 1.0;
    diff_alpha = (0.5 * (Q_ref) / pChannel->GetTopWidth(Q_ref, slope, mannings_n) / (slope_mult * pChannel->bedslope)) * SEC_PER_DAY / 2;   // D_ref / 2 of UpdateRoutingHydro (:2539,2556)
  } else {
    aRouteHydro[0] = 1.0;
  }
  double sum = 0.0;
  for (int n = 0; n < nQin; n++) sum += aRouteHydro[n];
  ExitGracefullyIf(sum < 0.2, "CSubBasin::GenerateRoutingHydrograph: bad routing hydrograph constructed");
  for (int n = 0; n < nQin; n++) aRouteHydro[n] /= sum;
}
void CSubBasin::InitializeFlowStates(double Qin_avg, double Qlat_avg, const optStruct &O) {
  if (disabled) return;
  for (int seg = 0; seg < nSegments; seg++) aQout[seg] = Qin_avg + Qlat_avg * (double)(seg + 1) / (double)(nSegments);
  QoutLast = aQout[nSegments - 1];
  QlatLast = Qlat_avg;
  channel_storage = 0.0;
  if (O.routing == RVN_ROUTE_NONE) channel_storage = 0.0;
  else if (O.routing == RVN_ROUTE_HYDROLOGIC) {   // src/SubBasin.cpp:1983-1989
    for (int seg = 0; seg < nSegments; seg++) channel_storage += pChannel->GetArea(aQout[seg], slope, mannings_n) * (reach_length / nSegments);
  } else if (O.routing == RVN_ROUTE_PLUG_FLOW || O.routing == RVN_ROUTE_DIFFUSIVE_WAVE) {
    double sum = 0;
    for (int n = (int)aQinHist.size() - 1; n > 0; n--) { sum += aRouteHydro[n]; channel_storage += aQinHist[n] * sum * (O.timestep * SEC_PER_DAY); }
  } else if (O.routing == RVN_ROUTE_MUSKINGUM || O.routing == RVN_ROUTE_MUSKINGUM_CUNGE) {
    double dx = reach_length / (double)(nSegments);
    double K = GetMuskingumK(dx), X = GetMuskingumX(dx);
    channel_storage = K * (X * Qin_avg + (1.0 - X) * aQout[0]);
    for (int seg = 1; seg < nSegments; seg++) channel_storage += K * (X * aQout[seg - 1] + (1.0 - X) * aQout[seg]);
  }
  channel_storage += aQout[nSegments - 1] * (O.timestep * SEC_PER_DAY);
  channel_storage -= Qlat_avg * (O.timestep * SEC_PER_DAY);
  double sum = 0.0;
  for (int n = (int)aQlatHist.size() - 1; n > 0; n--) { sum += aUnitHydro[n]; rivulet_storage += sum * Qlat_avg * (O.timestep * SEC_PER_DAY); }
  rivulet_storage += Qlat_avg * O.timestep * (O.timestep * SEC_PER_DAY);
}
void CSubBasin::Initialize(double Qin_avg, double Qlat_avg, double total_drain_area, const CModel *pModel, const optStruct &O) {
  ExitGracefullyIf((pChannel == NULL) && (O.routing != RVN_ROUTE_NONE),
                   "CSubBasin::Initialize: channel profile for basin " + std::to_string(ID) + " may only be 'NONE' if Routing=ROUTE_NONE or ROUTE_EXTERNAL");
  if (pInflowHydro != NULL) is_headwater = false;   // src/SubBasin.cpp:1797
  drainage_area = total_drain_area;
  if (disabled) return;
  const global_struct *G = pModel->GetGlobalParams();
  if (Q_ref == AUTO_COMPUTE) {
    ExitGracefullyIf(((Qin_avg + Qlat_avg) <= 0) && (!is_headwater),
                     "CSubBasin::Initialize: negative or zero average flow specified in initialization (basin " + std::to_string(ID) + ")");
    ResetReferenceFlow(10.0 * (Qin_avg + Qlat_avg));  // REFERENCE_FLOW_MULT autogenerates to 10 (src/GlobalParams.cpp:159-163; src/SubBasin.cpp Initialize)
  } else {
    ResetReferenceFlow(Q_ref);
  }
  if (reach_length == AUTO_COMPUTE) reach_length = pow(basin_area, 0.67) * M_PER_KM;
  if (t_conc == AUTO_COMPUTE) { t_conc = CalculateTOC(); t_conc *= G->v[RVN_G_TOC_MULTIPLIER]; }
  if (O.catchment_routing == ROUTE_GAMMA_CONVOLUTION) t_peak = AUTO_COMPUTE;
  if (O.catchment_routing == ROUTE_DUMP) t_peak = AUTO_COMPUTE;
  if (t_peak == AUTO_COMPUTE) { t_peak = 0.3 * t_conc; t_peak *= G->v[RVN_G_TIME_TO_PEAK_MULTIPLIER]; }
  if (t_lag == AUTO_COMPUTE) t_lag = 0.0;
  if (gamma_shape == AUTO_COMPUTE) { gamma_shape = 3.0; gamma_shape *= G->v[RVN_G_GAMMA_SHAPE_MULTIPLIER]; }
  if (gamma_scale == AUTO_COMPUTE) { gamma_scale = rvh_max((gamma_shape - 1.0) / t_peak, 0.01); gamma_scale *= G->v[RVN_G_GAMMA_SCALE_MULTIPLIER]; }
  ExitGracefullyIf(t_conc < t_peak, "CSubBasin::Initialize: time of concentration must be greater than time to peak");
  ExitGracefullyIf(t_peak <= 0, "CSubBasin::Initialize: time to peak must be greater than zero");
  aQout.assign(nSegments, 0.0);
  GenerateCatchmentHydrograph(Qlat_avg, O);
  GenerateRoutingHydrograph(Qin_avg, O);
  InitializeFlowStates(Qin_avg, Qlat_avg, O);
  if (pInflowHydro) pInflowHydro->Initialize(O.julian_start_day, O.julian_start_year, O.duration, O.timestep, O.calendar);      // src/SubBasin.cpp:1899-1904
  if (pInflowHydro2) pInflowHydro2->Initialize(O.julian_start_day, O.julian_start_year, O.duration, O.timestep, O.calendar);
  if (pReservoir && pReservoir->_pDemandTS) pReservoir->_pDemandTS->Initialize(O.julian_start_day, O.julian_start_year, O.duration, O.timestep, O.calendar);   // CDemand::Initialize, src/WaterDemands.cpp:190-198
}
void CSubBasin::SetQout(double Q) {
  for (int i = 0; i < nSegments; i++) aQout[i] = Q;
  QoutLast = Q;
}

// ---------------------------------------------------------------------------------- CModel: catalogue
CModel::CModel(int nsoillayers) : num_members(1), _nSoilVars(nsoillayers), _lake_sv(0), _nConvVariables(0), _WatershedArea(0), _pOpt(NULL),
                                  _f_ptab_stride(0), _f_nconv(0) {
  ExitGracefullyIf(nsoillayers < 1, "CModel constructor::improper number of soil layers. SoilModel not specified?");
  calib_SBID = DOESNT_EXIST; calib_obj = 0; calib_period = "ALL";
  // constructor defaults of the reference (src/Model.cpp:104-117)
  const sv_type base[5] = {RVN_SV_SURFACE_WATER, RVN_SV_ATMOSPHERE, RVN_SV_ATMOS_PRECIP, RVN_SV_PONDED_WATER, RVN_SV_RUNOFF};
  for (int i = 0; i < 5; i++) { _aStateVarType.push_back(base[i]); _aStateVarLayer.push_back(i == 4 ? (int)RVN_SV_RUNOFF : DOESNT_EXIST); }
  for (int m = 0; m < nsoillayers; m++) { _aStateVarType.push_back(RVN_SV_SOIL); _aStateVarLayer.push_back(m); }
  memset(&_desc, 0, sizeof(_desc));
}
CModel::~CModel() {
  for (size_t i = 0; i < _pHydroUnits.size(); i++) delete _pHydroUnits[i];
  for (size_t i = 0; i < _pSubBasins.size(); i++) delete _pSubBasins[i];
  for (size_t i = 0; i < _pProcesses.size(); i++) delete _pProcesses[i];
  for (size_t i = 0; i < _pGauges.size(); i++) delete _pGauges[i];
  for (size_t i = 0; i < channels.size(); i++) delete channels[i];
  for (size_t i = 0; i < observed.size(); i++) delete observed[i].pTS;
}
int CModel::FindHRUGroup(const std::string &name) const { for (size_t i = 0; i < hru_groups.size(); i++) if (hru_groups[i].name == name) return (int)i; return DOESNT_EXIST; }
int CModel::FindSubBasinGroup(const std::string &name) const { for (size_t i = 0; i < sb_groups.size(); i++) if (sb_groups[i].name == name) return (int)i; return DOESNT_EXIST; }
void CModel::AddForcingPerturbation(forcing_type f, int distribution, const double distpar[3], int kk, adjustment adj) {
  force_perturb P; P.forcing = f; P.distribution = distribution; P.kk = kk; P.adj_type = adj;
  for (int i = 0; i < 3; i++) P.distpar[i] = distpar[i];
  perturbations.push_back(P);
}
void CModel::SetPerturbationSample(int member, int step, int i, double eps) {
  const int nP = (int)perturbations.size();
  ExitGracefullyIf(member < 0 || member >= _desc.nMembers || step < 0 || step >= _desc.nSteps || i < 0 || i >= nP || _f_pert_eps.empty(),
                   "CModel::SetPerturbationSample: index outside the flattened perturbation table (call Flatten first)");
  _f_pert_eps[((size_t)member * _desc.nSteps + step) * nP + i] = eps;
}
const double *CModel::MemberPerturbations(int member) const { return _f_pert_eps.data() + (size_t)member * _desc.nSteps * perturbations.size(); }
int CModel::GetStateVarIndex(sv_type typ) const { return GetStateVarIndex(typ, DOESNT_EXIST); }
int CModel::GetStateVarLayer(int ii) const {
  int count = 0;
  for (int i = 0; i < ii; i++) if (_aStateVarType[i] == _aStateVarType[ii]) count++;
  return count;
}
int CModel::GetStateVarIndex(sv_type typ, int layer) const {
  // reference _aStateVarIndices[type][layer] table: layer DOESNT_EXIST maps to slot 0 (src/Model.cpp:789-797)
  int want = (layer == DOESNT_EXIST) ? 0 : layer;
  for (size_t i = 0; i < _aStateVarType.size(); i++) {
    if (_aStateVarType[i] != typ) continue;
    int have = (_aStateVarLayer[i] == DOESNT_EXIST || typ == RVN_SV_RUNOFF) ? 0 : _aStateVarLayer[i];
    if (have == want) return (int)i;
  }
  return DOESNT_EXIST;
}
void CModel::AddStateVariables(const sv_type *aSV, const int *aLev, int nSV) {
  for (int ii = 0; ii < nSV; ii++) {
    bool found = false;
    for (size_t i = 0; i < _aStateVarType.size(); i++)
      if ((_aStateVarType[i] == aSV[ii]) && (_aStateVarLayer[i] == aLev[ii])) found = true;
    if (!found) { _aStateVarType.push_back(aSV[ii]); _aStateVarLayer.push_back(aLev[ii]); }
  }
}
bool CModel::IsWaterStorage(sv_type typ) {
  switch (typ) {
    case RVN_SV_SURFACE_WATER: case RVN_SV_PONDED_WATER: case RVN_SV_ATMOSPHERE: case RVN_SV_ATMOS_PRECIP: case RVN_SV_SNOW:
    case RVN_SV_GROUNDWATER: case RVN_SV_SOIL: case RVN_SV_CANOPY: case RVN_SV_CANOPY_SNOW: case RVN_SV_ROOT: case RVN_SV_TRUNK:
    case RVN_SV_DEPRESSION: case RVN_SV_SNOW_LIQ: case RVN_SV_WETLAND: case RVN_SV_GLACIER: case RVN_SV_GLACIER_ICE: case RVN_SV_FIRN:
    case RVN_SV_CONVOLUTION: case RVN_SV_NEW_SNOW: case RVN_SV_SNOW_DRIFT: case RVN_SV_LAKE_STORAGE:
      return true;
    default: return false;
  }
}
static const struct { const char *name; sv_type t; } kSVNames[] = {
  {"SURFACE_WATER", RVN_SV_SURFACE_WATER}, {"ATMOSPHERE", RVN_SV_ATMOSPHERE}, {"ATMOS_PRECIP", RVN_SV_ATMOS_PRECIP},
  {"PONDED_WATER", RVN_SV_PONDED_WATER}, {"SOIL", RVN_SV_SOIL}, {"CANOPY", RVN_SV_CANOPY}, {"CANOPY_SNOW", RVN_SV_CANOPY_SNOW},
  {"GROUNDWATER", RVN_SV_GROUNDWATER}, {"DEPRESSION", RVN_SV_DEPRESSION}, {"SNOW", RVN_SV_SNOW}, {"NEW_SNOW", RVN_SV_NEW_SNOW},
  {"SNOW_LIQ", RVN_SV_SNOW_LIQ}, {"TOTAL_SWE", RVN_SV_TOTAL_SWE}, {"WETLAND", RVN_SV_WETLAND}, {"GLACIER", RVN_SV_GLACIER},
  {"GLACIER_ICE", RVN_SV_GLACIER_ICE}, {"LAKE_STORAGE", RVN_SV_LAKE_STORAGE}, {"GLACIER_MB", RVN_SV_GLACIER_MB},
  {"CONVOLUTION", RVN_SV_CONVOLUTION}, {"CONV_STOR", RVN_SV_CONV_STOR}, {"CUM_INFIL", RVN_SV_CUM_INFIL}, {"CUM_SNOWMELT", RVN_SV_CUM_SNOWMELT},
  {"AET", RVN_SV_AET}, {"RUNOFF", RVN_SV_RUNOFF}, {"SNOW_TEMP", RVN_SV_SNOW_TEMP}, {"COLD_CONTENT", RVN_SV_COLD_CONTENT},
  {"SNOW_DEPTH", RVN_SV_SNOW_DEPTH}, {"SNOW_COVER", RVN_SV_SNOW_COVER}, {"SNOW_DEFICIT", RVN_SV_SNOW_DEFICIT}, {"SNOW_AGE", RVN_SV_SNOW_AGE},
  {"SNOW_ALBEDO", RVN_SV_SNOW_ALBEDO}, {"CROP_HEAT_UNITS", RVN_SV_CROP_HEAT_UNITS}, {"GLACIER_CC", RVN_SV_GLACIER_CC}, {"ICE_THICKNESS", RVN_SV_ICE_THICKNESS}, {NULL, RVN_SV_NTYPES}};
sv_type CModel::StringToSVType(const std::string &sin, int &layer, bool strict) const {
  std::string s = sin;
  std::map<std::string, std::string>::const_iterator it = _aliases.find(s);
  if (it != _aliases.end()) s = it->second;
  layer = DOESNT_EXIST;
  size_t b = s.find('[');
  std::string base = s;
  if (b != std::string::npos) {
    size_t e = s.find(']', b);
    ExitGracefullyIf(e == std::string::npos, "StringToSVType: malformed state variable name " + sin);
    layer = s_to_i(s.substr(b + 1, e - b - 1));
    base = s.substr(0, b);
  }
  for (int i = 0; kSVNames[i].name; i++) if (base == kSVNames[i].name) return kSVNames[i].t;
  ExitGracefullyIf(strict, "StringToSVType: unrecognized state variable name " + sin);
  return RVN_SV_NTYPES;
}
int CModel::ParseSVTypeIndex(const std::string &s) {
  int layer; sv_type t = StringToSVType(s, layer, true);
  AddStateVariables(&t, &layer, 1);
  int i = GetStateVarIndex(t, layer);
  ExitGracefullyIf(i == DOESNT_EXIST, "ParseSVTypeIndex: invalid state variable " + s);
  return i;
}
std::string CModel::GetStateVarName(int i) const {
  sv_type t = _aStateVarType[i];
  std::string nm = "SV" + std::to_string((int)t);
  for (int q = 0; kSVNames[q].name; q++) if (kSVNames[q].t == t) nm = kSVNames[q].name;
  if (t == RVN_SV_SOIL || t == RVN_SV_CONVOLUTION || t == RVN_SV_CONV_STOR) nm += "[" + std::to_string(_aStateVarLayer[i]) + "]";
  return nm;
}
int CModel::FindSoilClass(const std::string &t) const { for (size_t i = 0; i < soil_classes.size(); i++) if (soil_classes[i].tag == t) return (int)i; return -1; }
int CModel::FindVegClass(const std::string &t) const { for (size_t i = 0; i < veg_classes.size(); i++) if (veg_classes[i].tag == t) return (int)i; return -1; }
int CModel::FindLUClass(const std::string &t) const { for (size_t i = 0; i < lu_classes.size(); i++) if (lu_classes[i].tag == t) return (int)i; return -1; }
int CModel::FindTerrainClass(const std::string &t) const { for (size_t i = 0; i < terrain_classes.size(); i++) if (terrain_classes[i].tag == t) return (int)i; return -1; }
int CModel::FindSoilProfile(const std::string &t) const { for (size_t i = 0; i < soil_profiles.size(); i++) if (soil_profiles[i].tag == t) return (int)i; return -1; }
const CChannelXSect *CModel::FindChannel(const std::string &t) const {   // case-insensitive like CModel::StringToChannelXSect (src/Model.cpp:2256)
  const std::string up = StringToUppercase(t);
  for (size_t i = 0; i < channels.size(); i++) if (StringToUppercase(channels[i]->name) == up) return channels[i];
  return NULL;
}
int CModel::GetSubBasinIndex(long long SBID) const {
  if (SBID < 0) return DOESNT_EXIST;
  for (size_t p = 0; p < _pSubBasins.size(); p++) if (_pSubBasins[p]->ID == SBID) return (int)p;
  return -2;  // INDEX_NOT_FOUND
}
void CModel::UpdateParameter(int ctype, const std::string &pname, const std::string &cname, double value) {
  // sets the raw field only; derived fields (e.g. cap_ratio) are NOT refreshed, like the reference
  if (ctype == CLASS_SOIL) {
    int f = SoilFieldFromName(pname), c = FindSoilClass(cname);
    ExitGracefullyIf(f < 0 || c < 0, "CModel::UpdateParameter: unknown soil parameter/class " + pname + "/" + cname);
    soil_classes[c].S.v[f] = value;
  } else if (ctype == CLASS_VEGETATION) {
    int f = VegFieldFromName(pname), c = FindVegClass(cname);
    ExitGracefullyIf(f < 0 || c < 0, "CModel::UpdateParameter: unknown vegetation parameter/class " + pname + "/" + cname);
    veg_classes[c].V.v[f] = value;
  } else if (ctype == CLASS_LANDUSE) {
    int f = LandUseFieldFromName(pname), c = FindLUClass(cname);
    ExitGracefullyIf(f < 0 || c < 0, "CModel::UpdateParameter: unknown land use parameter/class " + pname + "/" + cname);
    lu_classes[c].S.v[f] = value;
  } else if (ctype == CLASS_GLOBAL) {
    int f = GlobalFieldFromName(pname);
    ExitGracefullyIf(f < 0, "CModel::UpdateParameter: unknown global parameter " + pname);
    globals.v[f] = value;
  } else ExitGracefully("CModel::UpdateParameter: unsupported class type");
}
void CModel::AllocateInitialState() { initial_state.assign((size_t)GetNumHRUs() * GetNumStateVars(), 0.0); }

// ---------------------------------------------------------------------------------- CModel: initialisation
void CModel::InitializeRoutingNetwork() {
  const int NB = GetNumSubBasins();
  _aSubBasinOrder.assign(NB, 0); _aDownstreamInds.assign(NB, DOESNT_EXIST); _aOrderedSBind.assign(NB, 0);
  std::vector<int> inflow(NB, 0);
  for (int p = 0; p < NB; p++) {
    int pp = GetSubBasinIndex(_pSubBasins[p]->downstream_ID);
    ExitGracefullyIf(pp == -2, "CModel::InitializeRoutingNetwork: downstream basin ID not found");
    ExitGracefullyIf(pp == p, "CModel::InitializeRoutingNetwork: subbasin empties into itself: circular reference!");
    _aDownstreamInds[p] = pp;
  }
  for (int p = 0; p < NB; p++) if (_aDownstreamInds[p] != DOESNT_EXIST) inflow[_aDownstreamInds[p]]++;
  const int MAX_ITER = 1000;
  int last_ordersum, iter = 0, ordersum = 0, maxorder = 0;
  do {
    last_ordersum = ordersum;
    for (int p = 0; p < NB; p++) {
      int pTo = _aDownstreamInds[p];
      _aSubBasinOrder[p] = (pTo == DOESNT_EXIST) ? 0 : _aSubBasinOrder[pTo] + 1;
    }
    ordersum = 0;
    for (int p = 0; p < NB; p++) { ordersum += _aSubBasinOrder[p]; if (_aSubBasinOrder[p] > maxorder) maxorder = _aSubBasinOrder[p]; }
    iter++;
  } while ((ordersum > last_ordersum) && (iter < MAX_ITER));
  ExitGracefullyIf(iter >= MAX_ITER, "CModel::InitializeRoutingNetwork: exceeded maximum iterations. Circular reference in basin connections?");
  int pp = 0;
  for (int ord = maxorder; ord >= 0; ord--)
    for (int p = 0; p < NB; p++) if (_aSubBasinOrder[p] == ord) _aOrderedSBind[pp++] = p;
  for (int p = 0; p < NB; p++) _pSubBasins[p]->is_headwater = (inflow[p] == 0);
}
void CModel::InitializeBasins(const optStruct &O) {
  const int NB = GetNumSubBasins();
  std::vector<double> aSBArea(NB), aSBQin(NB), aSBQlat(NB);
  double runoff_est = 0.0;
  for (int p = 0; p < NB; p++) {
    aSBArea[p] = _pSubBasins[p]->GetBasinArea();
    if (globals.v[RVN_G_AVG_ANNUAL_RUNOFF] > 0) runoff_est = globals.v[RVN_G_AVG_ANNUAL_RUNOFF] / DAYS_PER_YEAR;
    aSBQlat[p] = runoff_est / MM_PER_METER * (aSBArea[p] * M2_PER_KM2) / SEC_PER_DAY;
    // specified inflows at "t = 0" (src/ModelInitialize.cpp:663-665): asked before the series are initialised, i.e. with a zero time
    // offset - the first value of the series whatever its start date
    aSBQlat[p] += _pSubBasins[p]->pInflowHydro2 ? _pSubBasins[p]->pInflowHydro2->FirstValue() : 0.0;
    aSBQin[p] = _pSubBasins[p]->pInflowHydro ? _pSubBasins[p]->pInflowHydro->FirstValue() : 0.0;
  }
  for (int pp = 0; pp < NB; pp++) {
    int p = GetOrderedSubBasinIndex(pp);
    int pTo = _aDownstreamInds[p];
    if ((pTo != DOESNT_EXIST) && (!_pSubBasins[pTo]->disabled)) aSBQin[pTo] += aSBQlat[p] + aSBQin[p];
    if (pTo != DOESNT_EXIST) aSBArea[pTo] += aSBArea[p];
    _pSubBasins[p]->Initialize(aSBQin[p], aSBQlat[p], aSBArea[p], this, O);
  }
}
void CModel::GenerateGaugeWeights(const optStruct &O) {
  const int nH = GetNumHRUs(), nG = GetNumGauges();
  _aGaugeWeights.assign((size_t)nH * nG, 0.0);
  ExitGracefullyIf(nG == 0, "GenerateGaugeWeights: no gauges present");
  if (O.interpolation == INTERP_FROM_FILE) {
    // first line "nGauges nHRUs", then one row of nGauges weights per HRU (src/ModelInitialize.cpp:977-1060)
    CParser p(O.interp_file);
    ExitGracefullyIf(!p.ok(), "GenerateGaugeWeights:: Cannot find gauge weighting file " + O.interp_file);
    std::vector<std::string> s;
    bool have_dims = false; int k = 0;
    while (p.Tokenize(s)) {
      if (s.empty() || s[0] == ":GaugeWeightTable") continue;
      if (s[0] == ":EndGaugeWeightTable") break;
      if (!have_dims) {
        ExitGracefullyIf(s.size() < 2 || s_to_i(s[0]) != nG || s_to_i(s[1]) != nH, "GenerateGaugeWeights: improper number of gauges/HRUs in gauge weighting file");
        have_dims = true; continue;
      }
      ExitGracefullyIf((int)s.size() != nG || k >= nH, "GenerateGaugeWeights: incorrect number of columns/rows in gauge weighting file");
      double sum = 0;
      for (int g = 0; g < nG; g++) { _aGaugeWeights[(size_t)k * nG + g] = s_to_d(s[g]); sum += _aGaugeWeights[(size_t)k * nG + g]; }
      ExitGracefullyIf(fabs(sum - 1.0) > 1e-4, "GenerateGaugeWeights: user-specified weights for gauge don't add up to 1.0");
      k++;
    }
    ExitGracefullyIf(k != nH, "GenerateGaugeWeights: too few rows in gauge weighting file");
  } else if (O.interpolation == INTERP_AVERAGE_ALL) {
    for (int k = 0; k < nH; k++) for (int g = 0; g < nG; g++) _aGaugeWeights[(size_t)k * nG + g] = 1.0 / (double)(nG);
  } else {
    // distance-based interpolation (src/ModelInitialize.cpp:868-975) in the UTM zone of the area-weighted mean HRU longitude (:108-124).
    // One weight table serves all forcings: every gauge must carry the same series (the reference builds a table per forcing from the gauges
    // that have it).
    for (int g = 1; g < nG; g++)
      for (int f = 0; f < F_NTYPES; f++)
        ExitGracefullyIf((_pGauges[g]->GetTimeSeries((forcing_type)f) != NULL) != (_pGauges[0]->GetTimeSeries((forcing_type)f) != NULL),
                         "GenerateGaugeWeights: gauges carrying different forcing series are not supported by the B200 build with distance-based interpolation");
    double cen_long = 0, area_tot = 0;
    for (int k = 0; k < nH; k++) { area_tot += _pHydroUnits[k]->GetArea(); cen_long += _pHydroUnits[k]->longitude * (_pHydroUnits[k]->GetArea()); }
    cen_long /= area_tot;
    const int zone = (int)(floor((cen_long + 180.0) / 6) + 1);
    std::vector<double> gx(nG), gy(nG);
    for (int g = 0; g < nG; g++) LatLonToUTMXY(_pGauges[g]->latitude, _pGauges[g]->longitude, zone, gx[g], gy[g]);
    const double IDW_POWER = 2.0;
    for (int k = 0; k < nH; k++) {
      double hx, hy;
      LatLonToUTMXY(_pHydroUnits[k]->latitude, _pHydroUnits[k]->longitude, zone, hx, hy);
      double *w = &_aGaugeWeights[(size_t)k * nG];
      if (O.interpolation == INTERP_NEAREST_NEIGHBOR) {
        int g_min = 0; double distmin = ALMOST_INF;
        for (int g = 0; g < nG; g++) {
          const double dist = pow(hx - gx[g], 2) + pow(hy - gy[g], 2);
          if (dist < distmin) { distmin = dist; g_min = g; }
          w[g] = 0.0;
        }
        w[g_min] = 1.0;
      } else {
        const bool by_elev = (O.interpolation == INTERP_INVERSE_DISTANCE_ELEVATION);
        int atop_gauge = DOESNT_EXIST; double denomsum = 0;
        std::vector<double> dist(nG);
        for (int g = 0; g < nG; g++) {
          dist[g] = by_elev ? fabs(_pHydroUnits[k]->GetElevation() - _pGauges[g]->elevation) : sqrt(pow(hx - gx[g], 2) + pow(hy - gy[g], 2));
          denomsum += pow(dist[g], -IDW_POWER);
          if (dist[g] < REAL_SMALL) atop_gauge = g;
        }
        for (int g = 0; g < nG; g++) {
          w[g] = 0.0;
          if (atop_gauge != DOESNT_EXIST) { w[g] = 0.0; w[atop_gauge] = 1.0; }
          else w[g] = pow(dist[g], -IDW_POWER) / denomsum;
        }
      }
    }
  }
}
void CModel::Initialize(optStruct &O) {
  _pOpt = &O;
  ExitGracefullyIf(GetNumHRUs() == 0, "CModel::Initialize: Must have at least one hydrologic unit");
  ExitGracefullyIf(GetNumSubBasins() == 0, "CModel::Initialize: Must have at least one SubBasin");
  ExitGracefullyIf(GetNumProcesses() == 0, "CModel::Initialize: must have at least one hydrological process included in model");
  ExitGracefullyIf(GetNumProcesses() > 64, "CModel::Initialize: the B200 build supports at most 64 processes");
  _WatershedArea = 0.0;
  for (int p = 0; p < GetNumSubBasins(); p++) {
    CSubBasin *b = _pSubBasins[p];
    b->basin_area = 0.0;
    for (size_t i = 0; i < b->hru_index.size(); i++) b->basin_area += _pHydroUnits[b->hru_index[i]]->GetArea();
    _WatershedArea += b->basin_area;
  }
  for (int k = 0; k < GetNumHRUs(); k++) {
    const CHydroUnit *h = _pHydroUnits[k];
    ExitGracefullyIf(soil_profiles[h->soil_profile].nHorizons != _nSoilVars && h->type == RVN_HRU_STANDARD,
                     "CModel::Initialize: soil profile " + soil_profiles[h->soil_profile].tag + " does not have one horizon per model soil layer");
  }
  {  // default evaluation period: the whole simulation (CDiagPeriod "ALL", src/ModelInitialize.cpp:287; clipping src/Diagnostics.cpp:1572-1573)
    bool have_all = false;
    for (size_t i = 0; i < diag_periods.size(); i++) if (diag_periods[i].name == "ALL") have_all = true;
    if (!have_all) { diag_period P; P.name = "ALL"; P.t_start = 0.0; P.t_end = O.duration; diag_periods.push_back(P); }
    for (size_t i = 0; i < diag_periods.size(); i++) {
      diag_periods[i].t_start = rvh_max(rvh_min(diag_periods[i].t_start, O.duration), 0.0);
      diag_periods[i].t_end = rvh_max(rvh_min(diag_periods[i].t_end, O.duration), 0.0);
    }
  }
  for (int g = 0; g < GetNumGauges(); g++) _pGauges[g]->Initialize(O);
  for (size_t i = 0; i < observed.size(); i++) observed[i].pTS->Initialize(O.julian_start_day, O.julian_start_year, O.duration, O.timestep, O.calendar, true);
  GenerateGaugeWeights(O);
  InitializeRoutingNetwork();
  ExitGracefullyIf(GetNumSubBasins() > 1 && !(globals.v[RVN_G_AVG_ANNUAL_RUNOFF] > 0) && O.routing != RVN_ROUTE_NONE,
                   "CModel::Initialize: :AvgAnnualRunoff is required in the .rvp file when more than one subbasin is routed");
  InitializeBasins(O);
}
