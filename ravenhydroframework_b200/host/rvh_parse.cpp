// rvh_parse.cpp -- readers for the Raven input subset the hot path needs:
//   .rvi  model options + :HydrologicProcesses      (reference src/ParseInput.cpp)
//   .rvp  classes + parameter lists                 (reference src/ParsePropertyFile.cpp)
//   .rvh  subbasins + HRUs                          (reference src/ParseHRUFile.cpp)
//   .rvt  gauges + forcing series                   (reference src/ParseTimeSeriesFile.cpp, src/TimeSeries.cpp)
//   .rvc  initial conditions                        (reference src/ParseInitialConditionFile.cpp)
//   .rve  ensemble description                      (reference src/ParseEnsembleFile.cpp)
// Anything outside the subset fails loudly (RavenError) instead of being silently ignored, except for
// the output/diagnostic commands listed in kIgnoredRVI which do not influence the water balance.
#include "rvh_model.h"
#include "rvh_process.h"
#include "rvh_ensemble.h"
#include <set>

// decimal reader with the arithmetic of the reference's time-series reader (fast_s_to_d,
// src/CommonFunctions.cpp:1267-1342): digits are accumulated in floating point, so the result can
// differ from strtod in the last bit.  Forcing values must be the same doubles the reference sees.
static double ts_s_to_d(const char *p) {
  while (*p == ' ' || *p == '\t') p++;
  double sign = 1.0;
  if (*p == '-') { sign = -1.0; p++; } else if (*p == '+') { p++; }
  double value = 0.0;
  for (; *p >= '0' && *p <= '9'; p++) value = value * 10.0 + (*p - '0');
  if (*p == '.') {
    double pow10 = 10.0;
    p++;
    while (*p >= '0' && *p <= '9') { value += (*p - '0') / pow10; pow10 *= 10.0; p++; }
  }
  int frac = 0;
  double scale = 1.0;
  if (*p == 'e' || *p == 'E') {
    unsigned int expon = 0;
    p++;
    if (*p == '-') { frac = 1; p++; } else if (*p == '+') { p++; }
    for (; *p >= '0' && *p <= '9'; p++) expon = expon * 10 + (unsigned)(*p - '0');
    if (expon > 308) expon = 308;
    while (expon >= 50) { scale *= 1E50; expon -= 50; }
    while (expon >= 8) { scale *= 1E8; expon -= 8; }
    while (expon > 0) { scale *= 10.0; expon -= 1; }
  }
  return sign * (frac ? (value / scale) : (value * scale));
}

static double FixTimestep(double tstep) {
  double tmp = floor(1.0 / tstep + 0.5);
  ExitGracefullyIf(fabs(tstep * tmp - 1.0) > 0.1, "FixTimestep: timesteps and time intervals must evenly divide into one day");
  return 1.0 / tmp;
}
static bool IsValidDateString(const std::string &s) {
  return (s.length() == 10) && (s[4] == '/' || s[4] == '-') && (s[7] == '/' || s[7] == '-');
}
static bool LooksLikeClock(const std::string &t) { return t.length() >= 3 && (t[2] == ':' || t[1] == ':'); }
static double ParseIntervalToken(const std::string &t, int calendar) {
  if (LooksLikeClock(t)) return FixTimestep(DateStringToTimeStruct("0000-01-01", t, calendar).julian_day);
  return FixTimestep(s_to_d(t));
}

// ================================================================================================ .rvi
static const char *kIgnoredRVI[] = {":FileType", ":WrittenBy", ":CreationDate", ":Description", ":SilentMode", ":NoisyMode", ":QuietMode",
  ":EvaluationMetrics", ":CustomOutput", ":WriteForcingFunctions", ":WriteMassBalanceFile", ":WriteEnsimFormat", ":WriteNetCDFFormat",
  ":RunName", ":OutputDirectory", ":OutputInterval", ":SuppressOutput", ":DontWriteWatershedStorage", ":SnapshotHydrograph",
  ":PavicsMode", ":DebugMode", ":WriteExhaustiveMB", ":HydrologicProcesses", ":EndHydrologicProcesses",
  ":WriteSubbasinFile", ":NetCDFAttribute", ":rvh_Filename", ":AggregatedVariable", ":BenchmarkingMode", NULL};

static rvn_pet ParsePET(const std::string &s) {
  if (s == "PET_DATA") return RVN_PET_DATA;
  if (s == "PET_FROMMONTHLY") return RVN_PET_FROMMONTHLY;
  if (s == "PET_CONSTANT") return RVN_PET_CONSTANT;
  if (s == "PET_NONE") return RVN_PET_NONE;
  if (s == "PET_HARGREAVES_1985") return RVN_PET_HARGREAVES_1985;
  if (s == "PET_OUDIN") return RVN_PET_OUDIN;
  if (s == "PET_HAMON") return RVN_PET_HAMON;
  if (s == "PET_LINEAR_TEMP") return RVN_PET_LINEAR_TEMP;
  if (s == "PET_MOHYSE") return RVN_PET_MOHYSE;
  if (s == "PET_TURC_1961") return RVN_PET_TURC_1961;
  if (s == "PET_MAKKINK_1957") return RVN_PET_MAKKINK_1957;
  if (s == "PET_PENMAN_SIMPLE33") return RVN_PET_PENMAN_SIMPLE33;
  if (s == "PET_PENMAN_SIMPLE39") return RVN_PET_PENMAN_SIMPLE39;
  if (s == "PET_VAPDEFICIT") return RVN_PET_VAPDEFICIT;
  if (s == "PET_LINACRE") return RVN_PET_LINACRE;
  if (s == "PET_HARGREAVES") return RVN_PET_HARGREAVES;
  if (s == "PET_PRIESTLEY_TAYLOR") return RVN_PET_PRIESTLEY_TAYLOR;
  if (s == "PET_PENMAN_COMBINATION") return RVN_PET_PENMAN_COMBINATION;
  ExitGracefully("ParseMainInputFile: evaporation method " + s + " is not implemented in the B200 build");
  return RVN_PET_NONE;
}
static rvn_orocorr ParseOro(const std::string &s) {
  if (s == "OROCORR_NONE") return RVN_OROCORR_NONE;
  if (s == "OROCORR_SIMPLELAPSE") return RVN_OROCORR_SIMPLELAPSE;
  if (s == "OROCORR_HBV") return RVN_OROCORR_HBV;
  ExitGracefully("ParseMainInputFile: orographic correction " + s + " is not implemented in the B200 build");
  return RVN_OROCORR_NONE;
}

static void ParseMainInputFile(CModel *&pModel, optStruct &Options) {
  CParser p(Options.rvi_filename);
  ExitGracefullyIf(!p.ok(), "ParseMainInputFile: cannot open " + Options.rvi_filename);
  std::vector<std::string> s;
  CHydroProcessABC *pLast = NULL;
  CProcessGroup *pGroup = NULL;   // open :ProcessGroup, if any
  std::set<std::string> ignored;
  for (int i = 0; kIgnoredRVI[i]; i++) ignored.insert(kIgnoredRVI[i]);
  std::vector<std::pair<std::string, std::string> > pending_alias;
  while (p.Tokenize(s)) {
    if (s.empty()) continue;
    const std::string &c = s[0];
    const int Len = (int)s.size();
    if (c[0] != ':') continue;  // stray table rows of ignored blocks
    if (ignored.count(c)) continue;
    if (c == ":BenchmarkingMode") { Options.benchmarking = true; continue; }
    if (c == ":StartDate") {
      time_struct tt = DateStringToTimeStruct(s[1], Len == 2 ? "00:00:00" : s[2], Options.calendar);
      Options.julian_start_day = tt.julian_day; Options.julian_start_year = tt.year; continue;
    }
    if (c == ":JulianStartDay") { Options.julian_start_day = s_to_d(s[1]); continue; }
    if (c == ":JulianStartYear") { Options.julian_start_year = s_to_i(s[1]); continue; }
    if (c == ":Duration") {
      double part = 0.0;
      if (Len >= 3 && !IsComment(s[2]) && LooksLikeClock(s[2])) part = DateStringToTimeStruct("0000-01-01", s[2], Options.calendar).julian_day;
      Options.duration = s_to_d(s[1]) + part; continue;
    }
    if (c == ":EndDate") {
      time_struct tt = DateStringToTimeStruct(s[1], Len == 2 ? "00:00:00" : s[2], Options.calendar);
      Options.duration = TimeDifference(Options.julian_start_day, Options.julian_start_year, tt.julian_day, tt.year, Options.calendar); continue;
    }
    if (c == ":TimeStep") { Options.timestep = ParseIntervalToken(s[1], Options.calendar); continue; }
    if (c == ":Method" || c == ":NumericalMethod") {
      if (s[1] == "ORDERED_SERIES" || s[1] == "STANDARD") Options.sol_method = RVN_SOL_ORDERED_SERIES;
      else if (s[1] == "EULER") Options.sol_method = RVN_SOL_EULER;
      else if (s[1] == "ITER_HEUN") {   // src/ParseInput.cpp:742-746: the keyword is ITER_HEUN and both numbers are read unconditionally
        ExitGracefullyIf(Len < 4, "ParseMainInputFile: :Method ITER_HEUN needs [convergence criterion] [max iterations] (the reference reads both without checking)");
        Options.sol_method = RVN_SOL_ITERATED_HEUN;
        Options.convergence_crit = s_to_d(s[2]); Options.max_iterations = (int)s_to_d(s[3]);
      }
      else ExitGracefully("ParseMainInputFile: only :Method ORDERED_SERIES, EULER and ITER_HEUN are implemented in the B200 build");
      continue;
    }
    if (c == ":SoilModel") {
      if (s[1] == "SOIL_ONE_LAYER") Options.num_soillayers = 1;
      else if (s[1] == "SOIL_TWO_LAYER") Options.num_soillayers = 2;
      else if (s[1] == "SOIL_MULTILAYER") Options.num_soillayers = s_to_i(s[2]);
      else ExitGracefully("ParseMainInputFile: Unrecognized Soil model");
      pModel = new CModel(Options.num_soillayers);
      for (size_t i = 0; i < pending_alias.size(); i++) pModel->AddAlias(pending_alias[i].first, pending_alias[i].second);
      continue;
    }
    if (c == ":AssimilateStreamflow") { Options.assimilate_flow = true; continue; }   // src/ParseInput.cpp:1813-1818 (the per-gauge command of the .rvt has the same name)
    if (c == ":AssimilationStartTime") {   // :1795-1812
      ExitGracefullyIf(s.size() < 3, "ParseMainInputFile: :AssimilationStartTime needs [yyyy-mm-dd] [hh:mm:ss]");
      time_struct tt = DateStringToTimeStruct(s[1], s[2], Options.calendar);
      Options.assimilation_start = TimeDifference(Options.julian_start_day, Options.julian_start_year, tt.julian_day, tt.year, Options.calendar);
      if (Options.assimilation_start > Options.duration) WriteWarning("Data assimilation start date is after model simulation end. No data assimilation will be applied.");
      continue;
    }
    if (c == ":AssimilationMethod") { ExitGracefullyIf(s.size() >= 2 && s[1] != "DA_ECCC", "ParseMainInputFile: only :AssimilationMethod DA_ECCC is supported"); continue; }
    if (c == ":Routing") {
      if (s[1] == "ROUTE_NONE") Options.routing = RVN_ROUTE_NONE;
      else if (s[1] == "ROUTE_MUSKINGUM") Options.routing = RVN_ROUTE_MUSKINGUM;
      else if (s[1] == "ROUTE_MUSKINGUM_CUNGE") Options.routing = RVN_ROUTE_MUSKINGUM_CUNGE;
      else if (s[1] == "ROUTE_STORAGECOEFF") Options.routing = RVN_ROUTE_STORAGECOEFF;
      else if (s[1] == "ROUTE_PLUG_FLOW") Options.routing = RVN_ROUTE_PLUG_FLOW;
      else if (s[1] == "ROUTE_HYDROLOGIC") Options.routing = RVN_ROUTE_HYDROLOGIC;
      else if (s[1] == "ROUTE_DIFFUSIVE_WAVE") Options.routing = RVN_ROUTE_DIFFUSIVE_WAVE;
      else if (s[1] == "ROUTE_DIFFUSIVE_VARY") Options.routing = RVN_ROUTE_DIFFUSIVE_VARY;
      else ExitGracefully("ParseMainInputFile: routing method " + s[1] + " is not implemented in the B200 build (SURVEY 8f)");
      continue;
    }
    if (c == ":CatchmentRoute" || c == ":CatchmentRouting") {
      if (s[1] == "ROUTE_DUMP" || s[1] == "DUMP") Options.catchment_routing = ROUTE_DUMP;
      else if (s[1] == "TRIANGULAR_UH" || s[1] == "ROUTE_TRI_CONVOLUTION" || s[1] == "TRI_CONVOLUTION") Options.catchment_routing = ROUTE_TRI_CONVOLUTION;
      else if (s[1] == "GAMMA_UH" || s[1] == "ROUTE_GAMMA_CONVOLUTION" || s[1] == "GAMMA_CONVOLUTION") Options.catchment_routing = ROUTE_GAMMA_CONVOLUTION;
      else if (s[1] == "ROUTE_DELAYED_FIRST_ORDER" || s[1] == "DELAYED_FIRST_ORDER") Options.catchment_routing = ROUTE_DELAYED_FIRST_ORDER;
      else ExitGracefully("ParseMainInputFile: catchment routing method " + s[1] + " is not implemented in the B200 build");
      continue;
    }
    if (c == ":Evaporation") { Options.evaporation = ParsePET(s[1]); continue; }
    if (c == ":OW_Evaporation") { Options.ow_evaporation = ParsePET(s[1]); continue; }
    if (c == ":OroTempCorrect") { Options.orocorr_temp = ParseOro(s[1]); continue; }
    if (c == ":OroPrecipCorrect") { Options.orocorr_precip = ParseOro(s[1]); continue; }
    if (c == ":OroPETCorrect") { Options.orocorr_PET = ParseOro(s[1]); continue; }
    if (c == ":RainSnowFraction" || c == ":RainSnowMethod") {
      if (s[1] == "RAINSNOW_DATA") Options.rainsnow = RVN_RAINSNOW_DATA;
      else if (s[1] == "RAINSNOW_DINGMAN") Options.rainsnow = RVN_RAINSNOW_DINGMAN;
      else if (s[1] == "RAINSNOW_HBV") Options.rainsnow = RVN_RAINSNOW_HBV;
      else if (s[1] == "RAINSNOW_UBC" || s[1] == "RAINSNOW_UBCWM") Options.rainsnow = RVN_RAINSNOW_UBCWM;
      else if (s[1] == "RAINSNOW_THRESHOLD") Options.rainsnow = RVN_RAINSNOW_THRESHOLD;
      else ExitGracefully("ParseMainInputFile: rain/snow method " + s[1] + " is not implemented in the B200 build");
      continue;
    }
    // atmospheric estimators (src/ParseInput.cpp cases 24, 26, 29, 37, 23, 36): honoured where built; anything else is recorded and
    // becomes fatal only if a consumed PET method needs the atmospheric chain (the reference derives it for every run, used or not)
    if (c == ":SWRadiationMethod") {
      if (Len < 2) continue;
      if (s[1] == "SW_RAD_DEFAULT") Options.SW_radiation = RVN_SW_RAD_DEFAULT;
      else if (s[1] == "SW_RAD_NONE") Options.SW_radiation = RVN_SW_RAD_NONE;
      else if (Options.unsupported_atmos.empty()) Options.unsupported_atmos = c + " " + s[1];
      continue;
    }
    if (c == ":RelativeHumidityMethod") {
      if (Len < 2) continue;
      if (s[1] == "CONSTANT" || s[1] == "RELHUM_CONSTANT") Options.rel_humidity = RVN_RELHUM_CONSTANT;
      else if (s[1] == "MINDEWPT" || s[1] == "RELHUM_MINDEWPT") Options.rel_humidity = RVN_RELHUM_MINDEWPT;
      else if (Options.unsupported_atmos.empty()) Options.unsupported_atmos = c + " " + s[1];
      continue;
    }
    if (c == ":AirPressureMethod") {
      if (Len < 2) continue;
      if (s[1] == "BASIC" || s[1] == "AIRPRESS_BASIC") Options.air_pressure = RVN_AIRPRESS_BASIC;
      else if (s[1] == "UBC" || s[1] == "AIRPRESS_UBC") Options.air_pressure = RVN_AIRPRESS_UBC;
      else if (s[1] == "CONSTANT" || s[1] == "AIRPRESS_CONST") Options.air_pressure = RVN_AIRPRESS_CONST;
      else if (Options.unsupported_atmos.empty()) Options.unsupported_atmos = c + " " + s[1];
      continue;
    }
    if (c == ":SWCloudCorrect") {
      if (Len < 2) continue;
      if (s[1] == "SW_CLOUD_CORR_NONE") Options.SW_cloudcovercorr = RVN_SW_CLOUD_CORR_NONE;
      else if (s[1] == "SW_CLOUD_CORR_DINGMAN") Options.SW_cloudcovercorr = RVN_SW_CLOUD_CORR_DINGMAN;
      else if (s[1] == "SW_CLOUD_CORR_ANNANDALE") Options.SW_cloudcovercorr = RVN_SW_CLOUD_CORR_ANNANDALE;
      else if (Options.unsupported_atmos.empty()) Options.unsupported_atmos = c + " " + s[1];
      continue;
    }
    if (c == ":WindspeedMethod") {
      if (Len < 2) continue;
      if (s[1] == "CONSTANT" || s[1] == "WINDVEL_CONSTANT") Options.wind_velocity = RVN_WINDVEL_CONSTANT;
      else if (s[1] == "WINDVEL_UBC_MOD") Options.wind_velocity = RVN_WINDVEL_UBC_MOD;
      else if (s[1] == "WINDVEL_SQRT") Options.wind_velocity = RVN_WINDVEL_SQRT;
      else if (Options.unsupported_atmos.empty()) Options.unsupported_atmos = c + " " + s[1];
      continue;
    }
    if (c == ":CloudCoverMethod") {
      if (Len >= 2 && s[1] != "NONE" && s[1] != "CLOUDCOV_NONE" && Options.unsupported_atmos.empty()) Options.unsupported_atmos = c + " " + s[1];
      continue;
    }
    if (c == ":SWCanopyCorrect") {   // src/ParseInput.cpp; CRadiation::SWCanopyCorrection (src/Radiation.cpp:377-415)
      if (Len >= 2 && (s[1] == "SW_CANOPY_CORR_NONE" || s[1] == "NONE")) Options.SW_canopycorr = 0;
      else if (Len >= 2 && (s[1] == "SW_CANOPY_CORR_STATIC" || s[1] == "STATIC")) Options.SW_canopycorr = 1;
      else if (Options.unsupported_radiation.empty()) Options.unsupported_radiation = c + " " + (Len >= 2 ? s[1] : std::string());
      continue;
    }
    if (c == ":LWRadiationMethod") {
      if (Len >= 2 && (s[1] == "LW_RAD_DEFAULT" || s[1] == "DEFAULT")) Options.LW_radiation = 0;
      else if (Len >= 2 && (s[1] == "LW_RAD_NONE" || s[1] == "NONE")) Options.LW_radiation = 1;
      else if (Options.unsupported_radiation.empty()) Options.unsupported_radiation = c + " " + (Len >= 2 ? s[1] : std::string());
      continue;
    }
    if (c == ":LWIncomingMethod") {
      if (!(Len >= 2 && (s[1] == "LW_INC_DEFAULT" || s[1] == "DEFAULT")) && Options.unsupported_radiation.empty()) Options.unsupported_radiation = c + " " + (Len >= 2 ? s[1] : std::string());
      continue;
    }
    if (c == ":PotentialMeltMethod" || c == ":PotentialMelt") {
      if (s[1] == "POTMELT_DEGREE_DAY") Options.pot_melt = RVN_POTMELT_DEGREE_DAY;
      else if (s[1] == "POTMELT_HBV") Options.pot_melt = RVN_POTMELT_HBV;
      else if (s[1] == "POTMELT_HMETS") Options.pot_melt = RVN_POTMELT_HMETS;
      else if (s[1] == "POTMELT_DATA") Options.pot_melt = RVN_POTMELT_DATA;
      else if (s[1] == "POTMELT_NONE") Options.pot_melt = RVN_POTMELT_NONE;
      else if (s[1] == "POTMELT_DD_FREEZE") Options.pot_melt = RVN_POTMELT_DD_FREEZE;
      else if (s[1] == "POTMELT_HBV_ROS") Options.pot_melt = RVN_POTMELT_HBV_ROS;
      else if (s[1] == "POTMELT_EB") Options.pot_melt = RVN_POTMELT_EB;
      else if (s[1] == "POTMELT_RESTRICTED") Options.pot_melt = RVN_POTMELT_RESTRICTED;
      else if (s[1] == "POTMELT_DD_RAIN")   // no parameter list in the reference (src/ModelParamCheck.cpp:98-175): DD_MELT_TEMP is then overridden by the -77777.7 "not needed" sentinel
        ExitGracefully("ParseMainInputFile: POTMELT_DD_RAIN is not supported by the B200 build (the reference evaluates it with an unset DD_MELT_TEMP)");
      else ExitGracefully("ParseMainInputFile: potential melt method " + s[1] + " is not implemented in the B200 build");
      continue;
    }
    if (c == ":PrecipIceptFract") {
      if (s[1] == "USER_SPECIFIED" || s[1] == "ICEPT_USER" || s[1] == "PRECIP_ICEPT_USER") Options.interception_factor = RVN_ICEPT_USER;
      else if (s[1] == "LAI_LINEAR" || s[1] == "ICEPT_LAI" || s[1] == "PRECIP_ICEPT_LAI") Options.interception_factor = RVN_ICEPT_LAI;
      else if (s[1] == "LAI_EXPONENTIAL" || s[1] == "ICEPT_EXPLAI" || s[1] == "PRECIP_ICEPT_EXPLAI") Options.interception_factor = RVN_ICEPT_EXPLAI;
      else if (s[1] == "PRECIP_ICEPT_NONE") Options.interception_factor = RVN_ICEPT_NONE;
      else ExitGracefully("ParseMainInputFile: :PrecipIceptFract " + s[1] + " is not implemented in the B200 build");   // HEDSTROM / STICKY read the canopy snow load
      continue;
    }
    if (c == ":SubdailyMethod" || c == ":SubDailyMethod") {
      if (s[1] == "NONE" || s[1] == "SUBDAILY_NONE") Options.subdaily = 0;
      else if (s[1] == "SIMPLE" || s[1] == "SUBDAILY_SIMPLE") Options.subdaily = 1;   // src/ParseInput.cpp:1186-1191
      else ExitGracefully("ParseMainInputFile: :SubdailyMethod " + s[1] + " is not implemented in the B200 build (SUBDAILY_UBC re-derives every temperature of the day per step)");
      continue;
    }
    if (c == ":MonthlyInterpolationMethod") {
      if (s[1] == "UNIFORM") Options.month_interp_method = MONTHINT_UNIFORM;
      else if (s[1] == "LINEAR_START_OF_MONTH" || s[1] == "LINEAR_FOM") Options.month_interp_method = MONTHINT_LINEAR_FOM;
      else if (s[1] == "LINEAR_21" || s[1] == "MONTHINT_LINEAR_21") Options.month_interp_method = MONTHINT_LINEAR_21;
      else if (s[1] == "LINEAR_MID" || s[1] == "MONTHINT_LINEAR_MID") Options.month_interp_method = MONTHINT_LINEAR_MID;
      else ExitGracefully("ParseMainInputFile: Unrecognized monthly interpolation");
      continue;
    }
    if (c == ":InterpolationMethod" || c == ":Interpolation" || c == ":MetGaugeInterpolation") {
      if (s[1] == "NEAREST_NEIGHBOR" || s[1] == "INTERP_NEAREST_NEIGHBOR") Options.interpolation = INTERP_NEAREST_NEIGHBOR;
      else if (s[1] == "AVERAGE_ALL" || s[1] == "INTERP_AVERAGE_ALL") Options.interpolation = INTERP_AVERAGE_ALL;
      else if (s[1] == "INVERSE_DISTANCE" || s[1] == "INTERP_INVERSE_DISTANCE") Options.interpolation = INTERP_INVERSE_DISTANCE;
      else if (s[1] == "INVERSE_DISTANCE_ELEVATION" || s[1] == "INTERP_INVERSE_DISTANCE_ELEVATION") Options.interpolation = INTERP_INVERSE_DISTANCE_ELEVATION;
      else if (s[1] == "INTERP_FROM_FILE") { Options.interpolation = INTERP_FROM_FILE; Options.interp_file = CorrectForRelativePath(s[2], Options.rvi_filename); }
      else ExitGracefully("ParseMainInputFile: Unrecognized interpolation method");
      continue;
    }
    if (c == ":Alias") {
      if (pModel) pModel->AddAlias(s[1], s[2]); else pending_alias.push_back(std::make_pair(s[1], s[2]));
      continue;
    }
    if (c == ":LakeStorage") {
      ExitGracefullyIf(!pModel, ":LakeStorage command must be after :SoilModel command in .rvi file");
      pModel->SetLakeStorage(pModel->ParseSVTypeIndex(s[1])); continue;
    }
    if (c == ":SuppressCompetitiveET") { Options.suppressCompetitiveET = true; continue; }
    if (c == ":SnowSuppressesPET") { Options.snow_suppressPET = true; continue; }
    if (c == ":AllowSoilOverfill") { Options.allow_soil_overfill = true; continue; }
    if (c == ":DirectEvaporation") { Options.direct_evap = true; continue; }
    if (c == ":EnsembleMode") {
      if (s[1] == "ENSEMBLE_MONTECARLO") Options.ensemble = ENSEMBLE_MONTECARLO;
      else if (s[1] == "ENSEMBLE_NONE") Options.ensemble = ENSEMBLE_NONE;
      else if (s[1] == "ENSEMBLE_ENKF") Options.ensemble = ENSEMBLE_ENKF;
      else if (s[1] == "ENSEMBLE_DDS") Options.ensemble = ENSEMBLE_DDS;
      else ExitGracefully("ParseMainInputFile: unrecognized ensemble mode");
      if (pModel && Len >= 3) pModel->num_members = s_to_i(s[2]);
      continue;
    }
    if (c == ":RandomSeed") {  // src/ParseInput.cpp:1941-1946: the seed is offset by the start date known at this point of the file
      int timestamp_corr = Options.julian_start_year + int(round(Options.julian_start_day * 48));
      Options.random_seed = (unsigned int)(s_to_i(s[1]) + timestamp_corr); Options.random_seed_set = true; continue;
    }
    // ---- process list -----------------------------------------------------------------------------
    if (c == ":-->Conditional" || c == ":Conditional") {
      ExitGracefullyIf(!pLast || Len < 4, "ParseMainInputFile: :-->Conditional must follow a process");
      condition_basis basis; comparison cmp;
      if (s[1] == "HRU_TYPE") basis = BASIS_HRU_TYPE;
      else if (s[1] == "LAND_CLASS") basis = BASIS_LANDCLASS;
      else if (s[1] == "VEGETATION") basis = BASIS_VEGETATION;
      else if (s[1] == "HRU_GROUP") basis = BASIS_HRU_GROUP;
      else { ExitGracefully("ParseMainInputFile: unrecognized condition basis"); basis = BASIS_HRU_TYPE; }
      if (s[2] == "IS") cmp = COMPARE_IS_EQUAL;
      else if (s[2] == "IS_NOT") cmp = COMPARE_NOT_EQUAL;
      else { ExitGracefully("ParseMainInputFile: unrecognized comparison in :-->Conditional"); cmp = COMPARE_IS_EQUAL; }
      pLast->AddCondition(basis, cmp, s[3]);
      continue;
    }
    if (c == ":-->RedirectFlow") {
      ExitGracefullyIf(!pLast || Len < 3, "ParseMainInputFile: :-->RedirectFlow must follow a process");
      pLast->Redirect(pModel->ParseSVTypeIndex(s[1]), pModel->ParseSVTypeIndex(s[2]));
      continue;
    }
    ExitGracefullyIf(!pModel && c != ":FileType", "ParseMainInputFile: " + c + " appears before :SoilModel, or is not supported by the B200 build");
    if (c == ":EvaluationPeriod") {   // :EvaluationPeriod [name] [yyyy-mm-dd start] [yyyy-mm-dd end]  (src/ParseInput.cpp case 73, src/Diagnostics.cpp:1547-1574)
      ExitGracefullyIf(Len < 4, "ParseMainInputFile: :EvaluationPeriod needs [name] [start] [end]");
      CModel::diag_period P; P.name = s[1];
      time_struct t1 = DateStringToTimeStruct(s[2], "00:00:00", Options.calendar), t2 = DateStringToTimeStruct(s[3], "00:00:00", Options.calendar);
      P.t_start = TimeDifference(Options.julian_start_day, Options.julian_start_year, t1.julian_day, t1.year, Options.calendar);
      P.t_end = TimeDifference(Options.julian_start_day, Options.julian_start_year, t2.julian_day, t2.year, Options.calendar);
      ExitGracefullyIf(!pModel, "ParseMainInputFile: :EvaluationPeriod appears before :SoilModel");
      pModel->diag_periods.push_back(P);
      continue;
    }
    if (c == ":DefineHRUGroups" || c == ":DefineHRUGroup") {   // groups are populated by the .rvh file
      for (int i = 1; i < Len; i++) if (pModel->FindHRUGroup(s[i]) == DOESNT_EXIST) { CHRUGroup G; G.name = s[i]; pModel->hru_groups.push_back(G); }
      continue;
    }
    if (c == ":ProcessGroup") {   // src/ParseInput.cpp case 295
      ExitGracefullyIf(pGroup != NULL, "ParseMainInputFile: nested :ProcessGroup commands are not supported by the B200 build");
      pGroup = new CProcessGroup(Len > 1 ? s[1] : "", pModel);
      continue;
    }
    if (c == ":EndProcessGroup") {   // case 296: {wt1 .. wtN} | CALCULATE_WTS {r1 .. rN-1}
      ExitGracefullyIf(pGroup == NULL, "ParseMainInputFile: :EndProcessGroup without :ProcessGroup");
      const int N = pGroup->GetGroupSize();
      std::vector<double> aWts(N > 0 ? N : 1, 0.0);
      if (Len == N + 1) {
        if (s[1] == "CALCULATE_WTS") { for (int i = 0; i < N - 1; i++) aWts[i] = s_to_d(s[i + 2]); pGroup->CalcWeights(aWts.data(), N - 1); }
        else { for (int i = 0; i < N; i++) aWts[i] = s_to_d(s[i + 1]); pGroup->SetWeights(aWts.data(), N); }
      } else if (Len > 1) WriteWarning("ParseMainInputFile: incorrect number of weights in :EndProcessGroup command. Contents ignored.");
      pModel->AddProcess(pGroup); pLast = pGroup; pGroup = NULL;
      continue;
    }
    CHydroProcessABC *pMover = CreateProcessFromRVI(s, pModel, Options);
    if (pMover) {
      if (pGroup) pGroup->AddProcess(pMover); else pModel->AddProcess(pMover);
      pLast = pMover; continue;
    }
    ExitGracefully("ParseMainInputFile: command " + c + " (line " + std::to_string(p.line()) + " of " + Options.rvi_filename +
                   ") is not supported by the B200 build");
  }
  ExitGracefullyIf(!pModel, "ParseMainInputFile: :SoilModel command missing from .rvi file");
  // Add TOTAL_SWE state variable if any snow is simulated (src/ParseInput.cpp:3775-3779)
  if (pModel->StateVarExists(RVN_SV_SNOW)) { sv_type t = RVN_SV_TOTAL_SWE; int l = 0; pModel->AddStateVariables(&t, &l, 1); }
  if (pModel->StateVarExists(RVN_SV_GLACIER_ICE)) { sv_type t = RVN_SV_GLACIER_MB; int l = 0; pModel->AddStateVariables(&t, &l, 1); }
}

// ================================================================================================ .rvp
struct ParamTable {  // [DEFAULT] row + per-class explicit cells for one class family
  std::map<int, double> def;                              // field -> default
  std::map<std::pair<int, int>, double> cell;             // (class, field) -> explicit value
  std::set<int> no_default;                               // fields whose [DEFAULT] row the reference never applies (its class structs start from a literal)
};
template <class FieldFn>
static void ParseParameterList(CParser &p, const char *endtok, FieldFn field_of, int (CModel::*find)(const std::string &) const,
                               CModel *pModel, ParamTable &T, const char *family) {
  std::vector<std::string> s;
  std::vector<int> fields;
  while (p.Tokenize(s)) {
    if (s.empty()) continue;
    if (s[0] == endtok || s[0].compare(0, 4, ":End") == 0) return;   // the reference's ParsePropArray stops at any :End... line
    if (s[0] == ":Parameters") {
      fields.clear();
      for (size_t i = 1; i < s.size(); i++) {
        int f = field_of(s[i]);
        ExitGracefullyIf(f == -2, std::string("ParsePropertyFile: ") + family + " parameter " + s[i] + " is not supported by the B200 build");
        fields.push_back(f);
      }
      continue;
    }
    if (s[0] == ":Units") continue;
    ExitGracefullyIf(s.size() != fields.size() + 1, std::string("ParsePropertyFile: wrong number of columns in ") + family + " parameter list row " + s[0]);
    bool is_default = (s[0] == "[DEFAULT]");
    int c = -1;
    if (!is_default) {
      c = (pModel->*find)(s[0]);
      ExitGracefullyIf(c < 0, std::string("ParsePropertyFile: unknown ") + family + " class " + s[0]);
    }
    for (size_t i = 0; i < fields.size(); i++) {
      if (fields[i] < 0) continue;  // recognised but irrelevant to the hot path
      const std::string &v = s[i + 1];
      if (v == "_DEFAULT" || v == "_DEF") continue;
      if (v == "_AUTO" || v == "AUTO") {   // AUTO_COMPUTE: the class value is autogenerated even where a [DEFAULT] is given (SetCalculableValue, src/CommonFunctions.cpp:101-116)
        if (is_default) T.def.erase(fields[i]); else T.cell[std::make_pair(c, fields[i])] = NOT_SPECIFIED;
        continue;
      }
      if (is_default) { if (!T.no_default.count(fields[i])) T.def[fields[i]] = s_to_d(v); }
      else T.cell[std::make_pair(c, fields[i])] = s_to_d(v);
    }
  }
}

// name -> field, with -1 = "known to the reference, irrelevant here" and -2 = unknown
static int SoilFieldOrIgnore(const std::string &n) {
  int f = SoilFieldFromName(n);
  if (f >= 0) return f;
  static const char *ign[] = {"SAND_CON", "CLAY_CON", "SILT_CON", "ORG_CON", "STONE_FRAC", "BULK_DENSITY", "HYDRAUL_COND", NULL};
  for (int i = 0; ign[i]; i++) if (n == ign[i]) return -1;
  return -2;
}
static int LUFieldOrIgnore(const std::string &n) {
  int f = LandUseFieldFromName(n);
  if (f >= 0) return f;
  static const char *ign[] = {"FOREST_PET_CORR", NULL};
  for (int i = 0; ign[i]; i++) if (n == ign[i]) return -1;
  return -2;
}
static int VegFieldOrIgnore(const std::string &n) {
  if (n == "TFRAIN") return 1000 + RVN_V_RAIN_ICEPT_PCT;   // throughfall fractions: icept = 1 - value (src/ParsePropertyFile.cpp)
  if (n == "TFSNOW") return 1000 + RVN_V_SNOW_ICEPT_PCT;
  if (n == "RELATIVE_HT") return 2000;    // one value for all twelve months (src/VegetationClass.cpp:388-389); autogenerated 1.0
  if (n == "RELATIVE_LAI") return 2001;
  int f = VegFieldFromName(n);
  if (f >= 0) return f;
  static const char *ign[] = {"MAX_LEAF_COND", "ALBEDO_WET", "MAX_INTERCEPT_RATE", NULL};
  for (int i = 0; ign[i]; i++) if (n == ign[i]) return -1;
  return -2;
}

static void ParseClassPropertiesFile(CModel *pModel, const optStruct &Options) {
  CParser p(Options.rvp_filename);
  ExitGracefullyIf(!p.ok(), "ParseClassPropertiesFile: cannot open " + Options.rvp_filename);
  std::vector<std::string> s;
  ParamTable Tsoil, Tlu, Tveg;
  // ALBEDO_WET / ALBEDO_DRY of a soil class start at 0.16 / 0.24 (src/SoilClass.cpp:397-398), i.e. count as specified: SetCalculableValue never
  // consults the [DEFAULT] template for them
  Tsoil.no_default.insert(RVN_S_ALBEDO_WET); Tsoil.no_default.insert(RVN_S_ALBEDO_DRY);
  for (int f = 0; f < RVN_G_COUNT; f++) pModel->globals.v[f] = NOT_SPECIFIED;
  pModel->globals.v[RVN_G_TOC_MULTIPLIER] = pModel->globals.v[RVN_G_TIME_TO_PEAK_MULTIPLIER] = 1.0;
  pModel->globals.v[RVN_G_GAMMA_SHAPE_MULTIPLIER] = pModel->globals.v[RVN_G_GAMMA_SCALE_MULTIPLIER] = 1.0;
  std::map<int, double> veg_lai_set;
  while (p.Tokenize(s)) {
    if (s.empty()) continue;
    const std::string &c = s[0];
    if (c == ":FileType" || c == ":WrittenBy" || c == ":CreationDate") continue;
    if (c == ":SoilClasses") {
      while (p.Tokenize(s)) {
        if (s.empty() || s[0] == ":Attributes" || s[0] == ":Units") continue;
        if (s[0] == ":EndSoilClasses") break;
        CSoilClass sc; sc.tag = s[0];
        for (int f = 0; f < RVN_S_COUNT; f++) sc.S.v[f] = NOT_SPECIFIED;
        pModel->soil_classes.push_back(sc);
      }
      continue;
    }
    if (c == ":VegetationClasses") {
      while (p.Tokenize(s)) {
        if (s.empty() || s[0] == ":Attributes" || s[0] == ":Units") continue;
        if (s[0] == ":EndVegetationClasses") break;
        ExitGracefullyIf(s.size() < 4, "ParseClassPropertiesFile: :VegetationClasses rows need NAME MAX_HT MAX_LAI MAX_LEAF_COND");
        CVegetationClass vc; vc.tag = s[0];
        for (int f = 0; f < RVN_V_COUNT; f++) vc.V.v[f] = NOT_SPECIFIED;
        vc.V.v[RVN_V_MAX_HEIGHT] = s_to_d(s[1]);
        vc.V.v[RVN_V_MAX_LAI] = s_to_d(s[2]);
        for (int m = 0; m < 12; m++) { vc.V.relative_ht[m] = 1.0; vc.V.relative_LAI[m] = 1.0; }
        pModel->veg_classes.push_back(vc);
      }
      continue;
    }
    if (c == ":LandUseClasses") {
      while (p.Tokenize(s)) {
        if (s.empty() || s[0] == ":Attributes" || s[0] == ":Units") continue;
        if (s[0] == ":EndLandUseClasses") break;
        ExitGracefullyIf(s.size() < 3, "ParseClassPropertiesFile: :LandUseClasses rows need NAME IMPERM FOREST_COV");
        CLandUseClass lc; lc.tag = s[0];
        for (int f = 0; f < RVN_LU_COUNT; f++) lc.S.v[f] = NOT_SPECIFIED;
        lc.S.v[RVN_LU_IMPERMEABLE_FRAC] = s_to_d(s[1]);
        lc.S.v[RVN_LU_FOREST_COVERAGE] = s_to_d(s[2]);
        pModel->lu_classes.push_back(lc);
      }
      continue;
    }
    if (c == ":TerrainClasses") {
      while (p.Tokenize(s)) {
        if (s.empty() || s[0] == ":Attributes" || s[0] == ":Units") continue;
        if (s[0] == ":EndTerrainClasses") break;
        CTerrainClass tc; tc.tag = s[0];
        tc.T.v[RVN_T_TOPMODEL_LAMBDA] = (s.size() == 4) ? s_to_d(s[3]) : 7.5;   // CTerrainClass::InitializeTerrainProperties (src/TerrainClass.cpp:83-91), src/ParsePropertyFile.cpp:1019-1024
        pModel->terrain_classes.push_back(tc);
      }
      continue;
    }
    if (c == ":SoilProfiles") {
      while (p.Tokenize(s)) {
        if (s.empty() || s[0] == ":Attributes" || s[0] == ":Units") continue;
        if (s[0] == ":EndSoilProfiles") break;
        ExitGracefullyIf(s.size() < 2, "ParseClassPropertiesFile: bad :SoilProfiles row");
        CSoilProfile sp; sp.tag = s[0]; sp.nHorizons = s_to_i(s[1]);
        ExitGracefullyIf((int)s.size() < 2 + 2 * sp.nHorizons, "ParseClassPropertiesFile: :SoilProfiles row " + s[0] + " has too few horizon entries");
        for (int m = 0; m < sp.nHorizons; m++) {
          int sc = pModel->FindSoilClass(s[2 + 2 * m]);
          ExitGracefullyIf(sc < 0, "ParseClassPropertiesFile: bad soiltype code '" + s[2 + 2 * m] + "' in soil profile " + sp.tag);
          sp.soil_class.push_back(sc);
          sp.thickness.push_back(s_to_d(s[3 + 2 * m]));
        }
        pModel->soil_profiles.push_back(sp);
      }
      continue;
    }
    if (c == ":SoilParameterList") { ParseParameterList(p, ":EndSoilParameterList", SoilFieldOrIgnore, &CModel::FindSoilClass, pModel, Tsoil, "soil"); continue; }
    if (c == ":LandUseParameterList" || c == ":SnowParameterList") {   // :SnowParameterList is the same command (src/ParsePropertyFile.cpp:278)
      ParseParameterList(p, ":EndLandUseParameterList", LUFieldOrIgnore, &CModel::FindLUClass, pModel, Tlu, "land use"); continue; }
    if (c == ":VegetationParameterList") {
      ParseParameterList(p, ":EndVegetationParameterList", VegFieldOrIgnore, &CModel::FindVegClass, pModel, Tveg, "vegetation");
      for (size_t k = 0; k < pModel->veg_classes.size(); k++)   // RELATIVE_HT / RELATIVE_LAI take effect here so that a later :Seasonal... block overrides them
        for (int key = 2000; key <= 2001; key++) {
          std::map<std::pair<int, int>, double>::iterator it = Tveg.cell.find(std::make_pair((int)k, key));
          double val = NOT_SPECIFIED;
          if (it != Tveg.cell.end()) val = it->second;
          else if (Tveg.def.count(key)) val = Tveg.def[key];
          else continue;
          if (val == NOT_SPECIFIED) val = 1.0;   // AUTO_COMPUTE -> 1.0 (src/VegetationClass.cpp:205-225)
          for (int m = 0; m < 12; m++) (key == 2001 ? pModel->veg_classes[k].V.relative_LAI : pModel->veg_classes[k].V.relative_ht)[m] = val;
        }
      for (int key = 2000; key <= 2001; key++) {
        Tveg.def.erase(key);
        for (size_t k = 0; k < pModel->veg_classes.size(); k++) Tveg.cell.erase(std::make_pair((int)k, key));
      }
      continue;
    }
    if (c == ":GlobalParameter" && s.size() >= 3 && (s[1] == "ASSIMILATION_FACT" || s[1] == "ASSIM_UPSTREAM_DECAY" || s[1] == "ASSIM_TIME_DECAY")) {
      // streamflow-assimilation globals: host-side only (they shape the per-step tables of CModel::Flatten)
      optStruct &Ow = const_cast<optStruct &>(Options);
      if (s[1] == "ASSIMILATION_FACT") Ow.assimilation_fact = s_to_d(s[2]);
      else if (s[1] == "ASSIM_UPSTREAM_DECAY") Ow.assim_upstream_decay = s_to_d(s[2]);
      else Ow.assim_time_decay = s_to_d(s[2]);
      continue;
    }
    if (c == ":GlobalParameter") {
      int f = GlobalFieldFromName(s[1]);
      ExitGracefullyIf(f < 0, "ParseClassPropertiesFile: global parameter " + s[1] + " is not supported by the B200 build");
      pModel->globals.v[f] = s_to_d(s[2]);
      continue;
    }
    if (c == ":AvgAnnualRunoff") { pModel->globals.v[RVN_G_AVG_ANNUAL_RUNOFF] = s_to_d(s[1]); continue; }
    if (c == ":AvgAnnualSnow") { pModel->globals.v[RVN_G_AVG_ANNUAL_SNOW] = s_to_d(s[1]); continue; }
    if (c == ":AdiabaticLapseRate") { pModel->globals.v[RVN_G_ADIABATIC_LAPSE] = s_to_d(s[1]); continue; }
    if (c == ":RainSnowTransition") { pModel->globals.v[RVN_G_RAINSNOW_TEMP] = s_to_d(s[1]); pModel->globals.v[RVN_G_RAINSNOW_DELTA] = s_to_d(s[2]); continue; }
    if (c == ":IrreducibleSnowSaturation") { pModel->globals.v[RVN_G_SNOW_SWI] = s_to_d(s[1]); continue; }
    if (c == ":PrecipitationLapseRate") { pModel->globals.v[RVN_G_PRECIP_LAPSE] = s_to_d(s[1]); continue; }
    if (c == ":AirSnowCoeff") { pModel->globals.v[RVN_G_AIRSNOW_COEFF] = s_to_d(s[1]); continue; }
    if (c == ":SeasonalRelativeLAI" || c == ":SeasonalRelativeHeight" || c == ":SeasonalCanopyLAI" || c == ":SeasonalCanopyHeight") {   // src/ParsePropertyFile.cpp:282-285
      const bool lai = (c == ":SeasonalRelativeLAI" || c == ":SeasonalCanopyLAI");
      const std::string endt = ":End" + c.substr(1);
      while (p.Tokenize(s)) {
        if (s.empty() || s[0] == ":Attributes" || s[0] == ":Units") continue;
        if (s[0] == endt) break;
        ExitGracefullyIf(s.size() < 13, "ParseClassPropertiesFile: seasonal vegetation rows need 12 monthly values");
        if (s[0] == "[DEFAULT]") {
          for (size_t v = 0; v < pModel->veg_classes.size(); v++)
            for (int m = 0; m < 12; m++) (lai ? pModel->veg_classes[v].V.relative_LAI : pModel->veg_classes[v].V.relative_ht)[m] = s_to_d(s[1 + m]);
        } else {
          int v = pModel->FindVegClass(s[0]);
          ExitGracefullyIf(v < 0, "ParseClassPropertiesFile: unknown vegetation class " + s[0]);
          for (int m = 0; m < 12; m++) (lai ? pModel->veg_classes[v].V.relative_LAI : pModel->veg_classes[v].V.relative_ht)[m] = s_to_d(s[1 + m]);
        }
      }
      continue;
    }
    if (c == ":ChannelProfile") {   // src/ParsePropertyFile.cpp:804-915
      ExitGracefullyIf(s.size() < 2, "ParseClassPropertiesFile: :ChannelProfile needs a name");
      CChannelXSect *ch = new CChannelXSect();
      ch->name = s[1]; ch->bedslope = 0.0; ch->mannings_n = 0.0;
      std::vector<double> x, y, xz, n;
      int mode = 0;
      while (p.Tokenize(s)) {
        if (s.empty()) continue;
        if (s[0] == ":EndChannelProfile") break;
        if (s[0] == ":Bedslope" && s.size() == 2) { ch->bedslope = s_to_d(s[1]); continue; }
        if (s[0] == ":SurveyPoints") { mode = 1; continue; }
        if (s[0] == ":RoughnessZones") { mode = 2; continue; }
        if (s[0] == ":EndSurveyPoints" || s[0] == ":EndRoughnessZones") { mode = 0; continue; }
        ExitGracefullyIf(mode == 0 || s.size() != 2, "ParseClassPropertiesFile: improper format in :ChannelProfile " + ch->name);
        if (mode == 1) { x.push_back(s_to_d(s[0])); y.push_back(s_to_d(s[1])); }
        else { xz.push_back(s_to_d(s[0])); n.push_back(s_to_d(s[1])); }
      }
      const int nSP = (int)x.size(), nRS = (int)n.size();
      ExitGracefullyIf(nSP < 2 || nRS < 1, "ParseClassPropertiesFile: :ChannelProfile " + ch->name + " needs survey points and roughness zones");
      // Manning's n of survey segment i = length-weighted mean of the roughness zones it crosses
      std::vector<double> nn(nSP, n[nRS - 1]);
      int j = 0;
      for (int i = 0; i + 1 < nSP; i++) {
        if (j == nRS - 1) nn[i] = n[j];
        else if (xz[j + 1] > x[i + 1]) nn[i] = n[j];
        else {
          const double Li = (x[i + 1] - x[i]);
          nn[i] = (xz[j + 1] - x[i]) / Li * n[j];
          while ((j < nRS - 1) && (xz[j + 1] < x[i + 1])) {
            j++;
            if (j == nRS - 1) nn[i] += (x[i + 1] - xz[j]) / Li * n[j];
            else nn[i] += (std::min(x[i + 1], xz[j + 1]) - xz[j]) / Li * n[j];
          }
        }
      }
      ch->GenerateFromProfile(x, y, nn);
      pModel->channels.push_back(ch);
      continue;
    }
    if (c == ":TrapezoidalChannel") {  // name bot_width bot_elev incline mannings_n bedslope (src/ParsePropertyFile.cpp:982-988)
      ExitGracefullyIf(s.size() < 7, "ParseClassPropertiesFile: :TrapezoidalChannel needs 6 arguments");
      CChannelXSect *ch = new CChannelXSect();
      ch->name = s[1]; ch->mannings_n = s_to_d(s[5]); ch->bedslope = s_to_d(s[6]);
      ch->GenerateTrapezoid(s_to_d(s[2]), s_to_d(s[3]), s_to_d(s[4]));
      pModel->channels.push_back(ch);
      continue;
    }
    if (c == ":RedirectToFile") {   // src/ParseClassPropertiesFile.cpp: the redirect file is parsed in place
      ExitGracefullyIf(s.size() < 2, "ParseClassPropertiesFile: :RedirectToFile needs a file name");
      const std::string f = CorrectForRelativePath(s[1], Options.rvp_filename);
      ExitGracefullyIf(!p.Include(f), "ParseClassPropertiesFile: cannot open redirect file " + f);
      continue;
    }
    ExitGracefully("ParseClassPropertiesFile: command " + c + " (line " + std::to_string(p.line()) + ") is not supported by the B200 build");
  }
  // resolve [DEFAULT] rows into the classes
  for (size_t k = 0; k < pModel->soil_classes.size(); k++)
    for (int f = 0; f < RVN_S_COUNT; f++) {
      std::map<std::pair<int, int>, double>::iterator it = Tsoil.cell.find(std::make_pair((int)k, f));
      if (it != Tsoil.cell.end()) pModel->soil_classes[k].S.v[f] = it->second;
      else if (Tsoil.def.count(f)) pModel->soil_classes[k].S.v[f] = Tsoil.def[f];
    }
  for (size_t k = 0; k < pModel->lu_classes.size(); k++)
    for (int f = 0; f < RVN_LU_COUNT; f++) {
      std::map<std::pair<int, int>, double>::iterator it = Tlu.cell.find(std::make_pair((int)k, f));
      if (it != Tlu.cell.end()) pModel->lu_classes[k].S.v[f] = it->second;
      else if (Tlu.def.count(f)) pModel->lu_classes[k].S.v[f] = Tlu.def[f];
    }
  for (size_t k = 0; k < pModel->veg_classes.size(); k++) {
    for (int f = 0; f < RVN_V_COUNT; f++) {
      for (int alt = 0; alt < 2; alt++) {  // alt=1: throughfall aliases (1 - value)
        int key = f + 1000 * alt;
        std::map<std::pair<int, int>, double>::iterator it = Tveg.cell.find(std::make_pair((int)k, key));
        double val = NOT_SPECIFIED; bool have = false;
        if (it != Tveg.cell.end()) { val = it->second; have = true; }
        else if (Tveg.def.count(key)) { val = Tveg.def[key]; have = true; }
        if (have) pModel->veg_classes[k].V.v[f] = alt ? (1.0 - val) : val;
      }
    }
  }
  // derived / autogenerated values the hot path reads
  for (size_t k = 0; k < pModel->soil_classes.size(); k++) {
    soil_struct &S = pModel->soil_classes[k].S;
    // conceptual-model default: POROSITY autogenerates to 1.0 when no sand content is given (src/SoilClass.cpp:84-88)
    if (S.v[RVN_S_POROSITY] == NOT_SPECIFIED) S.v[RVN_S_POROSITY] = 1.0;
    S.v[RVN_S_CAP_RATIO] = S.v[RVN_S_POROSITY] * (1.0 - 0.0);  // stone_frac autogenerates to 0 (src/SoilClass.cpp:93-99,113)
    if (S.v[RVN_S_PET_CORRECTION] == NOT_SPECIFIED) S.v[RVN_S_PET_CORRECTION] = 1.0;  // src/SoilClass.cpp PET_CORRECTION default
    // bare-soil albedo: the reference initialises every soil class with 0.16 / 0.24 (CSoilClass::InitializeSoilProperties, src/SoilClass.cpp:397-398),
    // so its sand-content autogeneration (:312-327) only runs for an explicit _AUTO
    if (S.v[RVN_S_ALBEDO_WET] == NOT_SPECIFIED) S.v[RVN_S_ALBEDO_WET] = 0.16;
    if (S.v[RVN_S_ALBEDO_DRY] == NOT_SPECIFIED) S.v[RVN_S_ALBEDO_DRY] = 0.24;
  }
  for (size_t k = 0; k < pModel->lu_classes.size(); k++) {
    surface_struct &S = pModel->lu_classes[k].S;
    if (S.v[RVN_LU_FOREST_SPARSENESS] == NOT_SPECIFIED) S.v[RVN_LU_FOREST_SPARSENESS] = 0.0;  // src/LandUseClass.cpp autocalc
    if (S.v[RVN_LU_LAKE_PET_CORR] == NOT_SPECIFIED) S.v[RVN_LU_LAKE_PET_CORR] = 1.0;
    if (S.v[RVN_LU_MAX_SAT_AREA_FRAC] == NOT_SPECIFIED) S.v[RVN_LU_MAX_SAT_AREA_FRAC] = 1.0;
    if (S.v[RVN_LU_RELHUM_CORR] == NOT_SPECIFIED) S.v[RVN_LU_RELHUM_CORR] = 1.0;   // src/LandUseClass.cpp:264-267  // src/LandUseClass.cpp:93-98
    if (S.v[RVN_LU_OW_PET_CORR] == NOT_SPECIFIED) S.v[RVN_LU_OW_PET_CORR] = 1.0;
    if (S.v[RVN_LU_MIN_WIND_SPEED] == NOT_SPECIFIED) S.v[RVN_LU_MIN_WIND_SPEED] = 1.0;   // src/LandUseClass.cpp:246-255
    if (S.v[RVN_LU_MAX_WIND_SPEED] == NOT_SPECIFIED) S.v[RVN_LU_MAX_WIND_SPEED] = 8.0;
    if (S.v[RVN_LU_ROUGHNESS] == NOT_SPECIFIED) S.v[RVN_LU_ROUGHNESS] = 0.0;   // src/LandUseClass.cpp:85-91
    if (S.v[RVN_LU_PRIESTLEYTAYLOR_COEFF] == NOT_SPECIFIED) S.v[RVN_LU_PRIESTLEYTAYLOR_COEFF] = 1.26;   // :190-194
    // autogenerated snow parameters (CLandUseClass::AutoCalculateLandUseProps, src/LandUseClass.cpp:100-160)
    if (S.v[RVN_LU_FOREST_COVERAGE] == NOT_SPECIFIED) S.v[RVN_LU_FOREST_COVERAGE] = 0.0;
    if (S.v[RVN_LU_MELT_FACTOR] == NOT_SPECIFIED) S.v[RVN_LU_MELT_FACTOR] = 5.04;
    if (S.v[RVN_LU_DD_MELT_TEMP] == NOT_SPECIFIED) S.v[RVN_LU_DD_MELT_TEMP] = 0.0;
    if (S.v[RVN_LU_MIN_MELT_FACTOR] == NOT_SPECIFIED) S.v[RVN_LU_MIN_MELT_FACTOR] = S.v[RVN_LU_MELT_FACTOR];
    if (S.v[RVN_LU_REFREEZE_FACTOR] == NOT_SPECIFIED) S.v[RVN_LU_REFREEZE_FACTOR] = 5.04;
    if (S.v[RVN_LU_DD_REFREEZE_TEMP] == NOT_SPECIFIED) S.v[RVN_LU_DD_REFREEZE_TEMP] = 0.0;
    if (S.v[RVN_LU_REFREEZE_EXP] == NOT_SPECIFIED) S.v[RVN_LU_REFREEZE_EXP] = 1.0;
    if (S.v[RVN_LU_SCS_IA_FRACTION] == NOT_SPECIFIED) S.v[RVN_LU_SCS_IA_FRACTION] = 0.2;   // src/LandUseClass.cpp:195-199
    if (S.v[RVN_LU_HBV_MELT_FOR_CORR] == NOT_SPECIFIED) S.v[RVN_LU_HBV_MELT_FOR_CORR] = 1.0;
    if (S.v[RVN_LU_HBV_MELT_ASP_CORR] == NOT_SPECIFIED) S.v[RVN_LU_HBV_MELT_ASP_CORR] = 0.0;
    if (S.v[RVN_LU_HBV_MELT_GLACIER_CORR] == NOT_SPECIFIED) S.v[RVN_LU_HBV_MELT_GLACIER_CORR] = 1.0;
  }
  {  // autogenerated global parameters (CGlobalParams::AutoCalculateGlobalParams, src/GlobalParams.cpp:80-325)
    global_struct &G = pModel->globals;
    static const struct { int f; double v; } gd[] = {
      {RVN_G_SNOW_SWI, 0.05}, {RVN_G_SNOW_SWI_MIN, 0.05}, {RVN_G_SNOW_SWI_MAX, 0.05}, {RVN_G_SNOW_TEMPERATURE, 0.0}, {RVN_G_RAINSNOW_TEMP, 0.15},
      {RVN_G_RAINSNOW_DELTA, 2.0}, {RVN_G_MAX_SWE_SURFACE, 100.0}, {RVN_G_ADIABATIC_LAPSE, 6.4}, {RVN_G_PRECIP_LAPSE, 0.0},
      {RVN_G_HBVEC_LAPSE_RATE, 0.08}, {RVN_G_HBVEC_LAPSE_UPPER, 0.0}, {RVN_G_HBVEC_LAPSE_ELEV, 5000.0}, {RVN_G_SNOW_ROUGHNESS, 0.0002},
      {RVN_G_UBC_FLASH_PONDING, 36.0}};
    for (size_t i = 0; i < sizeof(gd) / sizeof(gd[0]); i++) if (G.v[gd[i].f] == NOT_SPECIFIED) G.v[gd[i].f] = gd[i].v;
  }
  for (size_t k = 0; k < pModel->veg_classes.size(); k++) {
    veg_struct &V = pModel->veg_classes[k].V;
    if (V.v[RVN_V_SAI_HT_RATIO] == NOT_SPECIFIED) V.v[RVN_V_SAI_HT_RATIO] = 0.0;
    if (V.v[RVN_V_PET_VEG_CORR] == NOT_SPECIFIED) V.v[RVN_V_PET_VEG_CORR] = 1.0;
    // CVegetationClass::AutoCalculateVegetationProps (src/VegetationClass.cpp:118-195)
    if (V.v[RVN_V_RAIN_ICEPT_PCT] == NOT_SPECIFIED) V.v[RVN_V_RAIN_ICEPT_PCT] = 0.12;
    if (V.v[RVN_V_SNOW_ICEPT_PCT] == NOT_SPECIFIED) V.v[RVN_V_SNOW_ICEPT_PCT] = 0.10;
    if (V.v[RVN_V_RAIN_ICEPT_FACT] == NOT_SPECIFIED) V.v[RVN_V_RAIN_ICEPT_FACT] = 0.06;   // src/VegetationClass.cpp:104-116
    if (V.v[RVN_V_SNOW_ICEPT_FACT] == NOT_SPECIFIED) V.v[RVN_V_SNOW_ICEPT_FACT] = 0.04;
    if (V.v[RVN_V_SVF_EXTINCTION] == NOT_SPECIFIED) V.v[RVN_V_SVF_EXTINCTION] = 0.5;   // src/VegetationClass.cpp:79-86
    if (V.v[RVN_V_ALBEDO] == NOT_SPECIFIED) V.v[RVN_V_ALBEDO] = 0.14;                  // :88-95
    const double max_LAI = V.v[RVN_V_MAX_LAI], max_SAI = V.v[RVN_V_SAI_HT_RATIO] * V.v[RVN_V_MAX_HEIGHT];
    if (V.v[RVN_V_MAX_CAPACITY] == NOT_SPECIFIED) V.v[RVN_V_MAX_CAPACITY] = 0.15 * (max_LAI + max_SAI);
    if (V.v[RVN_V_MAX_SNOW_CAPACITY] == NOT_SPECIFIED) V.v[RVN_V_MAX_SNOW_CAPACITY] = 0.6 * (max_LAI + max_SAI);
  }
}

// ================================================================================================ .rvh
// :Reservoir [name] ... :EndReservoir (ReservoirParse, src/ParseHRUFile.cpp:1312-2119).  The formats whose routing the device implements:
// lake type (:WeirCoefficient / :CrestWidth / :MaxDepth / :LakeArea [/ :AbsoluteCrestHeight], optional LOOKUP_TABLE :VolumeStageRelation /
// :AreaStageRelation overrides), :StageRelations lookup tables, POWER_LAW / LINEAR relations.  Everything else fails loudly.
static void ReservoirParse(CParser &p, const std::string &name, CModel *pModel) {
  enum { CURVE_POWERLAW, CURVE_LINEAR, CURVE_DATA, CURVE_LAKE } type = CURVE_POWERLAW;
  long long SBID = DOESNT_EXIST, HRUID = DOESNT_EXIST;
  double a_V = 1000.0, b_V = 1.0, a_Q = 10.0, b_Q = 1.0, a_A = AUTO_COMPUTE, b_A = AUTO_COMPUTE;
  double cwidth = DOESNT_EXIST, max_depth = DOESNT_EXIST, weircoeff = DOESNT_EXIST, lakearea = DOESNT_EXIST, crestht = 0.0, max_capacity = 0, demand_mult = 1.0;
  std::vector<double> aQ, aQ_ht, aV, aV_ht, aA, aA_ht, aQund;
  std::vector<std::string> s;
  auto next = [&]() { do { ExitGracefullyIf(!p.Tokenize(s), "ReservoirParse: unexpected end of file inside :Reservoir block"); } while (s.empty()); };
  auto table = [&](std::vector<double> &x, std::vector<double> &y) {
    next(); const int n = s_to_i(s[0]);
    x.clear(); y.clear();
    for (int i = 0; i < n; i++) { next(); ExitGracefullyIf(s.size() < 2, "ReservoirParse: lookup table rows need two columns"); x.push_back(s_to_d(s[0])); y.push_back(s_to_d(s[1])); }
    next();   // :End...Relation
  };
  for (;;) {
    next();
    const std::string c = s[0];
    if (c == ":EndReservoir") break;
    else if (c == ":SubBasinID") SBID = s_to_ll(s[1]);
    else if (c == ":HRUID") HRUID = s_to_ll(s[1]);
    else if (c == ":Type") {}
    else if (c == ":CrestWidth") { cwidth = s_to_d(s[1]); type = CURVE_LAKE; }
    else if (c == ":MaxCapacity") max_capacity = s_to_d(s[1]);
    else if (c == ":WeirCoefficient") { weircoeff = s_to_d(s[1]); type = CURVE_LAKE; }
    else if (c == ":MaxDepth") { max_depth = s_to_d(s[1]); type = CURVE_LAKE; }
    else if (c == ":LakeArea") { lakearea = s_to_d(s[1]); type = CURVE_LAKE; }
    else if (c == ":AbsoluteCrestHeight") { crestht = s_to_d(s[1]); type = CURVE_LAKE; }
    else if (c == ":DemandMultiplier") demand_mult = s_to_d(s[1]);
    else if (c == ":LakebedThickness" || c == ":LakebedThermalConductivity" || c == ":LakeConvectionCoeff") {}   // thermal properties: transport only
    else if (c == ":VolumeStageRelation") {
      ExitGracefullyIf(s.size() < 2, "VolumeStageRelation : Must specify whether LOOKUP_TABLE or POWER_LAW");
      if (s[1] == "POWER_LAW") { type = CURVE_POWERLAW; next(); if (s.size() >= 2) { a_V = s_to_d(s[0]); b_V = s_to_d(s[1]); } next(); }
      else if (s[1] == "LINEAR") { next(); if (s.size() >= 1) { a_V = s_to_d(s[0]); b_V = 1.0; } next(); }
      else if (s[1] == "LOOKUP_TABLE") table(aV_ht, aV);
    } else if (c == ":AreaStageRelation") {
      ExitGracefullyIf(s.size() < 2, "AreaStageRelation : Must specify whether LOOKUP_TABLE or POWER_LAW");
      if (s[1] == "POWER_LAW") { type = CURVE_POWERLAW; next(); if (s.size() >= 2) { a_A = s_to_d(s[0]); b_A = s_to_d(s[1]); } next(); }
      else if (s[1] == "CONSTANT") { type = CURVE_LINEAR; next(); if (s.size() >= 1) { a_A = s_to_d(s[0]); b_A = 0.0; } next(); }
      else if (s[1] == "LINEAR") { type = CURVE_LINEAR; next(); if (s.size() >= 1) { a_A = s_to_d(s[0]); b_A = 1.0; } next(); }
      else if (s[1] == "LOOKUP_TABLE") table(aA_ht, aA);
    } else if (c == ":OutflowStageRelation") {
      ExitGracefullyIf(s.size() < 2, "OutflowStageRelation : Must specify whether LOOKUP_TABLE or POWER_LAW");
      if (s[1] == "POWER_LAW") { type = CURVE_POWERLAW; next(); if (s.size() >= 2) { a_Q = s_to_d(s[0]); b_Q = s_to_d(s[1]); } next(); }
      else if (s[1] == "LINEAR") { type = CURVE_LINEAR; next(); if (s.size() >= 1) { a_Q = s_to_d(s[0]); b_Q = 1.0; } next(); }
      // the reference tests s[0] (not s[1]) against LOOKUP_TABLE here (src/ParseHRUFile.cpp:1663), so that branch is never taken
    } else if (c == ":StageRelations") {
      type = CURVE_DATA;
      next(); const int n = s_to_i(s[0]);
      for (int i = 0; i < n; i++) {
        next();
        ExitGracefullyIf(s.size() < 4, "Incorrect line length (<4) in :Reservoir :StageRelations command");
        aQ_ht.push_back(s_to_d(s[0])); aQ.push_back(s_to_d(s[1])); aV.push_back(s_to_d(s[2])); aA.push_back(s_to_d(s[3]));
        aQund.push_back(s.size() >= 5 ? s_to_d(s[4]) : 0.0);
      }
      next();   // :EndStageRelations
    } else if (c == ":EndStageRelations") {}
    else ExitGracefully("ReservoirParse: command " + c + " in :Reservoir block (line " + std::to_string(p.line()) + ") is not supported by the B200 build");
  }
  CReservoir *pRes = new CReservoir(name, SBID);
  if (type == CURVE_POWERLAW || type == CURVE_LINEAR) {
    ExitGracefullyIf(max_depth == DOESNT_EXIST, "ReservoirParse: :MaxDepth must be given for POWER_LAW / LINEAR stage relations");
    pRes->SetPowerLaw(a_V, b_V, a_Q, b_Q, a_A, b_A, crestht, max_depth);
  } else if (type == CURVE_DATA) pRes->SetLookup(aQ_ht, aQ, aQund, aA, aV);
  else {
    ExitGracefullyIf(weircoeff == DOESNT_EXIST, "CReservoir::Parse: :WeirCoefficient must be specified for lake-type reservoirs");
    ExitGracefullyIf(cwidth == DOESNT_EXIST, "CReservoir::Parse: :CrestWidth must be specified for lake-type reservoirs");
    ExitGracefullyIf(lakearea == DOESNT_EXIST, "CReservoir::Parse: :LakeArea  must be specified for lake-type reservoirs");
    ExitGracefullyIf(max_depth == DOESNT_EXIST, "CReservoir::Parse: :LakeDepth must be specified for lake-type reservoirs");
    pRes->SetLake(weircoeff, cwidth, crestht, lakearea, max_depth);
  }
  if (max_capacity != 0) pRes->_max_capacity = max_capacity;
  if ((type == CURVE_LAKE) && !aV.empty() && !aV_ht.empty()) pRes->SetVolumeStageCurve(aV_ht, aV, weircoeff, cwidth);
  if ((type == CURVE_LAKE) && !aA.empty() && !aA_ht.empty()) pRes->SetAreaStageCurve(aA_ht, aA);
  pRes->_demand_mult = demand_mult;
  const int pb = pModel->GetSubBasinIndex(pRes->GetSubbasinID());
  ExitGracefullyIf(pb < 0, "ParseHRUProps: Bad Sub-basin index in :Reservoir command");
  if (HRUID != DOESNT_EXIST)
    for (int k = 0; k < pModel->GetNumHRUs(); k++) if (pModel->GetHydroUnit(k)->ID == HRUID) { pRes->_HRUID = HRUID; pRes->_hru_k = k; break; }
  ExitGracefullyIf(pModel->GetSubBasin(pb)->pReservoir != NULL, "CSubBasin::AddReservoir: only one inflow reservoir may be specified per basin");
  pModel->GetSubBasin(pb)->pReservoir = pRes;
}

static void ParseHRUPropsFileInner(CModel *pModel, const optStruct &Options, const std::string &filename);
static void ParseHRUPropsFile(CModel *pModel, const optStruct &Options) { ParseHRUPropsFileInner(pModel, Options, Options.rvh_filename); }
static void ParseHRUPropsFileInner(CModel *pModel, const optStruct &Options, const std::string &filename) {
  CParser p(filename);
  ExitGracefullyIf(!p.ok(), "ParseHRUPropsFile: cannot open " + filename);
  std::vector<std::string> s;
  while (p.Tokenize(s)) {
    if (s.empty()) continue;
    const std::string &c = s[0];
    if (c == ":FileType" || c == ":WrittenBy" || c == ":CreationDate") continue;
    if (c == ":RedirectToFile") { ExitGracefullyIf(s.size() < 2, "ParseHRUPropsFile: :RedirectToFile needs a file name"); ParseHRUPropsFileInner(pModel, Options, CorrectForRelativePath(s[1], filename)); continue; }
    if (c == ":Reservoir") { ExitGracefullyIf(s.size() < 2, "ParseHRUProps: No reservoir name provided in :Reservoir command "); const std::string nm = s[1]; ReservoirParse(p, nm, pModel); continue; }
    if (c == ":SubBasins") {
      while (p.Tokenize(s)) {
        if (s.empty() || s[0] == ":Attributes" || s[0] == ":Units") continue;
        if (s[0] == ":EndSubBasins") break;
        ExitGracefullyIf(s.size() < 6, "ParseHRUPropsFile: :SubBasins rows need ID NAME DOWNSTREAM_ID PROFILE REACH_LENGTH GAUGED");
        CSubBasin *b = new CSubBasin();
        b->ID = s_to_ll(s[0]); b->name = s[1]; b->downstream_ID = s_to_ll(s[2]); b->profile_name = s[3];
        if (b->downstream_ID < 0) b->downstream_ID = DOESNT_EXIST;
        b->pChannel = NULL;
        if (StringToUppercase(s[3]) != "NONE") {
          b->pChannel = pModel->FindChannel(s[3]);
          ExitGracefullyIf(!b->pChannel, "ParseHRUPropsFile: Unrecognized Channel profile code (" + s[3] + ") in SubBasins command");
        }
        if (s[4] == "_AUTO" || StringToUppercase(s[4]) == "AUTO") b->reach_length = AUTO_COMPUTE;
        else b->reach_length = s_to_d(s[4]) * M_PER_KM;
        b->gauged = s_to_i(s[5]) != 0;
        pModel->AddSubBasin(b);
      }
      continue;
    }
    if (c == ":HRUs") {
      while (p.Tokenize(s)) {
        if (s.empty() || s[0] == ":Attributes" || s[0] == ":Units") continue;
        if (s[0] == ":EndHRUs") break;
        ExitGracefullyIf(s.size() < 13, "ParseHRUPropsFile: :HRUs rows need 13 columns");
        CHydroUnit *h = new CHydroUnit();
        h->ID = s_to_ll(s[0]); h->area = s_to_d(s[1]); h->elevation = s_to_d(s[2]);
        h->latitude = s_to_d(s[3]); h->longitude = s_to_d(s[4]); h->SBID = s_to_ll(s[5]);
        h->lu_class = pModel->FindLUClass(s[6]);
        ExitGracefullyIf(h->lu_class < 0, "ParseHRUPropsFile: Unrecognized Land Use Code/index in HRU definition: " + s[6]);
        h->veg_class = pModel->FindVegClass(s[7]);
        ExitGracefullyIf(h->veg_class < 0, "ParseHRUPropsFile: Unrecognized Vegetation Code/index in HRU definition: " + s[7]);
        h->soil_profile = pModel->FindSoilProfile(s[8]);
        ExitGracefullyIf(h->soil_profile < 0, "ParseHRUPropsFile: Unrecognized Soil Profile Code/index in HRU definition: " + s[8]);
        h->terrain_class = (s[10] == "[NONE]") ? DOESNT_EXIST : pModel->FindTerrainClass(s[10]);
        // HRU type from the soil-profile tag (reference src/ParseHRUFile.cpp: LAKE/GLACIER/ROCK/WATER/WETLAND prefixes)
        std::string tag = pModel->soil_profiles[h->soil_profile].tag;
        h->type = RVN_HRU_STANDARD;
        if (tag.substr(0, 4) == "LAKE") h->type = RVN_HRU_LAKE;
        else if (tag.substr(0, 7) == "GLACIER") h->type = RVN_HRU_GLACIER;
        else if (tag.substr(0, 5) == "WATER") h->type = RVN_HRU_WATER;
        else if (tag.substr(0, 4) == "ROCK") h->type = RVN_HRU_ROCK;
        else if (tag.substr(0, 8) == "PAVEMENT") h->type = RVN_HRU_ROCK;
        else if (tag.substr(0, 7) == "WETLAND") h->type = RVN_HRU_WETLAND;
        h->slope = (s[11] == "[NONE]") ? 0.0 : s_to_d(s[11]);
        h->aspect = (s[12] == "[NONE]") ? 0.0 : s_to_d(s[12]);
        h->enabled = true;
        h->subbasin_index = pModel->GetSubBasinIndex(h->SBID);
        ExitGracefullyIf(h->subbasin_index < 0, "ParseHRUPropsFile: HRU " + s[0] + " references unknown subbasin " + s[5]);
        ExitGracefullyIf(h->area <= 0.0, "ParseHRUPropsFile: HRU area must be positive");
        pModel->GetSubBasin(h->subbasin_index)->hru_index.push_back(pModel->GetNumHRUs());
        pModel->AddHRU(h);
      }
      continue;
    }
    if (c == ":SubBasinProperties") {
      std::vector<std::string> names;
      while (p.Tokenize(s)) {
        if (s.empty() || s[0] == ":Units") continue;
        if (s[0] == ":EndSubBasinProperties") break;
        if (s[0] == ":Parameters" || s[0] == ":Attributes") { names.assign(s.begin() + 1, s.end()); continue; }
        int pb = pModel->GetSubBasinIndex(s_to_ll(s[0]));
        ExitGracefullyIf(pb < 0, "ParseHRUPropsFile: bad subbasin ID in :SubBasinProperties");
        for (size_t i = 0; i < names.size() && i + 1 < s.size(); i++)
          ExitGracefullyIf(!pModel->GetSubBasin(pb)->SetBasinProperties(names[i], s_to_d(s[i + 1])),
                           "ParseHRUPropsFile: subbasin property " + names[i] + " is not supported by the B200 build");
      }
      continue;
    }
    if (c == ":HRUGroup" || c == ":SubBasinGroup") {
      // {ID1,ID2,a-b,...} lines until :End...; members keep the listed order (src/ParseHRUFile.cpp:472-530, 660-720)
      const std::string cmd = c;   // `c` refers into the token list, which the loop below overwrites
      const bool hru = (cmd == ":HRUGroup");
      ExitGracefullyIf(s.size() < 2, "ParseHRUPropsFile: " + cmd + " needs a group name");
      const std::string gname = s[1];
      std::vector<int> members;
      std::map<long long, int> hru_of_id;
      if (hru) for (int k = 0; k < pModel->GetNumHRUs(); k++) hru_of_id.insert(std::make_pair(pModel->GetHydroUnit(k)->ID, k));   // first HRU with an ID wins, as in the reference's scan
      while (p.Tokenize(s)) {
        if (s.empty()) continue;
        if (s[0] == (hru ? ":EndHRUGroup" : ":EndSubBasinGroup")) break;
        ExitGracefullyIf(s[0][0] == ':', "ParseHRUPropsFile: command found between " + cmd + " ... :End commands; only IDs should be in this block");
        for (size_t i = 0; i < s.size(); i++) {
          long long a, b;
          size_t dash = s[i].find('-', 1);
          if (dash == std::string::npos) a = b = s_to_ll(s[i]);
          else { a = s_to_ll(s[i].substr(0, dash)); b = s_to_ll(s[i].substr(dash + 1)); }
          ExitGracefullyIf((b - a) > 10000000, "Parsing " + cmd + " command: invalid range of indices");
          for (long long id = a; id <= b; id++) {
            int idx = DOESNT_EXIST;
            if (hru) { std::map<long long, int>::const_iterator it = hru_of_id.find(id); if (it != hru_of_id.end()) idx = it->second; }
            else idx = pModel->GetSubBasinIndex(id);
            if (idx < 0) { if (a == b) WriteWarning("ID " + std::to_string(id) + " specified in " + cmd + " command for group " + gname + " does not exist"); continue; }
            members.push_back(idx);
          }
        }
      }
      if (hru) {
        int g = pModel->FindHRUGroup(gname);
        if (g == DOESNT_EXIST) { CHRUGroup G; G.name = gname; pModel->hru_groups.push_back(G); g = (int)pModel->hru_groups.size() - 1; }
        pModel->hru_groups[g].hru.insert(pModel->hru_groups[g].hru.end(), members.begin(), members.end());
      } else {
        int g = pModel->FindSubBasinGroup(gname);
        if (g == DOESNT_EXIST) { CSubbasinGroup G; G.name = gname; pModel->sb_groups.push_back(G); g = (int)pModel->sb_groups.size() - 1; }
        pModel->sb_groups[g].sb.insert(pModel->sb_groups[g].sb.end(), members.begin(), members.end());
      }
      continue;
    }
    if (c == ":LateralConnections") {   // src/ParseHRUFile.cpp:1134-1200: {source HRU ID} {recipient HRU ID} {weight} rows
      ExitGracefullyIf(s.size() < 2, "ParseHRUPropsFile: :LateralConnections needs a process name");
      ExitGracefullyIf(s[1] != "SNOW_REDISTRIBUTE", "ParseHRUPropsFile: :LateralConnections " + s[1] + " is not supported by the B200 build");
      std::vector<CmvLatRedistribute *> targets;
      for (int j = 0; j < pModel->GetNumProcesses(); j++)
        if (CmvLatRedistribute *r = dynamic_cast<CmvLatRedistribute *>(pModel->GetProcess(j))) { r->ClearConnections(); targets.push_back(r); }
      while (p.Tokenize(s)) {
        if (s.empty()) continue;
        if (s[0] == ":EndLateralConnections") break;
        ExitGracefullyIf(s.size() < 3, "ParseHRUPropsFile: improper format of a :LateralConnections row");
        for (size_t t = 0; t < targets.size(); t++) targets[t]->AddConnection(s_to_ll(s[0]), s_to_ll(s[1]), s_to_d(s[2]));
      }
      continue;
    }
    if (c == ":PopulateHRUGroup" || c == ":PopulateSubBasinGroup") continue;   // output-only groups
    ExitGracefully("ParseHRUPropsFile: command " + c + " (line " + std::to_string(p.line()) + ") is not supported by the B200 build");
  }
}

// ================================================================================================ .rvt
forcing_type ForcingFromString(const std::string &s) {
  if (s == "PRECIP") return F_PRECIP;
  if (s == "RAINFALL") return F_RAINFALL;
  if (s == "SNOWFALL") return F_SNOWFALL;
  if (s == "TEMP_AVE") return F_TEMP_AVE;
  if (s == "TEMP_DAILY_MIN" || s == "TEMP_MIN") return F_TEMP_DAILY_MIN;
  if (s == "TEMP_DAILY_MAX" || s == "TEMP_MAX") return F_TEMP_DAILY_MAX;
  if (s == "TEMP_DAILY_AVE") return F_TEMP_DAILY_AVE;
  if (s == "PET") return F_PET;
  if (s == "OW_PET") return F_OW_PET;
  if (s == "POTENTIAL_MELT") return F_POTENTIAL_MELT;
  return F_NTYPES;
}

static void ReadSeriesHeader(CParser &p, const optStruct &Options, double &start_day, int &start_yr, double &interval, int &n) {
  std::vector<std::string> s;
  do { ExitGracefullyIf(!p.Tokenize(s), "ParseTimeSeriesFile: unexpected end of file in time series header"); } while (s.empty());
  ExitGracefullyIf(s.size() < 4, "ParseTimeSeriesFile: time series header needs [date] [time] [interval] [N]");
  if (IsValidDateString(s[0])) {
    time_struct tt = DateStringToTimeStruct(s[0], s[1], Options.calendar);
    start_day = tt.julian_day; start_yr = tt.year;
    interval = ParseIntervalToken(s[2], Options.calendar);
    n = s_to_i(s[3]);
  } else {
    n = s_to_i(s[0]); start_day = s_to_d(s[1]); start_yr = s_to_i(s[2]); interval = FixTimestep(s_to_d(s[3]));
  }
}

static void ParseGaugeBody(CParser &p, CGauge *g, const optStruct &Options, const std::string &thisfile);
static void ParseMultiData(CParser &p, CGauge *g, const optStruct &Options) {
  double start_day, interval; int start_yr, n;
  ReadSeriesHeader(p, Options, start_day, start_yr, interval, n);
  std::vector<std::string> s;
  do { ExitGracefullyIf(!p.Tokenize(s), "ParseTimeSeriesFile: truncated :MultiData"); } while (s.empty());
  ExitGracefullyIf(s[0] != ":Parameters", "CTimeSeries::ParseMultiple : MultiData command improperly formatted");
  std::vector<forcing_type> types;
  for (size_t i = 1; i < s.size(); i++) {
    forcing_type f = ForcingFromString(s[i]);
    ExitGracefullyIf(f == F_NTYPES, "CTimeSeries::ParseMultiple : time series type " + s[i] + " is not supported by the B200 build");
    types.push_back(f);
  }
  do { ExitGracefullyIf(!p.Tokenize(s), "ParseTimeSeriesFile: truncated :MultiData"); } while (s.empty());
  ExitGracefullyIf(s[0] != ":Units", "CTimeSeries::ParseMultiple : MultiData command improperly formatted");
  std::vector<std::vector<double> > vals(types.size());
  for (size_t i = 0; i < types.size(); i++) vals[i].reserve(n);
  int cnt = 0;
  while (p.Tokenize(s)) {
    if (s.empty()) continue;
    if (s[0] == ":EndMultiData") break;
    ExitGracefullyIf(s.size() != types.size(), "CTimeSeries:ParseMultiple: wrong number of columns in :MultiData data. File: " + p.filename());
    ExitGracefullyIf(cnt >= n, "CTimeSeries::ParseMultiple: Bad number of time series points. File: " + p.filename());
    for (size_t i = 0; i < types.size(); i++) vals[i].push_back(ts_s_to_d(s[i].c_str()));
    cnt++;
  }
  ExitGracefullyIf(cnt != n, "CTimeSeries::ParseMultiple: Insufficient number of time series points. File: " + p.filename());
  static const char *nm[F_NTYPES] = {"PRECIP", "RAINFALL", "SNOWFALL", "TEMP_AVE", "TEMP_DAILY_MIN", "TEMP_DAILY_MAX", "TEMP_DAILY_AVE", "PET", "OW_PET", "POTENTIAL_MELT"};
  for (size_t i = 0; i < types.size(); i++) g->AddTimeSeries(new CTimeSeries(nm[types[i]], start_day, start_yr, interval, vals[i]), types[i]);
}
static void ParseData(CParser &p, CGauge *g, const std::vector<std::string> &hdr, const optStruct &Options) {
  forcing_type f = ForcingFromString(hdr[1]);
  ExitGracefullyIf(f == F_NTYPES, "ParseTimeSeriesFile: :Data type " + hdr[1] + " is not supported by the B200 build");
  double start_day, interval; int start_yr, n;
  ReadSeriesHeader(p, Options, start_day, start_yr, interval, n);
  std::vector<double> v; v.reserve(n);
  std::vector<std::string> s;
  while (p.Tokenize(s)) {
    if (s.empty()) continue;
    if (s[0] == ":EndData") break;
    for (size_t i = 0; i < s.size(); i++) v.push_back(ts_s_to_d(s[i].c_str()));
  }
  ExitGracefullyIf((int)v.size() != n, "CTimeSeries::Parse: wrong number of data points in :Data block. File: " + p.filename());
  g->AddTimeSeries(new CTimeSeries(hdr[1], start_day, start_yr, interval, v), f);
}
static void ParseGaugeBody(CParser &p, CGauge *g, const optStruct &Options, const std::string &thisfile) {
  std::vector<std::string> s;
  while (p.Tokenize(s)) {
    if (s.empty()) continue;
    const std::string &c = s[0];
    if (c == ":EndGauge") return;
    if (c == ":Latitude") g->latitude = s_to_d(s[1]);
    else if (c == ":Longitude") g->longitude = s_to_d(s[1]);
    else if (c == ":Elevation") g->elevation = s_to_d(s[1]);
    else if (c == ":RainCorrection") g->rainfall_corr = s_to_d(s[1]);
    else if (c == ":SnowCorrection") g->snowfall_corr = s_to_d(s[1]);
    else if (c == ":TemperatureCorrection") g->temperature_corr = s_to_d(s[1]);
    else if (c == ":MonthlyAveEvaporation") { for (int m = 0; m < 12; m++) g->aAvePET[m] = s_to_d(s[1 + m]); }
    else if (c == ":MonthlyAveTemperature") { for (int m = 0; m < 12; m++) g->aAveTemp[m] = s_to_d(s[1 + m]); }
    else if (c == ":MonthlyMinTemperature") { for (int m = 0; m < 12; m++) g->aMinTemp[m] = s_to_d(s[1 + m]); }
    else if (c == ":MonthlyMaxTemperature") { for (int m = 0; m < 12; m++) g->aMaxTemp[m] = s_to_d(s[1 + m]); }
    else if (c == ":MultiData") ParseMultiData(p, g, Options);
    else if (c == ":Data") ParseData(p, g, s, Options);
    else if (c == ":RedirectToFile") {
      std::string f = CorrectForRelativePath(s[1], thisfile);
      CParser q(f);
      ExitGracefullyIf(!q.ok(), "ParseTimeSeriesFile: cannot open redirect file " + f);
      ParseGaugeBody(q, g, Options, f);  // returns at EOF
    } else if (c == ":FileType" || c == ":WrittenBy" || c == ":CreationDate") {
    } else ExitGracefully("ParseTimeSeriesFile: command " + c + " inside :Gauge is not supported by the B200 build");
  }
}
static void SkipBlock(CParser &p, const char *endtok) {
  std::vector<std::string> s;
  while (p.Tokenize(s)) if (!s.empty() && s[0] == endtok) return;
}
// :AssimilateStreamflow [SBID] (.rvt or .rve): flags the subbasin when a continuous HYDROGRAPH observation exists for it
static void AssimilateStreamflowCommand(CModel *pModel, const std::vector<std::string> &s) {
  ExitGracefullyIf(s.size() < 2, ":AssimilateStreamflow needs a subbasin ID");
  const long long SBID = s_to_ll(s[1]);
  const int pb = pModel->GetSubBasinIndex(SBID);
  if (pb < 0) { WriteWarning("ParseTimeSeries::Trying to assimilate streamflow at non-existent subbasin " + std::to_string(SBID)); return; }
  bool ObsExists = false;
  for (size_t i = 0; i < pModel->observed.size(); i++) if (pModel->observed[i].name == "HYDROGRAPH" && pModel->observed[i].loc_ID == SBID) ObsExists = true;
  if (ObsExists) pModel->GetSubBasin(pb)->assimilate = true;
  else WriteWarning("AssimilateStreamflow: no observation time series associated with subbasin " + std::to_string(SBID) + ". Cannot assimilate flow in this subbasin.");
}
static void ParseTimeSeriesFile(CModel *pModel, const optStruct &Options, const std::string &filename, bool top) {
  CParser p(filename);
  ExitGracefullyIf(!p.ok(), "ParseTimeSeriesFile: cannot open " + filename);
  std::vector<std::string> s;
  while (p.Tokenize(s)) {
    if (s.empty()) continue;
    const std::string &c = s[0];
    if (c == ":FileType" || c == ":WrittenBy" || c == ":CreationDate") continue;
    if (c == ":Gauge") {
      CGauge *g = new CGauge(s.size() > 1 ? s[1] : ("Gauge" + std::to_string(pModel->GetNumGauges() + 1)));
      ParseGaugeBody(p, g, Options, filename);
      pModel->AddGauge(g);
      continue;
    }
    if (c == ":RedirectToFile") { ParseTimeSeriesFile(pModel, Options, CorrectForRelativePath(s[1], filename), false); continue; }
    if (c == ":ObservationData") {   // src/ParseTimeSeriesFile.cpp:431-497; used by the EnKF ensemble (HYDROGRAPH observations)
      ExitGracefullyIf(s.size() < 3, "ParseTimeSeriesFile: :ObservationData needs [data type] [SBID or HRUID]");
      const std::string dtype = s[1];
      const long long loc = s_to_ll(s[2]);
      double start_day, interval; int start_yr, n;
      ReadSeriesHeader(p, Options, start_day, start_yr, interval, n);
      if (dtype == "HYDROGRAPH") {   // hydrographs are stored period-ending: shifted by one model step (src/TimeSeries.cpp:1040-1046)
        start_day += Options.timestep;
        int leap = IsLeapYear(start_yr, Options.calendar) ? 1 : 0;
        if (start_day >= 365 + leap) { start_day -= 365 + leap; start_yr++; }
      }
      std::vector<double> v; v.reserve(n);
      std::vector<std::string> t;
      while (p.Tokenize(t)) {
        if (t.empty()) continue;
        if (t[0] == ":EndObservationData") break;
        for (size_t i = 0; i < t.size(); i++) v.push_back(ts_s_to_d(t[i].c_str()));
      }
      ExitGracefullyIf((int)v.size() != n, "CTimeSeries::Parse: wrong number of data points in :ObservationData block. File: " + p.filename());
      if (dtype == "HYDROGRAPH" && pModel->GetSubBasinIndex(loc) < 0) { WriteWarning("ParseTimeSeries:: Invalid subbasin ID in observation hydrograph time series. Will be ignored"); continue; }
      observed_ts o; o.name = dtype; o.loc_ID = loc; o.pTS = new CTimeSeries(dtype, start_day, start_yr, interval, v);
      pModel->observed.push_back(o);
      continue;
    }
    if (c == ":BasinInflowHydrograph" || c == ":BasinInflowHydrograph2") {   // src/ParseTimeSeriesFile.cpp:582-628 (series parsed as not pulse-based)
      ExitGracefullyIf(s.size() < 2, "ParseTimeSeriesFile: " + c + " needs a subbasin ID");
      const long long SBID = s_to_ll(s[1]);
      const std::string endtag = (c == ":BasinInflowHydrograph") ? ":EndBasinInflowHydrograph" : ":EndBasinInflowHydrograph2";
      double start_day, interval; int start_yr, n;
      ReadSeriesHeader(p, Options, start_day, start_yr, interval, n);
      std::vector<double> v; v.reserve(n);
      std::vector<std::string> t;
      while (p.Tokenize(t)) {
        if (t.empty()) continue;
        if (t[0] == endtag) break;
        for (size_t i = 0; i < t.size(); i++) v.push_back(ts_s_to_d(t[i].c_str()));
      }
      ExitGracefullyIf((int)v.size() != n, "CTimeSeries::Parse: wrong number of data points in " + c + " block. File: " + p.filename());
      const int pb = pModel->GetSubBasinIndex(SBID);
      if (pb < 0) { WriteWarning(c + " Subbasin " + std::to_string(SBID) + " not in model, cannot set inflow hydrograph"); continue; }
      CTimeSeries *ts = new CTimeSeries("Inflow_Hydrograph_" + std::to_string(SBID), start_day, start_yr, interval, v);
      CSubBasin *pSB = pModel->GetSubBasin(pb);
      if (c == ":BasinInflowHydrograph") { delete pSB->pInflowHydro; pSB->pInflowHydro = ts; }
      else { delete pSB->pInflowHydro2; pSB->pInflowHydro2 = ts; }
      continue;
    }
    if (c == ":ReservoirExtraction" || c == ":ReservoirDemand") {   // src/ParseTimeSeriesFile.cpp:627-667 (pulse-based series, [m3/s])
      ExitGracefullyIf(s.size() < 2, "ParseTimeSeriesFile: " + c + " needs a subbasin ID");
      const long long SBID = s_to_ll(s[1]);
      const std::string endtag = (c == ":ReservoirExtraction") ? ":EndReservoirExtraction" : ":EndReservoirDemand";
      double start_day, interval; int start_yr, n;
      ReadSeriesHeader(p, Options, start_day, start_yr, interval, n);
      std::vector<double> v; v.reserve(n);
      std::vector<std::string> t;
      // CTimeSeries::Parse (src/TimeSeries.cpp:700-760) reads lines until it has n values and then one more line as the :End command,
      // whatever that line says (LOTW's extraction files end with a run-together ":EndObservationData:EndReservoirExtraction")
      while ((int)v.size() < n && p.Tokenize(t)) {
        if (t.empty()) continue;
        ExitGracefullyIf(t[0][0] == ':', "CTimeSeries::Parse: wrong number of data points in " + c + " block. File: " + p.filename());
        for (size_t i = 0; i < t.size(); i++) v.push_back(ts_s_to_d(t[i].c_str()));
      }
      ExitGracefullyIf((int)v.size() != n, "CTimeSeries::Parse: wrong number of data points in " + c + " block. File: " + p.filename());
      do { if (!p.Tokenize(t)) break; } while (t.empty());   // the :End... line
      (void)endtag;
      const int pb = pModel->GetSubBasinIndex(SBID);
      if (pb < 0 || pModel->GetSubBasin(pb)->pReservoir == NULL) {
        WriteWarning(":ReservoirDemand Subbasin " + std::to_string(SBID) + " not in model or doesn't have reservoir, cannot set Reservoir Demand");
        continue;
      }
      CReservoir *pRes = pModel->GetSubBasin(pb)->pReservoir;
      // the reference keeps a list; its RouteWater uses the LAST demand only while the mass balance adds all of them (src/Reservoir.cpp:1643-1652,
      // 1316-1323) -- one series per reservoir is what is supported here
      ExitGracefullyIf(pRes->_pDemandTS != NULL, "ParseTimeSeriesFile: more than one :ReservoirExtraction series for subbasin " + std::to_string(SBID) + " is not supported by the B200 build");
      pRes->_pDemandTS = new CTimeSeries("ResDemand_" + std::to_string(SBID), start_day, start_yr, interval, v);
      continue;
    }
    if (c == ":AssimilateStreamflow") {   // src/ParseTimeSeriesFile.cpp:1225-1248
      AssimilateStreamflowCommand(pModel, s);
      continue;
    }
    if (c == ":ObservationWeights") { SkipBlock(p, ":EndObservationWeights"); continue; }
    ExitGracefully("ParseTimeSeriesFile: command " + c + " (line " + std::to_string(p.line()) + " of " + filename + ") is not supported by the B200 build");
  }
  (void)top;
}

// ================================================================================================ .rve
static void ParseEnsembleFile(CModel *pModel, optStruct &Options) {
  if (Options.rve_filename.empty()) return;
  CParser p(Options.rve_filename);
  if (!p.ok()) { ExitGracefullyIf(Options.ensemble != ENSEMBLE_NONE, "ParseEnsembleFile: cannot open " + Options.rve_filename); return; }
  std::vector<std::string> s;
  while (p.Tokenize(s)) {
    if (s.empty()) continue;
    const std::string &c = s[0];
    if (c == ":FileType" || c == ":WrittenBy" || c == ":CreationDate" || c == ":OutputDirectoryFormat" || c == ":RunNameFormat") continue;
    if (c == ":ParameterDistributions") {  // src/ParseEnsembleFile.cpp:150-215
      while (p.Tokenize(s)) {
        if (s.empty() || s[0] == ":Attributes" || s[0] == ":Units") continue;
        if (s[0] == ":EndParameterDistributions") break;
        // [param_name] [class_type] [class_name] [base_value] [dist_type] [distpar1] [distpar2] {distpar3}   (src/ParseEnsembleFile.cpp:168-220)
        ExitGracefullyIf(s.size() < 7, "Parse Ensemble File: incorrect number of terms in ParameterDistributions command.");
        param_dist d; d.param_name = s[0];
        const std::string &ct = s[1];
        if (ct == "SOIL") d.param_class = CLASS_SOIL; else if (ct == "VEGETATION") d.param_class = CLASS_VEGETATION;
        else if (ct == "LANDUSE") d.param_class = CLASS_LANDUSE; else if (ct == "TERRAIN") d.param_class = CLASS_TERRAIN;
        else if (ct == "GLOBALS") d.param_class = CLASS_GLOBAL;
        else { ExitGracefully("ParseEnsembleFile: invalid parameter class in :ParameterDistributions command"); d.param_class = CLASS_GLOBAL; }
        d.class_group = s[2];
        d.default_val = s_to_d(s[3]);
        bool fix = false;
        const std::string &dt = s[4];
        if (dt == "DIST_UNIFORM") d.distribution = 0;
        else if (dt == "DIST_NORMAL") d.distribution = 1;
        else if (dt == "DIST_LOGNORMAL") d.distribution = 2;
        else if (dt == "DIST_LOGNORMAL2") { d.distribution = 2; fix = true; }
        else if (dt == "DIST_GAMMA") d.distribution = 3;
        else { ExitGracefully("ParseEnsembleFile: invalid distribution type in :ParameterDistributions command"); d.distribution = 0; }
        d.distpar[0] = s_to_d(s[5]); d.distpar[1] = s_to_d(s[6]); d.distpar[2] = (s.size() > 7) ? s_to_d(s[7]) : 0.0;
        if (fix) {  // mean / std of x -> mu / std of ln(x)
          const double mean = d.distpar[0], sd = d.distpar[1];
          d.distpar[0] = log(mean * mean) - log(sqrt(mean * mean + sd * sd));
          d.distpar[1] = log(1.0 + sd * sd / mean / mean);
        }
        pModel->param_dists.push_back(d);
      }
      continue;
    }
    if (c == ":ObjectiveFunction") {   // :ObjectiveFunction [SBID] [Diagnostic] {Period}  (src/ParseEnsembleFile.cpp:232-260)
      ExitGracefullyIf(s.size() < 3, "ParseEnsembleFile: :ObjectiveFunction needs [SBID] [Diagnostic]");
      const int diag = (int)StringToDiagnostic(s[2]);
      ExitGracefullyIf(diag == (int)DIAG_UNRECOGNIZED, "ParseEnsembleFile::unknown diagnostic in :ObjectiveFunction command.");
      if (Options.ensemble == ENSEMBLE_DDS) { pModel->calib_SBID = s_to_ll(s[1]); pModel->calib_obj = diag; pModel->calib_period = (s.size() >= 4) ? s[3] : "ALL"; }
      else WriteWarning(":ObjectiveFunction command will be ignored; no calibration method specified.");
      continue;
    }
    const bool is_enkf = (Options.ensemble == ENSEMBLE_ENKF);
    if (c == ":EnKFMode") {   // src/ParseEnsembleFile.cpp:431-450
      ExitGracefullyIf(s.size() < 2, "ParseEnsembleFile: :EnKFMode needs a mode");
      EnKF_mode mode = ENKF_UNSPECIFIED;
      if (s[1] == "ENKF_SPINUP") mode = ENKF_SPINUP; else if (s[1] == "ENKF_CLOSED_LOOP") mode = ENKF_CLOSED_LOOP;
      else if (s[1] == "ENKF_OPEN_LOOP") mode = ENKF_OPEN_LOOP; else if (s[1] == "ENKF_FORECAST") mode = ENKF_FORECAST;
      else if (s[1] == "ENKF_OPEN_FORECAST") mode = ENKF_OPEN_FORECAST;
      else ExitGracefully("ParseEnsembleFile: EnKFMode-  invalid mode specified");
      if (is_enkf) pModel->enkf.mode = mode;
      continue;
    }
    if (c == ":WindowSize" || c == ":DataHorizon") { if (is_enkf) pModel->enkf.window_size = s_to_i(s[1]); continue; }
    if (c == ":ForcingPerturbation" || c == ":ObservationErrorModel") {   // src/ParseEnsembleFile.cpp:263-303, 375-420
      if (!is_enkf) { WriteWarning(c + " command will be ignored; only valid for EnKF ensemble simulation."); continue; }
      ExitGracefullyIf(s.size() < 6, "ParseEnsembleFile: " + c + " needs [type] [distrib] [par1] [par2] [adj_type]");
      int distrib = 1; bool fix = false;
      if (s[2] == "DIST_UNIFORM") distrib = 0; else if (s[2] == "DIST_NORMAL") distrib = 1; else if (s[2] == "DIST_LOGNORMAL") distrib = 2;
      else if (s[2] == "DIST_LOGNORMAL2") { distrib = 2; fix = true; } else if (s[2] == "DIST_GAMMA") distrib = 3;
      else ExitGracefully("ParseEnsembleFile: invalid distribution type in " + c + " command");
      double distpars[3] = {s_to_d(s[3]), s_to_d(s[4]), 0.0};
      if (fix) {
        const double mean = distpars[0], sd = distpars[1];
        distpars[0] = log(mean * mean) - log(sqrt(mean * mean + sd * sd));
        distpars[1] = log(1.0 + sd * sd / mean / mean);
        ExitGracefullyIf(mean <= 0.0, "ParseEnsembleFile: invalid (negative or zero) mean of lognormal distribution in " + c + " command");
      }
      adjustment adj = ADJ_ADDITIVE;
      if (s[5] == "ADDITIVE") adj = ADJ_ADDITIVE; else if (s[5] == "MULTIPLICATIVE") adj = ADJ_MULTIPLICATIVE;
      else ExitGracefully("ParseEnsembleFile: invalid perturbation adjustment type in " + c + " command");
      if (c == ":ForcingPerturbation") {
        forcing_type ftyp = ForcingFromString(s[1]);
        ExitGracefullyIf(ftyp == F_NTYPES, "ParseEnsembleFile: forcing type " + s[1] + " in :ForcingPerturbation is not supported by the B200 build");
        int kk = DOESNT_EXIST;
        if (s.size() >= 7) { kk = pModel->FindHRUGroup(s[6]); ExitGracefullyIf(kk == DOESNT_EXIST, "ParseEnsembleFile: :ForcingPerturbation HRU group does not exist"); }
        pModel->AddForcingPerturbation(ftyp, distrib, distpars, kk, adj);
      } else {
        int lay = DOESNT_EXIST;
        obs_perturb op; op.state = (s[1] == "STREAMFLOW") ? (sv_type)RVN_SV_NTYPES : pModel->StringToSVType(s[1], lay, true);
        op.distribution = distrib; op.kk = DOESNT_EXIST; op.adj_type = adj;
        for (int i = 0; i < 3; i++) op.distpar[i] = distpars[i];
        pModel->enkf.obs_perturbations.push_back(op);
      }
      continue;
    }
    if (c == ":AssimilatedState") {   // src/ParseEnsembleFile.cpp:305-345
      if (!is_enkf) { WriteWarning(":AssimilatedState command will be ignored; only valid for EnKF ensemble simulation."); continue; }
      ExitGracefullyIf(s.size() < 3, "ParseEnsembleFile: :AssimilatedState needs [state] [group]");
      assim_state a; a.layer = DOESNT_EXIST; a.is_streamflow = (s[1] == "STREAMFLOW");
      ExitGracefullyIf(s[1] == "RESERVOIR_STAGE", "ParseEnsembleFile: :AssimilatedState RESERVOIR_STAGE is not supported by the B200 build (no reservoirs)");
      if (a.is_streamflow) {
        a.sv = (sv_type)RVN_SV_NTYPES;
        a.group = pModel->FindSubBasinGroup(s[2]);
        ExitGracefullyIf(a.group == DOESNT_EXIST, "ParseEnsembleFile: :AssimilatedState subbasin/HRU group does not exist");
      } else {
        a.sv = pModel->StringToSVType(s[1], a.layer, true);
        if (pModel->GetStateVarIndex(a.sv, a.layer) == DOESNT_EXIST) {
          WriteWarning("State variable " + s[1] + " does not exist in this model and will be ignored in the :AssimilatedState command"); continue;
        }
        a.group = pModel->FindHRUGroup(s[2]);
        ExitGracefullyIf(a.group == DOESNT_EXIST, "ParseEnsembleFile: :AssimilatedState subbasin/HRU group does not exist");
      }
      pModel->enkf.assim_states.push_back(a);
      continue;
    }
    if (c == ":AssimilateStreamflow") { AssimilateStreamflowCommand(pModel, s); continue; }
    if (c == ":SolutionRunName" || c == ":EnsembleRVCFormat" || c == ":ExtraRVTFilename" || c == ":ForecastRVTFilename") continue;   // file plumbing of the external EnKF driver
    ExitGracefully("ParseEnsembleFile: command " + c + " is not supported by the B200 build");
  }
}

// ================================================================================================ drivers
bool ParseInputFiles(CModel *&pModel, optStruct &Options) {
  pModel = NULL;
  ParseMainInputFile(pModel, Options);
  ParseClassPropertiesFile(pModel, Options);
  ParseHRUPropsFile(pModel, Options);
  ParseTimeSeriesFile(pModel, Options, Options.rvt_filename, true);
  if (Options.ensemble != ENSEMBLE_NONE) ParseEnsembleFile(pModel, Options);
  for (int j = 0; j < pModel->GetNumProcesses(); j++) pModel->GetProcess(j)->Initialize();
  return true;
}

bool ParseInitialConditions(CModel *&pModel, const optStruct &Options) {
  pModel->AllocateInitialState();
  if (Options.rvc_filename.empty()) return true;
  CParser p(Options.rvc_filename);
  if (!p.ok()) return true;  // reference: .rvc is optional (all-zero initial state)
  const int NS = pModel->GetNumStateVars(), nH = pModel->GetNumHRUs();
  std::vector<std::string> s;
  double rvc_tstep = Options.timestep;
  std::map<long long, int> hru_of_id;
  for (int kk = 0; kk < nH; kk++) hru_of_id.insert(std::make_pair(pModel->GetHydroUnit(kk)->ID, kk));
  while (p.Tokenize(s)) {
    if (s.empty()) continue;
    const std::string c = s[0];   // a copy: several blocks below re-tokenise into s
    if (c == ":FileType" || c == ":WrittenBy" || c == ":CreationDate" || c == ":TimeStamp" || c == ":StartTime") continue;
    if (c == ":TimeStep") {   // time step of the run that wrote the file; histories are only taken over at the same step
      rvc_tstep = ParseIntervalToken(s[1], Options.calendar);
      continue;
    }
    if (c == ":BasinStateVariables") {   // src/ParseInitialConditionFile.cpp:539-660 (no reservoirs)
      CSubBasin *pBasin = NULL;
      // values may continue on following lines: pull `need` numbers starting at token i of the current line
      auto pull = [&](size_t i, int need, std::vector<double> &out) {
        out.clear();
        while ((int)out.size() < need) {
          if (i >= s.size()) { if (!p.Tokenize(s)) break; i = 0; if (s.empty()) continue; }
          out.push_back(s_to_d(s[i++]));
        }
      };
      while (p.Tokenize(s)) {
        if (s.empty()) continue;
        if (s[0] == ":EndBasinStateVariables") break;
        if (s[0] == ":BasinIndex") {
          const int pb = pModel->GetSubBasinIndex(s_to_ll(s[1]));
          ExitGracefullyIf(pb < 0, "ParseInitialConditionsFile: bad basin index in .rvc file");
          pBasin = pModel->GetSubBasin(pb);
          continue;
        }
        ExitGracefullyIf(pBasin == NULL, "ParseInitialConditionsFile: basin state given before :BasinIndex");
        ExitGracefullyIf(rvc_tstep != Options.timestep, "ParseInitialConditionsFile: the .rvc file was written with a different time step (history rescaling is not supported by the B200 build)");
        std::vector<double> v;
        if (s[0] == ":ChannelStorage") { if (s.size() >= 2) pBasin->channel_storage = s_to_d(s[1]); }
        else if (s[0] == ":RivuletStorage") { if (s.size() >= 2) pBasin->rivulet_storage = s_to_d(s[1]); }
        else if (s[0] == ":Qout") {   // [nsegs] [aQout x nsegs] [QoutLast]  -> CSubBasin::SetQoutArray (src/SubBasin.cpp:1407-1417)
          if (s.size() > 2) {
            const int nsegs = s_to_i(s[1]);
            pull(2, nsegs + 1, v);
            if (nsegs != pBasin->nSegments) WriteWarning("Number of reach segments in state file and input file are inconsistent in basin " + std::to_string(pBasin->ID) + ". Unable to read in-reach flow initial conditions");
            else if ((int)v.size() == nsegs + 1) { for (int i = 0; i < nsegs; i++) pBasin->aQout[i] = v[i]; pBasin->QoutLast = v[nsegs]; }
          }
        } else if (s[0] == ":Qlat") {   // [n] [aQlatHist x n] [QlatLast] -> SetQlatHist (:1435-1476): also recomputes _Qlocal
          if (s.size() > 2) {
            const int n = s_to_i(s[1]);
            pull(2, n + 1, v);
            if (n != 0 && (int)v.size() == n + 1) {
              pBasin->QlatLast = v[n];
              const int m = (n < pBasin->GetLatHistorySize()) ? n : pBasin->GetLatHistorySize();
              if (n != pBasin->GetLatHistorySize()) WriteWarning("CSubBasin::SetQlatHist: size of lateral flow history differs between current model and initial conditions file. Array will be truncated");
              for (int i = 0; i < m; i++) pBasin->aQlatHist[i] = v[i];
              pBasin->Qlocal = 0.0;
              for (int i = 0; i < pBasin->GetLatHistorySize(); i++) pBasin->Qlocal += pBasin->aUnitHydro[i] * pBasin->aQlatHist[i];
            }
          }
        } else if (s[0] == ":Qin") {   // [n] [aQinHist x n] -> SetQinHist (:1485-1518)
          if (s.size() > 2) {
            const int n = s_to_i(s[1]);
            pull(2, n, v);
            if (n != 0 && (int)v.size() == n) {
              const int m = (n < pBasin->GetInflowHistorySize()) ? n : pBasin->GetInflowHistorySize();
              if (n != pBasin->GetInflowHistorySize()) WriteWarning("CSubBasin::SetQinHist: size of inflow history differs between current model and initial conditions file. Array will be truncated");
              for (int i = 0; i < m; i++) pBasin->aQinHist[i] = v[i];
            }
          }
        } else if (s[0] == ":ResStage") {   // src/ParseInitialConditionFile.cpp:628-637
          if (s.size() >= 3) {
            if (pBasin->pReservoir == NULL) WriteWarning("ParseInitialConditionsFile: reservoir initial conditions for basin without reservoir will be ignored.");
            else pBasin->pReservoir->SetReservoirStage(s_to_d(s[1]), s_to_d(s[2]));
          }
        } else if (s[0] == ":ResFlow") {    // :638-646
          if (s.size() >= 3) {
            if (pBasin->pReservoir == NULL) WriteWarning("ParseInitialConditionsFile: reservoir initial conditions for basin without reservoir will be ignored.");
            else pBasin->pReservoir->SetInitialFlow(s_to_d(s[1]), s_to_d(s[2]), true);
          }
        } else ExitGracefully("ParseInitialConditionsFile: command " + s[0] + " inside :BasinStateVariables is not supported by the B200 build");
      }
      continue;
    }
    if (c == ":InitialReservoirFlow" || c == ":InitialReservoirStage") {   // src/ParseInitialConditionFile.cpp:668-699
      ExitGracefullyIf(s.size() < 3, "ParseInitialConditionsFile: " + c + " needs [SBID] [value]");
      const int pb = pModel->GetSubBasinIndex(s_to_ll(s[1]));
      ExitGracefullyIf(pb < 0, "ParseInitialConditionsFile: bad basin index in " + c + " command (.rvc file)");
      CReservoir *pRes = pModel->GetSubBasin(pb)->pReservoir;
      if (pRes == NULL) { WriteWarning("ParseInitialConditionsFile: bad basin index in " + c + " command (.rvc file), no reservoir exists in basin " + s[1] + ". Command will be ignored."); continue; }
      if (c == ":InitialReservoirFlow") pRes->SetInitialFlow(s_to_d(s[2]), s_to_d(s[2]), false);
      else pRes->SetReservoirStage(s_to_d(s[2]), s_to_d(s[2]));
      continue;
    }
    if (c == ":UniformInitialConditions") {
      int lay; sv_type t = pModel->StringToSVType(s[1], lay, false);
      int i = (t == RVN_SV_NTYPES) ? DOESNT_EXIST : pModel->GetStateVarIndex(t, lay);
      if (i == DOESNT_EXIST) { WriteWarning("Initial conditions specified for state variable not in model (" + s[1] + ")"); continue; }
      for (int k = 0; k < nH; k++) pModel->SetInitialStateVar(k, i, s_to_d(s[2]));
      continue;
    }
    if (c == ":HRUStateVariableTable") {
      std::vector<int> idx;
      while (p.Tokenize(s)) {
        if (s.empty() || s[0] == ":Units") continue;
        if (s[0] == ":EndHRUStateVariableTable") break;
        if (s[0] == ":Attributes") {
          idx.clear();
          for (size_t i = 1; i < s.size(); i++) {
            int lay; sv_type t = pModel->StringToSVType(s[i], lay, false);
            int ix = (t == RVN_SV_NTYPES) ? DOESNT_EXIST : pModel->GetStateVarIndex(t, lay);
            // initial conditions of cumulative precip, evaporation and glacier loss are ignored, left at zero (src/ParseInitialConditionFile.cpp:385-388;
            // Options.glacier_model_on is off in every model the host layer accepts)
            if (t == RVN_SV_ATMOS_PRECIP || t == RVN_SV_ATMOSPHERE || t == RVN_SV_GLACIER_ICE) ix = DOESNT_EXIST;
            idx.push_back(ix);
          }
          continue;
        }
        long long id = s_to_ll(s[0]);
        std::map<long long, int>::const_iterator it = hru_of_id.find(id);
        if (it == hru_of_id.end()) continue;
        const int k = it->second;
        for (size_t i = 0; i < idx.size() && i + 1 < s.size(); i++) if (idx[i] >= 0) pModel->SetInitialStateVar(k, idx[i], s_to_d(s[i + 1]));
      }
      continue;
    }
    if (c == ":InitialConditions") {  // one value per HRU (or a single value for all)
      int lay; sv_type t = pModel->StringToSVType(s[1], lay, false);
      int i = (t == RVN_SV_NTYPES) ? DOESNT_EXIST : pModel->GetStateVarIndex(t, lay);
      std::vector<double> v;
      while (p.Tokenize(s)) {
        if (s.empty()) continue;
        if (s[0] == ":EndInitialConditions") break;
        for (size_t q = 0; q < s.size(); q++) v.push_back(s_to_d(s[q]));
      }
      if (i >= 0) for (int k = 0; k < nH && k < (int)v.size(); k++) pModel->SetInitialStateVar(k, i, v[k]);
      continue;
    }
    if (c == ":BasinInitialConditions") {
      while (p.Tokenize(s)) {
        if (s.empty() || s[0] == ":Attributes" || s[0] == ":Units") continue;
        if (s[0] == ":EndBasinInitialConditions") break;
        int pb = pModel->GetSubBasinIndex(s_to_ll(s[0]));
        ExitGracefullyIf(pb < 0, "ParseInitialConditionsFile: bad basin index in :BasinInitialConditions (.rvc file)");
        pModel->GetSubBasin(pb)->SetQout(s_to_d(s[1]));
      }
      continue;
    }
    ExitGracefully("ParseInitialConditions: command " + c + " is not supported by the B200 build");
  }
  // quality check of the initial state (src/ParseInitialConditionFile.cpp:1050-1080): values above the state-independent
  // capacity (GetStateVarMax with ignorevar=true: SOIL capacity, DEPRESSION dep_max, SNOW_COVER 1) are clipped to it
  for (int k = 0; k < nH; k++) {
    const CHydroUnit *h = pModel->GetHydroUnit(k);
    const CSoilProfile &sp = pModel->soil_profiles[h->soil_profile];
    for (int i = 0; i < NS; i++) {
      double maxv = ALMOST_INF;
      const sv_type t = pModel->GetStateVarType(i);
      if (t == RVN_SV_SOIL) {
        const int m = pModel->GetStateVarLayer(i);
        maxv = (m < sp.nHorizons) ? sp.thickness[m] * MM_PER_METER * pModel->soil_classes[sp.soil_class[m]].S.v[RVN_S_CAP_RATIO] : 0.0;
      } else if (t == RVN_SV_DEPRESSION) {
        const double dm = pModel->lu_classes[h->lu_class].S.v[RVN_LU_DEP_MAX];
        if (dm >= 0) maxv = dm;
      } else if (t == RVN_SV_SNOW_COVER) maxv = 1.0;
      maxv = rvh_max(maxv, 0.0);
      const double v = pModel->initial_state[(size_t)k * NS + i];
      if ((v - maxv) > 1e-8) {
        WriteWarning("maximum state variable limit exceeded in initial conditions for " + pModel->GetStateVarName(i) + " (in HRU " + std::to_string(h->ID) + ") in .rvc file");
        pModel->SetInitialStateVar(k, i, maxv);
      }
    }
  }
  return true;
}
