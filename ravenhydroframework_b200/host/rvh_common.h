// rvh_common.h -- shared declarations of the Raven-surface host layer of the B200 build.
//
// The host layer keeps the reference's source-level surface for the hot path:
// optStruct / time_struct, CModel, CHydroUnit, CSubBasin, CHydroProcessABC, MassEnergyBalance(),
// CModel::UpdateHRUForcingFunctions(), the .rvi/.rvh/.rvp/.rvt/.rvc readers and the BMI facade,
// and compiles what it parsed into the flat rvn_model_desc of include/rvn_gpu.h.  No per-timestep
// arithmetic happens here: the time loop runs on the GPU behind the C ABI.
#ifndef RVH_COMMON_H
#define RVH_COMMON_H
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>
#include "rvn_gpu.h"

// ---- constants (values are the reference's physical/sentinel constants, src/RavenInclude.h:113-273) ----
const int    DOESNT_EXIST       = -1;
const double ALMOST_INF         = 1e99;
const double PRETTY_SMALL       = 1e-8;
const double REAL_SMALL         = 1e-12;
const double TIME_CORRECTION    = 0.0001;
const double DAYS_PER_YEAR      = 365.25;
const double SEC_PER_DAY        = 86400.0;
const double HR_PER_DAY         = 24.0;
const double MM_PER_METER       = 1000.0;
const double M_PER_KM           = 1000.0;
const double M2_PER_KM2         = 1e6;
const double RVH_PI             = 3.1415926535898;  // reference PI literal (src/RavenInclude.h)
const double AUTO_COMPUTE       = -11111.1;
const double NOT_SPECIFIED      = -33333.3;
const double USE_TEMPLATE_VALUE = -55555.5;
const double RAV_BLANK_DATA     = -1.2345;
const int    MAX_CONVOL_STORES  = RVN_CONV_STORES;
const int    MAX_RIVER_SEGS     = RVN_MAX_RIVER_SEGS;

typedef rvn_sv_type sv_type;

// Error behaviour of the reference: ExitGracefully(msg,code) (src/RavenInclude.h:311-336).  In library
// (BMI) mode the reference only prints; standalone it exits.  The host layer throws instead so that
// the C API / BMI facade can surface the message; the CLI front end catches and exits.
struct RavenError : public std::runtime_error {
  explicit RavenError(const std::string &m) : std::runtime_error(m) {}
};
void ExitGracefully(const std::string &statement);
inline void ExitGracefullyIf(bool cond, const std::string &statement) { if (cond) ExitGracefully(statement); }
void WriteWarning(const std::string &warn, bool noisy = false);
const std::vector<std::string> &RavenWarnings();

// ---- reference time_struct (src/RavenInclude.h) --------------------------------------------------
struct time_struct {
  std::string date_string;
  double model_time;
  double julian_day;
  int    day_of_month;
  int    month;
  int    year;
  bool   leap_yr;
  bool   day_changed;
  time_struct() : model_time(0), julian_day(0), day_of_month(1), month(1), year(1900), leap_yr(false), day_changed(true) {}
};

enum catchment_route { ROUTE_DUMP, ROUTE_TRI_CONVOLUTION, ROUTE_GAMMA_CONVOLUTION, ROUTE_DELAYED_FIRST_ORDER };
enum interp_method { INTERP_NEAREST_NEIGHBOR, INTERP_AVERAGE_ALL, INTERP_INVERSE_DISTANCE, INTERP_FROM_FILE, INTERP_INVERSE_DISTANCE_ELEVATION };
enum month_interp { MONTHINT_UNIFORM, MONTHINT_LINEAR_FOM, MONTHINT_LINEAR_MID, MONTHINT_LINEAR_21 };
enum ensemble_type { ENSEMBLE_NONE, ENSEMBLE_MONTECARLO, ENSEMBLE_DDS, ENSEMBLE_ENKF };

// ---- reference optStruct (src/RavenInclude.h:1065-1223), the subset that steers the hot path ----
struct optStruct {
  double julian_start_day;
  int    julian_start_year;
  double duration;
  double timestep;
  int    calendar;
  std::string rvi_filename, rvh_filename, rvp_filename, rvt_filename, rvc_filename, rve_filename;
  std::string output_dir, main_output_dir, run_name, working_dir;
  int    sol_method;            // rvn_sol_method: ORDERED_SERIES (0), EULER (1) or ITERATED_HEUN (2)
  double convergence_crit = 0.01; int max_iterations = 30;   // ITERATED_HEUN (src/ParseInput.cpp:242-243, 744-745)
  rvn_rainsnow rainsnow;
  rvn_potmelt  pot_melt;
  rvn_pet      evaporation, ow_evaporation, evap_infill, ow_evap_infill;
  rvn_orocorr  orocorr_temp, orocorr_precip, orocorr_PET;
  int    air_pressure, rel_humidity, SW_radiation, SW_cloudcovercorr;   // rvn_airpress / rvn_relhum / rvn_sw_rad / rvn_sw_cloudcorr
  int    wind_velocity;         // rvn_windvel
  int    interception_factor;   // rvn_icept
  int    LW_radiation = 0, SW_canopycorr = 0;   // rvn_options::lw_radiation / sw_canopycorr
  std::string unsupported_radiation;   // first radiation method command the B200 build cannot honour (fatal only if net radiation is consumed)
  bool keep_flux_balance = false;   // keep the per-connection cumulative fluxes (CModel::_aCumulativeBal) on the device: rvn_options::keep_flux_balance
  int subdaily = 0;   // :SubdailyMethod: 0 SUBDAILY_NONE, 1 SUBDAILY_SIMPLE
  std::string unsupported_atmos;   // first atmospheric method command the B200 build cannot honour (fatal only if a consumer needs it)
  rvn_routing  routing;
  catchment_route catchment_routing;
  interp_method interpolation;
  std::string  interp_file;
  month_interp month_interp_method;
  int    num_soillayers;
  bool   suppressCompetitiveET, snow_suppressPET, allow_soil_overfill, direct_evap;
  bool   silent, noisy, benchmarking;
  bool   assimilate_flow = false; double assimilation_start = -1.0;   // :AssimilateStreamflow / :AssimilationStartTime (src/ParseInput.cpp:335-339,1795-1818)
  double assimilation_fact = 1.0, assim_upstream_decay = 0.0, assim_time_decay = 0.2;   // globals ASSIMILATION_FACT / ASSIM_UPSTREAM_DECAY [1/km] / ASSIM_TIME_DECAY (src/GlobalParams.cpp:388-390)
  ensemble_type ensemble;
  unsigned int random_seed;
  bool   random_seed_set;
  optStruct();
};

// ---- small utilities (rvh_util.cpp) -------------------------------------------------------------
bool   IsLeapYear(int year, int calendar);
void   JulianConvert(double model_time, double start_date, int start_year, int calendar, time_struct &tt);
time_struct DateStringToTimeStruct(const std::string &sDate, const std::string &sTime, int calendar);
double TimeDifference(double jul_day1, int year1, double jul_day2, int year2, int calendar);
double DayAngle(double julian_day, int year, int calendar);
double SolarDeclination(double day_angle);                         // src/Radiation.cpp:484-497
double OpticalAirMass(double latrad, double declin, double day_length, double t_sol, bool avg_daily);   // src/Radiation.cpp:875-967
double EccentricityCorr(double day_angle);                         // src/Radiation.cpp:505-515
double CalcETRadiation2(double latrad, double aspect, double declin, double ecc, double slope, double t1, double t2, bool avg_daily);  // src/Radiation.cpp:637-822
double gamma2(double x);
double IncompleteGamma(double x, double a);
double GammaCumDist(double t, double alpha, double beta);
double TriCumDist(double t, double tc, double tp);
double DiffusiveWaveUH(int n, double L, double v, double D, double dt);   // src/CommonFunctions.cpp:1796-1806
double InterpolateMo(const double aVal[12], const time_struct &tt, month_interp method, const optStruct &Options);
inline double rvh_max(double r, double l) { return (r >= l) ? r : l; }   // tie -> first argument (src/RavenInclude.h:1456)
inline double rvh_min(double r, double l) { return (r <= l) ? r : l; }

// tokenizer in the spirit of the reference CParser (src/ParseLib.h): whitespace/comma/tab separated,
// '#' and '*' start comments, ':RedirectToFile' handled by the callers.
class CParser {
public:
  explicit CParser(const std::string &filename);
  bool ok() const { return _ok; }
  // returns false at EOF; empty/comment lines yield an empty token list
  bool Tokenize(std::vector<std::string> &tokens);
  bool Include(const std::string &filename);   // splice another file in at the current position (:RedirectToFile)
  int  line() const { return _line; }
  const std::string &filename() const { return _filename; }
private:
  std::vector<std::string> _lines;
  size_t _pos;
  int _line;
  bool _ok;
  std::string _filename;
};
std::string DirName(const std::string &path);
std::string CorrectForRelativePath(const std::string &filename, const std::string &relfile);
void LatLonToUTMXY(double lat_deg, double lon_deg, int zone, double &x, double &y);   // src/UTM_to_LatLong.cpp:216-227
std::string StringToUppercase(const std::string &s);
bool   IsComment(const std::string &tok);
double s_to_d(const std::string &s);
int    s_to_i(const std::string &s);
long long s_to_ll(const std::string &s);
#endif
