// rvh_util.cpp -- tokenizer, calendar arithmetic and the special functions the host needs to hoist
// per-(class,date) work out of the per-HRU-per-step loop.  Formulas follow the reference so that the
// hoisted values are the same doubles the reference recomputes for every HRU and step:
//   JulianConvert / IsLeapYear / TimeDifference  src/CommonFunctions.cpp:273-380, 673-692
//   gamma2 / IncompleteGamma / GammaCumDist      src/CommonFunctions.cpp:1588-1688
//   TriCumDist                                   src/CommonFunctions.cpp:1699-1715
//   InterpolateMo                                src/CommonFunctions.cpp:865-925
//   CRadiation::DayAngle                         src/Radiation.cpp:423-430
#include "rvh_common.h"
#include <fstream>
#include <sstream>
#include <algorithm>

static std::vector<std::string> g_warnings;
void ExitGracefully(const std::string &statement) { throw RavenError(statement); }
void WriteWarning(const std::string &warn, bool noisy) {
  g_warnings.push_back(warn);
  if (noisy) fprintf(stderr, "WARNING: %s\n", warn.c_str());
}
const std::vector<std::string> &RavenWarnings() { return g_warnings; }

optStruct::optStruct()
    : julian_start_day(0), julian_start_year(1666), duration(365), timestep(1.0), calendar(0), sol_method(0),
      rainsnow(RVN_RAINSNOW_DINGMAN), pot_melt(RVN_POTMELT_DEGREE_DAY), evaporation(RVN_PET_HARGREAVES_1985),
      ow_evaporation(RVN_PET_HARGREAVES_1985), evap_infill(RVN_PET_HARGREAVES_1985), ow_evap_infill(RVN_PET_HARGREAVES_1985),
      orocorr_temp(RVN_OROCORR_NONE), orocorr_precip(RVN_OROCORR_NONE), orocorr_PET(RVN_OROCORR_NONE),
      air_pressure(RVN_AIRPRESS_BASIC), rel_humidity(RVN_RELHUM_CONSTANT), SW_radiation(RVN_SW_RAD_DEFAULT), SW_cloudcovercorr(RVN_SW_CLOUD_CORR_NONE),
      wind_velocity(RVN_WINDVEL_CONSTANT), interception_factor(RVN_ICEPT_USER),
      routing(RVN_ROUTE_STORAGECOEFF), catchment_routing(ROUTE_DUMP), interpolation(INTERP_NEAREST_NEIGHBOR),
      month_interp_method(MONTHINT_LINEAR_MID), num_soillayers(-1), suppressCompetitiveET(false), snow_suppressPET(false),
      allow_soil_overfill(false), direct_evap(false), silent(false), noisy(false), benchmarking(false),
      ensemble(ENSEMBLE_NONE), random_seed(0), random_seed_set(false) {}

// ---------------------------------------------------------------------------------- strings / files
std::string StringToUppercase(const std::string &s) {
  std::string r = s;
  for (size_t i = 0; i < r.size(); i++) r[i] = (char)toupper((unsigned char)r[i]);
  return r;
}
bool IsComment(const std::string &t) { return !t.empty() && (t[0] == '#' || t[0] == '*'); }
double s_to_d(const std::string &s) { return atof(s.c_str()); }
int s_to_i(const std::string &s) { return atoi(s.c_str()); }
long long s_to_ll(const std::string &s) { return atoll(s.c_str()); }
std::string DirName(const std::string &path) {
  size_t p = path.find_last_of("/\\");
  return (p == std::string::npos) ? std::string("") : path.substr(0, p + 1);
}
std::string CorrectForRelativePath(const std::string &filename, const std::string &relfile) {
  if (!filename.empty() && (filename[0] == '/' || (filename.size() > 1 && filename[1] == ':'))) return filename;
  return DirName(relfile) + filename;
}

CParser::CParser(const std::string &filename) : _pos(0), _line(0), _ok(false), _filename(filename) {
  std::ifstream in(filename.c_str());
  if (!in.good()) return;
  _ok = true;
  std::string l;
  while (std::getline(in, l)) _lines.push_back(l);
}
bool CParser::Include(const std::string &filename) {   // :RedirectToFile: the lines of `filename` are read next, then the rest of this file
  std::ifstream in(filename.c_str());
  if (!in.good()) return false;
  std::vector<std::string> add;
  std::string l;
  while (std::getline(in, l)) add.push_back(l);
  _lines.insert(_lines.begin() + _pos, add.begin(), add.end());
  return true;
}
bool CParser::Tokenize(std::vector<std::string> &tokens) {
  tokens.clear();
  if (_pos >= _lines.size()) return false;
  const std::string &l = _lines[_pos++];
  _line++;
  std::string cur;
  for (size_t i = 0; i <= l.size(); i++) {
    char c = (i < l.size()) ? l[i] : ' ';
    if (c == ' ' || c == '\t' || c == ',' || c == '\r' || c == '\n') {
      if (!cur.empty()) { tokens.push_back(cur); cur.clear(); }
    } else if ((c == '#' || c == '*') && cur.empty()) {
      break;  // comment to end of line
    } else {
      cur.push_back(c);
    }
  }
  if (!cur.empty()) tokens.push_back(cur);
  return true;
}

// ---------------------------------------------------------------------------------- calendar
bool IsLeapYear(int year, int calendar) {
  (void)calendar;  // only the proleptic Gregorian calendar (reference default) is supported
  return (((year % 4 == 0) && (year % 100 != 0)) || (year % 400 == 0));
}

void JulianConvert(double model_time, double start_date, int start_year, int calendar, time_struct &tt) {
  // snap daily / hourly round-off exactly as the reference does before splitting the date
  if ((model_time - floor(model_time)) > (1 - TIME_CORRECTION)) model_time = floor(model_time + TIME_CORRECTION);
  if (model_time * HR_PER_DAY - floor(model_time * HR_PER_DAY) > (1.0 - TIME_CORRECTION) * HR_PER_DAY)
    model_time = floor(HR_PER_DAY * (model_time + TIME_CORRECTION)) / HR_PER_DAY;

  double ddate = start_date + model_time;
  int dyear = start_year;
  int leap = IsLeapYear(dyear, calendar) ? 1 : 0;
  while (ddate >= (365 + leap)) {
    ddate -= (365.0 + leap);
    dyear++;
    leap = IsLeapYear(dyear, calendar) ? 1 : 0;
  }
  static const int mdays[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
  int dmonth = 1;
  double days = 31, sum = 31 - TIME_CORRECTION;
  for (int m = 1; m < 12; m++) {
    if (ddate >= sum) {
      dmonth += 1;
      days = mdays[m] + ((m == 1) ? leap : 0);
      sum += days;
    }
  }
  double dday = ddate - sum + days;
  tt.model_time = model_time;
  tt.julian_day = ddate;
  tt.day_of_month = (int)(ceil(dday + REAL_SMALL));
  if (tt.day_of_month == 0) tt.day_of_month = 1;
  tt.month = dmonth;
  tt.year = dyear;
  tt.day_changed = false;
  if ((model_time <= PRETTY_SMALL) || (tt.julian_day - floor(tt.julian_day + TIME_CORRECTION) < 0.001)) tt.day_changed = true;
  char out[64];
  snprintf(out, sizeof(out), "%4.4d-%2.2i-%2.2d", dyear, tt.month, tt.day_of_month);
  tt.date_string = out;
  tt.leap_yr = IsLeapYear(tt.year, calendar);
}

time_struct DateStringToTimeStruct(const std::string &sDate, const std::string &sTimeIn, int calendar) {
  // yyyy-mm-dd (or yyyy/mm/dd) and hh:mm:ss[.ss]  (reference src/CommonFunctions.cpp:404-470)
  time_struct tt;
  std::string sTime = sTimeIn;
  ExitGracefullyIf(sDate.length() != 10, "DateStringToTimeStruct: Invalid date format used: " + sDate);
  if (sTime.length() == 7) sTime = "0" + sTime;  // h:mm:ss
  ExitGracefullyIf(sTime.length() < 8, "DateStringToTimeStruct: Invalid time format used: " + sTime);
  tt.year = s_to_i(sDate.substr(0, 4));
  tt.month = s_to_i(sDate.substr(5, 2));
  tt.day_of_month = s_to_i(sDate.substr(8, 2));
  ExitGracefullyIf(tt.month > 12 || tt.month < 1, "DateStringToTimeStruct: Invalid month in date " + sDate);
  static const int mdays[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
  int leap = IsLeapYear(tt.year, calendar) ? 1 : 0;
  tt.julian_day = tt.day_of_month - 1;
  for (int m = 0; m < tt.month - 1; m++) tt.julian_day += mdays[m] + ((m == 1) ? leap : 0);
  int hr = s_to_i(sTime.substr(0, 2));
  int mn = s_to_i(sTime.substr(3, 2));
  double sec = s_to_d(sTime.substr(6));
  tt.julian_day += (double)hr / HR_PER_DAY;
  tt.julian_day += (double)mn / 1440.0;
  tt.julian_day += sec / SEC_PER_DAY;
  tt.model_time = 0.0;
  tt.leap_yr = leap != 0;
  tt.date_string = sDate;
  return tt;
}

double TimeDifference(double jul_day1, int year1, double jul_day2, int year2, int calendar) {
  double diff = jul_day2 - jul_day1;
  int yr = year2 - 1;
  while (yr >= year1) { diff += 365 + (IsLeapYear(yr, calendar) ? 1 : 0); yr--; }
  yr = year2;
  while (yr < year1) { diff -= 365 + (IsLeapYear(yr, calendar) ? 1 : 0); yr++; }
  return diff;
}

double DayAngle(double julian_day, int year, int calendar) {
  double leap = IsLeapYear(year, calendar) ? 1.0 : 0.0;
  return 2.0 * RVH_PI * (julian_day / (365.0 + leap));
}

// ---------------------------------------------------------------------------------- special functions
double gamma2(double x) {
  // Taylor coefficients of 1/Gamma(z) (Zhang & Jin, "Computation of Special Functions", routine GAMMA),
  // the same published series the reference uses.
  static const double g[25] = {1.0, 0.5772156649015329, -0.6558780715202538, -0.420026350340952e-1, 0.1665386113822915,
                               -0.421977345555443e-1, -0.9621971527877e-2, 0.7218943246663e-2, -0.11651675918591e-2,
                               -0.2152416741149e-3, 0.1280502823882e-3, -0.201348547807e-4, -0.12504934821e-5,
                               0.1133027232e-5, -0.2056338417e-6, 0.6116095e-8, 0.50020075e-8, -0.11812746e-8,
                               0.1043427e-9, 0.77823e-11, -0.36968e-11, 0.51e-12, -0.206e-13, -0.54e-14, 0.14e-14};
  double ga = 0, gr, r = 0, z;
  if (x > 171.0) return 0.0;
  if (x == (int)x) {
    if (x > 0.0) {
      ga = 1.0;
      for (int i = 2; i < x; i++) ga *= i;
    } else {
      ExitGracefully("Gamma:negative integer values not allowed");
    }
  } else {
    z = x;
    r = 1.0;
    int m = 0;
    if (fabs(x) > 1.0) {
      z = fabs(x);
      m = (int)(z);
      r = 1.0;
      for (int k = 1; k <= m; k++) r *= (z - k);
      z -= m;
    }
    gr = g[24];
    for (int k = 23; k >= 0; k--) gr = gr * z + g[k];
    ga = 1.0 / (gr * z);
    if (fabs(x) > 1.0) {
      ga *= r;
      if (x < 0.0) ga = -RVH_PI / (x * ga * sin(RVH_PI * x));
    }
  }
  return ga;
}

double IncompleteGamma(double x, double a) {
  const int N = 100;
  if (x <= 0) return 0.0;
  if (x > 50) return IncompleteGamma(49.9, a);
  double num = 1.0, sum = 0.0, prod = 1.0;
  for (int n = 0; n < N; n++) {
    if (n > 0) num *= x;
    prod *= (a + n);
    sum += num / prod;
  }
  return sum * pow(x, a) * exp(-x);
}
double GammaCumDist(double t, double alpha, double beta) { return IncompleteGamma(beta * t, alpha) / gamma2(alpha); }

double TriCumDist(double t, double tc, double tp) {
  double b = 2.0 / tc, m;
  if (t < 0.0) return 0.0;
  if (t <= tp) {
    m = (b / tp);
    return 0.5 * m * t * t;
  } else if (t <= tc) {
    m = -b / (tc - tp);
    return tp / tc + b * (t - tp) + 0.5 * m * (t - tp) * (t - tp);
  }
  return 1.0;
}

double InterpolateMo(const double aVal[12], const time_struct &tt, month_interp method, const optStruct &Options) {
  static const double DAYS_PER_MONTH[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
  double wt;
  int leap = 0, mo, nextmo;
  int day = tt.day_of_month, month = tt.month, year = tt.year;
  if (method == MONTHINT_UNIFORM) return aVal[month - 1];
  if (method == MONTHINT_LINEAR_FOM) {
    mo = month - 1;
    nextmo = mo + 1;
    if (nextmo == 12) nextmo = 0;
    leap = (IsLeapYear(year, Options.calendar) && (mo == 1)) ? 1 : 0;
    wt = 1.0 - (double)(day) / (DAYS_PER_MONTH[mo] + leap);
    return wt * aVal[mo] + (1 - wt) * aVal[nextmo];
  }
  double pivot = 0.0;
  if (method == MONTHINT_LINEAR_21) pivot = 21;
  else {
    pivot = 0.5 * DAYS_PER_MONTH[month - 1];
    if (IsLeapYear(year, Options.calendar) && (month == 2)) pivot += 0.5;
  }
  if (day <= pivot) {
    mo = month - 2;
    nextmo = mo + 1;
    if (mo == -1) { mo = 11; nextmo = 0; }
    leap = (IsLeapYear(year, Options.calendar) && (mo == 1)) ? 1 : 0;
    wt = 1.0 - (double)(day + (DAYS_PER_MONTH[mo] + leap) - pivot) / (double)(DAYS_PER_MONTH[mo] + leap);
  } else {
    mo = month - 1;
    nextmo = mo + 1;
    int lastmo = mo - 1;
    if (lastmo == -1) lastmo = 11;
    if (nextmo == 12) nextmo = 0;
    leap = (IsLeapYear(year, Options.calendar) && (lastmo == 1)) ? 1 : 0;
    wt = 1.0 - (double)(day - pivot) / (double)(DAYS_PER_MONTH[lastmo] + leap);
  }
  return wt * aVal[mo] + (1 - wt) * aVal[nextmo];
}

// ---- extraterrestrial radiation on an inclined surface ---------------------------------------------------------
// Restates CRadiation::SolarDeclination / EccentricityCorr (src/Radiation.cpp:484-515) and CRadiation::CalcETRadiation2
// (src/Radiation.cpp:637-822; analytical integration of the beam over the sunlit interval after Allen et al. 2006,
// including the double-sunrise case of steep pole-facing slopes).  The result depends on (HRU geometry, date, step
// interval) only -- never on an ensemble member -- so the B200 build evaluates it once per distinct date on the host
// (same libm as the reference) instead of once per HRU per member per step.
double SolarDeclination(double day_angle) { return 0.40928 * sin(day_angle - 1.384); }
double EccentricityCorr(double day_angle) {
  return 1.000110 + 0.034221 * cos(1.0 * day_angle) + 0.001280 * sin(1.0 * day_angle) + 0.000719 * cos(2.0 * day_angle) + 0.000077 * sin(2.0 * day_angle);
}
namespace {
const double PI = RVH_PI;
const double SOLAR_CONSTANT = 118.1;   // [MJ/m2/d] src/RavenInclude.h:223
struct BeamGeom {   // Allen et al. (2006) eqn 11 constants + the trigonometric pieces the integral reuses
  double a, b, c, sd, cd, sl, cl, ss, cs, sr, cr;
  double cos_at(double w) const { return -a + b * cos(w) + c * sin(w); }   // eqn 14
  // eqn 5 integrated from w_lo to w_hi, terms in the reference's factored order
  double integral(double w_lo, double w_hi) const {
    return (sd * sl * cs - sd * cl * ss * cr) * (w_hi - w_lo) + (cd * cl * cs + cd * sl * ss * cr) * (sin(w_hi) - sin(w_lo)) - (cd * ss * sr) * (cos(w_hi) - cos(w_lo));
  }
};
}  // namespace
// CRadiation::MassFunct / OpticalAirMass (src/Radiation.cpp:875-967; X. Yin 1997): optical air mass averaged over the day or at
// the step's solar time.  Geometry and date only -> evaluated on the host with the other per-date tables.
static double MassFunct(double latrad, double declin, double cos_omt, double c1, double c2, double c3, double t) {
  const double EARTH_ANG_VEL = 6.283185;
  double a, b, tmp;
  a = c1 + sin(latrad) * sin(declin);
  b = cos(latrad) * cos(declin);
  if (a > b) {
    tmp = (b + a * cos_omt) / (a + b * cos_omt);
    return 1.0 / EARTH_ANG_VEL * c2 / sqrt(a * a - b * b) * acos(tmp) + c3 * t;
  } else if (a < b) {
    tmp = sqrt(b + a) * sqrt(1.0 + cos_omt) + sqrt(b - a) * sqrt(1.0 - cos_omt);
    tmp /= (sqrt(b + a) * sqrt(1.0 + cos_omt) - sqrt(b - a) * sqrt(1.0 - cos_omt));
    return 1.0 / EARTH_ANG_VEL * c2 / sqrt(b * b - a * a) * log(tmp) + c3 * t;
  } else {
    return 1.0 / EARTH_ANG_VEL * c2 / a * tan(0.5 * acos(cos_omt)) + c3 * t;
  }
}
double OpticalAirMass(double latrad, double declin, double day_length, double t_sol, bool avg_daily) {
  const double EARTH_ANG_VEL = 6.283185;
  const double c1a = 0.00830700, c2a = 1.02129227, c3a = -0.01287829, c1b = 0.03716000, c2b = 1.53832220, c3b = -1.69726057, m90 = 39.7;
  static const double deg80 = 80.0 / 180.0 * PI;
  static const double cos80 = cos(deg80);
  double t1, Zn, Z0, t80;
  t1 = 0.5 * day_length;
  if (t1 == 0.0) return m90;
  Zn = acos(sin(latrad) * sin(declin) - cos(latrad) * cos(declin));
  if (fabs(latrad + declin) <= 0.5 * PI) Zn = 0.5 * PI;
  if (avg_daily) {
    Z0 = 0.5 * PI;
    if (fabs(latrad - declin) <= 0.5 * PI) Z0 = latrad - declin;
    t80 = (1.0 / EARTH_ANG_VEL) * acos((cos80 - sin(latrad) * sin(declin)) / (cos(latrad) * cos(declin)));
    double cos_omt = cos(EARTH_ANG_VEL * t1);
    if (Zn <= deg80) return MassFunct(latrad, declin, cos_omt, c1a, c2a, c3a, t1) / t1;
    else if (Z0 >= deg80) return MassFunct(latrad, declin, cos_omt, c1b, c2b, c3b, t1) / t1;
    else {
      double cos_omt80 = cos(EARTH_ANG_VEL * t80);
      double sum = 0;
      sum += MassFunct(latrad, declin, cos_omt80, c1a, c2a, c3a, t80) / t1;
      sum += MassFunct(latrad, declin, cos_omt, c1b, c2b, c3b, t1) / t1;
      sum -= MassFunct(latrad, declin, cos_omt80, c1b, c2b, c3b, t80) / t1;
      return sum;
    }
  } else {
    double cosZ = sin(latrad) * sin(declin) + cos(latrad) * cos(declin) * cos(EARTH_ANG_VEL * t_sol);
    if (cosZ < 0) return m90;
    if (cosZ > cos80) return c2a / (c1a + cosZ) + c3a;
    else return c2b / (c1b + cosZ) + c3b;
  }
}
double CalcETRadiation2(double latrad, double aspect, double declin, double ecc, double slope, double t1, double t2, bool avg_daily) {
  double revasp = -(aspect + PI);
  if (revasp > 2 * PI) revasp -= 2 * PI;
  BeamGeom g;
  g.cs = cos(slope); g.ss = sin(slope); g.cd = cos(declin); g.sd = sin(declin);
  g.cl = cos(latrad); g.sl = sin(latrad); g.cr = cos(revasp); g.sr = sin(revasp);
  g.a = g.sd * (g.cl * g.ss * g.cr - g.sl * g.cs);
  g.b = g.cd * (g.cl * g.cs + g.sl * g.ss * g.cr);
  g.c = g.cd * g.ss * g.sr;
  double quad = g.b * g.b + g.c * g.c - g.a * g.a;
  quad = rvh_max(0.0001, quad);
  const double noon_cos_theta = (g.sd * g.sl + g.cd * g.cl) * g.cs + (g.cd * g.sl - g.sd * g.cl) * g.ss * g.cr;
  // horizontal-surface sunset / sunrise hour angles (eqn 8) with polar day / night
  double ws2 = acos(-g.sd / g.cd * g.sl / g.cl), ws1 = -ws2;
  if (latrad > 0) {
    if (declin + latrad > PI / 2.0) { ws1 = -PI; ws2 = PI; }
    else if (fabs(declin - latrad) > PI / 2.0) { ws1 = 0.0; ws2 = 0.0; }
  } else if (latrad < 0) {
    if (declin + latrad < -PI / 2.0) { ws1 = -PI; ws2 = PI; }
    else if (fabs(declin - latrad) < -PI / 2.0) { ws1 = 0; ws2 = 0; }
  }
  const double root = sqrt(quad), den = g.b * g.b + g.c * g.c;
  // sunrise on the slope (eqn 13a, steps A.2)
  double w1_24 = 0.0;
  {
    double s = (g.a * g.c - g.b * root) / den;
    s = rvh_max(-1.0, s); s = rvh_min(1.0, s);
    const double w1 = asin(s), w1x = -PI - w1;
    const double cos_w1 = g.cos_at(w1), cos_w1x = g.cos_at(w1x), cos_ws1 = g.cos_at(ws1);
    if ((cos_ws1 <= cos_w1) && (cos_w1 < 0.001)) w1_24 = w1;
    else if (cos_w1x > 0.001) w1_24 = ws1;
    else if (w1x <= ws1) w1_24 = ws1;
    else if (w1x > ws1) w1_24 = w1x;
    w1_24 = rvh_max(w1_24, ws1);
  }
  // sunset on the slope (eqn 13b, steps A.3)
  double w2_24 = 0.0;
  {
    double s = (g.a * g.c + g.b * root) / den;
    s = rvh_max(-1.0, s); s = rvh_min(1.0, s);
    const double w2 = asin(s), w2x = PI - w2;
    const double cos_w2 = g.cos_at(w2), cos_w2x = g.cos_at(w2x), cos_ws2 = g.cos_at(ws2);
    if ((cos_ws2 <= cos_w2) && (cos_w2 < 0.001)) w2_24 = w2;
    else if (cos_w2x > 0.001) w2_24 = ws2;
    else if (w2x >= ws2) w2_24 = ws2;
    else if (w2x < ws2) w2_24 = w2x;
    w2_24 = rvh_min(w2_24, ws2);
  }
  if (w2_24 < w1_24) w1_24 = w2_24;   // slope always shaded
  double cos_theta = 0.0;
  bool two_periods = false;
  if (noon_cos_theta < 0) {   // the slope may be shaded around noon: two sunlit periods (eqns 44-51)
    double sA = (g.a * g.c + g.b * root) / den; sA = rvh_max(-1.0, sA); sA = rvh_min(1.0, sA);
    double sB = (g.a * g.c - g.b * root) / den; sB = rvh_max(-1.0, sB); sB = rvh_min(1.0, sB);
    const double AA = asin(sA), BB = asin(sB);
    double w2_24b = rvh_min(AA, BB), w1_24b = rvh_max(AA, BB);
    const double cos_w2b = g.cos_at(w2_24b), cos_w1b = g.cos_at(w1_24b);
    if ((cos_w2b < -0.001) || (cos_w2b > 0.001)) w2_24b = -PI - w2_24b;
    if ((cos_w1b < -0.001) || (cos_w1b > 0.001)) w1_24b = PI - w1_24b;
    if ((w2_24b >= w1_24) && (w1_24b <= w2_24)) {
      // the reference evaluates this test integral term by term with un-factored products
      const double X = sin(declin) * sin(latrad) * cos(slope) * (w1_24b - w2_24b) - sin(declin) * cos(latrad) * sin(slope) * cos(revasp) * (w1_24b - w2_24b) +
                       cos(declin) * cos(latrad) * cos(slope) * (sin(w1_24b) - sin(w2_24b)) + cos(declin) * sin(latrad) * sin(slope) * cos(revasp) * (sin(w1_24b) - sin(w2_24b)) -
                       cos(declin) * sin(slope) * sin(revasp) * (cos(w1_24b) - cos(w2_24b));
      if (X < 0) {
        if (!avg_daily) {
          w1_24 = rvh_max(w1_24, t1 * 2.0 * PI); w2_24b = rvh_min(w2_24b, t2 * 2.0 * PI);
          if (w2_24b < w1_24) w2_24b = w1_24;
        }
        const double p1 = sin(declin) * sin(latrad) * cos(slope) * (w2_24b - w1_24) - sin(declin) * cos(latrad) * sin(slope) * cos(revasp) * (w2_24b - w1_24) +
                          cos(declin) * cos(latrad) * cos(slope) * (sin(w2_24b) - sin(w1_24)) + cos(declin) * sin(latrad) * sin(slope) * cos(revasp) * (sin(w2_24b) - sin(w1_24)) -
                          cos(declin) * sin(slope) * sin(revasp) * (cos(w2_24b) - cos(w1_24));
        if (!avg_daily) {
          w1_24b = rvh_max(w1_24b, t1 * 2.0 * PI); w2_24 = rvh_min(w2_24, t2 * 2.0 * PI);
          if (w2_24 < w1_24b) w2_24 = w1_24b;
        }
        const double p2 = sin(declin) * sin(latrad) * cos(slope) * (w2_24 - w1_24b) - sin(declin) * cos(latrad) * sin(slope) * cos(revasp) * (w2_24 - w1_24b) +
                          cos(declin) * cos(latrad) * cos(slope) * (sin(w2_24) - sin(w1_24b)) + cos(declin) * sin(latrad) * sin(slope) * cos(revasp) * (sin(w2_24) - sin(w1_24b)) -
                          cos(declin) * sin(slope) * sin(revasp) * (cos(w2_24) - cos(w1_24b));
        two_periods = true;
        cos_theta = p1 + p2;
      }
    }
  }
  if (!two_periods) {
    if (!avg_daily) {
      w1_24 = rvh_max(w1_24, t1 * 2.0 * PI); w2_24 = rvh_min(w2_24, t2 * 2.0 * PI);
      if (w2_24 <= w1_24) w2_24 = w1_24;
    }
    cos_theta = g.integral(w1_24, w2_24);
  }
  return rvh_max(0.0, SOLAR_CONSTANT * cos_theta * ecc) / 2.0 / PI / (t2 - t1);
}

// ---- diffusive-wave routing hydrograph (reference src/CommonFunctions.cpp:1487-1512, 1752-1806) --------------------------------
// erfc: rational approximations the reference uses (Abramowitz & Stegun 7.1.26 below 3, asymptotic series above); the routing
// hydrograph inherits their few-1e-7 accuracy, so the same formulas are used here.
static double rvh_erfc(double x) {
  const double a = fabs(x);
  double fun;
  if (a > 3.0) {
    const double f1 = (1.0 - 1.0 / (2.0 * a * a) + 3.0 / (4.0 * pow(a, 4)) - 5.0 / (6.0 * pow(a, 6)));
    fun = f1 * exp(-a * a) / (a * sqrt(RVH_PI));
  } else {
    const double t = 1.0 / (1.0 + (0.3275911 * a));
    const double poly = 0.254829592 * t - (0.284496736 * t * t) + (1.421413741 * pow(t, 3)) - (1.453152027 * pow(t, 4)) + (1.061405429 * pow(t, 5));
    fun = poly * exp(-a * a);
  }
  return (x < 0) ? 2.0 - fun : fun;
}
// Ogata & Banks (1961) cumulative arrival distribution of the advection-diffusion equation at distance L
static double ADRCumDist(double t, double L, double v, double D) {
  ExitGracefullyIf(D <= 0, "ADRCumDist: Invalid diffusivity");
  double F = 0;
  if (t <= 0) return 0.0;
  if (L < 500 * (D / v)) F = 0.5 * (exp((v * L) / D) * rvh_erfc((L + v * t) / sqrt(4.0 * D * t)));
  F += 0.5 * (rvh_erfc((L - v * t) / sqrt(4.0 * D * t)));
  return F;
}
static double G_integral(double t, double L, double v, double D) {
  double G = 0;
  if (t <= 0) return 0.0;
  if (L < 500 * (D / v)) G = -0.5 * (exp((v * L) / D) * rvh_erfc((L + v * t) / sqrt(4.0 * D * t)));
  G += 0.5 * (rvh_erfc((L - v * t) / sqrt(4.0 * D * t)));
  return L / v * G;
}
double DiffusiveWaveUH(int n, double L, double v, double D, double dt) {
  if (n == 0) return 1.0 / dt * (G_integral(0.0, L, v, D) - G_integral(dt, L, v, D)) + ADRCumDist(dt, L, v, D) - ADRCumDist(0.0, L, v, D);
  double UH = 1.0 / dt * (2 * G_integral(n * dt, L, v, D) - G_integral((n - 1) * dt, L, v, D) - G_integral((n + 1) * dt, L, v, D));
  UH += -(2 * n) * ADRCumDist(n * dt, L, v, D) + (n - 1) * ADRCumDist((n - 1) * dt, L, v, D) + (n + 1) * ADRCumDist((n + 1) * dt, L, v, D);
  return UH;
}

// ---- geographic -> UTM coordinates of HRU centroids and gauges, as the reference projects them for distance-based gauge interpolation
// (LatLonToUTMXY / MapLatLonToXY / ArcLengthOfMeridian, src/UTM_to_LatLong.cpp:20-37,88-129,216-227: the transverse-Mercator series of
// Hoffmann-Wellenhof et al. on the WGS84 ellipsoid).  The operation order is the reference's so that inverse-distance weights are identical.
void LatLonToUTMXY(double lat_deg, double lon_deg, int zone, double &x, double &y) {
  const double a = 6378137.0, b = 6356752.314, k0 = 0.9996, DEGREES_TO_RADIANS = 0.017453293;   // the rounded literal of src/RavenInclude.h:166
  const double phi = lat_deg * DEGREES_TO_RADIANS, lambda = lon_deg * DEGREES_TO_RADIANS, lambda0 = (-183.0 + (zone * 6.0)) * DEGREES_TO_RADIANS;
  // meridian arc from the equator to phi
  const double n = (a - b) / (a + b);
  const double alpha = ((a + b) / 2.0) * (1.0 + (pow(n, 2.0) / 4.0) + (pow(n, 4.0) / 64.0));
  const double beta = (-3.0 * n / 2.0) + (9.0 * pow(n, 3.0) / 16.0) + (-3.0 * pow(n, 5.0) / 32.0);
  const double gamma = (15.0 * pow(n, 2.0) / 16.0) + (-15.0 * pow(n, 4.0) / 32.0);
  const double delta = (-35.0 * pow(n, 3.0) / 48.0) + (105.0 * pow(n, 5.0) / 256.0);
  const double epsilon = (315.0 * pow(n, 4.0) / 512.0);
  const double arc = alpha * (phi + (beta * sin(2.0 * phi)) + (gamma * sin(4.0 * phi)) + (delta * sin(6.0 * phi)) + (epsilon * sin(8.0 * phi)));
  // transverse Mercator
  const double ep2 = (pow(a, 2.0) - pow(b, 2.0)) / pow(b, 2.0);
  const double nu2 = ep2 * pow(cos(phi), 2.0);
  const double N = pow(a, 2.0) / (b * sqrt(1 + nu2));
  const double t = tan(phi), t2 = t * t, l = lambda - lambda0;
  const double c3 = 1.0 - t2 + nu2;
  const double c4 = 5.0 - t2 + 9 * nu2 + 4.0 * (nu2 * nu2);
  const double c5 = 5.0 - 18.0 * t2 + (t2 * t2) + 14.0 * nu2 - 58.0 * t2 * nu2;
  const double c6 = 61.0 - 58.0 * t2 + (t2 * t2) + 270.0 * nu2 - 330.0 * t2 * nu2;
  const double c7 = 61.0 - 479.0 * t2 + 179.0 * (t2 * t2) - (t2 * t2 * t2);
  const double c8 = 1385.0 - 3111.0 * t2 + 543.0 * (t2 * t2) - (t2 * t2 * t2);
  x = N * cos(phi) * l + (N / 6.0 * pow(cos(phi), 3.0) * c3 * pow(l, 3.0)) + (N / 120.0 * pow(cos(phi), 5.0) * c5 * pow(l, 5.0)) +
      (N / 5040.0 * pow(cos(phi), 7.0) * c7 * pow(l, 7.0));
  y = arc + (t / 2.0 * N * pow(cos(phi), 2.0) * pow(l, 2.0)) + (t / 24.0 * N * pow(cos(phi), 4.0) * c4 * pow(l, 4.0)) +
      (t / 720.0 * N * pow(cos(phi), 6.0) * c6 * pow(l, 6.0)) + (t / 40320.0 * N * pow(cos(phi), 8.0) * c8 * pow(l, 8.0));
  x = x * k0 + 500000.0;
  y = y * k0;
  if (y < 0.0) y = y + 10000000.0;
}
