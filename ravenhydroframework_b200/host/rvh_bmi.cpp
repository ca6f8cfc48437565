// rvh_bmi.cpp -- see rvh_bmi.h.  Initialize/Update follow the call sequence of the reference (src/Raven_BMI.cpp:300-401);
// everything that happens per step runs on the GPU through CGPUSimulation (-> rvn_gpu_run).
#include "rvh_bmi.h"
#include <fstream>
#include <sstream>

static const char *kForcingNames[RVN_F_COUNT] = {"PRECIP", "SNOW_FRAC", "TEMP_AVE", "TEMP_DAILY_AVE", "TEMP_DAILY_MIN", "TEMP_DAILY_MAX",
                                                 "PET", "OW_PET", "POTENTIAL_MELT", "TEMP_MONTH_AVE", "PET_MONTH_AVE",
                                                 "AIR_PRES", "REL_HUMIDITY", "SW_RADIA", "WIND_VEL", "SW_RADIA_NET", "LW_RADIA_NET"};
static_assert(sizeof(kForcingNames) / sizeof(kForcingNames[0]) == RVN_F_COUNT, "one name per rvn_forcing_field");

CRavenBMI::CRavenBMI() : pModel(NULL), _sim(NULL), _duration(0.0) {}
CRavenBMI::~CRavenBMI() { delete _sim; delete pModel; }

static std::string trim(const std::string &s) {
  size_t a = s.find_first_not_of(" \t\r"), b = s.find_last_not_of(" \t\r");
  return (a == std::string::npos) ? std::string() : s.substr(a, b - a + 1);
}
// config file (src/Raven_BMI.cpp:113-270):  cli_args: <base> [-o dir]   duration: <days>   input_vars: / output_vars: lists of
// "- NAME: hru_state | subbasin_state | forcing"
void CRavenBMI::_ReadConfigFile(const std::string &config_file) {
  std::ifstream CONFIG(config_file.c_str());
  if (CONFIG.fail()) throw std::logic_error("Cannot find configuration file " + config_file);
  bool cli_args = false, in_inp = false, in_out = false;
  for (std::string line; std::getline(CONFIG, line);) {
    if (line.empty() || line[0] == '#') continue;
    size_t c = line.find(":");
    if (c == std::string::npos) throw std::logic_error("WARNING: Invalid line in Raven config file: '" + line + "'");
    std::string key = trim(line.substr(0, c)), val = trim(line.substr(c + 1));
    if (key == "cli_args") {
      cli_args = true; in_inp = in_out = false;
      std::istringstream iss(val);
      std::vector<std::string> tok;
      for (std::string t; iss >> t;) tok.push_back(t);
      if (tok.empty()) throw std::logic_error("RavenBMI:ReadConfigFile: cli_args needs the input file base name");
      _base = tok[0];
      for (size_t i = 1; i + 1 < tok.size(); i++) if (tok[i] == "-o") Options.output_dir = tok[i + 1];
      continue;
    }
    if (key == "duration") { _duration = s_to_d(val); in_inp = in_out = false; continue; }
    if (key == "input_vars") { in_inp = true; in_out = false; continue; }
    if (key == "output_vars") { in_out = true; in_inp = false; continue; }
    if (in_inp || in_out) {
      if (key.substr(0, 2) != "- ") throw std::logic_error("WARNING: Invalid input var line in Raven config file: " + line);
      key = trim(key.substr(2));
      bmi_var_kind kind;
      if (val == "forcing") kind = VAR_FORCING_FUNCTION;
      else if (val == "hru_state") kind = VAR_STATE_VAR;
      else if (val == "subbasin_state") {
        if (key == "STREAMFLOW") kind = VAR_STREAMFLOW;
        else throw std::logic_error("WARNING: Invalid subbasin state variable type in Raven config file input_vars: '" + line + "'");
      } else if (val == "flux") throw std::logic_error("WARNING: flux variables as input are not supported in the BMI interface yet.");
      else throw std::logic_error("WARNING: Invalid input var line in Raven config file: " + line);
      (in_inp ? _input_vars : _output_vars).push_back(rvn_var_data(kind, key));
      continue;
    }
    throw std::logic_error("WARNING: Invalid line in Raven config file: '" + line + "'");
  }
  if (!cli_args) throw std::logic_error("RavenBMI:ReadConfigFile: missing cli_args command in configuration file");
  if (_duration == 0.0) throw std::logic_error("RavenBMI:ReadConfigFile: missing duration command in configuration file");
}
void CRavenBMI::_CheckConfigVars(std::vector<rvn_var_data> &vars) {
  for (size_t i = 0; i < vars.size(); i++) {
    if (vars[i].type == VAR_STATE_VAR) {
      int layer; sv_type t = pModel->StringToSVType(vars[i].name, layer, false);
      vars[i].sv_index = (t == RVN_SV_NTYPES) ? DOESNT_EXIST : pModel->GetStateVarIndex(t, layer);
      if (vars[i].sv_index == DOESNT_EXIST) throw std::logic_error("WARNING: config variable '" + vars[i].name + "' has an invalid state variable type.");
    } else if (vars[i].type == VAR_FORCING_FUNCTION) {
      for (int f = 0; f < RVN_F_COUNT; f++) if (vars[i].name == kForcingNames[f]) vars[i].f_index = f;
      if (vars[i].f_index == DOESNT_EXIST) throw std::logic_error("WARNING: config variable '" + vars[i].name + "' has an invalid state variable type.");
    }
  }
}
void CRavenBMI::Initialize(std::string config_file) {
  _ReadConfigFile(config_file);
  std::string dir = DirName(config_file);
  std::string b = (!_base.empty() && _base[0] == '/') ? _base : dir + _base;
  Options.rvi_filename = b + ".rvi"; Options.rvh_filename = b + ".rvh"; Options.rvp_filename = b + ".rvp";
  Options.rvt_filename = b + ".rvt"; Options.rvc_filename = b + ".rvc"; Options.rve_filename = b + ".rve";
  try {
    ParseInputFiles(pModel, Options);
    if (_duration > Options.duration) throw std::logic_error("WARNING: Duration indicated in .rvi file must be greater than config file indicates");
    Options.duration = _duration;
    pModel->Initialize(Options);
    ParseInitialConditions(pModel, Options);
  } catch (const RavenError &e) { throw std::logic_error(e.what()); }
  _CheckConfigVars(_input_vars);
  _CheckConfigVars(_output_vars);
  try {
    _sim = new CGPUSimulation(pModel, Options, 1, 0);
    bool wants_forcing = false;
    for (size_t i = 0; i < _output_vars.size(); i++) if (_output_vars[i].type == VAR_FORCING_FUNCTION) wants_forcing = true;
    if (wants_forcing && rvn_gpu_enable_forcing_output(_sim->engine(), 1) != RVN_OK) throw RavenError(rvn_gpu_last_error());
  } catch (const RavenError &e) { throw std::logic_error(e.what()); }
  JulianConvert(0.0, Options.julian_start_day, Options.julian_start_year, Options.calendar, tt);
}
void CRavenBMI::Update() {
  if (!_sim) throw std::logic_error("CRavenBMI.Update: model is not initialized");
  try {
    _sim->Step();   // UpdateHRUForcingFunctions + MassEnergyBalance + IncrementCumulInput/CumOutflow for this step, on the GPU
  } catch (const RavenError &e) { throw std::logic_error(e.what()); }
  JulianConvert(tt.model_time + Options.timestep, Options.julian_start_day, Options.julian_start_year, Options.calendar, tt);
}
void CRavenBMI::UpdateUntil(double time) {
  for (double t = tt.model_time; t < (time + 0.1 * Options.timestep); t += Options.timestep) Update();
}
void CRavenBMI::Finalize() { if (_sim) _sim->Sync(); }
std::string CRavenBMI::GetComponentName() { return "Raven Hydrological Modelling Framework (B200 build)"; }
std::vector<std::string> CRavenBMI::GetInputVarNames() { std::vector<std::string> n; for (size_t i = 0; i < _input_vars.size(); i++) n.push_back(_input_vars[i].name); return n; }
std::vector<std::string> CRavenBMI::GetOutputVarNames() { std::vector<std::string> n; for (size_t i = 0; i < _output_vars.size(); i++) n.push_back(_output_vars[i].name); return n; }
const rvn_var_data &CRavenBMI::_Find(const std::string &name, bool input_first) const {
  const std::vector<rvn_var_data> *a = input_first ? &_input_vars : &_output_vars, *b = input_first ? &_output_vars : &_input_vars;
  for (size_t i = 0; i < a->size(); i++) if ((*a)[i].name == name) return (*a)[i];
  for (size_t i = 0; i < b->size(); i++) if ((*b)[i].name == name) return (*b)[i];
  throw std::logic_error("RavenBMI: variable '" + name + "' is invalid, not in config.txt, or not yet supported.");
}
int CRavenBMI::GetVarGrid(std::string name) { return (_Find(name, false).type == VAR_STREAMFLOW) ? GRID_SUBBASIN : GRID_HRU; }
std::string CRavenBMI::GetVarUnits(std::string name) {
  const rvn_var_data &v = _Find(name, false);
  if (v.type == VAR_STREAMFLOW) return "m3 s-1";
  if (v.type == VAR_FORCING_FUNCTION) return (v.name.substr(0, 4) == "TEMP") ? "C" : ((v.name == "SNOW_FRAC") ? "0-1" : "mm/d");
  return "mm";
}
int CRavenBMI::GetGridSize(const int grid) {
  if (grid == GRID_HRU) return pModel->GetNumHRUs();
  if (grid == GRID_SUBBASIN) return pModel->GetNumSubBasins();
  throw std::logic_error("RavenBMI.GetGridSize: invalid grid");
}
void CRavenBMI::GetGridX(const int grid, double *x) {
  if (grid != GRID_HRU) throw std::logic_error("RavenBMI.GetGridX: only the HRU grid has coordinates");
  for (int k = 0; k < pModel->GetNumHRUs(); k++) x[k] = pModel->GetHydroUnit(k)->longitude;
}
void CRavenBMI::GetGridY(const int grid, double *y) {
  if (grid != GRID_HRU) throw std::logic_error("RavenBMI.GetGridY: only the HRU grid has coordinates");
  for (int k = 0; k < pModel->GetNumHRUs(); k++) y[k] = pModel->GetHydroUnit(k)->latitude;
}
void CRavenBMI::GetGridZ(const int grid, double *z) {
  if (grid != GRID_HRU) throw std::logic_error("RavenBMI.GetGridZ: only the HRU grid has coordinates");
  for (int k = 0; k < pModel->GetNumHRUs(); k++) z[k] = pModel->GetHydroUnit(k)->elevation;
}
void CRavenBMI::_Fetch(const rvn_var_data &v, std::vector<double> &out) {
  const int nH = pModel->GetNumHRUs(), NS = pModel->GetNumStateVars(), NB = pModel->GetNumSubBasins();
  rvn_gpu_model *g = _sim->engine();
  if (v.type == VAR_STREAMFLOW) {   // CSubBasin::GetOutflowRate (src/Raven_BMI.cpp:592-603)
    std::vector<double> q((size_t)NB * RVN_MAX_RIVER_SEGS);
    if (rvn_gpu_download_basin_state(g, NULL, NULL, q.data(), NULL, 0, 1) != RVN_OK) throw std::logic_error(rvn_gpu_last_error());
    out.resize(NB);
    for (int p = 0; p < NB; p++) out[p] = q[(size_t)p * RVN_MAX_RIVER_SEGS + pModel->GetSubBasin(p)->GetNumSegments() - 1];
  } else if (v.type == VAR_STATE_VAR) {
    std::vector<double> st((size_t)nH * NS);
    if (rvn_gpu_download_state(g, st.data(), 0, 1) != RVN_OK) throw std::logic_error(rvn_gpu_last_error());
    out.resize(nH);
    for (int k = 0; k < nH; k++) out[k] = st[(size_t)k * NS + v.sv_index];
  } else {
    std::vector<double> f((size_t)nH * RVN_F_COUNT);
    if (rvn_gpu_download_forcings(g, f.data(), 0, 1) != RVN_OK) throw std::logic_error(rvn_gpu_last_error());
    out.resize(nH);
    for (int k = 0; k < nH; k++) out[k] = f[(size_t)k * RVN_F_COUNT + v.f_index];
  }
}
void CRavenBMI::GetValue(std::string name, void *dest) {
  const rvn_var_data *v = NULL;
  for (size_t i = 0; i < _output_vars.size(); i++) if (_output_vars[i].name == name) v = &_output_vars[i];
  if (!v) throw std::logic_error("CRavenBMI.GetValue: unsupported variable name or not included in config file.");
  std::vector<double> out;
  _Fetch(*v, out);
  memcpy(dest, out.data(), out.size() * sizeof(double));
}
void CRavenBMI::GetValueAtIndices(std::string name, void *dest, int *inds, int count) {
  const rvn_var_data *v = NULL;
  for (size_t i = 0; i < _output_vars.size(); i++) if (_output_vars[i].name == name) v = &_output_vars[i];
  if (!v) throw std::logic_error("CRavenBMI.GetValueAtIndices: unsupported variable name or not included in config file.");
  std::vector<double> out;
  _Fetch(*v, out);
  double *d = (double *)dest;
  for (int i = 0; i < count; i++) {
    if (inds[i] < 0 || inds[i] >= (int)out.size()) throw std::logic_error("CRavenBMI.GetValueAtIndices: index out of range");
    d[i] = out[inds[i]];
  }
}
void CRavenBMI::SetValue(std::string name, void *src) {   // src/Raven_BMI.cpp:700-760: state variables of all HRUs
  const rvn_var_data *v = NULL;
  for (size_t i = 0; i < _input_vars.size(); i++) if (_input_vars[i].name == name) v = &_input_vars[i];
  if (!v) throw std::logic_error("CRavenBMI.SetValue: unsupported variable name or not included in config file.");
  if (v->type != VAR_STATE_VAR) throw std::logic_error("CRavenBMI.SetValue: only hru_state variables can be set in the B200 build");
  const int nH = pModel->GetNumHRUs(), NS = pModel->GetNumStateVars();
  std::vector<double> st((size_t)nH * NS);
  rvn_gpu_model *g = _sim->engine();
  if (rvn_gpu_download_state(g, st.data(), 0, 1) != RVN_OK) throw std::logic_error(rvn_gpu_last_error());
  const double *in = (const double *)src;
  for (int k = 0; k < nH; k++) st[(size_t)k * NS + v->sv_index] = in[k];
  if (rvn_gpu_upload_state(g, st.data(), 0, 1) != RVN_OK) throw std::logic_error(rvn_gpu_last_error());
}
void CRavenBMI::SetValueAtIndices(std::string name, int *inds, int count, void *src) {
  const rvn_var_data *v = NULL;
  for (size_t i = 0; i < _input_vars.size(); i++) if (_input_vars[i].name == name) v = &_input_vars[i];
  if (!v) throw std::logic_error("CRavenBMI.SetValueAtIndices: unsupported variable name or not included in config file.");
  if (v->type != VAR_STATE_VAR) throw std::logic_error("CRavenBMI.SetValueAtIndices: only hru_state variables can be set in the B200 build");
  const int nH = pModel->GetNumHRUs(), NS = pModel->GetNumStateVars();
  std::vector<double> st((size_t)nH * NS);
  rvn_gpu_model *g = _sim->engine();
  if (rvn_gpu_download_state(g, st.data(), 0, 1) != RVN_OK) throw std::logic_error(rvn_gpu_last_error());
  const double *in = (const double *)src;
  for (int i = 0; i < count; i++) {
    if (inds[i] < 0 || inds[i] >= nH) throw std::logic_error("CRavenBMI.SetValueAtIndices: index out of range");
    st[(size_t)inds[i] * NS + v->sv_index] = in[i];
  }
  if (rvn_gpu_upload_state(g, st.data(), 0, 1) != RVN_OK) throw std::logic_error(rvn_gpu_last_error());
}

extern "C" {
CRavenBMI *bmi_model_create() { return new CRavenBMI(); }
void bmi_model_destroy(CRavenBMI *ptr) { delete ptr; }

// plain-C shims over the C++ object for callers without a C++ ABI (ctypes tests); 0 = ok, -1 = std::logic_error (text in bmi_last_error)
static std::string g_bmi_err;
const char *bmi_last_error() { return g_bmi_err.c_str(); }
#define BMI_TRY try {
#define BMI_CATCH } catch (const std::exception &e) { g_bmi_err = e.what(); return -1; } return 0;
int bmi_initialize(CRavenBMI *p, const char *cfg) { BMI_TRY p->Initialize(cfg); BMI_CATCH }
int bmi_update(CRavenBMI *p) { BMI_TRY p->Update(); BMI_CATCH }
int bmi_update_until(CRavenBMI *p, double t) { BMI_TRY p->UpdateUntil(t); BMI_CATCH }
int bmi_finalize(CRavenBMI *p) { BMI_TRY p->Finalize(); BMI_CATCH }
int bmi_get_value(CRavenBMI *p, const char *name, double *dest) { BMI_TRY p->GetValue(name, dest); BMI_CATCH }
int bmi_get_value_at_indices(CRavenBMI *p, const char *name, double *dest, int *inds, int count) { BMI_TRY p->GetValueAtIndices(name, dest, inds, count); BMI_CATCH }
int bmi_set_value(CRavenBMI *p, const char *name, double *src) { BMI_TRY p->SetValue(name, src); BMI_CATCH }
int bmi_get_var_grid(CRavenBMI *p, const char *name, int *grid) { BMI_TRY *grid = p->GetVarGrid(name); BMI_CATCH }
int bmi_get_grid_size(CRavenBMI *p, int grid, int *n) { BMI_TRY *n = p->GetGridSize(grid); BMI_CATCH }
int bmi_get_var_nbytes(CRavenBMI *p, const char *name, int *n) { BMI_TRY *n = p->GetVarNbytes(name); BMI_CATCH }
double bmi_get_current_time(CRavenBMI *p) { return p->GetCurrentTime(); }
double bmi_get_end_time(CRavenBMI *p) { return p->GetEndTime(); }
double bmi_get_time_step(CRavenBMI *p) { return p->GetTimeStep(); }
int bmi_get_item_counts(CRavenBMI *p, int *nin, int *nout) { *nin = p->GetInputItemCount(); *nout = p->GetOutputItemCount(); return 0; }
}
