// rvh_ensemble.cpp -- see rvh_ensemble.h.  Random numbers come from libc rand() exactly as in the reference, so that a
// member here has the parameters the reference would have given the same member (same seed, same draw order).
#include "rvh_ensemble.h"
#include <cstdlib>

double UniformRandom() { return (double)(rand()) / RAND_MAX; }
double GaussRandom() {   // Box-Muller, cosine branch only: two rand() calls per deviate
  double u1 = (double)(rand() + 1) / (RAND_MAX + 1.0);
  double u2 = (double)(rand()) / RAND_MAX;
  return sqrt(-2.0 * log(u1)) * cos(2.0 * RVH_PI * u2);
}
double SampleFromGamma(const double &shape, const double &scale) {   // Marsaglia & Tsang (2000); variable number of draws
  if (shape < 1.0) {
    double u = UniformRandom();
    return SampleFromGamma(shape + 1.0, scale) * pow(u, 1.0 / shape);
  }
  const double d = shape - 1.0 / 3.0;
  const double c = 1.0 / std::sqrt(9.0 * d);
  double x, v;
  while (true) {
    do {
      x = GaussRandom();
      v = 1.0 + c * x;
    } while (v <= 0.0);
    v = v * v * v;
    double u = UniformRandom();
    if (u < 1.0 - 0.0331 * x * x * x * x) return scale * d * v;
    if (log(u) < 0.5 * x * x + d * (1.0 - v + std::log(v))) return scale * d * v;
  }
}
double SampleFromDistribution(int distribution, const double distpar[3]) {
  double value = 0;
  if (distribution == DIST_UNIFORM) value = distpar[0] + (distpar[1] - distpar[0]) * UniformRandom();
  else if (distribution == DIST_NORMAL) value = distpar[0] + (distpar[1] * GaussRandom());
  else if (distribution == DIST_LOGNORMAL) value = exp(distpar[0] + distpar[1] * GaussRandom());
  else if (distribution == DIST_GAMMA) value = SampleFromGamma(distpar[0], distpar[1]);
  return value;
}

CEnsemble::CEnsemble(int num_members, const optStruct &Options) : _type(ENSEMBLE_NONE), _nMembers(num_members) { (void)Options; }
void CEnsemble::SetRandomSeed(unsigned int seed) { srand(seed); }
void CEnsemble::UpdateModel(CModel *pModel, optStruct &Options, int e) { (void)pModel; (void)Options; (void)e; }

CMonteCarloEnsemble::CMonteCarloEnsemble(int num_members, const optStruct &Options) : CEnsemble(num_members, Options) { _type = ENSEMBLE_MONTECARLO; }
void CMonteCarloEnsemble::Initialize(const CModel *pModel, const optStruct &Options) { (void)pModel; (void)Options; }
void CMonteCarloEnsemble::UpdateModel(CModel *pModel, optStruct &Options, int e) {   // src/ModelEnsemble.cpp:230-263
  CEnsemble::UpdateModel(pModel, Options, e);
  ExitGracefullyIf(e >= _nMembers, "CMonteCarloEnsemble::UpdateMode: invalid ensemble member index");
  _last.resize(_ParamDists.size());
  for (size_t i = 0; i < _ParamDists.size(); i++) {
    const double val = SampleFromDistribution(_ParamDists[i].distribution, _ParamDists[i].distpar);
    pModel->UpdateParameter(_ParamDists[i].param_class, _ParamDists[i].param_name, _ParamDists[i].class_group, val);
    _last[i] = val;
  }
  // the reference re-reads the .rvc here: every member starts from the same initial state (uploaded once per member by the caller)
}

// ----------------------------------------------------------------------------------------------------------------- DDS
diag_type StringToDiagnostic(const std::string &s) {
  if (s == "NASH_SUTCLIFFE") return DIAG_NASH_SUTCLIFFE;
  if (s == "RMSE") return DIAG_RMSE;
  if (s == "KLING_GUPTA") return DIAG_KLING_GUPTA;
  return DIAG_UNRECOGNIZED;
}
double CalculateDiagnostic(diag_type type, const std::vector<double> &mod, const std::vector<double> &obs, double starttime, double endtime,
                           double timestep, bool is_hydrograph) {
  const int nSampVal = (int)obs.size();
  auto index_of = [&](double t_mod) {   // CTimeSeries::GetTimeIndexFromModelTime (src/TimeSeries.cpp:311-314)
    int i = (int)(rvh_max(floor((t_mod + TIME_CORRECTION) / timestep), 0.0));
    return (i < nSampVal - 1) ? i : nSampVal - 1;
  };
  const int skip = is_hydrograph ? 1 : 0;   // averaged hydrographs: the value at index 0 is not a period average
  const int nnstart = index_of(starttime) + skip, nnend = index_of(endtime) + 1;
  std::vector<double> w(nnend > 0 ? nnend : 1, 0.0);
  for (int nn = nnstart; nn < nnend; nn++) {
    w[nn] = 1.0;
    if (obs[nn] == RAV_BLANK_DATA) w[nn] = 0.0;
    if (nn >= (int)mod.size() || mod[nn] == RAV_BLANK_DATA) w[nn] = 0.0;
  }
  double N = 0;
  if (type == DIAG_NASH_SUTCLIFFE) {
    double avgobs = 0.0;
    for (int nn = nnstart; nn < nnend; nn++) { avgobs += w[nn] * obs[nn]; N += w[nn]; }
    if (N > 0.0) avgobs /= N;
    double sum1 = 0.0, sum2 = 0.0;
    for (int nn = nnstart; nn < nnend; nn++) {
      const double m = (nn < (int)mod.size()) ? mod[nn] : 0.0;
      sum1 += w[nn] * pow(obs[nn] - m, 2);
      sum2 += w[nn] * pow(obs[nn] - avgobs, 2);
    }
    return (N > 0) ? 1.0 - (sum1 / sum2) : -ALMOST_INF;
  }
  if (type == DIAG_RMSE) {
    double sum = 0.0;
    for (int nn = nnstart; nn < nnend; nn++) { const double m = (nn < (int)mod.size()) ? mod[nn] : 0.0; sum += w[nn] * pow(obs[nn] - m, 2); N += w[nn]; }
    return (N > 0.0) ? sqrt(sum / N) : -ALMOST_INF;
  }
  if (type == DIAG_KLING_GUPTA) {
    double ObsSum = 0, ModSum = 0, ObsStd = 0, ModStd = 0, Cov = 0;
    for (int nn = nnstart; nn < nnend; nn++) { const double m = (nn < (int)mod.size()) ? mod[nn] : 0.0; ObsSum += obs[nn] * w[nn]; ModSum += m * w[nn]; N += w[nn]; }
    const double ObsAvg = ObsSum / N, ModAvg = ModSum / N;
    for (int nn = nnstart; nn < nnend; nn++) {
      const double m = (nn < (int)mod.size()) ? mod[nn] : 0.0;
      ObsStd += pow((obs[nn] - ObsAvg), 2) * w[nn];
      ModStd += pow((m - ModAvg), 2) * w[nn];
      Cov += (obs[nn] - ObsAvg) * (m - ModAvg) * w[nn];
    }
    ObsStd = sqrt(ObsStd / N); ModStd = sqrt(ModStd / N); Cov /= N;
    const double r = Cov / ObsStd / ModStd, Beta = ModAvg / ObsAvg, Alpha = ModStd / ObsStd;
    if ((N > 0) && ((ObsAvg != 0.0) || (Beta == 1.0)) && (ObsStd != 0.0) && (ModStd != 0.0)) return 1.0 - sqrt(pow((r - 1), 2) + pow((Alpha - 1), 2) + pow((Beta - 1), 2));
    return -ALMOST_INF;
  }
  ExitGracefully("CalculateDiagnostic: this diagnostic is not implemented in the B200 build");
  return 0.0;
}
CDDSEnsemble::CDDSEnsemble(int num_members, const optStruct &Options)
    : CEnsemble(num_members, Options), _Fbest(ALMOST_INF), _r_val(0.2), _calib_SBID(DOESNT_EXIST), _calib_Obj(DIAG_NASH_SUTCLIFFE), _calib_Period("ALL") {
  _type = ENSEMBLE_DDS;
}
void CDDSEnsemble::SetPerturbationValue(double perturb) {
  _r_val = perturb;
  ExitGracefullyIf((_r_val < 0) || (_r_val > 1.0), "CDDSEnsemble::SetPerturbationValue");
}
void CDDSEnsemble::SetCalibrationTarget(long long SBID, diag_type object_diag, const std::string &period) { _calib_SBID = SBID; _calib_Obj = object_diag; _calib_Period = period; }
void CDDSEnsemble::Initialize(const CModel *pModel, const optStruct &Options) {
  (void)pModel; (void)Options;
  ExitGracefullyIf(_calib_SBID == DOESNT_EXIST, "CDDSEnsemble::Initialize: DDS calibration target/objective function must be set");
  ExitGracefullyIf(_nMembers == 0, "CDDSEnsemble::Initialize: number of DDS iterations/ensemble members must be >0");
  ExitGracefullyIf(_ParamDists.empty(), "CDDSEnsemble::Initialize: at least one parameter distribution must be set using :ParameterDistributions command");
  _BestParams.assign(_ParamDists.size(), 0.0); _TestParams.assign(_ParamDists.size(), 0.0);
  for (size_t i = 0; i < _ParamDists.size(); i++) _BestParams[i] = _TestParams[i] = _ParamDists[i].default_val;
  _Fbest = ALMOST_INF; _log.clear();
}
double CDDSEnsemble::PerturbParam(double x_best, double lowerbound, double upperbound) {   // src/DDS.cpp:121-148, reflecting bounds
  const double x_min = lowerbound, x_max = upperbound;
  double x_new = x_best + GaussRandom() * _r_val * (x_max - x_min);
  if (x_new < x_min) { x_new = x_min + (x_min - x_new); if (x_new > x_max) x_new = x_min; }
  else if (x_new > x_max) { x_new = x_max - (x_new - x_max); if (x_new < x_min) x_new = x_max; }
  return x_new;
}
void CDDSEnsemble::UpdateModel(CModel *pModel, optStruct &Options, int e) {   // src/DDS.cpp:150-208
  CEnsemble::UpdateModel(pModel, Options, e);
  ExitGracefullyIf(e >= _nMembers, "CDDSEnsemble::UpdateMode: invalid ensemble member index");
  const int nP = (int)_ParamDists.size();
  const double Pn = 1.0 - log(double(e)) / log(double(_nMembers));   // probability of including a decision variable
  int dvn_count = 0;
  for (int k = 0; k < nP; k++) _TestParams[k] = _BestParams[k];
  for (int k = 0; k < nP; k++) {
    const double u = UniformRandom();
    if (u < Pn) { dvn_count++; _TestParams[k] = PerturbParam(_BestParams[k], _ParamDists[k].distpar[0], _ParamDists[k].distpar[1]); }
  }
  if (dvn_count == 0) {
    const double u = UniformRandom();
    const int dv = (int)(ceil((double)(nP)*u)) - 1;
    _TestParams[dv] = PerturbParam(_BestParams[dv], _ParamDists[dv].distpar[0], _ParamDists[dv].distpar[1]);
  }
  for (int k = 0; k < nP; k++) pModel->UpdateParameter(_ParamDists[k].param_class, _ParamDists[k].param_name, _ParamDists[k].class_group, _TestParams[k]);
  // the reference re-reads the .rvc here: every trial starts from the same initial state (uploaded by the caller)
}
void CDDSEnsemble::FinishEnsembleRun(double Ftest, int e) {   // src/DDS.cpp:210-240; minimisation
  if (Ftest <= _Fbest) {
    _Fbest = Ftest;
    _BestParams = _TestParams;
    _log.push_back(std::make_pair(e + 1, _Fbest));
  }
}
double CDDSEnsemble::ObjectiveFromHydrograph(const CModel *pModel, const optStruct &Options, const double *qout_rec, const double *qout_initial, int nsteps, int nSB) const {
  // CModel::GetObjFuncVal (src/StandardOutput.cpp:3282-3322)
  double starttime = 0.0, endtime = 0.0;
  for (size_t d = 0; d < pModel->diag_periods.size(); d++)
    if (pModel->diag_periods[d].name == _calib_Period) { starttime = pModel->diag_periods[d].t_start; endtime = pModel->diag_periods[d].t_end; }
  const observed_ts *pObs = NULL;
  for (size_t i = 0; i < pModel->observed.size(); i++) if (pModel->observed[i].name == "HYDROGRAPH" && pModel->observed[i].loc_ID == _calib_SBID) pObs = &pModel->observed[i];
  ExitGracefullyIf(pObs == NULL, "GetObjFuncVal: unable to find calibration target time series (hydrograph in basin :CalibrationSBID)");
  ExitGracefullyIf(endtime == 0.0, "GetObjFuncVal: unable to find calibration period with this name ");
  const int p = pModel->GetSubBasinIndex(_calib_SBID);
  // modelled HYDROGRAPH at time index nn (CModel::UpdateDiagnostics, src/Model.cpp:2929-2936): period average of the outflow for nn >= 1,
  // the initial outflow at nn = 0
  std::vector<double> mod(nsteps + 1), obs(pObs->pTS->GetNumSampledValues());
  mod[0] = qout_initial[p];
  for (int nn = 1; nn <= nsteps; nn++) {
    const double Qout = qout_rec[(size_t)(nn - 1) * nSB + p], QoutLast = (nn >= 2) ? qout_rec[(size_t)(nn - 2) * nSB + p] : qout_initial[p];
    mod[nn] = (0.5 * (Qout + QoutLast) * (Options.timestep * SEC_PER_DAY)) / (Options.timestep * SEC_PER_DAY);
  }
  for (size_t nn = 0; nn < obs.size(); nn++) obs[nn] = pObs->pTS->GetSampledValue((int)nn);
  double objval = CalculateDiagnostic(_calib_Obj, mod, obs, starttime, endtime, Options.timestep, true);
  if ((_calib_Obj == DIAG_NASH_SUTCLIFFE) || (_calib_Obj == DIAG_KLING_GUPTA)) objval *= -1;
  return objval;
}

void CDDSEnsemble::ObjectiveInputs(const CModel *pModel, const optStruct &Options, int &basin, int &type, int &nnstart, int &nnend, std::vector<double> &obs) const {
  double starttime = 0.0, endtime = 0.0;
  for (size_t d = 0; d < pModel->diag_periods.size(); d++)
    if (pModel->diag_periods[d].name == _calib_Period) { starttime = pModel->diag_periods[d].t_start; endtime = pModel->diag_periods[d].t_end; }
  const observed_ts *pObs = NULL;
  for (size_t i = 0; i < pModel->observed.size(); i++) if (pModel->observed[i].name == "HYDROGRAPH" && pModel->observed[i].loc_ID == _calib_SBID) pObs = &pModel->observed[i];
  ExitGracefullyIf(pObs == NULL, "GetObjFuncVal: unable to find calibration target time series (hydrograph in basin :CalibrationSBID)");
  ExitGracefullyIf(endtime == 0.0, "GetObjFuncVal: unable to find calibration period with this name ");
  basin = pModel->GetSubBasinIndex(_calib_SBID);
  type = (_calib_Obj == DIAG_NASH_SUTCLIFFE) ? 0 : (_calib_Obj == DIAG_RMSE) ? 1 : 2;
  obs.resize(pObs->pTS->GetNumSampledValues());
  for (size_t nn = 0; nn < obs.size(); nn++) obs[nn] = pObs->pTS->GetSampledValue((int)nn);
  const int nSampVal = (int)obs.size();
  auto index_of = [&](double t_mod) { int i = (int)(rvh_max(floor((t_mod + TIME_CORRECTION) / Options.timestep), 0.0)); return (i < nSampVal - 1) ? i : nSampVal - 1; };
  nnstart = index_of(starttime) + 1;   // averaged hydrographs: the value at index 0 is not a period average
  nnend = index_of(endtime) + 1;
}

// ---------------------------------------------------------------------------------------------------------------- EnKF
CEnKFEnsemble::CEnKFEnsemble(int num_members, const optStruct &Options)
    : CEnsemble(num_members, Options), _EnKF_mode(ENKF_UNSPECIFIED), _window_size(1), _nTimeSteps(0), _nObsDatapoints(0), _member0(0), _nLocal(num_members) {
  _type = ENSEMBLE_ENKF;
}
void CEnKFEnsemble::AddObsPerturbation(sv_type type, int distrib, const double *distpars, adjustment adj) {
  obs_perturb p; p.state = type; p.distribution = distrib; p.kk = DOESNT_EXIST; p.adj_type = adj;
  for (int i = 0; i < 3; i++) p.distpar[i] = distpars[i];
  _pObsPerturbations.push_back(p);
}
void CEnKFEnsemble::AddAssimilationState(sv_type sv, int lay, int assim_groupID, bool is_streamflow) {
  assim_state a; a.sv = sv; a.layer = lay; a.group = assim_groupID; a.is_streamflow = is_streamflow;
  _aAssim.push_back(a);
}
void CEnKFEnsemble::Initialize(const CModel *pModel, const optStruct &Options) {   // src/EnKF.cpp:197-418
  ExitGracefullyIf(_nMembers <= 1, "CEnKFEnsemble::Initialize: number of EnKF ensemble members must be >1");
  const bool forecast = (_EnKF_mode == ENKF_FORECAST) || (_EnKF_mode == ENKF_OPEN_FORECAST);
  ExitGracefullyIf((pModel->GetNumForcingPerturbations() == 0) && !forecast,
                   "CEnKFEnsemble::Initialize: at least one forcing perturbation must be set using :ForcingPerturbation command");
  ExitGracefullyIf(_aAssim.empty() && !forecast, "CEnKFEnsemble::Initialize: at least one assimilated state must be set in :ForcingPerturbation command");
  // rows of the state matrix, in the order of AddToStateMatrix / UpdateFromStateMatrix (:602-747)
  _rows.clear(); _state_names.clear();
  for (size_t i = 0; i < _aAssim.size(); i++) {
    const assim_state &a = _aAssim[i];
    if (a.is_streamflow) {
      const CSubbasinGroup &G = pModel->sb_groups[a.group];
      for (size_t pp = 0; pp < G.sb.size(); pp++) {
        const int p = G.sb[pp];
        const CSubBasin *pBasin = pModel->GetSubBasin(p);
        for (int n = 0; n < pBasin->GetInflowHistorySize(); n++) { rvn_enkf_row r = {RVN_XROW_QIN_HIST, p, n, 0}; _rows.push_back(r); _state_names.push_back("inflow_" + std::to_string(pp) + "_" + std::to_string(n)); }
        for (int n = 0; n < pBasin->GetNumSegments(); n++) { rvn_enkf_row r = {RVN_XROW_QOUT, p, n, 0}; _rows.push_back(r); _state_names.push_back("outflow_" + std::to_string(pp) + "_" + std::to_string(n)); }
        rvn_enkf_row r = {RVN_XROW_QOUT_LAST, p, 0, 0}; _rows.push_back(r); _state_names.push_back("outflowlast_" + std::to_string(pp));
      }
    } else {
      const CHRUGroup &G = pModel->hru_groups[a.group];
      const int iii = pModel->GetStateVarIndex(a.sv, a.layer);
      for (size_t n = 0; n < G.hru.size(); n++) {
        rvn_enkf_row r = {RVN_XROW_HRU_SV, G.hru[n], iii, 0}; _rows.push_back(r);
        _state_names.push_back(pModel->GetStateVarName(iii) + "_" + std::to_string(pModel->GetHydroUnit(G.hru[n])->ID));
      }
    }
  }
  // observations available in the data window
  _nTimeSteps = (int)(floor((Options.duration + TIME_CORRECTION) / Options.timestep));
  _window_size = (_window_size <= _nTimeSteps) ? _window_size : _nTimeSteps;
  std::vector<int> aObsIndices;
  _obs_basin.clear(); _obs_nn.clear(); _series_basin.clear(); _series_valid.clear();
  for (int i = 0; i < pModel->GetNumObservedTS(); i++) {
    const observed_ts &o = pModel->observed[i];
    const int p = pModel->GetSubBasinIndex(o.loc_ID);
    const bool good = (p >= 0) && pModel->GetSubBasin(p)->assimilate;
    if (good && o.name == "HYDROGRAPH") {
      aObsIndices.push_back(i);
      _series_basin.push_back(p);
      for (int nn = _nTimeSteps - _window_size + 1; nn <= _nTimeSteps; nn++) {
        const bool valid = (o.pTS->GetSampledValue(nn) != RAV_BLANK_DATA);
        _series_valid.push_back(valid ? 1 : 0);
        if (valid) { _obs_basin.push_back(p); _obs_nn.push_back(nn); }
      }
    }
  }
  _nObsDatapoints = (int)_obs_basin.size();
  ExitGracefullyIf(aObsIndices.empty() && !forecast,
                   "EnKF::Initialize : no observations available for assimilation. Use :AssimilateStreamflow command to indicate enabled observation time series");
  if ((_nObsDatapoints == 0) && !forecast) WriteWarning("EnKF::Initialize : no observations were available for assimilation. EnKF assimilation updates will not be performed.");
  // perturbed observations: one draw per (member, datapoint) in member-major order (:372-405)
  _obs_matrix.assign((size_t)_nMembers * _nObsDatapoints, 0.0);
  _noise_matrix.assign((size_t)_nMembers * _nObsDatapoints, 0.0);
  for (int e = 0; e < _nMembers; e++) {
    int j = 0;
    for (size_t ii = 0; ii < aObsIndices.size(); ii++) {
      const observed_ts &o = pModel->observed[aObsIndices[ii]];
      const obs_perturb *pPerturb = NULL;   // HYDROGRAPH observations are perturbed by the STREAMFLOW error model; the last one given wins
      for (size_t q = 0; q < _pObsPerturbations.size(); q++) if (_pObsPerturbations[q].state == (sv_type)RVN_SV_NTYPES) pPerturb = &_pObsPerturbations[q];
      for (int nn = _nTimeSteps - _window_size + 1; nn <= _nTimeSteps; nn++) {
        const double obsval = o.pTS->GetSampledValue(nn);
        if (obsval != RAV_BLANK_DATA) {
          double noise = 0;
          if (pPerturb != NULL) {
            const double eps = SampleFromDistribution(pPerturb->distribution, pPerturb->distpar);
            if (pPerturb->adj_type == ADJ_ADDITIVE) noise = eps;
            else if (pPerturb->adj_type == ADJ_MULTIPLICATIVE) noise = (eps * obsval) - obsval;
          }
          _noise_matrix[(size_t)e * _nObsDatapoints + j] = noise;
          _obs_matrix[(size_t)e * _nObsDatapoints + j] = obsval + noise;
          j++;
        }
      }
    }
  }
}
void CEnKFEnsemble::UpdateModel(CModel *pModel, optStruct &Options, int e) {
  CEnsemble::UpdateModel(pModel, Options, e);
  ExitGracefullyIf(e >= _nMembers, "CEnKFEnsemble::UpdateMode: invalid ensemble member index");
  if ((_EnKF_mode == ENKF_FORECAST) || (_EnKF_mode == ENKF_OPEN_FORECAST)) return;   // no perturbations drawn (src/EnKF.cpp:428-431)
  // the window's forcing-perturbation samples of member e, in the order StartTimeStepOps draws them: step by step, perturbation
  // by perturbation, nStepsPerDay (= 1 for the daily-or-longer steps supported here) samples each
  const int nP = pModel->GetNumForcingPerturbations();
  const int nsteps = (int)((Options.duration - TIME_CORRECTION) / Options.timestep) + 1;   // iterations of the reference time loop (src/RavenMain.cpp:147)
  for (int n = 0; n < nsteps; n++)
    for (int i = 0; i < nP; i++) {
      const force_perturb &P = pModel->perturbations[i];
      const double eps = SampleFromDistribution(P.distribution, P.distpar);
      const bool matched = !(P.forcing == F_TEMP_DAILY_MIN || P.forcing == F_TEMP_DAILY_MAX);
      if (matched && e >= _member0 && e < _member0 + _nLocal) pModel->SetPerturbationSample(e - _member0, n, i, eps);
    }
}
void CEnKFEnsemble::HarvestOutputs(const double *qout_rec, const double *qout_initial, int nSB, double timestep, double *out_row) const {
  // CloseTimeStepOps (src/EnKF.cpp:440-472) replayed for every step of the data window.  The modelled HYDROGRAPH at time index
  // nn is CSubBasin::GetIntegratedOutflow(tstep)/(tstep*SEC_PER_DAY) after step nn-1 (src/Model.cpp:2929-2936,
  // src/SubBasin.cpp:929-934; Options.ave_hydrograph is on by default).  The reference's slot counter j stops advancing at the
  // `break` once the current step is found, so with a window > 1 and several gauges later series land in earlier slots and
  // some slots keep their initial 0 -- reproduced as is (the analysis is defined on these numbers).
  for (int j = 0; j < _nObsDatapoints; j++) out_row[j] = 0.0;
  const int nSeries = (int)_series_basin.size();
  for (int curr_nn = 1; curr_nn <= _nTimeSteps; curr_nn++) {
    if (curr_nn < _nTimeSteps - _window_size + 1) continue;
    int j = 0;
    for (int ii = 0; ii < nSeries; ii++) {
      const int p = _series_basin[ii];
      for (int nn = _nTimeSteps - _window_size + 1; nn <= _nTimeSteps; nn++) {
        if (_series_valid[(size_t)ii * _window_size + (nn - (_nTimeSteps - _window_size + 1))]) {
          if (nn == curr_nn) {
            const double Qout = qout_rec[(size_t)(nn - 1) * nSB + p];
            const double QoutLast = (nn >= 2) ? qout_rec[(size_t)(nn - 2) * nSB + p] : qout_initial[p];
            out_row[j] = (0.5 * (Qout + QoutLast) * (timestep * SEC_PER_DAY)) / (timestep * SEC_PER_DAY);
            j++;
            break;
          } else j++;
        }
      }
    }
  }
}
// ---------------------------------------------------------------------------------------------------------------------
// Truncated least-squares solve of P x = b for many right-hand sides, the way the reference does it (SVD(), src/Matrix.cpp:144-390:
// the Golub-Reinsch singular value decomposition as published in Numerical Recipes' svdcmp/svbksb, singular values below
// tol * max set to zero).  P of the EnKF analysis is a sample covariance of strongly correlated flows with condition numbers up to
// the 1e8 cut-off, where two different (individually backward-stable) factorisations agree to ~1e-8 only; the 1e-9 contract of
// the analysis therefore needs the reference's own sequence of Householder reflections and implicit-shift QR sweeps, which this
// restates operation by operation (0-based, flat row-major storage; the reference decomposes P once per member with identical
// results - here once per analysis).
// ---------------------------------------------------------------------------------------------------------------------
namespace {
inline double copysign_nr(double a, double b) { return b >= 0.0 ? (a >= 0.0 ? a : -a) : (a >= 0.0 ? -a : a); }   // SIGN(a,b)
inline double hypot_nr(double a, double b) {                                                                     // pythag(a,b)
  const double absa = fabs(a), absb = fabs(b);
  if (absa > absb) return absa * sqrt(1.0 + (absb / absa) * (absb / absa));
  return (absb == 0.0) ? 0.0 : absb * sqrt(1.0 + (absa / absb) * (absa / absb));
}
struct GolubReinsch {
  int n;
  std::vector<double> u, v, w;   // P = u diag(w) v'
  double &U(int i, int j) { return u[(size_t)i * n + j]; }
  double &V(int i, int j) { return v[(size_t)i * n + j]; }

  void Decompose(const std::vector<double> &P, int n_) {
    n = n_;
    u = P; v.assign((size_t)n * n, 0.0); w.assign(n, 0.0);
    std::vector<double> e(n, 0.0);   // super-diagonal of the bidiagonal form (the reference's tmp[] / rv1)
    double g = 0.0, scale = 0.0, anorm = 0.0;
    int l = 0;
    // ---- Householder reduction to bidiagonal form
    for (int i = 0; i < n; i++) {
      l = i + 2;
      e[i] = scale * g;
      g = 0.0; scale = 0.0;
      double s_ = 0.0;
      for (int k = i; k < n; k++) scale += fabs(U(k, i));
      if (scale != 0.0) {
        for (int k = i; k < n; k++) { U(k, i) /= scale; s_ += U(k, i) * U(k, i); }
        double f = U(i, i);
        g = -copysign_nr(sqrt(s_), f);
        const double h = f * g - s_;
        U(i, i) = f - g;
        for (int j = l - 1; j < n; j++) {
          s_ = 0.0;
          for (int k = i; k < n; k++) s_ += U(k, i) * U(k, j);
          f = s_ / h;
          for (int k = i; k < n; k++) U(k, j) += f * U(k, i);
        }
        for (int k = i; k < n; k++) U(k, i) *= scale;
      }
      w[i] = scale * g;
      g = 0.0; scale = 0.0; s_ = 0.0;
      {   // row transformation (the reference's guard `i+1<=m && i!=n` always holds for a square matrix)
        for (int k = l - 1; k < n; k++) scale += fabs(U(i, k));
        if (scale != 0.0) {
          for (int k = l - 1; k < n; k++) { U(i, k) /= scale; s_ += U(i, k) * U(i, k); }
          const double f = U(i, l - 1);
          g = -copysign_nr(sqrt(s_), f);
          const double h = f * g - s_;
          U(i, l - 1) = f - g;
          for (int k = l - 1; k < n; k++) e[k] = U(i, k) / h;
          for (int j = l - 1; j < n; j++) {
            s_ = 0.0;
            for (int k = l - 1; k < n; k++) s_ += U(j, k) * U(i, k);
            for (int k = l - 1; k < n; k++) U(j, k) += s_ * e[k];
          }
          for (int k = l - 1; k < n; k++) U(i, k) *= scale;
        }
      }
      const double an = fabs(w[i]) + fabs(e[i]);
      if (an > anorm) anorm = an;
    }
    // ---- accumulation of the right-hand transformations
    for (int i = n - 1; i >= 0; i--) {
      if (i < n - 1) {
        if (g != 0.0) {
          for (int j = l; j < n; j++) V(j, i) = (U(i, j) / U(i, l)) / g;   // double division avoids underflow
          for (int j = l; j < n; j++) {
            double s_ = 0.0;
            for (int k = l; k < n; k++) s_ += U(i, k) * V(k, j);
            for (int k = l; k < n; k++) V(k, j) += s_ * V(k, i);
          }
        }
        for (int j = l; j < n; j++) V(i, j) = V(j, i) = 0.0;
      }
      V(i, i) = 1.0;
      g = e[i];
      l = i;
    }
    // ---- accumulation of the left-hand transformations
    for (int i = n - 1; i >= 0; i--) {
      l = i + 1;
      g = w[i];
      for (int j = l; j < n; j++) U(i, j) = 0.0;
      if (g != 0.0) {
        g = 1.0 / g;
        for (int j = l; j < n; j++) {
          double s_ = 0.0;
          for (int k = l; k < n; k++) s_ += U(k, i) * U(k, j);
          const double f = (s_ / U(i, i)) * g;
          for (int k = i; k < n; k++) U(k, j) += f * U(k, i);
        }
        for (int j = i; j < n; j++) U(j, i) *= g;
      } else {
        for (int j = i; j < n; j++) U(j, i) = 0.0;
      }
      ++U(i, i);
    }
    // ---- diagonalisation of the bidiagonal form: implicit-shift QR, at most 30 sweeps per singular value
    for (int k = n - 1; k >= 0; k--) {
      for (int its = 0; its < 30; its++) {
        bool cancel = true;
        int nm = 0;
        for (l = k; l >= 0; l--) {   // test for splitting (e[0] is always zero)
          nm = l - 1;
          if (fabs(e[l]) + anorm == anorm) { cancel = false; break; }
          if (fabs(w[nm]) + anorm == anorm) break;
        }
        if (cancel) {   // cancellation of e[l], l > 0
          double c = 0.0, s_ = 1.0;
          for (int i = l - 1; i < k + 1; i++) {   // (the reference starts this loop one row early, at i = l-1; kept)
            const double f = s_ * e[i];
            e[i] *= c;
            if (fabs(f) + anorm == anorm) break;
            g = w[i];
            double h = hypot_nr(f, g);
            w[i] = h;
            h = 1.0 / h;
            c = g * h;
            s_ = -f * h;
            for (int j = 0; j < n; j++) {
              const double y = U(j, nm), z = U(j, i);
              U(j, nm) = y * c + z * s_;
              U(j, i) = z * c - y * s_;
            }
          }
        }
        double z = w[k];
        if (l == k) {   // converged: make the singular value non-negative
          if (z < 0.0) { w[k] = -z; for (int j = 0; j < n; j++) V(j, k) = -V(j, k); }
          break;
        }
        ExitGracefullyIf(its == 29, "Singular Value Decomposition did not converge");
        double x = w[l];   // shift from the bottom 2x2 minor
        nm = k - 1;
        double y = w[nm];
        g = e[nm];
        double h = e[k];
        double f = ((y - z) * (y + z) + (g - h) * (g + h)) / (2.0 * h * y);
        g = hypot_nr(f, 1.0);
        f = ((x - z) * (x + z) + h * ((y / (f + copysign_nr(g, f))) - h)) / x;
        double c = 1.0, s_ = 1.0;
        for (int j = l; j <= nm; j++) {   // next QR transformation
          const int i = j + 1;
          g = e[i];
          y = w[i];
          h = s_ * g;
          g = c * g;
          z = hypot_nr(f, h);
          e[j] = z;
          c = f / z;
          s_ = h / z;
          f = x * c + g * s_;
          g = g * c - x * s_;
          h = y * s_;
          y *= c;
          for (int jj = 0; jj < n; jj++) {
            x = V(jj, j); z = V(jj, i);
            V(jj, j) = x * c + z * s_;
            V(jj, i) = z * c - x * s_;
          }
          z = hypot_nr(f, h);
          w[j] = z;
          if (z != 0.0) { z = 1.0 / z; c = f * z; s_ = h * z; }   // rotation can be arbitrary if z = 0
          f = c * g + s_ * y;
          x = c * y - s_ * g;
          for (int jj = 0; jj < n; jj++) {
            y = U(jj, j); z = U(jj, i);
            U(jj, j) = y * c + z * s_;
            U(jj, i) = z * c - y * s_;
          }
        }
        e[l] = 0.0;
        e[k] = f;
        w[k] = x;
      }
    }
  }
  // zero the singular values below tol * max, as the reference does before back-substituting
  void Truncate(double tol) {
    double maxw = -ALMOST_INF;
    for (int j = 0; j < n; j++) if (w[j] > maxw) maxw = w[j];
    for (int j = 0; j < n; j++) if (fabs(w[j] / maxw) < tol) w[j] = 0.0;
  }
  // x = V diag(1/w) U' b   (svbksb)
  void Solve(const double *b, int bstride, double *x, int xstride, std::vector<double> &t) {
    t.assign(n, 0.0);
    for (int j = 0; j < n; j++) {
      double s_ = 0.0;
      if (w[j] != 0.0) {
        for (int i = 0; i < n; i++) s_ += U(i, j) * b[(size_t)i * bstride];
        s_ /= w[j];
      }
      t[j] = s_;
    }
    for (int j = 0; j < n; j++) {
      double s_ = 0.0;
      for (int jj = 0; jj < n; jj++) s_ += V(j, jj) * t[jj];
      x[(size_t)j * xstride] = s_;
    }
  }
};
}  // namespace
void SolveSymmetricTruncated(const std::vector<double> &P, int n, double tol, const std::vector<double> &B, int nrhs, std::vector<double> &X) {
  GolubReinsch svd;
  svd.Decompose(P, n);
  svd.Truncate(tol);
  X.assign((size_t)n * nrhs, 0.0);
  std::vector<double> t;
  for (int r = 0; r < nrhs; r++) svd.Solve(&B[r], nrhs, &X[r], nrhs, t);   // B, X are [n][nrhs]
}
bool CEnKFEnsemble::AssimilationCalcs(const double *out_all, std::vector<double> &Z) const {   // src/EnKF.cpp:478-575
  const int N = _nMembers, Nobs = _nObsDatapoints;
  Z.assign((size_t)N * N, 0.0);
  if (Nobs == 0) return false;
  std::vector<double> HA((size_t)Nobs * N), P((size_t)Nobs * Nobs), tmp((size_t)Nobs * Nobs), D((size_t)Nobs * N), MM;
  for (int j = 0; j < Nobs; j++) {   // HA = O_sim - mean
    double outMean = 0;
    for (int i = 0; i < N; i++) outMean += out_all[(size_t)i * Nobs + j] / N;
    for (int i = 0; i < N; i++) HA[(size_t)j * N + i] = out_all[(size_t)i * Nobs + j] - outMean;
  }
  // P = 1/(N-1) * (HA*HA' + eQ*eQ'), each product accumulated over the members in order (MatMult, src/Matrix.cpp:38-49)
  for (int a = 0; a < Nobs; a++)
    for (int b = 0; b < Nobs; b++) {
      double s1 = 0.0, s2 = 0.0;
      for (int i = 0; i < N; i++) s1 += HA[(size_t)a * N + i] * HA[(size_t)b * N + i];
      for (int i = 0; i < N; i++) s2 += _noise_matrix[(size_t)i * Nobs + a] * _noise_matrix[(size_t)i * Nobs + b];
      P[(size_t)a * Nobs + b] = (s1 + s2) * (1.0 / (N - 1));
    }
  (void)tmp;
  for (int i = 0; i < N; i++) for (int j = 0; j < Nobs; j++) D[(size_t)j * N + i] = _obs_matrix[(size_t)i * Nobs + j] - out_all[(size_t)i * Nobs + j];
  SolveSymmetricTruncated(P, Nobs, 1e-8, D, N, MM);   // MM [Nobs][N] = inv(P)*(O_obs-O_sim)
  for (int a = 0; a < N; a++)                           // Z = HA' * MM
    for (int b = 0; b < N; b++) {
      double s_ = 0.0;
      for (int j = 0; j < Nobs; j++) s_ += HA[(size_t)j * N + a] * MM[(size_t)j * N + b];
      Z[(size_t)a * N + b] = s_;
    }
  return true;
}

CEnsemble *CreateEnsemble(CModel *pModel, const optStruct &Options) {
  CEnsemble *pE = NULL;
  if (Options.ensemble == ENSEMBLE_MONTECARLO) {
    CMonteCarloEnsemble *mc = new CMonteCarloEnsemble(pModel->num_members, Options);
    for (size_t i = 0; i < pModel->param_dists.size(); i++) mc->AddParamDist(pModel->param_dists[i]);
    pE = mc;
  } else if (Options.ensemble == ENSEMBLE_DDS) {
    CDDSEnsemble *dds = new CDDSEnsemble(pModel->num_members, Options);
    for (size_t i = 0; i < pModel->param_dists.size(); i++) dds->AddParamDist(pModel->param_dists[i]);
    if (pModel->calib_SBID != DOESNT_EXIST) dds->SetCalibrationTarget(pModel->calib_SBID, (diag_type)pModel->calib_obj, pModel->calib_period);
    pE = dds;
  } else if (Options.ensemble == ENSEMBLE_ENKF) {   // src/ParseInput.cpp:3830 + the .rve commands (src/ParseEnsembleFile.cpp)
    CEnKFEnsemble *en = new CEnKFEnsemble(pModel->num_members, Options);
    en->SetEnKFMode(pModel->enkf.mode);
    en->SetWindowSize(pModel->enkf.window_size);
    for (size_t i = 0; i < pModel->enkf.obs_perturbations.size(); i++) {
      const obs_perturb &o = pModel->enkf.obs_perturbations[i];
      en->AddObsPerturbation(o.state, o.distribution, o.distpar, o.adj_type);
    }
    for (size_t i = 0; i < pModel->enkf.assim_states.size(); i++) {
      const assim_state &a = pModel->enkf.assim_states[i];
      en->AddAssimilationState(a.sv, a.layer, a.group, a.is_streamflow);
    }
    pE = en;
  } else if (Options.ensemble == ENSEMBLE_NONE) {
    pE = new CEnsemble(1, Options);
  } else {
    ExitGracefully("CreateEnsemble: this ensemble mode is not implemented in the B200 build yet");
  }
  pE->SetRandomSeed(Options.random_seed);
  return pE;
}
