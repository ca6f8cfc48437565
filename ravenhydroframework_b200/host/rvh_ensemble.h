// rvh_ensemble.h -- ensemble member management with the reference's plugin surface (src/ModelEnsemble.h:34-95).
//
// The reference runs members one after another on the same CModel (src/RavenMain.cpp:115-196): UpdateModel(e) mutates
// class parameters in place, the whole simulation runs, then member e+1.  Here UpdateModel(e) is unchanged in meaning
// (same libc rand() stream, same draw order, same CModel::UpdateParameter calls) but it is called for ALL members before
// the time loop; after each call the member's parameter row is flattened and uploaded (rvn_gpu_upload_params), and the
// members then advance together on the GPU.  Members shard over GPUs by contiguous blocks: every rank replays the draws
// of all members up to the end of its block (they are consumed in member-serial order) and uploads its own.
#ifndef RVH_ENSEMBLE_H
#define RVH_ENSEMBLE_H
#include "rvh_model.h"

double UniformRandom();                                        // src/ModelEnsemble.cpp:10-13
double GaussRandom();                                          // src/ModelEnsemble.cpp:14-19
double SampleFromGamma(const double &shape, const double &scale);      // src/ModelEnsemble.cpp:265-292
double SampleFromDistribution(int distribution, const double distpar[3]);  // src/ModelEnsemble.cpp:310-340
enum disttype { DIST_UNIFORM = 0, DIST_NORMAL = 1, DIST_LOGNORMAL = 2, DIST_GAMMA = 3 };

class CEnsemble {
public:
  CEnsemble(int num_members, const optStruct &Options);
  virtual ~CEnsemble() {}
  int  GetNumMembers() const { return _nMembers; }
  ensemble_type GetType() const { return _type; }
  virtual double GetStartTime(int e) const { (void)e; return 0.0; }
  void SetRandomSeed(unsigned int seed);                       // srand(seed), src/ModelEnsemble.cpp:107-110
  virtual void Initialize(const CModel *pModel, const optStruct &Options) { (void)pModel; (void)Options; }
  virtual void UpdateModel(CModel *pModel, optStruct &Options, int e);
  virtual void StartTimeStepOps(CModel *, optStruct &, const time_struct &, int) {}
  virtual void CloseTimeStepOps(CModel *, optStruct &, const time_struct &, int) {}
  virtual void FinishEnsembleRun(CModel *, optStruct &, const time_struct &, int) {}
protected:
  ensemble_type _type;
  int _nMembers;
};

class CMonteCarloEnsemble : public CEnsemble {
public:
  CMonteCarloEnsemble(int num_members, const optStruct &Options);
  void AddParamDist(const param_dist &dist) { _ParamDists.push_back(dist); }
  void Initialize(const CModel *pModel, const optStruct &Options) override;
  void UpdateModel(CModel *pModel, optStruct &Options, int e) override;
  const std::vector<double> &LastSample() const { return _last; }   // values drawn by the latest UpdateModel (MonteCarloOutput.csv row)
private:
  std::vector<param_dist> _ParamDists;
  std::vector<double> _last;
};

// DDS calibration (reference CDDSEnsemble, src/DDS.cpp; Tolson & Shoemaker 2007).  Trial e+1 perturbs the best parameter vector
// found so far, so the trials are strictly sequential: every trial is one full GPU run of one member.  UpdateModel(e) draws from
// the same rand() stream in the same order as the reference; FinishEnsembleRun(e) takes the trial's objective function value
// (CModel::GetObjFuncVal, src/StandardOutput.cpp:3282-3322: the named diagnostic of the HYDROGRAPH observation at the calibration
// subbasin, negated for NSE / KGE because DDS minimises).
enum diag_type { DIAG_NASH_SUTCLIFFE = 0, DIAG_RMSE, DIAG_KLING_GUPTA, DIAG_UNRECOGNIZED };
diag_type StringToDiagnostic(const std::string &s);
// CDiagnostic::CalculateDiagnostic (src/Diagnostics.cpp:93-200,423-445,1072-1160) for the three objective functions above:
// mod/obs are sampled values at time indices 0..n-1 (blank = RAV_BLANK_DATA), evaluated over [starttime, endtime] model days
double CalculateDiagnostic(diag_type type, const std::vector<double> &mod, const std::vector<double> &obs, double starttime, double endtime,
                           double timestep, bool is_hydrograph);
class CDDSEnsemble : public CEnsemble {
public:
  CDDSEnsemble(int num_members, const optStruct &Options);
  void AddParamDist(const param_dist &dist) { _ParamDists.push_back(dist); }
  void SetPerturbationValue(double perturb);
  void SetCalibrationTarget(long long SBID, diag_type object_diag, const std::string &period);
  void Initialize(const CModel *pModel, const optStruct &Options) override;
  void UpdateModel(CModel *pModel, optStruct &Options, int e) override;
  // the reference signature reads the objective from the model; here the caller hands over the modelled series of the trial
  void FinishEnsembleRun(double Ftest, int e);
  double ObjectiveFromHydrograph(const CModel *pModel, const optStruct &Options, const double *qout_rec, const double *qout_initial, int nsteps, int nSB) const;
  // what rvn_gpu_diagnostic needs to evaluate the same objective on the device: basin index, diagnostic, sample window, observations
  void ObjectiveInputs(const CModel *pModel, const optStruct &Options, int &basin, int &type, int &nnstart, int &nnend, std::vector<double> &obs) const;
  const std::vector<double> &TestParams() const { return _TestParams; }
  const std::vector<double> &BestParams() const { return _BestParams; }
  double BestObjective() const { return _Fbest; }
  const std::vector<std::pair<int, double> > &Improvements() const { return _log; }   // the rows of DDSOutput.csv: (trial e+1, Fbest)
private:
  double PerturbParam(double x_best, double lowerbound, double upperbound);
  std::vector<param_dist> _ParamDists;
  std::vector<double> _BestParams, _TestParams;
  std::vector<std::pair<int, double> > _log;
  double _Fbest, _r_val;
  long long _calib_SBID;
  diag_type _calib_Obj;
  std::string _calib_Period;
};

// EnKF ensemble (reference CEnKFEnsemble, src/EnKF.h:31-89, src/EnKF.cpp).  The reference runs the N members one after
// another over one assimilation window and then updates the state matrix on the host; here
//   * Initialize() builds the state-row map, finds the observation datapoints of the window and draws the observation
//     noise (same rand() stream, same order: src/EnKF.cpp:197-418);
//   * UpdateModel(e) draws member e's forcing-perturbation samples for every step of the window -- what the reference draws
//     step by step in StartTimeStepOps -> CModel::PrepareForcingPerturbation (src/EnKF.cpp:424-432, src/Model.cpp:2829-2847);
//     members are visited in order, so the stream position is the reference's;
//   * the members then advance together on the GPU; HarvestOutputs() = CloseTimeStepOps (:440-472);
//   * AssimilationCalcs() does the N x Nobs part of the analysis on the host (HA, P, the N solves, Z = HA'.MM, :478-575) and
//     returns Z; the O(M N^2) update X += A.Z/(N-1) and UpdateFromStateMatrix run on the device (csrc/rvn_enkf.cuh).
class CEnKFEnsemble : public CEnsemble {
public:
  CEnKFEnsemble(int num_members, const optStruct &Options);
  void SetEnKFMode(EnKF_mode mode) { _EnKF_mode = mode; }
  EnKF_mode GetEnKFMode() const { return _EnKF_mode; }
  void SetWindowSize(int nTimesteps) { _window_size = nTimesteps; }
  void AddObsPerturbation(sv_type type, int distrib, const double *distpars, adjustment adj);
  void AddAssimilationState(sv_type sv, int lay, int assim_groupID, bool is_streamflow);
  void SetLocalMembers(int member0, int nlocal) { _member0 = member0; _nLocal = nlocal; }
  void Initialize(const CModel *pModel, const optStruct &Options) override;
  void UpdateModel(CModel *pModel, optStruct &Options, int e) override;
  // state matrix description
  int  GetNumStateVars() const { return (int)_rows.size(); }
  const std::vector<rvn_enkf_row> &GetStateRows() const { return _rows; }
  const std::string &GetStateName(int i) const { return _state_names[i]; }
  // observations of the window: datapoint j = subbasin _obs_basin[j] at time index _obs_nn[j]
  int  GetNumObsDatapoints() const { return _nObsDatapoints; }
  const std::vector<int> &ObsBasins() const { return _obs_basin; }
  const std::vector<int> &ObsTimeIndices() const { return _obs_nn; }
  const std::vector<double> &ObsMatrix() const { return _obs_matrix; }       // [N][Nobs]
  const std::vector<double> &NoiseMatrix() const { return _noise_matrix; }   // [N][Nobs]
  // CloseTimeStepOps for one member: simulated HYDROGRAPH values of the window from the recorded outflows
  // qout_rec[nsteps][nSB] (CSubBasin::GetOutflowRate after every step) and the outflows before the first step
  void HarvestOutputs(const double *qout_rec, const double *qout_initial, int nSB, double timestep, double *out_row) const;
  // Z [N][N] from the simulated outputs of ALL members out_all[N][Nobs]; false (Z = 0) when there are no observations
  bool AssimilationCalcs(const double *out_all, std::vector<double> &Z) const;
private:
  EnKF_mode _EnKF_mode;
  int _window_size, _nTimeSteps, _nObsDatapoints, _member0, _nLocal;
  std::vector<obs_perturb> _pObsPerturbations;
  std::vector<assim_state> _aAssim;
  std::vector<rvn_enkf_row> _rows;
  std::vector<std::string> _state_names;
  std::vector<int> _obs_basin, _obs_nn;
  std::vector<int> _series_basin;            // subbasin of every assimilated observation series
  std::vector<char> _series_valid;           // [series][window step] 1 = a (non-blank) observation exists
  std::vector<double> _obs_matrix, _noise_matrix;
};
// x = pinv(P) b for the symmetric positive semi-definite Nobs x Nobs matrix P, singular values below tol*max dropped
// (what the reference's SVD(P,b,x,size,tol) returns, src/Matrix.cpp:144-390)
void SolveSymmetricTruncated(const std::vector<double> &P, int n, double tol, const std::vector<double> &B, int nrhs, std::vector<double> &X);

// factory: the ensemble object the .rvi/.rve files describe (src/ParseInput.cpp:3826-3833), seeded
CEnsemble *CreateEnsemble(CModel *pModel, const optStruct &Options);
#endif
