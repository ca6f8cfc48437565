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

// factory: the ensemble object the .rvi/.rve files describe (src/ParseInput.cpp:3826-3833), seeded
CEnsemble *CreateEnsemble(CModel *pModel, const optStruct &Options);
#endif
