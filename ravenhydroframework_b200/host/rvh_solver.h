// rvh_solver.h -- step entry points with the reference's signatures + the batched GPU simulation object.
#ifndef RVH_SOLVER_H
#define RVH_SOLVER_H
#include "rvh_model.h"

class CGPUSimulation {
public:
  CGPUSimulation(CModel *pModel, const optStruct &Options, int nMembers, int device);
  ~CGPUSimulation();
  void UploadMemberParams(int member);   // after CModel::UpdateParameter for that member
  void Step();                           // one model step, all members
  void Run(int nsteps);                  // nsteps < 0: to the end of the simulation
  void Sync();
  int  CurrentStep() const { return _step; }
  int  NumSteps() const { return _nSteps; }
  int  NumMembers() const { return _nMembers; }
  rvn_gpu_model *engine() const { return _gpu; }
private:
  void Check(int rc) const;
  CModel *_pModel;
  rvn_gpu_model *_gpu;
  int _nMembers, _nSteps, _step;
};

// reference-named entry points (src/RavenMain.h, src/Model.h:536)
void MassEnergyBalance(CModel *pModel, const optStruct &Options, const time_struct &tt);
void UpdateHRUForcingFunctions(CModel *pModel, const optStruct &Options, const time_struct &tt);
CGPUSimulation *GetGPUSimulation(CModel *pModel, const optStruct &Options);
void ReleaseGPUSimulation();
#endif
