// rvh_solver.cpp -- the reference's step entry points, re-authored as thin calls into the GPU engine.
//
//   void MassEnergyBalance(CModel*, const optStruct&, const time_struct&)       reference src/Solvers.cpp:19-21
//   void CModel::UpdateHRUForcingFunctions(const optStruct&, const time_struct&) reference src/UpdateForcings.cpp:30
//
// In the reference these two are separate host sweeps over all HRUs; here the forcing derivation is
// fused into the device sweep, so UpdateHRUForcingFunctions only records the step index and
// MassEnergyBalance launches forcing update + process sweep + routing + balance reductions for that
// step through rvn_gpu_run().  RunSimulation() is the batched form used by the CLI/bench: all steps of
// all ensemble members without returning to the host in between.
#include "rvh_solver.h"

CGPUSimulation::CGPUSimulation(CModel *pModel, const optStruct &Options, int nMembers, int device)
    : _pModel(pModel), _gpu(NULL), _nMembers(nMembers), _step(0) {
  const rvn_model_desc *desc = pModel->Flatten(Options, nMembers);
  _nSteps = desc->nSteps;
  int rc = rvn_gpu_create(desc, device, &_gpu);
  if (rc != RVN_OK) ExitGracefully(std::string("rvn_gpu_create failed: ") + rvn_gpu_last_error());
  // initial HRU state: every member starts from the .rvc state (reference re-reads the .rvc per member,
  // src/ModelEnsemble.cpp:255-259)
  const size_t n1 = (size_t)pModel->GetNumHRUs() * pModel->GetNumStateVars();
  ExitGracefullyIf(pModel->initial_state.size() != n1, "CGPUSimulation: initial conditions were not parsed (call ParseInitialConditions)");
  std::vector<double> qin, qlat, qout, scal;
  pModel->PackBasinState(qin, qlat, qout, scal);
  for (int e = 0; e < nMembers; e++) {
    Check(rvn_gpu_upload_state(_gpu, pModel->initial_state.data(), e, 1));
    Check(rvn_gpu_upload_basin_state(_gpu, qin.data(), qlat.data(), qout.data(), scal.data(), e, 1));
  }
}
CGPUSimulation::~CGPUSimulation() { if (_gpu) rvn_gpu_destroy(_gpu); }
void CGPUSimulation::Check(int rc) const {
  if (rc != RVN_OK) ExitGracefully(std::string("GPU engine error: ") + rvn_gpu_last_error());
}
void CGPUSimulation::UploadMemberParams(int member) {
  _pModel->FlattenMemberParams(member);
  Check(rvn_gpu_upload_params(_gpu, _pModel->MemberParams(member), _pModel->MemberConvN(member), _pModel->MemberConvUH(member), member, 1));
}
void CGPUSimulation::Step() {
  ExitGracefullyIf(_step >= _nSteps, "CGPUSimulation::Step: simulation already reached its duration");
  Check(rvn_gpu_run(_gpu, _step, 1));
  _step++;
}
void CGPUSimulation::Run(int nsteps) {
  if (nsteps < 0 || _step + nsteps > _nSteps) nsteps = _nSteps - _step;
  Check(rvn_gpu_run(_gpu, _step, nsteps));
  _step += nsteps;
}
void CGPUSimulation::Sync() { Check(rvn_gpu_sync(_gpu)); }

static CGPUSimulation *g_sim = NULL;   // one model per process, like the reference's function-local statics (src/Solvers.cpp:42-62)
static CModel *g_sim_model = NULL;
CGPUSimulation *GetGPUSimulation(CModel *pModel, const optStruct &Options) {
  if (g_sim == NULL || g_sim_model != pModel) {
    delete g_sim;
    g_sim = new CGPUSimulation(pModel, Options, 1, 0);
    g_sim_model = pModel;
  }
  return g_sim;
}
void ReleaseGPUSimulation() { delete g_sim; g_sim = NULL; g_sim_model = NULL; }

void UpdateHRUForcingFunctions(CModel *pModel, const optStruct &Options, const time_struct &tt) {
  // forcing derivation is fused into the device sweep launched by MassEnergyBalance for the same step
  (void)pModel; (void)Options; (void)tt;
}
void MassEnergyBalance(CModel *pModel, const optStruct &Options, const time_struct &tt) {
  CGPUSimulation *sim = GetGPUSimulation(pModel, Options);
  int nn = (int)((tt.model_time + TIME_CORRECTION) / Options.timestep);
  ExitGracefullyIf(nn != sim->CurrentStep(), "MassEnergyBalance: steps must be taken in order (the device state is at step " +
                                                 std::to_string(sim->CurrentStep()) + ", asked for " + std::to_string(nn) + ")");
  sim->Step();
}
