// rvh_capi.cpp -- plain-C surface of the host layer (for ctypes-based tests/bench and foreign callers).
// It exposes: parse+initialise a model from Raven input files, obtain the flattened rvn_model_desc
// (to feed rvn_gpu_create or, in tests only, the CPU oracle), initial state, and name/meta lookups.
#include "rvh_model.h"
#include "rvh_process.h"
#include "rvh_solver.h"
#include "rvh_ensemble.h"

struct rvh_model {
  CModel *pModel;
  optStruct Options;
  CEnsemble *pEnsemble;
  rvh_model() : pModel(NULL), pEnsemble(NULL) {}
};
static std::string g_err;
#define RVH_TRY try {
#define RVH_CATCH(ret) } catch (const std::exception &e) { g_err = e.what(); return ret; }

extern "C" {
const char *rvh_last_error() { return g_err.c_str(); }
#ifndef RVH_BUILD_DIGEST
#define RVH_BUILD_DIGEST "unstamped"
#endif
const char *rvh_build_digest() { return RVH_BUILD_DIGEST; }   // sha256 of sources + flags, stamped by build.py

// base: path prefix of the input set ("dir/name" -> dir/name.rvi, .rvh, .rvp, .rvt, .rvc, .rve), like `Raven.exe base`
// duration_override > 0 shortens the run (the BMI config 'duration:' does the same, src/Raven_BMI.cpp:341-345)
rvh_model *rvh_model_load(const char *base, double duration_override) {
  RVH_TRY
  rvh_model *h = new rvh_model();
  std::string b(base);
  h->Options.rvi_filename = b + ".rvi"; h->Options.rvh_filename = b + ".rvh"; h->Options.rvp_filename = b + ".rvp";
  h->Options.rvt_filename = b + ".rvt"; h->Options.rvc_filename = b + ".rvc"; h->Options.rve_filename = b + ".rve";
  try {
    ParseInputFiles(h->pModel, h->Options);
    if (duration_override > 0) {
      if (duration_override > h->Options.duration) throw RavenError("rvh_model_load: duration override exceeds :Duration of the .rvi file");
      h->Options.duration = duration_override;
    }
    h->pModel->Initialize(h->Options);
    ParseInitialConditions(h->pModel, h->Options);
    h->pModel->ApplyInitialAssimilation(h->Options);
  } catch (...) { delete h->pModel; delete h; throw; }
  return h;
  RVH_CATCH(NULL)
}
void rvh_model_free(rvh_model *h) { if (h) { delete h->pEnsemble; delete h->pModel; delete h; } }
// ensemble hooks with the reference's call order (src/RavenMain.cpp:110-120): rvh_ensemble_begin = ParseInput's
// SetRandomSeed + CEnsemble::Initialize; rvh_ensemble_update_model(e) = CEnsemble::UpdateModel(pModel,Options,e).
// Members must be updated in order 0,1,2,... (one shared rand() stream).  Returns the number of sampled values written.
int rvh_ensemble_begin(rvh_model *h) {
  RVH_TRY
  delete h->pEnsemble;
  h->pEnsemble = CreateEnsemble(h->pModel, h->Options);
  h->pEnsemble->Initialize(h->pModel, h->Options);
  return h->pEnsemble->GetNumMembers();
  RVH_CATCH(-1)
}
// reparse_ic != 0: re-read the .rvc under the member's parameters, as CMonteCarloEnsemble / CDDSEnsemble::UpdateModel do after
// UpdateParameter (src/ModelEnsemble.cpp:259, src/DDS.cpp:197: stores above a sampled capacity are clipped); callers that only
// replay the rand() stream of members they do not own pass 0.
int rvh_ensemble_update_model_ex(rvh_model *h, int e, double *values, int nmax, int reparse_ic);
int rvh_ensemble_update_model(rvh_model *h, int e, double *values, int nmax) { return rvh_ensemble_update_model_ex(h, e, values, nmax, 1); }
int rvh_ensemble_update_model_ex(rvh_model *h, int e, double *values, int nmax, int reparse_ic) {
  RVH_TRY
  if (!h->pEnsemble) throw RavenError("rvh_ensemble_update_model: call rvh_ensemble_begin first");
  h->pEnsemble->UpdateModel(h->pModel, h->Options, e);
  if (reparse_ic && (dynamic_cast<const CMonteCarloEnsemble *>(h->pEnsemble) || dynamic_cast<const CDDSEnsemble *>(h->pEnsemble)))
    ParseInitialConditions(h->pModel, h->Options);
  const CMonteCarloEnsemble *mc = dynamic_cast<const CMonteCarloEnsemble *>(h->pEnsemble);
  const CDDSEnsemble *dds = dynamic_cast<const CDDSEnsemble *>(h->pEnsemble);
  int n = 0;
  if (mc && values) for (; n < (int)mc->LastSample().size() && n < nmax; n++) values[n] = mc->LastSample()[n];
  if (dds && values) for (; n < (int)dds->TestParams().size() && n < nmax; n++) values[n] = dds->TestParams()[n];
  return n;
  RVH_CATCH(-1)
}
// checkpoint of one member: the reference's solution.rvc from arrays in the layouts of rvn_gpu_download_state / _basin_state
int rvh_model_write_solution(rvh_model *h, const char *filename, double model_time, const double *state, const double *qin, const double *qlat,
                             const double *qout, const double *scal, int full_precision) {
  RVH_TRY
  h->pModel->WriteSolutionFile(filename, h->Options, model_time, state, qin, qlat, qout, scal, full_precision != 0);
  return 0;
  RVH_CATCH(-1)
}
// the same with the reservoir state [nRes][RVN_RES_STATE] of the member (:ResFlow / :ResStage lines)
int rvh_model_write_solution_res(rvh_model *h, const char *filename, double model_time, const double *state, const double *qin, const double *qlat,
                                 const double *qout, const double *scal, int full_precision, const double *res_state) {
  RVH_TRY
  h->pModel->WriteSolutionFile(filename, h->Options, model_time, state, qin, qlat, qout, scal, full_precision != 0, res_state);
  return 0;
  RVH_CATCH(-1)
}
// Hydrographs.csv of one member from the recorded hydrographs qout_rec[nsteps][nSB], the initial outflows q0 / q0last [nSB] and the recorded
// watershed-average precipitation precip[nsteps] (CModel::WriteMinorOutput, src/StandardOutput.cpp:790-836)
int rvh_model_write_hydrographs(rvh_model *h, const char *filename, const double *qout_rec, const double *q0, const double *q0last, const double *precip, int nsteps) {
  RVH_TRY
  h->pModel->WriteHydrographsFile(filename, h->Options, qout_rec, q0, q0last, precip, nsteps);
  return 0;
  RVH_CATCH(-1)
}
// ---- DDS (CDDSEnsemble): rvh_ensemble_begin + rvh_ensemble_update_model(e) perturb the best-so-far vector (the sampled values
// returned are the trial's parameters); after the trial's GPU run rvh_dds_finish_run(e, hydrographs) evaluates the objective
// function and updates the best solution.  qout_rec[nsteps][nSB] = CSubBasin::GetOutflowRate after every step of the single member.
double rvh_dds_finish_run(rvh_model *h, int e, const double *qout_rec, const double *qout_initial, int nsteps, double *Fbest) {
  RVH_TRY
  CDDSEnsemble *dds = dynamic_cast<CDDSEnsemble *>(h->pEnsemble);
  if (!dds) throw RavenError("rvh_dds_finish_run: the model is not an ENSEMBLE_DDS model or rvh_ensemble_begin was not called");
  const double F = dds->ObjectiveFromHydrograph(h->pModel, h->Options, qout_rec, qout_initial, nsteps, h->pModel->GetNumSubBasins());
  dds->FinishEnsembleRun(F, e);
  if (Fbest) *Fbest = dds->BestObjective();
  return F;
  RVH_CATCH(-1e300)
}
// device-side objective: the inputs of rvn_gpu_diagnostic (obs may be NULL to query its length), and the bookkeeping of a trial
// whose diagnostic value came back from the device (sign convention of CModel::GetObjFuncVal applied here)
int rvh_dds_objective_inputs(rvh_model *h, int *basin, int *type, int *nnstart, int *nnend, double *obs, int nmax) {
  RVH_TRY
  CDDSEnsemble *dds = dynamic_cast<CDDSEnsemble *>(h->pEnsemble);
  if (!dds) throw RavenError("rvh_dds_objective_inputs: not a DDS ensemble");
  std::vector<double> o;
  dds->ObjectiveInputs(h->pModel, h->Options, *basin, *type, *nnstart, *nnend, o);
  if (obs) for (int i = 0; i < (int)o.size() && i < nmax; i++) obs[i] = o[i];
  return (int)o.size();
  RVH_CATCH(-1)
}
double rvh_dds_finish_value(rvh_model *h, int e, double diagnostic, double *Fbest) {
  RVH_TRY
  CDDSEnsemble *dds = dynamic_cast<CDDSEnsemble *>(h->pEnsemble);
  if (!dds) throw RavenError("rvh_dds_finish_value: not a DDS ensemble");
  int basin, type, a, b; std::vector<double> o;
  dds->ObjectiveInputs(h->pModel, h->Options, basin, type, a, b, o);
  const double F = (type == 1) ? diagnostic : -1.0 * diagnostic;   // NSE / KGE are maximised: objval *= -1
  dds->FinishEnsembleRun(F, e);
  if (Fbest) *Fbest = dds->BestObjective();
  return F;
  RVH_CATCH(-1e300)
}
int rvh_dds_best(rvh_model *h, double *best, int nmax) {
  RVH_TRY
  CDDSEnsemble *dds = dynamic_cast<CDDSEnsemble *>(h->pEnsemble);
  if (!dds) throw RavenError("rvh_dds_best: not a DDS ensemble");
  int n = 0;
  for (; n < (int)dds->BestParams().size() && n < nmax; n++) best[n] = dds->BestParams()[n];
  return n;
  RVH_CATCH(-1)
}
// ---- EnKF (CEnKFEnsemble).  rvh_enkf_begin = SetRandomSeed + Initialize + UpdateModel(e) for e = 0 .. member0+nlocal-1: the
// observation noise and the forcing-perturbation samples of the local member block [member0, member0+nlocal) come out of the
// one shared rand() stream at the positions the reference's member loop would use.  The model must have been flattened for
// nlocal members; the samples land in the descriptor's pert_eps table (rvn_gpu_create / rvn_gpu_upload_perturbations).
// info[6] = {N members, M state rows, Nobs datapoints, nPert, window steps, EnKF mode}
static CEnKFEnsemble *enkf_of(rvh_model *h) {
  CEnKFEnsemble *en = dynamic_cast<CEnKFEnsemble *>(h->pEnsemble);
  if (!en) throw RavenError("rvh_enkf: the model is not an ENSEMBLE_ENKF model or rvh_enkf_begin was not called");
  return en;
}
int rvh_enkf_begin(rvh_model *h, int member0, int nlocal, int *info) {
  RVH_TRY
  if (h->Options.ensemble != ENSEMBLE_ENKF) throw RavenError("rvh_enkf_begin: :EnsembleMode is not ENSEMBLE_ENKF");
  delete h->pEnsemble;
  h->pEnsemble = CreateEnsemble(h->pModel, h->Options);
  CEnKFEnsemble *en = enkf_of(h);
  if (member0 < 0 || nlocal < 1 || member0 + nlocal > en->GetNumMembers()) throw RavenError("rvh_enkf_begin: bad member block");
  en->SetLocalMembers(member0, nlocal);
  en->Initialize(h->pModel, h->Options);
  for (int e = 0; e < member0 + nlocal; e++) en->UpdateModel(h->pModel, h->Options, e);
  info[0] = en->GetNumMembers(); info[1] = en->GetNumStateVars(); info[2] = en->GetNumObsDatapoints();
  info[3] = h->pModel->GetNumForcingPerturbations(); info[4] = (int)((h->Options.duration - TIME_CORRECTION) / h->Options.timestep) + 1;
  info[5] = (int)en->GetEnKFMode();
  return 0;
  RVH_CATCH(-1)
}
int rvh_enkf_rows(rvh_model *h, rvn_enkf_row *rows) {
  RVH_TRY
  const std::vector<rvn_enkf_row> &r = enkf_of(h)->GetStateRows();
  memcpy(rows, r.data(), r.size() * sizeof(rvn_enkf_row));
  return 0;
  RVH_CATCH(-1)
}
int rvh_enkf_row_name(rvh_model *h, int i, char *buf, int buflen) {
  RVH_TRY
  snprintf(buf, buflen, "%s", enkf_of(h)->GetStateName(i).c_str());
  return 0;
  RVH_CATCH(-1)
}
int rvh_enkf_observations(rvh_model *h, double *obs, double *noise, int32_t *basin, int32_t *nn) {
  RVH_TRY
  const CEnKFEnsemble *en = enkf_of(h);
  if (obs) memcpy(obs, en->ObsMatrix().data(), en->ObsMatrix().size() * 8);
  if (noise) memcpy(noise, en->NoiseMatrix().data(), en->NoiseMatrix().size() * 8);
  for (int j = 0; j < en->GetNumObsDatapoints(); j++) { if (basin) basin[j] = en->ObsBasins()[j]; if (nn) nn[j] = en->ObsTimeIndices()[j]; }
  return 0;
  RVH_CATCH(-1)
}
const double *rvh_enkf_member_eps(rvh_model *h, int member_local) { return h->pModel->MemberPerturbations(member_local); }
// qout_rec[nsteps][nmembers][nSB] (layout of rvn_gpu_download_hydrographs), qout_initial[nSB]; out[nmembers][Nobs]
int rvh_enkf_harvest_outputs(rvh_model *h, const double *qout_rec, const double *qout_initial, int nsteps, int nmembers, double *out) {
  RVH_TRY
  const CEnKFEnsemble *en = enkf_of(h);
  const int nSB = h->pModel->GetNumSubBasins(), Nobs = en->GetNumObsDatapoints();
  std::vector<double> one((size_t)nsteps * nSB);
  for (int e = 0; e < nmembers; e++) {
    for (int n = 0; n < nsteps; n++) memcpy(&one[(size_t)n * nSB], qout_rec + ((size_t)n * nmembers + e) * nSB, (size_t)nSB * 8);
    en->HarvestOutputs(one.data(), qout_initial, nSB, h->Options.timestep, out + (size_t)e * Nobs);
  }
  return 0;
  RVH_CATCH(-1)
}
// out_all[N][Nobs] -> Z[N][N]; returns 1 when an analysis is due, 0 when there are no observations (Z = 0)
int rvh_enkf_compute_Z(rvh_model *h, const double *out_all, double *Z) {
  RVH_TRY
  const CEnKFEnsemble *en = enkf_of(h);
  std::vector<double> z;
  const bool due = en->AssimilationCalcs(out_all, z);
  memcpy(Z, z.data(), z.size() * 8);
  return due ? 1 : 0;
  RVH_CATCH(-1)
}
const rvn_model_desc *rvh_model_flatten(rvh_model *h, int nMembers) {
  RVH_TRY
  return h->pModel->Flatten(h->Options, nMembers);
  RVH_CATCH(NULL)
}
// sizes a caller needs to allocate buffers for a flattened model: {nSteps,max_nqin,max_nqlat,nCore,nConvProc,ptab_stride,nMembers,nLU}
int rvh_desc_info(const rvn_model_desc *d, int *out8) {
  int ncore = 0;
  for (int i = 0; i < d->NS; i++) if (d->sv_type[i] != RVN_SV_CONV_STOR) ncore++;
  out8[0] = d->nSteps; out8[1] = d->max_nqin; out8[2] = d->max_nqlat; out8[3] = ncore; out8[4] = d->nConvProc; out8[5] = d->ptab_stride;
  out8[6] = d->nMembers; out8[7] = d->nLU;
  return 0;
}
// reservoirs of the flattened model: count, and the initial state [nRes][RVN_RES_STATE] + basin index of each
int rvh_desc_num_reservoirs(const rvn_model_desc *d) { return d->nRes; }
int rvh_desc_reservoir_state(const rvn_model_desc *d, double *res_state, int32_t *res_basin) {
  for (int i = 0; i < d->nRes * RVN_RES_STATE; i++) res_state[i] = d->res_state0[i];
  if (res_basin) for (int r = 0; r < d->nRes; r++) res_basin[r] = d->res_basin[r];
  return 0;
}
const double *rvh_desc_gauge_series(const rvn_model_desc *d) { return d->gauge_series; }
int rvh_desc_num_gauges(const rvn_model_desc *d) { return d->nGauges; }
int rvh_desc_hru_lu(const rvn_model_desc *d, int32_t *lu) { for (int k = 0; k < d->nHRU; k++) lu[k] = d->hru_lu[k]; return 0; }
// gridded forcing ingestion (CModel::SetCellForcing): takes effect at the next rvh_model_flatten
int rvh_model_set_cell_forcing(rvh_model *h, int nCells, int nSteps, const double *cell_elev, const double *series, const int32_t *wt_ptr,
                               const int32_t *wt_idx, const double *wt_val) {
  RVH_TRY
  h->pModel->SetCellForcing(nCells, nSteps, cell_elev, series, wt_ptr, wt_idx, wt_val);
  return 0;
  RVH_CATCH(-1)
}
// per-connection cumulative flux arrays (reference CModel::IncrementBalance): off by default, takes effect at the next rvh_model_flatten
int rvh_model_set_flux_balance(rvh_model *h, int on) {
  RVH_TRY
  h->Options.keep_flux_balance = (on != 0);   // the model reads its options through a pointer to this struct
  return 0;
  RVH_CATCH(-1)
}
int rvh_desc_num_connections(const rvn_model_desc *d) { return d->nConn; }
int rvh_desc_connections(const rvn_model_desc *d, int32_t *from, int32_t *to) { for (int q = 0; q < d->nConn; q++) { from[q] = d->conn_from[q]; to[q] = d->conn_to[q]; } return 0; }
const int32_t *rvh_desc_conv_n(const rvn_model_desc *d) { return d->conv_n; }
int rvh_desc_sv_table(const rvn_model_desc *d, int32_t *type, int32_t *layer) {
  for (int i = 0; i < d->NS; i++) { type[i] = d->sv_type[i]; layer[i] = d->sv_layer[i]; }
  return 0;
}
int rvh_desc_nseg(const rvn_model_desc *d, int32_t *nseg) { for (int p = 0; p < d->nSB; p++) nseg[p] = d->sb_nseg[p]; return 0; }
int rvh_desc_order(const rvn_model_desc *d, int32_t *order) { for (int p = 0; p < d->nSB; p++) order[p] = d->sb_order[p]; return 0; }
int rvh_desc_hru_area(const rvn_model_desc *d, double *area) { for (int k = 0; k < d->nHRU; k++) area[k] = d->hru_area[k]; return 0; }
// resolved class parameter of HRU k as the hot path will see it; kind = "G" | "LU" | "VEG" | "SOIL<m>"
double rvh_model_param_value(rvh_model *h, const char *kind, int k, const char *name) {
  const CModel *M = h->pModel;
  std::string kd(kind), nm(name);
  if (kd == "G") { int f = GlobalFieldFromName(nm); return f < 0 ? NOT_SPECIFIED : M->globals.v[f]; }
  const CHydroUnit *u = M->GetHydroUnit(k);
  if (kd == "LU") { int f = LandUseFieldFromName(nm); return f < 0 ? NOT_SPECIFIED : M->lu_classes[u->lu_class].S.v[f]; }
  if (kd == "VEG") {
    for (int f = 0; f < RVN_V_COUNT; f++) if (nm == VegFieldName(f)) {
      if (f == RVN_V_VAR_RAIN_ICEPT_PCT) f = RVN_V_RAIN_ICEPT_PCT;
      if (f == RVN_V_VAR_SNOW_ICEPT_PCT) f = RVN_V_SNOW_ICEPT_PCT;
      return M->veg_classes[u->veg_class].V.v[f];
    }
    return NOT_SPECIFIED;
  }
  if (kd.substr(0, 4) == "SOIL") {
    int m = atoi(kd.c_str() + 4);
    int f = SoilFieldFromName(nm);
    const CSoilProfile &sp = M->soil_profiles[u->soil_profile];
    if (f < 0 || m >= sp.nHorizons) return NOT_SPECIFIED;
    return M->soil_classes[sp.soil_class[m]].S.v[f];
  }
  return NOT_SPECIFIED;
}
int rvh_model_dims(rvh_model *h, int *nHRU, int *NS, int *nSB, int *nProc, int *nGauges) {
  *nHRU = h->pModel->GetNumHRUs(); *NS = h->pModel->GetNumStateVars(); *nSB = h->pModel->GetNumSubBasins();
  *nProc = h->pModel->GetNumProcesses(); *nGauges = h->pModel->GetNumGauges();
  return 0;
}
int rvh_model_initial_state(rvh_model *h, double *out) {
  memcpy(out, h->pModel->initial_state.data(), h->pModel->initial_state.size() * sizeof(double));
  return 0;
}
int rvh_model_basin_state(rvh_model *h, double *qin, double *qlat, double *qout, double *scal) {
  RVH_TRY
  std::vector<double> a, b, c, d;
  h->pModel->PackBasinState(a, b, c, d);
  memcpy(qin, a.data(), a.size() * 8); memcpy(qlat, b.data(), b.size() * 8); memcpy(qout, c.data(), c.size() * 8); memcpy(scal, d.data(), d.size() * 8);
  return 0;
  RVH_CATCH(-1)
}
int rvh_model_sv_name(rvh_model *h, int i, char *buf, int buflen) {
  std::string n = h->pModel->GetStateVarName(i);
  snprintf(buf, buflen, "%s", n.c_str());
  return 0;
}
int rvh_model_process_info(rvh_model *h, int j, char *name, int buflen, int *nconn, int *from8, int *to8) {
  const CHydroProcessABC *p = h->pModel->GetProcess(j);
  snprintf(name, buflen, "%s", p->GetProcessName().c_str());
  *nconn = p->GetNumConnections();
  for (int q = 0; q < 8; q++) { from8[q] = (q < *nconn) ? p->GetFromIndices()[q] : -1; to8[q] = (q < *nconn) ? p->GetToIndices()[q] : -1; }
  return 0;
}
// ensemble support: overwrite a class parameter (CModel::UpdateParameter) and refresh one member's flattened row
int rvh_model_update_parameter(rvh_model *h, int class_type, const char *pname, const char *cname, double value) {
  RVH_TRY
  h->pModel->UpdateParameter(class_type, pname, cname, value);
  return 0;
  RVH_CATCH(-1)
}
int rvh_model_flatten_member(rvh_model *h, int member, const double **ptab, const int32_t **conv_n, const double **conv_uh) {
  RVH_TRY
  h->pModel->FlattenMemberParams(member);
  *ptab = h->pModel->MemberParams(member); *conv_n = h->pModel->MemberConvN(member); *conv_uh = h->pModel->MemberConvUH(member);
  return 0;
  RVH_CATCH(-1)
}
int rvh_model_num_param_dists(rvh_model *h) { return (int)h->pModel->param_dists.size(); }
int rvh_model_param_dist(rvh_model *h, int i, char *pname, char *cname, int buflen, int *ctype, int *dist, double *par3) {
  const param_dist &d = h->pModel->param_dists[i];
  snprintf(pname, buflen, "%s", d.param_name.c_str()); snprintf(cname, buflen, "%s", d.class_group.c_str());
  *ctype = d.param_class; *dist = d.distribution; par3[0] = d.distpar[0]; par3[1] = d.distpar[1]; par3[2] = d.distpar[2];
  return 0;
}
int rvh_model_options(rvh_model *h, double *timestep, double *duration, unsigned *seed, int *ensemble, int *num_members) {
  *timestep = h->Options.timestep; *duration = h->Options.duration; *seed = h->Options.random_seed; *ensemble = (int)h->Options.ensemble;
  *num_members = h->pModel->num_members;
  return 0;
}
}  // extern "C"
