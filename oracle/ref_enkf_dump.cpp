// oracle/ref_enkf_dump.cpp -- TEST INFRASTRUCTURE ONLY (never on the product path).
//
// EnKF parity driver for the UNMODIFIED reference Raven, linked against oracle/_ref/libravenref.so (see build_ref.sh).
// It replays the reference's ensemble loop (reference src/RavenMain.cpp:93-196) for an ENSEMBLE_ENKF model: every member is
// run over the assimilation window with its forcing perturbations, then the analysis of CEnKFEnsemble (src/EnKF.cpp:478-596)
// is executed by the reference's own AssimilationCalcs().  Because EnKFOutput.csv carries 6 digits, the in-memory FP64
// matrices are dumped instead:
//     <out>/enkf_meta.txt      N (members), M (state rows), Nobs, nHRU, NS, NB, state row names
//     <out>/enkf_obs.f64       [N][Nobs]  _obs_matrix   (perturbed observations)
//     <out>/enkf_noise.f64     [N][Nobs]  _noise_matrix
//     <out>/enkf_out.f64       [N][Nobs]  _output_matrix (simulated observations)
//     <out>/enkf_Xpre.f64      [N][M]     _state_matrix before the analysis
//     <out>/enkf_Xpost.f64     [N][M]     _state_matrix after the analysis
//     <out>/enkf_state.f64     [N][nHRU][NS]  every member's full HRU state at the end of the window
//     <out>/enkf_qout.f64      [N][nsteps][NB] outflow of every basin after every step
//     <out>/enkf_eps.f64       [N][nsteps][nPert] perturbation samples used (eps[0] of each perturbation at each step)
// The solution-file rewriting at the end of FinishEnsembleRun (text round trip of every member) is not replayed.
//
// usage: ref_enkf_dump <rvi-base> <outdir/>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <fstream>
#include <iostream>
#include <sstream>
#include <map>
#include <set>
#include <algorithm>
#define private public
#define protected public
#include "RavenInclude.h"
#include "RavenMain.h"
#include "Model.h"
#include "EnKF.h"
#undef private
#undef protected

void ParseManagementFile(CModel *&pModel, optStruct &Options);
void MassEnergyBalance(CModel *pModel, const optStruct &Options, const time_struct &tt);

static void wr(FILE *f, const double *p, size_t n) { if (n && fwrite(p, 8, n, f) != n) { perror("fwrite"); exit(3); } }
static FILE *op(const std::string &n) { FILE *f = fopen(n.c_str(), "wb"); if (!f) { perror(n.c_str()); exit(3); } return f; }

int main(int argc, char **argv)
{
  if (argc < 3) { fprintf(stderr, "usage: ref_enkf_dump <base> <outdir/>\n"); return 2; }
  std::string outdir = argv[2];
  CModel *pModel = NULL;
  optStruct Options;
  time_struct tt;
  Options.version = "refdump";
  std::vector<char *> av;
  av.push_back(argv[0]); av.push_back(argv[1]);
  std::string o = "-o"; av.push_back((char *)o.c_str()); av.push_back(argv[2]);
  if (!ProcessExecutableArguments((int)av.size(), av.data(), Options)) return 1;
  PrepareOutputdirectory(Options);
  for (int i = 0; i < 10; i++) g_debug_vars[i] = 0;
  { std::ofstream W((Options.main_output_dir + "Raven_errors.txt").c_str()); W.close(); }
  if (!ParseInputFiles(pModel, Options)) { fprintf(stderr, "parse failed\n"); return 1; }
  CheckForErrorWarnings(true, pModel);
  pModel->Initialize(Options);
  ParseManagementFile(pModel, Options);
  pModel->InitializePostRVM(Options);
  ParseInitialConditions(pModel, Options);
  pModel->CalculateInitialWaterStorage(Options);
  pModel->GetEnsemble()->Initialize(pModel, Options);
  CheckForErrorWarnings(false, pModel);
  if (pModel->GetEnsemble()->GetType() != ENSEMBLE_ENKF) { fprintf(stderr, "not an ENSEMBLE_ENKF model\n"); return 1; }
  CEnKFEnsemble *pE = (CEnKFEnsemble *)pModel->GetEnsemble();

  const int nH = pModel->GetNumHRUs(), NS = pModel->GetNumStateVars(), NB = pModel->GetNumSubBasins();
  const int N = pE->GetNumMembers(), M = pE->_nStateVars, Nobs = pE->_nObsDatapoints, nPert = pModel->_nPerturbations;
  FILE *fstate = op(outdir + "enkf_state.f64"), *fq = op(outdir + "enkf_qout.f64"), *feps = op(outdir + "enkf_eps.f64");
  int nsteps = 0;

  for (int e = 0; e < N; e++) {
    pE->UpdateModel(pModel, Options, e);
    PrepareOutputdirectory(Options);
    pModel->WriteOutputFileHeaders(Options);
    double t_start = pE->GetStartTime(e);
    JulianConvert(t_start, Options.julian_start_day, Options.julian_start_year, Options.calendar, tt);
    pModel->RecalculateHRUDerivedParams(Options, tt);
    pModel->UpdateHRUForcingFunctions(Options, tt);
    pModel->UpdateDiagnostics(Options, tt);
    pModel->InitializePostRVC(Options);
    pModel->WriteMinorOutput(Options, tt);
    int step = 0;
    for (double t = t_start; t < Options.duration - TIME_CORRECTION; t += Options.timestep) {
      pModel->UpdateTransientParams(Options, tt);
      pModel->ApplyStateOverrrides(Options, tt);
      pModel->RecalculateHRUDerivedParams(Options, tt);
      pE->StartTimeStepOps(pModel, Options, tt, e);
      for (int i = 0; i < nPert; i++) wr(feps, &pModel->_pPerturbations[i]->eps[0], 1);
      pModel->UpdateHRUForcingFunctions(Options, tt);
      pModel->PrepareAssimilation(Options, tt);
      MassEnergyBalance(pModel, Options, tt);
      pModel->IncrementCumulInput(Options, tt);
      pModel->IncrementCumOutflow(Options, tt);
      JulianConvert(t + Options.timestep, Options.julian_start_day, Options.julian_start_year, Options.calendar, tt);
      pModel->WriteMinorOutput(Options, tt);
      pModel->UpdateDiagnostics(Options, tt);
      pE->CloseTimeStepOps(pModel, Options, tt, e);
      for (int p = 0; p < NB; p++) { double q = pModel->GetSubBasin(p)->GetOutflowRate(); wr(fq, &q, 1); }
      step++;
    }
    nsteps = step;
    pModel->UpdateDiagnostics(Options, tt);
    pModel->CloseOutputStreams();
    for (int k = 0; k < nH; k++) for (int i = 0; i < NS; i++) { double v = pModel->GetHydroUnit(k)->GetStateVarValue(i); wr(fstate, &v, 1); }
    // FinishEnsembleRun (src/EnKF.cpp:813-...) without the solution-file rewriting: harvest the state row; after the last
    // member dump, run the reference's analysis and dump again
    pE->AddToStateMatrix(pModel, Options, e);
    if (e == N - 1) {
      FILE *f = op(outdir + "enkf_Xpre.f64");
      for (int ee = 0; ee < N; ee++) wr(f, pE->_state_matrix[ee], M);
      fclose(f);
      f = op(outdir + "enkf_out.f64");
      for (int ee = 0; ee < N; ee++) wr(f, pE->_output_matrix[ee], Nobs);
      fclose(f);
      pE->AssimilationCalcs();
      f = op(outdir + "enkf_Xpost.f64");
      for (int ee = 0; ee < N; ee++) wr(f, pE->_state_matrix[ee], M);
      fclose(f);
      f = op(outdir + "enkf_obs.f64");
      for (int ee = 0; ee < N; ee++) wr(f, pE->_obs_matrix[ee], Nobs);
      fclose(f);
      f = op(outdir + "enkf_noise.f64");
      for (int ee = 0; ee < N; ee++) wr(f, pE->_noise_matrix[ee], Nobs);
      fclose(f);
    }
    pModel->RebootTimeVariables(Options);
  }
  fclose(fstate); fclose(fq); fclose(feps);
  {
    std::ofstream Mf((outdir + "enkf_meta.txt").c_str());
    Mf << "N " << N << "\nM " << M << "\nNobs " << Nobs << "\nnHRU " << nH << "\nNS " << NS << "\nNB " << NB << "\nnsteps " << nsteps
       << "\nnPert " << nPert << "\n";
    for (int i = 0; i < M; i++) Mf << "ROW " << i << " " << pE->_state_names[i] << "\n";
  }
  return 0;
}
