// oracle/ref_dump.cpp -- TEST INFRASTRUCTURE ONLY (never on the product path).
//
// Parity driver for the UNMODIFIED reference Raven, linked against oracle/_ref/libravenref.so
// (the reference compiled with -D_BMI_LIBRARY_ by oracle/build_ref.sh; headers are included from
// where they lie in the read-only reference tree, nothing is copied).
//
// It replays the reference's own CLI time loop (reference src/RavenMain.cpp:93-196) and, because
// Raven's text outputs carry only 5-6 significant digits, dumps the in-memory
// FP64 values after every step:
//     <out>/state.f64   [nsteps+1][nHRU][NS]   CHydroUnit::GetStateVarValue
//     <out>/forc.f64    [nsteps+1][nHRU][43]   force_struct (all doubles), as used for the step
//     <out>/basin.f64   [nsteps+1][NB][6]      Qout, Qlocal, channel stor, rivulet stor, Qin[0], Qlat[0]
//     <out>/wshed.f64   [nsteps+1][3]          _CumulInput, _CumulOutput, GetAveragePrecip
//     <out>/cumflux.f64 [nHRU][NS][2]          CModel::GetCumulativeFlux(k, i, to / from) at the end of the run
//     <out>/res.f64     [nsteps+1][NB][4]      reservoir stage, outflow, losses of the step [m3], storage [m3] (zeros for basins without one)
//     <out>/meta.txt    integer tables: SV catalogue, process connections, routing order, UH arrays
// Row 0 is the initial condition (after the t=0 forcing update), row n the state after step n.
//
// usage: ref_dump <rvi-base> <outdir/> <nsteps|-1> [extra Raven CLI args...]
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <fstream>
#include <iostream>
#include <sstream>
#define private public
#define protected public
#include "RavenInclude.h"
#include "RavenMain.h"
#include "Model.h"
#include "HydroProcessABC.h"
#undef private
#undef protected

void ParseManagementFile(CModel *&pModel, optStruct &Options);

static void wr(FILE *f, const double *p, size_t n) { if (fwrite(p, 8, n, f) != n) { perror("fwrite"); exit(3); } }

int main(int argc, char **argv)
{
  if (argc < 4) { fprintf(stderr, "usage: ref_dump <base> <outdir/> <nsteps|-1> [raven args]\n"); return 2; }
  std::string outdir = argv[2];
  int max_steps = atoi(argv[3]);

  CModel *pModel = NULL;
  optStruct Options;
  time_struct tt;
  Options.version = "refdump";

  std::vector<char *> av;
  av.push_back(argv[0]); av.push_back(argv[1]);
  std::string o = "-o"; av.push_back((char *)o.c_str()); av.push_back(argv[2]);
  for (int i = 4; i < argc; i++) av.push_back(argv[i]);
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

  const int nH = pModel->GetNumHRUs(), NS = pModel->GetNumStateVars(), NB = pModel->GetNumSubBasins();
  const int NF = (int)(sizeof(force_struct) / sizeof(double));
  int nsteps_total = (int)((Options.duration - TIME_CORRECTION) / Options.timestep) + 1;
  (void)nsteps_total;

  // REF_DUMP_MEMBER=m: dump ensemble member m.  Members before it are "updated" in order first so that the shared
  // rand() stream is where the reference's own member loop (src/RavenMain.cpp:117-120) would have left it.
  int e = getenv("REF_DUMP_MEMBER") ? atoi(getenv("REF_DUMP_MEMBER")) : 0;
  for (int ee = 0; ee < e; ee++) pModel->GetEnsemble()->UpdateModel(pModel, Options, ee);
  pModel->GetEnsemble()->UpdateModel(pModel, Options, e);
  PrepareOutputdirectory(Options);
  pModel->WriteOutputFileHeaders(Options);
  double t_start = pModel->GetEnsemble()->GetStartTime(e);
  JulianConvert(t_start, Options.julian_start_day, Options.julian_start_year, Options.calendar, tt);
  pModel->RecalculateHRUDerivedParams(Options, tt);
  pModel->UpdateHRUForcingFunctions(Options, tt);
  pModel->UpdateDiagnostics(Options, tt);
  pModel->InitializePostRVC(Options);
  pModel->WriteMinorOutput(Options, tt);

  // ---- meta -------------------------------------------------------------------------------
  {
    std::ofstream M((outdir + "meta.txt").c_str());
    M.precision(17);
    M << "nHRU " << nH << "\nNS " << NS << "\nNB " << NB << "\nNF " << NF << "\ndt " << Options.timestep << "\n";
    for (int i = 0; i < NS; i++)
      M << "SV " << i << " " << (int)pModel->GetStateVarType(i) << " " << pModel->GetStateVarLayer(i) << " "
        << CStateVariable::GetStateVarLongName(pModel->GetStateVarType(i), pModel->GetStateVarLayer(i), pModel->GetTransportModel()) << "\n";
    for (int j = 0; j < pModel->GetNumProcesses(); j++) {
      CHydroProcessABC *pr = pModel->GetProcess(j);
      int nc = pr->GetNumConnections();
      M << "PROC " << j << " " << (int)pr->GetProcessType() << " " << nc;
      // print at most the first 8 connections in full (convolutions have 101)
      for (int q = 0; q < nc && q < 8; q++) M << " " << pr->GetFromIndices()[q] << ">" << pr->GetToIndices()[q];
      M << "\n";
    }
    for (int pp = 0; pp < NB; pp++) M << "ORDER " << pp << " " << pModel->GetOrderedSubBasinIndex(pp) << "\n";
    for (int p = 0; p < NB; p++) {
      const CSubBasin *b = pModel->GetSubBasin(p);
      M << "BASIN " << p << " id " << b->GetID() << " down " << pModel->GetDownstreamBasin(p)
        << " nseg " << b->GetNumSegments() << " nQin " << b->GetInflowHistorySize() << " nQlat " << b->GetLatHistorySize()
        << " area " << b->GetBasinArea() << " reach " << b->GetReachLength() << " Qref " << b->GetReferenceFlow() << "\n";
      if (p < 64) {
        M << "UH " << p;
        for (int n = 0; n < b->GetLatHistorySize(); n++) M << " " << b->GetUnitHydrograph()[n];
        M << "\nRH " << p;
        for (int n = 0; n < b->GetInflowHistorySize(); n++) M << " " << b->GetRoutingHydrograph()[n];
        M << "\n";
      }
    }
    { // parameter values as the reference resolved them (defaults + autogeneration), by include/rvn_gpu.h field name
      const global_struct *G = pModel->GetGlobalParams()->GetParams();
#define GP(n, f) M << "PARAM G 0 " << n << " " << G->f << "\n";
      GP("SNOW_SWI", snow_SWI) GP("SNOW_SWI_MIN", snow_SWI_min) GP("SNOW_SWI_MAX", snow_SWI_max) GP("SWI_REDUCT_COEFF", SWI_reduct_coeff)
      GP("ADIABATIC_LAPSE", adiabatic_lapse) GP("PRECIP_LAPSE", precip_lapse) GP("RAINSNOW_TEMP", rainsnow_temp) GP("RAINSNOW_DELTA", rainsnow_delta)
      GP("AVG_ANNUAL_SNOW", avg_annual_snow) GP("SNOW_TEMPERATURE", snow_temperature) GP("HBVEC_LAPSE_UPPER", HBVEC_lapse_upper)
      GP("HBVEC_LAPSE_RATE", HBVEC_lapse_rate) GP("HBVEC_LAPSE_ELEV", HBVEC_lapse_elev) GP("AIRSNOW_COEFF", airsnow_coeff)
      GP("AVG_ANNUAL_RUNOFF", avg_annual_runoff) GP("MAX_SWE_SURFACE", max_SWE_surface) GP("TOC_MULTIPLIER", TOC_multiplier)
      GP("TIME_TO_PEAK_MULTIPLIER", TIME_TO_PEAK_multiplier) GP("GAMMA_SHAPE_MULTIPLIER", GAMMA_SHAPE_multiplier) GP("GAMMA_SCALE_MULTIPLIER", GAMMA_SCALE_multiplier)
      for (int k = 0; k < nH && k < 64; k++) {
        const CHydroUnit *h = pModel->GetHydroUnit(k);
        const surface_struct *S = h->GetSurfaceProps();
#define LP(n, f) M << "PARAM LU " << k << " " << n << " " << S->f << "\n";
        LP("IMPERMEABLE_FRAC", impermeable_frac) LP("FOREST_COVERAGE", forest_coverage) LP("FOREST_SPARSENESS", forest_sparseness)
        LP("MELT_FACTOR", melt_factor) LP("MIN_MELT_FACTOR", min_melt_factor) LP("MAX_MELT_FACTOR", max_melt_factor) LP("DD_MELT_TEMP", DD_melt_temp)
        LP("DD_AGGRADATION", DD_aggradation) LP("REFREEZE_FACTOR", refreeze_factor) LP("DD_REFREEZE_TEMP", DD_refreeze_temp) LP("REFREEZE_EXP", refreeze_exp)
        LP("HMETS_RUNOFF_COEFF", HMETS_runoff_coeff) LP("GAMMA_SHAPE", gamma_shape) LP("GAMMA_SCALE", gamma_scale) LP("GAMMA_SHAPE2", gamma_shape2)
        LP("GAMMA_SCALE2", gamma_scale2) LP("GR4J_X4", GR4J_x4) LP("HBV_MELT_FOR_CORR", HBV_melt_for_corr) LP("HBV_MELT_ASP_CORR", HBV_melt_asp_corr)
        LP("HBV_MELT_GLACIER_CORR", HBV_melt_glacier_corr) LP("HBV_GLACIER_KMIN", HBV_glacier_Kmin) LP("GLAC_STORAGE_COEFF", glac_storage_coeff)
        LP("HBV_GLACIER_AG", HBV_glacier_Ag) LP("DEP_MAX", dep_max) LP("LAKE_PET_CORR", lake_PET_corr) LP("OW_PET_CORR", ow_PET_corr)
        LP("PARTITION_COEFF", partition_coeff) LP("CC_DECAY_COEFF", CC_decay_coeff)
        const veg_struct *V = h->GetVegetationProps();
        const veg_var_struct *VV = h->GetVegVarProps();
#define VP(n, f) M << "PARAM VEG " << k << " " << n << " " << V->f << "\n";
#define VVP(n, f) M << "PARAM VEG " << k << " " << n << " " << VV->f << "\n";
        VP("MAX_HEIGHT", max_height) VP("MAX_LAI", max_LAI) VP("SAI_HT_RATIO", SAI_ht_ratio) VP("MAX_CAPACITY", max_capacity)
        VP("MAX_SNOW_CAPACITY", max_snow_capacity) VP("RAIN_ICEPT_PCT", rain_icept_pct) VP("SNOW_ICEPT_PCT", snow_icept_pct) VP("PET_VEG_CORR", PET_veg_corr)
        VVP("VAR_CAPACITY", capacity) VVP("VAR_SNOW_CAPACITY", snow_capacity) VVP("VAR_RAIN_ICEPT_PCT", rain_icept_pct) VVP("VAR_SNOW_ICEPT_PCT", snow_icept_pct)
        for (int m = 0; m < pModel->GetNumSoilLayers(); m++) {
          const soil_struct *P = h->GetSoilProps(m);
#define SP(n, f) M << "PARAM SOIL" << m << " " << k << " " << n << " " << P->f << "\n";
          SP("POROSITY", porosity) SP("CAP_RATIO", cap_ratio) SP("PERC_COEFF", perc_coeff) SP("PET_CORRECTION", PET_correction)
          SP("BASEFLOW_COEFF", baseflow_coeff) SP("BASEFLOW_N", baseflow_n) SP("FIELD_CAPACITY", field_capacity) SP("SAT_WILT", sat_wilt)
          SP("HBV_BETA", HBV_beta) SP("MAX_CAP_RISE_RATE", max_cap_rise_rate) SP("MAX_PERC_RATE", max_perc_rate) SP("GR4J_X2", GR4J_x2) SP("GR4J_X3", GR4J_x3)
          M << "PARAM SOIL" << m << " " << k << " THICKNESS " << h->GetSoilThickness(m) << "\n";
        }
      }
    }
    for (int k = 0; k < nH && k < 64; k++) {
      const CHydroUnit *h = pModel->GetHydroUnit(k);
      M << "HRU " << k << " area " << h->GetArea() << " elev " << h->GetElevation() << " sb " << h->GetSubBasinIndex()
        << " type " << (int)h->GetHRUType();
      for (int m = 0; m < pModel->GetNumSoilLayers(); m++) M << " cap" << m << " " << h->GetSoilCapacity(m);
      M << "\n";
    }
  }

  FILE *fs = fopen((outdir + "state.f64").c_str(), "wb");
  FILE *ff = fopen((outdir + "forc.f64").c_str(), "wb");
  FILE *fb = fopen((outdir + "basin.f64").c_str(), "wb");
  FILE *fw = fopen((outdir + "wshed.f64").c_str(), "wb");
  FILE *fr = fopen((outdir + "res.f64").c_str(), "wb");
  if (!fs || !ff || !fb || !fw || !fr) { perror("open dump"); return 3; }

  std::vector<double> row(NS);
  auto dump_state = [&]() {
    for (int k = 0; k < nH; k++) {
      const CHydroUnit *h = pModel->GetHydroUnit(k);
      for (int i = 0; i < NS; i++) row[i] = h->GetStateVarValue(i);
      wr(fs, row.data(), NS);
    }
    for (int p = 0; p < NB; p++) {
      const CSubBasin *b = pModel->GetSubBasin(p);
      double v[6] = { b->GetOutflowRate(), b->GetLocalOutflowRate(), b->GetChannelStorage(), b->GetRivuletStorage(),
                      b->GetInflowHistory()[0], b->GetLatHistory()[0] };
      wr(fb, v, 6);
    }
    for (int p = 0; p < NB; p++) {
      const CReservoir *r = pModel->GetSubBasin(p)->GetReservoir();
      double v[4] = { 0.0, 0.0, 0.0, 0.0 };
      if (r != NULL) { v[0] = r->GetResStage(); v[1] = r->GetOutflowRate(false); v[2] = r->GetReservoirLosses(Options.timestep); v[3] = r->GetStorage(); }
      wr(fr, v, 4);
    }
    double w[3] = { pModel->_CumulInput, pModel->_CumulOutput, pModel->GetAveragePrecip() };
    wr(fw, w, 3);
  };
  auto dump_forc = [&]() {
    for (int k = 0; k < nH; k++) wr(ff, (const double *)pModel->GetHydroUnit(k)->GetForcingFunctions(), NF);
  };
  dump_state();
  dump_forc();  // forcings at t=0 (row 0); row n+1 = forcings used for step n ... see below

  int step = 0;
  for (double t = t_start; t < Options.duration - TIME_CORRECTION; t += Options.timestep) {
    if (max_steps >= 0 && step >= max_steps) break;
    pModel->UpdateTransientParams(Options, tt);
    pModel->ApplyStateOverrrides(Options, tt);
    pModel->RecalculateHRUDerivedParams(Options, tt);
    pModel->GetEnsemble()->StartTimeStepOps(pModel, Options, tt, e);
    pModel->UpdateHRUForcingFunctions(Options, tt);
    pModel->PrepareAssimilation(Options, tt);
    MassEnergyBalance(pModel, Options, tt);
    pModel->IncrementCumulInput(Options, tt);
    pModel->IncrementCumOutflow(Options, tt);
    JulianConvert(t + Options.timestep, Options.julian_start_day, Options.julian_start_year, Options.calendar, tt);
    pModel->WriteMinorOutput(Options, tt);
    pModel->UpdateDiagnostics(Options, tt);
    pModel->GetEnsemble()->CloseTimeStepOps(pModel, Options, tt, e);
    dump_state();
    dump_forc();  // row step+1 = forcings that drove step `step`
    step++;
  }
  fclose(fs); fclose(ff); fclose(fb); fclose(fw); fclose(fr);
  {  // CModel::GetCumulativeFlux(k, i, to): cumulative [mm] moved into (to = true) / out of (to = false) state variable i of HRU k by all processes
    FILE *fc = fopen((outdir + "cumflux.f64").c_str(), "wb");
    std::vector<double> row((size_t)NS * 2);
    for (int k = 0; k < nH; k++) {
      for (int i = 0; i < NS; i++) { row[2 * i] = pModel->GetCumulativeFlux(k, i, true); row[2 * i + 1] = pModel->GetCumulativeFlux(k, i, false); }
      wr(fc, row.data(), row.size());
    }
    fclose(fc);
  }
  {
    std::ofstream M((outdir + "meta.txt").c_str(), std::ios::app);
    M << "nsteps " << step << "\n";
  }
  fprintf(stderr, "ref_dump: %d steps, nHRU=%d NS=%d NB=%d -> %s\n", step, nH, NS, NB, outdir.c_str());
  return 0;
}
