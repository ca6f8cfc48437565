// rvh_flatten.cpp -- "compile" the parsed model into the flat, device-ready rvn_model_desc.
// This is where per-(class,date) work the reference repeats for every HRU and step is hoisted:
// convolution unit hydrographs (reference rebuilds them per HRU per step, src/Convolution.cpp:258),
// the calendar table (JulianConvert per step, src/RavenMain.cpp:162) and gauge sampling.
#include <algorithm>
#include "rvh_model.h"
#include "rvh_process.h"

void CModel::FlattenMemberParams(int member) {
  double *row = _f_ptab.data() + (size_t)member * _f_ptab_stride;
  for (int f = 0; f < RVN_G_COUNT; f++) row[_desc.glob_off + f] = globals.v[f];
  for (size_t c = 0; c < lu_classes.size(); c++)
    for (int f = 0; f < RVN_LU_COUNT; f++) row[_desc.lu_off + c * RVN_LU_COUNT + f] = lu_classes[c].S.v[f];
  for (size_t c = 0; c < veg_classes.size(); c++) {
    veg_struct V = veg_classes[c].V;
    // PRECIP_ICEPT_USER: canopy interception fractions come straight from the class (src/VegetationParams.cpp:150-154)
    V.v[RVN_V_VAR_RAIN_ICEPT_PCT] = V.v[RVN_V_RAIN_ICEPT_PCT];
    V.v[RVN_V_VAR_SNOW_ICEPT_PCT] = V.v[RVN_V_SNOW_ICEPT_PCT];
    for (int f = 0; f < RVN_V_COUNT; f++) row[_desc.veg_off + c * RVN_V_COUNT + f] = V.v[f];
  }
  for (size_t c = 0; c < soil_classes.size(); c++)
    for (int f = 0; f < RVN_S_COUNT; f++) row[_desc.soil_off + c * RVN_S_COUNT + f] = soil_classes[c].S.v[f];
  for (size_t c = 0; c < terrain_classes.size(); c++)
    for (int f = 0; f < RVN_T_COUNT; f++) row[_desc.terr_off + c * RVN_T_COUNT + f] = terrain_classes[c].T.v[f];
  // convolution unit hydrographs for this member, per land-use class
  const int nLU = (int)lu_classes.size();
  int ci = 0;
  for (int j = 0; j < GetNumProcesses(); j++) {
    const CmvConvolution *cv = dynamic_cast<const CmvConvolution *>(_pProcesses[j]);
    if (!cv) continue;
    for (int c = 0; c < nLU; c++) {
      size_t slot = ((size_t)member * nLU + c) * _f_nconv + ci;
      int N = 0;
      cv->GenerateUnitHydrograph(lu_classes[c].S, *_pOpt, &_f_conv_uh[slot * RVN_CONV_STORES], N);
      _f_conv_n[slot] = N;
    }
    ci++;
  }
}
const int32_t *CModel::MemberConvN(int member) const { return _f_conv_n.data() + (size_t)member * lu_classes.size() * _f_nconv; }
const double *CModel::MemberConvUH(int member) const { return _f_conv_uh.data() + (size_t)member * lu_classes.size() * _f_nconv * RVN_CONV_STORES; }

void CModel::SetCellForcing(int nCells, int nSteps, const double *cell_elev, const double *series, const int32_t *wt_ptr, const int32_t *wt_idx, const double *wt_val) {
  const int nH = GetNumHRUs();
  ExitGracefullyIf(nCells < 1 || nSteps < 1 || !cell_elev || !series || !wt_ptr || !wt_idx || !wt_val, "CModel::SetCellForcing: bad arguments");
  ExitGracefullyIf(wt_ptr[0] != 0, "CModel::SetCellForcing: wt_ptr[0] must be 0");
  for (int k = 0; k < nH; k++) {
    ExitGracefullyIf(wt_ptr[k + 1] <= wt_ptr[k], "CModel::SetCellForcing: every HRU needs at least one weighted grid cell");
    double sum = 0.0;
    for (int i = wt_ptr[k]; i < wt_ptr[k + 1]; i++) {
      ExitGracefullyIf(wt_idx[i] < 0 || wt_idx[i] >= nCells, "CModel::SetCellForcing: grid cell index out of range");
      sum += wt_val[i];
    }
    ExitGracefullyIf(fabs(sum - 1.0) > 0.05, "CModel::SetCellForcing: grid weights of an HRU do not sum to 1");   // the reference's CheckWeightArray tolerance
  }
  const size_t nnz = (size_t)wt_ptr[nH];
  _cell_forcing = true; _nCells = nCells;
  _cell_elev.assign(cell_elev, cell_elev + nCells);
  _cell_series.assign(series, series + (size_t)RVN_GF_COUNT * nCells * nSteps);
  _cell_ptr.assign(wt_ptr, wt_ptr + nH + 1); _cell_idx.assign(wt_idx, wt_idx + nnz); _cell_wt.assign(wt_val, wt_val + nnz);
}

const rvn_model_desc *CModel::Flatten(const optStruct &O, int nMembers) {
  ExitGracefullyIf(_pOpt == NULL, "CModel::Flatten: model must be initialized first");
  rvn_model_desc &d = _desc;
  memset(&d, 0, sizeof(d));
  const int nH = GetNumHRUs(), NB = GetNumSubBasins(), NS = GetNumStateVars(), nG = _cell_forcing ? _nCells : GetNumGauges();
  d.abi_version = RVN_ABI_VERSION;
  d.nHRU = nH; d.nSB = NB; d.nMembers = nMembers; d.nGauges = nG;
  // ---- options
  d.opt.timestep = O.timestep;
  d.opt.rainsnow = O.rainsnow; d.opt.pot_melt = O.pot_melt; d.opt.evaporation = O.evaporation; d.opt.ow_evaporation = O.ow_evaporation;
  d.opt.orocorr_temp = O.orocorr_temp; d.opt.orocorr_precip = O.orocorr_precip; d.opt.orocorr_pet = O.orocorr_PET;
  d.opt.routing = O.routing; d.opt.suppress_competitive_ET = O.suppressCompetitiveET; d.opt.snow_suppress_PET = O.snow_suppressPET;
  d.opt.allow_soil_overfill = O.allow_soil_overfill; d.opt.direct_evap = O.direct_evap; d.opt.lake_sv = _lake_sv; d.opt.keep_flux_balance = O.keep_flux_balance ? 1 : 0; d.opt.sol_method = O.sol_method; d.opt.max_iterations = O.max_iterations; d.opt.convergence_crit = O.convergence_crit;
  // ---- state variables
  d.NS = NS; d.nSoilLayers = _nSoilVars;
  _f_svtype.resize(NS); _f_svlayer.resize(NS);
  for (int i = 0; i < NS; i++) { _f_svtype[i] = (int32_t)GetStateVarType(i); _f_svlayer[i] = GetStateVarLayer(i); }
  d.sv_type = _f_svtype.data(); d.sv_layer = _f_svlayer.data();
  // ---- process program
  // A :ProcessGroup is one process of the model (one gate) but one device record per member (include/rvn_gpu.h rvn_process).
  const int NP = GetNumProcesses();
  _f_proc.clear(); _f_from.clear(); _f_to.clear();
  std::vector<int> rec_owner;   // model process index of every device record
  _f_nconv = 0;
  auto emit = [&](const CHydroProcessABC *pr, int owner) -> rvn_process & {
    rvn_process rec; memset(&rec, 0, sizeof(rec));
    pr->Compile(rec);
    rec.group = 0; rec.gconn0 = 0; rec.gnconn = 0; rec.pad_ = 0; rec.weight = 1.0;
    if (rec.op == RVN_OP_CONVOLVE) { rec.conn0 = -1; _f_nconv++; }
    else {
      ExitGracefullyIf(pr->GetNumConnections() > RVN_MAX_CONN, "CModel::Flatten: process has too many connections");
      rec.conn0 = (int32_t)_f_from.size();
      for (int q = 0; q < pr->GetNumConnections(); q++) {
        ExitGracefullyIf(pr->GetFromIndices()[q] < 0 || pr->GetToIndices()[q] < 0, "CModel::Flatten: process " + pr->GetProcessName() + " has an unconnected state variable");
        _f_from.push_back(pr->GetFromIndices()[q]); _f_to.push_back(pr->GetToIndices()[q]);
      }
    }
    _f_proc.push_back(rec); rec_owner.push_back(owner);
    return _f_proc.back();
  };
  _f_lat.clear(); _f_lat_kfrom.clear(); _f_lat_kto.clear(); _f_lat_frac.clear(); _f_lat_flag.clear(); _f_lat_chunk.clear();
  for (int j = 0; j < NP; j++) {
    if (const CLateralExchangeProcessABC *lf = dynamic_cast<const CLateralExchangeProcessABC *>(_pProcesses[j])) {
      // CModel::ApplyLateralProcess (src/Model.cpp:3126-3165): no connections -> nothing; gate of the FIRST source HRU; every
      // participating HRU must be enabled.  All of it is static, so a process that does not apply is simply not listed.
      const int nc = lf->GetNumLatConnections();
      if (nc == 0) continue;
      if (!lf->ShouldApply(_pHydroUnits[lf->GetFromHRUIndices()[0]])) continue;
      bool all_enabled = true;
      for (int q = 0; q < nc; q++) if (!_pHydroUnits[lf->GetFromHRUIndices()[q]]->IsEnabled() || !_pHydroUnits[lf->GetToHRUIndices()[q]]->IsEnabled()) all_enabled = false;
      if (!all_enabled) continue;
      rvn_lateral L; memset(&L, 0, sizeof(L));
      L.kind = lf->GetLateralKind(); L.sv_from = lf->GetLateralFromIndex(); L.sv_to = lf->GetLateralToIndex(); L.mixing_rate = lf->GetMixingRate(); L.method = lf->GetMethod();
      L.nconn = nc; L.conn0 = (int32_t)_f_lat_kfrom.size();
      L.nchunks = (int32_t)lf->ChunkOffsets().size() - 1; L.chunk0 = (int32_t)_f_lat_chunk.size();
      for (int q = 0; q < nc; q++) { _f_lat_kfrom.push_back(lf->GetFromHRUIndices()[q]); _f_lat_kto.push_back(lf->GetToHRUIndices()[q]); _f_lat_frac.push_back(lf->GetFractions()[q]);
        _f_lat_flag.push_back(lf->GetLateralKind() == RVN_LAT_EQUILIBRATE ? lf->GetFlags()[q] : 0); }
      for (size_t c = 0; c < lf->ChunkOffsets().size(); c++) _f_lat_chunk.push_back(lf->ChunkOffsets()[c]);
      _f_lat.push_back(L);
      continue;
    }
    const CProcessGroup *grp = dynamic_cast<const CProcessGroup *>(_pProcesses[j]);
    if (!grp) { emit(_pProcesses[j], j); continue; }
    const int n = grp->GetGroupSize();
    if (n == 0) continue;
    const int gconn0 = (int32_t)_f_from.size();
    const size_t first = _f_proc.size();
    for (int i = 0; i < n; i++) {
      ExitGracefullyIf(dynamic_cast<const CmvConvolution *>(grp->GetSubProcess(i)) != NULL || dynamic_cast<const CProcessGroup *>(grp->GetSubProcess(i)) != NULL,
                       "CModel::Flatten: convolutions and nested groups inside a :ProcessGroup are not supported by the B200 build");
      rvn_process &rec = emit(grp->GetSubProcess(i), j);
      rec.group = RVN_GRP_MEMBER | (i == 0 ? RVN_GRP_FIRST : 0) | (i == n - 1 ? RVN_GRP_LAST : 0);
      rec.weight = grp->GetWeight(i);
    }
    const int gnconn = (int32_t)_f_from.size() - gconn0;
    for (size_t r = first; r < _f_proc.size(); r++) { _f_proc[r].gconn0 = gconn0; _f_proc[r].gnconn = gnconn; }
  }
  d.nLat = (int32_t)_f_lat.size(); d.nLatConn = (int32_t)_f_lat_kfrom.size();
  if (_f_lat_kfrom.empty()) { _f_lat_kfrom.push_back(0); _f_lat_kto.push_back(0); _f_lat_frac.push_back(0.0); _f_lat_flag.push_back(0); }
  if (_f_lat_chunk.empty()) _f_lat_chunk.push_back(0);
  d.lat = _f_lat.empty() ? NULL : _f_lat.data(); d.lat_kfrom = _f_lat_kfrom.data(); d.lat_kto = _f_lat_kto.data(); d.lat_frac = _f_lat_frac.data(); d.lat_flag = _f_lat_flag.data();
  d.lat_chunk_ptr = _f_lat_chunk.data();
  const int NR = (int)_f_proc.size();
  ExitGracefullyIf(NR > 64, "CModel::Flatten: more than 64 process records (the per-HRU process gate is a 64-bit mask)");
  d.nProc = NR; d.nConn = (int32_t)_f_from.size(); d.proc = _f_proc.data(); d.conn_from = _f_from.data(); d.conn_to = _f_to.data();
  _f_mask.assign(nH, 0);
  for (int k = 0; k < nH; k++)
    for (int r = 0; r < NR; r++) if (_pProcesses[rec_owner[r]]->ShouldApply(_pHydroUnits[k])) _f_mask[k] |= (1ull << r);
  d.apply_mask = _f_mask.data();
  // ---- HRUs
  for (int a = 0; a < 5; a++) _f_hru_d[a].resize(nH);
  for (int a = 0; a < 6; a++) _f_hru_i[a].resize(nH);
  for (int k = 0; k < nH; k++) {
    const CHydroUnit *h = _pHydroUnits[k];
    _f_hru_d[0][k] = h->GetArea(); _f_hru_d[1][k] = h->GetElevation(); _f_hru_d[2][k] = h->GetLatRad();
    _f_hru_d[3][k] = h->GetSlope(); _f_hru_d[4][k] = h->GetAspect();
    _f_hru_i[0][k] = (int32_t)h->GetHRUType(); _f_hru_i[1][k] = h->lu_class; _f_hru_i[2][k] = h->veg_class;
    _f_hru_i[3][k] = h->soil_profile; _f_hru_i[4][k] = h->subbasin_index; _f_hru_i[5][k] = h->enabled ? 1 : 0;
  }
  d.hru_area = _f_hru_d[0].data(); d.hru_elev = _f_hru_d[1].data(); d.hru_lat_rad = _f_hru_d[2].data();
  d.hru_slope = _f_hru_d[3].data(); d.hru_aspect = _f_hru_d[4].data();
  d.hru_type = _f_hru_i[0].data(); d.hru_lu = _f_hru_i[1].data(); d.hru_veg = _f_hru_i[2].data();
  d.hru_profile = _f_hru_i[3].data(); d.hru_sb = _f_hru_i[4].data(); d.hru_enabled = _f_hru_i[5].data();
  // ---- profiles / classes
  d.nProfiles = (int32_t)soil_profiles.size(); d.nLU = (int32_t)lu_classes.size(); d.nVeg = (int32_t)veg_classes.size(); d.nSoilClasses = (int32_t)soil_classes.size();
  _f_prof_class.assign((size_t)d.nProfiles * RVN_MAX_SOIL_LAYERS, 0); _f_prof_thick.assign((size_t)d.nProfiles * RVN_MAX_SOIL_LAYERS, 0.0);
  ExitGracefullyIf(_nSoilVars > RVN_MAX_SOIL_LAYERS, "CModel::Flatten: too many soil layers");
  for (int r = 0; r < d.nProfiles; r++)
    for (int m = 0; m < soil_profiles[r].nHorizons && m < RVN_MAX_SOIL_LAYERS; m++) {
      _f_prof_class[(size_t)r * RVN_MAX_SOIL_LAYERS + m] = soil_profiles[r].soil_class[m];
      _f_prof_thick[(size_t)r * RVN_MAX_SOIL_LAYERS + m] = soil_profiles[r].thickness[m];
    }
  d.prof_class = _f_prof_class.data(); d.prof_thick = _f_prof_thick.data();
  d.glob_off = 0; d.lu_off = RVN_G_COUNT; d.veg_off = d.lu_off + d.nLU * RVN_LU_COUNT; d.soil_off = d.veg_off + d.nVeg * RVN_V_COUNT;
  d.terr_off = d.soil_off + d.nSoilClasses * RVN_S_COUNT; d.nTerrain = (int32_t)terrain_classes.size();
  _f_ptab_stride = d.terr_off + (d.nTerrain > 0 ? d.nTerrain : 1) * RVN_T_COUNT;
  _f_hru_terr.resize(nH);
  for (int k = 0; k < nH; k++) _f_hru_terr[k] = _pHydroUnits[k]->terrain_class;
  d.hru_terrain = _f_hru_terr.data();
  for (int r = 0; r < d.nProc; r++)
    if (d.proc[r].op == RVN_OP_BASE_TOPMODEL)
      for (int k = 0; k < nH; k++)
        ExitGracefullyIf(_pHydroUnits[k]->terrain_class < 0 && ((_f_mask[k] >> r) & 1ull), "CModel::Flatten: BASE_TOPMODEL needs a terrain class (TOPMODEL_LAMBDA) for every HRU it applies to");
  // class parameters an algorithm needs must have been given (reference: "required parameter not specified" from
  // CSoilClass / CLandUseClass::AutoCalculate*Props fed by each process's GetParticipatingParamList)
  {
    struct Need { int op; char cls; int f[5]; };
    static const Need needs[] = {
      {RVN_OP_BASE_CONSTANT, 'S', {RVN_S_MAX_BASEFLOW_RATE, -1}}, {RVN_OP_BASE_LINEAR_CONSTRAIN, 'S', {RVN_S_BASEFLOW_COEFF, RVN_S_FIELD_CAPACITY, -1}},
      {RVN_OP_BASE_THRESH_POWER, 'S', {RVN_S_MAX_BASEFLOW_RATE, RVN_S_BASEFLOW_N, RVN_S_BASEFLOW_THRESH, -1}},
      {RVN_OP_BASE_THRESH_STOR, 'S', {RVN_S_BASEFLOW_COEFF2, RVN_S_STORAGE_THRESHOLD, -1}},
      {RVN_OP_PERC_GAWSER, 'S', {RVN_S_MAX_PERC_RATE, RVN_S_FIELD_CAPACITY, -1}}, {RVN_OP_PERC_GAWSER_CONSTRAIN, 'S', {RVN_S_MAX_PERC_RATE, RVN_S_FIELD_CAPACITY, -1}},
      {RVN_OP_PERC_POWER_LAW, 'S', {RVN_S_MAX_PERC_RATE, RVN_S_PERC_N, -1}}, {RVN_OP_PERC_LINEAR_ANALYTIC, 'S', {RVN_S_PERC_COEFF, -1}},
      {RVN_OP_PERC_PRMS, 'S', {RVN_S_MAX_PERC_RATE, RVN_S_PERC_N, RVN_S_FIELD_CAPACITY, RVN_S_SAT_WILT, -1}},
      {RVN_OP_PERC_SACRAMENTO, 'S', {RVN_S_MAX_BASEFLOW_RATE, RVN_S_SAC_PERC_ALPHA, RVN_S_SAC_PERC_EXPON, RVN_S_FIELD_CAPACITY, RVN_S_SAT_WILT}},
      {RVN_OP_PERC_ASPEN, 'S', {RVN_S_PERC_ASPEN, -1}},
      {RVN_OP_INF_RATIONAL, 'L', {RVN_LU_PARTITION_COEFF, -1}}, {RVN_OP_INF_PDM, 'L', {RVN_LU_PDM_B, -1}},
      {RVN_OP_INF_VIC, 'S', {RVN_S_VIC_ZMAX, RVN_S_VIC_ZMIN, RVN_S_VIC_ALPHA, -1}}, {RVN_OP_INF_PRMS, 'S', {RVN_S_FIELD_CAPACITY, RVN_S_SAT_WILT, -1}},
      {RVN_OP_SOILEVAP_LINEAR, 'L', {RVN_LU_AET_COEFF, -1}}, {RVN_OP_SOILEVAP_PDM, 'L', {RVN_LU_PDM_B, -1}},
      {RVN_OP_SOILEVAP_HYMOD2, 'L', {RVN_LU_PDM_B, RVN_LU_HYMOD2_G, RVN_LU_HYMOD2_KMAX, RVN_LU_HYMOD2_EXP, -1}},
      {RVN_OP_SOILEVAP_UBC, 'S', {RVN_S_UBC_EVAP_SOIL_DEF, RVN_S_UBC_INFIL_SOIL_DEF, -1}},
      {RVN_OP_SOILEVAP_VIC, 'S', {RVN_S_VIC_ZMAX, RVN_S_VIC_ZMIN, RVN_S_VIC_ALPHA, RVN_S_VIC_EVAP_GAMMA, -1}},
      {RVN_OP_SOILEVAP_HYPR, 'L', {RVN_LU_MAX_DEP_AREA_FRAC, RVN_LU_PONDED_EXP, RVN_LU_DEP_MAX, -1}},
      {RVN_OP_SOILEVAP_SEQUEN, 'S', {RVN_S_FIELD_CAPACITY, RVN_S_SAT_WILT, -1}}, {RVN_OP_SOILEVAP_ROOT, 'S', {RVN_S_FIELD_CAPACITY, RVN_S_SAT_WILT, -1}},
      {RVN_OP_SOILEVAP_ROOT_CONSTRAIN, 'S', {RVN_S_FIELD_CAPACITY, RVN_S_SAT_WILT, -1}},
      {RVN_OP_SOILEVAP_AWBM, 'L', {RVN_LU_AWBM_AREAFRAC1, RVN_LU_AWBM_AREAFRAC2, -1}},
      {RVN_OP_INF_GREEN_AMPT, 'S', {RVN_S_HYDRAUL_COND, RVN_S_WETTING_FRONT_PSI, -1}},
      {RVN_OP_SOILEVAP_SACSMA, 'S', {RVN_S_UNAVAIL_FRAC, -1}},
      {RVN_OP_EXCHANGE_FLOW, 'S', {RVN_S_EXCHANGE_FLOW, -1}},
      {RVN_OP_OPEN_WATER_RIPARIAN, 'L', {RVN_LU_STREAM_FRACTION, -1}},
      {RVN_OP_INF_SCS, 'L', {RVN_LU_SCS_CN, -1}}, {RVN_OP_INF_SCS_NOABSTRACTION, 'L', {RVN_LU_SCS_CN, -1}},
      {RVN_OP_INF_GA_SIMPLE, 'S', {RVN_S_HYDRAUL_COND, RVN_S_WETTING_FRONT_PSI, -1}},
    };
    for (int r = 0; r < d.nProc; r++)
      if (d.proc[r].op == RVN_OP_DEPRESSION_OVERFLOW || d.proc[r].op == RVN_OP_SEEPAGE) {   // src/DepressionProcesses.cpp:44-75,215-222
        static const int need_f[4][5] = {{RVN_LU_DEP_THRESHOLD, RVN_LU_DEP_N, RVN_LU_DEP_MAX_FLOW, RVN_LU_DEP_MAX, -1}, {RVN_LU_DEP_THRESHOLD, RVN_LU_DEP_K, -1, -1, -1},
                                         {RVN_LU_DEP_THRESHOLD, RVN_LU_DEP_CRESTRATIO, -1, -1, -1}, {RVN_LU_DEP_SEEP_K, -1, -1, -1, -1}};
        const int row = (d.proc[r].op == RVN_OP_SEEPAGE) ? 3 : d.proc[r].ia[0];
        for (int j = 0; j < 5 && need_f[row][j] >= 0; j++)
          for (size_t c = 0; c < lu_classes.size(); c++)
            ExitGracefullyIf(lu_classes[c].S.v[need_f[row][j]] == NOT_SPECIFIED, std::string("CLandUseClass::AutoCalculateLandUseProps: required parameter ") + LandUseFieldName(need_f[row][j]) + " not specified for land use class " + lu_classes[c].tag);
      }
    for (int r = 0; r < d.nProc; r++)
      if (d.proc[r].op == RVN_OP_SOILEVAP_CHU)   // src/SoilEvaporation.cpp:189-193
        for (size_t c = 0; c < veg_classes.size(); c++)
          ExitGracefullyIf(veg_classes[c].V.v[RVN_V_CHU_MATURITY] == NOT_SPECIFIED, "CVegetationClass::AutoCalculateVegetationProps: required parameter CHU_MATURITY not specified for vegetation class " + veg_classes[c].tag);
    for (int r = 0; r < d.nProc; r++)
      if (d.proc[r].op == RVN_OP_INF_UBC) {   // CmvInfiltration::GetParticipatingParamList, src/Infiltration.cpp:181-190 (UBC_FLASH_PONDING is autogenerated)
        ExitGracefullyIf(globals.v[RVN_G_UBC_GW_SPLIT] == NOT_SPECIFIED, "CGlobalParams::AutoCalculateGlobalParams: required global parameter UBC_GW_SPLIT not specified");
        for (size_t c = 0; c < soil_classes.size(); c++)
          ExitGracefullyIf(soil_classes[c].S.v[RVN_S_MAX_PERC_RATE] == NOT_SPECIFIED || soil_classes[c].S.v[RVN_S_UBC_INFIL_SOIL_DEF] == NOT_SPECIFIED,
                           "CSoilClass::AutoCalculateSoilProps: required parameter MAX_PERC_RATE / UBC_INFIL_SOIL_DEF not specified for soil class " + soil_classes[c].tag);
      }
    for (int r = 0; r < d.nProc; r++)
      if (d.proc[r].op == RVN_OP_ABSTRACTION) {   // CmvAbstraction::GetParticipatingParamList, src/Abstraction.cpp:100-150
        const int a = d.proc[r].ia[0];
        const int f = (a == RVN_ABST_PERCENTAGE) ? RVN_LU_ABST_PERCENT : (a == RVN_ABST_SCS) ? RVN_LU_SCS_CN : RVN_LU_DEP_MAX;
        for (size_t c = 0; c < lu_classes.size(); c++)
          ExitGracefullyIf(lu_classes[c].S.v[f] == NOT_SPECIFIED, std::string("CLandUseClass::AutoCalculateLandUseProps: required parameter ") + LandUseFieldName(f) + " not specified for land use class " + lu_classes[c].tag);
      }
    for (int r = 0; r < d.nProc; r++)
      for (size_t q = 0; q < sizeof(needs) / sizeof(needs[0]); q++) {
        if (needs[q].op != d.proc[r].op) continue;
        for (int j = 0; j < 5 && needs[q].f[j] >= 0; j++) {
          const int f = needs[q].f[j];
          if (needs[q].cls == 'S') {
            for (size_t c = 0; c < soil_classes.size(); c++)
              ExitGracefullyIf(soil_classes[c].S.v[f] == NOT_SPECIFIED, std::string("CSoilClass::AutoCalculateSoilProps: required parameter ") + SoilFieldName(f) + " not specified for soil class " + soil_classes[c].tag);
          } else {
            for (size_t c = 0; c < lu_classes.size(); c++)
              ExitGracefullyIf(lu_classes[c].S.v[f] == NOT_SPECIFIED, std::string("CLandUseClass::AutoCalculateLandUseProps: required parameter ") + LandUseFieldName(f) + " not specified for land use class " + lu_classes[c].tag);
          }
        }
      }
  }
  d.ptab_stride = _f_ptab_stride;
  _f_ptab.assign((size_t)nMembers * _f_ptab_stride, 0.0);
  d.nConvProc = _f_nconv;
  _f_conv_n.assign((size_t)nMembers * d.nLU * (_f_nconv > 0 ? _f_nconv : 1), 0);
  _f_conv_uh.assign((size_t)nMembers * d.nLU * (_f_nconv > 0 ? _f_nconv : 1) * RVN_CONV_STORES, 0.0);
  for (int e = 0; e < nMembers; e++) {
    if (e == 0) FlattenMemberParams(0);
    else {  // identical until the ensemble sampler overwrites them
      std::copy(_f_ptab.begin(), _f_ptab.begin() + _f_ptab_stride, _f_ptab.begin() + (size_t)e * _f_ptab_stride);
      size_t nn = (size_t)d.nLU * _f_nconv;
      std::copy(_f_conv_n.begin(), _f_conv_n.begin() + nn, _f_conv_n.begin() + e * nn);
      std::copy(_f_conv_uh.begin(), _f_conv_uh.begin() + nn * RVN_CONV_STORES, _f_conv_uh.begin() + e * nn * RVN_CONV_STORES);
    }
  }
  d.ptab = _f_ptab.data(); d.conv_n = _f_conv_n.data(); d.conv_uh = _f_conv_uh.data();
  // ---- calendar: replay the reference time loop  for(t=0; t<duration-TIME_CORRECTION; t+=timestep)  (src/RavenMain.cpp:145)
  _f_times.clear();
  _f_veg_season.clear();
  _f_spec_basin.clear(); _f_spec_in.clear(); _f_spec_down.clear(); _f_spec_cum.clear();
  for (int p = 0; p < NB; p++) if (_pSubBasins[p]->pInflowHydro || _pSubBasins[p]->pInflowHydro2) _f_spec_basin.push_back(p);
  // direct-insertion assimilation: CModel::PrepareAssimilation (src/Assimilate.cpp:125-275) replayed over the run; its arrays persist from step to step
  _f_da_override.clear(); _f_da_obsQ.clear(); _f_da_obsQ2.clear(); _f_da_decay.clear(); _f_da_wt.clear();
  std::vector<double> DAlength(NB, 0.0), DAtimesince(NB, 0.0), DAobsQ(NB, 0.0), DAobsQ2(NB, 0.0), DADrainSum(NB, 0.0), DADownSum(NB, 0.0);
  std::vector<char> DAoverride(NB, 0);
  if (O.assimilate_flow) for (int p = 0; p < NB; p++) ExitGracefullyIf(_pSubBasins[p]->pReservoir != NULL, "CModel::Flatten: streamflow assimilation in a model with reservoirs is not supported by the B200 build");
  // reservoirs: curves padded to the longest, initial state, linked HRU (src/ModelInitialize.cpp:165-175)
  _f_res_basin.clear(); _f_res_hru.clear(); _f_res_np.clear(); _f_res_curves.clear(); _f_res_min_stage.clear(); _f_res_extract.clear(); _f_res_state0.clear();
  int nResPts = 0;
  for (int p = 0; p < NB; p++) if (_pSubBasins[p]->pReservoir) { _f_res_basin.push_back(p); nResPts = std::max(nResPts, (int)_pSubBasins[p]->pReservoir->_aStage.size()); }
  for (size_t r = 0; r < _f_res_basin.size(); r++) {
    const CReservoir *R = _pSubBasins[_f_res_basin[r]]->pReservoir;
    const int np = (int)R->_aStage.size();
    _f_res_hru.push_back(R->_hru_k); _f_res_np.push_back(np); _f_res_min_stage.push_back(R->_min_stage);
    const std::vector<double> *cv[RVN_RES_CURVES] = {&R->_aStage, &R->_aQ, &R->_aQunder, &R->_aArea, &R->_aVolume};
    for (int c = 0; c < RVN_RES_CURVES; c++) for (int i = 0; i < nResPts; i++) _f_res_curves.push_back(i < np ? (*cv[c])[i] : (*cv[c])[np - 1]);
    const double st0[RVN_RES_STATE] = {R->_stage, R->_stage_last, R->_Qout, R->_Qout_last, 0.0, 0.0, 0.0, 0.0};
    _f_res_state0.insert(_f_res_state0.end(), st0, st0 + RVN_RES_STATE);
  }
  {
    time_struct tt;
    JulianConvert(0.0, O.julian_start_day, O.julian_start_year, O.calendar, tt);
    double last_angle = 0.0;
    for (double t = 0.0; t < O.duration - TIME_CORRECTION; t += O.timestep) {
      rvn_time r; memset(&r, 0, sizeof(r));
      r.model_time = tt.model_time; r.julian_day = tt.julian_day; r.month = tt.month; r.day_of_month = tt.day_of_month; r.year = tt.year;
      r.day_changed = tt.day_changed ? 1 : 0;
      r.nn = (int)((tt.model_time + TIME_CORRECTION) / O.timestep);
      if (tt.day_changed) {  // src/UpdateForcings.cpp:55,172-176
        double mid_day = floor(tt.julian_day + TIME_CORRECTION) + 0.5;
        last_angle = DayAngle(mid_day, tt.year, O.calendar);
      }
      r.day_angle = last_angle;
      _f_times.push_back(r);
      for (size_t i = 0; i < _f_spec_basin.size(); i++) {   // user-specified basin inflows of this step (src/Solvers.cpp:543,586; src/SubBasin.cpp:950-957)
        const CSubBasin *b = _pSubBasins[_f_spec_basin[i]];
        const double area = _WatershedArea * M2_PER_KM2;
        _f_spec_in.push_back(b->GetSpecifiedInflow(t + O.timestep));
        _f_spec_down.push_back(b->GetDownstreamInflow(t));
        double sum = 0.0;
        sum += 0.5 * (b->GetSpecifiedInflow(tt.model_time) + b->GetSpecifiedInflow(tt.model_time + O.timestep)) * (O.timestep * SEC_PER_DAY);
        sum += b->GetDownstreamInflow(tt.model_time) * (O.timestep * SEC_PER_DAY);
        _f_spec_cum.push_back(sum / area * MM_PER_METER);
      }
      if (O.assimilate_flow) {
        if (!(tt.model_time < (O.assimilation_start - O.timestep / 2.0))) {
          int p = 0, pdown;
          double Qobs = 0, Qobs2 = 0;
          for (p = 0; p < NB; p++) { DADrainSum[p] = 0.0; DAlength[p] = 0.0; }
          const int nn = (int)((tt.model_time + TIME_CORRECTION) / O.timestep) + 1;   // end-of-step index
          for (int pp = NB - 1; pp >= 0; pp--) {   // downstream to upstream
            p = _aOrderedSBind[pp];
            pdown = _aDownstreamInds[p];
            bool ObsExists = false;
            if (_pSubBasins[p]->assimilate) {
              const CTimeSeries *obs = FlowObservation(_pSubBasins[p]);
              if (obs != NULL) { Qobs = obs->GetSampledValue(nn); Qobs2 = obs->GetSampledValue(nn + 1); ObsExists = true; }
            }
            if (ObsExists) {
              if (Qobs != RAV_BLANK_DATA) {
                DAlength[p] = 0.0; DAtimesince[p] = 0.0; DAoverride[p] = 1; DAobsQ[p] = Qobs; DAobsQ2[p] = Qobs2;
              } else {
                if (pdown != DOESNT_EXIST) DAlength[p] += _pSubBasins[pdown]->GetReachLength(); else DAlength[p] = 0;
                DAtimesince[p] += O.timestep; DAoverride[p] = 0; DAobsQ[p] = 0.0; DAobsQ2[p] = 0.0;
              }
            } else {
              if ((pdown != DOESNT_EXIST) && (!DAoverride[p]) && (!_pSubBasins[p]->disabled) && (!_pSubBasins[pdown]->disabled)) {
                DAlength[p] += _pSubBasins[pdown]->GetReachLength(); DAtimesince[p] = DAtimesince[pdown]; DAoverride[p] = 0;
              } else { DAlength[p] = 0.0; DAtimesince[p] = 0.0; DAoverride[p] = 0; }
            }
          }
          for (int pp = 0; pp < NB; pp++) { DADrainSum[p] = 0.0; DADownSum[p] = 0.0; }   // the reference's loop: p is not advanced (:238-242)
          for (int pp = 0; pp < NB; pp++) {   // upstream to downstream
            p = _aOrderedSBind[pp]; pdown = _aDownstreamInds[p];
            if (DAoverride[p]) DADrainSum[p] = _pSubBasins[p]->drainage_area;
            else if (pdown != DOESNT_EXIST) DADrainSum[pdown] += DADrainSum[p];
          }
          for (int pp = NB - 1; pp >= 0; pp--) {   // downstream to upstream
            p = _aOrderedSBind[pp]; pdown = _aDownstreamInds[p];
            if (DAoverride[p]) DADownSum[p] = _pSubBasins[p]->drainage_area;
            else if (pdown != DOESNT_EXIST) DADownSum[p] = DADownSum[pdown];
          }
        }
        const double distfact = O.assim_upstream_decay / M_PER_KM;   // [1/km] -> [1/m]
        for (int p = 0; p < NB; p++) {   // what AssimilationOverride / AssimilationBackPropagate (:67-118, 283-339) read this step
          _f_da_override.push_back(DAoverride[p] ? 1 : 0); _f_da_obsQ.push_back(DAobsQ[p]); _f_da_obsQ2.push_back(DAobsQ2[p]);
          _f_da_decay.push_back((_aDownstreamInds[p] != DOESNT_EXIST) ? exp(-distfact * DAlength[p]) : 1.0);
          double ECCCwt = 1.0;
          if (DADrainSum[p] != 0.0) ECCCwt = (_pSubBasins[p]->drainage_area - DADrainSum[p]) / (DADownSum[p] - DADrainSum[p]);
          _f_da_wt.push_back(ECCCwt);
        }
      }
      for (size_t r = 0; r < _f_res_basin.size(); r++) {   // CDemand::UpdateDemand (src/WaterDemands.cpp:204-216): sampled value of the step, blank = 0
        const CReservoir *R = _pSubBasins[_f_res_basin[r]]->pReservoir;
        double Qirr = 0.0;
        if (R->_pDemandTS) { Qirr = R->_pDemandTS->GetSampledValue((int)((tt.model_time + TIME_CORRECTION) / O.timestep)); if (Qirr == RAV_BLANK_DATA) Qirr = 0.0; Qirr = 1.0 * Qirr; }
        _f_res_extract.push_back(Qirr);
      }
      // month-interpolated relative vegetation height / LAI (CVegetationClass::RecalculateCanopyParams, src/VegetationParams.cpp:102-107;
      // the reference repeats this per HRU per step although it depends on (vegetation class, date) only)
      for (size_t v = 0; v < veg_classes.size(); v++) {
        _f_veg_season.push_back(InterpolateMo(veg_classes[v].V.relative_ht, tt, O.month_interp_method, O));
        _f_veg_season.push_back(InterpolateMo(veg_classes[v].V.relative_LAI, tt, O.month_interp_method, O));
      }
      JulianConvert(t + O.timestep, O.julian_start_day, O.julian_start_year, O.calendar, tt);
    }
  }
  const int nSteps = (int)_f_times.size();
  d.nSteps = nSteps; d.times = _f_times.data();
  if (_f_veg_season.empty()) _f_veg_season.push_back(0.0);
  d.veg_season = _f_veg_season.data();
  // ---- extraterrestrial radiation (CRadiation::EstimateShortwaveRadiation -> ClearSkySolarRadiation -> CalcETRadiation2,
  // src/Radiation.cpp:24-57,992-1032): needed only when a PET that a process consumes is estimated from it.  It depends on
  // (HRU geometry, day angle, position of the step inside the day), not on the member: one row per distinct date.
  d.nEtRows = 0; d.et_radia = NULL; d.et_row = NULL;
  {
    bool uses_pet = false, uses_owpet = false;
    for (int j = 0; j < NP; j++) {
      const int op = _f_proc[j].op;
      if (op == RVN_OP_SOILEVAP_ALL || op == RVN_OP_SOILEVAP_HBV || op == RVN_OP_SOILEVAP_GR4J || op == RVN_OP_SOILEVAP_TOPMODEL ||
          (op >= RVN_OP_SOILEVAP_LINEAR && op <= RVN_OP_SOILEVAP_AWBM) || op == RVN_OP_SOILEVAP_SACSMA || op == RVN_OP_SOILEVAP_CHU || op == RVN_OP_CANEVP_ALL || op == RVN_OP_CANSNOWEVP_ALL ||
          op == RVN_OP_CANEVP_MAXIMUM || op == RVN_OP_CANEVP_RUTTER || op == RVN_OP_CANSNOWEVP_MAXIMUM) uses_pet = true;
      if (op == RVN_OP_OPEN_WATER_EVAP || op == RVN_OP_OPEN_WATER_RIPARIAN || op == RVN_OP_LAKE_EVAP_BASIC || op == RVN_OP_SOILEVAP_HYPR) uses_owpet = true;
    }
    for (int p = 0; p < NB; p++) if (_pSubBasins[p]->pReservoir && _pSubBasins[p]->pReservoir->_hru_k >= 0) uses_owpet = true;   // reservoir evaporation reads OW_PET of the linked HRU (src/Reservoir.cpp:1634)
    auto consumed = [&](int method) { return (uses_pet && O.evaporation == method) || (uses_owpet && O.ow_evaporation == method); };
    bool need_atmos = false;   // PET methods fed by air pressure / humidity / shortwave radiation
    for (int mth = RVN_PET_TURC_1961; mth <= RVN_PET_PENMAN_COMBINATION; mth++) if (consumed(mth)) need_atmos = true;
    // net shortwave / longwave radiation (src/UpdateForcings.cpp:591-606): energy-balance PET and potential-melt methods
    const bool need_radiation = consumed(RVN_PET_PRIESTLEY_TAYLOR) || consumed(RVN_PET_PENMAN_COMBINATION) || O.pot_melt == RVN_POTMELT_EB || O.pot_melt == RVN_POTMELT_RESTRICTED;
    if (need_radiation) need_atmos = true;
    ExitGracefullyIf(need_radiation && !O.unsupported_radiation.empty(), "ParseMainInputFile: " + O.unsupported_radiation + " is not implemented in the B200 build");
    ExitGracefullyIf(need_radiation && GetStateVarIndex(RVN_SV_SNOW_ALBEDO) != DOESNT_EXIST, "CModel::Flatten: net radiation with an evolving SNOW_ALBEDO state variable is not built (the default snow albedo 0.8 is)");
    d.opt.need_radiation = need_radiation ? 1 : 0; d.opt.lw_radiation = O.LW_radiation; d.opt.sw_canopycorr = O.SW_canopycorr; d.opt.pad_opt_ = 0;
    for (int j = 0; j < NP; j++) if (_f_proc[j].op == RVN_OP_SUBLIMATION) need_atmos = true;   // reads air pressure, humidity, wind velocity
    ExitGracefullyIf(need_atmos && O.wind_velocity == RVN_WINDVEL_SQRT && (globals.v[RVN_G_WINDVEL_ICEPT] == NOT_SPECIFIED || globals.v[RVN_G_WINDVEL_SCALE] == NOT_SPECIFIED),
                     "CGlobalParams::AutoCalculateGlobalParams: required global parameters WINDVEL_ICEPT / WINDVEL_SCALE not specified");
    ExitGracefullyIf(need_atmos && !O.unsupported_atmos.empty(), "ParseMainInputFile: " + O.unsupported_atmos + " is not implemented in the B200 build");
    ExitGracefullyIf(need_atmos && O.timestep < 1.0 - 0.0001, "CModel::Flatten: radiation / humidity based PET methods and snow sublimation are built for daily steps only in the B200 build");
    if (consumed(RVN_PET_VAPDEFICIT))
      for (size_t c = 0; c < lu_classes.size(); c++)
        ExitGracefullyIf(lu_classes[c].S.v[RVN_LU_PET_VAP_COEFF] == NOT_SPECIFIED, "CLandUseClass::AutoCalculateLandUseProps: required parameter PET_VAP_COEFF not specified for land use class " + lu_classes[c].tag);
    ExitGracefullyIf(_cell_forcing && (consumed(RVN_PET_HARGREAVES) || consumed(RVN_PET_FROMMONTHLY) || O.evaporation == RVN_PET_FROMMONTHLY),
                     "CModel::Flatten: PET methods that read monthly station normals are not available with gridded forcing input");
    if (consumed(RVN_PET_HARGREAVES))
      for (int g = 0; g < GetNumGauges(); g++)
        ExitGracefullyIf(_pGauges[g]->aMaxTemp[0] == NOT_SPECIFIED || _pGauges[g]->aMinTemp[0] == NOT_SPECIFIED, "PET_HARGREAVES requires minimum and maximum monthly temperatures");
    d.opt.need_atmos = need_atmos ? 1 : 0;
    d.opt.air_pressure = O.air_pressure; d.opt.rel_humidity = O.rel_humidity; d.opt.sw_radiation = O.SW_radiation; d.opt.sw_cloudcorr = O.SW_cloudcovercorr; d.opt.wind_velocity = O.wind_velocity; d.opt.interception_factor = O.interception_factor;
    const bool need_et = consumed(RVN_PET_HARGREAVES_1985) || consumed(RVN_PET_OUDIN) || need_atmos;
    bool uses_cc = false;   // ColdContentBalance reads force_struct::day_length (src/SnowBalance.cpp:746)
    for (int j = 0; j < NP; j++) if (_f_proc[j].op == RVN_OP_SNOBAL_COLD_CONTENT) uses_cc = true;
    const bool need_dl = consumed(RVN_PET_HAMON) || uses_cc, need_ws = consumed(RVN_PET_MOHYSE);
    ExitGracefullyIf(need_et && O.SW_radiation == RVN_SW_RAD_NONE, "CModel::Flatten: :SWRadiationMethod SW_RAD_NONE leaves the extraterrestrial radiation unset; PET methods that read it are not built for that case");
    ExitGracefullyIf((consumed(RVN_PET_OUDIN) || need_dl || need_ws || consumed(RVN_PET_LINEAR_TEMP)) && O.timestep < 1.0 - 0.0001,
                     "CModel::Flatten: PET_OUDIN / PET_HAMON / PET_MOHYSE / PET_LINEAR_TEMP are built for daily steps only in the B200 build");
    if (consumed(RVN_PET_LINEAR_TEMP))
      for (size_t c = 0; c < lu_classes.size(); c++)
        ExitGracefullyIf(lu_classes[c].S.v[RVN_LU_PET_LIN_COEFF] == NOT_SPECIFIED, "CLandUseClass::AutoCalculateLandUseProps: required parameter PET_LIN_COEFF not specified for land use class " + lu_classes[c].tag);
    ExitGracefullyIf(need_ws && globals.v[RVN_G_MOHYSE_PET_COEFF] == NOT_SPECIFIED, "CModel::Flatten: PET_MOHYSE needs :GlobalParameter MOHYSE_PET_COEFF");
    d.day_length = NULL; d.sunset_angle = NULL; d.et_flat = NULL; d.opt_air_mass = NULL;
    const bool need_sd = (O.subdaily == 1) && (O.timestep < 1.0);   // SUBDAILY_SIMPLE (CalculateSubDailyCorrection returns 1 for daily steps)
    d.subdaily_corr = NULL;
    if (need_et || need_dl || need_ws || need_sd) {
      _f_et_row.assign(nSteps, 0);
      _f_et.clear(); _f_daylen.clear(); _f_sunset.clear(); _f_etflat.clear(); _f_mopt.clear(); _f_subdaily.clear();
      std::map<std::pair<double, double>, int> rows;
      for (int n = 0; n < nSteps; n++) {
        const double t_sol = _f_times[n].julian_day - floor(_f_times[n].julian_day) - 0.5;
        const std::pair<double, double> key(_f_times[n].day_angle, t_sol);
        std::map<std::pair<double, double>, int>::iterator it = rows.find(key);
        if (it == rows.end()) {
          const int r = (int)rows.size();
          rows[key] = r;
          const double declin = SolarDeclination(_f_times[n].day_angle), ecc = EccentricityCorr(_f_times[n].day_angle);
          for (int k = 0; k < nH; k++) {
            if (need_et) _f_et.push_back(CalcETRadiation2(_f_hru_d[2][k], _f_hru_d[4][k], declin, ecc, _f_hru_d[3][k], t_sol, t_sol + O.timestep, O.timestep >= 1.0));
            if (need_dl) {   // CRadiation::DayLength, src/Radiation.cpp:438-445
              const double hd = -tan(declin) * tan(_f_hru_d[2][k]);
              _f_daylen.push_back((hd >= 1.0) ? 0.0 : ((hd <= -1.0) ? 1.0 : acos(hd) / RVH_PI));
            }
            if (need_atmos) {   // CRadiation::ClearSkySolarRadiation, src/Radiation.cpp:1015-1032
              const double hd = -tan(declin) * tan(_f_hru_d[2][k]);
              const double day_length = (hd >= 1.0) ? 0.0 : ((hd <= -1.0) ? 1.0 : acos(hd) / RVH_PI);
              _f_etflat.push_back(CalcETRadiation2(_f_hru_d[2][k], _f_hru_d[4][k], declin, ecc, 0.0, t_sol, t_sol + O.timestep, O.timestep >= 1.0));
              _f_mopt.push_back(OpticalAirMass(_f_hru_d[2][k], declin, day_length, t_sol, O.timestep >= 1.0));
            }
            if (need_ws) _f_sunset.push_back(acos(-tan(_f_hru_d[2][k]) * tan(declin)));   // src/Evaporation.cpp:481
            if (need_sd) {   // CModel::CalculateSubDailyCorrection, SUBDAILY_SIMPLE (src/UpdateForcings.cpp:964-982); F.day_length is the day's (:174-175)
              const double hd = -tan(declin) * tan(_f_hru_d[2][k]);
              const double DL = (hd >= 1.0) ? 0.0 : ((hd <= -1.0) ? 1.0 : acos(hd) / RVH_PI);
              const double dawn = 0.5 - 0.5 * DL, dusk = 0.5 + 0.5 * DL;
              const double t = _f_times[n].julian_day - floor(_f_times[n].julian_day), dt = O.timestep;
              double corr = 0.0;
              if ((t > dawn) && (t + dt <= dusk)) corr = -0.5 * (cos(RVH_PI * (t + dt - dawn) / DL) - cos(RVH_PI * (t - dawn) / DL)) / O.timestep;
              else if ((t < dawn) && (t + dt >= dawn)) corr = -0.5 * (cos(RVH_PI * (t + dt - dawn) / DL) - 1) / O.timestep;
              else if ((t < dusk) && (t + dt >= dusk)) corr = -0.5 * (-1 - cos(RVH_PI * (t - dawn) / DL)) / O.timestep;
              _f_subdaily.push_back(corr);
            }
          }
          _f_et_row[n] = r;
        } else _f_et_row[n] = it->second;
      }
      d.nEtRows = (int32_t)rows.size(); d.et_row = _f_et_row.data();
      if (need_et) d.et_radia = _f_et.data();
      if (need_dl) d.day_length = _f_daylen.data();
      if (need_ws) d.sunset_angle = _f_sunset.data();
      if (need_sd) d.subdaily_corr = _f_subdaily.data();
      if (need_atmos) { d.et_flat = _f_etflat.data(); d.opt_air_mass = _f_mopt.data(); }
    }
  }
  // ---- gauge series sampled on the model step (src/UpdateForcings.cpp:98-155)
  _f_series.assign((size_t)RVN_GF_COUNT * nG * nSteps, 0.0);
  _f_gconst.assign((size_t)nG * RVN_GC_COUNT, 0.0);
  if (_cell_forcing) {   // one station per grid cell: series as handed over, unit corrections (grid-level :RainCorrection etc. are not built)
    ExitGracefullyIf(_cell_series.size() != (size_t)RVN_GF_COUNT * nG * nSteps, "CModel::Flatten: the gridded forcing arrays were given for a different number of steps");
    _f_series = _cell_series;
    for (int g = 0; g < nG; g++) {
      double *C = &_f_gconst[(size_t)g * RVN_GC_COUNT];
      C[RVN_GC_ELEVATION] = _cell_elev[g]; C[RVN_GC_RAIN_CORR] = 1.0; C[RVN_GC_SNOW_CORR] = 1.0; C[RVN_GC_TEMP_CORR] = 0.0;
      for (int m = 0; m < 48; m++) C[RVN_GC_MONTHLY_AVE_TEMP0 + m] = NOT_SPECIFIED;
    }
  }
  for (int g = 0; g < (_cell_forcing ? 0 : nG); g++) {
    const CGauge *G = _pGauges[g];
    ExitGracefullyIf(!G->GetTimeSeries(F_PRECIP) || !G->GetTimeSeries(F_TEMP_AVE),
                     "CModel::Flatten: every gauge must carry precipitation and temperature series in the B200 build");
    for (int n = 0; n < nSteps; n++) {
      const int nn = _f_times[n].nn;
      double *S = _f_series.data();
#define SER(f) S[((size_t)(f) * nG + g) * nSteps + n]
      SER(RVN_GF_PRECIP) = G->GetForcingValue(F_PRECIP, nn);
      SER(RVN_GF_SNOW_FRAC) = G->GetAverageSnowFrac(nn);
      SER(RVN_GF_TEMP_AVE) = G->GetForcingValue(F_TEMP_AVE, nn);
      SER(RVN_GF_TEMP_DAILY_AVE) = G->GetForcingValue(F_TEMP_DAILY_AVE, nn);
      SER(RVN_GF_TEMP_DAILY_MIN) = G->GetForcingValue(F_TEMP_DAILY_MIN, nn);
      SER(RVN_GF_TEMP_DAILY_MAX) = G->GetForcingValue(F_TEMP_DAILY_MAX, nn);
      SER(RVN_GF_PET) = G->GetForcingValue(F_PET, nn);
      SER(RVN_GF_OW_PET) = G->GetForcingValue(F_OW_PET, nn);
      SER(RVN_GF_POTENTIAL_MELT) = G->GetForcingValue(F_POTENTIAL_MELT, nn);
#undef SER
    }
    double *C = &_f_gconst[(size_t)g * RVN_GC_COUNT];
    C[RVN_GC_ELEVATION] = G->elevation; C[RVN_GC_RAIN_CORR] = G->rainfall_corr; C[RVN_GC_SNOW_CORR] = G->snowfall_corr; C[RVN_GC_TEMP_CORR] = G->temperature_corr;
    for (int m = 0; m < 12; m++) {
      C[RVN_GC_MONTHLY_AVE_TEMP0 + m] = G->aAveTemp[m]; C[RVN_GC_MONTHLY_AVE_PET0 + m] = G->aAvePET[m];
      C[RVN_GC_MONTHLY_MIN_TEMP0 + m] = G->aMinTemp[m]; C[RVN_GC_MONTHLY_MAX_TEMP0 + m] = G->aMaxTemp[m];
    }
  }
  d.gauge_series = _f_series.data(); d.gauge_const = _f_gconst.data();
  {   // force_struct::precip_5day of the gauges (src/UpdateForcings.cpp:107), only when a process reads it
    bool uses_p5 = false;
    for (int j = 0; j < NP; j++)
      if (_f_proc[j].op == RVN_OP_INF_SCS || _f_proc[j].op == RVN_OP_INF_SCS_NOABSTRACTION || (_f_proc[j].op == RVN_OP_ABSTRACTION && _f_proc[j].ia[0] == RVN_ABST_SCS)) uses_p5 = true;
    d.gauge_precip5 = NULL;
    if (uses_p5) {
      ExitGracefullyIf(_cell_forcing, "CModel::Flatten: INF_SCS needs gauge precipitation series (the 5-day antecedent precipitation is not derived from gridded input in the B200 build)");
      _f_p5.assign((size_t)nG * nSteps, 0.0);
      for (int g = 0; g < nG; g++)
        for (int n = 0; n < nSteps; n++) _f_p5[(size_t)g * nSteps + n] = _pGauges[g]->GetTimeSeries(F_PRECIP)->GetAvgValue(_f_times[n].model_time - 5.0, 5.0) * 5.0;
      d.gauge_precip5 = _f_p5.data();
    }
  }
  // gauge weights -> CSR (the reference keeps three dense [nHRU][nGauges] tables; here they hold the same values)
  _f_wt_ptr.assign(nH + 1, 0); _f_wt_idx.clear(); _f_wt.clear();
  if (_cell_forcing) {
    _f_wt_ptr = _cell_ptr; _f_wt_idx = _cell_idx;
    _f_wt.resize(_cell_wt.size() * 3);
    for (size_t i = 0; i < _cell_wt.size(); i++) _f_wt[3 * i] = _f_wt[3 * i + 1] = _f_wt[3 * i + 2] = _cell_wt[i];
  }
  for (int k = 0; k < (_cell_forcing ? 0 : nH); k++) {
    for (int g = 0; g < nG; g++) {
      const double w = _aGaugeWeights[(size_t)k * nG + g];
      if (w != 0.0) { _f_wt_idx.push_back(g); _f_wt.push_back(w); _f_wt.push_back(w); _f_wt.push_back(w); }
    }
    _f_wt_ptr[k + 1] = (int32_t)_f_wt_idx.size();
  }
  (void)0;
  if (_f_wt_idx.empty()) { _f_wt_idx.push_back(0); _f_wt.assign(3, 0.0); }
  d.wt_ptr = _f_wt_ptr.data(); d.wt_idx = _f_wt_idx.data(); d.wt_val = _f_wt.data();
  // ---- routing network
  for (int a = 0; a < 7; a++) _f_sb_i[a].resize(NB);
  for (int a = 0; a < 7; a++) _f_sb_d[a].resize(NB);
  int maxqin = 1, maxqlat = 1;
  for (int p = 0; p < NB; p++) {
    const CSubBasin *b = _pSubBasins[p];
    if (b->GetInflowHistorySize() > maxqin) maxqin = b->GetInflowHistorySize();
    if (b->GetLatHistorySize() > maxqlat) maxqlat = b->GetLatHistorySize();
  }
  _f_uh.assign((size_t)NB * maxqlat, 0.0); _f_rh.assign((size_t)NB * maxqin, 0.0);
  for (int p = 0; p < NB; p++) {
    const CSubBasin *b = _pSubBasins[p];
    _f_sb_i[0][p] = _aDownstreamInds[p]; _f_sb_i[1][p] = _aOrderedSBind[p]; _f_sb_i[2][p] = _aSubBasinOrder[p];
    _f_sb_i[3][p] = b->is_headwater ? 1 : 0; _f_sb_i[4][p] = b->nSegments; _f_sb_i[5][p] = b->GetInflowHistorySize();
    _f_sb_d[0][p] = b->basin_area; _f_sb_d[1][p] = b->rain_corr; _f_sb_d[2][p] = b->snow_corr; _f_sb_d[3][p] = b->temperature_corr;
    double K = 0, X = 0;
    if ((O.routing == RVN_ROUTE_MUSKINGUM || O.routing == RVN_ROUTE_MUSKINGUM_CUNGE || O.routing == RVN_ROUTE_STORAGECOEFF) && b->pChannel) {
      double dx = b->reach_length / (double)(b->nSegments);
      K = b->GetMuskingumK(dx);
      if (O.routing != RVN_ROUTE_STORAGECOEFF) X = b->GetMuskingumX(dx);
    }
    _f_sb_d[4][p] = K; _f_sb_d[5][p] = X; _f_sb_d[6][p] = K;
    for (int n = 0; n < b->GetLatHistorySize(); n++) _f_uh[(size_t)p * maxqlat + n] = b->aUnitHydro[n];
    for (int n = 0; n < b->GetInflowHistorySize(); n++) _f_rh[(size_t)p * maxqin + n] = b->aRouteHydro[n];
  }
  d.sb_down = _f_sb_i[0].data(); d.sb_order = _f_sb_i[1].data(); d.sb_level = _f_sb_i[2].data(); d.sb_headwater = _f_sb_i[3].data();
  d.sb_nseg = _f_sb_i[4].data(); d.sb_nqin = _f_sb_i[5].data();
  for (int p = 0; p < NB; p++) _f_sb_i[6][p] = _pSubBasins[p]->GetLatHistorySize();
  d.sb_nqlat = _f_sb_i[6].data();
  d.sb_area = _f_sb_d[0].data(); d.sb_rain_corr = _f_sb_d[1].data(); d.sb_snow_corr = _f_sb_d[2].data(); d.sb_temp_corr = _f_sb_d[3].data();
  d.sb_musk_K = _f_sb_d[4].data(); d.sb_musk_X = _f_sb_d[5].data(); d.sb_K_reach = _f_sb_d[6].data();
  d.max_nqin = maxqin; d.max_nqlat = maxqlat; d.sb_unit_hydro = _f_uh.data(); d.sb_route_hydro = _f_rh.data();
  d.watershed_area = _WatershedArea;
  d.nRes = (int32_t)_f_res_basin.size(); d.nResPts = nResPts;
  d.res_basin = d.nRes ? _f_res_basin.data() : NULL; d.res_hru = d.nRes ? _f_res_hru.data() : NULL; d.res_np = d.nRes ? _f_res_np.data() : NULL;
  d.res_curves = d.nRes ? _f_res_curves.data() : NULL; d.res_min_stage = d.nRes ? _f_res_min_stage.data() : NULL;
  d.res_extract = d.nRes ? _f_res_extract.data() : NULL; d.res_state0 = d.nRes ? _f_res_state0.data() : NULL;
  d.assimilate_flow = O.assimilate_flow ? 1 : 0; d.assimilation_fact = O.assimilation_fact;
  d.da_override = NULL; d.da_obsQ = d.da_obsQ2 = d.da_decay = d.da_wt = d.sb_drain_area = d.sb_da_corr = NULL;
  if (O.assimilate_flow) {
    const double distfact = O.assim_upstream_decay / M_PER_KM;
    _f_sb_drain.assign(NB, 0.0); _f_sb_dacorr.assign(NB, 1.0);
    for (int p = 0; p < NB; p++) { _f_sb_drain[p] = _pSubBasins[p]->drainage_area; _f_sb_dacorr[p] = exp(-distfact * _pSubBasins[p]->reach_length); }
    d.da_override = _f_da_override.data(); d.da_obsQ = _f_da_obsQ.data(); d.da_obsQ2 = _f_da_obsQ2.data(); d.da_decay = _f_da_decay.data(); d.da_wt = _f_da_wt.data();
    d.sb_drain_area = _f_sb_drain.data(); d.sb_da_corr = _f_sb_dacorr.data();
  }
  d.nSpec = (int32_t)_f_spec_basin.size();
  d.spec_basin = d.nSpec ? _f_spec_basin.data() : NULL; d.spec_in = d.nSpec ? _f_spec_in.data() : NULL;
  d.spec_down = d.nSpec ? _f_spec_down.data() : NULL; d.spec_cum = d.nSpec ? _f_spec_cum.data() : NULL;
  // ---- ROUTE_HYDROLOGIC: rating curves flow -> cross-section area
  d.nChannels = 0; d.nChanPts = 0; d.chan_Q = d.chan_A = NULL; d.sb_channel = NULL; d.sb_Qmult = d.sb_reach_m = NULL;
  d.sb_c_ref = d.sb_diff_alpha = NULL;
  if (O.routing == RVN_ROUTE_HYDROLOGIC || O.routing == RVN_ROUTE_DIFFUSIVE_VARY) {
    int npts = 0;
    for (size_t c = 0; c < channels.size(); c++) npts = std::max(npts, (int)channels[c]->aQ.size());
    d.nChannels = (int32_t)channels.size(); d.nChanPts = npts;
    _f_chan_Q.assign((size_t)d.nChannels * npts, 0.0); _f_chan_A.assign((size_t)d.nChannels * npts, 0.0);
    for (size_t c = 0; c < channels.size(); c++) {
      const CChannelXSect *ch = channels[c];
      ExitGracefullyIf((int)ch->aQ.size() != npts, "CModel::Flatten: channel profiles with different rating-curve lengths are not supported");
      for (int i = 0; i < npts; i++) { _f_chan_Q[c * npts + i] = ch->aQ[i]; _f_chan_A[c * npts + i] = ch->aXArea[i]; }
    }
    _f_sb_channel.assign(NB, -1); _f_sb_Qmult.assign(NB, 1.0); _f_sb_reach.assign(NB, 0.0);
    for (int p = 0; p < NB; p++) {
      const CSubBasin *b = _pSubBasins[p];
      for (size_t c = 0; c < channels.size(); c++) if (channels[c] == b->pChannel) _f_sb_channel[p] = (int32_t)c;
      if (b->pChannel) _f_sb_Qmult[p] = b->pChannel->GetFlowMultiplier(b->slope, b->mannings_n);
      _f_sb_reach[p] = b->reach_length;
    }
    if (O.routing == RVN_ROUTE_DIFFUSIVE_VARY) {
      _f_sb_cref.assign(NB, 0.0); _f_sb_alpha.assign(NB, 0.0);
      for (int p = 0; p < NB; p++) { _f_sb_cref[p] = _pSubBasins[p]->c_ref; _f_sb_alpha[p] = _pSubBasins[p]->diff_alpha; }
      d.sb_c_ref = _f_sb_cref.data(); d.sb_diff_alpha = _f_sb_alpha.data();
    }
    d.chan_Q = _f_chan_Q.data(); d.chan_A = _f_chan_A.data(); d.sb_channel = _f_sb_channel.data(); d.sb_Qmult = _f_sb_Qmult.data(); d.sb_reach_m = _f_sb_reach.data();
  }
  // ---- forcing perturbations (EnKF): samples start neutral; CEnKFEnsemble::UpdateModel fills them per member
  const int nP = (int)perturbations.size();
  d.nPert = nP; d.pad3_ = 0; d.pert_forcing = NULL; d.pert_adj = NULL; d.pert_mask = NULL; d.pert_eps = NULL;
  if (nP > 0) {
    _f_pert_forcing.resize(nP); _f_pert_adj.resize(nP);
    bool any_group = false;
    for (int i = 0; i < nP; i++) {
      const force_perturb &P = perturbations[i];
      int f = -1;
      if (P.forcing == F_PRECIP) f = RVN_PF_PRECIP; else if (P.forcing == F_RAINFALL) f = RVN_PF_RAINFALL;
      else if (P.forcing == F_SNOWFALL) f = RVN_PF_SNOWFALL; else if (P.forcing == F_TEMP_AVE) f = RVN_PF_TEMP_AVE;
      // the reference only ever applies these four (src/UpdateForcings.cpp:474,558-560); TEMP_DAILY_MIN/MAX perturbations are never matched
      ExitGracefullyIf(f < 0 && P.forcing != F_TEMP_DAILY_MIN && P.forcing != F_TEMP_DAILY_MAX,
                       "CHydroUnit::AdjustHRUForcing: Only PRECIP, SNOWFALL, RAINFALL, TEMP_DAILY_MIN, and TEMP_AVE are supported for forcing perturbation. ");
      _f_pert_forcing[i] = (f < 0) ? RVN_PF_TEMP_AVE : f;
      _f_pert_adj[i] = (f < 0) ? RVN_ADJ_ADDITIVE : (int)P.adj_type;
      if (P.kk != DOESNT_EXIST) any_group = true;
    }
    if (any_group) {
      _f_pert_mask.assign((size_t)nP * nH, 0);
      for (int i = 0; i < nP; i++) {
        if (perturbations[i].kk == DOESNT_EXIST) { for (int k = 0; k < nH; k++) _f_pert_mask[(size_t)i * nH + k] = 1; continue; }
        const std::vector<int> &g = hru_groups[perturbations[i].kk].hru;
        for (size_t q = 0; q < g.size(); q++) _f_pert_mask[(size_t)i * nH + g[q]] = 1;
      }
      d.pert_mask = _f_pert_mask.data();
    }
    _f_pert_eps.assign((size_t)nMembers * d.nSteps * nP, 0.0);
    for (int e = 0; e < nMembers; e++)
      for (int n = 0; n < d.nSteps; n++)
        for (int i = 0; i < nP; i++) {
          const bool unmatched = (perturbations[i].forcing == F_TEMP_DAILY_MIN || perturbations[i].forcing == F_TEMP_DAILY_MAX);
          _f_pert_eps[((size_t)e * d.nSteps + n) * nP + i] = (!unmatched && perturbations[i].adj_type == ADJ_MULTIPLICATIVE) ? 1.0 : 0.0;
        }
    d.pert_forcing = _f_pert_forcing.data(); d.pert_adj = _f_pert_adj.data(); d.pert_eps = _f_pert_eps.data();
  }
  return &d;
}

void CModel::PackBasinState(std::vector<double> &qin, std::vector<double> &qlat, std::vector<double> &qout, std::vector<double> &scal) const {
  const int NB = GetNumSubBasins();
  qin.assign((size_t)NB * _desc.max_nqin, 0.0); qlat.assign((size_t)NB * _desc.max_nqlat, 0.0);
  qout.assign((size_t)NB * RVN_MAX_RIVER_SEGS, 0.0); scal.assign((size_t)NB * 8, 0.0);
  for (int p = 0; p < NB; p++) {
    const CSubBasin *b = _pSubBasins[p];
    for (int n = 0; n < b->GetInflowHistorySize(); n++) qin[(size_t)p * _desc.max_nqin + n] = b->aQinHist[n];
    for (int n = 0; n < b->GetLatHistorySize(); n++) qlat[(size_t)p * _desc.max_nqlat + n] = b->aQlatHist[n];
    for (int sgm = 0; sgm < b->nSegments; sgm++) qout[(size_t)p * RVN_MAX_RIVER_SEGS + sgm] = b->aQout[sgm];
    double *s = &scal[(size_t)p * 8];
    s[0] = b->QoutLast; s[1] = b->QlatLast; s[2] = b->Qlocal; s[3] = b->QlocLast; s[4] = b->channel_storage; s[5] = b->rivulet_storage;
  }
}

// ---- direct-insertion streamflow assimilation (src/Assimilate.cpp) -----------------------------------------------------------
const CTimeSeries *CModel::FlowObservation(const CSubBasin *b) const {   // IsContinuousFlowObs2 (:12-19): first HYDROGRAPH series of the basin
  for (size_t i = 0; i < observed.size(); i++) if (observed[i].pTS != NULL && observed[i].loc_ID == b->ID && observed[i].name == "HYDROGRAPH") return observed[i].pTS;
  return NULL;
}
void CModel::ApplyInitialAssimilation(const optStruct &O) {
  _pre_assim_q0.clear(); _pre_assim_q0last.clear();
  if (!O.assimilate_flow) return;
  if (0.0 < (O.assimilation_start - O.timestep / 2.0)) return;   // PrepareAssimilation returns before the assimilation start (:128)
  for (int p = 0; p < GetNumSubBasins(); p++) {
    CSubBasin *b = _pSubBasins[p];
    if (!b->assimilate) continue;
    const CTimeSeries *obs = FlowObservation(b);
    if (obs == NULL) continue;
    const double Qobs = obs->GetSampledValue(1);   // nn == 1: the end of the first step
    if (Qobs == RAV_BLANK_DATA) continue;
    ExitGracefullyIf(b->pReservoir != NULL, "ApplyInitialAssimilation: streamflow assimilation at a reservoir outlet is not supported by the B200 build");
    if (_pre_assim_q0.empty()) { _pre_assim_q0.assign(GetNumSubBasins(), 0.0); _pre_assim_q0last.assign(GetNumSubBasins(), 0.0); for (int q = 0; q < GetNumSubBasins(); q++) { _pre_assim_q0[q] = _pSubBasins[q]->GetOutflowRate(); _pre_assim_q0last[q] = _pSubBasins[q]->QoutLast; } }
    b->SetQout(Qobs);
    for (size_t n = 0; n < b->aQinHist.size(); n++) b->aQinHist[n] = Qobs;   // SetQinHist with a constant history
  }
}

static std::string DecDaysToHours(double dec_date) {   // src/CommonFunctions.cpp:385-402 (not truncated): hh:mm:ss.ss of the fractional day
  const double hours = (dec_date - floor(dec_date)) * 24.0, mind = (hours - floor(hours)) * 60.0;
  double sec = (mind - floor(mind)) * 60.0;
  int hr = (int)(floor(hours)), min = (int)(floor(mind));
  if (sec >= 59.995) { min++; sec = 0.0; }
  if (min == 60) { hr++; min = 0; }
  if (hr == 24) hr = 0;
  char b[32]; snprintf(b, sizeof b, "%2.2d:%2.2d:%05.2f", hr, min, sec);
  return b;
}
void CModel::WriteSolutionFile(const std::string &filename, const optStruct &O, double model_time, const double *state, const double *qin, const double *qlat,
                               const double *qout, const double *scal, bool full_precision, const double *res_state) const {
  FILE *f = fopen(filename.c_str(), "w");
  ExitGracefullyIf(f == NULL, "CModel::WriteMajorOutput: Unable to open output file " + filename + " for writing.");
  const char *fmt_s = full_precision ? "%.17g" : "%.5f";   // HRU table: std::fixed, precision 5
  const char *fmt_b = full_precision ? "%.17g" : "%.5f";   // the stream stays fixed/5 for the basin block as well
  time_struct tts, tt;
  JulianConvert(0, O.julian_start_day, O.julian_start_year, O.calendar, tts);
  JulianConvert(model_time, O.julian_start_day, O.julian_start_year, O.calendar, tt);
  fprintf(f, ":StartTime %s %s\n", tts.date_string.c_str(), DecDaysToHours(tts.julian_day).c_str());
  fprintf(f, ":TimeStamp %s %s\n", tt.date_string.c_str(), DecDaysToHours(tt.julian_day).c_str());
  fprintf(f, ":TimeStep %.17g\n", O.timestep);
  const int NS = GetNumStateVars(), nH = GetNumHRUs(), NB = GetNumSubBasins(), M = 80;
  for (int j = 0; j < (int)ceil(NS / (double)(M)); j++) {   // blocks of 80 state variables
    const int mini = j * M, maxi = (NS < (j + 1) * M) ? NS : (j + 1) * M;
    fprintf(f, ":HRUStateVariableTable\n  :Attributes,");
    for (int i = mini; i < maxi; i++) fprintf(f, "%s%s", GetStateVarName(i).c_str(), (i != NS - 1) ? "," : "");
    fprintf(f, "\n  :Units,");
    for (int i = mini; i < maxi; i++) fprintf(f, "%s%s", "none", (i != NS - 1) ? "," : "");
    fprintf(f, "\n");
    for (int k = 0; k < nH; k++) {
      fprintf(f, "  %lld,", _pHydroUnits[k]->ID);
      for (int i = mini; i < maxi; i++) { fprintf(f, fmt_s, state[(size_t)k * NS + i]); if (i != NS - 1) fprintf(f, ","); }
      fprintf(f, "\n");
    }
    fprintf(f, ":EndHRUStateVariableTable\n");
  }
  fprintf(f, ":BasinStateVariables\n");
  for (int p = 0; p < NB; p++) {
    const CSubBasin *b = _pSubBasins[p];
    const double *sc = scal + (size_t)p * 8;
    fprintf(f, "  :BasinIndex %lld,%s\n", b->ID, b->name.c_str());
    fprintf(f, "    :ChannelStorage, "); fprintf(f, fmt_b, sc[4]); fprintf(f, "\n");
    fprintf(f, "    :RivuletStorage, "); fprintf(f, fmt_b, sc[5]); fprintf(f, "\n");
    fprintf(f, "    :Qout,%d", b->nSegments);
    for (int i = 0; i < b->nSegments; i++) { fprintf(f, ","); fprintf(f, fmt_b, qout[(size_t)p * RVN_MAX_RIVER_SEGS + i]); }
    fprintf(f, ","); fprintf(f, fmt_b, sc[0]); fprintf(f, "\n");
    fprintf(f, "    :Qlat,%d", b->GetLatHistorySize());
    for (int i = 0; i < b->GetLatHistorySize(); i++) { fprintf(f, ","); fprintf(f, fmt_b, qlat[(size_t)p * _desc.max_nqlat + i]); if (i % 200 == 199) fprintf(f, "\n    "); }
    fprintf(f, ","); fprintf(f, fmt_b, sc[1]); fprintf(f, "\n");
    fprintf(f, "    :Qin ,%d", b->GetInflowHistorySize());
    for (int i = 0; i < b->GetInflowHistorySize(); i++) { fprintf(f, ","); fprintf(f, fmt_b, qin[(size_t)p * _desc.max_nqin + i]); if (i % 200 == 199) fprintf(f, "\n    "); }
    fprintf(f, "\n");
    if (b->pReservoir != NULL && res_state != NULL) {   // CReservoir::WriteToSolutionFile (src/Reservoir.cpp:1855-1863)
      int r = 0;
      for (int q = 0; q < p; q++) if (_pSubBasins[q]->pReservoir) r++;
      const double *rs = res_state + (size_t)r * RVN_RES_STATE;
      fprintf(f, "    :ResFlow, "); fprintf(f, fmt_b, rs[2]); fprintf(f, ","); fprintf(f, fmt_b, rs[3]); fprintf(f, "\n");
      fprintf(f, "    :ResStage, "); fprintf(f, fmt_b, rs[0]); fprintf(f, ","); fprintf(f, fmt_b, rs[1]); fprintf(f, "\n");
    }
  }
  fprintf(f, ":EndBasinStateVariables\n");
  fclose(f);
}

// ---- Hydrographs.csv (reference CModel::WriteOutputFileHeaders / WriteMinorOutput, src/StandardOutput.cpp:199-243,790-836) ----------------
// Written after the run from what the device recorded: qout_rec[nsteps][NB] = CSubBasin::GetOutflowRate after every step, q0[NB] the outflow
// and q0last[NB] the _QoutLast of the initial condition, precip[nsteps] = GetAveragePrecip.  Default options of the reference: period-ending
// time stamps, step-averaged hydrographs (0.5 * (Qout + QoutLast)), observed series next to the modelled one, 6 significant digits.
static std::string StreamDouble(double v) { char b[64]; snprintf(b, sizeof b, "%g", v); return b; }   // operator<<(double) of a default ostream
void CModel::WriteHydrographsFile(const std::string &filename, const optStruct &O, const double *qout_rec, const double *q0, const double *q0last,
                                  const double *precip, int nsteps) const {
  FILE *f = fopen(filename.c_str(), "w");
  ExitGracefullyIf(f == NULL, "CModel::WriteOutputFileHeaders: Unable to open output file " + filename + " for writing.");
  const int NB = GetNumSubBasins();
  for (int p = 0; p < NB; p++)
    ExitGracefullyIf(_pSubBasins[p]->gauged && !_pSubBasins[p]->disabled && _pSubBasins[p]->pReservoir != NULL,
                     "CModel::WriteHydrographsFile: the reservoir-inflow columns of gauged reservoir basins are not written by the B200 build");
  fprintf(f, "time,date,hour,precip [mm/day]");
  for (int p = 0; p < NB; p++) {
    const CSubBasin *b = _pSubBasins[p];
    if (!(b->gauged && !b->disabled)) continue;
    if (b->name == "") fprintf(f, ",ID=%lld [m3/s]", b->ID); else fprintf(f, ",%s [m3/s]", b->name.c_str());
    for (size_t i = 0; i < observed.size(); i++)
      if (observed[i].pTS != NULL && observed[i].loc_ID == b->ID && observed[i].name == "HYDROGRAPH") {
        if (b->name == "") fprintf(f, ",ID=%lld (observed) [m3/s]", b->ID); else fprintf(f, ",%s (observed) [m3/s]", b->name.c_str());
      }
  }
  fprintf(f, "\n");
  for (int n = 0; n <= nsteps; n++) {   // row 0 = the initial condition (WriteMinorOutput at t = 0), row n = the end of step n
    const double t = n * O.timestep;
    time_struct tt;
    JulianConvert(t, O.julian_start_day, O.julian_start_year, O.calendar, tt);
    fprintf(f, "%s,%s,%s", StreamDouble(tt.model_time).c_str(), tt.date_string.c_str(), DecDaysToHours(tt.julian_day).c_str());
    if (n != 0) fprintf(f, ",%s", StreamDouble(precip[n - 1]).c_str()); else fprintf(f, ",---");
    for (int p = 0; p < NB; p++) {
      const CSubBasin *b = _pSubBasins[p];
      if (!(b->gauged && !b->disabled)) continue;
      // row 0 is written before the first PrepareAssimilation call replaces the initial flows of the gauges by their first observation
      const double Qnow = (n == 0) ? (_pre_assim_q0.empty() ? q0[p] : _pre_assim_q0[p]) : qout_rec[(size_t)(n - 1) * NB + p];
      const double Qlast = (n == 0) ? (_pre_assim_q0.empty() ? q0last[p] : _pre_assim_q0last[p]) : ((n == 1) ? q0[p] : qout_rec[(size_t)(n - 2) * NB + p]);
      // CSubBasin::GetIntegratedOutflow(tstep) / (tstep * SEC_PER_DAY), src/SubBasin.cpp:929-934
      fprintf(f, ",%s", StreamDouble(0.5 * (Qnow + Qlast) * (O.timestep * SEC_PER_DAY) / (O.timestep * SEC_PER_DAY)).c_str());
      for (size_t i = 0; i < observed.size(); i++)
        if (observed[i].pTS != NULL && observed[i].loc_ID == b->ID && observed[i].name == "HYDROGRAPH") {
          const double val = observed[i].pTS->GetAvgValue(tt.model_time, O.timestep);
          if ((val != RAV_BLANK_DATA) && (tt.model_time > 0)) fprintf(f, ",%s", StreamDouble(val).c_str()); else fprintf(f, ",");
        }
    }
    fprintf(f, "\n");
  }
  fclose(f);
}
