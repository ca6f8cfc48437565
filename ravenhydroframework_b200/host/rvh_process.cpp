// rvh_process.cpp -- host half of the process plugins: state-variable registration order, connection
// lists and conditions exactly as the reference constructs them while parsing :HydrologicProcesses
// (src/ParseInput.cpp case 200..), so that the SV catalogue and connection tables are integer-identical
// to the reference build.  Rates are computed on the device (csrc/rvn_processes.cuh).
#include "rvh_process.h"
#include <algorithm>
#include <map>

rvn_hru_type StringToHRUType(const std::string &sin) {
  std::string s = StringToUppercase(sin);
  if (s == "GLACIER") return RVN_HRU_GLACIER;
  if (s == "LAKE") return RVN_HRU_LAKE;
  if (s == "WATER") return RVN_HRU_WATER;
  if (s == "ROCK") return RVN_HRU_ROCK;
  if (s == "WETLAND") return RVN_HRU_WETLAND;
  if (s == "STANDARD") return RVN_HRU_STANDARD;
  if (s == "MASKED_GLACIER") return RVN_HRU_MASKED_GLACIER;
  ExitGracefully("StringToHRUType: unrecognized HRU type " + sin);
  return RVN_HRU_STANDARD;
}

bool CHydroProcessABC::ShouldApply(const CHydroUnit *pHRU) const {
  if (!pHRU->IsEnabled()) return false;
  for (size_t i = 0; i < _conditions.size(); i++) {
    const condition &c = _conditions[i];
    if (c.basis == BASIS_HRU_TYPE) {
      rvn_hru_type typ2 = StringToHRUType(c.data);
      if ((c.compare_method == COMPARE_IS_EQUAL) && (pHRU->GetHRUType() != typ2)) return false;
      if ((c.compare_method == COMPARE_NOT_EQUAL) && (pHRU->GetHRUType() == typ2)) return false;
    } else if (c.basis == BASIS_LANDCLASS) {
      const std::string &nm = pModel->lu_classes[pHRU->lu_class].tag;
      if ((c.data != nm) && (c.compare_method == COMPARE_IS_EQUAL)) return false;
      if ((c.data == nm) && (c.compare_method == COMPARE_NOT_EQUAL)) return false;
    } else if (c.basis == BASIS_VEGETATION) {
      const std::string &nm = pModel->veg_classes[pHRU->veg_class].tag;
      if ((c.data != nm) && (c.compare_method == COMPARE_IS_EQUAL)) return false;
      if ((c.data == nm) && (c.compare_method == COMPARE_NOT_EQUAL)) return false;
    } else {   // BASIS_HRU_GROUP: CModel::IsInHRUGroup (src/Model.cpp:465-477; an unknown group name means "not in the group")
      bool is_in_group = false;
      for (size_t g = 0; g < pModel->hru_groups.size(); g++)
        if (pModel->hru_groups[g].name == c.data) { is_in_group = pModel->hru_groups[g].IsInGroup(pHRU->global_k); break; }
      if ((!is_in_group) && (c.compare_method == COMPARE_IS_EQUAL)) return false;
      if ((is_in_group) && (c.compare_method == COMPARE_NOT_EQUAL)) return false;
    }
  }
  return true;
}
void CHydroProcessABC::AddCondition(condition_basis basis, comparison cmp, const std::string &data) {
  condition c; c.basis = basis; c.compare_method = cmp; c.data = data;
  _conditions.push_back(c);
}
void CHydroProcessABC::Redirect(int toSVindex, int newToSVindex) {
  bool found = false;
  for (size_t q = 0; q < iTo.size(); q++)
    if (iTo[q] == toSVindex) { iTo[q] = newToSVindex; found = true; }
  ExitGracefullyIf(!found, "CHydroProcessABC::Redirect: state variable in :-->RedirectFlow command not a recipient of the preceding process");
}

void CmvGeneric::Compile(rvn_process &rec) const {
  rec.op = (int32_t)_op;
  rec.nconn = GetNumConnections();
  rec.conn0 = 0;
  for (int i = 0; i < 5; i++) rec.ia[i] = _ia[i];
  rec.da[0] = _da[0]; rec.da[1] = _da[1];
}

// ---- ProcessGroup -----------------------------------------------------------------------------------
void CalcWeightsFromUniformNums(const double *aVals, double *aWeights, int N) {
  double sum = 0.0;
  for (int q = 0; q < (N - 1); q++) {
    aWeights[q] = (1.0 - sum) * (1.0 - pow(1.0 - aVals[q], 1.0 / (N - q - 1)));
    sum += aWeights[q];
  }
  aWeights[N - 1] = 1.0 - sum;
}
void CProcessGroup::SetWeights(const double *aWts, int nVal) {
  if (nVal != GetGroupSize()) { WriteWarning("CProcessGroup::SetWeights: Incorrect number of weights for process group. Weights will be ignored"); return; }
  double sum = 0.0;
  for (int q = 0; q < nVal; q++) { _aWeights[q] = aWts[q]; sum += _aWeights[q]; }
  if (fabs(sum - 1.0) > 0.01) for (int q = 0; q < nVal; q++) _aWeights[q] /= sum;
  if (fabs(sum - 1.0) > REAL_SMALL) WriteWarning("CProcessGroup::SetWeights: Process group weights do not sum to 1.0");
}
void CProcessGroup::CalcWeights(const double *aVals, int nVal) {
  if (nVal != GetGroupSize() - 1) { WriteWarning("CProcessGroup::SetWeights: Incorrect number of numbers for process group. Weights will be ignored"); return; }
  CalcWeightsFromUniformNums(aVals, _aWeights.data(), GetGroupSize());
}
void CProcessGroup::Initialize() {
  iFrom.clear(); iTo.clear();
  for (size_t j = 0; j < _pSubProcesses.size(); j++) {
    _pSubProcesses[j]->Initialize();
    for (int q = 0; q < _pSubProcesses[j]->GetNumConnections(); q++) { iFrom.push_back(_pSubProcesses[j]->GetFromIndices()[q]); iTo.push_back(_pSubProcesses[j]->GetToIndices()[q]); }
  }
}

// ---- LateralFlush -----------------------------------------------------------------------------------
CmvLatFlush::CmvLatFlush(int from_sv_ind, int to_sv_ind, int from_HRU_grp, int to_HRU_grp, bool constrain_to_SBs, CModel *pM)
    : CLateralExchangeProcessABC("LateralFlush", RVN_LAT_FLUSH, pM), _kk_from(from_HRU_grp), _kk_to(to_HRU_grp), _constrain_to_SBs(constrain_to_SBs) {
  _iFromLat = from_sv_ind; _iToLat = to_sv_ind;
  ExitGracefullyIf(to_HRU_grp < 0 || to_HRU_grp >= (int)pM->hru_groups.size(), "CmvLatFlush::unrecognized 'to' HRU group specified in :LateralFlush/:LateralDivert command");
  ExitGracefullyIf(from_HRU_grp < 0 || from_HRU_grp >= (int)pM->hru_groups.size(), "CmvLatFlush::unrecognized 'from' HRU group specified in :LateralFlush/:LateralDivert command");
  ExitGracefullyIf(from_sv_ind == DOESNT_EXIST, "CmvLatFlush::unrecognized 'from' state variable specified in :LateralFlush/:LateralDivert command");
  ExitGracefullyIf(to_sv_ind == DOESNT_EXIST, "CmvLatFlush::unrecognized 'to' state variable specified in :LateralFlush/:LateralDivert command");
}
void CmvLatFlush::Initialize() {   // src/LatFlush.cpp:55-177
  _kFrom.clear(); _kTo.clear(); _aFrac.clear(); _aFlag.clear(); _chunk_ptr.assign(1, 0);
  const int nH = pModel->GetNumHRUs();
  std::vector<char> inFrom(nH, 0), inTo(nH, 0);
  { const std::vector<int> &g = pModel->hru_groups[_kk_from].hru; for (size_t i = 0; i < g.size(); i++) inFrom[g[i]] = 1; }
  { const std::vector<int> &g = pModel->hru_groups[_kk_to].hru; for (size_t i = 0; i < g.size(); i++) inTo[g[i]] = 1; }
  if (_constrain_to_SBs) {
    bool source_sink_issue = false;
    for (int p = 0; p < pModel->GetNumSubBasins(); p++) {
      const CSubBasin *b = pModel->GetSubBasin(p);
      if (b->disabled) continue;
      for (size_t ks = 0; ks < b->hru_index.size(); ks++) {   // sources
        const int k1 = b->hru_index[ks];
        if (!inFrom[k1]) continue;
        double Asum = 0.0;
        int nRecipients = 0;
        for (size_t ks2 = 0; ks2 < b->hru_index.size(); ks2++) {   // recipients
          const int k2 = b->hru_index[ks2];
          if (!inTo[k2]) continue;
          if (k1 != k2) {
            const double area = pModel->GetHydroUnit(k2)->GetArea();
            _kFrom.push_back(k1); _kTo.push_back(k2); _aFrac.push_back(area);
            Asum += area; nRecipients++;
          } else source_sink_issue = true;
        }
        const size_t q = _aFrac.size();
        for (int qq = 0; qq < nRecipients; qq++) _aFrac[q - qq - 1] = _aFrac[q - qq - 1] / Asum;   // fraction of the recipient areas
      }
      if ((int)_kFrom.size() > _chunk_ptr.back()) _chunk_ptr.push_back((int)_kFrom.size());
    }
    if (_kFrom.empty()) WriteWarning("CmvLatFlush::Initialize: no connections found between from and to HRU groups in any of the basins. if INTERBASIN flag not used, only transfer within basins is supported. ");
    if (source_sink_issue) WriteWarning("CmvLatFlush::Initialize: one or more HRUs belongs to both source and recipient HRU groups in LateralFlush or LateralDivert process. This may have unintentended consequences. ");
  } else {
    int kToSB = DOESNT_EXIST;
    for (int k = 0; k < nH; k++)
      if (inTo[k]) {
        ExitGracefullyIf(kToSB != DOESNT_EXIST, "LatFlush::Initialize - only one HRU in the model can recieve flush output if INTERBASIN is used. More than one recipient HRU found.");
        kToSB = k;
      }
    for (int k = 0; k < nH; k++)
      if (inFrom[k] && (kToSB != DOESNT_EXIST) && pModel->GetHydroUnit(k)->IsEnabled()) { _kFrom.push_back(k); _kTo.push_back(kToSB); _aFrac.push_back(1.0); }
    if (!_kFrom.empty()) _chunk_ptr.push_back((int)_kFrom.size());
    if (_kFrom.empty()) WriteWarning("CmvLatFlush::Initialize: no connections found between from and to HRU groups. HRU Group population of zero?");
  }
}

CmvLatEquilibrate::CmvLatEquilibrate(int sv_ind, int HRU_grp, double mixing_rate, bool constrain_to_SBs, CModel *pM)
    : CLateralExchangeProcessABC("LateralEquilibrate", RVN_LAT_EQUILIBRATE, pM), _kk(HRU_grp), _constrain_to_SBs(constrain_to_SBs) {
  _iFromLat = _iToLat = sv_ind; _mixing_rate = mixing_rate;
  ExitGracefullyIf(HRU_grp < 0 || HRU_grp >= (int)pM->hru_groups.size(), "CmvLatEquilibrate::unrecognized HRU group specified in :LatEquilibrate command");
  ExitGracefullyIf(sv_ind == DOESNT_EXIST, "CmvLatEquilibrate::unrecognized state variable specified in :LatEquilibrate command");
}
void CmvLatEquilibrate::Initialize() {   // src/LatEquilibrate.cpp:49-148 (connections) and the control flow of :176-210 (flags)
  _kFrom.clear(); _kTo.clear(); _aFrac.clear(); _aFlag.clear(); _chunk_ptr.assign(1, 0);
  const int nH = pModel->GetNumHRUs(), NB = pModel->GetNumSubBasins();
  std::vector<char> inGrp(nH, 0);
  { const std::vector<int> &g = pModel->hru_groups[_kk].hru; for (size_t i = 0; i < g.size(); i++) inGrp[g[i]] = 1; }
  std::vector<int> kFrom, kTo, halfconn(NB > 0 ? NB : 1, 0);
  std::vector<double> Asum(NB > 0 ? NB : 1, 0.0);
  if (_constrain_to_SBs) {
    for (int p = 0; p < NB; p++) {
      const CSubBasin *b = pModel->GetSubBasin(p);
      int kToSB = DOESNT_EXIST;
      for (size_t ks = 0; ks < b->hru_index.size(); ks++) {
        const int k = b->hru_index[ks];
        if (inGrp[k] && (kToSB == DOESNT_EXIST)) { kToSB = k; Asum[p] += pModel->GetHydroUnit(k)->GetArea(); }   // first found is the mixing unit
        else if (inGrp[k] && (kToSB != DOESNT_EXIST)) {
          kFrom.push_back(k); kTo.push_back(kToSB);
          Asum[p] += pModel->GetHydroUnit(k)->GetArea(); halfconn[p]++;
        }
      }
    }
    if (kFrom.empty()) WriteWarning("CmvLatEquilibrate::Initialize: no connections found between from and to HRU groups in any of the basins. if INTERBASIN flag not used, only transfer within basins is supported. ");
  } else {
    int kToSB = DOESNT_EXIST;
    for (int k = 0; k < nH; k++)
      if (inGrp[k] && (kToSB == DOESNT_EXIST)) { kToSB = k; Asum[0] += pModel->GetHydroUnit(k)->GetArea(); }
    for (int k = 0; k < nH; k++)   // includes the mixing unit itself, whose area is counted twice (as in the reference)
      if (inGrp[k] && (kToSB != DOESNT_EXIST)) { kFrom.push_back(k); kTo.push_back(kToSB); Asum[0] += pModel->GetHydroUnit(k)->GetArea(); halfconn[0]++; }
  }
  const int half = (int)kFrom.size(), nConn = 2 * half;
  std::vector<int> qF(nConn), qT(nConn), qP(nConn), qFlag(nConn);
  for (int q = 0; q < nConn; q++) {
    if (q < half) { qF[q] = kFrom[q]; qT[q] = kTo[q]; }
    else { qF[q] = kTo[q - half]; qT[q] = kFrom[q - half]; }
  }
  // GetLateralExchange's bookkeeping (p, plast, qstart) depends on the connection list only: replay it once
  int plast = -1, qstart = 0;
  for (int q = 0; q < nConn; q++) {
    int p = pModel->GetHydroUnit(qF[q])->GetSubBasinIndex();
    if (!_constrain_to_SBs) p = 0;
    int flag = 0;
    if (p != plast) { flag |= 1; plast = p; qstart = q; }                 // bit 0: add the recipient's share to the pool first
    if ((q >= qstart) && (q < qstart + halfconn[p])) flag |= 2;            // bit 1: step 1 (mix into the pool); else step 2 (hand back)
    qP[q] = p; qFlag[q] = flag;
  }
  // chunks: one per subbasin (its connections in the reference's order), a single one with INTERBASIN
  std::vector<int> plist;
  for (int q = 0; q < nConn; q++) if (std::find(plist.begin(), plist.end(), qP[q]) == plist.end()) plist.push_back(qP[q]);
  std::map<int, std::vector<int> > byp;
  for (int q = 0; q < nConn; q++) byp[qP[q]].push_back(q);
  for (size_t i = 0; i < plist.size(); i++) {
    const std::vector<int> &v = byp[plist[i]];
    for (size_t j = 0; j < v.size(); j++) {
      const int q = v[j];
      _kFrom.push_back(qF[q]); _kTo.push_back(qT[q]); _aFlag.push_back(qFlag[q]);
      _aFrac.push_back(pModel->GetHydroUnit(qT[q])->GetArea() / Asum[qP[q]]);   // (Ato / _Asum[p])
    }
    _chunk_ptr.push_back((int)_kFrom.size());
  }
}

CmvLatRedistribute::CmvLatRedistribute(int sv_ind, double max_snow_height, int method, CModel *pM)
    : CLateralExchangeProcessABC("SnowRedistribute", RVN_LAT_REDISTRIBUTE, pM) {
  _iFromLat = _iToLat = sv_ind; _mixing_rate = max_snow_height; _method = method;
  ExitGracefullyIf(sv_ind == DOESNT_EXIST, "CmvLatRedistribute::unrecognized state variable specified in :SnowRedistribute command");
}
void CmvLatRedistribute::Initialize() {   // src/SnowRedistribute.cpp:59-93
  _kFrom.clear(); _kTo.clear(); _aFrac.clear(); _aFlag.clear(); _chunk_ptr.assign(1, 0);
  ExitGracefullyIf(_src.empty(), "CmvLatRedistribute::Initialize(): no lateral connections specified. Missing the :LateralConnections command?");
  const int nH = pModel->GetNumHRUs(), nC = (int)_src.size();
  std::map<long long, int> hru_of_id;
  for (int k = 0; k < nH; k++) hru_of_id.insert(std::make_pair(pModel->GetHydroUnit(k)->ID, k));
  std::vector<int> kF(nC), kT(nC);
  for (int q = 0; q < nC; q++) {
    std::map<long long, int>::const_iterator a = hru_of_id.find(_src[q]), b = hru_of_id.find(_dst[q]);
    ExitGracefullyIf(a == hru_of_id.end() || b == hru_of_id.end(), "CmvLatRedistribute::Initialize: HRU ID in :LateralConnections does not exist");
    kF[q] = a->second; kT[q] = b->second;
  }
  // chunks = connected components of the connection graph (HRUs no other chunk touches), each in the table's order
  std::vector<int> parent(nH);
  for (int k = 0; k < nH; k++) parent[k] = k;
  struct UF { static int find(std::vector<int> &p, int x) { while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; } return x; } };
  for (int q = 0; q < nC; q++) { int a = UF::find(parent, kF[q]), b = UF::find(parent, kT[q]); if (a != b) parent[a] = b; }
  std::vector<int> roots;
  std::map<int, std::vector<int> > byroot;
  for (int q = 0; q < nC; q++) {
    const int r = UF::find(parent, kF[q]);
    if (!byroot.count(r)) roots.push_back(r);
    byroot[r].push_back(q);
  }
  for (size_t i = 0; i < roots.size(); i++) {
    const std::vector<int> &v = byroot[roots[i]];
    for (size_t j = 0; j < v.size(); j++) { _kFrom.push_back(kF[v[j]]); _kTo.push_back(kT[v[j]]); _aFrac.push_back(_w[v[j]]); _aFlag.push_back(0); }
    _chunk_ptr.push_back((int)_kFrom.size());
  }
}

// ---- Precipitation ---------------------------------------------------------------------------------
void CmvPrecipitation::GetParticipatingStateVarList(sv_type *aSV, int *aLev, int &nSV) {
  nSV = 1; aSV[0] = RVN_SV_ATMOS_PRECIP; aLev[0] = DOESNT_EXIST;
}
void CmvPrecipitation::Initialize() {
  int iAtmos = pModel->GetStateVarIndex(RVN_SV_ATMOS_PRECIP);
  ExitGracefullyIf(pModel->StateVarExists(RVN_SV_ICE_THICKNESS), "CmvPrecipitation: ICE_THICKNESS (frozen lake) connections are not supported yet");
  DynamicSpecifyConnections(13);
  const sv_type tgt[11] = {RVN_SV_PONDED_WATER, RVN_SV_CANOPY, RVN_SV_CANOPY_SNOW, RVN_SV_DEPRESSION, RVN_SV_ATMOSPHERE, RVN_SV_SNOW,
                           RVN_SV_SNOW_LIQ, RVN_SV_SNOW_DEFICIT, RVN_SV_NEW_SNOW, RVN_SV_NTYPES /*lake*/, RVN_SV_SURFACE_WATER};
  for (int q = 0; q < 11; q++) {
    iFrom[q] = iAtmos;
    iTo[q] = (q == 9) ? pModel->GetLakeStorageIndex() : pModel->GetStateVarIndex(tgt[q]);
  }
  iFrom[11] = iTo[11] = pModel->GetStateVarIndex(RVN_SV_COLD_CONTENT);
  iFrom[12] = iTo[12] = pModel->GetStateVarIndex(RVN_SV_SNOW_DEPTH);
  for (int q = 0; q < 13; q++) {  // dummy slots for missing state variables
    if (iTo[q] == DOESNT_EXIST) iTo[q] = iAtmos;
    if (iFrom[q] == DOESNT_EXIST) iFrom[q] = iAtmos;
  }
}

// ---- Convolution -------------------------------------------------------------------------------------
CmvConvolution::CmvConvolution(rvn_convol_type type, int to_index, CModel *pM, int conv_index)
    : CmvGeneric("Convolve", RVN_OP_CONVOLVE, pM), _type(type), _conv_index(conv_index) {
  const int N = MAX_CONVOL_STORES;
  DynamicSpecifyConnections(2 * N + 1);
  iFrom[0] = iTo[0] = pModel->GetStateVarIndex(RVN_SV_CONV_STOR, 0 + conv_index * N);
  for (int i = 0; i < N; i++) {
    iFrom[i + 1] = pModel->GetStateVarIndex(RVN_SV_CONV_STOR, i + conv_index * N);
    iTo[i + 1] = to_index;
    if (i < N - 1) {
      iFrom[N + i + 1] = pModel->GetStateVarIndex(RVN_SV_CONV_STOR, i + conv_index * N);
      iTo[N + i + 1] = pModel->GetStateVarIndex(RVN_SV_CONV_STOR, i + 1 + conv_index * N);
    }
  }
  iFrom[2 * N] = iTo[2 * N] = pModel->GetStateVarIndex(RVN_SV_CONVOLUTION, conv_index);
  SetIntArg(0, conv_index);
  SetIntArg(1, to_index);
  SetIntArg(2, iFrom[1]);
  SetIntArg(3, iFrom[2 * N]);
  SetIntArg(4, (int)type);
  // the device kernel streams the stores; they must be contiguous in the catalogue
  for (int i = 1; i < N; i++)
    ExitGracefullyIf(iFrom[i + 1] != iFrom[1] + i, "CmvConvolution: CONV_STOR state variables are not contiguous");
}
void CmvConvolution::GetParticipatingStateVarList(rvn_convol_type, sv_type *aSV, int *aLev, int &nSV, int conv_index) {
  const int N = MAX_CONVOL_STORES;
  nSV = N + 1;
  for (int i = 0; i < N; i++) { aSV[i] = RVN_SV_CONV_STOR; aLev[i] = conv_index * N + i; }
  aSV[N] = RVN_SV_CONVOLUTION; aLev[N] = conv_index;
}
double CmvConvolution::LocalCumulDist(double t, const surface_struct &S) const {
  if (_type == RVN_CONVOL_GR4J_1) {
    double x4 = S.v[RVN_LU_GR4J_X4];
    return rvh_min(pow(t / x4, 2.5), 1.0);
  } else if (_type == RVN_CONVOL_GR4J_2) {
    double x4 = S.v[RVN_LU_GR4J_X4];
    if (t / x4 < 1.0) return 0.5 * pow(t / x4, 2.5);
    else if (t / x4 < 2.0) return 1.0 - 0.5 * pow(2 - t / x4, 2.5);
    else return 1.0;
  } else if (_type == RVN_CONVOL_GAMMA) {
    return GammaCumDist(t, S.v[RVN_LU_GAMMA_SHAPE], S.v[RVN_LU_GAMMA_SCALE]);
  } else if (_type == RVN_CONVOL_GAMMA_2) {
    return GammaCumDist(t, S.v[RVN_LU_GAMMA_SHAPE2], S.v[RVN_LU_GAMMA_SCALE2]);
  }
  return 1.0;
}
void CmvConvolution::GenerateUnitHydrograph(const surface_struct &S, const optStruct &Options, double *aUnitHydro, int &N) const {
  double tstep = Options.timestep, max_time = 0;
  if (_type == RVN_CONVOL_GR4J_1) max_time = S.v[RVN_LU_GR4J_X4];
  else if (_type == RVN_CONVOL_GR4J_2) max_time = 2 * S.v[RVN_LU_GR4J_X4];
  else if (_type == RVN_CONVOL_GAMMA) max_time = 4.5 * pow(S.v[RVN_LU_GAMMA_SHAPE], 0.6) / S.v[RVN_LU_GAMMA_SCALE];
  else if (_type == RVN_CONVOL_GAMMA_2) max_time = 4.5 * pow(S.v[RVN_LU_GAMMA_SHAPE2], 0.6) / S.v[RVN_LU_GAMMA_SCALE2];
  ExitGracefullyIf(!(max_time > 0.0), "CmvConvolution::GenerateUnitHydrograph: negative or zero duration of unit hydrograph (bad GR4J_x4, or GAMMA_SHAPE/GAMMA_SCALE)");
  double sum = 0;
  max_time = rvh_min(MAX_CONVOL_STORES * tstep, max_time);
  N = (int)(ceil(max_time / tstep));
  if (N == 0) N = 1;
  for (int n = 0; n < N; n++) {
    aUnitHydro[n] = LocalCumulDist((n + 1) * tstep, S) - sum;
    sum += aUnitHydro[n];
  }
  ExitGracefullyIf(sum == 0.0, "CmvConvolution::GenerateUnitHydrograph: bad unit hydrograph constructed");
  for (int n = 0; n < N; n++) aUnitHydro[n] /= sum;
  for (int n = N; n < MAX_CONVOL_STORES; n++) aUnitHydro[n] = 0.0;
}

// ---- factory -------------------------------------------------------------------------------------------
static void AddSV1(CModel *pM, sv_type t, int lev = DOESNT_EXIST) { pM->AddStateVariables(&t, &lev, 1); }
static void AddSVs(CModel *pM, const std::vector<std::pair<sv_type, int> > &l) {
  for (size_t i = 0; i < l.size(); i++) AddSV1(pM, l[i].first, l[i].second);
}
typedef std::pair<sv_type, int> SVL;
static int SoilLayerOf(CModel *pM, int i) { return (pM->GetStateVarType(i) == RVN_SV_SOIL) ? pM->GetStateVarLayer(i) : DOESNT_EXIST; }

CHydroProcessABC *CreateProcessFromRVI(const std::vector<std::string> &s, CModel *pM, const optStruct &Options) {
  (void)Options;
  const std::string &cmd = s[0];
  const int Len = (int)s.size();
  const int iAtmos = pM->GetStateVarIndex(RVN_SV_ATMOSPHERE);
  const int iSW = pM->GetStateVarIndex(RVN_SV_SURFACE_WATER);
  const int iPond = pM->GetStateVarIndex(RVN_SV_PONDED_WATER);
  auto need = [&](int n) { ExitGracefullyIf(Len < n, "ParseMainInputFile: improper format of " + cmd + " command"); };

  if (cmd == ":Precipitation") {  // src/ParseInput.cpp case 212
    need(4);
    AddSV1(pM, RVN_SV_ATMOS_PRECIP);
    return new CmvPrecipitation(pM);
  }
  if (cmd == ":SnowBalance") {  // case 209; connections src/SnowBalance.cpp:31-176
    need(4);
    CmvGeneric *p = NULL;
    if (s[1] == "SNOBAL_HMETS") {
      AddSVs(pM, {SVL(RVN_SV_SNOW, -1), SVL(RVN_SV_SNOW_LIQ, -1), SVL(RVN_SV_PONDED_WATER, -1), SVL(RVN_SV_CUM_SNOWMELT, -1)});
      int iSnow = pM->GetStateVarIndex(RVN_SV_SNOW), iSL = pM->GetStateVarIndex(RVN_SV_SNOW_LIQ), iCM = pM->GetStateVarIndex(RVN_SV_CUM_SNOWMELT);
      p = new CmvGeneric("SnowBalance", RVN_OP_SNOBAL_HMETS, pM);
      p->Connect(iSnow, iSL); p->Connect(iSnow, iPond); p->Connect(iSL, iPond); p->Connect(iSL, iSnow); p->Connect(iCM, iCM);
    } else if (s[1] == "SNOBAL_SIMPLE_MELT") {
      AddSV1(pM, RVN_SV_SNOW);
      int to = pM->ParseSVTypeIndex(s[3]);
      p = new CmvGeneric("SnowBalance", RVN_OP_SNOBAL_SIMPLE_MELT, pM);
      p->Connect(pM->GetStateVarIndex(RVN_SV_SNOW), to);
    } else if (s[1] == "SNOBAL_HBV") {   // src/SnowBalance.cpp:56-70, 296-303
      AddSVs(pM, {SVL(RVN_SV_SNOW, -1), SVL(RVN_SV_SNOW_LIQ, -1), SVL(RVN_SV_SNOW_DEPTH, -1)});
      int iSnow = pM->GetStateVarIndex(RVN_SV_SNOW), iSL = pM->GetStateVarIndex(RVN_SV_SNOW_LIQ), iSoil = pM->GetStateVarIndex(RVN_SV_SOIL, 0);
      p = new CmvGeneric("SnowBalance", RVN_OP_SNOBAL_HBV, pM);
      p->Connect(iSnow, iSL); p->Connect(iSnow, iSoil); p->Connect(iSL, iSoil); p->Connect(iSnow, iPond); p->Connect(iSL, iPond);
    } else if (s[1] == "SNOBAL_GAWSER") {   // src/SnowBalance.cpp:123-145 (connections), :336-344 (state variables, in this order)
      AddSVs(pM, {SVL(RVN_SV_NEW_SNOW, -1), SVL(RVN_SV_SNOW, -1), SVL(RVN_SV_SNOW_LIQ, -1), SVL(RVN_SV_SNOW_DEPTH, -1), SVL(RVN_SV_PONDED_WATER, -1)});
      const int iNew = pM->GetStateVarIndex(RVN_SV_NEW_SNOW), iSWC = pM->GetStateVarIndex(RVN_SV_SNOW), iSD = pM->GetStateVarIndex(RVN_SV_SNOW_DEPTH);
      const int iLWC = pM->GetStateVarIndex(RVN_SV_SNOW_LIQ);
      p = new CmvGeneric("SnowBalance", RVN_OP_SNOBAL_GAWSER, pM);
      p->Connect(iSWC, iLWC); p->Connect(iPond, iLWC); p->Connect(iLWC, iSWC); p->Connect(iSD, iSD); p->Connect(iSD, iSD); p->Connect(iSD, iSD);
      p->Connect(iNew, iSWC); p->Connect(iSWC, iAtmos); p->Connect(iLWC, iPond); p->Connect(iSWC, iPond);
    } else if (s[1] == "SNOBAL_CRHM_EBSM") {   // src/SnowBalance.cpp:146-158 (connections), :345-352 (state variables)
      AddSVs(pM, {SVL(RVN_SV_SNOW, -1), SVL(RVN_SV_SNOW_LIQ, -1), SVL(RVN_SV_PONDED_WATER, -1), SVL(RVN_SV_COLD_CONTENT, -1)});
      const int iSnow = pM->GetStateVarIndex(RVN_SV_SNOW), iSL = pM->GetStateVarIndex(RVN_SV_SNOW_LIQ), iCC = pM->GetStateVarIndex(RVN_SV_COLD_CONTENT);
      p = new CmvGeneric("SnowBalance", RVN_OP_SNOBAL_CRHM_EBSM, pM);
      p->Connect(iSnow, iSL); p->Connect(iSnow, iPond); p->Connect(iSL, iPond); p->Connect(iSL, iSnow); p->Connect(iCC, iCC);
    } else if (s[1] == "SNOBAL_COLD_CONTENT") {   // src/SnowBalance.cpp:41-55 (connections), :287-295 (state variables, in this order)
      AddSVs(pM, {SVL(RVN_SV_SNOW_LIQ, -1), SVL(RVN_SV_COLD_CONTENT, -1), SVL(RVN_SV_ENERGY_LOSSES, -1), SVL(RVN_SV_SURFACE_WATER, -1), SVL(RVN_SV_SNOW, -1)});
      const int iSnow = pM->GetStateVarIndex(RVN_SV_SNOW), iSL = pM->GetStateVarIndex(RVN_SV_SNOW_LIQ);
      const int iCC = pM->GetStateVarIndex(RVN_SV_COLD_CONTENT), iEn = pM->GetStateVarIndex(RVN_SV_ENERGY_LOSSES);
      p = new CmvGeneric("SnowBalance", RVN_OP_SNOBAL_COLD_CONTENT, pM);
      p->Connect(iCC, iEn); p->Connect(iSL, iSnow); p->Connect(iEn, iCC); p->Connect(iSnow, iPond); p->Connect(iSL, iPond);
    } else if (s[1] == "SNOBAL_CEMA_NIEGE" || s[1] == "SNOBAL_CEMA_NEIGE") {
      AddSVs(pM, {SVL(RVN_SV_SNOW, -1), SVL(RVN_SV_PONDED_WATER, -1), SVL(RVN_SV_SNOW_COVER, -1)});
      int iSnow = pM->GetStateVarIndex(RVN_SV_SNOW), iSC = pM->GetStateVarIndex(RVN_SV_SNOW_COVER);
      p = new CmvGeneric("SnowBalance", RVN_OP_SNOBAL_CEMA_NEIGE, pM);
      p->Connect(iSnow, iPond); p->Connect(iSC, iSC);
    } else {
      ExitGracefully("ParseMainInputFile: snow balance algorithm " + s[1] + " is not implemented in the B200 build");
    }
    return p;
  }
  if (cmd == ":Infiltration") {  // case 204; src/Infiltration.cpp:19-66
    need(4);
    AddSVs(pM, {SVL(RVN_SV_PONDED_WATER, -1), SVL(RVN_SV_SURFACE_WATER, -1)});
    int iTop = pM->GetStateVarIndex(RVN_SV_SOIL, 0);
    CmvGeneric *p = NULL;
    if (s[1] == "INF_HMETS") {
      AddSVs(pM, {SVL(RVN_SV_CONVOLUTION, 0), SVL(RVN_SV_CONVOLUTION, 1)});
      p = new CmvGeneric("Infiltration", RVN_OP_INF_HMETS, pM);
      p->Connect(iPond, iTop); p->Connect(iPond, iSW);
      p->Connect(iPond, pM->GetStateVarIndex(RVN_SV_CONVOLUTION, 0)); p->Connect(iPond, pM->GetStateVarIndex(RVN_SV_CONVOLUTION, 1));
    } else if (s[1] == "INF_HBV" || s[1] == "INF_GR4J" || s[1] == "INF_VIC_ARNO") {
      p = new CmvGeneric("Infiltration", s[1] == "INF_HBV" ? RVN_OP_INF_HBV : (s[1] == "INF_GR4J" ? RVN_OP_INF_GR4J : RVN_OP_INF_VIC_ARNO), pM);
      p->Connect(iPond, iTop); p->Connect(iPond, iSW);
    } else if (s[1] == "INF_RATIONAL" || s[1] == "INF_PARTITION" || s[1] == "INF_ALL_INFILTRATES" || s[1] == "INF_VIC" || s[1] == "INF_PRMS" || s[1] == "INF_PDM") {
      const rvn_opcode op = (s[1] == "INF_ALL_INFILTRATES") ? RVN_OP_INF_ALL_INFILTRATES : (s[1] == "INF_VIC") ? RVN_OP_INF_VIC
                          : (s[1] == "INF_PRMS") ? RVN_OP_INF_PRMS : (s[1] == "INF_PDM") ? RVN_OP_INF_PDM : RVN_OP_INF_RATIONAL;
      p = new CmvGeneric("Infiltration", op, pM);
      p->Connect(iPond, iTop); p->Connect(iPond, iSW);
    } else if (s[1] == "INF_GREEN_AMPT") {   // src/Infiltration.cpp:20-23 (the two default connections), src/GreenAmpt.cpp
      p = new CmvGeneric("Infiltration", RVN_OP_INF_GREEN_AMPT, pM);
      p->Connect(iPond, iTop); p->Connect(iPond, iSW);
    } else if (s[1] == "INF_SCS" || s[1] == "INF_SCS_NOABSTRACTION") {   // src/Infiltration.cpp:20-23, GetSCSRunoff :638-680
      p = new CmvGeneric("Infiltration", s[1] == "INF_SCS" ? RVN_OP_INF_SCS : RVN_OP_INF_SCS_NOABSTRACTION, pM);
      p->Connect(iPond, iTop); p->Connect(iPond, iSW);
    } else if (s[1] == "INF_UBC") {   // src/Infiltration.cpp:34-43,250-256: topsoil, interflow, upper and lower groundwater stores
      ExitGracefullyIf(pM->GetNumSoilLayers() < 4, "INF_UBCWM infiltration algorithm requires at least 4 soil layers to operate. Please use a different :SoilModel or replace this infiltration algorithm.");
      AddSVs(pM, {SVL(RVN_SV_SOIL, 1), SVL(RVN_SV_SOIL, 2), SVL(RVN_SV_SOIL, 3)});
      p = new CmvGeneric("Infiltration", RVN_OP_INF_UBC, pM);
      p->Connect(iPond, iTop); p->Connect(iPond, iSW);
      for (int m = 1; m <= 3; m++) p->Connect(iPond, pM->GetStateVarIndex(RVN_SV_SOIL, m));
    } else if (s[1] == "INF_GA_SIMPLE") {   // src/Infiltration.cpp:25-33,244-249: tracks the cumulative infiltration of the running rainfall event
      AddSVs(pM, {SVL(RVN_SV_CUM_INFIL, -1), SVL(RVN_SV_GA_MOISTURE_INIT, -1)});
      p = new CmvGeneric("Infiltration", RVN_OP_INF_GA_SIMPLE, pM);
      p->Connect(iPond, iTop); p->Connect(iPond, iSW);
      const int iCum = pM->GetStateVarIndex(RVN_SV_CUM_INFIL), iGA = pM->GetStateVarIndex(RVN_SV_GA_MOISTURE_INIT);
      p->Connect(iCum, iCum); p->Connect(iGA, iGA);
    } else {
      ExitGracefully("ParseMainInputFile: infiltration algorithm " + s[1] + " is not implemented in the B200 build");
    }
    return p;
  }
  if (cmd == ":Overflow" || cmd == ":-->Overflow") {  // case 222
    need(4);
    int from = pM->ParseSVTypeIndex(s[2]), to = pM->ParseSVTypeIndex(s[3]);
    CmvGeneric *p = new CmvGeneric("Overflow", RVN_OP_OVERFLOW, pM);
    p->Connect(from, to);
    p->SetIntArg(0, SoilLayerOf(pM, from));
    return p;
  }
  if (cmd == ":DepressionOverflow") {  // case 230; src/DepressionProcesses.cpp:20-33,90-98
    need(4);
    int type = -1;
    if (s[1] == "DFLOW_THRESHPOW") type = 0; else if (s[1] == "DFLOW_LINEAR") type = 1; else if (s[1] == "DFLOW_WEIR") type = 2;
    ExitGracefullyIf(type < 0, "ParseMainInputFile: Unrecognized depression overflow algorithm");
    AddSVs(pM, {SVL(RVN_SV_DEPRESSION, -1), SVL(RVN_SV_SURFACE_WATER, -1)});
    CmvGeneric *p = new CmvGeneric("DepressionOverflow", RVN_OP_DEPRESSION_OVERFLOW, pM);
    p->Connect(pM->GetStateVarIndex(RVN_SV_DEPRESSION), iSW);
    p->SetIntArg(0, type);
    return p;
  }
  if (cmd == ":Seepage") {  // case 233; src/DepressionProcesses.cpp:180-214
    need(4);
    ExitGracefullyIf(s[1] != "SEEP_LINEAR", "ParseMainInputFile: Unrecognized seepage algorithm");
    AddSV1(pM, RVN_SV_DEPRESSION);
    const int to = pM->ParseSVTypeIndex(s[3]);
    ExitGracefullyIf(pM->GetStateVarType(to) != RVN_SV_SOIL && pM->GetStateVarType(to) != RVN_SV_GROUNDWATER, "CmvSeepage::Initialize: seepage must move water to soil or groundwater unit");
    CmvGeneric *p = new CmvGeneric("Seepage", RVN_OP_SEEPAGE, pM);
    p->Connect(pM->GetStateVarIndex(RVN_SV_DEPRESSION), to);
    p->SetIntArg(0, SoilLayerOf(pM, to));
    return p;
  }
  if (cmd == ":ExchangeFlow") {  // case 231: :ExchangeFlow RAVEN_DEFAULT [from] [mixing zone]; src/Flush.cpp:296-326
    need(4);
    int from = pM->ParseSVTypeIndex(s[2]), to = pM->ParseSVTypeIndex(s[3]);
    ExitGracefullyIf(SoilLayerOf(pM, from) < 0 || SoilLayerOf(pM, to) < 0, "CmvExchangeFlow::Initialize: exchange flow  must be between two soil units");
    CmvGeneric *p = new CmvGeneric("ExchangeFlow", RVN_OP_EXCHANGE_FLOW, pM);
    p->Connect(from, to); p->Connect(to, from);
    p->SetIntArg(0, SoilLayerOf(pM, from));
    return p;
  }
  if (cmd == ":Flush") {  // case 221
    need(4);
    int from = pM->ParseSVTypeIndex(s[2]), to = pM->ParseSVTypeIndex(s[3]);
    CmvGeneric *p = new CmvGeneric("Flush", RVN_OP_FLUSH, pM);
    p->Connect(from, to);
    p->SetDblArg(0, (Len >= 5) ? s_to_d(s[4]) : 1.0);
    return p;
  }
  if (cmd == ":Split") {  // case 227
    need(6);
    int from = pM->ParseSVTypeIndex(s[2]), to1 = pM->ParseSVTypeIndex(s[3]), to2 = pM->ParseSVTypeIndex(s[4]);
    CmvGeneric *p = new CmvGeneric("Split", RVN_OP_SPLIT, pM);
    p->Connect(from, to1); p->Connect(from, to2);
    p->SetDblArg(0, s_to_d(s[5]));
    return p;
  }
  if (cmd == ":Baseflow") {  // case 201
    need(4);
    AddSV1(pM, RVN_SV_SURFACE_WATER);
    int from = pM->ParseSVTypeIndex(s[2]);
    rvn_opcode op = RVN_OP_NOP;
    if (s[1] == "BASE_LINEAR") op = RVN_OP_BASE_LINEAR;
    else if (s[1] == "BASE_POWER_LAW") op = RVN_OP_BASE_POWER_LAW;
    else if (s[1] == "BASE_GR4J") op = RVN_OP_BASE_GR4J;
    else if (s[1] == "BASE_LINEAR_ANALYTIC") op = RVN_OP_BASE_LINEAR_ANALYTIC;
    else if (s[1] == "BASE_VIC") op = RVN_OP_BASE_VIC;
    else if (s[1] == "BASE_TOPMODEL") op = RVN_OP_BASE_TOPMODEL;
    else if (s[1] == "BASE_LINEAR_CONSTRAIN") op = RVN_OP_BASE_LINEAR_CONSTRAIN;
    else if (s[1] == "BASE_CONSTANT") op = RVN_OP_BASE_CONSTANT;
    else if (s[1] == "BASE_THRESH_POWER") op = RVN_OP_BASE_THRESH_POWER;
    else if (s[1] == "BASE_THRESH_STOR") op = RVN_OP_BASE_THRESH_STOR;
    else ExitGracefully("ParseMainInputFile: baseflow algorithm " + s[1] + " is not implemented in the B200 build");
    CmvGeneric *p = new CmvGeneric("Baseflow", op, pM);
    p->Connect(from, iSW);
    p->SetIntArg(0, SoilLayerOf(pM, from));
    return p;
  }
  if (cmd == ":Percolation") {  // case 205
    need(4);
    int from = pM->ParseSVTypeIndex(s[2]), to = pM->ParseSVTypeIndex(s[3]);
    rvn_opcode op = RVN_OP_NOP;
    if (s[1] == "PERC_LINEAR") op = RVN_OP_PERC_LINEAR;
    else if (s[1] == "PERC_CONSTANT") op = RVN_OP_PERC_CONSTANT;
    else if (s[1] == "PERC_GR4J") op = RVN_OP_PERC_GR4J;
    else if (s[1] == "PERC_GR4JEXCH") op = RVN_OP_PERC_GR4JEXCH;
    else if (s[1] == "PERC_GR4JEXCH2") op = RVN_OP_PERC_GR4JEXCH2;
    else if (s[1] == "PERC_GAWSER") op = RVN_OP_PERC_GAWSER;
    else if (s[1] == "PERC_GAWSER_CONSTRAIN") op = RVN_OP_PERC_GAWSER_CONSTRAIN;
    else if (s[1] == "PERC_POWER_LAW") op = RVN_OP_PERC_POWER_LAW;
    else if (s[1] == "PERC_LINEAR_ANALYTIC") op = RVN_OP_PERC_LINEAR_ANALYTIC;
    else if (s[1] == "PERC_PRMS") op = RVN_OP_PERC_PRMS;
    else if (s[1] == "PERC_SACRAMENTO") op = RVN_OP_PERC_SACRAMENTO;
    else if (s[1] == "PERC_ASPEN") op = RVN_OP_PERC_ASPEN;
    else ExitGracefully("ParseMainInputFile: percolation algorithm " + s[1] + " is not implemented in the B200 build");
    CmvGeneric *p = new CmvGeneric("Percolation", op, pM);
    p->Connect(from, to);
    p->SetIntArg(0, SoilLayerOf(pM, from));
    p->SetIntArg(1, SoilLayerOf(pM, to));
    return p;
  }
  if (cmd == ":SoilEvaporation") {  // case 208; src/SoilEvaporation.cpp:21-100,244-313
    need(4);
    rvn_opcode op = RVN_OP_NOP;
    if (s[1] == "SOILEVAP_ALL") op = RVN_OP_SOILEVAP_ALL;
    else if (s[1] == "SOILEVAP_HBV") op = RVN_OP_SOILEVAP_HBV;
    else if (s[1] == "SOILEVAP_GR4J") op = RVN_OP_SOILEVAP_GR4J;
    else if (s[1] == "SOILEVAP_TOPMODEL") op = RVN_OP_SOILEVAP_TOPMODEL;
    else if (s[1] == "SOILEVAP_LINEAR") op = RVN_OP_SOILEVAP_LINEAR;
    else if (s[1] == "SOILEVAP_PDM") op = RVN_OP_SOILEVAP_PDM;
    else if (s[1] == "SOILEVAP_HYMOD2") op = RVN_OP_SOILEVAP_HYMOD2;
    else if (s[1] == "SOILEVAP_UBC") op = RVN_OP_SOILEVAP_UBC;
    else if (s[1] == "SOILEVAP_VIC") op = RVN_OP_SOILEVAP_VIC;
    else if (s[1] == "SOILEVAP_HYPR") op = RVN_OP_SOILEVAP_HYPR;
    else if (s[1] == "SOILEVAP_SEQUEN") op = RVN_OP_SOILEVAP_SEQUEN;
    else if (s[1] == "SOILEVAP_ROOT") op = RVN_OP_SOILEVAP_ROOT;
    else if (s[1] == "SOILEVAP_ROOT_CONSTRAIN") op = RVN_OP_SOILEVAP_ROOT_CONSTRAIN;
    else if (s[1] == "SOILEVAP_AWBM") op = RVN_OP_SOILEVAP_AWBM;
    else if (s[1] == "SOILEVAP_SACSMA") op = RVN_OP_SOILEVAP_SACSMA;
    else if (s[1] == "SOILEVAP_CHU") op = RVN_OP_SOILEVAP_CHU;
    else ExitGracefully("ParseMainInputFile: soil evaporation algorithm " + s[1] + " is not implemented in the B200 build");
    // state variables and connections: src/SoilEvaporation.cpp:21-100 and :244-313
    if (op == RVN_OP_SOILEVAP_SACSMA) {   // :86-97, :294-301: UZT, UZF, LZT, LZFP, LZFS, ADIMC = SOIL[0..5]
      ExitGracefullyIf(pM->GetNumSoilLayers() < 6, "SOILEVAP_SACSMA needs the six Sacramento soil stores (:SoilModel SOIL_MULTILAYER 6)");
      for (int m = 0; m <= 5; m++) AddSV1(pM, RVN_SV_SOIL, m);
      AddSV1(pM, RVN_SV_ATMOSPHERE); AddSV1(pM, RVN_SV_AET);
      CmvGeneric *p = new CmvGeneric("SoilEvaporation", op, pM);
      const int iAET = pM->GetStateVarIndex(RVN_SV_AET);
      auto S = [&](int m) { return pM->GetStateVarIndex(RVN_SV_SOIL, m); };
      p->Connect(S(0), iAtmos); p->Connect(S(1), iAtmos); p->Connect(S(1), S(0)); p->Connect(S(2), iAtmos); p->Connect(S(4), S(2)); p->Connect(S(5), iAtmos);
      p->Connect(iAET, iAET);
      return p;
    }
    const bool two = (op == RVN_OP_SOILEVAP_SEQUEN || op == RVN_OP_SOILEVAP_ROOT || op == RVN_OP_SOILEVAP_ROOT_CONSTRAIN);
    AddSV1(pM, RVN_SV_SOIL, 0);
    if (two || op == RVN_OP_SOILEVAP_AWBM) AddSV1(pM, RVN_SV_SOIL, 1);
    if (op == RVN_OP_SOILEVAP_AWBM) AddSV1(pM, RVN_SV_SOIL, 2);
    AddSV1(pM, RVN_SV_ATMOSPHERE);
    if (op == RVN_OP_SOILEVAP_HYPR) AddSV1(pM, RVN_SV_DEPRESSION);
    if (op == RVN_OP_SOILEVAP_UBC) AddSV1(pM, RVN_SV_SNOW);
    if (op == RVN_OP_SOILEVAP_CHU) AddSV1(pM, RVN_SV_CROP_HEAT_UNITS);   // src/SoilEvaporation.cpp:271-276
    AddSV1(pM, RVN_SV_AET);
    CmvGeneric *p = new CmvGeneric("SoilEvaporation", op, pM);
    int iAET = pM->GetStateVarIndex(RVN_SV_AET);
    p->Connect(pM->GetStateVarIndex(RVN_SV_SOIL, 0), iAtmos);
    if (two || op == RVN_OP_SOILEVAP_AWBM) p->Connect(pM->GetStateVarIndex(RVN_SV_SOIL, 1), iAtmos);
    if (op == RVN_OP_SOILEVAP_AWBM) p->Connect(pM->GetStateVarIndex(RVN_SV_SOIL, 2), iAtmos);
    if (op == RVN_OP_SOILEVAP_HYPR) p->Connect(pM->GetStateVarIndex(RVN_SV_DEPRESSION), iAtmos);
    p->Connect(iAET, iAET);
    return p;
  }
  if (cmd == ":LateralFlush") {  // case 232: :LateralFlush RAVEN_DEFAULT [FROM_HRUGroup] [FROM_SV] To [TO_HRUGROUP] [TO_SV] {INTERBASIN}
    need(7);
    const int from = pM->ParseSVTypeIndex(s[3]), to = pM->ParseSVTypeIndex(s[6]);
    const bool interbasin = (Len >= 8) && (s[7] == "INTERBASIN");
    const int gF = pM->FindHRUGroup(s[2]), gT = pM->FindHRUGroup(s[5]);
    ExitGracefullyIf(gF == DOESNT_EXIST || gT == DOESNT_EXIST, "ParseInput: Lateral Flush - invalid 'to' or 'from' HRU Group used. Must define using :DefineHRUGroups command.");
    return new CmvLatFlush(from, to, gF, gT, !interbasin, pM);
  }
  if (cmd == ":LateralEquilibrate") {  // case 238: :LateralEquilibrate RAVEN_DEFAULT [HRUGroup] [SV] [mixing_rate] {INTERBASIN}
    need(5);
    const int sv = pM->ParseSVTypeIndex(s[3]);
    const bool interbasin = (Len >= 6) && (s[5] == "INTERBASIN");
    const int g = pM->FindHRUGroup(s[2]);
    ExitGracefullyIf(g == DOESNT_EXIST, "ParseInput: Lateral Equilibrate - invalid HRU Group used. Must define using :DefineHRUGroups command.");
    return new CmvLatEquilibrate(sv, g, s_to_d(s[4]), !interbasin, pM);
  }
  if (cmd == ":SnowRedistribute") {  // case 241: :SnowRedistribute [method] [SV] [max_snow_height]
    need(4);
    int method = RVN_REDIST_CONTINUOUS;
    if (s[1] == "CONTINUOUS") method = RVN_REDIST_CONTINUOUS;
    else if (s[1] == "THRESHOLD") method = RVN_REDIST_THRESHOLD;
    else ExitGracefully("ParseMainInputFile: Unrecognized snow redistribution method");
    return new CmvLatRedistribute(pM->ParseSVTypeIndex(s[2]), s_to_d(s[3]), method, pM);
  }
  if (cmd == ":Interflow") {  // case 210; src/Interflow.cpp:19-45,78-87
    need(4);
    ExitGracefullyIf(s[1] != "INTERFLOW_PRMS" && s[1] != "PRMS", "ParseMainInputFile: Unrecognized interflow process representation");
    AddSVs(pM, {SVL(RVN_SV_SOIL, 0), SVL(RVN_SV_SURFACE_WATER, -1)});
    int from = pM->ParseSVTypeIndex(s[2]);
    ExitGracefullyIf(pM->GetStateVarType(from) != RVN_SV_SOIL, "CmvInterflow Constructor: invalid 'from' compartment specified");
    CmvGeneric *p = new CmvGeneric("Interflow", RVN_OP_INTERFLOW_PRMS, pM);
    p->Connect(from, iSW);
    p->SetIntArg(0, SoilLayerOf(pM, from));
    return p;
  }
  if (cmd == ":SnowMelt") {  // case 213; src/SnowMeltRefreeze.cpp:22-75
    need(4);
    ExitGracefullyIf(s[1] != "MELT_POTMELT", "ParseMainInputFile: Unrecognized snow melt process representation");
    AddSV1(pM, RVN_SV_SNOW);
    int to = pM->ParseSVTypeIndex(s[3]);
    CmvGeneric *p = new CmvGeneric("SnowMelt", RVN_OP_SNOWMELT_POTMELT, pM);
    p->Connect(pM->GetStateVarIndex(RVN_SV_SNOW), to);
    return p;
  }
  if (cmd == ":SnowSqueeze") {  // case 218; src/SnowMeltRefreeze.cpp:137-166
    need(4);
    ExitGracefullyIf(s[1] != "SQUEEZE_RAVEN", "ParseMainInputFile: Unrecognized snow squeeze process representation");
    AddSV1(pM, RVN_SV_SNOW_LIQ);
    int to = pM->ParseSVTypeIndex(s[3]);
    ExitGracefullyIf(pM->StateVarExists(RVN_SV_SNOW_DEPTH), "CmvSnowSqueeze: models with a SNOW_DEPTH state variable are not supported by the B200 build");
    CmvGeneric *p = new CmvGeneric("SnowSqueeze", RVN_OP_SNOWSQUEEZE, pM);
    p->Connect(pM->GetStateVarIndex(RVN_SV_SNOW_LIQ), to);
    return p;
  }
  if (cmd == ":SnowRefreeze") {  // case 214; src/SnowMeltRefreeze.cpp:253-267
    need(4);
    ExitGracefullyIf(s[1] != "FREEZE_DEGREE_DAY", "ParseMainInputFile: Unrecognized snow refreeze process representation");
    AddSVs(pM, {SVL(RVN_SV_SNOW_LIQ, -1), SVL(RVN_SV_SNOW, -1)});
    CmvGeneric *p = new CmvGeneric("SnowRefreeze", RVN_OP_SNOWREFREEZE_DD, pM);
    p->Connect(pM->GetStateVarIndex(RVN_SV_SNOW_LIQ), pM->GetStateVarIndex(RVN_SV_SNOW));
    return p;
  }
  if (cmd == ":SnowTempEvolve") {  // case 229; src/SnowTempEvolve.cpp:14-22
    need(3);
    ExitGracefullyIf(s[1] != "SNOTEMP_NEWTONS", "ParseMainInputFile: Unrecognized snow temp evolution process representation");
    AddSV1(pM, RVN_SV_SNOW_TEMP);
    CmvGeneric *p = new CmvGeneric("SnowTempEvolve", RVN_OP_SNOTEMP_NEWTONS, pM);
    p->Connect(pM->GetStateVarIndex(RVN_SV_SNOW_TEMP), pM->GetStateVarIndex(RVN_SV_SNOW_TEMP));
    return p;
  }
  if (cmd == ":CanopyEvaporation") {  // case 202; src/VegetationMovers.cpp:20-32
    need(4);
    int op = -1;
    if (s[1] == "CANEVP_ALL") op = RVN_OP_CANEVP_ALL;
    else if (s[1] == "CANEVP_MAXIMUM") op = RVN_OP_CANEVP_MAXIMUM;
    else if (s[1] == "CANEVP_RUTTER") op = RVN_OP_CANEVP_RUTTER;   // trunk fraction is 0 without a TRUNK variable, which nothing built here creates
    ExitGracefullyIf(op < 0, "ParseMainInputFile: canopy evaporation algorithm " + s[1] + " is not implemented in the B200 build");
    AddSVs(pM, {SVL(RVN_SV_CANOPY, -1), SVL(RVN_SV_ATMOSPHERE, -1), SVL(RVN_SV_AET, -1)});
    CmvGeneric *p = new CmvGeneric("CanopyEvaporation", (rvn_opcode)op, pM);
    int iAET = pM->GetStateVarIndex(RVN_SV_AET);
    p->Connect(pM->GetStateVarIndex(RVN_SV_CANOPY), iAtmos); p->Connect(iAET, iAET);
    return p;
  }
  if (cmd == ":CanopySnowEvap" || cmd == ":CanopySnowEvaporation" || cmd == ":CanopySublimation") {  // case 221; src/VegetationMovers.cpp:205-216
    need(4);
    const bool snow_max = (s[1] == "CANEVP_MAXIMUM" || s[1] == "SUBLIM_MAXIMUM");
    ExitGracefullyIf(s[1] != "CANEVP_ALL" && s[1] != "SUBLIM_ALL" && !snow_max, "ParseMainInputFile: canopy sublimation algorithm " + s[1] + " is not implemented in the B200 build");
    AddSVs(pM, {SVL(RVN_SV_CANOPY_SNOW, -1), SVL(RVN_SV_ATMOSPHERE, -1), SVL(RVN_SV_AET, -1)});
    CmvGeneric *p = new CmvGeneric("CanopySublimation", snow_max ? RVN_OP_CANSNOWEVP_MAXIMUM : RVN_OP_CANSNOWEVP_ALL, pM);
    int iAET = pM->GetStateVarIndex(RVN_SV_AET);
    p->Connect(pM->GetStateVarIndex(RVN_SV_CANOPY_SNOW), iAtmos); p->Connect(iAET, iAET);
    return p;
  }
  if (cmd == ":CropHeatUnitEvolve") {  // case 224; src/CropGrowth.cpp:15-33,76-84
    need(2);
    ExitGracefullyIf(s[1] != "CHU_ONTARIO", "ParseMainInputFile: Unrecognized crop heat evolution algorithm");
    AddSV1(pM, RVN_SV_CROP_HEAT_UNITS);
    CmvGeneric *p = new CmvGeneric("CropHeatUnitEvolve", RVN_OP_CHU_ONTARIO, pM);
    const int iCHU = pM->GetStateVarIndex(RVN_SV_CROP_HEAT_UNITS);
    p->Connect(iCHU, iCHU);
    return p;
  }
  if (cmd == ":Abstraction") {  // case 225; src/Abstraction.cpp:20-33
    need(4);
    int type = -1;
    if (s[1] == "ABST_PERCENTAGE") type = RVN_ABST_PERCENTAGE;
    else if (s[1] == "ABST_FILL") type = RVN_ABST_FILL;
    else if (s[1] == "ABST_SCS") type = RVN_ABST_SCS;
    else if (s[1] == "ABST_PDMROF") type = RVN_ABST_PDMROF;
    ExitGracefullyIf(type < 0, "ParseMainInputFile: abstraction algorithm " + s[1] + " is not implemented in the B200 build");
    ExitGracefullyIf(s[2] != "PONDED_WATER" || s[3] != "DEPRESSION", "ParseMainInputFile: :Abstraction moves water from PONDED_WATER to DEPRESSION");
    AddSVs(pM, {SVL(RVN_SV_PONDED_WATER, -1), SVL(RVN_SV_DEPRESSION, -1)});
    CmvGeneric *p = new CmvGeneric("Abstraction", RVN_OP_ABSTRACTION, pM);
    p->Connect(pM->GetStateVarIndex(RVN_SV_PONDED_WATER), pM->GetStateVarIndex(RVN_SV_DEPRESSION));
    if (type == RVN_ABST_PDMROF) p->Connect(pM->GetStateVarIndex(RVN_SV_PONDED_WATER), iSW);   // runoff of what the depressions shed (src/Abstraction.cpp:33-41)
    p->SetIntArg(0, type);
    return p;
  }
  if (cmd == ":Sublimation") {  // case 210; src/Sublimation.cpp:20-44 (the from/to tokens must be SNOW ATMOSPHERE)
    need(4);
    int type = -1;
    if (s[1] == "SUBLIM_SVERDRUP") type = RVN_SUBLIM_SVERDRUP;
    else if (s[1] == "SUBLIM_KUZMIN") type = RVN_SUBLIM_KUZMIN;
    else if (s[1] == "SUBLIM_KUCHMENT_GELFAN") type = RVN_SUBLIM_KUCHMENT_GELFAN;
    else if (s[1] == "SUBLIM_BULK_AERO") type = RVN_SUBLIM_BULK_AERO;
    else if (s[1] == "SUBLIM_CENTRAL_SIERRA") type = RVN_SUBLIM_CENTRAL_SIERRA;
    ExitGracefullyIf(type < 0, "ParseMainInputFile: sublimation algorithm " + s[1] + " is not implemented in the B200 build");
    ExitGracefullyIf(s[2] != "SNOW" || s[3] != "ATMOSPHERE", "ParseMainInputFile: :Sublimation moves water from SNOW to ATMOSPHERE");
    AddSVs(pM, {SVL(RVN_SV_SNOW, -1), SVL(RVN_SV_ATMOSPHERE, -1)});
    CmvGeneric *p = new CmvGeneric("Sublimation", RVN_OP_SUBLIMATION, pM);
    p->Connect(pM->GetStateVarIndex(RVN_SV_SNOW), iAtmos);
    const int iSD = pM->GetStateVarIndex(RVN_SV_SNOW_DEPTH);   // depth change is modelled only if SNOW_DEPTH exists when the process is created
    if (iSD != DOESNT_EXIST) p->Connect(iSD, iSD);
    p->SetIntArg(0, type);
    return p;
  }
  if (cmd == ":OpenWaterEvaporation") {  // case 211; src/OpenWaterEvap.cpp:17-38
    need(4);
    ExitGracefullyIf(s[1] != "OPEN_WATER_EVAP" && s[1] != "OPEN_WATER_RIPARIAN", "ParseMainInputFile: open water evaporation algorithm " + s[1] + " is not implemented in the B200 build");
    AddSVs(pM, {SVL(RVN_SV_ATMOSPHERE, -1), SVL(RVN_SV_AET, -1)});
    int from = pM->ParseSVTypeIndex(s[2]);
    CmvGeneric *p = new CmvGeneric("OpenWaterEvaporation", s[1] == "OPEN_WATER_RIPARIAN" ? RVN_OP_OPEN_WATER_RIPARIAN : RVN_OP_OPEN_WATER_EVAP, pM);
    int iAET = pM->GetStateVarIndex(RVN_SV_AET);
    p->Connect(from, iAtmos); p->Connect(iAET, iAET);
    return p;
  }
  if (cmd == ":LakeEvaporation") {  // case 217; src/OpenWaterEvap.cpp:206-222
    need(4);
    ExitGracefullyIf(s[1] != "LAKE_EVAP_BASIC" && s[1] != "BASIC", "ParseMainInputFile: Unrecognized Lake Evaporation process representation");
    AddSVs(pM, {SVL(RVN_SV_ATMOSPHERE, -1), SVL(RVN_SV_AET, -1)});
    int lake_ind = pM->GetLakeStorageIndex();   // the 'from' token is ignored by the reference when four tokens are given
    ExitGracefullyIf(lake_ind == DOESNT_EXIST, "CmvLakeEvaporation Constructor: invalid 'from' compartment specified");
    CmvGeneric *p = new CmvGeneric("LakeEvaporation", RVN_OP_LAKE_EVAP_BASIC, pM);
    int iAET = pM->GetStateVarIndex(RVN_SV_AET);
    p->Connect(lake_ind, iAtmos); p->Connect(iAET, iAET);
    p->SetIntArg(0, 1);   // source == :LakeStorage variable
    return p;
  }
  if (cmd == ":GlacierMelt") {  // case 219; src/GlacierProcesses.cpp:17-36
    need(4);
    rvn_opcode op = RVN_OP_NOP;
    if (s[1] == "GMELT_HBV" || s[1] == "HBV") op = RVN_OP_GLACIERMELT_HBV;
    else if (s[1] == "GMELT_SIMPLE_MELT" || s[1] == "SIMPLE") op = RVN_OP_GLACIERMELT_SIMPLE;
    else if (s[1] == "GMELT_UBC" || s[1] == "UBC") op = RVN_OP_GLACIERMELT_UBC;
    else ExitGracefully("ParseMainInputFile: glacier melt algorithm " + s[1] + " is not implemented in the B200 build");
    AddSVs(pM, {SVL(RVN_SV_GLACIER_ICE, -1), SVL(RVN_SV_GLACIER, -1)});
    if (op == RVN_OP_GLACIERMELT_UBC) AddSV1(pM, RVN_SV_GLACIER_CC);   // src/GlacierProcesses.cpp:95-103
    CmvGeneric *p = new CmvGeneric("GlacierMelt", op, pM);
    if (op == RVN_OP_GLACIERMELT_UBC) {   // :29-36: melt goes to PONDED_WATER, cold content of the ice is a self-connection
      const int iCC = pM->GetStateVarIndex(RVN_SV_GLACIER_CC);
      p->Connect(pM->GetStateVarIndex(RVN_SV_GLACIER_ICE), iPond); p->Connect(iCC, iCC);
    } else p->Connect(pM->GetStateVarIndex(RVN_SV_GLACIER_ICE), pM->GetStateVarIndex(RVN_SV_GLACIER));
    return p;
  }
  if (cmd == ":GlacierRelease") {  // case 220; src/GlacierProcesses.cpp:200-211,240-251
    need(4);
    const bool hbvec = (s[1] == "GRELEASE_HBV_EC" || s[1] == "HBV_EC"), lin = (s[1] == "GRELEASE_LINEAR" || s[1] == "LINEAR_STORAGE"), ana = (s[1] == "GRELEASE_LINEAR_ANALYTIC");
    ExitGracefullyIf(!hbvec && !lin && !ana, "ParseMainInputFile: glacier release algorithm " + s[1] + " is not implemented in the B200 build");
    if (hbvec) AddSVs(pM, {SVL(RVN_SV_GLACIER, -1), SVL(RVN_SV_SNOW, -1), SVL(RVN_SV_SNOW_LIQ, -1)});
    else AddSV1(pM, RVN_SV_GLACIER);
    CmvGeneric *p = new CmvGeneric("GlacierRelease", hbvec ? RVN_OP_GLACIERRELEASE_HBVEC : RVN_OP_GLACIERRELEASE_LINEAR, pM);
    p->Connect(pM->GetStateVarIndex(RVN_SV_GLACIER), iSW);
    if (!hbvec) p->SetIntArg(0, ana ? 1 : 0);
    return p;
  }
  if (cmd == ":CapillaryRise") {  // case 216; src/CapillaryRise.cpp:20-35
    need(4);
    ExitGracefullyIf(s[1] != "RISE_HBV" && s[1] != "CRISE_HBV", "ParseMainInputFile: Unrecognized capillary rise process representation");
    int from = pM->ParseSVTypeIndex(s[2]), to = pM->ParseSVTypeIndex(s[3]);
    ExitGracefullyIf(pM->GetStateVarType(from) != RVN_SV_SOIL || pM->GetStateVarType(to) != RVN_SV_SOIL, "CmvCapillaryRise::Initialize:CapillaryRise must come from and go to soil units");
    CmvGeneric *p = new CmvGeneric("CapillaryRise", RVN_OP_CAPRISE_HBV, pM);
    p->Connect(from, to);
    p->SetIntArg(0, SoilLayerOf(pM, from));
    p->SetIntArg(1, SoilLayerOf(pM, to));
    return p;
  }
  if (cmd == ":Convolve") {  // case 228
    need(4);
    rvn_convol_type ct = RVN_CONVOL_GR4J_1;
    if (s[1] == "CONVOL_GR4J_1") ct = RVN_CONVOL_GR4J_1;
    else if (s[1] == "CONVOL_GR4J_2") ct = RVN_CONVOL_GR4J_2;
    else if (s[1] == "CONVOL_GAMMA") ct = RVN_CONVOL_GAMMA;
    else if (s[1] == "CONVOL_GAMMA_2") ct = RVN_CONVOL_GAMMA_2;
    else ExitGracefully("ParseMainInputFile: Unrecognized convolution process representation");
    pM->IncrementConvolutionCount();
    int conv_index = pM->GetNumConvolutionVariables() - 1;
    sv_type tS[MAX_CONVOL_STORES + 1]; int tL[MAX_CONVOL_STORES + 1]; int tN;
    CmvConvolution::GetParticipatingStateVarList(ct, tS, tL, tN, conv_index);
    pM->AddStateVariables(tS, tL, tN);
    return new CmvConvolution(ct, pM->ParseSVTypeIndex(s[3]), pM, conv_index);
  }
  return NULL;
}
