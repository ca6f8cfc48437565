// rvh_process.h -- host half of the hydrological-process plugin contract.
//
// Reference contract (src/HydroProcessABC.h:42-113): a process owns its iFrom[]/iTo[] connection
// lists, optional conditions (ShouldApply), names the parameters and state variables it needs, and
// returns rates through GetRatesOfChange()/ApplyConstraints().  In the B200 build the connection
// lists, conditions and parameter/state lists stay on the host with the same accessors, while the two
// rate functions are __device__ functions of the same names in csrc/rvn_processes.cuh, selected by the
// opcode that Compile() emits.  (There is deliberately no host implementation of the rate functions:
// no CPU fallback exists on the product path.)
#ifndef RVH_PROCESS_H
#define RVH_PROCESS_H
#include "rvh_model.h"

enum condition_basis { BASIS_HRU_TYPE, BASIS_HRU_GROUP, BASIS_LANDCLASS, BASIS_VEGETATION };
enum comparison { COMPARE_IS_EQUAL, COMPARE_NOT_EQUAL };
struct condition { condition_basis basis; comparison compare_method; std::string data; };
enum class_type { CLASS_SOIL = 0, CLASS_VEGETATION = 1, CLASS_LANDUSE = 2, CLASS_TERRAIN = 3, CLASS_GLOBAL = 4 };

rvn_hru_type StringToHRUType(const std::string &s);

class CHydroProcessABC {
public:
  CHydroProcessABC(const std::string &process_name, CModel *pM) : _name(process_name), pModel(pM) {}
  virtual ~CHydroProcessABC() {}
  const int *GetFromIndices() const { return iFrom.data(); }
  const int *GetToIndices() const { return iTo.data(); }
  int  GetNumConnections() const { return (int)iFrom.size(); }
  const std::string &GetProcessName() const { return _name; }
  bool ShouldApply(const CHydroUnit *pHRU) const;                       // src/HydroProcessABC.cpp:290-328
  void AddCondition(condition_basis basis, comparison cmp, const std::string &data);
  void Redirect(int toSVindex, int newToSVindex);                       // src/HydroProcessABC.cpp:273-283
  virtual void Initialize() {}
  virtual void GetParticipatingParamList(std::vector<std::string> &aP, std::vector<class_type> &aPC) const {}
  // emit the device opcode record; conn arrays are appended by the caller from iFrom/iTo
  virtual void Compile(rvn_process &rec) const = 0;
protected:
  void DynamicSpecifyConnections(int n) { iFrom.assign(n, DOESNT_EXIST); iTo.assign(n, DOESNT_EXIST); }
  std::string _name;
  CModel *pModel;
  std::vector<int> iFrom, iTo;
  std::vector<condition> _conditions;
};

// generic single-opcode mover: used for every process whose host half is only "connections + opcode"
class CmvGeneric : public CHydroProcessABC {
public:
  CmvGeneric(const std::string &pname, rvn_opcode op, CModel *pM) : CHydroProcessABC(pname, pM), _op(op) {
    for (int i = 0; i < 5; i++) _ia[i] = DOESNT_EXIST;
    _da[0] = _da[1] = 0.0;
  }
  void Connect(int from, int to) { iFrom.push_back(from); iTo.push_back(to); }
  void SetIntArg(int i, int v) { _ia[i] = v; }
  void SetDblArg(int i, double v) { _da[i] = v; }
  void Compile(rvn_process &rec) const override;
  rvn_opcode opcode() const { return _op; }
protected:
  rvn_opcode _op;
  int _ia[5];
  double _da[2];
};

// CmvPrecipitation: connections depend on which state variables exist once the whole process list is
// parsed, so they are built in Initialize() (reference src/PartitionPrecip.cpp:28-73)
class CmvPrecipitation : public CmvGeneric {
public:
  explicit CmvPrecipitation(CModel *pM) : CmvGeneric("Precipitation", RVN_OP_PRECIPITATION, pM) {}
  void Initialize() override;
  static void GetParticipatingStateVarList(sv_type *aSV, int *aLev, int &nSV);
};

// CmvConvolution (reference src/Convolution.cpp:31-64): 2N+1 connections, N=MAX_CONVOL_STORES
class CmvConvolution : public CmvGeneric {
public:
  CmvConvolution(rvn_convol_type type, int to_index, CModel *pM, int conv_index);
  static void GetParticipatingStateVarList(rvn_convol_type t, sv_type *aSV, int *aLev, int &nSV, int conv_index);
  rvn_convol_type type() const { return _type; }
  int conv_index() const { return _conv_index; }
  // unit hydrograph for one land-use class (reference GenerateUnitHydrograph, src/Convolution.cpp:107-154)
  void GenerateUnitHydrograph(const surface_struct &S, const optStruct &Options, double *aUnitHydro, int &N) const;
private:
  double LocalCumulDist(double t, const surface_struct &S) const;
  rvn_convol_type _type;
  int _conv_index;
};

// CProcessGroup (reference src/ProcessGroup.cpp): weighted blend of sub-process rates.  The group is ONE process of the model
// (one ShouldApply gate, connections = concatenation of the members'); on the device its members are consecutive records.
class CProcessGroup : public CHydroProcessABC {
public:
  CProcessGroup(const std::string &name, CModel *pM) : CHydroProcessABC("ProcessGroup", pM), _group_name(name) {}
  ~CProcessGroup() override { for (size_t i = 0; i < _pSubProcesses.size(); i++) delete _pSubProcesses[i]; }
  void AddProcess(CHydroProcessABC *pProc) { _pSubProcesses.push_back(pProc); _aWeights.assign(_pSubProcesses.size(), 1.0); }   // weights default to 1.0
  int  GetGroupSize() const { return (int)_pSubProcesses.size(); }
  const CHydroProcessABC *GetSubProcess(int i) const { return _pSubProcesses[i]; }
  double GetWeight(int i) const { return _aWeights[i]; }
  void SetWeights(const double *aWts, int nVal);     // src/ProcessGroup.cpp:180-196
  void CalcWeights(const double *aVals, int nVal);   // src/ProcessGroup.cpp:203-210 -> CalcWeightsFromUniformNums
  void Initialize() override;                        // src/ProcessGroup.cpp:32-56
  void Compile(rvn_process &rec) const override { (void)rec; ExitGracefully("CProcessGroup::Compile: groups are flattened member by member"); }
private:
  std::string _group_name;
  std::vector<CHydroProcessABC *> _pSubProcesses;
  std::vector<double> _aWeights;
};
// CLateralExchangeProcessABC (reference src/LateralExchangeABC.h:20-75): a process that moves water between HRUs.  No vertical
// connections; the lateral ones (_kFrom/_kTo with one state variable on either side) are built in Initialize() once the HRU
// groups are populated.  Connections are stored grouped into chunks of HRUs no other chunk touches (see rvn_model_desc.lat*).
class CLateralExchangeProcessABC : public CHydroProcessABC {
public:
  CLateralExchangeProcessABC(const char *name, int kind, CModel *pM) : CHydroProcessABC(name, pM), _kind(kind), _iFromLat(DOESNT_EXIST), _iToLat(DOESNT_EXIST), _method(0), _mixing_rate(0.0) {}
  void Compile(rvn_process &rec) const override { (void)rec; ExitGracefully("CLateralExchangeProcessABC::Compile: lateral processes have no vertical record"); }
  int  GetLateralKind() const { return _kind; }
  int  GetNumLatConnections() const { return (int)_kFrom.size(); }
  const int *GetFromHRUIndices() const { return _kFrom.data(); }
  const int *GetToHRUIndices() const { return _kTo.data(); }
  const double *GetFractions() const { return _aFrac.data(); }
  const int *GetFlags() const { return _aFlag.data(); }
  int  GetLateralFromIndex() const { return _iFromLat; }
  int  GetLateralToIndex() const { return _iToLat; }
  double GetMixingRate() const { return _mixing_rate; }
  int  GetMethod() const { return _method; }
  const std::vector<int> &ChunkOffsets() const { return _chunk_ptr; }
protected:
  int _kind, _iFromLat, _iToLat, _method;
  double _mixing_rate;
  std::vector<int> _kFrom, _kTo, _aFlag, _chunk_ptr;
  std::vector<double> _aFrac;
};
// CmvLatFlush (reference src/LatFlush.cpp:16-177): moves the whole content of one state variable of the 'from' group's HRUs to a
// state variable of the 'to' group's HRUs -- within each subbasin, split by recipient area, or to one single recipient HRU when
// INTERBASIN is given.
class CmvLatFlush : public CLateralExchangeProcessABC {
public:
  CmvLatFlush(int from_sv_ind, int to_sv_ind, int from_HRU_grp, int to_HRU_grp, bool constrain_to_SBs, CModel *pM);
  void Initialize() override;
  bool IsConstrainedToSubBasins() const { return _constrain_to_SBs; }
private:
  int _kk_from, _kk_to;
  bool _constrain_to_SBs;
};
// CmvLatEquilibrate (reference src/LatEquilibrate.cpp:14-212): partially mixes one state variable among the HRUs of a group, within
// each subbasin (or across the whole model with INTERBASIN): a fraction min(mixing_rate*dt,1) of every member's storage is pooled
// in the group's first HRU of the basin and handed back in proportion to area.
class CmvLatEquilibrate : public CLateralExchangeProcessABC {
public:
  CmvLatEquilibrate(int sv_ind, int HRU_grp, double mixing_rate, bool constrain_to_SBs, CModel *pM);
  void Initialize() override;
private:
  int _kk;
  bool _constrain_to_SBs;
};
// CmvLatRedistribute (reference src/SnowRedistribute.cpp:14-215): gravitational snow redistribution along the user's
// :LateralConnections SNOW_REDISTRIBUTE table (source HRU ID, recipient HRU ID, weight) of the .rvh file.
class CmvLatRedistribute : public CLateralExchangeProcessABC {
public:
  CmvLatRedistribute(int sv_ind, double max_snow_height, int method, CModel *pM);
  void AddConnection(long long src_id, long long dst_id, double w) { _src.push_back(src_id); _dst.push_back(dst_id); _w.push_back(w); }
  void ClearConnections() { _src.clear(); _dst.clear(); _w.clear(); }
  void Initialize() override;
private:
  std::vector<long long> _src, _dst;
  std::vector<double> _w;
};
void CalcWeightsFromUniformNums(const double *aVals, double *aWeights, int N);   // src/CommonFunctions.cpp:1845-1853

// factory used by the .rvi reader: builds the process for a ':ProcessName ALGORITHM FROM TO' line,
// registering its state variables in the reference's order.  Returns NULL for unknown commands.
CHydroProcessABC *CreateProcessFromRVI(const std::vector<std::string> &s, CModel *pModel, const optStruct &Options);
#endif
