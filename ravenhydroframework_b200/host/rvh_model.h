// rvh_model.h -- the Raven model container as seen by the hot path, re-authored for the B200 build.
//
// Class and method names mirror the reference so that Raven front ends and user code read the same
// (CModel src/Model.h, CHydroUnit src/HydroUnits.h:21-161, CSubBasin src/SubBasin.h:43-348,
// CGauge src/Gauge.h, CTimeSeries src/TimeSeries.h, class structs src/Properties.h), but the objects
// here are *descriptions*: the numerical state lives on the GPU in SoA form and these objects are
// flattened into rvn_model_desc (include/rvn_gpu.h) by CModel::Flatten().
#ifndef RVH_MODEL_H
#define RVH_MODEL_H
#include "rvh_common.h"

class CModel;
class CHydroProcessABC;

// ---- parameter classes (fixed-stride rows; see rvn_*_field enums) --------------------------------
struct soil_struct    { double v[RVN_S_COUNT]; };
struct surface_struct { double v[RVN_LU_COUNT]; };
struct veg_struct     { double v[RVN_V_COUNT]; double relative_ht[12]; double relative_LAI[12]; };
struct global_struct  { double v[RVN_G_COUNT]; };
int SoilFieldFromName(const std::string &name);     // "POROSITY" -> RVN_S_POROSITY, -1 if unknown
int LandUseFieldFromName(const std::string &name);
int VegFieldFromName(const std::string &name);
int GlobalFieldFromName(const std::string &name);
const char *SoilFieldName(int f);
const char *LandUseFieldName(int f);
const char *VegFieldName(int f);
const char *GlobalFieldName(int f);

struct CSoilClass    { std::string tag; soil_struct S; };
struct CVegetationClass { std::string tag; veg_struct V; };
struct CLandUseClass { std::string tag; surface_struct S; };
struct terrain_struct { double v[RVN_T_COUNT]; };
struct CTerrainClass { std::string tag; terrain_struct T; };
struct CSoilProfile  { std::string tag; int nHorizons; std::vector<int> soil_class; std::vector<double> thickness; };

// trapezoidal / surveyed channel cross-section with the reference's tabulated rating curve
// (CChannelXSect, src/ChannelXSect.cpp): only what Muskingum K/X needs (celerity + top width at Q_ref)
class CChannelXSect {
public:
  std::string name;
  double bedslope, mannings_n;
  std::vector<double> aQ, aStage, aTopWidth, aXArea, aPerim;
  void GenerateTrapezoid(double bottom_w, double bottom_elev, double incline_ratio);
  // surveyed profile (x, bed elevation, Manning's n of the segment to the right of each point) -> 20-point rating curves
  void GenerateFromProfile(const std::vector<double> &x, const std::vector<double> &elev, const std::vector<double> &mann);
  double GetCelerity(double Q, double SB_slope, double SB_n) const;
  double GetTopWidth(double Q, double SB_slope, double SB_n) const;
  double GetArea(double Q, double SB_slope, double SB_n) const;          // src/ChannelXSect.cpp:297-302
  double GetFlowMultiplier(double SB_slope, double SB_n) const;           // Q_mult of GetFlowCorrections (src/ChannelXSect.cpp:345-359)
private:
  double GetQfromQref(double Q, double SB_slope, double SB_n) const;
  void   Interp(double Q, int &i, double &wt) const;
  double min_mannings;
};

// ---- time series & gauges -------------------------------------------------------------------------
class CTimeSeries {
public:
  CTimeSeries(const std::string &name, double start_day, int start_yr, double interval, const std::vector<double> &vals)
      : _name(name), _start_day(start_day), _start_year(start_yr), _interval(interval), _aVal(vals), _t_corr(0), _sampInterval(1) {}
  // is_observation: no overlap QA, one extra sample for the last period-ending value (src/TimeSeries.cpp:214-262)
  void   Initialize(double model_start_day, int model_start_year, double model_duration, double timestep, int calendar, bool is_observation = false);
  double GetValue(double t) const;                  // pulse-type point value at model time t
  double GetValueInterp(double t) const;            // point value of a series parsed as not pulse-based (src/TimeSeries.cpp:320-334)
  double FirstValue() const { return _aVal[0]; }
  double GetAvgValue(double t, double tstep) const;
  double GetSampledValue(int nn) const;
  int    GetNumSampledValues() const { return (int)_aSampVal.size(); }
  bool   IsDaily() const { return _interval == 1.0; }
  double GetInterval() const { return _interval; }
  double GetStartDay() const { return _start_day; }
  int    GetStartYear() const { return _start_year; }
  const std::vector<double> &Values() const { return _aVal; }
  const std::string &GetName() const { return _name; }
private:
  int GetTimeIndex(double t_loc) const;
  std::string _name;
  double _start_day; int _start_year; double _interval;
  std::vector<double> _aVal;
  double _t_corr;
  std::vector<double> _aSampVal; double _sampInterval;
};

enum forcing_type { F_PRECIP, F_RAINFALL, F_SNOWFALL, F_TEMP_AVE, F_TEMP_DAILY_MIN, F_TEMP_DAILY_MAX, F_TEMP_DAILY_AVE,
                    F_PET, F_OW_PET, F_POTENTIAL_MELT, F_NTYPES };
forcing_type ForcingFromString(const std::string &s);  // F_NTYPES if unsupported

class CGauge {
public:
  explicit CGauge(const std::string &name);
  ~CGauge();
  void Initialize(const optStruct &Options);       // reference CGauge::Initialize, src/Gauge.cpp:70-270
  void AddTimeSeries(CTimeSeries *pTS, forcing_type f);
  CTimeSeries *GetTimeSeries(forcing_type f) const { return _pTS[f]; }
  double GetForcingValue(forcing_type f, int nn) const;   // src/Gauge.cpp:533-540
  double GetAverageSnowFrac(int nn) const;                // src/Gauge.cpp:486-492
  std::string name;
  double latitude, longitude, elevation;
  double rainfall_corr, snowfall_corr, temperature_corr;
  double aAveTemp[12], aMinTemp[12], aMaxTemp[12], aAvePET[12];
private:
  CTimeSeries *_pTS[F_NTYPES];
  bool _owned[F_NTYPES];
};

// ---- HRUs and subbasins ---------------------------------------------------------------------------
class CHydroUnit {
public:
  long long ID; int global_k;
  double area, elevation, latitude, longitude, slope, aspect;  // deg in file; GetLatRad()/GetSlope()/GetAspect() in rad
  long long SBID; int subbasin_index;
  int lu_class, veg_class, soil_profile, terrain_class;
  rvn_hru_type type;
  bool enabled;
  double GetArea() const { return area; }
  double GetElevation() const { return elevation; }
  double GetLatRad() const;
  double GetSlope() const;
  double GetAspect() const;
  int    GetSubBasinIndex() const { return subbasin_index; }
  rvn_hru_type GetHRUType() const { return type; }
  bool   IsEnabled() const { return enabled; }
};

// ---- reservoir / lake at a subbasin outlet (reference CReservoir, src/Reservoir.h) --------------------------------------------
// Host half only: the curves (built like the reference's constructors) and the initial state; the level-pool routing itself
// (CReservoir::RouteWater, src/Reservoir.cpp:1532-1853) runs on the device.  Not built: control structures, DZTR, seepage,
// stage / flow constraint time series, stage assimilation, :VaryingStageRelations -- the parser rejects them.
class CReservoir {
public:
  CReservoir(const std::string &name, long long SBID) : _name(name), _SBID(SBID), _HRUID(DOESNT_EXIST), _hru_k(DOESNT_EXIST), _stage(0), _stage_last(0),
      _min_stage(0), _max_stage(0), _Qout(0), _Qout_last(0), _crest_ht(0), _crest_width(DOESNT_EXIST), _max_capacity(0), _pDemandTS(NULL), _demand_mult(1.0) {}
  ~CReservoir() { delete _pDemandTS; }
  // power-law curves (src/Reservoir.cpp:130-180), lookup tables :StageRelations (:190-237), prismatic lake behind a weir (:311-358)
  void SetPowerLaw(double a_V, double b_V, double a_Q, double b_Q, double a_A, double b_A, double crestht, double depth);
  void SetLookup(const std::vector<double> &ht, const std::vector<double> &Q, const std::vector<double> &Qund, const std::vector<double> &A, const std::vector<double> &V);
  void SetLake(double weircoeff, double crestw, double crestht, double A, double depth);
  void SetVolumeStageCurve(const std::vector<double> &ht, const std::vector<double> &V, double weircoeff, double crestwidth);   // :1068-1116
  void SetAreaStageCurve(const std::vector<double> &ht, const std::vector<double> &A);                                           // :1124-1134
  void SetInitialFlow(double initQ, double initQlast, bool justFlow);                                                           // :1413-1485
  void SetReservoirStage(double ht, double ht_last) { if (ht_last != RAV_BLANK_DATA) _stage_last = ht_last; _stage = ht; }      // :1509-1513
  double GetVolume(double ht) const;        // InterpolateCurve(ht, _aStage, _aVolume, _Np, true)
  double GetArea(double ht) const;          // ... (_aArea, false)
  double GetWeirOutflow(double ht) const;   // underflow + curve outflow, no weir-height adjustment series
  long long GetSubbasinID() const { return _SBID; }
  std::string _name; long long _SBID, _HRUID; int _hru_k;
  double _stage, _stage_last, _min_stage, _max_stage, _Qout, _Qout_last, _crest_ht, _crest_width, _max_capacity;
  std::vector<double> _aStage, _aQ, _aQunder, _aArea, _aVolume;
  CTimeSeries *_pDemandTS;   // :ReservoirExtraction / :ReservoirDemand (one per reservoir)
  double _demand_mult;
};

class CSubBasin {
public:
  CSubBasin();
  long long ID; std::string name; long long downstream_ID; std::string profile_name;
  const CChannelXSect *pChannel;
  double reach_length;      // [m] or AUTO_COMPUTE
  bool gauged, is_headwater, disabled;
  bool assimilate;          // :AssimilateStreamflow (CSubBasin::IncludeInAssimilation / UseInFlowAssimilation)
  int  nSegments;
  double basin_area, drainage_area;  // [km2]
  double t_conc, t_peak, t_lag, gamma_shape, gamma_scale, Q_ref, c_ref, w_ref, mannings_n, slope;
  double rain_corr, snow_corr, temperature_corr;
  std::vector<int> hru_index;
  // routing arrays (reference _aUnitHydro/_aRouteHydro/_aQlatHist/_aQinHist/_aQout)
  std::vector<double> aUnitHydro, aRouteHydro, aQlatHist, aQinHist, aQout;
  double QoutLast, QlatLast, Qlocal, QlocLast, channel_storage, rivulet_storage;
  double diff_alpha;        // ROUTE_DIFFUSIVE_VARY: half the diffusivity at the reference flow [m2/d]
  // user-specified inflows [m3/s]: :BasinInflowHydrograph (upstream end of the reach) and :BasinInflowHydrograph2 (downstream end)
  CTimeSeries *pInflowHydro, *pInflowHydro2;
  CReservoir *pReservoir;   // NULL: none (CSubBasin::AddReservoir, src/SubBasin.cpp:1122-1132)
  double GetSpecifiedInflow(double t) const { return pInflowHydro ? pInflowHydro->GetValueInterp(t) : 0.0; }     // src/SubBasin.cpp:478-482
  double GetDownstreamInflow(double t) const { return pInflowHydro2 ? pInflowHydro2->GetValueInterp(t) : 0.0; }  // :488-492
  // reference-surface accessors
  long long GetID() const { return ID; }
  double GetBasinArea() const { return basin_area; }
  double GetReachLength() const { return reach_length; }
  int    GetNumSegments() const { return nSegments; }
  int    GetLatHistorySize() const { return (int)aQlatHist.size(); }
  int    GetInflowHistorySize() const { return (int)aQinHist.size(); }
  const double *GetUnitHydrograph() const { return aUnitHydro.data(); }
  const double *GetRoutingHydrograph() const { return aRouteHydro.data(); }
  double GetOutflowRate() const { return pReservoir ? pReservoir->_Qout : aQout[nSegments - 1]; }   // src/SubBasin.cpp:844-848
  double GetMuskingumK(double dx) const;    // src/SubBasin.cpp:2429-2437
  double GetMuskingumX(double dx) const;    // src/SubBasin.cpp:2443-2452
  bool   SetBasinProperties(const std::string &label, double value);  // src/SubBasin.cpp SetBasinProperties
  void   Initialize(double Qin_avg, double Qlat_avg, double total_drain_area, const CModel *pModel, const optStruct &Options);
  void   InitializeFlowStates(double Qin_avg, double Qlat_avg, const optStruct &Options);
  void   SetQout(double Q);                 // :BasinInitialConditions (src/SubBasin.cpp SetQout)
private:
  void   GenerateCatchmentHydrograph(double Qlat_avg, const optStruct &Options);
  void   GenerateRoutingHydrograph(double Qin_avg, const optStruct &Options);
  void   ResetReferenceFlow(double Qreference);
  double CalculateTOC() const;
};

// ---- groups, observations, forcing perturbations (what the EnKF ensemble refers to) ---------------------------------------
struct CHRUGroup      { std::string name; std::vector<int> hru; bool IsInGroup(int k) const { for (size_t i = 0; i < hru.size(); i++) if (hru[i] == k) return true; return false; } };   // CHRUGroup::IsInGroup (src/HRUGroup.cpp)
struct CSubbasinGroup { std::string name; std::vector<int> sb; };
struct observed_ts    { std::string name; long long loc_ID; CTimeSeries *pTS; };   // :ObservationData [type] [SBID|HRUID]
enum adjustment { ADJ_ADDITIVE = 0, ADJ_MULTIPLICATIVE = 1, ADJ_REPLACE = 2 };
struct force_perturb {  // reference force_perturb (src/Model.h), one :ForcingPerturbation command
  forcing_type forcing; int distribution; double distpar[3]; int kk; adjustment adj_type;
};
enum EnKF_mode { ENKF_SPINUP, ENKF_CLOSED_LOOP, ENKF_OPEN_LOOP, ENKF_FORECAST, ENKF_OPEN_FORECAST, ENKF_UNSPECIFIED };
struct obs_perturb { sv_type state; int distribution; double distpar[3]; int kk; adjustment adj_type; };   // src/EnKF.h:22-29
struct assim_state { sv_type sv; int layer; int group; bool is_streamflow; };
struct enkf_settings {   // what the .rve file says about the EnKF run (src/ParseEnsembleFile.cpp:263-470)
  EnKF_mode mode; int window_size;
  std::vector<obs_perturb> obs_perturbations;
  std::vector<assim_state> assim_states;
  enkf_settings() : mode(ENKF_UNSPECIFIED), window_size(1) {}
};

struct param_dist {  // reference ModelEnsemble.h:20-32
  std::string param_name; int param_class; std::string class_group; int distribution; double distpar[3]; double default_val;
};

// ---- the model ------------------------------------------------------------------------------------
class CModel {
public:
  explicit CModel(int nsoillayers);
  ~CModel();
  // state-variable catalogue (reference CModelABC surface, src/ModelABC.h:18-52)
  int     GetNumStateVars() const { return (int)_aStateVarType.size(); }
  sv_type GetStateVarType(int i) const { return _aStateVarType[i]; }
  int     GetStateVarLayer(int i) const;   // count of earlier SVs of the same type (src/Model.cpp:805-812)
  int     GetStateVarRawLayer(int i) const { return _aStateVarLayer[i]; }
  int     GetStateVarIndex(sv_type typ) const;
  int     GetStateVarIndex(sv_type typ, int layer) const;
  bool    StateVarExists(sv_type typ) const { return GetStateVarIndex(typ) != DOESNT_EXIST; }
  int     GetNumSoilLayers() const { return _nSoilVars; }
  int     GetLakeStorageIndex() const { return _lake_sv; }
  void    SetLakeStorage(int i) { _lake_sv = i; }
  void    AddStateVariables(const sv_type *aSV, const int *aLev, int nSV);  // src/Model.cpp:1561-1612
  std::string GetStateVarName(int i) const;
  // parse "SOIL[1]" / aliases into (type, layer); adds nothing
  sv_type StringToSVType(const std::string &s, int &layer, bool strict) const;
  int     ParseSVTypeIndex(const std::string &s);  // adds the SV if missing, returns its index
  void    AddAlias(const std::string &alias, const std::string &target) { _aliases[alias] = target; }
  static bool IsWaterStorage(sv_type typ);          // src/StateVariables.cpp:612-645 (conv_coverup=true)

  // containers
  int  GetNumHRUs() const { return (int)_pHydroUnits.size(); }
  int  GetNumSubBasins() const { return (int)_pSubBasins.size(); }
  int  GetNumProcesses() const { return (int)_pProcesses.size(); }
  int  GetNumGauges() const { return (int)_pGauges.size(); }
  CHydroUnit *GetHydroUnit(int k) const { return _pHydroUnits[k]; }
  CSubBasin  *GetSubBasin(int p) const { return _pSubBasins[p]; }
  CHydroProcessABC *GetProcess(int j) const { return _pProcesses[j]; }
  CGauge *GetGauge(int g) const { return _pGauges[g]; }
  int  GetSubBasinIndex(long long SBID) const;
  int  GetOrderedSubBasinIndex(int pp) const { return _aOrderedSBind[pp]; }
  int  GetDownstreamBasin(int p) const { return _aDownstreamInds[p]; }
  int  GetSubBasinOrder(int p) const { return _aSubBasinOrder[p]; }
  void AddHRU(CHydroUnit *p) { p->global_k = (int)_pHydroUnits.size(); _pHydroUnits.push_back(p); }
  void AddSubBasin(CSubBasin *p) { _pSubBasins.push_back(p); }
  void AddProcess(CHydroProcessABC *p) { _pProcesses.push_back(p); }
  void AddGauge(CGauge *p) { _pGauges.push_back(p); }
  int  IncrementConvolutionCount() { return _nConvVariables++; }
  int  GetNumConvolutionVariables() const { return _nConvVariables; }

  // classes
  std::vector<CSoilClass> soil_classes;
  std::vector<CVegetationClass> veg_classes;
  std::vector<CLandUseClass> lu_classes;
  std::vector<CTerrainClass> terrain_classes;
  std::vector<CSoilProfile> soil_profiles;
  std::vector<CChannelXSect *> channels;
  global_struct globals;
  int FindSoilClass(const std::string &t) const;
  int FindVegClass(const std::string &t) const;
  int FindLUClass(const std::string &t) const;
  int FindTerrainClass(const std::string &t) const;
  int FindSoilProfile(const std::string &t) const;
  const CChannelXSect *FindChannel(const std::string &t) const;
  const global_struct *GetGlobalParams() const { return &globals; }
  const optStruct *GetOptStruct() const { return _pOpt; }

  // reference CModel::UpdateParameter (src/Model.cpp:2692-2750): class 0=soil 1=veg 2=landuse 3=terrain 4=global
  void UpdateParameter(int ctype, const std::string &pname, const std::string &cname, double value);

  // initial conditions (per HRU rows of NS doubles; filled by the .rvc reader)
  std::vector<double> initial_state;   // [nHRU][NS]
  void   AllocateInitialState();
  void   SetInitialStateVar(int k, int i, double v) { initial_state[(size_t)k * GetNumStateVars() + i] = v; }

  // initialisation (reference CModel::Initialize, src/ModelInitialize.cpp:29-400)
  void Initialize(optStruct &Options);
  double GetWatershedArea() const { return _WatershedArea; }
  const std::vector<double> &GetGaugeWeights() const { return _aGaugeWeights; }  // [nHRU][nGauges]
  // Gridded forcing ingestion (the arithmetic of CForcingGrid::GetWeightedValue, src/ForcingGrid.cpp:1886-1900 = the gauge sum of
  // src/UpdateForcings.cpp:184-235 over one station per grid cell): cell series already sampled on the model step,
  // series[RVN_GF_COUNT][nCells][nSteps], reference elevation per cell, CSR weights (wt_ptr[nHRU+1], wt_idx, wt_val; rows sum to 1).
  // The NetCDF reader of the reference (CForcingGrid::ReadData) is out of scope; callers decode their files and hand the arrays over.
  void SetCellForcing(int nCells, int nSteps, const double *cell_elev, const double *series, const int32_t *wt_ptr, const int32_t *wt_idx, const double *wt_val);

  // ensemble description (.rve)
  std::vector<param_dist> param_dists;
  int num_members;
  enkf_settings enkf;
  // DDS calibration target (.rve :ObjectiveFunction) and the evaluation periods of the .rvi (:EvaluationPeriod; "ALL" always exists)
  long long calib_SBID; int calib_obj; std::string calib_period;
  struct diag_period { std::string name; double t_start, t_end; };
  std::vector<diag_period> diag_periods;
  // groups / observations / forcing perturbations
  std::vector<CHRUGroup> hru_groups;
  std::vector<CSubbasinGroup> sb_groups;
  int  FindHRUGroup(const std::string &name) const;       // DOESNT_EXIST if absent
  int  FindSubBasinGroup(const std::string &name) const;
  std::vector<observed_ts> observed;
  int  GetNumObservedTS() const { return (int)observed.size(); }
  std::vector<force_perturb> perturbations;
  int  GetNumForcingPerturbations() const { return (int)perturbations.size(); }
  void AddForcingPerturbation(forcing_type f, int distribution, const double distpar[3], int kk, adjustment adj);   // src/Model.cpp AddForcingPerturbation
  // sample of perturbation i for (local member, step) in the flattened table handed to the device (pert_eps)
  void SetPerturbationSample(int member, int step, int i, double eps);
  const double *MemberPerturbations(int member) const;

  // ---- compile to the device ABI --------------------------------------------------------------
  // Owns every buffer referenced by the returned descriptor until the next Flatten()/destruction.
  const rvn_model_desc *Flatten(const optStruct &Options, int nMembers);
  // refresh one member's parameter row + convolution UHs from the *current* class structs (after
  // UpdateParameter); writes into the flattened buffers and returns pointers for rvn_gpu_upload_params.
  void FlattenMemberParams(int member);
  const double *MemberParams(int member) const { return _f_ptab.data() + (size_t)member * _f_ptab_stride; }
  const int32_t *MemberConvN(int member) const;
  const double *MemberConvUH(int member) const;
  // checkpoint: the reference's {RunName}_solution.rvc (CModel::WriteMajorOutput, src/StandardOutput.cpp:1388-1451;
  // CSubBasin::WriteToSolutionFile, src/SubBasin.cpp:2869-2880) for ONE member, from arrays in the layouts of
  // rvn_gpu_download_state / rvn_gpu_download_basin_state.  full_precision: %.17g instead of the reference's 5 fixed decimals
  // (6 significant digits for the basin block), so that a restart continues bit-identically.
  // direct-insertion assimilation: the reference's first PrepareAssimilation call also overwrites the initial flows of gauged basins with the
  // first observation (src/Assimilate.cpp:160-178); call after the initial conditions were read
  void WriteHydrographsFile(const std::string &filename, const optStruct &Options, const double *qout_rec, const double *q0, const double *q0last,
                            const double *precip, int nsteps) const;   // Hydrographs.csv from the device's per-step records
  void ApplyInitialAssimilation(const optStruct &Options);
  const CTimeSeries *FlowObservation(const CSubBasin *b) const;   // the continuous HYDROGRAPH series linked to the basin, or NULL
  void WriteSolutionFile(const std::string &filename, const optStruct &Options, double model_time, const double *state, const double *qin,
                         const double *qlat, const double *qout, const double *scal, bool full_precision, const double *res_state = NULL) const;
  // initial basin routing state in the layout of rvn_gpu_upload_basin_state (one member)
  void PackBasinState(std::vector<double> &qin, std::vector<double> &qlat, std::vector<double> &qout, std::vector<double> &scal) const;

private:
  void InitializeRoutingNetwork();
  void InitializeBasins(const optStruct &Options);
  void GenerateGaugeWeights(const optStruct &Options);
  std::vector<sv_type> _aStateVarType;
  std::vector<int> _aStateVarLayer;
  int _nSoilVars, _lake_sv, _nConvVariables;
  std::map<std::string, std::string> _aliases;
  std::vector<CHydroUnit *> _pHydroUnits;
  std::vector<CSubBasin *> _pSubBasins;
  std::vector<CHydroProcessABC *> _pProcesses;
  std::vector<CGauge *> _pGauges;
  std::vector<int> _aOrderedSBind, _aDownstreamInds, _aSubBasinOrder;
  std::vector<double> _aGaugeWeights;
  // gridded forcing input (SetCellForcing): one station per grid cell with sparse HRU weights; replaces the :Gauge series at Flatten
  bool _cell_forcing = false;
  int _nCells = 0;
  std::vector<double> _f_subdaily;   // [nEtRows][nHRU] SUBDAILY_SIMPLE correction factors
  std::vector<double> _f_p5;   // [nGauges][nSteps] 5-day antecedent precipitation of the gauges (INF_SCS)
  std::vector<double> _cell_elev, _cell_series, _cell_wt;   // [nCells], [RVN_GF_COUNT][nCells][nSteps], [nnz]
  std::vector<int32_t> _cell_ptr, _cell_idx;                // CSR over HRUs
  double _WatershedArea;
  const optStruct *_pOpt;
  // flattened storage
  rvn_model_desc _desc;
  std::vector<int32_t> _f_svtype, _f_svlayer, _f_from, _f_to, _f_hru_i[6], _f_prof_class, _f_conv_n, _f_sb_i[7];
  std::vector<rvn_process> _f_proc;
  std::vector<uint64_t> _f_mask;
  std::vector<double> _f_hru_d[5], _f_prof_thick, _f_ptab, _f_conv_uh, _f_series, _f_gconst, _f_wt, _f_sb_d[7], _f_uh, _f_rh, _f_veg_season, _f_et;
  std::vector<int32_t> _f_et_row, _f_wt_ptr, _f_wt_idx, _f_pert_forcing, _f_pert_adj, _f_hru_terr, _f_spec_basin;
  std::vector<int32_t> _f_res_basin, _f_res_hru, _f_res_np;
  std::vector<double> _f_res_curves, _f_res_min_stage, _f_res_extract, _f_res_state0;
  std::vector<double> _f_sb_cref, _f_sb_alpha;
  std::vector<double> _pre_assim_q0, _pre_assim_q0last;   // outflows of the initial condition before ApplyInitialAssimilation (row 0 of Hydrographs.csv)
  std::vector<uint8_t> _f_da_override;
  std::vector<double> _f_da_obsQ, _f_da_obsQ2, _f_da_decay, _f_da_wt, _f_sb_drain, _f_sb_dacorr;
  std::vector<double> _f_spec_in, _f_spec_down, _f_spec_cum;
  std::vector<uint8_t> _f_pert_mask;
  std::vector<rvn_lateral> _f_lat;
  std::vector<int32_t> _f_lat_kfrom, _f_lat_kto, _f_lat_chunk;
  std::vector<int> _f_lat_flag;
  std::vector<double> _f_daylen, _f_sunset, _f_etflat, _f_mopt;
  std::vector<double> _f_lat_frac, _f_chan_Q, _f_chan_A, _f_sb_Qmult, _f_sb_reach;
  std::vector<int32_t> _f_sb_channel;
  std::vector<double> _f_pert_eps;
  std::vector<rvn_time> _f_times;
  int _f_ptab_stride, _f_nconv;
};

// front-end entry points that keep the reference names/signatures (src/RavenMain.h)
bool ParseInputFiles(CModel *&pModel, optStruct &Options);           // .rvi -> .rvp -> .rvh -> .rvt (-> .rve)
bool ParseInitialConditions(CModel *&pModel, const optStruct &Options); // .rvc
#endif
