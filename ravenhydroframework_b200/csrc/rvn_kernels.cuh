// rvn_kernels.cuh -- the sm_100a kernels behind include/rvn_gpu.h.
//
// Per model step two kernels run for ALL ensemble members at once:
//   hru_sweep_*        = CModel::UpdateHRUForcingFunctions (src/UpdateForcings.cpp:30-668, fused in) +
//                        the ORDERED_SERIES HRU loop of MassEnergyBalance (src/Solvers.cpp:155-251) +
//                        routing prep / TOTAL_SWE / write-back (:495-511, :697-729) + the per-HRU part
//                        of the output reductions (src/Model.cpp:1048-1060,1159-1174).
//                        One thread owns one (member, HRU) pair.  Two forms share every process function:
//        hru_sweep_static<Prog>   the process list is a compile-time program (rvn_static_programs.cuh): state and rates
//                                 live in registers, every connection index is a constant, the process loop is unrolled
//                                 ("one fused FP64 kernel per process group")
//        hru_sweep_interp         the same sweep as an interpreter over the DevProgram for arbitrary process lists:
//                                 state and rates live in shared-memory columns indexed at run time
//                        In both the 50-slot convolution stores are streamed HBM -> registers -> HBM, coalesced over
//                        neighbouring HRUs, with a register double buffer; no CTA-wide barrier inside the time-critical part.
//   route_level_kernel (one launch per level), balance_kernel
//                      = HRU->subbasin aggregation (:490-511), the upstream->downstream routing loop
//                        (:565-615; CSubBasin::RouteWater/UpdateOutflows, src/SubBasin.cpp:2584-2863,
//                        2320-2421) level by level, IncrementCumulInput/CumOutflow (src/Model.cpp:
//                        2458-2539) and the watershed-storage total (src/StandardOutput.cpp:745-785).
//                        One thread per (member, basin); the level order is the reference's precomputed topological order.
//                        They run on a second, higher-priority stream: routing of step s overlaps the sweep of step s+1.
// Bound: HBM bandwidth (state is read once and written once per step, ~0.5 flop/byte); no tensor cores.
#ifndef RVN_KERNELS_CUH
#define RVN_KERNELS_CUH
#include <cuda_runtime.h>
#include "rvn_device.cuh"
#include "rvn_processes.cuh"

// ------------------------------------------------------------------------------------------------
// contexts
// ------------------------------------------------------------------------------------------------
#ifndef RVN_F_SMEM
#define RVN_F_SMEM 0   // 1: the derived forcings of the thread's HRU live in a shared-memory column instead of registers
#endif
#if RVN_F_SMEM
struct FVec { double *p; RVN_DI double &operator[](int f) const { return p[f * RVN_BS]; } };   // forcing f of this thread at p[f * RVN_BS]
#define RVN_F_SMEM_DOUBLES (RVN_F_CARRIED * RVN_BS)
#else
struct FVec { double v[RVN_F_CARRIED]; RVN_DI double &operator[](int f) { return v[f]; } RVN_DI const double &operator[](int f) const { return v[f]; } };
#define RVN_F_SMEM_DOUBLES 0
#endif
struct CtxCommon {
  FVec F;                  // derived forcings of this HRU for this step
  // lagged (start-of-step) values (what pHRU->GetStateVarValue returns inside processes)
  double old_snow, old_snow_cover, old_snow_depth, old_snow_temp;
  double air_pres, rel_hum, wind_vel;   // atmospheric forcings a process may read (only derived when opt.need_atmos)
  double area, elev, lat, slope, aspect;
  double tstep;
  double canopy_cap, canopy_snow_cap;   // veg_var_struct capacity / snow_capacity of this HRU for this step
  double rain_icept_pct, snow_icept_pct; // veg_var_struct interception fractions (:PrecipIceptFract)
  double precip_5day; int month;          // force_struct::precip_5day and the calendar month of the step (INF_SCS; derived when M.gauge_p5 is given)
  double veg_LAI, veg_SAI, day_length;   // veg_var_struct LAI / SAI and force_struct day_length of this HRU and step: set when the process list holds SNOBAL_COLD_CONTENT (M.need_cc)
  int type, lu, veg, profile;
  bool linked;   // HRU linked to a reservoir (CHydroUnit::IsLinkedToReservoir)
  const double *ptab;      // member parameter row (shared memory)
  const rvn_options *popt;
  int glob_off, lu_base, veg_base, soil_off, terr_base;
  RVN_DI double TERR(int f) const { return ptab[terr_base + f]; }
  RVN_DI double G(int f) const { return ptab[glob_off + f]; }
  RVN_DI double LU(int f) const { return ptab[lu_base + f]; }
  RVN_DI double VEG(int f) const { return ptab[veg_base + f]; }
  RVN_DI const rvn_options &opt() const { return *popt; }
};

// run-time context of the interpreter: shared-memory columns (column-major over the CTA's threads so that dynamic
// indexing by state-variable slot is bank-conflict free)
struct DynCtx : CtxCommon {
  static constexpr bool kAtmos = true;   // atmospheric forcing chain / non-USER interception compiled in (the engine routes such models here)
  const DevProgram *P;     // program (shared memory)
  double *sv;              // &s_sv[tid]; slot i at sv[i*RVN_BS]  -- the state the processes read: newest (aPhinew[k], ORDERED_SERIES) or
                           //             start-of-step (aPhi[k], EULER)
  double *svw;             // the state the solver updates (aPhinew[k]); == sv for ORDERED_SERIES
  double *rt;              // &s_rates[tid]; connection q at rt[q*RVN_BS]
  double *grt;             // &s_grates[tid]: rate vector of the current process group
  const int32_t *prof_class; const double *prof_thick;   // global, row of this HRU's soil profile
  RVN_DI double &SV(int slot) const { return sv[slot * RVN_BS]; }
  RVN_DI double &SVW(int slot) const { return svw[slot * RVN_BS]; }
  RVN_DI double &R(int q) const { return rt[q * RVN_BS]; }
  RVN_DI double &GR(int q) const { return grt[q * RVN_BS]; }
  RVN_DI int iSV(int t) const { return P->iSV[t]; }
  RVN_DI int slot_type(int s) const { return P->slot_type[s]; }
  RVN_DI int slot_layer(int s) const { return P->slot_layer[s]; }
  RVN_DI int slot_water(int s) const { return P->slot_water[s]; }
  RVN_DI int soil_slot(int m) const { return P->soil_slot[m]; }
  RVN_DI double SOIL(int m, int f) const { return ptab[soil_off + prof_class[m] * RVN_S_COUNT + f]; }
  RVN_DI double thick(int m) const { return prof_thick[m]; }
};
struct DynProc {
  const rvn_process *p; const DevProgram *P;
  RVN_DI int from(int q) const { return P->conn_from[p->conn0 + q]; }
  RVN_DI int to(int q) const { return P->conn_to[p->conn0 + q]; }
  RVN_DI int nconn() const { return p->nconn; }
  RVN_DI int ia(int i) const { return p->ia[i]; }
  RVN_DI double da(int i) const { return p->da[i]; }
  RVN_DI int conn0() const { return p->conn0; }
  RVN_DI int gconn0() const { return p->gconn0; }
  RVN_DI int gnconn() const { return p->gnconn; }
  RVN_DI double weight() const { return p->weight; }
  RVN_DI int gfrom(int q) const { return P->conn_from[p->gconn0 + q]; }
  RVN_DI int gto(int q) const { return P->conn_to[p->gconn0 + q]; }
};

// compile-time context of the specialised kernels: everything indexed by constants -> registers
#ifndef RVN_SV_SMEM
#define RVN_SV_SMEM 1   // 1: the specialised sweep keeps the core state vector in shared-memory columns instead of registers
#endif
template <class Prog>
struct StatCtx : CtxCommon {
  static constexpr bool kAtmos = false;
#if RVN_SV_SMEM
  double *svp;   // &s_sv[tid]; slot s at svp[s * RVN_BS] (constant offsets; no register pressure, no local-memory spills)
#else
  double sv[Prog::nCore];
#endif
  double rt[Prog::maxConn];
  double grt[Prog::maxGroupConn > 0 ? Prog::maxGroupConn : 1];
  int sbase[Prog::nSoil > 0 ? Prog::nSoil : 1];
  double sthick[Prog::nSoil > 0 ? Prog::nSoil : 1];
#if RVN_SV_SMEM
  RVN_DI double &SV(int slot) { return svp[slot * RVN_BS]; }
  RVN_DI const double &SV(int slot) const { return svp[slot * RVN_BS]; }
#else
  RVN_DI double &SV(int slot) { return sv[slot]; }
  RVN_DI const double &SV(int slot) const { return sv[slot]; }
#endif
  RVN_DI double &SVW(int slot) { return SV(slot); }   // specialised kernels are ORDERED_SERIES only
  RVN_DI double &R(int q) { return rt[q]; }
  RVN_DI double &GR(int q) { return grt[q]; }
  static constexpr RVN_DI int iSV(int t) { return Prog::iSV(t); }
  static constexpr RVN_DI int slot_type(int s) { return Prog::slot_type(s); }
  static constexpr RVN_DI int slot_layer(int s) { return Prog::slot_layer(s); }
  static constexpr RVN_DI int slot_water(int s) { return Prog::slot_water(s); }
  static constexpr RVN_DI int soil_slot(int m) { return Prog::soil_slot(m); }
  RVN_DI double SOIL(int m, int f) const { return ptab[sbase[m] + f]; }
  RVN_DI double thick(int m) const { return sthick[m]; }
};
template <class Prog, int J>
struct StatProc {
  static constexpr RVN_DI int from(int q) { return Prog::conn_from(Prog::conn0(J) + q); }
  static constexpr RVN_DI int to(int q) { return Prog::conn_to(Prog::conn0(J) + q); }
  static constexpr RVN_DI int nconn() { return Prog::nconn(J); }
  static constexpr RVN_DI int ia(int i) { return Prog::ia(J, i); }
  static constexpr RVN_DI double da(int i) { return Prog::da(J, i); }
  static constexpr RVN_DI int conn0() { return Prog::conn0(J); }
  static constexpr RVN_DI int gconn0() { return Prog::gconn0(J); }
  static constexpr RVN_DI int gnconn() { return Prog::gnconn(J); }
  const double *w;   // run-time group weights [nProc] (the only part of a program that is not compile-time: calibration varies them)
  RVN_DI double weight() const { return __ldg(w + J); }
  static constexpr RVN_DI int gfrom(int q) { return Prog::conn_from(Prog::gconn0(J) + q); }
  static constexpr RVN_DI int gto(int q) { return Prog::conn_to(Prog::gconn0(J) + q); }
};

// ------------------------------------------------------------------------------------------------
// forcing derivation for one HRU (restates CModel::UpdateHRUForcingFunctions for gauge-based input)
// ------------------------------------------------------------------------------------------------
// CModel::ApplyForcingPerturbation (src/Model.cpp:2856-2879) for forcing type `ftype`: CHydroUnit::AdjustHRUForcing
// (src/HydroUnits.cpp:461-509) followed, at the start of a day (every step of a daily model), by AdjustDailyHRUForcings
// (:513-538).  One sample per perturbation and day: eps[nn] = eps[0] and the daily mean adjustment 0.0 + eps/1 = eps.
RVN_DI void apply_perturbations(FVec &F, const DevModel &M, const int ftype, const int e, const int k, const int step) {
  for (int i = 0; i < M.nPert; i++) {
    if (M.pert_forcing[i] != ftype) continue;
    if (M.pert_mask != nullptr && M.pert_mask[(size_t)i * M.nHp + k] == 0) continue;
    const double eps = M.pert_eps[((size_t)e * M.nSteps + step) * M.nPert + i];
    const int adj = M.pert_adj[i];
    if (ftype == RVN_PF_PRECIP) {
      if (adj == RVN_ADJ_MULTIPLICATIVE) F[RVN_F_PRECIP] *= eps;
      else if (adj == RVN_ADJ_ADDITIVE) F[RVN_F_PRECIP] += eps;
      else F[RVN_F_PRECIP] = eps;
      if (0.0 > F[RVN_F_PRECIP]) F[RVN_F_PRECIP] = 0.0;   // upperswap
    } else if (ftype == RVN_PF_RAINFALL) {
      const double sf = F[RVN_F_SNOW_FRAC], Po = F[RVN_F_PRECIP];
      if (adj == RVN_ADJ_MULTIPLICATIVE) F[RVN_F_PRECIP] += Po * (1.0 - sf) * (eps - 1.0);
      else if (adj == RVN_ADJ_ADDITIVE) F[RVN_F_PRECIP] += eps;
      else F[RVN_F_PRECIP] += eps - (1 - sf) * Po;
      if (0.0 > F[RVN_F_PRECIP]) F[RVN_F_PRECIP] = 0.0;
      if (F[RVN_F_PRECIP] == 0) F[RVN_F_SNOW_FRAC] = 0;
      else F[RVN_F_SNOW_FRAC] = F[RVN_F_SNOW_FRAC] * Po / F[RVN_F_PRECIP];
    } else if (ftype == RVN_PF_SNOWFALL) {
      const double sf = F[RVN_F_SNOW_FRAC], Po = F[RVN_F_PRECIP];
      if (adj == RVN_ADJ_MULTIPLICATIVE) F[RVN_F_PRECIP] += Po * (sf) * (eps - 1.0);
      else if (adj == RVN_ADJ_ADDITIVE) F[RVN_F_PRECIP] += eps;
      else F[RVN_F_PRECIP] += eps - (sf)*Po;
      if (0.0 > F[RVN_F_PRECIP]) F[RVN_F_PRECIP] = 0.0;
      if (F[RVN_F_PRECIP] == 0) F[RVN_F_SNOW_FRAC] = 0;
      else F[RVN_F_SNOW_FRAC] = 1.0 - (1.0 - F[RVN_F_SNOW_FRAC]) * Po / F[RVN_F_PRECIP];
    } else if (ftype == RVN_PF_TEMP_AVE) {
      if (adj == RVN_ADJ_MULTIPLICATIVE) F[RVN_F_TEMP_AVE] *= eps;
      else if (adj == RVN_ADJ_ADDITIVE) F[RVN_F_TEMP_AVE] += eps;
      else F[RVN_F_TEMP_AVE] = eps;
      const double adjust = 0.0 + eps / 1;
      if (adj == RVN_ADJ_MULTIPLICATIVE) { F[RVN_F_TEMP_DAILY_MIN] *= adjust; F[RVN_F_TEMP_DAILY_MAX] *= adjust; F[RVN_F_TEMP_DAILY_AVE] *= adjust; }
      else if (adj == RVN_ADJ_ADDITIVE) { F[RVN_F_TEMP_DAILY_AVE] += adjust; F[RVN_F_TEMP_DAILY_MIN] += adjust; F[RVN_F_TEMP_DAILY_MAX] += adjust; }
    }
  }
}

// GetSaturatedVaporPressure  src/CommonFunctions.cpp:935-949 [kPa]
RVN_DI double sat_vapor_pressure(const double T) {
  return (T >= 0) ? 0.61078 * rvn_exp(17.26939 * T / (T + 237.3)) : 0.61078 * rvn_exp(21.87456 * T / (T + 265.5));
}
template <class C>
RVN_DI void compute_forcings(C &c, const DevModel &M, int e, int k, int step, int sb) {
  const rvn_options &opt = c.opt();
  const rvn_time tt = M.times[step];
  const int nG = M.nGauges, mo = tt.month;
  const double elev = c.elev;
  FVec &F = c.F;
  double ref_elev_temp = 0.0, ref_elev_precip = 0.0, temp_month_min = 0, temp_month_max = 0, temp_month_ave = 0, PET_month_ave = 0;
#pragma unroll
  for (int f = 0; f < RVN_F_CARRIED; f++) F[f] = 0.0;
  // K3: sparse gather over this HRU's gauges / grid cells (CSR, ascending gauge index = the reference's summation order;
  // omitted zero weights add exact zeros there).  The step's series form one contiguous [gauge][forcing] slab: the 9 values of a
  // station / grid cell are 72 contiguous bytes, so the gather of a scattered HRU touches 2-3 sectors per cell instead of 9.
  const double *slab = M.gauge_series + (size_t)step * RVN_GF_COUNT * nG;
#define GS(f, g) slab[(size_t)(g) * RVN_GF_COUNT + (f)]
#define GC(g, f) M.gauge_const[(size_t)(g) * RVN_GC_COUNT + (f)]
  // single-station models (every weight exactly 1.0): no CSR loads at all, the loops below run once with g = 0, wt = 1.0
  const bool uw = (M.unit_weights != 0);
  const int w0 = uw ? 0 : M.wt_ptr[k], w1 = uw ? 1 : M.wt_ptr[k + 1];
#define WT(i, j) (uw ? 1.0 : M.wt_val[3 * (size_t)(i) + (j)])
#define WG(i) (uw ? 0 : M.wt_idx[i])
  for (int i = w0; i < w1; i++) {  // src/UpdateForcings.cpp:184-235
    const int g = WG(i);
    double wt = WT(i, 0);
    if (wt != 0.0) { F[RVN_F_PRECIP] += wt * GS(RVN_GF_PRECIP, g); F[RVN_F_SNOW_FRAC] += wt * GS(RVN_GF_SNOW_FRAC, g); ref_elev_precip += wt * GC(g, RVN_GC_ELEVATION); }
    wt = WT(i, 1);
    if (wt != 0.0) {
      temp_month_min += wt * GC(g, RVN_GC_MONTHLY_MIN_TEMP0 + mo - 1); temp_month_max += wt * GC(g, RVN_GC_MONTHLY_MAX_TEMP0 + mo - 1);
      temp_month_ave += wt * GC(g, RVN_GC_MONTHLY_AVE_TEMP0 + mo - 1);
      ref_elev_temp += wt * GC(g, RVN_GC_ELEVATION);
    }
    wt = WT(i, 2);
    if (wt != 0.0) {
      PET_month_ave += wt * GC(g, RVN_GC_MONTHLY_AVE_PET0 + mo - 1);
      F[RVN_F_POTENTIAL_MELT] += wt * GS(RVN_GF_POTENTIAL_MELT, g);
      F[RVN_F_PET] += wt * GS(RVN_GF_PET, g); F[RVN_F_OW_PET] += wt * GS(RVN_GF_OW_PET, g);
    }
  }
  {  // temperature gauge+basin corrections, recomputed over all gauges (:441-455)
    const double tc = M.sb_temp_corr[sb];
    for (int i = w0; i < w1; i++) {
      const int g = WG(i);
      double gauge_corr = tc + GC(g, RVN_GC_TEMP_CORR), wt = WT(i, 1);
      F[RVN_F_TEMP_AVE] += wt * (gauge_corr + GS(RVN_GF_TEMP_AVE, g));
      F[RVN_F_TEMP_DAILY_AVE] += wt * (gauge_corr + GS(RVN_GF_TEMP_DAILY_AVE, g));
      F[RVN_F_TEMP_DAILY_MAX] += wt * (gauge_corr + GS(RVN_GF_TEMP_DAILY_MAX, g));
      F[RVN_F_TEMP_DAILY_MIN] += wt * (gauge_corr + GS(RVN_GF_TEMP_DAILY_MIN, g));
    }
  }
  const double temp_ave_unc = F[RVN_F_TEMP_DAILY_AVE];  // :468
  // CModel::CorrectTemp  src/OrographicCorrections.cpp:24-58
  if (opt.orocorr_temp == RVN_OROCORR_SIMPLELAPSE || opt.orocorr_temp == RVN_OROCORR_HBV) {
    double lapse = c.G(RVN_G_ADIABATIC_LAPSE);
    lapse /= 1000.0;
    F[RVN_F_TEMP_AVE] -= lapse * (elev - ref_elev_temp);
    if (tt.day_changed) {
      F[RVN_F_TEMP_DAILY_AVE] -= lapse * (elev - ref_elev_temp);
      F[RVN_F_TEMP_DAILY_MIN] -= lapse * (elev - ref_elev_temp);
      F[RVN_F_TEMP_DAILY_MAX] -= lapse * (elev - ref_elev_temp);
      temp_month_max -= lapse * (elev - ref_elev_temp);
      temp_month_min -= lapse * (elev - ref_elev_temp);
      if (opt.orocorr_temp != RVN_OROCORR_HBV) temp_month_ave -= lapse * (elev - ref_elev_temp);
    }
  }
  if (M.nPert > 0) apply_perturbations(F, M, RVN_PF_TEMP_AVE, e, k, step);   // src/UpdateForcings.cpp:474
  if (M.daily != nullptr) {
    // sub-daily steps: within a day the daily items are the ones the HRU got at the first step of the day
    // (CHydroUnit::CopyDailyForcings -> CopyDailyForcingItems, src/UpdateForcings.cpp:480-482, src/Forcings.cpp:72-89)
    double *dly = M.daily + (size_t)e * 7 * M.nHp + k;
    if (!tt.day_changed) {
      F[RVN_F_TEMP_DAILY_MIN] = dly[0 * (size_t)M.nHp]; F[RVN_F_TEMP_DAILY_MAX] = dly[1 * (size_t)M.nHp]; F[RVN_F_TEMP_DAILY_AVE] = dly[2 * (size_t)M.nHp];
      temp_month_max = dly[3 * (size_t)M.nHp]; temp_month_min = dly[4 * (size_t)M.nHp]; temp_month_ave = dly[5 * (size_t)M.nHp];
      PET_month_ave = dly[6 * (size_t)M.nHp];
    } else {
      dly[0 * (size_t)M.nHp] = F[RVN_F_TEMP_DAILY_MIN]; dly[1 * (size_t)M.nHp] = F[RVN_F_TEMP_DAILY_MAX]; dly[2 * (size_t)M.nHp] = F[RVN_F_TEMP_DAILY_AVE];
      dly[3 * (size_t)M.nHp] = temp_month_max; dly[4 * (size_t)M.nHp] = temp_month_min; dly[5 * (size_t)M.nHp] = temp_month_ave;
      dly[6 * (size_t)M.nHp] = PET_month_ave;
    }
  }
  F[RVN_F_TEMP_MONTH_AVE] = temp_month_ave; F[RVN_F_PET_MONTH_AVE] = PET_month_ave;
  // CalculateSubDailyCorrection (src/UpdateForcings.cpp:949-1024): 1 for daily steps / SUBDAILY_NONE, else the host-tabulated SUBDAILY_SIMPLE share (interpreter sweep only)
  const double subdaily_corr = (C::kAtmos && M.subdaily != nullptr) ? M.subdaily[(size_t)M.et_row[step] * M.nHp + k] : 1.0;
  {  // CModel::EstimateSnowFraction  src/UpdateForcings.cpp:1068-1197
    double sf = F[RVN_F_SNOW_FRAC];
    const int method = opt.rainsnow;
    if (method == RVN_RAINSNOW_DINGMAN) {
      double temp = c.G(RVN_G_RAINSNOW_TEMP);
      if (F[RVN_F_TEMP_DAILY_MAX] <= temp) sf = 1.0;
      else if (F[RVN_F_TEMP_DAILY_MIN] >= temp) sf = 0.0;
      else sf = (temp - F[RVN_F_TEMP_DAILY_MIN]) / (F[RVN_F_TEMP_DAILY_MAX] - F[RVN_F_TEMP_DAILY_MIN]);
    } else if (method == RVN_RAINSNOW_THRESHOLD) {
      sf = (F[RVN_F_TEMP_AVE] <= c.G(RVN_G_RAINSNOW_TEMP)) ? 1.0 : 0.0;
    } else if (method == RVN_RAINSNOW_HBV || method == RVN_RAINSNOW_UBCWM) {
      double frac, delta = c.G(RVN_G_RAINSNOW_DELTA), temp = c.G(RVN_G_RAINSNOW_TEMP);
      if (F[RVN_F_TEMP_DAILY_AVE] <= (temp - 0.5 * delta)) frac = 1.0;
      else if (F[RVN_F_TEMP_DAILY_AVE] >= (temp + 0.5 * delta)) frac = 0.0;
      else frac = 0.5 + (temp - F[RVN_F_TEMP_DAILY_AVE]) / delta;
      sf = (method == RVN_RAINSNOW_UBCWM) ? frac : frac * (1.0 - F[RVN_F_SNOW_FRAC]) + F[RVN_F_SNOW_FRAC];
    }
    F[RVN_F_SNOW_FRAC] = sf;
  }
  double p5 = 0.0;   // precip_5day: everything that happens to the precipitation happens to it (only kept when a process reads it)
  {  // precipitation gauge+basin corrections (:507-527)
    const double rc = M.sb_rain_corr[sb], sc = M.sb_snow_corr[sb];
    F[RVN_F_PRECIP] = 0.0; p5 = 0.0;
    for (int i = w0; i < w1; i++) {
      const int g = WG(i);
      double gauge_corr = F[RVN_F_SNOW_FRAC] * sc * GC(g, RVN_GC_SNOW_CORR) + (1.0 - F[RVN_F_SNOW_FRAC]) * rc * GC(g, RVN_GC_RAIN_CORR);
      double wt = WT(i, 0);
      F[RVN_F_PRECIP] += wt * gauge_corr * GS(RVN_GF_PRECIP, g);
      if (C::kAtmos && M.gauge_p5 != nullptr) p5 += wt * gauge_corr * M.gauge_p5[(size_t)g * M.nSteps + step];
    }
  }
  // CModel::CorrectPrecip  src/OrographicCorrections.cpp:213-245
  if (opt.orocorr_precip == RVN_OROCORR_SIMPLELAPSE) {
    double lapse = c.G(RVN_G_PRECIP_LAPSE);
    lapse /= 1000;
    if (F[RVN_F_PRECIP] > D_REAL_SMALL) { F[RVN_F_PRECIP] = dmax(F[RVN_F_PRECIP] + lapse * (elev - ref_elev_precip), 0.0); p5 = dmax(p5 + lapse * (elev - ref_elev_precip), 0.0); }
  } else if (opt.orocorr_precip == RVN_OROCORR_HBV) {
    double corr_upper = c.G(RVN_G_HBVEC_LAPSE_UPPER) / 1000.0, corr = c.G(RVN_G_HBVEC_LAPSE_RATE) / 1000.0;
    double lapse_elev = c.G(RVN_G_HBVEC_LAPSE_ELEV), add = 0.0;
    if (elev > lapse_elev) add = (corr_upper - corr) * (elev - lapse_elev);
    F[RVN_F_PRECIP] *= dmax(1.0 + corr * (elev - ref_elev_precip) + add, 0.0);
    p5 *= dmax(1.0 + corr * (elev - ref_elev_precip) + add, 0.0);
  }
  if (M.nPert > 0) {   // src/UpdateForcings.cpp:558-560
    apply_perturbations(F, M, RVN_PF_PRECIP, e, k, step);
    apply_perturbations(F, M, RVN_PF_RAINFALL, e, k, step);
    apply_perturbations(F, M, RVN_PF_SNOWFALL, e, k, step);
  }
  if (c.type == RVN_HRU_MASKED_GLACIER) { F[RVN_F_PRECIP] = 0; p5 = 0; }
  if (C::kAtmos) { c.precip_5day = p5; c.month = mo; }   // interpreter only: INF_SCS / ABST_SCS / CHU_ONTARIO read them
  {  // CModel::EstimatePotentialMelt  src/PotentialMelt.cpp:20-246
    double pm = F[RVN_F_POTENTIAL_MELT];
    const int method = opt.pot_melt;
    if (method == RVN_POTMELT_DEGREE_DAY) {
      pm = c.LU(RVN_LU_MELT_FACTOR) * (F[RVN_F_TEMP_DAILY_AVE] - c.LU(RVN_LU_DD_MELT_TEMP)) * subdaily_corr;
    } else if (method == RVN_POTMELT_HBV || method == RVN_POTMELT_HBV_ROS) {
      double Ma = c.LU(RVN_LU_MELT_FACTOR), Ma_min = c.LU(RVN_LU_MIN_MELT_FACTOR), AM = c.LU(RVN_LU_HBV_MELT_ASP_CORR);
      double MRF = c.LU(RVN_LU_HBV_MELT_FOR_CORR), Fc = c.LU(RVN_LU_FOREST_COVERAGE), melt_temp = c.LU(RVN_LU_DD_MELT_TEMP);
      double slope_corr;
      const int south = (c.lat < 0.0) ? 1 : 0;   // adj = pi
      Ma = Ma_min + (Ma - Ma_min) * 0.5 * (1.0 - __ldg(M.melt_cos + 2 * step + south));   // cos(day_angle - WINTER_SOLSTICE_ANG - adj), host libm
      Ma *= ((1.0 - Fc) * 1.0 + (Fc)*MRF);
      slope_corr = ((1.0 - Fc) * 1.0 + (Fc)*__ldg(M.hru_sin_slope + k));                  // sin(slope)
      Ma *= dmax(1.0 - AM * slope_corr * __ldg(M.hru_cos_aspect + k), 0.0);               // cos(aspect - adj)
      double rain_energy = 0.0;
      if (method == RVN_POTMELT_HBV_ROS)   // rain-on-snow energy, src/PotentialMelt.cpp:77-81
        rain_energy = c.LU(RVN_LU_RAIN_MELT_MULT) * D_SPH_WATER / D_LH_FUSION * dmax(F[RVN_F_TEMP_AVE], 0.0) * (F[RVN_F_PRECIP] * (1.0 - F[RVN_F_SNOW_FRAC]));
      pm = Ma * (F[RVN_F_TEMP_DAILY_AVE] - melt_temp) * subdaily_corr + rain_energy;
    } else if (method == RVN_POTMELT_DD_FREEZE) {   // :40-49
      const double Ma = c.LU(RVN_LU_MELT_FACTOR), melt_temp = c.LU(RVN_LU_DD_MELT_TEMP), Mf = c.LU(RVN_LU_REFREEZE_FACTOR);
      if ((Mf > 0) && (F[RVN_F_TEMP_DAILY_AVE] - melt_temp) < 0.0) pm = Mf * (F[RVN_F_TEMP_DAILY_AVE] - melt_temp);
      else pm = Ma * (F[RVN_F_TEMP_DAILY_AVE] - melt_temp) * subdaily_corr;
    } else if (method == RVN_POTMELT_HMETS) {
      double Ma_min = c.LU(RVN_LU_MIN_MELT_FACTOR), Ma_max = c.LU(RVN_LU_MAX_MELT_FACTOR);
      double melt_temp = c.LU(RVN_LU_DD_MELT_TEMP), Kcum = c.LU(RVN_LU_DD_AGGRADATION);
      double cum_melt = 0.0;
      if (c.iSV(RVN_SV_CUM_SNOWMELT) >= 0) cum_melt = c.SV(c.iSV(RVN_SV_CUM_SNOWMELT));  // still the stored value here
      double Ma = dmin(Ma_max, Ma_min * (1 + Kcum * cum_melt));
      pm = Ma * (F[RVN_F_TEMP_DAILY_AVE] - melt_temp) * subdaily_corr;
    } else if (method == RVN_POTMELT_NONE) {
      pm = 0.0;
    }
    F[RVN_F_POTENTIAL_MELT] = pm;
  }
  // (the atmospheric block sits directly in front of the PET estimate that consumes it: its four results are not live across the melt block)
  // air pressure, relative humidity (src/UpdateForcings.cpp:493-497, 677-731) and incoming shortwave radiation (:591-593;
  // CRadiation::EstimateShortwaveRadiation / ClearSkySolarRadiation / SWCloudCoverCorrection, src/Radiation.cpp:24-57, 992-1038,
  // 333-364) -- only when a consumed PET method needs them.  Geometry-only pieces (ET radiation on the slope and on the flat,
  // optical air mass) come from the host's per-date tables.
  double air_pres = 0.0, rel_hum = 0.0, SW_radia = 0.0, ET_radia = 0.0;
  double SW_radia_net = 0.0, LW_radia_net = 0.0, veg_height = 0.0, veg_LAI = 0.0, veg_SAI = 0.0;   // net radiation and the canopy state behind it (opt.need_radiation)
  c.air_pres = 0.0; c.rel_hum = 0.0; c.wind_vel = 0.0;
  if (C::kAtmos && opt.need_atmos) {
    const double Tave = F[RVN_F_TEMP_AVE];
    if (opt.air_pressure == RVN_AIRPRESS_BASIC) air_pres = D_AMBIENT_AIR_PRESSURE * rvn_pow(1.0 - 0.0065 * elev / (D_ZERO_CELSIUS + Tave), 5.26);
    else if (opt.air_pressure == RVN_AIRPRESS_UBC) air_pres = D_KPA_PER_ATM * (1.0 - 0.0001 * elev);
    else air_pres = D_AMBIENT_AIR_PRESSURE;
    double relhum = 0.5;
    if (opt.rel_humidity == RVN_RELHUM_MINDEWPT) relhum = dmin(sat_vapor_pressure(F[RVN_F_TEMP_DAILY_MIN]) / sat_vapor_pressure(Tave), 1.0);
    rel_hum = dmin(relhum * c.LU(RVN_LU_RELHUM_CORR), 1.0);
    const size_t row = (size_t)M.et_row[step] * M.nHp + k;
    ET_radia = M.et_radia[row];
    if (opt.sw_radiation == RVN_SW_RAD_DEFAULT) {
      const double Ketp = ET_radia, Ket = M.et_flat[row], Mopt = M.opt_air_mass[row];
      double dew_pt;
      { const double a = 17.27, b = 237.7; const double tmp = (a * Tave / (b + Tave)) + rvn_log(rel_hum); dew_pt = b * tmp / (a - tmp); }   // GetDewPointTemp, src/CommonFunctions.cpp:1131-1139
      const double gamma_dust = 0.025;
      const double precip_WV = 1.12 * rvn_exp(0.0614 * dew_pt);
      const double tau = rvn_exp((-0.124 - 0.0207 * precip_WV) + (-0.0682 - 0.0248 * precip_WV) * Mopt) - gamma_dust;                    // Dingman E-9..E-13
      const double tau2 = 0.5 * (1.0 - rvn_exp((-0.0363 - 0.0084 * precip_WV) + (-0.0572 - 0.0173 * precip_WV) * Mopt) + gamma_dust);    // Dingman E-15..E-18
      SW_radia = tau * Ketp + tau2 * Ket + tau2 * D_GLOBAL_ALBEDO * (tau2 + tau) * Ketp;
    }
    double cc = 1.0;   // cloud cover is 0 (CLOUDCOV_NONE)
    if (opt.sw_cloudcorr == RVN_SW_CLOUD_CORR_DINGMAN) cc = 0.355 + 0.68 * (1 - 0.0);
    else if (opt.sw_cloudcorr == RVN_SW_CLOUD_CORR_ANNANDALE) {
      const double kRs = 0.16;
      if (F[RVN_F_TEMP_DAILY_MAX] > F[RVN_F_TEMP_DAILY_MIN]) cc = kRs * (1.0 + 2.7E-5 * elev) * sqrt(F[RVN_F_TEMP_DAILY_MAX] - F[RVN_F_TEMP_DAILY_MIN]);
      else cc = 0.0;
    }
    SW_radia *= cc;
    double wind_vel = 2.0;   // CModel::EstimateWindVelocity  src/UpdateForcings.cpp:739-777
    if (opt.wind_velocity == RVN_WINDVEL_UBC_MOD) {
      const double wt = dmin(0.04 * (F[RVN_F_TEMP_DAILY_MAX] - F[RVN_F_TEMP_DAILY_MIN]), 1.0);
      wind_vel = dmax(0.0, (1 - wt) * c.LU(RVN_LU_MIN_WIND_SPEED) + (wt)*c.LU(RVN_LU_MAX_WIND_SPEED));
    } else if (opt.wind_velocity == RVN_WINDVEL_SQRT) {
      wind_vel = dmax(0.0, c.G(RVN_G_WINDVEL_SCALE) * (F[RVN_F_TEMP_DAILY_MAX] - F[RVN_F_TEMP_DAILY_MIN]) + c.G(RVN_G_WINDVEL_ICEPT));
    }
    c.air_pres = air_pres; c.rel_hum = rel_hum; c.wind_vel = wind_vel;
    // net radiation (src/UpdateForcings.cpp:591-606): sub-canopy shortwave (CRadiation::SWCanopyCorrection, src/Radiation.cpp:377-415), total
    // albedo above and below the canopy (CHydroUnit::GetTotalAlbedo, src/HydroUnits.cpp:728-781; snow albedo = DEFAULT_SNOW_ALBEDO), net
    // longwave LW_RAD_DEFAULT with LW_INC_DEFAULT (src/Radiation.cpp:210-231), cloud cover 0; canopy variables of this step as
    // CVegetationClass::RecalculateCanopyParams derives them (src/VegetationParams.cpp:85-142)
    if (opt.need_radiation) {
      const double rel_ht = M.veg_season[((size_t)step * M.nVeg + c.veg) * 2 + 0], rel_LAI = M.veg_season[((size_t)step * M.nVeg + c.veg) * 2 + 1];
      const double max_height = c.VEG(RVN_V_MAX_HEIGHT), sparseness = c.LU(RVN_LU_FOREST_SPARSENESS), max_LAI = c.VEG(RVN_V_MAX_LAI);
      const double SAI_ht_ratio = c.VEG(RVN_V_SAI_HT_RATIO), extinction = c.VEG(RVN_V_SVF_EXTINCTION), Fc = c.LU(RVN_LU_FOREST_COVERAGE);
      veg_height = max_height * rel_ht;
      veg_LAI = (1.0 - sparseness) * max_LAI * rel_LAI;
      veg_SAI = (1.0 - sparseness) * (veg_height * SAI_ht_ratio);
      const double svf = rvn_exp(-extinction * (veg_LAI + veg_SAI));
      double canopy_corr = 1.0;
      if (opt.sw_canopycorr == 1) canopy_corr = (1.0 - Fc) * 1.0 + Fc * rvn_exp(-extinction * (veg_LAI + veg_SAI));
      const double SW_subcan = SW_radia * canopy_corr;
      double land_albedo = 0.0;
      const double snow_albedo = 0.8, snow_cov = SnowCover(c);
      if (c.type == RVN_HRU_STANDARD) {
        const double soil_sat = dmax(dmin(c.SV(c.soil_slot(0)) / SoilCapacity(c, 0), 1.0), 0.0);   // the stored state: nothing has moved yet this step
        land_albedo = (soil_sat)*c.SOIL(0, RVN_S_ALBEDO_WET) + (1.0 - soil_sat) * c.SOIL(0, RVN_S_ALBEDO_DRY);
      } else if (c.type == RVN_HRU_GLACIER || c.type == RVN_HRU_MASKED_GLACIER) land_albedo = 0.4;
      else if (c.type == RVN_HRU_LAKE) land_albedo = 0.1;
      else if (c.type == RVN_HRU_WATER) land_albedo = 0.1;
      else if (c.type == RVN_HRU_ROCK) land_albedo = 0.35;
      else if (c.type == RVN_HRU_WETLAND) land_albedo = 0.15;
      land_albedo = (snow_cov)*snow_albedo + (1.0 - snow_cov) * land_albedo;
      double veg_albedo = c.VEG(RVN_V_ALBEDO), svf2 = svf, land2 = land_albedo;
      if (veg_albedo < 0) veg_albedo = 0.14;
      if (svf2 > 1.0) svf2 = 0.0;
      if (land2 < 0) land2 = 0.3;
      const double albedo_above = (1.0 - svf2) * (Fc)*veg_albedo + ((svf2) * (Fc) + (1.0 - Fc)) * land2;
      SW_radia_net = SW_radia * (1 - albedo_above);
      const double SW_subcan_net = SW_subcan * (1 - land_albedo);
      if (opt.lw_radiation == 1) LW_radia_net = 0.0;
      else {
        const double emissivity = 0.99, Tair = Tave + D_ZERO_CELSIUS, Tsurf = Tave + D_ZERO_CELSIUS;
        const double ea = rel_hum * sat_vapor_pressure(Tave);
        const double emiss_eff = 1.72 * rvn_pow(ea / Tair, 1 / 7.0);
        double eps_at = emiss_eff * (1.0 + 0.22 * 0.0 * 0.0);
        eps_at = (1 - Fc) * eps_at + (Fc)*1.0;
        const double LW_incoming = 4.90e-9 * eps_at * rvn_pow(Tair, 4.0);
        LW_radia_net = LW_incoming - 4.90e-9 * emissivity * rvn_pow(Tsurf, 4.0);
      }
      // potential melt from the energy balance (CModel::EstimatePotentialMelt, src/PotentialMelt.cpp:103-136; the reference evaluates it after
      // the radiation terms, which is why these two methods sit here and not in the melt block above)
      if (opt.pot_melt == RVN_POTMELT_EB) {   // GetSensibleHeatSnow / GetLatentHeatSnow / GetRainHeatInput, src/SnowParams.cpp:111-190
        const double rainfall = F[RVN_F_PRECIP] * (1 - F[RVN_F_SNOW_FRAC]) + 0.0;
        const double ref_ht = 2.0, roughness = c.LU(RVN_LU_ROUGHNESS), surf_temp = SnowTemperature(c);
        double temp_var = rvn_pow(0.42 / rvn_log(ref_ht / roughness), 2.0);
        const double H = 1.2466 * D_SPH_AIR * temp_var * wind_vel * D_SEC_PER_DAY * (Tave - surf_temp);
        const double vap_pres = sat_vapor_pressure(Tave) * rel_hum, surf_pres = sat_vapor_pressure(surf_temp) * 1.0;
        const double CE = rvn_pow(0.42, 2.0) / rvn_pow(rvn_log(ref_ht / roughness), 2.0);
        temp_var = D_AIR_H20_MW_RAT * 1.2466 * CE / air_pres;
        const double LH = 2.845 * temp_var * wind_vel * D_SEC_PER_DAY * (vap_pres - surf_pres);
        double R = 0;
        const double le = rvn_log(sat_vapor_pressure(Tave) * rel_hum);
        const double rain_temp = (le + 0.4926) / (0.0708 - 0.00421 * le);   // GetDewPointTemp(e), src/CommonFunctions.cpp:1148-1156
        if (surf_temp < D_FREEZING_TEMP) R = D_DENSITY_WATER * (rainfall / D_MM_PER_METER) * (D_SPH_WATER * (rain_temp - D_FREEZING_TEMP) + D_LH_FUSION);
        if (surf_temp >= D_FREEZING_TEMP) R = D_DENSITY_WATER * (rainfall / D_MM_PER_METER) * (D_SPH_WATER * (rain_temp - D_FREEZING_TEMP));
        double S = SW_subcan_net + LW_radia_net + H + LH + R;
        S *= (D_MM_PER_METER / D_DENSITY_WATER / D_LH_FUSION);
        F[RVN_F_POTENTIAL_MELT] = S;
      } else if (opt.pot_melt == RVN_POTMELT_RESTRICTED) {
        const double tmp_rate = dmax(c.LU(RVN_LU_MELT_FACTOR) * (F[RVN_F_TEMP_DAILY_AVE] - c.LU(RVN_LU_DD_MELT_TEMP)), 0.0);
        const double rad = (SW_subcan_net + LW_radia_net);
        const double convert = D_MM_PER_METER / D_DENSITY_WATER / D_LH_FUSION;
        F[RVN_F_POTENTIAL_MELT] = tmp_rate * subdaily_corr + dmax(rad * convert, 0.0);
      }
      if (M.record_forcings) {
        double *gf = M.forc + (size_t)e * RVN_F_COUNT * M.nHp + k;
        gf[(size_t)RVN_F_SW_RADIA_NET * M.nHp] = SW_radia_net; gf[(size_t)RVN_F_LW_RADIA_NET * M.nHp] = LW_radia_net;
        gf[(size_t)RVN_F_POTENTIAL_MELT * M.nHp] = F[RVN_F_POTENTIAL_MELT];
      }
    }
    if (M.record_forcings) {
      M.forc[((size_t)e * RVN_F_COUNT + RVN_F_WIND_VEL) * M.nHp + k] = wind_vel;
      double *gf = M.forc + (size_t)e * RVN_F_COUNT * M.nHp + k;
      gf[(size_t)RVN_F_AIR_PRES * M.nHp] = air_pres; gf[(size_t)RVN_F_REL_HUMIDITY * M.nHp] = rel_hum; gf[(size_t)RVN_F_SW_RADIA * M.nHp] = SW_radia;
    }
  }
  // CModel::EstimatePET  src/Evaporation.cpp:282-540
#pragma unroll
  for (int ow = 0; ow < 2; ow++) {
    const int method = ow ? opt.ow_evaporation : opt.evaporation;
    double PET = ow ? F[RVN_F_OW_PET] : F[RVN_F_PET];
    if (method == RVN_PET_CONSTANT) PET = 3.0;
    else if (method == RVN_PET_NONE) PET = 0.0;
    else if (method == RVN_PET_HARGREAVES_1985) {  // Hargreaves1985Evap  src/Evaporation.cpp:215-228
      if (M.et_radia != nullptr) {
        double Ra = M.et_radia[(size_t)M.et_row[step] * M.nHp + k];   // F->ET_radia [MJ/m2/d] (hoisted CalcETRadiation2)
        Ra /= (D_LH_VAPOR * D_DENSITY_WATER / D_MM_PER_METER);
        double delT = F[RVN_F_TEMP_DAILY_MAX] - F[RVN_F_TEMP_DAILY_MIN];
        delT = dmax(delT, 0.0);
        PET = dmax(0.0, 0.0023 * Ra * sqrt(delT) * (F[RVN_F_TEMP_DAILY_AVE] + 17.8));
      }
    } else if (method == RVN_PET_OUDIN) {   // src/Evaporation.cpp:485-489
      if (M.et_radia != nullptr) {
        const double ET_radia = M.et_radia[(size_t)M.et_row[step] * M.nHp + k];
        PET = dmax(ET_radia / D_DENSITY_WATER / D_LH_VAPOR * D_MM_PER_METER * (F[RVN_F_TEMP_DAILY_AVE] + 5.0) / 100.0, 0.0);
      }
    } else if (method == RVN_PET_HAMON) {   // Hamon1961Evap  src/Evaporation.cpp:263-271; GetSaturatedVaporPressure  src/CommonFunctions.cpp:935-949
      if (M.day_length != nullptr) {
        const double T = F[RVN_F_TEMP_DAILY_AVE], day_length = M.day_length[(size_t)M.et_row[step] * M.nHp + k];
        const double sat_vap = (T >= 0) ? 0.61078 * rvn_exp(17.26939 * T / (T + 237.3)) : 0.61078 * rvn_exp(21.87456 * T / (T + 265.5));
        const double abs_hum = 216.7 * (sat_vap * 10) / (T + 273.16);
        PET = dmax(0.0, 0.0055 * 4.0 * abs_hum * day_length * day_length * 25.4);
      }
    } else if (method == RVN_PET_LINEAR_TEMP) {   // src/Evaporation.cpp:306-311
      PET = c.LU(RVN_LU_PET_LIN_COEFF) * dmax(F[RVN_F_TEMP_DAILY_AVE], 0.0);
    } else if (method == RVN_PET_MOHYSE) {   // src/Evaporation.cpp:476-483
      if (M.sunset_angle != nullptr) {
        const double cpet = c.G(RVN_G_MOHYSE_PET_COEFF), ws = M.sunset_angle[(size_t)M.et_row[step] * M.nHp + k];
        PET = cpet / D_PI * ws * rvn_exp((17.3 * F[RVN_F_TEMP_AVE]) / (238 + F[RVN_F_TEMP_AVE]));
      }
    } else if (C::kAtmos && (method == RVN_PET_PRIESTLEY_TAYLOR || method == RVN_PET_PENMAN_COMBINATION)) {
      const double T = F[RVN_F_TEMP_AVE];
      const double sat_vap = sat_vapor_pressure(T);
      const double de_dT = (T > 0) ? 17.26939 * 237.3 / rvn_sqr(T + 237.3) * sat_vap : 21.87456 * 265.5 / rvn_sqr(T + 265.5) * sat_vap;   // GetSatVapSlope
      const double LH_vapor = 2.495 - 0.002361 * T;
      const double gamma = D_SPH_AIR / D_AIR_H20_MW_RAT * air_pres / LH_vapor;
      if (method == RVN_PET_PRIESTLEY_TAYLOR) {   // PriestleyTaylorEvap  src/Evaporation.cpp:155-168
        PET = c.LU(RVN_LU_PRIESTLEYTAYLOR_COEFF) * (de_dT / (de_dT + gamma)) * dmax(SW_radia_net + LW_radia_net, 0.0) / LH_vapor / D_DENSITY_WATER * D_MM_PER_METER;
      } else {   // PET_PENMAN_COMBINATION  :404-416, PenmanCombinationEvap :119-144, canopy roughness after Brook90 ROUGH (src/VegetationParams.cpp:236-262)
        const double SAI_ht_ratio = c.VEG(RVN_V_SAI_HT_RATIO), lu_rough = c.LU(RVN_LU_ROUGHNESS);
        double closed_roughness, zero_pl, z0;
        if (veg_height >= 10.0) closed_roughness = 0.05 * veg_height;
        else if (veg_height <= 1.0) closed_roughness = 0.13 * veg_height;
        else { closed_roughness = 0.13 * 1.0; closed_roughness += (0.05 * 10.0 - 0.13 * 1.0) * (veg_height - 1.0) / (10.0 - 1.0); }
        const double closed_zerodisp = veg_height - closed_roughness / 0.3;
        if (lu_rough > closed_roughness) closed_roughness = lu_rough;
        const double ratio = (veg_LAI + veg_SAI) / (4.0 + SAI_ht_ratio * veg_height);
        if (ratio >= 1.0) { zero_pl = closed_zerodisp; z0 = closed_roughness; }
        else {
          double xx = 0.0;
          if (veg_height > D_REAL_SMALL) xx = ratio * rvn_pow(-1.0 + rvn_exp(0.909 - 3.03 * closed_zerodisp / veg_height), 4.0);
          zero_pl = 1.1 * veg_height * rvn_log(1.0 + rvn_pow(xx, 0.25));
          z0 = dmin(0.3 * (veg_height - zero_pl), lu_rough + 0.3 * veg_height * sqrt(xx));
        }
        const double ref_ht = veg_height + 2.0;
        double numer = D_AIR_H20_MW_RAT * 1.2466;
        double denom = air_pres * D_DENSITY_WATER * (1.0 / rvn_pow(0.42, 2.0) * (rvn_pow((rvn_log((ref_ht - zero_pl) / z0)), 2.0)));
        const double vert_trans = numer / denom;   // GetVerticalTransportEfficiency  src/CommonFunctions.cpp:1083-1094
        const double vapor_def = sat_vap * (1.0 - (rel_hum));
        numer = de_dT * dmax((SW_radia_net) + (LW_radia_net), 0.0);
        numer += gamma * vert_trans * D_DENSITY_WATER * LH_vapor * (c.wind_vel) * vapor_def;
        denom = D_DENSITY_WATER * LH_vapor * (de_dT + gamma);
        PET = dmax(0.0, (numer / denom)) * D_MM_PER_METER;
      }
    } else if (C::kAtmos && method >= RVN_PET_TURC_1961 && method <= RVN_PET_HARGREAVES) {
     if (method == RVN_PET_TURC_1961) {   // TurcEvap  src/Evaporation.cpp:61-68
      const double t = dmax(0.0, F[RVN_F_TEMP_DAILY_AVE]);
      double evp = 0.013 * t / (t + 15) * ((SW_radia)*23.8846 + 50);
      if (rel_hum < 0.5) evp *= (1 + (50 - rel_hum * 100) / 70);
      PET = dmax(0.0, evp);
    } else if (method == RVN_PET_MAKKINK_1957) {   // Makkink1957Evap  src/Evaporation.cpp:32-48
      const double T = F[RVN_F_TEMP_AVE];
      const double sat_vap = sat_vapor_pressure(T);
      const double de_dT = (T > 0) ? 17.26939 * 237.3 / rvn_sqr(T + 237.3) * sat_vap : 21.87456 * 265.5 / rvn_sqr(T + 265.5) * sat_vap;   // GetSatVapSlope
      const double LH_vapor = 2.495 - 0.002361 * T;
      const double gamma = D_SPH_AIR / D_AIR_H20_MW_RAT * air_pres / LH_vapor;
      PET = dmax(0.0, 0.61 * (de_dT / (de_dT + gamma)) * SW_radia * 23.8846 / 58.5 - 0.12);
    } else if (method == RVN_PET_PENMAN_SIMPLE33 || method == RVN_PET_PENMAN_SIMPLE39) {   // Valiantzas (2006)  src/Evaporation.cpp:454-474
      const double Rs = SW_radia, R_et = ET_radia, Tave = F[RVN_F_TEMP_AVE];
      if (method == RVN_PET_PENMAN_SIMPLE33) PET = 0.047 * Rs * sqrt(Tave + 9.5) - 2.4 * rvn_sqr(Rs / R_et) + 0.09 * (Tave + 20) * (1 - rel_hum);
      else PET = 0.038 * Rs * sqrt(Tave + 9.5) - 2.4 * rvn_sqr(Rs / R_et) + 0.075 * (Tave + 20) * (1 - rel_hum);
    } else if (method == RVN_PET_VAPDEFICIT) {   // src/Evaporation.cpp:506-518
      const double e_sat = sat_vapor_pressure(F[RVN_F_TEMP_AVE]);
      const double ea = rel_hum * e_sat;
      PET = c.LU(RVN_LU_PET_VAP_COEFF) * (e_sat - ea);
    } else if (method == RVN_PET_LINACRE) {   // src/Evaporation.cpp:491-504
      const double T = F[RVN_F_TEMP_DAILY_AVE];
      double Tdewpoint;
      { const double a = 17.27, b = 237.7; const double tmp = (a * T / (b + T)) + rvn_log(rel_hum); Tdewpoint = b * tmp / (a - tmp); }
      const double latit = c.lat * 57.29578951;
      PET = ((ow ? 700 : 500) * T / (100 - latit) + 15.0 * (T - Tdewpoint)) / (80.0 - T);
    } else if (method == RVN_PET_HARGREAVES) {   // HargreavesEvap  src/Evaporation.cpp:179-203
      double Ra = ET_radia;
      Ra /= (D_LH_VAPOR * D_DENSITY_WATER / D_MM_PER_METER);
      double Ct = 0.125;
      if (rel_hum >= 0.54) Ct = 0.035 * rvn_pow(100 * (1.0 - rel_hum), 0.333);
      const double delT = (1.8 * temp_month_max + 32.0) - (1.8 * temp_month_min + 32.0);
      PET = dmax(0.0, 0.0075 * Ra * Ct * sqrt(delT) * (1.8 * F[RVN_F_TEMP_DAILY_AVE] + 32.0));
     }
    } else if (method == RVN_PET_FROMMONTHLY) {
      double peRatio = 1.0 + D_HBV_PET_TEMP_CORR * (temp_ave_unc - temp_month_ave);
      peRatio = dmax(0.0, dmin(2.0, peRatio));
      PET = PET_month_ave * peRatio;
    }
    PET = dmax(PET, 0.0);
    PET = PET * c.VEG(RVN_V_PET_VEG_CORR);
    if (ow) F[RVN_F_OW_PET] = PET; else F[RVN_F_PET] = PET;
  }
  // CModel::CorrectPET  src/OrographicCorrections.cpp:350-440
  if (opt.orocorr_pet == RVN_OROCORR_HBV) {
    double factor = D_GLOBAL_PET_CORR * dmax(1.0 - D_HBV_PET_ELEV_CORR * (elev - ref_elev_temp), 0.0);
    F[RVN_F_PET] *= factor; F[RVN_F_OW_PET] *= factor;
  }
  if (opt.evaporation == RVN_PET_FROMMONTHLY || opt.evaporation == RVN_PET_CONSTANT || opt.evaporation == RVN_PET_HAMON ||
      opt.evaporation == RVN_PET_LINEAR_TEMP || opt.evaporation == RVN_PET_TURC_1961 || opt.evaporation == RVN_PET_LINACRE) F[RVN_F_PET] *= subdaily_corr;   // IsDailyPETmethod, src/Evaporation.cpp:540-556
  if (opt.snow_suppress_PET) {
    double SWE = 0.0, sc = 1.0;
    if (c.iSV(RVN_SV_SNOW) >= 0) SWE = c.old_snow;
    if (c.iSV(RVN_SV_SNOW_COVER) >= 0) sc = c.old_snow_cover;
    if (SWE > 0.1) { F[RVN_F_PET] *= (1.0 - sc); F[RVN_F_OW_PET] *= (1.0 - sc); }
  }
  if (c.type == RVN_HRU_STANDARD) F[RVN_F_PET] *= c.SOIL(0, RVN_S_PET_CORRECTION);
  if (c.type == RVN_HRU_MASKED_GLACIER) F[RVN_F_PET] = F[RVN_F_OW_PET] = 0.0;
  if (opt.direct_evap) {  // src/UpdateForcings.cpp:640-650
    double rainfall = F[RVN_F_PRECIP] * (1.0 - F[RVN_F_SNOW_FRAC]);
    double reduce = dmin(F[RVN_F_PET], rainfall + 0.0);
    double rainfrac = rainfall / (rainfall + 0.0);
    if ((rainfall + 0.0) == 0) rainfrac = 1.0;
    if (F[RVN_F_PRECIP] + 0.0 - reduce > 0.0) F[RVN_F_SNOW_FRAC] = 1.0 - (rainfall - reduce * rainfrac) / (F[RVN_F_PRECIP] - reduce * rainfrac);
    F[RVN_F_PRECIP] -= reduce * rainfrac;
    F[RVN_F_PET] -= reduce * rainfrac;
  }
#undef GS
#undef GC
#undef WT
#undef WG
}

// ------------------------------------------------------------------------------------------------
// CmvConvolution (src/Convolution.cpp:245-303) + the solver's per-connection update of its 2N+1 connections
// (src/Solvers.cpp:220-236) in ONE pass over the N live stores of this HRU: store i is loaded, released, shifted and
// written back.  Only S'[i-1] is carried between rows.  Stores >= N are never touched: their rates are zero in the
// reference as well.  UNIT_DT: tstep == 1.0, where x/tstep and r*tstep are exact identities (daily models).
//
// orig = (sumrem < REAL_SMALL) ? 0 : S/sumrem  (src/Convolution.cpp:280-281).  sumrem depends only on (member, land-use
// class, row), so the host stores it with its correctly rounded reciprocal y = RN(1/sumrem) in the ConvRow table; the
// quotient is then one Markstein correction step  q = S*y; r = fma(-sumrem, q, S); q' = fma(r, y, q)  which returns the
// correctly rounded S/sumrem (the same closing sequence the compiler emits for a/b, without its reciprocal refinement
// and range scaffolding).  y = 0 encodes the reference's guard (q = 0, r = S, q' = 0).  S = 0 (every dry day leaves one
// such store) gives 0 like the division.  Outside 1e-290 < |S| < 1e290 the sequence is only faithful, not correctly
// rounded -- storages are O(1e-12..1e5) mm.
// ------------------------------------------------------------------------------------------------
// x / y for a divisor that is fixed over a loop (the model time step): with yr = RN(1 / y) the same Markstein sequence as conv_orig
// returns the correctly rounded quotient (tests/test_glibc_math.py::test_division_by_constant_is_exact checks it against IEEE
// division for the step lengths in use) in 3 dependent operations instead of the ~25 instructions of a double-precision division.
RVN_DI double div_by(const double x, const double y, const double yr) {
  const double q = __dmul_rn(x, yr);
  const double r = __fma_rn(-y, q, x);
  return __fma_rn(r, yr, q);
}
RVN_DI double conv_orig(const double S, const ConvRow &t) {
  const double q = __dmul_rn(S, t.rcp);
  const double r = __fma_rn(-t.sumrem, q, S);
  return __fma_rn(r, t.rcp, q);
}
RVN_DI ConvRow ld_row(const ConvRow *__restrict__ tab, int i) {
  const double2 a = __ldg(reinterpret_cast<const double2 *>(tab + i));
  ConvRow t; t.uh = a.x; t.rcp = a.y; t.sumrem = __ldg(&tab[i].sumrem); t.pad = 0.0;
  return t;
}

#ifndef RVN_DA_ON
#define RVN_DA_ON 1   // 0 compiles the streamflow-assimilation code out of the routing kernels (layout experiments only)
#endif
#ifndef RVN_CONV_R
#define RVN_CONV_R 4   // rows per register buffer
#endif
#ifndef RVN_PF_DIST
#define RVN_PF_DIST 16  // rows of L2 prefetch run-ahead of the store stream (0 = off)
#endif
RVN_DI void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// warm L2 with the first rows of a convolution's stores (issued at kernel start, long before they are consumed)
#ifndef RVN_PF_HEAD
#define RVN_PF_HEAD RVN_PF_DIST   // rows of each convolution warmed at kernel start
#endif
RVN_DI void conv_prefetch_head(const double *g, const size_t stride, const int n) {
#pragma unroll
  for (int r = 0; r < RVN_PF_HEAD; r++) if (RVN_PF_HEAD <= RVN_PF_DIST || r < n) prefetch_l2(g + (size_t)r * stride);
}

template <bool UNIT_DT, class C>
RVN_DI void convolve_direct(C &c, const int iTarget, const int iTS, double *__restrict__ g /*store 0 of this HRU*/, const size_t stride,
                            const ConvRow *__restrict__ tab, const int N, double *__restrict__ sum_cache, const bool sum_valid) {
  if (N <= 0) return;
  const double tstep = c.tstep;
  const double rstep = UNIT_DT ? 1.0 : 1.0 / tstep;   // correctly rounded reciprocal of the step for div_by
  // ---- the reference's `sum` of the stores in index order.  The same sum, over the same values in the same order, is
  // accumulated while the stores are written at the end of the previous step, so it is normally read back from a
  // one-double-per-HRU cache; only after the state was replaced from outside is it recomputed.
  double sum = 0.0;
  if (sum_valid) sum = *sum_cache;
  else {
    const double *p = g;
    for (int i = 0; i < N; i++) { sum += *p; p += stride; }
  }
  const double TS_old = c.SV(iTS);
  const double added = TS_old - sum;
  double target = c.SV(iTarget), nsum = 0.0, TS_new = 0.0;
  {
    // ---- row 0: receives the new water (connection 0), releases, shifts out unless it is the only store
    const ConvRow t = ld_row(tab, 0);
    const double g0 = __ldcs(g);
    const double r0 = UNIT_DT ? added : div_by(added, tstep, rstep);                         // rates[0]
    const double Sin = g0 + added;                                             // local S[0]
    double act = UNIT_DT ? Sin : g0 + r0 * tstep;                              // store 0 after connection 0
    const double orig = conv_orig(Sin, t);
    const double r_ = UNIT_DT ? t.uh * orig : div_by(t.uh * orig, tstep, rstep);             // rates[1]
    const double rdt = UNIT_DT ? r_ : r_ * tstep;
    const double Sloc = Sin - rdt;                                             // local S[0] after release
    act = act - rdt;                                                           // solver: phinew[store 0] -= r*tstep
    target += rdt;                                                             // solver: phinew[target]  += r*tstep
    if (N == 1) {
      __stcs(g, act);
      nsum = act;
    } else {
      const double mv0 = UNIT_DT ? Sloc : div_by(Sloc, tstep, rstep) * tstep;              // m_0*tstep
      const double fin0 = act - mv0;                                           // connection N+1
      __stcs(g, fin0);
      nsum = fin0;
      double prev_move = mv0, prevS = Sloc;
      // ---- interior rows 1..N-2: every special case is statically decided; the loads of the next RVN_CONV_R rows are in
      // flight while the current ones are processed
      const int end = N - 1;
      double *p = g + stride;
      double cur[RVN_CONV_R];
#pragma unroll
      for (int r = 0; r < RVN_CONV_R; r++) cur[r] = (1 + r < end) ? __ldcs(p + (size_t)r * stride) : 0.0;
      for (int i = 1; i < end; i += RVN_CONV_R) {
        double nxt[RVN_CONV_R];
        const double *pn = p + (size_t)RVN_CONV_R * stride;
#pragma unroll
        for (int r = 0; r < RVN_CONV_R; r++) nxt[r] = (i + RVN_CONV_R + r < end) ? __ldcs(pn + (size_t)r * stride) : 0.0;
        if (RVN_PF_DIST > 0) {
#pragma unroll
          for (int r = 0; r < RVN_CONV_R; r++) if (i + RVN_PF_DIST + r < N) prefetch_l2(p + (size_t)(RVN_PF_DIST - 1 + r) * stride);
        }
#pragma unroll
        for (int r = 0; r < RVN_CONV_R; r++) {
          if (i + r < end) {
            const ConvRow t = ld_row(tab, i + r);
            const double gi = cur[r];
            const double orig = conv_orig(gi, t);
            const double rdt = UNIT_DT ? t.uh * orig : (div_by(t.uh * orig, tstep, rstep)) * tstep;   // rates[i+1]*tstep
            const double Sloc = gi - rdt;                               // local S[i] after release == store i after release
            target += rdt;
            const double mv = UNIT_DT ? Sloc : div_by(Sloc, tstep, rstep) * tstep; // m_i*tstep
            const double fin = (Sloc + prev_move) - mv;                 // connections N+i then N+i+1
            __stcs(p + (size_t)r * stride, fin);
            nsum += fin;
            TS_new += prevS;                                            // local S[i] after the shift loop = S'[i-1]
            prev_move = mv; prevS = Sloc;
          }
        }
#pragma unroll
        for (int r = 0; r < RVN_CONV_R; r++) cur[r] = nxt[r];
        p += (size_t)RVN_CONV_R * stride;
      }
      // ---- last row N-1: releases, receives the shift, keeps what is left
      {
        double *pl = g + (size_t)(N - 1) * stride;
        const ConvRow t = ld_row(tab, N - 1);
        const double gi = __ldcs(pl);
        const double orig = conv_orig(gi, t);
        const double r_ = UNIT_DT ? t.uh * orig : div_by(t.uh * orig, tstep, rstep);
        const double rdt = UNIT_DT ? r_ : r_ * tstep;
        const double Sloc = gi - rdt;
        target += rdt;
        const double fin = Sloc + prev_move;                            // connection 2N-1
        __stcs(pl, fin);
        nsum += fin;
        // reference TS_new = sum_{j=1..N-1} of the local array after its descending shift loop:
        //   local S[j] = S'[j-1] for j<=N-2 and S'[N-1]+S'[N-2] for j=N-1
        TS_new += (Sloc + prevS);
      }
    }
  }
  *sum_cache = nsum;
  c.SV(iTarget) = target;
  const double rTS = UNIT_DT ? (TS_new - TS_old) : div_by(TS_new - TS_old, tstep, rstep);   // rates[2N]
  c.SV(iTS) += UNIT_DT ? rTS : rTS * tstep;                                     // CONVOLUTION[c]: iTo==iFrom, updated in place
}

// ------------------------------------------------------------------------------------------------
// solver pieces shared by both sweeps
// ------------------------------------------------------------------------------------------------
// per-connection update of the newest state (src/Solvers.cpp:220-236)
template <class C, class PR>
RVN_DI void apply_connections(C &c, const PR &pr) {
  const double tstep = c.tstep;
#pragma unroll
  for (int q = 0; q < pr.nconn(); q++) {
    const int from = pr.from(q), to = pr.to(q);
    const double r = c.R(q);
    if (to != from) { c.SVW(from) -= r * tstep; c.SVW(to) += r * tstep; }
    else if (c.slot_water(from) == 1) { c.SVW(to) += 0.0; }   // redirect-to-self: rate forced to zero
    else c.SVW(to) += r * tstep;
  }
}

// one member of a :ProcessGroup (CProcessGroup::GetRatesOfChange, src/ProcessGroup.cpp:72-99): `grp` = the record's group flags.
// The local rate vector R (the reference's loc_rates) is zeroed once per group and NOT between members; every member's
// GetRatesOfChange + ApplyConstraints run on the same state; the group's connections are applied after the last member.
template <class C, class PR>
RVN_DI void run_group_member(C &c, const PR &pr, const int op, const int grp, const int maxConn) {
  if (grp & RVN_GRP_FIRST) {
#pragma unroll
    for (int q = 0; q < maxConn; q++) c.R(q) = 0.0;
  }
  dispatch_rates(c, pr, op);
  const int off = pr.conn0() - pr.gconn0();
#pragma unroll
  for (int q = 0; q < pr.nconn(); q++) c.GR(off + q) = pr.weight() * c.R(q);
  if (grp & RVN_GRP_LAST) {
    const double tstep = c.tstep;
#pragma unroll
    for (int q = 0; q < pr.gnconn(); q++) {   // src/Solvers.cpp:220-236 over the group's concatenated connections
      const int from = pr.gfrom(q), to = pr.gto(q);
      const double r = c.GR(q);
      if (to != from) { c.SVW(from) -= r * tstep; c.SVW(to) += r * tstep; }
      else if (c.slot_water(from) == 1) { c.SVW(to) += 0.0; }
      else c.SVW(to) += r * tstep;
    }
  }
}

RVN_DI double block_sum(double v, double *s_red) {
  // fixed-shape tree: warp shuffles then the (<=4) warp totals in order -> deterministic
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[w] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int i = 0; i < RVN_BS / 32; i++) t += s_red[i];
  return t;
}

template <class C>
RVN_DI void load_hru_statics(C &c, const DevModel &M, int k, const double *s_ptab) {
  c.ptab = s_ptab; c.popt = &M.opt;
  c.tstep = M.opt.timestep;
  c.area = M.hru_area[k]; c.elev = M.hru_elev[k]; c.lat = M.hru_lat[k]; c.slope = M.hru_slope[k]; c.aspect = M.hru_aspect[k];
  { const int t = M.hru_type[k]; c.type = t & 0xff; c.linked = (t & RVN_DEV_RES_LINKED) != 0; }
  c.lu = M.hru_lu[k]; c.veg = M.hru_veg[k]; c.profile = M.hru_profile[k];
  c.glob_off = M.glob_off; c.lu_base = M.lu_off + c.lu * RVN_LU_COUNT; c.veg_base = M.veg_off + c.veg * RVN_V_COUNT; c.soil_off = M.soil_off;
  {
    const int terr = (M.hru_terrain != nullptr) ? M.hru_terrain[k] : -1;
    c.terr_base = M.terr_off + ((terr >= 0) ? terr : 0) * RVN_T_COUNT;   // [NONE] terrain: the host rejects models whose processes read terrain parameters
  }
}

// canopy storage capacities of this HRU for this step: CVegetationClass::RecalculateCanopyParams (src/VegetationParams.cpp:85-135).
// The month-interpolated relative height / LAI of each vegetation class depend on the date only and are hoisted to the
// host (veg_season[step][veg][2]); everything else is evaluated here in the reference's order.
template <class C> RVN_DI void load_canopy(C &c, const DevModel &M, int step, int k) {
  c.canopy_cap = 0.0; c.canopy_snow_cap = 0.0;
  if (C::kAtmos && M.need_cc) {   // what ColdContentBalance reads from the HRU (src/SnowBalance.cpp:745-753)
    const double rel_ht = M.veg_season[((size_t)step * M.nVeg + c.veg) * 2 + 0], rel_LAI = M.veg_season[((size_t)step * M.nVeg + c.veg) * 2 + 1];
    const double sparseness = c.LU(RVN_LU_FOREST_SPARSENESS);
    c.veg_LAI = (1.0 - sparseness) * c.VEG(RVN_V_MAX_LAI) * rel_LAI;
    c.veg_SAI = (1.0 - sparseness) * ((c.VEG(RVN_V_MAX_HEIGHT) * rel_ht) * c.VEG(RVN_V_SAI_HT_RATIO));
    c.day_length = M.day_length[(size_t)M.et_row[step] * M.nHp + k];
  }
  const int icept = C::kAtmos ? c.opt().interception_factor : (int)RVN_ICEPT_USER;
  if (icept == RVN_ICEPT_USER) { c.rain_icept_pct = c.VEG(RVN_V_VAR_RAIN_ICEPT_PCT); c.snow_icept_pct = c.VEG(RVN_V_VAR_SNOW_ICEPT_PCT); }
  else { c.rain_icept_pct = 0.0; c.snow_icept_pct = 0.0; }
  if (c.iSV(RVN_SV_CANOPY) < 0 && c.iSV(RVN_SV_CANOPY_SNOW) < 0 && (icept == RVN_ICEPT_USER || icept == RVN_ICEPT_NONE)) return;
  const double rel_ht = M.veg_season[((size_t)step * M.nVeg + c.veg) * 2 + 0], rel_LAI = M.veg_season[((size_t)step * M.nVeg + c.veg) * 2 + 1];
  const double max_height = c.VEG(RVN_V_MAX_HEIGHT), sparseness = c.LU(RVN_LU_FOREST_SPARSENESS), max_LAI = c.VEG(RVN_V_MAX_LAI);
  const double SAI_ht_ratio = c.VEG(RVN_V_SAI_HT_RATIO);
  const double max_SAI = max_height * SAI_ht_ratio;
  const double height = max_height * rel_ht;
  const double LAI = (1.0 - sparseness) * max_LAI * rel_LAI;
  const double SAI = (1.0 - sparseness) * (height * SAI_ht_ratio);
  if (icept == RVN_ICEPT_LAI) {          // src/VegetationParams.cpp:145-149
    c.rain_icept_pct = dmin(c.VEG(RVN_V_RAIN_ICEPT_FACT) * (LAI + SAI), 1.0);
    c.snow_icept_pct = dmin(c.VEG(RVN_V_SNOW_ICEPT_FACT) * (LAI + SAI), 1.0);
  } else if (icept == RVN_ICEPT_EXPLAI) {  // :155-159
    c.rain_icept_pct = (1.0 - rvn_exp(-0.5 * (LAI + SAI)));
    c.snow_icept_pct = (1.0 - rvn_exp(-0.5 * (LAI + SAI)));
  }
  if ((max_LAI + max_SAI) > 0) {
    c.canopy_cap = ((LAI + SAI) / (max_LAI + max_SAI)) * c.VEG(RVN_V_MAX_CAPACITY);
    c.canopy_snow_cap = ((LAI + SAI) / (max_LAI + max_SAI)) * c.VEG(RVN_V_MAX_SNOW_CAPACITY);
  }
}

// open-water PET and its lake correction factor of a reservoir-linked HRU for this step: what CReservoir::RouteWater / GetAET read through
// _pHRU (src/Reservoir.cpp:1632-1636,1283-1292), handed to the routing kernels in ring slot `slot`
template <class C> RVN_DI void export_reservoir_forcing(C &c, const DevModel &M, const int e, const int k, const int slot) {
  double *rf = M.res_force + (((size_t)slot * M.E + e) * M.nRes + M.hru_res[k]) * 2;
  rf[0] = c.F[RVN_F_OW_PET]; rf[1] = c.LU(RVN_LU_LAKE_PET_CORR);
}
// AET / RUNOFF / GROUNDWATER reboot (src/Solvers.cpp:171-200)
template <class C> RVN_DI void reboot_diagnostics(C &c) {
  if (c.iSV(RVN_SV_AET) >= 0) c.SV(c.iSV(RVN_SV_AET)) = 0.0;
  if (c.iSV(RVN_SV_RUNOFF) >= 0) c.SV(c.iSV(RVN_SV_RUNOFF)) = 0.0;
  if (c.iSV(RVN_SV_GROUNDWATER) >= 0) c.SV(c.iSV(RVN_SV_GROUNDWATER)) = 0.0;
}
// routing prep (src/Solvers.cpp:495-511) and diagnostics (:697-729); returns the surface-water volume [m3]
template <class C> RVN_DI double finish_step(C &c, const int nCore) {
  const int iSW = c.iSV(RVN_SV_SURFACE_WATER);
  const double swv = (c.SV(iSW) / D_MM_PER_METER) * (c.area * D_M2_PER_KM2);
  if (c.iSV(RVN_SV_RUNOFF) >= 0) c.SV(c.iSV(RVN_SV_RUNOFF)) = c.SV(iSW);
  c.SV(iSW) = 0.0;
  if (c.iSV(RVN_SV_TOTAL_SWE) >= 0) {
    double t = 0.0;
#pragma unroll
    for (int s = 0; s < nCore; s++) if (c.slot_type(s) == RVN_SV_SNOW || c.slot_type(s) == RVN_SV_SNOW_LIQ) t += c.SV(s);
    c.SV(c.iSV(RVN_SV_TOTAL_SWE)) = t;
  }
  if (c.iSV(RVN_SV_GLACIER_MB) >= 0) {
    double t = 0.0;
#pragma unroll
    for (int s = 0; s < nCore; s++) {
      const int ty = c.slot_type(s);
      if (ty == RVN_SV_SNOW || ty == RVN_SV_SNOW_LIQ || ty == RVN_SV_FIRN || ty == RVN_SV_GLACIER_ICE || ty == RVN_SV_GLACIER) t += c.SV(s);
    }
    c.SV(c.iSV(RVN_SV_GLACIER_MB)) = t;
  }
  return swv;
}
// this HRU's water storage [mm] over the non-convolution-store variables (the stores are represented by CONVOLUTION[c])
template <class C> RVN_DI double hru_storage(C &c, const int nCore) {
  double stor = 0.0;
#pragma unroll
  for (int s = 0; s < nCore; s++)
    if ((c.slot_water(s) == 1 || c.slot_water(s) == 2) && c.slot_type(s) != RVN_SV_ATMOS_PRECIP) stor += c.SV(s);
  return stor;
}

extern __shared__ __align__(16) unsigned char rvn_smem[];

// ------------------------------------------------------------------------------------------------
// specialised sweep: the process list is the compile-time program Prog
// ------------------------------------------------------------------------------------------------
template <class Prog, int J>
struct PrefetchConvs {
  static RVN_DI void run(const double *gstate, const size_t stride, const int32_t *conv_n) {
    if constexpr (J < Prog::nProc) {
      if constexpr (Prog::op(J) == RVN_OP_CONVOLVE) conv_prefetch_head(gstate + (size_t)Prog::ia(J, 2) * stride, stride, (RVN_PF_HEAD > RVN_PF_DIST) ? conv_n[Prog::ia(J, 0)] : 0);
      PrefetchConvs<Prog, J + 1>::run(gstate, stride, conv_n);
    }
  }
};

template <class Prog, bool UNIT_DT, int J>
struct RunProcs {
  static RVN_DI void run(StatCtx<Prog> &c, const DevModel &M, const bool enabled, const unsigned long long mask, const int e, const int k,
                         double *gstate, const bool sum_valid) {
    if constexpr (J < Prog::nProc) {
      constexpr int op = Prog::op(J);
      const bool apply = enabled && ((mask >> J) & 1ull);   // CModel::ApplyProcess gate (src/Model.cpp:3061)
      if constexpr (op == RVN_OP_CONVOLVE) {
        if (apply) {
          constexpr int ci = Prog::ia(J, 0);
          const size_t slot = ((size_t)e * M.nLU + c.lu) * M.nConv + ci;
          convolve_direct<UNIT_DT>(c, Prog::ia(J, 1), Prog::ia(J, 3), gstate + (size_t)Prog::ia(J, 2) * RVN_TILE, (size_t)RVN_TILE,
                                   M.conv_tab + slot * RVN_CONV_STORES, M.conv_n[slot], M.conv_sum + ((size_t)e * M.nConv + ci) * M.nHp + k,
                                   sum_valid);
        }
      } else {
        if (apply) {
          const StatProc<Prog, J> pr{M.proc_weight};
          if constexpr (Prog::grp(J) != 0) run_group_member(c, pr, op, Prog::grp(J), Prog::maxConn);
          else {
#pragma unroll
            for (int q = 0; q < Prog::nconn(J); q++) c.R(q) = 0.0;          // src/Model.cpp:3066-3071
            dispatch_rates(c, pr, op);
            apply_connections(c, pr);
          }
        }
      }
      RunProcs<Prog, UNIT_DT, J + 1>::run(c, M, enabled, mask, e, k, gstate, sum_valid);
    }
  }
};

#ifndef RVN_STATIC_MIN_BLOCKS
#define RVN_STATIC_MIN_BLOCKS 7   // resident CTAs per SM the specialised sweep is compiled for (caps registers at 72/thread)
#endif
// One launch advances every (member, HRU) pair of the grid by `nsub` consecutive model steps (MULTI; 1 otherwise).  Without lateral
// exchange an HRU's steps depend on nothing but its own state and the shared forcings, so the state vector stays on chip between
// the sub-steps (temporal blocking: the core state is read and written once per launch instead of once per step) and only what
// routing and the balance need per step - the surface-water volume and the block partial sums - goes to the ring slot
// slot0 + sub-step of the hand-off buffers.
template <class Prog, bool UNIT_DT, bool MULTI>
__global__ void __launch_bounds__(RVN_BS, RVN_STATIC_MIN_BLOCKS) hru_sweep_static(const __grid_constant__ DevModel M, const int step0, const int nsub_, const int slot0) {
  double *s_ptab = reinterpret_cast<double *>(rvn_smem);
  double *s_red = s_ptab + ((M.ptab_stride + 1) & ~1);
  const int tid = threadIdx.x, e = blockIdx.y, k = blockIdx.x * RVN_BS + tid;
  {
    const double *prow = M.ptab + (size_t)e * M.ptab_stride;
    for (int i = tid; i < M.ptab_stride; i += RVN_BS) s_ptab[i] = prow[i];
  }
  const size_t nHp = (size_t)M.nHp;
  const bool in_range = (k < M.nH);
  const bool enabled = in_range && (M.hru_enabled[k] != 0);
  StatCtx<Prog> c;
#if RVN_SV_SMEM
  c.svp = s_red + 8 + tid;   // after the reduction scratch
#endif
#if RVN_F_SMEM
  c.F.p = s_red + 8 + (RVN_SV_SMEM ? Prog::nCore * RVN_BS : 0) + tid;
#endif
  load_hru_statics(c, M, k, s_ptab);
  if (UNIT_DT) c.tstep = 1.0;
#pragma unroll
  for (int m = 0; m < Prog::nSoil; m++) {
    c.sbase[m] = M.soil_off + M.prof_class[c.profile * RVN_MAX_SOIL_LAYERS + m] * RVN_S_COUNT;
    c.sthick[m] = M.prof_thick[c.profile * RVN_MAX_SOIL_LAYERS + m];
  }
  const int sb = M.hru_sb[k];
  const unsigned long long mask = M.apply_mask[k];
  // ---- state in: coalesced over k for every state variable (aPhi[k][i] = pHRU->GetStateVarValue(i))
  double *gstate = M.state + RVN_STATE_OFF(M, e, 0, k);   // this thread's column of its tile: state variable i at gstate[i * RVN_TILE]
#pragma unroll
  for (int s = 0; s < Prog::nCore; s++) c.SV(s) = __ldcs(gstate + (size_t)Prog::slot_sv(s) * RVN_TILE);
  if (RVN_PF_DIST > 0 && enabled) PrefetchConvs<Prog, 0>::run(gstate, (size_t)RVN_TILE, M.conv_n + ((size_t)e * M.nLU + c.lu) * M.nConv);
  __syncthreads();   // parameter row staged
  const int nsub = MULTI ? nsub_ : 1;
#pragma unroll 1
  for (int ss = 0; ss < nsub; ss++) {
    const int step = step0 + ss;
    c.old_snow = (Prog::iSV(RVN_SV_SNOW) >= 0) ? c.SV(Prog::iSV(RVN_SV_SNOW) >= 0 ? Prog::iSV(RVN_SV_SNOW) : 0) : 0.0;
    c.old_snow_cover = (Prog::iSV(RVN_SV_SNOW_COVER) >= 0) ? c.SV(Prog::iSV(RVN_SV_SNOW_COVER) >= 0 ? Prog::iSV(RVN_SV_SNOW_COVER) : 0) : 0.0;
    c.old_snow_depth = (Prog::iSV(RVN_SV_SNOW_DEPTH) >= 0) ? c.SV(Prog::iSV(RVN_SV_SNOW_DEPTH) >= 0 ? Prog::iSV(RVN_SV_SNOW_DEPTH) : 0) : 0.0;
    c.old_snow_temp = (Prog::iSV(RVN_SV_SNOW_TEMP) >= 0) ? c.SV(Prog::iSV(RVN_SV_SNOW_TEMP) >= 0 ? Prog::iSV(RVN_SV_SNOW_TEMP) : 0) : 0.0;
    // ---- forcings (on the stored state)
    if (enabled) compute_forcings(c, M, e, k, step, sb);
    if (M.nRes > 0 && enabled && c.linked) export_reservoir_forcing(c, M, e, k, slot0 + ss);
    if (!enabled) {
#pragma unroll
      for (int f = 0; f < RVN_F_CARRIED; f++) c.F[f] = 0.0;
    }
    load_canopy(c, M, step, k);   // (class, date) only: placed after the forcing chain so that its results are not live across it
    if (enabled) reboot_diagnostics(c);   // disabled HRUs keep their stored values (the reference never writes them back, :701)
    // ---- ordered-series process sweep on the newest state
    RunProcs<Prog, UNIT_DT, 0>::run(c, M, enabled, mask, e, k, gstate, (M.conv_sum_valid != 0) || (MULTI && ss > 0));
    // ---- routing prep, diagnostics + per-HRU part of the output reductions
    double swv = 0.0, stor = 0.0;
    if (enabled) {
      swv = finish_step(c, Prog::nCore);
      stor = hru_storage(c, Prog::nCore) * c.area;
      if (!MULTI || ss == nsub - 1) {   // ---- state out
#pragma unroll
        for (int s = 0; s < Prog::nCore; s++) __stcs(gstate + (size_t)Prog::slot_sv(s) * RVN_TILE, c.SV(s));
      }
      if (M.record_forcings) {
        double *gf = M.forc + (size_t)e * RVN_F_COUNT * nHp + k;
#pragma unroll
        for (int f = 0; f < RVN_F_CARRIED; f++) gf[(size_t)f * nHp] = c.F[f];
      }
    }
    const int slot = slot0 + ss;
    if (in_range) M.swvol[((size_t)slot * M.E + e) * nHp + k] = swv;
    const double pa = enabled ? c.F[RVN_F_PRECIP] * c.area : 0.0;
    const double bp = block_sum(pa, s_red);
    const double bs = block_sum(stor, s_red);
    if (tid == 0) {
      M.part_precip[((size_t)slot * M.E + e) * M.nBlocksX + blockIdx.x] = bp;
      M.part_storage[((size_t)slot * M.E + e) * M.nBlocksX + blockIdx.x] = bs;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// interpreter sweep: arbitrary process lists (DevProgram), state/rates in shared-memory columns
// ------------------------------------------------------------------------------------------------
#ifndef RVN_INTERP_MIN_BLOCKS
#define RVN_INTERP_MIN_BLOCKS 4   // caps the interpreter at 128 registers/thread (uncapped it compiles to 253 = 2 resident CTAs/SM; measured on 128 members x 10k HRUs HMETS: 1.47 ms uncapped, 1.11 ms at 3, 0.98 ms at 4)
#endif
__global__ void __launch_bounds__(RVN_BS, RVN_INTERP_MIN_BLOCKS) hru_sweep_interp(const __grid_constant__ DevModel M, const int step0, const int nsub, const int slot0) {
  // ---- shared memory carve-up
  DevProgram *P = reinterpret_cast<DevProgram *>(rvn_smem);
  double *s_ptab = reinterpret_cast<double *>(rvn_smem + ((sizeof(DevProgram) + 15) & ~size_t(15)));
  double *s_sv = s_ptab + ((M.ptab_stride + 1) & ~1);
  double *s_rt = s_sv + (size_t)M.nCore * RVN_BS;
  double *s_grt = s_rt + (size_t)M.maxConn * RVN_BS;
  double *s_svw = s_grt + (size_t)M.maxGroupConn * RVN_BS;                       // EULER only: aPhinew columns
  const bool euler = (M.opt.sol_method == RVN_SOL_EULER);
  const bool heun = (M.opt.sol_method == RVN_SOL_ITERATED_HEUN);   // columns: s_svw = aPhinew, then aPhiPrevIter and rate1
  double *s_prev = s_svw + ((euler || heun) ? (size_t)M.nCore * RVN_BS : 0);
  double *s_r1 = s_prev + (heun ? (size_t)M.nCore * RVN_BS : 0);
  double *s_red = s_r1 + (heun ? (size_t)M.maxConn * RVN_BS : 0);
  const int tid = threadIdx.x, e = blockIdx.y, k = blockIdx.x * RVN_BS + tid;
  {
    const uint32_t *src = reinterpret_cast<const uint32_t *>(M.prog);
    uint32_t *dst = reinterpret_cast<uint32_t *>(P);
    for (int i = tid; i < (int)(sizeof(DevProgram) / 4); i += RVN_BS) dst[i] = src[i];
    const double *prow = M.ptab + (size_t)e * M.ptab_stride;
    for (int i = tid; i < M.ptab_stride; i += RVN_BS) s_ptab[i] = prow[i];
  }
  __syncthreads();
  const int nCore = P->nCore;
  const size_t nHp = (size_t)M.nHp;
  const bool in_range = (k < M.nH);
  const bool enabled = in_range && (M.hru_enabled[k] != 0);
  DynCtx c;
  load_hru_statics(c, M, k, s_ptab);
  c.P = P; c.sv = s_sv + tid; c.rt = s_rt + tid; c.grt = s_grt + tid;
  c.svw = (euler || heun) ? (s_svw + tid) : c.sv;
#if RVN_F_SMEM
  c.F.p = s_red + 8 + tid;
#endif
  c.prof_class = M.prof_class + c.profile * RVN_MAX_SOIL_LAYERS; c.prof_thick = M.prof_thick + c.profile * RVN_MAX_SOIL_LAYERS;
  const int sb = M.hru_sb[k];
  const unsigned long long mask = M.apply_mask[k];
  double *gstate = M.state + RVN_STATE_OFF(M, e, 0, k);
  for (int s = 0; s < nCore; s++) c.SV(s) = __ldcs(gstate + (size_t)P->slot_sv[s] * RVN_TILE);
  double *const sv_read = c.sv;
  for (int ss = 0; ss < nsub; ss++) {   // sub-steps of this launch (see hru_sweep_static); nsub = 1 with lateral processes or EULER
  const int step = step0 + ss;
  c.sv = sv_read;
  c.old_snow = (P->iSV[RVN_SV_SNOW] >= 0) ? c.SV(P->iSV[RVN_SV_SNOW]) : 0.0;
  c.old_snow_cover = (P->iSV[RVN_SV_SNOW_COVER] >= 0) ? c.SV(P->iSV[RVN_SV_SNOW_COVER]) : 0.0;
  c.old_snow_depth = (P->iSV[RVN_SV_SNOW_DEPTH] >= 0) ? c.SV(P->iSV[RVN_SV_SNOW_DEPTH]) : 0.0;
  c.old_snow_temp = (P->iSV[RVN_SV_SNOW_TEMP] >= 0) ? c.SV(P->iSV[RVN_SV_SNOW_TEMP]) : 0.0;
  if (enabled) compute_forcings(c, M, e, k, step, sb);
  else { for (int f = 0; f < RVN_F_CARRIED; f++) c.F[f] = 0.0; }
  if (M.nRes > 0 && enabled && c.linked) export_reservoir_forcing(c, M, e, k, slot0 + ss);
  load_canopy(c, M, step, k);
  if (enabled) reboot_diagnostics(c);
  if (euler || heun) { for (int s = 0; s < nCore; s++) c.SVW(s) = c.SV(s); }   // aPhinew = aPhi (src/Solvers.cpp:155-200); rates come from aPhi
  const bool unit_dt = (c.tstep == 1.0);
  const int nProc = P->nProc;
  if (heun && enabled) {
    // ITERATED_HEUN (src/Solvers.cpp:304-397): every process is evaluated on the start-of-step vector aPhi (rate1) and on the previous
    // iterate aPhiPrevIter (rate2), the iterate aPhinew is rebuilt from aPhi with the mean rate (precipitation keeps rate1), until the
    // mean absolute change of the state variables falls below Options.convergence_crit or max_iterations is reached
    double *phi = c.sv, *nw = c.svw, *pv = s_prev + tid, *r1 = s_r1 + tid;
    const int iAtm = P->iSV[RVN_SV_ATMOS_PRECIP];
    const double tstep = c.tstep;
    int iter = 0;
    bool converg;
    do {
      iter++;
      for (int s = 0; s < nCore; s++) { pv[s * RVN_BS] = nw[s * RVN_BS]; nw[s * RVN_BS] = phi[s * RVN_BS]; }
      for (int j = 0; j < nProc; j++) {
        const rvn_process &pr = P->proc[j];
        if (!((mask >> j) & 1ull)) continue;
        const DynProc dp = {&pr, P};
        c.sv = phi;
        for (int q = 0; q < pr.nconn; q++) c.R(q) = 0.0;
        dispatch_rates(c, dp, pr.op);
        for (int q = 0; q < pr.nconn; q++) r1[q * RVN_BS] = c.R(q);
        c.sv = pv;
        for (int q = 0; q < pr.nconn; q++) c.R(q) = 0.0;
        dispatch_rates(c, dp, pr.op);
        for (int q = 0; q < pr.nconn; q++) {
          const int from = dp.from(q), to = dp.to(q);
          double rg = 0.5 * (r1[q * RVN_BS] + c.R(q));
          if (from == iAtm) rg = r1[q * RVN_BS];
          if (to != from) { nw[from * RVN_BS] -= rg * tstep; nw[to * RVN_BS] += rg * tstep; }
          else if (c.slot_water(from) == 1) { nw[to * RVN_BS] += 0.0; }
          else nw[to * RVN_BS] += rg * tstep;
        }
      }
      double check = 0.0;
      for (int s = 0; s < nCore; s++) check += (fabs(nw[s * RVN_BS] - pv[s * RVN_BS]) / nCore);
      converg = (check <= M.opt.convergence_crit) || (iter == M.opt.max_iterations);
    } while (!converg);
    c.sv = phi;
  }
  for (int j = 0; j < (heun ? 0 : nProc); j++) {
    const rvn_process &pr = P->proc[j];
    const bool apply = enabled && ((mask >> j) & 1ull);   // CModel::ApplyProcess gate (src/Model.cpp:3061)
    if (!apply) continue;
    if (pr.op == RVN_OP_CONVOLVE) {
      const int ci = pr.ia[0];
      const size_t slot = ((size_t)e * M.nLU + c.lu) * M.nConv + ci;
      double *g = gstate + (size_t)pr.ia[2] * RVN_TILE;
      double *sc = M.conv_sum + ((size_t)e * M.nConv + ci) * nHp + k;
      const bool sum_valid = (M.conv_sum_valid != 0) || ss > 0;
      if (unit_dt) convolve_direct<true>(c, pr.ia[1], pr.ia[3], g, (size_t)RVN_TILE, M.conv_tab + slot * RVN_CONV_STORES, M.conv_n[slot], sc, sum_valid);
      else convolve_direct<false>(c, pr.ia[1], pr.ia[3], g, (size_t)RVN_TILE, M.conv_tab + slot * RVN_CONV_STORES, M.conv_n[slot], sc, sum_valid);
      continue;
    }
    const DynProc dp = {&pr, P};
    if (pr.group != 0) {
      run_group_member(c, dp, pr.op, pr.group, P->maxConn);
      if (M.cum_flux != nullptr && (pr.group & RVN_GRP_LAST)) {   // IncrementBalance over the group's weighted connections (src/Solvers.cpp:234)
        for (int q = 0; q < pr.gnconn; q++) {
          const int from = dp.gfrom(q), to = dp.gto(q);
          const double r = (to == from && c.slot_water(from) == 1) ? 0.0 : c.GR(q);
          M.cum_flux[((size_t)e * M.nFlux + pr.gconn0 + q) * nHp + k] += r * c.tstep;
        }
      }
      continue;
    }
    for (int q = 0; q < pr.nconn; q++) c.R(q) = 0.0;          // src/Model.cpp:3066-3071
    dispatch_rates(c, dp, pr.op);
    apply_connections(c, dp);
    if (M.cum_flux != nullptr) {   // pModel->IncrementBalance(qs, k, rates_of_change[q] * tstep): a redirect-to-self of a water store counts zero (src/Solvers.cpp:227-234)
      for (int q = 0; q < pr.nconn; q++) {
        const int from = dp.from(q), to = dp.to(q);
        const double r = (to == from && c.slot_water(from) == 1) ? 0.0 : c.R(q);
        M.cum_flux[((size_t)e * M.nFlux + pr.conn0 + q) * nHp + k] += r * c.tstep;
      }
    }
  }
  c.sv = c.svw;   // from here on "the state" is the updated vector aPhinew
  double swv = 0.0, stor = 0.0;
  const bool defer = (M.defer_finish != 0);   // lateral exchange processes follow: routing prep / diagnostics move to finish_kernel
  if (enabled) {
    if (!defer) {
      swv = finish_step(c, nCore);
      stor = hru_storage(c, nCore) * c.area;
    }
    if (ss == nsub - 1) { for (int s = 0; s < nCore; s++) __stcs(gstate + (size_t)P->slot_sv[s] * RVN_TILE, c.SV(s)); }
    if (M.record_forcings) {
      double *gf = M.forc + (size_t)e * RVN_F_COUNT * nHp + k;
#pragma unroll
      for (int f = 0; f < RVN_F_CARRIED; f++) gf[(size_t)f * nHp] = c.F[f];
    }
  }
  const int slot = slot0 + ss;
  if (in_range && !defer) M.swvol[((size_t)slot * M.E + e) * nHp + k] = swv;
  const double pa = enabled ? c.F[RVN_F_PRECIP] * c.area : 0.0;
  const double bp = block_sum(pa, s_red);
  const double bs = block_sum(stor, s_red);
  if (tid == 0) {
    M.part_precip[((size_t)slot * M.E + e) * M.nBlocksX + blockIdx.x] = bp;
    if (!defer) M.part_storage[((size_t)slot * M.E + e) * M.nBlocksX + blockIdx.x] = bs;
  }
  }   // sub-steps
}

// ------------------------------------------------------------------------------------------------
// routing + balance
// ------------------------------------------------------------------------------------------------
// The inflow / lateral-inflow histories (_aQinHist, _aQlatHist) are ring buffers: the reference shifts every history by one
// slot per step (src/SubBasin.cpp:1531-1553), here only the newest value is written.  After t pushes since the last upload,
// history element n (0 = newest) of a basin with nQ slots lives at slot (n - t) mod nQ; uploads/downloads present the
// reference's ordered arrays (rvn_engine.cu rotates).
struct Ring {
  double *buf; int nQ, base;
  RVN_DI Ring(double *b, int n, int t) : buf(b), nQ(n) { base = n - (t % n); if (base == n) base = 0; }
  RVN_DI double &operator[](int n) const { int i = base + n; if (i >= nQ) i -= nQ; return buf[i]; }
};

// CChannelXSect::GetArea (src/ChannelXSect.cpp:297-302): InterpolateCurve(Q/Q_mult, aQ, aXArea, N, extrapbottom=true)
// (src/CommonFunctions.cpp:1982-2004); the interval is the one with aQ[i] <= x < aQ[i+1]
RVN_DI double channel_area(const DevModel &M, const int p, const double Q) {
  const int N = M.nChanPts;
  const double *xx = M.chan_Q + (size_t)M.sb_channel[p] * N, *y = M.chan_A + (size_t)M.sb_channel[p] * N;
  const double x = Q / M.sb_Qmult[p];
  if (x <= xx[0]) return y[0] + (y[1] - y[0]) / (xx[1] - xx[0]) * (x - xx[0]);
  if (x >= xx[N - 1]) return y[N - 1] + (y[N - 1] - y[N - 2]) / (xx[N - 1] - xx[N - 2]) * (x - xx[N - 1]);
  int lo = 0, hi = N - 1;   // invariant: xx[lo] <= x < xx[hi]
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (x >= xx[mid]) lo = mid; else hi = mid; }
  if (fabs(xx[lo + 1] - xx[lo]) < D_REAL_SMALL) return (y[lo] + y[lo + 1]) / 2;
  return y[lo] + (y[lo + 1] - y[lo]) / (xx[lo + 1] - xx[lo]) * (x - xx[lo]);
}

// ---- ROUTE_DIFFUSIVE_VARY (CSubBasin::UpdateRoutingHydro, src/SubBasin.cpp:2521-2572)
// rvn_erfc of the reference (src/CommonFunctions.cpp:1487-1512): its own rational approximation / asymptotic series, not libm's erfc
RVN_DI double raven_erfc(const double x) {
  const double tmp = fabs(x);
  double fun;
  if (tmp > 3.0) {
    const double f1 = (1.0 - 1.0 / (2.0 * tmp * tmp) + 3.0 / (4.0 * rvn_pow(tmp, 4.0)) - 5.0 / (6.0 * rvn_pow(tmp, 6.0)));
    fun = f1 * rvn_exp(-tmp * tmp) / (tmp * sqrt(D_PI));
  } else {
    const double tmp2 = 1.0 / (1.0 + (0.3275911 * tmp));
    const double tmp3 = 0.254829592 * tmp2 - (0.284496736 * tmp2 * tmp2) + (1.421413741 * rvn_pow(tmp2, 3.0)) - (1.453152027 * rvn_pow(tmp2, 4.0)) + (1.061405429 * rvn_pow(tmp2, 5.0));
    fun = tmp3 * rvn_exp(-tmp * tmp);
  }
  if (tmp == x) return fun;
  return (2 - fun);
}
// TimeVaryingADRCumDist (src/CommonFunctions.cpp:1819-1836): cumulative arrival at distance L of a pulse advected with the celerity history v[]
RVN_DI double tv_adr_cum_dist(const double t, const double L, const double *v, const double D, const double dt) {
  const int i = (int)(floor((t + 0.0001) / dt));
  double tstar = v[i] / v[0] * (t - (double)(i)*dt);
  for (int j = 0; j < i; j++) tstar += dt * v[j] / v[0];
  double F = 0;
  if (t <= 0) return 0.0;
  if (L < 500 * (D / v[0])) F = 0.5 * (rvn_exp((v[0] * L) / D) * raven_erfc((L + v[0] * tstar) / sqrt(4.0 * D * tstar)));
  F += 0.5 * (raven_erfc((L - v[0] * tstar) / sqrt(4.0 * D * tstar)));
  return F;
}
// this step's routing hydrograph of basin p, member e: celerity at the latest inflow Q off the rating curve (CChannelXSect::GetCelerity), pushed
// into the celerity history; rh[n] = increments of the time-varying advection-diffusion arrival curve, normalised
__device__ __noinline__ void update_routing_hydro(const DevModel &M, const int e, const int p, const int nQ, const double Q) {
  const double tstep = M.opt.timestep, L = M.sb_reach_m[p];
  double *c_hist = M.c_hist + ((size_t)e * M.nSB + p) * M.max_nqin, *rh = M.route_hydro_m + ((size_t)e * M.nSB + p) * M.max_nqin;
  double cel = 0.0;
  if (!(Q < 0)) {
    const int N = M.nChanPts;
    const double *aQ = M.chan_Q + (size_t)M.sb_channel[p] * N, *aA = M.chan_A + (size_t)M.sb_channel[p] * N;
    const double Q_mult = M.sb_Qmult[p];
    int i = 0;
    while ((Q > Q_mult * aQ[i + 1]) && (i < (N - 2))) i++;
    cel = Q_mult * (aQ[i + 1] - aQ[i]) / (aA[i + 1] - aA[i]);
  }
  for (int n = nQ - 1; n > 0; n--) c_hist[n] = c_hist[n - 1];   // the reference also writes one slot past the end, which nothing reads
  c_hist[0] = cel * D_SEC_PER_DAY;
  double sum = 0;
  const double alpha = M.sb_diff_alpha[p];
  for (int n = 0; n < nQ; n++) {
    rh[n] = dmax(tv_adr_cum_dist((n)*tstep, L, c_hist, alpha, tstep) - sum, 0.0);
    sum += rh[n];
  }
  for (int n = 0; n < nQ; n++) rh[n] /= sum;
  const double travel_time = L / M.sb_c_ref[p];   // [s] against a step in [d]: the reference's comparison, kept
  if (travel_time < tstep) {
    rh[0] = 1.0 - travel_time / tstep;
    rh[1] = travel_time / tstep;
    for (int n = 2; n < nQ; n++) rh[n] = 0.0;
  }
}

// one basin, one member, one step: HRU -> subbasin aggregation in HRU-index order (src/Solvers.cpp:490-511),
// CSubBasin::UpdateLateralInflow / UpdateInflow, RouteWater (src/SubBasin.cpp:2584-2863), UpdateOutflows (:2320-2421).
// `t` = pushes into the histories before this step.
// the per-basin tables the ordered part of routing chases (global memory, or the member kernel's shared-memory copies)
struct RouteTabs {
  const int32_t *nseg, *nqin, *nqlat, *headwater, *up_ptr, *up_idx;
  const double *musk_K, *musk_X;
  RVN_DI void from_model(const DevModel &M) {
    nseg = M.sb_nseg; nqin = M.sb_nqin; nqlat = M.sb_nqlat; headwater = M.sb_headwater; up_ptr = M.sb_up_ptr; up_idx = M.sb_up_idx;
    musk_K = M.sb_musk_K; musk_X = M.sb_musk_X;
  }
};

// HRU -> subbasin aggregation of one basin (src/Solvers.cpp:490-511) and CSubBasin::UpdateLateralInflow: independent of the routing
// order, so callers may run it for all basins at once.  The surface-water volumes are summed in HRU-index order (bit-identical to
// the reference's loop); the loads of a batch are issued together so that their latencies overlap.
__device__ void basin_lateral_inflow(const DevModel &M, int e, int p, int slot, int t) {
  const double tstep = M.opt.timestep;
  const Ring QlatHist(M.qlat + ((size_t)e * M.nSB + p) * M.max_nqlat, M.sb_nqlat[p], t + 1);
  const double *swvol = M.swvol + ((size_t)slot * M.E + e) * M.nHp;
  double routed = 0.0;
  const int i1 = M.sb_hru_ptr[p + 1];
  int i = M.sb_hru_ptr[p];
  // a reservoir-linked HRU does not drain into the catchment: its surface water is precipitation on the reservoir (src/Solvers.cpp:502-507);
  // here it contributes an exact zero to the sum
  const int r = (M.nRes > 0) ? M.sb_res[p] : -1;
  const int kres = (r >= 0) ? M.res_hru[r] : -1;
  for (; i + 8 <= i1; i += 8) {
    double v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { const int k = M.sb_hru_idx[i + j]; v[j] = (k == kres) ? 0.0 : swvol[k]; }
#pragma unroll
    for (int j = 0; j < 8; j++) routed += v[j];
  }
  for (; i < i1; i++) { const int k = M.sb_hru_idx[i]; routed += (k == kres) ? 0.0 : swvol[k]; }
  QlatHist[0] = routed / (tstep * D_SEC_PER_DAY);
  if (kres >= 0 && M.hru_enabled[kres]) M.res_state[((size_t)e * M.nRes + r) * RVN_RES_STATE + 4] = swvol[kres];   // CReservoir::SetPrecip [m3]
}

// InterpolateCurve (src/CommonFunctions.cpp:1982-2004) on one curve of reservoir r; the interval is the one with xx[i] <= x < xx[i+1]
RVN_DI double res_interp(const double x, const double *xx, const double *y, const int N, const bool extrapbottom) {
  if (x <= xx[0]) {
    if (extrapbottom) return y[0] + (y[1] - y[0]) / (xx[1] - xx[0]) * (x - xx[0]);
    return y[0];
  }
  if (x >= xx[N - 1]) return y[N - 1] + (y[N - 1] - y[N - 2]) / (xx[N - 1] - xx[N - 2]) * (x - xx[N - 1]);
  int lo = 0, hi = N - 1;   // invariant: xx[lo] <= x < xx[hi]
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (x >= xx[mid]) lo = mid; else hi = mid; }
  if (fabs(xx[lo + 1] - xx[lo]) < D_REAL_SMALL) return (y[lo] + y[lo + 1]) / 2;
  return y[lo] + (y[lo + 1] - y[lo]) / (xx[lo + 1] - xx[lo]) * (x - xx[lo]);
}
struct ResCurves {
  const double *stage, *Q, *Qunder, *area, *volume; int N;
  RVN_DI ResCurves(const DevModel &M, const int r) {
    stage = M.res_curves + (size_t)r * RVN_RES_CURVES * M.nResPts; Q = stage + M.nResPts; Qunder = Q + M.nResPts; area = Qunder + M.nResPts; volume = area + M.nResPts;
    N = M.res_np[r];
  }
  RVN_DI double GetVolume(const double ht) const { return res_interp(ht, stage, volume, N, true); }      // src/Reservoir.cpp:1871-1874
  RVN_DI double GetArea(const double ht) const { return res_interp(ht, stage, area, N, false); }          // :1880-1883
  RVN_DI double GetWeirOutflow(const double ht) const {                                                    // :1890-1894 (no weir-height series: adj = 0)
    const double underflow = res_interp(ht, stage, Qunder, N, false);
    return res_interp(ht - 0.0, stage, Q, N, false) + underflow;
  }
};
// CReservoir::RouteWater (src/Reservoir.cpp:1532-1853) for a reservoir without control structures, DZTR, seepage, constraint series or stage
// assimilation, then UpdateStage + UpdateMassBalance (:1264-1323) and the AET hand-back to the linked HRU (src/Solvers.cpp:608-612).
// Level-pool mass balance over the step solved for the new stage with Newton's method on a finite-difference derivative.
__device__ __noinline__ void reservoir_route(const DevModel &M, const int e, const int r, const int slot, const int step, const double Qin_old, const double Qin_new) {
  const ResCurves C(M, r);
  double *rs = M.res_state + ((size_t)e * M.nRes + r) * RVN_RES_STATE;   // stage, stage_last, Qout, Qout_last, precip, MB losses, AET
  const double tstep = M.opt.timestep;
  const double stage = rs[0], Qout = rs[2];
  const int k = M.res_hru[r];
  const double dh = 0.001;
  const double V_old = C.GetVolume(stage), A_old = C.GetArea(stage);
  double ET = 0.0, precip = 0.0, Evap = 0.0;
  if (k >= 0) {
    const double *rf = M.res_force + (((size_t)slot * M.E + e) * M.nRes + r) * 2;
    ET = rf[0] / D_SEC_PER_DAY / D_MM_PER_METER;
    ET *= rf[1];
    Evap = rf[0]; Evap *= rf[1];
    precip = rs[4] / M.opt.timestep / D_SEC_PER_DAY;
  }
  const double seep_old = 0.0;
  const double demand = M.res_extract[(size_t)step * M.nRes + r];
  const double ext_old = demand, ext_new = demand;
  double stage_new, res_outflow;
  const double gamma = V_old + ((Qin_old + Qin_new) - Qout + 2.0 * precip - ET * A_old - seep_old - (ext_old + ext_new)) / 2.0 * (tstep * D_SEC_PER_DAY);
  if (gamma < 0) {   // reservoir dried out
    res_outflow = 0.0; stage_new = M.res_min_stage[r];
  } else {
    double h_guess = stage, relax = 1.0, change;
    int iter = 0;
    do {
      double out = C.GetWeirOutflow(h_guess), out2 = C.GetWeirOutflow(h_guess + dh);
      out += ET * C.GetArea(h_guess) + 0.0 * (h_guess - 0.0);
      out2 += ET * C.GetArea(h_guess + dh) + 0.0 * (h_guess + dh - 0.0);
      const double f = (C.GetVolume(h_guess) + out / 2.0 * (tstep * D_SEC_PER_DAY));
      const double dfdh = ((C.GetVolume(h_guess + dh) + out2 / 2.0 * (tstep * D_SEC_PER_DAY)) - f) / dh;
      change = -(f - gamma) / dfdh;
      if (dfdh == 0) change = 1e-7;
      if (iter > 3) relax *= 0.98;
      h_guess += relax * change;
      iter++;
    } while ((iter < 100) && (fabs(change / relax) > 0.0001));
    stage_new = h_guess;
    res_outflow = C.GetWeirOutflow(stage_new);
    if ((res_outflow < 0.0) || (res_outflow > D_ALMOST_INF)) {   // minimum (0) / maximum flow violated: fix the flow, recompute the volume (:1775-1803)
      if (res_outflow < 0.0) res_outflow = 0.0;
      if (res_outflow > D_ALMOST_INF) res_outflow = D_ALMOST_INF;
      double A_guess = A_old, A_last;
      do {
        double V_new = V_old + ((Qin_old + Qin_new) - (Qout + res_outflow) + 2.0 * precip - ET * (A_old + A_guess) - (seep_old + 0.0) - (ext_old + ext_new)) / 2.0 * (tstep * D_SEC_PER_DAY);
        if (V_new < 0) {
          V_new = 0;
          res_outflow = -2.0 * (V_new - V_old) / (tstep * D_SEC_PER_DAY) + (-Qout + 2.0 * precip + (Qin_old + Qin_new) - ET * (A_old + 0.0) - (ext_old + ext_new));
        }
        stage_new = res_interp(V_new, C.volume, C.stage, C.N, false);
        A_last = A_guess;
        A_guess = C.GetArea(stage_new);
      } while (fabs(1.0 - A_guess / A_last) > 0.00001);
    }
  }
  // UpdateStage / UpdateMassBalance
  rs[1] = stage; rs[0] = stage_new; rs[3] = Qout; rs[2] = res_outflow;
  double losses = 0.0, AET_m3 = 0.0;
  if (k >= 0) {
    const double harea = M.hru_area[k] * D_M2_PER_KM2;
    const double aet = Evap * 0.5 * (C.GetArea(stage_new) + C.GetArea(stage)) / harea;   // CReservoir::GetAET [mm/d], normalised to the HRU area
    AET_m3 = aet * harea / D_MM_PER_METER * tstep;
    if (M.iAET >= 0) M.state[RVN_STATE_OFF(M, e, M.iAET, k)] = aet;
  }
  losses += AET_m3;
  losses += demand * D_SEC_PER_DAY * tstep;   // no series: an exact zero
  rs[5] = losses; rs[6] = AET_m3;
}
// CModel::AssimilationOverride (src/Assimilate.cpp:67-118, DA_ECCC): the gauge's outflow is nudged onto the observation.  Kept out of line so
// that models without assimilation (every benchmark shape) route with the register allocation they had before this existed.
__device__ __noinline__ void da_override(const DevModel &M, const int e, const int p, const int step, double *aQout, const int nSeg) {
  const double tstep = M.opt.timestep;
  const size_t ix = (size_t)step * M.nSB + p;
  double va = 0.0;
  if (M.da_override[ix]) {
    const double Qobs = M.da_obsQ[ix], Qobs2 = M.da_obsQ2[ix], alpha = M.assimilation_fact;
    const double Qmod = aQout[nSeg - 1];
    double adj = 0.0;
    if ((Qmod > 1e-8) && (Qobs != -1.2345)) adj = (Qobs2 == -1.2345) ? alpha * (Qobs - Qmod) : alpha * (0.5 * (Qobs2 + Qobs) - Qmod);
    M.da_adjust[(size_t)e * M.nSB + p] = adj;
    for (int i = 0; i < nSeg; i++) {   // CSubBasin::AdjustAllFlows(adjust, overriding, assimsite), src/SubBasin.cpp:1604-1611
      double q = aQout[i] + adj;
      if (q < 0.0) q = 0.0;
      aQout[i] = q;
      va += adj * tstep * D_SEC_PER_DAY;
    }
  }
  M.da_mass[(size_t)e * M.nSB + p] = va;
}
// the ordered part: CSubBasin::UpdateInflow, RouteWater (src/SubBasin.cpp:2584-2863), UpdateOutflows (:2320-2421); needs the lateral
// inflow of this step (basin_lateral_inflow) and the outflows of the basins above.  `t` = pushes into the histories before this step.
__device__ void basin_route(const DevModel &M, const RouteTabs &T, int e, int p, int t, int step, int slot) {
  const double tstep = M.opt.timestep;
  const int nSeg = T.nseg[p], nQin = T.nqin[p], nQlat = T.nqlat[p];
  const Ring QinHist(M.qin + ((size_t)e * M.nSB + p) * M.max_nqin, nQin, t + 1);
  const Ring QlatHist(M.qlat + ((size_t)e * M.nSB + p) * M.max_nqlat, nQlat, t + 1);
  double *aQout = M.qout + ((size_t)e * M.nSB + p) * RVN_MAX_RIVER_SEGS;
  double *sc = M.scal + ((size_t)e * M.nSB + p) * 8;
  const double *UH = M.sb_unit_hydro + (size_t)p * M.max_nqlat, *RH = M.sb_route_hydro + (size_t)p * M.max_nqin;
  // inflow from upstream basins, accumulated in the reference's processing order (ascending index inside the
  // level above; src/Solvers.cpp:602-606), then CSubBasin::UpdateInflow (src/SubBasin.cpp:1531-1537)
  const int ispec = (M.nSpec > 0) ? M.sb_spec_slot[p] : -1;
  double Qin_up = (ispec >= 0) ? 0.0 + M.spec_in[(size_t)step * M.nSpec + ispec] : 0.0;   // user-specified inflow first (src/Solvers.cpp:543)
  for (int u = T.up_ptr[p]; u < T.up_ptr[p + 1]; u++) {
    const int pu = T.up_idx[u];
    const int ru = (M.nRes > 0) ? M.sb_res[pu] : -1;   // CSubBasin::GetOutflowRate: the reservoir's outflow where there is one (src/SubBasin.cpp:844-848)
    Qin_up += (ru >= 0) ? M.res_state[((size_t)e * M.nRes + ru) * RVN_RES_STATE + 2] : M.qout[((size_t)e * M.nSB + pu) * RVN_MAX_RIVER_SEGS + T.nseg[pu] - 1];
  }
  QinHist[0] = Qin_up;
  const double Qin0 = Qin_up, Qin1 = QinHist[1], Qlat0 = QlatHist[0];
  // CSubBasin::RouteWater  src/SubBasin.cpp:2584-2863
  double Qlat_new = 0.0;
  for (int n = 0; n < nQlat; n++) Qlat_new += UH[n] * QlatHist[n];
  int method = M.opt.routing;
  if (T.headwater[p]) method = RVN_ROUTE_NONE;
  const double seg_fraction = 1.0 / (double)(nSeg);
  const double QoutLast = aQout[nSeg - 1];
  if (method == RVN_ROUTE_MUSKINGUM || method == RVN_ROUTE_MUSKINGUM_CUNGE) {
    const double K = T.musk_K[p], X = T.musk_X[p];
    const double cunge = (method == RVN_ROUTE_MUSKINGUM_CUNGE) ? 1.0 : 0.0;
    double dt = dmin(K, tstep);
    // aQout[] is advanced in place over the local sub-steps; the reference restores the stored array and then
    // overwrites it with the final values in UpdateOutflows, which is the same end state.
    double dt_c = -1.0, c1 = 0.0, c2 = 0.0, c3 = 0.0, c4 = 0.0;   // the coefficients change only when dt does (the last, shorter sub-step)
    for (double tt = 0; tt < tstep; tt += dt) {
      if (dt > (tstep - tt)) dt = tstep - tt;
      if (dt != dt_c) {
        const double denom = (2 * K * (1.0 - X) + dt);
        c1 = (dt - 2 * K * (X)) / denom; c2 = (dt + 2 * K * (X)) / denom; c3 = (-dt + 2 * K * (1 - X)) / denom; c4 = (dt) / denom;
        dt_c = dt;
      }
      double Qin = Qin1 + ((tt) / tstep) * (Qin0 - Qin1);
      double Qin_new = Qin1 + ((tt + dt) / tstep) * (Qin0 - Qin1);
      for (int s = 0; s < nSeg; s++) {
        const double old = aQout[s];
        const double qn = c1 * Qin_new + c2 * Qin + c3 * old + cunge * c4 * (Qlat_new * seg_fraction);
        Qin = old; Qin_new = qn;
        aQout[s] = qn;
      }
    }
  } else if (method == RVN_ROUTE_STORAGECOEFF) {
    double Qin_new = Qin0, Qin = Qin1;
    const double Qlocal = sc[2];
    for (int s = 0; s < nSeg; s++) {
      const double ttime = M.sb_K_reach[p] * seg_fraction;
      const double stc = dmin(1.0 / (ttime / tstep + 0.5), 1.0);
      const double c1 = stc / 2, c2 = stc / 2, c3 = (1.0 - stc), corr = 1.0;
      const double old = aQout[s];
      const double qn = c1 * Qin + c2 * Qin_new + c3 * (old - corr * Qlocal);
      Qin = old; Qin_new = qn;
      aQout[s] = qn;
    }
  } else if (method == RVN_ROUTE_HYDROLOGIC) {
    // level-pool routing of the reach with Newton's method on the rating curve (src/SubBasin.cpp:2685-2741): solve
    // V(Q) + Q/2*dt = V_old + (Qin_old + Qin_new - Qout_old)/2*dt for the new outflow; finite-difference derivative, relaxation
    // after 3 and 10 iterations, at most 20 iterations
    const double dts = tstep * D_SEC_PER_DAY, L = M.sb_reach_m[p];
    const double Qout_old = aQout[nSeg - 1] - sc[2];
    const double V_old = channel_area(M, p, Qout_old) * L;
    int iter = 0;
    double change = 0;
    const double gamma = V_old + (Qin1 + Qin0 - Qout_old) / 2.0 * dts;
    const double dQ = 0.1;
    double Q_guess = Qout_old, relax = 1.0;
    do {
      const double f = (channel_area(M, p, Q_guess) * L + (Q_guess) / 2.0 * dts);
      const double dfdQ = (channel_area(M, p, Q_guess + dQ) * L + (Q_guess + dQ) / 2.0 * dts - f) / dQ;
      change = -(f - gamma) / dfdQ;
      if (dfdQ == 0.0) change = 1e-7;
      Q_guess += relax * change;
      if (Q_guess < 0) { Q_guess = 0; change = 0.0; }
      iter++;
      if (iter > 3) relax = 0.9;
      if (iter > 10) relax = 0.7;
    } while ((iter < 20) && (fabs(change) > 0.0001));
    aQout[nSeg - 1] = Q_guess;
  } else if (method == RVN_ROUTE_DIFFUSIVE_VARY) {
    // CSubBasin::UpdateSubBasin runs before any basin is routed (src/Solvers.cpp:552): the celerity comes from the inflow of the previous step
    update_routing_hydro(M, e, p, nQin, Qin1);
    const double *rh = M.route_hydro_m + ((size_t)e * M.nSB + p) * M.max_nqin;
    double q = 0.0;
    for (int n = 0; n < nQin; n++) q += rh[n] * QinHist[n];
    aQout[nSeg - 1] = q;
  } else if (method == RVN_ROUTE_PLUG_FLOW || method == RVN_ROUTE_DIFFUSIVE_WAVE) {
    double q = 0.0;
    for (int n = 0; n < nQin; n++) q += RH[n] * QinHist[n];
    aQout[nSeg - 1] = q;
  } else {  // ROUTE_NONE
    for (int s = 0; s < nSeg; s++) aQout[s] = 0.0;
    aQout[nSeg - 1] = Qin0;
  }
  aQout[nSeg - 1] += Qlat_new;
  aQout[nSeg - 1] += (ispec >= 0) ? M.spec_down[(size_t)step * M.nSpec + ispec] + 0.0 : 0.0;  // down_Q = specified downstream inflow + return flows (none)
  if (M.nRes > 0 && M.sb_res[p] >= 0)   // src/Solvers.cpp:591-596: inflow to the reservoir = channel outflow at the start and at the end of the step
    reservoir_route(M, e, M.sb_res[p], slot, step, QoutLast, dmax((aQout[nSeg - 1] - 0.0 - 0.0), 0.0));
  // CSubBasin::UpdateOutflows  src/SubBasin.cpp:2320-2421 (no irrigation or diversions)
  sc[0] = QoutLast;
  const double irrQ = dmin(0.0, aQout[nSeg - 1]); aQout[nSeg - 1] -= irrQ;
  const double divQ = dmin(0.0, aQout[nSeg - 1]); aQout[nSeg - 1] -= divQ;
  const double dts = tstep * D_SEC_PER_DAY;
  sc[3] = sc[2]; sc[2] = Qlat_new;
  double dV = 0.0;
  dV += 0.5 * (Qin0 + Qin1) * dts;
  dV -= 0.5 * (aQout[nSeg - 1] + QoutLast) * dts;
  dV += 0.5 * (Qlat_new + sc[1]) * dts;
  dV -= 0.5 * (irrQ + 0.0) * dts;
  dV -= 0.5 * (divQ + 0.0) * dts;
  sc[4] += dV;
  dV = 0.0;
  dV += Qlat0 * dts;
  dV -= 0.5 * (Qlat_new + sc[1]) * dts;
  sc[5] += dV;
  sc[1] = Qlat_new;
  if (RVN_DA_ON && M.assimilate_flow) da_override(M, e, p, step, aQout, nSeg);
}
// CModel::AssimilationBackPropagate (src/Assimilate.cpp:283-339), first loop: a basin without a valid observation takes the correction of the
// basin below scaled by the ratio of drainage areas; downstream to upstream, i.e. level after level starting at the outlets
RVN_DI void da_propagate(const DevModel &M, const int e, const int p, const int step) {
  if (M.da_override[(size_t)step * M.nSB + p]) return;
  const int pdown = M.sb_down[p];
  double *adj = M.da_adjust + (size_t)e * M.nSB;
  adj[p] = (pdown >= 0) ? adj[pdown] * (M.sb_drain_area[p] / M.sb_drain_area[pdown]) : 0.0;
}
// second and third loop: distance decay and drainage-area weight of the propagated correction, then CSubBasin::AdjustAllFlows(adjust, false, override)
// (src/SubBasin.cpp:1585-1603): the inflow history (all but the newest entry) and, away from the gauges, the reach outflows move by the correction
__device__ void da_apply(const DevModel &M, const int e, const int p, const int step, const int t) {
  const size_t ix = (size_t)step * M.nSB + p;
  double *padj = M.da_adjust + (size_t)e * M.nSB + p;
  const bool site = M.da_override[ix] != 0;
  if (!site) {
    double a = *padj;
    if (M.sb_down[p] >= 0) a *= M.da_decay[ix];
    a = a * M.da_wt[ix];
    *padj = a;
  }
  const double adjust = *padj;
  const int nQin = M.sb_nqin[p], nSeg = M.sb_nseg[p];
  const Ring QinHist(M.qin + ((size_t)e * M.nSB + p) * M.max_nqin, nQin, t + 1);
  for (int n = 1; n < nQin; n++) {
    double q = QinHist[n] + adjust * (M.sb_drain_area[p] - M.sb_area[p]) / M.sb_drain_area[p] * M.sb_da_corr[p];
    if (q < 0.0) q = 0.0;
    QinHist[n] = q;
  }
  if (!site) {
    double *aQout = M.qout + ((size_t)e * M.nSB + p) * RVN_MAX_RIVER_SEGS;
    for (int i = 0; i < nSeg; i++) { double q = aQout[i] + adjust; if (q < 0.0) q = 0.0; aQout[i] = q; }
  }
}

// ---- routing runs as a short sequence of small kernels on its own stream (one thread per (member, basin)):
//   route_level_kernel x nLevels  aggregation + level-ordered routing, deepest level first; basins inside a level are independent
//   balance_kernel                IncrementCumulInput / IncrementCumOutflow + watershed storage + per-step records
// Small CTAs (64 threads x <=32 registers): routing CTAs slip into the register-file and thread slots the resident sweep
// CTAs leave free on an SM instead of displacing them (measured: 128 threads x 72 registers cost the sweep 0.2 ms/step)
#ifndef RVN_ROUTE_BS
#define RVN_ROUTE_BS 64
#endif
#ifndef RVN_ROUTE_MINB
#define RVN_ROUTE_MINB 32
#endif
#define RVN_ROUTE_LB ((RVN_ROUTE_BS) < 64 ? 64 : (RVN_ROUTE_BS))   // register budget of a 64-thread CTA (<= 32 registers) also when CTAs of one warp are launched
// HRU -> subbasin aggregation + lateral inflow of every (member, basin): no order among basins, one full-width launch per step, so that the
// per-level launches below carry only the ordered part of the chain
__global__ void __launch_bounds__(RVN_ROUTE_LB, RVN_ROUTE_MINB) lateral_all_kernel(const __grid_constant__ DevModel M, const int slot, const int t) {
  const long long idx = (long long)blockIdx.x * RVN_ROUTE_BS + threadIdx.x;
  if (idx >= (long long)M.E * M.nSB) return;
  const int e = (int)(idx / M.nSB), p = (int)(idx - (long long)e * M.nSB);
  basin_lateral_inflow(M, e, p, slot, t);
}
__global__ void __launch_bounds__(RVN_ROUTE_LB, RVN_ROUTE_MINB) route_level_kernel(const __grid_constant__ DevModel M, const int lev0, const int nlev, const int slot, const int t, const int step) {
  const long long idx = (long long)blockIdx.x * RVN_ROUTE_BS + threadIdx.x;
  if (idx >= (long long)M.E * nlev) return;
  const int e = (int)(idx / nlev), i = (int)(idx - (long long)e * nlev);
  const int p = M.level_idx[lev0 + i];
  RouteTabs T; T.from_model(M);
  basin_route(M, T, e, p, t, step, slot);
}
// the narrow levels towards the outlets (at most RVN_TAIL_MAX (member, basin) pairs each) in ONE launch: one CTA per member walks them in order with a CTA
// barrier in between.  A level of a handful of basins costs a full kernel's latency chain when it is launched on its own; a 50 000-basin
// tree has 11 such levels out of 16.
#ifndef RVN_TAIL_MAX
#define RVN_TAIL_MAX 2048
#endif
#define RVN_TAIL_BS 256
#ifndef RVN_TAIL_BS_NARROW
#define RVN_TAIL_BS_NARROW 32   // large ensembles: every tail level has a handful of basins per member - one warp per member, 32 registers (see RVN_ROUTE_BS)
#endif
template <int BS>
__global__ void __launch_bounds__(BS < 64 ? 64 : BS, BS < 64 ? 32 : 4) route_tail_kernel(const __grid_constant__ DevModel M, const int L0, const int slot, const int t, const int step) {
  const int e = blockIdx.x;
  RouteTabs T; T.from_model(M);
  for (int L = L0; L < M.nLevels; L++) {
    const int lev0 = M.level_ptr[L], lev1 = M.level_ptr[L + 1];
    for (int i = lev0 + threadIdx.x; i < lev1; i += BS) basin_route(M, T, e, M.level_idx[i], t, step, slot);
    __syncthreads();   // the next level reads this level's outflows (global memory, same CTA)
  }
}
__global__ void __launch_bounds__(RVN_ROUTE_BS) da_propagate_kernel(const __grid_constant__ DevModel M, const int lev0, const int nlev, const int step) {
  const long long idx = (long long)blockIdx.x * RVN_ROUTE_BS + threadIdx.x;
  if (idx >= (long long)M.E * nlev) return;
  const int e = (int)(idx / nlev), i = (int)(idx - (long long)e * nlev);
  da_propagate(M, e, M.level_idx[lev0 + i], step);
}
__global__ void __launch_bounds__(RVN_ROUTE_BS) da_apply_kernel(const __grid_constant__ DevModel M, const int step, const int t) {
  const long long idx = (long long)blockIdx.x * RVN_ROUTE_BS + threadIdx.x;
  if (idx >= (long long)M.E * M.nSB) return;
  const int e = (int)(idx / M.nSB), p = (int)(idx - (long long)e * M.nSB);
  da_apply(M, e, p, step, t);
}
// volume the assimilation overrides added (src/Assimilate.cpp:115-116): input when positive, negative output otherwise; few gauges, one thread
__device__ __noinline__ void da_book_mass(const DevModel &M, const int e, double *cum, const double area) {
  for (int p = 0; p < M.nSB; p++) {
    const double va = M.da_mass[(size_t)e * M.nSB + p];
    if (va > 0.0) cum[0] += va / area * D_MM_PER_METER; else cum[1] -= va / area * D_MM_PER_METER;
  }
}
// IncrementCumulInput / IncrementCumOutflow (src/Model.cpp:2458-2539) + watershed storage + the step's records for member e; executed
// by all NT threads of a CTA (s_red: NT/32 x 5 doubles of shared memory)
template <int NT>
RVN_DI void balance_member(const DevModel &M, const int e, const int slot, const int rec, const int step, double (*s_red)[5]) {
  const int tid = threadIdx.x, NB = M.nSB;
  const double tstep = M.opt.timestep;
  const double area = M.watershed_area * D_M2_PER_KM2;
  double v[5] = {0.0, 0.0, 0.0, 0.0, 0.0};   // outflow, channel storage, rivulet storage, precip partials, storage partials
  for (int p = tid; p < NB; p += NT) {
    const double *sc = M.scal + ((size_t)e * NB + p) * 8;
    double qo = M.qout[((size_t)e * NB + p) * RVN_MAX_RIVER_SEGS + M.sb_nseg[p] - 1];
    const int r = (M.nRes > 0) ? M.sb_res[p] : -1;
    if (r >= 0) {   // CSubBasin::GetOutflowRate / GetIntegratedOutflow / GetReservoirLosses / GetReservoirStorage with a reservoir (src/SubBasin.cpp:803-807,844-848,902-906,929-934)
      const double *rs = M.res_state + ((size_t)e * M.nRes + r) * RVN_RES_STATE;
      qo = rs[2];
      if (M.sb_level[p] == 0 || M.sb_down[p] < 0) v[0] += (0.5 * (dmax(rs[2] + 0.0, 0.0) + dmax(rs[3] + 0.0, 0.0)) * (tstep * D_SEC_PER_DAY)) / area * D_MM_PER_METER;
      v[0] += rs[5] / area * D_MM_PER_METER;
      v[1] += ResCurves(M, r).GetVolume(rs[0]);   // reservoir storage rides with the channel storage in the watershed record
    } else if (M.sb_level[p] == 0 || M.sb_down[p] < 0) v[0] += (0.5 * (qo + sc[0]) * (tstep * D_SEC_PER_DAY)) / area * D_MM_PER_METER;
    v[1] += sc[4]; v[2] += sc[5];
    if ((M.rec_flags & 1) && rec >= 0) M.rec_qout[((size_t)rec * M.E + e) * NB + p] = qo;
  }
  const double *part_p = M.part_precip + ((size_t)slot * M.E + e) * M.nBlocksX, *part_s = M.part_storage + ((size_t)slot * M.E + e) * M.nBlocksX;
  for (int b = tid; b < M.nBlocksX; b += NT) { v[3] += part_p[b]; v[4] += part_s[b]; }
  // fixed-shape tree (warp shuffles, then the warp totals in order): deterministic for a given NT
#pragma unroll
  for (int w = 0; w < 5; w++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[w] += __shfl_xor_sync(0xffffffffu, v[w], o);
    if ((tid & 31) == 0) s_red[tid >> 5][w] = v[w];
  }
  __syncthreads();
  if (tid == 0) {
    double t[5];
#pragma unroll
    for (int w = 0; w < 5; w++) { t[w] = 0.0; for (int i = 0; i < NT / 32; i++) t[w] += s_red[i][w]; }
    const double avg_precip = t[3] / M.watershed_area;
    double *cum = M.cumul + (size_t)e * 2;
    cum[0] += avg_precip * tstep;
    for (int i = 0; i < M.nSpec; i++) cum[0] += M.spec_cum[(size_t)step * M.nSpec + i];   // GetIntegratedSpecInflow, basin after basin (src/Model.cpp:2467-2469)
    cum[1] += t[0];
    if (RVN_DA_ON && M.assimilate_flow) da_book_mass(M, e, cum, area);
    if ((M.rec_flags & 2) && rec >= 0) {
      double *w = M.rec_wshed + ((size_t)rec * M.E + e) * 4;
      w[0] = cum[0]; w[1] = cum[1]; w[2] = avg_precip;
      w[3] = t[4] / M.watershed_area + t[1] / area * D_MM_PER_METER + t[2] / area * D_MM_PER_METER;
    }
  }
}
__global__ void __launch_bounds__(RVN_ROUTE_LB, RVN_ROUTE_MINB) balance_kernel(const __grid_constant__ DevModel M, const int slot, const int rec, const int step) {
  __shared__ double s_red[RVN_ROUTE_BS / 32][5];
  balance_member<RVN_ROUTE_BS>(M, blockIdx.x, slot, rec, step, s_red);
}

__global__ void __launch_bounds__(1024) balance_kernel_wide(const __grid_constant__ DevModel M, const int slot, const int rec, const int step) {
  __shared__ double s_red[1024 / 32][5];
  balance_member<1024>(M, blockIdx.x, slot, rec, step, s_red);
}

// ---- small ensembles: one CTA per member routes the whole network - every level in order with a CTA barrier between levels - and
// closes the balance, for `nsub` consecutive steps in ONE launch (the per-level form needs nLevels + 1 launches per step, which is
// what a single-member model spends its time on).  The sums of the balance use the same fixed-shape tree as balance_kernel but over
// RVN_MEMBER_BS threads: their last bits may differ from the per-level path (records are compared at 1e-9).
#ifndef RVN_MEMBER_BS
#define RVN_MEMBER_BS 512
#endif
#ifndef RVN_MEMBER_STAGE
#define RVN_MEMBER_STAGE 1024
#endif
__global__ void __launch_bounds__(RVN_MEMBER_BS) route_member_kernel(const __grid_constant__ DevModel M, const int nsub, const int slot0, const int t0, const int rec0, const int step0) {
  __shared__ double s_red[RVN_MEMBER_BS / 32][5];
  // networks of up to RVN_MEMBER_STAGE basins: the index / parameter tables the level loop chases are staged in shared memory once
  // per launch (a level of a deep network is a chain of dependent loads; from global memory that is ~8 L2 round trips per level)
  __shared__ int32_t s_i[7][RVN_MEMBER_STAGE + 1];
  __shared__ double s_d[2][RVN_MEMBER_STAGE];
  const int e = blockIdx.x, NB = M.nSB;
  RouteTabs T; T.from_model(M);
  const int32_t *level_idx = M.level_idx;
  if (NB <= RVN_MEMBER_STAGE) {
    for (int p = threadIdx.x; p < NB; p += RVN_MEMBER_BS) {
      s_i[0][p] = M.sb_nseg[p]; s_i[1][p] = M.sb_nqin[p]; s_i[2][p] = M.sb_nqlat[p]; s_i[3][p] = M.sb_headwater[p]; s_i[6][p] = M.level_idx[p];
      s_d[0][p] = M.sb_musk_K ? M.sb_musk_K[p] : 0.0; s_d[1][p] = M.sb_musk_X ? M.sb_musk_X[p] : 0.0;
    }
    for (int p = threadIdx.x; p <= NB; p += RVN_MEMBER_BS) s_i[4][p] = M.sb_up_ptr[p];
    const int nup = M.sb_up_ptr[NB];   // <= NB - 1 edges in a tree
    for (int u = threadIdx.x; u < nup; u += RVN_MEMBER_BS) s_i[5][u] = M.sb_up_idx[u];
    __syncthreads();
    T.nseg = s_i[0]; T.nqin = s_i[1]; T.nqlat = s_i[2]; T.headwater = s_i[3]; T.up_ptr = s_i[4]; T.up_idx = s_i[5]; T.musk_K = s_d[0]; T.musk_X = s_d[1];
    level_idx = s_i[6];
  }
  for (int ss = 0; ss < nsub; ss++) {
    for (int p = threadIdx.x; p < NB; p += RVN_MEMBER_BS) basin_lateral_inflow(M, e, p, slot0 + ss, t0 + ss);   // no order among basins
    __syncthreads();
    for (int L = 0; L < M.nLevels; L++) {
      const int lev0 = M.level_ptr[L], lev1 = M.level_ptr[L + 1];
      for (int i = lev0 + threadIdx.x; i < lev1; i += RVN_MEMBER_BS) basin_route(M, T, e, level_idx[i], t0 + ss, step0 + ss, slot0 + ss);
      __syncthreads();   // the next level reads this level's outflows
    }
    if (RVN_DA_ON && M.assimilate_flow) {   // AssimilationBackPropagate: outlets first, one level at a time, then the corrections are applied everywhere
      for (int L = M.nLevels - 1; L >= 0; L--) {
        const int lev0 = M.level_ptr[L], lev1 = M.level_ptr[L + 1];
        for (int i = lev0 + threadIdx.x; i < lev1; i += RVN_MEMBER_BS) da_propagate(M, e, level_idx[i], step0 + ss);
        __syncthreads();
      }
      for (int p = threadIdx.x; p < NB; p += RVN_MEMBER_BS) da_apply(M, e, p, step0 + ss, t0 + ss);
      __syncthreads();
    }
    balance_member<RVN_MEMBER_BS>(M, e, slot0 + ss, rec0 >= 0 ? rec0 + ss : -1, step0 + ss, s_red);
    __syncthreads();
  }
}

// CalculateDiagnostic (src/Diagnostics.cpp) of the recorded hydrograph of one basin, one thread per member, sums in the reference's order
__global__ void diagnostic_kernel(const __grid_constant__ DevModel M, const int type, const int basin, const double *__restrict__ obs, const int nobs,
                                  const int nnstart, const int nnend, const double *__restrict__ q0, const int member0, const int nmembers,
                                  const int nrec, double *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nmembers) return;
  const int e = member0 + i, NB = M.nSB, nmod = nrec + 1;
  const double dts = M.opt.timestep * D_SEC_PER_DAY;
  auto modv = [&](int nn) -> double {   // period-averaged outflow at sample nn
    if (nn == 0) return q0[i];
    const double Q = M.rec_qout[((size_t)(nn - 1) * M.E + e) * NB + basin];
    const double Ql = (nn >= 2) ? M.rec_qout[((size_t)(nn - 2) * M.E + e) * NB + basin] : q0[i];
    return (0.5 * (Q + Ql) * dts) / dts;
  };
  auto wt = [&](int nn) -> double { return (obs[nn] == RVN_BLANK || nn >= nmod) ? 0.0 : 1.0; };
  auto mv = [&](int nn) -> double { return (nn < nmod) ? modv(nn) : 0.0; };
  double N = 0.0, res = -D_ALMOST_INF;
  if (type == 0) {
    double avgobs = 0.0;
    for (int nn = nnstart; nn < nnend; nn++) { avgobs += wt(nn) * obs[nn]; N += wt(nn); }
    if (N > 0.0) avgobs /= N;
    double sum1 = 0.0, sum2 = 0.0;
    for (int nn = nnstart; nn < nnend; nn++) { const double w = wt(nn), m = mv(nn); sum1 += w * rvn_sqr(obs[nn] - m); sum2 += w * rvn_sqr(obs[nn] - avgobs); }
    if (N > 0) res = 1.0 - (sum1 / sum2);
  } else if (type == 1) {
    double sum = 0.0;
    for (int nn = nnstart; nn < nnend; nn++) { const double w = wt(nn); sum += w * rvn_sqr(obs[nn] - mv(nn)); N += w; }
    if (N > 0.0) res = sqrt(sum / N);
  } else {
    double ObsSum = 0, ModSum = 0, ObsStd = 0, ModStd = 0, Cov = 0;
    for (int nn = nnstart; nn < nnend; nn++) { const double w = wt(nn); ObsSum += obs[nn] * w; ModSum += mv(nn) * w; N += w; }
    const double ObsAvg = ObsSum / N, ModAvg = ModSum / N;
    for (int nn = nnstart; nn < nnend; nn++) {
      const double w = wt(nn), m = mv(nn);
      ObsStd += rvn_sqr(obs[nn] - ObsAvg) * w;
      ModStd += rvn_sqr(m - ModAvg) * w;
      Cov += (obs[nn] - ObsAvg) * (m - ModAvg) * w;
    }
    ObsStd = sqrt(ObsStd / N); ModStd = sqrt(ModStd / N); Cov /= N;
    const double r = Cov / ObsStd / ModStd, Beta = ModAvg / ObsAvg, Alpha = ModStd / ObsStd;
    if ((N > 0) && ((ObsAvg != 0.0) || (Beta == 1.0)) && (ObsStd != 0.0) && (ModStd != 0.0)) res = 1.0 - sqrt(rvn_sqr(r - 1) + rvn_sqr(Alpha - 1) + rvn_sqr(Beta - 1));
  }
  out[i] = res;
}

// area-weighted watershed average of every state variable (CModel::GetAvgStateVar, src/Model.cpp:1159-1174)
__global__ void avg_state_kernel(const __grid_constant__ DevModel M, double *out, int member0) {
  const int e = member0 + blockIdx.y, i = blockIdx.x, tid = threadIdx.x;
  __shared__ double s_red[256];
  double v = 0.0;
  for (int k = tid; k < M.nH; k += 256) if (M.hru_enabled[k]) v += M.state[RVN_STATE_OFF(M, e, i, k)] * M.hru_area[k];
  s_red[tid] = v;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (tid < o) s_red[tid] += s_red[tid + o]; __syncthreads(); }
  if (tid == 0) out[(size_t)blockIdx.y * M.NS + i] = s_red[0] / M.watershed_area;
}

// plain [member][field][nHp] rows (the recorded forcings) -> [member][hru][field] (host order)
__global__ void transpose_rows_out_kernel(const double *__restrict__ src, double *__restrict__ dst, int nH, int nHp, int NS) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x, e = blockIdx.y;
  if (k >= nH) return;
  for (int i = 0; i < NS; i++) dst[((size_t)e * nH + k) * NS + i] = src[((size_t)e * NS + i) * nHp + k];
}
// [member][hru][sv] (host order) <-> the tiled device order (RVN_STATE_OFF) of members member0 + blockIdx.y
__global__ void transpose_in_kernel(const double *__restrict__ src, const __grid_constant__ DevModel M, const int member0) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x, e = blockIdx.y;
  if (k >= M.nH) return;
  for (int i = 0; i < M.NS; i++) M.state[RVN_STATE_OFF(M, member0 + e, i, k)] = src[((size_t)e * M.nH + k) * M.NS + i];
}
__global__ void transpose_out_kernel(const __grid_constant__ DevModel M, double *__restrict__ dst, const int member0) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x, e = blockIdx.y;
  if (k >= M.nH) return;
  for (int i = 0; i < M.NS; i++) dst[((size_t)e * M.nH + k) * M.NS + i] = M.state[RVN_STATE_OFF(M, member0 + e, i, k)];
}
#endif
