// rvn_device.cuh -- device-side data model of the B200 engine (sm_100a, FP64 throughout).
//
// HBM layout (all FP64 unless noted; nHp = nHRU rounded up to the 128-thread CTA so that every row is a whole number of 1 KB CTA tiles):
//   state   [member][sv][nHp]        SoA: a warp reading one state variable of 32 neighbouring HRUs
//                                    touches 256 contiguous bytes (8 full sectors)
//   swvol   [2][member][nHp]         surface-water volume handed to in-catchment routing [m3] (double-buffered by step parity
//                                    so that routing of step s can overlap the HRU sweep of step s+1)
//   forc    [member][RVN_F_COUNT][nHp]  derived forcings (written only when forcing output is enabled)
//   ptab    [member][ptab_stride]    parameter rows (globals | land-use | vegetation | soil classes)
//   conv_tab[member][nLU][nConv][50] hoisted convolution unit hydrographs as ConvRow records, conv_n = their lengths
//   basin   qin[member][nSB][max_nqin] qlat[..][max_nqlat] qout[..][50] scal[..][8]
// Static, member-independent tables (HRU attributes, class handles, gauge series/weights, network
// CSR, calendar) are stored once and shared by all members.
#ifndef RVN_DEVICE_CUH
#define RVN_DEVICE_CUH
#include <stdint.h>
#include "rvn_gpu.h"

#define RVN_BS 128            // threads per CTA of the HRU sweep: one (member, HRU) pair per thread
#define RVN_MAX_CORE 40       // max non-CONV_STOR state variables
#define RVN_MAX_PROC 64
#define RVN_MAX_PROG_CONN 256
#define RVN_MAX_CONVPROC 4

// One ordinate of a hoisted convolution unit hydrograph (reference CmvConvolution::GenerateUnitHydrograph,
// src/Convolution.cpp:107-154) together with the two member/class constants the release of store i needs
// (src/Convolution.cpp:278-283): sumrem = 1 - sum(UH[0..i-1]) accumulated in the reference's order, and its correctly
// rounded reciprocal (0 where the reference's guard `sumrem < REAL_SMALL` sets the release to zero).
struct __align__(32) ConvRow { double uh, rcp, sumrem, pad; };

// process program + catalogue, read by the interpreter kernel (uniform reads -> broadcasts)
struct DevProgram {
  int32_t nProc, nCore, NS, nSoil;
  rvn_options opt;
  rvn_process proc[RVN_MAX_PROC];          // SV references translated to core slots (see below)
  int16_t conn_from[RVN_MAX_PROG_CONN];     // core slots
  int16_t conn_to[RVN_MAX_PROG_CONN];
  int16_t slot_sv[RVN_MAX_CORE];            // core slot -> catalogue index
  int8_t  slot_type[RVN_MAX_CORE];          // sv_type of slot
  int8_t  slot_layer[RVN_MAX_CORE];
  int8_t  slot_water[RVN_MAX_CORE];         // 1: water storage where a self-connection is a no-op; 2: CONVOLUTION; 0: other
  int16_t iSV[RVN_SV_NTYPES];               // first core slot of each SV type or -1
  int16_t soil_slot[RVN_MAX_SOIL_LAYERS];   // core slot of SOIL[m]
  int32_t maxConn, nConn, nConv, maxGroupConn;
};

// State layout in HBM: tiles of RVN_TILE (= one sweep CTA's) HRUs, state[member][tile][sv][RVN_TILE].  A CTA's rows are consecutive 1 KB
// lines, the row stride is a compile-time constant (row addresses of the specialised sweeps become immediate offsets), and the store
// stream of a CTA stays inside one ~100 KB region instead of touching one 2 MB page per row when a member has a million HRUs (with rows
// [sv][nHp] the stride of BASELINE config 4 was 8 MB: the sweep ran at 0.25 of the HBM peak, TLB / DRAM-page bound).
#define RVN_TILE RVN_BS
#define RVN_STATE_OFF(M, e, sv, k) ((((size_t)(e) * (M).nTiles + (size_t)((k) / RVN_TILE)) * (M).NS + (size_t)(sv)) * RVN_TILE + (size_t)((k) % RVN_TILE))
#define RVN_DEV_RES_LINKED 0x100   // device-side flag in DevModel::hru_type[]: the HRU is linked to a reservoir (CHydroUnit::IsLinkedToReservoir)
struct DevModel {
  int32_t nH, nHp, nTiles, nSB, E, NS, nCore, maxConn, maxGroupConn;   // nHp = nTiles * RVN_TILE
  rvn_options opt;
  int32_t glob_off, lu_off, veg_off, soil_off, ptab_stride;
  int32_t nLU, nVeg, nConv, nGauges, nSteps, record_forcings;
  // per-member arrays
  double *state, *swvol, *forc, *ptab;             // state: tiled, see RVN_STATE_OFF
  ConvRow *conv_tab;
  int32_t *conv_n;
  double *conv_sum;                               // [member][nConv][nHp] sum of each HRU's live convolution stores
  int32_t conv_sum_valid;                         // 0 after the state/parameters were replaced from the host
  double *qin, *qlat, *qout, *scal, *cumul;      // cumul [member][2]
  double *part_precip, *part_storage;             // [2][member][nBlocksX] block partial sums of the sweep (step parity)
  int32_t nBlocksX;
  // shared static tables
  const double *hru_area, *hru_elev, *hru_lat, *hru_slope, *hru_aspect;
  const int32_t *hru_type, *hru_lu, *hru_veg, *hru_profile, *hru_sb, *hru_enabled;
  const int32_t *hru_terrain; int32_t terr_off;   // terrain class of each HRU (or NULL) and its block in the parameter table
  const unsigned long long *apply_mask;
  const int32_t *prof_class; const double *prof_thick;
  const double *gauge_series;                      // [step][nGauges][RVN_GF_COUNT]: one contiguous slab per step, station-major
  const double *gauge_const;
  const int32_t *wt_ptr, *wt_idx; const double *wt_val;   // CSR gauge/grid-cell weights (rvn_gpu.h)
  int32_t unit_weights;                            // 1: one station, every HRU weight exactly 1.0 (CSR not read)
  const rvn_time *times;
  // trigonometric factors of POTMELT_HBV (src/PotentialMelt.cpp:60-75) evaluated on the host with the libm the reference links
  // (CUDA's sin / cos differ from it in the last place now and then): per step cos(day_angle - WINTER_SOLSTICE_ANG - adj) for
  // adj = 0 (northern) and pi (southern hemisphere) [nSteps][2]; per HRU sin(slope) and cos(aspect - adj)
  const double *melt_cos; const double *hru_sin_slope, *hru_cos_aspect;
  const double *veg_season;                        // [nSteps][nVeg][2] month-interpolated relative height, relative LAI
  const double *et_flat, *opt_air_mass;            // [nEtRows][nHp] flat-surface ET radiation and optical air mass (or NULL)
  const double *day_length, *sunset_angle;         // [nEtRows][nHp] day geometry for PET_HAMON / PET_MOHYSE (or NULL)
  const double *et_radia; const int32_t *et_row;   // [nEtRows][nHp] extraterrestrial radiation on the HRU's slope, row of each step (or NULL)
  // network
  const int32_t *sb_down, *sb_level, *sb_headwater, *sb_nseg, *sb_nqin, *sb_nqlat;
  const int32_t *sb_hru_ptr, *sb_hru_idx;          // CSR: HRUs of each basin in HRU-index order
  const int32_t *sb_up_ptr, *sb_up_idx;            // CSR: upstream basins, ascending index
  const int32_t *level_ptr, *level_idx;            // basins grouped by level, deepest level first, ascending index inside
  int32_t nLevels, max_nqin, max_nqlat;
  const double *sb_area, *sb_rain_corr, *sb_snow_corr, *sb_temp_corr, *sb_musk_K, *sb_musk_X, *sb_K_reach;
  const double *sb_unit_hydro, *sb_route_hydro;
  const double *chan_Q, *chan_A, *sb_Qmult, *sb_reach_m; const int32_t *sb_channel; int32_t nChanPts;   // ROUTE_HYDROLOGIC rating curves
  double watershed_area;
  // user-specified basin inflows (rvn_model_desc::spec_*): slot of each basin in the [nSteps][nSpec] tables or -1
  int32_t nSpec; const int32_t *sb_spec_slot; const double *spec_in, *spec_down, *spec_cum;
  // recording
  double *rec_qout, *rec_wshed;                    // [rec][member][nSB], [rec][member][4]
  int32_t rec_capacity, rec_flags;
  const DevProgram *prog;
  double *daily;                                   // sub-daily steps only: [member][7][nHp] daily forcing items carried through the day (CopyDailyForcingItems)
  int32_t defer_finish;                            // 1: lateral processes follow the sweep; routing prep / diagnostics run in finish_kernel
  const double *proc_weight;                       // [nProc] process-group member weights (1.0 outside groups)
  // forcing perturbations of the EnKF members (0 = none)
  int32_t nPert;
  int32_t pert_forcing[RVN_MAX_PERT], pert_adj[RVN_MAX_PERT];
  const uint8_t *pert_mask;                        // [nPert][nHp] or NULL
  const double *pert_eps;                          // [member][nSteps][nPert]
  // ---- fields only the routing kernels read are kept behind everything the sweeps read: the sweeps' constant-bank offsets do not move when
  // routing features are added (a 5 % difference of the headline sweep was measured between two layouts, profiles/r02/devmodel_layout.md)
  // ROUTE_DIFFUSIVE_VARY: per-member routing hydrographs [member][nSB][max_nqin] and celerity histories [member][nSB][max_nqin] rebuilt every step
  const double *sb_c_ref, *sb_diff_alpha; double *route_hydro_m, *c_hist;
  // reservoirs (rvn_model_desc::res_*): hru_type[] carries RVN_DEV_RES_LINKED for a reservoir-linked HRU, hru_res[k] = its reservoir or -1;
  // res_state [member][nRes][RVN_RES_STATE]; res_force [slot][member][nRes][2] = {OW_PET, LAKE_PET_CORR} of the linked HRU, written by the sweep
  int32_t nRes, nResPts; const int32_t *sb_res, *res_hru, *res_np, *hru_res; const double *res_curves, *res_min_stage, *res_extract;
  double *res_state, *res_force; int32_t iAET;
  // direct-insertion streamflow assimilation (rvn_model_desc::da_*): da_adjust [member][nSB] = _aDAQadjust, da_mass [member][nSB] = volume the
  // override added to basin p this step [m3] (booked by the balance)
  int32_t assimilate_flow; double assimilation_fact;
  const uint8_t *da_override; const double *da_obsQ, *da_obsQ2, *da_decay, *da_wt, *sb_drain_area, *sb_da_corr;
  double *da_adjust, *da_mass;
  const double *subdaily;   // [nEtRows][nHp] force_struct::subdaily_corr (rvn_model_desc::subdaily_corr) or NULL (= 1)
  const double *gauge_p5;   // [nGauges][nSteps] 5-day antecedent precipitation of the gauges (rvn_model_desc::gauge_precip5) or NULL
  double *cum_flux; int nFlux;   // [member][nFlux = nConn][nHp] cumulative per-connection fluxes (opt.keep_flux_balance) or NULL
  int need_cc;   // the process list holds SNOBAL_COLD_CONTENT: the sweep derives LAI / SAI / day length of every HRU (CtxCommon::veg_LAI ...)
};
#endif
