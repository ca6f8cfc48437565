// rvn_lateral.cuh -- lateral exchange between HRUs (reference CModel::ApplyLateralProcess, src/Model.cpp:3126-3165, applied by
// the solver after the vertical processes of ALL HRUs, src/Solvers.cpp:404-443) and the end-of-step pieces the sweep defers
// when lateral processes exist.
//
//   lateral_flush_kernel  CmvLatFlush::GetLateralExchange (src/LatFlush.cpp:204-234) or CmvLatEquilibrate::GetLateralExchange
//                         (src/LatEquilibrate.cpp:168-212) + the solver's sequential application
//                         aPhinew[kFrom][iFrom] -= rate/Afrom*tstep; aPhinew[kTo][iTo] += rate/Ato*tstep (src/Solvers.cpp:424-430).
//                         All rates of a process are computed from the same post-vertical state, then applied in connection
//                         order.  Connections are grouped into chunks whose HRUs no other chunk touches (one subbasin each:
//                         the reference only builds within-subbasin pairs unless INTERBASIN), so one thread owns one
//                         (member, chunk) and replays that chunk's connections in the reference's order -> bit-identical.
//   finish_kernel         routing prep (src/Solvers.cpp:495-511), TOTAL_SWE / GLACIER_MB (:697-729) and the per-HRU part of the
//                         watershed-storage reduction, i.e. what hru_sweep_interp does at its end when no lateral process exists.
#ifndef RVN_LATERAL_CUH
#define RVN_LATERAL_CUH
#include "rvn_kernels.cuh"

struct DevLateral {
  int32_t sv_from, sv_to, nconn, nchunks;
  const int32_t *kfrom, *kto, *chunk_ptr;
  const double *frac;
  double *rates;             // scratch [member][nconn]
  int32_t to_type, to_layer; // sv_type / layer of the receiving state variable (CHydroUnit::GetStateVarMax, src/HydroUnits.cpp:559-606)
  int32_t sv_snow;           // index of SNOW (for a SNOW_LIQ recipient) or -1
  int32_t kind;              // rvn_lateral_kind
  const int32_t *flag;       // LAT_EQUILIBRATE control-flow flags per connection
  double mixing_rate;        // LAT_EQUILIBRATE [1/d]; LAT_REDISTRIBUTE: max snow height [mm]
  int32_t method;            // LAT_REDISTRIBUTE: rvn_redist_method
  double tan80;              // tan(80 deg) from the host's libm (the reference's TAN_80_DEGREES)
};

// GetStateVarMax of the recipient for the state-variable types a flush may target here; CANOPY / CANOPY_SNOW recipients (date
// dependent capacities) are rejected at rvn_gpu_create
RVN_DI double lateral_max_storage(const DevModel &M, const DevLateral &L, const int e, const int k) {
  const double *ptab = M.ptab + (size_t)e * M.ptab_stride;
  if (L.to_type == RVN_SV_SOIL) {
    const int prof = M.hru_profile[k], m = L.to_layer;
    const int cls = M.prof_class[prof * RVN_MAX_SOIL_LAYERS + m];
    return M.prof_thick[prof * RVN_MAX_SOIL_LAYERS + m] * D_MM_PER_METER * ptab[M.soil_off + cls * RVN_S_COUNT + RVN_S_CAP_RATIO];
  }
  if (L.to_type == RVN_SV_DEPRESSION) { const double dm = ptab[M.lu_off + M.hru_lu[k] * RVN_LU_COUNT + RVN_LU_DEP_MAX]; return (dm >= 0) ? dm : D_ALMOST_INF; }
  if (L.to_type == RVN_SV_SNOW_COVER) return 1.0;
  if (L.to_type == RVN_SV_SNOW_LIQ) return ptab[M.glob_off + RVN_G_SNOW_SWI] * M.state[RVN_STATE_OFF(M, e, L.sv_snow, k)];
  return D_ALMOST_INF;
}

// one state variable of one member across HRUs in the tiled state layout
struct TiledRow {
  double *base; const DevModel *M; int e, sv;
  RVN_DI double &operator[](const int k) const { return base[RVN_STATE_OFF(*M, e, sv, k)]; }
};
__global__ void lateral_flush_kernel(const __grid_constant__ DevModel M, const DevLateral L) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)M.E * L.nchunks) return;
  const int e = (int)(idx / L.nchunks), c = (int)(idx - (long long)e * L.nchunks);
  const double tstep = M.opt.timestep;
  const TiledRow from_row{M.state, &M, e, L.sv_from}, to_row{M.state, &M, e, L.sv_to};   // row[k] = state variable of HRU k (tiled layout)
  double *rates = L.rates + (size_t)e * L.nconn;
  const int q0 = L.chunk_ptr[c], q1 = L.chunk_ptr[c + 1];
  if (L.kind == RVN_LAT_EQUILIBRATE) {   // CmvLatEquilibrate::GetLateralExchange, src/LatEquilibrate.cpp:168-212; one chunk = one subbasin's pool
    const double mix = dmin(L.mixing_rate * tstep, 1.0);
    double sum = 0.0;
    for (int q = q0; q < q1; q++) {
      const int kF = L.kfrom[q], kT = L.kto[q], flag = L.flag[q];
      const double stor = from_row[kF], to_stor = to_row[kT];
      const double Afrom = M.hru_area[kF], Ato = M.hru_area[kT];
      if (flag & 1) sum += mix * dmax(to_stor, 0.0) * Ato;
      double r;
      if (flag & 2) { r = mix * dmax(stor, 0.0) / tstep * Afrom; sum += r * tstep; }   // step 1: into the mixing unit
      else r = sum * L.frac[q] / tstep;                                                // step 2: back out by area
      rates[q] = r;
    }
  }
  else if (L.kind == RVN_LAT_REDISTRIBUTE) {   // CmvLatRedistribute::GetLateralExchange, src/SnowRedistribute.cpp:114-215
    const double RAD_TO_DEG = 180.0 / D_PI, SWE_TO_DEPTH_FACTOR = 1000.0 / 250.0;
    for (int q = q0; q < q1; q++) {
      const int kF = L.kfrom[q];
      const double weight = L.frac[q], areaFrom = M.hru_area[kF];
      const double slope_deg = M.hru_slope[kF] * RAD_TO_DEG;
      const double snowSWE = from_row[kF];
      double snowMoved = 0.0;
      if (L.method == RVN_REDIST_CONTINUOUS) {
        const double snowTransport = dmin(0.5, slope_deg / L.tan80) * dmin(1.0, snowSWE / L.mixing_rate);
        snowMoved = snowTransport * snowSWE;
      } else {   // Bernhardt & Schulz (2010) threshold with the FSM2 holding depth
        const double snowDepth = snowSWE * SWE_TO_DEPTH_FACTOR;
        double slope_thres = slope_deg;
        if (slope_thres < 10.0) slope_thres = 10.0;
        const double threshold_snow_height_m = 3178.4 * rvn_pow(slope_thres, -1.998);
        double threshold_snow_height = threshold_snow_height_m * D_MM_PER_METER;
        if (slope_deg > 60.0) threshold_snow_height = 0.0;
        if (snowDepth > threshold_snow_height) { const double excessSnowDepth = snowDepth - threshold_snow_height; snowMoved = excessSnowDepth / SWE_TO_DEPTH_FACTOR; }
      }
      rates[q] = (snowMoved * weight * areaFrom) / tstep;
    }
  }
  else for (int q = q0; q < q1; q++) {   // GetLateralExchange: every rate from the state before any of them is applied
    const int kF = L.kfrom[q], kT = L.kto[q];
    const double stor = from_row[kF], to_stor = to_row[kT];
    const double Afrom = M.hru_area[kF], Ato = M.hru_area[kT];
    const double max_to_stor = lateral_max_storage(M, L, e, kT);
    const double max_rate = dmax(max_to_stor - to_stor, 0.0) / tstep * Ato;
    const double mult = 1.0;
    double r = dmax(mult * stor, 0.0) / tstep * Afrom * L.frac[q];
    r = dmin(r, max_rate);
    rates[q] = r;
  }
  for (int q = q0; q < q1; q++) {   // src/Solvers.cpp:424-430, connection order
    const int kF = L.kfrom[q], kT = L.kto[q];
    const double r = rates[q];
    from_row[kF] -= r / M.hru_area[kF] * tstep;
    to_row[kT] += r / M.hru_area[kT] * tstep;
  }
}

// local-array context for the end-of-step helpers (finish_step / hru_storage are templates over the context)
struct FinCtx {
  const DevProgram *P;
  double sv[RVN_MAX_CORE];
  double area;
  RVN_DI double &SV(int slot) { return sv[slot]; }
  RVN_DI int iSV(int t) const { return P->iSV[t]; }
  RVN_DI int slot_type(int s) const { return P->slot_type[s]; }
  RVN_DI int slot_water(int s) const { return P->slot_water[s]; }
};
__global__ void __launch_bounds__(RVN_BS) finish_kernel(const __grid_constant__ DevModel M, const int slot) {
  __shared__ double s_red[RVN_BS / 32];
  const DevProgram *P = M.prog;
  const int tid = threadIdx.x, e = blockIdx.y, k = blockIdx.x * RVN_BS + tid;
  const size_t nHp = (size_t)M.nHp;
  const bool in_range = (k < M.nH);
  const bool enabled = in_range && (M.hru_enabled[k] != 0);
  const int nCore = P->nCore;
  double swv = 0.0, stor = 0.0;
  if (enabled) {
    FinCtx c; c.P = P; c.area = M.hru_area[k];
    double *gstate = M.state + RVN_STATE_OFF(M, e, 0, k);
    for (int s = 0; s < nCore; s++) c.sv[s] = gstate[(size_t)P->slot_sv[s] * RVN_TILE];
    swv = finish_step(c, nCore);
    stor = hru_storage(c, nCore) * c.area;
    for (int s = 0; s < nCore; s++) gstate[(size_t)P->slot_sv[s] * RVN_TILE] = c.sv[s];
  }
  if (in_range) M.swvol[((size_t)slot * M.E + e) * nHp + k] = swv;
  const double bs = block_sum(stor, s_red);
  if (tid == 0) M.part_storage[((size_t)slot * M.E + e) * M.nBlocksX + blockIdx.x] = bs;
}
#endif
