"""ctypes front end of the B200 build: host layer (libraven_b200.so) + CUDA engine (librvn_gpu.so).

PyTorch is not needed for the engine itself; bench.py uses torch only for distributed plumbing and
pinned host buffers.  Nothing here falls back to a CPU implementation: `GpuSimulation` raises if the
engine reports no CUDA device.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
F_COUNT = 17  # RVN_F_COUNT
FORCING_NAMES = ["PRECIP", "SNOW_FRAC", "TEMP_AVE", "TEMP_DAILY_AVE", "TEMP_DAILY_MIN", "TEMP_DAILY_MAX", "PET", "OW_PET",
                 "POTENTIAL_MELT", "TEMP_MONTH_AVE", "PET_MONTH_AVE", "AIR_PRES", "REL_HUMIDITY", "SW_RADIA", "WIND_VEL", "SW_RADIA_NET", "LW_RADIA_NET"]
MAX_RIVER_SEGS = 50
RES_STATE = 8  # RVN_RES_STATE

_engine = None
_host = None


class RavenError(RuntimeError):
    pass


def engine_lib():
    """librvn_gpu.so (C ABI of include/rvn_gpu.h)."""
    global _engine
    if _engine is None:
        path = os.path.join(HERE, "librvn_gpu.so")
        if not os.path.exists(path):
            raise RavenError("librvn_gpu.so is not built; run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
        lib.rvn_gpu_last_error.restype = C.c_char_p
        vp, i, dp = C.c_void_p, C.c_int, C.c_void_p
        lib.rvn_gpu_create.argtypes = [vp, i, C.POINTER(vp)]
        lib.rvn_gpu_create_ex.argtypes = [vp, i, i, C.POINTER(vp)]
        lib.rvn_gpu_destroy.argtypes = [vp]
        lib.rvn_gpu_destroy.restype = None
        lib.rvn_gpu_upload_state.argtypes = [vp, dp, i, i]
        lib.rvn_gpu_download_state.argtypes = [vp, dp, i, i]
        lib.rvn_gpu_upload_basin_state.argtypes = [vp, dp, dp, dp, dp, i, i]
        lib.rvn_gpu_download_basin_state.argtypes = [vp, dp, dp, dp, dp, i, i]
        lib.rvn_gpu_upload_reservoir_state.argtypes = [vp, dp, i, i]
        lib.rvn_gpu_download_reservoir_state.argtypes = [vp, dp, i, i]
        lib.rvn_gpu_upload_params.argtypes = [vp, dp, dp, dp, i, i]
        lib.rvn_gpu_upload_forcings.argtypes = [vp, dp, dp, i, i]
        lib.rvn_gpu_run.argtypes = [vp, i, i]
        lib.rvn_gpu_sync.argtypes = [vp]
        lib.rvn_gpu_step_async.argtypes = [vp, i, dp, dp, dp]
        lib.rvn_gpu_set_recording.argtypes = [vp, i, i]
        lib.rvn_gpu_enable_forcing_output.argtypes = [vp, i]
        lib.rvn_gpu_set_kernel_timing.argtypes = [vp, i]
        lib.rvn_gpu_download_hydrographs.argtypes = [vp, dp, i, i, i, i]
        lib.rvn_gpu_download_watershed.argtypes = [vp, dp, i, i, i, i]
        lib.rvn_gpu_download_forcings.argtypes = [vp, dp, i, i]
        lib.rvn_gpu_download_cumulative.argtypes = [vp, dp, i, i]
        lib.rvn_gpu_avg_state.argtypes = [vp, dp, i, i]
        lib.rvn_gpu_device_state_ptr.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        lib.rvn_gpu_last_run_stats.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]
        lib.rvn_gpu_kernel_variant.argtypes = [vp, C.c_char_p, i]
        lib.rvn_gpu_emit_static_program.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_int64]
        lib.rvn_gpu_upload_perturbations.argtypes = [vp, dp, i, i]
        lib.rvn_gpu_enkf_configure.argtypes = [vp, dp, C.c_int64]
        lib.rvn_gpu_enkf_pack.argtypes = [vp, dp]
        lib.rvn_gpu_enkf_update.argtypes = [vp, dp, dp, i, i]
        lib.rvn_gpu_enkf_pack_host.argtypes = [vp, dp]
        lib.rvn_gpu_enkf_update_host.argtypes = [vp, dp, dp, i, i]
        lib.rvn_gpu_enkf_last_ms.argtypes = [vp, C.POINTER(C.c_double)]
        lib.rvn_gpu_math_eval.argtypes = [i, i, C.c_int64, dp, dp, dp]
        lib.rvn_gpu_diagnostic.argtypes = [vp, i, i, dp, i, i, i, dp, i, i, dp]
        _engine = lib
    return _engine


def host_lib():
    """libraven_b200.so (Raven-surface C++ host layer)."""
    global _host
    if _host is None:
        engine_lib()
        path = os.path.join(HERE, "libraven_b200.so")
        if not os.path.exists(path):
            raise RavenError("libraven_b200.so is not built; run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = C.CDLL(path)
        vp = C.c_void_p
        lib.rvh_last_error.restype = C.c_char_p
        lib.rvh_model_load.restype = vp
        lib.rvh_model_load.argtypes = [C.c_char_p, C.c_double]
        lib.rvh_model_free.argtypes = [vp]
        lib.rvh_model_free.restype = None
        lib.rvh_model_flatten.restype = vp
        lib.rvh_model_flatten.argtypes = [vp, C.c_int]
        lib.rvh_model_dims.argtypes = [vp] + [C.POINTER(C.c_int)] * 5
        lib.rvh_model_initial_state.argtypes = [vp, vp]
        lib.rvh_model_basin_state.argtypes = [vp, vp, vp, vp, vp]
        lib.rvh_model_sv_name.argtypes = [vp, C.c_int, C.c_char_p, C.c_int]
        lib.rvh_model_process_info.argtypes = [vp, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.rvh_model_update_parameter.argtypes = [vp, C.c_int, C.c_char_p, C.c_char_p, C.c_double]
        lib.rvh_model_flatten_member.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
        lib.rvh_model_num_param_dists.argtypes = [vp]
        lib.rvh_model_param_dist.argtypes = [vp, C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double)]
        lib.rvh_model_options.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.rvh_desc_info.argtypes = [vp, C.POINTER(C.c_int)]
        lib.rvh_ensemble_begin.argtypes = [vp]
        lib.rvh_ensemble_update_model.argtypes = [vp, C.c_int, vp, C.c_int]
        lib.rvh_ensemble_update_model_ex.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int]
        lib.rvh_enkf_begin.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_int)]
        lib.rvh_enkf_rows.argtypes = [vp, vp]
        lib.rvh_enkf_row_name.argtypes = [vp, C.c_int, C.c_char_p, C.c_int]
        lib.rvh_enkf_observations.argtypes = [vp, vp, vp, vp, vp]
        lib.rvh_enkf_member_eps.argtypes = [vp, C.c_int]
        lib.rvh_enkf_member_eps.restype = vp
        lib.rvh_enkf_harvest_outputs.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp]
        lib.rvh_enkf_compute_Z.argtypes = [vp, vp, vp]
        lib.rvh_desc_nseg.argtypes = [vp, vp]
        lib.rvh_desc_num_reservoirs.argtypes = [vp]
        lib.rvh_desc_reservoir_state.argtypes = [vp, vp, vp]
        lib.rvh_model_write_solution.argtypes = [vp, C.c_char_p, C.c_double, vp, vp, vp, vp, vp, C.c_int]
        lib.rvh_model_write_hydrographs.argtypes = [vp, C.c_char_p, vp, vp, vp, vp, C.c_int]
        lib.rvh_model_write_solution_res.argtypes = [vp, C.c_char_p, C.c_double, vp, vp, vp, vp, vp, C.c_int, vp]
        _host = lib
    return _host


class RavenModel:
    """A Raven model parsed from <base>.rvi/.rvh/.rvp/.rvt/.rvc by the C++ host layer (CModel)."""

    def __init__(self, base, duration=0.0):
        self.lib = host_lib()
        self.h = self.lib.rvh_model_load(str(base).encode(), float(duration))
        if not self.h:
            raise RavenError(self.lib.rvh_last_error().decode())
        d = [C.c_int() for _ in range(5)]
        self.lib.rvh_model_dims(self.h, *[C.byref(x) for x in d])
        self.nHRU, self.NS, self.nSB, self.nProc, self.nGauges = [x.value for x in d]
        ts, du, seed, ens, nm = C.c_double(), C.c_double(), C.c_uint(), C.c_int(), C.c_int()
        self.lib.rvh_model_options(self.h, C.byref(ts), C.byref(du), C.byref(seed), C.byref(ens), C.byref(nm))
        self.timestep, self.duration, self.random_seed, self.ensemble_mode, self.num_members = ts.value, du.value, seed.value, ens.value, nm.value
        self.desc = None
        self.nMembers = 0

    def close(self):
        if self.h:
            self.lib.rvh_model_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def flatten(self, n_members=1):
        """Compile to rvn_model_desc (owned by the C++ model). Returns the descriptor pointer."""
        p = self.lib.rvh_model_flatten(self.h, int(n_members))
        if not p:
            raise RavenError(self.lib.rvh_last_error().decode())
        self.desc = p
        self.nMembers = int(n_members)
        info = (C.c_int * 8)()
        self.lib.rvh_desc_info(p, info)
        self.nSteps, self.max_nqin, self.max_nqlat, self.nCore, self.nConvProc, self.ptab_stride = [info[i] for i in range(6)]
        return p

    def gauge_series(self):
        """Copy of the descriptor's forcing series [RVN_GF_COUNT][nGauges][nSteps] (sampled on the model step)."""
        self.lib.rvh_desc_gauge_series.restype = C.c_void_p
        self.lib.rvh_desc_gauge_series.argtypes = [C.c_void_p]
        self.lib.rvh_desc_num_gauges.argtypes = [C.c_void_p]
        nG = self.lib.rvh_desc_num_gauges(self.desc)
        a = np.ctypeslib.as_array((C.c_double * (9 * nG * self.nSteps)).from_address(self.lib.rvh_desc_gauge_series(self.desc)))
        return a.reshape(9, nG, self.nSteps).copy()

    def set_cell_forcing(self, cell_elev, series, wt_ptr, wt_idx, wt_val):
        """Gridded forcing input: one station per grid cell.  series [RVN_GF_COUNT][nCells][nSteps] sampled on the model step, CSR weights
        over the HRUs (the sparse gather of reference CForcingGrid::GetWeightedValue).  Call flatten() afterwards."""
        series = np.ascontiguousarray(series, dtype=np.float64)
        cell_elev = np.ascontiguousarray(cell_elev, dtype=np.float64)
        wt_ptr = np.ascontiguousarray(wt_ptr, dtype=np.int32)
        wt_idx = np.ascontiguousarray(wt_idx, dtype=np.int32)
        wt_val = np.ascontiguousarray(wt_val, dtype=np.float64)
        assert series.ndim == 3 and series.shape[0] == 9 and series.shape[1] == cell_elev.size and wt_ptr.size == self.nHRU + 1
        self.lib.rvh_model_set_cell_forcing.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 5
        if self.lib.rvh_model_set_cell_forcing(self.h, int(series.shape[1]), int(series.shape[2]), cell_elev.ctypes.data, series.ctypes.data,
                                               wt_ptr.ctypes.data, wt_idx.ctypes.data, wt_val.ctypes.data):
            raise RavenError(self.lib.rvh_last_error().decode())
        self.desc = None

    def set_flux_balance(self, on=True):
        """Keep the per-connection cumulative fluxes (reference CModel::_aCumulativeBal) on the device; call flatten() afterwards."""
        self.lib.rvh_model_set_flux_balance.argtypes = [C.c_void_p, C.c_int]
        if self.lib.rvh_model_set_flux_balance(self.h, 1 if on else 0):
            raise RavenError(self.lib.rvh_last_error().decode())
        self.desc = None

    def connections(self):
        """(from, to) state-variable indices of every tabulated process connection, in process-list order."""
        self.lib.rvh_desc_num_connections.argtypes = [C.c_void_p]
        n = self.lib.rvh_desc_num_connections(self.desc)
        f = np.zeros(n, dtype=np.int32); t = np.zeros(n, dtype=np.int32)
        self.lib.rvh_desc_connections.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self.lib.rvh_desc_connections(self.desc, f.ctypes.data, t.ctypes.data)
        return f, t

    def sv_names(self):
        buf = C.create_string_buffer(64)
        out = []
        for i in range(self.NS):
            self.lib.rvh_model_sv_name(self.h, i, buf, 64)
            out.append(buf.value.decode())
        return out

    def processes(self):
        out = []
        buf = C.create_string_buffer(64)
        n = C.c_int()
        f8, t8 = (C.c_int * 8)(), (C.c_int * 8)()
        for j in range(self.nProc):
            self.lib.rvh_model_process_info(self.h, j, buf, 64, C.byref(n), f8, t8)
            m = min(n.value, 8)
            out.append((buf.value.decode(), n.value, list(f8)[:m], list(t8)[:m]))
        return out

    def initial_state(self):
        a = np.zeros((self.nHRU, self.NS))
        self.lib.rvh_model_initial_state(self.h, a.ctypes.data)
        return a

    def basin_state(self):
        assert self.desc, "flatten() first"
        qin = np.zeros((self.nSB, self.max_nqin))
        qlat = np.zeros((self.nSB, self.max_nqlat))
        qout = np.zeros((self.nSB, MAX_RIVER_SEGS))
        scal = np.zeros((self.nSB, 8))
        if self.lib.rvh_model_basin_state(self.h, qin.ctypes.data, qlat.ctypes.data, qout.ctypes.data, scal.ctypes.data):
            raise RavenError(self.lib.rvh_last_error().decode())
        return qin, qlat, qout, scal

    def write_hydrographs(self, filename, qout_rec, precip, qout0=None, qout0_last=None):
        """Hydrographs.csv in the reference's format (period-ending stamps, step-averaged flows, observed series beside the modelled ones,
        6 significant digits) from recorded hydrographs qout_rec [nsteps][nSB] and watershed-average precipitation precip [nsteps]."""
        q = np.ascontiguousarray(qout_rec, dtype=np.float64); pr = np.ascontiguousarray(precip, dtype=np.float64)
        if qout0 is None:
            _, _, qo, sc = self.basin_state()
            nseg = np.zeros(self.nSB, dtype=np.int32); self.lib.rvh_desc_nseg(self.desc, nseg.ctypes.data)
            qout0 = qo[np.arange(self.nSB), nseg - 1].copy(); qout0_last = sc[:, 0].copy()
            res, sb = self.reservoir_state()
            if len(sb):
                qout0[sb] = res[:, 2]; qout0_last[sb] = res[:, 3]
        q0 = np.ascontiguousarray(qout0, dtype=np.float64); q0l = np.ascontiguousarray(qout0_last, dtype=np.float64)
        if self.lib.rvh_model_write_hydrographs(self.h, str(filename).encode(), q.ctypes.data, q0.ctypes.data, q0l.ctypes.data, pr.ctypes.data, int(q.shape[0])):
            raise RavenError(self.lib.rvh_last_error().decode())

    def reservoir_state(self):
        """Initial reservoir state [nRes][8] = {stage, stage_last, Qout, Qout_last, precip m3, losses m3, AET m3, 0} and the basin index of
        each reservoir (rvn_model_desc::res_state0 / res_basin)."""
        assert self.desc, "flatten() first"
        n = self.lib.rvh_desc_num_reservoirs(self.desc)
        st = np.zeros((n, RES_STATE)); sb = np.zeros(n, dtype=np.int32)
        if n:
            self.lib.rvh_desc_reservoir_state(self.desc, st.ctypes.data, sb.ctypes.data)
        return st, sb

    def write_solution(self, filename, model_time, state, qin, qlat, qout, scal, full_precision=True, res=None):
        """Checkpoint of one member in the reference's solution.rvc format (CModel::WriteMajorOutput): state [nHRU][NS], basin arrays
        as returned by GpuSimulation.download_basin_state() for that member, res = its reservoir state [nRes][8] (models with reservoirs).
        full_precision writes %.17g so a restart is bit-exact."""
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in (state, qin, qlat, qout, scal)]
        r = np.ascontiguousarray(res, dtype=np.float64) if res is not None and np.size(res) else None
        if self.lib.rvh_model_write_solution_res(self.h, str(filename).encode(), float(model_time), *[x.ctypes.data for x in a], 1 if full_precision else 0,
                                                 r.ctypes.data if r is not None else None):
            raise RavenError(self.lib.rvh_last_error().decode())

    def update_parameter(self, class_type, pname, cname, value):
        """CModel::UpdateParameter: class_type 0=soil 1=vegetation 2=landuse 3=terrain 4=global."""
        if self.lib.rvh_model_update_parameter(self.h, int(class_type), pname.encode(), cname.encode(), float(value)):
            raise RavenError(self.lib.rvh_last_error().decode())

    def flatten_member(self, member):
        """Refresh member's parameter row / convolution UHs from the current class values; returns raw pointers."""
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        if self.lib.rvh_model_flatten_member(self.h, int(member), C.byref(a), C.byref(b), C.byref(c)):
            raise RavenError(self.lib.rvh_last_error().decode())
        return a, b, c

    def ensemble_begin(self):
        """Create and seed the ensemble object the .rvi/.rve describe (CEnsemble::SetRandomSeed + Initialize). Returns #members."""
        n = self.lib.rvh_ensemble_begin(self.h)
        if n < 0:
            raise RavenError(self.lib.rvh_last_error().decode())
        return n

    def ensemble_update_model(self, e, reparse_ic=True):
        """CEnsemble::UpdateModel(pModel, Options, e): draws member e's parameters from the shared rand() stream, applies them with
        CModel::UpdateParameter and (Monte Carlo / DDS, reparse_ic) re-reads the initial conditions under them, as the reference
        does. Members must be visited in order. Returns the sampled values."""
        vals = np.zeros(256)
        n = self.lib.rvh_ensemble_update_model_ex(self.h, int(e), vals.ctypes.data, 256, 1 if reparse_ic else 0)
        if n < 0:
            raise RavenError(self.lib.rvh_last_error().decode())
        return vals[:n].copy()

    # ---- EnKF (reference CEnKFEnsemble) -------------------------------------------------------------------------------
    def enkf_begin(self, member0=0, nlocal=None):
        """CEnsemble::SetRandomSeed + CEnKFEnsemble::Initialize + UpdateModel(e) for e < member0+nlocal: observation noise and the
        forcing-perturbation samples of the local member block are drawn from the shared rand() stream in the reference's order
        and written into the flattened descriptor (flatten(nlocal) first).  Returns dict(N, M, Nobs, nPert, nsteps, mode)."""
        nlocal = self.nMembers if nlocal is None else nlocal
        assert self.desc and self.nMembers == nlocal, "flatten(nlocal) first"
        info = (C.c_int * 6)()
        if self.lib.rvh_enkf_begin(self.h, int(member0), int(nlocal), info):
            raise RavenError(self.lib.rvh_last_error().decode())
        self.enkf = dict(N=info[0], M=info[1], Nobs=info[2], nPert=info[3], nsteps=info[4], mode=info[5], member0=member0, nlocal=nlocal)
        return self.enkf

    def enkf_rows(self):
        """State-matrix rows [M][4] int32 = (kind, index, sub, 0) in the reference's row order (rvn_enkf_row)."""
        rows = np.zeros((self.enkf["M"], 4), dtype=np.int32)
        if self.lib.rvh_enkf_rows(self.h, rows.ctypes.data):
            raise RavenError(self.lib.rvh_last_error().decode())
        return rows

    def enkf_row_names(self):
        buf = C.create_string_buffer(96)
        out = []
        for i in range(self.enkf["M"]):
            self.lib.rvh_enkf_row_name(self.h, i, buf, 96)
            out.append(buf.value.decode())
        return out

    def enkf_observations(self):
        """(obs [N][Nobs] perturbed observations, noise [N][Nobs], basin [Nobs], nn [Nobs])."""
        N, Nobs = self.enkf["N"], self.enkf["Nobs"]
        obs, noise = np.zeros((N, Nobs)), np.zeros((N, Nobs))
        basin, nn = np.zeros(Nobs, dtype=np.int32), np.zeros(Nobs, dtype=np.int32)
        if self.lib.rvh_enkf_observations(self.h, obs.ctypes.data, noise.ctypes.data, basin.ctypes.data, nn.ctypes.data):
            raise RavenError(self.lib.rvh_last_error().decode())
        return obs, noise, basin, nn

    def enkf_member_eps(self, member_local):
        """Forcing-perturbation samples of a local member [nsteps][nPert] (a copy)."""
        p = self.lib.rvh_enkf_member_eps(self.h, int(member_local))
        n = self.nSteps * self.enkf["nPert"]
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(n,)).reshape(self.nSteps, self.enkf["nPert"]).copy()

    def enkf_harvest_outputs(self, qout_rec, qout_initial):
        """CloseTimeStepOps: simulated observations [nmembers][Nobs] from recorded outflows qout_rec [nsteps][nmembers][nSB]."""
        q = np.ascontiguousarray(qout_rec, dtype=np.float64)
        q0 = np.ascontiguousarray(qout_initial, dtype=np.float64)
        nsteps, nmem, _ = q.shape
        out = np.zeros((nmem, self.enkf["Nobs"]))
        if self.lib.rvh_enkf_harvest_outputs(self.h, q.ctypes.data, q0.ctypes.data, nsteps, nmem, out.ctypes.data):
            raise RavenError(self.lib.rvh_last_error().decode())
        return out

    def enkf_compute_Z(self, out_all):
        """Host part of AssimilationCalcs: Z [N][N] from the simulated observations of all members; None when no analysis is due."""
        o = np.ascontiguousarray(out_all, dtype=np.float64)
        N = self.enkf["N"]
        assert o.shape == (N, self.enkf["Nobs"])
        Z = np.zeros((N, N))
        rc = self.lib.rvh_enkf_compute_Z(self.h, o.ctypes.data, Z.ctypes.data)
        if rc < 0:
            raise RavenError(self.lib.rvh_last_error().decode())
        return Z if rc == 1 else None

    def param_dists(self):
        out = []
        pn, cn = C.create_string_buffer(64), C.create_string_buffer(64)
        ct, di = C.c_int(), C.c_int()
        par = (C.c_double * 3)()
        for i in range(self.lib.rvh_model_num_param_dists(self.h)):
            self.lib.rvh_model_param_dist(self.h, i, pn, cn, 64, C.byref(ct), C.byref(di), par)
            out.append(dict(name=pn.value.decode(), cls=cn.value.decode(), class_type=ct.value, dist=di.value, par=list(par)))
        return out


class GpuSimulation:
    """Device-resident simulation of all members of a flattened model (rvn_gpu_* C ABI)."""

    def __init__(self, model, n_members=1, device=0, upload_initial=True, force_interpreter=False):
        self.lib = engine_lib()
        self.model = model
        if model.desc is None or model.nMembers != n_members:
            model.flatten(n_members)
        self.E, self.nHRU, self.NS, self.nSB = n_members, model.nHRU, model.NS, model.nSB
        self.nSteps = model.nSteps
        h = C.c_void_p()
        rc = self.lib.rvn_gpu_create_ex(model.desc, int(device), 1 if force_interpreter else 0, C.byref(h))
        if rc != 0:
            raise RavenError("rvn_gpu_create: %s" % self.lib.rvn_gpu_last_error().decode())
        self.h = h
        self.step = 0
        self._st0 = None
        if upload_initial:
            st = model.initial_state()
            self._st0 = st
            qin, qlat, qout, scal = model.basin_state()
            for e in range(n_members):
                self.upload_state(st, e, 1)
                self._ck(self.lib.rvn_gpu_upload_basin_state(self.h, qin.ctypes.data, qlat.ctypes.data, qout.ctypes.data, scal.ctypes.data, e, 1))

    def _ck(self, rc):
        if rc != 0:
            raise RavenError(self.lib.rvn_gpu_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.rvn_gpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload_state(self, state, member0=0, nmembers=None):
        state = np.ascontiguousarray(state, dtype=np.float64)
        if nmembers is None:
            nmembers = state.shape[0] if state.ndim == 3 else 1
        assert state.size == nmembers * self.nHRU * self.NS
        self._ck(self.lib.rvn_gpu_upload_state(self.h, state.ctypes.data, member0, nmembers))

    def download_state(self, member0=0, nmembers=None):
        nmembers = self.E - member0 if nmembers is None else nmembers
        out = np.empty((nmembers, self.nHRU, self.NS))
        self._ck(self.lib.rvn_gpu_download_state(self.h, out.ctypes.data, member0, nmembers))
        return out

    def download_basin_state(self, member0=0, nmembers=None):
        nmembers = self.E - member0 if nmembers is None else nmembers
        m = self.model
        qin = np.empty((nmembers, self.nSB, m.max_nqin)); qlat = np.empty((nmembers, self.nSB, m.max_nqlat))
        qout = np.empty((nmembers, self.nSB, MAX_RIVER_SEGS)); scal = np.empty((nmembers, self.nSB, 8))
        self._ck(self.lib.rvn_gpu_download_basin_state(self.h, qin.ctypes.data, qlat.ctypes.data, qout.ctypes.data, scal.ctypes.data, member0, nmembers))
        return qin, qlat, qout, scal

    def download_reservoir_state(self, member0=0, nmembers=None):
        nmembers = self.E - member0 if nmembers is None else nmembers
        n = self.model.reservoir_state()[0].shape[0]
        res = np.zeros((nmembers, n, RES_STATE))
        if n:
            self._ck(self.lib.rvn_gpu_download_reservoir_state(self.h, res.ctypes.data, member0, nmembers))
        return res

    def download_flux_balance(self, member0=0, nmembers=None):
        """[member][connection][HRU] cumulative amount moved through every process connection (needs RavenModel.set_flux_balance)."""
        nmembers = self.E - member0 if nmembers is None else nmembers
        f, _ = self.model.connections()
        cum = np.zeros((nmembers, f.size, self.model.nHRU))
        self.lib.rvn_gpu_download_flux_balance.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        self._ck(self.lib.rvn_gpu_download_flux_balance(self.h, cum.ctypes.data, member0, nmembers))
        return cum

    def upload_reservoir_state(self, res, member0=0, nmembers=None):
        res = np.ascontiguousarray(res, dtype=np.float64)
        nmembers = (res.shape[0] if res.ndim == 3 else 1) if nmembers is None else nmembers
        if res.size:
            self._ck(self.lib.rvn_gpu_upload_reservoir_state(self.h, res.ctypes.data, member0, nmembers))

    def write_solution(self, filename, member=0, full_precision=True):
        """solution.rvc of `member` at the current step, straight from the device buffers."""
        st = self.download_state(member, 1)[0]
        qin, qlat, qout, scal = [a[0] for a in self.download_basin_state(member, 1)]
        self.model.write_solution(filename, self.step * self.model.timestep, st, qin, qlat, qout, scal, full_precision, res=self.download_reservoir_state(member, 1)[0])

    def upload_member_params(self, member):
        a, b, c = self.model.flatten_member(member)
        self._ck(self.lib.rvn_gpu_upload_params(self.h, a, b, c, member, 1))

    def apply_ensemble(self, member0=0):
        """Give local member i the parameters of global ensemble member member0+i (reference CEnsemble::UpdateModel order):
        the draws of members 0..member0-1 are replayed and discarded, so any rank of a sharded ensemble sees the same stream."""
        m = self.model
        m.ensemble_begin()
        samples = []
        for e in range(member0 + self.E):
            v = m.ensemble_update_model(e, reparse_ic=(e >= member0))
            if e >= member0:
                self.upload_member_params(e - member0)
                # the member's initial state is the .rvc re-read under its parameters (clipped to sampled capacities); uploaded only
                # where it differs from what the handle already holds
                st = m.initial_state()
                if self._st0 is None or not np.array_equal(st, self._st0):
                    self.upload_state(st, e - member0, 1)
                samples.append(v)
        return np.array(samples)

    def upload_forcings(self, gauge_series, step0, nsteps):
        """gauge_series [RVN_GF_COUNT][nGauges][nsteps] host array replacing steps step0..step0+nsteps."""
        g = np.ascontiguousarray(gauge_series, dtype=np.float64)
        self._ck(self.lib.rvn_gpu_upload_forcings(self.h, g.ctypes.data, None, step0, nsteps))

    def set_recording(self, max_steps, hydrographs=True, watershed=True):
        self._ck(self.lib.rvn_gpu_set_recording(self.h, int(max_steps), (1 if hydrographs else 0) | (2 if watershed else 0)))
        self._rec = 0

    def enable_forcing_output(self, on=True):
        self._ck(self.lib.rvn_gpu_enable_forcing_output(self.h, 1 if on else 0))

    def set_kernel_timing(self, on=True):
        self._ck(self.lib.rvn_gpu_set_kernel_timing(self.h, 1 if on else 0))

    def run(self, nsteps=-1, sync=True):
        if nsteps < 0:
            nsteps = self.nSteps - self.step
        self._ck(self.lib.rvn_gpu_run(self.h, self.step, nsteps))
        self.step += nsteps
        if sync:
            self._ck(self.lib.rvn_gpu_sync(self.h))

    def sync(self):
        self._ck(self.lib.rvn_gpu_sync(self.h))

    def step_async(self, forcing_ptr=None, qout_ptr=None, wshed_ptr=None):
        """Streaming step (rvn_gpu_step_async): raw addresses of PINNED host buffers (or None); nothing waits, sync() completes."""
        self._ck(self.lib.rvn_gpu_step_async(self.h, self.step, forcing_ptr, qout_ptr, wshed_ptr))
        self.step += 1

    def hydrographs(self, rec0, nrec, member0=0, nmembers=None):
        nmembers = self.E - member0 if nmembers is None else nmembers
        out = np.empty((nrec, nmembers, self.nSB))
        self._ck(self.lib.rvn_gpu_download_hydrographs(self.h, out.ctypes.data, rec0, nrec, member0, nmembers))
        return out

    def watershed(self, rec0, nrec, member0=0, nmembers=None):
        nmembers = self.E - member0 if nmembers is None else nmembers
        out = np.empty((nrec, nmembers, 4))
        self._ck(self.lib.rvn_gpu_download_watershed(self.h, out.ctypes.data, rec0, nrec, member0, nmembers))
        return out

    def forcings(self, member0=0, nmembers=None):
        nmembers = self.E - member0 if nmembers is None else nmembers
        out = np.empty((nmembers, self.nHRU, F_COUNT))
        self._ck(self.lib.rvn_gpu_download_forcings(self.h, out.ctypes.data, member0, nmembers))
        return out

    def cumulative(self, member0=0, nmembers=None):
        nmembers = self.E - member0 if nmembers is None else nmembers
        out = np.empty((nmembers, 2))
        self._ck(self.lib.rvn_gpu_download_cumulative(self.h, out.ctypes.data, member0, nmembers))
        return out

    def avg_state(self, member0=0, nmembers=None):
        nmembers = self.E - member0 if nmembers is None else nmembers
        out = np.empty((nmembers, self.NS))
        self._ck(self.lib.rvn_gpu_avg_state(self.h, out.ctypes.data, member0, nmembers))
        return out

    # ---- EnKF analysis on the device ---------------------------------------------------------------------------------
    def upload_perturbations(self, eps, member0=0, nmembers=None):
        e = np.ascontiguousarray(eps, dtype=np.float64)
        nmembers = self.E - member0 if nmembers is None else nmembers
        self._ck(self.lib.rvn_gpu_upload_perturbations(self.h, e.ctypes.data, member0, nmembers))

    def enkf_configure(self, rows):
        r = np.ascontiguousarray(rows, dtype=np.int32)
        self.enkf_M = r.shape[0]
        self._ck(self.lib.rvn_gpu_enkf_configure(self.h, r.ctypes.data, r.shape[0]))

    def enkf_pack(self):
        """AddToStateMatrix for the local members: X [E][M] on the host."""
        X = np.zeros((self.E, self.enkf_M))
        self._ck(self.lib.rvn_gpu_enkf_pack_host(self.h, X.ctypes.data))
        return X

    def enkf_update(self, X_all, Z, member0_global=0):
        """Analysis update + UpdateFromStateMatrix for the local members; X_all [N][M] host array of ALL members (updated in place
        for the local rows).  Returns the updated local rows."""
        X = np.ascontiguousarray(X_all, dtype=np.float64)
        Zc = np.ascontiguousarray(Z, dtype=np.float64)
        N = X.shape[0]
        self._ck(self.lib.rvn_gpu_enkf_update_host(self.h, X.ctypes.data, Zc.ctypes.data, N, int(member0_global)))
        if X is not X_all:
            X_all[...] = X
        return X[member0_global:member0_global + self.E]

    def enkf_pack_device(self, dptr):
        self._ck(self.lib.rvn_gpu_enkf_pack(self.h, C.c_void_p(int(dptr))))

    def enkf_update_device(self, dptr_all, Z, N, member0_global):
        Zc = np.ascontiguousarray(Z, dtype=np.float64)
        self._ck(self.lib.rvn_gpu_enkf_update(self.h, C.c_void_p(int(dptr_all)), Zc.ctypes.data, int(N), int(member0_global)))

    def enkf_last_ms(self):
        v = C.c_double()
        self._ck(self.lib.rvn_gpu_enkf_last_ms(self.h, C.byref(v)))
        return v.value

    def diagnostic(self, kind, basin, obs, nnstart, nnend, qout_initial, member0=0, nmembers=None):
        """NASH_SUTCLIFFE (0) / RMSE (1) / KLING_GUPTA (2) of the recorded hydrograph of `basin` against obs, per member, on the device."""
        nmembers = self.E - member0 if nmembers is None else nmembers
        obs = np.ascontiguousarray(obs, dtype=np.float64)
        q0 = np.ascontiguousarray(np.broadcast_to(np.asarray(qout_initial, dtype=np.float64), (nmembers,)))
        out = np.zeros(nmembers)
        self._ck(self.lib.rvn_gpu_diagnostic(self.h, int(kind), int(basin), obs.ctypes.data, obs.size, int(nnstart), int(nnend), q0.ctypes.data, member0, nmembers, out.ctypes.data))
        return out

    def kernel_variant(self):
        """'static:<PROGRAM>' or 'interpreter' (which HRU-sweep kernel serves this model)."""
        buf = C.create_string_buffer(64)
        self._ck(self.lib.rvn_gpu_kernel_variant(self.h, buf, 64))
        return buf.value.decode()

    def last_run_stats(self):
        a, b, n = C.c_double(), C.c_double(), C.c_int64()
        self._ck(self.lib.rvn_gpu_last_run_stats(self.h, C.byref(a), C.byref(b), C.byref(n)))
        return dict(sweep_ms=a.value, total_ms=b.value, launches=n.value)
