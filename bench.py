#!/usr/bin/env python
"""bench.py -- HRU-timesteps/s of the fused per-timestep water balance on B200 (BASELINE.json metric).

Headline workload (config.workload): the per-GPU share of BASELINE config 3 -- an HMETS Monte-Carlo ensemble of
512 members x 10,000 HRUs (500 subbasins x 20 HRUs, binary-tree network, ROUTE_NONE/ROUTE_DUMP), synthetic daily
forcings.  Members shard over GPUs with no data-path collective (weak scaling: 512 members per GPU, N=8 is exactly
the 4096-member configuration).  A "step" is one model timestep of every (member, HRU) pair of the rank: forcing
derivation + ordered process sweep + routing + balance reductions.

  value     HRU-steps/s with everything resident in HBM (K steps in one rvn_gpu_run call)
  e2e       the same through the host-facing C ABI one step at a time (rvn_gpu_step_async): every step copies that step's
            gauge forcings from pinned host memory, runs it and reads the per-basin hydrographs + watershed balance back
  roofline  algorithmic bytes of the HRU sweep kernel / its CUDA-event duration vs the measured HBM peak
  cpu_baseline / --impl reference: the unmodified reference executable (oracle/_ref/Raven.exe) on the box's host cores,
            one single-threaded process per core, bounded sample of the same workload
  other_configs  the other named shapes of BASELINE.json, each at its full size with its own value / roofline / e2e /
            cpu_baseline, measured outside the headline's timed region:
              "2"  HBV-EC 10k HRUs / 500 subbasins, Muskingum, 1 member            (replicas only at N > 1)
              "4"  Blended 1M HRUs / 50k subbasins, 200 x 200 forcing grid, hourly   (replicas only at N > 1)
              "5"  EnKF, HBV-EC 512 members x 100k HRUs: forecast + one analysis; members sharded over ranks, the state
                   rows all-gathered over NCCL (the one collective of this design)
              "3_total4096"  config 3 with 4096 members in total (strong scaling: 4096 / N members per GPU)
            (--skip-other to leave them out; --only CFG to run just one)
"""
import argparse
import json
import os
import re
import shutil
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MEMBERS_PER_GPU = 512
N_SB, HRU_PER_SB = 500, 20
METRIC = "HRU-timesteps/sec (members x HRUs x steps)"
MIN_REF_STEPS = 25   # shortest reference sample (Raven's printout covers the whole time loop)


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel, units_per_launch):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel` from the committed `ncu --set full` captures
    (profiles/r0*/ncu_traffic.json, latest round first), or None when there is no capture of this kernel at this launch size."""
    for rnd in ("r02", "r01"):
        try:
            with open(os.path.join(ROOT, "profiles", rnd, "ncu_traffic.json")) as f:
                rec = json.load(f).get(kernel)
            if rec and int(rec["units_per_launch"]) == int(units_per_launch):
                return float(rec["dram_bytes_per_launch"])
        except Exception:
            pass
    return None


class ClockSampler:
    Q = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.p, self.path = index, None, None

    def start(self):
        if not shutil.which("nvidia-smi"):
            return
        fd, self.path = tempfile.mkstemp(suffix=".csv")
        os.close(fd)
        self.f = open(self.path, "w")
        self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                  stdout=self.f, stderr=subprocess.DEVNULL)

    def stop(self, t0=None, t1=None):
        """Samples taken between wall-clock times t0 and t1 (the timed region); all samples if none fall inside."""
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.p:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.close()
        rows = []
        for line in open(self.path):
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(c[1]), float(c[2]), c[5:9]))
            except ValueError:
                continue
        os.unlink(self.path)
        inside = [r for r in rows if t0 is not None and t0 <= r[0] <= t1]
        use = inside if inside else rows
        reasons = set()
        for r in use:
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if use:
            out = {"sm_mhz": statistics.median([r[1] for r in use]), "sm_max_mhz": max(r[2] for r in use), "reasons": sorted(reasons),
                   "samples": len(use), "window": "timed region" if inside else "whole run (no sample fell inside the timed region)"}
        return out


# ------------------------------------------------------------------------------------------------------------------ workloads
def write_config3(td, n_days, members=0):
    from ravenhydroframework_b200 import synth
    return synth.write_hmets(td, n_sb=N_SB, hru_per_sb=HRU_PER_SB, n_days=n_days, seed=1, routing="ROUTE_NONE", catchment_route="ROUTE_DUMP",
                             single_class=True, ensemble_members=members, season_offset=105)  # timed window = snowmelt season: every store is active


def write_config2(td, n_days, n_sb=500):
    from ravenhydroframework_b200 import synth
    return synth.write_hbv(td, n_sb=n_sb, hru_per_sb=20, n_days=n_days, seed=2, routing="ROUTE_MUSKINGUM", season_offset=100, name="cfg2")


def write_config4(td, n_hours, n_sb=50000, n_gauges=1):
    from ravenhydroframework_b200 import synth
    base = synth.write_blended(td, n_sb=n_sb, hru_per_sb=20, n_days=max(1, (n_hours + 23) // 24), seed=4, routing="ROUTE_MUSKINGUM",
                               catchment_route="TRIANGULAR_UH", season_offset=124, name="cfg4", n_gauges=n_gauges)
    rvi = open(base + ".rvi").read()
    open(base + ".rvi", "w").write(rvi.replace(":TimeStep              1.0", ":TimeStep              01:00:00"))
    return base


def write_config5(td, n_days, members, n_sb=5000):
    from ravenhydroframework_b200 import synth
    base = synth.write_hbv(td, n_sb=n_sb, hru_per_sb=20, n_days=n_days, seed=31, routing="ROUTE_MUSKINGUM", name="cfg5", season_offset=130)
    if members:
        synth.add_enkf(base, n_members=members, n_sb=n_sb, n_hru=n_sb * 20, n_days=n_days, obs_basins=tuple(range(1, 11)), window=1)
    return base


def run_reference_cpu(base, nproc, td):
    """Unmodified reference executable on the input set `base`, `nproc` concurrent single-threaded processes.  Returns the aggregate
    HRU-steps/s from Raven's own :BenchmarkingMode printout (time loop only, src/RavenMain.cpp:190-192) and the wall time."""
    exe = os.path.join(ROOT, "oracle", "_ref", "Raven.exe")
    if not os.path.exists(exe):
        return None
    with open(base + ".rvi", "a") as f:
        f.write(":BenchmarkingMode\n:SuppressOutput\n")
    procs = []
    t0 = time.time()
    for i in range(nproc):
        od = os.path.join(td, "refout%d" % i) + "/"
        os.makedirs(od, exist_ok=True)
        procs.append(subprocess.Popen([exe, os.path.basename(base), "-o", od], cwd=os.path.dirname(base), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    rates = []
    for p in procs:
        out = p.communicate()[0]
        m = re.search(r"([0-9.eE+]+)\s+HRU-time steps/second", out)
        if m:
            rates.append(float(m.group(1)))
    wall = time.time() - t0
    if not rates:
        return None
    return dict(value=sum(rates), wall_s=wall, nproc=len(rates), per_proc=statistics.median(rates))


def cpu_baseline_for(cfg, ncores):
    """Bounded reference sample of one named shape (about 10-20 s of CPU work)."""
    with tempfile.TemporaryDirectory() as td:
        if cfg == "3":
            steps = 100
            r = run_reference_cpu(write_config3(td, steps), min(ncores, 16), td)
            what = "%d concurrent single-threaded Raven.exe processes, one 10k-HRU HMETS member each, %d daily steps" % (r["nproc"] if r else 0, steps)
        elif cfg == "2":
            steps = 60
            r = run_reference_cpu(write_config2(td, steps), 1, td)
            what = "the full 10k-HRU / 500-subbasin HBV-EC model, %d daily steps, one process (the reference is single-threaded and the shape has one member)" % steps
        elif cfg == "4":
            steps = 12
            r = run_reference_cpu(write_config4(td, steps, n_sb=2000, n_gauges=9), 1, td)
            what = ("the same Blended process list on 40,000 HRUs / 2,000 subbasins (1/25 of the watershed), %d hourly steps, one process; the reference cannot "
                    "read gridded input here (no NetCDF): 9-station INTERP_FROM_FILE weights instead of the 200 x 200 grid" % steps)
        else:
            steps = 5
            r = run_reference_cpu(write_config5(td, steps, 0), min(ncores, 16), td)
            what = "%d concurrent single-threaded Raven.exe processes, one 100k-HRU HBV-EC member each (no perturbations), %d daily steps" % (r["nproc"] if r else 0, steps)
    if r is None:
        return None
    return {"value": r["value"], "unit": "HRU-steps/s", "cores": r["nproc"], "kind": "reference",
            "sample": what + " (%.1f s wall incl. parsing; per-process %.0f HRU-steps/s)" % (r["wall_s"], r["per_proc"])}


# ------------------------------------------------------------------------------------------------------------------ measurement
class Ctx:
    pass


def _live_stores(model, E):
    """Mean number of live convolution stores per (member, HRU): sum over convolution processes of N(member, land-use class of the HRU)."""
    import ctypes as C
    import numpy as np
    lib = model.lib
    lib.rvh_desc_conv_n.restype = C.c_void_p
    lib.rvh_desc_conv_n.argtypes = [C.c_void_p]
    info = (C.c_int * 8)()
    lib.rvh_desc_info(model.desc, info)
    nconv, nLU = info[4], info[7]
    if nconv == 0:
        return 0.0
    arr = np.ctypeslib.as_array((C.c_int32 * (E * nLU * nconv)).from_address(lib.rvh_desc_conv_n(model.desc))).reshape(E, nLU, nconv)
    lu = np.zeros(model.nHRU, dtype=np.int32)
    lib.rvh_desc_hru_lu.argtypes = [C.c_void_p, C.c_void_p]
    lib.rvh_desc_hru_lu(model.desc, lu.ctypes.data)
    hrus_per_class = np.bincount(lu, minlength=nLU).astype(np.float64)
    return float((arr.sum(axis=2) * hrus_per_class[None, :]).sum() / (E * model.nHRU))


def resident_pass(cx, sim, W, K, clock_index=None):
    """W warm-up steps, then exactly K timed steps in one rvn_gpu_run call; CUDA events on the engine's stream."""
    sim.set_kernel_timing(True)
    clocks = ClockSampler(clock_index) if clock_index is not None else None
    if clocks:
        clocks.start()
    if W:
        sim.run(W)
    cx.barrier()
    tw0 = time.time()
    t0 = time.perf_counter()
    sim.run(K)
    stats = sim.last_run_stats()
    cx.barrier()
    wall_ms = 1000.0 * (time.perf_counter() - t0)
    clk = clocks.stop(tw0, time.time()) if clocks else None
    sim.set_kernel_timing(False)
    return stats["total_ms"], stats["sweep_ms"], stats["launches"], wall_ms, clk


def stream_pass(cx, sim, model, W, K, E, NB):
    """End to end through rvn_gpu_step_async: per step the forcing slab comes from pinned host memory and the per-basin outflows +
    watershed record go back into that step's slot of a pinned host array; nothing waits until the final synchronisation."""
    import torch
    nG = model.gauge_series().shape[1]
    series = torch.from_numpy(model.gauge_series())                          # [9][nG][nSteps]
    s0 = sim.step
    slabs = torch.zeros((W + K, nG, 9), dtype=torch.float64).pin_memory()      # one station-major slab [nGauges][RVN_GF_COUNT] per step
    slabs[:] = series[:, :, s0:s0 + W + K].permute(2, 1, 0)
    hyd_all = torch.zeros((W + K, E, NB), dtype=torch.float64).pin_memory()
    wsh_all = torch.zeros((W + K, E, 4), dtype=torch.float64).pin_memory()
    sb, hb, wb = slabs.data_ptr(), hyd_all.data_ptr(), wsh_all.data_ptr()

    def step(i):
        sim.step_async(sb + i * 9 * nG * 8, hb + i * E * NB * 8, wb + i * E * 4 * 8)
    for i in range(W):
        step(i)
    sim.sync()
    cx.barrier()
    t1 = time.perf_counter()
    for i in range(W, W + K):
        step(i)
    sim.sync()              # every result of the K steps is in host memory here
    cx.barrier()
    ms = 1000.0 * (time.perf_counter() - t1)
    assert bool(torch.isfinite(hyd_all[W + K - 1]).all()) and float(hyd_all[W + K - 1].abs().sum()) > 0.0
    return ms, 9 * nG * 8, E * NB * 8 + E * 4 * 8, hyd_all


def roofline_of(sim, model, E, nH, sweep_ms, K, dev_ms, note=None):
    peak, peak_src = measured_peak()
    live = _live_stores(model, E)
    units = E * nH
    b_live = 16.0 * (model.nCore + live) + 8.0            # state in+out of the live variables + surface-water volume
    b_full = 16.0 * model.NS + 8.0                         # reference layout: every state variable in and out
    per_launch_ms = sweep_ms / K
    achieved = units * b_live / (per_launch_ms / 1000.0) / 1e9
    r = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(sim.kernel_variant(), units),
         "kernel": sim.kernel_variant(), "kernel_ms_per_launch": per_launch_ms, "kernel_share_of_step": sweep_ms / dev_ms, "units_per_launch": units,
         "bytes_per_hru_step": b_live, "bytes_per_hru_step_reference_layout": b_full, "live_conv_stores_mean": live, "peak_source": peak_src}
    if note:
        r["note"] = note
    return r


def max_over_ranks(cx, vals):
    import torch
    import torch.distributed as dist
    t = torch.tensor(vals, dtype=torch.float64, device="cuda")
    if cx.world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


# ---- single-member shapes (configs 2 and 4): replicas only at N > 1
def bench_single_member(cx, cfg, K, W):
    from ravenhydroframework_b200 import api, synth
    td = tempfile.mkdtemp(prefix="rvnb%s_%d_" % (cfg, cx.rank))
    try:
        t_setup = time.time()
        n = 2 * (W + K) + 2
        if cfg == "2":
            base = write_config2(td, n)
            model = api.RavenModel(base)
            model.flatten(1)
            wl = "HBV-EC semi-distributed, 10,000 HRUs / 500 subbasins, ROUTE_MUSKINGUM, daily step, 1 member (BASELINE config 2)"
            l2 = "working set (%.1f MB of state) is smaller than L2 by definition of this configuration: launch / latency bound, the HBM fraction is reported for completeness" % (model.nHRU * model.NS * 8 / 1e6)
        else:
            base = write_config4(td, n)
            model = api.RavenModel(base)
            model.flatten(1)
            ncell = synth.apply_cell_grid(model, base, 200, 200)
            model.flatten(1)
            wl = "Blended (process groups) distributed, 1,000,000 HRUs / 50,000 subbasins, forcing on a 200 x 200 grid (%d cells, 4 per HRU, CSR gather), hourly step, ROUTE_MUSKINGUM, 1 member (BASELINE config 4)" % ncell
            l2 = "inputs larger than L2 (%.2f GB of state swept every step)" % (model.nHRU * model.NS * 8 / 1e9)
        sim = api.GpuSimulation(model, n_members=1, device=cx.local_rank)
        setup_s = time.time() - t_setup
        nH, NB = model.nHRU, model.nSB
        sim.set_recording(n, hydrographs=False, watershed=True)
        dev_ms, sweep_ms, launches, wall_ms, _ = resident_pass(cx, sim, W, K)
        e2e_ms, h2d, d2h, _ = stream_pass(cx, sim, model, W, K, 1, NB)
        dev_ms, sweep_ms, e2e_ms = max_over_ranks(cx, [dev_ms, sweep_ms, e2e_ms])
        units = nH * cx.world
        out = {"workload": wl, "n_gpus": cx.world, "parallelism": "replicas only (one member; every rank runs its own copy)" if cx.world > 1 else "single GPU",
               "steps": K, "warmup": W, "value": units * K / (dev_ms / 1000.0), "unit": "HRU-steps/s", "ms_per_step": dev_ms / K, "l2_policy": l2,
               "roofline": roofline_of(sim, model, 1, nH, sweep_ms, K, dev_ms),
               "e2e": {"value": units * K / (e2e_ms / 1000.0), "unit": "HRU-steps/s", "ms_per_step": e2e_ms / K, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "api": "rvn_gpu_step_async (streaming) + rvn_gpu_sync"},
               "gpu_launches": int(launches), "launches_per_step": launches / float(K), "routing_levels": int(launches / K) - 2, "setup_s": setup_s}
        sim.close()
        return out
    finally:
        shutil.rmtree(td, ignore_errors=True)


# ---- EnKF window (config 5): members sharded over ranks, NCCL all-gather of the state rows at the analysis
def bench_enkf(cx, K, W, members=512, n_sb=5000):
    import numpy as np
    import torch
    from ravenhydroframework_b200 import api, enkf
    td = tempfile.mkdtemp(prefix="rvnb5_%d_" % cx.rank)
    try:
        t_setup = time.time()
        n = W + K
        base = write_config5(td, n, members, n_sb)
        model = api.RavenModel(base)
        sim = enkf.EnKFSimulation(model, device=cx.local_rank, rank=cx.rank, world=cx.world)
        setup_s = time.time() - t_setup
        N, M, nloc, nH = sim.N, sim.info["M"], sim.n_local, model.nHRU
        dev = torch.device("cuda", cx.local_rank)
        g = sim.sim
        # ---- forecast of the local members over the window (W warm-up + K timed steps of the same window), resident
        dev_ms, sweep_ms, launches, wall_ms, _ = resident_pass(cx, g, W, K)
        q = g.hydrographs(0, sim.nsteps)
        sim.out_local = model.enkf_harvest_outputs(q, sim.qout_initial)
        # ---- analysis, piece by piece (CUDA events / wall clock as noted)
        cx.barrier()
        t0 = time.perf_counter()
        out_all = enkf.gather_member_rows(torch.from_numpy(sim.out_local).to(dev), N).cpu().numpy()
        Z = model.enkf_compute_Z(out_all)
        z_ms = 1000.0 * (time.perf_counter() - t0)
        x_local = torch.empty((nloc, M), dtype=torch.float64, device=dev)
        g.enkf_pack_device(x_local.data_ptr())
        cx.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x_all = enkf.gather_member_rows(x_local, N).contiguous()          # warm-up of the collective (NCCL channel setup)
        cx.barrier()
        ev0.record()
        x_all = enkf.gather_member_rows(x_local, N).contiguous()
        ev1.record()
        cx.barrier()
        gather_ms = ev0.elapsed_time(ev1)
        x_pre = x_all.clone()
        t0 = time.perf_counter()
        g.enkf_update_device(x_all.data_ptr(), Z, N, sim.member0)
        torch.cuda.synchronize(dev)
        upd_call_ms = 1000.0 * (time.perf_counter() - t0)
        upd_ms = g.enkf_last_ms()
        # ---- checker (oracle, outside every timed region): a 512-row sample of the analysis, bit for bit
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import enkf_oracle as EO
        rows = np.random.RandomState(1).choice(M, size=min(M, 512), replace=False)
        obs, noise, _, _ = model.enkf_observations()
        rt = torch.from_numpy(rows).to(dev)
        Xa_ref, Zo = EO.assimilation_calcs(x_pre[:, rt].cpu().numpy(), out_all, obs, noise)
        got = x_all[sim.member0:sim.member0 + nloc][:, rt].cpu().numpy()
        same = bool(np.array_equal(got, Xa_ref[sim.member0:sim.member0 + nloc])) and bool(np.array_equal(Z, Zo))
        dev_ms, sweep_ms, gather_ms, upd_ms, upd_call_ms, z_ms = max_over_ranks(cx, [dev_ms, sweep_ms, gather_ms, upd_ms, upd_call_ms, z_ms])
        units = N * nH
        window_ms = dev_ms + z_ms + gather_ms + upd_call_ms
        recv_bytes = (N - nloc) * M * 8
        ntile = (nloc + 15) // 16
        out = {"workload": "EnKF streamflow assimilation, HBV-EC %d members x %d HRUs / %d subbasins, forcing perturbations, SOIL[0] + SNOW of every HRU and "
                           "STREAMFLOW of every subbasin assimilated at 10 gauged outlets, window 1 (BASELINE config 5)" % (N, nH, n_sb),
               "n_gpus": cx.world, "parallelism": "members sharded (%d per GPU), state rows all-gathered at the analysis" % nloc, "steps": K, "warmup": W,
               "value": units * K / (dev_ms / 1000.0), "unit": "HRU-steps/s", "ms_per_step": dev_ms / K,
               "l2_policy": "inputs larger than L2 (%.1f GB of state per GPU swept every step)" % (nloc * nH * model.NS * 8 / 1e9),
               "roofline": roofline_of(g, model, nloc, nH, sweep_ms, K, dev_ms),
               "analysis": {"state_matrix_rows": M, "members": N, "observations": int(obs.shape[1]), "host_Z_ms": z_ms, "allgather_ms": gather_ms,
                            "allgather_bytes_received_per_rank": recv_bytes, "allgather_GBps_per_rank": (recv_bytes / (gather_ms / 1000.0) / 1e9) if cx.world > 1 else None,
                            "backend": "nccl" if cx.world > 1 else "local (one rank holds every member)",
                            "update_kernel_ms": upd_ms, "update_call_ms": upd_call_ms, "update_fp64_gflops": 2.0 * N * M * nloc / (upd_ms * 1e-3) / 1e9,
                            "update_stream_GBps": (N * M * 8.0 * ntile + N * M * 8.0 + nloc * M * 8.0) / (upd_ms * 1e-3) / 1e9,
                            "matches_oracle_bitwise_on_512_rows": same},
               "e2e": {"value": units * K / (window_ms / 1000.0), "unit": "HRU-steps/s", "ms_per_window": window_ms,
                       "definition": "the K forecast steps + one complete analysis (gather of the simulated observations, host Z, all-gather of X, device update)",
                       "h2d_bytes_per_step": int(N * N * 8 / K), "d2h_bytes_per_step": int(nloc * sim.nsteps * model.nSB * 8 / K)},
               "gpu_launches": int(launches) + 3, "setup_s": setup_s}
        assert same, "EnKF analysis differs from the oracle"
        g.close()
        return out
    finally:
        shutil.rmtree(td, ignore_errors=True)


# ---- the HMETS ensemble (config 3): headline at 512 members per GPU, strong-scaling line at 4096 members in total
def bench_config3(cx, K, W, E, full=True):
    import torch
    from ravenhydroframework_b200 import api, ensemble
    n_days = 2 * (W + K) + 4
    td = tempfile.mkdtemp(prefix="rvnb3_%d_" % cx.rank)
    try:
        base = write_config3(td, n_days, members=E * cx.world)   # .rve: the 19 HMETS parameters, uniform +-20 %
        t_setup = time.time()
        model = api.RavenModel(base)
        model.flatten(E)
        sim = api.GpuSimulation(model, n_members=E, device=cx.local_rank)
        # Monte-Carlo members exactly as the reference's member loop would draw them (CMonteCarloEnsemble::UpdateModel,
        # one shared rand() stream): this rank owns global members rank*E .. rank*E+E-1
        member0, n_local = ensemble.shard_members(E * cx.world, cx.rank, cx.world)
        assert n_local == E
        sim.apply_ensemble(member0)
        setup_s = time.time() - t_setup
        nH, NB = model.nHRU, model.nSB
        sim.set_recording(n_days, hydrographs=False, watershed=True)
        dev_ms, sweep_ms, launches, wall_ms, clk = resident_pass(cx, sim, W, K, clock_index=cx.local_rank if full else None)
        res = dict(setup_s=setup_s, launches=launches, clk=clk, wall_ms=wall_ms, nH=nH, NB=NB)
        if full:
            e2e_ms, h2d, d2h, hyd_all = stream_pass(cx, sim, model, W, K, E, NB)
            res.update(h2d=h2d, d2h=d2h)
            # ---- final ensemble statistics over ALL members (outside the timed regions): outlet hydrograph of the end-to-end pass,
            # reduced with NCCL all-reduces (sum/count, squared deviations, min, max)
            q_out = hyd_all[W:W + K, :, 0].T.contiguous().cuda()     # [members, K]
            torch.cuda.synchronize()
            ts = time.perf_counter()
            st = ensemble.ensemble_statistics(q_out)
            torch.cuda.synchronize()
            res["ensemble_statistics"] = {"members": st["n"], "outlet_q_mean_last_step": float(st["mean"][-1]), "outlet_q_std_last_step": float(st["std"][-1]),
                                          "ms": 1000.0 * (time.perf_counter() - ts), "backend": "nccl" if cx.world > 1 else "local"}
        else:
            e2e_ms = dev_ms
        dev_ms, e2e_ms, sweep_ms, wall_ms = max_over_ranks(cx, [dev_ms, e2e_ms, sweep_ms, wall_ms])
        res.update(dev_ms=dev_ms, e2e_ms=e2e_ms, sweep_ms=sweep_ms, roofline=roofline_of(sim, model, E, nH, sweep_ms, K, dev_ms))
        sim.close()
        return res
    finally:
        shutil.rmtree(td, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--members", type=int, default=MEMBERS_PER_GPU, help="members per GPU (default: the 8-GPU share of config 3)")
    ap.add_argument("--total-members", type=int, default=4096, help="ensemble size of the strong-scaling line other_configs['3_total4096']")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-other", action="store_true", help="headline workload only")
    ap.add_argument("--only", default="", help="run just this entry of other_configs (2, 4, 5, 3_total4096) next to the headline")
    args = ap.parse_args()
    for v in os.environ:
        if v.startswith("RVN_") and v not in ("RVN_BENCH_TAG",):
            sys.stderr.write("bench.py: refusing to run with %s set (no environment switch may change what is measured)\n" % v)
            return 2
    K, W = args.steps, max(args.warmup, 0)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ncores = os.cpu_count() or 1
    config = {"workload": "HMETS Monte-Carlo ensemble, %d members/GPU x %d HRUs (%d subbasins x %d), ROUTE_NONE, daily step "
                          "(BASELINE config 3 sharded over GPUs)" % (args.members, N_SB * HRU_PER_SB, N_SB, HRU_PER_SB),
              "members_per_gpu": args.members, "hrus": N_SB * HRU_PER_SB, "subbasins": N_SB, "parallelism": "members sharded, no data-path collective",
              "l2_policy": "inputs larger than L2 (state %.1f GB per GPU is swept every step)" % (args.members * N_SB * HRU_PER_SB * 114 * 8 / 1e9)}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = max(K + W, MIN_REF_STEPS)
        with tempfile.TemporaryDirectory() as td:
            r = run_reference_cpu(write_config3(td, steps), ncores, td)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/Raven.exe missing (reference not built)"}))
            return 0
        # a "step" of this arm = one model step of the bounded sample it actually ran: nproc members x 10k HRUs
        sample_units = r["nproc"] * N_SB * HRU_PER_SB
        ms = 1000.0 * sample_units / r["value"]
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "HRU-steps/s", "n_gpus": args.gpus, "steps": K, "warmup": W,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": r["value"], "unit": "HRU-steps/s", "cores": r["nproc"], "kind": "reference",
                                 "sample": "%d concurrent single-threaded Raven.exe runs (one member each) of the same 10k-HRU HMETS inputs, %d daily steps, Raven's own "
                                           "benchmarking printout summed; ms_per_step is one step of this %d-member sample (%.1f s wall incl. parsing)"
                                           % (r["nproc"], steps, r["nproc"], r["wall_s"])},
                "e2e": {"value": r["value"], "unit": "HRU-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        sys.stderr.write("bench.py: no CUDA device -- the B200 arm has no CPU fallback\n")
        return 2
    # stdout carries exactly ONE line (the JSON): library chatter on fd 1 (NCCL prints its version banner there) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    cx = Ctx()
    cx.rank, cx.world, cx.local_rank = rank, world, local_rank

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    cx.barrier = barrier
    try:
        h = bench_config3(cx, K, W, args.members, full=True)
        units = args.members * h["nH"] * world
        line = {"metric": METRIC, "value": units * K / (h["dev_ms"] / 1000.0), "unit": "HRU-steps/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": h["dev_ms"] / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "roofline": h["roofline"],
                "e2e": {"value": units * K / (h["e2e_ms"] / 1000.0), "unit": "HRU-steps/s", "h2d_bytes_per_step": h["h2d"], "d2h_bytes_per_step": h["d2h"],
                        "ms_per_step": h["e2e_ms"] / K, "api": "rvn_gpu_step_async (streaming) + rvn_gpu_sync"},
                "gpu_launches": int(h["launches"]), "clocks": h["clk"], "wall_ms_timed_region": h["wall_ms"], "setup_s": h["setup_s"],
                "ensemble_statistics": h["ensemble_statistics"]}
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline_for("3", ncores)
            if cb:
                line["cpu_baseline"] = cb
        if not args.skip_other:
            other = {}
            want = [c for c in ("2", "4", "5", "3_total4096") if not args.only or c == args.only]
            for c in want:
                try:
                    if c in ("2", "4"):
                        o = bench_single_member(cx, c, 365 if c == "2" else 48, 5)
                    elif c == "5":
                        o = bench_enkf(cx, 10, 3)
                    else:
                        Et = args.total_members // world
                        s = bench_config3(cx, K, W, Et, full=False)
                        o = {"workload": "HMETS Monte-Carlo ensemble, %d members in total x 10,000 HRUs (BASELINE config 3 at its full size, strong scaling: %d members per GPU)" % (Et * world, Et),
                             "n_gpus": world, "scaling": "strong", "steps": K, "warmup": W, "value": Et * world * s["nH"] * K / (s["dev_ms"] / 1000.0), "unit": "HRU-steps/s",
                             "ms_per_step": s["dev_ms"] / K, "roofline": s["roofline"], "setup_s": s["setup_s"], "gpu_launches": int(s["launches"]),
                             "l2_policy": "inputs larger than L2 (state %.1f GB per GPU)" % (Et * s["nH"] * 114 * 8 / 1e9)}
                    if rank == 0 and world == 1 and not args.no_cpu_baseline and c in ("2", "4", "5"):
                        cb = cpu_baseline_for(c, ncores)
                        if cb:
                            o["cpu_baseline"] = cb
                    other[c] = o
                except Exception as ex:   # a shape that fails is reported, it must not take the headline line with it
                    other[c] = {"error": "%s: %s" % (type(ex).__name__, str(ex)[:300])}
                torch.cuda.empty_cache()
            line["other_configs"] = other
        if rank == 0:
            sys.stdout.flush()
            os.write(json_fd, (json.dumps(line) + "\n").encode())
    finally:
        if world > 1:
            dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
