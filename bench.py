#!/usr/bin/env python
"""bench.py -- HRU-timesteps/s of the fused per-timestep water balance on B200 (BASELINE.json metric).

Workload (config.workload): the per-GPU share of BASELINE config 3 -- an HMETS Monte-Carlo ensemble of
512 members x 10,000 HRUs (500 subbasins x 20 HRUs, binary-tree network, ROUTE_NONE/ROUTE_DUMP), synthetic
daily forcings.  Members shard over GPUs with no data-path collective (weak scaling: 512 members per GPU,
N=8 is exactly the 4096-member configuration).  A "step" is one model timestep of every (member, HRU)
pair of the rank: forcing derivation + ordered process sweep + routing + balance reductions.

  value     HRU-steps/s with everything resident in HBM (K steps in one rvn_gpu_run call)
  e2e       the same through the host-facing C ABI one step at a time: every step copies that step's
            gauge forcings from pinned host memory (rvn_gpu_upload_forcings), runs it and reads the
            per-basin hydrographs + watershed balance back to the host
  roofline  algorithmic bytes of the HRU sweep kernel / its CUDA-event duration vs the measured HBM peak
  cpu_baseline / --impl reference: the unmodified reference executable (oracle/_ref/Raven.exe) on the
            box's host cores, one single-threaded process per core, bounded sample of the same workload
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
N_SB, HRU_PER_SB = int(os.environ.get("RVN_BENCH_NSB", 500)), int(os.environ.get("RVN_BENCH_HPS", 20))   # env overrides: layout experiments only
METRIC = "HRU-timesteps/sec (members x HRUs x steps)"
CPU_SAMPLE_STEPS = 100   # bounded cpu_baseline sample: ~16 s of time loop per reference process at ~6e4 HRU-steps/s


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel, units_per_launch):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel` from the committed `ncu --set full` capture
    (profiles/r01/ncu_traffic.json), or None when there is no capture of this kernel at this launch size."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01", "ncu_traffic.json")) as f:
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


def write_workload(td, n_days, members=0):
    from ravenhydroframework_b200 import synth
    return synth.write_hmets(td, n_sb=N_SB, hru_per_sb=HRU_PER_SB, n_days=n_days, seed=1, routing="ROUTE_NONE", catchment_route="ROUTE_DUMP",
                             single_class=True, ensemble_members=members, season_offset=105)  # timed window = snowmelt season: every store is active


def run_reference_cpu(n_days, nproc, td):
    """Unmodified reference, one single-threaded process per core, same 10k-HRU HMETS inputs.  Returns aggregate HRU-steps/s
    from Raven's own :BenchmarkingMode printout (time loop only, src/RavenMain.cpp:190-192)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "Raven.exe")
    if not os.path.exists(exe):
        return None
    base = write_workload(td, n_days)
    with open(base + ".rvi", "a") as f:
        f.write(":BenchmarkingMode\n:SuppressOutput\n")
    procs = []
    t0 = time.time()
    for i in range(nproc):
        od = os.path.join(td, "out%d" % i) + "/"
        os.makedirs(od, exist_ok=True)
        procs.append(subprocess.Popen([exe, os.path.basename(base), "-o", od], cwd=td, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--members", type=int, default=MEMBERS_PER_GPU, help="members per GPU (default: the 8-GPU share of config 3)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
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
        with tempfile.TemporaryDirectory() as td:
            r = run_reference_cpu(max(K + W, 8), ncores, td)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/Raven.exe missing (reference not built)"}))
            return 0
        ms = 1000.0 * N_SB * HRU_PER_SB * args.members * world / r["value"]
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "HRU-steps/s", "n_gpus": args.gpus, "steps": K, "warmup": W,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": r["value"], "unit": "HRU-steps/s", "cores": r["nproc"], "kind": "reference",
                                 "sample": "%d concurrent single-threaded Raven.exe runs (one member each) of the same 10k-HRU HMETS inputs, %d daily steps, "
                                           "Raven's own benchmarking printout summed" % (r["nproc"], max(K + W, 8))},
                "e2e": {"value": r["value"], "unit": "HRU-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ B200 arm
    import numpy as np
    import torch
    import torch.distributed as dist
    from ravenhydroframework_b200 import api
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
    E = args.members
    n_days = 3 * (W + K) + 4  # resident pass + two end-to-end passes (streaming, synchronous)
    td = tempfile.mkdtemp(prefix="rvnbench%d_" % rank)
    try:
        base = write_workload(td, n_days, members=E * world)   # .rve: the 19 HMETS parameters, uniform +-20 %
        t_setup = time.time()
        model = api.RavenModel(base)
        model.flatten(E)
        sim = api.GpuSimulation(model, n_members=E, device=local_rank)
        # Monte-Carlo members exactly as the reference's member loop would draw them (CMonteCarloEnsemble::UpdateModel,
        # one shared rand() stream): this rank owns global members rank*E .. rank*E+E-1
        from ravenhydroframework_b200 import ensemble
        member0, n_local = ensemble.shard_members(E * world, rank, world)
        assert n_local == E
        sim.apply_ensemble(member0)
        setup_s = time.time() - t_setup
        nH, NS, NB = model.nHRU, model.NS, model.nSB
        units_per_step = E * nH  # HRU-steps per step on this rank

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        # ---- resident pass: W warm-up steps, then exactly K timed steps
        sim.set_recording(n_days, hydrographs=False, watershed=True)
        sim.set_kernel_timing(True)
        clocks = ClockSampler(local_rank)
        clocks.start()          # sampling runs from before the warm-up; only samples inside the timed region are reported
        sim.run(W)
        barrier()
        tw0 = time.time()
        t0 = time.perf_counter()
        sim.run(K)              # asynchronous launches on the engine's stream ...
        stats = sim.last_run_stats()   # ... synchronises on that stream and reads the CUDA events
        barrier()
        wall_ms = 1000.0 * (time.perf_counter() - t0)
        clk = clocks.stop(tw0, time.time())
        dev_ms = stats["total_ms"]
        sweep_ms = stats["sweep_ms"]
        launches = stats["launches"]
        sim.set_kernel_timing(False)

        # ---- end-to-end pass through the host-facing ABI, one step per call, host buffers pinned
        nG = model.nGauges
        series = torch.zeros((9, nG, n_days), dtype=torch.float64).pin_memory()
        # the host copy of the forcing series is re-sent step by step (values identical to what the model parsed)
        import ctypes as C
        desc_series = (C.c_double * (9 * nG * model.nSteps)).from_address(_desc_series_ptr(model))
        series.view(-1)[:] = torch.from_numpy(np.ctypeslib.as_array(desc_series).copy())
        hyd = torch.zeros((E, NB), dtype=torch.float64).pin_memory()
        wsh = torch.zeros((E, 4), dtype=torch.float64).pin_memory()
        lib = sim.lib
        step_buf = torch.zeros((9, nG, 1), dtype=torch.float64).pin_memory()
        sim.set_recording(n_days, hydrographs=True, watershed=True)

        def e2e_step2(i):
            s = sim.step
            step_buf[:, :, 0] = series[:, :, s]
            sim._ck(lib.rvn_gpu_upload_forcings(sim.h, step_buf.data_ptr(), None, s, 1))
            sim._ck(lib.rvn_gpu_run(sim.h, s, 1))
            sim.step += 1
            sim._ck(lib.rvn_gpu_download_hydrographs(sim.h, hyd.data_ptr(), i, 1, 0, E))
            sim._ck(lib.rvn_gpu_download_watershed(sim.h, wsh.data_ptr(), i, 1, 0, E))
        for i in range(W):
            e2e_step2(i)
        barrier()
        t1 = time.perf_counter()
        for i in range(W, W + K):
            e2e_step2(i)
        barrier()
        e2e_sync_ms = 1000.0 * (time.perf_counter() - t1)
        h2d = 9 * nG * 8
        d2h = E * NB * 8 + E * 4 * 8

        # ---- end-to-end pass, streaming form of the same ABI (rvn_gpu_step_async): per step the forcing slab is copied from pinned
        # host memory, the step runs, and the per-basin outflows + watershed record are copied back into that step's slot of a
        # pinned host array; nothing waits until the final synchronisation, so the copies overlap the next step's sweep
        slabs = torch.zeros((n_days, 9, nG), dtype=torch.float64).pin_memory()
        slabs[:] = series.permute(2, 0, 1)
        hyd_all = torch.zeros((W + K, E, NB), dtype=torch.float64).pin_memory()
        wsh_all = torch.zeros((W + K, E, 4), dtype=torch.float64).pin_memory()
        sb, hb, wb = slabs.data_ptr(), hyd_all.data_ptr(), wsh_all.data_ptr()

        def e2e_stream(i):
            s_ = sim.step
            sim.step_async(sb + s_ * 9 * nG * 8, hb + i * E * NB * 8, wb + i * E * 4 * 8)
        for i in range(W):
            e2e_stream(i)
        sim.sync()
        barrier()
        t1 = time.perf_counter()
        for i in range(W, W + K):
            e2e_stream(i)
        sim.sync()              # every result of the K steps is in host memory here
        barrier()
        e2e_ms = 1000.0 * (time.perf_counter() - t1)
        assert bool(torch.isfinite(hyd_all[W + K - 1]).all()) and float(hyd_all[W + K - 1].abs().sum()) > 0.0

        # ---- max over ranks
        t = torch.tensor([dev_ms, e2e_ms, sweep_ms, wall_ms, e2e_sync_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms, sweep_ms, wall_ms, e2e_sync_ms = [float(x) for x in t.tolist()]
        total_units = units_per_step * world
        value = total_units * K / (dev_ms / 1000.0)
        e2e_value = total_units * K / (e2e_ms / 1000.0)

        # ---- roofline of the dominant kernel (hru_sweep_kernel), per launch
        peak, peak_src = measured_peak()
        live = _live_stores(model, E)
        b_alg_live = 16.0 * (model.nCore + live) + 8.0      # state in+out of the live variables + surface-water volume
        b_alg_full = 16.0 * NS + 8.0                          # reference layout: every state variable in and out
        sweep_per_launch_ms = sweep_ms / K
        achieved = units_per_step * b_alg_live / (sweep_per_launch_ms / 1000.0) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": ncu_traffic(sim.kernel_variant(), units_per_step),
                    "kernel": sim.kernel_variant(), "kernel_ms_per_launch": sweep_per_launch_ms, "kernel_share_of_step": sweep_ms / dev_ms,
                    "bytes_per_hru_step": b_alg_live, "bytes_per_hru_step_reference_layout": b_alg_full, "live_conv_stores_mean": live,
                    "peak_source": peak_src}

        line = {"metric": METRIC, "value": value, "unit": "HRU-steps/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": dev_ms / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "roofline": roofline, "e2e": {"value": e2e_value, "unit": "HRU-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                                             "ms_per_step": e2e_ms / K, "api": "rvn_gpu_step_async (streaming) + rvn_gpu_sync",
                                             "synchronous_api_ms_per_step": e2e_sync_ms / K},
                "gpu_launches": int(launches), "clocks": clk, "wall_ms_timed_region": wall_ms, "setup_s": setup_s}
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            with tempfile.TemporaryDirectory() as td2:
                r = run_reference_cpu(CPU_SAMPLE_STEPS, min(ncores, 16), td2)
            if r is not None:
                line["cpu_baseline"] = {"value": r["value"], "unit": "HRU-steps/s", "cores": r["nproc"], "kind": "reference",
                                        "sample": "%d concurrent single-threaded Raven.exe processes, one 10k-HRU HMETS member each, %d daily steps "
                                                  "(%.1f s wall incl. parsing); per-process %.0f HRU-steps/s" % (r["nproc"], CPU_SAMPLE_STEPS, r["wall_s"], r["per_proc"])}
        # ---- final ensemble statistics over ALL members (outside the timed regions): outlet hydrograph of the end-to-end pass,
        # reduced with one NCCL all-reduce per statistic (sum/sumsq, min, max)
        q_out = torch.from_numpy(sim.hydrographs(W, K)[:, :, 0].T.copy()).cuda()     # [members, K]
        torch.cuda.synchronize()
        ts = time.perf_counter()
        stats = ensemble.ensemble_statistics(q_out)
        torch.cuda.synchronize()
        line["ensemble_statistics"] = {"members": stats["n"], "outlet_q_mean_last_step": float(stats["mean"][-1]), "outlet_q_std_last_step": float(stats["std"][-1]),
                                       "ms": 1000.0 * (time.perf_counter() - ts), "backend": "nccl" if world > 1 else "local"}
        if rank == 0:
            sys.stdout.flush()
            os.write(json_fd, (json.dumps(line) + "\n").encode())
        sim.close()
    finally:
        shutil.rmtree(td, ignore_errors=True)
        if world > 1:
            dist.destroy_process_group()
    return 0


def _desc_series_ptr(model):
    """Address of rvn_model_desc.gauge_series (host copy of the sampled forcing series)."""
    import ctypes as C
    lib = model.lib
    lib.rvh_desc_gauge_series.restype = C.c_void_p
    lib.rvh_desc_gauge_series.argtypes = [C.c_void_p]
    return lib.rvh_desc_gauge_series(model.desc)


def _live_stores(model, E):
    """Mean number of live convolution stores per (member, HRU): sum over convolution processes of N(member, class)."""
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
    # single land-use class in the bench workload -> class 0 of every member
    return float(arr[:, 0, :].sum(axis=1).mean())


if __name__ == "__main__":
    sys.exit(main())
