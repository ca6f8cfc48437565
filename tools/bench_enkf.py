#!/usr/bin/env python
"""tools/bench_enkf.py -- EnKF assimilation window on B200 at BASELINE config-5 shape (HBV-EC, E members x nHRU HRUs, SOIL[0] + SNOW
of every HRU and STREAMFLOW of every subbasin assimilated, gauged outlets, window 1): forecast of the local members with their
forcing perturbations, then the analysis (NCCL all-gather of the state rows over NVLink when launched with torchrun).

One JSON line on rank 0:
  forecast    HRU-steps/s of the perturbed-member forecast (all ranks), ms per step
  analysis    device ms of the update kernel, its algorithmic FP64 rate (2*N*M*N_local mul+add) and HBM/L2 streaming rate
              (N*M*8 B read N_local/ET times + N_local*M*8 written), all-gather ms, total analysis ms
  checks      size-independent properties at this size: Z = 0 leaves X unchanged (idempotence); the update of a random column
              sample equals the numpy restatement in oracle/enkf_oracle.py bit for bit (checker only).

python tools/bench_enkf.py [--members 512] [--sb 5000] [--hru-per-sb 20] [--days 10]
torchrun --nproc-per-node 2 ... tools/bench_enkf.py   (members sharded over ranks)
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--members", type=int, default=512)
    ap.add_argument("--sb", type=int, default=5000)
    ap.add_argument("--hru-per-sb", type=int, default=20)
    ap.add_argument("--days", type=int, default=10)
    ap.add_argument("--gauges", type=int, default=10)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from ravenhydroframework_b200 import api, enkf, synth
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nH = a.sb * a.hru_per_sb
    td = tempfile.mkdtemp(prefix="enkf_bench_r%d_" % rank)
    base = synth.write_hbv(td, n_sb=a.sb, hru_per_sb=a.hru_per_sb, n_days=a.days, seed=31, routing="ROUTE_MUSKINGUM", name="model", season_offset=130)
    obs_basins = tuple(range(1, a.gauges + 1))
    synth.add_enkf(base, n_members=a.members, n_sb=a.sb, n_hru=nH, n_days=a.days, obs_basins=obs_basins, window=1)
    t0 = time.time()
    model = api.RavenModel(base)
    sim = enkf.EnKFSimulation(model, device=local, rank=rank, world=world)
    setup_s = time.time() - t0
    N, M, nloc = sim.N, sim.info["M"], sim.n_local
    dev = torch.device("cuda", local)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- forecast
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    sim.sim.run(sim.nsteps)
    q = sim.sim.hydrographs(0, sim.nsteps)
    barrier()
    fc_ms = (time.time() - t0) * 1e3
    stats = sim.sim.last_run_stats()
    sim.out_local = model.enkf_harvest_outputs(q, sim.qout_initial)
    # ---- analysis (timed pieces)
    t0 = time.time()
    out_all = enkf.gather_member_rows(torch.from_numpy(sim.out_local).to(dev), N).cpu().numpy()
    Z = model.enkf_compute_Z(out_all)
    z_ms = (time.time() - t0) * 1e3
    x_local = torch.empty((nloc, M), dtype=torch.float64, device=dev)
    sim.sim.enkf_pack_device(x_local.data_ptr())
    barrier()
    ev0.record()
    x_all = enkf.gather_member_rows(x_local, N).contiguous()
    ev1.record()
    barrier()
    gather_ms = ev0.elapsed_time(ev1)
    x_pre = x_all.clone()
    # property: Z = 0 -> no change
    zero = x_all.clone()
    sim.sim.enkf_update_device(zero.data_ptr(), np.zeros((N, N)), N, sim.member0)
    idem = bool(torch.equal(zero[sim.member0:sim.member0 + nloc], x_pre[sim.member0:sim.member0 + nloc]))
    t0 = time.time()
    sim.sim.enkf_update_device(x_all.data_ptr(), Z, N, sim.member0)
    torch.cuda.synchronize(dev)
    upd_wall_ms = (time.time() - t0) * 1e3
    upd_ms = sim.sim.enkf_last_ms()
    # checker (oracle, a column sample): identical bits expected
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import enkf_oracle as EO
    cols = np.random.RandomState(1).choice(M, size=min(M, 512), replace=False)
    obs, noise, _, _ = model.enkf_observations()
    Xs = x_pre[:, torch.from_numpy(cols).to(dev)].cpu().numpy()
    Xa_ref, Zo = EO.assimilation_calcs(Xs, out_all, obs, noise)
    # use the product's Z for the column check so that only the update kernel is compared
    XdT = np.zeros((len(cols), N))
    Xmean = np.zeros(len(cols))
    for i in range(N):
        Xmean += Xs[i] / N
    A = (Xs - Xmean[None, :]).T
    for i in range(N):
        XdT += np.outer(A[:, i], Z[i, :])
    Xa_prod = Xs + (XdT * (1.0 / (N - 1))).T
    got = x_all[sim.member0:sim.member0 + nloc][:, torch.from_numpy(cols).to(dev)].cpu().numpy()
    bitexact = bool(np.array_equal(got, Xa_prod[sim.member0:sim.member0 + nloc]))
    rel_vs_oracle = float(np.max(np.abs(got - Xa_ref[sim.member0:sim.member0 + nloc]) / np.maximum(np.abs(Xa_ref[sim.member0:sim.member0 + nloc]), 1e-3)))
    # ---- report (max over ranks)
    vals = torch.tensor([fc_ms, upd_ms, gather_ms, upd_wall_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    fc_ms, upd_ms, gather_ms, upd_wall_ms = [float(v) for v in vals.cpu()]
    if rank == 0:
        ntile = (nloc + 15) // 16
        flop = 2.0 * N * M * nloc
        bytes_ = N * M * 8.0 * ntile + N * M * 8.0 + nloc * M * 8.0
        print(json.dumps({
            "workload": "EnKF window, HBV-EC %d members x %d HRUs / %d subbasins, %d daily steps, %d gauges, state matrix M=%d rows" % (N, nH, a.sb, sim.nsteps, a.gauges, M),
            "n_gpus": world, "members_per_gpu": nloc, "setup_s": setup_s,
            "forecast": {"ms_total": fc_ms, "ms_per_step": fc_ms / sim.nsteps, "hru_steps_per_s": N * nH * sim.nsteps / (fc_ms * 1e-3),
                         "device_ms": stats["total_ms"], "kernel": sim.sim.kernel_variant()},
            "analysis": {"update_kernel_ms": upd_ms, "update_call_ms": upd_wall_ms, "allgather_X_ms": gather_ms, "host_Z_ms": z_ms,
                         "X_bytes": N * M * 8, "fp64_gflops": flop / (upd_ms * 1e-3) / 1e9, "stream_GBs": bytes_ / (upd_ms * 1e-3) / 1e9},
            "checks": {"zero_Z_is_identity": idem, "update_bitexact_vs_restatement_on_512_columns": bitexact, "rel_err_vs_oracle_full_analysis": rel_vs_oracle},
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
