#!/usr/bin/env python
"""tools/bench_routing.py -- routing at depth, one member (E = 1): ms per step and kernel launches per step of the whole step loop for
  tree500    HBV-EC, 10,000 HRUs / 500 subbasins, binary tree (9 levels), ROUTE_MUSKINGUM        (BASELINE config 2)
  chain500   the same model on a 500-deep chain (500 levels of one basin each)
  tree50k    Blended, 1,000,000 HRUs / 50,000 subbasins, binary tree (16 levels), hourly           (BASELINE config 4, one station)
Each case is also compared with the oracle on its first steps (checker only, outside the timing).  One JSON line."""
import json, os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from ravenhydroframework_b200 import api, synth


def case(name, steps):
    td = tempfile.mkdtemp()
    synth.TOPOLOGY = "chain" if name == "chain500" else "tree"
    if name == "tree50k":
        base = synth.write_blended(td, n_sb=50000, hru_per_sb=20, n_days=(steps + 23) // 24 + 1, seed=4, routing="ROUTE_MUSKINGUM", catchment_route="TRIANGULAR_UH", season_offset=124)
        rvi = open(base + ".rvi").read()
        open(base + ".rvi", "w").write(rvi.replace(":TimeStep              1.0", ":TimeStep              01:00:00"))
    else:
        base = synth.write_hbv(td, n_sb=500, hru_per_sb=20, n_days=steps + 1, seed=2, routing="ROUTE_MUSKINGUM", season_offset=100)
    synth.TOPOLOGY = "tree"
    model = api.RavenModel(base)
    model.flatten(1)
    sim = api.GpuSimulation(model)
    ncheck = 6 if name == "tree50k" else 40
    sim.set_recording(ncheck)
    sim.run(ncheck)
    ref = H.oracle_run(model, ncheck, record_state=False)
    err = max(H.rel_err(sim.download_state()[0], ref["state"]), H.rel_err(sim.hydrographs(0, ncheck)[:, 0, :], ref["qout_rec"]))
    sim.set_recording(0)
    sim.run(5)
    t0 = time.perf_counter()
    sim.run(steps - ncheck - 5)
    st = sim.last_run_stats()
    wall = (time.perf_counter() - t0) * 1e3
    n = steps - ncheck - 5
    out = {"hrus": model.nHRU, "subbasins": model.nSB, "kernel": sim.kernel_variant(), "steps": n, "ms_per_step": st["total_ms"] / n, "wall_ms_per_step": wall / n,
           "launches_per_step": st["launches"] / float(n), "hru_steps_per_s": model.nHRU * n / (st["total_ms"] * 1e-3), "rel_err_vs_oracle": err}
    sim.close()
    return out


if __name__ == "__main__":
    res = {}
    for name, steps in (("tree500", 2045), ("chain500", 2045), ("tree50k", 107)):
        res[name] = case(name, steps)
    print(json.dumps(res))
