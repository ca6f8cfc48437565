#!/usr/bin/env python
"""tools/enkf_multigpu_check.py -- run under torchrun on N GPUs: the golden EnKF cases with members sharded over the ranks
(NCCL all-gather of the state rows on the device), every rank checks its own members against the reference's matrices.
Prints one line per rank; exit code 0 only if every check passes on every rank."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import helpers as H
    from ravenhydroframework_b200 import api, enkf
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for case in ("enkf_hbv", "enkf_hmets"):
        ref = dict(np.load(os.path.join(H.GOLDEN, case, "reference_enkf.npz")))
        model = api.RavenModel(os.path.join(H.GOLDEN, case, "model"))
        sim = enkf.EnKFSimulation(model, device=local, rank=rank, world=world)
        out = sim.forecast()
        sl = slice(sim.member0, sim.member0 + sim.n_local)
        e_out = H.rel_err(out, ref["out"][sl])
        pre, post = sim.analysis()
        e_pre, e_post = H.rel_err(pre, ref["Xpre"][sl]), H.rel_err(post, ref["Xpost"][sl])
        good = max(e_out, e_pre, e_post) < 1e-9
        ok = ok and good
        print("rank %d/%d %s members [%d,%d): out %.2e  Xpre %.2e  Xpost %.2e  %s" % (rank, world, case, sim.member0, sim.member0 + sim.n_local, e_out, e_pre, e_post,
                                                                                   "OK" if good else "FAIL"), flush=True)
        sim.sim.close()
    flag = torch.tensor([0 if ok else 1], device=torch.device("cuda", local))
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 0 else 1)


if __name__ == "__main__":
    main()
