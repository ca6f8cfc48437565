"""EnKF streamflow/state assimilation across GPUs (reference CEnKFEnsemble, src/EnKF.cpp).

Members shard over ranks in contiguous blocks and advance through the assimilation window without any exchange.  At the end
of the window the analysis needs every member once:
  1. simulated observations  HA rows [n_local][Nobs]  -> all-gather (tiny)      -> Z [N][N] on every rank (host, redundant)
  2. state rows              X  rows [n_local][M]     -> all-gather ON DEVICE (NCCL over NVLink) -> X_all [N][M]
  3. every rank updates its own members' rows with the device kernel (csrc/rvn_enkf.cuh) and writes max(Xa, 0) into its state.
With no process group (or world size 1) the gathers are identities.  torch is used for the collectives and the device
buffers only.
"""
import numpy as np

from .ensemble import shard_members


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist
    except ImportError:
        pass
    return None


def gather_member_rows(local_rows, n_members, group=None):
    """All-gather of per-member rows [n_local, ...] (torch tensor, any device) into [N, ...] in global member order.  Blocks may
    differ by one member (shard_members), so every rank pads to the largest block."""
    import torch
    dist = _dist()
    if dist is None:
        return local_rows
    world = dist.get_world_size(group)
    sizes = [shard_members(n_members, r, world)[1] for r in range(world)]
    mx = max(sizes)
    pad = local_rows
    if local_rows.shape[0] < mx:
        pad = torch.cat([local_rows, torch.zeros((mx - local_rows.shape[0],) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype, device=local_rows.device)])
    out = torch.empty((world * mx,) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype, device=local_rows.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    if all(s == mx for s in sizes):
        return out
    return torch.cat([out[r * mx:r * mx + sizes[r]] for r in range(world)])


class EnKFSimulation:
    """One assimilation window of an ENSEMBLE_ENKF model on this rank's GPU: forecast of the local members with their forcing
    perturbations, then the analysis.  `model` is an api.RavenModel of the EnKF input set."""

    def __init__(self, model, device=0, rank=0, world=1):
        from . import api
        self.model = model
        self.N = model.num_members
        if world > self.N:
            raise ValueError("EnKFSimulation: %d ranks for %d ensemble members - every rank needs at least one member "
                             "(launch with --nproc-per-node <= %d)" % (world, self.N, self.N))
        self.member0, self.n_local = shard_members(self.N, rank, world)
        model.flatten(self.n_local)
        self.info = model.enkf_begin(self.member0, self.n_local)
        self.sim = api.GpuSimulation(model, n_members=self.n_local, device=device)
        self.sim.enkf_configure(model.enkf_rows())
        self.nsteps = model.nSteps
        self.sim.set_recording(self.nsteps, hydrographs=True, watershed=False)
        # CSubBasin::GetOutflowRate before the first step: last river segment of every basin
        import ctypes as C
        nseg = (C.c_int32 * model.nSB)()
        api.host_lib().rvh_desc_nseg(model.desc, nseg)
        qout = model.basin_state()[2]
        self.qout_initial = np.array([qout[p, nseg[p] - 1] for p in range(model.nSB)])
        self.device = device

    def forecast(self):
        """Run every local member over the window; returns the simulated observations [n_local][Nobs] (CloseTimeStepOps)."""
        self.sim.run(self.nsteps)
        q = self.sim.hydrographs(0, self.nsteps)
        self.out_local = self.model.enkf_harvest_outputs(q, self.qout_initial)
        return self.out_local

    def analysis(self):
        """AssimilationCalcs + UpdateFromStateMatrix.  Returns (X_pre_local, X_post_local) as numpy arrays [n_local][M]."""
        import torch
        dist = _dist()
        if dist is None:
            Z = self.model.enkf_compute_Z(self.out_local)
            X = self.sim.enkf_pack()
            pre = X.copy()
            if Z is None:
                return pre, pre
            post = self.sim.enkf_update(X, Z, 0).copy()
            return pre, post
        dev = torch.device("cuda", self.device)
        out_all = gather_member_rows(torch.from_numpy(self.out_local).to(dev), self.N).cpu().numpy()
        Z = self.model.enkf_compute_Z(out_all)
        M = self.info["M"]
        x_local = torch.empty((self.n_local, M), dtype=torch.float64, device=dev)
        self.sim.enkf_pack_device(x_local.data_ptr())
        pre = x_local.cpu().numpy()
        if Z is None:
            return pre, pre
        x_all = gather_member_rows(x_local, self.N).contiguous()
        torch.cuda.synchronize(dev)
        self.sim.enkf_update_device(x_all.data_ptr(), Z, self.N, self.member0)
        post = x_all[self.member0:self.member0 + self.n_local].cpu().numpy()
        return pre, post
