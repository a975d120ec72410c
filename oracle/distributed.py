"""x-slab driver for the CPU oracle: sequences the oracle's stages (include/ocean_b200.h :: ob200_stage) and
exchanges halos through torch.distributed (gloo on CPU).  TEST INFRASTRUCTURE ONLY -- it exercises the host-side
partition / halo logic of the N > 1 path without a GPU.  A slab run must reproduce the single-domain run bit for bit.

The product exchanges a wide barotropic halo once per step (NCCL, inside libocean_b200.so); this driver uses the
other legal schedule -- one-column exchanges after every half substep -- so the two cross-check each other.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

HALO = 7
FID = {"U": 0, "V": 1, "T": 3, "S": 4, "ETA": 16, "UTRANS": 19, "VTRANS": 20}
ST = {"BAROTROPIC_FORCING": 0, "AB2_IMPLICIT": 1, "BAROTROPIC_INIT": 2, "BAROTROPIC_ETA": 3, "BAROTROPIC_VEL": 4,
      "BAROTROPIC_FINISH": 5, "CORRECT_AND_STORE": 6, "MASK_AND_LOCAL_HALOS": 7, "AUXILIARIES": 8, "TENDENCIES": 9}


class SlabStepper:
    def __init__(self, model):
        self.m = model
        self.b = model.backend
        g = model.grid
        self.Nx, self.Ny, self.Nz = g.size()
        self.rank, self.n = g.rank, g.nranks
        self.west, self.east = (self.rank - 1) % self.n, (self.rank + 1) % self.n
        self.nsub = len(model.free_surface.averaging_weights)

    def _count(self, name, width):
        ny = self.Ny + 1 if name in ("V", "VTRANS") else self.Ny
        nz = 1 if name in ("ETA", "UTRANS", "VTRANS") else self.Nz
        return width * ny * nz

    def exchange(self, names, width):
        """Fill the `width` x-halo columns of the named fields from the ring neighbours."""
        if self.n == 1:
            return
        ops, recvs = [], []
        for t, name in enumerate(names):
            n = self._count(name, width)
            wedge = torch.from_numpy(self.b.get_strip(FID[name], n, 0, width, True))   # my west edge -> west neighbour
            eedge = torch.from_numpy(self.b.get_strip(FID[name], n, 1, width, True))   # my east edge -> east neighbour
            whalo, ehalo = torch.empty(n, dtype=torch.float64), torch.empty(n, dtype=torch.float64)
            ops += [dist.P2POp(dist.isend, eedge, self.east, tag=2 * t), dist.P2POp(dist.isend, wedge, self.west, tag=2 * t + 1),
                    dist.P2POp(dist.irecv, whalo, self.west, tag=2 * t), dist.P2POp(dist.irecv, ehalo, self.east, tag=2 * t + 1)]
            recvs.append((name, whalo, ehalo))
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        for name, whalo, ehalo in recvs:
            self.b.set_strip(FID[name], whalo.numpy(), 0, width, False)
            self.b.set_strip(FID[name], ehalo.numpy(), 1, width, False)

    def update_state(self):
        self.b.run_stage(ST["MASK_AND_LOCAL_HALOS"], 0.0, 1)    # mask + y/z halos
        self.exchange(["U", "V", "T", "S", "ETA"], HALO)
        self.b.run_stage(ST["MASK_AND_LOCAL_HALOS"], 0.0, 1)    # y/z halos of the received columns (corners)
        self.b.run_stage(ST["AUXILIARIES"])
        self.b.run_stage(ST["TENDENCIES"])

    def time_step(self, dt):
        _, it, last_dt = self.b.clock()
        if it == 0:
            self.update_state()
        if dt != last_dt:
            self.b.zero_prev_tendencies()
        self.b.run_stage(ST["BAROTROPIC_FORCING"], dt)
        self.b.run_stage(ST["AB2_IMPLICIT"], dt)
        self.b.run_stage(ST["BAROTROPIC_INIT"], dt)
        self.exchange(["UTRANS", "VTRANS"], 1)
        for m in range(self.nsub):
            self.b.run_stage(ST["BAROTROPIC_ETA"], dt, m)
            self.exchange(["ETA"], 1)
            self.b.run_stage(ST["BAROTROPIC_VEL"], dt, m)
            self.exchange(["UTRANS", "VTRANS"], 1)
        self.b.run_stage(ST["BAROTROPIC_FINISH"], dt)
        self.b.run_stage(ST["CORRECT_AND_STORE"], dt)
        self.b.tick(dt)
        self.update_state()


def gather_global(model, name, shape_fn):
    """Concatenate every rank's interior slab of `name` along x on rank 0 (returns None elsewhere)."""
    a = model.get(name)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return a
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, a)
    return np.concatenate(parts, axis=-1)
