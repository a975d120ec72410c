"""Host-side mirror of /root/reference/src/near_global_simulation.jl and experiments/run.jl.

`scaling_test_simulation` keeps the reference's signature and defaults; `time_step!` is
the C-ABI call into libocean_b200.so.  MPI is replaced by torch.distributed (bootstrap,
min/sum all-reduce of scalars) + NCCL inside the library (halo exchange).
"""
from __future__ import annotations

import json
import os
import time as _time
from dataclasses import dataclass, field
from typing import Callable, Optional

import numpy as np

from .bathymetry_and_forcings import exponential_z_faces
from .boundary_and_initial_conditions import initialize_model, set_boundary_conditions
from .grid_load_balance import load_balanced_grid
from .model import (HydrostaticFreeSurfaceModel, QGLeith, RiBasedVerticalDiffusivity, SplitExplicitFreeSurface,
                    VerticalScalarDiffusivity, barotropic_substeps, substeps, time_step)

minutes = 60.0
days = 86400.0


def experiment_depth(E: str) -> float:
    """near_global_simulation.jl:26"""
    return 5244.5 if E == "RealisticOcean" else 3000


def equation_of_state(experiment: str) -> str:
    """near_global_simulation.jl:23-24: TEOS10 except DoubleDrake (linear)."""
    return "linear" if experiment == "DoubleDrake" else "teos10"


class Comm:
    """What the reference gets from MPI.COMM_WORLD, on top of torch.distributed (if initialised)."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist if (dist.is_available() and dist.is_initialized()) else None

    @property
    def rank(self):
        return self.dist.get_rank() if self.dist else 0

    @property
    def size(self):
        return self.dist.get_world_size() if self.dist else 1

    def allreduce(self, x: float, op: str) -> float:
        if not self.dist:
            return x
        import torch
        dev = "cuda" if self.dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN if op == "min" else self.dist.ReduceOp.SUM)
        return float(t.item())

    def bcast_bytes(self, b: Optional[bytes]) -> bytes:
        if not self.dist:
            return b
        obj = [b]
        self.dist.broadcast_object_list(obj, src=0)
        return obj[0]

    def barrier(self):
        if self.dist:
            self.dist.barrier()


@dataclass
class Simulation:
    """Simulation(model; Δt, stop_time) with the reference's callbacks (near_global_simulation.jl:134-162)."""
    model: HydrostaticFreeSurfaceModel
    dt: float
    stop_time: float
    max_dt: float = 0.0
    min_dt: float = 0.0
    stop_iteration: Optional[int] = None
    wizard_cfl: float = 0.35
    wizard_max_change: float = 1.1
    progress: Optional[Callable] = None
    run_wall_time: float = 0.0


def run(sim: Simulation) -> None:
    """Oceananigans `run!` loop: TimeStepWizard every 10 iterations, progress every 100."""
    m = sim.model
    t0 = _time.time()
    while True:
        clk = m.clock
        if clk["time"] >= sim.stop_time or (sim.stop_iteration is not None and clk["iteration"] >= sim.stop_iteration):
            break
        if clk["iteration"] % 10 == 0 and sim.max_dt > sim.min_dt:
            _, cfl_dt = m.backend.diagnostics()
            new_dt = sim.wizard_cfl * cfl_dt
            new_dt = min(sim.wizard_max_change * sim.dt, new_dt)
            new_dt = max(0.5 * sim.dt, new_dt)           # min_change default 0.5
            sim.dt = float(min(max(new_dt, sim.min_dt), sim.max_dt))
        if clk["iteration"] % 100 == 0 and sim.progress is not None:
            sim.progress(sim)
        time_step(m, sim.dt)
    m.backend.sync()
    sim.run_wall_time = _time.time() - t0


def scaling_test_simulation(resolution, ranks, dt, stop_time, *, device: int = 0, experiment: str = "Quiescent",
                            Depth: Optional[float] = None, latitude=(-75, 75), restart: str = "",
                            z_faces_function=exponential_z_faces, Nz: int = 100, profile: bool = False,
                            with_fluxes: bool = True, with_restoring: bool = True, loadbalance: bool = True,
                            boundary_layer_parameterization=None, extra_closures=(), backend_factory=None,
                            realistic_inputs=None, profile_steps=(10, 10, 30)):
    """near_global_simulation.jl:36-176.  `ranks` = (Rx, 1, 1); this process's slab is rank `comm.rank`."""
    comm = Comm()
    if Depth is None:
        Depth = experiment_depth(experiment)
    min_dt, max_dt = (dt, dt) if np.isscalar(dt) else dt

    Lphi = latitude[1] - latitude[0]
    Nx = int(360 * resolution)
    Ny = int(Lphi * resolution)
    z_faces = z_faces_function(Nz, Depth)

    bathy = realistic_inputs["bathymetry"] if realistic_inputs else None
    grid = load_balanced_grid(comm.rank, ranks[0], (Nx, Ny, Nz), latitude, z_faces, resolution, loadbalance,
                              experiment, realistic_bathymetry=bathy)

    # physics (near_global_simulation.jl:70-89)
    vertical_diffusivity = VerticalScalarDiffusivity(nu=5e-4, kappa=3e-5)
    if boundary_layer_parameterization is None:
        boundary_layer_parameterization = RiBasedVerticalDiffusivity()
    nsub = barotropic_substeps(max_dt, grid, allreduce_min=lambda x: comm.allreduce(x, "min"))
    free_surface = SplitExplicitFreeSurface(substeps=nsub)
    if comm.rank == 0:
        print(f"[ Info: running with {substeps(free_surface)} barotropic substeps", flush=True)
    closure = (vertical_diffusivity, boundary_layer_parameterization) + tuple(extra_closures)

    forcing = None
    if experiment == "RealisticOcean" and realistic_inputs and with_fluxes:
        forcing = realistic_inputs["forcing"](grid)
    boundary_conditions = set_boundary_conditions(experiment, grid, with_fluxes=with_fluxes,
                                                  with_restoring=with_restoring, forcing_arrays=forcing)

    model = HydrostaticFreeSurfaceModel(grid=grid, free_surface=free_surface, eos=equation_of_state(experiment),
                                        closure=closure, boundary_conditions=boundary_conditions, device=device,
                                        backend_factory=backend_factory, comm=comm)

    if profile:
        return profiled_time_steps(model, max_dt, resolution, comm=comm, warmup=profile_steps[0],
                                   gc_steps=profile_steps[1], profiled_steps=profile_steps[2])

    kw = {}
    if experiment == "RealisticOcean" and realistic_inputs:
        T0, S0 = realistic_inputs["initial"](grid)
        kw = dict(initial_T=T0, initial_S=S0)
    initialize_model(model, experiment, restart=restart, **kw)

    sim = Simulation(model, dt=min_dt, stop_time=stop_time, max_dt=max_dt, min_dt=min_dt)

    def progress(s):
        mx, _ = s.model.backend.diagnostics()
        c = s.model.clock
        print("R: %02d, T: %12.1f s, it: %d, dt: %.2f, vels: %.2e %.2e %.2e, eta: %.2e, trac: %.2e %.2e" %
              (comm.rank, c["time"], c["iteration"], s.dt, *mx), flush=True)

    sim.progress = progress
    return sim


def profiled_time_steps(model, dt, resolution, *, comm=None, warmup=10, gc_steps=10, profiled_steps=30,
                        outdir: str = "."):
    """near_global_simulation.jl:178-216: warm-up steps, then gc_steps x profiled_steps timed steps, mean
    wall time per step averaged over ranks.  Unlike the reference, the device is synchronised around the
    timed region (the reference's @elapsed has no sync).  Returns seconds per step."""
    comm = comm or Comm()
    for _ in range(warmup):
        time_step(model, dt)
    model.backend.sync()
    elapsed = 0.0
    for _ in range(gc_steps):
        t0 = _time.perf_counter()
        model.backend.time_steps(dt, profiled_steps)   # each step carries the NVTX range "one time step"
        model.backend.sync()
        elapsed += _time.perf_counter() - t0
    elapsed = elapsed / comm.size / gc_steps / profiled_steps
    elapsed = comm.allreduce(elapsed, "sum")
    if comm.rank == 0:
        name = f"time_res{resolution}_NZ{model.grid.Nz}_ranks{comm.size}_precFloat64.json"
        path = os.path.join(outdir, name)
        while os.path.isfile(path):
            path = os.path.join(outdir, "new_" + os.path.basename(path))
        with open(path, "w") as f:
            json.dump({"elapsed_time": elapsed}, f)
    comm.barrier()
    return elapsed
