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
from .boundary_and_initial_conditions import (initialize_model, set_boundary_conditions, update_fluxes,
                                              update_restoring)
from .grid_load_balance import load_balanced_grid
from .model import (HydrostaticFreeSurfaceModel, RiBasedVerticalDiffusivity, SplitExplicitFreeSurface,
                    VerticalScalarDiffusivity, barotropic_substeps, substeps, time_step)
from .simulation_outputs import IterationInterval, TimeInterval, set_from_checkpoint

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
class Callback:
    """Oceananigans `Callback(func, schedule)`."""
    func: Callable
    schedule: object

    def __call__(self, sim):
        return self.func(sim)


@dataclass
class Simulation:
    """Simulation(model; Δt, stop_time) with the reference's callbacks (near_global_simulation.jl:134-173) and output
    writers (simulation_outputs.jl:29-40)."""
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
    callbacks: dict = field(default_factory=dict)
    output_writers: dict = field(default_factory=dict)
    align_time_step: bool = True
    initialized: bool = False
    comm: Optional[Comm] = None        # x-slab runs: the wizard's time scale is reduced over the ranks


def time_step_wizard(sim: Simulation) -> None:
    """TimeStepWizard(cfl=0.35; max_change=1.1, max_Δt, min_Δt) -- near_global_simulation.jl:160."""
    if not sim.max_dt > sim.min_dt:
        return
    _, cfl_dt = sim.model.backend.diagnostics()
    # the advective time scale is a minimum over the WHOLE domain (Oceananigans reduces it over the ranks): without this
    # every x-slab would pick its own dt, the clocks diverge and the next NCCL exchange pairs mismatched steps
    comm = sim.comm if sim.comm is not None else (Comm() if sim.model.grid.nranks > 1 else None)
    if comm is not None and comm.size > 1:
        cfl_dt = comm.allreduce(cfl_dt, "min")
    new_dt = sim.wizard_cfl * cfl_dt
    new_dt = min(sim.wizard_max_change * sim.dt, new_dt)
    new_dt = max(0.5 * sim.dt, new_dt)           # min_change default 0.5
    sim.dt = float(min(max(new_dt, sim.min_dt), sim.max_dt))


def _all_callbacks(sim: Simulation) -> list:
    """The reference's callbacks in insertion order: progress (every 100 iterations), wizard (every 10), then the
    experiment's (update_fluxes / update_restoring) and the user's."""
    cbs = []
    if sim.progress is not None:
        cbs.append(Callback(sim.progress, IterationInterval(100)))
    cbs.append(Callback(time_step_wizard, IterationInterval(10)))
    cbs.extend(sim.callbacks.values())
    return cbs


def aligned_time_step(sim: Simulation, dt: float) -> float:
    """Oceananigans `aligned_time_step`: shorten Δt so that the model lands exactly on the next actuation time of every
    TimeInterval schedule (callbacks and output writers) and on the stop time."""
    t = sim.model.clock["time"]
    aligned = dt
    for obj in list(sim.callbacks.values()) + list(sim.output_writers.values()):
        sched = obj.schedule
        if isinstance(sched, TimeInterval):
            gap = sched.next_actuation_time() - t
            if gap > 0:
                aligned = min(aligned, gap)
    aligned = min(aligned, sim.stop_time - t)
    return max(aligned, 0.0)


def initialize_simulation(sim: Simulation) -> None:
    """Oceananigans `initialize!(sim)`: every callback runs once and every writer writes the initial state."""
    m = sim.model
    for obj in list(sim.callbacks.values()) + list(sim.output_writers.values()):
        obj.schedule.initialize(m)
    for cb in _all_callbacks(sim):
        cb(sim)
    for w in sim.output_writers.values():
        w.write(m)
    sim.initialized = True


def run(sim: Simulation, pickup=False) -> None:
    """Oceananigans `run!(simulation; pickup)`: restore from a checkpoint file when `pickup` names one
    (experiments/run.jl:59-60), then step until stop_time / stop_iteration with callbacks and writers after each step."""
    m = sim.model
    if pickup:
        set_from_checkpoint(m, pickup)
    t0 = _time.time()
    if not sim.initialized:
        initialize_simulation(sim)
    cbs = _all_callbacks(sim)
    while True:
        clk = m.clock
        if clk["time"] >= sim.stop_time or (sim.stop_iteration is not None and clk["iteration"] >= sim.stop_iteration):
            break
        dt = aligned_time_step(sim, sim.dt) if sim.align_time_step else sim.dt
        time_step(m, dt)
        for cb in cbs:
            if cb.schedule(m):
                cb(sim)
        for w in sim.output_writers.values():
            if w.schedule(m):
                w.write(m)
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
    providers = {k: realistic_inputs[k] for k in ("fluxes", "restoring") if realistic_inputs and k in realistic_inputs}
    boundary_conditions = set_boundary_conditions(experiment, grid, with_fluxes=with_fluxes,
                                                  with_restoring=with_restoring, forcing_arrays=forcing,
                                                  providers=providers)

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

    sim = Simulation(model, dt=min_dt, stop_time=stop_time, max_dt=max_dt, min_dt=min_dt, comm=comm)

    def progress(s):
        mx, _ = s.model.backend.diagnostics()
        c = s.model.clock
        print("R: %02d, T: %12.1f s, it: %d, dt: %.2f, vels: %.2e %.2e %.2e, eta: %.2e, trac: %.2e %.2e" %
              (comm.rank, c["time"], c["iteration"], s.dt, *mx), flush=True)

    sim.progress = progress
    if experiment == "RealisticOcean":      # near_global_simulation.jl:163-172
        if with_fluxes:
            sim.callbacks["update_fluxes"] = Callback(update_fluxes, TimeInterval(5 * days))
        if with_restoring:
            repeat_year = os.environ.get("REPEATYEAR", "true").lower() in ("true", "1")
            sim.callbacks["update_restoring"] = Callback(update_restoring, TimeInterval((1 if repeat_year else 15) * days))
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
