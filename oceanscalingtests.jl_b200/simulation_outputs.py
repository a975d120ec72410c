"""Host-side mirror of /root/reference/src/simulation_outputs.jl: the two callers on the far side of `time_step!`
that read the model state -- the surface-field writer and the checkpointer -- and the restore path `set!(model, file)`.

File container.  The reference writes JLD2 (HDF5-style groups).  Neither JLD2 nor h5py exists here, so the same group
names are kept as entry names of an uncompressed zip of .npy members (a `.npz` that numpy loads directly):
    u/data, v/data, T/data, S/data, η/data,
    timestepper/Gⁿ/{u,v,T,S}/data, timestepper/G⁻/{u,v,T,S}/data, clock/time, clock/iteration
Arrays are Oceananigans *parent* arrays (interior + 7-cell halos, simulation_outputs.jl:54-57,91-99), stored C-ordered
as (k, j, i), i.e. the memory image of the Julia (i, j, k) column-major array.  `julia/OceanB200.jl` documents the
two-line JLD2 <-> npz conversion for a site that has Julia.

Beyond the reference's keys the checkpointer adds `b200/*` members (time-filtered diffusivities, barotropic averages,
the last Δt): Oceananigans' checkpoint does not hold them, so a reference restart is not bit-reproducible; with them a
restart from this library continues bit for bit (tests/test_outputs_cpu.py, tests/test_parity_gpu.py).
"""
from __future__ import annotations

import io
import os
import zipfile
from dataclasses import dataclass
from typing import Dict

import numpy as np

from . import _abi
from .grid_load_balance import HALO

days = 86400.0

_PROGNOSTIC = (("u", "U"), ("v", "V"), ("T", "T"), ("S", "S"))
_TENDENCY = {"u": ("GU", "GU_PREV"), "v": ("GV", "GV_PREV"), "T": ("GT", "GT_PREV"), "S": ("GS", "GS_PREV")}
_EXTRA = ("KAPPA_C", "KAPPA_U", "UBAR", "VBAR")


# ---------------------------------------------------------------------------------------
# container: named groups appended to one zip of .npy members
# ---------------------------------------------------------------------------------------
class GroupFile:
    """Append-only store of named arrays; `np.load(path)` reads it back (keys = group paths)."""

    def __init__(self, path: str, mode: str = "a"):
        self.path = path
        self.mode = mode

    def write(self, entries: Dict[str, np.ndarray]) -> None:
        with zipfile.ZipFile(self.path, self.mode, zipfile.ZIP_STORED, allowZip64=True) as zf:
            present = set(zf.namelist())
            dup = [key for key in entries if key + ".npy" in present]
            if dup:
                raise ValueError(f"{self.path}: group {dup[0]} already written")
            for key, arr in entries.items():
                name = key + ".npy"
                buf = io.BytesIO()
                np.lib.format.write_array(buf, np.asanyarray(arr), allow_pickle=False)
                zf.writestr(name, buf.getvalue())
        self.mode = "a"


def _parent(interior: np.ndarray, halo: int = HALO) -> np.ndarray:
    """Parent array (zero halos) around an interior array; a 2-D field gets a singleton k dimension."""
    a = interior[None] if interior.ndim == 2 else interior
    hz = 0 if interior.ndim == 2 else halo
    out = np.zeros((a.shape[0] + 2 * hz, a.shape[1] + 2 * halo, a.shape[2] + 2 * halo), dtype=np.float64)
    out[hz:hz + a.shape[0], halo:halo + a.shape[1], halo:halo + a.shape[2]] = a
    return out


def _interior(parent: np.ndarray, shape) -> np.ndarray:
    """Interior of a parent array whatever its halo width ("allow different halo size", simulation_outputs.jl:55)."""
    p = np.asarray(parent)
    if len(shape) == 2 and p.ndim == 3:
        p = p[p.shape[0] // 2]          # windowed field: singleton k
    shape3 = tuple(shape)
    off = [(ps - s) for ps, s in zip(p.shape, shape3)]
    if any(o < 0 or o % 2 for o in off):
        raise ValueError(f"checkpoint array {p.shape} does not contain an interior of {tuple(shape3)}")
    sl = tuple(slice(o // 2, o // 2 + s) for o, s in zip(off, shape3))
    return np.ascontiguousarray(p[sl])


# ---------------------------------------------------------------------------------------
# schedules (Oceananigans.Utils.TimeInterval / IterationInterval)
# ---------------------------------------------------------------------------------------
@dataclass
class TimeInterval:
    interval: float
    first_actuation_time: float = 0.0
    actuations: int = 0

    def next_actuation_time(self) -> float:
        return self.first_actuation_time + (self.actuations + 1) * self.interval

    def __call__(self, model) -> bool:
        if model.clock["time"] >= self.next_actuation_time():
            self.actuations += 1
            return True
        return False

    def initialize(self, model) -> None:
        """Picking up at t > 0: past actuations are not replayed."""
        t = model.clock["time"]
        self.actuations = max(self.actuations, int((t - self.first_actuation_time) // self.interval))


@dataclass
class IterationInterval:
    interval: int
    offset: int = 0

    def __call__(self, model) -> bool:
        return (model.clock["iteration"] - self.offset) % self.interval == 0

    def initialize(self, model) -> None:
        pass


# ---------------------------------------------------------------------------------------
# writers
# ---------------------------------------------------------------------------------------
def vertical_vorticity_surface(model) -> np.ndarray:
    """VerticalVorticityField(grid, velocities; indices=(:, :, Nz)) (simulation_outputs.jl:25): ζ at (f, f, c) on the
    top level, (Ny + 1, Nx); the west neighbour column comes from the x-halo (periodic image or neighbouring slab).
    Wall rows (j = 1, Ny + 1) are boundary nodes: zero."""
    g = model.grid
    k = g.Nz - 1
    u = model.get_level("U", k, 1, x_halo_data=True)[1:-1, 1:-1]     # (Ny, Nx)
    v = model.get_level("V", k, 1, x_halo_data=True)[1:-1, :-1]      # (Ny + 1, Nx + 1): column -1 first
    rad = np.pi / 180.0
    dxfc = g.radius * np.cos(g.phi_c() * rad) * (g.dlam * rad)       # Δx at (f, c)
    dy = g.radius * g.dphi * rad
    phif = g.phi_f()
    azff = g.radius ** 2 * (g.dlam * rad) * (np.sin((phif + 0.5 * g.dphi) * rad) - np.sin((phif - 0.5 * g.dphi) * rad))
    zeta = np.zeros((g.Ny + 1, g.Nx))
    dxv = dy * (v[:, 1:] - v[:, :-1])                                  # δx(Δy v)
    dyu = dxfc[1:, None] * u[1:, :] - dxfc[:-1, None] * u[:-1, :]      # δy(Δx u) at rows 1..Ny-1
    zeta[1:-1, :] = (dxv[1:-1, :] - dyu) / azff[1:-1, None]
    return zeta


@dataclass
class SurfaceWriter:
    """JLD2OutputWriter(model, outputs; schedule=TimeInterval(surface_time), filename="<Experiment>_fields_<rank>",
    with_halos=true) -- simulation_outputs.jl:29-34.  Groups `timeseries/<name>/<iteration>`, `timeseries/t/<iteration>`."""
    path: str
    schedule: TimeInterval
    overwrite_existing: bool = True
    _started: bool = False

    def write(self, model) -> None:
        if not self._started:
            if os.path.exists(self.path) and self.overwrite_existing:
                os.remove(self.path)
            self._started = True
        it = model.clock["iteration"]
        top = model.grid.Nz - 1                     # indices = (:, :, Nz): also for w, whose top face is Nz + 1
        out = {}
        for name, fid in (("u", "U"), ("v", "V"), ("w", "W"), ("T", "T"), ("S", "S")):
            out[f"timeseries/{name}/{it}"] = model.get_level(fid, top, HALO, x_halo_data=True)[None]
        out[f"timeseries/η/{it}"] = _parent(model.get("ETA"))
        out[f"timeseries/ζ/{it}"] = _parent(vertical_vorticity_surface(model))
        out[f"timeseries/t/{it}"] = np.float64(model.clock["time"])
        try:
            GroupFile(self.path).write(out)
        except ValueError:
            pass    # this iteration is already in the file (initialisation after a pick-up)


@dataclass
class Checkpointer:
    """Checkpointer(model; schedule=TimeInterval(checkpoint_time), prefix="<Experiment>_checkpoint_<rank>")
    -- simulation_outputs.jl:36-40.  One file per actuation: <prefix>_iteration<N>.npz."""
    dir: str
    prefix: str
    schedule: TimeInterval
    overwrite_existing: bool = True
    extended: bool = True

    def filepath(self, iteration: int) -> str:
        return os.path.join(self.dir, f"{self.prefix}_iteration{iteration}.npz")

    def write(self, model) -> str:
        clk = model.clock
        path = self.filepath(clk["iteration"])
        if os.path.exists(path):
            if not self.overwrite_existing:
                raise FileExistsError(path)
            os.remove(path)
        out = {}
        for name, fid in _PROGNOSTIC:
            out[f"{name}/data"] = _parent(model.get(fid))
            gn, gm = _TENDENCY[name]
            out[f"timestepper/Gⁿ/{name}/data"] = _parent(model.get(gn))
            out[f"timestepper/G⁻/{name}/data"] = _parent(model.get(gm))
        out["η/data"] = _parent(model.get("ETA"))
        out["clock/time"] = np.float64(clk["time"])
        out["clock/iteration"] = np.int64(clk["iteration"])
        if self.extended:
            for fid in _EXTRA:       # their x-halo columns are state too (not recomputed by update_state!)
                out[f"b200/{fid}/data"] = model.get_parent(fid, HALO, x_halo_data=True)
            out["b200/last_dt"] = np.float64(clk["last_dt"])
        GroupFile(path, "w").write(out)
        return path


def set_outputs(simulation, experiment: str, *, overwrite_existing: bool = True, surface_time: float = 1 * days,
                checkpoint_time: float = 5 * days, output_dir: str = "./") -> None:
    """set_outputs!(simulation, Val(Experiment); ...) -- simulation_outputs.jl:9-43 (Quiescent: nothing, :43)."""
    if experiment == "Quiescent":
        return
    rank = simulation.model.grid.rank
    simulation.output_writers["surface"] = SurfaceWriter(
        os.path.join(output_dir, f"{experiment}_fields_{rank}.npz"), TimeInterval(surface_time), overwrite_existing)
    simulation.output_writers["checkpointer"] = Checkpointer(
        output_dir, f"{experiment}_checkpoint_{rank}", TimeInterval(checkpoint_time), overwrite_existing)


# ---------------------------------------------------------------------------------------
# restore: set!(model, filepath) -- simulation_outputs.jl:45-102
# ---------------------------------------------------------------------------------------
def load_free_surface(model, file) -> None:
    """load_free_surface! (:72-81): the stored η may carry any horizontal halo width (distributed runs widen it)."""
    shape = _abi.field_shape("ETA", *model.grid.size())
    model.set("ETA", _interior(file["η/data"], shape))


def set_time_stepper_tendencies(model, file) -> bool:
    """set_time_stepper_tendencies! (:83-102): G^n and G^- of u, v, T, S when present."""
    found = True
    for name, _ in _PROGNOSTIC:
        key = f"timestepper/Gⁿ/{name}/data"
        if key not in file:
            print(f"[ Warning: Tendencies for {name} do not exist in checkpoint and could not be restored.", flush=True)
            found = False
            continue
        gn, gm = _TENDENCY[name]
        model.set(gn, _interior(file[key], _abi.field_shape(gn, *model.grid.size())))
        model.set(gm, _interior(file[f"timestepper/G⁻/{name}/data"], _abi.field_shape(gm, *model.grid.size())))
    return found


def set_from_checkpoint(model, filepath: str) -> None:
    """set!(model, filepath::AbstractString) (:45-70): prognostic fields (interior only, so the halo width may differ),
    η, the time-stepper tendencies and the clock.  The grid is NOT validated, like the reference."""
    with np.load(filepath) as file:
        for name, fid in _PROGNOSTIC:
            if f"{name}/data" in file:            # "Test if variable exist in checkpoint"
                model.set(fid, _interior(file[f"{name}/data"], _abi.field_shape(fid, *model.grid.size())))
        load_free_surface(model, file)
        # the clock FIRST: update_state! evaluates the time-interpolated surface fluxes (forcing of the tendencies, Q_b of the
        # Ri-based entrainment term) at the model time -- with the clock of a fresh model a checkpoint that carries no
        # b200/* members (one converted from the reference) would restore a diffusivity filtered with the wrong Q_b
        last_dt = float(file["b200/last_dt"]) if "b200/last_dt" in file else 0.0
        model.backend.set_clock(float(file["clock/time"]), int(file["clock/iteration"]), last_dt)
        # auxiliaries (w, pHY', diffusivities) and halos of the restored state; then the stored tendencies and the
        # state Oceananigans does not checkpoint overwrite what update_state! recomputed
        model.backend.update_state()
        set_time_stepper_tendencies(model, file)
        if "b200/last_dt" in file:
            for fid in _EXTRA:
                model.set_parent(fid, file[f"b200/{fid}/data"], x_halo_data=True)
