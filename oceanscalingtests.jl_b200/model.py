"""Host-side mirror of the model objects the reference builds in
/root/reference/src/near_global_simulation.jl:66-113 and of Oceananigans'
`time_step!` entry point, backed by libocean_b200.so through the C ABI.

Names follow the reference: SplitExplicitFreeSurface, RiBasedVerticalDiffusivity,
VerticalScalarDiffusivity, QGLeith, HydrostaticFreeSurfaceModel, time_step, set_model.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _abi
from .grid_load_balance import LatitudeLongitudeGrid

g_Earth = 9.80665        # Oceananigans.BuoyancyModels.g_Earth
Omega_Earth = 7.292115e-5  # Oceananigans.Coriolis.Ω_Earth


# ---------------------------------------------------------------------------------------
# split-explicit free surface settings (near_global_simulation.jl:11-21,82)
# ---------------------------------------------------------------------------------------
def barotropic_substeps(dt: float, grid: LatitudeLongitudeGrid, *, g: float = g_Earth, CFL: float = 0.75,
                        allreduce_min=None) -> int:
    """near_global_simulation.jl:11-19.  `allreduce_min` replaces MPI.Allreduce(local_Δ, min, comm)."""
    wave_speed = math.sqrt(g * grid.Lz)
    local_delta = 1 / math.sqrt(1 / grid.minimum_xspacing() ** 2 + 1 / grid.minimum_yspacing() ** 2)
    global_delta = allreduce_min(local_delta) if allreduce_min is not None else local_delta
    return max(int(math.ceil(2 * dt / (CFL / wave_speed * global_delta))), 10)


def averaging_shape_function(tau: float, *, p: int = 2, q: int = 4, r: float = 0.18927) -> float:
    tau0 = (p + 2) * (p + q + 2) / (p + 1) / (p + q + 1)
    return (tau / tau0) ** p * (1 - (tau / tau0) ** q) - r * (tau / tau0)


def _searchsortedlast_rev(v, x) -> int:
    """Julia `searchsortedlast(v, x, rev=true)` (binary search; returns a 1-based index)."""
    lo, hi = 0, len(v) + 1
    while lo < hi - 1:
        m = lo + ((hi - lo) >> 1)
        if v[m - 1] < x:      # lt(Reverse, x, v[m]) == isless(v[m], x)
            hi = m
        else:
            lo = m
    return lo


def weights_from_substeps(substeps: int):
    """Oceananigans `weights_from_substeps` [OCN-recall]: tau_f = range(0, 2, substeps+1);
    weights truncated at the last positive one and normalised.  Returns (fractional dtau, weights)."""
    tau_f = np.linspace(0.0, 2.0, substeps + 1)
    dtau = float(tau_f[1] - tau_f[0])
    w = np.array([averaging_shape_function(t) for t in tau_f[1:]])
    idx = _searchsortedlast_rev(w, 0.0)
    w = w[:idx]
    w = w / np.sum(w)
    return dtau, np.ascontiguousarray(w)


@dataclass
class SplitExplicitFreeSurface:
    """SplitExplicitFreeSurface(precision; substeps) -- near_global_simulation.jl:82"""
    substeps: int
    gravitational_acceleration: float = g_Earth
    fractional_step_size: float = field(init=False)
    averaging_weights: np.ndarray = field(init=False)

    def __post_init__(self):
        self.fractional_step_size, self.averaging_weights = weights_from_substeps(self.substeps)


def substeps(free_surface: SplitExplicitFreeSurface) -> int:
    """near_global_simulation.jl:21"""
    return len(free_surface.averaging_weights)


@dataclass
class VerticalScalarDiffusivity:
    """VerticalScalarDiffusivity(VerticallyImplicitTimeDiscretization(); nu, kappa) -- :73"""
    nu: float = 5e-4
    kappa: float = 3e-5


@dataclass
class RiBasedVerticalDiffusivity:
    """Oceananigans defaults [OCN-recall], near_global_simulation.jl:49"""
    nu0: float = 0.7
    kappa0: float = 0.5
    kappa_ca: float = 1.7
    Cen: float = 0.1
    Cav: float = 0.6
    Ri0: float = 0.1
    Ridelta: float = 0.4


@dataclass
class QGLeith:
    """quasi_geostrophic_leith.jl:37-48 (optional; not in the default closure tuple)"""
    C: float = 2.0
    min_N2: float = 1e-20
    Vscale: float = 1.0


@dataclass
class BoundaryConditions:
    """What set_boundary_conditions (boundary_and_initial_conditions.jl:48-137) produces."""
    kind: int = _abi.OB200_BC_QUIESCENT
    wind_coeffs: tuple = ()
    salt_coeffs: tuple = ()
    T_restoring_lambda: float = 0.0
    drag_mu: float = 0.0
    forcing_arrays: dict = field(default_factory=dict)  # RealisticOcean: taux,tauy,Qs,Fs,Tr,Sr
    providers: dict = field(default_factory=dict)       # "fluxes" / "restoring": (grid, filenum) -> arrays (data/*.jld2)


def build_config(grid: LatitudeLongitudeGrid, free_surface: SplitExplicitFreeSurface, *, eos: str,
                 closure, boundary_conditions: BoundaryConditions, advection_orders=(9, 5, 7), chi: float = 0.1,
                 rotation_rate: float = Omega_Earth, barotropic_exchange: str = "auto"):
    """Assemble `ob200_config` (include/ocean_b200.h).  Returns (config, keepalive arrays)."""
    cfg = _abi.ob200_config()
    cfg.struct_bytes = C.sizeof(_abi.ob200_config)
    cfg.Nx, cfg.Ny, cfg.Nz = grid.Nx, grid.Ny, grid.Nz
    cfg.Nx_global, cfg.i_offset = grid.Nx_global, grid.i_offset
    cfg.rank, cfg.nranks = grid.rank, grid.nranks
    cfg.lon_west, cfg.lon_east = grid.longitude
    cfg.lat_south, cfg.lat_north = grid.latitude
    cfg.radius, cfg.gravity, cfg.rotation_rate = grid.radius, free_surface.gravitational_acceleration, rotation_rate
    cfg.eos_kind = _abi.OB200_EOS_LINEAR if eos == "linear" else _abi.OB200_EOS_TEOS10
    cfg.eos_alpha, cfg.eos_beta, cfg.eos_rho0 = 1.67e-4, 7.80e-4, 1020.0
    ri = None
    leith = None
    vsd = VerticalScalarDiffusivity()
    for c in closure:
        if isinstance(c, VerticalScalarDiffusivity):
            vsd = c
        elif isinstance(c, RiBasedVerticalDiffusivity):
            ri = c
        elif isinstance(c, QGLeith):
            leith = c
        else:
            raise TypeError(f"unsupported closure {type(c).__name__}")
    cfg.nu_z, cfg.kappa_z = vsd.nu, vsd.kappa
    cfg.ri_enabled = int(ri is not None)
    if ri is not None:
        cfg.ri_nu0, cfg.ri_kappa0, cfg.ri_kappa_ca, cfg.ri_Cen = ri.nu0, ri.kappa0, ri.kappa_ca, ri.Cen
        cfg.ri_Cav, cfg.ri_Ri0, cfg.ri_Ridelta = ri.Cav, ri.Ri0, ri.Ridelta
    cfg.order_vorticity, cfg.order_divergence, cfg.order_tracer = advection_orders
    cfg.ab2_chi = chi
    w = np.ascontiguousarray(free_surface.averaging_weights, dtype=np.float64)
    cfg.n_substeps, cfg.frac_dtau = len(w), free_surface.fractional_step_size
    bc = boundary_conditions
    cfg.bc_kind = bc.kind
    wc = np.zeros(24)
    sc = np.zeros(16)
    if bc.wind_coeffs:
        wc[:] = np.asarray(bc.wind_coeffs, dtype=np.float64).reshape(-1)
    if bc.salt_coeffs:
        sc[:] = np.asarray(bc.salt_coeffs, dtype=np.float64).reshape(-1)
    cfg.wind_coeffs = (C.c_double * 24)(*wc)
    cfg.salt_coeffs = (C.c_double * 16)(*sc)
    cfg.T_restoring_lambda, cfg.drag_mu = bc.T_restoring_lambda, bc.drag_mu
    cfg.qgleith_enabled = int(leith is not None)
    if leith is not None:
        cfg.leith_C, cfg.leith_min_N2, cfg.leith_Vscale = leith.C, leith.min_N2, leith.Vscale
    zf = np.ascontiguousarray(grid.z_faces, dtype=np.float64)
    # x-slabs carry a barotropic halo of M + 1 columns, exchanged once per step
    halo_b = grid.halo if grid.nranks == 1 else max(grid.halo, len(w) + 1)
    cfg.bottom_halo_columns = halo_b
    # x-slabs: wide halo exchanged once per step, or one-column exchanges every half substep (narrow slabs: the reference
    # clamps wet-cell-balanced slabs at >= 20 columns, grid_load_balance.jl:132-145, fewer than the M + 1 a wide halo needs)
    cfg.barotropic_exchange = {"auto": _abi.OB200_BAROX_AUTO, "wide": _abi.OB200_BAROX_WIDE,
                               "substep": _abi.OB200_BAROX_SUBSTEP}[barotropic_exchange]
    bh = np.ascontiguousarray(grid.bottom_height_local(halo_b), dtype=np.float64)
    cfg.z_faces = _abi.as_double_ptr(zf)
    cfg.bottom_height = _abi.as_double_ptr(bh)
    cfg.averaging_weights = _abi.as_double_ptr(w)
    return cfg, (zf, bh, w)


# ---------------------------------------------------------------------------------------
# backends: thin objects that own a native handle.  The product has exactly one: B200.
# ---------------------------------------------------------------------------------------
class B200Backend:
    """libocean_b200.so through ctypes (the Julia glue does the same with ccall)."""
    name = "b200"

    def __init__(self, cfg, device: int = 0, precision: str = "Float64"):
        self.lib = _abi.load_library(precision=precision)
        self.precision = precision
        self.dtype = np.float64 if self.lib.ob200_real_bytes() == 8 else np.float32   # element type of every field buffer
        self.h = C.c_void_p()
        rc = self.lib.ob200_create(C.byref(cfg), device, C.byref(self.h))
        if rc != 0:
            msg = self.lib.ob200_last_error(None)
            raise RuntimeError(f"ob200_create failed ({rc}): {msg.decode() if msg else ''}")

    def _check(self, rc):
        if rc != 0:
            msg = self.lib.ob200_last_error(self.h)
            raise RuntimeError(f"libocean_b200 error {rc}: {msg.decode() if msg else ''}")

    def set_field(self, fid, arr: np.ndarray, halo=(0, 0, 0)):
        """`arr`: interior-shaped, or parent-shaped with `halo` = (hx, hy, hz) cells around the interior."""
        arr = np.ascontiguousarray(arr, dtype=self.dtype)
        self._check(self.lib.ob200_set_field(self.h, fid, arr.ctypes.data, 0, *halo))

    def get_field(self, fid, shape, halo=(0, 0, 0)) -> np.ndarray:
        out = np.zeros(shape, dtype=self.dtype) if any(halo) else np.empty(shape, dtype=self.dtype)   # halo cells are not written
        self._check(self.lib.ob200_get_field(self.h, fid, out.ctypes.data, 0, *halo))
        return out

    def get_level(self, fid, k: int, shape2d, halo=(0, 0)) -> np.ndarray:
        """One horizontal level of a 3-D field, (ny + 2 hy, Nx + 2 hx)."""
        out = np.zeros(shape2d, dtype=self.dtype) if any(halo) else np.empty(shape2d, dtype=self.dtype)
        self._check(self.lib.ob200_get_level(self.h, fid, int(k), out.ctypes.data, 0, *halo))
        return out

    def set_field_device(self, fid, ptr: int, halo=(0, 0, 0)):
        self._check(self.lib.ob200_set_field(self.h, fid, ptr, 1, *halo))

    def get_field_device(self, fid, ptr: int, halo=(0, 0, 0)):
        self._check(self.lib.ob200_get_field(self.h, fid, ptr, 1, *halo))

    def bind(self, fid, ptr: int, halo=(0, 0, 0)):
        """Register a caller-owned device parent array (what julia/OceanB200.jl does with `pointer(parent(field))`)."""
        self._check(self.lib.ob200_bind(self.h, fid, ptr, *halo))

    def unbind(self, fid):
        self._check(self.lib.ob200_unbind(self.h, fid))

    def sync_outputs(self):
        self._check(self.lib.ob200_sync_outputs(self.h))

    def sync_inputs(self):
        self._check(self.lib.ob200_sync_inputs(self.h))

    def set_forcing(self, which, arr):
        arr = np.ascontiguousarray(arr, dtype=self.dtype)
        self._check(self.lib.ob200_set_forcing(self.h, which, arr.ctypes.data, 0))

    def update_state(self):
        self._check(self.lib.ob200_update_state(self.h))

    def time_step(self, dt):
        self._check(self.lib.ob200_time_step(self.h, float(dt)))

    def time_steps(self, dt, n):
        self._check(self.lib.ob200_time_steps(self.h, float(dt), int(n)))

    def run_stage(self, stage, dt=0.0, arg=0):
        self._check(self.lib.ob200_run_stage(self.h, stage, float(dt), int(arg)))

    def sync(self):
        self._check(self.lib.ob200_sync(self.h))

    def clock(self):
        t, it, ldt = C.c_double(), C.c_int64(), C.c_double()
        self._check(self.lib.ob200_get_clock(self.h, C.byref(t), C.byref(it), C.byref(ldt)))
        return t.value, it.value, ldt.value

    def set_clock(self, time, iteration, last_dt):
        self._check(self.lib.ob200_set_clock(self.h, float(time), int(iteration), float(last_dt)))

    def diagnostics(self):
        m = (C.c_double * 6)()
        cfl = C.c_double()
        self._check(self.lib.ob200_diagnostics(self.h, m, C.byref(cfl)))
        return list(m), cfl.value

    def timing(self):
        t = _abi.ob200_timing()
        self._check(self.lib.ob200_get_timing(self.h, C.byref(t)))
        return [t.ms[i] for i in range(t.n)], int(t.launches)

    def set_option(self, name: str, value: int):
        self._check(self.lib.ob200_set_option(self.h, name.encode(), int(value)))

    def comm_init(self, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        self._check(self.lib.ob200_comm_init(self.h, buf))

    @staticmethod
    def comm_unique_id() -> bytes:
        lib = _abi.load_library()
        buf = C.create_string_buffer(128)
        rc = lib.ob200_comm_unique_id(buf)
        if rc != 0:
            raise RuntimeError("ob200_comm_unique_id failed")
        return buf.raw

    def close(self):
        if self.h:
            self.lib.ob200_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HydrostaticFreeSurfaceModel:
    """HydrostaticFreeSurfaceModel(; grid, free_surface, momentum_advection, tracer_advection, coriolis,
    buoyancy, tracers, boundary_conditions, closure) -- near_global_simulation.jl:105-113.

    All device state lives inside the native handle; `model.fields[name]` copies a field out.
    Like the Oceananigans constructor, construction ends with one `update_state!`.
    """

    def __init__(self, *, grid: LatitudeLongitudeGrid, free_surface: SplitExplicitFreeSurface, eos: str = "linear",
                 closure=(VerticalScalarDiffusivity(), RiBasedVerticalDiffusivity()),
                 boundary_conditions: Optional[BoundaryConditions] = None, advection_orders=(9, 5, 7),
                 device: int = 0, backend_factory=None, comm=None, barotropic_exchange: str = "auto",
                 precision: str = "Float64"):
        self.grid = grid
        self.free_surface = free_surface
        self.closure = tuple(closure)
        self.boundary_conditions = boundary_conditions or BoundaryConditions()
        self.eos = eos
        self.cfg, self._keep = build_config(grid, free_surface, eos=eos, closure=self.closure,
                                            boundary_conditions=self.boundary_conditions,
                                            advection_orders=advection_orders, barotropic_exchange=barotropic_exchange)
        factory = backend_factory or B200Backend
        # `precision`: the reference's PRECISION (experiments/run.jl:38, near_global_simulation.jl:48): which build of the library
        self.precision = precision
        self.backend = factory(self.cfg, device, precision) if factory is B200Backend else factory(self.cfg)
        if grid.nranks > 1 and hasattr(self.backend, "comm_init"):
            # NCCL ring bootstrap (replaces MPI.Init + Distributed(...), experiments/run.jl:9-27): rank 0 creates the
            # unique id, `comm.bcast_bytes` (torch.distributed here, MPI.Bcast in Julia) hands it to every rank
            if comm is None:
                raise ValueError("x-slab models need `comm` (an object with bcast_bytes) to bootstrap NCCL")
            uid = self.backend.comm_unique_id() if grid.rank == 0 else None
            self.backend.comm_init(comm.bcast_bytes(uid))
        for i, key in enumerate(("taux", "tauy", "Qs", "Fs", "Tr", "Sr")):
            if key in self.boundary_conditions.forcing_arrays:
                self.backend.set_forcing(i, self.boundary_conditions.forcing_arrays[key])
        self.backend.update_state()

    # --- field access (interior(field) as numpy, (k, j, i) order) ---------------------
    def get(self, name: str) -> np.ndarray:
        shape = _abi.field_shape(name, *self.grid.size())
        return self.backend.get_field(_abi.FIELD_ID[name], shape)

    def set(self, name: str, arr) -> None:
        shape = _abi.field_shape(name, *self.grid.size())
        a = np.broadcast_to(np.asarray(arr, dtype=np.float64), shape)
        self.backend.set_field(_abi.FIELD_ID[name], np.ascontiguousarray(a))

    def parent_shape(self, name: str, halo: int = 7):
        shape = _abi.field_shape(name, *self.grid.size())
        if len(shape) == 2:
            return (1, shape[0] + 2 * halo, shape[1] + 2 * halo)
        return tuple(n + 2 * halo for n in shape)

    def get_parent(self, name: str, halo: int = 7, x_halo_data: bool = False) -> np.ndarray:
        """Oceananigans parent array (k, j, i) with `halo` cells around the interior.  Halo cells are zero unless
        `x_halo_data`, which also copies the x-halo columns the device holds (option transfer_x_halo)."""
        two_d = len(_abi.field_shape(name, *self.grid.size())) == 2
        if x_halo_data:
            self.backend.set_option("transfer_x_halo", halo)
        try:
            return self.backend.get_field(_abi.FIELD_ID[name], self.parent_shape(name, halo), (halo, halo, 0 if two_d else halo))
        finally:
            if x_halo_data:
                self.backend.set_option("transfer_x_halo", 0)

    def get_level(self, name: str, k: int, halo: int = 0, x_halo_data: bool = False) -> np.ndarray:
        """Level k (0-based) of a 3-D field as a (ny + 2 halo, Nx + 2 halo) window -- the surface writer's
        `Field(...; indices = (:, :, Nz))` (simulation_outputs.jl:17-25) without moving the whole field."""
        nz, ny, nx = _abi.field_shape(name, *self.grid.size())
        if x_halo_data:
            self.backend.set_option("transfer_x_halo", halo)
        try:
            return self.backend.get_level(_abi.FIELD_ID[name], k, (ny + 2 * halo, nx + 2 * halo), (halo, halo))
        finally:
            if x_halo_data:
                self.backend.set_option("transfer_x_halo", 0)

    def set_parent(self, name: str, parent, x_halo_data: bool = False) -> None:
        """Inverse of get_parent; the halo width is inferred from the array ("allow different halo size")."""
        shape = _abi.field_shape(name, *self.grid.size())
        p = np.asarray(parent, dtype=np.float64)
        if len(shape) == 2 and p.ndim == 2:
            p = p[None]
        core = (1,) + tuple(shape) if len(shape) == 2 else tuple(shape)
        off = [a - b for a, b in zip(p.shape, core)]
        if any(o < 0 or o % 2 for o in off) or off[1] != off[2]:
            raise ValueError(f"{name}: parent array {p.shape} does not wrap an interior of {core}")
        halo = (off[2] // 2, off[1] // 2, off[0] // 2)
        if x_halo_data:
            self.backend.set_option("transfer_x_halo", halo[0])
        try:
            self.backend.set_field(_abi.FIELD_ID[name], p, halo)
        finally:
            if x_halo_data:
                self.backend.set_option("transfer_x_halo", 0)

    @property
    def clock(self):
        t, it, ldt = self.backend.clock()
        return {"time": t, "iteration": it, "last_dt": ldt}

    def prognostic(self) -> dict:
        return {n: self.get(n) for n in ("U", "V", "W", "ETA", "T", "S")}


def set_model(model: HydrostaticFreeSurfaceModel, **fields) -> None:
    """set!(model, T = ..., S = ...) -- boundary_and_initial_conditions.jl:15.
    Values may be arrays of interior shape, scalars, or functions f(lam, phi, z).
    Ends with update_state!, as Oceananigans' set! does."""
    g = model.grid
    names = {"u": "U", "v": "V", "T": "T", "S": "S", "eta": "ETA", "η": "ETA"}
    for key, val in fields.items():
        name = names.get(key, key)
        shape = _abi.field_shape(name, *g.size())
        if callable(val):
            lam = g.lam_c()
            if name == "U":
                lam = lam - 0.5 * g.dlam
            phi = g.phi_f() if name == "V" else g.phi_c()
            if len(shape) == 3:
                val = val(lam[None, None, :], phi[None, :, None], g.z_c()[:, None, None])
            else:
                val = val(lam[None, :], phi[:, None])
        model.set(name, val)
    model.backend.update_state()


def initialize_free_surface(model: HydrostaticFreeSurfaceModel) -> None:
    """Seed the averaged barotropic transports with the vertical integral of the current velocities,
    U-bar, V-bar <- sum_k dz (u, v)  (Oceananigans' `barotropic_mode!`; [OCN-recall, low confidence] later Oceananigans versions
    do this in `initialize_free_surface!` when `run!` initialises the model).

    Every step starts its substeps from the PREVIOUS step's averages (U <- U-bar); a fresh model has U-bar = 0, so
    without this call the barotropic part of an initial velocity field is dropped by the first step's corrector -- harmless for
    the reference's runs, which start from rest (boundary_and_initial_conditions.jl:9-18) or in profile mode, and for this
    repository's own checkpoints, which carry U-bar, V-bar (simulation_outputs.py `b200/*`).  NOT called by set_model or
    time_step: the library and the oracle follow SURVEY A9 literally.  Velocities are zero on masked nodes, so the plain sum is
    the sum over non-peripheral nodes."""
    dz = np.diff(np.asarray(model.grid.z_faces))[:, None, None]
    model.set("UBAR", (model.get("U") * dz).sum(axis=0))
    model.set("VBAR", (model.get("V") * dz).sum(axis=0))


def time_step(model: HydrostaticFreeSurfaceModel, dt: float) -> None:
    """time_step!(model, Δt) -- the hot path (near_global_simulation.jl:181,196)."""
    model.backend.time_step(dt)
