"""Shared problem definitions for tests, smoke() and bench.py (closed-form, no RNG).

A case = grid size + experiment + time step + a deterministic initial state that exercises
every term of the step: the reference's DoubleDrake initial condition
(boundary_and_initial_conditions.jl:9-18) plus smooth closed-form perturbations of u, v, eta
and of T, S (so that WENO weights, Coriolis, pressure gradient and the barotropic solver all
see non-trivial inputs).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np


@dataclass
class Case:
    name: str
    Nx: int
    Ny: int
    Nz: int
    experiment: str = "DoubleDrake"
    dt: float = 1200.0
    perturb: bool = True
    depth: float = 3000.0
    qgleith: bool = False
    bathymetry: str = "experiment"    # "experiment" | "bumpy" (k-dependent masks)
    ranks: int = 1
    balance: str = "equal"            # "equal" (reference: Partition(ranks...)) | "cost" (kernel-cost-balanced slabs)
    # a periodic SECTOR of a larger grid: Nx columns starting at `sector_start` of a grid with `sector_of` columns (same
    # spacing, levels, bathymetry rule and forcing as the full grid) -- BASELINE shapes at a size the oracle finishes
    sector_of: int = 0
    sector_start: int = 0
    uniform_TS: tuple = ()            # (T, S): a fixed T, S everywhere (TEOS-10 N^2 = 0 check)

    def grid(self, pkg, rank=0, ranks=None):
        ranks = ranks or self.ranks
        zf = pkg.exponential_z_faces(self.Nz, self.depth)
        if self.sector_of:
            P = self.sector_of
            g0 = pkg.LatitudeLongitudeGrid(P, self.Ny, self.Nz, zf)
            bottom = pkg.double_drake_bathymetry(g0.lam_c()[None, :], g0.phi_c()[:, None])
            bottom = np.ascontiguousarray(bottom[:, self.sector_start:self.sector_start + self.Nx])
            lon0 = -180.0 + 360.0 * self.sector_start / P
            return pkg.LatitudeLongitudeGrid(self.Nx, self.Ny, self.Nz, zf, longitude=(lon0, lon0 + 360.0 * self.Nx / P),
                                             rank=rank, partition=pkg.Partition.equal(self.Nx, ranks), bottom_height_global=bottom)
        g0 = pkg.LatitudeLongitudeGrid(self.Nx, self.Ny, self.Nz, zf)
        lam, phi = g0.lam_c()[None, :], g0.phi_c()[:, None]
        if self.experiment == "Quiescent" and self.bathymetry == "experiment":
            bottom = None
        elif self.experiment == "RealisticOcean":
            bottom = pkg.synthetic.realistic_bathymetry(self.Nx, self.Ny, self.depth)
        elif self.bathymetry == "bumpy":
            # smooth seamounts + a continent: exercises k-dependent immersed masks
            bottom = -self.depth * (1.0 - 0.85 * np.exp(-((lam - 40.0) / 25.0) ** 2 - ((phi - 10.0) / 18.0) ** 2)
                                    - 0.6 * np.exp(-((lam + 100.0) / 30.0) ** 2 - ((phi + 30.0) / 15.0) ** 2))
            bottom = np.where((np.abs(lam + 20.0) < 12.0) & (np.abs(phi - 35.0) < 14.0), 100.0, bottom)
            bottom = bottom + 0 * lam
        else:
            bottom = pkg.double_drake_bathymetry(lam, phi)
        if self.balance == "cost" and ranks > 1:
            part = pkg.Partition.balanced(pkg.kernel_cost_per_column(bottom, self.Nx, self.Ny), ranks, min_columns=40)
        elif self.balance == "wet" and ranks > 1 and bottom is not None:
            # LOADBALANCE=1: the reference's wet-cell prefix split + memory clamp (grid_load_balance.jl:49-62,97-146)
            load = pkg.assess_x_load_from_bottom(g0.z_c(), bottom)
            local_Nx = pkg.calculate_local_N(load, self.Nx, ranks)
            pkg.redistribute_size_to_fulfill_memory_limitation(local_Nx, 1150)
            part = pkg.Partition(tuple(int(n) for n in local_Nx))
        else:
            part = pkg.Partition.equal(self.Nx, ranks)
        return pkg.LatitudeLongitudeGrid(self.Nx, self.Ny, self.Nz, zf, rank=rank, partition=part,
                                         bottom_height_global=bottom)

    def make_model(self, pkg, backend_factory, rank=0, ranks=None, device=0, initialize=True, comm=None, precision="Float64"):
        grid = self.grid(pkg, rank, ranks)
        res = (self.sector_of or self.Nx) / 360.0
        max_dt = 3240.0 / res
        # the substep count follows the FULL grid's spacing (near_global_simulation.jl:11-19 reduces over all ranks)
        g_sub = pkg.LatitudeLongitudeGrid(self.sector_of, self.Ny, self.Nz, grid.z_faces) if self.sector_of else grid
        nsub = pkg.barotropic_substeps(max_dt, g_sub)
        fs = pkg.SplitExplicitFreeSurface(substeps=nsub)
        if self.experiment == "RealisticOcean":
            bcs = pkg.set_boundary_conditions("RealisticOcean", grid, forcing_arrays=pkg.synthetic.forcing_arrays(grid))
        else:
            bcs = pkg.set_boundary_conditions("DoubleDrake" if self.experiment == "DoubleDrake" else "Quiescent", grid)
        closure = (pkg.VerticalScalarDiffusivity(), pkg.RiBasedVerticalDiffusivity())
        if self.qgleith:
            closure = closure + (pkg.QGLeith(),)
        eos = "linear" if self.experiment == "DoubleDrake" else "teos10"
        model = pkg.HydrostaticFreeSurfaceModel(grid=grid, free_surface=fs, eos=eos, closure=closure,
                                                boundary_conditions=bcs, device=device,
                                                backend_factory=backend_factory, comm=comm, precision=precision)
        if initialize:
            self.initialize(pkg, model)
        return model

    def initialize(self, pkg, model):
        if self.uniform_TS:
            pkg.set_model(model, T=self.uniform_TS[0], S=self.uniform_TS[1])
            return
        if self.experiment == "RealisticOcean":
            T0, S0 = pkg.synthetic.initial_T_S(model.grid)
            r = math.pi / 180.0
            pkg.set_model(model, T=T0, S=S0,
                          u=lambda lam, phi, z: 0.05 * np.sin(2 * lam * r) * np.cos(phi * r) * np.exp(z / 2000.0),
                          v=lambda lam, phi, z: 0.03 * np.cos(3 * lam * r) * np.sin(2 * phi * r) * np.exp(z / 1000.0))
            model.backend.set_clock(1.7 * 86400.0, 0, float("inf"))   # mid-interval: exercises the time interpolation
            return
        if self.experiment == "DoubleDrake":
            pkg.initialize_model(model, "DoubleDrake")
        if not self.perturb:
            return
        D = self.depth
        r = math.pi / 180.0

        def Tf(lam, phi, z):
            base = pkg.exponential_profile(z, Lz=D, h=D / 15.0) * np.maximum(pkg.T_reference(phi), 2.0)
            return base + 0.5 * np.sin(3 * lam * r) * np.cos(2 * phi * r) * np.exp(z / 500.0)

        def Sf(lam, phi, z):
            return 35.0 + 0.2 * np.cos(2 * lam * r + 0.3) * np.sin(3 * phi * r) * np.exp(z / 800.0)

        def uf(lam, phi, z):
            return 0.1 * np.sin(2 * lam * r) * np.cos(phi * r) * (1.0 + z / D) + 0 * z

        def vf(lam, phi, z):
            return 0.05 * np.cos(3 * lam * r) * np.sin(2 * phi * r) * np.exp(z / 1000.0)

        def ef(lam, phi):
            return 0.1 * np.cos(phi * r) * np.sin(2 * lam * r)

        pkg.set_model(model, T=Tf, S=Sf, u=uf, v=vf, eta=ef)


CASES = {
    # the reference's own test size (test/runtests.jl:11-13: resolution 1/5 -> 72 x 30, Nz = 10)
    "ref_test": Case("ref_test", 72, 30, 10, dt=600.0),
    "dd_small": Case("dd_small", 90, 40, 8, dt=1200.0),
    "dd_mid": Case("dd_mid", 180, 75, 12, dt=1200.0),
    # 2 x-slabs of 30 columns: narrower than the wide barotropic halo (M + 1 = 38) -> per-substep one-column exchanges
    "dd_narrow": Case("dd_narrow", 60, 40, 8, dt=1200.0),
    "dd_leith": Case("dd_leith", 90, 40, 8, dt=1200.0, qgleith=True),
    "bumpy": Case("bumpy", 96, 48, 12, dt=900.0, bathymetry="bumpy"),
    # RealisticOcean-shaped (BASELINE configs[3]): TEOS-10, synthetic continents, time-interpolated fluxes +
    # restoring, quadratic bottom drag
    "realistic": Case("realistic", 96, 48, 12, experiment="RealisticOcean", dt=600.0, depth=5244.5),
    "quiescent": Case("quiescent", 90, 40, 8, experiment="Quiescent", dt=3240.0, perturb=False),
    # C1: BASELINE.json configs[0]
    "c1": Case("c1", 360, 150, 10, experiment="Quiescent", dt=3240.0, perturb=False),
    # ---- BASELINE shapes at sizes the oracle finishes in seconds (parity only on toy grids) ----
    # configs[2] DoubleDrake 1/6 deg NZ = 120: a 37-column periodic sector around the ridge at 0..1 deg E (the sample
    # bench.py's CPU legs time): full Ny = 900 (57 tile rows), 15 level chunks of the tendency kernels, 1/6 deg metrics
    "c3_sector": Case("c3_sector", 37, 900, 120, dt=540.0, sector_of=2160, sector_start=1072),
    # the same grid at the width of ONE x-slab of the 8-GPU partition (270 columns: 17 tile columns, 32 x 4 column blocks)
    "c3_slab270": Case("c3_slab270", 270, 900, 120, dt=540.0, sector_of=2160, sector_start=1000),
    # configs[3] RealisticOcean-shaped 1/4 deg per GPU, NZ = 100: a 1440-wide TEOS-10 strip with synthetic continents,
    # time-interpolated fluxes + restoring, quadratic + immersed drag
    "c4_strip": Case("c4_strip", 1440, 48, 100, experiment="RealisticOcean", dt=810.0, depth=5244.5),
    # TEOS-10 with a fixed T, S over depth: N^2 must be exactly 0 (no compressibility term in d b / dz)
    "teos_uniform": Case("teos_uniform", 96, 48, 12, experiment="Quiescent", dt=900.0, bathymetry="bumpy", uniform_TS=(10.0, 35.0)),
}


def build_case(name: str) -> Case:
    return CASES[name]


def signature(fields: dict) -> dict:
    """Compact fingerprint of a set of fields: bottom / middle / top horizontal slices plus global statistics."""
    out = {}
    for n, a in fields.items():
        if a.ndim == 3:
            ks = sorted(set([0, a.shape[0] // 2, a.shape[0] - 1]))
            out[n + "_slices"] = np.ascontiguousarray(a[ks])
        else:
            out[n + "_slices"] = np.ascontiguousarray(a)
        out[n + "_stats"] = np.array([np.sum(a), np.sum(a * a), np.max(np.abs(a))])
    return out


def check_signature(fields: dict, gold, tol: float) -> list:
    """Compare fields against a stored signature; returns a list of mismatch descriptions."""
    bad = []
    sig = signature(fields)
    for key, val in sig.items():
        ref = gold[key]
        scale = max(1.0, float(np.max(np.abs(ref))))
        err = float(np.max(np.abs(val - ref)))
        if key.endswith("_stats"):
            # sums of O(1e5) terms: compare relative to the sum of absolute values' scale
            scale = max(scale, float(np.abs(ref[1])) ** 0.5 * 10)
        if not err <= tol * scale:
            bad.append(f"{key}: err {err:.3e} (scale {scale:.3e})")
    return bad
