"""Host-side mirror of /root/reference/src/grid_load_balance.jl.

Grid description (LatitudeLongitudeGrid + GridFittedBottom) and the x-slab
partitioning: equal split for Quiescent / DoubleDrake (grid_load_balance.jl:16-33),
wet-cell-balanced slabs for RealisticOcean with LOADBALANCE=1 (:35-75).  The slab
axis is x = longitude, as in the reference (`ranks = (Nranks, 1, 1)`,
experiments/run.jl:22-27); x is periodic so the ranks form a ring.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional, Sequence

import numpy as np

from .bathymetry_and_forcings import double_drake_bathymetry

HALO = 7  # grid_load_balance.jl:25  halo = (7, 7, 7)


@dataclass
class Partition:
    """x-slab partition of the global grid: sizes[r] columns on rank r."""
    sizes: tuple

    @classmethod
    def equal(cls, Nx: int, ranks: int) -> "Partition":
        # Partition(ranks...) -> Nx / Rx columns each; remainder spread over the first ranks
        base, rem = divmod(Nx, ranks)
        return cls(tuple(base + (1 if r < rem else 0) for r in range(ranks)))

    def offset(self, rank: int) -> int:
        return int(sum(self.sizes[:rank]))

    @classmethod
    def balanced(cls, load_per_column, ranks: int, min_columns: int = 20) -> "Partition":
        """x-slabs with (approximately) equal load, by the reference's own greedy prefix rule
        (`calculate_local_N`, grid_load_balance.jl:97-114) applied to an arbitrary per-column load."""
        load = np.asarray(load_per_column, dtype=np.float64)
        if ranks == 1:
            return cls((len(load),))
        target = load.sum() / ranks
        sizes, idx = [], 0
        for r in range(ranks - 1):
            acc, n = 0.0, 0
            while acc <= target and idx < len(load) - (ranks - 1 - r) * min_columns:
                acc += load[idx]; n += 1; idx += 1
            # the reference's rule overshoots by one column on every rank and starves the last one: step back when
            # that is closer to the target
            if n > min_columns and abs(acc - load[idx - 1] - target) < abs(acc - target):
                n -= 1; idx -= 1
            sizes.append(max(n, min_columns))
        sizes.append(len(load) - sum(sizes))
        return cls(tuple(int(x) for x in sizes))


@dataclass
class LatitudeLongitudeGrid:
    """The local (per-rank) view of the near-global grid, plus what is needed to build
    `ob200_config`.  Mirrors LatitudeLongitudeGrid(arch, FT; size, longitude=(-180,180),
    latitude, halo=(7,7,7), z) wrapped in ImmersedBoundaryGrid(GridFittedBottom(...))."""
    Nx_global: int
    Ny: int
    Nz: int
    z_faces: np.ndarray
    latitude: tuple = (-75.0, 75.0)
    longitude: tuple = (-180.0, 180.0)
    rank: int = 0
    partition: Partition = None
    bottom_height_global: Optional[np.ndarray] = None  # (Ny, Nx_global) or None (no bathymetry)
    radius: float = 6371.0e3
    halo: int = HALO

    def __post_init__(self):
        if self.partition is None:
            self.partition = Partition.equal(self.Nx_global, 1)
        self.z_faces = np.ascontiguousarray(self.z_faces, dtype=np.float64)
        assert self.z_faces.shape == (self.Nz + 1,)
        assert sum(self.partition.sizes) == self.Nx_global

    # --- sizes -------------------------------------------------------------
    @property
    def nranks(self) -> int:
        return len(self.partition.sizes)

    @property
    def Nx(self) -> int:
        return int(self.partition.sizes[self.rank])

    @property
    def i_offset(self) -> int:
        return self.partition.offset(self.rank)

    @property
    def Lz(self) -> float:
        return float(self.z_faces[-1] - self.z_faces[0])

    @property
    def Ly(self) -> float:
        return float(self.latitude[1] - self.latitude[0])

    def size(self):
        return (self.Nx, self.Ny, self.Nz)

    # --- nodes -------------------------------------------------------------
    @property
    def dlam(self) -> float:
        return (self.longitude[1] - self.longitude[0]) / self.Nx_global

    @property
    def dphi(self) -> float:
        return (self.latitude[1] - self.latitude[0]) / self.Ny

    def lam_c(self, with_halo: bool = False) -> np.ndarray:
        """Cell-centre longitudes of the local slab (optionally with the 7 halo columns)."""
        h = self.halo if with_halo else 0
        i = np.arange(-h, self.Nx + h) + self.i_offset
        return self.longitude[0] + (i + 0.5) * self.dlam

    def phi_c(self) -> np.ndarray:
        return self.latitude[0] + (np.arange(self.Ny) + 0.5) * self.dphi

    def phi_f(self) -> np.ndarray:
        return self.latitude[0] + np.arange(self.Ny + 1) * self.dphi

    def z_c(self) -> np.ndarray:
        return 0.5 * (self.z_faces[1:] + self.z_faces[:-1])

    # --- metrics used on the host ----------------------------------------------
    def minimum_xspacing(self) -> float:
        """min over the interior of dx^ccc (Oceananigans.Grids.minimum_xspacing)."""
        return float(np.min(self.radius * np.cos(np.pi * self.phi_c() / 180) * (self.dlam * np.pi / 180)))

    def minimum_yspacing(self) -> float:
        return float(self.radius * (self.dphi * np.pi / 180))

    # --- bathymetry ----------------------------------------------------------------
    def bottom_height_local(self, halo: Optional[int] = None) -> np.ndarray:
        """(Ny, Nx + 2*halo) bottom height including the periodic / neighbour x-halo columns
        (`partition_global_array`, bathymetry_and_forcings.jl:70-71, plus halo columns)."""
        h = self.halo if halo is None else halo
        cols = (np.arange(-h, self.Nx + h) + self.i_offset) % self.Nx_global
        if self.bottom_height_global is None:
            return np.full((self.Ny, self.Nx + 2 * h), -1.0e10)
        return np.ascontiguousarray(self.bottom_height_global[:, cols], dtype=np.float64)

    def immersed_cell_mask(self) -> np.ndarray:
        """(Nz, Ny, Nx_global) bool: GridFittedBottom centre condition z_c <= bottom."""
        if self.bottom_height_global is None:
            return np.zeros((self.Nz, self.Ny, self.Nx_global), dtype=bool)
        return self.z_c()[:, None, None] <= self.bottom_height_global[None, :, :]

    def with_rank(self, rank: int) -> "LatitudeLongitudeGrid":
        g = LatitudeLongitudeGrid(self.Nx_global, self.Ny, self.Nz, self.z_faces, self.latitude, self.longitude,
                                  rank, self.partition, self.bottom_height_global, self.radius, self.halo)
        return g


def kernel_cost_per_column(bottom_height: Optional[np.ndarray], Nx: int, Ny: int, tile: int = 16,
                           masked_weight: float = 0.55) -> np.ndarray:
    """Relative cost of one longitude column for THIS implementation: tendency tiles that touch a solid column or a
    wall run the masked (order-reducing) kernels, which cost ~1.5x on ~57 % of the step."""
    masked = np.zeros((Ny, Nx), dtype=bool)
    masked[:tile + 5, :] = True
    masked[-(tile + 5):, :] = True
    if bottom_height is not None:
        solid = bottom_height >= 0.0
        reach = tile + 5
        acc = np.zeros_like(solid)
        for d in range(-reach, reach + 1):
            acc |= np.roll(solid, d, axis=1)
        rows = np.zeros_like(acc)
        for d in range(-reach, reach + 1):
            sh = np.roll(acc, d, axis=0)
            if d > 0:
                sh[:d, :] = False
            elif d < 0:
                sh[d:, :] = False
            rows |= sh
        masked |= rows
    return Ny * 1.0 + masked_weight * masked.sum(axis=0)


def assess_x_load(immersed: np.ndarray) -> np.ndarray:
    """grid_load_balance.jl:77-85 -- number of wet cells per longitude index."""
    return np.sum(~immersed, axis=(0, 1)).astype(np.int64)


def assess_x_load_from_bottom(z_c: np.ndarray, bottom_height: np.ndarray) -> np.ndarray:
    """The same count as `assess_x_load(immersed_cell_mask)` without materialising the 3-D mask (777 M cells at 1/12 deg):
    a GridFittedBottom column is solid from the floor up to the first centre above the bottom height."""
    zc = np.asarray(z_c, dtype=np.float64)
    solid_levels = np.searchsorted(zc, np.asarray(bottom_height, dtype=np.float64), side="right")   # z_c <= bottom
    return np.sum(len(zc) - solid_levels, axis=0).astype(np.int64)


def assess_y_load(immersed: np.ndarray) -> np.ndarray:
    """grid_load_balance.jl:87-95 (dead code in the reference; kept for the y-slab option)."""
    return np.sum(~immersed, axis=(0, 2)).astype(np.int64)


def calculate_local_N(load_per_slab: Sequence[int], N: int, ranks: int) -> list:
    """grid_load_balance.jl:97-114 -- greedy prefix split, bit-for-bit: advance while
    `local_load <= total / ranks`; the last rank takes the remainder."""
    active_cells = int(np.sum(load_per_slab))
    active_load = active_cells / ranks
    local_N = [0] * ranks
    idx = 0
    for r in range(ranks - 1):
        local_load = 0
        while local_load <= active_load:
            local_load += int(load_per_slab[idx])  # IndexError past the end == Julia BoundsError
            local_N[r] += 1
            idx += 1
    local_N[-1] = N - sum(local_N[:-1])
    return local_N


def redistribute_size_to_fulfill_memory_limitation(l: list, m: int = 700) -> list:
    """grid_load_balance.jl:116-146 (in place, like the reference's `!` function)."""
    n = len(l)
    while any(x > m for x in l):
        xp = max(l); ip = l.index(xp)   # findmax / findmin return the first extremum
        xm = min(l); im = l.index(xm)
        if xp == m + 1:
            l[ip] -= 1
            l[im] += 1
        else:
            d = l[ip] - m
            q = d // n
            for t in range(n):
                l[t] += q
            l[ip] -= d
            l[im] += d % n
    while any(x < 20 for x in l):
        xp = max(l); ip = l.index(xp)
        xm = min(l); im = l.index(xm)
        if xm == 19:
            l[im] += 1
            l[ip] -= 1
        else:
            d = 20 - l[im]
            q = d // n
            for t in range(n):
                l[t] -= q
            l[ip] -= d % n
            l[im] += d
    return l


def load_balanced_grid(rank: int, ranks: int, N, latitude, z_faces, resolution, balance: bool, experiment: str,
                       *, realistic_bathymetry: Optional[Callable] = None) -> LatitudeLongitudeGrid:
    """grid_load_balance.jl:16-75.

    `realistic_bathymetry(Nx, Ny) -> (Ny, Nx) array` replaces the reference's
    `jldopen("bathymetry/bathymetry$(resolution).jld2")` (offline product, no network here).
    """
    Nx, Ny, Nz = N
    if experiment == "RealisticOcean":
        if realistic_bathymetry is None:
            raise ValueError("RealisticOcean needs a bathymetry provider (see synthetic.realistic_bathymetry)")
        bottom = np.asarray(realistic_bathymetry(Nx, Ny), dtype=np.float64)
    elif experiment == "DoubleDrake":
        g0 = LatitudeLongitudeGrid(Nx, Ny, Nz, z_faces, tuple(latitude))
        bottom = double_drake_bathymetry(g0.lam_c()[None, :], g0.phi_c()[:, None])
    else:
        bottom = None  # Quiescent: plain underlying grid

    partition = Partition.equal(Nx, ranks)
    if balance and experiment == "RealisticOcean" and ranks > 1:
        g0 = LatitudeLongitudeGrid(Nx, Ny, Nz, z_faces, tuple(latitude), bottom_height_global=bottom)
        load_per_x_slab = assess_x_load_from_bottom(g0.z_c(), bottom)      # == assess_x_load(g0.immersed_cell_mask())
        local_Nx = calculate_local_N(load_per_x_slab, Nx, ranks)
        redistribute_size_to_fulfill_memory_limitation(local_Nx, 1150)
        partition = Partition(tuple(local_Nx))
    return LatitudeLongitudeGrid(Nx, Ny, Nz, z_faces, tuple(latitude), rank=rank, partition=partition,
                                 bottom_height_global=bottom)
