"""Synthetic, closed-form stand-ins for the RealisticOcean input files (no network here):
`bathymetry/bathymetry$(res).jld2`, `data/initial_{T,S}_at_k*.jld2`, `data/fluxes_*.jld2`, `data/restoring_*.jld2`
(/root/reference/src/bathymetry_and_forcings.jl:70-73, src/boundary_and_initial_conditions.jl:20-46,139-188).
Shapes and value ranges follow the ECCO-derived products the reference regrids; formulas are fixed,
no RNG, so every rank can evaluate its own slab.
"""
from __future__ import annotations

import numpy as np


def realistic_bathymetry(Nx: int, Ny: int, depth: float = 5244.5) -> np.ndarray:
    """(Ny, Nx) bottom height: smooth analytic continents (~30 % land columns), depths 6 .. `depth` m."""
    lam = -180.0 + (np.arange(Nx) + 0.5) * 360.0 / Nx
    phi = -75.0 + (np.arange(Ny) + 0.5) * 150.0 / Ny
    L, P = np.meshgrid(lam, phi)
    r = np.pi / 180.0
    land = (1.1 * np.exp(-((L + 100.0) / 28.0) ** 2 - ((P - 35.0) / 30.0) ** 2)        # "North America"
            + 1.0 * np.exp(-((L + 60.0) / 14.0) ** 2 - ((P + 15.0) / 28.0) ** 2)       # "South America"
            + 1.2 * np.exp(-((L - 20.0) / 22.0) ** 2 - ((P - 5.0) / 32.0) ** 2)        # "Africa"
            + 1.3 * np.exp(-((L - 90.0) / 45.0) ** 2 - ((P - 50.0) / 22.0) ** 2)       # "Eurasia"
            + 0.9 * np.exp(-((L - 135.0) / 16.0) ** 2 - ((P + 25.0) / 11.0) ** 2))     # "Australia"
    ridges = 0.35 * (np.sin(3 * L * r) * np.cos(2 * P * r)) ** 2
    h = -depth * np.clip(1.0 - 1.6 * np.clip(land - 0.25, 0.0, None) - ridges, 0.0012, 1.0)
    return np.where(land > 0.62, 50.0, h)


def initial_T_S(grid):
    """(Nz, Ny, Nx) T in [-2, 30], S in [33, 37] (zonally varying analytic profiles)."""
    r = np.pi / 180.0
    lam, phi, z = grid.lam_c()[None, None, :], grid.phi_c()[None, :, None], grid.z_c()[:, None, None]
    T = -2.0 + 32.0 * np.cos(phi * r) ** 2 * np.exp(z / 900.0) * (1.0 + 0.05 * np.sin(2 * lam * r))
    S = 35.0 + 1.5 * np.cos(2 * phi * r) * np.exp(z / 1500.0) + 0.3 * np.sin(3 * lam * r) * np.exp(z / 600.0)
    return np.clip(T, -2.0, 30.0), np.clip(S, 33.0, 37.0)


def forcing_arrays(grid, flux_filenum: int = 1, restoring_filenum: int = 1) -> dict:
    """taux (6,Ny,Nx), tauy (6,Ny+1,Nx), Qs, Fs, Tr, Sr (6,Ny,Nx): six time slices -- days 0..5 of flux file
    `flux_filenum` (data/fluxes_<n>.jld2) and days 0, 3, .., 15 of restoring file `restoring_filenum`
    (data/restoring_<n>.jld2).  Consecutive files are continuous: slice 6 of file n equals slice 1 of file n + 1."""
    r = np.pi / 180.0
    lam = grid.lam_c()[None, None, :]
    phi, phif = grid.phi_c()[None, :, None], grid.phi_f()[None, :, None]
    t = (np.arange(6, dtype=np.float64) + 5.0 * (flux_filenum - 1))[:, None, None]
    tr = (3.0 * np.arange(6, dtype=np.float64) + 15.0 * (restoring_filenum - 1))[:, None, None] / 3.0
    taux = -1e-4 * np.cos(3 * phi * r) * (1.0 + 0.1 * np.sin(lam * r + 0.4 * t))
    tauy = 2e-5 * np.sin(2 * phif * r) * np.cos(2 * lam * r + 0.3 * t)
    Qs = 2e-5 * np.cos(2 * phi * r) * (1.0 + 0.2 * np.sin(lam * r - 0.2 * t))
    Fs = 1e-6 * np.sin(3 * phi * r) * np.cos(lam * r + 0.1 * t)
    Tr = -2.0 + 31.0 * np.cos(phi * r) ** 2 + 0.2 * np.sin(lam * r + 0.5 * tr)
    Sr = 35.0 + 1.2 * np.cos(2 * phi * r) + 0.05 * np.cos(2 * lam * r - 0.3 * tr)
    return {k: np.ascontiguousarray(v) for k, v in dict(taux=taux, tauy=tauy, Qs=Qs, Fs=Fs, Tr=Tr, Sr=Sr).items()}


def flux_file(grid, filenum: int) -> dict:
    """Stand-in for `jldopen("data/fluxes_$(filenum).jld2")` already partitioned to the rank's slab."""
    a = forcing_arrays(grid, flux_filenum=filenum)
    return {k: a[k] for k in ("taux", "tauy", "Qs", "Fs")}


def restoring_file(grid, filenum: int) -> dict:
    """Stand-in for `jldopen("data/restoring_$(filenum).jld2")` already partitioned to the rank's slab."""
    a = forcing_arrays(grid, restoring_filenum=filenum)
    return {k: a[k] for k in ("Tr", "Sr")}
