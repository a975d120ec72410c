"""Host-side mirror of /root/reference/src/boundary_and_initial_conditions.jl."""
from __future__ import annotations

import os

import numpy as np

from . import _abi
from .bathymetry_and_forcings import (T_reference, exponential_profile, salinity_flux_coefficients,
                                      wind_stress_coefficients)
from .model import BoundaryConditions, HydrostaticFreeSurfaceModel, set_model

days = 86400.0
meters = 1.0


def set_boundary_conditions(experiment: str, grid, *, with_fluxes: bool = True, with_restoring: bool = True,
                            forcing_arrays=None, providers=None) -> BoundaryConditions:
    """boundary_and_initial_conditions.jl:48-137"""
    if experiment == "Quiescent":          # :48-56 -- zero top fluxes
        return BoundaryConditions(kind=_abi.OB200_BC_QUIESCENT)
    if experiment == "DoubleDrake":        # :58-81
        Ly = grid.Ly
        S_reference = 35.0
        u_coeffs = wind_stress_coefficients(Ly / 2)
        S_coeffs = salinity_flux_coefficients(Ly / 2, S_reference)
        lam = 5 * meters / (7 * days)      # pumping velocity of temperature restoring
        mu = 0.003                         # linear drag coefficient
        return BoundaryConditions(kind=_abi.OB200_BC_DOUBLEDRAKE, wind_coeffs=u_coeffs, salt_coeffs=S_coeffs,
                                  T_restoring_lambda=lam, drag_mu=mu)
    if experiment == "RealisticOcean":     # :83-137
        arrays = {}
        if with_fluxes:
            if forcing_arrays is None:
                raise ValueError("RealisticOcean with fluxes needs forcing arrays (see synthetic.py)")
            for k in ("taux", "tauy", "Qs", "Fs"):
                arrays[k] = forcing_arrays[k]
            if with_restoring:
                arrays["Tr"] = forcing_arrays["Tr"]
                arrays["Sr"] = forcing_arrays["Sr"]
        return BoundaryConditions(kind=_abi.OB200_BC_REALISTIC, drag_mu=0.001, forcing_arrays=arrays,
                                  providers=dict(providers or {}))
    raise ValueError(f"unknown experiment {experiment}")


_FORCING_SLOT = {"taux": 0, "tauy": 1, "Qs": 2, "Fs": 3, "Tr": 4, "Sr": 5}


def partition_global_array(grid, a: np.ndarray) -> np.ndarray:
    """Oceananigans.DistributedComputations.partition_global_array for x-slabs: this rank's columns of a global array
    whose last axis is longitude (boundary_and_initial_conditions.jl:154-157, bathymetry_and_forcings.jl:70-71)."""
    a = np.asarray(a)
    if a.shape[-1] == grid.Nx:
        return np.ascontiguousarray(a, dtype=np.float64)
    if a.shape[-1] != grid.Nx_global:
        raise ValueError(f"array with {a.shape[-1]} columns is neither the slab ({grid.Nx}) nor the globe ({grid.Nx_global})")
    return np.ascontiguousarray(a[..., grid.i_offset:grid.i_offset + grid.Nx], dtype=np.float64)


def _file_provider(model, kind: str):
    """The array source that stands in for `jldopen("data/<kind>_<n>.jld2")`: `boundary_conditions.providers[kind]`
    (a callable (grid, filenum) -> dict) or `data/<kind>_<n>.npz` with the reference's dataset names."""
    prov = getattr(model.boundary_conditions, "providers", {}).get(kind)
    if prov is not None:
        return prov

    def from_disk(grid, filenum):
        with np.load(os.path.join("data", f"{kind}_{filenum}.npz")) as f:
            ren = {"τx": "taux", "τy": "tauy"}
            return {ren.get(k, k): f[k] for k in f.files}
    return from_disk


def load_fluxes(model, filenum: int) -> None:
    """load_fluxes! (boundary_and_initial_conditions.jl:139-163): replace the four (Nx, Ny(+1), 6) flux arrays."""
    if model.grid.rank == 0:
        print(f"[ Info: loading fluxes from file fluxes_{filenum}", flush=True)
    arrays = _file_provider(model, "fluxes")(model.grid, filenum)
    for key in ("taux", "tauy", "Qs", "Fs"):
        a = partition_global_array(model.grid, arrays[key])
        model.boundary_conditions.forcing_arrays[key] = a
        model.backend.set_forcing(_FORCING_SLOT[key], a)


def load_restoring(model, filenum: int) -> None:
    """load_restoring! (:165-188)."""
    if model.grid.rank == 0:
        print(f"[ Info: loading fluxes from file restoring_{filenum}", flush=True)
    arrays = _file_provider(model, "restoring")(model.grid, filenum)
    for key in ("Tr", "Sr"):
        a = partition_global_array(model.grid, arrays[key])
        model.boundary_conditions.forcing_arrays[key] = a
        model.backend.set_forcing(_FORCING_SLOT[key], a)


def _repeat_year() -> bool:
    return os.environ.get("REPEATYEAR", "true").lower() in ("true", "1")


def flux_filenum(time: float, repeat_year: bool) -> int:
    """:194-196 -- `Int(mod(t ÷ days, 365) ÷ 5) + 1` (repeat year) or `Int(t ÷ 5days) + 1`."""
    if repeat_year:
        return int(((time // days) % 365) // 5) + 1
    return int(time // (5 * days)) + 1


def restoring_filenum(time: float, repeat_year: bool):
    """:217-223 -- every 15 days; in a repeat year only on days that are multiples of 15."""
    if repeat_year:
        year_days = (time // days) % 365
        return int(year_days // 15) + 1 if year_days % 15 == 0 else None
    return int(time // (15 * days)) + 1


def update_fluxes(sim) -> None:
    """update_fluxes! -- Callback every 5 days (boundary_and_initial_conditions.jl:190-213)."""
    load_fluxes(sim.model, flux_filenum(sim.model.clock["time"], _repeat_year()))


def update_restoring(sim) -> None:
    """update_restoring! -- Callback every day (repeat year) or 15 days (:215-238)."""
    filenum = restoring_filenum(sim.model.clock["time"], _repeat_year())
    if filenum is not None:
        load_restoring(sim.model, filenum)


def initialize_model(model: HydrostaticFreeSurfaceModel, experiment: str, *, restart: str = "",
                     initial_T=None, initial_S=None) -> None:
    """initialize_model! -- boundary_and_initial_conditions.jl:7-46"""
    if experiment == "Quiescent":          # :7
        return
    if experiment == "DoubleDrake":        # :9-18
        Depth = model.grid.Lz
        exp_length = Depth / 15.0

        def init_temperature(lam, phi, z):
            return exponential_profile(z, Lz=Depth, h=exp_length) * np.maximum(T_reference(phi), 2.0) + 0 * lam

        set_model(model, T=init_temperature, S=35.0)
        return
    if experiment == "RealisticOcean":     # :20-46 (files replaced by caller-provided arrays)
        if restart:
            return
        set_model(model, T=initial_T, S=initial_S)
        return
    raise ValueError(experiment)
