"""Host-side mirror of /root/reference/src/boundary_and_initial_conditions.jl."""
from __future__ import annotations

import numpy as np

from . import _abi
from .bathymetry_and_forcings import (T_reference, exponential_profile, salinity_flux_coefficients,
                                      wind_stress_coefficients)
from .model import BoundaryConditions, HydrostaticFreeSurfaceModel, set_model

days = 86400.0
meters = 1.0


def set_boundary_conditions(experiment: str, grid, *, with_fluxes: bool = True, with_restoring: bool = True,
                            forcing_arrays=None) -> BoundaryConditions:
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
        return BoundaryConditions(kind=_abi.OB200_BC_REALISTIC, drag_mu=0.001, forcing_arrays=arrays)
    raise ValueError(f"unknown experiment {experiment}")


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
