"""Host-side mirror of /root/reference/src/bathymetry_and_forcings.jl.

Vertical grids, the DoubleDrake bathymetry and the idealised piecewise-cubic
wind-stress / salinity-flux profiles.  Pure host logic (numpy); the per-cell
evaluation of the fluxes happens inside the CUDA tendency kernel, which receives
the cubic coefficients computed here through `ob200_config`.
"""
from __future__ import annotations

import math

import numpy as np


def exponential_profile(z, *, Lz, h):
    """bathymetry_and_forcings.jl:45"""
    return (np.exp(z / h) - np.exp(-Lz / h)) / (1 - np.exp(-Lz / h))


def exponential_z_faces(Nz: int, Depth: float, *, h: float | None = None) -> np.ndarray:
    """bathymetry_and_forcings.jl:47-58 -- Nz+1 faces, bottom (-Depth) to top (0)."""
    if h is None:
        h = Nz / 4.5
    z_faces = exponential_profile(np.arange(1, Nz + 2, dtype=np.float64), Lz=Nz, h=h)
    z_faces = z_faces - z_faces[0]
    z_faces = z_faces * (-Depth / z_faces[-1])
    z_faces[0] = 0.0
    return np.ascontiguousarray(z_faces[::-1])


def linear_z_faces(Nz: int, Depth: float, *, first_delta: float = 5) -> np.ndarray:
    """bathymetry_and_forcings.jl:33-43"""
    m = (Depth - (Nz * first_delta)) / sum(range(1, Nz))
    delta = m * np.arange(0, Nz) + first_delta
    z_faces = np.zeros(Nz + 1)
    for k in range(Nz - 1, -1, -1):  # Julia k = Nz:-1:1
        z_faces[k] = z_faces[k + 1] - delta[Nz - 1 - k]
    return z_faces


def double_drake_bathymetry(lam, phi):
    """bathymetry_and_forcings.jl:60-67 (vectorised): two one-degree ridges north of 35S."""
    lam = np.asarray(lam, dtype=np.float64)
    phi = np.asarray(phi, dtype=np.float64)
    lam, phi = np.broadcast_arrays(lam, phi)
    ridge = (phi > -35) & (((lam > 0) & (lam < 1)) | ((lam > 90) & (lam < 91)))
    return np.where(ridge, 0.0, -10000.0)


def cubic_profile(x1, x2, y1, y2, d1, d2):
    """bathymetry_and_forcings.jl:75-86 -- dense 4x4 solve (Julia `A \\ b` = LAPACK LU, as numpy)."""
    A = np.array([[x1**3, x1**2, x1, 1.0],
                  [x2**3, x2**2, x2, 1.0],
                  [3 * x1**2, 2 * x1, 1.0, 0.0],
                  [3 * x2**2, 2 * x2, 1.0, 0.0]], dtype=np.float64)
    b = np.array([y1, y2, d1, d2], dtype=np.float64)
    return tuple(np.linalg.solve(A, b))


def wind_stress_coefficients(south_north_limit):
    """bathymetry_and_forcings.jl:89-98"""
    L = float(south_north_limit)
    pieces = [cubic_profile(-L, -45.0, 0.0, 0.2, 0.0, 0.0),
              cubic_profile(-45.0, -15.0, 0.2, -0.1, 0.0, 0.0),
              cubic_profile(-15.0, 0.0, -0.1, -0.02, 0.0, 0.0),
              cubic_profile(0.0, 15.0, -0.02, -0.1, 0.0, 0.0),
              cubic_profile(15.0, 45.0, -0.1, 0.1, 0.0, 0.0),
              cubic_profile(45.0, L, 0.1, 0.0, 0.0, 0.0)]
    return tuple(tuple(c / 1000 for c in p) for p in pieces)


def salinity_flux_coefficients(south_north_limit, initial_salinity):
    """bathymetry_and_forcings.jl:101-108"""
    L = float(south_north_limit)
    pieces = [cubic_profile(-L, -20.0, -2e-8, 2e-8, 0.0, 0.0),
              cubic_profile(-20.0, 0.0, 2e-8, -4e-8, 0.0, 0.0),
              cubic_profile(0.0, 20.0, -4e-8, 2e-8, 0.0, 0.0),
              cubic_profile(20.0, L, 2e-8, -2e-8, 0.0, 0.0)]
    return tuple(tuple(c * initial_salinity for c in p) for p in pieces)


def wind_stress(phi, coeffs):
    """bathymetry_and_forcings.jl:110-119"""
    c = (coeffs[0] if phi < -45 else coeffs[1] if phi < -15 else coeffs[2] if phi < 0
         else coeffs[3] if phi < 15 else coeffs[4] if phi < 45 else coeffs[5])
    return c[0] * phi**3 + c[1] * phi**2 + c[2] * phi + c[3]


def salinity_flux(phi, coeffs):
    """bathymetry_and_forcings.jl:121-128"""
    c = coeffs[0] if phi < -20 else coeffs[1] if phi < 0 else coeffs[2] if phi < 20 else coeffs[3]
    return c[0] * phi**3 + c[1] * phi**2 + c[2] * phi + c[3]


def T_reference(phi):
    """bathymetry_and_forcings.jl:154"""
    return np.maximum(0.0, 30.0 * np.cos(1.2 * math.pi * np.asarray(phi, dtype=np.float64) / 180))
