// Flux boundary-condition functions evaluated on the device
// (reference: src/bathymetry_and_forcings.jl:130-193, src/boundary_and_initial_conditions.jl:58-137).
#pragma once
#include "kernels.h"

// flux_from_interpolated_array (bathymetry_and_forcings.jl:162-169): arrays are Nx x nyf x 6, compact
__device__ __forceinline__ real interp_flux(const Geom& G, const real* F, int i, int j, int nyf, double time,
                                            bool restoring) {
    if (F == nullptr) return RL(0.0);
    // the interpolation index comes from the Float64 model clock in both precisions
    const double days = time / 86400.0; /*f64*/
    double n; /*f64*/
    if (!restoring) n = fmod(days, 5.0) + 1.0; /*f64*/
    else n = floor(fmod(days, 15.0) / 3.0) + 1.0;   // `mod(t, 15) ÷ 3 + 1` (:180) /*f64*/
    const int n1 = (int)floor(n), n2 = n1 + 1;
    const size_t plane = (size_t)G.Nx * nyf;
    const real a = F[(size_t)(n1 - 1) * plane + (size_t)j * G.Nx + i];
    const real b = F[(size_t)(n2 - 1) * plane + (size_t)j * G.Nx + i];
    return a * (real)(n2 - n) + b * (real)(n - n1);
}

__device__ __forceinline__ real dz_min(const Geom& G) {
    real m = RL(1e300);
    for (int k = 0; k < G.Nz; k++) m = fmin(m, G.dzc[k]);
    return m;
}

__device__ __forceinline__ real top_flux_T(const Geom& G, const Fields& F, int i, int j, real Ttop, double time) {
    if (G.bc_kind == OB200_BC_DOUBLEDRAKE) return G.lambda_T * (Ttop - G.tref[j]);   // T_relaxation (:156-159)
    if (G.bc_kind == OB200_BC_REALISTIC) {
        const bool inx = (unsigned)i < (unsigned)G.Nx;
        if (!inx) return RL(0.0);
        real fl = interp_flux(G, F.forcing[2], i, j, G.Ny, time, false);
        if (F.forcing[4]) fl += (dz_min(G) / (RL(30.0) * RL(86400.0))) * (Ttop - interp_flux(G, F.forcing[4], i, j, G.Ny, time, true));
        return fl;
    }
    return RL(0.0);
}
__device__ __forceinline__ real top_flux_S(const Geom& G, const Fields& F, int i, int j, real Stop, double time) {
    if (G.bc_kind == OB200_BC_DOUBLEDRAKE) return G.sflux[j];                         // surface_salinity_flux (:149-152)
    if (G.bc_kind == OB200_BC_REALISTIC) {
        const bool inx = (unsigned)i < (unsigned)G.Nx;
        if (!inx) return RL(0.0);
        real fl = interp_flux(G, F.forcing[3], i, j, G.Ny, time, false);
        if (F.forcing[5]) fl += (dz_min(G) / (RL(60.0) * RL(86400.0))) * (Stop - interp_flux(G, F.forcing[5], i, j, G.Ny, time, true));
        return fl;
    }
    return RL(0.0);
}
__device__ __forceinline__ real top_buoyancy_flux(const Geom& G, const Fields& F, int i, int j, double time) {
    if (G.bc_kind == OB200_BC_QUIESCENT) return RL(0.0);
    const long long c = idx3(G, i, j, G.Nz - 1);
    const real Tt = F.T[c], St = F.S[c];
    real al = G.alpha, be = G.beta_s;
    if (G.eos == OB200_EOS_TEOS10) {
        real dT, dS;
        teos_rprime<true>(Tt, St, G.zc[G.Nz - 1], &dT, &dS);
        al = -dT / G.rho0; be = dS / G.rho0;
    }
    return G.g * (al * top_flux_T(G, F, i, j, Tt, time) - be * top_flux_S(G, F, i, j, St, time));
}
