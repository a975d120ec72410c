// QG-Leith eddy viscosity (reference: src/quasi_geostrophic_leith.jl:37-169) -- optional closure.
//   k_leith_ld   calculate_deformation_radius! (:126-135)   Ld = sum_k dz sqrt(max(0, N2)) / (pi |f|)
//   k_leith_q    compute_stretching!           (:114-121)   q = f / max(N2, minN2) * I(d b)   at (c,c,f)
//   k_leith_nu   calculate_qgleith_viscosity!  (:78-104)    nu = (C ds / pi)^3 sqrt(min-limited |grad q|^2 + |grad div|^2)
// Julia's min/max propagate NaN; jmin/jmax reproduce that.
#include "kernels.h"

#define LEITH_E 2   // columns/rows beyond the interior needed by the stress divergence

__device__ __forceinline__ real jmin(real a, real b) { return (isnan(a) || isnan(b)) ? nan("") : fmin(a, b); }
__device__ __forceinline__ real jmax(real a, real b) { return (isnan(a) || isnan(b)) ? nan("") : fmax(a, b); }
__device__ __forceinline__ real f_ccc(const Geom& G, int j) { return RL(0.25) * (G.fff[j] + G.fff[j] + G.fff[j + 1] + G.fff[j + 1]); }
__device__ __forceinline__ real b_at(const Geom& G, const Fields& F, int i, int j, int k) {
    const long long c = idx3(G, i, j, k);
    return buoyancy_of(G, F.T[c], F.S[c], k);
}
__device__ __forceinline__ real dzb(const Geom& G, const Fields& F, int i, int j, int k) {   // N^2, see n2_face
    if (!active(G, i, j, k - 1) || !active(G, i, j, k)) return RL(0.0);
    const long long c = idx3(G, i, j, k);
    return n2_face(G, F.T[c - G.sz], F.S[c - G.sz], F.T[c], F.S[c], k);
}
__device__ __forceinline__ real dxb(const Geom& G, const Fields& F, int i, int j, int k) {
    if (!active(G, i - 1, j, k) || !active(G, i, j, k)) return RL(0.0);
    return (b_at(G, F, i, j, k) - b_at(G, F, i - 1, j, k)) / G.dxfc[j];
}
__device__ __forceinline__ real dyb(const Geom& G, const Fields& F, int i, int j, int k) {
    if (!active(G, i, j - 1, k) || !active(G, i, j, k)) return RL(0.0);
    return (b_at(G, F, i, j, k) - b_at(G, F, i, j - 1, k)) / G.dy;
}
__device__ __forceinline__ real divxy(const Geom& G, const Fields& F, int i, int j, int k) {
    const long long c = idx3(G, i, j, k);
    return (G.dy * F.u[c + 1] - G.dy * F.u[c] + G.dxcf[j + 1] * F.v[c + G.pitch] - G.dxcf[j] * F.v[c]) * G.razcc[j];
}
__device__ __forceinline__ real zeta_c(const Geom& G, const Fields& F, int i, int j, int k) {
    const long long c = idx3(G, i, j, k);
    real dxv = G.dy * F.v[c] - G.dy * F.v[c - 1];
    real dyu = G.dxfc[j] * F.u[c] - G.dxfc[j - 1] * F.u[c - G.pitch];
    if (inactive_cfc(G, i - 1, j, k) || inactive_cfc(G, i, j, k)) dxv = RL(0.0);
    if (inactive_fcc(G, i, j - 1, k) || inactive_fcc(G, i, j, k)) dyu = RL(0.0);
    return (dxv - dyu) * G.razcf[j];
}

__global__ void k_leith_ld(Geom G, Fields F) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x - LEITH_E;
    const int j = (int)blockIdx.y - LEITH_E;
    if (i >= G.Nx + LEITH_E) return;
    real s = RL(0.0);
    const real fc = fabs(f_ccc(G, j));
    for (int k = 0; k < G.Nz; k++) s += G.dzf[k] * (sqrt(jmax(RL(0.0), dzb(G, F, i, j, k))) / RL(M_PI) / fc);
    F.Ld[idx2(G, i, j)] = s;
}
__global__ void k_leith_q(Geom G, Fields F) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x - LEITH_E;
    const int j = (int)blockIdx.y - LEITH_E, k = blockIdx.z;   // k = 0..Nz
    if (i >= G.Nx + LEITH_E) return;
    const real coef = f_ccc(G, j) / jmax(G.leith_minN2, dzb(G, F, i, j, k));
    const real bx = RL(0.25) * (dxb(G, F, i, j, k - 1) + dxb(G, F, i + 1, j, k - 1) + dxb(G, F, i, j, k) + dxb(G, F, i + 1, j, k));
    const real by = RL(0.25) * (dyb(G, F, i, j, k - 1) + dyb(G, F, i, j + 1, k - 1) + dyb(G, F, i, j, k) + dyb(G, F, i, j + 1, k));
    const long long c = idx3(G, i, j, k);
    F.qx[c] = coef * bx;
    F.qy[c] = coef * by;
}
__global__ void k_leith_nu(Geom G, Fields F) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x - LEITH_E;
    const int j = (int)blockIdx.y - LEITH_E, k = blockIdx.z;
    if (i >= G.Nx + LEITH_E) return;
    const long long c = idx3(G, i, j, k);
    auto dxz = [&](int a, int b) { return (zeta_c(G, F, a + 1, b, k) - zeta_c(G, F, a, b, k)) / G.dxcf[b]; };
    auto dyz = [&](int a, int b) { return (zeta_c(G, F, a, b + 1, k) - zeta_c(G, F, a, b, k)) / G.dy; };
    real dzx = RL(0.5) * (dxz(i, j) + dxz(i, j + 1));
    real dzy = RL(0.5) * (dyz(i, j) + dyz(i + 1, j));
    const real dfy = RL(0.5) * ((G.fff[j + 1] - G.fff[j]) / G.dy + (G.fff[j + 1] - G.fff[j]) / G.dy);
    dzx += RL(0.0); dzy += dfy;
    const bool act = active(G, i, j, k);
    const real dqx = act ? (F.qx[c + G.sz] - F.qx[c]) / G.dzc[k] : RL(0.0);
    const real dqy = act ? (F.qy[c + G.sz] - F.qy[c]) / G.dzc[k] : RL(0.0);
    auto ddx = [&](int a, int b) {
        if (!active(G, a - 1, b, k) || !active(G, a, b, k)) return RL(0.0);
        return (divxy(G, F, a, b, k) - divxy(G, F, a - 1, b, k)) / G.dxfc[b];
    };
    auto ddy = [&](int a, int b) {
        if (!active(G, a, b - 1, k) || !active(G, a, b, k)) return RL(0.0);
        return (divxy(G, F, a, b, k) - divxy(G, F, a, b - 1, k)) / G.dy;
    };
    const real ddx_ = RL(0.5) * (ddx(i, j) + ddx(i + 1, j)), ddy_ = RL(0.5) * (ddy(i, j) + ddy(i, j + 1));
    const real dd2 = ddx_ * ddx_ + ddy_ * ddy_;
    const real fc = f_ccc(G, j);
    const real dz2 = dzx * dzx + dzy * dzy;
    const real dq2 = (dqx + dzx) * (dqx + dzx) + (dqy + dzy) * (dqy + dzy);
    const real A = RL(2.0) * (RL(1.0) / (RL(1.0) / (G.dxfc[j] * G.dxfc[j]) + RL(1.0) / (G.dy * G.dy)));
    const real Ds = pow(A, RL(0.5));
    const real ld = F.Ld[idx2(G, i, j)];
    const real Bu = ld * ld / A;
    const real Ro = G.leith_Vscale / (fc * Ds);
    const real t1 = RL(1.0) + RL(1.0) / Bu, t2 = RL(1.0) + RL(1.0) / (Ro * Ro);
    real dQ2 = jmin(dq2, dz2 * (t1 * t1));
    dQ2 = jmin(dQ2, dz2 * (t2 * t2));
    F.nuL[c] = pow(G.leith_C * Ds / RL(M_PI), RL(3.0)) * sqrt(dQ2 + dd2);
}

void launch_leith(const Geom& G, const Fields& F, cudaStream_t st) {
    dim3 b(128);
    dim3 g2((G.Nx + 2 * LEITH_E + 127) / 128, G.Ny + 2 * LEITH_E);
    k_leith_ld<<<g2, b, 0, st>>>(G, F);
    OB_COUNT_LAUNCH(1);
    dim3 g3((G.Nx + 2 * LEITH_E + 127) / 128, G.Ny + 2 * LEITH_E, G.Nz + 1);
    k_leith_q<<<g3, b, 0, st>>>(G, F);
    OB_COUNT_LAUNCH(1);
    dim3 g4((G.Nx + 2 * LEITH_E + 127) / 128, G.Ny + 2 * LEITH_E, G.Nz);
    k_leith_nu<<<g4, b, 0, st>>>(G, F);
    OB_COUNT_LAUNCH(1);
}
