// Shared device-side definitions of libocean_b200: geometry, indexing, node predicates,
// equation of state, WENO wrappers.  sm_100a only; `real` = Float64 (default build) or Float32 (-DOB200_F32).
//
// HBM layout (DESIGN.md "Data layout"): every 3-D field is one array of (Nz+1) planes x NyP rows x
// `pitch` doubles, i (longitude) contiguous.  Interior column 0 sits at element X0 = 16 of a row (128-byte
// aligned since pitch % 16 == 0 and the base is 256-byte aligned); 7 halo columns on each side are used
// (periodic copy on one GPU, NCCL-exchanged neighbour columns on several).  Rows j = -7 .. Ny+7 (v has the
// extra face row Ny).  No vertical halos are stored: vertical stencils are 2-point and clamp.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/ocean_b200.h"

// Arithmetic type of the library (the reference's PRECISION, experiments/run.jl:38): Float64, or Float32 when built with
// -DOB200_F32 (libocean_b200_f32.so).  Every floating literal of the device code is written RL(x) so that the Float32
// build has no Float64 arithmetic; the model clock (`time`) and the diagnostics stay Float64 in both.
#ifdef OB200_F32
typedef float real;
#else
typedef double real;
#endif
#define RL(x) ((real)(x))

#define OB_H 7     // horizontal halo used by stencils
#define OB_X0 16   // element offset of interior column 0 inside a row

struct Geom {
    int Nx, Ny, Nz;
    int pitch, NyP;          // row pitch (elements), number of rows
    int pitch2, X0b, Wb;     // barotropic 2-D arrays: own pitch / origin so that they can carry a halo of Wb columns
    long long sz;            // plane stride = pitch * NyP
    // per-latitude metrics, indexable for j = -OB_H-1 .. Ny+OB_H+1 (pointers pre-offset so [j] works)
    const real *dxfc, *dxcf, *azcc, *azcf, *fff, *taux, *sflux, *tref;
    const real *rdxfc, *razcc, *razcf;   // reciprocals used by the tendency kernels (DESIGN.md spec decision 8)
    // vertical metrics: dzc[k], zc[k] for k = -1..Nz (pre-offset), dzf[k] for k = 0..Nz
    const real *dzc, *dzf, *zc, *rdzc, *rdzf, *rupz, *rloz;   // rupz[k] = 1/(dzc[k] dzf[k+1]), rloz[k] = 1/(dzc[k] dzf[k])
    const real* zfa;       // z of the (c,c,f) faces, k = 0..Nz (the caller's z_faces)
    const int* kb;           // first active level per column (2-D, field layout); Nz = no active cell
    const int* kfast;        // level from which the full-order, mask-free stencils are valid (2-D)
    real dy, rdy;
    real ab2_ca, ab2_cb;     // 3/2 + chi, 1/2 + chi of the regular AB2 step (the tendency kernels write ca G^n - cb G^-)
    real g, alpha, beta_s, rho0;
    int eos;
    real nu_z, kappa_z;
    int ri_enabled;
    real ri_nu0, ri_kappa0, ri_kappa_ca, ri_Cen, ri_Cav, ri_Ri0, ri_Ridelta, r1cav, rRidelta;
    int bc_kind;
    real lambda_T, mu_drag;
    int qgleith;
    real leith_C, leith_minN2, leith_Vscale;
};

__device__ __forceinline__ long long idx3(const Geom& G, int i, int j, int k) {
    return (long long)k * G.sz + (long long)(j + OB_H) * G.pitch + (i + OB_X0);
}
__device__ __forceinline__ int idx2(const Geom& G, int i, int j) { return (j + OB_H) * G.pitch + (i + OB_X0); }
// barotropic arrays (eta, U, V, averages, G^U, G^V, H): wide x-halo on multi-GPU runs
__device__ __forceinline__ int idx2b(const Geom& G, int i, int j) { return (j + OB_H) * G.pitch2 + (i + G.X0b); }

// ---- node predicates.  kb rows outside the y-domain hold Nz, so the j-range test is implicit.
__device__ __forceinline__ bool active(const Geom& G, int i, int j, int k) {
    return (unsigned)k < (unsigned)G.Nz && k >= __ldg(&G.kb[idx2(G, i, j)]);
}
__device__ __forceinline__ bool in_domain(const Geom& G, int j, int k) {
    return (unsigned)j < (unsigned)G.Ny && (unsigned)k < (unsigned)G.Nz;
}
__device__ __forceinline__ bool inactive_fcc(const Geom& G, int i, int j, int k) { return !active(G, i - 1, j, k) && !active(G, i, j, k); }
__device__ __forceinline__ bool inactive_cfc(const Geom& G, int i, int j, int k) { return !active(G, i, j - 1, k) && !active(G, i, j, k); }
__device__ __forceinline__ bool peripheral_fcc(const Geom& G, int i, int j, int k) { return !active(G, i - 1, j, k) || !active(G, i, j, k); }
__device__ __forceinline__ bool peripheral_cfc(const Geom& G, int i, int j, int k) { return !active(G, i, j - 1, k) || !active(G, i, j, k); }
__device__ __forceinline__ bool peripheral_ccf(const Geom& G, int i, int j, int k) { return !active(G, i, j, k - 1) || !active(G, i, j, k); }
__device__ __forceinline__ bool peripheral_fcf(const Geom& G, int i, int j, int k) {
    return !active(G, i - 1, j, k - 1) || !active(G, i, j, k - 1) || !active(G, i - 1, j, k) || !active(G, i, j, k);
}
__device__ __forceinline__ bool peripheral_cff(const Geom& G, int i, int j, int k) {
    return !active(G, i, j - 1, k - 1) || !active(G, i, j, k - 1) || !active(G, i, j - 1, k) || !active(G, i, j, k);
}

// ---- TEOS-10 55-term polynomial (Roquet et al. 2015; SeawaterPolynomials 0.3.2) [OCN-recall coefficients].
// Stored as R[k][j][i] (powers of zeta, tau, s); evaluated by nested Horner over the triangular support.
static __device__ __constant__ real TEOS_R[4][7][7] = {
    {{RL(8.0189615746e+02), RL(8.6672408165e+02), -RL(1.7864682637e+03), RL(2.0375295546e+03), -RL(1.2849161071e+03), RL(4.3227585684e+02), -RL(6.0579916612e+01)},
     {RL(2.6010145068e+01), -RL(6.5281885265e+01), RL(8.1770425108e+01), -RL(5.6888046321e+01), RL(1.7681814114e+01), -RL(1.9193502195e+00), 0},
     {-RL(3.7074170417e+01), RL(6.1548258127e+01), -RL(6.0362551501e+01), RL(2.9130021253e+01), -RL(5.4723692739e+00), 0, 0},
     {RL(2.1661789529e+01), -RL(3.3449108469e+01), RL(1.9717078466e+01), -RL(3.1742946532e+00), 0, 0, 0},
     {-RL(8.3627885467e+00), RL(1.1311538584e+01), -RL(5.3563304045e+00), 0, 0, 0, 0},
     {RL(5.4048723791e-01), RL(4.8169980163e-01), 0, 0, 0, 0, 0},
     {-RL(1.9083568888e-01), 0, 0, 0, 0, 0, 0}},
    {{RL(1.9681925209e+01), -RL(4.2549998214e+01), RL(5.0774768218e+01), -RL(3.0938076334e+01), RL(6.6051753097e+00), 0, 0},
     {-RL(1.3336301113e+01), -RL(4.4870114575e+00), RL(5.0042598061e+00), -RL(6.5399043664e-01), 0, 0, 0},
     {RL(6.7080479603e+00), RL(3.5063081279e+00), -RL(1.8795372996e+00), 0, 0, 0, 0},
     {-RL(2.4649669534e+00), -RL(5.5077101279e-01), 0, 0, 0, 0, 0},
     {RL(5.5927935970e-01), 0, 0, 0, 0, 0, 0},
     {0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0}},
    {{RL(2.0660924175e+00), -RL(4.9527603989e+00), RL(2.5019633244e+00), 0, 0, 0, 0},
     {RL(2.0564311499e+00), -RL(2.1311365518e-01), 0, 0, 0, 0, 0},
     {-RL(1.2419983026e+00), 0, 0, 0, 0, 0, 0},
     {0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0}},
    {{-RL(2.3342758797e-02), -RL(1.8507636718e-02), 0, 0, 0, 0, 0},
     {RL(3.7969820455e-01), 0, 0, 0, 0, 0, 0},
     {0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0}}};

// r'(tau, s, zeta) and optionally its derivatives w.r.t. Theta and S_A
template <bool DERIV>
__device__ __forceinline__ real teos_rprime(real Th, real SA, real Z, real* dTh, real* dSA) {
    const real Sau = RL(40.0) * RL(35.16504) / RL(35.0), Thu = RL(40.0), Zu = RL(1e4), dS = RL(32.0);
    const real tau = Th / Thu, s = sqrt((SA + dS) / Sau), ze = -Z / Zu;
    const int deg[4] = {6, 4, 2, 1};
    real r = RL(0.0), rt = RL(0.0), rs = RL(0.0);
#pragma unroll
    for (int k = 3; k >= 0; k--) {
        real pj = RL(0.0), pjt = RL(0.0), pjs = RL(0.0);
#pragma unroll
        for (int j = 6; j >= 0; j--) {
            if (j > deg[k]) continue;
            real pi = RL(0.0), pis = RL(0.0);
#pragma unroll
            for (int i = 6; i >= 0; i--) {
                if (i + j > deg[k]) continue;
                if (DERIV) pis = pis * s + pi;
                pi = pi * s + TEOS_R[k][j][i];
            }
            if (DERIV) { pjt = pjt * tau + pj; pjs = pjs * tau + pis; }
            pj = pj * tau + pi;
        }
        r = r * ze + pj;
        if (DERIV) { rt = rt * ze + pjt; rs = rs * ze + pjs; }
    }
    if (DERIV) { *dTh = rt / Thu; *dSA = rs / (RL(2.0) * s * Sau); }
    return r;
}

// compile-time equation of state (the column sweeps are instantiated per EOS: the TEOS-10 polynomial costs registers the
// linear instance does not have under its 32-register occupancy bound)
template <bool TEOS>
__device__ __forceinline__ real buoyancy_t(const Geom& G, real T, real S, int k) {
    if (!TEOS) return G.g * (G.alpha * T - G.beta_s * S);
    return -G.g * teos_rprime<false>(T, S, G.zc[k], nullptr, nullptr) / G.rho0;
}
__device__ __forceinline__ real buoyancy_of(const Geom& G, real T, real S, int k) {
    return G.eos == OB200_EOS_LINEAR ? buoyancy_t<false>(G, T, S, k) : buoyancy_t<true>(G, T, S, k);
}

// N^2 on the (c,c,f) face kf between the cells kf - 1 (Tm, Sm) and kf (Tk, Sk): Oceananigans' `∂z_b` for SeawaterBuoyancy,
//     g (alpha ∂z T - beta ∂z S),   alpha, beta evaluated AT THE FACE from the z-averaged T, S and the face depth
// -- not a difference of in-situ buoyancies, which for TEOS-10 would add the compressibility term ∂rho'/∂z at fixed T, S.
// The caller applies the immersed conditional (zero when either cell is inactive).
template <bool TEOS>
__device__ __forceinline__ real n2_face_t(const Geom& G, real Tm, real Sm, real Tk, real Sk, int kf) {
    real al = G.alpha, be = G.beta_s;
    if (TEOS) {
        real dTh, dSA;
        teos_rprime<true>(RL(0.5) * (Tm + Tk), RL(0.5) * (Sm + Sk), G.zfa[kf], &dTh, &dSA);
        al = -dTh / G.rho0; be = dSA / G.rho0;
    }
    const real rdzf = G.rdzf[kf];
    return G.g * (al * ((Tk - Tm) * rdzf) - be * ((Sk - Sm) * rdzf));
}
__device__ __forceinline__ real n2_face(const Geom& G, real Tm, real Sm, real Tk, real Sk, int kf) {
    return G.eos == OB200_EOS_LINEAR ? n2_face_t<false>(G, Tm, Sm, Tk, Sk, kf) : n2_face_t<true>(G, Tm, Sm, Tk, Sk, kf);
}

// Deterministic tanh from +,-,*,/ only (the library is compiled with -fmad=false): the same operation
// sequence as the CPU oracle, so kappa agrees bit for bit; |error| < 3e-16 absolute.
__device__ __forceinline__ real det_exp_neg(real y) {   // exp(y), y <= 0
    const real LOG2E = RL(1.4426950408889634), LN2_HI = RL(6.93147180369123816490e-01), LN2_LO = RL(1.90821492927058770002e-10);
    const real n = rint(y * LOG2E);
    const real r = (y - n * LN2_HI) - n * LN2_LO;
    real p = RL(1.0) / RL(6227020800.0);
    const real inv_fact[13] = {RL(1.0) / RL(479001600.0), RL(1.0) / RL(39916800.0), RL(1.0) / RL(3628800.0), RL(1.0) / RL(362880.0), RL(1.0) / RL(40320.0),
                                 RL(1.0) / RL(5040.0), RL(1.0) / RL(720.0), RL(1.0) / RL(120.0), RL(1.0) / RL(24.0), RL(1.0) / RL(6.0), RL(0.5), RL(1.0), RL(1.0)};
#pragma unroll
    for (int m = 0; m < 13; m++) p = p * r + inv_fact[m];
    return ldexp(p, (int)n);
}
__device__ __forceinline__ real det_tanh(real x) {
    const real a = fabs(x);
    if (a > RL(20.0)) return x > 0 ? RL(1.0) : -RL(1.0);
    const real t = det_exp_neg(-RL(2.0) * a);
    const real r = (RL(1.0) - t) / (RL(1.0) + t);
    return x < 0 ? -r : r;
}

__device__ __forceinline__ real upwind_flux(real vel, real L, real R) {
    const real a = fabs(vel);
    return RL(0.5) * ((vel + a) * L + (vel - a) * R);
}

#define OB_CUDA_CHECK(h, expr)                                                         \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) return (h)->fail(OB200_ECUDA, #expr, cudaGetErrorString(_e)); \
    } while (0)
