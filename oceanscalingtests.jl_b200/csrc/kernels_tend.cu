// Tendency kernels: fused momentum (Gu, Gv) and fused tracer (GT, GS) tendencies incl. flux BCs.
//
//   Gu = +upwind(vhat, zeta)        WENO-9 vorticity flux, VelocityStencil smoothness
//        -(1/V)(Phi_delta + dz(W u)) WENO-5 self-upwinded divergence flux + centred vertical flux
//        -dx(K) + f * <v> - dx(pHY') + BCs            (Gv mirrors)
//   Gc = -(1/V)(dx Fx + dy Fy + dz Fz) + BCs          WENO-7 in x and y, centred in z
//
// Upwinding evaluates ONLY the upwind-side reconstruction: 0.5*((a+|a|)L + (a-|a|)R) equals a*L (a > 0) or
// a*R (a < 0) exactly in IEEE arithmetic, so this halves the WENO work without changing a bit.
// Two paths per cell: a mask-free full-order path (k >= kfast(i,j): every cell within 5 columns/rows is wet and
// inside the domain) and a general path with per-reconstruction order reduction and conditional operators.
#include "ob_bc.cuh"
#include "weno_gen.cuh"

struct Mom {
    const Geom& G;
    const double* __restrict__ u;
    const double* __restrict__ v;
    __device__ __forceinline__ double U(int i, int j, int k) const { return u[idx3(G, i, j, k)]; }
    __device__ __forceinline__ double V(int i, int j, int k) const { return v[idx3(G, i, j, k)]; }
    // vertical vorticity at (f,f,c); GEN applies the immersed conditional differences
    template <bool GEN>
    __device__ __forceinline__ double zeta(int i, int j, int k) const {
        const long long c = idx3(G, i, j, k);
        double dxv = G.dy * v[c] - G.dy * v[c - 1];
        double dyu = G.dxfc[j] * u[c] - G.dxfc[j - 1] * u[c - G.pitch];
        if (GEN) {
            if (inactive_cfc(G, i - 1, j, k) || inactive_cfc(G, i, j, k)) dxv = 0.0;
            if (inactive_fcc(G, i, j - 1, k) || inactive_fcc(G, i, j, k)) dyu = 0.0;
        }
        return (dxv - dyu) / G.azcf[j];
    }
    __device__ __forceinline__ double dxU(int i, int j, int k) const {
        const long long c = idx3(G, i, j, k);
        const double ax = G.dy * G.dzc[k];
        return ax * u[c + 1] - ax * u[c];
    }
    __device__ __forceinline__ double dyV(int i, int j, int k) const {
        const long long c = idx3(G, i, j, k);
        return G.dxcf[j + 1] * G.dzc[k] * v[c + G.pitch] - G.dxcf[j] * G.dzc[k] * v[c];
    }
    __device__ __forceinline__ double Kh(int i, int j, int k) const {
        const long long c = idx3(G, i, j, k);
        const double a = u[c], b = u[c + 1], cc = v[c], d = v[c + G.pitch];
        return (0.5 * (a * a + b * b) + 0.5 * (cc * cc + d * d)) / 2.0;
    }
};

// ---- biased reconstructions along a line; `base + sgn*m` walks from the most upwind-ward point -------------
template <int N, bool GEN>
__device__ __forceinline__ double recon_vort_y(const Mom& M, int i, int j, int k, bool left) {
    const int base = left ? j - N + 1 : j + N, sgn = left ? 1 : -1;
    double zq[2 * N - 1], us[2 * N - 1], vs[2 * N - 1];
#pragma unroll
    for (int m = 0; m < 2 * N - 1; m++) {
        const int jj = base + sgn * m;
        zq[m] = M.template zeta<GEN>(i, jj, k);
        us[m] = 0.5 * (M.U(i, jj - 1, k) + M.U(i, jj, k));
        vs[m] = 0.5 * (M.V(i - 1, jj, k) + M.V(i, jj, k));
    }
    if (N == 1) return zq[0];
    double bu[N], bv[N];
    Weno<N>::beta(us, bu);
    Weno<N>::beta(vs, bv);
#pragma unroll
    for (int s = 0; s < N; s++) bu[s] = 0.5 * (bu[s] + bv[s]);
    return Weno<N>::combine(zq, bu);
}
template <int N, bool GEN>
__device__ __forceinline__ double recon_vort_x(const Mom& M, int i, int j, int k, bool left) {
    const int base = left ? i - N + 1 : i + N, sgn = left ? 1 : -1;
    double zq[2 * N - 1], us[2 * N - 1], vs[2 * N - 1];
#pragma unroll
    for (int m = 0; m < 2 * N - 1; m++) {
        const int ii = base + sgn * m;
        zq[m] = M.template zeta<GEN>(ii, j, k);
        us[m] = 0.5 * (M.U(ii, j - 1, k) + M.U(ii, j, k));
        vs[m] = 0.5 * (M.V(ii - 1, j, k) + M.V(ii, j, k));
    }
    if (N == 1) return zq[0];
    double bu[N], bv[N];
    Weno<N>::beta(us, bu);
    Weno<N>::beta(vs, bv);
#pragma unroll
    for (int s = 0; s < N; s++) bu[s] = 0.5 * (bu[s] + bv[s]);
    return Weno<N>::combine(zq, bu);
}
template <int N>
__device__ __forceinline__ double recon_div_x(const Mom& M, int i, int j, int k, bool left) {
    const int base = left ? i - N : i + N - 1, sgn = left ? 1 : -1;
    double q[2 * N - 1], s[2 * N - 1];
#pragma unroll
    for (int m = 0; m < 2 * N - 1; m++) {
        const int ii = base + sgn * m;
        q[m] = M.dxU(ii, j, k);
        s[m] = q[m] + M.dyV(ii, j, k);
    }
    if (N == 1) return q[0];
    double b[N];
    Weno<N>::beta(s, b);
    return Weno<N>::combine(q, b);
}
template <int N>
__device__ __forceinline__ double recon_div_y(const Mom& M, int i, int j, int k, bool left) {
    const int base = left ? j - N : j + N - 1, sgn = left ? 1 : -1;
    double q[2 * N - 1], s[2 * N - 1];
#pragma unroll
    for (int m = 0; m < 2 * N - 1; m++) {
        const int jj = base + sgn * m;
        q[m] = M.dyV(i, jj, k);
        s[m] = M.dxU(i, jj, k) + q[m];
    }
    if (N == 1) return q[0];
    double b[N];
    Weno<N>::beta(s, b);
    return Weno<N>::combine(q, b);
}
template <int N>
__device__ __forceinline__ double recon_cell(const double* __restrict__ c, long long base, long long stride) {
    double q[2 * N - 1];
#pragma unroll
    for (int m = 0; m < 2 * N - 1; m++) q[m] = c[base + stride * m];
    if (N == 1) return q[0];
    double b[N];
    Weno<N>::beta(q, b);
    return Weno<N>::combine(q, b);
}

// runtime-order dispatch used by the general path
#define OB_DISPATCH(N, CALL1, CALL2, CALL3, CALL4, CALL5) \
    ((N) >= 5 ? (CALL5) : (N) == 4 ? (CALL4) : (N) == 3 ? (CALL3) : (N) == 2 ? (CALL2) : (CALL1))

__device__ __forceinline__ int order_cells(const Geom& G, int i, int j, int k, int Nmax, bool along_x) {
    for (int N = Nmax; N >= 1; N--) {
        bool ok = true;
        for (int m = -N; m <= N - 1 && ok; m++) ok = along_x ? active(G, i + m, j, k) : active(G, i, j + m, k);
        if (ok) return N;
    }
    return 0;
}
__device__ __forceinline__ int order_faces(const Geom& G, int i, int j, int k, int Nmax, bool along_x) {
    for (int N = Nmax; N >= 1; N--) {
        bool ok = true;
        for (int m = -N + 1; m <= N && ok; m++) ok = along_x ? !inactive_fcc(G, i + m, j, k) : !inactive_cfc(G, i, j + m, k);
        if (ok) return N;
    }
    return 0;
}

// ================================================================================================
// momentum tendencies
// ================================================================================================
template <bool GEN>
__device__ __forceinline__ double tend_u(const Geom& G, const Fields& F, int i, int j, int k, int NV, int ND, double time) {
    const Mom M{G, F.u, F.v};
    const long long c = idx3(G, i, j, k);
    // (1) vorticity flux
    const double q00 = G.dxcf[j] * F.v[c - 1], q01 = G.dxcf[j + 1] * F.v[c - 1 + G.pitch];
    const double q10 = G.dxcf[j] * F.v[c], q11 = G.dxcf[j + 1] * F.v[c + G.pitch];
    const double vavg = 0.5 * (0.5 * (q00 + q01) + 0.5 * (q10 + q11));
    const double vhat = vavg / G.dxfc[j];
    double Gt = 0.0;
    if (vhat != 0.0) {
        const bool left = vhat > 0.0;
        double z;
        if (!GEN) z = recon_vort_y<5, false>(M, i, j, k, left);
        else {
            const int N = order_faces(G, i, j, k, NV, false);
            z = OB_DISPATCH(N, (recon_vort_y<1, true>(M, i, j, k, left)), (recon_vort_y<2, true>(M, i, j, k, left)),
                            (recon_vort_y<3, true>(M, i, j, k, left)), (recon_vort_y<4, true>(M, i, j, k, left)),
                            (recon_vort_y<5, true>(M, i, j, k, left)));
        }
        Gt = vhat * z;
    }
    // (2) divergence flux + vertical advection
    const double uh = F.u[c];
    double Phi = 0.0;
    if (uh != 0.0) {
        const bool left = uh > 0.0;
        double d;
        if (!GEN) d = recon_div_x<3>(M, i, j, k, left);
        else {
            const int N = order_cells(G, i, j, k, ND, true);
            d = OB_DISPATCH(N, (recon_div_x<1>(M, i, j, k, left)), (recon_div_x<2>(M, i, j, k, left)),
                            (recon_div_x<3>(M, i, j, k, left)), (recon_div_x<3>(M, i, j, k, left)),
                            (recon_div_x<3>(M, i, j, k, left)));
        }
        Phi = uh * d + uh * 0.5 * (M.dyV(i - 1, j, k) + M.dyV(i, j, k));
    }
    const double az = G.azcc[j];
    const int km = max(k - 1, 0), kp = min(k + 1, G.Nz - 1);
    double fl_lo, fl_hi;
    {
        const double Wlo = 0.5 * (az * F.w[c - 1] + az * F.w[c]);
        const double Whi = 0.5 * (az * F.w[c - 1 + G.sz] + az * F.w[c + G.sz]);
        fl_lo = Wlo * 0.5 * (F.u[idx3(G, i, j, km)] + uh);
        fl_hi = Whi * 0.5 * (uh + F.u[idx3(G, i, j, kp)]);
        if (GEN) {
            // immersed-peripheral (f,c,f) faces carry no flux: the node below is peripheral but inside the domain
            if (k > 0 && peripheral_fcc(G, i, j, k - 1)) fl_lo = 0.0;
            // (the node above an active node is active for a grid-fitted bottom; the top face is a domain boundary)
        }
    }
    Gt -= (Phi + fl_hi - fl_lo) / (az * G.dzc[k]);
    // (3) Bernoulli head  (4) Coriolis  (5) pressure gradient
    Gt -= (M.Kh(i, j, k) - M.Kh(i - 1, j, k)) / G.dxfc[j];
    int nact = 4;
    if (GEN) {
        nact = 0;
        for (int a = 0; a < 2; a++)
            for (int b = 0; b < 2; b++) nact += peripheral_cfc(G, i - 1 + a, j + b, k) ? 0 : 1;
    }
    const double fbar = 0.5 * (G.fff[j] + G.fff[j + 1]);
    const double cor = nact == 0 ? 0.0 : vavg / (0.25 * nact);
    Gt += fbar * cor / G.dxfc[j];
    Gt -= (F.p[c] - F.p[c - 1]) / G.dxfc[j];
    // (6) QG-Leith stress divergence
    if (G.qgleith) {
        const double ax = G.dy * G.dzc[k];
        auto fccc = [&](int a, int b) { return active(G, a, b, k) ? -F.nuL[idx3(G, a, b, k)] * ((M.dxU(a, b, k) / G.dzc[k] + M.dyV(a, b, k) / G.dzc[k]) / G.azcc[b]) : 0.0; };
        auto fffc = [&](int a, int b) {
            const bool per = !active(G, a - 1, b - 1, k) || !active(G, a, b - 1, k) || !active(G, a - 1, b, k) || !active(G, a, b, k);
            if (per && in_domain(G, b - 1, k) && in_domain(G, b, k)) return 0.0;
            const long long q = idx3(G, a, b, k);
            const double nu = 0.25 * (F.nuL[q - 1 - G.pitch] + F.nuL[q - G.pitch] + F.nuL[q - 1] + F.nuL[q]);
            return nu * M.zeta<true>(a, b, k);
        };
        const double t = ax * fccc(i, j) - ax * fccc(i - 1, j) +
                         (G.dxcf[j + 1] * G.dzc[k] * fffc(i, j + 1) - G.dxcf[j] * G.dzc[k] * fffc(i, j));
        Gt -= t / (az * G.dzc[k]);
    }
    // (7) flux boundary conditions
    if (k == G.Nz - 1) {
        double J = 0.0;
        if (G.bc_kind == OB200_BC_DOUBLEDRAKE) J = G.taux[j];
        else if (G.bc_kind == OB200_BC_REALISTIC) J = interp_flux(G, F.forcing[0], i, j, G.Ny, time, false);
        Gt -= J / G.dzc[k];
    }
    if (k == 0) {
        double J = 0.0;
        if (G.bc_kind == OB200_BC_DOUBLEDRAKE) J = -G.mu_drag * uh;
        else if (G.bc_kind == OB200_BC_REALISTIC) {
            const double v0 = F.v[c - 1], v1 = F.v[c - 1 + G.pitch], v2 = F.v[c], v3 = F.v[c + G.pitch];
            J = -G.mu_drag * uh * sqrt(uh * uh + 0.25 * (v0 * v0 + v1 * v1 + v2 * v2 + v3 * v3));
        }
        Gt += J / G.dzc[k];
    }
    return Gt;
}

template <bool GEN>
__device__ __forceinline__ double tend_v(const Geom& G, const Fields& F, int i, int j, int k, int NV, int ND, double time) {
    const Mom M{G, F.u, F.v};
    const long long c = idx3(G, i, j, k);
    const double dy = G.dy;
    const double q00 = dy * F.u[c - G.pitch], q10 = dy * F.u[c + 1 - G.pitch];
    const double q01 = dy * F.u[c], q11 = dy * F.u[c + 1];
    const double uavg = 0.5 * (0.5 * (q00 + q10) + 0.5 * (q01 + q11));
    const double uhat = uavg / dy;
    double Gt = 0.0;
    if (uhat != 0.0) {
        const bool left = uhat > 0.0;
        double z;
        if (!GEN) z = recon_vort_x<5, false>(M, i, j, k, left);
        else {
            const int N = order_faces(G, i, j, k, NV, true);
            z = OB_DISPATCH(N, (recon_vort_x<1, true>(M, i, j, k, left)), (recon_vort_x<2, true>(M, i, j, k, left)),
                            (recon_vort_x<3, true>(M, i, j, k, left)), (recon_vort_x<4, true>(M, i, j, k, left)),
                            (recon_vort_x<5, true>(M, i, j, k, left)));
        }
        Gt = -(uhat * z);
    }
    const double vh = F.v[c];
    double Phi = 0.0;
    if (vh != 0.0) {
        const bool left = vh > 0.0;
        double d;
        if (!GEN) d = recon_div_y<3>(M, i, j, k, left);
        else {
            const int N = order_cells(G, i, j, k, ND, false);
            d = OB_DISPATCH(N, (recon_div_y<1>(M, i, j, k, left)), (recon_div_y<2>(M, i, j, k, left)),
                            (recon_div_y<3>(M, i, j, k, left)), (recon_div_y<3>(M, i, j, k, left)),
                            (recon_div_y<3>(M, i, j, k, left)));
        }
        Phi = vh * d + vh * 0.5 * (M.dxU(i, j - 1, k) + M.dxU(i, j, k));
    }
    const int km = max(k - 1, 0), kp = min(k + 1, G.Nz - 1);
    const double Wlo = 0.5 * (G.azcc[j - 1] * F.w[c - G.pitch] + G.azcc[j] * F.w[c]);
    const double Whi = 0.5 * (G.azcc[j - 1] * F.w[c - G.pitch + G.sz] + G.azcc[j] * F.w[c + G.sz]);
    double fl_lo = Wlo * 0.5 * (F.v[idx3(G, i, j, km)] + vh);
    const double fl_hi = Whi * 0.5 * (vh + F.v[idx3(G, i, j, kp)]);
    if (GEN) {
        if (k > 0 && peripheral_cfc(G, i, j, k - 1)) fl_lo = 0.0;
    }
    Gt -= (Phi + fl_hi - fl_lo) / (G.azcf[j] * G.dzc[k]);
    Gt -= (M.Kh(i, j, k) - M.Kh(i, j - 1, k)) / dy;
    int nact = 4;
    if (GEN) {
        nact = 0;
        for (int a = 0; a < 2; a++)
            for (int b = 0; b < 2; b++) nact += peripheral_fcc(G, i + a, j - 1 + b, k) ? 0 : 1;
    }
    const double fbar = 0.5 * (G.fff[j] + G.fff[j]);
    const double cor = nact == 0 ? 0.0 : uavg / (0.25 * nact);
    Gt -= fbar * cor / dy;
    Gt -= (F.p[c] - F.p[c - G.pitch]) / dy;
    if (G.qgleith) {
        const double ax = dy * G.dzc[k];
        auto fccc = [&](int a, int b) { return active(G, a, b, k) ? -F.nuL[idx3(G, a, b, k)] * ((M.dxU(a, b, k) / G.dzc[k] + M.dyV(a, b, k) / G.dzc[k]) / G.azcc[b]) : 0.0; };
        auto fffc = [&](int a, int b) {
            const bool per = !active(G, a - 1, b - 1, k) || !active(G, a, b - 1, k) || !active(G, a - 1, b, k) || !active(G, a, b, k);
            if (per && in_domain(G, b - 1, k) && in_domain(G, b, k)) return 0.0;
            const long long q = idx3(G, a, b, k);
            const double nu = 0.25 * (F.nuL[q - 1 - G.pitch] + F.nuL[q - G.pitch] + F.nuL[q - 1] + F.nuL[q]);
            return nu * M.zeta<true>(a, b, k);
        };
        const double t = ax * (-fffc(i + 1, j)) - ax * (-fffc(i, j)) +
                         (G.dxfc[j] * G.dzc[k] * fccc(i, j) - G.dxfc[j - 1] * G.dzc[k] * fccc(i, j - 1));
        Gt -= t / (G.azcf[j] * G.dzc[k]);
    }
    if (k == G.Nz - 1 && G.bc_kind == OB200_BC_REALISTIC) Gt -= interp_flux(G, F.forcing[1], i, j, G.Ny + 1, time, false) / G.dzc[k];
    if (k == 0) {
        double J = 0.0;
        if (G.bc_kind == OB200_BC_DOUBLEDRAKE) J = -G.mu_drag * vh;
        else if (G.bc_kind == OB200_BC_REALISTIC) {
            const double u0 = F.u[c - G.pitch], u1 = F.u[c + 1 - G.pitch], u2 = F.u[c], u3 = F.u[c + 1];
            J = -G.mu_drag * vh * sqrt(vh * vh + 0.25 * (u0 * u0 + u2 * u2 + u1 * u1 + u3 * u3));
        }
        Gt += J / G.dzc[k];
    }
    return Gt;
}

__global__ void __launch_bounds__(128) k_momentum(Geom G, Fields F, int NV, int ND, double time) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= G.Nx || j >= G.Ny) return;
    const long long c = idx3(G, i, j, k);
    const int c2 = idx2(G, i, j);
    const bool fast = k >= __ldg(&G.kfast[c2]);
    double gu, gv;
    if (fast) {
        gu = tend_u<false>(G, F, i, j, k, NV, ND, time);
        gv = tend_v<false>(G, F, i, j, k, NV, ND, time);
    } else {
        gu = peripheral_fcc(G, i, j, k) ? 0.0 : tend_u<true>(G, F, i, j, k, NV, ND, time);
        gv = peripheral_cfc(G, i, j, k) ? 0.0 : tend_v<true>(G, F, i, j, k, NV, ND, time);
    }
    F.Gu[c] = gu;
    F.Gv[c] = gv;
}

// ================================================================================================
// tracer tendencies (T and S share the velocity loads)
// ================================================================================================
template <bool GEN>
__device__ __forceinline__ double flux_x(const Geom& G, const double* __restrict__ cf, const double* __restrict__ u,
                                         int i, int j, int k, int NT) {
    const long long c = idx3(G, i, j, k);
    const double vel = u[c];
    if (vel == 0.0) return 0.0;
    const bool left = vel > 0.0;
    double r;
    if (!GEN) r = recon_cell<4>(cf, left ? c - 4 : c + 3, left ? 1 : -1);
    else {
        if (peripheral_fcc(G, i, j, k)) return 0.0;
        const int N = order_cells(G, i, j, k, NT, true);
        const long long b = left ? c - N : c + N - 1, s = left ? 1 : -1;
        r = OB_DISPATCH(N, recon_cell<1>(cf, b, s), recon_cell<2>(cf, b, s), recon_cell<3>(cf, b, s), recon_cell<4>(cf, b, s),
                        recon_cell<4>(cf, b, s));
    }
    return G.dy * G.dzc[k] * (vel * r);
}
template <bool GEN>
__device__ __forceinline__ double flux_y(const Geom& G, const double* __restrict__ cf, const double* __restrict__ v,
                                         int i, int j, int k, int NT) {
    const long long c = idx3(G, i, j, k);
    const double vel = v[c];
    if (vel == 0.0) return 0.0;
    const bool left = vel > 0.0;
    double r;
    if (!GEN) r = recon_cell<4>(cf, left ? c - 4 * (long long)G.pitch : c + 3 * (long long)G.pitch, left ? G.pitch : -G.pitch);
    else {
        if (peripheral_cfc(G, i, j, k)) return 0.0;
        const int N = order_cells(G, i, j, k, NT, false);
        const long long s = left ? G.pitch : -G.pitch;
        const long long b = left ? c - (long long)N * G.pitch : c + (long long)(N - 1) * G.pitch;
        r = OB_DISPATCH(N, recon_cell<1>(cf, b, s), recon_cell<2>(cf, b, s), recon_cell<3>(cf, b, s), recon_cell<4>(cf, b, s),
                        recon_cell<4>(cf, b, s));
    }
    return G.dxcf[j] * G.dzc[k] * (vel * r);
}

template <bool GEN>
__device__ __forceinline__ double tend_c(const Geom& G, const Fields& F, const double* __restrict__ cf, int i, int j, int k,
                                         int NT, double topflux) {
    const long long c = idx3(G, i, j, k);
    const double fxw = flux_x<GEN>(G, cf, F.u, i, j, k, NT), fxe = flux_x<GEN>(G, cf, F.u, i + 1, j, k, NT);
    const double fys = flux_y<GEN>(G, cf, F.v, i, j, k, NT), fyn = flux_y<GEN>(G, cf, F.v, i, j + 1, k, NT);
    const int km = max(k - 1, 0), kp = min(k + 1, G.Nz - 1);
    const double ck = cf[c];
    double fzl = G.azcc[j] * F.w[c] * 0.5 * (cf[idx3(G, i, j, km)] + ck);
    const double fzh = G.azcc[j] * F.w[c + G.sz] * 0.5 * (ck + cf[idx3(G, i, j, kp)]);
    if (GEN) {
        if (k > 0 && !active(G, i, j, k - 1)) fzl = 0.0;
    }
    double Gt = -((fxe - fxw) + (fyn - fys) + (fzh - fzl)) / (G.azcc[j] * G.dzc[k]);
    if (k == G.Nz - 1) Gt -= topflux / G.dzc[k];
    return Gt;
}

__global__ void __launch_bounds__(128) k_tracers(Geom G, Fields F, int NT, double time) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= G.Nx || j >= G.Ny) return;
    const long long c = idx3(G, i, j, k);
    const int c2 = idx2(G, i, j);
    const bool fast = k >= __ldg(&G.kfast[c2]);
    double jT = 0.0, jS = 0.0;
    if (k == G.Nz - 1) {
        jT = top_flux_T(G, F, i, j, F.T[c], time);
        jS = top_flux_S(G, F, i, j, F.S[c], time);
    }
    double gT = 0.0, gS = 0.0;
    if (fast) {
        gT = tend_c<false>(G, F, F.T, i, j, k, NT, jT);
        gS = tend_c<false>(G, F, F.S, i, j, k, NT, jS);
    } else if (active(G, i, j, k)) {
        gT = tend_c<true>(G, F, F.T, i, j, k, NT, jT);
        gS = tend_c<true>(G, F, F.S, i, j, k, NT, jS);
    }
    F.GT[c] = gT;
    F.GS[c] = gS;
}

void launch_momentum(const Geom& G, const Fields& F, int ord_vort, int ord_div, double time, cudaStream_t st) {
    dim3 b(32, 4), g((G.Nx + 31) / 32, (G.Ny + 3) / 4, G.Nz);
    k_momentum<<<g, b, 0, st>>>(G, F, (ord_vort + 1) / 2, (ord_div + 1) / 2, time);
    OB_COUNT_LAUNCH(1);
}
void launch_tracers(const Geom& G, const Fields& F, int ord_tracer, double time, cudaStream_t st) {
    dim3 b(32, 4), g((G.Nx + 31) / 32, (G.Ny + 3) / 4, G.Nz);
    k_tracers<<<g, b, 0, st>>>(G, F, (ord_tracer + 1) / 2, time);
    OB_COUNT_LAUNCH(1);
}
