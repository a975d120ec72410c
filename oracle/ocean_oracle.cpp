// =============================================================================
// oracle/ocean_oracle.cpp -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A plain, unfused, kernel-by-kernel CPU restatement of the hot path that
// OceanScalingTests.jl benchmarks: Oceananigans' HydrostaticFreeSurfaceModel
// `time_step!` as configured at /root/reference/src/near_global_simulation.jl:36-113
// (call sites :181,:196).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library.  The product
// (libocean_b200.so) never links, loads or calls it.
//
// PARITY UNPINNED: the arithmetic lives in the third-party dependency
// Oceananigans.jl v0.90.1 (Manifest.toml:643-649, repo-rev ss/no-immersed-cells2),
// which is absent from /root/reference and cannot run here (no Julia).  The
// reference's own tests pin no numbers (test/runtests.jl:11-19,33-34).  The
// numerics below follow SURVEY.md Appendix A (recollection of v0.90.1) plus the
// in-repo configuration files cited per function.  Deviations chosen for
// definiteness are listed in DESIGN.md ("Spec decisions").
//
// TWO ARITHMETIC MODES (option "reference_arithmetic"):
//   0 (default) PRODUCT arithmetic: the few places where the CUDA product deliberately departs from the reference's
//     operation sequence by <= 1 ulp per operation -- reciprocal metrics, division-free Z-weights, explicit fma in the
//     WENO forms, a +,-,*,/ tanh, the corrector's column sum accumulated top-down -- are evaluated the same way here,
//     so that GPU-vs-oracle regression tests can demand BIT equality.
//   1 REFERENCE arithmetic (FROZEN -- do not edit to follow the product): what Oceananigans v0.90.1 evaluates as far
//     as SURVEY Appendix A records it: a true division for every metric, alpha_s = C_s (1 + (tau / (beta_s + eps))^2)
//     with its N + N divisions, plain multiply-add chains (no fma), libm tanh, kappa / dz / dz in the solver,
//     bottom-up column sums.  tests/test_parity_gpu.py reports the GPU's relative L-infinity distance to THIS mode
//     after 100 steps against the north-star bar of 1e-10.
//
// Structure mirrors the reference's launch sequence: one function
// per Oceananigans kernel, full halo'd arrays, both upwind sides evaluated,
// runtime-order WENO, explicit masking -- deliberately NOT the fused/tiled
// structure of the CUDA product, so that the two are independent restatements.
// =============================================================================
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../include/ocean_b200.h"

// Arithmetic type: Float64, or Float32 when built with -DOB200_F32 (libocean_oracle_f32.so, the checker of
// libocean_b200_f32.so).  As in the product, every floating literal of the model arithmetic is written RL(x); host-side
// setup (metrics, bathymetry, forcing polynomials), the model clock and the C interface stay Float64.
#ifdef OB200_F32
typedef float real;
#else
typedef double real; /*f64*/
#endif
#define RL(x) ((real)(x))
#include "weno_tables.h"

namespace {

constexpr int H = OB200_HALO;  // horizontal halo (grid_load_balance.jl:25)
constexpr int HZ = 1;          // vertical halo kept by the oracle (2nd-order vertical ops only)
constexpr real WENO_EPS = RL(1e-8);  // SURVEY A8 (Oceananigans `const eps = 1e-8`) [OCN-recall]

// Deterministic tanh built from +,-,*,/ only (no libm), so that CPU and GPU agree bit for bit: the Ri-based
// diffusivity switches (N2 < 0 ...) amplify one-ulp differences of tanh into O(1e-9) differences of kappa
// within 100 steps.  |error| < 3e-16 absolute; Oceananigans itself calls libm tanh (DESIGN.md "Arithmetic contract").
static inline real det_exp_neg(real y) {  // exp(y) for y <= 0
    const real LOG2E = RL(1.4426950408889634), LN2_HI = RL(6.93147180369123816490e-01), LN2_LO = RL(1.90821492927058770002e-10);
    const real n = std::nearbyint(y * LOG2E);
    const real r = (y - n * LN2_HI) - n * LN2_LO;
    real p = RL(1.0) / RL(6227020800.0);  // 1/13!
    const real inv_fact[13] = {RL(1.0) / RL(479001600.0), RL(1.0) / RL(39916800.0), RL(1.0) / RL(3628800.0), RL(1.0) / RL(362880.0), RL(1.0) / RL(40320.0),
                                 RL(1.0) / RL(5040.0), RL(1.0) / RL(720.0), RL(1.0) / RL(120.0), RL(1.0) / RL(24.0), RL(1.0) / RL(6.0), RL(0.5), RL(1.0), RL(1.0)};
    for (int m = 0; m < 13; m++) p = p * r + inv_fact[m];
    return std::ldexp(p, (int)n);
}
static inline real det_tanh(real x) {
    const real a = std::fabs(x);
    if (a > RL(20.0)) return x > 0 ? RL(1.0) : -RL(1.0);
    const real t = det_exp_neg(-RL(2.0) * a);
    const real r = (RL(1.0) - t) / (RL(1.0) + t);
    return x < 0 ? -r : r;
}

struct F3 {  // 3-D array with halos; (i,j,k) -> i fastest
    int NxP = 0, NyP = 0, NzP = 0;
    std::vector<real> a;
    void alloc(int Nx, int Ny, int Nz) {
        NxP = Nx + 2 * H; NyP = Ny + 1 + 2 * H; NzP = Nz + 1 + 2 * HZ;
        a.assign((size_t)NxP * NyP * NzP, RL(0.0));
    }
    inline real& operator()(int i, int j, int k) {
        return a[((size_t)(k + HZ) * NyP + (j + H)) * NxP + (i + H)];
    }
    inline real operator()(int i, int j, int k) const {
        return a[((size_t)(k + HZ) * NyP + (j + H)) * NxP + (i + H)];
    }
};
struct F2 {
    int NxP = 0, NyP = 0;
    std::vector<real> a;
    void alloc(int Nx, int Ny) { NxP = Nx + 2 * H; NyP = Ny + 1 + 2 * H; a.assign((size_t)NxP * NyP, RL(0.0)); }
    inline real& operator()(int i, int j) { return a[(size_t)(j + H) * NxP + (i + H)]; }
    inline real operator()(int i, int j) const { return a[(size_t)(j + H) * NxP + (i + H)]; }
};

// ---- TEOS-10 55-term polynomial (Roquet et al. 2015) as used by SeawaterPolynomials 0.3.2
// (Manifest.toml:844-847) [OCN-recall: coefficients from memory, unverifiable here].
// R[i][j][k]: power of s, tau, zeta.
struct Teos {
    real R[7][7][4];
    Teos() {
        std::memset(R, 0, sizeof(R));
        R[0][0][0] = RL(8.0189615746e+02); R[1][0][0] = RL(8.6672408165e+02); R[2][0][0] = -RL(1.7864682637e+03);
        R[3][0][0] = RL(2.0375295546e+03); R[4][0][0] = -RL(1.2849161071e+03); R[5][0][0] = RL(4.3227585684e+02);
        R[6][0][0] = -RL(6.0579916612e+01); R[0][1][0] = RL(2.6010145068e+01); R[1][1][0] = -RL(6.5281885265e+01);
        R[2][1][0] = RL(8.1770425108e+01); R[3][1][0] = -RL(5.6888046321e+01); R[4][1][0] = RL(1.7681814114e+01);
        R[5][1][0] = -RL(1.9193502195e+00); R[0][2][0] = -RL(3.7074170417e+01); R[1][2][0] = RL(6.1548258127e+01);
        R[2][2][0] = -RL(6.0362551501e+01); R[3][2][0] = RL(2.9130021253e+01); R[4][2][0] = -RL(5.4723692739e+00);
        R[0][3][0] = RL(2.1661789529e+01); R[1][3][0] = -RL(3.3449108469e+01); R[2][3][0] = RL(1.9717078466e+01);
        R[3][3][0] = -RL(3.1742946532e+00); R[0][4][0] = -RL(8.3627885467e+00); R[1][4][0] = RL(1.1311538584e+01);
        R[2][4][0] = -RL(5.3563304045e+00); R[0][5][0] = RL(5.4048723791e-01); R[1][5][0] = RL(4.8169980163e-01);
        R[0][6][0] = -RL(1.9083568888e-01);
        R[0][0][1] = RL(1.9681925209e+01); R[1][0][1] = -RL(4.2549998214e+01); R[2][0][1] = RL(5.0774768218e+01);
        R[3][0][1] = -RL(3.0938076334e+01); R[4][0][1] = RL(6.6051753097e+00); R[0][1][1] = -RL(1.3336301113e+01);
        R[1][1][1] = -RL(4.4870114575e+00); R[2][1][1] = RL(5.0042598061e+00); R[3][1][1] = -RL(6.5399043664e-01);
        R[0][2][1] = RL(6.7080479603e+00); R[1][2][1] = RL(3.5063081279e+00); R[2][2][1] = -RL(1.8795372996e+00);
        R[0][3][1] = -RL(2.4649669534e+00); R[1][3][1] = -RL(5.5077101279e-01); R[0][4][1] = RL(5.5927935970e-01);
        R[0][0][2] = RL(2.0660924175e+00); R[1][0][2] = -RL(4.9527603989e+00); R[2][0][2] = RL(2.5019633244e+00);
        R[0][1][2] = RL(2.0564311499e+00); R[1][1][2] = -RL(2.1311365518e-01); R[0][2][2] = -RL(1.2419983026e+00);
        R[0][0][3] = -RL(2.3342758797e-02); R[1][0][3] = -RL(1.8507636718e-02); R[0][1][3] = RL(3.7969820455e-01);
    }
    static constexpr real Sau = RL(40.0) * RL(35.16504) / RL(35.0), Thu = RL(40.0), Zu = RL(1e4), dS = RL(32.0);
    // returns r'(tau,s,zeta); optionally d/dTheta and d/dSA.  Nested Horner over the triangular support
    // (powers of zeta outermost), the same evaluation order as the product so the two agree bit for bit.
    real rprime(real Th, real SA, real Z, real* dTh = nullptr, real* dSA = nullptr) const {
        const real tau = Th / Thu, s = std::sqrt((SA + dS) / Sau), ze = -Z / Zu;
        const int deg[4] = {6, 4, 2, 1};
        real r = RL(0.0), rt = RL(0.0), rs = RL(0.0);
        for (int k = 3; k >= 0; k--) {
            real pj = RL(0.0), pjt = RL(0.0), pjs = RL(0.0);
            for (int j = deg[k]; j >= 0; j--) {
                real pi = RL(0.0), pis = RL(0.0);
                for (int i = deg[k] - j; i >= 0; i--) {
                    pis = pis * s + pi;
                    pi = pi * s + R[i][j][k];
                }
                pjt = pjt * tau + pj; pjs = pjs * tau + pis;
                pj = pj * tau + pi;
            }
            r = r * ze + pj;
            rt = rt * ze + pjt; rs = rs * ze + pjs;
        }
        if (dTh) *dTh = rt / Thu;
        if (dSA) *dSA = rs / (RL(2.0) * s * Sau);
        return r;
    }
};

struct Oracle {
    ob200_config c;
    int Nx, Ny, Nz;
    bool periodic_local;
    std::vector<double> zf, weights; /*f64*/
    std::vector<real> forcing[6];
    // host-side metrics are computed and kept in Float64; the accessors below hand them to the model arithmetic rounded to
    // `real` -- exactly what the product uploads (ocean_b200.cu upload_real)
    // k-indexed (offset +1 so that k = -1 .. Nz are valid)
    std::vector<double> zc_, dzc_, dzf_; /*f64*/
    // j-indexed with halo offset H: j = -H .. Ny+H
    std::vector<double> phif_, phic_, dxfc_, dxcf_, azcc_, azcf_, fff_; /*f64*/
    // reciprocals (1.0 / x, correctly rounded) used by the tendency operators: a multiplication by the
    // precomputed reciprocal replaces each division by a grid metric (<= 1 ulp from the reference's division)
    std::vector<double> rdxfc_, razcc_, razcf_, rdzc_, rdzf_, rupz_, rloz_; /*f64*/
    real rdy = 0, r1cav = 0, rRidelta = 0;
    real dy = 0;
    std::vector<int> kb_;  // (NyP x NxP) first active level per column, Nz = none
    int NxP, NyP;
    Teos teos;
    F3 u, v, w, T, S, p, kc, ku, Ri, Gu, Gv, GT, GS, Gu0, Gv0, GT0, GS0;
    F3 nuL, qx, qy;  // QG-Leith
    F2 eta, U, V, etab, Ub, Vb, GU, GV, Hfc, Hcf, Ld;
    double time = 0, last_dt = INFINITY; /*f64*/
    // the configuration's real-valued parameters in the model's arithmetic type
    struct RealParams {
        real gravity, eos_alpha, eos_beta, eos_rho0, nu_z, kappa_z, ri_nu0, ri_kappa0, ri_kappa_ca, ri_Cen, ri_Cav, ri_Ri0,
            ri_Ridelta, T_restoring_lambda, drag_mu, leith_C, leith_min_N2, leith_Vscale;
    } p_;
    int64_t iteration = 0;
    std::string err;
    bool refar = false;   // reference arithmetic (see the header)
    int refbits = 7;   // bisect mask: 1 metrics, 2 WENO, 4 tanh
    // Timing runs only (bench.py's CPU legs): evaluate just the upwind-side reconstruction.  0.5 ((a + |a|) L + (a - |a|) R) is a L or
    // a R exactly, so the fields are bit-identical (tests/test_oracle_cpu.py); the default evaluates both sides, as the reference does.
    bool upwind1 = false;
    inline bool skip_side(int side, real vel) const { return upwind1 && (side == 0 ? !(vel > RL(0.0)) : !(vel < RL(0.0))); }
    // x / metric: a true division (reference) or a multiplication by the correctly rounded reciprocal (product)
    inline real mdiv(real x, real m, real rm) const { return (refar && (refbits & 1)) ? x / m : x * rm; }
    // 1 / V = 1 / (Az dz): Oceananigans evaluates `1 / V * (...)`
    inline real rvol(real az, real raz, int k) const { return (refar && (refbits & 1)) ? RL(1.0) / (az * dzc(k)) : raz * rdzc(k); }
    inline real tanh_(real x) const { return (refar && (refbits & 4)) ? std::tanh(x) : det_tanh(x); }

    inline real zc(int k) const { return (real)zc_[k + 1]; }
    inline real dzc(int k) const { return (real)dzc_[k + 1]; }
    inline real dzf(int k) const { return (real)dzf_[k]; }  // k = 0..Nz
    inline double phif(int j) const { return phif_[j + H]; } /*f64*/
    inline double phic(int j) const { return phic_[j + H]; } /*f64*/
    inline real dxfc(int j) const { return (real)dxfc_[j + H]; }
    inline real dxcf(int j) const { return (real)dxcf_[j + H]; }
    inline real azcc(int j) const { return (real)azcc_[j + H]; }
    inline real azcf(int j) const { return (real)azcf_[j + H]; }
    inline real fff(int j) const { return (real)fff_[j + H]; }
    inline real rdxfc(int j) const { return (real)rdxfc_[j + H]; }
    inline real razcc(int j) const { return (real)razcc_[j + H]; }
    inline real razcf(int j) const { return (real)razcf_[j + H]; }
    inline real rdzc(int k) const { return (real)rdzc_[k + 1]; }
    inline real rdzf(int k) const { return (real)rdzf_[k]; }
    inline real rupz(int k) const { return (real)rupz_[k]; }
    inline real rloz(int k) const { return (real)rloz_[k]; }
    inline int kb(int i, int j) const { return kb_[(size_t)(j + H) * NxP + (i + H)]; }

    // ---- node predicates ---------------------------------------
    inline bool in_domain(int j, int k) const { return j >= 0 && j < Ny && k >= 0 && k < Nz; }
    inline bool active(int i, int j, int k) const { return in_domain(j, k) && k >= kb(i, j); }
    inline bool immersed(int i, int j, int k) const { return in_domain(j, k) && k < kb(i, j); }
    inline bool inactive_fcc(int i, int j, int k) const { return !active(i - 1, j, k) && !active(i, j, k); }
    inline bool inactive_cfc(int i, int j, int k) const { return !active(i, j - 1, k) && !active(i, j, k); }
    inline bool peripheral_fcc(int i, int j, int k) const { return !active(i - 1, j, k) || !active(i, j, k); }
    inline bool peripheral_cfc(int i, int j, int k) const { return !active(i, j - 1, k) || !active(i, j, k); }
    inline bool peripheral_ccf(int i, int j, int k) const { return !active(i, j, k - 1) || !active(i, j, k); }
    inline bool peripheral_fcf(int i, int j, int k) const {
        return !active(i - 1, j, k - 1) || !active(i, j, k - 1) || !active(i - 1, j, k) || !active(i, j, k);
    }
    inline bool peripheral_cff(int i, int j, int k) const {
        return !active(i, j - 1, k - 1) || !active(i, j, k - 1) || !active(i, j - 1, k) || !active(i, j, k);
    }
    // immersed_peripheral = peripheral on the immersed grid but not on the underlying grid
    inline bool imm_peripheral_fcc(int i, int j, int k) const { return peripheral_fcc(i, j, k) && in_domain(j, k); }
    inline bool imm_peripheral_cfc(int i, int j, int k) const {
        return peripheral_cfc(i, j, k) && in_domain(j - 1, k) && in_domain(j, k);
    }
    inline bool imm_peripheral_ccf(int i, int j, int k) const {
        return peripheral_ccf(i, j, k) && in_domain(j, k - 1) && in_domain(j, k);
    }
    inline bool imm_peripheral_fcf(int i, int j, int k) const {
        return peripheral_fcf(i, j, k) && in_domain(j, k - 1) && in_domain(j, k);
    }
    inline bool imm_peripheral_cff(int i, int j, int k) const {
        return peripheral_cff(i, j, k) && in_domain(j - 1, k - 1) && in_domain(j, k);
    }

    // ---- setup (host side: Float64 throughout, as in ocean_b200.cu setup_geometry) ------------
    /*f64-begin*/
    int setup(const ob200_config* cfg) {
        c = *cfg;
        Nx = c.Nx; Ny = c.Ny; Nz = c.Nz;
        if (Nx < 1 || Ny < 2 || Nz < 2 || c.n_substeps < 1 || c.n_substeps > OB200_MAX_SUBSTEPS) {
            err = "bad sizes"; return OB200_EINVAL;
        }
        periodic_local = (c.nranks <= 1);
        NxP = Nx + 2 * H; NyP = Ny + 1 + 2 * H;
        zf.assign(c.z_faces, c.z_faces + Nz + 1);
        weights.assign(c.averaging_weights, c.averaging_weights + c.n_substeps);
        // vertical grid: halo spacings repeat the adjacent interior spacing
        zc_.assign(Nz + 2, 0); dzc_.assign(Nz + 2, 0); dzf_.assign(Nz + 1, 0);
        for (int k = 0; k < Nz; k++) { dzc_[k + 1] = zf[k + 1] - zf[k]; zc_[k + 1] = 0.5 * (zf[k + 1] + zf[k]); }
        dzc_[0] = dzc_[1]; dzc_[Nz + 1] = dzc_[Nz];
        zc_[0] = zf[0] - 0.5 * dzc_[0]; zc_[Nz + 1] = zf[Nz] + 0.5 * dzc_[Nz + 1];
        for (int k = 0; k <= Nz; k++) dzf_[k] = zc_[k + 1] - zc_[k];  // zc(k) - zc(k-1)
        // horizontal grid, metrics per latitude row
        const double pi = M_PI;
        double dlam = (c.lon_east - c.lon_west) / c.Nx_global;
        double dphi = (c.lat_north - c.lat_south) / Ny;
        const double dlam_rad = dlam * pi / 180.0;
        const double dy_d = c.radius * (dphi * pi / 180.0);
        dy = (real)dy_d;
        int nj = Ny + 1 + 2 * H + 1;
        phif_.assign(nj, 0); phic_.assign(nj, 0); dxfc_.assign(nj, 0); dxcf_.assign(nj, 0);
        azcc_.assign(nj, 0); azcf_.assign(nj, 0); fff_.assign(nj, 0);
        auto pf = [&](int j) { return (double)((long double)c.lat_south + (long double)j * (long double)dphi); };
        auto pc = [&](int j) { return (double)((long double)c.lat_south + ((long double)j + 0.5L) * (long double)dphi); };
        auto cosd = [&](double x) { return std::cos(pi * x / 180.0); };
        auto sind = [&](double x) { return std::sin(pi * x / 180.0); };
        for (int jj = 0; jj < nj; jj++) {
            int j = jj - H;
            phif_[jj] = pf(j); phic_[jj] = pc(j);
            dxfc_[jj] = c.radius * cosd(pc(j)) * dlam_rad;
            dxcf_[jj] = c.radius * cosd(pf(j)) * dlam_rad;
            azcc_[jj] = c.radius * c.radius * dlam_rad * (sind(pf(j + 1)) - sind(pf(j)));
            azcf_[jj] = c.radius * c.radius * dlam_rad * (sind(pc(j)) - sind(pc(j - 1)));
            fff_[jj] = 2.0 * c.rotation_rate * sind(pf(j));
        }
        rdxfc_.resize(nj); razcc_.resize(nj); razcf_.resize(nj); rdzc_.resize(Nz + 2);
        for (int jj = 0; jj < nj; jj++) { rdxfc_[jj] = 1.0 / dxfc_[jj]; razcc_[jj] = 1.0 / azcc_[jj]; razcf_[jj] = 1.0 / azcf_[jj]; }
        for (int k = 0; k < Nz + 2; k++) rdzc_[k] = 1.0 / dzc_[k];
        rdy = (real)(1.0 / dy_d);
        rdzf_.resize(Nz + 1); rupz_.assign(Nz + 1, 0.0); rloz_.assign(Nz + 1, 0.0);
        for (int k = 0; k <= Nz; k++) rdzf_[k] = 1.0 / dzf_[k];
        for (int k = 0; k < Nz; k++) { rupz_[k] = 1.0 / (dzc_[k + 1] * dzf_[k + 1]); rloz_[k] = 1.0 / (dzc_[k + 1] * dzf_[k]); }
        r1cav = (real)(1.0 / (1.0 + c.ri_Cav)); rRidelta = (real)(1.0 / c.ri_Ridelta);
        p_.gravity = (real)c.gravity; p_.eos_alpha = (real)c.eos_alpha; p_.eos_beta = (real)c.eos_beta; p_.eos_rho0 = (real)c.eos_rho0;
        p_.nu_z = (real)c.nu_z; p_.kappa_z = (real)c.kappa_z; p_.ri_nu0 = (real)c.ri_nu0; p_.ri_kappa0 = (real)c.ri_kappa0;
        p_.ri_kappa_ca = (real)c.ri_kappa_ca; p_.ri_Cen = (real)c.ri_Cen; p_.ri_Cav = (real)c.ri_Cav; p_.ri_Ri0 = (real)c.ri_Ri0;
        p_.ri_Ridelta = (real)c.ri_Ridelta; p_.T_restoring_lambda = (real)c.T_restoring_lambda; p_.drag_mu = (real)c.drag_mu;
        p_.leith_C = (real)c.leith_C; p_.leith_min_N2 = (real)c.leith_min_N2; p_.leith_Vscale = (real)c.leith_Vscale;
        // bathymetry -> first active level (GridFittedBottom, centre condition: zc <= h is solid)
        kb_.assign((size_t)NxP * NyP, Nz);
        const int bh = c.bottom_halo_columns, nxh = Nx + 2 * bh;   // the host may carry a wider halo than H
        if (bh < H) { err = "bottom_halo_columns < 7"; return OB200_EINVAL; }
        for (int j = 0; j < Ny; j++)
            for (int ii = 0; ii < NxP; ii++) {
                double hgt = c.bottom_height[(size_t)j * nxh + (ii - H + bh)];
                int k0 = 0;
                while (k0 < Nz && zc_[k0 + 1] <= hgt) k0++;
                kb_[(size_t)(j + H) * NxP + ii] = k0;
            }
        for (F3* f : {&u, &v, &w, &T, &S, &p, &kc, &ku, &Ri, &Gu, &Gv, &GT, &GS, &Gu0, &Gv0, &GT0, &GS0, &nuL, &qx, &qy})
            f->alloc(Nx, Ny, Nz);
        for (F2* f : {&eta, &U, &V, &etab, &Ub, &Vb, &GU, &GV, &Hfc, &Hcf, &Ld}) f->alloc(Nx, Ny);
        // column depths at u / v points: sum of dz over non-peripheral nodes
        for (int j = -H + 1; j < Ny + H; j++)
            for (int i = -H + 1; i < Nx + H; i++) {
                double a = 0, b = 0;
                for (int k = 0; k < Nz; k++) {
                    if (!peripheral_fcc(i, j, k)) a += dzc_[k + 1];
                    if (!peripheral_cfc(i, j, k)) b += dzc_[k + 1];
                }
                Hfc(i, j) = (real)a; Hcf(i, j) = (real)b;
            }
        return OB200_OK;
    }
    /*f64-end*/

    // ---- equation of state --------------------------------------------------
    inline real buoyancy(int i, int j, int k) const {  // buoyancy perturbation at cell centre
        if (c.eos_kind == OB200_EOS_LINEAR) return p_.gravity * (p_.eos_alpha * T(i, j, k) - p_.eos_beta * S(i, j, k));
        return -p_.gravity * teos.rprime(T(i, j, k), S(i, j, k), zc(k)) / p_.eos_rho0;
    }
    // N^2 = d b / dz at the (c,c,f) face k.  Oceananigans' `∂z_b` for SeawaterBuoyancy is
    //     g (thermal_expansionᶜᶜᶠ ∂zᶜᶜᶠ(T) − haline_contractionᶜᶜᶠ ∂zᶜᶜᶠ(S)),
    // the coefficients evaluated AT THE FACE from the z-averaged T, S and the face depth [OCN-recall] -- not a difference
    // of in-situ buoyancies, which for TEOS-10 would add the compressibility term ∂ρ'/∂z at fixed T, S (1e-6 .. 4e-6 s^-2).
    // ∂z is the immersed conditional difference: zero when either cell is inactive.
    inline real dz_b(int i, int j, int k) const {
        if (!active(i, j, k - 1) || !active(i, j, k)) return RL(0.0);
        real al = p_.eos_alpha, be = p_.eos_beta;
        if (c.eos_kind == OB200_EOS_TEOS10) {
            real dTh, dSA;
            teos.rprime(RL(0.5) * (T(i, j, k - 1) + T(i, j, k)), RL(0.5) * (S(i, j, k - 1) + S(i, j, k)), (real)zf[k], &dTh, &dSA);
            al = -dTh / p_.eos_rho0; be = dSA / p_.eos_rho0;
        }
        const real dTz = mdiv(T(i, j, k) - T(i, j, k - 1), dzf(k), rdzf(k));
        const real dSz = mdiv(S(i, j, k) - S(i, j, k - 1), dzf(k), rdzf(k));
        return p_.gravity * (al * dTz - be * dSz);
    }

    // ---- boundary-condition functions (bathymetry_and_forcings.jl:110-159) ----
    /*f64-begin*/  // per-latitude tables in the product (computed on the host in Float64, stored rounded)
    inline double wind_stress(double phi) const {
        const double* q = phi < -45 ? c.wind_coeffs : phi < -15 ? c.wind_coeffs + 4 : phi < 0 ? c.wind_coeffs + 8
                        : phi < 15 ? c.wind_coeffs + 12 : phi < 45 ? c.wind_coeffs + 16 : c.wind_coeffs + 20;
        return q[0] * phi * phi * phi + q[1] * phi * phi + q[2] * phi + q[3];
    }
    inline double salinity_flux(double phi) const {
        const double* q = phi < -20 ? c.salt_coeffs : phi < 0 ? c.salt_coeffs + 4 : phi < 20 ? c.salt_coeffs + 8
                                                                                          : c.salt_coeffs + 12;
        return q[0] * phi * phi * phi + q[1] * phi * phi + q[2] * phi + q[3];
    }
    static inline double T_reference(double phi) { return std::max(0.0, 30.0 * std::cos(1.2 * M_PI * phi / 180.0)); }
    /*f64-end*/
    // flux_from_interpolated_array (bathymetry_and_forcings.jl:162-169)
    inline real interp_flux(const std::vector<real>& F, int i, int j, int nyf, real period_days, bool div3) const {
        if (F.empty() || i < 0 || i >= Nx) return RL(RL(0.0));
        // the interpolation index comes from the Float64 model clock in both precisions (ob_bc.cuh interp_flux)
        /*f64-begin*/
        double days = time / 86400.0;
        double n;
        if (!div3) n = std::fmod(days, period_days) + 1;
        else n = std::floor(std::fmod(days, 15.0) / 3.0) + 1;   // `mod(t,15) ÷ 3 + 1` (:180)
        int n1 = (int)std::floor(n), n2 = n1 + 1;
        /*f64-end*/
        size_t plane = (size_t)Nx * nyf;
        real a = F[(size_t)(n1 - 1) * plane + (size_t)j * Nx + i], b = F[(size_t)(n2 - 1) * plane + (size_t)j * Nx + i];
        return a * (real)(n2 - n) + b * (real)(n - n1);
    }
    inline real top_flux_T(int i, int j) const {
        switch (c.bc_kind) {
            case OB200_BC_DOUBLEDRAKE: return p_.T_restoring_lambda * (T(i, j, Nz - 1) - (real)T_reference(phic(j)));
            case OB200_BC_REALISTIC: {
                real fl = interp_flux(forcing[2], i, j, Ny, RL(5.0), false);
                if (!forcing[4].empty())
                    fl += (dzmin() / (RL(30.0) * RL(86400.0))) * (T(i, j, Nz - 1) - interp_flux(forcing[4], i, j, Ny, RL(15.0), true));
                return fl;
            }
            default: return RL(0.0);
        }
    }
    inline real top_flux_S(int i, int j) const {
        switch (c.bc_kind) {
            case OB200_BC_DOUBLEDRAKE: return (real)salinity_flux(phic(j));
            case OB200_BC_REALISTIC: {
                real fl = interp_flux(forcing[3], i, j, Ny, RL(5.0), false);
                if (!forcing[5].empty())
                    fl += (dzmin() / (RL(60.0) * RL(86400.0))) * (S(i, j, Nz - 1) - interp_flux(forcing[5], i, j, Ny, RL(15.0), true));
                return fl;
            }
            default: return RL(0.0);
        }
    }
    inline real dzmin() const { real m = RL(1e300); for (int k = 0; k < Nz; k++) m = std::min(m, dzc(k)); return m; }
    inline real top_buoyancy_flux(int i, int j) const {
        if (c.bc_kind == OB200_BC_QUIESCENT) return RL(0.0);
        real al = p_.eos_alpha, be = p_.eos_beta;
        if (c.eos_kind == OB200_EOS_TEOS10) {
            real dT, dSv;
            teos.rprime(T(i, j, Nz - 1), S(i, j, Nz - 1), zc(Nz - 1), &dT, &dSv);
            al = -dT / p_.eos_rho0; be = dSv / p_.eos_rho0;
        }
        return p_.gravity * (al * top_flux_T(i, j) - be * top_flux_S(i, j));
    }

    // ---- halos & masks ----------------------------------------
    void fill_x(F3& f, int nk) {
        if (!periodic_local) return;
        for (int k = -HZ; k < nk + HZ; k++)
            for (int j = -H; j < Ny + 1 + H; j++)
                for (int h = 0; h < H; h++) { f(-1 - h, j, k) = f(Nx - 1 - h, j, k); f(Nx + h, j, k) = f(h, j, k); }
    }
    void fill_x(F2& f) {
        if (!periodic_local) return;
        for (int j = -H; j < Ny + 1 + H; j++)
            for (int h = 0; h < H; h++) { f(-1 - h, j) = f(Nx - 1 - h, j); f(Nx + h, j) = f(h, j); }
    }
    void fill_y_center(F3& f, int nk) {  // zero-gradient (no-flux) copy into every halo row
        for (int k = 0; k < nk; k++)
            for (int i = -H; i < Nx + H; i++)
                for (int h = 1; h <= H; h++) { f(i, -h, k) = f(i, 0, k); f(i, Ny - 1 + h, k) = f(i, Ny - 1, k); }
    }
    void fill_y_center(F2& f) {
        for (int i = -H; i < Nx + H; i++)
            for (int h = 1; h <= H; h++) { f(i, -h) = f(i, 0); f(i, Ny - 1 + h) = f(i, Ny - 1); }
    }
    void fill_y_face(F3& f, int nk) {  // impenetrable walls: v = 0 on and beyond the wall faces
        for (int k = 0; k < nk; k++)
            for (int i = -H; i < Nx + H; i++) {
                for (int h = 0; h <= H; h++) f(i, -h, k) = RL(0.0);
                for (int h = 0; h <= H; h++) f(i, Ny + h, k) = RL(0.0);
            }
    }
    void fill_z(F3& f) {
        for (int j = -H; j < Ny + 1 + H; j++)
            for (int i = -H; i < Nx + H; i++) { f(i, j, -1) = f(i, j, 0); f(i, j, Nz) = f(i, j, Nz - 1); }
    }
    void mask_fields() {  // mask_immersed_field! on u,v,T,S; eta with inactive_node at the top face
        for (int k = 0; k < Nz; k++)
            for (int j = 0; j <= Ny; j++)
                for (int i = 0; i < Nx; i++) {
                    if (j < Ny) {
                        if (peripheral_fcc(i, j, k)) u(i, j, k) = RL(0.0);
                        if (!active(i, j, k)) { T(i, j, k) = RL(0.0); S(i, j, k) = RL(0.0); }
                    }
                    if (peripheral_cfc(i, j, k)) v(i, j, k) = RL(0.0);
                }
        for (int j = 0; j < Ny; j++)
            for (int i = 0; i < Nx; i++)
                if (!active(i, j, Nz - 1)) eta(i, j) = RL(0.0);
    }
    void local_halos() {
        fill_x(u, Nz); fill_x(v, Nz); fill_x(T, Nz); fill_x(S, Nz); fill_x(eta);
        local_halos_yz();
    }
    void local_halos_yz() {
        fill_y_center(u, Nz); fill_y_center(T, Nz); fill_y_center(S, Nz); fill_y_center(eta);
        fill_y_face(v, Nz);
        fill_z(u); fill_z(v); fill_z(T); fill_z(S);
    }

    // ---- _compute_w_from_continuity! --------------------------------
    inline real div_xy(int i, int j, int k) const {
        return mdiv(dy * u(i + 1, j, k) - dy * u(i, j, k) + dxcf(j + 1) * v(i, j + 1, k) - dxcf(j) * v(i, j, k), azcc(j), razcc(j));
    }
    void compute_w() {
#pragma omp parallel for
        for (int j = -H + 1; j < Ny + H - 1; j++)
            for (int i = -H + 1; i < Nx + H - 1; i++) {
                w(i, j, 0) = RL(0.0);
                for (int k = 1; k <= Nz; k++) w(i, j, k) = w(i, j, k - 1) - dzc(k - 1) * div_xy(i, j, k - 1);
            }
    }
    // ---- _update_hydrostatic_pressure! --------------------------------
    void compute_p() {
#pragma omp parallel for
        for (int j = -1; j < Ny + 1; j++)
            for (int i = -1; i < Nx + 1; i++) {
                p(i, j, Nz - 1) = -RL(0.5) * (buoyancy(i, j, Nz) + buoyancy(i, j, Nz - 1)) * dzf(Nz);
                for (int k = Nz - 2; k >= 0; k--)
                    p(i, j, k) = p(i, j, k + 1) - RL(0.5) * (buoyancy(i, j, k + 1) + buoyancy(i, j, k)) * dzf(k + 1);
            }
    }
    // ---- RiBasedVerticalDiffusivity -----------------------------------
    inline real dz_u(int i, int j, int k) const {  // d/dz at (f,c,f), conditional on both u-nodes being active
        if (inactive_fcc(i, j, k) || inactive_fcc(i, j, k - 1)) return RL(0.0);
        return mdiv(u(i, j, k) - u(i, j, k - 1), dzf(k), rdzf(k));
    }
    inline real dz_v(int i, int j, int k) const {
        if (inactive_cfc(i, j, k) || inactive_cfc(i, j, k - 1)) return RL(0.0);
        return mdiv(v(i, j, k) - v(i, j, k - 1), dzf(k), rdzf(k));
    }
    void compute_ri_number() {
#pragma omp parallel for
        for (int k = 0; k < Nz; k++)
            for (int j = 0; j < Ny; j++)
                for (int i = -2; i < Nx + 2; i++) {
                    real a0 = dz_u(i, j, k), a1 = dz_u(i + 1, j, k), b0 = dz_v(i, j, k), b1 = dz_v(i, j + 1, k);
                    real S2 = RL(0.5) * (a0 * a0 + a1 * a1) + RL(0.5) * (b0 * b0 + b1 * b1);
                    real N2 = dz_b(i, j, k);
                    Ri(i, j, k) = (N2 <= RL(0.0)) ? RL(0.0) : N2 / S2;
                }
        fill_y_center(Ri, Nz);
    }
    void compute_ri_diffusivities() {
        if (!c.ri_enabled) return;
        compute_ri_number();
#pragma omp parallel for
        for (int j = -1; j < Ny + 1; j++)
            for (int i = -1; i < Nx + 1; i++) {
                bool col_in = (j >= 0 && j < Ny);
                real Qb = col_in ? top_buoyancy_flux(i, j) : RL(0.0);
                for (int k = 0; k < Nz; k++) {
                    real kcp = RL(0.0), kup = RL(0.0);
                    if (!peripheral_ccf(i, j, k)) {
                        real N2 = dz_b(i, j, k), N2a = dz_b(i, j, k + 1);
                        bool convecting = N2 < RL(0.0);
                        bool entraining = (N2 > RL(0.0)) && (N2a < RL(0.0)) && (Qb > RL(0.0));
                        real kca = convecting ? p_.ri_kappa_ca : RL(0.0);
                        real ken = entraining ? p_.ri_Cen * Qb / N2 : RL(0.0);
                        auto ff = [&](int a, int b) {  // Ixy^ff(Ri)
                            return RL(0.25) * (Ri(a - 1, b - 1, k) + Ri(a, b - 1, k) + Ri(a - 1, b, k) + Ri(a, b, k));
                        };
                        real Rif = RL(0.25) * (ff(i, j) + ff(i + 1, j) + ff(i, j + 1) + ff(i + 1, j + 1));
                        real tau = RL(0.5) * (RL(1.0) - tanh_(mdiv(Rif - p_.ri_Ri0, p_.ri_Ridelta, rRidelta)));
                        kcp = p_.ri_kappa0 * tau + kca + ken;
                        kup = p_.ri_nu0 * tau;
                    }
                    kc(i, j, k) = mdiv(p_.ri_Cav * kc(i, j, k) + kcp, RL(1.0) + p_.ri_Cav, r1cav);
                    ku(i, j, k) = mdiv(p_.ri_Cav * ku(i, j, k) + kup, RL(1.0) + p_.ri_Cav, r1cav);
                }
            }
    }

    // ---- WENO machinery --------------------------------------------
    // q[0..2N-2]: stencil values ordered upwind-ward mirrored so that the same routine
    // serves left- and right-biased reconstructions.  sa/sb: fields the smoothness is
    // measured on (VelocityStencil: two fields averaged; otherwise sb == nullptr).
    // reference arithmetic: the Horner-like sum psi_a * (C_aa psi_a + C_ab psi_b + ...) of Oceananigans' `smoothness_sum`
    // with plain multiplications and additions, left to right [OCN-recall]
    static real beta_of_ref(int N, int s, const real* ps) {
        const real* C = WENO_SMOOTH[N][s];
        real b = RL(0.0); int m = 0;
        for (int a = 0; a < N; a++) {
            real inner = C[m++] * ps[a];
            for (int bb = a + 1; bb < N; bb++) inner = inner + C[m++] * ps[bb];
            b = (a == 0) ? ps[a] * inner : b + ps[a] * inner;
        }
        return b;
    }
    static real weno_ref(int N, const real* q, const real* sa, const real* sb) {
        if (N == 1) return q[0];
        real beta[5], alpha[5], asum = RL(0.0);
        for (int s = 0; s < N; s++) {
            int off = N - 1 - s;
            beta[s] = beta_of_ref(N, s, sa + off);
            if (sb) beta[s] = (beta[s] + beta_of_ref(N, s, sb + off)) / 2;
        }
        real tau;
        switch (N) {
            case 2: tau = std::fabs(beta[0] - beta[1]); break;
            case 3: tau = std::fabs(beta[0] - beta[2]); break;
            case 4: tau = std::fabs(beta[0] + 3 * beta[1] - 3 * beta[2] - beta[3]); break;
            default: tau = std::fabs(beta[0] + 2 * beta[1] - 6 * beta[2] + 2 * beta[3] + beta[4]); break;
        }
        // alpha_s = C_s (1 + (tau / (beta_s + eps))^2);  w = alpha / sum(alpha);  result = sum(w_s p_s)
        for (int s = 0; s < N; s++) {
            const real r = tau / (beta[s] + WENO_EPS);
            alpha[s] = WENO_OPT[N][s] * (RL(1.0) + r * r);
            asum = (s == 0) ? alpha[s] : asum + alpha[s];
        }
        real res = RL(0.0);
        for (int s = 0; s < N; s++) {
            int off = N - 1 - s;
            real ps = WENO_COEFF[N][s][0] * q[off];
            for (int a = 1; a < N; a++) ps = ps + WENO_COEFF[N][s][a] * q[off + a];
            const real w = alpha[s] / asum;
            res = (s == 0) ? w * ps : res + w * ps;
        }
        return res;
    }
    static real beta_of(int N, int s, const real* ps) {
        const real* C = WENO_SMOOTH[N][s];
        // explicit fma in a fixed order (see DESIGN.md "Arithmetic contract"): the quadratic form cancels
        // catastrophically on fields with a large mean, so the evaluation order is part of the contract
        real b = RL(0.0); int m = 0;
        for (int a = 0; a < N; a++) {
            real inner = RL(0.0);
            for (int bb = a; bb < N; bb++) inner = std::fma(C[m++], ps[bb], inner);
            b = std::fma(ps[a], inner, b);
        }
        return b;
    }
    static real weno(int N, const real* q, const real* sa, const real* sb, bool ref_arith = false) {
        if (N == 1) return q[0];  // UpwindBiased(order = 1)
        if (ref_arith) return weno_ref(N, q, sa, sb);
        real beta[5], alpha[5], asum = RL(0.0), r = RL(0.0);
        for (int s = 0; s < N; s++) {
            int off = N - 1 - s;
            beta[s] = beta_of(N, s, sa + off);
            if (sb) beta[s] = RL(0.5) * (beta[s] + beta_of(N, s, sb + off));
        }
        real tau;
        switch (N) {
            case 2: tau = std::fabs(beta[0] - beta[1]); break;
            case 3: tau = std::fabs(beta[0] - beta[2]); break;
            case 4: tau = std::fabs(beta[0] + 3 * beta[1] - 3 * beta[2] - beta[3]); break;
            default: tau = std::fabs(beta[0] + 2 * beta[1] - 6 * beta[2] + 2 * beta[3] + beta[4]); break;
        }
        // Z-weights alpha_s = C_s (1 + (tau / d_s)^2), d_s = beta_s + eps, evaluated division-free after scaling
        // by P^2, P = prod d_r:  alpha_s P^2 = C_s (P^2 + (tau P_s)^2) with P_s = prod_{r != s} d_r (prefix * suffix).
        // Algebraically identical to the reference's formula; one division per reconstruction instead of N + 1.
#ifdef OB200_F32
        // Float32 build (weno_gen.cuh, same branch): the reference's own form alpha_s = C_s (1 + (tau / (beta_s + eps))^2) --
        // the product of all (beta_s + eps) below underflows in Float32 -- with the denominator kept >= 1e-12 in magnitude,
        // sign preserved (beta_s + eps cancels to exactly zero about once per step at 3e7 cells in Float32)
        for (int s = 0; s < N; s++) {
            const real dd = beta[s] + WENO_EPS;
            const real t = tau / std::copysign(std::max(std::fabs(dd), RL(1e-12)), dd);
            alpha[s] = WENO_OPT[N][s] * std::fma(t, t, RL(1.0));
            asum += alpha[s];
            int off = N - 1 - s;
            real ps = RL(0.0);
            for (int a = 0; a < N; a++) ps = std::fma(WENO_COEFF[N][s][a], q[off + a], ps);
            r = std::fma(alpha[s], ps, r);
        }
        return r / asum;
#endif
        real d[5], L[5], R[5];
        for (int s = 0; s < N; s++) d[s] = beta[s] + WENO_EPS;
        L[0] = RL(1.0); for (int s = 0; s < N - 1; s++) L[s + 1] = L[s] * d[s];
        R[N - 1] = RL(1.0); for (int s = N - 1; s > 0; s--) R[s - 1] = R[s] * d[s];
        const real Pt = L[N - 1] * d[N - 1], PP = Pt * Pt;
        for (int s = 0; s < N; s++) {
            real t = tau * (L[s] * R[s]);
            alpha[s] = WENO_OPT[N][s] * std::fma(t, t, PP);
            asum += alpha[s];
            int off = N - 1 - s;
            real ps = RL(0.0);
            for (int a = 0; a < N; a++) ps = std::fma(WENO_COEFF[N][s][a], q[off + a], ps);
            r = std::fma(alpha[s], ps, r);
        }
        return r / asum;
    }
    static inline real upwind(real vel, real L, real R) {
        return RL(0.5) * ((vel + std::fabs(vel)) * L + (vel - std::fabs(vel)) * R);
    }

    // order selection: largest N whose full (L u R) stencil contains no inactive node
    inline int order_cells_x(int i, int j, int k, int Nmax) const {
        for (int N = Nmax; N >= 1; N--) {
            bool ok = true;
            for (int m = i - N; m <= i + N - 1 && ok; m++) ok = active(m, j, k);
            if (ok) return N;
        }
        return 0;
    }
    inline int order_cells_y(int i, int j, int k, int Nmax) const {
        for (int N = Nmax; N >= 1; N--) {
            bool ok = true;
            for (int m = j - N; m <= j + N - 1 && ok; m++) ok = active(i, m, k);
            if (ok) return N;
        }
        return 0;
    }
    inline int order_faces_y(int i, int j, int k, int Nmax) const {  // zeta(ffc) -> fcc, nodes checked at (c,f,c)
        for (int N = Nmax; N >= 1; N--) {
            bool ok = true;
            for (int m = j - N + 1; m <= j + N && ok; m++) ok = !inactive_cfc(i, m, k);
            if (ok) return N;
        }
        return 0;
    }
    inline int order_faces_x(int i, int j, int k, int Nmax) const {  // zeta(ffc) -> cfc, nodes checked at (f,c,c)
        for (int N = Nmax; N >= 1; N--) {
            bool ok = true;
            for (int m = i - N + 1; m <= i + N && ok; m++) ok = !inactive_fcc(m, j, k);
            if (ok) return N;
        }
        return 0;
    }

    // ---- operators used by the momentum tendencies ------------------------------------
    inline real zeta(int i, int j, int k) const {  // vertical vorticity at (f,f,c), conditional differences
        real dxv = (inactive_cfc(i - 1, j, k) || inactive_cfc(i, j, k)) ? RL(0.0) : dy * v(i, j, k) - dy * v(i - 1, j, k);
        real dyu = (inactive_fcc(i, j - 1, k) || inactive_fcc(i, j, k)) ? RL(0.0)
                                                                          : dxfc(j) * u(i, j, k) - dxfc(j - 1) * u(i, j - 1, k);
        return mdiv(dxv - dyu, azcf(j), razcf(j));
    }
    inline real dxU(int i, int j, int k) const { return dy * dzc(k) * u(i + 1, j, k) - dy * dzc(k) * u(i, j, k); }
    inline real dyV(int i, int j, int k) const {
        return dxcf(j + 1) * dzc(k) * v(i, j + 1, k) - dxcf(j) * dzc(k) * v(i, j, k);
    }
    inline real Kh(int i, int j, int k) const {
        real a = u(i, j, k), b = u(i + 1, j, k), cc = v(i, j, k), d = v(i, j + 1, k);
        return (RL(0.5) * (a * a + b * b) + RL(0.5) * (cc * cc + d * d)) / RL(2.0);
    }

    real tendency_u(int i, int j, int k) const {
        // (1) vorticity flux: - upwind(vhat, zetaL, zetaR), WENO(order_vorticity) in y, VelocityStencil
        real qv[2][2];
        for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) qv[a][b] = dxcf(j + b) * v(i - 1 + a, j + b, k);
        real vhat = mdiv(RL(0.5) * (RL(0.5) * (qv[0][0] + qv[0][1]) + RL(0.5) * (qv[1][0] + qv[1][1])), dxfc(j), rdxfc(j));
        int N = order_faces_y(i, j, k, (c.order_vorticity + 1) / 2);
        real zq[9], us[9], vs[9], zL, zR;
        for (int side = 0; side < 2; side++) {
            if (skip_side(side, vhat)) { (side == 0 ? zL : zR) = RL(0.0); continue; }
            for (int m = 0; m < 2 * N - 1; m++) {
                int jj = side == 0 ? j - N + 1 + m : j + N - m;
                zq[m] = zeta(i, jj, k);
                us[m] = RL(0.5) * (u(i, jj - 1, k) + u(i, jj, k));
                vs[m] = RL(0.5) * (v(i - 1, jj, k) + v(i, jj, k));
            }
            (side == 0 ? zL : zR) = weno(N, zq, us, vs, refar && (refbits & 2));
        }
        real G = upwind(vhat, zL, zR);
        // (2) vertical advection + upwinded horizontal-divergence flux (OnlySelfUpwinding, cross = Centered)
        int Nd = order_cells_x(i, j, k, (c.order_divergence + 1) / 2);
        real dq[5], ds[5], dL, dR;
        for (int side = 0; side < 2; side++) {
            if (skip_side(side, u(i, j, k))) { (side == 0 ? dL : dR) = RL(0.0); continue; }
            for (int m = 0; m < 2 * Nd - 1; m++) {
                int ii = side == 0 ? i - Nd + m : i + Nd - 1 - m;
                dq[m] = dxU(ii, j, k);
                ds[m] = dxU(ii, j, k) + dyV(ii, j, k);
            }
            (side == 0 ? dL : dR) = weno(Nd, dq, ds, nullptr, refar && (refbits & 2));
        }
        real uh = u(i, j, k);
        real Phi = upwind(uh, dL, dR) + uh * RL(0.5) * (dyV(i - 1, j, k) + dyV(i, j, k));
        auto wflux = [&](int kf) {
            if (imm_peripheral_fcf(i, j, kf)) return RL(0.0);
            real W = RL(0.5) * (azcc(j) * w(i - 1, j, kf) + azcc(j) * w(i, j, kf));
            return W * RL(0.5) * (u(i, j, kf - 1) + u(i, j, kf));
        };
        G -= (Phi + wflux(k + 1) - wflux(k)) * rvol(azcc(j), razcc(j), k);
        // (3) Bernoulli head, centred
        G -= mdiv(Kh(i, j, k) - Kh(i - 1, j, k), dxfc(j), rdxfc(j));
        // (4) Coriolis, ActiveCellEnstrophyConserving: x_f_cross_U = -f * <dx v>_active / dx
        int nact = 0;
        for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) nact += peripheral_cfc(i - 1 + a, j + b, k) ? 0 : 1;
        real fbar = RL(0.5) * (fff(j) + fff(j + 1));
        real vavg = RL(0.5) * (RL(0.5) * (qv[0][0] + qv[0][1]) + RL(0.5) * (qv[1][0] + qv[1][1]));
        real cor = nact == 0 ? RL(0.0) : nact == 4 ? vavg : vavg / (RL(0.25) * nact);
        G += mdiv(fbar * cor, dxfc(j), rdxfc(j));
        // (5) hydrostatic pressure gradient
        G -= mdiv(p(i, j, k) - p(i - 1, j, k), dxfc(j), rdxfc(j));
        // (6) optional QG-Leith horizontal viscosity, vector-invariant form
        if (c.qgleith_enabled) G -= leith_div_tau_u(i, j, k);
        // (7) flux boundary conditions (_apply_z_bcs!)
        if (k == Nz - 1) G -= mdiv(top_flux_u(i, j), dzc(k), rdzc(k));
        if (k == 0) G += mdiv(bottom_flux_u(i, j), dzc(k), rdzc(k));
        // ImmersedBoundaryCondition(bottom = u_immersed_quadratic_bottom_drag) (boundary_and_initial_conditions.jl:124-134,
        // bathymetry_and_forcings.jl:138): the same drag law where the sea floor lies above level 0, i.e. on the first
        // non-peripheral node above a peripheral one [OCN-recall for where the immersed flux is applied]
        if (c.bc_kind == OB200_BC_REALISTIC && k > 0 && peripheral_fcc(i, j, k - 1))
            G += mdiv(-p_.drag_mu * u(i, j, k) * speed_fcc(i, j, k), dzc(k), rdzc(k));
        return G;
    }

    real tendency_v(int i, int j, int k) const {
        real qu[2][2];
        for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) qu[a][b] = dy * u(i + a, j - 1 + b, k);
        real uhat = mdiv(RL(0.5) * (RL(0.5) * (qu[0][0] + qu[1][0]) + RL(0.5) * (qu[0][1] + qu[1][1])), dy, rdy);
        int N = order_faces_x(i, j, k, (c.order_vorticity + 1) / 2);
        real zq[9], us[9], vs[9], zL, zR;
        for (int side = 0; side < 2; side++) {
            if (skip_side(side, uhat)) { (side == 0 ? zL : zR) = RL(0.0); continue; }
            for (int m = 0; m < 2 * N - 1; m++) {
                int ii = side == 0 ? i - N + 1 + m : i + N - m;
                zq[m] = zeta(ii, j, k);
                us[m] = RL(0.5) * (u(ii, j - 1, k) + u(ii, j, k));
                vs[m] = RL(0.5) * (v(ii - 1, j, k) + v(ii, j, k));
            }
            (side == 0 ? zL : zR) = weno(N, zq, us, vs, refar && (refbits & 2));
        }
        real G = -upwind(uhat, zL, zR);
        int Nd = order_cells_y(i, j, k, (c.order_divergence + 1) / 2);
        real dq[5], ds[5], dL, dR;
        for (int side = 0; side < 2; side++) {
            if (skip_side(side, v(i, j, k))) { (side == 0 ? dL : dR) = RL(0.0); continue; }
            for (int m = 0; m < 2 * Nd - 1; m++) {
                int jj = side == 0 ? j - Nd + m : j + Nd - 1 - m;
                dq[m] = dyV(i, jj, k);
                ds[m] = dxU(i, jj, k) + dyV(i, jj, k);
            }
            (side == 0 ? dL : dR) = weno(Nd, dq, ds, nullptr, refar && (refbits & 2));
        }
        real vh = v(i, j, k);
        real Phi = upwind(vh, dL, dR) + vh * RL(0.5) * (dxU(i, j - 1, k) + dxU(i, j, k));
        auto wflux = [&](int kf) {
            if (imm_peripheral_cff(i, j, kf)) return RL(0.0);
            real W = RL(0.5) * (azcc(j - 1) * w(i, j - 1, kf) + azcc(j) * w(i, j, kf));
            return W * RL(0.5) * (v(i, j, kf - 1) + v(i, j, kf));
        };
        G -= (Phi + wflux(k + 1) - wflux(k)) * rvol(azcf(j), razcf(j), k);
        G -= mdiv(Kh(i, j, k) - Kh(i, j - 1, k), dy, rdy);
        int nact = 0;
        for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) nact += peripheral_fcc(i + a, j - 1 + b, k) ? 0 : 1;
        real fbar = RL(0.5) * (fff(j) + fff(j));
        real uavg = RL(0.5) * (RL(0.5) * (qu[0][0] + qu[1][0]) + RL(0.5) * (qu[0][1] + qu[1][1]));
        real cor = nact == 0 ? RL(0.0) : nact == 4 ? uavg : uavg / (RL(0.25) * nact);
        G -= mdiv(fbar * cor, dy, rdy);
        G -= mdiv(p(i, j, k) - p(i, j - 1, k), dy, rdy);
        if (c.qgleith_enabled) G -= leith_div_tau_v(i, j, k);
        if (k == Nz - 1) G -= mdiv(top_flux_v(i, j), dzc(k), rdzc(k));
        if (k == 0) G += mdiv(bottom_flux_v(i, j), dzc(k), rdzc(k));
        if (c.bc_kind == OB200_BC_REALISTIC && k > 0 && peripheral_cfc(i, j, k - 1))
            G += mdiv(-p_.drag_mu * v(i, j, k) * speed_cfc(i, j, k), dzc(k), rdzc(k));
        return G;
    }

    inline real top_flux_u(int i, int j) const {
        if (c.bc_kind == OB200_BC_DOUBLEDRAKE) return (real)wind_stress(phic(j));
        if (c.bc_kind == OB200_BC_REALISTIC) return interp_flux(forcing[0], i, j, Ny, RL(5.0), false);
        return RL(0.0);
    }
    inline real top_flux_v(int i, int j) const {
        if (c.bc_kind == OB200_BC_REALISTIC) return interp_flux(forcing[1], i, j, Ny + 1, RL(5.0), false);
        return RL(0.0);
    }
    inline real speed_fcc(int i, int j, int k) const {  // bathymetry_and_forcings.jl:131
        real s = 0; for (int a = -1; a <= 0; a++) for (int b = 0; b <= 1; b++) s += v(i + a, j + b, k) * v(i + a, j + b, k);
        return std::sqrt(u(i, j, k) * u(i, j, k) + RL(0.25) * s);
    }
    inline real speed_cfc(int i, int j, int k) const {  // :132
        real s = 0; for (int a = 0; a <= 1; a++) for (int b = -1; b <= 0; b++) s += u(i + a, j + b, k) * u(i + a, j + b, k);
        return std::sqrt(v(i, j, k) * v(i, j, k) + RL(0.25) * s);
    }
    inline real bottom_flux_u(int i, int j) const {
        if (c.bc_kind == OB200_BC_DOUBLEDRAKE) return -p_.drag_mu * u(i, j, 0);                       // :141
        if (c.bc_kind == OB200_BC_REALISTIC) return -p_.drag_mu * u(i, j, 0) * speed_fcc(i, j, 0);    // :135
        return RL(0.0);
    }
    inline real bottom_flux_v(int i, int j) const {
        if (c.bc_kind == OB200_BC_DOUBLEDRAKE) return -p_.drag_mu * v(i, j, 0);
        if (c.bc_kind == OB200_BC_REALISTIC) return -p_.drag_mu * v(i, j, 0) * speed_cfc(i, j, 0);
        return RL(0.0);
    }

    real tendency_c(const F3& cfld, int i, int j, int k, real top_flux) const {
        int Nt = (c.order_tracer + 1) / 2;
        auto flux_x = [&](int ii) {
            if (imm_peripheral_fcc(ii, j, k)) return RL(0.0);
            int N = order_cells_x(ii, j, k, Nt);
            if (N == 0) return RL(0.0);
            real q[7], L = RL(0.0), R = RL(0.0);
            if (!skip_side(0, u(ii, j, k))) {
                for (int m = 0; m < 2 * N - 1; m++) q[m] = cfld(ii - N + m, j, k);
                L = weno(N, q, q, nullptr, refar && (refbits & 2));
            }
            if (!skip_side(1, u(ii, j, k))) {
                for (int m = 0; m < 2 * N - 1; m++) q[m] = cfld(ii + N - 1 - m, j, k);
                R = weno(N, q, q, nullptr, refar && (refbits & 2));
            }
            return dy * dzc(k) * upwind(u(ii, j, k), L, R);
        };
        auto flux_y = [&](int jj) {
            if (imm_peripheral_cfc(i, jj, k)) return RL(0.0);
            int N = order_cells_y(i, jj, k, Nt);
            if (N == 0) return RL(0.0);  // wall face: v = 0
            real q[7], L = RL(0.0), R = RL(0.0);
            if (!skip_side(0, v(i, jj, k))) {
                for (int m = 0; m < 2 * N - 1; m++) q[m] = cfld(i, jj - N + m, k);
                L = weno(N, q, q, nullptr, refar && (refbits & 2));
            }
            if (!skip_side(1, v(i, jj, k))) {
                for (int m = 0; m < 2 * N - 1; m++) q[m] = cfld(i, jj + N - 1 - m, k);
                R = weno(N, q, q, nullptr, refar && (refbits & 2));
            }
            return dxcf(jj) * dzc(k) * upwind(v(i, jj, k), L, R);
        };
        auto flux_z = [&](int kf) {
            if (imm_peripheral_ccf(i, j, kf)) return RL(0.0);
            return azcc(j) * w(i, j, kf) * RL(0.5) * (cfld(i, j, kf - 1) + cfld(i, j, kf));
        };
        real G = -((flux_x(i + 1) - flux_x(i)) + (flux_y(j + 1) - flux_y(j)) + (flux_z(k + 1) - flux_z(k))) *
                   rvol(azcc(j), razcc(j), k);
        if (k == Nz - 1) G -= mdiv(top_flux, dzc(k), rdzc(k));
        return G;
    }

    void compute_tendencies() {
#pragma omp parallel for schedule(dynamic, 1)
        for (int k = 0; k < Nz; k++)
            for (int j = 0; j < Ny; j++)
                for (int i = 0; i < Nx; i++) {
                    Gu(i, j, k) = peripheral_fcc(i, j, k) ? RL(0.0) : tendency_u(i, j, k);
                    Gv(i, j, k) = peripheral_cfc(i, j, k) ? RL(0.0) : tendency_v(i, j, k);
                    if (active(i, j, k)) {
                        GT(i, j, k) = tendency_c(T, i, j, k, k == Nz - 1 ? top_flux_T(i, j) : RL(0.0));
                        GS(i, j, k) = tendency_c(S, i, j, k, k == Nz - 1 ? top_flux_S(i, j) : RL(0.0));
                    } else { GT(i, j, k) = RL(0.0); GS(i, j, k) = RL(0.0); }
                }
    }

    // ---- QG-Leith (src/quasi_geostrophic_leith.jl:37-169) --------------------------------
    // Julia min/max propagate NaN; C++ std::min/max do not -> emulate.
    static inline real jmin(real a, real b) { return (std::isnan(a) || std::isnan(b)) ? NAN : std::min(a, b); }
    static inline real jmax(real a, real b) { return (std::isnan(a) || std::isnan(b)) ? NAN : std::max(a, b); }
    inline real f_ccc(int j) const { return RL(0.25) * (fff(j) + fff(j) + fff(j + 1) + fff(j + 1)); }  // Ixy^cc(f^ff)
    inline real dx_b(int i, int j, int k) const {  // d b/dx at fcc, conditional
        if (!active(i - 1, j, k) || !active(i, j, k)) return RL(0.0);
        return (buoyancy(i, j, k) - buoyancy(i - 1, j, k)) / dxfc(j);
    }
    inline real dy_b(int i, int j, int k) const {
        if (!active(i, j - 1, k) || !active(i, j, k)) return RL(0.0);
        return (buoyancy(i, j, k) - buoyancy(i, j - 1, k)) / dy;
    }
    void compute_leith() {
        if (!c.qgleith_enabled) return;
        const int e = 2;  // extent beyond the interior needed by the stress divergence
        // calculate_deformation_radius! (:126-135)
#pragma omp parallel for
        for (int j = -e; j < Ny + e; j++)
            for (int i = -e; i < Nx + e; i++) {
                real s = RL(0.0);
                for (int k = 0; k < Nz; k++)
                    s += dzf(k) * (std::sqrt(jmax(RL(0.0), dz_b(i, j, k))) / RL(M_PI) / std::fabs(f_ccc(j)));
                Ld(i, j) = s;
            }
        // compute_stretching! (:114-121): q at (c,c,f)
#pragma omp parallel for
        for (int k = 0; k <= Nz; k++)
            for (int j = -e; j < Ny + e; j++)
                for (int i = -e; i < Nx + e; i++) {
                    real coef = f_ccc(j) / jmax(p_.leith_min_N2, dz_b(i, j, k));
                    real bx = RL(0.25) * (dx_b(i, j, k - 1) + dx_b(i + 1, j, k - 1) + dx_b(i, j, k) + dx_b(i + 1, j, k));
                    real by = RL(0.25) * (dy_b(i, j, k - 1) + dy_b(i, j + 1, k - 1) + dy_b(i, j, k) + dy_b(i, j + 1, k));
                    qx(i, j, k) = coef * bx; qy(i, j, k) = coef * by;
                }
        // calculate_qgleith_viscosity! (:78-104)
#pragma omp parallel for
        for (int k = 0; k < Nz; k++)
            for (int j = -e; j < Ny + e; j++)
                for (int i = -e; i < Nx + e; i++) {
                    auto dxz = [&](int a, int b) {  // d zeta/dx at (c,f,c), conditional on the two ffc nodes
                        return (zeta(a + 1, b, k) - zeta(a, b, k)) / dxcf(b);
                    };
                    auto dyz = [&](int a, int b) { return (zeta(a, b + 1, k) - zeta(a, b, k)) / dy; };
                    real dzx = RL(0.5) * (dxz(i, j) + dxz(i, j + 1));
                    real dzy = RL(0.5) * (dyz(i, j) + dyz(i + 1, j));
                    real dfx = RL(0.0);  // f depends on latitude only
                    real dfy = RL(0.5) * ((fff(j + 1) - fff(j)) / dy + (fff(j + 1) - fff(j)) / dy);
                    dzx += dfx; dzy += dfy;
                    real dqx = active(i, j, k) ? (qx(i, j, k + 1) - qx(i, j, k)) / dzc(k) : RL(0.0);
                    real dqy = active(i, j, k) ? (qy(i, j, k + 1) - qy(i, j, k)) / dzc(k) : RL(0.0);
                    auto ddx = [&](int a, int b) {  // d div/dx at fcc
                        if (!active(a - 1, b, k) || !active(a, b, k)) return RL(0.0);
                        return (div_xy(a, b, k) - div_xy(a - 1, b, k)) / dxfc(b);
                    };
                    auto ddy = [&](int a, int b) {
                        if (!active(a, b - 1, k) || !active(a, b, k)) return RL(0.0);
                        return (div_xy(a, b, k) - div_xy(a, b - 1, k)) / dy;
                    };
                    real ddx_ = RL(0.5) * (ddx(i, j) + ddx(i + 1, j)), ddy_ = RL(0.5) * (ddy(i, j) + ddy(i, j + 1));
                    real dd2 = ddx_ * ddx_ + ddy_ * ddy_;
                    real fc = f_ccc(j);
                    real dz2 = dzx * dzx + dzy * dzy;
                    real dq2 = (dqx + dzx) * (dqx + dzx) + (dqy + dzy) * (dqy + dzy);
                    real A = RL(2.0) * (RL(1.0) / (RL(1.0) / (dxfc(j) * dxfc(j)) + RL(1.0) / (dy * dy)));
                    real Ds = std::pow(A, RL(0.5));
                    real Bu = Ld(i, j) * Ld(i, j) / A;
                    real Ro = p_.leith_Vscale / (fc * Ds);
                    real t1 = RL(1.0) + RL(1.0) / Bu, t2 = RL(1.0) + RL(1.0) / (Ro * Ro);
                    real dQ2 = jmin(dq2, dz2 * (t1 * t1));
                    dQ2 = jmin(dQ2, dz2 * (t2 * t2));
                    nuL(i, j, k) = std::pow(p_.leith_C * Ds / RL(M_PI), RL(3.0)) * std::sqrt(dQ2 + dd2);
                }
    }
    // horizontal viscous stress divergence in vector-invariant form: nu*div at ccc, nu*zeta at ffc
    inline real nu_ffc(int i, int j, int k) const {
        return RL(0.25) * (nuL(i - 1, j - 1, k) + nuL(i, j - 1, k) + nuL(i - 1, j, k) + nuL(i, j, k));
    }
    inline real flux_ccc(int i, int j, int k) const {  // -nu * delta, zero inside solid
        return active(i, j, k) ? -nuL(i, j, k) * div_xy(i, j, k) : RL(0.0);
    }
    inline real flux_ffc(int i, int j, int k) const {  // nu_ff * zeta, zero on immersed-peripheral ffc nodes
        bool per = !active(i - 1, j - 1, k) || !active(i, j - 1, k) || !active(i - 1, j, k) || !active(i, j, k);
        bool dom = in_domain(j - 1, k) && in_domain(j, k);
        if (per && dom) return RL(0.0);
        return nu_ffc(i, j, k) * zeta(i, j, k);
    }
    real leith_div_tau_u(int i, int j, int k) const {  // (1/V)[dx(Ax * -nu delta) + dy(Ay * +nu_ff zeta)]
        real ax = dy * dzc(k);
        real t = ax * flux_ccc(i, j, k) - ax * flux_ccc(i - 1, j, k) +
                   (dxcf(j + 1) * dzc(k) * flux_ffc(i, j + 1, k) - dxcf(j) * dzc(k) * flux_ffc(i, j, k));
        return t / (azcc(j) * dzc(k));
    }
    real leith_div_tau_v(int i, int j, int k) const {  // (1/V)[dx(Ax * -nu_ff zeta) + dy(Ay * -nu delta)]
        real ax = dy * dzc(k);
        real t = ax * (-flux_ffc(i + 1, j, k)) - ax * (-flux_ffc(i, j, k)) +
                   (dxfc(j) * dzc(k) * flux_ccc(i, j, k) - dxfc(j - 1) * dzc(k) * flux_ccc(i, j - 1, k));
        return t / (azcf(j) * dzc(k));
    }

    // ---- AB2 + implicit vertical diffusion ------------------------------------
    void barotropic_forcing(real chi) {  // _compute_integrated_ab2_tendencies!
#pragma omp parallel for
        for (int j = 0; j <= Ny; j++)
            for (int i = 0; i < Nx; i++) {
                real a = 0, b = 0;
                for (int k = 0; k < Nz; k++) {
                    if (j < Ny && !peripheral_fcc(i, j, k)) a += dzc(k) * ((RL(1.5) + chi) * Gu(i, j, k) - Gu0(i, j, k) * (RL(0.5) + chi));
                    if (!peripheral_cfc(i, j, k)) b += dzc(k) * ((RL(1.5) + chi) * Gv(i, j, k) - Gv0(i, j, k) * (RL(0.5) + chi));
                }
                GU(i, j) = a; GV(i, j) = b;
            }
    }
    // solve (I + L) phi+ = phi* on one column; Kf(k) = diffusivity on the face below level k (k = 1..Nz-1)
    template <class KF>
    void implicit_column(F3& f, int i, int j, real dt, KF Kface) {
        // scratch of the column solve: per-thread, allocated once (the reference's solver owns a scratch field `t`)
        static thread_local std::vector<real> t, lo, up;
        if ((int)t.size() < Nz + 1) { t.assign(Nz + 1, RL(0.0)); lo.assign(Nz + 1, RL(0.0)); up.assign(Nz + 1, RL(0.0)); }
        for (int k = 0; k < Nz; k++) {
            // reference: -Δt * κ_Δz²  with  κ_Δz² = κ / Δzᶜ(k) / Δzᶠ(face)  (two successive divisions)
            up[k] = (k < Nz - 1) ? ((refar && (refbits & 1)) ? -dt * (Kface(k + 1) / dzc(k) / dzf(k + 1)) : -dt * Kface(k + 1) * rupz(k)) : RL(0.0);
            lo[k] = (k > 0) ? ((refar && (refbits & 1)) ? -dt * (Kface(k) / dzc(k) / dzf(k)) : -dt * Kface(k) * rloz(k)) : RL(0.0);  // coupling of row k to k-1
        }
        real beta = RL(1.0) - up[0] - lo[0];
        if (!refar) {
            // product arithmetic: ONE reciprocal per level, r_k = 1 / beta_k, t_k = up_{k-1} r_{k-1}, f'_k = (...) r_k
            real r = RL(1.0) / beta;
            f(i, j, 0) = f(i, j, 0) * r;
            for (int k = 1; k < Nz; k++) {
                t[k] = up[k - 1] * r;
                r = RL(1.0) / ((RL(1.0) - up[k] - lo[k]) - lo[k] * t[k]);
                f(i, j, k) = (f(i, j, k) - lo[k] * f(i, j, k - 1)) * r;
            }
        } else {
            // reference: solve_batched_tridiagonal_system_kernel! -- two divisions by beta per level
            f(i, j, 0) = f(i, j, 0) / beta;
            for (int k = 1; k < Nz; k++) {
                t[k] = up[k - 1] / beta;
                beta = (RL(1.0) - up[k] - lo[k]) - lo[k] * t[k];
                if (!(std::fabs(beta) > 10 * RL(2.220446049250313e-16))) break;
                f(i, j, k) = (f(i, j, k) - lo[k] * f(i, j, k - 1)) / beta;
            }
        }
        for (int k = Nz - 2; k >= 0; k--) f(i, j, k) -= t[k + 1] * f(i, j, k + 1);
    }
    void ab2_implicit(real dt, real chi) {
        auto ab2 = [&](F3& f, const F3& G, const F3& G0, int nyf) {
#pragma omp parallel for
            for (int k = 0; k < Nz; k++)
                for (int j = 0; j < nyf; j++)
                    for (int i = 0; i < Nx; i++)
                        f(i, j, k) += dt * ((RL(1.5) + chi) * G(i, j, k) - (RL(0.5) + chi) * G0(i, j, k));
        };
        ab2(u, Gu, Gu0, Ny);
#pragma omp parallel for
        for (int j = 0; j < Ny; j++)
            for (int i = 0; i < Nx; i++)
                implicit_column(u, i, j, dt, [&](int kf) {
                    if (peripheral_fcf(i, j, kf)) return RL(0.0);
                    return p_.nu_z + (c.ri_enabled ? RL(0.5) * (ku(i - 1, j, kf) + ku(i, j, kf)) : RL(0.0));
                });
        ab2(v, Gv, Gv0, Ny + 1);
#pragma omp parallel for
        for (int j = 0; j <= Ny; j++)
            for (int i = 0; i < Nx; i++)
                implicit_column(v, i, j, dt, [&](int kf) {
                    if (peripheral_cff(i, j, kf)) return RL(0.0);
                    return p_.nu_z + (c.ri_enabled ? RL(0.5) * (ku(i, j - 1, kf) + ku(i, j, kf)) : RL(0.0));
                });
        for (int tr = 0; tr < 2; tr++) {
            F3& f = tr == 0 ? T : S;
            ab2(f, tr == 0 ? GT : GS, tr == 0 ? GT0 : GS0, Ny);
#pragma omp parallel for
            for (int j = 0; j < Ny; j++)
                for (int i = 0; i < Nx; i++)
                    implicit_column(f, i, j, dt, [&](int kf) {
                        if (peripheral_ccf(i, j, kf)) return RL(0.0);
                        return p_.kappa_z + (c.ri_enabled ? kc(i, j, kf) : RL(0.0));
                    });
        }
    }

    // ---- split-explicit free surface ---------------------------------------------
    void baro_init() {
        U.a = Ub.a; V.a = Vb.a;
        std::fill(etab.a.begin(), etab.a.end(), RL(0.0));
        std::fill(Ub.a.begin(), Ub.a.end(), RL(0.0));
        std::fill(Vb.a.begin(), Vb.a.end(), RL(0.0));
    }
    // ranges: x from -ex to Nx+ex (wide-halo variant used by distributed runs)
    void baro_eta(real dtau, int ex) {
#pragma omp parallel for
        for (int j = 0; j < Ny; j++)
            for (int i = -ex; i < Nx + ex; i++) {
                real Vn = (j + 1 == Ny) ? RL(0.0) : V(i, j + 1), Vs = (j == 0) ? RL(0.0) : V(i, j);
                eta(i, j) -= dtau * ((dy * U(i + 1, j) - dy * U(i, j)) + (dxcf(j + 1) * Vn - dxcf(j) * Vs)) / azcc(j);
            }
    }
    void baro_vel(real dtau, real wgt, int ex) {
#pragma omp parallel for
        for (int j = 0; j < Ny; j++)
            for (int i = -ex; i < Nx + ex; i++) {
                real detax = (eta(i, j) - eta(i - 1, j)) / dxfc(j);
                real detay = (j == 0) ? RL(0.0) : (eta(i, j) - eta(i, j - 1)) / dy;
                U(i, j) += dtau * (-p_.gravity * Hfc(i, j) * detax + GU(i, j));
                V(i, j) += dtau * (-p_.gravity * Hcf(i, j) * detay + GV(i, j));
                etab(i, j) += wgt * eta(i, j);
                Ub(i, j) += wgt * U(i, j);
                Vb(i, j) += wgt * V(i, j);
            }
    }
    void baro_finish() { eta.a = etab.a; }
    void barotropic_loop(double dt) { /*f64*/
        baro_init();
        if (periodic_local) { fill_x(U); fill_x(V); fill_x(eta); }
        double dtau = c.frac_dtau * dt;   // formed in Float64, rounded once when it enters the kernels (as the product does) /*f64*/
        for (int m = 0; m < c.n_substeps; m++) {
            baro_eta(dtau, 0);
            if (periodic_local) fill_x(eta);
            baro_vel(dtau, weights[m], 0);
            if (periodic_local) { fill_x(U); fill_x(V); }
        }
        baro_finish();
    }

    // ---- barotropic corrector + store tendencies -----------------------------------------------
    void correct_and_store() {
#pragma omp parallel for
        for (int j = 0; j <= Ny; j++)
            for (int i = 0; i < Nx; i++) {
                real a = 0, b = 0;
                if (refar) {   // barotropic_mode_kernel!: bottom-up
                    for (int k = 0; k < Nz; k++) { if (j < Ny) a += dzc(k) * u(i, j, k); b += dzc(k) * v(i, j, k); }
                } else {       // product: accumulated in the back-substitution of the implicit solve, i.e. top-down
                    for (int k = Nz - 1; k >= 0; k--) { if (j < Ny) a += dzc(k) * u(i, j, k); b += dzc(k) * v(i, j, k); }
                }
                if (j < Ny) U(i, j) = a;
                V(i, j) = b;
                for (int k = 0; k < Nz; k++) {
                    if (j < Ny && !peripheral_fcc(i, j, k)) u(i, j, k) = u(i, j, k) + (Ub(i, j) - U(i, j)) / Hfc(i, j);
                    if (!peripheral_cfc(i, j, k)) v(i, j, k) = v(i, j, k) + (Vb(i, j) - V(i, j)) / Hcf(i, j);
                }
            }
        Gu0.a = Gu.a; Gv0.a = Gv.a; GT0.a = GT.a; GS0.a = GS.a;
    }

    // ---- sequencing -----------------------------------------------------------------
    void auxiliaries() { compute_w(); compute_ri_diffusivities(); compute_leith(); compute_p(); }
    void update_state() { mask_fields(); local_halos(); auxiliaries(); compute_tendencies(); }
    int run_stage(int st, double dt, int arg) { /*f64*/
        double chi = (dt != last_dt) ? -0.5 : c.ab2_chi; /*f64*/
        switch (st) {
            case OB200_ST_BAROTROPIC_FORCING: barotropic_forcing(chi); break;
            case OB200_ST_AB2_IMPLICIT: ab2_implicit(dt, chi); break;
            case OB200_ST_BAROTROPIC_INIT: baro_init(); break;
            case OB200_ST_BAROTROPIC_ETA: baro_eta(c.frac_dtau * dt, 0); break;
            case OB200_ST_BAROTROPIC_VEL: baro_vel(c.frac_dtau * dt, weights[arg], 0); break;
            case OB200_ST_BAROTROPIC_FINISH: baro_finish(); break;
            case OB200_ST_CORRECT_AND_STORE: correct_and_store(); break;
            case OB200_ST_MASK_AND_LOCAL_HALOS: mask_fields(); if (arg == 1) local_halos_yz(); else local_halos(); break;
            case OB200_ST_AUXILIARIES: auxiliaries(); break;
            case OB200_ST_TENDENCIES: compute_tendencies(); break;
            case OB200_ST_BAROTROPIC_LOOP: barotropic_loop(dt); break;
            default: err = "unknown stage"; return OB200_EINVAL;
        }
        return OB200_OK;
    }
    void time_step(double dt) { /*f64*/
        if (iteration == 0) update_state();
        bool euler = (dt != last_dt);
        double chi = euler ? -0.5 : c.ab2_chi; /*f64*/
        if (euler) for (F3* g : {&Gu0, &Gv0, &GT0, &GS0}) std::fill(g->a.begin(), g->a.end(), RL(0.0));
        barotropic_forcing(chi);
        ab2_implicit(dt, chi);
        barotropic_loop(dt);
        correct_and_store();
        time += dt; iteration += 1; last_dt = dt;
        update_state();
    }

    // ---- field access ------------------------------------------------------------------------------
    struct Ref { F3* f3; F2* f2; int ny, nz; };
    Ref ref(int id) {
        switch (id) {
            case OB200_F_U: return {&u, 0, Ny, Nz};
            case OB200_F_V: return {&v, 0, Ny + 1, Nz};
            case OB200_F_W: return {&w, 0, Ny, Nz + 1};
            case OB200_F_T: return {&T, 0, Ny, Nz};
            case OB200_F_S: return {&S, 0, Ny, Nz};
            case OB200_F_P: return {&p, 0, Ny, Nz};
            case OB200_F_KAPPA_C: return {&kc, 0, Ny, Nz + 1};
            case OB200_F_KAPPA_U: return {&ku, 0, Ny, Nz + 1};
            case OB200_F_GU: return {&Gu, 0, Ny, Nz};
            case OB200_F_GV: return {&Gv, 0, Ny + 1, Nz};
            case OB200_F_GT: return {&GT, 0, Ny, Nz};
            case OB200_F_GS: return {&GS, 0, Ny, Nz};
            case OB200_F_GU_PREV: return {&Gu0, 0, Ny, Nz};
            case OB200_F_GV_PREV: return {&Gv0, 0, Ny + 1, Nz};
            case OB200_F_GT_PREV: return {&GT0, 0, Ny, Nz};
            case OB200_F_GS_PREV: return {&GS0, 0, Ny, Nz};
            case OB200_F_ETA: return {0, &eta, Ny, 1};
            case OB200_F_UBAR: return {0, &Ub, Ny, 1};
            case OB200_F_VBAR: return {0, &Vb, Ny + 1, 1};
            case OB200_F_UTRANS: return {0, &U, Ny, 1};
            case OB200_F_VTRANS: return {0, &V, Ny + 1, 1};
            case OB200_F_GUBT: return {0, &GU, Ny, 1};
            case OB200_F_GVBT: return {0, &GV, Ny + 1, 1};
            case OB200_F_NU_LEITH: return {&nuL, 0, Ny, Nz};
            case OB200_F_QX: return {&qx, 0, Ny, Nz + 1};
            case OB200_F_QY: return {&qy, 0, Ny, Nz + 1};
            case OB200_F_LD: return {0, &Ld, Ny, 1};
            default: return {0, 0, 0, 0};
        }
    }
    int xfer_halo = 0;
    int copy_field(int id, double* buf, int hx, int hy, int hz, bool set) { /*f64*/
        Ref r = ref(id);
        if (!r.f3 && !r.f2) { err = "bad field id"; return OB200_EINVAL; }
        size_t sx = Nx + 2 * hx, sy = r.ny + 2 * hy;
        const int e = std::min(std::min(xfer_halo, hx), H);   // option transfer_x_halo: x-halo columns travel too
        for (int k = 0; k < r.nz; k++)
            for (int j = 0; j < r.ny; j++)
                for (int i = -e; i < Nx + e; i++) {
                    double& b = buf[((size_t)(k + (r.f3 ? hz : 0)) * sy + (j + hy)) * sx + (i + hx)]; /*f64*/
                    real& x = r.f3 ? (*r.f3)(i, j, k) : (*r.f2)(i, j);
                    if (set) x = (real)b; else b = (double)x; /*f64*/
                }
        return OB200_OK;
    }
    // x-halo strips for host-driven exchange (distributed CPU tests): side 0 = west, 1 = east
    int copy_strip(int id, double* buf, int side, int width, bool interior_edge, bool set) { /*f64*/
        Ref r = ref(id);
        if (!r.f3 && !r.f2) { err = "bad field id"; return OB200_EINVAL; }
        int i0 = interior_edge ? (side == 0 ? 0 : Nx - width) : (side == 0 ? -width : Nx);
        size_t n = 0;
        for (int k = 0; k < r.nz; k++)
            for (int j = 0; j < r.ny; j++)
                for (int i = i0; i < i0 + width; i++, n++) {
                    real& x = r.f3 ? (*r.f3)(i, j, k) : (*r.f2)(i, j);
                    if (set) x = (real)buf[n]; else buf[n] = (double)x; /*f64*/
                }
        return OB200_OK;
    }
};

}  // namespace

/*f64-begin*/  // the C interface is Float64 in both builds (buffers are converted at copy_field / copy_strip / set_forcing)
extern "C" {
// The oracle deliberately exposes its own symbol names (oracle_*), not the product's.
void* oracle_create(const ob200_config* cfg) {
    if (!cfg || cfg->struct_bytes != (int32_t)sizeof(ob200_config)) return nullptr;
    Oracle* o = new Oracle();
    if (o->setup(cfg) != OB200_OK) { delete o; return nullptr; }
    return o;
}
void oracle_destroy(void* h) { delete (Oracle*)h; }
int oracle_set_field(void* h, int id, const double* buf, int hx, int hy, int hz) {
    return ((Oracle*)h)->copy_field(id, const_cast<double*>(buf), hx, hy, hz, true);
}
int oracle_get_field(void* h, int id, double* buf, int hx, int hy, int hz) {
    return ((Oracle*)h)->copy_field(id, buf, hx, hy, hz, false);
}
int oracle_set_option(void* h, const char* name, int value) {
    if (!std::strcmp(name, "transfer_x_halo")) { ((Oracle*)h)->xfer_halo = value; return OB200_OK; }
    if (!std::strcmp(name, "reference_arithmetic")) { ((Oracle*)h)->refar = value != 0; return OB200_OK; }
    if (!std::strcmp(name, "reference_arithmetic_mask")) { ((Oracle*)h)->refbits = value; return OB200_OK; }
    if (!std::strcmp(name, "upwind_side_only")) { ((Oracle*)h)->upwind1 = value != 0; return OB200_OK; }
    return OB200_EINVAL;
}
int oracle_get_strip(void* h, int id, double* buf, int side, int width, int interior_edge) {
    return ((Oracle*)h)->copy_strip(id, buf, side, width, interior_edge != 0, false);
}
int oracle_set_strip(void* h, int id, const double* buf, int side, int width, int interior_edge) {
    return ((Oracle*)h)->copy_strip(id, const_cast<double*>(buf), side, width, interior_edge != 0, true);
}
int oracle_set_forcing(void* h, int which, const double* buf) {
    Oracle* o = (Oracle*)h;
    if (which < 0 || which > 5) return OB200_EINVAL;
    size_t n = (size_t)o->Nx * (which == 1 ? o->Ny + 1 : o->Ny) * 6;
    o->forcing[which].assign(buf, buf + n);
    return OB200_OK;
}
int oracle_update_state(void* h) { ((Oracle*)h)->update_state(); return OB200_OK; }
int oracle_time_step(void* h, double dt) { ((Oracle*)h)->time_step(dt); return OB200_OK; }
int oracle_run_stage(void* h, int st, double dt, int arg) { return ((Oracle*)h)->run_stage(st, dt, arg); }
int oracle_get_clock(void* h, double* t, int64_t* it, double* last_dt) {
    Oracle* o = (Oracle*)h; *t = o->time; *it = o->iteration; *last_dt = o->last_dt; return OB200_OK;
}
int oracle_set_clock(void* h, double t, int64_t it, double last_dt) {
    Oracle* o = (Oracle*)h; o->time = t; o->iteration = it; o->last_dt = last_dt; return OB200_OK;
}
// after a step whose stages were driven from the host (distributed tests)
int oracle_tick(void* h, double dt) {
    Oracle* o = (Oracle*)h; o->time += dt; o->iteration += 1; o->last_dt = dt; return OB200_OK;
}
int oracle_zero_prev_tendencies(void* h) {
    Oracle* o = (Oracle*)h;
    for (F3* g : {&o->Gu0, &o->Gv0, &o->GT0, &o->GS0}) std::fill(g->a.begin(), g->a.end(), 0.0);
    return OB200_OK;
}
// standalone WENO entry point for unit tests: biased reconstruction from 2N-1 values
static double weno_entry(int N, const double* q, const double* sa, const double* sb, bool ref) {
    real rq[9], ra[9], rb[9];
    for (int m = 0; m < 2 * N - 1; m++) { rq[m] = (real)q[m]; ra[m] = (real)sa[m]; if (sb) rb[m] = (real)sb[m]; }
    return (double)Oracle::weno(N, rq, ra, sb ? rb : nullptr, ref);
}
double oracle_weno(int N, const double* q, const double* sa, const double* sb) { return weno_entry(N, q, sa, sb, false); }
double oracle_weno_ref(int N, const double* q, const double* sa, const double* sb) { return weno_entry(N, q, sa, sb, true); }
double oracle_teos10_rprime(double Th, double SA, double Z) { static Teos t; return (double)t.rprime((real)Th, (real)SA, (real)Z); }
// r' and its derivatives with respect to Conservative Temperature and Absolute Salinity (what N^2 at faces uses)
double oracle_teos10_rprime_derivatives(double Th, double SA, double Z, double* dTh, double* dSA) {
    static Teos t;
    real a = RL(0.0), b = RL(0.0);
    const real r = t.rprime((real)Th, (real)SA, (real)Z, &a, &b);
    *dTh = (double)a; *dSA = (double)b;
    return (double)r;
}
double oracle_det_tanh(double x) { return (double)det_tanh((real)x); }
int oracle_real_bytes(void) { return (int)sizeof(real); }
// OpenMP team size: torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, which would silently turn the
// reference arm of bench.py into a single-threaded run
void oracle_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int oracle_get_max_threads(void) { return omp_get_max_threads(); }
int oracle_threads_in_parallel_region(void) {
    int n = 1;
#pragma omp parallel
    {
#pragma omp single
        n = omp_get_num_threads();
    }
    return n;
}
}
/*f64-end*/
