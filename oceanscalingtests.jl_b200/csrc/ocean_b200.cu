// libocean_b200.so -- C ABI (include/ocean_b200.h): handle, HBM allocation, sequencing of one
// HydrostaticFreeSurfaceModel time step, x-slab halo exchange over NCCL.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cudaTypedefs.h>
#include <nccl.h>
#include <nvtx3/nvToolsExt.h>

#include "kernels.h"

#ifdef OB200_F32
#define OB_NCCL_REAL ncclFloat
#else
#define OB_NCCL_REAL ncclDouble
#endif

static thread_local std::string g_last_error;
long long ob_launch_count = 0;

struct ob200_handle {
    ob200_config cfg;
    int device = 0;
    Geom G;
    Fields F;
    MomPlan mom_plan;
    real *F_Hfc = nullptr, *F_Hcf = nullptr;
    real *bsendW = nullptr, *bsendE = nullptr, *brecvW = nullptr, *brecvE = nullptr;
    size_t bcount = 0;
    bool substep_exchange = false;   // x-slabs: one-column eta / U exchanges every half substep instead of the wide halo
    real *e_send_eta = nullptr, *e_recv_eta = nullptr, *e_send_U = nullptr, *e_recv_U = nullptr;   // Ny doubles each
    bool periodic_local = true;
    bool own_stream = true;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[20];
    bool timing = false, use_graph = true, persistent_baro = false, overlap = false;
    bool onchip_baro = false;   // option "onchip_barotropic": persistent kernel with eta, U, V resident in shared memory
    bool overlap_exchange = false;  // x-slabs: halo exchange on stream2 beside the interior columns of the auxiliary sweeps (option)
    int transfer_x_halo = 0;   // x-halo columns copied by set/get_field besides the interior (checkpoint of halo-carrying state)
    cudaStream_t stream2 = nullptr;            // tracer work that can overlap the momentum / barotropic chain
    cudaStream_t stream3 = nullptr;            // masked tendency tiles beside the full-order tiles
    cudaEvent_t ev_tend[2] = {nullptr, nullptr};   // auxiliaries done (fork), masked tiles done (join)
    bool concurrent_masked = false;            // option "concurrent_masked_tiles" (measured: no gain, 53.0 vs 52.9 ms at C3)
    bool onchip_solve = false;                 // option "onchip_solve": implicit solve with t, f' in shared memory
    cudaEvent_t ev_join[4];                    // step start, T/S solved, w ready, tracer tendencies done
    cudaEvent_t ev2[4];                        // timing events on stream2
    int n_timed = 0;
    std::vector<void*> allocs;
    struct Bound { real* ptr = nullptr; int hx = 0, hy = 0, hz = 0; };
    Bound bound[OB200_F_COUNT];   // caller-owned device parent arrays (ob200_bind)
    std::vector<double> weights;
    real* d_weights = nullptr;
    double* d_diag = nullptr;
    // clock
    double time = 0.0, last_dt = INFINITY;
    int64_t iteration = 0;
    bool fields_dirty = true;   // set_field since the last mask
    // F.Gab = ca G^n - cb G^- of the CURRENT G^n, G^- (written by the last tendencies launch, nothing touched G since): the next
    // regular AB2 step's implicit solve reads it instead of G^n and G^-
    bool gab_valid = false;
    // NCCL ring
    ncclComm_t comm = nullptr;
    real *sendW[2] = {nullptr, nullptr}, *sendE[2] = {nullptr, nullptr}, *recvW[2] = {nullptr, nullptr}, *recvE[2] = {nullptr, nullptr};
    size_t xcount[2] = {0, 0};      // per exchange group (OB_XG_UVETA, OB_XG_TS)
    bool early_ts_exchange = true;  // option "early_ts_exchange": T, S halos travel right after their solve, on stream2
    cudaEvent_t ev_ts[2] = {nullptr, nullptr};   // T, S solved (fork) / their halos unpacked (join)
    // CUDA graph of the barotropic substep launches at fixed dt
    cudaGraphExec_t graph_exec = nullptr;
    double graph_dt = NAN;
    std::string err;

    int fail(int code, const char* what, const char* detail) {
        err = std::string(what) + ": " + (detail ? detail : "");
        g_last_error = err;
        return code;
    }
};

namespace {

template <class T>
int dev_alloc(ob200_handle* h, T** p, size_t n, bool zero = true) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, n * sizeof(T));
    if (e != cudaSuccess) return h->fail(OB200_ENOMEM, "cudaMalloc", cudaGetErrorString(e));
    if (zero) cudaMemset(q, 0, n * sizeof(T));
    h->allocs.push_back(q);
    *p = (T*)q;
    return OB200_OK;
}

// host arrays are computed in Float64 and stored in the library's precision (identity copy in the Float64 build)
int upload_real(ob200_handle* h, const double* src, size_t n, real** dst) {
    if (int rc = dev_alloc(h, dst, n, false)) return rc;
#ifdef OB200_F32
    std::vector<real> t(src, src + n);
    cudaMemcpy(*dst, t.data(), n * sizeof(real), cudaMemcpyHostToDevice);
#else
    cudaMemcpy(*dst, src, n * sizeof(real), cudaMemcpyHostToDevice);
#endif
    return OB200_OK;
}

double wind_stress(const ob200_config& c, double phi) {   // bathymetry_and_forcings.jl:110-119
    const double* q = phi < -45 ? c.wind_coeffs : phi < -15 ? c.wind_coeffs + 4 : phi < 0 ? c.wind_coeffs + 8
                    : phi < 15 ? c.wind_coeffs + 12 : phi < 45 ? c.wind_coeffs + 16 : c.wind_coeffs + 20;
    return q[0] * phi * phi * phi + q[1] * phi * phi + q[2] * phi + q[3];
}
double salinity_flux(const ob200_config& c, double phi) {   // :121-128
    const double* q = phi < -20 ? c.salt_coeffs : phi < 0 ? c.salt_coeffs + 4 : phi < 20 ? c.salt_coeffs + 8 : c.salt_coeffs + 12;
    return q[0] * phi * phi * phi + q[1] * phi * phi + q[2] * phi + q[3];
}

int setup_geometry(ob200_handle* h) {
    const ob200_config& c = h->cfg;
    Geom& G = h->G;
    std::memset(&G, 0, sizeof(G));
    G.Nx = c.Nx; G.Ny = c.Ny; G.Nz = c.Nz;
    G.pitch = OB_X0 + ((c.Nx + OB_H + 15) / 16) * 16;
    G.NyP = c.Ny + 1 + 2 * OB_H;
    G.sz = (long long)G.pitch * G.NyP;
    // barotropic arrays: halo of Wb columns (M + 1 on x-slabs: exchanged once, then no communication in the loop)
    G.Wb = (h->periodic_local || h->substep_exchange) ? OB_H : std::max(OB_H, c.n_substeps + 1);
    G.X0b = ((G.Wb + 15) / 16) * 16;
    G.pitch2 = G.X0b + ((c.Nx + G.Wb + 15) / 16) * 16;
    const double pi = M_PI;
    const double dlam = (c.lon_east - c.lon_west) / c.Nx_global;
    const double dphi = (c.lat_north - c.lat_south) / c.Ny;
    const double dlam_rad = dlam * pi / 180.0;
    const double dy = c.radius * (dphi * pi / 180.0);
    G.dy = dy;
    // per-latitude metrics for j = -(H+1) .. Ny+H+1
    const int jo = OB_H + 1, nj = c.Ny + 2 * jo + 1;
    std::vector<double> m[11];
    for (auto& v : m) v.assign(nj, 0.0);
    auto pf = [&](int j) { return (double)((long double)c.lat_south + (long double)j * (long double)dphi); };
    auto pc = [&](int j) { return (double)((long double)c.lat_south + ((long double)j + 0.5L) * (long double)dphi); };
    auto cosd = [&](double x) { return std::cos(pi * x / 180.0); };
    auto sind = [&](double x) { return std::sin(pi * x / 180.0); };
    for (int jj = 0; jj < nj; jj++) {
        const int j = jj - jo;
        m[0][jj] = c.radius * cosd(pc(j)) * dlam_rad;                                        // dxfc = dxcc
        m[1][jj] = c.radius * cosd(pf(j)) * dlam_rad;                                        // dxcf = dxff
        m[2][jj] = c.radius * c.radius * dlam_rad * (sind(pf(j + 1)) - sind(pf(j)));         // azcc = azfc
        m[3][jj] = c.radius * c.radius * dlam_rad * (sind(pc(j)) - sind(pc(j - 1)));         // azcf = azff
        m[4][jj] = 2.0 * c.rotation_rate * sind(pf(j));                                      // f at (f,f)
        m[5][jj] = wind_stress(c, pc(j));
        m[6][jj] = salinity_flux(c, pc(j));
        m[7][jj] = std::fmax(0.0, 30.0 * std::cos(1.2 * pi * pc(j) / 180.0));                // T_reference (:154)
        m[8][jj] = 1.0 / m[0][jj]; m[9][jj] = 1.0 / m[2][jj]; m[10][jj] = 1.0 / m[3][jj];
    }
    G.rdy = 1.0 / dy;
    G.ab2_ca = (real)1.5 + (real)c.ab2_chi; G.ab2_cb = (real)0.5 + (real)c.ab2_chi;   // as k_ab2_implicit forms them
    const real** dst[11] = {&G.dxfc, &G.dxcf, &G.azcc, &G.azcf, &G.fff, &G.taux, &G.sflux, &G.tref, &G.rdxfc, &G.razcc, &G.razcf};
    for (int q = 0; q < 11; q++) {
        real* d = nullptr;
        if (int rc = upload_real(h, m[q].data(), nj, &d)) return rc;
        *dst[q] = d + jo;
    }
    // vertical metrics
    const int Nz = c.Nz;
    std::vector<double> zc(Nz + 2), dzc(Nz + 2), dzf(Nz + 1);
    for (int k = 0; k < Nz; k++) { dzc[k + 1] = c.z_faces[k + 1] - c.z_faces[k]; zc[k + 1] = 0.5 * (c.z_faces[k + 1] + c.z_faces[k]); }
    dzc[0] = dzc[1]; dzc[Nz + 1] = dzc[Nz];
    zc[0] = c.z_faces[0] - 0.5 * dzc[0]; zc[Nz + 1] = c.z_faces[Nz] + 0.5 * dzc[Nz + 1];
    for (int k = 0; k <= Nz; k++) dzf[k] = zc[k + 1] - zc[k];
    real *dzc_d, *zc_d, *dzf_d, *rdzc_d;
    std::vector<double> rdzc(Nz + 2);
    for (int k = 0; k < Nz + 2; k++) rdzc[k] = 1.0 / dzc[k];
    if (int rc = upload_real(h, rdzc.data(), Nz + 2, &rdzc_d)) return rc;
    G.rdzc = rdzc_d + 1;
    {
        std::vector<double> rdzf(Nz + 1), rupz(Nz + 1, 0.0), rloz(Nz + 1, 0.0);
        for (int k = 0; k <= Nz; k++) rdzf[k] = 1.0 / dzf[k];
        for (int k = 0; k < Nz; k++) { rupz[k] = 1.0 / (dzc[k + 1] * dzf[k + 1]); rloz[k] = 1.0 / (dzc[k + 1] * dzf[k]); }
        real* d[3];
        const std::vector<double>* src[3] = {&rdzf, &rupz, &rloz};
        for (int q = 0; q < 3; q++)
            if (int rc = upload_real(h, src[q]->data(), Nz + 1, &d[q])) return rc;
        G.rdzf = d[0]; G.rupz = d[1]; G.rloz = d[2];
    }
    if (int rc = upload_real(h, dzc.data(), Nz + 2, &dzc_d)) return rc;
    if (int rc = upload_real(h, zc.data(), Nz + 2, &zc_d)) return rc;
    if (int rc = upload_real(h, dzf.data(), Nz + 1, &dzf_d)) return rc;
    G.dzc = dzc_d + 1; G.zc = zc_d + 1; G.dzf = dzf_d;
    {
        real* zfa_d;
        if (int rc = upload_real(h, c.z_faces, Nz + 1, &zfa_d)) return rc;
        G.zfa = zfa_d;
    }
    // bathymetry -> first active level per column; rows outside the y-domain are fully inactive
    const size_t n2 = (size_t)G.pitch * G.NyP;
    std::vector<int> kb(n2, Nz), kfast(n2, Nz);
    const int bh = c.bottom_halo_columns, nxh = c.Nx + 2 * bh;
    std::vector<int> kbw((size_t)nxh * (c.Ny + 2), Nz);   // wide copy incl. one inactive row south and north
    for (int j = 0; j < c.Ny; j++)
        for (int ii = 0; ii < nxh; ii++) {
            const double hgt = c.bottom_height[(size_t)j * nxh + ii];
            int k0 = 0;
            while (k0 < Nz && zc[k0 + 1] <= hgt) k0++;
            kbw[(size_t)(j + 1) * nxh + ii] = k0;
            const int i = ii - bh;
            if (i >= -OB_H && i < c.Nx + OB_H) kb[(size_t)(j + OB_H) * G.pitch + (i + OB_X0)] = k0;
        }
    {   // wet column depths at u / v points: sum of dz over non-peripheral nodes, wide halo
        const size_t n2b = (size_t)G.pitch2 * G.NyP;
        std::vector<double> Hfc(n2b, 0.0), Hcf(n2b, 0.0);
        const int lim = std::min(bh, G.Wb);
        for (int j = 0; j <= c.Ny; j++)
            for (int i = -lim + 1; i < c.Nx + lim; i++) {
                const int ii = i + bh;
                const int kc = kbw[(size_t)(j + 1) * nxh + ii];
                const int ku = std::max(kc, kbw[(size_t)(j + 1) * nxh + ii - 1]);
                const int kv = std::max(kc, kbw[(size_t)j * nxh + ii]);
                double a = 0.0, b = 0.0;
                for (int k = ku; k < Nz; k++) a += dzc[k + 1];
                for (int k = kv; k < Nz; k++) b += dzc[k + 1];
                Hfc[(size_t)(j + OB_H) * G.pitch2 + (i + G.X0b)] = a;
                Hcf[(size_t)(j + OB_H) * G.pitch2 + (i + G.X0b)] = b;
            }
        if (int rc = upload_real(h, Hfc.data(), n2b, &h->F_Hfc)) return rc;
        if (int rc = upload_real(h, Hcf.data(), n2b, &h->F_Hcf)) return rc;
    }
    // first level from which the mask-free full-order stencils are valid: every cell within 5 columns / rows
    // wet and in the domain; the level just above a neighbour's sea floor still needs the conditional
    // vertical flux, hence K + 1 unless K == 0.  Only when the configured orders are the defaults (9, 5, 7).
    const bool default_orders = c.order_vorticity == 9 && c.order_divergence == 5 && c.order_tracer == 7;
    if (default_orders)
        for (int j = 0; j < c.Ny; j++)
            for (int i = 0; i < c.Nx; i++) {
                int K = 0;
                for (int b = -5; b <= 5; b++)
                    for (int a = -5; a <= 5; a++) {
                        const int jj = j + b;
                        const int v = (jj < 0 || jj >= c.Ny) ? Nz : kb[(size_t)(jj + OB_H) * G.pitch + (i + a + OB_X0)];
                        K = v > K ? v : K;
                    }
                kfast[(size_t)(j + OB_H) * G.pitch + (i + OB_X0)] = (K == 0) ? 0 : (K >= Nz ? Nz : K + 1);
            }
    int *kb_d, *kf_d;
    if (int rc = dev_alloc(h, &kb_d, n2, false)) return rc;
    if (int rc = dev_alloc(h, &kf_d, n2, false)) return rc;
    cudaMemcpy(kb_d, kb.data(), n2 * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(kf_d, kfast.data(), n2 * sizeof(int), cudaMemcpyHostToDevice);
    G.kb = kb_d; G.kfast = kf_d;
    {   // per-tile plan for the momentum kernels (kernels.h MomPlan)
        int tx, ty;
        mom_plan_tile_shape(&tx, &ty);
        MomPlan& P = h->mom_plan;
        P.ntx = (c.Nx + tx - 1) / tx; P.nty = (c.Ny + ty - 1) / ty;
        std::vector<int> tk((size_t)P.ntx * P.nty, 0), slow;
        for (int b = 0; b < P.nty; b++)
            for (int a = 0; a < P.ntx; a++) {
                int K = 0, kmin = Nz;
                for (int j = b * ty; j < std::min((b + 1) * ty, c.Ny); j++)
                    for (int i = a * tx; i < std::min((a + 1) * tx, c.Nx); i++) {
                        K = std::max(K, kfast[(size_t)(j + OB_H) * G.pitch + (i + OB_X0)]);
                        kmin = std::min(kmin, kb[(size_t)(j + OB_H) * G.pitch + (i + OB_X0)]);
                    }
                tk[(size_t)b * P.ntx + a] = K;
                if (K < Nz) P.any_fast = true;
                // levels below kmin hold no active cell of the tile: every G there is a masked zero (the arrays are
                // zero-initialised and those entries are never written) -- the tile-level analogue of the
                // reference's active_cells_map (grid_load_balance.jl:21-32); all-solid tiles are skipped entirely
                if (K > 0 && kmin < K) { slow.push_back(b * P.ntx + a); slow.push_back(K); slow.push_back(kmin); }
            }
        P.nslow = (int)slow.size() / 3;
        if (int rc = dev_alloc(h, &P.tile_kfast, tk.size(), false)) return rc;
        cudaMemcpy(P.tile_kfast, tk.data(), tk.size() * sizeof(int), cudaMemcpyHostToDevice);
        if (P.nslow) {
            if (int rc = dev_alloc(h, &P.slow_tiles, slow.size(), false)) return rc;
            cudaMemcpy(P.slow_tiles, slow.data(), slow.size() * sizeof(int), cudaMemcpyHostToDevice);
        }
    }
    // parameters
    G.g = c.gravity; G.alpha = c.eos_alpha; G.beta_s = c.eos_beta; G.rho0 = c.eos_rho0; G.eos = c.eos_kind;
    G.nu_z = c.nu_z; G.kappa_z = c.kappa_z; G.ri_enabled = c.ri_enabled;
    G.ri_nu0 = c.ri_nu0; G.ri_kappa0 = c.ri_kappa0; G.ri_kappa_ca = c.ri_kappa_ca; G.ri_Cen = c.ri_Cen;
    G.ri_Cav = c.ri_Cav; G.ri_Ri0 = c.ri_Ri0; G.ri_Ridelta = c.ri_Ridelta;
    G.r1cav = 1.0 / (1.0 + c.ri_Cav); G.rRidelta = 1.0 / c.ri_Ridelta;
    G.bc_kind = c.bc_kind; G.lambda_T = c.T_restoring_lambda; G.mu_drag = c.drag_mu;
    G.qgleith = c.qgleith_enabled; G.leith_C = c.leith_C; G.leith_minN2 = c.leith_min_N2; G.leith_Vscale = c.leith_Vscale;
    return OB200_OK;
}

// TMA descriptor of one 3-D field: dims (pitch, NyP, Nz+1) elements, box (bx, by, 1); out-of-range elements read 0
int encode_tile_map(ob200_handle* h, CUtensorMap* tm, real* base, int bx, int by) {
    static PFN_cuTensorMapEncodeTiled encode = nullptr;
    if (!encode) {
        cudaDriverEntryPointQueryResult q;
        void* fn = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn)
            return h->fail(OB200_ECUDA, "cuTensorMapEncodeTiled", "driver entry point not found");
        encode = (PFN_cuTensorMapEncodeTiled)fn;
    }
    const Geom& G = h->G;
    cuuint64_t dims[3] = {(cuuint64_t)G.pitch, (cuuint64_t)G.NyP, (cuuint64_t)(G.Nz + 1)};
    cuuint64_t strides[2] = {(cuuint64_t)G.pitch * sizeof(real), (cuuint64_t)G.sz * sizeof(real)};
    cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode(tm, sizeof(real) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return h->fail(OB200_ECUDA, "cuTensorMapEncodeTiled", "encode failed");
    return OB200_OK;
}

int alloc_fields(ob200_handle* h) {
    const Geom& G = h->G;
    Fields& F = h->F;
    std::memset(&F, 0, sizeof(F));
    const size_t n3 = (size_t)G.sz * (G.Nz + 1), n2 = (size_t)G.sz;
    real** f3[] = {&F.u, &F.v, &F.w, &F.T, &F.S, &F.p, &F.kc, &F.ku, &F.Ri, &F.scratch, &F.scratch2,
                     &F.Gu, &F.Gv, &F.GT, &F.GS, &F.Gu0, &F.Gv0, &F.GT0, &F.GS0};
    for (real** p : f3)
        if (int rc = dev_alloc(h, p, n3)) return rc;
    if (G.qgleith)
        for (real** p : {&F.nuL, &F.qx, &F.qy})
            if (int rc = dev_alloc(h, p, n3)) return rc;
    const size_t n2b = (size_t)G.pitch2 * G.NyP;
    real** f2[] = {&F.eta, &F.U, &F.V, &F.etab, &F.Ub, &F.Vb, &F.GU, &F.GV, &F.Usum, &F.Vsum};
    for (real** p : f2)
        if (int rc = dev_alloc(h, p, n2b)) return rc;
    F.Hfc = h->F_Hfc; F.Hcf = h->F_Hcf;
    if (int rc = dev_alloc(h, &F.Ld, n2)) return rc;
    {   // combined AB2 tendencies (kernels.h Fields::Gab): four more 3-D arrays, OPTION OB200_AB2_COMBINED=1.  Bit-identical
        // (tests/test_parity_gpu.py with the variable set), measured at C3 on one B200: the implicit solve drops from 11.14
        // to 9.94 ms (7.5 GB less DRAM traffic), but the tendency kernels pay 0.66 + 0.93 ms for the extra load and store
        // (51.98 vs 51.60 ms per step) -- off by default
        const char* e = getenv("OB200_AB2_COMBINED");
        if (e && e[0] == '1')
            for (int q = 0; q < 4; q++)
                if (int rc = dev_alloc(h, &F.Gab[q], n3)) return rc;
    }
    {   // TMA descriptors for the tile staging of the tendency kernels (OB200_NO_TMA=1 falls back to register staging)
        MomPlan& P = h->mom_plan;
        int mx, my, tx, ty;
        tend_tile_boxes(&mx, &my, &tx, &ty);
        const char* e = getenv("OB200_NO_TMA");
        P.use_tma = !(e && e[0] == '1');
#ifdef OB200_F32
        P.use_tma = false;   // the tile boxes (start column, row bytes) are laid out for 8-byte elements: register staging
#endif
        if (P.use_tma) {
            if (int rc = encode_tile_map(h, &P.tm_u, F.u, mx, my)) return rc;
            if (int rc = encode_tile_map(h, &P.tm_v, F.v, mx, my)) return rc;
            if (int rc = encode_tile_map(h, &P.tm_T, F.T, tx, ty)) return rc;
            if (int rc = encode_tile_map(h, &P.tm_S, F.S, tx, ty)) return rc;
        }
    }
    return OB200_OK;
}

struct FieldRef { real* p; int ny, nz; int layout; };   // layout: 0 = 3-D, 1 = 2-D plane layout, 2 = barotropic
FieldRef field_ref(ob200_handle* h, int id) {
    const Geom& G = h->G;
    Fields& F = h->F;
    const int Ny = G.Ny, Nz = G.Nz;
    switch (id) {
        case OB200_F_U: return {F.u, Ny, Nz, 0};
        case OB200_F_V: return {F.v, Ny + 1, Nz, 0};
        case OB200_F_W: return {F.w, Ny, Nz + 1, 0};
        case OB200_F_T: return {F.T, Ny, Nz, 0};
        case OB200_F_S: return {F.S, Ny, Nz, 0};
        case OB200_F_P: return {F.p, Ny, Nz, 0};
        case OB200_F_KAPPA_C: return {F.kc, Ny, Nz + 1, 0};
        case OB200_F_KAPPA_U: return {F.ku, Ny, Nz + 1, 0};
        case OB200_F_GU: return {F.Gu, Ny, Nz, 0};
        case OB200_F_GV: return {F.Gv, Ny + 1, Nz, 0};
        case OB200_F_GT: return {F.GT, Ny, Nz, 0};
        case OB200_F_GS: return {F.GS, Ny, Nz, 0};
        case OB200_F_GU_PREV: return {F.Gu0, Ny, Nz, 0};
        case OB200_F_GV_PREV: return {F.Gv0, Ny + 1, Nz, 0};
        case OB200_F_GT_PREV: return {F.GT0, Ny, Nz, 0};
        case OB200_F_GS_PREV: return {F.GS0, Ny, Nz, 0};
        case OB200_F_ETA: return {F.eta, Ny, 1, 2};
        case OB200_F_UBAR: return {F.Ub, Ny, 1, 2};
        case OB200_F_VBAR: return {F.Vb, Ny + 1, 1, 2};
        case OB200_F_UTRANS: return {F.U, Ny, 1, 2};
        case OB200_F_VTRANS: return {F.V, Ny + 1, 1, 2};
        case OB200_F_GUBT: return {F.GU, Ny, 1, 2};
        case OB200_F_GVBT: return {F.GV, Ny + 1, 1, 2};
        case OB200_F_NU_LEITH: return {F.nuL, Ny, Nz, 0};
        case OB200_F_QX: return {F.qx, Ny, Nz + 1, 0};
        case OB200_F_QY: return {F.qy, Ny, Nz + 1, 0};
        case OB200_F_LD: return {F.Ld, Ny, 1, 1};
        default: return {nullptr, 0, 0, 0};
    }
}

// ---- stages ------------------------------------------------------------------------------------------
// groups: bit 0 = u, v, eta; bit 1 = T, S
int exchange_halos(ob200_handle* h, cudaStream_t st = nullptr, int groups = 3) {
    if (h->periodic_local) return OB200_OK;
    if (!st) st = h->stream;
    if (!h->comm) return h->fail(OB200_ESTATE, "exchange_halos", "nranks > 1 but ob200_comm_init was not called");
    const int r = h->cfg.rank, n = h->cfg.nranks;
    const int west = (r + n - 1) % n, east = (r + 1) % n;
    for (int g = 0; g < 2; g++)
        if (groups & (1 << g)) launch_pack_x(h->G, h->F, h->sendW[g], h->sendE[g], g, st);
    // order matters on a 2-rank ring (west == east == the same peer): NCCL pairs sends and receives between two
    // ranks in issue order, and my west halo must receive the peer's EAST edge -> east-bound message first
    ncclGroupStart();
    for (int g = 0; g < 2; g++)
        if (groups & (1 << g)) {
            ncclSend(h->sendE[g], h->xcount[g], OB_NCCL_REAL, east, h->comm, st);
            ncclSend(h->sendW[g], h->xcount[g], OB_NCCL_REAL, west, h->comm, st);
            ncclRecv(h->recvW[g], h->xcount[g], OB_NCCL_REAL, west, h->comm, st);
            ncclRecv(h->recvE[g], h->xcount[g], OB_NCCL_REAL, east, h->comm, st);
        }
    ncclResult_t rc = ncclGroupEnd();
    if (rc != ncclSuccess) return h->fail(OB200_ENCCL, "halo exchange", ncclGetErrorString(rc));
    for (int g = 0; g < 2; g++)
        if (groups & (1 << g)) launch_unpack_x(h->G, h->F, h->recvW[g], h->recvE[g], g, st);
    return OB200_OK;
}

int barotropic_loop(ob200_handle* h, double dt);

// `ts_in_flight`: the T, S halos were sent on stream2 right after their solve (ev_ts[1] marks their arrival)
int mask_and_halos(ob200_handle* h, bool ts_in_flight = false) {
    if (h->fields_dirty) { launch_mask_fields(h->G, h->F, h->stream); h->fields_dirty = false; }
    if (int rc = exchange_halos(h, nullptr, ts_in_flight ? 1 : 3)) return rc;
    if (ts_in_flight) cudaStreamWaitEvent(h->stream, h->ev_ts[1], 0);
    launch_fill_halos(h->G, h->F, h->periodic_local, h->stream);
    return OB200_OK;
}
// T, S are final once their implicit solve is done: their x-halos travel on stream2 while the main stream goes on with the
// u, v solve, the barotropic loop and the corrector (half of the exchange payload off the critical path)
int early_ts_exchange(ob200_handle* h) {
    cudaEventRecord(h->ev_ts[0], h->stream);
    cudaStreamWaitEvent(h->stream2, h->ev_ts[0], 0);
    if (int rc = exchange_halos(h, h->stream2, 2)) return rc;
    cudaEventRecord(h->ev_ts[1], h->stream2);
    return OB200_OK;
}
void auxiliaries(ob200_handle* h) {
    launch_w_p(h->G, h->F, h->stream);
    launch_ri_kappa(h->G, h->F, h->time, h->stream);
    if (h->G.qgleith) launch_leith(h->G, h->F, h->stream);
}
// Tendencies.  The masked tiles (walls, ridges, sea floor: latency-bound, 2 resident blocks per SM) write (tile, level)
// sets disjoint from those of the full-order tiles (FP64-pipe-bound), so their two launches run on a third stream BESIDE
// the full-order launches and fill the issue slots those leave idle; joined before anything reads G.
// `mid` / `end` (events or null): recorded on the main stream between momentum and tracers / after the join (stage timing).
void tendencies(ob200_handle* h, cudaEvent_t mid = nullptr, cudaEvent_t end = nullptr) {
    const Geom& G = h->G;
    cudaStream_t s1 = h->stream, s3 = h->stream3;
    const MomPlan& P = h->mom_plan;
    const bool fork = h->concurrent_masked && P.nslow > 0 && P.any_fast;
    if (fork) {
        cudaEventRecord(h->ev_tend[0], s1);
        cudaStreamWaitEvent(s3, h->ev_tend[0], 0);
        launch_momentum(G, h->F, P, h->cfg.order_vorticity, h->cfg.order_divergence, h->time, s3, 2);
        launch_tracers(G, h->F, P, h->cfg.order_tracer, h->time, s3, 2);
        cudaEventRecord(h->ev_tend[1], s3);
    }
    launch_momentum(G, h->F, P, h->cfg.order_vorticity, h->cfg.order_divergence, h->time, s1, fork ? 1 : 3);
    if (mid) cudaEventRecord(mid, s1);
    launch_tracers(G, h->F, P, h->cfg.order_tracer, h->time, s1, fork ? 1 : 3);
    h->gab_valid = true;
    if (fork) cudaStreamWaitEvent(s1, h->ev_tend[1], 0);
    if (end) cudaEventRecord(end, s1);
}
// mask + halos + the two auxiliary sweeps.  On x-slabs the NCCL exchange of the 7-column halos runs on the second
// stream BESIDE the sweeps of the interior columns (w + Ri needs u(i + 1): columns [0, Nx - 1); pHY' + kappa reads Ri at
// i - 1 .. i + 1: columns [1, Nx - 2)); the few columns whose stencils reach into the halo follow once it has arrived.
// `ev` (3 events or null): after the halos were issued / w + Ri / pHY' + kappa, for the stage timing.
// Option "overlap_exchange", off by default: bit-identical, but measured SLOWER on 2 B200 at C3 (halo + sweeps 3.89 vs
// 3.58 ms): a column sweep is latency-bound (one thread marches 120 levels), so the extra launches for the few
// halo-dependent columns cost a full sweep latency each (~0.35 ms), more than the 0.17 ms of exchange they hide.
int halos_and_auxiliaries(ob200_handle* h, cudaEvent_t* ev, bool ts_in_flight = false) {
    const Geom& G = h->G;
    cudaStream_t s1 = h->stream, s2 = h->stream2;
    const int Nx = G.Nx;
    if (h->periodic_local || !h->overlap_exchange) {
        if (int rc = mask_and_halos(h, ts_in_flight)) return rc;
        if (ev) cudaEventRecord(ev[0], s1);
        launch_w_p(G, h->F, s1);
        if (ev) cudaEventRecord(ev[1], s1);
        launch_ri_kappa(G, h->F, h->time, s1);
        if (G.qgleith) launch_leith(G, h->F, s1);
        if (ev) cudaEventRecord(ev[2], s1);
        return OB200_OK;
    }
    if (h->fields_dirty) { launch_mask_fields(G, h->F, s1); h->fields_dirty = false; }
    cudaEventRecord(h->ev_join[2], s1);
    cudaStreamWaitEvent(s2, h->ev_join[2], 0);
    if (int rc = exchange_halos(h, s2)) return rc;
    cudaEventRecord(h->ev_join[3], s2);
    if (ev) cudaEventRecord(ev[0], s1);
    launch_w_p(G, h->F, 0, Nx - 1, 0, 0, s1);
    launch_ri_kappa(G, h->F, h->time, 1, Nx - 3, 0, 0, s1);
    cudaStreamWaitEvent(s1, h->ev_join[3], 0);
    launch_fill_halos(G, h->F, false, s1);
    launch_w_p(G, h->F, -2, 2, Nx - 1, 3, s1);
    if (ev) cudaEventRecord(ev[1], s1);
    launch_ri_kappa(G, h->F, h->time, -2, 3, Nx - 2, 4, s1);
    if (G.qgleith) launch_leith(G, h->F, s1);
    if (ev) cudaEventRecord(ev[2], s1);
    return OB200_OK;
}
int update_state(ob200_handle* h) {
    if (int rc = halos_and_auxiliaries(h, nullptr)) return rc;
    tendencies(h);
    return OB200_OK;
}
void swap_tendencies(ob200_handle* h) {   // store_tendencies!: G^- <- G^n by pointer swap
    Fields& F = h->F;
    std::swap(F.Gu, F.Gu0); std::swap(F.Gv, F.Gv0); std::swap(F.GT, F.GT0); std::swap(F.GS, F.GS0);
    h->gab_valid = false;   // until the next tendencies
}

int barotropic_wide_exchange(ob200_handle* h) {
    // one exchange of Wb-column halos of Ubar, Vbar, eta, G^U, G^V before the substep loop
    if (!h->comm) return h->fail(OB200_ESTATE, "barotropic_loop", "nranks > 1 but ob200_comm_init was not called");
    const int r = h->cfg.rank, n = h->cfg.nranks, W = h->G.Wb;
    const int west = (r + n - 1) % n, east = (r + 1) % n;
    launch_pack_baro(h->G, h->F, h->bsendW, h->bsendE, W, false, h->stream);
    ncclGroupStart();   // east-bound first: see exchange_halos
    ncclSend(h->bsendE, h->bcount, OB_NCCL_REAL, east, h->comm, h->stream);
    ncclSend(h->bsendW, h->bcount, OB_NCCL_REAL, west, h->comm, h->stream);
    ncclRecv(h->brecvW, h->bcount, OB_NCCL_REAL, west, h->comm, h->stream);
    ncclRecv(h->brecvE, h->bcount, OB_NCCL_REAL, east, h->comm, h->stream);
    ncclResult_t rc = ncclGroupEnd();
    if (rc != ncclSuccess) return h->fail(OB200_ENCCL, "barotropic halo exchange", ncclGetErrorString(rc));
    launch_pack_baro(h->G, h->F, h->brecvW, h->brecvE, W, true, h->stream);
    return OB200_OK;
}

// one-column ring exchange: my `send` goes to `to`, `recv` comes from the opposite neighbour
int edge_exchange(ob200_handle* h, const real* send, real* recv, int to, int from) {
    ncclGroupStart();
    ncclSend(send, h->G.Ny, OB_NCCL_REAL, to, h->comm, h->stream);
    ncclRecv(recv, h->G.Ny, OB_NCCL_REAL, from, h->comm, h->stream);
    ncclResult_t rc = ncclGroupEnd();
    if (rc != ncclSuccess) return h->fail(OB200_ENCCL, "barotropic edge exchange", ncclGetErrorString(rc));
    return OB200_OK;
}
// substeps with one-column exchanges: eta(Nx - 1) travels east after every eta half step, U(0) west after every velocity
// half step (2 M - 1 + 1 NCCL groups per baroclinic step; the loop is captured in a CUDA graph like the wide-halo loop)
int barotropic_substeps_edge(ob200_handle* h, double dtau) {
    const int r = h->cfg.rank, n = h->cfg.nranks, west = (r + n - 1) % n, east = (r + 1) % n, M = h->cfg.n_substeps;
    launch_baro_edge_first(h->G, h->F, h->e_send_U, h->stream);
    if (int rc = edge_exchange(h, h->e_send_U, h->e_recv_U, west, east)) return rc;
    for (int m = 0; m < M; m++) {
        launch_baro_eta_edge(h->G, h->F, dtau, h->e_send_eta, h->e_recv_eta, h->e_send_U, h->e_recv_U, h->stream);
        if (int rc = edge_exchange(h, h->e_send_eta, h->e_recv_eta, east, west)) return rc;
        launch_baro_vel_edge(h->G, h->F, dtau, h->weights[m], h->e_send_eta, h->e_recv_eta, h->e_send_U, h->e_recv_U, h->stream);
        if (m + 1 < M)
            if (int rc = edge_exchange(h, h->e_send_U, h->e_recv_U, west, east)) return rc;
    }
    launch_baro_finish(h->G, h->F, h->stream);
    return OB200_OK;
}

int barotropic_loop(ob200_handle* h, double dt) {
    const double dtau = h->cfg.frac_dtau * dt;
    const bool per = h->periodic_local;
    const int ex = per ? 0 : h->G.Wb - 1;
    if (!per && h->substep_exchange) {
        if (!h->comm) return h->fail(OB200_ESTATE, "barotropic_loop", "nranks > 1 but ob200_comm_init was not called");
        launch_baro_init(h->G, h->F, false, 0, h->stream);
        if (!h->use_graph) return barotropic_substeps_edge(h, dtau);
        if (!h->graph_exec || h->graph_dt != dt) {
            if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }
            cudaGraph_t graph = nullptr;
            const long long before = ob_launch_count;
            OB_CUDA_CHECK(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
            const int rc = barotropic_substeps_edge(h, dtau);
            const cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
            ob_launch_count = before;
            if (rc) return rc;
            OB_CUDA_CHECK(h, ce);
            OB_CUDA_CHECK(h, cudaGraphInstantiate(&h->graph_exec, graph, 0));
            cudaGraphDestroy(graph);
            h->graph_dt = dt;
        }
        OB_CUDA_CHECK(h, cudaGraphLaunch(h->graph_exec, h->stream));
        OB_COUNT_LAUNCH(2 * h->cfg.n_substeps + 2);
        return OB200_OK;
    }
    if (!per)
        if (int rc = barotropic_wide_exchange(h)) return rc;
    launch_baro_init(h->G, h->F, per, ex, h->stream);
    if (h->onchip_baro && baro_onchip_fits(h->G, ex)) {
        // one cooperative launch, state in shared memory; writes eta <- eta_bar itself (no k_baro_finish)
        if (launch_baro_onchip(h->G, h->F, dtau, h->d_weights, h->cfg.n_substeps, per, ex, h->stream) != 0)
            return h->fail(OB200_ECUDA, "cooperative launch (on-chip barotropic)", cudaGetErrorString(cudaGetLastError()));
        return OB200_OK;
    }
    if (h->persistent_baro) {
        if (launch_baro_persistent(h->G, h->F, dtau, h->d_weights, h->cfg.n_substeps, per, ex, h->stream) != 0)
            return h->fail(OB200_ECUDA, "cooperative launch", cudaGetErrorString(cudaGetLastError()));
    } else if (h->use_graph) {
        // the 2 M tiny launches are launch-rate-bound on narrow slabs: replay them as one CUDA graph (re-captured
        // when dt changes; all kernel arguments are baked in)
        if (!h->graph_exec || h->graph_dt != dt) {
            if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }
            cudaGraph_t graph = nullptr;
            const long long before = ob_launch_count;
            OB_CUDA_CHECK(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
            for (int m = 0; m < h->cfg.n_substeps; m++) {
                launch_baro_eta(h->G, h->F, dtau, per, ex, h->stream);
                launch_baro_vel(h->G, h->F, dtau, h->weights[m], per, ex, h->stream);
            }
            launch_baro_finish(h->G, h->F, h->stream);
            OB_CUDA_CHECK(h, cudaStreamEndCapture(h->stream, &graph));
            ob_launch_count = before;   // capture enqueues nothing
            OB_CUDA_CHECK(h, cudaGraphInstantiate(&h->graph_exec, graph, 0));
            cudaGraphDestroy(graph);
            h->graph_dt = dt;
        }
        OB_CUDA_CHECK(h, cudaGraphLaunch(h->graph_exec, h->stream));
        OB_COUNT_LAUNCH(2 * h->cfg.n_substeps + 1);
        return OB200_OK;
    } else {
        for (int m = 0; m < h->cfg.n_substeps; m++) {
            launch_baro_eta(h->G, h->F, dtau, per, ex, h->stream);
            launch_baro_vel(h->G, h->F, dtau, h->weights[m], per, ex, h->stream);
        }
    }
    launch_baro_finish(h->G, h->F, h->stream);
    return OB200_OK;
}

// One time step.  Option "overlap": the T, S solve (HBM-bound, independent of the momentum / barotropic
// chain until the halo exchange) runs on a second stream beside the launch-latency-bound barotropic substeps; joined
// by events before the exchange.  Bit-identical to the serial order; measured without gain on one B200 at C3 (the
// long-lived solver blocks fill every SM, the substep kernels queue behind them), so the default is the serial order.
int step_body_serial(ob200_handle* h, double dt, bool euler);

int step_body(ob200_handle* h, double dt, bool euler) {
    if (!h->overlap) return step_body_serial(h, dt, euler);
    const double chi = euler ? -0.5 : h->cfg.ab2_chi;
    const Geom& G = h->G;
    cudaStream_t s1 = h->stream, s2 = h->stream2;
    if (euler) {
        const size_t n3 = (size_t)G.sz * (G.Nz + 1) * sizeof(real);
        for (real* g : {h->F.Gu0, h->F.Gv0, h->F.GT0, h->F.GS0}) cudaMemsetAsync(g, 0, n3, s1);
    }
    int e = 0;
    auto mark = [&]() { if (h->timing) cudaEventRecord(h->ev[e++], s1); };
    auto mark2 = [&](int q) { if (h->timing) cudaEventRecord(h->ev2[q], s2); };
    mark();
    const bool gab = !euler && h->gab_valid;
    launch_ab2_implicit_uv(G, h->F, dt, chi, h->onchip_solve, gab, s1);   // also produces G^U, G^V (barotropic forcing)
    mark();
    // the T, S solve (HBM-bound) runs beside the barotropic substeps (launch-latency-bound, SMs mostly idle)
    cudaEventRecord(h->ev_join[0], s1); cudaStreamWaitEvent(s2, h->ev_join[0], 0);
    mark2(0);
    launch_ab2_implicit_ts(G, h->F, dt, chi, h->onchip_solve, gab, s2);
    mark2(1);
    const bool ts_early = !h->periodic_local && h->early_ts_exchange && !h->overlap_exchange && !h->fields_dirty;
    if (ts_early) {   // the T, S halos follow their solve on the same side stream
        if (int rc = exchange_halos(h, s2, 2)) return rc;
        cudaEventRecord(h->ev_ts[1], s2);
    }
    cudaEventRecord(h->ev_join[1], s2);
    if (int rc = barotropic_loop(h, dt)) return rc;
    mark();
    launch_correct(G, h->F, s1);
    swap_tendencies(h);
    mark();
    h->time += dt; h->iteration += 1; h->last_dt = dt;
    cudaStreamWaitEvent(s1, h->ev_join[1], 0);   // halos need the new T, S
    if (int rc = mask_and_halos(h, ts_early)) return rc;
    mark();
    launch_w_p(G, h->F, s1);
    mark();
    launch_ri_kappa(G, h->F, h->time, s1);
    if (G.qgleith) launch_leith(G, h->F, s1);
    mark();
    if (h->timing) { tendencies(h, h->ev[e], h->ev2[3]); e++; }   // ev[e]: momentum done; ev2[3]: tracers + masked tiles done
    else tendencies(h);
    h->n_timed = e - 1;
    return OB200_OK;
}

// Default: everything on one stream, in the reference's order (53.1 vs 53.0 ms/step with "overlap" at C3).
int step_body_serial(ob200_handle* h, double dt, bool euler) {
    const double chi = euler ? -0.5 : h->cfg.ab2_chi;
    const Geom& G = h->G;
    cudaStream_t s1 = h->stream;
    if (euler) {
        const size_t n3 = (size_t)G.sz * (G.Nz + 1) * sizeof(real);
        for (real* g : {h->F.Gu0, h->F.Gv0, h->F.GT0, h->F.GS0}) cudaMemsetAsync(g, 0, n3, s1);
    }
    int e = 0;
    auto mark = [&]() { if (h->timing) cudaEventRecord(h->ev[e++], s1); };
    const bool gab = !euler && h->gab_valid;
    if (h->timing) cudaEventRecord(h->ev2[0], s1);
    launch_ab2_implicit_ts(G, h->F, dt, chi, h->onchip_solve, gab, s1);
    if (h->timing) cudaEventRecord(h->ev2[1], s1);
    // set_field since the last step (fields_dirty): T, S still have to be masked before they travel -> regular exchange
    const bool ts_early = !h->periodic_local && h->early_ts_exchange && !h->overlap_exchange && !h->fields_dirty;
    if (ts_early)
        if (int rc = early_ts_exchange(h)) return rc;
    mark();
    launch_ab2_implicit_uv(G, h->F, dt, chi, h->onchip_solve, gab, s1);   // also produces G^U, G^V (barotropic forcing)
    mark();
    if (int rc = barotropic_loop(h, dt)) return rc;
    mark();
    launch_correct(G, h->F, s1);
    swap_tendencies(h);
    mark();
    h->time += dt; h->iteration += 1; h->last_dt = dt;
    if (int rc = halos_and_auxiliaries(h, h->timing ? &h->ev[e] : nullptr, ts_early)) return rc;
    if (h->timing) e += 3;
    if (h->timing) { tendencies(h, h->ev[e], h->ev2[3]); e++; }   // ev[e]: momentum done; ev2[3]: tracers + masked tiles done
    else tendencies(h);
    h->n_timed = e - 1;
    return OB200_OK;
}

}  // namespace

// ==========================================================================================================
extern "C" {

const char* ob200_version(void) { return sizeof(real) == 8 ? "ocean_b200 0.2 (sm_100a, f64)" : "ocean_b200 0.2 (sm_100a, f32)"; }
int32_t ob200_real_bytes(void) { return (int32_t)sizeof(real); }
const char* ob200_last_error(const ob200_handle* h) { return h ? h->err.c_str() : g_last_error.c_str(); }

int ob200_create(const ob200_config* cfg, int32_t device, ob200_handle** out) {
    if (!cfg || !out) { g_last_error = "null argument"; return OB200_EINVAL; }
    if (cfg->struct_bytes != (int32_t)sizeof(ob200_config)) { g_last_error = "ob200_config size mismatch (ABI)"; return OB200_EINVAL; }
    if (cfg->Nx < 2 * OB_H || cfg->Ny < 2 || cfg->Nz < 2 || cfg->Nz > 32767 || cfg->n_substeps < 1 || cfg->n_substeps > OB200_MAX_SUBSTEPS ||
        !cfg->z_faces || !cfg->bottom_height || !cfg->averaging_weights || cfg->nranks < 1 || cfg->bottom_halo_columns < OB_H ||
        cfg->barotropic_exchange < OB200_BAROX_AUTO || cfg->barotropic_exchange > OB200_BAROX_SUBSTEP) {
        g_last_error = "invalid ob200_config (sizes / null arrays; Nx >= 14; Nz <= 32767; bottom_halo_columns >= 7)";
        return OB200_EINVAL;
    }
    const bool wide_ok = cfg->Nx >= cfg->n_substeps + 1 && cfg->bottom_halo_columns >= cfg->n_substeps + 1;
    if (cfg->nranks > 1 && cfg->barotropic_exchange == OB200_BAROX_WIDE && !wide_ok) {
        char msg[256];
        std::snprintf(msg, sizeof msg, "x-slab of %d columns (bathymetry halo %d) is narrower than the wide barotropic halo of n_substeps + 1 = %d "
                      "columns: use barotropic_exchange = OB200_BAROX_SUBSTEP (or AUTO)", cfg->Nx, cfg->bottom_halo_columns, cfg->n_substeps + 1);
        g_last_error = msg;
        return OB200_EINVAL;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_last_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (libocean_b200 has no CPU fallback)";
        return OB200_ECUDA;
    }
    if (device < 0 || device >= ndev) { g_last_error = "bad device ordinal"; return OB200_EINVAL; }
    ob200_handle* h = new ob200_handle();
    h->cfg = *cfg;
    h->device = device;
    cudaSetDevice(device);
    h->weights.assign(cfg->averaging_weights, cfg->averaging_weights + cfg->n_substeps);
    h->periodic_local = cfg->nranks <= 1;
    {
        int mode = cfg->barotropic_exchange;
        const char* e = getenv("OB200_BARO_EXCHANGE");      // tuning / microbenchmark override: "wide" | "substep"
        if (e && !std::strcmp(e, "substep")) mode = OB200_BAROX_SUBSTEP;
        if (e && !std::strcmp(e, "wide") && wide_ok) mode = OB200_BAROX_WIDE;
        h->substep_exchange = cfg->nranks > 1 && (mode == OB200_BAROX_SUBSTEP || (mode == OB200_BAROX_AUTO && !wide_ok));
    }
    int rc = setup_geometry(h);
    if (!rc) rc = alloc_fields(h);
    if (!rc) rc = upload_real(h, h->weights.data(), cfg->n_substeps, &h->d_weights);
    if (!rc) rc = dev_alloc(h, &h->d_diag, 8);
    if (rc) { g_last_error = h->err; ob200_destroy(h); return rc; }
    h->cfg.z_faces = nullptr; h->cfg.bottom_height = nullptr; h->cfg.averaging_weights = nullptr;  // borrowed only during create
    {   // the shared-memory solve is an OPTION: measured slower than the HBM-scratch sweep
        const char* e = getenv("OB200_ONCHIP_SOLVE");
        h->onchip_solve = ab2_onchip_fits(h->G) && (e && e[0] == '1');
    }
    {   // the main stream outranks the side streams: its (small, latency-bound) barotropic kernels get free SM slots first
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, hi);
        cudaStreamCreateWithPriority(&h->stream2, cudaStreamNonBlocking, lo);
    }
    cudaStreamCreateWithFlags(&h->stream3, cudaStreamNonBlocking);
    for (auto& ev : h->ev_tend) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    for (auto& ev : h->ev_ts) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    for (auto& ev : h->ev) cudaEventCreate(&ev);
    for (auto& ev : h->ev2) cudaEventCreate(&ev);
    for (auto& ev : h->ev_join) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) { g_last_error = std::string("setup kernels: ") + cudaGetErrorString(e); ob200_destroy(h); return OB200_ECUDA; }
    *out = h;
    return OB200_OK;
}

void ob200_destroy(ob200_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
    if (h->comm) ncclCommDestroy(h->comm);
    for (void* p : h->allocs) cudaFree(p);
    for (auto& ev : h->ev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : h->ev2) if (ev) cudaEventDestroy(ev);
    for (auto& ev : h->ev_join) if (ev) cudaEventDestroy(ev);
    for (auto& ev : h->ev_tend) if (ev) cudaEventDestroy(ev);
    for (auto& ev : h->ev_ts) if (ev) cudaEventDestroy(ev);
    if (h->stream2) cudaStreamDestroy(h->stream2);
    if (h->stream3) cudaStreamDestroy(h->stream3);
    if (h->stream && h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
}

static int copy_common(ob200_handle* h, int32_t field, real* buf, int on_device, int hx, int hy, int hz, bool to_field,
                       int level = -1) {
    if (!h || !buf) return OB200_EINVAL;
    cudaSetDevice(h->device);
    FieldRef r = field_ref(h, field);
    if (!r.p) return h->fail(OB200_EINVAL, "field", "unknown or unallocated field id");
    if (hx < 0 || hy < 0 || hz < 0) return h->fail(OB200_EINVAL, "halo", "negative halo");
    const Geom& G = h->G;
    if (level >= 0) {   // one horizontal level of a 3-D field, as a (1, ny + 2 hy, Nx + 2 hx) window
        if (r.layout != 0 || level >= r.nz) return h->fail(OB200_EINVAL, "level", "not a level of a 3-D field");
        r.p += (size_t)level * G.sz; r.nz = 1; hz = 0;
    }
    const int ex = std::min(std::min(h->transfer_x_halo, hx), OB_H);
    if (on_device) {
        launch_copy_field(G, r.p, buf, r.ny, r.nz, hx, hy, hz, r.layout, to_field, ex, h->stream);
    } else {
        // host buffer: ONE strided DMA between the dense parent-array window and the pitched device layout (no staging
        // copy, no staging allocation); pinned host memory moves at the PCIe rate
        const int pitch = r.layout == 2 ? G.pitch2 : G.pitch, x0 = r.layout == 2 ? G.X0b : OB_X0;
        const size_t sx = (size_t)G.Nx + 2 * hx, sy = (size_t)r.ny + 2 * hy;
        cudaMemcpy3DParms p = {};
        const cudaPitchedPtr host = make_cudaPitchedPtr(buf, sx * sizeof(real), sx, sy);
        const cudaPitchedPtr dev = make_cudaPitchedPtr(r.p, (size_t)pitch * sizeof(real), pitch, G.NyP);
        const cudaPos hpos = make_cudaPos((size_t)(hx - ex) * sizeof(real), hy, r.layout == 0 ? hz : 0);
        const cudaPos dpos = make_cudaPos((size_t)(x0 - ex) * sizeof(real), OB_H, 0);
        p.srcPtr = to_field ? host : dev; p.srcPos = to_field ? hpos : dpos;
        p.dstPtr = to_field ? dev : host; p.dstPos = to_field ? dpos : hpos;
        p.extent = make_cudaExtent((size_t)(G.Nx + 2 * ex) * sizeof(real), r.ny, r.nz);
        p.kind = to_field ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
        OB_CUDA_CHECK(h, cudaMemcpy3DAsync(&p, h->stream));
        OB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
    }
    OB_CUDA_CHECK(h, cudaGetLastError());
    if (to_field) { h->fields_dirty = true; h->gab_valid = false; }   // the caller may have written G^n or G^-
    return OB200_OK;
}
int ob200_set_field(ob200_handle* h, int32_t field, const ob200_real* buf, int32_t on_device, int32_t hx, int32_t hy, int32_t hz) {
    return copy_common(h, field, const_cast<real*>(buf), on_device, hx, hy, hz, true);
}
int ob200_get_field(ob200_handle* h, int32_t field, ob200_real* buf, int32_t on_device, int32_t hx, int32_t hy, int32_t hz) {
    return copy_common(h, field, buf, on_device, hx, hy, hz, false);
}

int ob200_bind(ob200_handle* h, int32_t field, ob200_real* dev_ptr, int32_t hx, int32_t hy, int32_t hz) {
    if (!h) return OB200_EINVAL;
    if (field < 0 || field >= OB200_F_COUNT || !field_ref(h, field).p) return h->fail(OB200_EINVAL, "bind", "unknown or unallocated field id");
    if (!dev_ptr || hx < 0 || hy < 0 || hz < 0) return h->fail(OB200_EINVAL, "bind", "null pointer or negative halo");
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, dev_ptr) != cudaSuccess || (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged)) {
        cudaGetLastError();
        return h->fail(OB200_EINVAL, "bind", "not a device pointer");
    }
    h->bound[field].ptr = dev_ptr; h->bound[field].hx = hx; h->bound[field].hy = hy; h->bound[field].hz = hz;
    return OB200_OK;
}
int ob200_unbind(ob200_handle* h, int32_t field) {
    if (!h || field < 0 || field >= OB200_F_COUNT) return OB200_EINVAL;
    h->bound[field] = ob200_handle::Bound();
    return OB200_OK;
}
int ob200_sync_outputs(ob200_handle* h) {
    if (!h) return OB200_EINVAL;
    for (int f = 0; f < OB200_F_COUNT; f++)
        if (h->bound[f].ptr)
            if (int rc = copy_common(h, f, h->bound[f].ptr, 1, h->bound[f].hx, h->bound[f].hy, h->bound[f].hz, false)) return rc;
    OB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
    return OB200_OK;
}
int ob200_sync_inputs(ob200_handle* h) {
    if (!h) return OB200_EINVAL;
    bool any = false;
    for (int f = 0; f < OB200_F_COUNT; f++)
        if (h->bound[f].ptr) {
            if (int rc = copy_common(h, f, h->bound[f].ptr, 1, h->bound[f].hx, h->bound[f].hy, h->bound[f].hz, true)) return rc;
            any = true;
        }
    return any ? ob200_update_state(h) : OB200_OK;
}

int ob200_get_level(ob200_handle* h, int32_t field, int32_t k, ob200_real* buf, int32_t on_device, int32_t hx, int32_t hy) {
    return copy_common(h, field, buf, on_device, hx, hy, 0, false, k);
}

int ob200_set_forcing(ob200_handle* h, int32_t which, const ob200_real* buf, int32_t on_device) {
    if (!h || !buf || which < 0 || which > 5) return OB200_EINVAL;
    cudaSetDevice(h->device);
    const size_t n = (size_t)h->G.Nx * (which == 1 ? h->G.Ny + 1 : h->G.Ny) * 6;
    if (!h->F.forcing[which])
        if (int rc = dev_alloc(h, &h->F.forcing[which], n, false)) return rc;
    OB_CUDA_CHECK(h, cudaMemcpyAsync(h->F.forcing[which], buf, n * sizeof(real),
                                     on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
    OB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
    if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }
    return OB200_OK;
}

int ob200_update_state(ob200_handle* h) {
    if (!h) return OB200_EINVAL;
    cudaSetDevice(h->device);
    if (int rc = update_state(h)) return rc;
    OB_CUDA_CHECK(h, cudaGetLastError());
    return OB200_OK;
}

int ob200_time_step(ob200_handle* h, double dt) {
    if (!h) return OB200_EINVAL;
    cudaSetDevice(h->device);
    nvtxRangePushA("one time step");   // same range name as near_global_simulation.jl:195
    int rc = OB200_OK;
    if (h->iteration == 0) rc = update_state(h);
    const bool euler = (dt != h->last_dt);
    if (!rc) rc = step_body(h, dt, euler);
    nvtxRangePop();
    if (rc) return rc;
    OB_CUDA_CHECK(h, cudaGetLastError());
    return OB200_OK;
}

int ob200_time_steps(ob200_handle* h, double dt, int32_t n) {
    for (int s = 0; s < n; s++)
        if (int rc = ob200_time_step(h, dt)) return rc;
    return OB200_OK;
}

int ob200_run_stage(ob200_handle* h, int32_t stage, double dt, int32_t arg) {
    if (!h) return OB200_EINVAL;
    cudaSetDevice(h->device);
    const double chi = (dt != h->last_dt) ? -0.5 : h->cfg.ab2_chi;
    const double dtau = h->cfg.frac_dtau * dt;
    int rc = OB200_OK;
    switch (stage) {
        case OB200_ST_BAROTROPIC_FORCING: launch_barotropic_forcing(h->G, h->F, chi, h->stream); break;
        case OB200_ST_AB2_IMPLICIT:
            launch_ab2_implicit_uv(h->G, h->F, dt, chi, h->onchip_solve, dt == h->last_dt && h->gab_valid, h->stream);
            launch_ab2_implicit_ts(h->G, h->F, dt, chi, h->onchip_solve, dt == h->last_dt && h->gab_valid, h->stream);
            break;
        case OB200_ST_BAROTROPIC_INIT: launch_baro_init(h->G, h->F, h->periodic_local, 0, h->stream); break;
        case OB200_ST_BAROTROPIC_ETA: launch_baro_eta(h->G, h->F, dtau, h->periodic_local, 0, h->stream); break;
        case OB200_ST_BAROTROPIC_VEL:
            if (arg < 0 || arg >= h->cfg.n_substeps) return h->fail(OB200_EINVAL, "substep", "index out of range");
            launch_baro_vel(h->G, h->F, dtau, h->weights[arg], h->periodic_local, 0, h->stream);
            break;
        case OB200_ST_BAROTROPIC_FINISH: launch_baro_finish(h->G, h->F, h->stream); break;
        case OB200_ST_CORRECT_AND_STORE: launch_correct(h->G, h->F, h->stream); swap_tendencies(h); break;
        case OB200_ST_MASK_AND_LOCAL_HALOS:
            if (h->fields_dirty) { launch_mask_fields(h->G, h->F, h->stream); h->fields_dirty = false; }
            launch_fill_halos(h->G, h->F, h->periodic_local, h->stream);
            break;
        case OB200_ST_AUXILIARIES: auxiliaries(h); break;
        case OB200_ST_TENDENCIES: tendencies(h); break;
        case OB200_ST_BAROTROPIC_LOOP: rc = barotropic_loop(h, dt); break;
        case OB200_ST_EXCHANGE_HALOS: rc = exchange_halos(h); break;
        default: return h->fail(OB200_EINVAL, "stage", "unknown stage id");
    }
    if (rc) return rc;
    OB_CUDA_CHECK(h, cudaGetLastError());
    return OB200_OK;
}

int ob200_sync(ob200_handle* h) {
    if (!h) return OB200_EINVAL;
    cudaSetDevice(h->device);
    OB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
    return OB200_OK;
}

int ob200_get_clock(ob200_handle* h, double* time, int64_t* iteration, double* last_dt) {
    if (!h) return OB200_EINVAL;
    if (time) *time = h->time;
    if (iteration) *iteration = h->iteration;
    if (last_dt) *last_dt = h->last_dt;
    return OB200_OK;
}
int ob200_set_clock(ob200_handle* h, double time, int64_t iteration, double last_dt) {
    if (!h) return OB200_EINVAL;
    h->time = time; h->iteration = iteration; h->last_dt = last_dt;
    return OB200_OK;
}

int ob200_diagnostics(ob200_handle* h, double maxabs[6], double* cfl_dt) {
    if (!h) return OB200_EINVAL;
    cudaSetDevice(h->device);
    launch_diagnostics(h->G, h->F, h->d_diag, h->stream);
    double out[8];
    OB_CUDA_CHECK(h, cudaMemcpyAsync(out, h->d_diag, sizeof(out), cudaMemcpyDeviceToHost, h->stream));
    OB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
    if (maxabs) for (int q = 0; q < 6; q++) maxabs[q] = out[q];
    if (cfl_dt) *cfl_dt = out[6] != out[6] ? NAN : (out[6] > 0.0 ? 1.0 / out[6] : INFINITY);
    return OB200_OK;
}

int ob200_get_timing(ob200_handle* h, ob200_timing* t) {
    if (!h || !t) return OB200_EINVAL;
    cudaSetDevice(h->device);
    t->n = 0;
    t->launches = ob_launch_count;
    if (!h->timing || h->n_timed <= 0) return OB200_OK;
    OB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
    // stream-1 marks: ab2(u,v) | barotropic | corrector | halos | w+Ri | pHY'+kappa | momentum ; stream 2: T,S solve, tracers
    float g[8] = {0};
    for (int q = 0; q < 7 && q < h->n_timed; q++) cudaEventElapsedTime(&g[q], h->ev[q], h->ev[q + 1]);
    float ts = 0, tr = 0;
    cudaEventElapsedTime(&ts, h->ev2[0], h->ev2[1]);
    cudaEventElapsedTime(&tr, h->ev[7], h->ev2[3]);
    t->ms[0] = g[0] + ts; t->ms[1] = g[1]; t->ms[2] = g[2]; t->ms[3] = g[3]; t->ms[4] = g[4]; t->ms[5] = g[5];
    t->ms[6] = g[6]; t->ms[7] = tr; t->ms[8] = ts;
    t->n = 9;
    return OB200_OK;
}

int ob200_set_stream(ob200_handle* h, void* cuda_stream) {
    if (!h) return OB200_EINVAL;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    h->stream = (cudaStream_t)cuda_stream;
    h->own_stream = false;
    return OB200_OK;
}

int ob200_set_option(ob200_handle* h, const char* name, int32_t value) {
    if (!h || !name) return OB200_EINVAL;
    if (!std::strcmp(name, "timing")) h->timing = value != 0;
    else if (!std::strcmp(name, "persistent_barotropic")) h->persistent_baro = value != 0;
    else if (!std::strcmp(name, "onchip_barotropic")) h->onchip_baro = value != 0;
    else if (!std::strcmp(name, "overlap")) h->overlap = value != 0;
    else if (!std::strcmp(name, "barotropic_graph")) h->use_graph = value != 0;
    else if (!std::strcmp(name, "overlap_exchange")) h->overlap_exchange = value != 0;
    else if (!std::strcmp(name, "concurrent_masked_tiles")) h->concurrent_masked = value != 0;
    else if (!std::strcmp(name, "early_ts_exchange")) h->early_ts_exchange = value != 0;
    else if (!std::strcmp(name, "onchip_solve")) {
        if (value && !ab2_onchip_fits(h->G)) return h->fail(OB200_EINVAL, "onchip_solve", "Nz too large for the shared-memory solve");
        h->onchip_solve = value != 0;
    }
    else if (!std::strcmp(name, "transfer_x_halo")) h->transfer_x_halo = value < 0 ? 0 : value;
    else return h->fail(OB200_EINVAL, "option", name);
    return OB200_OK;
}

int ob200_comm_unique_id(void* id128) {
    if (!id128) return OB200_EINVAL;
    static_assert(sizeof(ncclUniqueId) <= 128, "unique id size");
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) { g_last_error = "ncclGetUniqueId failed"; return OB200_ENCCL; }
    std::memset(id128, 0, 128);
    std::memcpy(id128, &id, sizeof(id));
    return OB200_OK;
}

int ob200_comm_init(ob200_handle* h, const void* id128) {
    if (!h || !id128) return OB200_EINVAL;
    cudaSetDevice(h->device);
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    ncclResult_t rc = ncclCommInitRank(&h->comm, h->cfg.nranks, id, h->cfg.rank);
    if (rc != ncclSuccess) return h->fail(OB200_ENCCL, "ncclCommInitRank", ncclGetErrorString(rc));
    for (int g = 0; g < 2; g++) {
        h->xcount[g] = pack_x_count(h->G, g);
        for (real** p : {&h->sendW[g], &h->sendE[g], &h->recvW[g], &h->recvE[g]})
            if (int r2 = dev_alloc(h, p, h->xcount[g])) return r2;
    }
    h->bcount = pack_baro_count(h->G, h->G.Wb);
    for (real** p : {&h->bsendW, &h->bsendE, &h->brecvW, &h->brecvE})
        if (int r2 = dev_alloc(h, p, h->bcount)) return r2;
    for (real** p : {&h->e_send_eta, &h->e_recv_eta, &h->e_send_U, &h->e_recv_U})
        if (int r2 = dev_alloc(h, p, (size_t)h->G.Ny + 1)) return r2;
    return OB200_OK;
}

}  // extern "C"
