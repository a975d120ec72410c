// Host-callable launchers of the CUDA kernels (one translation unit per kernel family).
#pragma once
#include <cuda.h>   // CUtensorMap (driver types only; the encoder is fetched with cudaGetDriverEntryPoint)

#include "ob_common.cuh"

extern long long ob_launch_count;   // kernels launched by this library (bench.py reports it as gpu_launches)
#define OB_COUNT_LAUNCH(n) (ob_launch_count += (n))

struct Fields {
    // 3-D (Nz+1 planes)
    real *u, *v, *w, *T, *S, *p, *kc, *ku, *Ri, *scratch, *scratch2;
    real *Gu, *Gv, *GT, *GS, *Gu0, *Gv0, *GT0, *GS0;
    // OPTION (OB200_AB2_COMBINED=1, measured slower overall): the AB2 combination (3/2 + chi) G^n - (1/2 + chi) G^- of u, v, T, S,
    // written by the tendency kernels beside G^n so that the HBM-bound implicit solve reads ONE tendency array per field
    // instead of two; null = not allocated
    real *Gab[4];
    real *nuL, *qx, *qy;
    // 2-D
    real *eta, *U, *V, *etab, *Ub, *Vb, *GU, *GV, *Hfc, *Hcf, *Ld;
    real *Usum, *Vsum;   // sum_k dz u, sum_k dz v of the freshly solved u, v (barotropic layout), produced by the implicit solve
    // RealisticOcean forcing (Nx x Ny(+1) x 6, compact), may be null
    real* forcing[6];
};

// kernels_halo.cu
void launch_fill_halos(const Geom& G, const Fields& F, bool periodic_x, cudaStream_t st);
void launch_mask_fields(const Geom& G, const Fields& F, cudaStream_t st);
// layout: 0 = 3-D field, 1 = 2-D in the 3-D plane layout (Ld), 2 = 2-D barotropic layout
void launch_copy_field(const Geom& G, real* field, real* buf, int ny, int nz, int hx, int hy, int hz, int layout,
                       bool to_field, int ex, cudaStream_t st);
// exchange groups of the 7-column x-halos: u, v, eta (final after the corrector) | T, S (final after their implicit solve)
#define OB_XG_UVETA 0
#define OB_XG_TS 1
void launch_pack_x(const Geom& G, const Fields& F, real* sendW, real* sendE, int group, cudaStream_t st);
void launch_unpack_x(const Geom& G, const Fields& F, const real* recvW, const real* recvE, int group, cudaStream_t st);
size_t pack_x_count(const Geom& G, int group);
void launch_diagnostics(const Geom& G, const Fields& F, double* out8, cudaStream_t st); /*f64*/

// kernels_column.cu
void launch_w_p(const Geom& G, const Fields& F, cudaStream_t st);                       // all columns [-2, Nx + 2)
void launch_ri_kappa(const Geom& G, const Fields& F, double time, cudaStream_t st);
// the same sweeps on the columns [a0, a0 + n0) U [a1, a1 + n1)
void launch_w_p(const Geom& G, const Fields& F, int a0, int n0, int a1, int n1, cudaStream_t st);
void launch_ri_kappa(const Geom& G, const Fields& F, double time, int a0, int n0, int a1, int n1, cudaStream_t st);
void launch_barotropic_forcing(const Geom& G, const Fields& F, real chi, cudaStream_t st);
// momentum (u, v: also G^U, G^V) and tracer (T, S) solves can run on different streams (separate scratch arrays)
// `onchip`: eliminated coefficients and forward values in shared memory (every array crosses HBM once); needs
// ab2_onchip_fits(G), otherwise the variant with an HBM scratch array.  Same arithmetic, same bits.
bool ab2_onchip_fits(const Geom& G);
// `gab`: F.Gab holds the AB2 combination of the current G^n, G^- for the configured chi (valid, regular AB2 step)
void launch_ab2_implicit_uv(const Geom& G, const Fields& F, real dt, real chi, bool onchip, bool gab, cudaStream_t st);
void launch_ab2_implicit_ts(const Geom& G, const Fields& F, real dt, real chi, bool onchip, bool gab, cudaStream_t st);
void launch_correct(const Geom& G, const Fields& F, cudaStream_t st);

// kernels_baro.cu
// `ex` = columns beyond the interior computed on each side (0 on one GPU, Wb - 1 on x-slabs)
void launch_baro_init(const Geom& G, const Fields& F, bool periodic_x, int ex, cudaStream_t st);
void launch_baro_eta(const Geom& G, const Fields& F, real dtau, bool periodic_x, int ex, cudaStream_t st);
void launch_baro_vel(const Geom& G, const Fields& F, real dtau, real wgt, bool periodic_x, int ex, cudaStream_t st);
void launch_baro_finish(const Geom& G, const Fields& F, cudaStream_t st);
// per-substep slab exchange: interior columns only; edge columns through contiguous buffers of Ny doubles (kernels_baro.cu)
void launch_baro_eta_edge(const Geom& G, const Fields& F, real dtau, real* send_eta, const real* recv_eta, real* send_U,
                          const real* recv_U, cudaStream_t st);
void launch_baro_vel_edge(const Geom& G, const Fields& F, real dtau, real wgt, real* send_eta, const real* recv_eta,
                          real* send_U, const real* recv_U, cudaStream_t st);
void launch_baro_edge_first(const Geom& G, const Fields& F, real* send_U, cudaStream_t st);
int launch_baro_persistent(const Geom& G, const Fields& F, real dtau, const real* weights_dev, int M, bool periodic_x,
                           int ex, cudaStream_t st);
// persistent ON-CHIP variant: eta, U, V resident in shared memory for all M substeps (returns -2 when the slab does not fit)
bool baro_onchip_fits(const Geom& G, int ex);
int launch_baro_onchip(const Geom& G, const Fields& F, real dtau, const real* weights_dev, int M, bool periodic_x, int ex,
                       cudaStream_t st);
size_t pack_baro_count(const Geom& G, int W);
void launch_pack_baro(const Geom& G, const Fields& F, real* bufW, real* bufE, int W, bool unpack, cudaStream_t st);

// kernels_tend.cu
// Per-tile plan of the momentum kernels: tile_kfast[t] = first level the shared-memory fast kernel may take for
// tile t (max of kfast over the tile's columns; Nz = never); slow_tiles = (tile, kfast, kmin) triples with kfast > 0,
// handled on the levels [kmin, kfast) by the general masked kernel (kmin = first level with an active cell in the tile;
// below it everything is a masked zero that is never written).
struct MomPlan {
    int ntx = 0, nty = 0, nslow = 0;
    bool any_fast = false;
    int* tile_kfast = nullptr;   // device, ntx * nty
    int* slow_tiles = nullptr;   // device, 3 * nslow: (tile, kfast, first level with an active cell in the tile)
    // TMA descriptors (cp.async.bulk.tensor) of the 3-D fields u, v (momentum tile box) and T, S (tracer tile box)
    bool use_tma = false;
    CUtensorMap tm_u, tm_v, tm_T, tm_S;
};
// tile boxes (elements) the tendency kernels fetch with one TMA copy per field and level
void tend_tile_boxes(int* mom_x, int* mom_y, int* trc_x, int* trc_y);
void mom_plan_tile_shape(int* tx, int* ty);
// `which`: 1 = full-order tiles, 2 = masked tiles, 3 = both (they write disjoint (tile, level) sets)
void launch_momentum(const Geom& G, const Fields& F, const MomPlan& P, int ord_vort, int ord_div, double time, cudaStream_t st, int which = 3);
void launch_tracers(const Geom& G, const Fields& F, const MomPlan& P, int ord_tracer, double time, cudaStream_t st, int which = 3);
// kernels_leith.cu
void launch_leith(const Geom& G, const Fields& F, cudaStream_t st);
