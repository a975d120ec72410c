// Split-explicit free surface: barotropic substeps (forward-backward scheme).
//   eta -= dtau * div(U, V) / Az ;  U, V += dtau * (-g H d(eta) + G^{U,V}) ; averages += w_m * (eta, U, V)
// One GPU: the producing kernel also maintains the periodic x-halo column its consumer needs (eta at i = -1,
// U at i = Nx), so there are no halo launches inside the loop.  Several GPUs: the barotropic arrays carry a
// halo of Wb = M + 1 columns that is exchanged ONCE before the loop (as Oceananigans' distributed solver does,
// SURVEY F9); every substep is computed over [-ex, Nx + ex) and the invalid fringe, which grows by one column
// per substep, never reaches the interior.
// launch_baro_persistent: one cooperative launch for all M substeps (grid-wide barrier between half steps).
#include <algorithm>
#include <utility>
#include <vector>

#include <cooperative_groups.h>

#include "kernels.h"
namespace cg = cooperative_groups;

__global__ void k_baro_init(Geom G, Fields F, bool periodic, int ex) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x - ex;
    const int j = blockIdx.y;  // 0..Ny
    if (i >= G.Nx + ex) return;
    const int c = idx2b(G, i, j);
    const real a = F.Ub[c], b = F.Vb[c];
    F.U[c] = a;
    F.V[c] = b;
    if (periodic) {   // the first eta substep reads U at the east halo column
        if (i < OB_H) { F.U[c + G.Nx] = a; F.V[c + G.Nx] = b; }
        if (i >= G.Nx - OB_H) { F.U[c - G.Nx] = a; F.V[c - G.Nx] = b; }
    }
    F.etab[c] = RL(0.0); F.Ub[c] = RL(0.0); F.Vb[c] = RL(0.0);
}

// Per-substep slab exchange (x-slabs narrower than the wide halo, and the microbenchmark of BASELINE configs[4]): the one
// column a neighbour needs travels through contiguous edge buffers -- the producer writes its edge value there as well,
// the consumer reads the neighbour's from there -- so there are no pack / unpack launches between the NCCL calls.
struct BaroEdge {
    real* send_eta;         // my eta(Nx - 1, j)      -> east neighbour's eta(-1, j)
    const real* recv_eta;   // west neighbour's eta(Nx - 1, j)
    real* send_U;           // my U(0, j)             -> west neighbour's U(Nx, j)
    const real* recv_U;     // east neighbour's U(0, j)
};

__device__ __forceinline__ void eta_update(const Geom& G, const Fields& F, int i, int j, real dtau, bool periodic,
                                           const BaroEdge* E = nullptr) {
    const int c = idx2b(G, i, j);
    const real Vn = (j + 1 == G.Ny) ? RL(0.0) : F.V[c + G.pitch2];
    const real Vs = (j == 0) ? RL(0.0) : F.V[c];
    const real Ue = (E && i == G.Nx - 1) ? E->recv_U[j] : F.U[c + 1];
    const real e = F.eta[c] - dtau * ((G.dy * Ue - G.dy * F.U[c]) + (G.dxcf[j + 1] * Vn - G.dxcf[j] * Vs)) / G.azcc[j];
    F.eta[c] = e;
    if (periodic) {
        if (i == G.Nx - 1) F.eta[c - G.Nx] = e;   // west halo column i = -1
        if (i == 0) F.eta[c + G.Nx] = e;          // east halo column i = Nx
    }
    if (E && i == G.Nx - 1) E->send_eta[j] = e;
}

__device__ __forceinline__ void vel_update(const Geom& G, const Fields& F, int i, int j, real dtau, real wgt,
                                           bool periodic, const BaroEdge* E = nullptr) {
    const int c = idx2b(G, i, j);
    const real e = F.eta[c];
    const real ew = (E && i == 0) ? E->recv_eta[j] : F.eta[c - 1];
    const real detax = (e - ew) / G.dxfc[j];
    const real detay = (j == 0) ? RL(0.0) : (e - F.eta[c - G.pitch2]) / G.dy;
    const real Un = F.U[c] + dtau * (-G.g * F.Hfc[c] * detax + F.GU[c]);
    const real Vn = F.V[c] + dtau * (-G.g * F.Hcf[c] * detay + F.GV[c]);
    F.U[c] = Un; F.V[c] = Vn;
    F.etab[c] += wgt * e;
    F.Ub[c] += wgt * Un;
    F.Vb[c] += wgt * Vn;
    if (periodic) {
        if (i == 0) { F.U[c + G.Nx] = Un; F.V[c + G.Nx] = Vn; }
        if (i == G.Nx - 1) { F.U[c - G.Nx] = Un; F.V[c - G.Nx] = Vn; }
    }
    if (E && i == 0) E->send_U[j] = Un;
}

__global__ void __launch_bounds__(128) k_baro_eta(Geom G, Fields F, real dtau, bool periodic, int ex) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x - ex;
    if (i >= G.Nx + ex) return;
    eta_update(G, F, i, blockIdx.y, dtau, periodic);
}
__global__ void __launch_bounds__(128) k_baro_vel(Geom G, Fields F, real dtau, real wgt, bool periodic, int ex) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x - ex;
    if (i >= G.Nx + ex) return;
    vel_update(G, F, i, blockIdx.y, dtau, wgt, periodic);
}
__global__ void __launch_bounds__(128) k_baro_eta_edge(Geom G, Fields F, real dtau, BaroEdge E) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G.Nx) return;
    eta_update(G, F, i, blockIdx.y, dtau, false, &E);
}
__global__ void __launch_bounds__(128) k_baro_vel_edge(Geom G, Fields F, real dtau, real wgt, BaroEdge E) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G.Nx) return;
    vel_update(G, F, i, blockIdx.y, dtau, wgt, false, &E);
}
// after baro_init: the first eta substep needs U(Nx) = the east neighbour's U(0)
__global__ void k_baro_edge_first(Geom G, Fields F, BaroEdge E) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < G.Ny) E.send_U[j] = F.U[idx2b(G, 0, j)];
}
__global__ void k_baro_finish(Geom G, Fields F) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i >= G.Nx) return;
    const int c = idx2b(G, i, j);
    F.eta[c] = F.etab[c];
}

void launch_baro_init(const Geom& G, const Fields& F, bool periodic, int ex, cudaStream_t st) {
    dim3 b(128), g((G.Nx + 2 * ex + 127) / 128, G.Ny + 1);
    k_baro_init<<<g, b, 0, st>>>(G, F, periodic, ex);
    OB_COUNT_LAUNCH(1);
}
void launch_baro_eta(const Geom& G, const Fields& F, real dtau, bool periodic, int ex, cudaStream_t st) {
    dim3 b(128), g((G.Nx + 2 * ex + 127) / 128, G.Ny);
    k_baro_eta<<<g, b, 0, st>>>(G, F, dtau, periodic, ex);
    OB_COUNT_LAUNCH(1);
}
void launch_baro_vel(const Geom& G, const Fields& F, real dtau, real wgt, bool periodic, int ex, cudaStream_t st) {
    dim3 b(128), g((G.Nx + 2 * ex + 127) / 128, G.Ny);
    k_baro_vel<<<g, b, 0, st>>>(G, F, dtau, wgt, periodic, ex);
    OB_COUNT_LAUNCH(1);
}
void launch_baro_eta_edge(const Geom& G, const Fields& F, real dtau, real* send_eta, const real* recv_eta, real* send_U,
                          const real* recv_U, cudaStream_t st) {
    dim3 b(128), g((G.Nx + 127) / 128, G.Ny);
    k_baro_eta_edge<<<g, b, 0, st>>>(G, F, dtau, BaroEdge{send_eta, recv_eta, send_U, recv_U});
    OB_COUNT_LAUNCH(1);
}
void launch_baro_vel_edge(const Geom& G, const Fields& F, real dtau, real wgt, real* send_eta, const real* recv_eta,
                          real* send_U, const real* recv_U, cudaStream_t st) {
    dim3 b(128), g((G.Nx + 127) / 128, G.Ny);
    k_baro_vel_edge<<<g, b, 0, st>>>(G, F, dtau, wgt, BaroEdge{send_eta, recv_eta, send_U, recv_U});
    OB_COUNT_LAUNCH(1);
}
void launch_baro_edge_first(const Geom& G, const Fields& F, real* send_U, cudaStream_t st) {
    k_baro_edge_first<<<(G.Ny + 127) / 128, 128, 0, st>>>(G, F, BaroEdge{nullptr, nullptr, send_U, nullptr});
    OB_COUNT_LAUNCH(1);
}
void launch_baro_finish(const Geom& G, const Fields& F, cudaStream_t st) {
    dim3 b(128), g((G.Nx + 127) / 128, G.Ny);
    k_baro_finish<<<g, b, 0, st>>>(G, F);
    OB_COUNT_LAUNCH(1);
}

// ---- persistent variant: all substeps in one cooperative launch --------------------------------------
__global__ void __launch_bounds__(256) k_baro_persistent(Geom G, Fields F, real dtau, const real* __restrict__ wts,
                                                          int M, bool periodic, int ex) {
    // rows are dealt to warps, columns to lanes: no integer division in the substep loop, coalesced rows
    cg::grid_group grid = cg::this_grid();
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int xchunks = (G.Nx + 2 * ex + 255) / 256;           // a warp sweeps 256 columns of one row at a time
    const int nitems = G.Ny * xchunks;
    for (int m = 0; m < M; m++) {
        for (int it = warp; it < nitems; it += nwarps) {
            const int j = it / xchunks, i0 = (it - j * xchunks) * 256 - ex;
            for (int i = i0 + lane; i < min(i0 + 256, G.Nx + ex); i += 32) eta_update(G, F, i, j, dtau, periodic);
        }
        grid.sync();
        const real wgt = wts[m];
        for (int it = warp; it < nitems; it += nwarps) {
            const int j = it / xchunks, i0 = (it - j * xchunks) * 256 - ex;
            for (int i = i0 + lane; i < min(i0 + 256, G.Nx + ex); i += 32) vel_update(G, F, i, j, dtau, wgt, periodic);
        }
        grid.sync();
    }
}

int launch_baro_persistent(const Geom& G, const Fields& F, real dtau, const real* wts, int M, bool periodic, int ex,
                           cudaStream_t st) {
    static int nblk = 0;
    if (nblk == 0) {
        int dev = 0, sms = 0, per = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_baro_persistent, 256, 0);
        nblk = sms * per;
    }
    Geom g = G; Fields f = F;
    void* args[] = {&g, &f, &dtau, &wts, &M, &periodic, &ex};
    cudaError_t e = cudaLaunchCooperativeKernel((void*)k_baro_persistent, dim3(nblk), dim3(256), args, 0, st);
    OB_COUNT_LAUNCH(1);
    return e == cudaSuccess ? 0 : -1;
}

// ---- wide-halo exchange buffers: [Ubar | Vbar | eta | GU | GV] x (Ny + 1) rows x W columns --------------------
__global__ void k_pack_baro(Geom G, Fields F, real* bufW, real* bufE, int W, bool unpack) {
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y, q = blockIdx.z;
    if (h >= W) return;
    real* fs[5] = {F.Ub, F.Vb, F.eta, F.GU, F.GV};
    real* f = fs[q];
    const size_t b = ((size_t)q * (G.Ny + 1) + j) * W + h;
    const int c = idx2b(G, 0, j);
    if (!unpack) { bufW[b] = f[c + h]; bufE[b] = f[c + G.Nx - W + h]; }
    else { f[c - W + h] = bufW[b]; f[c + G.Nx + h] = bufE[b]; }
}
size_t pack_baro_count(const Geom& G, int W) { return (size_t)5 * (G.Ny + 1) * W; }
void launch_pack_baro(const Geom& G, const Fields& F, real* bufW, real* bufE, int W, bool unpack, cudaStream_t st) {
    dim3 b(64), g((W + 63) / 64, G.Ny + 1, 5);
    k_pack_baro<<<g, b, 0, st>>>(G, F, bufW, bufE, W, unpack);
    OB_COUNT_LAUNCH(1);
}

// ==========================================================================================================
// Persistent ON-CHIP variant (BASELINE north_star (4)): the whole substep loop in ONE cooperative launch with the
// stencil-coupled state (eta, U, V) of the slab RESIDENT IN SHARED MEMORY across all M substeps and the averages
// (eta_bar, U_bar, V_bar) in registers; global memory is touched once at the start and once at the end, plus one halo
// row per block and half step.  Decomposition: block b owns the rows [b R, (b + 1) R) at full slab width, so all
// x-neighbours are on chip; the two y-neighbours a block needs (eta of the row below its first, V of the row above its
// last) are published by their owners in global memory, made visible by ONE grid-wide barrier per half step, and
// staged into the block's halo rows with a bulk async copy (cp.async.bulk, mbarrier completion).
// Fits when (3 R + 2) rows x (Nx + 2 ex + 2) doubles <= 227 KB with R = ceil(Ny / #SMs): C3 slabs at 2, 4, 8 GPUs, not
// the single-GPU grid (1.94 M columns x 24 B = 47 MB > 148 x 227 KB).  Same operations per column as k_baro_eta / _vel.
// ==========================================================================================================
__device__ __forceinline__ unsigned bsm_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_row_load(real* dst, const real* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bsm_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(bsm_u32(dst)), "l"(src), "r"(bytes), "r"(bsm_u32(bar)) : "memory");
}
__device__ __forceinline__ void bar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "OBB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra OBB_DONE;\n"
        "bra OBB_WAIT;\n"
        "OBB_DONE:\n"
        "}\n" ::"r"(bsm_u32(bar)), "r"(parity) : "memory");
}

template <int MAXC>
__global__ void __launch_bounds__(1024, 1) k_baro_onchip(Geom G, Fields F, real dtau, const real* __restrict__ wts, int M,
                                                         bool periodic, int ex, int R, int Wp, int a0) {
    // Wp = padded row length in shared memory (even, covers array columns [a0, a0 + Wp) of a barotropic row: the slab's
    // columns [-ex - 1, Nx + ex] plus alignment padding, a0 even so that bulk copies are 16-byte aligned)
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(128) real bsm[];
    __shared__ unsigned long long bar[2];
    const int tid = threadIdx.x, NT = blockDim.x;
    const int j0 = blockIdx.x * R, j1 = min(j0 + R, G.Ny), nr = max(j1 - j0, 0);
    real* sE = bsm;                       // (R + 1) rows: row 0 = eta(j0 - 1), row 1 + r = eta(j0 + r)
    real* sU = sE + (size_t)(R + 1) * Wp;  // R rows
    real* sV = sU + (size_t)R * Wp;        // (R + 1) rows: row r = V(j0 + r), row R(=nr) = V(j1) halo
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bsm_u32(&bar[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bsm_u32(&bar[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int ncol = G.Nx + 2 * ex;          // columns updated per row: i in [-ex, Nx + ex)
    const int nwork = nr * ncol;
    const int xo = G.X0b - a0;               // shared-memory column of array column X0b (i = 0)
    // load the block's rows (+ the two halo rows) once
    for (int r = -1; r <= nr; r++) {
        const int j = j0 + r;
        if (j < 0 || j > G.Ny) continue;
        const real* ge = F.eta + (size_t)(j + OB_H) * G.pitch2 + a0;
        const real* gu = F.U + (size_t)(j + OB_H) * G.pitch2 + a0;
        const real* gv = F.V + (size_t)(j + OB_H) * G.pitch2 + a0;
        for (int x = tid; x < Wp; x += NT) {
            if (r < nr) sE[(size_t)(r + 1) * Wp + x] = ge[x];
            if (r >= 0 && r < nr) sU[(size_t)r * Wp + x] = gu[x];
            if (r >= 0) sV[(size_t)r * Wp + x] = gv[x];
        }
    }
    real aE[MAXC], aU[MAXC], aV[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; c++) { aE[c] = RL(0.0); aU[c] = RL(0.0); aV[c] = RL(0.0); }
    __syncthreads();
    unsigned ph0 = 0, ph1 = 0;
    const unsigned row_bytes = (unsigned)Wp * sizeof(real);
    for (int m = 0; m < M; m++) {
        // ---- eta half step: eta -= dtau * div(U, V) / Az
#pragma unroll
        for (int c = 0; c < MAXC; c++) {
            const int wi = tid + c * NT;
            if (wi < nwork) {
                const int r = wi / ncol, i = wi - r * ncol - ex, j = j0 + r;
                const int x = xo + i;
                const real Vn = (j + 1 == G.Ny) ? RL(0.0) : sV[(size_t)(r + 1) * Wp + x];
                const real Vs = (j == 0) ? RL(0.0) : sV[(size_t)r * Wp + x];
                const real* u = sU + (size_t)r * Wp + x;
                real* e = sE + (size_t)(r + 1) * Wp + x;
                const real en = e[0] - dtau * ((G.dy * u[1] - G.dy * u[0]) + (G.dxcf[j + 1] * Vn - G.dxcf[j] * Vs)) / G.azcc[j];
                e[0] = en;
                if (periodic) {
                    if (i == G.Nx - 1) e[-G.Nx] = en;   // west halo column i = -1
                    if (i == 0) e[G.Nx] = en;           // east halo column i = Nx
                }
            }
        }
        __syncthreads();
        // publish my LAST row of eta for the block above, barrier, stage the row below my first
        if (nr > 0 && j1 < G.Ny) {
            real* ge = F.eta + (size_t)(j1 - 1 + OB_H) * G.pitch2 + a0;
            const real* se = sE + (size_t)nr * Wp;
            for (int x = tid; x < Wp; x += NT) ge[x] = se[x];
        }
        grid.sync();
        if (nr > 0 && j0 > 0) {
            if (tid == 0) {
                asm volatile("fence.proxy.async;" ::: "memory");
                bulk_row_load(sE, F.eta + (size_t)(j0 - 1 + OB_H) * G.pitch2 + a0, row_bytes, &bar[0]);
            }
            bar_wait(&bar[0], ph0); ph0 ^= 1;
        }
        // ---- velocity half step + averaging
        const real wgt = wts[m];
#pragma unroll
        for (int c = 0; c < MAXC; c++) {
            const int wi = tid + c * NT;
            if (wi < nwork) {
                const int r = wi / ncol, i = wi - r * ncol - ex, j = j0 + r;
                const int x = xo + i;
                const int cg2 = idx2b(G, i, j);
                const real* e = sE + (size_t)(r + 1) * Wp + x;
                const real en = e[0];
                const real detax = (en - e[-1]) / G.dxfc[j];
                const real detay = (j == 0) ? RL(0.0) : (en - e[-Wp]) / G.dy;
                real* u = sU + (size_t)r * Wp + x;
                real* v = sV + (size_t)r * Wp + x;
                const real Un = u[0] + dtau * (-G.g * F.Hfc[cg2] * detax + F.GU[cg2]);
                const real Vn = v[0] + dtau * (-G.g * F.Hcf[cg2] * detay + F.GV[cg2]);
                u[0] = Un; v[0] = Vn;
                aE[c] += wgt * en; aU[c] += wgt * Un; aV[c] += wgt * Vn;
                if (periodic) {
                    if (i == 0) { u[G.Nx] = Un; v[G.Nx] = Vn; }
                    if (i == G.Nx - 1) { u[-G.Nx] = Un; v[-G.Nx] = Vn; }
                }
            }
        }
        __syncthreads();
        if (m + 1 < M) {
            // publish my FIRST row of V for the block below, barrier, stage the row above my last
            if (nr > 0 && j0 > 0) {
                real* gv = F.V + (size_t)(j0 + OB_H) * G.pitch2 + a0;
                for (int x = tid; x < Wp; x += NT) gv[x] = sV[x];
            }
            grid.sync();
            if (nr > 0 && j1 < G.Ny) {
                if (tid == 0) {
                    asm volatile("fence.proxy.async;" ::: "memory");
                    bulk_row_load(sV + (size_t)nr * Wp, F.V + (size_t)(j1 + OB_H) * G.pitch2 + a0, row_bytes, &bar[1]);
                }
                bar_wait(&bar[1], ph1); ph1 ^= 1;
            }
        }
    }
    // write back: eta <- eta_bar (interior columns, as k_baro_finish), the averages, the final transports
#pragma unroll
    for (int c = 0; c < MAXC; c++) {
        const int wi = tid + c * NT;
        if (wi < nwork) {
            const int r = wi / ncol, i = wi - r * ncol - ex, j = j0 + r;
            const int cg2 = idx2b(G, i, j);
            F.etab[cg2] = aE[c]; F.Ub[cg2] = aU[c]; F.Vb[cg2] = aV[c];
            F.U[cg2] = sU[(size_t)r * Wp + xo + i];
            F.V[cg2] = sV[(size_t)r * Wp + xo + i];
            F.eta[cg2] = ((unsigned)i < (unsigned)G.Nx) ? aE[c] : sE[(size_t)(r + 1) * Wp + xo + i];
        }
    }
}

// geometry of the on-chip variant; returns false when the slab does not fit one block per row band
static bool onchip_plan(const Geom& G, int ex, int* R, int* Wp, int* a0, int* nblk, int* maxc, size_t* smem) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    *R = (G.Ny + sms - 1) / sms;
    *nblk = (G.Ny + *R - 1) / *R;
    const int lo = G.X0b - ex - 1;                  // first array column needed (i = -ex - 1)
    *a0 = lo & ~1;                                  // even -> 16-byte aligned bulk copies
    const int hi = G.X0b + G.Nx + ex + 1;           // one past the last (i = Nx + ex)
    *Wp = ((hi - *a0) + 1) & ~1;
    if (*a0 < 0 || *a0 + *Wp > G.pitch2) return false;
    *smem = (size_t)(3 * *R + 2) * *Wp * sizeof(real);
    const int per_thread = ((*R) * (G.Nx + 2 * ex) + 1023) / 1024;
    *maxc = per_thread <= 3 ? 3 : per_thread <= 6 ? 6 : per_thread <= 10 ? 10 : 0;
    return *smem <= 220 * 1024 && *maxc > 0;
}
bool baro_onchip_fits(const Geom& G, int ex) {
#ifdef OB200_F32
    return false;   // the bulk row copies are laid out for 8-byte elements (16-byte granules): Float64 build only
#endif
    int R, Wp, a0, nblk, maxc; size_t smem;
    return onchip_plan(G, ex, &R, &Wp, &a0, &nblk, &maxc, &smem);
}
int launch_baro_onchip(const Geom& G, const Fields& F, real dtau, const real* wts, int M, bool periodic, int ex, cudaStream_t st) {
    int R, Wp, a0, nblk, maxc; size_t smem;
    if (!onchip_plan(G, ex, &R, &Wp, &a0, &nblk, &maxc, &smem)) return -2;
    void* kern = maxc == 3 ? (void*)k_baro_onchip<3> : maxc == 6 ? (void*)k_baro_onchip<6> : (void*)k_baro_onchip<10>;
    static std::vector<std::pair<const void*, int>> done;
    int dev = 0;
    cudaGetDevice(&dev);
    const std::pair<const void*, int> key(kern, dev);
    if (std::find(done.begin(), done.end(), key) == done.end()) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);   // 227 KB minus the static mbarriers
        done.push_back(key);
    }
    Geom g = G; Fields f = F;
    void* args[] = {&g, &f, &dtau, &wts, &M, &periodic, &ex, &R, &Wp, &a0};
    cudaError_t e = cudaLaunchCooperativeKernel(kern, dim3(nblk), dim3(1024), args, smem, st);
    OB_COUNT_LAUNCH(1);
    return e == cudaSuccess ? 0 : -1;
}
