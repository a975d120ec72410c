// Split-explicit free surface: barotropic substeps (forward-backward scheme).
//   eta -= dtau * div(U, V) / Az ;  U, V += dtau * (-g H d(eta) + G^{U,V}) ; averages += w_m * (eta, U, V)
// One GPU: the producing kernel also maintains the periodic x-halo column its consumer needs (eta at i = -1,
// U at i = Nx), so there are no halo launches inside the loop.  Several GPUs: the barotropic arrays carry a
// halo of Wb = M + 1 columns that is exchanged ONCE before the loop (as Oceananigans' distributed solver does,
// SURVEY F9); every substep is computed over [-ex, Nx + ex) and the invalid fringe, which grows by one column
// per substep, never reaches the interior.
// launch_baro_persistent: one cooperative launch for all M substeps (grid-wide barrier between half steps).
#include <cooperative_groups.h>

#include "kernels.h"
namespace cg = cooperative_groups;

__global__ void k_baro_init(Geom G, Fields F, bool periodic, int ex) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x - ex;
    const int j = blockIdx.y;  // 0..Ny
    if (i >= G.Nx + ex) return;
    const int c = idx2b(G, i, j);
    const double a = F.Ub[c], b = F.Vb[c];
    F.U[c] = a;
    F.V[c] = b;
    if (periodic) {   // the first eta substep reads U at the east halo column
        if (i < OB_H) { F.U[c + G.Nx] = a; F.V[c + G.Nx] = b; }
        if (i >= G.Nx - OB_H) { F.U[c - G.Nx] = a; F.V[c - G.Nx] = b; }
    }
    F.etab[c] = 0.0; F.Ub[c] = 0.0; F.Vb[c] = 0.0;
}

// Per-substep slab exchange (x-slabs narrower than the wide halo, and the microbenchmark of BASELINE configs[4]): the one
// column a neighbour needs travels through contiguous edge buffers -- the producer writes its edge value there as well,
// the consumer reads the neighbour's from there -- so there are no pack / unpack launches between the NCCL calls.
struct BaroEdge {
    double* send_eta;         // my eta(Nx - 1, j)      -> east neighbour's eta(-1, j)
    const double* recv_eta;   // west neighbour's eta(Nx - 1, j)
    double* send_U;           // my U(0, j)             -> west neighbour's U(Nx, j)
    const double* recv_U;     // east neighbour's U(0, j)
};

__device__ __forceinline__ void eta_update(const Geom& G, const Fields& F, int i, int j, double dtau, bool periodic,
                                           const BaroEdge* E = nullptr) {
    const int c = idx2b(G, i, j);
    const double Vn = (j + 1 == G.Ny) ? 0.0 : F.V[c + G.pitch2];
    const double Vs = (j == 0) ? 0.0 : F.V[c];
    const double Ue = (E && i == G.Nx - 1) ? E->recv_U[j] : F.U[c + 1];
    const double e = F.eta[c] - dtau * ((G.dy * Ue - G.dy * F.U[c]) + (G.dxcf[j + 1] * Vn - G.dxcf[j] * Vs)) / G.azcc[j];
    F.eta[c] = e;
    if (periodic) {
        if (i == G.Nx - 1) F.eta[c - G.Nx] = e;   // west halo column i = -1
        if (i == 0) F.eta[c + G.Nx] = e;          // east halo column i = Nx
    }
    if (E && i == G.Nx - 1) E->send_eta[j] = e;
}

__device__ __forceinline__ void vel_update(const Geom& G, const Fields& F, int i, int j, double dtau, double wgt,
                                           bool periodic, const BaroEdge* E = nullptr) {
    const int c = idx2b(G, i, j);
    const double e = F.eta[c];
    const double ew = (E && i == 0) ? E->recv_eta[j] : F.eta[c - 1];
    const double detax = (e - ew) / G.dxfc[j];
    const double detay = (j == 0) ? 0.0 : (e - F.eta[c - G.pitch2]) / G.dy;
    const double Un = F.U[c] + dtau * (-G.g * F.Hfc[c] * detax + F.GU[c]);
    const double Vn = F.V[c] + dtau * (-G.g * F.Hcf[c] * detay + F.GV[c]);
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

__global__ void __launch_bounds__(128) k_baro_eta(Geom G, Fields F, double dtau, bool periodic, int ex) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x - ex;
    if (i >= G.Nx + ex) return;
    eta_update(G, F, i, blockIdx.y, dtau, periodic);
}
__global__ void __launch_bounds__(128) k_baro_vel(Geom G, Fields F, double dtau, double wgt, bool periodic, int ex) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x - ex;
    if (i >= G.Nx + ex) return;
    vel_update(G, F, i, blockIdx.y, dtau, wgt, periodic);
}
__global__ void __launch_bounds__(128) k_baro_eta_edge(Geom G, Fields F, double dtau, BaroEdge E) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G.Nx) return;
    eta_update(G, F, i, blockIdx.y, dtau, false, &E);
}
__global__ void __launch_bounds__(128) k_baro_vel_edge(Geom G, Fields F, double dtau, double wgt, BaroEdge E) {
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
void launch_baro_eta(const Geom& G, const Fields& F, double dtau, bool periodic, int ex, cudaStream_t st) {
    dim3 b(128), g((G.Nx + 2 * ex + 127) / 128, G.Ny);
    k_baro_eta<<<g, b, 0, st>>>(G, F, dtau, periodic, ex);
    OB_COUNT_LAUNCH(1);
}
void launch_baro_vel(const Geom& G, const Fields& F, double dtau, double wgt, bool periodic, int ex, cudaStream_t st) {
    dim3 b(128), g((G.Nx + 2 * ex + 127) / 128, G.Ny);
    k_baro_vel<<<g, b, 0, st>>>(G, F, dtau, wgt, periodic, ex);
    OB_COUNT_LAUNCH(1);
}
void launch_baro_eta_edge(const Geom& G, const Fields& F, double dtau, double* send_eta, const double* recv_eta, double* send_U,
                          const double* recv_U, cudaStream_t st) {
    dim3 b(128), g((G.Nx + 127) / 128, G.Ny);
    k_baro_eta_edge<<<g, b, 0, st>>>(G, F, dtau, BaroEdge{send_eta, recv_eta, send_U, recv_U});
    OB_COUNT_LAUNCH(1);
}
void launch_baro_vel_edge(const Geom& G, const Fields& F, double dtau, double wgt, double* send_eta, const double* recv_eta,
                          double* send_U, const double* recv_U, cudaStream_t st) {
    dim3 b(128), g((G.Nx + 127) / 128, G.Ny);
    k_baro_vel_edge<<<g, b, 0, st>>>(G, F, dtau, wgt, BaroEdge{send_eta, recv_eta, send_U, recv_U});
    OB_COUNT_LAUNCH(1);
}
void launch_baro_edge_first(const Geom& G, const Fields& F, double* send_U, cudaStream_t st) {
    k_baro_edge_first<<<(G.Ny + 127) / 128, 128, 0, st>>>(G, F, BaroEdge{nullptr, nullptr, send_U, nullptr});
    OB_COUNT_LAUNCH(1);
}
void launch_baro_finish(const Geom& G, const Fields& F, cudaStream_t st) {
    dim3 b(128), g((G.Nx + 127) / 128, G.Ny);
    k_baro_finish<<<g, b, 0, st>>>(G, F);
    OB_COUNT_LAUNCH(1);
}

// ---- persistent variant: all substeps in one cooperative launch --------------------------------------
__global__ void __launch_bounds__(256) k_baro_persistent(Geom G, Fields F, double dtau, const double* __restrict__ wts,
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
        const double wgt = wts[m];
        for (int it = warp; it < nitems; it += nwarps) {
            const int j = it / xchunks, i0 = (it - j * xchunks) * 256 - ex;
            for (int i = i0 + lane; i < min(i0 + 256, G.Nx + ex); i += 32) vel_update(G, F, i, j, dtau, wgt, periodic);
        }
        grid.sync();
    }
}

int launch_baro_persistent(const Geom& G, const Fields& F, double dtau, const double* wts, int M, bool periodic, int ex,
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
__global__ void k_pack_baro(Geom G, Fields F, double* bufW, double* bufE, int W, bool unpack) {
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y, q = blockIdx.z;
    if (h >= W) return;
    double* fs[5] = {F.Ub, F.Vb, F.eta, F.GU, F.GV};
    double* f = fs[q];
    const size_t b = ((size_t)q * (G.Ny + 1) + j) * W + h;
    const int c = idx2b(G, 0, j);
    if (!unpack) { bufW[b] = f[c + h]; bufE[b] = f[c + G.Nx - W + h]; }
    else { f[c - W + h] = bufW[b]; f[c + G.Nx + h] = bufE[b]; }
}
size_t pack_baro_count(const Geom& G, int W) { return (size_t)5 * (G.Ny + 1) * W; }
void launch_pack_baro(const Geom& G, const Fields& F, double* bufW, double* bufE, int W, bool unpack, cudaStream_t st) {
    dim3 b(64), g((W + 63) / 64, G.Ny + 1, 5);
    k_pack_baro<<<g, b, 0, st>>>(G, F, bufW, bufE, W, unpack);
    OB_COUNT_LAUNCH(1);
}
