// Halo fills, immersed masking, field transfer, x-slab pack/unpack, diagnostics.
#include "kernels.h"

// x: periodic copy of OB_H columns (single rank).  Threads over (h, row j, plane k).
__global__ void k_fill_x(Geom G, Fields F) {
    const int h = threadIdx.x;                        // 0..OB_H-1
    const int j = blockIdx.x * blockDim.y + threadIdx.y;   // 0..Ny (interior rows, plus the v wall row)
    const int k = blockIdx.y;
    if (h >= OB_H || j > G.Ny) return;
    real* fs[4] = {F.u, F.v, F.T, F.S};
    const long long c = idx3(G, 0, j, k);
    if (k < G.Nz) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            real* f = fs[q];
            f[c - 1 - h] = f[c + G.Nx - 1 - h];
            f[c + G.Nx + h] = f[c + h];
        }
    } else {  // 2-D eta rides along as "plane Nz"
        const int c2 = idx2b(G, 0, j);
        F.eta[c2 - 1 - h] = F.eta[c2 + G.Nx - 1 - h];
        F.eta[c2 + G.Nx + h] = F.eta[c2 + h];
    }
}

// y: centre-located fields get zero-gradient copies of the boundary row in every halo row; v is zero on and
// beyond the wall faces.  Threads over all columns incl. x-halos.
__global__ void k_fill_y(Geom G, Fields F) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x - OB_H;
    const int k = blockIdx.y;
    if (i >= G.Nx + OB_H) return;
    if (k < G.Nz) {
        real* fs[3] = {F.u, F.T, F.S};
        const long long s0 = idx3(G, i, 0, k), s1 = idx3(G, i, G.Ny - 1, k);
#pragma unroll
        for (int q = 0; q < 3; q++) {
            real* f = fs[q];
            const real a = f[s0], b = f[s1];
            for (int h = 1; h <= OB_H; h++) { f[s0 - (long long)h * G.pitch] = a; f[s1 + (long long)h * G.pitch] = b; }
        }
        const long long v0 = idx3(G, i, 0, k), v1 = idx3(G, i, G.Ny, k);
        for (int h = 0; h <= OB_H; h++) { F.v[v0 - (long long)h * G.pitch] = RL(0.0); F.v[v1 + (long long)h * G.pitch] = RL(0.0); }
    } else {
        const int s0 = idx2b(G, i, 0), s1 = idx2b(G, i, G.Ny - 1);
        const real a = F.eta[s0], b = F.eta[s1];
        for (int h = 1; h <= OB_H; h++) { F.eta[s0 - h * G.pitch2] = a; F.eta[s1 + h * G.pitch2] = b; }
    }
}

void launch_fill_halos(const Geom& G, const Fields& F, bool periodic_x, cudaStream_t st) {
    if (periodic_x) {
        dim3 b(8, 16), g((G.Ny + 1 + 15) / 16, G.Nz + 1);
        k_fill_x<<<g, b, 0, st>>>(G, F);
        OB_COUNT_LAUNCH(1);
    }
    dim3 b2(128), g2((G.Nx + 2 * OB_H + 127) / 128, G.Nz + 1);
    k_fill_y<<<g2, b2, 0, st>>>(G, F);
    OB_COUNT_LAUNCH(1);
}

// mask_immersed_field! for u, v, T, S and eta; needed after set_field only: the step never
// writes a masked node.
__global__ void k_mask(Geom G, Fields F) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y, k = blockIdx.z;
    if (i >= G.Nx) return;
    const long long c = idx3(G, i, j, k);
    if (j < G.Ny) {
        if (peripheral_fcc(G, i, j, k)) F.u[c] = RL(0.0);
        if (!active(G, i, j, k)) { F.T[c] = RL(0.0); F.S[c] = RL(0.0); }
        if (k == G.Nz - 1 && !active(G, i, j, k)) F.eta[idx2b(G, i, j)] = RL(0.0);
    }
    if (peripheral_cfc(G, i, j, k)) F.v[c] = RL(0.0);
}
void launch_mask_fields(const Geom& G, const Fields& F, cudaStream_t st) {
    dim3 b(128), g((G.Nx + 127) / 128, G.Ny + 1, G.Nz);
    k_mask<<<g, b, 0, st>>>(G, F);
    OB_COUNT_LAUNCH(1);
}

// interior <-> user buffer laid out like an Oceananigans parent array with halos (hx, hy, hz)
__global__ void k_copy_field(Geom G, real* field, real* buf, int ny, int nz, int hx, int hy, int hz, int layout,
                             bool to_field, int ex) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x - ex;   // ex x-halo columns travel too (option transfer_x_halo)
    const int j = blockIdx.y, k = blockIdx.z;
    if (i >= G.Nx + ex) return;
    const size_t sx = G.Nx + 2 * hx, sy = ny + 2 * hy;
    const size_t b = ((size_t)(k + (layout == 0 ? hz : 0)) * sy + (j + hy)) * sx + (i + hx);
    const long long c = layout == 0 ? idx3(G, i, j, k) : layout == 1 ? (long long)idx2(G, i, j) : (long long)idx2b(G, i, j);
    if (to_field) field[c] = buf[b]; else buf[b] = field[c];
}
void launch_copy_field(const Geom& G, real* field, real* buf, int ny, int nz, int hx, int hy, int hz, int layout,
                       bool to_field, int ex, cudaStream_t st) {
    dim3 b(128), g((G.Nx + 2 * ex + 127) / 128, ny, nz);
    k_copy_field<<<g, b, 0, st>>>(G, field, buf, ny, nz, hx, hy, hz, layout, to_field, ex);
    OB_COUNT_LAUNCH(1);
}

// ---- x-slab exchange buffers: per field Nz planes x (Ny+1) rows x OB_H cols; eta ((Ny+1) x OB_H) rides with u, v.
// Two groups with their own buffers so that they can travel at different times: T, S are final after their implicit solve
// (their exchange hides behind the u, v solve and the barotropic loop), u, v, eta only after the corrector.
size_t pack_x_count(const Geom& G, int group) {
    return (size_t)OB_H * (G.Ny + 1) * (2 * (size_t)G.Nz + (group == OB_XG_UVETA ? 1 : 0));
}

__global__ void k_pack_x(Geom G, Fields F, real* sendW, real* sendE, bool unpack, int group) {
    const int h = threadIdx.x;
    const int j = blockIdx.x * blockDim.y + threadIdx.y;
    const int k = blockIdx.y;   // 0..Nz (Nz = eta, group u, v, eta only)
    if (h >= OB_H || j > G.Ny) return;
    const size_t per = (size_t)OB_H * (G.Ny + 1);
    const size_t o = (size_t)j * OB_H + h;
    if (k < G.Nz) {
        real* fs[2] = {group == OB_XG_UVETA ? F.u : F.T, group == OB_XG_UVETA ? F.v : F.S};
        const long long c = idx3(G, 0, j, k);
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const size_t b = ((size_t)q * G.Nz + k) * per + o;
            if (!unpack) { sendW[b] = fs[q][c + h]; sendE[b] = fs[q][c + G.Nx - OB_H + h]; }
            else { fs[q][c - OB_H + h] = sendW[b]; fs[q][c + G.Nx + h] = sendE[b]; }
        }
    } else {
        const size_t b = (size_t)2 * G.Nz * per + o;
        const int c = idx2b(G, 0, j);
        if (!unpack) { sendW[b] = F.eta[c + h]; sendE[b] = F.eta[c + G.Nx - OB_H + h]; }
        else { F.eta[c - OB_H + h] = sendW[b]; F.eta[c + G.Nx + h] = sendE[b]; }
    }
}
void launch_pack_x(const Geom& G, const Fields& F, real* sendW, real* sendE, int group, cudaStream_t st) {
    dim3 b(8, 16), g((G.Ny + 1 + 15) / 16, G.Nz + (group == OB_XG_UVETA ? 1 : 0));
    k_pack_x<<<g, b, 0, st>>>(G, F, sendW, sendE, false, group);
    OB_COUNT_LAUNCH(1);
}
// recvW = data for the west halo (from the west neighbour's east edge), recvE likewise
void launch_unpack_x(const Geom& G, const Fields& F, const real* recvW, const real* recvE, int group, cudaStream_t st) {
    dim3 b(8, 16), g((G.Ny + 1 + 15) / 16, G.Nz + (group == OB_XG_UVETA ? 1 : 0));
    k_pack_x<<<g, b, 0, st>>>(G, F, const_cast<real*>(recvW), const_cast<real*>(recvE), true, group);
    OB_COUNT_LAUNCH(1);
}

// ---- diagnostics: max |u|,|v|,|w|,|eta|,|T|,|S| and the advective CFL time scale (TimeStepWizard) ----------
__device__ __forceinline__ void atomic_max_pos(double* addr, double v) {   // v >= 0 or the canonical (positive) NaN, /*f64*/
    atomicMax((unsigned long long*)addr, (unsigned long long)__double_as_longlong(v));   // whose bit pattern tops every finite one /*f64*/
}
// max that PROPAGATES NaN like Julia's `maximum(abs, u)` (fmax drops it): after a blow-up the progress line and the wizard
// must see NaN, not the maximum over the cells that are still finite
__device__ __forceinline__ real max_nan(real a, real b) { return (a != a || b != b) ? nan("") : fmax(a, b); }
__global__ void __launch_bounds__(256) k_diag(Geom G, Fields F, double* out) { /*f64*/
    // one block per (row j, level-chunk): threads stride over i and 8 levels, then one atomic per block and quantity
    __shared__ real red[7][8];
    const int j = blockIdx.y, k0 = blockIdx.z * 8;
    real m[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int k = k0; k < min(k0 + 8, G.Nz); k++)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < G.Nx; i += gridDim.x * blockDim.x) {
            const long long c = idx3(G, i, j, k);
            const real au = fabs(F.u[c]), av = fabs(F.v[c]), aw = fabs(F.w[c]);
            m[0] = max_nan(m[0], au); m[1] = max_nan(m[1], av);
            m[2] = max_nan(m[2], max_nan(aw, k == G.Nz - 1 ? fabs(F.w[c + G.sz]) : RL(0.0)));
            if (k == 0) m[3] = max_nan(m[3], fabs(F.eta[idx2b(G, i, j)]));
            m[4] = max_nan(m[4], fabs(F.T[c])); m[5] = max_nan(m[5], fabs(F.S[c]));
            // 1 / cell advection time scale: |u|/dx + |v|/dy + |w|/dz
            m[6] = max_nan(m[6], au / G.dxfc[j] + av / G.dy + aw / G.dzc[k]);
        }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int q = 0; q < 7; q++) {
        real v = m[q];
        for (int o = 16; o > 0; o >>= 1) v = max_nan(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (lane == 0) red[q][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < 7) {
        real v = RL(0.0);
        for (int w = 0; w < 8; w++) v = max_nan(v, red[threadIdx.x][w]);
        if (v > RL(0.0) || v != v) atomic_max_pos(out + threadIdx.x, v);
    }
}
void launch_diagnostics(const Geom& G, const Fields& F, double* out8, cudaStream_t st) { /*f64*/
    cudaMemsetAsync(out8, 0, 8 * sizeof(double), st); /*f64*/
    dim3 b(256), g(2, G.Ny, (G.Nz + 7) / 8);
    k_diag<<<g, b, 0, st>>>(G, F, out8);
    OB_COUNT_LAUNCH(1);
}
