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
#ifndef TEND_LEITH
#define TEND_LEITH 1   // 0: compile the QG-Leith stress out (instruction-count studies only)
#endif

// ---- TMA (cp.async.bulk.tensor) + mbarrier helpers: one elected thread fetches a whole (box_x, box_y, 1) tile of a
// 3-D field into shared memory; completion is signalled on an mbarrier by transaction bytes.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "OB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra OB_DONE;\n"
        "bra OB_WAIT;\n"
        "OB_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_tile(void* dst, const CUtensorMap* tm, unsigned long long* bar, int x, int y, int z) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}

// G^n of field q (0 u, 1 v, 2 T, 3 S) and, when allocated, the AB2 combination the next regular step's implicit solve reads
// instead of G^n and G^- (same expression, same operand order as k_ab2_implicit: bit-identical)
__device__ __forceinline__ void store_tendency(const Geom& G, const Fields& F, int q, long long c, real g) {
    real* Gn = q == 0 ? F.Gu : q == 1 ? F.Gv : q == 2 ? F.GT : F.GS;
    Gn[c] = g;
    if (F.Gab[q]) {
        const real* G0 = q == 0 ? F.Gu0 : q == 1 ? F.Gv0 : q == 2 ? F.GT0 : F.GS0;
        F.Gab[q][c] = G.ab2_ca * g - G.ab2_cb * G0[c];
    }
}

// ---- accessors: the same operators read either straight from HBM/L1 (MomG) or from a shared-memory tile
// in which zeta, the tangential-velocity stencils, the divergence pieces and K were computed once (MomS).
struct MomG {
    const Geom& G;
    const real* __restrict__ u;
    const real* __restrict__ v;
    __device__ __forceinline__ real U(int i, int j, int k) const { return u[idx3(G, i, j, k)]; }
    __device__ __forceinline__ real V(int i, int j, int k) const { return v[idx3(G, i, j, k)]; }
    __device__ __forceinline__ real usm(int i, int j, int k) const { return RL(0.5) * (U(i, j - 1, k) + U(i, j, k)); }
    __device__ __forceinline__ real vsm(int i, int j, int k) const { return RL(0.5) * (V(i - 1, j, k) + V(i, j, k)); }
    // vertical vorticity at (f,f,c); GEN applies the immersed conditional differences
    template <bool GEN>
    __device__ __forceinline__ real zeta(int i, int j, int k) const {
        const long long c = idx3(G, i, j, k);
        real dxv = G.dy * v[c] - G.dy * v[c - 1];
        real dyu = G.dxfc[j] * u[c] - G.dxfc[j - 1] * u[c - G.pitch];
        if (GEN) {
            if (inactive_cfc(G, i - 1, j, k) || inactive_cfc(G, i, j, k)) dxv = RL(0.0);
            if (inactive_fcc(G, i, j - 1, k) || inactive_fcc(G, i, j, k)) dyu = RL(0.0);
        }
        return (dxv - dyu) * G.razcf[j];
    }
    __device__ __forceinline__ real dxU(int i, int j, int k) const {
        const long long c = idx3(G, i, j, k);
        const real ax = G.dy * G.dzc[k];
        return ax * u[c + 1] - ax * u[c];
    }
    __device__ __forceinline__ real dyV(int i, int j, int k) const {
        const long long c = idx3(G, i, j, k);
        return G.dxcf[j + 1] * G.dzc[k] * v[c + G.pitch] - G.dxcf[j] * G.dzc[k] * v[c];
    }
    __device__ __forceinline__ real Kh(int i, int j, int k) const {
        const long long c = idx3(G, i, j, k);
        const real a = u[c], b = u[c + 1], cc = v[c], d = v[c + G.pitch];
        return (RL(0.5) * (a * a + b * b) + RL(0.5) * (cc * cc + d * d)) / RL(2.0);
    }
    __device__ __forceinline__ real divxy(int i, int j, int k) const {   // horizontal divergence at ccc
        const long long c = idx3(G, i, j, k);
        return (G.dy * u[c + 1] - G.dy * u[c] + G.dxcf[j + 1] * v[c + G.pitch] - G.dxcf[j] * v[c]) * G.razcc[j];
    }
};

// per-latitude metrics of a tile's rows in shared memory: [metric][row], row r <-> j = j0 - MT_J0 + r.  One LDS with an
// immediate offset replaces the constant-bank pointer load + address arithmetic + global load of `G.metric[j]`.
enum { MT_DXFC = 0, MT_DXCF, MT_RAZCF, MT_RDXFC, MT_AZCC, MT_RAZCC, MT_FFF, MT_TAUX, MT_N };
#define MT_ROWS 28
#define MT_J0 6
__device__ __forceinline__ void load_metric_table(const Geom& G, real* smt, int j0, int tid, int nt) {
    const real* src[MT_N] = {G.dxfc, G.dxcf, G.razcf, G.rdxfc, G.azcc, G.razcc, G.fff, G.taux};
    for (int e = tid; e < MT_N * MT_ROWS; e += nt) {
        const int m = e / MT_ROWS, r = e % MT_ROWS;
        const int j = min(j0 - MT_J0 + r, G.Ny + OB_H + 1);   // the arrays are indexable for j = -8 .. Ny + 8
        real v = RL(0.0);
#pragma unroll
        for (int q = 0; q < MT_N; q++) if (q == m) v = src[q][j];
        smt[e] = v;
    }
}
#define MET(mrow, name, dj) ((mrow)[MT_##name * MT_ROWS + (dj)])

// MOM_DB = 1: the derived fields of consecutive levels alternate between two shared-memory buffers, so the barrier that
// used to end a level (protecting them from the next level's staging) goes away: warps that finish a level early stage the
// next one while the slow ones still compute.  77 KB per block -> 3 resident blocks (was 4 with 2 barriers per level).
// Measured at C3 on one B200: 19.01 ms against 18.41 ms for the single-buffered kernel with 4 resident blocks -- the
// occupancy is worth more than the barrier -- so it is OFF by default.
#ifndef MOM_DB
#define MOM_DB 0
#endif
// MOM_WSM = 1: the vertical velocity of the tile (levels k and k + 1) is staged in a two-slot shared-memory ring -- one global
// load per thread and level (issued at the top of the derived-field phase, stored at its end) instead of eight with 64-bit
// address arithmetic in the middle of the tendency phase.  Needs the end-of-level barrier (not combined with MOM_DB).
#ifndef MOM_WSM
#define MOM_WSM (MOM_DB ? 0 : 1)
#endif
// shared-memory tile geometry for the fast path (TX x TY cells per block)
template <int TX, int TY>
struct TileDims {
    static constexpr int HU = 5;                         // halo of the u, v tiles
    // i in [i0-6, i0+TX+5], j in [j0-5, j0+TY+5]: one extra column on the left because a TMA box must start on a
    // 16-byte boundary in global memory (an odd start column of doubles raises "illegal instruction") and its rows
    // must be a multiple of 16 bytes
    static constexpr int HUX = HU + 1;
    static constexpr int PU = TX + HUX + HU + 1, RU = TY + 2 * HU + 1;
    static constexpr int HZ = 4;                         // zeta / us / vs: i in [i0-4, i0+TX+4], j likewise
    static constexpr int PZ = TX + 2 * HZ + 1, RZ = TY + 2 * HZ + 1;
    static constexpr int HD = 3;                         // dxU, dyV: [i0-3, i0+TX+2]
    static constexpr int PD = TX + 2 * HD, RD = TY + 2 * HD;
    static constexpr int PK = TX + 1, RK = TY + 1;       // K: [i0-1, i0+TX-1]
    static constexpr int NU = PU * RU, NZ = PZ * RZ, ND = PD * RD, NK = PK * RK;
    static constexpr int NUP = (NU + 15) / 16 * 16;      // tile slot padded to 128 bytes (TMA destination alignment)
    static constexpr int NDER = 3 * NZ + 2 * ND + NK;     // derived fields of one level: zeta, I_y u, I_x v, dxU, dyV, K
    // doubles: u, v double-buffered (TMA) + derived fields (double-buffered when MOM_DB: ONE block barrier per level) + metrics
    // w at the cells [i0 - 1, i0 + TX - 1] x [j0 - 1, j0 + TY - 1], levels k and k + 1 (ring of two slots)
    static constexpr int PW = TX + 1, NW = PW * (TY + 1);
    static constexpr int TOTAL = 4 * NUP + (MOM_DB ? 2 : 1) * NDER + MT_N * MT_ROWS + (MOM_WSM ? 2 * NW : 0);
};

template <int TX, int TY>
struct MomS {
    using D = TileDims<TX, TY>;
    const real *su, *sv, *sz, *sus, *svs, *sdx, *sdy, *sK;
    int i0, j0;
    __device__ __forceinline__ real U(int i, int j, int) const { return su[(j - j0 + D::HU) * D::PU + (i - i0 + D::HUX)]; }
    __device__ __forceinline__ real V(int i, int j, int) const { return sv[(j - j0 + D::HU) * D::PU + (i - i0 + D::HUX)]; }
    __device__ __forceinline__ real usm(int i, int j, int) const { return sus[(j - j0 + D::HZ) * D::PZ + (i - i0 + D::HZ)]; }
    __device__ __forceinline__ real vsm(int i, int j, int) const { return svs[(j - j0 + D::HZ) * D::PZ + (i - i0 + D::HZ)]; }
    template <bool GEN>
    __device__ __forceinline__ real zeta(int i, int j, int) const { return sz[(j - j0 + D::HZ) * D::PZ + (i - i0 + D::HZ)]; }
    __device__ __forceinline__ real dxU(int i, int j, int) const { return sdx[(j - j0 + D::HD) * D::PD + (i - i0 + D::HD)]; }
    __device__ __forceinline__ real dyV(int i, int j, int) const { return sdy[(j - j0 + D::HD) * D::PD + (i - i0 + D::HD)]; }
    __device__ __forceinline__ real Kh(int i, int j, int) const { return sK[(j - j0 + 1) * D::PK + (i - i0 + 1)]; }
};

// ---- biased reconstructions along a line; `base + sgn*m` walks from the most upwind-ward point -------------
template <int N, bool GEN, class A>
__device__ __forceinline__ real recon_vort_y(const A& M, int i, int j, int k, bool left) {
    const int base = left ? j - N + 1 : j + N, sgn = left ? 1 : -1;
    real zq[2 * N - 1], us[2 * N - 1], vs[2 * N - 1];
#pragma unroll
    for (int m = 0; m < 2 * N - 1; m++) {
        const int jj = base + sgn * m;
        zq[m] = M.template zeta<GEN>(i, jj, k);
        us[m] = M.usm(i, jj, k);
        vs[m] = M.vsm(i, jj, k);
    }
    if (N == 1) return zq[0];
    real bu[N], bv[N];
    Weno<N>::beta2(us, vs, bu, bv);
#pragma unroll
    for (int s = 0; s < N; s++) bu[s] = RL(0.5) * (bu[s] + bv[s]);
    return Weno<N>::combine(zq, bu);
}
template <int N, bool GEN, class A>
__device__ __forceinline__ real recon_vort_x(const A& M, int i, int j, int k, bool left) {
    const int base = left ? i - N + 1 : i + N, sgn = left ? 1 : -1;
    real zq[2 * N - 1], us[2 * N - 1], vs[2 * N - 1];
#pragma unroll
    for (int m = 0; m < 2 * N - 1; m++) {
        const int ii = base + sgn * m;
        zq[m] = M.template zeta<GEN>(ii, j, k);
        us[m] = M.usm(ii, j, k);
        vs[m] = M.vsm(ii, j, k);
    }
    if (N == 1) return zq[0];
    real bu[N], bv[N];
    Weno<N>::beta2(us, vs, bu, bv);
#pragma unroll
    for (int s = 0; s < N; s++) bu[s] = RL(0.5) * (bu[s] + bv[s]);
    return Weno<N>::combine(zq, bu);
}
template <int N, class A>
__device__ __forceinline__ real recon_div_x(const A& M, int i, int j, int k, bool left) {
    const int base = left ? i - N : i + N - 1, sgn = left ? 1 : -1;
    real q[2 * N - 1], s[2 * N - 1];
#pragma unroll
    for (int m = 0; m < 2 * N - 1; m++) {
        const int ii = base + sgn * m;
        q[m] = M.dxU(ii, j, k);
        s[m] = q[m] + M.dyV(ii, j, k);
    }
    if (N == 1) return q[0];
    real b[N];
    Weno<N>::beta(s, b);
    return Weno<N>::combine(q, b);
}
template <int N, class A>
__device__ __forceinline__ real recon_div_y(const A& M, int i, int j, int k, bool left) {
    const int base = left ? j - N : j + N - 1, sgn = left ? 1 : -1;
    real q[2 * N - 1], s[2 * N - 1];
#pragma unroll
    for (int m = 0; m < 2 * N - 1; m++) {
        const int jj = base + sgn * m;
        q[m] = M.dyV(i, jj, k);
        s[m] = M.dxU(i, jj, k) + q[m];
    }
    if (N == 1) return q[0];
    real b[N];
    Weno<N>::beta(s, b);
    return Weno<N>::combine(q, b);
}
template <int N>
__device__ __forceinline__ real recon_cell(const real* __restrict__ c, long long base, long long stride) {
    real q[2 * N - 1];
#pragma unroll
    for (int m = 0; m < 2 * N - 1; m++) q[m] = c[base + stride * m];
    if (N == 1) return q[0];
    real b[N];
    Weno<N>::beta(q, b);
    return Weno<N>::combine(q, b);
}

// WENO-7 upwind-side reconstruction at the low face of tile cell q along a compile-time stride S: one 8-value
// window (immediate LDS offsets) serves both upwind directions; the right-biased stencil is the mirror image.
template <int S>
__device__ __forceinline__ real recon7_window(const real* __restrict__ c, int q, bool left) {
    real w[8];
#pragma unroll
    for (int m = 0; m < 8; m++) w[m] = c[q + (m - 4) * S];
    real qq[7];
#pragma unroll
    for (int m = 0; m < 7; m++) qq[m] = left ? w[m] : w[7 - m];
    real b[4];
    Weno<4>::beta(qq, b);
    return Weno<4>::combine(qq, b);
}

// runtime-order dispatch used by the general path
#define OB_DISPATCH(N, CALL1, CALL2, CALL3, CALL4, CALL5) \
    ((N) >= 5 ? (CALL5) : (N) == 4 ? (CALL4) : (N) == 3 ? (CALL3) : (N) == 2 ? (CALL2) : (CALL1))

// ---- masked tiles: every node predicate is a threshold in k (active(i,j,k) <=> k >= kb(i,j) for a grid-fitted
// bottom), so the order / peripheral / Coriolis tests of a column are folded ONCE per block into level thresholds
// instead of ~500 kb look-ups per thread per level.
#define OB_NEVER 0x7fffffff
__device__ __forceinline__ int kb_at(const Geom& G, int i, int j) { return __ldg(&G.kb[idx2(G, i, j)]); }
struct GenThr {
    int per_u, per_v;      // (f,c,c) / (c,f,c) node peripheral below this level
    int kfast;             // full-order, mask-free stencils from this level
    int vy[6], vx[6];      // vorticity reconstruction of half-order N valid from vy[N] (G^u, along y) / vx[N] (G^v, along x)
    int dx[4], dy[4];      // divergence reconstruction of half-order N valid from dx[N] / dy[N]
    int cu[4], cv[4];      // Coriolis: the four tangential-velocity nodes are non-peripheral from these levels
};
__device__ __forceinline__ void gen_thresholds(const Geom& G, int i, int j, int NV, int ND, GenThr& T) {
    T.per_u = max(kb_at(G, i - 1, j), kb_at(G, i, j));
    T.per_v = max(kb_at(G, i, j - 1), kb_at(G, i, j));
    T.kfast = __ldg(&G.kfast[idx2(G, i, j)]);
    int ty = 0, tx = 0;
    T.vy[0] = T.vx[0] = 0;
#pragma unroll
    for (int N = 1; N <= 5; N++) {   // order N adds the face nodes j-N+1 and j+N (resp. i-N+1, i+N)
        ty = max(ty, max(min(kb_at(G, i, j - N), kb_at(G, i, j - N + 1)), min(kb_at(G, i, j + N - 1), kb_at(G, i, j + N))));
        tx = max(tx, max(min(kb_at(G, i - N, j), kb_at(G, i - N + 1, j)), min(kb_at(G, i + N - 1, j), kb_at(G, i + N, j))));
        T.vy[N] = N <= NV ? ty : OB_NEVER;
        T.vx[N] = N <= NV ? tx : OB_NEVER;
    }
    ty = tx = 0;
    T.dx[0] = T.dy[0] = 0;
#pragma unroll
    for (int N = 1; N <= 3; N++) {   // order N adds the cells i-N and i+N-1 (resp. j-N, j+N-1)
        tx = max(tx, max(kb_at(G, i - N, j), kb_at(G, i + N - 1, j)));
        ty = max(ty, max(kb_at(G, i, j - N), kb_at(G, i, j + N - 1)));
        T.dx[N] = N <= ND ? tx : OB_NEVER;
        T.dy[N] = N <= ND ? ty : OB_NEVER;
    }
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 2; b++) {
            T.cu[2 * a + b] = max(kb_at(G, i - 1 + a, j + b - 1), kb_at(G, i - 1 + a, j + b));      // (c,f,c) node (i-1+a, j+b)
            T.cv[2 * a + b] = max(kb_at(G, i + a - 1, j - 1 + b), kb_at(G, i + a, j - 1 + b));      // (f,c,c) node (i+a, j-1+b)
        }
}
__device__ __forceinline__ int order_from(const int* thr, int Nmax, int k) {
    int N = 0;
#pragma unroll
    for (int n = 1; n <= 5; n++)
        if (n <= Nmax && k >= thr[n]) N = n;   // thresholds are non-decreasing in n
    return N;
}

// ================================================================================================
// momentum tendencies
// ================================================================================================
template <bool GEN, class A>
__device__ __forceinline__ real tend_u(const Geom& G, const Fields& F, const A& M, int i, int j, int k, int NV, int ND,
                                         double time, const GenThr& T, const real* __restrict__ mrow,
                                         const real u_below, const real u_above, const real* __restrict__ wk) {
    const long long c = idx3(G, i, j, k);
    // (1) vorticity flux
    const real dxcf0 = MET(mrow, DXCF, 0), dxcf1 = MET(mrow, DXCF, 1), rdxfc = MET(mrow, RDXFC, 0);
    const real q00 = dxcf0 * M.V(i - 1, j, k), q01 = dxcf1 * M.V(i - 1, j + 1, k);
    const real q10 = dxcf0 * M.V(i, j, k), q11 = dxcf1 * M.V(i, j + 1, k);
    const real vavg = RL(0.5) * (RL(0.5) * (q00 + q01) + RL(0.5) * (q10 + q11));
    const real vhat = vavg * rdxfc;
    real Gt = RL(0.0);
    if (vhat != RL(0.0)) {
        const bool left = vhat > RL(0.0);
        real z;
        if (!GEN) z = recon_vort_y<5, false>(M, i, j, k, left);
        else {
            const int N = order_from(T.vy, 5, k);
            z = OB_DISPATCH(N, (recon_vort_y<1, true>(M, i, j, k, left)), (recon_vort_y<2, true>(M, i, j, k, left)),
                            (recon_vort_y<3, true>(M, i, j, k, left)), (recon_vort_y<4, true>(M, i, j, k, left)),
                            (recon_vort_y<5, true>(M, i, j, k, left)));
        }
        Gt = vhat * z;
    }
    // (2) divergence flux + vertical advection
    const real uh = M.U(i, j, k);
    real Phi = RL(0.0);
    if (uh != RL(0.0)) {
        const bool left = uh > RL(0.0);
        real d;
        if (!GEN) d = recon_div_x<3>(M, i, j, k, left);
        else {
            const int N = order_from(T.dx, 3, k);
            d = OB_DISPATCH(N, (recon_div_x<1>(M, i, j, k, left)), (recon_div_x<2>(M, i, j, k, left)),
                            (recon_div_x<3>(M, i, j, k, left)), (recon_div_x<3>(M, i, j, k, left)),
                            (recon_div_x<3>(M, i, j, k, left)));
        }
        Phi = uh * d + uh * RL(0.5) * (M.dyV(i - 1, j, k) + M.dyV(i, j, k));
    }
    const real az = MET(mrow, AZCC, 0);
    real fl_lo, fl_hi;
    {
        // wk: {w(i,j,k), w(i-1,j,k), w(i,j-1,k), w(i,j,k+1), w(i-1,j,k+1), w(i,j-1,k+1)}
        const real Wlo = RL(0.5) * (az * wk[1] + az * wk[0]);
        const real Whi = RL(0.5) * (az * wk[4] + az * wk[3]);
        fl_lo = Wlo * RL(0.5) * (u_below + uh);    // u(k - 1), clamped at the bottom level
        fl_hi = Whi * RL(0.5) * (uh + u_above);    // u(k + 1), clamped at the top level
        if (GEN) {
            // immersed-peripheral (f,c,f) faces carry no flux: the node below is peripheral but inside the domain
            if (k > 0 && k <= T.per_u) fl_lo = RL(0.0);
            // (the node above an active node is active for a grid-fitted bottom; the top face is a domain boundary)
        }
    }
    Gt -= (Phi + fl_hi - fl_lo) * (MET(mrow, RAZCC, 0) * G.rdzc[k]);
    // (3) Bernoulli head  (4) Coriolis  (5) pressure gradient
    Gt -= (M.Kh(i, j, k) - M.Kh(i - 1, j, k)) * rdxfc;
    int nact = 4;
    if (GEN) {
        nact = 0;
#pragma unroll
        for (int n = 0; n < 4; n++) nact += k >= T.cu[n] ? 1 : 0;
    }
    const real fbar = RL(0.5) * (MET(mrow, FFF, 0) + MET(mrow, FFF, 1));
    const real cor = nact == 0 ? RL(0.0) : nact == 4 ? vavg : vavg / (RL(0.25) * nact);
    Gt += fbar * cor * rdxfc;
    Gt -= (F.p[c] - F.p[c - 1]) * rdxfc;
    // (6) QG-Leith stress divergence
    if (TEND_LEITH && G.qgleith) {
        const MomG Mg{G, F.u, F.v};
        const real ax = G.dy * G.dzc[k];
        auto fccc = [&](int a, int b) { return active(G, a, b, k) ? -F.nuL[idx3(G, a, b, k)] * Mg.divxy(a, b, k) : RL(0.0); };
        auto fffc = [&](int a, int b) {
            const bool per = !active(G, a - 1, b - 1, k) || !active(G, a, b - 1, k) || !active(G, a - 1, b, k) || !active(G, a, b, k);
            if (per && in_domain(G, b - 1, k) && in_domain(G, b, k)) return RL(0.0);
            const long long q = idx3(G, a, b, k);
            const real nu = RL(0.25) * (F.nuL[q - 1 - G.pitch] + F.nuL[q - G.pitch] + F.nuL[q - 1] + F.nuL[q]);
            return nu * Mg.zeta<true>(a, b, k);
        };
        const real t = ax * fccc(i, j) - ax * fccc(i - 1, j) +
                         (G.dxcf[j + 1] * G.dzc[k] * fffc(i, j + 1) - G.dxcf[j] * G.dzc[k] * fffc(i, j));
        Gt -= t / (az * G.dzc[k]);
    }
    // (7) flux boundary conditions
    if (k == G.Nz - 1) {
        real J = RL(0.0);
        if (G.bc_kind == OB200_BC_DOUBLEDRAKE) J = MET(mrow, TAUX, 0);
        else if (G.bc_kind == OB200_BC_REALISTIC) J = interp_flux(G, F.forcing[0], i, j, G.Ny, time, false);
        Gt -= J * G.rdzc[k];
    }
    if (k == 0) {
        real J = RL(0.0);
        if (G.bc_kind == OB200_BC_DOUBLEDRAKE) J = -G.mu_drag * uh;
        else if (G.bc_kind == OB200_BC_REALISTIC) {
            const real v0 = M.V(i - 1, j, k), v1 = M.V(i - 1, j + 1, k), v2 = M.V(i, j, k), v3 = M.V(i, j + 1, k);
            J = -G.mu_drag * uh * sqrt(uh * uh + RL(0.25) * (v0 * v0 + v1 * v1 + v2 * v2 + v3 * v3));
        }
        Gt += J * G.rdzc[k];
    }
    // immersed quadratic bottom drag (RealisticOcean): first non-peripheral node above the sea floor
    if (GEN && G.bc_kind == OB200_BC_REALISTIC && k > 0 && k <= T.per_u) {
        const real v0 = M.V(i - 1, j, k), v1 = M.V(i - 1, j + 1, k), v2 = M.V(i, j, k), v3 = M.V(i, j + 1, k);
        Gt += -G.mu_drag * uh * sqrt(uh * uh + RL(0.25) * (v0 * v0 + v1 * v1 + v2 * v2 + v3 * v3)) * G.rdzc[k];
    }
    return Gt;
}

template <bool GEN, class A>
__device__ __forceinline__ real tend_v(const Geom& G, const Fields& F, const A& M, int i, int j, int k, int NV, int ND,
                                         double time, const GenThr& T, const real* __restrict__ mrow,
                                         const real v_below, const real v_above, const real* __restrict__ wk) {
    const long long c = idx3(G, i, j, k);
    const real dy = G.dy;
    const real q00 = dy * M.U(i, j - 1, k), q10 = dy * M.U(i + 1, j - 1, k);
    const real q01 = dy * M.U(i, j, k), q11 = dy * M.U(i + 1, j, k);
    const real uavg = RL(0.5) * (RL(0.5) * (q00 + q10) + RL(0.5) * (q01 + q11));
    const real uhat = uavg * G.rdy;
    real Gt = RL(0.0);
    if (uhat != RL(0.0)) {
        const bool left = uhat > RL(0.0);
        real z;
        if (!GEN) z = recon_vort_x<5, false>(M, i, j, k, left);
        else {
            const int N = order_from(T.vx, 5, k);
            z = OB_DISPATCH(N, (recon_vort_x<1, true>(M, i, j, k, left)), (recon_vort_x<2, true>(M, i, j, k, left)),
                            (recon_vort_x<3, true>(M, i, j, k, left)), (recon_vort_x<4, true>(M, i, j, k, left)),
                            (recon_vort_x<5, true>(M, i, j, k, left)));
        }
        Gt = -(uhat * z);
    }
    const real vh = M.V(i, j, k);
    real Phi = RL(0.0);
    if (vh != RL(0.0)) {
        const bool left = vh > RL(0.0);
        real d;
        if (!GEN) d = recon_div_y<3>(M, i, j, k, left);
        else {
            const int N = order_from(T.dy, 3, k);
            d = OB_DISPATCH(N, (recon_div_y<1>(M, i, j, k, left)), (recon_div_y<2>(M, i, j, k, left)),
                            (recon_div_y<3>(M, i, j, k, left)), (recon_div_y<3>(M, i, j, k, left)),
                            (recon_div_y<3>(M, i, j, k, left)));
        }
        Phi = vh * d + vh * RL(0.5) * (M.dxU(i, j - 1, k) + M.dxU(i, j, k));
    }
    const real azm = MET(mrow, AZCC, -1), az0 = MET(mrow, AZCC, 0);
    const real Wlo = RL(0.5) * (azm * wk[2] + az0 * wk[0]);
    const real Whi = RL(0.5) * (azm * wk[5] + az0 * wk[3]);
    real fl_lo = Wlo * RL(0.5) * (v_below + vh);
    const real fl_hi = Whi * RL(0.5) * (vh + v_above);
    if (GEN) {
        if (k > 0 && k <= T.per_v) fl_lo = RL(0.0);
    }
    Gt -= (Phi + fl_hi - fl_lo) * (MET(mrow, RAZCF, 0) * G.rdzc[k]);
    Gt -= (M.Kh(i, j, k) - M.Kh(i, j - 1, k)) * G.rdy;
    int nact = 4;
    if (GEN) {
        nact = 0;
#pragma unroll
        for (int n = 0; n < 4; n++) nact += k >= T.cv[n] ? 1 : 0;
    }
    const real f0 = MET(mrow, FFF, 0);
    const real fbar = RL(0.5) * (f0 + f0);
    const real cor = nact == 0 ? RL(0.0) : nact == 4 ? uavg : uavg / (RL(0.25) * nact);
    Gt -= fbar * cor * G.rdy;
    Gt -= (F.p[c] - F.p[c - G.pitch]) * G.rdy;
    if (TEND_LEITH && G.qgleith) {
        const MomG Mg{G, F.u, F.v};
        const real ax = dy * G.dzc[k];
        auto fccc = [&](int a, int b) { return active(G, a, b, k) ? -F.nuL[idx3(G, a, b, k)] * Mg.divxy(a, b, k) : RL(0.0); };
        auto fffc = [&](int a, int b) {
            const bool per = !active(G, a - 1, b - 1, k) || !active(G, a, b - 1, k) || !active(G, a - 1, b, k) || !active(G, a, b, k);
            if (per && in_domain(G, b - 1, k) && in_domain(G, b, k)) return RL(0.0);
            const long long q = idx3(G, a, b, k);
            const real nu = RL(0.25) * (F.nuL[q - 1 - G.pitch] + F.nuL[q - G.pitch] + F.nuL[q - 1] + F.nuL[q]);
            return nu * Mg.zeta<true>(a, b, k);
        };
        const real t = ax * (-fffc(i + 1, j)) - ax * (-fffc(i, j)) +
                         (G.dxfc[j] * G.dzc[k] * fccc(i, j) - G.dxfc[j - 1] * G.dzc[k] * fccc(i, j - 1));
        Gt -= t / (G.azcf[j] * G.dzc[k]);
    }
    if (k == G.Nz - 1 && G.bc_kind == OB200_BC_REALISTIC) Gt -= interp_flux(G, F.forcing[1], i, j, G.Ny + 1, time, false) * G.rdzc[k];
    if (k == 0) {
        real J = RL(0.0);
        if (G.bc_kind == OB200_BC_DOUBLEDRAKE) J = -G.mu_drag * vh;
        else if (G.bc_kind == OB200_BC_REALISTIC) {
            const real u0 = M.U(i, j - 1, k), u1 = M.U(i + 1, j - 1, k), u2 = M.U(i, j, k), u3 = M.U(i + 1, j, k);
            J = -G.mu_drag * vh * sqrt(vh * vh + RL(0.25) * (u0 * u0 + u2 * u2 + u1 * u1 + u3 * u3));
        }
        Gt += J * G.rdzc[k];
    }
    if (GEN && G.bc_kind == OB200_BC_REALISTIC && k > 0 && k <= T.per_v) {
        const real u0 = M.U(i, j - 1, k), u1 = M.U(i + 1, j - 1, k), u2 = M.U(i, j, k), u3 = M.U(i + 1, j, k);
        Gt += -G.mu_drag * vh * sqrt(vh * vh + RL(0.25) * (u0 * u0 + u2 * u2 + u1 * u1 + u3 * u3)) * G.rdzc[k];
    }
    return Gt;
}

// ---- shared-memory tile kernel -----------------------------------------------------------------------------
// Block = TX x TY cells of one level; u, v are staged (register-prefetched one level ahead), then zeta, I_y(u),
// I_x(v), dxU, dyV, K are computed ONCE per tile point (instead of once per stencil point per thread) and every
// thread reads its 9/9/5-point stencils from shared memory.
//   GEN = false: tiles whose every stencil is full order and mask free (tile_kfast <= k)
//   GEN = true : the remaining (tile, level) pairs near ridges, walls and the sea floor: conditional zeta,
//                per-reconstruction order reduction, masked outputs.
#ifndef MOM_MINB
#define MOM_MINB (MOM_DB ? 3 : 4)   // 3 (80 registers, no spills) measured slower than 4 (64, small spills): 18.98 vs 18.81 ms; tracers 11.76 vs 11.51
#endif
#ifndef MOM_KNBR
#define MOM_KNBR 1   // own-column u, v at k - 1 / k + 1 from registers / the next TMA tile instead of global memory
#endif
#ifndef TRC_MINB
#define TRC_MINB 4
#endif
#ifndef GEN_MINB
#define GEN_MINB 2   // masked-tile instances: resident blocks per SM
#endif
//   TMA = true : the u, v tiles of level k+1 are fetched by one thread with cp.async.bulk.tensor into the second
//                shared-memory buffer while level k is being computed (mbarrier completion); false: register staging
template <int TX, int TY, bool GEN, bool TMA>
__global__ void __launch_bounds__(TX * TY, GEN ? GEN_MINB : MOM_MINB) k_momentum_tiled(Geom G, Fields F, double time, const int* __restrict__ tile_kfast,
                                                             const int* __restrict__ slow_tiles, int ntx, int klevels,
                                                             int NV, int ND, const __grid_constant__ CUtensorMap tmu,
                                                             const __grid_constant__ CUtensorMap tmv) {
    using D = TileDims<TX, TY>;
    constexpr int NT = TX * TY;
    constexpr int RU = (D::NU + NT - 1) / NT;
    extern __shared__ __align__(128) real sm[];
    __shared__ unsigned long long bar[2];
    constexpr bool DB = MOM_DB && TMA;          // one barrier per level (derived fields double-buffered)
    real* der = sm + 4 * D::NUP;  // [u0 | v0 | u1 | v1] tile buffers first (128-byte aligned slots), then the derived fields
    real* smt = der + (MOM_DB ? 2 : 1) * D::NDER;       // metrics table [MT_N][MT_ROWS]
    real* swr = smt + MT_N * MT_ROWS;                   // w ring [2][NW] (MOM_WSM)
    int tile, kf;
    int kskip = 0;   // masked tiles: levels below the shallowest sea floor of the tile are all solid -> G stays 0 (never written)
    if (GEN) { tile = slow_tiles[3 * blockIdx.x]; kf = slow_tiles[3 * blockIdx.x + 1]; kskip = slow_tiles[3 * blockIdx.x + 2]; }
    else { tile = blockIdx.y * gridDim.x + blockIdx.x; kf = tile_kfast[tile]; }
    const int i0 = (tile % ntx) * TX, j0 = (tile / ntx) * TY;
    const int tid = threadIdx.y * TX + threadIdx.x;
    const int zc = GEN ? blockIdx.y : blockIdx.z;
    const int k_begin = GEN ? max(zc * klevels, kskip) : max(zc * klevels, kf);
    const int k_end = GEN ? min(min((zc + 1) * klevels, kf), G.Nz) : min((zc + 1) * klevels, G.Nz);
    if (k_begin >= k_end) return;
    const int i = i0 + threadIdx.x, j = j0 + threadIdx.y;
    const bool mine = i < G.Nx && j < G.Ny;
    const int tx0 = i0 - D::HUX + OB_X0, ty0 = j0 - D::HU + OB_H;   // tile origin in array coordinates (even column)
    constexpr unsigned TILE_BYTES = D::NU * sizeof(real);
    load_metric_table(G, smt, j0, tid, NT);
    const real* mrow = smt + (threadIdx.y + MT_J0);               // this thread's row j: MET(mrow, X, dj) = X[j + dj]
    // staging map: thread (tx, ty) stages the tile points (tx + TX p, ty + TY r) -- every smem offset below is this
    // thread's base plus a compile-time constant (no div / mod, no per-element index arithmetic in the level loop)
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int ez0 = ty * D::PZ + tx, qz0 = (ty + D::HU - D::HZ) * D::PU + (tx + D::HUX - D::HZ);
    const int ed0 = ty * D::PD + tx, qd0 = (ty + D::HU - D::HD) * D::PU + (tx + D::HUX - D::HD);
    const int ek0 = ty * D::PK + tx, qk0 = (ty + D::HU - 1) * D::PU + (tx + D::HUX - 1);
    constexpr int NRXZ = (D::PZ + TX - 1) / TX, NRYZ = (D::RZ + TY - 1) / TY;
    constexpr int NRXD = (D::PD + TX - 1) / TX, NRYD = (D::RD + TY - 1) / TY;
    constexpr int NRXK = (D::PK + TX - 1) / TX, NRYK = (D::RK + TY - 1) / TY;
    int offu[TMA ? 1 : RU];
    real ru[TMA ? 1 : RU], rv[TMA ? 1 : RU];
    auto prefetch = [&](int k) {
        const long long lev = (long long)k * G.sz;
#pragma unroll
        for (int r = 0; r < (TMA ? 0 : RU); r++) { ru[r] = F.u[offu[r] + lev]; rv[r] = F.v[offu[r] + lev]; }
    };
    if (TMA) {
        if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&bar[0], 2 * TILE_BYTES);
            tma_load_tile(sm, &tmu, &bar[0], tx0, ty0, k_begin);
            tma_load_tile(sm + D::NUP, &tmv, &bar[0], tx0, ty0, k_begin);
            if (DB && k_begin + 1 < k_end) {      // both tile buffers are free at the start: the second level follows at once
                mbar_expect_tx(&bar[1], 2 * TILE_BYTES);
                tma_load_tile(sm + 2 * D::NUP, &tmu, &bar[1], tx0, ty0, k_begin + 1);
                tma_load_tile(sm + 3 * D::NUP, &tmv, &bar[1], tx0, ty0, k_begin + 1);
            }
        }
    } else {
#pragma unroll
        for (int r = 0; r < (TMA ? 0 : RU); r++) {
            const int e = min(tid + r * NT, D::NU - 1);
            const int a = e % D::PU, b = e / D::PU;
            offu[r] = (int)idx3(G, min(i0 - D::HUX + a, G.Nx + OB_H - 1), min(j0 - D::HU + b, G.Ny + OB_H), 0);
        }
        prefetch(k_begin);
    }
    // masked tiles: level thresholds of this thread's column and of the zeta points it stages (k independent)
    GenThr T;
    int zthr[GEN ? NRXZ * NRYZ : 1];
    if (GEN) {
        if (mine) gen_thresholds(G, i, j, NV, ND, T);
#pragma unroll
        for (int r = 0; r < NRXZ * NRYZ; r++) {
            const int a = min(tx + TX * (r % NRXZ), D::PZ - 1), b = min(ty + TY * (r / NRXZ), D::RZ - 1);
            const int ic = min(i0 - D::HZ + a, G.Nx + OB_H - 1), jc = min(j0 - D::HZ + b, G.Ny + OB_H);
            const int k00 = kb_at(G, ic - 1, jc - 1), k10 = kb_at(G, ic, jc - 1), k01 = kb_at(G, ic - 1, jc), k11 = kb_at(G, ic, jc);
            const int kzx = max(min(k00, k01), min(k10, k11));   // delta_x v touches an inactive (c,f,c) node below this level
            const int kzy = max(min(k00, k10), min(k01, k11));   // delta_y u touches an inactive (f,c,c) node below this level
            zthr[r] = (kzx << 16) | kzy;
        }
    }
    // own-column u, v one level below (clamped at the bottom): read once, then carried from level to level; the level above
    // comes from the tile that TMA is already fetching for the next iteration (MOM_KNBR = 0: four global loads per cell)
    const int own = (threadIdx.y + D::HU) * D::PU + (threadIdx.x + D::HUX);
    real ub = RL(0.0), vb = RL(0.0);
    if (MOM_KNBR && mine) {
        const long long cb = idx3(G, i, j, max(k_begin - 1, 0));
        ub = F.u[cb]; vb = F.v[cb];
    }
    // w ring: thread (tx, ty) stages w of its own cell at slot (tx + 1, ty + 1); the tx = 0 / ty = 0 threads add the column
    // i0 - 1 / the row j0 - 1
    const int wown = (threadIdx.y + 1) * D::PW + (threadIdx.x + 1);
    auto stage_w_load = [&](int kk, real& a, real& b, real& c3) {
        const long long cw = idx3(G, i, j, kk);
        a = F.w[cw];
        if (threadIdx.x == 0) b = F.w[cw - 1];
        if (threadIdx.y == 0) c3 = F.w[cw - G.pitch];
    };
    auto stage_w_store = [&](real* dst, real a, real b, real c3) {
        dst[wown] = a;
        if (threadIdx.x == 0) dst[wown - 1] = b;
        if (threadIdx.y == 0) dst[wown - D::PW] = c3;
    };
    if (MOM_WSM && mine) {   // level k_begin; made visible by the first mid-level barrier
        real a = RL(0.0), b = RL(0.0), c3 = RL(0.0);
        stage_w_load(k_begin, a, b, c3);
        stage_w_store(swr, a, b, c3);
    }
    int it = 0;
    for (int k = k_begin; k < k_end; k++, it++) {
        const int buf = TMA ? (it & 1) : 0;
        real* su = sm + buf * 2 * D::NUP;
        real* sv = su + D::NUP;
        real* sz = der + (DB ? buf : 0) * D::NDER;
        real* sus = sz + D::NZ;
        real* svs = sus + D::NZ;
        real* sdx = svs + D::NZ;
        real* sdy = sdx + D::ND;
        real* sK = sdy + D::ND;
        const MomS<TX, TY> M{su, sv, sz, sus, svs, sdx, sdy, sK, i0, j0};
        if (TMA) {
            // the other buffer was last read in the previous iteration, which ended with a block barrier
            if (!DB && tid == 0 && k + 1 < k_end) {
                fence_proxy_async();
                real* nu = sm + (buf ^ 1) * 2 * D::NUP;
                mbar_expect_tx(&bar[buf ^ 1], 2 * TILE_BYTES);
                tma_load_tile(nu, &tmu, &bar[buf ^ 1], tx0, ty0, k + 1);
                tma_load_tile(nu + D::NUP, &tmv, &bar[buf ^ 1], tx0, ty0, k + 1);
            }
            mbar_wait(&bar[buf], (it >> 1) & 1);
        } else {
#pragma unroll
            for (int r = 0; r < (TMA ? 0 : RU); r++) {
                const int e = tid + r * NT;
                if (e < D::NU) { su[e] = ru[r]; sv[e] = rv[r]; }
            }
            __syncthreads();
            if (k + 1 < k_end) prefetch(k + 1);
        }
        const real dzk = G.dzc[k];
        // w(k + 1) of the tile: requested here, stored at the end of this phase into the slot w(k - 1) occupied
        real wa = RL(0.0), wb = RL(0.0), wc = RL(0.0);
        if (MOM_WSM && mine) stage_w_load(k + 1, wa, wb, wc);
        // zeta, I_y u, I_x v on [i0 - 4, i0 + TX + 4] x [j0 - 4, j0 + TY + 4]
#pragma unroll
        for (int r = 0; r < NRYZ; r++)
#pragma unroll
            for (int p = 0; p < NRXZ; p++) {
                if ((p == 0 || tx + TX * p < D::PZ) && (r == 0 || ty + TY * r < D::RZ)) {
                    const int e = ez0 + TX * p + TY * r * D::PZ, q = qz0 + TX * p + TY * r * D::PU;
                    const real* mz = smt + (ty + TY * r + MT_J0 - D::HZ);     // row jj = j0 - HZ + b
                    real dxv = G.dy * sv[q] - G.dy * sv[q - 1];
                    real dyu = MET(mz, DXFC, 0) * su[q] - MET(mz, DXFC, -1) * su[q - D::PU];
                    if (GEN) {   // conditional differences: a difference that touches an inactive node is zero
                        const int th = zthr[r * NRXZ + p];
                        if (k < (th >> 16)) dxv = RL(0.0);
                        if (k < (th & 0xffff)) dyu = RL(0.0);
                    }
                    sz[e] = (dxv - dyu) * MET(mz, RAZCF, 0);
                    sus[e] = RL(0.5) * (su[q - D::PU] + su[q]);
                    svs[e] = RL(0.5) * (sv[q - 1] + sv[q]);
                }
            }
        // dx(Ax u), dy(Ay v) on [i0 - 3, i0 + TX + 2] x [j0 - 3, j0 + TY + 2]
#pragma unroll
        for (int r = 0; r < NRYD; r++)
#pragma unroll
            for (int p = 0; p < NRXD; p++) {
                if ((p == 0 || tx + TX * p < D::PD) && (r == 0 || ty + TY * r < D::RD)) {
                    const int e = ed0 + TX * p + TY * r * D::PD, q = qd0 + TX * p + TY * r * D::PU;
                    const real* md = smt + (ty + TY * r + MT_J0 - D::HD);     // row jj = j0 - HD + b
                    const real ax = G.dy * dzk;
                    sdx[e] = ax * su[q + 1] - ax * su[q];
                    sdy[e] = MET(md, DXCF, 1) * dzk * sv[q + D::PU] - MET(md, DXCF, 0) * dzk * sv[q];
                }
            }
        // K on [i0 - 1, i0 + TX - 1] x [j0 - 1, j0 + TY - 1]
#pragma unroll
        for (int r = 0; r < NRYK; r++)
#pragma unroll
            for (int p = 0; p < NRXK; p++) {
                if ((p == 0 || tx + TX * p < D::PK) && (r == 0 || ty + TY * r < D::RK)) {
                    const int e = ek0 + TX * p + TY * r * D::PK, q = qk0 + TX * p + TY * r * D::PU;
                    const real ua = su[q], ub = su[q + 1], va = sv[q], vb = sv[q + D::PU];
                    sK[e] = (RL(0.5) * (ua * ua + ub * ub) + RL(0.5) * (va * va + vb * vb)) / RL(2.0);
                }
            }
        if (MOM_WSM && mine) stage_w_store(swr + ((it + 1) & 1) * D::NW, wa, wb, wc);
        __syncthreads();
        if (DB && tid == 0 && it >= 1 && k + 1 < k_end) {
            // every warp is past the compute phase of level k - 1, the last reader of the other tile buffer: refill it
            fence_proxy_async();
            real* nu = sm + (buf ^ 1) * 2 * D::NUP;
            mbar_expect_tx(&bar[buf ^ 1], 2 * TILE_BYTES);
            tma_load_tile(nu, &tmu, &bar[buf ^ 1], tx0, ty0, k + 1);
            tma_load_tile(nu + D::NUP, &tmv, &bar[buf ^ 1], tx0, ty0, k + 1);
        }
        if (mine) {
            const long long c = idx3(G, i, j, k);
            real ua, va;
            if (MOM_KNBR && TMA && k + 1 < k_end) {
                // the next level's tiles were requested at the top of this iteration: by now they have normally landed
                mbar_wait(&bar[buf ^ 1], ((it + 1) >> 1) & 1);
                const real* nu = sm + (buf ^ 1) * 2 * D::NUP;
                ua = nu[own]; va = nu[D::NUP + own];
            } else {
                const long long ca = idx3(G, i, j, min(k + 1, G.Nz - 1));
                ua = F.u[ca]; va = F.v[ca];
            }
            if (!MOM_KNBR) { const long long cb = idx3(G, i, j, max(k - 1, 0)); ub = F.u[cb]; vb = F.v[cb]; }
            real wk[6];
            if (MOM_WSM) {
                const real* w0 = swr + (it & 1) * D::NW + wown;
                const real* w1 = swr + ((it + 1) & 1) * D::NW + wown;
                wk[0] = w0[0]; wk[1] = w0[-1]; wk[2] = w0[-D::PW];
                wk[3] = w1[0]; wk[4] = w1[-1]; wk[5] = w1[-D::PW];
            } else {
                wk[0] = F.w[c]; wk[1] = F.w[c - 1]; wk[2] = F.w[c - G.pitch];
                wk[3] = F.w[c + G.sz]; wk[4] = F.w[c - 1 + G.sz]; wk[5] = F.w[c - G.pitch + G.sz];
            }
            if (!GEN) {
                store_tendency(G, F, 0, c, tend_u<false>(G, F, M, i, j, k, 5, 3, time, T, mrow, ub, ua, wk));
                store_tendency(G, F, 1, c, tend_v<false>(G, F, M, i, j, k, 5, 3, time, T, mrow, vb, va, wk));
            } else if (k >= T.kfast) {
                // most cells of a masked tile are still far from any boundary: full-order, check-free path
                store_tendency(G, F, 0, c, tend_u<false>(G, F, M, i, j, k, 5, 3, time, T, mrow, ub, ua, wk));
                store_tendency(G, F, 1, c, tend_v<false>(G, F, M, i, j, k, 5, 3, time, T, mrow, vb, va, wk));
            } else {
                store_tendency(G, F, 0, c, k < T.per_u ? RL(0.0) : tend_u<true>(G, F, M, i, j, k, NV, ND, time, T, mrow, ub, ua, wk));
                store_tendency(G, F, 1, c, k < T.per_v ? RL(0.0) : tend_v<true>(G, F, M, i, j, k, NV, ND, time, T, mrow, vb, va, wk));
            }
            if (MOM_KNBR) { ub = su[own]; vb = sv[own]; }
        }
        if (!DB) __syncthreads();
    }
}

// ================================================================================================
// tracer tendencies (T and S share the velocity loads)
// ================================================================================================
// ---- fast path: T and S tiles in shared memory, every face flux computed ONCE (the reference computes each
// face twice, once per adjacent cell), upwind side only, then each thread differences the fluxes of its cell.
// TRC_PIPE = 1 (TMA instantiations): fluxes and face velocities alternate between two shared-memory slots, so the barrier that
// ended a level goes away (two block barriers per level instead of three); the own-column values of the vertical flux come
// from registers (k - 1: the centre value of the previous level; w(k): the previous level's w(k + 1)) and from the tile TMA is
// already fetching for the next level (k + 1) -- one global load per cell and level instead of six.
#ifndef TRC_PIPE
#define TRC_PIPE 0   // measured at C3 on one B200: 12.76 ms against 11.93 ms without (the carried values spill at 64 registers)
#endif
template <int TX, int TY>
struct TracerTile {
    static constexpr int HC = 4;                                   // WENO-7 halo
    static constexpr int PC = TX + 2 * HC, RC = TY + 2 * HC;       // cells [i0-4, i0+TX+3] x [j0-4, j0+TY+3]
    static constexpr int NC = PC * RC;
    static constexpr int NCP = (NC + 15) / 16 * 16;                // tile slot padded to 128 bytes
    static constexpr int NFX = (TX + 1) * TY, NFY = TX * (TY + 1); // faces per tracer
    static constexpr int TOTAL = 4 * NCP + (TRC_PIPE ? 2 : 1) * 3 * (NFX + NFY);  // T, S tiles double-buffered + fluxes + face velocities
};

template <int TX, int TY, bool GEN, bool TMA>
__global__ void __launch_bounds__(TX * TY, GEN ? GEN_MINB : TRC_MINB) k_tracers_tiled(Geom G, Fields F, double time, const int* __restrict__ tile_kfast,
                                                            const int* __restrict__ slow_tiles, int ntx, int klevels, int NTR,
                                                            const __grid_constant__ CUtensorMap tmT,
                                                            const __grid_constant__ CUtensorMap tmS) {
    using D = TracerTile<TX, TY>;
    constexpr int NT = TX * TY;
    constexpr int RC = (D::NC + NT - 1) / NT;                 // tile elements per thread (per tracer)
    constexpr int NFV = D::NFX + D::NFY;                      // face velocities per level
    constexpr int RV = (NFV + NT - 1) / NT;
    extern __shared__ __align__(128) real sm[];
    __shared__ unsigned long long bar[2];
    constexpr bool PIPE = TRC_PIPE && TMA;
    real* sfx0 = sm + 4 * D::NCP;        // after the [T0 | S0 | T1 | S1] tile buffers: slots of [2 tracers][(TX+1)*TY x-faces | TX*(TY+1) y-faces]
    real* svel0 = sfx0 + (TRC_PIPE ? 2 : 1) * 2 * NFV;   // slots of [NFX + NFY] face velocities u (x-faces) then v (y-faces)
    int tile, kf;
    int kskip = 0;   // masked tiles: levels below the shallowest sea floor of the tile are all solid -> G stays 0 (never written)
    if (GEN) { tile = slow_tiles[3 * blockIdx.x]; kf = slow_tiles[3 * blockIdx.x + 1]; kskip = slow_tiles[3 * blockIdx.x + 2]; }
    else { tile = blockIdx.y * gridDim.x + blockIdx.x; kf = tile_kfast[tile]; }
    const int i0 = (tile % ntx) * TX, j0 = (tile / ntx) * TY;
    const int tid = threadIdx.y * TX + threadIdx.x;
    const int zc = GEN ? blockIdx.y : blockIdx.z;
    const int k_begin = GEN ? max(zc * klevels, kskip) : max(zc * klevels, kf);
    const int k_end = GEN ? min(min((zc + 1) * klevels, kf), G.Nz) : min((zc + 1) * klevels, G.Nz);
    if (k_begin >= k_end) return;
    const int i = i0 + threadIdx.x, j = j0 + threadIdx.y;
    const bool mine = i < G.Nx && j < G.Ny;
    // per-thread global offsets (within a level) of the tile elements / face velocities this thread stages
    int offc[TMA ? 1 : RC], offv[RV];   // element offsets within one level (a plane is < 2^31 elements)
#pragma unroll
    for (int r = 0; r < (TMA ? 0 : RC); r++) {
        const int e = min(tid + r * NT, D::NC - 1);
        const int a = e % D::PC, b = e / D::PC;
        offc[r] = (int)idx3(G, min(i0 - D::HC + a, G.Nx + OB_H - 1), min(j0 - D::HC + b, G.Ny + OB_H), 0);
    }
#pragma unroll
    for (int r = 0; r < RV; r++) {
        const int e = min(tid + r * NT, NFV - 1);
        if (e < D::NFX) {
            const int a = e % (TX + 1), b = e / (TX + 1);
            offv[r] = (int)idx3(G, min(i0 + a, G.Nx + OB_H - 1), min(j0 + b, G.Ny - 1), 0);
        } else {
            const int f = e - D::NFX, a = f % TX, b = f / TX;
            offv[r] = (int)idx3(G, min(i0 + a, G.Nx + OB_H - 1), min(j0 + b, G.Ny), 0);
        }
    }
    real rT[TMA ? 1 : RC], rS[TMA ? 1 : RC], rV[RV];
    const int tx0 = i0 - D::HC + OB_X0, ty0 = j0 - D::HC + OB_H;   // tile origin in array coordinates
    constexpr unsigned TILE_BYTES = D::NC * sizeof(real);
    auto prefetch = [&](int k) {   // issue the loads of level k; consumed one iteration later
        const long long lev = (long long)k * G.sz;
#pragma unroll
        for (int r = 0; r < (TMA ? 0 : RC); r++) { rT[r] = F.T[offc[r] + lev]; rS[r] = F.S[offc[r] + lev]; }
#pragma unroll
        for (int r = 0; r < RV; r++) {
            const int e = min(tid + r * NT, NFV - 1);
            rV[r] = (e < D::NFX ? F.u : F.v)[offv[r] + lev];
        }
    };
    if (TMA) {
        if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&bar[0], 2 * TILE_BYTES);
            tma_load_tile(sm, &tmT, &bar[0], tx0, ty0, k_begin);
            tma_load_tile(sm + D::NCP, &tmS, &bar[0], tx0, ty0, k_begin);
        }
    }
    prefetch(k_begin);
    const long long cbase = mine ? idx3(G, i, j, 0) : 0;
    const real az = G.azcc[min(j, G.Ny - 1)], raz = G.razcc[min(j, G.Ny - 1)];
    // masked tiles: per-face level thresholds of the reconstruction order (packed 16 bit), k independent
    constexpr int RFE = (NFV + NT - 1) / NT;
    int fthr[GEN ? 2 * RFE : 1];
    int kb_own = 0;
    if (GEN) {
        kb_own = mine ? kb_at(G, i, j) : 0;
#pragma unroll
        for (int r = 0; r < RFE; r++) {
            const int f = min(tid + r * NT, NFV - 1);
            const bool isx = f < D::NFX;
            const int g = isx ? f : f - D::NFX;
            const int w = isx ? TX + 1 : TX;
            const int a = g % w, b = g / w;
            const int ii = min(i0 + a, G.Nx + OB_H - 1), jj = min(j0 + b, isx ? G.Ny - 1 : G.Ny);
            int t = 0, thr[5];
#pragma unroll
            for (int N = 1; N <= 4; N++) {   // order N adds the cells -N and N-1 along the face normal
                t = isx ? max(t, max(kb_at(G, ii - N, jj), kb_at(G, ii + N - 1, jj)))
                        : max(t, max(kb_at(G, ii, jj - N), kb_at(G, ii, jj + N - 1)));
                thr[N] = N <= NTR ? t : 0x7fff;
            }
            fthr[2 * r] = (thr[2] << 16) | thr[1];
            fthr[2 * r + 1] = (thr[4] << 16) | thr[3];
        }
    }
    // PIPE: own-column values carried from level to level
    const int qown = (threadIdx.y + D::HC) * D::PC + (threadIdx.x + D::HC);
    real wlo_c = RL(0.0), Tm_c = RL(0.0), Sm_c = RL(0.0);
    if (PIPE && mine) {
        const long long cm = cbase + (long long)max(k_begin - 1, 0) * G.sz;
        wlo_c = F.w[cbase + (long long)k_begin * G.sz]; Tm_c = F.T[cm]; Sm_c = F.S[cm];
    }
    int it = 0;
    for (int k = k_begin; k < k_end; k++, it++) {
        const int buf = TMA ? (it & 1) : 0;
        real* sc = sm + buf * 2 * D::NCP;   // T tile, then S tile at + NCP
        real* sfx = sfx0 + (PIPE ? buf : 0) * 2 * NFV;
        real* svel = svel0 + (PIPE ? buf : 0) * NFV;
        if (TMA && tid == 0 && k + 1 < k_end) {
            // the other buffer was last read before the second barrier of the previous iteration (PIPE: the centre values are
            // taken in the flux phase), otherwise the previous iteration ended with a block barrier
            fence_proxy_async();
            real* nx = sm + (buf ^ 1) * 2 * D::NCP;
            mbar_expect_tx(&bar[buf ^ 1], 2 * TILE_BYTES);
            tma_load_tile(nx, &tmT, &bar[buf ^ 1], tx0, ty0, k + 1);
            tma_load_tile(nx + D::NCP, &tmS, &bar[buf ^ 1], tx0, ty0, k + 1);
        }
#pragma unroll
        for (int r = 0; r < (TMA ? 0 : RC); r++) {
            const int e = tid + r * NT;
            if (e < D::NC) { sc[e] = rT[r]; sc[D::NCP + e] = rS[r]; }
        }
#pragma unroll
        for (int r = 0; r < RV; r++) {
            const int e = tid + r * NT;
            if (e < NFV) svel[e] = rV[r];
        }
        __syncthreads();
        if (TMA) mbar_wait(&bar[buf], (it >> 1) & 1);
        if (k + 1 < k_end) prefetch(k + 1);
        // own-column values for the vertical flux, requested now, consumed after the flux phase
        const int km = max(k - 1, 0), kp = min(k + 1, G.Nz - 1);
        real wlo = 0, whi = 0, Tm = 0, Tp = 0, Sm = 0, Sp = 0, ckT = 0, ckS = 0;
        if (mine) {
            const long long c = cbase + (long long)k * G.sz;
            whi = F.w[c + G.sz];
            if (PIPE) { wlo = wlo_c; Tm = Tm_c; Sm = Sm_c; ckT = sc[qown]; ckS = sc[D::NCP + qown]; }
            else {
                wlo = F.w[c];
                Tm = F.T[cbase + (long long)km * G.sz]; Tp = F.T[cbase + (long long)kp * G.sz];
                Sm = F.S[cbase + (long long)km * G.sz]; Sp = F.S[cbase + (long long)kp * G.sz];
            }
        }
        const real dzk = G.dzc[k], rdzk = G.rdzc[k];
        if (!GEN) {
            // own cell: west face (x) and south face (y) of both tracers, compile-time strides
            {
                const int q = (threadIdx.y + D::HC) * D::PC + (threadIdx.x + D::HC);
                const int fx = threadIdx.y * (TX + 1) + threadIdx.x, fy = D::NFX + threadIdx.y * TX + threadIdx.x;
                const real ux = svel[fx], vy = svel[fy];
                const real mx = G.dy * dzk, my = G.dxcf[min(j, G.Ny)] * dzk;
#pragma unroll
                for (int tr = 0; tr < 2; tr++) {
                    const real* cc = sc + tr * D::NCP;
                    real flx = RL(0.0), fly = RL(0.0);
                    if (ux != RL(0.0)) flx = G.dy * dzk * (ux * recon7_window<1>(cc, q, ux > RL(0.0)));
                    if (vy != RL(0.0)) fly = G.dxcf[min(j, G.Ny)] * dzk * (vy * recon7_window<D::PC>(cc, q, vy > RL(0.0)));
                    sfx[tr * NFV + fx] = flx;
                    sfx[tr * NFV + fy] = fly;
                }
                (void)mx; (void)my;
            }
            // the tile's east column of x-faces and north row of y-faces: (TY + TX) faces x 2 tracers
            if (tid < 2 * (TX + TY)) {
                const int tr = tid / (TX + TY), g = tid - tr * (TX + TY);
                const real* cc = sc + tr * D::NCP;
                const bool isx = g < TY;
                const int a = isx ? TX : g - TY, b = isx ? g : TY;
                const int f = isx ? b * (TX + 1) + a : D::NFX + b * TX + a;
                const int q = (b + D::HC) * D::PC + (a + D::HC);
                const real vel = svel[f];
                real flux = RL(0.0);
                if (vel != RL(0.0)) {
                    const real r = isx ? recon7_window<1>(cc, q, vel > RL(0.0)) : recon7_window<D::PC>(cc, q, vel > RL(0.0));
                    flux = (isx ? G.dy : G.dxcf[min(j0 + b, G.Ny)]) * dzk * (vel * r);
                }
                sfx[tr * NFV + f] = flux;
            }
        } else {
            // flat work list over (direction, face): every face flux once, upwind side only.  The direction only
            // changes the stride / metric, so x- and y-faces share one code path; the reconstruction order of a face
            // comes from its precomputed level thresholds (order 0 = peripheral face, no flux).
#pragma unroll
            for (int r = 0; r < RFE; r++) {
                const int f = tid + r * NT;
                if (f < NFV) {
                    const real vel = svel[f];
                    const bool isx = f < D::NFX;
                    const int g = isx ? f : f - D::NFX;
                    const int w = isx ? TX + 1 : TX;
                    const int a = g % w, b = g / w;                      // face between cells (a-1,b),(a,b) or (a,b-1),(a,b)
                    const int stride = isx ? 1 : D::PC;
                    const int t12 = fthr[2 * r], t34 = fthr[2 * r + 1];
                    int N = 0;
                    if (k >= (t12 & 0xffff)) N = 1;
                    if (k >= (t12 >> 16)) N = 2;
                    if (k >= (t34 & 0xffff)) N = 3;
                    if (k >= (t34 >> 16)) N = 4;
                    real flux[2] = {RL(0.0), RL(0.0)};
                    if (vel != RL(0.0) && N > 0) {
                        const bool left = vel > RL(0.0);
                        const int q = (b + D::HC) * D::PC + (a + D::HC);   // the cell on the high side of the face
                        const int jj = min(j0 + b, isx ? G.Ny - 1 : G.Ny);
                        const real metric = isx ? G.dy : G.dxcf[jj];
                        const long long st = left ? stride : -stride;
                        const long long bq = left ? q - N * stride : q + (N - 1) * stride;
#pragma unroll
                        for (int tr = 0; tr < 2; tr++) {
                            const real* cc = sc + tr * D::NCP;
                            const real rr = OB_DISPATCH(N, recon_cell<1>(cc, bq, st), recon_cell<2>(cc, bq, st),
                                                          recon_cell<3>(cc, bq, st), recon_cell<4>(cc, bq, st), recon_cell<4>(cc, bq, st));
                            flux[tr] = metric * dzk * (vel * rr);
                        }
                    }
                    sfx[f] = flux[0];   // [tracer][x-faces | y-faces]
                    sfx[NFV + f] = flux[1];
                }
            }
        }
        __syncthreads();
        if (mine) {
            const long long c = cbase + (long long)k * G.sz;
            const int fxi = threadIdx.y * (TX + 1) + threadIdx.x, fyi = threadIdx.y * TX + threadIdx.x;
            const int q = qown;
            if (PIPE) {
                if (k + 1 < k_end) {   // the next level's tiles, requested at the top of this iteration
                    mbar_wait(&bar[buf ^ 1], ((it + 1) >> 1) & 1);
                    const real* nx = sm + (buf ^ 1) * 2 * D::NCP;
                    Tp = nx[q]; Sp = nx[D::NCP + q];
                } else {
                    Tp = F.T[cbase + (long long)kp * G.sz]; Sp = F.S[cbase + (long long)kp * G.sz];
                }
                wlo_c = whi; Tm_c = ckT; Sm_c = ckS;
            }
#pragma unroll
            for (int tr = 0; tr < 2; tr++) {
                const real ck = PIPE ? (tr == 0 ? ckT : ckS) : sc[tr * D::NCP + q];
                const real cm = tr == 0 ? Tm : Sm, cp = tr == 0 ? Tp : Sp;
                real fzl = az * wlo * RL(0.5) * (cm + ck);
                const real fzh = az * whi * RL(0.5) * (ck + cp);
                if (GEN && k > 0 && k <= kb_own) fzl = RL(0.0);   // immersed-peripheral bottom face
                const real fxw = sfx[tr * NFV + fxi], fxe = sfx[tr * NFV + fxi + 1];
                const real fys = sfx[tr * NFV + D::NFX + fyi], fyn = sfx[tr * NFV + D::NFX + fyi + TX];
                real Gt = -((fxe - fxw) + (fyn - fys) + (fzh - fzl)) * (raz * rdzk);
                if (k == G.Nz - 1) {
                    const real J = tr == 0 ? top_flux_T(G, F, i, j, ck, time) : top_flux_S(G, F, i, j, ck, time);
                    Gt -= J * rdzk;
                }
                if (GEN && k < kb_own) Gt = RL(0.0);
                store_tendency(G, F, 2 + tr, c, Gt);
            }
        }
        if (!PIPE) __syncthreads();
    }
}

// 16 x 16 tiles: same speed as 32 x 8 on the full grid, but 270-column slabs (8 GPUs) fill 17 tiles to 99 %
#ifndef MOM_TX
#define MOM_TX 16
#endif
#ifndef MOM_TY
#define MOM_TY 16
#endif
#ifndef TEND_KLEV
#define TEND_KLEV 20   // levels marched by one block (measured at C3: 8 -> 18.85 + 11.62 ms, 20 -> 18.44 + 11.39, 60 -> 18.66 + 11.78)
#endif
// `which`: bit 0 = the full-order tiles, bit 1 = the masked tiles (disjoint (tile, level) sets of G: the two instances may
// run concurrently on different streams)
template <bool TMA>
static void launch_momentum_t(const Geom& G, const Fields& F, const MomPlan& P, int ord_vort, int ord_div, double time, cudaStream_t st, int which) {
    using D = TileDims<MOM_TX, MOM_TY>;
    static bool attr_set[64] = {};   // the opt-in to > 48 KB of dynamic shared memory is per device
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t smem = D::TOTAL * sizeof(real);
    if (!attr_set[dev & 63]) {
        cudaFuncSetAttribute(k_momentum_tiled<MOM_TX, MOM_TY, false, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_momentum_tiled<MOM_TX, MOM_TY, true, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set[dev & 63] = true;
    }
    const int NV = (ord_vort + 1) / 2, ND = (ord_div + 1) / 2, nz = (G.Nz + TEND_KLEV - 1) / TEND_KLEV;
    dim3 b(MOM_TX, MOM_TY);
    if (P.any_fast && (which & 1)) {
        k_momentum_tiled<MOM_TX, MOM_TY, false, TMA><<<dim3(P.ntx, P.nty, nz), b, smem, st>>>(
            G, F, time, P.tile_kfast, P.slow_tiles, P.ntx, TEND_KLEV, NV, ND, P.tm_u, P.tm_v);
        OB_COUNT_LAUNCH(1);
    }
    if (P.nslow > 0 && (which & 2)) {
        k_momentum_tiled<MOM_TX, MOM_TY, true, TMA><<<dim3(P.nslow, nz), b, smem, st>>>(
            G, F, time, P.tile_kfast, P.slow_tiles, P.ntx, TEND_KLEV, NV, ND, P.tm_u, P.tm_v);
        OB_COUNT_LAUNCH(1);
    }
}
void launch_momentum(const Geom& G, const Fields& F, const MomPlan& P, int ord_vort, int ord_div, double time, cudaStream_t st, int which) {
    if (P.use_tma) launch_momentum_t<true>(G, F, P, ord_vort, ord_div, time, st, which);
    else launch_momentum_t<false>(G, F, P, ord_vort, ord_div, time, st, which);
}
void mom_plan_tile_shape(int* tx, int* ty) { *tx = MOM_TX; *ty = MOM_TY; }
void tend_tile_boxes(int* mom_x, int* mom_y, int* trc_x, int* trc_y) {
    *mom_x = TileDims<MOM_TX, MOM_TY>::PU; *mom_y = TileDims<MOM_TX, MOM_TY>::RU;
    *trc_x = TracerTile<MOM_TX, MOM_TY>::PC; *trc_y = TracerTile<MOM_TX, MOM_TY>::RC;
}
template <bool TMA>
static void launch_tracers_t(const Geom& G, const Fields& F, const MomPlan& P, int ord_tracer, double time, cudaStream_t st, int which) {
    using D = TracerTile<MOM_TX, MOM_TY>;
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t smem = D::TOTAL * sizeof(real);
    if (!attr_set[dev & 63]) {
        cudaFuncSetAttribute(k_tracers_tiled<MOM_TX, MOM_TY, false, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_tracers_tiled<MOM_TX, MOM_TY, true, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set[dev & 63] = true;
    }
    const int NTR = (ord_tracer + 1) / 2, nz = (G.Nz + TEND_KLEV - 1) / TEND_KLEV;
    dim3 b(MOM_TX, MOM_TY);
    if (P.any_fast && (which & 1)) {
        k_tracers_tiled<MOM_TX, MOM_TY, false, TMA><<<dim3(P.ntx, P.nty, nz), b, smem, st>>>(
            G, F, time, P.tile_kfast, P.slow_tiles, P.ntx, TEND_KLEV, NTR, P.tm_T, P.tm_S);
        OB_COUNT_LAUNCH(1);
    }
    if (P.nslow > 0 && (which & 2)) {
        k_tracers_tiled<MOM_TX, MOM_TY, true, TMA><<<dim3(P.nslow, nz), b, smem, st>>>(
            G, F, time, P.tile_kfast, P.slow_tiles, P.ntx, TEND_KLEV, NTR, P.tm_T, P.tm_S);
        OB_COUNT_LAUNCH(1);
    }
}
void launch_tracers(const Geom& G, const Fields& F, const MomPlan& P, int ord_tracer, double time, cudaStream_t st, int which) {
    if (P.use_tma) launch_tracers_t<true>(G, F, P, ord_tracer, time, st, which);
    else launch_tracers_t<false>(G, F, P, ord_tracer, time, st, which);
}
