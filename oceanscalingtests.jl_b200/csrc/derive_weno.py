#!/usr/bin/env python3
"""Derive the WENO-N (N = 2..5, order 2N-1) tables symbolically and emit C headers.

Build-time generator (its outputs are committed, so a normal build does not need sympy).  The arithmetic that the reference
benchmark runs lives in Oceananigans.jl v0.90.1 (Manifest.toml:643-649), whose
source is NOT available in this container; SURVEY.md Appendix A10 records the
tables from recollection.  This script re-derives every table from first
principles (cell-average reconstruction polynomials on a uniform grid) with
sympy rationals, checks them against the recalled decimal literals, and writes

  oceanscalingtests.jl_b200/csrc/weno_gen.cuh   straight-line device code, constants folded (CUDA kernels)
  oracle/weno_tables.h                          plain tables for the CPU oracle's runtime-order loops

Conventions:
  * left-biased value at face i (between cells i-1 and i); stencil s (0 = most
    downwind) covers cells (i-1-s) .. (i-1-s+N-1); values listed in increasing
    cell index;
  * smoothness beta_s = sum_{a<=b} C[s][ab] psi_a psi_b, upper-triangular
    row-major, scaled by 1 (N=2), 3 (N=3), 0.24 (N=4), 0.0504 (N=5) relative to
    sum_l int (d^l p_s/dx^l)^2 dx over the cell upwind of the face (dx = 1);
  * optimal weights C*_s such that sum_s C*_s p_s = the (2N-1)-order
    upwind-biased reconstruction.
"""
import os
import re
import sys
import sympy as sp

x = sp.Symbol("x")

SCALE = {2: sp.Integer(1), 3: sp.Integer(3), 4: sp.Rational(24, 100), 5: sp.Rational(504, 10000)}


def recon_poly(cells, vals):
    """Polynomial p(x) of degree len(cells)-1 whose cell averages over
    [c, c+1] equal vals (cells are integer left edges)."""
    n = len(cells)
    a = sp.symbols("a0:%d" % n)
    p = sum(a[m] * x**m for m in range(n))
    eqs = [sp.integrate(p, (x, c, c + 1)) - v for c, v in zip(cells, vals)]
    sol = sp.solve(eqs, a, dict=True)[0]
    return sp.expand(p.subs(sol))


def derive(N):
    """Tables for the left-biased reconstruction at the face x = 0
    (cell -1 is [-1,0], cell 0 is [0,1])."""
    coeffs, smooth = [], []
    for s in range(N):
        cells = list(range(-1 - s, -1 - s + N))
        psi = sp.symbols("psi0:%d" % N)
        p = recon_poly(cells, psi)
        face = sp.expand(p.subs(x, 0))
        coeffs.append([sp.Rational(face.coeff(psi[a])) for a in range(N)])
        beta = 0
        for l in range(1, N):
            beta += sp.integrate(sp.diff(p, x, l) ** 2, (x, -1, 0))
        beta = sp.expand(beta * SCALE[N])
        row = []
        for a in range(N):
            for b in range(a, N):
                row.append(sp.Rational(beta.coeff(psi[a], 1).coeff(psi[b], 1)) if a != b
                           else sp.Rational(beta.coeff(psi[a], 2)))
        smooth.append(row)
    # optimal weights: match the full (2N-1)-cell reconstruction
    cells = list(range(-N, N - 1))
    q = sp.symbols("q0:%d" % (2 * N - 1))
    full = sp.expand(recon_poly(cells, q).subs(x, 0))
    g = sp.symbols("g0:%d" % N)
    combo = 0
    for s in range(N):
        for a in range(N):
            cell = -1 - s + a
            combo += g[s] * coeffs[s][a] * q[cell + N]
    eqs = [sp.expand(combo - full).coeff(q[m]) for m in range(2 * N - 1)]
    sol = sp.solve(eqs, g, dict=True)[0]
    opt = [sp.Rational(sol[g[s]]) for s in range(N)]
    return coeffs, opt, smooth


# decimal literals recalled in SURVEY.md A10 (the scale factors rest on these)
RECALLED_SMOOTH = {
    2: [[1, -2, 1], [1, -2, 1]],
    3: [[10, -31, 11, 25, -19, 4], [4, -13, 5, 13, -13, 4], [4, -19, 11, 25, -31, 10]],
    4: [[2.107, -9.402, 7.042, -1.854, 11.003, -17.246, 4.642, 7.043, -3.882, 0.547],
        [0.547, -2.522, 1.922, -0.494, 3.443, -5.966, 1.602, 2.843, -1.642, 0.267],
        [0.267, -1.642, 1.602, -0.494, 2.843, -5.966, 1.922, 3.443, -2.522, 0.547],
        [0.547, -3.882, 4.642, -1.854, 7.043, -17.246, 7.042, 11.003, -9.402, 2.107]],
    5: [[1.07918, -6.49501, 7.58823, -4.11487, 0.86329, 10.20563, -24.62076, 13.58458, -2.88007,
         15.21393, -17.04396, 3.64863, 4.82963, -2.08501, 0.22658],
        [0.22658, -1.40251, 1.65153, -0.88297, 0.18079, 2.42723, -6.11976, 3.37018, -0.70237,
         4.06293, -4.64976, 0.99213, 1.38563, -0.60871, 0.06908],
        [0.06908, -0.51001, 0.67923, -0.38947, 0.08209, 1.04963, -2.99076, 1.79098, -0.38947,
         2.31153, -2.99076, 0.67923, 1.04963, -0.51001, 0.06908],
        [0.06908, -0.60871, 0.99213, -0.70237, 0.18079, 1.38563, -4.64976, 3.37018, -0.88297,
         4.06293, -6.11976, 1.65153, 2.42723, -1.40251, 0.22658],
        [0.22658, -2.08501, 3.64863, -2.88007, 0.86329, 4.82963, -17.04396, 13.58458, -4.11487,
         15.21393, -24.62076, 7.58823, 10.20563, -6.49501, 1.07918]],
}
RECALLED_OPT = {
    2: [sp.Rational(2, 3), sp.Rational(1, 3)],
    3: [sp.Rational(3, 10), sp.Rational(3, 5), sp.Rational(1, 10)],
    4: [sp.Rational(4, 35), sp.Rational(18, 35), sp.Rational(12, 35), sp.Rational(1, 35)],
    5: [sp.Rational(5, 126), sp.Rational(20, 63), sp.Rational(10, 21), sp.Rational(10, 63), sp.Rational(1, 126)],
}


def lit(r):
    return repr(float(r))


def emit(tables, path, cuda):
    q = "__device__ __constant__ const" if cuda else "static const"
    guard = "OB200_WENO_TABLES_%s" % ("CUH" if cuda else "H")
    out = ["// GENERATED by oracle/derive_weno.py -- do not edit.",
           "// WENO-N tables, N = 2..5 (index [N][s][...], rows for N < 2 unused).",
           "// `real` is the includer's arithmetic type (double, or float with -DOB200_F32); the literals are Float64 and",
           "// are rounded once, exactly like RL(x) in the device code.",
           "#ifndef %s" % guard, "#define %s" % guard, ""]
    def arr3(name, key, width):
        out.append("%s real %s[6][5][%d] = {" % (q, name, width))
        for N in range(6):
            out.append("  {")
            for s in range(5):
                vals = ["0.0"] * width
                if N in tables and s < N:
                    src = tables[N][key][s]
                    for m, v in enumerate(src):
                        vals[m] = lit(v)
                out.append("    {" + ", ".join(vals) + "},")
            out.append("  },")
        out.append("};")
    arr3("WENO_COEFF", 0, 5)
    out.append("%s real WENO_OPT[6][5] = {" % q)
    for N in range(6):
        vals = ["0.0"] * 5
        if N in tables:
            for s in range(N):
                vals[s] = lit(tables[N][1][s])
        out.append("  {" + ", ".join(vals) + "},")
    out.append("};")
    arr3("WENO_SMOOTH", 2, 15)
    out += ["", "#endif", ""]
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        f.write("\n".join(out))


def emit_device_code(tables, path):
    """Straight-line __device__ code per order N: q[0..2N-2] are the stencil values ordered from the
    most upwind-ward cell (q[0]) to the most downwind (q[2N-2]); stencil s uses q[N-1-s .. 2N-2-s].
    Weno<N>::beta(q, b) fills the N smoothness indicators; Weno<N>::combine(q, b) applies Z-weights.
    Every multiply-add is an EXPLICIT fma() in a fixed order, identical to the oracle's loops: the quadratic
    forms cancel catastrophically on fields with a large mean (S ~ 35), so beta is only reproducible to
    ~1e-8 relative between two evaluation orders; pinning the order keeps CUDA and oracle bit-compatible
    while the library is compiled with -fmad=false (no implicit contraction anywhere else)."""
    out = ["// GENERATED by csrc/derive_weno.py -- do not edit.",
           "#pragma once", "",
           "template <int N> struct Weno;", ""]
    taus = {2: "fabs(b[0] - b[1])", 3: "fabs(b[0] - b[2])",
            4: "fabs(b[0] + 3.0 * b[1] - 3.0 * b[2] - b[3])",
            5: "fabs(b[0] + 2.0 * b[1] - 6.0 * b[2] + 2.0 * b[3] + b[4])"}
    # coefficients live in __constant__ arrays indexed with compile-time constants, so that DFMA/DMUL take
    # them as constant-bank operands (64-bit literals would be materialised with two UMOVs each on sm_100a:
    # measured 14-20 % of all issued instructions in the first tiled kernels)
    for N in (2, 3, 4, 5):
        c, o, b = tables[N]
        nb = N * (N + 1) // 2
        out.append("static __device__ __constant__ double WB%d[%d] = {%s};" % (
            N, N * nb, ", ".join(lit(b[s][m]) for s in range(N) for m in range(nb))))
        out.append("static __device__ __constant__ double WP%d[%d] = {%s};" % (
            N, N * N, ", ".join(lit(c[s][a]) for s in range(N) for a in range(N))))
        out.append("static __device__ __constant__ double WO%d[%d] = {%s};" % (N, N, ", ".join(lit(o[s]) for s in range(N))))
    out.append("")
    for N in (2, 3, 4, 5):
        c, o, b = tables[N]
        nb = N * (N + 1) // 2
        out.append("template <> struct Weno<%d> {" % N)
        out.append("  static constexpr int NQ = %d;" % (2 * N - 1))
        out.append("  __device__ __forceinline__ static void beta(const double* __restrict__ q, double* __restrict__ b) {")
        for s in range(N):
            off = N - 1 - s
            m = 0
            acc = None
            for a in range(N):
                inner = None
                for bb in range(a, N):
                    cst = "WB%d[%d]" % (N, s * nb + m)
                    term = "%s * q[%d]" % (cst, off + bb) if inner is None else \
                        "fma(%s, q[%d], %s)" % (cst, off + bb, inner)
                    inner = term
                    m += 1
                acc = "q[%d] * (%s)" % (off + a, inner) if acc is None else "fma(q[%d], %s, %s)" % (off + a, inner, acc)
            out.append("    b[%d] = %s;" % (s, acc))
        out.append("  }")
        # two fields at once (VelocityStencil: u and v): statement pairs share each coefficient load
        out.append("  __device__ __forceinline__ static void beta2(const double* __restrict__ q, const double* __restrict__ p,")
        out.append("                                                double* __restrict__ b, double* __restrict__ c) {")
        for s in range(N):
            off = N - 1 - s
            for (src, dst) in (("q", "b"), ("p", "c")):
                m = 0
                acc = None
                for a in range(N):
                    inner = None
                    for bb in range(a, N):
                        cst = "WB%d[%d]" % (N, s * nb + m)
                        term = "%s * %s[%d]" % (cst, src, off + bb) if inner is None else \
                            "fma(%s, %s[%d], %s)" % (cst, src, off + bb, inner)
                        inner = term
                        m += 1
                    acc = "%s[%d] * (%s)" % (src, off + a, inner) if acc is None else "fma(%s[%d], %s, %s)" % (src, off + a, inner, acc)
                out.append("    %s[%d] = %s;" % (dst, s, acc))
        out.append("  }")
        out.append("  __device__ __forceinline__ static double combine(const double* __restrict__ q, const double* __restrict__ b) {")
        out.append("    const double tau = %s;" % taus[N])
        # Float32 build: the reference's own form alpha_s = C_s (1 + (tau / (beta_s + eps))^2) -- the product of all
        # (beta_s + eps) of the division-free form below underflows in Float32 ((1e-8)^N squared).  beta_s >= 0
        # mathematically, but the expanded quadratic form rounds to values of either sign when the field is smooth, and in
        # Float32 beta_s + eps is then EXACTLY zero about once per step at 3e7 cells (measured: isolated NaN tendencies after one
        # step of a 270 x 900 x 120 grid).  The denominator is therefore kept at least 1e-12 away from zero, sign preserved
        # (clamping beta_s at zero instead hands a stencil whose beta_s is rounding noise a weight 1e10 times the others':
        # measured unstable within ten steps).
        polys = []
        for s_ in range(N):
            off_ = N - 1 - s_
            poly_ = None
            for a_ in range(N):
                cst_ = "WP%d[%d]" % (N, s_ * N + a_)
                poly_ = "%s * q[%d]" % (cst_, off_ + a_) if poly_ is None else "fma(%s, q[%d], %s)" % (cst_, off_ + a_, poly_)
            polys.append(poly_)
        out.append("#ifdef OB200_F32")
        out.append("    double asum = 0.0, r = 0.0;")
        for s_ in range(N):
            out.append("    { const double d = b[%d] + 1e-8; const double t = tau / copysign(fmax(fabs(d), 1e-12), d);" % s_)
            out.append("      const double al = WO%d[%d] * fma(t, t, 1.0);" % (N, s_))
            out.append("      asum += al; r = fma(al, %s, r); }" % polys[s_])
        out.append("    return r / asum;")
        out.append("#else")
        # division-free Z-weights: alpha_s * P^2 = C_s * (P^2 + (tau * P_s)^2), P_s = prod_{r != s} d_r, d = beta + eps
        out.append("    double d[%d], L[%d], R[%d];" % (N, N, N))
        out.append("    " + " ".join("d[%d] = b[%d] + 1e-8;" % (s, s) for s in range(N)))
        out.append("    L[0] = 1.0; " + " ".join("L[%d] = L[%d] * d[%d];" % (s + 1, s, s) for s in range(N - 1)))
        out.append("    R[%d] = 1.0; " % (N - 1) + " ".join("R[%d] = R[%d] * d[%d];" % (s - 1, s, s) for s in range(N - 1, 0, -1)))
        out.append("    const double Pt = L[%d] * d[%d]; const double PP = Pt * Pt;" % (N - 1, N - 1))
        out.append("    double asum = 0.0, r = 0.0;")
        for s in range(N):
            off = N - 1 - s
            poly = None
            for a in range(N):
                cst = "WP%d[%d]" % (N, s * N + a)
                poly = "%s * q[%d]" % (cst, off + a) if poly is None else \
                    "fma(%s, q[%d], %s)" % (cst, off + a, poly)
            out.append("    { const double t = tau * (L[%d] * R[%d]); const double al = WO%d[%d] * fma(t, t, PP);" % (s, s, N, s))
            out.append("      asum += al; r = fma(al, %s, r); }" % poly)
        out.append("    return r / asum;")
        out.append("#endif")
        out.append("  }")
        out.append("};")
        out.append("")
    out.append("template <> struct Weno<1> {")
    out.append("  static constexpr int NQ = 1;")
    out.append("  __device__ __forceinline__ static void beta(const double*, double*) {}")
    out.append("  __device__ __forceinline__ static void beta2(const double*, const double*, double*, double*) {}")
    out.append("  __device__ __forceinline__ static double combine(const double* q, const double*) { return q[0]; }")
    out.append("};")
    out.append("")
    with open(path, "w") as f:
        f.write(to_real("\n".join(out)))


_LIT = re.compile(r'(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])')


def to_real(text):
    """The device code is written against the library's `real` typedef (Float64, or Float32 with -DOB200_F32): `double`
    becomes `real` and every floating literal x becomes RL(x) = ((real)(x)) -- same rule as tools/to_real.py."""
    lines = []
    for line in text.split("\n"):
        i = line.find("//")
        code, comment = (line, "") if i < 0 else (line[:i], line[i:])
        code = re.sub(r"\bdouble\b", "real", code)
        code = _LIT.sub(lambda m: "RL(%s)" % m.group(1), code)
        lines.append(code + comment)
    return "\n".join(lines)


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(os.path.dirname(here))
    tables = {}
    for N in (2, 3, 4, 5):
        c, o, b = derive(N)
        tables[N] = (c, o, b)
        assert o == RECALLED_OPT[N], (N, o)
        for s in range(N):
            for got, want in zip(b[s], RECALLED_SMOOTH[N][s]):
                assert abs(float(got) - want) < 1e-9, (N, s, float(got), want)
            assert sum(c[s]) == 1
    # spot checks of the classic coefficients
    assert tables[3][0][0] == [sp.Rational(1, 3), sp.Rational(5, 6), sp.Rational(-1, 6)]
    assert tables[3][0][2] == [sp.Rational(1, 3), sp.Rational(-7, 6), sp.Rational(11, 6)]
    emit(tables, os.path.join(root, "oracle", "weno_tables.h"), cuda=False)
    emit_device_code(tables, os.path.join(here, "weno_gen.cuh"))
    print("weno tables derived and verified against SURVEY A10 literals")


if __name__ == "__main__":
    sys.exit(main())
