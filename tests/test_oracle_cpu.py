"""CPU tests of the oracle (test infrastructure): WENO machinery against independent numpy restatements,
physical invariants that hold regardless of recollection (SURVEY 7 "hard parts"), and the committed golden
vectors.  The reference's own tests pin no numbers (test/runtests.jl:33-34): parity with Oceananigans is UNPINNED."""
import os

import numpy as np
import pytest

import cases

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---- WENO ---------------------------------------------------------------------------------------------
def numpy_weno_js(N, q):
    """Independent restatement: Z-weighted WENO from Lagrange-type cell-average polynomials fitted with numpy."""
    q = np.asarray(q, dtype=np.float64)
    optimal = {2: [2 / 3, 1 / 3], 3: [3 / 10, 3 / 5, 1 / 10], 4: [4 / 35, 18 / 35, 12 / 35, 1 / 35],
               5: [5 / 126, 20 / 63, 10 / 21, 10 / 63, 1 / 126]}[N]
    scale = {2: 1.0, 3: 3.0, 4: 0.24, 5: 0.0504}[N]
    from numpy.polynomial import polynomial as P
    cand, beta = [], []
    for s in range(N):
        cells = np.arange(-1 - s, -1 - s + N)          # left edges; face at x = 0
        vals = q[N - 1 - s: 2 * N - 1 - s]
        A = np.array([[((c + 1) ** (m + 1) - c ** (m + 1)) / (m + 1) for m in range(N)] for c in cells])
        a = np.linalg.solve(A, vals)                   # polynomial with the given cell averages
        cand.append(a[0])                              # value at x = 0
        b = 0.0
        for l in range(1, N):
            d = P.polyder(a, l)
            d2 = P.polymul(d, d)
            I = P.polyint(d2)
            b += P.polyval(0.0, I) - P.polyval(-1.0, I)
        beta.append(scale * b)
    beta = np.array(beta)
    tau_w = {2: [1, -1], 3: [1, 0, -1], 4: [1, 3, -3, -1], 5: [1, 2, -6, 2, 1]}[N]
    tau = abs(float(np.dot(tau_w, beta)))
    alpha = np.array(optimal) * (1 + (tau / (beta + 1e-8)) ** 2)
    return float(np.sum(alpha / alpha.sum() * np.array(cand)))


@pytest.mark.parametrize("N", [2, 3, 4, 5])
def test_weno_matches_independent_numpy_restatement(orc, N):
    for trial, f in enumerate([np.sin, np.exp, lambda x: np.tanh(5 * x), lambda x: np.where(x > 0.1, 1.0, 0.0)]):
        x = (np.arange(2 * N - 1) - N + 0.5) * 0.3 + 0.05 * trial
        q = f(x)
        got = orc.weno(N, q)
        want = numpy_weno_js(N, q)
        assert abs(got - want) <= 5e-12 * max(1.0, abs(want)), (N, trial, got, want)


@pytest.mark.parametrize("N", [1, 2, 3, 4, 5])
def test_weno_exact_on_constants_and_high_order_on_smooth_data(orc, N):
    q = np.full(2 * N - 1, 35.0)
    assert abs(orc.weno(N, q) - 35.0) < 1e-13
    if N == 1:
        return
    # cell averages of a degree-(N-1) polynomial are reconstructed exactly by every candidate stencil
    rng = np.random.default_rng(20240611)
    coef = rng.standard_normal(N)
    from numpy.polynomial import polynomial as P
    I = P.polyint(coef)
    edges = np.arange(-N, N)
    q = np.array([P.polyval(e + 1, I) - P.polyval(e, I) for e in edges[:2 * N - 1]])
    assert abs(orc.weno(N, q) - P.polyval(0.0, coef)) < 1e-11
    # velocity-stencil smoothness: with smooth u, v stencils the weights tend to the optimal ones (order 2N-1)
    h = 0.05
    xs = (np.arange(2 * N - 1) - N) * h
    Iq = lambda a, b: (np.cos(a) - np.cos(b)) / (b - a)       # cell average of sin
    q = np.array([Iq(a, a + h) for a in xs])
    got = orc.weno(N, q, sa=np.cos(xs), sb=np.sin(2 * xs))
    assert abs(got - np.sin(0.0)) < 10 * h ** (2 * N - 1)


def test_deterministic_tanh_and_teos10_known_values(orc):
    lib = orc.load()
    xs = np.concatenate([np.linspace(-25, 25, 2001), [1e-9, -1e-9, 0.0, 0.25, -0.25]])
    err = max(abs(lib.oracle_det_tanh(float(x)) - np.tanh(x)) for x in xs)
    assert err < 3e-16
    # TEOS-10 polynomial: fresh water at 0 C, 0 dbar -> 999.843 kg/m3 (known value); sea water is denser
    assert abs(lib.oracle_teos10_rprime(0.0, 0.0, 0.0) - 999.843) < 1e-3
    # published known-answer value of the 55-term polynomial of Roquet et al. (2015) (polyTEOS10-bsq), quoted in NEMO's
    # eosbn2.F90: "rho = 1028.21993233072 kg/m^3 for z = 3000 dbar, ct = 3 Celsius, sa = 35.5 g/kg" -- all 14 digits, which
    # pins every one of the 52 non-zero coefficients recalled in oracle/ocean_oracle.cpp and csrc/ob_common.cuh
    assert abs(lib.oracle_teos10_rprime(3.0, 35.5, -3000.0) - 1028.21993233072) < 5e-11
    # the analytic derivatives N^2 is built from (alpha, beta at faces) against central differences of the polynomial
    import ctypes as C
    dT, dS = C.c_double(), C.c_double()
    for Th, SA, Z in [(10.0, 35.16504, 0.0), (3.0, 35.5, -3000.0), (25.0, 34.0, -100.0), (-1.5, 33.0, -5000.0)]:
        r = lib.oracle_teos10_rprime_derivatives(Th, SA, Z, C.byref(dT), C.byref(dS))
        assert r == lib.oracle_teos10_rprime(Th, SA, Z)
        h = 1e-4
        fdT = (lib.oracle_teos10_rprime(Th + h, SA, Z) - lib.oracle_teos10_rprime(Th - h, SA, Z)) / (2 * h)
        fdS = (lib.oracle_teos10_rprime(Th, SA + h, Z) - lib.oracle_teos10_rprime(Th, SA - h, Z)) / (2 * h)
        assert abs(dT.value - fdT) < 2e-8 * abs(fdT) + 1e-9 and abs(dS.value - fdS) < 2e-8 * abs(fdS), (Th, SA, Z)
        assert dS.value > 0 and (dT.value < 0 or Th < 0)      # saltier is denser; warmer is lighter (above ~0 C at depth)
    assert 1026.0 < lib.oracle_teos10_rprime(10.0, 35.0, 0.0) < 1028.0
    assert lib.oracle_teos10_rprime(10.0, 35.0, 0.0) > lib.oracle_teos10_rprime(20.0, 35.0, 0.0)


# ---- physical invariants -----------------------------------------------------------------------------------
def test_quiescent_stays_at_rest(pkg, orc):
    c = cases.build_case("quiescent")
    m = c.make_model(pkg, orc.OracleBackend)
    for _ in range(5):
        pkg.time_step(m, c.dt)
    for n in ("U", "V", "W", "ETA", "T", "S"):
        assert np.max(np.abs(m.get(n))) == 0.0


def test_stratified_rest_state_has_no_spurious_flow(pkg, orc):
    """Horizontally uniform T(z), S on the DoubleDrake bathymetry with the surface forcing switched off:
    pressure gradients vanish identically, so the ocean stays at rest (vertical diffusion only)."""
    c = cases.Case("rest", 90, 40, 8, experiment="DoubleDrake", perturb=False)
    grid = c.grid(pkg)
    fs = pkg.SplitExplicitFreeSurface(substeps=30)
    m = pkg.HydrostaticFreeSurfaceModel(grid=grid, free_surface=fs, eos="linear",
                                        boundary_conditions=pkg.set_boundary_conditions("Quiescent", grid),
                                        backend_factory=orc.OracleBackend)
    pkg.set_model(m, T=lambda lam, phi, z: 20.0 * np.exp(z / 400.0) + 0 * lam + 0 * phi, S=35.0)
    for _ in range(10):
        pkg.time_step(m, 1200.0)
    assert np.max(np.abs(m.get("U"))) < 1e-14 and np.max(np.abs(m.get("V"))) < 1e-14
    assert np.max(np.abs(m.get("ETA"))) < 1e-12


def volume_integral(pkg, model, name):
    g = model.grid
    dz = np.diff(g.z_faces)[:, None, None]
    phif = g.phi_f()
    az = (np.sin(np.pi * phif[1:] / 180) - np.sin(np.pi * phif[:-1] / 180))[None, :, None]
    return float(np.sum(model.get(name) * dz * az))


def test_tracer_content_changes_only_through_surface_fluxes(pkg, orc):
    """Flux-form advection + no-flux walls: with zero surface fluxes the volume integral of S is conserved up
    to the free-surface term (w at the top face carries tracer through the linearised surface)."""
    c = cases.Case("cons", 90, 40, 8, experiment="DoubleDrake", perturb=True)
    grid = c.grid(pkg)
    fs = pkg.SplitExplicitFreeSurface(substeps=30)
    m = pkg.HydrostaticFreeSurfaceModel(grid=grid, free_surface=fs, eos="linear",
                                        boundary_conditions=pkg.set_boundary_conditions("Quiescent", grid),
                                        backend_factory=orc.OracleBackend)
    c.initialize(pkg, m) if False else None
    r = np.pi / 180
    pkg.set_model(m, T=lambda lam, phi, z: 10.0 + 5 * np.exp(z / 500.0) * np.cos(phi * r) + 0 * lam,
                  S=lambda lam, phi, z: 35.0 + 0.3 * np.sin(2 * lam * r) * np.cos(phi * r) * np.exp(z / 800.0),
                  u=lambda lam, phi, z: 0.1 * np.sin(2 * lam * r) * np.cos(phi * r) + 0 * z)
    s0 = volume_integral(pkg, m, "S")
    for _ in range(10):
        pkg.time_step(m, 600.0)
    s1 = volume_integral(pkg, m, "S")
    # surface term: int S w_top dt dA / int S dV ~ eta / depth ~ 1e-4; everything else telescopes exactly
    assert abs(s1 - s0) / abs(s0) < 5e-5
    # uniform salinity is preserved to rounding whatever the flow (constancy preservation of div(U c))
    pkg.set_model(m, S=35.0)
    for _ in range(5):
        pkg.time_step(m, 600.0)
    S = m.get("S")
    wet = S != 0
    assert np.max(np.abs(S[wet] - 35.0)) < 2e-4   # w_top * S / dz free-surface term only


def test_zonally_symmetric_setup_stays_zonally_symmetric(pkg, orc):
    """No bathymetry + zonally symmetric forcing and state => every field independent of longitude."""
    c = cases.Case("zonal", 48, 40, 8, experiment="Quiescent", perturb=False)
    grid = c.grid(pkg)
    fs = pkg.SplitExplicitFreeSurface(substeps=30)
    bcs = pkg.set_boundary_conditions("DoubleDrake", grid)
    m = pkg.HydrostaticFreeSurfaceModel(grid=grid, free_surface=fs, eos="linear", boundary_conditions=bcs,
                                        backend_factory=orc.OracleBackend)
    pkg.initialize_model(m, "DoubleDrake")
    for _ in range(20):
        pkg.time_step(m, 1200.0)
    for n in ("U", "V", "T", "S", "ETA"):
        a = m.get(n)
        assert np.max(np.abs(a - a[..., :1])) <= 1e-13 * max(1.0, np.max(np.abs(a))), n
    assert np.max(np.abs(m.get("U"))) > 1e-4      # the wind does spin the ocean up


def test_masked_nodes_stay_zero(pkg, orc):
    c = cases.build_case("bumpy")
    m = c.make_model(pkg, orc.OracleBackend)
    for _ in range(5):
        pkg.time_step(m, c.dt)
    imm = m.grid.immersed_cell_mask()
    assert np.all(m.get("T")[imm] == 0) and np.all(m.get("S")[imm] == 0)
    u = m.get("U")
    assert np.all(u[imm] == 0) and np.all(u[np.roll(imm, 1, axis=2)] == 0)
    assert np.all(np.isfinite(u)) and np.max(np.abs(u)) < 5.0


# ---- golden vectors (generated by tests/golden/make_golden.py from this oracle; regression pins) ---------------
@pytest.mark.parametrize("name", ["ref_test", "dd_small", "bumpy"])
def test_oracle_reproduces_golden_vectors(pkg, orc, name):
    path = os.path.join(GOLDEN, f"{name}.npz")
    gold = np.load(path)
    c = cases.build_case(name)
    m = c.make_model(pkg, orc.OracleBackend)
    for _ in range(int(gold["steps"])):
        pkg.time_step(m, c.dt)
    bad = cases.check_signature({n: m.get(n) for n in ("U", "V", "W", "ETA", "T", "S")}, gold, 1e-11)
    assert not bad, bad


def test_implicit_vertical_diffusion_solves_the_backward_euler_system(pkg, orc):
    """AB2 + implicit step of T against an independent restatement: T* = T + dt G^n (first step: forward Euler), then
    (I - dt d/dz kappa d/dz) T+ = T* per column as a banded system solved by LAPACK (scipy), with kappa = kappa_z +
    kappa_c on interior faces and zero on faces that touch the sea floor or a solid cell."""
    from scipy.linalg import solve_banded
    c = cases.build_case("bumpy")
    m = c.make_model(pkg, orc.OracleBackend)
    g = m.grid
    T0, GT, kc = m.get("T"), m.get("GT"), m.get("KAPPA_C")
    dt = c.dt
    m.backend.run_stage(pkg._abi.STAGE_ID["AB2_IMPLICIT"], dt, 0)
    T1 = m.get("T")
    zf, zc = g.z_faces, g.z_c()
    dzc = zf[1:] - zf[:-1]
    dzf = np.empty(g.Nz + 1)
    dzf[1:-1] = zc[1:] - zc[:-1]
    dzf[0] = dzf[-1] = np.nan                       # boundary faces carry no flux
    solid = g.immersed_cell_mask()                   # (Nz, Ny, Nx)
    kappa_z = 3e-5                                   # near_global_simulation.jl:71
    worst = 0.0
    rng = [(j, i) for j in range(0, g.Ny, 5) for i in range(0, g.Nx, 7)]
    for j, i in rng:
        wet = ~solid[:, j, i]
        if not wet.any():
            continue
        rhs = T0[:, j, i] + dt * GT[:, j, i]
        K = np.zeros(g.Nz + 1)
        for kf in range(1, g.Nz):
            if wet[kf - 1] and wet[kf]:
                K[kf] = kappa_z + kc[kf, j, i]
        ab = np.zeros((3, g.Nz))
        for k in range(g.Nz):
            up = -dt * K[k + 1] / (dzc[k] * dzf[k + 1]) if k < g.Nz - 1 else 0.0
            lo = -dt * K[k] / (dzc[k] * dzf[k]) if k > 0 else 0.0
            ab[1, k] = 1.0 - up - lo
            if k < g.Nz - 1:
                ab[0, k + 1] = up
            if k > 0:
                ab[2, k - 1] = lo
        sol = solve_banded((1, 1), ab, rhs)
        scale = max(np.max(np.abs(sol)), 1e-30)
        worst = max(worst, float(np.max(np.abs(sol - T1[:, j, i])) / scale))
    assert worst < 1e-12, worst


def test_w_satisfies_discrete_continuity(pkg, orc):
    """_compute_w_from_continuity!: w(k+1) = w(k) - dz_k * div_xy(u, v)_k with w = 0 at the bottom,
    restated with numpy on the interior (periodic in x, v = 0 on the walls)."""
    c = cases.build_case("dd_small")
    m = c.make_model(pkg, orc.OracleBackend)
    g = m.grid
    u, v, w = m.get("U"), m.get("V"), m.get("W")
    rad = np.pi / 180.0
    dy = g.radius * g.dphi * rad
    dxcf = g.radius * np.cos(g.phi_f() * rad) * g.dlam * rad                       # dx at (c, f)
    az = g.radius ** 2 * g.dlam * rad * (np.sin(g.phi_f()[1:] * rad) - np.sin(g.phi_f()[:-1] * rad))   # Az at (c, c)
    dz = g.z_faces[1:] - g.z_faces[:-1]
    div = (dy * (np.roll(u, -1, axis=2) - u) + dxcf[None, 1:, None] * v[:, 1:, :] - dxcf[None, :-1, None] * v[:, :-1, :]) / az[None, :, None]
    wref = np.zeros_like(w)
    for k in range(g.Nz):
        wref[k + 1] = wref[k] - dz[k] * div[k]
    scale = np.max(np.abs(wref))
    assert scale > 0 and np.max(np.abs(w - wref)) / scale < 1e-12


def test_hydrostatic_pressure_is_the_vertical_integral_of_buoyancy(pkg, orc):
    """_update_hydrostatic_pressure! with the linear equation of state b = g (alpha T - beta S): top-down
    trapezoidal integral, restated with numpy.  The top half-cell uses b(Nz+1) = b(Nz) (halo = top value)."""
    c = cases.build_case("dd_small")
    m = c.make_model(pkg, orc.OracleBackend)
    g = m.grid
    T, S, p = m.get("T"), m.get("S"), m.get("P")
    b = m.cfg.gravity * (m.cfg.eos_alpha * T - m.cfg.eos_beta * S)
    zc = g.z_c()
    dzf = np.empty(g.Nz + 1)
    dzf[1:-1] = zc[1:] - zc[:-1]
    dzf[-1] = 2.0 * (g.z_faces[-1] - zc[-1])      # centre-to-centre spacing across the surface (reflected top cell)
    pref = np.zeros_like(p)
    pref[-1] = -0.5 * (b[-1] + b[-1]) * dzf[-1]
    for k in range(g.Nz - 2, -1, -1):
        pref[k] = pref[k + 1] - 0.5 * (b[k + 1] + b[k]) * dzf[k + 1]
    wet = ~g.immersed_cell_mask()
    scale = np.max(np.abs(pref[wet]))
    assert scale > 0 and np.max(np.abs((p - pref)[wet])) / scale < 1e-12


def test_free_surface_volume_is_conserved(pkg, orc):
    """The split-explicit substeps are in flux form on a closed (walls N/S, periodic E/W) domain: sum(eta * Az) over
    wet columns does not change, whatever the flow."""
    c = cases.build_case("bumpy")
    m = c.make_model(pkg, orc.OracleBackend)
    g = m.grid
    rad = np.pi / 180.0
    az = g.radius ** 2 * g.dlam * rad * (np.sin(g.phi_f()[1:] * rad) - np.sin(g.phi_f()[:-1] * rad))
    vol0 = float(np.sum(m.get("ETA") * az[:, None]))
    ref = float(np.sum(np.abs(m.get("ETA")) * az[:, None]))
    for _ in range(5):
        pkg.time_step(m, c.dt)
    vol1 = float(np.sum(m.get("ETA") * az[:, None]))
    assert abs(vol1 - vol0) / ref < 1e-12


# ---- the two arithmetic modes of the oracle (oracle/ocean_oracle.cpp header) ----------------------------------------
def test_reference_arithmetic_weno_equals_product_arithmetic_to_rounding(orc):
    """alpha_s = C_s (1 + (tau / (beta_s + eps))^2), w = alpha / sum(alpha) (reference) against the division-free
    fma form (product): the same numbers up to rounding, on smooth and on kinked stencils, with and without a large mean."""
    import ctypes as C
    lib = orc.load()
    lib.oracle_weno_ref.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.oracle_weno_ref.restype = C.c_double
    rng = np.random.default_rng(20240611)
    worst = 0.0
    for N in (2, 3, 4, 5):
        for mean in (0.0, 35.0):
            for kind in ("smooth", "kink"):
                for _ in range(20):
                    x = np.arange(2 * N - 1, dtype=np.float64)
                    q = mean + (np.sin(0.3 * x + rng.uniform(0, 6)) if kind == "smooth" else np.where(x < N - 1, 0.0, 1.0) + 0.01 * rng.standard_normal(2 * N - 1))
                    q = np.ascontiguousarray(q)
                    a = lib.oracle_weno(N, q.ctypes.data, q.ctypes.data, None)
                    b = lib.oracle_weno_ref(N, q.ctypes.data, q.ctypes.data, None)
                    worst = max(worst, abs(a - b) / max(abs(b), 1.0))
    # beta of a field with mean 35 cancels ~ 35^2 / (range^2): the weights are only reproducible to ~1e-8 relative between
    # two evaluation orders, the reconstruction (a convex combination of nearby values) to ~1e-10 of the range
    assert worst < 1e-9, worst


def test_reference_arithmetic_mode_is_the_same_algorithm(pkg, orc):
    """One step in both modes: every field agrees to rounding (1e-13); 20 steps: still within 1e-11 for u, v, eta, T, S."""
    c = cases.build_case("bumpy")
    ma = c.make_model(pkg, orc.OracleBackend)
    mb = c.make_model(pkg, orc.OracleBackend, initialize=False)
    mb.backend.set_option("reference_arithmetic", 1)
    c.initialize(pkg, mb)
    for n, tol in ((1, 1e-13), (20, 1e-11)):
        while ma.clock["iteration"] < n:
            pkg.time_step(ma, c.dt)
            pkg.time_step(mb, c.dt)
        for f in ("U", "V", "ETA", "T", "S"):
            a, b = ma.get(f), mb.get(f)
            assert np.max(np.abs(a - b)) <= tol * np.max(np.abs(b)), (n, f)
        assert not np.array_equal(ma.get("U"), mb.get("U")) or n == 0      # the modes do differ in the last bits


def test_teos10_fixed_TS_column_has_zero_N2_oracle(pkg, orc):
    """N^2 = g (alpha dT/dz - beta dS/dz), alpha / beta at the face (Oceananigans' ∂z_b): zero for a fixed T, S column even
    though the in-situ buoyancy varies with depth (TEOS-10 compressibility) -> kappa takes its Ri = 0 value."""
    c = cases.build_case("teos_uniform")
    m = c.make_model(pkg, orc.OracleBackend)
    kc = m.get("KAPPA_C")
    ri = pkg.RiBasedVerticalDiffusivity()
    tau0 = 0.5 * (1.0 - np.tanh((0.0 - ri.Ri0) / ri.Ridelta))
    f = 1.0 / (1.0 + ri.Cav)
    wet = kc[1:-1][kc[1:-1] > 0]
    assert wet.size > 0.5 * kc[1:-1].size
    assert np.allclose(wet, ri.kappa0 * tau0 * (f + ri.Cav * f * f), rtol=1e-12)
    # in-situ buoyancy DOES vary with depth for this column (what the old finite difference picked up)
    b_top = orc.load().oracle_teos10_rprime(10.0, 35.0, -5.0)
    b_deep = orc.load().oracle_teos10_rprime(10.0, 35.0, -4000.0)
    assert abs(b_deep - b_top) > 0.1


def test_compare_dump_tool_on_a_self_made_dump(pkg, orc, tmp_path):
    """tests/compare_dump.py (the consumer of julia/compare_with_oceananigans.jl) on a dump written by the oracle itself
    in the layout the Julia script writes: the reference-arithmetic row must read 0, the product-arithmetic row a few ulps."""
    import compare_dump
    steps = 5
    fields = compare_dump.run_case(pkg, orc.OracleBackend, steps, 600.0, reference_arithmetic=True)
    path = tmp_path / "dump.npz"
    np.savez(path, steps=steps, dt=600.0, **fields)
    assert compare_dump.main([str(path)]) == 0
    prod = compare_dump.run_case(pkg, orc.OracleBackend, steps, 600.0)
    d = compare_dump.distances(prod, np.load(path))
    assert 0.0 < max(d.values()) < 1e-11


# ---- Float32 build of the oracle (the checker of libocean_b200_f32.so) --------------------------------------------------
F32_TOL = {"T": 3e-3, "S": 6e-4, "U": 5e-2, "V": 2.5e-2, "ETA": 2e-2, "W": 2e-1}   # as in tests/test_parity_gpu.py


@pytest.mark.parametrize("name", ["dd_small", "realistic"])
def test_float32_oracle_tracks_the_float64_oracle(pkg, orc, name):
    """libocean_oracle_f32.so is ocean_oracle.cpp compiled with real = float.  It starts from the Float64 initial state
    rounded once, returns Float32 values exactly (its C interface is Float64), and after 10 steps sits inside the
    per-field bars the GPU test holds the Float32 library to."""
    c = cases.build_case(name)
    m32, m64 = c.make_model(pkg, orc.OracleBackendF32), c.make_model(pkg, orc.OracleBackend)
    for n in ("U", "V", "ETA", "T", "S"):
        a = m32.get(n)
        assert np.array_equal(a, a.astype(np.float32).astype(np.float64)), n      # Float32 values, held exactly
        assert np.array_equal(a, m64.get(n).astype(np.float32).astype(np.float64)), n
    for _ in range(10):
        pkg.time_step(m32, c.dt)
        pkg.time_step(m64, c.dt)
    for n, tol in F32_TOL.items():
        a, b = m32.get(n), m64.get(n)
        assert np.all(np.isfinite(a)), n
        assert np.array_equal(a, a.astype(np.float32).astype(np.float64)), n
        e = float(np.max(np.abs(a - b))) / float(np.max(np.abs(b)))
        assert 0.0 < e <= tol, (n, e)     # > 0: it really computes in Float32


def test_float32_oracle_quiescent_stays_at_rest(pkg, orc):
    c = cases.build_case("c1")
    m = c.make_model(pkg, orc.OracleBackendF32)
    for _ in range(5):
        pkg.time_step(m, c.dt)
    for n in ("U", "V", "W", "ETA", "T", "S"):
        assert np.max(np.abs(m.get(n))) == 0.0, n


# ---- single terms of the tendencies against closed forms (states built so that every other term vanishes) -----------------
# Independent of the oracle's code structure: the expected values are written from the discrete formulas of SURVEY Appendix A
# with numpy.  Grid metrics as Oceananigans defines them on a LatitudeLongitudeGrid: dx = R cos(phi) dlam, dy = R dphi,
# Az(c,c) = R^2 dlam (sin phi_f[j+1] - sin phi_f[j]).
EARTH_R, EARTH_OMEGA = 6371.0e3, 7.292115e-5


def _metrics(g):
    phic, phif = np.deg2rad(g.phi_c()), np.deg2rad(g.phi_f())
    dlam, dphi = np.deg2rad(g.dlam), np.deg2rad(g.dphi)
    return dict(phic=phic, phif=phif, dlam=dlam, dphi=dphi, dy=EARTH_R * dphi, dxfc=EARTH_R * np.cos(phic) * dlam,
                Az=EARTH_R ** 2 * dlam * (np.sin(phif[1:]) - np.sin(phif[:-1])))


def test_coriolis_term_alone(pkg, orc):
    """u = 0, v uniform, T = S = 0 on a flat bottom: zeta = 0, dK/dx = 0, u-advection = 0, no pressure gradient, so
    G^u = f-bar * (dx-weighted 4-point average of v) / dx(f,c)  (HydrostaticSphericalCoriolis, enstrophy-conserving form)."""
    m = cases.build_case("quiescent").make_model(pkg, orc.OracleBackend)
    M = _metrics(m.grid)
    V0 = 0.1
    pkg.set_model(m, v=lambda lam, phi, z: V0 + 0 * lam + 0 * phi + 0 * z)
    f = 2 * EARTH_OMEGA * np.sin(M["phif"])
    expect = 0.5 * (f[:-1] + f[1:]) * 0.5 * (np.cos(M["phif"][:-1]) + np.cos(M["phif"][1:])) / np.cos(M["phic"]) * V0
    GU = m.get("GU")
    inner = slice(8, -8)      # away from the walls, where v = 0 enters the stencils
    assert np.max(np.abs(GU[:, inner, :] - expect[None, inner, None])) <= 1e-14 * np.max(np.abs(expect))


def test_pressure_gradient_term_alone(pkg, orc):
    """u = v = 0 with horizontally varying T, S (TEOS-10): G^u = -dx(pHY') / dx(f,c), G^v = -dy(pHY') / dy, nothing else."""
    m = cases.build_case("quiescent").make_model(pkg, orc.OracleBackend)
    M = _metrics(m.grid)
    r = np.pi / 180
    pkg.set_model(m, T=lambda lam, phi, z: 10 + 3 * np.sin(2 * lam * r) * np.cos(3 * phi * r) * np.exp(z / 800.0),
                  S=lambda lam, phi, z: 35 + 0.3 * np.cos(lam * r) * np.sin(2 * phi * r) + 0 * z)
    P, GU, GV = m.get("P"), m.get("GU"), m.get("GV")
    eGU = -(P - np.roll(P, 1, axis=2)) / M["dxfc"][None, :, None]
    eGV = -(P[:, 1:, :] - P[:, :-1, :]) / M["dy"]
    assert np.max(np.abs(eGU)) > 1e-6
    assert np.max(np.abs(GU - eGU)) <= 1e-14 * np.max(np.abs(eGU))
    assert np.max(np.abs(GV[:, 1:-1, :] - eGV)) <= 1e-14 * np.max(np.abs(eGV))
    assert np.max(np.abs(GV[:, 0, :])) == 0.0 and np.max(np.abs(GV[:, -1, :])) == 0.0      # walls


@pytest.mark.parametrize("speed", [0.3, -0.2])
def test_tracer_advection_of_a_linear_profile_is_exact(pkg, orc, speed):
    """Non-divergent flows (zonal: u = U0 cos(phi); meridional: dx(c,f) v = const) advecting a tracer that is linear in the cell
    index: every WENO candidate reproduces the face value exactly, so the flux-form tendency is -(A u) dT / V for either
    upwind direction -- pins the flux form, the metrics (spherical Az in the volume) and the direction handling."""
    m = cases.build_case("quiescent").make_model(pkg, orc.OracleBackend)
    g, M = m.grid, _metrics(m.grid)
    r, s = np.pi / 180, 0.01
    lam0, phi0 = g.lam_c()[0], g.phi_c()[0]
    away = slice(6, -6)      # away from the periodic wrap of the linear profile / from the walls
    pkg.set_model(m, u=lambda lam, phi, z: speed * np.cos(phi * r) + 0 * lam + 0 * z,
                  T=lambda lam, phi, z: 5 + s * (lam - lam0) / g.dlam + 0 * phi + 0 * z, S=35.0)
    expect = -(speed * np.cos(M["phic"])) * s * M["dy"] / M["Az"]
    assert np.max(np.abs(m.get("W"))) == 0.0
    assert np.max(np.abs(m.get("GT")[:, :, away] - expect[None, :, None])) <= 1e-11 * np.max(np.abs(expect))
    pkg.set_model(m, u=0.0, v=lambda lam, phi, z: speed / np.cos(phi * r) + 0 * lam + 0 * z,
                  T=lambda lam, phi, z: 5 + s * (phi - phi0) / g.dphi + 0 * lam + 0 * z, S=35.0)
    expect = -(speed * EARTH_R * M["dlam"]) * s / M["Az"]
    assert np.max(np.abs(m.get("W")[:, away, :])) <= 1e-17
    assert np.max(np.abs(m.get("GT")[:, away, :] - expect[None, away, None])) <= 1e-11 * np.max(np.abs(expect[away]))


def test_surface_flux_boundary_conditions_alone(pkg, orc):
    """DoubleDrake at rest with uniform T, S: the only tendencies are the top-level flux divergences -J / dz of the wind stress,
    the salinity flux and the temperature restoring (src/bathymetry_and_forcings.jl:110-159,
    src/boundary_and_initial_conditions.jl:58-81)."""
    c = cases.build_case("dd_small")
    m = c.make_model(pkg, orc.OracleBackend, initialize=False)
    g = m.grid
    T0, S0 = 10.0, 35.0
    pkg.set_model(m, T=T0, S=S0)
    dz_top = g.z_faces[-1] - g.z_faces[-2]
    phi = g.phi_c()
    L = 75.0
    tau = np.array([pkg.wind_stress(p, pkg.wind_stress_coefficients(L)) for p in phi])
    Fs = np.array([pkg.salinity_flux(p, pkg.salinity_flux_coefficients(L, 35.0)) for p in phi])
    lam_T = 5.0 / (7 * 86400.0)
    Qt = lam_T * (T0 - pkg.T_reference(phi))
    GU, GV, GT, GS = (m.get(n) for n in ("GU", "GV", "GT", "GS"))
    wet = ~g.immersed_cell_mask()[-1]                      # top level, (Ny, Nx)
    wet_u = wet & np.roll(wet, 1, axis=1)
    for G, J, w in ((GU, tau, wet_u), (GS, Fs, wet), (GT, Qt, wet)):
        assert np.max(np.abs(G[:-1])) == 0.0                                   # nothing below the top level
        e = np.where(w, -J[:, None] / dz_top, 0.0)
        assert np.max(np.abs(G[-1] - e)) <= 1e-13 * np.max(np.abs(e))
    assert np.max(np.abs(GV)) == 0.0


def test_split_explicit_loop_and_corrector_against_a_numpy_restatement(pkg, orc):
    """A free-surface bump over a resting, unstratified flat-bottom ocean: the 3-D tendencies vanish, so one time step is the
    barotropic substep loop (eta first with U^m, then U, V with eta^(m+1); averages with the substep weights; eta <- average)
    followed by the corrector u(k) = Ubar / H.  Restated here with whole-array numpy operations."""
    c = cases.build_case("quiescent")
    m = c.make_model(pkg, orc.OracleBackend)
    g, fs = m.grid, m.free_surface
    M = _metrics(g)
    r = np.pi / 180
    pkg.set_model(m, eta=lambda lam, phi: 0.1 * np.cos(phi * r) * np.sin(2 * lam * r))
    eta = m.get("ETA").copy()
    Ny, Nx = eta.shape
    grav, H = 9.80665, g.Lz
    dxcf = EARTH_R * np.cos(M["phif"]) * M["dlam"]
    dtau = fs.fractional_step_size * c.dt
    U, V = np.zeros((Ny, Nx)), np.zeros((Ny + 1, Nx))       # V = 0 on the two walls
    etab, Ub, Vb = np.zeros_like(eta), np.zeros_like(U), np.zeros_like(V)
    for w in fs.averaging_weights:
        div = M["dy"] * (np.roll(U, -1, axis=1) - U) + dxcf[1:, None] * V[1:] - dxcf[:-1, None] * V[:-1]
        eta = eta - dtau * div / M["Az"][:, None]
        U = U + dtau * (-grav * H * (eta - np.roll(eta, 1, axis=1)) / M["dxfc"][:, None])
        V[1:Ny] = V[1:Ny] + dtau * (-grav * H * (eta[1:] - eta[:-1]) / M["dy"])
        etab += w * eta
        Ub += w * U
        Vb += w * V
    pkg.time_step(m, c.dt)

    def rel(a, b):
        return np.max(np.abs(a - b)) / np.max(np.abs(b))
    assert np.max(np.abs(Ub)) > 1.0      # the bump did set the ocean in motion (transport in m^2/s)
    assert rel(m.get("ETA"), etab) <= 1e-13 and rel(m.get("UBAR"), Ub) <= 1e-13 and rel(m.get("VBAR"), Vb) <= 1e-13
    u, v = m.get("U"), m.get("V")
    assert rel(u, np.broadcast_to(Ub / H, u.shape)) <= 1e-13 and rel(v, np.broadcast_to(Vb / H, v.shape)) <= 1e-13


@pytest.mark.parametrize("stable", [True, False])
def test_ri_based_diffusivities_against_the_closed_form(pkg, orc, stable):
    """Horizontally uniform u(z), T(z) on the flat part of DoubleDrake (linear EOS): Ri = N^2 / (du/dz)^2 on faces,
    tau = (1 - tanh((Ri - Ri0) / Ri_delta)) / 2, kappa_c = kappa0 tau (+ kappa_ca where N^2 < 0), kappa_u = nu0 tau.  The
    diffusivities are time filtered, so update_state! is repeated until the filter has forgotten its history."""
    grav, alpha = 9.80665, 1.67e-4
    m = cases.build_case("dd_small").make_model(pkg, orc.OracleBackend, initialize=False)
    zc = m.grid.z_c()

    def uf(z):
        return 0.4 * np.exp(z / 150.0) + 0.3 * z / 3000.0

    def Tf(z):
        return 10 + 0.02 * np.exp(z / 200.0) + 1e-5 * z if stable else 10 - 8 * np.exp(z / 400.0)
    pkg.set_model(m, u=lambda lam, phi, z: uf(z) + 0 * lam + 0 * phi, T=lambda lam, phi, z: Tf(z) + 0 * lam + 0 * phi, S=35.0)
    for _ in range(80):          # the filter keeps 0.6 / 1.6 of its previous value per update: 0.375^80 ~ 1e-34
        m.backend.update_state()
    dzf = zc[1:] - zc[:-1]
    N2 = grav * alpha * (Tf(zc)[1:] - Tf(zc)[:-1]) / dzf
    S2 = ((uf(zc)[1:] - uf(zc)[:-1]) / dzf) ** 2
    Ri = np.where(N2 <= 0, 0.0, N2 / S2)
    tau = 0.5 * (1 - np.tanh((Ri - 0.1) / 0.4))
    assert (np.all(N2 > 0) and Ri.min() < 0.1 < 1.0 < Ri.max()) if stable else np.all(N2 < 0)
    ekc = 0.5 * tau + np.where(N2 < 0, 1.7, 0.0)
    eku = 0.7 * tau
    kc, ku = m.get("KAPPA_C"), m.get("KAPPA_U")
    j, i = 20, 45                # far from the ridges and the walls
    assert np.max(np.abs(kc[1:-1, j, i] / ekc - 1)) <= 1e-12 and np.max(np.abs(ku[1:-1, j, i] / eku - 1)) <= 1e-12
    assert kc[0, j, i] == 0.0 and kc[-1, j, i] == 0.0


def test_second_step_uses_the_quasi_adams_bashforth_combination(pkg, orc):
    """After one (forward-Euler) step with the same dt, the next AB2 + implicit stage of S must be the backward-Euler solve of
    S + dt ((3/2 + chi) G^n - (1/2 + chi) G^-) with chi = 0.1 (near_global_simulation.jl:105-113: QuasiAdamsBashforth2 default)."""
    from scipy.linalg import solve_banded
    c = cases.build_case("bumpy")
    m = c.make_model(pkg, orc.OracleBackend)
    g, dt, chi = m.grid, c.dt, 0.1
    pkg.time_step(m, dt)
    S0, Gn, G0, kc = m.get("S"), m.get("GS"), m.get("GS_PREV"), m.get("KAPPA_C")
    assert np.max(np.abs(G0)) > 0 and np.max(np.abs(Gn - G0)) > 0
    m.backend.run_stage(pkg._abi.STAGE_ID["AB2_IMPLICIT"], dt, 0)
    S1 = m.get("S")
    zf, zc = g.z_faces, g.z_c()
    dzc = zf[1:] - zf[:-1]
    dzf = np.full(g.Nz + 1, np.nan)
    dzf[1:-1] = zc[1:] - zc[:-1]
    solid = g.immersed_cell_mask()
    worst, wrong = 0.0, 0.0
    for j in range(2, g.Ny, 5):
        for i in range(1, g.Nx, 7):
            wet = ~solid[:, j, i]
            if not wet.any():
                continue
            rhs = S0[:, j, i] + dt * ((1.5 + chi) * Gn[:, j, i] - (0.5 + chi) * G0[:, j, i])
            euler = S0[:, j, i] + dt * Gn[:, j, i]
            K = np.zeros(g.Nz + 1)
            for kf in range(1, g.Nz):
                if wet[kf - 1] and wet[kf]:
                    K[kf] = 3e-5 + kc[kf, j, i]
            ab = np.zeros((3, g.Nz))
            for k in range(g.Nz):
                up = -dt * K[k + 1] / (dzc[k] * dzf[k + 1]) if k < g.Nz - 1 else 0.0
                lo = -dt * K[k] / (dzc[k] * dzf[k]) if k > 0 else 0.0
                ab[1, k] = 1.0 - up - lo
                if k < g.Nz - 1:
                    ab[0, k + 1] = up
                if k > 0:
                    ab[2, k - 1] = lo
            worst = max(worst, float(np.max(np.abs(solve_banded((1, 1), ab, rhs) - S1[:, j, i]))) / 35.0)
            wrong = max(wrong, float(np.max(np.abs(solve_banded((1, 1), ab, euler) - S1[:, j, i]))) / 35.0)
    assert worst < 1e-14, worst
    assert wrong > 1e3 * worst       # the test can tell AB2 from forward Euler


# ---- momentum tendencies: the whole assembly at sample points against an independent numpy restatement ------------------
def _numpy_beta_and_candidates(N, vals):
    """Per sub-stencil: the value at the face of the polynomial with the given cell averages, and its smoothness indicator
    (sum over derivatives of the integral over the upwind cell), fitted numerically -- no coefficient tables."""
    from numpy.polynomial import polynomial as P
    scale = {2: 1.0, 3: 3.0, 4: 0.24, 5: 0.0504}[N]
    cand, beta = [], []
    for s in range(N):
        cells = np.arange(-1 - s, -1 - s + N)
        v = vals[N - 1 - s: 2 * N - 1 - s]
        A = np.array([[((c + 1) ** (m + 1) - c ** (m + 1)) / (m + 1) for m in range(N)] for c in cells])
        a = np.linalg.solve(A, v)
        cand.append(a[0])
        b = 0.0
        for l in range(1, N):
            d = P.polyder(a, l)
            I = P.polyint(P.polymul(d, d))
            b += P.polyval(0.0, I) - P.polyval(-1.0, I)
        beta.append(scale * b)
    return np.array(cand), np.array(beta)


def numpy_weno_z(N, q, smooth):
    """Z-weighted WENO of `q` (2N - 1 values, most upwind-ward first) with the smoothness measured on the fields in `smooth`
    (their indicators averaged): VelocityStencil for the vorticity, the full divergence for the divergence flux."""
    optimal = {2: [2 / 3, 1 / 3], 3: [3 / 10, 3 / 5, 1 / 10], 4: [4 / 35, 18 / 35, 12 / 35, 1 / 35],
               5: [5 / 126, 20 / 63, 10 / 21, 10 / 63, 1 / 126]}[N]
    tau_w = {2: [1, -1], 3: [1, 0, -1], 4: [1, 3, -3, -1], 5: [1, 2, -6, 2, 1]}[N]
    cand, _ = _numpy_beta_and_candidates(N, np.asarray(q, dtype=np.float64))
    beta = np.mean([_numpy_beta_and_candidates(N, np.asarray(f, dtype=np.float64))[1] for f in smooth], axis=0)
    tau = abs(float(np.dot(tau_w, beta)))
    alpha = np.array(optimal) * (1 + (tau / (beta + 1e-8)) ** 2)
    return float(np.sum(alpha / alpha.sum() * cand))


def test_momentum_tendencies_against_a_numpy_restatement_at_sample_points(pkg, orc):
    """G^u, G^v of the vector-invariant form on a flat-bottom ocean with uniform T, S (no pressure gradient), assembled term by
    term from SURVEY Appendix A with numpy: WENO-9 upwinded vorticity flux with velocity-stencil smoothness, WENO-5 upwinded
    divergence flux with the full divergence as smoothness field, centred vertical advection, centred kinetic-energy gradient,
    enstrophy-conserving Coriolis.  Points are at least 8 rows from the walls so that every stencil is full order."""
    c = cases.build_case("quiescent")
    m = c.make_model(pkg, orc.OracleBackend)
    g = m.grid
    M = _metrics(g)
    r, D = np.pi / 180, g.Lz
    pkg.set_model(m, u=lambda lam, phi, z: 0.1 * np.sin(2 * lam * r) * np.cos(phi * r) * (1.0 + z / D),
                  v=lambda lam, phi, z: 0.05 * np.cos(3 * lam * r) * np.sin(2 * phi * r) * np.exp(z / 1000.0))
    U, V, W, GU, GV = (m.get(n) for n in ("U", "V", "W", "GU", "GV"))
    Nz, Ny, Nx = U.shape
    dz = np.diff(g.z_faces)
    dy, dxfc, Azc = M["dy"], M["dxfc"], M["Az"]                       # dx at (f,c) latitudes, Az at centre latitudes
    dxcf = EARTH_R * np.cos(M["phif"]) * M["dlam"]                    # dx at face latitudes
    Azf = np.full(Ny + 1, np.nan)
    Azf[1:Ny] = EARTH_R ** 2 * M["dlam"] * (np.sin(M["phic"][1:]) - np.sin(M["phic"][:-1]))   # Az at face latitudes
    f = 2 * EARTH_OMEGA * np.sin(M["phif"])
    ix = lambda i: i % Nx

    def u_(k, j, i): return U[k, j, ix(i)]
    def v_(k, j, i): return V[k, j, ix(i)]
    def w_(k, j, i): return W[k, j, ix(i)]
    def zeta(k, j, i): return (dy * v_(k, j, i) - dy * v_(k, j, i - 1) - (dxfc[j] * u_(k, j, i) - dxfc[j - 1] * u_(k, j - 1, i))) / Azf[j]
    def us(k, j, i): return 0.5 * (u_(k, j - 1, i) + u_(k, j, i))
    def vs(k, j, i): return 0.5 * (v_(k, j, i - 1) + v_(k, j, i))
    def dxU(k, j, i): return dy * dz[k] * (u_(k, j, i + 1) - u_(k, j, i))
    def dyV(k, j, i): return dz[k] * (dxcf[j + 1] * v_(k, j + 1, i) - dxcf[j] * v_(k, j, i))
    def K(k, j, i): return (0.5 * (u_(k, j, i) ** 2 + u_(k, j, i + 1) ** 2) + 0.5 * (v_(k, j, i) ** 2 + v_(k, j + 1, i) ** 2)) / 2

    def Gu(k, j, i):
        vavg = 0.25 * (dxcf[j] * v_(k, j, i - 1) + dxcf[j + 1] * v_(k, j + 1, i - 1) + dxcf[j] * v_(k, j, i) + dxcf[j + 1] * v_(k, j + 1, i))
        vhat = vavg / dxfc[j]
        js = range(j - 4, j + 5) if vhat > 0 else range(j + 5, j - 4, -1)       # zeta lives on face latitudes; the node sits between j and j + 1
        zr = numpy_weno_z(5, [zeta(k, jj, i) for jj in js], [[us(k, jj, i) for jj in js], [vs(k, jj, i) for jj in js]])
        uh = u_(k, j, i)
        is_ = range(i - 3, i + 2) if uh > 0 else range(i + 2, i - 3, -1)        # divergence lives in cells; the node sits between i - 1 and i
        dr = numpy_weno_z(3, [dxU(k, j, ii) for ii in is_], [[dxU(k, j, ii) + dyV(k, j, ii) for ii in is_]])
        Phi = uh * dr + uh * 0.5 * (dyV(k, j, i - 1) + dyV(k, j, i))
        ub, ua = u_(max(k - 1, 0), j, i), u_(min(k + 1, Nz - 1), j, i)
        fl_lo = 0.5 * Azc[j] * (w_(k, j, i - 1) + w_(k, j, i)) * 0.5 * (ub + uh)
        fl_hi = 0.5 * Azc[j] * (w_(k + 1, j, i - 1) + w_(k + 1, j, i)) * 0.5 * (uh + ua)
        return (vhat * zr - (Phi + fl_hi - fl_lo) / (Azc[j] * dz[k]) - (K(k, j, i) - K(k, j, i - 1)) / dxfc[j]
                + 0.5 * (f[j] + f[j + 1]) * vavg / dxfc[j])

    def Gv(k, j, i):
        uavg = 0.25 * (u_(k, j - 1, i) + u_(k, j - 1, i + 1) + u_(k, j, i) + u_(k, j, i + 1))
        is_ = range(i - 4, i + 5) if uavg > 0 else range(i + 5, i - 4, -1)
        zr = numpy_weno_z(5, [zeta(k, j, ii) for ii in is_], [[us(k, j, ii) for ii in is_], [vs(k, j, ii) for ii in is_]])
        vh = v_(k, j, i)
        js = range(j - 3, j + 2) if vh > 0 else range(j + 2, j - 3, -1)
        dr = numpy_weno_z(3, [dyV(k, jj, i) for jj in js], [[dxU(k, jj, i) + dyV(k, jj, i) for jj in js]])
        Phi = vh * dr + vh * 0.5 * (dxU(k, j - 1, i) + dxU(k, j, i))
        vb, va = v_(max(k - 1, 0), j, i), v_(min(k + 1, Nz - 1), j, i)
        fl_lo = 0.5 * (Azc[j - 1] * w_(k, j - 1, i) + Azc[j] * w_(k, j, i)) * 0.5 * (vb + vh)
        fl_hi = 0.5 * (Azc[j - 1] * w_(k + 1, j - 1, i) + Azc[j] * w_(k + 1, j, i)) * 0.5 * (vh + va)
        return (-uavg * zr - (Phi + fl_hi - fl_lo) / (Azf[j] * dz[k]) - (K(k, j, i) - K(k, j - 1, i)) / dy - f[j] * uavg)

    su, sv = float(np.max(np.abs(GU))), float(np.max(np.abs(GV)))
    worst_u = worst_v = 0.0
    seen = set()
    for k in (0, 3, Nz - 1):
        for j in (9, 14, 20, 26, 30):
            for i in (0, 1, 17, 44, 45, 71, Nx - 1):
                worst_u = max(worst_u, abs(Gu(k, j, i) - GU[k, j, i]) / su)
                worst_v = max(worst_v, abs(Gv(k, j, i) - GV[k, j, i]) / sv)
                seen.add((V[k, j, i] > 0, U[k, j, i] > 0))
    assert len(seen) == 4                       # both upwind directions of both components were exercised
    assert worst_u <= 1e-12 and worst_v <= 1e-12, (worst_u, worst_v)


def test_tracer_tendencies_against_a_numpy_restatement_at_sample_points(pkg, orc):
    """G^T, G^S in flux form: WENO-7 upwinded reconstructions of the tracer at the x and y faces (smoothness on the tracer itself),
    centred vertical flux, divided by the cell volume Az dz -- at sample points of a flat-bottom ocean, both upwind directions."""
    c = cases.build_case("quiescent")
    m = c.make_model(pkg, orc.OracleBackend)
    g = m.grid
    M = _metrics(g)
    r, D = np.pi / 180, g.Lz
    pkg.set_model(m, u=lambda lam, phi, z: 0.1 * np.sin(2 * lam * r) * np.cos(phi * r) * (1.0 + z / D),
                  v=lambda lam, phi, z: 0.05 * np.cos(3 * lam * r) * np.sin(2 * phi * r) * np.exp(z / 1000.0),
                  T=lambda lam, phi, z: 10 + 3 * np.sin(3 * lam * r) * np.cos(2 * phi * r) * np.exp(z / 500.0),
                  S=lambda lam, phi, z: 35 + 0.2 * np.cos(2 * lam * r + 0.3) * np.sin(3 * phi * r) * np.exp(z / 800.0))
    U, V, W = (m.get(n) for n in ("U", "V", "W"))
    Nz, Ny, Nx = U.shape
    dz = np.diff(g.z_faces)
    dy, Azc = M["dy"], M["Az"]
    dxcf = EARTH_R * np.cos(M["phif"]) * M["dlam"]
    seen = set()
    for name, gname in (("T", "GT"), ("S", "GS")):
        C, G = m.get(name), m.get(gname)

        def c_(k, j, i): return C[k, j, i % Nx]

        def Fx(k, j, i):          # through the face between cells i - 1 and i
            a = U[k, j, i % Nx]
            cells = range(i - 4, i + 3) if a > 0 else range(i + 3, i - 4, -1)
            q = [c_(k, j, ii) for ii in cells]
            return dy * dz[k] * a * numpy_weno_z(4, q, [q])

        def Fy(k, j, i):          # through the face between cells j - 1 and j
            a = V[k, j, i % Nx]
            cells = range(j - 4, j + 3) if a > 0 else range(j + 3, j - 4, -1)
            q = [c_(k, jj, i) for jj in cells]
            return dxcf[j] * dz[k] * a * numpy_weno_z(4, q, [q])

        def Fz(k, j, i):          # through the face below level k; the face above the top level sees the top value on both sides
            return Azc[j] * W[k, j, i % Nx] * 0.5 * (c_(max(k - 1, 0), j, i) + c_(min(k, Nz - 1), j, i))
        scale, worst = float(np.max(np.abs(G))), 0.0
        for k in (0, 3, Nz - 1):
            for j in (9, 14, 21, 26, 30):
                for i in (0, 1, 17, 44, 45, 71, Nx - 1):
                    e = -((Fx(k, j, i + 1) - Fx(k, j, i)) + (Fy(k, j + 1, i) - Fy(k, j, i)) + (Fz(k + 1, j, i) - Fz(k, j, i))) / (Azc[j] * dz[k])
                    worst = max(worst, abs(e - G[k, j, i]) / scale)
                    seen.add((U[k, j, i] > 0, V[k, j, i] > 0))
        assert worst <= 1e-11, (name, worst)
    assert len(seen) == 4


def test_tracer_tendencies_next_to_immersed_boundaries_and_walls(pkg, orc):
    """The same restatement on the k-dependent "bumpy" bathymetry, at wet cells that have a solid cell or a wall within four
    cells: a face is reconstructed at the largest order N <= 4 whose 2N cells (i - N .. i + N - 1, the union of the two biased
    stencils) are all wet and inside the domain, down to first-order upwind (DESIGN.md spec decision 1)."""
    c = cases.build_case("bumpy")
    m = c.make_model(pkg, orc.OracleBackend)
    g = m.grid
    M = _metrics(g)
    U, V, W, C, G = (m.get(n) for n in ("U", "V", "W", "T", "GT"))
    Nz, Ny, Nx = U.shape
    solid = g.immersed_cell_mask()
    dz = np.diff(g.z_faces)
    dy, Azc = M["dy"], M["Az"]
    dxcf = EARTH_R * np.cos(M["phif"]) * M["dlam"]

    def wet(k, j, i): return 0 <= j < Ny and not solid[k, j, i % Nx]

    def recon(values_at, ok_at, a):
        """values_at(n) / ok_at(n): cell n relative to the face (n = -1 is the cell behind it, n = 0 the one ahead)"""
        N = next((n for n in (4, 3, 2, 1) if all(ok_at(q) for q in range(-n, n))), 0)
        if N == 0:
            return 0.0          # a face between a wet and a solid cell: the velocity there is masked to zero
        cells = range(-N, N - 1) if a > 0 else range(N - 1, -N, -1)
        q = [values_at(n) for n in cells]
        return q[0] if N == 1 else numpy_weno_z(N, q, [q])

    def Fx(k, j, i):
        a = U[k, j, i % Nx]
        return dy * dz[k] * a * recon(lambda n: C[k, j, (i + n) % Nx], lambda n: wet(k, j, i + n), a)

    def Fy(k, j, i):
        a = V[k, j, i % Nx]
        return dxcf[j] * dz[k] * a * recon(lambda n: C[k, j + n, i % Nx], lambda n: wet(k, j + n, i), a)

    def Fz(k, j, i):
        return Azc[j] * W[k, j, i % Nx] * 0.5 * (C[max(k - 1, 0), j, i % Nx] + C[min(k, Nz - 1), j, i % Nx])

    near = [(k, j, i) for k in range(Nz) for j in range(Ny) for i in range(Nx)
            if wet(k, j, i) and not all(wet(k, j + dj, i + di) for dj, di in
                                        [(d, 0) for d in range(-4, 5)] + [(0, d) for d in range(-4, 5)])]
    assert len(near) > 500
    scale, worst, orders = float(np.max(np.abs(G))), 0.0, set()
    for k, j, i in near[::max(1, len(near) // 150)]:
        e = -((Fx(k, j, i + 1) - Fx(k, j, i)) + (Fy(k, j + 1, i) - Fy(k, j, i)) + (Fz(k + 1, j, i) - Fz(k, j, i))) / (Azc[j] * dz[k])
        if k == Nz - 1:         # DoubleDrake restores the surface temperature (boundary_and_initial_conditions.jl:58-81)
            e -= 5.0 / (7 * 86400.0) * (C[k, j, i] - pkg.T_reference(g.phi_c()[j])) / dz[k]
        worst = max(worst, abs(e - G[k, j, i]) / scale)
        orders.add(next((n for n in (4, 3, 2, 1) if all(wet(k, j, i + q) for q in range(-n, n))), 0))
    assert orders >= {1, 2, 3}, orders          # reduced orders were really exercised
    assert worst <= 1e-11, worst
    assert np.max(np.abs(G[solid])) == 0.0      # tendencies are defined as zero in solid cells (spec decision 2)


# ---- consistency with the continuous equations (independent of any recollection of Oceananigans' source) -----------------
def test_tendencies_converge_to_the_primitive_equations(pkg, orc):
    """For a smooth analytic state on the sphere the discrete tendencies must converge to the tendencies of the hydrostatic
    primitive equations in vector-invariant form,
        G^u = (zeta + f) v - dK/dx - w du/dz - dp'/dx,  G^v = -(zeta + f) u - dK/dy - w dv/dz - dp'/dy,  G^T = -u . grad T,
    with w from continuity and p' the hydrostatic integral of the (linear-EOS) buoyancy, all evaluated with sympy.  Signs,
    metric factors (cos phi), the Coriolis parameter, the kinetic-energy and pressure gradients and the vertical advection are
    checked against the PDE itself.  The top level is left out: the zero-gradient tracer / velocity value above the surface
    makes the vertical advection of the top cell first-order with an O(1) coefficient (a property of the scheme, not an error)."""
    import sympy as sp
    a_, Om, grav, alpha, D = EARTH_R, EARTH_OMEGA, 9.80665, 1.67e-4, sp.Integer(3000)
    lam, phi, z, zp = sp.symbols("lam phi z zp", real=True)
    fu, fv, fT = 1 + z / D, sp.exp(z / 1000), sp.exp(z / 500)                   # vertical structure (integrated over z alone)
    Uh, Vh, Th = sp.sin(2 * lam) * sp.cos(phi) / 10, sp.cos(3 * lam) * sp.sin(2 * phi) / 20, 3 * sp.sin(3 * lam) * sp.cos(2 * phi)
    u, v, T = Uh * fu, Vh * fv, 10 + Th * fT
    cosp = sp.cos(phi)
    Iu, Iv = sp.integrate(fu.subs(z, zp), (zp, -D, z)), sp.integrate(fv.subs(z, zp), (zp, -D, z))
    w = -(sp.diff(Uh, lam) * Iu + sp.diff(Vh * cosp, phi) * Iv) / (a_ * cosp)     # w(-D) = 0, dw/dz = -div_h(u, v)
    zeta = (sp.diff(v, lam) - sp.diff(u * cosp, phi)) / (a_ * cosp)
    K = (u ** 2 + v ** 2) / 2
    p = -grav * alpha * (-10 * z + Th * sp.integrate(fT.subs(z, zp), (zp, z, 0)))   # dp'/dz = b = g alpha T, p'(0) = 0
    f = 2 * Om * sp.sin(phi)
    exprs = {"u": u, "v": v, "T": T, "W": w,
             "GU": (zeta + f) * v - sp.diff(K, lam) / (a_ * cosp) - w * sp.diff(u, z) - sp.diff(p, lam) / (a_ * cosp),
             "GV": -(zeta + f) * u - sp.diff(K, phi) / a_ - w * sp.diff(v, z) - sp.diff(p, phi) / a_,
             "GT": -u * sp.diff(T, lam) / (a_ * cosp) - v * sp.diff(T, phi) / a_ - w * sp.diff(T, z)}
    fn = {n: sp.lambdify((lam, phi, z), e, "numpy") for n, e in exprs.items()}
    r = np.pi / 180

    def errors(Nx, Ny, Nz):
        grid = cases.Case("conv", Nx, Ny, Nz, experiment="Quiescent", perturb=False).grid(pkg)       # flat bottom, no forcing
        m = pkg.HydrostaticFreeSurfaceModel(grid=grid, free_surface=pkg.SplitExplicitFreeSurface(substeps=30), eos="linear",
                                            boundary_conditions=pkg.set_boundary_conditions("Quiescent", grid),
                                            backend_factory=orc.OracleBackend)
        pkg.set_model(m, u=lambda l, q, zz: fn["u"](l * r, q * r, zz), v=lambda l, q, zz: fn["v"](l * r, q * r, zz),
                      T=lambda l, q, zz: fn["T"](l * r, q * r, zz), S=35.0)
        g = m.grid
        lc, lf, pc, pf, zc = g.lam_c() * r, (g.lam_c() - 0.5 * g.dlam) * r, g.phi_c() * r, g.phi_f() * r, g.z_c()
        out = {}
        for name, L, P, Z in (("GU", lf, pc, zc), ("GV", lc, pf, zc), ("GT", lc, pc, zc), ("W", lc, pc, g.z_faces)):
            ZZ, PP, LL = np.meshgrid(Z, P, L, indexing="ij")
            A = fn[name](LL, PP, ZZ)
            sel = (np.abs(PP) < 50 * r) & ((ZZ < zc[-1]) | (name == "W"))      # away from the walls; below the top level
            out[name] = float(np.max(np.abs(m.get(name) - A)[sel]) / np.max(np.abs(A[sel])))
        return out
    coarse, fine = errors(90, 40, 8), errors(180, 80, 16)
    print("\nrelative L-infinity distance to the PDE tendencies, 90x40x8 -> 180x80x16:",
          {n: (round(coarse[n], 5), round(fine[n], 5)) for n in coarse})
    for n, bar in (("GU", 0.025), ("GV", 0.015), ("GT", 0.07), ("W", 0.01)):
        assert coarse[n] < bar, (n, coarse[n])
        assert fine[n] < 0.4 * coarse[n], (n, coarse[n], fine[n])


def test_equatorial_kelvin_wave_travels_east_at_the_gravity_wave_speed(pkg, orc):
    """Whole-model physics, independent of any recollection: an equatorially trapped free-surface bump with u = (g / c) eta, uniform
    in depth, over a flat 3000 m ocean is an equatorial Kelvin wave; it must travel EASTWARD at c = sqrt(g H) = 171.5 m/s.
    Thirty steps of 1200 s carry it 55.5 degrees of longitude (through the split-explicit solver, the barotropic corrector,
    Coriolis and the beta effect of the spherical grid)."""
    R, Om, grav = EARTH_R, EARTH_OMEGA, 9.80665
    c = cases.build_case("quiescent")
    m = c.make_model(pkg, orc.OracleBackend)
    g = m.grid
    cw = np.sqrt(grav * g.Lz)
    trap = np.sqrt(cw / (2 * Om / R))                      # equatorial deformation radius sqrt(c / beta), 24.6 degrees
    r = np.pi / 180

    def eta0(lam, phi):
        return 0.01 * np.exp(-((lam + 60.0) / 12.0) ** 2) * np.exp(-0.5 * (R * phi * r / trap) ** 2)
    pkg.set_model(m, eta=eta0, u=lambda lam, phi, z: (grav / cw) * eta0(lam, phi) + 0 * z)
    pkg.initialize_free_surface(m)        # the substeps start from U-bar: give them the barotropic part of the initial u

    def peak(e):
        row = 0.5 * (e[g.Ny // 2 - 1] + e[g.Ny // 2])      # the two rows next to the equator
        i = int(np.argmax(row))
        a, b, d = row[i - 1], row[i], row[(i + 1) % g.Nx]
        return g.lam_c()[i] + 0.5 * (a - d) / (a - 2 * b + d) * g.dlam, float(b)
    lam0, amp0 = peak(m.get("ETA"))
    steps, dt = 30, 1200.0
    for _ in range(steps):
        pkg.time_step(m, dt)
    lam1, amp1 = peak(m.get("ETA"))
    expected = cw * steps * dt / (R * r)
    assert abs(lam0 + 60.0) < 1e-9 and abs(expected - 55.53) < 0.01
    assert abs((lam1 - lam0) - expected) < 0.03 * expected, (lam1 - lam0, expected)
    assert amp1 > 0.8 * amp0
    eta = m.get("ETA")
    equatorial = np.abs(g.phi_c()) < 8.0
    west = eta[equatorial][:, (g.lam_c() > -170) & (g.lam_c() < -80)]
    assert np.max(np.abs(west)) < 0.25 * amp1             # along the equator nothing comparable went west


def test_teos10_pressure_gradient_converges_to_the_hydrostatic_integral(pkg, orc):
    """Ocean at rest with horizontally varying T, S and the TEOS-10 equation of state: G^u must converge to -dp'/dx with
    p'(z) = -int_z^0 b dz', b = -g rho'(T, S, z) / rho_0, evaluated here by Gauss-Legendre quadrature of the (known-answer-pinned)
    polynomial -- sign of the buoyancy, reference density, depth convention and the pressure integral against the continuous
    statement."""
    R, grav, rho0 = EARTH_R, 9.80665, 1020.0
    lib = orc.load()
    r = np.pi / 180

    def Tf(lam, phi, z): return 10 + 3 * np.sin(2 * lam * r) * np.cos(3 * phi * r) * np.exp(z / 800.0)
    def Sf(lam, phi, z): return 35 + 0.3 * np.cos(lam * r) * np.sin(2 * phi * r) + 0 * z
    xs, ws = np.polynomial.legendre.leggauss(24)

    def pprime(lam, phi, z):
        zz, wq = -0.5 * z * xs + 0.5 * z, -0.5 * z * ws
        b = [-grav * lib.oracle_teos10_rprime(float(Tf(lam, phi, q)), float(Sf(lam, phi, q)), float(q)) / rho0 for q in zz]
        return -float(np.sum(wq * np.array(b)))

    def error(Nx, Ny, Nz):
        m = cases.Case("t", Nx, Ny, Nz, experiment="Quiescent", perturb=False).make_model(pkg, orc.OracleBackend)
        g = m.grid
        pkg.set_model(m, T=Tf, S=Sf)
        GU, zc, lamf, phic = m.get("GU"), g.z_c(), g.lam_c() - 0.5 * g.dlam, g.phi_c()
        worst = scale = 0.0
        d = 0.05       # degrees: centred difference of the quadrature
        for k in range(0, Nz, Nz // 4):
            for j in range(3, Ny - 3, Ny // 8):
                for i in range(0, Nx, Nx // 9):
                    e = -(pprime(lamf[i] + d, phic[j], zc[k]) - pprime(lamf[i] - d, phic[j], zc[k])) / (2 * d * r * R * np.cos(phic[j] * r))
                    worst, scale = max(worst, abs(GU[k, j, i] - e)), max(scale, abs(e))
        return worst / scale
    coarse, fine = error(90, 40, 8), error(180, 80, 16)
    assert coarse < 0.05 and fine < 0.4 * coarse, (coarse, fine)


@pytest.mark.parametrize("name", ["bumpy", "realistic", "dd_leith"])
def test_upwind_side_only_timing_option_does_not_change_a_bit(pkg, orc, name):
    """bench.py's CPU legs time the oracle with only the upwind-side reconstruction evaluated (0.5 ((a + |a|) L + (a - |a|) R)
    is a L or a R exactly); the checker's default evaluates both sides.  Same bits."""
    c = cases.build_case(name)
    a, b = c.make_model(pkg, orc.OracleBackend), c.make_model(pkg, orc.OracleBackend)
    b.backend.set_option("upwind_side_only", 1)
    for _ in range(4):
        pkg.time_step(a, c.dt)
        pkg.time_step(b, c.dt)
    for n in ("U", "V", "W", "ETA", "T", "S", "GU", "GV", "GT", "GS"):
        assert np.array_equal(a.get(n), b.get(n)), n


def test_momentum_tendencies_next_to_immersed_boundaries_and_walls(pkg, orc):
    """G^u, G^v on the k-dependent "bumpy" DoubleDrake bathymetry at nodes with a solid cell or a wall within five cells, restated
    with numpy from the immersed-boundary rules of SURVEY Appendix A / DESIGN.md spec decisions 1-2:
      * a node on a face is INACTIVE when both cells it separates are solid or outside the domain, PERIPHERAL when either is;
      * the differences inside zeta are zero when they touch an inactive velocity node (conditional differences);
      * the vorticity is reconstructed at the largest order N <= 5 whose face nodes j - N + 1 .. j + N ((c,f,c) nodes for G^u;
        i - N + 1 .. i + N, (f,c,c) nodes for G^v) are all active; the divergence at the largest N <= 3 whose cells are all wet;
      * vertical fluxes through immersed-peripheral faces vanish; Coriolis averages over the non-peripheral nodes only;
      * tendencies are zero on peripheral nodes.
    Plus the DoubleDrake boundary conditions (wind stress on u at the top, linear drag at the bottom level)."""
    c = cases.build_case("bumpy")
    m = c.make_model(pkg, orc.OracleBackend)
    g = m.grid
    M = _metrics(g)
    U, V, W, P, GU, GV = (m.get(n) for n in ("U", "V", "W", "P", "GU", "GV"))
    Nz, Ny, Nx = U.shape
    solid = g.immersed_cell_mask()
    dz = np.diff(g.z_faces)
    dy, dxfc, Azc = M["dy"], M["dxfc"], M["Az"]
    dxcf = EARTH_R * np.cos(M["phif"]) * M["dlam"]
    Azf = np.full(Ny + 1, np.nan)
    Azf[1:Ny] = EARTH_R ** 2 * M["dlam"] * (np.sin(M["phic"][1:]) - np.sin(M["phic"][:-1]))
    f = 2 * EARTH_OMEGA * np.sin(M["phif"])
    tau = [pkg.wind_stress(p, pkg.wind_stress_coefficients(75.0)) for p in g.phi_c()]
    mu = 0.003

    def wet(k, j, i): return 0 <= j < Ny and 0 <= k < Nz and not solid[k, j, i % Nx]
    def u_(k, j, i): return U[k, j, i % Nx] if 0 <= j < Ny else 0.0
    def v_(k, j, i): return V[k, j, i % Nx] if 0 <= j <= Ny else 0.0
    def inactive_u(k, j, i): return not wet(k, j, i - 1) and not wet(k, j, i)          # (f,c,c)
    def inactive_v(k, j, i): return not wet(k, j - 1, i) and not wet(k, j, i)          # (c,f,c)
    def periph_u(k, j, i): return not wet(k, j, i - 1) or not wet(k, j, i)
    def periph_v(k, j, i): return not wet(k, j - 1, i) or not wet(k, j, i)

    def zeta(k, j, i):
        dxv = 0.0 if inactive_v(k, j, i - 1) or inactive_v(k, j, i) else dy * v_(k, j, i) - dy * v_(k, j, i - 1)
        dyu = 0.0 if inactive_u(k, j - 1, i) or inactive_u(k, j, i) else dxfc[j] * u_(k, j, i) - dxfc[j - 1] * u_(k, j - 1, i)
        return (dxv - dyu) / Azf[j]
    def us(k, j, i): return 0.5 * (u_(k, j - 1, i) + u_(k, j, i))
    def vs(k, j, i): return 0.5 * (v_(k, j, i - 1) + v_(k, j, i))
    def dxU(k, j, i): return dy * dz[k] * (u_(k, j, i + 1) - u_(k, j, i))
    def dyV(k, j, i): return dz[k] * (dxcf[j + 1] * v_(k, j + 1, i) - dxcf[j] * v_(k, j, i))
    def K(k, j, i): return (0.5 * (u_(k, j, i) ** 2 + u_(k, j, i + 1) ** 2) + 0.5 * (v_(k, j, i) ** 2 + v_(k, j + 1, i) ** 2)) / 2

    def biased(N, a):        # offsets of the 2N - 1 points, most upwind-ward first, relative to the point just behind the node
        return range(-N + 1, N) if a > 0 else range(N, -N + 1, -1)

    orders = set()

    def weno(N, q, smooth):
        orders.add((len(smooth), N))          # (2 smoothness fields: vorticity; 1: divergence), order
        return q[0] if N == 1 else numpy_weno_z(N, q, smooth)

    def Gu(k, j, i):
        if periph_u(k, j, i):
            return 0.0
        vavg = 0.25 * (dxcf[j] * v_(k, j, i - 1) + dxcf[j + 1] * v_(k, j + 1, i - 1) + dxcf[j] * v_(k, j, i) + dxcf[j + 1] * v_(k, j + 1, i))
        vhat = vavg / dxfc[j]
        G = 0.0
        if vhat != 0.0:
            N = next(n for n in (5, 4, 3, 2, 1) if not any(inactive_v(k, jj, i) for jj in range(j - n + 1, j + n + 1)))
            js = [j + o for o in biased(N, vhat)]
            G = vhat * weno(N, [zeta(k, jj, i) for jj in js], [[us(k, jj, i) for jj in js], [vs(k, jj, i) for jj in js]])
        uh = u_(k, j, i)
        Phi = 0.0
        if uh != 0.0:
            N = next(n for n in (3, 2, 1) if all(wet(k, j, ii) for ii in range(i - n, i + n)))
            is_ = [i - 1 + o for o in biased(N, uh)]
            Phi = uh * weno(N, [dxU(k, j, ii) for ii in is_], [[dxU(k, j, ii) + dyV(k, j, ii) for ii in is_]]) \
                + uh * 0.5 * (dyV(k, j, i - 1) + dyV(k, j, i))
        def wflux(kf):          # through the (f,c,f) face below level kf
            if 0 < kf < Nz and (periph_u(kf - 1, j, i) or periph_u(kf, j, i)):
                return 0.0
            return 0.5 * Azc[j] * (W[kf, j, (i - 1) % Nx] + W[kf, j, i % Nx]) * 0.5 * (u_(max(kf - 1, 0), j, i) + u_(min(kf, Nz - 1), j, i))
        G -= (Phi + wflux(k + 1) - wflux(k)) / (Azc[j] * dz[k])
        G -= (K(k, j, i) - K(k, j, i - 1)) / dxfc[j]
        nact = sum(not periph_v(k, j + b, i - 1 + a) for a in (0, 1) for b in (0, 1))
        if nact:
            G += 0.5 * (f[j] + f[j + 1]) * (vavg / (0.25 * nact)) / dxfc[j]
        G -= (P[k, j, i % Nx] - P[k, j, (i - 1) % Nx]) / dxfc[j]
        if k == Nz - 1:
            G -= tau[j] / dz[k]
        if k == 0:
            G += -mu * uh / dz[k]
        return G

    def Gv(k, j, i):
        if periph_v(k, j, i):
            return 0.0
        uavg = 0.25 * (u_(k, j - 1, i) + u_(k, j - 1, i + 1) + u_(k, j, i) + u_(k, j, i + 1))
        G = 0.0
        if uavg != 0.0:
            N = next(n for n in (5, 4, 3, 2, 1) if not any(inactive_u(k, j, ii) for ii in range(i - n + 1, i + n + 1)))
            is_ = [i + o for o in biased(N, uavg)]
            G = -uavg * weno(N, [zeta(k, j, ii) for ii in is_], [[us(k, j, ii) for ii in is_], [vs(k, j, ii) for ii in is_]])
        vh = v_(k, j, i)
        Phi = 0.0
        if vh != 0.0:
            N = next(n for n in (3, 2, 1) if all(wet(k, jj, i) for jj in range(j - n, j + n)))
            js = [j - 1 + o for o in biased(N, vh)]
            Phi = vh * weno(N, [dyV(k, jj, i) for jj in js], [[dxU(k, jj, i) + dyV(k, jj, i) for jj in js]]) \
                + vh * 0.5 * (dxU(k, j - 1, i) + dxU(k, j, i))
        def wflux(kf):
            if 0 < kf < Nz and (periph_v(kf - 1, j, i) or periph_v(kf, j, i)):
                return 0.0
            return 0.5 * (Azc[j - 1] * W[kf, j - 1, i % Nx] + Azc[j] * W[kf, j, i % Nx]) * 0.5 * (v_(max(kf - 1, 0), j, i) + v_(min(kf, Nz - 1), j, i))
        G -= (Phi + wflux(k + 1) - wflux(k)) / (Azf[j] * dz[k])
        G -= (K(k, j, i) - K(k, j - 1, i)) / dy
        nact = sum(not periph_u(k, j - 1 + b, i + a) for a in (0, 1) for b in (0, 1))
        if nact:
            G -= f[j] * (uavg / (0.25 * nact))
        G -= (P[k, j, i % Nx] - P[k, j - 1, i % Nx]) / dy
        if k == 0:
            G += -mu * vh / dz[k]
        return G

    def near_boundary(k, j, i):
        return not all(wet(k, j + dj, i + di) for dj, di in [(d, 0) for d in range(-5, 6)] + [(0, d) for d in range(-5, 6)])
    su, sv = float(np.max(np.abs(GU))), float(np.max(np.abs(GV)))
    pts = [(k, j, i) for k in range(Nz) for j in range(Ny) for i in range(Nx) if wet(k, j, i) and near_boundary(k, j, i)]
    assert len(pts) > 1000
    worst_u = worst_v = 0.0
    nper = 0
    for k, j, i in pts[::max(1, len(pts) // 160)]:
        worst_u = max(worst_u, abs(Gu(k, j, i) - GU[k, j, i]) / su)
        worst_v = max(worst_v, abs(Gv(k, j, i) - GV[k, j, i]) / sv)
        nper += periph_u(k, j, i) + periph_v(k, j, i)
    assert nper > 10                               # peripheral nodes (zero tendency) were among the samples
    assert {(2, n) for n in (1, 2, 3, 4, 5)} <= orders and {(1, n) for n in (1, 2, 3)} <= orders, orders   # every order was used
    assert worst_u <= 1e-11 and worst_v <= 1e-11, (worst_u, worst_v)


@pytest.mark.parametrize("day", [1.7, 4.2, 7.9, 13.1])
def test_realistic_ocean_surface_fluxes_follow_the_reference_formulas(pkg, orc, day):
    """RealisticOcean at rest: the top-level tendencies are -J / dz with the reference's own formulas
    (src/bathymetry_and_forcings.jl:162-193): fluxes interpolated linearly between the daily slices n1 = floor(mod(t, 5)) + 1 and
    n1 + 1; restoring values taken from slice mod(t, 15) div 3 + 1 WITHOUT interpolation (`÷` makes n an integer, so the weight of
    slice n2 is zero -- a quirk of the reference that is kept); lambda = min dz / 30 days (T), / 60 days (S)."""
    c = cases.build_case("realistic")
    m = c.make_model(pkg, orc.OracleBackend, initialize=False)
    g = m.grid
    T0, S0 = pkg.synthetic.initial_T_S(g)
    pkg.set_model(m, T=T0, S=S0)
    m.backend.set_clock(day * 86400.0, 0, float("inf"))
    m.backend.update_state()
    F = pkg.synthetic.forcing_arrays(g)
    n = day % 5.0
    n1 = int(np.floor(n))
    w2 = n - n1

    def interp(a): return a[n1] * (1.0 - w2) + a[n1 + 1] * w2
    nr = int((day % 15.0) // 3.0)
    dz = np.diff(g.z_faces)
    lamT, lamS = dz.min() / (30 * 86400.0), dz.min() / (60 * 86400.0)
    solid = g.immersed_cell_mask()[-1]
    wet_u = ~solid & ~np.roll(solid, 1, axis=1)
    wet_v = np.zeros((g.Ny + 1, g.Nx), bool)
    wet_v[1:-1] = ~solid[1:] & ~solid[:-1]
    T, S = m.get("T")[-1], m.get("S")[-1]
    expect = {"GU": np.where(wet_u, -interp(F["taux"]) / dz[-1], 0.0), "GV": np.where(wet_v, -interp(F["tauy"]) / dz[-1], 0.0),
              "GT": np.where(~solid, -(interp(F["Qs"]) + lamT * (T - F["Tr"][nr])) / dz[-1], 0.0),
              "GS": np.where(~solid, -(interp(F["Fs"]) + lamS * (S - F["Sr"][nr])) / dz[-1], 0.0)}
    for name, e in expect.items():
        G = m.get(name)[-1]
        if name in ("GU", "GV"):      # the pressure gradient of the stratified state is there as well: compare what is left
            P = m.get("P")[-1]
            M = _metrics(g)
            G = G + ((P - np.roll(P, 1, axis=1)) / M["dxfc"][:, None] if name == "GU"
                     else np.vstack([np.zeros((1, g.Nx)), (P[1:] - P[:-1]) / M["dy"], np.zeros((1, g.Nx))]))
            G = np.where(wet_u if name == "GU" else wet_v, G, 0.0)
        assert np.max(np.abs(G - e)) <= 1e-12 * np.max(np.abs(e)) + 1e-22, (name, day)
