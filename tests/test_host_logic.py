"""CPU tests of the host-side mirror of the reference's Julia configuration code."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest


def test_exponential_z_faces_matches_reference_formula(pkg):
    # bathymetry_and_forcings.jl:47-58, Nz = 120, Depth = 3000 (BASELINE configs 2, 3)
    zf = pkg.exponential_z_faces(120, 3000)
    assert zf.shape == (121,)
    assert abs(zf[0] + 3000.0) < 1e-9 and zf[-1] == 0.0
    dz = np.diff(zf)
    assert np.all(dz > 0)
    # dz from 1.29 m at the top to 111.7 m at the bottom
    assert abs(dz[-1] - 1.29) < 0.01 and abs(dz[0] - 111.7) < 0.1


def test_linear_z_faces(pkg):
    zf = pkg.linear_z_faces(10, 1000.0)
    assert abs(zf[0] + 1000.0) < 1e-9 and zf[-1] == 0.0
    assert abs((zf[-1] - zf[-2]) - 5.0) < 1e-12


def test_double_drake_bathymetry(pkg):
    f = pkg.double_drake_bathymetry
    assert f(0.5, 0.0) == 0.0 and f(90.5, 10.0) == 0.0
    assert f(0.5, -40.0) == -10000.0 and f(1.0, 0.0) == -10000.0 and f(45.0, 0.0) == -10000.0
    # C3: 7 920 of 1 944 000 columns solid
    g = pkg.LatitudeLongitudeGrid(2160, 900, 2, np.array([-1.0, -0.5, 0.0]))
    solid = f(g.lam_c()[None, :], g.phi_c()[:, None]) == 0.0
    assert int(solid.sum()) == 7920


def test_cubic_profiles_hit_their_knots(pkg):
    wc = pkg.wind_stress_coefficients(75.0)
    sc = pkg.salinity_flux_coefficients(75.0, 35.0)
    knots = [(-75, 0.0), (-45, 0.2), (-15, -0.1), (0, -0.02), (15, -0.1), (45, 0.1), (75, 0.0)]
    for phi, val in knots:
        for eps in (-1e-9, 1e-9):
            if abs(phi + eps) <= 75:
                assert abs(pkg.wind_stress(phi + eps, wc) - val / 1000) < 1e-9
    for phi, val in [(-75, -2e-8), (-20, 2e-8), (0, -4e-8), (20, 2e-8), (75, -2e-8)]:
        assert abs(pkg.salinity_flux(phi - 1e-9 if phi > 0 else phi + 1e-9, sc) - val * 35.0) < 1e-12


def test_barotropic_substeps_match_survey_appendix_b(pkg):
    # C3 (1/6 deg, 3000 m, dt = 540 s) -> N = 53..54 -> M = 38
    zf = pkg.exponential_z_faces(120, 3000)
    g = pkg.LatitudeLongitudeGrid(2160, 900, 120, zf)
    N = pkg.barotropic_substeps(540.0, g)
    assert N in (53, 54)
    fs = pkg.SplitExplicitFreeSurface(substeps=N)
    assert pkg.substeps(fs) == 38
    assert abs(np.sum(fs.averaging_weights) - 1.0) < 1e-14
    assert fs.averaging_weights[-1] > 0 and abs(fs.fractional_step_size - 2.0 / N) < 1e-15
    # minimum of 10 (near_global_simulation.jl:18)
    assert pkg.barotropic_substeps(1.0, g) == 10


def test_averaging_weights_are_normalised_and_centred_on_the_new_time_level(pkg):
    """Split-explicit averaging (Shchepetkin & McWilliams 2005 shape function): the weights sum to one and their first moment
    sum_m w_m m dtau sits at the new baroclinic time level (within the ~1 % the discrete truncation + renormalisation leaves),
    for every substep count; a wrong shape constant would move the centroid by tens of per cent."""
    for n in (10, 30, 38, 53, 80, 160):
        fs = pkg.SplitExplicitFreeSurface(substeps=n)
        w = np.asarray(fs.averaging_weights)
        m = np.arange(1, len(w) + 1)
        assert abs(w.sum() - 1.0) < 1e-14 and np.all(w[:-1] != 0.0)
        assert abs((w * m).sum() * fs.fractional_step_size - 1.0) < 0.02, n
        assert 1.3 < len(w) * fs.fractional_step_size < 1.5          # the substeps reach ~1.4 dt past the old time level


def test_calculate_local_N_reference_algorithm(pkg):
    # uniform load: greedy `<=` advance takes one more column than the mean per rank (grid_load_balance.jl:102-109)
    load = np.full(100, 10)
    assert pkg.calculate_local_N(load, 100, 4) == [26, 26, 26, 22]
    # ragged load with land
    load = np.array([0] * 10 + [5] * 20 + [1] * 30 + [8] * 40)
    ln = pkg.calculate_local_N(load, 100, 3)
    assert sum(ln) == 100 and all(n > 0 for n in ln)
    total = load.sum() / 3
    assert load[:ln[0]].sum() > total >= load[:ln[0] - 1].sum()
    # single rank
    assert pkg.calculate_local_N(load, 100, 1) == [100]


def test_redistribute_size(pkg):
    l = [1200, 1100, 900, 1120]
    pkg.redistribute_size_to_fulfill_memory_limitation(l, 1150)
    assert sum(l) == 4320 and max(l) <= 1150
    l = [10, 300, 300, 300]
    pkg.redistribute_size_to_fulfill_memory_limitation(l, 1150)
    assert sum(l) == 910 and min(l) >= 20
    l = [19, 300, 300]
    pkg.redistribute_size_to_fulfill_memory_limitation(l, 1150)
    assert l == [20, 299, 300]


def test_load_balanced_grid_partitions(pkg):
    zf = pkg.exponential_z_faces(10, 3000)
    grids = [pkg.load_balanced_grid(r, 4, (360, 150, 10), (-75, 75), zf, 1, False, "DoubleDrake") for r in range(4)]
    assert [g.Nx for g in grids] == [90] * 4 and [g.i_offset for g in grids] == [0, 90, 180, 270]
    # bottom height with halo columns wraps periodically and matches the neighbours' interiors
    b0, b3 = grids[0].bottom_height_local(), grids[3].bottom_height_local()
    assert b0.shape == (150, 90 + 14)
    assert np.array_equal(b0[:, :7], b3[:, -14:-7])

    def bathy(Nx, Ny):   # synthetic continents: land in the western half
        lam = -180 + (np.arange(Nx) + 0.5) * 360 / Nx
        return np.where(lam[None, :] < -60, 100.0, -4000.0) * np.ones((Ny, 1))
    gb = [pkg.load_balanced_grid(r, 4, (360, 150, 10), (-75, 75), zf, 1, True, "RealisticOcean",
                                 realistic_bathymetry=bathy) for r in range(4)]
    sizes = [g.Nx for g in gb]
    assert sum(sizes) == 360 and sizes[0] > sizes[1]     # the land-heavy slab is wider
    wet = [int((~g.immersed_cell_mask()[:, :, g.i_offset:g.i_offset + g.Nx]).sum()) for g in gb]
    assert max(wet[:3]) - min(wet[:3]) <= 2 * 150 * 10   # balanced to within ~one column each way


def test_abi_library_exports_every_declared_symbol(pkg):
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "ocean_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(ob200_[a-z_0-9]+)\s*\(", hdr)))
    assert declared == sorted(pkg._abi.EXPORTED_SYMBOLS)
    for precision, (lib, dtype) in pkg._abi.PRECISIONS.items():   # the Float64 library and its Float32 build: same entry points
        assert os.path.exists(lib), f"{os.path.basename(lib)} not built (run __graft_entry__.build())"
        out = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True, check=True).stdout
        exported = set(line.split()[-1] for line in out.splitlines() if line.strip())
        missing = [s for s in declared if s not in exported]
        assert not missing, (precision, missing)
        # loads without a GPU, and says which element type its field buffers have
        assert pkg._abi.load_library(precision=precision).ob200_real_bytes() == np.dtype(dtype).itemsize
    # the struct the host passes has the size the header's C compiler sees
    src = '#include "ocean_b200.h"\n#include <stdio.h>\nint main(){printf("%zu", sizeof(ob200_config));return 0;}\n'
    exe = "/tmp/ob200_sizeof"
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(root, "include"), "-o", exe], input=src, text=True, check=True)
    size = int(subprocess.run([exe], capture_output=True, text=True, check=True).stdout)
    assert size == C.sizeof(pkg._abi.ob200_config)


def test_product_fails_loudly_without_gpu(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import cases
    with pytest.raises(RuntimeError, match="no CUDA device|ob200_create failed"):
        cases.build_case("dd_small").make_model(pkg, None)


def test_cost_balanced_partition(pkg):
    import cases
    c = cases.Case("c3", 2160, 900, 4, ranks=8, balance="cost")
    g = c.grid(pkg, 0, 8)
    sizes = g.partition.sizes
    assert sum(sizes) == 2160 and min(sizes) >= 40
    cost = pkg.kernel_cost_per_column(g.bottom_height_global, 2160, 900)
    loads = [cost[g.partition.offset(r):g.partition.offset(r) + sizes[r]].sum() for r in range(8)]
    assert max(loads) / (sum(loads) / 8) < 1.02          # balanced to 2 %
    eq = pkg.Partition.equal(2160, 8)
    eq_loads = [cost[eq.offset(r):eq.offset(r) + 270].sum() for r in range(8)]
    assert max(loads) < max(eq_loads)                    # better than the equal split
    assert pkg.Partition.balanced(np.ones(100), 1).sizes == (100,)
    assert pkg.Partition.balanced(np.ones(100), 4).sizes == (25, 25, 25, 25)


def test_reference_smoke_test_through_the_mirror(pkg, orc):
    """The reference's own test (test/runtests.jl:11-19,33-34): scaling_test_simulation(1/5, ranks, 10minutes, 365days;
    Nz = 10, experiment) and 10 iterations must run -- here through the host mirror on the CPU oracle backend."""
    for experiment in ("Quiescent", "DoubleDrake"):
        sim = pkg.scaling_test_simulation(1 / 5, (1, 1, 1), 10 * 60.0, 365 * 86400.0, Nz=10, experiment=experiment,
                                          backend_factory=orc.OracleBackend)
        assert sim.model.grid.size() == (72, 30, 10)
        sim.stop_iteration = 10
        sim.progress = None
        # the oracle backend has no device diagnostics: min_dt == max_dt disables the wizard, as a scalar dt does
        pkg.run(sim)
        assert sim.model.clock["iteration"] == 10
        for n in ("U", "V", "T", "S", "ETA"):
            assert np.all(np.isfinite(sim.model.get(n)))
    # profile mode: warm-up + timed steps, returns seconds per step and writes time_res*.json (near_global_simulation.jl:178-216)
    import tempfile, os, glob
    with tempfile.TemporaryDirectory() as d:
        cwd = os.getcwd()
        os.chdir(d)
        try:
            t = pkg.scaling_test_simulation(1 / 5, (1, 1, 1), 600.0, 1e9, Nz=10, experiment="DoubleDrake", profile=True,
                                            backend_factory=orc.OracleBackend, profile_steps=(1, 1, 2))
            assert t > 0 and glob.glob("time_res*_NZ10_ranks1_precFloat64.json")
        finally:
            os.chdir(cwd)


def test_env_var_names_match_run_jl():
    """run.py parses the same environment variables as experiments/run.jl:30-42."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "oceanscalingtests.jl_b200", "run.py")).read()
    for var in ("RESOLUTION", "EXPERIMENT", "WITHFLUXES", "WITHRESTORING", "PROFILE", "RESTART", "NZ", "LOADBALANCE", "PRECISION"):
        assert f'"{var}"' in src, var


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU oracle on a bounded sample) runs without a GPU, prints exactly one JSON line
    with the contract's keys on rank 0, and nothing on the other ranks of a torchrun launch."""
    import json
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "small", "--steps", "2",
           "--warmup", "1", "--cpu-columns", "16"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, RANK="0", WORLD_SIZE="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "grid-cell-steps/s" and d["unit"] == "cell-steps/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["steps"] == 2
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "cell-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    r1 = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert r1.returncode == 0 and r1.stdout.strip() == ""
