"""Host-side callers around `time_step!` driven through the host mirror with the CPU oracle as
the backend: checkpoint -> restore -> continue, the surface writer, time-aligned schedules and the 5-day flux reload.
The same scenarios run against libocean_b200.so in tests/test_parity_gpu.py."""
import os

import numpy as np
import pytest

from cases import build_case

days = 86400.0


def _prognostic(model):
    return {n: model.get(n) for n in ("U", "V", "W", "T", "S", "ETA", "GU", "GV", "GT", "GS", "KAPPA_C")}


def checkpoint_restart_roundtrip(pkg, factory, tmp_path, case_name="dd_small", n1=4, n2=4):
    """n1 steps, checkpoint, n2 more steps; a second model restored from the file must continue bit for bit."""
    case = build_case(case_name)
    m1 = case.make_model(pkg, factory)
    for _ in range(n1):
        pkg.time_step(m1, case.dt)
    ck = pkg.Checkpointer(str(tmp_path), f"{case.experiment}_checkpoint_0", pkg.TimeInterval(5 * days))
    path = ck.write(m1)
    assert os.path.basename(path) == f"{case.experiment}_checkpoint_0_iteration{n1}.npz"
    for _ in range(n2):
        pkg.time_step(m1, case.dt)
    ref = _prognostic(m1)

    m2 = case.make_model(pkg, factory, initialize=False)
    pkg.set_from_checkpoint(m2, path)
    assert m2.clock["iteration"] == n1 and m2.clock["time"] == m1.clock["time"] - n2 * case.dt
    for _ in range(n2):
        pkg.time_step(m2, case.dt)
    got = _prognostic(m2)
    for n in ref:
        assert np.array_equal(ref[n], got[n]), f"{n} differs after restart"
    return path


def test_checkpoint_layout_and_bitwise_restart(pkg, orc, tmp_path):
    path = checkpoint_restart_roundtrip(pkg, orc.OracleBackend, tmp_path)
    case = build_case("dd_small")
    with np.load(path) as f:
        keys = set(f.files)
        # the reference's group names (simulation_outputs.jl:54,73,90,96,62)
        for name in "uvTS":
            assert f"{name}/data" in keys
            assert f"timestepper/Gⁿ/{name}/data" in keys and f"timestepper/G⁻/{name}/data" in keys
        assert "η/data" in keys and "clock/time" in keys and "clock/iteration" in keys
        # parent arrays: interior + 7-cell halos, v with Ny + 1 rows, η windowed (singleton k)
        assert f["u/data"].shape == (case.Nz + 14, case.Ny + 14, case.Nx + 14)
        assert f["v/data"].shape == (case.Nz + 14, case.Ny + 1 + 14, case.Nx + 14)
        assert f["η/data"].shape == (1, case.Ny + 14, case.Nx + 14)
        assert int(f["clock/iteration"]) == 4


def test_restore_accepts_other_halo_widths_and_reference_only_keys(pkg, orc, tmp_path):
    """load_free_surface! strips whatever halo the file has (distributed runs widen η's); a checkpoint holding only the
    reference's keys restores state, tendencies and clock (first step after it is an Euler step, like a fresh Δt)."""
    case = build_case("dd_small")
    m1 = case.make_model(pkg, orc.OracleBackend)
    for _ in range(3):
        pkg.time_step(m1, case.dt)
    path = pkg.Checkpointer(str(tmp_path), "ck", pkg.TimeInterval(days), extended=False).write(m1)
    with np.load(path) as f:
        entries = {k: f[k] for k in f.files}
    assert not any(k.startswith("b200/") for k in entries)
    eta = entries["η/data"]
    wide = np.zeros((1, eta.shape[1] + 2 * 20, eta.shape[2] + 2 * 33))
    wide[:, 20:-20, 33:-33] = eta
    entries["η/data"] = wide
    path2 = str(tmp_path / "wide.npz")
    pkg.GroupFile(path2, "w").write(entries)
    m2 = case.make_model(pkg, orc.OracleBackend, initialize=False)
    pkg.set_from_checkpoint(m2, path2)
    for n in ("U", "V", "T", "S", "ETA", "GU", "GT", "GU_PREV", "GS_PREV"):
        assert np.array_equal(m1.get(n), m2.get(n)), n
    assert m2.clock["iteration"] == 3 and m2.clock["time"] == 3 * case.dt


def test_surface_writer_groups(pkg, orc, tmp_path):
    case = build_case("dd_small")
    m = case.make_model(pkg, orc.OracleBackend)
    sim = pkg.Simulation(m, dt=case.dt, stop_time=3 * case.dt)
    pkg.set_outputs(sim, "DoubleDrake", surface_time=2 * case.dt, checkpoint_time=3 * case.dt, output_dir=str(tmp_path))
    pkg.run(sim)
    with np.load(tmp_path / "DoubleDrake_fields_0.npz") as f:
        its = sorted(int(k.split("/")[-1]) for k in f.files if k.startswith("timeseries/t/"))
        assert its == [0, 2]                       # initial state + one TimeInterval actuation
        assert f["timeseries/t/2"] == 2 * case.dt
        top = f["timeseries/T/2"]
        assert top.shape == (1, case.Ny + 14, case.Nx + 14)
        # the state at the time of the write is not kept; compare the last write with a re-run
        assert np.all(np.isfinite(f["timeseries/ζ/2"])) and np.abs(f["timeseries/ζ/2"]).max() > 0
        assert np.abs(f["timeseries/w/2"]).max() > 0
    assert os.path.exists(tmp_path / "DoubleDrake_checkpoint_0_iteration0.npz")
    assert os.path.exists(tmp_path / "DoubleDrake_checkpoint_0_iteration3.npz")
    # Quiescent: set_outputs! is a no-op (simulation_outputs.jl:43)
    sim_q = pkg.Simulation(m, dt=case.dt, stop_time=0.0)
    pkg.set_outputs(sim_q, "Quiescent")
    assert sim_q.output_writers == {}


def test_time_steps_align_with_schedules(pkg, orc):
    """A TimeInterval schedule shortens the step that would overshoot it (Oceananigans aligned_time_step)."""
    case = build_case("dd_small")
    m = case.make_model(pkg, orc.OracleBackend)
    fired = []
    sim = pkg.Simulation(m, dt=case.dt, stop_time=2.5 * case.dt + 7.0)
    sim.callbacks["probe"] = pkg.Callback(lambda s: fired.append(s.model.clock["time"]), pkg.TimeInterval(1.5 * case.dt))
    pkg.run(sim)
    # initial call, then exactly at 1.5 dt (after a half step); the run ends exactly on stop_time
    assert fired == [0.0, 1.5 * case.dt]
    assert m.clock["time"] == 2.5 * case.dt + 7.0
    assert m.clock["iteration"] == 4       # dt, dt/2, dt, 7 s


def test_flux_file_numbers_follow_the_reference(pkg):
    f, r = pkg.flux_filenum, pkg.restoring_filenum
    assert [f(d * days, True) for d in (0, 4.99, 5, 364, 365, 370)] == [1, 1, 2, 73, 1, 2]
    assert [f(d * days, False) for d in (0, 5, 365, 370)] == [1, 2, 74, 75]
    assert [r(d * days, True) for d in (0, 1, 15, 16, 30, 365)] == [1, None, 2, None, 3, 1]
    assert [r(d * days, False) for d in (0, 14, 15, 31)] == [1, 1, 2, 3]


def flux_reload_run(pkg, factory, steps=6):
    """RealisticOcean stepping across the day-5 boundary: update_fluxes! swaps in file 2 exactly at t = 5 days."""
    case = build_case("realistic")
    grid = case.grid(pkg)
    loads = []

    def fluxes(g, n):
        loads.append(n)
        return pkg.synthetic.flux_file(g, n)

    bcs = pkg.set_boundary_conditions("RealisticOcean", grid, forcing_arrays=pkg.synthetic.forcing_arrays(grid),
                                      providers={"fluxes": fluxes, "restoring": pkg.synthetic.restoring_file})
    nsub = pkg.barotropic_substeps(3240.0 / (case.Nx / 360.0), grid)
    m = pkg.HydrostaticFreeSurfaceModel(grid=grid, free_surface=pkg.SplitExplicitFreeSurface(substeps=nsub), eos="teos10",
                                        boundary_conditions=bcs, backend_factory=factory)
    case.initialize(pkg, m)
    t0 = 5 * days - 2.5 * case.dt
    m.backend.set_clock(t0, 0, float("inf"))
    sim = pkg.Simulation(m, dt=case.dt, stop_time=t0 + 1e9, stop_iteration=steps)
    sim.callbacks["update_fluxes"] = pkg.Callback(pkg.update_fluxes, pkg.TimeInterval(5 * days))
    sim.callbacks["update_restoring"] = pkg.Callback(pkg.update_restoring, pkg.TimeInterval(1 * days))
    pkg.run(sim)
    return m, loads


def test_flux_reload_callback_fires_on_the_five_day_boundary(pkg, orc):
    m, loads = flux_reload_run(pkg, orc.OracleBackend)
    assert loads == [1, 2]                               # initialisation (file 1), then file 2 at day 5
    # steps: dt, dt, dt/2 (aligned to day 5), then full steps
    assert m.clock["iteration"] == 6
    assert abs(m.clock["time"] - (5 * days + 3 * 600.0)) < 1e-6
    taux2 = pkg.synthetic.flux_file(m.grid, 2)["taux"]
    assert np.array_equal(m.boundary_conditions.forcing_arrays["taux"], taux2)
    assert np.all(np.isfinite(m.get("T")))


def test_partition_global_array(pkg):
    case = build_case("dd_small")
    g = case.grid(pkg, rank=1, ranks=3)
    a = np.arange(6 * case.Ny * case.Nx, dtype=np.float64).reshape(6, case.Ny, case.Nx)
    loc = pkg.partition_global_array(g, a)
    assert loc.shape == (6, case.Ny, g.Nx) and np.array_equal(loc, a[..., g.i_offset:g.i_offset + g.Nx])
    assert pkg.partition_global_array(g, loc) is not None
    with pytest.raises(ValueError):
        pkg.partition_global_array(g, a[..., :7])


def test_time_step_wizard_follows_the_cfl(pkg, orc):
    """TimeStepWizard(cfl=0.35; max_change=1.1, max_Δt, min_Δt) every 10 iterations (near_global_simulation.jl:160-162)
    driven through `run` on the oracle: Δt grows by at most 10 % per update, never leaves [min_Δt, max_Δt], and a
    changed Δt is followed by a forward-Euler step (last_dt bookkeeping)."""
    case = build_case("dd_small")
    m = case.make_model(pkg, orc.OracleBackend)
    m.backend.grid_for_diagnostics = m.grid
    sim = pkg.Simulation(m, dt=5.0, stop_time=1e9, max_dt=300.0, min_dt=5.0, stop_iteration=35)
    seen = []
    sim.callbacks["probe"] = pkg.Callback(lambda s: seen.append(s.dt), pkg.IterationInterval(1))
    pkg.run(sim)
    assert m.clock["iteration"] == 35
    changes = [b / a for a, b in zip(seen[:-1], seen[1:]) if b != a]
    assert len(changes) == 3 and all(abs(c - 1.1) < 1e-12 for c in changes)     # iterations 10, 20, 30 (and 0: 5 -> 5.5)
    assert all(5.0 <= d <= 300.0 for d in seen)
    _, cfl_dt = m.backend.diagnostics()
    assert 0.35 * cfl_dt > seen[-1]            # still CFL-limited by max_change, not by the flow
