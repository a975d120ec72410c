"""N > 1 host logic on CPU: world_size-2 (and 3) gloo runs of the oracle on x-slabs must reproduce the
single-domain run bit for bit (partitioning, halo indexing, ring neighbours, substep exchange schedule)."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, case_name, nsteps, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import importlib.util
    import __graft_entry__ as ge
    import cases
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg, orc = ge.package(), ge.oracle_module()
    spec = importlib.util.spec_from_file_location("oracle_distributed", os.path.join(ROOT, "oracle", "distributed.py"))
    od = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(od)
    c = cases.build_case(case_name)
    model = c.make_model(pkg, orc.OracleBackend, rank=rank, ranks=world, initialize=False)
    stepper = od.SlabStepper(model)
    # set!(model, ...) ends with update_state!: in slab mode that needs the exchange, so initialise field by field
    model.backend.update_state = stepper.update_state
    c.initialize(pkg, model)
    for _ in range(nsteps):
        stepper.time_step(c.dt)
    out = {n: od.gather_global(model, n, None) for n in ("U", "V", "W", "ETA", "T", "S", "KAPPA_U")}
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,case_name", [(2, "dd_small"), (3, "bumpy")])
def test_slab_oracle_matches_single_domain(pkg, orc, world, case_name):
    import cases
    nsteps = 4
    c = cases.build_case(case_name)
    ref = c.make_model(pkg, orc.OracleBackend)
    for _ in range(nsteps):
        pkg.time_step(ref, c.dt)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, case_name, nsteps, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for n, a in out.items():
        assert np.array_equal(a, ref.get(n)), f"{n}: slab run differs from the single-domain run"


def test_partition_offsets_and_ring(pkg):
    p = pkg.Partition.equal(90, 4)
    assert p.sizes == (23, 23, 22, 22) and [p.offset(r) for r in range(4)] == [0, 23, 46, 68]
    import cases
    g = cases.build_case("dd_small").grid(pkg, rank=1, ranks=2)
    assert g.Nx == 45 and g.i_offset == 45 and g.nranks == 2
    # the wide barotropic halo requested for x-slabs: M + 1 columns of bathymetry on each side
    assert g.bottom_height_local(37).shape == (40, 45 + 74)
    lam = g.lam_c()
    assert abs(lam[0] - (-180 + 45.5 * 4.0)) < 1e-12


def _restart_worker(rank, world, port, case_name, n1, n2, tmpdir, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import importlib.util
    import __graft_entry__ as ge
    import cases
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg, orc = ge.package(), ge.oracle_module()
    spec = importlib.util.spec_from_file_location("oracle_distributed", os.path.join(ROOT, "oracle", "distributed.py"))
    od = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(od)
    c = cases.build_case(case_name)

    def slab_model(initialize):
        m = c.make_model(pkg, orc.OracleBackend, rank=rank, ranks=world, initialize=False)
        st = od.SlabStepper(m)
        m.backend.update_state = st.update_state        # update_state! on a slab needs the neighbour exchange
        if initialize:
            c.initialize(pkg, m)
        return m, st

    m1, s1 = slab_model(True)
    for _ in range(n1):
        s1.time_step(c.dt)
    path = pkg.Checkpointer(tmpdir, f"DoubleDrake_checkpoint_{rank}", pkg.TimeInterval(86400.0)).write(m1)
    for _ in range(n2):
        s1.time_step(c.dt)
    m2, s2 = slab_model(False)
    pkg.set_from_checkpoint(m2, path)                   # per-rank file, as simulation_outputs.jl:36-40 names them
    for _ in range(n2):
        s2.time_step(c.dt)
    same = all(np.array_equal(m1.get(n), m2.get(n)) for n in ("U", "V", "W", "ETA", "T", "S", "KAPPA_C", "KAPPA_U", "GU", "GT"))
    q.put((rank, bool(same), os.path.basename(path)))
    dist.barrier()
    dist.destroy_process_group()


def test_slab_checkpoint_restart_is_bit_reproducible(pkg, orc, tmp_path):
    """2 x-slabs of the oracle over gloo: checkpoint per rank, restore into fresh slab models, continue -- the halo
    columns of the state that update_state! does not recompute travel in each rank's file."""
    world, n1, n2 = 2, 3, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 7
    procs = [ctx.Process(target=_restart_worker, args=(r, world, port, "dd_small", n1, n2, str(tmp_path), q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got == [(0, True, "DoubleDrake_checkpoint_0_iteration3.npz"), (1, True, "DoubleDrake_checkpoint_1_iteration3.npz")]


def _wizard_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import __graft_entry__ as ge
    import cases
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg, orc = ge.package(), ge.oracle_module()
    c = cases.build_case("dd_small")
    m = c.make_model(pkg, orc.OracleBackend, rank=rank, ranks=world, initialize=False)
    m.backend.grid_for_diagnostics = m.grid
    # asymmetric flow: the CFL limit binds on rank 1 only (set fields directly: update_state! on a slab needs the exchange)
    m.set("U", 0.05 if rank == 0 else 100.0)
    sim = pkg.Simulation(m, dt=1000.0, stop_time=1e9, max_dt=5000.0, min_dt=5.0, comm=pkg.Comm())
    _, local_cfl = m.backend.diagnostics()
    pkg.near_global_simulation.time_step_wizard(sim)
    q.put((rank, float(local_cfl), float(sim.dt)))
    dist.barrier()
    dist.destroy_process_group()


def test_time_step_wizard_agrees_across_ranks(pkg, orc):
    """the wizard's advective time scale is min-reduced over the x-slabs, so every rank takes the same dt."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 11
    procs = [ctx.Process(target=_wizard_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    (_, cfl0, dt0), (_, cfl1, dt1) = got
    assert cfl1 < 0.1 * cfl0                      # the local time scales differ by more than 10x ...
    assert dt0 == dt1                             # ... but both ranks take the same step,
    assert dt0 == max(0.5 * 1000.0, min(1.1 * 1000.0, 0.35 * cfl1)) and dt0 < 1000.0   # the tighter rank dictates it
