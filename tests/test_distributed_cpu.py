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
