"""Multi-GPU parity worker, launched by torchrun (one rank per GPU): the x-slab run through libocean_b200.so
(NCCL halo exchange + wide-halo barotropic solver) must reproduce the single-GPU run of the same case."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import __graft_entry__ as ge  # noqa: E402
import cases  # noqa: E402


def main():
    case_name, nsteps = sys.argv[1], int(sys.argv[2])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = ge.package()
    c = cases.build_case(case_name)
    precision = os.environ.get("OB200_TEST_PRECISION", "Float64")     # Float32: libocean_b200_f32.so, ncclFloat exchanges
    m = c.make_model(pkg, None, rank=rank, ranks=world, device=local, comm=pkg.Comm(), precision=precision)
    if os.environ.get("OB200_OVERLAP_EXCHANGE") == "1":
        # halo exchange on a second stream beside the interior columns of the auxiliary sweeps (option, same bits)
        m.backend.set_option("overlap_exchange", 1)
    if os.environ.get("OB200_ONCHIP_BARO") == "1":
        m.backend.set_option("onchip_barotropic", 1)
    if os.environ.get("OB200_BARO_GRAPH") == "0":
        m.backend.set_option("barotropic_graph", 0)
    restart = len(sys.argv) > 3 and sys.argv[3] == "restart"
    if restart:
        # per-rank checkpoint half way (simulation_outputs.jl:36-40 names the files by rank), then a fresh handle
        # restored from it finishes the run: the slab halos of the state Oceananigans omits travel in the file
        import tempfile
        half = nsteps // 2
        for _ in range(half):
            pkg.time_step(m, c.dt)
        d = tempfile.mkdtemp(prefix=f"ob200_ck_{rank}_")
        path = pkg.Checkpointer(d, f"DoubleDrake_checkpoint_{rank}", pkg.TimeInterval(86400.0)).write(m)
        m.backend.close()
        m = c.make_model(pkg, None, rank=rank, ranks=world, device=local, comm=pkg.Comm(), initialize=False, precision=precision)
        pkg.set_from_checkpoint(m, path)
        for _ in range(nsteps - half):
            pkg.time_step(m, c.dt)
    else:
        for _ in range(nsteps):
            pkg.time_step(m, c.dt)
    m.backend.sync()
    names = ("U", "V", "W", "ETA", "T", "S", "KAPPA_U", "GU", "GT")
    ok = True
    for n in names:
        parts = [None] * world
        dist.all_gather_object(parts, m.get(n))
        if rank == 0:
            full = np.concatenate(parts, axis=-1)
            if n == names[0]:
                ref = c.make_model(pkg, None, device=local, precision=precision)
                for _ in range(nsteps):
                    pkg.time_step(ref, c.dt)
            same = np.array_equal(full, ref.get(n))
            err = float(np.max(np.abs(full - ref.get(n))))
            print(f"mgpu {case_name} x{world} {precision} {n}: identical={same} max|diff|={err:.3e}", flush=True)
            ok = ok and same
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
