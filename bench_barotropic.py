#!/usr/bin/env python3
"""BASELINE.json configs[4]: split-explicit barotropic substep microbenchmark at 1/6 degree -- the number of substeps
per baroclinic step is swept, on 1 GPU or (under torchrun) on x-slabs with the wide-halo exchange.  Prints one JSON line
per substep count: ms per baroclinic step spent in the barotropic stage (CUDA events on the library's stream, max over
ranks) for the launch loop / CUDA-graph replay / persistent cooperative kernel.

    python bench_barotropic.py                       # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 ... bench_barotropic.py
"""
import json
import os
import sys


ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
import cases  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    saved = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = ge.package()
    Nx, Ny, Nz, dt = 2160, 900, 4, 540.0      # the 2-D problem of C3; a few levels only (the 3-D stages are not timed)
    lines = []
    for nsub in (10, 20, 38, 53, 80, 160):
        if world > 1 and Nx // world < nsub + 1:
            continue
        case = cases.Case("baro", Nx, Ny, Nz, experiment="DoubleDrake", dt=dt, ranks=world)
        grid = case.grid(pkg, rank, world)
        fs = pkg.SplitExplicitFreeSurface(substeps=nsub)
        res = {}
        for mode in ("launch_loop", "cuda_graph", "persistent"):
            model = pkg.HydrostaticFreeSurfaceModel(grid=grid, free_surface=fs, eos="linear",
                                                    boundary_conditions=pkg.set_boundary_conditions("DoubleDrake", grid),
                                                    device=local, comm=pkg.Comm())
            case.initialize(pkg, model)
            model.backend.set_option("barotropic_graph", int(mode == "cuda_graph"))
            model.backend.set_option("persistent_barotropic", int(mode == "persistent"))
            model.backend.set_option("timing", 1)
            for _ in range(3):
                pkg.time_step(model, dt)
            acc = 0.0
            for _ in range(10):
                pkg.time_step(model, dt)
                ms, _ = model.backend.timing()
                acc += ms[1]
            t = torch.tensor([acc / 10], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            res[mode] = float(t.item())
            model.backend.close()
        lines.append({"metric": "barotropic stage ms per baroclinic step", "n_gpus": world, "columns": Nx * Ny,
                      "substeps_requested": nsub, "substeps_run": len(fs.averaging_weights), **res})
    if rank == 0:
        os.dup2(saved, 1)
        for ln in lines:
            print(json.dumps(ln), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
