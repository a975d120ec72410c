#!/usr/bin/env python3
"""BASELINE.json configs[4]: split-explicit barotropic substep + halo-exchange microbenchmark at 1/6 degree.

The number of substeps M actually run per baroclinic step is swept over {10, 20, 38, 51, 80, 160}; for every M the
barotropic stage (wide exchange / per-substep exchanges + the substep loop) of one baroclinic step is timed with CUDA
events on the library's stream, max over ranks, for every strategy the library has:

  1 GPU     launch_loop | cuda_graph | persistent (cooperative, global memory) | onchip (cooperative, state in shared
            memory; only when the domain fits the aggregate shared memory, else it falls back to the graph)
  x-slabs   wide_graph | wide_loop | wide_persistent | wide_onchip
                                        halo of M + 1 columns exchanged once, no communication inside the loop
            substep_graph | substep_loop  one-column eta / U exchanges (NCCL send/recv) every half substep

    python bench_barotropic.py                                   # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 ... bench_barotropic.py

One JSON line per M on stdout (rank 0).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
import cases  # noqa: E402


def substeps_for(pkg, M):
    """Smallest `substeps` argument whose averaging weights have exactly M entries (M ~ 0.71 N)."""
    for N in range(M, 4 * M + 8):
        if len(pkg.SplitExplicitFreeSurface(substeps=N).averaging_weights) == M:
            return N
    raise ValueError(M)


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    saved = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = ge.package()
    Nx, Ny, Nz, dt = 2160, 900, 4, 540.0      # the 2-D problem of C3; a few levels only (the 3-D stages are not timed)
    steps = int(os.environ.get("OB200_BARO_STEPS", "10"))
    if world == 1:
        modes = {"launch_loop": ("auto", 0, 0), "cuda_graph": ("auto", 1, 0), "persistent": ("auto", 0, 1),
                 "onchip": ("auto", 0, 2)}
    else:
        modes = {"wide_graph": ("wide", 1, 0), "wide_loop": ("wide", 0, 0), "wide_persistent": ("wide", 0, 1),
                 "wide_onchip": ("wide", 0, 2), "substep_graph": ("substep", 1, 0), "substep_loop": ("substep", 0, 0)}
    lines = []
    for M in (10, 20, 38, 51, 80, 160):
        case = cases.Case("baro", Nx, Ny, Nz, experiment="DoubleDrake", dt=dt, ranks=world)
        grid = case.grid(pkg, rank, world)
        fs = pkg.SplitExplicitFreeSurface(substeps=substeps_for(pkg, M))
        res = {}
        for mode, (exchange, graph, persistent) in modes.items():
            if exchange == "wide" and grid.Nx < M + 1:
                res[mode] = None
                continue
            model = pkg.HydrostaticFreeSurfaceModel(grid=grid, free_surface=fs, eos="linear",
                                                    boundary_conditions=pkg.set_boundary_conditions("DoubleDrake", grid),
                                                    device=local, comm=pkg.Comm(), barotropic_exchange=exchange)
            case.initialize(pkg, model)
            model.backend.set_option("barotropic_graph", graph)
            model.backend.set_option("persistent_barotropic", int(persistent == 1))
            model.backend.set_option("onchip_barotropic", int(persistent == 2))      # falls back to the graph when it does not fit
            model.backend.set_option("timing", 1)
            for _ in range(3):
                pkg.time_step(model, dt)
            acc = 0.0
            for _ in range(steps):
                pkg.time_step(model, dt)
                ms, _ = model.backend.timing()
                acc += ms[1]
            t = torch.tensor([acc / steps], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            res[mode] = round(float(t.item()), 4)
            model.backend.close()
        best = min((v, k) for k, v in res.items() if v is not None)[1]
        lines.append({"metric": "barotropic stage ms per baroclinic step (max over ranks)", "n_gpus": world,
                      "columns": Nx * Ny, "columns_per_gpu": grid.Nx * Ny, "substeps_run": M, "substeps_requested": fs.substeps,
                      **res, "fastest": best})
    if rank == 0:
        os.dup2(saved, 1)
        for ln in lines:
            print(json.dumps(ln), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
