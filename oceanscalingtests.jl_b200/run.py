"""Environment-variable driver mirroring /root/reference/experiments/run.jl (:30-65).

    EXPERIMENT=DoubleDrake RESOLUTION=3 NZ=120 PROFILE=1 python -m torch.distributed.run ... run.py
    (or plain `python oceanscalingtests.jl_b200/run.py` on one GPU)

Same variables and defaults as the reference: RESOLUTION (3), EXPERIMENT (DoubleDrake),
NZ (100), PROFILE (1), LOADBALANCE (0), WITHFLUXES (0), WITHRESTORING (0), RESTART (""), OUTPUTDIR (./).
PRECISION must be Float64 (the only precision built).  MPI.Init is replaced by torch.distributed (one rank per
GPU, NCCL); RealisticOcean inputs are the synthetic stand-ins of synthetic.py (no network for the real files).
"""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _bool(name, default):
    return os.environ.get(name, default).strip().lower() in ("1", "true", "yes")


def main():
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    pkg = ge.package()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ranks = (world, 1, 1)                                   # run.jl:22-27: Rx = Nranks, Ry = 1

    resolution = int(os.environ.get("RESOLUTION", "3"))
    experiment = os.environ.get("EXPERIMENT", "DoubleDrake")
    with_fluxes = _bool("WITHFLUXES", "0")
    with_restoring = _bool("WITHRESTORING", "0")
    profile = _bool("PROFILE", "1")
    restart = os.environ.get("RESTART", "")
    Nz = int(os.environ.get("NZ", "100"))
    loadbalance = _bool("LOADBALANCE", "0")
    if os.environ.get("PRECISION", "Float64") != "Float64":
        raise SystemExit("only PRECISION=Float64 is built (SURVEY 8f rank 4)")

    max_dt = 45 * 48 * 1.5 / resolution                     # run.jl:44
    min_dt = 5
    stop_time = 7300 * 86400.0
    if rank == 0:
        print(f"[ Info: Scaling test ranks={ranks} resolution={resolution} max_dt={max_dt} min_dt={min_dt} "
              f"experiment={experiment} profile={profile}", flush=True)
    realistic = None
    if experiment == "RealisticOcean":
        depth = pkg.experiment_depth(experiment)
        realistic = {"bathymetry": lambda Nx, Ny: pkg.synthetic.realistic_bathymetry(Nx, Ny, depth),
                     "forcing": pkg.synthetic.forcing_arrays, "initial": pkg.synthetic.initial_T_S,
                     "fluxes": pkg.synthetic.flux_file, "restoring": pkg.synthetic.restoring_file}
    sim = pkg.scaling_test_simulation(resolution, ranks, (min_dt, max_dt), stop_time, device=local, Nz=Nz,
                                      experiment=experiment, restart=restart, profile=profile, with_fluxes=with_fluxes,
                                      with_restoring=with_restoring, loadbalance=loadbalance, realistic_inputs=realistic)
    if profile:
        if rank == 0:
            print(f"[ Info: mean wall time per step: {sim:.6f} s", flush=True)
    else:
        stop_it = os.environ.get("STOP_ITERATION")
        if stop_it:
            sim.stop_iteration = int(stop_it)
        output_dir = os.environ.get("OUTPUTDIR", "./")       # run.jl:39,57-60
        pkg.set_outputs(sim, experiment, output_dir=output_dir, overwrite_existing=False, checkpoint_time=10 * 86400.0)
        pickup = os.path.join(output_dir, f"RealisticOcean_checkpoint_{rank}_iteration{restart}.npz") if restart else False
        pkg.run(sim, pickup=pickup)
        if rank == 0:
            print(f"[ Info: simulation took {sim.run_wall_time:.2f} s", flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
