"""x-slab runs on 2+ GPUs must equal the single-GPU run bit for bit (skipped unless >= 2 GPUs are visible)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("case_name,precision", [("dd_mid", "Float64"), ("bumpy", "Float64"), ("dd_mid", "Float32")])
def test_slabs_match_single_gpu(case_name, precision):
    """Float32: the slab run of libocean_b200_f32.so (ncclFloat exchanges) against its own single-GPU run -- the same
    operations in the same order, so also bit for bit."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29731", os.path.join(HERE, "mgpu_worker.py"), case_name, "6"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, OB200_TEST_PRECISION=precision))
    sys.stdout.write(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


def test_slab_checkpoint_restart_matches_single_gpu():
    """2 slabs: 3 steps, per-rank checkpoint, fresh handles restored from the files, 3 more steps == 6 steps on 1 GPU."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29733", os.path.join(HERE, "mgpu_worker.py"), "dd_mid", "6", "restart"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


def test_overlapped_exchange_option_matches_single_gpu():
    """Option "overlap_exchange": NCCL halo exchange on a second stream beside the interior column sweeps; same bits."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29735", os.path.join(HERE, "mgpu_worker.py"), "bumpy", "6"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, OB200_OVERLAP_EXCHANGE="1"))
    sys.stdout.write(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("case_name,env", [("dd_mid", {"OB200_BARO_EXCHANGE": "substep"}), ("dd_narrow", {}),
                                           ("dd_mid", {"OB200_BARO_EXCHANGE": "substep", "OB200_BARO_GRAPH": "0"}),
                                           ("dd_mid", {"OB200_ONCHIP_BARO": "1"}), ("bumpy", {"OB200_ONCHIP_BARO": "1"})])
def test_per_substep_barotropic_exchange_matches_single_gpu(case_name, env):
    """The other barotropic strategies on x-slabs: one-column eta / U exchanges every half substep (BASELINE configs[4]'s
    second strategy; the only one possible on slabs narrower than the wide halo, e.g. the reference's 20-column clamp) and
    the persistent on-chip kernel over the wide-halo slab: same bits as one GPU."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29737", os.path.join(HERE, "mgpu_worker.py"), case_name, "6"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, **env))
    sys.stdout.write(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
