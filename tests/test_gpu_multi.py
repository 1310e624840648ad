"""Multi-GPU parity (needs >= 2 CUDA devices on the box): latitude bands + NCCL halo rows vs the oracle."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpu():
    import torch

    return torch.cuda.device_count()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("nt,nphi,steps,prec,halo", [(64, 128, 120, 0, "nccl"), (64, 128, 140, 2, "nccl"),
                                                     (61, 50, 30, 1, "nccl"), (256, 1024, 8, 0, "nccl"),
                                                     (64, 128, 120, 0, "peer"), (256, 1024, 60, 0, "peer"),
                                                     (16, 256, 40, 0, "peer"),
                                                     (64, 128, 90, 0, "peer_mix"), (1300, 256, 30, 0, "peer_mix"),
                                                     (1300, 256, 20, 0, "peer_reset"),
                                                     (64, 128, 120, 0, "host"), (1300, 256, 12, 0, "host"),
                                                     (64, 128, 40, 0, "foam"), (32, 64, 30, 1, "foam"),
                                                     (64, 128, 120, 1, "peer"), (256, 1024, 40, 1, "peer"),
                                                     (64, 128, 12, 0, "wgsl_nccl"), (64, 128, 12, 0, "wgsl_peer"),
                                                     (256, 1024, 30, 0, "wgsl_peer"),
                                                     # tall bands: the NCCL path runs boundary bands + exchange beside the interior
                                                     (1536, 256, 10, 0, "nccl"), (1536, 256, 10, 1, "nccl"), (1030, 100, 6, 2, "nccl")])
def test_sharded_matches_oracle(world, nt, nphi, steps, prec, halo):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    only = os.environ.get("SBX_TEST_WORLDS")  # e.g. "8": spend a big box's minutes on the sizes it alone can run
    if only and str(world) not in only.split(","):
        pytest.skip(f"SBX_TEST_WORLDS={only}")
    if nt < 2 * world:
        pytest.skip("fewer than two rows per band")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(HERE, "mgpu_worker.py"), str(nt), str(nphi), str(steps), str(prec), halo]
    env = dict(os.environ)
    if halo in ("host", "peer_mix"):
        env["SBX_HOST_CHUNK_ROWS"] = "56"  # several chunks per band: first / last chunk wait for the halo rows
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    sys.stdout.write(r.stdout[-2000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "ok=True" in r.stdout
