"""The multi-GPU fabric on REAL peer memory (CUDA IPC between processes, NVLink): needs >= 2 GPUs, skipped
otherwise (tests/test_gpu_fabric.py covers the same kernels with several ranks on one GPU)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    from dpgo_ros_b200 import capi
    return capi.lib().dpgo_b200_device_count()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("mode", ["sync", "parallel"])
def test_fabric_across_processes_matches_single_team(mode):
    n = _gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "_fabric_worker.py"), mode]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert f"fabric {mode} ok" in out.stdout


def test_per_robot_api_across_processes_through_shared_memory():
    """The e2e arm at N > 1 (dist.ShmHostTeam): stand-alone agents in several processes, host buffers, shared-memory
    mailboxes.  Two processes also work on a single GPU (no device-side waiting between them)."""
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "_fabric_worker.py"), "shm"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "fabric shm ok" in out.stdout
