"""Real multi-GPU runs (one process per GPU over NVLink peer memory and NCCL), launched with torchrun from inside pytest.
Skipped when fewer than two GPUs are visible; the single-GPU emulation of the same paths is tests/test_peer_lockstep_gpu.py."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _torchrun(nproc, script, *args, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + os.getpid() % 500), str(ROOT / "tools" / script), *[str(a) for a in args]]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=str(ROOT))
    assert res.returncode == 0, res.stdout[-4000:] + "\n" + res.stderr[-4000:]
    return res.stdout


@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_peer_paths_vs_oracle(nproc):
    if _n_gpus() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    out = _torchrun(nproc, "dist_check_peer.py", 300_000, 300_000, 50_000, 6)
    assert out.count("OK") >= 3 * nproc and "MISMATCH" not in out


@pytest.mark.parametrize("nproc", [2])
def test_nccl_paths_vs_oracle(nproc):
    if _n_gpus() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    out = _torchrun(nproc, "dist_check_lean.py", 300_000)
    assert "MISMATCH" not in out and out.count("OK") >= nproc
    out = _torchrun(nproc, "dist_check_icp.py", 300_000, 50_000, 6)
    assert "MISMATCH" not in out and out.count("OK") >= nproc
