"""GPU, >= 2 devices: the NCCL row-partitioned path (skipped on a 1-GPU box)."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("tail_rows", ["0", "131072", "600"])
@pytest.mark.parametrize("world", [2, 4])
def test_distributed_spmv_and_solve(world, tail_rows):
    """tail_rows: global size below which a level is replicated on every rank (0: the whole hierarchy stays distributed;
    600: the switch happens in the middle of the hierarchy)"""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    import os
    env = dict(os.environ, AMGXB_TAIL_ROWS=tail_rows, AMGXB_RUN_UNVALIDATED="1")      # every section ran on 2 x B200 in round 2 (both exchange paths)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29611 + world), str(ROOT / "tests" / "dist_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(ROOT), env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DIST_GPU_OK" in r.stdout
    assert r.stdout.count("DIST_BLOCK_DILU_OK") == 4, r.stdout[-3000:]
    assert "DIST_NONSYMMETRIC_OK" in r.stdout, r.stdout[-3000:]
    assert all(t in r.stdout for t in ("DIST_PARTITION_VECTOR_OK", "DIST_COMM_MAPS_OK", "DIST_READ_SYSTEM_OK")), r.stdout[-3000:]
