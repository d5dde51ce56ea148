"""CPU, world_size 2 and 3 over gloo: host-side logic of the multi-GPU path (partition planner + halo protocol)."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.parametrize("world,partition", [(2, "offsets"), (3, "offsets"), (2, "vector"), (3, "vector"), (2, "maps"), (3, "maps")])
def test_partition_plan_and_halo_protocol(world, partition):
    import os
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29511 + world + {"offsets": 0, "vector": 10, "maps": 20}[partition]), str(ROOT / "tests" / "dist_cpu_worker.py")]
    env = dict(os.environ, AMGXB_TEST_PARTITION=partition)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=str(ROOT), env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DIST_CPU_OK" in r.stdout
