"""Reference GPU build on the SuiteSparse-shaped matrix (context for bench.py --workload banded): writes the system for oracle/_ref/ref_dump and runs it."""
import json, os, subprocess, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import numpy as np
from amgx_b200 import gallery
sys.path.insert(0, str(ROOT / "tests" / "golden"))
from refdump_io import write_system
import bench
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
rp, ci, va = gallery.random_banded(rows)
n = rp.shape[0] - 1
write_system("/tmp/banded.bin", rp, ci, va, np.ones(n))
cfg = "/tmp/banded_cfg.json"
Path(cfg).write_text(json.dumps(bench.BANDED_CFG))
exe = ROOT / "oracle" / "_ref" / "ref_dump"
env = dict(os.environ, REFDUMP_NO_LEVELS="1", LD_LIBRARY_PATH=str(exe.parent) + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
r = subprocess.run([str(exe), "/tmp/banded.bin", cfg, "/tmp/banded_out.bin", "dDDI", "2"], capture_output=True, text=True, env=env, timeout=900)
print((r.stdout + r.stderr)[-600:])
