"""Host logic of bench.py without a GPU: the engine is replaced by a stand-in (tests/bench_dryrun_worker.py), so these tests say
nothing about kernels -- they pin how the JSON line is assembled, that a failing / slow context object never costs the line, and that
the multi-GPU flow (run here over gloo, world size 2) ends even when a collective context object hangs."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
WORKER = ROOT / "tests" / "bench_dryrun_worker.py"
sys.path.insert(0, str(ROOT))

CONTRACT_KEYS = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                 "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline", "reference_gpu", "parity", "final_relative_residual"]


def _line(stdout):
    lines = [ln for ln in stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, stdout[-2000:]          # ONE JSON line
    return json.loads(lines[0])


def test_other_workloads_children(tmp_path):
    """Every extra workload is a child `bench.py <flags> --no-cpu-baseline --no-reference-gpu --no-extras`; a child that prints no line,
    fails or overruns its time costs only its own entry."""
    import bench
    small = [("poisson_small", ["--workload", "poisson", "--grid", "16", "--steps", "2", "--warmup", "1"]),
             ("classical_small", ["--workload", "classical", "--grid", "16", "--steps", "2", "--warmup", "1"]),
             ("banded_small", ["--workload", "banded", "--rows", "20000", "--steps", "2", "--warmup", "1", "+reference-gpu"]),
             ("block_small", ["--workload", "block", "--grid", "8", "--mode", "dDFI", "--steps", "2", "--warmup", "1", "+reference-gpu"])]
    out = bench.other_workloads(budget_s=200, per_run_s=100, workloads=small, script=WORKER)
    for name, _ in small:
        assert "error" not in out[name] and "skipped" not in out[name], out[name]
        assert out[name]["metric"] == bench.METRIC and out[name]["value"] > 0 and out[name]["n_gpus"] == 1
        assert "other_workloads" not in out[name] and "cpu_baseline" not in out[name]           # --no-extras reached the child; None-valued keys dropped
    assert "unavailable" in out["banded_small"]["reference_gpu"] and "reference_gpu" not in out["poisson_small"]        # the harness ran (and found no GPU) for that one only
    assert "SuiteSparse-shaped" in out["banded_small"]["config"]["workload"] and "20000 rows" in out["banded_small"]["config"]["workload"]
    assert "classical AMG" in out["classical_small"]["config"]["workload"] and "iteration" not in out["classical_small"]["roofline"]
    assert "unavailable" in out["block_small"]["reference_gpu"]
    assert out["block_small"]["dtype"] == "f32 matrix / f64 vectors" and "block4" in out["block_small"]["roofline"]["kernel"]
    # the flags of the real list parse
    for _, flags in bench.EXTRA_WORKLOADS:
        r = subprocess.run([sys.executable, str(ROOT / "bench.py"), *[f for f in flags if f != "+reference-gpu"], "--no-cpu-baseline", "--no-reference-gpu", "--no-extras", "--help"],
                           capture_output=True, text=True)
        assert r.returncode == 0

    bad = tmp_path / "bad.py"
    bad.write_text("import sys, time\nif '--grid' in sys.argv:\n    time.sleep(60)\nprint('no json here'); sys.exit(3)\n")
    out = bench.other_workloads(budget_s=42.5, per_run_s=3, workloads=[("fails", []), ("slow", ["--grid", "1"]), ("late", [])], script=bad)
    assert "timed out" in out["slow"]["error"]
    assert out["fails"]["error"] == "no JSON line" and out["fails"]["returncode"] == 3
    assert "skipped" in out["late"]                                                              # fewer than 40 s of the budget left


def test_single_gpu_line_survives_failing_extras():
    """N = 1 default flow: the children (the real bench.py) stop at once without a device -- the line is printed all the same."""
    r = subprocess.run([sys.executable, str(WORKER), "--steps", "2", "--warmup", "1", "--no-cpu-baseline", "--no-reference-gpu"], capture_output=True, text=True, timeout=600)
    d = _line(r.stdout)
    for k in CONTRACT_KEYS:
        assert k in d, k
    assert d["n_gpus"] == 1 and d["scaling"] == "weak" and d["config"]["workload"].startswith("7-pt Poisson 256x256x256")
    assert d["roofline"]["bound"] == "hbm" and d["roofline"]["spmv"]["frac"] > 0 and d["roofline"]["iteration"]["levels"] == 3
    assert set(d["other_workloads"]) == {"poisson512", "classical512", "banded4m", "block160_dDFI"}
    assert all("error" in v for v in d["other_workloads"].values())
    assert "strong_512" not in d


def _torchrun(port, extra_env=None, args=()):
    env = dict(os.environ, AMGXB_BENCH_BLOCK_NX="10", **(extra_env or {}))      # block_weak generates its slab in numpy: keep it small here
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port),
           str(WORKER), "--gpus", "2", "--steps", "2", "--warmup", "1", "--no-parity", *args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)


def test_two_rank_line_has_strong_512():
    r = _torchrun(29541)
    assert r.returncode == 0, r.stderr[-2000:]
    d = _line(r.stdout)
    assert d["n_gpus"] == 2 and d["value"] == 2 * d["config"]["global_iterations_per_sec"] and d["roofline"] is None
    s = d["strong_512"]
    assert s["scaling"] == "strong" and s["rows"] == 512 ** 3 and "z-slabs of 256 planes" in s["workload"] and s["warmup"] >= 3
    w = d["block_weak"]
    assert w["scaling"] == "weak" and w["block_rows_global"] == 2 * 10 ** 3 and w["value"] == 2 * w["global_iterations_per_sec"] and "dDFI" in w["workload"]
    assert "other_workloads" not in d and "note" not in d
    d = _line(_torchrun(29542, args=("--no-extras",)).stdout)
    assert "strong_512" not in d and "block_weak" not in d
    d = _line(_torchrun(29543, args=("--strong", "--grid", "512")).stdout)
    assert "strong_512" not in d and "block_weak" not in d and d["scaling"] == "strong" and d["value"] == d["config"]["global_iterations_per_sec"]


def test_two_rank_line_is_printed_when_a_context_object_hangs():
    """Rank 1 never returns from its first solve of the strong_512 problem: after AMGXB_BENCH_GUARD_S seconds rank 0 prints the line
    without the object and every rank ends."""
    r = _torchrun(29544, extra_env={"BENCH_DRYRUN_HANG": "strong", "AMGXB_BENCH_GUARD_S": "5"})
    d = _line(r.stdout)
    assert d["n_gpus"] == 2 and d["value"] > 0 and "strong_512" not in d
    assert "did not finish" in d["note"]


def test_bench_configurations_pass_config_check():
    """every configuration dictionary bench.py hands to the engine names only components the engine provides (no GPU needed)"""
    import bench
    from amgx_b200 import capi
    capi.load_library()
    for name in ("CLASSICAL_CFG", "BLOCK_CFG", "BANDED_CFG", "HOST_PATH_CFG"):
        cfg = capi.Config(getattr(bench, name))
        ok, msg = capi.config_check(cfg)
        cfg.destroy()
        assert ok, (name, msg)


def test_sigterm_during_the_context_objects_prints_the_line():
    """whoever launched the bench may end it while the context objects run (they can take minutes): the line, whose headline numbers
    are final by then, is printed as it stands"""
    import signal
    import time
    code = ("import sys, time; sys.argv = ['bench.py', '--steps', '2', '--warmup', '1', '--no-cpu-baseline', '--no-reference-gpu']\n"
            "import tests.bench_dryrun_worker as w\n"
            "def slow(**k):\n    print('CONTEXT_STARTED', file=sys.stderr, flush=True); time.sleep(200)\n"
            "w.bench.other_workloads = slow\n"
            "w.bench.main()\n")
    p = subprocess.Popen([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=str(ROOT))
    t0 = time.time()
    for ln in p.stderr:                      # wait until the context phase has begun
        if "CONTEXT_STARTED" in ln or time.time() - t0 > 120:
            break
    time.sleep(0.5)
    p.send_signal(signal.SIGTERM)
    out, _ = p.communicate(timeout=60)
    d = _line(out)
    assert p.returncode == 0 and d["value"] > 0 and "other_workloads" not in d
    assert "terminated from outside" in d["note"]
