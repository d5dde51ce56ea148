"""Build libamgx_b200.so (the C-ABI engine) in-tree with nvcc for sm_100a.

One nvcc invocation per translation unit (parallel, incremental by mtime), then one link.
No torch involved: the library depends only on the CUDA runtime and (for multi-GPU) NCCL.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
OBJ = ROOT / "_build"
LIB = ROOT / "libamgx_b200.so"

NVCC = os.environ.get("NVCC", "nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "--extended-lambda", "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wno-unused-function",
          "-DAMGXB200_BUILD", "-w"]


# translation units whose arithmetic must be reproducible bit for bit by a sequential CPU restatement: no implicit FMA contraction
EXTRA = {"classical.cu": ["-fmad=false"], "spgemm.cu": ["-fmad=false"]}


def _nccl_flags():
    """Prefer the system NCCL (headers + libnccl.so); fall back to torch's bundled one."""
    inc, libs = [], []
    if Path("/usr/include/nccl.h").exists():
        libs = ["-lnccl"]
    else:  # pragma: no cover
        try:
            import nvidia.nccl as n  # type: ignore
            base = Path(n.__file__).parent
            inc = [f"-I{base / 'include'}"]
            libs = [f"-L{base / 'lib'}", "-l:libnccl.so.2", f"-Xlinker=-rpath={base / 'lib'}"]
        except Exception:
            libs = []
    return inc, libs


def sources():
    return sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cpp")))


def _needs(obj: Path, src: Path, headers_mtime: float) -> bool:
    if not obj.exists():
        return True
    m = obj.stat().st_mtime
    return m < src.stat().st_mtime or m < headers_mtime


def build(verbose: bool = False, force: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    inc, libs = _nccl_flags()
    hdrs = list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + list((ROOT.parent / "include").glob("*.h"))
    hm = max(h.stat().st_mtime for h in hdrs)
    jobs = []
    for src in sources():
        obj = OBJ / (src.stem + ".o")
        if force or _needs(obj, src, hm):
            cmd = [NVCC, *ARCH, *COMMON, *EXTRA.get(src.name, []), *inc, "-c", str(src), "-o", str(obj)]
            if src.suffix == ".cpp":
                cmd = [NVCC, *ARCH, *COMMON, "-x", "cu", *inc, "-c", str(src), "-o", str(obj)]
            jobs.append((src, cmd))

    def run(job):
        src, cmd = job
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r

    failed = False
    with ThreadPoolExecutor(max_workers=int(os.environ.get("AMGXB_BUILD_JOBS", "6"))) as ex:
        for src, r in ex.map(run, jobs):
            if r.returncode != 0:
                failed = True
                sys.stderr.write(f"--- nvcc failed on {src.name} ---\n{r.stdout}\n{r.stderr}\n")
            elif verbose and r.stderr.strip():
                sys.stderr.write(r.stderr)
    if failed:
        raise RuntimeError("amgx_b200: compilation failed")
    objs = [str(OBJ / (s.stem + ".o")) for s in sources()]
    if jobs or not LIB.exists():
        cmd = [NVCC, *ARCH, "-shared", "-o", str(LIB), *objs, "-lcudart", *libs]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("amgx_b200: link failed")
        # drop-in name used by applications linked against the reference (CMakeLists.txt:247-251)
        alias = ROOT / "libamgxsh.so"
        if alias.exists() or alias.is_symlink():
            alias.unlink()
        alias.symlink_to(LIB.name)
    return LIB


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print(p)
