"""Write profiles/r02_sass.md: SASS mnemonic counts per translation unit of the shipped sm_100a objects + excerpts (run HERE after a build;
cuobjdump needs no GPU).  What to look for: UBLKCP = TMA 1-D bulk copy (cp.async.bulk), SYNCS = mbarrier, UCGABAR_* = thread-block-cluster
barrier, *.STRONG.SYS = system-scope loads / stores of the NVLink peer-memory protocol, DFMA = the fp64 FMA chain."""
import re, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
OBJ = ROOT / "amgx_b200" / "_build"
units = ["k_spmv", "k_spmv_enc", "k_spmv_win", "k_blas", "k_transfer", "k_block", "dilu", "p2p", "fgmres", "dist"]
pats = {"UBLKCP": r"\bUBLKCP", "UBLKPF.L2 (bulk L2 prefetch)": r"\bUBLKPF", "SYNCS (mbarrier)": r"\bSYNCS", "UTMALDG": r"UTMALDG", "UCGABAR (cluster barrier)": r"UCGABAR", "LDG/STG .STRONG.SYS": r"(LDG|STG)\.E[.0-9A-Z]*\.STRONG\.SYS",
        "DFMA": r"\bDFMA", "FFMA": r"\bFFMA", "LDS": r"\bLDS", "SHFL": r"\bSHFL", "HMMA (legacy tensor)": r"\bHMMA", "UTC*MMA (tcgen05)": r"UTC[A-Z]*MMA"}
out = ["# r02 -- SASS evidence of the shipped sm_100a objects (`cuobjdump -sass amgx_b200/_build/*.o`, tools/r2/sass_evidence.py)\n",
       "No tensor-core instructions are expected: every hot kernel is an HBM-bound sparse / level-1 op (DESIGN 3).  TMA here is the 1-D bulk copy",
       "(`cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes` -> `UBLKCP`), not the tensor-map form (`UTMALDG`): CSR streams are 1-D.\n",
       "| object | " + " | ".join(pats) + " |", "|---|" + "---|" * len(pats)]
sass = {}
for u in units:
    f = OBJ / f"{u}.o"
    if not f.exists():
        continue
    txt = subprocess.run(["cuobjdump", "-sass", str(f)], capture_output=True, text=True).stdout
    sass[u] = txt
    out.append(f"| {u}.o | " + " | ".join(str(len(re.findall(p, txt))) for p in pats.values()) + " |")

def excerpt(unit, func_pat, line_pat, before=3, after=6, title=""):
    txt = sass.get(unit, "")
    m = re.search(r"Function : (\S*" + func_pat + r"\S*)", txt)
    if not m:
        return
    body = txt[m.start():]
    nxt = body.find("Function :", 10)
    body = body[:nxt] if nxt > 0 else body
    lines = body.splitlines()
    for i, l in enumerate(lines):
        if re.search(line_pat, l):
            out.append(f"\n## {title}\n\n`{m.group(1)[:150]}`\n\n```")
            out.extend(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", x)[:150] for x in lines[max(0, i - before):i + after])
            out.append("```")
            return

excerpt("k_spmv", "csr_tile_kernelIddLi256ELi2ELb0", r"UBLKCP", 6, 14, "producer warp of the CSR tile kernel (fused Jacobi): expect_tx on the stage's mbarrier, three bulk copies (row_ptr slice, values, columns)")
excerpt("k_spmv_enc", "csr_tile_enc_kernelIddLi256ELi2E", r"UBLKCP", 4, 20, "producer of the coded-stream tile kernel: code streams + dictionaries by bulk copy")
excerpt("k_spmv_win", "csr_window_kernelIddLi2E", r"UBLKCP", 4, 26, "producer of the sliding-window kernel (fused Jacobi): sliced-ELL values and 16-bit offsets, row map, slice offsets and the x ring chunk by bulk copy; UBLKPF.L2 = prefetch of the b / d slices")
excerpt("k_spmv", "csr_tile_kernelIddLi256ELi2ELb0", r"SYNCS\.PHASECHK|SYNCS\.ARRIVE", 2, 6, "consumer side: mbarrier phase check / arrive")
excerpt("dilu", "dilu_level_kernelIfdLi4", r"UCGABAR_ARV", 4, 6, "fused DILU level kernel: colour boundary = cluster barrier (barrier.cluster.arrive.release / wait.acquire)")
excerpt("p2p", "p2p_exchange_kernelId", r"STG\.E\.64\.STRONG\.SYS", 6, 8, "peer-memory exchange kernel: release store of the epoch flag into the neighbour's window")
excerpt("p2p", "p2p_allreduce_kernel", r"LDG\.E\.64\.STRONG\.SYS", 3, 8, "scalar all-reduce over peer memory: acquire loads of the ranks' flags")
(ROOT / "profiles" / "r02_sass.md").write_text("\n".join(out) + "\n")
print("\n".join(out[:20]))
