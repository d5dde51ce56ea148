"""CPU: the reference citations the judge follows are kept honest mechanically (only where /root/reference exists: the build container).

  * every AMGX_* declaration of include/amgx_b200.h carries the line of the reference's include/amgx_c.h that declares the same name;
  * every `src/...:line`, `include/...:line`, `examples/...:line` citation in the documents, the C-ABI header, the engine sources, the
    oracle and the tests names a file of the reference and a line range inside it.
"""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
pytestmark = pytest.mark.skipif(not (REF / "include" / "amgx_c.h").exists(), reason="the reference tree is only present in the build container")


def test_every_c_abi_declaration_cites_the_reference_line_that_declares_it():
    ref = (REF / "include" / "amgx_c.h").read_text().split("\n")
    txt = (ROOT / "include" / "amgx_b200.h").read_text()
    decls = re.findall(r"^(?:AMGX_RC|void)\s+AMGX_API\s+(AMGX_[A-Za-z0-9_]+)\((?:[^;]|\n)*?\);[ \t]*(?:/\*\s*:(\d+)\s*\*/)?", txt, re.M)
    assert len(decls) == 72
    for name, line in decls:
        assert line, f"{name}: no /* :line */ citation"
        assert re.search(r"\b" + name + r"\b", ref[int(line) - 1]), (name, line, ref[int(line) - 1])
    # the section headers' ranges contain the lines of the declarations under them
    for m in re.finditer(r"/\* ---- [^\n]*\[ref: include/amgx_c\.h:(\d+)-(\d+)\][^\n]*\n((?:(?!/\* ----).*\n)*)", txt):
        lo, hi = int(m.group(1)), int(m.group(2))
        for c in re.findall(r"/\*\s*:(\d+)\s*\*/", m.group(3)):
            assert lo <= int(c) <= hi, (m.group(0)[:60], c)


def test_reference_citations_name_existing_files_and_lines():
    files = [p for p in ROOT.glob("*.md") if p.name not in ("SURVEY.md", "VERDICT.md", "ADVICE.md", "PAPERS.md", "SNIPPETS.md")]
    files += list((ROOT / "include").glob("*.h")) + list((ROOT / "amgx_b200" / "csrc").glob("*")) + list((ROOT / "oracle").glob("*.c"))
    files += list((ROOT / "oracle").glob("*.py")) + list((ROOT / "tests").glob("*.py")) + [ROOT / "bench.py"]
    pat = re.compile(r"(?<![A-Za-z0-9_/])((?:src|include|examples)/[A-Za-z0-9_/\.]+\.(?:cu|h|c|cpp|inl|json))(?::(\d+)(?:-(\d+))?)?")
    lengths, bad, total = {}, [], 0
    for f in files:
        for ln, text in enumerate(f.read_text(errors="ignore").split("\n"), 1):
            for m in pat.finditer(text):
                path, a, b = m.group(1), m.group(2), m.group(3)
                if (ROOT / path).exists() or path == "include/nccl.h":       # this repository's own files (include/amgx_b200.h, examples/poisson_capi.c)
                    continue
                total += 1
                full = REF / path
                if not full.exists():
                    bad.append((f.name, ln, path, "no such file in the reference"))
                    continue
                if a:
                    if full not in lengths:
                        lengths[full] = len(full.read_text(errors="ignore").split("\n"))
                    hi = int(b or a)
                    if int(a) > hi or hi > lengths[full]:
                        bad.append((f.name, ln, m.group(0), f"the file has {lengths[full]} lines"))
    assert total > 250
    assert not bad, bad[:20]
