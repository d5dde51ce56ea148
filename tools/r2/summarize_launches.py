"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: share / count / average per kernel (template arguments kept short)."""
import collections, csv, re, sys
path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
lines = [l for l in open(path, errors="replace") if not l.startswith("==")]
tot = collections.defaultdict(lambda: [0, 0.0])
for d in csv.DictReader(lines):
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    try:
        v = float(d["Metric Value"].replace(",", ""))
    except Exception:
        continue
    u = d["Metric Unit"]
    ns = v * 1e3 if u in ("us", "usecond") else v * 1e6 if u in ("ms", "msecond") else v
    name = re.sub(r"amgxb::|\(anonymous namespace\)::|<unnamed>::|void ", "", d["Kernel Name"])
    name = re.sub(r"\(.*", "", name)[:110]
    tot[name][0] += 1
    tot[name][1] += ns
T = sum(v[1] for v in tot.values()) or 1.0
print(f"# {title}\n\ntotal kernel time {T/1e3:.1f} us over {sum(v[0] for v in tot.values())} launches (cold-cache, serialised: compare SHARES)\n")
print("| share | launches | avg us | kernel |\n|---|---|---|---|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"| {100*v[1]/T:.2f}% | {v[0]} | {v[1]/v[0]/1e3:.1f} | {k} |")
