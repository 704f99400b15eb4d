"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total time, share."""
import csv
import re
import sys
from collections import defaultdict

path, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
hdr = next(r for r in rows if "Kernel Name" in r)
ix = {h: i for i, h in enumerate(hdr)}
tot, cnt, grids = defaultdict(float), defaultdict(int), defaultdict(set)
for r in rows:
    if r is hdr or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "")
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0)
    tot[name] += v
    cnt[name] += 1
    grids[name].add(r[ix["Grid Size"]])
total = sum(tot.values())
print(f"# {title}\n")
print("| kernel | launches | total us | share | grids |")
print("|---|---|---|---|---|")
for k in sorted(tot, key=lambda k: -tot[k]):
    g = ", ".join(sorted(grids[k])[:3])
    print(f"| `{k}` | {cnt[k]} | {tot[k]:.1f} | {100 * tot[k] / total:.1f}% | {g} |")
print(f"\nTotal {total:.1f} us over {sum(cnt.values())} launches.")
