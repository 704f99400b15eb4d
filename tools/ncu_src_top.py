"""Summarise an `ncu --page source --csv` dump: top SASS instructions by stall samples + stall reasons."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for blk in blocks:
    hdr, data = blk["rows"][0], [r for r in blk["rows"][1:] if len(r) == len(blk["rows"][0])]
    ix = {h: i for i, h in enumerate(hdr)}
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {h: sum(int(r[ix[h]]) for r in data) for h in reasons}
    print("==", blk["name"][:60], "samples", tot)
    print("   stall reasons:", ", ".join(f"{k[6:]}={v * 100 // max(tot, 1)}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 100 >= tot))
    top = sorted(range(len(data)), key=lambda k: -int(data[k][ix["# Samples"]]))[:topn]
    for k in sorted(top):
        r = data[k]
        why = max(reasons, key=lambda h: int(r[ix[h]]))
        print(f"{k:5d} {int(r[ix['# Samples']]) * 100.0 / max(tot,1):5.1f}%  {why[6:]:10s} exec={r[ix['Instructions Executed']]:>9s} {r[ix['Source']].strip()[:90]}")
