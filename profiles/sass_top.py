"""Summarise `ncu --page source --csv --print-source sass` output: top stall locations.
usage: ncu -i X.ncu-rep --page source --csv --print-source sass > x.csv; python profiles/sass_top.py x.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hi = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
hdr = rows[hi]
col = {k: i for i, k in enumerate(hdr)}
stalls = [k for k in hdr if k.startswith("stall_")]
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        samp = float(r[col["Warp Stall Sampling (All Samples)"]] or 0)
    except ValueError:
        continue
    data.append((samp, r))
tot = sum(d[0] for d in data) or 1.0
print(f"total samples {tot:.0f}, instructions {len(data)}")
agg = {}
for samp, r in data:
    for k in stalls:
        try:
            agg[k] = agg.get(k, 0.0) + float(r[col[k]] or 0)
        except ValueError:
            pass
print("stall reasons (share of samples):")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {k:28s} {100 * v / tot:6.2f}%")
print("top instructions:")
for idx, (samp, r) in enumerate(data):
    r.append(idx)
for samp, r in sorted(data, key=lambda d: -d[0])[:n]:
    top = sorted(((float(r[col[k]] or 0), k) for k in stalls), reverse=True)[:2]
    print(f"  {100 * samp / tot:5.2f}%  #{r[-1]:5d} {r[col['Source']][:70]:70s} "
          + " ".join(f"{k[6:]}={v:.0f}" for v, k in top if v))
