"""Per-instruction warp-state samples of a SASS index range of an `ncu --page source --csv --print-source sass` dump.
usage: python profiles/sass_region.py x.csv [lo hi]   (no range: samples per bucket of 50 instructions)"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
hdr = rows[hi]
col = {k: i for i, k in enumerate(hdr)}
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
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
if len(sys.argv) < 4:
    for b in range(0, len(data), 50):
        s = sum(d[0] for d in data[b:b + 50])
        if s / tot > 0.004:
            print(f"#{b:5d}-{b + 49:5d}  {100 * s / tot:6.2f}%  {data[b][1][col['Source']][:60]}")
    sys.exit(0)
lo, hi_ = int(sys.argv[2]), int(sys.argv[3])
sel = data[lo:hi_ + 1]
ssum = sum(d[0] for d in sel)
print(f"range #{lo}..#{hi_}: {ssum:.0f} samples = {100 * ssum / tot:.1f}% of the kernel's")
agg = {}
for samp, r in sel:
    for k in stalls:
        try:
            agg[k] = agg.get(k, 0.0) + float(r[col[k]] or 0)
        except ValueError:
            pass
print("  " + "  ".join(f"{k[6:]}={100 * v / max(ssum, 1):.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for i, (samp, r) in enumerate(sel):
    top = sorted(((float(r[col[k]] or 0), k) for k in stalls), reverse=True)[:2]
    print(f"#{lo + i:5d} {samp:5.0f} {r[col['Source']][:64]:64s} " + " ".join(f"{k[6:]}={v:.0f}" for v, k in top if v))
