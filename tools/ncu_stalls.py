"""Per-region stall-reason breakdown of an ncu source-page CSV: python tools/ncu_stalls.py src.csv [lo hi]
(regions = runs of instructions with the same execution count, as tools/ncu_region.py)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iA, iS, iI, iSm = (hdr.index(x) for x in ('Address', 'Source', 'Instructions Executed', '# Samples'))
stall = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = []
for r in rows[2:]:
    try:
        data.append((int(r[iA], 16), r[iS].strip(), int(r[iI]), int(r[iSm]), [int(r[i] or 0) for i, _ in stall]))
    except Exception:
        pass
base = data[0][0]
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 40
tot = collections.Counter(); n = 0
for a, s, i, sm, st in data:
    if lo <= a - base <= hi:
        n += sm
        for (_, name), v in zip(stall, st):
            tot[name] += v
allsm = sum(d[3] for d in data)
print(f"range {lo:#x}-{hi:#x}: {n} samples ({100 * n / allsm:.1f} % of all)")
for k, v in tot.most_common(12):
    print(f"  {k:20s} {v:7d}  {100 * v / max(1, n):5.1f} %")
if len(sys.argv) > 4:  # top instructions of one reason
    j = [name for _, name in stall].index(sys.argv[4])
    top = sorted(((d[4][j], d[0] - base, d[1]) for d in data if lo <= d[0] - base <= hi), reverse=True)[:25]
    for v, a, s in top:
        print(f"    {v:6d} {a:05x} {s}")
