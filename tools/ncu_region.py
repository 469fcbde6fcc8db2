"""Summarise an ncu source-page CSV (SASS view): instruction share / stall samples per code region."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iA, iS, iI, iSm, iT = (hdr.index(x) for x in ('Address', 'Source', 'Instructions Executed', '# Samples', 'Thread Instructions Executed'))
data = []
for r in rows[2:]:
    try:
        data.append((int(r[iA], 16), r[iS].strip(), int(r[iI]), int(r[iSm]), int(r[iT])))
    except Exception:
        pass
base = data[0][0]
tot = sum(d[2] for d in data); tots = sum(d[3] for d in data)
print("total warp inst", tot, "samples", tots)
runs = []
for a, s, i, sm, t in data:
    op = s.split()[1] if s.startswith('@') else s.split()[0]
    if runs and abs(runs[-1]['cnt'] - i) <= max(3, 0.03 * i):
        r = runs[-1]; r['n'] += 1; r['sm'] += sm; r['end'] = a - base; r['thr'] += t; r['inst'] += i; r['ops'].append(op)
    else:
        runs.append(dict(start=a - base, end=a - base, cnt=i, n=1, sm=sm, thr=t, inst=i, ops=[op]))
for r in runs:
    if r['inst'] / tot > 0.004 or r['sm'] / tots > 0.01:
        c = collections.Counter(o.split('.')[0] for o in r['ops'])
        print("%05x-%05x cnt=%9d n=%3d inst%%=%5.2f smp%%=%5.2f thr=%4.1f %s" % (r['start'], r['end'], r['cnt'], r['n'], 100 * r['inst'] / tot, 100 * r['sm'] / tots, r['thr'] / max(1, r['inst']), dict(c.most_common(7))))
