"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of one step."""
import collections, csv, re, sys
path = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else -2          # which EMA-delimited step to print
lines = [l for l in open(path) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
names = [r['Kernel Name'] for r in rows]
idx = [i for i, n in enumerate(names) if 'ema_multi' in n]
a, b = idx[which], idx[which + 1] if which + 1 < len(idx) and which + 1 != 0 else len(rows)
agg = collections.OrderedDict(); tot = 0.0
for r in rows[a:b]:
    n = re.sub(r'<.*', '', r['Kernel Name'])[:64]
    t = float(r['Metric Value']) / 1000.0
    tot += t
    agg.setdefault(n, [0, 0.0]); agg[n][0] += 1; agg[n][1] += t
print(f"step {which}: {b - a} kernels, {tot:.1f} us (ncu: cold caches, serialised -- compare shares)")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:8.1f} us {100 * t / tot:5.1f}%  x{c:<3d} {n}")
