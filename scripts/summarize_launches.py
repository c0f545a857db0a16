"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of `bench.py`:
per-kernel totals of ONE replayed criterion step (the kernels between two L2-flush reads) and its shares."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
names = [r['Kernel Name'] for r in rows]
dur = [float(r['Metric Value']) / 1000.0 for r in rows]
flush = [i for i, n in enumerate(names) if 'reduce_kernel' in n and dur[i] > 40.0]      # the 256 MiB read-flush
# pick the last window between two flushes that contains the tcgen05 kernel exactly once (a graph replay)
win = None
for a, b in zip(flush[:-1], flush[1:]):
    seg = names[a + 1:b]
    if sum('nce_tc2_kernel' in n or 'nce_tc_kernel' in n for n in seg) == 1 and any('ema_multi' in n for n in seg):
        win = (a + 1, b)
if win is None:
    print("no complete step window found"); sys.exit(1)
a, b = win
agg = collections.OrderedDict(); tot = 0.0
for i in range(a, b):
    if 'unrolled_elementwise' in names[i] and dur[i] > 40.0:
        continue                                                                         # flush fill
    n = names[i].replace('(anonymous namespace)::', '').replace('<unnamed>::', '').replace('unnamed>::', '')
    n = re.sub(r'\(.*', '', re.sub(r'<.*', '', n)).replace('void ', '')[:60]
    tot += dur[i]
    agg.setdefault(n, [0, 0.0]); agg[n][0] += 1; agg[n][1] += dur[i]
print(f"one replayed step: {sum(c for c, _ in agg.values())} kernels, {tot:.1f} us summed (ncu: cold caches, serialised -- compare SHARES)")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:8.1f} us {100 * t / tot:5.1f}%  x{c:<3d} {n}")
