"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of scripts/run_step_once.py (or bench.py):
per-kernel totals of ONE criterion step -- the kernels between two L2-flush reads -- and their shares.
    python scripts/summarize_launches.py gpurun_out/p_launches_c3.csv"""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
names = [r['Kernel Name'] for r in rows]
dur = [float(r['Metric Value']) / 1000.0 for r in rows]
flush = [i for i, n in enumerate(names) if 'reduce_kernel' in n and 'nce_reduce' not in n and dur[i] > 30.0]   # the 256 MiB flush read
win = None
for a, b in zip(flush[:-1], flush[1:]):          # the last complete window that holds exactly one tcgen05 launch
    seg = names[a + 1:b]
    if sum(bool(re.search(r'nce_tc\d?_kernel', n)) for n in seg) == 1 and any('ema_multi' in n for n in seg):
        win = (a + 1, b)
if win is None:
    print("no complete step window found"); sys.exit(1)
a, b = win
agg = collections.OrderedDict(); tot = 0.0
for i in range(a, b):
    n = names[i].replace('(anonymous namespace)::', '').replace('<unnamed>::', '').replace('unnamed>::', '')
    n = re.sub(r'\(.*', '', re.sub(r'<.*', '', n)).replace('void ', '')[:60]
    tot += dur[i]
    agg.setdefault(n, [0, 0.0]); agg[n][0] += 1; agg[n][1] += dur[i]
print(f"one criterion step: {sum(c for c, _ in agg.values())} kernels, {tot:.1f} us summed (ncu: cold caches, serialised -- compare SHARES)")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:8.1f} us {100 * t / tot:5.1f}%  x{c:<3d} {n}")
