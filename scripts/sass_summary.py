"""Opcode counts per kernel of the shipped library (cuobjdump -sass): which kernels use tcgen05 (UTC*MMA, LDTM/STTM),
TMA (UTMALDG), warp-level MMA (HMMA), MUFU.   python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "moma_b200", "lib", "libmoma_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kern, counts = None, collections.OrderedDict()
pats = collections.OrderedDict([("UTC*MMA (tcgen05.mma)", r"\bUTC[A-Z]*MMA"), ("LDTM (tcgen05.ld)", r"\bLDTM"), ("STTM (tcgen05.st)", r"\bSTTM"),
                                ("UTMALDG (TMA load)", r"\bUTMALDG"), ("SYNCS (mbarrier)", r"\bSYNCS"), ("HMMA (mma.sync)", r"\bHMMA"),
                                ("MUFU.EX2", r"MUFU\.EX2"), ("FFMA2/FADD2", r"\bF(FMA|ADD)2"), ("LDGSTS (cp.async)", r"\bLDGSTS"),
                                ("total", r"^\s*/\*[0-9a-f]{4,}\*/")])
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("void ", "")
        counts[kern] = collections.Counter()
        continue
    if kern:
        for name, pat in pats.items():
            if re.search(pat, line):
                counts[kern][name] += 1
print("# cuobjdump -sass moma_b200/lib/libmoma_b200.so : instruction counts per kernel (sm_100a)")
print(f"{'kernel':70s} " + " ".join(f"{n.split(' ')[0]:>9s}" for n in pats))
for k, c in counts.items():
    if c["total"] == 0:
        continue
    print(f"{k[:70]:70s} " + " ".join(f"{c[n]:9d}" for n in pats))
