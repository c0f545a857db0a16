"""What does a LAUNCH of the InfoNCE kernel's shape cost before any work is done?  Empty kernel, event-timed exactly like
bench.py times one kernel (L2 flush read, event, launch, event), for several dynamic shared-memory sizes / TMEM use."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moma_b200 import _lib
lib = _lib.load()
flush = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")
out = torch.zeros(1, dtype=torch.float32, device="cuda")
st = lambda: torch.cuda.current_stream().cuda_stream
def timed(fn, cold=True, reps=25):
    ts = []
    for _ in range(reps):
        if cold: torch.sum(flush, dim=(0,), keepdim=True, out=out)
        else: torch.cuda._sleep(100000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
print("ctas threads smem_KB tmem pdl | us after flush | us after sleep")
for ctas, thr, smem, tm, pdl in [(148, 384, 0, 0, 0), (148, 384, 0, 0, 1), (148, 384, 48, 0, 1), (148, 384, 100, 0, 1), (148, 384, 136, 0, 1),
                                 (148, 384, 196, 0, 1), (148, 384, 196, 1, 1), (148, 384, 196, 1, 0), (148, 128, 196, 1, 1), (74, 384, 196, 1, 1),
                                 (1, 32, 0, 0, 1)]:
    f = lambda: _lib.check(lib.moma_debug_probe_launch(ctas, thr, smem * 1024, tm, pdl, st()))
    for _ in range(3): f()
    print(f"{ctas:4d} {thr:4d} {smem:4d} {tm} {pdl} | {timed(f):6.2f} | {timed(f, cold=False):6.2f}")
# two back-to-back launches of the big shape: the second one's marginal cost
f2 = lambda: (_lib.check(lib.moma_debug_probe_launch(148, 384, 196 * 1024, 1, 1, st())), _lib.check(lib.moma_debug_probe_launch(148, 384, 196 * 1024, 1, 1, st())))
print(f"two launches 148x384 196KB tmem pdl: {timed(f2):6.2f} us")
f3 = lambda: (_lib.check(lib.moma_debug_probe_launch(148, 128, 0, 0, 1, st())), _lib.check(lib.moma_debug_probe_launch(148, 384, 196 * 1024, 1, 1, st())))
print(f"small-smem kernel then the big shape: {timed(f3):6.2f} us")
