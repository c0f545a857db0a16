"""Life-cycle trace of one launch of the tcgen05 InfoNCE kernel (hooked library: scripts/ablate_nce.py --build):
where the fixed cost of a launch goes.  clock64 stamps of the first and the last CTA of the grid, relative to the CTA's
first instruction; the launch is also timed with CUDA events (cold L2) for comparison.
    MOMA_B200_LIB=moma_b200/lib/libmoma_b200_ablate.so python scripts/trace_nce_life.py [B K]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from moma_b200 import _lib
from moma_b200._lib import BF16, check
lib = _lib.load()
D = 128
cases = [(512, 65536), (256, 16384)] if len(sys.argv) < 3 else [(int(sys.argv[1]), int(sys.argv[2]))]
names = ["entry", "setup done (bar init, TMEM alloc, sync)", "after griddepcontrol.wait", "producer: first queue tile issued",
         "issuer: Q + first tile landed", "issuer: first S issued", "softmax: first S complete", "softmax: first P handed over",
         "issuer: last PV issued (o_final commit)", "softmax: O final", "softmax: partials stored", "CTA exit"]
flush = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")
for B, K in cases:
    splits = lib.moma_nce_num_splits(B, D, K, BF16)
    q = torch.randn(B, D, device="cuda").to(torch.bfloat16)
    queue = torch.nn.functional.normalize(torch.randn(K, D, device="cuda")).to(torch.bfloat16)
    st = torch.empty((3, splits, B), device="cuda"); O = torch.empty((splits, B, D), device="cuda")
    os.environ["MOMA_TC_ABLATE"] = str(1024)
    dbg = torch.zeros(2 * B * 128 + 2 * (4096 + 64) + 64, device="cuda")
    ts = []
    for _ in range(5):
        flush.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib.moma_debug_nce_tc(q.data_ptr(), queue.data_ptr(), B, D, K, 1 / 0.15, splits, st[0].data_ptr(), st[1].data_ptr(),
                                    st[2].data_ptr(), O.data_ptr(), dbg.data_ptr(), torch.cuda.current_stream().cuda_stream))
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    raw = dbg[2 * B * 128:].view(torch.int64).cpu().numpy()
    print(f"== B={B} K={K} splits={splits} tiles/CTA={K / 128 / splits:.1f}  event-timed (cold) {sorted(ts)[2]:.1f} us")
    for tag, off in (("first CTA", 4096), ("last CTA", 4096 + 32)):
        life = raw[off:off + 32]
        wall = (life[21] - life[20]) / 1e3
        clk = life[11] - life[0]
        ghz = clk / max(wall * 1e3, 1)
        print(f"-- {tag}: {int(life[12])} tiles, CTA lifetime {wall:.2f} us = {clk} clk ({ghz:.2f} GHz)")
        for k, nm in enumerate(names):
            print(f"   {(life[k] - life[0]):8d} clk  {(life[k] - life[0]) / max(ghz, 1e-3) / 1e3:7.2f} us  {nm}")
    g = raw[4096 + 20], raw[4096 + 32 + 21]
    print(f"   first CTA entry -> last CTA exit (globaltimer): {(g[1] - g[0]) / 1e3:.2f} us")
