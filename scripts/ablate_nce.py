"""Timing ablations of the v2 tcgen05 InfoNCE kernel (results are wrong on purpose; timing only).

Build the hooked library first:  python scripts/ablate_nce.py --build   (-> moma_b200/lib/libmoma_b200_ablate.so)
then on the GPU box:             MOMA_B200_LIB=moma_b200/lib/libmoma_b200_ablate.so python scripts/ablate_nce.py
bits: 64 no TMA of the queue tiles, 1 no exp2, 2 no tcgen05.ld of S, 4 no PV MMAs, 8 no S MMAs, 16 no tcgen05.st of P, 32 no max exchange.
"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ABL = os.path.join(ROOT, "moma_b200", "lib", "libmoma_b200_ablate.so")
if "--build" in sys.argv:
    from moma_b200 import _build
    cmd = ["nvcc"] + _build.NVCC_FLAGS + ["-DMOMA_TC_ABLATE", "-I", _build.INCLUDE, "-o", ABL] + _build.sources()
    subprocess.check_call(cmd); print(ABL); sys.exit(0)
import torch
from moma_b200 import _lib
from moma_b200._lib import BF16, check
lib = _lib.load()
dev = torch.device("cuda")
B, D, K = 512, 128, 1 << 20
splits = lib.moma_nce_num_splits(B, D, K, BF16)
q = torch.randn(B, D, device=dev).to(torch.bfloat16)
queue = torch.nn.functional.normalize(torch.randn(K, D, device=dev)).to(torch.bfloat16)
st = torch.empty((3, splits, B), device=dev); O = torch.empty((splits, B, D), device=dev)
tiles = K / 128 / splits
names = {0: "full kernel", 64: "no TMA", 1: "no exp2", 2: "no LDTM(S)", 16: "no STTM(P)", 32: "no max exchange", 2 | 16: "no TMEM ld/st",
         4: "no PV MMA", 8: "no S MMA", 12: "no MMA at all", 51: "MMA only", 51 | 64: "MMA only, no TMA", 127: "barriers only"}
dbg = torch.zeros(2 * B * 128 + 64, device=dev)
print(f"B={B} D={D} K={K} splits={splits} tiles/CTA={tiles:.1f}")
for mask, name in names.items():
    os.environ["MOMA_TC_ABLATE"] = str(mask | 256)
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib.moma_debug_nce_tc(q.data_ptr(), queue.data_ptr(), B, D, K, 1 / 0.15, splits, st[0].data_ptr(), st[1].data_ptr(),
                                    st[2].data_ptr(), O.data_ptr(), dbg.data_ptr(), torch.cuda.current_stream().cuda_stream))
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); us = ts[len(ts) // 2]
    tr = dbg[2 * B * 128:].view(torch.int64).cpu().numpy()
    clk = (tr[1] - tr[0]) / max(int(tr[2]), 1)
    print(f"mask {mask:3d} {name:24s} {us:8.1f} us  {us / tiles * 1e3:7.0f} ns/tile  {clk:7.0f} clk/tile  ({clk / (us / tiles * 1e3):.2f} GHz)")
