"""clock() trace of the v2 InfoNCE kernel, CTA (0,0), tiles 64..79 (hooked library: scripts/ablate_nce.py --build)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from moma_b200 import _lib
from moma_b200._lib import check
lib = _lib.load()
B, D, K, splits = 512, 128, 1 << 20, 37
q = torch.randn(B, D, device="cuda").to(torch.bfloat16)
queue = torch.nn.functional.normalize(torch.randn(K, D, device="cuda")).to(torch.bfloat16)
st = torch.empty((3, splits, B), device="cuda"); O = torch.empty((splits, B, D), device="cuda")
for mask in [int(a) for a in sys.argv[1:]] or [0, 127]:
    os.environ["MOMA_TC_ABLATE"] = str(mask | 256 | 512)
    dbg = torch.zeros(2 * B * 128 + 8192, device="cuda")
    for _ in range(3):
        check(lib.moma_debug_nce_tc(q.data_ptr(), queue.data_ptr(), B, D, K, 1 / 0.15, splits, st[0].data_ptr(), st[1].data_ptr(),
                                    st[2].data_ptr(), O.data_ptr(), dbg.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    raw = dbg[2 * B * 128:].view(torch.int32).cpu().numpy().astype(np.int64)
    w4 = raw[64:64 + 128].reshape(16, 8); w8 = raw[64 + 1024:64 + 1024 + 128].reshape(16, 8)
    iss = raw[64 + 2048:64 + 2048 + 128].reshape(16, 8)
    t0 = w4[0, 0]
    rel = lambda x: (x - t0) & 0xffffffff
    print(f"== mask {mask}: clk relative to warp 4's first stamp; tiles 64..79")
    print("tile | warp4: top  S seen  ld done  max xchg  exps+st  st waited  arrived | warp8: top S seen arrived | issuer: top  P seen  PV+cmt  kv seen  S+cmt")
    for i in range(14):
        a = " ".join(f"{rel(x):7d}" for x in w4[i, :7]); b_ = " ".join(f"{rel(x):7d}" for x in (w8[i, 0], w8[i, 1], w8[i, 6]))
        c = " ".join(f"{rel(x):7d}" for x in iss[i, :5])
        print(f"{64 + i:4d} | {a} | {b_} | {c}")
