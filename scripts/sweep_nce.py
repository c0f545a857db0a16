"""Timing sweep of the tcgen05 InfoNCE kernel (run on the GPU box): fixed overhead vs per-tile cost."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moma_b200 import _lib, ops
from moma_b200._lib import BF16, check

lib = _lib.load()
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def time_kernel(B, D, K, splits=None, cold=True, reps=15, qscale=1.0):
    q = (torch.randn(B, D, device=dev) * qscale).to(torch.bfloat16)
    queue = torch.nn.functional.normalize(torch.randn(K, D, device=dev)).to(torch.bfloat16)
    if splits is None:
        splits = lib.moma_nce_num_splits(B, D, K, BF16)
    st = torch.empty((3, splits, B), device=dev)
    O = torch.empty((splits, B, D), device=dev)
    ts = []
    for _ in range(reps):
        if cold:
            flush.fill_(1)
        else:
            torch.cuda._sleep(200000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib.moma_nce_partial(q.data_ptr(), queue.data_ptr(), B, D, K, 1 / 0.15, BF16, splits, st[0].data_ptr(),
                                   st[1].data_ptr(), st[2].data_ptr(), O.data_ptr(), torch.cuda.current_stream().cuda_stream))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], splits


if __name__ == "__main__":
    print("B D K splits tiles/cta us(cold) us(warm) TFLOP/s(cold)")
    qscale = float(os.environ.get("QSCALE", "1.0"))
    for (B, D, K) in [(512, 128, 16384), (512, 128, 65536), (512, 128, 262144), (512, 128, 1048576),
                      (256, 128, 16384), (256, 128, 65536), (1024, 128, 65536), (1024, 256, 131072), (512, 64, 65536)]:
        for splits in [None]:
            try:
                us, sp = time_kernel(B, D, K, splits, qscale=qscale)
                usw, _ = time_kernel(B, D, K, splits, cold=False, qscale=qscale)
            except Exception as e:
                print(B, D, K, splits, "ERR", e); continue
            bn = 64 if D == 256 else 128
            nq = 1 if (D == 256 or B <= 128) else 2
            mg = -(-B // (128 * nq))
            print(B, D, K, sp, f"{K / bn / sp:.1f}", f"{us:.1f}", f"{usw:.1f}", f"{4.0 * B * K * D / us / 1e6:.0f}", f"ctas={mg * sp}")
