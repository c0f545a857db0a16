"""bf16 InfoNCE pass vs the fp64 oracle fed the same bf16-rounded operands: worst relative error of loss rows and dq."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import moma_b200
from moma_b200 import ops
from oracle import moma_oracle as O
moma_b200.set_precision("bf16")
dev = torch.device("cuda")
worst = 0.0
for (B, D, K, T, seed) in [(256, 128, 16384, 0.15, 0), (512, 128, 65536, 0.15, 1), (64, 128, 4096, 0.15, 2), (128, 64, 2048, 0.07, 3),
                           (128, 256, 8192, 0.07, 4), (64, 128, 2048, 0.15, 5)]:
    torch.manual_seed(seed)
    q = torch.nn.functional.normalize(torch.randn(B, D, device=dev)) * 1.0
    k = torch.nn.functional.normalize(torch.randn(B, D, device=dev))
    queue = torch.nn.functional.normalize(torch.randn(K, D, device=dev))
    qb = queue.to(torch.bfloat16)
    out = ops.nce_rows(q.requires_grad_(), k, queue, qb, T, "bf16")
    dq = torch.autograd.grad(out.loss, q)[0]
    r = lambda t: O.round_bf16(t.detach().float().cpu().numpy()).astype(np.float64)
    loss_o, rows_o, dq_o, pim = O.nce_loss_and_grad(r(q), r(k), r(queue), T)
    e_rows = np.abs(out.rows.detach().cpu().numpy() - rows_o).max() / np.abs(rows_o).max()
    e_dq = np.linalg.norm(dq.cpu().numpy() - dq_o) / np.linalg.norm(dq_o)
    worst = max(worst, e_rows, e_dq)
    print(f"B={B} D={D} K={K} T={T}: splits={ops.nce_num_splits(B, D, K, ops.BF16)} rows err {e_rows:.2e}  dq err {e_dq:.2e}")
print(f"worst {worst:.2e}  (MOMA_B200_NCE_MIN_TILES={os.environ.get('MOMA_B200_NCE_MIN_TILES', 'default 2')})")
