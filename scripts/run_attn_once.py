"""One attention forward + backward at the C2 shape (N=256 tokens, C=128, 4 heads) -- target for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moma_b200 import ops
dev = torch.device("cuda")
torch.manual_seed(0)
N, C, H = 256, 128, 4
x = torch.randn(N, C, device=dev, requires_grad=True)
wq = (torch.randn(3 * C, C, device=dev) / C ** 0.5).requires_grad_()
bq = torch.zeros(3 * C, device=dev, requires_grad=True)
wp = (torch.randn(C, C, device=dev) / C ** 0.5).requires_grad_()
bp = torch.zeros(C, device=dev, requires_grad=True)
for _ in range(3):
    y = ops.attention(x, wq, bq, wp, bp, H)
    y.square().sum().backward()
torch.cuda.synchronize()
print("ok")
