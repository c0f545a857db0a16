"""Accuracy and timing of moma_linear_fwd/bwd (3xTF32 tensor-core GEMM) against fp64 and the IEEE-fp32 library GEMM."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moma_b200 import ops
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda")
torch.cuda.set_stream(torch.cuda.Stream())        # leaves must be born on the stream the graphs are captured on


def rel(a, b):
    return float((a.detach().double() - b.detach().double()).norm() / b.detach().double().norm())


def timed(fn, reps=20, inner=10):
    """GPU time per call: `inner` calls captured in one CUDA graph (no Python between the kernels)."""
    s = torch.cuda.current_stream()
    if True:
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(inner):
                fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps / inner


print("   M    N    K relu | fwd err(ours) err(fp32 lib) | dX err  dW err  db err | fwd us (ours / lib)  bwd us (ours / lib)")
for (M, N, K, relu) in [(256, 512, 512, 1), (256, 128, 512, 0), (256, 384, 128, 0), (256, 128, 128, 0), (256, 2048, 2048, 1),
                        (256, 128, 2048, 0), (7, 33, 19, 1), (1, 128, 512, 0), (64, 512, 512, 1), (1024, 512, 512, 1)]:
    torch.manual_seed(M + N + K)
    x = torch.randn(M, K, device=dev, requires_grad=True)
    w = (torch.randn(N, K, device=dev) / K ** 0.5).requires_grad_()
    b = torch.randn(N, device=dev, requires_grad=True)
    gy = torch.randn(M, N, device=dev)
    y = ops.linear(x, w, b, relu=bool(relu))
    gx, gw, gb = torch.autograd.grad(y, (x, w, b), gy)
    xd, wd, bd = x.detach().double().requires_grad_(), w.detach().double().requires_grad_(), b.detach().double().requires_grad_()
    yd = torch.nn.functional.linear(xd, wd, bd)
    if relu:
        yd = torch.relu(yd)
    gxd, gwd, gbd = torch.autograd.grad(yd, (xd, wd, bd), gy.double())
    yl = torch.nn.functional.linear(x, w, b)
    if relu:
        yl = torch.relu(yl)

    def ours_f():
        with torch.no_grad():
            ops.linear(x, w, b, relu=bool(relu))

    def lib_f():
        with torch.no_grad():
            t = torch.nn.functional.linear(x, w, b)
            if relu:
                torch.relu_(t)

    def ours_b():
        torch.autograd.grad(ops.linear(x, w, b, relu=bool(relu)), (x, w, b), gy)

    def lib_b():
        t = torch.nn.functional.linear(x, w, b)
        if relu:
            t = torch.relu(t)
        torch.autograd.grad(t, (x, w, b), gy)

    tf_o, tf_l = timed(ours_f), timed(lib_f)
    tb_o, tb_l = timed(ours_b) - tf_o, timed(lib_b) - tf_l
    print(f"{M:5d} {N:4d} {K:4d} {relu:4d} | {rel(y, yd):.2e} {rel(yl, yd):.2e} | {rel(gx, gxd):.2e} {rel(gw, gwd):.2e} {rel(gb, gbd):.2e} |"
          f" {tf_o:6.1f} / {tf_l:6.1f}   {tb_o:6.1f} / {tb_l:6.1f}")
