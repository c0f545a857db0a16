#!/usr/bin/env python
"""tcgen05 3xTF32 GEMM (csrc/gemm_tc.cu) against fp64 and against the warp-level 3xTF32 kernel (csrc/gemm.cu):
max error relative to max|C| and amortised time per launch (back-to-back launches, rotating outputs) for the Linear shapes
of the criterion step."""
import ctypes
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from moma_b200 import _lib


def vp(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def call_tc(L, A, a_mn, mask, B, b_mn, bias, C, M, N, K, relu, ws, h):
    return L.moma_debug_gemm_tc(vp(A), A.stride(0), a_mn, vp(mask), vp(B), B.stride(0), b_mn, vp(bias), vp(C), C.stride(0), M, N, K, relu,
                                vp(ws), ws.numel() * 4 if ws is not None else 0, ctypes.c_void_p(h))


def graph_time(fn, reps=20, replays=5):
    """us per launch of fn(stream_handle) captured `reps` times into one CUDA graph"""
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        fn(side.cuda_stream)
        side.synchronize()
        with torch.cuda.graph(g, stream=side):
            for _ in range(reps):
                fn(side.cuda_stream)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * replays)


def main():
    L = _lib.load()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    st = torch.cuda.current_stream().cuda_stream
    shapes = [(512, 512, 512), (512, 128, 512), (512, 128, 2048), (512, 2048, 2048), (512, 384, 128), (1024, 512, 512),
              (130, 72, 100), (4096, 128, 128), (512, 512, 128), (512, 2048, 128)]
    print("M N K mode err_tc err_simt us_tc us_simt")
    for (M, N, K) in shapes:
        for mode in ("NT", "NN", "TN"):                    # forward, d input (masked), d weight (masked, A stored [K, M])
            a_mn, b_mn = (mode == "TN"), (mode != "NT")
            if (b_mn and N % 4) or (a_mn and M % 4) or ((not a_mn) and K % 4) or ((not b_mn) and K % 4):
                continue
            A = torch.randn(K, M, device=dev) if a_mn else torch.randn(M, K, device=dev)
            Bm = torch.randn(K, N, device=dev) if b_mn else torch.randn(N, K, device=dev)
            bias = torch.randn(N, device=dev) if mode == "NT" else None
            mask = torch.randn_like(A) if mode != "NT" else None
            relu = 1 if mode == "NT" else 0
            wsb = L.moma_debug_gemm_tc_workspace_bytes(M, N, K)
            ws = torch.zeros(wsb // 4 + 64, device=dev)
            C = torch.empty(M, N, device=dev)
            rc = call_tc(L, A, int(a_mn), mask, Bm, int(b_mn), bias, C, M, N, K, relu, ws, st)
            torch.cuda.synchronize()
            err = L.moma_debug_gemm_tc_error()
            if rc != 0 or err != 0:
                print(M, N, K, mode, "rc", rc, "device error", err)
                continue
            A64 = A.double() * ((mask > 0).double() if mask is not None else 1.0)
            if a_mn:
                A64 = A64.t()
            ref = A64 @ (Bm.double() if b_mn else Bm.double().t())
            if bias is not None:
                ref = ref + bias.double()
            if relu:
                ref = ref.clamp_min(0)
            scale = ref.abs().max().item()
            e_tc = (C.double() - ref).abs().max().item() / scale
            e_simt, us_simt = float("nan"), float("nan")
            if mode == "NT":                                   # warp-level kernel through the product entry point
                wl = L.moma_linear_workspace_bytes(M, N, K)
                w2 = torch.zeros(wl // 4 + 64, device=dev)
                Y = torch.empty(M, N, device=dev)
                L.moma_linear_fwd(vp(A), vp(Bm), vp(bias), M, N, K, relu, vp(Y), vp(w2), w2.numel() * 4, ctypes.c_void_p(st))
                torch.cuda.synchronize()
                e_simt = (Y.double() - ref).abs().max().item() / scale
                us_simt = graph_time(lambda h: L.moma_linear_fwd(vp(A), vp(Bm), vp(bias), M, N, K, relu, vp(Y), vp(w2), w2.numel() * 4,
                                                                 ctypes.c_void_p(h)))
            us_tc = graph_time(lambda h: call_tc(L, A, int(a_mn), mask, Bm, int(b_mn), bias, C, M, N, K, relu, ws, h))
            print(M, N, K, mode, f"{e_tc:.2e} {e_simt:.2e} {us_tc:.1f} {us_simt:.1f}", flush=True)


if __name__ == "__main__":
    main()
