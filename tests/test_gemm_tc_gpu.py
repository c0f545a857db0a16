"""tcgen05 3xTF32 GEMM (csrc/gemm_tc.cu) through the C-ABI test entry point against fp64: the three operand layouts of a
Linear layer (forward, d input, d weight), ReLU masks, bias / ReLU epilogue, ragged shapes, split-K and single-split."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _vp(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _run(L, M, N, K, mode, use_mask, use_ws, dev):
    a_mn, b_mn = int(mode == "TN"), int(mode != "NT")
    A = torch.randn(K, M, device=dev) if a_mn else torch.randn(M, K, device=dev)
    B = torch.randn(K, N, device=dev) if b_mn else torch.randn(N, K, device=dev)
    bias = torch.randn(N, device=dev) if mode == "NT" else None
    relu = int(mode == "NT")
    mask = torch.randn_like(A) if use_mask else None
    ws = None
    if use_ws:
        ws = torch.zeros(L.moma_debug_gemm_tc_workspace_bytes(M, N, K) // 4 + 64, device=dev)
    C = torch.full((M, N), float("nan"), device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):                                   # twice: the tickets must be left at zero
        rc = L.moma_debug_gemm_tc(_vp(A), A.stride(0), a_mn, _vp(mask), _vp(B), B.stride(0), b_mn, _vp(bias), _vp(C), N, M, N, K,
                                  relu, _vp(ws), ws.numel() * 4 if ws is not None else 0, ctypes.c_void_p(st))
        assert rc == 0
    torch.cuda.synchronize()
    assert L.moma_debug_gemm_tc_error() == 0
    A64 = A.double() * ((mask > 0).double() if use_mask else 1.0)
    if a_mn:
        A64 = A64.t()
    ref = A64 @ (B.double() if b_mn else B.double().t())
    if bias is not None:
        ref = ref + bias.double()
    if relu:
        ref = ref.clamp_min(0)
    return (C.double() - ref).abs().max().item() / ref.abs().max().item()


@pytest.mark.parametrize("mode", ["NT", "NN", "TN"])
@pytest.mark.parametrize("shape", [(512, 512, 512), (512, 128, 2048), (128, 32, 32), (132, 72, 100), (1000, 200, 36), (4096, 384, 128)])
def test_gemm_tc_matches_fp64(mode, shape):
    from moma_b200 import _lib
    L = _lib.load()
    M, N, K = shape
    if (mode != "NT" and N % 4) or (mode == "TN" and M % 4) or (mode != "TN" and K % 4):
        pytest.skip("contiguous dimension not a multiple of 4 floats")
    dev = torch.device("cuda:0")
    torch.manual_seed(M + N + K)
    for use_mask in (False, True):
        if use_mask and mode == "NT":
            continue
        for use_ws in (True, False):
            err = _run(L, M, N, K, mode, use_mask, use_ws, dev)
            # 3xTF32: fp32-level.  One accumulation chain over K = 2048 (no workspace -> no split-K) drifts to a few 1e-6
            tol = 2e-6 if (use_ws or K <= 512) else 1e-5
            assert err < tol, (mode, shape, use_mask, use_ws, err)


def test_linear_dispatches_to_tc_and_matches_fp64():
    """moma_linear_fwd / bwd on a head-sized layer: the product path (gemm_nt -> gemm_tc) against fp64 autograd."""
    from moma_b200 import ops
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    x = torch.randn(512, 512, device=dev, requires_grad=True)
    w = torch.randn(512, 512, device=dev, requires_grad=True) * 0.05
    w = w.detach().requires_grad_(True)
    b = torch.randn(512, device=dev, requires_grad=True)
    y = ops.linear(x, w, b, relu=True)
    g = torch.randn_like(y)
    y.backward(g)
    x64, w64, b64 = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    y64 = torch.relu(x64 @ w64.t() + b64)
    y64.backward(g.double())
    for got, ref in ((y, y64), (x.grad, x64.grad), (w.grad, w64.grad), (b.grad, b64.grad)):
        assert float((got.double() - ref).abs().max() / ref.abs().max()) < 2e-6


@pytest.mark.parametrize("case", [(2048, 128, 8, 3, 4, 512), (1100, 64, 4, 1, 2, 500), (4096, 128, 8, 5, 8, 512), (1536, 64, 8, 0, 1, 40)])
def test_attention_rows_keys_split_over_warps(case):
    """Owned-rows attention with many keys (the split-KV kernel of csrc/attn_tc.cu) against fp64 attention over all
    tokens, restricted to the same rows."""
    from moma_b200 import ops
    N, C, H, q_start, q_stride, q_count = case
    dev = torch.device("cuda:0")
    torch.manual_seed(N + C)
    x = torch.randn(N, C, device=dev)
    wq, bq = torch.randn(3 * C, C, device=dev) * C ** -0.5, torch.randn(3 * C, device=dev) * 0.1
    wp, bp = torch.randn(C, C, device=dev) * C ** -0.5, torch.randn(C, device=dev) * 0.1
    y = ops.attention_rows(x, wq, bq, wp, bp, H, q_start, q_stride, q_count)
    x64, wq64, bq64, wp64, bp64 = (t.double() for t in (x, wq, bq, wp, bp))
    qkv = (x64 @ wq64.t() + bq64).reshape(N, 3, H, C // H).permute(1, 2, 0, 3)          # [3, H, N, hd]
    att = torch.softmax(qkv[0] @ qkv[1].transpose(-1, -2) * (C // H) ** -0.5, dim=-1)
    ref = ((att @ qkv[2]).transpose(0, 1).reshape(N, C) @ wp64.t() + bp64)[q_start::q_stride][:q_count]
    assert float((y.double() - ref).abs().max() / ref.abs().max()) < 1e-5
    qkv32 = (x64 @ wq64.t() + bq64).float()
    y2 = ops.attention_rows_from_qkv(qkv32, wp, bp, H, q_start, q_stride, q_count)
    assert float((y2.double() - ref).abs().max() / ref.abs().max()) < 1e-5
