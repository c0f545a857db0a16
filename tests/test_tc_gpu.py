"""GPU parity tests of the BF16 tcgen05/TMEM/TMA InfoNCE kernel (csrc/nce_tc.cu).

The oracle is fed the SAME bf16-representable operands as the kernel;
tolerance 1e-3 relative (north_star, BF16 mode), gradients compared norm-wise.
"""
import ctypes

import numpy as np
import pytest
import torch

from oracle import moma_oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-3


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def npy(t):
    return t.detach().float().cpu().numpy()


@pytest.fixture(scope="module")
def lib():
    from moma_b200 import _lib
    l = _lib.load()
    if not l.moma_has_tcgen05():
        pytest.skip("device is not compute capability 10.x")
    return l


def run_tc(lib, q, queue, T, n_splits=None, want_dbg=False):
    """q [B,D], queue [K,D] fp32 numpy -> kernel on their bf16 roundings."""
    from moma_b200 import ops
    from moma_b200._lib import BF16, check
    qb = torch.as_tensor(q).cuda().to(torch.bfloat16).contiguous()
    kb = torch.as_tensor(queue).cuda().to(torch.bfloat16).contiguous()
    B, D = qb.shape
    K = kb.shape[0]
    if n_splits is None:
        n_splits = lib.moma_nce_num_splits(B, D, K, BF16)
    BN = 64 if D == 256 else 128
    stats = torch.full((3, n_splits, B), float("nan"), device="cuda")
    Op = torch.full((n_splits, B, D), float("nan"), device="cuda")
    dbg = torch.zeros((2 * B, BN), device="cuda") if want_dbg else None
    check(lib.moma_debug_nce_tc(qb.data_ptr(), kb.data_ptr(), B, D, K, 1.0 / T, n_splits, stats[0].data_ptr(),
                                stats[1].data_ptr(), stats[2].data_ptr(), Op.data_ptr(),
                                None if dbg is None else dbg.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert lib.moma_debug_tc_error() == 0
    if dbg is not None:
        run_tc.info = dbg[B:].clone()
        dbg = dbg[:B]
    return stats, Op, dbg, qb, kb


def merged(stats, Op):
    """Merge split partials into (LSE over negatives, normalised O) in float64."""
    m = npy(stats[0]).astype(np.float64); l = npy(stats[1]).astype(np.float64); O_ = npy(Op).astype(np.float64)
    mstar = m.max(axis=0)
    w = np.exp(m - mstar)
    L = (w * l).sum(axis=0)
    Om = (w[:, :, None] * O_).sum(axis=0)
    return np.log(L) + mstar, Om / L[:, None], npy(stats[2]).max(axis=0)


def oracle_neg(qb, kb, T):
    q = npy(qb).astype(np.float64); qu = npy(kb).astype(np.float64)
    s = q @ qu.T / T
    mx = s.max(axis=1)
    e = np.exp(s - mx[:, None])
    return np.log(e.sum(1)) + mx, (e @ qu) / e.sum(1)[:, None], mx


def data(B, D, K, seed, qscale=0.7):
    rng = np.random.default_rng(seed)
    q = (rng.standard_normal((B, D)) * qscale).astype(np.float32)
    queue = O.normalize(rng.standard_normal((K, D))).astype(np.float32)
    return q, queue


def test_single_tile_scores_and_partials(lib):
    """One 128-row queue tile, one split: checks TMA + swizzle + both UMMA descriptor forms."""
    q, queue = data(128, 128, 128, 0)
    stats, Op, dbg, qb, kb = run_tc(lib, q, queue, 0.15, n_splits=1, want_dbg=True)
    S = npy(qb).astype(np.float64) @ npy(kb).astype(np.float64).T
    assert np.allclose(npy(dbg), S, rtol=1e-5, atol=1e-5)                 # S = Q . Tile^T (fp32 accumulate)
    lse, Onorm, mx = merged(stats, Op)
    lse_o, O_o, mx_o = oracle_neg(qb, kb, 0.15)
    assert np.allclose(mx, mx_o, rtol=1e-5, atol=1e-5)                    # true max tracked exactly
    assert np.allclose(lse, lse_o, rtol=TOL, atol=TOL)
    assert rel(Onorm, O_o) < 5e-3                                         # P rounded to bf16


@pytest.mark.parametrize("B,D,K", [(256, 128, 128), (256, 128, 1024), (512, 128, 4096), (200, 128, 1000),
                                   (64, 128, 2048), (128, 64, 1024), (300, 64, 777), (128, 256, 512),
                                   (512, 256, 2048), (70, 256, 1000)])
def test_shapes_vs_oracle(lib, B, D, K):
    q, queue = data(B, D, K, B + D + K)
    for splits in (None, 1, 3):
        BN = 64 if D == 256 else 128
        if splits is not None and splits > (K + BN - 1) // BN:
            continue
        stats, Op, _, qb, kb = run_tc(lib, q, queue, 0.15, n_splits=splits)
        lse, Onorm, mx = merged(stats, Op)
        lse_o, O_o, mx_o = oracle_neg(qb, kb, 0.15)
        assert np.allclose(mx, mx_o, rtol=1e-5, atol=1e-5), splits
        assert np.abs(lse - lse_o).max() < TOL * np.abs(lse_o).max(), splits
        assert rel(Onorm, O_o) < 5e-3, splits


def test_lazy_rescale_path(lib):
    """Queue rows whose scores grow tile after tile force the rare O-rescale branch."""
    B, D, K = 256, 128, 2048
    rng = np.random.default_rng(9)
    q = O.normalize(rng.standard_normal((B, D))).astype(np.float32)
    queue = O.normalize(rng.standard_normal((K, D))).astype(np.float32)
    # tile t (128 rows) gets rows strongly aligned with q_0 direction scaled up with t
    direction = q.mean(axis=0); direction /= np.linalg.norm(direction)
    for t in range(K // 128):
        queue[t * 128 + 5] = direction * (0.5 + 0.8 * t) + 0.01 * queue[t * 128 + 5]
    for splits in (1, 2):
        stats, Op, _, qb, kb = run_tc(lib, q, queue, 0.07, n_splits=splits)
        lse, Onorm, mx = merged(stats, Op)
        lse_o, O_o, mx_o = oracle_neg(qb, kb, 0.07)
        assert np.allclose(mx, mx_o, rtol=1e-5, atol=1e-4)
        assert np.abs(lse - lse_o).max() < TOL * np.abs(lse_o).max()
        assert rel(Onorm, O_o) < 5e-3


@pytest.mark.parametrize("B,D,K,T", [(32, 128, 4096, 0.15), (256, 128, 16384, 0.15), (512, 128, 65536, 0.15),
                                     (256, 256, 8192, 0.07), (256, 64, 4096, 0.07)])
def test_end_to_end_loss_and_grad(lib, B, D, K, T):
    """MoCo.forward in bf16 mode vs the oracle on bf16-rounded operands: loss and dq within 1e-3."""
    import moma_b200
    from moma_b200 import MoCo, LazyLogits
    moma_b200.set_precision("bf16")
    torch.manual_seed(B + K)
    m = MoCo(D, K, T).cuda()
    q = (torch.randn(B, D, device="cuda") * 0.6).requires_grad_()
    k = torch.randn(B, D, device="cuda") * 0.6
    mem0 = m.memory.clone()
    logits, labels = m(q, k)
    assert isinstance(logits, LazyLogits)
    loss = torch.nn.CrossEntropyLoss()(logits, labels)
    loss.backward()
    rb = lambda t: npy(t.detach().to(torch.bfloat16)).astype(np.float64)
    loss_o, rows_o, dq_o, pim_o = O.nce_loss_and_grad(rb(q), rb(k), rb(mem0), T)
    assert abs(loss.item() - loss_o) < TOL * abs(loss_o)
    assert rel(npy(q.grad), dq_o) < TOL
    # enqueue: fp32 master verbatim, shadow = bf16 rounding, pointer advanced
    assert torch.equal(m.memory[:B], k) and torch.equal(m.memory[B:], mem0[B:]) and m.index == B % K
    sh = m._shadow_of(m.memory, create=False)
    assert sh is not None and torch.equal(sh, m.memory.to(torch.bfloat16))
    # fp32 mode on the same inputs agrees with the bf16 mode to bf16 accuracy
    moma_b200.set_precision("fp32")
    m2 = MoCo(D, K, T).cuda(); m2.memory.copy_(mem0)
    q2 = q.detach().clone().requires_grad_()
    lg2, lab2 = m2(q2, k)
    l2 = torch.nn.CrossEntropyLoss()(lg2, lab2); l2.backward()
    moma_b200.set_precision("bf16")
    assert abs(l2.item() - loss.item()) < 2e-2 * abs(l2.item())
    assert rel(npy(q.grad), npy(q2.grad)) < 3e-2
