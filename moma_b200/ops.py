"""torch-facing wrappers over the C ABI (include/moma_b200.h).

PyTorch is plumbing here: it owns device memory and streams; every arithmetic op
of the hot path is a kernel in libmoma_b200.so, called with raw pointers on the
current CUDA stream.  There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import BF16, F32, check

_PRECISION = os.environ.get("MOMA_B200_PRECISION", "bf16").lower()


def set_precision(mode: str) -> None:
    """'bf16' (default; tcgen05 tensor-core InfoNCE, 1e-3 parity) or 'fp32' (SIMT, 1e-5 parity)."""
    global _PRECISION
    mode = mode.lower()
    if mode not in ("bf16", "fp32"):
        raise ValueError("precision must be 'bf16' or 'fp32'")
    _PRECISION = mode


def get_precision() -> str:
    return _PRECISION


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _need_cuda(*ts: torch.Tensor) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("moma_b200: expected CUDA tensors (there is no CPU fallback); got a "
                               f"{t.device} tensor")


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


# ----------------------------------------------------------------------------- EMA
class EmaPlan:
    """Chunk table for one (model, model_ema) parameter list (built once, reused)."""

    def __init__(self, srcs: Sequence[torch.Tensor], dsts: Sequence[torch.Tensor]):
        lib = _lib.load()
        n = len(srcs)
        self.key = plan_key(srcs, dsts)
        numels = (ctypes.c_int64 * n)(*[int(d.numel()) for d in dsts])
        n_chunks, nbytes = ctypes.c_int64(0), ctypes.c_size_t(0)
        check(lib.moma_ema_plan_size(n, numels, ctypes.byref(n_chunks), ctypes.byref(nbytes)))
        self.n_chunks = n_chunks.value
        self.elements = sum(int(d.numel()) for d in dsts)
        host = torch.empty(max(nbytes.value, 32), dtype=torch.uint8).pin_memory() if torch.cuda.is_available() \
            else torch.empty(max(nbytes.value, 32), dtype=torch.uint8)
        sp = (ctypes.c_void_p * n)(*[s.data_ptr() for s in srcs])
        dp = (ctypes.c_void_p * n)(*[d.data_ptr() for d in dsts])
        check(lib.moma_ema_plan_fill(n, sp, dp, numels, host.data_ptr(), host.numel()))
        self.host_table = host
        self.table = host.to(dsts[0].device, non_blocking=False) if n else host

    def run(self, m: float) -> None:
        check(_lib.load().moma_ema_multi(_p(self.table), self.n_chunks, float(m), float(1 - m), _stream()))


def plan_key(srcs, dsts) -> Tuple:
    """The chunk table encodes exactly (source address, destination address, element count) per tensor, so this
    key is the table's full content: a parameter re-allocated at the same address with the same element count
    yields the same (still valid) table, any other change yields another key."""
    return tuple((s.data_ptr(), d.data_ptr(), d.numel()) for s, d in zip(srcs, dsts))


_EMA_PLANS = {}


def ema_update(srcs: Sequence[torch.Tensor], dsts: Sequence[torch.Tensor], m: float) -> None:
    """dst = dst*m + (1-m)*src for every pair, one launch.  Mirrors the error
    behaviour of learning/contrast_trainer.py:207-211 (shape mismatch raises)."""
    srcs, dsts = list(srcs), list(dsts)
    n = min(len(srcs), len(dsts))                      # zip() semantics of the reference
    srcs, dsts = srcs[:n], dsts[:n]
    if n == 0:
        return
    for s, d in zip(srcs, dsts):
        if s.shape != d.shape:
            raise RuntimeError(f"The size of tensor a {tuple(d.shape)} must match the size of tensor b "
                               f"{tuple(s.shape)} (momentum_update)")
        if s.dtype != torch.float32 or d.dtype != torch.float32:
            raise RuntimeError("moma_b200.ema_update: only float32 parameters are supported")
        if not (s.is_contiguous() and d.is_contiguous()):
            raise RuntimeError("moma_b200.ema_update: parameters must be contiguous")
    _need_cuda(*srcs, *dsts)
    key = plan_key(srcs, dsts)
    plan = _EMA_PLANS.get(key)
    if plan is None:
        if len(_EMA_PLANS) > 64:
            _EMA_PLANS.clear()
        plan = _EMA_PLANS[key] = EmaPlan(srcs, dsts)
    plan.run(m)


# ------------------------------------------------------------------------ Normalize
class _L2Norm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eps):
        xc = _f32c(x)
        y = torch.empty_like(xc)
        check(_lib.load().moma_l2norm_fwd(_p(xc), _p(y), xc.shape[0], xc.shape[1], eps, _stream()))
        ctx.save_for_backward(xc)
        ctx.eps = eps
        return y

    @staticmethod
    def backward(ctx, g):
        (xc,) = ctx.saved_tensors
        g = _f32c(g)
        dx = torch.empty_like(xc)
        check(_lib.load().moma_l2norm_bwd(_p(xc), _p(g), _p(dx), xc.shape[0], xc.shape[1], ctx.eps, _stream()))
        return dx, None


def l2_normalize(x: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """F.normalize(x, p=2, dim=1) for a 2-D CUDA tensor."""
    _need_cuda(x)
    if x.dim() != 2 or x.shape[1] % 4 != 0:
        raise RuntimeError("moma_b200.l2_normalize: expects [rows, D] with D % 4 == 0")
    return _L2Norm.apply(x, eps)


# -------------------------------------------------------------------------- enqueue
def cast_bf16(src: torch.Tensor, dst: torch.Tensor) -> None:
    check(_lib.load().moma_cast_bf16(_p(src), _p(dst), src.numel(), _stream()))


def enqueue(keys: torch.Tensor, queue: torch.Tensor, shadow: Optional[torch.Tensor], K: int, index: int,
            rank: int = 0, world: int = 1, normalize: bool = False, eps: float = 1e-12,
            index_dev: Optional[torch.Tensor] = None, key_start: int = 0, key_stride: int = 1) -> None:
    """queue[(index + j) % K] = keys[j]  (mem_moco.py:17-27), cyclically sharded when world > 1.
    With key_start / key_stride, keys[i] stands for row key_start + i * key_stride of the step's key list."""
    _need_cuda(keys, queue, shadow)
    keys = _f32c(keys.detach())
    if key_start != 0 or key_stride != 1:
        check(_lib.load().moma_enqueue_strided(_p(keys), keys.shape[0], keys.shape[1], _p(queue), _p(shadow), K,
                                               int(index), _p(index_dev), rank, world, int(key_start),
                                               int(key_stride), _stream()))
        return
    check(_lib.load().moma_enqueue(_p(keys), keys.shape[0], keys.shape[1], _p(queue), _p(shadow), K, int(index),
                                   _p(index_dev), rank, world, int(normalize), eps, _stream()))


def pointer_advance(index_dev: torch.Tensor, n: int, K: int) -> None:
    check(_lib.load().moma_pointer_advance(_p(index_dev), int(n), int(K), _stream()))


def enqueue_ids(n: int, index: int, K: int, device, index_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ids = fmod(arange(n) + index, K).long()  (mem_moco.py:24-26); index read from the device mirror when given."""
    out = torch.empty(n, dtype=torch.int64, device=device)
    check(_lib.load().moma_enqueue_ids(n, int(index), _p(index_dev), K, _p(out), _stream()))
    return out


# ---------------------------------------------------------------------------- InfoNCE
def nce_num_splits(B: int, D: int, K_local: int, dtype: int) -> int:
    return int(_lib.load().moma_nce_num_splits(B, D, K_local, dtype))


def nce_partial(q: torch.Tensor, queue: torch.Tensor, inv_T: float, dtype: int, n_splits: Optional[int] = None):
    """Per-split partials (m, l, mmax, O) of q against the local queue rows."""
    B, D = q.shape
    K_local = queue.shape[0]
    if n_splits is None:
        n_splits = nce_num_splits(B, D, K_local, dtype)
    dev = q.device
    stats = torch.empty((3, n_splits, B), dtype=torch.float32, device=dev)
    O = torch.empty((n_splits, B, D), dtype=torch.float32, device=dev)
    check(_lib.load().moma_nce_partial(_p(q), _p(queue), B, D, K_local, inv_T, dtype, n_splits,
                                       _p(stats[0]), _p(stats[1]), _p(stats[2]), _p(O), _stream()))
    return stats, O


def nce_combine(stats: torch.Tensor, O: torch.Tensor, q_f32: torch.Tensor, k_f32: torch.Tensor, inv_T: float,
                round_bf16: bool = False, dq_scale: float = 1.0, want_mean: bool = False):
    """Merge partials (+ positive column) ->
    (loss_rows [B], dq [B, D] (= dq_unit * dq_scale), pos_is_max [B] int32, max_logit [B],
     loss_mean [] and acc_pct [1] when want_mean)."""
    n_parts, B = stats.shape[1], stats.shape[2]
    D = O.shape[2]
    dev = q_f32.device
    rows = torch.empty(B, dtype=torch.float32, device=dev)
    dq = torch.empty((B, D), dtype=torch.float32, device=dev)
    pim = torch.empty(B, dtype=torch.int32, device=dev)
    mx = torch.empty(B, dtype=torch.float32, device=dev)
    fin = torch.empty(2, dtype=torch.float32, device=dev) if want_mean else None
    check(_lib.load().moma_nce_combine(_p(stats[0]), _p(stats[1]), _p(stats[2]), _p(O), n_parts, _p(q_f32),
                                       _p(k_f32), B, D, inv_T, int(round_bf16), float(dq_scale), _p(rows), _p(dq),
                                       _p(pim), _p(mx), _p(fin), None if fin is None else fin.data_ptr() + 4,
                                       _stream()))
    if want_mean:
        return rows, dq, pim, mx, fin[0], fin[1:2]
    return rows, dq, pim, mx


def nce_merge(stats: torch.Tensor, O: torch.Tensor):
    """Fold the split partials [3, S, B] / [S, B, D] into one partial per row: ([3, 1, B], [1, B, D])."""
    n_parts, B = stats.shape[1], stats.shape[2]
    D = O.shape[2]
    if n_parts == 1:
        return stats, O
    out_s = torch.empty((3, 1, B), dtype=torch.float32, device=O.device)
    out_O = torch.empty((1, B, D), dtype=torch.float32, device=O.device)
    check(_lib.load().moma_nce_merge(_p(stats[0]), _p(stats[1]), _p(stats[2]), _p(O), n_parts, B, D,
                                     _p(out_s[0]), _p(out_s[1]), _p(out_s[2]), _p(out_O), _stream()))
    return out_s, out_O


def nce_merge_packed(stats: torch.Tensor, O: torch.Tensor) -> torch.Tensor:
    """Fold split partials into one packed record per row: [B, D + 4] = (O | m | l | mmax | pad)."""
    n_parts, B = stats.shape[1], stats.shape[2]
    D = O.shape[2]
    packed = torch.empty((B, D + 4), dtype=torch.float32, device=O.device)
    check(_lib.load().moma_nce_merge_packed(_p(stats[0]), _p(stats[1]), _p(stats[2]), _p(O), n_parts, B, D,
                                            _p(packed), _stream()))
    return packed


def nce_combine_packed(packed: torch.Tensor, q_f32: torch.Tensor, k_f32: torch.Tensor, inv_T: float,
                       round_bf16: bool, dq_scale: float):
    """Combine packed records [n_parts, B, D + 4] (+ positive column) ->
    (rows, dq, pos_is_max, max_logit, loss_mean, acc_pct)."""
    n_parts, B, P = packed.shape
    D = P - 4
    dev = q_f32.device
    rows = torch.empty(B, dtype=torch.float32, device=dev)
    dq = torch.empty((B, D), dtype=torch.float32, device=dev)
    pim = torch.empty(B, dtype=torch.int32, device=dev)
    mx = torch.empty(B, dtype=torch.float32, device=dev)
    fin = torch.empty(2, dtype=torch.float32, device=dev)
    check(_lib.load().moma_nce_combine_packed(_p(packed), n_parts, _p(q_f32), _p(k_f32), B, D, inv_T,
                                              int(round_bf16), float(dq_scale), _p(rows), _p(dq), _p(pim), _p(mx),
                                              _p(fin), fin.data_ptr() + 4, _stream()))
    return rows, dq, pim, mx, fin[0], fin[1:2]


def nce_merge_push(stats: torch.Tensor, O: torch.Tensor, peer, channel: int) -> None:
    """K-sharded queue: fold the split partials of all W * B_local query rows and store each merged record straight into
    the receive region of the rank that owns the query (``peer``: moma_b200.peer.PeerExchange)."""
    n_parts, n = stats.shape[1], stats.shape[2]
    D = O.shape[2]
    check(_lib.load().moma_nce_merge_push(_p(stats[0]), _p(stats[1]), _p(stats[2]), _p(O), n_parts, n // peer.world, D,
                                          *peer.link(channel, 2 * n * (D + 4) * 4), _stream()))
    peer.note_use(channel)


def nce_combine_poll(peer, channel: int, q_f32: torch.Tensor, k_f32: torch.Tensor, inv_T: float, round_bf16: bool,
                     dq_scale: float):
    """Counterpart of ``nce_merge_push``: poll this rank's records from every rank, add the positive column ->
    (rows, dq, pos_is_max, max_logit, loss_mean, acc_pct) as ``nce_combine_packed``."""
    B, D = q_f32.shape
    dev = q_f32.device
    rows = torch.empty(B, dtype=torch.float32, device=dev)
    dq = torch.empty((B, D), dtype=torch.float32, device=dev)
    pim = torch.empty(B, dtype=torch.int32, device=dev)
    mx = torch.empty(B, dtype=torch.float32, device=dev)
    fin = torch.empty(2, dtype=torch.float32, device=dev)
    check(_lib.load().moma_nce_combine_poll(_p(q_f32), _p(k_f32), B, D, inv_T, int(round_bf16), float(dq_scale),
                                            *peer.link(channel, 2 * peer.world * B * (D + 4) * 4),
                                            _p(rows), _p(dq), _p(pim), _p(mx), _p(fin), fin.data_ptr() + 4, _stream()))
    peer.note_use(channel)
    return rows, dq, pim, mx, fin[0], fin[1:2]


# ---- one-launch InfoNCE (tcgen05 pass + in-kernel combine) ---------------------------------------------------------
_FUSE_COUNTERS = {}


def _fuse_counters(device) -> torch.Tensor:
    """A zeroed block of 256 uint32 ticket counters for one launch.  The kernel leaves its block zero; consecutive
    launches rotate through 16 blocks per device so that launches running concurrently on different streams (parallel
    graph branches, the dual-queue variants) never share one."""
    ent = _FUSE_COUNTERS.get(device)
    if ent is None:
        ent = _FUSE_COUNTERS[device] = [torch.zeros((16, 256), dtype=torch.int32, device=device), 0]
    ent[1] = (ent[1] + 1) % 16
    return ent[0][ent[1]]


def nce_fused_supported(B: int, D: int, K_local: int) -> bool:
    return os.environ.get("MOMA_B200_NCE_FUSED", "") != "never" and bool(_lib.load().moma_nce_fused_supported(B, D, K_local))


def nce_fused_enabled(B: int, D: int, K_local: int) -> bool:
    """MOMA_B200_NCE_FUSED=1 runs the whole InfoNCE pass as ONE launch (tcgen05 kernel with the in-kernel combine).
    Default off: measured on the B200 the step takes the same time either way (C2 0.171 / 0.171 ms, C3 0.214 / 0.212 ms
    fused / three launches) -- the combine is bound by L2 latency, and in the kernel's tail only 8 warps per SM work on it
    while the three-launch path spreads it over 512 small CTAs behind programmatic dependent launch -- and the lean kernel
    keeps the tensor-core pass free of that latency-bound tail (ncu: 35.6 us fused vs 20 us + 8 us)."""
    if os.environ.get("MOMA_B200_NCE_FUSED", "0") != "1":
        return False
    return bool(_lib.load().moma_nce_fused_supported(B, D, K_local))


def nce_fused(q_bf16: torch.Tensor, queue_bf16: torch.Tensor, q_f32: torch.Tensor, k_f32: torch.Tensor, inv_T: float,
              round_bf16: bool, dq_scale: float):
    """(rows, dq, pos_is_max, max_logit, loss_mean, acc_pct) from ONE kernel launch."""
    lib = _lib.load()
    B, D = q_bf16.shape
    K_local = queue_bf16.shape[0]
    dev = q_bf16.device
    nbytes = int(lib.moma_nce_fused_workspace_bytes(B, D, K_local))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    rows = torch.empty(B, dtype=torch.float32, device=dev)
    dq = torch.empty((B, D), dtype=torch.float32, device=dev)
    pim = torch.empty(B, dtype=torch.int32, device=dev)
    mx = torch.empty(B, dtype=torch.float32, device=dev)
    fin = torch.empty(2, dtype=torch.float32, device=dev)
    check(lib.moma_nce_fused(_p(q_bf16), _p(queue_bf16), _p(q_f32), _p(k_f32), B, D, K_local, inv_T, int(round_bf16),
                             float(dq_scale), _p(ws), nbytes, _p(_fuse_counters(dev)), _p(rows), _p(dq), _p(pim), _p(mx),
                             _p(fin), fin.data_ptr() + 4, _stream()))
    return rows, dq, pim, mx, fin[0], fin[1:2]


def nce_fused_packed(q_bf16: torch.Tensor, queue_bf16: torch.Tensor, inv_T: float) -> torch.Tensor:
    """Merged packed records [B, D + 4] = (O | m | l | mmax | pad) of q against the local queue rows, one launch."""
    lib = _lib.load()
    B, D = q_bf16.shape
    K_local = queue_bf16.shape[0]
    dev = q_bf16.device
    nbytes = int(lib.moma_nce_fused_workspace_bytes(B, D, K_local))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    packed = torch.empty((B, D + 4), dtype=torch.float32, device=dev)
    check(lib.moma_nce_fused_packed(_p(q_bf16), _p(queue_bf16), B, D, K_local, inv_T, _p(ws), nbytes,
                                    _p(_fuse_counters(dev)), _p(packed), _stream()))
    return packed


def nce_operands(q: torch.Tensor, k: torch.Tensor, precision: str):
    """Operands of the chosen mode: (q_op, dtype, q_f32, k_f32, round_bf16).  In bf16 mode the
    partial kernel reads a bf16 copy of q; the combine kernel rounds q / k inline (round_bf16)."""
    q32, k32 = _f32c(q.detach()), _f32c(k.detach())
    if precision == "bf16":
        cached = getattr(q, "_moma_bf16", None)       # left by ops.attention when q is its (unmodified) output
        if cached is not None and cached[1] == q._version and cached[0].shape == q32.shape and q.dtype == torch.float32:
            return cached[0], BF16, q32, k32, True
        qb = torch.empty(q32.shape, dtype=torch.bfloat16, device=q32.device)
        cast_bf16(q32, qb)
        return qb, BF16, q32, k32, True
    return q32, F32, q32, k32, False


class NceOut:
    """Results of one fused InfoNCE pass."""
    __slots__ = ("loss", "rows", "pos_is_max", "max_logit", "acc")

    def __init__(self, loss, rows, pos_is_max, max_logit, acc):
        self.loss, self.rows, self.pos_is_max, self.max_logit, self.acc = loss, rows, pos_is_max, max_logit, acc


class _NceFused(torch.autograd.Function):
    """loss = mean_i(LSE_i - l_i0) (and the per-row terms) with, from the SAME pass over the queue,
    d loss / d q.  ``compute(q, k)`` runs the kernels (single GPU or sharded) and returns
    (loss_mean, rows, pos_is_max, max_logit, acc_pct, dq_mean) with dq_mean = d loss_mean / d q."""

    @staticmethod
    def forward(ctx, q, k, compute):
        loss, rows, pim, mx, acc, dq_mean = compute(q, k)
        ctx.save_for_backward(dq_mean)
        ctx.q_dtype, ctx.B = q.dtype, q.shape[0]
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(pim, mx, acc)
        return loss, rows, pim, mx, acc

    @staticmethod
    def backward(ctx, g_loss, g_rows, *unused):
        (dq_mean,) = ctx.saved_tensors
        grad = None
        if g_loss is not None:
            if (g_loss.is_cuda and g_loss.numel() == 1 and g_loss.dtype == torch.float32 and dq_mean.is_contiguous()
                    and not torch.is_grad_enabled()):
                grad = torch.empty_like(dq_mean)          # one PDL-chained launch instead of a broadcasting torch multiply
                check(_lib.load().moma_scale_by_scalar(_p(dq_mean), _p(g_loss), _p(grad), dq_mean.numel(), _stream()))
            else:
                grad = dq_mean * g_loss
        if g_rows is not None:
            gr = (g_rows.unsqueeze(1) * dq_mean) * float(ctx.B)
            grad = gr if grad is None else grad + gr
        return (None if grad is None else grad.to(ctx.q_dtype)), None, None


def nce_autograd(q: torch.Tensor, k: torch.Tensor, compute) -> NceOut:
    return NceOut(*_NceFused.apply(q, k.detach(), compute))


def nce_rows(q: torch.Tensor, k: torch.Tensor, queue_f32: torch.Tensor, queue_bf16: Optional[torch.Tensor],
             T: float, precision: Optional[str] = None) -> NceOut:
    """Single-GPU fused InfoNCE over a replicated queue."""
    _need_cuda(q, k, queue_f32)
    precision = precision or _PRECISION
    if precision == "bf16" and (queue_bf16 is None or not bf16_supported(q.shape[1])):
        precision = "fp32"
    inv_T = 1.0 / float(T)

    def compute(q_, k_):
        q_op, dtype, q32, k32, rnd = nce_operands(q_, k_, precision)
        queue = queue_bf16 if dtype == BF16 else queue_f32
        if dtype == BF16 and nce_fused_enabled(q_op.shape[0], q_op.shape[1], queue.shape[0]):
            rows, dq, pim, mx, loss, acc = nce_fused(q_op, queue, q32, k32, inv_T, rnd, 1.0 / q_.shape[0])
            return loss, rows, pim, mx, acc, dq
        stats, O = nce_partial(q_op, queue, inv_T, dtype)
        rows, dq, pim, mx, loss, acc = nce_combine(stats, O, q32, k32, inv_T, rnd, 1.0 / q_.shape[0], want_mean=True)
        return loss, rows, pim, mx, acc, dq

    return nce_autograd(q, k, compute)


def bf16_supported(D: int) -> bool:
    """Shapes the tcgen05 kernel handles (see csrc/nce_tc.cu)."""
    return bool(_lib.load().moma_has_tcgen05()) and D in (64, 128, 256)


def fused_nce_supported(D: int) -> bool:
    return D % 4 == 0 and D <= 512


class QueueGuard:
    """Rows of the queue that an enqueue overwrites between a dense-logits forward and its backward.

    The dense path saves the LIVE queue for backward (no K x D clone per step, unlike mem_moco.py:89); the
    enqueue kernel then overwrites n of its rows through a raw pointer, invisible to autograd's version
    counter.  The module that runs the enqueue records (ids, old rows) here first; backward then computes
    d loss/d q against the rows the logits were actually computed from."""
    __slots__ = ("ids", "old")

    def __init__(self):
        self.ids, self.old = None, None

    def remember(self, ids: torch.Tensor, old_rows: torch.Tensor) -> None:
        self.ids, self.old = ids, old_rows


class _NceLogits(torch.autograd.Function):
    """Dense logits [B, K+1] (mem_moco.py:29-49); backward = dense dlogits -> dq (k, queue detached)."""

    @staticmethod
    def forward(ctx, q, k, queue, T, guard):
        q32, k32 = _f32c(q), _f32c(k)
        B, D = q32.shape
        K = queue.shape[0]
        out = torch.empty((B, K + 1), dtype=torch.float32, device=q.device)
        dtype = BF16 if queue.dtype == torch.bfloat16 else F32
        if dtype == BF16:
            qo, ko = q32.to(torch.bfloat16), k32.to(torch.bfloat16)
        else:
            qo, ko = q32, k32
        check(_lib.load().moma_nce_logits(_p(qo), _p(ko), _p(queue), B, D, K, T, dtype, _p(out), _stream()))
        if ctx.needs_input_grad[0]:
            # without a guard the queue is copied (the reference's clone, mem_moco.py:89): the caller may
            # overwrite it before backward runs
            ctx.save_for_backward(k32, queue if guard is not None else queue.clone())
        ctx.T, ctx.guard = T, guard
        return out

    @staticmethod
    def backward(ctx, g):
        k32, queue = ctx.saved_tensors
        g = g / ctx.T
        dq = g[:, :1] * k32 + g[:, 1:] @ queue.float()      # escape hatch only: library GEMM
        guard = ctx.guard
        if guard is not None and guard.ids is not None:
            # the enqueue replaced rows `ids` after the forward: swap their contribution for the old rows'
            cols = g.index_select(1, guard.ids + 1)
            dq = dq + cols @ (guard.old.float() - queue.index_select(0, guard.ids).float())
        return dq, None, None, None, None


def nce_logits(q, k, queue, T, guard: Optional[QueueGuard] = None) -> torch.Tensor:
    """Dense logits.  ``guard``: see QueueGuard -- pass one (and fill it before the enqueue) to avoid the
    queue copy that otherwise protects the backward from a later in-place queue update."""
    _need_cuda(q, k, queue)
    return _NceLogits.apply(q, k.detach(), queue.detach(), float(T), guard)


def nce_logits_qk(q, k, T) -> torch.Tensor:
    """mem_moco.py:51-66 (positives only, [B]).  Tiny; when either operand carries gradients (MoCoAtt 'dual2': the
    attended k is NOT detached in the reference, :116 detaches before the attention) it stays on autograd ops."""
    _need_cuda(q, k)
    if torch.is_grad_enabled() and (q.requires_grad or k.requires_grad):
        return (_f32c(q) * _f32c(k)).sum(1) / T
    q32, k32 = _f32c(q.detach()), _f32c(k.detach())
    out = torch.empty(q32.shape[0], dtype=torch.float32, device=q.device)
    check(_lib.load().moma_nce_logits_qk(_p(q32), _p(k32), q32.shape[0], q32.shape[1], float(T), _p(out), _stream()))
    return out


# -------------------------------------------------------------------------- attention
class _Attention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w_qkv, b_qkv, w_proj, b_proj, H, want_probs, want_bf16=False):
        lib = _lib.load()
        xc, wq, wp, bp = _f32c(x), _f32c(w_qkv), _f32c(w_proj), _f32c(b_proj)
        bq = _f32c(b_qkv) if b_qkv is not None else None
        N, C = xc.shape
        dev = xc.device
        y = torch.empty((N, C), dtype=torch.float32, device=dev)
        qkv = torch.empty((N, 3 * C), dtype=torch.float32, device=dev)
        o = torch.empty((N, C), dtype=torch.float32, device=dev)
        lse = torch.empty((H, N), dtype=torch.float32, device=dev)
        probs = torch.empty((1, H, N, N), dtype=torch.float32, device=dev) if want_probs else None
        y16 = torch.empty((N, C), dtype=torch.bfloat16, device=dev) if want_bf16 else None
        check(lib.moma_attn_fwd(_p(xc), _p(wq), _p(bq), _p(wp), _p(bp), N, C, H, _p(y), _p(qkv), _p(o), _p(lse),
                                _p(probs), _p(y16), _stream()))
        ctx.save_for_backward(xc, wq, wp, qkv, o, lse)
        ctx.H = H
        ctx.has_bq = b_qkv is not None
        ctx.set_materialize_grads(False)      # no zero-fill kernel for the (non-differentiable) second output's gradient
        if want_probs:
            ctx.mark_non_differentiable(probs)
            return y, probs
        if want_bf16:
            ctx.mark_non_differentiable(y16)
            return y, y16
        return y

    @staticmethod
    def backward(ctx, gy, *unused):
        lib = _lib.load()
        xc, wq, wp, qkv, o, lse = ctx.saved_tensors
        N, C = xc.shape
        H = ctx.H
        if gy is None:                        # the output was not used downstream
            return (None,) * 8
        gy = _f32c(gy)
        need = ctx.needs_input_grad
        dev = xc.device
        gx = torch.empty_like(xc) if need[0] else None
        gwq = torch.empty_like(wq) if need[1] else None
        gbq = torch.empty(3 * C, dtype=torch.float32, device=dev) if (need[2] and ctx.has_bq) else None
        gwp = torch.empty_like(wp) if need[3] else None
        gbp = torch.empty(C, dtype=torch.float32, device=dev) if need[4] else None
        ws_bytes = int(lib.moma_attn_bwd_workspace_bytes(N, C, H))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        check(lib.moma_attn_bwd(_p(xc), _p(wq), _p(wp), _p(qkv), _p(o), _p(lse), _p(gy), N, C, H, _p(gx), _p(gwq),
                                _p(gbq), _p(gwp), _p(gbp), _p(ws), ws_bytes, _stream()))
        return gx, gwq, gbq, gwp, gbp, None, None, None


def attention(x, w_qkv, b_qkv, w_proj, b_proj, num_heads: int, want_probs: bool = False):
    """Attention.forward (criterion_moco_att.py:153-167) on x [N, C]."""
    _need_cuda(x, w_qkv, w_proj)
    if x.dim() != 2:
        raise RuntimeError("moma_b200.attention: expects [N, C]")
    if not want_probs and _PRECISION == "bf16" and bf16_supported(x.shape[1]):
        # the projection's epilogue also emits y in bf16: if y becomes the query of the InfoNCE pass, its tensor-core
        # operand is already there (nce_operands) and no cast kernel is launched
        y, y16 = _Attention.apply(x, w_qkv, b_qkv, w_proj, b_proj, int(num_heads), False, True)
        y._moma_bf16 = (y16, y._version)
        return y
    return _Attention.apply(x, w_qkv, b_qkv, w_proj, b_proj, int(num_heads), bool(want_probs))


def attention_rows(x, w_qkv, b_qkv, w_proj, b_proj, num_heads: int, q_start: int, q_stride: int, q_count: int):
    """Attention output for the query rows q_start + i * q_stride only (keys / values from all rows of x).
    Forward only: used for the keys a rank of the K-sharded queue will enqueue."""
    _need_cuda(x, w_qkv, w_proj)
    with torch.no_grad():
        xc, wq, wp, bp = _f32c(x), _f32c(w_qkv), _f32c(w_proj), _f32c(b_proj)
        bq = _f32c(b_qkv) if b_qkv is not None else None
        N, C = xc.shape
        dev = xc.device
        y = torch.empty((q_count, C), dtype=torch.float32, device=dev)
        qkv = torch.empty((N, 3 * C), dtype=torch.float32, device=dev)
        o = torch.empty((q_count, C), dtype=torch.float32, device=dev)
        lse = torch.empty((num_heads, q_count), dtype=torch.float32, device=dev)
        check(_lib.load().moma_attn_fwd_rows(_p(xc), _p(wq), _p(bq), _p(wp), _p(bp), N, C, int(num_heads),
                                             int(q_start), int(q_stride), int(q_count), _p(y), _p(qkv), _p(o),
                                             _p(lse), _stream()))
    return y


def attention_rows_from_qkv(qkv, w_proj, b_proj, num_heads: int, q_start: int, q_stride: int, q_count: int):
    """As attention_rows, from precomputed projections qkv [N, 3C] of all N tokens (forward only)."""
    _need_cuda(qkv, w_proj)
    with torch.no_grad():
        qc, wp, bp = _f32c(qkv), _f32c(w_proj), _f32c(b_proj)
        N, C = qc.shape[0], qc.shape[1] // 3
        dev = qc.device
        y = torch.empty((q_count, C), dtype=torch.float32, device=dev)
        o = torch.empty((q_count, C), dtype=torch.float32, device=dev)
        lse = torch.empty((num_heads, q_count), dtype=torch.float32, device=dev)
        check(_lib.load().moma_attn_fwd_rows(None, None, None, _p(wp), _p(bp), N, C, int(num_heads), int(q_start),
                                             int(q_stride), int(q_count), _p(y), _p(qc), _p(o), _p(lse), _stream()))
    return y


def attention_supported(C: int, H: int) -> bool:
    return C % H == 0 and (C // H) in (8, 16, 32, 64, 128)


# -------------------------------------------------------------------------- projection-head Linear (+ReLU)
def _linear_workspace(owner: torch.Tensor, key, nbytes: int, device):
    """Zeroed split-K workspace of one call site (a layer's forward or backward).  The kernels leave the ticket
    counters zero, so it is zeroed exactly once; a call site never runs concurrently with itself.  The workspace
    lives ON the layer's weight tensor (attribute ``_moma_ws``), so its lifetime is the layer's: nothing is keyed by
    ``id()`` / addresses that a later allocation could reuse."""
    table = getattr(owner, "_moma_ws", None)
    if table is None:
        table = {}
        try:
            owner._moma_ws = table
        except AttributeError:                      # exotic tensor subclass: fall back to a fresh workspace per call
            return torch.zeros(max(nbytes, 16), dtype=torch.uint8, device=device)
    ws = table.get(key)
    if ws is None or ws.numel() < nbytes or ws.device != device:
        ws = table[key] = torch.zeros(max(nbytes, 16), dtype=torch.uint8, device=device)
    return ws


class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, relu, key):
        lib = _lib.load()
        xc, wc = _f32c(x), _f32c(w)
        bc = _f32c(b) if b is not None else None
        M, K = xc.shape
        N = wc.shape[0]
        y = torch.empty((M, N), dtype=torch.float32, device=xc.device)
        nbytes = int(lib.moma_linear_workspace_bytes(M, N, K))
        ws = _linear_workspace(w, (key, "f", M, N, K), nbytes, xc.device)
        check(lib.moma_linear_fwd(_p(xc), _p(wc), _p(bc), M, N, K, int(relu), _p(y), _p(ws), nbytes, _stream()))
        ctx.save_for_backward(xc, wc, y if relu else None)
        ctx.relu, ctx.key, ctx.has_b, ctx.owner = bool(relu), key, b is not None, w
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = _lib.load()
        xc, wc, y = ctx.saved_tensors
        M, K = xc.shape
        N = wc.shape[0]
        gy = _f32c(gy)
        need = ctx.needs_input_grad
        gx = torch.empty_like(xc) if need[0] else None
        gw = torch.empty_like(wc) if need[1] else None
        gb = torch.empty(N, dtype=torch.float32, device=xc.device) if (need[2] and ctx.has_b) else None
        nbytes = int(lib.moma_linear_workspace_bytes(M, N, K))
        ws = _linear_workspace(ctx.owner, (ctx.key, "b", M, N, K), nbytes, xc.device)
        check(lib.moma_linear_bwd(_p(xc), _p(wc), _p(y), _p(gy), M, N, K, int(ctx.relu), _p(gx), _p(gw), _p(gb),
                                  _p(ws), nbytes, _stream()))
        return gx, gw, gb, None, None


def linear(x, weight, bias=None, relu: bool = False, key=None):
    """nn.Linear (+ nn.ReLU) of the projection heads (criterion_moco_att.py:254-305) on x [M, K]:
    3xTF32 tensor-core GEMM with the bias / ReLU epilogue fused; autograd through the same kernels."""
    _need_cuda(x, weight)
    if x.dim() != 2 or weight.dim() != 2 or x.shape[1] != weight.shape[1]:
        raise RuntimeError(f"moma_b200.linear: bad shapes x {tuple(x.shape)} weight {tuple(weight.shape)}")
    return _Linear.apply(x, weight, bias, bool(relu), key if key is not None else "linear")


# -------------------------------------------------------------------------- classification CE + KD + top-1
class _ClsKd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logit_s, logit_t, labels, T):
        ls, lt = _f32c(logit_s), _f32c(logit_t.detach())
        B, C = ls.shape
        out = torch.empty(3, dtype=torch.float32, device=ls.device)
        g_cls, g_div = torch.empty_like(ls), torch.empty_like(ls)
        check(_lib.load().moma_cls_kd(_p(ls), _p(lt), _p(labels.contiguous()), B, C, float(T), _p(out), _p(g_cls), _p(g_div),
                                      _stream()))
        ctx.save_for_backward(g_cls, g_div)
        ctx.set_materialize_grads(False)
        acc = out[2:3]
        ctx.mark_non_differentiable(acc)
        return out[0], out[1], acc

    @staticmethod
    def backward(ctx, g1, g2, *unused):
        g_cls, g_div = ctx.saved_tensors
        grad = None
        if g1 is not None:
            grad = g_cls * g1
        if g2 is not None:
            grad = g_div * g2 if grad is None else torch.addcmul(grad, g_div, g2)
        return grad, None, None, None


def cls_kd_losses(logit_s, logit_t, labels, T: float):
    """(loss_cls, loss_div, acc_top1 [1]) = (CrossEntropyLoss()(logit_s, labels), DistillKL(T)(logit_s, logit_t),
    accuracy(logit_s, labels)[0]) from one kernel launch; differentiable with respect to logit_s
    (helper/loops_moma.py:278-279,350 / distiller_zoo/KD.py:7-17 / helper/util.py:71-85)."""
    _need_cuda(logit_s, logit_t, labels)
    if logit_s.dim() != 2 or logit_s.shape != logit_t.shape or labels.dtype != torch.int64 or labels.shape[0] != logit_s.shape[0]:
        raise RuntimeError("moma_b200.cls_kd_losses: expects logit_s, logit_t [B, n_cls] and int64 labels [B]")
    return _ClsKd.apply(logit_s, logit_t, labels, float(T))


# -------------------------------------------------------------------------- SGD + EMA in one pass
class FusedSgdEma:
    """``torch.optim.SGD(params, lr, momentum, weight_decay).step()`` followed by
    ``ContrastTrainer.momentum_update(model, model_ema, m)`` as ONE multi-tensor launch
    (train_student_moma.py:389-392, helper/loops_moma.py:361 -> :309; same rounding sequence as the two library steps).

        opt = FusedSgdEma(model.parameters(), model_ema.parameters(), lr=0.05, momentum=0.9, weight_decay=5e-4, m=0.999)
        loss.backward(); opt.step()            # instead of optimizer.step() ... trainer.momentum_update(...)

    Parameters without a gradient are skipped by the SGD part of torch; here every parameter must have a ``.grad``.
    The pointer table is rebuilt only when a gradient tensor was re-allocated (``zero_grad(set_to_none=True)``)."""

    def __init__(self, params, ema_params, lr, momentum=0.0, weight_decay=0.0, m=0.999):
        self.params, self.emas = list(params), list(ema_params)
        if len(self.params) != len(self.emas):
            raise RuntimeError("FusedSgdEma: parameter lists differ in length")
        for p, e in zip(self.params, self.emas):
            if p.shape != e.shape:
                raise RuntimeError(f"The size of tensor a {tuple(e.shape)} must match the size of tensor b {tuple(p.shape)}")
            if p.dtype != torch.float32 or e.dtype != torch.float32 or not (p.is_contiguous() and e.is_contiguous()):
                raise RuntimeError("FusedSgdEma: contiguous float32 parameters only")
        _need_cuda(*self.params, *self.emas)
        self.lr, self.momentum, self.weight_decay, self.m = float(lr), float(momentum), float(weight_decay), float(m)
        self.bufs = [torch.zeros_like(p) for p in self.params]
        self.steps = 0
        self._key, self._table, self._n_chunks = None, None, 0

    def _plan(self):
        grads = []
        for p in self.params:
            if p.grad is None:
                raise RuntimeError("FusedSgdEma.step: a parameter has no gradient")
            g = p.grad
            if g.dtype != torch.float32 or not g.is_contiguous():
                raise RuntimeError("FusedSgdEma.step: contiguous float32 gradients only")
            grads.append(g)
        key = tuple(g.data_ptr() for g in grads)
        if key == self._key:
            return
        lib = _lib.load()
        n = len(self.params)
        numels = (ctypes.c_int64 * n)(*[int(p.numel()) for p in self.params])
        n_chunks, nbytes = ctypes.c_int64(0), ctypes.c_size_t(0)
        check(lib.moma_sgd_ema_plan_size(n, numels, ctypes.byref(n_chunks), ctypes.byref(nbytes)))
        host = torch.empty(max(nbytes.value, 64), dtype=torch.uint8).pin_memory()
        arr = lambda ts: (ctypes.c_void_p * n)(*[t.data_ptr() for t in ts])
        check(lib.moma_sgd_ema_plan_fill(n, arr([p.detach() for p in self.params]), arr(grads), arr(self.bufs),
                                         arr([e.detach() for e in self.emas]), numels, host.data_ptr(), host.numel()))
        self._host = host
        self._table = host.to(self.params[0].device, non_blocking=False)
        self._n_chunks, self._key = n_chunks.value, key

    @torch.no_grad()
    def step(self):
        self._plan()
        check(_lib.load().moma_sgd_ema_multi(_p(self._table), self._n_chunks, self.lr, self.momentum, self.weight_decay,
                                             int(self.steps == 0), self.m, float(1 - self.m), _stream()))
        self.steps += 1

    def zero_grad(self, set_to_none: bool = False):
        for p in self.params:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()
