"""Momentum memory bank -- host-side mirror of the reference MoMA/mem_moco.py.

Same classes, constructor signatures, buffers (``memory`` / ``memory_s`` /
``memory_t``, fp32 ``[K, D]`` in ``state_dict``) and the ``index`` attribute, so
helper/loops_moma.py:331 and train_student_moma.py:333-336 use it unchanged.
The arithmetic is done by the sm_100a kernels behind include/moma_b200.h:

  * logits + CrossEntropy + d loss/d q : one fused pass (``ops.nce_rows``); the
    ``[B, K+1]`` logits are returned as a :class:`LazyLogits` handle and never
    written to HBM (reference: mem_moco.py:29-49 + contrast_trainer.py:189-205);
  * the per-step ``memory.clone()`` (mem_moco.py:89) is eliminated: kernels are
    stream-ordered, the loss pass reads the queue before the enqueue overwrites it;
  * ring enqueue + pointer (mem_moco.py:14-27): ``ops.enqueue`` (ids computed
    in-kernel, fp32 master + bf16 shadow written together).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .lazy_logits import LazyLogits


class BaseMoCo(nn.Module):
    """base class for MoCo-style memory cache (reference mem_moco.py:6-66)"""

    def __init__(self, K=65536, T=0.07):
        super().__init__()
        self.K = K
        self.T = T
        self.index = 0
        self._shadows = {}          # registered buffer name -> (bf16 shadow, version) ; non-persistent
        self._zero_labels = {}
        self._pending = []          # (queue data_ptr, LazyLogits state | ops.QueueGuard) awaiting the next enqueue
        self.checkpoint_pointer = False   # True: state_dict() also carries the ring pointer (key '<prefix>index')

    # ---- pointer / enqueue ------------------------------------------------
    def _update_pointer(self, bsz):
        # mem_moco.py:14-15
        self.index = (self.index + bsz) % self.K
        if self._index_dev is not None:         # device mirror, advanced on the stream (CUDA graphs)
            ops.pointer_advance(self._index_dev, bsz, self.K)

    _index_dev = None

    def use_device_pointer(self, device=None, enable=True):
        """Mirror ``index`` in a device-resident int64 that the enqueue kernel reads and a tiny
        kernel advances, so a captured CUDA graph of the step stays correct across replays.  The
        host ``index`` attribute keeps the reference's semantics (``replayed(n)`` re-syncs it)."""
        if not enable:
            self._index_dev = None
            return
        device = device or next(self.buffers()).device
        self._index_dev = torch.tensor([self.index], dtype=torch.int64, device=device)

    def replayed(self, n_rows):
        """Account on the host for one graph replay that enqueued ``n_rows`` rows."""
        self.index = (self.index + n_rows) % self.K

    def _buffer_name(self, queue: torch.Tensor):
        """Name of the registered buffer `queue` is (or is a detached alias of), else None."""
        for name, b in self._buffers.items():
            if b is not None and (b is queue or (b.data_ptr() == queue.data_ptr() and b.shape == queue.shape)):
                return name
        return None

    def _shadow_of(self, queue: torch.Tensor, create: bool = True):
        """bf16 shadow of a queue.  Shadows are cached only for REGISTERED buffers (by buffer name) and rebuilt
        when the fp32 master was replaced (``.cuda()``, ``load_state_dict``) or modified through autograd-visible
        in-place ops (``_version``); writers that bypass the version counter (``memory.data.copy_``,
        ``dist.broadcast``) must call ``invalidate_shadows()``.  A transient queue (e.g. the attended queue of
        MoCoAtt, a new tensor every step whose address the allocator recycles) is cast every time."""
        if not queue.is_cuda or queue.shape[1] % 8 != 0:
            return None
        name = self._buffer_name(queue)
        if name is None:
            if not create:
                return None
            sh = torch.empty(queue.shape, dtype=torch.bfloat16, device=queue.device)
            ops.cast_bf16(queue.contiguous(), sh)
            return sh
        ent = self._shadows.get(name)
        if ent is not None:
            sh, ptr, ver = ent
            if ptr == queue.data_ptr() and sh.device == queue.device and sh.shape == queue.shape:
                if ver != queue._version:
                    ops.cast_bf16(queue, sh)
                    self._shadows[name] = (sh, ptr, queue._version)
                return sh
        if not create:
            return None
        sh = torch.empty(queue.shape, dtype=torch.bfloat16, device=queue.device)
        ops.cast_bf16(queue, sh)
        self._shadows[name] = (sh, queue.data_ptr(), queue._version)
        return sh

    def invalidate_shadows(self):
        """Drop the bf16 shadows (call after writing a queue buffer behind autograd's back)."""
        self._shadows.clear()

    def _update_memory(self, k, queue):
        """queue[(index + j) % K] = k[j]   (mem_moco.py:17-27)"""
        with torch.no_grad():
            n = k.shape[0]
            if n > self.K:
                raise RuntimeError(f"enqueue of {n} rows into a queue of K={self.K}: duplicate ids "
                                   "(undefined in the reference, mem_moco.py:24-27)")
            shadow = self._shadow_of(queue, create=ops.get_precision() == "bf16")
            self._settle_pending(queue, shadow, n)
            ops.enqueue(k, queue, shadow, self.K, self.index, index_dev=self._index_dev)

    def _settle_pending(self, queue, shadow, n):
        """Before `queue` is overwritten: hand the n rows about to be replaced to everything that was computed
        from the pre-enqueue queue and may still need it -- lazy logits handles that allow late materialisation
        and the autograd nodes of dense logits (the reference's ``memory.clone()``, mem_moco.py:89, without
        copying K x D every step)."""
        ptr = queue.data_ptr()
        mine = [o for (p, o) in self._pending if p == ptr]
        self._pending = [(p, o) for (p, o) in self._pending if p != ptr]
        if not mine or n == 0:
            return
        ids = ops.enqueue_ids(n, self.index, self.K, queue.device, index_dev=self._index_dev)
        old32 = old16 = None
        for o in mine:
            if isinstance(o, ops.QueueGuard):
                if old32 is None:
                    old32 = queue.index_select(0, ids)
                o.remember(ids, old32)
            else:                                   # LazyLogits enqueue state
                if shadow is not None and ops.get_precision() == "bf16":
                    if old16 is None:
                        old16 = shadow.index_select(0, ids)
                    o["saved"] = (ids, old16)
                else:
                    if old32 is None:
                        old32 = queue.index_select(0, ids)
                    o["saved"] = (ids, old32)

    # ---- checkpointing of the ring pointer (the reference never saves it, train_student_moma.py:549-573)
    def _save_to_state_dict(self, destination, prefix, keep_vars):
        super()._save_to_state_dict(destination, prefix, keep_vars)
        if self.checkpoint_pointer:                  # opt-in: the default key set stays the reference's
            destination[prefix + "index"] = torch.tensor(int(self.index), dtype=torch.int64)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        key = prefix + "index"
        if key in state_dict:
            self.index = int(state_dict.pop(key)) % self.K
            if self._index_dev is not None:
                self._index_dev.fill_(self.index)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)
        self._shadows.clear()                        # the masters were just rewritten

    # ---- logits -----------------------------------------------------------
    def _compute_logit(self, q, k, queue):
        """Dense logits (mem_moco.py:29-49); used by the variants' fallbacks and tests.  When `queue` is one of
        this module's live buffers, the rows a later enqueue overwrites are handed to the autograd node first
        (``_settle_pending``), so ``backward()`` after the queue update still differentiates the logits that
        were returned (the reference clones the queue for this, :89)."""
        guard = None
        if torch.is_grad_enabled() and q.requires_grad:
            guard = ops.QueueGuard()                # a transient queue (never enqueued into) needs no copy either
            if self._buffer_name(queue) is not None:
                self._pending.append((queue.data_ptr(), guard))
        out = ops.nce_logits(q, k, queue, self.T, guard)
        return out.squeeze().contiguous()

    def _compute_logit_qk(self, q, k):
        """mem_moco.py:51-66"""
        return ops.nce_logits_qk(q, k, self.T).squeeze().contiguous()

    def _labels(self, bsz, device):
        key = (bsz, str(device))
        lab = self._zero_labels.get(key)
        if lab is None:
            lab = self._zero_labels[key] = torch.zeros(bsz, dtype=torch.long, device=device)
        return lab

    def _fused_logits(self, q, k, queue, labels):
        """Fused loss pass over `queue`; returns the lazy handle standing in for
        ``_compute_logit(q, k, queue.clone())``."""
        bsz, D = q.shape
        K = queue.shape[0]
        if not ops.fused_nce_supported(D):
            return self._compute_logit(q, k, queue)          # dense escape hatch (e.g. D = 1280, 2048)
        precision = ops.get_precision()
        shadow = self._shadow_of(queue) if precision == "bf16" and ops.bf16_supported(D) else None
        nce = ops.nce_rows(q, k, queue, shadow, self.T, precision)
        T = self.T
        # rows about to be overwritten by this step's enqueue: keep them so a late
        # materialisation still sees the pre-enqueue queue (the reference's clone, :89)
        state = {"index": self.index, "saved": None}
        q_d, k_d = q.detach(), k.detach()

        def materialize():
            src = shadow if shadow is not None else queue
            dense = ops.nce_logits(q, k_d, src, T)
            if state["saved"] is not None:
                ids, old = state["saved"]
                patch = ops.nce_logits(q, k_d, old, T)[:, 1:]
                dense = dense.index_copy(1, ids + 1, patch)
            return dense.squeeze().contiguous()

        shape = (bsz, K + 1) if bsz != 1 else (K + 1,)
        handle = LazyLogits(shape, q.device, nce, labels, materialize)
        handle._enqueue_state = state
        if self._buffer_name(queue) is not None:
            if getattr(self, "track_overwritten", False):
                self._pending.append((queue.data_ptr(), state))
            else:
                state["stale"] = True               # armed by _arm_stale() once the queue is updated
        return handle

    @staticmethod
    def _arm_stale(handles):
        """After the enqueue the dense logits of a handle that did not keep the overwritten rows are gone."""
        for h in handles:
            if isinstance(h, LazyLogits) and h._enqueue_state.get("stale") and h._dense is None:
                h._materialize = _stale_after_enqueue


class MoCo(BaseMoCo):
    """Single Modal (e.g., RGB) MoCo-style cache (reference mem_moco.py:69-100)"""

    def __init__(self, n_dim, K=65536, T=0.07, mem_name='memory'):
        super().__init__(K, T)
        # same RNG draw as the reference: K*n_dim normals from the global CPU generator, then L2 rows
        self.register_buffer(mem_name, torch.randn(K, n_dim))
        self.memory = F.normalize(self.memory)
        self.track_overwritten = False      # True: keep the n overwritten rows so the dense logits can still be
                                            # materialised after the enqueue (2 tiny extra launches per step)

    def forward(self, q, k, all_k=None, defer_enqueue=False):
        """
        Args:
          q: query on current node
          k: key on current node
          all_k: gather of feats across nodes; otherwise use k
          defer_enqueue: (extension) skip the queue update here; the caller runs ``enqueue(all_k)`` later in the
            step.  The loss of a step does not depend on the keys it enqueues, so a scheduler can take the branch
            that produces ``all_k`` (the queue attention) off the critical path.
        Returns (logits, labels) as mem_moco.py:77-100.
        """
        bsz = q.size(0)
        k = k.detach()
        labels = self._labels(bsz, q.device)
        logits = self._fused_logits(q, k, self.memory, labels)
        if defer_enqueue:
            self._arm_stale([logits])
            return logits, labels
        all_k = all_k if all_k is not None else k
        self.enqueue(all_k)
        self._arm_stale([logits])
        return logits, labels

    def enqueue(self, all_k):
        """The queue update of forward (reference mem_moco.py:14-27, :97-99) on its own."""
        self._update_memory(all_k, self.memory)
        self._update_pointer(all_k.size(0))


def _stale_after_enqueue():
    raise RuntimeError(
        "LazyLogits: the dense [B, K+1] logits were requested after the queue was updated. Only "
        "CrossEntropyLoss (zero labels) and top-1 accuracy are served without materialising; set "
        "`contrast.track_overwritten = True` (or MOMA_B200_LOGITS=dense) to allow other uses.")


class MoCoAtt(BaseMoCo):
    """MoCo cache with attention applied inside the memory (reference mem_moco.py:103-161)"""

    def __init__(self, n_dim, K=65536, T=0.07, mem_name='memory'):
        super().__init__(K, T)
        self.register_buffer(mem_name, torch.randn(K, n_dim))
        self.memory = F.normalize(self.memory)

    def forward(self, q, k, all_k=None, attn=None, criterion_kd=None):
        bsz = q.size(0)
        k = k.detach()
        # The reference clones the queue (:119).  Where the queue only feeds the fused loss pass, no copy is needed
        # (stream order: the pass reads it before the enqueue overwrites it).  Where it is an INPUT OF AN ATTENTION
        # MODULE under autograd ('all', 'dual' and the default mode), that module saves its input for the backward of
        # its weights, and the enqueue below would overwrite the saved rows behind autograd's back: copy it.
        queue = self.memory.detach()
        if torch.is_grad_enabled() and attn not in ('qk', 'dual2', 'self_qk', 'self_qkv2'):
            queue = queue.clone()
        if attn == 'all':
            out = criterion_kd.atts(torch.cat([q, k, queue], dim=0))
            q, k, queue = out[:bsz], out[bsz:2 * bsz], out[2 * bsz:]
        elif attn == 'qk':
            out = criterion_kd.atts(torch.cat([q, k], dim=0))
            q, k = out[:bsz], out[bsz:]
        elif attn == 'dual':
            out_p = criterion_kd.atts_p(torch.cat([q, queue], dim=0))
            q, queue = out_p[:bsz], out_p[bsz:]
            out_n = criterion_kd.atts_n(torch.cat([k, queue], dim=0))
            k, queue = out_n[:bsz], out_n[bsz:]
        elif attn == 'dual2':
            out_p = criterion_kd.atts_p(torch.cat([q, k], dim=0))
            q = out_p[:bsz]
            out_n = criterion_kd.atts_n(torch.cat([k, q], dim=0))
            k = out_n[:bsz]
        elif attn in ['self_qk', 'self_qkv2']:
            q = criterion_kd.atts_q(q)
            k = criterion_kd.atts_k(k)
        else:
            q = criterion_kd.atts_q(q)
            k = criterion_kd.atts_k(k)
            queue = criterion_kd.atts_queue(queue)

        labels = self._labels(bsz, q.device)
        if attn == 'dual2':
            logits = self._compute_logit_qk(q, k)
        elif torch.is_grad_enabled() and (k.requires_grad or queue.requires_grad):
            # the attended k / queue carry gradients into the attention parameters (the reference
            # only detaches k *before* the attention, :116): keep full autograd on the dense form.
            # autograd saves `queue` for backward and the enqueue below overwrites the live buffer
            # behind its back -> copy it first (the reference's clone, :119)
            if queue.data_ptr() == self.memory.data_ptr():
                queue = queue.clone()
            logits = _dense_logits_autograd(q, k, queue, self.T)
        else:
            logits = self._fused_logits(q, k, queue.contiguous(), labels)

        all_k = all_k if all_k is not None else k
        self._update_memory(all_k, self.memory)
        self._update_pointer(all_k.size(0))
        self._arm_stale([logits])
        return logits, labels


def _dense_logits_autograd(q, k, queue, T):
    """Reference formulation with full autograd (k / queue may carry gradients in the
    MoCoAtt concat modes, where the reference does not detach the attended tensors)."""
    bsz = q.shape[0]
    pos = (q * k).sum(1, keepdim=True)
    neg = q @ queue.t()
    return (torch.cat((pos, neg), dim=1) / T).squeeze().contiguous()


class _DualQueue(BaseMoCo):
    def __init__(self, n_dim, K=65536, T=0.07):
        super().__init__(K, T)
        self.register_buffer('memory_s', torch.randn(K, n_dim))
        self.register_buffer('memory_t', torch.randn(K, n_dim))
        self.memory_s = F.normalize(self.memory_s)
        self.memory_t = F.normalize(self.memory_t)

    def _finish(self, handles, k, k_t, all_k, all_k_t):
        all_k = all_k if all_k is not None else k
        all_k_t = all_k_t if all_k_t is not None else k_t
        self._update_memory(all_k, self.memory_s)
        self._update_memory(all_k_t, self.memory_t)
        self._update_pointer(all_k.size(0))
        self._arm_stale(handles)


class MoCoST(_DualQueue):
    """Two queues, student->student and student->teacher logits (reference mem_moco.py:165-204)"""

    def forward(self, q, k, k_t, all_k=None, all_k_t=None):
        bsz = q.size(0)
        k = k.detach()
        k_t = k_t.detach()
        labels = self._labels(bsz, q.device)
        logits_ss = self._fused_logits(q, k, self.memory_s, labels)
        logits_st = self._fused_logits(q, k_t, self.memory_t, labels)
        self._finish([logits_ss, logits_st], k, k_t, all_k, all_k_t)
        return logits_ss, logits_st, labels


class MoCoSSTT(_DualQueue):
    """Two queues, up to four logits sharing one pointer (reference mem_moco.py:208-253)"""

    def forward(self, q, k, q_t=None, k_t=None, all_k=None, all_k_t=None):
        bsz = q.size(0)
        k = k.detach()
        k_t = k_t.detach()
        labels = self._labels(bsz, q.device)
        logits_ss = self._fused_logits(q, k, self.memory_s, labels)
        logits_st = self._fused_logits(q, k_t, self.memory_t, labels)
        outs = [logits_ss, logits_st]
        if q_t is not None:
            outs.append(self._fused_logits(q_t, k, self.memory_s, labels))
            outs.append(self._fused_logits(q_t, k_t, self.memory_t, labels))
        self._finish(outs, k, k_t, all_k, all_k_t)
        return (*outs, labels)


def build_mem(opt):
    """Factory on opt.mem (reference mem_moco.py:256-273).  With ``opt.shard_queue`` (or
    MOMA_B200_SHARD_QUEUE=1) and an initialised process group of world size > 1, MoCo is
    built with its queue sharded by K across ranks (moma_b200.sharded.ShardedMoCo)."""
    import os
    if opt.mem == 'MoCoSSTT':
        return MoCoSSTT(opt.feat_dim, opt.nce_k, opt.nce_t)
    if opt.mem == 'MoCoST':
        return MoCoST(opt.feat_dim, opt.nce_k, opt.nce_t)
    if opt.mem == 'MoCoAtt':
        return MoCoAtt(opt.feat_dim, opt.nce_k, opt.nce_t)
    want_shard = bool(getattr(opt, "shard_queue", False)) or os.environ.get("MOMA_B200_SHARD_QUEUE") == "1"
    if want_shard:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            from .sharded import ShardedMoCo
            return ShardedMoCo(opt.feat_dim, opt.nce_k, opt.nce_t)
    return MoCo(opt.feat_dim, opt.nce_k, opt.nce_t)
