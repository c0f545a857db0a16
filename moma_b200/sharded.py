"""K-sharded momentum memory bank across the ranks of one node.

The reference replicates the whole ``[K, D]`` queue on every rank and each rank scans all K rows
(MoMA/mem_moco.py:97-99 with learning/contrast_trainer.py:124).  Here the queue rows are sharded
*cyclically*: global row ``g`` lives on rank ``g % W`` at local slot ``g // W`` (K % W == 0).  The
global ids are exactly the reference's ``(index + j) % K`` (bit-exact), and because a step's ids are
consecutive every rank receives ``n / W`` of the new rows (balanced enqueue).

Per step (all latency-bound, tiny payloads):
  1. all-gather of the queries ``[B_local, D]`` -> ``[n, D]`` (bf16 in bf16 mode) so each rank scores
     all ``n`` global queries against its ``K / W`` rows -- same FLOPs per rank as one GPU at B_local;
  2. fused partial pass over the local shard (+ local merge of the split partials);
  3. exchange: rank j receives the W partials ``(m, l, mmax, O)`` of its own ``B_local`` queries
     (all-to-all on NCCL; all-gather + slice where the backend lacks all-to-all, e.g. gloo on CPU tests);
  4. combine with the positive column (local q, k) -> loss rows, d loss/d q, top-1 flags;
  5. enqueue: every rank writes the rows it owns out of the all-gathered keys (the caller's
     ``_global_gather``, learning/contrast_trainer.py:83-88).
``state_dict()['memory']`` still presents the full ``[K, D]`` fp32 queue (gathered), see ``memory``.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import ops
from .lazy_logits import LazyLogits
from .peer import CH_PARTIALS, CH_QUERIES, PeerExchange
from .mem_moco import BaseMoCo, _stale_after_enqueue


def _peer_fused(world: int) -> bool:
    """Exchange fused into the merge / combine kernels (MOMA_B200_PEER_FUSED=1 / 0; default: only for 2 ranks).
    Measured on B200s, C3 weak: 2 GPUs 0.243 (fused) vs 0.251 ms per step; 8 GPUs 0.317 vs 0.285 -- with 8 ranks the records
    of one combine CTA come from 8 merge kernels at 8 different points of their own steps, and B polling CTAs wait where
    the separate exchange kernel waits with at most 37."""
    env = os.environ.get("MOMA_B200_PEER_FUSED")
    if env is not None:
        return env != "0"
    return world <= 2


def cyclic_shard(full: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Rows g with g % world == rank, in slot order g // world."""
    return full[rank::world].contiguous()


def cyclic_unshard(shards) -> torch.Tensor:
    world = len(shards)
    rows = shards[0].shape[0]
    out = shards[0].new_empty((rows * world,) + tuple(shards[0].shape[1:]))
    for r, s in enumerate(shards):
        out[r::world] = s
    return out


class ShardedMoCo(BaseMoCo):
    """Drop-in for MoCo (same ctor / forward signature) with the queue sharded by K over the
    default process group."""

    is_sharded = True
    after_query_gather = None      # optional callable, invoked right after the query all-gather has been enqueued

    def __init__(self, n_dim, K=65536, T=0.07, mem_name='memory', group=None):
        super().__init__(K, T)
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if K % self.world != 0:
            raise ValueError(f"sharded queue needs K % world == 0 (K={K}, world={self.world})")
        # identical RNG draw to the reference on every rank (same seed -> same queue), then keep our rows
        full = F.normalize(torch.randn(K, n_dim))
        self.n_dim = n_dim
        self.register_buffer("memory_shard", cyclic_shard(full, self.rank, self.world), persistent=False)

    # ---- full-queue views (collective: call on every rank) ------------------------------------
    def gather_full(self) -> torch.Tensor:
        return cyclic_unshard(list(self._all_gather(self.memory_shard).unbind(0)))

    def _all_gather(self, x: torch.Tensor) -> torch.Tensor:
        """[W, *x.shape]; the output is laid out as the dim-0 concatenation (accepted by NCCL and gloo)."""
        x = x.contiguous()
        out = torch.empty((self.world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x, group=self.group)
        return out.view((self.world,) + tuple(x.shape))

    @property
    def memory(self) -> torch.Tensor:
        """Full [K, D] fp32 queue in the reference's row order (collective)."""
        return self.gather_full()

    def set_full(self, full: torch.Tensor) -> None:
        with torch.no_grad():
            self.memory_shard.copy_(cyclic_shard(full.to(self.memory_shard.device), self.rank, self.world))
        self.invalidate_shadows()

    def broadcast_from_rank0(self) -> None:
        """ContrastTrainer.broadcast_memory for the sharded queue (reference :71-81)."""
        full = self.gather_full()
        # rank 0's view of every shard is authoritative only for its own rows; rebuild from rank 0's RNG draw
        dist.broadcast(full, 0, group=self.group)
        self.set_full(full)

    # ---- checkpointing ------------------------------------------------------------------------------
    # ``state_dict()`` is free of collectives (the usual ``if rank == 0: torch.save(contrast.state_dict())`` must not
    # deadlock): it carries this rank's rows under 'memory_shard' plus the layout (rank, world) and the ring pointer.
    # ``full_state_dict()`` is the explicit COLLECTIVE that presents the reference's key set ('memory' = full [K, D]).
    # ``load_state_dict`` accepts either form.
    def _save_to_state_dict(self, destination, prefix, keep_vars):
        super()._save_to_state_dict(destination, prefix, keep_vars)
        destination[prefix + "memory_shard"] = self.memory_shard if keep_vars else self.memory_shard.detach()
        destination[prefix + "shard_layout"] = torch.tensor([self.rank, self.world, int(self.index)], dtype=torch.int64)

    def full_state_dict(self):
        """COLLECTIVE (call on every rank): {'memory': full [K, D] fp32 queue in the reference's row order,
        'index': ring pointer} -- loadable by MoCo, by ShardedMoCo of any world size, and ('memory' alone) by
        the reference's MoCo."""
        return {"memory": self.gather_full(), "index": torch.tensor(int(self.index), dtype=torch.int64)}

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        full, shard, layout = prefix + "memory", prefix + "memory_shard", prefix + "shard_layout"
        if full in state_dict:
            self.set_full(state_dict.pop(full))
            state_dict.pop(shard, None)
            state_dict.pop(layout, None)
        elif shard in state_dict:
            lay = state_dict.pop(layout, None)
            if lay is not None:
                r, w, idx = (int(v) for v in lay)
                if (r, w) != (self.rank, self.world):
                    error_msgs.append(f"ShardedMoCo: checkpoint shard is rank {r} of {w}, this process is rank "
                                      f"{self.rank} of {self.world} (save with full_state_dict() to re-shard)")
                    state_dict.pop(shard)
                    return
                self.index = idx % self.K
                if self._index_dev is not None:
                    self._index_dev.fill_(self.index)
            self.memory_shard.copy_(state_dict.pop(shard).to(self.memory_shard.device))
        elif strict:
            missing_keys.append(shard)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)

    # ---- exchange of the per-rank partials ------------------------------------------------------
    def _exchange(self, packed: torch.Tensor) -> torch.Tensor:
        """packed [W(dst), B_local, D + 4] -> [W(src), B_local, D + 4] for this rank's queries."""
        backend = dist.get_backend(self.group)
        out = torch.empty_like(packed)
        if backend == "nccl":
            dist.all_to_all_single(out, packed, group=self.group)
            return out
        return self._all_gather(packed)[:, self.rank].contiguous()

    def owned_rows(self, n: int):
        """(start, stride, count) of the rows of this step's n all-gathered keys that this rank will
        enqueue: global id (index + j) % K is owned iff (index + j) % W == rank.  Constant across steps
        when n % W == 0 (the pointer advances by n)."""
        W = self.world
        if n % W != 0:
            raise ValueError("owned_rows needs n % world == 0")
        return (self.rank - self.index) % W, W, n // W

    def forward(self, q, k, all_k=None, owned_k=None, defer_enqueue=False):
        """As MoCo.forward.  ``owned_k`` (optional, instead of ``all_k``): only the rows
        ``owned_rows(n)`` of the step's keys, e.g. from ``Attention.forward_rows``.  ``defer_enqueue``: skip the
        queue update; the caller runs ``enqueue(all_k=... | owned_k=...)`` later in the step."""
        bsz, D = q.shape
        W = self.world
        k = k.detach()
        labels = self._labels(bsz, q.device)
        precision = ops.get_precision()
        use_bf16 = precision == "bf16" and ops.bf16_supported(D)
        shadow = self._shadow_of(self.memory_shard) if use_bf16 else None
        inv_T = 1.0 / self.T

        group, rank = self.group, self.rank
        memory_shard = self.memory_shard

        # NVLink peer-memory exchange (one kernel per collective) where available, torch.distributed otherwise
        peer = PeerExchange.create(group, q.device, 2 * W * bsz * (D + 4) * 4) if q.is_cuda else None

        def compute(q_, k_):
            # 1. all-gather the queries (every rank must contribute the same B_local)
            if peer is not None:
                q32, k32 = ops._f32c(q_.detach()), ops._f32c(k_.detach())
                dtype, rnd = (ops.BF16, True) if use_bf16 else (ops.F32, False)
                all_q = peer.allgather(q32, CH_QUERIES, to_bf16=use_bf16)      # bf16 cast fused into the push
            else:
                q_op, dtype, q32, k32, rnd = ops.nce_operands(q_, k_, "bf16" if use_bf16 else "fp32")
                all_q = torch.empty((W * bsz, D), dtype=q_op.dtype, device=q_.device)
                dist.all_gather_into_tensor(all_q, q_op.contiguous(), group=group)
            hook = self.after_query_gather
            if hook is not None:            # the step's other exchanges are sequenced behind this one (see step.py)
                hook()
            # 2. local pass over this rank's K / W rows for all n queries, then fold the splits
            queue = shadow if use_bf16 else memory_shard
            # MOMA_B200_NCE_FUSED=few: with few K-splits per query tile (many ranks -> many query tiles) the merge of the splits
            # runs in the tail of the tensor-core kernel instead of a separate merge launch.  Measured at 8 GPUs together
            # with the other schedule changes of round 2: not faster than the separate launch, so off by default
            few_splits = use_bf16 and os.environ.get("MOMA_B200_NCE_FUSED") == "few" \
                and ops.nce_num_splits(all_q.shape[0], D, queue.shape[0], ops.BF16) <= 8
            if use_bf16 and (few_splits or ops.nce_fused_enabled(all_q.shape[0], D, queue.shape[0])) \
                    and ops.nce_fused_supported(all_q.shape[0], D, queue.shape[0]):
                packed = ops.nce_fused_packed(all_q, queue, inv_T)       # pass + merge of the K-splits in ONE launch
            else:
                stats, Opart = ops.nce_partial(all_q, queue, inv_T, dtype)
                if peer is not None and D <= 512 and _peer_fused(W):
                    # 3 + 4 without an exchange launch: the merge kernel stores each record straight into its owner's
                    # receive region, the combine kernel polls its rows there
                    ops.nce_merge_push(stats, Opart, peer, CH_PARTIALS)
                    rows, dq, pim, mx, loss, acc = ops.nce_combine_poll(peer, CH_PARTIALS, q32, k32, inv_T, rnd, 1.0 / bsz)
                    return loss, rows, pim, mx, acc, dq
                packed = ops.nce_merge_packed(stats, Opart)             # [n, D + 4] = (O | m | l | mmax | pad)
            # 3. exchange: rows are ordered by owner rank, so the records route with one all-to-all
            if peer is not None:
                recv = peer.alltoall(packed.view(W, bsz, D + 4), CH_PARTIALS)
            else:
                recv = self._exchange(packed.view(W, bsz, D + 4))       # [W(src), bsz, D + 4]
            # 4. combine with the positive column (reads the receive buffer in place)
            rows, dq, pim, mx, loss, acc = ops.nce_combine_packed(recv, q32, k32, inv_T, rnd, 1.0 / bsz)
            return loss, rows, pim, mx, acc, dq

        nce = ops.nce_autograd(q, k, compute)
        shape = (bsz, self.K + 1) if bsz != 1 else (self.K + 1,)
        logits = LazyLogits(shape, q.device, nce, labels, _stale_after_enqueue)
        if not defer_enqueue:
            self.enqueue(all_k=all_k if all_k is not None else k, owned_k=owned_k)
        return logits, labels

    def enqueue(self, all_k=None, owned_k=None):
        """5. enqueue the rows this rank owns (from the full key list or from the owned rows only)."""
        W = self.world
        use_bf16 = ops.get_precision() == "bf16" and ops.bf16_supported(self.n_dim)
        with torch.no_grad():
            shadow = self._shadow_of(self.memory_shard) if use_bf16 else None
            sh = shadow if use_bf16 else self._shadow_of(self.memory_shard, create=False)
            if owned_k is not None:
                n = owned_k.shape[0] * W
                start, stride, count = self.owned_rows(n)
                ops.enqueue(owned_k, self.memory_shard, sh, self.K, self.index, rank=self.rank, world=W,
                            index_dev=self._index_dev, key_start=start, key_stride=stride)
            else:
                n = all_k.shape[0]
                if n > self.K:
                    raise RuntimeError("enqueue of more rows than K (duplicate ids)")
                ops.enqueue(all_k, self.memory_shard, sh, self.K, self.index, rank=self.rank, world=W,
                            index_dev=self._index_dev)
        self._update_pointer(n)
