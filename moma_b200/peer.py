"""Peer-memory exchange (NVLink / NVSwitch) for the K-sharded queue -- host side of csrc/peer.cu.

One symmetric allocation per process group (torch.distributed._symmetric_memory: every rank maps every
peer's copy), a control block plus two receive regions per channel.  ``allgather`` / ``alltoall`` are
ONE kernel launch each (push of (word, epoch-tag) pairs + poll + copy-out) and are CUDA-graph capturable: the epoch that
sequences the calls lives in device memory.  Where symmetric memory is unavailable (CPU / gloo tests,
no peer access) ``PeerExchange.create`` returns None and the callers use torch.distributed collectives.
"""
from __future__ import annotations

import os
import warnings
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check

# One channel per CALL SITE: a channel carries one epoch sequence, so two call sites that could run on different
# streams (the step's branches) must not share one.  CH_KEYS = ContrastTrainer._global_gather (keys / qkv projections).
CH_QUERIES, CH_PARTIALS, CH_KEYS, CH_SPARE = 0, 1, 2, 3
_N_CHANNELS = 4


class PeerExchange:
    _cache = {}

    def __init__(self, group, device, region_bytes: int):
        import torch.distributed._symmetric_memory as symm_mem
        lib = _lib.load()
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.device = device
        self.ctrl_bytes = int(lib.moma_peer_ctrl_bytes())
        self.region_bytes = (max(int(region_bytes), 32 << 20) + 255) // 256 * 256    # roomy: re-creation is collective and
        # must not happen while a captured graph still holds the old buffer's addresses
        total = self.ctrl_bytes + 2 * _N_CHANNELS * self.region_bytes
        self.buf = symm_mem.empty(total, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, self.group)
        torch.cuda.synchronize(device)
        self.handle.barrier()                       # every rank's control block is zero before anyone pushes
        self.bases_dev = int(self.handle.buffer_ptrs_dev)
        self._last_use = {}                         # channel -> (stream id, event after its last call)

    # ---------------------------------------------------------------- construction
    @classmethod
    def create(cls, group, device, region_bytes: int) -> Optional["PeerExchange"]:
        """The shared instance for (group, device), grown on demand; None when peer memory is unavailable.

        COLLECTIVE on first use (and on growth): every rank of the group must call it at the same point.  The
        ranks AGREE on the outcome (all-reduce of a success flag), so a rendezvous that fails on one rank cannot
        leave the group on two different code paths; the fallback to torch.distributed collectives is logged
        once (``PeerExchange.status()`` / the bench JSON expose it).  Only intra-node groups qualify: the kernel
        stores straight into peer-mapped memory over NVLink."""
        if device.type != "cuda" or not dist.is_initialized():
            return None
        group = group if group is not None else dist.group.WORLD
        key = (id(group), device.index)
        inst = cls._cache.get(key)
        if inst is False:
            return None
        if inst is not None and inst.region_bytes >= region_bytes:
            return inst
        why = None
        if dist.get_backend(group) != "nccl":
            why = f"backend {dist.get_backend(group)} (needs nccl)"
        elif dist.get_world_size(group) > 16:
            why = "group larger than 16 ranks"
        elif os.environ.get("MOMA_B200_PEER", "1") == "0":
            why = "disabled by MOMA_B200_PEER=0"
        elif not cls._single_node(group):
            why = "group spans several nodes"
        if why is not None:                              # deterministic on every rank: no agreement round needed
            cls._fallback(key, why)
            return None
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("PeerExchange: the symmetric buffer must be created before CUDA-graph capture "
                               "(run one eager step first)")
        err = None
        try:
            inst = cls(group, device, region_bytes)
        except Exception as e:                           # no symmetric-memory support on this system
            inst, err = None, f"{type(e).__name__}: {e}"
        ok = torch.tensor([1 if inst is not None else 0], dtype=torch.int32, device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            cls._fallback(key, err or "symmetric-memory rendezvous failed on a peer rank")
            return None
        cls._cache[key] = inst
        cls._status[key] = "peer-memory exchange kernels over NVLink (torch symmetric memory)"
        return inst

    _status = {}

    @staticmethod
    def _single_node(group) -> bool:
        """All ranks of the group on this host?  LOCAL_WORLD_SIZE (torchrun) or one visible device per rank."""
        world = dist.get_world_size(group)
        lws = os.environ.get("LOCAL_WORLD_SIZE")
        if lws is not None:
            return int(lws) >= world
        return torch.cuda.device_count() >= world

    @classmethod
    def _fallback(cls, key, why: str) -> None:
        cls._cache[key] = False
        cls._status[key] = f"torch.distributed collectives (peer exchange unavailable: {why})"
        warnings.warn(f"moma_b200: peer-memory exchange unavailable ({why}); using torch.distributed "
                      "collectives for the sharded queue", RuntimeWarning, stacklevel=3)

    @classmethod
    def status(cls):
        """Human-readable transport per (group, device) seen so far."""
        return sorted(set(cls._status.values()))

    # ---------------------------------------------------------------- collectives
    def _order(self, channel):
        """A channel is one epoch sequence: its calls must be ordered.  If this call comes from another stream than the
        channel's previous call, order it behind that call (inside a capture the step's own fork/join does it)."""
        cur = torch.cuda.current_stream()
        last = self._last_use.get(channel)
        if last is not None and last[0] != cur.cuda_stream and not torch.cuda.is_current_stream_capturing():
            cur.wait_event(last[1])
        return cur

    def note_use(self, channel):
        if not torch.cuda.is_current_stream_capturing():
            cur = torch.cuda.current_stream()
            ev = torch.cuda.Event()
            ev.record(cur)
            self._last_use[channel] = (cur.cuda_stream, ev)

    def link(self, channel: int, cell_bytes: int):
        """The C-ABI arguments (peer_bases_dev, ctrl_off, data_off, region_bytes, rank, world, channel) of one channel for
        kernels that speak the exchange protocol themselves (ops.nce_merge_push / nce_combine_poll).  ``cell_bytes``: the
        bytes one call writes into a receive region (tagged cells: twice the payload)."""
        if cell_bytes > self.region_bytes:
            raise RuntimeError("PeerExchange: message larger than the receive region")
        self._order(channel)
        return (self.bases_dev, 0, self.ctrl_bytes, self.region_bytes, self.rank, self.world, channel)

    def _run(self, src, stride_bytes, bytes_per_rank, cast, channel, out):
        if 2 * bytes_per_rank * self.world > self.region_bytes:          # 8-byte (word, tag) cells: 2x the payload
            raise RuntimeError("PeerExchange: message larger than the receive region")
        cur = self._order(channel)
        check(_lib.load().moma_peer_exchange(src.data_ptr(), stride_bytes, bytes_per_rank, int(cast), self.bases_dev,
                                             0, self.ctrl_bytes, self.region_bytes, self.rank, self.world, channel,
                                             out.data_ptr(), cur.cuda_stream))
        self.note_use(channel)
        return out

    def allgather(self, x: torch.Tensor, channel: int, to_bf16: bool = False) -> torch.Tensor:
        """[rows, ...] on every rank -> [world * rows, ...] (optionally cast fp32 -> bf16 on the way out)."""
        x = x.contiguous()
        dtype = torch.bfloat16 if to_bf16 else x.dtype
        if to_bf16 and x.dtype != torch.float32:
            raise RuntimeError("PeerExchange.allgather: the fused cast expects fp32 input")
        out = torch.empty((self.world * x.shape[0],) + tuple(x.shape[1:]), dtype=dtype, device=x.device)
        nbytes = x.numel() * out.element_size()
        if nbytes % 16:
            raise RuntimeError("PeerExchange: rows must be a multiple of 16 bytes")
        return self._run(x, 0, nbytes, to_bf16, channel, out)

    def alltoall(self, x: torch.Tensor, channel: int) -> torch.Tensor:
        """x[p] goes to rank p; returns y with y[s] = what rank s sent here.  x: [world, ...]."""
        x = x.contiguous()
        if x.shape[0] != self.world:
            raise RuntimeError("PeerExchange.alltoall: leading dimension must be the world size")
        out = torch.empty_like(x)
        nbytes = x[0].numel() * x.element_size()
        if nbytes % 16:
            raise RuntimeError("PeerExchange: blocks must be a multiple of 16 bytes")
        return self._run(x, nbytes, nbytes, False, channel, out)
