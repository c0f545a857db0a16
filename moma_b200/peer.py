"""Peer-memory exchange (NVLink / NVSwitch) for the K-sharded queue -- host side of csrc/peer.cu.

One symmetric allocation per process group (torch.distributed._symmetric_memory: every rank maps every
peer's copy), a control block plus two receive regions per channel.  ``allgather`` / ``alltoall`` are
ONE kernel launch each (push of (word, epoch-tag) pairs + poll + copy-out) and are CUDA-graph capturable: the epoch that
sequences the calls lives in device memory.  Where symmetric memory is unavailable (CPU / gloo tests,
no peer access) ``PeerExchange.create`` returns None and the callers use torch.distributed collectives.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check

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
        self.region_bytes = (max(int(region_bytes), 4 << 20) + 255) // 256 * 256     # roomy: re-creation is collective
        total = self.ctrl_bytes + 2 * _N_CHANNELS * self.region_bytes
        self.buf = symm_mem.empty(total, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, self.group)
        torch.cuda.synchronize(device)
        self.handle.barrier()                       # every rank's control block is zero before anyone pushes
        self.bases_dev = int(self.handle.buffer_ptrs_dev)

    # ---------------------------------------------------------------- construction
    @classmethod
    def create(cls, group, device, region_bytes: int) -> Optional["PeerExchange"]:
        """The shared instance for (group, device), grown on demand; None when peer memory is unavailable."""
        if device.type != "cuda" or not dist.is_initialized():
            return None
        group = group if group is not None else dist.group.WORLD
        if dist.get_backend(group) != "nccl" or dist.get_world_size(group) > 16:
            return None
        key = (id(group), device.index)
        inst = cls._cache.get(key)
        if inst is False:
            return None
        if inst is not None and inst.region_bytes >= region_bytes:
            return inst
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("PeerExchange: the symmetric buffer must be created before CUDA-graph capture "
                               "(run one eager step first)")
        try:
            inst = cls(group, device, region_bytes)
        except Exception:                                # no symmetric-memory support on this system
            cls._cache[key] = False
            return None
        cls._cache[key] = inst
        return inst

    # ---------------------------------------------------------------- collectives
    def _run(self, src, stride_bytes, bytes_per_rank, cast, channel, out):
        if 2 * bytes_per_rank * self.world > self.region_bytes:          # 8-byte (word, tag) cells: 2x the payload
            raise RuntimeError("PeerExchange: message larger than the receive region")
        check(_lib.load().moma_peer_exchange(src.data_ptr(), stride_bytes, bytes_per_rank, int(cast), self.bases_dev,
                                             0, self.ctrl_bytes, self.region_bytes, self.rank, self.world, channel,
                                             out.data_ptr(), torch.cuda.current_stream().cuda_stream))
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
