"""World-size-2 (gloo, CPU) test of the K-sharded queue's host logic: query all-gather, routing of
the partials to their owner rank, merge, owned-row enqueue, pointer, gather-on-save.

There is no GPU here, so the kernel entry points in ``moma_b200.ops`` are replaced by the CPU
oracle (tests may do that; the product never does).  The expected values are the golden vectors of
the REFERENCE run under 2-rank gloo (tests/golden/kat_gloo.npz): replicated queue, same seeds.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _install_oracle_backend():
    from moma_b200 import ops
    from oracle import moma_oracle as O

    def nce_partial(q, queue, inv_T, dtype, n_splits=None):
        n_splits = n_splits or 2
        qn, qu = q.float().numpy().astype(np.float64), queue.float().numpy().astype(np.float64)
        K = qu.shape[0]
        bounds = [K * s // n_splits for s in range(n_splits + 1)]
        parts = [O.nce_partial(qn, qu[bounds[s]:bounds[s + 1]], 1.0 / inv_T) for s in range(n_splits)]
        stats = np.stack([np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]),
                          np.stack([p[0] for p in parts])])
        return torch.from_numpy(stats).float(), torch.from_numpy(np.stack([p[2] for p in parts])).float()

    def nce_merge_packed(stats, Op):
        m, l, mm, O_ = stats[0].double(), stats[1].double(), stats[2].double(), Op.double()
        mref = m.max(0).values
        w = torch.exp(m - mref)
        Om = (w.unsqueeze(-1) * O_).sum(0)
        B = Om.shape[0]
        return torch.cat([Om, mref[:, None], (w * l).sum(0)[:, None], mm.max(0).values[:, None],
                          torch.zeros(B, 1, dtype=torch.float64)], dim=1).float()

    def nce_combine_packed(packed, q32, k32, inv_T, round_bf16, dq_scale):
        D = packed.shape[2] - 4
        parts = [(packed[s, :, D].double().numpy(), packed[s, :, D + 1].double().numpy(),
                  packed[s, :, :D].double().numpy()) for s in range(packed.shape[0])]
        rows, dq, pim = O.nce_merge(parts, q32.double().numpy(), k32.double().numpy(), 1.0 / inv_T)
        mx = np.maximum(packed[:, :, D + 2].double().numpy().max(0),
                        (q32.double().numpy() * k32.double().numpy()).sum(1) * inv_T)
        return (torch.from_numpy(rows).float(), torch.from_numpy(dq * dq_scale).float(),
                torch.from_numpy(pim.astype(np.int32)), torch.from_numpy(mx).float(),
                torch.tensor(rows.mean(), dtype=torch.float32), torch.tensor([pim.mean() * 100.0], dtype=torch.float32))

    def enqueue(keys, queue, shadow, K, index, rank=0, world=1, normalize=False, eps=1e-12, index_dev=None):
        ids = O.enqueue_ids(keys.shape[0], index, K)
        owner, slot = O.shard_owner_slot(ids, world)
        for j in range(keys.shape[0]):
            if owner[j] == rank:
                queue[int(slot[j])] = keys[j].detach()

    ops.nce_partial, ops.enqueue = nce_partial, enqueue
    ops.nce_merge_packed, ops.nce_combine_packed = nce_merge_packed, nce_combine_packed
    ops.bf16_supported = lambda D: False
    ops.set_precision("fp32")


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", world_size=world, rank=rank)
    _install_oracle_backend()
    from argparse import Namespace

    from moma_b200 import ContrastTrainer, build_mem
    from moma_b200.sharded import ShardedMoCo
    torch.manual_seed(99)
    m = build_mem(Namespace(mem="MoCo", feat_dim=16, nce_k=32, nce_t=0.15, shard_queue=True))
    assert isinstance(m, ShardedMoCo) and tuple(m.memory_shard.shape) == (16, 16)
    ContrastTrainer(Namespace(mem="MoCo")).broadcast_memory(m)
    mem0 = m.memory.clone()                                   # collective gather
    torch.manual_seed(100 + rank)
    q = torch.randn(4, 16, requires_grad=True); k = torch.randn(4, 16)
    all_k = ContrastTrainer._global_gather(k)
    m.index = 28
    logits, labels = m(q, k, defer_enqueue=True)              # the queue update is scheduled after the backward
    losses, accs = ContrastTrainer._compute_loss_accuracy([logits], labels, torch.nn.CrossEntropyLoss())
    losses[0].backward()
    m.enqueue(all_k=all_k)
    # state_dict() is collective-free (a rank-0-only save must not deadlock): only rank 0 calls it here
    local_keys = []
    if rank == 0:
        sd_local = m.state_dict()
        local_keys = list(sd_local.keys())
        m3 = ShardedMoCo(16, 32, 0.15)
        m3.load_state_dict(sd_local)                          # per-rank shard + pointer round trip
        assert torch.equal(m3.memory_shard, m.memory_shard) and m3.index == m.index
    sd = m.full_state_dict()                                  # explicit collective: the reference's key set + pointer
    ret[rank] = dict(mem0=mem0.numpy(), mem1=sd["memory"].numpy(), index=m.index, loss=losses[0].item(),
                     acc=accs[0].item(), dq=q.grad.numpy(), all_k=all_k.numpy(), shape=tuple(logits.shape),
                     shard=m.memory_shard.numpy().copy(), keys=list(sd.keys()), local_keys=local_keys)
    # load_state_dict scatters the full queue back into shards and restores the pointer
    m2 = ShardedMoCo(16, 32, 0.15)
    m2.load_state_dict(sd)
    assert torch.equal(m2.memory_shard, m.memory_shard) and m2.index == m.index == 4
    dist.barrier(); dist.destroy_process_group()


def test_sharded_queue_matches_reference_two_rank_run(golden):
    g = golden("kat_gloo")
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29733, ret), nprocs=2, join=True)
    from oracle import moma_oracle as O
    for r in (0, 1):
        out = ret[r]
        assert np.array_equal(out["mem0"], g[f"r{r}_mem0"])                # same RNG draw, gathered view
        assert np.array_equal(out["all_k"], g[f"r{r}_all_k"])              # KAT7
        assert out["shape"] == tuple(g[f"r{r}_logits"].shape)
        want_loss, rows = O.cross_entropy_zero_label(g[f"r{r}_logits"].astype(np.float64))
        assert abs(out["loss"] - want_loss) < 1e-5 * abs(want_loss)
        assert abs(out["loss"] - float(g[f"r{r}_loss"])) < 1e-5 * abs(want_loss)
        _, _, dq_o, pim = O.nce_loss_and_grad(g[f"r{r}_q"].astype(np.float64), g[f"r{r}_k"].astype(np.float64),
                                              g[f"r{r}_mem0"].astype(np.float64), 0.15)
        assert np.linalg.norm(out["dq"] - dq_o) < 1e-5 * np.linalg.norm(dq_o)
        assert out["acc"] == pytest.approx(pim.mean() * 100)
        assert np.array_equal(out["mem1"], g[f"r{r}_mem1"])                # sharded enqueue == replicated, bit-exact
        assert out["index"] == int(g[f"r{r}_index"]) == 4
        assert np.array_equal(out["shard"], g[f"r{r}_mem1"][r::2])         # cyclic ownership
        assert "memory" in out["keys"] and "memory_shard" not in out["keys"]
    assert "memory_shard" in ret[0]["local_keys"] and "memory" not in ret[0]["local_keys"]
