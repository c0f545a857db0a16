"""Multi-rank check (torchrun, NCCL): the K-sharded queue against the replicated one.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 scripts/check_sharded.py

Every rank holds a replicated MoCo (the reference's layout) and a ShardedMoCo with the same initial
queue; each step both see the same local (q, k) and the same all-gathered keys.  Loss rows, dq and
the top-1 flags must agree (fp32: 1e-5; bf16: 2e-3 between the two layouts, each 1e-3 from the oracle) and the gathered shards must equal the
replicated queue bit-exactly after every step.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import moma_b200
from moma_b200 import ContrastTrainer, MoCo
from moma_b200.sharded import ShardedMoCo


def check_peer(rank, world):
    """moma_peer_exchange against the NCCL collectives, bit for bit, eagerly and replayed from a CUDA graph."""
    from moma_b200.peer import CH_KEYS, CH_PARTIALS, CH_QUERIES, PeerExchange
    dev = torch.device("cuda", torch.cuda.current_device())
    peer = PeerExchange.create(None, dev, 1 << 20)
    assert peer is not None, "symmetric memory unavailable on this box"
    torch.manual_seed(1234 + rank)
    for it in range(6):                                  # epochs advance; both parities get reused
        rows = 64 * (1 + it % 3)
        x = torch.randn(rows, 128, device=dev)
        want = torch.empty(world * rows, 128, device=dev); dist.all_gather_into_tensor(want, x)
        assert torch.equal(peer.allgather(x, CH_KEYS), want)
        got16 = peer.allgather(x, CH_QUERIES, to_bf16=True)
        assert got16.dtype == torch.bfloat16 and torch.equal(got16, want.to(torch.bfloat16))
        y = torch.randn(world, rows, 132, device=dev)
        want2 = torch.empty_like(y); dist.all_to_all_single(want2, y)
        assert torch.equal(peer.alltoall(y, CH_PARTIALS), want2)
    # graph replay: the epoch lives in device memory, so one captured exchange replays correctly many times
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        x = torch.zeros(256, 128, device=dev)
        g = torch.cuda.CUDAGraph()
        peer.allgather(x, CH_KEYS)                        # warm-up outside the capture
        side.synchronize()
        with torch.cuda.graph(g, stream=side):
            out = peer.allgather(x, CH_KEYS)
        for it in range(5):
            x.fill_(float(rank * 10 + it))
            g.replay()
            side.synchronize()
            for r in range(world):
                assert float(out[r * 256, 0]) == float(r * 10 + it) and float(out[r * 256 + 255, 127]) == float(r * 10 + it)
    dist.barrier()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    check_peer(rank, world)
    ce = torch.nn.CrossEntropyLoss()
    worst = {}
    # bf16: both layouts are separately within 1e-3 of the oracle (tests/test_parity_gpu.py; measured 1e-4 on normalised
    # inputs, scripts/nce_bf16_error.py); two such results may differ by the sum of their errors, and the un-normalised
    # queries used here (|q| ~ 7, logits up to +-140 at T = 0.07) are the worst case for the bf16 rounding of P
    for precision, tol in (("fp32", 1e-5), ("bf16", 2e-3)):
        moma_b200.set_precision(precision)
        for (B, D, K, T) in ((64, 128, 4096, 0.15), (128, 256, 8192, 0.07), (32, 64, 1024, 0.15)):
            torch.manual_seed(7)
            rep = MoCo(D, K, T).cuda()
            torch.manual_seed(7)
            sh = ShardedMoCo(D, K, T).cuda()
            assert torch.equal(sh.memory, rep.memory)
            rep.index = sh.index = K - B * world // 2          # the enqueue wraps
            torch.manual_seed(100 + rank)
            for step in range(3):
                q1 = (torch.randn(B, D, device="cuda") * 0.6).requires_grad_()
                q2 = q1.detach().clone().requires_grad_()
                k = torch.randn(B, D, device="cuda") * 0.6
                all_k = ContrastTrainer._global_gather(k)
                lr, labr = rep(q1, k, all_k)
                if step == 1:       # only the rows this rank owns (what Attention.forward_rows would deliver)
                    start, stride, count = sh.owned_rows(all_k.shape[0])
                    ls, labs = sh(q2, k, owned_k=all_k[start::stride][:count].contiguous())
                else:
                    ls, labs = sh(q2, k, all_k)
                loss_r, loss_s = ce(lr, labr), ce(ls, labs)
                loss_r.backward(); loss_s.backward()
                e_loss = abs(loss_r.item() - loss_s.item()) / abs(loss_r.item())
                e_dq = ((q1.grad - q2.grad).norm() / q1.grad.norm()).item()
                assert e_loss < tol and e_dq < tol, (precision, B, D, K, step, e_loss, e_dq)
                assert torch.equal(lr.pos_is_max, ls.pos_is_max)
                assert torch.equal(sh.memory, rep.memory), "sharded queue diverged from the replicated one"
                assert sh.index == rep.index
                worst[precision] = max(worst.get(precision, 0.0), e_loss, e_dq)
    dist.barrier()
    if rank == 0:
        print(f"check_sharded OK world={world} worst rel err {worst}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
