"""Multi-rank check (torchrun, NCCL): the K-sharded queue against the replicated one.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 scripts/check_sharded.py

Every rank holds a replicated MoCo (the reference's layout) and a ShardedMoCo with the same initial
queue; each step both see the same local (q, k) and the same all-gathered keys.  Loss rows, dq and
the top-1 flags must agree (fp32: 1e-5; bf16: 1e-3) and the gathered shards must equal the
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


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ce = torch.nn.CrossEntropyLoss()
    worst = {}
    for precision, tol in (("fp32", 1e-5), ("bf16", 1e-3)):
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
