#!/usr/bin/env python
"""Drive the UNMODIFIED reference loop ``helper/loops_moma.py:train_distill_moma`` (vendored copy under baseline/_ref,
scripts/vendor_ref.sh) on one GPU, with the set-up of ``train_student_moma.py:323-410`` for ``--distill moma``:

    --arm ours        MoMA.* and learning.* resolve to THIS repository (sys.path order), everything else --
                      helper/, models/, distiller_zoo/ -- to the reference: the drop-in as INTEGRATION.md describes it
    --arm reference   every module is the reference's own (its true GPU behaviour, stock PyTorch kernels)

Same seeds in both arms -> same backbones, same CMO parameters, same queue, same batches.  Prints one JSON line:
per-iteration total loss (cls + div + beta * kd), the final queue pointer, a checksum of the queue, and -- with
``--time`` -- the wall-clock step time of the whole loop (L2 of SURVEY 8d: backbones included, the reference loop's
two ``.item()`` syncs per step included).

Used by tests/test_reference_loop_gpu.py (loss sequences of the two arms must agree) and bench.py (``l2_end_to_end``).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arm", choices=["ours", "reference"], required=True)
    ap.add_argument("--ref", default=os.environ.get("MOMA_REFERENCE", os.path.join(ROOT, "baseline", "_ref")))
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--K", type=int, default=256)
    ap.add_argument("--D", type=int, default=128)
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--lr", type=float, default=2e-4,
                    help="SGD learning rate.  Iteration 1 starts from identical state in both arms; later iterations start "
                         "from parameters updated with gradients that agree to ~1e-5, and a batch-16 train-mode-BatchNorm "
                         "ResNet amplifies such differences by ~10x per step in proportion to the step size (measured: "
                         "lr 0.05 -> loss 7, 18, 58 and 1e-3 apart at step 3; lr 0.002 -> 4e-5 apart).  Parity of the LOOP is "
                         "checked on a gently updating trajectory; kernels are checked on identical inputs elsewhere.")
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--port", type=int, default=29777)
    a = ap.parse_args()
    if not os.path.isdir(os.path.join(a.ref, "helper")):
        print(json.dumps({"unavailable": f"{a.ref} missing (run scripts/vendor_ref.sh in the build container)"}))
        return 0

    sys.modules.setdefault("tensorboard_logger", types.ModuleType("tensorboard_logger"))   # learning/base_trainer.py:9
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != ROOT]
    if a.arm == "ours":
        sys.path.insert(0, a.ref)
        sys.path.insert(0, ROOT)                       # MoMA/, learning/ of this repo shadow the reference's
    else:
        sys.path.insert(0, a.ref)

    import torch
    import torch.distributed as dist
    import torch.nn as nn
    import torch.optim as optim

    from distiller_zoo import DistillKL                               # reference
    from helper.loops_moma import train_distill_moma                   # reference, unchanged
    from learning.contrast_trainer import ContrastTrainer              # ours | reference
    from MoMA.criterion_moco_att import CMO                            # ours | reference
    from MoMA.mem_moco import build_mem                                # ours | reference
    from models import model_dict                                      # reference
    import MoMA.mem_moco as mm
    origin = os.path.abspath(mm.__file__)
    if a.arm == "ours":
        assert origin.startswith(ROOT) and "baseline" not in origin, origin
        import moma_b200
        moma_b200.set_precision(a.precision)
    else:
        assert os.path.abspath(a.ref) in origin, origin

    torch.cuda.set_device(0)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{a.port}", world_size=1, rank=0)

    opt = argparse.Namespace(
        distill="moma", head="mlp", attn="self", mem="MoCo", feat_dim=a.D, nce_k=a.K, nce_t=0.15, alpha=0.999,
        gpu=0, rank=0, multiprocessing_distributed=True, batch_size=a.batch, print_freq=10 ** 9,
        cls=1.0, div=1.0, beta=1.0, kd_T=4, learning_rate=a.lr, momentum=0.9, weight_decay=5e-4,
        local_rank=0, node_rank=0, ngpus_per_node=1, world_size=1)
    trainer = ContrastTrainer(opt)
    trainer.local_group = dist.new_group(ranks=[0])

    torch.manual_seed(12345)                                           # train_student_moma.py:55 default --seed
    model_t = model_dict["ResNet18"](num_classes=4)
    model_s = model_dict["ResNet18"](num_classes=4)
    data = torch.randn(2, 3, a.size, a.size)
    model_t.eval(); model_s.eval()
    with torch.no_grad():
        feat_t, _ = model_t(data, is_feat=True)
        feat_s, _ = model_s(data, is_feat=True)
    opt.s_dim, opt.t_dim = feat_s[-1].shape[1], feat_t[-1].shape[1]     # :331-332
    module_list = nn.ModuleList([model_s])
    trainable_list = nn.ModuleList([model_s])
    contrast = build_mem(opt)                                          # :336
    contrast.cuda()
    trainer.broadcast_memory(contrast)                                 # :339
    criterion_kd = CMO(opt)                                            # :341
    module_list.append(criterion_kd.embed_s); module_list.append(criterion_kd.embed_t)
    trainable_list.append(criterion_kd.embed_s)
    criterion_kd.embed_t.eval()
    trainable_list.append(criterion_kd.atts_q); trainable_list.append(criterion_kd.atts_k)
    trainable_list.append(criterion_kd.atts_queue)
    criterion_list = nn.ModuleList([nn.CrossEntropyLoss(), DistillKL(opt.kd_T), criterion_kd])
    module_list.append(model_t)
    optimizer = optim.SGD(trainable_list.parameters(), lr=opt.learning_rate, momentum=opt.momentum,
                          weight_decay=opt.weight_decay)
    module_list.cuda(0)
    DDP = torch.nn.parallel.DistributedDataParallel
    module_list = [DDP(model_s, device_ids=[0]), model_t.cuda()]        # :407-411
    criterion_list.cuda(0)

    g = torch.Generator().manual_seed(777)
    batches = [(torch.randn(a.batch, 3, a.size, a.size, generator=g), torch.randint(0, 4, (a.batch,), generator=g))
               for _ in range(a.iters)]
    losses, states = [], []

    def csum(ts):
        return float(sum(float(t.detach().double().abs().sum().item()) for t in ts))

    for it, batch in enumerate(batches):                               # one-batch "epochs": the loop returns that step's loss
        _, loss_avg = train_distill_moma(it + 1, [batch], module_list, criterion_list, trainer, contrast, optimizer, opt)
        losses.append(float(loss_avg))
        states.append({                                                # |.|-sums of every piece of state after the step
            "student": csum(model_s.parameters()), "teacher": csum(model_t.parameters()),
            "teacher_bn": csum(b for n, b in model_t.named_buffers() if "running" in n),
            "embed_s": csum(criterion_kd.embed_s.parameters()), "embed_t": csum(criterion_kd.embed_t.parameters()),
            "atts_q": csum(criterion_kd.atts_q.parameters()), "atts_k": csum(criterion_kd.atts_k.parameters()),
            "atts_queue": csum(criterion_kd.atts_queue.parameters()), "queue": csum([contrast.memory]),
            "grad_student": csum(p.grad for p in model_s.parameters() if p.grad is not None),
            "grad_embed_s": csum(p.grad for p in criterion_kd.embed_s.parameters() if p.grad is not None),
            "grad_atts_q": csum(p.grad for p in criterion_kd.atts_q.parameters() if p.grad is not None),
            "grad_atts_k": csum(p.grad for p in criterion_kd.atts_k.parameters() if p.grad is not None),
        })
    torch.cuda.synchronize()
    mem = contrast.memory if hasattr(contrast, "memory") else None
    trained = [p for p in trainable_list.parameters()]
    out = {"arm": a.arm, "param_abs_sum": csum(trained),
           "grad_abs_sum": csum(p.grad for p in trained if p.grad is not None), "origin": origin, "precision": a.precision if a.arm == "ours" else "fp32 (stock)",
           "losses": losses, "states": states, "index": int(contrast.index), "lr": a.lr,
           "queue_sum": float(mem.double().sum().item()), "queue_abs_sum": float(mem.double().abs().sum().item()),
           "config": {"batch": a.batch, "size": a.size, "K": a.K, "D": a.D, "s_dim": opt.s_dim, "t_dim": opt.t_dim}}
    if a.time:
        n_t = max(a.iters, 10)
        loader = [batches[i % len(batches)] for i in range(n_t)]
        train_distill_moma(99, loader[:3], module_list, criterion_list, trainer, contrast, optimizer, opt)   # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        train_distill_moma(100, loader, module_list, criterion_list, trainer, contrast, optimizer, opt)
        torch.cuda.synchronize()
        out["ms_per_step"] = (time.perf_counter() - t0) / n_t * 1e3
        out["timed_steps"] = n_t
    dist.destroy_process_group()
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
