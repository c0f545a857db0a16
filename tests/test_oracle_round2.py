"""Pins for the round-2 oracle restatements (MoCoAtt modes, the non-'mlp' heads, ShuffleBN index logic)
against golden vectors from the unmodified reference (tests/golden/make_golden.py: kat_mocoatt, kat_heads,
kat_shufflebn), and the host-side ShuffleBN of moma_b200 on 2 gloo ranks.  Runs on CPU."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import moma_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODES = ["all", "qk", "dual", "dual2", "self_qk", "self"]


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def _att_from_golden(g, tag, H=4, dt=np.float64):
    def att(name, x):
        sd = lambda n: g[f"{tag}sd_{name}.{n}"].astype(dt)
        return O.attention_forward(x, sd("qkv.weight"), sd("qkv.bias"), sd("proj.weight"), sd("proj.bias"), H)
    return att


@pytest.mark.parametrize("mode", MODES)
def test_mocoatt_forward_oracle(golden, mode):
    """MoMA/mem_moco.py:111-161, every attention mode: logits, enqueued rows (bit-exact) and pointer."""
    g = golden("kat_mocoatt")
    tag = mode + "_"
    mem = g[tag + "mem0"].astype(np.float64)
    logits, labels, idx, q2, k2 = O.mocoatt_forward(mem, 20, g[tag + "q"].astype(np.float64),
                                                    g[tag + "k"].astype(np.float64), mode,
                                                    _att_from_golden(g, tag), 0.15)
    assert logits.shape == g[tag + "logits"].shape
    assert rel(logits, g[tag + "logits"]) < 2e-6
    assert idx == int(g[tag + "index"]) == (20 + 6) % 24
    assert (labels == 0).all() and labels.dtype == np.int64
    # rows 20..23, 0..1 now hold the (attended) keys; everything else is untouched
    ids = O.enqueue_ids(6, 20, 24)
    untouched = np.setdiff1d(np.arange(24), ids)
    assert np.array_equal(g[tag + "mem1"][untouched], g[tag + "mem0"][untouched])
    assert rel(mem[ids], g[tag + "mem1"][ids]) < 2e-6
    # which attention modules train in which mode (KAT6 generalised): recorded from the reference run
    trained = sorted({k[len(tag) + 8:].split(".")[0] for k in g.files
                      if k.startswith(tag + "hasgrad_") and bool(g[k])})
    want = {"all": ["atts"], "qk": ["atts"], "dual": ["atts_n", "atts_p"], "dual2": ["atts_n", "atts_p"],
            "self_qk": ["atts_k", "atts_q"], "self": ["atts_k", "atts_q", "atts_queue"]}[mode]
    assert trained == want, (mode, trained)


def test_heads_oracle(golden):
    """criterion_moco_att.py:269-305: 'mlp_byol', 'linear' and the bare Normalize head."""
    g = golden("kat_heads")
    sd = lambda h, n: g[f"{h}_sd0_{n}"].astype(np.float64)
    y = O.embed_mlp_byol(g["mlp_byol_x"].astype(np.float64), sd("mlp_byol", "1.weight"), sd("mlp_byol", "1.bias"),
                         sd("mlp_byol", "2.weight"), sd("mlp_byol", "2.bias"), sd("mlp_byol", "4.weight"),
                         sd("mlp_byol", "4.bias"))
    assert rel(y, g["mlp_byol_y"]) < 2e-6
    y = O.embed_linear(g["linear_x"].astype(np.float64), sd("linear", "1.weight"), sd("linear", "1.bias"))
    assert rel(y, g["linear_y"]) < 2e-6
    assert rel(O.normalize(g["none_x"].astype(np.float64)), g["none_y"]) < 2e-6


def test_shuffle_bn_indices_are_a_permutation_and_its_inverse():
    rng = np.random.default_rng(0)
    perm = rng.permutation(12)
    rows = np.arange(12)
    fed = np.concatenate([O.shuffle_bn_indices(perm, r, 6)[0] for r in range(2)])     # what the ranks compute on
    back = np.concatenate([O.shuffle_bn_indices(perm, r, 6)[1] for r in range(2)])
    assert sorted(fed) == list(rows)
    assert np.array_equal(fed[back], rows)               # node_k[reverse] restores the original order


# ----------------------------------------------------------------------- ShuffleBN, 2 gloo ranks, our trainer
class TinyTeacher(torch.nn.Module):
    """Same stand-in momentum encoder as tests/golden/make_golden.py (definition must match)."""

    def __init__(self):
        super().__init__()
        self.conv = torch.nn.Conv2d(3, 5, 3, padding=1)
        self.bn = torch.nn.BatchNorm2d(5)
        self.fc = torch.nn.Linear(5, 4)

    def forward(self, x, is_feat=False):
        f = torch.relu(self.bn(self.conv(x))).mean(dim=(2, 3))
        logit = self.fc(f)
        return ([f], logit) if is_feat else logit


def _worker(rank, world, port, ret, mode="slices"):
    os.environ["MOMA_B200_SHUFFLE_BN"] = mode
    sys.path.insert(0, ROOT)
    from argparse import Namespace
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", world_size=world, rank=rank)
    from moma_b200 import ContrastTrainer

    class Normalize(torch.nn.Module):            # the product's Normalize is CUDA-only (no CPU fallback by design)
        def __init__(self, p):
            super().__init__()

        def forward(self, t):
            return torch.nn.functional.normalize(t, p=2, dim=1)
    args = Namespace(local_rank=rank, node_rank=0, ngpus_per_node=world, rank=rank, mem="MoCo")
    tr = ContrastTrainer(args)
    tr.local_group = dist.new_group(ranks=list(range(world)), backend="gloo")
    torch.manual_seed(321)
    teacher = TinyTeacher().train()
    head = torch.nn.Sequential(torch.nn.Linear(5, 8), Normalize(2))
    torch.manual_seed(400 + rank)
    x = torch.randn(6, 3, 4, 4)
    torch.manual_seed(77)
    k, all_k = tr._shuffle_bn(x, teacher, head)
    ret[rank] = dict(k=k.numpy().copy(), all_k=all_k.numpy().copy(), bn_mean=teacher.bn.running_mean.numpy().copy(),
                     bn_var=teacher.bn.running_var.numpy().copy())
    dist.barrier(); dist.destroy_process_group()


@pytest.mark.parametrize("mode,port", [("slices", 29741), ("gather", 29743)])
def test_shuffle_bn_two_ranks_matches_reference(golden, mode, port):
    """learning/contrast_trainer.py:90-133: returned k / all_k and the BatchNorm statistics the shuffled
    batches leave behind equal the unmodified reference's 2-rank run (same seeds -> same permutation), both with the
    all-to-all of image slices (default: each rank receives only the B images it feeds to the momentum encoder) and
    with the reference's all-gather of the node's images."""
    g = golden("kat_shufflebn")
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret, mode), nprocs=2, join=True)
    for r in (0, 1):
        for key in ("k", "all_k", "bn_mean", "bn_var"):
            assert np.allclose(ret[r][key], g[f"r{r}_{key}"], rtol=1e-6, atol=1e-7), (r, key)
    assert np.allclose(ret[0]["all_k"], ret[1]["all_k"])


def test_criterion_step_oracle_matches_reference_run(golden):
    """O.criterion_step (loss and d loss/d feat_s through heads + attention + InfoNCE) against the unmodified
    reference's 3-step run (tests/golden/criterion_step.npz) -- the checker bench.py's parity self-check uses."""
    g = golden("criterion_step")
    names = ["1.weight", "1.bias", "3.weight", "3.bias"]
    sd = {k[4:]: g[k] for k in g.files if k.startswith("sd0_")}
    mem = g["mem0"]
    for st in range(3):
        ema = [sd["embed_t." + n].astype(np.float32).copy() for n in names]
        O.momentum_update(ema, [sd["embed_s." + n].astype(np.float32) for n in names], 0.999)    # loops_moma.py:310-312
        for n, e in zip(names, ema):
            sd["embed_t." + n] = e
        out = O.criterion_step(g[f"st{st}_feat_s"], g[f"st{st}_feat_t"], sd, mem, 0.15, 4)
        assert abs(out["loss"] - float(g[f"st{st}_loss"])) < 2e-6 * abs(float(g[f"st{st}_loss"]))
        assert rel(out["dfeat_s"], g[f"st{st}_dfeat_s"]) < 5e-6
        assert rel(out["f_s"], g[f"st{st}_f_s"]) < 2e-6 and rel(out["k"], g[f"st{st}_k2"]) < 2e-6
        sd = {k[len(f"st{st}_sd_"):]: g[k] for k in g.files if k.startswith(f"st{st}_sd_")}      # after the SGD step
        mem = g[f"st{st}_mem"]


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_cls_kd_oracle(golden, tag):
    """CrossEntropyLoss / DistillKL / their gradients restated (helper/loops_moma.py:278-279, distiller_zoo/KD.py:7-17)
    against the unmodified reference modules."""
    g = golden("kat_kd")
    ys, yt, lab, T = g[f"{tag}_ys"].astype(np.float64), g[f"{tag}_yt"].astype(np.float64), g[f"{tag}_lab"], float(g[f"{tag}_T"])
    assert abs(O.cross_entropy(ys, lab) - float(g[f"{tag}_cls"])) < 2e-6 * abs(float(g[f"{tag}_cls"]))
    assert abs(O.distill_kl(ys, yt, T) - float(g[f"{tag}_div"])) < 2e-6 * abs(float(g[f"{tag}_div"]))
    gc, gd = O.cls_kd_grads(ys, yt, lab, T)
    assert rel(gc, g[f"{tag}_gcls"]) < 2e-6 and rel(gd, g[f"{tag}_gdiv"]) < 2e-6
    assert O.accuracy_top1(ys, lab) == pytest.approx(float(g[f"{tag}_acc"][0]))


def test_sgd_ema_oracle_matches_torch_sgd_then_momentum_update():
    """The C restatement of SGD(momentum, weight_decay).step() + momentum_update against torch.optim.SGD on CPU and the
    reference's own momentum_update formula, 3 steps (the first one creates the momentum buffers)."""
    torch.manual_seed(5)
    shapes = [(7,), (33, 17), (4097,), (64, 3, 3, 3)]
    ps = [torch.nn.Parameter(torch.randn(*s)) for s in shapes]
    es = [torch.randn(*s) for s in shapes]
    opt = torch.optim.SGD(ps, lr=0.05, momentum=0.9, weight_decay=5e-4)
    P = [p.detach().numpy().copy() for p in ps]
    E = [e.numpy().copy() for e in es]
    Bf = [np.zeros_like(p) for p in P]
    for step in range(3):
        grads = [torch.randn(*s) for s in shapes]
        for p, g in zip(ps, grads):
            p.grad = g.clone()
        opt.step()
        for p1, p2 in zip(ps, es):                                   # learning/contrast_trainer.py:209-211
            p2.data.mul_(0.999).add_(p1.detach().data, alpha=(1 - 0.999))
        O.sgd_ema_step(P, [g.numpy().copy() for g in grads], Bf, E, 0.05, 0.9, 5e-4, step == 0, 0.999)
        for a, b in zip(P, ps):
            assert np.array_equal(a, b.detach().numpy()), step
        for a, b in zip(E, es):
            assert np.array_equal(a, b.numpy()), step
