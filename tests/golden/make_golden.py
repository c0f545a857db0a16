#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):   python tests/golden/make_golden.py

Shims (none of them touch arithmetic):
  * a stub ``tensorboard_logger`` module (learning/base_trainer.py:9 imports it),
  * ``Tensor.cuda`` / ``Module.cuda`` -> identity on this CUDA-less host (the
    reference calls ``.cuda()`` unconditionally, MoMA/mem_moco.py:25,94).

Everything is seeded; fp32; torch CPU, 1 thread for reproducibility.
"""
import os
import sys
import types
from argparse import Namespace

import numpy as np
import torch
import torch.nn as nn

REF = os.environ.get("MOMA_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))

sys.modules.setdefault("tensorboard_logger", types.ModuleType("tensorboard_logger"))
if not torch.cuda.is_available():
    torch.Tensor.cuda = lambda self, *a, **k: self
    nn.Module.cuda = lambda self, *a, **k: self
sys.path.insert(0, REF)

from MoMA.mem_moco import MoCo, MoCoST, MoCoSSTT, build_mem          # noqa: E402
from MoMA.criterion_moco_att import CMO, Attention, Normalize        # noqa: E402
from learning.contrast_trainer import ContrastTrainer                # noqa: E402
import torch.nn.functional as F                                      # noqa: E402

torch.set_num_threads(1)


def npy(t):
    return t.detach().cpu().numpy().copy()


def save(name, **arrs):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"{name}.npz: {os.path.getsize(path)/1024:.1f} KiB  keys={len(arrs)}")


# ---------------------------------------------------------------- KAT1: MoCo
def kat_moco():
    out = {}
    # KAT1 of SURVEY 8c: scalars only (the full logits would be 0.5 MB)
    torch.manual_seed(0)
    m = MoCo(128, 4096, 0.15)
    q = F.normalize(torch.randn(32, 128)).requires_grad_()
    k = F.normalize(torch.randn(32, 128))
    mem0 = m.memory.clone()
    logits, labels = m(q, k)
    loss = nn.CrossEntropyLoss()(logits, labels)
    loss.backward()
    out.update(kat1_q=npy(q), kat1_k=npy(k), kat1_mem_rows=npy(mem0[:40]),
               kat1_mem_sum=np.float64(mem0.double().sum().item()),
               kat1_loss=np.float32(loss.item()), kat1_gradnorm=np.float32(q.grad.norm().item()),
               kat1_logits_head=npy(logits[:4, :8]), kat1_index=np.int64(m.index),
               kat1_mem_after_rows=npy(m.memory[:40]), kat1_labels=npy(labels),
               kat1_dq=npy(q.grad))
    # the queue itself is needed to recompute: store it as fp16-free exact fp32 but
    # only for a smaller case below.  Small full case:
    torch.manual_seed(1)
    m = MoCo(32, 64, 0.07)
    mem0 = m.memory.clone()
    q = torch.randn(8, 32).requires_grad_()          # NOT normalised (post-attention regime)
    k = torch.randn(8, 32)
    all_k = torch.randn(24, 32)
    m.index = 56                                      # wraps: rows 56..63, 0..15
    logits, labels = m(q, k, all_k)
    loss = nn.CrossEntropyLoss()(logits, labels)
    loss.backward()
    acc = ContrastTrainer._compute_loss_accuracy([logits], labels, nn.CrossEntropyLoss())[1][0]
    out.update(s_q=npy(q), s_k=npy(k), s_allk=npy(all_k), s_mem0=npy(mem0), s_logits=npy(logits),
               s_labels=npy(labels), s_loss=np.float32(loss.item()), s_dq=npy(q.grad),
               s_mem1=npy(m.memory), s_index=np.int64(m.index), s_acc=npy(acc))
    # KAT4: B == 1 squeeze
    torch.manual_seed(2)
    m = MoCo(16, 32, 0.2)
    mem0 = m.memory.clone()
    q = torch.randn(1, 16); k = torch.randn(1, 16)
    logits, labels = m(q, k)
    out.update(b1_q=npy(q), b1_k=npy(k), b1_mem0=npy(mem0), b1_logits=npy(logits),
               b1_labels=npy(labels))
    save("kat_moco", **out)


# --------------------------------------------- KAT2/KAT3: pointer + ids
def kat_pointer():
    out = {}
    m = MoCo(4, 10, 0.07)
    m.index = 8
    m.memory.zero_()
    k = torch.arange(16, dtype=torch.float32).view(4, 4) + 1
    m._update_memory(k, m.memory); m._update_pointer(4)
    out.update(kat2_mem=npy(m.memory), kat2_index=np.int64(m.index))
    for K, n, steps in ((4096, 96, 45), (10, 4, 9), (65536, 512, 130), (131072, 1024, 130), (7, 7, 3), (12, 5, 11)):
        m = MoCo(2, K, 0.07)
        idxs, first_ids, last_ids = [], [], []
        for s in range(steps):
            # replicate the id computation verbatim (mem_moco.py:25-26)
            ids = torch.fmod(torch.arange(n) + m.index, m.K).long()
            first_ids.append(int(ids[0])); last_ids.append(int(ids[-1]))
            m._update_pointer(n)
            idxs.append(m.index)
        out[f"ptr_K{K}_n{n}_index"] = np.array(idxs, dtype=np.int64)
        out[f"ptr_K{K}_n{n}_first"] = np.array(first_ids, dtype=np.int64)
        out[f"ptr_K{K}_n{n}_last"] = np.array(last_ids, dtype=np.int64)
    # a full ids vector across a wrap
    m = MoCo(2, 4096, 0.07); m.index = 4032
    out["ids_wrap"] = npy(torch.fmod(torch.arange(96) + m.index, m.K).long())
    save("kat_pointer", **out)


# ----------------------------------------------------- Attention fwd/bwd
def kat_attention():
    out = {}
    for tag, (N, C, H, bias) in {"a": (24, 32, 4, True), "b": (40, 64, 8, True),
                                 "c": (17, 32, 4, False)}.items():
        torch.manual_seed(10 + N)
        att = Attention(C, num_heads=H, qkv_bias=bias)
        x = torch.randn(N, C).requires_grad_()
        y = att(x)
        dy = torch.randn(N, C)
        y.backward(dy)
        out.update({f"{tag}_x": npy(x), f"{tag}_y": npy(y), f"{tag}_dy": npy(dy),
                    f"{tag}_H": np.int64(H), f"{tag}_scale": np.float64(att.scale),
                    f"{tag}_wqkv": npy(att.qkv.weight), f"{tag}_wproj": npy(att.proj.weight),
                    f"{tag}_bproj": npy(att.proj.bias),
                    f"{tag}_dx": npy(x.grad), f"{tag}_dwqkv": npy(att.qkv.weight.grad),
                    f"{tag}_dwproj": npy(att.proj.weight.grad), f"{tag}_dbproj": npy(att.proj.bias.grad)})
        if bias:
            out[f"{tag}_bqkv"] = npy(att.qkv.bias); out[f"{tag}_dbqkv"] = npy(att.qkv.bias.grad)
    save("kat_attention", **out)


# ----------------------------------------------------------- Normalize
def kat_normalize():
    torch.manual_seed(3)
    x = torch.randn(12, 48)
    x[3] = 0.0
    x[5] *= 1e-14
    x[7] *= 1e3
    x = x.requires_grad_()
    y = Normalize(2)(x)
    g = torch.randn(12, 48)
    y.backward(g)
    save("kat_normalize", x=npy(x), y=npy(y), g=npy(g), dx=npy(x.grad))


# ----------------------------------------------------------------- EMA
def kat_ema():
    torch.manual_seed(4)
    shapes = [(7,), (64, 3, 7, 7), (64,), (129,), (33, 17), (1,), (96, 40)]
    ms = nn.ParameterList([nn.Parameter(torch.randn(*s)) for s in shapes])
    me = nn.ParameterList([nn.Parameter(torch.randn(*s) * 3) for s in shapes])
    out = {f"src{i}": npy(p) for i, p in enumerate(ms)}
    out.update({f"ema{i}_0": npy(p) for i, p in enumerate(me)})
    for step in (1, 2, 3):
        ContrastTrainer.momentum_update(ms, me, 0.999)
        out.update({f"ema{i}_{step}": npy(p) for i, p in enumerate(me)})
    # another momentum value
    me2 = nn.ParameterList([nn.Parameter(torch.randn(*s)) for s in shapes])
    out.update({f"emb{i}_0": npy(p) for i, p in enumerate(me2)})
    ContrastTrainer.momentum_update(ms, me2, 0.5)
    out.update({f"emb{i}_1": npy(p) for i, p in enumerate(me2)})
    out["n"] = np.int64(len(shapes))
    # shape mismatch must raise
    try:
        ContrastTrainer.momentum_update(nn.ParameterList([nn.Parameter(torch.randn(3, 4))]),
                                        nn.ParameterList([nn.Parameter(torch.randn(4, 3))]), 0.9)
        raised = False
    except RuntimeError:
        raised = True
    out["mismatch_raises"] = np.bool_(raised)
    save("kat_ema", **out)


# -------------------------------------- full criterion step (moma branch)
def criterion_step():
    """helper/loops_moma.py:308-335 driven with synthetic features (the
    backbones are outside the path); 3 steps so the enqueue wraps."""
    torch.manual_seed(12345)
    opt = Namespace(head="mlp", s_dim=24, t_dim=24, feat_dim=32, attn="self", mem="MoCo",
                    nce_k=40, nce_t=0.15, alpha=0.999)
    contrast = build_mem(opt)
    crit = CMO(opt)
    B = 16
    out = {"mem0": npy(contrast.memory)}
    for name, p in crit.state_dict().items():
        out["sd0_" + name] = npy(p)
    sgd = torch.optim.SGD([p for n, p in crit.named_parameters() if not n.startswith("embed_t")], lr=0.05)
    for step in range(3):
        feat_s = torch.randn(B, opt.s_dim).requires_grad_()
        feat_t = torch.randn(B, opt.t_dim)
        # :310-312
        crit.embed_t.eval()
        ContrastTrainer.momentum_update(crit.embed_s, crit.embed_t, opt.alpha)
        with torch.no_grad():
            k = crit.embed_t(feat_t)                    # _shuffle_bn's head call (:121), W=1
        all_k = k                                       # _global_gather at W=1 (:124)
        f_s = crit.embed_s(feat_s)                      # :323-324
        f_s = crit.atts_q(f_s); k2 = crit.atts_k(k); all_k2 = crit.atts_queue(all_k)   # :326-329
        output = contrast(q=f_s, k=k2, all_k=all_k2)    # :331
        losses, accs = ContrastTrainer._compute_loss_accuracy(output[:-1], output[-1],
                                                              nn.CrossEntropyLoss())
        loss = losses[0]
        sgd.zero_grad(); loss.backward()
        out.update({f"st{step}_feat_s": npy(feat_s), f"st{step}_feat_t": npy(feat_t),
                    f"st{step}_k": npy(k), f"st{step}_f_s": npy(f_s), f"st{step}_k2": npy(k2),
                    f"st{step}_allk2": npy(all_k2), f"st{step}_loss": np.float32(loss.item()),
                    f"st{step}_acc": npy(accs[0]), f"st{step}_dfeat_s": npy(feat_s.grad),
                    f"st{step}_mem": npy(contrast.memory), f"st{step}_index": np.int64(contrast.index)})
        for n, p in crit.named_parameters():
            out[f"st{step}_grad_{n}"] = npy(p.grad) if p.grad is not None else np.zeros(0, np.float32)
            out[f"st{step}_hasgrad_{n}"] = np.bool_(p.grad is not None)
        sgd.step()
        for n, p in crit.state_dict().items():
            out[f"st{step}_sd_{n}"] = npy(p)
    save("criterion_step", **out)


# ------------------------------------------------- MoCoST / MoCoSSTT
def kat_dual():
    torch.manual_seed(7)
    out = {}
    m = MoCoST(16, 24, 0.1)
    out.update(st_ms0=npy(m.memory_s), st_mt0=npy(m.memory_t))
    q, k, kt = torch.randn(4, 16), torch.randn(4, 16), torch.randn(4, 16)
    m.index = 22
    lss, lst, lab = m(q, k, kt)
    out.update(st_q=npy(q), st_k=npy(k), st_kt=npy(kt), st_lss=npy(lss), st_lst=npy(lst),
               st_ms1=npy(m.memory_s), st_mt1=npy(m.memory_t), st_index=np.int64(m.index))
    m = MoCoSSTT(16, 24, 0.1)
    out.update(sstt_ms0=npy(m.memory_s), sstt_mt0=npy(m.memory_t))
    qt = torch.randn(4, 16)
    res = m(q, k, qt, kt)
    out.update(sstt_qt=npy(qt), sstt_lss=npy(res[0]), sstt_lst=npy(res[1]), sstt_lts=npy(res[2]),
               sstt_ltt=npy(res[3]), sstt_index=np.int64(m.index))
    save("kat_dual", **out)


# ------------------------------------- 2-rank gloo: _global_gather + MoCo
def _gloo_worker(rank, world, port, ret):
    import torch.distributed as dist
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", world_size=world, rank=rank)
    torch.manual_seed(99)                       # same seed -> same queue on every rank
    m = MoCo(16, 32, 0.15)
    mem0 = m.memory.clone()
    torch.manual_seed(100 + rank)
    q = torch.randn(4, 16); k = torch.randn(4, 16)
    all_k = ContrastTrainer._global_gather(k)
    m.index = 28
    logits, labels = m(q, k, all_k)
    loss = nn.CrossEntropyLoss()(logits, labels)
    ret[rank] = dict(q=npy(q), k=npy(k), all_k=npy(all_k), logits=npy(logits), mem0=npy(mem0),
                     mem1=npy(m.memory), index=m.index, loss=loss.item())
    dist.barrier(); dist.destroy_process_group()


def kat_gloo():
    import torch.multiprocessing as mp
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_gloo_worker, args=(2, 29611, ret), nprocs=2, join=True)
    out = {}
    for r in (0, 1):
        for key, v in ret[r].items():
            out[f"r{r}_{key}"] = np.asarray(v)
    save("kat_gloo", **out)


# ------------------------------------------------- MoCoAtt: every attention mode
MOCOATT_MODES = [("all", "all"), ("qk", "qk"), ("dual", "dual"), ("dual2", "dual2"), ("self_qk", "self_qk"),
                 ("self", "self")]           # (CMO opt.attn, MoCoAtt.forward attn=)


def kat_mocoatt():
    """MoMA/mem_moco.py:111-161 with the CMO attention set each mode needs (criterion_moco_att.py:308-338).
    Small K because 'all' / 'dual' / 'self' attend the whole queue (N = K tokens)."""
    from MoMA.mem_moco import MoCoAtt
    out = {}
    D, K, B, T = 32, 24, 6, 0.15
    for opt_attn, mode in MOCOATT_MODES:
        torch.manual_seed(500 + len(mode))
        opt = Namespace(head="linear", s_dim=8, t_dim=8, feat_dim=D, attn=opt_attn)
        crit = CMO(opt)
        m = MoCoAtt(D, K, T)
        m.index = 20                                   # wraps: rows 20..23, 0..1
        q = torch.randn(B, D).requires_grad_()
        k = torch.randn(B, D)
        mem0 = m.memory.clone()
        logits, labels = m(q, k, attn=mode, criterion_kd=crit)
        if mode == "dual2":
            loss = logits.sum()                        # [B] positives only (:51-66): no CE over one column
        else:
            loss = nn.CrossEntropyLoss()(logits, labels)
        loss.backward()
        tag = f"{mode}_"
        out.update({tag + "q": npy(q), tag + "k": npy(k), tag + "mem0": npy(mem0), tag + "logits": npy(logits),
                    tag + "loss": np.float32(loss.item()), tag + "dq": npy(q.grad), tag + "mem1": npy(m.memory),
                    tag + "index": np.int64(m.index)})
        for n, p_ in crit.state_dict().items():
            out[tag + "sd_" + n] = npy(p_)
        for n, p_ in crit.named_parameters():
            out[tag + "hasgrad_" + n] = np.bool_(p_.grad is not None)
            if p_.grad is not None:
                out[tag + "grad_" + n] = npy(p_.grad)
    save("kat_mocoatt", **out)


# ------------------------------------------------- projection heads other than 'mlp'
def kat_heads():
    """criterion_moco_att.py:269-305: 'mlp_byol' (train-mode BatchNorm1d), 'linear', and the bare Normalize head."""
    out = {}
    for head in ("mlp_byol", "linear", "none"):
        torch.manual_seed(700 + len(head))
        opt = Namespace(head=head, s_dim=20, t_dim=12, feat_dim=16, attn="self")
        crit = CMO(opt)
        x = torch.randn(10, 20 if head != "none" else 16).requires_grad_()
        for n, p_ in crit.embed_s.state_dict().items():
            out[f"{head}_sd0_{n}"] = npy(p_)
        y = crit.embed_s(x)
        g = torch.randn_like(y)
        y.backward(g)
        out.update({f"{head}_x": npy(x), f"{head}_y": npy(y), f"{head}_g": npy(g), f"{head}_dx": npy(x.grad)})
        for n, p_ in crit.embed_s.named_parameters():
            out[f"{head}_grad_{n}"] = npy(p_.grad)
        for n, p_ in crit.embed_s.state_dict().items():
            out[f"{head}_sd1_{n}"] = npy(p_)            # BatchNorm running stats after the train-mode forward
    save("kat_heads", **out)


# ------------------------------------- 2-rank gloo: ShuffleBN (contrast_trainer.py:90-133)
class TinyTeacher(nn.Module):
    """Stand-in momentum encoder with the (feats, logit) = model(x, is_feat=True) convention of the reference
    backbones; BatchNorm in train mode makes the output depend on WHICH samples share a device -- exactly what
    ShuffleBN permutes."""

    def __init__(self):
        super().__init__()
        self.conv = nn.Conv2d(3, 5, 3, padding=1)
        self.bn = nn.BatchNorm2d(5)
        self.fc = nn.Linear(5, 4)

    def forward(self, x, is_feat=False):
        f = torch.relu(self.bn(self.conv(x))).mean(dim=(2, 3))
        logit = self.fc(f)
        return ([f], logit) if is_feat else logit


def _shufflebn_worker(rank, world, port, ret):
    import torch.distributed as dist
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", world_size=world, rank=rank)
    args = Namespace(local_rank=rank, node_rank=0, ngpus_per_node=world, rank=rank, mem="MoCo")
    tr = ContrastTrainer(args)
    tr.local_group = dist.new_group(ranks=list(range(world)), backend="gloo")
    torch.manual_seed(321)                        # same weights on every rank
    teacher = TinyTeacher().train()
    head = nn.Sequential(nn.Linear(5, 8), Normalize(2))
    torch.manual_seed(400 + rank)
    x = torch.randn(6, 3, 4, 4)
    torch.manual_seed(77)                         # the permutation is rank 0's draw (broadcast), seeded here
    k, all_k = tr._shuffle_bn(x, teacher, head)
    ret[rank] = dict(x=npy(x), k=npy(k), all_k=npy(all_k),
                     bn_mean=npy(teacher.bn.running_mean), bn_var=npy(teacher.bn.running_var))
    dist.barrier(); dist.destroy_process_group()


def kat_shufflebn():
    import torch.multiprocessing as mp
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_shufflebn_worker, args=(2, 29613, ret), nprocs=2, join=True)
    out = {}
    for r in (0, 1):
        for key, v in ret[r].items():
            out[f"r{r}_{key}"] = np.asarray(v)
    save("kat_shufflebn", **out)


# ------------------------------------- classification CE + DistillKL + accuracy
def kat_kd():
    """helper/loops_moma.py:278-279,350: CrossEntropyLoss, distiller_zoo.DistillKL(4), helper.util.accuracy on [B, n_cls]."""
    from distiller_zoo import DistillKL
    from helper.util import accuracy as acc_fn
    out = {}
    for tag, (B, C, T) in {"a": (16, 4, 4.0), "b": (37, 8, 2.0), "c": (5, 100, 1.0)}.items():
        torch.manual_seed(900 + B)
        ys = (torch.randn(B, C) * 3).requires_grad_()
        yt = torch.randn(B, C) * 3
        lab = torch.randint(0, C, (B,))
        l_cls = nn.CrossEntropyLoss()(ys, lab)
        l_div = DistillKL(T)(ys, yt)
        acc = acc_fn(ys, lab, topk=(1,))[0]
        g_cls, = torch.autograd.grad(l_cls, ys, retain_graph=True)
        g_div, = torch.autograd.grad(l_div, ys)
        out.update({f"{tag}_ys": npy(ys), f"{tag}_yt": npy(yt), f"{tag}_lab": npy(lab), f"{tag}_T": np.float32(T),
                    f"{tag}_cls": np.float32(l_cls.item()), f"{tag}_div": np.float32(l_div.item()), f"{tag}_acc": npy(acc),
                    f"{tag}_gcls": npy(g_cls), f"{tag}_gdiv": npy(g_div)})
    save("kat_kd", **out)


if __name__ == "__main__":
    only = sys.argv[1:]
    for fn in (kat_moco, kat_pointer, kat_attention, kat_normalize, kat_ema, criterion_step, kat_dual, kat_gloo,
               kat_mocoatt, kat_heads, kat_shufflebn, kat_kd):
        if not only or fn.__name__ in only:
            fn()
