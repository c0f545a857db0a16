"""torch port of the reference's MoMA criterion step -- TEST / BASELINE INFRASTRUCTURE ONLY.

This is the CPU arm that bench.py times (``cpu_baseline.kind = "port"`` and
``--impl reference``): the reference's own op sequence on its own arithmetic type
(fp32, ATen CPU kernels, all host threads), restated from

    helper/loops_moma.py:308-335          (the moma branch of train_distill_moma)
    MoMA/mem_moco.py:14-49,77-100         (MoCo.forward, _compute_logit, _update_memory/_pointer)
    MoMA/criterion_moco_att.py:12-27,141-167,251-338   (Normalize, Flatten, Attention, CMO)
    learning/contrast_trainer.py:83-88,189-211, learning/util.py:25-41

The unmodified reference cannot travel to the GPU box (/root/reference does not exist there),
so this port stands in for it; it is pinned to the reference by tests/test_torch_port.py
(golden vectors generated from the real modules).  Nothing in moma_b200 imports this file.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class PortAttention(nn.Module):
    """criterion_moco_att.py:141-167"""

    def __init__(self, dim, num_heads=4, qkv_bias=True):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        x = x.unsqueeze(0)
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        attn = ((q @ k.transpose(-2, -1)) * self.scale).softmax(dim=-1)
        x = (attn @ v).transpose(1, 2).reshape(N, C)
        return self.proj(x)


def port_head(in_dim, feat_dim):
    """criterion_moco_att.py:254-268 ('mlp' head)"""
    class _Norm(nn.Module):
        def forward(self, x):
            return F.normalize(x, p=2, dim=1)

    class _Flat(nn.Module):
        def forward(self, x):
            return torch.flatten(x, 1)

    return nn.Sequential(_Flat(), nn.Linear(in_dim, in_dim), nn.ReLU(inplace=True), nn.Linear(in_dim, feat_dim),
                         _Norm())


class PortMoCo(nn.Module):
    """mem_moco.py:69-100.  ``operand_dtype`` (GPU baseline arm only): dtype of the q / queue operands of the
    negatives GEMM -- torch.bfloat16 is "the reference with bf16 operands" of SURVEY 8d(iii): the library
    (cuBLAS sm_100) tensor-core GEMM the fused kernel has to beat; None = the reference's fp32."""

    def __init__(self, n_dim, K, T, operand_dtype=None):
        super().__init__()
        self.K, self.T, self.index = K, T, 0
        self.operand_dtype = operand_dtype
        self.device_constants = False     # True (graph-captured GPU arm): labels / ids are created on the device --
                                          # a capture cannot contain the reference's per-step pageable H2D copies
        self.register_buffer("memory", F.normalize(torch.randn(K, n_dim)))

    def forward(self, q, k, all_k=None):
        bsz = q.size(0)
        dev = q.device
        k = k.detach()
        queue = self.memory.clone().detach()                              # :89
        pos = torch.bmm(q.view(bsz, 1, -1), k.view(bsz, -1, 1)).view(bsz, 1)    # :38-39
        if self.operand_dtype is None:
            neg = torch.mm(queue, q.transpose(1, 0)).transpose(0, 1)      # :42-43
        else:
            od = self.operand_dtype
            neg = torch.mm(queue.to(od), q.to(od).transpose(1, 0)).transpose(0, 1).float()
        out = torch.div(torch.cat((pos, neg), dim=1), self.T).squeeze().contiguous()   # :45-47
        if self.device_constants:
            labels = torch.zeros(bsz, dtype=torch.long, device=dev)
        else:
            labels = torch.zeros(bsz, dtype=torch.long).to(dev)           # :94 (CPU tensor, then .cuda())
        all_k = all_k if all_k is not None else k
        with torch.no_grad():                                             # :23-27
            if self.device_constants:
                ids = torch.fmod(torch.arange(all_k.shape[0], device=dev) + self.index, self.K).long()
            else:
                ids = torch.fmod(torch.arange(all_k.shape[0]) + self.index, self.K).long().to(dev)
            self.memory.index_copy_(0, ids, all_k)
        self.index = (self.index + all_k.size(0)) % self.K                # :14-15
        return out, labels


def port_accuracy(output, target):
    """learning/util.py:25-41, topk=(1,)"""
    with torch.no_grad():
        _, pred = output.topk(1, 1, True, True)
        correct = pred.t().eq(target.view(1, -1))
        return correct[:1].reshape(-1).float().sum(0, keepdim=True).mul_(100.0 / target.size(0))


def port_momentum_update(params, params_ema, m):
    """contrast_trainer.py:207-211"""
    for p1, p2 in zip(params, params_ema):
        p2.data.mul_(m).add_(p1.detach().data, alpha=(1 - m))


class PortCriterionStep:
    """One criterion step on CPU with the reference's op sequence (L1 of SURVEY 8d)."""

    def __init__(self, s_dim, t_dim, feat_dim, K, T, alpha, num_heads, ema_shapes, seed=12345, device="cpu",
                 operand_dtype=None):
        """device='cpu': the CPU arm (the default, what --impl reference times).  device='cuda': the same op
        sequence executed by the stock PyTorch/ATen/cuBLAS GPU kernels -- the reference's own GPU behaviour,
        timed by bench.py beside the repo's kernels (``gpu_reference``); never part of the product path."""
        torch.manual_seed(seed)
        self.contrast = PortMoCo(feat_dim, K, T, operand_dtype).to(device)
        self.embed_s = port_head(s_dim, feat_dim).to(device)
        self.embed_t = port_head(t_dim, feat_dim).to(device)
        self.atts_q = PortAttention(feat_dim, num_heads).to(device)
        self.atts_k = PortAttention(feat_dim, num_heads).to(device)
        self.atts_queue = PortAttention(feat_dim, num_heads).to(device)
        self.alpha = alpha
        self.student = [torch.randn(*s).to(device) for s in ema_shapes]
        self.teacher = [torch.randn(*s).to(device) for s in ema_shapes]
        self.head_ema = s_dim == t_dim
        self.ce = nn.CrossEntropyLoss()

    def step(self, feat_s, feat_t):
        port_momentum_update(self.student, self.teacher, self.alpha)             # loops_moma.py:309
        if self.head_ema:                                                        # :310-312
            port_momentum_update(list(self.embed_s.parameters()), list(self.embed_t.parameters()), self.alpha)
        with torch.no_grad():
            k = self.embed_t(feat_t)                                             # contrast_trainer.py:121
        all_k = k                                                                # _global_gather at W=1
        f_s = self.embed_s(feat_s)                                               # :323-324
        f_s = self.atts_q(f_s); k = self.atts_k(k); all_k = self.atts_queue(all_k)   # :326-329
        logits, labels = self.contrast(f_s, k, all_k)                            # :331
        loss = self.ce(logits, labels)                                           # contrast_trainer.py:197
        acc = port_accuracy(logits, labels)                                      # :200-204
        for p in self.params():
            p.grad = None
        loss.backward()                                                          # loops_moma.py:360
        return loss, acc

    def params(self):
        mods = (self.embed_s, self.atts_q, self.atts_k, self.atts_queue)
        return [p for m in mods for p in m.parameters()]
