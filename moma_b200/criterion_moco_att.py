"""MoMA criterion container -- host-side mirror of the reference
MoMA/criterion_moco_att.py (same class names, constructor arguments, sub-module
and parameter names, construction order -> identical RNG draws and state_dict).

``Normalize``, the ``Attention`` family and the projection heads' Linear(+ReLU) layers
run on the sm_100a kernels behind include/moma_b200.h; the modules keep their ``nn.Linear``
parameters (same state_dict), only the computation is routed.
"""
from __future__ import annotations

import os

import torch
import torch.nn.functional as F
from torch import nn

from . import ops

eps = 1e-7
_GATHER_PROJECTIONS_MAX_BYTES = int(os.environ.get("MOMA_B200_GATHER_PROJECTIONS_MAX_BYTES", 1 << 30))


class Normalize(nn.Module):
    """F.normalize(x, p, dim=1)  (reference :12-18)"""

    def __init__(self, p=2):
        super().__init__()
        self.p = p

    def forward(self, x):
        if self.p == 2 and x.dim() == 2 and x.shape[1] % 4 == 0 and x.dtype == torch.float32:
            return ops.l2_normalize(x)
        ops._need_cuda(x)
        return F.normalize(x, p=self.p, dim=1)          # other p / ranks: not on the MoMA path


class Flatten(nn.Module):
    """torch.flatten(x, 1)  (reference :21-27)"""

    def __init__(self):
        super().__init__()

    @staticmethod
    def forward(x):
        return torch.flatten(x, 1)


# ---- random-Fourier-feature heads (reference :31-112; unused by the shipped drivers) ----
def input_mapping_torch(x, B_w, B_b):
    return torch.cos(torch.matmul(x, B_w) + B_b)


def _rff_basis(in_dim, out_dim, w_scale, b_scale, device):
    B_w = torch.empty((in_dim, out_dim)).normal_(mean=0, std=1).to(device)
    B_b = torch.distributions.uniform.Uniform(0, 6.283).sample([1, out_dim]).to(device)
    return B_w * w_scale, B_b * b_scale


class RFF_ST(nn.Module):
    def __init__(self, w_scale=1., b_scale=1., b_init='gauss01', RFF_init='gauss01', out_dim=128):
        super().__init__()
        self.w_scale, self.b_scale, self.b_init, self.out_dim, self.RFF_init = w_scale, b_scale, b_init, out_dim, RFF_init

    def forward(self, x, xt):
        x, xt = x.flatten(start_dim=1), xt.flatten(start_dim=1)
        B_w, B_b = _rff_basis(x.shape[-1], self.out_dim, self.w_scale, self.b_scale, x.device)
        return input_mapping_torch(x, B_w, B_b), input_mapping_torch(xt, B_w, B_b)


class RFF(nn.Module):
    def __init__(self, w_scale=1., b_scale=1., b_init='gauss01', RFF_init='gauss01', out_dim=128):
        super().__init__()
        self.w_scale, self.b_scale, self.b_init, self.out_dim, self.RFF_init = w_scale, b_scale, b_init, out_dim, RFF_init

    def forward(self, x):
        x = x.flatten(start_dim=1)
        B_w, B_b = _rff_basis(x.shape[-1], self.out_dim, self.w_scale, self.b_scale, x.device)
        out = input_mapping_torch(x, B_w, B_b)
        return (2 / self.in_dim) ** 0.5 * out       # AttributeError as in the reference (:84)


class RFF_fixed(nn.Module):
    def __init__(self, in_dim, w_scale=1., b_scale=1., b_init='gauss01', RFF_init='gauss01', out_dim=128):
        super().__init__()
        self.w_scale, self.b_scale, self.b_init, self.out_dim, self.RFF_init = w_scale, b_scale, b_init, out_dim, RFF_init
        self.in_dim = in_dim
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        self.B_w, self.B_b = _rff_basis(in_dim, out_dim, w_scale, b_scale, dev)

    def forward(self, x):
        x = x.flatten(start_dim=1)
        return (2 / self.in_dim) ** 0.5 * input_mapping_torch(x, self.B_w, self.B_b)


# ---- attention over the batch axis -------------------------------------------------------
class Attention(nn.Module):
    """reference :141-167.  x [N, dim] -> [N, dim]; the batch is the token axis."""

    def __init__(self, dim, num_heads=12, qkv_bias=False, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = head_dim ** -0.5

        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def _fusable(self, x):
        drop = self.training and (self.attn_drop.p > 0 or self.proj_drop.p > 0)
        return (not drop) and x.dim() == 2 and ops.attention_supported(x.shape[1], self.num_heads)

    def _composed(self, x, want_probs=False):
        """The reference formulation op by op (dropout > 0 or an unsupported head_dim)."""
        ops._need_cuda(x)
        x = x.unsqueeze(0)
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        attn = (q @ k.transpose(-2, -1)) * self.scale
        attn = self.attn_drop(attn.softmax(dim=-1))
        x = (attn @ v).transpose(1, 2).reshape(N, C)
        x = self.proj_drop(self.proj(x))
        return (x, attn) if want_probs else x

    def forward(self, x):
        if not self._fusable(x):
            return self._composed(x)
        return ops.attention(x, self.qkv.weight, self.qkv.bias, self.proj.weight, self.proj.bias,
                             self.num_heads)


    def forward_rows(self, x, q_start, q_stride, q_count):
        """Rows q_start + i * q_stride (i < q_count) of ``self(x)`` without computing the others; no autograd.
        (K-sharded queue: a rank only enqueues every W-th attended key.)"""
        if not self._fusable(x):
            with torch.no_grad():
                return self._composed(x)[q_start::q_stride][:q_count]
        return ops.attention_rows(x, self.qkv.weight, self.qkv.bias, self.proj.weight, self.proj.bias,
                                  self.num_heads, q_start, q_stride, q_count)


    def forward_rows_gathered(self, x_local, gather, q_start, q_stride, q_count):
        """``forward_rows`` over the concatenation of every rank's ``x_local`` (``gather``: [B, .] -> [W * B, .], e.g.
        ContrastTrainer._global_gather).  No autograd.  Two schedules:
          * default: every rank projects only its own rows (qkv Linear) and the PROJECTIONS are all-gathered
            (3C floats per token) -- no rank projects a token twice;
          * above MOMA_B200_GATHER_PROJECTIONS_MAX_BYTES: the raw tokens are all-gathered (C floats per token, the
            reference's key gather, learning/contrast_trainer.py:124) and projected locally -- 3x fewer bytes on the wire
            for one extra GEMM over all W x B tokens.  Measured at 8 GPUs (C3): slower end to end (0.338 vs 0.291 ms per
            step: the extra 4096-token GEMM competes with the critical path), hence not the default."""
        with torch.no_grad():
            if not self._fusable(x_local):
                return self._composed(gather(x_local))[q_start::q_stride][:q_count]
            n_tokens = q_stride * q_count if q_stride > 1 else x_local.shape[0]
            if n_tokens * 3 * x_local.shape[1] * 4 > _GATHER_PROJECTIONS_MAX_BYTES:
                return ops.attention_rows(gather(x_local), self.qkv.weight, self.qkv.bias, self.proj.weight, self.proj.bias,
                                          self.num_heads, q_start, q_stride, q_count)
            qkv_local = ops.linear(x_local, self.qkv.weight, self.qkv.bias, key="rows_gathered")
            return ops.attention_rows_from_qkv(gather(qkv_local), self.proj.weight, self.proj.bias, self.num_heads,
                                               q_start, q_stride, q_count)


class Attention_viz(Attention):
    """reference :171-197: also returns the attention map [1, H, N, N]."""

    def forward(self, x):
        if not self._fusable(x):
            return self._composed(x, want_probs=True)
        return ops.attention(x, self.qkv.weight, self.qkv.bias, self.proj.weight, self.proj.bias,
                             self.num_heads, want_probs=True)


class Attention_(Attention):
    """reference :201-225: takes [N, C] without the unsqueeze; its reshape/permute is
    shape-inconsistent in the reference (4-D reshape, 5-index permute) and raises there too."""

    def forward(self, x):
        N, C = x.shape
        return self.qkv(x).reshape(N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)


class Attention2(nn.Module):
    """reference :227-233: LayerNorm(Attention(x) + x)"""

    def __init__(self, dim, num_heads=12, qkv_bias=False, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.attn_layer = Attention(dim, num_heads, qkv_bias, attn_drop, proj_drop)
        self.norm = nn.LayerNorm(dim)

    def forward(self, x):
        return self.norm(self.attn_layer(x) + x)


class _Head(nn.Sequential):
    """Projection head (same sub-module indices / state_dict keys as the reference's nn.Sequential).

    On the GPU every ``nn.Linear`` (+ the ``nn.ReLU`` that follows it) runs as one ``moma_linear_fwd`` launch --
    3xTF32 tensor-core GEMM with the bias / ReLU epilogue fused, fp32-level accuracy -- and back-propagates
    through ``moma_linear_bwd``.  One exception: in bf16 mode a LARGE layer evaluated WITHOUT autograd (the
    momentum teacher's ``embed_t`` inside ``_shuffle_bn``, learning/contrast_trainer.py:117-121) uses a
    single-pass TF32 library GEMM: no gradient flows through it and its output is rounded to bf16 by the InfoNCE
    kernel anyway (measured end-to-end effect 5e-5 on the gradients, scripts/tf32_heads_error.py)."""

    _TF32_MIN_WEIGHT = 1 << 20          # elements; below this the 3xTF32 kernel is as fast as the library call

    def _tf32_library(self, lin):
        return (not torch.is_grad_enabled() and ops.get_precision() == "bf16"
                and lin.weight.numel() >= self._TF32_MIN_WEIGHT)

    def forward(self, x):
        if not x.is_cuda or os.environ.get("MOMA_B200_GEMM") == "simt":
            return super().forward(x)
        mods = list(self)
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, nn.Linear) and x.dim() == 2:
                if self._tf32_library(m):
                    relu = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU) and m.bias is not None
                    prev = torch.backends.cuda.matmul.allow_tf32
                    torch.backends.cuda.matmul.allow_tf32 = True
                    try:
                        if relu:        # bias + ReLU in the GEMM epilogue (one launch instead of two)
                            x = torch._addmm_activation(m.bias, x, m.weight.t(), use_gelu=False)
                            i += 1
                        else:
                            x = m(x)
                    finally:
                        torch.backends.cuda.matmul.allow_tf32 = prev
                else:
                    relu = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
                    x = ops.linear(x, m.weight, m.bias, relu=relu)
                    i += 1 if relu else 0
            else:
                x = m(x)
            i += 1
        return x


def _head(kind, in_dim, feat_dim):
    """Projection heads of CMO (reference :254-305)."""
    if kind == 'mlp':
        return _Head(Flatten(), nn.Linear(in_dim, in_dim), nn.ReLU(inplace=True),
                             nn.Linear(in_dim, feat_dim), Normalize(2))
    if kind == 'mlp_byol':
        return _Head(Flatten(), nn.Linear(in_dim, in_dim), nn.BatchNorm1d(in_dim),
                             nn.ReLU(inplace=True), nn.Linear(in_dim, feat_dim), Normalize(2))
    if kind == 'linear':
        return _Head(Flatten(), nn.Linear(in_dim, feat_dim), Normalize(2))
    return _Head(Flatten(), Normalize(2))


class CMO(nn.Module):
    """Container of the projection heads and the attention modules of the MoMA loss
    (reference :236-338).  No forward: helper/loops_moma.py:308-335 drives the parts.

    opt.head in {'mlp', 'mlp_byol', 'linear', other}; opt.attn selects the attention set.
    ``opt.num_heads`` (optional, default 4 = the reference's hard-coded value) sets the heads.
    """

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.embed_s = _head(opt.head, opt.s_dim, opt.feat_dim)
        self.embed_t = _head(opt.head, opt.t_dim, opt.feat_dim)

        self.norm1 = nn.LayerNorm
        self.qkv_bias = True
        H = int(getattr(opt, "num_heads", 4))
        D = opt.feat_dim

        def att(cls=Attention):
            return cls(D, num_heads=H, qkv_bias=self.qkv_bias, attn_drop=0., proj_drop=0.)

        if opt.attn in ['all', 'self_mix', 'qk']:
            self.atts = att()
        elif opt.attn in ['dual', 'dual2']:
            self.atts_p = att()
            self.atts_n = att()
        elif opt.attn in ['self_qk', 'self_nomix']:
            self.atts_q = att()
            self.atts_k = att()
        elif opt.attn in ['self_qkv2']:
            self.atts_q = att(Attention2)
            self.atts_k = att(Attention2)
        elif opt.attn in ['selfv2']:
            self.atts_q = att(Attention2)
            self.atts_k = att(Attention2)
            self.atts_queue = att(Attention2)
        elif opt.attn == 'self_viz':
            self.atts_q = att(Attention_viz)
            self.atts_k = att(Attention_viz)
            self.atts_queue = att(Attention_viz)
        else:  # 'self'
            self.atts_q = att()
            self.atts_k = att()
            self.atts_queue = att()


class CMO_EmaTec(nn.Module):
    """EMA-teacher variant (reference :344-419): embed_s / embed_ema / embed_t."""

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        if opt.head == 'mlp':
            def mlp(d_in, d_mid):
                return nn.Sequential(Flatten(), nn.Linear(d_in, d_in), nn.ReLU(inplace=True),
                                     nn.Linear(d_mid, opt.feat_dim), Normalize(2))
            self.embed_s = mlp(opt.s_dim, opt.t_dim)       # the reference feeds t_dim here (:367)
            self.embed_ema = mlp(opt.s_dim, opt.t_dim)
            self.embed_t = mlp(opt.t_dim, opt.t_dim)
        elif opt.head in ('RFF_fixed', 'RFF'):
            self.embed_s = RFF_fixed(in_dim=opt.s_dim, out_dim=opt.feat_dim)
            self.embed_ema = RFF_fixed(in_dim=opt.s_dim, out_dim=opt.feat_dim)
            self.embed_t = RFF_fixed(in_dim=opt.s_dim, out_dim=opt.feat_dim)
        else:
            self.embed_s = nn.Sequential(Flatten(), Normalize(2))
            self.embed_ema = nn.Sequential(Flatten(), Normalize(2))
            self.embed_t = nn.Sequential(Flatten(), Normalize(2))

    def forward(self, f_s, f_ema, f_t):
        return self.embed_s(f_s), self.embed_ema(f_ema), self.embed_t(f_t)
