"""Host-side logic that needs no GPU: RNG / state_dict parity of the module mirrors with the
reference run, the lazy-logits dispatch, pointer arithmetic, and loud failure without CUDA."""
from argparse import Namespace

import numpy as np
import pytest
import torch
import torch.nn as nn

import moma_b200
from moma_b200 import CMO, ContrastTrainer, LazyLogits, MoCo, MoCoST, accuracy, build_mem
from moma_b200.ops import NceOut


def test_same_seed_same_queue_and_parameters(golden):
    """Same constructor order and RNG draws as the reference: identical initial queue and weights."""
    g = golden("criterion_step")
    torch.manual_seed(12345)
    opt = Namespace(head="mlp", s_dim=24, t_dim=24, feat_dim=32, attn="self", mem="MoCo", nce_k=40, nce_t=0.15)
    contrast = build_mem(opt)
    crit = CMO(opt)
    assert isinstance(contrast, MoCo) and contrast.index == 0 and contrast.K == 40 and contrast.T == 0.15
    assert np.array_equal(contrast.memory.numpy(), g["mem0"])
    assert list(contrast.state_dict().keys()) == ["memory"]
    want = sorted(k[4:] for k in g.files if k.startswith("sd0_"))
    assert sorted(crit.state_dict().keys()) == want
    for name, p in crit.state_dict().items():
        assert np.array_equal(p.numpy(), g["sd0_" + name]), name
    # KAT1 queue
    g1 = golden("kat_moco")
    torch.manual_seed(0)
    m = MoCo(128, 4096, 0.15)
    assert np.array_equal(m.memory.numpy()[:40], g1["kat1_mem_rows"])


def test_cmo_attention_sets():
    base = dict(head="linear", s_dim=16, t_dim=24, feat_dim=32)
    names = lambda attn: sorted({k.split(".")[0] for k in CMO(Namespace(attn=attn, **base)).state_dict()})
    assert names("self") == ["atts_k", "atts_q", "atts_queue", "embed_s", "embed_t"]
    assert names("all") == ["atts", "embed_s", "embed_t"] == names("qk") == names("self_mix")
    assert names("dual") == ["atts_n", "atts_p", "embed_s", "embed_t"]
    assert names("self_qk") == ["atts_k", "atts_q", "embed_s", "embed_t"]
    c = CMO(Namespace(attn="selfv2", **base))
    assert "atts_queue.norm.weight" in c.state_dict() and "atts_q.attn_layer.qkv.weight" in c.state_dict()
    assert CMO(Namespace(attn="self", num_heads=8, **base)).atts_q.num_heads == 8
    assert CMO(Namespace(attn="self", **base)).atts_q.num_heads == 4          # the reference's hard-coded value
    assert build_mem(Namespace(mem="MoCoST", feat_dim=8, nce_k=16, nce_t=0.1)).__class__ is MoCoST


def test_pointer_sequence(golden):
    g = golden("kat_pointer")
    for K, n in ((4096, 96), (10, 4), (7, 7), (12, 5)):
        m = MoCo(4, K, 0.07)
        for want in g[f"ptr_K{K}_n{n}_index"]:
            m._update_pointer(n)
            assert m.index == want


def _fake_handle(B=4, K=9):
    q = torch.randn(B, requires_grad=True)
    rows = q * 2
    nce = NceOut(rows.mean(), rows, torch.tensor([1, 0, 1, 1], dtype=torch.int32), torch.arange(B, dtype=torch.float32),
                 torch.tensor([75.0]))
    labels = torch.zeros(B, dtype=torch.long)
    dense = torch.randn(B, K + 1)
    return LazyLogits((B, K + 1), torch.device("cpu"), nce, labels, lambda: dense), labels, dense, q


def test_lazy_logits_dispatch():
    h, labels, dense, q = _fake_handle()
    assert isinstance(h, torch.Tensor) and tuple(h.shape) == (4, 10) and h.dtype == torch.float32 and h.dim() == 2
    loss = nn.CrossEntropyLoss()(h, labels)                      # served from the fused results
    loss.backward()
    assert torch.allclose(q.grad, torch.full((4,), 0.5))
    assert torch.allclose(nn.CrossEntropyLoss(reduction="sum")(h, labels), (q * 2).sum())
    _, pred = h.topk(1, 1, True, True)                           # learning/util.py:31
    assert pred.t().eq(labels.view(1, -1)).tolist() == [[True, False, True, True]]
    losses, accs = ContrastTrainer._compute_loss_accuracy([h], labels, nn.CrossEntropyLoss())
    assert accs[0].item() == 75.0 and accuracy(h, labels)[0].item() == 75.0
    # anything else (or foreign labels) falls back to the materialised logits
    other = torch.zeros(4, dtype=torch.long)
    assert torch.allclose(nn.CrossEntropyLoss()(h, other), nn.CrossEntropyLoss()(dense, other))
    assert torch.allclose(h + 1, dense + 1) and torch.allclose(h[:, 0], dense[:, 0])
    assert torch.allclose(accuracy(dense, labels)[0], (dense.argmax(1) == 0).float().mean(0, keepdim=True) * 100)


def test_no_cpu_fallback():
    m = MoCo(16, 32, 0.1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(4, 16), torch.randn(4, 16))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        moma_b200.Attention(16, num_heads=2)(torch.randn(4, 16))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        moma_b200.Normalize(2)(torch.randn(4, 16))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ContrastTrainer.momentum_update(nn.Linear(4, 4), nn.Linear(4, 4), 0.9)
    # the reference's error on heterogeneous student / teacher comes before any launch
    with pytest.raises(RuntimeError, match="must match"):
        ContrastTrainer.momentum_update(nn.Linear(4, 4), nn.Linear(4, 5), 0.9)
    with pytest.raises(ValueError):
        moma_b200.set_precision("fp8")


def test_import_compat_packages():
    """The reference's import lines (train_student_moma.py:37-39) resolve to this implementation."""
    from learning.contrast_trainer import ContrastTrainer as CT
    from MoMA.criterion_moco_att import CMO as C2, Attention, Normalize, Flatten
    from MoMA.mem_moco import build_mem as bm, MoCo as M2, MoCoAtt, MoCoSSTT
    assert CT is ContrastTrainer and C2 is CMO and bm is build_mem and M2 is MoCo
    x = torch.randn(3, 2, 5)
    assert Flatten()(x).shape == (3, 10)
    for name in ("init_ddp_environment", "broadcast_memory", "_shuffle_bn", "_global_gather",
                 "_compute_loss_accuracy", "momentum_update", "adjust_learning_rate", "warmup_learning_rate"):
        assert hasattr(CT, name), name
