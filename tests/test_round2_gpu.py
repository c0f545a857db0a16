"""Round-2 GPU parity tests: what round 1 only mirrored or left unprotected.

  * dense-logits training path (D > 512, MoCoAtt): backward AFTER the enqueue differentiates the pre-enqueue queue
  * MoCoAtt.forward, every attention mode, against the unmodified reference's run (tests/golden/kat_mocoatt.npz)
  * the non-'mlp' projection heads through CMO/_Head (kat_heads.npz)
  * the single-pass TF32 teacher layer of bf16 mode as a tolerance test
  * transient queues never hit a stale bf16 shadow; checkpointing of the ring pointer
  * the captured OVERLAPPED step (what bench.py times) == the sequential step, bit-exact queue; K-shards emulated in
    one process through nce_merge_packed / nce_combine_packed inside a captured graph
"""
from argparse import Namespace

import numpy as np
import pytest
import torch

from oracle import moma_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def cu(x):
    return torch.as_tensor(np.ascontiguousarray(x)).cuda()


def npy(t):
    return t.detach().float().cpu().numpy()


@pytest.fixture()
def fp32():
    import moma_b200
    moma_b200.set_precision("fp32")
    yield moma_b200
    moma_b200.set_precision("bf16")


@pytest.fixture()
def bf16():
    import moma_b200
    moma_b200.set_precision("bf16")
    yield moma_b200


# ------------------------------------------------------------------ dense path: backward after the enqueue
@pytest.mark.parametrize("D", [1280, 640])
def test_dense_path_backward_after_enqueue(fp32, D):
    """head='None' makes feat_dim = s_dim (1280 for EfficientNet-B0 / MobileNetV2, train_student_moma.py:331-332):
    the fused kernel does not cover D > 512, MoCo.forward takes the dense path, whose autograd node keeps the LIVE
    queue; the enqueue then overwrites n rows of it.  d loss/d q must still be the gradient of the logits that were
    returned (the reference clones the queue, mem_moco.py:89)."""
    from moma_b200 import MoCo
    torch.manual_seed(0)
    K, B, T = 96, 24, 0.15
    m = MoCo(D, K, T).cuda()
    m.index = 80                                        # wraps: rows 80..95, 0..7
    mem0 = npy(m.memory).astype(np.float64)
    q = (torch.randn(B, D, device="cuda") * 0.1).requires_grad_()
    k = torch.randn(B, D, device="cuda") * 0.1
    all_k = torch.randn(B, D, device="cuda")            # very different rows: a stale read would be far off
    logits, labels = m(q, k, all_k)
    assert not type(logits).__name__ == "LazyLogits"
    loss = torch.nn.CrossEntropyLoss()(logits, labels)
    assert m.index == (80 + B) % K                      # the queue HAS been updated before backward runs
    loss.backward()
    loss_o, _, dq_o, _ = O.nce_loss_and_grad(npy(q).astype(np.float64), npy(k).astype(np.float64), mem0, T)
    assert abs(loss.item() - loss_o) < TOL * abs(loss_o)
    assert rel(npy(q.grad), dq_o) < TOL
    want = mem0.astype(np.float32)
    O.update_memory(want, npy(all_k), 80)
    assert np.array_equal(npy(m.memory), want)


def test_dense_path_two_steps_do_not_leak_guards(fp32):
    from moma_b200 import MoCo
    torch.manual_seed(1)
    m = MoCo(640, 64, 0.2).cuda()
    for step in range(3):
        mem0 = npy(m.memory).astype(np.float64)
        q = (torch.randn(8, 640, device="cuda") * 0.1).requires_grad_()
        k = torch.randn(8, 640, device="cuda") * 0.1
        logits, labels = m(q, k)
        torch.nn.functional.cross_entropy(logits, labels).backward()
        _, _, dq_o, _ = O.nce_loss_and_grad(npy(q).astype(np.float64), npy(k).astype(np.float64), mem0, 0.2)
        assert rel(npy(q.grad), dq_o) < TOL, step
        assert m._pending == []


# ------------------------------------------------------------------ MoCoAtt: every mode vs the reference's run
MODES = [("all", "all"), ("qk", "qk"), ("dual", "dual"), ("dual2", "dual2"), ("self_qk", "self_qk"), ("self", "self")]


def _mocoatt_setup(g, opt_attn, mode):
    from moma_b200 import CMO, MoCoAtt
    tag = mode + "_"
    torch.manual_seed(500 + len(mode))                  # the generator's seed: identical RNG draws -> identical init
    crit = CMO(Namespace(head="linear", s_dim=8, t_dim=8, feat_dim=32, attn=opt_attn))
    m = MoCoAtt(32, 24, 0.15)
    for n_, p_ in crit.state_dict().items():
        assert np.array_equal(p_.numpy(), g[tag + "sd_" + n_]), n_
    assert np.array_equal(m.memory.numpy(), g[tag + "mem0"])
    m.index = 20
    return crit.cuda(), m.cuda(), tag


@pytest.mark.parametrize("opt_attn,mode", MODES)
def test_mocoatt_modes_golden(fp32, golden, opt_attn, mode):
    """MoMA/mem_moco.py:111-161: logits, d loss/d q, the gradients of every attention parameter that trains in this
    mode (and None for those that do not), the enqueued rows and the pointer."""
    g = golden("kat_mocoatt")
    crit, m, tag = _mocoatt_setup(g, opt_attn, mode)
    q = cu(g[tag + "q"]).requires_grad_()
    k = cu(g[tag + "k"])
    logits, labels = m(q, k, attn=mode, criterion_kd=crit)
    assert tuple(logits.shape) == g[tag + "logits"].shape
    loss = logits.sum() if mode == "dual2" else torch.nn.CrossEntropyLoss()(logits, labels)
    loss.backward()
    assert abs(loss.item() - float(g[tag + "loss"])) < 2 * TOL * abs(float(g[tag + "loss"]))
    if mode != "dual2" and type(logits).__name__ != "LazyLogits":
        assert rel(npy(logits), g[tag + "logits"]) < 2 * TOL
    assert rel(npy(q.grad), g[tag + "dq"]) < 5 * TOL
    for n_, p_ in crit.named_parameters():
        has = bool(g[tag + "hasgrad_" + n_])
        assert (p_.grad is not None) == has, (mode, n_)
        if has:
            # both sides are fp32 runs (the golden is the reference on CPU): 1e-4, plus an absolute floor for
            # gradients that are mathematically zero (the key part of qkv.bias: softmax is shift-invariant)
            want = g[tag + "grad_" + n_]
            assert np.linalg.norm(npy(p_.grad) - want) < 1e-4 * np.linalg.norm(want) + 1e-6, (mode, n_)
    assert m.index == int(g[tag + "index"])
    ids = O.enqueue_ids(6, 20, 24)
    untouched = np.setdiff1d(np.arange(24), ids)
    assert np.array_equal(npy(m.memory)[untouched], g[tag + "mem1"][untouched])       # bit-exact index logic
    assert rel(npy(m.memory)[ids], g[tag + "mem1"][ids]) < 2 * TOL


def test_mocoatt_attended_queue_bf16_no_stale_shadow(bf16, golden):
    """attn='self' attends the whole queue: a NEW tensor every step whose address the caching allocator recycles.
    The bf16 operand of the fused kernel must be cast from THIS step's attended queue (round-1 bug: a shadow cached
    by data_ptr was reused).  Two steps, each against the oracle fed the bf16-rounded operands."""
    from moma_b200 import CMO, MoCoAtt
    torch.manual_seed(3)
    D, K, B, T, H = 64, 256, 32, 0.15, 4
    crit = CMO(Namespace(head="linear", s_dim=8, t_dim=8, feat_dim=D, attn="self")).cuda()
    m = MoCoAtt(D, K, T).cuda()
    r = lambda a: O.round_bf16(a.astype(np.float32)).astype(np.float64)
    sd = {n: npy(p).astype(np.float64) for n, p in crit.state_dict().items()}
    att = lambda name, x: O.attention_forward(x, sd[name + ".qkv.weight"], sd[name + ".qkv.bias"],
                                              sd[name + ".proj.weight"], sd[name + ".proj.bias"], H)
    for step in range(2):
        mem = npy(m.memory).astype(np.float64)
        q = torch.randn(B, D, device="cuda")
        k = torch.randn(B, D, device="cuda")
        with torch.no_grad():
            logits, labels = m(q, k, attn="self", criterion_kd=crit)
            loss = torch.nn.functional.cross_entropy(logits, labels)
        q2, k2, queue2 = att("atts_q", npy(q).astype(np.float64)), att("atts_k", npy(k).astype(np.float64)), \
            att("atts_queue", mem)
        loss_o, _, _, _ = O.nce_loss_and_grad(r(q2), r(k2), r(queue2), T)
        assert abs(loss.item() - loss_o) < 1e-3 * abs(loss_o), step
        assert set(m._shadows) <= {"memory"}           # only the registered buffer's shadow is cached, none for the
                                                       # transient attended queue


# ------------------------------------------------------------------ heads other than 'mlp'
@pytest.mark.parametrize("head", ["mlp_byol", "linear", "none"])
def test_heads_golden(fp32, golden, head):
    """criterion_moco_att.py:269-305 through CMO's _Head: forward, d/dx, parameter gradients, and the BatchNorm
    running statistics the train-mode forward leaves behind."""
    from moma_b200 import CMO
    g = golden("kat_heads")
    torch.manual_seed(700 + len(head))
    crit = CMO(Namespace(head=head, s_dim=20, t_dim=12, feat_dim=16, attn="self"))
    for n_, p_ in crit.embed_s.state_dict().items():
        assert np.array_equal(p_.numpy(), g[f"{head}_sd0_{n_}"]), n_
    emb = crit.embed_s.cuda()
    x = cu(g[f"{head}_x"]).requires_grad_()
    y = emb(x)
    y.backward(cu(g[f"{head}_g"]))
    assert rel(npy(y), g[f"{head}_y"]) < 2 * TOL
    assert rel(npy(x.grad), g[f"{head}_dx"]) < 5 * TOL
    for n_, p_ in emb.named_parameters():
        want = g[f"{head}_grad_{n_}"]                    # absolute floor: the bias in front of a train-mode BatchNorm
        assert np.linalg.norm(npy(p_.grad) - want) < 5 * TOL * np.linalg.norm(want) + 1e-6, n_     # has zero gradient
    for n_, p_ in emb.state_dict().items():
        assert rel(npy(p_), g[f"{head}_sd1_{n_}"]) < 2 * TOL, n_


# ------------------------------------------------------------------ the TF32 library layer of bf16 mode
def test_tf32_teacher_layer_tolerance(bf16):
    """In bf16 mode the no-grad 2048x2048 first layer of embed_t runs as ONE single-pass TF32 library GEMM
    (criterion_moco_att.py:_Head._tf32_library).  Bound its effect: the head's output (unit rows) stays within 1e-3
    of float64, and with grad enabled (any training use) the layer is back on the 3xTF32 kernel (2e-6)."""
    from moma_b200 import CMO
    torch.manual_seed(11)
    crit = CMO(Namespace(head="mlp", s_dim=512, t_dim=2048, feat_dim=128, attn="self")).cuda()
    x = torch.randn(256, 2048, device="cuda")
    sd = {n: npy(p).astype(np.float64) for n, p in crit.embed_t.state_dict().items()}
    want = O.embed_mlp(npy(x).astype(np.float64), sd["1.weight"], sd["1.bias"], sd["3.weight"], sd["3.bias"])
    with torch.no_grad():
        y_tf32 = crit.embed_t(x)
    y_3x = crit.embed_t(x.clone().requires_grad_())
    e_tf32, e_3x = rel(npy(y_tf32), want), rel(npy(y_3x), want)
    assert e_3x < 2e-6, e_3x
    assert e_tf32 < 1e-3, e_tf32
    assert e_tf32 > e_3x                                 # the switch really took the single-pass path


# ------------------------------------------------------------------ checkpointing of the ring pointer
def test_pointer_checkpoint_roundtrip(fp32):
    from moma_b200 import MoCo
    torch.manual_seed(2)
    m = MoCo(32, 40, 0.15).cuda()
    assert sorted(m.state_dict().keys()) == ["memory"]                 # default key set = the reference's
    m(torch.randn(8, 32, device="cuda"), torch.randn(8, 32, device="cuda"))
    m(torch.randn(8, 32, device="cuda"), torch.randn(8, 32, device="cuda"))
    m.checkpoint_pointer = True
    sd = m.state_dict()
    assert int(sd["index"]) == 16
    m2 = MoCo(32, 40, 0.15).cuda()
    m2.use_device_pointer()
    m2.load_state_dict(sd)
    assert m2.index == 16 and int(m2._index_dev.item()) == 16 and torch.equal(m2.memory, m.memory)
    ref_only = {"memory": sd["memory"]}                                # a reference-format checkpoint still loads
    m3 = MoCo(32, 40, 0.15).cuda()
    m3.load_state_dict(ref_only)
    assert m3.index == 0 and torch.equal(m3.memory, m.memory)


# ------------------------------------------------------------------ the graph bench.py times
def _small_cfg(B=64, K=1024, D=128, H=4):
    return dict(B=B, s_dim=96, t_dim=160, D=D, K=K, H=H, desc="test")


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_overlapped_graph_equals_sequential_step(precision):
    """CriterionStep.step_overlapped captured in a CUDA graph (side-stream branches, deferred enqueue, device-resident
    pointer -- exactly what bench.py replays) against CriterionStep.step run eagerly in the reference loop's order:
    identical loss, gradients, queue (bit-exact) and pointer over replays that wrap the ring."""
    import moma_b200
    from moma_b200.graphed import GraphedStep
    from moma_b200.step import CriterionStep
    dev = torch.device("cuda", 0)
    side = torch.cuda.Stream()
    try:
        with torch.cuda.stream(side):
            cfg = _small_cfg(B=96, K=480)
            a = CriterionStep(cfg, 0, 1, dev, precision=precision, ema_shapes=[(64, 3, 3, 3), (64,), (10, 64)])
            b = CriterionStep(cfg, 0, 1, dev, precision=precision, ema_shapes=[(64, 3, 3, 3), (64,), (10, 64)])
            g = GraphedStep(a.step_overlapped, contrast=a.contrast, rows_per_step=96, warmup=3)
            for _ in range(3):
                b.step()
            for i in range(7):                          # 10 steps x 96 rows = 2 wraps of K = 480
                la = g.replay()
                lb = b.step()
                torch.cuda.synchronize()
                assert abs(la.item() - lb.item()) <= 1e-6 * abs(lb.item()), i
                assert torch.allclose(a.feat_s.grad, b.feat_s.grad, rtol=1e-5, atol=1e-9), i
                for (n_, pa), (_, pb) in zip(a.crit.named_parameters(), b.crit.named_parameters()):
                    if pb.grad is not None:
                        assert torch.allclose(pa.grad, pb.grad, rtol=1e-5, atol=1e-9), (i, n_)
                assert torch.equal(a.contrast.memory, b.contrast.memory), i
                assert a.contrast.index == b.contrast.index == ((4 + i) * 96) % 480
                for pa, pb in zip(a.teacher, b.teacher):
                    assert torch.equal(pa, pb)          # backbone EMA, bit-exact
    finally:
        moma_b200.set_precision("bf16")


@pytest.mark.parametrize("W", [2, 4])
def test_emulated_shards_through_a_captured_graph(bf16, W):
    """The sharded loss pass (ShardedMoCo.forward steps 2-4) emulated in ONE process: every 'rank' scores all n
    queries against its cyclic K/W shard (nce_partial -> nce_merge_packed), the packed records of one rank's B queries
    are combined (nce_combine_packed) -- all captured in a CUDA graph and replayed -- against the replicated pass and
    the oracle fed the same bf16 operands."""
    from moma_b200 import ops
    from moma_b200._lib import BF16
    torch.manual_seed(7)
    B, D, K, T = 128, 128, 4096, 0.15
    n = B * W
    queue = torch.nn.functional.normalize(torch.randn(K, D, device="cuda"))
    all_q = torch.randn(n, D, device="cuda") * 0.5
    all_kpos = torch.randn(n, D, device="cuda") * 0.5
    shards = [queue[r::W].contiguous().to(torch.bfloat16) for r in range(W)]
    q16 = all_q.to(torch.bfloat16)
    me = 1                                                   # the emulated rank whose queries are combined
    q32, k32 = all_q[me * B:(me + 1) * B].contiguous(), all_kpos[me * B:(me + 1) * B].contiguous()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        def run():
            recs = []
            for r in range(W):
                stats, Op = ops.nce_partial(q16, shards[r], 1 / T, BF16)
                recs.append(ops.nce_merge_packed(stats, Op).view(W, B, D + 4)[me])
            recv = torch.stack(recs)                         # [W(src), B, D + 4]: what the all-to-all delivers
            return ops.nce_combine_packed(recv, q32, k32, 1 / T, True, 1.0 / B)
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            rows, dq, pim, mx, loss, acc = run()
        for _ in range(2):
            graph.replay()
        torch.cuda.synchronize()
    r_ = lambda t: O.round_bf16(npy(t)).astype(np.float64)
    loss_o, rows_o, dq_o, pim_o = O.nce_loss_and_grad(r_(q32), r_(k32), r_(queue), T)
    assert abs(loss.item() - loss_o) < 1e-3 * abs(loss_o)
    assert rel(npy(rows), rows_o) < 1e-3 and rel(npy(dq), dq_o) < 1e-3
    assert np.array_equal(npy(pim).astype(bool), pim_o)
    # and against the replicated single pass over the whole queue
    nce = ops.nce_rows(q32, k32, queue, queue.to(torch.bfloat16), T, "bf16")
    assert abs(loss.item() - nce.loss.item()) < 5e-4 * abs(loss_o)


# ------------------------------------------------------------------ one-launch InfoNCE (in-kernel combine)
@pytest.mark.parametrize("B,D,K", [(256, 128, 16384), (512, 128, 65536), (200, 128, 1000), (512, 64, 8192), (128, 128, 128)])
def test_fused_nce_equals_three_launch_path_and_oracle(bf16, B, D, K):
    """moma_nce_fused (tcgen05 pass + combine + finalize in ONE launch) against moma_nce_partial + moma_nce_combine on
    the same operands, and against the oracle fed the bf16-rounded operands; repeated launches (the ticket counters
    must come back to zero), ragged K and a partial last query tile included."""
    from moma_b200 import _lib, ops
    from moma_b200._lib import BF16
    lib = _lib.load()
    torch.manual_seed(B + K)
    T = 0.15
    q = torch.randn(B, D, device="cuda") * 0.5
    k = torch.randn(B, D, device="cuda") * 0.5
    queue = torch.nn.functional.normalize(torch.randn(K, D, device="cuda"))
    q16, queue16 = q.to(torch.bfloat16), queue.to(torch.bfloat16)
    assert lib.moma_nce_fused_supported(B, D, K) == 1
    stats, Op = ops.nce_partial(q16, queue16, 1 / T, BF16)
    rows0, dq0, pim0, mx0, loss0, acc0 = ops.nce_combine(stats, Op, q, k, 1 / T, True, 1.0 / B, want_mean=True)
    for rep in range(3):
        rows, dq, pim, mx, loss, acc = ops.nce_fused(q16, queue16, q, k, 1 / T, True, 1.0 / B)
        torch.cuda.synchronize()
        assert torch.allclose(rows, rows0, rtol=2e-6, atol=2e-6), rep
        assert rel(npy(dq), npy(dq0)) < 1e-5, rep                      # same partials, different summation grouping
        assert torch.equal(pim, pim0) and torch.allclose(mx, mx0, rtol=1e-6)
        assert abs(loss.item() - loss0.item()) < 2e-6 * abs(loss0.item()) and acc.item() == acc0.item()
    assert int(ops._FUSE_COUNTERS[q.device][0].abs().sum().item()) == 0          # every block back to zero
    # and through the module API with the one-launch path switched on
    import os
    from moma_b200 import MoCo
    os.environ["MOMA_B200_NCE_FUSED"] = "1"
    try:
        m = MoCo(D, K, T).cuda()
        with torch.no_grad():
            m.memory.copy_(queue)
        m.invalidate_shadows()
        qq = q.clone().requires_grad_()
        logits, labels = m(qq, k)
        lm = torch.nn.functional.cross_entropy(logits, labels)
        lm.backward()
        assert abs(lm.item() - loss0.item()) < 2e-6 * abs(loss0.item()) and rel(npy(qq.grad), npy(dq0)) < 1e-5
    finally:
        os.environ["MOMA_B200_NCE_FUSED"] = "0"
    r_ = lambda t: O.round_bf16(npy(t)).astype(np.float64)
    loss_o, rows_o, dq_o, pim_o = O.nce_loss_and_grad(r_(q), r_(k), r_(queue), T)
    assert abs(loss.item() - loss_o) < 1e-3 * abs(loss_o) and rel(npy(dq), dq_o) < 1e-3
    assert np.array_equal(npy(pim).astype(bool), pim_o)
    # the packed variant == partial + merge_packed
    packed = ops.nce_fused_packed(q16, queue16, 1 / T)
    want = ops.nce_merge_packed(stats, Op)
    torch.cuda.synchronize()
    assert rel(npy(packed[:, :D + 3]), npy(want[:, :D + 3])) < 1e-5


# ------------------------------------------------------------------ classification CE + KD + top-1 in one launch (8f-3)
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_cls_kd_fused_golden(golden, tag):
    """ops.cls_kd_losses == (CrossEntropyLoss, DistillKL(T), accuracy) of the reference on its own inputs, values and
    gradients (through an arbitrary weighting cls * l_cls + div * l_div as at helper/loops_moma.py:345)."""
    from moma_b200 import ops
    g = golden("kat_kd")
    ys = cu(g[f"{tag}_ys"]).requires_grad_()
    yt, lab, T = cu(g[f"{tag}_yt"]), cu(g[f"{tag}_lab"]), float(g[f"{tag}_T"])
    l_cls, l_div, acc = ops.cls_kd_losses(ys, yt, lab, T)
    assert abs(l_cls.item() - float(g[f"{tag}_cls"])) < 2e-6 * abs(float(g[f"{tag}_cls"]))
    assert abs(l_div.item() - float(g[f"{tag}_div"])) < 5e-6 * abs(float(g[f"{tag}_div"]))
    assert acc.item() == pytest.approx(float(g[f"{tag}_acc"][0]))
    (0.7 * l_cls + 1.3 * l_div).backward()
    want = 0.7 * g[f"{tag}_gcls"] + 1.3 * g[f"{tag}_gdiv"]
    assert rel(npy(ys.grad), want) < TOL


# ------------------------------------------------------------------ SGD + EMA in one pass (8f-2)
def test_fused_sgd_ema_bit_exact():
    """ops.FusedSgdEma.step() == torch.optim.SGD(momentum 0.9, wd 5e-4).step() followed by momentum_update, bit for bit,
    against torch's CUDA kernels and against the C oracle; odd sizes, an unaligned view, re-allocated gradients."""
    from moma_b200 import ContrastTrainer, ops
    torch.manual_seed(9)
    shapes = [(7,), (33, 17), (4097,), (64, 3, 3, 3), (100003,), (512, 512)]
    base = [torch.randn(*s) for s in shapes]
    ema0 = [torch.randn(*s) for s in shapes]
    pa = torch.nn.ParameterList([torch.nn.Parameter(b.clone()) for b in base]).cuda()       # fused
    ea = torch.nn.ParameterList([torch.nn.Parameter(e.clone()) for e in ema0]).cuda()
    pb = torch.nn.ParameterList([torch.nn.Parameter(b.clone()) for b in base]).cuda()       # torch SGD + momentum_update
    eb = torch.nn.ParameterList([torch.nn.Parameter(e.clone()) for e in ema0]).cuda()
    fused = ops.FusedSgdEma(pa, ea, lr=0.05, momentum=0.9, weight_decay=5e-4, m=0.999)
    opt = torch.optim.SGD(pb, lr=0.05, momentum=0.9, weight_decay=5e-4)
    P = [b.numpy().copy() for b in base]; E = [e.numpy().copy() for e in ema0]; Bf = [np.zeros_like(p) for p in P]
    for step in range(4):
        grads = [torch.randn(*s) for s in shapes]
        for p, q, g in zip(pa, pb, grads):
            p.grad = g.clone().cuda()                     # new tensors every step: the pointer table must follow
            q.grad = g.clone().cuda()
        fused.step()
        opt.step()
        ContrastTrainer.momentum_update(pb, eb, 0.999)
        O.sgd_ema_step(P, [g.numpy().copy() for g in grads], Bf, E, 0.05, 0.9, 5e-4, step == 0, 0.999)
        for i in range(len(shapes)):
            assert torch.equal(pa[i], pb[i]), (step, i)
            assert torch.equal(ea[i], eb[i]), (step, i)
            assert np.array_equal(npy(pa[i]), P[i]) and np.array_equal(npy(ea[i]), E[i]), (step, i)
    with pytest.raises(RuntimeError):
        ops.FusedSgdEma([torch.zeros(3, 4, device="cuda")], [torch.zeros(4, 3, device="cuda")], lr=0.1)
