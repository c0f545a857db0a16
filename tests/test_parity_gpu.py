"""GPU parity tests: every kernel behind the C ABI against the CPU oracle (pinned to the
reference by tests/test_oracle_golden.py) and against the committed golden vectors.

Bars: bit-exact for ids / pointer / enqueued rows / EMA; <= 1e-5 relative (norm-wise for
gradients) for the FP32 floating-point kernels.  The bf16 tensor-core kernel is in test_tc_gpu.py.
"""
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


@pytest.fixture(scope="module")
def ops():
    import moma_b200
    from moma_b200 import ops as _ops
    moma_b200.set_precision("fp32")
    yield _ops
    moma_b200.set_precision("bf16")


# ------------------------------------------------------------------------------ EMA
def test_ema_golden_bit_exact(ops, golden):
    g = golden("kat_ema")
    n = int(g["n"])
    src = [cu(g[f"src{i}"]) for i in range(n)]
    ema = [cu(g[f"ema{i}_0"]) for i in range(n)]
    for step in (1, 2, 3):
        ops.ema_update(src, ema, 0.999)
        for i in range(n):
            assert np.array_equal(npy(ema[i]), g[f"ema{i}_{step}"]), (i, step)
    emb = [cu(g[f"emb{i}_0"]) for i in range(n)]
    ops.ema_update(src, emb, 0.5)
    for i in range(n):
        assert np.array_equal(npy(emb[i]), g[f"emb{i}_1"])
    with pytest.raises(RuntimeError):
        ops.ema_update([torch.zeros(3, 4).cuda()], [torch.zeros(4, 3).cuda()], 0.9)


def test_ema_large_ragged_and_unaligned(ops):
    rng = np.random.default_rng(0)
    sizes = [1, 3, 8191, 8192, 8193, 100003, 2_000_001, 64, 5]
    src = [rng.standard_normal(s).astype(np.float32) for s in sizes]
    ema = [(rng.standard_normal(s) * 7).astype(np.float32) for s in sizes]
    want = [e.copy() for e in ema]
    O.momentum_update(want, src, 0.999)
    ds, de = [cu(s) for s in src], [cu(e) for e in ema]
    ops.ema_update(ds, de, 0.999)
    for w, d in zip(want, de):
        assert np.array_equal(npy(d), w)
    # views at a 4-byte (not 16-byte) aligned offset take the scalar path
    big_s, big_e = cu(rng.standard_normal(5001).astype(np.float32)), cu(rng.standard_normal(5001).astype(np.float32))
    w = npy(big_e)[1:].copy()
    O.momentum_update([w], [npy(big_s)[1:].copy()], 0.9)
    ops.ema_update([big_s[1:]], [big_e[1:]], 0.9)
    assert np.array_equal(npy(big_e)[1:], w)


def test_momentum_update_module_api(ops):
    """ContrastTrainer.momentum_update(model, model_ema, m) on real modules (+ plan reuse)."""
    from moma_b200 import ContrastTrainer
    torch.manual_seed(0)
    a = torch.nn.Sequential(torch.nn.Linear(33, 17), torch.nn.BatchNorm1d(17), torch.nn.Linear(17, 5)).cuda()
    b = torch.nn.Sequential(torch.nn.Linear(33, 17), torch.nn.BatchNorm1d(17), torch.nn.Linear(17, 5)).cuda()
    want = [npy(p).copy() for p in b.parameters()]
    for _ in range(2):
        O.momentum_update(want, [npy(p) for p in a.parameters()], 0.999)
        ContrastTrainer.momentum_update(a, b, 0.999)
    for w, p in zip(want, b.parameters()):
        assert np.array_equal(npy(p), w)


# ------------------------------------------------------------------------ Normalize
def test_normalize_golden(ops, golden):
    g = golden("kat_normalize")
    x = cu(g["x"]).requires_grad_()
    y = ops.l2_normalize(x)
    assert np.allclose(npy(y), g["y"], rtol=2e-6, atol=1e-7)
    y.backward(cu(g["g"]))
    for r in range(12):
        assert rel(npy(x.grad)[r], g["dx"][r]) < TOL, r


def test_normalize_random(ops):
    rng = np.random.default_rng(1)
    for rows, D in ((1, 4), (37, 128), (256, 512), (5, 2048)):
        x = rng.standard_normal((rows, D)).astype(np.float32)
        gy = rng.standard_normal((rows, D)).astype(np.float32)
        xt = cu(x).requires_grad_()
        y = ops.l2_normalize(xt)
        y.backward(cu(gy))
        assert rel(npy(y), O.normalize(x.astype(np.float64))) < 1e-6
        assert rel(npy(xt.grad), O.normalize_backward(x.astype(np.float64), gy.astype(np.float64))) < TOL


# -------------------------------------------------------------------------- enqueue
def test_enqueue_ids_and_pointer_golden(ops, golden):
    g = golden("kat_pointer")
    from moma_b200 import MoCo
    for K, n in ((4096, 96), (10, 4), (65536, 512), (131072, 1024), (7, 7), (12, 5)):
        idx = 0
        seq = g[f"ptr_K{K}_n{n}_index"]
        for s in (0, 1, len(seq) // 2, len(seq) - 1):
            start = 0 if s == 0 else int(seq[s - 1])
            ids = ops.enqueue_ids(n, start, K, "cuda").cpu().numpy()
            assert ids.dtype == np.int64
            assert np.array_equal(ids, O.enqueue_ids(n, start, K))
            assert ids[0] == g[f"ptr_K{K}_n{n}_first"][s] and ids[-1] == g[f"ptr_K{K}_n{n}_last"][s]
    assert np.array_equal(ops.enqueue_ids(96, 4032, 4096, "cuda").cpu().numpy(), g["ids_wrap"])
    m = MoCo(4, 10, 0.07).cuda()
    m.memory.zero_()
    m.index = 8
    k = torch.arange(16, dtype=torch.float32).view(4, 4).cuda() + 1
    m._update_memory(k, m.memory); m._update_pointer(4)
    assert np.array_equal(npy(m.memory), g["kat2_mem"]) and m.index == int(g["kat2_index"]) == 2
    # KAT3 pointer sequence through the module
    m = MoCo(4, 4096, 0.07)
    for want in g["ptr_K4096_n96_index"]:
        m._update_pointer(96)
        assert m.index == want


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_enqueue_sharded_matches_replicated(ops, world):
    rng = np.random.default_rng(2)
    K, D, n = 64 * world, 32, 24
    full = rng.standard_normal((K, D)).astype(np.float32)
    shards = [cu(full[r::world].copy()) for r in range(world)]
    shadows = [torch.zeros((K // world, D), dtype=torch.bfloat16, device="cuda") for _ in range(world)]
    for r in range(world):
        ops.cast_bf16(shards[r], shadows[r])
    index = K - 10                         # wraps
    for step in range(5):
        keys = rng.standard_normal((n, D)).astype(np.float32)
        O.update_memory(full, keys, index)
        for r in range(world):
            ops.enqueue(cu(keys), shards[r], shadows[r], K, index, rank=r, world=world)
        index = O.update_pointer(index, n, K)
    re = np.empty_like(full)
    for r in range(world):
        re[r::world] = npy(shards[r])
    assert np.array_equal(re, full)                                   # rows copied verbatim
    for r in range(world):
        assert np.array_equal(npy(shadows[r]), O.round_bf16(full[r::world]))   # shadow = RNE bf16


def test_enqueue_fused_normalize_and_errors(ops):
    rng = np.random.default_rng(3)
    K, D, n = 40, 64, 16
    q = cu(np.zeros((K, D), np.float32))
    keys = rng.standard_normal((n, D)).astype(np.float32)
    ops.enqueue(cu(keys), q, None, K, 30, normalize=True)
    want = np.zeros((K, D), np.float32)
    O.update_memory(want, O.normalize(keys), 30)
    assert np.allclose(npy(q), want, rtol=2e-6, atol=1e-7)
    with pytest.raises(RuntimeError, match="duplicate ids"):
        ops.enqueue(cu(np.zeros((K + 1, D), np.float32)), q, None, K, 0)
    with pytest.raises(RuntimeError):
        ops.enqueue(torch.zeros(4, D), q, None, K, 0)          # CPU tensor: no fallback


# ----------------------------------------------------------------- InfoNCE (fp32 path)
def nce_gpu(ops, q, k, queue, T, n_splits=None):
    from moma_b200._lib import F32
    qt, kt, mt = cu(q), cu(k), cu(queue)
    stats, Opart = ops.nce_partial(qt, mt, 1.0 / T, F32, n_splits)
    rows, dq, pim, mx = ops.nce_combine(stats, Opart, qt, kt, 1.0 / T)
    return npy(rows), npy(dq), pim.cpu().numpy(), npy(mx)


def test_nce_golden_small(ops, golden):
    g = golden("kat_moco")
    rows, dq, pim, mx = nce_gpu(ops, g["s_q"], g["s_k"], g["s_mem0"], 0.07)
    assert abs(rows.mean() - float(g["s_loss"])) < TOL * abs(float(g["s_loss"]))
    assert rel(dq / 8, g["s_dq"]) < TOL
    assert pim.mean() * 100 == pytest.approx(float(g["s_acc"][0]))
    assert np.allclose(mx, g["s_logits"].max(axis=1), rtol=1e-5)


def test_nce_kat1_through_module(ops, golden):
    """SURVEY 8c KAT1 end to end: same seed -> same queue (RNG parity), loss 8.53796864,
    grad norm 1.18056488, index 32, rows 0..31 == k bit-exact."""
    from moma_b200 import MoCo
    g = golden("kat_moco")
    torch.manual_seed(0)
    m = MoCo(128, 4096, 0.15)
    assert np.array_equal(m.memory.numpy()[:40], g["kat1_mem_rows"])      # initial queue: same RNG draw
    assert abs(m.memory.double().sum().item() - float(g["kat1_mem_sum"])) < 1e-9
    m = m.cuda()
    q = cu(g["kat1_q"]).requires_grad_()
    k = cu(g["kat1_k"])
    logits, labels = m(q, k)
    assert tuple(logits.shape) == (32, 4097) and labels.dtype == torch.int64 and int(labels.abs().sum()) == 0
    loss = torch.nn.CrossEntropyLoss()(logits, labels)
    loss.backward()
    assert abs(loss.item() - 8.53796864) < 8.6e-5            # 1e-5 relative
    assert abs(q.grad.norm().item() - 1.18056488) < 1.2e-5
    assert rel(npy(q.grad), g["kat1_dq"]) < TOL
    assert m.index == 32
    assert np.array_equal(npy(m.memory)[:32], g["kat1_k"])
    assert np.array_equal(npy(m.memory)[32:40], g["kat1_mem_rows"][32:])


@pytest.mark.parametrize("B,D,K,T", [(1, 16, 33, 0.2), (5, 48, 100, 0.07), (32, 128, 4096, 0.15),
                                     (70, 256, 1000, 0.15), (33, 512, 300, 0.07), (256, 128, 16384, 0.15)])
def test_nce_fp32_random(ops, B, D, K, T):
    rng = np.random.default_rng(B * 7 + D)
    q = (rng.standard_normal((B, D)) * 0.7).astype(np.float32)          # un-normalised (post-attention regime)
    k = (rng.standard_normal((B, D)) * 0.7).astype(np.float32)
    queue = O.normalize(rng.standard_normal((K, D))).astype(np.float32)
    loss, rows_o, dq_o, pim_o = O.nce_loss_and_grad(q.astype(np.float64), k.astype(np.float64), queue.astype(np.float64), T)
    rows, dq, pim, mx = nce_gpu(ops, q, k, queue, T)
    assert rel(rows, rows_o) < TOL
    assert rel(dq / B, dq_o) < TOL
    assert np.array_equal(pim.astype(bool), pim_o)
    # split-count invariance (partials merge associatively)
    rows1, dq1, _, _ = nce_gpu(ops, q, k, queue, T, n_splits=1)
    rows3, dq3, _, _ = nce_gpu(ops, q, k, queue, T, n_splits=3)
    assert rel(rows1, rows_o) < TOL and rel(rows3, rows_o) < TOL and rel(dq3 / B, dq_o) < TOL


def test_nce_logits_materialise(ops, golden):
    g = golden("kat_moco")
    dense = ops.nce_logits(cu(g["s_q"]), cu(g["s_k"]), cu(g["s_mem0"]), 0.07)
    assert rel(npy(dense), g["s_logits"]) < 2e-6
    from moma_b200 import MoCo
    m = MoCo(16, 32, 0.2).cuda()
    m.memory.copy_(cu(g["b1_mem0"]))
    lg = m._compute_logit(cu(g["b1_q"]), cu(g["b1_k"]), m.memory)
    assert tuple(lg.shape) == (33,) and rel(npy(lg), g["b1_logits"]) < 2e-6        # KAT4 squeeze
    qk = m._compute_logit_qk(cu(g["s_q"][:, :16].copy()), cu(g["s_k"][:, :16].copy()))
    assert rel(npy(qk), O.compute_logit_qk(g["s_q"][:, :16], g["s_k"][:, :16], 0.2)) < 2e-6


def test_lazy_logits_on_gpu(ops, golden):
    """CrossEntropy / top-1 served lazily == dense path; other uses materialise (when allowed)."""
    from moma_b200 import MoCo, LazyLogits, ContrastTrainer
    g = golden("kat_moco")
    m = MoCo(32, 64, 0.07).cuda()
    m.memory.copy_(cu(g["s_mem0"]))
    m.index = 56
    m.track_overwritten = True
    q = cu(g["s_q"]).requires_grad_()
    logits, labels = m(q, cu(g["s_k"]), cu(g["s_allk"]))
    assert isinstance(logits, LazyLogits) and isinstance(logits, torch.Tensor)
    losses, accs = ContrastTrainer._compute_loss_accuracy([logits], labels, torch.nn.CrossEntropyLoss())
    assert abs(losses[0].item() - float(g["s_loss"])) < TOL * abs(float(g["s_loss"]))
    assert accs[0].item() == pytest.approx(float(g["s_acc"][0]))
    losses[0].backward()
    assert rel(npy(q.grad), g["s_dq"]) < TOL
    # enqueue happened: wrap 56..63, 0..15 ; pointer advanced
    assert np.array_equal(npy(m.memory), g["s_mem1"]) and m.index == int(g["s_index"])
    # late materialisation still sees the PRE-enqueue queue (the reference's clone)
    dense = logits.materialize()
    assert rel(npy(dense), g["s_logits"]) < 2e-6
    _, pred = logits.topk(1, 1, True, True)
    assert ((pred.squeeze(1) == 0).cpu().numpy() == (g["s_logits"].argmax(1) == 0)).all()
    # default mode refuses a late materialisation loudly
    m2 = MoCo(32, 64, 0.07).cuda()
    lg2, _ = m2(cu(g["s_q"]), cu(g["s_k"]))
    with pytest.raises(RuntimeError, match="after the queue was updated"):
        (lg2 + 1).sum()


def test_dual_queue_variants(ops, golden):
    from moma_b200 import MoCoST, MoCoSSTT
    g = golden("kat_dual")
    ce = torch.nn.CrossEntropyLoss()
    m = MoCoST(16, 24, 0.1).cuda()
    m.memory_s.copy_(cu(g["st_ms0"])); m.memory_t.copy_(cu(g["st_mt0"]))
    m.index = 22
    lss, lst, lab = m(cu(g["st_q"]), cu(g["st_k"]), cu(g["st_kt"]))
    want_ss, _ = O.cross_entropy_zero_label(g["st_lss"].astype(np.float64))
    want_st, _ = O.cross_entropy_zero_label(g["st_lst"].astype(np.float64))
    assert abs(ce(lss, lab).item() - want_ss) < TOL * want_ss and abs(ce(lst, lab).item() - want_st) < TOL * want_st
    assert np.array_equal(npy(m.memory_s), g["st_ms1"]) and np.array_equal(npy(m.memory_t), g["st_mt1"])
    assert m.index == int(g["st_index"])
    m = MoCoSSTT(16, 24, 0.1).cuda()
    m.memory_s.copy_(cu(g["sstt_ms0"])); m.memory_t.copy_(cu(g["sstt_mt0"]))
    out = m(cu(g["st_q"]), cu(g["st_k"]), cu(g["sstt_qt"]), cu(g["st_kt"]))
    assert len(out) == 5
    for lg, key in zip(out[:4], ("sstt_lss", "sstt_lst", "sstt_lts", "sstt_ltt")):
        want, _ = O.cross_entropy_zero_label(g[key].astype(np.float64))
        assert abs(ce(lg, out[4]).item() - want) < TOL * want
    assert m.index == int(g["sstt_index"])


# ------------------------------------------------------------------------- attention
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_attention_golden(ops, golden, tag):
    g = golden("kat_attention")
    H = int(g[f"{tag}_H"])
    has_b = f"{tag}_bqkv" in g.files
    x = cu(g[f"{tag}_x"]).requires_grad_()
    wq = cu(g[f"{tag}_wqkv"]).requires_grad_()
    bq = cu(g[f"{tag}_bqkv"]).requires_grad_() if has_b else None
    wp = cu(g[f"{tag}_wproj"]).requires_grad_()
    bp = cu(g[f"{tag}_bproj"]).requires_grad_()
    y = ops.attention(x, wq, bq, wp, bp, H)
    assert rel(npy(y), g[f"{tag}_y"]) < TOL
    y.backward(cu(g[f"{tag}_dy"]))
    assert rel(npy(x.grad), g[f"{tag}_dx"]) < TOL
    assert rel(npy(wq.grad), g[f"{tag}_dwqkv"]) < TOL
    assert rel(npy(wp.grad), g[f"{tag}_dwproj"]) < TOL
    assert rel(npy(bp.grad), g[f"{tag}_dbproj"]) < TOL
    if has_b:
        assert rel(npy(bq.grad), g[f"{tag}_dbqkv"]) < TOL


@pytest.mark.parametrize("N,C,H", [(1, 64, 4), (100, 128, 8), (256, 128, 4), (130, 256, 4), (65, 512, 4), (1024, 256, 4)])
def test_attention_random(ops, N, C, H):
    rng = np.random.default_rng(N + C)
    x = rng.standard_normal((N, C)).astype(np.float32)
    wq = (rng.standard_normal((3 * C, C)) / np.sqrt(C)).astype(np.float32)
    bq = (rng.standard_normal(3 * C) * 0.1).astype(np.float32)
    wp = (rng.standard_normal((C, C)) / np.sqrt(C)).astype(np.float32)
    bp = (rng.standard_normal(C) * 0.1).astype(np.float32)
    dy = rng.standard_normal((N, C)).astype(np.float32)
    f64 = lambda *a: [v.astype(np.float64) for v in a]
    yo = O.attention_forward(*f64(x, wq, bq, wp, bp), H)
    go = O.attention_backward(*f64(x, wq, bq, wp, bp), H, dy.astype(np.float64))
    t = [cu(v).requires_grad_() for v in (x, wq, bq, wp, bp)]
    y = ops.attention(*t, H)
    y.backward(cu(dy))
    assert rel(npy(y), yo) < TOL
    for got, key in zip(t, ("dx", "d_wqkv", "d_bqkv", "d_wproj", "d_bproj")):
        assert rel(npy(got.grad), go[key]) < 2 * TOL, key


def test_attention_module_and_viz(ops, golden):
    from moma_b200 import Attention, Attention2, Attention_viz
    g = golden("kat_attention")
    att = Attention(32, num_heads=4, qkv_bias=True).cuda()
    with torch.no_grad():
        att.qkv.weight.copy_(cu(g["a_wqkv"])); att.qkv.bias.copy_(cu(g["a_bqkv"]))
        att.proj.weight.copy_(cu(g["a_wproj"])); att.proj.bias.copy_(cu(g["a_bproj"]))
    assert rel(npy(att(cu(g["a_x"]))), g["a_y"]) < TOL
    assert att.scale == float(g["a_scale"])
    viz = Attention_viz(32, num_heads=4, qkv_bias=True).cuda()
    viz.load_state_dict(att.state_dict())
    y, probs = viz(cu(g["a_x"]))
    _, cache = O.attention_forward(g["a_x"].astype(np.float64), g["a_wqkv"], g["a_bqkv"], g["a_wproj"], g["a_bproj"],
                                   4, return_cache=True)
    assert tuple(probs.shape) == (1, 4, 24, 24) and rel(npy(probs)[0], cache["attn"]) < TOL
    a2 = Attention2(32, num_heads=4, qkv_bias=True).cuda()
    out = a2(cu(g["a_x"]))
    assert tuple(out.shape) == (24, 32)


# -------------------------------------------------- whole criterion step vs the reference run
def test_criterion_step_golden(ops, golden):
    """helper/loops_moma.py:308-335 with our modules, 3 steps, against the reference's run
    (same seed -> identical initial queue and parameters)."""
    from argparse import Namespace
    from moma_b200 import CMO, ContrastTrainer, build_mem
    g = golden("criterion_step")
    torch.manual_seed(12345)
    opt = Namespace(head="mlp", s_dim=24, t_dim=24, feat_dim=32, attn="self", mem="MoCo", nce_k=40, nce_t=0.15,
                    alpha=0.999)
    contrast = build_mem(opt)
    crit = CMO(opt)
    assert np.array_equal(contrast.memory.numpy(), g["mem0"])                  # RNG parity of the queue
    for name, p in crit.state_dict().items():
        assert np.array_equal(p.numpy(), g["sd0_" + name]), name               # and of every parameter
    contrast, crit = contrast.cuda(), crit.cuda()
    sgd = torch.optim.SGD([p for n, p in crit.named_parameters() if not n.startswith("embed_t")], lr=0.05)
    ce = torch.nn.CrossEntropyLoss()
    for st in range(3):
        feat_s = cu(g[f"st{st}_feat_s"]).requires_grad_()
        feat_t = cu(g[f"st{st}_feat_t"])
        crit.embed_t.eval()
        ContrastTrainer.momentum_update(crit.embed_s, crit.embed_t, opt.alpha)
        with torch.no_grad():
            k = crit.embed_t(feat_t)
        assert rel(npy(k), g[f"st{st}_k"]) < TOL
        f_s = crit.embed_s(feat_s)
        f_s = crit.atts_q(f_s); k2 = crit.atts_k(k); allk2 = crit.atts_queue(k)
        assert rel(npy(f_s), g[f"st{st}_f_s"]) < 2 * TOL and rel(npy(allk2), g[f"st{st}_allk2"]) < 2 * TOL
        output = contrast(q=f_s, k=k2, all_k=allk2)
        losses, accs = ContrastTrainer._compute_loss_accuracy(output[:-1], output[-1], ce)
        loss = losses[0]
        sgd.zero_grad(); loss.backward()
        assert abs(loss.item() - float(g[f"st{st}_loss"])) < 2 * TOL * abs(float(g[f"st{st}_loss"]))
        assert accs[0].item() == pytest.approx(float(g[f"st{st}_acc"][0]))
        assert rel(npy(feat_s.grad), g[f"st{st}_dfeat_s"]) < 5 * TOL
        for n, p in crit.named_parameters():
            has = bool(g[f"st{st}_hasgrad_{n}"])
            assert (p.grad is not None) == has, n                              # KAT6
            if has:
                assert rel(npy(p.grad), g[f"st{st}_grad_{n}"]) < 5 * TOL, n
        assert contrast.index == int(g[f"st{st}_index"])
        assert rel(npy(contrast.memory), g[f"st{st}_mem"]) < 2 * TOL
        sgd.step()
        for n, p in crit.state_dict().items():
            assert rel(npy(p), g[f"st{st}_sd_{n}"]) < 2 * TOL, n


# ------------------------------------------------- full-size, size-independent properties
def test_full_size_properties(ops):
    """BASELINE C2/C3 sizes: fused loss == CE of the materialised logits; split invariance;
    gradient rows are (convex combination of queue/k rows - k)/T; enqueue/pointer round trip."""
    from moma_b200._lib import F32
    torch.manual_seed(5)
    for B, D, K in ((256, 128, 16384), (512, 128, 65536)):
        T = 0.15
        q = torch.randn(B, D, device="cuda") * 0.6
        k = torch.randn(B, D, device="cuda") * 0.6
        queue = torch.nn.functional.normalize(torch.randn(K, D, device="cuda"))
        stats, Op = ops.nce_partial(q, queue, 1 / T, F32)
        rows, dq, pim, mx = ops.nce_combine(stats, Op, q, k, 1 / T)
        dense = ops.nce_logits(q, k, queue, T)
        ref_rows = torch.nn.functional.cross_entropy(dense.double(), torch.zeros(B, dtype=torch.long, device="cuda"),
                                                     reduction="none")
        assert rel(npy(rows), ref_rows.cpu().numpy()) < TOL
        assert torch.equal(pim.bool(), dense.argmax(1) == 0)
        stats1, Op1 = ops.nce_partial(q, queue, 1 / T, F32, 1)
        rows1, dq1, _, _ = ops.nce_combine(stats1, Op1, q, k, 1 / T)
        assert rel(npy(rows1), npy(rows)) < TOL and rel(npy(dq1), npy(dq)) < TOL
        # dq_unit * T + k = sum_j p_j c_j : inside the convex hull -> norm bounded by max row norm
        mix = dq * T + k
        assert float(mix.norm(dim=1).max()) <= float(torch.maximum(queue.norm(dim=1).max(), k.norm(dim=1).max())) * (1 + 1e-4)
        # ring: K/n enqueues bring the pointer back and replace every row exactly once
        n = B
        shadow = torch.empty(K, D, dtype=torch.bfloat16, device="cuda")
        idx = 12345 % K
        start = idx
        stamp = torch.arange(n, device="cuda", dtype=torch.float32).unsqueeze(1).expand(n, D).contiguous()
        for s in range(K // n):
            ops.enqueue(stamp + s * n, queue, shadow, K, idx)
            idx = O.update_pointer(idx, n, K)
        assert idx == start
        want = (torch.arange(K, device="cuda") - start) % K
        assert torch.equal(queue[:, 0].long(), want)


# ------------------------------------------------------------- CUDA-graph replay == eager
def test_graphed_step_matches_eager(ops):
    """The captured graph (device-resident queue pointer) reproduces the eager step exactly:
    same losses, gradients, queue contents and pointer over several replays incl. a wrap."""
    from argparse import Namespace
    from moma_b200 import CMO, ContrastTrainer, build_mem
    from moma_b200.graphed import GraphedStep

    def make():
        torch.manual_seed(3)
        opt = Namespace(head="mlp", s_dim=32, t_dim=32, feat_dim=64, attn="self", mem="MoCo", nce_k=160, nce_t=0.15)
        contrast, crit = build_mem(opt).cuda(), CMO(opt).cuda()
        fs = torch.randn(48, 32, device="cuda", requires_grad=True)
        ft = torch.randn(48, 32, device="cuda")
        ce = torch.nn.CrossEntropyLoss()

        def step():
            ContrastTrainer.momentum_update(crit.embed_s, crit.embed_t, 0.999)
            with torch.no_grad():
                k = crit.embed_t(ft)
            q = crit.atts_q(crit.embed_s(fs)); k2 = crit.atts_k(k); ak = crit.atts_queue(k)
            logits, labels = contrast(q=q, k=k2, all_k=ak)
            loss = ce(logits, labels)
            fs.grad = None
            for p in crit.parameters():
                p.grad = None
            loss.backward()
            return loss
        return contrast, crit, fs, step

    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        c_e, crit_e, fs_e, step_e = make()
        eager = []
        for _ in range(3 + 5):                        # GraphedStep warms up 3 times before capturing
            l = step_e()
            eager.append((l.item(), fs_e.grad.clone(), c_e.memory.clone(), c_e.index))
        c_g, crit_g, fs_g, step_g = make()
        g = GraphedStep(step_g, contrast=c_g, rows_per_step=48, warmup=3)
        assert c_g.index == eager[2][3]
        for i in range(5):
            l = g.replay()
            torch.cuda.synchronize()
            want = eager[3 + i]
            assert abs(l.item() - want[0]) < 1e-6 * abs(want[0])
            assert torch.allclose(fs_g.grad, want[1], rtol=1e-5, atol=1e-8)
            assert torch.equal(c_g.memory, want[2]) and c_g.index == want[3]
        assert c_g.index == (8 * 48) % 160                                   # wrapped twice


def test_packed_merge_and_combine(ops):
    """The packed partial records used for the cross-rank exchange give the same result as the
    unpacked path, and emulated shards (cyclic split of the queue) merge to the replicated answer."""
    from moma_b200._lib import F32
    rng = np.random.default_rng(11)
    B, D, K, T, W = 96, 64, 1024, 0.15, 4
    q = (rng.standard_normal((B, D)) * 0.7).astype(np.float32)
    k = (rng.standard_normal((B, D)) * 0.7).astype(np.float32)
    queue = O.normalize(rng.standard_normal((K, D))).astype(np.float32)
    loss_o, rows_o, dq_o, pim_o = O.nce_loss_and_grad(q.astype(np.float64), k.astype(np.float64), queue.astype(np.float64), T)
    qt, kt = cu(q), cu(k)
    recs = []
    for r in range(W):                                   # rank r owns rows r, r + W, ...
        shard = cu(queue[r::W].copy())
        stats, Op = ops.nce_partial(qt, shard, 1 / T, F32, 3)
        packed = ops.nce_merge_packed(stats, Op)
        assert tuple(packed.shape) == (B, D + 4)
        s1, O1 = ops.nce_merge(stats, Op)
        assert torch.allclose(packed[:, :D], O1[0], rtol=1e-6, atol=1e-7)
        assert torch.allclose(packed[:, D:D + 3].t(), s1[:, 0], rtol=1e-6, atol=1e-7)
        recs.append(packed)
    rows, dq, pim, mx, loss, acc = ops.nce_combine_packed(torch.stack(recs), qt, kt, 1 / T, False, 1.0 / B)
    assert rel(npy(rows), rows_o) < TOL and rel(npy(dq), dq_o) < TOL
    assert abs(loss.item() - loss_o) < TOL * abs(loss_o)
    assert acc.item() == pytest.approx(pim_o.mean() * 100) and np.array_equal(pim.cpu().numpy().astype(bool), pim_o)


def test_attention_row_subset_and_strided_enqueue(ops):
    """forward_rows == the same rows of the full attention; strided enqueue == enqueue of the full list."""
    from moma_b200 import Attention
    torch.manual_seed(1)
    att = Attention(128, num_heads=4, qkv_bias=True).cuda()
    x = torch.randn(200, 128, device="cuda")
    with torch.no_grad():
        full = att(x)
    for start, stride in ((0, 1), (3, 4), (1, 8), (7, 8)):
        count = (200 - start + stride - 1) // stride
        sub = att.forward_rows(x, start, stride, count)
        assert torch.allclose(sub, full[start::stride], rtol=1e-5, atol=1e-6), (start, stride)
        # the K-sharded variant: per-rank projections, gathered (here: two "ranks" holding x[:100] / x[100:])
        parts = {}

        def fake_gather(qkv_local, parts=parts):
            parts.setdefault("other", att.qkv(x[100:]).detach())
            return torch.cat([qkv_local, parts["other"]], dim=0)
        sub2 = att.forward_rows_gathered(x[:100].contiguous(), fake_gather, start, stride, count)
        assert torch.allclose(sub2, full[start::stride], rtol=1e-5, atol=1e-6), (start, stride)
    rng = np.random.default_rng(5)
    K, D, n, W = 64, 32, 16, 4
    full_q = rng.standard_normal((K, D)).astype(np.float32)
    keys = rng.standard_normal((n, D)).astype(np.float32)
    for rank in range(W):
        index = 58                                    # wraps
        want = full_q.copy(); O.update_memory(want, keys, index)
        shard = cu(full_q[rank::W].copy())
        start = (rank - index) % W
        ops.enqueue(cu(keys[start::W].copy()), shard, None, K, index, rank=rank, world=W, key_start=start, key_stride=W)
        assert np.array_equal(npy(shard), want[rank::W])


# ----------------------------------------------------------------------------- projection-head Linear (3xTF32 GEMM)
@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K,relu,bias", [
    (256, 512, 512, True, True),      # embed_s layer 1 (C2)
    (256, 128, 512, False, True),     # embed_s layer 2: split-K
    (64, 128, 2048, False, True),     # long K, many splits
    (7, 33, 19, True, True),          # ragged: scalar copies, partial tiles
    (1, 128, 512, False, False),      # one row (B == 1), no bias
    (130, 36, 100, True, False),      # K % 32 != 0, M % 32 != 0
])
def test_linear_matches_fp64_reference(M, N, K, relu, bias):
    """moma_linear_fwd / moma_linear_bwd against torch in float64 (the reference's nn.Linear + nn.ReLU,
    criterion_moco_att.py:254-305).  Tolerance 2e-6 relative: the 3xTF32 product keeps fp32-level accuracy."""
    import torch
    from moma_b200 import ops
    dev = torch.device("cuda")
    torch.manual_seed(M * 7 + N * 3 + K)
    x = torch.randn(M, K, device=dev, requires_grad=True)
    w = (torch.randn(N, K, device=dev) / K ** 0.5).requires_grad_()
    b = torch.randn(N, device=dev, requires_grad=True) if bias else None
    gy = torch.randn(M, N, device=dev)
    for rep in range(2):                                   # twice: the split-K ticket counters must reset themselves
        y = ops.linear(x, w, b, relu=relu)
        grads = torch.autograd.grad(y, (x, w) + ((b,) if bias else ()), gy)
    xd, wd = x.detach().double().requires_grad_(), w.detach().double().requires_grad_()
    bd = b.detach().double().requires_grad_() if bias else None
    yd = torch.nn.functional.linear(xd, wd, bd)
    if relu:
        # same ReLU mask as the kernel's own output (an element within rounding of 0 may legitimately differ)
        yd = yd * (y.detach() > 0).double()
    gd = torch.autograd.grad(yd, (xd, wd) + ((bd,) if bias else ()), gy.double())

    def rel(a, ref):
        return float((a.double() - ref).norm() / ref.norm().clamp_min(1e-30))
    assert rel(y.detach(), yd.detach()) < 2e-6
    for got, want in zip(grads, gd):
        assert rel(got, want) < 2e-6
    # and against the numpy oracle (oracle/moma_oracle.py: linear / linear_backward) in float64
    xo, wo = x.detach().double().cpu().numpy(), w.detach().double().cpu().numpy()
    bo = b.detach().double().cpu().numpy() if bias else 0.0
    yo = O.linear(xo, wo, bo)
    yo = np.maximum(yo, 0) if relu else yo
    assert np.linalg.norm(y.detach().cpu().numpy() - yo) < 2e-6 * np.linalg.norm(yo)
    ox, ow, ob = O.linear_backward(xo, wo, y.detach().double().cpu().numpy(), gy.double().cpu().numpy(), relu=relu)
    for got, want in zip(grads, (ox, ow, ob)):
        assert np.linalg.norm(got.double().cpu().numpy() - want) < 2e-6 * max(np.linalg.norm(want), 1e-30)


@pytest.mark.gpu
def test_linear_is_deterministic_and_head_uses_it():
    """Split-K reduces in split order: bit-identical results run to run; the `mlp` head routes its Linears here."""
    import torch
    from argparse import Namespace
    from moma_b200 import CMO, _lib
    dev = torch.device("cuda")
    torch.manual_seed(5)
    crit = CMO(Namespace(head="mlp", attn="self", s_dim=512, t_dim=2048, feat_dim=128)).to(dev)
    x = torch.randn(64, 512, device=dev, requires_grad=True)
    lib = _lib.load()
    lib.moma_debug_launch_count(1)
    y1 = crit.embed_s(x)
    g1 = torch.autograd.grad(y1.square().sum() + y1[:, 0].sum(), x)[0]
    n = int(lib.moma_debug_launch_count(0))
    assert n >= 7          # 2 linear fwd + l2norm fwd, l2norm bwd, 2 x (dX, dW, db) backward launches
    y2 = crit.embed_s(x)
    g2 = torch.autograd.grad(y2.square().sum() + y2[:, 0].sum(), x)[0]
    assert torch.equal(y1, y2) and torch.equal(g1, g2)
    ref = torch.nn.Sequential(*list(crit.embed_s))        # plain torch modules over the same parameters
    yr = ref(x.detach())
    assert float((y1 - yr).norm() / yr.norm()) < 2e-6


@pytest.mark.gpu
def test_deferred_enqueue_equals_inline():
    """forward(..., defer_enqueue=True) + enqueue(all_k) leaves the same loss, gradient, queue and pointer as the
    reference-ordered forward (mem_moco.py:77-100): the step's loss never reads the keys it enqueues."""
    import torch
    from moma_b200 import MoCo
    torch.manual_seed(3)
    a = MoCo(128, 1024, 0.15).cuda()
    torch.manual_seed(3)
    b = MoCo(128, 1024, 0.15).cuda()
    a.index = b.index = 1000                      # wraps
    ce = torch.nn.CrossEntropyLoss()
    for _ in range(2):
        q1 = torch.randn(48, 128, device="cuda", requires_grad=True)
        q2 = q1.detach().clone().requires_grad_()
        k = torch.randn(48, 128, device="cuda")
        all_k = torch.randn(96, 128, device="cuda")
        la, lab = a(q1, k, all_k)
        lb, lbb = b(q2, k, defer_enqueue=True)
        ce(la, lab).backward()
        ce(lb, lbb).backward()
        b.enqueue(all_k)
        assert torch.equal(q1.grad, q2.grad) and torch.equal(a.memory, b.memory) and a.index == b.index


@pytest.mark.gpu
def test_attention_emits_the_bf16_query_operand():
    """In bf16 mode the attention's output projection also writes y rounded to bf16 (moma_attn_fwd's y_bf16): it must be
    bit-identical to the separate cast, be picked up by the InfoNCE pass, and be ignored once y was modified in place."""
    import torch
    import moma_b200
    from moma_b200 import Attention, MoCo, ops
    if not ops.bf16_supported(128):
        pytest.skip("no tcgen05 path on this device")
    moma_b200.set_precision("bf16")
    try:
        torch.manual_seed(11)
        att = Attention(128, num_heads=4, qkv_bias=True).cuda()
        x = torch.randn(96, 128, device="cuda", requires_grad=True)
        y = att(x)
        y16, ver = y._moma_bf16
        assert ver == y._version and torch.equal(y16, y.detach().to(torch.bfloat16))
        lib = ops._lib.load()
        lib.moma_debug_launch_count(1)
        q_op = ops.nce_operands(y, torch.randn(96, 128, device="cuda"), "bf16")[0]
        assert q_op is y16 and int(lib.moma_debug_launch_count(0)) == 0           # no cast kernel
        # the loss / gradient through the cached operand equal the ones through an explicit cast
        torch.manual_seed(12)
        a = MoCo(128, 1024, 0.15).cuda()
        torch.manual_seed(12)
        b = MoCo(128, 1024, 0.15).cuda()
        k = torch.randn(96, 128, device="cuda")
        ce = torch.nn.CrossEntropyLoss()
        la, lab = a(y, k)
        ga = torch.autograd.grad(ce(la, lab), x, retain_graph=True)[0]
        y2 = y * 1.0                                                               # same values, no cached operand
        assert getattr(y2, "_moma_bf16", None) is None
        lb, lbb = b(y2, k)
        gb = torch.autograd.grad(ce(lb, lbb), x)[0]
        assert torch.equal(ga, gb)
        with torch.no_grad():
            y.mul_(2.0)                                                            # in-place change: cache is stale
        assert ops.nce_operands(y, k, "bf16")[0] is not y16
    finally:
        moma_b200.set_precision("bf16")
