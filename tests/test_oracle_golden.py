"""Pin the CPU oracle against golden vectors produced by the unmodified
reference (tests/golden/make_golden.py).  Runs on CPU."""
import numpy as np
import pytest

from oracle import moma_oracle as O


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def test_kat1_moco_scalars(golden):
    g = golden("kat_moco")
    # SURVEY 8c KAT1 known answers (recorded independently during the survey)
    assert abs(float(g["kat1_loss"]) - 8.53796864) < 2e-6
    assert abs(float(g["kat1_gradnorm"]) - 1.18056488) < 2e-6
    assert np.allclose(g["kat1_logits_head"][0, :3], [-0.0396851, -0.4447779, -1.1489086], atol=1e-6)
    assert int(g["kat1_index"]) == 32
    assert (g["kat1_labels"] == 0).all() and g["kat1_labels"].dtype == np.int64
    # enqueue wrote k verbatim into rows 0..31, rows 32.. untouched
    assert np.array_equal(g["kat1_mem_after_rows"][:32], g["kat1_k"])
    assert np.array_equal(g["kat1_mem_after_rows"][32:], g["kat1_mem_rows"][32:])
    # positive logit column restated
    pos = O.compute_logit_qk(g["kat1_q"], g["kat1_k"], 0.15)
    assert np.allclose(pos[:4], g["kat1_logits_head"][:, 0], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("dt,tol", [(np.float32, 2e-6), (np.float64, 2e-6)])
def test_small_moco_forward(golden, dt, tol):
    g = golden("kat_moco")
    mem = g["s_mem0"].astype(dt)
    q, k, allk = g["s_q"].astype(dt), g["s_k"].astype(dt), g["s_allk"].astype(dt)
    logits, labels, idx = O.moco_forward(mem, 56, q, k, allk, T=0.07)
    assert logits.shape == g["s_logits"].shape
    assert rel(logits, g["s_logits"]) < tol
    assert np.array_equal(labels, g["s_labels"]) and labels.dtype == np.int64
    assert idx == int(g["s_index"]) == (56 + 24) % 64
    assert np.array_equal(mem.astype(np.float32), g["s_mem1"])        # rows copied verbatim
    loss, rows = O.cross_entropy_zero_label(logits)
    assert abs(loss - float(g["s_loss"])) < 1e-5 * abs(float(g["s_loss"]))
    loss2, rows2, dq, pos_is_max = O.nce_loss_and_grad(q, k, g["s_mem0"].astype(dt), 0.07)
    assert abs(loss2 - float(g["s_loss"])) < 1e-5 * abs(float(g["s_loss"]))
    assert rel(dq, g["s_dq"]) < 1e-5
    assert O.accuracy_top1(logits, labels) == pytest.approx(float(g["s_acc"][0]))
    assert pos_is_max.mean() * 100 == pytest.approx(float(g["s_acc"][0]))


def test_split_merge_matches_closed_form(golden):
    g = golden("kat_moco")
    q, k, mem = (g[n].astype(np.float64) for n in ("s_q", "s_k", "s_mem0"))
    loss, rows, dq, pim = O.nce_loss_and_grad(q, k, mem, 0.07)
    # cyclic shards (world 4) and contiguous splits both merge to the same answer
    for parts in ([mem[r::4] for r in range(4)], [mem[:10], mem[10:33], mem[33:]]):
        partials = [O.nce_partial(q, p, 0.07) for p in parts]
        rows2, dq_unit, pim2 = O.nce_merge(partials, q, k, 0.07)
        assert np.allclose(rows2, rows, rtol=1e-12, atol=1e-12)
        assert np.allclose(dq_unit / q.shape[0], dq, rtol=1e-10, atol=1e-13)
        assert np.array_equal(pim, pim2)


def test_b1_squeeze(golden):
    g = golden("kat_moco")
    mem = g["b1_mem0"].copy()
    logits, labels, idx = O.moco_forward(mem, 0, g["b1_q"], g["b1_k"], None, T=0.2)
    assert logits.shape == g["b1_logits"].shape == (33,)          # KAT4: 1-D when B == 1
    assert rel(logits, g["b1_logits"]) < 2e-6
    assert labels.shape == (1,)


def test_pointer_and_ids(golden):
    g = golden("kat_pointer")
    mem = np.zeros((10, 4), np.float32)
    k = (np.arange(16, dtype=np.float32).reshape(4, 4) + 1)
    O.update_memory(mem, k, 8)
    assert np.array_equal(mem, g["kat2_mem"])                      # KAT2: rows 8,9,0,1
    assert O.update_pointer(8, 4, 10) == int(g["kat2_index"]) == 2
    for K, n in ((4096, 96), (10, 4), (65536, 512), (131072, 1024), (7, 7), (12, 5)):
        idx = 0
        for s, want in enumerate(g[f"ptr_K{K}_n{n}_index"]):
            ids = O.enqueue_ids(n, idx, K)
            assert ids.dtype == np.int64
            assert np.array_equal(ids, O.enqueue_ids_c(n, idx, K))
            assert ids[0] == g[f"ptr_K{K}_n{n}_first"][s] and ids[-1] == g[f"ptr_K{K}_n{n}_last"][s]
            idx = O.update_pointer(idx, n, K)
            assert idx == want
    # KAT3: K=4096, n=96 -> index 4032 / 32 / 128 after 42 / 43 / 44 steps
    seq = g["ptr_K4096_n96_index"]
    assert (seq[41], seq[42], seq[43]) == (4032, 32, 128)
    assert np.array_equal(O.enqueue_ids(96, 4032, 4096), g["ids_wrap"])


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_attention(golden, tag):
    g = golden("kat_attention")
    H = int(g[f"{tag}_H"])
    bq = g[f"{tag}_bqkv"] if f"{tag}_bqkv" in g.files else None
    args = (g[f"{tag}_x"], g[f"{tag}_wqkv"], bq, g[f"{tag}_wproj"], g[f"{tag}_bproj"], H)
    y = O.attention_forward(*args)
    assert rel(y, g[f"{tag}_y"]) < 5e-6
    gr = O.attention_backward(*args, g[f"{tag}_dy"])
    assert rel(gr["dx"], g[f"{tag}_dx"]) < 2e-5
    assert rel(gr["d_wqkv"], g[f"{tag}_dwqkv"]) < 2e-5
    assert rel(gr["d_wproj"], g[f"{tag}_dwproj"]) < 2e-5
    assert rel(gr["d_bproj"], g[f"{tag}_dbproj"]) < 2e-5
    if bq is not None:
        assert rel(gr["d_bqkv"], g[f"{tag}_dbqkv"]) < 2e-5
    # KAT5: scale = hd^-0.5
    assert float(g[f"{tag}_scale"]) == (g[f"{tag}_x"].shape[1] // H) ** -0.5
    # float64 restatement agrees too
    y64 = O.attention_forward(*[a.astype(np.float64) if isinstance(a, np.ndarray) else a for a in args])
    assert rel(y64, g[f"{tag}_y"]) < 5e-6


def test_normalize(golden):
    g = golden("kat_normalize")
    y = O.normalize(g["x"])
    assert np.allclose(y, g["y"], rtol=2e-6, atol=1e-7)
    assert (y[3] == 0).all()                                   # zero row stays zero (eps clamp)
    dx = O.normalize_backward(g["x"].astype(np.float64), g["g"].astype(np.float64))
    live = [i for i in range(12) if i not in (3, 5)]
    assert rel(dx[live], g["dx"][live]) < 1e-5
    assert rel(dx[5], g["dx"][5]) < 1e-5                        # below-eps row: g / eps
    assert rel(dx[3], g["dx"][3]) < 1e-5


def test_ema_bit_exact(golden):
    g = golden("kat_ema")
    n = int(g["n"])
    src = [g[f"src{i}"].copy() for i in range(n)]
    ema = [g[f"ema{i}_0"].copy() for i in range(n)]
    for step in (1, 2, 3):
        O.momentum_update(ema, src, 0.999)
        for i in range(n):
            assert np.array_equal(ema[i], g[f"ema{i}_{step}"]), (i, step)   # bit-exact fp32
    emb = [g[f"emb{i}_0"].copy() for i in range(n)]
    O.momentum_update(emb, src, 0.5)
    for i in range(n):
        assert np.array_equal(emb[i], g[f"emb{i}_1"])
    assert bool(g["mismatch_raises"])
    with pytest.raises(RuntimeError):
        O.momentum_update([np.zeros((4, 3), np.float32)], [np.zeros((3, 4), np.float32)], 0.9)


def test_dual_queue_variants(golden):
    g = golden("kat_dual")
    q, k, kt = g["st_q"], g["st_k"], g["st_kt"]
    ms, mt = g["st_ms0"].copy(), g["st_mt0"].copy()
    lss = O.compute_logit(q, k, ms, 0.1); lst = O.compute_logit(q, kt, mt, 0.1)
    assert rel(lss, g["st_lss"]) < 2e-6 and rel(lst, g["st_lst"]) < 2e-6
    O.update_memory(ms, k, 22); O.update_memory(mt, kt, 22)
    assert np.array_equal(ms, g["st_ms1"]) and np.array_equal(mt, g["st_mt1"])
    assert O.update_pointer(22, 4, 24) == int(g["st_index"])
    ms, mt, qt = g["sstt_ms0"], g["sstt_mt0"], g["sstt_qt"]
    assert rel(O.compute_logit(qt, k, ms, 0.1), g["sstt_lts"]) < 2e-6
    assert rel(O.compute_logit(qt, kt, mt, 0.1), g["sstt_ltt"]) < 2e-6


def test_gloo_two_rank_golden(golden):
    g = golden("kat_gloo")
    all_k = O.global_gather([g["r0_k"], g["r1_k"]])
    for r in (0, 1):
        assert np.array_equal(g[f"r{r}_all_k"], all_k)                  # KAT7
        mem = g[f"r{r}_mem0"].copy()
        logits, labels, idx = O.moco_forward(mem, 28, g[f"r{r}_q"], g[f"r{r}_k"], all_k, T=0.15)
        assert rel(logits, g[f"r{r}_logits"]) < 2e-6
        assert np.array_equal(mem, g[f"r{r}_mem1"]) and idx == int(g[f"r{r}_index"]) == 4
    assert np.array_equal(g["r0_mem1"], g["r1_mem1"])                   # replicated queue identical
    # cyclic shards of the updated queue reassemble to the replicated queue
    mem1 = g["r0_mem1"]
    ids = np.arange(mem1.shape[0])
    owner, slot = O.shard_owner_slot(ids, 2)
    shards = [mem1[owner == r] for r in (0, 1)]
    re = np.empty_like(mem1)
    for r in (0, 1):
        re[ids[owner == r]] = shards[r][slot[owner == r]]
    assert np.array_equal(re, mem1)


def test_criterion_step_golden(golden):
    """The whole moma branch (helper/loops_moma.py:308-335) restated with the
    oracle against the reference run, 3 steps (enqueue wraps at K=40, n=16)."""
    g = golden("criterion_step")
    sd = {k[4:]: g[k].copy() for k in g.files if k.startswith("sd0_")}
    mem = g["mem0"].copy(); index = 0
    H, T, alpha = 4, 0.15, 0.999
    for st in range(3):
        fs, ft = g[f"st{st}_feat_s"], g[f"st{st}_feat_t"]
        # EMA of the head (loops_moma.py:310-312)
        names = ["1.weight", "1.bias", "3.weight", "3.bias"]
        O.momentum_update([sd["embed_t." + n] for n in names], [sd["embed_s." + n] for n in names], alpha)
        k = O.embed_mlp(ft, *[sd["embed_t." + n] for n in names])
        assert rel(k, g[f"st{st}_k"]) < 5e-6
        f_s = O.embed_mlp(fs, *[sd["embed_s." + n] for n in names])
        att = lambda p, x: O.attention_forward(x, sd[p + ".qkv.weight"], sd[p + ".qkv.bias"],
                                               sd[p + ".proj.weight"], sd[p + ".proj.bias"], H)
        f_s = att("atts_q", f_s); k2 = att("atts_k", k); allk2 = att("atts_queue", k)
        assert rel(f_s, g[f"st{st}_f_s"]) < 2e-5 and rel(allk2, g[f"st{st}_allk2"]) < 2e-5
        loss, rows, dq, pim = O.nce_loss_and_grad(f_s, k2, mem, T)
        assert abs(loss - float(g[f"st{st}_loss"])) < 2e-5 * abs(float(g[f"st{st}_loss"]))
        assert pim.mean() * 100 == pytest.approx(float(g[f"st{st}_acc"][0]))
        # enqueue the reference's own post-attention keys: queue must then match bit-exactly
        O.update_memory(mem, g[f"st{st}_allk2"], index)
        index = O.update_pointer(index, 16, 40)
        assert np.array_equal(mem, g[f"st{st}_mem"]) and index == int(g[f"st{st}_index"])
        # KAT6: only atts_q and embed_s receive gradients
        for n in [k[len(f"st{st}_hasgrad_"):] for k in g.files if k.startswith(f"st{st}_hasgrad_")]:
            want = n.startswith("atts_q.") or n.startswith("embed_s.")
            assert bool(g[f"st{st}_hasgrad_{n}"]) == want, n
        # next step uses the reference's post-SGD parameters
        for n in list(sd):
            sd[n] = g[f"st{st}_sd_{n}"].copy()
    assert index == 48 % 40


def test_linear_backward_matches_torch_autograd():
    """oracle.linear / linear_backward (the heads' nn.Linear + nn.ReLU) against torch autograd in float64."""
    import torch
    rng = np.random.default_rng(3)
    x, w, b = rng.standard_normal((9, 7)), rng.standard_normal((5, 7)), rng.standard_normal(5)
    gy = rng.standard_normal((9, 5))
    for relu in (False, True):
        xt, wt, bt = (torch.tensor(a, requires_grad=True) for a in (x, w, b))
        yt = torch.nn.functional.linear(xt, wt, bt)
        yt = torch.relu(yt) if relu else yt
        gx, gw, gb = torch.autograd.grad(yt, (xt, wt, bt), torch.tensor(gy))
        y = O.linear(x, w, b)
        y = np.maximum(y, 0) if relu else y
        assert np.allclose(y, yt.detach().numpy(), rtol=1e-13, atol=1e-13)
        ox, ow, ob = O.linear_backward(x, w, y, gy, relu=relu)
        for got, want in ((ox, gx), (ow, gw), (ob, gb)):
            assert np.allclose(got, want.numpy(), rtol=1e-12, atol=1e-12)
