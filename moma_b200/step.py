"""The MoMA criterion step (level L1 of SURVEY 8d) as one callable object.

``CriterionStep`` performs exactly the module calls of the moma branch of the reference loop
(helper/loops_moma.py:308-335 and the backward at :360) on backbone FEATURES -- the backbones are outside
the path -- in the reference's call order:

    momentum_update(student, teacher)            # :309   backbone EMA (student -> same-architecture twin)
    momentum_update(embed_s, embed_t)            # :310-312 (only when s_dim == t_dim, as the reference requires)
    k      = embed_t(feat_t)                     # contrast_trainer.py:121 (no grad)
    f_s    = atts_q(embed_s(feat_s))             # :323-326
    k2     = atts_k(k); all_k = atts_queue(gather(k))      # :327-329, contrast_trainer.py:124
    logits, labels = contrast(f_s, k2, all_k)    # :331
    loss   = CrossEntropyLoss()(logits, labels); accuracy(...)     # contrast_trainer.py:189-205
    loss.backward()                              # :360

``step()`` issues them sequentially through the public module API (what the unchanged loop does).
``step_overlapped()`` issues the SAME calls with the independent branches forked onto side streams, so that a
CUDA-graph capture (``moma_b200.graphed.GraphedStep``) records the step's real dependency DAG; ``queue_layout``
selects the replicated queue (the reference's layout) or the K-sharded one.  bench.py, the parity self-check
and tests/ all drive this one object.
"""
from __future__ import annotations

import os
from argparse import Namespace

import torch

from . import ops
from .contrast_trainer import ContrastTrainer
from .criterion_moco_att import CMO
from .mem_moco import build_mem

T_NCE, ALPHA, SEED = 0.15, 0.999, 12345


def resnet18_param_shapes(num_classes=4):
    """Parameter shapes of the reference ResNet-18 (models/resnet_imagenet.py; 62 tensors, 11,178,564 elements)
    in parameters() order -- the EMA pair is (student, same-architecture momentum twin) because the reference's
    momentum_update raises on heterogeneous pairs."""
    shapes = [(64, 3, 7, 7), (64,), (64,)]
    cin = 64
    for cout, stride in ((64, 1), (128, 2), (256, 2), (512, 2)):
        for blk in range(2):
            shapes += [(cout, cin if blk == 0 else cout, 3, 3), (cout,), (cout,), (cout, cout, 3, 3), (cout,), (cout,)]
            if blk == 0 and (stride != 1 or cin != cout):
                shapes += [(cout, cin, 1, 1), (cout,), (cout,)]
        cin = cout
    shapes += [(num_classes, 512), (num_classes,)]
    return shapes


class CriterionStep:
    """cfg: dict(B, s_dim, t_dim, D, K, H).  B is the PER-RANK batch.  ``sharded``: K-sharded queue (needs an
    initialised process group of size ``world`` > 1); otherwise every rank keeps the full queue and enqueues the
    all-gathered keys (the reference's layout, contrast_trainer.py:124 + mem_moco.py:97-99)."""

    def __init__(self, cfg, rank, world, device, sharded=None, ema_shapes=None, precision="bf16", seed=SEED):
        self.cfg, self.rank, self.world, self.dev = cfg, rank, world, device
        self.sharded = (world > 1) if sharded is None else bool(sharded)
        ops.set_precision(precision)
        torch.manual_seed(seed)                       # same seed on every rank -> identical init (reference :241-246)
        opt = Namespace(head="mlp", s_dim=cfg["s_dim"], t_dim=cfg["t_dim"], feat_dim=cfg["D"], attn="self", mem="MoCo",
                        nce_k=cfg["K"], nce_t=T_NCE, alpha=ALPHA, num_heads=cfg["H"], shard_queue=self.sharded)
        self.opt = opt
        self.contrast = build_mem(opt).to(device)
        self.crit = CMO(opt).to(device)
        self.trainer = ContrastTrainer
        shapes = resnet18_param_shapes() if ema_shapes is None else ema_shapes
        self.student = torch.nn.ParameterList([torch.nn.Parameter(torch.randn(*s)) for s in shapes]).to(device)
        self.teacher = torch.nn.ParameterList([torch.nn.Parameter(torch.randn(*s)) for s in shapes]).to(device)
        self.ema_elems = sum(p.numel() for p in self.student)
        self.ce = torch.nn.CrossEntropyLoss()
        self.head_ema = cfg["s_dim"] == cfg["t_dim"]
        self.params = [p for n, p in self.crit.named_parameters() if not n.startswith("embed_t")]
        torch.manual_seed(seed + 1 + rank)            # data differs per rank
        B = cfg["B"]
        self.feat_s = torch.randn(B, cfg["s_dim"], device=device, requires_grad=True)
        self.feat_t = torch.randn(B, cfg["t_dim"], device=device)
        self.host_s = torch.randn(B, cfg["s_dim"])
        self.host_t = torch.randn(B, cfg["t_dim"])
        if device.type == "cuda":
            self.host_s, self.host_t = self.host_s.pin_memory(), self.host_t.pin_memory()
        self.h2d_bytes = (self.host_s.numel() + self.host_t.numel()) * 4
        self.loss = self.acc = None
        self._unit = None
        self.last = {}                                # tensors of the latest step, for the parity self-check

    # ------------------------------------------------------------------ keys of the step
    def _queue_keys(self, k0):
        """(all_k, owned_k): the keys this step enqueues.  Replicated layout: atts_queue over the all-gathered
        keys, every rank writes all n rows.  Sharded layout: this rank only enqueues every W-th attended key,
        so it attends those rows only, and every rank projects only its own keys (the qkv projections are
        all-gathered instead of the raw keys)."""
        crit = self.crit
        if self.world > 1 and self.sharded:
            owned = crit.atts_queue.forward_rows_gathered(k0, self.trainer._global_gather,
                                                          *self.contrast.owned_rows(k0.shape[0] * self.world))
            return None, owned
        gathered = self.trainer._global_gather(k0) if self.world > 1 else k0
        return crit.atts_queue(gathered), None

    # ------------------------------------------------------------------ sequential (the reference loop's order)
    def step(self, feat_s=None, feat_t=None):
        crit, opt = self.crit, self.opt
        feat_s = self.feat_s if feat_s is None else feat_s
        feat_t = self.feat_t if feat_t is None else feat_t
        self.trainer.momentum_update(self.student, self.teacher, opt.alpha)
        if self.head_ema:
            self.trainer.momentum_update(crit.embed_s, crit.embed_t, opt.alpha)
        with torch.no_grad():
            k0 = crit.embed_t(feat_t)
        f_s = crit.embed_s(feat_s)
        f_s = crit.atts_q(f_s)
        k = crit.atts_k(k0)
        all_k, owned = self._queue_keys(k0)
        return self._loss_and_backward(f_s, k, all_k, feat_s, owned_k=owned)

    def _loss_and_backward(self, f_s, k, all_k, feat_s, owned_k=None, enqueue_stream=None, ema_stream=None):
        if enqueue_stream is not None:
            output = self.contrast(q=f_s, k=k, defer_enqueue=True)
            late = getattr(self, "_late_box", None)
            if late is not None:                        # the queue branch was launched from inside the forward (see below)
                self.contrast.after_query_gather = None
                all_k, owned_k = late.pop("keys")
                self._late_box = None
        else:
            output = self.contrast(q=f_s, k=k, owned_k=owned_k) if owned_k is not None else \
                self.contrast(q=f_s, k=k, all_k=all_k)
        losses, accs = self.trainer._compute_loss_accuracy(output[:-1], output[-1], self.ce)
        for p in self.params:
            p.grad = None
        feat_s.grad = None
        if ema_stream is not None:
            # The backbone EMA is pure HBM streaming (270 MB) and depends on nothing in the step.  It is forked HERE, behind
            # the InfoNCE pass -- the one kernel of the step that streams from HBM itself (the queue) -- so that it overlaps
            # the backward, whose kernels are latency-bound and L2-resident.  (Forked earlier it shares HBM with the
            # InfoNCE pass: 40 us instead of 23 us for that kernel at K = 65536.)
            ema_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(ema_stream):
                self.trainer.momentum_update(self.student, self.teacher, self.opt.alpha)
        if enqueue_stream is not None:
            # the queue update only has to follow the InfoNCE pass that reads the old queue: it runs on the branch
            # that produced the new keys, concurrently with the backward
            enqueue_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(enqueue_stream), torch.no_grad():
                if owned_k is not None:
                    self.contrast.enqueue(owned_k=owned_k)
                else:
                    self.contrast.enqueue(all_k)
        # loss.backward() as at helper/loops_moma.py:360; the unit upstream gradient is a preallocated device scalar
        # (autograd would otherwise launch a fill kernel for it every step)
        if self._unit is None:
            self._unit = torch.ones((), dtype=torch.float32, device=self.dev)
        losses[0].backward(gradient=self._unit)
        self.loss, self.acc = losses[0], accs[0]
        self.last = {"q": f_s, "k": k, "all_k": all_k, "owned_k": owned_k}
        return losses[0]

    # ------------------------------------------------------------------ same calls, real dependency DAG
    def step_overlapped(self):
        """Same module calls, with the independent branches forked onto side streams so the captured graph exposes
        the step's real dependencies: the backbone EMA touches nothing else in the step, and the teacher branch
        (embed_t -> atts_k / atts_queue) only meets the student branch (embed_s -> atts_q) at the InfoNCE pass."""
        crit, opt = self.crit, self.opt
        main = torch.cuda.current_stream()
        if not hasattr(self, "_side"):
            # EMA and the queue-attention branch at normal priority; the teacher branch at the priority of the student
            # chain (the launching stream): whichever of the two is longer is the step's critical path (the teacher's
            # 2048-wide head at C3, the student's at C2)
            hi = -1 if os.environ.get("MOMA_B200_TEACHER_PRIORITY", "high") == "high" else 0
            self._side = [torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev, priority=hi), torch.cuda.Stream(self.dev)]
        s_ema, s_t, s_u = self._side
        s_t.wait_stream(main)
        with torch.cuda.stream(s_t):
            if self.head_ema:
                self.trainer.momentum_update(crit.embed_s, crit.embed_t, opt.alpha)
            with torch.no_grad():
                k0 = crit.embed_t(self.feat_t)
            late_queue = self.world > 1 and self.sharded and os.environ.get("MOMA_B200_QUEUE_BRANCH", "sequenced") == "sequenced"
            all_k = owned = None
            if not late_queue:
                s_u.wait_stream(s_t)
                with torch.cuda.stream(s_u):
                    all_k, owned = self._queue_keys(k0)
            k = crit.atts_k(k0)
        if late_queue:
            # Sharded queue: the queue-attention branch starts with a LARGE exchange (the projections of every rank's keys).
            # Measured at 8 GPUs: while it is in flight, the small query all-gather on the critical path takes 70 us
            # instead of 8.  So the branch is launched from a hook that fires right after the query all-gather has been
            # enqueued, and waits for it: the exchanges of a step run one after the other, the branch still overlaps the
            # InfoNCE pass and the backward.
            box = {}

            def launch_queue_branch():
                s_u.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s_u), torch.no_grad():
                    box["keys"] = self._queue_keys(k0)
            self.contrast.after_query_gather = launch_queue_branch
            self._late_box = box
        f_s = crit.embed_s(self.feat_s)
        f_s = crit.atts_q(f_s)
        ema_late = os.environ.get("MOMA_B200_EMA_FORK", "early") != "early"       # A/B switch (see _loss_and_backward)
        if not ema_late:
            s_ema.wait_stream(s_t)
            with torch.cuda.stream(s_ema):
                self.trainer.momentum_update(self.student, self.teacher, opt.alpha)
        # The loss of this step needs q, the local positive keys and the OLD queue -- not the keys enqueued for later
        # steps: only the teacher branch (s_t) joins here, the queue-attention branch (s_u) joins after the backward.
        main.wait_stream(s_t)
        loss = self._loss_and_backward(f_s, k, all_k, self.feat_s, owned_k=owned, enqueue_stream=s_u, ema_stream=s_ema if ema_late else None)
        main.wait_stream(s_u)
        main.wait_stream(s_ema)
        return loss

    # ------------------------------------------------------------------ views for checks
    def full_queue(self) -> torch.Tensor:
        """The whole [K, D] fp32 queue in the reference's row order (collective for the sharded layout)."""
        c = self.contrast
        return c.gather_full() if getattr(c, "is_sharded", False) else c.memory
