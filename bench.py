#!/usr/bin/env python
"""bench.py -- MoMA criterion step throughput (samples/s) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C2|C3|C5]

A "step" is one pass of the hot path over one synthetic batch (level L1):
backbone EMA (student -> momentum twin) + projection heads + 3x attention + fused InfoNCE
logits/CE forward AND backward (to d loss / d feat_s and the criterion parameters) + key
all-gather (N > 1) + ring enqueue.  Backbones are outside the path: inputs are their features.

Default workload = BASELINE.json configs[1] (C2): ResNet-50 teacher -> ResNet-18 student,
batch 256 per GPU, D = 128, K = 16384, bf16 InfoNCE operands.  With N > 1 the per-GPU batch is
fixed (weak scaling) and the queue is sharded by K across ranks.

Prints ONE JSON line (rank 0).  `value` is device-resident throughput; `e2e` goes through the
public module API with HOST (pinned) inputs and a host read of the loss inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

CONFIGS = {
    # name: per-GPU batch, s_dim, t_dim, D, K, heads, description
    "C1": dict(B=32, s_dim=512, t_dim=2048, D=128, K=4096, H=4,
               desc="C1 R18<-R50 B32 D128 K4096 (the reference's CPU-runnable case)"),
    "C2": dict(B=256, s_dim=512, t_dim=2048, D=128, K=16384, H=4,
               desc="C2 MoMA criterion step, ResNet-50 teacher -> ResNet-18 student, B256/GPU D128 K16384 bf16"),
    "C3": dict(B=512, s_dim=512, t_dim=2048, D=128, K=65536, H=8,
               desc="C3 large memory bank K65536, 8 heads, B512/GPU D128"),
    "C5": dict(B=1024, s_dim=384, t_dim=768, D=256, K=131072, H=4,
               desc="C5 ViT-S<-ViT-B features, B1024/GPU D256 K131072"),
}
T_NCE, ALPHA, SEED = 0.15, 0.999, 12345


def resnet18_param_shapes(num_classes=4):
    """Parameter shapes of the reference ResNet-18 (models/resnet_imagenet.py; 62 tensors,
    11,178,564 elements) in parameters() order -- the EMA pair is (student, same-architecture
    momentum twin) because the reference's momentum_update raises on heterogeneous pairs."""
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


# ----------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.stop_flag, self.max_mhz = [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksEventReasonHwPowerBrakeSlowdown: "hw_power_brake"} if hasattr(
            nv, "nvmlClocksEventReasonHwSlowdown") else {}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ----------------------------------------------------------------------------- our arm
class CriterionStep:
    """The criterion step through the public (reference-shaped) module API of moma_b200."""

    def __init__(self, cfg, rank, world, device):
        from argparse import Namespace

        import moma_b200
        from moma_b200 import CMO, ContrastTrainer, build_mem
        self.cfg, self.rank, self.world, self.dev = cfg, rank, world, device
        moma_b200.set_precision("bf16")
        torch.manual_seed(SEED)                       # same seed on every rank -> identical init (reference :241-246)
        opt = Namespace(head="mlp", s_dim=cfg["s_dim"], t_dim=cfg["t_dim"], feat_dim=cfg["D"], attn="self", mem="MoCo",
                        nce_k=cfg["K"], nce_t=T_NCE, alpha=ALPHA, num_heads=cfg["H"], shard_queue=world > 1)
        self.opt = opt
        self.contrast = build_mem(opt).to(device)
        self.crit = CMO(opt).to(device)
        self.trainer = ContrastTrainer
        shapes = resnet18_param_shapes()
        self.student = torch.nn.ParameterList([torch.nn.Parameter(torch.randn(*s)) for s in shapes]).to(device)
        self.teacher = torch.nn.ParameterList([torch.nn.Parameter(torch.randn(*s)) for s in shapes]).to(device)
        self.ema_elems = sum(p.numel() for p in self.student)
        self.ce = torch.nn.CrossEntropyLoss()
        self.head_ema = cfg["s_dim"] == cfg["t_dim"]
        self.params = [p for n, p in self.crit.named_parameters() if not n.startswith("embed_t")]
        torch.manual_seed(SEED + 1 + rank)            # data differs per rank
        B = cfg["B"]
        self.feat_s = torch.randn(B, cfg["s_dim"], device=device, requires_grad=True)
        self.feat_t = torch.randn(B, cfg["t_dim"], device=device)
        self.host_s = torch.randn(B, cfg["s_dim"]).pin_memory()
        self.host_t = torch.randn(B, cfg["t_dim"]).pin_memory()
        self.h2d_bytes = (self.host_s.numel() + self.host_t.numel()) * 4
        self.loss = None

    def step(self, feat_s=None, feat_t=None):
        """helper/loops_moma.py:308-335 + :360 (backward) on features, in the reference's call order."""
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
        if self.world > 1:
            # K-sharded queue: this rank only enqueues every W-th attended key -> attend those rows only, and every
            # rank projects only its own keys (the qkv projections are all-gathered instead of the raw keys)
            owned = crit.atts_queue.forward_rows_gathered(k0, self.trainer._global_gather,
                                                          *self.contrast.owned_rows(k0.shape[0] * self.world))
            return self._loss_and_backward(f_s, k, None, feat_s, owned_k=owned)
        all_k = crit.atts_queue(k0)
        return self._loss_and_backward(f_s, k, all_k, feat_s)

    def _loss_and_backward(self, f_s, k, all_k, feat_s, owned_k=None, enqueue_stream=None):
        if enqueue_stream is not None:
            output = self.contrast(q=f_s, k=k, defer_enqueue=True)
        else:
            output = self.contrast(q=f_s, k=k, owned_k=owned_k) if owned_k is not None else \
                self.contrast(q=f_s, k=k, all_k=all_k)
        losses, accs = self.trainer._compute_loss_accuracy(output[:-1], output[-1], self.ce)
        for p in self.params:
            p.grad = None
        feat_s.grad = None
        if enqueue_stream is not None:
            # the queue update only has to follow the InfoNCE pass that reads the old queue: it runs on the branch
            # that produced the new keys, concurrently with the backward
            enqueue_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(enqueue_stream), torch.no_grad():
                if owned_k is not None:
                    self.contrast.enqueue(owned_k=owned_k)
                else:
                    self.contrast.enqueue(all_k)
        losses[0].backward()
        self.loss = losses[0]
        return losses[0]

    def step_overlapped(self):
        """Same module calls, with the independent branches forked onto side streams so the captured
        graph exposes the step's real dependencies: the backbone EMA touches nothing else in the step,
        and the teacher branch (embed_t -> atts_k / atts_queue) only meets the student branch
        (embed_s -> atts_q) at the InfoNCE pass."""
        crit, opt = self.crit, self.opt
        main = torch.cuda.current_stream()
        if not hasattr(self, "_side"):
            self._side = [torch.cuda.Stream(self.dev) for _ in range(3)]
        s_ema, s_t, s_u = self._side
        s_t.wait_stream(main)
        with torch.cuda.stream(s_t):
            if self.head_ema:
                self.trainer.momentum_update(crit.embed_s, crit.embed_t, opt.alpha)
            with torch.no_grad():
                k0 = crit.embed_t(self.feat_t)
            s_u.wait_stream(s_t)
            owned = all_k = None
            with torch.cuda.stream(s_u):
                if self.world > 1:
                    owned = crit.atts_queue.forward_rows_gathered(k0, self.trainer._global_gather,
                                                                  *self.contrast.owned_rows(k0.shape[0] * self.world))
                else:
                    all_k = crit.atts_queue(k0)
            k = crit.atts_k(k0)
        f_s = crit.embed_s(self.feat_s)
        f_s = crit.atts_q(f_s)
        # The backbone EMA (bandwidth-bound, 270 MB of traffic) is forked behind the teacher branch, which finishes well
        # before the student chain: it then overlaps the InfoNCE pass and the backward (all latency-bound, L2-resident)
        # instead of the projection heads, the only other kernels of the step that miss in L2 (cold weights).
        # Measured per step: forked at the start 178 us, after the heads 160.3, after the attention 163.7, here 159.6.
        s_ema.wait_stream(s_t)
        with torch.cuda.stream(s_ema):
            self.trainer.momentum_update(self.student, self.teacher, opt.alpha)
        # The loss of this step needs q, the local positive keys and the OLD queue -- not the keys enqueued for later
        # steps: only the teacher branch (s_t) joins here, the queue-attention branch (s_u) joins after the backward.
        main.wait_stream(s_t)
        loss = self._loss_and_backward(f_s, k, all_k, self.feat_s, owned_k=owned, enqueue_stream=s_u)
        main.wait_stream(s_u)
        main.wait_stream(s_ema)
        return loss

    def step_e2e(self):
        fs = self.host_s.to(self.dev, non_blocking=True).requires_grad_()
        ft = self.host_t.to(self.dev, non_blocking=True)
        return float(self.step(fs, ft).item())          # D2H read of the loss inside the timed region


def run_ours(args, cfg, rank, world, local_rank):
    import torch.distributed as dist

    import moma_b200
    from moma_b200 import _lib, ops
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    lib = _lib.load()
    # everything runs on one side stream from the first call on (graph capture needs a non-default
    # stream, and autograd binds gradient accumulation to the stream a leaf was first used on)
    # High priority: the student chain (this stream) is the critical path of the step; when its small CTAs compete with
    # the teacher branch's library GEMM for SM slots they go first (measured: 163.1 -> 160.6 us per step).
    torch.cuda.set_stream(torch.cuda.Stream(device, priority=-1))
    cs = CriterionStep(cfg, rank, world, device)
    B, D, K = cfg["B"], cfg["D"], cfg["K"]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass

    # L2 flush between steps: READ a 256 MiB buffer (leaves the 126 MB L2 full of clean foreign lines; a
    # write-flush would leave it dirty and charge the write-back to the next kernel)
    flush_buf = None if args.no_flush else torch.empty(64 << 20, dtype=torch.int32, device=device).zero_()
    flush_out = torch.zeros(1, dtype=torch.int64, device=device)

    class _Flush:
        @staticmethod
        def fill_(_v):
            torch.sum(flush_buf, dim=(0,), keepdim=True, out=flush_out)
    flush = None if args.no_flush else _Flush

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """Per-step CUDA-event timing on the launching stream (L2 flushed before each step), summed."""
        evs = []
        barrier()
        for _ in range(steps):
            if flush is not None:
                flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        if world > 1:
            t = torch.tensor([ms], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # warm-up (also builds the EMA plan, TMA descriptors, bf16 shadow, cuBLAS handles)
    for _ in range(max(args.warmup, 3)):
        cs.step()
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---- eager module API (what the unchanged helper/loops_moma.py drives)
    eager_ms = timed(cs.step, args.steps)

    # ---- the same step captured once as a CUDA graph and replayed (no Python between kernels)
    from moma_b200.graphed import GraphedStep
    rows_per_step = B * world
    lib.moma_debug_launch_count(1)
    step_fn = cs.step if os.environ.get("MOMA_BENCH_SEQ") else cs.step_overlapped
    graphed = GraphedStep(step_fn, contrast=cs.contrast, rows_per_step=rows_per_step, warmup=3)
    launches_per_step = int(lib.moma_debug_launch_count(1)) // 4       # 3 warm-up calls + 1 captured call
    for _ in range(3):
        graphed.replay()
    total_ms = timed(graphed.replay, args.steps)
    launches = launches_per_step * args.steps

    # ---- end to end: host (pinned) inputs -> device, replay, host read of the loss.  The H2D copy of the NEXT
    #      step's features runs on a copy stream while the current step computes (double buffering through a
    #      staging buffer); every step still performs one H2D of its inputs and one D2H of its loss.
    copy_stream = torch.cuda.Stream(device)
    stage_s, stage_t = torch.empty_like(cs.feat_s, requires_grad=False), torch.empty_like(cs.feat_t)
    ev_copied, ev_consumed = torch.cuda.Event(), torch.cuda.Event()

    def prefetch():
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_consumed)
            stage_s.copy_(cs.host_s, non_blocking=True)
            stage_t.copy_(cs.host_t, non_blocking=True)
            ev_copied.record(copy_stream)

    ev_consumed.record()
    prefetch()

    def step_e2e():
        main = torch.cuda.current_stream()
        main.wait_event(ev_copied)
        with torch.no_grad():
            cs.feat_s.copy_(stage_s)
            cs.feat_t.copy_(stage_t)
        ev_consumed.record(main)
        loss = graphed.replay()
        prefetch()                                  # next step's inputs, overlapped with this step's compute
        return float(loss.item())

    for _ in range(3):
        step_e2e()
    e2e_ms = timed(step_e2e, args.steps)
    sampler.stop_flag = True
    sampler.join(timeout=1)

    # ---- per-kernel CUDA-event timing for the rooflines: cold L2 (flush first; the flush also hides
    #      the launch gap, so e0 -> e1 brackets the kernel alone), median of 20 launches
    def kernel_us(launch):
        ts = []
        for _ in range(20):
            if flush is not None:
                flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); launch(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        return ts[len(ts) // 2]

    from moma_b200._lib import BF16
    plan = ops._EMA_PLANS[ops.plan_key([p.detach() for p in cs.student], [p.detach() for p in cs.teacher])]
    ema_us = kernel_us(lambda: plan.run(ALPHA))
    n_q = B * world
    q_bf = torch.randn(n_q, D, device=device).to(torch.bfloat16)
    shadow = cs.contrast._shadow_of(cs.contrast.memory_shard if world > 1 else cs.contrast.memory)
    n_splits = ops.nce_num_splits(n_q, D, shadow.shape[0], BF16)
    st_buf = torch.empty((3, n_splits, n_q), device=device)
    o_buf = torch.empty((n_splits, n_q, D), device=device)
    nce_us = kernel_us(lambda: _lib.check(lib.moma_nce_partial(
        q_bf.data_ptr(), shadow.data_ptr(), n_q, D, shadow.shape[0], 1.0 / T_NCE, BF16, n_splits, st_buf[0].data_ptr(),
        st_buf[1].data_ptr(), st_buf[2].data_ptr(), o_buf.data_ptr(), torch.cuda.current_stream().cuda_stream)))

    ms_per_step = total_ms / args.steps
    value = B * world / (ms_per_step * 1e-3)
    e2e_value = B * world / (e2e_ms / args.steps * 1e-3)

    # ---- rooflines (algorithmic work per launch)
    k_local = K // world                              # n_q = B * world queries scored per rank (all-gathered)
    nce_flop = 4.0 * n_q * k_local * D                # S = Q.Queue^T and P.Queue, 2 FLOP/MAC each
    ema_bytes = 12.0 * cs.ema_elems                   # read ema, read src, write ema (fp32)
    # DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of one `ncu --set full` capture of the same
    # kernels at this configuration, committed under profiles/ (null when no capture exists for this config / world size)
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(f"{args.config}_n{world}", {})
    except Exception:
        pass
    tf_peak = peaks.get("bf16_tflops_sustained", 1421.9)
    bw_peak = peaks.get("hbm_gbs", 6452.2)
    roof_nce = {"kernel": "nce_tc_kernel (tcgen05 InfoNCE logits+CE fwd/bwd)", "bound": "tensor",
                "achieved": nce_flop / (nce_us * 1e-6) / 1e12 if nce_us else None, "peak": tf_peak, "unit": "TFLOP/s",
                "frac": (nce_flop / (nce_us * 1e-6) / 1e12 / tf_peak) if nce_us else None, "traffic": traffic.get("nce_tc2_kernel"),
                "us_per_launch": nce_us, "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback"}
    roof_ema = {"kernel": "ema_multi_kernel (momentum update)", "bound": "hbm",
                "achieved": ema_bytes / (ema_us * 1e-6) / 1e9 if ema_us else None, "peak": bw_peak, "unit": "GB/s",
                "frac": (ema_bytes / (ema_us * 1e-6) / 1e9 / bw_peak) if ema_us else None, "traffic": traffic.get("ema_multi_kernel"),
                "algorithmic_bytes": ema_bytes,
                "us_per_launch": ema_us, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback"}
    dominant = roof_ema if ema_us >= nce_us else roof_nce
    other = roof_nce if dominant is roof_ema else roof_ema

    out = {
        "metric": "MoMA criterion samples/sec", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": cfg["desc"], "global_batch": B * world, "feat_dim": D, "queue_K": K,
                   "s_dim": cfg["s_dim"], "t_dim": cfg["t_dim"], "heads": cfg["H"], "nce_t": T_NCE, "alpha": ALPHA,
                   "head": "mlp", "attn": "self", "ema_pair": "ResNet-18 student -> ResNet-18 momentum twin (11.18M)",
                   "queue": "replicated" if world == 1 else f"sharded by K over {world} ranks (cyclic)",
                   "l2": "no flush" if args.no_flush else "L2 flushed (256 MiB read) before every step; per-step CUDA events summed",
                   "timed_region": "EMA + projection heads + 3x attention + fused InfoNCE/CE fwd+bwd + enqueue (criterion step, L1)"},
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": cs.h2d_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_ms / args.steps,
                "path": "pinned host features -> H2D on a copy stream (double-buffered, overlapping the previous step) -> "
                        "static device buffers, graph replay, loss.item()"},
        "gpu_launches": launches,
        "gpu_launches_per_step": launches_per_step,
        "eager": {"ms_per_step": eager_ms / args.steps, "value": B * world / (eager_ms / args.steps * 1e-3),
                  "note": "same step through the eager module API (Python between kernels)"},
        "execution": "whole criterion step (fwd+bwd) captured once as a CUDA graph and replayed; the teacher branch, the "
                     "queue-attention + enqueue branch and the backbone EMA are forked onto side streams inside the capture; "
                     "kernels are chained with programmatic dependent launch",
        "roofline": dominant, "roofline_other": other,
        "clocks": sampler.summary(),
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(cfg, budget_s=args.cpu_seconds)
    return out


# ----------------------------------------------------------------------------- CPU arm
def cpu_baseline(cfg, budget_s=12.0):
    """The reference's CPU path (oracle/torch_port.py, the pinned port of the reference op sequence)
    on the host cores: bounded sample of the same workload."""
    from oracle.torch_port import PortCriterionStep
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    port = PortCriterionStep(cfg["s_dim"], cfg["t_dim"], cfg["D"], cfg["K"], T_NCE, ALPHA, cfg["H"],
                             resnet18_param_shapes(), seed=SEED)
    B = cfg["B"]
    torch.manual_seed(SEED + 1)
    fs = torch.randn(B, cfg["s_dim"], requires_grad=True)
    ft = torch.randn(B, cfg["t_dim"])
    for _ in range(2):
        port.step(fs, ft)
    times = []
    t_end = time.perf_counter() + budget_s
    while len(times) < 5 or (time.perf_counter() < t_end and len(times) < 400):
        t0 = time.perf_counter()
        port.step(fs, ft)
        times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return {"value": B / med, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{len(times)} full criterion steps of the same workload (median {med*1e3:.2f} ms/step), fp32, "
                      f"torch {torch.__version__} CPU kernels", "ms_per_step": med * 1e3}


def run_reference(args, cfg, rank, world):
    if rank != 0:
        return None
    steps = max(args.steps, 1)
    from oracle.torch_port import PortCriterionStep
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    port = PortCriterionStep(cfg["s_dim"], cfg["t_dim"], cfg["D"], cfg["K"], T_NCE, ALPHA, cfg["H"],
                             resnet18_param_shapes(), seed=SEED)
    B = cfg["B"]
    torch.manual_seed(SEED + 1)
    fs = torch.randn(B, cfg["s_dim"], requires_grad=True)
    ft = torch.randn(B, cfg["t_dim"])
    steps = min(steps, 200)
    for _ in range(min(max(args.warmup, 1), 5)):
        port.step(fs, ft)
    t0 = time.perf_counter()
    for _ in range(steps):
        port.step(fs, ft)
    dt = time.perf_counter() - t0
    value = B * steps / dt
    return {
        "impl": "reference", "metric": "MoMA criterion samples/sec", "value": value, "unit": "samples/s",
        "n_gpus": world, "steps": steps, "warmup": min(max(args.warmup, 1), 5), "ms_per_step": dt / steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["desc"], "global_batch": B, "feat_dim": cfg["D"], "queue_K": cfg["K"],
                   "note": "reference CPU path (pinned torch port of the reference op sequence), one replica on rank 0"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{steps} full criterion steps, fp32, all host threads"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS))
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    cfg = CONFIGS[args.config]

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))

    if args.impl == "reference":
        out = run_reference(args, cfg, rank, world)
        if out is not None:
            print(json.dumps(out), flush=True)
        return

    import torch.distributed as dist
    # stdout carries exactly one JSON line: everything else any library prints on fd 1 (e.g. NCCL's version
    # banner) is sent to stderr; the JSON goes to a private duplicate of the original stdout
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    out = run_ours(args, cfg, rank, world, local_rank)
    if rank == 0:
        json_out.write(json.dumps(out) + "\n")
        json_out.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
