"""Kernel-level timeline of graph replays of the criterion step via torch.profiler (CUPTI)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from bench import CONFIGS
from moma_b200.step import CriterionStep
from moma_b200.graphed import GraphedStep

cfg = CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C2"]
mode = sys.argv[2] if len(sys.argv) > 2 else "seq"
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
    if rank != 0:
        sys.stdout = open(os.devnull, "w")
torch.cuda.set_stream(torch.cuda.Stream(dev, priority=-1))
cs = CriterionStep(cfg, rank, world, dev)
g = GraphedStep(cs.step if mode == "seq" else cs.step_overlapped, contrast=cs.contrast, rows_per_step=cfg["B"] * world)
for _ in range(5):
    g.replay()
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if len(sys.argv) > 3 and sys.argv[3] == "flush" else None
# event timing as bench.py does it
for cold in (True, False):
    ts = []
    for _ in range(30):
        if cold and flush is not None:
            flush.fill_(1)
        else:
            torch.cuda._sleep(100000)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    print("event-timed replay", "cold" if cold and flush is not None else "warm", f"{ts[len(ts)//2]:.1f} us")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        if flush is not None:
            flush.fill_(1)
        g.replay()
    torch.cuda.synchronize()
kin = [e for e in prof.profiler.kineto_results.events() if str(e.device_type()).endswith("CUDA") and e.duration_ns() > 0]
kin.sort(key=lambda e: e.start_ns())
n = len(kin) // 3
last = kin[-n:]
t0 = last[0].start_ns()
streams = {}
print(f"{'start':>8s} {'end':>8s} {'dur':>7s}  st  kernel")
busy = 0.0; end_max = t0
for e in last:
    sid = streams.setdefault(e.device_resource_id(), len(streams))
    st, du = (e.start_ns() - t0) / 1e3, e.duration_ns() / 1e3
    print(f"{st:8.1f} {st + du:8.1f} {du:7.1f}  {sid:2d}  {'    ' * sid}{e.name()[:70]}")
    busy += du; end_max = max(end_max, e.start_ns() + e.duration_ns())
print(f"total span {(end_max - t0) / 1e3:.1f} us, sum of kernel durations {busy:.1f} us, kernels {len(last)}")
