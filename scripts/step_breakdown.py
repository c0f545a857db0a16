"""Real (CUDA-event, graph-replay) timings of the pieces of the criterion step (run on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from bench import CONFIGS
from moma_b200.step import CriterionStep

cfg = CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C2"]
dev = torch.device("cuda", 0)
torch.cuda.set_stream(torch.cuda.Stream(dev))
cs = CriterionStep(cfg, 0, 1, dev)
crit, opt = cs.crit, cs.opt
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
cs.contrast.use_device_pointer()


def graph_of(fn, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=torch.cuda.current_stream()):
        fn()
    return g


def timeit(g, cold, reps=30):
    ts = []
    for _ in range(reps):
        if cold:
            flush.fill_(1)
        else:
            torch.cuda._sleep(100000)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


state = {}
def ema(): cs.trainer.momentum_update(cs.student, cs.teacher, opt.alpha)
def teacher():
    with torch.no_grad():
        k = crit.embed_t(cs.feat_t)
    state["k"] = crit.atts_k(k); state["allk"] = crit.atts_queue(k)
def embed_t_only():
    with torch.no_grad():
        state["k0"] = crit.embed_t(cs.feat_t)
def atts_k_only(): state["k"] = crit.atts_k(state["k0"])
def student_fwd():
    f = crit.embed_s(cs.feat_s); state["fs"] = crit.atts_q(f)
def embed_s_only(): state["f0"] = crit.embed_s(cs.feat_s)
def nce_fwd():
    with torch.no_grad():
        out = cs.contrast(q=state["fs"].detach(), k=state["k"].detach(), all_k=state["allk"].detach())
        state["l"] = torch.nn.functional.cross_entropy(out[0], out[1])
def full_seq(): cs.step()
def full_ovl(): cs.step_overlapped()
def fwd_only():
    with torch.no_grad():
        k = crit.embed_t(cs.feat_t); f = crit.embed_s(cs.feat_s)
        f = crit.atts_q(f); k2 = crit.atts_k(k); a2 = crit.atts_queue(k)
        out = cs.contrast(q=f, k=k2, all_k=a2)
        state["l"] = torch.nn.functional.cross_entropy(out[0], out[1])
def empty(): pass

print(f"{'piece':16s} {'cold us':>9s} {'warm us':>9s}")
teacher(); student_fwd(); embed_t_only()
for name, fn in [("ema", ema), ("embed_t", embed_t_only), ("atts_k", atts_k_only), ("teacher(all)", teacher),
                 ("embed_s", embed_s_only), ("student_fwd", student_fwd), ("nce_fwd+enq", nce_fwd), ("fwd_only", fwd_only),
                 ("full_seq", full_seq), ("full_overlap", full_ovl)]:
    g = graph_of(fn)
    print(f"{name:16s} {timeit(g, True):9.1f} {timeit(g, False):9.1f}")
