"""How much do TF32 head GEMMs move the criterion step's loss / gradients? (GPU box)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import CONFIGS
from moma_b200.step import CriterionStep

def run(cfg, tf32, precision):
    import moma_b200
    torch.backends.cuda.matmul.allow_tf32 = tf32 is True
    cs = CriterionStep(cfg, 0, 1, torch.device("cuda", 0))
    if tf32 == "teacher":
        orig = cs.crit.embed_t.forward
        def fwd(x):
            torch.backends.cuda.matmul.allow_tf32 = True
            try:
                return orig(x)
            finally:
                torch.backends.cuda.matmul.allow_tf32 = False
        cs.crit.embed_t.forward = fwd
    moma_b200.set_precision(precision)
    loss = cs.step()
    grads = {n: p.grad.clone() for n, p in cs.crit.named_parameters() if p.grad is not None}
    return loss.item(), cs.feat_s.grad.clone(), grads

for name in ("C2", "C3"):
    cfg = CONFIGS[name]
    ref = run(cfg, False, "fp32")
    for tf32, prec in ((False, "bf16"), ("teacher", "bf16"), ("teacher", "fp32")):
        out = run(cfg, tf32, prec)
        rel = lambda a, b: ((a - b).norm() / b.norm()).item()
        worst = max(rel(out[2][n], ref[2][n]) for n in ref[2])
        print(f"{name} tf32={tf32} nce={prec}: loss rel {abs(out[0]-ref[0])/abs(ref[0]):.2e}  dfeat_s rel {rel(out[1], ref[1]):.2e}  worst param-grad rel {worst:.2e}")
