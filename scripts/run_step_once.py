"""A few eager criterion steps (CriterionStep.step_overlapped) of one configuration -- the short command the ncu captures
under profiles/ are taken from:   python scripts/run_step_once.py [C2|C3|C5] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import CONFIGS
from moma_b200.step import CriterionStep

cfg = CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C3"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
torch.cuda.set_stream(torch.cuda.Stream(dev))
cs = CriterionStep(cfg, 0, 1, dev)
flush = torch.zeros(64 << 20, dtype=torch.float32, device=dev)
for i in range(steps):
    flush.sum()
    loss = cs.step_overlapped()
torch.cuda.synchronize()
print("loss", float(loss.item()), "index", cs.contrast.index)
