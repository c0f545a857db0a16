"""Launch the tcgen05 InfoNCE kernel a few times at one shape (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.sweep_nce import time_kernel
B, D, K = (int(v) for v in (sys.argv[1:4] or (512, 128, 65536)))
us, sp = time_kernel(B, D, K, None, cold=True, reps=5)
print(f"B{B} D{D} K{K} splits {sp}: {us:.1f} us")
