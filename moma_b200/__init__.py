"""moma_b200 -- B200-native (sm_100a) implementation of MoMA's momentum-contrastive,
multi-head-attention distillation criterion step.

Layout
  csrc/                    hand-written CUDA kernels + the C ABI (include/moma_b200.h)
  _lib.py / ops.py         ctypes binding and torch-facing wrappers (PyTorch = plumbing)
  mem_moco.py              mirror of reference MoMA/mem_moco.py      (MoCo, build_mem, ...)
  criterion_moco_att.py    mirror of reference MoMA/criterion_moco_att.py (CMO, Attention, ...)
  contrast_trainer.py      mirror of reference learning/contrast_trainer.py (ContrastTrainer)
  sharded.py               K-sharded queue across ranks (NCCL all-gather + partial-LSE combine)
The top-level packages ``MoMA`` and ``learning`` re-export these under the reference's import
paths so helper/loops_moma.py and train_student_moma.py run unchanged.
"""
from . import ops  # noqa: F401
from .ops import get_precision, set_precision  # noqa: F401
from .lazy_logits import LazyLogits  # noqa: F401
from .mem_moco import BaseMoCo, MoCo, MoCoAtt, MoCoST, MoCoSSTT, build_mem  # noqa: F401
from .criterion_moco_att import CMO, CMO_EmaTec, Attention, Attention2, Attention_viz, Normalize, Flatten  # noqa: F401
from .contrast_trainer import ContrastTrainer, accuracy  # noqa: F401

__version__ = "0.1.0"
