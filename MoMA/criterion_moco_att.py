"""``from MoMA.criterion_moco_att import CMO`` (train_student_moma.py:39) -> moma_b200."""
from moma_b200.criterion_moco_att import (  # noqa: F401
    eps, Normalize, Flatten, input_mapping_torch, RFF_ST, RFF, RFF_fixed,
    Attention, Attention_viz, Attention_, Attention2, CMO, CMO_EmaTec)
