"""``from MoMA.mem_moco import build_mem`` (train_student_moma.py:38) -> moma_b200."""
from moma_b200.mem_moco import BaseMoCo, MoCo, MoCoAtt, MoCoST, MoCoSSTT, build_mem  # noqa: F401
