"""``from learning.contrast_trainer import ContrastTrainer`` (train_student_moma.py:37) -> moma_b200."""
from moma_b200.contrast_trainer import ContrastTrainer, AverageMeter, accuracy  # noqa: F401
from .base_trainer import BaseTrainer  # noqa: F401
