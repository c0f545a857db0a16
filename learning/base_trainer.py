from moma_b200.contrast_trainer import BaseTrainer  # noqa: F401
