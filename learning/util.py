from moma_b200.contrast_trainer import AverageMeter, accuracy  # noqa: F401
