#!/usr/bin/env python
"""Parameter shapes (parameters() order) of the small-student backbones of BASELINE config C4, taken from the
UNMODIFIED reference model definitions -- run in the build container only (needs /root/reference):

    python scripts/make_c4_shapes.py      ->  bench_data/c4_param_shapes.json

EfficientNet-B0: models/efficientnet_pytorch/model.py:483-494 (EfficientNet.from_name, num_classes=4)
MobileNetV2:     models/__init__.py:42 key 'MobileNetV2_Imagenet' (models/mobilenetv2_imagenet.py)
ResNet-18/50:    models/resnet_imagenet.py (for cross-checking moma_b200.step.resnet18_param_shapes)
bench.py times the EMA (momentum_update, learning/contrast_trainer.py:207-211) of a student against its
same-architecture momentum twin on exactly these tensor lists."""
import json
import os
import sys

REF = os.environ.get("MOMA_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from models import model_dict                                      # noqa: E402
from models.efficientnet_pytorch.model import efficientnet         # noqa: E402

nets = {
    "efficientnet_b0": efficientnet(pretrained=False, num_classes=4),
    "mobilenetv2_imagenet": model_dict["MobileNetV2_Imagenet"](num_classes=4),
    "resnet18": model_dict["ResNet18"](num_classes=4),
    "resnet50": model_dict["ResNet50"](num_classes=4),
}
out = {}
for name, net in nets.items():
    shapes = [list(p.shape) for p in net.parameters()]
    out[name] = {"tensors": len(shapes), "elements": sum(p.numel() for p in net.parameters()), "shapes": shapes}
    print(name, out[name]["tensors"], out[name]["elements"])
with open(os.path.join(ROOT, "bench_data", "c4_param_shapes.json"), "w") as f:
    json.dump(out, f, separators=(",", ":"))
