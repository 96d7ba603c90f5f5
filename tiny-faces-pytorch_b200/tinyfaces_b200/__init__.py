"""tinyfaces_b200 -- Blackwell-native drop-in for the hot path of varunagrawal/tiny-faces-pytorch.

Mirrors the reference's call surface for that path only:
    tinyfaces_b200.models.model.DetectionModel      <- tinyfaces/models/model.py:7-128
    tinyfaces_b200.models.loss.DetectionCriterion   <- tinyfaces/models/loss.py:24-97
    tinyfaces_b200.models.utils.get_bboxes          <- tinyfaces/models/utils.py:4-76
    tinyfaces_b200.evaluation.get_detections / nms  <- tinyfaces/evaluation.py:20-87 (+ torchvision.ops.nms)
    tinyfaces_b200.trainer.train                    <- tinyfaces/trainer.py:68-90
All compute goes through the C-ABI library libtinyfaces_b200.so (include/tinyfaces_b200.h).
"""
__version__ = "0.1.0"
