"""Loss and metrics of the hot path (the reference's general/callbacks/plots utilities are
orchestration and are used unchanged from the reference)."""
from .loss import BalancedCELoss, DeepSupervisionLoss, DiceBCELoss, DiceLoss, create_loss_function
from .metrics import SegmentationMetrics, compute_dice, compute_iou

__all__ = ["DiceLoss", "BalancedCELoss", "DiceBCELoss", "DeepSupervisionLoss", "create_loss_function",
           "SegmentationMetrics", "compute_iou", "compute_dice"]
