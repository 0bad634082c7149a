"""Loss, metrics and the EMA update of the hot path (the rest of the reference's general / callbacks /
plots utilities is orchestration and is used unchanged from the reference)."""
from .general import ModelEMA
from .loss import BalancedCELoss, DeepSupervisionLoss, DiceBCELoss, DiceLoss, create_loss_function
from .metrics import SegmentationMetrics, compute_dice, compute_iou

__all__ = ["DiceLoss", "BalancedCELoss", "DiceBCELoss", "DeepSupervisionLoss", "create_loss_function",
           "SegmentationMetrics", "compute_iou", "compute_dice", "ModelEMA"]

# names the reference's unet/utils/__init__.py re-exports from modules this package does not rebuild
# (set_seed, get_device, load_config, EarlyStopping, ModelCheckpoint) come from an attached checkout
import sys as _sys

from .. import overlay as _overlay

_overlay.install(_sys.modules[__name__])
__getattr__ = _overlay.package_getattr(__name__, ("general", "callbacks"))
