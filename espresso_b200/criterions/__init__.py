from .ctc_loss import CtcLossCriterion  # noqa: F401
from .label_smoothed_cross_entropy_v2 import LabelSmoothedCrossEntropyV2Criterion  # noqa: F401
from .transducer_loss import TransducerLossCriterion  # noqa: F401
