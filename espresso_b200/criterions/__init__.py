from .ctc_loss import CtcLossCriterion  # noqa: F401
