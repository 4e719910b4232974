from .lr_scheduler import (NoamLRScheduler, PolynomialDecayV2LRScheduler, ReduceLROnPlateauV2LRScheduler,  # noqa: F401
                           TriStageLRScheduler)
