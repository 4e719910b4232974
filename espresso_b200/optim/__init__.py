from .lr_scheduler import NoamLRScheduler, TriStageLRScheduler  # noqa: F401
