def compose(*a, **k):
    raise RuntimeError("hydra stub: compose unavailable")


def initialize(*a, **k):
    raise RuntimeError("hydra stub: initialize unavailable")
