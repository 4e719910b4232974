import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


class _ParityLog:
    """`parity(name, achieved, bound)`: asserts achieved <= bound AND appends the achieved error to
    gpurun_out/parity_report.txt, so that every run on the GPU box leaves a record of the errors actually reached next to
    the asserted bounds (copied to profiles/ per round)."""

    def __init__(self):
        self.path = os.path.join(ROOT, "gpurun_out", "parity_report.txt")

    def __call__(self, name, achieved, bound):
        achieved, bound = float(achieved), float(bound)
        try:
            os.makedirs(os.path.dirname(self.path), exist_ok=True)
            with open(self.path, "a") as f:
                f.write("%-84s achieved %.3e   bound %.3e\n" % (name, achieved, bound))
        except OSError:
            pass
        assert achieved <= bound, "%s: achieved %.3e > bound %.3e" % (name, achieved, bound)


@pytest.fixture(scope="session")
def parity():
    return _ParityLog()
