"""espresso_b200 -- B200-native (sm_100a) implementation of the Espresso ASR hot path.

Host code is Python/PyTorch (device memory, streams, torch.distributed); every piece of arithmetic on
the path is a hand-written CUDA kernel in ``libespresso_b200.so`` reached through the C ABI declared in
``include/espresso_b200.h``.  There is no CPU or library fallback: importing :mod:`espresso_b200.lib`
raises if the shared library has not been built.
"""
__version__ = "0.1.0"
