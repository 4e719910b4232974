"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements (numpy / plain PyTorch fp32) of the reference algorithms on the Espresso hot path,
each citing the reference file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package, and only as the checker -- the product
(espresso_b200/) never imports it and has no CPU fallback.

Pinning: every restatement here is checked against the REAL reference (imported from /root/reference
through oracle/refshim in the authoring container) by oracle/pin_against_reference.py, which also writes
the golden fixtures under tests/golden/.  The reference's own tests hold no vectors for fbank / CMVN /
SpecAugment / rel-pos attention / Conformer / CTC (SURVEY.md §4 "Gap that matters"), so outputs of the
reference itself are the pin.
"""
