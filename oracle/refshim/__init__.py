"""Import shim for the *real* reference (/root/reference) -- authoring container only.

Follows SURVEY.md §8(c): stub packages for the missing third-party deps and a
`dataclasses.dataclass` wrapper that lets fairseq's mutable dataclass defaults
construct on Python >= 3.11.  Used only by `oracle/pin_against_reference.py`
(golden-fixture generation + pinning the oracle restatement).  /root/reference
does not exist on the GPU box, so nothing under tests/ -m gpu, smoke() or
bench.py imports this module.
"""
import dataclasses
import os
import sys

REFERENCE_ROOT = os.environ.get("ESPRESSO_REFERENCE_ROOT", "/root/reference")
_done = False


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "espresso"))


def activate():
    """Make `import fairseq, espresso` resolve to the reference tree."""
    global _done
    if _done:
        return
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    import torch  # noqa: F401  (must be imported before the dataclass patch)
    import torchaudio  # noqa: F401

    stubs = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")
    sys.path.insert(0, REFERENCE_ROOT)
    sys.path.insert(0, stubs)

    orig = dataclasses.dataclass

    def patched(cls=None, **kw):
        def wrap(c):
            if c.__module__.split(".")[0] in ("fairseq", "espresso", "examples"):
                kw2 = dict(kw)
                kw2.setdefault("unsafe_hash", True)
                kw2.pop("eq", None)
                return orig(c, **kw2)
            return orig(c, **kw)

        return wrap if cls is None else wrap(cls)

    dataclasses.dataclass = patched
    try:
        import fairseq  # noqa: F401
        import espresso  # noqa: F401
    finally:
        dataclasses.dataclass = orig
    _done = True
