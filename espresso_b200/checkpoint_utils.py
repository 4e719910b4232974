"""Checkpoint interchange with the reference: read and write fairseq's `checkpoint_*.pt` layout
(fairseq/checkpoint_utils.py:281-345 `load_checkpoint_to_cpu`, :579-712 `_upgrade_state_dict`, fairseq/trainer.py:395-480
`state_dict` / `save_checkpoint`):

    {"cfg": ..., "args": None, "model": model.state_dict(), "criterion": {...},
     "optimizer_history": [{"criterion_name", "optimizer_name", "lr_scheduler_state", "num_updates"}],
     "extra_state": {"train_iterator": {"epoch", "iterations_in_epoch"}, ...},
     "last_optimizer_state": {"state": {0: {"step", "exp_avg", "exp_avg_sq"}}, "param_groups": [...]}}

The model part interchanges because the state-dict keys are the reference's (tests/test_host_orchestration.py).  The
optimizer part follows fairseq's bf16 optimizer: ONE flat fp32 parameter in `model.parameters()` order without padding
(fairseq/optim/fp16_optimizer.py:40-94) -- the flat buffers here are aligned and grouped differently, so Adam's moments are
re-packed tensor by tensor.  A checkpoint written by a fairseq that pickled its omegaconf / argparse config objects is
readable without fairseq installed: unknown classes are replaced by inert stand-ins while unpickling.
"""
import io
import pickle
from collections import OrderedDict

import torch


class _Inert:
    """Stand-in for a pickled class that is not importable here (omegaconf nodes, fairseq dataclasses, ...)."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self.__dict__["_state"] = state


class _TolerantUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        try:
            return super().find_class(module, name)
        except Exception:
            return type(name, (_Inert,), {"__module__": module})


class _TolerantPickle:
    __name__ = "tolerant_pickle"
    Unpickler = _TolerantUnpickler
    load = staticmethod(lambda f, **kw: _TolerantUnpickler(f, **kw).load())
    loads = staticmethod(lambda b, **kw: _TolerantUnpickler(io.BytesIO(b), **kw).load())


def load_checkpoint_to_cpu(path):
    """-> the checkpoint dict with at least `model` (tensors on CPU); old layouts are upgraded like the reference does."""
    with open(path, "rb") as f:
        state = torch.load(f, map_location="cpu", weights_only=False, pickle_module=_TolerantPickle)
    if "model" not in state:
        raise KeyError("%s is not a fairseq checkpoint: no 'model' entry" % path)
    if "optimizer_history" not in state:  # fairseq/checkpoint_utils.py:583-589
        state["optimizer_history"] = [{"criterion_name": "CrossEntropyCriterion", "best_loss": state.get("best_loss")}]
        state["last_optimizer_state"] = state.pop("optimizer", None)
    hist = state["optimizer_history"][-1]
    hist.setdefault("num_updates", 0)
    state.setdefault("extra_state", {})
    state["extra_state"].setdefault("train_iterator", {"epoch": state["extra_state"].get("epoch", 0),
                                                       "iterations_in_epoch": state["extra_state"].get("batch_offset", 0)})
    return state


def load_model_state(model, state, strict=True):
    """model.load_state_dict(state['model']) tolerant of the reference's bookkeeping buffers, then refresh the fp32 master
    weights of a finalised model."""
    sd = state["model"] if "model" in state else state
    own = model.state_dict()
    skip = lambda k: k.endswith("version") or k.endswith("_float_tensor") or k.endswith("num_batches_tracked")  # noqa: E731
    missing = [k for k in own if k not in sd and not skip(k)]
    unexpected = [k for k in sd if k not in own and not skip(k)]
    if strict and (missing or unexpected):
        raise RuntimeError("checkpoint does not match the model: missing %s, unexpected %s" % (missing[:5], unexpected[:5]))
    flat = getattr(model, "flat", None)
    has_master = flat is not None and getattr(flat, "p32", None) is not None
    with torch.no_grad():
        for k, v in sd.items():
            if k in own:
                own[k].copy_(v.to(own[k].dtype))
                if has_master and k in flat.offsets:  # keep the checkpoint's fp32 precision in the master copy
                    flat._view(flat.p32, k).copy_(v.float())
    return missing, unexpected


def _fairseq_param_order(model):
    return [(n, p) for n, p in model.named_parameters()]


def optimizer_state_to_fairseq(trainer):
    """Adam moments of the flat buffers -> fairseq FP16Optimizer layout (one flat fp32 tensor, parameter order, no padding)."""
    flat = trainer.flat
    m = torch.cat([flat._view(flat.m, n).reshape(-1).float().cpu() for n, _ in _fairseq_param_order(trainer.model)])
    v = torch.cat([flat._view(flat.v, n).reshape(-1).float().cpu() for n, _ in _fairseq_param_order(trainer.model)])
    return {"state": {0: {"step": trainer.num_updates, "exp_avg": m, "exp_avg_sq": v}},
            "param_groups": [{"lr": trainer.get_lr(), "betas": tuple(trainer.betas), "eps": trainer.eps,
                              "weight_decay": trainer.weight_decay, "params": [0]}]}


def optimizer_state_from_fairseq(trainer, opt_state):
    st = opt_state["state"][0] if 0 in opt_state["state"] else next(iter(opt_state["state"].values()))
    flat, off = trainer.flat, 0
    for n, p in _fairseq_param_order(trainer.model):
        k = p.numel()
        flat._view(flat.m, n).copy_(st["exp_avg"][off: off + k].view(p.shape))
        flat._view(flat.v, n).copy_(st["exp_avg_sq"][off: off + k].view(p.shape))
        off += k
    assert off == st["exp_avg"].numel(), "optimizer state size does not match the model"
    trainer.num_updates = int(st.get("step", trainer.num_updates))


def save_checkpoint(path, model, trainer=None, cfg=None, extra_state=None, criterion_name="CtcLossCriterion"):
    """Write a checkpoint the reference's `load_checkpoint_to_cpu` + `load_state_dict` accept.  Model weights are saved from
    the fp32 master copy when there is one (what fairseq keeps in its fp32 optimizer copy), else in parameter precision."""
    flat = getattr(model, "flat", None)
    sd = OrderedDict()
    for k, v in model.state_dict().items():
        if flat is not None and getattr(flat, "p32", None) is not None and k in flat.offsets:
            sd[k] = flat._view(flat.p32, k).detach().cpu().clone()
        else:
            sd[k] = v.detach().cpu().clone()
    state = {"args": None, "cfg": cfg, "model": sd, "criterion": {},
             "optimizer_history": [{"criterion_name": criterion_name, "optimizer_name": "FP16Optimizer",
                                    "lr_scheduler_state": {"best": None},
                                    "num_updates": trainer.num_updates if trainer is not None else 0}],
             "extra_state": dict({"train_iterator": {"epoch": 1, "iterations_in_epoch": 0}}, **(extra_state or {})),
             "last_optimizer_state": optimizer_state_to_fairseq(trainer) if trainer is not None else None}
    torch.save(state, path)
    return state
