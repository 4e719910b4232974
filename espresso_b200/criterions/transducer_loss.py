"""`transducer_loss` criterion (espresso/criterions/transducer_loss.py:40-192), B200-native: the RNN-T loss and its
gradient w.r.t. the joint logits run in esp_rnnt_loss (fused log-softmax; the reference calls
torchaudio.functional.rnnt_loss(..., clamp=-1.0, reduction="sum") :130-140).  Targets exclude eos (include_eos=False,
:78-92): target = sample["target"][:, :-1], lengths = non-pad & non-eos count."""
import math

import torch

from .. import ops as _ops
from ..registry import register_criterion


class _RnntFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, V, t_lens, u_lens, targets, blank, unit_grad):
        loss, grad = _ops.rnnt_loss(logits, V, t_lens, u_lens, targets, blank, 1.0, True)
        ctx.grad, ctx.unit_grad = grad, unit_grad
        return loss

    @staticmethod
    def backward(ctx, dloss):
        g = ctx.grad
        ctx.grad = None
        if not ctx.unit_grad:
            g = (g.float() * dloss.view(-1, 1, 1, 1)).to(g.dtype)
        return g, None, None, None, None, None, None


@register_criterion("transducer_loss")
class TransducerLossCriterion(torch.nn.Module):
    def __init__(self, task=None, sentence_avg=True, pad_idx=None, eos_idx=None, blank_idx=None, unit_grad_output=False):
        super().__init__()
        d = getattr(task, "target_dictionary", None)
        self.pad_idx = pad_idx if pad_idx is not None else d.pad()
        self.eos_idx = eos_idx if eos_idx is not None else d.eos()
        self.blank_idx = blank_idx if blank_idx is not None else d.index("<s>")
        self.sentence_avg = sentence_avg
        self.unit_grad_output = unit_grad_output

    def forward(self, model, sample, reduce=True):
        net_output, enc_lens = model(**sample["net_input"])
        out = model._b200_out  # [B, T', U+1, ldV] bf16
        V = net_output.size(-1)
        target = sample["target"]
        u_lens = ((target != self.pad_idx) & (target != self.eos_idx)).sum(-1).to(torch.int32)
        tg = target[:, :-1].to(torch.int32).contiguous()
        loss_b = _RnntFn.apply(out, V, enc_lens.to(torch.int32), u_lens, tg, self.blank_idx, self.unit_grad_output)
        loss = loss_b.sum() if reduce else loss_b
        ntokens = sample["ntokens"] if "ntokens" in sample else sample["target"].ne(self.pad_idx).sum()
        nsent = target.size(0)
        sample_size = nsent if self.sentence_avg else ntokens
        return loss, sample_size, {"loss": loss.detach(), "ntokens": ntokens, "nsentences": nsent, "sample_size": sample_size}

    @staticmethod
    def reduce_metrics(logging_outputs):
        f = lambda v: float(v.item()) if torch.is_tensor(v) else float(v)  # noqa: E731
        ls = sum(f(l.get("loss", 0)) for l in logging_outputs)
        ss = sum(f(l.get("sample_size", 0)) for l in logging_outputs)
        return {"loss": ls / max(ss, 1) / math.log(2), "sample_size": ss,
                "ntokens": sum(f(l.get("ntokens", 0)) for l in logging_outputs)}

    @staticmethod
    def logging_outputs_can_be_summed():
        return True
