"""`ctc_loss` criterion (espresso/criterions/ctc_loss.py:40-169), B200-native.

forward(model, sample) -> (loss [sum-reduced], sample_size, logging_output), same contract as
FairseqCriterion.  The fp32 log-softmax + alpha/beta + gradient run in the fused CTC kernels
(esp_ctc_loss); the [T',B,V] fp32 log-prob tensor of the reference is never materialised.  Logging scalars
stay on the device (the trainer ships them through the gradient all-reduce tail) -- no per-step
`.item()` sync as in ctc_loss.py:96-106.
"""
import math

import torch

from .. import ops as _ops
from ..registry import register_criterion


class _CtcFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits_bt, V, in_lens, targets, tgt_lens, blank, zero_infinity, unit_grad):
        loss_b, grad = _ops.ctc_loss(logits_bt, V, in_lens, targets, tgt_lens, blank, zero_infinity, 1.0, True)
        ctx.grad = grad
        ctx.unit_grad = unit_grad
        return loss_b

    @staticmethod
    def backward(ctx, dloss_b):
        g = ctx.grad
        ctx.grad = None
        if not ctx.unit_grad:  # general autograd use; the B200 trainer back-propagates d(sum loss) = 1
            g = (g.float() * dloss_b[:, None, None]).to(g.dtype)
        return g, None, None, None, None, None, None, None


def compact_targets(target, pad_idx, eos_idx):
    """targets = non-pad & non-eos tokens, left-compacted (ctc_loss.py:76-83) -> int32 [B, U], int32 [B]."""
    keep = (target != pad_idx) & (target != eos_idx)
    order = torch.argsort((~keep).to(torch.int8), dim=1, stable=True)
    comp = torch.gather(target, 1, order)
    return comp.to(torch.int32).contiguous(), keep.sum(-1).to(torch.int32)


@register_criterion("ctc_loss")
class CtcLossCriterion(torch.nn.Module):
    def __init__(self, task=None, zero_infinity=True, sentence_avg=True, pad_idx=None, eos_idx=None, blank_idx=None,
                 unit_grad_output=False):
        super().__init__()
        d = getattr(task, "target_dictionary", None)
        self.pad_idx = pad_idx if pad_idx is not None else d.pad()
        self.eos_idx = eos_idx if eos_idx is not None else d.eos()
        # blank = <s> (bos) index: espresso/tasks/speech_recognition.py:324-328
        self.blank_idx = blank_idx if blank_idx is not None else d.index(getattr(task, "blank_symbol", "<s>"))
        self.zero_infinity = zero_infinity
        self.sentence_avg = sentence_avg
        self.unit_grad_output = unit_grad_output

    def forward(self, model, sample, reduce=True):
        net_output = model(**sample["net_input"])
        out = net_output["b200_out"]  # [B, T', ld] batch-major logits
        V = net_output["encoder_out"][0].size(-1)
        in_lens = net_output["src_lengths"][0].to(torch.int32)
        targets, tgt_lens = compact_targets(sample["target"], self.pad_idx, self.eos_idx)
        loss_b = _CtcFn.apply(out, V, in_lens, targets, tgt_lens, self.blank_idx, self.zero_infinity, self.unit_grad_output)
        loss = loss_b.sum() if reduce else loss_b
        # collate's count: every non-pad target token, eos included (espresso/data/asr_dataset.py:110-125); device scalar
        ntokens = sample["ntokens"] if "ntokens" in sample else sample["target"].ne(self.pad_idx).sum()
        nsent = sample["target"].size(0)
        sample_size = nsent if self.sentence_avg else ntokens
        logging_output = {"loss": loss.detach(), "ntokens": ntokens, "nsentences": nsent, "sample_size": sample_size}
        return loss, sample_size, logging_output

    @staticmethod
    def reduce_metrics(logging_outputs):
        """Aggregate like ctc_loss.py:133-160 (loss reported in base 2 per sample)."""
        f = lambda v: float(v.item()) if torch.is_tensor(v) else float(v)  # noqa: E731
        loss_sum = sum(f(l.get("loss", 0)) for l in logging_outputs)
        ntok = sum(f(l.get("ntokens", 0)) for l in logging_outputs)
        nsent = sum(f(l.get("nsentences", 0)) for l in logging_outputs)
        ss = sum(f(l.get("sample_size", 0)) for l in logging_outputs)
        return {"loss": loss_sum / max(ss, 1) / math.log(2), "ntokens": ntok, "nsentences": nsent, "sample_size": ss}

    @staticmethod
    def logging_outputs_can_be_summed():
        return True
