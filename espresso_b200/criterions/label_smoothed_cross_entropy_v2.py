"""`label_smoothed_cross_entropy_v2` criterion (espresso/criterions/label_smoothed_cross_entropy_v2.py:49-243;
uniform, unigram and temporal smoothing), B200-native: fp32 log-softmax + NLL + smoothing + their gradient in one fused kernel
(esp_lsce_loss); the [B*U, V] fp32 log-prob tensor of the reference is never materialised.  The model is called
with `epoch=` like the reference does (:167)."""
import math

import torch

from .. import ops as _ops
from ..registry import register_criterion


class _LsceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits_bu, V, targets, pad_idx, eps, unit_grad, smoothing=0, unigram=None):
        B, U, ld = logits_bu.shape
        loss, nll, grad = _ops.lsce_loss(logits_bu.view(B * U, ld), V, targets, pad_idx, eps, 1.0, True,
                                         smoothing=smoothing, unigram=unigram, U=U)
        ctx.grad = grad.view(B, U, ld)
        ctx.unit_grad = unit_grad
        ctx.mark_non_differentiable(nll)
        return loss, nll

    @staticmethod
    def backward(ctx, dloss, dnll):
        g = ctx.grad
        ctx.grad = None
        if not ctx.unit_grad:
            g = (g.float() * dloss.view(g.shape[0], g.shape[1], 1)).to(g.dtype)
        return g, None, None, None, None, None, None, None


@register_criterion("label_smoothed_cross_entropy_v2")
class LabelSmoothedCrossEntropyV2Criterion(torch.nn.Module):
    _SMOOTHING = {"uniform": 0, "unigram": 1, "temporal": 2}

    def __init__(self, task=None, sentence_avg=False, label_smoothing=0.1, smoothing_type="uniform", pad_idx=None,
                 unit_grad_output=False, unigram_pseudo_count=1.0, unigram_counts=None):
        """unigram smoothing builds the distribution like the reference (:151-155): dictionary.count (or
        `unigram_counts`) + pseudo count, normalised."""
        super().__init__()
        if smoothing_type not in self._SMOOTHING:
            raise ValueError("Unsupported smoothing type: {}".format(smoothing_type))
        self.smoothing_type = smoothing_type
        self.unigram_tensor = None
        if smoothing_type == "unigram":
            counts = unigram_counts if unigram_counts is not None else task.target_dictionary.count
            u = torch.as_tensor(counts, dtype=torch.float32).clone() + unigram_pseudo_count
            self.unigram_tensor = u / u.sum()
        self.padding_idx = pad_idx if pad_idx is not None else task.target_dictionary.pad()
        self.eps = label_smoothing
        self.sentence_avg = sentence_avg
        self.unit_grad_output = unit_grad_output

    def forward(self, model, sample, reduce=True, epoch=1):
        net_output = model(**sample["net_input"], epoch=epoch)
        out = net_output[1]["b200_out"]  # [B, U, ldV] bf16
        V = net_output[0].size(-1)
        target = sample["target"]
        tgt = target.reshape(-1).to(torch.int32)
        if self.unigram_tensor is not None and self.unigram_tensor.device != out.device:
            self.unigram_tensor = self.unigram_tensor.to(out.device)
        loss_r, nll_r = _LsceFn.apply(out, V, tgt, self.padding_idx, self.eps, self.unit_grad_output,
                                      self._SMOOTHING[self.smoothing_type], self.unigram_tensor)
        loss, nll = loss_r.sum(), nll_r.sum()
        ntokens = sample["ntokens"] if "ntokens" in sample else target.ne(self.padding_idx).sum()
        nsent = target.size(0)
        sample_size = nsent if self.sentence_avg else ntokens
        logging_output = {"loss": loss.detach(), "nll_loss": nll.detach(), "ntokens": ntokens, "nsentences": nsent,
                          "sample_size": sample_size}
        return loss, sample_size, logging_output

    @staticmethod
    def reduce_metrics(logging_outputs):
        f = lambda v: float(v.item()) if torch.is_tensor(v) else float(v)  # noqa: E731
        loss = sum(f(l.get("loss", 0)) for l in logging_outputs)
        nll = sum(f(l.get("nll_loss", 0)) for l in logging_outputs)
        ntok = sum(f(l.get("ntokens", 0)) for l in logging_outputs)
        ss = sum(f(l.get("sample_size", 0)) for l in logging_outputs)
        return {"loss": loss / max(ss, 1) / math.log(2), "nll_loss": nll / max(ntok, 1) / math.log(2),
                "ppl": 2 ** (nll / max(ntok, 1) / math.log(2)), "ntokens": ntok, "sample_size": ss}

    @staticmethod
    def logging_outputs_can_be_summed():
        return True
