"""fairseq `--user-dir` plugin: registers the B200-native model / criterion with the REAL fairseq registries.

    fairseq-train ... --user-dir /path/to/repo/espresso_b200_plugin \
        --arch speech_conformer_encoder_model_b200 --criterion ctc_loss_b200          (same recipe otherwise)

`utils.import_user_module` (fairseq/utils.py:464-511; first call of fairseq_cli/train.py:48 and
espresso/speech_recognize.py:75) imports this package with the repository root on sys.path, so `espresso_b200`
resolves next to it.  What is registered:

  * model  `speech_transformer_encoder_model_b200` (+ archs `speech_transformer_encoder_model_b200`,
    `speech_conformer_encoder_model_b200`): the B200 encoder model class mixed with `BaseFairseqModel`, so it passes
    `issubclass(cls, BaseFairseqModel)` (fairseq/models/__init__.py:128-134), keeps the reference's dataclass
    (`SpeechTransformerConfig`: same yaml / CLI) and the reference's state-dict keys.  fairseq moves / casts models after
    construction (trainer.py:105-107), so the flat bf16/fp32 buffers are (re)built lazily on the first forward; after each
    backward the gradients of the flat fp32 buffer are handed to autograd's `.grad` slots (bf16, what
    fairseq/optim/fp16_optimizer.py:109-145 reads), so fairseq's own trainer / optimizer / DDP work unchanged.
  * criterion `ctc_loss_b200`: `FairseqCriterion` subclass (fairseq/registry.py:64-82) with the reference's config
    dataclass and `reduce_metrics`, forwarding to the fused CTC kernels.

The product's own `espresso_b200.trainer.Trainer` (flat buffers, one all-reduce, fused Adam, CUDA graphs) is the fast
path; this plugin is the drop-in path under an unmodified fairseq.
"""
import torch

from fairseq.criterions import FairseqCriterion, register_criterion
from fairseq.models import BaseFairseqModel, register_model, register_model_architecture

from espresso.criterions.ctc_loss import CtcLossCriterion as _RefCtc
from espresso.criterions.ctc_loss import CtcLossCriterionConfig
from espresso.models.transformer.speech_transformer_config import SpeechTransformerConfig as _RefConfig

from espresso_b200.criterions import CtcLossCriterion as _B200Ctc
from espresso_b200.models.transformer import speech_transformer_encoder_model as _m


@register_model("speech_transformer_encoder_model_b200", dataclass=_RefConfig)
class SpeechTransformerEncoderModelB200(_m.SpeechTransformerEncoderModel, BaseFairseqModel):
    """espresso/models/transformer/speech_transformer_encoder_model.py:36-150 on the sm_100a kernels."""

    def __init__(self, cfg, encoder):
        _m.SpeechTransformerEncoderModel.__init__(self, cfg, encoder)   # nn.Module.__init__ via the MRO
        self._is_generation_fast = False
        self._flat_device = None

    # fairseq moves (.to / .cuda) and casts (.half / .bfloat16) models after build_model: flat views would be lost
    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        self._flat_device = None
        return out

    def _ensure_flat(self, device):
        p = next(self.parameters())
        if self._flat_device != device or self.encoder.engine is None or p.dtype != torch.bfloat16:
            self.finalize_(device)
            self._flat_device = device
            # end-of-backward hook (queued by the encoder's autograd node): flat fp32 gradients -> .grad
            self.encoder.engine.after_backward = self._publish_grads

    def forward(self, src_tokens, src_lengths, **kwargs):
        self._ensure_flat(src_tokens.device)
        return _m.SpeechTransformerEncoderModel.forward(self, src_tokens, src_lengths, **kwargs)

    def _publish_grads(self):
        flat = self.flat
        for n, p in self.named_parameters():
            if n.startswith("encoder.pre_encoder.") and p.grad is not None and getattr(p, "_b200_seen", None) is not p.grad:
                # torch-executed conv front: autograd already produced this gradient
                p._b200_seen = p.grad
                continue
            g = flat.grad(n).to(p.dtype)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.add_(g)
            p._b200_seen = p.grad
        flat.g32.zero_()


@register_model_architecture("speech_transformer_encoder_model_b200", "speech_conformer_encoder_model_b200")
def speech_conformer_encoder_model_b200(cfg):
    cfg.encoder.layer_type = "conformer"


@register_criterion("ctc_loss_b200", dataclass=CtcLossCriterionConfig)
class CtcLossCriterionB200(FairseqCriterion):
    """espresso/criterions/ctc_loss.py:40-169 with the fused CTC kernels."""

    def __init__(self, cfg, task):
        super().__init__(task)
        self.impl = _B200Ctc(task, zero_infinity=cfg.zero_infinity, sentence_avg=cfg.sentence_avg)

    def forward(self, model, sample, reduce=True):
        return self.impl(model, sample, reduce=reduce)

    reduce_metrics = staticmethod(_RefCtc.reduce_metrics)

    @staticmethod
    def logging_outputs_can_be_summed() -> bool:
        return True
