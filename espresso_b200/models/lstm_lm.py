"""`lstm_lm_espresso`: the LSTM language model the LibriSpeech recipe fuses at decoding time
(espresso/models/lstm_lm.py:88-252; examples/asr_librispeech/run_torchaudio.sh:50,180-198).  It is the attention-free
case of the speech_lstm decoder (embedding -> LSTMCell stack [-> additional_fc] -> output projection, optionally sharing
the embedding), with the reference's state-dict keys (`decoder.embed_tokens.weight`, `decoder.layers.{i}.*`, ...).

The LSTM cells run through ATen / cuDNN (not a north-star kernel); what this class adds is the incremental
protocol of espresso_b200.sequence_generator -- `init_incremental_state` / `decode_step` with beam reordering of the
cached (h, c) -- so the model can be used as `lm_model=` for shallow fusion, where the fusion itself (log-softmax of
both logit rows, `lprobs += lm_weight * lm`) runs in esp_beam_merge."""
from dataclasses import dataclass

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..registry import register_model
from .speech_lstm import SpeechLSTMDecoder


@dataclass
class LSTMLanguageModelEspressoConfig:
    dropout: float = 0.1
    decoder_embed_dim: int = 48
    decoder_hidden_size: int = 650
    decoder_layers: int = 2
    decoder_out_embed_dim: int = 650
    decoder_rnn_residual: bool = False
    share_embed: bool = False
    max_target_positions: int = 1024
    is_wordlm: bool = False  # word LM (look-ahead / multi-level fusion at decoding time), espresso/models/lstm_lm.py:64-70


@register_model("lstm_lm_espresso", dataclass=LSTMLanguageModelEspressoConfig)
class LSTMLanguageModelEspresso(nn.Module):
    def __init__(self, decoder, is_wordlm=False):
        super().__init__()
        self.decoder = decoder
        self.is_wordlm = is_wordlm

    @classmethod
    def build_model(cls, cfg, task):
        if cfg.share_embed and cfg.decoder_embed_dim != cfg.decoder_out_embed_dim:
            raise ValueError("--share-embed requires --decoder-embed-dim to match --decoder-out-embed-dim")
        # a word LM uses the task's word dictionary when there is one (lstm_lm.py:155-158)
        d = task.word_dictionary if getattr(cfg, "is_wordlm", False) and hasattr(task, "word_dictionary") else task.target_dictionary
        dec = SpeechLSTMDecoder(d, cfg.decoder_embed_dim, cfg.decoder_hidden_size, cfg.decoder_out_embed_dim, cfg.decoder_layers,
                                cfg.dropout, cfg.dropout, encoder_output_units=0, attn_dim=0,
                                residual=False,  # the reference does not forward decoder_rnn_residual to the decoder (lstm_lm.py:175-191)
                                share_input_output_embed=cfg.share_embed, max_target_positions=cfg.max_target_positions)
        return cls(dec, is_wordlm=getattr(cfg, "is_wordlm", False))

    def finalize_(self, device, dtype=torch.bfloat16):
        self.to(device=device, dtype=dtype)
        self.eval()
        return self

    def forward(self, src_tokens, **unused):
        """Teacher-forced logits [B, U, V] (fairseq language-model convention: the input is `src_tokens`)."""
        return self.decoder(src_tokens), None

    def get_normalized_probs(self, net_output, log_probs, sample=None):
        x = net_output[0].float()
        return F.log_softmax(x, dim=-1) if log_probs else F.softmax(x, dim=-1)

    # ---- generator protocol (espresso_b200/sequence_generator.py) --------------------------------------------------
    def max_decoder_positions(self):
        return self.decoder.max_positions()

    def forward_encoder(self, net_input):
        return None

    def init_incremental_state(self, encoder_out, bsz, beam):
        N = bsz * beam
        w = self.decoder.embed_tokens.weight
        z = lambda: [w.new_zeros(N, self.decoder.hidden_size) for _ in self.decoder.layers]  # noqa: E731
        return {"h": z(), "c": z()}

    @torch.no_grad()
    def decode_step(self, step, tokens, state, new_order):
        """tokens int32 [N, L]; consumes column `step`; new_order (or None): row permutation chosen by the last search
        step, applied to the cached (h, c) like reorder_incremental_state (speech_lstm.py:974-1000)."""
        if new_order is not None:
            idx = new_order.long()
            state["h"] = [h.index_select(0, idx) for h in state["h"]]
            state["c"] = [c.index_select(0, idx) for c in state["c"]]
        x = self.decoder.embed_tokens(tokens[:, step].long())
        y, state["h"], state["c"], _ = self.decoder.step(x, state["h"], state["c"], None)
        logits = self.decoder.output_layer(y)
        V = logits.size(-1)
        ldV = (V + 7) // 8 * 8
        if logits.dtype == torch.bfloat16 and ldV != V:  # 16-byte rows for the fused log-softmax / fusion kernel
            logits = F.pad(logits, (0, ldV - V))
        return logits.contiguous(), True
