"""Beam search for transducer models: the "adaptive expansion search" the transducer recipes decode with
(examples/asr_librispeech/run_transformer_transducer.sh:263-265: --beam 5 --transducer-expansion-beta 2
--transducer-expansion-gamma 2.3 --transducer-prefix-alpha 1), following
espresso/tools/transducer_beam_search_decoder.py:21-601 and espresso/tools/transducer_utils.py:17-757.

Per utterance and encoder frame: (1) hypotheses are sorted by length and the score of every hypothesis that is a
prefix of a longer one is merged into the longer one (prefix search, limited to `prefix_alpha` extra tokens);
(2) up to `max_num_expansions_per_step` rounds: each live hypothesis proposes its beam+beta best next symbols, those
within `expansion_gamma` of its best survive, the best beam+beta of all proposals are kept; proposals that end with
blank are set aside for the next frame, the others advance the prediction network and go on; (3) after the last round
the remaining ones pay the blank score and the best `beam` hypotheses overall move to the next frame.  Scores are
ranked normalised by the number of emissions, the final n-best by score per output token.

The search is host logic over small tensors (a handful of hypotheses); the model enters through four callbacks, which
`TransducerBeamSearchDecoder` wires to the native kernels (prediction-network step = embedding + GEMMs, joint =
LayerNorm / add+ReLU / output GEMM) exactly like the greedy decoder."""
import torch
import torch.nn.functional as F

from .transducer_greedy_decoder import LN_EPS, TransducerGreedyDecoder
from .. import ops as _ops


class _Hyps:
    """A set of hypotheses as parallel tensors (batch dim 0): scores, seqs (pad-filled, starts with bos), lens, nemit,
    prev (last proposed symbol), hs / cs (prediction-network state [n, L, H]), dec (its output per sequence position
    [n, U, H]) and, with an LM, lm_scores, lhs / lcs, ldec."""
    PADDED = ("seqs", "dec", "ldec")

    def __init__(self, **f):
        self.f = f

    def __getattr__(self, k):
        try:
            return self.__dict__["f"][k]
        except KeyError:
            raise AttributeError(k)

    def size(self):
        return self.f["scores"].size(0)

    def take(self, index):
        """Rows `index` (a permutation / subset), padding columns beyond the longest kept sequence dropped."""
        lens = self.f["lens"].index_select(0, index)
        width = int(lens.max()) if lens.numel() else 0
        out = {}
        for k, v in self.f.items():
            if v is None:
                out[k] = None
            elif k in self.PADDED:
                out[k] = v[:, :width].index_select(0, index)
            else:
                out[k] = v.index_select(0, index)
        return _Hyps(**out)

    def where(self, mask):
        return self if bool(mask.all()) else self.take(mask.nonzero().squeeze(1))

    def repeat(self, k):
        return _Hyps(**{n: (None if v is None else v.repeat_interleave(k, dim=0)) for n, v in self.f.items()})

    @staticmethod
    def cat(a, b, pad):
        if b.size() == 0:
            return a
        if a.size() == 0:
            return b
        out = {}
        for k, v in a.f.items():
            w = b.f[k]
            if v is None:
                out[k] = None
                continue
            if k in _Hyps.PADDED:
                width = max(v.size(1), w.size(1))
                fill = pad if k == "seqs" else 0
                grow = lambda t: t if t.size(1) == width else F.pad(t, ((0, width - t.size(1)) if t.dim() == 2 else (0, 0, 0, width - t.size(1))), value=fill)  # noqa: E731
                v, w = grow(v), grow(w)
            out[k] = torch.cat((v, w), dim=0)
        return _Hyps(**out)

    def ranked(self, normalize):
        return self.f["scores"] / self.f["nemit"] if normalize else self.f["scores"]

    def sort_by_score(self, normalize, descending=True):
        return self if self.size() == 0 else self.take(self.ranked(normalize).argsort(descending=descending))

    def top_k(self, k, normalize):
        if k > self.size():
            return self.sort_by_score(normalize)
        return self.take(torch.topk(self.ranked(normalize), k, largest=True, sorted=True)[1])

    def last(self, key):
        v = self.f[key]
        idx = (self.f["lens"] - 1).view(-1, 1, 1).expand(-1, 1, v.size(2))
        return v.gather(1, idx).squeeze(1)

    def set_last_(self, key, value):
        v = self.f[key]
        idx = (self.f["lens"] - 1).view(-1, 1, 1).expand(-1, 1, v.size(2))
        v.scatter_(1, idx, value.unsqueeze(1).to(v.dtype))

    def append_(self, tokens, blank, pad):
        """Record the proposed symbols: non-blank ones extend their sequence (Hypotheses.append_tokens_)."""
        self.f["prev"] = tokens.clone()
        is_blank = tokens == blank
        if bool(is_blank.all()):
            return  # (the reference leaves the emission counters untouched in this case)
        lens = self.f["lens"]
        if bool((tokens[lens == lens.max()] != blank).any()):
            self.f["seqs"] = F.pad(self.f["seqs"], (0, 1), value=pad)
            for k in ("dec", "ldec"):
                if self.f.get(k) is not None:
                    self.f[k] = F.pad(self.f[k], (0, 0, 0, 1))
        self.f["seqs"].scatter_(1, lens.unsqueeze(1), tokens.masked_fill(is_blank, pad).unsqueeze(1))
        self.f["lens"] = lens + (~is_blank).long()
        self.f["nemit"] = self.f["nemit"] + 1


class AdaptiveExpansionSearch:
    """Model-agnostic search.  Callbacks:
         pred_step(prev [n] int64, hs, cs) -> (out [n, H], hs, cs)          prediction network, one step
         joint_lprobs(frame, out [n, H]) -> fp32 [n, V]                     log-probs of the joint at encoder frame `frame`
         lm_step(prev [n], lhs, lcs) -> (feat [n, H'], lhs, lcs)            LM, one step (optional)
         lm_lprobs(feat [n, H']) -> fp32 [n, V_lm]                          LM log-probs (optional)
       `init_state(n)` / `lm_init_state(n)` return zero states (hs, cs) with batch dim 0."""

    def __init__(self, vocab_size, blank, pad, eos, bos, beam_size, max_num_expansions_per_step=2, expansion_beta=0,
                 expansion_gamma=None, prefix_alpha=None, normalize_scores=True, model_predicts_eos=False, lm_weight=1.0,
                 no_blank_in_lm=False):
        self.V, self.blank, self.pad, self.eos, self.bos = vocab_size, blank, pad, eos, bos
        self.beam = min(beam_size, vocab_size - (1 if pad != blank else 0))
        self.E, self.beta, self.gamma, self.alpha = max_num_expansions_per_step, expansion_beta, expansion_gamma, prefix_alpha
        assert self.E > 0 and expansion_beta >= 0 and (expansion_gamma is None or expansion_gamma > 0.0)
        assert prefix_alpha is None or prefix_alpha > 0
        self.normalize, self.predicts_eos = normalize_scores, model_predicts_eos
        self.lm_weight, self.no_blank_in_lm = lm_weight, no_blank_in_lm

    # ---- LM fusion over the non-blank symbols, transducer's blank / non-blank split preserved -------------------
    def _fuse(self, lp, lm_lp):
        nb = self._nonblank.to(lp.device)
        if not self.no_blank_in_lm:
            lm_lp = lm_lp[:, nb]
        lp_nb = lp[:, nb]
        fused = lp_nb + self.lm_weight * lm_lp
        scale = lp_nb.exp().sum(1).log() - fused.exp().sum(1).log()
        lp = lp.clone()
        lp[:, nb] = fused + scale[:, None]
        padded = torch.cat((lm_lp[:, : self.blank], lm_lp.new_zeros(lm_lp.size(0), 1), lm_lp[:, self.blank:]), dim=1)
        return lp, padded, scale

    def _lm_token(self, tok):
        return torch.where(tok > self.blank, tok - 1, tok) if self.no_blank_in_lm else tok

    def search(self, n_frames, cb, device, bos_token=None, use_lm=False):
        """One utterance.  Returns (sequences int64 [n, U] without the leading bos, scores [n]) sorted best first."""
        self._nonblank = torch.ones(self.V, dtype=torch.bool)
        self._nonblank[self.blank] = False
        prev = torch.full((1,), self.bos if bos_token is None else bos_token, dtype=torch.long, device=device)
        hs, cs = cb["init_state"](1)
        out, hs, cs = cb["pred_step"](prev, hs, cs)
        f = dict(scores=torch.zeros(1, device=device), seqs=prev.unsqueeze(1).clone(), lens=torch.ones(1, dtype=torch.long, device=device),
                 nemit=torch.zeros(1, dtype=torch.long, device=device), prev=prev, hs=hs, cs=cs, dec=out.unsqueeze(1).clone(),
                 lm_scores=None, lhs=None, lcs=None, ldec=None)
        if use_lm:
            lhs, lcs = cb["lm_init_state"](1)
            lfeat, lhs, lcs = cb["lm_step"](self._lm_token(prev), lhs, lcs)
            f.update(lm_scores=torch.zeros(1, device=device), lhs=lhs, lcs=lcs, ldec=lfeat.unsqueeze(1).clone())
        nxt = _Hyps(**f)
        for t in range(n_frames):
            nxt = nxt.take(nxt.lens.argsort(descending=True))
            hyps = self._prefix_merge(nxt, t, cb, use_lm)
            set_aside = None
            for e in range(self.E):
                lp = cb["joint_lprobs"](t, hyps.last("dec"))
                lm_padded = None
                if use_lm:
                    lp, lm_padded, _ = self._fuse(lp, cb["lm_lprobs"](hyps.last("ldec")))
                if self.predicts_eos:
                    lp[:, self.blank] = torch.logaddexp(lp[:, self.blank], lp[:, self.eos])
                    lp[:, self.eos] = float("-inf")
                cand = self._expand(hyps, lp, lm_padded)
                is_blank = cand.prev == self.blank
                blanks = cand.where(is_blank)
                set_aside = blanks if e == 0 else _Hyps.cat(set_aside, blanks, self.pad)
                live = cand.where(~is_blank)
                if live.size() == 0:
                    nxt = set_aside.top_k(self.beam, self.normalize)
                    break
                out, live.f["hs"], live.f["cs"] = cb["pred_step"](live.prev, live.hs, live.cs)
                live.set_last_("dec", out)
                if use_lm:
                    lfeat, live.f["lhs"], live.f["lcs"] = cb["lm_step"](self._lm_token(live.prev), live.lhs, live.lcs)
                    live.set_last_("ldec", lfeat)
                if e < self.E - 1:
                    hyps = live
                else:  # out of expansion rounds: the survivors take the blank of this frame
                    lp = cb["joint_lprobs"](t, out)
                    live.f["scores"] = live.scores + lp[:, self.blank]
                    live.f["prev"] = torch.full_like(live.prev, self.blank)
                    live.f["nemit"] = live.nemit + 1
                    nxt = _Hyps.cat(set_aside, live, self.pad).top_k(self.beam, self.normalize)
        nxt.f["scores"] = nxt.scores / (nxt.lens - 1)
        nxt = nxt.sort_by_score(False)
        return nxt.seqs[:, 1:], nxt.scores

    def _expand(self, hyps, lp, lm_padded):
        """Each hypothesis proposes its k best symbols; prune by value; keep the k best proposals (select_k_expansions)."""
        tot = lp + hyps.scores.unsqueeze(-1)
        k = min(self.beam + self.beta, tot.size(1) - (1 if self.pad != self.blank else 0))
        scores, idx = torch.topk(tot, k=k)
        cand = hyps.repeat(k)
        cand.f["scores"] = scores.reshape(-1)
        if lm_padded is not None:
            cand.f["lm_scores"] = cand.lm_scores + lm_padded.gather(1, idx).reshape(-1)
        cand.append_(idx.reshape(-1), self.blank, self.pad)
        if self.gamma is not None:
            keep = scores >= (scores[:, :1] - self.gamma)
            if not bool(keep.all()):
                cand = cand.where(keep.reshape(-1))
        return cand.top_k(k, self.normalize)

    def _prefix_merge(self, hyps, t, cb, use_lm):
        """hyps sorted by non-increasing length.  If hypothesis i is a prefix of the longer j (at most alpha tokens
        longer), the probability of reaching j's sequence from i within this frame is added to j's score."""
        n = hyps.size()
        lens = hyps.lens
        seqs = hyps.seqs
        rel = torch.zeros(n, n, dtype=torch.bool)
        for j in range(n - 1):
            for i in range(j + 1, n):
                li = int(lens[i])
                rel[i, j] = bool(lens[i] < lens[j]) and bool((seqs[i, :li] == seqs[j, :li]).all())
        if self.alpha is not None:
            rel = rel & (lens.cpu().unsqueeze(1) + self.alpha >= lens.cpu().unsqueeze(0))
        if not bool(rel.any()):
            return hyps
        for j in range(n - 1):
            for i in range(j + 1, n):
                if not bool(rel[i, j]):
                    continue
                li, lj = int(lens[i]), int(lens[j])
                # positions li-1 (in i's history) then li .. lj-2 (in j's): each emits j's next token
                steps = [(i, li - 1)] + [(j, k) for k in range(li, lj - 1)]
                score = hyps.scores[i].clone()
                lm_score = hyps.lm_scores[i].clone() if use_lm else None
                for (row, pos) in steps:
                    lp = cb["joint_lprobs"](t, hyps.dec[row: row + 1, pos])[0]
                    tok = int(seqs[j, pos + 1])
                    score = score + lp[tok]
                    if use_lm:
                        lm_lp = cb["lm_lprobs"](hyps.ldec[row: row + 1, pos])[0]
                        lt = tok - 1 if (self.no_blank_in_lm and tok > self.blank) else tok
                        lm_score = lm_score + lm_lp[lt]
                        _, _, scale = self._fuse(lp.unsqueeze(0), lm_lp.unsqueeze(0))
                        score = score + self.lm_weight * lm_lp[lt] + scale[0]
                hyps.f["scores"][j] = torch.logaddexp(hyps.scores[j], score)
                if use_lm:
                    hyps.f["lm_scores"][j] = torch.logaddexp(hyps.lm_scores[j], lm_score)
        return hyps


class TransducerBeamSearchDecoder(TransducerGreedyDecoder):
    """The reference's constructor arguments (transducer_beam_search_decoder.py:22-42); decode() returns the 1-best like
    the reference (tokens [B, U] padded, scores [B], None), generate() the n-best lists."""

    def __init__(self, models, dictionary, beam_size=1, max_len=0, max_num_expansions_per_step=2, expansion_beta=0,
                 expansion_gamma=None, prefix_alpha=None, normalize_scores=True, temperature=1.0, eos=None, bos=None, blank=None,
                 pad=None, model_predicts_eos=False, symbols_to_strip_from_output=None, lm_model=None, lm_weight=1.0, **unused):
        super().__init__(models, dictionary, max_len=max_len, max_num_expansions_per_step=max_num_expansions_per_step,
                         temperature=temperature, eos=eos, bos=bos, blank=blank, model_predicts_eos=model_predicts_eos,
                         symbols_to_strip_from_output=symbols_to_strip_from_output, lm_model=lm_model, lm_weight=lm_weight)
        if pad is not None:
            self.pad = pad
        self.core = AdaptiveExpansionSearch(self.vocab_size, self.blank, self.pad, self.eos, self.bos, beam_size,
                                            max_num_expansions_per_step, expansion_beta, expansion_gamma, prefix_alpha,
                                            normalize_scores, model_predicts_eos, lm_weight, self.no_blank_in_lm)

    @torch.no_grad()
    def _generate(self, sample, bos_token=None):
        m = self.model
        m.eval()
        P = m.flat.param
        ni = sample["net_input"]
        src, src_len = ni["src_tokens"], ni["src_lengths"]
        if src.dim() == 2:
            src, src_len = m.frontend(src, src_len, None, None)
        enc_out = m.encoder(src, src_len, src_lengths_cpu=ni.get("src_lengths_cpu"))
        enc = enc_out["b200_out"]
        enc_lens = enc_out["src_lengths"][0].cpu().tolist()
        B, T, d = enc.shape
        dev = enc.device
        V = self.vocab_size
        J = P("proj_encoder.weight").shape[0]
        pe_all, _, _ = _ops.layer_norm_fwd(_ops.linear(enc.reshape(B * T, d), P("proj_encoder.weight"), P("proj_encoder.bias")),
                                           P("laynorm_proj_encoder.weight"), P("laynorm_proj_encoder.bias"), LN_EPS)
        pe_all = pe_all.view(B, T, J)
        g, v = P("fc_out.weight_g").float(), P("fc_out.weight_v").float()
        W = (g * v / v.norm(dim=1, keepdim=True)).to(torch.bfloat16).contiguous()
        ldV = (V + 7) // 8 * 8
        nl = len(m.decoder.layers)
        hid = m.decoder.layers[0].weight_hh.shape[1]
        lm = self.lm_model
        results = []
        for b in range(B):
            def pred_step(prev, hs, cs):
                top, nh, nc = self._predictor_step(prev, [hs[:, i] for i in range(nl)], [cs[:, i] for i in range(nl)])
                return top, torch.stack(nh, dim=1), torch.stack(nc, dim=1)

            def joint_lprobs(t, out, _b=b):
                n = out.size(0)
                pd, _, _ = _ops.layer_norm_fwd(_ops.linear(out.contiguous(), P("proj_decoder.weight"), P("proj_decoder.bias")),
                                               P("laynorm_proj_decoder.weight"), P("laynorm_proj_decoder.bias"), LN_EPS)
                fj = _ops.joint_fwd(pe_all[_b, t].contiguous().view(1, 1, J), pd.view(1, n, J)).view(n, J)
                logits = torch.zeros(n, ldV, dtype=torch.bfloat16, device=dev) if ldV != V else \
                    torch.empty(n, ldV, dtype=torch.bfloat16, device=dev)
                _ops.gemm(fj, W, logits, n, V, J, J, J, ldV, bias=P("fc_out.bias"))
                return torch.log_softmax(logits[:, :V].float() / self.temperature, dim=-1)

            cb = dict(pred_step=pred_step, joint_lprobs=joint_lprobs,
                      init_state=lambda n: (torch.zeros(n, nl, hid, dtype=torch.bfloat16, device=dev),
                                            torch.zeros(n, nl, hid, dtype=torch.float32, device=dev)))
            if lm is not None:
                lm.eval()
                ld = lm.decoder
                lnl, lh = len(ld.layers), ld.hidden_size
                w = ld.embed_tokens.weight

                def lm_step(prev, lhs, lcs):
                    y, nh, nc, _ = ld.step(ld.embed_tokens(prev), [lhs[:, i] for i in range(lnl)], [lcs[:, i] for i in range(lnl)], None)
                    return y, torch.stack(nh, dim=1), torch.stack(nc, dim=1)

                cb.update(lm_step=lm_step, lm_lprobs=lambda feat: torch.log_softmax(ld.output_layer(feat.to(w.dtype)).float(), dim=-1),
                          lm_init_state=lambda n: (w.new_zeros(n, lnl, lh), w.new_zeros(n, lnl, lh)))
            n_frames = min(enc_lens[b], self.max_len) if self.max_len > 0 else enc_lens[b]
            results.append(self.core.search(n_frames, cb, dev, bos_token=bos_token, use_lm=lm is not None))
        return results

    @torch.no_grad()
    def decode(self, models, sample, bos_token=None, **unused):
        res = self._generate(sample, bos_token)
        width = max(max(int((s[0] != self.pad).sum()) for s, _ in res), 1)
        tokens = torch.full((len(res), width), self.pad, dtype=torch.long, device=res[0][0].device)
        for b, (s, _) in enumerate(res):
            row = s[0][s[0] != self.pad]
            tokens[b, : row.numel()] = row
        return tokens, torch.stack([sc[0] for _, sc in res]), None

    def generate(self, models, sample, bos_token=None, **unused):
        strip = sorted(self.symbols_to_strip_from_output | {self.pad})
        out = []
        for seqs, scores in self._generate(sample, bos_token):
            st = torch.tensor(strip, device=seqs.device)
            out.append([{"tokens": seqs[k][~torch.isin(seqs[k], st)], "score": scores[k], "attention": None, "alignment": None,
                         "positional_scores": None} for k in range(seqs.size(0))])
        return out
