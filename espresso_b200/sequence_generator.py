"""Batched beam search with LM shallow fusion, B200-native mirror of `SequenceGenerator`
(fairseq/sequence_generator.py:27-621 with the Espresso patches: lm_model / lm_weight :385-393, eos_factor
:404-410) and `BeamSearch.step` (fairseq/search.py:103-144).

generate(models, sample) -> List[bsz] of List[<= beam] of {"tokens" (ends with eos), "score", "attention",
"alignment", "positional_scores"}, sorted by score -- the reference's contract (:744-752).

The per-step tensor work is three native kernels (esp_beam_merge, esp_beam_topk, esp_beam_bookkeep) plus
esp_gather_rows for the incremental state; hypothesis bookkeeping and finalisation live in device buffers, and the
only host<->device traffic per step is one 4-byte "how many sentences are unfinished" read.  Finished sentences are
kept in the batch (static shapes) instead of being compacted away (:507-541): results are identical.

Model protocol (implemented by espresso_b200 models and by test doubles):
    enc = model.forward_encoder(net_input)                       # any object
    model.max_decoder_positions() -> int
    state = model.init_incremental_state(enc, bsz, beam)          # encoder state replicated per beam if needed
    values, is_logits = model.decode_step(step, tokens, state, new_order)
        tokens: int32 [bsz*beam, max_len+2] device buffer (columns 0..step are valid);
        new_order: int32 [bsz*beam] source row of each hypothesis since the previous step (None at step 0);
        values: [bsz*beam, >= V] bf16/fp32 logits (is_logits) or fp32 log-probabilities.
"""
import math

import torch

from . import ops as _ops


class _SearchState:
    pass


class SequenceGenerator:
    def __init__(self, models, tgt_dict, beam_size=1, max_len_a=0, max_len_b=200, max_len=0, min_len=1,
                 normalize_scores=True, len_penalty=1.0, unk_penalty=0.0, temperature=1.0, lm_model=None, lm_weight=1.0,
                 eos_factor=None, eos=None, use_cuda_graphs=True, **unused):
        self.models = models if isinstance(models, (list, tuple)) else [models]
        assert len(self.models) >= 1
        self.tgt_dict = tgt_dict
        self.pad, self.unk = tgt_dict.pad(), tgt_dict.unk()
        self.eos = tgt_dict.eos() if eos is None else eos
        self.vocab_size = len(tgt_dict)
        self.beam_size = min(beam_size, self.vocab_size - 1)
        self.max_len_a, self.max_len_b, self.min_len = max_len_a, max_len_b, min_len
        self.max_len = max_len or self.models[0].max_decoder_positions()
        self.normalize_scores, self.len_penalty, self.unk_penalty = normalize_scores, len_penalty, unk_penalty
        self.temperature = temperature
        assert temperature > 0, "--temperature must be greater than 0"
        self.lm_model, self.lm_weight = lm_model, lm_weight
        self.eos_factor = eos_factor
        assert eos_factor is None or eos_factor >= 1.0, "--eos-factor must be >= 1.0 if set"
        # One search step (decoder + LM step, merge, top-k, bookkeeping: ~140 kernel launches) is host-bound when launched
        # from Python.  With use_cuda_graphs every step index gets its own CUDA graph over PERSISTENT search / decoder state
        # (first batch of a shape: eager, second: capture, then replay); the per-step "all sentences finished?" read stays
        # outside the graphs.  Single-model search only (ensembles run eagerly).
        self.use_cuda_graphs = use_cuda_graphs
        self._graph_cache = {}

    @torch.no_grad()
    def generate(self, models, sample, **kwargs):
        return self._generate(sample, **kwargs)

    def forward(self, sample, **kwargs):
        return self._generate(sample, **kwargs)

    @torch.no_grad()
    def _generate(self, sample, bos_token=None, **unused):
        model = self.models[0]
        net_input = sample["net_input"]
        src_tokens = net_input["src_tokens"]
        dev = src_tokens.device
        bsz, src_len = src_tokens.shape[:2]
        if src_tokens.dim() == 2 and src_tokens.is_floating_point():
            # raw waveform batch [B, samples] (on-device front end): the reference's T_src is the padded number of
            # feature FRAMES (src_tokens.size(1) of [B, T, F], sequence_generator.py:270,285-288)
            src_len = 1 + (src_len - 400) // 160 if src_len >= 400 else 0
        beam, V = self.beam_size, self.vocab_size
        max_len = min(int(self.max_len_a * src_len + self.max_len_b), self.max_len - 1)  # :285-288
        assert self.min_len <= max_len, "min_len cannot be larger than max_len, please adjust these!"
        N, L = bsz * beam, max_len + 2

        for m_ in tuple(self.models) + (self.lm_model,):
            if m_ is not None and hasattr(m_, "t_max_hint"):
                m_.t_max_hint = max_len + 1
        graphs_on = (self.use_cuda_graphs and dev.type == "cuda" and len(self.models) == 1
                     and hasattr(model, "advance_state_without_compute")
                     and (self.lm_model is None or hasattr(self.lm_model, "advance_state_without_compute")))
        enc = model.forward_encoder(net_input)
        cache = None
        if graphs_on:
            enc_T = enc["b200_out"].shape[1] if isinstance(enc, dict) and "b200_out" in enc else -1
            has_pad = isinstance(enc, dict) and len(enc.get("encoder_padding_mask", [])) > 0
            key = (bsz, beam, max_len, enc_T, has_pad, bos_token, str(dev))
            cache = self._graph_cache.get(key)
            if cache is None:
                if len(self._graph_cache) >= 4:  # bounded: search shapes come in a handful of buckets
                    self._graph_cache.pop(next(iter(self._graph_cache)))
                cache = self._graph_cache[key] = {"calls": 0, "graphs": {}, "pool": None, "state": None, "lm_state": None,
                                                  "search": None}
        # an ensemble = one encoder pass and one incremental state per model; their step log-probs are averaged in the
        # probability domain (EnsembleModel.forward_decoder, fairseq/sequence_generator.py:836-901)
        if graphs_on:
            states = [model.init_incremental_state(enc, bsz, beam, reuse=cache["state"])]
            cache["state"] = states[0]
        else:
            states = [model.init_incremental_state(enc, bsz, beam)]
            for m_ in self.models[1:]:
                states.append(m_.init_incremental_state(m_.forward_encoder(net_input), bsz, beam))
        state = states[0]
        lm_state = None
        if self.lm_model is not None:
            if graphs_on:
                lm_state = cache["lm_state"] = self.lm_model.init_incremental_state(None, bsz, beam, reuse=cache["lm_state"])
            else:
                lm_state = self.lm_model.init_incremental_state(None, bsz, beam)
        for s_ in states + [lm_state]:  # models may size their caches from this
            if isinstance(s_, dict):
                s_["max_len"] = max_len

        st = cache["search"] if graphs_on else None
        if st is None:
            st = _SearchState()
            st.buf_tokens = [torch.empty((N, L), dtype=torch.int32, device=dev) for _ in range(2)]
            st.buf_scores = [torch.empty(N, L - 1, dtype=torch.float32, device=dev) for _ in range(2)]
            st.ignore = torch.empty(N, dtype=torch.uint8, device=dev)
            st.finished = torch.empty(bsz, dtype=torch.uint8, device=dev)
            st.nfin = torch.empty(bsz, dtype=torch.int32, device=dev)
            st.fin_tokens = torch.empty((bsz, beam, L - 1), dtype=torch.int32, device=dev)
            st.fin_len = torch.empty(bsz, beam, dtype=torch.int32, device=dev)
            st.fin_score = torch.empty(bsz, beam, dtype=torch.float32, device=dev)
            st.fin_pos = torch.empty(bsz, beam, L - 1, dtype=torch.float32, device=dev)
            st.new_order = torch.empty(N, dtype=torch.int32, device=dev)
            st.n_unfinished = torch.empty((1,), dtype=torch.int32, device=dev)
            st.cand = torch.empty(N, V, dtype=torch.float32, device=dev)
            st.prev = torch.empty(N, dtype=torch.float32, device=dev)
            st.arange = torch.arange(N, dtype=torch.int32, device=dev)
            if graphs_on:
                cache["search"] = st
        # (re-)initialise in place: the same buffers, in the same ping-pong roles, for every batch of this shape
        st.tokens, st.tokens_alt = st.buf_tokens
        st.scores, st.scores_alt = st.buf_scores
        st.tokens.fill_(self.pad)
        st.tokens[:, 0] = self.eos if bos_token is None else bos_token
        st.scores.zero_()
        st.scores_alt.zero_()
        st.ignore.zero_()
        st.finished.zero_()
        st.nfin.zero_()
        st.fin_tokens.fill_(self.pad)
        st.fin_len.zero_()
        st.fin_score.zero_()
        st.fin_pos.zero_()
        st.new_order.copy_(st.arange)
        st.n_unfinished.fill_(bsz)
        cand, prev = st.cand, st.prev

        def device_step(step):
            """All device work of search step `step` (this is what a step graph records)."""
            new_order = st.new_order if step > 0 else None
            values, is_logits = model.decode_step(step, st.tokens, state, new_order)
            if len(self.models) > 1:
                lps = []
                for m_, s_ in zip(self.models, states):
                    v_, il_ = (values, is_logits) if m_ is model else m_.decode_step(step, st.tokens, s_, new_order)
                    v_ = v_[:, :V].float()
                    lps.append(torch.log_softmax(v_ / self.temperature, dim=-1) if il_ else v_)
                values = torch.logsumexp(torch.stack(lps, dim=0), dim=0) - math.log(len(self.models))
                is_logits = False  # temperature already applied per model (sequence_generator.py:866-870)
            lm_vals, lm_logits = (None, True)
            if self.lm_model is not None:
                lm_vals, lm_logits = self.lm_model.decode_step(step, st.tokens, lm_state, new_order)
            if step > 0:
                prev.copy_(st.scores[:, step - 1])
            _ops.beam_merge(values, V, is_logits, cand, prev_scores=prev if step > 0 else None,
                            temperature=self.temperature if (is_logits or len(self.models) == 1) else 1.0,
                            lm=lm_vals, lm_is_logits=lm_logits, lm_weight=self.lm_weight, pad=self.pad, unk=self.unk,
                            unk_penalty=self.unk_penalty, eos=self.eos, force_eos=step >= max_len,
                            eos_factor=self.eos_factor, ban_eos=step < self.min_len)
            n_cand = V if step == 0 else beam * V  # step 0: all beams are identical, use the first (search.py:119-122)
            K = min(2 * beam, n_cand - 1)
            cs, ct, cb = _ops.beam_topk(cand, bsz, beam * V, n_cand, K, V)
            _ops.beam_bookkeep(step, max_len, bsz, beam, K, self.eos, self.pad, self.normalize_scores, self.len_penalty,
                               cs, ct, cb, st)

        def host_swaps():
            """What device_step changes on the host side (ping-pong roles); replays must repeat it."""
            st.tokens, st.tokens_alt = st.tokens_alt, st.tokens
            st.scores, st.scores_alt = st.scores_alt, st.scores
            model.advance_state_without_compute(state)
            if self.lm_model is not None:
                self.lm_model.advance_state_without_compute(lm_state)

        # "all sentences finished?" is read one step LATE through pinned memory: the host never waits for the step it has
        # just queued (finished sentences are inert, so at most one surplus step runs after the last hypothesis ended)
        lagged = dev.type == "cuda"
        if lagged:
            if not hasattr(self, "_nu_host"):
                self._nu_host = [torch.zeros(1, dtype=torch.int32).pin_memory() for _ in range(2)]
                self._nu_ev = [torch.cuda.Event() for _ in range(2)]
        for step in range(max_len + 1):  # one extra step for the eos marker
            graph = cache["graphs"].get(step) if graphs_on else None
            if graph is not None:
                graph.replay()
                host_swaps()
            elif graphs_on and cache["calls"] >= 1:
                # second batch of this shape: record the step (recording executes nothing), then run it by replaying
                graph = torch.cuda.CUDAGraph()
                torch.cuda.synchronize()
                with torch.cuda.graph(graph, pool=cache["pool"]):
                    device_step(step)
                if cache["pool"] is None:
                    cache["pool"] = graph.pool()
                cache["graphs"][step] = graph
                graph.replay()
            else:
                device_step(step)
            if lagged:
                k = step & 1
                self._nu_host[k].copy_(st.n_unfinished, non_blocking=True)
                self._nu_ev[k].record()
                if step > 0:
                    self._nu_ev[1 - k].synchronize()
                    if int(self._nu_host[1 - k][0]) == 0:
                        break
            elif int(st.n_unfinished.item()) == 0:
                break
        if graphs_on:
            cache["calls"] += 1

        # ---- collect (:611-620): sort each sentence's hypotheses by score, descending
        nfin = st.nfin.cpu().tolist()
        fl, fs = st.fin_len.cpu(), st.fin_score.cpu()
        ft, fp = st.fin_tokens.cpu(), st.fin_pos.cpu()
        finalized = []
        for s in range(bsz):
            hyps = []
            for k in range(nfin[s]):
                n = int(fl[s, k])
                hyps.append({"tokens": ft[s, k, :n].long(), "score": fs[s, k].clone(), "attention": None, "alignment": None,
                             "positional_scores": fp[s, k, :n].clone()})
            order = sorted(range(len(hyps)), key=lambda i: -float(hyps[i]["score"]))
            finalized.append([hyps[i] for i in order])
        return finalized


class TableDecoderModel:
    """Decoder double driven by per-step probability tables (the reference's TestIncrementalDecoder,
    tests/utils.py:546-603): used to run the reference's known-answer beam-search tests against this generator."""

    def __init__(self, beam_probs, vocab_size, eos, max_positions=100):
        self.beam_probs, self.V, self.eos, self.max_pos = beam_probs, vocab_size, eos, max_positions

    def max_decoder_positions(self):
        return self.max_pos

    def forward_encoder(self, net_input):
        return None

    def init_incremental_state(self, enc, bsz, beam):
        return {"order": None}

    def decode_step(self, step, tokens, state, new_order):
        N = tokens.shape[0]
        probs = torch.zeros(N, self.V, dtype=torch.float32)
        if step < len(self.beam_probs):
            t = self.beam_probs[step]
            probs[:, self.eos:] = t if t.shape[0] == N else t[: N]
        else:
            probs[:, self.eos] = 1.0
        return probs.log().to(tokens.device), False
