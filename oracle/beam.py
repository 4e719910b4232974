"""Oracle: batched beam search as the reference runs it, restated in plain Python/PyTorch on CPU.
TEST INFRASTRUCTURE ONLY.  Follows fairseq/sequence_generator.py:212-621 (incl. finalize_hypos :657-766 and the
batch compaction :507-541 that the product omits) and fairseq/search.py:103-144, with a step function
`lprobs_fn(step, tokens[:, :step+1], reorder_state) -> lprobs [rows, V]` standing in for EnsembleModel.forward_decoder
(+ LM fusion).  Pinned twice: by the reference's own known-answer tests (tests/test_sequence_generator.py:202-283), which
tests/test_beam_search.py replays against this function and against the product generator, and by the REAL reference
SequenceGenerator run with LM shallow fusion / eos_factor on table models (oracle/pin_against_reference.py::pin_beam ->
tests/golden/beam_reference.npz).
"""
import math

import torch


def topk_lex(x, k):
    """torch.topk made deterministic: (value desc, index asc)."""
    vals, idx = [], []
    for row in x:
        order = sorted(range(row.numel()), key=lambda j: (-float(row[j]), j))[:k]
        idx.append(order)
        vals.append([float(row[j]) for j in order])
    return torch.tensor(vals, dtype=x.dtype), torch.tensor(idx, dtype=torch.long)


def generate(lprobs_fn, bsz, src_len, V, pad, unk, eos, beam_size=1, max_len_a=0, max_len_b=200, model_max_len=100, min_len=1,
             normalize_scores=True, len_penalty=1.0, unk_penalty=0.0, eos_factor=None):
    beam = min(beam_size, V - 1)
    max_len = min(int(max_len_a * src_len + max_len_b), model_max_len - 1)
    NEG = -math.inf
    scores = torch.zeros(bsz * beam, max_len + 1)
    tokens = torch.full((bsz * beam, max_len + 2), pad, dtype=torch.long)
    tokens[:, 0] = eos
    cands_to_ignore = torch.zeros(bsz, beam, dtype=torch.bool)
    finalized = [[] for _ in range(bsz)]
    finished = [False] * bsz
    num_remaining = bsz
    cand_size = 2 * beam
    bbsz_offsets = (torch.arange(bsz) * beam).unsqueeze(1)
    cand_offsets = torch.arange(cand_size)
    reorder_state, batch_idxs = None, None
    for step in range(max_len + 1):
        if reorder_state is not None and batch_idxs is not None:
            corr = batch_idxs - torch.arange(batch_idxs.numel())
            reorder_state.view(-1, beam).add_(corr.unsqueeze(-1) * beam)
        lprobs = lprobs_fn(step, tokens[:, : step + 1], reorder_state).clone().float()
        lprobs[lprobs != lprobs] = NEG
        lprobs[:, pad] = NEG
        lprobs[:, unk] -= unk_penalty
        if step >= max_len:
            lprobs[:, :eos] = NEG
            lprobs[:, eos + 1:] = NEG
        elif eos_factor is not None:
            dis = lprobs[:, eos] < eos_factor * lprobs.max(dim=1)[0]
            lprobs[dis, eos] = NEG
        if step < min_len:
            lprobs[:, eos] = NEG
        # BeamSearch.step
        lp = lprobs.view(bsz, -1, V)
        if step == 0:
            lp = lp[:, ::beam, :].contiguous()
        else:
            lp = lp + scores.view(bsz, beam, -1)[:, :, step - 1].unsqueeze(-1)
        cand_scores, flat = topk_lex(lp.view(bsz, -1), min(beam * 2, lp.view(bsz, -1).size(1) - 1))
        cand_beams, cand_indices = flat // V, flat % V
        cand_bbsz_idx = cand_beams + bbsz_offsets
        eos_mask = cand_indices.eq(eos) & cand_scores.ne(NEG)
        eos_mask[:, :beam][cands_to_ignore] = False
        eos_bbsz_idx = torch.masked_select(cand_bbsz_idx[:, :beam], eos_mask[:, :beam])
        finalized_sents = []
        if eos_bbsz_idx.numel() > 0:
            eos_scores = torch.masked_select(cand_scores[:, :beam], eos_mask[:, :beam])
            # ---- finalize_hypos
            tok_clone = tokens.index_select(0, eos_bbsz_idx)[:, 1: step + 2].clone()
            tok_clone[:, step] = eos
            pos = scores.index_select(0, eos_bbsz_idx)[:, : step + 1].clone()
            pos[:, step] = eos_scores
            pos[:, 1:] = pos[:, 1:] - pos[:, :-1]
            if normalize_scores:
                eos_scores = eos_scores / (step + 1) ** len_penalty
            cum_unfin, prev = [], 0
            for f in finished:
                if f:
                    prev += 1
                else:
                    cum_unfin.append(prev)
            unfin_idx = eos_bbsz_idx // beam
            sent = unfin_idx + torch.tensor(cum_unfin)[unfin_idx]
            seen = sorted(set(zip(sent.tolist(), unfin_idx.tolist())))
            for i, sidx in enumerate(sent.tolist()):
                if len(finalized[sidx]) < beam:
                    finalized[sidx].append({"tokens": tok_clone[i], "score": eos_scores[i], "positional_scores": pos[i]})
            for us, uidx in seen:
                if not finished[us] and (len(finalized[us]) == beam or step == max_len):
                    finished[us] = True
                    finalized_sents.append(uidx)
            num_remaining -= len(finalized_sents)
        if num_remaining == 0:
            break
        assert step < max_len
        if finalized_sents:
            new_bsz = bsz - len(finalized_sents)
            batch_mask = torch.ones(bsz, dtype=torch.bool)
            batch_mask[finalized_sents] = False
            batch_idxs = torch.arange(bsz).masked_select(batch_mask)
            eos_mask, cand_beams = eos_mask[batch_idxs], cand_beams[batch_idxs]
            bbsz_offsets = bbsz_offsets[:new_bsz]
            cand_bbsz_idx = cand_beams + bbsz_offsets
            cand_scores, cand_indices = cand_scores[batch_idxs], cand_indices[batch_idxs]
            cands_to_ignore = cands_to_ignore[batch_idxs]
            scores = scores.view(bsz, -1)[batch_idxs].view(new_bsz * beam, -1)
            tokens = tokens.view(bsz, -1)[batch_idxs].view(new_bsz * beam, -1)
            bsz = new_bsz
        else:
            batch_idxs = None
        eos_mask[:, :beam] = ~((~cands_to_ignore) & (~eos_mask[:, :beam]))
        active_mask = eos_mask.long() * cand_size + cand_offsets[: eos_mask.size(1)]
        new_ignore, active_hypos = torch.topk(active_mask, k=beam, dim=1, largest=False)
        cands_to_ignore = new_ignore.ge(cand_size)[:, :beam]
        active_bbsz_idx = torch.gather(cand_bbsz_idx, 1, active_hypos).view(-1)
        tokens[:, : step + 1] = torch.index_select(tokens[:, : step + 1], 0, active_bbsz_idx)
        tokens.view(bsz, beam, -1)[:, :, step + 1] = torch.gather(cand_indices, 1, active_hypos)
        if step > 0:
            scores[:, :step] = torch.index_select(scores[:, :step], 0, active_bbsz_idx)
        scores.view(bsz, beam, -1)[:, :, step] = torch.gather(cand_scores, 1, active_hypos)
        reorder_state = active_bbsz_idx
    out = []
    for hyps in finalized:
        order = sorted(range(len(hyps)), key=lambda i: -float(hyps[i]["score"]))
        out.append([hyps[i] for i in order])
    return out
