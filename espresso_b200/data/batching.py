"""Length-bucketed batching and rank sharding with the reference's semantics (host logic).

  ordered_indices   espresso/data/asr_dataset.py:392-408   (sort by target then source length, stable)
  batch_by_size     fairseq/data/data_utils.py:282-365 -> data_utils_fast.pyx:20-105 (greedy packing of
                    consecutive indices under max_tokens = max_len * n_sentences and max_sentences)
  shard_batches     fairseq/data/iterators.py:623-658 (ShardedIterator: rank r takes batches r, r+W, ...;
                    the tail is padded with empty batches so every rank joins every collective)
"""
import numpy as np


def ordered_indices(src_sizes, tgt_sizes=None, shuffle_seed=None):
    n = len(src_sizes)
    idx = np.arange(n)
    if shuffle_seed is not None:
        idx = np.random.RandomState(shuffle_seed).permutation(n)
    if tgt_sizes is not None:
        idx = idx[np.argsort(np.asarray(tgt_sizes)[idx], kind="mergesort")]
    return idx[np.argsort(np.asarray(src_sizes)[idx], kind="mergesort")]


def batch_by_size(indices, num_tokens, max_tokens=None, max_sentences=None, required_batch_size_multiple=1):
    """Greedy packing identical to fairseq's batch_by_size_vec: a batch closes when adding the next sample
    would exceed max_tokens (= longest sample * batch size) or max_sentences."""
    max_tokens = int(max_tokens) if max_tokens else -1
    max_sentences = int(max_sentences) if max_sentences else -1
    mult = required_batch_size_multiple
    batches, cur, cur_max = [], [], 0
    for i in indices:
        nt = int(num_tokens[i])
        assert max_tokens <= 0 or nt <= max_tokens, "sentence at index %d exceeds max_tokens limit" % i
        new_max = max(cur_max, nt)
        n_after = len(cur) + 1
        overflow = (max_sentences > 0 and len(cur) == max_sentences) or (max_tokens > 0 and n_after * new_max > max_tokens)
        if overflow and cur:
            keep = max(mult * (len(cur) // mult), len(cur) % mult)
            batches.append(np.array(cur[:keep], dtype=np.int64))
            cur = cur[keep:]
            cur_max = max([int(num_tokens[j]) for j in cur], default=0)
            new_max = max(cur_max, nt)
        cur.append(int(i))
        cur_max = new_max
    if cur:
        batches.append(np.array(cur, dtype=np.int64))
    return batches


def shard_batches(batches, world_size, rank, fill_value=None):
    """ShardedIterator semantics: every rank gets ceil(len/W) entries; missing ones are `fill_value`."""
    n = (len(batches) + world_size - 1) // world_size
    out = []
    for i in range(n):
        j = i * world_size + rank
        out.append(batches[j] if j < len(batches) else (fill_value if fill_value is not None else np.array([], dtype=np.int64)))
    return out
