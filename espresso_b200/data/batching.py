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
    """fairseq.data.data_utils.batch_by_size (fairseq/data/data_utils.py:282-365): pack the samples `indices` (in that
    order) into batches under max_tokens (= longest sample * batch size), max_sentences and a batch-size multiple.
    `num_tokens[i]` is the size of sample i.  The packing itself is native (esp_batch_by_size in the C ABI, the
    counterpart of the reference's Cython batch_by_size_vec); returns a list of int64 index arrays."""
    import ctypes

    from .. import lib as _lib

    indices = np.ascontiguousarray(np.asarray(indices, dtype=np.int64))
    n = int(indices.shape[0])
    if n == 0:
        return []
    sizes = np.ascontiguousarray(np.asarray(num_tokens, dtype=np.int64)[indices])
    ends = np.zeros(n, dtype=np.int32)
    k = _lib.load().esp_batch_by_size(sizes.ctypes.data_as(ctypes.c_void_p), n, int(max_tokens) if max_tokens else -1,
                                      int(max_sentences) if max_sentences else -1, int(required_batch_size_multiple),
                                      ends.ctypes.data_as(ctypes.c_void_p))
    if k < 0:
        raise AssertionError(_lib.load().esp_last_error().decode())
    return np.split(indices, ends[:k])


def shard_batches(batches, world_size, rank, fill_value=None):
    """ShardedIterator semantics: every rank gets ceil(len/W) entries; missing ones are `fill_value`."""
    n = (len(batches) + world_size - 1) // world_size
    out = []
    for i in range(n):
        j = i * world_size + rank
        out.append(batches[j] if j < len(batches) else (fill_value if fill_value is not None else np.array([], dtype=np.int64)))
    return out
