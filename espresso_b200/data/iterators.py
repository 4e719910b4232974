"""Epoch iteration with background batch assembly (host side; fairseq/data/iterators.py: EpochBatchIterator :283-600,
BufferedIterator / BackgroundConsumer :661-750).

The reference runs dataset.__getitem__ + collater in DataLoader worker PROCESSES because its items are expensive (fbank on
the CPU).  Here an item is a file read (the front end runs on the GPU), so a few THREADS that read, collate and PIN batches a
bounded number of steps ahead are enough; the consumer gets pinned tensors it can upload with non_blocking=True while the
previous step computes.  Same epoch semantics as the reference: batches are frozen once (batch_by_size over the dataset's
ordered_indices), shuffled per epoch under numpy_seed(seed + epoch), sharded round-robin over ranks with dummy padding
batches so every rank does the same number of updates, resumable from (epoch, iterations_in_epoch)."""
import threading

import numpy as np
import torch

from . import batching
from .specaugment import numpy_seed


def _pin(x):
    if torch.is_tensor(x):
        return x.pin_memory() if torch.cuda.is_available() else x
    if isinstance(x, dict):
        return {k: _pin(v) for k, v in x.items()}
    return x


class PrefetchIterator:
    """Iterate `fn(job)` for job in jobs, computed by `num_workers` threads at most `buffer_size` results ahead, in order."""

    _END = object()

    def __init__(self, jobs, fn, num_workers=2, buffer_size=4):
        self.jobs, self.fn, self.total = list(jobs), fn, len(jobs)
        self.n_workers, self.buffer_size = max(1, num_workers), max(1, buffer_size)
        self._results, self._cv, self._next_job, self._next_out = {}, threading.Condition(), 0, 0
        self._stop, self._err, self._threads = False, None, []

    def __len__(self):
        return self.total

    def _work(self):
        while True:
            with self._cv:
                while True:
                    if self._stop or self._next_job >= self.total:
                        return
                    if self._next_job - self._next_out < self.buffer_size:
                        break
                    self._cv.wait()
                k = self._next_job
                self._next_job += 1
            try:
                r = self.fn(self.jobs[k])
            except BaseException as e:  # surfaced in the consumer, like BackgroundConsumer (:688-692)
                with self._cv:
                    self._err = e
                    self._cv.notify_all()
                return
            with self._cv:
                self._results[k] = r
                self._cv.notify_all()

    def __iter__(self):
        self._threads = [threading.Thread(target=self._work, daemon=True) for _ in range(self.n_workers)]
        for t in self._threads:
            t.start()
        try:
            while self._next_out < self.total:
                with self._cv:
                    while self._next_out not in self._results and self._err is None:
                        self._cv.wait()
                    if self._err is not None:
                        raise self._err
                    r = self._results.pop(self._next_out)
                    self._next_out += 1
                    self._cv.notify_all()
                yield r
        finally:
            with self._cv:
                self._stop = True
                self._cv.notify_all()


class EpochBatchIterator:
    def __init__(self, dataset, batch_sampler, seed=1, num_shards=1, shard_id=0, num_workers=2, buffer_size=4, epoch=1,
                 pin_memory=True):
        self.dataset, self.frozen_batches = dataset, [np.asarray(b) for b in batch_sampler]
        self.seed, self.num_shards, self.shard_id = seed, num_shards, shard_id
        self.num_workers, self.buffer_size, self.pin_memory = num_workers, buffer_size, pin_memory
        self.epoch, self._offset, self._cur = max(epoch, 1), 0, None
        self.shuffle = True

    def __len__(self):
        return -(-len(self.frozen_batches) // self.num_shards)

    @property
    def next_epoch_idx(self):
        """iterators.py:352-361"""
        if self._cur is not None and self.end_of_epoch():
            return self.epoch + 1
        return self.epoch

    def end_of_epoch(self):
        return self._cur is not None and self._cur.n >= len(self)

    @property
    def iterations_in_epoch(self):
        return self._cur.n if self._cur is not None else self._offset

    def next_epoch_itr(self, shuffle=True):
        self.epoch = self.next_epoch_idx
        if hasattr(self.dataset, "set_epoch"):
            self.dataset.set_epoch(self.epoch)
        batches = list(self.frozen_batches)
        if shuffle:
            with numpy_seed(self.seed + self.epoch):        # iterators.py:537-545 shuffle_batches
                np.random.shuffle(batches)
        mine = batching.shard_batches(batches, self.num_shards, self.shard_id, fill_value=[])
        offset, self._offset = self._offset, 0
        self._cur = _Counting(PrefetchIterator(mine[offset:], self._load, self.num_workers, self.buffer_size), offset, len(mine))
        self.shuffle = shuffle
        return self._cur

    def _load(self, idx):
        if len(idx) == 0:
            return {}                                        # dummy batch: joins the collective, contributes nothing
        b = self.dataset.collater([self.dataset[int(i)] for i in idx])
        return _pin(b) if self.pin_memory else b

    def state_dict(self):
        """iterators.py:395-409"""
        epoch = self.epoch + 1 if self.end_of_epoch() else self.epoch
        return {"version": 2, "epoch": epoch, "iterations_in_epoch": 0 if self.end_of_epoch() else self.iterations_in_epoch,
                "shuffle": self.shuffle}

    def load_state_dict(self, sd):
        self.epoch, self._offset, self._cur = sd["epoch"], sd.get("iterations_in_epoch", 0), None
        self.shuffle = sd.get("shuffle", True)


class _Counting:
    """CountingIterator (iterators.py:27-94): knows how many batches were consumed, supports len()."""

    def __init__(self, it, start, total):
        self._it, self.n, self.total = iter(it), start, total

    def __len__(self):
        return self.total

    def __iter__(self):
        return self

    def __next__(self):
        x = next(self._it)
        self.n += 1
        return x

    def has_next(self):
        return self.n < self.total
