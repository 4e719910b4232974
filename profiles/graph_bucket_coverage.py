"""How many CUDA graphs does LibriSpeech-shape data need, and how often is a step a replay?  (host logic only, no GPU)

Draws a large pool of utterance durations from the bench distribution, packs it with the bench's batch_by_size settings and
computes, for every batch, the key the trainer captures graphs under (espresso_b200/trainer.py::_signature: batch size, waveform
length padded to a multiple of `bucket_frames` feature frames, mask-descriptor counts padded to 4, target length padded to
`bucket_tokens`, and the "some utterance is shorter than the padded encoder axis" predicate).  A bucket's first occurrence runs
eagerly, its second captures, everything after replays.

    python profiles/graph_bucket_coverage.py [pool_size]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from espresso_b200.data import batching, specaugment as SA  # noqa: E402

pool = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
BF, BT = 64, 16  # Trainer defaults: bucket_frames, bucket_tokens
rs = np.random.RandomState(123)
durs = np.clip(rs.gamma(6.1, 2.0, size=pool), 1.0, 35.0)
n_samples = np.round(durs * 16000).astype(np.int64)
frames = 1 + (n_samples - 400) // 160
order = batching.ordered_indices(frames)
batches = batching.batch_by_size(order, frames, bench.MAX_TOKENS, bench.MAX_SENTENCES)
np.random.RandomState(5).shuffle(batches)
cfg = SA.AdaptiveSpecAugmentConfig.from_config_dict(bench.SPECAUG)


def up(n, m):
    return (n + m - 1) // m * m


def key_of(idx):
    f = frames[idx]
    fmax = int(f.max())
    nf = nt = 0
    for i in idx:
        with SA.numpy_seed(1, 1, int(i)):
            fm, tm = SA.draw_masks(cfg, int(frames[i]), 80)
        nf, nt = max(nf, len(fm)), max(nt, len(tm))
    U = max(max(1, int(round(4.0 * durs[i]))) for i in idx) + 1
    tp = lambda t: -(-(-(-t // 2)) // 2)  # noqa: E731  frames after the two stride-2 convolutions
    pads = bool((np.array([tp(int(x)) for x in f]) < tp(up(fmax, BF))).any())
    return (len(idx), up(fmax, BF), up(nf, 4), up(nt, 4), up(U, BT), pads)


seen = {}
state = []  # 0 eager (first sight), 1 capture (second sight), 2 replay
for b in batches:
    k = key_of(b)
    c = seen.get(k, 0)
    state.append(min(c, 2))
    seen[k] = c + 1
state = np.array(state)
n = len(batches)
print("pool of %d utterances -> %d batches (max_tokens %d, max_sentences %d), one epoch in shuffled order" % (pool, n, bench.MAX_TOKENS, bench.MAX_SENTENCES))
print("distinct graph keys: %d  (bucket_frames %d, bucket_tokens %d)" % (len(seen), BF, BT))
for lo, hi in ((0, 100), (100, 500), (500, 1000), (1000, n)):
    if lo >= n:
        break
    s = state[lo:min(hi, n)]
    print("steps %5d-%5d: replay %5.1f %%  capture %4.1f %%  eager %4.1f %%" % (lo, min(hi, n), 100 * (s == 2).mean(), 100 * (s == 1).mean(), 100 * (s == 0).mean()))
print("whole epoch: replay %.1f %%; a second epoch over the same data replays %.1f %% of its steps"
      % (100 * (state == 2).mean(), 100 * np.mean([seen[key_of(b)] >= 2 for b in batches[:400]])))
sizes = sorted(seen.values(), reverse=True)
print("most frequent keys cover: top 16 -> %.1f %%, top 32 -> %.1f %%, top 64 -> %.1f %% of the batches" % tuple(100.0 * sum(sizes[:k]) / n for k in (16, 32, 64)))
