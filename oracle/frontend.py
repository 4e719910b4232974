"""Oracle: Kaldi-compatible log-mel filterbank + global CMVN + adaptive SpecAugment (numpy).

Restates, for the defaults Espresso uses (espresso/tools/utils.py:426-454):
  torchaudio/compliance/kaldi.py:154-218 (_get_window), :436-512 (get_mel_banks), :514-646 (fbank)
  fairseq/data/audio/feature_transforms/global_cmvn.py:26-29
  espresso/data/feature_transforms/adaptive_specaugment.py:77-136
  fairseq/data/data_utils.py:127-140 (numpy_seed)
Test infrastructure only (see oracle/__init__.py).
"""
import contextlib
import math

import numpy as np

FRAME_LEN, FRAME_SHIFT, NFFT, NUM_BINS = 400, 160, 512, 80
EPS = np.float32(1.1920928955078125e-07)


def num_frames(n_samples: int) -> int:
    # snip_edges=True (espresso/tools/utils.py:457-486)
    return 0 if n_samples < FRAME_LEN else 1 + (n_samples - FRAME_LEN) // FRAME_SHIFT


def povey_window(dtype=np.float32):
    n = np.arange(FRAME_LEN, dtype=np.float64)
    hann = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / (FRAME_LEN - 1))
    return np.power(hann.astype(dtype), dtype(0.85)).astype(dtype)  # kaldi.py:98-100


def mel_banks(dtype=np.float32):
    """[80, 256] triangular filters, low 20 Hz, high Nyquist (kaldi.py:436-512)."""
    def mel(f):
        return dtype(1127.0) * np.log(dtype(1.0) + np.asarray(f, dtype=dtype) / dtype(700.0))
    low, high = mel(20.0), mel(8000.0)
    delta = (high - low) / dtype(NUM_BINS + 1)
    b = np.arange(NUM_BINS, dtype=dtype)[:, None]
    left, center, right = low + b * delta, low + (b + 1.0) * delta, low + (b + 2.0) * delta
    m = mel(dtype(16000.0 / NFFT) * np.arange(NFFT // 2, dtype=dtype))[None, :]
    up = (m - left) / (center - left)
    down = (right - m) / (right - center)
    return np.maximum(dtype(0.0), np.minimum(up, down)).astype(dtype)


def kaldi_fbank(wave, dtype=np.float32):
    """wave: 1-D array in int16 value range -> [m, 80] log-mel energies."""
    wave = np.asarray(wave, dtype=dtype)
    m = num_frames(len(wave))
    if m == 0:
        return np.zeros((0, NUM_BINS), dtype=dtype)
    idx = np.arange(m)[:, None] * FRAME_SHIFT + np.arange(FRAME_LEN)[None, :]
    fr = wave[idx]                                              # strided framing (kaldi.py:171)
    fr = fr - fr.mean(axis=1, keepdims=True, dtype=dtype)       # remove_dc_offset (:183-186)
    prev = np.concatenate([fr[:, :1], fr[:, :-1]], axis=1)      # replicate pad (:193-198)
    fr = fr - dtype(0.97) * prev
    fr = fr * povey_window(dtype)[None, :]
    fr = np.pad(fr, ((0, 0), (0, NFFT - FRAME_LEN)))            # :207-212
    spec = np.abs(np.fft.rfft(fr, axis=1)).astype(dtype) ** 2   # :616-618 (power spectrum)
    mel = spec[:, : NFFT // 2] @ mel_banks(dtype).T             # :621-630 (last fft bin has zero weight)
    return np.log(np.maximum(mel.astype(dtype), EPS)).astype(dtype)


def global_cmvn(x, mean, std):
    # float64 promotion is the reference behaviour (global_cmvn.py:26-29 with float64 stats)
    return (x - np.asarray(mean, dtype=np.float64)) / np.asarray(std, dtype=np.float64)


@contextlib.contextmanager
def numpy_seed(seed, *addl):
    """fairseq/data/data_utils.py:127-140."""
    if seed is None:
        yield
        return
    if len(addl) > 0:
        seed = int(hash((seed, *addl)) % 1e6)
    state = np.random.get_state()
    np.random.seed(seed)
    try:
        yield
    finally:
        np.random.set_state(state)


def adaptive_specaugment(spec, freq_mask_n=2, freq_mask_f=27, time_mask_pm=0.04, time_mask_ps=0.04,
                         time_mask_n=2, time_mask_t=40, time_mask_p=1.0, return_masks=False):
    """adaptive_specaugment.py:77-136 with time_warp_W = 0; draws from the global numpy RNG."""
    out = spec.copy()
    m, nf = spec.shape
    fill = spec.mean()
    fmasks, tmasks = [], []
    done = m == 0 or nf < freq_mask_f
    if not done:
        for _ in range(freq_mask_n):
            f = np.random.randint(0, freq_mask_f)
            f0 = np.random.randint(0, nf - f)
            fmasks.append((f0, f))
            if f != 0:
                out[:, f0:f0 + f] = fill
        max_t = (min(time_mask_t, math.floor(m * time_mask_p)) if time_mask_ps is None
                 else math.floor(m * time_mask_ps))
        if max_t >= 1:
            n = time_mask_n if time_mask_pm is None else min(20, math.floor(m * time_mask_pm))
            for _ in range(n):
                t = np.random.randint(0, max_t)
                t0 = np.random.randint(0, m - t)
                tmasks.append((t0, t))
                if t != 0:
                    out[t0:t0 + t, :] = fill
    if return_masks:
        return out, fmasks, tmasks
    return out


def synth_waveform(i, dur_s):
    """SURVEY.md §8(d) synthetic utterance i: Gaussian noise + a sine, int16 range, float32."""
    rs = np.random.RandomState(1000 + i)
    n = int(round(dur_s * 16000))
    f0 = rs.uniform(80.0, 400.0)
    t = np.arange(n) / 16000.0
    x = np.round(3000.0 * rs.randn(n) + 1500.0 * np.sin(2 * np.pi * f0 * t))
    return np.clip(x, -32767, 32767).astype(np.float32)
