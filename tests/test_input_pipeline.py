"""Input pipeline and wire formats on the host: WAV decoding, Kaldi archives, the JSON
manifest -> datasets -> collated batches, epoch iteration with background prefetch, sharding and resumption."""
import io
import json
import os
import struct
import wave

import numpy as np
import pytest
import torch


def _write_wav16(path, x, rate=16000, channels=1):
    with wave.open(path, "wb") as w:
        w.setnchannels(channels)
        w.setsampwidth(2)
        w.setframerate(rate)
        w.writeframes(np.asarray(x, dtype="<i2").tobytes())


def test_wav_decoding_matches_libsndfile_conventions(tmp_path):
    from espresso_b200.data.audio_io import get_waveform

    rs = np.random.RandomState(0)
    x = rs.randint(-32768, 32767, size=4000).astype(np.int16)
    p = str(tmp_path / "a.wav")
    _write_wav16(p, x)
    w, sr = get_waveform(p, normalization=False)
    assert sr == 16000 and w.shape == (1, 4000) and w.dtype == np.float32
    assert np.array_equal(w[0], x.astype(np.float32))            # int16 scale exactly (float / 32768 * 32768)
    wn, _ = get_waveform(p, normalization=True, always_2d=False)
    assert wn.shape == (4000,) and np.allclose(wn, x / 32768.0)
    # stereo -> mono by averaging; frames / start windows; file objects
    st = rs.randint(-1000, 1000, size=(500, 2)).astype(np.int16)
    p2 = str(tmp_path / "s.wav")
    _write_wav16(p2, st.reshape(-1), channels=2)
    w2, _ = get_waveform(p2, normalization=False)
    assert np.allclose(w2[0], st.astype(np.float32).mean(1))
    w3, _ = get_waveform(open(p2, "rb"), normalization=False, mono=False, start=10, frames=20)
    assert w3.shape == (2, 20) and np.array_equal(w3[1], st[10:30, 1].astype(np.float32))
    # 24-bit PCM and 32-bit float, extensible header, a LIST chunk before data, streamed (unknown) data size
    v24 = np.array([0, 1, -1, 8388607, -8388608, 123456], dtype=np.int64)
    b24 = b"".join(struct.pack("<i", int(v))[:3] for v in v24)
    hdr = struct.pack("<HHIIHH", 0xFFFE, 1, 8000, 24000, 3, 24) + struct.pack("<HHI", 22, 24, 4) + struct.pack("<H", 1) + b"\x00" * 14
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(hdr)) + hdr + b"LIST" + struct.pack("<I", 3) + b"abc\x00" + b"data" + struct.pack("<I", 0xFFFFFFFF) + b24
    w24, sr24 = get_waveform(io.BytesIO(b"RIFF" + struct.pack("<I", 4 + len(body)) + body), normalization=True)
    assert sr24 == 8000 and np.allclose(w24[0], v24 / 8388608.0)
    f32 = np.array([0.5, -0.25, 1.0], dtype="<f4")
    body = b"WAVE" + b"fmt " + struct.pack("<I", 16) + struct.pack("<HHIIHH", 3, 1, 16000, 64000, 4, 32) + b"data" + struct.pack("<I", 12) + f32.tobytes()
    wf, _ = get_waveform(io.BytesIO(b"RIFF" + struct.pack("<I", 4 + len(body)) + body), normalization=False)
    assert np.allclose(wf[0], f32 * 32768.0)
    with pytest.raises(ValueError):
        get_waveform(str(tmp_path / "x.mp3"))


def test_kaldi_matrices_round_trip_and_compressed_known_answers(tmp_path):
    from espresso_b200.data.audio_io import num_frames_of, read_kaldi_mat, write_kaldi_mat

    rs = np.random.RandomState(1)
    mats = {"utt%d" % i: rs.randn(5 + 3 * i, 7).astype(np.float32 if i % 2 == 0 else np.float64) for i in range(4)}
    ark = str(tmp_path / "feats.ark")
    scp = {}
    with open(ark, "wb") as f:
        for k, m in mats.items():
            scp[k] = "%s:%d" % (ark, write_kaldi_mat(f, k, m))
    for k, m in mats.items():
        got = read_kaldi_mat(scp[k])
        assert got.dtype == np.float32 and np.allclose(got, m.astype(np.float32)) and num_frames_of(scp[k]) == m.shape[0]
    # compressed layouts (kaldi/src/matrix/compressed-matrix.h): build the byte streams from the published formulas
    gmin, grange, rows, cols = -2.0, 8.0, 3, 2
    glob = struct.pack("<ffii", gmin, grange, rows, cols)
    u16 = np.array([[0, 65535], [1000, 30000], [65535, 12345]], dtype="<u2")
    got = read_kaldi_mat(io.BytesIO(b"\x00BCM2 " + glob + u16.tobytes()))
    assert np.allclose(got, gmin + grange * u16.astype(np.float64) / 65535.0, atol=1e-6)
    u8 = np.array([[0, 255], [10, 200], [128, 64]], dtype=np.uint8)
    got = read_kaldi_mat(io.BytesIO(b"\x00BCM3 " + glob + u8.tobytes()))
    assert np.allclose(got, gmin + grange * u8.astype(np.float64) / 255.0, atol=1e-6)
    perc = np.array([[0, 16384, 49152, 65535], [100, 200, 300, 400]], dtype="<u2")      # per column p0 p25 p75 p100
    data = np.array([[0, 64, 255], [32, 128, 200]], dtype=np.uint8)                        # column-major [cols, rows]
    got = read_kaldi_mat(io.BytesIO(b"\x00BCM " + glob + perc.tobytes() + data.tobytes()))
    p = gmin + grange * perc.astype(np.float64) / 65535.0
    want = np.empty((rows, cols))
    for c in range(cols):
        for r in range(rows):
            v = float(data[c, r])
            p0, p25, p75, p100 = p[c]
            want[r, c] = (p0 + (p25 - p0) * v / 64.0) if v <= 64 else (p25 + (p75 - p25) * (v - 64) / 128.0) if v <= 192 else \
                (p75 + (p100 - p75) * (v - 192) / 63.0)
    assert np.allclose(got, want, atol=1e-5)
    with pytest.raises(ValueError):
        read_kaldi_mat(io.BytesIO(b"text matrix"))


def _corpus(tmp_path, n=11, fmt="wave"):
    from espresso_b200.data.asr_dictionary import AsrDictionary
    from espresso_b200.data.audio_io import write_kaldi_mat

    rs = np.random.RandomState(2)
    d = AsrDictionary()
    for c in "abcdefghij":
        d.add_symbol(c)
    d.add_symbol("<space>")
    d.space_index = d.index("<space>")
    d.build_bpe("characters_asr")
    man = {}
    ark = open(str(tmp_path / "f.ark"), "wb") if fmt == "feat" else None
    for i in range(n):
        uid = "u%02d" % i
        dur = 3000 + 700 * ((i * 5) % n)
        text = " ".join("".join(rs.choice(list("abcdefghij"), size=rs.randint(1, 5))) for _ in range(1 + i % 3))
        if fmt == "wave":
            p = str(tmp_path / (uid + ".wav"))
            _write_wav16(p, rs.randint(-3000, 3000, size=dur))
            man[uid] = {"wave": p, "text": text, "utt2num_frames": str(1 + (dur - 400) // 160)}
        elif fmt == "command":
            p = str(tmp_path / (uid + ".wav"))
            _write_wav16(p, rs.randint(-3000, 3000, size=dur))
            man[uid] = {"command": "cat %s |" % p, "text": text}
        else:
            m = rs.randn(20 + 3 * i, 40).astype(np.float32)
            man[uid] = {"feat": "%s:%d" % (str(tmp_path / "f.ark"), write_kaldi_mat(ark, uid, m)), "text": text,
                        "utt2num_frames": str(m.shape[0])}
    if ark:
        ark.close()
    with open(str(tmp_path / "train.json"), "w") as f:
        json.dump(man, f)
    return d, man


@pytest.mark.parametrize("fmt", ["wave", "command", "feat"])
def test_manifest_to_collated_batches(tmp_path, fmt):
    from espresso_b200.data.asr_dataset import get_asr_dataset_from_json

    d, man = _corpus(tmp_path, fmt=fmt)
    spec = "{'freq_mask_F': 10, 'freq_mask_N': 2, 'time_mask_pm': 0.2, 'time_mask_ps': 0.2, 'time_warp_W': 0}"
    ds = get_asr_dataset_from_json(str(tmp_path), "train", d, is_training_set=True, specaugment_config=spec, seed=3)
    assert len(ds) == len(man) and ds.src.input_format == fmt and ds.src.feat_dim == (40 if fmt == "feat" else 80)
    if fmt != "feat":   # frame counts: from the manifest, or counted from the audio (snip-edges)
        assert ds.src_sizes.tolist() == [1 + (3000 + 700 * ((i * 5) % len(man)) - 400) // 160 for i in range(len(man))]
    it = ds[4]
    assert it["utt_id"] == "u04" and it["text"] == man["u04"]["text"]
    toks = [d[int(t)] for t in it["target"]]
    # characters_asr: one unit per character, <space> between words AND after the last one, then </s>
    assert toks[-2:] == ["<space>", "</s>"] and "".join(t if t != "<space>" else " " for t in toks[:-2]) == it["text"]
    assert ds.tgt_sizes[4] == len(toks)
    order = ds.ordered_indices()
    assert sorted(order.tolist()) == list(range(len(ds)))
    batches = ds.batch_by_size(order, max_tokens=int(ds.src_sizes.max()) * 3, max_sentences=4)
    assert sorted(int(i) for b in batches for i in b) == list(range(len(ds)))
    b = ds.collater([ds[int(i)] for i in batches[0]])
    src, lens = b["net_input"]["src_tokens"], b["net_input"]["src_lengths"]
    assert lens.dtype == torch.int32 and (lens[:-1] >= lens[1:]).all() and src.shape[0] == len(batches[0])
    assert src.dim() == (3 if fmt == "feat" else 2) and b["target"].shape[0] == len(batches[0])
    assert b["net_input"]["prev_output_tokens"][:, 0].eq(d.eos()).all()
    assert "freq_masks" in b["net_input"] and b["net_input"]["freq_masks"].shape[0] == len(batches[0])
    # SpecAugment draws follow numpy_seed(seed, epoch, index): same epoch -> same masks, next epoch -> new ones
    again = ds[4]
    assert np.array_equal(np.asarray(again["time_masks"]), np.asarray(it["time_masks"]))
    ds.set_epoch(2)
    assert not np.array_equal(np.asarray(ds[4]["time_masks"]), np.asarray(it["time_masks"])) or len(it["time_masks"]) == 0
    # evaluation set: no masks
    ev = get_asr_dataset_from_json(str(tmp_path), "train", d, is_training_set=False, specaugment_config=spec)
    assert ev[0]["freq_masks"] is None
    with pytest.raises(FileNotFoundError):
        get_asr_dataset_from_json(str(tmp_path), "valid", d)


def test_pairing_drops_unmatched_utterances_and_combines_splits(tmp_path):
    from espresso_b200.data.asr_dataset import AsrDataset, AsrTextDataset, AudioFeatDataset, get_asr_dataset_from_json

    d, man = _corpus(tmp_path, n=6)
    ids = list(man)
    src = AudioFeatDataset(ids[:5], [man[u]["wave"] for u in ids[:5]], feat_dim=80, feature_type="fbank")
    tgt = AsrTextDataset(list(reversed(ids[1:])), [man[u]["text"] for u in reversed(ids[1:])], d)
    ds = AsrDataset(src, src.sizes, tgt, tgt.sizes, d)
    assert ds.src.utt_ids == ids[1:5] == ds.tgt.utt_ids and [ds[i]["text"] for i in range(4)] == [man[u]["text"] for u in ids[1:5]]
    # train.json + train1.json with the primary set up-sampled twice
    os.rename(str(tmp_path / "train.json"), str(tmp_path / "x.json"))
    a = {u: man[u] for u in ids[:2]}
    b = {u: man[u] for u in ids[2:]}
    json.dump(a, open(str(tmp_path / "train.json"), "w"))
    json.dump(b, open(str(tmp_path / "train1.json"), "w"))
    both = get_asr_dataset_from_json(str(tmp_path), "train", d, combine=True, upsample_primary=2)
    assert len(both) == 2 * 2 + 4 and both.src.utt_ids == ids[:2] * 2 + ids[2:]
    assert len(get_asr_dataset_from_json(str(tmp_path), "train", d, combine=False)) == 2


def test_epoch_iterator_prefetch_sharding_and_resume(tmp_path):
    from espresso_b200.data.asr_dataset import get_asr_dataset_from_json
    from espresso_b200.data.iterators import EpochBatchIterator, PrefetchIterator

    d, man = _corpus(tmp_path, n=11)
    ds = get_asr_dataset_from_json(str(tmp_path), "train", d, seed=5)
    batches = ds.batch_by_size(ds.ordered_indices(), max_sentences=2)
    assert len(batches) == 6

    def run(shards, rank, workers, state=None, take=None):
        it = EpochBatchIterator(ds, batches, seed=5, num_shards=shards, shard_id=rank, num_workers=workers, buffer_size=2, pin_memory=False)
        if state:
            it.load_state_dict(state)
        out, itr = [], it.next_epoch_itr()
        for b in itr:
            out.append(tuple(b["utt_id"]) if b else ())
            if take and len(out) == take:
                break
        return out, it

    one, it1 = run(1, 0, 3)
    assert len(one) == 6 and sorted(u for b in one for u in b) == sorted(man) and it1.end_of_epoch() and it1.next_epoch_idx == 2
    assert run(1, 0, 1)[0] == one                                  # order does not depend on the number of workers
    r0, r1 = run(2, 0, 2)[0], run(2, 1, 2)[0]
    assert len(r0) == len(r1) == 3 and [x for p in zip(r0, r1) for x in p] == one   # round-robin shards of the same order
    r = [run(4, k, 2)[0] for k in range(4)]
    assert all(len(x) == 2 for x in r) and r[2][1] == () and r[3][1] == ()          # dummy batches pad the last round
    part, itp = run(1, 0, 2, take=2)
    sd = itp.state_dict()
    assert sd["epoch"] == 1 and sd["iterations_in_epoch"] == 2
    rest, _ = run(1, 0, 2, state=sd)
    assert part + rest == one
    second = EpochBatchIterator(ds, batches, seed=5, pin_memory=False)
    second.load_state_dict({"epoch": 2, "iterations_in_epoch": 0})
    e2 = [tuple(b["utt_id"]) for b in second.next_epoch_itr()]
    assert sorted(e2) == sorted(one) and e2 != one                                   # reshuffled per epoch
    # worker exceptions reach the consumer; results stay ordered under uneven job times
    import time
    got = list(PrefetchIterator(range(20), lambda k: (time.sleep(0.002 * (k % 3)), k)[1], num_workers=4, buffer_size=3))
    assert got == list(range(20))

    def boom(k):
        if k == 3:
            raise RuntimeError("bad file")
        return k

    with pytest.raises(RuntimeError):
        list(PrefetchIterator(range(6), boom, num_workers=2))


def test_task_load_dataset_and_batch_iterator(tmp_path):
    """speech_recognition_espresso task: manifest -> dataset -> frozen batches -> prefetching epoch iterator."""
    from espresso_b200.tasks.speech_recognition import SpeechRecognitionEspressoConfig, SpeechRecognitionEspressoTask

    d, man = _corpus(tmp_path, n=9)
    d.save(str(tmp_path / "dict.txt"))
    os.rename(str(tmp_path / "train.json"), str(tmp_path / "train_si.json"))
    cfg = SpeechRecognitionEspressoConfig(criterion_name="ctc_loss", dict=str(tmp_path / "dict.txt"), data=str(tmp_path),
                                          train_subset="train_si", seed=4,
                                          specaugment_config="{'freq_mask_F': 5, 'freq_mask_N': 1, 'time_mask_pm': 0.1, 'time_mask_ps': 0.1, 'time_warp_W': 0}")
    task = SpeechRecognitionEspressoTask.setup_task(cfg)
    task.tgt_dict.build_bpe("characters_asr")
    ds = task.load_dataset("train_si")
    assert task.dataset("train_si") is ds and len(ds) == 9
    assert task.tgt_dict.count[task.tgt_dict.eos()] == 9 and task.tgt_dict.count[task.tgt_dict.unk()] == 0
    itr = task.get_batch_iterator(ds, max_tokens=int(ds.src_sizes.max()) * 2, max_sentences=3, seed=4, num_workers=2)
    seen = []
    for b in itr.next_epoch_itr():
        assert b["net_input"]["src_tokens"].dim() == 2 and b["net_input"]["freq_masks"] is not None   # raw waveforms + mask descriptors
        assert b["net_input"]["src_tokens"].shape[0] <= 3
        seen += b["utt_id"]
    assert sorted(seen) == sorted(man)
    with pytest.raises(KeyError):
        task.dataset("valid")


def test_global_cmvn_stats_match_the_reference_tool(monkeypatch, tmp_path):
    """espresso_b200.tools.compute_global_cmvn_stats (device front end + float64 sums) against the statistics the reference's
    tool produces with its own fbank and pooling formula (tests/golden/global_cmvn.npz); host path over oracle/ops_ref, incl.
    the command-line entry reading WAVE files."""
    from espresso_b200 import ops
    from espresso_b200.tools import compute_global_cmvn_stats as G
    from oracle import frontend as OF
    from oracle import ops_ref

    monkeypatch.setattr(ops, "frontend_fbank", ops_ref.frontend_fbank)
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "global_cmvn.npz"))
    waves = [OF.synth_waveform(int(g["seed0"]) + i, float(d)) for i, d in enumerate(g["durs"])]
    mean, std, n = G.global_cmvn_stats(waves, torch.device("cpu"), batch_seconds=5.0)  # several batches
    assert n == int(g["frames"])
    assert np.abs(mean - g["mean"]).max() < 2e-4 and np.abs(std - g["std"]).max() < 2e-4
    # the command-line entry: wav.scp -> gcmvn.npz readable by the on-the-fly front end
    lines = []
    for i, w in enumerate(waves):
        p = str(tmp_path / ("u%d.wav" % i))
        _write_wav16(p, w.astype(np.int16))
        lines.append("utt%d %s\n" % (i, p))
    (tmp_path / "wav.scp").write_text("".join(lines))
    G.main([str(tmp_path / "wav.scp"), str(tmp_path / "out"), "--device", "cpu", "--max-num-utts", "6"])
    st = np.load(str(tmp_path / "out" / "gcmvn.npz"))
    assert np.abs(st["mean"] - g["mean"]).max() < 2e-4 and np.abs(st["std"] - g["std"]).max() < 2e-4
    from espresso_b200.data.frontend import OnTheFlyFbank

    assert OnTheFlyFbank.from_npz(str(tmp_path / "out" / "gcmvn.npz")) is not None


def test_recognize_loop_from_manifest_to_wer(tmp_path, monkeypatch):
    """espresso_b200.speech_recognize.recognize (the reference's decode-and-score loop, espresso/speech_recognize.py:60-400) end
    to end on the host path: WAVE files + JSON manifest -> dataset -> batch iterator -> on-the-fly front end -> encoder-decoder
    beam search -> T- / H- lines -> Scorer."""
    from argparse import Namespace

    from espresso_b200 import ops
    from espresso_b200.data.frontend import OnTheFlyFbank
    from espresso_b200.models import SpeechTransformerConfig, SpeechTransformerModelBase
    from espresso_b200.speech_recognize import recognize
    from espresso_b200.tasks.speech_recognition import SpeechRecognitionEspressoConfig, SpeechRecognitionEspressoTask
    from oracle import ops_ref

    for name in dir(ops_ref):   # host orchestration over the oracle's statement of every kernel
        if not name.startswith("_") and callable(getattr(ops_ref, name)) and hasattr(ops, name):
            monkeypatch.setattr(ops, name, getattr(ops_ref, name))
    d, man = _corpus(tmp_path, n=7, fmt="wave")
    os.rename(str(tmp_path / "train.json"), str(tmp_path / "test.json"))
    cfg = SpeechRecognitionEspressoConfig(criterion_name="label_smoothed_cross_entropy_v2", data=str(tmp_path), autoregressive=True)
    task = SpeechRecognitionEspressoTask(cfg, d, feat_dim=80)
    torch.manual_seed(3)
    mcfg = SpeechTransformerConfig.from_dict(dict(
        dropout=0.0, attention_dropout=0.0, activation_dropout=0.0, layernorm_embedding=False, max_target_positions=64,
        encoder=dict(embed_dim=64, ffn_embed_dim=128, layers=1, attention_heads=4, normalize_before=True, learned_pos=False,
                     relative_positional_embeddings=True, layer_type="transformer"),
        decoder=dict(embed_dim=64, ffn_embed_dim=128, layers=1, attention_heads=4, normalize_before=True, learned_pos=False,
                     relative_positional_embeddings=False, input_dim=64, output_dim=64)))
    m = SpeechTransformerModelBase.build_model(mcfg, task).finalize_(torch.device("cpu"))
    m.frontend = OnTheFlyFbank(None, None)
    args = Namespace(beam=2, max_len_a=0.0, max_len_b=6, lm_weight=0.0)
    out = recognize(task, [m], subset="test", gen_args=args, max_tokens=None, max_sentences=3, device="cpu", quiet=True)
    sc = out["scorer"]
    assert out["num_sentences"] == len(man) == len(sc.results) == len(sc.aligned_results)
    t_lines = [l for l in out["lines"] if l.startswith("T-")]
    h_lines = [l for l in out["lines"] if l.startswith("H-")]
    assert len(t_lines) == len(h_lines) == len(man)
    assert {l.split("\t")[0][2:]: l.split("\t")[1] for l in t_lines} == {u: v["text"] for u, v in man.items()}
    for l in h_lines:   # the printed hypothesis is the word-level decoding of what the scorer stored; scores are base-2 logs
        uid, hyp, score = l.split("\t")
        assert sc.results[uid[2:]].rstrip("\n") == hyp and float(score) <= 0.0
    wer, sub, ins, dele = sc.wer()
    assert wer >= 0 and abs(wer - (sub + ins + dele)) < 1e-9 and sc.tot_word_count() == sum(len(v["text"].split()) for v in man.values())
    # the loop is deterministic and independent of how the subset is sharded
    again = recognize(task, [m], subset="test", gen_args=args, max_tokens=None, max_sentences=3, device="cpu")
    assert again["scorer"].print_results() == sc.print_results()
    parts = [recognize(task, [m], subset="test", gen_args=args, max_sentences=3, max_tokens=None, device="cpu", num_shards=2, shard_id=r)
             for r in range(2)]
    assert sum(p["num_sentences"] for p in parts) == len(man)
    assert set(parts[0]["scorer"].results) | set(parts[1]["scorer"].results) == set(man)
