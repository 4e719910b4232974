"""Word / character error scoring of decoded hypotheses: the object `speech_recognize.py` feeds with every hypothesis
(espresso/tools/wer.py:16-220; alignment and the three-line aligned print: espresso/tools/utils.py:265-423).  Host logic.

`Scorer(dictionary, wer_output_filter=None)` keeps
  * the sub-word ("char") level edit counts between the token strings,
  * the word level counts after `dictionary.wordpiece_decode` and the optional sed-style filters (`s/pat/repl/g`, `s:pat:repl:g`),
  * per utterance the decoded strings and an aligned REF / HYP / STP block,
and reports CER / WER with their substitution / insertion / deletion shares.

Alignment: Levenshtein with unit costs; among equally cheap alignments the back-trace prefers, in this order, a match, a
substitution, an insertion (extra hypothesis word) and a deletion -- the order that fixes which of several optimal alignments is
printed and how errors split into sub / ins / del."""
import re
from collections import Counter, OrderedDict

import numpy as np


def align(ref, hyp):
    """ref, hyp: lists of words.  Returns the edit operations ('corr' | 'sub' | 'ins' | 'del') of one optimal alignment,
    in sentence order."""
    n, m = len(ref), len(hyp)
    d = np.zeros((n + 1, m + 1), dtype=np.int64)
    d[:, 0] = np.arange(n + 1)
    d[0, :] = np.arange(m + 1)
    for i in range(1, n + 1):
        ri = ref[i - 1]
        for j in range(1, m + 1):
            d[i, j] = d[i - 1, j - 1] if ri == hyp[j - 1] else 1 + min(d[i - 1, j - 1], d[i, j - 1], d[i - 1, j])
    ops = []
    i, j = n, m
    while i or j:
        if i and j and ref[i - 1] == hyp[j - 1] and d[i, j] == d[i - 1, j - 1]:
            ops.append("corr")
            i, j = i - 1, j - 1
        elif i and j and d[i, j] == d[i - 1, j - 1] + 1:
            ops.append("sub")
            i, j = i - 1, j - 1
        elif j and d[i, j] == d[i, j - 1] + 1:
            ops.append("ins")
            j -= 1
        else:
            ops.append("del")
            i -= 1
    ops.reverse()
    return ops


def edit_counts(ref, hyp):
    """Counter with 'words' (= len(ref)), 'corr', 'sub', 'ins', 'del', and the operations themselves."""
    ops = align(ref, hyp)
    c = Counter({"words": len(ref), "corr": 0, "sub": 0, "ins": 0, "del": 0})
    c.update(ops)
    return c, ops


def aligned_block(ref, hyp, ops):
    """The REF / HYP / STP / WER block: one column per operation, as wide as the longer of its two words; an insertion leaves the
    REF cell blank, a deletion the HYP cell; STP marks S / I / D."""
    if not ops:
        return "REF: \nHYP: \nSTP: \nWER: %.2f%%\n\n" % 0.0
    cols = []
    ri = hi = 0
    for op in ops:
        r = ref[ri] if op != "ins" else ""
        h = hyp[hi] if op != "del" else ""
        ri += op != "ins"
        hi += op != "del"
        w = max(len(r), len(h))
        mark = {"corr": " ", "sub": "S", "ins": "I", "del": "D"}[op]
        cols.append((r.ljust(w), h.ljust(w), mark.ljust(w)))
    errs = sum(op != "corr" for op in ops)
    wer = 100.0 * errs / len(ref) if ref else 0.0
    lines = ["%s: %s\n" % (tag, " ".join(c[k] for c in cols)) for k, tag in enumerate(("REF", "HYP", "STP"))]
    return "".join(lines) + "WER: %.2f%%\n\n" % wer


class Scorer:
    def __init__(self, dictionary, wer_output_filter=None):
        self.dictionary = dictionary
        self.ordered_utt_list = None
        self.word_filters = []
        if wer_output_filter:
            self._read_filters(wer_output_filter)
        self.reset()

    def reset(self):
        self.char_counter, self.word_counter = Counter(), Counter()
        self.char_results, self.results, self.aligned_results = OrderedDict(), OrderedDict(), OrderedDict()

    def _read_filters(self, path):
        with open(path, "r", encoding="utf-8") as f:
            for line in f:
                line = line.strip()
                if not line or line.startswith("#!"):
                    continue
                m = re.match(r"s/(.+)/(.*)/g", line) if line.startswith("s/") else (
                    re.match(r"s:(.+):(.*):g", line) if line.startswith("s:") else None)
                if m is not None:
                    self.word_filters.append((m.group(1), m.group(2)))  # other patterns are ignored, like the reference does

    @staticmethod
    def _check(utt_id, *strings):
        if not isinstance(utt_id, str):
            raise TypeError("utt_id must be a string(got {})".format(type(utt_id)))
        for s in strings:
            if not isinstance(s, str):
                raise TypeError("ref / pred must be strings (got {})".format(type(s)))

    def add_prediction(self, utt_id, pred):
        self._check(utt_id, pred)
        assert utt_id not in self.char_results and utt_id not in self.results, "Duplicated utterance id detected: %s" % utt_id
        self.char_results[utt_id] = pred + "\n"
        self.results[utt_id] = self.dictionary.wordpiece_decode(pred) + "\n"

    def add_evaluation(self, utt_id, ref, pred):
        self._check(utt_id, ref, pred)
        nls = getattr(self.dictionary, "non_lang_syms", None)
        if nls:  # non-linguistic symbols count neither as words nor as errors
            ref = " ".join(x for x in ref.strip().split() if x not in nls)
            pred = " ".join(x for x in pred.strip().split() if x not in nls)
        c, _ = edit_counts(ref.strip().split(), pred.strip().split())
        self.char_counter += c
        ref_words, pred_words = self.dictionary.wordpiece_decode(ref), self.dictionary.wordpiece_decode(pred)
        for pat, repl in self.word_filters:
            ref_words, pred_words = re.sub(pat, repl, ref_words), re.sub(pat, repl, pred_words)
        rw, pw = ref_words.split(), pred_words.split()
        c, ops = edit_counts(rw, pw)
        self.word_counter += c
        assert utt_id not in self.aligned_results, "Duplicated utterance id detected: %s" % utt_id
        self.aligned_results[utt_id] = aligned_block(rw, pw, ops)

    @staticmethod
    def _rates(c):
        assert c["words"] > 0
        n = float(c["words"])
        return (100.0 * (c["sub"] + c["ins"] + c["del"]) / n, 100.0 * c["sub"] / n, 100.0 * c["ins"] / n, 100.0 * c["del"] / n)

    def cer(self):
        return self._rates(self.char_counter)

    def wer(self):
        return self._rates(self.word_counter)

    def tot_word_error(self):
        return self.word_counter["sub"] + self.word_counter["ins"] + self.word_counter["del"]

    def tot_word_count(self):
        return self.word_counter["words"]

    def tot_char_error(self):
        return self.char_counter["sub"] + self.char_counter["ins"] + self.char_counter["del"]

    def tot_char_count(self):
        return self.char_counter["words"]

    def add_ordered_utt_list(self, *args):
        """Either one list of utterance ids, or text files whose first column is the id."""
        if len(args) == 1 and isinstance(args[0], list):
            self.ordered_utt_list = args[0]
            return
        self.ordered_utt_list = []
        for path in args:
            with open(path, "r", encoding="utf-8") as f:
                self.ordered_utt_list.extend(line.strip().split()[0] for line in f)
        for table in (self.char_results, self.results, self.aligned_results):
            if len(table):
                assert set(self.ordered_utt_list) == set(table.keys())

    def _dump(self, table, sep):
        ids = self.ordered_utt_list if self.ordered_utt_list is not None else list(table)
        if self.ordered_utt_list is not None:
            assert set(ids) == set(table.keys())
        return "".join(u + sep + table[u] for u in ids)

    def print_char_results(self):
        return self._dump(self.char_results, " ")

    def print_results(self):
        return self._dump(self.results, " ")

    def print_aligned_results(self):
        return self._dump(self.aligned_results, "\n")
