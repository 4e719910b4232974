"""Text <-> token-string encoders of the ASR task (host logic).

CharactersAsr: the `characters_asr` BPE of the character-level recipes (espresso/data/encoders/characters_asr.py:19-43 over
espresso/tools/utils.py:36-58 `tokenize`): a transcript becomes space-separated characters, blanks become the space
symbol, listed non-linguistic symbols (e.g. "<noise>") stay whole; decode() undoes it."""
import re


def tokenize(sent, space="<space>", non_lang_syms=None):
    sent = " ".join(sent.strip().split())
    pieces, i = [], 0
    if non_lang_syms:
        for mt in re.finditer("|".join(re.escape(s) for s in non_lang_syms), sent):
            pieces.extend(sent[i: mt.start()])
            pieces.append(mt.group(0))
            i = mt.end()
    pieces.extend(sent[i:])
    return " ".join(space if p == " " else p for p in pieces)


class CharactersAsr:
    def __init__(self, space_symbol="<space>", ends_with_space=True, non_lang_syms=None):
        self.space_symbol, self.ends_with_space, self.non_lang_syms = space_symbol, ends_with_space, non_lang_syms

    def encode(self, x):
        y = tokenize(x, space=self.space_symbol, non_lang_syms=self.non_lang_syms)
        return y + " " + self.space_symbol if self.ends_with_space else y

    def decode(self, x):
        return x.replace(" ", "").replace(self.space_symbol, " ").strip()


class SentencepieceBPE:
    """fairseq/data/encoders/sentencepiece_bpe.py:31-70 (the `sentencepiece` bpe of the LibriSpeech / SWBD recipes)."""

    def __init__(self, model_path, enable_sampling=False, alpha=None):
        import sentencepiece as spm

        self.enable_sampling, self.alpha = enable_sampling, alpha
        self.sp = spm.SentencePieceProcessor()
        self.sp.Load(model_path)

    def encode(self, x):
        return " ".join(self.sp.Encode(x, out_type=str, enable_sampling=self.enable_sampling, alpha=self.alpha))

    def decode(self, x):
        return x.replace(" ", "").replace("\u2581", " ").strip()
