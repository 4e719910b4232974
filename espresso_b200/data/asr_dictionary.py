"""Symbol <-> index mapping with the reference's conventions (host logic; espresso/data/asr_dictionary.py:17-110 on top
of fairseq/data/dictionary.py:17-330).

Index order: [<s> only if enable_bos] <pad> </s> <unk>, then the symbols of the dictionary file in file order
("<symbol> <count>" per line).  With enable_bos the ASR task uses <s> (index 0) as the CTC / transducer blank
(espresso/tasks/speech_recognition.py:324-328).  `count` feeds unigram label smoothing."""
import torch


class AsrDictionary:
    def __init__(self, bos="<s>", pad="<pad>", eos="</s>", unk="<unk>", space="<space>", enable_bos=False,
                 extra_special_symbols=None):
        self.bos_word, self.unk_word, self.pad_word, self.eos_word, self.space_word = bos, unk, pad, eos, space
        self.symbols, self.count, self.indices = [], [], {}
        if enable_bos:  # no bos in the dictionary unless asked for
            self.bos_index = self.add_symbol(bos)
        self.pad_index = self.add_symbol(pad, n=0)
        self.eos_index = self.add_symbol(eos, n=0)
        self.unk_index = self.add_symbol(unk, n=0)
        for s in extra_special_symbols or ():
            self.add_symbol(s, n=0)
        self.nspecial = len(self.symbols)
        self.space_index = -1
        self.non_lang_syms = None
        self.tokenizer = None
        self.bpe = None

    # ---- container protocol ----------------------------------------------------------------------
    def __len__(self):
        return len(self.symbols)

    def __getitem__(self, idx):
        return self.symbols[idx] if idx < len(self.symbols) else self.unk_word

    def __contains__(self, sym):
        return sym in self.indices

    def index(self, sym):
        return self.indices.get(sym, self.unk_index)

    def get_count(self, idx):
        return self.count[idx]

    def add_symbol(self, word, n=1, overwrite=False):
        if word in self.indices and not overwrite:
            idx = self.indices[word]
            self.count[idx] += n
            return idx
        idx = len(self.symbols)
        self.indices[word] = idx
        self.symbols.append(word)
        self.count.append(n)
        return idx

    # ---- special symbols ---------------------------------------------------------------------------
    def bos(self):
        if hasattr(self, "bos_index"):
            return self.bos_index
        raise NotImplementedError("this dictionary was built without <s> (enable_bos=False)")

    def pad(self):
        return self.pad_index

    def eos(self):
        return self.eos_index

    def unk(self):
        return self.unk_index

    def space(self):
        return self.space_index

    # ---- files ---------------------------------------------------------------------------------------
    @classmethod
    def load(cls, f, enable_bos=False, f_non_lang_syms=None):
        d = cls(enable_bos=enable_bos)
        d.add_from_file(f)
        d.space_index = d.indices.get(d.space_word, -1)
        if f_non_lang_syms is not None:
            with open(f_non_lang_syms, "r", encoding="utf-8") as fd:
                syms = [x.rstrip() for x in fd.readlines()]
            for s in syms:
                assert d.index(s) != d.unk(), "{} in {} is not in the dictionary".format(s, f_non_lang_syms)
            d.non_lang_syms = syms
        return d

    def add_from_file(self, f):
        if isinstance(f, str):
            with open(f, "r", encoding="utf-8") as fd:
                return self.add_from_file(fd)
        for line in f.readlines():
            try:
                word, field = line.rstrip().rsplit(" ", 1)
                overwrite = field == "#fairseq:overwrite"
                if overwrite:
                    word, field = word.rsplit(" ", 1)
                count = int(field)
            except ValueError:
                raise ValueError("Incorrect dictionary format, expected '<token> <cnt> [flags]': \"{}\"".format(line))
            if word in self and not overwrite:
                raise RuntimeError("Duplicate word found when loading Dictionary: '{}'.".format(word))
            self.add_symbol(word, n=count, overwrite=overwrite)

    def save(self, f):
        if isinstance(f, str):
            with open(f, "w", encoding="utf-8") as fd:
                return self.save(fd)
        for s, c in zip(self.symbols[self.nspecial:], self.count[self.nspecial:]):
            print("{} {}".format(s, c), file=f)

    # ---- word pieces (espresso/data/asr_dictionary.py:117-142) ----------------------------------------------------
    def build_bpe(self, bpe=None, sentencepiece_model=None, sentencepiece_enable_sampling=False, sentencepiece_alpha=None):
        """bpe: "characters_asr" (character units with this dictionary's space symbol and non-linguistic symbols),
        "sentencepiece" (needs sentencepiece_model), or None."""
        from .encoders import CharactersAsr, SentencepieceBPE

        if bpe in (None, "none"):
            self.bpe = None
        elif bpe == "characters_asr":
            self.bpe = CharactersAsr(space_symbol=self.space_word, non_lang_syms=self.non_lang_syms)
        elif bpe == "sentencepiece":
            self.bpe = SentencepieceBPE(sentencepiece_model, sentencepiece_enable_sampling, sentencepiece_alpha)
        else:
            raise ValueError("unsupported bpe %r (the ASR recipes use characters_asr and sentencepiece)" % (bpe,))

    def wordpiece_encode(self, x):
        if self.tokenizer is not None:
            x = self.tokenizer.encode(x)
        if self.bpe is not None:
            x = self.bpe.encode(x)
        return x

    def wordpiece_decode(self, x):
        if self.bpe is not None:
            x = self.bpe.decode(x)
        if self.tokenizer is not None:
            x = self.tokenizer.decode(x)
        return x

    # ---- text <-> indices --------------------------------------------------------------------------------
    def unk_string(self, escape=False):
        return "<{}>".format(self.unk_word) if escape else self.unk_word

    def string(self, tensor, bpe_symbol=None, escape_unk=False, extra_symbols_to_ignore=None, unk_string=None,
               include_eos=False, separator=" "):
        """Indices -> text: drops eos (unless include_eos) and bos, maps <unk>, optionally undoes sentencepiece /
        subword_nmt style BPE (fairseq/data/data_utils.py post_process)."""
        if torch.is_tensor(tensor) and tensor.dim() == 2:
            return "\n".join(self.string(t, bpe_symbol, escape_unk, extra_symbols_to_ignore, include_eos=include_eos) for t in tensor)
        ignore = set(extra_symbols_to_ignore or [])
        if not include_eos:
            ignore.add(self.eos())
        if hasattr(self, "bos_index"):
            ignore.add(self.bos())

        def tok(i):
            if i == self.unk():
                return unk_string if unk_string is not None else self.unk_string(escape_unk)
            return self[i]

        sent = separator.join(tok(int(i)) for i in tensor if int(i) not in ignore)
        if bpe_symbol == "sentencepiece":
            sent = sent.replace(" ", "").replace("▁", " ").strip()
        elif bpe_symbol == "wordpiece":
            sent = sent.replace(" ", "").replace("_", " ").strip()
        elif bpe_symbol == "letter":
            sent = sent.replace(" ", "").replace("|", " ").strip()
        elif bpe_symbol in ("subword_nmt", "@@ ", "@@"):
            sym = "@@ " if bpe_symbol == "subword_nmt" else bpe_symbol
            sent = (sent + " ").replace(sym, "").rstrip()
        elif bpe_symbol is not None and bpe_symbol != "none":
            sent = (sent + " ").replace(bpe_symbol, "").rstrip()
        return sent

    def encode_line(self, line, line_tokenizer=None, add_if_not_exist=False, append_eos=True, reverse_order=False):
        words = (line_tokenizer or (lambda s: s.strip().split()))(line)
        if reverse_order:
            words = list(reversed(words))
        ids = torch.empty(len(words) + (1 if append_eos else 0), dtype=torch.int32)
        for i, w in enumerate(words):
            ids[i] = self.add_symbol(w) if add_if_not_exist else self.index(w)
        if append_eos:
            ids[len(words)] = self.eos_index
        return ids

    def dummy_sentence(self, length):
        t = torch.empty(length).uniform_(self.nspecial, len(self)).long()
        t[-1] = self.eos()
        return t
