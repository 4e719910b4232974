__version__ = "2.0.0"


class BLEU:
    TOKENIZERS = ["none", "13a", "intl", "zh", "ja-mecab"]


class _M:
    BLEU = BLEU


metrics = _M()
