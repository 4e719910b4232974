from .transformer import (  # noqa: F401
    SpeechTransformerConfig,
    SpeechTransformerDecoderBase,
    SpeechTransformerEncoderForPrediction,
    SpeechTransformerEncoderModel,
    SpeechTransformerModel,
    SpeechTransformerModelBase,
    SpeechTransformerTransducerModelBase,
)
