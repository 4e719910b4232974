from .transformer import (  # noqa: F401
    SpeechTransformerConfig,
    SpeechTransformerEncoderForPrediction,
    SpeechTransformerEncoderModel,
)
