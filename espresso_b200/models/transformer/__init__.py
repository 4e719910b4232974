from .speech_transformer_config import SpeechTransformerConfig  # noqa: F401
from .speech_transformer_encoder_model import (  # noqa: F401
    SpeechTransformerEncoderForPrediction,
    SpeechTransformerEncoderModel,
)
from .speech_transformer_base import SpeechTransformerDecoderBase, SpeechTransformerModel, SpeechTransformerModelBase  # noqa: F401
from .speech_transformer_transducer_base import SpeechTransformerTransducerModelBase  # noqa: F401
