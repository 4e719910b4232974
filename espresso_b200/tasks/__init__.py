from .speech_recognition import SpeechRecognitionEspressoConfig, SpeechRecognitionEspressoTask  # noqa: F401
