"""neuspeech1_b200 -- B200-native (sm_100a) implementation of NeuSpeech's EEG-conditioned Whisper hot path.

Python host code over a C-ABI shared library of hand-written CUDA kernels (neuspeech1_b200/csrc, include/neuspeech_b200.h).
"""
from ._abi import NeuSpeechB200Error, LIB_PATH  # noqa: F401

__version__ = "0.1.0"
