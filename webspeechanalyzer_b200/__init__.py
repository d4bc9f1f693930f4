"""webspeechanalyzer_b200 -- B200 (sm_100a) implementation of the formantanalyzer hot path.

Public surface = the reference's (formantanalyzer @1.1.6, inner module 1 of /root/reference/dist/main.js):
configure / LaunchAudioNodes / StopAudioNodes / set_predicted_label_for_segment, plus the batch Engine over
the C-ABI (include/fa_b200.h).  Importing the package does not load the CUDA library; the first call does,
and fails loudly if libfa_b200.so is missing (there is no CPU fallback).
"""
from ._ctypes_defs import FaConfig, FaCounts, FaSegment, FaSyllable, N_FEATURES  # noqa: F401
from .api import (LaunchAudioNodes, LaunchError, StopAudioNodes, configure,  # noqa: F401
                  set_predicted_label_for_segment)
from .engine import Engine, PinnedBuffer, synth_speech, synth_speech_i16_batch  # noqa: F401
