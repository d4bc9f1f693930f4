"""webspeechanalyzer_b200 -- B200 (sm_100a) implementation of the formantanalyzer hot path."""
from ._ctypes_defs import FaConfig, FaCounts, FaSegment, FaSyllable, N_FEATURES  # noqa: F401
