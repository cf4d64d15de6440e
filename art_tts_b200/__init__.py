"""art_tts_b200 -- B200-native (sm_100a) Monotonic Alignment Search for art-tts / Grad-TTS.

Scope: the one hot path  log_prior(mu_x, y) -> maximum_path -> path -> durations
(see DESIGN.md).  Python here is only the host-side mirror of the reference's operator
surface; the work happens in art_tts_b200/csrc (CUDA) behind include/mas_b200.h.
"""
from . import alignment, monotonic_align, utils  # noqa: F401
from .monotonic_align import (  # noqa: F401
    lengths_from_mask,
    maximum_path,
    maximum_path_from_prior,
    maximum_path_lengths,
)
from .alignment import alignment_losses  # noqa: F401
from .utils import generate_path, sequence_mask  # noqa: F401

__version__ = "0.1.0"
