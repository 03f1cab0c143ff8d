"""
classifying-vae-lstm_b200: B200-native (sm_100a) training + sampling hot path of the Classifying VAE
and Classifying VAE+LSTM (mobeets/classifying-vae-lstm), behind a C-ABI shared library
(include/clv_b200.h, libclv_b200.so) with a Python host that mirrors the reference's
get_model / fit / load_model / generate_sample surface.  No CPU fallback.
"""
from . import _lib  # noqa: F401
from ._lib import ClvError, lib  # noqa: F401

__all__ = ["ClvError", "lib"]
