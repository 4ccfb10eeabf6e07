"""flo_b200 -- B200-native (sm_100a) implementation of flo's lossless ALPC encode path.

Drop-in for `libflo_audio::Encoder::{new, with_compression, encode}` only; see DESIGN.md.
"""
from ._lib import FMT_F32, FMT_PCM16, FMT_S32, FMT_U8, FloError, SO_PATH
from .encoder import Context, Decoder, Encoder, TrackSpec, default_context, encode_batch
from . import analysis, reflo
from .streaming import EncodedFrame, StreamingEncoder

__all__ = ["Encoder", "Decoder", "Context", "TrackSpec", "encode_batch", "default_context", "FloError", "FMT_F32", "FMT_PCM16", "FMT_U8", "FMT_S32",
           "SO_PATH", "reflo", "analysis", "StreamingEncoder", "EncodedFrame"]
