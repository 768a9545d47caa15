"""qoaudio_b200 — B200-native batched QOA codec engine with the API surface of the `qoaudio` Rust crate.

Hand-written sm_100a CUDA kernels (qoaudio_b200/csrc) behind a C ABI (include/qoa_b200.h); this package is the
host-side mirror of the crate's API.  There is no CPU fallback.
"""
from .api import (MAX_SLICES_PER_CHANNEL_PER_FRAME, QOA_FRAME_LEN, QOA_HEADER_SIZE, QOA_LMS_LEN, QOA_MAGIC,  # noqa: F401
                  QOA_MAX_CHANNELS, QOA_SLICE_LEN, QOA_SLICES_PER_FRAME, DecodedQoa, DecodeError, EncodeError,
                  FrameHeader, ProcessingMode, QoaDecoder, QoaDesc, QoaEncoder, QoaEncoderBatch, QoaItem, decode_all,
                  decode_all_batch, encode_all, encode_all_batch, open_and_decode_all)
from . import wavio  # noqa: F401
from .engine import DecodePlan, EncoderBatch, Engine, MultiEngine, default_engine, partition, probe  # noqa: F401
