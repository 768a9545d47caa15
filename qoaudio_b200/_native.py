"""ctypes binding of libqoa_b200.so — the C ABI declared in include/qoa_b200.h.

The library is built in-tree by `make -C qoaudio_b200/csrc` (see __graft_entry__.build).  There is no CPU fallback:
if the shared object is missing this module raises, and every compute call fails when no sm_100 device is usable.
"""
import ctypes as C
import os

_DIR = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_DIR, "libqoa_b200.so")

OK = 0
ENC_INVALID_CHANNELS, ENC_INVALID_SAMPLE_RATE, ENC_INVALID_SAMPLES, ENC_IO_ERROR = 1, 2, 3, 4
DEC_NOT_QOA_FILE, DEC_NO_SAMPLES, DEC_INVALID_FRAME_HEADER, DEC_INCOMPATIBLE_FRAME, DEC_IO_ERROR = 11, 12, 13, 14, 15
DEC_OUTPUT_TOO_SMALL = 16
E_ARG, E_CUDA, E_NOMEM = -1, -2, -3
DECODE_ITER, DECODE_ALL, DECODE_FRAMES = 0, 1, 2
INFO_UNIFORM_CH, INFO_ALIGNED16 = 1, 2


class Desc(C.Structure):
    _fields_ = [("channels", C.c_uint32), ("sample_rate", C.c_uint32), ("samples", C.c_uint32)]


class StreamInfo(C.Structure):
    _fields_ = [("status", C.c_int32), ("streaming", C.c_int32), ("header_samples", C.c_uint32),
                ("channels", C.c_uint32), ("sample_rate", C.c_uint32), ("index_frames", C.c_uint32),
                ("samples", C.c_uint64), ("frames", C.c_uint64), ("max_channels", C.c_uint32), ("flags", C.c_uint32)]


class FrameHeaderC(C.Structure):
    _fields_ = [("num_channels", C.c_uint32), ("sample_rate", C.c_uint32), ("num_samples_per_channel", C.c_uint32)]


# every symbol include/qoa_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_SZ = C.c_size_t
SYMBOLS = {
    "qoa_b200_version": (C.c_char_p, []),
    "qoa_b200_ctx_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "qoa_b200_ctx_destroy": (None, [_P]),
    "qoa_b200_last_error": (C.c_char_p, [_P]),
    "qoa_b200_kernel_launches": (C.c_uint64, [_P]),
    "qoa_b200_encoded_size": (_SZ, [C.c_uint32, C.c_uint32]),
    "qoa_b200_frame_size": (_SZ, [C.c_uint32, C.c_uint32]),
    "qoa_b200_encode_all_batch": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _SZ]),
    "qoa_b200_encode_all_batch_dev": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "qoa_b200_encoder_batch_new": (C.c_int, [_P, _P, _SZ, _P, C.POINTER(_P)]),
    "qoa_b200_encoder_batch_free": (None, [_P]),
    "qoa_b200_encoder_batch_write_header": (C.c_int, [_P, _P]),
    "qoa_b200_encoder_batch_encode_frame": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "qoa_b200_encoder_batch_encode_chunk": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "qoa_b200_encoder_batch_encode_chunk_dev": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "qoa_b200_encoder_batch_get_lms": (C.c_int, [_P, _SZ, C.c_uint32, _P, _P]),
    "qoa_b200_decode_plan_create": (C.c_int, [_P, _P, _P, _SZ, C.c_int, C.POINTER(_P)]),
    "qoa_b200_decode_plan_create_dev": (C.c_int, [_P, _P, _P, _P, _SZ, C.c_int, _P, C.POINTER(_P)]),
    "qoa_b200_decode_plan_destroy": (None, [_P]),
    "qoa_b200_decode_plan_info": (C.c_int, [_P, _P]),
    "qoa_b200_decode_plan_frame_headers": (_SZ, [_P, _SZ, _P, _SZ]),
    "qoa_b200_decode_plan_run": (C.c_int, [_P, _P, _P, _P]),
    "qoa_b200_decode_plan_run_dev": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "qoa_b200_probe": (C.c_int, [_P, _SZ, C.c_int, _P]),
    "qoa_b200_decode_all_batch": (C.c_int, [_P, _P, _P, _SZ, C.c_int, _P, _P, _P, _P]),
    "qoa_b200_gen_pcm_dev": (C.c_int, [_P, _P, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64,
                                       C.c_int, _P]),
    "qoa_b200_gen_qoa_dev": (C.c_int, [_P, _P, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
                                       C.c_uint64, _P]),
    "qoa_b200_int_issue_peak": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double)]),
    "qoa_b200_ctx_create_multi": (C.c_int, [_P, C.c_int, C.POINTER(_P)]),
    "qoa_b200_mctx_destroy": (None, [_P]),
    "qoa_b200_mctx_device_count": (C.c_int, [_P]),
    "qoa_b200_mctx_device": (C.c_int, [_P, C.c_int]),
    "qoa_b200_mctx_ctx": (_P, [_P, C.c_int]),
    "qoa_b200_mctx_last_error": (C.c_char_p, [_P]),
    "qoa_b200_mctx_kernel_launches": (C.c_uint64, [_P]),
    "qoa_b200_partition": (C.c_int, [_P, _SZ, C.c_int, _P]),
    "qoa_b200_mctx_partition": (C.c_int, [_P, _P, _SZ, _P]),
    "qoa_b200_mctx_encode_all_batch": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _SZ]),
    "qoa_b200_mctx_decode_all_batch": (C.c_int, [_P, _P, _P, _SZ, C.c_int, _P, _P, _P, _P]),
}

_lib = None


def lib():
    """Loads libqoa_b200.so (no compute happens at load time, so this also works on a CPU-only box)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                f"{SO_PATH} is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; "
                "g.build()' or make -C qoaudio_b200/csrc).  qoaudio_b200 has no CPU fallback.")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError here = header and library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class NativeError(RuntimeError):
    """Call-level failure of the C ABI (bad argument, CUDA failure, out of memory)."""


def check(ctx_handle, rc, what):
    if rc != 0:
        msg = lib().qoa_b200_last_error(ctx_handle)
        raise NativeError(f"{what} failed with {rc}: {msg.decode() if msg else '?'}")
