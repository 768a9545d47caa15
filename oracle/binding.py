"""ctypes binding of oracle/libqoa_oracle.so (TEST INFRASTRUCTURE ONLY; see oracle/qoa_oracle.h)."""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "libqoa_oracle.so")

OK = 0
ENC_INVALID_CHANNELS, ENC_INVALID_SAMPLE_RATE, ENC_INVALID_SAMPLES, ENC_IO_ERROR = 1, 2, 3, 4
DEC_NOT_QOA_FILE, DEC_NO_SAMPLES, DEC_INVALID_FRAME_HEADER, DEC_INCOMPATIBLE_FRAME, DEC_IO_ERROR = 11, 12, 13, 14, 15


def build(force=False):
    src = [os.path.join(_DIR, f) for f in ("qoa_oracle.c", "qoa_oracle.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-s", "-C", _DIR, "libqoa_oracle.so"])
    return _SO


class Lms(C.Structure):
    _fields_ = [("history", C.c_int32 * 4), ("weights", C.c_int32 * 4)]


class Encoder(C.Structure):
    _fields_ = [("channels", C.c_uint32), ("sample_rate", C.c_uint32), ("samples", C.c_uint32),
                ("lms", Lms * 8), ("brute", C.c_int), ("candidate_samples", C.c_uint64), ("slices", C.c_uint64)]


class FrameHeader(C.Structure):
    _fields_ = [("num_channels", C.c_uint32), ("sample_rate", C.c_uint32), ("num_samples_per_channel", C.c_uint32)]


class DecodeInfo(C.Structure):
    _fields_ = [("status", C.c_int32), ("streaming", C.c_int32), ("header_samples", C.c_uint32),
                ("channels", C.c_uint32), ("sample_rate", C.c_uint32), ("samples", C.c_uint64),
                ("frames", C.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.qoa_o_encoded_size.restype = C.c_size_t
        L.qoa_o_encoded_size.argtypes = [C.c_uint32, C.c_uint32]
        L.qoa_o_encoder_new.argtypes = [C.POINTER(Encoder), C.c_uint32, C.c_uint32, C.c_uint32]
        L.qoa_o_encoder_encode.argtypes = [C.POINTER(Encoder), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                           C.POINTER(C.c_size_t)]
        L.qoa_o_encoder_write_header.argtypes = [C.POINTER(Encoder), C.c_void_p]
        L.qoa_o_encoder_write_header.restype = None
        L.qoa_o_encoder_encode_frame.argtypes = [C.POINTER(Encoder), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                                 C.POINTER(C.c_size_t)]
        L.qoa_o_encode_all.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int,
                                       C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.qoa_o_decode_iter.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                        C.POINTER(DecodeInfo)]
        L.qoa_o_decode_iter.restype = None
        L.qoa_o_decode_all.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(DecodeInfo)]
        L.qoa_o_decode_all.restype = None
        L.qoa_o_stream_decoder_new.restype = C.c_void_p
        L.qoa_o_stream_decoder_free.argtypes = [C.c_void_p]
        L.qoa_o_stream_decoder_free.restype = None
        L.qoa_o_stream_decoder_decode_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                                        C.POINTER(C.c_size_t)]
        L.qoa_o_div.argtypes = [C.c_int32, C.c_int]
        L.qoa_o_div.restype = C.c_int32
        L.qoa_o_lms_predict.argtypes = [C.POINTER(Lms)]
        L.qoa_o_lms_predict.restype = C.c_int32
        L.qoa_o_lms_update.argtypes = [C.POINTER(Lms), C.c_int16, C.c_int32]
        L.qoa_o_lms_update.restype = None
        L.qoa_o_encode_slice.argtypes = [C.POINTER(Lms), C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int),
                                         C.POINTER(Lms), C.POINTER(C.c_uint64)]
        L.qoa_o_encode_slice.restype = C.c_uint64
        L.qoa_o_mt_encode.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int]
        L.qoa_o_mt_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        _lib = L
    return _lib


def _pcm(a):
    a = np.ascontiguousarray(a, dtype=np.int16)
    return a


def encoded_size(channels, samples):
    return lib().qoa_o_encoded_size(channels, samples)


def encode_all(pcm, channels, sample_rate, samples, brute=False, stats=None):
    """Returns (status, bytes)."""
    pcm = _pcm(pcm)
    enc = Encoder()
    st = lib().qoa_o_encoder_new(C.byref(enc), channels, sample_rate, samples)
    if st != OK:
        return st, b""
    enc.brute = 1 if brute else 0
    cap = encoded_size(channels, samples)
    out = np.empty(cap, dtype=np.uint8)
    n = C.c_size_t(0)
    st = lib().qoa_o_encoder_encode(C.byref(enc), pcm.ctypes.data, pcm.size, out.ctypes.data, cap, C.byref(n))
    if stats is not None:
        stats["candidate_samples"] = enc.candidate_samples
        stats["slices"] = enc.slices
    if st != OK:
        return st, b""
    return st, out[:n.value].tobytes()


class StreamEncoder:
    """QoaEncoder streaming API (write_header / encode_frame)."""

    def __init__(self, channels, sample_rate, samples, brute=False):
        self.enc = Encoder()
        self.status = lib().qoa_o_encoder_new(C.byref(self.enc), channels, sample_rate, samples)
        self.enc.brute = 1 if brute else 0

    def write_header(self):
        out = np.empty(8, dtype=np.uint8)
        lib().qoa_o_encoder_write_header(C.byref(self.enc), out.ctypes.data)
        return out.tobytes()

    def encode_frame(self, pcm):
        pcm = _pcm(pcm)
        cap = 8 + 16 * 8 + 8 * 256 * 8
        out = np.empty(cap, dtype=np.uint8)
        n = C.c_size_t(0)
        st = lib().qoa_o_encoder_encode_frame(C.byref(self.enc), pcm.ctypes.data, pcm.size, out.ctypes.data, cap,
                                              C.byref(n))
        return st, out[:n.value].tobytes()

    def lms(self, c):
        return list(self.enc.lms[c].history), list(self.enc.lms[c].weights)


def _bytes_arr(data):
    return np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data


def decode_iter(data, max_samples=None, max_headers=1 << 16):
    """QoaDecoder iteration.  Returns (info dict, samples np.int16, headers list)."""
    buf = _bytes_arr(data)
    if max_samples is None:
        max_samples = max(0, (len(buf) // 8 + 1) * 20)
    out = np.zeros(max_samples, dtype=np.int16)
    hdrs = (FrameHeader * max_headers)()
    info = DecodeInfo()
    lib().qoa_o_decode_iter(buf.ctypes.data if len(buf) else None, len(buf), out.ctypes.data, max_samples,
                            C.cast(hdrs, C.c_void_p), max_headers, C.byref(info))
    n = min(int(info.samples), max_samples)
    hl = [(h.num_channels, h.sample_rate, h.num_samples_per_channel) for h in hdrs[:min(int(info.frames), max_headers)]]
    return _info(info), out[:n], hl


def decode_all(data, max_samples=None):
    """decode_all semantics.  Returns (info dict, samples np.int16)."""
    buf = _bytes_arr(data)
    if max_samples is None:
        max_samples = max(0, (len(buf) // 8 + 1) * 20)
    out = np.zeros(max_samples, dtype=np.int16)
    info = DecodeInfo()
    lib().qoa_o_decode_all(buf.ctypes.data if len(buf) else None, len(buf), out.ctypes.data, max_samples,
                           C.byref(info))
    n = min(int(info.samples), max_samples)
    return _info(info), out[:n]


def _info(i):
    return dict(status=i.status, streaming=bool(i.streaming), header_samples=i.header_samples, channels=i.channels,
                sample_rate=i.sample_rate, samples=int(i.samples), frames=int(i.frames))


class StreamDecoder:
    def __init__(self):
        self.h = lib().qoa_o_stream_decoder_new()

    def decode_frame(self, frame):
        buf = _bytes_arr(frame)
        cap = (len(buf) // 8 + 1) * 20
        out = np.zeros(cap, dtype=np.int16)
        n = C.c_size_t(0)
        st = lib().qoa_o_stream_decoder_decode_frame(self.h, buf.ctypes.data if len(buf) else None, len(buf),
                                                     out.ctypes.data, cap, C.byref(n))
        return st, out[:n.value]

    def __del__(self):
        if getattr(self, "h", None):
            lib().qoa_o_stream_decoder_free(self.h)
            self.h = None


def mt_encode(pcm, n, channels, sample_rate, samples, threads):
    pcm = _pcm(pcm)
    out = np.empty(n * encoded_size(channels, samples), dtype=np.uint8)
    rc = lib().qoa_o_mt_encode(pcm.ctypes.data, n, channels, sample_rate, samples, out.ctypes.data, threads)
    return rc, out


def mt_decode(qoa, n, bytes_per_stream, out_stride, threads):
    qoa = np.ascontiguousarray(qoa, dtype=np.uint8)
    out = np.empty(n * out_stride, dtype=np.int16)
    rc = lib().qoa_o_mt_decode(qoa.ctypes.data, n, bytes_per_stream, out.ctypes.data, out_stride, threads)
    return rc, out
