"""Crate-shaped API: the names, argument meaning and error behaviour of rafaelcaricio/qoaudio (src/lib.rs), executed by
the B200 engine.  A user of the crate finds: QOA_* constants (:12-19), QoaDesc (:94-102), EncodeError (:105-115),
DecodeError (:866-873), ProcessingMode (:30-45), FrameHeader (:773-781), QoaItem (:655-659), encode_all (:760),
QoaEncoder::{new, encode, write_header, encode_frame} (:335-439), QoaDecoder::{new, open, new_streaming, decode_frame,
mode, current_frame_header, total_duration, into_inner} + Iterator (:127-201, :619-693), decode_all (:713),
open_and_decode_all (:750), DecodedQoa (:696-704) — plus *_batch variants, which are the point of this engine.
"""
import io
from dataclasses import dataclass

import numpy as np

from . import _native as N
from .engine import default_engine

QOA_SLICE_LEN = 20
QOA_LMS_LEN = 4
QOA_HEADER_SIZE = 8
QOA_MAGIC = 0x716F6166
QOA_SLICES_PER_FRAME = 256
QOA_FRAME_LEN = QOA_SLICES_PER_FRAME * QOA_SLICE_LEN
QOA_MAX_CHANNELS = 8
MAX_SLICES_PER_CHANNEL_PER_FRAME = 256


@dataclass
class QoaDesc:
    """src/lib.rs:94-102"""
    channels: int
    sample_rate: int
    samples: int


class EncodeError(Exception):
    """src/lib.rs:105-115; .kind is one of InvalidChannels, InvalidSampleRate, InvalidSamples, IoError."""
    _MSG = {"InvalidChannels": "Invalid number of channels (must be 1-8)",
            "InvalidSampleRate": "Invalid sample rate (must be > 0)",
            "InvalidSamples": "Invalid number of samples (must be > 0)"}

    def __init__(self, kind, detail=None):
        self.kind = kind
        super().__init__(self._MSG.get(kind, f"IO error: {detail}"))


class DecodeError(Exception):
    """src/lib.rs:866-873; .kind is one of NotQoaFile, NoSamples, InvalidFrameHeader, IncompatibleFrame, IoError."""
    _MSG = {"NotQoaFile": "File is not a qoa file", "NoSamples": "File has no samples",
            "InvalidFrameHeader": "File has invalid frame header", "IncompatibleFrame": "Incompatible frame header"}

    def __init__(self, kind, detail=None):
        self.kind = kind
        super().__init__(self._MSG.get(kind, f"IO error: {detail or 'failed to fill whole buffer'}"))


_ENC_KIND = {N.ENC_INVALID_CHANNELS: "InvalidChannels", N.ENC_INVALID_SAMPLE_RATE: "InvalidSampleRate",
             N.ENC_INVALID_SAMPLES: "InvalidSamples", N.ENC_IO_ERROR: "IoError"}
_DEC_KIND = {N.DEC_NOT_QOA_FILE: "NotQoaFile", N.DEC_NO_SAMPLES: "NoSamples",
             N.DEC_INVALID_FRAME_HEADER: "InvalidFrameHeader", N.DEC_INCOMPATIBLE_FRAME: "IncompatibleFrame",
             N.DEC_IO_ERROR: "IoError", N.DEC_OUTPUT_TOO_SMALL: "IoError"}


@dataclass(frozen=True)
class FrameHeader:
    """src/lib.rs:773-781"""
    num_channels: int = 0
    sample_rate: int = 0
    num_samples_per_channel: int = 0


class ProcessingMode:
    """src/lib.rs:30-45"""

    @dataclass(frozen=True)
    class FixedSamples:
        channels: int
        sample_rate: int
        samples: int

    @dataclass(frozen=True)
    class _Streaming:
        def __repr__(self):
            return "Streaming"

    Streaming = _Streaming()


class QoaItem:
    """src/lib.rs:655-659: either QoaItem.Sample(i16) or QoaItem.FrameHeader(FrameHeader)."""

    @dataclass(frozen=True)
    class Sample:
        value: int

    @dataclass(frozen=True)
    class FrameHeader:
        header: "FrameHeader"


@dataclass
class DecodedQoa:
    """src/lib.rs:696-704"""
    num_channels: int
    sample_rate: int
    samples: np.ndarray


def _as_pcm(samples):
    return np.ascontiguousarray(samples, dtype=np.int16).reshape(-1)


def _validate_desc(desc):
    """QoaEncoder::new, src/lib.rs:336-345 (this order)."""
    if desc.channels == 0 or desc.channels > QOA_MAX_CHANNELS:
        raise EncodeError("InvalidChannels")
    if desc.sample_rate == 0:
        raise EncodeError("InvalidSampleRate")
    if desc.samples == 0:
        raise EncodeError("InvalidSamples")


# ---- encode ------------------------------------------------------------------------------------------------------

def encode_all_batch(items, engine=None):
    """Batched encode_all: items = [(samples, QoaDesc), ...] -> list of bytes or EncodeError (one per stream)."""
    eng = engine or default_engine()
    status, outs = eng.encode_all_batch([_as_pcm(s) for s, _ in items],
                                        [(d.channels, d.sample_rate, d.samples) for _, d in items])
    return [o.tobytes() if st == N.OK else EncodeError(_ENC_KIND[st]) for st, o in zip(status, outs)]


def encode_all(sample_data, desc, engine=None):
    """src/lib.rs:760-763"""
    res = encode_all_batch([(sample_data, desc)], engine)[0]
    if isinstance(res, EncodeError):
        raise res
    return res


class QoaEncoder:
    """src/lib.rs:85-91, :333-439.  One stream; the LMS state lives on the GPU between encode_frame calls."""

    def __init__(self, desc, engine=None):
        _validate_desc(desc)
        self._desc = QoaDesc(desc.channels, desc.sample_rate, desc.samples)
        self._eng = engine or default_engine()
        self._batch = None

    new = classmethod(lambda cls, desc, engine=None: cls(desc, engine))

    def encode(self, sample_data):
        """src/lib.rs:367-398.  Like the reference, encoding starts from this encoder's CURRENT LMS state; for a
        fresh encoder that is the initial state and the result equals encode_all."""
        pcm = _as_pcm(sample_data)
        d = self._desc
        if pcm.size != d.samples * d.channels:
            raise EncodeError("InvalidSamples")
        st, outs = self._state().encode_chunk([pcm])
        if st[0] != N.OK:
            raise EncodeError(_ENC_KIND[st[0]])
        return b"qoaf" + int(d.samples).to_bytes(4, "big") + outs[0]

    def _state(self):
        if self._batch is None:
            d = self._desc
            self._batch = self._eng.encoder_batch([(d.channels, d.sample_rate, d.samples)])
        return self._batch

    def write_header(self, writer):
        """src/lib.rs:404-408"""
        s = self._desc.samples
        writer.write(b"qoaf" + int(s).to_bytes(4, "big"))

    def encode_frame(self, sample_data, writer):
        """src/lib.rs:418-439; returns frame_len, the samples per channel encoded (Ok(frame_len), :438)."""
        pcm = _as_pcm(sample_data)
        st, outs = self._state().encode_frame([pcm])
        if st[0] != N.OK:
            raise EncodeError(_ENC_KIND[st[0]])
        try:
            writer.write(outs[0])
        except OSError as e:
            raise EncodeError("IoError", str(e))
        return pcm.size // self._desc.channels


class QoaEncoderBatch:
    """A batch of QoaEncoders advancing in lock-step (one encode_frame per stream per call)."""

    def __init__(self, descs, engine=None):
        self._eng = engine or default_engine()
        self.descs = list(descs)
        self._batch = self._eng.encoder_batch([(d.channels, d.sample_rate, d.samples) for d in self.descs])
        self.errors = [None if st == N.OK else EncodeError(_ENC_KIND[st]) for st in self._batch.status]

    def write_header(self):
        return self._batch.write_header()

    def encode_frame(self, frames):
        st, outs = self._batch.encode_frame([None if f is None else _as_pcm(f) for f in frames])
        return [None if f is None else (o if s == N.OK else EncodeError(_ENC_KIND[s]))
                for f, s, o in zip(frames, st, outs)]


# ---- decode ------------------------------------------------------------------------------------------------------

def _read_all(reader):
    if isinstance(reader, (bytes, bytearray, memoryview, np.ndarray)):
        return bytes(reader)
    return reader.read()


class QoaDecoder:
    """src/lib.rs:59-68, :117-331, Iterator :661-693.  The whole stream is indexed and decoded on the GPU when the
    decoder is created; iteration replays the reference's item order (FrameHeader, then that frame's samples, ...)
    and raises the reference's error at the position the reference would."""

    def __init__(self, reader, engine=None, _streaming_new=False):
        self._eng = engine or default_engine()
        self._reader = reader
        self._frame_mode = _streaming_new
        self._current = FrameHeader()
        if _streaming_new:
            self._mode = ProcessingMode.Streaming
            self._items = iter(())
            self._first_header_pending = True
            return
        data = _read_all(reader)
        plan = self._eng.decode_plan([data], N.DECODE_ITER)
        try:
            info = plan.info[0]
            if info["status"] != N.OK and info["frames"] == 0:  # QoaDecoder::new failed (:127-169)
                raise DecodeError(_DEC_KIND[info["status"]])
            headers = plan.frame_headers(0)
            samples = plan.run()[0]
        finally:
            plan.close()
        if info["streaming"]:
            self._mode = ProcessingMode.Streaming
        else:
            self._mode = ProcessingMode.FixedSamples(info["channels"], info["sample_rate"], info["header_samples"])
            self._current = FrameHeader(*headers[0])
        self._info, self._headers, self._samples = info, headers, samples
        self._items = self._iterate()

    # constructors -------------------------------------------------------------------------------------------
    new = classmethod(lambda cls, reader, engine=None: cls(reader, engine))

    @classmethod
    def open(cls, path, engine=None):
        """src/lib.rs:621-624"""
        try:
            f = open(path, "rb")
        except OSError as e:
            raise DecodeError("IoError", str(e))
        with f:
            return cls(io.BytesIO(f.read()), engine)

    @classmethod
    def new_streaming(cls, engine=None):
        """src/lib.rs:632-638"""
        return cls(None, engine, _streaming_new=True)

    # accessors ----------------------------------------------------------------------------------------------
    def mode(self):
        return self._mode

    def current_frame_header(self):
        return self._current

    def total_duration(self):
        """src/lib.rs:190-201: seconds as float, or None in streaming mode."""
        if isinstance(self._mode, ProcessingMode.FixedSamples):
            return self._mode.samples / self._mode.sample_rate
        return None

    def into_inner(self):
        return self._reader

    # iteration ----------------------------------------------------------------------------------------------
    def _iterate(self):
        info = self._info
        if info["streaming"]:
            yield QoaItem.FrameHeader(FrameHeader())  # :674-677 default header first
        pos = 0
        for (ch, rate, nspc) in self._headers:
            self._current = FrameHeader(ch, rate, nspc)
            yield QoaItem.FrameHeader(self._current)
            n = min(nspc * ch, self._samples.size - pos)
            for v in self._samples[pos:pos + n]:
                yield QoaItem.Sample(int(v))
            pos += n
        if info["status"] != N.OK:
            raise DecodeError(_DEC_KIND[info["status"]])

    def __iter__(self):
        return self

    def __next__(self):
        return next(self._items)

    # streaming ----------------------------------------------------------------------------------------------
    def decode_frame(self, frame_data):
        """src/lib.rs:641-651: decode one (or more) whole frames given as bytes; returns the samples.
        As in the reference, after an error the decoder must not be used again.

        The reference appends `frame_data` to a persistent Cursor and iterates to EOF.  Because every QOA frame carries
        its own LMS state (:270-281) and Streaming mode never raises IncompatibleFrame (:255-258), the only state that
        survives a successful call is the cursor position, which then sits at the end of the buffer: a call whose bytes
        are whole frames therefore decodes exactly as a stateless call does, which is what this method is.  Bytes that
        are NOT whole frames behave as in the reference, too: a frame cut after its 8-byte header is IoError
        (UnexpectedEof, and the reference decoder is unusable afterwards, :666-667), and 1-7 trailing bytes are consumed
        silently (read_exact on a Cursor moves to the end on a short read); neither implementation can resume a frame that
        was split across two calls."""
        if not self._frame_mode:
            raise TypeError("decode_frame is only available on QoaDecoder.new_streaming()")
        infos, outs = self._eng.decode_batch([bytes(frame_data)], N.DECODE_FRAMES)
        if infos[0]["status"] != N.OK:
            raise DecodeError(_DEC_KIND[infos[0]["status"]])
        return outs[0]


def decode_all_batch(readers, engine=None):
    """Batched decode_all (src/lib.rs:713-747): list of DecodedQoa or DecodeError, one per stream."""
    eng = engine or default_engine()
    infos, outs = eng.decode_batch([_read_all(r) for r in readers], N.DECODE_ALL)
    res = []
    for info, o in zip(infos, outs):
        if info["status"] != N.OK:
            res.append(DecodeError(_DEC_KIND[info["status"]]))
        else:
            res.append(DecodedQoa(info["channels"], info["sample_rate"], o))
    return res


def decode_all(reader, engine=None):
    """src/lib.rs:713-747"""
    res = decode_all_batch([reader], engine)[0]
    if isinstance(res, DecodeError):
        raise res
    return res


def open_and_decode_all(path, engine=None):
    """src/lib.rs:750-754"""
    try:
        with open(path, "rb") as f:
            data = f.read()
    except OSError as e:
        raise DecodeError("IoError", str(e))
    return decode_all(data, engine)
