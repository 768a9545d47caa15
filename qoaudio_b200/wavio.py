"""16-bit PCM WAV files for the example tools (the reference's examples use the `hound` crate for this,
examples/encode.rs:32-58, examples/decode.rs:39-50).  Reads RIFF/WAVE with a plain PCM or WAVE_FORMAT_EXTENSIBLE(PCM)
`fmt ` chunk; writes the canonical 44-byte header (EXTENSIBLE above two channels, as hound does)."""
import struct

import numpy as np

_PCM, _EXTENSIBLE = 1, 0xFFFE
_PCM_GUID = bytes.fromhex("0100000000001000800000aa00389b71")


class WavError(ValueError):
    pass


def read_wav(path):
    """Returns (channels, sample_rate, bits_per_sample, samples) with samples an interleaved np.int16 array; files that
    are not 16-bit integer PCM come back with samples = None so that the caller can report the bit depth."""
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 12 or data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise WavError("not a RIFF/WAVE file")
    pos, fmt, pcm = 12, None, None
    while pos + 8 <= len(data):
        tag, size = data[pos:pos + 4], struct.unpack_from("<I", data, pos + 4)[0]
        body = data[pos + 8:pos + 8 + size]
        if tag == b"fmt ":
            if len(body) < 16:
                raise WavError("short fmt chunk")
            fmt = struct.unpack_from("<HHIIHH", body)
            if fmt[0] == _EXTENSIBLE and len(body) >= 40:
                fmt = ((_PCM if body[24:40] == _PCM_GUID else 0),) + fmt[1:]
        elif tag == b"data":
            pcm = body  # a truncated final chunk yields what is there
        pos += 8 + size + (size & 1)
    if fmt is None or pcm is None:
        raise WavError("missing fmt or data chunk")
    tag, channels, rate, _, _, bits = fmt
    if tag != _PCM or bits != 16:
        return channels, rate, bits if tag == _PCM else 0, None
    return channels, rate, bits, np.frombuffer(pcm[:len(pcm) & ~1], dtype="<i2").astype(np.int16, copy=False)


def write_wav(path, channels, sample_rate, samples):
    """samples: interleaved int16."""
    pcm = np.ascontiguousarray(samples, dtype="<i2").tobytes()
    block = 2 * channels
    if channels <= 2:
        fmt = struct.pack("<HHIIHH", _PCM, channels, sample_rate, sample_rate * block, block, 16)
    else:
        mask = (1 << channels) - 1
        fmt = struct.pack("<HHIIHHHHI", _EXTENSIBLE, channels, sample_rate, sample_rate * block, block, 16, 22, 16, mask)
        fmt += _PCM_GUID
    chunks = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"data" + struct.pack("<I", len(pcm)) + pcm
    if len(pcm) & 1:
        chunks += b"\0"
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", len(chunks)) + chunks)
