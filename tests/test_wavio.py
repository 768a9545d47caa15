"""16-bit PCM WAV reader/writer used by the example tools (CPU only), checked against the stdlib `wave` module."""
import struct
import wave

import numpy as np
import pytest

from qoaudio_b200 import wavio


@pytest.mark.parametrize("channels", [1, 2, 6, 8])
def test_write_then_read(tmp_path, channels):
    rng = np.random.default_rng(channels)
    pcm = rng.integers(-32768, 32768, size=channels * 1001, dtype=np.int16)
    path = tmp_path / "a.wav"
    wavio.write_wav(path, channels, 44100, pcm)
    ch, rate, bits, got = wavio.read_wav(path)
    assert (ch, rate, bits) == (channels, 44100, 16) and np.array_equal(got, pcm)
    with wave.open(str(path), "rb") as w:  # an independent reader agrees (3.12 reads WAVE_FORMAT_EXTENSIBLE)
        assert (w.getnchannels(), w.getframerate(), w.getsampwidth(), w.getnframes()) == (channels, 44100, 2, 1001)
        assert w.readframes(1001) == pcm.astype("<i2").tobytes()


def test_reads_stdlib_output_and_skips_unknown_chunks(tmp_path):
    pcm = np.arange(-300, 300, dtype=np.int16)
    path = tmp_path / "b.wav"
    with wave.open(str(path), "wb") as w:
        w.setnchannels(2)
        w.setsampwidth(2)
        w.setframerate(8000)
        w.writeframes(pcm.tobytes())
    raw = path.read_bytes()
    junk = b"LIST" + struct.pack("<I", 5) + b"hello" + b"\0"  # odd-sized chunk + pad byte before fmt
    patched = raw[:12] + junk + raw[12:]
    patched = patched[:4] + struct.pack("<I", len(patched) - 8) + patched[8:]
    path.write_bytes(patched)
    ch, rate, bits, got = wavio.read_wav(path)
    assert (ch, rate, bits) == (2, 8000, 16) and np.array_equal(got, pcm)


def test_rejects_other_formats(tmp_path):
    path = tmp_path / "c.wav"
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(3)
        w.setframerate(8000)
        w.writeframes(bytes(30))
    assert wavio.read_wav(path)[2:] == (24, None)  # the tool reports the bit depth, as examples/encode.rs:36-42 does
    path.write_bytes(b"RIFX" + bytes(40))
    with pytest.raises(wavio.WavError):
        wavio.read_wav(path)
