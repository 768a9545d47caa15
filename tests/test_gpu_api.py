"""The reference crate's own unit tests (src/lib.rs:991-1298) re-expressed against the crate-shaped Python API, plus
the two fuzz targets as seeded randomised tests.  GPU only."""
import io
import os
import math

import numpy as np
import pytest

import signals
from oracle import binding as ob

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def q():
    import qoaudio_b200
    qoaudio_b200.default_engine()
    return qoaudio_b200


def test_iterating_through_whole_file(q, fixture_bytes):
    """src/lib.rs:998-1032"""
    dec = q.QoaDecoder(io.BytesIO(fixture_bytes))
    assert dec.mode() == q.ProcessingMode.FixedSamples(2, 44100, 2394122)
    assert abs(dec.total_duration() - 2394122 / 44100) < 1e-12
    frames, samples = [], 0
    for item in dec:
        if isinstance(item, q.QoaItem.FrameHeader):
            frames.append(item.header)
        else:
            samples += 1
    assert len(frames) == 468
    for h in frames[:-1]:
        assert (h.num_channels, h.sample_rate, h.num_samples_per_channel) == (2, 44100, 5120)
    assert frames[-1].num_samples_per_channel == 3082
    assert samples == 2394122 * 2
    assert dec.current_frame_header() == frames[-1]


def test_item_order_and_values_small(q):
    pcm = signals.synth_pcm(1, 0, 3, 5130)
    data = q.encode_all(pcm, q.QoaDesc(3, 22050, 5130))
    info, opcm, ohdrs = ob.decode_iter(data)
    items = list(q.QoaDecoder(data))
    assert isinstance(items[0], q.QoaItem.FrameHeader) and items[0].header.num_samples_per_channel == 5120
    assert isinstance(items[1], q.QoaItem.Sample)
    assert isinstance(items[1 + 5120 * 3], q.QoaItem.FrameHeader)
    vals = np.array([it.value for it in items if isinstance(it, q.QoaItem.Sample)], dtype=np.int16)
    assert np.array_equal(vals, opcm)
    hdrs = [(it.header.num_channels, it.header.sample_rate, it.header.num_samples_per_channel) for it in items
            if isinstance(it, q.QoaItem.FrameHeader)]
    assert hdrs == ohdrs


def test_decode_streaming_frames(q, fixture_bytes):
    """src/lib.rs:1034-1066"""
    dec = q.QoaDecoder.new_streaming()
    assert dec.mode() is q.ProcessingMode.Streaming and dec.total_duration() is None
    pos = 8
    for _ in range(2):
        size = int.from_bytes(fixture_bytes[pos + 6:pos + 8], "big")
        out = dec.decode_frame(fixture_bytes[pos:pos + size])
        assert len(out) == 5120 * 2
        pos += size


def test_decode_all(q, fixture_bytes):
    """src/lib.rs:1068-1074"""
    d = q.decode_all(io.BytesIO(fixture_bytes))
    assert (d.sample_rate, d.num_channels, len(d.samples)) == (44100, 2, 2394122 * 2)


@pytest.mark.parametrize("channels", [1, 2])
def test_encode_decode_simple(q, channels):
    """src/lib.rs:1076-1146: 0.1 s 440 Hz sine/cosine."""
    rate, n = 44100, 4410
    t = np.arange(n) / rate
    cols = [np.sin(2 * math.pi * 440 * t) * 16384, np.cos(2 * math.pi * 440 * t) * 16384][:channels]
    pcm = np.stack(cols, axis=1).astype(np.int16).reshape(-1)
    enc = q.encode_all(pcm, q.QoaDesc(channels, rate, n))
    assert enc == ob.encode_all(pcm, channels, rate, n)[1]
    d = q.decode_all(enc)
    assert (d.sample_rate, d.num_channels, len(d.samples)) == (rate, channels, n * channels)


def test_encoder_errors(q):
    """src/lib.rs:1148-1183"""
    for desc, kind in ((q.QoaDesc(0, 44100, 1000), "InvalidChannels"), (q.QoaDesc(1, 0, 1000), "InvalidSampleRate"),
                       (q.QoaDesc(1, 44100, 0), "InvalidSamples"), (q.QoaDesc(9, 0, 0), "InvalidChannels")):
        with pytest.raises(q.EncodeError) as e:
            q.QoaEncoder(desc)
        assert e.value.kind == kind
        with pytest.raises(q.EncodeError) as e:
            q.encode_all(np.zeros(10, np.int16), desc)
        assert e.value.kind == kind
    with pytest.raises(q.EncodeError) as e:
        q.encode_all(np.zeros(10, np.int16), q.QoaDesc(1, 44100, 11))
    assert e.value.kind == "InvalidSamples"


def test_round_trip_audio(q):
    """src/lib.rs:1185-1230: sawtooth, max abs diff < 8000 (the reference measures 5665)."""
    pcm = np.array([((i % 200) - 100) * 100 for i in range(1000)], dtype=np.int16)
    d = q.decode_all(q.encode_all(pcm, q.QoaDesc(1, 44100, 1000)))
    diff = np.abs(d.samples.astype(np.int32) - pcm.astype(np.int32)).max()
    assert diff < 8000 and diff == 5665


def test_full_file_round_trip(q, fixture_bytes):
    """src/lib.rs:1232-1259: RMS error < 500."""
    d = q.decode_all(fixture_bytes)
    enc = q.encode_all(d.samples, q.QoaDesc(d.num_channels, d.sample_rate, len(d.samples) // d.num_channels))
    d2 = q.decode_all(enc)
    assert len(d2.samples) == len(d.samples)
    rms = math.sqrt(np.mean((d.samples.astype(np.float64) - d2.samples.astype(np.float64)) ** 2))
    assert rms < 500


def test_streaming_encode(q, fixture_bytes):
    """src/lib.rs:1261-1297: write_header + encode_frame bytes == one-shot encode bytes."""
    d = q.decode_all(fixture_bytes)
    pcm = d.samples[:5120 * 2 * 5 + 2 * 1234]
    n = len(pcm) // 2
    desc = q.QoaDesc(2, 44100, n)
    one_shot = q.QoaEncoder(desc).encode(pcm)
    assert one_shot == q.encode_all(pcm, desc)
    enc = q.QoaEncoder(desc)
    out = io.BytesIO()
    enc.write_header(out)
    total = 0
    for i in range(0, len(pcm), 5120 * 2):
        total += enc.encode_frame(pcm[i:i + 5120 * 2], out)  # Ok(frame_len): samples per channel (src/lib.rs:438)
    assert out.getvalue() == one_shot and total == n
    with pytest.raises(q.EncodeError):
        enc.encode_frame(np.zeros(3, np.int16), io.BytesIO())


def test_decoder_errors(q):
    with pytest.raises(q.DecodeError) as e:
        q.QoaDecoder(b"nope....")
    assert e.value.kind == "NotQoaFile"
    with pytest.raises(q.DecodeError) as e:
        q.QoaDecoder(b"qoaf\0\0\0\5")
    assert e.value.kind == "NoSamples"
    with pytest.raises(q.DecodeError) as e:
        q.decode_all(b"qo")
    assert e.value.kind == "IoError"
    good = q.encode_all(signals.synth_pcm(1, 0, 2, 100), q.QoaDesc(2, 8000, 100))
    dec = q.QoaDecoder(good[:-5])
    with pytest.raises(q.DecodeError) as e:
        list(dec)
    assert e.value.kind == "IoError"
    res = q.decode_all_batch([good, good[:-5], b"xxxxxxxxxxxx"])
    assert isinstance(res[0], q.DecodedQoa) and res[1].kind == "IoError" and res[2].kind == "NotQoaFile"


def test_fuzz_iter_all(q):
    """fuzz/fuzz_targets/iter_all.rs: arbitrary bytes never crash the decoder; results equal the oracle's."""
    rng = np.random.default_rng(1234)
    base = q.encode_all(signals.synth_pcm(5, 1, 2, 300), q.QoaDesc(2, 44100, 300))
    cases = []
    for k in range(200):
        b = bytearray(base)
        for _ in range(int(rng.integers(1, 6))):
            b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        if k % 3 == 0:
            b = b[:int(rng.integers(0, len(b)))]
        cases.append(bytes(b))
    cases += [bytes(rng.integers(0, 256, size=int(rng.integers(0, 200)), dtype=np.uint8)) for _ in range(50)]
    eng = q.default_engine()
    infos, outs = eng.decode_batch(cases, 0)
    for c, info, o in zip(cases, infos, outs):
        oi, opcm, _ = ob.decode_iter(c)
        assert {k: info[k] for k in oi} == oi
        assert np.array_equal(o, opcm)


def test_fuzz_encode_round_trip(q):
    """fuzz/fuzz_targets/encode_round_trip.rs: ch = data[0] % 8 + 1, LE i16 from the rest, encode -> decode ok."""
    rng = np.random.default_rng(99)
    items, shapes = [], []
    for _ in range(64):
        data = rng.integers(0, 256, size=int(rng.integers(3, 4000)), dtype=np.uint8).tobytes()
        ch = data[0] % 8 + 1
        body = data[1:]
        vals = np.frombuffer(body[:len(body) // 2 * 2], dtype="<i2")
        n = len(vals) // ch
        if n == 0:
            continue
        pcm = vals[:n * ch]
        items.append((pcm, q.QoaDesc(ch, 44100, n)))
        shapes.append((ch, n))
    encs = q.encode_all_batch(items)
    decs = q.decode_all_batch(encs)
    for (pcm, desc), enc, dec, (ch, n) in zip(items, encs, decs, shapes):
        assert enc == ob.encode_all(pcm, ch, 44100, n)[1]
        assert (dec.num_channels, dec.sample_rate, len(dec.samples)) == (ch, 44100, n * ch)


def test_example_tools_wav_to_qoa_and_back(q, tmp_path):
    """examples/encode.py / decode.py (the reference's examples/encode.rs, decode.rs): single-file and batch forms."""
    import subprocess
    import sys as _sys
    from conftest import ROOT
    shapes = [(1, 8000, 777), (2, 44100, 5120 * 2 + 33), (6, 48000, 6000)]
    wavs, pcms = [], []
    for i, (ch, rate, n) in enumerate(shapes):
        pcm = signals.synth_pcm(31, i, ch, n)
        path = tmp_path / f"in{i}.wav"
        q.wavio.write_wav(path, ch, rate, pcm)
        wavs.append(str(path))
        pcms.append(pcm)
    (tmp_path / "bad.wav").write_bytes(b"RIFF" + bytes(20))
    enc = [_sys.executable, os.path.join(ROOT, "examples", "encode.py")]
    dec = [_sys.executable, os.path.join(ROOT, "examples", "decode.py")]
    r = subprocess.run(enc + ["--out-dir", str(tmp_path / "q")] + wavs + [str(tmp_path / "bad.wav"), str(tmp_path / "nope.wav")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 1 and "does not exist" in r.stderr and "cannot read" in r.stderr, r.stderr
    qoas = []
    for i, (ch, rate, n) in enumerate(shapes):
        data = (tmp_path / "q" / f"in{i}.qoa").read_bytes()
        assert data == ob.encode_all(pcms[i], ch, rate, n)[1]
        qoas.append(str(tmp_path / "q" / f"in{i}.qoa"))
    r = subprocess.run(dec + ["--out-dir", str(tmp_path / "w")] + qoas, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    for i, (ch, rate, n) in enumerate(shapes):
        got = q.wavio.read_wav(tmp_path / "w" / f"in{i}.wav")
        assert got[:3] == (ch, rate, 16)
        assert np.array_equal(got[3], ob.decode_all((tmp_path / "q" / f"in{i}.qoa").read_bytes())[1])
    r = subprocess.run(enc + [wavs[1], str(tmp_path / "one.qoa")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and (tmp_path / "one.qoa").read_bytes() == (tmp_path / "q" / "in1.qoa").read_bytes()
    r = subprocess.run(dec + [str(tmp_path / "one.qoa"), str(tmp_path / "one.wav")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and (tmp_path / "one.wav").read_bytes() == (tmp_path / "w" / "in1.wav").read_bytes()
