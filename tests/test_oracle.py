"""CPU tests pinning the oracle (oracle/qoa_oracle.c) — re-expressions of the reference's own tests
(src/lib.rs:991-1298), the SURVEY Appendix B known answers, and cross-checks between the three restatements."""
import hashlib
import math

import numpy as np
import pytest

import signals
from oracle import binding as ob
from oracle import qoa_pyref as py


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


# ---- reference test re-expressions --------------------------------------------------------------------------
def test_iterating_through_whole_file(fixture_bytes):  # src/lib.rs:998-1032
    info, pcm, hdrs = ob.decode_iter(fixture_bytes)
    assert info["status"] == ob.OK and not info["streaming"]
    assert (info["channels"], info["sample_rate"], info["header_samples"]) == (2, 44100, 2394122)
    assert len(hdrs) == 468
    for k, (ch, rate, nspc) in enumerate(hdrs):
        assert ch == 2 and rate == 44100
        assert nspc == (5120 if k < 467 else 3082)
    assert info["samples"] == 2394122 * 2 == len(pcm)


def test_bench_item_count(fixture_bytes):  # benches/qoa_benchmarks.rs:11-12
    info, _, _ = ob.decode_iter(fixture_bytes, max_samples=0)
    assert info["samples"] + info["frames"] == 2394122 * 2 + 468


def test_decode_streaming_frames(fixture_bytes):  # src/lib.rs:1034-1066
    dec = ob.StreamDecoder()
    pos = 8
    whole = []
    for _ in range(2):
        fsize = int.from_bytes(fixture_bytes[pos + 6:pos + 8], "big")
        st, s = dec.decode_frame(fixture_bytes[pos:pos + fsize])
        assert st == ob.OK and len(s) == 5120 * 2
        whole.append(s)
        pos += fsize
    _, pcm, _ = ob.decode_iter(fixture_bytes)
    assert np.array_equal(np.concatenate(whole), pcm[:2 * 5120 * 2])


def test_decode_all(fixture_bytes, golden):  # src/lib.rs:1068-1074
    info, pcm = ob.decode_all(fixture_bytes)
    assert info["status"] == ob.OK and info["sample_rate"] == 44100 and info["channels"] == 2
    assert len(pcm) == 2394122 * 2
    assert sha(pcm.tobytes()) == golden["appendix_b"]["fixture_decoded_pcm_sha256"]


@pytest.mark.parametrize("channels", [1, 2])
def test_encode_decode_sine(channels):  # src/lib.rs:1076-1146
    rate = 44100
    n = int(rate * 0.1)
    t = np.arange(n) / rate
    cols = [np.sin(t * 440.0 * 2 * math.pi) * 16384.0, np.cos(t * 440.0 * 2 * math.pi) * 16384.0][:channels]
    pcm = np.stack(cols, axis=1).astype(np.int16).reshape(-1)  # `as i16` truncates toward zero like astype
    st, enc = ob.encode_all(pcm, channels, rate, n)
    assert st == ob.OK and len(enc) > 8
    info, dec = ob.decode_all(enc)
    assert (info["status"], info["sample_rate"], info["channels"], len(dec)) == (ob.OK, rate, channels, len(pcm))


def test_encoder_errors():  # src/lib.rs:1148-1183 and precedence :336-344
    z = np.zeros(1000, dtype=np.int16)
    assert ob.encode_all(z, 0, 44100, 1000)[0] == ob.ENC_INVALID_CHANNELS
    assert ob.encode_all(z, 9, 44100, 1000)[0] == ob.ENC_INVALID_CHANNELS
    assert ob.encode_all(z, 1, 0, 1000)[0] == ob.ENC_INVALID_SAMPLE_RATE
    assert ob.encode_all(z, 1, 44100, 0)[0] == ob.ENC_INVALID_SAMPLES
    assert ob.encode_all(z, 0, 0, 0)[0] == ob.ENC_INVALID_CHANNELS
    assert ob.encode_all(z, 1, 0, 0)[0] == ob.ENC_INVALID_SAMPLE_RATE
    assert ob.encode_all(z, 1, 44100, 999)[0] == ob.ENC_INVALID_SAMPLES  # len != samples*channels :368
    e = ob.StreamEncoder(2, 44100, 10)
    assert e.encode_frame(np.zeros(0, dtype=np.int16))[0] == ob.ENC_INVALID_SAMPLES  # :424
    assert e.encode_frame(np.zeros(3, dtype=np.int16))[0] == ob.ENC_INVALID_SAMPLES
    assert e.encode_frame(np.zeros(5121 * 2, dtype=np.int16))[0] == ob.ENC_INVALID_SAMPLES  # :428
    assert e.encode_frame(np.zeros(5120 * 2, dtype=np.int16))[0] == ob.OK


def test_round_trip_audio():  # src/lib.rs:1185-1230
    saw, ch, rate, n = signals.kat_signals()["saw1000_mono"]
    st, enc = ob.encode_all(saw, ch, rate, n)
    info, dec = ob.decode_all(enc)
    assert len(dec) == len(saw)
    assert int(np.abs(saw.astype(np.int32) - dec).max()) < 8000


def test_full_file_round_trip_and_streaming_encode(fixture_bytes, golden):  # src/lib.rs:1232-1297
    _, pcm = ob.decode_all(fixture_bytes)
    st, enc = ob.encode_all(pcm, 2, 44100, 2394122)
    assert st == ob.OK and sha(enc) == golden["appendix_b"]["fixture_reencode_sha256"]
    info, dec = ob.decode_all(enc)
    assert len(dec) == len(pcm)
    rms = float(np.sqrt(np.mean((pcm.astype(np.float64) - dec) ** 2)))
    assert rms < 500.0
    assert sha(dec.tobytes()) == golden["appendix_b"]["fixture_reencode_decoded_sha256"]
    # streaming == one-shot (:1261-1297)
    e = ob.StreamEncoder(2, 44100, 2394122)
    parts = [e.write_header()]
    for off in range(0, 2394122, 5120):
        n = min(5120, 2394122 - off)
        st, b = e.encode_frame(pcm[off * 2:(off + n) * 2])
        assert st == ob.OK
        parts.append(b)
    assert b"".join(parts) == enc


# ---- known answers ----------------------------------------------------------------------------------------------
def test_appendix_b_kats(golden):
    for name, (sig, ch, rate, n) in signals.kat_signals().items():
        exp = golden["appendix_b"]["kats"][name]
        for brute in (False, True):
            st, enc = ob.encode_all(sig, ch, rate, n, brute=brute)
            assert st == ob.OK and len(enc) == exp["len"] and sha(enc) == exp["sha256"], (name, brute)


def test_silence_is_not_zero():  # DEQUANT[0][0] = +1: SURVEY §8a note
    st, enc = ob.encode_all(np.zeros(21, dtype=np.int16), 1, 8000, 21)
    assert enc[32:40].hex() == "00d10d0200051451"


def test_synth_generator_pins(golden):
    for key, exp in golden["synth"].items():
        kind, seed, stream, ch, n = key.split(":")
        p = signals.synth_pcm(int(seed), int(stream), int(ch), int(n), kind)
        assert sha(p.tobytes()) == exp["pcm_sha256"], key
        st, enc = ob.encode_all(p, int(ch), 44100, int(n))
        assert sha(enc) == exp["qoa_sha256"] and len(enc) == exp["qoa_len"]


# ---- restatement cross-checks -------------------------------------------------------------------------------------
LENGTHS = [1, 2, 19, 20, 21, 39, 40, 41, 5119, 5120, 5121, 10259]


@pytest.mark.parametrize("channels", range(1, 9))
def test_faithful_equals_brute_force(channels):
    rng = np.random.default_rng(channels)
    for n in LENGTHS:
        for kind in ("music", "noise", "zeros", "const", "alt"):
            if kind in ("music", "noise"):
                pcm = signals.synth_pcm(3, n, channels, n, kind)
            elif kind == "zeros":
                pcm = np.zeros(n * channels, dtype=np.int16)
            elif kind == "const":
                pcm = np.repeat(rng.integers(-32768, 32768, size=channels).astype(np.int16)[None, :], n, 0).reshape(-1)
            else:
                pcm = np.where(np.arange(n * channels) // channels % 2 == 0, 32767, -32768).astype(np.int16)
            a = ob.encode_all(pcm, channels, 48000, n)
            b = ob.encode_all(pcm, channels, 48000, n, brute=True)
            assert a[0] == ob.OK and a == b, (channels, n, kind)
            info, dec = ob.decode_all(a[1])
            assert info["status"] == ob.OK and len(dec) == n * channels and info["frames"] == -(-n // 5120)


def test_c_oracle_equals_python_restatement():
    cases = [(1, 61), (2, 45), (3, 41), (5, 23), (8, 20)]
    for ch, n in cases:
        for name, sig in signals.edge_signals().items():
            pcm = np.resize(sig, n * ch).astype(np.int16)
            st, enc = ob.encode_all(pcm, ch, 22050, n)
            assert st == ob.OK
            assert py.encode_all([int(v) for v in pcm], ch, 22050, n) == enc, (ch, n, name)
            pch, prate, pdec, pframes = py.decode_frames(enc)
            info, dec = ob.decode_all(enc)
            assert (pch, prate, pframes) == (ch, 22050, [n]) and pdec == dec.tolist()


def test_python_restatement_multi_frame():
    pcm = signals.synth_pcm(1, 0, 2, 5120 + 33, "music")
    st, enc = ob.encode_all(pcm, 2, 44100, 5120 + 33)
    assert py.encode_all([int(v) for v in pcm], 2, 44100, 5120 + 33) == enc
    assert py.decode_frames(enc)[2] == ob.decode_all(enc)[1].tolist()


def test_primitives_against_python():
    rng = np.random.default_rng(0)
    for _ in range(2000):
        v = int(rng.integers(-(1 << 31), 1 << 31))
        sf = int(rng.integers(0, 16))
        assert ob.lib().qoa_o_div(v, sf) == py.qoa_div(v, sf)
    for _ in range(500):
        l = ob.Lms()
        h = [int(x) for x in rng.integers(-32768, 32768, size=4)]
        w = [int(x) for x in rng.integers(-(1 << 31), 1 << 31, size=4)]
        for i in range(4):
            l.history[i], l.weights[i] = h[i], w[i]
        assert ob.lib().qoa_o_lms_predict(l) == py.predict(h, w)
        s, r = int(rng.integers(-32768, 32768)), int(rng.integers(-14336, 14337))
        ob.lib().qoa_o_lms_update(l, s, r)
        nh, nw = py.update(h, w, s, r)
        assert list(l.history) == nh and list(l.weights) == nw


def test_pruning_work_profile(fixture_bytes):
    """The faithful search does a small fraction of brute-force work (~13 % + first pass)."""
    _, pcm = ob.decode_all(fixture_bytes)
    n = 5120 * 20
    s1, s2 = {}, {}
    ob.encode_all(pcm[:n * 2], 2, 44100, n, stats=s1)
    ob.encode_all(pcm[:n * 2], 2, 44100, n, brute=True, stats=s2)
    assert s2["candidate_samples"] == 16 * n * 2
    assert s1["candidate_samples"] < 0.3 * s2["candidate_samples"]


# ---- decoder error semantics (src/lib.rs:127-169, :205-289, :713-747) ---------------------------------------------
def _mk(ch=2, n=45, rate=8000):
    pcm = signals.synth_pcm(5, 1, ch, n)
    return ob.encode_all(pcm, ch, rate, n)[1]


def test_decoder_errors():
    good = _mk()
    assert ob.decode_iter(b"")[0]["status"] == ob.DEC_IO_ERROR
    assert ob.decode_iter(b"qoa")[0]["status"] == ob.DEC_IO_ERROR
    assert ob.decode_iter(b"riff\0\0\0\1")[0]["status"] == ob.DEC_NOT_QOA_FILE
    assert ob.decode_iter(b"qoaf\0\0")[0]["status"] == ob.DEC_IO_ERROR
    assert ob.decode_iter(good[:8])[0]["status"] == ob.DEC_NO_SAMPLES
    assert ob.decode_iter(good[:15])[0]["status"] == ob.DEC_NO_SAMPLES  # short header read => Ok(false) => NoSamples
    assert ob.decode_iter(good[:16])[0]["status"] == ob.DEC_IO_ERROR    # EOF inside LMS
    assert ob.decode_iter(good[:47])[0]["status"] == ob.DEC_IO_ERROR
    # EOF inside the slices: groups before it are yielded, then IoError
    info, pcm, hdrs = ob.decode_iter(good[:48 + 16 + 8])
    assert info["status"] == ob.DEC_IO_ERROR and info["samples"] == 40 and len(hdrs) == 1
    full = ob.decode_iter(good)[1]
    assert np.array_equal(pcm, full[:40])
    # trailing garbage shorter than a header is a clean end
    info, pcm, _ = ob.decode_iter(good + b"\1\2\3")
    assert info["status"] == ob.OK and info["samples"] == 90

    def hdr(ch, rate, nspc, fsize):
        return ((ch << 56) | (rate << 32) | (nspc << 16) | fsize).to_bytes(8, "big")

    base = good[:8]
    body = good[16:]
    assert ob.decode_iter(base + hdr(0, 8000, 45, 88) + body)[0]["status"] == ob.DEC_INVALID_FRAME_HEADER
    assert ob.decode_iter(base + hdr(2, 0, 45, 88) + body)[0]["status"] == ob.DEC_INVALID_FRAME_HEADER
    assert ob.decode_iter(base + hdr(2, 8000, 45, 40) + body)[0]["status"] == ob.DEC_INVALID_FRAME_HEADER  # <= 8+16ch
    assert ob.decode_iter(base + hdr(2, 8000, 45, 40 + 24) + body)[0]["status"] == ob.DEC_INVALID_FRAME_HEADER  # 3 % 2
    assert ob.decode_iter(base + hdr(2, 8000, 45, 40 + 16 * 257) + body)[0]["status"] == ob.DEC_INVALID_FRAME_HEADER
    # frame_size is only validated, not used to walk: a wrong-but-valid size still decodes
    assert ob.decode_iter(base + hdr(2, 8000, 45, 40 + 16) + body)[0]["status"] == ob.OK
    # second frame with different channels/rate in Fixed mode => IncompatibleFrame, first frame's samples yielded
    other = _mk(ch=2, n=20, rate=16000)
    info, pcm, hdrs = ob.decode_iter(good + other[8:])
    assert info["status"] == ob.DEC_INCOMPATIBLE_FRAME and info["samples"] == 90 and len(hdrs) == 1
    same = _mk(ch=2, n=20, rate=8000)
    info, pcm, hdrs = ob.decode_iter(good + same[8:])  # partial interior frame accepted (:53-57)
    assert info["status"] == ob.OK and info["samples"] == 90 + 40 and len(hdrs) == 2


def test_streaming_mode_semantics():
    good = _mk()
    other = _mk(ch=3, n=25, rate=16000)
    s = b"qoaf\0\0\0\0" + good[8:] + other[8:]
    info, pcm, hdrs = ob.decode_iter(s)  # iterator: dummy header first, channel/rate changes allowed
    assert info["status"] == ob.OK and info["streaming"] and hdrs[0] == (0, 0, 0)
    assert hdrs[1:] == [(2, 8000, 45), (3, 16000, 25)] and info["samples"] == 90 + 75
    assert ob.decode_all(s)[0]["status"] == ob.DEC_INCOMPATIBLE_FRAME  # decode_all quirk
    info, pcm = ob.decode_all(b"qoaf\0\0\0\0")
    assert info["status"] == ob.OK and info["samples"] == 0 and info["channels"] == 0


def test_mt_helpers_match_single_thread():
    n, ch, samples = 5, 2, 6000
    pcm = np.concatenate([signals.synth_pcm(1, s, ch, samples) for s in range(n)])
    rc, enc = ob.mt_encode(pcm, n, ch, 44100, samples, threads=3)
    assert rc == 0
    sz = ob.encoded_size(ch, samples)
    for s in range(n):
        assert enc[s * sz:(s + 1) * sz].tobytes() == ob.encode_all(pcm[s * samples * ch:(s + 1) * samples * ch], ch,
                                                                    44100, samples)[1]
    rc, dec = ob.mt_decode(enc, n, sz, samples * ch, threads=3)
    assert rc == 0 and np.array_equal(dec[:samples * ch], ob.decode_all(enc[:sz])[1])


def test_oracle_under_asan_ubsan():
    """the C restatement runs clean under -fsanitize=address,undefined WITHOUT -fwrapv (every wrapping op
    is done in uint32_t), on encode + streaming encode + both searches + decode of 1..8 channels, every truncation of
    small files, mutated headers and random payloads (oracle/san_check.c)."""
    import os
    import subprocess
    from conftest import ROOT
    odir = os.path.join(ROOT, "oracle")
    b = subprocess.run(["make", "-s", "-C", odir, "san_check"], capture_output=True, text=True)
    if b.returncode != 0:
        pytest.skip("no sanitizer runtime for this compiler: " + b.stderr.strip().splitlines()[-1][:120])
    r = subprocess.run([os.path.join(odir, "san_check")], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1"))
    assert r.returncode == 0 and "san_check ok" in r.stdout and "runtime error" not in r.stderr, r.stdout + r.stderr[-2000:]


def test_golden_dump_checker(tmp_path):
    """rust/golden-dump runs the reference crate on exported inputs; `make_golden.py --check-against` compares its output
    with the oracle.  No Rust toolchain here, so the checker itself is tested: a directory filled by the oracle passes,
    one flipped byte in one file is reported."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden as mg
    inp, out = tmp_path / "in", tmp_path / "out"
    mg.export_inputs(str(inp))
    assert (inp / "saw1000_mono.pcm").stat().st_size == 2000 and (inp / "saw1000_mono.desc").read_text().split() == ["1", "44100", "1000"]
    out.mkdir()
    for name, (sig, ch, rate, n) in mg.parity_inputs().items():
        st, enc = ob.encode_all(sig, ch, rate, n)
        (out / f"{name}.qoa").write_bytes(bytes(enc))
        ob.decode_all(bytes(enc))[1].astype("<i2").tofile(str(out / f"{name}.dec.pcm"))
    fixture = (inp / "fixture.qoa").read_bytes()
    pcm = ob.decode_all(fixture)[1]
    pcm.astype("<i2").tofile(str(out / "fixture.decoded.pcm"))
    (out / "fixture.reencoded.qoa").write_bytes(bytes(ob.encode_all(pcm, 2, 44100, 2394122)[1]))
    assert mg.check_against(str(out)) == 0
    blob = bytearray((out / "mix45_3ch.qoa").read_bytes())
    blob[40] ^= 1
    (out / "mix45_3ch.qoa").write_bytes(bytes(blob))
    assert mg.check_against(str(out)) == 1
    # the Rust side exists as source and names the same files
    main_rs = open(os.path.join(ROOT, "rust", "golden-dump", "src", "main.rs")).read()
    for token in (".dec.pcm", ".decoded.pcm", ".reencoded.qoa", "encode_all", "decode_all", "encode_frame"):
        assert token in main_rs


def test_probe_matches_oracle_walk(fixture_bytes):
    """qoa_b200_probe walks frame headers on the HOST (no GPU, no sample decoded) with the code the index pre-pass runs
    on the device; its info must equal what the oracle's decoder reports, for good files, every truncation of small
    files and malformed headers, in all three modes."""
    import qoaudio_b200 as q
    keys = ("status", "streaming", "header_samples", "channels", "sample_rate", "samples", "frames")

    def same(blob, mode):
        got = q.probe(blob, mode)
        oi = (ob.decode_iter(blob)[0] if mode == 0 else ob.decode_all(blob)[0])
        assert {k: got[k] for k in keys} == {k: oi[k] for k in keys}, (len(blob), mode, got, oi)

    same(fixture_bytes, 0)
    same(fixture_bytes, 1)
    same(fixture_bytes[:100000], 0)
    small = [ob.encode_all(signals.synth_pcm(3, i, ch, n), ch, 8000, n)[1] for i, (ch, n) in
             enumerate([(1, 45), (2, 5121), (3, 20), (8, 21)])]
    for blob in small:
        for cut in list(range(0, min(len(blob), 200))) + [len(blob) - 1, len(blob)]:
            same(blob[:cut], 0)
            same(blob[:cut], 1)
    rng = np.random.default_rng(9)
    base = bytearray(small[1])
    for _ in range(300):                       # header mutations: the shapes of fuzz/fuzz_targets/iter_all.rs
        m = bytearray(base)
        for _ in range(int(rng.integers(1, 4))):
            m[int(rng.integers(0, 40))] ^= 1 << int(rng.integers(0, 8))
        same(bytes(m), int(rng.integers(0, 2)))
    streaming = small[0][:4] + bytes(4) + small[0][8:]  # header samples == 0: Streaming mode and its dummy first header
    same(streaming, 0)
    same(streaming, 1)
    assert q.probe(small[1][8:], 2)["samples"] == 2 * 5121  # DECODE_FRAMES: frames without the file header
