"""CPU tests of oracle/vector_verify.py — the vectorised verifier (closed-form weight trajectories, slice-parallel
16-candidate search) — and, through it, two pins of the C oracle that do not depend on the oracle's own lineage:

* decode VALUES are pinned by the reference's own fixture (fixtures/julien_baker_sprained_ankle.qoa, written by
  somebody else's encoder): the LMS state at the end of every decoded frame must equal the state that encoder stored
  in the next frame header (src/lib.rs:270-281 reads it) — 934 x 128 bits, each a function of all 5120 samples of its frame;
* the oracle's sorted + pruned search (src/lib.rs:516-593) picks, slice for slice, what the slice-parallel
  lexicographic arg-min over all 16 candidates picks.
"""
import numpy as np
import pytest
import torch

import signals
from oracle import binding as ob
from oracle import vector_verify as vv

CLEAN_DECODE = ("bad_file_headers", "bad_frame_headers", "bad_samples", "bad_streams", "chain_history_bad",
                "chain_weights_bad", "chain_bad_streams")
CLEAN_ENCODE = CLEAN_DECODE + ("bad_slices", "bad_initial_state")


def u8(rows):
    return torch.from_numpy(np.stack([np.frombuffer(bytes(r), dtype=np.uint8) for r in rows]).copy())


def i16(rows):
    return torch.from_numpy(np.stack([np.asarray(r, dtype=np.int16).reshape(-1) for r in rows]).copy())


def test_fixture_decode_values_pinned_by_embedded_encoder_state(fixture_bytes):
    info, pcm = ob.decode_all(fixture_bytes)
    assert info["status"] == 0
    ch, n = info["channels"], info["header_samples"]
    r = vv.verify_decode(u8([fixture_bytes]), i16([pcm]), ch, n, info["sample_rate"])
    assert all(r[k] == 0 for k in CLEAN_DECODE), r
    assert r["samples"] == 4788244 and r["boundaries"] == 467 * 2 and r["truncated_boundaries"] == 0
    # and the verifier does see a wrong sample (the four predictions it feeds)
    bad = pcm.copy()
    bad[1234567] += 1
    r = vv.verify_decode(u8([fixture_bytes]), i16([bad]), ch, n)
    assert 1 <= r["bad_samples"] <= 5 and r["bad_streams"] == 1


def test_embedded_state_detects_any_tampered_slice(fixture_bytes):
    """How sharp the fixture pin is: take the first 12 frames of the fixture as a file of its own, flip ONE code bit in a
    random slice of a random frame (16 variants), decode the tampered file with the oracle (so every sample IS the
    decode of the file) — the state at the end of the tampered frame no longer matches the next header, every time."""
    ch, frames = 2, 12
    fbytes = 8 + 16 * ch + 2048 * ch
    n = frames * 5120
    base = bytearray(b"qoaf" + n.to_bytes(4, "big") + fixture_bytes[8:8 + frames * fbytes])
    rng = np.random.default_rng(11)
    files, pcms = [bytes(base)], []
    for _ in range(16):
        f, k, c = int(rng.integers(0, frames - 1)), int(rng.integers(0, 256)), int(rng.integers(0, ch))
        pos = 8 + f * fbytes + 8 + 16 * ch + (k * ch + c) * 8
        t = bytearray(base)
        bit = int(rng.integers(0, 60))                      # one of the 60 code bits (not the scalefactor nibble)
        t[pos + 7 - bit // 8] ^= 1 << (bit % 8)
        files.append(bytes(t))
    for f in files:
        info, pcm = ob.decode_all(f)
        assert info["status"] == 0 and pcm.size == n * ch
        pcms.append(pcm)
    clean = vv.verify_decode(u8(files[:1]), i16(pcms[:1]), ch, n)
    assert all(clean[k] == 0 for k in CLEAN_DECODE) and clean["boundaries"] == (frames - 1) * ch
    r = vv.verify_decode(u8(files[1:]), i16(pcms[1:]), ch, n)
    assert r["bad_samples"] == 0            # each is the true decode of its (tampered) file ...
    assert r["chain_bad_streams"] == 16     # ... and each contradicts the state the original encoder stored


SHAPES = [(2, 5120 * 3 + 777, "music"), (1, 20, "music"), (1, 1, "noise"), (1, 19, "music"), (3, 5119, "music"),
          (8, 10259, "music"), (2, 5120 * 2, "noise"), (5, 5121, "noise"), (4, 21, "music"), (6, 5120, "music"),
          (7, 123, "noise")]


@pytest.mark.parametrize("ch,n,kind", SHAPES)
def test_vectorised_search_agrees_with_oracle(ch, n, kind):
    """Oracle encode (sorted + pruned search) -> oracle decode -> every sample and every slice pass the verifier."""
    S = 3
    src = [signals.synth_pcm(7, s, ch, n, kind) for s in range(S)]
    enc = [ob.encode_all(p, ch, 44100, n)[1] for p in src]
    dec = [ob.decode_all(e)[1] for e in enc]
    r = vv.verify_encode(u8(enc), i16(src), i16(dec), ch, n, 44100)
    assert all(r[k] == 0 for k in CLEAN_ENCODE), r
    assert r["slices"] == S * ch * (-(-n // 20)) and r["samples"] == S * ch * n
    if kind == "music":
        assert r["unverifiable_slices"] == 0


def test_kats_and_edge_signals_pass_the_verifier():
    for name, (pcm, ch, rate, n) in signals.kat_signals().items():
        st, enc = ob.encode_all(pcm, ch, rate, n)
        dec = ob.decode_all(enc)[1]
        r = vv.verify_encode(u8([enc]), i16([pcm]), i16([dec]), ch, n, rate)
        assert st == 0 and all(r[k] == 0 for k in CLEAN_ENCODE), (name, r)
    for name, pcm in signals.edge_signals().items():
        for ch in (1, 2, 3):
            n = pcm.size // ch
            if n == 0:
                continue
            p = pcm[:n * ch]
            st, enc = ob.encode_all(p, ch, 8000, n)
            dec = ob.decode_all(enc)[1]
            r = vv.verify_encode(u8([enc]), i16([p]), i16([dec]), ch, n, 8000)
            assert st == 0 and all(r[k] == 0 for k in CLEAN_ENCODE), (name, ch, r)


def test_encode_verifier_flags_a_suboptimal_slice():
    """Replace one slice by the same samples encoded with another scalefactor (what a wrong search would emit): the
    file still decodes, the verifier names the slice."""
    ch, n = 2, 5120 * 2 + 40
    src = signals.synth_pcm(3, 0, ch, n)
    enc = bytearray(ob.encode_all(src, ch, 48000, n)[1])
    f, k, c = 1, 17, 1
    pos = 8 + f * (8 + 16 * ch + 2048 * ch) + 8 + 16 * ch + (k * ch + c) * 8
    sf = enc[pos] >> 4
    enc[pos] = (enc[pos] & 0x0F) | (((sf + 3) % 16) << 4)
    info, dec = ob.decode_all(bytes(enc))
    assert info["status"] == 0
    r = vv.verify_encode(u8([enc]), i16([src]), i16([dec]), ch, n, 48000)
    assert r["bad_samples"] == 0 and r["bad_slices"] >= 1 and r["chain_bad_streams"] == 1


def test_fixture_reencode_every_slice(fixture_bytes):
    """The oracle's re-encode of the decoded fixture (239,414 slices): every slice is the arg-min candidate."""
    info, pcm = ob.decode_all(fixture_bytes)
    ch, n, rate = info["channels"], info["header_samples"], info["sample_rate"]
    st, enc = ob.encode_all(pcm, ch, rate, n)
    dec = ob.decode_all(enc)[1]
    r = vv.verify_encode(u8([enc]), i16([pcm]), i16([dec]), ch, n, rate, max_lanes=1 << 22)
    assert st == 0 and all(r[k] == 0 for k in CLEAN_ENCODE), r
    assert r["slices"] == ch * (-(-n // 20)) and r["unverifiable_slices"] == 0


def random_payload_stream(rng, ch, rate, n):
    """Valid headers + random LMS + random slice words (what qoa_b200_gen_qoa_dev makes on the device)."""
    size = vv.encoded_size(ch, n)
    buf = bytearray(rng.bytes(size))
    buf[0:8] = b"qoaf" + n.to_bytes(4, "big")
    pos, left = 8, n
    while left:
        flen = min(5120, left)
        k = -(-flen // 20)
        buf[pos:pos + 8] = ((ch << 56) | (rate << 32) | (flen << 16) | (8 + 16 * ch + 8 * k * ch)).to_bytes(8, "big")
        pos += 8 + 16 * ch + 8 * k * ch
        left -= flen
    return bytes(buf)


def test_chunked_drivers_on_cpu():
    """tests/verify_chunks.py is what the -m gpu tests run on CUDA tensors at full size; same code, small batch, CPU."""
    import verify_chunks as vc
    rng = np.random.default_rng(2)
    ch, n, S = 2, 5120 * 3 + 4080, 11
    files = [random_payload_stream(rng, ch, 44100, n) for _ in range(S)]
    stride = (len(files[0]) + 15) // 16 * 16
    pcms = [ob.decode_all(f)[1] for f in files]
    qoa = torch.zeros(S, stride, dtype=torch.uint8)
    qoa[:, :len(files[0])] = u8(files)
    pcm = torch.full((S, n * ch + 24), 1234, dtype=torch.int16)   # guard values behind every stream, as in the GPU test
    pcm[:, :n * ch] = i16(pcms)
    r = vc.decode_in_chunks(qoa, pcm, ch, n, 44100, chunk=4)
    assert r["streams"] == S and r["samples"] == S * n * ch and r["bad_samples"] == 0 and r["bad_frame_headers"] == 0
    assert r["chain_bad_streams"] == S                       # random LMS in every header: frames do not chain
    r = vc.decode_in_chunks(qoa, pcm, ch, n, 44100, streams=[0, 5, 10], chunk=2)
    assert r["streams"] == 3 and r["bad_samples"] == 0
    pcm[7, 4321] ^= 1
    r = vc.decode_in_chunks(qoa, pcm, ch, n, 44100, chunk=4)
    assert r["bad_streams"] == 1 and r["bad_samples"] >= 1
    # encode driver: a partial last frame with a partial last slice
    n = 5120 * 2 + 333
    src = [signals.synth_pcm(5, s, ch, n, "noise" if s == 3 else "music") for s in range(S)]
    enc = [ob.encode_all(p, ch, 48000, n)[1] for p in src]
    dec = [ob.decode_all(e)[1] for e in enc]
    stride = (len(enc[0]) + 15) // 16 * 16
    qoa = torch.zeros(S, stride, dtype=torch.uint8)
    qoa[:, :len(enc[0])] = u8(enc)
    r = vc.encode_in_chunks(qoa, i16(src), i16(dec), ch, n, 48000, chunk=4, max_lanes=1 << 16)
    assert all(r[k] == 0 for k in CLEAN_ENCODE) and r["streams"] == S and r["slices"] == S * ch * (-(-n // 20)), r


@pytest.mark.parametrize("ch", [1, 2, 3, 4, 5, 6, 7, 8])
def test_random_payload_decode_all_channel_counts(ch):
    """Random slice words and random 16-bit LMS states (weights far outside what an encoder produces: the i32 wrap of
    predict, src/lib.rs:799-805, is exercised) for 1..8 channels, with a partial last frame and slice: the oracle's
    decode is the fixed point of the recurrences; one flipped sample is found."""
    rng = np.random.default_rng(100 + ch)
    n, S = 5120 * 2 + 20 * 37 + 11, 4
    files = [random_payload_stream(rng, ch, 32000, n) for _ in range(S)]
    pcms = [ob.decode_all(f)[1] for f in files]
    assert all(p.size == n * ch for p in pcms)
    r = vv.verify_decode(u8(files), i16(pcms), ch, n, 32000)
    assert r["bad_samples"] == 0 and r["bad_frame_headers"] == 0 and r["bad_file_headers"] == 0 and r["samples"] == S * n * ch
    pcms[2][n * ch - 1] ^= 0x40                              # the very last sample (inside the partial slice)
    r = vv.verify_decode(u8(files), i16(pcms), ch, n, 32000)
    assert r["bad_samples"] == 1 and r["bad_streams"] == 1
