"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the C ABI, against the CPU oracle
and the committed golden vectors.  Bit-exact: identical .qoa bytes on encode, identical i16 samples on decode."""
import hashlib
import os

import numpy as np
import pytest

import signals
from oracle import binding as ob

pytestmark = pytest.mark.gpu


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


@pytest.fixture(scope="module")
def eng():
    import qoaudio_b200
    return qoaudio_b200.default_engine()


@pytest.fixture(scope="module")
def N():
    import qoaudio_b200._native as n
    return n


INFO_KEYS = ("status", "streaming", "header_samples", "channels", "sample_rate", "samples", "frames")


def oracle_info(data, mode):
    if mode == 0:
        info, pcm, _ = ob.decode_iter(data)
    else:
        info, pcm = ob.decode_all(data)
    return info, pcm


def check_decode_batch(eng, streams, mode):
    infos, outs = eng.decode_batch(streams, mode)
    for i, s in enumerate(streams):
        oi, opcm = oracle_info(s, mode)
        got = {k: infos[i][k] for k in INFO_KEYS}
        assert got == oi, (i, len(s), got, oi)
        assert np.array_equal(outs[i], opcm), (i, "samples differ")


# ---- config 1: the reference fixture -----------------------------------------------------------------------------

def test_fixture_decode_matches_golden(eng, fixture_bytes, golden):
    infos, outs = eng.decode_batch([fixture_bytes], 0)
    info = infos[0]
    assert info["status"] == 0 and not info["streaming"]
    assert (info["channels"], info["sample_rate"], info["header_samples"]) == (2, 44100, 2394122)  # src/lib.rs:1005-1012
    assert info["frames"] == 468 and info["samples"] == 2394122 * 2                                   # :1018-1031
    assert sha(outs[0].tobytes()) == golden["appendix_b"]["fixture_decoded_pcm_sha256"]
    _, opcm, _ = ob.decode_iter(fixture_bytes)
    assert np.array_equal(outs[0], opcm)


def test_fixture_reencode_matches_golden(eng, fixture_bytes, golden):
    _, pcm, _ = ob.decode_iter(fixture_bytes)
    st, outs = eng.encode_all_batch([pcm], [(2, 44100, 2394122)])
    assert st == [0]
    assert len(outs[0]) == 1934040
    assert sha(outs[0].tobytes()) == golden["appendix_b"]["fixture_reencode_sha256"]
    infos, dec = eng.decode_batch([outs[0].tobytes()], 1)
    assert sha(dec[0].tobytes()) == golden["appendix_b"]["fixture_reencode_decoded_sha256"]


# ---- known-answer encodes ---------------------------------------------------------------------------------------------

def test_kat_encodes(eng, golden):
    kats = signals.kat_signals()
    names = sorted(kats)
    st, outs = eng.encode_all_batch([kats[k][0] for k in names], [kats[k][1:] for k in names])
    assert st == [0] * len(names)
    for k, o in zip(names, outs):
        g = golden["appendix_b"]["kats"][k]
        assert len(o) == g["len"], k
        assert sha(o.tobytes()) == g["sha256"], k
        ost, obytes = ob.encode_all(kats[k][0], *kats[k][1:])
        assert o.tobytes() == obytes


def test_edge_signals_all_channel_counts(eng):
    pcm_list, descs = [], []
    for name, sig in signals.edge_signals().items():
        for ch in range(1, 9):
            n = sig.size // ch
            if n == 0:
                continue
            pcm_list.append(sig[:n * ch])
            descs.append((ch, 8000 + 1000 * ch, n))
    st, outs = eng.encode_all_batch(pcm_list, descs)
    assert st == [0] * len(descs)
    for p, d, o in zip(pcm_list, descs, outs):
        ost, obytes = ob.encode_all(p, *d)
        assert ost == 0 and o.tobytes() == obytes, d
    check_decode_batch(eng, [o.tobytes() for o in outs], 0)
    check_decode_batch(eng, [o.tobytes() for o in outs], 1)


# ---- config 5: mixed batch --------------------------------------------------------------------------------------------

def mixed_batch(count=256, seed=1, long_every=16):
    rates = [8000, 11025, 16000, 22050, 32000, 44100, 48000, 96000]
    lengths = [1, 19, 20, 21, 39, 40, 41, 5119, 5120, 5121, 10259, 12345]  # + 123457 / ch and rate * 7 + 13 below
    pcm_list, descs = [], []
    for i in range(count):
        ch = i % 8 + 1
        rate = rates[(i // 8) % len(rates)]
        n = lengths[(i * 7 + i // 8) % len(lengths)]
        if long_every and i % long_every == long_every - 1:
            n = 123457 // ch
        elif long_every and i % long_every == 7:
            n = (rate * 7 + 13) // (8 if rate > 16000 else 1)
        kind = "noise" if i % 5 == 4 else "music"
        pcm_list.append(signals.synth_pcm(seed, i, ch, n, kind))
        descs.append((ch, rate, n))
    return pcm_list, descs


def test_mixed_batch_roundtrip(eng):
    """BASELINE.json configs[4]: 256 mixed streams, 1-8 channels, 8-96 kHz, odd lengths with partial frames/slices."""
    pcm_list, descs = mixed_batch()
    assert len(descs) == 256 and {d[0] for d in descs} == set(range(1, 9))
    st, outs = eng.encode_all_batch(pcm_list, descs)
    assert st == [0] * len(descs)
    enc = []
    for p, d, o in zip(pcm_list, descs, outs):
        ost, obytes = ob.encode_all(p, *d)
        assert ost == 0
        assert len(o) == len(obytes), d
        assert o.tobytes() == obytes, d
        enc.append(obytes)
    infos, dec = eng.decode_batch(enc, 1)
    for i, (d, e) in enumerate(zip(descs, enc)):
        oi, opcm = ob.decode_all(e)
        assert infos[i]["status"] == 0 and infos[i]["samples"] == d[0] * d[2]
        assert (infos[i]["channels"], infos[i]["sample_rate"]) == (d[0], d[1])
        assert np.array_equal(dec[i], opcm), d


def test_stereo_batch_fast_path(eng):
    """All-stereo batches take the fast stereo kernel: full frames, partial last frames and partial last slices."""
    lens = [5120, 5120 * 3, 5120 * 2 + 20, 5120 + 1, 5120 + 19, 20, 7, 4080, 2 * 5120 + 3082]
    pcm_list = [signals.synth_pcm(3, i, 2, n) for i, n in enumerate(lens)]
    descs = [(2, 44100, n) for n in lens]
    st, outs = eng.encode_all_batch(pcm_list, descs)
    assert st == [0] * len(lens)
    enc = [o.tobytes() for o in outs]
    for p, d, e in zip(pcm_list, descs, enc):
        assert e == ob.encode_all(p, *d)[1]
    check_decode_batch(eng, enc, 0)


def test_mono_batch_fast_path(eng):
    """All-mono batches take the same TMA-staged kernel with one chain per thread (6 slices per staged row)."""
    lens = [5120, 5120 * 3, 5120 * 2 + 20, 5120 + 1, 5120 + 19, 20, 7, 1, 4080, 2 * 5120 + 3081, 5120 * 40 + 119]
    pcm_list = [signals.synth_pcm(4, i, 1, n, "noise" if i % 3 == 0 else "music") for i, n in enumerate(lens)]
    descs = [(1, 22050, n) for n in lens]
    st, outs = eng.encode_all_batch(pcm_list, descs)
    assert st == [0] * len(lens)
    enc = [o.tobytes() for o in outs]
    for p, d, e in zip(pcm_list, descs, enc):
        assert e == ob.encode_all(p, *d)[1]
    check_decode_batch(eng, enc, 0)
    check_decode_batch(eng, enc * 7, 1)      # enough frames for regular 32-frame warps and stream-straddling warps
    check_decode_batch(eng, [e[:len(e) - 9] for e in enc[:4]] + enc, 0)  # truncated mono streams in the batch


@pytest.mark.parametrize("ch", [3, 4, 5, 6, 7, 8])
def test_multichannel_batch_fast_path(eng, ch):
    """Uniform 3..8-channel batches take the pair-per-lane kernel: warps whose frames are full-size neighbours in one
    stream leave through one tensor store per group (the 40-frame stream: 16 frames per warp at most), every other
    warp through the contiguous row copy."""
    lens = [5120, 5120 * 2 + 20, 5120 + 1, 5120 + 19, 20, 7, 1, 4080, 5120 * 3 + 3081, 5120 * 9 + 119, 5120 * 40 + 13]
    pcm_list = [signals.synth_pcm(8, i, ch, n, "noise" if i % 4 == 1 else "music") for i, n in enumerate(lens)]
    descs = [(ch, 48000, n) for n in lens]
    st, outs = eng.encode_all_batch(pcm_list, descs)
    assert st == [0] * len(lens)
    enc = [o.tobytes() for o in outs]
    for pcm, d, e in zip(pcm_list, descs, enc):
        assert e == ob.encode_all(pcm, *d)[1]
    check_decode_batch(eng, enc, 0)
    check_decode_batch(eng, enc * 3 + [e[:len(e) - 17] for e in enc[:3]], 1)  # more frames + truncated streams


def test_noise_brute_equivalence(eng):
    """White noise is the worst case for the reference's pruning; the GPU's full search must pick the same winners."""
    pcm = signals.synth_pcm(2, 5, 2, 30000, "noise")
    st, outs = eng.encode_all_batch([pcm], [(2, 48000, 30000)])
    assert st == [0]
    assert outs[0].tobytes() == ob.encode_all(pcm, 2, 48000, 30000)[1]
    assert outs[0].tobytes() == ob.encode_all(pcm, 2, 48000, 30000, brute=True)[1]


# ---- encoder errors and streaming ------------------------------------------------------------------------------------

def test_encoder_errors(eng, N):
    """src/lib.rs:1148-1183 + :368 + precedence :336-344."""
    z = np.zeros(10, dtype=np.int16)
    descs = [(0, 44100, 10), (9, 44100, 10), (1, 0, 10), (1, 44100, 0), (0, 0, 0), (2, 0, 0), (1, 44100, 11), (1, 44100, 10)]
    st, outs = eng.encode_all_batch([z] * len(descs), descs)
    assert st == [N.ENC_INVALID_CHANNELS, N.ENC_INVALID_CHANNELS, N.ENC_INVALID_SAMPLE_RATE, N.ENC_INVALID_SAMPLES,
                  N.ENC_INVALID_CHANNELS, N.ENC_INVALID_SAMPLE_RATE, N.ENC_INVALID_SAMPLES, 0]
    for d, s in zip(descs, st):
        ost, _ = ob.encode_all(z, *d)
        assert ost == s
    assert outs[-1].tobytes() == ob.encode_all(z, 1, 44100, 10)[1]


def test_streaming_encode_equals_one_shot(eng):
    """src/lib.rs:1261-1297, for a batch of differently shaped streams; also checks the carried i32 LMS state."""
    shapes = [(1, 44100, 5120 * 2 + 777), (2, 48000, 5120 * 3), (3, 8000, 5120 + 1), (8, 96000, 5120 * 2 + 19)]
    pcm_list = [signals.synth_pcm(9, i, c, n, "noise" if i == 2 else "music") for i, (c, r, n) in enumerate(shapes)]
    st, one_shot = eng.encode_all_batch(pcm_list, shapes)
    assert st == [0] * len(shapes)
    batch = eng.encoder_batch(shapes)
    assert batch.status == [0] * len(shapes)
    streams = [bytearray(h) for h in batch.write_header()]
    oracles = [ob.StreamEncoder(*s) for s in shapes]
    rounds = max((n + 5119) // 5120 for (_, _, n) in shapes)
    for r in range(rounds):
        frames = []
        for (c, _, n), p in zip(shapes, pcm_list):
            lo, hi = r * 5120, min(n, (r + 1) * 5120)
            frames.append(p[lo * c:hi * c] if lo < n else None)
        fst, outs = batch.encode_frame(frames)
        for i, f in enumerate(frames):
            if f is None:
                assert outs[i] == b""
                continue
            assert fst[i] == 0
            ost, obytes = oracles[i].encode_frame(f)
            assert outs[i] == obytes
            streams[i] += outs[i]
            for c in range(shapes[i][0]):
                assert batch.lms(i, c) == tuple(oracles[i].lms(c))
    for s, o in zip(streams, one_shot):
        assert bytes(s) == o.tobytes()
    batch.close()


def test_encode_frame_errors(eng, N):
    """src/lib.rs:424-430"""
    batch = eng.encoder_batch([(2, 44100, 100), (2, 44100, 100), (2, 44100, 100), (0, 1, 1)])
    assert batch.status == [0, 0, 0, N.ENC_INVALID_CHANNELS]
    st, outs = batch.encode_frame([np.zeros(0, np.int16), np.zeros(3, np.int16), np.zeros(5121 * 2, np.int16),
                                   np.zeros(2, np.int16)])
    assert st == [N.ENC_INVALID_SAMPLES] * 3 + [N.ENC_INVALID_CHANNELS]
    assert outs == [b""] * 4
    batch.close()


# ---- decoder: malformed / foreign inputs -----------------------------------------------------------------------------

def small_file(ch=2, n=45, rate=8000, seed=4):
    pcm = signals.synth_pcm(seed, 0, ch, n)
    st, b = ob.encode_all(pcm, ch, rate, n)
    assert st == 0
    return b


def test_truncated_inputs_every_length(eng):
    data = small_file(ch=2, n=45) + small_file(ch=2, n=45)[8:]  # two frames
    cuts = [data[:k] for k in range(len(data) + 1)]
    check_decode_batch(eng, cuts, 0)
    check_decode_batch(eng, cuts, 1)
    big = small_file(ch=3, n=5120 + 30, rate=44100)
    cuts = [big[:k] for k in list(range(0, 80)) + list(range(len(big) - 200, len(big) + 1)) + [24632, 24640, 24641]]
    check_decode_batch(eng, cuts, 0)


def be64(v):
    return int(v).to_bytes(8, "big")


def frame(ch, rate, nspc, seed=0, frame_size=None, slices=None):
    """A hand-built frame with pseudo-random LMS and slices."""
    nsl = (nspc + 19) // 20 if slices is None else slices
    fs = 8 + 16 * ch + 8 * nsl * ch if frame_size is None else frame_size
    rng = np.random.default_rng(seed)
    body = rng.integers(0, 256, size=16 * ch + 8 * nsl * ch, dtype=np.uint8).tobytes()
    return be64((ch << 56) | (rate << 32) | (nspc << 16) | (fs & 0xffff)) + body


def header(samples):
    return b"qoaf" + int(samples).to_bytes(4, "big")


def test_malformed_headers(eng):
    f = frame(2, 44100, 40, 1)
    cases = [
        b"", b"qoa", b"qoaf", b"qoag\0\0\0\1", b"qoaf\0\0\0", header(40),                  # IoError/NotQoaFile/NoSamples
        header(40) + f,                                                                       # ok
        header(40) + frame(0, 44100, 40), header(40) + frame(2, 0, 40),                      # :227
        header(40) + frame(2, 44100, 40, frame_size=8 + 32),                                  # :233
        header(40) + frame(2, 44100, 40, frame_size=8 + 32 + 8 * 3),                          # :239
        header(40) + frame(1, 44100, 5120, frame_size=8 + 16 + 8 * 257, slices=257),          # :242
        header(40) + f + frame(1, 44100, 40, 2),                                              # IncompatibleFrame (channels)
        header(40) + f + frame(2, 48000, 40, 2),                                              # IncompatibleFrame (rate)
        header(40) + f + b"\1\2\3",                                                           # trailing junk < 8: clean EOF
        header(40) + f + frame(2, 44100, 0, 3, frame_size=8 + 32 + 16) + frame(2, 44100, 25, 4),  # nspc == 0 frame
        header(0) + f + frame(3, 48000, 25, 5) + frame(1, 8000, 1, 6),                        # streaming file: changes allowed
        header(0),                                                                            # streaming, no frames
        header(99) + frame(2, 44100, 40, 7, frame_size=8 + 32 + 16 * 5) + f,                  # frame_size ignored for walking
        header(40) + frame(2, 44100, 65535, 8, frame_size=8 + 32 + 16, slices=3277)[:4000],   # huge nspc, truncated
        header(40) + frame(255, 16777215, 21, 9, frame_size=(8 + 16 * 255 + 16 * 255) & 0xffff),  # 255 channels
        header(7) + frame(2, 44100, 3, 10) + frame(2, 44100, 4, 11),                          # odd-length interior frame
    ]
    check_decode_batch(eng, cases, 0)
    check_decode_batch(eng, cases, 1)


def test_speculative_index_falls_back_to_walking(eng):
    """Conforming-length files take the closed-form index; a single foreign frame header anywhere in the batch must
    send the batch back to the serial walk and give exactly the walk's (= the reference's) result."""
    good = [ob.encode_all(signals.synth_pcm(6, i, c, n), c, 44100, n)[1] for i, (c, n) in
            enumerate([(2, 5120 * 3 + 77), (1, 5120 * 2), (2, 25), (3, 5121)])]
    fsz = 8 + 32 + 2048 * 2
    cases = list(good)
    for pos, patch in ((8 + fsz + 1, 0x12), (8 + 2 * fsz + 5, 0x99), (8 + fsz, 0x03), (8 + 3 * fsz + 7, 0x00), (8 + fsz + 4, 0xff)):
        b = bytearray(good[0])
        b[pos] = patch  # rate / nspc / channels / frame_size of an interior or the last frame header
        cases.append(bytes(b))
    b = bytearray(good[0])
    b[8 + fsz + 8 + 3] ^= 0x55  # LMS bytes are payload, not layout: still regular, different samples
    cases.append(bytes(b))
    for mode in (0, 1):
        check_decode_batch(eng, cases, mode)
        check_decode_batch(eng, good, mode)
    infos, _ = eng.decode_batch(good, 0)
    assert all(i["flags"] & 4 for i in infos)          # QOA_B200_INFO_REGULAR
    infos, _ = eng.decode_batch(cases, 0)
    assert not any(i["flags"] & 4 for i in infos)      # the whole batch was walked


def test_streaming_decode_frames(eng, fixture_bytes, N):
    """src/lib.rs:1034-1066: new_streaming + decode_frame on the first two fixture frames."""
    pos = 8
    frames = []
    for _ in range(3):
        size = int.from_bytes(fixture_bytes[pos + 6:pos + 8], "big")
        frames.append(fixture_bytes[pos:pos + size])
        pos += size
    infos, outs = eng.decode_batch(frames + [frames[0] + frames[1]], N.DECODE_FRAMES)
    sd = ob.StreamDecoder()
    for i in range(3):
        st, opcm = sd.decode_frame(frames[i])
        assert st == 0 and infos[i]["status"] == 0
        assert outs[i].size == 5120 * 2
        assert np.array_equal(outs[i], opcm)
    assert np.array_equal(outs[3], np.concatenate([outs[0], outs[1]]))


# ---- generators --------------------------------------------------------------------------------------------------------

def test_gen_pcm_matches_numpy(eng):
    import torch
    for kind in ("music", "noise"):
        for ch, n, start in ((2, 1000, 0), (3, 777, 12345), (1, 64, 2 ** 33 + 5)):
            t = torch.empty(4 * ch * n, dtype=torch.int16, device="cuda")
            eng.gen_pcm_dev(t.data_ptr(), 4, 10, ch, n, start, 7, kind)
            torch.cuda.synchronize()
            got = t.cpu().numpy().reshape(4, -1)
            for s in range(4):
                assert np.array_equal(got[s], signals.synth_pcm(7, 10 + s, ch, n, kind, start)), (kind, ch, n, s)


def test_gen_qoa_decodes_like_oracle(eng):
    import torch
    ch, rate, n = 2, 44100, 5120 * 2 + 4080
    size = ob.encoded_size(ch, n)
    stride = (size + 15) // 16 * 16
    t = torch.zeros(3 * stride, dtype=torch.uint8, device="cuda")
    eng.gen_qoa_dev(t.data_ptr(), stride, 3, 0, ch, rate, n, 1)
    torch.cuda.synchronize()
    host = t.cpu().numpy()
    streams = [host[i * stride:i * stride + size].tobytes() for i in range(3)]
    assert len(set(streams)) == 3
    check_decode_batch(eng, streams, 0)
    # the same through the device-resident API
    plan = eng.decode_plan_dev(t.data_ptr(), [i * stride for i in range(3)], [size] * 3, 1)
    out = torch.zeros(3 * n * ch, dtype=torch.int16, device="cuda")
    st = plan.run_dev(out.data_ptr(), [i * n * ch for i in range(3)], [n * ch] * 3)
    torch.cuda.synchronize()
    assert list(st) == [0, 0, 0]
    got = out.cpu().numpy().reshape(3, -1)
    for i in range(3):
        assert np.array_equal(got[i], ob.decode_all(streams[i])[1])
    plan.close()


def test_mixed_classes_device_offsets(eng):
    """A shuffled batch of mono / stereo / 5.1 / channel-switching streams: the plan groups the frame table by kernel
    class and launches each class on its own range.  Device-resident API with caller-chosen output offsets; one mono
    slot is deliberately not 16-byte aligned, which sends only the mono class through the generic kernel."""
    import torch
    rng = np.random.default_rng(11)
    streams, expect = [], []
    for i in range(48):
        ch = (1, 2, 6, 2, 1, 3)[i % 6]
        n = int(rng.integers(1, 5120 * 4))
        pcm = signals.synth_pcm(12, i, ch, n, "noise" if i % 7 == 3 else "music")
        e = ob.encode_all(pcm, ch, 32000, n)[1]
        if i % 12 == 5:  # stereo frames followed by mono frames in one stream (the iterator re-reads channels per frame)
            m = ob.encode_all(signals.synth_pcm(13, i, 1, 6000), 1, 32000, 6000)[1]
            e = e[:4] + bytes(4) + e[8:] + m[8:]  # streaming header: the channel switch is legal
        streams.append(e)
        expect.append(ob.decode_iter(e))
    sizes = [len(s) for s in streams]
    in_off = np.concatenate([[0], np.cumsum([(s + 15) // 16 * 16 for s in sizes])]).astype(np.uint64)
    arena = np.zeros(int(in_off[-1]), dtype=np.uint8)
    for o, s in zip(in_off, streams):
        arena[int(o):int(o) + len(s)] = np.frombuffer(s, dtype=np.uint8)
    d_arena = torch.from_numpy(arena).cuda()
    for misalign in (0, 3):
        caps = [len(x[1]) for x in expect]
        out_off, total = [], 0
        for i, c in enumerate(caps):
            total = (total + 7) // 8 * 8 + (misalign if i == 6 else 0)
            out_off.append(total)
            total += c
        out = torch.full((total + 8,), -21846, dtype=torch.int16, device="cuda")
        plan = eng.decode_plan_dev(d_arena.data_ptr(), in_off[:-1], sizes, 0)
        st = plan.run_dev(out.data_ptr(), out_off, caps)
        torch.cuda.synchronize()
        host = out.cpu().numpy()
        for i, (oi, opcm, _) in enumerate(expect):
            assert st[i] == oi["status"], (i, st[i], oi)
            assert {k: plan.info[i][k] for k in INFO_KEYS} == oi, i
            assert np.array_equal(host[out_off[i]:out_off[i] + caps[i]], opcm), (i, misalign)
        plan.close()
    # Base pointers that are not 16-byte aligned (a torch view with a storage offset): the output arena then takes the
    # generic kernel for every class (the fast kernels store 16-byte rows), a misaligned input arena is refused.
    caps = [len(x[1]) for x in expect]
    out_off = np.concatenate([[0], np.cumsum([(c + 7) // 8 * 8 for c in caps])[:-1]]).astype(np.uint64)
    big = torch.full((int(out_off[-1]) + caps[-1] + 16,), -21846, dtype=torch.int16, device="cuda")
    view = big[3:]                                   # data_ptr() is 6 bytes past a 256-byte boundary
    assert view.data_ptr() % 16 == 6
    plan = eng.decode_plan_dev(d_arena.data_ptr(), in_off[:-1], sizes, 0)
    st = plan.run_dev(view.data_ptr(), out_off, caps)
    torch.cuda.synchronize()
    host = view.cpu().numpy()
    for i, (oi, opcm, _) in enumerate(expect):
        assert st[i] == oi["status"] and np.array_equal(host[int(out_off[i]):int(out_off[i]) + caps[i]], opcm), i
    assert (big[:3].cpu().numpy() == -21846).all()
    plan.close()
    import qoaudio_b200._native as NN
    with pytest.raises(NN.NativeError):
        eng.decode_plan_dev(d_arena.data_ptr() + 8, in_off[:-1], sizes, 0)
    enc_out = torch.zeros(4096, dtype=torch.uint8, device="cuda")
    pcm_d = torch.zeros(64, dtype=torch.int16, device="cuda")
    with pytest.raises(NN.NativeError):              # 8-byte slice stores need an 8-byte aligned output arena
        eng.encode_all_batch_dev([(1, 8000, 20)], pcm_d.data_ptr(), [0], [20], enc_out.data_ptr() + 4, [0], [4096])


def test_device_encode_chunked_equals_one_shot(eng):
    """Time-chunked device-resident encode with persistent LMS state (the config-4 path) == encode_all."""
    import torch
    n_streams, ch, rate, n = 6, 2, 48000, 5120 * 5 + 2560
    pcm = torch.empty(n_streams * ch * n, dtype=torch.int16, device="cuda")
    eng.gen_pcm_dev(pcm.data_ptr(), n_streams, 100, ch, n, 0, 1, "music")
    size = ob.encoded_size(ch, n)
    stride = (size + 15) // 16 * 16
    out = torch.zeros(n_streams * stride, dtype=torch.uint8, device="cuda")
    st, lens = eng.encode_all_batch_dev([(ch, rate, n)] * n_streams, pcm.data_ptr(), [i * ch * n for i in range(n_streams)],
                                        [ch * n] * n_streams, out.data_ptr(), [i * stride for i in range(n_streams)],
                                        [stride] * n_streams)
    torch.cuda.synchronize()
    assert st == [0] * n_streams and lens == [size] * n_streams
    one = out.cpu().numpy().reshape(n_streams, stride)[:, :size]
    host_pcm = pcm.cpu().numpy().reshape(n_streams, -1)
    for i in range(n_streams):
        assert np.array_equal(host_pcm[i], signals.synth_pcm(1, 100 + i, ch, n))
        assert one[i].tobytes() == ob.encode_all(host_pcm[i], ch, rate, n)[1]
    # chunked: 2 frames per call
    out2 = torch.zeros(n_streams * stride, dtype=torch.uint8, device="cuda")
    batch = eng.encoder_batch([(ch, rate, n)] * n_streams)
    fbytes = 8 + 16 * ch + 2048 * ch
    done = 0
    while done < n:
        step = min(2 * 5120, n - done)
        batch.encode_chunk_dev(pcm.data_ptr(), [i * ch * n + done * ch for i in range(n_streams)], [step] * n_streams,
                               out2.data_ptr(), [i * stride + 8 + (done // 5120) * fbytes for i in range(n_streams)])
        done += step
    torch.cuda.synchronize()
    two = out2.cpu().numpy().reshape(n_streams, stride)[:, 8:size]
    assert np.array_equal(two, one[:, 8:])
    batch.close()


def test_decode_all_batch_pipelined(eng):
    """qoa_b200_decode_all_batch pipelines large host batches over chunks of streams on three CUDA streams; the
    result must equal the one-shot plan path and the oracle, including failing streams in the middle of a chunk."""
    import torch
    ch, rate, n = 2, 44100, 5120 * 200 + 4080
    size = ob.encoded_size(ch, n)
    stride = (size + 15) // 16 * 16
    count = 96  # ~80 MB of input: above the pipelining threshold
    t = torch.zeros(count * stride, dtype=torch.uint8, device="cuda")
    eng.gen_qoa_dev(t.data_ptr(), stride, count, 0, ch, rate, n, 5)
    torch.cuda.synchronize()
    host = t.cpu().numpy()
    streams = [host[i * stride:i * stride + size] for i in range(count)]
    streams[7] = streams[7][:size - 1000]          # truncated: IoError, nothing written beyond what was decoded
    streams[40] = np.frombuffer(b"nope" + bytes(60), dtype=np.uint8)
    caps = [n * ch] * count
    info, status, outs = eng.decode_all_batch(streams, caps)
    for i in (0, 1, 7, 39, 40, 41, 95):
        oi, opcm = ob.decode_all(streams[i].tobytes())
        assert status[i] == oi["status"] and int(info["samples"][i]) == oi["samples"], i
        assert np.array_equal(outs[i][:oi["samples"]], opcm), i
    infos2, outs2 = eng.decode_batch([s.tobytes() for s in streams[:12]], 1)
    for i in range(12):
        assert np.array_equal(outs[i][:infos2[i]["samples"]], outs2[i])


def test_pipelined_host_paths_bound_device_memory(N):
    """The pipelined host paths hold the arenas of about three pipeline stages, not of the whole batch: a block released
    by a stage on pipeline stream k is handed out again to later stages of the same stream.  A fresh context decodes
    1.9 GB of PCM (360 x 2.6 M samples) and then encodes it again; the device memory its pool keeps afterwards (= the
    most it ever held) stays far below the batch."""
    import torch
    import qoaudio_b200 as q
    e2 = q.Engine(0)
    ch, rate, n = 2, 44100, 5120 * 256
    size = ob.encoded_size(ch, n)
    stride = (size + 15) // 16 * 16
    count = 360
    t = torch.zeros(count * stride, dtype=torch.uint8, device="cuda")
    e2.gen_qoa_dev(t.data_ptr(), stride, count, 0, ch, rate, n, 9)
    torch.cuda.synchronize()
    host = t.cpu().numpy()
    del t
    torch.cuda.empty_cache()
    streams = [host[i * stride:i * stride + size] for i in range(count)]
    out = np.empty((count, n * ch), dtype=np.int16)
    free0, _ = torch.cuda.mem_get_info()
    info, status, views = e2.decode_all_batch(streams, [n * ch] * count, out=out)
    free1, _ = torch.cuda.mem_get_info()
    assert not status.any()
    batch_bytes = count * (size + n * ch * 2)
    assert batch_bytes > 2.2e9 and free0 - free1 < 0.45 * batch_bytes, (free0 - free1, batch_bytes)
    for i in (0, 179, 359):
        assert np.array_equal(out[i], ob.decode_all(streams[i].tobytes())[1]), i
    st, enc = e2.encode_all_batch(list(out), [(ch, rate, n)] * count)        # 1.9 GB of PCM: time-chunked pipeline
    free2, _ = torch.cuda.mem_get_info()
    assert st == [0] * count and free0 - free2 < 0.6 * batch_bytes, (free0 - free2, batch_bytes)
    for i in (0, 359):
        assert enc[i].tobytes() == ob.encode_all(out[i], ch, rate, n)[1], i
    e2.close()


def test_batch_tensor_rows_move_as_strided_copies(eng):
    """Equal-length streams held as rows of one array (input and output) travel as single strided copies
    (cudaMemcpy2DAsync) instead of one copy per stream; results must not depend on how the bytes travelled."""
    ch, rate, n, count = 2, 32000, 5120 * 2 + 1234, 24
    pcm = np.stack([signals.synth_pcm(21, i, ch, n, "noise" if i % 5 == 2 else "music") for i in range(count)])
    st, outs = eng.encode_all_batch(list(pcm), [(ch, rate, n)] * count)      # rows of `pcm`: strided H2D
    assert st == [0] * count
    size = ob.encoded_size(ch, n)
    rows = np.zeros((count, (size + 15) // 16 * 16), dtype=np.uint8)
    for i in range(count):
        assert outs[i].tobytes() == ob.encode_all(pcm[i], ch, rate, n)[1], i
        rows[i, :size] = outs[i]
    dec = np.full((count, n * ch + 24), 77, dtype=np.int16)
    info, status, views = eng.decode_all_batch([rows[i, :size] for i in range(count)], [n * ch] * count, out=dec)
    assert not status.any() and (info["samples"] == n * ch).all()
    for i in range(count):
        assert np.array_equal(dec[i, :n * ch], ob.decode_all(rows[i, :size].tobytes())[1]), i
    assert (dec[:, n * ch:] == 77).all()                                      # nothing written past a row's samples


@pytest.mark.parametrize("tmap", ["2", "1", "0"])
def test_stereo_decode_kernel_variants(tmap):
    """The mono / stereo kernel has two loops: warps whose 32 frames are full-size neighbours in one stream take the
    stripped uniform loop (one 32-row tensor store per group), every other warp the general loop (8-row boxes, row
    copies for irregular boxes).  QOA_B200_TMAP caps the tensor-map mode (read per launch; a subprocess keeps the
    environment clean): 2 = both loops, 1 = general loop only, 0 = no tensor stores at all.  All three must agree with
    the oracle on full frames, partial frames and slices, stream boundaries inside a warp and truncated streams."""
    import subprocess
    import sys
    code = """
import sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + '/tests')
import numpy as np, signals, qoaudio_b200
from oracle import binding as ob
eng = qoaudio_b200.default_engine()
lens = [5120, 5120 * 3, 5120 * 2 + 20, 5120 + 1, 5120 + 19, 20, 7, 4080, 2 * 5120 + 3082, 5120 * 40 + 119]
enc = [ob.encode_all(signals.synth_pcm(51, i, 2, n, 'noise' if i % 3 == 1 else 'music'), 2, 44100, n)[1] for i, n in enumerate(lens)]
long_ = [ob.encode_all(signals.synth_pcm(53, i, 2, 5120 * 70 + 333 * i), 2, 44100, 5120 * 70 + 333 * i)[1] for i in range(3)]  # 70 full frames: whole warps are uniform
streams = enc * 5 + long_ + [e[:len(e) - 11] for e in enc[:4]] + [open({root!r} + '/tests/golden/julien_baker_sprained_ankle.qoa', 'rb').read()]
mono = [ob.encode_all(signals.synth_pcm(52, i, 1, n + 1, 'noise' if i % 4 == 2 else 'music'), 1, 16000, n + 1)[1] for i, n in enumerate(lens)]
mono += [ob.encode_all(signals.synth_pcm(54, 0, 1, 5120 * 40 + 7), 1, 8000, 5120 * 40 + 7)[1]]
for batch in (streams, mono * 4 + [e[:len(e) - 9] for e in mono[:3]]):   # all-stereo, then all-mono: each takes its own kernel
    for mode in (0, 1):
        infos, outs = eng.decode_batch(batch, mode)
        for i, s in enumerate(batch):
            oi, opcm = (ob.decode_iter(s)[:2] if mode == 0 else ob.decode_all(s))
            assert infos[i]['status'] == oi['status'] and infos[i]['samples'] == oi['samples'], (mode, i)
            assert np.array_equal(outs[i], opcm), (mode, i)
print('variants ok')
"""
    from conftest import ROOT
    env = dict(os.environ, QOA_B200_TMAP=tmap)
    r = subprocess.run([sys.executable, "-c", code.format(root=ROOT)], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "variants ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("fast", ["0", "1"])
def test_encode_kernel_variants(fast):
    """Small batches take the latency-mode kernel (threshold search, lazily applied LMS update, exact redo of slices
    that leave its domain), large ones the throughput kernel.  Both are forced in turn here (QOA_B200_ENC_FAST is read
    once per process, hence the subprocess) on signals that do and do not trip the redo path."""
    import subprocess
    import sys
    code = """
import sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + '/tests')
import numpy as np, signals, qoaudio_b200
from oracle import binding as ob
eng = qoaudio_b200.default_engine()
pcm, descs = [], []
for i, (ch, n, kind) in enumerate([(2, 5120 * 3 + 777, 'music'), (1, 20000, 'noise'), (2, 30000, 'noise'), (6, 7001, 'music'),
                                   (1, 5120, 'music'), (3, 19, 'noise')]):
    pcm.append(signals.synth_pcm(41, i, ch, n, kind)); descs.append((ch, 44100, n))
for name, sig in signals.edge_signals().items():
    pcm.append(sig); descs.append((1, 8000, sig.size))
loud = np.tile(np.array([32767, -32768, 32767, 32767, -32768, -32768], dtype=np.int16), 4000)  # worst-case residuals
pcm.append(loud); descs.append((2, 48000, loud.size // 2))
st, outs = eng.encode_all_batch(pcm, descs)
assert st == [0] * len(descs)
for p, d, o in zip(pcm, descs, outs):
    assert o.tobytes() == ob.encode_all(p, *d)[1], d
print('variants ok')
"""
    from conftest import ROOT
    env = dict(os.environ, QOA_B200_ENC_FAST=fast)
    r = subprocess.run([sys.executable, "-c", code.format(root=ROOT)], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "variants ok" in r.stdout, r.stdout + r.stderr


def test_encode_all_batch_pipelined(eng, N):
    """Host batches above ~0.5 GB of PCM are encoded in time chunks on three CUDA streams with the LMS state carried on
    the device between chunks; the bytes must equal the one-shot result (= the oracle), invalid streams included."""
    import torch
    n_streams, ch, rate, n = 62, 2, 48000, 5120 * 850 + 777
    pcm = torch.empty(n_streams * ch * n, dtype=torch.int16, device="cuda")
    eng.gen_pcm_dev(pcm.data_ptr(), n_streams, 300, ch, n, 0, 2, "music")
    host = pcm.cpu().numpy().reshape(n_streams, -1)
    extra = [(1, 8000, 12345), (8, 96000, 5120 * 3 + 19), (3, 44100, 1)]
    pcm_list = [host[i] for i in range(n_streams)] + [signals.synth_pcm(5, 900 + j, c, m) for j, (c, r, m) in enumerate(extra)]
    descs = [(ch, rate, n)] * n_streams + extra
    pcm_list.append(np.zeros(10, np.int16))
    descs.append((2, 0, 5))  # InvalidSampleRate in the middle of a pipelined batch
    assert sum(p.size for p in pcm_list) >= 512 << 20
    st, outs = eng.encode_all_batch(pcm_list, descs)
    assert st == [0] * (n_streams + len(extra)) + [N.ENC_INVALID_SAMPLE_RATE]
    rc, ref = ob.mt_encode(host.reshape(-1), n_streams, ch, rate, n, 16)
    assert rc == 0
    size = ob.encoded_size(ch, n)
    ref = ref.reshape(n_streams, size)
    for i in range(n_streams):
        assert outs[i].size == size and np.array_equal(outs[i], ref[i]), i
    for j, d in enumerate(extra):
        assert outs[n_streams + j].tobytes() == ob.encode_all(pcm_list[n_streams + j], *d)[1], d


def test_multi_device_context(eng, N):
    """qoa_b200_ctx_create_multi: one process, several GPUs, batches sharded by stream (contiguous, balanced by work),
    results in caller order and identical to the single-device path and to the oracle.  With one visible GPU the same
    device is listed twice: two contexts, two host threads, same sharding logic."""
    import torch
    import qoaudio_b200 as q
    ndev = torch.cuda.device_count()
    devices = list(range(ndev)) if ndev >= 2 else [0, 0]
    m = q.MultiEngine(devices)
    assert m.devices == devices
    shapes = [(1 + i % 8, [8000, 44100, 48000][i % 3], n) for i, n in
              enumerate([1, 19, 5120, 5121, 30000, 777, 12345, 20, 65000, 2 * 5120 + 3, 40000, 9])]
    pcm = [signals.synth_pcm(61, i, c, n, "noise" if i % 4 == 3 else "music") for i, (c, r, n) in enumerate(shapes)]
    shapes.insert(5, (0, 44100, 10))       # a failing stream in the middle of a shard
    pcm.insert(5, np.zeros(10, np.int16))
    st, enc = m.encode_all_batch(pcm, shapes)
    st1, enc1 = eng.encode_all_batch(pcm, shapes)
    assert st == st1 and st[5] == N.ENC_INVALID_CHANNELS
    for i, (p, s) in enumerate(zip(pcm, shapes)):
        if st[i] == 0:
            assert enc[i].tobytes() == enc1[i].tobytes() == ob.encode_all(p, *s)[1], s
    streams = [e for i, e in enumerate(enc) if st[i] == 0] + [np.frombuffer(b"nope" + bytes(60), dtype=np.uint8)]
    caps = [s[0] * s[2] for i, s in enumerate(shapes) if st[i] == 0] + [16]
    info, dst, outs = m.decode_all_batch(streams, caps)
    info1, dst1, outs1 = eng.decode_all_batch(streams, caps)
    assert (dst == dst1).all() and dst[-1] == N.DEC_NOT_QOA_FILE and (info == info1).all()
    for i, s in enumerate(streams[:-1]):
        assert np.array_equal(outs[i], outs1[i]) and np.array_equal(outs[i], ob.decode_all(s.tobytes())[1]), i
    infos2, outs2 = m.decode_batch(streams)          # sizes from qoa_b200_probe (host-side header walk), iterator mode
    for i, s in enumerate(streams):
        oi, opcm = ob.decode_iter(s.tobytes())[:2]
        assert infos2[i]["status"] == oi["status"] and infos2[i]["samples"] == oi["samples"] and np.array_equal(outs2[i], opcm), i
    # the split is the documented one: contiguous, balanced by encoded length
    parts = m.partition([len(s) for s in streams])
    assert parts[0][0] == 0 and parts[-1][1] == len(streams) and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    assert m.kernel_launches() > 0
    # big enough to take the pipelined host path on every device
    big = [np.frombuffer(ob.encode_all(signals.synth_pcm(62, i, 2, 5120 * 40), 2, 44100, 5120 * 40)[1], dtype=np.uint8)
           for i in range(4)] * 48
    info, dst, outs = m.decode_all_batch(big, [2 * 5120 * 40] * len(big))
    ref = [ob.decode_all(b.tobytes())[1] for b in big[:4]]
    assert (dst == 0).all() and all(np.array_equal(o, ref[i % 4]) for i, o in enumerate(outs))
    m.close()


def test_full_size_decode_batch_properties(eng, N):
    """BASELINE.json configs[2] at FULL size (8192 x 60 s stereo, 104 GB of traffic) through size-independent
    properties: (1) per-stream checksums of the whole batch decoded in one plan equal those of the same streams decoded
    in three uneven sub-batches (different frame-table offsets, so different warps / loops see every frame); (2) decoding
    twice is idempotent; (3) a random handful of streams equals the oracle sample for sample; (4) bytes around every
    stream's slot stay untouched."""
    import torch
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    S = 8192 if free > 150e9 else 1024
    ch, rate, n = 2, 44100, 2646000
    size = 8 + ((n + 5119) // 5120) * (8 + 16 * ch) + ((n + 19) // 20) * 8 * ch
    stride = (size + 15) // 16 * 16
    per = n * ch
    slot = per + 24                                        # 24 guard values behind every stream (keeps 16-byte alignment)
    stream = torch.cuda.current_stream().cuda_stream
    qoa = torch.empty(S * stride, dtype=torch.uint8, device="cuda")
    eng.gen_qoa_dev(qoa.data_ptr(), stride, S, 0, ch, rate, n, 7, stream)
    in_off = np.arange(S, dtype=np.uint64) * np.uint64(stride)
    lens = np.full(S, size, dtype=np.uint64)
    out_off = np.arange(S, dtype=np.uint64) * np.uint64(slot)
    caps = np.full(S, per, dtype=np.uint64)
    w = torch.arange(1, 4097, dtype=torch.int32, device="cuda")

    def checksums(buf):                                    # position-weighted sums per stream, int64, chunked to bound memory
        v = buf.view(S, slot)
        out = torch.zeros(S, dtype=torch.int64, device="cuda")
        step = 4096 * 16                                   # 2 GB of int32 temporaries per block at S = 8192
        for c0 in range(0, per, step):
            c1 = min(per, c0 + step)
            blk = v[:, c0:c1].to(torch.int32)
            pad = (-blk.shape[1]) % 4096
            if pad:
                blk = torch.nn.functional.pad(blk, (0, pad))
            out += (blk.view(S, -1, 4096) * w).sum(dim=(1, 2), dtype=torch.int64) * (1 + c0 // step)
            del blk
        return out, v[:, per:].clone()

    def decode(ranges, fill):
        pcm = torch.full((S * slot,), fill, dtype=torch.int16, device="cuda")
        for b, e in ranges:
            plan = eng.decode_plan_dev(qoa.data_ptr(), in_off[b:e], lens[b:e], N.DECODE_ALL, stream)
            st = plan.run_dev(pcm.data_ptr(), out_off[b:e], caps[b:e], stream)
            assert (st == 0).all() and (plan.info_array["samples"] == per).all()
            plan.close()
        torch.cuda.synchronize()
        return pcm

    pcm = decode([(0, S)], 1234)
    sums, guards = checksums(pcm)
    assert (guards == 1234).all()
    rng = np.random.default_rng(4)
    for s in sorted(set([0, S - 1] + [int(v) for v in rng.integers(0, S, size=6)])):
        blob = qoa[s * stride:s * stride + size].cpu().numpy().tobytes()
        assert np.array_equal(pcm[s * slot:s * slot + per].cpu().numpy(), ob.decode_all(blob)[1]), s
    del pcm
    cut1, cut2 = S // 3 + 5, (2 * S) // 3 + 11             # uneven thirds: frame-table offsets that are no multiple of 32
    pcm2 = decode([(0, cut1), (cut1, cut2), (cut2, S)], -77)
    sums2, guards2 = checksums(pcm2)
    assert torch.equal(sums, sums2) and (guards2 == -77).all()
    del pcm2
    pcm3 = decode([(0, S)], 5)                             # idempotence
    assert torch.equal(checksums(pcm3)[0], sums)


def test_full_size_encode_chunking_property(eng, N):
    """BASELINE.json configs[3] shape at full width (4096 stereo 48 kHz streams): 32 frames of every stream encoded as one
    chunk, as 2 x 16 frames and as 5 + 27 frames (LMS state carried on the device, src/lib.rs:481-482) give identical
    bytes for ALL streams; streams 0, 1777 and last equal the oracle."""
    import torch
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    S = 4096 if free > 60e9 else 512
    ch, rate, frames = 2, 48000, 32
    n = frames * 5120
    fbytes = 8 + 16 * ch + 2048 * ch
    stream = torch.cuda.current_stream().cuda_stream
    pcm = torch.empty(S * n * ch, dtype=torch.int16, device="cuda")
    eng.gen_pcm_dev(pcm.data_ptr(), S, 0, ch, n, 0, 3, "music", stream)
    p_off = np.arange(S, dtype=np.uint64) * np.uint64(n * ch)
    o_off = np.arange(S, dtype=np.uint64) * np.uint64(frames * fbytes)
    results = []
    for cuts in ([32], [16, 16], [5, 27]):
        out = torch.zeros(S * frames * fbytes, dtype=torch.uint8, device="cuda")
        batch = eng.encoder_batch([(ch, rate, n)] * S)
        done = 0
        for f in cuts:
            batch.encode_chunk_dev(pcm.data_ptr(), p_off + np.uint64(done * 5120 * ch), np.full(S, f * 5120, dtype=np.uint32),
                                   out.data_ptr(), o_off + np.uint64(done * fbytes), stream)
            done += f
        torch.cuda.synchronize()
        batch.close()
        results.append(out)
    assert torch.equal(results[0], results[1]) and torch.equal(results[0], results[2])
    host = pcm.view(S, n * ch)
    for s in sorted(set([0, 1777 % S, S - 1])):
        ref = ob.encode_all(host[s].cpu().numpy(), ch, rate, n)[1]
        assert results[0][s * frames * fbytes:(s + 1) * frames * fbytes].cpu().numpy().tobytes() == bytes(ref)[8:], s
