"""GPU tests (-m gpu) that check the CUDA kernels' output with the vectorised verifier (oracle/vector_verify.py) running
on the same device through torch's own elementwise kernels: EVERY sample of a decode and EVERY slice of an encode is
checked against the reference's recurrences (src/lib.rs:291-330, :495-596, :613-617, :797-828) at batch sizes the
sequential CPU oracle could not cover in the test budget.  The verifier itself is pinned on the CPU
(tests/test_vector_verify.py: it agrees with the C oracle slice for slice, and the reference's fixture satisfies it
including the encoder states embedded in the frame headers)."""
import numpy as np
import pytest

import verify_chunks as vc

pytestmark = pytest.mark.gpu

CLEAN_DECODE = ("bad_file_headers", "bad_frame_headers", "bad_samples", "bad_streams")
CLEAN_CHAIN = ("chain_history_bad", "chain_weights_bad", "chain_bad_streams")
CLEAN_ENCODE = CLEAN_DECODE + CLEAN_CHAIN + ("bad_slices", "bad_initial_state")


@pytest.fixture(scope="module")
def eng():
    import qoaudio_b200
    return qoaudio_b200.default_engine()


@pytest.fixture(scope="module")
def N():
    import qoaudio_b200._native as n
    return n


def assert_clean_encode(r, kind):
    """Every sample, slice and initial state must verify.  The frame-to-frame chain is exact unless a weight left the
    16-bit range at a frame boundary (src/lib.rs:459-460 truncates what :482 carries in full; never seen on these
    signals, 4096 noise boundaries checked on the CPU): behind such a boundary encoder and decoder legitimately differ,
    the slices there are counted as unverifiable and the chain counters are not asserted."""
    assert all(r[k] == 0 for k in CLEAN_DECODE + ("bad_slices", "bad_initial_state")), r
    if r["truncated_boundaries"] == 0:
        assert all(r[k] == 0 for k in CLEAN_CHAIN) and r["unverifiable_slices"] == 0, r
    if kind == "music":
        assert r["truncated_boundaries"] == 0, r


def enc_size(ch, n):
    return 8 + ((n + 5119) // 5120) * (8 + 16 * ch) + ((n + 19) // 20) * 8 * ch


def test_fixture_gpu_decode_matches_the_encoder_state_in_its_headers(eng, fixture_bytes):
    """The reference's fixture was written by somebody else's encoder, which stored its LMS state in every frame header.
    The GPU's samples must be the fixed point of the decode recurrences AND leave, at the end of every frame, exactly the
    state the next header holds: 467 boundaries x 2 channels x 128 bits held by the reference's own file."""
    import torch
    infos, outs = eng.decode_batch([fixture_bytes], 1)
    assert infos[0]["status"] == 0
    ch, n, rate = infos[0]["channels"], infos[0]["header_samples"], infos[0]["sample_rate"]
    qoa = torch.frombuffer(bytearray(fixture_bytes), dtype=torch.uint8).unsqueeze(0).cuda()
    pcm = torch.from_numpy(outs[0].copy()).unsqueeze(0).cuda()
    r = vc.decode_in_chunks(qoa, pcm, ch, n, rate)
    assert all(r[k] == 0 for k in CLEAN_DECODE + CLEAN_CHAIN), r
    assert r["samples"] == 4788244 and r["boundaries"] == 467 * 2


def test_full_size_decode_every_sample_of_every_fourth_stream(eng, N):
    """BASELINE.json configs[2] (8192 x 60 s stereo, device-resident, one plan): all 5,292,000 samples of every fourth
    stream (and of the first and last eight) are checked against the recurrences on the device — 10.9 G samples, not a
    handful of spot checks.  (The generator's streams carry random LMS states in their headers, so frames do not chain;
    the chain is covered by the fixture and by the encode test below.)"""
    import torch
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    S = 8192 if free > 150e9 else 1024 if free > 24e9 else 128 if free > 4e9 else 12
    ch, rate, n = 2, 44100, 2646000
    size = enc_size(ch, n)
    stride = (size + 15) // 16 * 16
    per = n * ch
    stream = torch.cuda.current_stream().cuda_stream
    qoa = torch.empty(S * stride, dtype=torch.uint8, device="cuda")
    pcm = torch.full((S * per,), 77, dtype=torch.int16, device="cuda")
    eng.gen_qoa_dev(qoa.data_ptr(), stride, S, 0, ch, rate, n, 11, stream)
    in_off = np.arange(S, dtype=np.uint64) * np.uint64(stride)
    lens = np.full(S, size, dtype=np.uint64)
    out_off = np.arange(S, dtype=np.uint64) * np.uint64(per)
    caps = np.full(S, per, dtype=np.uint64)
    plan = eng.decode_plan_dev(qoa.data_ptr(), in_off, lens, N.DECODE_ALL, stream)
    st = plan.run_dev(pcm.data_ptr(), out_off, caps, stream)
    torch.cuda.synchronize()
    assert (st == 0).all() and (plan.info_array["samples"] == per).all()
    plan.close()
    pick = sorted(set(range(0, S, 4)) | set(range(min(8, S))) | set(range(max(0, S - 8), S)))
    r = vc.decode_in_chunks(qoa.view(S, stride), pcm.view(S, per), ch, n, rate, streams=pick, chunk=8)
    assert all(r[k] == 0 for k in CLEAN_DECODE), r
    assert r["streams"] == len(pick) and r["samples"] == len(pick) * per


@pytest.mark.parametrize("kind,streams", [("music", 4096), ("noise", 512)])
def test_full_width_encode_every_slice(eng, N, kind, streams):
    """BASELINE.json configs[3]'s width (4096 stereo 48 kHz streams; 512 for white noise), 8 frames + a partial frame with
    a partial last slice per stream: GPU encode -> GPU decode -> every sample is the fixed point of the decode recurrences,
    every frame header carries the encoder's state, and every one of the slices is the lexicographic arg-min of
    (total rank, first-sample rank, scalefactor) over all 16 candidates (src/lib.rs:495-596)."""
    import torch
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    S = streams if free > 40e9 else min(streams, 256) if free > 4e9 else 24
    ch, rate, n = 2, 48000, 5120 * 8 + 333
    size = enc_size(ch, n)
    stride = (size + 15) // 16 * 16
    per = n * ch
    stream = torch.cuda.current_stream().cuda_stream
    src = torch.empty(S * per, dtype=torch.int16, device="cuda")
    eng.gen_pcm_dev(src.data_ptr(), S, 0, ch, n, 0, 4, kind, stream)
    qoa = torch.zeros(S * stride, dtype=torch.uint8, device="cuda")
    offs = np.arange(S, dtype=np.uint64)
    st, out_len = eng.encode_all_batch_dev([(ch, rate, n)] * S, src.data_ptr(), offs * np.uint64(per),
                                           np.full(S, per, dtype=np.uint64), qoa.data_ptr(), offs * np.uint64(stride),
                                           np.full(S, size, dtype=np.uint64), stream)
    assert st == [0] * S and out_len == [size] * S
    dec = torch.full((S * per,), -5, dtype=torch.int16, device="cuda")
    plan = eng.decode_plan_dev(qoa.data_ptr(), offs * np.uint64(stride), np.full(S, size, dtype=np.uint64), N.DECODE_ALL, stream)
    dst = plan.run_dev(dec.data_ptr(), offs * np.uint64(per), np.full(S, per, dtype=np.uint64), stream)
    torch.cuda.synchronize()
    assert (dst == 0).all() and (plan.info_array["samples"] == per).all()
    plan.close()
    r = vc.encode_in_chunks(qoa.view(S, stride), src.view(S, per), dec.view(S, per), ch, n, rate, chunk=128,
                            max_lanes=1 << 25)
    assert_clean_encode(r, kind)
    assert r["streams"] == S and r["slices"] == S * ch * ((n + 19) // 20) and r["boundaries"] == S * ch * 8


@pytest.mark.parametrize("ch", [1, 3, 4, 5, 6, 7, 8])
def test_long_streams_every_sample_other_channel_counts(eng, N, ch):
    """The mono fast kernel and the 3..8-channel kernels on long streams (40 full frames + a partial one with a partial
    last slice: uniform-warp loop / tensor-store path, stream boundaries inside warps, the row-copy path for the
    tails): every sample of every stream against the recurrences."""
    import torch
    free, _ = torch.cuda.mem_get_info()
    S = 96 if free > 4e9 else 5
    rate, n = 48000, 5120 * 40 + 1234
    size = enc_size(ch, n)
    stride = (size + 15) // 16 * 16
    per = (n * ch + 7) // 8 * 8                              # 16-byte aligned slots
    stream = torch.cuda.current_stream().cuda_stream
    qoa = torch.empty(S * stride, dtype=torch.uint8, device="cuda")
    pcm = torch.full((S * per,), 99, dtype=torch.int16, device="cuda")
    eng.gen_qoa_dev(qoa.data_ptr(), stride, S, 0, ch, rate, n, 20 + ch, stream)
    offs = np.arange(S, dtype=np.uint64)
    plan = eng.decode_plan_dev(qoa.data_ptr(), offs * np.uint64(stride), np.full(S, size, dtype=np.uint64), N.DECODE_ALL, stream)
    st = plan.run_dev(pcm.data_ptr(), offs * np.uint64(per), np.full(S, n * ch, dtype=np.uint64), stream)
    torch.cuda.synchronize()
    assert (st == 0).all() and (plan.info_array["samples"] == n * ch).all()
    plan.close()
    r = vc.decode_in_chunks(qoa.view(S, stride), pcm.view(S, per), ch, n, rate, chunk=8)
    assert all(r[k] == 0 for k in CLEAN_DECODE), r
    assert r["streams"] == S and r["samples"] == S * n * ch
    assert (pcm.view(S, per)[:, n * ch:] == 99).all()        # nothing written behind a stream


@pytest.mark.parametrize("ch,kind", [(1, "music"), (3, "noise"), (8, "music")])
def test_encode_every_slice_other_channel_counts(eng, N, ch, kind):
    """Mono, 3- and 8-channel encodes (chains = streams x channels, interleaved input): every slice is the arg-min
    candidate, every header chains."""
    import torch
    free, _ = torch.cuda.mem_get_info()
    S = 192 if free > 4e9 else 6
    rate, n = 44100, 5120 * 3 + 4097
    size = enc_size(ch, n)
    stride = (size + 15) // 16 * 16
    per = n * ch
    stream = torch.cuda.current_stream().cuda_stream
    src = torch.empty(S * per, dtype=torch.int16, device="cuda")
    eng.gen_pcm_dev(src.data_ptr(), S, 0, ch, n, 0, 6, kind, stream)
    qoa = torch.zeros(S * stride, dtype=torch.uint8, device="cuda")
    offs = np.arange(S, dtype=np.uint64)
    st, out_len = eng.encode_all_batch_dev([(ch, rate, n)] * S, src.data_ptr(), offs * np.uint64(per),
                                           np.full(S, per, dtype=np.uint64), qoa.data_ptr(), offs * np.uint64(stride),
                                           np.full(S, size, dtype=np.uint64), stream)
    assert st == [0] * S and out_len == [size] * S
    dper = (per + 7) // 8 * 8
    dec = torch.full((S * dper,), -5, dtype=torch.int16, device="cuda")
    plan = eng.decode_plan_dev(qoa.data_ptr(), offs * np.uint64(stride), np.full(S, size, dtype=np.uint64), N.DECODE_ALL, stream)
    dst = plan.run_dev(dec.data_ptr(), offs * np.uint64(dper), np.full(S, per, dtype=np.uint64), stream)
    torch.cuda.synchronize()
    assert (dst == 0).all() and (plan.info_array["samples"] == per).all()
    plan.close()
    r = vc.encode_in_chunks(qoa.view(S, stride), src.view(S, per), dec.view(S, dper), ch, n, rate, chunk=64, max_lanes=1 << 24)
    assert_clean_encode(r, kind)
    assert r["streams"] == S and r["slices"] == S * ch * ((n + 19) // 20)
