"""Long differential fuzz of the CUDA path against the oracle (development aid; the committed tests hold a short,
seeded version of the same thing).  Usage: python tools/fuzz_gpu.py [rounds] [seed]
Each round: random streams (1-8 channels, lengths 1..12000, signals from silence to full-scale square waves) are
encoded in one small batch (latency-mode kernel) and once more inside a large batch (throughput kernel), compared byte
for byte with the oracle, then decoded — intact, truncated and corrupted — and compared sample for sample."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qoaudio_b200 as q
from oracle import binding as ob

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 5
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
eng = q.default_engine()


def signal(n):
    kind = rng.integers(0, 9)
    t = np.arange(n)
    if kind == 0:
        v = np.zeros(n)
    elif kind == 1:
        v = rng.integers(-32768, 32768, n)
    elif kind == 2:
        v = np.where((t // rng.integers(1, 50)) % 2, 32767, -32768)
    elif kind == 3:
        v = 30000 * np.sin(t * rng.uniform(0.001, 3.0)) + rng.integers(-2000, 2000, n)
    elif kind == 4:
        v = np.full(n, rng.integers(-32768, 32768))
    elif kind == 5:
        v = np.where(rng.random(n) < 0.02, rng.choice([-32768, 32767], n), 0)
    elif kind == 6:
        v = np.cumsum(rng.integers(-900, 900, n))
    elif kind == 7:
        v = rng.integers(-40, 40, n)
    else:
        v = 32767 * np.sign(np.sin(t * 0.3)) * (t % 7 != 0)
    return np.clip(v, -32768, 32767).astype(np.int16)


t0 = time.time()
checked = 0
for r in range(rounds):
    pcm, descs = [], []
    for i in range(400):
        ch = int(rng.integers(1, 9))
        n = int(rng.choice([1, 19, 20, 21, 5119, 5120, 5121, int(rng.integers(1, 12000))]))
        pcm.append(signal(n * ch))
        descs.append((ch, int(rng.choice([8000, 44100, 48000, 16777215])), n))
    for i in range(4):  # long pathological streams: the LMS weights get time to run away
        n = int(rng.integers(100000, 300000))
        pcm.append(signal(n))
        descs.append((1, 44100, n))
    for ch in range(1, 9):  # long enough for whole warps of full frames: the uniform-warp / tensor-store decode paths
        n = int(rng.integers(5120 * 40, 5120 * 70))
        pcm.append(signal(n * ch))
        descs.append((ch, 48000, n))
    order = rng.permutation(len(pcm))
    pcm, descs = [pcm[k] for k in order], [descs[k] for k in order]
    expect = [ob.encode_all(p, *d)[1] for p, d in zip(pcm, descs)]
    st, outs = eng.encode_all_batch(pcm[:150], descs[:150])          # <= 1184 warps: latency-mode kernel
    assert not any(st)
    for k in range(150):
        assert outs[k].tobytes() == expect[k], ("fast", r, k, descs[k])
    st, outs = eng.encode_all_batch(pcm * 4, descs * 4)               # > 1184 warps: throughput kernel
    assert not any(st)
    for k in range(len(pcm)):
        assert outs[k].tobytes() == expect[k] and outs[k + 3 * len(pcm)].tobytes() == expect[k], ("exact", r, k, descs[k])
    streams = list(expect)
    for e in expect[:120]:
        b = bytearray(e)
        cut = int(rng.integers(0, len(b) + 1))
        streams.append(bytes(b[:cut]))
        if len(b) > 16:
            pos = int(rng.integers(0, len(b)))
            b[pos] ^= 1 << int(rng.integers(0, 8))
            streams.append(bytes(b))
    for mode in (0, 1):
        infos, dec = eng.decode_batch(streams, mode)
        for k, s in enumerate(streams):
            oi, opcm = (ob.decode_iter(s)[:2] if mode == 0 else ob.decode_all(s))
            got = {key: infos[k][key] for key in ("status", "streaming", "header_samples", "channels", "sample_rate", "samples", "frames")}
            assert got == oi, (r, mode, k, got, oi)
            assert np.array_equal(dec[k], opcm), (r, mode, k)
    checked += len(pcm) * 5 + 150 + 2 * len(streams)
    print(f"round {r}: ok ({checked} stream checks, {time.time() - t0:.0f} s)", flush=True)
print("fuzz ok")
