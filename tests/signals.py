"""Deterministic integer-only test signals shared by CPU and GPU parity tests."""
import numpy as np

M64 = (1 << 64) - 1


def splitmix64(x):
    """Vectorised splitmix64 finaliser over uint64 arrays (same as csrc/qoa_gen.cu)."""
    x = (np.asarray(x, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15))
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def stream_key(seed, stream, channel):
    with np.errstate(over="ignore"):
        return splitmix64(np.uint64(seed) * np.uint64(0x100000001B3) + np.uint64(stream) * np.uint64(257)
                          + np.uint64(channel))


def _tri(phase):
    ph = phase & np.uint32(0xFFFF)
    t = np.where(ph < 32768, ph, np.uint32(65535) - ph).astype(np.int64)  # 0..32767
    return t * 2 - 32767


def synth_pcm(seed, stream, channels, n, kind="music", start=0):
    """Interleaved int16 [n*channels] for samples start..start+n of the given stream."""
    out = np.empty((n, channels), dtype=np.int16)
    with np.errstate(over="ignore"):
        i = (np.arange(n, dtype=np.uint64) + np.uint64(start))
        for c in range(channels):
            k = stream_key(seed, stream, c)
            h = splitmix64(k + i)
            if kind == "noise":
                out[:, c] = (h & np.uint64(0xFFFF)).astype(np.uint16).view(np.int16)
                continue
            kk = int(k)
            f1 = 40 + (kk & 0x3FF)
            f2 = 300 + ((kk >> 10) & 0xFFF)
            a1 = 2000 + ((kk >> 22) & 0x1FFF)
            a2 = 500 + ((kk >> 35) & 0xFFF)
            i32 = i.astype(np.uint32)
            v = (_tri(i32 * np.uint32(f1)) * a1 >> 15) + (_tri(i32 * np.uint32(f2)) * a2 >> 15)
            v = v + (h >> np.uint64(52)).astype(np.int64) - 2048
            out[:, c] = np.clip(v, -32768, 32767).astype(np.int16)
    return out.reshape(-1)


def kat_signals():
    """The SURVEY.md Appendix B known-answer inputs: name -> (pcm int16, channels, rate, samples)."""
    saw = np.array([((i % 200) - 100) * 100 for i in range(1000)], dtype=np.int16)  # src/lib.rs:1194-1197
    sil = np.zeros(21, dtype=np.int16)
    mix = np.zeros(45 * 3, dtype=np.int64)
    for i in range(45):
        mix[3 * i] = i * 700 - 15000
        mix[3 * i + 1] = 32767 if (i & 1) else -32768
        mix[3 * i + 2] = (i * i * 37) % 4001 - 2000
    x = 12345
    lcg = np.zeros(5121 * 8, dtype=np.int16)
    vals = []
    for _ in range(5121 * 8):
        x = (x * 1664525 + 1013904223) & 0xFFFFFFFF
        vals.append((x >> 16) & 0xFFFF)
    lcg = np.array(vals, dtype=np.uint16).view(np.int16)
    return {
        "saw1000_mono": (saw, 1, 44100, 1000),
        "silence21_mono": (sil, 1, 8000, 21),
        "mix45_3ch": (mix.astype(np.int16), 3, 96000, 45),
        "lcg5121_8ch": (lcg, 8, 48000, 5121),
    }


def edge_signals(seed=7):
    """Small adversarial signals: constants, extremes, alternating full-scale, steps, ramps."""
    rng = np.random.default_rng(seed)
    sig = {
        "zeros": np.zeros(61, dtype=np.int16),
        "max": np.full(61, 32767, dtype=np.int16),
        "min": np.full(61, -32768, dtype=np.int16),
        "alt": np.array([32767 if i & 1 else -32768 for i in range(101)], dtype=np.int16),
        "step": np.concatenate([np.zeros(30), np.full(35, 30000), np.full(36, -30000)]).astype(np.int16),
        "ramp": np.arange(-32768, 32768, 257, dtype=np.int32).astype(np.int16),
        "uniform": rng.integers(-32768, 32768, size=401).astype(np.int16),
        "small": rng.integers(-3, 4, size=203).astype(np.int16),
        "one": np.array([-12345], dtype=np.int16),
    }
    return sig
