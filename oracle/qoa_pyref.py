"""Independent pure-Python restatement of the reference's QOA encode/decode arithmetic.

TEST INFRASTRUCTURE ONLY (see oracle/qoa_oracle.h).  Written separately from qoa_oracle.c, in a
different language with arbitrary-precision ints and explicit 32-bit wrapping, so that a shared
transcription slip between the two is unlikely.  Small inputs only (pure-Python loops).

The encoder here uses the *brute-force* formulation: evaluate all 16 scalefactors over the whole slice
and take the lexicographic minimum of (total_rank, first_sample_rank, sf) — SURVEY.md §0 fact 2 — whereas
qoa_oracle.c's default path follows the reference's sorted + pruned control flow (src/lib.rs:516-593).

Citations: /root/reference/src/lib.rs.
"""

QUANT_TAB = [7, 7, 7, 5, 5, 3, 3, 1, 0, 0, 2, 2, 4, 4, 6, 6, 6]  # :22
RECIP_TAB = [65536, 9363, 3121, 1457, 781, 475, 311, 216, 156, 117, 90, 71, 57, 47, 39, 32]  # :25-27
_MAG = [  # :847-864, magnitudes (a, b, c, d) per scalefactor; codes are +a,-a,+b,-b,+c,-c,+d,-d
    (1, 3, 5, 7), (5, 18, 32, 49), (16, 53, 95, 147), (34, 113, 203, 315),
    (63, 210, 378, 588), (104, 345, 621, 966), (158, 528, 950, 1477), (228, 760, 1368, 2128),
    (316, 1053, 1895, 2947), (422, 1405, 2529, 3934), (548, 1828, 3290, 5117), (696, 2320, 4176, 6496),
    (868, 2893, 5207, 8099), (1064, 3548, 6386, 9933), (1286, 4288, 7718, 12005), (1536, 5120, 9216, 14336),
]
DEQUANT_TAB = [[(-m if (q & 1) else m) for q in range(8) for m in [_MAG[sf][q >> 1]]] for sf in range(16)]

SLICE_LEN = 20
FRAME_LEN = 5120
MAGIC = 0x716F6166


def _i32(x):
    """wrap to signed 32-bit"""
    x &= 0xFFFFFFFF
    return x - (1 << 32) if x & 0x80000000 else x


def _i16(x):
    x &= 0xFFFF
    return x - (1 << 16) if x & 0x8000 else x


def predict(h, w):  # :799-805   (python >> on negative ints is arithmetic)
    return _i32(w[0] * h[0] + w[1] * h[1] + w[2] * h[2] + w[3] * h[3]) >> 13


def penalty_sq(w):  # :814-817
    pen = (_i32(w[0] * w[0] + w[1] * w[1] + w[2] * w[2] + w[3] * w[3]) >> 18) - 0x8FF
    pen = max(pen, 0)
    return pen * pen


def update(h, w, sample, residual):  # :821-828
    delta = residual >> 4
    w = [_i32(w[i] + (-delta if h[i] < 0 else delta)) for i in range(4)]
    h = [h[1], h[2], h[3], sample]
    return h, w


def qoa_div(v, sf):  # :613-617
    n = _i32(v * RECIP_TAB[sf] + (1 << 15)) >> 16
    return n + ((v > 0) - (v < 0)) - ((n > 0) - (n < 0))


def _candidate(h, w, samples, sf):
    """Run scalefactor sf over the slice; returns (total, first, word, h, w)."""
    word = sf
    total = 0
    first = 0
    for i, s in enumerate(samples):
        p = predict(h, w)
        pen = penalty_sq(w)
        scaled = qoa_div(s - p, sf)
        clamped = max(-8, min(8, scaled))
        q = QUANT_TAB[clamped + 8]
        dq = DEQUANT_TAB[sf][q]
        r = max(-32768, min(32767, p + dq))
        rank = (s - r) * (s - r) + pen
        if i == 0:
            first = rank
        total += rank
        h, w = update(h, w, r, dq)
        word = (word << 3) | q
    return total, first, word, h, w


def encode_slice(h, w, samples):
    """:495-596 as lexicographic arg-min.  Returns (unshifted word, sf, h, w)."""
    best = None
    for sf in range(16):
        total, first, word, nh, nw = _candidate(h, w, samples, sf)
        key = (total, first, sf)
        if best is None or key < best[0]:
            best = (key, word, sf, nh, nw)
    return best[1], best[2], best[3], best[4]


class Encoder:
    """QoaEncoder (:85-91, :333-493)."""

    def __init__(self, channels, sample_rate, samples):
        if channels == 0 or channels > 8:
            raise ValueError("InvalidChannels")
        if sample_rate == 0:
            raise ValueError("InvalidSampleRate")
        if samples == 0:
            raise ValueError("InvalidSamples")
        self.channels, self.sample_rate, self.samples = channels, sample_rate, samples
        self.h = [[0, 0, 0, 0] for _ in range(channels)]
        self.w = [[0, 0, -(1 << 13), 1 << 14] for _ in range(channels)]  # :348-351

    def header(self):  # :404-408
        return MAGIC.to_bytes(4, "big") + self.samples.to_bytes(4, "big")

    def frame(self, pcm):  # :418-493 ; pcm = flat interleaved list
        ch = self.channels
        if len(pcm) == 0 or len(pcm) % ch:
            raise ValueError("InvalidSamples")
        n = len(pcm) // ch
        if n > FRAME_LEN:
            raise ValueError("InvalidSamples")
        slices = -(-n // SLICE_LEN)
        fsize = (8 + 16 * ch + 8 * slices * ch) & 0xFFFF
        out = bytearray()
        out += ((ch << 56) | (self.sample_rate << 32) | (n << 16) | fsize).to_bytes(8, "big")
        for c in range(ch):
            for v in self.h[c]:
                out += (v & 0xFFFF).to_bytes(2, "big")
            for v in self.w[c]:
                out += (v & 0xFFFF).to_bytes(2, "big")
        for s0 in range(0, n, SLICE_LEN):
            ln = min(SLICE_LEN, n - s0)
            for c in range(ch):
                smp = [pcm[(s0 + i) * ch + c] for i in range(ln)]
                word, _sf, self.h[c], self.w[c] = encode_slice(self.h[c], self.w[c], smp)
                word <<= (SLICE_LEN - ln) * 3  # :485-487
                out += word.to_bytes(8, "big")
        return bytes(out)

    def encode(self, pcm):  # :367-398
        ch = self.channels
        if len(pcm) != self.samples * ch:
            raise ValueError("InvalidSamples")
        out = bytearray(self.header())
        for i in range(0, self.samples, FRAME_LEN):
            n = min(FRAME_LEN, self.samples - i)
            out += self.frame(pcm[i * ch:(i + n) * ch])
        return bytes(out)


def encode_all(pcm, channels, sample_rate, samples):  # :760-763
    return Encoder(channels, sample_rate, samples).encode(list(pcm))


def decode_frames(data):
    """Plain conforming-file decode (:205-330) walking by num_samples_per_channel.
    Returns (channels, rate, flat interleaved samples, [frame nspc...]).  No error semantics: asserts."""
    assert int.from_bytes(data[0:4], "big") == MAGIC
    pos = 8
    out = []
    frames = []
    ch0 = rate0 = None
    while len(data) - pos >= 8:
        hd = int.from_bytes(data[pos:pos + 8], "big")
        pos += 8
        ch, rate, nspc = hd >> 56, (hd >> 32) & 0xFFFFFF, (hd >> 16) & 0xFFFF
        if ch0 is None:
            ch0, rate0 = ch, rate
        assert (ch, rate) == (ch0, rate0)
        hs, ws = [], []
        for _c in range(ch):
            hs.append([_i16(int.from_bytes(data[pos + 2 * i:pos + 2 * i + 2], "big")) for i in range(4)])
            ws.append([_i16(int.from_bytes(data[pos + 8 + 2 * i:pos + 10 + 2 * i], "big")) for i in range(4)])
            pos += 16
        frames.append(nspc)
        remaining = nspc
        while remaining > 0:
            group = [0] * (SLICE_LEN * ch)
            for c in range(ch):
                word = int.from_bytes(data[pos:pos + 8], "big")
                pos += 8
                sf = word >> 60
                h, w = hs[c], ws[c]
                for i in range(SLICE_LEN):
                    p = predict(h, w)
                    q = (word >> (57 - 3 * i)) & 7
                    dq = DEQUANT_TAB[sf][q]
                    r = max(-32768, min(32767, p + dq))
                    group[i * ch + c] = r
                    h, w = update(h, w, r, dq)
                hs[c], ws[c] = h, w
            n = min(remaining, SLICE_LEN)
            out += group[:n * ch]
            remaining -= n
    return ch0, rate0, out, frames
