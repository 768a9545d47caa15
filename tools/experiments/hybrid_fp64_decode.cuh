// EXPERIMENT (round 2), not compiled into the library: stereo decode with the left chain on the FP64 pipe.
// Bit-exact against the oracle on random-payload streams, but slower than the all-integer kernel on B200:
// 33.9 ms against 27.5 ms for 8192 x 60 s stereo (ncu, 1280 streams: issue active 64.8 %, FP64 pipe 32 %, ALU 47 %,
// FMA-heavy 35 %, new top stall = dispatch_stall 1.39 per issue; 36 instructions per sample pair against 32).
// tools/pipe_probe.cu: DFMA issues at the IMAD rate (18.4 T thread-ops/s) on its own pipe, IMAD:DFMA 1:1 reaches 95 % of
// the issue rate, but the three-pipe mix of this loop (8 IMAD : 9 FP64 : 12 ALU) only 81 %.  Kept for the record.
// It plugged into decode_fast_kernel as PairLoopD<0>::run(ad, b, row_a, row_b, wa, wb, o) with a second 16 KiB table.

// ---- hybrid stereo step: the right chain on the integer pipes, the left chain on the FP64 pipe ---------------------
// The stereo loop saturates the FMA-heavy pipe (16 IMAD per sample pair at one warp instruction per two cycles) at the
// same moment as the issue port.  B200 has an FP64 pipe of the same width that the loop never touches, and the LMS
// arithmetic of a frame is exact in binary64:
//   * weights: |w| <= 32768 + 65535 steps x 896 < 2^26, history |h| <= 2^15, so sum(w * h) is an integer below 2^43;
//     accumulated by four DFMAs on top of 2^52 + 2^51, the low word of the result IS the wrapping i32 sum of
//     QoaLms::predict (src/lib.rs:799-805) - no conversion instruction;
//   * the reconstructed sample enters the history as a double without a conversion either: the low half of the packed
//     (L, R) word, sign bit flipped, under the exponent 0x433 is the HIGH word of the double 2^52 + (r + 32768) * 2^32
//     (low word zero), and one DADD of -(2^52 + 2^47) leaves r * 2^32 exactly - again with a zero low word, so the
//     history doubles are edited in place, high word only (weights and deltas carry the matching factor 2^-32);
//   * sign(h) as +-1.0 is the sign bit of that double OR-ed into the high word of 1.0 (zero counts as positive,
//     :823-826), so update() is four DFMAs.
// Per sample pair: 8 IMAD + 9 FP64 + 12 ALU instead of 16 IMAD + 13 ALU: no pipe above 75 % of the issue rate.
// History and signs rotate by index (I & 3) instead of moving, so no 64-bit register moves are needed.
struct ChainD {
    double W[4]; // weights * 2^-32
    double H[4]; // history * 2^32 (low word always zero), slot (I + j) & 3 holds tap j at sample I
    double S[4]; // +-1.0 by the sign of the history value in the same slot (low word always zero)
};

constexpr int kLutDBytes = 128 * 8 * 16; // entry (sf * 8 + code), 8 copies (lane & 7) of {dq, dq >> 4, delta_d.lo, delta_d.hi}

__device__ __forceinline__ void fill_lut_d(uint8_t *lut)
{
    int4 *v = reinterpret_cast<int4 *>(lut);
    for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {
        const int dq = c_dequant[i >> 3];
        const double dd = (double)(dq >> 4) * (1.0 / 4294967296.0);
        v[i] = make_int4(dq, dq >> 4, __double2loint(dd), __double2hiint(dd));
    }
    __syncthreads();
}

__device__ __forceinline__ void load_lms_d(ChainD &c, const uint8_t *p)
{
    const uint2 h = ldg_be64(p);
    const uint2 w = ldg_be64(p + 8);
    const int hv[4] = {(int)h.x >> 16, (int)(h.x << 16) >> 16, (int)h.y >> 16, (int)(h.y << 16) >> 16};
    const int wv[4] = {(int)w.x >> 16, (int)(w.x << 16) >> 16, (int)w.y >> 16, (int)(w.y << 16) >> 16};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        c.W[j] = (double)wv[j] * (1.0 / 4294967296.0);
        c.H[j] = (double)hv[j] * 4294967296.0;
        c.S[j] = hv[j] < 0 ? -1.0 : 1.0;
    }
}

// Replaces the high word of a double in place by (x & m) ^ c / (x & m) | c; the low word stays in its register.
__device__ __forceinline__ void set_hi_and_xor(double &d, uint32_t x, uint32_t m, uint32_t c)
{
    asm("{ .reg .b32 lo, hi; mov.b64 {lo, hi}, %0; lop3.b32 hi, %1, %2, %3, 0x6a; mov.b64 %0, {lo, hi}; }"
        : "+d"(d)
        : "r"(x), "r"(m), "r"(c));
}
__device__ __forceinline__ void set_hi_and_or(double &d, uint32_t x, uint32_t m, uint32_t c)
{
    asm("{ .reg .b32 lo, hi; mov.b64 {lo, hi}, %0; lop3.b32 hi, %1, %2, %3, 0xea; mov.b64 %0, {lo, hi}; }"
        : "+d"(d)
        : "r"(x), "r"(m), "r"(c));
}

// a = left channel (low half of the packed word) on the FP64 pipe, b = right channel on the integer pipes.
template <int I>
__device__ __forceinline__ uint32_t pair_core_d(ChainD &a, Chain &b, int4 ta, int2 tb)
{
    double acc = fma(a.W[0], a.H[(I + 0) & 3], 6755399441055744.0); // 2^52 + 2^51; predict :799-805
    acc = fma(a.W[1], a.H[(I + 1) & 3], acc);
    acc = fma(a.W[2], a.H[(I + 2) & 3], acc);
    acc = fma(a.W[3], a.H[(I + 3) & 3], acc);
    const int xa = (__double2loint(acc) >> 13) + ta.x;                                     // :310
    const int xb = ((b.w0 * b.h0 + b.w1 * b.h1 + b.w2 * b.h2 + b.w3 * b.h3) >> 13) + tb.x;
    uint32_t packed; // (clamp_i16(xb) << 16) | (clamp_i16(xa) & 0xffff)
    asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(packed) : "r"(xb), "r"(xa));
    const int rb = (int)packed >> 16;
    b.w0 += b.s0 * tb.y; // update :821-828
    b.w1 += b.s1 * tb.y;
    b.w2 += b.s2 * tb.y;
    b.w3 += b.s3 * tb.y;
    b.h0 = b.h1; b.h1 = b.h2; b.h2 = b.h3; b.h3 = rb;
    b.s0 = b.s1; b.s1 = b.s2; b.s2 = b.s3; b.s3 = sign_pm1(rb);
    const double dd = __hiloint2double(ta.w, ta.z);
#pragma unroll
    for (int j = 0; j < 4; ++j) a.W[j] = fma(a.S[(I + j) & 3], dd, a.W[j]);
    set_hi_and_xor(a.H[I & 3], packed, 0x0000ffffu, 0x43308000u); // 2^52 + (ra + 32768) * 2^32
    a.H[I & 3] -= 4503599627370496.0 + 140737488355328.0;        // 2^52 + 2^47  ->  ra * 2^32
    set_hi_and_or(a.S[I & 3], (uint32_t)__double2hiint(a.H[I & 3]), 0x80000000u, 0x3ff00000u);
    return packed;
}

template <int I>
struct PairLoopD {
    static __device__ __forceinline__ void run(ChainD &a, Chain &b, uint32_t row_a, uint32_t row_b, uint2 wa, uint2 wb,
                                               uint32_t *o)
    {
        const int4 ta = lds_v4(row_a | code_off<I>(wa.x, wa.y));
        const int2 tb = lds_v2(row_b | code_off<I>(wb.x, wb.y));
        o[I] = pair_core_d<I>(a, b, ta, tb);
        PairLoopD<I + 1>::run(a, b, row_a, row_b, wa, wb, o);
    }
};
template <>
struct PairLoopD<20> {
    static __device__ __forceinline__ void run(ChainD &, Chain &, uint32_t, uint32_t, uint2, uint2, uint32_t *) {}
};

