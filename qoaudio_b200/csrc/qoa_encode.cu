// qoa_encode.cu — batched QOA encoder: per-slice scalefactor search as a warp kernel (sm_100a).
//
// Reference semantics (rafaelcaricio/qoaudio, src/lib.rs):
//   QoaEncoder::encode_slice        :495-596   16-candidate search, sorted order, strict-< replacement
//   encode_frame_to_buf             :441-493   frame header, LMS dump (truncating `as u16`), slice order, partial slice
//   QoaLms::predict_and_penalty_sq  :808-818
//   QoaLms::update                  :821-828
//   qoa_div                         :613-617
//
// Parallel shape: LMS state flows slice -> slice and frame -> frame (:481-482), so one (stream, channel) "chain" is
// sequential.  A chain is one half-warp; lane = scalefactor candidate.  Every lane runs the full 20-sample candidate
// (the reference's pruning never changes the winner: ranks are monotone).  The winner is the lexicographic minimum of
// (total rank, first-sample rank, scalefactor) — exactly what the reference's stable insertion sort by first-sample
// rank (:535-541) followed by strict `<` replacement (:587) selects.
//
// Instruction diet (B200 issues IMAD on the FMA pipe and everything else on the ALU pipe, each at half rate, so the
// step is balanced between the two): the +/-1 sign of every history tap is kept in a register so the weight update is
// four IMADs; the LUT returns {dequantized, quantized, delta} in one 128-bit shared load; clamp/add pairs use the
// fused VIADDMNMX; the squared error accumulates through one IMAD.WIDE; the 3-bit codes accumulate in two 30-bit
// halves; PCM samples are staged de-interleaved in shared memory 16 slices at a time so the per-step fetch is one
// broadcast LDS with an immediate offset.
#include <mutex>

#include "qoa_kernels.h"

namespace qoa {

constexpr int kEncThreads = 128;                         // 4 warps = 8 chains per CTA
constexpr int kEncWarps = kEncThreads / 32;
constexpr int kChunkSlices = 16;                         // slices staged per refill
constexpr int kChunkSamples = kChunkSlices * kSliceLen;  // 320 samples per chain
constexpr uint32_t kFull = 0xffffffffu;
constexpr uint32_t kFastMaxWarps = 148 * 8;              // latency mode up to two warps per scheduler (measured crossover: see DESIGN.md)

// Candidate LUT: entry (idx, sf), idx = clamp(scaled, -8, 8) + 8, at byte (idx * 16 + sf) * 16 holding
// {dequantized, quantized, dequantized >> 4, 0}.  A quarter-warp (8 scalefactors) reads 8 distinct 16-byte bank groups
// whatever the indices are: conflict-free.
struct __align__(16) EncSmem {
    int4 lut[17 * 16];
    int16_t stage[kEncWarps][2][kChunkSamples];
};

struct Cand {
    int h0, h1, h2, h3;
    int w0, w1, w2, w3;
    int s0, s1, s2, s3; // +1 / -1: the sign update() derives from history (:823-826; zero counts as positive)
};

__device__ __forceinline__ int sign_pm1(int v) { return (v >> 31) | 1; }

// One (candidate, sample) step :564-585.  err accumulates error^2 (64-bit), pen accumulates penalty^2 (fits 32 bits:
// penalty < 2^13, 20 terms), codes accumulates the 3-bit quantized values.
// `zero` is an opaque 0 (a kernel argument): a three-input add can only issue as IADD3 on the ALU pipe, which keeps the
// two subtractions off the FMA pipe (the busier one in this step: 14 IMADs + the wide multiply-add).
__device__ __forceinline__ void candidate_step(Cand &l, int s, int recip, uint32_t lut_lane, long long &err, uint32_t &pen,
                                               uint32_t &codes, int zero)
{
    // predict_and_penalty_sq :808-818 (all sums wrap in i32)
    const int p = (l.w0 * l.h0 + l.w1 * l.h1 + l.w2 * l.h2 + l.w3 * l.h3) >> 13;
    const int ws = (l.w0 * l.w0 + l.w1 * l.w1 + l.w2 * l.w2 + l.w3 * l.w3) >> 18;
    const int penalty = __viaddmax_s32(ws, -0x8ff, 0);
    const int v = s - p + zero; // :567
    // qoa_div :613-617 — wrapping multiply, round half up, then push away from zero
    const int n = (int)((uint32_t)v * (uint32_t)recip + 32768u) >> 16;
    const int t = n + max(-1, min(1, v)) - max(-1, min(1, n));
    const int c = max(-8, min(8, t)); // :568-569
    int dq, q, d, unused;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(dq), "=r"(q), "=r"(d), "=r"(unused)
                 : "r"(lut_lane + (uint32_t)(c * 256))); // :570-572 (lut_lane is pre-biased by +8 entries)
    const int r = max(__viaddmin_s32(p, dq, 32767), -32768); // :573
    const int e = s - r + zero;                              // :575
    err += (long long)e * (long long)e;                      // :576
    pen += (uint32_t)penalty * (uint32_t)penalty;
    l.w0 += l.s0 * d; // update :821-828 (d = dequantized >> 4)
    l.w1 += l.s1 * d;
    l.w2 += l.s2 * d;
    l.w3 += l.s3 * d;
    l.h0 = l.h1; l.h1 = l.h2; l.h2 = l.h3; l.h3 = r;
    l.s0 = l.s1; l.s1 = l.s2; l.s2 = l.s3; l.s3 = sign_pm1(r);
    codes = codes * 8u + (uint32_t)q; // :584
}

__device__ __forceinline__ int lds_s16(uint32_t addr)
{
    int v;
    asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// ---- latency mode (few chains: every warp is alone on its scheduler and the step is a dependency chain) -------------
// For a fixed scalefactor the map residual -> (quantized code, dequantized value) of :567-572 (qoa_div, clamp, two
// table lookups) is a monotone step function of the residual with seven thresholds, as long as the 32-bit product in
// qoa_div does not wrap (|v| < 32768 for scalefactor 0, < 229356 for scalefactor 1, never for the others given
// |v| <= 32768 + 2^18).  Seven compares and a three-level select tree replace the multiply, the sign fix-up, the clamp
// and the ~30-cycle shared-memory lookup on the critical path.  The predictor is re-associated so that only three
// instructions sit between the previous sample and the residual: with negated weights nw = -w and
// K = 8192 * s + 8191,   s - ((sum w*h) >> 13) == (K + sum nw*h) >> 13   (exact while K + sum nw*h fits 32 bits),
// and the weight update of the previous step is applied lazily:  sum_{i<3} nw_i' * h_i = D + delta * NB  with D and NB
// built from values known one step earlier.  A step that leaves the exact domain (|v| >= lim, which also catches the
// 32-bit overflow of K + sum nw*h: the wrapped residual is then >= 229376 in magnitude) sets a sticky flag and the
// whole slice is redone with candidate_step(); on music that is ~1 % of the slices, on full-scale noise about half.
// The table is built ON THE DEVICE from the very expressions candidate_step() uses (fast_table_kernel).
struct FastTab {
    int b[7]; // ascending thresholds: the code changes at v == b[k]
    int x[8]; // (dequantized << 3) | quantized for v < b[0], b[0] <= v < b[1], ..., v >= b[6]
    int lim;  // exact for -lim <= v < lim
};
__device__ FastTab g_fast_tab[16];
__device__ int g_fast_ok;

// One CTA per scalefactor: the threads scan the residual domain in parallel and record every residual at which the
// quantized code changes (round 1 scanned serially: 22.8 ms on the first small-batch encode of a process).
__global__ void __launch_bounds__(256) fast_table_kernel()
{
    __shared__ int s_nb, s_v[8], s_q[8];
    const int sf = blockIdx.x;
    const int recip = c_recip[sf];
    const int lim = sf == 0 ? 32768 : 196608;
    auto quant = [&](int v) { // :567-572 exactly as candidate_step() evaluates it
        const int n = (int)((uint32_t)v * (uint32_t)recip + 32768u) >> 16;
        const int f = n + max(-1, min(1, v)) - max(-1, min(1, n));
        return c_quant[max(-8, min(8, f)) + 8];
    };
    if (threadIdx.x == 0) s_nb = 0;
    __syncthreads();
    for (int v = -lim + 1 + (int)threadIdx.x; v < lim; v += (int)blockDim.x) {
        const int q = quant(v);
        if (q != quant(v - 1)) {
            const int k = atomicAdd(&s_nb, 1);
            if (k < 8) {
                s_v[k] = v;
                s_q[k] = q;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    FastTab t;
    bool ok = s_nb == 7;
    if (ok) {
        for (int i = 1; i < 7; ++i) // sort the seven thresholds ascending
            for (int j = i; j > 0 && s_v[j] < s_v[j - 1]; --j) {
                const int tv = s_v[j], tq = s_q[j];
                s_v[j] = s_v[j - 1];
                s_q[j] = s_q[j - 1];
                s_v[j - 1] = tv;
                s_q[j - 1] = tq;
            }
        const int q0 = quant(-lim);
        t.x[0] = c_dequant[sf * 8 + q0] * 8 + q0;
        for (int k = 0; k < 7; ++k) {
            t.b[k] = s_v[k];
            t.x[k + 1] = c_dequant[sf * 8 + s_q[k]] * 8 + s_q[k];
        }
        for (int k = 0; k < 8; ++k) ok = ok && (t.x[k] & 7) == (k < 4 ? 7 - 2 * k : 2 * (k - 4)); // codes 7,5,3,1,0,2,4,6
    }
    t.lim = lim;
    if (ok) g_fast_tab[sf] = t;
    if (!ok) atomicExch(&g_fast_ok, 0);
}

__device__ __forceinline__ int keep(int x)
{
    asm("mov.b32 %0, %0;" : "+r"(x));
    return x;
}

struct FCand {
    int a, b, c, d;         // history, d newest
    int nw0, nw1, nw2, nw3; // negated weights as of one step ago (the update by dp is pending)
    int nsx, nsa, nsb, nsc; // negated +/-1 signs of the history of one step ago: the pending update's signs
    int ga, gb, gc;         // nsx * a, nsa * b, nsb * c
    int dp;                 // pending delta (dequantized >> 4 of the previous step)
};

__device__ __forceinline__ void fast_step(FCand &l, const FastTab &t, int s, long long &err, uint32_t &pen, uint32_t &codes,
                                          uint32_t &trip)
{
    // keep() pins the association: the compiler would otherwise re-chain the whole sum behind the newest sample
    const int K = s * 8192 + 8191;
    const int D = keep(l.nw0 * l.a + l.nw1 * l.b + l.nw2 * l.c + K);
    const int NB = keep(l.ga + l.gb + l.gc);
    const int E = keep(l.dp * NB + D);
    const int nw3 = keep(l.nsc * l.dp + l.nw3);
    const int v = (nw3 * l.d + E) >> 13; // == s - predict
    l.nw0 += l.nsx * l.dp;               // the pending update :821-828
    l.nw1 += l.nsa * l.dp;
    l.nw2 += l.nsb * l.dp;
    l.nw3 = nw3;
    const int ws = (l.nw0 * l.nw0 + l.nw1 * l.nw1 + l.nw2 * l.nw2 + l.nw3 * l.nw3) >> 18; // :814-817
    const int penalty = __viaddmax_s32(ws, -0x8ff, 0);
    const int x01 = v >= t.b[0] ? t.x[1] : t.x[0];
    const int x23 = v >= t.b[2] ? t.x[3] : t.x[2];
    const int x45 = v >= t.b[4] ? t.x[5] : t.x[4];
    const int x67 = v >= t.b[6] ? t.x[7] : t.x[6];
    const int xl = v >= t.b[1] ? x23 : x01;
    const int xh = v >= t.b[5] ? x67 : x45;
    const int x = v >= t.b[3] ? xh : xl;
    trip |= (uint32_t)(v + t.lim) >= 2u * (uint32_t)t.lim;
    const int dq = x >> 3;
    const int p = s - v;
    const int r = max(__viaddmin_s32(p, dq, 32767), -32768); // :573
    const int e = s - r;                                     // :575
    err += (long long)e * (long long)e;                      // :576
    pen += (uint32_t)penalty * (uint32_t)penalty;
    codes = codes * 8u + (uint32_t)(x & 7); // :584
    const int nsd = (l.d >> 31) * -2 - 1; // -sign: one shift + one IMAD
    l.ga = l.gb;
    l.gb = l.gc;
    l.gc = l.nsc * l.d;
    l.nsx = l.nsa; l.nsa = l.nsb; l.nsb = l.nsc; l.nsc = nsd;
    l.a = l.b; l.b = l.c; l.c = l.d; l.d = r;
    l.dp = x >> 7; // dequantized >> 4
}

// Lexicographic (total, first, sf) arg-min over a half-warp; only used when two candidates tie on the total rank.
__device__ __noinline__ uint32_t winner_slow(uint64_t total, uint64_t first, uint32_t sf)
{
#pragma unroll
    for (int m = 8; m >= 1; m >>= 1) {
        const uint64_t ot = __shfl_xor_sync(kFull, total, m);
        const uint64_t of = __shfl_xor_sync(kFull, first, m);
        const uint32_t os = __shfl_xor_sync(kFull, sf, m);
        const bool less = ot != total ? ot < total : (of != first ? of < first : os < sf);
        if (less) {
            total = ot;
            first = of;
            sf = os;
        }
    }
    return sf;
}

// One warp = two chains (half-warps).  Both halves run the same number of iterations (the longer chain's slice count);
// a half whose chain is exhausted computes on zeros and commits nothing, so every shuffle is warp-converged.
template <bool FAST>
__global__ void __launch_bounds__(kEncThreads, FAST ? 1 : 7)
encode_kernel(const int16_t *__restrict__ pcm, uint8_t *__restrict__ out, const EncJob *__restrict__ jobs,
              const EncChain *__restrict__ chains, uint32_t n_chains, int32_t *__restrict__ lms_state, int zero)
{
    __shared__ EncSmem sm;
    for (int i = threadIdx.x; i < 17 * 16; i += kEncThreads) {
        const int idx = i >> 4, sf = i & 15;
        const int q = c_quant[idx];
        const int dq = c_dequant[sf * 8 + q];
        sm.lut[i] = make_int4(dq, q, dq >> 4, 0);
    }
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31;
    const uint32_t sf = lane & 15;
    const uint32_t half = lane >> 4;
    const uint32_t warp_in_cta = threadIdx.x >> 5;
    const uint32_t warp = (blockIdx.x * kEncThreads + threadIdx.x) >> 5;
    const uint32_t chain = warp * 2 + half;
    const bool have_chain = chain < n_chains;
    const uint32_t lut_lane = smem_u32(&sm.lut[0]) + sf * 16 + 8 * 256;
    const uint32_t stage = smem_u32(&sm.stage[warp_in_cta][half][0]);
    const int recip = c_recip[sf];
    FastTab ft;
    if (FAST) ft = g_fast_tab[sf];

    EncJob job;
    job.pcm_off = 0; job.out_off = 0; job.channels = 1; job.rate = 0; job.nsamples = 0; job.lms_base = 0;
    job.file_samples = 0; job.flags = 0;
    uint32_t c = 0;
    if (have_chain) {
        const EncChain ec = chains[chain];
        job = jobs[ec.job];
        c = ec.channel;
    }
    const uint32_t ch = job.channels;
    const uint32_t n = job.nsamples;
    const uint32_t my_slices = n / kSliceLen + (n % kSliceLen != 0); // no n + 19: samples may be anything up to 2^32 - 1
    const uint32_t total_slices = max(my_slices, __shfl_xor_sync(kFull, my_slices, 16));
    if (total_slices == 0) return; // both chains of this warp idle

    int32_t *st = lms_state + (size_t)(job.lms_base + c) * 8;
    Cand cur;
    if (my_slices) {
        cur.h0 = st[0]; cur.h1 = st[1]; cur.h2 = st[2]; cur.h3 = st[3];
        cur.w0 = st[4]; cur.w1 = st[5]; cur.w2 = st[6]; cur.w3 = st[7];
    } else {
        cur.h0 = cur.h1 = cur.h2 = cur.h3 = 0;
        cur.w0 = cur.w1 = cur.w2 = cur.w3 = 0;
    }
    cur.s0 = sign_pm1(cur.h0); cur.s1 = sign_pm1(cur.h1); cur.s2 = sign_pm1(cur.h2); cur.s3 = sign_pm1(cur.h3);

    const int16_t *src = pcm + job.pcm_off + c;
    uint8_t *frames_out = out + job.out_off;
    if (job.flags & 1u) { // file header "qoaf" + samples (write_header :404-408)
        if (my_slices && c == 0 && sf == 0) st_be64(frames_out, kMagic, job.file_samples);
        frames_out += 8;
    }
    const size_t full_frame_bytes = 8 + 16 * (size_t)ch + 8 * (size_t)kSlicesPerFrame * ch; // qoa_frame_size :602-604

#pragma unroll 1
    for (uint32_t j = 0; j < total_slices; ++j) {
        // ---- stage the next 16 slices of this chain, de-interleaved (:502-508), zero-filled past the end
        if ((j & (kChunkSlices - 1)) == 0) {
            __syncwarp();
            const uint32_t base = j * kSliceLen + sf;
            int v[kChunkSamples / 16];
#pragma unroll
            for (int m = 0; m < kChunkSamples / 16; ++m) {
                const uint32_t s = base + 16 * m;
                v[m] = s < n ? (int)__ldg(src + (size_t)s * ch) : 0;
            }
#pragma unroll
            for (int m = 0; m < kChunkSamples / 16; ++m)
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(stage + (sf + 16 * m) * 2), "h"((short)v[m]) : "memory");
            __syncwarp();
        }
        const bool active = j < my_slices;
        const uint32_t f = j >> 8, k = j & 255;
        const uint32_t s0 = j * kSliceLen; // first sample (per channel) of this slice
        const uint32_t len = active ? min((uint32_t)kSliceLen, n - s0) : 0u;
        uint8_t *frame = frames_out + (size_t)f * full_frame_bytes;

        if (active && k == 0 && sf == 0) { // frame start :448-466
            const uint32_t frame_len = min((uint32_t)kFrameLen, n - f * kFrameLen);
            const uint32_t slices = (frame_len + kSliceLen - 1) / kSliceLen;
            const uint32_t frame_size = (8 + 16 * ch + 8 * slices * ch) & 0xffffu; // `as u16`
            if (c == 0) st_be64(frame, (ch << 24) | job.rate, (frame_len << 16) | frame_size); // rate OR-ed unmasked (:449)
            uint8_t *lp = frame + 8 + 16 * c; // truncating `as u16` dump of the carried i32 state :459-460
            st_be64(lp, ((uint32_t)cur.h0 << 16) | ((uint32_t)cur.h1 & 0xffffu), ((uint32_t)cur.h2 << 16) | ((uint32_t)cur.h3 & 0xffffu));
            st_be64(lp + 8, ((uint32_t)cur.w0 << 16) | ((uint32_t)cur.w1 & 0xffffu), ((uint32_t)cur.w2 << 16) | ((uint32_t)cur.w3 & 0xffffu));
        }

        // ---- the candidates: every lane runs the slice with its own scalefactor
        Cand l = cur;
        long long err = 0;
        uint32_t pen = 0, first_err = 0, first_pen = 0;
        uint32_t codes_a = 0, codes_b = 0; // samples 0..9 and 10..19, 3 bits each
        const uint32_t sp = stage + (j & (kChunkSlices - 1)) * (kSliceLen * 2);
        bool done = false;
        if (FAST && __all_sync(kFull, len == kSliceLen || !active)) { // latency mode: threshold search, lazy update
            FCand fl;
            fl.a = cur.h0; fl.b = cur.h1; fl.c = cur.h2; fl.d = cur.h3;
            fl.nw0 = -cur.w0; fl.nw1 = -cur.w1; fl.nw2 = -cur.w2; fl.nw3 = -cur.w3;
            fl.nsx = 0; fl.nsa = -cur.s0; fl.nsb = -cur.s1; fl.nsc = -cur.s2;
            fl.ga = 0; fl.gb = fl.nsa * fl.b; fl.gc = fl.nsb * fl.c;
            fl.dp = 0;
            uint32_t trip = 0;
            fast_step(fl, ft, lds_s16(sp), err, pen, codes_a, trip);
            first_err = (uint32_t)err;
            first_pen = pen;
#pragma unroll
            for (int i = 1; i < 10; ++i) fast_step(fl, ft, lds_s16(sp + i * 2), err, pen, codes_a, trip);
#pragma unroll
            for (int i = 10; i < 20; ++i) fast_step(fl, ft, lds_s16(sp + i * 2), err, pen, codes_b, trip);
            if (!__any_sync(kFull, trip != 0)) {
                done = true;
                l.h0 = fl.a; l.h1 = fl.b; l.h2 = fl.c; l.h3 = fl.d;
                l.w0 = -(fl.nw0 + fl.nsx * fl.dp); // the last pending update
                l.w1 = -(fl.nw1 + fl.nsa * fl.dp);
                l.w2 = -(fl.nw2 + fl.nsb * fl.dp);
                l.w3 = -(fl.nw3 + fl.nsc * fl.dp);
                l.s0 = -fl.nsa; l.s1 = -fl.nsb; l.s2 = -fl.nsc; l.s3 = sign_pm1(fl.d);
            } else { // some candidate left the exact domain: redo the slice with the reference arithmetic
                err = 0;
                pen = 0;
                codes_a = codes_b = 0;
            }
        }
        if (done) {
        } else if (__all_sync(kFull, len == kSliceLen || !active)) { // steady state (an idle half computes on zeros)
            candidate_step(l, lds_s16(sp), recip, lut_lane, err, pen, codes_a, zero);
            first_err = (uint32_t)err; // error^2 of one sample fits 32 bits
            first_pen = pen;
#pragma unroll
            for (int i = 1; i < 10; ++i) candidate_step(l, lds_s16(sp + i * 2), recip, lut_lane, err, pen, codes_a, zero);
#pragma unroll
            for (int i = 10; i < 20; ++i) candidate_step(l, lds_s16(sp + i * 2), recip, lut_lane, err, pen, codes_b, zero);
        } else { // a partial last slice somewhere in the warp
#pragma unroll 1
            for (uint32_t i = 0; i < kSliceLen; ++i) {
                if (i < len) {
                    candidate_step(l, lds_s16(sp + i * 2), recip, lut_lane, err, pen, i < 10 ? codes_a : codes_b, zero);
                    if (i == 0) {
                        first_err = (uint32_t)err;
                        first_pen = pen;
                    }
                }
            }
            // left-justify: the reference shifts the whole word by (20 - len) * 3 (:485-487)
            if (len <= 10)
                codes_a <<= (10 - len) * 3;
            else
                codes_b <<= (20 - len) * 3;
        }

        // ---- winner within the half-warp: min total rank; ties (rare) resolved by (first-sample rank, sf)
        const uint64_t total = (uint64_t)err + pen;
        uint32_t hi = (uint32_t)(total >> 32), lo = (uint32_t)total;
        // per-half minimum.  Latency mode: one warp-wide REDUX per half (the other half contributes the neutral
        // element) instead of four dependent shuffle + min stages; in the throughput kernel the butterfly is faster.
        auto half_min = [&](uint32_t x) {
            if (FAST) {
                const uint32_t m0 = __reduce_min_sync(kFull, half ? 0xffffffffu : x);
                const uint32_t m1 = __reduce_min_sync(kFull, half ? x : 0xffffffffu);
                return half ? m1 : m0;
            }
#pragma unroll
            for (int m = 8; m >= 1; m >>= 1) x = min(x, __shfl_xor_sync(kFull, x, m));
            return x;
        };
        uint32_t mh = hi;
        if (__any_sync(kFull, hi != 0)) {
            mh = half_min(hi);
            if (hi != mh) lo = 0xffffffffu; // out of the race (a true 0xffffffff with hi == mh still ties correctly below)
        }
        const uint32_t ml = half_min(lo);
        const bool is_min = hi == mh && (uint32_t)total == ml;
        const uint32_t ballot = __ballot_sync(kFull, is_min);
        const uint32_t votes = (ballot >> (half * 16)) & 0xffffu;
        uint32_t win = __ffs(votes) - 1;
        const uint32_t b0 = ballot & 0xffffu, b1 = ballot >> 16; // every lane sees both halves: no second vote needed
        if ((b0 & (b0 - 1)) | (b1 & (b1 - 1))) // some half has a tie on the total rank (warp-uniform)
            win = winner_slow(total, (uint64_t)first_err + first_pen, sf);
        const int src_lane = (int)((lane & 16) | win);
        const int nh0 = __shfl_sync(kFull, l.h0, src_lane), nh1 = __shfl_sync(kFull, l.h1, src_lane);
        const int nh2 = __shfl_sync(kFull, l.h2, src_lane), nh3 = __shfl_sync(kFull, l.h3, src_lane);
        const int nw0 = __shfl_sync(kFull, l.w0, src_lane), nw1 = __shfl_sync(kFull, l.w1, src_lane);
        const int nw2 = __shfl_sync(kFull, l.w2, src_lane), nw3 = __shfl_sync(kFull, l.w3, src_lane);
        const int ns0 = __shfl_sync(kFull, l.s0, src_lane), ns1 = __shfl_sync(kFull, l.s1, src_lane);
        const int ns2 = __shfl_sync(kFull, l.s2, src_lane), ns3 = __shfl_sync(kFull, l.s3, src_lane);
        if (active) { // carry the winner's full i32 state (:481-482)
            cur.h0 = nh0; cur.h1 = nh1; cur.h2 = nh2; cur.h3 = nh3;
            cur.w0 = nw0; cur.w1 = nw1; cur.w2 = nw2; cur.w3 = nw3;
            cur.s0 = ns0; cur.s1 = ns1; cur.s2 = ns2; cur.s3 = ns3;
            if (sf == win) { // the winning lane packs and stores its slice :488-489: sf:4 | codes_a:30 | codes_b:30
                const uint32_t word_hi = (sf << 28) | (codes_a >> 2);
                const uint32_t word_lo = (codes_a << 30) | codes_b;
                st_be64(frame + 8 + 16 * (size_t)ch + ((size_t)k * ch + c) * 8, word_hi, word_lo);
            }
        }
    }

    if (my_slices && sf == 0) { // carry the full i32 state to the next call (:482)
        st[0] = cur.h0; st[1] = cur.h1; st[2] = cur.h2; st[3] = cur.h3;
        st[4] = cur.w0; st[5] = cur.w1; st[6] = cur.w2; st[7] = cur.w3;
    }
}

static cudaError_t ensure_fast_table()
{
    static bool ready[64] = {false}, usable[64] = {false};
    static std::mutex mu; // contexts on different host threads may get here at the same time
    std::lock_guard<std::mutex> lock(mu);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (!ready[dev]) {
        const int one = 1;
        int ok = 0;
        if ((e = cudaMemcpyToSymbol(g_fast_ok, &one, sizeof one)) != cudaSuccess) return e;
        fast_table_kernel<<<16, 256>>>();
        if ((e = cudaMemcpyFromSymbol(&ok, g_fast_ok, sizeof ok)) != cudaSuccess) return e;
        usable[dev] = ok != 0;
        ready[dev] = true;
    }
    return usable[dev] ? cudaSuccess : cudaErrorNotSupported;
}

cudaError_t launch_encode(const int16_t *d_pcm, uint8_t *d_out, const EncJob *d_jobs, const EncChain *d_chains,
                          uint32_t n_chains, int32_t *d_lms, cudaStream_t s, int *n_launches)
{
    if (n_chains == 0) return cudaSuccess;
    const uint32_t warps = (n_chains + 1) / 2;
    // Few chains: the step is a pure dependency chain (one warp per scheduler or less) -> the latency-mode kernel,
    // one warp per CTA so that the warps spread over the SMs.  QOA_B200_ENC_FAST = 0 / 1 forces the choice
    // (development aid).
    static const int forced = getenv("QOA_B200_ENC_FAST") ? atoi(getenv("QOA_B200_ENC_FAST")) : -1;
    bool fast = forced >= 0 ? forced != 0 : warps <= kFastMaxWarps;
    if (fast) {
        const cudaError_t e = ensure_fast_table();
        if (e == cudaErrorNotSupported)
            fast = false; // the table did not validate: keep the exact kernel
        else if (e != cudaSuccess)
            return e;
    }
    if (fast) {
        const uint32_t blocks = (warps + kEncWarps - 1) / kEncWarps;
        encode_kernel<true><<<blocks, kEncThreads, 0, s>>>(d_pcm, d_out, d_jobs, d_chains, n_chains, d_lms, 0);
    } else {
        const uint32_t blocks = (warps + kEncWarps - 1) / kEncWarps;
        encode_kernel<false><<<blocks, kEncThreads, 0, s>>>(d_pcm, d_out, d_jobs, d_chains, n_chains, d_lms, 0);
    }
    if (n_launches) *n_launches += 1;
    return cudaGetLastError();
}

} // namespace qoa
