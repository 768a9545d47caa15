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
#include "qoa_kernels.h"

namespace qoa {

constexpr int kEncThreads = 128; // 4 warps = 8 chains per CTA

// Candidate LUT in shared memory: entry idx = clamp(scaled, -8, 8) + 8 of scalefactor sf at byte (idx * 16 + sf) * 8,
// holding {dequantized, quantized}.  A half-warp (16 scalefactors) reads 16 consecutive 8-byte words: conflict-free.
struct __align__(16) EncLut {
    int2 v[17 * 16];
};

struct Lms {
    int h0, h1, h2, h3;
    int w0, w1, w2, w3;
};

// qoa_div :613-617 — wrapping multiply, round half up, then push away from zero.
__device__ __forceinline__ int qoa_div(int v, int recip)
{
    const int n = (int)((uint32_t)v * (uint32_t)recip + 32768u) >> 16;
    const int sv = max(-1, min(1, v)); // (v > 0) - (v < 0)
    const int sn = max(-1, min(1, n));
    return n + sv - sn;
}

// One (candidate, sample) step :564-585.  Returns the rank contribution err^2 + penalty^2 (< 2^33).
__device__ __forceinline__ uint64_t candidate_step(Lms &l, int s, int recip, uint32_t lut_lane, uint32_t &codes)
{
    // predict_and_penalty_sq :808-818 (all sums wrap in i32)
    const int p = (l.w0 * l.h0 + l.w1 * l.h1 + l.w2 * l.h2 + l.w3 * l.h3) >> 13;
    const int ws = (l.w0 * l.w0 + l.w1 * l.w1 + l.w2 * l.w2 + l.w3 * l.w3) >> 18;
    const int pen = max(ws - 0x8ff, 0);
    const int v = s - p;                                 // :567
    const int c = max(-8, min(8, qoa_div(v, recip)));    // :568-569
    int dq, q;
    asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(dq), "=r"(q) : "r"(lut_lane + (uint32_t)(c + 8) * 128u)); // :570-572
    const int r = clamp_i16(p + dq);                     // :573
    const int e = s - r;                                 // :575
    const uint64_t rank = (uint64_t)((int64_t)e * (int64_t)e) + (uint64_t)((uint32_t)pen * (uint32_t)pen); // :576
    const int d = dq >> 4;                               // update :821-828
    l.w0 += l.h0 < 0 ? -d : d;
    l.w1 += l.h1 < 0 ? -d : d;
    l.w2 += l.h2 < 0 ? -d : d;
    l.w3 += l.h3 < 0 ? -d : d;
    l.h0 = l.h1;
    l.h1 = l.h2;
    l.h2 = l.h3;
    l.h3 = r;
    codes = (codes << 3) | (uint32_t)q;                  // :584
    return rank;
}

struct Key {
    uint64_t total, first;
    uint32_t sf;
};
__device__ __forceinline__ bool key_less(const Key &a, const Key &b)
{
    if (a.total != b.total) return a.total < b.total;
    if (a.first != b.first) return a.first < b.first;
    return a.sf < b.sf;
}

// One warp = two chains (half-warps).  Both halves run the same number of iterations (the longer chain's slice count);
// a half whose chain is exhausted computes on zeros and stores nothing, so every shuffle is warp-converged.
__global__ void __launch_bounds__(kEncThreads)
encode_kernel(const int16_t *__restrict__ pcm, uint8_t *__restrict__ out, const EncJob *__restrict__ jobs,
              const EncChain *__restrict__ chains, uint32_t n_chains, int32_t *__restrict__ lms_state)
{
    __shared__ EncLut lut;
    for (int i = threadIdx.x; i < 17 * 16; i += kEncThreads) {
        const int idx = i >> 4, sf = i & 15;
        const int q = c_quant[idx];
        lut.v[i] = make_int2(c_dequant[sf * 8 + q], q);
    }
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31;
    const uint32_t sf = lane & 15;
    const uint32_t warp = (blockIdx.x * kEncThreads + threadIdx.x) >> 5;
    const uint32_t chain = warp * 2 + (lane >> 4);
    const bool have_chain = chain < n_chains;
    const uint32_t lut_lane = smem_u32(&lut.v[0]) + sf * 8;
    const int recip = c_recip[sf];

    EncJob job;
    job.pcm_off = 0; job.out_off = 0; job.channels = 1; job.rate = 0; job.nsamples = 0; job.lms_base = 0; job.file_samples = 0; job.flags = 0;
    uint32_t c = 0;
    if (have_chain) {
        const EncChain ec = chains[chain];
        job = jobs[ec.job];
        c = ec.channel;
    }
    const uint32_t ch = job.channels;
    const uint32_t n = job.nsamples;
    const uint32_t my_slices = (n + kSliceLen - 1) / kSliceLen;
    const uint32_t other_slices = __shfl_xor_sync(0xffffffffu, my_slices, 16);
    const uint32_t total_slices = max(my_slices, other_slices);
    if (total_slices == 0) return; // both chains of this warp idle

    int32_t *st = lms_state + (size_t)(job.lms_base + c) * 8;
    Lms cur;
    if (my_slices) {
        cur.h0 = st[0]; cur.h1 = st[1]; cur.h2 = st[2]; cur.h3 = st[3];
        cur.w0 = st[4]; cur.w1 = st[5]; cur.w2 = st[6]; cur.w3 = st[7];
    } else {
        cur.h0 = cur.h1 = cur.h2 = cur.h3 = 0;
        cur.w0 = cur.w1 = cur.w2 = cur.w3 = 0;
    }

    const int16_t *src = pcm + job.pcm_off + c;
    uint8_t *frames_out = out + job.out_off;
    if (job.flags & 1u) { // file header "qoaf" + samples (write_header :404-408)
        if (my_slices && c == 0 && sf == 0) st_be64(frames_out, kMagic, job.file_samples);
        frames_out += 8;
    }
    const size_t full_frame_bytes = 8 + 16 * (size_t)ch + 8 * (size_t)kSlicesPerFrame * ch; // qoa_frame_size :602-604

    for (uint32_t j = 0; j < total_slices; ++j) {
        const bool active = j < my_slices;
        const uint32_t f = j >> 8, k = j & 255;
        const uint32_t s0 = j * kSliceLen;                       // first sample (per channel) of this slice
        const uint32_t len = active ? min((uint32_t)kSliceLen, n - s0) : 0u;
        uint8_t *frame = frames_out + (size_t)f * full_frame_bytes;

        if (active && k == 0 && sf == 0) { // frame start :448-466
            const uint32_t frame_len = min((uint32_t)kFrameLen, n - f * kFrameLen);
            const uint32_t slices = (frame_len + kSliceLen - 1) / kSliceLen;
            const uint32_t frame_size = (8 + 16 * ch + 8 * slices * ch) & 0xffffu; // `as u16`
            if (c == 0) st_be64(frame, (ch << 24) | job.rate, (frame_len << 16) | frame_size); // rate is OR-ed unmasked like :449
            uint8_t *lp = frame + 8 + 16 * c; // truncating `as u16` dump of the carried i32 state :459-460
            st_be64(lp, ((uint32_t)cur.h0 << 16) | ((uint32_t)cur.h1 & 0xffffu), ((uint32_t)cur.h2 << 16) | ((uint32_t)cur.h3 & 0xffffu));
            st_be64(lp + 8, ((uint32_t)cur.w0 << 16) | ((uint32_t)cur.w1 & 0xffffu), ((uint32_t)cur.w2 << 16) | ((uint32_t)cur.w3 & 0xffffu));
        }

        // the candidate: every lane runs all `len` samples with its own scalefactor
        Lms l = cur;
        uint64_t total = 0, first = 0;
        uint32_t codes_a = 0, codes_b = 0; // samples 0..9 and 10..19, 3 bits each
        if (len == kSliceLen) {
            const int16_t *sp = src + (size_t)s0 * ch;
#pragma unroll
            for (int i = 0; i < 10; ++i) {
                const uint64_t r = candidate_step(l, (int)__ldg(sp + (size_t)i * ch), recip, lut_lane, codes_a);
                if (i == 0) first = r;
                total += r;
            }
#pragma unroll
            for (int i = 10; i < 20; ++i) total += candidate_step(l, (int)__ldg(sp + (size_t)i * ch), recip, lut_lane, codes_b);
        } else {
            const int16_t *sp = src + (size_t)s0 * ch;
            for (uint32_t i = 0; i < len; ++i) {
                const uint64_t r = candidate_step(l, (int)__ldg(sp + (size_t)i * ch), recip, lut_lane, i < 10 ? codes_a : codes_b);
                if (i == 0) first = r;
                total += r;
            }
            // left-justify: the reference shifts the whole word by (20 - len) * 3 (:485-487)
            if (len <= 10) {
                codes_a <<= (10 - len) * 3;
            } else {
                codes_b <<= (20 - len) * 3;
            }
        }

        // winner within the half-warp: lexicographic min of (total, first, sf)
        Key best{total, first, sf};
#pragma unroll
        for (int m = 8; m >= 1; m >>= 1) {
            Key o;
            o.total = __shfl_xor_sync(0xffffffffu, best.total, m);
            o.first = __shfl_xor_sync(0xffffffffu, best.first, m);
            o.sf = __shfl_xor_sync(0xffffffffu, best.sf, m);
            if (key_less(o, best)) best = o;
        }
        const int src_lane = (int)((lane & 16) | best.sf);
        cur.h0 = __shfl_sync(0xffffffffu, l.h0, src_lane);
        cur.h1 = __shfl_sync(0xffffffffu, l.h1, src_lane);
        cur.h2 = __shfl_sync(0xffffffffu, l.h2, src_lane);
        cur.h3 = __shfl_sync(0xffffffffu, l.h3, src_lane);
        cur.w0 = __shfl_sync(0xffffffffu, l.w0, src_lane);
        cur.w1 = __shfl_sync(0xffffffffu, l.w1, src_lane);
        cur.w2 = __shfl_sync(0xffffffffu, l.w2, src_lane);
        cur.w3 = __shfl_sync(0xffffffffu, l.w3, src_lane);

        if (active && sf == best.sf) { // the winning lane packs and stores its slice :488-489
            // word = sf:4 | codes_a:30 | codes_b:30
            const uint64_t word = ((uint64_t)sf << 60) | ((uint64_t)codes_a << 30) | (uint64_t)codes_b;
            uint8_t *dst = frame + 8 + 16 * (size_t)ch + ((size_t)k * ch + c) * 8;
            st_be64(dst, (uint32_t)(word >> 32), (uint32_t)word);
        }
    }

    if (my_slices && sf == 0) { // carry the full i32 state to the next call (:482)
        st[0] = cur.h0; st[1] = cur.h1; st[2] = cur.h2; st[3] = cur.h3;
        st[4] = cur.w0; st[5] = cur.w1; st[6] = cur.w2; st[7] = cur.w3;
    }
}

cudaError_t launch_encode(const int16_t *d_pcm, uint8_t *d_out, const EncJob *d_jobs, const EncChain *d_chains,
                          uint32_t n_chains, int32_t *d_lms, cudaStream_t s, int *n_launches)
{
    if (n_chains == 0) return cudaSuccess;
    const uint32_t warps = (n_chains + 1) / 2;
    const uint32_t blocks = (warps * 32 + kEncThreads - 1) / kEncThreads;
    encode_kernel<<<blocks, kEncThreads, 0, s>>>(d_pcm, d_out, d_jobs, d_chains, n_chains, d_lms);
    if (n_launches) *n_launches += 1;
    return cudaGetLastError();
}

} // namespace qoa
