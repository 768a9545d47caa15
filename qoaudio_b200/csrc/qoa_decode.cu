// qoa_decode.cu — frame index pre-pass and frame-parallel decode kernels (sm_100a).
//
// Reference semantics (rafaelcaricio/qoaudio, src/lib.rs):
//   QoaDecoder::new                       :127-169   file header, mode, first frame header
//   decode_frame_header_and_lms           :205-289   header parse + validation + LMS load, EOF rules
//   decode_one_slice_per_channel          :291-330   the 20-sample LMS loop
//   Iterator::next                        :661-693   item order, streaming-mode dummy header
//   decode_all                            :713-747   extra IncompatibleFrame check
// Every frame header carries its own LMS state (:270-281), so frames decode independently: one thread owns one frame
// (mono / stereo fast path), one (frame, channel pair) or one (frame, channel) and walks its slices; nothing crosses a
// frame.
#include <cuda.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "qoa_kernels.h"

namespace qoa {

// ================================================================================================================
// Index pre-pass
// ================================================================================================================

// Big-endian loads usable from both sides: the walk below also runs on the host for qoa_b200_probe (header parsing only;
// no sample is ever decoded on the CPU).
__host__ __device__ __forceinline__ uint32_t ld_be32(const uint8_t *p) // 4-byte aligned on the device
{
#ifdef __CUDA_ARCH__
    return bswap32(__ldg(reinterpret_cast<const uint32_t *>(p)));
#else
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
#endif
}
__host__ __device__ __forceinline__ uint2 ld_be64(const uint8_t *p) // 8-byte aligned on the device; (hi, lo)
{
#ifdef __CUDA_ARCH__
    return ldg_be64(p);
#else
    return make_uint2(ld_be32(p), ld_be32(p + 4));
#endif
}

// Walks one stream exactly like QoaDecoder::new + Iterator::next would, without decoding samples.
template <bool FILL>
__host__ __device__ void walk_stream(const uint8_t *__restrict__ arena, uint64_t off, uint64_t len, int mode, uint32_t stream,
                            FrameRec *__restrict__ recs, qoa_b200_stream_info *__restrict__ out)
{
    const uint8_t *p = arena + off;
    qoa_b200_stream_info info;
    info.status = QOA_B200_OK;
    info.streaming = 0;
    info.header_samples = 0;
    info.channels = 0;
    info.sample_rate = 0;
    info.index_frames = 0;
    info.max_channels = 0;
    info.flags = 0;
    info.samples = 0;
    info.frames = 0;

    uint64_t pos = 0;
    bool streaming = true;
    bool ok = true;
    if (mode != QOA_B200_DECODE_FRAMES) { // QoaDecoder::new :128-144
        if (len < 4) {
            info.status = QOA_B200_DEC_IO_ERROR;
            ok = false;
        } else if (ld_be32(p) != kMagic) {
            info.status = QOA_B200_DEC_NOT_QOA_FILE;
            ok = false;
        } else if (len < 8) {
            info.status = QOA_B200_DEC_IO_ERROR;
            ok = false;
        } else {
            info.header_samples = ld_be32(p + 4);
            streaming = info.header_samples == 0;
            pos = 8;
        }
    }
    info.streaming = (ok && streaming) ? 1 : 0;

    uint64_t items = 0, table_frames = 0, samples = 0;
    uint32_t first_ch = 0, first_rate = 0, fixed_ch = 0, fixed_rate = 0, max_ch = 0, min_ch = 0xffffffffu;
    bool aligned = true;
    bool have_first = false, first = true;
    if (ok && streaming && mode != QOA_B200_DECODE_FRAMES) { // :674-677 default header replayed first
        items = 1;
        have_first = true;
    }
    while (ok) {
        if (len - pos < 8) { // :206-216 EOF at a header is a clean end ...
            if (!streaming && first) info.status = QOA_B200_DEC_NO_SAMPLES; // ... except in new() :163-165
            break;
        }
        const uint64_t frame_pos = pos;
        const uint2 hw = ld_be64(p + pos);
        pos += 8;
        const uint32_t ch = hw.x >> 24;            // :217
        const uint32_t rate = hw.x & 0xffffffu;     // :218
        const uint32_t nspc = hw.y >> 16;           // :219
        const uint32_t frame_size = hw.y & 0xffffu; // :220
        if (ch == 0 || rate == 0) { // :227
            info.status = QOA_B200_DEC_INVALID_FRAME_HEADER;
            break;
        }
        const uint32_t non_sample = 8 + 16 * ch; // :232
        if (frame_size <= non_sample) {           // :233
            info.status = QOA_B200_DEC_INVALID_FRAME_HEADER;
            break;
        }
        const uint32_t num_slices = (frame_size - non_sample) / 8;     // :236-237
        if (num_slices % ch != 0 || num_slices / ch > kSlicesPerFrame) { // :239, :242
            info.status = QOA_B200_DEC_INVALID_FRAME_HEADER;
            break;
        }
        if (!streaming) { // :246-259
            if (first) {
                fixed_ch = ch;
                fixed_rate = rate;
            } else if (ch != fixed_ch || rate != fixed_rate) {
                info.status = QOA_B200_DEC_INCOMPATIBLE_FRAME;
                break;
            }
        }
        if (len - pos < 16ull * ch) { // :271-281 read_exact of the LMS block fails
            info.status = QOA_B200_DEC_IO_ERROR;
            break;
        }
        pos += 16ull * ch;
        // the FrameHeader item is yielded now (:674-678 / :685)
        if (mode != QOA_B200_DECODE_FRAMES) items++;
        if (!have_first) {
            have_first = true;
            first_ch = ch;
            first_rate = rate;
        } else if (mode == QOA_B200_DECODE_ALL && (ch != first_ch || rate != first_rate)) { // decode_all :735-738
            info.status = QOA_B200_DEC_INCOMPATIBLE_FRAME;
            break;
        }
        if (mode == QOA_B200_DECODE_FRAMES) items++;
        // slices: the decoder is driven by nspc, not by frame_size (:679-688, :320-328)
        const uint64_t groups_needed = (nspc + kSliceLen - 1) / kSliceLen;
        const uint64_t group_bytes = 8ull * ch;
        const uint64_t groups_avail = (len - pos) / group_bytes;
        uint32_t nsamp = nspc;
        bool truncated = false;
        if (groups_avail < groups_needed) { // read_exact fails inside a slice group (:301): whole groups before it were yielded
            nsamp = (uint32_t)groups_avail * kSliceLen;
            truncated = true;
        }
        if (FILL) {
            FrameRec r;
            r.in_off = off + frame_pos;
            r.out_rel = samples;
            r.stream = stream;
            r.rate = rate;
            r.nsamp = (uint16_t)nsamp;
            r.nspc = (uint16_t)nspc;
            r.ch = (uint8_t)ch;
            r.pad[0] = r.pad[1] = r.pad[2] = 0;
            recs[table_frames] = r;
        }
        table_frames++;
        max_ch = ch > max_ch ? ch : max_ch;
        min_ch = ch < min_ch ? ch : min_ch;
        aligned = aligned && ((samples & 7) == 0);
        samples += (uint64_t)nsamp * ch;
        if (truncated) {
            info.status = QOA_B200_DEC_IO_ERROR;
            break;
        }
        pos += groups_needed * group_bytes;
        first = false;
    }
    if (!FILL) {
        info.channels = first_ch;
        info.sample_rate = first_rate;
        info.samples = samples;
        info.frames = items;
        info.index_frames = (uint32_t)table_frames;
        info.max_channels = max_ch;
        info.flags = ((table_frames == 0 || max_ch == min_ch) ? QOA_B200_INFO_UNIFORM_CH : 0u) |
                     (aligned ? QOA_B200_INFO_ALIGNED16 : 0u);
        *out = info;
    }
}

// ---- closed-form index for conforming streams ------------------------------------------------------------------
// A stream written by a conforming encoder (the crate's included) has every frame but the last at 5120 samples, so the
// frame offsets, sizes and header words follow from the file header and the first frame header alone.  The pre-pass
// SPECULATES on that: the probe accepts a stream as "regular" when its length equals the closed-form size and its
// first frame header is the expected word; the fill kernel then checks EVERY frame header against its expected word in
// parallel (one thread per frame).  Any mismatch raises *bad and the host falls back to the serial walk for the whole
// batch, so the result is always what walking would have produced — it just skips 517 dependent loads per stream.
__device__ __forceinline__ uint2 expected_header(uint32_t ch, uint32_t rate, uint32_t frame_len)
{
    const uint32_t slices = (frame_len + kSliceLen - 1) / kSliceLen;
    return make_uint2((ch << 24) | rate, (frame_len << 16) | ((8 + 16 * ch + 8 * slices * ch) & 0xffffu));
}

__device__ bool probe_regular(const uint8_t *p, uint64_t len, qoa_b200_stream_info *out)
{
    if (len < 8 + 8 + 16 + 8) return false;
    if (ld_be32(p) != kMagic) return false;
    const uint32_t n = ld_be32(p + 4);
    if (n == 0) return false; // streaming mode: walk
    const uint2 h = ldg_be64(p + 8);
    const uint32_t ch = h.x >> 24, rate = h.x & 0xffffffu;
    if (ch == 0 || ch > 8 || rate == 0) return false;
    const uint64_t frames = ((uint64_t)n + kFrameLen - 1) / kFrameLen;
    const uint64_t slices = ((uint64_t)n + kSliceLen - 1) / kSliceLen;
    if (len != 8 + frames * (8 + 16ull * ch) + slices * 8ull * ch) return false;
    const uint2 e = expected_header(ch, rate, (uint32_t)min((uint64_t)kFrameLen, (uint64_t)n));
    if (h.x != e.x || h.y != e.y) return false;
    qoa_b200_stream_info info;
    info.status = QOA_B200_OK;
    info.streaming = 0;
    info.header_samples = n;
    info.channels = ch;
    info.sample_rate = rate;
    info.index_frames = (uint32_t)frames;
    info.samples = (uint64_t)n * ch;
    info.frames = frames;
    info.max_channels = ch;
    // frame f starts at f * 5120 * ch PCM elements: always a multiple of 8 (16 bytes)
    info.flags = QOA_B200_INFO_UNIFORM_CH | QOA_B200_INFO_REGULAR | QOA_B200_INFO_ALIGNED16;
    *out = info;
    return true;
}

// `host_info` is a host-mapped (pinned) mirror of the table: the host needs the per-stream results to size the frame
// index, and reading them through mapped memory keeps the copy engines — busy with payload copies in the pipelined
// host path — out of the dependency chain.
__global__ void index_count_kernel(const uint8_t *__restrict__ arena, const uint64_t *__restrict__ in_off,
                                   const uint64_t *__restrict__ in_len, uint32_t n, int mode, int speculate,
                                   qoa_b200_stream_info *__restrict__ info, qoa_b200_stream_info *__restrict__ host_info)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!(speculate && mode != QOA_B200_DECODE_FRAMES && probe_regular(arena + in_off[i], in_len[i], info + i)))
        walk_stream<false>(arena, in_off[i], in_len[i], mode, i, nullptr, info + i);
    host_info[i] = info[i];
}

__global__ void index_fill_kernel(const uint8_t *__restrict__ arena, const uint64_t *__restrict__ in_off,
                                  const uint64_t *__restrict__ in_len, uint32_t n, int mode,
                                  const qoa_b200_stream_info *__restrict__ info,
                                  const uint64_t *__restrict__ frame_base, FrameRec *__restrict__ recs)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (info[i].flags & QOA_B200_INFO_REGULAR)) return;
    walk_stream<true>(arena, in_off[i], in_len[i], mode, i, recs + frame_base[i], nullptr);
}

// One thread per (stream, frame) of the regular streams: verify the header word, write the FrameRec.
__global__ void index_fill_regular_kernel(const uint8_t *__restrict__ arena, const uint64_t *__restrict__ in_off,
                                          uint32_t n, const qoa_b200_stream_info *__restrict__ info,
                                          const uint64_t *__restrict__ frame_base, FrameRec *__restrict__ recs,
                                          int *__restrict__ bad)
{
    const uint32_t i = blockIdx.x;
    const qoa_b200_stream_info inf = info[i];
    if (!(inf.flags & QOA_B200_INFO_REGULAR)) return;
    const uint32_t ch = inf.channels, rate = inf.sample_rate, total = inf.header_samples;
    const uint64_t full = 8 + 16ull * ch + 8ull * kSlicesPerFrame * ch;
    for (uint32_t f = blockIdx.y * blockDim.x + threadIdx.x; f < inf.index_frames; f += gridDim.y * blockDim.x) {
        const uint32_t frame_len = min((uint32_t)kFrameLen, total - f * kFrameLen);
        const uint64_t pos = 8 + (uint64_t)f * full;
        const uint2 h = ldg_be64(arena + in_off[i] + pos);
        const uint2 e = expected_header(ch, rate, frame_len);
        if (h.x != e.x || h.y != e.y) *reinterpret_cast<volatile int *>(bad) = 1; // host-mapped flag; every writer stores 1
        FrameRec r;
        r.in_off = in_off[i] + pos;
        r.out_rel = (uint64_t)f * kFrameLen * ch;
        r.stream = i;
        r.rate = rate;
        r.nsamp = (uint16_t)frame_len;
        r.nspc = (uint16_t)frame_len;
        r.ch = (uint8_t)ch;
        r.pad[0] = r.pad[1] = r.pad[2] = 0;
        recs[frame_base[i] + f] = r;
    }
}

cudaError_t launch_index_count(const uint8_t *d_arena, const uint64_t *d_in_off, const uint64_t *d_in_len, size_t n,
                               int mode, int speculate, qoa_b200_stream_info *d_info, qoa_b200_stream_info *h_info,
                               cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    const int block = 32; // one dependent-load chain per thread: spread the streams over as many SMs as possible
    index_count_kernel<<<(unsigned)((n + block - 1) / block), block, 0, s>>>(d_arena, d_in_off, d_in_len, (uint32_t)n,
                                                                             mode, speculate, d_info, h_info);
    return cudaGetLastError();
}

cudaError_t launch_index_fill(const uint8_t *d_arena, const uint64_t *d_in_off, const uint64_t *d_in_len, size_t n,
                              int mode, const qoa_b200_stream_info *d_info, const uint64_t *d_frame_base,
                              FrameRec *d_recs, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    const int block = 32;
    index_fill_kernel<<<(unsigned)((n + block - 1) / block), block, 0, s>>>(d_arena, d_in_off, d_in_len, (uint32_t)n,
                                                                            mode, d_info, d_frame_base, d_recs);
    return cudaGetLastError();
}

cudaError_t launch_index_fill_regular(const uint8_t *d_arena, const uint64_t *d_in_off, size_t n, uint32_t max_frames,
                                      const qoa_b200_stream_info *d_info, const uint64_t *d_frame_base,
                                      FrameRec *d_recs, int *d_bad, cudaStream_t s)
{
    if (n == 0 || max_frames == 0) return cudaSuccess;
    const int block = 128;
    const unsigned gy = (unsigned)std::min<uint64_t>((max_frames + block - 1) / block, 65535);
    for (size_t first = 0; first < n; first += 0x7fffffffu) { // grid.x is 31-bit
        const unsigned gx = (unsigned)std::min<size_t>(n - first, 0x7fffffffu);
        index_fill_regular_kernel<<<dim3(gx, gy), block, 0, s>>>(d_arena, d_in_off + first, gx, d_info + first,
                                                                d_frame_base + first, d_recs, d_bad);
    }
    return cudaGetLastError();
}

// cudaFuncSetAttribute is per device: remember which devices have been configured for each kernel instantiation.
struct PerDeviceOnce {
    std::atomic<bool> done[64];
    PerDeviceOnce() { for (auto &d : done) d.store(false); }
    // returns the slot of the current device, or nullptr when it has to be configured on every launch
    std::atomic<bool> *slot()
    {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
        return &done[dev];
    }
};

// ================================================================================================================
// Decode
// ================================================================================================================

// Dequant LUT in shared memory: entry (sf*8 + code) holds {dequantized, dequantized >> 4} (the residual and the LMS
// delta of update(), src/lib.rs:822) and is replicated 16 times so that the 16 lanes of a half-warp (one wavefront of
// a 64-bit shared load) read 16 distinct bank pairs: copy (L & 15) of entry e lives at byte e*128 + (L & 15)*8
// (16 KiB, conflict-free).
// (Round 1 also had a 64 KiB table indexed by two codes at once - one LDS.128 per two samples, 8.5 % fewer instructions,
// but two-way bank conflicts on almost every load: 499 M conflicts per 1280 streams, L1TEX 90 %, +2 % net.  With the
// uniform-warp loop of round 2 this conflict-free table is as fast at 98 M conflicts and a quarter of the shared
// memory, so the two-code path is gone; an 8-copy conflict-free version of it only fits next to 80-byte staging rows
// and is then bound by the TMA engine's row rate, 36.0 ms against 27.2 ms: profiles/r02_decode_experiments.md.)
constexpr int kLutBytes = 128 * 16 * 8;
constexpr uint32_t kLutAlign = 1024; // bits 7..9 of the base must be zero: code offsets are OR-ed into the address

__device__ __forceinline__ void fill_lut(uint8_t *lut)
{
    int2 *v = reinterpret_cast<int2 *>(lut);
    for (int i = threadIdx.x; i < 128 * 16; i += blockDim.x) {
        const int dq = c_dequant[i >> 4];
        v[i] = make_int2(dq, dq >> 4);
    }
    __syncthreads();
}

// LMS state of one channel inside a thread.  s[i] = +1 / -1 is the sign the reference's update() derives from
// history[i] (src/lib.rs:823-826: history < 0 ? -delta : delta, zero counts as positive).
struct Chain {
    int h0, h1, h2, h3;
    int w0, w1, w2, w3;
    int s0, s1, s2, s3;
};

__device__ __forceinline__ int sign_pm1(int v) { return (v >> 31) | 1; }

// 16 bytes: 4 x BE i16 history, 4 x BE i16 weights, sign-extended (:271-281)
__device__ __forceinline__ void load_lms(Chain &c, const uint8_t *p)
{
    const uint2 h = ldg_be64(p);
    const uint2 w = ldg_be64(p + 8);
    c.h0 = (int)h.x >> 16;
    c.h1 = (int)(h.x << 16) >> 16;
    c.h2 = (int)h.y >> 16;
    c.h3 = (int)(h.y << 16) >> 16;
    c.w0 = (int)w.x >> 16;
    c.w1 = (int)(w.x << 16) >> 16;
    c.w2 = (int)w.y >> 16;
    c.w3 = (int)(w.y << 16) >> 16;
    c.s0 = sign_pm1(c.h0);
    c.s1 = sign_pm1(c.h1);
    c.s2 = sign_pm1(c.h2);
    c.s3 = sign_pm1(c.h3);
}

// Byte offset (code * 128) of sample I's 3-bit code inside the replicated LUT.
// Slice word layout (:303, :308): sf in bits 63..60, sample i's code in bits 59-3i .. 57-3i.
template <int I>
__device__ __forceinline__ uint32_t code_off(uint32_t hi, uint32_t lo)
{
    constexpr int pos = 57 - 3 * I; // LSB of the code within the 64-bit word
    constexpr int sh = pos - 7;     // shift that brings the code to bits 9..7
    uint32_t v;
    if constexpr (sh >= 32)
        v = hi >> (sh - 32);
    else if constexpr (sh > 0)
        v = __funnelshift_r(lo, hi, sh);
    else if constexpr (sh == 0)
        v = lo;
    else
        v = lo << (-sh);
    return v & 0x380u;
}

__device__ __forceinline__ int2 lds_v2(uint32_t addr)
{
    int2 v;
    asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}

// One sample of one channel (:306-317).  `row` = shared address of the lane's LUT entries for this slice's
// scalefactor (1 KiB-aligned apart from the lane offset, so OR-ing the code offset in is an add).
__device__ __forceinline__ int step_core(Chain &c, int2 t);

template <int I>
__device__ __forceinline__ int decode_step(Chain &c, uint32_t row, uint32_t hi, uint32_t lo)
{
    return step_core(c, lds_v2(row | code_off<I>(hi, lo))); // {dequantized, delta}
}

__device__ __forceinline__ int step_core(Chain &c, int2 t)
{
    const int p = (c.w0 * c.h0 + c.w1 * c.h1 + c.w2 * c.h2 + c.w3 * c.h3) >> 13; // predict :799-805
    const int r = clamp_i16(p + t.x);                                             // :310
    c.w0 += c.s0 * t.y;                                                           // update :821-828
    c.w1 += c.s1 * t.y;
    c.w2 += c.s2 * t.y;
    c.w3 += c.s3 * t.y;
    c.h0 = c.h1;
    c.h1 = c.h2;
    c.h2 = c.h3;
    c.h3 = r;
    c.s0 = c.s1;
    c.s1 = c.s2;
    c.s2 = c.s3;
    c.s3 = sign_pm1(r);
    return r;
}

// Both channels of one sample.  The two reconstructed values are clamped to i16 AND packed into the (L, R) output
// word by one I2IP (cvt.pack.sat.s16.s32); the clamped values are then read back out of the packed word for the LMS
// history.  Same arithmetic as decode_step, two instructions fewer per sample pair.
__device__ __forceinline__ uint32_t pair_core(Chain &a, Chain &b, int2 ta, int2 tb)
{
    const int xa = ((a.w0 * a.h0 + a.w1 * a.h1 + a.w2 * a.h2 + a.w3 * a.h3) >> 13) + ta.x; // predict :799-805, :310
    const int xb = ((b.w0 * b.h0 + b.w1 * b.h1 + b.w2 * b.h2 + b.w3 * b.h3) >> 13) + tb.x;
    uint32_t packed; // (clamp_i16(xb) << 16) | (clamp_i16(xa) & 0xffff)
    asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(packed) : "r"(xb), "r"(xa));
    int ra; // sign-extend the low half: selector nibble 9 = replicate the sign of byte 1
    asm("prmt.b32 %0, %1, %1, 0x9910;" : "=r"(ra) : "r"(packed));
    const int rb = (int)packed >> 16;
    a.w0 += a.s0 * ta.y; // update :821-828
    a.w1 += a.s1 * ta.y;
    a.w2 += a.s2 * ta.y;
    a.w3 += a.s3 * ta.y;
    b.w0 += b.s0 * tb.y;
    b.w1 += b.s1 * tb.y;
    b.w2 += b.s2 * tb.y;
    b.w3 += b.s3 * tb.y;
    a.h0 = a.h1; a.h1 = a.h2; a.h2 = a.h3; a.h3 = ra;
    b.h0 = b.h1; b.h1 = b.h2; b.h2 = b.h3; b.h3 = rb;
    a.s0 = a.s1; a.s1 = a.s2; a.s2 = a.s3; a.s3 = sign_pm1(ra);
    b.s0 = b.s1; b.s1 = b.s2; b.s2 = b.s3; b.s3 = sign_pm1(rb);
    return packed;
}

template <int I>
__device__ __forceinline__ uint32_t decode_pair_step(Chain &a, Chain &b, uint32_t row_a, uint32_t row_b, uint2 wa, uint2 wb)
{
    const int2 ta = lds_v2(row_a | code_off<I>(wa.x, wa.y)); // {dequantized, delta}
    const int2 tb = lds_v2(row_b | code_off<I>(wb.x, wb.y));
    return pair_core(a, b, ta, tb);
}

template <int I>
struct PairLoop {
    // Decodes samples I..19 of a slice for two channels; o[i] = (r_b << 16) | (r_a & 0xffff).
    static __device__ __forceinline__ void run(Chain &a, Chain &b, uint32_t row_a, uint32_t row_b, uint2 wa, uint2 wb,
                                               uint32_t *o)
    {
        o[I] = decode_pair_step<I>(a, b, row_a, row_b, wa, wb);
        PairLoop<I + 1>::run(a, b, row_a, row_b, wa, wb, o);
    }
};
template <>
struct PairLoop<20> {
    static __device__ __forceinline__ void run(Chain &, Chain &, uint32_t, uint32_t, uint2, uint2, uint32_t *) {}
};

// Two consecutive samples of ONE channel -> the output word (r1 << 16) | (r0 & 0xffff).  The second sample is clamped
// by the same I2IP that builds the word (its other operand, r0, is already in range), so the pair costs one
// instruction less than two clamps plus a pack.
__device__ __forceinline__ uint32_t two_core(Chain &c, int2 t0, int2 t);

template <int I>
__device__ __forceinline__ uint32_t decode_two_steps(Chain &c, uint32_t row, uint32_t hi, uint32_t lo)
{
    return two_core(c, lds_v2(row | code_off<I>(hi, lo)), lds_v2(row | code_off<I + 1>(hi, lo)));
}

__device__ __forceinline__ uint32_t two_core(Chain &c, int2 t0, int2 t)
{
    const int r0 = step_core(c, t0);
    const int x1 = ((c.w0 * c.h0 + c.w1 * c.h1 + c.w2 * c.h2 + c.w3 * c.h3) >> 13) + t.x;
    uint32_t packed;
    asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(packed) : "r"(x1), "r"(r0));
    const int r1 = (int)packed >> 16;
    c.w0 += c.s0 * t.y;
    c.w1 += c.s1 * t.y;
    c.w2 += c.s2 * t.y;
    c.w3 += c.s3 * t.y;
    c.h0 = c.h1; c.h1 = c.h2; c.h2 = c.h3; c.h3 = r1;
    c.s0 = c.s1; c.s1 = c.s2; c.s2 = c.s3; c.s3 = sign_pm1(r1);
    return packed;
}

template <int I>
struct SingleLoop {
    // Decodes samples I..19 of a slice for one channel, two per step; o[i / 2] = samples (i, i + 1) as one word.
    static __device__ __forceinline__ void run(Chain &a, uint32_t row_a, uint2 wa, uint32_t *o)
    {
        o[I / 2] = decode_two_steps<I>(a, row_a, wa.x, wa.y);
        SingleLoop<I + 2>::run(a, row_a, wa, o);
    }
};
template <>
struct SingleLoop<20> {
    static __device__ __forceinline__ void run(Chain &, uint32_t, uint2, uint32_t *) {}
};

// ---- fast path for uniform mono / stereo batches ---------------------------------------------------------------
// One thread = one frame, all of its channels: for stereo two independent LMS chains (ILP 2; an (L, R) pair packs into
// one 32-bit word), for mono one chain (two consecutive samples per word).
// Requirements (checked by the host from the index flags): every frame has CH channels and starts 16-byte aligned in
// the PCM output.
//
// Memory path.  Lanes of a warp own different frames (4136 B apart in the input, 20480 B apart in the output for
// stereo), so direct per-lane stores would cost one L1 wavefront per lane per instruction.  Instead every lane stages
// S slices (240 B of PCM: S = 3 stereo, S = 6 mono) in its own row of a per-warp shared-memory tile [32 lanes][240 B]
// and the tile goes to global memory through the TMA engine:
//   * regular boxes (8 consecutive full frames of one stream, i.e. rows exactly one frame of PCM apart in the
//     output): one 2-D tensor store per box per group (cp.async.bulk.tensor.2d, box = 60 words x 8 rows of a tensor
//     map that views the output arena as rows of one frame of PCM, dim 0 extended by one box so that a box may start
//     anywhere in a row) — no copy-out instructions at all; the last, shorter group of a frame (256 mod S slices) is
//     staged densely and leaves through a second map with a narrower box;
//   * everything else (the box that contains a stream boundary, short last frames, partial slices; 1-2 % of the
//     rows): the warp copies those rows with 16-byte shared loads / global stores, one half-warp per 240-byte row.
//     (Measured alternatives: per-lane cp.async.bulk copies — UBLKCP takes uniform operands, so the compiler
//     serialises it over the 32 lanes, ~480 instructions per group; per-lane 16-byte stores — 32 scattered sectors
//     per instruction, groups on that path ran 1.8x slower.)
// The row pitch (240 bytes) is an odd multiple of 16: the 128-bit staging stores of a quarter-warp hit eight distinct
// bank groups.  NB tiles per warp form a ring.
// Slice words arrive through a per-lane ring of four 16-byte vectors filled by cp.async two to three slices ahead
// of use (frames are only 8-byte aligned: a stereo slice pair is either one aligned vector or the upper half of one
// and the lower half of the next — reading the ring with 8-byte shared loads needs no alignment selects, and unlike a
// register prefetch the asynchronous copy never shares a scoreboard slot with the LUT loads).
template <int CH, int S, int NB>
struct FastCfg {
    static_assert(CH == 1 || CH == 2, "fast path: mono or stereo");
    static constexpr uint32_t kOutRowBytes = kFrameLen * CH * 2;  // PCM bytes of one full frame = tensor-map row pitch
    static constexpr int kSliceBytes = kSliceLen * CH * 2;        // PCM bytes one lane produces per slice
    static constexpr int kRowBytes = S * kSliceBytes;             // ... per group
    static constexpr int kGroupSamples = S * kSliceLen;           // per channel
    static constexpr int kTileBytes = 32 * kRowBytes;             // one warp's tile
    static constexpr int kWarpBytes = NB * kTileBytes;
    static constexpr int kBoxRows = 8;                            // rows (frames) per tensor store
    static constexpr int kInRing = 4;                             // 16-byte input vectors in flight per lane
    static constexpr int kInWarpBytes = kInRing * 32 * 16;        // slot-major: slot s of lane l at s*512 + l*16
    static_assert((kRowBytes / 16) % 2 == 1 && kRowBytes % 16 == 0, "row pitch must be an odd multiple of 16 bytes");
    static_assert((kBoxRows * kRowBytes) % 128 == 0 && (kBoxRows * (kSlicesPerFrame % S) * kSliceBytes) % 128 == 0,
                  "every box handed to the TMA engine must start 128-byte aligned in shared memory");
};

__device__ __forceinline__ void tensor_store_2d(const CUtensorMap *map, uint32_t saddr, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0),
                 "r"(c1), "r"(saddr)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void *gptr) // L2 only: the ring is the cache
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint2 lds64(uint32_t saddr)
{
    uint2 v;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t saddr, uint32_t a, uint32_t b)
{
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(saddr), "r"(a), "r"(b) : "memory");
}

template <int CH, int S, int NB, int THREADS>
__global__ void __launch_bounds__(THREADS, (THREADS <= 128 ? 5 : THREADS <= 192 ? 3 : THREADS <= 352 ? 2 : 1))
decode_fast_kernel(const uint8_t *__restrict__ arena, const FrameRec *__restrict__ recs, uint64_t n_frames,
                   const uint64_t *__restrict__ out_base, int16_t *__restrict__ out,
                   const __grid_constant__ CUtensorMap out_map, const __grid_constant__ CUtensorMap out_map_rem,
                   const __grid_constant__ CUtensorMap out_map32, const __grid_constant__ CUtensorMap out_map32_rem,
                   int use_map)
{
    using Cfg = FastCfg<CH, S, NB>;
    constexpr uint32_t kRem = kSlicesPerFrame % S; // slices in the last group of a full frame (its own tensor map)
    constexpr uint32_t kFull = 0xffffffffu;
    constexpr uint32_t kRowWords = Cfg::kOutRowBytes / 4;
    constexpr uint32_t kSliceWords = Cfg::kSliceBytes / 4;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t pad = (kLutAlign - (smem_u32(smem_raw) & (kLutAlign - 1))) & (kLutAlign - 1);
    uint8_t *smem = smem_raw + pad;
    fill_lut(smem);
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t fi = (uint64_t)blockIdx.x * THREADS + threadIdx.x;

    // ---- per-lane frame setup (lanes without work keep nslices == 0 and only take part in the warp-level steps)
    uint32_t nsamp = 0;
    uint64_t out_byte = 0; // byte offset of this frame's PCM inside the output arena
    FrameRec rec;
    rec.in_off = 0;
    if (fi < n_frames) {
        rec = recs[fi];
        const uint64_t ob = out_base[rec.stream];
        if (ob != ~0ull) {
            nsamp = rec.nsamp;
            out_byte = (ob + rec.out_rel) * 2;
        }
    }
    const uint32_t nslices = (nsamp + kSliceLen - 1) / kSliceLen;
    const uint32_t nsl_max = __reduce_max_sync(kFull, nslices);
    if (nsl_max == 0) return;
    // The 32 rows of a warp's tile go out as 32 / RB tensor stores of RB rows each.  A box is "regular" when its RB
    // frames sit exactly one full frame of PCM apart in the output (same stream, consecutive frames): only boxes that
    // contain a stream boundary (or idle lanes) take the row-by-row copy path.
    constexpr uint32_t RB = Cfg::kBoxRows;
    const uint32_t leader = lane & ~(RB - 1);
    const uint64_t out0 = __shfl_sync(kFull, out_byte, leader);
    const bool lane_ok = use_map && nsamp != 0 && out_byte == out0 + (uint64_t)(lane - leader) * Cfg::kOutRowBytes;
    const uint32_t ok_mask = __ballot_sync(kFull, lane_ok);
    const uint32_t box_mask = (RB == 32 ? 0xffffffffu : ((1u << RB) - 1u)) << leader;
    const bool regular = (ok_mask & box_mask) == box_mask;
    uint32_t nsamp_min = nsamp; // over the box
#pragma unroll
    for (uint32_t m = RB / 2; m >= 1; m >>= 1) nsamp_min = min(nsamp_min, __shfl_xor_sync(kFull, nsamp_min, m));
    int map_c1 = (int)(out0 / Cfg::kOutRowBytes);       // tensor-map row of the box leader's frame start
    int map_c0 = (int)((out0 % Cfg::kOutRowBytes) / 4); // 32-bit word inside that row

    const uint8_t *fp = arena + rec.in_off;
    Chain a, b;
    const uint8_t *sp = fp + 8 + 16 * CH;                           // first slice word (8-byte aligned)
    const uint32_t skew = (uint32_t)(reinterpret_cast<uintptr_t>(sp) & 8);
    const uint4 *vp = reinterpret_cast<const uint4 *>(sp - skew);    // 16-byte aligned superset
    const uint32_t nvec = nslices ? (skew + nslices * 8 * CH + 15) / 16 : 0; // vectors holding any byte of the slices
    // input ring: vector j of this lane lives in slot j % 4
    const uint32_t in_ring = smem_u32(smem + kLutBytes) + (THREADS / 32) * Cfg::kWarpBytes +
                             (threadIdx.x >> 5) * Cfg::kInWarpBytes + lane * 16;
    auto fetch = [&](uint32_t j) {
        if (j < nvec) cp_async16(in_ring + (j & 3) * 512, vp + j);
    };
    // stereo: one vector per slice, three in flight.  mono: one vector per two slices, fetched on even slices.
    fetch(0);
    cp_async_commit();
    fetch(1);
    cp_async_commit();
    if (CH == 2) {
        fetch(2);
        cp_async_commit();
    }
    a = Chain{0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1};
    b = a;
    if (nslices) {
        load_lms(a, fp + 8);
        if (CH == 2) load_lms(b, fp + 24);
    }
    uint8_t *op = reinterpret_cast<uint8_t *>(out) + out_byte;
    const uint32_t lane_row = smem_u32(smem) + (lane & 15) * 8;
    const uint32_t warp_tiles = smem_u32(smem + kLutBytes) + (threadIdx.x >> 5) * Cfg::kWarpBytes;
    // per-warp table {PCM pointer, nsamp} of the 32 rows for the row-by-row copy path
    const uint32_t rows_meta = smem_u32(smem + kLutBytes) + (THREADS / 32) * (Cfg::kWarpBytes + Cfg::kInWarpBytes) +
                               (threadIdx.x >> 5) * 512;
    {
        const uint64_t p64 = reinterpret_cast<uint64_t>(op);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rows_meta + lane * 16), "r"((uint32_t)p64),
                     "r"((uint32_t)(p64 >> 32)), "r"(nsamp), "r"(0u)
                     : "memory");
        __syncwarp();
    }

    // One slice of this lane's frame: read its words out of the input ring, decode 20 samples per channel, stage the
    // PCM at dst_s.  (Both loops below inline it.)
    auto fetch_ahead = [&](uint32_t k) {
        if (CH == 2) {
            fetch(k + 3);
            cp_async_commit();
            cp_async_wait<2>(); // vectors k and k + 1 have landed (only k + 2 and k + 3 may still be in flight)
        } else {
            if ((k & 1) == 0) fetch(k / 2 + 2);
            cp_async_commit();
            cp_async_wait<2>(); // everything fetched up to two slices ago has landed: vectors <= (k + 1) / 2
        }
    };
    auto slice_body = [&](uint32_t k, uint32_t dst_s) {
        if (CH == 2) {
            // slice k = bytes [skew + 16k, skew + 16k + 16) of the vector stream: no alignment selects needed
            const uint2 ra = lds64(in_ring + (k & 3) * 512 + skew);
            const uint2 rb = lds64(in_ring + ((k + (skew >> 3)) & 3) * 512 + (8 - skew));
            const uint2 wa = make_uint2(bswap32(ra.x), bswap32(ra.y));
            const uint2 wb = make_uint2(bswap32(rb.x), bswap32(rb.y));
            const uint32_t row_a = lane_row + (wa.x >> 28) * 1024; // sf = top nibble (:303)
            const uint32_t row_b = lane_row + (wb.x >> 28) * 1024;
            uint32_t o[20];
            PairLoop<0>::run(a, b, row_a, row_b, wa, wb, o);
#pragma unroll
            for (int i = 0; i < 5; ++i) sts128(dst_s + i * 16, o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
        } else {
            // slice k = bytes [skew + 8k, skew + 8k + 8) of the vector stream
            const uint32_t byte = skew + 8 * k;
            const uint2 ra = lds64(in_ring + ((byte >> 4) & 3) * 512 + (byte & 8));
            const uint2 wa = make_uint2(bswap32(ra.x), bswap32(ra.y));
            const uint32_t row_a = lane_row + (wa.x >> 28) * 1024;
            uint32_t o[10];
            SingleLoop<0>::run(a, row_a, wa, o);
#pragma unroll
            for (int i = 0; i < 5; ++i) sts64(dst_s + i * 8, o[2 * i], o[2 * i + 1]); // 40 bytes per slice: rows stay 8-byte aligned
        }
    };

    // ---- uniform warps: 32 full frames that are neighbours in one stream (94 % of the warps of a 60 s batch) ------
    // Nothing is data dependent here: every lane runs 256 slices, every group of S slices leaves as ONE 32-row tensor
    // store issued by lane 0, and none of the bookkeeping of the general loop below (per-lane slice counts, per-box
    // regularity, the row-by-row copy path) is needed: ~35 instructions per slice less.
    if (NB == 1 && use_map == 2) {
        const uint64_t out00 = __shfl_sync(kFull, out_byte, 0);
        if (__all_sync(kFull, nsamp == (uint32_t)kFrameLen && out_byte == out00 + (uint64_t)lane * Cfg::kOutRowBytes)) {
            constexpr uint32_t kMain = kSlicesPerFrame - kRem; // slices that leave in full groups
            const uint32_t tile = warp_tiles;
            int c1 = (int)(out00 / Cfg::kOutRowBytes), c0 = (int)((out00 % Cfg::kOutRowBytes) / 4);
            uint32_t dst = tile + lane * Cfg::kRowBytes, j = 0;
#pragma unroll 1
            for (uint32_t k = 0; k < (uint32_t)kSlicesPerFrame; ++k) {
                if (j == 0) { // the tensor store that last read this tile has drained it (moving this wait behind the
                    bulk_wait_read<0>(); // slice's arithmetic was measured: no gain, the other warps hide it)
                    __syncwarp();
                }
                fetch_ahead(k);
                slice_body(k, dst);
                dst += Cfg::kSliceBytes;
                if (++j == S) {
                    fence_async_smem(); // staging stores -> visible to the async proxy
                    __syncwarp();
                    if (lane == 0) tensor_store_2d(&out_map32, tile, c0, c1);
                    bulk_commit();
                    c0 += S * kSliceWords;
                    if (c0 >= (int)kRowWords) {
                        c0 -= (int)kRowWords;
                        c1 += 1;
                    }
                    j = 0; // the last, shorter group of the frame is staged densely (pitch = its box width)
                    dst = tile + lane * (k + 1 == kMain ? kRem * Cfg::kSliceBytes : (uint32_t)Cfg::kRowBytes);
                }
            }
            if (kRem) {
                fence_async_smem();
                __syncwarp();
                if (lane == 0) tensor_store_2d(&out_map32_rem, tile, c0, c1);
                bulk_commit();
            }
            bulk_wait_read<0>(); // shared memory must stay valid until the TMA engine has read the tile
            return;
        }
    }

    uint32_t ring = 0;   // tile of the ring being filled
    uint32_t in_row = 0; // slices already staged in it
    uint32_t first = 0;  // per-channel sample index of the tile's first sample
    uint32_t pitch = Cfg::kRowBytes; // row pitch of the tile being filled: the warp's last, shorter group is staged densely

#pragma unroll 1
    for (uint32_t k = 0; k < nsl_max; ++k) {
        const uint32_t tile = warp_tiles + ring * Cfg::kTileBytes;
        if (in_row == 0) { // the copies that last read this tile have drained it
            bulk_wait_read<NB - 1>();
            __syncwarp();
            pitch = min((uint32_t)S, nsl_max - k) * Cfg::kSliceBytes; // TMA reads the tile densely: pitch = box width
        }
        const uint32_t row = tile + lane * pitch;
        fetch_ahead(k);
        if (k < nslices) slice_body(k, row + in_row * Cfg::kSliceBytes);
        ++in_row;
        if (in_row == S || k + 1 == nsl_max) {
            fence_async_smem(); // staging stores -> visible to the async proxy
            __syncwarp();
            const bool whole = regular && (in_row == S || in_row == kRem) && first + in_row * kSliceLen <= nsamp_min &&
                               (use_map == 2 || map_c0 + in_row * kSliceWords <= kRowWords);
            const uint32_t whole_rows = __ballot_sync(kFull, whole); // bit r: row r leaves through a tensor store
            if (whole && lane == leader) // every lane of the box staged the same number of full slices
                tensor_store_2d(in_row == S ? &out_map : &out_map_rem, tile + leader * pitch, map_c0, map_c1);
            if (whole_rows != kFull) {
                // Rows of irregular boxes: the warp copies them one half-warp per row (16 lanes x 16 bytes cover a
                // 240-byte row), so every store instruction writes two contiguous runs instead of 32 scattered 16-byte
                // pieces.  Samples past a row's nsamp (partial last slice, :320-325) are not emitted.
#pragma unroll 1
                for (uint32_t it = 0; it < 16; ++it) {
                    if (((whole_rows >> (2 * it)) & 3u) == 3u) continue;
                    const uint32_t r = it * 2 + (lane >> 4);
                    if ((whole_rows >> r) & 1u) continue;
                    const uint32_t off = (lane & 15) * 16;
                    const uint2 pr = lds64(rows_meta + r * 16);
                    uint32_t ns;
                    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(ns) : "r"(rows_meta + r * 16 + 8));
                    const uint32_t bytes = first < ns ? min(in_row * (uint32_t)kSliceLen, ns - first) * (2 * CH) : 0u;
                    uint8_t *dst = reinterpret_cast<uint8_t *>(((uint64_t)pr.y << 32) | pr.x) + (uint64_t)first * (2 * CH);
                    const uint32_t src = tile + r * pitch;
                    if (off + 16 <= bytes) {
                        uint4 v;
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(src + off));
                        *reinterpret_cast<uint4 *>(dst + off) = v;
                    } else {
                        for (uint32_t t = off; t < bytes; t += 2) { // trailing samples of a row that is not a multiple of 16 bytes
                            uint16_t v;
                            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(src + t));
                            *reinterpret_cast<uint16_t *>(dst + t) = v;
                        }
                    }
                }
            }
            bulk_commit();
            first += in_row * kSliceLen;
            map_c0 += in_row * kSliceWords;
            if (map_c0 >= (int)kRowWords) {
                map_c0 -= (int)kRowWords;
                map_c1 += 1;
            }
            in_row = 0;
            ring = (ring + 1 == NB) ? 0 : ring + 1;
        }
    }
    bulk_wait_read<0>(); // shared memory must stay valid until the TMA engine has read every tile
}

// Encodes the tensor map that views the output arena as rows of `row_bytes` (the PCM of one full frame) of 32-bit words;
// box = box_words x box_rows.
// Returns 0 when no map could be made (driver entry point unavailable, arena not 16-byte aligned: the kernel then
// copies every row itself), 2 when the map's rows overlap (dim 0 is extended by one box beyond the row pitch,
// so a box may start anywhere in a row and run into the next one), 1 when the driver refused overlapping rows (boxes
// that would cross a row end are then left to the row copy path).
static int make_out_map(CUtensorMap *map, void *d_out, uint64_t out_bytes, uint32_t box_words, uint32_t row_bytes,
                        uint32_t box_rows)
{
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                 const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    static std::once_flag looked_up; // contexts on different host threads (qoa_b200_mctx) may get here at the same time
    std::call_once(looked_up, [] {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            encode = reinterpret_cast<EncodeFn>(fn);
        else
            cudaGetLastError();
    });
    if (!encode || (reinterpret_cast<uintptr_t>(d_out) & 15) || out_bytes == 0) return 0;
    const uint64_t rows = (out_bytes + row_bytes - 1) / row_bytes + 1;
    if (rows > 0x7fffffffull) return 0;
    const cuuint64_t strides[1] = {row_bytes};
    const cuuint32_t box[2] = {box_words, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    for (int overlap = 1; overlap >= 0; --overlap) {
        const cuuint64_t dims[2] = {row_bytes / 4 + (overlap ? box_words : 0), rows};
        if (encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d_out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
            return overlap ? 2 : 1;
    }
    return 0;
}

// ---- uniform multi-channel batches (3..8 channels) ---------------------------------------------------------------
// One thread = two channel chains, the same two-chain slice loop as stereo.  Even channel counts: the lane owns a
// channel pair of one frame, the ch / 2 lanes of a frame are neighbours.  Odd channel counts: frames are taken two at
// a time and the 2 x (ch / 2) + 1 = ch lanes of such a frame pair are (ch / 2) pairs of the first frame, (ch / 2)
// pairs of the second, and ONE lane whose two chains are the last channel of each of the two frames - every chain a
// lane computes is useful (round 1 gave every frame its own half-empty last pair: 3 channels used 3/4 of the
// arithmetic, 5 channels 5/6).  Every chain writes its samples into its FRAME's row of a per-warp shared-memory tile
// ([frames][S slices x 20 samples x ch] interleaved i16; S = 1 for even, 2 for odd channel counts so that a row is a
// multiple of 16 bytes).  After S slices the tile leaves
//   * through ONE 2-D tensor store (rows = frames, box = one group of one frame's interleaved PCM) when the warp's
//     frames are full-size neighbours in one stream - no copy-out instructions (round 1 copied every tile out with
//     LDS.128 + STG.128: 368 of 1097 instructions per slice for 8 channels);
//   * otherwise row by row, one half-warp per 256-byte run, so global stores are contiguous runs.
// The tensor-store path stages densely (the TMA engine reads the box as one dense block), which costs bank conflicts on
// the per-sample staging stores; the copy path pads the row pitch to an odd multiple of 16 bytes instead.
// Slice words arrive through a per-lane cp.async ring three slices ahead.  Requirements (host-checked): every frame
// has CH channels and starts 16-byte aligned.
constexpr int kMultiThreads = 256;
constexpr int kMultiTileBytes = 5632; // largest tile: 4 channels = 16 frames x (2 slices x 160 B + 16 B pad) = 5376

template <int CH>
struct MultiCfg {
    static constexpr bool kOdd = (CH & 1) != 0;
    static constexpr uint32_t H = CH / 2;                              // whole channel pairs per frame
    static constexpr uint32_t kGroupLanes = kOdd ? CH : H;             // lanes per frame (even) / per frame pair (odd)
    static constexpr uint32_t kGroups = 32 / kGroupLanes;
    static constexpr uint32_t kFpw = kOdd ? 2 * kGroups : kGroups;     // frames per warp
    static constexpr uint32_t S = kOdd ? 2u : 1u;                      // slices per tile row
    static constexpr uint32_t kSliceBytes = kSliceLen * CH * 2;        // interleaved PCM of one slice of one frame
    static constexpr uint32_t kRowBytes = S * kSliceBytes;
    static constexpr uint32_t kPadPitch = kRowBytes + (((kRowBytes >> 4) & 1) ? 0 : 16); // copy path: odd multiple of 16
    static constexpr uint32_t kOutRowBytes = kFrameLen * CH * 2;       // PCM bytes of one full frame = tensor-map row pitch
    static_assert(kRowBytes % 16 == 0 && kFpw * kPadPitch <= (uint32_t)kMultiTileBytes, "tile layout");
    static_assert(kRowBytes / 4 <= 256 && kFpw <= 256 && (kSlicesPerFrame % S) == 0, "tensor-map box limits");
};

__device__ __forceinline__ void cp_async8(uint32_t saddr, const void *gptr)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(saddr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void sts_u16(uint32_t saddr, uint32_t v)
{
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(saddr), "h"((uint16_t)v) : "memory");
}

template <int CH>
__global__ void __launch_bounds__(kMultiThreads, 2)
decode_multi_kernel(const uint8_t *__restrict__ arena, const FrameRec *__restrict__ recs, uint64_t n_frames,
                    const uint64_t *__restrict__ out_base, int16_t *__restrict__ out,
                    const __grid_constant__ CUtensorMap out_map, int use_map)
{
    using Cfg = MultiCfg<CH>;
    constexpr uint32_t kFull = 0xffffffffu;
    constexpr int kWarps = kMultiThreads / 32;
    constexpr uint32_t ch = CH, fpw = Cfg::kFpw, S = Cfg::S, row_bytes = Cfg::kRowBytes;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t pad = (kLutAlign - (smem_u32(smem_raw) & (kLutAlign - 1))) & (kLutAlign - 1);
    uint8_t *smem = smem_raw + pad;
    fill_lut(smem);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // this lane's two chains: (frame fa, channel ca) and (frame fb, channel cb); frames are indices inside the warp
    const uint32_t g = lane / Cfg::kGroupLanes, q = lane - g * Cfg::kGroupLanes;
    uint32_t fa, fb, ca, cb;
    if (!Cfg::kOdd) {
        fa = fb = g;
        ca = 2 * q;
        cb = ca + 1;
    } else if (q < 2 * Cfg::H) {
        fa = fb = 2 * g + (q >= Cfg::H ? 1u : 0u);
        ca = 2 * (q >= Cfg::H ? q - Cfg::H : q);
        cb = ca + 1;
    } else { // the odd channels of both frames of the pair
        fa = 2 * g;
        fb = fa + 1;
        ca = cb = ch - 1;
    }
    const bool lane_on = g < Cfg::kGroups;
    const bool same = fa == fb;
    const uint64_t f0 = ((uint64_t)blockIdx.x * kWarps + warp) * fpw;

    uint32_t nsamp_a = 0, nsamp_b = 0;
    uint64_t ob_a = 0, ob_b = 0; // byte offsets of the two frames' PCM inside the output arena
    FrameRec ra_, rb_;
    ra_.in_off = rb_.in_off = 0;
    if (lane_on && f0 + fa < n_frames) {
        ra_ = recs[f0 + fa];
        const uint64_t ob = out_base[ra_.stream];
        if (ob != ~0ull) {
            nsamp_a = ra_.nsamp;
            ob_a = (ob + ra_.out_rel) * 2;
        }
    }
    if (same) {
        rb_ = ra_;
        nsamp_b = nsamp_a;
        ob_b = ob_a;
    } else if (lane_on && f0 + fb < n_frames) {
        rb_ = recs[f0 + fb];
        const uint64_t ob = out_base[rb_.stream];
        if (ob != ~0ull) {
            nsamp_b = rb_.nsamp;
            ob_b = (ob + rb_.out_rel) * 2;
        }
    }
    const uint32_t nsl_a = (nsamp_a + kSliceLen - 1) / kSliceLen, nsl_b = (nsamp_b + kSliceLen - 1) / kSliceLen;
    const uint32_t nsl_max = __reduce_max_sync(kFull, max(nsl_a, nsl_b));
    if (nsl_max == 0) return;
    // Tensor-store mode (warp-uniform): every frame of the warp is a full frame and frame f sits f rows below frame 0.
    const uint64_t out00 = __shfl_sync(kFull, ob_a, 0);
    const bool tma = use_map == 2 &&
                     __all_sync(kFull, !lane_on || (nsamp_a == (uint32_t)kFrameLen && nsamp_b == (uint32_t)kFrameLen &&
                                                    ob_a == out00 + (uint64_t)fa * Cfg::kOutRowBytes &&
                                                    ob_b == out00 + (uint64_t)fb * Cfg::kOutRowBytes));
    const uint32_t pitch = tma ? row_bytes : Cfg::kPadPitch;
    int map_c1 = (int)(out00 / Cfg::kOutRowBytes), map_c0 = (int)((out00 % Cfg::kOutRowBytes) / 4);

    const uint8_t *fp_a = arena + ra_.in_off, *fp_b = arena + rb_.in_off;
    Chain a, b;
    a = Chain{0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1};
    b = a;
    if (nsl_a) load_lms(a, fp_a + 8 + 16 * ca);
    if (nsl_b) load_lms(b, fp_b + 8 + 16 * cb);
    // the chains' words of slice 0; slice k is k * ch * 8 bytes further
    const uint8_t *sp_a = fp_a + 8 + 16 * ch + 8 * ca, *sp_b = fp_b + 8 + 16 * ch + 8 * cb;
    const uint32_t lane_row = smem_u32(smem) + (lane & 15) * 8;
    // shared memory after the LUT: per warp [tile: frames x pitch <= kMultiTileBytes][ring: 4 x 32 x 16][row table: 32 x 16]
    const uint32_t warp_base = smem_u32(smem + kLutBytes) + warp * (kMultiTileBytes + 2048 + 512);
    const uint32_t tile = warp_base;
    const uint32_t in_ring = warp_base + kMultiTileBytes + lane * 16;
    const uint32_t rows_meta = warp_base + kMultiTileBytes + 2048;
    if (lane_on && ca == 0) { // one table entry per frame: {PCM pointer, nsamp}; the lane of channel pair 0 writes it
        const uint64_t p64 = reinterpret_cast<uint64_t>(reinterpret_cast<uint8_t *>(out) + ob_a);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rows_meta + fa * 16), "r"((uint32_t)p64),
                     "r"((uint32_t)(p64 >> 32)), "r"(nsamp_a), "r"(0u)
                     : "memory");
    }
    auto fetch = [&](uint32_t k) { // slice k's words into ring slot k % 4
        if (k < nsl_a) cp_async8(in_ring + (k & 3) * 512, sp_a + (uint64_t)k * ch * 8);
        if (k < nsl_b) cp_async8(in_ring + (k & 3) * 512 + 8, sp_b + (uint64_t)k * ch * 8);
        cp_async_commit();
    };
    fetch(0);
    fetch(1);
    fetch(2);
    __syncwarp();

    // the chains' first samples inside their frames' rows
    const uint32_t row_a0 = tile + fa * pitch + ca * 2, row_b0 = tile + fb * pitch + cb * 2;
    uint32_t in_row = 0, first = 0;
#pragma unroll 1
    for (uint32_t k = 0; k < nsl_max; ++k) {
        fetch(k + 3);
        cp_async_wait<3>(); // slice k's words have landed
        uint32_t o[20];
        const bool on_a = k < nsl_a, on_b = k < nsl_b;
        if (on_a || on_b) {
            uint2 xa = lds64(in_ring + (k & 3) * 512), xb = lds64(in_ring + (k & 3) * 512 + 8);
            if (!on_a) xa = xb; // a finished chain decodes its neighbour's word (valid table addresses), unused
            if (!on_b) xb = xa;
            const uint2 wa = make_uint2(bswap32(xa.x), bswap32(xa.y));
            const uint2 wb = make_uint2(bswap32(xb.x), bswap32(xb.y));
            const uint32_t row_a = lane_row + (wa.x >> 28) * 1024; // sf = top nibble (:303)
            const uint32_t row_b = lane_row + (wb.x >> 28) * 1024;
            PairLoop<0>::run(a, b, row_a, row_b, wa, wb, o);
        }
        if (tma && in_row == 0) { // the tensor store that last read the tile has drained it: with one-slice groups this
            bulk_wait_read<0>();  // wait comes every slice, so it sits BEHIND the slice's arithmetic (in front of it the
            __syncwarp();         // warps spent 2.4 stall cycles per issue on it: profiles/r02_decode_multi_8ch_*.txt)
        }
        const uint32_t da = row_a0 + in_row * Cfg::kSliceBytes, db = row_b0 + in_row * Cfg::kSliceBytes;
        if (!Cfg::kOdd) { // even channel counts: the pair is one aligned 32-bit word per sample (both chains, one frame)
            if (on_a) {
#pragma unroll
                for (int i = 0; i < 20; ++i) asm volatile("st.shared.b32 [%0], %1;" ::"r"(da + i * ch * 2), "r"(o[i]) : "memory");
            }
        } else {
            if (on_a) {
#pragma unroll
                for (int i = 0; i < 20; ++i) sts_u16(da + i * ch * 2, o[i] & 0xffffu);
            }
            if (on_b) {
#pragma unroll
                for (int i = 0; i < 20; ++i) sts_u16(db + i * ch * 2, o[i] >> 16);
            }
        }
        ++in_row;
        if (in_row == S || k + 1 == nsl_max) {
            if (tma) { // (full frames: in_row == S here)
                fence_async_smem(); // staging stores -> visible to the async proxy
                __syncwarp();
                if (lane == 0) tensor_store_2d(&out_map, tile, map_c0, map_c1);
                bulk_commit();
                map_c0 += (int)(row_bytes / 4);
                if (map_c0 >= (int)(Cfg::kOutRowBytes / 4)) {
                    map_c0 -= (int)(Cfg::kOutRowBytes / 4);
                    map_c1 += 1;
                }
                first += in_row * kSliceLen;
                in_row = 0;
                continue;
            }
            __syncwarp();
            // copy the tile out: one half-warp per 256-byte run of a row
            constexpr uint32_t runs = (row_bytes + 255) / 256;
            for (uint32_t it = lane >> 4; it < fpw * runs; it += 2) {
                const uint32_t r = it / runs, run = it - r * runs;
                const uint32_t off = run * 256 + (lane & 15) * 16;
                const uint2 pr = lds64(rows_meta + r * 16);
                uint32_t ns;
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(ns) : "r"(rows_meta + r * 16 + 8));
                // Samples past nsamp (partial last slice, :320-325) are not emitted.
                const uint32_t bytes = first < ns ? min(in_row * (uint32_t)kSliceLen, ns - first) * ch * 2 : 0u;
                uint8_t *gdst = reinterpret_cast<uint8_t *>(((uint64_t)pr.y << 32) | pr.x) + (uint64_t)first * ch * 2;
                const uint32_t src = tile + r * pitch;
                if (off + 16 <= bytes) {
                    uint4 v;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(src + off));
                    *reinterpret_cast<uint4 *>(gdst + off) = v;
                } else {
                    for (uint32_t t = off; t < min(bytes, off + 16); t += 2) {
                        uint16_t v;
                        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(src + t));
                        *reinterpret_cast<uint16_t *>(gdst + t) = v;
                    }
                }
            }
            __syncwarp();
            first += in_row * kSliceLen;
            in_row = 0;
        }
    }
    cp_async_wait<0>();
    bulk_wait_read<0>(); // shared memory must stay valid until the TMA engine has read the tile
}

template <int CH>
static cudaError_t launch_multi(const uint8_t *d_arena, const FrameRec *d_recs, uint64_t n_frames,
                                const uint64_t *d_out_base, int16_t *d_out, uint64_t out_bytes, cudaStream_t s)
{
    using Cfg = MultiCfg<CH>;
    const size_t smem = kLutAlign + kLutBytes + (size_t)(kMultiThreads / 32) * (kMultiTileBytes + 2048 + 512);
    static PerDeviceOnce configured;
    std::atomic<bool> *cfg = configured.slot();
    if (!cfg || !cfg->load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(decode_multi_kernel<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (cfg) cfg->store(true, std::memory_order_release);
    }
    CUtensorMap map;
    memset(&map, 0, sizeof map);
    int use_map = make_out_map(&map, d_out, out_bytes, Cfg::kRowBytes / 4, Cfg::kOutRowBytes, Cfg::kFpw);
    if (const char *dbg = getenv("QOA_B200_TMAP")) use_map = std::min(use_map, atoi(dbg)); // development aid
    const uint64_t warps = (n_frames + Cfg::kFpw - 1) / Cfg::kFpw;
    const uint64_t blocks = (warps + kMultiThreads / 32 - 1) / (kMultiThreads / 32);
    if (blocks > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    decode_multi_kernel<CH><<<(unsigned)blocks, kMultiThreads, smem, s>>>(d_arena, d_recs, n_frames, d_out_base, d_out, map,
                                                                          use_map);
    return cudaGetLastError();
}

// ---- generic path: one thread = one (frame, channel); any channel count (1..255), any alignment ---------------
constexpr int kGenThreads = 128;

__global__ void __launch_bounds__(kGenThreads)
decode_generic_kernel(const uint8_t *__restrict__ arena, const FrameRec *__restrict__ recs, uint64_t n_units,
                      uint32_t max_ch, const uint64_t *__restrict__ out_base, int16_t *__restrict__ out)
{
    // static shared memory starts at window offset 0x400 and __align__ is relative to that: align by hand
    __shared__ __align__(16) uint8_t lut_raw[kLutBytes + kLutAlign];
    uint8_t *lut = lut_raw + ((kLutAlign - (smem_u32(lut_raw) & (kLutAlign - 1))) & (kLutAlign - 1));
    fill_lut(lut);
    const uint64_t u = (uint64_t)blockIdx.x * kGenThreads + threadIdx.x;
    if (u >= n_units) return;
    const uint64_t fi = u / max_ch;
    const uint32_t c = (uint32_t)(u - fi * max_ch);
    const FrameRec rec = recs[fi];
    const uint32_t ch = rec.ch;
    const uint32_t nsamp = rec.nsamp;
    if (c >= ch || nsamp == 0) return;
    const uint64_t ob = out_base[rec.stream];
    if (ob == ~0ull) return;

    const uint8_t *fp = arena + rec.in_off;
    Chain a;
    load_lms(a, fp + 8 + 16 * c);
    const uint8_t *sp = fp + 8 + 16 * ch + 8 * c;
    const uint32_t ngroups = (nsamp + kSliceLen - 1) / kSliceLen;
    int16_t *op = out + ob + rec.out_rel + c;
    const uint32_t lane_row = smem_u32(lut) + (threadIdx.x & 15) * 8;

    for (uint32_t k = 0; k < ngroups; ++k) {
        const uint2 wa = ldg_be64(sp + (uint64_t)k * ch * 8);
        const uint32_t row_a = lane_row + (wa.x >> 28) * 1024;
        uint32_t o[10];
        SingleLoop<0>::run(a, row_a, wa, o);
        const uint32_t base = k * kSliceLen;
        const uint32_t emit = min((uint32_t)kSliceLen, nsamp - base); // :320-328
#pragma unroll
        for (int i = 0; i < 20; ++i)
            if ((uint32_t)i < emit) op[(uint64_t)(base + i) * ch] = (int16_t)(o[i / 2] >> (16 * (i & 1)));
    }
}

template <int CH, int S, int NB, int THREADS>
static cudaError_t launch_fast(const uint8_t *d_arena, const FrameRec *d_recs, uint64_t n_frames,
                               const uint64_t *d_out_base, int16_t *d_out, uint64_t out_bytes, cudaStream_t s)
{
    using Cfg = FastCfg<CH, S, NB>;
    const size_t smem = kLutAlign + kLutBytes +
                        (size_t)(THREADS / 32) * (Cfg::kWarpBytes + Cfg::kInWarpBytes + 512);
    static PerDeviceOnce configured;
    std::atomic<bool> *cfg = configured.slot();
    if (!cfg || !cfg->load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(decode_fast_kernel<CH, S, NB, THREADS>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (cfg) cfg->store(true, std::memory_order_release);
    }
    CUtensorMap map, map_rem, map32, map32_rem;
    memset(&map, 0, sizeof map);
    memset(&map_rem, 0, sizeof map_rem);
    memset(&map32, 0, sizeof map32);
    memset(&map32_rem, 0, sizeof map32_rem);
    constexpr uint32_t kRem = kSlicesPerFrame % S;
    int use_map = make_out_map(&map, d_out, out_bytes, Cfg::kRowBytes / 4, Cfg::kOutRowBytes, Cfg::kBoxRows);
    if (use_map && kRem) // the last group of a full frame holds 256 % S slices: same tile pitch, narrower box
        use_map = std::min(use_map, make_out_map(&map_rem, d_out, out_bytes, kRem * Cfg::kSliceBytes / 4, Cfg::kOutRowBytes,
                                                 Cfg::kBoxRows));
    if (use_map == 2) { // the same views with 32-row boxes: warps whose 32 frames are neighbours in one stream
        int wide = make_out_map(&map32, d_out, out_bytes, Cfg::kRowBytes / 4, Cfg::kOutRowBytes, 32);
        if (wide == 2 && kRem)
            wide = make_out_map(&map32_rem, d_out, out_bytes, kRem * Cfg::kSliceBytes / 4, Cfg::kOutRowBytes, 32);
        if (wide != 2) use_map = 1; // the uniform-warp loop needs use_map == 2
    }
    if (const char *dbg = getenv("QOA_B200_TMAP")) use_map = std::min(use_map, atoi(dbg)); // development aid: 0 = no tensor stores, 1 = no overlapping rows
    const uint64_t blocks = (n_frames + THREADS - 1) / THREADS;
    if (blocks > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    decode_fast_kernel<CH, S, NB, THREADS><<<(unsigned)blocks, THREADS, smem, s>>>(
        d_arena, d_recs, n_frames, d_out_base, d_out, map, map_rem, map32, map32_rem, use_map);
    return cudaGetLastError();
}

void probe_stream(const uint8_t *qoa, uint64_t len, int mode, qoa_b200_stream_info *info)
{
    walk_stream<false>(qoa, 0, len, mode, 0, nullptr, info);
}

cudaError_t launch_decode(const uint8_t *d_arena, const FrameRec *d_recs, uint64_t n_frames, uint32_t max_ch,
                          uint32_t uniform_ch, const uint64_t *d_out_base, int16_t *d_out, uint64_t out_bytes,
                          cudaStream_t s, int *n_launches)
{
    if (n_frames == 0 || max_ch == 0) return cudaSuccess;
    cudaError_t e;
    if (uniform_ch == 2) {
        // 256 threads x 2 CTAs per SM: the measured best of 128 / 192 / 256 / 288 / 320 / 512 / 640 threads
        e = launch_fast<2, 3, 1, 256>(d_arena, d_recs, n_frames, d_out_base, d_out, out_bytes, s);
    } else if (uniform_ch == 1) {
        e = launch_fast<1, 6, 1, 256>(d_arena, d_recs, n_frames, d_out_base, d_out, out_bytes, s);
    } else if (uniform_ch >= 3 && uniform_ch <= 8) {
        switch (uniform_ch) {
        case 3: e = launch_multi<3>(d_arena, d_recs, n_frames, d_out_base, d_out, out_bytes, s); break;
        case 4: e = launch_multi<4>(d_arena, d_recs, n_frames, d_out_base, d_out, out_bytes, s); break;
        case 5: e = launch_multi<5>(d_arena, d_recs, n_frames, d_out_base, d_out, out_bytes, s); break;
        case 6: e = launch_multi<6>(d_arena, d_recs, n_frames, d_out_base, d_out, out_bytes, s); break;
        case 7: e = launch_multi<7>(d_arena, d_recs, n_frames, d_out_base, d_out, out_bytes, s); break;
        default: e = launch_multi<8>(d_arena, d_recs, n_frames, d_out_base, d_out, out_bytes, s); break;
        }
    } else {
        const uint64_t n_units = n_frames * max_ch;
        const uint64_t blocks = (n_units + kGenThreads - 1) / kGenThreads;
        if (blocks > 0x7fffffffull) return cudaErrorInvalidConfiguration;
        decode_generic_kernel<<<(unsigned)blocks, kGenThreads, 0, s>>>(d_arena, d_recs, n_units, max_ch, d_out_base, d_out);
        e = cudaGetLastError();
    }
    if (n_launches) *n_launches += 1;
    return e;
}

} // namespace qoa
