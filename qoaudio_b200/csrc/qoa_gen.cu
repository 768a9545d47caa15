// qoa_gen.cu — on-device synthetic workload generators and the integer-issue microbenchmark (sm_100a).
//
// The PCM recipe is integer-only and counter-based so that tests/signals.py (numpy) and this kernel produce
// bit-identical data; it is measurement plumbing, not part of the reference's path.
#include "qoa_kernels.h"

namespace qoa {

__device__ __forceinline__ int tri16(uint32_t phase)
{
    const uint32_t ph = phase & 0xffffu;
    const int t = (int)(ph < 32768u ? ph : 65535u - ph); // 0..32767
    return t * 2 - 32767;
}

__global__ void gen_pcm_kernel(int16_t *__restrict__ pcm, uint64_t n_streams, uint64_t first_stream, uint32_t channels,
                               uint64_t samples, uint64_t start, uint64_t seed, int kind)
{
    const uint64_t per_stream = samples * channels;
    const uint64_t total = n_streams * per_stream;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t s = e / per_stream;
        const uint64_t r = e - s * per_stream;
        const uint64_t i = r / channels + start;
        const uint32_t c = (uint32_t)(r % channels);
        const uint64_t k = stream_key(seed, first_stream + s, c);
        const uint64_t h = splitmix64(k + i);
        int v;
        if (kind == 1) {
            v = (int)(int16_t)(uint16_t)(h & 0xffffu);
        } else {
            const uint32_t f1 = 40u + (uint32_t)(k & 0x3ffu);
            const uint32_t f2 = 300u + (uint32_t)((k >> 10) & 0xfffu);
            const int a1 = 2000 + (int)((k >> 22) & 0x1fffu);
            const int a2 = 500 + (int)((k >> 35) & 0xfffu);
            const uint32_t i32 = (uint32_t)i;
            v = ((tri16(i32 * f1) * a1) >> 15) + ((tri16(i32 * f2) * a2) >> 15);
            v += (int)(h >> 52) - 2048;
            v = max(-32768, min(32767, v));
        }
        pcm[e] = (int16_t)v;
    }
}

cudaError_t launch_gen_pcm(int16_t *d_pcm, uint64_t n_streams, uint64_t first_stream, uint32_t channels,
                           uint64_t samples, uint64_t start, uint64_t seed, int kind, cudaStream_t s)
{
    if (n_streams == 0 || samples == 0 || channels == 0) return cudaSuccess;
    gen_pcm_kernel<<<148 * 16, 256, 0, s>>>(d_pcm, n_streams, first_stream, channels, samples, start, seed, kind);
    return cudaGetLastError();
}

// Valid .qoa streams with pseudo-random LMS states and slice words.  One thread per 8-byte word.
__global__ void gen_qoa_kernel(uint8_t *__restrict__ qoa, uint64_t stride, uint64_t n_streams, uint64_t first_stream,
                               uint32_t channels, uint32_t rate, uint32_t samples, uint64_t seed)
{
    const uint64_t frames = ((uint64_t)samples + kFrameLen - 1) / kFrameLen;
    const uint64_t slices = ((uint64_t)samples + kSliceLen - 1) / kSliceLen;
    const uint64_t words = 1 + frames * (1 + 2ull * channels) + slices * channels; // encoded size / 8 (:374-378)
    const uint64_t wpf = 1 + 2ull * channels + (uint64_t)kSlicesPerFrame * channels;
    const uint64_t total = n_streams * words;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t s = e / words;
        const uint64_t j = e - s * words;
        uint32_t hi, lo;
        if (j == 0) {
            hi = kMagic;
            lo = samples;
        } else {
            const uint64_t f = (j - 1) / wpf;
            const uint64_t r = (j - 1) - f * wpf;
            if (r == 0) {
                const uint32_t frame_len = (uint32_t)min((uint64_t)kFrameLen, (uint64_t)samples - f * kFrameLen);
                const uint32_t nsl = (frame_len + kSliceLen - 1) / kSliceLen;
                hi = (channels << 24) | rate;
                lo = (frame_len << 16) | ((8 + 16 * channels + 8 * nsl * channels) & 0xffffu);
            } else {
                const uint64_t h = splitmix64(stream_key(seed, first_stream + s, 0) + j);
                hi = (uint32_t)(h >> 32);
                lo = (uint32_t)h;
            }
        }
        st_be64(qoa + s * stride + j * 8, hi, lo);
    }
}

cudaError_t launch_gen_qoa(uint8_t *d_qoa, uint64_t stride, uint64_t n_streams, uint64_t first_stream,
                           uint32_t channels, uint32_t rate, uint32_t samples, uint64_t seed, cudaStream_t s)
{
    if (n_streams == 0 || samples == 0 || channels == 0) return cudaSuccess;
    gen_qoa_kernel<<<148 * 16, 256, 0, s>>>(d_qoa, stride, n_streams, first_stream, channels, rate, samples, seed);
    return cudaGetLastError();
}

// ---- batched host <-> device copies through mapped pinned memory -------------------------------------------------
// A batch of thousands of per-stream buffers costs one cudaMemcpyAsync call each (~4 us) — more than the transfer for
// streams of a few hundred KB.  When the caller's buffers are pinned (device-accessible), ONE kernel moves them all:
// block (item, tile) copies a 64 KB tile with 16-byte accesses straight over PCIe.
constexpr uint32_t kCopyTile = 64 * 1024;

__global__ void __launch_bounds__(256) copy_items_kernel(const CopyItem *__restrict__ items)
{
    const CopyItem it = items[blockIdx.x];
    const uint64_t t0 = (uint64_t)blockIdx.y * kCopyTile;
    if (t0 >= it.bytes) return;
    const uint64_t len = min((uint64_t)kCopyTile, it.bytes - t0);
    const uint8_t *src = static_cast<const uint8_t *>(it.src) + t0;
    uint8_t *dst = static_cast<uint8_t *>(it.dst) + t0;
    const uint32_t nvec = (uint32_t)(len / 16);
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
    uint4 *d4 = reinterpret_cast<uint4 *>(dst);
    uint32_t i = threadIdx.x;
    for (; i + 3 * 256 < nvec; i += 4 * 256) { // four loads in flight per thread
        const uint4 a = s4[i], b = s4[i + 256], c = s4[i + 512], d = s4[i + 768];
        d4[i] = a;
        d4[i + 256] = b;
        d4[i + 512] = c;
        d4[i + 768] = d;
    }
    for (; i < nvec; i += 256) d4[i] = s4[i];
    for (uint32_t b = nvec * 16 + threadIdx.x; b < len; b += 256) dst[b] = src[b];
}

// Variant that moves each 64 KB tile through shared memory with the bulk-copy engine (four 16 KB stages, all loads in
// flight at once, each forwarded as it lands): the engine writes host memory in larger PCIe transactions than the
// 32-byte sectors SM stores produce.
constexpr uint32_t kBulkChunk = 16 * 1024, kBulkStages = kCopyTile / kBulkChunk;

__global__ void __launch_bounds__(32) copy_items_bulk_kernel(const CopyItem *__restrict__ items)
{
    extern __shared__ __align__(128) uint8_t bulk_smem[];
    const CopyItem it = items[blockIdx.x];
    const uint64_t t0 = (uint64_t)blockIdx.y * kCopyTile;
    if (t0 >= it.bytes || threadIdx.x != 0) return;
    const uint32_t len = (uint32_t)min((uint64_t)kCopyTile, it.bytes - t0);
    const uint8_t *src = static_cast<const uint8_t *>(it.src) + t0;
    uint8_t *dst = static_cast<uint8_t *>(it.dst) + t0;
    const uint32_t body = len & ~15u;
    const uint32_t sm = smem_u32(bulk_smem), bars = sm + kCopyTile;
    const uint32_t chunks = (body + kBulkChunk - 1) / kBulkChunk;
    for (uint32_t c = 0; c < chunks; ++c) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + 8 * c));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (uint32_t c = 0; c < chunks; ++c) {
        const uint32_t bytes = min(kBulkChunk, body - c * kBulkChunk);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bars + 8 * c), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         sm + c * kBulkChunk),
                     "l"(src + (size_t)c * kBulkChunk), "r"(bytes), "r"(bars + 8 * c)
                     : "memory");
    }
    for (uint32_t c = 0; c < chunks; ++c) {
        const uint32_t bytes = min(kBulkChunk, body - c * kBulkChunk);
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done)
                         : "r"(bars + 8 * c)
                         : "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + (size_t)c * kBulkChunk),
                     "r"(sm + c * kBulkChunk), "r"(bytes)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    for (uint32_t b = body; b < len; ++b) dst[b] = src[b];
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

cudaError_t launch_copy_items(const CopyItem *d_items, uint32_t n_items, uint64_t max_bytes, bool to_host, cudaStream_t s)
{
    if (n_items == 0 || max_bytes == 0) return cudaSuccess;
    const uint64_t tiles = (max_bytes + kCopyTile - 1) / kCopyTile;
    if (tiles > 65535) return cudaErrorInvalidConfiguration;
    // Device -> host goes through the bulk-copy engine (measured 35-42 GB/s against 32-37 GB/s for SM stores); host ->
    // device reads reach ~22 GB/s either way and keep the shared-memory-free kernel, which can share an SM with the
    // codec kernels.  QOA_B200_COPY_BULK = 0 / 1 forces one kernel (development aid).
    static const int forced = getenv("QOA_B200_COPY_BULK") ? atoi(getenv("QOA_B200_COPY_BULK")) : -1;
    if (forced >= 0 ? forced != 0 : to_host) {
        const int smem = kCopyTile + 8 * kBulkStages;
        static cudaError_t attr = cudaFuncSetAttribute(copy_items_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (attr != cudaSuccess) return attr;
        copy_items_bulk_kernel<<<dim3(n_items, (unsigned)tiles), 32, smem, s>>>(d_items);
        return cudaGetLastError();
    }
    copy_items_kernel<<<dim3(n_items, (unsigned)tiles), 256, 0, s>>>(d_items);
    return cudaGetLastError();
}

// ---- integer issue-rate microbenchmark -----------------------------------------------------------------------
// 8 independent dependency chains per thread, 16 warps per SM x 4 CTAs: enough to saturate every scheduler.
// kind 0 = IMAD, 1 = LOP3, 2 = IMAD + LOP3 1:1 (the two-pipe issue peak).  Kinds >= 3 time single SASS opcodes the
// kernels lean on (development aid): 3 I2IP.SAT, 4 VIMNMX, 5 SHF, 6 PRMT, 7 LEA.HI, 8 IMAD.HI, 9 IADD3, 10 SEL-free
// ISETP+SEL pair, 11 IMAD.WIDE, 12 IMAD imm, 13 IMUL, 14 LOP3 imm, 15 SHF one-register, 16 OR imm, 17 IADD3 three-register,
// 18 IMAD + one-register SHF 1:1, 19 IMAD : LOP3 : SHF 2:1:1.
#define QOA_OP8(STMT)                                                                                              \
    do {                                                                                                           \
        { int &x = x0; STMT; } { int &x = x1; STMT; } { int &x = x2; STMT; } { int &x = x3; STMT; }                \
        { int &x = x4; STMT; } { int &x = x5; STMT; } { int &x = x6; STMT; } { int &x = x7; STMT; }                \
    } while (0);

template <int KIND>
__global__ void __launch_bounds__(512) int_issue_kernel(int *out, int iters, int a, int b)
{
    int x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (KIND == 0) QOA_OP8(asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b)))
            if (KIND == 1) QOA_OP8(asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(a), "r"(b)))
            if (KIND == 2) {
                asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x0) : "r"(a), "r"(b));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x4) : "r"(a), "r"(b));
                asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x1) : "r"(a), "r"(b));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x5) : "r"(a), "r"(b));
                asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x2) : "r"(a), "r"(b));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x6) : "r"(a), "r"(b));
                asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x3) : "r"(a), "r"(b));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x7) : "r"(a), "r"(b));
            }
            if (KIND == 3) QOA_OP8(asm volatile("cvt.pack.sat.s16.s32 %0, %0, %1;" : "+r"(x) : "r"(a)))
            if (KIND == 4) { // alternate min / max so that ptxas cannot fold min(min(x, a), a)
                if (u & 1)
                    QOA_OP8(asm volatile("min.s32 %0, %0, %1;" : "+r"(x) : "r"(b)))
                else
                    QOA_OP8(asm volatile("max.s32 %0, %0, %1;" : "+r"(x) : "r"(a)))
            }
            if (KIND == 5) QOA_OP8(asm volatile("shf.r.clamp.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b)))
            if (KIND == 6) QOA_OP8(asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b)))
            if (KIND == 7) QOA_OP8(asm volatile("{ .reg .s32 t; shr.s32 t, %0, 13; add.s32 %0, t, %1; }" : "+r"(x) : "r"(a)))
            if (KIND == 8) QOA_OP8(asm volatile("mad.hi.s32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b)))
            if (KIND == 9) QOA_OP8(asm volatile("add.s32 %0, %0, %1;" : "+r"(x) : "r"(a)))
            if (KIND == 10) QOA_OP8(asm volatile("{ .reg .pred p; setp.lt.s32 p, %0, %1; selp.s32 %0, %2, %0, p; }" : "+r"(x) : "r"(a), "r"(b)))
            if (KIND == 12) QOA_OP8(asm volatile("mad.lo.s32 %0, %0, 5, %1;" : "+r"(x) : "r"(a)))
            if (KIND == 13) QOA_OP8(asm volatile("mul.lo.s32 %0, %0, %1;" : "+r"(x) : "r"(a)))
            if (KIND == 14) QOA_OP8(asm volatile("lop3.b32 %0, %0, %1, 0x55, 0x96;" : "+r"(x) : "r"(a)))
            if (KIND == 15) QOA_OP8(asm volatile("shr.s32 %0, %0, 1;" : "+r"(x)))
            if (KIND == 16) QOA_OP8(asm volatile("or.b32 %0, %0, 1;" : "+r"(x)))
            if (KIND == 17) QOA_OP8(asm volatile("{ .reg .s32 t; add.s32 t, %0, %1; add.s32 %0, t, %2; }" : "+r"(x) : "r"(a), "r"(b)))
            if (KIND == 18) {
                asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x0) : "r"(a), "r"(b));
                asm volatile("shr.s32 %0, %0, 1;" : "+r"(x4));
                asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x1) : "r"(a), "r"(b));
                asm volatile("shr.s32 %0, %0, 1;" : "+r"(x5));
                asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x2) : "r"(a), "r"(b));
                asm volatile("shr.s32 %0, %0, 1;" : "+r"(x6));
                asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x3) : "r"(a), "r"(b));
                asm volatile("shr.s32 %0, %0, 1;" : "+r"(x7));
            }
            if (KIND == 19) { // 2 IMAD : 1 LOP3 : 1 one-register SHF
                asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x0) : "r"(a), "r"(b));
                asm volatile("shr.s32 %0, %0, 1;" : "+r"(x4));
                asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x1) : "r"(a), "r"(b));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x5) : "r"(a), "r"(b));
                asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x2) : "r"(a), "r"(b));
                asm volatile("shr.s32 %0, %0, 1;" : "+r"(x6));
                asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x3) : "r"(a), "r"(b));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x7) : "r"(a), "r"(b));
            }
            if (KIND == 20 || KIND == 21) { // decode-like mixes over 17 instructions: 20 = 5 IMAD : 4 FFMA : 8 LOP3, 21 = 9 IMAD : 8 LOP3
                float f0 = __int_as_float(x0), f1 = __int_as_float(x1), f2 = __int_as_float(x2), f3 = __int_as_float(x3);
                const float fa = __int_as_float(a), fb = __int_as_float(b);
#define QOA_I(x) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b))
#define QOA_L(x) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(a), "r"(b))
#define QOA_F(x) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(fa), "f"(fb))
                if (KIND == 20) {
                    QOA_I(x4); QOA_L(x5); QOA_F(f0); QOA_L(x6); QOA_I(x7); QOA_L(x5); QOA_F(f1); QOA_L(x6); QOA_I(x4);
                    QOA_L(x5); QOA_F(f2); QOA_L(x6); QOA_I(x7); QOA_L(x5); QOA_F(f3); QOA_L(x6); QOA_I(x4);
                } else {
                    QOA_I(x4); QOA_L(x5); QOA_I(x0); QOA_L(x6); QOA_I(x7); QOA_L(x5); QOA_I(x1); QOA_L(x6); QOA_I(x4);
                    QOA_L(x5); QOA_I(x2); QOA_L(x6); QOA_I(x7); QOA_L(x5); QOA_I(x3); QOA_L(x6); QOA_I(x4);
                }
#undef QOA_I
#undef QOA_L
#undef QOA_F
                if (KIND == 20) { x0 = __float_as_int(f0); x1 = __float_as_int(f1); x2 = __float_as_int(f2); x3 = __float_as_int(f3); }
            }
            if (KIND == 11) {
                long long y = x0;
                asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(y) : "r"(x1), "r"(a));
                asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(y) : "r"(x2), "r"(a));
                asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(y) : "r"(x3), "r"(a));
                asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(y) : "r"(x4), "r"(a));
                asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(y) : "r"(x5), "r"(a));
                asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(y) : "r"(x6), "r"(a));
                asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(y) : "r"(x7), "r"(a));
                asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(y) : "r"(x1), "r"(b));
                x0 = (int)y ^ (int)(y >> 32);
            }
        }
    }
    const int r = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
    if (r == 0x7fffffff) out[0] = r; // practically never; keeps the chains alive
}

template <int KIND>
static void launch_issue(int grid, int block, int *d_out, int iters)
{
    int_issue_kernel<KIND><<<grid, block>>>(d_out, iters, 3, 7);
}

cudaError_t run_int_issue_peak(int kind, int sm_count, double *thread_ops_per_s)
{
    int *d_out = nullptr;
    cudaError_t e = cudaMalloc(&d_out, sizeof(int));
    if (e != cudaSuccess) return e;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    const int iters = 4096, grid = sm_count * 4, block = 512;
    float best_ms = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(t0);
        switch (kind) {
        case 0: launch_issue<0>(grid, block, d_out, iters); break;
        case 1: launch_issue<1>(grid, block, d_out, iters); break;
        case 2: launch_issue<2>(grid, block, d_out, iters); break;
        case 3: launch_issue<3>(grid, block, d_out, iters); break;
        case 4: launch_issue<4>(grid, block, d_out, iters); break;
        case 5: launch_issue<5>(grid, block, d_out, iters); break;
        case 6: launch_issue<6>(grid, block, d_out, iters); break;
        case 7: launch_issue<7>(grid, block, d_out, iters); break;
        case 8: launch_issue<8>(grid, block, d_out, iters); break;
        case 9: launch_issue<9>(grid, block, d_out, iters); break;
        case 10: launch_issue<10>(grid, block, d_out, iters); break;
        case 11: launch_issue<11>(grid, block, d_out, iters); break;
        case 12: launch_issue<12>(grid, block, d_out, iters); break;
        case 13: launch_issue<13>(grid, block, d_out, iters); break;
        case 14: launch_issue<14>(grid, block, d_out, iters); break;
        case 15: launch_issue<15>(grid, block, d_out, iters); break;
        case 16: launch_issue<16>(grid, block, d_out, iters); break;
        case 17: launch_issue<17>(grid, block, d_out, iters); break;
        case 18: launch_issue<18>(grid, block, d_out, iters); break;
        case 19: launch_issue<19>(grid, block, d_out, iters); break;
        case 20: launch_issue<20>(grid, block, d_out, iters); break;
        default: launch_issue<21>(grid, block, d_out, iters); break;
        }
        cudaEventRecord(t1);
        e = cudaEventSynchronize(t1);
        if (e != cudaSuccess) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, t0, t1);
        if (rep > 0 && ms < best_ms) best_ms = ms;
    }
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    cudaFree(d_out);
    if (e != cudaSuccess) return e;
    const double ops = (double)grid * block * (double)iters * (kind >= 20 ? 8.0 * 17.0 : 64.0); // 8 unrolled x 8 chains (or the 17-instruction mixes)
    *thread_ops_per_s = ops / (best_ms * 1e-3);
    return cudaGetLastError();
}

} // namespace qoa
