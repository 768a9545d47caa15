// core_probe.cu — how fast can the stereo decode inner loop run when nothing else is in the way?  (development aid)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I qoaudio_b200/csrc -o tools/core_probe tools/core_probe.cu -lcuda
// Kinds: 0 = pair_core only (tables replaced by kernel arguments): 16 IMAD + 9 ALU per sample pair
//        1 = + single-code lookups from the shared-memory table (SHF + LOP3 + LDS.64 per sample and channel)
//        4..7 = pieces of the lookup in isolation (see ProbeLoop); 8 = two frames (four chains) per thread.  (Round 2 also measured the since-removed two-code
//        table here: 22.4 ms, profiles/r02_decode_experiments.md.)
// Each thread decodes `slices` pseudo slices (the slice words come from a counter hash, no global loads), 20 steps each,
// and folds the output words into a checksum.  Reported: ns per sample pair per SM-scheduler and the implied time of the
// headline batch (8192 x 60 s stereo = 21.68 G sample pairs) on 148 SMs.
#include "../qoaudio_b200/csrc/qoa_decode.cu"
#include <cstdio>
using namespace qoa;

__device__ __forceinline__ int lds_s32(uint32_t addr)
{
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// 4: LDS.32 of dq, delta = dq >> 4 by SHF.  5: address arithmetic only (no load).  6: loads only (fixed address).
// 7: LDS.32 of dq only, delta from the kernel argument (no shift)
template <int I, int K>
struct ProbeLoop {
    static __device__ __forceinline__ void run(Chain &a, Chain &b, uint32_t row_a, uint32_t row_b, uint2 wa, uint2 wb,
                                               uint32_t *o, int2 ka, int2 kb)
    {
        int2 ta, tb;
        if (K == 4) {
            ta.x = lds_s32(row_a | code_off<I>(wa.x, wa.y)); ta.y = ta.x >> 4;
            tb.x = lds_s32(row_b | code_off<I>(wb.x, wb.y)); tb.y = tb.x >> 4;
        } else if (K == 5) {
            ta = make_int2((int)(row_a | code_off<I>(wa.x, wa.y)), ka.y);
            tb = make_int2((int)(row_b | code_off<I>(wb.x, wb.y)), kb.y);
        } else if (K == 6) {
            ta = lds_v2(row_a + 128 * (I & 7));
            tb = lds_v2(row_b + 128 * (I & 7));
        } else {
            ta = make_int2(lds_s32(row_a | code_off<I>(wa.x, wa.y)), ka.y);
            tb = make_int2(lds_s32(row_b | code_off<I>(wb.x, wb.y)), kb.y);
        }
        o[I] = pair_core(a, b, ta, tb);
        ProbeLoop<I + 1, K>::run(a, b, row_a, row_b, wa, wb, o, ka, kb);
    }
};
template <int K>
struct ProbeLoop<20, K> {
    static __device__ __forceinline__ void run(Chain &, Chain &, uint32_t, uint32_t, uint2, uint2, uint32_t *, int2, int2) {}
};

template <int I>
struct Pair2Loop { // two independent (L, R) pairs per thread, interleaved step by step
    static __device__ __forceinline__ void run(Chain &a, Chain &b, Chain &c, Chain &d, uint32_t row_a, uint32_t row_b,
                                               uint32_t row_c, uint32_t row_d, uint2 wa, uint2 wb, uint2 wc, uint2 wd,
                                               uint32_t *o, uint32_t *o2)
    {
        o[I] = decode_pair_step<I>(a, b, row_a, row_b, wa, wb);
        o2[I] = decode_pair_step<I>(c, d, row_c, row_d, wc, wd);
        Pair2Loop<I + 1>::run(a, b, c, d, row_a, row_b, row_c, row_d, wa, wb, wc, wd, o, o2);
    }
};
template <>
struct Pair2Loop<20> {
    static __device__ __forceinline__ void run(Chain &, Chain &, Chain &, Chain &, uint32_t, uint32_t, uint32_t, uint32_t, uint2,
                                               uint2, uint2, uint2, uint32_t *, uint32_t *)
    {
    }
};

template <int KIND, int THREADS, int CTAS>
__global__ void __launch_bounds__(THREADS, CTAS) core_kernel(uint32_t *out, int slices, int2 ka, int2 kb)
{
    extern __shared__ uint8_t smem_raw[];
    constexpr uint32_t align = kLutAlign;
    uint8_t *smem = smem_raw + ((align - (smem_u32(smem_raw) & (align - 1))) & (align - 1));
    if (KIND == 1 || KIND >= 4) fill_lut(smem);
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t tid = blockIdx.x * THREADS + threadIdx.x;
    Chain a{(int)tid, 3, -5, 7, 100, -200, -8192, 16384, 1, 1, -1, 1}, b = a, c2 = a, d2 = a;
    b.h0 = 11;
    c2.h1 = 13;
    d2.h2 = 17;
    uint32_t sum = 0;
    uint32_t lane_row = smem_u32(smem) + (lane & 15) * 8;
    uint2 wa = make_uint2(tid * 2654435761u, tid * 40503u + 17), wb = make_uint2(~wa.x, wa.y ^ 0x5bd1e995u);
#pragma unroll 1
    for (int k = 0; k < slices; ++k) {
        uint32_t o[20];
        if (KIND == 0) {
#pragma unroll
            for (int i = 0; i < 20; ++i) o[i] = pair_core(a, b, ka, kb);
        } else {
            const uint32_t row_a = lane_row + (wa.x >> 28) * 1024;
            const uint32_t row_b = lane_row + (wb.x >> 28) * 1024;
            if (KIND == 1) PairLoop<0>::run(a, b, row_a, row_b, wa, wb, o);
            if (KIND == 8) { // ILP 4: a second, independent pair of chains in the same thread (two frames per thread)
                const uint32_t row_c = lane_row + (wa.y >> 28) * 1024, row_d = lane_row + (wb.y >> 28) * 1024;
                uint32_t o2[20];
                Pair2Loop<0>::run(a, b, c2, d2, row_a, row_b, row_c, row_d, wa, wb, make_uint2(wb.y, wa.x), make_uint2(wa.y, wb.x), o, o2);
#pragma unroll
                for (int i = 0; i < 20; ++i) sum ^= o2[i];
            }
            if (KIND == 4) ProbeLoop<0, 4>::run(a, b, row_a, row_b, wa, wb, o, ka, kb);
            if (KIND == 5) ProbeLoop<0, 5>::run(a, b, row_a, row_b, wa, wb, o, ka, kb);
            if (KIND == 6) ProbeLoop<0, 6>::run(a, b, row_a, row_b, wa, wb, o, ka, kb);
            if (KIND == 7) ProbeLoop<0, 7>::run(a, b, row_a, row_b, wa, wb, o, ka, kb);
            wa.x = wa.x * 1664525u + 1013904223u; wa.y ^= wa.x >> 3;   // next pseudo slice words
            wb.x = wb.x * 22695477u + 1u; wb.y ^= wb.x >> 5;
        }
#pragma unroll
        for (int i = 0; i < 20; ++i) sum ^= o[i];
    }
    if (sum == 0x12345678u) out[0] = sum;
}

template <int KIND, int THREADS, int CTAS>
static void run(int sms, uint32_t *d_out, const char *name)
{
    const int slices = 4096;
    const size_t smem = (KIND >= 1 ? 1024 + kLutBytes : 0) +
                        (CTAS >= 2 ? 60000 : 120000); // plus a tile footprint, so that exactly CTAS CTAs fit on an SM
    cudaFuncSetAttribute(core_kernel<KIND, THREADS, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0); cudaEventCreate(&t1);
    const int ctas = sms * CTAS;
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(t0);
        core_kernel<KIND, THREADS, CTAS><<<ctas, THREADS, smem>>>(d_out, slices, make_int2(37, 2), make_int2(-113, -8));
        cudaEventRecord(t1);
        if (cudaEventSynchronize(t1) != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("%s: launch failed\n", name); return; }
        float ms; cudaEventElapsedTime(&ms, t0, t1);
        if (ms < best) best = ms;
    }
    const double pairs = (double)ctas * THREADS * slices * 20.0 * (KIND == 8 ? 2.0 : 1.0);
    const double c3_pairs = 8192.0 * 2646000.0;
    printf("%-44s %4d thr x %d: %8.3f ms, %.2f T pairs/s -> headline batch in %.2f ms\n", name, THREADS, CTAS, best, pairs / best / 1e9,
           c3_pairs / (pairs / best));
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    uint32_t *d; cudaMalloc(&d, 4);
    const int sms = p.multiProcessorCount;
    run<0, 256, 2>(sms, d, "pair_core only (16 IMAD + 9 ALU)");
    run<1, 256, 2>(sms, d, "+ single-code lookups (4 ALU + 2 LDS.64)");
    run<5, 256, 2>(sms, d, "+ address arithmetic only (4 ALU)");
    run<6, 256, 2>(sms, d, "+ loads only (2 LDS.64, fixed address)");
    run<7, 256, 2>(sms, d, "+ 4 ALU + 2 LDS.32 (dq only)");
    run<4, 256, 2>(sms, d, "+ 4 ALU + 2 LDS.32 + 2 SHF (delta = dq >> 4)");
    run<8, 128, 2>(sms, d, "two frames per thread (ILP 4), lookups");
    run<8, 256, 1>(sms, d, "two frames per thread (ILP 4), lookups");
    run<8, 192, 2>(sms, d, "two frames per thread (ILP 4), lookups");
    return 0;
}
