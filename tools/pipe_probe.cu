// pipe_probe.cu — per-pipe and mixed instruction throughput on the current GPU (development aid, not product code).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_probe tools/pipe_probe.cu
// Every kind runs 8 independent dependency chains per thread, 512 threads x 4 CTAs per SM, and reports warp-level
// thread-ops per second; "mix" kinds interleave opcodes in the ratio of the hybrid stereo decode loop.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define I_(x) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b))
#define L_(x) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(a), "r"(b))
#define D_(x) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(da), "d"(db))
#define A_(x) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x) : "d"(db))
#define F_(x) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(fa), "f"(fb))

template <int KIND>
__global__ void __launch_bounds__(512) probe(int *out, int iters, int a, int b, double da, double db, float fa, float fb)
{
    int x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    double d0 = x0, d1 = x1, d2 = x2, d3 = x3, d4 = x4, d5 = x5, d6 = x6, d7 = x7;
    float f0 = x0, f1 = x1, f2 = x2, f3 = x3;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (KIND == 0) { I_(x0); I_(x1); I_(x2); I_(x3); I_(x4); I_(x5); I_(x6); I_(x7); }              // 8 IMAD
            if (KIND == 1) { L_(x0); L_(x1); L_(x2); L_(x3); L_(x4); L_(x5); L_(x6); L_(x7); }              // 8 LOP3
            if (KIND == 2) { D_(d0); D_(d1); D_(d2); D_(d3); D_(d4); D_(d5); D_(d6); D_(d7); }              // 8 DFMA
            if (KIND == 3) { A_(d0); A_(d1); A_(d2); A_(d3); A_(d4); A_(d5); A_(d6); A_(d7); }              // 8 DADD
            if (KIND == 4) { I_(x0); D_(d0); I_(x1); D_(d1); I_(x2); D_(d2); I_(x3); D_(d3); }              // IMAD:DFMA 1:1
            if (KIND == 5) { L_(x0); D_(d0); L_(x1); D_(d1); L_(x2); D_(d2); L_(x3); D_(d3); }              // LOP3:DFMA 1:1
            if (KIND == 6) { I_(x0); L_(x4); D_(d0); I_(x1); L_(x5); D_(d1); I_(x2); L_(x6); D_(d2); }      // 1:1:1
            if (KIND == 7) { // hybrid decode loop mix per sample pair: 8 IMAD : 9 FP64 : 12 ALU
                I_(x0); D_(d0); L_(x4); I_(x1); D_(d1); L_(x5); I_(x2); D_(d2); L_(x6); I_(x3); D_(d3); L_(x7);
                I_(x0); D_(d4); L_(x4); I_(x1); D_(d5); L_(x5); I_(x2); D_(d6); L_(x6); I_(x3); D_(d7); L_(x7);
                A_(d0); L_(x4); L_(x5); L_(x6); L_(x7);
            }
            if (KIND == 8) { // the all-integer stereo loop mix: 16 IMAD : 13 ALU
                I_(x0); L_(x4); I_(x1); L_(x5); I_(x2); L_(x6); I_(x3); L_(x7); I_(x0); L_(x4); I_(x1); L_(x5); I_(x2);
                L_(x6); I_(x3); L_(x7); I_(x0); L_(x4); I_(x1); L_(x5); I_(x2); L_(x6); I_(x3); L_(x7); I_(x0); L_(x4);
                I_(x1); I_(x2); I_(x3);
            }
            if (KIND == 9) { F_(f0); F_(f1); F_(f2); F_(f3); F_(f0); F_(f1); F_(f2); F_(f3); }              // 8 FFMA
            if (KIND == 10) { I_(x0); F_(f0); I_(x1); F_(f1); I_(x2); F_(f2); I_(x3); F_(f3); }             // IMAD:FFMA 1:1
            if (KIND == 11) { I_(x0); F_(f0); D_(d0); L_(x4); I_(x1); F_(f1); D_(d1); L_(x5); }             // 1:1:1:1
        }
    }
    int r = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
    double d = d0 + d1 + d2 + d3 + d4 + d5 + d6 + d7;
    float f = f0 + f1 + f2 + f3;
    if (iters < 0) out[threadIdx.x] = r + (int)d + (int)f; // never taken: keeps every chain live
}

static const int kOps[] = {8, 8, 8, 8, 8, 8, 9, 29, 29, 8, 8, 8};
static const char *kNames[] = {"IMAD", "LOP3", "DFMA", "DADD", "IMAD:DFMA 1:1", "LOP3:DFMA 1:1", "IMAD:LOP3:DFMA 1:1:1",
                               "hybrid mix 8 IMAD:9 FP64:12 ALU", "integer mix 16 IMAD:13 ALU", "FFMA", "IMAD:FFMA 1:1",
                               "IMAD:FFMA:DFMA:LOP3 1:1:1:1"};

template <int KIND>
static void run(int sms, int *d_out)
{
    const int iters = 2048, grid = sms * 4, block = 512;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(t0);
        probe<KIND><<<grid, block>>>(d_out, iters, 3, 7, 1.0000001, 0.5, 1.0001f, 0.5f);
        cudaEventRecord(t1);
        if (cudaEventSynchronize(t1) != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("kind %d: launch failed\n", KIND); return; }
        float ms;
        cudaEventElapsedTime(&ms, t0, t1);
        if (rep && ms < best) best = ms;
    }
    const double ops = (double)grid * block * iters * 4.0 * kOps[KIND];
    printf("%2d %-34s %7.2f T thread-ops/s  (%.3f ms)\n", KIND, kNames[KIND], ops / (best * 1e-3) / 1e12, best);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int *d_out;
    cudaMalloc(&d_out, 4);
    printf("%s, %d SMs, %d MHz: nominal 1 warp-instr/clk/SMSP = %.2f T thread-ops/s\n", p.name, p.multiProcessorCount,
           p.clockRate / 1000, p.multiProcessorCount * 128.0 * p.clockRate * 1e3 / 1e12);
    run<0>(p.multiProcessorCount, d_out); run<1>(p.multiProcessorCount, d_out); run<2>(p.multiProcessorCount, d_out);
    run<3>(p.multiProcessorCount, d_out); run<4>(p.multiProcessorCount, d_out); run<5>(p.multiProcessorCount, d_out);
    run<6>(p.multiProcessorCount, d_out); run<7>(p.multiProcessorCount, d_out); run<8>(p.multiProcessorCount, d_out);
    run<9>(p.multiProcessorCount, d_out); run<10>(p.multiProcessorCount, d_out); run<11>(p.multiProcessorCount, d_out);
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : 1;
}
