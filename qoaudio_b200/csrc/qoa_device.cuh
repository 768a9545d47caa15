// qoa_device.cuh — device-side building blocks shared by the decode / encode / generator kernels.
//
// Constants and arithmetic follow the reference crate (paths relative to its repo root):
//   tables            src/lib.rs:22, :25-27, :831-864
//   QoaLms            src/lib.rs:70-74, :797-829
//   qoa_div           src/lib.rs:613-617
// All arithmetic is two's-complement i32 exactly like the crate's wrapping_* ops (SURVEY.md §8a notes).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qoa {

constexpr int kSliceLen = 20;          // QOA_SLICE_LEN        src/lib.rs:12
constexpr int kFrameLen = 5120;        // QOA_FRAME_LEN        src/lib.rs:17
constexpr int kSlicesPerFrame = 256;   // QOA_SLICES_PER_FRAME src/lib.rs:16
constexpr uint32_t kMagic = 0x716f6166u;

// QOA_DEQUANT_TAB src/lib.rs:847-864, flattened [sf*8 + code].
static __constant__ int c_dequant[128] = {
    1,    -1,    3,    -3,    5,    -5,    7,     -7,     5,    -5,    18,   -18,   32,   -32,   49,    -49,
    16,   -16,   53,   -53,   95,   -95,   147,   -147,   34,   -34,   113,  -113,  203,  -203,  315,   -315,
    63,   -63,   210,  -210,  378,  -378,  588,   -588,   104,  -104,  345,  -345,  621,  -621,  966,   -966,
    158,  -158,  528,  -528,  950,  -950,  1477,  -1477,  228,  -228,  760,  -760,  1368, -1368, 2128,  -2128,
    316,  -316,  1053, -1053, 1895, -1895, 2947,  -2947,  422,  -422,  1405, -1405, 2529, -2529, 3934,  -3934,
    548,  -548,  1828, -1828, 3290, -3290, 5117,  -5117,  696,  -696,  2320, -2320, 4176, -4176, 6496,  -6496,
    868,  -868,  2893, -2893, 5207, -5207, 8099,  -8099,  1064, -1064, 3548, -3548, 6386, -6386, 9933,  -9933,
    1286, -1286, 4288, -4288, 7718, -7718, 12005, -12005, 1536, -1536, 5120, -5120, 9216, -9216, 14336, -14336};
// QOA_RECIPROCAL_TAB src/lib.rs:25-27
static __constant__ int c_recip[16] = {65536, 9363, 3121, 1457, 781, 475, 311, 216, 156, 117, 90, 71, 57, 47, 39, 32};
// QOA_QUANT_TAB src/lib.rs:22, index = clamp(scaled, -8, 8) + 8
static __constant__ int c_quant[17] = {7, 7, 7, 5, 5, 3, 3, 1, 0, 0, 2, 2, 4, 4, 6, 6, 6};

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

// Big-endian 64-bit word as (hi, lo) host-order halves; p must be 8-byte aligned.
__device__ __forceinline__ uint2 ldg_be64(const uint8_t *p)
{
    uint2 v = __ldg(reinterpret_cast<const uint2 *>(p));
    return make_uint2(bswap32(v.x), bswap32(v.y)); // .x = high 32 bits, .y = low 32 bits
}
__device__ __forceinline__ void st_be64(uint8_t *p, uint32_t hi, uint32_t lo)
{
    *reinterpret_cast<uint2 *>(p) = make_uint2(bswap32(hi), bswap32(lo));
}

__device__ __forceinline__ int clamp_i16(int v) { return max(-32768, min(32767, v)); }

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// splitmix64 finaliser — the synthetic-data recipe shared with tests/signals.py.
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t stream_key(uint64_t seed, uint64_t stream, uint64_t channel)
{
    return splitmix64(seed * 0x100000001B3ull + stream * 257ull + channel);
}

// One frame of the decode index (32 bytes).  Built by the index pre-pass, consumed by the decode kernel.
struct FrameRec {
    uint64_t in_off;   // byte offset of the frame header inside the input arena
    uint64_t out_rel;  // i16 elements from the start of the stream's output to this frame's first sample
    uint32_t stream;   // index of the stream in the batch
    uint32_t rate;     // frame header sample rate
    uint16_t nsamp;    // samples per channel this frame EMITS (== nspc unless the input is truncated mid-frame)
    uint16_t nspc;     // frame header num_samples_per_channel
    uint8_t ch;        // frame header num_channels (1..255)
    uint8_t pad[3];
};
static_assert(sizeof(FrameRec) == 32, "FrameRec layout");

} // namespace qoa
