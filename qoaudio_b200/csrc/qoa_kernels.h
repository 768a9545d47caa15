// qoa_kernels.h — host-callable launchers of the sm_100a kernels (internal; the public surface is include/qoa_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/qoa_b200.h"
#include "qoa_device.cuh"

namespace qoa {

// ---- decode --------------------------------------------------------------------------------------------------
// Index pre-pass, pass 1: one thread per stream.  With `speculate`, a stream whose length and first frame header match
// the closed-form layout of a conforming encoder is accepted as "regular" (QOA_B200_INFO_REGULAR) without walking;
// every other stream is walked frame header by frame header (src/lib.rs:205-289 semantics).  Fills d_info[i]
// (info.index_frames = number of frames that go into the frame table).
// h_info is a host-mapped (pinned) mirror of d_info that the kernel also writes.
cudaError_t launch_index_count(const uint8_t *d_arena, const uint64_t *d_in_off, const uint64_t *d_in_len, size_t n,
                               int mode, int speculate, qoa_b200_stream_info *d_info, qoa_b200_stream_info *h_info,
                               cudaStream_t s);
// Pass 2 for walked streams: same walk, writes FrameRec entries at d_recs[d_frame_base[i] + f].
cudaError_t launch_index_fill(const uint8_t *d_arena, const uint64_t *d_in_off, const uint64_t *d_in_len, size_t n,
                              int mode, const qoa_b200_stream_info *d_info, const uint64_t *d_frame_base,
                              FrameRec *d_recs, cudaStream_t s);
// Pass 2 for regular streams: one thread per frame verifies the header word and writes the FrameRec; *d_bad is set
// (a host-mapped int) if any header differs from the closed-form expectation (the caller then re-indexes without
// speculation).
cudaError_t launch_index_fill_regular(const uint8_t *d_arena, const uint64_t *d_in_off, size_t n, uint32_t max_frames,
                                      const qoa_b200_stream_info *d_info, const uint64_t *d_frame_base,
                                      FrameRec *d_recs, int *d_bad, cudaStream_t s);
// Frame-parallel decode.  max_ch = largest channel count of any frame in the table; d_out_base[stream] = i16 element
// offset of the stream's output in d_out (UINT64_MAX = do not write this stream); out_bytes = extent of the output arena
// that the batch may touch (upper bound of any written byte offset).
// uniform_ch = 1..8 when every frame has that many channels and starts 16-byte aligned in the output (fast paths:
// TMA-staged kernel for mono / stereo, pair-per-lane kernel for 3..8 channels), else 0 (generic path).
cudaError_t launch_decode(const uint8_t *d_arena, const FrameRec *d_recs, uint64_t n_frames, uint32_t max_ch,
                          uint32_t uniform_ch, const uint64_t *d_out_base, int16_t *d_out, uint64_t out_bytes,
                          cudaStream_t s, int *n_launches);

// The index walk of ONE stream on the host (frame headers only): what qoa_b200_probe returns.
void probe_stream(const uint8_t *qoa, uint64_t len, int mode, qoa_b200_stream_info *info);

// ---- encode --------------------------------------------------------------------------------------------------
// Per-stream encode job (device table).
struct EncJob {
    uint64_t pcm_off;   // i16 element offset of the stream's first sample of this call in the PCM arena
    uint64_t out_off;   // byte offset of the first frame header this call writes in the output arena
    uint32_t channels;  // 1..8
    uint32_t rate;
    uint32_t nsamples;  // samples per channel to encode in this call (0 = skip)
    uint32_t lms_base;  // index of the stream's channel-0 LMS state in the state array (state = 8 x i32 per chain)
    uint32_t file_samples; // value of the file header's `samples` field (used when flags & 1)
    uint32_t flags;        // bit 0: write the 8-byte file header at out_off; frames then start at out_off + 8
};
// chain table entry: (job index, channel)
struct EncChain {
    uint32_t job;
    uint32_t channel;
};
cudaError_t launch_encode(const int16_t *d_pcm, uint8_t *d_out, const EncJob *d_jobs, const EncChain *d_chains,
                          uint32_t n_chains, int32_t *d_lms, cudaStream_t s, int *n_launches);

// ---- batched copies -------------------------------------------------------------------------------------------
// One kernel copies n_items buffers (src/dst 16-byte aligned, device-accessible: device memory or pinned host memory).
struct CopyItem {
    const void *src;
    void *dst;
    uint64_t bytes;
};
cudaError_t launch_copy_items(const CopyItem *d_items, uint32_t n_items, uint64_t max_bytes, bool to_host, cudaStream_t s);

// ---- generators / microbenchmarks ----------------------------------------------------------------------------
cudaError_t launch_gen_pcm(int16_t *d_pcm, uint64_t n_streams, uint64_t first_stream, uint32_t channels,
                           uint64_t samples, uint64_t start, uint64_t seed, int kind, cudaStream_t s);
cudaError_t launch_gen_qoa(uint8_t *d_qoa, uint64_t stride, uint64_t n_streams, uint64_t first_stream,
                           uint32_t channels, uint32_t rate, uint32_t samples, uint64_t seed, cudaStream_t s);
cudaError_t run_int_issue_peak(int kind, int sm_count, double *thread_ops_per_s);

} // namespace qoa
