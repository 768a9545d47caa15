/*
 * qoa_b200.h — C ABI of the B200-native batched QOA codec engine (libqoa_b200.so).
 *
 * This is the drop-in boundary for the hot path of rafaelcaricio/qoaudio (a pure-Rust crate with no FFI of its
 * own): batched `encode_all` / `QoaEncoder::{write_header, encode_frame}` and batched `decode_all` /
 * `QoaDecoder` iteration / `QoaDecoder::{new_streaming, decode_frame}`.  Every entry point cites the reference
 * item (src/lib.rs:line) it replaces.  A Rust `-sys` shim binding exactly these symbols is shown in
 * INTEGRATION.md and rust/.
 *
 * Conventions
 *   - plain pointers and sizes only; no CUDA or torch types (streams are passed as `void*` = cudaStream_t).
 *   - the function return value is the CALL status (0 = ok, <0 = argument/CUDA failure, see QOA_B200_E_*);
 *     per-stream results mirror the reference's error enums 1:1 in `int32_t status[]`
 *     (a failing stream never aborts the batch).
 *   - "samples" counts are per channel where the reference says so (QoaDesc.samples, frame headers) and
 *     interleaved i16 VALUES for buffer lengths, exactly like the reference's `&[i16]` / `Vec<i16>`.
 *   - `*_dev` variants take DEVICE pointers for the payload (PCM / .qoa bytes) and HOST pointers for the small
 *     per-stream tables; they enqueue on the given stream and only synchronise where documented.
 *   - there is NO CPU fallback: every compute entry point fails with QOA_B200_E_CUDA if no sm_100 device is
 *     usable.
 */
#ifndef QOA_B200_H
#define QOA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/lib.rs:12-19 */
#define QOA_B200_SLICE_LEN 20
#define QOA_B200_LMS_LEN 4
#define QOA_B200_HEADER_SIZE 8
#define QOA_B200_MAGIC 0x716f6166u /* "qoaf" */
#define QOA_B200_SLICES_PER_FRAME 256
#define QOA_B200_FRAME_LEN 5120 /* QOA_FRAME_LEN = 256 * 20 */
#define QOA_B200_MAX_CHANNELS 8 /* encoder; the decoder accepts up to 255 like the reference (u8) */

/* Per-stream status.  EncodeError src/lib.rs:105-115, DecodeError src/lib.rs:866-873. */
enum {
    QOA_B200_OK = 0,
    QOA_B200_ENC_INVALID_CHANNELS = 1,
    QOA_B200_ENC_INVALID_SAMPLE_RATE = 2,
    QOA_B200_ENC_INVALID_SAMPLES = 3,
    QOA_B200_ENC_IO_ERROR = 4,         /* output buffer too small (the reference's io::Write failing) */
    QOA_B200_DEC_NOT_QOA_FILE = 11,
    QOA_B200_DEC_NO_SAMPLES = 12,
    QOA_B200_DEC_INVALID_FRAME_HEADER = 13,
    QOA_B200_DEC_INCOMPATIBLE_FRAME = 14,
    QOA_B200_DEC_IO_ERROR = 15,        /* io::ErrorKind::UnexpectedEof (truncated input) */
    QOA_B200_DEC_OUTPUT_TOO_SMALL = 16 /* engine-specific: caller's out_cap < samples decoded; nothing written past cap */
};

/* Call-level status. */
enum {
    QOA_B200_E_ARG = -1,    /* bad argument (null pointer, misaligned device offset, ...) */
    QOA_B200_E_CUDA = -2,   /* CUDA runtime failure or no usable sm_100 device; see qoa_b200_last_error */
    QOA_B200_E_NOMEM = -3
};

typedef struct qoa_b200_ctx qoa_b200_ctx;

/* QoaDesc, src/lib.rs:94-102 (channels is u8 there; wider here so that 0 and >8 can be reported). */
typedef struct {
    uint32_t channels;
    uint32_t sample_rate;
    uint32_t samples; /* per channel */
} qoa_b200_desc;

/* What iterating QoaDecoder (src/lib.rs:661-693) to its end or first error yields, per stream. */
typedef struct {
    int32_t status;          /* 0 on clean end, else the DecodeError that ended iteration */
    int32_t streaming;       /* ProcessingMode::Streaming (file header samples == 0), src/lib.rs:134-144 */
    uint32_t header_samples; /* file header `samples` field */
    uint32_t channels;       /* of the first FrameHeader item (0 for the streaming-mode dummy, :674-677) */
    uint32_t sample_rate;    /* idem */
    uint32_t index_frames;   /* frames held by the plan's index (real frames that yield samples or a header) */
    uint64_t samples;        /* QoaItem::Sample items yielded (interleaved i16 values) */
    uint64_t frames;         /* QoaItem::FrameHeader items yielded */
    uint32_t max_channels;   /* largest channel count of any indexed frame */
    uint32_t flags;          /* QOA_B200_INFO_* */
} qoa_b200_stream_info;
#define QOA_B200_INFO_UNIFORM_CH 1u  /* every indexed frame has `max_channels` channels */
#define QOA_B200_INFO_ALIGNED16 2u   /* every indexed frame starts 16-byte aligned in the stream's PCM output */
#define QOA_B200_INFO_REGULAR 4u     /* conforming layout (all frames but the last hold 5120 samples): indexed in closed form */

/* FrameHeader, src/lib.rs:773-781 */
typedef struct {
    uint32_t num_channels;
    uint32_t sample_rate;
    uint32_t num_samples_per_channel;
} qoa_b200_frame_header;

/* ---- context ------------------------------------------------------------------------------------------- */

/* One context per GPU (one process per GPU at scale; batches are sharded by stream, no collective). */
int qoa_b200_ctx_create(int device, qoa_b200_ctx **ctx);
void qoa_b200_ctx_destroy(qoa_b200_ctx *ctx);
/* Message of the last failing call on this context (or of ctx_create when ctx == NULL). */
const char *qoa_b200_last_error(const qoa_b200_ctx *ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t qoa_b200_kernel_launches(const qoa_b200_ctx *ctx);
const char *qoa_b200_version(void);

/* ---- encode ---------------------------------------------------------------------------------------------- */

/* Size formula of QoaEncoder::encode, src/lib.rs:374-378. */
size_t qoa_b200_encoded_size(uint32_t channels, uint32_t samples);
/* qoa_frame_size, src/lib.rs:602-604 (not truncated to u16). */
size_t qoa_b200_frame_size(uint32_t channels, uint32_t frame_samples);

/* Batched encode_all (src/lib.rs:760-763 = QoaEncoder::new :335 + encode :367), HOST buffers.
 * pcm[i] holds pcm_len[i] interleaved i16; must equal descs[i].samples * channels else INVALID_SAMPLES (:368).
 * out[i] receives out_len[i] = qoa_b200_encoded_size(...) bytes (ENC_IO_ERROR if out_cap[i] is smaller). */
int qoa_b200_encode_all_batch(qoa_b200_ctx *ctx, const qoa_b200_desc *descs, const int16_t *const *pcm,
                              const size_t *pcm_len, uint8_t *const *out, const size_t *out_cap, size_t *out_len,
                              int32_t *status, size_t n);

/* Same, payload resident on the device.  d_pcm/d_out are device arenas (d_out 8-byte aligned, else QOA_B200_E_ARG);
 * pcm_off[i] (in i16 elements, even) and out_off[i] (in bytes, multiple of 8) are HOST arrays locating stream i.
 * Validation as above (pcm_len/out_cap are HOST arrays).  Enqueues on `stream` (cudaStream_t) and returns without
 * synchronising. */
int qoa_b200_encode_all_batch_dev(qoa_b200_ctx *ctx, const qoa_b200_desc *descs, const int16_t *d_pcm,
                                  const uint64_t *pcm_off, const size_t *pcm_len, uint8_t *d_out,
                                  const uint64_t *out_off, const size_t *out_cap, size_t *out_len, int32_t *status,
                                  size_t n, void *stream);

/* Streaming encoder for a batch of streams: QoaEncoder::{new :335, write_header :404, encode_frame :418}.
 * The LMS state of every (stream, channel) persists on the device between calls (src/lib.rs:89, :482). */
typedef struct qoa_b200_encoder_batch qoa_b200_encoder_batch;
/* status[i] receives the QoaEncoder::new result for stream i; invalid streams are skipped by later calls. */
int qoa_b200_encoder_batch_new(qoa_b200_ctx *ctx, const qoa_b200_desc *descs, size_t n, int32_t *status,
                               qoa_b200_encoder_batch **enc);
void qoa_b200_encoder_batch_free(qoa_b200_encoder_batch *enc);
/* write_header: 8 bytes ("qoaf" + samples BE) into out[i], HOST. */
int qoa_b200_encoder_batch_write_header(const qoa_b200_encoder_batch *enc, uint8_t *const *out);
/* encode_frame for every stream, HOST buffers: pcm[i] = pcm_len[i] interleaved i16 of ONE frame
 * (non-empty, multiple of channels, <= 5120 per channel else INVALID_SAMPLES, :424-430); written[i] bytes go to
 * out[i].  A stream with pcm_len[i] == 0 and pcm[i] == NULL is skipped (status untouched, written 0). */
int qoa_b200_encoder_batch_encode_frame(qoa_b200_encoder_batch *enc, const int16_t *const *pcm,
                                        const size_t *pcm_len, uint8_t *const *out, const size_t *out_cap,
                                        size_t *written, int32_t *status);
/* Same with any number of samples per stream (a multiple of channels): cut into frames of 5120 — the bytes that
 * many encode_frame calls fed 5120-sample chunks would produce, i.e. the frame part of QoaEncoder::encode (:388-394)
 * starting from the encoder's current LMS state.  written[i] = qoa_b200_encoded_size(ch, samples) - 8. */
int qoa_b200_encoder_batch_encode_chunk(qoa_b200_encoder_batch *enc, const int16_t *const *pcm,
                                        const size_t *pcm_len, uint8_t *const *out, const size_t *out_cap,
                                        size_t *written, int32_t *status);
/* Device-resident chunk: for stream i encode chunk_samples[i] samples per channel (any number; cut into frames
 * of 5120, i.e. the same bytes as that many encode_frame calls fed 5120-sample chunks) from d_pcm + pcm_off[i]
 * to d_out + out_off[i] (frames back to back, no file header).  chunk_samples[i] == 0 skips the stream.
 * Enqueues on `stream`, no synchronisation. */
int qoa_b200_encoder_batch_encode_chunk_dev(qoa_b200_encoder_batch *enc, const int16_t *d_pcm,
                                            const uint64_t *pcm_off, const uint32_t *chunk_samples, uint8_t *d_out,
                                            const uint64_t *out_off, void *stream);
/* Copies the current LMS state of stream i, channel c to host (history[4], weights[4]); synchronises. */
int qoa_b200_encoder_batch_get_lms(qoa_b200_encoder_batch *enc, size_t i, uint32_t c, int32_t history[4],
                                   int32_t weights[4]);

/* ---- decode ------------------------------------------------------------------------------------------------ */

/* Decode semantics selector. */
enum {
    QOA_B200_DECODE_ITER = 0,     /* QoaDecoder::new + iterate (src/lib.rs:127-169, :661-693) */
    QOA_B200_DECODE_ALL = 1,      /* decode_all (src/lib.rs:713-747): adds its IncompatibleFrame check */
    QOA_B200_DECODE_FRAMES = 2    /* new_streaming + decode_frame (src/lib.rs:627-652): input = frames, no file header */
};

/* A decode plan = the frame index of a batch (the pre-pass that walks every stream's frame headers exactly like
 * decode_frame_header_and_lms, src/lib.rs:205-289, and validates them) plus the device copy of the input when it
 * came from host memory. */
typedef struct qoa_b200_decode_plan qoa_b200_decode_plan;

/* HOST input: copies the n streams to the device, indexes them.  Synchronises. */
int qoa_b200_decode_plan_create(qoa_b200_ctx *ctx, const uint8_t *const *qoa, const size_t *len, size_t n,
                                int mode, qoa_b200_decode_plan **plan);
/* DEVICE input: stream i occupies d_qoa[in_off[i] .. in_off[i]+len[i]) (d_qoa 16-byte aligned and in_off a multiple
 * of 16, else QOA_B200_E_ARG; HOST arrays).  The decode kernels read the arena in 16-byte vectors: the bytes up to the
 * next 16-byte boundary behind a stream's end must be readable (any content; they are never interpreted) — true for
 * every stream but possibly the last of an arena that ends exactly at in_off + len with len % 16 == 8.
 * Runs the index kernels on `stream` and synchronises it (the per-stream info is needed on the host). */
int qoa_b200_decode_plan_create_dev(qoa_b200_ctx *ctx, const uint8_t *d_qoa, const uint64_t *in_off,
                                    const size_t *len, size_t n, int mode, void *stream,
                                    qoa_b200_decode_plan **plan);
void qoa_b200_decode_plan_destroy(qoa_b200_decode_plan *plan);
/* Per-stream info (n entries) — what the caller needs to size outputs: info[i].samples i16 values. */
int qoa_b200_decode_plan_info(const qoa_b200_decode_plan *plan, qoa_b200_stream_info *info);
/* Frame headers of stream i in iteration order (real frames only); returns how many exist, writes <= cap. */
size_t qoa_b200_decode_plan_frame_headers(const qoa_b200_decode_plan *plan, size_t i, qoa_b200_frame_header *hdrs,
                                          size_t cap);
/* Decode into HOST buffers out[i] (capacity out_cap[i] i16 values).  status[i] = info[i].status, or
 * DEC_OUTPUT_TOO_SMALL (nothing is written for that stream).  Synchronises. */
int qoa_b200_decode_plan_run(qoa_b200_decode_plan *plan, int16_t *const *out, const size_t *out_cap,
                             int32_t *status);
/* Decode into a DEVICE arena: stream i goes to d_out + out_off[i] (i16 elements; HOST array).  The fast kernels need
 * 16-byte aligned stream slots (d_out 16-byte aligned and out_off a multiple of 8); streams whose slot is not take
 * the generic kernel (correct, ~5x slower).  Enqueues on `stream`; no synchronisation. */
int qoa_b200_decode_plan_run_dev(qoa_b200_decode_plan *plan, int16_t *d_out, const uint64_t *out_off,
                                 const size_t *out_cap, int32_t *status, void *stream);

/* Sizes ONE host-resident stream without a GPU: walks its frame headers exactly like the index pre-pass does on the
 * device (QoaDecoder::new + decode_frame_header_and_lms, src/lib.rs:127-169, :205-289) and fills *info (status,
 * channels, sample_rate, info->samples = the i16 values a decode yields, frames, ...).  Header parsing only: no sample is
 * decoded on the CPU.  This is what a caller of the one-call entry points below uses to allocate out[i]. */
int qoa_b200_probe(const uint8_t *qoa, size_t len, int mode, qoa_b200_stream_info *info);

/* One-call convenience for HOST buffers: plan_create + plan_run + destroy.  info may be NULL. */
int qoa_b200_decode_all_batch(qoa_b200_ctx *ctx, const uint8_t *const *qoa, const size_t *len, size_t n, int mode,
                              int16_t *const *out, const size_t *out_cap, qoa_b200_stream_info *info,
                              int32_t *status);

/* ---- several GPUs from one process --------------------------------------------------------------------------- */

/* A multi-device context: one qoa_b200_ctx per listed device (devices == NULL or n <= 0: every visible device).
 * The batch calls below cut the batch into contiguous ranges of streams, one per device, balanced by work
 * (qoa_b200_partition), and run the single-device host path on every range from its own host thread: no collective
 * (the path has no cross-stream dependency, src/lib.rs:59-68, :85-91), results in caller order.  This is the call a
 * user of encode_all (src/lib.rs:760) / decode_all (:713) makes for a batch on an 8-GPU box. */
typedef struct qoa_b200_mctx qoa_b200_mctx;
int qoa_b200_ctx_create_multi(const int *devices, int n, qoa_b200_mctx **mctx);
void qoa_b200_mctx_destroy(qoa_b200_mctx *mctx);
int qoa_b200_mctx_device_count(const qoa_b200_mctx *mctx);
/* CUDA device index / single-device context of slot i (the *_dev calls of a slot's context work as usual). */
int qoa_b200_mctx_device(const qoa_b200_mctx *mctx, int i);
qoa_b200_ctx *qoa_b200_mctx_ctx(qoa_b200_mctx *mctx, int i);
/* Message of the last failing call ("device D: ..."); mctx == NULL: why ctx_create_multi failed on this thread. */
const char *qoa_b200_mctx_last_error(const qoa_b200_mctx *mctx);
uint64_t qoa_b200_mctx_kernel_launches(const qoa_b200_mctx *mctx);
/* Splits range(n) into `parts` contiguous ranges of near-equal summed weight: range p = first[p] .. first[p + 1]
 * (parts + 1 entries, ranges may be empty).  Host arithmetic only.  The batch calls weigh a stream by
 * samples x channels (encode) or by its encoded length (decode). */
int qoa_b200_partition(const uint64_t *weight, size_t n, int parts, size_t *first);
int qoa_b200_mctx_partition(const qoa_b200_mctx *mctx, const uint64_t *weight, size_t n, size_t *first);
/* qoa_b200_encode_all_batch / qoa_b200_decode_all_batch over every device of the context; same arguments, same
 * per-stream status values, same bytes. */
int qoa_b200_mctx_encode_all_batch(qoa_b200_mctx *mctx, const qoa_b200_desc *descs, const int16_t *const *pcm,
                                   const size_t *pcm_len, uint8_t *const *out, const size_t *out_cap, size_t *out_len,
                                   int32_t *status, size_t n);
int qoa_b200_mctx_decode_all_batch(qoa_b200_mctx *mctx, const uint8_t *const *qoa, const size_t *len, size_t n,
                                   int mode, int16_t *const *out, const size_t *out_cap, qoa_b200_stream_info *info,
                                   int32_t *status);

/* ---- synthetic workload generators (bench + tests; same integer recipes as tests/signals.py) ------------- */

/* Fills d_pcm with `n_streams` streams of `channels` x `samples` interleaved i16 starting at sample index
 * `start` of each stream (kind 0 = "music", 1 = white noise).  Stream i is stream id first_stream + i. */
int qoa_b200_gen_pcm_dev(qoa_b200_ctx *ctx, int16_t *d_pcm, uint64_t n_streams, uint64_t first_stream,
                         uint32_t channels, uint64_t samples, uint64_t start, uint64_t seed, int kind, void *stream);
/* Fills d_qoa with n_streams valid .qoa streams of the given shape (each qoa_b200_encoded_size bytes, padded
 * to `stride` bytes): correct headers, pseudo-random LMS states and slice words (decode cost is data-independent). */
int qoa_b200_gen_qoa_dev(qoa_b200_ctx *ctx, uint8_t *d_qoa, uint64_t stride, uint64_t n_streams,
                         uint64_t first_stream, uint32_t channels, uint32_t sample_rate, uint32_t samples,
                         uint64_t seed, void *stream);
/* Integer-pipe issue-rate microbenchmark (the measured INT32 roofline denominator for encode): runs `iters`
 * dependent-free integer ops per thread on every SM and returns the achieved thread-ops/s.
 * kind 0 = IMAD only, 1 = LOP3/IADD3 only, 2 = 1:1 mix.  Synchronises. */
int qoa_b200_int_issue_peak(qoa_b200_ctx *ctx, int kind, double *thread_ops_per_s);

#ifdef __cplusplus
}
#endif
#endif
