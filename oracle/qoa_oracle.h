/*
 * qoa_oracle.h — CPU restatement of rafaelcaricio/qoaudio's QOA encode/decode path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle for the B200 engine in
 * qoaudio_b200/.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may link or call it.  The product (libqoa_b200.so) never does.
 *
 * Parity pin status:
 *   - decode: pinned by the reference's own fixture tests (src/lib.rs:998-1074): mode,
 *     468 frame headers, per-frame sample counts, 2394122*2 samples.  The reference's tests
 *     assert no sample VALUES, but its fixture holds them: the file was written by another
 *     encoder, which stored its LMS state in every frame header, and the state this decoder
 *     reaches at the end of each frame equals the next header's, 934 times x 128 bits
 *     (oracle/vector_verify.py, tests/test_vector_verify.py) - a reference-held check of the
 *     decode arithmetic over every sample of every full frame.
 *   - encode: PARITY UNPINNED by golden vectors.  The reference holds no byte-level encoder
 *     golden; the Rust crate cannot be built here (no rustc/cargo).  The pin is this
 *     restatement + an independent brute-force search (encode_slice_brute) + an independent
 *     pure-Python restatement (oracle/qoa_pyref.py) + a slice-parallel torch verifier
 *     (oracle/vector_verify.py: all 16 candidates of every slice, lexicographic arg-min)
 *     + SURVEY.md Appendix B hashes.
 *
 * Every function cites the reference file:line it follows (paths relative to the reference
 * repo root, i.e. src/lib.rs).
 */
#ifndef QOA_ORACLE_H
#define QOA_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/lib.rs:12-19 */
#define QOA_O_SLICE_LEN 20
#define QOA_O_LMS_LEN 4
#define QOA_O_HEADER_SIZE 8
#define QOA_O_MAGIC 0x716f6166u /* "qoaf" */
#define QOA_O_SLICES_PER_FRAME 256
#define QOA_O_FRAME_LEN (QOA_O_SLICES_PER_FRAME * QOA_O_SLICE_LEN)
#define QOA_O_MAX_CHANNELS 8
#define QOA_O_MAX_SLICES_PER_CHANNEL_PER_FRAME 256

/* Status codes.  EncodeError src/lib.rs:105-115, DecodeError src/lib.rs:866-873.
 * The numeric values are shared with include/qoa_b200.h so tests can compare directly. */
enum {
    QOA_O_OK = 0,
    QOA_O_ENC_INVALID_CHANNELS = 1,
    QOA_O_ENC_INVALID_SAMPLE_RATE = 2,
    QOA_O_ENC_INVALID_SAMPLES = 3,
    QOA_O_ENC_IO_ERROR = 4,
    QOA_O_DEC_NOT_QOA_FILE = 11,
    QOA_O_DEC_NO_SAMPLES = 12,
    QOA_O_DEC_INVALID_FRAME_HEADER = 13,
    QOA_O_DEC_INCOMPATIBLE_FRAME = 14,
    QOA_O_DEC_IO_ERROR = 15 /* io::ErrorKind::UnexpectedEof */
};

/* src/lib.rs:70-74 */
typedef struct {
    int32_t history[QOA_O_LMS_LEN];
    int32_t weights[QOA_O_LMS_LEN];
} qoa_o_lms;

/* src/lib.rs:85-91 (prev_scalefactor is dead state in the reference: written :481, never read) */
typedef struct {
    uint32_t channels;
    uint32_t sample_rate;
    uint32_t samples;
    qoa_o_lms lms[QOA_O_MAX_CHANNELS];
    int brute;                    /* 0: faithful sorted+pruned search; 1: brute-force lexicographic arg-min */
    uint64_t candidate_samples;   /* work counter: (candidate, sample) evaluations performed */
    uint64_t slices;              /* slices encoded */
} qoa_o_encoder;

/* QoaEncoder::new src/lib.rs:335-361.  `channels` is u8 in the reference; values > 255 cannot occur there. */
int qoa_o_encoder_new(qoa_o_encoder *enc, uint32_t channels, uint32_t sample_rate, uint32_t samples);
/* size formula src/lib.rs:374-378 */
size_t qoa_o_encoded_size(uint32_t channels, uint32_t samples);
/* QoaEncoder::encode src/lib.rs:367-398 */
int qoa_o_encoder_encode(qoa_o_encoder *enc, const int16_t *pcm, size_t pcm_len, uint8_t *out, size_t out_cap,
                         size_t *out_len);
/* QoaEncoder::write_header src/lib.rs:404-408 */
void qoa_o_encoder_write_header(const qoa_o_encoder *enc, uint8_t out[8]);
/* QoaEncoder::encode_frame src/lib.rs:418-439.  *written = bytes written; returns status. */
int qoa_o_encoder_encode_frame(qoa_o_encoder *enc, const int16_t *pcm, size_t pcm_len, uint8_t *out, size_t out_cap,
                               size_t *written);
/* encode_all src/lib.rs:760-763 */
int qoa_o_encode_all(const int16_t *pcm, size_t pcm_len, uint32_t channels, uint32_t sample_rate, uint32_t samples,
                     int brute, uint8_t *out, size_t out_cap, size_t *out_len);

/* FrameHeader src/lib.rs:773-781 */
typedef struct {
    uint32_t num_channels;
    uint32_t sample_rate;
    uint32_t num_samples_per_channel;
} qoa_o_frame_header;

/* Result of running the QoaDecoder iterator (src/lib.rs:661-693) to its end or first error. */
typedef struct {
    int32_t status;          /* QOA_O_OK on clean end, else the DecodeError that ended iteration */
    int32_t streaming;       /* ProcessingMode::Streaming (file header samples == 0), src/lib.rs:134-144 */
    uint32_t header_samples; /* file header `samples` field */
    uint32_t channels;       /* of the first frame header ITEM (0 for the streaming-mode dummy) */
    uint32_t sample_rate;    /* idem */
    uint64_t samples;        /* QoaItem::Sample items yielded (i16 values, all channels) */
    uint64_t frames;         /* QoaItem::FrameHeader items yielded (includes the streaming-mode dummy, src/lib.rs:674-677) */
} qoa_o_decode_info;

/* QoaDecoder::new + iterate to end/error (src/lib.rs:127-169, :661-693).
 * Writes up to out_cap samples (counts all of them in info->samples) and up to hdr_cap frame headers.
 * If QoaDecoder::new itself fails, info->status is that error and nothing is yielded. */
void qoa_o_decode_iter(const uint8_t *data, size_t len, int16_t *out, size_t out_cap, qoa_o_frame_header *hdrs,
                       size_t hdr_cap, qoa_o_decode_info *info);

/* decode_all src/lib.rs:713-747 (incl. its streaming-file quirk: header samples==0 and at least one frame
 * => IncompatibleFrame; on any error the reference returns Err and no samples).  info->samples counts what
 * had been pushed before the error. */
void qoa_o_decode_all(const uint8_t *data, size_t len, int16_t *out, size_t out_cap, qoa_o_decode_info *info);

/* QoaDecoder::new_streaming + decode_frame (src/lib.rs:627-652): stateful cursor over appended bytes. */
typedef struct qoa_o_stream_decoder qoa_o_stream_decoder;
qoa_o_stream_decoder *qoa_o_stream_decoder_new(void);
void qoa_o_stream_decoder_free(qoa_o_stream_decoder *d);
/* returns status; *n_out = samples produced by this call (also on error: what was collected before it) */
int qoa_o_stream_decoder_decode_frame(qoa_o_stream_decoder *d, const uint8_t *frame, size_t len, int16_t *out,
                                      size_t out_cap, size_t *n_out);

/* Exposed primitives for unit tests (src/lib.rs:613-617, :797-829, :495-596). */
int32_t qoa_o_div(int32_t v, int sf);
int32_t qoa_o_lms_predict(const qoa_o_lms *lms);
void qoa_o_lms_update(qoa_o_lms *lms, int16_t sample, int32_t residual);
uint64_t qoa_o_encode_slice(const qoa_o_lms *lms, const int32_t *samples, int slice_len, int brute, int *best_sf,
                            qoa_o_lms *best_lms, uint64_t *work);

/* Multi-threaded CPU baseline helpers (one stream per thread, bench.py cpu_baseline / --impl reference).
 * Encodes/decodes `n` equally shaped streams laid out back to back, using `threads` pthreads.
 * Returns 0 on success. */
int qoa_o_mt_encode(const int16_t *pcm, size_t n, uint32_t channels, uint32_t sample_rate, uint32_t samples,
                    uint8_t *out, int threads);
int qoa_o_mt_decode(const uint8_t *qoa, size_t n, size_t bytes_per_stream, int16_t *out, size_t out_stride,
                    int threads);

#ifdef __cplusplus
}
#endif
#endif
