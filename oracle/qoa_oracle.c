/*
 * qoa_oracle.c — CPU restatement (plain C) of the QOA encode/decode path of rafaelcaricio/qoaudio.
 *
 * TEST INFRASTRUCTURE ONLY — see qoa_oracle.h.  Encode parity is UNPINNED by reference golden
 * vectors (none exist); decode is pinned structurally by the reference's fixture tests.
 *
 * All wrapping arithmetic is done in uint32_t so the file is UB-free even without -fwrapv
 * (the recipe passes -fwrapv anyway).  `>>` on negative int32_t is arithmetic with gcc.
 * Citations are to the reference's src/lib.rs.
 */
#include "qoa_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ---- tables: src/lib.rs:22, :25-27, :847-864 ------------------------------------------- */
static const int32_t QUANT_TAB[17] = {7, 7, 7, 5, 5, 3, 3, 1, 0, 0, 2, 2, 4, 4, 6, 6, 6};
static const int32_t RECIPROCAL_TAB[16] = {65536, 9363, 3121, 1457, 781, 475, 311, 216,
                                           156,   117,  90,   71,   57,  47,  39,  32};
static const int32_t DEQUANT_TAB[16][8] = {
    {1, -1, 3, -3, 5, -5, 7, -7},
    {5, -5, 18, -18, 32, -32, 49, -49},
    {16, -16, 53, -53, 95, -95, 147, -147},
    {34, -34, 113, -113, 203, -203, 315, -315},
    {63, -63, 210, -210, 378, -378, 588, -588},
    {104, -104, 345, -345, 621, -621, 966, -966},
    {158, -158, 528, -528, 950, -950, 1477, -1477},
    {228, -228, 760, -760, 1368, -1368, 2128, -2128},
    {316, -316, 1053, -1053, 1895, -1895, 2947, -2947},
    {422, -422, 1405, -1405, 2529, -2529, 3934, -3934},
    {548, -548, 1828, -1828, 3290, -3290, 5117, -5117},
    {696, -696, 2320, -2320, 4176, -4176, 6496, -6496},
    {868, -868, 2893, -2893, 5207, -5207, 8099, -8099},
    {1064, -1064, 3548, -3548, 6386, -6386, 9933, -9933},
    {1286, -1286, 4288, -4288, 7718, -7718, 12005, -12005},
    {1536, -1536, 5120, -5120, 9216, -9216, 14336, -14336},
};

static inline int32_t wmul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
static inline int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static inline int32_t wsub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
static inline int32_t clampi(int32_t v, int32_t lo, int32_t hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int sgn(int32_t v) { return (v > 0) - (v < 0); }

/* ---- LMS: src/lib.rs:797-829 -------------------------------------------------------------- */
int32_t qoa_o_lms_predict(const qoa_o_lms *l) /* :799-805 */
{
    int32_t p01 = wadd(wmul(l->weights[0], l->history[0]), wmul(l->weights[1], l->history[1]));
    int32_t p23 = wadd(wmul(l->weights[2], l->history[2]), wmul(l->weights[3], l->history[3]));
    return wadd(p01, p23) >> 13;
}

static inline int32_t predict_and_penalty_sq(const qoa_o_lms *l, uint64_t *penalty_sq) /* :808-818 */
{
    const int32_t *w = l->weights;
    int32_t prediction = qoa_o_lms_predict(l);
    int32_t s01 = wadd(wmul(w[0], w[0]), wmul(w[1], w[1]));
    int32_t s23 = wadd(wmul(w[2], w[2]), wmul(w[3], w[3]));
    int32_t pen = wsub(wadd(s01, s23) >> 18, 0x8ff);
    int64_t penalty = pen > 0 ? pen : 0;
    *penalty_sq = (uint64_t)(penalty * penalty);
    return prediction;
}

void qoa_o_lms_update(qoa_o_lms *l, int16_t sample, int32_t residual) /* :821-828 */
{
    int32_t delta = residual >> 4;
    for (int i = 0; i < QOA_O_LMS_LEN; i++)
        l->weights[i] = wadd(l->weights[i], l->history[i] < 0 ? wsub(0, delta) : delta);
    l->history[0] = l->history[1];
    l->history[1] = l->history[2];
    l->history[2] = l->history[3];
    l->history[3] = (int32_t)sample;
}

/* qoa_div: src/lib.rs:613-617 */
int32_t qoa_o_div(int32_t v, int sf)
{
    int32_t n = wadd(wmul(v, RECIPROCAL_TAB[sf]), 1 << 15) >> 16;
    return n + sgn(v) - sgn(n);
}

/* One (candidate, sample) step with scalefactor sf on state `l`.  Body of src/lib.rs:564-585
 * (and :524-533 for the first sample).  Returns this sample's rank contribution. */
static inline uint64_t candidate_step(qoa_o_lms *l, int32_t sample, int sf, uint64_t *slice)
{
    uint64_t penalty_sq;
    int32_t predicted = predict_and_penalty_sq(l, &penalty_sq);
    int32_t residual = wsub(sample, predicted);
    int32_t scaled = qoa_o_div(residual, sf);
    int32_t clamped = clampi(scaled, -8, 8);
    int32_t quantized = QUANT_TAB[clamped + 8];            /* fused LUT :831-845 */
    int32_t dequantized = DEQUANT_TAB[sf][quantized];
    int32_t reconstructed = clampi(wadd(predicted, dequantized), -32768, 32767);
    int64_t error = (int64_t)wsub(sample, reconstructed);
    qoa_o_lms_update(l, (int16_t)reconstructed, dequantized);
    *slice = (*slice << 3) | (uint64_t)quantized;
    return (uint64_t)(error * error) + penalty_sq;
}

/* encode_slice: src/lib.rs:495-596.
 * brute == 0: the reference's control flow — first-sample pass for all 16 scalefactors (:516-542), stable
 *             insertion sort by first-sample rank (:535-541), sorted search with break (:555-557) and pruning
 *             (:578-581), strict `<` replacement (:587).
 * brute == 1: independent formulation — evaluate all 16 candidates fully and take the lexicographic arg-min of
 *             (total_rank, first_sample_rank, sf).  Must give identical results.
 * Returns the unshifted slice word (sf followed by slice_len 3-bit codes). */
uint64_t qoa_o_encode_slice(const qoa_o_lms *lms, const int32_t *samples, int slice_len, int brute, int *best_sf_out,
                            qoa_o_lms *best_lms_out, uint64_t *work)
{
    uint64_t best_rank = UINT64_MAX, best_slice = 0;
    int best_sf = 0;
    qoa_o_lms best_lms;
    memset(&best_lms, 0, sizeof best_lms);
    uint64_t evals = 0;

    if (brute) {
        uint64_t best_first = UINT64_MAX;
        for (int sf = 0; sf < 16; sf++) {
            qoa_o_lms l = *lms;
            uint64_t slice = (uint64_t)sf, total = 0, first = 0;
            for (int i = 0; i < slice_len; i++) {
                uint64_t r = candidate_step(&l, samples[i], sf, &slice);
                if (i == 0) first = r;
                total += r;
                evals++;
            }
            if (total < best_rank || (total == best_rank && first < best_first)) { /* ascending sf breaks the rest */
                best_rank = total;
                best_first = first;
                best_slice = slice;
                best_sf = sf;
                best_lms = l;
            }
        }
    } else {
        /* first-sample pass :516-542 */
        uint64_t first_rank[16], first_slice[16];
        qoa_o_lms first_lms[16];
        int order[16], count = 0;
        for (int sf = 0; sf < 16; sf++) {
            first_lms[sf] = *lms;
            first_slice[sf] = (uint64_t)sf;
            first_rank[sf] = candidate_step(&first_lms[sf], samples[0], sf, &first_slice[sf]);
            evals++;
            int pos = count;
            while (pos > 0 && first_rank[order[pos - 1]] > first_rank[sf]) { /* strict > : stable */
                order[pos] = order[pos - 1];
                pos--;
            }
            order[pos] = sf;
            count++;
        }
        /* sorted, pruned search :544-593 */
        for (int k = 0; k < 16; k++) {
            int sf = order[k];
            uint64_t current = first_rank[sf];
            if (current > best_rank) break; /* :555-557 */
            qoa_o_lms l = first_lms[sf];
            uint64_t slice = first_slice[sf];
            int valid = 1;
            for (int i = 1; i < slice_len; i++) {
                /* the reference adds the rank and tests BEFORE update/pack (:575-584); a pruned candidate's
                 * state is discarded, so stepping first and testing after is equivalent. */
                current += candidate_step(&l, samples[i], sf, &slice);
                evals++;
                if (current > best_rank) { /* :578-581 */
                    valid = 0;
                    break;
                }
            }
            if (valid && current < best_rank) { /* :587 */
                best_rank = current;
                best_slice = slice;
                best_sf = sf;
                best_lms = l;
            }
        }
    }
    if (best_sf_out) *best_sf_out = best_sf;
    if (best_lms_out) *best_lms_out = best_lms;
    if (work) *work += evals;
    return best_slice;
}

/* ---- big-endian helpers -------------------------------------------------------------------- */
static inline void put_be32(uint8_t *p, uint32_t v)
{
    p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v;
}
static inline void put_be64(uint8_t *p, uint64_t v)
{
    put_be32(p, (uint32_t)(v >> 32));
    put_be32(p + 4, (uint32_t)v);
}
static inline uint32_t get_be32(const uint8_t *p)
{
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
}
static inline uint64_t get_be64(const uint8_t *p) { return ((uint64_t)get_be32(p) << 32) | get_be32(p + 4); }

/* ---- encoder --------------------------------------------------------------------------------- */
int qoa_o_encoder_new(qoa_o_encoder *enc, uint32_t channels, uint32_t sample_rate, uint32_t samples) /* :335-361 */
{
    memset(enc, 0, sizeof *enc);
    if (channels == 0 || channels > QOA_O_MAX_CHANNELS) return QOA_O_ENC_INVALID_CHANNELS;
    if (sample_rate == 0) return QOA_O_ENC_INVALID_SAMPLE_RATE;
    if (samples == 0) return QOA_O_ENC_INVALID_SAMPLES;
    enc->channels = channels;
    enc->sample_rate = sample_rate;
    enc->samples = samples;
    for (uint32_t c = 0; c < channels; c++) { /* :347-352 */
        enc->lms[c].weights[2] = -(1 << 13);
        enc->lms[c].weights[3] = 1 << 14;
    }
    return QOA_O_OK;
}

size_t qoa_o_encoded_size(uint32_t channels, uint32_t samples) /* :374-378 */
{
    size_t total = samples, ch = channels;
    size_t num_slices = (total + QOA_O_SLICE_LEN - 1) / QOA_O_SLICE_LEN;
    size_t num_frames = (total + QOA_O_FRAME_LEN - 1) / QOA_O_FRAME_LEN;
    return QOA_O_HEADER_SIZE + num_frames * (8 + QOA_O_LMS_LEN * 4 * ch) + num_slices * 8 * ch;
}

static inline uint32_t frame_size_u16(size_t channels, size_t slices) /* qoa_frame_size :602-604 (as u16) */
{
    return (uint32_t)((8 + QOA_O_LMS_LEN * 4 * channels + 8 * slices * channels) & 0xffff);
}

/* encode_frame_to_buf: src/lib.rs:441-493.  pcm holds frame_len*channels interleaved samples. */
static size_t encode_frame_to_buf(qoa_o_encoder *enc, const int16_t *pcm, size_t frame_len, uint8_t *buf)
{
    size_t ch = enc->channels;
    size_t slices = (frame_len + QOA_O_SLICE_LEN - 1) / QOA_O_SLICE_LEN;
    size_t p = 0;
    uint64_t header = ((uint64_t)(enc->channels & 0xff) << 56) | ((uint64_t)enc->sample_rate << 32) |
                      ((uint64_t)frame_len << 16) | (uint64_t)frame_size_u16(ch, slices); /* :448-451 */
    put_be64(buf + p, header);
    p += 8;
    for (size_t c = 0; c < ch; c++) { /* :455-466: truncating `as u16` */
        uint64_t history = 0, weights = 0;
        for (int i = 0; i < QOA_O_LMS_LEN; i++) {
            history = (history << 16) | (uint64_t)(uint16_t)enc->lms[c].history[i];
            weights = (weights << 16) | (uint64_t)(uint16_t)enc->lms[c].weights[i];
        }
        put_be64(buf + p, history);
        put_be64(buf + p + 8, weights);
        p += 16;
    }
    for (size_t s0 = 0; s0 < frame_len; s0 += QOA_O_SLICE_LEN) { /* :468-491 slice-major, channel-minor */
        size_t slice_len = frame_len - s0 < QOA_O_SLICE_LEN ? frame_len - s0 : QOA_O_SLICE_LEN;
        for (size_t c = 0; c < ch; c++) {
            int32_t samples[QOA_O_SLICE_LEN];
            for (size_t i = 0; i < slice_len; i++) samples[i] = (int32_t)pcm[(s0 + i) * ch + c]; /* :502-508 */
            qoa_o_lms best_lms;
            int best_sf;
            uint64_t slice = qoa_o_encode_slice(&enc->lms[c], samples, (int)slice_len, enc->brute, &best_sf,
                                                &best_lms, &enc->candidate_samples);
            enc->lms[c] = best_lms; /* :482 — full i32 state carried, never truncated */
            enc->slices++;
            if (slice_len < QOA_O_SLICE_LEN) slice <<= (QOA_O_SLICE_LEN - slice_len) * 3; /* :485-487 */
            put_be64(buf + p, slice);
            p += 8;
        }
    }
    return p;
}

int qoa_o_encoder_encode(qoa_o_encoder *enc, const int16_t *pcm, size_t pcm_len, uint8_t *out, size_t out_cap,
                         size_t *out_len) /* :367-398 */
{
    if (pcm_len != (size_t)enc->samples * enc->channels) return QOA_O_ENC_INVALID_SAMPLES;
    size_t need = qoa_o_encoded_size(enc->channels, enc->samples);
    if (out_cap < need) return QOA_O_ENC_IO_ERROR;
    size_t p = 0, ch = enc->channels, total = enc->samples;
    qoa_o_encoder_write_header(enc, out);
    p += 8;
    for (size_t idx = 0; idx < total;) {
        size_t frame_len = total - idx < QOA_O_FRAME_LEN ? total - idx : QOA_O_FRAME_LEN;
        p += encode_frame_to_buf(enc, pcm + idx * ch, frame_len, out + p);
        idx += frame_len;
    }
    *out_len = p;
    return QOA_O_OK;
}

void qoa_o_encoder_write_header(const qoa_o_encoder *enc, uint8_t out[8]) /* :404-408 */
{
    put_be32(out, QOA_O_MAGIC);
    put_be32(out + 4, enc->samples);
}

int qoa_o_encoder_encode_frame(qoa_o_encoder *enc, const int16_t *pcm, size_t pcm_len, uint8_t *out, size_t out_cap,
                               size_t *written) /* :418-439 */
{
    size_t ch = enc->channels;
    *written = 0;
    if (pcm_len == 0 || pcm_len % ch != 0) return QOA_O_ENC_INVALID_SAMPLES;
    size_t frame_len = pcm_len / ch;
    if (frame_len > QOA_O_FRAME_LEN) return QOA_O_ENC_INVALID_SAMPLES;
    size_t slices = (frame_len + QOA_O_SLICE_LEN - 1) / QOA_O_SLICE_LEN;
    size_t frame_size = 8 + 16 * ch + 8 * slices * ch;
    if (out_cap < frame_size) return QOA_O_ENC_IO_ERROR;
    *written = encode_frame_to_buf(enc, pcm, frame_len, out);
    return QOA_O_OK;
}

int qoa_o_encode_all(const int16_t *pcm, size_t pcm_len, uint32_t channels, uint32_t sample_rate, uint32_t samples,
                     int brute, uint8_t *out, size_t out_cap, size_t *out_len) /* :760-763 */
{
    qoa_o_encoder enc;
    int st = qoa_o_encoder_new(&enc, channels, sample_rate, samples);
    if (st != QOA_O_OK) return st;
    enc.brute = brute;
    return qoa_o_encoder_encode(&enc, pcm, pcm_len, out, out_cap, out_len);
}

/* ---- decoder ---------------------------------------------------------------------------------- */

/* Mirrors QoaDecoder<R> over a Cursor (src/lib.rs:59-68). */
typedef struct {
    const uint8_t *data;
    size_t len, pos;
    int streaming;                 /* ProcessingMode */
    uint32_t header_samples;
    uint32_t fixed_channels, fixed_rate;
    uint32_t nlms;                 /* self.lms.len() */
    qoa_o_lms lms[255];
    qoa_o_frame_header cur;        /* current_frame.header */
    uint32_t remaining;            /* num_samples_per_channel_remaining */
    int returned_first_header;
} dec_state;

/* read_exact on a Cursor: on short read the cursor moves to the end and UnexpectedEof is returned
 * (src/lib.rs:783-795 over std::io::Cursor). */
static int rd(dec_state *d, size_t n, const uint8_t **p)
{
    if (d->len - d->pos < n) {
        d->pos = d->len;
        return 0;
    }
    *p = d->data + d->pos;
    d->pos += n;
    return 1;
}

/* decode_frame_header_and_lms: src/lib.rs:205-289.  Returns 1 frame read, 0 clean EOF, <0 = -(error). */
static int dec_frame_header_and_lms(dec_state *d, int first)
{
    const uint8_t *p;
    if (!rd(d, 8, &p)) return 0; /* :206-216 UnexpectedEof at header => Ok(false) */
    uint64_t h = get_be64(p);
    uint32_t num_channels = (uint32_t)((h >> 56) & 0xff);
    uint32_t sample_rate = (uint32_t)((h >> 32) & 0xffffff);
    uint32_t nspc = (uint32_t)((h >> 16) & 0xffff);
    uint32_t frame_size = (uint32_t)(h & 0xffff);

    if (num_channels == 0 || sample_rate == 0) return -QOA_O_DEC_INVALID_FRAME_HEADER; /* :227 */
    size_t non_sample = QOA_O_HEADER_SIZE + QOA_O_LMS_LEN * 4 * (size_t)num_channels;    /* :232 */
    if ((size_t)frame_size <= non_sample) return -QOA_O_DEC_INVALID_FRAME_HEADER;       /* :233 */
    size_t num_slices = ((size_t)frame_size - non_sample) / 8;                           /* :236-237 */
    if (num_slices % num_channels != 0) return -QOA_O_DEC_INVALID_FRAME_HEADER;          /* :239 */
    if (num_slices / num_channels > QOA_O_MAX_SLICES_PER_CHANNEL_PER_FRAME)              /* :242 */
        return -QOA_O_DEC_INVALID_FRAME_HEADER;

    if (!d->streaming) { /* :246-259 */
        if (first) {
            d->fixed_channels = num_channels;
            d->fixed_rate = sample_rate;
        } else if (num_channels != d->fixed_channels || sample_rate != d->fixed_rate) {
            return -QOA_O_DEC_INCOMPATIBLE_FRAME;
        }
    }
    if (d->nlms != num_channels) { /* :263-268 resize_with(Default): keeps the common prefix, zero-fills growth */
        for (uint32_t c = d->nlms; c < num_channels; c++) memset(&d->lms[c], 0, sizeof d->lms[c]);
        d->nlms = num_channels;
    }
    for (uint32_t c = 0; c < num_channels; c++) { /* :271-281 */
        const uint8_t *ph, *pw;
        if (!rd(d, 8, &ph)) return -QOA_O_DEC_IO_ERROR;
        if (!rd(d, 8, &pw)) return -QOA_O_DEC_IO_ERROR;
        uint64_t history = get_be64(ph), weights = get_be64(pw);
        for (int i = 0; i < QOA_O_LMS_LEN; i++) {
            d->lms[c].history[i] = (int32_t)(int16_t)(history >> 48);
            history <<= 16;
            d->lms[c].weights[i] = (int32_t)(int16_t)(weights >> 48);
            weights <<= 16;
        }
    }
    d->cur.num_channels = num_channels; /* :283-286 */
    d->cur.sample_rate = sample_rate;
    d->cur.num_samples_per_channel = nspc;
    d->remaining = nspc;
    return 1;
}

/* QoaDecoder::new: src/lib.rs:127-169.  Returns 0 or the DecodeError. */
static int dec_new(dec_state *d, const uint8_t *data, size_t len)
{
    memset(d, 0, offsetof(dec_state, lms));
    d->data = data;
    d->len = len;
    d->nlms = 0;
    memset(&d->cur, 0, sizeof d->cur);
    d->remaining = 0;
    d->returned_first_header = 0;
    const uint8_t *p;
    if (!rd(d, 4, &p)) return QOA_O_DEC_IO_ERROR;
    if (get_be32(p) != QOA_O_MAGIC) return QOA_O_DEC_NOT_QOA_FILE;
    if (!rd(d, 4, &p)) return QOA_O_DEC_IO_ERROR;
    d->header_samples = get_be32(p);
    d->streaming = d->header_samples == 0;
    if (!d->streaming) {
        int r = dec_frame_header_and_lms(d, 1);
        if (r < 0) return -r;
        if (r == 0) return QOA_O_DEC_NO_SAMPLES;
    }
    return 0;
}

/* decode_one_slice_per_channel: src/lib.rs:291-330.  pending must hold 20*channels samples.
 * Returns number of samples to emit (>0) or -(error). */
static long dec_slice_group(dec_state *d, int16_t *pending)
{
    uint32_t channels = d->cur.num_channels;
    for (uint32_t c = 0; c < channels; c++) {
        const uint8_t *p;
        if (!rd(d, 8, &p)) return -QOA_O_DEC_IO_ERROR; /* :301 */
        uint64_t slice = get_be64(p);
        int sf = (int)((slice >> 60) & 0xf);
        qoa_o_lms l = d->lms[c];
        for (int i = 0; i < QOA_O_SLICE_LEN; i++) {
            int32_t prediction = qoa_o_lms_predict(&l);
            int q = (int)((slice >> 57) & 0x7);
            int32_t dequantized = DEQUANT_TAB[sf][q];
            int32_t reconstructed = clampi(wadd(prediction, dequantized), -32768, 32767);
            pending[(size_t)i * channels + c] = (int16_t)reconstructed;
            slice <<= 3;
            qoa_o_lms_update(&l, (int16_t)reconstructed, dequantized);
        }
        d->lms[c] = l;
    }
    uint32_t n = d->remaining; /* :320-328 */
    if (n < QOA_O_SLICE_LEN) {
        d->remaining = 0;
        return (long)n * channels;
    }
    d->remaining -= QOA_O_SLICE_LEN;
    return (long)QOA_O_SLICE_LEN * channels;
}

typedef struct {
    int16_t *out;
    size_t out_cap;
    qoa_o_frame_header *hdrs;
    size_t hdr_cap;
    uint64_t samples, frames;
    int have_first;
    qoa_o_frame_header first;
    int check_compat; /* decode_all's own check :735-738 */
} sink;

static void sink_header(sink *s, const qoa_o_frame_header *h)
{
    if (s->hdrs && s->frames < s->hdr_cap) s->hdrs[s->frames] = *h;
    if (!s->have_first) {
        s->first = *h;
        s->have_first = 1;
    }
    s->frames++;
}

/* Iterator::next driven to completion: src/lib.rs:661-693.  Returns 0 at clean end or the DecodeError. */
static int dec_iterate(dec_state *d, sink *s)
{
    int16_t pending[QOA_O_SLICE_LEN * 255];
    for (;;) {
        if (!d->returned_first_header) { /* :674-678 (streaming mode: a default all-zero header) */
            d->returned_first_header = 1;
            sink_header(s, &d->cur);
            continue;
        }
        if (d->remaining > 0) { /* :679-682 */
            long n = dec_slice_group(d, pending);
            if (n < 0) return (int)-n;
            for (long i = 0; i < n; i++) {
                if (s->out && s->samples < s->out_cap) s->out[s->samples] = pending[i];
                s->samples++;
            }
        } else { /* :683-688 */
            int r = dec_frame_header_and_lms(d, 0);
            if (r == 0) return 0;
            if (r < 0) return -r;
            sink_header(s, &d->cur);
            if (s->check_compat && (d->cur.num_channels != s->first.num_channels ||
                                    d->cur.sample_rate != s->first.sample_rate))
                return QOA_O_DEC_INCOMPATIBLE_FRAME; /* decode_all :736-738 */
        }
    }
}

static void run_decode(const uint8_t *data, size_t len, int16_t *out, size_t out_cap, qoa_o_frame_header *hdrs,
                       size_t hdr_cap, qoa_o_decode_info *info, int check_compat)
{
    dec_state *d = (dec_state *)malloc(sizeof *d);
    memset(info, 0, sizeof *info);
    int st = dec_new(d, data, len);
    info->streaming = d->streaming;
    info->header_samples = d->header_samples;
    if (st != 0) {
        info->status = st;
        free(d);
        return;
    }
    sink s;
    memset(&s, 0, sizeof s);
    s.out = out;
    s.out_cap = out_cap;
    s.hdrs = hdrs;
    s.hdr_cap = hdr_cap;
    s.check_compat = check_compat;
    info->status = dec_iterate(d, &s);
    info->samples = s.samples;
    info->frames = s.frames;
    info->channels = s.first.num_channels;
    info->sample_rate = s.first.sample_rate;
    free(d);
}

void qoa_o_decode_iter(const uint8_t *data, size_t len, int16_t *out, size_t out_cap, qoa_o_frame_header *hdrs,
                       size_t hdr_cap, qoa_o_decode_info *info)
{
    run_decode(data, len, out, out_cap, hdrs, hdr_cap, info, 0);
}

void qoa_o_decode_all(const uint8_t *data, size_t len, int16_t *out, size_t out_cap, qoa_o_decode_info *info)
{
    run_decode(data, len, out, out_cap, NULL, 0, info, 1);
}

/* ---- streaming decoder: src/lib.rs:627-652 ------------------------------------------------------ */
struct qoa_o_stream_decoder {
    dec_state d;
    uint8_t *buf;
    size_t cap;
};

qoa_o_stream_decoder *qoa_o_stream_decoder_new(void)
{
    qoa_o_stream_decoder *sd = (qoa_o_stream_decoder *)calloc(1, sizeof *sd);
    sd->cap = 64;
    sd->buf = (uint8_t *)malloc(sd->cap);
    put_be32(sd->buf, QOA_O_MAGIC); /* :633-637 */
    put_be32(sd->buf + 4, 0);
    int st = dec_new(&sd->d, sd->buf, 8);
    (void)st; /* cannot fail: magic ok, streaming */
    return sd;
}

void qoa_o_stream_decoder_free(qoa_o_stream_decoder *sd)
{
    if (!sd) return;
    free(sd->buf);
    free(sd);
}

int qoa_o_stream_decoder_decode_frame(qoa_o_stream_decoder *sd, const uint8_t *frame, size_t len, int16_t *out,
                                      size_t out_cap, size_t *n_out) /* :641-651 */
{
    if (sd->d.len + len > sd->cap) {
        while (sd->d.len + len > sd->cap) sd->cap *= 2;
        sd->buf = (uint8_t *)realloc(sd->buf, sd->cap);
    }
    memcpy(sd->buf + sd->d.len, frame, len); /* extend_from_slice :642 */
    sd->d.data = sd->buf;
    sd->d.len += len;
    sink s;
    memset(&s, 0, sizeof s);
    s.out = out;
    s.out_cap = out_cap;
    int st = dec_iterate(&sd->d, &s);
    *n_out = (size_t)s.samples;
    return st;
}

/* ---- multi-threaded CPU baseline (one stream per thread) ----------------------------------------- */
typedef struct {
    const int16_t *pcm;
    const uint8_t *qoa;
    uint8_t *enc_out;
    int16_t *dec_out;
    size_t n, bytes_per_stream, out_stride;
    uint32_t channels, sample_rate, samples;
    int tid, threads, failed;
} mt_job;

static void *mt_encode_worker(void *arg)
{
    mt_job *j = (mt_job *)arg;
    size_t in_stride = (size_t)j->samples * j->channels;
    size_t out_stride = qoa_o_encoded_size(j->channels, j->samples);
    for (size_t i = (size_t)j->tid; i < j->n; i += (size_t)j->threads) {
        size_t out_len = 0;
        int st = qoa_o_encode_all(j->pcm + i * in_stride, in_stride, j->channels, j->sample_rate, j->samples, 0,
                                  j->enc_out + i * out_stride, out_stride, &out_len);
        if (st != QOA_O_OK || out_len != out_stride) j->failed = 1;
    }
    return NULL;
}

static void *mt_decode_worker(void *arg)
{
    mt_job *j = (mt_job *)arg;
    for (size_t i = (size_t)j->tid; i < j->n; i += (size_t)j->threads) {
        qoa_o_decode_info info;
        qoa_o_decode_all(j->qoa + i * j->bytes_per_stream, j->bytes_per_stream, j->dec_out + i * j->out_stride,
                         j->out_stride, &info);
        if (info.status != QOA_O_OK) j->failed = 1;
    }
    return NULL;
}

static int mt_run(mt_job *proto, void *(*fn)(void *))
{
    int threads = proto->threads < 1 ? 1 : proto->threads;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    mt_job *jobs = (mt_job *)malloc(sizeof(mt_job) * (size_t)threads);
    int failed = 0;
    for (int t = 0; t < threads; t++) {
        jobs[t] = *proto;
        jobs[t].tid = t;
        jobs[t].threads = threads;
        pthread_create(&th[t], NULL, fn, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) {
        pthread_join(th[t], NULL);
        failed |= jobs[t].failed;
    }
    free(th);
    free(jobs);
    return failed;
}

int qoa_o_mt_encode(const int16_t *pcm, size_t n, uint32_t channels, uint32_t sample_rate, uint32_t samples,
                    uint8_t *out, int threads)
{
    mt_job j;
    memset(&j, 0, sizeof j);
    j.pcm = pcm;
    j.enc_out = out;
    j.n = n;
    j.channels = channels;
    j.sample_rate = sample_rate;
    j.samples = samples;
    j.threads = threads;
    return mt_run(&j, mt_encode_worker);
}

int qoa_o_mt_decode(const uint8_t *qoa, size_t n, size_t bytes_per_stream, int16_t *out, size_t out_stride,
                    int threads)
{
    mt_job j;
    memset(&j, 0, sizeof j);
    j.qoa = qoa;
    j.dec_out = out;
    j.n = n;
    j.bytes_per_stream = bytes_per_stream;
    j.out_stride = out_stride;
    j.threads = threads;
    return mt_run(&j, mt_decode_worker);
}
