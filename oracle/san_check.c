/*
 * san_check.c — runs the CPU oracle under AddressSanitizer + UndefinedBehaviorSanitizer (TEST INFRASTRUCTURE ONLY).
 *
 * Built by `make -C oracle san_check` as  gcc -fsanitize=address,undefined -fno-sanitize-recover=all  together with
 * qoa_oracle.c (no -fwrapv: every signed overflow the restatement relies on would be reported, so wrapping ops must
 * be done in uint32_t exactly as SURVEY.md §5 asks).  Exercises: encode (faithful + brute-force search) of ramps,
 * full-scale square waves and LCG noise for 1..8 channels and odd lengths, streaming encode, decode of the results,
 * decode of every truncation of a small file and of mutated headers (the shapes of the reference's fuzz targets,
 * fuzz/fuzz_targets/iter_all.rs:5-10 and encode_round_trip.rs:5-47).  Exit code 0 = clean.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "qoa_oracle.h"

static uint32_t lcg_state = 12345u;
static uint32_t lcg(void)
{
    lcg_state = lcg_state * 1664525u + 1013904223u;
    return lcg_state >> 8;
}

static int fail(const char *what, int a, int b)
{
    fprintf(stderr, "san_check: %s (%d, %d)\n", what, a, b);
    return 1;
}

int main(void)
{
    static const uint32_t lens[] = {1, 19, 20, 21, 255, 5119, 5120, 5121, 7777};
    int checks = 0;
    for (uint32_t ch = 1; ch <= 8; ++ch) {
        for (size_t li = 0; li < sizeof lens / sizeof lens[0]; ++li) {
            const uint32_t n = lens[li];
            if (ch > 2 && n > 5121) continue;
            for (int kind = 0; kind < 4; ++kind) {
                int16_t *pcm = malloc((size_t)n * ch * sizeof *pcm);
                for (size_t i = 0; i < (size_t)n * ch; ++i) {
                    const uint32_t t = (uint32_t)(i / ch), c = (uint32_t)(i % ch);
                    int v;
                    switch (kind) {
                    case 0: v = (int)((t * 97u + c * 1000u) % 65536u) - 32768; break;      /* ramp */
                    case 1: v = ((t / (3 + c)) & 1) ? 32767 : -32768; break;              /* full-scale square */
                    case 2: v = (int)(lcg() & 0xffffu) - 32768; break;                    /* white noise */
                    default: v = 0; break;                                                /* silence */
                    }
                    pcm[i] = (int16_t)v;
                }
                const size_t cap = qoa_o_encoded_size(ch, n);
                uint8_t *a = malloc(cap), *b = malloc(cap), *s = malloc(cap);
                size_t la = 0, lb = 0;
                if (qoa_o_encode_all(pcm, (size_t)n * ch, ch, 8000 + 1000 * ch, n, 0, a, cap, &la) != QOA_O_OK || la != cap)
                    return fail("encode_all", (int)ch, (int)n);
                if (n <= 5121 && kind != 3) { /* the brute-force search must pick the same bytes */
                    if (qoa_o_encode_all(pcm, (size_t)n * ch, ch, 8000 + 1000 * ch, n, 1, b, cap, &lb) != QOA_O_OK ||
                        lb != la || memcmp(a, b, la) != 0)
                        return fail("brute-force search differs", (int)ch, (int)n);
                }
                /* streaming encode == one-shot (src/lib.rs:1261-1297) */
                qoa_o_encoder e;
                if (qoa_o_encoder_new(&e, ch, 8000 + 1000 * ch, n) != QOA_O_OK) return fail("encoder_new", (int)ch, (int)n);
                qoa_o_encoder_write_header(&e, s);
                size_t at = 8;
                for (uint32_t f0 = 0; f0 < n; f0 += QOA_O_FRAME_LEN) {
                    const uint32_t fl = n - f0 < QOA_O_FRAME_LEN ? n - f0 : QOA_O_FRAME_LEN;
                    size_t w = 0;
                    if (qoa_o_encoder_encode_frame(&e, pcm + (size_t)f0 * ch, (size_t)fl * ch, s + at, cap - at, &w) != QOA_O_OK)
                        return fail("encode_frame", (int)ch, (int)n);
                    at += w;
                }
                if (at != la || memcmp(a, s, la) != 0) return fail("streaming encode differs", (int)ch, (int)n);
                /* decode it back, both entry points */
                int16_t *dec = malloc((size_t)n * ch * sizeof *dec + 2);
                qoa_o_frame_header hdrs[4];
                qoa_o_decode_info info;
                qoa_o_decode_all(a, la, dec, (size_t)n * ch, &info);
                if (info.status != QOA_O_OK || info.samples != (uint64_t)n * ch) return fail("decode_all", (int)ch, (int)n);
                qoa_o_decode_iter(a, la, dec, (size_t)n * ch, hdrs, 4, &info);
                if (info.status != QOA_O_OK || info.samples != (uint64_t)n * ch) return fail("decode_iter", (int)ch, (int)n);
                /* every truncation and a few header mutations of small files: must end with a status, never a fault */
                if (la <= 600) {
                    for (size_t cut = 0; cut <= la; ++cut) {
                        uint8_t *t = malloc(cut ? cut : 1); /* exact-size copy: any over-read trips ASan */
                        memcpy(t, a, cut);
                        qoa_o_decode_iter(t, cut, dec, (size_t)n * ch, hdrs, 4, &info);
                        qoa_o_decode_all(t, cut, dec, (size_t)n * ch, &info);
                        free(t);
                    }
                    for (int m = 0; m < 64; ++m) {
                        uint8_t *t = malloc(la);
                        memcpy(t, a, la);
                        t[lcg() % (la < 40 ? la : 40)] ^= (uint8_t)(1u << (lcg() & 7));
                        qoa_o_decode_iter(t, la, dec, (size_t)n * ch, hdrs, 4, &info);
                        qoa_o_decode_all(t, la, dec, (size_t)n * ch, &info);
                        free(t);
                    }
                }
                /* random payload behind valid headers: predictions and weights wrap all over the i32 range */
                for (size_t i = 8 + 8; i < la; ++i) a[i] = (uint8_t)lcg();
                {
                    const uint32_t frames = (n + QOA_O_FRAME_LEN - 1) / QOA_O_FRAME_LEN;
                    const size_t fsz = 8 + 16 * (size_t)ch + 8 * (size_t)QOA_O_SLICES_PER_FRAME * ch;
                    for (uint32_t f = 1; f < frames; ++f) memcpy(a + 8 + f * fsz, b ? s + 8 + f * fsz : a, 8); /* keep frame headers valid */
                    memcpy(a + 8, s + 8, 8);
                }
                qoa_o_decode_all(a, la, dec, (size_t)n * ch, &info);
                if (info.status != QOA_O_OK) return fail("random payload decode", (int)ch, (int)n);
                ++checks;
                free(dec); free(a); free(b); free(s); free(pcm);
            }
        }
    }
    /* streaming decoder, frame by frame */
    {
        const uint32_t ch = 2, n = 5120 + 333;
        int16_t *pcm = calloc((size_t)n * ch, sizeof *pcm);
        for (size_t i = 0; i < (size_t)n * ch; ++i) pcm[i] = (int16_t)((int)(lcg() & 0x3fff) - 8192);
        const size_t cap = qoa_o_encoded_size(ch, n);
        uint8_t *a = malloc(cap);
        size_t la = 0;
        qoa_o_encode_all(pcm, (size_t)n * ch, ch, 44100, n, 0, a, cap, &la);
        qoa_o_stream_decoder *d = qoa_o_stream_decoder_new();
        int16_t *out = malloc((size_t)QOA_O_FRAME_LEN * ch * sizeof *out);
        const size_t f0 = 8 + 16 * ch + 8 * 256 * ch;
        size_t got = 0;
        if (qoa_o_stream_decoder_decode_frame(d, a + 8, f0, out, (size_t)QOA_O_FRAME_LEN * ch, &got) != QOA_O_OK || got != 5120u * ch)
            return fail("stream decoder frame 0", (int)got, 0);
        if (qoa_o_stream_decoder_decode_frame(d, a + 8 + f0, la - 8 - f0, out, (size_t)QOA_O_FRAME_LEN * ch, &got) != QOA_O_OK ||
            got != 333u * ch)
            return fail("stream decoder frame 1", (int)got, 0);
        qoa_o_stream_decoder_free(d);
        free(out); free(a); free(pcm);
    }
    printf("san_check ok: %d shapes, no sanitizer report\n", checks);
    return 0;
}
