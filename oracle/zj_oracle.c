/*
 * zj_oracle.c -- CPU oracle: literal C restatement of the reference's post-entropy path.
 * TEST INFRASTRUCTURE ONLY -- see zj_oracle.h for the rules and the parity-pinning statement.
 *
 * Conventions used to stay literal:
 *   - Rust release-mode integer semantics: + - * wrap at the operand width; `>>` on signed is arithmetic.
 *     w16()/w32() below make every wrap explicit.
 *   - Where the Rust would panic (assert!, unwrap on None, slice index out of range) the function returns
 *     ZJO_ERR_PANIC instead.
 *   - The SIMD functions are written against simd_compat.h so the same source builds with the real
 *     intrinsics (-DZJO_REAL_SIMD -mavx2 -msse4.1) or with the portable emulation.
 */
#include "zj_oracle.h"

#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>

#include "simd_compat.h"

#define PANIC_IF(c) do { if (c) return ZJO_ERR_PANIC; } while (0)

static inline int16_t w16(int32_t x) { return (int16_t)(uint16_t)(uint32_t)x; }
static inline int32_t wmul32(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
static inline int32_t wadd32(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static inline int32_t wsub32(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
static inline int32_t wshl32(int32_t a, int s) { return (int32_t)((uint32_t)a << s); }
static inline int16_t wmul16(int16_t a, int16_t b) { return w16((int32_t)((uint32_t)(int32_t)a * (uint32_t)(int32_t)b)); }

int zjo_uses_real_simd(void)
{
#ifdef ZJO_REAL_SIMD
    return 1;
#else
    return 0;
#endif
}

/* =====================================================================================================
 * IDCT, SCALAR variant -- src/idct/scalar.rs:19-282
 * ===================================================================================================== */

#define SCALE_BITS (512 + 65536 + (128 << 17)) /* idct/scalar.rs:6, idct/avx2.rs:33 */

static inline int32_t dequantize(int16_t a, int32_t b) { return wmul32((int32_t)a, b); } /* scalar.rs:308-311 */
static inline int32_t fsh(int32_t x) { return wshl32(x, 12); }                            /* scalar.rs:294-297 */
static inline int16_t clamp_i32(int32_t a) { return (int16_t)(a < 0 ? 0 : (a > 255 ? 255 : a)); } /* scalar.rs:302-305 */

int zjo_idct_scalar(const int16_t *vector, size_t len, const int32_t *qt, size_t stride,
                    size_t samp_factors, size_t v_samp, int16_t *out_all)
{
    int32_t tmp[64];
    PANIC_IF(samp_factors == 0);
    const size_t chunks = len * v_samp / samp_factors; /* scalar.rs:30 */
    PANIC_IF(chunks == 0);                             /* chunks_exact(0) panics */

    for (size_t c0 = 0; c0 + chunks <= len; c0 += chunks) { /* scalar.rs:32-34 */
        const int16_t *in_vector = vector + c0;
        int16_t *out_vector = out_all + c0;
        size_t pos = 0, x = 0;

        for (size_t b0 = 0; b0 + 64 <= chunks; b0 += 64) { /* scalar.rs:40 */
            const int16_t *v = in_vector + b0;
            int all_zero = 1;
            for (int i = 1; i < 64; i++) if (v[i] != 0) { all_zero = 0; break; } /* scalar.rs:45 */

            if (all_zero) {
                /* scalar.rs:48: ((vector[0].wrapping_mul(qt[0] as i16)) >> 3) + 128, i16, NOT clamped */
                int16_t coeff = w16((int32_t)(int16_t)(wmul16(v[0], (int16_t)qt[0]) >> 3) + 128);
                for (int r = 0; r < 8; r++) { /* scalar.rs:65-73 */
                    PANIC_IF(pos + 8 > chunks);
                    for (int k = 0; k < 8; k++) out_vector[pos + k] = coeff;
                    pos += stride;
                }
            } else {
                /* pass 1: down the columns, scalar.rs:79-167 */
                for (int ptr = 0; ptr < 8; ptr++) {
                    int32_t p2 = dequantize(v[ptr + 16], qt[ptr + 16]);
                    int32_t p3 = dequantize(v[ptr + 48], qt[ptr + 48]);
                    int32_t p1 = wmul32(wadd32(p2, p3), 2217);
                    int32_t t2 = wadd32(p1, wmul32(p3, -7567));
                    int32_t t3 = wadd32(p1, wmul32(p2, 3135));
                    p2 = dequantize(v[ptr], qt[ptr]);
                    p3 = dequantize(v[32 + ptr], qt[32 + ptr]);
                    int32_t t0 = fsh(wadd32(p2, p3));
                    int32_t t1 = fsh(wsub32(p2, p3));
                    int32_t x0 = wadd32(wadd32(t0, t3), 512);
                    int32_t x3 = wadd32(wsub32(t0, t3), 512);
                    int32_t x1 = wadd32(wadd32(t1, t2), 512);
                    int32_t x2 = wadd32(wsub32(t1, t2), 512);
                    /* odd part */
                    t0 = dequantize(v[ptr + 56], qt[ptr + 56]);
                    t1 = dequantize(v[ptr + 40], qt[ptr + 40]);
                    t2 = dequantize(v[ptr + 24], qt[ptr + 24]);
                    t3 = dequantize(v[ptr + 8], qt[ptr + 8]);
                    p3 = wadd32(t0, t2);
                    int32_t p4 = wadd32(t1, t3);
                    p1 = wadd32(t0, t3);
                    p2 = wadd32(t1, t2);
                    int32_t p5 = wmul32(wadd32(p3, p4), 4816);
                    t0 = wmul32(t0, 1223);
                    t1 = wmul32(t1, 8410);
                    t2 = wmul32(t2, 12586);
                    t3 = wmul32(t3, 6149);
                    p1 = wadd32(p5, wmul32(p1, -3685));
                    p2 = wadd32(p5, wmul32(p2, -10497));
                    p3 = wmul32(p3, -8034);
                    p4 = wmul32(p4, -1597);
                    t3 = wadd32(t3, wadd32(p1, p4));
                    t2 = wadd32(t2, wadd32(p2, p3));
                    t1 = wadd32(t1, wadd32(p2, p4));
                    t0 = wadd32(t0, wadd32(p1, p3));
                    tmp[ptr] = wadd32(x0, t3) >> 10;
                    tmp[ptr + 8] = wadd32(x1, t2) >> 10;
                    tmp[ptr + 16] = wadd32(x2, t1) >> 10;
                    tmp[ptr + 24] = wadd32(x3, t0) >> 10;
                    tmp[ptr + 32] = wsub32(x3, t0) >> 10;
                    tmp[ptr + 40] = wsub32(x2, t1) >> 10;
                    tmp[ptr + 48] = wsub32(x1, t2) >> 10;
                    tmp[ptr + 56] = wsub32(x0, t3) >> 10;
                }
                /* pass 2: along the rows, scalar.rs:170-274 */
                for (int i = 0; i < 64; i += 8) {
                    int32_t p2 = tmp[i + 2];
                    int32_t p3 = tmp[i + 6];
                    int32_t p1 = wmul32(wadd32(p2, p3), 2217);
                    int32_t t2 = wadd32(p1, wmul32(p3, -7567));
                    int32_t t3 = wadd32(p1, wmul32(p2, 3135));
                    p2 = tmp[i];
                    p3 = tmp[i + 4];
                    int32_t t0 = fsh(wadd32(p2, p3));
                    int32_t t1 = fsh(wsub32(p2, p3));
                    int32_t x0 = wadd32(wadd32(t0, t3), SCALE_BITS);
                    int32_t x3 = wadd32(wsub32(t0, t3), SCALE_BITS);
                    int32_t x1 = wadd32(wadd32(t1, t2), SCALE_BITS);
                    int32_t x2 = wadd32(wsub32(t1, t2), SCALE_BITS);
                    t0 = tmp[i + 7];
                    t1 = tmp[i + 5];
                    t2 = tmp[i + 3];
                    t3 = tmp[i + 1];
                    p3 = wadd32(t0, t2);
                    int32_t p4 = wadd32(t1, t3);
                    p1 = wadd32(t0, t3);
                    p2 = wadd32(t1, t2);
                    int32_t p5 = wmul32(wadd32(p3, p4), 4816); /* f2f(1.175875602) = 4816, scalar.rs:224,287-290 */
                    t0 = wmul32(t0, 1223);
                    t1 = wmul32(t1, 8410);
                    t2 = wmul32(t2, 12586);
                    t3 = wmul32(t3, 6149);
                    p1 = wadd32(p5, wmul32(p1, -3685));
                    p2 = wadd32(p5, wmul32(p2, -10497));
                    p3 = wmul32(p3, -8034);
                    p4 = wmul32(p4, -1597);
                    t3 = wadd32(t3, wadd32(p1, p4));
                    t2 = wadd32(t2, wadd32(p2, p3));
                    t1 = wadd32(t1, wadd32(p2, p4));
                    t0 = wadd32(t0, wadd32(p1, p3));
                    PANIC_IF(pos + 8 > chunks); /* get_mut(pos..pos+8).unwrap(), scalar.rs:249-253 */
                    int16_t *o = out_vector + pos;
                    o[0] = clamp_i32(wadd32(x0, t3) >> 17);
                    o[1] = clamp_i32(wadd32(x1, t2) >> 17);
                    o[2] = clamp_i32(wadd32(x2, t1) >> 17);
                    o[3] = clamp_i32(wadd32(x3, t0) >> 17);
                    o[4] = clamp_i32(wsub32(x3, t0) >> 17);
                    o[5] = clamp_i32(wsub32(x2, t1) >> 17);
                    o[6] = clamp_i32(wsub32(x1, t2) >> 17);
                    o[7] = clamp_i32(wsub32(x0, t3) >> 17);
                    pos += stride;
                }
            }
            x += 8;   /* scalar.rs:277-278 */
            pos = x;
        }
    }
    return ZJO_OK;
}

/* =====================================================================================================
 * IDCT, X86 variant -- src/idct/avx2.rs:64-398 (+ YmmRegister operators, src/unsafe_utils.rs:16-158)
 * ===================================================================================================== */

#define SHUF(z, y, x, w) (((z) << 6) | ((y) << 4) | ((x) << 2) | (w)) /* avx2.rs:491-494 */

/* avx2.rs:422-486 */
static void transpose8(zv256 r[8])
{
    zv256 w0, w1, w2, w3, w4, w5, w6, w7, x0, x1, x2, x3, x4, x5, x6, x7, va, vb;
#define MERGE_EPI32(v0, v1, o2, o3) /* avx2.rs:427-437 */ \
    va = zv256_permute4x64_epi64(v0, SHUF(3, 1, 2, 0)); vb = zv256_permute4x64_epi64(v1, SHUF(3, 1, 2, 0)); \
    o2 = zv256_unpacklo_epi32(va, vb); o3 = zv256_unpackhi_epi32(va, vb);
#define MERGE_EPI64(v0, v1, o2, o3) /* avx2.rs:439-449 */ \
    va = zv256_permute4x64_epi64(v0, SHUF(3, 1, 2, 0)); vb = zv256_permute4x64_epi64(v1, SHUF(3, 1, 2, 0)); \
    o2 = zv256_unpacklo_epi64(va, vb); o3 = zv256_unpackhi_epi64(va, vb);
#define MERGE_SI128(v0, v1, o2, o3) /* avx2.rs:451-457 */ \
    o2 = zv256_permute2x128(v0, v1, SHUF(0, 2, 0, 0)); o3 = zv256_permute2x128(v0, v1, SHUF(0, 3, 0, 1));
    MERGE_EPI32(r[0], r[1], w0, w1)
    MERGE_EPI32(r[2], r[3], w2, w3)
    MERGE_EPI32(r[4], r[5], w4, w5)
    MERGE_EPI32(r[6], r[7], w6, w7)
    MERGE_EPI64(w0, w2, x0, x1)
    MERGE_EPI64(w1, w3, x2, x3)
    MERGE_EPI64(w4, w6, x4, x5)
    MERGE_EPI64(w5, w7, x6, x7)
    MERGE_SI128(x0, x4, r[0], r[1])
    MERGE_SI128(x1, x5, r[2], r[3])
    MERGE_SI128(x2, x6, r[4], r[5])
    MERGE_SI128(x3, x7, r[6], r[7])
#undef MERGE_EPI32
#undef MERGE_EPI64
#undef MERGE_SI128
}

#define YADD(a, b) zv256_add_epi32((a), (b))
#define YSUB(a, b) zv256_sub_epi32((a), (b))
#define YMULI(a, k) zv256_mullo_epi32((a), zv256_set1_epi32(k))
#define YADDI(a, k) zv256_add_epi32((a), zv256_set1_epi32(k))

/* dct_pass! macro, avx2.rs:251-331.  `scale` must be a literal for the real srai intrinsic. */
#define DCT_PASS(row, SCALE_B, scale) do { \
    zv256 p1 = YMULI(YADD(row[2], row[6]), 2217); \
    zv256 t2 = YADD(p1, YMULI(row[6], -7567)); \
    zv256 t3 = YADD(p1, YMULI(row[2], 3135)); \
    zv256 t0 = zv256_slli_epi32(YADD(row[0], row[4]), 12); \
    zv256 t1 = zv256_slli_epi32(YSUB(row[0], row[4]), 12); \
    zv256 x0 = YADDI(YADD(t0, t3), SCALE_B); \
    zv256 x3 = YADDI(YSUB(t0, t3), SCALE_B); \
    zv256 x1 = YADDI(YADD(t1, t2), SCALE_B); \
    zv256 x2 = YADDI(YSUB(t1, t2), SCALE_B); \
    zv256 p3 = YADD(row[7], row[3]); \
    zv256 p4 = YADD(row[5], row[1]); \
    p1 = YADD(row[7], row[1]); \
    zv256 p2 = YADD(row[5], row[3]); \
    zv256 p5 = YMULI(YADD(p3, p4), 4816); \
    t0 = YMULI(row[7], 1223); \
    t1 = YMULI(row[5], 8410); \
    t2 = YMULI(row[3], 12586); \
    t3 = YMULI(row[1], 6149); \
    p1 = YADD(p5, YMULI(p1, -3685)); \
    p2 = YADD(p5, YMULI(p2, -10497)); \
    p3 = YMULI(p3, -8034); \
    p4 = YMULI(p4, -1597); \
    t3 = YADD(t3, YADD(p1, p4)); \
    t2 = YADD(t2, YADD(p2, p3)); \
    t1 = YADD(t1, YADD(p2, p4)); \
    t0 = YADD(t0, YADD(p1, p3)); \
    row[0] = zv256_srai_epi32(YADD(x0, t3), scale); \
    row[1] = zv256_srai_epi32(YADD(x1, t2), scale); \
    row[2] = zv256_srai_epi32(YADD(x2, t1), scale); \
    row[3] = zv256_srai_epi32(YADD(x3, t0), scale); \
    row[4] = zv256_srai_epi32(YSUB(x3, t0), scale); \
    row[5] = zv256_srai_epi32(YSUB(x2, t1), scale); \
    row[6] = zv256_srai_epi32(YSUB(x1, t2), scale); \
    row[7] = zv256_srai_epi32(YSUB(x0, t3), scale); \
} while (0)

/* clamp_avx, avx2.rs:402-413 and color_convert/avx.rs:403-414 */
static inline zv256 clamp_avx(zv256 reg)
{
    zv256 min_s = zv256_set1_epi16(0), max_s = zv256_set1_epi16(255);
    return zv256_min_epi16(zv256_max_epi16(reg, min_s), max_s);
}

int zjo_idct_avx2(const int16_t *coeff, size_t len, const int32_t *qt, size_t stride,
                  size_t samp_factors, size_t v_samp, int16_t *tmp_vector)
{
    zv256 qt_row[8];
    for (int i = 0; i < 8; i++) qt_row[i] = zv256_loadu(qt + 8 * i); /* avx2.rs:73-87 */
    PANIC_IF(samp_factors == 0);
    const size_t chunks = len * v_samp / samp_factors; /* avx2.rs:89 */
    PANIC_IF(chunks == 0);

    for (size_t c0 = 0; c0 + chunks <= len; c0 += chunks) { /* avx2.rs:92-94 */
        const int16_t *in_vector = coeff + c0;
        int16_t *out_vector = tmp_vector + c0;
        size_t pos = 0, x = 0;
        for (size_t b0 = 0; b0 + 64 <= chunks; b0 += 64) { /* avx2.rs:100 */
            const int16_t *vector = in_vector + b0;
            zv128 rw[8];
            for (int i = 0; i < 8; i++) rw[i] = zv128_loadu(vector + 8 * i); /* avx2.rs:107-121 */
            {
                /* avx2.rs:136-157: OR of elements 1..=63, PTEST */
                zv128 zero_test = zv128_loadu(vector + 1);
                for (int i = 1; i < 8; i++) zero_test = zv128_or(rw[i], zero_test);
                if (zv128_test_all_zeros(zero_test, zero_test) == 1) {
                    /* avx2.rs:163-167: ((v0.wrapping_mul(qt0 as i16) >> 3) + 128).max(0).min(255) */
                    int16_t val = w16((int32_t)(int16_t)(wmul16(vector[0], (int16_t)qt[0]) >> 3) + 128);
                    val = val < 0 ? 0 : (val > 255 ? 255 : val);
                    zv128 idct_value = zv128_set1_epi16(val);
                    for (int r = 0; r < 8; r++) { /* avx2.rs:182-190 */
                        PANIC_IF(pos + 8 > chunks);
                        zv128_storeu(out_vector + pos, idct_value);
                        pos += stride;
                    }
                    x += 8;
                    pos = x;
                    continue;
                }
            }
            zv256 row[8];
            for (int i = 0; i < 8; i++) row[i] = zv256_cvtepi16_epi32(rw[i]);         /* avx2.rs:202-232 */
            for (int i = 0; i < 8; i++) row[i] = zv256_mullo_epi32(row[i], qt_row[i]); /* avx2.rs:235-249 */

            transpose8(row);              /* avx2.rs:333-336 */
            DCT_PASS(row, 512, 10);       /* avx2.rs:339 */
            transpose8(row);              /* avx2.rs:341-344 */
            DCT_PASS(row, SCALE_BITS, 17);/* avx2.rs:347 */

            /* permute_store!, avx2.rs:355-391 */
            for (int i = 0; i < 8; i += 2) {
                zv256 a = zv256_packs_epi32(row[i], row[i + 1]);
                zv256 b = clamp_avx(a);
                zv256 c = zv256_permute4x64_epi64(b, SHUF(3, 1, 2, 0));
                PANIC_IF(pos + 8 > chunks);
                zv128_storeu(out_vector + pos, zv256_extract128(c, 0));
                pos += stride;
                PANIC_IF(pos + 8 > chunks);
                zv128_storeu(out_vector + pos, zv256_extract128(c, 1));
                pos += stride;
            }
            x += 8; /* avx2.rs:393-394 */
            pos = x;
        }
    }
    return ZJO_OK;
}

/* =====================================================================================================
 * Up-samplers
 * ===================================================================================================== */

/* src/upsampler/scalar.rs:5-60 */
int zjo_upsample_horizontal_scalar(const int16_t *input, size_t n, int16_t *out, size_t out_len)
{
    memset(out, 0, out_len * sizeof(int16_t));
    PANIC_IF(!(out_len > 4 && n > 2)); /* scalar.rs:9-12 */
    out[0] = input[0];
    out[1] = (int16_t)(w16(w16(w16(input[0] * 3) + input[1]) + 2) >> 2); /* scalar.rs:15 */
    /* scalar.rs:30-42: out[2..].chunks_exact_mut(2) zipped with input.windows(3) */
    size_t n_out_chunks = (out_len - 2) / 2, n_win = n - 2;
    size_t cnt = n_out_chunks < n_win ? n_out_chunks : n_win;
    for (size_t k = 0; k < cnt; k++) {
        int16_t sample = w16(w16(3 * input[k + 1]) + 2);
        out[2 + 2 * k] = (int16_t)(w16(sample + input[k]) >> 2);
        out[3 + 2 * k] = (int16_t)(w16(sample + input[k + 2]) >> 2);
    }
    /* scalar.rs:46-57 */
    size_t ol = out_len - 2, il = n - 2;
    out[ol] = (int16_t)(w16(w16(w16(3 * input[il]) + input[il + 1]) + 2) >> 2);
    out[ol + 1] = input[il + 1];
    return ZJO_OK;
}

/* src/upsampler/sse.rs:24-134 */
int zjo_upsample_horizontal_sse(const int16_t *input, size_t n, int16_t *out, size_t out_len)
{
    memset(out, 0, out_len * sizeof(int16_t));
    PANIC_IF(!(out_len > 8 && n > 5)); /* sse.rs:33 */
#define T3(a, b) ((int16_t)(w16(w16(w16((a) * 3) + (b)) + 2) >> 2))
    out[0] = input[0];                  /* sse.rs:41-55 */
    out[1] = T3(input[0], input[1]);
    out[2] = T3(input[1], input[0]);
    out[3] = T3(input[1], input[2]);
    out[4] = T3(input[2], input[1]);
    out[5] = T3(input[2], input[3]);
    out[6] = T3(input[3], input[2]);
    out[7] = T3(input[3], input[4]);
    size_t inl = n;
    size_t hi = (inl >> 2) - 1; /* n > 5 so inl>>2 >= 1 */
    for (size_t i = 1; i < hi; i++) { /* sse.rs:69-108 */
        size_t pos = i << 2;
        zv128 yn = zv128_loadl64(input + pos);
        yn = zv128_unpacklo_epi16(yn, yn);
        zv128 v = zv128_loadl64(input + pos - 1);
        zv128 y = zv128_loadl64(input + pos + 1);
        zv128 even = zv128_unpacklo_epi16(v, v);
        zv128 odd = zv128_unpacklo_epi16(y, y);
        zv128 nn = zv128_blend_epi16(even, odd, 0xAA);
        zv128 an = zv128_add_epi16(zv128_slli_epi16(yn, 1), yn);
        zv128 bn = zv128_add_epi16(nn, zv128_set1_epi16(2));
        zv128 cn = zv128_srai_epi16(zv128_add_epi16(an, bn), 2);
        PANIC_IF(i * 8 + 8 > out_len);
        zv128_storeu(out + i * 8, cn);
    }
    size_t ol = out_len - 8, il = n - 4; /* sse.rs:111-131 */
    out[ol + 0] = T3(input[il], input[il - 1]);
    out[ol + 1] = T3(input[il], input[il + 1]);
    out[ol + 2] = T3(input[il + 1], input[il]);
    out[ol + 3] = T3(input[il + 1], input[il + 1]);
    out[ol + 4] = T3(input[il + 2], input[il + 2]);
    out[ol + 5] = T3(input[il + 2], input[il + 1]);
    out[ol + 6] = T3(input[il + 2], input[il + 3]);
    out[ol + 7] = input[il + 3];
    return ZJO_OK;
}

/* src/upsampler/scalar.rs:64-147 -- the iterator dance is kept as explicit cursors */
int zjo_upsample_vertical(const int16_t *input, size_t n, int16_t *out, size_t out_len)
{
    memset(out, 0, out_len * sizeof(int16_t));
    size_t stride = n >> 3;  /* scalar.rs:73 */
    PANIC_IF(stride == 0);   /* chunks_exact(0) */
    size_t n_rows = n / stride;
    size_t near_next = 0, far_next = 0; /* the two chunks_exact iterators */
    PANIC_IF(n_rows == 0);
    const int16_t *rw_n = input + stride * near_next++; /* scalar.rs:84 */
    const int16_t *rw_f = input + stride * far_next++;  /* scalar.rs:86 */
    const int16_t *previous;
    size_t i = 0;
    int next_row = 1;
    for (int it = 0; it < 8; it++) {
        PANIC_IF(out_len < i || out_len - i < stride); /* split_at_mut(stride), scalar.rs:111 */
        int16_t *out_near = out + i;
        int16_t *remainder = out + i + stride;
        size_t rem_len = out_len - i - stride;
        size_t cnt = stride < rem_len ? stride : rem_len; /* zip stops at the shortest */
        for (size_t k = 0; k < cnt; k++) { /* scalar.rs:113-127 */
            int16_t near = rw_n[k], far = rw_f[k];
            out_near[k] = (int16_t)(w16(w16(w16(near * 3) + far) + 2) >> 2);
            remainder[k] = (int16_t)(w16(w16(w16(far * 3) + near) + 2) >> 2);
        }
        i += stride * 2;
        previous = rw_n;
        rw_n = (near_next < n_rows) ? input + stride * near_next++ : previous; /* scalar.rs:134 */
        rw_f = (far_next < n_rows) ? input + stride * far_next++ : rw_n;       /* scalar.rs:136 */
        if (next_row) {                                                         /* scalar.rs:140-144 */
            rw_f = (far_next < n_rows) ? input + stride * far_next++ : rw_n;
            next_row = 0;
        }
    }
    return ZJO_OK;
}

/* src/upsampler/scalar.rs:148-166 */
int zjo_upsample_hv_scalar(const int16_t *input, size_t n, int16_t *out, size_t out_len)
{
    int16_t *first = (int16_t *)malloc((n > 0 ? n * 2 : 1) * sizeof(int16_t));
    if (!first) return ZJO_ERR_ARG;
    int rc = zjo_upsample_vertical(input, n, first, n * 2);
    if (rc == ZJO_OK) rc = zjo_upsample_horizontal_scalar(first, n * 2, out, out_len);
    free(first);
    return rc;
}

/* src/upsampler/avx2.rs:29-342 */
static int upsample_hv_avx(const int16_t *input, size_t n, int16_t *output, size_t output_len)
{
    memset(output, 0, output_len * sizeof(int16_t));
    size_t stride = 0;
    int modify_stride = 1;
    size_t pos = 0, output_position = 0;
    PANIC_IF(n <= 16);
    int16_t prev = (int16_t)(w16(w16(w16(3 * input[0]) + input[stride]) + 2) >> 2);                   /* avx2.rs:67 */
    int16_t pixel_far = (int16_t)(w16(w16(w16(3 * input[pos + 16]) + input[pos + stride + 16]) + 2) >> 2); /* :68 */
    const zv256 three = zv256_set1_epi16(3), two = zv256_set1_epi16(2);

/* pack_shuffle!, avx2.rs:73-82 */
#define PACK_SHUFFLE(x, y, value) do { \
    zv256 v_ = zv256_permute2x128((x), (x), (value)); \
    zv256 rwn_hi = zv256_unpackhi_epi16(v_, v_); \
    zv256 rwn_lo = zv256_unpacklo_epi16(v_, v_); \
    (y) = zv256_permute2x128(rwn_lo, rwn_hi, 0x30); \
} while (0)

/* upsample_horizontal!, avx2.rs:84-201 */
#define UPSAMPLE_HORIZONTAL(row, ostride) do { \
    zv256 next_arr = zv256_alignr_epi8(zv256_permute2x128((row), (row), SHUF(2, 0, 0, 1)), (row), 2); \
    zv256 prev_arr = zv256_alignr_epi8((row), zv256_permute2x128((row), (row), SHUF(0, 0, 2, 0)), 14); \
    prev_arr = zv256_insert_epi16(prev_arr, prev, 0); \
    next_arr = zv256_insert_epi16(next_arr, pixel_far, 15); \
    zv256 near_lo, near_hi, prev_lo, prev_hi, next_lo, next_hi; \
    PACK_SHUFFLE((row), near_lo, 0); \
    PACK_SHUFFLE((row), near_hi, 0x33); \
    PACK_SHUFFLE(prev_arr, prev_lo, 0); \
    PACK_SHUFFLE(prev_arr, prev_hi, 0x33); \
    PACK_SHUFFLE(next_arr, next_lo, 0); \
    PACK_SHUFFLE(next_arr, next_hi, 0x33); \
    zv256 nn = zv256_blend_epi16(prev_lo, next_lo, 0xAA); \
    zv256 an = zv256_mullo_epi16(near_lo, three); \
    zv256 bn = zv256_add_epi16(nn, two); \
    zv256 cn = zv256_srai_epi16(zv256_add_epi16(an, bn), 2); \
    PANIC_IF(output_position + (ostride) + 16 > output_len); \
    zv256_storeu(output + output_position + (ostride), cn); \
    nn = zv256_blend_epi16(prev_hi, next_hi, 0xAA); \
    an = zv256_mullo_epi16(near_hi, three); \
    bn = zv256_add_epi16(nn, two); \
    cn = zv256_srai_epi16(zv256_add_epi16(an, bn), 2); \
    PANIC_IF(output_position + (ostride) + 32 > output_len); \
    zv256_storeu(output + output_position + (ostride) + 16, cn); \
} while (0)

    const size_t len = output_len / 16;                        /* avx2.rs:221 */
    const size_t end = (n >> 7) > 0 ? (n >> 7) - 1 : 0;        /* avx2.rs:222 */
    const size_t v = (output_len / 16) > end * 32 ? (output_len / 16) - end * 32 : 0; /* avx2.rs:224 */

    for (int j = 0; j < 8; j++) { /* avx2.rs:229 */
        for (size_t t = 0; t < end; t++) { /* avx2.rs:235 */
            PANIC_IF(pos > n || pos + stride > n); /* input[pos..], input[pos+stride..] */
            PANIC_IF(pos + stride + 16 > n);       /* the raw 256-bit loads must stay inside (UB otherwise) */
            zv256 load_near = zv256_loadu(input + pos);
            zv256 load_far = zv256_loadu(input + pos + stride);
            /* upsample_vertical!, avx2.rs:203-217 */
            zv256 t2 = zv256_add_epi16(load_far, two);
            zv256 t1 = zv256_mullo_epi16(load_near, three);
            zv256 row_near = zv256_srai_epi16(zv256_add_epi16(t1, t2), 2);
            zv256 t3 = zv256_add_epi16(load_near, two);
            zv256 t4 = zv256_mullo_epi16(load_far, three);
            zv256 row_far = zv256_srai_epi16(zv256_add_epi16(t3, t4), 2);

            UPSAMPLE_HORIZONTAL(row_near, 0);   /* avx2.rs:252 */
            UPSAMPLE_HORIZONTAL(row_far, len);  /* avx2.rs:254 */
            output_position += 32;
            pos += 16;
            PANIC_IF(!(n >= pos + stride + 16)); /* avx2.rs:262 */
            /* avx2.rs:264-270: 3 * (a.wrapping_add(b).wrapping_add(2)) >> 2  (method calls bind before `*`) */
            {
                int16_t a = pos < n ? input[pos] : 0;
                int16_t b = pos + stride < n ? input[pos + stride] : 0;
                prev = (int16_t)(w16(3 * (int32_t)w16(w16(a + b) + 2)) >> 2);
                int16_t c = pos + 16 < n ? input[pos + 16] : 0;
                int16_t d = pos + stride + 16 < n ? input[pos + stride + 16] : 0;
                pixel_far = (int16_t)(w16(3 * (int32_t)w16(w16(c + d) + 2)) >> 2);
            }
        }
        /* scalar tail, avx2.rs:277-307 */
        PANIC_IF(output_position + v > output_len);   /* split_at_mut */
        int16_t *a_ptr = output;
        size_t a_len = output_position + v;
        int16_t *b_ptr = output + a_len;
        size_t b_len = output_len - a_len;
        size_t c = a_len - v;
        int16_t *unwritten = a_ptr + c;               /* v elements */
        PANIC_IF(len < v || len > b_len);
        int16_t *unwritten_stride = b_ptr + (len - v); /* v elements */

        PANIC_IF(pos < v / 2 + 2 || pos > n);         /* input[pos-v/2-2..pos] */
        {
            const int16_t *w = input + (pos - v / 2 - 2);
            size_t wl = v / 2 + 2;
            size_t nwin = wl >= 3 ? wl - 2 : 0, nch = v / 2;
            size_t cnt = nwin < nch ? nwin : nch;
            for (size_t k = 0; k < cnt; k++) {
                int16_t sample = w16(w16(3 * w[k + 1]) + 2);
                unwritten[2 * k] = (int16_t)(w16(sample + w[k]) >> 2);
                unwritten[2 * k + 1] = (int16_t)(w16(sample + w[k + 2]) >> 2);
            }
        }
        PANIC_IF(v == 0 || pos >= n);
        unwritten[v - 1] = input[pos];                /* avx2.rs:295 */
        PANIC_IF(pos + stride < v / 2 + 2 || pos + stride > n); /* input[pos-v/2+stride-2..pos+stride] */
        {
            const int16_t *w = input + (pos - v / 2 + stride - 2);
            size_t wl = v / 2 + 2;
            size_t nwin = wl >= 3 ? wl - 2 : 0, nch = v / 2;
            size_t cnt = nwin < nch ? nwin : nch;
            for (size_t k = 0; k < cnt; k++) {
                int16_t sample = w16(w16(3 * w[k + 1]) + 2);
                unwritten_stride[2 * k] = (int16_t)(w16(sample + w[k]) >> 2);
                unwritten_stride[2 * k + 1] = (int16_t)(w16(sample + w[k + 2]) >> 2);
            }
        }
        PANIC_IF(pos + stride >= n);
        unwritten_stride[v - 1] = input[pos + stride]; /* avx2.rs:307 */

        output_position += len + v; /* avx2.rs:311-312 */
        pos += v / 2;
        if (modify_stride) { stride = n / 8; modify_stride = 0; } /* avx2.rs:314-319 */
        if (j == 6) stride = 0;                                   /* avx2.rs:320-328 */
        /* avx2.rs:330-338 */
        PANIC_IF(output_position > output_len || output_position < len + 4);
        PANIC_IF(output_position - len + 1 >= output_len);
        output[output_position - len] = output[output_position - len + 1];
        output[output_position - len - 2] = output[output_position - len - 4];
        output[output_position - len - 1] = output[output_position - len - 3];
        output[output_position - 2] = output[output_position - 4];
        output[output_position - 1] = output[output_position - 3];
    }
    return ZJO_OK;
#undef PACK_SHUFFLE
#undef UPSAMPLE_HORIZONTAL
}

/* src/upsampler/avx2.rs:14-22 */
int zjo_upsample_hv_simd(const int16_t *input, size_t n, int16_t *out, size_t out_len)
{
    if (n < 500) return zjo_upsample_hv_scalar(input, n, out, out_len);
    return upsample_hv_avx(input, n, out, out_len);
}

/* =====================================================================================================
 * Colour conversion
 * ===================================================================================================== */

static inline uint8_t clamp_u8_i16(int16_t a) { return (uint8_t)(a < 0 ? 0 : (a > 255 ? 255 : a)); }

/* ycbcr_to_rgb_avx2 (color_convert/avx.rs:67-192) and ycbcr_to_rgb_16_scalar (color_convert/scalar.rs:52-89):
 * both compute, per pixel in wrapping i16,
 *   cb -= 128; cr -= 128; r = y + ((45*cr) >> 5); g = y - ((11*cb + 23*cr) >> 5); b = y + ((113*cb) >> 6)
 * clamp to [0,255] and store 48 interleaved bytes at out[*pos..*pos+48]; *pos += 48.
 * (X86 on an AVX2 host uses the AVX2 one; the `variant` argument selects which source is being restated but
 * the arithmetic is identical -- kept so a future divergence has a switch.) */
int zjo_ycbcr_to_rgb_16(const int16_t *y, const int16_t *cb, const int16_t *cr, uint8_t *out,
                        size_t out_len, size_t *pos, int variant)
{
    (void)variant;
    PANIC_IF(*pos + 48 > out_len); /* .get_mut(..).expect("Slice to small cannot write") */
    uint8_t *opt = out + *pos;
    for (int i = 0; i < 16; i++) {
        int16_t crr = w16(cr[i] - 128);
        int16_t cbb = w16(cb[i] - 128);
        int16_t r = w16(y[i] + (int16_t)(wmul16(45, crr) >> 5));
        int16_t g = w16(y[i] - (int16_t)(w16(wmul16(11, cbb) + wmul16(23, crr)) >> 5));
        int16_t b = w16(y[i] + (int16_t)(wmul16(113, cbb) >> 6));
        opt[3 * i] = clamp_u8_i16(r);
        opt[3 * i + 1] = clamp_u8_i16(g);
        opt[3 * i + 2] = clamp_u8_i16(b);
    }
    *pos += 48;
    return ZJO_OK;
}

/* color_convert/scalar.rs:91-114 */
int zjo_ycbcr_to_grayscale(const int16_t *y, size_t len, size_t width, uint8_t *output, size_t out_len)
{
    PANIC_IF(width == 0);
    size_t width_mcu = len / width;          /* scalar.rs:97 */
    PANIC_IF(width_mcu == 0);
    size_t width_chunk = len / width_mcu;    /* scalar.rs:99 */
    PANIC_IF(width_chunk == 0);
    size_t start = 0, end = width;
    for (size_t c0 = 0; c0 + width_chunk <= len; c0 += width_chunk) { /* scalar.rs:105 */
        PANIC_IF(width > width_chunk);       /* &chunk[0..width] */
        PANIC_IF(end > out_len);             /* output[start..end] */
        for (size_t k = 0; k < width; k++) output[start + k] = (uint8_t)(uint16_t)y[c0 + k]; /* `*x as u8` */
        start += width;
        end += width;
    }
    return ZJO_OK;
}

/* color_convert/scalar.rs:119-169 */
int zjo_ycbcr_to_ycbcr(const int16_t *const ch[3], size_t len0, size_t width, size_t h_samp,
                       size_t v_samp, uint8_t *output, size_t out_len)
{
    size_t mcu_chunks = len0 / (h_samp * v_samp);
    size_t stride = width * 3, start = 0, end = width * 3;
    size_t width_chunk = mcu_chunks >> 3;
    PANIC_IF(width_chunk == 0);
    for (size_t c0 = 0; c0 + width_chunk <= len0; c0 += width_chunk) {
        PANIC_IF(stride > width_chunk * 3); /* &temp_output[0..stride] */
        PANIC_IF(end > out_len);
        for (size_t k = 0; k < width; k++) {
            output[start + 3 * k] = (uint8_t)(uint16_t)ch[0][c0 + k];
            output[start + 3 * k + 1] = (uint8_t)(uint16_t)ch[1][c0 + k];
            output[start + 3 * k + 2] = (uint8_t)(uint16_t)ch[2][c0 + k];
        }
        start += stride;
        end += stride;
    }
    return ZJO_OK;
}

static size_t cs_components(uint32_t cs) /* ColorSpace::num_components, misc.rs:108-118 */
{
    switch (cs) {
    case ZJ_CS_RGB: case ZJ_CS_YCBCR: return 3;
    case ZJ_CS_CMYK: case ZJ_CS_RGBA: case ZJ_CS_RGBX: case ZJ_CS_YCCK: return 4;
    case ZJ_CS_GRAYSCALE: return 1;
    default: return 0;
    }
}

/* worker.rs:143-251 */
static int color_convert_ycbcr(int16_t *const mcu_block[3], size_t len0, size_t width, size_t h_samp,
                               size_t v_samp, uint32_t out_cs, int variant, uint8_t *output, size_t output_len)
{
    size_t nc = cs_components(out_cs);
    size_t mcu_chunks = len0 / (h_samp * v_samp); /* worker.rs:148 */
    size_t width_chunk = mcu_chunks >> 3;         /* worker.rs:150 */
    size_t stride = width * nc;
    size_t start = 0, end = stride;
    uint8_t temp[64];
    size_t temp_len = 16 * nc;
    PANIC_IF(width_chunk == 0);
    if (width < 16) memset(temp, 0, sizeof(temp)); /* worker.rs:158-163 */

    /* the three chunks_exact iterators have the same length for valid input (all planes are Y-sized) */
    for (size_t c0 = 0; c0 + width_chunk <= len0; c0 += width_chunk) {
        const int16_t *y_width = mcu_block[0] + c0, *cb_width = mcu_block[1] + c0, *cr_width = mcu_block[2] + c0;
        size_t elements = (width_chunk / 16) > 0 ? (width_chunk / 16) - 1 : 0; /* worker.rs:171 */
        size_t position = 0;
        PANIC_IF(end > output_len);  /* &mut output[start..end] */
        uint8_t *out = output + start;
        size_t out_len = end - start;

        if (width < 16) { /* worker.rs:176-198 */
            int16_t y_out[16] = {0}, cb_out[16] = {0}, cr_out[16] = {0};
            PANIC_IF(width_chunk > 16);
            memcpy(y_out, y_width, width_chunk * 2);
            memcpy(cb_out, cb_width, width_chunk * 2);
            memcpy(cr_out, cr_width, width_chunk * 2);
            size_t zero = 0;
            int rc = zjo_ycbcr_to_rgb_16(y_out, cb_out, cr_out, temp, temp_len, &zero, variant);
            if (rc) return rc;
            PANIC_IF(width * nc > out_len || width * nc > temp_len);
            memcpy(out + position, temp, width * nc);
            start += stride;
            end += stride;
            continue;
        }
        for (size_t k = 0; k < elements; k++) { /* worker.rs:201-214; chunks_exact(16).take(elements) */
            int rc = zjo_ycbcr_to_rgb_16(y_width + 16 * k, cb_width + 16 * k, cr_width + 16 * k, out, out_len, &position, variant);
            if (rc) return rc;
        }
        /* worker.rs:221-223 */
        size_t room = stride > position ? stride - position : 0;
        size_t diff = 64 > room ? 64 - room : 0;
        position = position > diff ? position - diff : 0;
        PANIC_IF(width_chunk < 16); /* rchunks_exact(16).next().unwrap() */
        {
            size_t o = width_chunk - 16;
            int rc = zjo_ycbcr_to_rgb_16(y_width + o, cb_width + o, cr_width + o, out, out_len, &position, variant);
            if (rc) return rc;
        }
        start += stride;
        end += stride;
    }
    return ZJO_OK;
}

/* =====================================================================================================
 * worker::post_process for one strip -- src/worker.rs:32-134
 * ===================================================================================================== */

typedef int (*upsampler_fn)(const int16_t *, size_t, int16_t *, size_t);

/* Decoder::set_upsampling (decoder.rs:468-523) + choose_*_samp_function (upsampler.rs:82-111) */
static upsampler_fn choose_upsampler(size_t h_max, size_t v_max, int variant)
{
    if (h_max == 2 && v_max == 1) return variant == ZJ_VARIANT_X86 ? zjo_upsample_horizontal_sse : zjo_upsample_horizontal_scalar;
    if (h_max == 1 && v_max == 2) return zjo_upsample_vertical;
    if (h_max == 2 && v_max == 2) return variant == ZJ_VARIANT_X86 ? zjo_upsample_hv_simd : zjo_upsample_hv_scalar;
    return NULL;
}

int zjo_post_process(const int16_t *const coeff[3], const size_t len[3], const zj_image *img,
                     uint8_t *output, size_t output_len)
{
    const size_t h_samp = img->comp[0].h_samp, v_samp = img->comp[0].v_samp; /* worker.rs:44-45 */
    const uint32_t in_cs = img->n_comp == 1 ? ZJ_CS_GRAYSCALE : ZJ_CS_YCBCR;
    const size_t in_n = cs_components(in_cs), out_n = cs_components(img->out_cs);
    const size_t x = in_n < out_n ? in_n : out_n; /* worker.rs:59-62 */
    int16_t *unprocessed[3] = {NULL, NULL, NULL};
    size_t ulen[3] = {0, 0, 0};
    int rc = ZJO_OK;

    for (size_t z = 0; z < x && rc == ZJO_OK; z++) { /* worker.rs:66-81 */
        size_t v_samp_idct = z == 0 ? 1 : v_samp;
        unprocessed[z] = (int16_t *)calloc(len[z] ? len[z] : 1, sizeof(int16_t));
        ulen[z] = len[z];
        if (!unprocessed[z]) { rc = ZJO_ERR_ARG; break; }
        if (img->variant == ZJ_VARIANT_X86) /* choose_idct_func, idct.rs:40-61 */
            rc = zjo_idct_avx2(coeff[z], len[z], img->comp[z].qt, img->comp[z].width_stride, h_samp * v_samp, v_samp_idct, unprocessed[z]);
        else
            rc = zjo_idct_scalar(coeff[z], len[z], img->comp[z].qt, img->comp[z].width_stride, h_samp * v_samp, v_samp_idct, unprocessed[z]);
    }
    /* post_process_inner, worker.rs:88-134 */
    if (rc == ZJO_OK && (h_samp != 1 || v_samp != 1)) {
        upsampler_fn up = choose_upsampler(h_samp, v_samp, (int)img->variant);
        for (size_t i = 1; i < x && rc == ZJO_OK; i++) { /* worker.rs:106-109 */
            if (!up) { rc = ZJO_ERR_UNSUPPORTED; break; }
            size_t olen = ulen[0];
            int16_t *o = (int16_t *)malloc((olen ? olen : 1) * sizeof(int16_t));
            if (!o) { rc = ZJO_ERR_ARG; break; }
            rc = up(unprocessed[i], ulen[i], o, olen);
            free(unprocessed[i]);
            unprocessed[i] = o;
            ulen[i] = olen;
        }
    }
    if (rc == ZJO_OK) {
        uint32_t oc = img->out_cs;
        if (oc == ZJ_CS_GRAYSCALE) { /* worker.rs:115-118: (YCbCr | GRAYSCALE, GRAYSCALE) */
            rc = zjo_ycbcr_to_grayscale(unprocessed[0], ulen[0], img->width, output, output_len);
        } else if (in_cs == ZJ_CS_YCBCR && oc == ZJ_CS_YCBCR) { /* worker.rs:120-123 */
            const int16_t *ch[3] = {unprocessed[0], unprocessed[1], unprocessed[2]};
            rc = zjo_ycbcr_to_ycbcr(ch, ulen[0], img->width, h_samp, v_samp, output, output_len);
        } else if (in_cs == ZJ_CS_YCBCR && (oc == ZJ_CS_RGB || oc == ZJ_CS_RGBA || oc == ZJ_CS_RGBX)) { /* worker.rs:125-129 */
            rc = color_convert_ycbcr(unprocessed, ulen[0], img->width, h_samp, v_samp, oc, (int)img->variant, output, output_len);
        } /* worker.rs:131-132: anything else writes nothing */
    }
    for (int z = 0; z < 3; z++) free(unprocessed[z]);
    return rc;
}

/* =====================================================================================================
 * Whole image: strip geometry of mcu.rs:139-226,354-379 and mcu_prog.rs:132-241
 * ===================================================================================================== */

typedef struct {
    size_t n_strips, y_chunk, c_chunk, out_chunk, out_size, nc;
} geometry;

static int compute_geometry(const zj_image *img, geometry *g)
{
    if (!img || (img->n_comp != 1 && img->n_comp != 3)) return ZJO_ERR_ARG;
    if (img->width == 0 || img->height == 0 || img->width > 65535 || img->height > 65535) return ZJO_ERR_ARG;
    size_t nc = cs_components(img->out_cs);
    if (nc == 0) return ZJO_ERR_ARG;
    if (img->variant != ZJ_VARIANT_X86 && img->variant != ZJ_VARIANT_SCALAR) return ZJO_ERR_ARG;
    size_t h_max = img->comp[0].h_samp, v_max = img->comp[0].v_samp;
    /* chroma must be 1x1, decoder.rs:634-643; a 1-component image reaches the path only as 1x1
     * (mcu.rs:171-196 resets it) */
    if (img->n_comp == 1 && (h_max != 1 || v_max != 1)) return ZJO_ERR_UNSUPPORTED;
    for (uint32_t z = 1; z < img->n_comp; z++)
        if (img->comp[z].h_samp != 1 || img->comp[z].v_samp != 1) return ZJO_ERR_UNSUPPORTED;
    if (!((h_max == 1 || h_max == 2) && (v_max == 1 || v_max == 2))) return ZJO_ERR_UNSUPPORTED; /* decoder.rs:512-519 */
    const size_t w = img->width, h = img->height;
    const int interleaved = (h_max != 1 || v_max != 1);           /* headers.rs:320-325 */
    const size_t mcu_x = (w + 8 * h_max - 1) / (8 * h_max);       /* headers.rs:316 */
    const size_t mcu_y = (h + 8 * v_max - 1) / (8 * v_max);       /* headers.rs:318 */
    for (uint32_t z = 0; z < img->n_comp; z++)                    /* headers.rs:338, decoder.rs:622-632 */
        if (img->comp[z].width_stride != img->comp[z].h_samp * mcu_x * 8) return ZJO_ERR_ARG;

    size_t mcu_width, n_strips, bias = 1;
    if (h_max == 2 && v_max == 1) { mcu_width = mcu_x * 2; n_strips = mcu_y / 2; }             /* mcu.rs:147-154, mcu_prog.rs:138-141 */
    else if (h_max == 2 && v_max == 2) { mcu_width = mcu_x; n_strips = mcu_y / 2; bias = 2; }  /* mcu.rs:155-159, mcu_prog.rs:142-144 */
    else if (interleaved) { mcu_width = mcu_x; n_strips = mcu_y; }                             /* mcu.rs:160-163 */
    else { mcu_width = (w + 7) / 8; n_strips = (h + 7) / 8; }                                  /* mcu.rs:164-169 */

    g->nc = nc;
    g->y_chunk = mcu_width * 64 * v_max * h_max * bias;   /* mcu.rs:246, mcu_prog.rs:191-192 */
    g->c_chunk = mcu_width * 64 * bias;                   /* mcu_prog.rs:199-200 */
    g->out_chunk = w * nc * 8 * h_max * v_max;            /* mcu.rs:226, mcu_prog.rs:188 */
    g->out_size = w * h * nc;                             /* mcu.rs:375-379 */
    /* over-allocated output, mcu.rs:198,207,222 / mcu_prog.rs:173-176; `width + 8` is u16 arithmetic */
    size_t capacity = (size_t)(uint16_t)(w + 8) * (size_t)(uint16_t)(h + 8);
    size_t extra = (interleaved ? 1u : 0u) * 128 * h * nc;
    size_t alloc = capacity * nc + extra;
    size_t avail = alloc / g->out_chunk;
    if (n_strips > avail) {
        if (img->flags & ZJ_FLAG_PROGRESSIVE) n_strips = avail; /* zip stops, mcu_prog.rs:206-209 */
        else return ZJO_ERR_PANIC;                              /* chunks.next().unwrap(), mcu.rs:354 */
    }
    g->n_strips = n_strips;
    return ZJO_OK;
}

size_t zjo_output_size(const zj_image *img)
{
    geometry g;
    if (compute_geometry(img, &g) != ZJO_OK) return 0;
    return g.out_size;
}

/* Strip-parallel driver (the role of scoped_threadpool in mcu.rs:135,356-368 / mcu_prog.rs:196-233): a PERSISTENT pool --
 * workers are created once and woken per image, strips are handed out by an atomic counter -- so that the CPU baseline
 * bench.py times does not pay thread creation per image or a lock per strip. */
typedef struct {
    const zj_image *img;
    const geometry *g;
    uint8_t *out;
    atomic_size_t next;  /* next strip to take */
    atomic_int rc;
} strip_job;

static int run_strip(const zj_image *img, const geometry *g, uint8_t *out, size_t s)
{
    const uint32_t in_n = img->n_comp, out_n = (uint32_t)cs_components(img->out_cs);
    const uint32_t x = in_n < out_n ? in_n : out_n;
    const int16_t *coeff[3] = {NULL, NULL, NULL};
    size_t len[3] = {0, 0, 0};
    for (uint32_t z = 0; z < x; z++) {
        size_t chunk = z == 0 ? g->y_chunk : g->c_chunk;
        coeff[z] = img->comp[z].coeff + s * chunk;
        len[z] = chunk;
    }
    size_t off = s * g->out_chunk;
    if (off + g->out_chunk <= g->out_size) {
        return zjo_post_process(coeff, len, img, out + off, g->out_chunk);
    }
    /* strip straddles the truncation point (mcu.rs:375): run it in a scratch chunk, keep the prefix */
    uint8_t *tmp = (uint8_t *)calloc(g->out_chunk, 1);
    if (!tmp) return ZJO_ERR_ARG;
    int rc = zjo_post_process(coeff, len, img, tmp, g->out_chunk);
    if (rc == ZJO_OK && off < g->out_size) memcpy(out + off, tmp, g->out_size - off);
    free(tmp);
    return rc;
}

static void run_strips(strip_job *job)
{
    for (;;) {
        if (atomic_load_explicit(&job->rc, memory_order_relaxed) != ZJO_OK) break;
        size_t s = atomic_fetch_add_explicit(&job->next, 1, memory_order_relaxed);
        if (s >= job->g->n_strips) break;
        int rc = run_strip(job->img, job->g, job->out, s);
        if (rc != ZJO_OK) { int ok = ZJO_OK; atomic_compare_exchange_strong(&job->rc, &ok, rc); }
    }
}

#define ZJO_MAX_THREADS 256
static struct {
    pthread_mutex_t call;              /* one image at a time uses the pool */
    pthread_mutex_t mu;
    pthread_cond_t work, done;
    int n_workers;                     /* threads created so far */
    int wanted;                        /* workers that should take part in the current job */
    unsigned generation;
    int running;                       /* workers still inside the current job */
    strip_job *job;
    pthread_t th[ZJO_MAX_THREADS];
} pool = {PTHREAD_MUTEX_INITIALIZER, PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, PTHREAD_COND_INITIALIZER, 0, 0, 0, 0, NULL, {0}};

static void *pool_worker(void *p)
{
    const int id = (int)(intptr_t)p;
    unsigned seen = 0;
    pthread_mutex_lock(&pool.mu);
    for (;;) {
        while (pool.generation == seen) pthread_cond_wait(&pool.work, &pool.mu);
        seen = pool.generation;
        if (id >= pool.wanted) continue;
        strip_job *job = pool.job;
        pthread_mutex_unlock(&pool.mu);
        run_strips(job);
        pthread_mutex_lock(&pool.mu);
        if (--pool.running == 0) pthread_cond_signal(&pool.done);
    }
    return NULL;
}

int zjo_reconstruct_image(const zj_image *img, uint8_t *out, size_t out_len, int threads)
{
    geometry g;
    int rc = compute_geometry(img, &g);
    if (rc != ZJO_OK) return rc;
    if (!out || out_len < g.out_size) return ZJO_ERR_SHORT_OUTPUT;
    const uint32_t in_n = img->n_comp, out_n = (uint32_t)cs_components(img->out_cs);
    const uint32_t x = in_n < out_n ? in_n : out_n;
    for (uint32_t z = 0; z < x; z++) {
        size_t chunk = z == 0 ? g.y_chunk : g.c_chunk;
        if (!img->comp[z].coeff || img->comp[z].n_i16 < g.n_strips * chunk) return ZJO_ERR_SHORT_PLANE;
    }
    memset(out, 0, g.out_size); /* vec![0; ...], mcu.rs:222 */

    if (threads <= 1) {
        for (size_t s = 0; s < g.n_strips; s++) {
            rc = run_strip(img, &g, out, s);
            if (rc != ZJO_OK) return rc;
        }
        return ZJO_OK;
    }
    strip_job job;
    job.img = img; job.g = &g; job.out = out;
    atomic_init(&job.next, 0);
    atomic_init(&job.rc, ZJO_OK);
    if (threads > ZJO_MAX_THREADS) threads = ZJO_MAX_THREADS;
    pthread_mutex_lock(&pool.call);
    pthread_mutex_lock(&pool.mu);
    while (pool.n_workers < threads - 1) {   /* the caller is the last worker */
        if (pthread_create(&pool.th[pool.n_workers], NULL, pool_worker, (void *)(intptr_t)pool.n_workers) != 0) break;
        pthread_detach(pool.th[pool.n_workers]);
        pool.n_workers++;
    }
    pool.wanted = threads - 1 < pool.n_workers ? threads - 1 : pool.n_workers;
    pool.running = pool.wanted;
    pool.job = &job;
    pool.generation++;
    pthread_cond_broadcast(&pool.work);
    pthread_mutex_unlock(&pool.mu);
    run_strips(&job);
    pthread_mutex_lock(&pool.mu);
    while (pool.running > 0) pthread_cond_wait(&pool.done, &pool.mu);
    pthread_mutex_unlock(&pool.mu);
    pthread_mutex_unlock(&pool.call);
    return atomic_load(&job.rc);
}
