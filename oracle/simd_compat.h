/*
 * simd_compat.h -- the handful of x86 SIMD operations the reference's X86
 * variant uses, under neutral names (zv_*), in two interchangeable builds:
 *
 *   -DZJO_REAL_SIMD  : each zv_* is the real SSE4.1/AVX2 intrinsic (immintrin.h)
 *   (default)        : each zv_* is a portable byte-level emulation written from
 *                      the Intel Intrinsics Guide pseudo-code
 *
 * TEST INFRASTRUCTURE ONLY (part of oracle/).  The oracle is compiled both ways
 * (oracle/Makefile) and tests/test_oracle_simd.py checks the two builds agree
 * bit-for-bit on random data, which pins the emulation (and therefore my reading
 * of every intrinsic the reference calls) against the real hardware.
 */
#ifndef ZJO_SIMD_COMPAT_H
#define ZJO_SIMD_COMPAT_H

#include <stdint.h>
#include <string.h>

#ifdef ZJO_REAL_SIMD
/* ------------------------------------------------------------------ real */
#include <immintrin.h>

typedef __m128i zv128;
typedef __m256i zv256;

#define zv128_loadu(p) _mm_loadu_si128((const __m128i *)(p))
#define zv128_loadl64(p) _mm_loadl_epi64((const __m128i *)(p))
#define zv128_storeu(p, v) _mm_storeu_si128((__m128i *)(p), (v))
#define zv128_or(a, b) _mm_or_si128((a), (b))
#define zv128_test_all_zeros(a, b) _mm_test_all_zeros((a), (b))
#define zv128_set1_epi16(x) _mm_set1_epi16((short)(x))
#define zv128_unpacklo_epi16(a, b) _mm_unpacklo_epi16((a), (b))
#define zv128_blend_epi16(a, b, imm) _mm_blend_epi16((a), (b), (imm))
#define zv128_add_epi16(a, b) _mm_add_epi16((a), (b))
#define zv128_slli_epi16(a, imm) _mm_slli_epi16((a), (imm))
#define zv128_srai_epi16(a, imm) _mm_srai_epi16((a), (imm))

#define zv256_loadu(p) _mm256_loadu_si256((const __m256i *)(p))
#define zv256_storeu(p, v) _mm256_storeu_si256((__m256i *)(p), (v))
#define zv256_set1_epi16(x) _mm256_set1_epi16((short)(x))
#define zv256_set1_epi32(x) _mm256_set1_epi32((int)(x))
#define zv256_cvtepi16_epi32(a) _mm256_cvtepi16_epi32((a))
#define zv256_mullo_epi32(a, b) _mm256_mullo_epi32((a), (b))
#define zv256_mullo_epi16(a, b) _mm256_mullo_epi16((a), (b))
#define zv256_add_epi32(a, b) _mm256_add_epi32((a), (b))
#define zv256_sub_epi32(a, b) _mm256_sub_epi32((a), (b))
#define zv256_add_epi16(a, b) _mm256_add_epi16((a), (b))
#define zv256_sub_epi16(a, b) _mm256_sub_epi16((a), (b))
#define zv256_slli_epi32(a, imm) _mm256_slli_epi32((a), (imm))
#define zv256_srai_epi32(a, imm) _mm256_srai_epi32((a), (imm))
#define zv256_srai_epi16(a, imm) _mm256_srai_epi16((a), (imm))
#define zv256_packs_epi32(a, b) _mm256_packs_epi32((a), (b))
#define zv256_max_epi16(a, b) _mm256_max_epi16((a), (b))
#define zv256_min_epi16(a, b) _mm256_min_epi16((a), (b))
#define zv256_permute4x64_epi64(a, imm) _mm256_permute4x64_epi64((a), (imm))
#define zv256_permute2x128(a, b, imm) _mm256_permute2x128_si256((a), (b), (imm))
#define zv256_extract128(a, imm) _mm256_extractf128_si256((a), (imm))
#define zv256_unpacklo_epi32(a, b) _mm256_unpacklo_epi32((a), (b))
#define zv256_unpackhi_epi32(a, b) _mm256_unpackhi_epi32((a), (b))
#define zv256_unpacklo_epi64(a, b) _mm256_unpacklo_epi64((a), (b))
#define zv256_unpackhi_epi64(a, b) _mm256_unpackhi_epi64((a), (b))
#define zv256_unpacklo_epi16(a, b) _mm256_unpacklo_epi16((a), (b))
#define zv256_unpackhi_epi16(a, b) _mm256_unpackhi_epi16((a), (b))
#define zv256_alignr_epi8(a, b, imm) _mm256_alignr_epi8((a), (b), (imm))
#define zv256_blend_epi16(a, b, imm) _mm256_blend_epi16((a), (b), (imm))
#define zv256_insert_epi16(a, x, idx) _mm256_insert_epi16((a), (short)(x), (idx))

#else
/* -------------------------------------------------------------- emulated */

typedef struct { uint8_t b[16]; } zv128;
typedef struct { uint8_t b[32]; } zv256;

#define ZV_I16(v, i) (((int16_t *)(void *)(v).b)[i])
#define ZV_I32(v, i) (((int32_t *)(void *)(v).b)[i])
#define ZV_I64(v, i) (((int64_t *)(void *)(v).b)[i])

static inline int16_t zv_ld16(const uint8_t *b, int i) { int16_t x; memcpy(&x, b + 2 * i, 2); return x; }
static inline void zv_st16(uint8_t *b, int i, int16_t x) { memcpy(b + 2 * i, &x, 2); }
static inline int32_t zv_ld32(const uint8_t *b, int i) { int32_t x; memcpy(&x, b + 4 * i, 4); return x; }
static inline void zv_st32(uint8_t *b, int i, int32_t x) { memcpy(b + 4 * i, &x, 4); }
static inline uint64_t zv_ld64(const uint8_t *b, int i) { uint64_t x; memcpy(&x, b + 8 * i, 8); return x; }
static inline void zv_st64(uint8_t *b, int i, uint64_t x) { memcpy(b + 8 * i, &x, 8); }

/* ---- 128-bit ---- */
static inline zv128 zv128_loadu(const void *p) { zv128 r; memcpy(r.b, p, 16); return r; }
/* MOVQ: low 64 bits loaded, high 64 zeroed */
static inline zv128 zv128_loadl64(const void *p) { zv128 r; memset(r.b, 0, 16); memcpy(r.b, p, 8); return r; }
static inline void zv128_storeu(void *p, zv128 v) { memcpy(p, v.b, 16); }
static inline zv128 zv128_or(zv128 a, zv128 b) { zv128 r; for (int i = 0; i < 16; i++) r.b[i] = a.b[i] | b.b[i]; return r; }
/* PTEST: returns ZF = ((a AND b) == 0) */
static inline int zv128_test_all_zeros(zv128 a, zv128 b) { for (int i = 0; i < 16; i++) if (a.b[i] & b.b[i]) return 0; return 1; }
static inline zv128 zv128_set1_epi16(int x) { zv128 r; for (int i = 0; i < 8; i++) zv_st16(r.b, i, (int16_t)x); return r; }
/* PUNPCKLWD: interleave low four words */
static inline zv128 zv128_unpacklo_epi16(zv128 a, zv128 b) {
    zv128 r; for (int i = 0; i < 4; i++) { zv_st16(r.b, 2 * i, zv_ld16(a.b, i)); zv_st16(r.b, 2 * i + 1, zv_ld16(b.b, i)); } return r; }
/* PBLENDW: bit j of imm selects b for word j */
static inline zv128 zv128_blend_epi16(zv128 a, zv128 b, int imm) {
    zv128 r; for (int i = 0; i < 8; i++) zv_st16(r.b, i, ((imm >> i) & 1) ? zv_ld16(b.b, i) : zv_ld16(a.b, i)); return r; }
static inline zv128 zv128_add_epi16(zv128 a, zv128 b) {
    zv128 r; for (int i = 0; i < 8; i++) zv_st16(r.b, i, (int16_t)(uint16_t)((uint16_t)zv_ld16(a.b, i) + (uint16_t)zv_ld16(b.b, i))); return r; }
static inline zv128 zv128_slli_epi16(zv128 a, int imm) {
    zv128 r; for (int i = 0; i < 8; i++) zv_st16(r.b, i, imm > 15 ? 0 : (int16_t)(uint16_t)((uint16_t)zv_ld16(a.b, i) << imm)); return r; }
static inline zv128 zv128_srai_epi16(zv128 a, int imm) {
    zv128 r; if (imm > 15) imm = 15; for (int i = 0; i < 8; i++) zv_st16(r.b, i, (int16_t)(zv_ld16(a.b, i) >> imm)); return r; }

/* ---- 256-bit ---- */
static inline zv256 zv256_loadu(const void *p) { zv256 r; memcpy(r.b, p, 32); return r; }
static inline void zv256_storeu(void *p, zv256 v) { memcpy(p, v.b, 32); }
static inline zv256 zv256_set1_epi16(int x) { zv256 r; for (int i = 0; i < 16; i++) zv_st16(r.b, i, (int16_t)x); return r; }
static inline zv256 zv256_set1_epi32(int32_t x) { zv256 r; for (int i = 0; i < 8; i++) zv_st32(r.b, i, x); return r; }
/* VPMOVSXWD */
static inline zv256 zv256_cvtepi16_epi32(zv128 a) { zv256 r; for (int i = 0; i < 8; i++) zv_st32(r.b, i, (int32_t)zv_ld16(a.b, i)); return r; }
static inline zv256 zv256_mullo_epi32(zv256 a, zv256 b) {
    zv256 r; for (int i = 0; i < 8; i++) zv_st32(r.b, i, (int32_t)((uint32_t)zv_ld32(a.b, i) * (uint32_t)zv_ld32(b.b, i))); return r; }
static inline zv256 zv256_mullo_epi16(zv256 a, zv256 b) {
    zv256 r; for (int i = 0; i < 16; i++) zv_st16(r.b, i, (int16_t)(uint16_t)((uint32_t)(uint16_t)zv_ld16(a.b, i) * (uint32_t)(uint16_t)zv_ld16(b.b, i))); return r; }
static inline zv256 zv256_add_epi32(zv256 a, zv256 b) {
    zv256 r; for (int i = 0; i < 8; i++) zv_st32(r.b, i, (int32_t)((uint32_t)zv_ld32(a.b, i) + (uint32_t)zv_ld32(b.b, i))); return r; }
static inline zv256 zv256_sub_epi32(zv256 a, zv256 b) {
    zv256 r; for (int i = 0; i < 8; i++) zv_st32(r.b, i, (int32_t)((uint32_t)zv_ld32(a.b, i) - (uint32_t)zv_ld32(b.b, i))); return r; }
static inline zv256 zv256_add_epi16(zv256 a, zv256 b) {
    zv256 r; for (int i = 0; i < 16; i++) zv_st16(r.b, i, (int16_t)(uint16_t)((uint16_t)zv_ld16(a.b, i) + (uint16_t)zv_ld16(b.b, i))); return r; }
static inline zv256 zv256_sub_epi16(zv256 a, zv256 b) {
    zv256 r; for (int i = 0; i < 16; i++) zv_st16(r.b, i, (int16_t)(uint16_t)((uint16_t)zv_ld16(a.b, i) - (uint16_t)zv_ld16(b.b, i))); return r; }
static inline zv256 zv256_slli_epi32(zv256 a, int imm) {
    zv256 r; for (int i = 0; i < 8; i++) zv_st32(r.b, i, imm > 31 ? 0 : (int32_t)((uint32_t)zv_ld32(a.b, i) << imm)); return r; }
static inline zv256 zv256_srai_epi32(zv256 a, int imm) {
    zv256 r; if (imm > 31) imm = 31; for (int i = 0; i < 8; i++) zv_st32(r.b, i, zv_ld32(a.b, i) >> imm); return r; }
static inline zv256 zv256_srai_epi16(zv256 a, int imm) {
    zv256 r; if (imm > 15) imm = 15; for (int i = 0; i < 16; i++) zv_st16(r.b, i, (int16_t)(zv_ld16(a.b, i) >> imm)); return r; }
static inline int16_t zv_sat16(int32_t x) { return x > 32767 ? 32767 : (x < -32768 ? -32768 : (int16_t)x); }
/* VPACKSSDW: per 128-bit lane, a's four dwords then b's four dwords, signed saturation */
static inline zv256 zv256_packs_epi32(zv256 a, zv256 b) {
    zv256 r;
    for (int lane = 0; lane < 2; lane++)
        for (int i = 0; i < 4; i++) {
            zv_st16(r.b, lane * 8 + i, zv_sat16(zv_ld32(a.b, lane * 4 + i)));
            zv_st16(r.b, lane * 8 + 4 + i, zv_sat16(zv_ld32(b.b, lane * 4 + i)));
        }
    return r;
}
static inline zv256 zv256_max_epi16(zv256 a, zv256 b) {
    zv256 r; for (int i = 0; i < 16; i++) { int16_t x = zv_ld16(a.b, i), y = zv_ld16(b.b, i); zv_st16(r.b, i, x > y ? x : y); } return r; }
static inline zv256 zv256_min_epi16(zv256 a, zv256 b) {
    zv256 r; for (int i = 0; i < 16; i++) { int16_t x = zv_ld16(a.b, i), y = zv_ld16(b.b, i); zv_st16(r.b, i, x < y ? x : y); } return r; }
/* VPERMQ: qword j of the result = qword ((imm >> 2j) & 3) of a */
static inline zv256 zv256_permute4x64_epi64(zv256 a, int imm) {
    zv256 r; for (int j = 0; j < 4; j++) zv_st64(r.b, j, zv_ld64(a.b, (imm >> (2 * j)) & 3)); return r; }
/* VPERM2I128: each result half chosen by a 4-bit field: 0=a.lo 1=a.hi 2=b.lo 3=b.hi, bit 3 = zero */
static inline zv256 zv256_permute2x128(zv256 a, zv256 b, int imm) {
    zv256 r;
    for (int h = 0; h < 2; h++) {
        int ctl = (imm >> (4 * h)) & 0xF;
        const uint8_t *src = (ctl & 2) ? b.b : a.b;
        if (ctl & 8) memset(r.b + 16 * h, 0, 16);
        else memcpy(r.b + 16 * h, src + 16 * (ctl & 1), 16);
    }
    return r;
}
static inline zv128 zv256_extract128(zv256 a, int imm) { zv128 r; memcpy(r.b, a.b + 16 * (imm & 1), 16); return r; }
/* VPUNPCK{L,H}DQ / QDQ / WD: per 128-bit lane interleave of the low / high half */
static inline zv256 zv256_unpacklo_epi32(zv256 a, zv256 b) {
    zv256 r; for (int l = 0; l < 2; l++) for (int i = 0; i < 2; i++) {
        zv_st32(r.b, l * 4 + 2 * i, zv_ld32(a.b, l * 4 + i)); zv_st32(r.b, l * 4 + 2 * i + 1, zv_ld32(b.b, l * 4 + i)); } return r; }
static inline zv256 zv256_unpackhi_epi32(zv256 a, zv256 b) {
    zv256 r; for (int l = 0; l < 2; l++) for (int i = 0; i < 2; i++) {
        zv_st32(r.b, l * 4 + 2 * i, zv_ld32(a.b, l * 4 + 2 + i)); zv_st32(r.b, l * 4 + 2 * i + 1, zv_ld32(b.b, l * 4 + 2 + i)); } return r; }
static inline zv256 zv256_unpacklo_epi64(zv256 a, zv256 b) {
    zv256 r; for (int l = 0; l < 2; l++) { zv_st64(r.b, l * 2, zv_ld64(a.b, l * 2)); zv_st64(r.b, l * 2 + 1, zv_ld64(b.b, l * 2)); } return r; }
static inline zv256 zv256_unpackhi_epi64(zv256 a, zv256 b) {
    zv256 r; for (int l = 0; l < 2; l++) { zv_st64(r.b, l * 2, zv_ld64(a.b, l * 2 + 1)); zv_st64(r.b, l * 2 + 1, zv_ld64(b.b, l * 2 + 1)); } return r; }
static inline zv256 zv256_unpacklo_epi16(zv256 a, zv256 b) {
    zv256 r; for (int l = 0; l < 2; l++) for (int i = 0; i < 4; i++) {
        zv_st16(r.b, l * 8 + 2 * i, zv_ld16(a.b, l * 8 + i)); zv_st16(r.b, l * 8 + 2 * i + 1, zv_ld16(b.b, l * 8 + i)); } return r; }
static inline zv256 zv256_unpackhi_epi16(zv256 a, zv256 b) {
    zv256 r; for (int l = 0; l < 2; l++) for (int i = 0; i < 4; i++) {
        zv_st16(r.b, l * 8 + 2 * i, zv_ld16(a.b, l * 8 + 4 + i)); zv_st16(r.b, l * 8 + 2 * i + 1, zv_ld16(b.b, l * 8 + 4 + i)); } return r; }
/* VPALIGNR: per 128-bit lane, (a.lane : b.lane) >> imm bytes, low 16 bytes kept */
static inline zv256 zv256_alignr_epi8(zv256 a, zv256 b, int imm) {
    zv256 r;
    for (int l = 0; l < 2; l++) {
        uint8_t tmp[32];
        memcpy(tmp, b.b + 16 * l, 16);
        memcpy(tmp + 16, a.b + 16 * l, 16);
        for (int i = 0; i < 16; i++) r.b[16 * l + i] = (imm + i < 32) ? tmp[imm + i] : 0;
    }
    return r;
}
/* VPBLENDW: the 8-bit imm is applied to each 128-bit lane */
static inline zv256 zv256_blend_epi16(zv256 a, zv256 b, int imm) {
    zv256 r; for (int i = 0; i < 16; i++) zv_st16(r.b, i, ((imm >> (i & 7)) & 1) ? zv_ld16(b.b, i) : zv_ld16(a.b, i)); return r; }
static inline zv256 zv256_insert_epi16(zv256 a, int x, int idx) { zv_st16(a.b, idx & 15, (int16_t)x); return a; }

#endif /* ZJO_REAL_SIMD */
#endif /* ZJO_SIMD_COMPAT_H */
