/*
 * zj_oracle.h -- CPU oracle for the pixel-reconstruction path of etemesi254/zune-jpeg 0.2.0.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library, and only as the checker / reported CPU baseline.  Nothing
 * under zune-jpeg_b200/ links, imports or executes it.
 *
 * It is a line-by-line C restatement of /root/reference/src/worker.rs, src/idct/, src/upsampler/ and
 * src/color_convert/ plus the strip geometry of mcu.rs / mcu_prog.rs; every function cites the lines it
 * follows.  The reference is Rust and there is no rustc/cargo in the build environment, so it cannot be
 * compiled into oracle/_ref; see oracle/README.md.
 *
 * PARITY PINNING: the reference's own tests pin only (a) the three IDCT known-answer vectors
 * src/idct.rs:66-127 and (b) the SSE==scalar ramp equalities src/upsampler.rs:126-150.  This oracle
 * reproduces all of them (tests/test_oracle_kat.py).  Everything else on the path -- colour conversion,
 * V/HV upsampling, row-tail handling, strip rules -- has no vector in the reference and the reference cannot
 * be executed here, so for those stages parity is UNPINNED: the oracle is authoritative only by literal
 * fidelity to the cited lines (cross-checked by an independent closed-form model in tests/ref_model.py and
 * by building the SIMD parts both with real intrinsics and with an emulation, see simd_compat.h).
 */
#ifndef ZJ_ORACLE_H
#define ZJ_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#include "../include/zune_jpeg_b200.h" /* zj_image / zj_component descriptors only (types, no code) */

#ifdef __cplusplus
extern "C" {
#endif

#define ZJO_OK 0
#define ZJO_ERR_ARG (-1)
#define ZJO_ERR_UNSUPPORTED (-2)
#define ZJO_ERR_SHORT_PLANE (-3)
#define ZJO_ERR_SHORT_OUTPUT (-4)
#define ZJO_ERR_PANIC (-5) /* the reference would panic (assert!/unwrap/slice index) at this point */

/* 1 when built with real SSE4.1/AVX2 intrinsics, 0 when built with the portable emulation */
int zjo_uses_real_simd(void);

/* ---- IDCT: IDCTPtr (decoder.rs:56).  out has `len` i16 and must be zeroed by the caller (vec![0; len]) */
int zjo_idct_scalar(const int16_t *vector, size_t len, const int32_t *qt, size_t stride,
                    size_t samp_factors, size_t v_samp, int16_t *out); /* idct/scalar.rs:19-282 */
int zjo_idct_avx2(const int16_t *coeff, size_t len, const int32_t *qt, size_t stride,
                  size_t samp_factors, size_t v_samp, int16_t *out); /* idct/avx2.rs:64-398 */

/* ---- up-samplers: UpSampler (components.rs:14).  out has out_len i16; the functions zero it first */
int zjo_upsample_horizontal_scalar(const int16_t *in, size_t n, int16_t *out, size_t out_len); /* upsampler/scalar.rs:5-60 */
int zjo_upsample_horizontal_sse(const int16_t *in, size_t n, int16_t *out, size_t out_len);    /* upsampler/sse.rs:24-134 */
int zjo_upsample_vertical(const int16_t *in, size_t n, int16_t *out, size_t out_len);          /* upsampler/scalar.rs:64-147 */
int zjo_upsample_hv_scalar(const int16_t *in, size_t n, int16_t *out, size_t out_len);         /* upsampler/scalar.rs:148-166 */
int zjo_upsample_hv_simd(const int16_t *in, size_t n, int16_t *out, size_t out_len);           /* upsampler/avx2.rs:14-342 */

/* ---- colour: ColorConvert16Ptr (decoder.rs:47) and the two slice writers */
int zjo_ycbcr_to_rgb_16(const int16_t *y, const int16_t *cb, const int16_t *cr, uint8_t *out,
                        size_t out_len, size_t *pos, int variant); /* color_convert/avx.rs:67-192, scalar.rs:52-89 */
int zjo_ycbcr_to_grayscale(const int16_t *y, size_t len, size_t width, uint8_t *out, size_t out_len); /* scalar.rs:91-114 */
int zjo_ycbcr_to_ycbcr(const int16_t *const ch[3], size_t len0, size_t width, size_t h_samp,
                       size_t v_samp, uint8_t *out, size_t out_len); /* scalar.rs:119-169 */

/* ---- worker::post_process for ONE strip (worker.rs:32-251).  coeff[z]/len[z] are the strip slices. */
int zjo_post_process(const int16_t *const coeff[3], const size_t len[3], const zj_image *img,
                     uint8_t *output, size_t output_len);

/* ---- whole image: the strip loop of mcu.rs:198-226,354-379 / mcu_prog.rs:173-241.
 * out must hold zj-output-size bytes = width*height*out_components; it is fully written (zeros where the
 * reference leaves its zero-initialised Vec untouched).  threads<=1: strips in order on the caller's
 * thread; threads>1: strips handed to a pool of that many threads, as scoped_threadpool does
 * (mcu.rs:135,356-368). */
int zjo_reconstruct_image(const zj_image *img, uint8_t *out, size_t out_len, int threads);
size_t zjo_output_size(const zj_image *img);

#ifdef __cplusplus
}
#endif
#endif
