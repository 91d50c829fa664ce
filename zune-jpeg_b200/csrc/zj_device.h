// zj_device.h -- descriptors shared by the host-side planner (zj_capi.cu) and the kernels (zj_kernels.cu).
//
// Vocabulary follows the reference: an image is cut into MCU-row *strips* (the unit of
// worker::post_process, reference src/worker.rs:32; geometry src/mcu.rs:139-226, src/mcu_prog.rs:132-203);
// a strip is cut into *tiles* of TM MCU columns, one CTA per tile.
#pragma once
#include <stdint.h>

namespace zj {

enum Mode : int { MODE_NONE = 0, MODE_H = 1, MODE_V = 2, MODE_HV = 3 };  // SubSampRatios, components.rs:130
enum OutKind : int {
    OUT_RGB = 0,   // YCbCr -> RGB | "RGBA" | "RGBX": the 3-byte conv16 + row-tail rule (worker.rs:143-251)
    OUT_YCC = 1,   // YCbCr -> YCbCr interleave (color_convert/scalar.rs:119-169)
    OUT_GRAY = 2,  // (YCbCr | GRAYSCALE) -> GRAYSCALE (color_convert/scalar.rs:91-114)
    OUT_ZERO = 3   // every other pair: nothing is written (worker.rs:131-132)
};

// MCU columns per tile, chosen so that one CTA of ZJ_THREADS threads has about one 8x8 block per thread
// (including the chroma halo blocks):                 blocks / MCU column        + halo
//   (per 128 threads) NONE: 3 blocks -> TM 40 -> 120  H: 8 -> TM 15 -> 120 + 8
//   V:    4 blocks  -> TM 32 -> 128                   HV: 12 -> TM 10 -> 120 + 8 (+4 in tile 0)
#ifndef ZJ_CFG_THREADS
#define ZJ_CFG_THREADS 256
#endif
#ifndef ZJ_CFG_MINBLOCKS
#define ZJ_CFG_MINBLOCKS 3
#endif
constexpr int ZJ_THREADS = ZJ_CFG_THREADS;
constexpr int ZJ_MINBLOCKS = ZJ_CFG_MINBLOCKS;  // CTAs per SM the register allocation is capped for
constexpr int ZJ_SLOW_CAP = 192;  // edge units per tile queued for the generic path
// tile widths scale with the CTA size (128 threads: 40 / 15 / 32 / 10)
constexpr int TM_NONE = 40 * ZJ_THREADS / 128, TM_H = 15 * ZJ_THREADS / 128, TM_V = 32 * ZJ_THREADS / 128, TM_HV = 10 * ZJ_THREADS / 128, TM_GRAY = 128;

// The fast kernel (X86 variant): 256 threads = 128 producers (IDCT, two 8x8 blocks per thread and strip) + 128
// consumers (up-sampling / colour / stores, two 16-sample units per thread and strip).
#ifndef ZF_CFG_MINBLOCKS
#define ZF_CFG_MINBLOCKS 3
#endif
#ifndef ZF_CFG_SPC
#define ZF_CFG_SPC 40
#endif
constexpr int ZF_THREADS = 256, ZF_PRODUCERS = 128, ZF_CONSUMERS = 128;
constexpr int ZF_MINBLOCKS = ZF_CFG_MINBLOCKS;
constexpr int ZF_DEFAULT_SPC = ZF_CFG_SPC;      // target strips per CTA (see strips_per_cta)
// unit columns (16 luma samples) per tile: 2 * ZF_CONSUMERS / row groups per strip
constexpr int ZF_XU_GRAY = 32;                  // luma-only fast kernel: 512-sample tiles
// (4:2:2: 15 instead of 16 -- the strip-tile's block list, 2 x 30 luma + 2 x 2 x 15 chroma + 8 halo blocks, is then exactly
// 128 = one IDCT pass of the four producer warps; with 16 it is 136 and the producers' critical path a second pass)
#ifndef ZF_CFG_XU_H
#define ZF_CFG_XU_H 15
#endif
constexpr int ZF_XU_NONE = 2 * ZF_CONSUMERS / 8, ZF_XU_H = ZF_CFG_XU_H, ZF_XU_V = 2 * ZF_CONSUMERS / 8, ZF_XU_HV = 2 * ZF_CONSUMERS / 16;

struct DevImage {
    const int16_t *coeff[3];  // device pointers, whole-image planes
    uint8_t *out;             // device pointer, width*height*nc bytes
    uint32_t qtw[3][32];      // quantisation tables packed for dp2a: word k = q[2k] | q[2k+1] << 24
    uint32_t width, height;
    uint32_t nc;              // output bytes per pixel
    uint32_t out_kind;        // OutKind
    uint32_t mcu_x;           // MCU columns (headers.rs:316)
    uint32_t n_strips;        // strips the reference processes (Q1: may not cover the image)
    uint32_t n_tiles;         // tiles per strip
    uint32_t Wp;              // luma plane row width = comp[0].width_stride
    uint32_t W;               // chroma plane row width = comp[1].width_stride (0 if no chroma)
    uint32_t stride;          // output row stride in bytes = width*nc
    // colour writer constants (worker.rs:171,221-223)
    uint32_t n_norm;          // samples [0, n_norm) of a row are written at byte 3*s ("normal" chunks)
    uint32_t P;               // 3*n_norm for RGB; bytes >= P not covered by the tail stay zero
    uint32_t T;               // tail chunk (samples Wp-16..Wp-1) lands at bytes [T, T+48); 0xffffffff = none
    uint32_t small_width;     // width < 16: temp-buffer path (worker.rs:158-163,176-198)
    uint32_t hv_avx;          // HV + X86 + chroma strip >= 500 samples -> AVX2 closed form
    uint32_t gray_rows_ok;    // Q7 resolved: 1 = plain row copy is what the reference does
    uint32_t gray_brows;      // luma-only output: block rows of the Y plane the reference processes
    uint32_t tile_q, tile_r;  // tile t covers MCU columns (fast kernel: 16-sample unit columns) [t*q + min(t,r), ...): the first r tiles are one wider
    uint64_t magic_w;         // ceil(2^40 / W): idx / W == (idx * magic_w) >> 40 for idx < 2^20
};

// Host-side launch plan entry: one kernel launch per (mode, variant, out-kind class) group.
struct LaunchGroup {
    int mode, variant, gray;  // gray = luma-only kernel
    int fast;                 // reconstruct_fast_kernel (X86 variant, aligned output)
    uint32_t first, count;    // images [first, first+count) of the sorted device array
    uint32_t max_tiles, max_strips;
};

// One image of a consumer launch (zj_consumer.cu): interleaved u8 pixels in, the layout / type / scale of zj_output_desc out.
struct ConvImage {
    const uint8_t *src;           // u8, interleaved, width * height * nc
    void *dst;
    uint32_t width, height, nc;   // source geometry
    uint32_t ow, oh, oc;          // output geometry (oc = channels written)
};

}  // namespace zj
