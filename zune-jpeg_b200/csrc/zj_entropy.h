// zj_entropy.h -- descriptors of the GPU form of the baseline entropy stage (zj_entropy.cu), filled by the host
// front-end (zj_host_decoder.cpp).
//
// Scope: baseline scans with restart markers (DRI).  Every restart interval starts from a known state (fresh bit
// reader right behind its RSTn, predictors 0, countdown = DRI), so one GPU thread per interval runs the reference's MCU
// loop (src/mcu.rs:253-351, handle_rst :386-418) with the reference's bit reader (src/bitstream.rs:159-402) and writes
// coefficients straight into the device planes the reconstruction kernels read.  An interval that does not end exactly
// the way the sequential loop would end it is reported, and the image goes through the host stage instead.
#pragma once
#include <stdint.h>

namespace zj {

struct EntTable {          // HuffmanTable, src/huffman.rs:14-41 (the members the decode loop reads)
    int32_t lookup[512];
    int16_t ac_lookup[512];
    int32_t maxcode[18];
    int32_t offset[18];
    uint8_t values[256];
};

struct EntImage {
    const uint8_t *data;          // the whole JPEG file (device)
    uint32_t len;                 // its length (the reader's bounds, bitstream.rs:696-703)
    const uint32_t *seg_start;    // n_seg + 1 entries: reader position at the start of interval k; entry n_seg is unused
    uint8_t *status;              // n_seg entries, written by the kernel: 0 = ended like the sequential loop ends it
    const EntTable *tables;       // 2 * ncomp entries: [dc, ac] of component 0, 1, 2
    int16_t *plane[3];            // whole-image planes (zeroed before the launch); nullptr = component not output (mcu.rs:244)
    uint32_t n_seg, per_seg, total_mcus, restart_interval;
    uint32_t mcu_w, bias, ncomp, is_hv;
    uint32_t width_stride, hv_width_stride;
    uint32_t strip_len[3];
    uint32_t h_samp[3], v_samp[3], is_y[3];
};

// One thread per restart interval, LANES of them per warp (the other lanes leave at once).  The lanes of a warp sit in
// different branches of the symbol loop and a warp step costs what all the branches it holds cost together, so fewer intervals
// per warp means shorter steps and more warps to hide the dependent table / byte loads behind -- as long as the GPU has warp
// slots to spare.  Measured (JPEG bytes -> pixels in HBM, 4K 4:2:0, 135 intervals per image): 32 images 29.0 / 13.5 / 8.9 / 12.1 ms
// with 32 / 16 / 8 / 4 lanes, 256 images 27.2 / 30.1 / 30.7 / 33.4 ms: launch_entropy picks 8, 16 or 32 by the number of
// intervals of the launch.  grid = (ceil(max n_seg / intervals per CTA), images)
constexpr int ENT_THREADS = 256;                         // threads per CTA
int launch_entropy(const EntImage *d_images, uint32_t n_images, uint32_t max_seg, void *stream);

}  // namespace zj
