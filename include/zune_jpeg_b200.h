/*
 * zune_jpeg_b200.h -- C ABI of the B200-native pixel-reconstruction path of
 * etemesi254/zune-jpeg (snapshot 0.2.0).
 *
 * The reference has no FFI; its boundary for this path is the crate-private
 *
 *   worker::post_process(coeff:&[&[i16];3], component_data:&[Components], idct_func, color_convert_16,
 *                        input_colorspace, output_colorspace, output:&mut [u8], width)   (src/worker.rs:32-41)
 *
 * called once per MCU-row strip from src/mcu.rs:356-368 (baseline) and src/mcu_prog.rs:205-233
 * (progressive).  A strip is far too small a GPU launch, so this ABI takes WHOLE-IMAGE coefficient planes
 * (the layout mcu_prog.rs:73-79 allocates, and that the concatenation of mcu.rs's per-strip buffers equals)
 * and performs every strip's post_process -- dequantise + IDCT + upsample + colour-convert + row de-padding --
 * for a batch of images in one call.  See INTEGRATION.md for the Rust `extern "C"` block that binds it.
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every function returns 0 (ZJ_OK) or a negative
 * zj_status and never unwinds; the caller owns every buffer; a call is thread-safe per (device, stream).
 * There is NO CPU fallback: without a CUDA device every compute entry point returns ZJ_ERR_NO_DEVICE.
 */
#ifndef ZUNE_JPEG_B200_H
#define ZUNE_JPEG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define ZJ_API __declspec(dllexport)
#else
#define ZJ_API __attribute__((visibility("default")))
#endif

/* ------------------------------------------------------------------ enums */

/* Ordinals of `ColorSpace` (src/misc.rs:88-106). */
typedef enum zj_colorspace {
    ZJ_CS_RGB = 0,
    ZJ_CS_GRAYSCALE = 1,
    ZJ_CS_YCBCR = 2,
    ZJ_CS_CMYK = 3,
    ZJ_CS_YCCK = 4,
    ZJ_CS_RGBA = 5,
    ZJ_CS_RGBX = 6
} zj_colorspace;

/* Which of the reference's two CPU code paths the output must be bit-identical to.
 * X86    = ZuneJpegOptions::use_unsafe == true on an AVX2+SSE4.1 host (the default, src/options.rs:31):
 *          AVX2 IDCT (src/idct/avx2.rs), SSE H upsampler (src/upsampler/sse.rs), scalar V, AVX2 HV
 *          (src/upsampler/avx2.rs, scalar below 500 samples), AVX2 RGB (src/color_convert/avx.rs).
 * SCALAR = use_unsafe == false or `--no-default-features`: src/idct/scalar.rs, src/upsampler/scalar.rs,
 *          src/color_convert/scalar.rs. */
typedef enum zj_variant { ZJ_VARIANT_X86 = 0, ZJ_VARIANT_SCALAR = 1 } zj_variant;

typedef enum zj_status {
    ZJ_OK = 0,
    ZJ_ERR_INVALID_ARG = -1,     /* null pointer, bad enum, inconsistent descriptor                        */
    ZJ_ERR_UNSUPPORTED = -2,     /* sampling the reference rejects (decoder.rs:512-519, :609-646)          */
    ZJ_ERR_SHORT_PLANE = -3,     /* a coefficient plane is smaller than the strips it must feed            */
    ZJ_ERR_SHORT_OUTPUT = -4,    /* out_len[i] < width*height*out_components                               */
    ZJ_ERR_REF_PANIC = -5,       /* the reference panics on this geometry (tiny widths)     */
    ZJ_ERR_NO_DEVICE = -6,       /* no CUDA device / bad ordinal -- there is no CPU fallback               */
    ZJ_ERR_CUDA = -7,            /* a CUDA runtime call failed; zj_gpu_last_cuda_error() has the text      */
    ZJ_ERR_OOM = -8,             /* device or pinned allocation failed                                     */
    ZJ_ERR_DECODE = -9           /* host front-end (headers/entropy) error; zj_decoder_error() has details */
} zj_status;

/* ------------------------------------------------------------ descriptors */

/* One image component as the path sees it: the fields of `Components` (src/components.rs:18-43) that
 * post_process reads, plus the coefficient plane. */
typedef struct zj_component {
    const int16_t *coeff;  /* whole-image plane: blocks in raster order, width_stride/8 blocks per block-row,
                              64 i16 per block in NATURAL (de-zigzagged, row-major) order -- what
                              bitstream.rs:343,359 / mcu_prog.rs:298,404-406 produce.  Host or device pointer
                              depending on the entry point.  May be NULL for components the output colourspace
                              does not need (worker.rs:59). */
    uint64_t n_i16;        /* number of i16 in the plane                                                    */
    int32_t qt[64];        /* quantisation table, natural order (headers.rs:533-543)                        */
    uint32_t h_samp;       /* Components::horizontal_sample                                                 */
    uint32_t v_samp;       /* Components::vertical_sample                                                   */
    uint32_t width_stride; /* Components::width_stride = h_samp * mcu_x * 8 (headers.rs:338)                */
    uint32_t reserved;
} zj_component;

#define ZJ_FLAG_PROGRESSIVE 1u /* strips come from mcu_prog.rs:188-233 (zip stops silently) instead of
                                  mcu.rs:225-369 (chunks.next().unwrap()) -- only matters when the
                                  over-allocated output runs out of strips, see DESIGN.md                   */

typedef struct zj_image {
    uint32_t width, height; /* ImageInfo::width/height (decoder.rs:652-668)                                 */
    uint32_t n_comp;        /* 1 (GRAYSCALE in) or 3 (YCbCr in) = input_colorspace.num_components()         */
    uint32_t out_cs;        /* zj_colorspace: ZuneJpegOptions::out_colorspace                               */
    uint32_t variant;       /* zj_variant                                                                   */
    uint32_t flags;         /* ZJ_FLAG_*                                                                    */
    zj_component comp[3];   /* comp[0] = Y and carries the maximum sampling factors (decoder.rs:609-646)    */
} zj_image;

/* ------------------------------------------------------------- GPU path  */

/* Number of CUDA devices visible (0 when there is none; never fails). */
ZJ_API int zj_gpu_device_count(void);

/* Bytes the reference's decode_buffer returns for this image: width*height*out_components
 * (mcu.rs:375-379).  Returns 0 on a bad descriptor. */
ZJ_API size_t zj_output_size(const zj_image *img);

/* Validate a descriptor exactly as the GPU entry points do, without touching a device. */
ZJ_API int zj_validate_image(const zj_image *img);

/* Replaces every worker::post_process call of `n` images (worker.rs:32-85).
 * HOST entry point: imgs[i].comp[*].coeff and out[i] are host pointers (pinned memory from
 * zj_gpu_pinned_alloc makes the copies asynchronous); the call stages H2D, launches, copies D2H and
 * returns after `stream` has drained.  out[i] receives exactly zj_output_size(&imgs[i]) bytes; bytes the
 * reference leaves untouched in its zero-initialised Vec (mcu.rs:222) are written as 0. */
ZJ_API int zj_gpu_reconstruct(int device, void *stream, const zj_image *imgs, size_t n,
                              uint8_t *const *out, const size_t *out_len);

/* The same call in two halves, for callers that have host work to do meanwhile (mcu.rs:230-369 entropy-decodes strip k+1
 * while its pool post-processes strip k): _submit queues the uploads, kernels and downloads and returns; the planes and the
 * outputs must stay untouched until _finish(pending) has returned, which waits (the thread sleeps, it does not spin),
 * releases `pending` and reports the status of the whole call.  A failed _submit leaves nothing to finish. */
typedef struct zj_pending zj_pending;
ZJ_API int zj_gpu_reconstruct_submit(int device, void *stream, const zj_image *imgs, size_t n,
                                     uint8_t *const *out, const size_t *out_len, zj_pending **pending);
ZJ_API int zj_gpu_reconstruct_finish(zj_pending *pending);

/* Several devices of one box (north_star: "batches of images (and, for single huge images, restart-interval strips) are
 * partitioned across the GPUs with independent streams, no NCCL: the work shards with no reduction").  n >= n_dev: contiguous
 * image ranges, one per device (zj_partition).  n < n_dev: every image is cut into contiguous STRIP ranges, one per device
 * (zj_image_strip_range): worker::post_process (src/worker.rs:32-85) is called per strip and every rule in it is local to
 * that strip, so a device needs only its strips' coefficient rows and writes only its rows.  One host thread per device
 * drives zj_gpu_reconstruct on that device's streams.  Host pointers, as zj_gpu_reconstruct. */
ZJ_API int zj_gpu_reconstruct_multi(const int *devices, size_t n_dev, const zj_image *imgs, size_t n,
                                    uint8_t *const *out, const size_t *out_len);
/* [begin, end) of part `part` when n_items are cut into n_parts contiguous, balanced ranges (sizes differ by at most one). */
ZJ_API void zj_partition(size_t n_items, size_t n_parts, size_t part, size_t *begin, size_t *end);
/* Strips [strip_begin, strip_end) of `img` as an image of their own: *sub = the descriptor (planes advanced to the first
 * strip, height = the range's rows; the range that ends at the last strip also owns the rows below it), *out_offset /
 * *out_bytes = where its pixels live inside the whole image's output.  *n_strips = strips of the whole image (pass sub,
 * out_offset, out_bytes = NULL to query only that).  Reconstructing every range of a partition gives exactly the bytes of
 * reconstructing the whole image.  ZJ_ERR_UNSUPPORTED: this image cannot be cut (its strip count was limited by the
 * reference's output-capacity rule, mcu_prog.rs:206-209). */
ZJ_API int zj_image_strip_range(const zj_image *img, uint32_t strip_begin, uint32_t strip_end, zj_image *sub,
                                size_t *out_offset, size_t *out_bytes, uint32_t *n_strips);

/* DEVICE entry point: coefficient planes and outputs already live in the memory of `device`.
 * Asynchronous on `stream` (a cudaStream_t, NULL = legacy default stream); no host<->device pixel traffic.
 * Device coefficient planes must start on a 16-byte boundary (ZJ_ERR_INVALID_ARG otherwise; cudaMalloc'ed memory
 * always does); outputs may have any alignment (4-byte aligned outputs take the fastest kernel). */
ZJ_API int zj_gpu_reconstruct_device(int device, void *stream, const zj_image *imgs, size_t n,
                                     uint8_t *const *out_dev, const size_t *out_len);

/* A reusable plan for a fixed batch of device-resident images: descriptors, strip/tile work lists and qt
 * tables are uploaded once; zj_batch_run only launches kernels (CUDA-graph friendly). */
typedef struct zj_batch zj_batch;
ZJ_API int zj_batch_create(int device, const zj_image *imgs, size_t n, uint8_t *const *out_dev,
                           const size_t *out_len, zj_batch **plan);
ZJ_API int zj_batch_run(zj_batch *plan, void *stream);
/* kernels launched by one zj_batch_run (for accounting) */
ZJ_API int zj_batch_launches(const zj_batch *plan);
/* algorithmic bytes one zj_batch_run moves: coefficient bytes read + output bytes written */
ZJ_API uint64_t zj_batch_algorithmic_bytes(const zj_batch *plan);
ZJ_API void zj_batch_destroy(zj_batch *plan);

/* ---------------------------------------------------- device-side consumers */
/* What a GPU consumer of the pixels reads, produced without leaving the device.  The reference's writers stop at
 * interleaved u8 (the per-colourspace dispatch of src/worker.rs:113-133, stores of src/color_convert/scalar.rs:52-169);
 * this descriptor extends them.  Every value is a function of the EXACT u8 the reference writes:
 *   u8:            the byte itself; at half size (a + b + c + d + 2) >> 2 over the 2x2 box (odd last row / column dropped)
 *   f32:           (float(u8) - mean[c]) * inv_std[c]    -- two IEEE fp32 operations (subtract, then multiply), no FMA;
 *                  at half size the first operand is float(a + b + c + d) * 0.25f
 *   f16:           the f32 value rounded to nearest even
 * The default descriptor (zj_output_desc_default: HWC, u8, full size, all channels) is the reference's output unchanged. */
typedef enum zj_layout { ZJ_LAYOUT_HWC = 0, ZJ_LAYOUT_CHW = 1 } zj_layout;
typedef enum zj_dtype { ZJ_DTYPE_U8 = 0, ZJ_DTYPE_F16 = 1, ZJ_DTYPE_F32 = 2 } zj_dtype;
typedef struct zj_output_desc {
    uint32_t layout;      /* zj_layout                                                                          */
    uint32_t dtype;       /* zj_dtype                                                                           */
    uint32_t scale_log2;  /* 0 = full size, 1 = 2x2 box average to (width / 2) x (height / 2)                    */
    uint32_t channels;    /* 0 = every byte of the colourspace's pixel; 3 = drop the 4th byte of RGBA / RGBX    */
    float mean[4];        /* per channel, in u8 units (ignored for u8 output)                                   */
    float inv_std[4];
} zj_output_desc;
ZJ_API void zj_output_desc_default(zj_output_desc *d);
/* 1 when `d` asks for nothing but the reference's bytes */
ZJ_API int zj_output_desc_is_default(const zj_output_desc *d);
/* bytes / width / height / channels of image `img` under descriptor `d` (0 on a bad descriptor) */
ZJ_API size_t zj_consumer_output_size(const zj_image *img, const zj_output_desc *d);
ZJ_API int zj_consumer_output_shape(const zj_image *img, const zj_output_desc *d, uint32_t *out_w, uint32_t *out_h, uint32_t *out_c);
/* The consumer alone: `src_dev` = width*height*nc interleaved u8 in device memory (nc = 1, 3 or 4) -> dst_dev.  Asynchronous on `stream`. */
ZJ_API int zj_gpu_convert_device(int device, void *stream, const uint8_t *src_dev, uint32_t width, uint32_t height, uint32_t nc,
                                 const zj_output_desc *d, void *dst_dev, size_t dst_len);
/* zj_gpu_reconstruct_device followed by the consumer: out_dev[i] receives zj_consumer_output_size(&imgs[i], d) bytes.  The
 * interleaved u8 intermediate lives in a stream-ordered scratch buffer and is produced and consumed in sub-batches of at most
 * ZJ_CONSUMER_CHUNK_MB megabytes (default 1024).  Returns after `stream` has drained. */
ZJ_API int zj_gpu_reconstruct_device_ex(int device, void *stream, const zj_image *imgs, size_t n, const zj_output_desc *d,
                                        void *const *out_dev, const size_t *out_len);

/* Memory helpers so a non-CUDA host language can stage buffers without linking the CUDA runtime. */
ZJ_API int zj_gpu_pinned_alloc(size_t bytes, void **p);
ZJ_API int zj_gpu_pinned_free(void *p);
ZJ_API int zj_gpu_device_alloc(int device, size_t bytes, void **p);
ZJ_API int zj_gpu_device_free(int device, void *p);
ZJ_API int zj_gpu_memcpy_h2d(int device, void *stream, void *dst_dev, const void *src_host, size_t bytes);
ZJ_API int zj_gpu_memcpy_d2h(int device, void *stream, void *dst_host, const void *src_dev, size_t bytes);
ZJ_API int zj_gpu_memset(int device, void *stream, void *dst_dev, int value, size_t bytes);
ZJ_API int zj_gpu_stream_create(int device, void **stream);
ZJ_API int zj_gpu_stream_destroy(int device, void *stream);
ZJ_API int zj_gpu_stream_synchronize(int device, void *stream);
/* CUDA-event timing on `stream` (events see only the stream they are recorded on). */
ZJ_API int zj_gpu_event_create(int device, void **event);
ZJ_API int zj_gpu_event_record(int device, void *event, void *stream);
ZJ_API int zj_gpu_event_elapsed_ms(int device, void *start, void *stop, float *ms);
ZJ_API int zj_gpu_event_destroy(int device, void *event);

ZJ_API const char *zj_gpu_strerror(int status);
ZJ_API const char *zj_gpu_last_cuda_error(void);
/* number of kernel launches this library has issued in this process (monotonic) */
ZJ_API uint64_t zj_gpu_launch_count(void);

/* ------------------------------------------------- host front-end (C++)  */
/* The reference's host stage -- headers (src/headers.rs, src/marker.rs, src/decoder.rs:239-411) and entropy
 * decode (src/bitstream.rs, src/huffman.rs, src/mcu.rs:253-351, src/mcu_prog.rs:49-129,249-430) -- stays on
 * the CPU.  There is no Rust toolchain in this build environment, so it is restated in C++ here; in a Rust
 * deployment these symbols are not needed. */

typedef struct zj_options {       /* ZuneJpegOptions (src/options.rs:6-40), same defaults */
    uint32_t use_unsafe;          /* 1                                                   */
    uint32_t out_colorspace;      /* ZJ_CS_RGB                                           */
    uint32_t num_threads;         /* 4                                                   */
    uint32_t max_width;           /* 16384                                               */
    uint32_t max_height;          /* 16384                                               */
    uint32_t max_scans;           /* 64                                                  */
    uint32_t strict_mode;         /* 0                                                   */
    int32_t device;               /* CUDA ordinal used by zj_decoder_decode_buffer (0)   */
} zj_options;

/* DecodeErrors variants (src/errors.rs:16-43) */
typedef enum zj_decode_error_kind {
    ZJ_DE_NONE = 0,
    ZJ_DE_FORMAT = 1,
    ZJ_DE_FORMAT_STATIC = 2,
    ZJ_DE_ILLEGAL_MAGIC_BYTES = 3,
    ZJ_DE_HUFFMAN_DECODE = 4,
    ZJ_DE_ZERO_ERROR = 5,
    ZJ_DE_DQT_ERROR = 6,
    ZJ_DE_SOS_ERROR = 7,
    ZJ_DE_SOF_ERROR = 8,
    ZJ_DE_UNSUPPORTED = 9,
    ZJ_DE_MCU_ERROR = 10,
    ZJ_DE_EXHAUSTED_DATA = 11,
    ZJ_DE_LARGE_DIMENSIONS = 12,
    ZJ_DE_GPU = 13 /* not in the reference: the GPU stage failed, message = zj_gpu_strerror */
} zj_decode_error_kind;

typedef struct zj_image_info {    /* ImageInfo (src/decoder.rs:652-668) */
    uint16_t width, height;
    uint8_t pixel_density;
    uint8_t sof;                  /* SOFMarkers ordinal (src/misc.rs): 0 BaselineDct, 2 ProgressiveDctHuffman */
    uint16_t x_density, y_density;
    uint8_t components;
    uint8_t valid;                /* 0 until headers were parsed (Decoder::info() -> None, decoder.rs:210)    */
} zj_image_info;

typedef struct zj_decoder zj_decoder;

ZJ_API void zj_options_default(zj_options *o);
ZJ_API zj_decoder *zj_decoder_new(const zj_options *o);                /* Decoder::new_with_options           */
ZJ_API void zj_decoder_free(zj_decoder *d);
ZJ_API int zj_decoder_read_headers(zj_decoder *d, const uint8_t *buf, size_t len); /* Decoder::read_headers  */
ZJ_API int zj_decoder_info(const zj_decoder *d, zj_image_info *info);  /* Decoder::info                       */
ZJ_API uint32_t zj_decoder_out_colorspace(const zj_decoder *d);        /* Decoder::get_output_colorspace      */
/* Host stage only: headers + entropy decode into coefficient planes owned by the decoder (pinned when a
 * device is present).  `img` is filled with a descriptor pointing at them (valid until the next call). */
ZJ_API int zj_decoder_decode_coefficients(zj_decoder *d, const uint8_t *buf, size_t len, zj_image *img);
/* Baseline scans with restart markers (DRI) are entropy-decoded by up to `num_threads` host threads, one restart interval
 * per thread at a time (src/mcu.rs:253-351 and 386-418 run per interval); the result is kept only when every interval ended
 * exactly where the next one starts, otherwise the scan is redone by the reference's sequential loop, so the planes are
 * always what that loop produces.  Returns how many intervals the last decode ran side by side (0 = sequential loop). */
ZJ_API size_t zj_decoder_entropy_segments(const zj_decoder *d);
/* Decoder::decode_buffer: host stage, then zj_gpu_reconstruct.  *out is malloc'd (zj_buffer_free). */
ZJ_API int zj_decoder_decode_buffer(zj_decoder *d, const uint8_t *buf, size_t len, uint8_t **out,
                                    size_t *out_len);
ZJ_API void zj_buffer_free(uint8_t *p);
/* decode_into of later zune-jpeg releases: the same, into the caller's buffer of out_cap bytes (pinned memory from
 * zj_gpu_pinned_alloc makes every copy asynchronous); *out_len = bytes written.  Baseline images of >= 4 MP run as a strip
 * pipeline the way the reference's driver does (src/mcu.rs:230-369 entropy-decodes strip k+1 while its pool post-processes
 * strip k): finished strip ranges (zj_image_strip_range) are uploaded, reconstructed and downloaded while the host --
 * sequentially, or with its restart intervals side by side -- is still entropy-decoding the rest of the image. */
ZJ_API int zj_decoder_decode_into(zj_decoder *d, const uint8_t *buf, size_t len, uint8_t *out, size_t out_cap,
                                  size_t *out_len);
/* Batch front door (the reference has none: it parallelises the strips of ONE image, mcu.rs:230-369): n JPEGs are
 * decoded by `o->num_threads` host threads (0 = one per hardware thread), one image per thread at a time; each thread
 * runs the host stage, hands its planes to zj_gpu_reconstruct_submit and starts on its next image (a second set of planes)
 * while the GPU works, so entropy decoding overlaps transfer and reconstruction of this thread's and the others' images.  out[i] non-NULL on entry = caller buffer of out_len[i] bytes (pinned memory copies
 * fastest); NULL = malloc'ed here (zj_buffer_free).  status[i] = per-image zj_status; returns the number of failed
 * images (0 = all decoded) or a negative zj_status for invalid arguments. */
ZJ_API int zj_decode_batch(const zj_options *o, const uint8_t *const *bufs, const size_t *lens, size_t n,
                           uint8_t **out, size_t *out_len, int *status);
/* The DecodeErrors of image i of the LAST batch call made by the calling thread (any of the zj_decode_batch* front doors):
 * the variant (zj_decode_error_kind) and Display text Decoder::decode_buffer would have reported for that input
 * (src/errors.rs:16-113); ZJ_DE_NONE / "" for images that decoded.  Valid until the thread's next batch call. */
ZJ_API int zj_batch_error_kind(size_t i);
ZJ_API const char *zj_batch_error(size_t i);
/* zj_decode_batch keeps its workers' decoders (pinned coefficient planes sized for the largest image seen, two per host
 * thread) for the next call; this frees them. */
/* zj_decode_batch over several devices of one box: contiguous image ranges, one per device, each range decoded by its own
 * share of the host threads (o->num_threads / n_dev each; o->device is ignored) and reconstructed on its device.  With
 * fewer images than devices every image is entropy-decoded by all host threads (restart intervals side by side) and its
 * strips are spread over the devices (zj_gpu_reconstruct_multi). */
ZJ_API int zj_decode_batch_multi(const zj_options *o, const int *devices, size_t n_dev, const uint8_t *const *bufs,
                                 const size_t *lens, size_t n, uint8_t **out, size_t *out_len, int *status);
ZJ_API void zj_release_host_caches(void);
/* Device-side state kept between calls, and its release.  zj_gpu_reconstruct[_submit] keeps, per host thread that called it,
 * its staging streams and up to three device staging buffers (256 MB sub-batches by default); zj_decode_batch_gpu[_device]
 * keeps two slots of device staging memory (>= 1 GB each once used, up to the 4 / 8 GB sub-batch budget; a slot that a call
 * grew beyond ZJ_RETAIN_MB megabytes is freed when that call ends -- default 16384, i.e. never: re-allocating a 6.7 GB slot
 * was measured at up to 119 ms per call), their streams and small pinned blocks.  zj_release_device_caches frees all of
 * it except what a thread is using at that moment (it waits for a zj_decode_batch_gpu call in flight); the next call
 * allocates again.  Call it when the process wants its HBM back, e.g. a data loader that shares the GPU with a training job. */
ZJ_API void zj_release_device_caches(void);
/* The same call with the entropy stage on the GPU too, for the JPEGs that allow it: baseline scans with restart markers (DRI)
 * whose every interval ends the way the reference's sequential loop (src/mcu.rs:253-351, 386-418) ends it.  Their files are
 * uploaded instead of their coefficient planes, one GPU thread per restart interval runs the reference's bit reader and MCU
 * loop (src/bitstream.rs:159-402) into device planes, which are reconstructed in place.  All other images (no DRI,
 * progressive, header errors, an interval that ends differently) take the zj_decode_batch route, so pixels, statuses and
 * errors are those of zj_decode_batch.  *n_gpu_entropy (may be NULL) = images whose entropy stage ran on the GPU. */
ZJ_API int zj_decode_batch_gpu(const zj_options *o, const uint8_t *const *bufs, const size_t *lens, size_t n,
                               uint8_t **out, size_t *out_len, int *status, size_t *n_gpu_entropy);
/* ... and with the pixels left in device memory (no PCIe download at all: what is uploaded is the JPEG file): out_dev[i] is a
 * device buffer of out_len[i] >= width*height*components bytes (sizes: zj_decoder_read_headers + zj_decoder_info); images
 * that take the host route are uploaded after decoding.  The call returns when all pixels are in place. */
ZJ_API int zj_decode_batch_gpu_device(const zj_options *o, const uint8_t *const *bufs, const size_t *lens, size_t n,
                                      uint8_t *const *out_dev, size_t *out_len, int *status, size_t *n_gpu_entropy);
/* ... and handed to a device-side consumer (zj_output_desc above): out_dev[i] receives zj_consumer_output_size bytes in the
 * layout / type / scale of `d` (sizes before decoding: zj_decoder_read_headers + zj_decoder_info give width, height and
 * components; out_len[i] is checked).  On return out_len[i] = bytes written. */
ZJ_API int zj_decode_batch_gpu_device_ex(const zj_options *o, const uint8_t *const *bufs, const size_t *lens, size_t n,
                                         const zj_output_desc *d, void *const *out_dev, size_t *out_len, int *status,
                                         size_t *n_gpu_entropy);
/* TEST-ONLY switches for three bugs of the reference's host stage that this library reproduces by default so that its
 * coefficient planes are the reference's.  A cleared bit makes the cited lines behave like libjpeg:
 *   Q9  src/huffman.rs:249-252   fast-AC entry `(k << 10) + ...` is an i16: |k| >= 32 loses its top bits
 *   Q10 src/bitstream.rs:278-281 decode_dc refills only below 16 buffered bits but may consume 27: the reader under-runs
 *   Q11 src/mcu.rs:337-342       the MCU loop breaks on a PREFETCHED EOI: trailing components of the last MCU stay zero
 * Process-wide; takes effect for tables built / streams started afterwards.  With any bit cleared the GPU entropy route
 * (zj_decode_batch_gpu*) sends every image through the host stage.  tests/test_quirks.py is the only caller. */
enum { ZJ_QUIRK_Q9_FAST_AC_I16 = 1u, ZJ_QUIRK_Q10_DC_REFILL = 2u, ZJ_QUIRK_Q11_EOI_BREAK = 4u, ZJ_QUIRK_ALL = 7u };
ZJ_API void zj_host_set_quirks(uint32_t mask);
ZJ_API uint32_t zj_host_get_quirks(void);
ZJ_API int zj_decoder_error_kind(const zj_decoder *d);                 /* zj_decode_error_kind                */
ZJ_API const char *zj_decoder_error(const zj_decoder *d);              /* Display text of the DecodeErrors    */

#ifdef __cplusplus
}
#endif
#endif /* ZUNE_JPEG_B200_H */
