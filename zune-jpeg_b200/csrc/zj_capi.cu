// zj_capi.cu -- the C ABI of include/zune_jpeg_b200.h: descriptor validation, strip/tile planning, device
// memory plumbing and kernel launches.  No CPU fallback lives here: without a CUDA device every compute
// entry point returns ZJ_ERR_NO_DEVICE.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "../../include/zune_jpeg_b200.h"
#include "zj_device.h"

namespace zj {
cudaError_t launch_group(const DevImage *d_images, const LaunchGroup &g, cudaStream_t stream);
cudaError_t launch_convert(const ConvImage *d_images, uint32_t count, uint32_t nc, uint32_t max_ow, uint32_t max_oh, const zj_output_desc &d, cudaStream_t s);
}

using namespace zj;

static thread_local char g_cuda_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

static int cuda_fail(cudaError_t e, const char *what)
{
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", what, cudaGetErrorString(e));
    if (e == cudaErrorMemoryAllocation) return ZJ_ERR_OOM;
    return ZJ_ERR_CUDA;
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(e_, #call); } while (0)

static int num_components(uint32_t cs)  // ColorSpace::num_components, reference src/misc.rs:108-118
{
    switch (cs) {
    case ZJ_CS_RGB: case ZJ_CS_YCBCR: return 3;
    case ZJ_CS_CMYK: case ZJ_CS_YCCK: case ZJ_CS_RGBA: case ZJ_CS_RGBX: return 4;
    case ZJ_CS_GRAYSCALE: return 1;
    default: return 0;
    }
}

// --------------------------------------------------------------------------------------------- planning
struct Plan {
    int mode, variant, gray, zero_only, fast;
    uint32_t n_strips, rows, ncomp_used;
    size_t chunk[3];   // i16 per strip per component
    size_t out_size;
    DevImage dev;      // pointers filled by the caller
    uint32_t grid_tiles, grid_strips;   // launch extents (grid_strips: strips as the kernel counts them)
};

// Strip geometry of reference src/mcu.rs:139-226 (baseline) / src/mcu_prog.rs:132-203 (progressive) and the
// colour-writer constants of src/worker.rs:143-251.
static int plan_image(const zj_image *img, Plan *pl, const void *out = nullptr)
{
    if (!img) return ZJ_ERR_INVALID_ARG;
    if (img->n_comp != 1 && img->n_comp != 3) return ZJ_ERR_INVALID_ARG;
    if (img->width == 0 || img->height == 0 || img->width > 65535 || img->height > 65535) return ZJ_ERR_INVALID_ARG;
    const int nc = num_components(img->out_cs);
    if (nc == 0) return ZJ_ERR_INVALID_ARG;
    if (img->variant != ZJ_VARIANT_X86 && img->variant != ZJ_VARIANT_SCALAR) return ZJ_ERR_INVALID_ARG;
    const uint32_t h = img->comp[0].h_samp, v = img->comp[0].v_samp;
    if (img->n_comp == 1 && (h != 1 || v != 1)) return ZJ_ERR_UNSUPPORTED;  // mcu.rs:171-196 resets these before the path
    for (uint32_t z = 1; z < img->n_comp; z++)
        if (img->comp[z].h_samp != 1 || img->comp[z].v_samp != 1) return ZJ_ERR_UNSUPPORTED;  // decoder.rs:634-643
    int mode;
    if (h == 1 && v == 1) mode = MODE_NONE;
    else if (h == 2 && v == 1) mode = MODE_H;
    else if (h == 1 && v == 2) mode = MODE_V;
    else if (h == 2 && v == 2) mode = MODE_HV;
    else return ZJ_ERR_UNSUPPORTED;  // decoder.rs:512-519

    const uint32_t w = img->width, hh = img->height;
    const uint32_t mcu_x = (w + 8 * h - 1) / (8 * h), mcu_y = (hh + 8 * v - 1) / (8 * v);  // headers.rs:316-318
    for (uint32_t z = 0; z < img->n_comp; z++) {
        if (img->comp[z].width_stride != img->comp[z].h_samp * mcu_x * 8) return ZJ_ERR_INVALID_ARG;  // headers.rs:338
        for (int k = 0; k < 64; k++)
            if (img->comp[z].qt[k] < 0 || img->comp[z].qt[k] > 255) return ZJ_ERR_INVALID_ARG;  // 8-bit tables only, headers.rs:156-173
    }
    memset(pl, 0, sizeof(*pl));
    pl->mode = mode;
    pl->variant = (int)img->variant;
    uint32_t n_strips, ybr, cbr;
    switch (mode) {
    case MODE_H: n_strips = mcu_y / 2; ybr = 2; cbr = 2; break;   // mcu.rs:147-154 -- Q1: odd MCU row dropped
    case MODE_HV: n_strips = mcu_y / 2; ybr = 4; cbr = 2; break;  // mcu.rs:155-159
    case MODE_V: n_strips = mcu_y; ybr = 2; cbr = 1; break;       // mcu.rs:160-163
    default: n_strips = (hh + 7) / 8; ybr = 1; cbr = 1; break;    // mcu.rs:164-169
    }
    const uint32_t rows = 8 * h * v;
    const size_t out_chunk = (size_t)w * nc * rows;               // mcu.rs:226
    const size_t capacity = (size_t)(uint16_t)(w + 8) * (size_t)(uint16_t)(hh + 8);  // mcu.rs:198 (u16 add)
    const size_t extra = (mode != MODE_NONE ? 128u : 0u) * (size_t)hh * nc;           // mcu.rs:207
    const size_t avail = (capacity * nc + extra) / out_chunk;
    if (n_strips > avail) {
        if (img->flags & ZJ_FLAG_PROGRESSIVE) n_strips = (uint32_t)avail;  // mcu_prog.rs:206-209, zip stops
        else return ZJ_ERR_REF_PANIC;                                      // mcu.rs:354, chunks.next().unwrap()
    }
    pl->n_strips = n_strips;
    pl->grid_strips = n_strips;
    pl->rows = rows;
    pl->out_size = (size_t)w * hh * nc;
    pl->chunk[0] = (size_t)ybr * h * mcu_x * 64;
    pl->chunk[1] = pl->chunk[2] = (size_t)cbr * mcu_x * 64;

    const uint32_t in_cs = img->n_comp == 1 ? ZJ_CS_GRAYSCALE : ZJ_CS_YCBCR;
    const uint32_t oc = img->out_cs;
    uint32_t kind;
    if (oc == ZJ_CS_GRAYSCALE) kind = OUT_GRAY;                                       // worker.rs:115-118
    else if (in_cs == ZJ_CS_YCBCR && oc == ZJ_CS_YCBCR) kind = OUT_YCC;               // worker.rs:120-123
    else if (in_cs == ZJ_CS_YCBCR && (oc == ZJ_CS_RGB || oc == ZJ_CS_RGBA || oc == ZJ_CS_RGBX)) kind = OUT_RGB;  // :125-129
    else kind = OUT_ZERO;                                                             // worker.rs:131-132
    pl->gray = kind == OUT_GRAY;
    pl->zero_only = kind == OUT_ZERO;
    pl->ncomp_used = kind == OUT_GRAY ? 1u : (kind == OUT_ZERO ? 0u : 3u);  // min(in, out) comps are IDCT'd; unused results are dropped

    DevImage &d = pl->dev;
    d.width = w; d.height = hh; d.nc = (uint32_t)nc; d.out_kind = kind;
    d.mcu_x = mcu_x; d.n_strips = n_strips;
    d.Wp = img->comp[0].width_stride;
    d.W = img->n_comp == 3 ? img->comp[1].width_stride : 0;
    d.stride = w * (uint32_t)nc;
    d.T = 0xffffffffu;
    d.small_width = 0;
    d.hv_avx = (mode == MODE_HV && img->variant == ZJ_VARIANT_X86 && (size_t)16 * d.W >= 500) ? 1u : 0u;  // upsampler/avx2.rs:16
    for (uint32_t z = 0; z < img->n_comp; z++)
        for (int k = 0; k < 32; k++)
            d.qtw[z][k] = (uint32_t)img->comp[z].qt[2 * k] | ((uint32_t)img->comp[z].qt[2 * k + 1] << 24);

    if (kind == OUT_GRAY) {
        // ycbcr_to_grayscale (color_convert/scalar.rs:97-112): width_mcu = len / width must equal the strip's row
        // count, otherwise the chunking overruns the strip's output slice and the reference panics (Q7).
        const size_t len = (size_t)rows * d.Wp;
        if (n_strips > 0 && len / w != rows) return ZJ_ERR_REF_PANIC;   // (no strip, no call, no panic: the output stays zero)
        d.gray_rows_ok = 1;
        d.gray_brows = n_strips * (rows / 8);
        // X86 variant: producer / consumer kernel over pairs of block rows; SCALAR (unclamped DC-only samples): gray_kernel
        pl->fast = img->variant == ZJ_VARIANT_X86 && !getenv("ZJ_NO_FAST");
        if (pl->fast) {
            const uint32_t ucols = (d.Wp + 15) / 16;
            d.n_tiles = (ucols + ZF_XU_GRAY - 1) / ZF_XU_GRAY;
            d.tile_q = ucols / d.n_tiles;
            d.tile_r = ucols % d.n_tiles;
            pl->grid_strips = (d.gray_brows + 1) / 2;
        } else {
            d.n_tiles = (d.Wp / 8 + ZJ_THREADS - 1) / ZJ_THREADS;
        }
        pl->grid_tiles = d.n_tiles;
    } else if (kind == OUT_YCC) {
        d.n_norm = w; d.P = 3 * w;
    } else if (kind == OUT_RGB) {
        if (w < 16) {
            if (d.Wp > 16) return ZJ_ERR_REF_PANIC;  // copy_from_slice into [0;16], worker.rs:183
            d.small_width = 1;
        } else {
            const uint32_t E = d.Wp / 16 > 0 ? d.Wp / 16 - 1 : 0;  // worker.rs:171
            d.n_norm = 16 * E;
            d.P = 48 * E;
            if (d.P > d.stride) return ZJ_ERR_REF_PANIC;
            const uint32_t room = d.stride - d.P;
            const uint32_t diff = room < 64 ? 64 - room : 0;       // worker.rs:221
            d.T = d.P > diff ? d.P - diff : 0;                     // worker.rs:223
            if (d.T + 48 > d.stride) return ZJ_ERR_REF_PANIC;
        }
    }
    if (kind != OUT_GRAY) {
        // the fast kernel covers the X86 variant with word-aligned rows; everything else runs the generic kernel
        pl->fast = img->variant == ZJ_VARIANT_X86 && (kind == OUT_RGB || kind == OUT_YCC) && !d.small_width && (d.stride & 3u) == 0 &&
                   (mode != MODE_HV || d.hv_avx) && (reinterpret_cast<uintptr_t>(out) & 3) == 0 && !getenv("ZJ_NO_FAST");
        if (pl->fast) {
            const uint32_t xuv[4] = {ZF_XU_NONE, ZF_XU_H, ZF_XU_V, ZF_XU_HV};
            const uint32_t ucols = (d.Wp + 15) / 16;            // 16-sample unit columns
            d.n_tiles = (ucols + xuv[mode] - 1) / xuv[mode];
            d.tile_q = ucols / d.n_tiles;
            d.tile_r = ucols % d.n_tiles;
        } else {
            const int tmv[4] = {TM_NONE, TM_H, TM_V, TM_HV};
            d.n_tiles = (mcu_x + tmv[mode] - 1) / tmv[mode];
            d.tile_q = mcu_x / d.n_tiles;
            d.tile_r = mcu_x % d.n_tiles;
        }
        pl->grid_tiles = d.n_tiles;
        d.magic_w = d.W ? (((uint64_t)1 << 40) + d.W - 1) / d.W : 0;
    }
    return ZJ_OK;
}

static int check_buffers(const zj_image *img, const Plan &pl, const void *out, size_t out_len)
{
    for (uint32_t z = 0; z < pl.ncomp_used; z++) {
        if (!img->comp[z].coeff) return ZJ_ERR_INVALID_ARG;
        if (img->comp[z].n_i16 < (uint64_t)pl.n_strips * pl.chunk[z]) return ZJ_ERR_SHORT_PLANE;
    }
    if (!out) return ZJ_ERR_INVALID_ARG;
    if (out_len < pl.out_size) return ZJ_ERR_SHORT_OUTPUT;
    return ZJ_OK;
}

// device-resident planes are read with 128-bit accesses: every plane must start on a 16-byte boundary
static int check_device_alignment(const zj_image *img, const Plan &pl)
{
    for (uint32_t z = 0; z < pl.ncomp_used; z++)
        if (reinterpret_cast<uintptr_t>(img->comp[z].coeff) & 15) return ZJ_ERR_INVALID_ARG;
    return ZJ_OK;
}

// ------------------------------------------------------------------------------------------------ batches
struct zj_batch {
    int device;
    std::vector<LaunchGroup> groups;
    std::vector<std::pair<uint8_t *, size_t>> zero_only;  // outputs that are just memset (worker.rs:131-132)
    DevImage *d_images;
    size_t n_images;
    uint64_t algo_bytes;
    cudaStream_t pool_stream;   // non-null + pooled: descriptors come from the stream-ordered pool of this stream
    bool pooled;
};

// Stream-ordered allocations of this library (descriptor arrays, the consumers' scratch) come from a PRIVATE memory pool per
// device whose release threshold keeps its memory across synchronisations -- by default a pool returns everything to the driver
// at every synchronisation, which would re-map the staging memory on each call.  The process's default pool is left alone
// (other cudaMallocAsync users of the host application keep the behaviour they configured).
static cudaMemPool_t g_pool[64] = {};
static std::mutex g_pool_mu;
static cudaMemPool_t device_pool(int device)
{
    if (device < 0 || device >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(g_pool_mu);
    if (!g_pool[device]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        cudaMemPool_t pool = nullptr;
        if (cudaMemPoolCreate(&pool, &props) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            g_pool[device] = pool;
        } else cudaGetLastError();
    }
    return g_pool[device];
}
// cudaMallocAsync from the private pool (falls back to the default pool if it could not be created)
static cudaError_t pool_alloc(void **p, size_t bytes, int device, cudaStream_t s)
{
    cudaMemPool_t pool = device_pool(device);
    return pool ? cudaMallocFromPoolAsync(p, bytes, pool, s) : cudaMallocAsync(p, bytes, s);
}
// gives the pool's cached memory back to the driver (zj_release_device_caches)
extern "C" void zj_capi_trim_pools(void)
{
    std::lock_guard<std::mutex> lock(g_pool_mu);
    for (int d = 0; d < 64; d++) if (g_pool[d]) cudaMemPoolTrimTo(g_pool[d], 0);
    cudaGetLastError();
}

static int set_device(int device)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return ZJ_ERR_NO_DEVICE; }
    if (device < 0 || device >= n) return ZJ_ERR_NO_DEVICE;
    CU(cudaSetDevice(device));
    return ZJ_OK;
}

extern "C" {

int zj_gpu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

size_t zj_output_size(const zj_image *img)
{
    Plan pl;
    int rc = plan_image(img, &pl);
    if (rc != ZJ_OK && rc != ZJ_ERR_REF_PANIC) return 0;
    return img ? (size_t)img->width * img->height * num_components(img->out_cs) : 0;
}

int zj_validate_image(const zj_image *img)
{
    Plan pl;
    int rc = plan_image(img, &pl);
    if (rc != ZJ_OK) return rc;
    for (uint32_t z = 0; z < pl.ncomp_used; z++)
        if (img->comp[z].n_i16 < (uint64_t)pl.n_strips * pl.chunk[z]) return ZJ_ERR_SHORT_PLANE;
    return ZJ_OK;
}

// `pooled`: allocate / upload the descriptors in stream order on `ps` (no device-wide synchronisation: cudaMalloc and
// above all cudaFree serialise every thread that feeds the device, which is what zj_decode_batch's workers are)
static int batch_create_impl(int device, const zj_image *imgs, size_t n, uint8_t *const *out_dev, const size_t *out_len, zj_batch **plan,
                             bool pooled, cudaStream_t ps);

int zj_batch_create(int device, const zj_image *imgs, size_t n, uint8_t *const *out_dev, const size_t *out_len, zj_batch **plan)
{
    return batch_create_impl(device, imgs, n, out_dev, out_len, plan, false, nullptr);
}

static int batch_create_impl(int device, const zj_image *imgs, size_t n, uint8_t *const *out_dev, const size_t *out_len, zj_batch **plan,
                             bool pooled, cudaStream_t ps)
{
    if (!plan) return ZJ_ERR_INVALID_ARG;
    *plan = nullptr;
    if ((!imgs || !out_dev || !out_len) && n) return ZJ_ERR_INVALID_ARG;
    int rc = set_device(device);
    if (rc) return rc;
    std::vector<Plan> plans(n);
    for (size_t i = 0; i < n; i++) {
        rc = plan_image(&imgs[i], &plans[i], out_dev[i]);
        if (rc) return rc;
        rc = check_buffers(&imgs[i], plans[i], out_dev[i], out_len[i]);
        if (rc) return rc;
        rc = check_device_alignment(&imgs[i], plans[i]);
        if (rc) return rc;
        for (int z = 0; z < 3; z++) plans[i].dev.coeff[z] = (uint32_t)z < plans[i].ncomp_used ? imgs[i].comp[z].coeff : nullptr;
        plans[i].dev.out = out_dev[i];
    }
    zj_batch *b = new zj_batch();
    b->device = device;
    b->d_images = nullptr;
    b->n_images = 0;
    b->algo_bytes = 0;
    // group by kernel: (gray, mode, variant); stable so images keep their relative order
    std::vector<size_t> order;
    for (size_t i = 0; i < n; i++) {
        if (plans[i].zero_only) b->zero_only.push_back({out_dev[i], plans[i].out_size});
        else order.push_back(i);
        b->algo_bytes += plans[i].out_size;
        for (uint32_t z = 0; z < plans[i].ncomp_used; z++) b->algo_bytes += (uint64_t)plans[i].n_strips * plans[i].chunk[z] * 2;
    }
    auto key = [&](size_t i) { return plans[i].gray * 32 + plans[i].fast * 16 + plans[i].mode * 2 + plans[i].variant; };
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t c) { return key(a) < key(c); });
    std::vector<DevImage> host(order.size());
    for (size_t k = 0; k < order.size(); k++) host[k] = plans[order[k]].dev;
    for (size_t k = 0; k < order.size();) {
        size_t e = k;
        LaunchGroup g{};
        g.gray = plans[order[k]].gray; g.mode = plans[order[k]].mode; g.variant = plans[order[k]].variant; g.fast = plans[order[k]].fast;
        g.first = (uint32_t)k;
        while (e < order.size() && key(order[e]) == key(order[k]) && e - k < 65535) {
            g.max_tiles = std::max(g.max_tiles, plans[order[e]].grid_tiles);
            g.max_strips = std::max(g.max_strips, plans[order[e]].grid_strips);
            e++;
        }
        g.count = (uint32_t)(e - k);
        b->groups.push_back(g);
        k = e;
    }
    if (!host.empty()) {
        cudaError_t e = pooled ? pool_alloc((void **)&b->d_images, host.size() * sizeof(DevImage), device, ps)
                               : cudaMalloc(&b->d_images, host.size() * sizeof(DevImage));
        if (e != cudaSuccess) { delete b; return cuda_fail(e, "cudaMalloc(descriptors)"); }
        // (pageable source: the async copy returns once the data sits in the driver's staging buffer)
        e = pooled ? cudaMemcpyAsync(b->d_images, host.data(), host.size() * sizeof(DevImage), cudaMemcpyHostToDevice, ps)
                   : cudaMemcpy(b->d_images, host.data(), host.size() * sizeof(DevImage), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { if (pooled) cudaFreeAsync(b->d_images, ps); else cudaFree(b->d_images); delete b; return cuda_fail(e, "cudaMemcpy(descriptors)"); }
        b->n_images = host.size();
    }
    b->pooled = pooled;
    b->pool_stream = ps;
    *plan = b;
    return ZJ_OK;
}

int zj_batch_run(zj_batch *b, void *stream)
{
    if (!b) return ZJ_ERR_INVALID_ARG;
    int rc = set_device(b->device);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    for (auto &z : b->zero_only) CU(cudaMemsetAsync(z.first, 0, z.second, s));
    for (auto &g : b->groups) {
        CU(launch_group(b->d_images, g, s));
        g_launches.fetch_add(1);
    }
    return ZJ_OK;
}

int zj_batch_launches(const zj_batch *b) { return b ? (int)b->groups.size() : 0; }
uint64_t zj_batch_algorithmic_bytes(const zj_batch *b) { return b ? b->algo_bytes : 0; }

void zj_batch_destroy(zj_batch *b)
{
    if (!b) return;
    if (b->d_images) {
        cudaSetDevice(b->device);
        if (b->pooled) cudaFreeAsync(b->d_images, b->pool_stream); else cudaFree(b->d_images);
    }
    delete b;
}

int zj_gpu_reconstruct_device(int device, void *stream, const zj_image *imgs, size_t n, uint8_t *const *out_dev, const size_t *out_len)
{
    zj_batch *b = nullptr;
    int rc = zj_batch_create(device, imgs, n, out_dev, out_len, &b);
    if (rc) return rc;
    rc = zj_batch_run(b, stream);
    // the descriptor array must outlive the kernels
    if (rc == ZJ_OK) { cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream); if (e != cudaSuccess) rc = cuda_fail(e, "cudaStreamSynchronize"); }
    zj_batch_destroy(b);
    return rc;
}

// ------------------------------------------------------------------------------------- device-side consumers
void zj_output_desc_default(zj_output_desc *d)
{
    if (!d) return;
    memset(d, 0, sizeof(*d));
    for (int c = 0; c < 4; c++) d->inv_std[c] = 1.0f;
}

static int desc_check(const zj_output_desc *d)
{
    if (!d || d->layout > ZJ_LAYOUT_CHW || d->dtype > ZJ_DTYPE_F32 || d->scale_log2 > 1 || (d->channels != 0 && d->channels != 3)) return ZJ_ERR_INVALID_ARG;
    return ZJ_OK;
}

int zj_output_desc_is_default(const zj_output_desc *d)
{
    return d && d->layout == ZJ_LAYOUT_HWC && d->dtype == ZJ_DTYPE_U8 && d->scale_log2 == 0 && d->channels == 0;
}

static size_t dtype_size(uint32_t dt) { return dt == ZJ_DTYPE_U8 ? 1 : (dt == ZJ_DTYPE_F16 ? 2 : 4); }

static int conv_shape(uint32_t width, uint32_t height, uint32_t nc, const zj_output_desc *d, uint32_t *ow, uint32_t *oh, uint32_t *oc)
{
    int rc = desc_check(d);
    if (rc) return rc;
    if (nc != 1 && nc != 3 && nc != 4) return ZJ_ERR_INVALID_ARG;
    *ow = width >> d->scale_log2;
    *oh = height >> d->scale_log2;
    *oc = (d->channels == 3 && nc == 4) ? 3u : nc;
    return ZJ_OK;
}

int zj_consumer_output_shape(const zj_image *img, const zj_output_desc *d, uint32_t *out_w, uint32_t *out_h, uint32_t *out_c)
{
    if (!img || !out_w || !out_h || !out_c) return ZJ_ERR_INVALID_ARG;
    return conv_shape(img->width, img->height, (uint32_t)num_components(img->out_cs), d, out_w, out_h, out_c);
}

size_t zj_consumer_output_size(const zj_image *img, const zj_output_desc *d)
{
    uint32_t ow, oh, oc;
    if (!img || zj_output_size(img) == 0 || zj_consumer_output_shape(img, d, &ow, &oh, &oc) != ZJ_OK) return 0;
    return (size_t)ow * oh * oc * dtype_size(d->dtype);
}

int zj_gpu_convert_device(int device, void *stream, const uint8_t *src_dev, uint32_t width, uint32_t height, uint32_t nc,
                          const zj_output_desc *d, void *dst_dev, size_t dst_len)
{
    ConvImage ci{};
    int rc = conv_shape(width, height, nc, d, &ci.ow, &ci.oh, &ci.oc);
    if (rc) return rc;
    if (!src_dev || !dst_dev) return ZJ_ERR_INVALID_ARG;
    if (dst_len < (size_t)ci.ow * ci.oh * ci.oc * dtype_size(d->dtype)) return ZJ_ERR_SHORT_OUTPUT;
    rc = set_device(device);
    if (rc) return rc;
    if (ci.ow == 0 || ci.oh == 0) return ZJ_OK;
    ci.src = src_dev; ci.dst = dst_dev; ci.width = width; ci.height = height; ci.nc = nc;
    cudaStream_t s = (cudaStream_t)stream;
    ConvImage *d_ci = nullptr;
    CU(pool_alloc((void **)&d_ci, sizeof(ci), device, s));
    cudaError_t e = cudaMemcpyAsync(d_ci, &ci, sizeof(ci), cudaMemcpyHostToDevice, s);   // (pageable source: staged before the call returns)
    if (e == cudaSuccess) { e = launch_convert(d_ci, 1, nc, ci.ow, ci.oh, *d, s); g_launches.fetch_add(1); }
    cudaFreeAsync(d_ci, s);
    if (e != cudaSuccess) return cuda_fail(e, "zj_gpu_convert_device");
    return ZJ_OK;
}

// Hidden helpers for zj_decode_batch_gpu_device_ex (zj_host_decoder.cpp): scratch from this library's stream-ordered pool, and
// the consumer over many images in as few launches as their pixel formats allow.
extern "C" int zj_capi_pool_alloc(void **p, size_t bytes, int device, void *stream)
{
    int rc = set_device(device);
    if (rc) return rc;
    CU(pool_alloc(p, bytes, device, (cudaStream_t)stream));
    return ZJ_OK;
}
extern "C" void zj_capi_pool_free(void *p, void *stream) { if (p) cudaFreeAsync(p, (cudaStream_t)stream); }
extern "C" int zj_capi_convert_many(int device, void *stream, const uint8_t *const *src, const uint32_t *w, const uint32_t *h, const uint32_t *nc,
                                    void *const *dst, size_t n, const zj_output_desc *d)
{
    int rc = desc_check(d);
    if (rc) return rc;
    rc = set_device(device);
    if (rc) return rc;
    if (n == 0) return ZJ_OK;
    cudaStream_t s = (cudaStream_t)stream;
    std::vector<ConvImage> conv(n);
    for (size_t i = 0; i < n; i++) {
        ConvImage &ci = conv[i];
        ci.src = src[i]; ci.dst = dst[i]; ci.width = w[i]; ci.height = h[i]; ci.nc = nc[i];
        rc = conv_shape(w[i], h[i], nc[i], d, &ci.ow, &ci.oh, &ci.oc);
        if (rc) return rc;
    }
    ConvImage *d_conv = nullptr;
    CU(pool_alloc((void **)&d_conv, n * sizeof(ConvImage), device, s));
    cudaError_t e = cudaMemcpyAsync(d_conv, conv.data(), n * sizeof(ConvImage), cudaMemcpyHostToDevice, s);   // (pageable: staged before the call returns)
    for (size_t k = 0; k < n && e == cudaSuccess;) {
        size_t k1 = k;
        uint32_t mw = 0, mh = 0;
        while (k1 < n && conv[k1].nc == conv[k].nc && k1 - k < 65535) { mw = std::max(mw, conv[k1].ow); mh = std::max(mh, conv[k1].oh); k1++; }
        if (mw && mh) { e = launch_convert(d_conv + k, (uint32_t)(k1 - k), conv[k].nc, mw, mh, *d, s); g_launches.fetch_add(1); }
        k = k1;
    }
    cudaFreeAsync(d_conv, s);
    if (e != cudaSuccess) return cuda_fail(e, "consumer launch");
    return ZJ_OK;
}

// Sub-batch budget of the u8 intermediate (scratch memory of the call).  Measured on 128 4K images (profiles/README.md):
// sub-batches small enough to stay in the 126 MB L2 between the two kernels (48 MB: 245 GP/s) lose more to under-filled
// launches than the cache saves; 320 MB: 292 GP/s, the whole batch at once: 320 GP/s.  1 GB keeps the launches large and the
// scratch bounded.
static size_t consumer_chunk_bytes()
{
    static const size_t v = [] {
        const char *e = getenv("ZJ_CONSUMER_CHUNK_MB");
        long mb = e ? atol(e) : 1024;
        if (mb < 1) mb = 1;
        return (size_t)mb << 20;
    }();
    return v;
}

int zj_gpu_reconstruct_device_ex(int device, void *stream, const zj_image *imgs, size_t n, const zj_output_desc *d,
                                 void *const *out_dev, const size_t *out_len)
{
    int rc = desc_check(d);
    if (rc) return rc;
    if (zj_output_desc_is_default(d)) return zj_gpu_reconstruct_device(device, stream, imgs, n, reinterpret_cast<uint8_t *const *>(out_dev), out_len);
    if ((!imgs || !out_dev || !out_len) && n) return ZJ_ERR_INVALID_ARG;
    rc = set_device(device);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    // geometry, sizes, sub-batches of consecutive images
    std::vector<ConvImage> conv(n);
    std::vector<size_t> u8_size(n), u8_off(n);
    struct Chunk { size_t first, count, bytes; };
    std::vector<Chunk> chunks;
    size_t scratch_bytes = 0;
    for (size_t i = 0; i < n; i++) {
        u8_size[i] = zj_output_size(&imgs[i]);
        if (u8_size[i] == 0) { rc = zj_validate_image(&imgs[i]); return rc ? rc : (int)ZJ_ERR_INVALID_ARG; }
        ConvImage &ci = conv[i];
        ci.width = imgs[i].width; ci.height = imgs[i].height; ci.nc = (uint32_t)num_components(imgs[i].out_cs);
        rc = conv_shape(ci.width, ci.height, ci.nc, d, &ci.ow, &ci.oh, &ci.oc);
        if (rc) return rc;
        if (!out_dev[i]) return ZJ_ERR_INVALID_ARG;
        if (out_len[i] < (size_t)ci.ow * ci.oh * ci.oc * dtype_size(d->dtype)) return ZJ_ERR_SHORT_OUTPUT;
        ci.dst = out_dev[i];
        const size_t padded = (u8_size[i] + 255) & ~(size_t)255;
        if (chunks.empty() || chunks.back().bytes + padded > consumer_chunk_bytes()) chunks.push_back({i, 0, 0});
        u8_off[i] = chunks.back().bytes;
        chunks.back().bytes += padded;
        chunks.back().count++;
        scratch_bytes = std::max(scratch_bytes, chunks.back().bytes);
    }
    if (n == 0) return ZJ_OK;
    uint8_t *scratch = nullptr;
    ConvImage *d_conv = nullptr;
    CU(pool_alloc((void **)&scratch, scratch_bytes, device, s));
    cudaError_t e = pool_alloc((void **)&d_conv, n * sizeof(ConvImage), device, s);
    if (e != cudaSuccess) { cudaFreeAsync(scratch, s); return cuda_fail(e, "cudaMallocAsync(consumer descriptors)"); }
    for (size_t i = 0; i < n; i++) conv[i].src = scratch + u8_off[i];
    e = cudaMemcpyAsync(d_conv, conv.data(), n * sizeof(ConvImage), cudaMemcpyHostToDevice, s);
    std::vector<zj_batch *> batches;
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemcpyAsync(consumer descriptors)");
    for (size_t c = 0; c < chunks.size() && rc == ZJ_OK; c++) {
        const Chunk &ch = chunks[c];
        std::vector<uint8_t *> outs(ch.count);
        std::vector<size_t> lens(ch.count);
        for (size_t k = 0; k < ch.count; k++) { outs[k] = scratch + u8_off[ch.first + k]; lens[k] = u8_size[ch.first + k]; }
        zj_batch *b = nullptr;
        rc = batch_create_impl(device, imgs + ch.first, ch.count, outs.data(), lens.data(), &b, true, s);
        if (rc) break;
        batches.push_back(b);
        rc = zj_batch_run(b, s);
        // one consumer launch per run of images with the same number of source bytes per pixel
        for (size_t k = 0; k < ch.count && rc == ZJ_OK;) {
            size_t k1 = k;
            uint32_t mw = 0, mh = 0;
            while (k1 < ch.count && conv[ch.first + k1].nc == conv[ch.first + k].nc) {
                mw = std::max(mw, conv[ch.first + k1].ow); mh = std::max(mh, conv[ch.first + k1].oh);
                k1++;
            }
            if (mw && mh) {
                e = launch_convert(d_conv + ch.first + k, (uint32_t)(k1 - k), conv[ch.first + k].nc, mw, mh, *d, s);
                g_launches.fetch_add(1);
                if (e != cudaSuccess) rc = cuda_fail(e, "consumer launch");
            }
            k = k1;
        }
    }
    // the descriptor arrays and the scratch must outlive the kernels
    e = cudaStreamSynchronize(s);
    if (e != cudaSuccess && rc == ZJ_OK) rc = cuda_fail(e, "cudaStreamSynchronize");
    for (zj_batch *b : batches) zj_batch_destroy(b);
    cudaFreeAsync(d_conv, s);
    cudaFreeAsync(scratch, s);
    return rc;
}

// Staging state of zj_gpu_reconstruct_submit, leased per host thread (see there); the idle ones are what
// zj_release_device_caches frees.
constexpr int ZJ_NS_MAX = 4;
struct StreamCache {
    cudaStream_t s[64][ZJ_NS_MAX] = {};
    uint8_t *buf[64][ZJ_NS_MAX] = {};
    size_t cap[64][ZJ_NS_MAX] = {};
    cudaEvent_t ev[64][ZJ_NS_MAX][3] = {};   // per staging buffer: planes uploaded / kernels done / pixels downloaded
    bool ev_rec[64][ZJ_NS_MAX] = {};
};
struct CachePool {
    std::mutex mu;
    std::vector<StreamCache *> idle;
};
static CachePool g_stream_pool;

// frees the streams, events and device staging buffers of every cache no thread holds (hidden symbol; the exported entry
// point is zj_release_device_caches in zj_host_decoder.cpp)
extern "C" void zj_capi_release_stream_caches(void)
{
    std::vector<StreamCache *> drop;
    {
        std::lock_guard<std::mutex> lock(g_stream_pool.mu);
        drop.swap(g_stream_pool.idle);
    }
    int prev = -1;
    cudaGetDevice(&prev);
    for (StreamCache *c : drop) {
        for (int dev = 0; dev < 64; dev++) {
            bool any = false;
            for (int k = 0; k < ZJ_NS_MAX; k++) any = any || c->s[dev][k] || c->buf[dev][k];
            if (!any || cudaSetDevice(dev) != cudaSuccess) continue;
            for (int k = 0; k < ZJ_NS_MAX; k++) {
                if (c->s[dev][k]) cudaStreamSynchronize(c->s[dev][k]);
                if (c->buf[dev][k]) cudaFree(c->buf[dev][k]);
                for (int e = 0; e < 3; e++) if (c->ev[dev][k][e]) cudaEventDestroy(c->ev[dev][k][e]);
                if (c->s[dev][k]) cudaStreamDestroy(c->s[dev][k]);
            }
        }
        delete c;
    }
    if (prev >= 0) cudaSetDevice(prev);
    cudaGetLastError();
}

// Host entry point: stage planes H2D, run, copy pixels D2H.  Images are processed in sub-batches on internal streams so
// that the copies of one sub-batch overlap the kernels of the other.  Split in two halves: zj_gpu_reconstruct_submit queues
// everything and returns while the GPU works (the caller's planes and outputs stay in use), zj_gpu_reconstruct_finish waits
// and releases the per-call state; zj_gpu_reconstruct = both.
struct zj_pending {
    int device = 0;
    int ns = 0;
    cudaStream_t st[4] = {};
    cudaStream_t user = nullptr;
    cudaEvent_t done[4] = {};
    std::vector<zj_batch *> batches;
};

int zj_gpu_reconstruct_finish(zj_pending *p)
{
    if (!p) return ZJ_ERR_INVALID_ARG;
    int rc = set_device(p->device);
    for (int k = 0; k < p->ns && rc == ZJ_OK; k++) {
        // a blocking-sync event: the host thread sleeps instead of spinning, its core is free for another thread's Huffman work
        cudaError_t e = p->done[k] ? cudaEventSynchronize(p->done[k]) : cudaStreamSynchronize(p->st[k]);
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaEventSynchronize");
    }
    for (int k = 0; k < p->ns; k++) if (p->done[k]) cudaEventDestroy(p->done[k]);
    for (zj_batch *b : p->batches) zj_batch_destroy(b);
    // (user == NULL is the legacy default stream: synchronising it would serialise this thread against every blocking stream
    // of the process, and nothing of this call was queued on it -- the internal streams were waited for above)
    if (rc == ZJ_OK && p->user) { cudaError_t e = cudaStreamSynchronize(p->user); if (e != cudaSuccess) rc = cuda_fail(e, "cudaStreamSynchronize(user)"); }
    delete p;
    return rc;
}

int zj_gpu_reconstruct_submit(int device, void *stream, const zj_image *imgs, size_t n, uint8_t *const *out, const size_t *out_len, zj_pending **pending)
{
    if (!pending) return ZJ_ERR_INVALID_ARG;
    *pending = nullptr;
    if ((!imgs || !out || !out_len) && n) return ZJ_ERR_INVALID_ARG;
    int rc = set_device(device);
    if (rc) return rc;
    std::vector<Plan> plans(n);
    for (size_t i = 0; i < n; i++) {
        rc = plan_image(&imgs[i], &plans[i]);
        if (rc) return rc;
        rc = check_buffers(&imgs[i], plans[i], out[i], out_len[i]);
        if (rc) return rc;
    }
    static const bool trace = getenv("ZJ_SUBMIT_TRACE") != nullptr;
    auto now_ms = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_in = trace ? now_ms() : 0;
    auto lap = [&](const char *what) { if (trace) fprintf(stderr, "[zj submit] %-22s +%.3f ms\n", what, now_ms() - t_in); };
    cudaStream_t user = (cudaStream_t)stream;
    constexpr int NS_MAX = 4;
    const char *env_ns = getenv("ZJ_E2E_STREAMS"), *env_mb = getenv("ZJ_E2E_BUDGET_MB");
    // ZJ_E2E_PIPE=1: one stream per direction (uploads / kernels / downloads) over a ring of four staging buffers chained by
    // events, instead of whole sub-batches alternating over the streams
    const bool pipe = getenv("ZJ_E2E_PIPE") != nullptr && atoi(getenv("ZJ_E2E_PIPE")) != 0;
    const int NS = pipe ? 3 : std::max(1, std::min(NS_MAX, env_ns ? atoi(env_ns) : 3));
    cudaStream_t st[NS_MAX];
    // the staging streams are created once per host thread and device and reused: creating / destroying streams takes
    // driver-wide locks, which hurts when many threads call in (zj_decode_batch)
    // ... and so is the device staging buffer of every stream (grown on demand, reused in stream order): memory that the
    // stream-ordered pool hands from one stream to another makes the second stream wait for the first one's work, which
    // serialises the upload of one sub-batch behind the download of the previous one
    struct Lease {
        StreamCache *c = nullptr;
        ~Lease()
        {
            if (!c) return;
            std::lock_guard<std::mutex> lock(g_stream_pool.mu);
            g_stream_pool.idle.push_back(c);
        }
    };
    thread_local Lease lease;
    if (!lease.c) {
        std::lock_guard<std::mutex> lock(g_stream_pool.mu);
        if (!g_stream_pool.idle.empty()) { lease.c = g_stream_pool.idle.back(); g_stream_pool.idle.pop_back(); }
    }
    if (!lease.c) lease.c = new (std::nothrow) StreamCache;
    if (!lease.c) return ZJ_ERR_OOM;
    StreamCache &cache = *lease.c;
    if (device >= 64) return ZJ_ERR_NO_DEVICE;
    for (int k = 0; k < NS; k++) {
        if (!cache.s[device][k]) CU(cudaStreamCreateWithFlags(&cache.s[device][k], cudaStreamNonBlocking));
        st[k] = cache.s[device][k];
    }
    // everything issued on `user` before this call must be visible
    cudaEvent_t ev_in;
    CU(cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming));
    CU(cudaEventRecord(ev_in, user));
    for (int k = 0; k < NS; k++) CU(cudaStreamWaitEvent(st[k], ev_in, 0));
    cudaEventDestroy(ev_in);   // (released once the recorded work has completed)
    lap("streams + entry event");

    zj_pending *pd = new (std::nothrow) zj_pending;
    if (!pd) return ZJ_ERR_OOM;
    pd->device = device;
    pd->ns = NS;
    pd->user = user;
    for (int k = 0; k < NS; k++) pd->st[k] = st[k];

#define CUB(call) { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rc = cuda_fail(e_, #call); break; } }   // (inside the loop: fall through to the common clean-up)
    const size_t budget = (size_t)(env_mb ? std::max(16, atoi(env_mb)) : 256) << 20;  // device staging per sub-batch
    static const bool prime = getenv("ZJ_E2E_NO_PRIME") == nullptr;
    const size_t first_budget = prime ? budget / 8 : budget;
    size_t i = 0;
    int which = 0;
    rc = ZJ_OK;
    while (i < n && rc == ZJ_OK) {
        size_t j = i, bytes = 0;
        while (j < n) {
            size_t need = plans[j].out_size, coef = 0;
            for (uint32_t z = 0; z < plans[j].ncomp_used; z++) coef += (size_t)plans[j].n_strips * plans[j].chunk[z] * 2;
            need += coef + coef / 8 + 4096;      // (planes uploaded as one span may carry the gaps between them)
            // (the first sub-batch is a small one: kernels and downloads start after 1/8 of the usual upload)
            if (j > i && bytes + need > (i == 0 ? first_budget : budget)) break;
            bytes += need + 4 * 256;
            j++;
        }
        cudaStream_t s = st[pipe ? 0 : which];
        which = (which + 1) % (pipe ? NS_MAX : NS);
        const int slot = (which + (pipe ? NS_MAX : NS) - 1) % (pipe ? NS_MAX : NS);          // index of stream s / of the staging buffer
        cudaStream_t s_up = pipe ? st[0] : s, s_k = pipe ? st[1] : s, s_dn = pipe ? st[2] : s;
        cudaError_t e = cudaSuccess;
        if (pipe) {
            for (int q = 0; q < 3; q++)
                if (!cache.ev[device][slot][q]) CUB(cudaEventCreateWithFlags(&cache.ev[device][slot][q], cudaEventDisableTiming));
            if (rc != ZJ_OK) break;
            // the buffer is free once the pixels of its last sub-batch have been downloaded
            if (cache.ev_rec[device][slot]) CUB(cudaStreamWaitEvent(s_up, cache.ev[device][slot][2], 0));
        }
        if (cache.cap[device][slot] < bytes) {
            if (cache.buf[device][slot]) {
                if (pipe) { for (int q = 0; q < 3; q++) cudaStreamSynchronize(st[q]); } else cudaStreamSynchronize(s);
                cudaFree(cache.buf[device][slot]); cache.buf[device][slot] = nullptr; cache.cap[device][slot] = 0;
            }
            e = cudaMalloc((void **)&cache.buf[device][slot], bytes + bytes / 8);
            if (e != cudaSuccess) { cache.buf[device][slot] = nullptr; rc = cuda_fail(e, "cudaMalloc(staging)"); break; }
            cache.cap[device][slot] = bytes + bytes / 8;
        }
        lap("staging buffer");
        uint8_t *pool = cache.buf[device][slot];
        std::vector<zj_image> dimgs(imgs + i, imgs + j);
        std::vector<uint8_t *> douts(j - i);
        std::vector<size_t> dlens(j - i);
        size_t off = 0;
        auto take = [&](size_t nbytes) { uint8_t *p = pool + off; off += (nbytes + 255) & ~(size_t)255; return p; };
        // device addresses first, then the descriptors, then the copies: the descriptor upload comes from pageable memory, and such a
        // copy first waits for everything queued on its stream -- behind the plane uploads it would hold the host back until they
        // are done, and nothing of the next sub-batch could be queued meanwhile
        // Planes that follow each other in host memory (the host stage keeps an image's planes in one block) keep their
        // relative positions on the device and go up as ONE copy: a 4 MB chroma plane alone reaches 44 GB/s over PCIe, the
        // 25 MB of a whole 4K image 49.
        std::vector<size_t> span(j - i, 0);     // bytes of the merged copy (0 = one copy per plane)
        for (size_t k = i; k < j; k++) {
            const uint32_t nz = plans[k].ncomp_used;
            size_t nbz[3] = {0, 0, 0}, rel[3] = {0, 0, 0}, sum = 0;
            bool merge = nz > 1;
            for (uint32_t z = 0; z < nz; z++) { nbz[z] = (size_t)plans[k].n_strips * plans[k].chunk[z] * 2; sum += nbz[z]; }
            for (uint32_t z = 0; z < nz; z++) {
                const uintptr_t a0 = reinterpret_cast<uintptr_t>(imgs[k].comp[0].coeff), az = reinterpret_cast<uintptr_t>(imgs[k].comp[z].coeff);
                if (az < a0) { merge = false; break; }
                rel[z] = az - a0;
                if ((rel[z] & 15) || (z > 0 && rel[z] < rel[z - 1] + nbz[z - 1])) { merge = false; break; }
            }
            const size_t total = merge ? rel[nz - 1] + nbz[nz - 1] : 0;
            if (merge && total <= sum + sum / 8 + 4096) {
                uint8_t *base = take(total);
                for (uint32_t z = 0; z < nz; z++) dimgs[k - i].comp[z].coeff = (const int16_t *)(base + rel[z]);
                span[k - i] = total;
            } else {
                for (uint32_t z = 0; z < nz; z++) dimgs[k - i].comp[z].coeff = (const int16_t *)take(nbz[z]);
            }
            douts[k - i] = take(plans[k].out_size);
            dlens[k - i] = plans[k].out_size;
        }
        zj_batch *b = nullptr;
        rc = batch_create_impl(device, dimgs.data(), dimgs.size(), douts.data(), dlens.data(), &b, true, s_k);
        if (rc != ZJ_OK) break;
        for (size_t k = i; k < j && rc == ZJ_OK; k++) {
            if (span[k - i]) {
                e = cudaMemcpyAsync((void *)dimgs[k - i].comp[0].coeff, imgs[k].comp[0].coeff, span[k - i], cudaMemcpyHostToDevice, s_up);
                if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemcpyAsync(H2D)");
                continue;
            }
            for (uint32_t z = 0; z < plans[k].ncomp_used; z++) {
                const size_t nb = (size_t)plans[k].n_strips * plans[k].chunk[z] * 2;
                e = cudaMemcpyAsync((void *)dimgs[k - i].comp[z].coeff, imgs[k].comp[z].coeff, nb, cudaMemcpyHostToDevice, s_up);
                if (e != cudaSuccess) { rc = cuda_fail(e, "cudaMemcpyAsync(H2D)"); break; }
            }
        }
        if (rc != ZJ_OK) { zj_batch_destroy(b); break; }
        lap("descriptors + uploads");
        pd->batches.push_back(b);
        if (pipe) { CUB(cudaEventRecord(cache.ev[device][slot][0], s_up)); CUB(cudaStreamWaitEvent(s_k, cache.ev[device][slot][0], 0)); }
        rc = zj_batch_run(b, s_k);
        if (rc != ZJ_OK) break;
        if (pipe) { CUB(cudaEventRecord(cache.ev[device][slot][1], s_k)); CUB(cudaStreamWaitEvent(s_dn, cache.ev[device][slot][1], 0)); }
        for (size_t k = i; k < j; k++) {
            e = cudaMemcpyAsync(out[k], douts[k - i], plans[k].out_size, cudaMemcpyDeviceToHost, s_dn);
            if (e != cudaSuccess) { rc = cuda_fail(e, "cudaMemcpyAsync(D2H)"); break; }
        }
        if (pipe && rc == ZJ_OK) { CUB(cudaEventRecord(cache.ev[device][slot][2], s_dn)); cache.ev_rec[device][slot] = true; }
        lap("kernels + downloads");
        i = j;
    }
    if (rc == ZJ_OK) {
        for (int k = 0; k < NS; k++) {
            if (cudaEventCreateWithFlags(&pd->done[k], cudaEventDisableTiming | cudaEventBlockingSync) != cudaSuccess) { pd->done[k] = nullptr; cudaGetLastError(); continue; }
            if (cudaEventRecord(pd->done[k], st[k]) != cudaSuccess) { cudaEventDestroy(pd->done[k]); pd->done[k] = nullptr; cudaGetLastError(); }
        }
        *pending = pd;
        return ZJ_OK;
    }
#undef CUB
    const int rc_submit = rc;
    zj_gpu_reconstruct_finish(pd);   // waits for what was queued, releases the descriptors
    return rc_submit;
}

// ------------------------------------------------------------------------------------------ several devices
// The path shards with no exchange step: images are independent, and so are the strips of one image --
// every rule of worker::post_process is local to the strip it is called for (worker.rs:32-85 sees one strip's coefficients
// and one strip's output rows).  So a batch is cut into contiguous image ranges, one per device, and a batch with fewer
// images than devices is cut inside the images: contiguous STRIP ranges, each a "virtual image" whose planes start at the
// range's first strip and whose output starts at the range's first row.  One host thread per device runs its share through
// zj_gpu_reconstruct on that device's own streams; nothing is exchanged between devices.
void zj_partition(size_t n_items, size_t n_parts, size_t part, size_t *begin, size_t *end)
{
    size_t lo = 0, hi = 0;
    if (n_parts && part < n_parts) {
        const size_t base = n_items / n_parts, rem = n_items % n_parts;
        lo = part * base + std::min(part, rem);
        hi = lo + base + (part < rem ? 1 : 0);
    }
    if (begin) *begin = lo;
    if (end) *end = hi;
}

int zj_image_strip_range(const zj_image *img, uint32_t strip_begin, uint32_t strip_end, zj_image *sub, size_t *out_offset, size_t *out_bytes,
                         uint32_t *n_strips)
{
    Plan pl;
    int rc = plan_image(img, &pl);
    if (rc) return rc;
    if (n_strips) *n_strips = pl.n_strips;
    if (!sub && !out_offset && !out_bytes) return ZJ_OK;           // a query for the strip count
    if (strip_begin > strip_end || strip_end > pl.n_strips) return ZJ_ERR_INVALID_ARG;
    const size_t row_bytes = (size_t)img->width * num_components(img->out_cs);
    // the range that holds the last strip also owns the rows below it; an image without strips is one (empty) range at 0
    const bool last = strip_end == pl.n_strips && (strip_begin < strip_end || pl.n_strips == 0);
    if (strip_begin == strip_end && !last) {
        if (out_offset) *out_offset = 0;
        if (out_bytes) *out_bytes = 0;
        if (sub) { *sub = *img; sub->height = 0; }
        return ZJ_OK;
    }
    const uint64_t row0 = (uint64_t)strip_begin * pl.rows;
    if (row0 > img->height) return ZJ_ERR_INVALID_ARG;
    // the last range also owns the rows below the last strip (a partial strip, and the rows the reference never writes: Q1)
    const uint32_t h = last ? img->height - (uint32_t)row0 : (strip_end - strip_begin) * pl.rows;
    if (!last && row0 + h > img->height) return ZJ_ERR_INVALID_ARG;

    if (out_offset) *out_offset = (size_t)row0 * row_bytes;
    if (out_bytes) *out_bytes = (size_t)h * row_bytes;
    if (sub) {
        *sub = *img;
        sub->height = h;
        for (uint32_t z = 0; z < img->n_comp; z++) {
            const uint64_t skip = (uint64_t)strip_begin * pl.chunk[z];
            if (img->comp[z].coeff) sub->comp[z].coeff = img->comp[z].coeff + skip;
            sub->comp[z].n_i16 = img->comp[z].n_i16 > skip ? img->comp[z].n_i16 - skip : 0;
        }
        if (h == 0) return ZJ_OK;
        // the virtual image must plan to exactly the strips of the range (it does unless the whole image's strip count was
        // cut by the reference's output-capacity rule, mcu_prog.rs:206-209 -- widths next to 65535 only)
        Plan ps;
        rc = plan_image(sub, &ps);
        if (rc) return rc;
        if (ps.n_strips != strip_end - strip_begin) return ZJ_ERR_UNSUPPORTED;
    }
    return ZJ_OK;
}

int zj_gpu_reconstruct_multi(const int *devices, size_t n_dev, const zj_image *imgs, size_t n, uint8_t *const *out, const size_t *out_len)
{
    if (!devices || n_dev == 0 || ((!imgs || !out || !out_len) && n)) return ZJ_ERR_INVALID_ARG;
    if (n == 0) return ZJ_OK;
    // per device: a list of (virtual) images with their output slices
    struct Share { std::vector<zj_image> imgs; std::vector<uint8_t *> out; std::vector<size_t> len; int rc = ZJ_OK; };
    std::vector<Share> share(n_dev);
    bool by_strips = n < n_dev;
    if (by_strips) {
        for (size_t i = 0; i < n && by_strips; i++) {
            uint32_t ns = 0;
            int rc = zj_image_strip_range(&imgs[i], 0, 0, nullptr, nullptr, nullptr, &ns);
            if (rc) return rc;
            if (!out[i]) return ZJ_ERR_INVALID_ARG;
            if (out_len[i] < zj_output_size(&imgs[i])) return ZJ_ERR_SHORT_OUTPUT;
            std::vector<Share> trial(n_dev);
            for (size_t k = 0; k < n_dev; k++) {
                size_t s0, s1;
                zj_partition(ns, n_dev, k, &s0, &s1);
                if (s0 == s1 && !(ns == 0 && k == 0)) continue;   // (an image without strips: its never-written rows go to the first device)
                zj_image sub;
                size_t off = 0, bytes = 0;
                rc = zj_image_strip_range(&imgs[i], (uint32_t)s0, (uint32_t)s1, &sub, &off, &bytes, nullptr);
                if (rc == ZJ_ERR_UNSUPPORTED) { by_strips = false; break; }   // cannot be cut: whole images per device instead
                if (rc) return rc;
                if (bytes == 0) continue;
                share[k].imgs.push_back(sub); share[k].out.push_back(out[i] + off); share[k].len.push_back(bytes);
            }
        }
        if (!by_strips) for (auto &sh : share) { sh.imgs.clear(); sh.out.clear(); sh.len.clear(); }
    }
    if (!by_strips) {
        for (size_t k = 0; k < n_dev; k++) {
            size_t lo, hi;
            zj_partition(n, n_dev, k, &lo, &hi);
            share[k].imgs.assign(imgs + lo, imgs + hi);
            share[k].out.assign(out + lo, out + hi);
            share[k].len.assign(out_len + lo, out_len + hi);
        }
    }
    auto run = [&](size_t k) {
        Share &sh = share[k];
        if (sh.imgs.empty()) return;
        try { sh.rc = zj_gpu_reconstruct(devices[k], nullptr, sh.imgs.data(), sh.imgs.size(), sh.out.data(), sh.len.data()); }
        catch (...) { sh.rc = ZJ_ERR_OOM; }
    };
    std::vector<std::thread> pool;
    for (size_t k = 1; k < n_dev; k++) {
        try { pool.emplace_back(run, k); } catch (...) { run(k); }
    }
    run(0);
    for (auto &t : pool) t.join();
    for (size_t k = 0; k < n_dev; k++) if (share[k].rc != ZJ_OK) return share[k].rc;
    return ZJ_OK;
}

int zj_gpu_reconstruct(int device, void *stream, const zj_image *imgs, size_t n, uint8_t *const *out, const size_t *out_len)
{
    zj_pending *pd = nullptr;
    int rc = zj_gpu_reconstruct_submit(device, stream, imgs, n, out, out_len, &pd);
    if (rc != ZJ_OK) return rc;
    return zj_gpu_reconstruct_finish(pd);
}

// ------------------------------------------------------------------------------------- memory helpers
int zj_gpu_pinned_alloc(size_t bytes, void **p)
{
    if (!p) return ZJ_ERR_INVALID_ARG;
    *p = nullptr;
    if (zj_gpu_device_count() == 0) return ZJ_ERR_NO_DEVICE;
    CU(cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocPortable));
    return ZJ_OK;
}
int zj_gpu_pinned_free(void *p) { if (p) CU(cudaFreeHost(p)); return ZJ_OK; }
int zj_gpu_device_alloc(int device, size_t bytes, void **p)
{
    if (!p) return ZJ_ERR_INVALID_ARG;
    *p = nullptr;
    int rc = set_device(device);
    if (rc) return rc;
    CU(cudaMalloc(p, bytes ? bytes : 1));
    return ZJ_OK;
}
int zj_gpu_device_free(int device, void *p)
{
    int rc = set_device(device);
    if (rc) return rc;
    if (p) CU(cudaFree(p));
    return ZJ_OK;
}
int zj_gpu_memcpy_h2d(int device, void *stream, void *dst, const void *src, size_t bytes)
{
    int rc = set_device(device);
    if (rc) return rc;
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return ZJ_OK;
}
int zj_gpu_memcpy_d2h(int device, void *stream, void *dst, const void *src, size_t bytes)
{
    int rc = set_device(device);
    if (rc) return rc;
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return ZJ_OK;
}
int zj_gpu_memset(int device, void *stream, void *dst, int value, size_t bytes)
{
    int rc = set_device(device);
    if (rc) return rc;
    CU(cudaMemsetAsync(dst, value, bytes, (cudaStream_t)stream));
    return ZJ_OK;
}
int zj_gpu_stream_create(int device, void **stream)
{
    if (!stream) return ZJ_ERR_INVALID_ARG;
    int rc = set_device(device);
    if (rc) return rc;
    cudaStream_t s;
    CU(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = (void *)s;
    return ZJ_OK;
}
int zj_gpu_stream_destroy(int device, void *stream)
{
    int rc = set_device(device);
    if (rc) return rc;
    CU(cudaStreamDestroy((cudaStream_t)stream));
    return ZJ_OK;
}
int zj_gpu_stream_synchronize(int device, void *stream)
{
    int rc = set_device(device);
    if (rc) return rc;
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    return ZJ_OK;
}
int zj_gpu_event_create(int device, void **event)
{
    if (!event) return ZJ_ERR_INVALID_ARG;
    int rc = set_device(device);
    if (rc) return rc;
    cudaEvent_t e;
    CU(cudaEventCreate(&e));
    *event = (void *)e;
    return ZJ_OK;
}
int zj_gpu_event_record(int device, void *event, void *stream)
{
    int rc = set_device(device);
    if (rc) return rc;
    CU(cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream));
    return ZJ_OK;
}
int zj_gpu_event_elapsed_ms(int device, void *start, void *stop, float *ms)
{
    if (!ms) return ZJ_ERR_INVALID_ARG;
    int rc = set_device(device);
    if (rc) return rc;
    CU(cudaEventSynchronize((cudaEvent_t)stop));
    CU(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
    return ZJ_OK;
}
int zj_gpu_event_destroy(int device, void *event)
{
    int rc = set_device(device);
    if (rc) return rc;
    CU(cudaEventDestroy((cudaEvent_t)event));
    return ZJ_OK;
}

const char *zj_gpu_strerror(int status)
{
    switch (status) {
    case ZJ_OK: return "ok";
    case ZJ_ERR_INVALID_ARG: return "invalid argument or inconsistent image descriptor";
    case ZJ_ERR_UNSUPPORTED: return "Unknown down-sampling method, cannot continue";  // decoder.rs:512-519
    case ZJ_ERR_SHORT_PLANE: return "coefficient plane smaller than the strips it must feed";
    case ZJ_ERR_SHORT_OUTPUT: return "output buffer smaller than width*height*components";
    case ZJ_ERR_REF_PANIC: return "the reference decoder panics on this geometry";
    case ZJ_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU fallback)";
    case ZJ_ERR_CUDA: return "CUDA runtime error (see zj_gpu_last_cuda_error)";
    case ZJ_ERR_OOM: return "out of device or pinned memory";
    case ZJ_ERR_DECODE: return "JPEG header / entropy decode error (see zj_decoder_error)";
    default: return "unknown status";
    }
}
const char *zj_gpu_last_cuda_error(void) { return g_cuda_err; }
uint64_t zj_gpu_launch_count(void) { return g_launches.load(); }

}  // extern "C"
