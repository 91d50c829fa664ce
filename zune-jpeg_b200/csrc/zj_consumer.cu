// zj_consumer.cu -- device-side consumers of the reconstructed pixels: the pixels the fused kernels
// wrote (u8, interleaved, exactly the bytes of the reference's decode_buffer) are turned into what a GPU consumer reads --
// planar (CHW) or interleaved (HWC) u8 / f16 / f32, per-channel `(x - mean) * inv_std`, optionally a 2x2 box down-scale,
// optionally without the alpha / padding byte of "RGBA" / "RGBX" -- without leaving the device.  The reference's writers
// that this extends are the per-colourspace dispatch of src/worker.rs:113-133 and the interleaving stores of
// src/color_convert/scalar.rs:52-169: every value here is a function of those writers' exact u8 results, so the u8 / HWC /
// full-size descriptor stays bit-identical to the reference and the other descriptors are specified (and tested) as
//     f = (float(u8) - mean[c]) * inv_std[c]                 (two IEEE fp32 operations, no fused multiply-add)
//     u8 at half size = (a + b + c + d + 2) >> 2,  f at half size = (float(a + b + c + d) * 0.25f - mean[c]) * inv_std[c]
//     f16 = round-to-nearest-even of the fp32 value
// applied to the oracle's bytes.
//
// One thread produces 8 consecutive output pixels of one row (all channels): its inputs are 8 * nc (or 2 rows of 16 * nc)
// contiguous bytes, read as 32-bit words when the row allows it, and its outputs are 16-byte stores when the row allows it.
// HBM-bound: 3 B/px in + 6 B/px out for RGB -> f16 (measured: 6.3 TB/s of the 6.5 TB/s copy bandwidth).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <vector>

#include "../../include/zune_jpeg_b200.h"
#include "zj_device.h"

namespace zj {

struct ConvParams {
    float mean[4], inv_std[4];
};

typedef unsigned int u32;

template <typename T> __device__ __forceinline__ T cvt(float f);
template <> __device__ __forceinline__ float cvt<float>(float f) { return f; }
template <> __device__ __forceinline__ uint8_t cvt<uint8_t>(float) { return 0; }   // (never taken: u8 outputs do not go through floats)
template <> __device__ __forceinline__ __half cvt<__half>(float f) { return __float2half_rn(f); }

// 8 elements of type T to `p`: one or two 16-byte stores when `vec`, element stores otherwise (n = valid elements)
template <typename T>
__device__ __forceinline__ void store8(T *p, const T (&v)[8], const int n, const bool vec)
{
    if (vec && n == 8) {
        if (sizeof(T) == 1) {
            uint2 w;
            memcpy(&w, v, 8);
            *reinterpret_cast<uint2 *>(p) = w;
        } else if (sizeof(T) == 2) {
            uint4 w;
            memcpy(&w, v, 16);
            *reinterpret_cast<uint4 *>(p) = w;
        } else {
            uint4 w0, w1;
            memcpy(&w0, v, 16);
            memcpy(&w1, reinterpret_cast<const char *>(v) + 16, 16);
            reinterpret_cast<uint4 *>(p)[0] = w0;
            reinterpret_cast<uint4 *>(p)[1] = w1;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++)
            if (i < n) p[i] = v[i];
    }
}

// NB contiguous bytes from `p` into words (little endian): aligned 32-bit loads when possible
template <int NW>
__device__ __forceinline__ void load_bytes(const uint8_t *p, const int nbytes, u32 (&w)[NW])
{
    if ((reinterpret_cast<uintptr_t>(p) & 3) == 0 && nbytes == NW * 4) {
#pragma unroll
        for (int k = 0; k < NW; k++) w[k] = __ldg(reinterpret_cast<const u32 *>(p) + k);
    } else {
#pragma unroll
        for (int k = 0; k < NW; k++) {
            u32 v = 0;
#pragma unroll
            for (int b = 0; b < 4; b++)
                if (4 * k + b < nbytes) v |= (u32)__ldg(p + 4 * k + b) << (8 * b);
            w[k] = v;
        }
    }
}
template <int NW> __device__ __forceinline__ u32 byte_at(const u32 (&w)[NW], const int k) { return (w[k >> 2] >> (8 * (k & 3))) & 0xffu; }

// T: output element (uint8_t, __half, float); CHW: planar output; HALF: 2x2 box down-scale; NC: source bytes per pixel;
// OC: channels written (NC, or 3 of 4)
template <typename T, bool CHW, bool HALF, int NC, int OC>
__global__ void __launch_bounds__(256) convert_kernel(const ConvImage *__restrict__ images, const ConvParams prm)
{
    const ConvImage im = images[blockIdx.z];
    const u32 y = blockIdx.y;
    if (y >= im.oh) return;
    const u32 x0 = 8u * (blockIdx.x * 256u + threadIdx.x);
    if (x0 >= im.ow) return;
    const int n = (int)min(8u, im.ow - x0);
    constexpr int S = HALF ? 2 : 1;
    constexpr int NW = 8 * S * NC / 4;   // words per source row segment
    u32 sum[8][NC];                      // per output pixel and channel: the u8 (or the sum of the four)
    {
        const size_t srow = (size_t)im.width * NC;
        const uint8_t *p = im.src + (size_t)(S * y) * srow + (size_t)(S * x0) * NC;
        u32 w[NW];
        load_bytes<NW>(p, n * S * NC, w);
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int c = 0; c < NC; c++)
                sum[i][c] = HALF ? byte_at<NW>(w, (2 * i) * NC + c) + byte_at<NW>(w, (2 * i + 1) * NC + c) : byte_at<NW>(w, i * NC + c);
        if (HALF) {
            load_bytes<NW>(p + srow, n * S * NC, w);
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int c = 0; c < NC; c++) sum[i][c] += byte_at<NW>(w, (2 * i) * NC + c) + byte_at<NW>(w, (2 * i + 1) * NC + c);
        }
    }
    T *const dst = reinterpret_cast<T *>(im.dst);
    const size_t plane = (size_t)im.ow * im.oh;
    if (CHW) {
#pragma unroll
        for (int c = 0; c < OC; c++) {
            T v[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (sizeof(T) == 1) v[i] = (T)(HALF ? (sum[i][c] + 2u) >> 2 : sum[i][c]);
                else {
                    const float f = HALF ? __fmul_rn((float)sum[i][c], 0.25f) : (float)sum[i][c];
                    v[i] = cvt<T>(__fmul_rn(__fsub_rn(f, prm.mean[c]), prm.inv_std[c]));
                }
            }
            T *q = dst + (size_t)c * plane + (size_t)y * im.ow + x0;
            store8<T>(q, v, n, (reinterpret_cast<uintptr_t>(q) & 15) == 0);
        }
    } else {
        // interleaved: 8 * OC consecutive elements, stored in OC groups of 8
        T *q = dst + ((size_t)y * im.ow + x0) * OC;
        const bool vec = (reinterpret_cast<uintptr_t>(q) & 15) == 0 && n == 8;
#pragma unroll
        for (int g = 0; g < OC; g++) {
            T v[8];
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const int px = (8 * g + e) / OC, ch = (8 * g + e) % OC;   // compile-time after unrolling
                const u32 s = sum[px][ch];
                if (sizeof(T) == 1) v[e] = (T)(HALF ? (s + 2u) >> 2 : s);
                else {
                    const float f = HALF ? __fmul_rn((float)s, 0.25f) : (float)s;
                    v[e] = cvt<T>(__fmul_rn(__fsub_rn(f, prm.mean[ch]), prm.inv_std[ch]));
                }
            }
            const int left = n * OC - 8 * g;   // valid elements of this group
            if (left <= 0) break;
            store8<T>(q + 8 * g, v, left < 8 ? left : 8, vec);
        }
    }
}

template <typename T, bool CHW, bool HALF>
static cudaError_t launch_nc(const ConvImage *d_images, uint32_t nc, uint32_t oc, dim3 grid, const ConvParams &prm, cudaStream_t s)
{
    if (nc == 1) convert_kernel<T, CHW, HALF, 1, 1><<<grid, 256, 0, s>>>(d_images, prm);
    else if (nc == 3) convert_kernel<T, CHW, HALF, 3, 3><<<grid, 256, 0, s>>>(d_images, prm);
    else if (oc == 3) convert_kernel<T, CHW, HALF, 4, 3><<<grid, 256, 0, s>>>(d_images, prm);
    else convert_kernel<T, CHW, HALF, 4, 4><<<grid, 256, 0, s>>>(d_images, prm);
    return cudaGetLastError();
}

template <typename T>
static cudaError_t launch_t(const ConvImage *d_images, uint32_t nc, uint32_t oc, bool chw, bool half, dim3 grid, const ConvParams &prm, cudaStream_t s)
{
    if (chw) return half ? launch_nc<T, true, true>(d_images, nc, oc, grid, prm, s) : launch_nc<T, true, false>(d_images, nc, oc, grid, prm, s);
    return half ? launch_nc<T, false, true>(d_images, nc, oc, grid, prm, s) : launch_nc<T, false, false>(d_images, nc, oc, grid, prm, s);
}

// images [0, count) of d_images share nc; max_ow / max_oh bound the grid
cudaError_t launch_convert(const ConvImage *d_images, uint32_t count, uint32_t nc, uint32_t max_ow, uint32_t max_oh, const zj_output_desc &d, cudaStream_t s)
{
    ConvParams prm;
    for (int c = 0; c < 4; c++) { prm.mean[c] = d.mean[c]; prm.inv_std[c] = d.inv_std[c]; }
    dim3 grid((max_ow + 8 * 256 - 1) / (8 * 256), max_oh, count);
    const bool chw = d.layout == ZJ_LAYOUT_CHW, half = d.scale_log2 == 1;
    const uint32_t oc = (d.channels == 3 && nc == 4) ? 3u : nc;
    switch (d.dtype) {
    case ZJ_DTYPE_U8: return launch_t<uint8_t>(d_images, nc, oc, chw, half, grid, prm, s);
    case ZJ_DTYPE_F16: return launch_t<__half>(d_images, nc, oc, chw, half, grid, prm, s);
    default: return launch_t<float>(d_images, nc, oc, chw, half, grid, prm, s);
    }
}

}  // namespace zj
