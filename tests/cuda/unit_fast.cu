// Device-side unit check of the packed fast-path helpers against scalar arithmetic (debug aid).
#include <cstdio>
#include <cstdlib>
#include "../../zune-jpeg_b200/csrc/zj_kernels.cu"
using namespace zj;
__global__ void k(const int *R6, const int *ycc, int *bad)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int *R = R6 + 6 * t;
    u32 h = R[0] | (R[5] << 16), a = R[1] | (R[2] << 16), b = R[3] | (R[4] << 16);
    u32 o[4];
    hfilter8(h, a, b, o);
    for (int i = 0; i < 4; i++) {
        int w0 = (3 * R[i + 1] + R[i] + 2) >> 2, w1 = (3 * R[i + 1] + R[i + 2] + 2) >> 2;
        if ((int)(o[i] & 0xffff) != w0 || (int)(o[i] >> 16) != w1) { atomicAdd(bad, 1); if (t < 4) printf("hf t=%d i=%d got %u %u want %d %d\n", t, i, o[i] & 0xffff, o[i] >> 16, w0, w1); }
    }
    const int *c = ycc + 6 * t;
    u32 r, g, bb;
    convert_pair(c[0] | (c[1] << 16), c[2] | (c[3] << 16), c[4] | (c[5] << 16), r, g, bb);
    for (int l = 0; l < 2; l++) {
        u32 pr, pg, pb;
        ycc_to_rgb(c[l], c[2 + l], c[4 + l], pr, pg, pb);
        u32 gr = (r >> (16 * l)) & 0xff, gg = (g >> (16 * l)) & 0xff, gb = (bb >> (16 * l)) & 0xff;
        if (gr != pr || gg != pg || gb != pb) { atomicAdd(bad + 1, 1); if (atomicAdd(bad + 2, 1) < 12) printf("cv t=%d l=%d got %u %u %u want %u %u %u (y %d cb %d cr %d)\n", t, l, gr, gg, gb, pr, pg, pb, c[l], c[2+l], c[4+l]); }
    }
}
int main()
{
    const int N = 1 << 16;
    int *hR = (int *)malloc(N * 6 * 4), *hC = (int *)malloc(N * 6 * 4);
    srand(1);
    for (int i = 0; i < N * 6; i++) { hR[i] = rand() % 385; hC[i] = (i % 6 < 2) ? rand() % 256 : rand() % 288; }
    int *dR, *dC, *dbad;
    cudaMalloc(&dR, N * 24); cudaMalloc(&dC, N * 24); cudaMalloc(&dbad, 16);
    cudaMemcpy(dR, hR, N * 24, cudaMemcpyHostToDevice); cudaMemcpy(dC, hC, N * 24, cudaMemcpyHostToDevice);
    cudaMemset(dbad, 0, 16);
    k<<<N / 128, 128>>>(dR, dC, dbad);
    int bad[2];
    cudaMemcpy(bad, dbad, 8, cudaMemcpyDeviceToHost);
    printf("hfilter8 mismatches %d, convert_pair mismatches %d (%s)\n", bad[0], bad[1], cudaGetErrorString(cudaGetLastError()));
    return bad[0] || bad[1];
}
