// Throughput microbenchmark of the integer instructions the kernels lean on (warp-instructions / clk / SM).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define REP 64
#define ITERS 2048
template <int OP> __device__ __forceinline__ uint32_t op(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    if (OP == 0) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    else if (OP == 1) asm volatile("mad.hi.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    else if (OP == 2) asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    else if (OP == 3) asm volatile("shr.s32 %0, %1, 10;" : "=r"(d) : "r"(a));
    else if (OP == 4) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    else if (OP == 5) asm volatile("prmt.b32 %0, %1, %2, 0x5410;" : "=r"(d) : "r"(a), "r"(b));
    else if (OP == 6) asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    else if (OP == 7) asm volatile("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    else if (OP == 8) asm volatile("min.relu.s16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    else if (OP == 9) asm volatile("add.s16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    else if (OP == 10) asm volatile("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    else if (OP == 11) asm volatile("shf.l.wrap.b32 %0, %1, %2, 7;" : "=r"(d) : "r"(a), "r"(b));
    else if (OP == 12) asm volatile("min.s32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    else if (OP == 13) { asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); uint32_t e; asm volatile("add.u32 %0, %1, %2;" : "=r"(e) : "r"(d), "r"(b)); d = e; }  // 1:1 fma:alu mix
    else if (OP == 14) { asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); uint32_t e; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(e) : "r"(d), "r"(b), "r"(c)); d = e; }
    else if (OP == 16) { asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); uint32_t e; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(e) : "r"(d), "r"(b), "r"(c)); d = e; }
    else if (OP == 17) { asm volatile("prmt.b32 %0, %1, %2, 0x5410;" : "=r"(d) : "r"(a), "r"(b)); uint32_t e; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(e) : "r"(d), "r"(b), "r"(c)); d = e; }
    else if (OP == 18) { asm volatile("prmt.b32 %0, %1, %2, 0x5410;" : "=r"(d) : "r"(a), "r"(b)); uint32_t e; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(e) : "r"(d), "r"(b), "r"(c)); d = e; }
    else if (OP == 19) { asm volatile("min.relu.s16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); uint32_t e; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(e) : "r"(d), "r"(b), "r"(c)); d = e; }
    else if (OP == 20) { asm volatile("min.relu.s16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); uint32_t e; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(e) : "r"(d), "r"(b), "r"(c)); d = e; }
    else if (OP == 21) { asm volatile("add.s16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); uint32_t e; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(e) : "r"(d), "r"(b), "r"(c)); d = e; }
    else if (OP == 22) { asm volatile("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); uint32_t e; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(e) : "r"(d), "r"(b), "r"(c)); d = e; }
    else if (OP == 23) { asm volatile("shr.s32 %0, %1, 3;" : "=r"(d) : "r"(a)); uint32_t e; asm volatile("xor.b32 %0, %1, %2;" : "=r"(e) : "r"(d), "r"(b)); d = e; }
    else if (OP == 24) { asm volatile("shr.s32 %0, %1, 3;" : "=r"(d) : "r"(a)); uint32_t e; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(e) : "r"(d), "r"(b), "r"(c)); d = e; }
    else if (OP == 25) { asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); uint32_t e; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(e) : "r"(d), "r"(b), "r"(c)); d = e; }
    else if (OP == 26) asm volatile("mad.lo.u32 %0, %1, 2217, %2;" : "=r"(d) : "r"(a), "r"(c));                 // IMAD R, R, imm, R
    else if (OP == 27) { asm volatile("mad.lo.u32 %0, %1, 2217, %2;" : "=r"(d) : "r"(a), "r"(c)); uint32_t e; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(e) : "r"(d), "r"(b), "r"(c)); d = e; }
    else if (OP == 28) { asm volatile("mad.lo.u32 %0, %1, 2217, %2;" : "=r"(d) : "r"(a), "r"(c)); uint32_t e; asm volatile("mad.lo.u32 %0, %1, 3135, %2;" : "=r"(e) : "r"(d), "r"(b)); uint32_t f; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(f) : "r"(e), "r"(b), "r"(c)); d = f; }   // 2 IMAD-imm : 1 LOP3
    else if (OP == 29) asm volatile("add.u32 %0, %1, 0x511f511f;" : "=r"(d) : "r"(a));                              // add with a 32-bit immediate
    else if (OP == 30) asm volatile("shr.u32 %0, %1, 2;" : "=r"(d) : "r"(a));
    else if (OP == 31) { asm volatile("shr.u32 %0, %1, 2;" : "=r"(d) : "r"(a)); uint32_t e; asm volatile("and.b32 %0, %1, 0x01ff01ff;" : "=r"(e) : "r"(d)); d = e; }
    else if (OP == 32) { uint32_t e; asm volatile("add.u32 %0, %1, %2;" : "=r"(e) : "r"(a), "r"(b)); asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(e), "r"(c)); }   // a + b + c (IADD3 r,r,r)
    else if (OP == 33) asm volatile("mul.lo.u32 %0, %1, 45;" : "=r"(d) : "r"(a));                                     // IMAD R, R, imm, RZ
    else if (OP == 34) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));                // FFMA r,r,r (bit patterns as floats)
    else if (OP == 35) asm volatile("fma.rn.f32 %0, %1, 0f3FC00000, %2;" : "=r"(d) : "r"(a), "r"(c));                // FFMA r,imm,r
    else if (OP == 36) { asm volatile("fma.rn.f32 %0, %1, 0f3FC00000, %2;" : "=r"(d) : "r"(a), "r"(c)); uint32_t e; asm volatile("mad.lo.u32 %0, %1, 2217, %2;" : "=r"(e) : "r"(d), "r"(c)); d = e; }   // FFMA-imm + IMAD-imm
    else if (OP == 37) { asm volatile("fma.rn.f32 %0, %1, 0f3FC00000, %2;" : "=r"(d) : "r"(a), "r"(c)); uint32_t e; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(e) : "r"(d), "r"(b), "r"(c)); d = e; }   // FFMA-imm + LOP3
    else if (OP == 38) { asm volatile("fma.rn.f32 %0, %1, 0f3FC00000, %2;" : "=r"(d) : "r"(a), "r"(c)); uint32_t e; asm volatile("mad.lo.u32 %0, %1, 2217, %2;" : "=r"(e) : "r"(d), "r"(c)); uint32_t f; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(f) : "r"(e), "r"(b), "r"(c)); d = f; }   // FFMA-imm + IMAD-imm + LOP3
    else if (OP == 39) asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));              // HFMA2 r,r,r
    else if (OP == 40) { asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); uint32_t e; asm volatile("mad.lo.u32 %0, %1, 2217, %2;" : "=r"(e) : "r"(d), "r"(c)); uint32_t f; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(f) : "r"(e), "r"(b), "r"(c)); d = f; }   // HFMA2 + IMAD-imm + LOP3
    else if (OP == 41) asm volatile("shl.b32 %0, %1, 5;" : "=r"(d) : "r"(a));
    else asm volatile("sub.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
template <int OP> __global__ void k(uint32_t *out, uint32_t seed, long long *cyc)
{
    uint32_t r[8];
    for (int i = 0; i < 8; i++) r[i] = seed * (threadIdx.x + 1) + i;
    uint32_t b = seed | 1, c = seed + 3;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int j = 0; j < REP / 8; j++) {
#pragma unroll
            for (int i = 0; i < 8; i++) r[i] = op<OP>(r[i], b, c);
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
    for (int i = 0; i < 8; i++) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP> void run(const char *name, int mult)
{
    uint32_t *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    k<OP><<<148, 1024>>>(out, 12345, cyc);  // 32 warps / SM = 8 per SMSP
    cudaDeviceSynchronize();
    k<OP><<<148, 1024>>>(out, 12345, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double winst = 32.0 * ITERS * REP * mult;  // warp-instructions per SM
    printf("%-28s %.2f warp-inst/clk/SM  (%.2f per SMSP)\n", name, winst / h, winst / h / 4);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    run<0>("IMAD (mad.lo)", 1); run<1>("IMAD.HI (mad.hi.s32)", 1); run<2>("IADD", 1); run<3>("SHF.R.S32 (shr.s32)", 1);
    run<4>("LOP3", 1); run<5>("PRMT", 1); run<6>("IDP.2A", 1); run<7>("IDP.4A", 1); run<8>("VIMNMX.S16x2.RELU", 1);
    run<9>("VIADD.16x2", 1); run<10>("I2IP.U8.S32.SAT", 1); run<11>("SHF funnel", 1); run<12>("IMNMX (min.s32)", 1);
    run<13>("IMAD+IADD mix", 2); run<14>("IDP.2A+IMAD mix", 2); run<15>("ISUB", 1);
    run<16>("IDP.2A+LOP3 mix", 2); run<17>("PRMT+LOP3 mix", 2); run<18>("PRMT+IMAD mix", 2); run<19>("VIMNMX2+LOP3 mix", 2); run<20>("VIMNMX2+IMAD mix", 2);
    run<21>("VIADD2+IMAD mix", 2); run<22>("I2IP+IMAD mix", 2); run<23>("SHR+XOR mix", 2); run<24>("SHR+IMAD mix", 2); run<25>("IADD+LOP3 mix", 2);
    run<26>("IMAD r,imm,r", 1); run<33>("IMAD r,imm,RZ (mul)", 1); run<29>("add r,imm32", 1); run<30>("SHF.R.U32 imm", 1); run<41>("SHL imm", 1); run<32>("a+b+c", 1);
    run<27>("IMAD-imm + LOP3", 2); run<28>("2 IMAD-imm + LOP3", 3); run<31>("SHR + AND-imm", 2);
    run<34>("FFMA r,r,r", 1); run<35>("FFMA r,imm,r", 1); run<36>("FFMA-imm + IMAD-imm", 2); run<37>("FFMA-imm + LOP3", 2); run<38>("FFMA-imm + IMAD-imm + LOP3", 3);
    run<39>("HFMA2 r,r,r", 1); run<40>("HFMA2 + IMAD-imm + LOP3", 3);
    return 0;
}
