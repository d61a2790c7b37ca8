// I-cache probe: straight-line integer code of increasing size executed by warps that
// are deliberately de-synchronised (like the residual kernel's warps are).  Prints
// instructions per clock per SM for each body size.   nvcc -arch=sm_100a -O3 -o probe
#include <cstdio>
#include <cuda_runtime.h>

template <int BODY>  // BODY = number of 8-instruction groups in the straight-line body
__global__ void __launch_bounds__(64) probe(int *out, int iters, int a, int b, int skew) {
    int x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    a += threadIdx.x >> 20; b += threadIdx.x >> 21;
    // de-synchronise warps: each warp spins a different amount first
    const int w = (blockIdx.x * 2 + (threadIdx.x >> 5)) % 16;
    for (int s = 0; s < w * skew; s++) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x0) : "r"(a), "r"(b));
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int g = 0; g < BODY; g++) {
            asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x0) : "r"(a), "r"(b));
            asm volatile("add.s32 %0, %0, %1;" : "+r"(x1) : "r"(a));
            asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %0;" : "+r"(x2) : "r"(x3), "r"(a));
            asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x3) : "r"(a), "r"(b));
            asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x4) : "r"(a), "r"(b));
            asm volatile("add.s32 %0, %0, %1;" : "+r"(x5) : "r"(b));
            asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %0;" : "+r"(x6) : "r"(x7), "r"(a));
            asm volatile("add.s32 %0, %0, %1;" : "+r"(x7) : "r"(a));
        }
    }
    int s = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
    if (s == 0x7fffffff) out[0] = s;
}

template <int BODY>
void run(int *d, int sms, int ctas_per_sm, int skew) {
    const long total_groups = 1 << 15;  // same work for every body size
    const int iters = (int)(total_groups / BODY);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<BODY><<<sms * ctas_per_sm, 64>>>(d, iters / 8, 3, 5, skew);
    cudaEventRecord(e0);
    probe<BODY><<<sms * ctas_per_sm, 64>>>(d, iters, 3, 5, skew);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double instr = (double)sms * ctas_per_sm * 2 * (double)iters * BODY * 8;  // warp-instructions
    const double cycles = ms * 1e-3 * clk * 1e3;
    printf("body %6d instr (%4d KB)  warps/SM %2d  skew %4d : %.3f warp-instr/clk/SM  (%.3f per SMSP)  %.3f ms\n",
           BODY * 8, BODY * 8 * 16 / 1024, ctas_per_sm * 2, skew, instr / cycles / sms, instr / cycles / sms / 4, ms);
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int *d; cudaMalloc(&d, 4);
    for (int skew : {0, 997}) {
        for (int cps : {4, 9, 16}) {
            run<16>(d, sms, cps, skew);
            run<64>(d, sms, cps, skew);
            run<128>(d, sms, cps, skew);
            run<256>(d, sms, cps, skew);
            run<512>(d, sms, cps, skew);
            run<1024>(d, sms, cps, skew);
            run<2048>(d, sms, cps, skew);
        }
    }
    return 0;
}
