// Register-resident integer-pipe microbenchmark ("INT32 peak ... must be
// measured").  Each kind issues one instruction type (or an interleave of two) on 16
// independent accumulators per thread; inline PTX keeps nvcc from folding the chains.
// The figure reported is lane-instructions per second (an IDP.2A counts once although it
// performs two multiply-accumulates).
#include <cuda_runtime.h>

#include "internal.h"

namespace p265 {

constexpr int kAcc = 16;
constexpr int kUnroll = 8;

template <int KIND>
__device__ __forceinline__ void step(int (&x)[kAcc], int a, int b) {
#pragma unroll
    for (int i = 0; i < kAcc; i++) {
        if (KIND == 0) {
            asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
        } else if (KIND == 1) {
            asm volatile("{ .reg .s32 t; add.s32 t, %0, %1; add.s32 %0, t, %2; }" : "+r"(x[i]) : "r"(a), "r"(b));
        } else if (KIND == 2) {
            if (i & 1) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
            else asm volatile("{ .reg .s32 t; add.s32 t, %0, %1; add.s32 %0, t, %2; }" : "+r"(x[i]) : "r"(a), "r"(b));
        } else if (KIND == 3) {
            asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(x[(i + 1) % kAcc]), "r"(a));
        } else if (KIND == 4) {
            asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
        } else {
            if (i & 1) asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(x[(i + 1) % kAcc]), "r"(a));
            else asm volatile("{ .reg .s32 t; add.s32 t, %0, %1; add.s32 %0, t, %2; }" : "+r"(x[i]) : "r"(a), "r"(b));
        }
    }
}

template <int KIND>
__global__ void __launch_bounds__(256) peak_kernel(int *out, int iters, int a, int b) {
    int x[kAcc];
#pragma unroll
    for (int i = 0; i < kAcc; i++) x[i] = threadIdx.x * 7 + i;
    // keep the operands in ordinary registers (not re-read from the constant bank)
    a += (int)(threadIdx.x >> 20);
    b += (int)(threadIdx.x >> 21);
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < kUnroll; u++) step<KIND>(x, a, b);
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < kAcc; i++) s ^= x[i];
    if (s == 0x7fffffff) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // never true in practice
}

template <int KIND>
static int time_kind(p265_ctx *ctx, int *d_out, int grid, int iters, float *ms) {
    cudaEvent_t e0, e1;
    P265_CUDA(cudaEventCreate(&e0));
    P265_CUDA(cudaEventCreate(&e1));
    peak_kernel<KIND><<<grid, 256, 0, ctx->stream>>>(d_out, iters / 8, 3, 5);  // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        P265_CUDA(cudaEventRecord(e0, ctx->stream));
        peak_kernel<KIND><<<grid, 256, 0, ctx->stream>>>(d_out, iters, 3, 5);
        P265_CUDA(cudaEventRecord(e1, ctx->stream));
        P265_CUDA(cudaEventSynchronize(e1));
        float t;
        P265_CUDA(cudaEventElapsedTime(&t, e0, e1));
        best = t < best ? t : best;
    }
    P265_CUDA(cudaGetLastError());
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms = best;
    return P265_OK;
}

int run_int_peak(p265_ctx *ctx, int kind, double *ops_per_s, double *ms_out) {
    const int grid = ctx->sm_count * 8, iters = 2048;
    int *d_out = nullptr;
    P265_CUDA(cudaMalloc(&d_out, sizeof(int) * (size_t)grid * 256));
    float ms = 0;
    int rc;
    switch (kind) {
        case 0: rc = time_kind<0>(ctx, d_out, grid, iters, &ms); break;
        case 1: rc = time_kind<1>(ctx, d_out, grid, iters, &ms); break;
        case 2: rc = time_kind<2>(ctx, d_out, grid, iters, &ms); break;
        case 3: rc = time_kind<3>(ctx, d_out, grid, iters, &ms); break;
        case 4: rc = time_kind<4>(ctx, d_out, grid, iters, &ms); break;
        case 5: rc = time_kind<5>(ctx, d_out, grid, iters, &ms); break;
        default: cudaFree(d_out); return set_error(P265_EINVAL, "p265_int_peak: kind must be 0..5");
    }
    cudaFree(d_out);
    if (rc) return rc;
    // instructions per lane: kinds with the two-add sequence may fuse into one IADD3; the
    // count below is SASS-level (one IADD3 per sequence), checked in profiles/*sass*.
    const double lane_instr = (double)grid * 256.0 * iters * kUnroll * kAcc;
    *ops_per_s = lane_instr / (ms * 1e-3);
    *ms_out = ms;
    return P265_OK;
}

}  // namespace p265
