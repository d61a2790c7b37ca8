// Reconstruction: rec = Clip1(pred + residual) over whole planes (reconstruction.py:4-27,
// H.265 8.6.7) -- the step between the residual planes and the SAO input (SURVEY.md 8(f)
// rank 1).  Element-wise and HBM-bound: 16-byte loads of the prediction and residual rows,
// packed 2 x 16-bit arithmetic (the int16 residual is first clamped to +-maxVal so the
// half-word add cannot overflow; VIADDMNMX.S16x2.RELU then adds and clips in one go).
#include <cuda_runtime.h>

#include "internal.h"

namespace p265 {

struct ReconArgs {
    const void *pred;
    const int16_t *res;
    void *rec;
    int64_t plane_off[3];
    int64_t pic_stride;
    int32_t width, height, stride_y, stride_c;
    int32_t bit_depth_y, bit_depth_c;
};

__device__ __forceinline__ uint32_t recon_word(uint32_t p, uint32_t r, uint32_t maxv2, uint32_t negmax2) {
    r = __vmins2(__vmaxs2(r, negmax2), maxv2);
    return __viaddmin_s16x2_relu(p, r, maxv2);
}

__device__ __forceinline__ uint32_t prmt_r(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// grid = (vector chunks, 3 planes, pictures); a vector = 8 consecutive samples of a row.  A
// thread handles kReconVecs vectors one CTA width apart (every warp access stays a run of
// whole 128-byte lines) and issues all of its loads before the first use: the kernel is
// latency-bound, so the bytes in flight per thread are what matters.
constexpr int kReconThreads = 256;
constexpr int kReconVecs = 4;

template <typename T>
__global__ void __launch_bounds__(kReconThreads) recon_kernel(const __grid_constant__ ReconArgs a) {
    const int c = blockIdx.y, pic = blockIdx.z;
    const int w = c ? a.width >> 1 : a.width, h = c ? a.height >> 1 : a.height;
    const int stride = c ? a.stride_c : a.stride_y;
    const int bd = c ? a.bit_depth_c : a.bit_depth_y;
    const uint32_t vec_per_row = (uint32_t)(w + 7) >> 3;
    const uint32_t total = vec_per_row * (uint32_t)h;
    const uint32_t first = blockIdx.x * (uint32_t)(kReconThreads * kReconVecs) + threadIdx.x;
    if (first >= total) return;
    const int64_t base = (int64_t)pic * a.pic_stride + a.plane_off[c];
    const uint32_t maxv2 = ((1u << bd) - 1u) * 0x00010001u;
    const uint32_t negmax2 = (0x10000u - ((1u << bd) - 1u)) * 0x00010001u;  // -maxVal in both halves
    int64_t off[kReconVecs];
    bool ok[kReconVecs];
    uint4 r[kReconVecs], v[kReconVecs];
#pragma unroll
    for (int k = 0; k < kReconVecs; k++) {
        const uint32_t idx = first + (uint32_t)(k * kReconThreads);
        ok[k] = idx < total;
        const uint32_t y = idx / vec_per_row, xv = idx - y * vec_per_row;
        off[k] = base + (int64_t)y * stride + xv * 8;
    }
#pragma unroll
    for (int k = 0; k < kReconVecs; k++) {
        if (!ok[k]) continue;
        r[k] = *reinterpret_cast<const uint4 *>(a.res + off[k]);
        if (sizeof(T) == 2) {
            v[k] = *reinterpret_cast<const uint4 *>(reinterpret_cast<const uint16_t *>(a.pred) + off[k]);
        } else {
            const uint2 t = *reinterpret_cast<const uint2 *>(reinterpret_cast<const uint8_t *>(a.pred) + off[k]);
            v[k] = make_uint4(prmt_r(t.x, 0, 0x4140), prmt_r(t.x, 0, 0x4342), prmt_r(t.y, 0, 0x4140),
                              prmt_r(t.y, 0, 0x4342));
        }
    }
#pragma unroll
    for (int k = 0; k < kReconVecs; k++) {
        if (!ok[k]) continue;
        const uint32_t o0 = recon_word(v[k].x, r[k].x, maxv2, negmax2), o1 = recon_word(v[k].y, r[k].y, maxv2, negmax2);
        const uint32_t o2 = recon_word(v[k].z, r[k].z, maxv2, negmax2), o3 = recon_word(v[k].w, r[k].w, maxv2, negmax2);
        if (sizeof(T) == 2)
            *reinterpret_cast<uint4 *>(reinterpret_cast<uint16_t *>(a.rec) + off[k]) = make_uint4(o0, o1, o2, o3);
        else
            *reinterpret_cast<uint2 *>(reinterpret_cast<uint8_t *>(a.rec) + off[k]) =
                make_uint2(prmt_r(o0, o1, 0x6420), prmt_r(o2, o3, 0x6420));
    }
}

int launch_recon(p265_ctx *ctx, const void *d_pred, const int16_t *d_res, void *d_rec, const p265_pic_geom *g) {
    ReconArgs a;
    a.pred = d_pred;
    a.res = d_res;
    a.rec = d_rec;
    for (int c = 0; c < 3; c++) a.plane_off[c] = g->plane_off[c];
    a.pic_stride = g->pic_stride;
    a.width = g->width;
    a.height = g->height;
    a.stride_y = g->stride_y;
    a.stride_c = g->stride_c;
    a.bit_depth_y = g->bit_depth_y;
    a.bit_depth_c = g->bit_depth_c;
    if (g->n_pics > 65535) return set_error(P265_EINVAL, "too many pictures in one reconstruction batch");
    const int64_t vecs = (int64_t)((g->width + 7) >> 3) * g->height;  // luma plane is the largest
    if (vecs >= (int64_t)1 << 31) return set_error(P265_EINVAL, "picture too large for the reconstruction kernel");
    const int per_cta = kReconThreads * kReconVecs;
    const dim3 grid((unsigned)((vecs + per_cta - 1) / per_cta), 3, g->n_pics);
    if (g->bit_depth_y > 8 || g->bit_depth_c > 8) recon_kernel<uint16_t><<<grid, kReconThreads, 0, ctx->stream>>>(a);
    else recon_kernel<uint8_t><<<grid, kReconThreads, 0, ctx->stream>>>(a);
    P265_CUDA(cudaGetLastError());
    ctx->launches++;
    return P265_OK;
}

}  // namespace p265
