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

// grid = (row chunks, 3 planes, pictures); a thread handles 8 consecutive samples of a row
template <typename T>
__global__ void __launch_bounds__(256) recon_kernel(const __grid_constant__ ReconArgs a) {
    const int c = blockIdx.y, pic = blockIdx.z;
    const int w = c ? a.width >> 1 : a.width, h = c ? a.height >> 1 : a.height;
    const int stride = c ? a.stride_c : a.stride_y;
    const int bd = c ? a.bit_depth_c : a.bit_depth_y;
    const int vec_per_row = (w + 7) >> 3;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)vec_per_row * h) return;
    const int y = (int)(idx / vec_per_row), xv = (int)(idx - (int64_t)y * vec_per_row);
    const int64_t off = (int64_t)pic * a.pic_stride + a.plane_off[c] + (int64_t)y * stride + xv * 8;
    const uint32_t maxv2 = ((1u << bd) - 1u) * 0x00010001u;
    const uint32_t negmax2 = (0x10000u - ((1u << bd) - 1u)) * 0x00010001u;  // -maxVal in both halves
    const uint4 r = *reinterpret_cast<const uint4 *>(a.res + off);
    uint32_t p[4];
    if (sizeof(T) == 2) {
        const uint4 v = *reinterpret_cast<const uint4 *>(reinterpret_cast<const uint16_t *>(a.pred) + off);
        p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w;
    } else {
        const uint2 v = *reinterpret_cast<const uint2 *>(reinterpret_cast<const uint8_t *>(a.pred) + off);
        p[0] = prmt_r(v.x, 0, 0x4140); p[1] = prmt_r(v.x, 0, 0x4342);
        p[2] = prmt_r(v.y, 0, 0x4140); p[3] = prmt_r(v.y, 0, 0x4342);
    }
    const uint32_t o0 = recon_word(p[0], r.x, maxv2, negmax2), o1 = recon_word(p[1], r.y, maxv2, negmax2);
    const uint32_t o2 = recon_word(p[2], r.z, maxv2, negmax2), o3 = recon_word(p[3], r.w, maxv2, negmax2);
    if (sizeof(T) == 2)
        *reinterpret_cast<uint4 *>(reinterpret_cast<uint16_t *>(a.rec) + off) = make_uint4(o0, o1, o2, o3);
    else
        *reinterpret_cast<uint2 *>(reinterpret_cast<uint8_t *>(a.rec) + off) =
            make_uint2(prmt_r(o0, o1, 0x6420), prmt_r(o2, o3, 0x6420));
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
    const dim3 grid((unsigned)((vecs + 255) / 256), 3, g->n_pics);
    if (g->bit_depth_y > 8 || g->bit_depth_c > 8) recon_kernel<uint16_t><<<grid, 256, 0, ctx->stream>>>(a);
    else recon_kernel<uint8_t><<<grid, 256, 0, ctx->stream>>>(a);
    P265_CUDA(cudaGetLastError());
    ctx->launches++;
    return P265_OK;
}

}  // namespace p265
