// sm_100a kernels of the residual path: one batched launch, binned by TB size, that
// dequantises (scaling.py:4-47) and inverse-transforms (transform.py:89-109, as the
// standard specifies it) every coded TB of a batch of pictures into int16 residual
// planes.  The arithmetic lives in residual_core.cuh (shared with the host emulation).
//
// Grid: one warp per work item (64 TB columns: 2/4/8/16 TBs of 32/16/8/4), 4 warps per
// CTA.  Warps never talk to each other: every hand-over (global -> smem tile, stage 1
// -> stage 2 transpose) is warp-private shared memory fenced by __syncwarp(), so there
// is no __syncthreads() in the kernel and CTAs are only a scheduling container.
#include <cuda_runtime.h>

#include "internal.h"
#include "residual_core.cuh"

namespace p265 {

constexpr int kWarpsPerCta = 4;

template <int LOG2N, bool HAS_SF>
__device__ __forceinline__ void run_item(const KernelArgs &a, int item, int lane, unsigned char *wsmem) {
    bool valid;
    const int tb = lane_tb<LOG2N>(a, item, lane, valid);
    const TbParams t = make_params(a, tb, valid);
    phase_load<LOG2N>(lane, t, wsmem);
    const bool slow = __any_sync(0xffffffffu, t.lsh != 0);  // also fences the tile (bar.warp.sync semantics)
    __syncwarp();
    int p[2][(1 << LOG2N) / 2];
    if (slow) phase_gather<LOG2N, HAS_SF, true>(lane, t, wsmem, p);  // rare
    else phase_gather<LOG2N, HAS_SF, false>(lane, t, wsmem, p);
    __syncwarp();  // every lane has read its columns: the tile may be overwritten by g
    phase_stage1<LOG2N>(lane, t, wsmem, p);
    __syncwarp();
    phase_stage2<LOG2N>(lane, t, wsmem);
}

template <bool HAS_SF>
__global__ void __launch_bounds__(kWarpsPerCta * 32) residual_kernel(const __grid_constant__ KernelArgs a) {
    __shared__ __align__(128) unsigned char smem[kWarpsPerCta * kWarpSmemBytes];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = blockIdx.x * kWarpsPerCta + warp;
    if (w >= a.first_item[4]) return;
    unsigned char *wsmem = smem + warp * kWarpSmemBytes;
    if (w < a.first_item[1]) run_item<5, HAS_SF>(a, w - a.first_item[0], lane, wsmem);
    else if (w < a.first_item[2]) run_item<4, HAS_SF>(a, w - a.first_item[1], lane, wsmem);
    else if (w < a.first_item[3]) run_item<3, HAS_SF>(a, w - a.first_item[2], lane, wsmem);
    else run_item<2, HAS_SF>(a, w - a.first_item[3], lane, wsmem);
}

// ---- auxiliary, non-hot kernels ------------------------------------------------------
// scaling.inverse_scaling alone: d[] in arena layout (what pu.scaled_samples receives).
__global__ void dequant_kernel(const __grid_constant__ KernelArgs a, int n_tus, int16_t *scaled) {
    const int lane = threadIdx.x & 31;
    const int tb = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tb >= n_tus) return;
    const TbParams t = make_params(a, tb, true);
    const int log2n = a.tus[tb].log2n;
    const int n2 = 1 << (2 * log2n);
    int16_t *dst = scaled + (t.src - a.coeffs);
    for (int e = lane; e < n2; e += 32) {
        const int m = t.sf ? (int)t.sf[e] * t.w : t.w;
        int d = dequant((int)t.src[e], m, t);
        d = max(-32768, min(32767, d));
        dst[e] = (int16_t)d;
    }
}

__constant__ Basis g_basis = Basis();

__device__ __forceinline__ int kDstDev(int i, int j) {
    const int v[16] = {29, 55, 74, 84, 74, 74, 0, -74, 84, -29, -74, 55, 55, -84, 74, -29};
    return v[i * 4 + j];
}

// transform.py:89-109 exactly as written: parity-test-only.
//   C[i][j] = DST[i][j] (4x4 luma) or DCT32[i][j * 32 / N]      (transform.py:79-85)
//   e[:, col] = C . d_xy[:, col]; g = clip16((e + 64) >> 7)       (transform.py:100-106)
//   r[row, :] = C . g[:, N-1] for every row                       (transform.py:108-109)
// in: d[] arena ([y][x] per TB); out: int32 arena, [x][y] per TB like the reference.
__global__ void ref_literal_kernel(const p265_tu_desc *tus, int n_tus, const int16_t *scaled, int32_t *out) {
    const int lane = threadIdx.x & 31;
    const int tb = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tb >= n_tus) return;
    const p265_tu_desc t = tus[tb];
    const int n = 1 << t.log2n;
    const bool dst = (n == 4 && t.c_idx == 0);
    const int step = 32 >> t.log2n;
    const size_t off = (size_t)t.coeff_off * 16;
    const int16_t *d = scaled + off;
    const int i = lane & (n - 1);
    int s = 0;
    for (int j = 0; j < n; j++) {
        const int c = dst ? kDstDev(i, j) : (int)g_basis.m[i][j * step];
        s += c * (int)d[(n - 1) * n + j];  // d_xy[j][n-1] == d_yx[n-1][j]
    }
    const int gl = max(-32768, min(32767, (s + 64) >> 7));
    int r = 0;
    for (int j = 0; j < n; j++) {
        const int c = dst ? kDstDev(i, j) : (int)g_basis.m[i][j * step];
        r += c * __shfl_sync(0xffffffffu, gl, j);
    }
    if (lane < n)
        for (int x = 0; x < n; x++) out[off + (size_t)x * n + lane] = r;
}

// transform.inverse_transform_1d on one vector (helper of the reference surface, not hot)
__global__ void idct1d_kernel(const int32_t *x, int log2size, int tr_type, int mode, int32_t *y) {
    const int n = 1 << log2size, i = threadIdx.x, step = 32 >> log2size;
    if (i >= n) return;
    int s = 0;
    for (int j = 0; j < n; j++) {
        int c;
        if (tr_type == 1) c = mode ? kDstDev(i, j) : kDstDev(j, i);
        else c = mode ? (int)g_basis.m[i][j * step] : (int)g_basis.m[j * step][i];
        s += c * x[j];
    }
    y[i] = s;
}

int launch_idct1d(p265_ctx *ctx, const int32_t *d_x, int log2size, int tr_type, int mode, int32_t *d_y) {
    idct1d_kernel<<<1, 32, 0, ctx->stream>>>(d_x, log2size, tr_type, mode, d_y);
    P265_CUDA(cudaGetLastError());
    ctx->launches++;
    return P265_OK;
}

// ---- launchers -----------------------------------------------------------------------
static int fill_args(KernelArgs &a, const p265_tu_desc *d_tus, const int32_t bin_counts[4], const int16_t *d_coeffs,
                     const uint8_t *d_sf, const p265_pic_geom *g, int16_t *d_out) {
    a.tus = d_tus;
    a.coeffs = d_coeffs;
    a.sf = d_sf;
    a.out = d_out;
    for (int c = 0; c < 3; c++) a.plane_off[c] = g->plane_off[c];
    a.pic_stride = g->pic_stride;
    a.stride_y = g->stride_y;
    a.stride_c = g->stride_c;
    a.bit_depth_y = g->bit_depth_y;
    a.bit_depth_c = g->bit_depth_c;
    int64_t first = 0, items = 0;
    for (int b = 0; b < 4; b++) {
        a.first_tb[b] = (int32_t)first;
        a.n_tb[b] = bin_counts[b];
        first += bin_counts[b];
        a.first_item[b] = (int32_t)items;
        const int per = 2 << b;
        items += (bin_counts[b] + per - 1) / per;
    }
    if (first > INT32_MAX || items > INT32_MAX) return set_error(P265_EINVAL, "too many TBs in one batch");
    a.first_item[4] = (int32_t)items;
    return P265_OK;
}

int launch_residual(p265_ctx *ctx, const p265_tu_desc *d_tus, const int32_t bin_counts[4], const int16_t *d_coeffs,
                    const uint8_t *d_sf, const p265_pic_geom *g, int16_t *d_out, int flags) {
    KernelArgs a;
    int rc = fill_args(a, d_tus, bin_counts, d_coeffs, d_sf, g, d_out);
    if (rc) return rc;
    if (flags & P265_RES_ZERO_FILL)
        P265_CUDA(cudaMemsetAsync(d_out, 0, sizeof(int16_t) * (size_t)g->pic_stride * g->n_pics, ctx->stream));
    const int items = a.first_item[4];
    if (items == 0) return P265_OK;
    const int grid = (items + kWarpsPerCta - 1) / kWarpsPerCta;
    if (a.sf) residual_kernel<true><<<grid, kWarpsPerCta * 32, 0, ctx->stream>>>(a);
    else residual_kernel<false><<<grid, kWarpsPerCta * 32, 0, ctx->stream>>>(a);
    P265_CUDA(cudaGetLastError());
    ctx->launches++;
    return P265_OK;
}

int launch_dequant(p265_ctx *ctx, const p265_tu_desc *d_tus, int n_tus, const int16_t *d_coeffs, const uint8_t *d_sf,
                   int bit_depth_y, int bit_depth_c, int16_t *d_scaled) {
    if (n_tus == 0) return P265_OK;
    KernelArgs a = {};
    a.tus = d_tus;
    a.coeffs = d_coeffs;
    a.sf = d_sf;
    a.out = nullptr;
    a.bit_depth_y = bit_depth_y;
    a.bit_depth_c = bit_depth_c;
    dequant_kernel<<<(n_tus + 3) / 4, 128, 0, ctx->stream>>>(a, n_tus, d_scaled);
    P265_CUDA(cudaGetLastError());
    ctx->launches++;
    return P265_OK;
}

int launch_ref_literal(p265_ctx *ctx, const p265_tu_desc *d_tus, int n_tus, const int16_t *d_scaled, int32_t *d_out) {
    if (n_tus == 0) return P265_OK;
    ref_literal_kernel<<<(n_tus + 3) / 4, 128, 0, ctx->stream>>>(d_tus, n_tus, d_scaled, d_out);
    P265_CUDA(cudaGetLastError());
    ctx->launches++;
    return P265_OK;
}

}  // namespace p265
