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
#ifndef P265_CTAS_PER_SM
#define P265_CTAS_PER_SM 4
#endif
constexpr int kCtasPerSm = P265_CTAS_PER_SM;  // 4: 128 regs/thread, 5: 96 regs (small spills)
constexpr int kCtaSmemBytes = kWarpsPerCta * 2 * kWarpSmemBytes;  // two tiles per warp

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int KEEP>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(KEEP) : "memory");
}

__device__ __forceinline__ int bin_of(const KernelArgs &a, int item) {
    return (item >= a.first_item[1]) + (item >= a.first_item[2]) + (item >= a.first_item[3]);
}

// descriptor of the TB this lane owns in work item `item`
__device__ __forceinline__ uint4 fetch_desc(const KernelArgs &a, int item, int lane, bool &valid) {
    const int bin = bin_of(a, item);
    int tb;
    switch (bin) {
        case 0: tb = lane_tb<5>(a, item - a.first_item[0], lane, valid); break;
        case 1: tb = lane_tb<4>(a, item - a.first_item[1], lane, valid); break;
        case 2: tb = lane_tb<3>(a, item - a.first_item[2], lane, valid); break;
        default: tb = lane_tb<2>(a, item - a.first_item[3], lane, valid); break;
    }
    return load_desc(a, tb, valid);
}

__device__ __forceinline__ void issue_tile(const KernelArgs &a, int item, int lane, const TbParams &t,
                                           unsigned char *buf) {
    switch (bin_of(a, item)) {
        case 0: tile_issue<5>(lane, t, buf); break;
        case 1: tile_issue<4>(lane, t, buf); break;
        case 2: tile_issue<3>(lane, t, buf); break;
        default: tile_issue<2>(lane, t, buf); break;
    }
}

// Out of line on purpose: each size path gets its own register allocation instead of
// inflating the pipeline loop (one call per ~4000-instruction work item).
template <int LOG2N, int SF>
__device__ __noinline__ void compute_item(int lane, const TbParams &t, unsigned char *buf) {
    const bool slow = __any_sync(0xffffffffu, t.lsh != 0);
    const bool special = __any_sync(0xffffffffu, t.valid && (t.flags & (P265_TU_SKIP | P265_TU_BYPASS)) != 0);
    if (special) phase_special<LOG2N>(lane, t, buf);
    int p[2][(1 << LOG2N) / 2];
    if (slow) phase_gather<LOG2N, SF, true>(lane, t, buf, p);  // rare
    else phase_gather<LOG2N, SF, false>(lane, t, buf, p);
    __syncwarp();  // every lane has read its columns: the tile may be overwritten by g
    phase_stage1<LOG2N>(lane, t, buf, p);
    __syncwarp();
    phase_stage2<LOG2N>(lane, t, buf);
}

// Persistent warps: warp w handles work items w, w + W, w + 2W, ... of the size-sorted
// list (all warps therefore sit in the same size bin at any time -> one code path hot in
// the instruction cache).  Three-deep software pipeline per warp:
//   descriptor of item k+2 (LDG)  |  tile of item k+1 (cp.async into the other buffer)
//   |  transform of item k
// so both dependent global-memory latencies of an item are hidden behind arithmetic.
template <int SF>
__global__ void __launch_bounds__(kWarpsPerCta * 32, kCtasPerSm) residual_kernel(const __grid_constant__ KernelArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stride = gridDim.x * kWarpsPerCta, n_items = a.first_item[4];
    int it = blockIdx.x * kWarpsPerCta + warp;
    if (it >= n_items) return;
    unsigned char *buf = smem + warp * (2 * kWarpSmemBytes);
    int b = 0;

    bool v_cur, v_next = false;
    const uint4 d_cur = fetch_desc(a, it, lane, v_cur);
    TbParams t_cur = make_params(a, d_cur, v_cur);
    issue_tile(a, it, lane, t_cur, buf);
    cp_async_commit();
    int it_next = it + stride;
    uint4 d_next = make_uint4(0, 0, 0, 0);
    if (it_next < n_items) d_next = fetch_desc(a, it_next, lane, v_next);

    while (true) {
        const bool has_next = it_next < n_items;
        TbParams t_next = t_cur;
        if (has_next) {
            t_next = make_params(a, d_next, v_next);
            issue_tile(a, it_next, lane, t_next, buf + (b ^ 1) * kWarpSmemBytes);
        }
        cp_async_commit();
        const int it_nn = it_next + stride;
        bool v_nn = false;
        uint4 d_nn = make_uint4(0, 0, 0, 0);
        if (it_nn < n_items) d_nn = fetch_desc(a, it_nn, lane, v_nn);

        cp_async_wait<1>();  // the current tile has landed (the newest group may be in flight)
        __syncwarp();
        unsigned char *cur = buf + b * kWarpSmemBytes;
        switch (bin_of(a, it)) {
            case 0: compute_item<5, SF>(lane, t_cur, cur); break;
            case 1: compute_item<4, SF>(lane, t_cur, cur); break;
            case 2: compute_item<3, SF>(lane, t_cur, cur); break;
            default: compute_item<2, SF>(lane, t_cur, cur); break;
        }
        __syncwarp();
        if (!has_next) break;
        it = it_next;
        t_cur = t_next;
        it_next = it_nn;
        d_next = d_nn;
        v_next = v_nn;
        b ^= 1;
    }
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
    a.sf_replicated = 0;
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

template <int SF>
static int launch_residual_sf(p265_ctx *ctx, const KernelArgs &a) {
    static int occ = 0;  // CTAs per SM the kernel really gets (same for every device of a box)
    if (!occ) {
        P265_CUDA(cudaFuncSetAttribute(residual_kernel<SF>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       cudaSharedmemCarveoutMaxShared));
        P265_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, residual_kernel<SF>, kWarpsPerCta * 32,
                                                                kCtaSmemBytes));
        if (occ < 1) occ = 1;
    }
    const int items = a.first_item[4];
    int grid = (items + kWarpsPerCta - 1) / kWarpsPerCta;
    const int persistent = ctx->sm_count * occ;
    if (grid > persistent) grid = persistent;
    residual_kernel<SF><<<grid, kWarpsPerCta * 32, kCtaSmemBytes, ctx->stream>>>(a);
    P265_CUDA(cudaGetLastError());
    ctx->launches++;
    return P265_OK;
}

int launch_residual(p265_ctx *ctx, const p265_tu_desc *d_tus, const int32_t bin_counts[4], const int16_t *d_coeffs,
                    const uint8_t *d_sf, const p265_pic_geom *g, int16_t *d_out, int flags) {
    KernelArgs a;
    int rc = fill_args(a, d_tus, bin_counts, d_coeffs, d_sf, g, d_out);
    if (rc) return rc;
    a.sf_replicated = (flags & P265_RES_SF_REPLICATED) != 0;
    if (flags & P265_RES_ZERO_FILL)
        P265_CUDA(cudaMemsetAsync(d_out, 0, sizeof(int16_t) * (size_t)g->pic_stride * g->n_pics, ctx->stream));
    if (a.first_item[4] == 0) return P265_OK;
    if (!a.sf) return launch_residual_sf<SF_NONE>(ctx, a);
    if (a.sf_replicated) return launch_residual_sf<SF_REPLICATED>(ctx, a);
    return launch_residual_sf<SF_GENERAL>(ctx, a);
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
