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

// Occupancy plan (per SM): 11 CTAs x 2 warps = 22 warps, <= 93 registers/thread,
// 11 x 18.5 KB = 204 KB of shared memory (tile + g + descriptor ring per warp).
#ifndef P265_WARPS_PER_CTA
#define P265_WARPS_PER_CTA 2
#endif
#ifndef P265_CTAS_PER_SM
#define P265_CTAS_PER_SM 11
#endif
constexpr int kWarpsPerCta = P265_WARPS_PER_CTA;
constexpr int kCtasPerSm = P265_CTAS_PER_SM;
constexpr int kDescRingBytes = 2 * 32 * 16;                          // 2 slots x 32 lanes x 16 B
constexpr int kWarpBytes = 2 * kWarpSmemBytes + kDescRingBytes;      // in + g + ring = 9472
constexpr int kCtaSmemBytes = kWarpsPerCta * kWarpBytes;

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int KEEP>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(KEEP) : "memory");
}

// Out-of-line wrappers (see residual_core.cuh): one copy of each 1-D pass per size, shared
// by the two columns / rows a lane owns.  Inside a loop ptxas would hoist the ~90 packed
// basis constants of the 32-point butterfly and need > 180 registers; as functions each
// pass keeps a small allocation and fetches its constants just in time (LDCU).
template <int LOG2N, int SF, bool SLOW>
__device__ __noinline__ void stage1_call(const unsigned char *in, unsigned char *g, int x, int tl, int half,
                                         const uint8_t *sf, int w, int rnd, int sh, int lsh, int dst_flag) {
    stage1_column<LOG2N, SF, SLOW>(in, g, x, tl, half, sf, w, rnd, sh, lsh, dst_flag);
}
template <int LOG2N>
__device__ __noinline__ void stage2_call(const unsigned char *g, int row, int16_t *dst, int rnd2, int sh2,
                                         int dst_flag) {
    stage2_row<LOG2N>(g, row, dst, rnd2, sh2, dst_flag);
}

// One size bin, one warp: items w, w + W, w + 2W, ... of the bin.  Software pipeline,
// everything asynchronous (cp.async / LDGSTS, no register staging):
//     descriptor of item k+2  ->  lane-private slot of a 2-entry ring in shared memory
//     tile of item k+1        ->  `in`, issued as soon as stage 1 of item k has consumed it
//     stage 2 of item k       ->  overlaps the tile copy
// so both dependent global-memory latencies of an item hide behind arithmetic.
template <int LOG2N, int SF>
__device__ __forceinline__ void run_bin(const KernelArgs &a, int gw, int stride, int lane, unsigned char *wbase) {
    using L = Layout<LOG2N>;
    constexpr int N = L::N, bin = 5 - LOG2N;
    const int n_items = a.first_item[bin + 1] - a.first_item[bin];
    if (gw >= n_items) return;
    unsigned char *in_base = wbase, *g_base = wbase + kWarpSmemBytes;
    uint4 *ring = reinterpret_cast<uint4 *>(wbase + 2 * kWarpSmemBytes) + lane;  // slot s at ring[32 * s]
    const int tb_l = lane / L::TPB, tl = lane % L::TPB;
    const unsigned char *in = in_base + tb_l * L::TB_BYTES;
    unsigned char *g = g_base + tb_l * L::TB_BYTES;
    const int x0 = slot_index_rt(N, tl, 0), x1 = slot_index_rt(N, tl, 1);
    bool valid;
    {   // prologue: first descriptor by plain load, its tile and the second descriptor async
        const int tb = lane_tb<LOG2N>(a, gw, lane, valid);
        const uint4 d0 = load_desc(a, tb, valid);
        ring[0] = d0;
        tile_issue<LOG2N>(lane, a.coeffs + (size_t)d0.z * 16, valid, in_base);
        if (gw + stride < n_items) {
            bool v1;
            const int tb1 = lane_tb<LOG2N>(a, gw + stride, lane, v1);
            if (v1) copy16_async(&ring[32], &a.tus[tb1]);
        }
        cp_async_commit();
    }
    int k = 0;
    for (int it = gw; it < n_items; it += stride, k ^= 1) {
        cp_async_wait<0>();  // tile k and descriptor k+1 have landed
        __syncwarp();        // ... for every lane; also: all lanes are done with g of item k-1
        lane_tb<LOG2N>(a, it, lane, valid);
        const TbParams t = make_params(a, ring[32 * k], valid);
        const bool is_special = (t.flags & (P265_TU_SKIP | P265_TU_BYPASS)) != 0;
        const bool slow = __any_sync(0xffffffffu, t.lsh != 0);
        if (__any_sync(0xffffffffu, t.valid && is_special)) phase_special<LOG2N>(lane, t, in_base);
        const int dstf = t.flags & P265_TU_DST;
        if (!slow) {
            stage1_call<LOG2N, SF, false>(in, g, x0, tl, 0, t.sf, t.w, t.rnd, t.sh, 0, dstf);
            stage1_call<LOG2N, SF, false>(in, g, x1, tl, 1, t.sf, t.w, t.rnd, t.sh, 0, dstf);
        } else {  // rare
            stage1_call<LOG2N, SF, true>(in, g, x0, tl, 0, t.sf, t.w, t.rnd, t.sh, t.lsh, dstf);
            stage1_call<LOG2N, SF, true>(in, g, x1, tl, 1, t.sf, t.w, t.rnd, t.sh, t.lsh, dstf);
        }
        __syncwarp();  // `in` is consumed, g is complete
        if (it + stride < n_items) {
            bool v1;
            lane_tb<LOG2N>(a, it + stride, lane, v1);
            const uint4 dn = ring[32 * (k ^ 1)];
            tile_issue<LOG2N>(lane, a.coeffs + (size_t)dn.z * 16, v1, in_base);
            if (it + 2 * stride < n_items) {  // slot k is free: its descriptor sits in `t`
                bool v2;
                const int tb2 = lane_tb<LOG2N>(a, it + 2 * stride, lane, v2);
                if (v2) copy16_async(&ring[32 * k], &a.tus[tb2]);
            }
        }
        cp_async_commit();
        if (t.valid && !is_special) {
            stage2_call<LOG2N>(g, tl, t.dst + (size_t)tl * t.stride, t.rnd2, t.sh2, dstf);
            stage2_call<LOG2N>(g, tl + L::TPB, t.dst + (size_t)(tl + L::TPB) * t.stride, t.rnd2, t.sh2, dstf);
        }
    }
    cp_async_wait<0>();
    __syncwarp();
}

// Persistent warps walk the size-sorted work-item list bin by bin (all warps therefore
// run the same size path at any time -> one code path hot in the instruction cache).
template <int SF>
__global__ void __launch_bounds__(kWarpsPerCta * 32, kCtasPerSm) residual_kernel(const __grid_constant__ KernelArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stride = gridDim.x * kWarpsPerCta;
    const int gw = blockIdx.x * kWarpsPerCta + warp;
    unsigned char *wbase = smem + warp * kWarpBytes;
    run_bin<5, SF>(a, gw, stride, lane, wbase);
    run_bin<4, SF>(a, gw, stride, lane, wbase);
    run_bin<3, SF>(a, gw, stride, lane, wbase);
    run_bin<2, SF>(a, gw, stride, lane, wbase);
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
