// sm_100a deblocking kernel: H.265 8.7.2 over reconstructed pictures, in place, one pass.
// The reference has no deblocking filter (it only parses the control flags, pps.py:121-131,
// slice.py:170-179); parity is against the spec oracle, which is
// pinned by libavcodec's decode of sanity.bin.
//
// Decomposition.  The standard filters every vertical edge of the picture, then every
// horizontal edge on the result.  Edges lie on the 8-sample grid and a filtered edge reads
// and writes only the 4 samples on either side of it, so the picture splits into 8x8 blocks
// SHIFTED by (-4, -4): block (i, j) = [8i-4, 8i+4) x [8j-4, 8j+4) holds exactly the
// vertical-edge samples of x = 8i (two 4-row segments) and the horizontal-edge samples of
// y = 8j (two 4-column segments), and the horizontal filter's inputs are this block's own
// vertically filtered samples.  Shifted blocks are therefore independent: one lane loads
// one block (8 rows x two 8-byte halves; a warp covers 256 contiguous samples per row),
// filters both directions in registers and stores it back -- one read and one write per
// sample, no shared-memory exchange, no second pass over HBM.  Blocks whose four segments
// all have Bs = 0 are neither loaded nor stored.  Chroma planes use the same scheme on their
// own 8-sample grid (Bs = 2 only, one sample per side).
//
// Per segment the lane derives beta / tC from the edge map (QpY of the two CUs, Bs) and the
// CTB's slice offsets through two small tables staged in shared memory.
#include <cuda_runtime.h>

#include "internal.h"

namespace p265 {

constexpr int kDbkThreads = 128;

struct DbkArgs {
    void *pix;
    const p265_dbk_blk *blk;
    const p265_dbk_ctb *ctb;
    int64_t plane_off[3];
    int64_t pic_stride;
    int32_t width, height, stride_y, stride_c;
    int32_t bit_depth_y, bit_depth_c;
    int32_t w8, h8;              // luma 8x8 blocks per row / column
    int32_t ctb_shift;           // ctb_log2 - 3
    int32_t ctbs_w, ctbs_h;
    int32_t chunks_y, rows_y;    // luma: 32-block chunks per block row, block rows
};

__constant__ uint8_t c_beta[52] = {0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  6,  7,
                                   8,  9,  10, 11, 12, 13, 14, 15, 16, 17, 18, 20, 22, 24, 26, 28, 30, 32,
                                   34, 36, 38, 40, 42, 44, 46, 48, 50, 52, 54, 56, 58, 60, 62, 64};
__constant__ uint8_t c_tc[54] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1,  1,  1,  1,  1,  1,  1,  1,
                                 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 5, 5, 6, 6, 7, 8, 9, 10, 11, 13, 14, 16, 18, 20, 22, 24};

__device__ __forceinline__ int clip3i(int lo, int hi, int v) { return min(max(v, lo), hi); }
__device__ __forceinline__ int blk_qp(uint32_t e) { return ((int)(e << 17)) >> 25; }  // bits 8..14, sign-extended

// Table 8-10, ChromaArrayType == 1
__device__ __forceinline__ int chroma_qp(int qpi) {
    if (qpi < 30) return qpi;
    if (qpi >= 44) return qpi - 6;
    // 30..43 -> 29,30,31,32,33,33,34,34,35,35,36,36,37,37
    return qpi <= 34 ? qpi - 1 : 33 + ((qpi - 34) >> 1);
}

struct Seg {
    int bs;      // 0: nothing to do
    int beta, tc;
    bool no_p, no_q;
};

// 8.7.2.5.3 (beta, tC) for one segment; eq / ep = edge-map entries of the blocks holding q0 / p0
template <bool CHROMA>
__device__ __forceinline__ Seg make_seg(int bs, uint32_t eq, uint32_t ep, const p265_dbk_ctb par, int c, int bd,
                                        const uint8_t *s_beta, const uint8_t *s_tc) {
    Seg s;
    s.bs = CHROMA ? (bs == 2 ? 2 : 0) : bs;
    s.no_p = (ep & P265_DBK_NO_FILTER) != 0;
    s.no_q = (eq & P265_DBK_NO_FILTER) != 0;
    const int qpl = (blk_qp(eq) + blk_qp(ep) + 1) >> 1;
    if (CHROMA) {
        const int qpc = chroma_qp(qpl + (c == 1 ? par.cb_qp_offset : par.cr_qp_offset));
        s.tc = (int)s_tc[clip3i(0, 53, qpc + 2 + 2 * par.tc_offset_div2)] << (bd - 8);
        s.beta = 0;
    } else {
        s.beta = (int)s_beta[clip3i(0, 51, qpl + 2 * par.beta_offset_div2)] << (bd - 8);
        s.tc = (int)s_tc[clip3i(0, 53, qpl + 2 * (s.bs - 1) + 2 * par.tc_offset_div2)] << (bd - 8);
    }
    return s;
}

// Decisions of a luma segment (8.7.2.5.3): a = line 0, b = line 3, each p3..p0 q0..q3.
struct Dec {
    bool on, strong, dep, deq;
};
__device__ __forceinline__ bool dsam(int dpq2, const int (&l)[8], int beta, int tc) {
    return dpq2 < (beta >> 2) && abs(l[0] - l[3]) + abs(l[4] - l[7]) < (beta >> 3) &&
           abs(l[3] - l[4]) < ((5 * tc + 1) >> 1);
}
__device__ __forceinline__ Dec decide(const int (&a)[8], const int (&b)[8], const Seg &s) {
    // l[0..3] = p3 p2 p1 p0, l[4..7] = q0 q1 q2 q3
    const int dp0 = abs(a[1] - 2 * a[2] + a[3]), dp3 = abs(b[1] - 2 * b[2] + b[3]);
    const int dq0 = abs(a[6] - 2 * a[5] + a[4]), dq3 = abs(b[6] - 2 * b[5] + b[4]);
    const int dpq0 = dp0 + dq0, dpq3 = dp3 + dq3;
    Dec d;
    d.on = s.bs > 0 && dpq0 + dpq3 < s.beta;
    d.strong = dsam(2 * dpq0, a, s.beta, s.tc) && dsam(2 * dpq3, b, s.beta, s.tc);
    const int side = (s.beta + (s.beta >> 1)) >> 3;
    d.dep = dp0 + dp3 < side;
    d.deq = dq0 + dq3 < side;
    return d;
}

// 8.7.2.5.7 on one line l = p3 p2 p1 p0 q0 q1 q2 q3
__device__ __forceinline__ void luma_line(int (&l)[8], const Dec &d, const Seg &s, int maxv) {
    const int p3 = l[0], p2 = l[1], p1 = l[2], p0 = l[3], q0 = l[4], q1 = l[5], q2 = l[6], q3 = l[7];
    const int tc = s.tc;
    int n[8] = {p3, p2, p1, p0, q0, q1, q2, q3};
    if (d.strong) {
        const int t2 = 2 * tc;
        n[3] = clip3i(p0 - t2, p0 + t2, (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
        n[2] = clip3i(p1 - t2, p1 + t2, (p2 + p1 + p0 + q0 + 2) >> 2);
        n[1] = clip3i(p2 - t2, p2 + t2, (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
        n[4] = clip3i(q0 - t2, q0 + t2, (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
        n[5] = clip3i(q1 - t2, q1 + t2, (p0 + q0 + q1 + q2 + 2) >> 2);
        n[6] = clip3i(q2 - t2, q2 + t2, (p0 + q0 + q1 + 3 * q2 + 2 * q3 + 4) >> 3);
    } else {
        int delta = (9 * (q0 - p0) - 3 * (q1 - p1) + 8) >> 4;
        if (abs(delta) < tc * 10) {
            delta = clip3i(-tc, tc, delta);
            n[3] = clip3i(0, maxv, p0 + delta);
            n[4] = clip3i(0, maxv, q0 - delta);
            const int th = tc >> 1;
            if (d.dep) n[2] = clip3i(0, maxv, p1 + clip3i(-th, th, (((p2 + p0 + 1) >> 1) - p1 + delta) >> 1));
            if (d.deq) n[5] = clip3i(0, maxv, q1 + clip3i(-th, th, (((q2 + q0 + 1) >> 1) - q1 - delta) >> 1));
        }
    }
    if (!s.no_p) { l[1] = n[1]; l[2] = n[2]; l[3] = n[3]; }
    if (!s.no_q) { l[4] = n[4]; l[5] = n[5]; l[6] = n[6]; }
}

// 8.7.2.5.8 on one line (only p1 p0 q0 q1 = l[2..5] are used)
__device__ __forceinline__ void chroma_line(int (&l)[8], const Seg &s, int maxv) {
    const int delta = clip3i(-s.tc, s.tc, ((((l[4] - l[3]) << 2) + l[2] - l[5] + 4) >> 3));
    if (!s.no_p) l[3] = clip3i(0, maxv, l[3] + delta);
    if (!s.no_q) l[4] = clip3i(0, maxv, l[4] - delta);
}

template <typename T>
__device__ __forceinline__ void load4(const T *p, int (&v)[8], int at);
template <>
__device__ __forceinline__ void load4<uint16_t>(const uint16_t *p, int (&v)[8], int at) {
    const uint2 w = *reinterpret_cast<const uint2 *>(p);
    v[at] = w.x & 0xffff; v[at + 1] = w.x >> 16; v[at + 2] = w.y & 0xffff; v[at + 3] = w.y >> 16;
}
template <>
__device__ __forceinline__ void load4<uint8_t>(const uint8_t *p, int (&v)[8], int at) {
    const uint32_t w = *reinterpret_cast<const uint32_t *>(p);
    v[at] = w & 0xff; v[at + 1] = (w >> 8) & 0xff; v[at + 2] = (w >> 16) & 0xff; v[at + 3] = w >> 24;
}
template <typename T>
__device__ __forceinline__ void store4(T *p, const int (&v)[8], int at);
template <>
__device__ __forceinline__ void store4<uint16_t>(uint16_t *p, const int (&v)[8], int at) {
    *reinterpret_cast<uint2 *>(p) = make_uint2((uint32_t)v[at] | ((uint32_t)v[at + 1] << 16),
                                               (uint32_t)v[at + 2] | ((uint32_t)v[at + 3] << 16));
}
template <>
__device__ __forceinline__ void store4<uint8_t>(uint8_t *p, const int (&v)[8], int at) {
    *reinterpret_cast<uint32_t *>(p) = (uint32_t)v[at] | ((uint32_t)v[at + 1] << 8) | ((uint32_t)v[at + 2] << 16) |
                                       ((uint32_t)v[at + 3] << 24);
}

// grid = (ceil(items / warps per CTA), 3 components, pictures); chroma planes use the first
// part of the luma-sized item range.
template <typename T>
__global__ void __launch_bounds__(kDbkThreads) deblock_kernel(const __grid_constant__ DbkArgs a) {
    __shared__ uint8_t s_beta[52], s_tc[54];
    if (threadIdx.x < 52) s_beta[threadIdx.x] = c_beta[threadIdx.x];
    if (threadIdx.x >= 64 && threadIdx.x < 64 + 54) s_tc[threadIdx.x - 64] = c_tc[threadIdx.x - 64];
    __syncthreads();

    const int c = blockIdx.y, pic = blockIdx.z;
    const int cs = c ? 1 : 0;
    const int w = a.width >> cs, h = a.height >> cs;
    const int nbx = ((w + 3) >> 3) + 1, nby = ((h + 3) >> 3) + 1;
    const int chunks = (nbx + 31) >> 5;
    const int item = blockIdx.x * (kDbkThreads / 32) + (threadIdx.x >> 5);
    if (item >= chunks * nby) return;
    const int j = item / chunks;
    const int i = (item - j * chunks) * 32 + (threadIdx.x & 31);
    if (i >= nbx) return;

    // ---- edge map: entries of the four 8x8 luma blocks around the block's centre ----------
    const int I = i << cs, J = j << cs;
    const p265_dbk_blk *bp = a.blk + (size_t)pic * a.w8 * a.h8;
    const bool in_i = I < a.w8, in_i1 = I >= 1 && I - 1 < a.w8, in_j = J < a.h8, in_j1 = J >= 1 && J - 1 < a.h8;
    const uint32_t e11 = (in_i && in_j) ? bp[J * a.w8 + I] : 0u;
    const uint32_t e01 = (in_i1 && in_j) ? bp[J * a.w8 + I - 1] : 0u;
    const uint32_t e10 = (in_i && in_j1) ? bp[(J - 1) * a.w8 + I] : 0u;
    const uint32_t e00 = (in_i1 && in_j1) ? bp[(J - 1) * a.w8 + I - 1] : 0u;
    int bs_vu = (e10 >> (cs ? P265_DBK_BS_V0 : P265_DBK_BS_V1)) & 3, bs_vl = (e11 >> P265_DBK_BS_V0) & 3;
    int bs_hl = (e01 >> (cs ? P265_DBK_BS_H0 : P265_DBK_BS_H1)) & 3, bs_hr = (e11 >> P265_DBK_BS_H0) & 3;
    if (cs) {
        bs_vu &= 2; bs_vl &= 2; bs_hl &= 2; bs_hr &= 2;   // chroma: Bs == 2 only (1 never has bit 1)
    }
    if (i == 0) bs_vu = bs_vl = 0;   // x = 0 / y = 0 are picture boundaries, whatever the map says
    if (j == 0) bs_hl = bs_hr = 0;
    if ((bs_vu | bs_vl | bs_hl | bs_hr) == 0) return;

    const int bd = c ? a.bit_depth_c : a.bit_depth_y;
    const int maxv = (1 << bd) - 1;
    const int stride = c ? a.stride_c : a.stride_y;
    T *base = reinterpret_cast<T *>(a.pix) + (size_t)pic * a.pic_stride + a.plane_off[c];
    const int x0 = 8 * i - 4, y0 = 8 * j - 4;
    const bool has_l = i > 0, has_r = 8 * i < w;

    // ---- load the shifted block --------------------------------------------------------
    int v[8][8];
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const int y = y0 + r;
        const bool row_ok = y >= 0 && y < h;
#pragma unroll
        for (int k = 0; k < 8; k++) v[r][k] = 0;
        if (row_ok && has_l) load4<T>(base + (size_t)y * stride + x0, v[r], 0);
        if (row_ok && has_r) load4<T>(base + (size_t)y * stride + x0 + 4, v[r], 4);
    }

    // ---- per-segment parameters ----------------------------------------------------------
    const p265_dbk_ctb *cp = a.ctb + (size_t)pic * a.ctbs_w * a.ctbs_h;
    auto ctb_of = [&](int bi, int bj) {
        const int ci = min(max(bi, 0) >> a.ctb_shift, a.ctbs_w - 1), cj = min(max(bj, 0) >> a.ctb_shift, a.ctbs_h - 1);
        return cp[cj * a.ctbs_w + ci];
    };
    Seg sv[2], sh[2];
    if (cs) {
        sv[0] = make_seg<true>(bs_vu, e10, e00, ctb_of(I, J - 1), c, bd, s_beta, s_tc);
        sv[1] = make_seg<true>(bs_vl, e11, e01, ctb_of(I, J), c, bd, s_beta, s_tc);
        sh[0] = make_seg<true>(bs_hl, e01, e00, ctb_of(I - 1, J), c, bd, s_beta, s_tc);
        sh[1] = make_seg<true>(bs_hr, e11, e10, ctb_of(I, J), c, bd, s_beta, s_tc);
    } else {
        sv[0] = make_seg<false>(bs_vu, e10, e00, ctb_of(I, J - 1), c, bd, s_beta, s_tc);
        sv[1] = make_seg<false>(bs_vl, e11, e01, ctb_of(I, J), c, bd, s_beta, s_tc);
        sh[0] = make_seg<false>(bs_hl, e01, e00, ctb_of(I - 1, J), c, bd, s_beta, s_tc);
        sh[1] = make_seg<false>(bs_hr, e11, e10, ctb_of(I, J), c, bd, s_beta, s_tc);
    }

    // ---- vertical edge x = 8i: rows 0-3 and 4-7, across = columns -----------------------
#pragma unroll
    for (int sgm = 0; sgm < 2; sgm++) {
        const Seg s = sv[sgm];
        if (s.bs == 0) continue;
        if (cs) {
#pragma unroll
            for (int r = 0; r < 4; r++) chroma_line(v[sgm * 4 + r], s, maxv);
        } else {
            const Dec d = decide(v[sgm * 4], v[sgm * 4 + 3], s);
            if (d.on) {
#pragma unroll
                for (int r = 0; r < 4; r++) luma_line(v[sgm * 4 + r], d, s, maxv);
            }
        }
    }
    // ---- horizontal edge y = 8j: columns 0-3 and 4-7, across = rows ----------------------
#pragma unroll
    for (int sgm = 0; sgm < 2; sgm++) {
        const Seg s = sh[sgm];
        if (s.bs == 0) continue;
        if (cs) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int col = sgm * 4 + k;
                int l[8];
#pragma unroll
                for (int r = 0; r < 8; r++) l[r] = v[r][col];
                chroma_line(l, s, maxv);
                v[3][col] = l[3];
                v[4][col] = l[4];
            }
        } else {
            int la[8], lb[8];
#pragma unroll
            for (int r = 0; r < 8; r++) {
                la[r] = v[r][sgm * 4];
                lb[r] = v[r][sgm * 4 + 3];
            }
            const Dec d = decide(la, lb, s);
            if (d.on) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int col = sgm * 4 + k;
                    int l[8];
#pragma unroll
                    for (int r = 0; r < 8; r++) l[r] = v[r][col];
                    luma_line(l, d, s, maxv);
#pragma unroll
                    for (int r = 1; r < 7; r++) v[r][col] = l[r];
                }
            }
        }
    }

    // ---- store -----------------------------------------------------------------------------
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const int y = y0 + r;
        const bool row_ok = y >= 0 && y < h;
        if (row_ok && has_l) store4<T>(base + (size_t)y * stride + x0, v[r], 0);
        if (row_ok && has_r) store4<T>(base + (size_t)y * stride + x0 + 4, v[r], 4);
    }
}

int launch_deblock(p265_ctx *ctx, void *d_pix, const p265_pic_geom *g, int ctb_log2, const p265_dbk_blk *d_blk,
                   const p265_dbk_ctb *d_ctb) {
    DbkArgs a;
    a.pix = d_pix;
    a.blk = d_blk;
    a.ctb = d_ctb;
    for (int c = 0; c < 3; c++) a.plane_off[c] = g->plane_off[c];
    a.pic_stride = g->pic_stride;
    a.width = g->width;
    a.height = g->height;
    a.stride_y = g->stride_y;
    a.stride_c = g->stride_c;
    a.bit_depth_y = g->bit_depth_y;
    a.bit_depth_c = g->bit_depth_c;
    a.w8 = g->width / 8;
    a.h8 = g->height / 8;
    a.ctb_shift = ctb_log2 - 3;
    const int ctb = 1 << ctb_log2;
    a.ctbs_w = (g->width + ctb - 1) / ctb;
    a.ctbs_h = (g->height + ctb - 1) / ctb;
    a.chunks_y = ((a.w8 + 1) + 31) / 32;
    a.rows_y = a.h8 + 1;
    if (g->n_pics > 65535) return set_error(P265_EINVAL, "too many pictures in one deblocking batch");
    const int items = a.chunks_y * a.rows_y;
    const int warps = kDbkThreads / 32;
    const dim3 grid((unsigned)((items + warps - 1) / warps), 3, g->n_pics);
    if (g->bit_depth_y > 8 || g->bit_depth_c > 8) deblock_kernel<uint16_t><<<grid, kDbkThreads, 0, ctx->stream>>>(a);
    else deblock_kernel<uint8_t><<<grid, kDbkThreads, 0, ctx->stream>>>(a);
    P265_CUDA(cudaGetLastError());
    ctx->launches++;
    return P265_OK;
}

}  // namespace p265
