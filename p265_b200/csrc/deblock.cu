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
// CTB's slice offsets (beta' in closed form, tC' through a 54-byte read-only table).
#include <cuda_runtime.h>

#include "internal.h"

namespace p265 {

static_assert(sizeof(p265_dbk_ctb) == 4, "p265_dbk_ctb is read as one 32-bit word");
constexpr int kDbkThreads = 128;
#ifndef P265_DBK_CTAS
#define P265_DBK_CTAS 6  // 80 registers; measured best of 4..8 on the 4K 10-bit workload
#endif

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
    int32_t items_y, items_c;    // warp items (32 shifted blocks of one block row) per luma / chroma plane
};

__device__ const uint8_t g_tc[54] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1,  1,  1,  1,  1,  1,  1,  1,
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

// beta' of Table 8-11 in closed form: 0 below 16, Q - 10 up to 28, 2Q - 38 from 29 on
__device__ __forceinline__ int beta_prime(int q) { return q < 16 ? 0 : (q < 29 ? q - 10 : 2 * q - 38); }

// p265_dbk_ctb read as one 32-bit word: beta_offset_div2 | tc_offset_div2 << 8 | cb << 16 | cr << 24
__device__ __forceinline__ int ctb_field(uint32_t par, int byte) { return ((int)(par << (24 - 8 * byte))) >> 24; }

// 8.7.2.5.3 (beta, tC) for one segment; eq / ep = edge-map entries of the blocks holding q0 / p0
template <bool CHROMA>
__device__ __forceinline__ Seg make_seg(int bs, uint32_t eq, uint32_t ep, uint32_t par, int c, int bd) {
    Seg s;
    s.bs = CHROMA ? (bs == 2 ? 2 : 0) : bs;
    s.no_p = (ep & P265_DBK_NO_FILTER) != 0;
    s.no_q = (eq & P265_DBK_NO_FILTER) != 0;
    const int qpl = (blk_qp(eq) + blk_qp(ep) + 1) >> 1;
    const int tc_off = 2 * ctb_field(par, 1);
    if (CHROMA) {
        const int qpc = chroma_qp(qpl + ctb_field(par, c == 1 ? 2 : 3));
        s.tc = (int)__ldg(&g_tc[clip3i(0, 53, qpc + 2 + tc_off)]) << (bd - 8);
        s.beta = 0;
    } else {
        s.beta = beta_prime(clip3i(0, 51, qpl + 2 * ctb_field(par, 0))) << (bd - 8);
        s.tc = (int)__ldg(&g_tc[clip3i(0, 53, qpl + 2 * (s.bs - 1) + tc_off)]) << (bd - 8);
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


// ======================================================================================
// Packed path (bit depths <= 11): two lines per instruction.  A "line pair" is eight 32-bit
// registers l[0..7] = p3 p2 p1 p0 q0 q1 q2 q3, each holding the same tap of two adjacent lines
// as 2 x 16 bit.  All quantities are kept non-negative per half-word (biases are multiples of
// the following shift) so that plain 32-bit adds / subtracts / shifts never carry or borrow
// between the halves; clips are VIMNMX.S16x2 / VIADDMNMX.S16x2(.RELU).
// ======================================================================================
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
__device__ __forceinline__ uint32_t rep2(int v) { return ((uint32_t)v & 0xffffu) * 0x00010001u; }
__device__ __forceinline__ uint32_t sel2(uint32_t m, uint32_t a, uint32_t b) { return (a & m) | (b & ~m); }

// 8.7.2.5.7 strong filter on a line pair
__device__ __forceinline__ void strong_pair(uint32_t (&l)[8], int tc, bool no_p, bool no_q) {
    const uint32_t p3 = l[0], p2 = l[1], p1 = l[2], p0 = l[3], q0 = l[4], q1 = l[5], q2 = l[6], q3 = l[7];
    const uint32_t tp = rep2(2 * tc), tn = rep2(-2 * tc);
    const uint32_t K4 = 0x00040004u, K2 = 0x00020002u;
    const uint32_t s = p0 + q0, A = p1 + s, B = q1 + s;
    auto clipd = [&](uint32_t r, uint32_t x) { return __viaddmax_s16x2(x, tn, __viaddmin_s16x2(x, tp, r)); };
    if (!no_p) {
        l[3] = clipd((((p2 + q1 + K4) + (A << 1)) >> 3) & 0x1fff1fffu, p0);
        l[2] = clipd(((p2 + A + K2) >> 2) & 0x3fff3fffu, p1);
        l[1] = clipd(((((p3 + p2) << 1) + p2 + A + K4) >> 3) & 0x1fff1fffu, p2);
    }
    if (!no_q) {
        l[4] = clipd((((p1 + q2 + K4) + (B << 1)) >> 3) & 0x1fff1fffu, q0);
        l[5] = clipd(((q2 + B + K2) >> 2) & 0x3fff3fffu, q1);
        l[6] = clipd(((((q3 + q2) << 1) + q2 + B + K4) >> 3) & 0x1fff1fffu, q2);
    }
}

// 8.7.2.5.7 weak filter on a line pair (per-line condition abs(delta) < 10 * tC)
__device__ __forceinline__ void weak_pair(uint32_t (&l)[8], int tc, bool dep, bool deq, bool no_p, bool no_q,
                                          uint32_t maxv2) {
    const uint32_t p2 = l[1], p1 = l[2], p0 = l[3], q0 = l[4], q1 = l[5], q2 = l[6];
    // D = 9*(q0 - p0) - 3*(q1 - p1) + 8 + 0x6000  (|.| <= 12 * 2047 < 0x6000)
    const uint32_t X = (q0 << 3) + q0 + (p1 << 1) + p1, Y = (p0 << 3) + p0 + (q1 << 1) + q1;
    const uint32_t D = X + 0x60086008u - Y;
    const uint32_t db = (D >> 4) & 0x0fff0fffu;                            // delta + 0x600, in [1, 3071]
    const uint32_t ad = __vmaxu2(db, 0x0c000c00u - db) - 0x06000600u;      // abs(delta)
    const uint32_t t = ad + 0x80008000u - rep2(10 * tc);                   // bit 15 of a half: abs(delta) >= 10 tC
    const uint32_t ok = ((~t >> 15) & 0x00010001u) * 0xffffu;
    const uint32_t dc = __vminu2(__vmaxu2(db, 0x06000600u - rep2(tc)), 0x06000600u + rep2(tc));
    const uint32_t n600 = rep2(-0x600), n800 = rep2(-0x800);
    const uint32_t th = rep2(tc >> 1);
    const uint32_t mp = no_p ? 0u : ok, mq = no_q ? 0u : ok;
    l[3] = sel2(mp, __viaddmin_s16x2_relu(p0 + dc, n600, maxv2), p0);
    l[4] = sel2(mq, __viaddmin_s16x2_relu(q0 + 0x0c000c00u - dc, n600, maxv2), q0);
    if (dep) {
        const uint32_t avg = ((p2 + p0 + 0x00010001u) >> 1) & 0x7fff7fffu;
        const uint32_t E = avg + dc + 0x0a000a00u - p1;                    // (avg - p1 + delta) + 0x1000
        uint32_t e2 = (E >> 1) & 0x7fff7fffu;                              // (.. >> 1) + 0x800
        e2 = __vminu2(__vmaxu2(e2, 0x08000800u - th), 0x08000800u + th);
        l[2] = sel2(mp, __viaddmin_s16x2_relu(p1 + e2, n800, maxv2), p1);
    }
    if (deq) {
        const uint32_t avg = ((q2 + q0 + 0x00010001u) >> 1) & 0x7fff7fffu;
        const uint32_t E = avg + 0x16001600u - q1 - dc;                    // (avg - q1 - delta) + 0x1000
        uint32_t e2 = (E >> 1) & 0x7fff7fffu;
        e2 = __vminu2(__vmaxu2(e2, 0x08000800u - th), 0x08000800u + th);
        l[5] = sel2(mq, __viaddmin_s16x2_relu(q1 + e2, n800, maxv2), q1);
    }
}

// 8.7.2.5.8 chroma filter on a line pair (taps p1 p0 q0 q1 = l[2..5])
__device__ __forceinline__ void chroma_pair(uint32_t (&l)[8], int tc, bool no_p, bool no_q, uint32_t maxv2) {
    const uint32_t p1 = l[2], p0 = l[3], q0 = l[4], q1 = l[5];
    // D = 4*(q0 - p0) + p1 - q1 + 4 + 0x6000  (|.| <= 5 * 4095 < 0x6000)
    const uint32_t D = (q0 << 2) + p1 + 0x60046004u - ((p0 << 2) + q1);
    const uint32_t db = (D >> 3) & 0x1fff1fffu;                            // delta + 0xc00
    const uint32_t dc = __vminu2(__vmaxu2(db, 0x0c000c00u - rep2(tc)), 0x0c000c00u + rep2(tc));
    const uint32_t nc00 = rep2(-0xc00);
    if (!no_p) l[3] = __viaddmin_s16x2_relu(p0 + dc, nc00, maxv2);
    if (!no_q) l[4] = __viaddmin_s16x2_relu(q0 + 0x18001800u - dc, nc00, maxv2);
}

// one luma segment = two line pairs; `a` / `b` = the segment's first / last line, unpacked
__device__ __forceinline__ void luma_segment_packed(uint32_t (&u)[8], uint32_t (&w)[8], const int (&a)[8],
                                                    const int (&b)[8], const Seg &s, uint32_t maxv2) {
    const Dec d = decide(a, b, s);
    if (!d.on) return;
    if (d.strong) {
        strong_pair(u, s.tc, s.no_p, s.no_q);
        strong_pair(w, s.tc, s.no_p, s.no_q);
    } else {
        weak_pair(u, s.tc, d.dep, d.deq, s.no_p, s.no_q, maxv2);
        weak_pair(w, s.tc, d.dep, d.deq, s.no_p, s.no_q, maxv2);
    }
}

// half-row (4 samples) <-> two packed words
template <typename T>
__device__ __forceinline__ void loadw(const T *p, uint32_t &w0, uint32_t &w1);
template <>
__device__ __forceinline__ void loadw<uint16_t>(const uint16_t *p, uint32_t &w0, uint32_t &w1) {
    const uint2 v = *reinterpret_cast<const uint2 *>(p);
    w0 = v.x; w1 = v.y;
}
template <>
__device__ __forceinline__ void loadw<uint8_t>(const uint8_t *p, uint32_t &w0, uint32_t &w1) {
    const uint32_t v = *reinterpret_cast<const uint32_t *>(p);
    w0 = prmt(v, 0, 0x4140); w1 = prmt(v, 0, 0x4342);
}
template <typename T>
__device__ __forceinline__ void storew(T *p, uint32_t w0, uint32_t w1);
template <>
__device__ __forceinline__ void storew<uint16_t>(uint16_t *p, uint32_t w0, uint32_t w1) {
    *reinterpret_cast<uint2 *>(p) = make_uint2(w0, w1);
}
template <>
__device__ __forceinline__ void storew<uint8_t>(uint8_t *p, uint32_t w0, uint32_t w1) {
    *reinterpret_cast<uint32_t *>(p) = prmt(w0, w1, 0x6420);
}

// grid = (ceil(items / warps per CTA), 3 components, pictures); chroma planes use the first
// part of the luma-sized item range.
template <typename T, bool PACKED>
__global__ void __launch_bounds__(kDbkThreads, PACKED ? P265_DBK_CTAS : 1) deblock_kernel(const __grid_constant__ DbkArgs a) {
    // grid = (luma items + 2 x chroma items, 1, pictures); an item = 32 consecutive shifted blocks
    const int pic = blockIdx.z;
    int item = blockIdx.x * (kDbkThreads / 32) + (threadIdx.x >> 5);
    int c = 0;
    if (item >= a.items_y) {
        item -= a.items_y;
        c = 1;
        if (item >= a.items_c) {
            item -= a.items_c;
            c = 2;
            if (item >= a.items_c) return;
        }
    }
    const int cs = c ? 1 : 0;
    const int w = a.width >> cs, h = a.height >> cs;
    const int nbx = ((w + 3) >> 3) + 1;
    const int chunks = (nbx + 31) >> 5;
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

    // ---- packed path: the block's 16 loads go out before the segment set-up, which needs two
    // more dependent global reads (CTB parameters, tC') and ~400 instructions: their latency
    // hides behind that work instead of being waited for afterwards.
    // W[r][k]: row r, samples (2k, 2k+1) of the shifted block
    // Blocks that lie inside the picture (all but the first / last block row and column) take
    // unpredicated loads and stores off one row pointer: the per-row bounds tests and address
    // arithmetic were ~13 % of the kernel's instructions.
    uint32_t W[8][4];
    const bool interior = has_l && has_r && y0 >= 0 && y0 + 8 <= h;
    T *const blk0 = base + (ptrdiff_t)y0 * stride + x0;   // dereferenced only where the row / half exists
    if constexpr (PACKED) {
        if (interior) {
#pragma unroll
            for (int r = 0; r < 8; r++) {
                loadw<T>(blk0 + (ptrdiff_t)r * stride, W[r][0], W[r][1]);
                loadw<T>(blk0 + (ptrdiff_t)r * stride + 4, W[r][2], W[r][3]);
            }
        } else {
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const int y = y0 + r;
                const bool row_ok = y >= 0 && y < h;
                W[r][0] = W[r][1] = W[r][2] = W[r][3] = 0u;
                if (row_ok && has_l) loadw<T>(blk0 + (ptrdiff_t)r * stride, W[r][0], W[r][1]);
                if (row_ok && has_r) loadw<T>(blk0 + (ptrdiff_t)r * stride + 4, W[r][2], W[r][3]);
            }
        }
    }

    // ---- per-segment parameters ----------------------------------------------------------
    const uint32_t *cp = reinterpret_cast<const uint32_t *>(a.ctb) + (size_t)pic * a.ctbs_w * a.ctbs_h;
    auto ctb_of = [&](int bi, int bj) {
        const int ci = min(max(bi, 0) >> a.ctb_shift, a.ctbs_w - 1), cj = min(max(bj, 0) >> a.ctb_shift, a.ctbs_h - 1);
        return __ldg(cp + cj * a.ctbs_w + ci);
    };
    const uint32_t par11 = ctb_of(I, J), par10 = ctb_of(I, J - 1), par01 = ctb_of(I - 1, J);
    Seg sv[2], sh[2];
    if (cs) {
        sv[0] = make_seg<true>(bs_vu, e10, e00, par10, c, bd);
        sv[1] = make_seg<true>(bs_vl, e11, e01, par11, c, bd);
        sh[0] = make_seg<true>(bs_hl, e01, e00, par01, c, bd);
        sh[1] = make_seg<true>(bs_hr, e11, e10, par11, c, bd);
    } else {
        sv[0] = make_seg<false>(bs_vu, e10, e00, par10, c, bd);
        sv[1] = make_seg<false>(bs_vl, e11, e01, par11, c, bd);
        sh[0] = make_seg<false>(bs_hl, e01, e00, par01, c, bd);
        sh[1] = make_seg<false>(bs_hr, e11, e10, par11, c, bd);
    }

    if constexpr (PACKED) {
        const uint32_t maxv2 = rep2(maxv);
        // ---- vertical edge: pair rows (2rp, 2rp+1) -> X[col][rp], filter, transpose back -------
        if (sv[0].bs | sv[1].bs) {
            uint32_t X[8][4];
#pragma unroll
            for (int rp = 0; rp < 4; rp++) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    X[2 * k][rp] = prmt(W[2 * rp][k], W[2 * rp + 1][k], 0x5410);
                    X[2 * k + 1][rp] = prmt(W[2 * rp][k], W[2 * rp + 1][k], 0x7632);
                }
            }
#pragma unroll
            for (int sgm = 0; sgm < 2; sgm++) {
                const Seg s = sv[sgm];
                if (s.bs == 0) continue;
                uint32_t u[8], w2[8];
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    u[t] = X[t][2 * sgm];
                    w2[t] = X[t][2 * sgm + 1];
                }
                if (cs) {
                    chroma_pair(u, s.tc, s.no_p, s.no_q, maxv2);
                    chroma_pair(w2, s.tc, s.no_p, s.no_q, maxv2);
                } else {
                    int la[8], lb[8];
#pragma unroll
                    for (int t = 0; t < 8; t++) {
                        la[t] = (int)(u[t] & 0xffffu);      // row 4 sgm
                        lb[t] = (int)(w2[t] >> 16);         // row 4 sgm + 3
                    }
                    luma_segment_packed(u, w2, la, lb, s, maxv2);
                }
#pragma unroll
                for (int t = 1; t < 7; t++) {
                    X[t][2 * sgm] = u[t];
                    X[t][2 * sgm + 1] = w2[t];
                }
            }
#pragma unroll
            for (int rp = 0; rp < 4; rp++) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    W[2 * rp][k] = prmt(X[2 * k][rp], X[2 * k + 1][rp], 0x5410);
                    W[2 * rp + 1][k] = prmt(X[2 * k][rp], X[2 * k + 1][rp], 0x7632);
                }
            }
        }
        // ---- horizontal edge: the words already pair columns (2k, 2k+1) ----------------------
#pragma unroll
        for (int sgm = 0; sgm < 2; sgm++) {
            const Seg s = sh[sgm];
            if (s.bs == 0) continue;
            uint32_t u[8], w2[8];
#pragma unroll
            for (int r = 0; r < 8; r++) {
                u[r] = W[r][2 * sgm];
                w2[r] = W[r][2 * sgm + 1];
            }
            if (cs) {
                chroma_pair(u, s.tc, s.no_p, s.no_q, maxv2);
                chroma_pair(w2, s.tc, s.no_p, s.no_q, maxv2);
            } else {
                int la[8], lb[8];
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    la[r] = (int)(u[r] & 0xffffu);          // column 4 sgm
                    lb[r] = (int)(w2[r] >> 16);             // column 4 sgm + 3
                }
                luma_segment_packed(u, w2, la, lb, s, maxv2);
            }
#pragma unroll
            for (int r = 1; r < 7; r++) {
                W[r][2 * sgm] = u[r];
                W[r][2 * sgm + 1] = w2[r];
            }
        }
        if (interior) {
#pragma unroll
            for (int r = 0; r < 8; r++) {
                storew<T>(blk0 + (ptrdiff_t)r * stride, W[r][0], W[r][1]);
                storew<T>(blk0 + (ptrdiff_t)r * stride + 4, W[r][2], W[r][3]);
            }
        } else {
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const int y = y0 + r;
                const bool row_ok = y >= 0 && y < h;
                if (row_ok && has_l) storew<T>(blk0 + (ptrdiff_t)r * stride, W[r][0], W[r][1]);
                if (row_ok && has_r) storew<T>(blk0 + (ptrdiff_t)r * stride + 4, W[r][2], W[r][3]);
            }
        }
    } else {
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
    auto items_of = [](int w, int h) { return ((((w + 3) >> 3) + 1 + 31) >> 5) * (((h + 3) >> 3) + 1); };
    a.items_y = items_of(g->width, g->height);
    a.items_c = items_of(g->width / 2, g->height / 2);
    if (g->n_pics > 65535) return set_error(P265_EINVAL, "too many pictures in one deblocking batch");
    const int items = a.items_y + 2 * a.items_c;
    const int warps = kDbkThreads / 32;
    const dim3 grid((unsigned)((items + warps - 1) / warps), 1, g->n_pics);
    // packed 2 x 16-bit arithmetic needs 12 * maxVal < 0x6000: bit depths up to 11
    const int bd_max = g->bit_depth_y > g->bit_depth_c ? g->bit_depth_y : g->bit_depth_c;
    if (bd_max > 11) deblock_kernel<uint16_t, false><<<grid, kDbkThreads, 0, ctx->stream>>>(a);
    else if (bd_max > 8) deblock_kernel<uint16_t, true><<<grid, kDbkThreads, 0, ctx->stream>>>(a);
    else deblock_kernel<uint8_t, true><<<grid, kDbkThreads, 0, ctx->stream>>>(a);
    P265_CUDA(cudaGetLastError());
    ctx->launches++;
    return P265_OK;
}

}  // namespace p265
