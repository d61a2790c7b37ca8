// sm_100a SAO kernel: H.265 8.7.3 sample adaptive offset (band offset and the four
// edge-offset classes) over reconstructed pictures, out of place, driven by the per-CTB
// parameters sao.Sao.parse() produces (sao.py:15-136).  The reference has no SAO filter
// at all; parity is against the spec oracle.
//
// One warp per tile of (up to) 1024 samples of one CTB: 64x16 for a 64x64 luma CTB (four
// tiles per CTB), the whole CTB for 32x32 and smaller.  A lane owns an 8-sample-wide,
// R <= 4 rows tall strip; all R + 2 row loads of the strip (16 bytes at 10 bits, 8 bytes
// at 8 bits, coalesced) are issued before the first use, so every warp keeps six
// independent loads in flight -- the kernel is bandwidth-, not latency-limited.
// Horizontal neighbours come through warp shuffles, vertical ones from the lane's own
// registers; no shared memory.  All arithmetic is packed 2 x 16 bit in one 32-bit register:
//   1 + sign(c - n)  : (c + 0x4000_4000 - n), then add -0x3fff, min 2, relu in one
//                      VIADDMNMX.S16x2.RELU
//   SaoOffsetVal[..] : byte-permute LUT (PRMT with sign replication) over the CTB's four
//                      int8 offsets
//   Clip1(c + off)   : VIADDMNMX.S16x2.RELU
// Type / class / offsets are uniform per warp, so every branch is warp-uniform.
// CTBs inside the picture whose eight neighbours are all available take a mask-free edge path.
#include <cuda_runtime.h>
#include <stdlib.h>

#include "internal.h"

namespace p265 {

constexpr int kSaoWarpsPerCta = 4;
#ifndef P265_SAO_CTAS
#define P265_SAO_CTAS 8  // 64 registers -> 32 warps per SM
#endif

struct SaoArgs {
    const void *rec;
    void *out;
    const p265_sao_ctb *params;
    const uint8_t *no_filter;
    int64_t plane_off[3];
    int64_t pic_stride;
    int32_t width, height, stride_y, stride_c;
    int32_t bit_depth_y, bit_depth_c;
    int32_t ctb_log2, ctbs_w, ctbs_h, n_pics;
    int32_t ctbs;        // ctbs_w * ctbs_h
    int32_t tiles_y, tiles_c;  // 1024-sample tiles per luma / chroma CTB
    int32_t items;       // n_pics * ctbs * (tiles_y + 2 * tiles_c)
    int32_t pf_rows;     // L2 prefetch distance in CTB rows (0 = off), <= ctbs_h
    int32_t pf_quota;    // 128-byte lines each CTA prefetches (covers one full CTB row of all three planes)
};

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// A row of the strip: e[1..4] = the lane's 8 samples as 4 x u16x2, e[0] = word whose
// HIGH half is the sample left of the strip, e[5] = word whose LOW half is the sample
// right of it.
struct Row {
    uint32_t e[6];
};

template <typename T>
__device__ __forceinline__ void load8(const T *p, uint32_t (&w)[4]);
template <>
__device__ __forceinline__ void load8<uint16_t>(const uint16_t *p, uint32_t (&w)[4]) {
    const uint4 v = *reinterpret_cast<const uint4 *>(p);
    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
}
template <>
__device__ __forceinline__ void load8<uint8_t>(const uint8_t *p, uint32_t (&w)[4]) {
    const uint2 v = *reinterpret_cast<const uint2 *>(p);
    w[0] = prmt(v.x, 0, 0x4140); w[1] = prmt(v.x, 0, 0x4342);
    w[2] = prmt(v.y, 0, 0x4140); w[3] = prmt(v.y, 0, 0x4342);
}
template <typename T>
__device__ __forceinline__ void store8(T *p, const uint32_t (&w)[4]);
template <>
__device__ __forceinline__ void store8<uint16_t>(uint16_t *p, const uint32_t (&w)[4]) {
    *reinterpret_cast<uint4 *>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}
template <>
__device__ __forceinline__ void store8<uint8_t>(uint8_t *p, const uint32_t (&w)[4]) {
    *reinterpret_cast<uint2 *>(p) = make_uint2(prmt(w[0], w[1], 0x6420), prmt(w[2], w[3], 0x6420));
}

struct Strip {
    int lx, lanes_w;         // lane column inside the CTB, lanes per CTB row
    bool first_col, last_col;  // lane touches the CTB's left / right edge
    bool mem_left, mem_right;  // a sample exists in memory left / right of the CTB
};

// Loading a row is split in two so that several rows can be in flight at once:
// row_issue() only issues the global loads (strip + the two edge samples a CTB-border
// lane needs), row_finish() exchanges the horizontal neighbours through shuffles.
// Every lane of the warp must call row_finish(); `p` points at the lane's first sample.
template <typename T, bool HALO>
__device__ __forceinline__ void row_issue(Row &r, const T *p, const Strip &s) {
    uint32_t w[4];
    load8<T>(p, w);
    r.e[1] = w[0]; r.e[2] = w[1]; r.e[3] = w[2]; r.e[4] = w[3];
    if (HALO) {
        r.e[0] = (s.first_col && s.mem_left) ? ((uint32_t)p[-1] << 16) : 0u;
        r.e[5] = (s.last_col && s.mem_right) ? (uint32_t)p[8] : 0u;
    }
}
template <bool HALO>
__device__ __forceinline__ void row_finish(Row &r, const Strip &s) {
    if (HALO) {
        const uint32_t l = __shfl_up_sync(0xffffffffu, r.e[4], 1);
        const uint32_t rr = __shfl_down_sync(0xffffffffu, r.e[1], 1);
        if (!s.first_col) r.e[0] = l;
        if (!s.last_col) r.e[5] = rr;
    }
}

struct ItemConst {
    uint32_t pool_lo, pool_hi;  // int8 offsets, indexable by PRMT
    uint32_t maxv2;             // (1 << bitDepth) - 1 in both halves
    uint32_t band_k;            // (32 - band_position) in both halves
    int band_shift;
};

__device__ __forceinline__ uint32_t apply_offset(uint32_t c, uint32_t idx2, const ItemConst &k) {
    // idx2: LUT index (0..4) per half-word -> packed int16 offsets -> clip(c + off)
    const uint32_t j = prmt(idx2, 0, 0x4420);   // idx_lo | idx_hi << 8
    const uint32_t sel = j * 0x11u + 0x8080u;   // nibbles: idx, idx|8 (sign replicate)
    const uint32_t off = prmt(k.pool_lo, k.pool_hi, sel);
    return __viaddmin_s16x2_relu(c, off, k.maxv2);
}

__device__ __forceinline__ uint32_t edge_word(uint32_t c, uint32_t a, uint32_t b, const ItemConst &k) {
    uint32_t da = c + 0x40004000u - a;  // 0x4000 + (c - a) per half, no cross-half borrow
    uint32_t db = c + 0x40004000u - b;
    // 1 + sign(c - a) = max(min((c - a) + 1, 2), 0): add, min and relu are one VIADDMNMX
    da = __viaddmin_s16x2_relu(da, 0xc001c001u, 0x00020002u);
    db = __viaddmin_s16x2_relu(db, 0xc001c001u, 0x00020002u);
    const uint32_t idx2 = da + db;  // 2 + sign + sign: 0..4
    return apply_offset(c, idx2, k);               // LUT order folds the edgeIdx remap
}

__device__ __forceinline__ uint32_t band_word(uint32_t c, const ItemConst &k) {
    const uint32_t band = (c >> k.band_shift) & 0x001f001fu;
    const uint32_t rel = (band + k.band_k) & 0x001f001fu;  // (band - band_position) mod 32
    return apply_offset(c, __vminu2(rel, 0x00040004u), k);
}

// keep-original masks of a row: 0xffff per half-word whose neighbour is unavailable.
struct MaskCtx {
    uint32_t col_l[4], col_r[4], beyond[4];  // sample is first / last valid column / outside
    bool aL, aR, aU, aD, aUL, aUR, aDL, aDR;
};

template <int CLS>
__device__ __forceinline__ void row_mask(const MaskCtx &m, bool top, bool bot, uint32_t (&out)[4]) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t k = m.beyond[i];
        if (CLS == 0) {
            if (!m.aL) k |= m.col_l[i];
            if (!m.aR) k |= m.col_r[i];
        } else if (CLS == 1) {
            if ((top && !m.aU) || (bot && !m.aD)) k = 0xffffffffu;
        } else if (CLS == 2) {  // a = (-1,-1), b = (+1,+1)
            if (top) k |= (m.aUL ? 0u : m.col_l[i]) | (m.aU ? 0u : ~m.col_l[i]);
            else if (!m.aL) k |= m.col_l[i];
            if (bot) k |= (m.aDR ? 0u : m.col_r[i]) | (m.aD ? 0u : ~m.col_r[i]);
            else if (!m.aR) k |= m.col_r[i];
        } else {                // a = (+1,-1), b = (-1,+1)
            if (top) k |= (m.aUR ? 0u : m.col_r[i]) | (m.aU ? 0u : ~m.col_r[i]);
            else if (!m.aR) k |= m.col_r[i];
            if (bot) k |= (m.aDL ? 0u : m.col_l[i]) | (m.aD ? 0u : ~m.col_l[i]);
            else if (!m.aL) k |= m.col_l[i];
        }
        out[i] = k;
    }
}

template <int CLS, bool FULL = false>
__device__ __forceinline__ void edge_row(const Row &P, const Row &C, const Row &N, const ItemConst &k,
                                         const uint32_t (&keep)[4], uint32_t (&out)[4]) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t c = C.e[i + 1];
        uint32_t a, b;
        if (CLS == 0) {
            a = prmt(C.e[i], C.e[i + 1], 0x5432);
            b = prmt(C.e[i + 1], C.e[i + 2], 0x5432);
        } else if (CLS == 1) {
            a = P.e[i + 1];
            b = N.e[i + 1];
        } else if (CLS == 2) {
            a = prmt(P.e[i], P.e[i + 1], 0x5432);
            b = prmt(N.e[i + 1], N.e[i + 2], 0x5432);
        } else {
            a = prmt(P.e[i + 1], P.e[i + 2], 0x5432);
            b = prmt(N.e[i], N.e[i + 1], 0x5432);
        }
        const uint32_t r = edge_word(c, a, b, k);
        out[i] = FULL ? r : ((c & keep[i]) | (r & ~keep[i]));
    }
}

constexpr int kMaxRows = 4;  // rows per lane
struct Geo {
    int y0, vh, row0, rows;  // CTB top row, valid rows, lane's first row inside the CTB, rows per lane
    int h;                   // plane height
    int64_t stride;
    bool active;             // lane covers valid columns
};

// FULL: the CTB lies inside the picture with all eight neighbours available and no no-filter map
// (nine CTBs out of ten of a 4K picture): every sample is filtered, no keep masks at all.
template <typename T, int CLS, bool NOFILT, bool FULL = false>
__device__ __forceinline__ void edge_strip(const T *in, T *out, const Geo &g, const Strip &s, const MaskCtx &m,
                                           const ItemConst &k, const uint8_t *nf_row, int nf_stride, int nf_shift, bool nf_second) {
    // `in` / `out` point at (row y0, lane's first column)
    constexpr bool HALO = CLS != 1;
    auto rowptr = [&](int ly) {
        if (FULL) return in + (int64_t)ly * g.stride;  // rows -1 .. cs exist: no clamping
        int y = g.y0 + ly;
        y = y < 0 ? 0 : (y >= g.h ? g.h - 1 : y);
        return in + (int64_t)(y - g.y0) * g.stride;
    };
    // rows row0-1 .. row0+rows of the CTB: every load is issued before the first use
    Row rw[kMaxRows + 2];
#pragma unroll
    for (int j = 0; j < kMaxRows + 2; j++)
        if (j < g.rows + 2) row_issue<T, HALO>(rw[j], rowptr(g.row0 - 1 + j), s);
#pragma unroll
    for (int j = 0; j < kMaxRows + 2; j++)
        if (j < g.rows + 2) row_finish<HALO>(rw[j], s);
    if (FULL) {
#pragma unroll
        for (int j = 0; j < kMaxRows; j++) {
            if (j >= g.rows) break;  // warp-uniform
            uint32_t o[4];
            const uint32_t none[4] = {0u, 0u, 0u, 0u};
            edge_row<CLS, true>(rw[j], rw[j + 1], rw[j + 2], k, none, o);
            if (g.active) store8<T>(out + (int64_t)(g.row0 + j) * g.stride, o);
        }
        return;
    }
    uint32_t mid[4];
    row_mask<CLS>(m, false, false, mid);
#pragma unroll
    for (int j = 0; j < kMaxRows; j++) {
        if (j >= g.rows) break;  // warp-uniform
        const int ly = g.row0 + j;
        const bool top = ly == 0, bot = ly == g.vh - 1;
        uint32_t keep[4], o[4];
        if (top || bot) row_mask<CLS>(m, top, bot, keep);
        else {
#pragma unroll
            for (int i = 0; i < 4; i++) keep[i] = mid[i];
        }
        if (NOFILT) {
            const uint8_t *f = nf_row + (int64_t)(min(ly, g.vh - 1) >> nf_shift) * nf_stride;
            if (nf_shift == 3) {
                if (f[0]) keep[0] = keep[1] = keep[2] = keep[3] = 0xffffffffu;
            } else {
                if (f[0]) keep[0] = keep[1] = 0xffffffffu;
                if (nf_second && f[1]) keep[2] = keep[3] = 0xffffffffu;
            }
        }
        edge_row<CLS>(rw[j], rw[j + 1], rw[j + 2], k, keep, o);
        if (g.active && ly < g.vh) store8<T>(out + (int64_t)ly * g.stride, o);
    }
}

template <typename T, bool NOFILT>
__device__ __forceinline__ void sao_item(const SaoArgs &a, int pic, int c, int rx, int ry, int tile, int lane) {
    // field-by-field reads (a by-value copy indexed by `c` would live in local memory)
    // 32-bit CTB index (the launcher rejects batches with 2^31 or more CTBs): one IMAD.WIDE
    const p265_sao_ctb *qp = a.params + ((uint32_t)(pic * a.ctbs_h + ry) * (uint32_t)a.ctbs_w + (uint32_t)rx);
    struct {
        int band_pos, eo_class, avail;
        int offset_val[4];
    } q;
    const int type = qp->type[c];
    q.band_pos = qp->band_pos[c];
    q.eo_class = qp->eo_class[c];
    q.avail = qp->avail;
#pragma unroll
    for (int i = 0; i < 4; i++) q.offset_val[i] = qp->offset_val[c][i];
    const int cs_log2 = a.ctb_log2 - (c ? 1 : 0);
    const int cs = 1 << cs_log2;
    const int w = c ? a.width >> 1 : a.width, h = c ? a.height >> 1 : a.height;
    const int bd = c ? a.bit_depth_c : a.bit_depth_y;
    Geo g;
    g.stride = c ? a.stride_c : a.stride_y;
    g.h = h;
    const int x0 = rx * cs;
    g.y0 = ry * cs;
    const int vw = min(cs, w - x0);
    g.vh = min(cs, h - g.y0);

    // tile = th rows of the CTB starting at row tile * th; lanes_w x lane_rows lanes, R rows each
    // (all sizes are powers of two: shifts, no integer divisions)
    Strip s;
    const int lw_log2 = max(cs_log2 - 3, 0);
    s.lanes_w = 1 << lw_log2;
    const int th = min(cs, 1024 >> cs_log2);        // 16 (cs 64), 32, 16, 8
    g.rows = max(1, (th << lw_log2) >> 5);          // 4, 4, 1, 1
    s.lx = lane & (s.lanes_w - 1);
    const int ly = lane >> lw_log2;
    g.row0 = tile * th + ly * g.rows;
    g.active = s.lx * 8 < vw && g.row0 < g.vh && ly * g.rows < th;
    if (ly * g.rows >= th) {  // lane has no rows at all (8x8 chroma CTBs): park it on the tile's first row
        g.row0 = tile * th;
        g.active = false;
    }
    s.first_col = s.lx == 0;
    s.last_col = s.lx == s.lanes_w - 1;
    s.mem_left = x0 > 0;
    s.mem_right = x0 + cs < (int)g.stride;  // row padding is readable; masked if outside

    const int64_t base = (int64_t)pic * a.pic_stride + a.plane_off[c] + (int64_t)g.y0 * g.stride + x0 + s.lx * 8;
    const T *in = reinterpret_cast<const T *>(a.rec) + base;
    T *out = reinterpret_cast<T *>(a.out) + base;

    const uint8_t *nf_row = nullptr;
    int nf_stride = 0, nf_shift = 3;
    bool nf_second = true;
    if (NOFILT) {
        nf_stride = (a.width + 7) >> 3;
        const int h8 = (a.height + 7) >> 3;
        nf_shift = c ? 2 : 3;
        // luma: lane strip = one 8x8 block column; chroma: two 4x4 block columns
        const int bx = c ? ((x0 + s.lx * 8) >> 2) : ((x0 + s.lx * 8) >> 3);
        const int by0 = g.y0 >> nf_shift;
        // a chroma strip at the right picture edge may own a single 8x8 luma block column (odd
        // number of columns): its second flag does not exist
        nf_row = a.no_filter + ((int64_t)pic * h8 + by0) * nf_stride + min(bx, nf_stride - 1);
        if (c && bx + 1 >= nf_stride) nf_second = false;
    }

    ItemConst k;
    k.maxv2 = ((1u << bd) - 1u) * 0x00010001u;
    k.band_shift = bd - 5;
    k.band_k = (uint32_t)(32 - q.band_pos) * 0x00010001u;
    const uint32_t o1 = (uint8_t)q.offset_val[0], o2 = (uint8_t)q.offset_val[1];
    const uint32_t o3 = (uint8_t)q.offset_val[2], o4 = (uint8_t)q.offset_val[3];

    if (type != 2) {
        // off: copy; band: bandTable lookup (pool index = band - band_position, 4 = none)
        k.pool_lo = o1 | (o2 << 8) | (o3 << 16) | (o4 << 24);
        k.pool_hi = 0;
        uint32_t wv[kMaxRows][4];
#pragma unroll
        for (int j = 0; j < kMaxRows; j++) {
            const int lyr = g.row0 + j;
            if (j < g.rows && g.active && lyr < g.vh) load8<T>(in + (int64_t)lyr * g.stride, wv[j]);
        }
        if (type != 1) {  // off (warp-uniform): plain copy, no band arithmetic to select away
#pragma unroll
            for (int j = 0; j < kMaxRows; j++) {
                const int lyr = g.row0 + j;
                if (j < g.rows && g.active && lyr < g.vh) store8<T>(out + (int64_t)lyr * g.stride, wv[j]);
            }
            return;
        }
#pragma unroll
        for (int j = 0; j < kMaxRows; j++) {
            const int lyr = g.row0 + j;
            if (!(j < g.rows && g.active && lyr < g.vh)) continue;
            uint32_t o[4];
            bool skip[2] = {false, false};
            if (NOFILT) {
                const uint8_t *f = nf_row + (int64_t)(lyr >> nf_shift) * nf_stride;
                skip[0] = f[0] != 0;
                skip[1] = nf_shift == 3 ? skip[0] : (nf_second && f[1] != 0);
            }
#pragma unroll
            for (int i = 0; i < 4; i++) o[i] = skip[i >> 1] ? wv[j][i] : band_word(wv[j][i], k);
            store8<T>(out + (int64_t)lyr * g.stride, o);
        }
        return;
    }

    // edge offset: pool index = 2 + sign + sign -> {off1, off2, 0, off3, off4}
    k.pool_lo = o1 | (o2 << 8) | (o3 << 24);
    k.pool_hi = o4;
    MaskCtx m;
    const uint32_t av = q.avail;
    const bool has_l = x0 > 0, has_r = x0 + cs < w, has_u = g.y0 > 0, has_d = g.y0 + cs < h;
    if (!NOFILT && has_l && has_r && has_u && has_d && (av & 0x1efu) == 0x1efu) {  // warp-uniform; vw == vh == cs follows
        switch (q.eo_class) {
            case 0: edge_strip<T, 0, false, true>(in, out, g, s, m, k, nullptr, 0, 3, true); break;
            case 1: edge_strip<T, 1, false, true>(in, out, g, s, m, k, nullptr, 0, 3, true); break;
            case 2: edge_strip<T, 2, false, true>(in, out, g, s, m, k, nullptr, 0, 3, true); break;
            default: edge_strip<T, 3, false, true>(in, out, g, s, m, k, nullptr, 0, 3, true); break;
        }
        return;
    }
    m.aUL = has_u && has_l && (av >> 0 & 1);
    m.aU = has_u && (av >> 1 & 1);
    m.aUR = has_u && has_r && (av >> 2 & 1);
    m.aL = has_l && (av >> 3 & 1);
    m.aR = has_r && (av >> 5 & 1);
    m.aDL = has_d && has_l && (av >> 6 & 1);
    m.aD = has_d && (av >> 7 & 1);
    m.aDR = has_d && has_r && (av >> 8 & 1);
    // first / last valid column of the CTB and columns beyond the picture, as half-word
    // masks of this lane's four words (vw is a multiple of 4: 8 for luma, 4 for chroma)
    {
        const int xr = vw - 1 - s.lx * 8;  // position of the last valid column inside the strip
#pragma unroll
        for (int i = 0; i < 4; i++) {
            m.col_l[i] = (i == 0 && s.lx == 0) ? 0x0000ffffu : 0u;
            m.col_r[i] = (xr >> 1) == i ? ((xr & 1) ? 0xffff0000u : 0x0000ffffu) : 0u;
            m.beyond[i] = (2 * i > xr) ? 0xffffffffu : ((2 * i + 1 > xr) ? 0xffff0000u : 0u);
        }
    }
    switch (q.eo_class) {
        case 0: edge_strip<T, 0, NOFILT>(in, out, g, s, m, k, nf_row, nf_stride, nf_shift, nf_second); break;
        case 1: edge_strip<T, 1, NOFILT>(in, out, g, s, m, k, nf_row, nf_stride, nf_shift, nf_second); break;
        case 2: edge_strip<T, 2, NOFILT>(in, out, g, s, m, k, nf_row, nf_stride, nf_shift, nf_second); break;
        default: edge_strip<T, 3, NOFILT>(in, out, g, s, m, k, nf_row, nf_stride, nf_shift, nf_second); break;
    }
}

// Grid = (tiles of a CTB row / 4, CTB row, picture), 4 warps per CTA, one warp per tile.
// Tiles of a CTB: the luma tiles first (PER_CTB - 2 of them), then Cb, then Cr; all cover
// (up to) 1024 samples.  PER_CTB is a template constant so that no run-time division is
// left in the per-warp setup.
template <typename T, bool NOFILT, int PER_CTB>
__global__ void __launch_bounds__(kSaoWarpsPerCta * 32, P265_SAO_CTAS) sao_kernel(const __grid_constant__ SaoArgs a) {
    // Software prefetch into L2.  The kernel is latency-bound (every warp's loads depend on its
    // parameter record, and registers cap the bytes a warp keeps in flight), and its grid walks
    // the pictures CTB row by CTB row: the CTAs of a CTB row prefetch, one 128-byte line per
    // thread, the input rows of the CTB row `pf_rows` further down (the next picture's when the
    // current one ends), so the demand loads of those later CTAs find their lines in L2.
    // Only the first warp of a CTA prefetches (pf_quota lines per CTA, lane-strided), and the
    // target row is found without a division (the launcher keeps pf_rows <= ctbs_h): when every
    // thread derived its own line, this preamble was 16 % of the kernel's instructions.
    if (a.pf_rows && threadIdx.x < 32) {
        int rr = blockIdx.y + a.pf_rows, pp = blockIdx.z;
        if (rr >= a.ctbs_h) {
            rr -= a.ctbs_h;
            pp++;
        }
        if (pp < a.n_pics) {
            const int ctb = 1 << a.ctb_log2;
            const int rows_y = min(ctb, a.height - rr * ctb), rows_c = min(ctb >> 1, (a.height >> 1) - rr * (ctb >> 1));
            const int lines_y = (rows_y * a.stride_y * (int)sizeof(T)) >> 7;
            const int lines_c = (rows_c * a.stride_c * (int)sizeof(T)) >> 7;
            const char *base = reinterpret_cast<const char *>(reinterpret_cast<const T *>(a.rec) + (int64_t)pp * a.pic_stride);
            const char *py = base + (a.plane_off[0] + (int64_t)rr * ctb * a.stride_y) * (int64_t)sizeof(T);
            const char *pcb = base + (a.plane_off[1] + (int64_t)rr * (ctb >> 1) * a.stride_c) * (int64_t)sizeof(T);
            const char *pcr = base + (a.plane_off[2] + (int64_t)rr * (ctb >> 1) * a.stride_c) * (int64_t)sizeof(T);
            const int t0 = blockIdx.x * a.pf_quota;
            const int t1 = min(t0 + a.pf_quota, lines_y + 2 * lines_c);
            for (int t = t0 + (int)threadIdx.x; t < t1; t += 32) {
                const char *q = t < lines_y ? py + ((int64_t)t << 7)
                                : (t < lines_y + lines_c ? pcb + ((int64_t)(t - lines_y) << 7)
                                                         : pcr + ((int64_t)(t - lines_y - lines_c) << 7));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
            }
        }
    }
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * kSaoWarpsPerCta + (threadIdx.x >> 5);
    const int rx = item / PER_CTB, sub = item - rx * PER_CTB;
    if (rx >= a.ctbs_w) return;
    const int c = sub < PER_CTB - 2 ? 0 : sub - (PER_CTB - 3);
    const int tile = c ? 0 : sub;
    sao_item<T, NOFILT>(a, blockIdx.z, c, rx, blockIdx.y, tile, lane);
}

template <typename T, bool NOFILT>
static int launch_sao_t(p265_ctx *ctx, SaoArgs a, int n_pics) {
    const int per_ctb = a.tiles_y + 2 * a.tiles_c;  // 6 for 64x64 CTBs, 3 otherwise
    const dim3 grid((a.ctbs_w * per_ctb + kSaoWarpsPerCta - 1) / kSaoWarpsPerCta, a.ctbs_h, n_pics);
    {
        const int ctb = 1 << a.ctb_log2;
        const int64_t lines = ((int64_t)ctb * a.stride_y + 2 * (int64_t)(ctb >> 1) * a.stride_c) * (int64_t)sizeof(T) >> 7;
        a.pf_quota = (int)((lines + grid.x - 1) / grid.x);
    }
    const int threads = kSaoWarpsPerCta * 32;
    if (per_ctb == 6) sao_kernel<T, NOFILT, 6><<<grid, threads, 0, ctx->stream>>>(a);
    else sao_kernel<T, NOFILT, 3><<<grid, threads, 0, ctx->stream>>>(a);
    P265_CUDA(cudaGetLastError());
    ctx->launches++;
    return P265_OK;
}

int launch_sao(p265_ctx *ctx, const void *d_rec, void *d_out, const p265_pic_geom *g, int ctb_log2,
               const p265_sao_ctb *d_params, const uint8_t *d_no_filter) {
    SaoArgs a;
    a.rec = d_rec;
    a.out = d_out;
    a.params = d_params;
    a.no_filter = d_no_filter;
    for (int c = 0; c < 3; c++) a.plane_off[c] = g->plane_off[c];
    a.pic_stride = g->pic_stride;
    a.width = g->width;
    a.height = g->height;
    a.stride_y = g->stride_y;
    a.stride_c = g->stride_c;
    a.bit_depth_y = g->bit_depth_y;
    a.bit_depth_c = g->bit_depth_c;
    a.ctb_log2 = ctb_log2;
    const int ctb = 1 << ctb_log2;
    a.ctbs_w = (g->width + ctb - 1) / ctb;
    a.ctbs_h = (g->height + ctb - 1) / ctb;
    a.n_pics = g->n_pics;
    a.ctbs = a.ctbs_w * a.ctbs_h;
    const int cs_c = ctb >> 1;
    a.tiles_y = ctb / (ctb < 1024 / ctb ? ctb : 1024 / ctb);
    a.tiles_c = cs_c / (cs_c < 1024 / cs_c ? cs_c : 1024 / cs_c);
    a.items = 0;
    {
        static int pf = -1;  // tuning knob: P265_SAO_PF = prefetch distance in CTB rows
        if (pf < 0) {
            const char *e = getenv("P265_SAO_PF");
            pf = e ? atoi(e) : 4;  // measured on 4K 10-bit: 2..8 alike (0.87 of HBM), 0 = off 0.76, >= 24 worse
            if (pf < 0 || pf > 4096) pf = 4;
        }
        a.pf_rows = pf < a.ctbs_h ? pf : a.ctbs_h;  // the kernel wraps into the next picture at most once
    }
    if (a.ctbs_h > 65535 || g->n_pics > 65535) return set_error(P265_EINVAL, "too many CTB rows / pictures in one SAO batch");
    if ((int64_t)a.ctbs * g->n_pics >= ((int64_t)1 << 31)) return set_error(P265_EINVAL, "too many CTBs in one SAO batch");
    const bool wide = g->bit_depth_y > 8 || g->bit_depth_c > 8;
    if (wide) return d_no_filter ? launch_sao_t<uint16_t, true>(ctx, a, g->n_pics) : launch_sao_t<uint16_t, false>(ctx, a, g->n_pics);
    return d_no_filter ? launch_sao_t<uint8_t, true>(ctx, a, g->n_pics) : launch_sao_t<uint8_t, false>(ctx, a, g->n_pics);
}

}  // namespace p265
