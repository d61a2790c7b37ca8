// Residual path core: dequantisation (scaling.py:23-47, H.265 8.6.3) + two-stage inverse
// transform with the 16-bit clip between the stages (transform.py:89-106 structure,
// H.265 8.6.4.1-2) + final bdShift (8.6.2), for ONE warp work item.
//
// The code is written against a "lane" index instead of threadIdx so that the very same
// functions run (a) inside the sm_100a kernel, 32 lanes in lock step separated by
// __syncwarp(), and (b) on the host, lane after lane, phase after phase
// (tests/host_core.cpp) -- the indexing, permutations and packed constants can be
// checked bit-exactly against the oracle without a GPU.
//
// Arithmetic design (B200): every multiply-accumulate is an IDP.2A (dp2a): two int16
// data values packed in one register times two int8 basis coefficients from a uniform
// register, accumulated in int32.  All basis coefficients fit int8 (|c| <= 90), all
// data are int16 by the standard's clips, all sums stay below 2^27, so the result is
// exact.  The even/odd partial butterfly is kept (N^2/2 -> ~N^2/5.3 MACs for N=32) and
// halves again because one IDP.2A does two MACs.  Saturating packs (I2IP.S16.S32.SAT)
// implement both 16-bit clips for free while building the packed operands of the next
// stage.
//
// Work item = 64 TB columns: 2 TBs of 32x32, 4 of 16x16, 8 of 8x8 or 16 of 4x4.  Each
// lane owns two columns in stage 1 (the two columns that form one packed operand
// "slot" of stage 2) and two rows in stage 2.
#pragma once

#include <stdint.h>

#include "p265_b200.h"

#if defined(__CUDACC__)
#define P265_HD __host__ __device__ __forceinline__
#define P265_UNROLL _Pragma("unroll")
#else
#define P265_HD inline
#define P265_UNROLL
// host-only build (tests/host_core.cpp): minimal stand-ins for the CUDA vector types
struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
#endif

namespace p265 {

// ------------------------------------------------------------------ primitives
P265_HD int dp2a_lo(int a, int b, int c) {
#if defined(__CUDA_ARCH__)
    return __dp2a_lo(a, b, c);
#else
    return c + (int)(int16_t)(a & 0xffff) * (int)(int8_t)(b & 0xff) +
           (int)(int16_t)((uint32_t)a >> 16) * (int)(int8_t)((b >> 8) & 0xff);
#endif
}
P265_HD int dp2a_hi(int a, int b, int c) {
#if defined(__CUDA_ARCH__)
    return __dp2a_hi(a, b, c);
#else
    return c + (int)(int16_t)(a & 0xffff) * (int)(int8_t)((b >> 16) & 0xff) +
           (int)(int16_t)((uint32_t)a >> 16) * (int)(int8_t)((b >> 24) & 0xff);
#endif
}
// signed int16 pair x UNSIGNED byte pair (ScalingFactor bytes reach 255): IDP.2A.{LO,HI}.S16.U8
P265_HD int dp2a_lo_su(int a, uint32_t b, int c) {
#if defined(__CUDA_ARCH__)
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#else
    return c + (int)(int16_t)(a & 0xffff) * (int)(b & 0xff) + (int)(int16_t)((uint32_t)a >> 16) * (int)((b >> 8) & 0xff);
#endif
}
P265_HD int dp2a_hi_su(int a, uint32_t b, int c) {
#if defined(__CUDA_ARCH__)
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#else
    return c + (int)(int16_t)(a & 0xffff) * (int)((b >> 16) & 0xff) + (int)(int16_t)((uint32_t)a >> 16) * (int)(b >> 24);
#endif
}
// (f0, 0, 0, f1): bytes 2p and 2p+1 (p = 0, 1) of a word of four ScalingFactor bytes, placed so that
// dp2a_lo_su / dp2a_hi_su of a packed coefficient pair give level_lo * f0 and level_hi * f1 -- the
// extraction of both coefficients, of both factors and the two multiplications in three instructions
P265_HD uint32_t sf_pair(uint32_t m4, int p) {
#if defined(__CUDA_ARCH__)
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(m4), "r"(0u), "r"(p ? 0x3442u : 0x1440u));
    return d;
#else
    const uint32_t f0 = (m4 >> (16 * p)) & 0xff, f1 = (m4 >> (16 * p + 8)) & 0xff;
    return f0 | (f1 << 24);
#endif
}

// {lo, hi} int32 -> s16x2 with signed saturation
P265_HD int pack_sat(int lo, int hi) {
#if defined(__CUDA_ARCH__)
    int r;
    asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(r) : "r"(hi), "r"(lo));
    return r;
#else
    int l = lo < -32768 ? -32768 : (lo > 32767 ? 32767 : lo);
    int h = hi < -32768 ? -32768 : (hi > 32767 ? 32767 : hi);
    return (int)(((uint32_t)(uint16_t)(int16_t)h << 16) | (uint16_t)(int16_t)l);
#endif
}
// 2e - s in one 3-input add (IADD3 e, e, -s): the mirrored butterfly output e - O from e and
// s = e + O.  Wraps like the separate subtraction would (two's complement).
P265_HD int mirror(int e, int s) {
#if defined(__CUDA_ARCH__)
    return e + e - s;
#else
    return (int)((uint32_t)e + (uint32_t)e - (uint32_t)s);
#endif
}
P265_HD int ilog2(unsigned v) {
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)v);
#else
    return 31 - __builtin_clz(v);
#endif
}

// global -> shared tile copies: LDGSTS (cp.async) on the device so the tile of the NEXT
// work item streams in while the current one is being transformed; memcpy on the host.
P265_HD void copy16_async(void *smem_dst, const void *gmem_src) {
#if defined(__CUDA_ARCH__)
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
#else
    const uint4 v = *reinterpret_cast<const uint4 *>(gmem_src);
    *reinterpret_cast<uint4 *>(smem_dst) = v;
#endif
}
P265_HD void copy8_async(void *smem_dst, const void *gmem_src) {
#if defined(__CUDA_ARCH__)
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
#else
    const uint2 v = *reinterpret_cast<const uint2 *>(gmem_src);
    *reinterpret_cast<uint2 *>(smem_dst) = v;
#endif
}

// ------------------------------------------------------------------ basis tables
// H.265 8.6.4.2 transMatrix, rebuilt from its cosine structure (see oracle for the
// same derivation; tests compare both with transform.py:7-72).
struct Basis {
    int8_t m[32][32];
    constexpr Basis() : m() {
        const int mag[33] = {64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67, 64,
                             61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9,  4,  0};
        for (int j = 0; j < 32; j++)
            for (int i = 0; i < 32; i++) {
                int a = (j * (2 * i + 1)) % 128;
                if (a > 64) a = 128 - a;
                m[j][i] = (int8_t)(j == 0 ? 64 : (a <= 32 ? mag[a] : -mag[64 - a]));
            }
    }
};
constexpr int kDst[4][4] = {{29, 55, 74, 84}, {74, 74, 0, -74}, {84, -29, -74, 55}, {55, -84, 74, -29}};

constexpr int pack4(int b0, int b1, int b2, int b3) {
    return (int)((uint32_t)(b0 & 0xff) | ((uint32_t)(b1 & 0xff) << 8) | ((uint32_t)(b2 & 0xff) << 16) |
                 ((uint32_t)(b3 & 0xff) << 24));
}

// Packed odd-part constants.  For an N-point transform the odd part is
//   O[k] = sum over odd rows j of T_N[j][k] * x[j],  k < N/2,  T_N[j] = T_32[j*32/N].
// Word q of output k holds the coefficients of odd rows 8q+1, 8q+3 (dp2a.lo) and 8q+5,
// 8q+7 (dp2a.hi); the matching data registers are odd slots 2q and 2q+1.
struct Consts {
    int o32[16][4];
    int o16[8][2];
    int o8[4][1];
    int e4;     // (64, 64 | 64, -64)   on slot (x0, x2): lo -> E0, hi -> E1
    int o4;     // (83, 36 | 36, -83)   on slot (x1, x3): lo -> O0, hi -> O1
    int dst[4]; // DST-VII: (D[0][i], D[2][i] | D[1][i], D[3][i])
    constexpr Consts() : o32(), o16(), o8(), e4(0), o4(0), dst() {
        Basis b;
        for (int k = 0; k < 16; k++)
            for (int q = 0; q < 4; q++)
                o32[k][q] = pack4(b.m[8 * q + 1][k], b.m[8 * q + 3][k], b.m[8 * q + 5][k], b.m[8 * q + 7][k]);
        for (int k = 0; k < 8; k++)
            for (int q = 0; q < 2; q++)
                o16[k][q] = pack4(b.m[2 * (8 * q + 1)][k], b.m[2 * (8 * q + 3)][k], b.m[2 * (8 * q + 5)][k],
                                  b.m[2 * (8 * q + 7)][k]);
        for (int k = 0; k < 4; k++) o8[k][0] = pack4(b.m[4][k], b.m[12][k], b.m[20][k], b.m[28][k]);
        e4 = pack4(b.m[0][0], b.m[16][0], b.m[0][1], b.m[16][1]);
        o4 = pack4(b.m[8][0], b.m[24][0], b.m[8][1], b.m[24][1]);
        for (int i = 0; i < 4; i++) dst[i] = pack4(kDst[0][i], kDst[2][i], kDst[1][i], kDst[3][i]);
    }
};

#if defined(__CUDA_ARCH__)
#define P265_K(x) (g_consts.x)
#else
#define P265_K(x) (h_consts.x)
#endif
#if defined(__CUDACC__)
__constant__ Consts g_consts = Consts();
#endif
static const Consts h_consts = Consts();

// Which natural index (row of a column / column of a row) sits in packed-operand slot
// `s`, half `h` (0 = low 16 bits) of an N-point transform.  Slot order is the order the
// recursive even/odd split consumes its inputs: slot 0 = (0, N/2); slot 1 = (N/4, 3N/4);
// slots [2^L, 2^(L+1)) hold the odd rows of the (N >> (log2N-2-L))-point sub-transform.
P265_HD constexpr int slot_index(int n, int s, int h) {
    if (s == 0) return h * (n >> 1);
    const int l = s >= 8 ? 3 : (s >= 4 ? 2 : (s >= 2 ? 1 : 0));
    int r = s - (1 << l);
    int scale = (n >> 2) >> l;
    return scale * (4 * r + 2 * h + 1);
}
P265_HD int slot_index_rt(int n, int s, int h) {
    if (s == 0) return h * (n >> 1);
    int l = ilog2((unsigned)s);
    int r = s - (1 << l);
    int scale = (n >> 2) >> l;
    return scale * (4 * r + 2 * h + 1);
}

// ------------------------------------------------------------------ 1-D transforms
// C independent vectors in lock step (they share every uniform-register constant).
// p[c][s]: packed operands in slot order; out[c][i]: natural order, int32, includes
// `rnd` (the rounding offset of the shift that follows) exactly once per output.
//
// Z = zero-extent code (the parser knows the last significant position, tu.py:145-148): every input
// with natural index >= N >> Z is zero.  An input that is zero contributes nothing, so its products are
// simply not issued -- bit-exact by construction.  In slot order the non-zero inputs are the first
// max(1, 2^L >> Z) slots of every level [2^L, 2^(L+1)) (extent_slot_needed), and the even part of an
// N-point transform with extent N >> Z is an N/2-point transform with extent (N/2) >> Z: the same Z
// all the way down.  32-point pass: 172 / 88 / 46 IDP.2A for Z = 0 / 1 / 2.
P265_HD constexpr bool extent_slot_needed(int s, int z) {
    if (s == 0) return true;
    if (s == 1) return z < 2;  // rows (N/4, 3N/4)
    const int l = s >= 8 ? 3 : (s >= 4 ? 2 : 1);
    const int keep = (1 << l) >> z;
    return (s - (1 << l)) < (keep < 1 ? 1 : keep);
}

template <int N, int C, int Z = 0>
struct Idct;

template <int C, int Z>
struct Idct<4, C, Z> {
    static P265_HD void run(const int (&p)[C][2], int rnd, int (&out)[C][4]) {
        P265_UNROLL
        for (int c = 0; c < C; c++) {
            // the odd chain accumulates on top of the even value (no separate add); the
            // mirrored output is 2e - (e + o), one 3-input IADD3 on the ALU pipe -- the FMA
            // pipe (IDP.2A / IMAD, one warp-instruction per 2 cycles) is the scarce one
            const int e0 = dp2a_lo(p[c][0], P265_K(e4), rnd);
            const int e1 = dp2a_hi(p[c][0], P265_K(e4), rnd);
            if (Z >= 2) {
                // inputs 1 and 3 of the 4-point stage (natural rows N/4, 3N/4) lie beyond the extent
                out[c][0] = e0; out[c][3] = e0;
                out[c][1] = e1; out[c][2] = e1;
            } else if (C == 4) {
                // one-lane-per-4x4-TB path: latency-bound, keep the even and odd products independent
                const int o0 = dp2a_lo(p[c][1], P265_K(o4), 0);
                const int o1 = dp2a_hi(p[c][1], P265_K(o4), 0);
                out[c][0] = e0 + o0;
                out[c][3] = e0 - o0;
                out[c][1] = e1 + o1;
                out[c][2] = e1 - o1;
            } else {
                out[c][0] = dp2a_lo(p[c][1], P265_K(o4), e0);
                out[c][1] = dp2a_hi(p[c][1], P265_K(o4), e1);
                out[c][3] = mirror(e0, out[c][0]);
                out[c][2] = mirror(e1, out[c][1]);
            }
        }
    }
};

template <int N>
struct OddK;
template <>
struct OddK<8> {
    static P265_HD int w(int k, int q) { return P265_K(o8)[k][q]; }
};
template <>
struct OddK<16> {
    static P265_HD int w(int k, int q) { return P265_K(o16)[k][q]; }
};
template <>
struct OddK<32> {
    static P265_HD int w(int k, int q) { return P265_K(o32)[k][q]; }
};

template <int N, int C, int Z>
struct Idct {
    // odd-part slots that can hold a non-zero input
    static constexpr int SN = ((N / 4) >> Z) < 1 ? 1 : ((N / 4) >> Z);
    static P265_HD void run(const int (&p)[C][N / 2], int rnd, int (&out)[C][N]) {
        int pe[C][N / 4];
        int e[C][N / 2];
        P265_UNROLL
        for (int c = 0; c < C; c++) {
            P265_UNROLL
            for (int s = 0; s < N / 4; s++) pe[c][s] = p[c][s];
        }
        Idct<N / 2, C, Z>::run(pe, rnd, e);
        P265_UNROLL
        for (int k = 0; k < N / 2; k++) {
            int o[C];  // e + O[k]: the odd chain starts from the even value
            P265_UNROLL
            for (int c = 0; c < C; c++) o[c] = e[c][k];
            P265_UNROLL
            for (int s = 0; s < SN; s++) {
                const int w = OddK<N>::w(k, s >> 1);
                P265_UNROLL
                for (int c = 0; c < C; c++)
                    o[c] = (s & 1) ? dp2a_hi(p[c][N / 4 + s], w, o[c]) : dp2a_lo(p[c][N / 4 + s], w, o[c]);
            }
            P265_UNROLL
            for (int c = 0; c < C; c++) {
                out[c][k] = o[c];
                out[c][N - 1 - k] = mirror(e[c][k], o[c]);
            }
        }
    }
};

template <int C>
P265_HD void dst4(const int (&p)[C][2], int rnd, int (&out)[C][4]) {
    P265_UNROLL
    for (int c = 0; c < C; c++) {
        P265_UNROLL
        for (int i = 0; i < 4; i++) out[c][i] = dp2a_hi(p[c][1], P265_K(dst)[i], dp2a_lo(p[c][0], P265_K(dst)[i], rnd));
    }
}

// ------------------------------------------------------------------ per-TB parameters
struct TbParams {
    const int16_t *src;  // coefficients of this TB (arena)
    int16_t *dst;        // top-left of the TB in the residual plane
    const uint8_t *sf;   // ScalingFactor matrix for this TB ([y][x]) or nullptr
    int stride;          // plane stride (elements)
    int w;               // 16 * levelScale[qP % 6]  (or levelScale when sf != nullptr)
    int rnd, sh;         // dequant: (lv * m' + rnd) >> sh      (per < bdShift)
    int lsh;             // dequant: clip16(lv * m') << lsh      (per >= bdShift)
    int rnd2, sh2;       // final 8.6.2 shift: bdShift = 20 - BitDepth
    int flags;
    bool valid;
};

struct KernelArgs {
    const p265_tu_desc *tus;
    const uint4 *xtus;  // expanded descriptors (expand_desc), one per TB, same order as `tus`
    const int16_t *coeffs;
    const uint8_t *sf;  // P265_SF_BYTES or nullptr
    int32_t sf_replicated;  // 16x16 / 32x32 matrices obey the 7.4.5 up-sampling (+ DC at [0][0])
    int32_t dense_arena;    // P265_RES_DENSE_ARENA: inside the 8x8 and 4x4 bins, TB i's coefficients follow TB i-1's
    int16_t *out;
    int64_t plane_off[3];
    int64_t pic_stride;
    int32_t stride_y, stride_c;
    int32_t bit_depth_y, bit_depth_c;
    int32_t first_tb[4];  // index of the first descriptor of each size bin (32,16,8,4)
    int32_t n_tb[4];
    int32_t first_item[5];  // warp-item prefix sums per bin
    int32_t wait_prev;      // this launch directly follows expand_kernel: wait for it (griddepcontrol.wait)
    int32_t zext;           // P265_RES_ZERO_EXTENTS: the big bins honour the zero-extent codes of their records
};

P265_HD int sf_matrix_offset(int log2n, int c_idx, int intra) {
    // [sizeId][matrixId][y][x]; matrixId per scaling.py:33-42.  Branch-free: the four
    // sizeId bases {0, 96, 480, 2016} sit in one 64-bit constant (16 bits each).
    const int size_id = log2n - 2;
    const int base = (int)((0x07e001e000600000ull >> (16 * size_id)) & 0xffff);
    const int not_intra = intra ? 0 : 1;
    const int mid = log2n == 5 ? not_intra : c_idx + 3 * not_intra;
    return base + (mid << (2 * log2n));
}

// levelScale[rem] = {40,45,51,57,64,72} (scaling.py:28) as a byte LUT
P265_HD int level_scale(int rem) {
#if defined(__CUDA_ARCH__)
    return (int)(__byte_perm(0x39332d28u, 0x00004840u, (unsigned)rem | 0x7770u));  // bytes 1-3 <- pool byte 7 = 0
#else
    const uint64_t lut = 0x0000484039332d28ull;
    return (int)((lut >> (8 * rem)) & 0xff);
#endif
}

P265_HD uint4 load_desc(const KernelArgs &a, int tb_index, bool valid) {
    // 16-byte descriptor read as one vector
    return valid ? *reinterpret_cast<const uint4 *>(&a.tus[tb_index]) : make_uint4(0, 0, 0, 0);
}

P265_HD TbParams make_params(const KernelArgs &a, const uint4 d, bool valid);
P265_HD TbParams make_params(const KernelArgs &a, int tb_index, bool valid) {
    return make_params(a, load_desc(a, tb_index, valid), valid);
}
P265_HD TbParams make_params(const KernelArgs &a, const uint4 d, bool valid) {
    TbParams t;
    t.valid = valid;
    if (!valid) {
        t.src = a.coeffs; t.dst = a.out; t.sf = nullptr;  // never dereferenced; not nullptr so that the address space stays known
        t.stride = 0; t.w = 0; t.rnd = 0; t.sh = 0;
        t.lsh = 0; t.rnd2 = 0; t.sh2 = 0; t.flags = 0;
        return t;
    }
    const int x = (int)(d.x & 0xffff), y = (int)(d.x >> 16);
    const int log2n = (int)(d.y & 0xff), c_idx = (int)((d.y >> 8) & 0xff);
    const int qp = (int)((d.y >> 16) & 0xff);
    t.flags = (int)(d.y >> 24);
    const uint32_t coeff_off = d.z;
    const int pic = (int)(d.w & 0xffff);
    const int bit_depth = c_idx ? a.bit_depth_c : a.bit_depth_y;
    t.stride = c_idx ? a.stride_c : a.stride_y;
    t.src = a.coeffs + (size_t)coeff_off * 16;
    // element offset inside the residual buffer in 32 bits (the launcher rejects batches
    // whose planes exceed 2^32 elements)
    const uint32_t off = (uint32_t)pic * (uint32_t)a.pic_stride + (uint32_t)a.plane_off[c_idx] +
                         (uint32_t)y * (uint32_t)t.stride + (uint32_t)x;
    t.dst = a.out + off;
    const int per = (qp * 43) >> 8;  // qp / 6 for qp < 128
    const int rem = qp - per * 6;
    const int ls = level_scale(rem);
    const int bd_shift = bit_depth + log2n - 5;
    if (a.sf) {
        t.sf = a.sf + sf_matrix_offset(log2n, c_idx, (t.flags & P265_TU_INTRA) != 0);
        t.w = ls;
    } else {
        t.sf = nullptr;
        t.w = ls * 16;
    }
    const int ds = bd_shift - per;  // > 0: right shift with rounding; <= 0: clip, left shift
    t.sh = ds > 0 ? ds : 0;
    t.rnd = ds > 0 ? (1 << (ds - 1)) : 0;
    t.lsh = ds > 0 ? 0 : -ds;
    if (t.flags & P265_TU_PRESCALED) {  // arena holds d[] already: identity "dequantisation"
        t.sf = nullptr;
        t.w = 1;
        t.sh = 0;
        t.rnd = 0;
        t.lsh = 0;
    }
    t.sh2 = 20 - bit_depth;
    t.rnd2 = 1 << (t.sh2 - 1);
    return t;
}

P265_HD int sf_matrix_id(int log2n, int c_idx, int flags) {
    const int not_intra = (flags & P265_TU_INTRA) ? 0 : 1;
    return log2n == 5 ? not_intra : c_idx + 3 * not_intra;
}

// ------------------------------------------------------------------ expanded descriptors
// Everything make_params derives from a public 16-byte descriptor is TB-uniform, but inside
// the bin kernels every lane of every item would recompute it (~100 instructions per item).
// expand_kernel does it once per TB, into a 16-byte record whose fields the pipeline code uses
// as they are:
//   .x  element offset of the TB's top-left in the residual buffer
//   .y  flags | matrixId << 8 (3 bits; 6 = prescaled, no table) | (20 - BitDepth) << 11 (4 bits)
//       | (plane stride / 8) << 15
//   .z  coefficient arena offset, units of 16 coefficients       (unchanged)
//   .w  w (16 bits: 16 * levelScale, levelScale with a table, 1 when prescaled) | sh << 16 (5 bits)
//       | zr << 21 (2 bits) | lsh << 24 (4 bits) | zc << 28 (2 bits)
//       zr / zc = zero-extent codes of the TB (P265_TU_ZR / P265_TU_ZC of the public descriptor): rows >=
//       N >> zr and columns >= N >> zc hold no coefficient (0 = nothing known)
// (the TB size is the bin's; strides are multiples of 8 elements, checked by the launcher)
P265_HD uint32_t xd_flags(const uint4 x) { return x.y & 0xff; }
P265_HD uint32_t xd_mid(const uint4 x) { return (x.y >> 8) & 7; }
P265_HD int xd_sh2(const uint4 x) { return (int)((x.y >> 11) & 15); }
P265_HD int xd_stride(const uint4 x) { return (int)((x.y >> 15) << 3); }
P265_HD int xd_w(const uint4 x) { return (int)(x.w & 0xffff); }
P265_HD int xd_sh(const uint4 x) { return (int)((x.w >> 16) & 0x1f); }
P265_HD int xd_lsh(const uint4 x) { return (int)((x.w >> 24) & 0xf); }
P265_HD int xd_zr(const uint4 x) { return (int)((x.w >> 21) & 3); }
P265_HD int xd_zc(const uint4 x) { return (int)((x.w >> 28) & 3); }
// zero-extent codes of a public descriptor (p265_tu_desc.rsvd, bits 11-12 and 13-14); code 3 is not defined
// (the host entry points reject it) and reads as "nothing known"
P265_HD int desc_zr(const uint4 d) { const int z = (int)((d.w >> (16 + P265_TU_ZR_SHIFT)) & 3); return z == 3 ? 0 : z; }
P265_HD int desc_zc(const uint4 d) { const int z = (int)((d.w >> (16 + P265_TU_ZC_SHIFT)) & 3); return z == 3 ? 0 : z; }

P265_HD uint4 expand_desc(const KernelArgs &a, const uint4 d) {
    const TbParams t = make_params(a, d, true);
    const int log2n = (int)(d.y & 0xff), c_idx = (int)((d.y >> 8) & 0xff);
    const uint32_t dst_off = (uint32_t)(t.dst - a.out);
    // PRESCALED (identity dequantisation) is folded into w = 1, sh = lsh = 0 and "no table":
    // matrixId 6, where the small-bin kernels keep an all-ones matrix
    const int mid = (t.flags & P265_TU_PRESCALED) ? 6 : sf_matrix_id(log2n, c_idx, t.flags);
    const uint32_t flags = (uint32_t)t.flags & 0xff;
    // extents only mean something where a transform runs: a bypass TB's "coefficients" are residual samples and
    // the element-wise path reads all of them -- such a TB promises nothing (and with it its work item)
    // (likewise a TB with the rare left-shift dequantisation: its item runs the one full-extent copy of that pass)
    const bool elementwise = !a.zext || (t.flags & (P265_TU_SKIP | P265_TU_BYPASS)) != 0 || t.lsh != 0;
    const uint32_t zr = elementwise ? 0u : (uint32_t)desc_zr(d), zc = elementwise ? 0u : (uint32_t)desc_zc(d);
    return make_uint4(dst_off, flags | ((uint32_t)mid << 8) | ((uint32_t)t.sh2 << 11) | ((uint32_t)(t.stride >> 3) << 15),
                      d.z, (uint32_t)t.w | ((uint32_t)t.sh << 16) | (zr << 21) | ((uint32_t)t.lsh << 24) | (zc << 28));
}

P265_HD TbParams params_from_x(const KernelArgs &a, const uint4 x, bool valid, int log2n) {
    TbParams t;
    t.valid = valid;
    if (!valid) {
        t.src = a.coeffs; t.dst = a.out; t.sf = nullptr;  // never dereferenced; not nullptr so that the address space stays known
        t.stride = 0; t.w = 0; t.rnd = 0; t.sh = 0;
        t.lsh = 0; t.rnd2 = 0; t.sh2 = 0; t.flags = 0;
        return t;
    }
    t.flags = (int)xd_flags(x);
    t.src = a.coeffs + (size_t)x.z * 16;
    t.dst = a.out + x.x;
    t.stride = xd_stride(x);
    t.w = xd_w(x);
    t.sh = xd_sh(x);
    t.lsh = xd_lsh(x);
    t.rnd = (1 << t.sh) >> 1;
    t.sh2 = xd_sh2(x);
    t.rnd2 = 1 << (t.sh2 - 1);
    t.sf = nullptr;
    if (a.sf && !(t.flags & P265_TU_PRESCALED)) {
        const int base = (int)((0x07e001e000600000ull >> (16 * (log2n - 2))) & 0xffff);
        t.sf = a.sf + base + ((int)xd_mid(x) << (2 * log2n));
    }
    return t;
}

P265_HD uint4 load_xdesc(const KernelArgs &a, int tb_index, bool valid) {
    return valid ? a.xtus[tb_index] : make_uint4(0, 0, 0, 0);
}

// d = Clip3(-32768, 32767, (lv * m * levelScale << per + round) >> bdShift) without the
// final clip (the saturating pack that follows applies it).  Exact in int32:
// |lv * m * ls| <= 32768 * 255 * 72 < 2^30.
P265_HD int dequant_fast(int lv, int m, const TbParams &t) {  // valid when t.lsh == 0
    return (lv * m + t.rnd) >> t.sh;
}
P265_HD int dequant(int lv, int m, const TbParams &t) {
    int v = lv * m;
    if (t.lsh == 0) return (v + t.rnd) >> t.sh;
    v = v < -32768 ? -32768 : (v > 32767 ? 32767 : v);  // clip commutes with << for lsh >= 0
    return v << t.lsh;
}

// ------------------------------------------------------------------ shared layout
// Per warp: one coefficient tile buffer `in` and one stage-1 output buffer `g`, the same
// size (N*N int16 = N rows of N/2 packed words), plus one pad row per TB so that the TBs
// of a warp start in different banks.  `g` rows are XOR-swizzled in 16-byte chunks so that
// both the 16-bit column stores of stage 1 and the 128-bit row loads of stage 2 are free
// of bank conflicts without padding.
template <int LOG2N>
struct Layout {
    static constexpr int N = 1 << LOG2N;
    static constexpr int TPB = N / 2;         // lanes per TB
    static constexpr int TBS = 64 / N;        // TBs per warp item
    static constexpr int ROW_BYTES = N * 2;   // tile row: N int16; g row: N/2 packed words
    static constexpr int TB_BYTES = N * N * 2 + 2 * N;
    static constexpr int WARP_BYTES = TB_BYTES * TBS;
    // 16-byte chunk swizzle of g row y (chunks per row: N/8)
    static P265_HD constexpr int swz(int y) { return N == 32 ? ((y >> 1) & 3) : (N == 16 ? ((y >> 2) & 1) : 0); }
};
constexpr int kWarpSmemBytes = 4224;  // == Layout<5>::WARP_BYTES, multiple of 128
static_assert(Layout<5>::WARP_BYTES <= kWarpSmemBytes, "smem");
static_assert(Layout<4>::WARP_BYTES <= kWarpSmemBytes, "smem");
static_assert(Layout<3>::WARP_BYTES <= kWarpSmemBytes, "smem");
static_assert(Layout<2>::WARP_BYTES <= kWarpSmemBytes, "smem");

P265_HD int lds_s16(const unsigned char *smem, int byte_off) {
    return (int)*reinterpret_cast<const int16_t *>(smem + byte_off);
}
// int32 -> int16 with signed saturation (one I2IP against zero)
P265_HD uint16_t sat_s16(int v) { return (uint16_t)(pack_sat(v, 0) & 0xffff); }

// ---------------------------------------------------------------- phase 0: global -> smem
// Each lane copies its share (two columns' worth = 4N bytes) of its TB, 16 bytes at a
// time, fully coalesced across the TB's lanes.  Asynchronous on the device: the caller
// commits / waits the cp.async group and __syncwarp()s before anyone reads the tile.
// `rows` (a multiple of 4, <= N): only the first `rows` rows of the TB are copied -- what a zero-extent code
// leaves to read (iteration i of the loop moves rows 4i .. 4i+3 for both shared-memory sizes).
template <int LOG2N>
P265_HD void tile_issue(int lane, const int16_t *src, bool valid, unsigned char *in_base, int rows = 1 << LOG2N) {
    using L = Layout<LOG2N>;
    constexpr int N = L::N;
    const int tb = lane / L::TPB, tl = lane % L::TPB;
    if (!valid) return;
    unsigned char *in = in_base + tb * L::TB_BYTES;
    P265_UNROLL
    for (int i = 0; i < N / 4; i++) {
        if (N >= 16 && 4 * i >= rows) break;
        const int chunk = tl + i * L::TPB;  // 16-byte chunk index inside the TB
        if (N >= 8) {
            copy16_async(in + chunk * 16, src + chunk * 8);
        } else {  // 4x4 tiles are only 8-byte aligned in shared memory (40-byte pitch)
            copy8_async(in + chunk * 16, src + chunk * 8);
            copy8_async(in + chunk * 16 + 8, src + chunk * 8 + 4);
        }
    }
}

// transform-skip (8.6.4.2, tsShift = 7) / transquant-bypass (8.6.2) TBs: element-wise,
// straight from the tile to the plane.
template <int LOG2N>
P265_HD void phase_special(int lane, const TbParams &t, const unsigned char *in_base) {
    using L = Layout<LOG2N>;
    constexpr int N = L::N;
    const int tb = lane / L::TPB, tl = lane % L::TPB;
    if (!t.valid || !(t.flags & (P265_TU_SKIP | P265_TU_BYPASS))) return;
    const unsigned char *in = in_base + tb * L::TB_BYTES;
    for (int i = 0; i < N / 4; i++) {
        const int chunk = tl + i * L::TPB;
        uint32_t w[4];
        if (N >= 8) {
            const uint4 v = *reinterpret_cast<const uint4 *>(in + chunk * 16);
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        } else {
            const uint2 v0 = *reinterpret_cast<const uint2 *>(in + chunk * 16);
            const uint2 v1 = *reinterpret_cast<const uint2 *>(in + chunk * 16 + 8);
            w[0] = v0.x; w[1] = v0.y; w[2] = v1.x; w[3] = v1.y;
        }
        // 8 consecutive coefficients: one row segment (N >= 8) or two rows (N == 4)
        uint32_t o[4];
        P265_UNROLL
        for (int k = 0; k < 4; k++) {
            int r[2];
            P265_UNROLL
            for (int h = 0; h < 2; h++) {
                const int lv = (int)(int16_t)(h ? (w[k] >> 16) : (w[k] & 0xffff));
                if (t.flags & P265_TU_BYPASS) {
                    r[h] = lv;
                } else {
                    const int e = chunk * 8 + k * 2 + h;
                    const int m = t.sf ? (int)t.sf[e] * t.w : t.w;
                    int d = dequant(lv, m, t);
                    d = d < -32768 ? -32768 : (d > 32767 ? 32767 : d);
                    r[h] = (d * 128 + t.rnd2) >> t.sh2;
                }
            }
            o[k] = (uint32_t)pack_sat(r[0], r[1]);
        }
        const int e0 = chunk * 8;
        if (N >= 8) {
            int16_t *dst = t.dst + (size_t)(e0 / N) * t.stride + (e0 % N);
            *reinterpret_cast<uint4 *>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
        } else {
            int16_t *dst = t.dst + (size_t)(e0 / 4) * t.stride;
            *reinterpret_cast<uint2 *>(dst) = make_uint2(o[0], o[1]);
            *reinterpret_cast<uint2 *>(dst + t.stride) = make_uint2(o[2], o[3]);
        }
    }
}

// Scaling-factor access modes of stage 1.
//   SF_NONE        flat m = 16 (scaling_list_enabled_flag == 0, scaling.py:32-33)
//   SF_GENERAL     m[y][x] read per coefficient from the 4064-byte table (any table)
//   SF_REPLICATED  the table obeys 7.4.5: 16x16 / 32x32 entries are the 8x8 list
//                  up-sampled 2x / 4x with the DC value at [0][0]; a column then has only
//                  8 distinct factors, fetched once per column and pre-multiplied by
//                  levelScale -- no per-coefficient load or multiply is left.
enum { SF_NONE = 0, SF_GENERAL = 1, SF_REPLICATED = 2 };

// SF_REPLICATED: compact, transposed copy of the ScalingFactor matrices of one size,
// built once per CTA in shared memory: matrix m occupies kSfcStride bytes --
// [x8][y8] (the 8x8 list, column-major: the 8 factors of a column are 8 contiguous
// bytes) followed by the DC value at byte 64.
constexpr int kSfcStride = 80;
template <int LOG2N>
P265_HD void build_sf_compact(const uint8_t *table, uint8_t *out, int tid, int nthreads) {
    constexpr int N = 1 << LOG2N, REP = N / 8, NM = LOG2N == 5 ? 2 : 6;
    const uint8_t *base = table + sf_matrix_offset(LOG2N, 0, 1);  // matrixId 0 of this size
    for (int i = tid; i < NM * 65; i += nthreads) {
        const int m = i / 65, e = i - m * 65;
        const uint8_t *mat = base + m * N * N;
        if (e == 64) {
            out[m * kSfcStride + 64] = mat[0];
        } else {
            const int x8 = e >> 3, y8 = e & 7;
            out[m * kSfcStride + e] = mat[(y8 * REP + (REP > 1 ? 1 : 0)) * N + x8 * REP + (REP - 1)];
        }
    }
}

// ------------------------------------------------- stage 1: ONE column of one TB
// Dequantise column x of the tile (8.6.3), inverse-transform it (8.6.4.2, vertical pass),
// clip16((e + 64) >> 7) and store the N results as the `half`-th 16-bit half of slot
// `tl` in the rows of g.  Called twice per lane and item (the two columns that form one
// packed operand of stage 2); kept out of line on the device so that both calls share one
// copy of the code (instruction-cache footprint) and one register allocation.
// SLOW is chosen warp-uniformly: some TB of the warp has per >= bdShift (left-shift
// dequantisation, only reachable at very high qP on small TBs).
template <int LOG2N, int SF, bool SLOW>
P265_HD void stage1_column(const unsigned char *in, unsigned char *g, int x, int tl, int half, const uint8_t *sf,
                           int w, int rnd, int sh, int lsh, int dst_flag) {
    using L = Layout<LOG2N>;
    constexpr int N = L::N;
    constexpr int K = N < 8 ? N : 8;  // distinct factors per column (SF_REPLICATED)
    constexpr int REP = N / K;        // 1, 1, 2, 4
    if (SF != SF_NONE && sf == nullptr) return;  // lane without a TB (tail of a bin)
    TbParams t;                       // only the dequantisation fields are used
    t.rnd = rnd; t.sh = sh; t.lsh = lsh;
    int mw[K];
    int dc = 0;
    if (SF == SF_REPLICATED) {
        // `sf` = this TB's compact matrix (build_sf_compact): one 8-byte load per column
        const uint2 v = *reinterpret_cast<const uint2 *>(sf + (x / REP) * 8);
        P265_UNROLL
        for (int k = 0; k < K; k++) mw[k] = (int)(((k < 4 ? v.x : v.y) >> (8 * (k & 3))) & 0xff) * w;
        dc = (x == 0) ? (int)sf[64] * w : mw[0];
    }
    int p[1][N / 2];
    P265_UNROLL
    for (int s = 0; s < N / 2; s++) {
        const int y0 = slot_index(N, s, 0), y1 = slot_index(N, s, 1);
        const int e0 = y0 * N + x, e1 = y1 * N + x;
        const int l0 = lds_s16(in, e0 * 2), l1 = lds_s16(in, e1 * 2);
        int m0 = w, m1 = w;
        if (SF == SF_GENERAL) {
            m0 *= (int)sf[e0];
            m1 *= (int)sf[e1];
        } else if (SF == SF_REPLICATED) {
            m0 = y0 == 0 ? dc : mw[y0 / REP];
            m1 = mw[y1 / REP];
        }
        if (!SLOW) p[0][s] = pack_sat(dequant_fast(l0, m0, t), dequant_fast(l1, m1, t));
        else p[0][s] = pack_sat(dequant(l0, m0, t), dequant(l1, m1, t));
    }
    int e[1][N];
    if (N == 4 && dst_flag) {
        int p4[1][2] = {{p[0][0], p[0][1]}};
        int e4[1][4];
        dst4<1>(p4, 64, e4);
        P265_UNROLL
        for (int i = 0; i < 4; i++) e[0][i] = e4[0][i];
    } else {
        Idct<N, 1>::run(p, 64, e);
    }
    // g[y][slot tl].half = clip16((e[y] + 64) >> 7); chunk index XOR-swizzled per row
    unsigned char *base[4];
    P265_UNROLL
    for (int q = 0; q < 4; q++) base[q] = g + ((((tl >> 2) ^ q) << 4) | ((tl & 3) << 2) | (half << 1));
    P265_UNROLL
    for (int y = 0; y < N; y++)
        *reinterpret_cast<uint16_t *>(base[L::swz(y)] + y * L::ROW_BYTES) = sat_s16(e[0][y] >> 7);
}

// ------------------------------------------------- stage 1: BOTH columns of a lane in lock step
// Same arithmetic as two stage1_column calls (columns x0 = slot tl half 0, x1 = half 1), but
// the two columns share every basis constant and their results leave as one packed 32-bit
// store per row (half the I2IP / STS of the one-column form).  Needs twice the registers of
// the transform state, so it is used where that fits the occupancy target (16x16).
// Z (zero-extent code, see Idct): the TB's rows >= N >> Z hold no coefficient -- they are neither read
// nor dequantised, and the column transform skips their products.
template <int LOG2N, int SF, bool SLOW, int Z = 0>
P265_HD void stage1_pair(const unsigned char *in, unsigned char *g, int x0, int x1, int tl, const uint8_t *sf, int w,
                         int rnd, int sh, int lsh) {
    using L = Layout<LOG2N>;
    constexpr int N = L::N;
    constexpr int K = N < 8 ? N : 8;
    constexpr int REP = N / K;
    static_assert(N >= 8, "pair form is for the shared-memory sizes");
    if (SF != SF_NONE && sf == nullptr) return;
    TbParams t;
    t.rnd = rnd; t.sh = sh; t.lsh = lsh;
    int p[2][N / 2];
    P265_UNROLL
    for (int c = 0; c < 2; c++) {
        const int x = c ? x1 : x0;
        int mw[K];
        int dc = 0;
        if (SF == SF_REPLICATED) {
            const uint2 v = *reinterpret_cast<const uint2 *>(sf + (x / REP) * 8);
            P265_UNROLL
            for (int k = 0; k < K; k++) mw[k] = (int)(((k < 4 ? v.x : v.y) >> (8 * (k & 3))) & 0xff) * w;
            dc = (x == 0) ? (int)sf[64] * w : mw[0];
        }
        P265_UNROLL
        for (int s = 0; s < N / 2; s++) {
            if (!extent_slot_needed(s, Z)) {  // both rows of the slot lie beyond the extent
                p[c][s] = 0;
                continue;
            }
            const int y0 = slot_index(N, s, 0), y1 = slot_index(N, s, 1);
            const int e0 = y0 * N + x, e1 = y1 * N + x;
            const bool has1 = y1 < (N >> Z);  // the slot's second row may already be outside
            const int l0 = lds_s16(in, e0 * 2), l1 = has1 ? lds_s16(in, e1 * 2) : 0;
            int m0 = w, m1 = w;
            if (SF == SF_GENERAL) {
                m0 *= (int)sf[e0];
                if (has1) m1 *= (int)sf[e1];
            } else if (SF == SF_REPLICATED) {
                m0 = y0 == 0 ? dc : mw[y0 / REP];
                m1 = mw[y1 / REP];
            }
            if (!has1) p[c][s] = pack_sat(SLOW ? dequant(l0, m0, t) : dequant_fast(l0, m0, t), 0);
            else if (!SLOW) p[c][s] = pack_sat(dequant_fast(l0, m0, t), dequant_fast(l1, m1, t));
            else p[c][s] = pack_sat(dequant(l0, m0, t), dequant(l1, m1, t));
        }
    }
    int e[2][N];
    Idct<N, 2, Z>::run(p, 64, e);
    // g[y][slot tl] = (clip16((e0[y] + 64) >> 7), clip16((e1[y] + 64) >> 7)); chunk index XOR-swizzled per row
    unsigned char *base[4];
    P265_UNROLL
    for (int q = 0; q < 4; q++) base[q] = g + ((((tl >> 2) ^ q) << 4) | ((tl & 3) << 2));
    P265_UNROLL
    for (int y = 0; y < N; y++)
        *reinterpret_cast<uint32_t *>(base[L::swz(y)] + y * L::ROW_BYTES) = (uint32_t)pack_sat(e[0][y] >> 7, e[1][y] >> 7);
}

// ------------------------------------------------- stage 2: ONE row of one TB
// Horizontal pass (8.6.4.2) over row `row` of g, final bdShift rounding (8.6.2), int16
// saturation: the row leaves as N/2 packed words.
// Z: the TB's columns >= N >> Z hold no coefficient, so those columns of g are zero (dequantisation and
// the column transform map 0 to 0) and the row transform skips their products.
template <int LOG2N, int Z = 0>
P265_HD void stage2_compute(const unsigned char *g, int row, int rnd2, int sh2, int dst_flag,
                            uint32_t (&w)[(1 << LOG2N) / 2]) {
    using L = Layout<LOG2N>;
    constexpr int N = L::N;
    const unsigned char *grow = g + row * L::ROW_BYTES;
    const int sw = N == 32 ? ((row >> 1) & 3) : (N == 16 ? ((row >> 2) & 1) : 0);
    int p[1][N / 2];
    if (N >= 8) {
        P265_UNROLL
        for (int q = 0; q < N / 8; q++) {
            const uint4 v = *reinterpret_cast<const uint4 *>(grow + ((q ^ sw) << 4));
            p[0][4 * q + 0] = (int)v.x;
            p[0][4 * q + 1] = (int)v.y;
            p[0][4 * q + 2] = (int)v.z;
            p[0][4 * q + 3] = (int)v.w;
        }
    } else {
        const uint2 v = *reinterpret_cast<const uint2 *>(grow);
        p[0][0] = (int)v.x;
        p[0][1] = (int)v.y;
    }
    int r[1][N];
    if (N == 4 && dst_flag) {
        int p4[1][2] = {{p[0][0], p[0][1]}};
        int r4[1][4];
        dst4<1>(p4, rnd2, r4);
        P265_UNROLL
        for (int i = 0; i < 4; i++) r[0][i] = r4[0][i];
    } else {
        Idct<N, 1, Z>::run(p, rnd2, r);
    }
    P265_UNROLL
    for (int i = 0; i < N / 2; i++) w[i] = (uint32_t)pack_sat(r[0][2 * i] >> sh2, r[0][2 * i + 1] >> sh2);
}

// ... and the 2N-byte row store straight into the residual plane (one lane = one row: a warp's
// store instruction touches 32 different rows)
template <int LOG2N>
P265_HD void stage2_row(const unsigned char *g, int row, int16_t *dst, int rnd2, int sh2, int dst_flag) {
    constexpr int N = 1 << LOG2N;
    uint32_t w[N / 2];
    stage2_compute<LOG2N>(g, row, rnd2, sh2, dst_flag, w);
    if (N >= 8) {
        P265_UNROLL
        for (int q = 0; q < N / 8; q++)
            *reinterpret_cast<uint4 *>(dst + 8 * q) = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
    } else {
        *reinterpret_cast<uint2 *>(dst) = make_uint2(w[0], w[1]);
    }
}

// ... or back into the lane's own row of g (consumed by the loads above; same chunk swizzle), for
// the coalesced copy-out below: the scattered form costs 32 LSU wavefronts per store instruction
// (32 rows), measured as 16-20 % of the big-size bins.
template <int LOG2N, int Z = 0>
P265_HD void stage2_row_g(unsigned char *g, int row, int rnd2, int sh2) {
    using L = Layout<LOG2N>;
    constexpr int N = L::N;
    static_assert(N >= 16, "shared-memory sizes only");
    uint32_t w[N / 2];
    stage2_compute<LOG2N, Z>(g, row, rnd2, sh2, 0, w);
    unsigned char *grow = g + row * L::ROW_BYTES;
    const int sw = L::swz(row);
    P265_UNROLL
    for (int q = 0; q < N / 8; q++)
        *reinterpret_cast<uint4 *>(grow + ((q ^ sw) << 4)) = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
}

// Copy-out of a finished item: store instruction `i` of the warp moves rows 4i .. 4i+3 of EVERY TB of
// the item -- whole rows (64 bytes for 32x32, 32 bytes for 16x16) instead of one 16-byte chunk of 32
// different rows -- and a lane only ever touches the TB it also computes (lane / TPB), so it needs
// no other TB's record: destination = its own row pointer advanced by four rows per instruction.
template <int LOG2N>
struct OutMap {
    static constexpr int N = 1 << LOG2N;
    static constexpr int CPR = N / 8;                            // 16-byte chunks per row
    static constexpr int TPB = Layout<LOG2N>::TPB;               // lanes per TB
    static constexpr int RPI = TPB / CPR;                        // rows per TB and store instruction (4)
    static constexpr int ITERS = N / RPI;                        // store instructions per item
    static P265_HD int row0(int lane) { return (lane % TPB) / CPR; }   // row inside instruction 0
    static P265_HD int part(int lane) { return lane % CPR; }
};
// chunk of the lane in store instruction i, read from the lane's own TB buffer `g`
template <int LOG2N>
P265_HD uint4 out_chunk_load(const unsigned char *g, int i, int lane) {
    using L = Layout<LOG2N>;
    using M = OutMap<LOG2N>;
    const int row = i * M::RPI + M::row0(lane);
    return *reinterpret_cast<const uint4 *>(g + row * L::ROW_BYTES + ((M::part(lane) ^ L::swz(row)) << 4));
}

// ======================================================================================
// Small TBs: one lane = one TB, everything in registers.
// For 4x4 and 8x8 blocks the shared-memory choreography above costs more than the
// arithmetic (per-item overhead is amortised over only 8 / 16 samples per lane).  Here a
// work item is 32 TBs, one per lane: the lane reads its TB, dequantises, runs both
// transform stages and writes its rows without ever leaving its registers -- no
// transposes through shared memory, no warp synchronisation.
// ======================================================================================
P265_HD int sx_lo(uint32_t w) { return (int)(int16_t)(w & 0xffff); }
P265_HD int sx_hi(uint32_t w) { return (int)w >> 16; }

// transform-skip / bypass of a TB held as packed rows (N*N/2 words, row-major)
template <int N>
P265_HD void lane_special(const TbParams &t, const uint32_t *w, const uint8_t *sfm) {
    P265_UNROLL
    for (int y = 0; y < N; y++) {
        uint32_t o[N / 2];
        P265_UNROLL
        for (int k = 0; k < N / 2; k++) {
            int r[2];
            P265_UNROLL
            for (int h = 0; h < 2; h++) {
                const int lv = h ? sx_hi(w[y * (N / 2) + k]) : sx_lo(w[y * (N / 2) + k]);
                if (t.flags & P265_TU_BYPASS) {
                    r[h] = lv;
                } else {
                    const int m = sfm ? (int)sfm[y * N + 2 * k + h] * t.w : t.w;
                    int d = dequant(lv, m, t);
                    d = d < -32768 ? -32768 : (d > 32767 ? 32767 : d);
                    r[h] = (d * 128 + t.rnd2) >> t.sh2;
                }
            }
            o[k] = (uint32_t)pack_sat(r[0], r[1]);
        }
        int16_t *dst = t.dst + (size_t)y * t.stride;
        if (N == 4) *reinterpret_cast<uint2 *>(dst) = make_uint2(o[0], o[1]);
        else *reinterpret_cast<uint4 *>(dst) = make_uint4(o[0], o[1], o[N / 2 - 2], o[N / 2 - 1]);
    }
}

// 4x4: `w` = the TB's 8 packed words (rows 0..3, two words per row)
// `sf` = the TB's ScalingFactor matrix ([y][x] bytes) or nullptr: t.sf on the host, the CTA's
// shared-memory copy on the device (small_sf_ptr below).
template <int SF, bool SLOW>
P265_HD void tb4_lane(const TbParams &t, const uint32_t (&w)[8], const uint8_t *sf) {
    if (!t.valid) return;
    const uint8_t *sfm = (SF != SF_NONE) ? sf : nullptr;
    uint32_t mrow[4] = {0, 0, 0, 0};  // ScalingFactor bytes, one word per row
    if (SF != SF_NONE && sfm) {
        const uint4 mv = *reinterpret_cast<const uint4 *>(sfm);
        mrow[0] = mv.x; mrow[1] = mv.y; mrow[2] = mv.z; mrow[3] = mv.w;
    }
    if (t.flags & (P265_TU_SKIP | P265_TU_BYPASS)) {
        // transform-skip / bypass, element-wise (8.6.2, 8.6.4.2).  Same packed level * factor products as
        // the transform path below: the byte-by-byte form of lane_special (16 dependent shared-memory
        // byte loads per TB) made the 10 % transform-skip TBs of the 4K mix 20 % of this bin's stall samples
        if (SF != SF_NONE && !sfm) mrow[0] = mrow[1] = mrow[2] = mrow[3] = 0x01010101u;  // host only
        P265_UNROLL
        for (int y = 0; y < 4; y++) {
            uint32_t o[2];
            P265_UNROLL
            for (int pr = 0; pr < 2; pr++) {
                const uint32_t ww = w[y * 2 + pr];
                int r0, r1;
                if (t.flags & P265_TU_BYPASS) {
                    r0 = sx_lo(ww);
                    r1 = sx_hi(ww);
                } else {
                    int v0, v1;
                    if (SF != SF_NONE) {
                        const uint32_t f = sf_pair(mrow[y], pr);
                        v0 = dp2a_lo_su((int)ww, f, 0);
                        v1 = dp2a_hi_su((int)ww, f, 0);
                    } else {
                        v0 = sx_lo(ww);
                        v1 = sx_hi(ww);
                    }
                    int d0 = dequant(v0, t.w, t), d1 = dequant(v1, t.w, t);
                    d0 = d0 < -32768 ? -32768 : (d0 > 32767 ? 32767 : d0);
                    d1 = d1 < -32768 ? -32768 : (d1 > 32767 ? 32767 : d1);
                    r0 = (d0 * 128 + t.rnd2) >> t.sh2;
                    r1 = (d1 * 128 + t.rnd2) >> t.sh2;
                }
                o[pr] = (uint32_t)pack_sat(r0, r1);
            }
            *reinterpret_cast<uint2 *>(t.dst + (size_t)y * t.stride) = make_uint2(o[0], o[1]);
        }
        return;
    }
    int d[4][4];  // [y][x]
    if (SF != SF_NONE) {
        // level * factor by mixed-sign dp2a straight from the packed words (sf_pair), then * w:
        // (level * f) * w == level * (f * w), all below 2^30
        if (!sfm) mrow[0] = mrow[1] = mrow[2] = mrow[3] = 0x01010101u;  // host only: the device always has a matrix
        P265_UNROLL
        for (int y = 0; y < 4; y++) {
            P265_UNROLL
            for (int pr = 0; pr < 2; pr++) {
                const int ww = (int)w[y * 2 + pr];
                const uint32_t f = sf_pair(mrow[y], pr);
                const int v0 = dp2a_lo_su(ww, f, 0), v1 = dp2a_hi_su(ww, f, 0);
                d[y][2 * pr] = SLOW ? dequant(v0, t.w, t) : dequant_fast(v0, t.w, t);
                d[y][2 * pr + 1] = SLOW ? dequant(v1, t.w, t) : dequant_fast(v1, t.w, t);
            }
        }
    } else {
        P265_UNROLL
        for (int y = 0; y < 4; y++) {
            P265_UNROLL
            for (int x = 0; x < 4; x++) {
                const uint32_t ww = w[y * 2 + (x >> 1)];
                const int lv = (x & 1) ? sx_hi(ww) : sx_lo(ww);
                d[y][x] = SLOW ? dequant(lv, t.w, t) : dequant_fast(lv, t.w, t);
            }
        }
    }
    // stage 1: four columns in lock step; slot 0 = rows (0,2), slot 1 = rows (1,3)
    int p[4][2], e[4][4];
    P265_UNROLL
    for (int x = 0; x < 4; x++) {
        p[x][0] = pack_sat(d[0][x], d[2][x]);
        p[x][1] = pack_sat(d[1][x], d[3][x]);
    }
    if (t.flags & P265_TU_DST) dst4<4>(p, 64, e);
    else Idct<4, 4>::run(p, 64, e);
    // stage 2: four rows in lock step; slot 0 = columns (0,2), slot 1 = columns (1,3)
    int q[4][2], r[4][4];
    P265_UNROLL
    for (int y = 0; y < 4; y++) {
        q[y][0] = pack_sat(e[0][y] >> 7, e[2][y] >> 7);
        q[y][1] = pack_sat(e[1][y] >> 7, e[3][y] >> 7);
    }
    if (t.flags & P265_TU_DST) dst4<4>(q, t.rnd2, r);
    else Idct<4, 4>::run(q, t.rnd2, r);
    P265_UNROLL
    for (int y = 0; y < 4; y++) {
        const uint32_t o0 = (uint32_t)pack_sat(r[y][0] >> t.sh2, r[y][1] >> t.sh2);
        const uint32_t o1 = (uint32_t)pack_sat(r[y][2] >> t.sh2, r[y][3] >> t.sh2);
        *reinterpret_cast<uint2 *>(t.dst + (size_t)y * t.stride) = make_uint2(o0, o1);
    }
}

// 8x8: the TB sits in shared memory (lane-private 128 bytes, 16-byte chunks XOR-swizzled
// by the lane index so that the 128-bit row loads of a quarter-warp hit distinct banks).
P265_HD int tb8_chunk_off(int lane, int row) { return lane * 128 + ((row ^ (lane & 7)) << 4); }

template <int SF, bool SLOW>
P265_HD void tb8_lane(const TbParams &t, const unsigned char *tile, int lane, const uint8_t *sf) {
    if (!t.valid) return;
    const uint8_t *sfm = (SF != SF_NONE) ? sf : nullptr;
    if (t.flags & P265_TU_BYPASS) {
        uint32_t w[32];
        P265_UNROLL
        for (int y = 0; y < 8; y++) {
            const uint4 v = *reinterpret_cast<const uint4 *>(tile + tb8_chunk_off(lane, y));
            w[4 * y] = v.x; w[4 * y + 1] = v.y; w[4 * y + 2] = v.z; w[4 * y + 3] = v.w;
        }
        lane_special<8>(t, w, sfm);
        return;
    }
    // dequantise row pairs (0,4), (2,6), (1,3), (5,7) into packed stage-1 operands P[x][s]
    int P[8][4];
    P265_UNROLL
    for (int s = 0; s < 4; s++) {
        const int y0 = slot_index(8, s, 0), y1 = slot_index(8, s, 1);
        const uint4 a = *reinterpret_cast<const uint4 *>(tile + tb8_chunk_off(lane, y0));
        const uint4 b = *reinterpret_cast<const uint4 *>(tile + tb8_chunk_off(lane, y1));
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
        uint32_t ma[2] = {0, 0}, mb[2] = {0, 0};
        if (SF != SF_NONE && sfm) {
            const uint2 va = *reinterpret_cast<const uint2 *>(sfm + y0 * 8);
            const uint2 vb = *reinterpret_cast<const uint2 *>(sfm + y1 * 8);
            ma[0] = va.x; ma[1] = va.y; mb[0] = vb.x; mb[1] = vb.y;
        }
        if (SF != SF_NONE) {
            if (!sfm) ma[0] = ma[1] = mb[0] = mb[1] = 0x01010101u;  // host only: the device always has a matrix
            P265_UNROLL
            for (int pr = 0; pr < 4; pr++) {  // columns 2 pr, 2 pr + 1: level * factor by mixed-sign dp2a (sf_pair)
                const uint32_t fa = sf_pair(ma[pr >> 1], pr & 1), fb = sf_pair(mb[pr >> 1], pr & 1);
                const int a0 = dp2a_lo_su((int)aw[pr], fa, 0), a1 = dp2a_hi_su((int)aw[pr], fa, 0);
                const int b0 = dp2a_lo_su((int)bw[pr], fb, 0), b1 = dp2a_hi_su((int)bw[pr], fb, 0);
                if (!SLOW) {
                    P[2 * pr][s] = pack_sat(dequant_fast(a0, t.w, t), dequant_fast(b0, t.w, t));
                    P[2 * pr + 1][s] = pack_sat(dequant_fast(a1, t.w, t), dequant_fast(b1, t.w, t));
                } else {
                    P[2 * pr][s] = pack_sat(dequant(a0, t.w, t), dequant(b0, t.w, t));
                    P[2 * pr + 1][s] = pack_sat(dequant(a1, t.w, t), dequant(b1, t.w, t));
                }
            }
        } else {
            P265_UNROLL
            for (int x = 0; x < 8; x++) {
                const int la = (x & 1) ? sx_hi(aw[x >> 1]) : sx_lo(aw[x >> 1]);
                const int lb = (x & 1) ? sx_hi(bw[x >> 1]) : sx_lo(bw[x >> 1]);
                if (!SLOW) P[x][s] = pack_sat(dequant_fast(la, t.w, t), dequant_fast(lb, t.w, t));
                else P[x][s] = pack_sat(dequant(la, t.w, t), dequant(lb, t.w, t));
            }
        }
    }
    // stage 1: column pairs = stage-2 slots (0,4), (2,6), (1,3), (5,7)
    int G[8][4];
    P265_UNROLL
    for (int sl = 0; sl < 4; sl++) {
        const int x0 = slot_index(8, sl, 0), x1 = slot_index(8, sl, 1);
        int pp[2][4], e[2][8];
        P265_UNROLL
        for (int k = 0; k < 4; k++) {
            pp[0][k] = P[x0][k];
            pp[1][k] = P[x1][k];
        }
        Idct<8, 2>::run(pp, 64, e);
        P265_UNROLL
        for (int y = 0; y < 8; y++) G[y][sl] = pack_sat(e[0][y] >> 7, e[1][y] >> 7);
    }
    // stage 2: two rows at a time
    P265_UNROLL
    for (int y = 0; y < 8; y += 2) {
        int pp[2][4], r[2][8];
        P265_UNROLL
        for (int k = 0; k < 4; k++) {
            pp[0][k] = G[y][k];
            pp[1][k] = G[y + 1][k];
        }
        Idct<8, 2>::run(pp, t.rnd2, r);
        P265_UNROLL
        for (int c = 0; c < 2; c++) {
            uint32_t o[4];
            P265_UNROLL
            for (int k = 0; k < 4; k++) o[k] = (uint32_t)pack_sat(r[c][2 * k] >> t.sh2, r[c][2 * k + 1] >> t.sh2);
            *reinterpret_cast<uint4 *>(t.dst + (size_t)(y + c) * t.stride) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

// TBs per warp work item, by bin (32x32, 16x16, 8x8, 4x4): the two big sizes use the
// shared-memory column/row scheme (64 columns per item), the two small ones one lane per TB.
P265_HD constexpr int tbs_per_item(int bin) { return bin == 0 ? 2 : (bin == 1 ? 4 : 32); }

// TB index handled by `lane` of warp item `item` in bin LOG2N (bins: 0 -> 32x32 ...)
template <int LOG2N>
P265_HD int lane_tb(const KernelArgs &a, int item, int lane, bool &valid) {
    using L = Layout<LOG2N>;
    constexpr int bin = 5 - LOG2N;
    const int local = item * L::TBS + lane / L::TPB;
    valid = local < a.n_tb[bin];
    return a.first_tb[bin] + local;
}

}  // namespace p265
