// Packed coefficient stream -> dense coefficient arena (include/p265_b200.h, P265_TU_LEVELS8).
//
// The host -> device transport of TransCoeffLevel is sparse (significance bitmap + the non-zero
// levels, tu.py:331 stores exactly those); the residual kernels want TB-contiguous dense int16
// tiles.  unpack_kernel expands one into the other on the device, where the bytes are cheap:
// 3.7 MB in, 25 MB out per 4K picture at HBM speed against 25 MB over PCIe.
//
// One lane = one 32-bit word of a TB's bitmap = 32 coefficients = 64 bytes of the arena (a 4x4 TB
// is half a word).  A warp item is 32 words: 1 TB of 32x32, 4 of 16x16, 16 of 8x8, 32 of 4x4.  The
// index of a lane's first level is the number of set bits in the TB's earlier words: a segmented
// warp scan.  The arena is written in descriptor order, so the small bins come out dense
// (P265_RES_DENSE_ARENA) and a TB's offset follows from its index -- no prefix sum over TBs.
#include <cuda_runtime.h>

#include "internal.h"

namespace p265 {

struct UnpackArgs {
    const p265_tu_desc *tus;
    p265_tu_desc *tus_out;
    const uint8_t *stream;
    int16_t *arena;
    int32_t first_tb[4], n_tb[4];
    uint32_t arena_base[4];  // first arena unit (16 coefficients) of each bin
    int32_t first_item[5];
};

template <int LOG2N>
__device__ __forceinline__ void unpack_item(const UnpackArgs &a, int item, int lane) {
    constexpr int bin = 5 - LOG2N, NN = 1 << (2 * LOG2N);
    constexpr int WPT = NN >= 32 ? NN / 32 : 1;  // bitmap words (lanes) per TB
    constexpr int BITS = NN >= 32 ? 32 : 16;     // coefficients per lane
    constexpr int TBS = 32 / WPT;
    const int tb = item * TBS + lane / WPT, wi = lane % WPT;
    const bool valid = tb < a.n_tb[bin];
    uint4 d = make_uint4(0, 0, 0, 0);
    if (valid) d = *reinterpret_cast<const uint4 *>(&a.tus[a.first_tb[bin] + tb]);
    const uint8_t *rec = a.stream + (size_t)d.z * 4;
    uint32_t word = 0;
    if (valid) word = BITS == 32 ? reinterpret_cast<const uint32_t *>(rec)[wi] : (uint32_t) * reinterpret_cast<const uint16_t *>(rec);
    // exclusive prefix of the set bits inside the TB's lanes
    int incl = __popc(word);
#pragma unroll
    for (int o = 1; o < WPT; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o, WPT);
        if (wi >= o) incl += v;
    }
    int idx = incl - __popc(word);
    const int n_levels = (int)((d.w >> 16) & P265_TU_LEVELS_MASK);  // descriptor field rsvd: the record holds exactly this many levels --
                                            // never read past them, whatever the bitmap claims (the host entry
                                            // point bounds records by this field, not by counting bits)
    const bool narrow = ((d.y >> 24) & P265_TU_LEVELS8) != 0;
    const uint8_t *lv = rec + NN / 8;
    uint32_t out[BITS / 2];
#pragma unroll
    for (int b = 0; b < BITS; b++) {
        uint32_t v = 0;
        if (((word >> b) & 1u) && idx < n_levels) {
            v = narrow ? (uint32_t)(int)reinterpret_cast<const int8_t *>(lv)[idx]
                       : (uint32_t)reinterpret_cast<const uint16_t *>(lv)[idx];
            idx++;
        }
        if (b & 1) out[b >> 1] |= v << 16;
        else out[b >> 1] = v & 0xffffu;
    }
    // Zero-extent codes of the 16x16 / 32x32 TBs, straight from the significance bitmap (a bit that is set
    // beyond n_levels only makes the extent larger: conservative).  The residual kernels skip the products
    // of rows >= N >> zr and columns >= N >> zc (residual.cu: "Zero-aware passes").
    uint32_t zbits = 0;
    if (LOG2N >= 4) {
        constexpr int N = 1 << LOG2N;
        int last_row, last_col;
        if (LOG2N == 5) {  // one TB per warp, lane = row
            const uint32_t rows = __ballot_sync(0xffffffffu, word != 0);
            const uint32_t cols = __reduce_or_sync(0xffffffffu, word);
            last_row = 31 - __clz((int)rows);
            last_col = 31 - __clz((int)cols);
        } else {           // 8 lanes per TB, lane = two rows of 16 columns
            int r = word ? 2 * wi + ((word >> 16) ? 1 : 0) : -1;
            uint32_t c = (word | (word >> 16)) & 0xffffu;
#pragma unroll
            for (int o = 1; o < WPT; o <<= 1) {
                r = max(r, __shfl_xor_sync(0xffffffffu, r, o, WPT));
                c |= __shfl_xor_sync(0xffffffffu, c, o, WPT);
            }
            last_row = r;
            last_col = 31 - __clz((int)c);
        }
        const uint32_t zr = last_row < N / 4 ? 2u : (last_row < N / 2 ? 1u : 0u);
        const uint32_t zc = last_col < N / 4 ? 2u : (last_col < N / 2 ? 1u : 0u);
        zbits = (zr << (16 + P265_TU_ZR_SHIFT)) | (zc << (16 + P265_TU_ZC_SHIFT));
    }
    if (!valid) return;
    const uint32_t unit = a.arena_base[bin] + (uint32_t)tb * (NN / 16);
    uint4 *dst = reinterpret_cast<uint4 *>(a.arena + (size_t)unit * 16 + (size_t)wi * BITS);
#pragma unroll
    for (int q = 0; q < BITS / 8; q++) dst[q] = make_uint4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
    if (wi == 0) {
        d.z = unit;
        d.w = (d.w & 0xffffu) | zbits;  // the level count is a packed-stream field; the extent codes stay
        d.y &= ~((uint32_t)P265_TU_LEVELS8 << 24);
        *reinterpret_cast<uint4 *>(&a.tus_out[a.first_tb[bin] + tb]) = d;
    }
}

__global__ void __launch_bounds__(256) unpack_kernel(const __grid_constant__ UnpackArgs a) {
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= a.first_item[4]) return;
    if (item < a.first_item[1]) unpack_item<5>(a, item, lane);
    else if (item < a.first_item[2]) unpack_item<4>(a, item - a.first_item[1], lane);
    else if (item < a.first_item[3]) unpack_item<3>(a, item - a.first_item[2], lane);
    else unpack_item<2>(a, item - a.first_item[3], lane);
}

int launch_unpack(p265_ctx *ctx, const p265_tu_desc *d_tus, const int32_t bin_counts[4], const uint8_t *d_stream,
                  int16_t *d_arena, p265_tu_desc *d_tus_out) {
    UnpackArgs a;
    a.tus = d_tus;
    a.tus_out = d_tus_out;
    a.stream = d_stream;
    a.arena = d_arena;
    int64_t first = 0, items = 0, unit = 0;
    for (int b = 0; b < 4; b++) {
        const int nn = 1 << (2 * (5 - b));
        const int tbs = nn >= 32 ? 32 / (nn / 32) : 32;
        a.first_tb[b] = (int32_t)first;
        a.n_tb[b] = bin_counts[b];
        a.arena_base[b] = (uint32_t)unit;
        a.first_item[b] = (int32_t)items;
        first += bin_counts[b];
        unit += (int64_t)bin_counts[b] * (nn / 16);
        items += (bin_counts[b] + tbs - 1) / tbs;
    }
    if (first > INT32_MAX || items > INT32_MAX || unit > 0xffffffffll)
        return set_error(P265_EINVAL, "too many TBs in one batch");
    a.first_item[4] = (int32_t)items;
    if (!items) return P265_OK;
    unpack_kernel<<<(unsigned)((items + 7) / 8), 256, 0, ctx->stream>>>(a);
    P265_CUDA(cudaGetLastError());
    ctx->launches++;
    return P265_OK;
}

}  // namespace p265
