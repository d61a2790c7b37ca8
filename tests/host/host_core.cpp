// Host emulation of one residual launch: runs the SAME per-lane phase functions the
// sm_100a kernel runs (p265_b200/csrc/residual_core.cuh), lane after lane, phase after
// phase -- the only synchronisation in the kernel is __syncwarp() between phases, so
// this is an exact functional model of the device code's indexing and arithmetic.
// Built and driven by tests/test_host_core.py; never part of the product library.
#include <cstring>
#include <vector>

#include "../../p265_b200/csrc/residual_core.cuh"

using namespace p265;

template <int LOG2N, int SF>
static void run_item_sf(const KernelArgs &a, int item) {
    using L = Layout<LOG2N>;
    constexpr int N = L::N;
    alignas(16) unsigned char in_buf[kWarpSmemBytes], g_buf[kWarpSmemBytes];
    std::memset(in_buf, 0xA5, sizeof in_buf);
    std::memset(g_buf, 0x5A, sizeof g_buf);
    TbParams t[32];
    int lane_tb_index[32];
    bool slow = false;
    int zr = 2, zc = 2;  // the item's zero-extent codes: the weakest promise among its TBs (run_bin)
    for (int lane = 0; lane < 32; lane++) {
        bool valid;
        int tb = lane_tb<LOG2N>(a, item, lane, valid);
        lane_tb_index[lane] = valid ? tb : 0;
        const uint4 x = valid ? expand_desc(a, load_desc(a, tb, true)) : make_uint4(0, 0, 0, 0);
        t[lane] = params_from_x(a, x, valid, LOG2N);
        slow |= t[lane].lsh != 0;
        if (valid) {
            zr = xd_zr(x) < zr ? xd_zr(x) : zr;
            zc = xd_zc(x) < zc ? xd_zc(x) : zc;
        }
    }
    // the kernel copies only the rows the item's row code leaves (the rest of the buffer keeps the 0xA5 fill:
    // a pass that read beyond its extent would not match the oracle)
    for (int lane = 0; lane < 32; lane++) {
        if constexpr (LOG2N >= 4) tile_issue<LOG2N>(lane, t[lane].src, t[lane].valid, in_buf, N >> zr);
        else tile_issue<LOG2N>(lane, t[lane].src, t[lane].valid, in_buf);
    }
    for (int lane = 0; lane < 32; lane++) phase_special<LOG2N>(lane, t[lane], in_buf);
    alignas(16) uint8_t sfc[512];
    if (SF == SF_REPLICATED)
        for (int tid = 0; tid < 64; tid++) build_sf_compact<LOG2N>(a.sf, sfc, tid, 64);
    for (int lane = 0; lane < 32; lane++) {
        const int tb_l = lane / L::TPB, tl = lane % L::TPB;
        const unsigned char *in = in_buf + tb_l * L::TB_BYTES;
        unsigned char *g = g_buf + tb_l * L::TB_BYTES;
        const TbParams &q = t[lane];
        const uint8_t *sf1 = q.sf;
        if (SF == SF_REPLICATED && q.sf) {
            const p265_tu_desc &d = a.tus[lane_tb_index[lane]];
            sf1 = sfc + sf_matrix_id(LOG2N, d.c_idx, q.flags) * kSfcStride;
        }
        if constexpr (LOG2N >= 4) {  // the kernel runs both columns of a lane in lock step
            const int xa = slot_index_rt(N, tl, 0), xb = slot_index_rt(N, tl, 1);
            if (slow) stage1_pair<LOG2N, SF, true, 0>(in, g, xa, xb, tl, sf1, q.w, q.rnd, q.sh, q.lsh);
            else if (zr == 0) stage1_pair<LOG2N, SF, false, 0>(in, g, xa, xb, tl, sf1, q.w, q.rnd, q.sh, 0);
            else if (zr == 1) stage1_pair<LOG2N, SF, false, 1>(in, g, xa, xb, tl, sf1, q.w, q.rnd, q.sh, 0);
            else stage1_pair<LOG2N, SF, false, 2>(in, g, xa, xb, tl, sf1, q.w, q.rnd, q.sh, 0);
            continue;
        }
        for (int half = 0; half < 2; half++) {
            const int x = slot_index_rt(N, tl, half);
            const int dstf = q.flags & P265_TU_DST;
            if (slow) stage1_column<LOG2N, SF, true>(in, g, x, tl, half, sf1, q.w, q.rnd, q.sh, q.lsh, dstf);
            else stage1_column<LOG2N, SF, false>(in, g, x, tl, half, sf1, q.w, q.rnd, q.sh, 0, dstf);
        }
    }
    for (int lane = 0; lane < 32; lane++) {
        const int tb_l = lane / L::TPB, tl = lane % L::TPB;
        unsigned char *g = g_buf + tb_l * L::TB_BYTES;
        const TbParams &q = t[lane];
        if (!q.valid || (q.flags & (P265_TU_SKIP | P265_TU_BYPASS))) continue;
        for (int c = 0; c < 2; c++) {
            const int row = tl + c * L::TPB;
            if constexpr (LOG2N >= 4) {  // the kernel's path
                if (zc == 0) stage2_row_g<LOG2N, 0>(g, row, q.rnd2, q.sh2);
                else if (zc == 1) stage2_row_g<LOG2N, 1>(g, row, q.rnd2, q.sh2);
                else stage2_row_g<LOG2N, 2>(g, row, q.rnd2, q.sh2);
            }
            else stage2_row<LOG2N>(g, row, q.dst + (size_t)row * q.stride, q.rnd2, q.sh2, q.flags & P265_TU_DST);
        }
    }
    if constexpr (LOG2N >= 4) {  // coalesced copy-out of the result rows (run_bin)
        using M = OutMap<LOG2N>;
        for (int i = 0; i < M::ITERS; i++)
            for (int lane = 0; lane < 32; lane++) {
                const TbParams &q = t[lane];
                if (!q.valid || (q.flags & (P265_TU_SKIP | P265_TU_BYPASS))) continue;
                const uint4 v = out_chunk_load<LOG2N>(g_buf + (lane / L::TPB) * L::TB_BYTES, i, lane);
                std::memcpy(q.dst + (size_t)(i * M::RPI + M::row0(lane)) * q.stride + M::part(lane) * 8, &v, 16);
            }
    }
}

template <int SF>
static void run_small_item(const KernelArgs &a, int bin, int item) {
    // one lane = one TB (tb8_lane / tb4_lane), 32 TBs per item
    alignas(16) unsigned char tile[kWarpSmemBytes];
    std::memset(tile, 0xA5, sizeof tile);
    TbParams t[32];
    bool slow = false;
    for (int lane = 0; lane < 32; lane++) {
        const int local = item * 32 + lane;
        const bool valid = local < a.n_tb[bin];
        t[lane] = params_from_x(a, valid ? expand_desc(a, load_desc(a, a.first_tb[bin] + local, true)) : make_uint4(0, 0, 0, 0), valid, 5 - bin);
        slow |= valid && t[lane].lsh != 0;
    }
    for (int lane = 0; lane < 32; lane++) {
        const TbParams &q = t[lane];
        if (!q.valid) continue;
        if (bin == 2) {
            for (int r = 0; r < 8; r++) copy16_async(tile + tb8_chunk_off(lane, r), q.src + r * 8);
            if (slow) tb8_lane<SF, true>(q, tile, lane, q.sf);
            else tb8_lane<SF, false>(q, tile, lane, q.sf);
        } else {
            uint32_t w[8];
            std::memcpy(w, q.src, 32);
            if (slow) tb4_lane<SF, true>(q, w, q.sf);
            else tb4_lane<SF, false>(q, w, q.sf);
        }
    }
}

template <int LOG2N>
static void run_item(const KernelArgs &a, int item) {
    if (LOG2N <= 3) {
        const int bin = 5 - LOG2N;
        if (!a.sf) run_small_item<SF_NONE>(a, bin, item);
        else run_small_item<SF_GENERAL>(a, bin, item);
        return;
    }
    if (!a.sf) run_item_sf<LOG2N, SF_NONE>(a, item);
    else if (a.sf_replicated) run_item_sf<LOG2N, SF_REPLICATED>(a, item);
    else run_item_sf<LOG2N, SF_GENERAL>(a, item);
}

extern "C" int host_residual_batch(const p265_tu_desc *tus, const int32_t bin_counts[4], const int16_t *coeffs,
                                   const uint8_t *sf, int sf_replicated, const p265_pic_geom *g, int16_t *out) {
    KernelArgs a;
    a.tus = tus; a.xtus = nullptr; a.wait_prev = 0; a.zext = 1; a.coeffs = coeffs; a.sf = sf; a.out = out; a.sf_replicated = sf_replicated;
    for (int c = 0; c < 3; c++) a.plane_off[c] = g->plane_off[c];
    a.pic_stride = g->pic_stride;
    a.stride_y = g->stride_y; a.stride_c = g->stride_c;
    a.bit_depth_y = g->bit_depth_y; a.bit_depth_c = g->bit_depth_c;
    int first = 0, items = 0;
    for (int b = 0; b < 4; b++) {
        a.first_tb[b] = first; a.n_tb[b] = bin_counts[b];
        first += bin_counts[b];
        a.first_item[b] = items;
        const int per = tbs_per_item(b);
        items += (bin_counts[b] + per - 1) / per;
    }
    a.first_item[4] = items;
    for (int w = 0; w < items; w++) {
        if (w < a.first_item[1]) run_item<5>(a, w - a.first_item[0]);
        else if (w < a.first_item[2]) run_item<4>(a, w - a.first_item[1]);
        else if (w < a.first_item[3]) run_item<3>(a, w - a.first_item[2]);
        else run_item<2>(a, w - a.first_item[3]);
    }
    return 0;
}

extern "C" int host_slot_index(int n, int s, int h) { return slot_index_rt(n, s, h); }
