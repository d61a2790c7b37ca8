// Host emulation of one residual launch: runs the SAME per-lane phase functions the
// sm_100a kernel runs (p265_b200/csrc/residual_core.cuh), lane after lane, phase after
// phase -- the only synchronisation in the kernel is __syncwarp() between phases, so
// this is an exact functional model of the device code's indexing and arithmetic.
// Built and driven by tests/test_host_core.py; never part of the product library.
#include <cstring>
#include <vector>

#include "../../p265_b200/csrc/residual_core.cuh"

using namespace p265;

template <int LOG2N>
static void run_item(const KernelArgs &a, int item) {
    using L = Layout<LOG2N>;
    alignas(16) unsigned char smem[kWarpSmemBytes];
    std::memset(smem, 0xA5, sizeof smem);
    TbParams t[32];
    for (int lane = 0; lane < 32; lane++) {
        bool valid;
        int tb = lane_tb<LOG2N>(a, item, lane, valid);
        t[lane] = make_params(a, tb, valid);
    }
    for (int lane = 0; lane < 32; lane++) tile_issue<LOG2N>(lane, t[lane], smem);
    for (int lane = 0; lane < 32; lane++) phase_special<LOG2N>(lane, t[lane], smem);
    static int p[32][2][L::N / 2];
    bool slow = false;
    for (int lane = 0; lane < 32; lane++) slow |= t[lane].lsh != 0;
    const int sf = !a.sf ? SF_NONE : (a.sf_replicated ? SF_REPLICATED : SF_GENERAL);
    for (int lane = 0; lane < 32; lane++) {
        if (sf == SF_NONE && slow) phase_gather<LOG2N, SF_NONE, true>(lane, t[lane], smem, p[lane]);
        else if (sf == SF_NONE) phase_gather<LOG2N, SF_NONE, false>(lane, t[lane], smem, p[lane]);
        else if (sf == SF_GENERAL && slow) phase_gather<LOG2N, SF_GENERAL, true>(lane, t[lane], smem, p[lane]);
        else if (sf == SF_GENERAL) phase_gather<LOG2N, SF_GENERAL, false>(lane, t[lane], smem, p[lane]);
        else if (slow) phase_gather<LOG2N, SF_REPLICATED, true>(lane, t[lane], smem, p[lane]);
        else phase_gather<LOG2N, SF_REPLICATED, false>(lane, t[lane], smem, p[lane]);
    }
    for (int lane = 0; lane < 32; lane++) phase_stage1<LOG2N>(lane, t[lane], smem, p[lane]);
    for (int lane = 0; lane < 32; lane++) phase_stage2<LOG2N>(lane, t[lane], smem);
}

extern "C" int host_residual_batch(const p265_tu_desc *tus, const int32_t bin_counts[4], const int16_t *coeffs,
                                   const uint8_t *sf, int sf_replicated, const p265_pic_geom *g, int16_t *out) {
    KernelArgs a;
    a.tus = tus; a.coeffs = coeffs; a.sf = sf; a.out = out; a.sf_replicated = sf_replicated;
    for (int c = 0; c < 3; c++) a.plane_off[c] = g->plane_off[c];
    a.pic_stride = g->pic_stride;
    a.stride_y = g->stride_y; a.stride_c = g->stride_c;
    a.bit_depth_y = g->bit_depth_y; a.bit_depth_c = g->bit_depth_c;
    int first = 0, items = 0;
    for (int b = 0; b < 4; b++) {
        a.first_tb[b] = first; a.n_tb[b] = bin_counts[b];
        first += bin_counts[b];
        a.first_item[b] = items;
        const int per = 2 << b;  // 2, 4, 8, 16 TBs per warp item
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
