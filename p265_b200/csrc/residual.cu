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
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "internal.h"
#include "residual_core.cuh"

namespace p265 {

// Occupancy plan (per SM) and CTA shape per bin.  The warps of a bin never talk to each other, so the CTA is
// only a scheduling container -- but its shape matters at the seams between the bin kernels (programmatic
// dependent launch: CTAs of the next bin start in the slots this bin frees) and for the order in which an SM
// walks the item list (the warps of a CTA take consecutive items).  Measured on B200 (round 2,
// profiles/r2_cta_shapes.txt): alone, the 32x32 bin is 6.6 % faster as ONE 16-warp CTA per SM and the 16x16
// bin 3.8 % faster as one 32-warp CTA, but a single large CTA holds its slots until its last warp is done and
// the next bin starts late; in the chain of bins, with the launch order of launch_residual_sf, the best of
// some 40 combinations over three sweeps is
//   32x32  2 CTAs x  8 warps = 16 warps, 128 registers (lock-step two-column pass), 8.4 KB smem per warp
//   16x16  4 CTAs x  8 warps = 32 warps,  64 registers
//   8x8    3 CTAs x  8 warps = 24 warps,  80 registers
//   4x4    2 CTAs x 16 warps = 32 warps,  64 registers
// = 0.2246 -> 0.2172 ms per 16 4K pictures (config 3) against the round-1 shape (4 warps per CTA everywhere,
// bins launched by size); config 2 0.1186 -> 0.1193 ms per 32 1080p pictures (unchanged within the noise).
#ifndef P265_WARPS_PER_CTA
#define P265_WARPS_PER_CTA 4
#endif
#ifndef P265_CTAS_PER_SM
#define P265_CTAS_PER_SM 3   // 8x8 bin
#endif
#ifndef P265_CTAS_BIN0
#define P265_CTAS_BIN0 2
#endif
#ifndef P265_WARPS_BIN0
#define P265_WARPS_BIN0 8
#endif
#ifndef P265_WARPS_BIN1
#define P265_WARPS_BIN1 8
#endif
#ifndef P265_WARPS_BIN2
#define P265_WARPS_BIN2 8
#endif
#ifndef P265_WARPS_BIN3
#define P265_WARPS_BIN3 16
#endif
constexpr int kCtasPerSm = P265_CTAS_PER_SM;
constexpr int kDescRingBytes = 2 * 32 * 16;                          // 2 slots x 32 lanes x 16 B
constexpr int kWarpBytes = 2 * kWarpSmemBytes + kDescRingBytes;      // in + g + ring = 9472

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int KEEP>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(KEEP) : "memory");
}

// Out-of-line wrappers (see residual_core.cuh): one copy of each 1-D pass per size, shared
// by the two columns / rows a lane owns.  Inside a loop ptxas would hoist the ~90 packed
// basis constants of the 32-point butterfly and need > 180 registers; as functions each
// pass keeps a small allocation and fetches its constants just in time (LDCU).
// The hot wrappers take 32-bit SHARED-WINDOW addresses for the warp's buffers (and for the compact
// ScalingFactor entry): a generic pointer costs a 64-bit argument plus the window-base arithmetic
// (S2R / ULEA) on both sides of every call.
__device__ __forceinline__ unsigned char *smem_ptr(uint32_t s) {
    return static_cast<unsigned char *>(__cvta_shared_to_generic(s));
}
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// sf: SF_REPLICATED -> shared address of the TB's compact matrix; SF_GENERAL -> global pointer; SF_NONE -> unused
template <int LOG2N, int SF, bool SLOW, int Z>
__device__ __noinline__ void stage1_pair_call(uint32_t in_s, uint32_t g_s, int x0, int x1, int tl, uint64_t sf, int w,
                                              int rnd, int sh, int lsh) {
    const uint8_t *sfp = SF == SF_REPLICATED ? smem_ptr((uint32_t)sf)
                                             : reinterpret_cast<const uint8_t *>(static_cast<uintptr_t>(sf));
    stage1_pair<LOG2N, SF, SLOW, Z>(smem_ptr(in_s), smem_ptr(g_s), x0, x1, tl, sfp, w, rnd, sh, lsh);
}
// Stage 1 runs both columns of a lane in lock step (stage1_pair: -6 % on the 16x16 bin, -10 % on the
// 32x32 bin at 128 registers); the same form for the two rows of stage 2 was measured twice: no gain.
// Both rows of a lane in one call, one after the other (not unrolled: one copy of the pass).  The
// result rows go back into g (stage2_row_g) and leave through the coalesced copy-out of run_bin.
template <int LOG2N, int Z>
__device__ __noinline__ void stage2_call(uint32_t g_s, int row, int rnd2, int sh2) {
    unsigned char *g = smem_ptr(g_s);
#pragma unroll 1
    for (int r = 0; r < 2; r++) {
        stage2_row_g<LOG2N, Z>(g, row, rnd2, sh2);
        row += Layout<LOG2N>::TPB;
    }
}

// Zero-aware passes (the parser knows last_sig_coeff_x / y, tu.py:145-148): a TB whose
// coefficients all lie in rows < N >> zr and columns < N >> zc (codes in the expanded record, from the
// descriptor's rsvd bits or from unpack_kernel, which sees the significance bitmap) runs shortened
// column / row passes -- 172 / 88 / 46 IDP.2A per 32-point pass for code 0 / 1 / 2, and only the rows
// inside the extent are read and dequantised.  The choice is per work item and warp-uniform: the weakest
// promise among the item's TBs.  One out-of-line copy of each pass per code; a picture whose TBs all
// promise nothing (the benchmark's coefficient model: 99.85 % of its 32x32 TBs have a level in the last
// quarter of their rows) only ever touches the code-0 copies, so its instruction-cache footprint is
// unchanged.  -DP265_ZERO_EXTENT=0 compiles the dispatch out (A/B runs).
#ifndef P265_ZERO_EXTENT
#define P265_ZERO_EXTENT 1
#endif

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// Small bins with a dense arena: software prefetch of a later item's tile into L2, P265_SMALL_PREFETCH items
// ahead of the cp.async that fetches it (0 = off; the top stall of both small bins is the cp.async wait at the
// head of the item loop, profiles/r2s3 source view).
#ifndef P265_SMALL_PREFETCH
#define P265_SMALL_PREFETCH 0
#endif

// One size bin, one warp: items w, w + W, w + 2W, ... of the bin.  Software pipeline,
// everything asynchronous (cp.async / LDGSTS, no register staging):
//     descriptor of item k+2  ->  slot (k+2) % 3 of a ring in shared memory (one entry per TB)
//     tile of item k+1        ->  prefetched into L2 at the top of item k (one line per lane), copied into
//                                 `in` as soon as stage 1 of item k has consumed it (an L2 hit by then)
//     stage 2 of item k       ->  overlaps the tile copy; its result rows go back into g
//     copy-out of item k      ->  whole rows per store instruction (OutMap)
// so both dependent global-memory latencies of an item hide behind arithmetic.  The per-item
// set-up reads the expanded record's fields as they are (no parameter derivation in the loop).
template <int LOG2N, int SF>
__device__ __forceinline__ void run_bin(const KernelArgs &a, int gw, int stride, int lane, unsigned char *wbase,
                                        const uint8_t *sfc) {
    using L = Layout<LOG2N>;
    using M = OutMap<LOG2N>;
    constexpr int N = L::N, bin = 5 - LOG2N;
    const int n_items = a.first_item[bin + 1] - a.first_item[bin];
    if (gw >= n_items) return;
    const int n_tb = a.n_tb[bin];
    const uint4 *xt = a.xtus + a.first_tb[bin];
    unsigned char *in_base = wbase, *g_base = wbase + L::WARP_BYTES;
    const int tb_l = lane / L::TPB, tl = lane % L::TPB;
    constexpr int RS = L::TBS;  // ring slot s = ring0[RS * s .. RS * s + RS)
    uint4 *ring0 = reinterpret_cast<uint4 *>(wbase + 2 * L::WARP_BYTES);
    uint4 *ring = ring0 + tb_l;  // the lane's own TB
    const unsigned char *in = in_base + tb_l * L::TB_BYTES;
    unsigned char *g = g_base + tb_l * L::TB_BYTES;
    const uint32_t in_s = smem_addr(in), g_s = smem_addr(g), sfc_s = smem_addr(sfc);
    const int x0 = slot_index_rt(N, tl, 0), x1 = slot_index_rt(N, tl, 1);
    constexpr int kLines = N * N * 2 / 128;  // 128-byte lines per TB (<= lanes per TB)
    {   // prologue: first descriptor by plain load, its tile and the second descriptor async
        const int tb = gw * L::TBS + tb_l;
        const bool valid = tb < n_tb;
        const uint4 d0 = valid ? xt[tb] : make_uint4(0, 0, 0, 0);
        ring[0] = d0;
        tile_issue<LOG2N>(lane, a.coeffs + (size_t)d0.z * 16, valid, in_base);
        const int tb1 = (gw + stride) * L::TBS + tb_l;
        if (gw + stride < n_items && tb1 < n_tb && tl == 0) copy16_async(&ring[RS], &xt[tb1]);
        cp_async_commit();
    }
    const uint32_t out_row0 = (uint32_t)M::row0(lane), out_part8 = (uint32_t)M::part(lane) * 8;
    int k = 0, k1 = 1, k2 = 2;  // ring slots of items it, it + stride, it + 2 * stride
    for (int it = gw; it < n_items; it += stride) {
        cp_async_wait<0>();  // tile k and descriptor k+1 have landed
        __syncwarp();        // ... for every lane; also: all lanes are done with g of item k-1
        const bool valid = it * L::TBS + tb_l < n_tb;
        const bool more = it + stride < n_items;
        const bool v1 = more && (it + stride) * L::TBS + tb_l < n_tb;
        const uint4 d = ring[RS * k];
        if (v1 && tl < kLines)  // next tile -> L2 while this item is transformed
            prefetch_l2(a.coeffs + (size_t)ring[RS * k1].z * 16 + tl * 64);
        const int flags = valid ? (int)xd_flags(d) : 0;
        const int w = valid ? xd_w(d) : 0, sh = xd_sh(d), lsh = valid ? xd_lsh(d) : 0;
        const int rnd = (1 << sh) >> 1;
        // SF_REPLICATED: stage 1 reads the CTA's compact copy of this TB's matrix (matrixId from the record)
        // (always a valid shared-memory address -- no generic-pointer select: a lane without a TB runs
        // with w = 0 into its own unused part of g; matrixId 6 = prescaled TB = the all-ones entry)
        uint64_t sf1 = 0;
        if (SF == SF_REPLICATED) sf1 = sfc_s + (valid ? xd_mid(d) : 0u) * kSfcStride;
        else if (SF == SF_GENERAL && valid && !(flags & P265_TU_PRESCALED))
            sf1 = reinterpret_cast<uintptr_t>(a.sf + sf_matrix_offset(LOG2N, 0, 1) + (xd_mid(d) << (2 * LOG2N)));
        // rare, warp-uniform: transform-skip / bypass TBs, left-shift dequantisation
        const bool is_special = (flags & (P265_TU_SKIP | P265_TU_BYPASS)) != 0;
        bool slow = false;
        if (__any_sync(0xffffffffu, is_special || lsh != 0)) {
            slow = __any_sync(0xffffffffu, lsh != 0);
            if (__any_sync(0xffffffffu, is_special)) phase_special<LOG2N>(lane, params_from_x(a, d, valid, LOG2N), in_base);
        }
        if (!slow) stage1_pair_call<LOG2N, SF, false, 0>(in_s, g_s, x0, x1, tl, sf1, w, rnd, sh, 0);
        else stage1_pair_call<LOG2N, SF, true, 0>(in_s, g_s, x0, x1, tl, sf1, w, rnd, sh, lsh);  // rare
        __syncwarp();  // `in` is consumed, g is complete
        if (more) {
            tile_issue<LOG2N>(lane, a.coeffs + (size_t)ring[RS * k1].z * 16, v1, in_base);
            const int tb2 = (it + 2 * stride) * L::TBS + tb_l;
            if (it + 2 * stride < n_items && tb2 < n_tb && tl == 0) copy16_async(&ring[RS * k2], &xt[tb2]);
        }
        cp_async_commit();
        if (valid && !is_special) {
            const int sh2 = xd_sh2(d);
            stage2_call<LOG2N, 0>(g_s, tl, 1 << (sh2 - 1), sh2);
        }
        __syncwarp();  // every result row of the item sits in g
        // copy-out: store instruction i = rows 4i .. 4i+3 of every TB of the item; the lane stays
        // inside its own TB (32-bit element offsets: the launcher keeps a batch below 2^32 elements)
        if (valid && !is_special) {
            const uint32_t row_step = (uint32_t)xd_stride(d);
            const uint32_t e0 = d.x + out_row0 * row_step + out_part8;
#pragma unroll
            for (int i = 0; i < M::ITERS; i++)
                *reinterpret_cast<uint4 *>(a.out + (e0 + (uint32_t)(i * M::RPI) * row_step)) = out_chunk_load<LOG2N>(g, i, lane);
        }
        const int kk = k; k = k1; k1 = k2; k2 = kk;
    }
    cp_async_wait<0>();
    __syncwarp();
}

// The same bin with the zero-extent codes of the records honoured (KernelArgs.zext).  A separate copy of the
// loop, so that batches without codes -- the benchmark's coefficient model -- run exactly the code above.
// Differences:
//   * the item's code pair = the weakest promise among its TBs (records of special / left-shift TBs carry 0),
//     found one item ahead, when the next item's records have landed in the ring;
//   * with the code known that early, the next tile is requested at the TOP of the current item, into a second
//     `in` buffer (the launch reserves it), and only its first N >> zr rows are copied: a shortened item is
//     too short to hide a tile request that is issued only after its own column pass (measured: the 32x32 bin
//     with every TB at code (2,2) executed 44 % fewer instructions and ran 15 % faster, long-scoreboard
//     stalls 0.2 -> 2.0 warps per issue);
//   * column and row passes are chosen by code: 172 / 88 / 46 IDP.2A per 32-point pass.
template <int LOG2N, int SF>
__device__ __forceinline__ void run_bin_zext(const KernelArgs &a, int gw, int stride, int lane, unsigned char *wbase,
                                             int in1_off, const uint8_t *sfc) {
    using L = Layout<LOG2N>;
    using M = OutMap<LOG2N>;
    constexpr int N = L::N, bin = 5 - LOG2N;
    const int n_items = a.first_item[bin + 1] - a.first_item[bin];
    if (gw >= n_items) return;
    const int n_tb = a.n_tb[bin];
    const uint4 *xt = a.xtus + a.first_tb[bin];
    auto inb = [&](int which) { return wbase + (which ? in1_off : 0); };  // the two tile buffers of the warp
    unsigned char *g_base = wbase + L::WARP_BYTES;
    const int tb_l = lane / L::TPB, tl = lane % L::TPB;
    constexpr int RS = L::TBS;
    uint4 *ring0 = reinterpret_cast<uint4 *>(wbase + 2 * L::WARP_BYTES);
    uint4 *ring = ring0 + tb_l;
    unsigned char *g = g_base + tb_l * L::TB_BYTES;
    const uint32_t in_s0 = smem_addr(wbase + tb_l * L::TB_BYTES);
    const uint32_t g_s = smem_addr(g), sfc_s = smem_addr(sfc);
    const int x0 = slot_index_rt(N, tl, 0), x1 = slot_index_rt(N, tl, 1);
    auto item_codes = [&](const uint4 d, bool v, int &zr, int &zc) {  // lanes without a TB promise everything
        zr = __reduce_min_sync(0xffffffffu, v ? xd_zr(d) : 2);
        zc = __reduce_min_sync(0xffffffffu, v ? xd_zc(d) : 2);
    };
    int zr, zc;
    {
        const int tb = gw * L::TBS + tb_l;
        const bool valid = tb < n_tb;
        const uint4 d0 = valid ? xt[tb] : make_uint4(0, 0, 0, 0);
        ring[0] = d0;
        item_codes(d0, valid, zr, zc);
        tile_issue<LOG2N>(lane, a.coeffs + (size_t)d0.z * 16, valid, inb(0), N >> zr);
        const int tb1 = (gw + stride) * L::TBS + tb_l;
        if (gw + stride < n_items && tb1 < n_tb && tl == 0) copy16_async(&ring[RS], &xt[tb1]);
        cp_async_commit();
    }
    const uint32_t out_row0 = (uint32_t)M::row0(lane), out_part8 = (uint32_t)M::part(lane) * 8;
    int k = 0, k1 = 1, k2 = 2, par = 0;
    for (int it = gw; it < n_items; it += stride) {
        cp_async_wait<0>();  // tile k and descriptor k+1 have landed
        __syncwarp();        // ... for every lane; all lanes are done with g and with the other `in` buffer
        const bool valid = it * L::TBS + tb_l < n_tb;
        const bool more = it + stride < n_items;
        const bool v1 = more && (it + stride) * L::TBS + tb_l < n_tb;
        const uint4 d = ring[RS * k];
        int zr1 = 0, zc1 = 0;
        if (more) {  // the next item: its codes, then its tile (the rows its row code leaves) and the record after it
            const uint4 dn = ring[RS * k1];
            item_codes(dn, v1, zr1, zc1);
            tile_issue<LOG2N>(lane, a.coeffs + (size_t)dn.z * 16, v1, inb(par ^ 1), N >> zr1);
            const int tb2 = (it + 2 * stride) * L::TBS + tb_l;
            if (it + 2 * stride < n_items && tb2 < n_tb && tl == 0) copy16_async(&ring[RS * k2], &xt[tb2]);
        }
        cp_async_commit();
        const int flags = valid ? (int)xd_flags(d) : 0;
        const int w = valid ? xd_w(d) : 0, sh = xd_sh(d), lsh = valid ? xd_lsh(d) : 0;
        const int rnd = (1 << sh) >> 1;
        uint64_t sf1 = 0;
        if (SF == SF_REPLICATED) sf1 = sfc_s + (valid ? xd_mid(d) : 0u) * kSfcStride;
        else if (SF == SF_GENERAL && valid && !(flags & P265_TU_PRESCALED))
            sf1 = reinterpret_cast<uintptr_t>(a.sf + sf_matrix_offset(LOG2N, 0, 1) + (xd_mid(d) << (2 * LOG2N)));
        const bool is_special = (flags & (P265_TU_SKIP | P265_TU_BYPASS)) != 0;
        bool slow = false;
        if (__any_sync(0xffffffffu, is_special || lsh != 0)) {  // such an item has code (0, 0): the whole tile is there
            slow = __any_sync(0xffffffffu, lsh != 0);
            if (__any_sync(0xffffffffu, is_special)) phase_special<LOG2N>(lane, params_from_x(a, d, valid, LOG2N), inb(par));
        }
        const uint32_t in_cur = in_s0 + (par ? (uint32_t)in1_off : 0u);
        if (slow) stage1_pair_call<LOG2N, SF, true, 0>(in_cur, g_s, x0, x1, tl, sf1, w, rnd, sh, lsh);  // rare
        else if (zr == 0) stage1_pair_call<LOG2N, SF, false, 0>(in_cur, g_s, x0, x1, tl, sf1, w, rnd, sh, 0);
        else if (zr == 1) stage1_pair_call<LOG2N, SF, false, 1>(in_cur, g_s, x0, x1, tl, sf1, w, rnd, sh, 0);
        else stage1_pair_call<LOG2N, SF, false, 2>(in_cur, g_s, x0, x1, tl, sf1, w, rnd, sh, 0);
        __syncwarp();  // g is complete
        if (valid && !is_special) {
            const int sh2 = xd_sh2(d);
            if (zc == 0) stage2_call<LOG2N, 0>(g_s, tl, 1 << (sh2 - 1), sh2);
            else if (zc == 1) stage2_call<LOG2N, 1>(g_s, tl, 1 << (sh2 - 1), sh2);
            else stage2_call<LOG2N, 2>(g_s, tl, 1 << (sh2 - 1), sh2);
        }
        __syncwarp();  // every result row of the item sits in g
        if (valid && !is_special) {
            const uint32_t row_step = (uint32_t)xd_stride(d);
            const uint32_t e0 = d.x + out_row0 * row_step + out_part8;
#pragma unroll
            for (int i = 0; i < M::ITERS; i++)
                *reinterpret_cast<uint4 *>(a.out + (e0 + (uint32_t)(i * M::RPI) * row_step)) = out_chunk_load<LOG2N>(g, i, lane);
        }
        const int kk = k; k = k1; k1 = k2; k2 = kk;
        par ^= 1;
        zr = zr1;
        zc = zc1;
    }
    cp_async_wait<0>();
    __syncwarp();
}

// ---- small TBs: one lane = one TB (residual_core.cuh: tb8_lane / tb4_lane) -----------
// Whether the 8x8 / 4x4 bins read expanded descriptors too.  No instruction is saved there (one
// lane per TB), but with a ScalingFactor table the per-TB table pointer is cheaper from the
// expanded record and expand_kernel's pass leaves the descriptors in L2: measured on the 4K
// 10-bit mix, +3 % with a table, -3 % without -> chosen by the SF mode.
template <int SF, int LOG2N>
struct SmallDesc {
    static constexpr bool X = SF != SF_NONE;
    static __device__ __forceinline__ uint4 load(const KernelArgs &a, int i, bool v) {
        return X ? load_xdesc(a, i, v) : load_desc(a, i, v);
    }
    static __device__ __forceinline__ const uint4 *ptr(const KernelArgs &a, int i) {
        return X ? &a.xtus[i] : reinterpret_cast<const uint4 *>(&a.tus[i]);
    }
    static __device__ __forceinline__ TbParams params(const KernelArgs &a, const uint4 d, bool v) {
        return X ? params_from_x(a, d, v, LOG2N) : make_params(a, d, v);
    }
};
// Small bins with a ScalingFactor table: the CTA keeps the 6 matrices of its size ([y][x] bytes,
// as in the table) plus an all-ones matrix (matrixId 6 of the expanded record = prescaled TB,
// m = 1) at the start of its shared memory, so a TB's factors come from an LDS issued next to
// its tile reads instead of a dependent global load at the head of every item.
template <int LOG2N>
__device__ __forceinline__ void build_sf_small(const uint8_t *table, uint8_t *out, int tid, int nthreads) {
    constexpr int NN = 1 << (2 * LOG2N);
    const uint8_t *base = table + sf_matrix_offset(LOG2N, 0, 1);
    for (int i = tid; i < 7 * NN; i += nthreads) out[i] = i < 6 * NN ? base[i] : (uint8_t)1;
}
template <int SF, int LOG2N>
__device__ __forceinline__ const uint8_t *small_sf_ptr(const uint8_t *sfs, const uint4 d) {
    return SF != SF_NONE ? sfs + (xd_mid(d) << (2 * LOG2N)) : nullptr;
}

template <int SF, bool SLOW>
__device__ __noinline__ void tb8_call(const KernelArgs &a, const uint4 d, bool valid, uint32_t tile_s, int lane,
                                      uint32_t sfs_s) {  // shared-window addresses, like the big-bin wrappers
    const TbParams t = SmallDesc<SF, 3>::params(a, d, valid);
    tb8_lane<SF, SLOW>(t, smem_ptr(tile_s), lane, small_sf_ptr<SF, 3>(smem_ptr(sfs_s), d));
}

// 8x8 bin: 32 TBs per item; the lane's 128-byte TB is copied asynchronously into a
// lane-private, chunk-swizzled slot of one of the warp's two tile buffers while the
// previous item is being transformed; descriptors run two items ahead in the ring.
template <int SF>
__device__ __forceinline__ void run_bin8(const KernelArgs &a, int gw, int stride, int lane, unsigned char *wbase,
                                         const uint8_t *sfs) {
    const int n_tb = a.n_tb[2], first = a.first_tb[2];
    const int n_items = (n_tb + 31) >> 5;
    if (gw >= n_items) return;
    const uint32_t wbase_s = smem_addr(wbase), sfs_s = smem_addr(sfs);
    uint4 *ring0 = reinterpret_cast<uint4 *>(wbase + 2 * kWarpSmemBytes);  // slot s, TB t at ring0[32 * s + t]
    uint4 *ring = ring0 + lane;
    // Cooperative tile copy: copy instruction i moves the 8 chunks (rows) of TBs 4i .. 4i+3, i.e. four
    // whole 128-byte lines when the arena is dense, instead of one 16-byte chunk of 32 different TBs
    // (32 LSU wavefronts per instruction).  The TB's arena offset comes from its ring entry.
    // Dense arena (P265_RES_DENSE_ARENA): TB i of the bin sits 64 coefficients behind TB i-1, so a
    // tile's address follows from the bin's first offset and the copy does not wait for descriptors:
    // a warp sees only a handful of items, its two-deep start-up chain (descriptor -> tile) was a
    // large part of its lifetime.
    const bool dense = a.dense_arena != 0;
    const uint32_t z0 = dense ? a.tus[first].coeff_off : 0u;
    auto issue = [&](int slot, int item, unsigned char *tile) {
        const int n_here = n_tb - item * 32;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int t = (lane >> 3) + 4 * i;
            if (t < n_here) {
                const uint32_t z = dense ? z0 + (uint32_t)(item * 32 + t) * 4u : ring0[32 * slot + t].z;
                copy16_async(tile + tb8_chunk_off(t, lane & 7), a.coeffs + (size_t)z * 16 + (lane & 7) * 8);
            }
        }
    };
    bool valid = gw * 32 + lane < n_tb;
    {
        if (dense) issue(0, gw, wbase);
        const uint4 d0 = SmallDesc<SF, 3>::load(a, first + gw * 32 + lane, valid);
        ring[0] = d0;
        __syncwarp();
        if (!dense) issue(0, gw, wbase);
        const int i1 = (gw + stride) * 32 + lane;
        if (gw + stride < n_items && i1 < n_tb) copy16_async(&ring[32], SmallDesc<SF, 3>::ptr(a, first + i1));
        cp_async_commit();
    }
    int k = 0;
    for (int it = gw; it < n_items; it += stride, k ^= 1) {
        if (P265_SMALL_PREFETCH > 0 && dense) {  // a later item's tile -> L2 (one 128-byte line = one TB per lane)
            const int pit = it + P265_SMALL_PREFETCH * stride;
            if (pit * 32 + lane < n_tb) prefetch_l2(a.coeffs + (size_t)(z0 + (uint32_t)(pit * 32 + lane) * 4u) * 16);
        }
        cp_async_wait<0>();  // tile k and descriptor k+1 ...
        __syncwarp();        // ... of every lane; all lanes are done with the other tile buffer
        valid = it * 32 + lane < n_tb;
        const uint4 d_cur = ring[32 * k];
        if (it + stride < n_items) {
            issue(k ^ 1, it + stride, wbase + (k ^ 1) * kWarpSmemBytes);
            const int i2 = (it + 2 * stride) * 32 + lane;
            if (it + 2 * stride < n_items && i2 < n_tb) copy16_async(&ring[32 * k], SmallDesc<SF, 3>::ptr(a, first + i2));
        }
        cp_async_commit();
        // per >= bdShift needs qP >= 6 * (bitDepth - 2): impossible for 8x8 below 14 bits,
        // but the descriptor is caller data: decide warp-uniformly like the other sizes
        bool slow_lane;
        if (SmallDesc<SF, 3>::X) {
            slow_lane = xd_lsh(d_cur) != 0;
        } else {
            const int qp = (int)((d_cur.y >> 16) & 0xff), c_idx = (int)((d_cur.y >> 8) & 0xff);
            slow_lane = ((qp * 43) >> 8) >= (c_idx ? a.bit_depth_c : a.bit_depth_y) - 2;
        }
        const bool slow = __any_sync(0xffffffffu, valid && slow_lane);
        if (slow) tb8_call<SF, true>(a, d_cur, valid, wbase_s + k * kWarpSmemBytes, lane, sfs_s);
        else tb8_call<SF, false>(a, d_cur, valid, wbase_s + k * kWarpSmemBytes, lane, sfs_s);
    }
    cp_async_wait<0>();
}

// 4x4 bin: 32 TBs per item, one per lane.  Three-stage cp.async pipeline in lane-private
// shared memory: the 32 bytes of coefficients of item k+2 and the descriptor of item k+3
// are in flight while item k is transformed in registers (the bin is latency-bound: every
// TB is only 32 + 16 bytes of input).
constexpr int kBin4Stages = 3;
constexpr int kBin4WarpBytes = kBin4Stages * 32 * 32 + 4 * 32 * 16;  // tiles + 4-slot descriptor ring = 5120

__device__ __forceinline__ int tb4_slot_off(int lane, int half) {
    // 32-byte lane slots; the two 16-byte halves are swapped for odd lane quads so that a
    // quarter-warp's 128-bit reads hit distinct banks
    return lane * 32 + ((half ^ ((lane >> 2) & 1)) << 4);
}

template <int SF>
__device__ __forceinline__ void run_bin4(const KernelArgs &a, int gw, int stride, int lane, unsigned char *wbase,
                                         const uint8_t *sfs) {
    const int n_tb = a.n_tb[3], first = a.first_tb[3];
    const int n_items = (n_tb + 31) >> 5;
    if (gw >= n_items) return;
    unsigned char *tiles = wbase;                                                     // [stage][lane][32 B]
    uint4 *ring = reinterpret_cast<uint4 *>(wbase + kBin4Stages * 1024) + lane;      // slot s at ring[32 * s]
    auto desc_async = [&](int item, int slot) {
        const int i = item * 32 + lane;
        if (item < n_items && i < n_tb) copy16_async(&ring[32 * slot], SmallDesc<SF, 2>::ptr(a, first + i));
    };
    const bool dense = a.dense_arena != 0;  // see run_bin8: a 4x4 TB is one 16-coefficient unit
    const uint32_t z0 = dense ? a.tus[first].coeff_off : 0u;
    auto tile_async = [&](int item, const uint4 d, int stage) {
        const int i = item * 32 + lane;
        if (item < n_items && i < n_tb) {
            const int16_t *src = a.coeffs + (size_t)(dense ? z0 + (uint32_t)i : d.z) * 16;
            unsigned char *dst = tiles + stage * 1024;
            copy16_async(dst + tb4_slot_off(lane, 0), src);
            copy16_async(dst + tb4_slot_off(lane, 1), src + 8);
        }
    };
    // cp.async groups: G0 = {tile 0}, G_j = {tile j, descriptor j+2} for j >= 1.  At item k
    // all groups but the newest (G_{k+1}) are complete: tile k and descriptors <= k+2.
    {   // prologue: descriptors 0..2 by plain loads (paid once per warp and bin); with a dense arena
        // the first two tiles are requested before them
        const uint4 none = make_uint4(0, 0, 0, 0);
        if (dense) {
            tile_async(gw, none, 0);
            cp_async_commit();
            tile_async(gw + stride, none, 1);
            desc_async(gw + 3 * stride, 3);
            cp_async_commit();
        }
        for (int j = 0; j < 3; j++) {
            const int i = (gw + j * stride) * 32 + lane;
            ring[32 * j] = SmallDesc<SF, 2>::load(a, first + i, gw + j * stride < n_items && i < n_tb);
        }
        if (!dense) {
            tile_async(gw, ring[0], 0);
            cp_async_commit();
            tile_async(gw + stride, ring[32], 1);
            desc_async(gw + 3 * stride, 3);
            cp_async_commit();
        }
    }
    int k = 0;  // item counter: tile stage k % 3, descriptor slot k % 4
#pragma unroll 1
    for (int it = gw; it < n_items; it += stride, k++) {
        if (P265_SMALL_PREFETCH > 0 && dense && lane < 8) {  // a later item's tile -> L2 (8 lines of 4 TBs)
            const int pit = it + (P265_SMALL_PREFETCH + 2) * stride;
            if (pit * 32 + lane * 4 < n_tb) prefetch_l2(a.coeffs + (size_t)(z0 + (uint32_t)(pit * 32 + lane * 4)) * 16);
        }
        cp_async_wait<1>();
        const bool valid = it * 32 + lane < n_tb;
        const int st = k % kBin4Stages;
        const uint4 d_cur = ring[32 * (k & 3)];  // read before its slot is refilled below
        tile_async(it + 2 * stride, ring[32 * ((k + 2) & 3)], (k + 2) % kBin4Stages);
        desc_async(it + 4 * stride, k & 3);
        cp_async_commit();
        const TbParams t = SmallDesc<SF, 2>::params(a, d_cur, valid);
        uint32_t w[8];
        {
            const uint4 v0 = *reinterpret_cast<const uint4 *>(tiles + st * 1024 + tb4_slot_off(lane, 0));
            const uint4 v1 = *reinterpret_cast<const uint4 *>(tiles + st * 1024 + tb4_slot_off(lane, 1));
            w[0] = v0.x; w[1] = v0.y; w[2] = v0.z; w[3] = v0.w;
            w[4] = v1.x; w[5] = v1.y; w[6] = v1.z; w[7] = v1.w;
        }
        const bool slow = __any_sync(0xffffffffu, t.lsh != 0);
        const uint8_t *sfm = small_sf_ptr<SF, 2>(sfs, d_cur);
        if (slow) tb4_lane<SF, true>(t, w, sfm);
        else tb4_lane<SF, false>(t, w, sfm);
    }
    cp_async_wait<0>();
}

// ---- small TBs, streaming form: one warp = ONE item, no software pipeline ---------------
// The persistent pipelines above hide the two global latencies of an item behind the previous
// item's arithmetic -- but a warp of a small bin only ever sees a handful of items (the 4x4 bin of
// 16 4K pictures is 8 items per warp), so its lifetime is mostly pipeline fill: the ncu source view
// of the round-1 4x4 bin shows 18 % of all stall samples on the cp.async wait at the head of the
// item loop and an achieved occupancy of 34 %.  Here the hardware does the hiding instead, as in the
// SAO kernel: a grid of one-item warps, every load of an item issued up front (descriptor, both
// halves of the tile; with a dense arena the tile address does not depend on the descriptor), 32
// warps per SM in flight, the block scheduler balancing the tail.  Measured on 16 4K pictures (B200,
// profiles/r2_small_bins.txt): alone, the 8x8 bin gains (46.2 -> 42.5 us) and the 4x4 bin does not
// (40.7 -> 42.3 us: its items are too short to amortise the per-CTA set-up); in the mix BOTH lose
// (0.2246 -> 0.2330 ms): a multi-wave grid releases its successor (griddepcontrol.launch_dependents)
// only when its last CTA has started, so 10 us of the tail overlap between the bins is gone.  The
// pipelined form therefore stays the default; P265_STREAM_BIN8 / P265_STREAM_BIN4 (compile time)
// select the streaming form per bin for A/B runs.
#ifndef P265_STREAM_BIN8
#define P265_STREAM_BIN8 0
#endif
#ifndef P265_STREAM_BIN4
#define P265_STREAM_BIN4 0
#endif
template <int BIN>
struct SmallStream {
    static constexpr bool on = BIN == 2 ? (P265_STREAM_BIN8 != 0) : (BIN == 3 ? (P265_STREAM_BIN4 != 0) : false);
};

template <int SF>
__device__ __forceinline__ void stream_bin4(const KernelArgs &a, int item, int lane, const uint8_t *sfs) {
    const int n_tb = a.n_tb[3], first = a.first_tb[3];
    const int i = item * 32 + lane;
    const bool valid = i < n_tb;
    uint4 v0 = make_uint4(0, 0, 0, 0), v1 = v0;
    uint4 d;
    if (a.dense_arena) {  // tile address from the TB's index: all three loads of the lane are independent
        const uint32_t z0 = __ldg(&a.tus[first].coeff_off);
        d = SmallDesc<SF, 2>::load(a, first + i, valid);
        if (valid) {
            const uint4 *src = reinterpret_cast<const uint4 *>(a.coeffs + (size_t)(z0 + (uint32_t)i) * 16);
            v0 = __ldg(src);
            v1 = __ldg(src + 1);
        }
    } else {
        d = SmallDesc<SF, 2>::load(a, first + i, valid);
        if (valid) {
            const uint4 *src = reinterpret_cast<const uint4 *>(a.coeffs + (size_t)d.z * 16);
            v0 = __ldg(src);
            v1 = __ldg(src + 1);
        }
    }
    const TbParams t = SmallDesc<SF, 2>::params(a, d, valid);
    const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    const bool slow = __any_sync(0xffffffffu, t.lsh != 0);
    const uint8_t *sfm = small_sf_ptr<SF, 2>(sfs, d);
    if (slow) tb4_lane<SF, true>(t, w, sfm);
    else tb4_lane<SF, false>(t, w, sfm);
}

template <int SF>
__device__ __forceinline__ void stream_bin8(const KernelArgs &a, int item, int lane, unsigned char *wbase,
                                            const uint8_t *sfs) {
    const int n_tb = a.n_tb[2], first = a.first_tb[2];
    const int n_here = n_tb - item * 32;
    const bool valid = lane < n_here;
    const uint32_t wbase_s = smem_addr(wbase), sfs_s = smem_addr(sfs);
    uint4 d;
    // cooperative tile copy (see run_bin8): copy instruction i = the 8 chunks of TBs 4i .. 4i+3
    if (a.dense_arena) {
        const uint32_t z0 = __ldg(&a.tus[first].coeff_off);
        d = SmallDesc<SF, 3>::load(a, first + item * 32 + lane, valid);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int t = (lane >> 3) + 4 * i;
            if (t < n_here)
                copy16_async(wbase + tb8_chunk_off(t, lane & 7),
                             a.coeffs + (size_t)(z0 + (uint32_t)(item * 32 + t) * 4u) * 16 + (lane & 7) * 8);
        }
    } else {
        d = SmallDesc<SF, 3>::load(a, first + item * 32 + lane, valid);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int t = (lane >> 3) + 4 * i;
            const uint32_t z = __shfl_sync(0xffffffffu, d.z, t);
            if (t < n_here) copy16_async(wbase + tb8_chunk_off(t, lane & 7), a.coeffs + (size_t)z * 16 + (lane & 7) * 8);
        }
    }
    cp_async_commit();
    bool slow_lane;
    if (SmallDesc<SF, 3>::X) {
        slow_lane = xd_lsh(d) != 0;
    } else {
        const int qp = (int)((d.y >> 16) & 0xff), c_idx = (int)((d.y >> 8) & 0xff);
        slow_lane = ((qp * 43) >> 8) >= (c_idx ? a.bit_depth_c : a.bit_depth_y) - 2;
    }
    const bool slow = __any_sync(0xffffffffu, valid && slow_lane);
    cp_async_wait<0>();
    __syncwarp();
    if (slow) tb8_call<SF, true>(a, d, valid, wbase_s, lane, sfs_s);
    else tb8_call<SF, false>(a, d, valid, wbase_s, lane, sfs_s);
}

// CTAs per SM per bin: 32x32 is shared-memory limited (9.25 KB per warp); 16x16 needs only
// 5.25 KB per warp and fits 64 registers; 8x8 keeps 64 packed words live per lane; 4x4 is
// register-only.
#ifndef P265_CTAS_BIN1
#define P265_CTAS_BIN1 4
#endif
#ifndef P265_CTAS_BIN3
#define P265_CTAS_BIN3 2
#endif
constexpr int kSfcBytes = 640;  // compact ScalingFactor copy at the start of a CTA's shared memory (7 x 80 B)
template <int BIN>
struct BinCfg {
    static constexpr int ctas = BIN == 0 ? P265_CTAS_BIN0 : (BIN == 1 ? P265_CTAS_BIN1 : (BIN == 3 ? P265_CTAS_BIN3 : kCtasPerSm));
    // warps per CTA (P265_WARPS_BIN0..3 override P265_WARPS_PER_CTA per bin; ctas x warps = resident warps per SM)
    static constexpr int warps = BIN == 0 ? P265_WARPS_BIN0 : (BIN == 1 ? P265_WARPS_BIN1 : (BIN == 2 ? P265_WARPS_BIN2 : P265_WARPS_BIN3));
    // per warp: tile + g buffers + descriptor ring (2 slots x TBs per item x 16 B for the big sizes)
    static constexpr int smem =
        BIN == 3 ? kSfcBytes + (SmallStream<3>::on ? 0 : P265_WARPS_BIN3 * kBin4WarpBytes)
        : BIN == 0 ? kSfcBytes + P265_WARPS_BIN0 * (2 * Layout<5>::WARP_BYTES + 3 * Layout<5>::TBS * 16 + 32)
        : BIN == 1 ? kSfcBytes + P265_WARPS_BIN1 * (2 * Layout<4>::WARP_BYTES + 3 * Layout<4>::TBS * 16)
                   : kSfcBytes + P265_WARPS_BIN2 * (SmallStream<2>::on ? kWarpSmemBytes : kWarpBytes);
    // with zero-extent codes (run_bin_zext): a second tile buffer per warp of the big bins, behind the ring
    static constexpr int zext_extra = BIN == 0 ? Layout<5>::WARP_BYTES : (BIN == 1 ? Layout<4>::WARP_BYTES : 0);
};

template <int BIN, int SF>
__global__ void __launch_bounds__(BinCfg<BIN>::warps * 32, BinCfg<BIN>::ctas) residual_kernel(const __grid_constant__ KernelArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    // The four bin kernels are independent (disjoint TBs): let the next one be scheduled
    // as soon as CTAs of this one retire (programmatic dependent launch) so the tail of a
    // bin overlaps the head of the next instead of idling SMs.
    // The first bin kernel of a batch is launched on top of expand_kernel and must see its
    // output: it waits for that grid before letting its own successor in, so every later bin
    // is ordered behind expand_kernel as well.
    if (a.wait_prev) asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kWarps = BinCfg<BIN>::warps;
    const int stride = gridDim.x * kWarps;
    const int gw = blockIdx.x * kWarps + warp;
    constexpr int warp_bytes0 = (BinCfg<BIN>::smem - kSfcBytes) / kWarps;
    const bool zext = P265_ZERO_EXTENT && BIN <= 1 && a.zext;
    const int warp_bytes = warp_bytes0 + (zext ? BinCfg<BIN>::zext_extra : 0);
    unsigned char *wbase = smem + kSfcBytes + warp * warp_bytes;
    if (SF == SF_REPLICATED && BIN <= 1) {
        if (BIN == 0) build_sf_compact<5>(a.sf, smem, threadIdx.x, blockDim.x);
        else build_sf_compact<4>(a.sf, smem, threadIdx.x, blockDim.x);
        for (int i = threadIdx.x; i < kSfcStride; i += blockDim.x) smem[6 * kSfcStride + i] = 1;  // matrixId 6: m = 1
        __syncthreads();  // the only block-wide barrier: once per persistent CTA
    }
    if (SF != SF_NONE && BIN >= 2) {
        if (BIN == 2) build_sf_small<3>(a.sf, smem, threadIdx.x, blockDim.x);
        else build_sf_small<2>(a.sf, smem, threadIdx.x, blockDim.x);
        __syncthreads();
    }
    if (BIN == 0) {
        if (zext) run_bin_zext<5, SF>(a, gw, stride, lane, wbase, warp_bytes0, smem);
        else run_bin<5, SF>(a, gw, stride, lane, wbase, smem);
    } else if (BIN == 1) {
        if (zext) run_bin_zext<4, SF>(a, gw, stride, lane, wbase, warp_bytes0, smem);
        else run_bin<4, SF>(a, gw, stride, lane, wbase, smem);
    }
    else if (SmallStream<BIN>::on) {  // one item per warp; the grid covers the bin
        if (gw < a.first_item[BIN + 1] - a.first_item[BIN]) {
            if (BIN == 2) stream_bin8<SF>(a, gw, lane, wbase, smem);
            else stream_bin4<SF>(a, gw, lane, smem);
        }
    } else if (BIN == 2) run_bin8<SF>(a, gw, stride, lane, wbase, smem);
    else run_bin4<SF>(a, gw, stride, lane, wbase, smem);
    // Make the chain transitive: a bin launched with programmatic stream serialization may start
    // (and finish) before its predecessor has finished, so "the last bin is complete" would not
    // imply "every bin is complete" for whatever follows on the stream (the D2H copy of the planes,
    // SAO, the next call's expand_kernel overwriting the expanded records).  Every bin therefore
    // waits for its predecessor grid before it exits -- after its own work, so the overlap is kept
    // (a no-op for a grid launched without the attribute; the first bin already waited above).
    if (!a.wait_prev) asm volatile("griddepcontrol.wait;" ::: "memory");
}

// One thread per TB: public descriptor -> expanded record (residual_core.cuh: expand_desc).
__global__ void __launch_bounds__(256) expand_kernel(const __grid_constant__ KernelArgs a, int n_tus, uint4 *out) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // the first bin kernel may be scheduled early
    // four records per thread, all four loads in flight before the first is used
    const int i0 = blockIdx.x * (blockDim.x * 4) + threadIdx.x;
    uint4 d[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int i = i0 + j * blockDim.x;
        d[j] = i < n_tus ? load_desc(a, i, true) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int i = i0 + j * blockDim.x;
        if (i < n_tus) out[i] = expand_desc(a, d[j]);
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
    a.xtus = nullptr;
    a.wait_prev = 0;
    a.coeffs = d_coeffs;
    a.sf = d_sf;
    a.sf_replicated = 0;
    a.dense_arena = 0;
    a.zext = 0;
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
        const int per = tbs_per_item(b);
        items += (bin_counts[b] + per - 1) / per;
    }
    if (first > INT32_MAX || items > INT32_MAX) return set_error(P265_EINVAL, "too many TBs in one batch");
    a.first_item[4] = (int32_t)items;
    return P265_OK;
}

// plain = first kernel of a chain on the auxiliary stream: ordered behind expand_kernel by an event, no
// programmatic launch, nothing to wait for inside the kernel.  pct_limit (1..100, 0 = the P265_GRID_PCT
// knob): share of the occupancy-maximal persistent grid to launch (two chains sharing the SMs).
template <int BIN, int SF>
static int launch_bin(p265_ctx *ctx, KernelArgs a, bool first, cudaStream_t stream = nullptr, bool plain = false,
                      int pct_limit = 0) {
    // behind expand_kernel (first big-size bin: wait for it) or overlapping the previous bin
    const bool expanded = a.n_tb[0] + a.n_tb[1] + (a.sf ? a.n_tb[2] + a.n_tb[3] : 0) > 0;
    const bool overlap_previous = !plain && (!first || expanded);
    a.wait_prev = (!plain && first && expanded) ? 1 : 0;
    const int items = a.first_item[BIN + 1] - a.first_item[BIN];
    if (items == 0) return P265_OK;
    const bool zext = P265_ZERO_EXTENT && BIN <= 1 && a.zext;
    constexpr int kWarps = BinCfg<BIN>::warps;
    const int smem = BinCfg<BIN>::smem + (zext ? kWarps * BinCfg<BIN>::zext_extra : 0);
    // CTAs per SM the kernel really gets; function attributes are per device, so the
    // carve-out hint is set once for every device this process uses
    static int occ_dev[2][64] = {{0}};
    int &occ = occ_dev[zext ? 1 : 0][ctx->device & 63];
    if (!occ) {
        int o = 0;
        P265_CUDA(cudaFuncSetAttribute(residual_kernel<BIN, SF>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       cudaSharedmemCarveoutMaxShared));
        P265_CUDA(cudaFuncSetAttribute(residual_kernel<BIN, SF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       BinCfg<BIN>::smem + kWarps * BinCfg<BIN>::zext_extra));
        P265_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, residual_kernel<BIN, SF>, kWarps * 32,
                                                                smem));
        occ = o < 1 ? 1 : o;
    }
    // persistent grid, trimmed so that every warp gets the same number of items
    // tuning knobs: P265_GRID_PCT (all bins) / P265_GRID_PCT_BINS="a,b,c,d" (per bin 32,16,8,4): size of the
    // persistent grid in per cent of the CTAs that are resident at once.  < 100: kernels sharing the SMs;
    // > 100: a short second wave, so that the CTAs of the NEXT bin (programmatic dependent launch) start one by
    // one in the slots the first wave frees instead of all at once when an equal-work single wave ends
    static int pct = -1, pct_bin[4] = {0, 0, 0, 0};
    if (pct < 0) {
        const char *e = getenv("P265_GRID_PCT");
        pct = e ? atoi(e) : 100;
        if (pct < 1 || pct > 400) pct = 100;
        const char *b = getenv("P265_GRID_PCT_BINS");
        int v[4];
        if (b && sscanf(b, "%d,%d,%d,%d", &v[0], &v[1], &v[2], &v[3]) == 4)
            for (int i = 0; i < 4; i++) pct_bin[i] = (v[i] >= 1 && v[i] <= 400) ? v[i] : 0;
    }
    const int pct_use = pct_limit > 0 ? pct_limit : (pct_bin[BIN] ? pct_bin[BIN] : pct);
    // CTAs over the whole device (not per SM): a 70 % grid of a 4-CTA kernel is 2.8 CTAs per SM on average
    int max_ctas = (int)((int64_t)ctx->sm_count * occ * pct_use / 100);
    if (max_ctas < 1) max_ctas = 1;
    const int max_warps = max_ctas * kWarps;
    const int rounds = (items + max_warps - 1) / max_warps;
    const int warps = SmallStream<BIN>::on ? items : (items + rounds - 1) / rounds;  // streaming bins: one warp per item
    const int grid = (warps + kWarps - 1) / kWarps;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kWarps * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream ? stream : ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = overlap_previous ? 1 : 0;
    P265_CUDA(cudaLaunchKernelEx(&cfg, residual_kernel<BIN, SF>, a));
    ctx->launches++;
    return P265_OK;
}

template <int SF>
static int launch_residual_sf(p265_ctx *ctx, const KernelArgs &a) {
    // the first kernel of the batch is ordered normally behind whatever precedes it on the
    // stream (zero fill, copies); the following bins may overlap their predecessor
    // the bins are independent, so their launch order is free (tuning knob P265_BIN_ORDER, e.g. "0312")
    // Default: the bin with the least work first, the others by size (32, 16, 8, 4).  Measured on B200 with the
    // round-2 CTA shapes, 20 of the 24 orders (profiles/r2_cta_shapes.txt): the 4K 10-bit mix,
    // whose lightest bin is 4x4, runs 1.3 % faster as 4,32,16,8 (1.7 % on the benchmark's residual + SAO step) and
    // 9 % SLOWER with the 1080p 8-bit mix, whose lightest bin is 32x32 and whose best orders start with it.
    // Work per TB in ns (per-bin times of both mixes alone): 1.2 / 0.28 / 0.095 / 0.035.
    static int env_order[4] = {-1, 0, 0, 0};
    if (env_order[0] < 0) {
        const char *e = getenv("P265_BIN_ORDER");
        int o[4] = {4, 4, 4, 4};   // 4 = not given
        if (e && strlen(e) == 4) {
            int seen = 0;
            for (int i = 0; i < 4; i++) { o[i] = e[i] - '0'; if (o[i] >= 0 && o[i] < 4) seen |= 1 << o[i]; }
            if (seen != 15) o[0] = o[1] = o[2] = o[3] = 4;
        }
        env_order[1] = o[1]; env_order[2] = o[2]; env_order[3] = o[3]; env_order[0] = o[0];
    }
    int order[4] = {env_order[0], env_order[1], env_order[2], env_order[3]};
    if (order[0] > 3) {
        const double ns_per_tb[4] = {1.2, 0.28, 0.095, 0.035};
        int first = -1;
        double least = 0;
        for (int b = 0; b < 4; b++) {
            if (!a.n_tb[b]) continue;
            const double w = ns_per_tb[b] * a.n_tb[b];
            if (first < 0 || w < least) { first = b; least = w; }
        }
        if (first < 0) first = 0;
        int n = 0;
        order[n++] = first;
        for (int b = 0; b < 4; b++) if (b != first) order[n++] = b;
    }
    int rc = P265_OK;
    // Two chains sharing the SMs (tuning knob P265_SPLIT, e.g. "03|12": bins 32x32 + 4x4 on the context's
    // stream, 16x16 + 8x8 on an auxiliary stream, forked and joined by events; P265_SPLIT_PCT "a,b" = share
    // of each chain's persistent grids).  The issue-bound big bins and the latency-bound small ones then
    // run side by side instead of one after the other.
    static int split_a = -1, split_b = 0, pct_a = 70, pct_b = 70;
    if (split_a < 0) {
        split_a = 0;
        const char *e = getenv("P265_SPLIT");
        if (e && strlen(e) == 5 && e[2] == '|') {
            int seen = 0, ma = 0, mb = 0;
            for (int i = 0; i < 5; i++) {
                if (i == 2) continue;
                const int b = e[i] - '0';
                if (b < 0 || b > 3) { seen = 0; break; }
                seen |= 1 << b;
                (i < 2 ? ma : mb) |= 1 << b;
            }
            if (seen == 15) { split_a = ma; split_b = mb; }
        }
        const char *p = getenv("P265_SPLIT_PCT");
        if (p) {
            int x = 0, y = 0;
            if (sscanf(p, "%d,%d", &x, &y) == 2 && x >= 10 && x <= 100 && y >= 10 && y <= 100) { pct_a = x; pct_b = y; }
        }
    }
    auto chain = [&](int mask, cudaStream_t st, bool aux, int pct_limit) -> int {
        bool first = true;
        for (int i = 0; i < 4; i++) {
            const int b = order[i];
            if (!a.n_tb[b] || !(mask & (1 << b))) continue;
            const bool plain = aux && first;
            int r;
            switch (b) {
                case 0: r = launch_bin<0, SF>(ctx, a, first, st, plain, pct_limit); break;
                case 1: r = launch_bin<1, SF>(ctx, a, first, st, plain, pct_limit); break;
                case 2: r = launch_bin<2, SF>(ctx, a, first, st, plain, pct_limit); break;
                default: r = launch_bin<3, SF>(ctx, a, first, st, plain, pct_limit); break;
            }
            if (r) return r;
            first = false;
        }
        return P265_OK;
    };
    auto has = [&](int mask) { for (int b = 0; b < 4; b++) if ((mask & (1 << b)) && a.n_tb[b]) return true; return false; };
    if (split_a && has(split_a) && has(split_b)) {
        if (!ctx->aux_stream) {
            P265_CUDA(cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
            P265_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
            P265_CUDA(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
        }
        P265_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));            // expand_kernel (and everything before) is done
        P265_CUDA(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
        if ((rc = chain(split_b, ctx->aux_stream, true, pct_b))) return rc;
        if ((rc = chain(split_a, ctx->stream, false, pct_a))) return rc;
        P265_CUDA(cudaEventRecord(ctx->ev_join, ctx->aux_stream));
        P265_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
        return P265_OK;
    }
    return chain(15, ctx->stream, false, 0);
}

int launch_residual(p265_ctx *ctx, const p265_tu_desc *d_tus, const int32_t bin_counts[4], const int16_t *d_coeffs,
                    const uint8_t *d_sf, const p265_pic_geom *g, int16_t *d_out, int flags) {
    if ((uint64_t)g->pic_stride * (uint64_t)g->n_pics > 0xffffffffull)
        return set_error(P265_EINVAL, "residual planes of one batch must stay below 2^32 elements: split the batch");
    KernelArgs a;
    int rc = fill_args(a, d_tus, bin_counts, d_coeffs, d_sf, g, d_out);
    if (rc) return rc;
    a.sf_replicated = (flags & P265_RES_SF_REPLICATED) != 0;
    a.dense_arena = (flags & P265_RES_DENSE_ARENA) != 0;
    a.zext = (flags & P265_RES_ZERO_EXTENTS) != 0;
    if (flags & P265_RES_ZERO_FILL)
        P265_CUDA(cudaMemsetAsync(d_out, 0, sizeof(int16_t) * (size_t)g->pic_stride * g->n_pics, ctx->stream));
    if (a.first_item[4] == 0) return P265_OK;
    // Expanded descriptors for the 32x32 / 16x16 bins (16 or 8 lanes share a TB there and would
    // all re-derive its parameters) and, with a ScalingFactor table, for the small bins as well
    // (SmallDesc), in a grow-only device buffer of the context.  The list is sorted by size.
    const int n_tus = a.n_tb[0] + a.n_tb[1] + (a.sf ? a.n_tb[2] + a.n_tb[3] : 0);
    const size_t need = sizeof(uint4) * (size_t)n_tus;
    if (ctx->xtus_bytes < need) {
        if (ctx->xtus) P265_CUDA(cudaFree(ctx->xtus));
        ctx->xtus = nullptr;
        ctx->xtus_bytes = 0;
        P265_CUDA(cudaMalloc(&ctx->xtus, need + need / 4));
        ctx->xtus_bytes = need + need / 4;
    }
    a.xtus = static_cast<const uint4 *>(ctx->xtus);
    if (n_tus) {
        expand_kernel<<<(n_tus + 1023) / 1024, 256, 0, ctx->stream>>>(a, n_tus, static_cast<uint4 *>(ctx->xtus));
        P265_CUDA(cudaGetLastError());
        ctx->launches++;
    }
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
