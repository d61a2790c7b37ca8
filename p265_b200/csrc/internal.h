// Internal glue shared by the translation units of libp265b200.so (not installed).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "p265_b200.h"

struct p265_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    int sm_count = 0;
    bool async_mode = false;  // host entry points return after enqueueing (p265_ctx_set_async)
    uint64_t launches = 0;  // kernels launched by this context (bench.py: gpu_launches)
    // grow-only device scratch for the host-buffer entry points
    // (slots are shared between entry points; the context's single stream serialises their use)
    enum { kScratchSlots = 12 };
    void *scratch[kScratchSlots] = {nullptr};
    size_t scratch_bytes[kScratchSlots] = {0};
    // optional timeline of the host entry points (p265_ctx_set_trace): events on the context's stream
    // around the phases of a call (kind: 1 residual, 2 SAO, 3 loop filter; phase: 0 start, 1 inputs
    // copied, 2 kernels done, 3 outputs back)
    bool trace = false;
    struct TraceMark { int kind, phase; cudaEvent_t ev; };
    std::vector<TraceMark> marks;
    cudaStream_t aux_stream = nullptr;  // second chain of residual bins (tuning knob P265_SPLIT), created on demand
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // host copy of the ScalingFactor table that scratch slot 3 currently holds on the device: the table of a
    // stream changes with its parameter sets, not with its pictures, so it is uploaded when it differs
    uint8_t sf_shadow[P265_SF_BYTES] = {0};
    bool sf_shadow_valid = false;
    void *xtus = nullptr;  // expanded TU descriptors of the current residual launch (grow-only)
    size_t xtus_bytes = 0;
};

namespace p265 {

int set_error(int code, const char *fmt, ...);
void trace_mark(p265_ctx *ctx, int kind, int phase);
int cuda_error(cudaError_t e, const char *what, const char *file, int line);

#define P265_CUDA(expr)                                                        \
    do {                                                                       \
        cudaError_t e__ = (expr);                                              \
        if (e__ != cudaSuccess) return ::p265::cuda_error(e__, #expr, __FILE__, __LINE__); \
    } while (0)

int launch_residual(p265_ctx *ctx, const p265_tu_desc *d_tus, const int32_t bin_counts[4], const int16_t *d_coeffs,
                    const uint8_t *d_sf, const p265_pic_geom *g, int16_t *d_out, int flags);
int launch_unpack(p265_ctx *ctx, const p265_tu_desc *d_tus, const int32_t bin_counts[4], const uint8_t *d_stream,
                  int16_t *d_arena, p265_tu_desc *d_tus_out);
// SAO write-back into page-locked host memory: only the CTB components with sao type != 0 (h_out is the
// device-visible address of the caller's buffer, which already holds the unfiltered samples)
int launch_sao_writeback(p265_ctx *ctx, const void *d_out, void *h_out, const p265_pic_geom *g, int ctb_log2,
                         const p265_sao_ctb *d_params);
int run_pcie_probe(p265_ctx *ctx, size_t bytes, int n_buffers, int reps, double *h2d, double *d2h);
int launch_dequant(p265_ctx *ctx, const p265_tu_desc *d_tus, int n_tus, const int16_t *d_coeffs, const uint8_t *d_sf,
                   int bit_depth_y, int bit_depth_c, int16_t *d_scaled);
int launch_ref_literal(p265_ctx *ctx, const p265_tu_desc *d_tus, int n_tus, const int16_t *d_scaled, int32_t *d_out);
int launch_idct1d(p265_ctx *ctx, const int32_t *d_x, int log2size, int tr_type, int mode, int32_t *d_y);
int launch_sao(p265_ctx *ctx, const void *d_rec, void *d_out, const p265_pic_geom *g, int ctb_log2,
               const p265_sao_ctb *d_params, const uint8_t *d_no_filter);
int launch_recon(p265_ctx *ctx, const void *d_pred, const int16_t *d_res, void *d_rec, const p265_pic_geom *g);
int launch_deblock(p265_ctx *ctx, void *d_pix, const p265_pic_geom *g, int ctb_log2, const p265_dbk_blk *d_blk,
                   const p265_dbk_ctb *d_ctb);
int run_int_peak(p265_ctx *ctx, int kind, double *ops_per_s, double *ms);

}  // namespace p265
