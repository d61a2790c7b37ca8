// C-ABI of libp265b200.so (include/p265_b200.h): contexts, argument validation, the
// host-buffer entry points (H2D -> kernel -> D2H on the context's stream) and the
// device-resident entry points the benchmark times.  No CPU fallback anywhere: every
// entry point needs a live CUDA context or fails with P265_ECUDA.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "internal.h"

namespace p265 {

static thread_local char g_err[512] = "";

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

int cuda_error(cudaError_t e, const char *what, const char *file, int line) {
    const char *base = strrchr(file, '/');
    set_error(P265_ECUDA, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), base ? base + 1 : file,
              line, what);
    return e == cudaErrorMemoryAllocation ? P265_ENOMEM : P265_ECUDA;
}

// process-wide time origin of the traces (one per device would do; events of one device share a clock)
static cudaEvent_t g_trace_origin[64] = {nullptr};

void trace_mark(p265_ctx *ctx, int kind, int phase) {
    if (!ctx->trace) return;
    cudaEvent_t ev;
    if (cudaEventCreate(&ev) != cudaSuccess) return;
    cudaEventRecord(ev, ctx->stream);
    ctx->marks.push_back({kind, phase, ev});
}

static int ensure(p265_ctx *ctx, int slot, size_t bytes, void **out);

// ScalingFactor table -> scratch slot 3, copied only when its content changed since the last upload of this
// context (one 4 KB copy less per picture on the H2D engine; the stream orders a new upload behind the kernels
// that still read the old table)
static int upload_sf(p265_ctx *ctx, const uint8_t *scaling_factor, void **d_sf) {
    int rc = ensure(ctx, 3, P265_SF_BYTES, d_sf);
    if (rc) return rc;
    if (ctx->sf_shadow_valid && memcmp(ctx->sf_shadow, scaling_factor, P265_SF_BYTES) == 0) return P265_OK;
    memcpy(ctx->sf_shadow, scaling_factor, P265_SF_BYTES);   // the copy below reads OUR copy: the caller's may be gone by then
    ctx->sf_shadow_valid = false;
    P265_CUDA(cudaMemcpyAsync(*d_sf, ctx->sf_shadow, P265_SF_BYTES, cudaMemcpyHostToDevice, ctx->stream));
    ctx->sf_shadow_valid = true;
    return P265_OK;
}

static int ensure(p265_ctx *ctx, int slot, size_t bytes, void **out) {
    if (bytes == 0) bytes = 16;
    if (ctx->scratch_bytes[slot] < bytes) {
        if (ctx->scratch[slot]) P265_CUDA(cudaFree(ctx->scratch[slot]));
        ctx->scratch[slot] = nullptr;
        ctx->scratch_bytes[slot] = 0;
        if (slot == 3) ctx->sf_shadow_valid = false;
        size_t cap = bytes + bytes / 4;  // grow-only with head-room
        P265_CUDA(cudaMalloc(&ctx->scratch[slot], cap));
        ctx->scratch_bytes[slot] = cap;
    }
    *out = ctx->scratch[slot];
    return P265_OK;
}

static int check_geom(const p265_pic_geom *g, int elem_align) {
    if (!g) return set_error(P265_EINVAL, "geometry is NULL");
    if (g->width <= 0 || g->height <= 0 || (g->width & 1) || (g->height & 1))
        return set_error(P265_EINVAL, "picture size %dx%d must be positive and even", g->width, g->height);
    if (g->n_pics <= 0) return set_error(P265_EINVAL, "n_pics must be positive");
    if (g->bit_depth_y < 8 || g->bit_depth_y > 12 || g->bit_depth_c < 8 || g->bit_depth_c > 12)
        return set_error(P265_EINVAL, "bit depths %d/%d outside 8..12", g->bit_depth_y, g->bit_depth_c);
    if (g->stride_y < g->width || g->stride_c < g->width / 2)
        return set_error(P265_EINVAL, "strides %d/%d smaller than the plane widths", g->stride_y, g->stride_c);
    if (g->stride_y % elem_align || g->stride_c % elem_align || g->pic_stride % elem_align)
        return set_error(P265_EINVAL, "strides must be multiples of %d elements (16-byte rows)", elem_align);
    for (int c = 0; c < 3; c++)
        if (g->plane_off[c] < 0 || g->plane_off[c] % elem_align)
            return set_error(P265_EINVAL, "plane_off[%d] must be a non-negative multiple of %d", c, elem_align);
    const int64_t need_y = (int64_t)g->stride_y * g->height, need_c = (int64_t)g->stride_c * (g->height / 2);
    if (g->plane_off[0] + need_y > g->pic_stride || g->plane_off[1] + need_c > g->pic_stride ||
        g->plane_off[2] + need_c > g->pic_stride)
        return set_error(P265_EINVAL, "planes do not fit inside pic_stride");
    return P265_OK;
}

// host-side validation of a descriptor list (host entry points only)
// *dense_small = inside the 8x8 bin and inside the 4x4 bin every TB's coefficients directly follow
// the previous TB's (what P265_RES_DENSE_ARENA asserts on the device entry point)
// packed: the descriptors index a packed coefficient stream of n_coeffs BYTES (records of significance
// bitmap + levels, include/p265_b200.h; `rsvd` = number of levels) instead of a dense arena of n_coeffs
// int16.  A record is bounded by its descriptor alone (bitmap size + rsvd levels): the device never reads
// more than rsvd levels of a TB whatever its bitmap says, so no pass over the stream is needed here.
struct TuLimits {
    int wmax[3], hmax[3], qmax[3];
    int n_pics;
    size_t n_coeffs;
};

template <bool PACKED>
static inline unsigned tu_bad(const p265_tu_desc &t, int log2n, unsigned bad_flags, const TuLimits &L) {
    const int n = 1 << log2n;
    const unsigned c = t.c_idx < 3 ? t.c_idx : 0;
    unsigned bad = (unsigned)(t.log2n != log2n) | (unsigned)(t.c_idx > 2) | (unsigned)(((t.x | t.y) & (n - 1)) != 0) |
                   (unsigned)(t.x + n > L.wmax[c]) | (unsigned)(t.y + n > L.hmax[c]) | (unsigned)(t.pic >= L.n_pics) |
                   (unsigned)((t.flags & bad_flags) != 0) | (unsigned)(((t.flags & P265_TU_DST) != 0) & (t.c_idx != 0)) |
                   (unsigned)(t.qp > L.qmax[c]);
    if (PACKED) {
        const unsigned lv = t.rsvd & P265_TU_LEVELS_MASK;
        const size_t end = (size_t)t.coeff_off * 4 + (size_t)(n * n) / 8 + (size_t)lv * ((t.flags & P265_TU_LEVELS8) ? 1 : 2);
        bad |= (unsigned)(end > L.n_coeffs) | (unsigned)(lv > (unsigned)(n * n));
    } else {
        bad |= (unsigned)((size_t)t.coeff_off * 16 + (size_t)(n * n) > L.n_coeffs);
        // zero-extent codes (a promise about the dense arena; the packed path derives its own): 3 is undefined
        bad |= (unsigned)(((t.rsvd >> P265_TU_ZR_SHIFT) & 3) == 3) | (unsigned)(((t.rsvd >> P265_TU_ZC_SHIFT) & 3) == 3);
    }
    return bad;
}

// the first offending descriptor, in words (only reached when the fast pass found one)
static int diagnose_tu(const p265_tu_desc &t, long long k, int log2n, bool have_table, bool packed, const TuLimits &L) {
    const int n = 1 << log2n;
    const unsigned c = t.c_idx < 3 ? t.c_idx : 0;
    if (t.log2n != log2n)
        return set_error(P265_EINVAL, "descriptor %lld: log2n %d where bin expects %d (list must be sorted 32,16,8,4)", k,
                         t.log2n, log2n);
    if (t.c_idx > 2) return set_error(P265_EINVAL, "descriptor %lld: c_idx %d", k, t.c_idx);
    if (t.x % n || t.y % n || t.x + n > L.wmax[c] || t.y + n > L.hmax[c])
        return set_error(P265_EINVAL, "descriptor %lld: %dx%d block at (%d,%d) outside the %dx%d plane or unaligned", k, n, n,
                         t.x, t.y, L.wmax[c], L.hmax[c]);
    if (t.pic >= L.n_pics) return set_error(P265_EINVAL, "descriptor %lld: picture %d", k, t.pic);
    if (packed && (t.rsvd & P265_TU_LEVELS_MASK) > n * n)
        return set_error(P265_EINVAL, "descriptor %lld: %d levels in a %dx%d block", k, t.rsvd & P265_TU_LEVELS_MASK, n, n);
    if (!packed && (((t.rsvd >> P265_TU_ZR_SHIFT) & 3) == 3 || ((t.rsvd >> P265_TU_ZC_SHIFT) & 3) == 3))
        return set_error(P265_EINVAL, "descriptor %lld: zero-extent code 3 is not defined (rsvd 0x%04x)", k, t.rsvd);
    if (tu_bad<true>(t, log2n, 0, L) && packed)
        return set_error(P265_EINVAL, "descriptor %lld: coefficients beyond the packed stream", k);
    if (!packed && (size_t)t.coeff_off * 16 + (size_t)(n * n) > L.n_coeffs)
        return set_error(P265_EINVAL, "descriptor %lld: coefficients beyond the arena", k);
    if (t.flags & (packed ? 0xc0u : (0xc0u | P265_TU_LEVELS8)))
        return set_error(P265_EINVAL, "descriptor %lld: flags 0x%02x not defined for this entry point", k, t.flags);
    if ((t.flags & P265_TU_SKIP) && log2n != 2)
        return set_error(P265_EINVAL, "descriptor %lld: transform_skip on a %dx%d block", k, n, n);
    if ((t.flags & P265_TU_DST) && (log2n != 2 || t.c_idx != 0))
        return set_error(P265_EINVAL, "descriptor %lld: DST on a non-4x4-luma block", k);
    if (have_table && (t.flags & P265_TU_PRESCALED))
        return set_error(P265_EINVAL, "descriptor %lld: PRESCALED needs scaling_factor == NULL", k);
    return set_error(P265_EINVAL, "descriptor %lld: qP %d out of range", k, t.qp);
}

// *any_codes = some 16x16 / 32x32 descriptor of a dense arena carries a zero-extent code (P265_RES_ZERO_EXTENTS)
static int check_tus(const p265_tu_desc *tus, const int32_t bin_counts[4], size_t n_coeffs, const p265_pic_geom *g,
                     bool have_table, bool *dense_small, bool packed = false, bool *any_codes = nullptr) {
    // one branch-free pass over caller data, on the latency path of every host call (119 k descriptors per
    // 4K picture); the diagnostics are formatted by a second look only when something is wrong
    TuLimits L;
    for (int c = 0; c < 3; c++) {
        L.wmax[c] = c ? g->width / 2 : g->width;
        L.hmax[c] = c ? g->height / 2 : g->height;
        L.qmax[c] = 51 + 6 * ((c ? g->bit_depth_c : g->bit_depth_y) - 8);
    }
    L.n_pics = g->n_pics;
    L.n_coeffs = n_coeffs;
    int64_t k = 0;
    uint32_t not_dense = 0, codes = 0;
    for (int b = 0; b < 4; b++) {
        if (bin_counts[b] < 0) return set_error(P265_EINVAL, "negative bin count");
        const int log2n = 5 - b, n = 1 << log2n;
        const uint32_t units = (uint32_t)(n * n) / 16, z0 = bin_counts[b] ? tus[k].coeff_off : 0u;
        const unsigned bad_flags = (log2n != 2 ? (P265_TU_SKIP | P265_TU_DST) : 0u) | (have_table ? P265_TU_PRESCALED : 0u) |
                                   (packed ? 0u : P265_TU_LEVELS8) | 0xc0u;
        const p265_tu_desc *t = tus + k;
        const int32_t cnt = bin_counts[b];
        unsigned bad = 0;
        if (packed) for (int32_t i = 0; i < cnt; i++) bad |= tu_bad<true>(t[i], log2n, bad_flags, L);
        else for (int32_t i = 0; i < cnt; i++) bad |= tu_bad<false>(t[i], log2n, bad_flags, L);
        if (b >= 2) for (int32_t i = 0; i < cnt; i++) not_dense |= t[i].coeff_off ^ (z0 + (uint32_t)i * units);
        else for (int32_t i = 0; i < cnt; i++) codes |= t[i].rsvd;
        if (bad)
            for (int32_t i = 0; i < cnt; i++)
                if (packed ? tu_bad<true>(t[i], log2n, bad_flags, L) : tu_bad<false>(t[i], log2n, bad_flags, L))
                    return diagnose_tu(t[i], (long long)(k + i), log2n, have_table, packed, L);
        k += cnt;
    }
    *dense_small = not_dense == 0;
    if (any_codes) *any_codes = (codes >> P265_TU_ZR_SHIFT) != 0;
    return P265_OK;
}

}  // namespace p265

using namespace p265;

extern "C" {
#pragma GCC visibility push(default)

int p265_abi_version(void) { return P265_ABI_VERSION; }

const char *p265_last_error(void) { return g_err; }

int p265_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int p265_ctx_create(int device, void *stream, p265_ctx **out) {
    if (!out) return set_error(P265_EINVAL, "p265_ctx_create: out is NULL");
    *out = nullptr;
    int n = 0;
    P265_CUDA(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) return set_error(P265_EINVAL, "device %d out of range (have %d)", device, n);
    P265_CUDA(cudaSetDevice(device));
    p265_ctx *ctx = new p265_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        delete ctx;
        return cuda_error(e, "cudaGetDeviceProperties", __FILE__, __LINE__);
    }
    if (prop.major < 10) {
        delete ctx;
        return set_error(P265_ECUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                         prop.minor);
    }
    ctx->sm_count = prop.multiProcessorCount;
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            delete ctx;
            return cuda_error(e, "cudaStreamCreateWithFlags", __FILE__, __LINE__);
        }
        ctx->owns_stream = true;
    }
    *out = ctx;
    return P265_OK;
}

int p265_ctx_destroy(p265_ctx *ctx) {
    if (!ctx) return P265_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < p265_ctx::kScratchSlots; i++)
        if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
    if (ctx->xtus) cudaFree(ctx->xtus);
    for (const auto &m : ctx->marks) cudaEventDestroy(m.ev);
    if (ctx->aux_stream) {
        cudaStreamSynchronize(ctx->aux_stream);
        cudaStreamDestroy(ctx->aux_stream);
        cudaEventDestroy(ctx->ev_fork);
        cudaEventDestroy(ctx->ev_join);
    }
    if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return P265_OK;
}

int p265_sync(p265_ctx *ctx) {
    if (!ctx) return set_error(P265_EINVAL, "ctx is NULL");
    P265_CUDA(cudaStreamSynchronize(ctx->stream));
    return P265_OK;
}

int p265_ctx_set_async(p265_ctx *ctx, int enable) {
    if (!ctx) return set_error(P265_EINVAL, "ctx is NULL");
    ctx->async_mode = enable != 0;
    return P265_OK;
}

// end of a host-buffer entry point: wait for the result unless the context is asynchronous
static int finish(p265_ctx *ctx) {
    if (ctx->async_mode) return P265_OK;
    P265_CUDA(cudaStreamSynchronize(ctx->stream));
    return P265_OK;
}

int p265_ctx_set_trace(p265_ctx *ctx, int enable) {
    if (!ctx) return set_error(P265_EINVAL, "ctx is NULL");
    P265_CUDA(cudaSetDevice(ctx->device));
    cudaEvent_t &origin = g_trace_origin[ctx->device & 63];
    if (enable && !origin) {
        P265_CUDA(cudaEventCreate(&origin));
        P265_CUDA(cudaEventRecord(origin, ctx->stream));
        P265_CUDA(cudaEventSynchronize(origin));
    }
    ctx->trace = enable != 0;
    return P265_OK;
}

int p265_trace_read(p265_ctx *ctx, double *out, int max_marks) {
    if (!ctx || (max_marks > 0 && !out)) return set_error(P265_EINVAL, "p265_trace_read: NULL argument");
    P265_CUDA(cudaSetDevice(ctx->device));
    P265_CUDA(cudaStreamSynchronize(ctx->stream));
    int n = 0;
    for (const auto &m : ctx->marks) {
        float ms = 0;
        if (n < max_marks && cudaEventElapsedTime(&ms, g_trace_origin[ctx->device & 63], m.ev) == cudaSuccess) {
            out[3 * n] = m.kind;
            out[3 * n + 1] = m.phase;
            out[3 * n + 2] = ms;
            n++;
        }
        cudaEventDestroy(m.ev);
    }
    cudaGetLastError();
    ctx->marks.clear();
    return n;
}

int p265_sm_count(p265_ctx *ctx) { return ctx ? ctx->sm_count : 0; }

uint64_t p265_launch_count(p265_ctx *ctx) { return ctx ? ctx->launches : 0; }

int p265_residual_batch_dev(p265_ctx *ctx, const p265_tu_desc *d_tus, const int32_t bin_counts[4],
                            const int16_t *d_coeffs, const uint8_t *d_sf, const p265_pic_geom *geom,
                            int16_t *d_residual, int flags) {
    if (!ctx || !bin_counts || !d_residual) return set_error(P265_EINVAL, "p265_residual_batch_dev: NULL argument");
    int rc = check_geom(geom, 8);
    if (rc) return rc;
    int64_t n = 0;
    for (int b = 0; b < 4; b++) {
        if (bin_counts[b] < 0) return set_error(P265_EINVAL, "negative bin count");
        n += bin_counts[b];
    }
    if (n && (!d_tus || !d_coeffs)) return set_error(P265_EINVAL, "descriptor / coefficient pointer is NULL");
    P265_CUDA(cudaSetDevice(ctx->device));
    return launch_residual(ctx, d_tus, bin_counts, d_coeffs, d_sf, geom, d_residual, flags);
}

int p265_residual_batch(p265_ctx *ctx, const p265_tu_desc *tus, const int32_t bin_counts[4], const int16_t *coeffs,
                        size_t n_coeffs, const uint8_t *scaling_factor, const p265_pic_geom *geom, int16_t *residual,
                        int flags) {
    if (!ctx || !bin_counts || !residual) return set_error(P265_EINVAL, "p265_residual_batch: NULL argument");
    int rc = check_geom(geom, 8);
    if (rc) return rc;
    int64_t n = 0;
    for (int b = 0; b < 4; b++) n += bin_counts[b] > 0 ? bin_counts[b] : 0;
    if (n && (!tus || !coeffs)) return set_error(P265_EINVAL, "descriptor / coefficient pointer is NULL");
    bool dense_small = false, any_codes = false;
    if ((rc = check_tus(tus, bin_counts, n_coeffs, geom, scaling_factor != nullptr, &dense_small, false, &any_codes))) return rc;
    flags = dense_small ? (flags | P265_RES_DENSE_ARENA) : (flags & ~P265_RES_DENSE_ARENA);  // found out here, not asserted
    flags = any_codes ? (flags | P265_RES_ZERO_EXTENTS) : (flags & ~P265_RES_ZERO_EXTENTS);  // likewise
    // the whole buffer travels back to the host (row padding and inter-plane gaps included) and the
    // scratch slot is shared with other entry points: never return stale bytes of an earlier call
    flags |= P265_RES_ZERO_FILL;
    P265_CUDA(cudaSetDevice(ctx->device));
    void *d_tus, *d_co, *d_sf = nullptr, *d_out;
    const size_t out_bytes = sizeof(int16_t) * (size_t)geom->pic_stride * geom->n_pics;
    if ((rc = ensure(ctx, 0, sizeof(p265_tu_desc) * (size_t)n, &d_tus))) return rc;
    if ((rc = ensure(ctx, 1, sizeof(int16_t) * n_coeffs + 64, &d_co))) return rc;
    if ((rc = ensure(ctx, 2, out_bytes, &d_out))) return rc;
    if (scaling_factor) {
        if ((rc = upload_sf(ctx, scaling_factor, &d_sf))) return rc;
    }
    if (n) {
        P265_CUDA(cudaMemcpyAsync(d_tus, tus, sizeof(p265_tu_desc) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
        P265_CUDA(cudaMemcpyAsync(d_co, coeffs, sizeof(int16_t) * n_coeffs, cudaMemcpyHostToDevice, ctx->stream));
    }
    rc = launch_residual(ctx, (const p265_tu_desc *)d_tus, bin_counts, (const int16_t *)d_co, (const uint8_t *)d_sf,
                         geom, (int16_t *)d_out, flags);
    if (rc) return rc;
    P265_CUDA(cudaMemcpyAsync(residual, d_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return finish(ctx);
}

static int64_t dense_elems(const int32_t bin_counts[4]) {
    int64_t e = 0;
    for (int b = 0; b < 4; b++) e += (int64_t)(bin_counts[b] > 0 ? bin_counts[b] : 0) << (2 * (5 - b));
    return e;
}

int p265_residual_batch_packed_dev(p265_ctx *ctx, const p265_tu_desc *d_tus, const int32_t bin_counts[4],
                                   const uint8_t *d_stream, const uint8_t *d_sf, const p265_pic_geom *geom,
                                   int16_t *d_arena, p265_tu_desc *d_tus_out, int16_t *d_residual, int flags) {
    if (!ctx || !bin_counts || !d_residual) return set_error(P265_EINVAL, "p265_residual_batch_packed_dev: NULL argument");
    int rc = check_geom(geom, 8);
    if (rc) return rc;
    int64_t n = 0;
    for (int b = 0; b < 4; b++) {
        if (bin_counts[b] < 0) return set_error(P265_EINVAL, "negative bin count");
        n += bin_counts[b];
    }
    if (n && (!d_tus || !d_stream || !d_arena || !d_tus_out))
        return set_error(P265_EINVAL, "descriptor / stream / arena pointer is NULL");
    P265_CUDA(cudaSetDevice(ctx->device));
    if ((rc = launch_unpack(ctx, d_tus, bin_counts, d_stream, d_arena, d_tus_out))) return rc;
    // the arena comes out in descriptor order: the small bins are dense by construction
    return launch_residual(ctx, d_tus_out, bin_counts, d_arena, d_sf, geom, d_residual,
                           flags | P265_RES_DENSE_ARENA | P265_RES_ZERO_EXTENTS);
}

int p265_residual_batch_packed(p265_ctx *ctx, const p265_tu_desc *tus, const int32_t bin_counts[4], const uint8_t *stream,
                               size_t stream_bytes, const uint8_t *scaling_factor, const p265_pic_geom *geom,
                               int16_t *residual, int flags) {
    if (!ctx || !bin_counts || !residual) return set_error(P265_EINVAL, "p265_residual_batch_packed: NULL argument");
    int rc = check_geom(geom, 8);
    if (rc) return rc;
    int64_t n = 0;
    for (int b = 0; b < 4; b++) n += bin_counts[b] > 0 ? bin_counts[b] : 0;
    if (n && (!tus || !stream)) return set_error(P265_EINVAL, "descriptor / stream pointer is NULL");
    bool dense_small = false;
    if ((rc = check_tus(tus, bin_counts, stream_bytes, geom, scaling_factor != nullptr, &dense_small, true))) return rc;
    flags |= P265_RES_ZERO_FILL;  // see p265_residual_batch
    P265_CUDA(cudaSetDevice(ctx->device));
    void *d_tus, *d_st, *d_sf = nullptr, *d_out, *d_arena, *d_tus2;
    const size_t out_bytes = sizeof(int16_t) * (size_t)geom->pic_stride * geom->n_pics;
    if ((rc = ensure(ctx, 0, sizeof(p265_tu_desc) * (size_t)n, &d_tus))) return rc;
    if ((rc = ensure(ctx, 1, stream_bytes + 64, &d_st))) return rc;
    if ((rc = ensure(ctx, 2, out_bytes, &d_out))) return rc;
    if ((rc = ensure(ctx, 8, sizeof(int16_t) * (size_t)dense_elems(bin_counts) + 64, &d_arena))) return rc;
    if ((rc = ensure(ctx, 9, sizeof(p265_tu_desc) * (size_t)n, &d_tus2))) return rc;
    if (scaling_factor) {
        if ((rc = upload_sf(ctx, scaling_factor, &d_sf))) return rc;
    }
    trace_mark(ctx, 1, 0);
    if (n) {
        P265_CUDA(cudaMemcpyAsync(d_tus, tus, sizeof(p265_tu_desc) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
        P265_CUDA(cudaMemcpyAsync(d_st, stream, stream_bytes, cudaMemcpyHostToDevice, ctx->stream));
        trace_mark(ctx, 1, 1);
        if ((rc = launch_unpack(ctx, (const p265_tu_desc *)d_tus, bin_counts, (const uint8_t *)d_st, (int16_t *)d_arena,
                                (p265_tu_desc *)d_tus2)))
            return rc;
    }
    rc = launch_residual(ctx, (const p265_tu_desc *)d_tus2, bin_counts, (const int16_t *)d_arena, (const uint8_t *)d_sf,
                         geom, (int16_t *)d_out, flags | P265_RES_DENSE_ARENA | P265_RES_ZERO_EXTENTS);
    if (rc) return rc;
    trace_mark(ctx, 1, 2);
    P265_CUDA(cudaMemcpyAsync(residual, d_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    trace_mark(ctx, 1, 3);
    return finish(ctx);
}

int p265_dequant_batch(p265_ctx *ctx, const p265_tu_desc *tus, int32_t n_tus, const int16_t *coeffs, size_t n_coeffs,
                       const uint8_t *scaling_factor, int bit_depth_y, int bit_depth_c, int16_t *scaled) {
    if (!ctx || n_tus < 0 || (n_tus && (!tus || !coeffs || !scaled)))
        return set_error(P265_EINVAL, "p265_dequant_batch: bad argument");
    if (bit_depth_y < 8 || bit_depth_y > 12 || bit_depth_c < 8 || bit_depth_c > 12)
        return set_error(P265_EINVAL, "bit depths %d/%d outside 8..12", bit_depth_y, bit_depth_c);
    const int qmax[3] = {51 + 6 * (bit_depth_y - 8), 51 + 6 * (bit_depth_c - 8), 51 + 6 * (bit_depth_c - 8)};
    for (int32_t i = 0; i < n_tus; i++) {
        const int n = 1 << tus[i].log2n;
        if (tus[i].log2n < 2 || tus[i].log2n > 5 || tus[i].c_idx > 2 ||
            (size_t)tus[i].coeff_off * 16 + (size_t)n * n > n_coeffs)
            return set_error(P265_EINVAL, "descriptor %d is malformed", i);
        if (tus[i].qp > qmax[tus[i].c_idx])  // same bound as check_tus: keeps make_params' qp / 6 and shifts defined
            return set_error(P265_EINVAL, "descriptor %d: qP %d out of range", i, tus[i].qp);
    }
    if (n_tus == 0) return P265_OK;
    P265_CUDA(cudaSetDevice(ctx->device));
    int rc;
    void *d_tus, *d_co, *d_sf = nullptr, *d_out;
    if ((rc = ensure(ctx, 0, sizeof(p265_tu_desc) * (size_t)n_tus, &d_tus))) return rc;
    if ((rc = ensure(ctx, 1, sizeof(int16_t) * n_coeffs + 64, &d_co))) return rc;
    if ((rc = ensure(ctx, 4, sizeof(int16_t) * n_coeffs, &d_out))) return rc;
    if (scaling_factor) {
        if ((rc = upload_sf(ctx, scaling_factor, &d_sf))) return rc;
    }
    P265_CUDA(cudaMemcpyAsync(d_tus, tus, sizeof(p265_tu_desc) * (size_t)n_tus, cudaMemcpyHostToDevice, ctx->stream));
    P265_CUDA(cudaMemcpyAsync(d_co, coeffs, sizeof(int16_t) * n_coeffs, cudaMemcpyHostToDevice, ctx->stream));
    P265_CUDA(cudaMemsetAsync(d_out, 0, sizeof(int16_t) * n_coeffs, ctx->stream));
    if ((rc = launch_dequant(ctx, (const p265_tu_desc *)d_tus, n_tus, (const int16_t *)d_co, (const uint8_t *)d_sf,
                             bit_depth_y, bit_depth_c, (int16_t *)d_out)))
        return rc;
    P265_CUDA(cudaMemcpyAsync(scaled, d_out, sizeof(int16_t) * n_coeffs, cudaMemcpyDeviceToHost, ctx->stream));
    P265_CUDA(cudaStreamSynchronize(ctx->stream));
    return P265_OK;
}

int p265_ref_literal_batch(p265_ctx *ctx, const p265_tu_desc *tus, int32_t n_tus, const int16_t *scaled,
                           size_t n_coeffs, int32_t *out) {
    if (!ctx || n_tus < 0 || (n_tus && (!tus || !scaled || !out)))
        return set_error(P265_EINVAL, "p265_ref_literal_batch: bad argument");
    for (int32_t i = 0; i < n_tus; i++) {
        const int n = 1 << tus[i].log2n;
        if (tus[i].log2n < 2 || tus[i].log2n > 5 || tus[i].c_idx > 2 ||
            (size_t)tus[i].coeff_off * 16 + (size_t)n * n > n_coeffs)
            return set_error(P265_EINVAL, "descriptor %d is malformed", i);
    }
    if (n_tus == 0) return P265_OK;
    P265_CUDA(cudaSetDevice(ctx->device));
    int rc;
    void *d_tus, *d_in, *d_out;
    if ((rc = ensure(ctx, 0, sizeof(p265_tu_desc) * (size_t)n_tus, &d_tus))) return rc;
    if ((rc = ensure(ctx, 1, sizeof(int16_t) * n_coeffs + 64, &d_in))) return rc;
    if ((rc = ensure(ctx, 5, sizeof(int32_t) * n_coeffs, &d_out))) return rc;
    P265_CUDA(cudaMemcpyAsync(d_tus, tus, sizeof(p265_tu_desc) * (size_t)n_tus, cudaMemcpyHostToDevice, ctx->stream));
    P265_CUDA(cudaMemcpyAsync(d_in, scaled, sizeof(int16_t) * n_coeffs, cudaMemcpyHostToDevice, ctx->stream));
    P265_CUDA(cudaMemsetAsync(d_out, 0, sizeof(int32_t) * n_coeffs, ctx->stream));
    if ((rc = launch_ref_literal(ctx, (const p265_tu_desc *)d_tus, n_tus, (const int16_t *)d_in, (int32_t *)d_out)))
        return rc;
    P265_CUDA(cudaMemcpyAsync(out, d_out, sizeof(int32_t) * n_coeffs, cudaMemcpyDeviceToHost, ctx->stream));
    P265_CUDA(cudaStreamSynchronize(ctx->stream));
    return P265_OK;
}

int p265_idct_1d(p265_ctx *ctx, const int32_t *x, int log2size, int tr_type, int mode, int32_t *y) {
    if (!ctx || !x || !y) return set_error(P265_EINVAL, "p265_idct_1d: NULL argument");
    if (log2size < 2 || log2size > 5) return set_error(P265_EINVAL, "log2size %d outside 2..5", log2size);
    if (tr_type != 0 && !(tr_type == 1 && log2size == 2))
        return set_error(P265_EINVAL, "tr_type %d invalid for log2size %d", tr_type, log2size);
    P265_CUDA(cudaSetDevice(ctx->device));
    const int n = 1 << log2size;
    void *d;
    int rc = ensure(ctx, 5, sizeof(int32_t) * 64, &d);
    if (rc) return rc;
    int32_t *d_x = (int32_t *)d, *d_y = d_x + 32;
    P265_CUDA(cudaMemcpyAsync(d_x, x, sizeof(int32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = launch_idct1d(ctx, d_x, log2size, tr_type, mode ? 1 : 0, d_y))) return rc;
    P265_CUDA(cudaMemcpyAsync(y, d_y, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    P265_CUDA(cudaStreamSynchronize(ctx->stream));
    return P265_OK;
}

static int check_sao(const p265_pic_geom *geom, int ctb_log2, int *bytes) {
    *bytes = (geom->bit_depth_y > 8 || geom->bit_depth_c > 8) ? 2 : 1;
    int rc = check_geom(geom, 16 / *bytes);
    if (rc) return rc;
    if (ctb_log2 < 4 || ctb_log2 > 6) return set_error(P265_EINVAL, "ctb_log2 %d outside 4..6", ctb_log2);
    if (geom->width % 8 || geom->height % 8)
        return set_error(P265_EINVAL, "picture size must be a multiple of MinCbSize 8 (got %dx%d)", geom->width,
                         geom->height);
    const int ctb = 1 << ctb_log2;
    const int ctbs_w = (geom->width + ctb - 1) / ctb;
    if (geom->stride_y < ctbs_w * ctb || geom->stride_c < ctbs_w * ctb / 2)
        return set_error(P265_EINVAL, "strides must cover whole CTB columns (%d / %d elements)", ctbs_w * ctb,
                         ctbs_w * ctb / 2);
    return P265_OK;
}

int p265_sao_batch_dev(p265_ctx *ctx, const void *d_rec, void *d_out, const p265_pic_geom *geom, int ctb_log2,
                       const p265_sao_ctb *d_params, const uint8_t *d_no_filter) {
    if (!ctx || !d_rec || !d_out || !d_params || !geom) return set_error(P265_EINVAL, "p265_sao_batch_dev: NULL argument");
    if (d_rec == d_out) return set_error(P265_EINVAL, "SAO runs out of place: rec and out must differ");
    int bytes;
    int rc = check_sao(geom, ctb_log2, &bytes);
    if (rc) return rc;
    P265_CUDA(cudaSetDevice(ctx->device));
    return launch_sao(ctx, d_rec, d_out, geom, ctb_log2, d_params, d_no_filter);
}

// Device-visible address of a host buffer when it is page-locked and mapped (cudaHostAlloc / cudaHostRegister
// memory under unified addressing), nullptr for pageable memory.  P265_NO_ZERO_COPY=1 disables the direct
// write-back (A/B measurements).
static void *mapped_host_pointer(const void *p) {
    static int off = -1;
    if (off < 0) {
        const char *e = getenv("P265_NO_ZERO_COPY");
        off = (e && *e && *e != '0') ? 1 : 0;
    }
    if (off) return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}

// D2H of the plane rows only (row padding and inter-plane gaps of the host buffer stay untouched)
static int copy_planes_to_host(p265_ctx *ctx, void *host, const void *dev, const p265_pic_geom *g, int bytes) {
    for (int p = 0; p < g->n_pics; p++)
        for (int c = 0; c < 3; c++) {
            const size_t off = ((size_t)p * g->pic_stride + g->plane_off[c]) * bytes;
            const size_t pitch = (size_t)(c ? g->stride_c : g->stride_y) * bytes;
            const size_t row = (size_t)(c ? g->width / 2 : g->width) * bytes;
            const size_t rows = (size_t)(c ? g->height / 2 : g->height);
            if (row == pitch) {
                P265_CUDA(cudaMemcpyAsync((char *)host + off, (const char *)dev + off, row * rows, cudaMemcpyDeviceToHost, ctx->stream));
            } else {
                P265_CUDA(cudaMemcpy2DAsync((char *)host + off, pitch, (const char *)dev + off, pitch, row, rows,
                                            cudaMemcpyDeviceToHost, ctx->stream));
            }
        }
    return P265_OK;
}

static int check_sao_params(const p265_pic_geom *geom, int ctb_log2, const p265_sao_ctb *params, size_t *n_ctbs) {
    const int ctb = 1 << ctb_log2;
    const size_t ctbs = (size_t)((geom->width + ctb - 1) / ctb) * ((geom->height + ctb - 1) / ctb) * geom->n_pics;
    for (size_t i = 0; i < ctbs; i++)
        for (int c = 0; c < 3; c++)
            if (params[i].type[c] > 2 || params[i].eo_class[c] > 3 || params[i].band_pos[c] > 31)
                return set_error(P265_EINVAL, "SAO parameters of CTB %zu component %d out of range", i, c);
    *n_ctbs = ctbs;
    return P265_OK;
}

int p265_sao_batch(p265_ctx *ctx, const void *rec, void *out, const p265_pic_geom *geom, int ctb_log2,
                   const p265_sao_ctb *params, const uint8_t *no_filter) {
    if (!ctx || !rec || !out || !params || !geom) return set_error(P265_EINVAL, "p265_sao_batch: NULL argument");
    int bytes;
    int rc = check_sao(geom, ctb_log2, &bytes);
    if (rc) return rc;
    size_t ctbs;
    if ((rc = check_sao_params(geom, ctb_log2, params, &ctbs))) return rc;
    P265_CUDA(cudaSetDevice(ctx->device));
    const size_t plane_bytes = (size_t)bytes * geom->pic_stride * geom->n_pics;
    const size_t nf_bytes = (size_t)((geom->width + 7) / 8) * ((geom->height + 7) / 8) * geom->n_pics;
    void *d_rec, *d_out, *d_par, *d_nf = nullptr;
    if ((rc = ensure(ctx, 2, plane_bytes, &d_rec))) return rc;
    if ((rc = ensure(ctx, 6, plane_bytes, &d_out))) return rc;
    if ((rc = ensure(ctx, 0, sizeof(p265_sao_ctb) * ctbs, &d_par))) return rc;
    trace_mark(ctx, 2, 0);
    P265_CUDA(cudaMemcpyAsync(d_rec, rec, plane_bytes, cudaMemcpyHostToDevice, ctx->stream));
    P265_CUDA(cudaMemcpyAsync(d_par, params, sizeof(p265_sao_ctb) * ctbs, cudaMemcpyHostToDevice, ctx->stream));
    if (no_filter) {
        if ((rc = ensure(ctx, 7, nf_bytes, &d_nf))) return rc;
        P265_CUDA(cudaMemcpyAsync(d_nf, no_filter, nf_bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    trace_mark(ctx, 2, 1);
    if ((rc = launch_sao(ctx, d_rec, d_out, geom, ctb_log2, (const p265_sao_ctb *)d_par, (const uint8_t *)d_nf)))
        return rc;
    trace_mark(ctx, 2, 2);
    // In place on a page-locked host buffer: the host already holds every sample SAO leaves alone, so only
    // the CTB components with sao type != 0 cross the bus again, stored by a kernel straight into the
    // caller's buffer.  Otherwise: the plane rows by the copy engine.
    void *h_dev = out == rec ? mapped_host_pointer(out) : nullptr;
    if (h_dev) rc = launch_sao_writeback(ctx, d_out, h_dev, geom, ctb_log2, (const p265_sao_ctb *)d_par);
    else rc = copy_planes_to_host(ctx, out, d_out, geom, bytes);
    if (rc) return rc;
    trace_mark(ctx, 2, 3);
    return finish(ctx);
}

int p265_reconstruct_batch_dev(p265_ctx *ctx, const void *d_pred, const int16_t *d_residual, void *d_rec,
                               const p265_pic_geom *geom) {
    if (!ctx || !d_pred || !d_residual || !d_rec) return set_error(P265_EINVAL, "p265_reconstruct_batch_dev: NULL argument");
    int rc = check_geom(geom, 16);  // 16-byte rows for both the uint8 and the int16 planes
    if (rc) return rc;
    P265_CUDA(cudaSetDevice(ctx->device));
    return launch_recon(ctx, d_pred, d_residual, d_rec, geom);
}

int p265_reconstruct_batch(p265_ctx *ctx, const void *pred, const int16_t *residual, void *rec,
                           const p265_pic_geom *geom) {
    if (!ctx || !pred || !residual || !rec) return set_error(P265_EINVAL, "p265_reconstruct_batch: NULL argument");
    int rc = check_geom(geom, 16);
    if (rc) return rc;
    P265_CUDA(cudaSetDevice(ctx->device));
    const int bytes = (geom->bit_depth_y > 8 || geom->bit_depth_c > 8) ? 2 : 1;
    const size_t elems = (size_t)geom->pic_stride * geom->n_pics;
    void *d_pred, *d_res, *d_rec;
    if ((rc = ensure(ctx, 2, elems * bytes, &d_pred))) return rc;
    if ((rc = ensure(ctx, 4, elems * 2, &d_res))) return rc;
    if ((rc = ensure(ctx, 6, elems * bytes, &d_rec))) return rc;
    P265_CUDA(cudaMemcpyAsync(d_pred, pred, elems * bytes, cudaMemcpyHostToDevice, ctx->stream));
    P265_CUDA(cudaMemcpyAsync(d_res, residual, elems * 2, cudaMemcpyHostToDevice, ctx->stream));
    P265_CUDA(cudaMemcpyAsync(d_rec, d_pred, elems * bytes, cudaMemcpyDeviceToDevice, ctx->stream));  // padding
    if ((rc = launch_recon(ctx, d_pred, (const int16_t *)d_res, d_rec, geom))) return rc;
    P265_CUDA(cudaMemcpyAsync(rec, d_rec, elems * bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return finish(ctx);
}

static int check_deblock(const p265_pic_geom *geom, int ctb_log2, int *bytes) {
    *bytes = (geom->bit_depth_y > 8 || geom->bit_depth_c > 8) ? 2 : 1;
    int rc = check_geom(geom, 8 / *bytes);   // rows start on 8-byte boundaries (half-block loads)
    if (rc) return rc;
    if (ctb_log2 < 4 || ctb_log2 > 6) return set_error(P265_EINVAL, "ctb_log2 %d outside 4..6", ctb_log2);
    if (geom->width % 8 || geom->height % 8)
        return set_error(P265_EINVAL, "picture size must be a multiple of MinCbSize 8 (got %dx%d)", geom->width,
                         geom->height);
    return P265_OK;
}

static int check_deblock_maps(const p265_pic_geom *geom, int ctb_log2, const p265_dbk_blk *blk, const p265_dbk_ctb *ctb,
                              size_t *n_blk_out, size_t *n_ctb_out) {
    const int cs = 1 << ctb_log2;
    const size_t n_blk = (size_t)(geom->width / 8) * (geom->height / 8) * geom->n_pics;
    const size_t n_ctb = (size_t)((geom->width + cs - 1) / cs) * ((geom->height + cs - 1) / cs) * geom->n_pics;
    for (size_t i = 0; i < n_blk; i++)
        if ((blk[i] & 3) == 3 || ((blk[i] >> 2) & 3) == 3 || ((blk[i] >> 4) & 3) == 3 || ((blk[i] >> 6) & 3) == 3)
            return set_error(P265_EINVAL, "edge map entry %zu holds a boundary strength of 3", i);
    for (size_t i = 0; i < n_ctb; i++)
        if (ctb[i].beta_offset_div2 < -6 || ctb[i].beta_offset_div2 > 6 || ctb[i].tc_offset_div2 < -6 ||
            ctb[i].tc_offset_div2 > 6 || ctb[i].cb_qp_offset < -12 || ctb[i].cb_qp_offset > 12 ||
            ctb[i].cr_qp_offset < -12 || ctb[i].cr_qp_offset > 12)
            return set_error(P265_EINVAL, "deblocking parameters of CTB %zu out of range", i);
    *n_blk_out = n_blk;
    *n_ctb_out = n_ctb;
    return P265_OK;
}

int p265_deblock_batch_dev(p265_ctx *ctx, void *d_planes, const p265_pic_geom *geom, int ctb_log2,
                           const p265_dbk_blk *d_blk, const p265_dbk_ctb *d_ctb) {
    if (!ctx || !d_planes || !d_blk || !d_ctb || !geom)
        return set_error(P265_EINVAL, "p265_deblock_batch_dev: NULL argument");
    int bytes;
    int rc = check_deblock(geom, ctb_log2, &bytes);
    if (rc) return rc;
    P265_CUDA(cudaSetDevice(ctx->device));
    return launch_deblock(ctx, d_planes, geom, ctb_log2, d_blk, d_ctb);
}

int p265_deblock_batch(p265_ctx *ctx, void *planes, const p265_pic_geom *geom, int ctb_log2, const p265_dbk_blk *blk,
                       const p265_dbk_ctb *ctb) {
    if (!ctx || !planes || !blk || !ctb || !geom) return set_error(P265_EINVAL, "p265_deblock_batch: NULL argument");
    int bytes;
    int rc = check_deblock(geom, ctb_log2, &bytes);
    if (rc) return rc;
    size_t n_blk, n_ctb;
    if ((rc = check_deblock_maps(geom, ctb_log2, blk, ctb, &n_blk, &n_ctb))) return rc;
    P265_CUDA(cudaSetDevice(ctx->device));
    const size_t plane_bytes = (size_t)bytes * geom->pic_stride * geom->n_pics;
    void *d_pix, *d_blk, *d_ctb;
    if ((rc = ensure(ctx, 2, plane_bytes, &d_pix))) return rc;
    if ((rc = ensure(ctx, 0, sizeof(p265_dbk_blk) * n_blk, &d_blk))) return rc;
    if ((rc = ensure(ctx, 7, sizeof(p265_dbk_ctb) * n_ctb, &d_ctb))) return rc;
    P265_CUDA(cudaMemcpyAsync(d_pix, planes, plane_bytes, cudaMemcpyHostToDevice, ctx->stream));
    P265_CUDA(cudaMemcpyAsync(d_blk, blk, sizeof(p265_dbk_blk) * n_blk, cudaMemcpyHostToDevice, ctx->stream));
    P265_CUDA(cudaMemcpyAsync(d_ctb, ctb, sizeof(p265_dbk_ctb) * n_ctb, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = launch_deblock(ctx, d_pix, geom, ctb_log2, (const p265_dbk_blk *)d_blk, (const p265_dbk_ctb *)d_ctb)))
        return rc;
    P265_CUDA(cudaMemcpyAsync(planes, d_pix, plane_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return finish(ctx);
}

int p265_loop_filter_batch(p265_ctx *ctx, void *planes, const p265_pic_geom *geom, int ctb_log2, const p265_dbk_blk *blk,
                           const p265_dbk_ctb *dbk_ctb, const p265_sao_ctb *sao, const uint8_t *no_filter) {
    if (!ctx || !planes || !geom) return set_error(P265_EINVAL, "p265_loop_filter_batch: NULL argument");
    if ((blk == nullptr) != (dbk_ctb == nullptr))
        return set_error(P265_EINVAL, "p265_loop_filter_batch: edge map and per-CTB deblocking parameters go together");
    if (!blk && !sao) return set_error(P265_EINVAL, "p265_loop_filter_batch: neither deblocking nor SAO requested");
    int bytes;
    int rc = blk ? check_deblock(geom, ctb_log2, &bytes) : P265_OK;
    if (!rc && sao) rc = check_sao(geom, ctb_log2, &bytes);
    if (rc) return rc;
    size_t n_blk = 0, n_ctb = 0, n_sao = 0;
    if (blk && (rc = check_deblock_maps(geom, ctb_log2, blk, dbk_ctb, &n_blk, &n_ctb))) return rc;
    if (sao && (rc = check_sao_params(geom, ctb_log2, sao, &n_sao))) return rc;
    P265_CUDA(cudaSetDevice(ctx->device));
    const size_t plane_bytes = (size_t)bytes * geom->pic_stride * geom->n_pics;
    const size_t nf_bytes = (size_t)((geom->width + 7) / 8) * ((geom->height + 7) / 8) * geom->n_pics;
    void *d_pix, *d_out = nullptr, *d_blk = nullptr, *d_ctb = nullptr, *d_par = nullptr, *d_nf = nullptr;
    if ((rc = ensure(ctx, 2, plane_bytes, &d_pix))) return rc;
    P265_CUDA(cudaMemcpyAsync(d_pix, planes, plane_bytes, cudaMemcpyHostToDevice, ctx->stream));  // the one H2D of the planes
    if (blk) {
        if ((rc = ensure(ctx, 0, sizeof(p265_dbk_blk) * n_blk, &d_blk))) return rc;
        if ((rc = ensure(ctx, 7, sizeof(p265_dbk_ctb) * n_ctb, &d_ctb))) return rc;
        P265_CUDA(cudaMemcpyAsync(d_blk, blk, sizeof(p265_dbk_blk) * n_blk, cudaMemcpyHostToDevice, ctx->stream));
        P265_CUDA(cudaMemcpyAsync(d_ctb, dbk_ctb, sizeof(p265_dbk_ctb) * n_ctb, cudaMemcpyHostToDevice, ctx->stream));
        if ((rc = launch_deblock(ctx, d_pix, geom, ctb_log2, (const p265_dbk_blk *)d_blk, (const p265_dbk_ctb *)d_ctb)))
            return rc;
    }
    const void *d_final = d_pix;
    if (sao) {
        if ((rc = ensure(ctx, 6, plane_bytes, &d_out))) return rc;
        if ((rc = ensure(ctx, 10, sizeof(p265_sao_ctb) * n_sao, &d_par))) return rc;
        P265_CUDA(cudaMemcpyAsync(d_par, sao, sizeof(p265_sao_ctb) * n_sao, cudaMemcpyHostToDevice, ctx->stream));
        if (no_filter) {
            if ((rc = ensure(ctx, 11, nf_bytes, &d_nf))) return rc;
            P265_CUDA(cudaMemcpyAsync(d_nf, no_filter, nf_bytes, cudaMemcpyHostToDevice, ctx->stream));
        }
        if ((rc = launch_sao(ctx, d_pix, d_out, geom, ctb_log2, (const p265_sao_ctb *)d_par, (const uint8_t *)d_nf)))
            return rc;
        d_final = d_out;
    }
    // the one copy back.  Without deblocking the host still holds every sample SAO leaves alone (see p265_sao_batch)
    void *h_dev = (!blk && sao) ? mapped_host_pointer(planes) : nullptr;
    if (h_dev) rc = launch_sao_writeback(ctx, d_final, h_dev, geom, ctb_log2, (const p265_sao_ctb *)d_par);
    else rc = copy_planes_to_host(ctx, planes, d_final, geom, bytes);
    if (rc) return rc;
    return finish(ctx);
}

int p265_pcie_probe(p265_ctx *ctx, size_t bytes, int n_buffers, int reps, double *h2d_bytes_per_s,
                    double *d2h_bytes_per_s) {
    if (!ctx || (!h2d_bytes_per_s && !d2h_bytes_per_s)) return set_error(P265_EINVAL, "p265_pcie_probe: NULL argument");
    P265_CUDA(cudaSetDevice(ctx->device));
    return run_pcie_probe(ctx, bytes, n_buffers, reps, h2d_bytes_per_s, d2h_bytes_per_s);
}

int p265_int_peak(p265_ctx *ctx, int kind, double *ops_per_s, double *ms) {
    if (!ctx || !ops_per_s || !ms) return set_error(P265_EINVAL, "p265_int_peak: NULL argument");
    P265_CUDA(cudaSetDevice(ctx->device));
    return run_int_peak(ctx, kind, ops_per_s, ms);
}

#pragma GCC visibility pop
}  // extern "C"
