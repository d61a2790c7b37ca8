/*
 * TEST INFRASTRUCTURE ONLY -- plain-C restatement of the residual + SAO path.
 *
 * Second, independent restatement next to oracle/spec_oracle.py (numpy); the two are
 * cross-checked in tests/test_oracle_c.py and both are pinned against the reference's
 * own scaling.py / transform.py outputs (tests/golden, .npz files).  It works on the same
 * packed structures as the C-ABI (include/p265_b200.h) so that full-size 4K batches
 * can be compared with the GPU in seconds, and bench.py times it (all host threads,
 * pthreads; no OpenMP runtime in this image) as the "port" CPU baseline.  The product library never links it.
 *
 * Follows: scaling.py:23-47 (8.6.3), transform.py:89-106 structure with the standard's
 * orientation / stage order / final shift (8.6.2, 8.6.4.1-2), 8.6.2 transform-skip and
 * bypass, 8.7.3 SAO.  Deliberately naive: full N x N matrix products, int64
 * arithmetic, one sample at a time.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "p265_b200.h"


/* ------------------------------------------------------------- tiny parallel-for */
static int g_threads = 0; /* 0 = all online cores */
void oracle_set_threads(int n) { g_threads = n; }
int oracle_get_threads(void) {
    if (g_threads > 0) return g_threads;
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
typedef void (*range_fn)(int64_t lo, int64_t hi, void *arg);
typedef struct { range_fn fn; void *arg; int64_t lo, hi; } job_t;
static void *job_main(void *p) { job_t *j = (job_t *)p; j->fn(j->lo, j->hi, j->arg); return NULL; }
static void parallel_for(int64_t n, range_fn fn, void *arg) {
    int nt = oracle_get_threads();
    if (nt > n) nt = n > 0 ? (int)n : 1;
    if (nt <= 1) { fn(0, n, arg); return; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nt);
    job_t *jobs = (job_t *)malloc(sizeof(job_t) * nt);
    for (int i = 0; i < nt; i++) {
        jobs[i].fn = fn; jobs[i].arg = arg;
        jobs[i].lo = n * i / nt; jobs[i].hi = n * (i + 1) / nt;
        pthread_create(&th[i], NULL, job_main, &jobs[i]);
    }
    for (int i = 0; i < nt; i++) pthread_join(th[i], NULL);
    free(th); free(jobs);
}

static int g_dct[32][32];
static int g_init;
static const int g_dst[4][4] = {{29, 55, 74, 84}, {74, 74, 0, -74}, {84, -29, -74, 55}, {55, -84, 74, -29}};
static const int g_level_scale[6] = {40, 45, 51, 57, 64, 72};
static const int g_sf_off[4] = {0, 96, 480, 2016};

static void init_tables(void) {
    /* 8.6.4.2 matrix from its cosine structure: |entry| = mag[k], k = folded angle */
    static const int mag[33] = {64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67, 64,
                                61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9,  4,  0};
    if (g_init) return;
    for (int j = 0; j < 32; j++)
        for (int i = 0; i < 32; i++) {
            int a = (j * (2 * i + 1)) % 128;
            if (a > 64) a = 128 - a;
            g_dct[j][i] = j == 0 ? 64 : (a <= 32 ? mag[a] : -mag[64 - a]);
        }
    g_init = 1;
}

void oracle_dct32(int32_t *out) {
    init_tables();
    for (int j = 0; j < 32; j++)
        for (int i = 0; i < 32; i++) out[j * 32 + i] = g_dct[j][i];
}

static inline int64_t clip3(int64_t lo, int64_t hi, int64_t v) { return v < lo ? lo : (v > hi ? hi : v); }

static inline int coef(int n, int dst, int j, int i) { return dst ? g_dst[j][i] : g_dct[j * (32 / n)][i]; }

static int matrix_id(int log2n, int c_idx, int intra) {
    if (log2n == 5) return intra ? 0 : 1;
    return intra ? c_idx : c_idx + 3;
}

/* d[y][x] (8.6.3) for one TB */
static void dequant_tb(const p265_tu_desc *t, const int16_t *lv, const uint8_t *sf, int bit_depth,
                       int64_t *d) {
    int n = 1 << t->log2n;
    int bd_shift = bit_depth + t->log2n - 5;
    int64_t scale = (int64_t)g_level_scale[t->qp % 6] << (t->qp / 6);
    const uint8_t *m = NULL;
    if (sf) m = sf + g_sf_off[t->log2n - 2] + matrix_id(t->log2n, t->c_idx, (t->flags & P265_TU_INTRA) != 0) * n * n;
    for (int i = 0; i < n * n; i++) {
        int64_t mm = m ? m[i] : 16;
        d[i] = clip3(-32768, 32767, ((int64_t)lv[i] * mm * scale + ((int64_t)1 << (bd_shift - 1))) >> bd_shift);
    }
}

/* r[y][x] (8.6.2 + 8.6.4), saturated to int16 (see spec_oracle.py:sat16) */
static void residual_tb(const p265_tu_desc *t, const int16_t *lv, const uint8_t *sf, int bit_depth,
                        int64_t *r) {
    int n = 1 << t->log2n;
    int64_t d[1024], e[1024], g[1024];
    if (t->flags & P265_TU_BYPASS) {
        for (int i = 0; i < n * n; i++) r[i] = lv[i];
        return;
    }
    dequant_tb(t, lv, sf, bit_depth, d);
    int bd2 = 20 - bit_depth;
    if (t->flags & P265_TU_SKIP) {
        for (int i = 0; i < n * n; i++) r[i] = ((d[i] * 128) + ((int64_t)1 << (bd2 - 1))) >> bd2;
        return;
    }
    int dst = (t->flags & P265_TU_DST) != 0;
    for (int x = 0; x < n; x++)          /* stage 1: every column */
        for (int i = 0; i < n; i++) {
            int64_t s = 0;
            for (int j = 0; j < n; j++) s += (int64_t)coef(n, dst, j, i) * d[j * n + x];
            e[i * n + x] = s;
        }
    for (int i = 0; i < n * n; i++) g[i] = clip3(-32768, 32767, (e[i] + 64) >> 7);
    for (int y = 0; y < n; y++)          /* stage 2: every row */
        for (int i = 0; i < n; i++) {
            int64_t s = 0;
            for (int j = 0; j < n; j++) s += (int64_t)coef(n, dst, j, i) * g[y * n + j];
            r[y * n + i] = (s + ((int64_t)1 << (bd2 - 1))) >> bd2;
        }
}

typedef struct {
    const p265_tu_desc *tus; const int16_t *coeffs; const uint8_t *sf; const p265_pic_geom *g;
    int16_t *out; int bdy, bdc;
} res_job;

static void residual_range(int64_t lo, int64_t hi, void *arg) {
    res_job *j = (res_job *)arg;
    const p265_pic_geom *g = j->g;
    for (int64_t k = lo; k < hi; k++) {
        const p265_tu_desc *t = &j->tus[k];
        int n = 1 << t->log2n;
        int64_t r[1024];
        int bd = t->c_idx ? g->bit_depth_c : g->bit_depth_y;
        int stride = t->c_idx ? g->stride_c : g->stride_y;
        residual_tb(t, j->coeffs + (size_t)t->coeff_off * 16, j->sf, bd, r);
        int16_t *dst = j->out + (size_t)t->pic * g->pic_stride + g->plane_off[t->c_idx] + (size_t)t->y * stride + t->x;
        for (int y = 0; y < n; y++)
            for (int x = 0; x < n; x++) dst[(size_t)y * stride + x] = (int16_t)clip3(-32768, 32767, r[y * n + x]);
    }
}

int oracle_residual_batch(const p265_tu_desc *tus, int32_t n_tus, const int16_t *coeffs,
                          const uint8_t *sf, const p265_pic_geom *g, int16_t *out, int flags) {
    init_tables();
    if (flags & P265_RES_ZERO_FILL) memset(out, 0, sizeof(int16_t) * (size_t)g->pic_stride * g->n_pics);
    res_job j = {tus, coeffs, sf, g, out, 0, 0};
    parallel_for(n_tus, residual_range, &j);
    return 0;
}

static void dequant_range(int64_t lo, int64_t hi, void *arg) {
    res_job *j = (res_job *)arg;
    for (int64_t k = lo; k < hi; k++) {
        const p265_tu_desc *t = &j->tus[k];
        int n = 1 << t->log2n;
        int64_t d[1024];
        size_t off = (size_t)t->coeff_off * 16;
        dequant_tb(t, j->coeffs + off, j->sf, t->c_idx ? j->bdc : j->bdy, d);
        for (int i = 0; i < n * n; i++) j->out[off + i] = (int16_t)d[i];
    }
}

int oracle_dequant_batch(const p265_tu_desc *tus, int32_t n_tus, const int16_t *coeffs, const uint8_t *sf,
                         int bit_depth_y, int bit_depth_c, int16_t *scaled) {
    init_tables();
    res_job j = {tus, coeffs, sf, NULL, scaled, bit_depth_y, bit_depth_c};
    parallel_for(n_tus, dequant_range, &j);
    return 0;
}

/* transform.py:89-109 as written: out[x][y], in d[y][x] */
int oracle_ref_literal_batch(const p265_tu_desc *tus, int32_t n_tus, const int16_t *scaled, int32_t *out) {
    init_tables();
    for (int32_t k = 0; k < n_tus; k++) {
        const p265_tu_desc *t = &tus[k];
        int n = 1 << t->log2n;
        int dst = (n == 4 && t->c_idx == 0);
        size_t off = (size_t)t->coeff_off * 16;
        const int16_t *d = scaled + off;
        int64_t gl[32], row[32];
        /* C[i][j] = dst ? DST[i][j] : DCT32[i][j * 32 / n];  e[:, col] = C . d_xy[:, col];
         * only col = n-1 survives (stale loop variable, transform.py:108-109)          */
        for (int i = 0; i < n; i++) {
            int64_t s = 0;
            for (int j = 0; j < n; j++) {
                int c = dst ? g_dst[i][j] : g_dct[i][j * (32 / n)];
                s += (int64_t)c * d[(n - 1) * n + j]; /* d_xy[j][n-1] == d_yx[n-1][j] */
            }
            gl[i] = clip3(-32768, 32767, (s + 64) >> 7);
        }
        for (int i = 0; i < n; i++) {
            int64_t s = 0;
            for (int j = 0; j < n; j++) {
                int c = dst ? g_dst[i][j] : g_dct[i][j * (32 / n)];
                s += (int64_t)c * gl[j];
            }
            row[i] = s;
        }
        for (int x = 0; x < n; x++)
            for (int y = 0; y < n; y++) out[off + (size_t)x * n + y] = (int32_t)row[y];
    }
    return 0;
}

/* ---------------------------------------------------------------------------- SAO */
static inline int sample_at(const void *p, int bytes, size_t i) {
    return bytes == 1 ? ((const uint8_t *)p)[i] : ((const uint16_t *)p)[i];
}
static inline void sample_put(void *p, int bytes, size_t i, int v) {
    if (bytes == 1) ((uint8_t *)p)[i] = (uint8_t)v;
    else ((uint16_t *)p)[i] = (uint16_t)v;
}

typedef struct {
    const void *rec; void *out; const p265_pic_geom *g; int ctb_log2; const p265_sao_ctb *params;
    const uint8_t *no_filter; int p, c;
} sao_job;

static void sao_rows(int64_t lo, int64_t hi, void *arg) {
    static const int hpos[4][2] = {{-1, 1}, {0, 0}, {-1, 1}, {1, -1}};
    static const int vpos[4][2] = {{0, 0}, {-1, 1}, {-1, 1}, {-1, 1}};
    static const int remap[5] = {1, 2, 0, 3, 4};
    sao_job *j = (sao_job *)arg;
    const p265_pic_geom *g = j->g;
    int p = j->p, c = j->c;
    int bytes = (g->bit_depth_y > 8 || g->bit_depth_c > 8) ? 2 : 1;
    int ctb = 1 << j->ctb_log2;
    int ctbs_w = (g->width + ctb - 1) / ctb, ctbs_h = (g->height + ctb - 1) / ctb;
    int w8 = (g->width + 7) / 8, h8 = (g->height + 7) / 8;
    int w = c ? g->width / 2 : g->width, h = c ? g->height / 2 : g->height;
    int stride = c ? g->stride_c : g->stride_y;
    int cs = c ? ctb / 2 : ctb; /* CTB size in this plane */
    int bd = c ? g->bit_depth_c : g->bit_depth_y;
    int maxv = (1 << bd) - 1;
    size_t base = (size_t)p * g->pic_stride + g->plane_off[c];
    for (int y = (int)lo; y < (int)hi; y++)
        for (int x = 0; x < w; x++) {
            int ry = y / cs, rx = x / cs;
            const p265_sao_ctb *q = &j->params[((size_t)p * ctbs_h + ry) * ctbs_w + rx];
            int v = sample_at(j->rec, bytes, base + (size_t)y * stride + x);
            int idx = 0;
            int skip = 0;
            if (j->no_filter) {
                int by = c ? y / 4 : y / 8, bx = c ? x / 4 : x / 8;
                skip = j->no_filter[((size_t)p * h8 + by) * w8 + bx] != 0;
            }
            if (q->type[c] == 1 && !skip) {
                int band = v >> (bd - 5);
                int k = (band - q->band_pos[c]) & 31;
                idx = k < 4 ? k + 1 : 0;
            } else if (q->type[c] == 2 && !skip) {
                int cls = q->eo_class[c];
                int e = 2;
                for (int k = 0; k < 2; k++) {
                    int ny = y + vpos[cls][k], nx = x + hpos[cls][k];
                    if (ny < 0 || ny >= h || nx < 0 || nx >= w) { e = -1; break; }
                    int bit = (ny / cs - ry + 1) * 3 + (nx / cs - rx + 1);
                    if (bit != 4 && !((q->avail >> bit) & 1)) { e = -1; break; }
                    int nv = sample_at(j->rec, bytes, base + (size_t)ny * stride + nx);
                    e += (v > nv) - (v < nv);
                }
                idx = e < 0 ? 0 : remap[e];
            }
            int o = idx ? q->offset_val[c][idx - 1] : 0;
            sample_put(j->out, bytes, base + (size_t)y * stride + x, (int)clip3(0, maxv, v + o));
        }
}

int oracle_sao_batch(const void *rec, void *out, const p265_pic_geom *g, int ctb_log2,
                     const p265_sao_ctb *params, const uint8_t *no_filter) {
    for (int p = 0; p < g->n_pics; p++)
        for (int c = 0; c < 3; c++) {
            sao_job j = {rec, out, g, ctb_log2, params, no_filter, p, c};
            parallel_for(c ? g->height / 2 : g->height, sao_rows, &j);
        }
    return 0;
}

/* --------------------------------------------------------------------- deblocking */
/* H.265 8.7.2 restated from the standard (the reference has no deblocking filter; only its
 * control flags are parsed, pps.py:121-131, slice.py:170-179).  Classic two passes over the
 * whole picture: every vertical edge, then every horizontal edge on the result.  Pinned by
 * the libavcodec decode of sanity.bin (tests/test_decode_sanity.py).                       */
static const uint8_t k_beta[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15,
                                   16, 17, 18, 20, 22, 24, 26, 28, 30, 32, 34, 36, 38, 40, 42, 44, 46, 48, 50, 52,
                                   54, 56, 58, 60, 62, 64};
static const uint8_t k_tc[54] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2,
                                 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 5, 5, 6, 6, 7, 8, 9, 10, 11, 13, 14, 16, 18, 20, 22, 24};
static int qpc_of(int qpi) {
    static const int t[14] = {29, 30, 31, 32, 33, 33, 34, 34, 35, 35, 36, 36, 37, 37};
    return qpi < 30 ? qpi : (qpi >= 44 ? qpi - 6 : t[qpi - 30]);
}

typedef struct {
    void *pix; const p265_pic_geom *g; int ctb_log2; const p265_dbk_blk *blk; const p265_dbk_ctb *ctb;
    int p, c, vertical;
} dbk_job;

static inline int blk_qp(uint16_t e) { int q = (e >> P265_DBK_QP_SHIFT) & 0x7f; return q >= 64 ? q - 128 : q; }

/* one 4-sample segment; (x, y) first q0 sample in the plane of component c */
static void dbk_segment(const dbk_job *j, int x, int y) {
    const p265_pic_geom *g = j->g;
    int c = j->c, v = j->vertical;
    int bytes = (g->bit_depth_y > 8 || g->bit_depth_c > 8) ? 2 : 1;
    int w8 = g->width / 8, h8 = g->height / 8;
    int ctb = 1 << j->ctb_log2;
    int ctbs_w = (g->width + ctb - 1) / ctb, ctbs_h = (g->height + ctb - 1) / ctb;
    int xl = c ? x * 2 : x, yl = c ? y * 2 : y;             /* luma position of q0 */
    int xp = v ? xl - 1 : xl, yp = v ? yl : yl - 1;         /* luma position of p0 */
    const p265_dbk_blk *bp = j->blk + (size_t)j->p * w8 * h8;
    uint16_t eq = bp[(yl >> 3) * w8 + (xl >> 3)], ep = bp[(yp >> 3) * w8 + (xp >> 3)];
    int bs = v ? (eq >> ((yl & 4) ? P265_DBK_BS_V1 : P265_DBK_BS_V0)) & 3
               : (eq >> ((xl & 4) ? P265_DBK_BS_H1 : P265_DBK_BS_H0)) & 3;
    if (bs == 0 || (c && bs != 2)) return;
    const p265_dbk_ctb *par = &j->ctb[((size_t)j->p * ctbs_h + (yl >> j->ctb_log2)) * ctbs_w + (xl >> j->ctb_log2)];
    int no_p = (ep & P265_DBK_NO_FILTER) != 0, no_q = (eq & P265_DBK_NO_FILTER) != 0;
    int qpl = (blk_qp(eq) + blk_qp(ep) + 1) >> 1;
    int stride = c ? g->stride_c : g->stride_y;
    int bd = c ? g->bit_depth_c : g->bit_depth_y;
    int maxv = (1 << bd) - 1;
    size_t base = (size_t)j->p * g->pic_stride + g->plane_off[c];
    ptrdiff_t across = v ? 1 : stride, along = v ? stride : 1;
    size_t q0 = base + (size_t)y * stride + x;
#define PX(side, i, k) sample_at(j->pix, bytes, (size_t)((ptrdiff_t)q0 + (side) * across * ((i) + ((side) < 0)) + (k) * along))
#define P(i, k) PX(-1, i, k)
#define Q(i, k) PX(1, i, k)
#define PUTP(i, k, val) sample_put(j->pix, bytes, (size_t)((ptrdiff_t)q0 - across * ((i) + 1) + (k) * along), (int)(val))
#define PUTQ(i, k, val) sample_put(j->pix, bytes, (size_t)((ptrdiff_t)q0 + across * (i) + (k) * along), (int)(val))
    if (c) {
        int qpc = qpc_of(qpl + (c == 1 ? par->cb_qp_offset : par->cr_qp_offset));
        int tc = k_tc[clip3(0, 53, qpc + 2 + (par->tc_offset_div2 * 2))] << (bd - 8);
        for (int k = 0; k < 4; k++) {
            int p0 = P(0, k), p1 = P(1, k), q0v = Q(0, k), q1 = Q(1, k);
            int d = (int)clip3(-tc, tc, ((((q0v - p0) * 4) + p1 - q1 + 4) >> 3));
            if (!no_p) PUTP(0, k, clip3(0, maxv, p0 + d));
            if (!no_q) PUTQ(0, k, clip3(0, maxv, q0v - d));
        }
        return;
    }
    int beta = k_beta[clip3(0, 51, qpl + par->beta_offset_div2 * 2)] << (bd - 8);
    int tc = k_tc[clip3(0, 53, qpl + 2 * (bs - 1) + par->tc_offset_div2 * 2)] << (bd - 8);
    int dp0 = abs(P(2, 0) - 2 * P(1, 0) + P(0, 0)), dp3 = abs(P(2, 3) - 2 * P(1, 3) + P(0, 3));
    int dq0 = abs(Q(2, 0) - 2 * Q(1, 0) + Q(0, 0)), dq3 = abs(Q(2, 3) - 2 * Q(1, 3) + Q(0, 3));
    int dpq0 = dp0 + dq0, dpq3 = dp3 + dq3, dp = dp0 + dp3, dq = dq0 + dq3;
    if (dpq0 + dpq3 >= beta) return;
    int s0 = 2 * dpq0 < (beta >> 2) && abs(P(3, 0) - P(0, 0)) + abs(Q(0, 0) - Q(3, 0)) < (beta >> 3) &&
             abs(P(0, 0) - Q(0, 0)) < ((5 * tc + 1) >> 1);
    int s3 = 2 * dpq3 < (beta >> 2) && abs(P(3, 3) - P(0, 3)) + abs(Q(0, 3) - Q(3, 3)) < (beta >> 3) &&
             abs(P(0, 3) - Q(0, 3)) < ((5 * tc + 1) >> 1);
    int side = (beta + (beta >> 1)) >> 3;
    int dep = dp < side, deq = dq < side;
    for (int k = 0; k < 4; k++) {
        int p0 = P(0, k), p1 = P(1, k), p2 = P(2, k), p3 = P(3, k);
        int q0v = Q(0, k), q1 = Q(1, k), q2 = Q(2, k), q3 = Q(3, k);
        if (s0 && s3) {
            if (!no_p) {
                PUTP(0, k, clip3(p0 - 2 * tc, p0 + 2 * tc, (p2 + 2 * p1 + 2 * p0 + 2 * q0v + q1 + 4) >> 3));
                PUTP(1, k, clip3(p1 - 2 * tc, p1 + 2 * tc, (p2 + p1 + p0 + q0v + 2) >> 2));
                PUTP(2, k, clip3(p2 - 2 * tc, p2 + 2 * tc, (2 * p3 + 3 * p2 + p1 + p0 + q0v + 4) >> 3));
            }
            if (!no_q) {
                PUTQ(0, k, clip3(q0v - 2 * tc, q0v + 2 * tc, (p1 + 2 * p0 + 2 * q0v + 2 * q1 + q2 + 4) >> 3));
                PUTQ(1, k, clip3(q1 - 2 * tc, q1 + 2 * tc, (p0 + q0v + q1 + q2 + 2) >> 2));
                PUTQ(2, k, clip3(q2 - 2 * tc, q2 + 2 * tc, (p0 + q0v + q1 + 3 * q2 + 2 * q3 + 4) >> 3));
            }
        } else {
            int d = (9 * (q0v - p0) - 3 * (q1 - p1) + 8) >> 4;
            if (abs(d) >= tc * 10) continue;
            d = (int)clip3(-tc, tc, d);
            if (!no_p) {
                PUTP(0, k, clip3(0, maxv, p0 + d));
                if (dep) PUTP(1, k, clip3(0, maxv, p1 + clip3(-(tc >> 1), tc >> 1, (((p2 + p0 + 1) >> 1) - p1 + d) >> 1)));
            }
            if (!no_q) {
                PUTQ(0, k, clip3(0, maxv, q0v - d));
                if (deq) PUTQ(1, k, clip3(0, maxv, q1 + clip3(-(tc >> 1), tc >> 1, (((q2 + q0v + 1) >> 1) - q1 - d) >> 1)));
            }
        }
    }
#undef PX
#undef P
#undef Q
#undef PUTP
#undef PUTQ
}

/* unit u = one row of 4-sample segments (vertical pass) or one edge row (horizontal pass) */
static void dbk_units(int64_t lo, int64_t hi, void *arg) {
    const dbk_job *j = (const dbk_job *)arg;
    int w = j->c ? j->g->width / 2 : j->g->width, h = j->c ? j->g->height / 2 : j->g->height;
    for (int64_t u = lo; u < hi; u++) {
        if (j->vertical) {
            for (int x = 8; x < w; x += 8) dbk_segment(j, x, (int)u * 4);
        } else {
            int y = ((int)u + 1) * 8;
            if (y >= h) continue;
            for (int x = 0; x < w; x += 4) dbk_segment(j, x, y);
        }
    }
}

int oracle_deblock_batch(void *planes, const p265_pic_geom *g, int ctb_log2, const p265_dbk_blk *blk,
                         const p265_dbk_ctb *ctb) {
    if (g->width % 8 || g->height % 8) return -1;
    for (int p = 0; p < g->n_pics; p++)
        for (int vertical = 1; vertical >= 0; vertical--)
            for (int c = 0; c < 3; c++) {
                dbk_job j = {planes, g, ctb_log2, blk, ctb, p, c, vertical};
                int h = c ? g->height / 2 : g->height;
                parallel_for(vertical ? h / 4 : h / 8, dbk_units, &j);
            }
    return 0;
}
