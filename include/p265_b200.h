/*
 * p265_b200 -- C-ABI of the B200 (sm_100a) residual + SAO path.
 *
 * Reference: jacke121/p265 is pure Python and has no FFI; the interface this
 * library replaces is the Python function surface of
 *     decoder/scaling.py:4     inverse_scaling(pu, x0, y0, log2size)
 *     decoder/transform.py:89  inverse_transform(pu, x0, y0, log2size)
 *     decoder/sao.py:4         class Sao   (per-CTB parameters; the filter itself
 *                              does not exist in the reference)
 * re-cut at picture granularity: the host packs every coded TB of
 * a batch of pictures into a descriptor list + coefficient arena and one call
 * produces the residual planes; a second call applies SAO to reconstructed planes.
 * p265_b200/{_lib,engine}.py bind these entry points with ctypes; the drop-in modules
 * p265_b200/dropin/{scaling,transform,reconstruction,sao,sld}.py keep the reference's
 * per-TB names on top.
 *
 * All entry points return 0 on success or a negative p265_status; the message of
 * the last failure on the calling thread is available from p265_last_error().
 * There is no CPU fallback: without a CUDA device p265_ctx_create() fails.
 *
 * Thread-safety: one p265_ctx per host thread / GPU; calls on one context are
 * serialised by the caller; every context owns (or borrows) exactly one stream.
 */
#ifndef P265_B200_H
#define P265_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define P265_ABI_VERSION 3

typedef enum p265_status {
    P265_OK = 0,
    P265_EINVAL = -1,  /* bad argument (Python wrapper raises ValueError)   */
    P265_ECUDA = -2,   /* CUDA runtime failure (RuntimeError)               */
    P265_ENOMEM = -3   /* device or pinned allocation failed (RuntimeError) */
} p265_status;

typedef struct p265_ctx p265_ctx;

/* One coded transform block.  Replaces the per-coefficient
 * tu.get_trans_coeff_level(x, y, c_idx) walk (tu.py:667-684) and the per-call
 * pu / cu / sps attribute reads of scaling.py:13-44.                               */
typedef struct p265_tu_desc {
    uint16_t x, y;       /* top-left sample in the component's plane                */
    uint8_t log2n;       /* 2..5                                                    */
    uint8_t c_idx;       /* 0 Y, 1 Cb, 2 Cr                                         */
    uint8_t qp;          /* qP including QpBdOffset (scaling.py:13-18)              */
    uint8_t flags;       /* P265_TU_*                                               */
    uint32_t coeff_off;  /* offset into the coefficient arena, units of 16 coeffs   */
    uint16_t pic;        /* picture index inside the batch                          */
    uint16_t rsvd;       /* bits 0-10  packed stream: number of non-zero levels in the
                                       TB's record (= set bits of its bitmap); dense
                                       arena: ignored
                            bits 11-12 zr, bits 13-14 zc: zero-extent codes, below     */
} p265_tu_desc;

/* Zero-extent codes (16x16 and 32x32 TBs; ignored for smaller ones).  The parser knows the last
 * significant position of every TB (last_sig_coeff_x / y, tu.py:145-148).  Code z in the row field
 * promises that every coefficient in rows >= N >> z is zero, in the column field the same for
 * columns: 0 = nothing known, 1 = upper / left half only, 2 = first quarter only; 3 is rejected.
 * The residual kernels then skip the products of the empty rows / columns (zero contributes zero:
 * results are identical).  The choice is made per work item of the kernels (2 consecutive 32x32 TBs,
 * 4 consecutive 16x16 TBs, counted from the first descriptor of the size: the weakest promise among
 * them).  Any order is correct; fastest is a list in which the TBs of an item are spatial neighbours
 * (decoding order) and items with the same code pair follow each other (p265_b200/picture.py:
 * size_kind_order does both).  Dense arena (p265_residual_batch,
 * p265_residual_batch_dev with P265_RES_ZERO_EXTENTS): the codes are the caller's promise; a TB that
 * breaks it gets a wrong residual, nothing else is affected.  Packed stream
 * (p265_residual_batch_packed[_dev]): the codes are ignored -- the device derives them from the
 * record's significance bitmap while it expands the stream.                                     */
#define P265_TU_ZR_SHIFT 11
#define P265_TU_ZC_SHIFT 13
#define P265_TU_LEVELS_MASK 0x07ffu

#define P265_TU_DST 1u    /* trType 1: 4x4 luma of an intra CU (transform.py:97)     */
#define P265_TU_SKIP 2u   /* transform_skip_flag (tu.py:142-143)                     */
#define P265_TU_BYPASS 4u /* cu_transquant_bypass_flag (scaling.py:20-21 raises)     */
#define P265_TU_INTRA 8u  /* selects matrixId (scaling.py:33-42)                     */
#define P265_TU_PRESCALED 16u /* arena already holds d[] (pu.scaled_samples): skip 8.6.3;
                                 only with scaling_factor == NULL                      */
#define P265_TU_LEVELS8 32u   /* packed coefficient stream only: this TB's levels are int8 */

/* Packed coefficient stream -- the compact host -> device transport of TransCoeffLevel
 * (p265_residual_batch_packed).  The parser's natural output is sparse: it knows every
 * significant position when it stores a level (tu.py:331).  One record per coded TB, at byte
 * offset 4 * coeff_off of the stream:
 *     significance bitmap   N*N bits, bit y*N + x of the TB in byte (y*N + x) >> 3, bit
 *                           (y*N + x) & 7 (least significant bit first)
 *     levels                the non-zero TransCoeffLevel values in raster order of their
 *                           bitmap bits: int8 each when the descriptor has P265_TU_LEVELS8
 *                           (every |level| of the TB <= 127), little-endian int16 otherwise
 *     padding               to the next multiple of 4 bytes
 * The descriptor's `rsvd` field holds the number of levels, so a record's extent follows from
 * its descriptor alone (validation needs no pass over the stream; the device never reads more
 * than `rsvd` levels of a TB).
 * A 4K 10-bit picture of the benchmark mix is 25.1 MB as a dense int16 arena and 3.7 MB as a
 * stream.  The device expands it (unpack_kernel) into its own dense arena in descriptor
 * order and runs the same residual kernels on it.                                          */

/* Planes of a batch of 4:2:0 pictures inside ONE buffer: picture p, component c
 * starts at element p * pic_stride + plane_off[c]; rows are stride_{y,c} elements
 * apart.  Element = int16 for residual planes, uint8 (both bit depths <= 8) or
 * uint16 for sample planes.                                                        */
typedef struct p265_pic_geom {
    int32_t width, height; /* luma samples; chroma planes are width/2 x height/2    */
    int32_t n_pics;
    int32_t bit_depth_y, bit_depth_c;
    int32_t stride_y, stride_c;
    int32_t rsvd;
    int64_t plane_off[3];
    int64_t pic_stride;
} p265_pic_geom;

/* SAO parameters of one CTB: the fields sao.Sao.parse() fills (sao.py:43-77), with
 * SaoOffsetVal[1..4] already derived (7.4.9.3.2).                                  */
typedef struct p265_sao_ctb {
    uint8_t type[3];          /* 0 off, 1 band, 2 edge                              */
    uint8_t band_pos[3];
    uint8_t eo_class[3];
    int8_t offset_val[3][4];
    uint8_t pad;
    uint16_t avail;           /* bit (dy+1)*3+(dx+1): neighbour CTB usable (8.7.3)   */
} p265_sao_ctb;

/* Deblocking edge map (8.7.2; the reference only parses the control flags, pps.py:121-131,
 * slice.py:170-179): one uint16 per 8x8 luma block, raster order, (height/8) x (width/8)
 * per picture.  Bs values are final (filterEdgeFlag, slice_deblocking_filter_disabled_flag
 * and the picture / slice / tile boundary rules are already folded in by the host).    */
typedef uint16_t p265_dbk_blk;
#define P265_DBK_BS_V0 0       /* bits 0-1: Bs of the vertical edge x = 8*bx, rows 8*by .. +3   */
#define P265_DBK_BS_V1 2       /* bits 2-3: same edge, rows 8*by+4 .. +7                        */
#define P265_DBK_BS_H0 4       /* bits 4-5: Bs of the horizontal edge y = 8*by, cols 8*bx .. +3 */
#define P265_DBK_BS_H1 6       /* bits 6-7: same edge, cols 8*bx+4 .. +7                        */
#define P265_DBK_QP_SHIFT 8    /* bits 8-14: QpY of the CU (cu.py:566), 7-bit two's complement  */
#define P265_DBK_NO_FILTER 0x8000u /* pcm + pcm_loop_filter_disabled / cu_transquant_bypass     */

/* Per-CTB deblocking parameters: the offsets of the slice the CTB belongs to
 * (slice_beta_offset_div2, slice_tc_offset_div2) and the PPS chroma QP offsets (cQpPicOffset). */
typedef struct p265_dbk_ctb {
    int8_t beta_offset_div2, tc_offset_div2, cb_qp_offset, cr_qp_offset;
} p265_dbk_ctb;

#define P265_SF_BYTES 4064 /* ScalingFactor table: [sizeId][matrixId][y][x] uint8    */

#define P265_RES_ZERO_FILL 1 /* clear the planes first (TBs do not cover them)       */
#define P265_RES_SF_REPLICATED 2 /* the 16x16 / 32x32 matrices of scaling_factor are the
                                    7.4.5 up-sampling of an 8x8 list (+ DC at [0][0]), as
                                    every conformant stream has them: enables the fast
                                    per-column factor path.  Unset: any table works.   */
#define P265_RES_ZERO_EXTENTS 8 /* device entry point only (the host entry point finds out by itself, the
                                   packed entry points always use the codes their unpack pass derives): the
                                   descriptors of the 16x16 / 32x32 TBs carry zero-extent codes in rsvd.
                                   Unset: the codes are ignored (full passes for every TB)                */
#define P265_RES_DENSE_ARENA 4 /* device entry point only (the host entry point finds out by
                                  itself): inside the 8x8 bin and inside the 4x4 bin the
                                  coefficients of descriptor i directly follow those of
                                  descriptor i-1 (coeff_off grows by 4 / by 1), as the packer
                                  emits them: tile copies start without their descriptors */

/* ---- context ------------------------------------------------------------------ */
int p265_abi_version(void);
const char *p265_last_error(void);
int p265_device_count(void);
/* stream == NULL: the context creates its own non-blocking stream                  */
int p265_ctx_create(int device, void *stream, p265_ctx **out);
int p265_ctx_destroy(p265_ctx *ctx);
int p265_sync(p265_ctx *ctx);
/* enable != 0: the host-buffer batch entry points (p265_residual_batch[_packed],
 * p265_sao_batch, p265_loop_filter_batch, p265_reconstruct_batch, p265_deblock_batch) return as
 * soon as their copies and kernels are queued on the context's stream; results are valid
 * after p265_sync().  Contract of an asynchronous context:
 *   - every host buffer passed in (inputs AND outputs) stays alive and unmodified until the
 *     next p265_sync() on this context; page-locked memory is needed for real overlap
 *     (pageable memory makes the copies synchronous, not wrong);
 *   - the context owns ONE stream and ONE set of grow-only device scratch buffers which the
 *     entry points share (the descriptor / parameter slot, the plane slots): calls queued on
 *     one context are serialised by that stream, which is what keeps the sharing correct --
 *     a call never starts before the previous call's kernels and copies have finished.  A
 *     scratch buffer that has to grow while work is queued is released only after the stream
 *     has drained (the call that grows it synchronises first);
 *   - p265_dequant_batch, p265_ref_literal_batch and p265_idct_1d (parity helpers) are always
 *     synchronous;
 *   - overlap between pictures comes from SEVERAL contexts (one per in-flight picture): the
 *     H2D copy of one picture then overlaps the kernels and the D2H copy of another.        */
int p265_ctx_set_async(p265_ctx *ctx, int enable);
/* Timeline of the host entry points (p265_residual_batch_packed, p265_sao_batch): with tracing on, every
 * call leaves four marks on the context's stream (0 call start, 1 inputs copied, 2 kernels done, 3 outputs
 * back).  p265_trace_read synchronises the context, writes up to max_marks triples (kind: 1 residual,
 * 2 SAO; phase; milliseconds since the process enabled tracing on this device) into out[3 * max_marks],
 * forgets them and returns how many it wrote.  The reference has no tracing of this path (its log.py
 * writes syntax-element logs); this is what tools/e2e_trace.py draws the copy-engine occupancy from.  */
int p265_ctx_set_trace(p265_ctx *ctx, int enable);
int p265_trace_read(p265_ctx *ctx, double *out, int max_marks);
int p265_sm_count(p265_ctx *ctx);
/* kernels launched through this context so far (bench.py reports it as gpu_launches) */
uint64_t p265_launch_count(p265_ctx *ctx);

/* ---- residual: dequantisation + inverse transform (scaling.py + transform.py) --- */
/* tus sorted by log2n descending; bin_counts = number of 32x32,16x16,8x8,4x4 TBs.  */
int p265_residual_batch(p265_ctx *ctx, const p265_tu_desc *tus, const int32_t bin_counts[4],
                        const int16_t *coeffs, size_t n_coeffs,
                        const uint8_t *scaling_factor /* P265_SF_BYTES or NULL = flat 16 */,
                        const p265_pic_geom *geom, int16_t *residual /* host out */,
                        int flags);
/* same result from the packed coefficient stream (see P265_TU_LEVELS8 above): tus[i].coeff_off
 * is the record's offset in units of 4 bytes.  What a parser that emits (position, level)
 * pairs hands over; 4-7x fewer host -> device bytes than the dense arena.                  */
int p265_residual_batch_packed(p265_ctx *ctx, const p265_tu_desc *tus, const int32_t bin_counts[4],
                               const uint8_t *stream, size_t stream_bytes,
                               const uint8_t *scaling_factor, const p265_pic_geom *geom,
                               int16_t *residual /* host out */, int flags);
/* device-resident variant: d_arena (>= sum of N*N int16 over all TBs, + 32 bytes) and d_tus_out
 * (n descriptors) receive the expanded arena and the descriptors that index it             */
int p265_residual_batch_packed_dev(p265_ctx *ctx, const p265_tu_desc *d_tus,
                                   const int32_t bin_counts[4], const uint8_t *d_stream,
                                   const uint8_t *d_scaling_factor, const p265_pic_geom *geom,
                                   int16_t *d_arena, p265_tu_desc *d_tus_out, int16_t *d_residual,
                                   int flags);
/* same, every pointer already resident in device memory; asynchronous on the stream */
int p265_residual_batch_dev(p265_ctx *ctx, const p265_tu_desc *d_tus,
                            const int32_t bin_counts[4], const int16_t *d_coeffs,
                            const uint8_t *d_scaling_factor, const p265_pic_geom *geom,
                            int16_t *d_residual, int flags);
/* scaling.inverse_scaling only: d[] in arena layout (pu.scaled_samples content)    */
int p265_dequant_batch(p265_ctx *ctx, const p265_tu_desc *tus, int32_t n_tus,
                       const int16_t *coeffs, size_t n_coeffs, const uint8_t *scaling_factor,
                       int bit_depth_y, int bit_depth_c, int16_t *scaled /* host out */);
/* transform.py:89-109 exactly as written (parity-test-only): in d[]
 * arena ([y][x] per TB), out r[] int32 arena, [x][y] per TB like the reference     */
int p265_ref_literal_batch(p265_ctx *ctx, const p265_tu_desc *tus, int32_t n_tus,
                           const int16_t *scaled, size_t n_coeffs, int32_t *out);

/* transform.inverse_transform_1d(x, log2size, tr_type) (transform.py:74-87) on one
 * vector: mode 0 = the standard's orientation y[i] = sum_j M[j][i] x[j] (8.6.4.2),
 * mode 1 = as written in the reference, y[i] = sum_j M[i][j*32/N] x[j].  Host in/out. */
int p265_idct_1d(p265_ctx *ctx, const int32_t *x, int log2size, int tr_type, int mode,
                 int32_t *y);

/* ---- SAO (8.7.3; parameters from sao.py) ---------------------------------------- */
/* params: n_pics * ctbs_h * ctbs_w records, raster order per picture.
 * no_filter: NULL or n_pics * ceil(h/8) * ceil(w/8) bytes, non-zero = luma 8x8 block
 * (and its 4x4 chroma blocks) keeps its samples (pcm + pcm_loop_filter_disabled,
 * cu_transquant_bypass).                                                           */
/* out == rec is allowed for the host entry point (the device works out of place): with
 * page-locked memory only the CTBs that SAO modifies are then written back.  out != rec: the
 * plane rows of `out` are written, row padding and inter-plane gaps are left untouched.      */
int p265_sao_batch(p265_ctx *ctx, const void *rec /* host in */, void *out /* host out */,
                   const p265_pic_geom *geom, int ctb_log2, const p265_sao_ctb *params,
                   const uint8_t *no_filter);
int p265_sao_batch_dev(p265_ctx *ctx, const void *d_rec, void *d_out,
                       const p265_pic_geom *geom, int ctb_log2, const p265_sao_ctb *d_params,
                       const uint8_t *d_no_filter);

/* ---- reconstruction (reconstruction.py:4-27) ------------- */
/* rec = Clip1(pred + residual) over whole planes: pred / rec are sample planes (uint8 or
 * uint16 like the SAO planes), residual the int16 planes p265_residual_batch produced;
 * all three share `geom`.  rec may alias pred.                                        */
int p265_reconstruct_batch(p265_ctx *ctx, const void *pred /* host in */,
                           const int16_t *residual /* host in */, void *rec /* host out */,
                           const p265_pic_geom *geom);
int p265_reconstruct_batch_dev(p265_ctx *ctx, const void *d_pred, const int16_t *d_residual,
                               void *d_rec, const p265_pic_geom *geom);

/* ---- deblocking (8.7.2) ----------------------------------- */
/* In place on reconstructed sample planes (same element type rule as SAO).  blk: n_pics *
 * (height/8) * (width/8) edge-map entries; ctb: n_pics * ctbs_h * ctbs_w records.
 * width and height must be multiples of 8.                                              */
int p265_deblock_batch(p265_ctx *ctx, void *planes /* host in/out */, const p265_pic_geom *geom,
                       int ctb_log2, const p265_dbk_blk *blk, const p265_dbk_ctb *ctb);
int p265_deblock_batch_dev(p265_ctx *ctx, void *d_planes, const p265_pic_geom *geom, int ctb_log2,
                           const p265_dbk_blk *d_blk, const p265_dbk_ctb *d_ctb);

/* ---- loop filters in one call (8.7.2 then 8.7.3) --------------------------------- */
/* Deblocking in place, then SAO, on reconstructed planes: ONE host -> device copy of the planes
 * and ONE copy back instead of the two round trips of p265_deblock_batch + p265_sao_batch.
 * The reference parses the controls (pps.py:121-131, slice.py:170-179) and has neither filter.
 * blk / dbk_ctb == NULL skips deblocking (slice_deblocking_filter_disabled_flag everywhere),
 * sao == NULL skips SAO.  `planes` is read and overwritten with the final samples.           */
int p265_loop_filter_batch(p265_ctx *ctx, void *planes /* host in/out */, const p265_pic_geom *geom,
                           int ctb_log2, const p265_dbk_blk *blk, const p265_dbk_ctb *dbk_ctb,
                           const p265_sao_ctb *sao, const uint8_t *no_filter);

/* ---- measurement helpers ------------------------------------------------------- */
/* Register-resident integer-pipe microbenchmark; kind: 0 IMAD, 1 IADD3, 2 IMAD+IADD3
 * interleaved, 3 DP2A, 4 SHF, 5 DP2A+IADD3.  Returns lane-ops per second.           */
int p265_int_peak(p265_ctx *ctx, int kind, double *ops_per_s, double *ms);
/* Plain page-locked host <-> device copies on two streams at once (what the host entry points are made
 * of): copies of `bytes`, `reps` per direction, both directions concurrently, cycling through `n_buffers`
 * distinct host buffers per direction (1 = the same buffer over and over, which stays in the host's
 * last-level cache and is what copy benchmarks usually report; a decoder moves pictures between many
 * buffers).  Returns the two rates in bytes per second -- the ceiling bench.py compares its end-to-end
 * number with (e2e.pcie_frac).  One direction alone: pass NULL for the other rate.                    */
int p265_pcie_probe(p265_ctx *ctx, size_t bytes, int n_buffers, int reps, double *h2d_bytes_per_s,
                    double *d2h_bytes_per_s);

#ifdef __cplusplus
}
#endif
#endif /* P265_B200_H */
