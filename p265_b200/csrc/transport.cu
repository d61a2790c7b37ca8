// Host <-> device transport helpers of the host-buffer entry points (api.cu):
//   * sao_writeback_kernel: after SAO on a picture whose host buffer is filtered in place, only
//     the CTB components with sao type != 0 differ from what the host already holds (8.7.3 leaves
//     the others untouched).  The kernel stores exactly those rectangles straight into the
//     caller's page-locked buffer over PCIe (device-visible host address), instead of a full-plane
//     D2H copy: a third of the CTBs of the benchmark's config 4 never cross the bus again.
//   * run_pcie_probe: plain page-locked copies in both directions at once -- the ceiling bench.py
//     compares the end-to-end number with.
#include <cuda_runtime.h>
#include <string.h>

#include <vector>

#include "internal.h"

namespace p265 {

struct WbArgs {
    const unsigned char *src;  // device planes (SAO output)
    unsigned char *dst;        // device-visible address of the host planes
    const p265_sao_ctb *params;
    int64_t plane_off_b[3];    // bytes
    int64_t pic_stride_b;
    int32_t row_b[3], stride_b[3], rows[3];  // plane row bytes / pitch / height per component
    int32_t ctb_row_b[3], ctb_log2_rows[3];  // CTB width in bytes and log2 of its height, per component
    int32_t ctbs_w, ctbs;
};

// One thread = one CHUNK-byte piece of one sample row; a warp covers 32 consecutive pieces of the row
// (512 contiguous bytes when the CTBs underneath are all filtered), blockIdx.y = row over the three
// planes, blockIdx.z = picture.  CHUNK divides every CTB width and row length, so a piece never
// straddles a CTB or the end of a row.
template <int CHUNK>
__global__ void __launch_bounds__(128) sao_writeback_kernel(const __grid_constant__ WbArgs a) {
    int y = blockIdx.y, c = 0;
    if (y >= a.rows[0]) { y -= a.rows[0]; c = 1; }
    if (c == 1 && y >= a.rows[1]) { y -= a.rows[1]; c = 2; }
    const int xb = (blockIdx.x * blockDim.x + threadIdx.x) * CHUNK;
    if (xb >= a.row_b[c]) return;
    const int pic = blockIdx.z;
    const p265_sao_ctb *q = a.params + (size_t)pic * a.ctbs + (size_t)(y >> a.ctb_log2_rows[c]) * a.ctbs_w + xb / a.ctb_row_b[c];
    if (q->type[c] == 0) return;
    const size_t off = (size_t)pic * a.pic_stride_b + a.plane_off_b[c] + (size_t)y * a.stride_b[c] + xb;
    if (CHUNK == 16) *reinterpret_cast<uint4 *>(a.dst + off) = *reinterpret_cast<const uint4 *>(a.src + off);
    else if (CHUNK == 8) *reinterpret_cast<uint2 *>(a.dst + off) = *reinterpret_cast<const uint2 *>(a.src + off);
    else *reinterpret_cast<uint32_t *>(a.dst + off) = *reinterpret_cast<const uint32_t *>(a.src + off);
}

int launch_sao_writeback(p265_ctx *ctx, const void *d_out, void *h_out, const p265_pic_geom *g, int ctb_log2,
                         const p265_sao_ctb *d_params) {
    const int eb = (g->bit_depth_y > 8 || g->bit_depth_c > 8) ? 2 : 1;
    WbArgs a;
    a.src = static_cast<const unsigned char *>(d_out);
    a.dst = static_cast<unsigned char *>(h_out);
    a.params = d_params;
    const int ctb = 1 << ctb_log2;
    for (int c = 0; c < 3; c++) {
        a.plane_off_b[c] = g->plane_off[c] * eb;
        a.row_b[c] = (c ? g->width / 2 : g->width) * eb;
        a.stride_b[c] = (c ? g->stride_c : g->stride_y) * eb;
        a.rows[c] = c ? g->height / 2 : g->height;
        a.ctb_row_b[c] = (c ? ctb / 2 : ctb) * eb;
        a.ctb_log2_rows[c] = c ? ctb_log2 - 1 : ctb_log2;
    }
    a.pic_stride_b = g->pic_stride * eb;
    a.ctbs_w = (g->width + ctb - 1) / ctb;
    a.ctbs = a.ctbs_w * ((g->height + ctb - 1) / ctb);
    int chunk = 16;  // largest power of two dividing every row length and CTB width (bytes)
    while (chunk > 4 && (a.row_b[0] % chunk || a.row_b[1] % chunk || a.ctb_row_b[1] % chunk)) chunk >>= 1;
    if (a.row_b[1] % chunk) return set_error(P265_EINVAL, "picture width %d cannot be written back in 4-byte pieces", g->width);
    const int rows = a.rows[0] + a.rows[1] + a.rows[2];
    if (rows > 65535 || g->n_pics > 65535) return set_error(P265_EINVAL, "too many rows / pictures in one SAO batch");
    const dim3 grid((a.row_b[0] / chunk + 127) / 128, rows, g->n_pics);
    if (chunk == 16) sao_writeback_kernel<16><<<grid, 128, 0, ctx->stream>>>(a);
    else if (chunk == 8) sao_writeback_kernel<8><<<grid, 128, 0, ctx->stream>>>(a);
    else sao_writeback_kernel<4><<<grid, 128, 0, ctx->stream>>>(a);
    P265_CUDA(cudaGetLastError());
    ctx->launches++;
    return P265_OK;
}

int run_pcie_probe(p265_ctx *ctx, size_t bytes, int n_buffers, int reps, double *h2d, double *d2h) {
    if (bytes < 4096 || reps < 1 || n_buffers < 1 || n_buffers > 64)
        return set_error(P265_EINVAL, "p265_pcie_probe: bytes >= 4096, 1 <= n_buffers <= 64 and reps >= 1 expected");
    // n_buffers distinct page-locked buffers per direction, used round robin: one buffer copied over and over
    // stays in the host's last-level cache (measured on the B200 boxes: 48 / 50 GB/s both ways at once), a
    // pipeline that moves pictures between many buffers does not (43 / 46 GB/s)
    std::vector<void *> h_in(n_buffers, nullptr), h_out(n_buffers, nullptr);
    void *d_in = nullptr, *d_out = nullptr;
    cudaStream_t s[2] = {nullptr, nullptr};
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    int rc = P265_OK;
    auto fail = [&](cudaError_t e, const char *what) { rc = cuda_error(e, what, __FILE__, __LINE__); };
    cudaError_t e;
    do {
        for (int b = 0; b < n_buffers && rc == P265_OK; b++) {
            if (h2d && (e = cudaHostAlloc(&h_in[b], bytes, cudaHostAllocDefault)) != cudaSuccess) { fail(e, "cudaHostAlloc"); break; }
            if (d2h && (e = cudaHostAlloc(&h_out[b], bytes, cudaHostAllocDefault)) != cudaSuccess) { fail(e, "cudaHostAlloc"); break; }
            if (h_in[b]) memset(h_in[b], 1, bytes);   // first touch on the calling thread's NUMA node
            if (h_out[b]) memset(h_out[b], 2, bytes);
        }
        if (rc) break;
        if (h2d && (e = cudaMalloc(&d_in, bytes)) != cudaSuccess) { fail(e, "cudaMalloc"); break; }
        if (d2h && (e = cudaMalloc(&d_out, bytes)) != cudaSuccess) { fail(e, "cudaMalloc"); break; }
        for (int i = 0; i < 2 && rc == P265_OK; i++)
            if ((e = cudaStreamCreateWithFlags(&s[i], cudaStreamNonBlocking)) != cudaSuccess) fail(e, "cudaStreamCreate");
        for (int i = 0; i < 4 && rc == P265_OK; i++)
            if ((e = cudaEventCreate(&ev[i])) != cudaSuccess) fail(e, "cudaEventCreate");
        if (rc) break;
        for (int pass = 0; pass < 2 && rc == P265_OK; pass++) {  // pass 0 = warm-up
            const int n = pass ? reps : (n_buffers < 2 ? 1 : 2);
            if (h2d) cudaEventRecord(ev[0], s[0]);
            if (d2h) cudaEventRecord(ev[2], s[1]);
            for (int r = 0; r < n; r++) {  // interleaved issue: both copy engines busy from the start
                if (h2d) cudaMemcpyAsync(d_in, h_in[r % n_buffers], bytes, cudaMemcpyHostToDevice, s[0]);
                if (d2h) cudaMemcpyAsync(h_out[r % n_buffers], d_out, bytes, cudaMemcpyDeviceToHost, s[1]);
            }
            if (h2d) cudaEventRecord(ev[1], s[0]);
            if (d2h) cudaEventRecord(ev[3], s[1]);
            if ((e = cudaStreamSynchronize(s[0])) != cudaSuccess) { fail(e, "cudaStreamSynchronize"); break; }
            if ((e = cudaStreamSynchronize(s[1])) != cudaSuccess) { fail(e, "cudaStreamSynchronize"); break; }
        }
        if (rc) break;
        float ms = 0;
        if (h2d) { cudaEventElapsedTime(&ms, ev[0], ev[1]); *h2d = (double)bytes * reps / (ms * 1e-3); }
        if (d2h) { cudaEventElapsedTime(&ms, ev[2], ev[3]); *d2h = (double)bytes * reps / (ms * 1e-3); }
    } while (0);
    for (int i = 0; i < 4; i++) if (ev[i]) cudaEventDestroy(ev[i]);
    for (int i = 0; i < 2; i++) if (s[i]) cudaStreamDestroy(s[i]);
    if (d_in) cudaFree(d_in);
    if (d_out) cudaFree(d_out);
    for (void *p : h_in) if (p) cudaFreeHost(p);
    for (void *p : h_out) if (p) cudaFreeHost(p);
    return rc;
}

}  // namespace p265
