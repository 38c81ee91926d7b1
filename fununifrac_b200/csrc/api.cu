// api.cu — the host-buffer entry point: the compute section of the reference driver
// (src/compute_all_pairwise_fununifrac.py:64-73: push every profile up, then pairwise_computation)
// behind one C call with HOST pointers.
//
// Pipeline: X is copied to the device in row blocks on a copy stream (double-buffered staging),
// each block is pushed up on the compute stream as soon as it has landed, then the all-pairs kernel
// runs over the complete node-major P, followed by the float64 refinement pass over the few pairs the
// fp32 operands cannot certify (refine.cu).  The matrix is always produced in device memory; when D_h
// is pinned memory the tile epilogue ALSO writes it straight into D_h over PCIe (coalesced 1 KB rows),
// so the device-to-host transfer overlaps the kernel; otherwise D is copied back at the end.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

using namespace fuf;

static bool is_pinned_host(const void *p)
{
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost;
}

extern "C" int fuf_all_pairs_host(fuf_ctx *ctx, const void *X_h, int x_dtype, int64_t N, int64_t ldx, int mode,
                                  double *D_h, int64_t ldd, int flags)
{
    FUF_REQUIRE(ctx && X_h && D_h, "fuf_all_pairs_host: null argument");
    FUF_REQUIRE(mode == FUF_MODE_L1 || mode == FUF_MODE_L2, "fuf_all_pairs_host: bad mode %d", mode);
    FUF_REQUIRE(x_dtype == FUF_F32 || x_dtype == FUF_F64, "fuf_all_pairs_host: bad x_dtype %d", x_dtype);
    FUF_REQUIRE(N >= 0 && ldx >= ctx->n && ldd >= N, "fuf_all_pairs_host: bad shape N=%lld ldx=%lld ldd=%lld",
                (long long)N, (long long)ldx, (long long)ldd);
    const bool validate = (flags & FUF_AP_VALIDATE) != 0, do_refine = !(flags & FUF_AP_NO_REFINE) && !validate;
    FUF_REQUIRE(!validate || x_dtype == FUF_F64, "fuf_all_pairs_host: validation mode needs float64 profiles");
    if (N == 0) return FUF_OK;
    FUF_CUDA(cudaSetDevice(ctx->device));
    const int32_t n = ctx->n;
    const size_t esz = x_dtype == FUF_F32 ? 4 : 8;
    const size_t psz = validate ? 8 : 4;
    const int64_t ldp = fuf_padded_samples(N);

    int rc;
    if ((rc = ctx->ws_p.reserve((size_t)n * ldp * psz))) return rc;
    double *P64T = nullptr, *center = nullptr, *stat = nullptr;
    if (!validate) {
        // centre [n] | error statistic [ldp] | P64T [n][ldp] (only when refining)
        const size_t head = ((size_t)n + (size_t)ldp) * sizeof(double);
        if ((rc = ctx->ws_p64.reserve(head + (do_refine ? (size_t)n * ldp * sizeof(double) : 0)))) return rc;
        center = (double *)ctx->ws_p64.ptr;
        stat = center + n;
        if (do_refine) P64T = stat + ldp;
    }
    // row blocks of ~64 MB, multiples of 128 samples
    int64_t rb = (int64_t)(((size_t)64 << 20) / ((size_t)n * esz));
    rb = std::max<int64_t>(128, rb / 128 * 128);
    rb = std::min<int64_t>(rb, ldp);
    const size_t blk_bytes = (size_t)rb * n * esz;
    if ((rc = ctx->ws_x.reserve(2 * blk_bytes))) return rc;
    const bool d_pinned = is_pinned_host(D_h) && !getenv("FUF_NO_ZEROCOPY");
    if ((rc = ctx->ws_d.reserve((size_t)N * N * sizeof(double)))) return rc;
    double *D_dev = (double *)ctx->ws_d.ptr, *D_zc = nullptr;
    if (d_pinned) FUF_CUDA(cudaHostGetDevicePointer((void **)&D_zc, D_h, 0));
    cudaStream_t sc = ctx->s_copy, sk = ctx->s_comp;
    cudaEvent_t e_start = ctx->ev[0], e_push = ctx->ev[1], e_end = ctx->ev[2];
    cudaEvent_t e_copied[2] = {ctx->ev[3], ctx->ev[4]}, e_free[2] = {ctx->ev[5], ctx->ev[6]};

    FUF_CUDA(cudaEventRecord(e_start, sk));
    FUF_CUDA(cudaStreamWaitEvent(sc, e_start, 0));
    int64_t done = 0;
    for (int b = 0; done < N; b++) {
        const int slot = b & 1;
        const int64_t rows = std::min<int64_t>(rb, N - done);
        char *xbuf = (char *)ctx->ws_x.ptr + (size_t)slot * blk_bytes;
        if (b >= 2) FUF_CUDA(cudaStreamWaitEvent(sc, e_free[slot], 0));
        FUF_CUDA(cudaMemcpy2DAsync(xbuf, (size_t)n * esz, (const char *)X_h + (size_t)done * ldx * esz,
                                   (size_t)ldx * esz, (size_t)n * esz, (size_t)rows, cudaMemcpyHostToDevice, sc));
        FUF_CUDA(cudaEventRecord(e_copied[slot], sc));
        FUF_CUDA(cudaStreamWaitEvent(sk, e_copied[slot], 0));
        if (validate) {
            rc = push_up_f64_exact(ctx, (const double *)xbuf, rows, n, mode, (double *)ctx->ws_p.ptr, ldp, done, sk);
        } else {
            if (b == 0) {  // the centre of the whole job: pushed mean of the first few samples
                rc = push_up_center(ctx, xbuf, x_dtype, std::min<int64_t>(rows, FUF_CENTER_SAMPLES), n, mode, center, sk);
                if (rc) return rc;
            }
            rc = push_up_f32(ctx, xbuf, x_dtype, rows, n, mode, center, (float *)ctx->ws_p.ptr, ldp, P64T, ldp, stat,
                             done, sk);
        }
        if (rc) return rc;
        FUF_CUDA(cudaEventRecord(e_free[slot], sk));
        done += rows;
    }
    FUF_CUDA(cudaEventRecord(e_push, sk));
    if (validate) {
        rc = pairwise_f64(ctx, (const double *)ctx->ws_p.ptr, N, ldp, mode, D_dev, N, sk);
    } else {
        rc = pairwise_f32(ctx, (const float *)ctx->ws_p.ptr, N, ldp, 0, 0, mode, nullptr, 0, D_dev, N, 0, sk, D_zc, ldd);
        if (!rc && do_refine)
            rc = refine(ctx, P64T, N, ldp, 0, 0, mode, stat, 0.0, nullptr, 0, D_dev, N, sk, D_zc, ldd);
    }
    if (rc) return rc;
    if (!d_pinned || validate)
        FUF_CUDA(cudaMemcpy2DAsync(D_h, (size_t)ldd * sizeof(double), D_dev, (size_t)N * sizeof(double),
                                   (size_t)N * sizeof(double), (size_t)N, cudaMemcpyDeviceToHost, sk));
    FUF_CUDA(cudaEventRecord(e_end, sk));
    FUF_CUDA(cudaStreamSynchronize(sk));
    FUF_CUDA(cudaStreamSynchronize(sc));
    cudaEventElapsedTime(&ctx->t_h2d_push, e_start, e_push);
    cudaEventElapsedTime(&ctx->t_pair, e_push, e_end);
    cudaEventElapsedTime(&ctx->t_total, e_start, e_end);
    ctx->last_refined = 0;
    if (do_refine && ctx->d_refined) {
        unsigned long long v = 0;
        FUF_CUDA(cudaMemcpy(&v, ctx->d_refined, sizeof(v), cudaMemcpyDeviceToHost));
        ctx->last_refined = (int64_t)v;
    }
    return FUF_OK;
}

extern "C" int fuf_last_timing(const fuf_ctx *ctx, float *h2d_pushup_ms, float *pairwise_ms, float *total_ms)
{
    FUF_REQUIRE(ctx != nullptr, "fuf_last_timing: null context");
    if (h2d_pushup_ms) *h2d_pushup_ms = ctx->t_h2d_push;
    if (pairwise_ms) *pairwise_ms = ctx->t_pair;
    if (total_ms) *total_ms = ctx->t_total;
    return FUF_OK;
}
