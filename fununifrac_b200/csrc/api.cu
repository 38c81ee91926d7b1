// api.cu — the host-buffer entry point: the compute section of the reference driver
// (src/compute_all_pairwise_fununifrac.py:64-73: push every profile up, then pairwise_computation)
// behind one C call with HOST pointers.
//
// Pipeline: X is copied to the device in sample bands on a copy stream (all copies queued up front into a
// full-size device copy), each band is pushed up on the compute stream as soon as it has landed, then the all-pairs
// kernel runs band by band (see below), each band followed by the float64 refinement pass over the few pairs the
// fp32 operands cannot certify (refine.cu) and by the device-to-host copy of the finished part of D on a
// third stream.  (Letting the tile epilogues write D straight into pinned host memory was measured: the
// kernel then stalls for exactly the PCIe time of the matrix, so the DMA engines do it instead.)  --L2 from 512 samples
// on is ONE launch of the tensor-core Gram kernel; it counts its finished tile rows in completion segments, and the
// third stream, waiting on those counters, copies the finished rows and mirrored columns while the launch goes on.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

using namespace fuf;

// Every early return (an error in the middle of the pipeline) must leave no copy in flight that still reads the
// caller's X_h or writes its D_h: drain the three streams unless the success path has already done so.
namespace {
struct StreamDrain {
    fuf_ctx *ctx;
    bool armed = true;
    explicit StreamDrain(fuf_ctx *c) : ctx(c) {}
    ~StreamDrain()
    {
        if (!armed) return;
        cudaStreamSynchronize(ctx->s_copy);
        cudaStreamSynchronize(ctx->s_comp);
        cudaStreamSynchronize(ctx->s_d2h);
        cudaGetLastError();
    }
};
}  // namespace

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
    double *center = nullptr, *stat = nullptr;
    if (!validate) {
        // centre [n] | error statistic [ldp] | Gram norms [ldp] | Gram error statistic [ldp]
        if ((rc = ctx->ws_p64.reserve(((size_t)n + 3 * (size_t)ldp) * sizeof(double)))) return rc;
        center = (double *)ctx->ws_p64.ptr;
        stat = center + n;
    }
    // The whole X gets a device copy (no staging ring: the copy stream never waits for compute).  Samples are
    // cut into up to 16 bands of whole 128-sample tiles; band b is copied on the copy stream, and as soon as it has
    // landed the compute stream pushes it up and computes the column band of D that involves it (tiles
    // (ti, tj) with tj in the band, ti <= tj) while the later bands are still crossing PCIe.
    const int64_t tiles_per_dim = ldp / 128;
    const bool use_gram = !validate && mode == FUF_MODE_L2 && N >= FUF_GRAM_MIN_SAMPLES && !(flags & FUF_AP_NO_GRAM);
    int nb = (validate || use_gram) ? 1 : (int)std::min<int64_t>(16, std::max<int64_t>(1, tiles_per_dim / 4));
    if (getenv("FUF_HOST_BANDS")) nb = std::max(1, std::min(64, atoi(getenv("FUF_HOST_BANDS"))));
    nb = (int)std::min<int64_t>(nb, tiles_per_dim);
    // device copy of X with a 128-byte aligned row pitch: the TMA push-up kernel's tensor map needs 16 bytes, and with
    // 128 every 128-byte box row of a sample is exactly four 32-byte sectors (an odd pitch costs ~10 % more DRAM reads)
    const int64_t ldx_dev = ((int64_t)n * (int64_t)esz + 127) / 128 * 128 / (int64_t)esz;
    if ((rc = ctx->ws_x.reserve((size_t)N * ldx_dev * esz))) return rc;
    const bool d_pinned = is_pinned_host(D_h);
    if ((rc = ctx->ws_d.reserve((size_t)N * N * sizeof(double)))) return rc;
    double *D_dev = (double *)ctx->ws_d.ptr;
    // pinned D_h: every finished band of D is copied while the next one computes; pageable D_h: one copy at the end
    const bool band_d2h = !validate && d_pinned;
    cudaStream_t sc = ctx->s_copy, sk = ctx->s_comp, sd = ctx->s_d2h;
    // per-pair refinement budget: 0.1 % of the pairs (beyond it the float64 tile kernel redoes the whole matrix)
    const int64_t refine_budget = std::max<int64_t>(1024, N * (N - 1) / 2 / 1000);
    cudaEvent_t e_start = ctx->ev[0], e_push = ctx->ev[1], e_end = ctx->ev[2];
    std::vector<cudaEvent_t> &e_band = ctx->ev_band;
    while ((int)e_band.size() < 2 * nb) {
        cudaEvent_t e;
        FUF_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        e_band.push_back(e);
    }
    std::vector<int64_t> band_tile(nb + 1);
    for (int b = 0; b <= nb; b++) band_tile[b] = tiles_per_dim * b / nb;

    StreamDrain drain(ctx);
    FUF_CUDA(cudaEventRecord(e_start, sk));
    FUF_CUDA(cudaStreamWaitEvent(sc, e_start, 0));
    for (int b = 0; b < nb; b++) {  // all copies are queued up front
        const int64_t s0 = band_tile[b] * 128, s1 = std::min<int64_t>(N, band_tile[b + 1] * 128);
        if (s1 > s0)
            FUF_RC(copy2d_async((char *)ctx->ws_x.ptr + (size_t)s0 * ldx_dev * esz, (size_t)ldx_dev * esz,
                                       (const char *)X_h + (size_t)s0 * ldx * esz, (size_t)ldx * esz, (size_t)n * esz,
                                       (size_t)(s1 - s0), sc));
        FUF_CUDA(cudaEventRecord(e_band[b], sc));
    }
    for (int b = 0; b < nb; b++) {
        const int64_t s0 = band_tile[b] * 128, s1 = std::min<int64_t>(N, band_tile[b + 1] * 128);
        FUF_CUDA(cudaStreamWaitEvent(sk, e_band[b], 0));
        if (s1 <= s0) continue;
        const char *xb = (const char *)ctx->ws_x.ptr + (size_t)s0 * ldx_dev * esz;
        if (validate) {
            rc = push_up_f64_exact(ctx, (const double *)xb, s1 - s0, ldx_dev, mode, (double *)ctx->ws_p.ptr, ldp, s0, sk);
        } else {
            if (b == 0) {  // the centre of the whole job: pushed mean of the first few samples
                rc = push_up_center(ctx, xb, x_dtype, std::min<int64_t>(s1 - s0, FUF_CENTER_SAMPLES), ldx_dev, mode, center, sk);
                if (rc) return rc;
            }
            // fp32 operands + statistics only: the float64 profiles are produced further down, and only if some pair
            // turns out to need them
            rc = push_up_f32(ctx, xb, x_dtype, s1 - s0, ldx_dev, mode, center, (float *)ctx->ws_p.ptr, ldp, nullptr, 0,
                             stat, s0, sk);
        }
        if (rc) return rc;
        if (b == 0) FUF_CUDA(cudaEventRecord(e_push, sk));
        if (use_gram) {
            // --L2 on the tensor cores: the Gram kernel needs every sample (node means), so no banding here
            double *qg = stat + ldp, *eg = qg + ldp;
            // With a page-locked D_h the matrix leaves while it is being computed: the launch is cut into completion
            // segments of whole tile rows, the kernel counts the finished items of every segment, and the device-to-host
            // stream waits on each counter in turn (cuStreamWaitValue32) before it copies the segment's rows — final in
            // every column by then: right of the diagonal from the segment itself, left of it mirrored by the earlier
            // ones.  20 GB at N = 50,000 take 0.37-0.40 s over PCIe against 0.26 s of tensor-core work: the copies start
            // ~10 ms into the launch instead of after it (0.88 s -> 0.67 s for the whole call).
            // (Storing the tiles into the mapped host matrix from the kernel was measured first: the epilogue warps then
            // run at the link's pace — or far below it with narrow stores — and hold the tensor pipe back.)
            GramSegments segs;
            if (band_d2h && !getenv("FUF_GRAM_NO_SEGMENTS")) {
                constexpr size_t max_seg = 4096;
                if ((rc = ctx->ws_seg.reserve(max_seg * sizeof(uint32_t)))) return rc;
                segs.seg_done = (uint32_t *)ctx->ws_seg.ptr;
                segs.n_wanted = (int32_t)std::max<int64_t>(1, std::min<int64_t>(32, tiles_per_dim / 2));
                if (getenv("FUF_GRAM_SEGMENTS")) segs.n_wanted = std::max(1, std::min(4000, atoi(getenv("FUF_GRAM_SEGMENTS"))));
                // counters zeroed now; the copy stream may evaluate its waits from here on
                FUF_CUDA(cudaMemsetAsync(segs.seg_done, 0, max_seg * sizeof(uint32_t), sk));
                FUF_CUDA(cudaEventRecord(e_band[nb], sk));
                FUF_CUDA(cudaStreamWaitEvent(sd, e_band[nb], 0));
            }
            rc = pairwise_gram(ctx, (const float *)ctx->ws_p.ptr, N, ldp, 0, 0, stat, qg, eg, nullptr, 0, D_dev, N, sk,
                               segs.seg_done ? &segs : nullptr);
            if (rc) return rc;
            int64_t R0 = 0;
            for (size_t sg = 0; sg < segs.row_end.size(); sg++) {
                const int64_t R1 = std::min<int64_t>(N, (int64_t)segs.row_end[sg] * 128);
                if ((rc = stream_wait_u32(segs.seg_done + sg, segs.expected[sg], sd))) return rc;
                FUF_RC(copy2d_async(D_h + (size_t)R0 * ldd, (size_t)ldd * sizeof(double), D_dev + (size_t)R0 * N,
                                           (size_t)N * sizeof(double), (size_t)N * sizeof(double), (size_t)(R1 - R0), sd));
                R0 = R1;
            }
            const bool gram_segmented = !segs.row_end.empty();
            if (do_refine) {
                // one counting scan, both criteria: rounding statistic of the operands, and D small against the Gram norms
                rc = refine(ctx, nullptr, N, ldp, 0, 0, mode, eg, 0.0, nullptr, 0, D_dev, N, sk, 0, N, true,
                            FUF_CRIT_ROUNDING, qg, 2.9e-3, 0);
                if (rc) return rc;
            }
            if (band_d2h && !gram_segmented) {
                FUF_RC(copy2d_async(D_h, (size_t)ldd * sizeof(double), D_dev, (size_t)N * sizeof(double),
                                           (size_t)N * sizeof(double), (size_t)N, sk));
            }
        } else if (!validate) {
            rc = pairwise_f32(ctx, (const float *)ctx->ws_p.ptr, s1, ldp, 0, 0, mode, nullptr, 0, D_dev, N, 0, sk,
                              (int32_t)band_tile[b], (int32_t)band_tile[b + 1]);
            if (rc) return rc;
            if (do_refine) {   // counting scan of the band: which pairs can the fp32 operands not certify?
                rc = refine(ctx, nullptr, s1, ldp, 0, 0, mode, stat, 0.0, nullptr, 0, D_dev, N, sk, s0, s1, b == 0,
                            FUF_CRIT_ROUNDING, nullptr, 0.0, 0);
                if (rc) return rc;
            }
            if (band_d2h && !use_gram) {
                // finished: columns [s0, s1) of rows [0, s1), and (mirror) rows [s0, s1) left of the band
                FUF_CUDA(cudaEventRecord(e_band[nb + b], sk));
                FUF_CUDA(cudaStreamWaitEvent(sd, e_band[nb + b], 0));
                FUF_RC(copy2d_async(D_h + s0, (size_t)ldd * sizeof(double), D_dev + s0, (size_t)N * sizeof(double),
                                           (size_t)(s1 - s0) * sizeof(double), (size_t)s1, sd));
                if (s0 > 0)
                    FUF_RC(copy2d_async(D_h + (size_t)s0 * ldd, (size_t)ldd * sizeof(double),
                                               D_dev + (size_t)s0 * N, (size_t)N * sizeof(double),
                                               (size_t)s0 * sizeof(double), (size_t)(s1 - s0), sd));
            }
        }
    }
    if (validate) rc = pairwise_f64(ctx, (const double *)ctx->ws_p.ptr, N, ldp, 0, 0, mode, nullptr, 0, D_dev, N, sk);
    if (band_d2h) {  // the compute stream's end event waits for the last band copy
        FUF_CUDA(cudaEventRecord(e_band[0], sd));
        FUF_CUDA(cudaStreamWaitEvent(sk, e_band[0], 0));
    }
    if (rc) return rc;
    if (!band_d2h)
        FUF_RC(copy2d_async(D_h, (size_t)ldd * sizeof(double), D_dev, (size_t)N * sizeof(double),
                                   (size_t)N * sizeof(double), (size_t)N, sk));
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
        if (v > 0) {
            // Some pairs (near-duplicate samples) need float64 operands.  Only now are the float64 profiles produced
            // — one more push-up of the resident X, in the reference's operation order — and the flagged pairs
            // recomputed; beyond the budget (e.g. --unweighted --L2, or a data set of near-duplicates) the float64 tile
            // kernel redoes the whole matrix instead.  The finished matrix then crosses PCIe once more.
            if ((rc = ctx->ws_p64t.reserve((size_t)n * ldp * sizeof(double)))) return rc;
            double *P64T = (double *)ctx->ws_p64t.ptr;
            rc = push_up_f32(ctx, ctx->ws_x.ptr, x_dtype, N, ldx_dev, mode, nullptr, nullptr, 0, P64T, ldp, nullptr, 0, sk);
            if (rc) return rc;
            if ((int64_t)v > refine_budget) {
                rc = pairwise_f64(ctx, P64T, N, ldp, 0, 0, mode, nullptr, 0, D_dev, N, sk);
            } else if (use_gram) {
                double *qg = stat + ldp, *eg = qg + ldp;
                rc = refine(ctx, P64T, N, ldp, 0, 0, mode, eg, 0.0, nullptr, 0, D_dev, N, sk, 0, N, true, FUF_CRIT_ROUNDING,
                            qg, 2.9e-3, 0);
            } else {
                rc = refine(ctx, P64T, N, ldp, 0, 0, mode, stat, 0.0, nullptr, 0, D_dev, N, sk, 0, N, true,
                            FUF_CRIT_ROUNDING, nullptr, 0.0, 0);
            }
            if (rc) return rc;
            FUF_RC(copy2d_async(D_h, (size_t)ldd * sizeof(double), D_dev, (size_t)N * sizeof(double),
                                       (size_t)N * sizeof(double), (size_t)N, sk));
            FUF_CUDA(cudaStreamSynchronize(sk));
        }
    }
    drain.armed = false;   // sk and sc are idle, and sk waited for the last copy on sd
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
