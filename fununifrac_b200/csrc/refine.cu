// refine.cu — float64 refinement of the pairs the fp32 production path cannot certify to 1e-6 relative.
//
// The production kernel (pairwise.cu) evaluates get_L1 / get_L2 (src/objects/profile_vector.py:9-19) on pushed
// profiles that were rounded to fp32.  For a pair (i, j) that rounding perturbs the distance by at most
//     L1:  rho_i + rho_j                      rho_s = sum_k |P_s[k] - fl32(P_s[k])|      (triangle inequality)
//     L2:  2 sqrt(D) (r_i + r_j) + (r_i+r_j)^2   r_s  = sqrt(sum_k (P_s[k] - fl32(P_s[k]))^2)  (Cauchy-Schwarz)
// (rho, r^2 are the per-sample statistics the push-up kernel emits).  The bound is only a problem for
// near-duplicate samples, whose distance is tiny compared with their mass.  K6 refine_kernel scans the distance
// matrix once, and wherever the bound exceeds 0.5e-6 of the computed distance,
//     L1:  D < kappa (rho_i + rho_j),   kappa = 2e6        L2:  D < kappa (r_i + r_j)^2,   kappa = 1.6e13
// it recomputes the pair from the float64 pushed profiles P64T (sum over all n entries in double) and overwrites
// D[i,j] (and D[j,i]).  Cost: one pass over the upper triangle of D plus ~2 n sector reads per refined pair
// (~1 us): beyond `max_pairs` flagged pairs the kernel only counts, and the caller recomputes the whole matrix with
// the float64 tile kernel (3 ns per pair) instead.
#include "common.cuh"

namespace fuf {

constexpr int RSEG = 256;  // columns of one row a warp examines per work item

struct RefineParams {
    const double *P64T;
    int64_t ldp64, shard, shard_stride;
    int32_t n;
    int64_t N;
    const double *stat;
    double kappa;
    const int32_t *tile_rows;  // nullptr: full matrix; else compact row blocks
    int32_t n_rows_total;      // rows to scan (N, or n_tile_rows * 128)
    double *D;
    int64_t ldd;
    int mirror;                // also write D[j,i]
    int64_t j_begin, j_end;    // only pairs with j in [j_begin, j_end) are examined
    int linear;                // criterion D < kappa (s_i + s_j) on the raw statistic, whatever the mode
    const double *stat2;       // optional second statistic: additional linear criterion D < kappa2 (t_i + t_j)
    double kappa2;
    unsigned long long budget; // pairs beyond this many flagged ones are counted but not recomputed
    unsigned long long *counter;
};

__device__ __forceinline__ const double *col_ptr(const RefineParams &p, int64_t s)
{
    if (p.shard == 0) return p.P64T + s;
    return p.P64T + (s / p.shard) * p.shard_stride + (s % p.shard);
}

template <int MODE>
__global__ void __launch_bounds__(256)
refine_kernel(const RefineParams p)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t segs_per_row = (p.j_end - p.j_begin + RSEG - 1) / RSEG;
    const int64_t n_items = (int64_t)p.n_rows_total * segs_per_row;
    const int64_t kstride = p.shard == 0 ? p.ldp64 : p.shard;
    for (int64_t item = warp; item < n_items; item += n_warps) {
        const int64_t row = item / segs_per_row, seg = item % segs_per_row;
        int64_t gi = row;
        if (p.tile_rows) {
            const int32_t t = p.tile_rows[row / FUF_TILE];
            if (t < 0) continue;
            gi = (int64_t)t * FUF_TILE + row % FUF_TILE;
        }
        if (gi >= p.N) continue;
        const int64_t j0 = p.j_begin + seg * RSEG;
        if (j0 + RSEG <= gi + 1) continue;  // entirely on or below the diagonal
        const bool square = (MODE == FUF_MODE_L2) && !p.linear;
        const double si = square ? sqrt(p.stat[gi]) : p.stat[gi];
        const double ti = p.stat2 ? p.stat2[gi] : 0.0;
        double *drow = p.D + row * p.ldd;
        // all the loads of the segment first (8 independent 256-byte requests per warp in flight), then the tests
        double dv[RSEG / 32], sv[RSEG / 32], tv[RSEG / 32];
#pragma unroll
        for (int u = 0; u < RSEG / 32; u++) {
            const int64_t gj = j0 + u * 32 + lane;
            const bool in = gj > gi && gj < p.j_end;
            dv[u] = in ? drow[gj] : 0.0;
            sv[u] = in ? p.stat[gj] : 0.0;
            tv[u] = (in && p.stat2) ? p.stat2[gj] : 0.0;
        }
        unsigned flags = 0;
#pragma unroll
        for (int u = 0; u < RSEG / 32; u++) {
            const int64_t gj = j0 + u * 32 + lane;
            bool flag = false;
            if (gj > gi && gj < p.j_end) {
                const double d = dv[u];
                const double sj = square ? sqrt(sv[u]) : sv[u];
                const double b = si + sj;
                flag = square ? !(d >= p.kappa * b * b) : !(d >= p.kappa * b);
                if (b == 0.0) flag = false;  // both samples exactly representable in fp32: nothing to gain
                if (p.stat2) flag = flag || !(d >= p.kappa2 * (ti + tv[u]));
            }
            flags |= (flag ? 1u : 0u) << u;
        }
        if (__ballot_sync(0xffffffffu, flags != 0) == 0) continue;
#pragma unroll 1
        for (int u = 0; u < RSEG / 32; u++) {
            const bool flag = (flags >> u) & 1u;
            unsigned mask = __ballot_sync(0xffffffffu, flag);
            if (mask == 0) continue;
            // the counter counts FLAGGED pairs; only the first `budget` of them are recomputed here (beyond that the
            // caller is better off recomputing everything with the float64 tile kernel)
            unsigned long long first = 0;
            if (lane == 0) first = atomicAdd(p.counter, (unsigned long long)__popc(mask));
            first = __shfl_sync(0xffffffffu, first, 0);
            if (p.P64T == nullptr) continue;  // count-only mode
            while (mask) {
                const int src = __ffs(mask) - 1;
                mask &= mask - 1;
                if (first++ >= p.budget) break;
                const int64_t j = j0 + u * 32 + src;
                const double *a = col_ptr(p, gi), *b = col_ptr(p, j);
                double acc[4] = {0.0, 0.0, 0.0, 0.0};
                int32_t k = lane;
                for (; k + 96 < p.n; k += 128) {
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const double x = a[(int64_t)(k + 32 * q) * kstride], y = b[(int64_t)(k + 32 * q) * kstride];
                        const double d = fabs(__dsub_rn(x, y));
                        acc[q] = __dadd_rn(acc[q], MODE == FUF_MODE_L2 ? __dmul_rn(d, d) : d);
                    }
                }
                for (; k < p.n; k += 32) {
                    const double d = fabs(__dsub_rn(a[(int64_t)k * kstride], b[(int64_t)k * kstride]));
                    acc[0] = __dadd_rn(acc[0], MODE == FUF_MODE_L2 ? __dmul_rn(d, d) : d);
                }
                double v = (acc[0] + acc[1]) + (acc[2] + acc[3]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) {
                    drow[j] = v;
                    if (p.mirror) p.D[j * p.ldd + gi] = v;
                }
            }
        }
    }
}

// K6l refine_list_kernel: the pairs a tile epilogue flagged (fuf_pairwise_tiles), one warp per pair, recomputed from the
// float64 profiles and written to D[i, j] and D[j, i] of the full matrix (which may live on a peer GPU or in mapped
// host memory: plain stores).
template <int MODE>
__global__ void __launch_bounds__(256)
refine_list_kernel(const long long *__restrict__ list, long long cap, const double *__restrict__ P64T, int64_t ldp64,
                   int64_t shard, int64_t shard_stride, int32_t n, double *__restrict__ D, int64_t ldd)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long count = min(list[0], cap);
    const int64_t kstride = shard == 0 ? ldp64 : shard;
    for (long long q = warp; q < count; q += n_warps) {
        const int64_t i = list[1 + 2 * q], j = list[2 + 2 * q];
        const double *a = P64T + (shard == 0 ? i : (i / shard) * shard_stride + i % shard);
        const double *b = P64T + (shard == 0 ? j : (j / shard) * shard_stride + j % shard);
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        int32_t k = lane;
        for (; k + 96 < n; k += 128) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const double d = fabs(__dsub_rn(a[(int64_t)(k + 32 * u) * kstride], b[(int64_t)(k + 32 * u) * kstride]));
                acc[u] = __dadd_rn(acc[u], MODE == FUF_MODE_L2 ? __dmul_rn(d, d) : d);
            }
        }
        for (; k < n; k += 32) {
            const double d = fabs(__dsub_rn(a[(int64_t)k * kstride], b[(int64_t)k * kstride]));
            acc[0] = __dadd_rn(acc[0], MODE == FUF_MODE_L2 ? __dmul_rn(d, d) : d);
        }
        double v = (acc[0] + acc[1]) + (acc[2] + acc[3]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) {
            D[i * ldd + j] = v;
            D[j * ldd + i] = v;
        }
    }
}

int refine(fuf_ctx *ctx, const double *P64T, int64_t N, int64_t ldp64, int64_t shard, int64_t shard_stride, int mode,
           const double *err_stat, double kappa, const int32_t *tile_rows_h, int32_t n_tile_rows, double *D,
           int64_t ldd, cudaStream_t stream, int64_t j_begin, int64_t j_end, bool reset_counter, int criterion,
           const double *lin_stat, double lin_kappa, int64_t max_pairs)
{
    if (!ctx->d_refined) FUF_CUDA(cudaMalloc((void **)&ctx->d_refined, sizeof(unsigned long long)));
    if (reset_counter) FUF_CUDA(cudaMemsetAsync(ctx->d_refined, 0, sizeof(unsigned long long), stream));
    if (N < 2) return FUF_OK;
    if (j_end <= 0 || j_end > N) j_end = N;
    if (j_begin < 0) j_begin = 0;
    if (j_begin >= j_end) return FUF_OK;
    RefineParams p;
    p.P64T = P64T;
    p.ldp64 = ldp64;
    p.shard = shard;
    p.shard_stride = shard_stride;
    p.n = ctx->n;
    p.N = N;
    p.stat = err_stat;
    p.budget = max_pairs > 0 ? (unsigned long long)max_pairs : ~0ull;
    p.stat2 = lin_stat;
    p.kappa2 = lin_kappa > 0 ? lin_kappa : 0.02;
    p.linear = (criterion == FUF_CRIT_LINEAR);
    p.kappa = kappa > 0 ? kappa : (p.linear ? 0.02 : (mode == FUF_MODE_L2 ? 1.6e13 : 2e6));
    p.D = D;
    p.ldd = ldd;
    p.counter = ctx->d_refined;
    p.j_begin = j_begin;
    p.j_end = j_end;
    if (tile_rows_h) {
        if (n_tile_rows == 0) return FUF_OK;
        void *d_rows = nullptr;
        int rc = ctx->tables.get_by_content("rows", tile_rows_h, (size_t)n_tile_rows * sizeof(int32_t), &d_rows);
        if (rc) return rc;
        p.tile_rows = (const int32_t *)d_rows;
        p.n_rows_total = n_tile_rows * FUF_TILE;
        p.mirror = 0;
    } else {
        p.tile_rows = nullptr;
        p.n_rows_total = (int32_t)j_end;   // rows i < j < j_end
        p.mirror = 1;
    }
    const int grid = ctx->sm_count * 8;
    if (mode == FUF_MODE_L1)
        refine_kernel<FUF_MODE_L1><<<grid, 256, 0, stream>>>(p);
    else
        refine_kernel<FUF_MODE_L2><<<grid, 256, 0, stream>>>(p);
    FUF_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return FUF_OK;
}

}  // namespace fuf

using namespace fuf;

extern "C" int fuf_refine(fuf_ctx *ctx, const double *P64T, int64_t N, int64_t ldp64, int64_t shard,
                          int64_t shard_stride, int mode, const double *err_stat, double kappa, int criterion,
                          const int32_t *tile_rows_h, int32_t n_tile_rows, double *D, int64_t ldd,
                          const double *lin_stat, double lin_kappa, int64_t max_pairs, void *stream)
{
    FUF_REQUIRE(ctx && err_stat && D, "fuf_refine: null argument");
    FUF_REQUIRE(mode == FUF_MODE_L1 || mode == FUF_MODE_L2, "fuf_refine: bad mode %d", mode);
    FUF_REQUIRE(criterion == FUF_CRIT_ROUNDING || criterion == FUF_CRIT_LINEAR, "fuf_refine: bad criterion %d", criterion);
    FUF_REQUIRE(N >= 0 && ldd >= N && N < ((int64_t)1 << 31), "fuf_refine: bad shape N=%lld ldd=%lld", (long long)N,
                (long long)ldd);
    FUF_REQUIRE(tile_rows_h != nullptr || n_tile_rows == 0, "fuf_refine: n_tile_rows without tile_rows_h");
    if (!P64T) {
    } else if (shard == 0)
        FUF_REQUIRE(ldp64 >= N, "fuf_refine: ldp64=%lld < N=%lld", (long long)ldp64, (long long)N);
    else
        FUF_REQUIRE(shard > 0 && shard_stride >= shard * (int64_t)ctx->n, "fuf_refine: bad shard layout");
    FUF_CUDA(cudaSetDevice(ctx->device));
    return refine(ctx, P64T, N, ldp64, shard, shard_stride, mode, err_stat, kappa, tile_rows_h, n_tile_rows, D, ldd,
                  (cudaStream_t)stream, 0, N, true, criterion, lin_stat, lin_kappa, max_pairs);
}

extern "C" int fuf_refine_list(fuf_ctx *ctx, const double *P64T, int64_t N, int64_t ldp64, int64_t shard,
                               int64_t shard_stride, int mode, const int64_t *flag_list, int64_t list_cap, double *D,
                               int64_t ldd, void *stream)
{
    FUF_REQUIRE(ctx && P64T && flag_list && D, "fuf_refine_list: null argument");
    FUF_REQUIRE(mode == FUF_MODE_L1 || mode == FUF_MODE_L2, "fuf_refine_list: bad mode %d", mode);
    FUF_REQUIRE(N >= 0 && ldd >= N && list_cap >= 0, "fuf_refine_list: bad shape");
    if (shard == 0)
        FUF_REQUIRE(ldp64 >= N, "fuf_refine_list: ldp64=%lld < N=%lld", (long long)ldp64, (long long)N);
    else
        FUF_REQUIRE(shard > 0 && shard_stride >= shard * (int64_t)ctx->n, "fuf_refine_list: bad shard layout");
    if (list_cap == 0 || N < 2) return FUF_OK;
    FUF_CUDA(cudaSetDevice(ctx->device));
    const int grid = ctx->sm_count * 4;
    if (mode == FUF_MODE_L1)
        refine_list_kernel<FUF_MODE_L1><<<grid, 256, 0, (cudaStream_t)stream>>>((const long long *)flag_list, list_cap, P64T, ldp64,
                                                                                 shard, shard_stride, ctx->n, D, ldd);
    else
        refine_list_kernel<FUF_MODE_L2><<<grid, 256, 0, (cudaStream_t)stream>>>((const long long *)flag_list, list_cap, P64T, ldp64,
                                                                                 shard, shard_stride, ctx->n, D, ldd);
    FUF_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return FUF_OK;
}

extern "C" int fuf_refine_count(fuf_ctx *ctx, int64_t *refined, void *stream)
{
    FUF_REQUIRE(ctx && refined, "fuf_refine_count: null argument");
    *refined = 0;
    if (!ctx->d_refined) return FUF_OK;
    FUF_CUDA(cudaSetDevice(ctx->device));
    unsigned long long v = 0;
    FUF_CUDA(cudaMemcpyAsync(&v, ctx->d_refined, sizeof(v), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    FUF_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    *refined = (int64_t)v;
    return FUF_OK;
}

extern "C" int fuf_refine_count_device(fuf_ctx *ctx, int64_t *count_dev, void *stream)
{
    FUF_REQUIRE(ctx && count_dev, "fuf_refine_count_device: null argument");
    FUF_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->d_refined) {
        FUF_CUDA(cudaMemsetAsync(count_dev, 0, sizeof(int64_t), (cudaStream_t)stream));
        return FUF_OK;
    }
    FUF_CUDA(cudaMemcpyAsync(count_dev, ctx->d_refined, sizeof(int64_t), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return FUF_OK;
}
