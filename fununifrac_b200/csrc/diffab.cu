// diffab.cu — per-pair differential-abundance rows (--diffab).
//
// Replaces get_L1_diffab / get_L2_diffab (src/objects/profile_vector.py:13-15,21-23) evaluated for every
// pair in itertools.combinations order (src/algorithms/emd_unifrac.py:49,56-67):
//     out[p][k] = P_i[k] - P_j[k]      (L1, signed, already length-weighted)
//                 (P_i[k] - P_j[k])^2  (L2)
// The host turns the dense rows into the reference's COO (and its negated transpose).
//
// K4 diffab_kernel (bound by the link to wherever `out` lives: HBM, or PCIe for pinned host memory):
// 32 consecutive pairs form a block (within a row i the j are consecutive); the node-major profiles are read
// coalesced along j, transposed through shared memory, and written coalesced along k (128 B per warp store).
// With float64 profiles and a float32 output the subtraction is done in double and rounded once.
#include "common.cuh"

namespace fuf {

// pair index (itertools.combinations order) -> (i, j): row i owns pairs [off(i), off(i+1)), off(i) = i N - i (i+1) / 2
__device__ __forceinline__ void pair_of_index(int64_t p, int64_t N, int32_t &i_out, int32_t &j_out)
{
    const double b = 2.0 * (double)N - 1.0;
    int64_t i = (int64_t)floor((b - sqrt(b * b - 8.0 * (double)p)) * 0.5);
    i = max((int64_t)0, min(i, N - 2));
    while (i > 0 && i * N - i * (i + 1) / 2 > p) i--;
    while ((i + 1) * N - (i + 1) * (i + 2) / 2 <= p) i++;
    i_out = (int32_t)i;
    j_out = (int32_t)(p - (i * N - i * (i + 1) / 2) + i + 1);
}

// One block = 32 consecutive pairs x 32 consecutive nodes; the pair -> (i, j) map is evaluated in the kernel, so the
// launch needs no host-built table (and no synchronisation to upload one).
template <typename TP, typename TO, int MODE>
__global__ void __launch_bounds__(256)
diffab_kernel(int64_t pair_begin, int64_t pair_end, int64_t N, const TP *__restrict__ P, int64_t ldp, int32_t n,
              TO *__restrict__ out)
{
    __shared__ TP tile[32][33];
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int64_t p0 = pair_begin + (int64_t)blockIdx.x * 32;
    const int count = (int)min((int64_t)32, pair_end - p0);
    const int32_t k0 = blockIdx.y * 32;
    int32_t i = 0, j = 0;
    if (lx < count) pair_of_index(p0 + lx, N, i, j);
    for (int kk = ly; kk < 32; kk += 8) {
        const int32_t k = k0 + kk;
        TP d = TP(0);
        if (k < n && lx < count) {
            const TP a = P[(int64_t)k * ldp + i];
            const TP q = P[(int64_t)k * ldp + j];
            d = a - q;
            if (MODE == FUF_MODE_L2) d = d * d;
        }
        tile[kk][lx] = d;
    }
    __syncthreads();
    const int32_t k = k0 + lx;
    if (k < n)
        for (int q = ly; q < count; q += 8) out[(p0 - pair_begin + q) * (int64_t)n + k] = (TO)tile[lx][q];
}

int diffab_block(fuf_ctx *ctx, const void *PT, int p_dtype, int64_t N, int64_t ldp, int mode, int64_t pair_begin,
                 int64_t pair_end, void *out, int out_dtype, cudaStream_t stream)
{
    const int64_t total = N * (N - 1) / 2;
    FUF_REQUIRE(pair_begin >= 0 && pair_begin <= pair_end && pair_end <= total,
                "fuf_diffab_block: pair range [%lld,%lld) outside [0,%lld)", (long long)pair_begin,
                (long long)pair_end, (long long)total);
    if (pair_begin == pair_end) return FUF_OK;
    const int64_t n_blocks = (pair_end - pair_begin + 31) / 32;
    FUF_REQUIRE(n_blocks < ((int64_t)1 << 31), "fuf_diffab_block: too many pairs for one call");
    dim3 grid((unsigned)n_blocks, (unsigned)((ctx->n + 31) / 32));
    FUF_REQUIRE(grid.y <= 65535, "fuf_diffab_block: tree too large");
    const int32_t n = ctx->n;
#define LAUNCH(TP, TO)                                                                                       \
    do {                                                                                                     \
        if (mode == FUF_MODE_L1)                                                                             \
            diffab_kernel<TP, TO, FUF_MODE_L1><<<grid, 256, 0, stream>>>(pair_begin, pair_end, N, (const TP *)PT, ldp, n, (TO *)out); \
        else                                                                                                 \
            diffab_kernel<TP, TO, FUF_MODE_L2><<<grid, 256, 0, stream>>>(pair_begin, pair_end, N, (const TP *)PT, ldp, n, (TO *)out); \
    } while (0)
    if (p_dtype == FUF_F32 && out_dtype == FUF_F32) LAUNCH(float, float);
    else if (p_dtype == FUF_F32 && out_dtype == FUF_F64) LAUNCH(float, double);
    else if (p_dtype == FUF_F64 && out_dtype == FUF_F32) LAUNCH(double, float);
    else LAUNCH(double, double);
#undef LAUNCH
    FUF_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return FUF_OK;
}

// K4b diffab_gather_kernel: the sub-block T[rows_i, rows_j, nodes] of the antisymmetric diffab tensor the reference
// materialises as COO minus its transpose (src/algorithms/emd_unifrac.py:56-72) and slices through
// DiffabArrayIndexer.get_diffab / get_diffab_for_node (src/utility/differential_abundance.py:18-55).  The store is the
// pushed matrix itself (N x n instead of N x N x n); same transpose-through-shared-memory tile as K4.
template <typename TP, typename TO, int MODE>
__global__ void __launch_bounds__(256)
diffab_gather_kernel(const TP *__restrict__ P, int64_t ldp, const int32_t *__restrict__ rows_i,
                     const int32_t *__restrict__ rows_j, int32_t nb, const int32_t *__restrict__ nodes, int32_t nk,
                     TO *__restrict__ out)
{
    __shared__ TP tile[32][33];
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int32_t c0 = blockIdx.x * 32, b0 = blockIdx.y * 32, a = blockIdx.z;
    const int32_t i = rows_i[a];
    const int32_t b = b0 + lx;
    const int32_t j = b < nb ? rows_j[b] : -1;
    for (int cc = ly; cc < 32; cc += 8) {
        const int32_t c = c0 + cc;
        TP d = TP(0);
        if (c < nk && j >= 0) {
            const int64_t k = nodes ? nodes[c] : c;
            d = P[k * ldp + i] - P[k * ldp + j];
            if (MODE == FUF_MODE_L2) d = (j > i) ? d * d : -(d * d);   // i == j: d == 0 either way
        }
        tile[cc][lx] = d;
    }
    __syncthreads();
    const int32_t c = c0 + lx;
    const int32_t cnt = min(32, nb - b0);
    if (c < nk)
        for (int q = ly; q < cnt; q += 8) out[((int64_t)a * nb + b0 + q) * nk + c] = (TO)tile[lx][q];
}

int diffab_gather(fuf_ctx *ctx, const void *PT, int p_dtype, int64_t ldp, int mode, const int32_t *rows_i, int64_t na,
                  const int32_t *rows_j, int64_t nb, const int32_t *nodes, int64_t nk, void *out, int out_dtype,
                  cudaStream_t stream)
{
    if (na == 0 || nb == 0 || nk == 0) return FUF_OK;
    dim3 grid((unsigned)((nk + 31) / 32), (unsigned)((nb + 31) / 32), (unsigned)na);
    FUF_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "fuf_diffab_gather: more than 65535 rows_i or 2097120 rows_j");
#define LAUNCH(TP, TO)                                                                                       \
    do {                                                                                                     \
        if (mode == FUF_MODE_L1)                                                                             \
            diffab_gather_kernel<TP, TO, FUF_MODE_L1><<<grid, 256, 0, stream>>>(                             \
                (const TP *)PT, ldp, rows_i, rows_j, (int32_t)nb, nodes, (int32_t)nk, (TO *)out);            \
        else                                                                                                 \
            diffab_gather_kernel<TP, TO, FUF_MODE_L2><<<grid, 256, 0, stream>>>(                             \
                (const TP *)PT, ldp, rows_i, rows_j, (int32_t)nb, nodes, (int32_t)nk, (TO *)out);            \
    } while (0)
    if (p_dtype == FUF_F32 && out_dtype == FUF_F32) LAUNCH(float, float);
    else if (p_dtype == FUF_F32 && out_dtype == FUF_F64) LAUNCH(float, double);
    else if (p_dtype == FUF_F64 && out_dtype == FUF_F32) LAUNCH(double, float);
    else LAUNCH(double, double);
#undef LAUNCH
    FUF_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return FUF_OK;
}

}  // namespace fuf

using namespace fuf;

extern "C" int fuf_diffab_block(fuf_ctx *ctx, const void *PT, int p_dtype, int64_t N, int64_t ldp, int mode,
                                int64_t pair_begin, int64_t pair_end, void *out, int out_dtype, void *stream)
{
    FUF_REQUIRE(ctx && PT && out, "fuf_diffab_block: null argument");
    FUF_REQUIRE(mode == FUF_MODE_L1 || mode == FUF_MODE_L2, "fuf_diffab_block: bad mode %d", mode);
    FUF_REQUIRE((p_dtype == FUF_F32 || p_dtype == FUF_F64) && (out_dtype == FUF_F32 || out_dtype == FUF_F64),
                "fuf_diffab_block: bad dtype");
    FUF_REQUIRE(N >= 0 && ldp >= N, "fuf_diffab_block: bad shape");
    FUF_CUDA(cudaSetDevice(ctx->device));
    return diffab_block(ctx, PT, p_dtype, N, ldp, mode, pair_begin, pair_end, out, out_dtype, (cudaStream_t)stream);
}

extern "C" int fuf_diffab_gather(fuf_ctx *ctx, const void *PT, int p_dtype, int64_t N, int64_t ldp, int mode,
                                 const int32_t *rows_i, int64_t na, const int32_t *rows_j, int64_t nb,
                                 const int32_t *nodes, int64_t nk, void *out, int out_dtype, void *stream)
{
    FUF_REQUIRE(ctx, "fuf_diffab_gather: null context");
    if (na == 0 || nb == 0 || nk == 0) return FUF_OK;
    FUF_REQUIRE(PT && out && rows_i && rows_j, "fuf_diffab_gather: null argument");
    FUF_REQUIRE(mode == FUF_MODE_L1 || mode == FUF_MODE_L2, "fuf_diffab_gather: bad mode %d", mode);
    FUF_REQUIRE((p_dtype == FUF_F32 || p_dtype == FUF_F64) && (out_dtype == FUF_F32 || out_dtype == FUF_F64),
                "fuf_diffab_gather: bad dtype");
    FUF_REQUIRE(N >= 0 && ldp >= N && na >= 0 && nb >= 0 && nk >= 0, "fuf_diffab_gather: bad shape");
    FUF_REQUIRE(nodes || nk == ctx->n, "fuf_diffab_gather: nodes == NULL needs nk == n (%d), got %lld", ctx->n,
                (long long)nk);
    FUF_CUDA(cudaSetDevice(ctx->device));
    return diffab_gather(ctx, PT, p_dtype, ldp, mode, rows_i, na, rows_j, nb, nodes, nk, out, out_dtype,
                         (cudaStream_t)stream);
}
