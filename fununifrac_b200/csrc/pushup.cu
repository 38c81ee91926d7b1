// pushup.cu — push-up kernels: per-sample abundances X -> per-edge cumulative mass P (node-major).
//
// Replaces push_up_L1 / push_up_L2 of the reference (src/algorithms/emd_unifrac.py:20-39):
//     for i in range(n-1): P[Tint[i]] += P[i]; P[i] *= w_i        (w = L or sqrt(L), 0 -> eps)
// i.e. P[k] = w_k * (sum of X over the subtree of k), the root unscaled.  In a DFS postorder the
// subtree of k is the contiguous index range [start_k, k], which is what the production kernel uses.
//
// Production path (HBM-bound, one pass over X and one over P; algorithmic bytes = sizeof(X)+4 per entry):
//   K1a pushup_chunk_kernel   one CTA = 128 samples x one chunk of 32 consecutive nodes.  X rows are read
//                             coalesced (128 B / 256 B per warp request) and transposed through shared memory;
//                             each thread then owns one sample and replays the reference loop restricted to
//                             the chunk with double accumulators in shared memory.  Nodes whose subtree lies
//                             inside the chunk are final and written (coalesced along samples) as float.
//                             For the rest ("spanning" nodes, a few hundred of 30K) the chunk leaves behind
//                             its total mass T[chunk] and chunk-local prefix sums at checkpoint positions.
//   K1b pushup_scan_kernel    exclusive scan of T over chunks (per sample) -> global prefix at chunk starts.
//   K1c pushup_span_kernel    spanning node k: S = prefix(k) - prefix(start_k - 1); writes w_k * S.
//
// Validation path (fuf_push_up_f64): transpose to node-major double, then one thread per sample replays the
// reference loop verbatim -> bit-identical to the reference's float64 result.
#include "common.cuh"

namespace fuf {

constexpr int KT = kPushChunk;  // nodes per chunk
constexpr int TI = 128;         // samples per CTA

template <typename TX>
__global__ void __launch_bounds__(TI)
pushup_chunk_kernel(const TX *__restrict__ X, int64_t ldx, int64_t N, int32_t n, const double *__restrict__ w,
                    const int32_t *__restrict__ parent_local, const int32_t *__restrict__ cp_after,
                    const uint8_t *__restrict__ is_span, float *__restrict__ PT, int64_t ldp, int64_t col0,
                    double *__restrict__ T, double *__restrict__ CP, int64_t ldt)
{
    // dynamic shared memory: acc[KT][TI] double | xs[KT][TI+1] TX   (48.5 KB float, 65 KB double)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double(*acc)[TI] = reinterpret_cast<double(*)[TI]>(smem_raw);
    TX(*xs)[TI + 1] = reinterpret_cast<TX(*)[TI + 1]>(smem_raw + sizeof(double) * KT * TI);
    __shared__ double s_w[KT];
    __shared__ int32_t s_pl[KT], s_cp[KT];
    __shared__ uint8_t s_span[KT];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int32_t k0 = blockIdx.x * KT;
    const int64_t i0 = (int64_t)blockIdx.y * TI;
    const int kt = min(KT, n - k0);

    if (tid < KT) {
        int32_t k = k0 + tid;
        bool ok = tid < kt;
        s_w[tid] = ok ? w[k] : 0.0;
        s_pl[tid] = ok ? parent_local[k] : -1;
        s_cp[tid] = ok ? cp_after[k] : -1;
        s_span[tid] = ok ? is_span[k] : 1;
    }
    // coalesced load of the X tile: one warp request = one sample row segment of KT entries
#pragma unroll 4
    for (int it = 0; it < TI / 4; it++) {
        int row = it * 4 + warp;
        int64_t s = i0 + row;
        TX v = TX(0);
        if (s < N && lane < kt) v = X[s * ldx + k0 + lane];
        xs[lane][row] = v;
    }
#pragma unroll
    for (int kk = 0; kk < KT; kk++) acc[kk][tid] = 0.0;
    __syncthreads();

    const int64_t s = i0 + tid;
    if (s >= N) return;
    double r = 0.0;  // running sum of X over the chunk (chunk-local prefix)
    float *out = PT + col0 + s;
    for (int kk = 0; kk < kt; kk++) {
        double x = (double)xs[kk][tid];
        double S = x + acc[kk][tid];
        r += x;
        int pl = s_pl[kk];
        if (pl >= 0) acc[pl][tid] += S;
        if (!s_span[kk]) out[(int64_t)(k0 + kk) * ldp] = (float)(s_w[kk] * S);
        int cp = s_cp[kk];
        if (cp >= 0) CP[(int64_t)cp * ldt + s] = r;
    }
    if (T) T[(int64_t)blockIdx.x * ldt + s] = r;
}

// T[c][s] (chunk totals) -> exclusive prefix over chunks, in place.
__global__ void pushup_scan_kernel(double *__restrict__ T, int32_t n_chunks, int64_t N, int64_t ldt)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    double run = 0.0;
    int c = 0;
    for (; c + 8 <= n_chunks; c += 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = T[(int64_t)(c + u) * ldt + s];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            T[(int64_t)(c + u) * ldt + s] = run;
            run += v[u];
        }
    }
    for (; c < n_chunks; c++) {
        double v = T[(int64_t)c * ldt + s];
        T[(int64_t)c * ldt + s] = run;
        run += v;
    }
}

__global__ void pushup_span_kernel(const SpanNode *__restrict__ span, int32_t n_span, const double *__restrict__ w,
                                   const double *__restrict__ CT, const double *__restrict__ CP, int64_t ldt,
                                   int64_t N, float *__restrict__ PT, int64_t ldp, int64_t col0)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    for (int q = blockIdx.y; q < n_span; q += gridDim.y) {
        SpanNode sn = span[q];
        double hi = CT[(int64_t)sn.chunk_last * ldt + s] + CP[(int64_t)sn.slot_k * ldt + s];
        double lo = CT[(int64_t)sn.chunk_first * ldt + s];
        if (sn.slot_start >= 0) lo += CP[(int64_t)sn.slot_start * ldt + s];
        PT[(int64_t)sn.k * ldp + col0 + s] = (float)(w[sn.k] * (hi - lo));
    }
}

// ---------------------------------------------------------------------------------------------
// validation path
// ---------------------------------------------------------------------------------------------
__global__ void transpose_f64_kernel(const double *__restrict__ X, int64_t ldx, int64_t N, int32_t n,
                                     double *__restrict__ PT, int64_t ldp, int64_t col0)
{
    __shared__ double tile[32][33];
    int64_t s0 = (int64_t)blockIdx.y * 32;
    int32_t k0 = blockIdx.x * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int64_t s = s0 + r;
        int32_t k = k0 + threadIdx.x;
        tile[r][threadIdx.x] = (s < N && k < n) ? X[s * ldx + k] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int32_t k = k0 + r;
        int64_t s = s0 + threadIdx.x;
        if (k < n && s < N) PT[(int64_t)k * ldp + col0 + s] = tile[threadIdx.x][r];
    }
}

// sample-major rows -> node-major columns with a dtype change (already pushed profiles).
template <typename TI_, typename TO_>
__global__ void transpose_cast_kernel(const TI_ *__restrict__ X, int64_t ldx, int64_t N, int32_t n,
                                      TO_ *__restrict__ PT, int64_t ldp, int64_t col0)
{
    __shared__ TI_ tile[32][33];
    int64_t s0 = (int64_t)blockIdx.y * 32;
    int32_t k0 = blockIdx.x * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int64_t s = s0 + r;
        int32_t k = k0 + threadIdx.x;
        tile[r][threadIdx.x] = (s < N && k < n) ? X[s * ldx + k] : TI_(0);
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int32_t k = k0 + r;
        int64_t s = s0 + threadIdx.x;
        if (k < n && s < N) PT[(int64_t)k * ldp + col0 + s] = (TO_)tile[threadIdx.x][r];
    }
}

// One thread per sample replays the reference loop in its exact operation order.
__global__ void pushup_exact_kernel(double *__restrict__ PT, int64_t ldp, int64_t col0, int64_t N, int32_t n,
                                    const int32_t *__restrict__ parent, const double *__restrict__ w)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    double *col = PT + col0 + s;
    for (int32_t i = 0; i < n - 1; i++) {
        double v = col[(int64_t)i * ldp];
        col[(int64_t)parent[i] * ldp] += v;  // push mass up
        col[(int64_t)i * ldp] = v * w[i];
    }
}

__global__ void f64_to_f32_kernel(const double *__restrict__ src, float *__restrict__ dst, int64_t ld_src,
                                  int64_t ld_dst, int64_t col0, int64_t N, int32_t n)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int32_t k = blockIdx.y;
    if (s < N && k < n) dst[(int64_t)k * ld_dst + col0 + s] = (float)src[(int64_t)k * ld_src + s];
}

int push_up_f64_exact(fuf_ctx *ctx, const double *X, int64_t N, int64_t ldx, int mode, double *P64T, int64_t ldp,
                      int64_t col0, cudaStream_t stream)
{
    if (N == 0) return FUF_OK;
    dim3 tb(32, 8), tg((ctx->n + 31) / 32, (unsigned)((N + 31) / 32));
    transpose_f64_kernel<<<tg, tb, 0, stream>>>(X, ldx, N, ctx->n, P64T, ldp, col0);
    FUF_CUDA(cudaGetLastError());
    pushup_exact_kernel<<<(unsigned)((N + 63) / 64), 64, 0, stream>>>(P64T, ldp, col0, N, ctx->n, ctx->d_parent,
                                                                        ctx->d_w[mode]);
    FUF_CUDA(cudaGetLastError());
    ctx->launches += 2;
    return FUF_OK;
}

int push_up_f32(fuf_ctx *ctx, const void *X, int x_dtype, int64_t N, int64_t ldx, int mode, float *PT, int64_t ldp,
                int64_t col0, cudaStream_t stream)
{
    if (N == 0) return FUF_OK;
    const int32_t n = ctx->n;
    if (!ctx->contiguous) {
        // Not a DFS postorder (hand-built input): subtree ranges are not contiguous.  Use the exact kernel
        // (any child<parent order) in double and round once.
        FUF_REQUIRE(x_dtype == FUF_F64, "fuf_push_up: a tree that is not in DFS postorder needs float64 input");
        int rc = ctx->ws_push.reserve((size_t)n * (size_t)N * sizeof(double));
        if (rc) return rc;
        double *tmp = (double *)ctx->ws_push.ptr;
        rc = push_up_f64_exact(ctx, (const double *)X, N, ldx, mode, tmp, N, 0, stream);
        if (rc) return rc;
        dim3 g((unsigned)((N + 255) / 256), n);
        f64_to_f32_kernel<<<g, 256, 0, stream>>>(tmp, PT, N, ldp, col0, N, n);
        FUF_CUDA(cudaGetLastError());
        ctx->launches += 1;
        return FUF_OK;
    }
    const int64_t ldt = (N + 31) / 32 * 32;
    double *T = nullptr, *CP = nullptr;
    if (ctx->n_span > 0) {
        size_t need = ((size_t)ctx->n_chunks + (size_t)ctx->n_cp) * (size_t)ldt * sizeof(double);
        int rc = ctx->ws_push.reserve(need);
        if (rc) return rc;
        T = (double *)ctx->ws_push.ptr;
        CP = T + (size_t)ctx->n_chunks * ldt;
    }
    FUF_REQUIRE((N + TI - 1) / TI <= 65535, "fuf_push_up: N = %lld too large for one call (max %d)", (long long)N,
                65535 * TI);
    dim3 grid(ctx->n_chunks, (unsigned)((N + TI - 1) / TI));
    const size_t smem_f = sizeof(double) * KT * TI + sizeof(float) * KT * (TI + 1);
    const size_t smem_d = sizeof(double) * KT * TI + sizeof(double) * KT * (TI + 1);
    static bool attr_set = false;
    if (!attr_set) {
        FUF_CUDA(cudaFuncSetAttribute(pushup_chunk_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem_f));
        FUF_CUDA(cudaFuncSetAttribute(pushup_chunk_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem_d));
        attr_set = true;
    }
    if (x_dtype == FUF_F32)
        pushup_chunk_kernel<float><<<grid, TI, smem_f, stream>>>((const float *)X, ldx, N, n, ctx->d_w[mode],
                                                            ctx->d_parent_local, ctx->d_cp_after, ctx->d_is_span, PT,
                                                            ldp, col0, T, CP, ldt);
    else
        pushup_chunk_kernel<double><<<grid, TI, smem_d, stream>>>((const double *)X, ldx, N, n, ctx->d_w[mode],
                                                             ctx->d_parent_local, ctx->d_cp_after, ctx->d_is_span, PT,
                                                             ldp, col0, T, CP, ldt);
    FUF_CUDA(cudaGetLastError());
    ctx->launches += 1;
    if (ctx->n_span > 0) {
        pushup_scan_kernel<<<(unsigned)((N + 127) / 128), 128, 0, stream>>>(T, ctx->n_chunks, N, ldt);
        FUF_CUDA(cudaGetLastError());
        dim3 g2((unsigned)((N + 127) / 128), (unsigned)min(ctx->n_span, 1024));
        pushup_span_kernel<<<g2, 128, 0, stream>>>(ctx->d_span, ctx->n_span, ctx->d_w[mode], T, CP, ldt, N, PT, ldp,
                                                   col0);
        FUF_CUDA(cudaGetLastError());
        ctx->launches += 2;
    }
    return FUF_OK;
}

}  // namespace fuf

using namespace fuf;

extern "C" int fuf_push_up(fuf_ctx *ctx, const void *X, int x_dtype, int64_t N, int64_t ldx, int mode, float *PT,
                           int64_t ldp, int64_t col0, void *stream)
{
    FUF_REQUIRE(ctx && X && PT, "fuf_push_up: null argument");
    FUF_REQUIRE(mode == FUF_MODE_L1 || mode == FUF_MODE_L2, "fuf_push_up: bad mode %d", mode);
    FUF_REQUIRE(x_dtype == FUF_F32 || x_dtype == FUF_F64, "fuf_push_up: bad x_dtype %d", x_dtype);
    FUF_REQUIRE(N >= 0 && ldx >= ctx->n && col0 >= 0 && col0 + N <= ldp,
                "fuf_push_up: bad shape N=%lld ldx=%lld ldp=%lld col0=%lld n=%d", (long long)N, (long long)ldx,
                (long long)ldp, (long long)col0, ctx->n);
    FUF_CUDA(cudaSetDevice(ctx->device));
    return push_up_f32(ctx, X, x_dtype, N, ldx, mode, PT, ldp, col0, (cudaStream_t)stream);
}

extern "C" int fuf_push_up_f64(fuf_ctx *ctx, const double *X, int64_t N, int64_t ldx, int mode, double *P64T,
                               int64_t ldp, int64_t col0, void *stream)
{
    FUF_REQUIRE(ctx && X && P64T, "fuf_push_up_f64: null argument");
    FUF_REQUIRE(mode == FUF_MODE_L1 || mode == FUF_MODE_L2, "fuf_push_up_f64: bad mode %d", mode);
    FUF_REQUIRE(N >= 0 && ldx >= ctx->n && col0 >= 0 && col0 + N <= ldp,
                "fuf_push_up_f64: bad shape N=%lld ldx=%lld ldp=%lld col0=%lld n=%d", (long long)N, (long long)ldx,
                (long long)ldp, (long long)col0, ctx->n);
    FUF_CUDA(cudaSetDevice(ctx->device));
    return push_up_f64_exact(ctx, X, N, ldx, mode, P64T, ldp, col0, (cudaStream_t)stream);
}

extern "C" int fuf_load_pushed(fuf_ctx *ctx, const void *P_rows, int p_dtype, int64_t N, int64_t ld, void *PT,
                               int out_dtype, int64_t ldp, int64_t col0, void *stream)
{
    FUF_REQUIRE(ctx && P_rows && PT, "fuf_load_pushed: null argument");
    FUF_REQUIRE((p_dtype == FUF_F32 || p_dtype == FUF_F64) && (out_dtype == FUF_F32 || out_dtype == FUF_F64),
                "fuf_load_pushed: bad dtype");
    FUF_REQUIRE(N >= 0 && ld >= ctx->n && col0 >= 0 && col0 + N <= ldp, "fuf_load_pushed: bad shape");
    if (N == 0) return FUF_OK;
    FUF_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    dim3 tb(32, 8), tg((ctx->n + 31) / 32, (unsigned)((N + 31) / 32));
    FUF_REQUIRE(tg.y <= 65535, "fuf_load_pushed: N too large for one call");
    if (p_dtype == FUF_F32 && out_dtype == FUF_F32)
        transpose_cast_kernel<float, float><<<tg, tb, 0, st>>>((const float *)P_rows, ld, N, ctx->n, (float *)PT, ldp, col0);
    else if (p_dtype == FUF_F32)
        transpose_cast_kernel<float, double><<<tg, tb, 0, st>>>((const float *)P_rows, ld, N, ctx->n, (double *)PT, ldp, col0);
    else if (out_dtype == FUF_F32)
        transpose_cast_kernel<double, float><<<tg, tb, 0, st>>>((const double *)P_rows, ld, N, ctx->n, (float *)PT, ldp, col0);
    else
        transpose_cast_kernel<double, double><<<tg, tb, 0, st>>>((const double *)P_rows, ld, N, ctx->n, (double *)PT, ldp, col0);
    FUF_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return FUF_OK;
}
