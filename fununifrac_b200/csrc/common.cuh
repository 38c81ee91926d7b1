// common.cuh — shared declarations of libfununifrac.so (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/fununifrac.h"

namespace fuf {

void set_error(const char *fmt, ...);

#define FUF_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t err__ = (call);                                                            \
        if (err__ != cudaSuccess) {                                                            \
            fuf::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call,                  \
                           cudaGetErrorString(err__));                                         \
            return FUF_ECUDA;                                                                  \
        }                                                                                      \
    } while (0)

// propagate the error code of a library-internal call (its message is already set)
#define FUF_RC(call)                                                                           \
    do {                                                                                       \
        int rc__ = (call);                                                                     \
        if (rc__ != FUF_OK) return rc__;                                                       \
    } while (0)

#define FUF_REQUIRE(cond, ...)                                                                 \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            fuf::set_error(__VA_ARGS__);                                                       \
            return FUF_EINVAL;                                                                 \
        }                                                                                      \
    } while (0)

constexpr int kPushChunk = 32;  // nodes per push-up chunk (KT)

// One spanning node = a node whose postorder range [start, k] crosses a push-up chunk boundary.
struct SpanNode {
    int32_t k;           // postorder index
    int32_t chunk_first; // chunk holding `start`
    int32_t chunk_last;  // chunk holding `k`
    int32_t slot_start;  // checkpoint slot with the chunk-local prefix sum up to start-1, -1 if start is a chunk start
    int32_t slot_k;      // checkpoint slot with the chunk-local prefix sum up to k (inclusive)
};

// Per-node record of the TMA push-up kernel (pushup.cu): the weight and the packed plan word in one 16-byte load.
//   bit 0      leaf-like: the subtree is the node itself, S = x
//   bit 1      skip: spanning node (written by the span kernel) or padding beyond the tree
//   bit 3      this node has a checkpoint slot (bits 16-31)
//   bits 4-7   park slot + 1: keep the chunk-local prefix after this node (an internal node further on starts right after
//              it); 0 = nothing to keep
//   bits 8-11  park slot + 1 of the prefix to subtract: S = prefix - parked; 0 = the subtree starts with the chunk, S = prefix
//   bits 16-31 checkpoint slot + 1 for the span kernel (0 = none)
constexpr int kPushParkSlots = 12;  // parked prefixes per chunk (24 KB of shared memory per CTA)
struct PushNode {
    double w;
    int32_t meta;
    int32_t pad;
};

// Grow-only device buffer owned by a context.
struct DeviceBuffer {
    void *ptr = nullptr;
    size_t bytes = 0;
    int reserve(size_t want);
    void release();
};

// Pinned host staging buffer.
struct PinnedBuffer {
    void *ptr = nullptr;
    size_t bytes = 0;
    int reserve(size_t want);
    void release();
};

// Small read-only device tables (tile lists, tile-row lists, slot tables), keyed by the parameters / content they
// were built from.  A table is uploaded once with a blocking copy and then reused: the launches that need it never
// synchronise the stream for a host-to-device copy of pageable planning data.
struct TableCache {
    std::unordered_map<std::string, void *> map;
    size_t bytes = 0;
    void *find(const std::string &key) const;
    // uploads `host` under `key` (a miss is assumed); *d_out = device pointer
    int put(const std::string &key, const void *host, size_t nbytes, void **d_out);
    // find, else put with the content itself as the key
    int get_by_content(const char *tag, const void *host, size_t nbytes, void **d_out);
    void release();
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda); nullptr if unavailable
EncodeTiledFn tensor_map_encoder();

template <typename T>
inline void key_append(std::string &key, const T &v)
{
    key.append(reinterpret_cast<const char *>(&v), sizeof(T));
}

// Tables of a sharded Gram job, kept in the context between its phases (fuf_gram_plan .. fuf_gram_tiles)
struct GramJob {
    int32_t n_rows = 0, n_direct = 0, n_direct_pad = 0, n_parts = 0, n_groups = 0;
    int32_t *d_map = nullptr;
    float *d_scl = nullptr;
    void *d_groups = nullptr, *d_parts = nullptr;
    int32_t *d_dn = nullptr;
};

}  // namespace fuf

struct fuf_ctx {
    int device = 0;
    int sm_count = 0;          // SMs the persistent grids are sized for (fuf_set_sm_limit)
    int sm_count_default = 0;  // ... at context creation: the device's count, or FUF_SM_LIMIT
    int32_t n = 0;
    int32_t n_leaves = 0, depth = 0, n_zero_len = 0;
    bool contiguous = true;  // every subtree is a contiguous postorder range (true DFS postorder)

    // host copies
    std::vector<int32_t> parent;
    std::vector<int32_t> start;  // first postorder index of the subtree of k
    std::vector<double> w[2];    // [mode][k]: L1 -> L (0 -> eps), L2 -> sqrt(L); root -> 1

    // device tree data
    int32_t *d_parent = nullptr;
    double *d_w[2] = {nullptr, nullptr};
    int32_t *d_parent_local = nullptr;  // parent index relative to the node's chunk base, -1 if outside
    int32_t *d_cp_after = nullptr;      // checkpoint slot to store the chunk prefix after node k, -1 if none
    uint8_t *d_is_span = nullptr;       // 1 if node k is a spanning node (written by the fix-up kernel)
    uint8_t *d_has_child = nullptr;     // 1 for internal nodes (the centre is zero on leaves)
    fuf::SpanNode *d_span = nullptr;
    fuf::PushNode *d_push_node[2] = {nullptr, nullptr};  // [mode][n_chunks * 32] (TMA push-up kernel), null if unusable
    int32_t n_chunks = 0, n_span = 0, n_cp = 0;

    // scratch
    fuf::DeviceBuffer ws_push;     // T, CT, CP arrays of the push-up
    fuf::DeviceBuffer ws_pair;     // tile list + split-K partials of the pairwise kernel
    fuf::DeviceBuffer ws_x;        // X staging blocks (fuf_all_pairs_host)
    fuf::DeviceBuffer ws_p;        // PT / P64T (fuf_all_pairs_host)
    fuf::DeviceBuffer ws_d;        // D on device when the host buffer cannot be written in place
    fuf::DeviceBuffer ws_center;   // column mean scratch of fuf_push_up_center
    fuf::DeviceBuffer ws_rec;      // per-job node records of the TMA push-up kernel
    fuf::DeviceBuffer ws_rows;     // tile-row list of fuf_refine
    fuf::DeviceBuffer ws_panel;    // transposed tile panels of fuf_store_tile_rows_host
    fuf::DeviceBuffer ws_gram;     // int8 digit planes + tables of fuf_pairwise_gram
    fuf::DeviceBuffer ws_gram_small;  // column maxima
    fuf::DeviceBuffer ws_gram_tab;    // plan tables of a sharded Gram job
    fuf::DeviceBuffer ws_seg;         // completion counters of the Gram launch of fuf_all_pairs_host
    fuf::GramJob gram_job;
    fuf::DeviceBuffer ws_p64;      // centre + error statistics (fuf_all_pairs_host)
    fuf::DeviceBuffer ws_p64t;     // float64 pushed profiles, only when a pair has to be refined (fuf_all_pairs_host)
    unsigned long long *d_refined = nullptr;  // pairs recomputed by the last fuf_refine
    int64_t last_refined = 0;
    fuf::TableCache tables;        // cached device copies of host-planned tables
    cudaStream_t s_copy = nullptr, s_comp = nullptr, s_d2h = nullptr;
    cudaEvent_t ev[8] = {};
    cudaEvent_t ev_gram[2] = {};       // around the last gram_tile_kernel launch
    double gram_flops = 0.0;           // MMA flops that launch executed
    int gram_groups = 0;               // exponent buckets of the last Gram call
    int gram_bypassed = 0;             // node columns the last Gram call computed in float64
    std::vector<cudaEvent_t> ev_band;  // one per sample band of fuf_all_pairs_host
    float t_h2d_push = 0.f, t_pair = 0.f, t_total = 0.f;

    int64_t launches = 0;
};

namespace fuf {

// pushup.cu
int push_up_f32(fuf_ctx *ctx, const void *X, int x_dtype, int64_t N, int64_t ldx, int mode, const double *center,
                float *PT, int64_t ldp, double *P64T, int64_t ldp64, double *err_stat, int64_t col0,
                cudaStream_t stream);
int push_up_center(fuf_ctx *ctx, const void *X, int x_dtype, int64_t m, int64_t ldx, int mode, double *center_out,
                   cudaStream_t stream);
int push_up_f64_exact(fuf_ctx *ctx, const double *X, int64_t N, int64_t ldx, int mode, double *P64T, int64_t ldp,
                      int64_t col0, cudaStream_t stream);
// pairwise.cu
// Explicit tile list of fuf_pairwise_tiles (multi-GPU: operand shards arrive while the kernel runs; results go straight
// into the full matrix, wherever it lives; refinement flags are evaluated in the tile epilogues).
struct PairTileList {
    const int32_t *tiles_h = nullptr;      // (ti, tj, dependency) triples, ti <= tj; dependency = shard flag index or -1
    int32_t n_tiles = 0;
    const uint32_t *dep_flags = nullptr;   // device
    uint32_t epoch = 0;
    const double *stat = nullptr;          // device: per-sample rounding statistics (all samples)
    double kappa = 0.0;
    long long *flag_list = nullptr;        // device: [0] = count, then (i, j) pairs
    long long list_cap = 0;
    const int32_t *seg_h = nullptr;        // segment of every tile (or -1), n_tiles entries
    uint32_t *seg_done = nullptr;          // device: completion counters, one per segment (+1 per finished tile)
};
int pairwise_f32(fuf_ctx *ctx, const float *PT, int64_t N, int64_t ldp, int64_t shard, int64_t shard_stride, int mode,
                 const int32_t *tile_rows_h, int32_t n_tile_rows, double *D, int64_t ldd, int flags,
                 cudaStream_t stream, int32_t col_tile_begin = 0, int32_t col_tile_end = 0, int32_t n_rows_override = 0,
                 int32_t flush_stages = 0, const PairTileList *list = nullptr);
int pairwise_f64(fuf_ctx *ctx, const double *P64T, int64_t N, int64_t ldp, int64_t shard, int64_t shard_stride, int mode,
                 const int32_t *tile_rows_h, int32_t n_tile_rows, double *D, int64_t ldd, cudaStream_t stream);
int assemble(fuf_ctx *ctx, const double *blocks, const int32_t *tile_rows_h, int32_t n_blocks, int64_t N, int64_t ldb,
             double *D, int64_t ldd, cudaStream_t stream);
int pack_tile_rows(fuf_ctx *ctx, const double *blocks, int64_t ldb, const int32_t *tile_rows_h, const int64_t *offsets_h,
                   int32_t n_blocks, int64_t N, double *packed, cudaStream_t stream);
int assemble_packed(fuf_ctx *ctx, const double *packed, const int32_t *tile_rows_h, const int64_t *offsets_h,
                    int32_t n_blocks, int64_t N, double *D, int64_t ldd, cudaStream_t stream);
int store_tile_rows_host(fuf_ctx *ctx, double *blocks, const int32_t *tile_rows_h, int32_t n_blocks, int64_t N,
                         int64_t ldb, double *D_h, int64_t ldd, cudaStream_t stream);
// tree.cu
int stream_wait_u32(const uint32_t *dptr, uint32_t value, cudaStream_t stream);
// cudaMemcpy2DAsync (cudaMemcpyDefault), cut at the boundaries of a host range that fuf_host_register page-locked in pieces
int copy2d_async(void *dst, size_t dst_pitch, const void *src, size_t src_pitch, size_t width, size_t rows, cudaStream_t stream);
// gram.cu
// Completion segments of a full-matrix Gram launch: in — device counters and the number of segments wanted; out — the
// tile-row boundary of every segment and the counter value that says it is complete (empty: no segments were made).
struct GramSegments {
    uint32_t *seg_done = nullptr;
    int32_t n_wanted = 0;
    std::vector<int32_t> row_end;
    std::vector<uint32_t> expected;
};
int pairwise_gram(fuf_ctx *ctx, const float *PT, int64_t N, int64_t ldp, int64_t shard, int64_t shard_stride,
                  const double *err_in, double *q, double *err_out, const int32_t *tile_rows_h, int32_t n_tile_rows,
                  double *D, int64_t ldd, cudaStream_t stream, GramSegments *segs = nullptr);
// refine.cu
int refine(fuf_ctx *ctx, const double *P64T, int64_t N, int64_t ldp64, int64_t shard, int64_t shard_stride, int mode,
           const double *err_stat, double kappa, const int32_t *tile_rows_h, int32_t n_tile_rows, double *D,
           int64_t ldd, cudaStream_t stream, int64_t j_begin = 0, int64_t j_end = 0, bool reset_counter = true,
           int criterion = 0, const double *lin_stat = nullptr, double lin_kappa = 0.0, int64_t max_pairs = 0);
// diffab.cu
int diffab_block(fuf_ctx *ctx, const void *PT, int p_dtype, int64_t N, int64_t ldp, int mode, int64_t pair_begin,
                 int64_t pair_end, void *out, int out_dtype, cudaStream_t stream);

}  // namespace fuf
