// tree.cu — context lifetime and tree preprocessing (host side) for libfununifrac.so.
//
// Replaces the dict-based tree view of the reference (Tint / lint of tree_to_EMDU_input,
// src/factory/make_emd_input.py:10-54) with dense postorder arrays on the device, plus the
// metadata the chunked push-up kernel needs (see pushup.cu).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <map>

#include <cstdlib>

#include <mutex>

#include "common.cuh"

namespace fuf {

static thread_local std::string g_last_error;

void set_error(const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

int DeviceBuffer::reserve(size_t want)
{
    if (want <= bytes) return FUF_OK;
    if (ptr) {
        cudaFree(ptr);
        ptr = nullptr;
        bytes = 0;
    }
    size_t rounded = (want + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
    cudaError_t err = cudaMalloc(&ptr, rounded);
    if (err != cudaSuccess) {
        ptr = nullptr;
        set_error("cudaMalloc(%zu bytes) failed: %s", rounded, cudaGetErrorString(err));
        cudaGetLastError();
        return FUF_ENOMEM;
    }
    bytes = rounded;
    return FUF_OK;
}

void DeviceBuffer::release()
{
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
}

int PinnedBuffer::reserve(size_t want)
{
    if (want <= bytes) return FUF_OK;
    if (ptr) {
        cudaFreeHost(ptr);
        ptr = nullptr;
        bytes = 0;
    }
    cudaError_t err = cudaHostAlloc(&ptr, want, cudaHostAllocDefault);
    if (err != cudaSuccess) {
        ptr = nullptr;
        set_error("cudaHostAlloc(%zu bytes) failed: %s", want, cudaGetErrorString(err));
        cudaGetLastError();
        return FUF_ENOMEM;
    }
    bytes = want;
    return FUF_OK;
}

void PinnedBuffer::release()
{
    if (ptr) cudaFreeHost(ptr);
    ptr = nullptr;
    bytes = 0;
}

void *TableCache::find(const std::string &key) const
{
    auto it = map.find(key);
    return it == map.end() ? nullptr : it->second;
}

int TableCache::put(const std::string &key, const void *host, size_t nbytes, void **d_out)
{
    if (map.size() >= 512 || bytes + nbytes > ((size_t)256 << 20)) release();  // cudaFree waits for the device
    void *d = nullptr;
    FUF_CUDA(cudaMalloc(&d, nbytes ? nbytes : 1));
    if (nbytes) {
        cudaError_t err = cudaMemcpy(d, host, nbytes, cudaMemcpyHostToDevice);
        if (err != cudaSuccess) {
            cudaFree(d);
            set_error("table upload (%zu bytes) failed: %s", nbytes, cudaGetErrorString(err));
            return FUF_ECUDA;
        }
    }
    map.emplace(key, d);
    bytes += nbytes;
    *d_out = d;
    return FUF_OK;
}

int TableCache::get_by_content(const char *tag, const void *host, size_t nbytes, void **d_out)
{
    std::string key(tag);
    key.push_back('\0');
    key.append(static_cast<const char *>(host), nbytes);
    if (void *d = find(key)) {
        *d_out = d;
        return FUF_OK;
    }
    return put(key, host, nbytes, d_out);
}

void TableCache::release()
{
    for (auto &kv : map) cudaFree(kv.second);
    map.clear();
    bytes = 0;
}

EncodeTiledFn tensor_map_encoder()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

template <typename T>
static int upload(T **dst, const std::vector<T> &src)
{
    size_t bytes = src.size() * sizeof(T);
    if (bytes == 0) bytes = sizeof(T);
    FUF_CUDA(cudaMalloc((void **)dst, bytes));
    if (!src.empty()) FUF_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return FUF_OK;
}

// Host-only planning of the chunked push-up (no device work): subtree ranges, which nodes span a chunk
// boundary, and where chunk-local prefix sums must be checkpointed.  Shared by fuf_tree_create and the
// test hook fuf_debug_tree_plan.
struct TreePlan {
    std::vector<int32_t> start, parent_local, cp_after;
    std::vector<uint8_t> is_span, has_child;
    std::vector<SpanNode> span;
    int32_t n_leaves = 0, depth = 0, n_cp = 0;
    bool contiguous = true;
};

static void build_tree_plan(const int32_t *parent_h, int32_t n, TreePlan &plan)
{
    const int32_t KT = kPushChunk;
    std::vector<int32_t> size(n, 1), depth(n, 0);
    std::vector<uint8_t> has_child(n, 0);
    for (int32_t i = 0; i < n - 1; i++) {
        size[parent_h[i]] += size[i];
        has_child[parent_h[i]] = 1;
    }
    for (int32_t i = n - 2; i >= 0; i--) depth[i] = depth[parent_h[i]] + 1;
    plan.start.resize(n);
    plan.has_child = has_child;
    for (int32_t i = 0; i < n; i++) {
        plan.start[i] = i - size[i] + 1;
        if (!has_child[i]) plan.n_leaves++;
        if (depth[i] > plan.depth) plan.depth = depth[i];
    }
    // every subtree is the contiguous range [start, k] iff each child's range nests in its parent's
    plan.contiguous = true;
    for (int32_t i = 0; i < n - 1; i++)
        if (plan.start[parent_h[i]] > plan.start[i] || plan.start[i] < 0) {
            plan.contiguous = false;
            break;
        }
    plan.parent_local.assign(n, -1);
    plan.cp_after.assign(n, -1);
    plan.is_span.assign(n, 0);
    plan.span.clear();
    plan.n_cp = 0;
    if (!plan.contiguous) return;
    std::map<int32_t, int32_t> slot_of_pos;  // node position -> checkpoint slot
    auto slot_for = [&](int32_t pos) {
        auto it = slot_of_pos.find(pos);
        if (it != slot_of_pos.end()) return it->second;
        int32_t s = (int32_t)slot_of_pos.size();
        slot_of_pos[pos] = s;
        plan.cp_after[pos] = s;
        return s;
    };
    for (int32_t k = 0; k < n; k++) {
        if (k < n - 1 && parent_h[k] / KT == k / KT) plan.parent_local[k] = parent_h[k] % KT;
        const int32_t st = plan.start[k];
        if (st / KT != k / KT) {
            SpanNode s;
            s.k = k;
            s.chunk_first = st / KT;
            s.chunk_last = k / KT;
            s.slot_start = (st % KT == 0) ? -1 : slot_for(st - 1);
            s.slot_k = slot_for(k);
            plan.span.push_back(s);
            plan.is_span[k] = 1;
        }
    }
    plan.n_cp = (int32_t)slot_of_pos.size();
}

// Plan words of the TMA push-up kernel (see PushNode in common.cuh), padded to whole chunks with "skip".  Returns false
// when a chunk would need more parked prefixes than the kernel has slots for (kPushParkSlots).
static bool build_push_meta(const TreePlan &plan, int32_t n, std::vector<int32_t> &meta)
{
    const int32_t KT = kPushChunk, padded = (n + KT - 1) / KT * KT;
    std::vector<int32_t> park_slot(n, -1);   // slot (within the chunk) that keeps the prefix after node k
    meta.assign(padded, 2);
    for (int32_t c0 = 0; c0 < n; c0 += KT) {
        int32_t used = 0;
        const int32_t c1 = std::min(n, c0 + KT);
        for (int32_t k = c0; k < c1; k++)   // an internal node whose subtree starts inside its chunk, after node start - 1
            if (!plan.is_span[k] && plan.start[k] != k && plan.start[k] > c0 && park_slot[plan.start[k] - 1] < 0) {
                if (used == kPushParkSlots) return false;
                park_slot[plan.start[k] - 1] = used++;
            }
    }
    for (int32_t k = 0; k < n; k++) {
        int32_t m = 0;
        if (plan.is_span[k])
            m |= 2;
        else if (plan.start[k] == k)
            m |= 1;
        else if (plan.start[k] > (k / KT) * KT)
            m |= (park_slot[plan.start[k] - 1] + 1) << 8;      // S = prefix - parked prefix; 0: S = prefix
        if (park_slot[k] >= 0) m |= (park_slot[k] + 1) << 4;
        if (plan.cp_after[k] >= 0) m |= 8;
        m |= (plan.cp_after[k] + 1) << 16;
        meta[k] = m;
    }
    return true;
}

}  // namespace fuf

using namespace fuf;

// Test hook (host only): the per-node plan words of the TMA push-up kernel, n_chunks * 32 entries; *usable = 0 when the
// tree cannot use that kernel (not a DFS postorder, or more than 65,534 checkpoint slots).
extern "C" int fuf_debug_push_meta(const int32_t *parent_h, int32_t n, int32_t *meta_out, int32_t *usable)
{
    FUF_REQUIRE(parent_h && meta_out && usable, "fuf_debug_push_meta: null argument");
    FUF_REQUIRE(n >= 1 && parent_h[n - 1] == -1, "fuf_debug_push_meta: root must be last");
    for (int32_t i = 0; i < n - 1; i++)
        FUF_REQUIRE(parent_h[i] > i && parent_h[i] < n, "fuf_debug_push_meta: parent[%d] = %d is not a postorder", i,
                    parent_h[i]);
    TreePlan plan;
    build_tree_plan(parent_h, n, plan);
    std::vector<int32_t> meta;
    *usable = (plan.contiguous && plan.n_cp < 65535 && build_push_meta(plan, n, meta)) ? 1 : 0;
    if (!*usable) return FUF_OK;
    for (size_t k = 0; k < meta.size(); k++) meta_out[k] = meta[k];
    return FUF_OK;
}

// Test hook (host only, no device needed): the push-up plan for a postorder parent array.
// span_out receives 5 ints per spanning node (k, chunk_first, chunk_last, slot_start, slot_k); capacity n nodes.
extern "C" int fuf_debug_tree_plan(const int32_t *parent_h, int32_t n, int32_t *parent_local, int32_t *cp_after,
                                   uint8_t *is_span, int32_t *span_out, int32_t *n_span, int32_t *n_cp,
                                   int32_t *contiguous, int32_t *chunk)
{
    FUF_REQUIRE(parent_h && parent_local && cp_after && is_span && span_out && n_span && n_cp && contiguous && chunk,
                "fuf_debug_tree_plan: null argument");
    FUF_REQUIRE(n >= 1 && parent_h[n - 1] == -1, "fuf_debug_tree_plan: root must be last");
    for (int32_t i = 0; i < n - 1; i++)
        FUF_REQUIRE(parent_h[i] > i && parent_h[i] < n, "fuf_debug_tree_plan: parent[%d] = %d is not a postorder", i,
                    parent_h[i]);
    TreePlan plan;
    build_tree_plan(parent_h, n, plan);
    for (int32_t i = 0; i < n; i++) {
        parent_local[i] = plan.parent_local[i];
        cp_after[i] = plan.cp_after[i];
        is_span[i] = plan.is_span[i];
    }
    for (size_t q = 0; q < plan.span.size(); q++) {
        span_out[5 * q + 0] = plan.span[q].k;
        span_out[5 * q + 1] = plan.span[q].chunk_first;
        span_out[5 * q + 2] = plan.span[q].chunk_last;
        span_out[5 * q + 3] = plan.span[q].slot_start;
        span_out[5 * q + 4] = plan.span[q].slot_k;
    }
    *n_span = (int32_t)plan.span.size();
    *n_cp = plan.n_cp;
    *contiguous = plan.contiguous ? 1 : 0;
    *chunk = kPushChunk;
    return FUF_OK;
}

extern "C" int fuf_version(void) { return FUF_ABI_VERSION; }

extern "C" const char *fuf_last_error(void) { return g_last_error.c_str(); }

extern "C" int64_t fuf_padded_samples(int64_t N)
{
    if (N <= 0) return FUF_TILE;
    return (N + FUF_TILE - 1) / FUF_TILE * FUF_TILE;
}

extern "C" int fuf_tree_create(fuf_ctx **out, int device, const int32_t *parent_h, const double *length_h, int32_t n)
{
    FUF_REQUIRE(out != nullptr && parent_h != nullptr && length_h != nullptr, "fuf_tree_create: null argument");
    FUF_REQUIRE(n >= 1, "fuf_tree_create: n must be >= 1 (got %d)", n);
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("fuf_tree_create: no CUDA device available (this library has no CPU fallback)");
        return FUF_ENODEV;
    }
    if (device < 0) FUF_CUDA(cudaGetDevice(&device));
    FUF_REQUIRE(device < ndev, "fuf_tree_create: device %d out of range (%d devices)", device, ndev);
    cudaDeviceProp prop;
    FUF_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("fuf_tree_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                  prop.minor);
        return FUF_ENODEV;
    }
    // ---- validate the postorder contract --------------------------------------------------
    FUF_REQUIRE(parent_h[n - 1] == -1, "fuf_tree_create: parent[n-1] must be -1 (root last), got %d", parent_h[n - 1]);
    for (int32_t i = 0; i < n - 1; i++) {
        FUF_REQUIRE(parent_h[i] > i && parent_h[i] < n,
                    "fuf_tree_create: parent[%d] = %d violates child < parent < n (postorder)", i, parent_h[i]);
        FUF_REQUIRE(std::isfinite(length_h[i]), "fuf_tree_create: edge length of node %d is not finite", i);
    }
    fuf_ctx *ctx = new (std::nothrow) fuf_ctx();
    if (!ctx) {
        set_error("fuf_tree_create: out of host memory");
        return FUF_ENOMEM;
    }
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    // FUF_SM_LIMIT: size the persistent grids for fewer SMs than the device has — for a device shared with other clients
    // (the single-GPU emulation of the multi-GPU job runs several ranks on one device, where the "peer" pulls are
    // same-device copies the driver may run as kernels: they need an SM that no persistent CTA is spinning on)
    if (const char *lim = getenv("FUF_SM_LIMIT")) {
        const int v = atoi(lim);
        if (v >= 2 && v < ctx->sm_count) ctx->sm_count = v;
    }
    ctx->sm_count_default = ctx->sm_count;
    ctx->n = n;
    ctx->parent.assign(parent_h, parent_h + n);

    // weights: zero -> DBL_EPSILON (emd_unifrac.py:24-25,35-36), L2 takes the square root (:28); the root is unscaled
    ctx->w[0].resize(n);
    ctx->w[1].resize(n);
    for (int32_t i = 0; i < n - 1; i++) {
        double l = length_h[i];
        if (l == 0.0) {
            l = DBL_EPSILON;
            ctx->n_zero_len++;
        }
        ctx->w[0][i] = l;
        ctx->w[1][i] = std::sqrt(l);
    }
    ctx->w[0][n - 1] = 1.0;
    ctx->w[1][n - 1] = 1.0;

    TreePlan plan;
    build_tree_plan(parent_h, n, plan);
    ctx->start = plan.start;
    ctx->n_leaves = plan.n_leaves;
    ctx->depth = plan.depth;
    ctx->contiguous = plan.contiguous;
    ctx->n_chunks = (n + kPushChunk - 1) / kPushChunk;
    ctx->n_cp = plan.n_cp;
    ctx->n_span = (int32_t)plan.span.size();
    std::vector<int32_t> &parent_local = plan.parent_local, &cp_after = plan.cp_after;
    std::vector<uint8_t> &is_span = plan.is_span;
    std::vector<SpanNode> &span = plan.span;

    // per-node records of the TMA push-up kernel (only for DFS postorders whose checkpoint slots fit 16 bits)
    std::vector<PushNode> push_node[2];
    std::vector<int32_t> meta;
    if (plan.contiguous && plan.n_cp < 65535 && build_push_meta(plan, n, meta)) {
        for (int mode = 0; mode < 2; mode++) {
            push_node[mode].assign(meta.size(), PushNode{0.0, 2, 0});
            for (size_t k = 0; k < meta.size(); k++)
                push_node[mode][k] = PushNode{k < (size_t)n ? ctx->w[mode][k] : 0.0, meta[k], 0};
        }
    }

    int rc = FUF_OK;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        set_error("cudaSetDevice(%d) failed: %s", device, cudaGetErrorString(e));
        delete ctx;
        return FUF_ECUDA;
    }
    if ((rc = upload(&ctx->d_parent, ctx->parent)) || (rc = upload(&ctx->d_w[0], ctx->w[0])) ||
        (rc = upload(&ctx->d_w[1], ctx->w[1])) || (rc = upload(&ctx->d_parent_local, parent_local)) ||
        (rc = upload(&ctx->d_cp_after, cp_after)) || (rc = upload(&ctx->d_is_span, is_span)) ||
        (rc = upload(&ctx->d_span, span)) || (rc = upload(&ctx->d_has_child, plan.has_child)) ||
        (!push_node[0].empty() && ((rc = upload(&ctx->d_push_node[0], push_node[0])) ||
                                   (rc = upload(&ctx->d_push_node[1], push_node[1]))))) {
        fuf_tree_destroy(ctx);
        return rc;
    }
    if (cudaStreamCreateWithFlags(&ctx->s_copy, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->s_comp, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking) != cudaSuccess) {
        set_error("fuf_tree_create: cudaStreamCreate failed");
        fuf_tree_destroy(ctx);
        return FUF_ECUDA;
    }
    for (auto &ev : ctx->ev)
        if (cudaEventCreate(&ev) != cudaSuccess) {
            set_error("fuf_tree_create: cudaEventCreate failed");
            fuf_tree_destroy(ctx);
            return FUF_ECUDA;
        }
    *out = ctx;
    return FUF_OK;
}

extern "C" int fuf_tree_destroy(fuf_ctx *ctx)
{
    if (!ctx) return FUF_OK;
    cudaSetDevice(ctx->device);
    cudaFree(ctx->d_parent);
    cudaFree(ctx->d_w[0]);
    cudaFree(ctx->d_w[1]);
    cudaFree(ctx->d_parent_local);
    cudaFree(ctx->d_cp_after);
    cudaFree(ctx->d_is_span);
    cudaFree(ctx->d_span);
    cudaFree(ctx->d_has_child);
    cudaFree(ctx->d_push_node[0]);
    cudaFree(ctx->d_push_node[1]);
    cudaFree(ctx->d_refined);
    ctx->ws_center.release();
    ctx->ws_rec.release();
    ctx->ws_rows.release();
    ctx->ws_panel.release();
    ctx->ws_gram.release();
    ctx->ws_gram_small.release();
    ctx->ws_gram_tab.release();
    ctx->ws_seg.release();
    ctx->ws_p64.release();
    ctx->ws_p64t.release();
    ctx->ws_push.release();
    ctx->ws_pair.release();
    ctx->ws_x.release();
    ctx->ws_p.release();
    ctx->ws_d.release();
    ctx->tables.release();
    if (ctx->s_copy) cudaStreamDestroy(ctx->s_copy);
    if (ctx->s_comp) cudaStreamDestroy(ctx->s_comp);
    if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
    for (auto &ev : ctx->ev)
        if (ev) cudaEventDestroy(ev);
    for (auto &ev : ctx->ev_band) cudaEventDestroy(ev);
    for (auto &ev : ctx->ev_gram)
        if (ev) cudaEventDestroy(ev);
    cudaGetLastError();
    delete ctx;
    return FUF_OK;
}

extern "C" int fuf_tree_get_info(const fuf_ctx *ctx, fuf_tree_info *info)
{
    FUF_REQUIRE(ctx != nullptr && info != nullptr, "fuf_tree_get_info: null argument");
    info->n = ctx->n;
    info->n_leaves = ctx->n_leaves;
    info->depth = ctx->depth;
    info->chunk = kPushChunk;
    info->n_chunks = ctx->n_chunks;
    info->n_spanning = ctx->contiguous ? ctx->n_span : -1;
    info->n_zero_len = ctx->n_zero_len;
    info->device = ctx->device;
    return FUF_OK;
}

// ---- peer memory: one process per GPU on one node shares device buffers through CUDA IPC handles ----------------
extern "C" int fuf_peer_alloc(fuf_ctx *ctx, size_t bytes, void **dptr, unsigned char handle[64])
{
    FUF_REQUIRE(ctx && dptr && handle, "fuf_peer_alloc: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handles are exchanged as 64 bytes");
    FUF_CUDA(cudaSetDevice(ctx->device));
    void *p = nullptr;
    cudaError_t err = cudaMalloc(&p, bytes ? bytes : 1);
    if (err != cudaSuccess) {
        set_error("fuf_peer_alloc: cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(err));
        cudaGetLastError();
        return FUF_ENOMEM;
    }
    cudaIpcMemHandle_t h;
    err = cudaIpcGetMemHandle(&h, p);
    if (err != cudaSuccess) {
        cudaFree(p);
        set_error("fuf_peer_alloc: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(err));
        cudaGetLastError();
        return FUF_ECUDA;
    }
    FUF_CUDA(cudaMemset(p, 0, bytes ? bytes : 1));
    memcpy(handle, &h, 64);
    *dptr = p;
    return FUF_OK;
}

extern "C" int fuf_peer_open(fuf_ctx *ctx, const unsigned char handle[64], void **dptr)
{
    FUF_REQUIRE(ctx && dptr && handle, "fuf_peer_open: null argument");
    FUF_CUDA(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    *dptr = nullptr;
    FUF_CUDA(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return FUF_OK;
}

extern "C" int fuf_peer_close(fuf_ctx *ctx, void *dptr)
{
    FUF_REQUIRE(ctx, "fuf_peer_close: null context");
    if (!dptr) return FUF_OK;
    FUF_CUDA(cudaSetDevice(ctx->device));
    FUF_CUDA(cudaIpcCloseMemHandle(dptr));
    return FUF_OK;
}

extern "C" int fuf_peer_free(fuf_ctx *ctx, void *dptr)
{
    FUF_REQUIRE(ctx, "fuf_peer_free: null context");
    if (!dptr) return FUF_OK;
    FUF_CUDA(cudaSetDevice(ctx->device));
    FUF_CUDA(cudaFree(dptr));
    return FUF_OK;
}

// Page-lock a range of host memory another allocator owns (the result matrix the ranks of a node share).  Some boxes
// refuse one registration of tens of gigabytes (cudaErrorOperatingSystem) but take it in pieces; with unified addressing
// the pieces still form one contiguous device-visible range.  *n_pieces receives how many registrations were made.
namespace {
// Ranges fuf_host_register had to page-lock in pieces.  A copy-engine transfer must lie inside ONE registration, so the
// pitched copies into such a range are cut at the piece boundaries (fuf::copy2d_async).
struct PiecedRange {
    uintptr_t base;
    size_t bytes, step;
};
std::mutex g_pieced_mutex;
std::vector<PiecedRange> g_pieced;

bool pieced_range_of(const void *p, PiecedRange *out)
{
    std::lock_guard<std::mutex> lock(g_pieced_mutex);
    for (const PiecedRange &r : g_pieced)
        if ((uintptr_t)p >= r.base && (uintptr_t)p < r.base + r.bytes) {
            *out = r;
            return true;
        }
    return false;
}
}  // namespace

namespace fuf {
int copy2d_async(void *dst, size_t dst_pitch, const void *src, size_t src_pitch, size_t width, size_t rows, cudaStream_t stream)
{
    if (!width || !rows) return FUF_OK;
    PiecedRange r;
    const bool dst_pieced = pieced_range_of(dst, &r);
    if (!dst_pieced && !pieced_range_of(src, &r)) {
        FUF_CUDA(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width, rows, cudaMemcpyDefault, stream));
        return FUF_OK;
    }
    // p walks the side that lives in the pieced range: runs of whole rows inside one piece go as one pitched copy, a row
    // whose segment straddles a boundary as plain copies of its parts
    const uintptr_t p0 = (uintptr_t)(dst_pieced ? dst : src);
    const size_t pitch = dst_pieced ? dst_pitch : src_pitch;
    size_t i = 0;
    while (i < rows) {
        const uintptr_t p = p0 + i * pitch;
        const uintptr_t piece_end = r.base + ((p - r.base) / r.step + 1) * r.step;
        char *d = (char *)dst + i * dst_pitch;
        const char *s = (const char *)src + i * src_pitch;
        if (p + width <= piece_end) {
            const size_t fit = pitch ? (piece_end - width - p) / pitch + 1 : rows - i;   // rows i .. i + fit - 1 end inside the piece
            const size_t run = std::min(rows - i, fit);
            FUF_CUDA(cudaMemcpy2DAsync(d, dst_pitch, s, src_pitch, width, run, cudaMemcpyDefault, stream));
            i += run;
        } else {
            for (size_t at = 0; at < width;) {   // (a row wider than a piece crosses several boundaries)
                const uintptr_t q = p + at;
                const size_t part = std::min<size_t>(width - at, r.base + ((q - r.base) / r.step + 1) * r.step - q);
                FUF_CUDA(cudaMemcpyAsync(d + at, s + at, part, cudaMemcpyDefault, stream));
                at += part;
            }
            i += 1;
        }
    }
    return FUF_OK;
}
}  // namespace fuf

extern "C" int fuf_host_register(void *ptr, size_t bytes, int32_t *n_pieces)
{
    FUF_REQUIRE(ptr && n_pieces, "fuf_host_register: null argument");
    *n_pieces = 0;
    if (bytes == 0) return FUF_OK;
    // FUF_REGISTER_STEP (bytes, a multiple of the page size): register in pieces of that size right away — a test hook for
    // the piece-boundary handling of the pitched copies
    size_t step = (size_t)1 << 30;
    const char *forced = getenv("FUF_REGISTER_STEP");
    if (forced && atoll(forced) >= 4096) {
        step = (size_t)atoll(forced);
    } else if (cudaHostRegister(ptr, bytes, cudaHostRegisterDefault) == cudaSuccess) {
        *n_pieces = 1;
        return FUF_OK;
    }
    cudaGetLastError();   // not sticky: clear it before trying the pieces
    int32_t done = 0;
    for (size_t off = 0; off < bytes; off += step, done++) {
        cudaError_t err = cudaHostRegister((char *)ptr + off, std::min(step, bytes - off), cudaHostRegisterDefault);
        if (err != cudaSuccess) {
            cudaGetLastError();
            for (int32_t q = 0; q < done; q++) cudaHostUnregister((char *)ptr + (size_t)q * step);
            cudaGetLastError();
            set_error("fuf_host_register: cudaHostRegister failed at byte %zu of %zu: %s", off, bytes, cudaGetErrorString(err));
            return FUF_ECUDA;
        }
    }
    *n_pieces = done;
    {
        std::lock_guard<std::mutex> lock(g_pieced_mutex);
        g_pieced.push_back(PiecedRange{(uintptr_t)ptr, bytes, step});
    }
    return FUF_OK;
}

extern "C" int fuf_host_unregister(void *ptr, size_t bytes, int32_t n_pieces)
{
    if (!ptr || n_pieces <= 0) return FUF_OK;
    size_t step = 0;
    {
        std::lock_guard<std::mutex> lock(g_pieced_mutex);
        for (size_t q = 0; q < g_pieced.size(); q++)
            if (g_pieced[q].base == (uintptr_t)ptr) {
                step = g_pieced[q].step;
                g_pieced.erase(g_pieced.begin() + q);
                break;
            }
    }
    if (step == 0) {
        cudaHostUnregister(ptr);
    } else {
        for (size_t off = 0; off < bytes; off += step) cudaHostUnregister((char *)ptr + off);
    }
    cudaGetLastError();
    return FUF_OK;
}

// Device-side address of page-locked (cudaHostAlloc / cudaHostRegister) host memory: what a kernel stores through when
// the result matrix lives in host memory shared by the ranks of a node.
extern "C" int fuf_host_device_pointer(fuf_ctx *ctx, void *host_ptr, void **dptr)
{
    FUF_REQUIRE(ctx && host_ptr && dptr, "fuf_host_device_pointer: null argument");
    FUF_CUDA(cudaSetDevice(ctx->device));
    FUF_CUDA(cudaHostGetDevicePointer(dptr, host_ptr, 0));
    return FUF_OK;
}

// Asynchronous copy between any two device-accessible buffers (own memory, a peer's IPC mapping): the copy engines
// pull a peer's shard over NVLink while the SMs compute.
extern "C" int fuf_copy_async(fuf_ctx *ctx, void *dst, const void *src, size_t bytes, void *stream)
{
    FUF_REQUIRE(ctx && dst && src, "fuf_copy_async: null argument");
    FUF_CUDA(cudaSetDevice(ctx->device));
    if (bytes) FUF_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return FUF_OK;
}

// Stream-ordered 32-bit write without an SM (cuStreamWriteValue32): the "shard has landed" flag a running kernel polls.
extern "C" int fuf_stream_write_u32(fuf_ctx *ctx, uint32_t *dptr, uint32_t value, void *stream)
{
    FUF_REQUIRE(ctx && dptr, "fuf_stream_write_u32: null argument");
    FUF_REQUIRE(((uintptr_t)dptr & 3) == 0, "fuf_stream_write_u32: unaligned address");
    FUF_CUDA(cudaSetDevice(ctx->device));
    typedef CUresult (*WriteFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
    static WriteFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            set_error("fuf_stream_write_u32: cuStreamWriteValue32 is not available from this driver");
            return FUF_ECUDA;
        }
        fn = (WriteFn)ptr;
    }
    CUresult r = fn((CUstream)stream, (CUdeviceptr)(uintptr_t)dptr, value, 0);
    if (r != CUDA_SUCCESS) {
        set_error("fuf_stream_write_u32: cuStreamWriteValue32 failed with CUresult %d", (int)r);
        return FUF_ECUDA;
    }
    return FUF_OK;
}

// Stream-ordered wait until a 32-bit device word has reached a value (cuStreamWaitValue32, >=): the stream's later work
// — e.g. the copy of a finished segment of D — starts once a running kernel has counted that far.
namespace fuf {
int stream_wait_u32(const uint32_t *dptr, uint32_t value, cudaStream_t stream)
{
    typedef CUresult (*WaitFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
    static WaitFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            set_error("fuf_stream_wait_u32: cuStreamWaitValue32 is not available from this driver");
            return FUF_ECUDA;
        }
        fn = (WaitFn)ptr;
    }
    CUresult r = fn((CUstream)stream, (CUdeviceptr)(uintptr_t)dptr, value, CU_STREAM_WAIT_VALUE_GEQ);
    if (r != CUDA_SUCCESS) {
        set_error("fuf_stream_wait_u32: cuStreamWaitValue32 failed with CUresult %d", (int)r);
        return FUF_ECUDA;
    }
    return FUF_OK;
}
}  // namespace fuf

extern "C" int fuf_stream_wait_u32(fuf_ctx *ctx, const uint32_t *dptr, uint32_t value, void *stream)
{
    FUF_REQUIRE(ctx && dptr, "fuf_stream_wait_u32: null argument");
    FUF_REQUIRE(((uintptr_t)dptr & 3) == 0, "fuf_stream_wait_u32: unaligned address");
    FUF_CUDA(cudaSetDevice(ctx->device));
    return stream_wait_u32(dptr, value, (cudaStream_t)stream);
}

// Pitched copy (device, peer or page-locked host memory on either side) by a copy engine: `rows` rows of `width` bytes.
extern "C" int fuf_copy2d_async(fuf_ctx *ctx, void *dst, size_t dst_pitch, const void *src, size_t src_pitch, size_t width,
                                size_t rows, void *stream)
{
    FUF_REQUIRE(ctx && dst && src, "fuf_copy2d_async: null argument");
    FUF_REQUIRE(dst_pitch >= width && src_pitch >= width, "fuf_copy2d_async: pitch smaller than the row width");
    FUF_CUDA(cudaSetDevice(ctx->device));
    return copy2d_async(dst, dst_pitch, src, src_pitch, width, rows, (cudaStream_t)stream);
}

// Size the persistent grids of this context for `sms` SMs (0: the context's default; < 0: that many fewer than the default).  A caller that makes a running tile
// kernel wait for a strided peer copy must leave SMs free: the driver may run such a copy as an SM kernel, and with a
// persistent CTA spinning on every SM it would never start.
extern "C" int fuf_set_sm_limit(fuf_ctx *ctx, int32_t sms)
{
    FUF_REQUIRE(ctx != nullptr, "fuf_set_sm_limit: null context");
    // never above what the context started with (the device's count, or FUF_SM_LIMIT for a device shared with others)
    const int want = sms < 0 ? ctx->sm_count_default + sms : sms;
    ctx->sm_count = (sms != 0 && want >= 2 && want < ctx->sm_count_default) ? want : ctx->sm_count_default;
    return FUF_OK;
}

extern "C" int64_t fuf_launch_count(const fuf_ctx *ctx) { return ctx ? ctx->launches : 0; }
