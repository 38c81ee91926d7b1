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
//                             its total mass T[chunk] and chunk-local prefix sums CP at checkpoint positions.
//   K1b pushup_span_kernel    spanning node k: S = tail of its first chunk + totals of the chunks in between
//                             + head of its last chunk (same-sign terms only); writes w_k * S.
//   K1c pushup_stat_kernel    sums the per-chunk (and per-span-block) rounding-error statistics of the fp32 outputs per
//                             sample, in fixed order.
//
// Production outputs per entry (all optional):
//   PT    float, CENTRED:  (float)(w_k * S_k - c_k).  Distances are invariant under subtracting one fixed vector c
//         from every profile; with c = the pushed mean of a few samples on the internal nodes (fuf_push_up_center)
//         the large, nearly sample-independent upper-level entries shrink to their sample-to-sample variation, so
//         rounding them to fp32 costs 2^-24 of that variation instead of 2^-24 of their magnitude.
//   P64T  double, uncentred w_k * S_k: the operands of the float64 refinement pass (refine.cu).
//   err   per sample: sum_k |rounding error of PT| (L1) or sum_k (rounding error)^2 (L2), the input to the
//         refinement criterion.
//
// Validation path (fuf_push_up_f64): transpose to node-major double, then one thread per sample replays the
// reference loop verbatim -> bit-identical to the reference's float64 result.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace fuf {

constexpr int KT = kPushChunk;  // nodes per chunk
constexpr int TI = 128;         // samples per CTA
constexpr int PT_SAMPLES = 256;                 // samples per tile = consumer threads
constexpr int PT_GROUP = 8;                     // chunks per work item

template <typename TX, int MODE, bool HAS_PT, bool HAS_P64>
__global__ void __launch_bounds__(TI)
pushup_chunk_kernel(const TX *__restrict__ X, int64_t ldx, int64_t N, int32_t n, const double *__restrict__ w,
                    const int32_t *__restrict__ parent_local, const int32_t *__restrict__ cp_after,
                    const uint8_t *__restrict__ is_span, const double *__restrict__ center, float *__restrict__ PT,
                    int64_t ldp, double *__restrict__ P64T, int64_t ldp64, int64_t col0, double *__restrict__ T,
                    double *__restrict__ CP, double *__restrict__ R, int64_t ldt)
{
    // dynamic shared memory: one array S[KT][TI+1] of doubles (33 KB -> 6 CTAs per SM).  It first receives the X tile,
    // transposed (coalesced global reads, conflict-free padded stores); each thread then lifts ITS sample's KT values
    // into registers and reuses its own column of S as the accumulators of the replayed loop.
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double(*S)[TI + 1] = reinterpret_cast<double(*)[TI + 1]>(smem_raw);
    __shared__ double s_w[KT], s_c[KT];
    __shared__ int32_t s_meta[KT];  // (cp + 1) << 8 | span << 7 | (parent_local + 1); span = 1 beyond the tree

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int32_t k0 = blockIdx.x * KT;
    const int64_t i0 = (int64_t)blockIdx.y * TI;
    const int kt = min(KT, n - k0);

    if (tid < KT) {
        int32_t k = k0 + tid;
        bool ok = tid < kt;
        s_w[tid] = ok ? w[k] : 0.0;
        s_c[tid] = (ok && center) ? center[k] : 0.0;
        s_meta[tid] = ok ? (((cp_after[k] + 1) << 8) | ((int32_t)is_span[k] << 7) | (parent_local[k] + 1)) : (1 << 7);
    }
    // coalesced load of the X tile: one warp request = one sample row segment of KT entries.  All TI/4 requests
    // of a warp are issued before the first value is used (32 x 256 B in flight per warp): the kernel streams
    // X once and is bound by how many bytes it keeps in flight.
    {
        TX v[TI / 4];
#pragma unroll
        for (int it = 0; it < TI / 4; it++) {
            const int64_t s = i0 + it * 4 + warp;
            v[it] = (s < N && lane < kt) ? __ldcs(X + s * ldx + k0 + lane) : TX(0);
        }
#pragma unroll
        for (int it = 0; it < TI / 4; it++) S[lane][it * 4 + warp] = (double)v[it];
    }
    __syncthreads();

    const int64_t s = i0 + tid;
    if (s >= N) return;
    double x[KT];
#pragma unroll
    for (int kk = 0; kk < KT; kk++) {
        x[kk] = S[kk][tid];
        S[kk][tid] = 0.0;  // from here on: sum of the already processed children of node kk
    }
    double r = 0.0;    // running sum of X over the chunk (chunk-local prefix)
    double err = 0.0;  // rounding-error statistic of the float outputs of this chunk
    float *pt = HAS_PT ? PT + (int64_t)k0 * ldp + col0 + s : nullptr;          // walks down the node rows
    double *p64 = HAS_P64 ? P64T + (int64_t)k0 * ldp64 + col0 + s : nullptr;
    double *cps = CP + s;
    double *mine = &S[0][tid];
#pragma unroll
    for (int kk = 0; kk < KT; kk++) {
        const int32_t meta = s_meta[kk];
        const double Sk = x[kk] + mine[kk * (TI + 1)];
        r += x[kk];
        const int pl = (meta & 0x7f) - 1;
        if (pl >= 0) mine[pl * (TI + 1)] += Sk;
        if (!(meta & 0x80)) {
            const double v = s_w[kk] * Sk;
            if (HAS_P64) *p64 = v;
            if (HAS_PT) {
                const double c = v - s_c[kk];
                const float f = (float)c;
                *pt = f;
                const double e = c - (double)f;
                err += (MODE == FUF_MODE_L2) ? e * e : fabs(e);
            }
        }
        const int cp = (meta >> 8) - 1;
        if (cp >= 0) cps[(int64_t)cp * ldt] = r;
        if (HAS_PT) pt += ldp;
        if (HAS_P64) p64 += ldp64;
    }
    if (T) T[(int64_t)blockIdx.x * ldt + s] = r;
    if (R) R[(int64_t)blockIdx.x * ldt + s] = err;
}

// ---------------------------------------------------------------------------------------------------------------
// K1t pushup_tma_kernel: the production push-up when only the centred fp32 profiles (+ statistics) are wanted.
//   * persistent, warp-specialised, two CTAs per SM: warp 8 is the TMA producer, warps 0-7 own 256 samples (one per
//     thread);
//   * work item = (block of 256 samples, group of 8 consecutive chunks).  One pipeline stage = one TMA box of
//     256 samples x 128 bytes of X (16 float64 nodes = half a chunk, or 32 float32 nodes = a chunk; 32 KB,
//     cp.async.bulk.tensor.2d, SWIZZLE_128B) in a 2-stage ring guarded by full/empty mbarriers; rows and nodes beyond
//     the matrix are zero-filled by the TMA unit.  The stage's node records (weight, centre, plan; 32 bytes per node,
//     built per job from the static plan and the centre) arrive with it through one bulk copy;
//   * a thread lifts its sample's values with 16-byte shared loads (the swizzle makes a warp's 32 row reads hit all
//     banks: 4 wavefronts per 512 bytes, the minimum), keeps a running prefix r over the chunk in a register and
//     gets every subtree sum inside the chunk as S = x (leaf) or S = r - r[start - 1] (internal; the few prefixes
//     that are needed again are parked in a [slot][sample] array, at most 12 per chunk): no per-node
//     read-modify-write of shared memory and no dependency chain through it.  The difference of two prefixes of at
//     most 32 same-sign terms is accurate to ~1e-16 of the chunk's mass — far below the fp32 rounding of the output;
//     the float64 profiles the refinement pass needs are produced on demand by the chunk kernel above, in the
//     reference's operation order;
//   * the per-node code is branch-free (selects, predicated accesses; the parked prefixes go through inline PTX that
//     the compiler knows not to alias the record loads), so the nodes of a stage overlap: with four warps per
//     scheduler it is this instruction-level parallelism that hides the conversion and shared-memory latencies;
//   * stores: PT rows, 128 bytes per warp and node; chunk totals T, group totals GT and checkpoints CP for the span
//     kernel; one rounding-error partial R per (group of 8 chunks, sample) — a fixed grouping, so the statistic does
//     not depend on the launch geometry.
// ---------------------------------------------------------------------------------------------------------------
constexpr int PT_STAGES = 2;
constexpr uint32_t PT_STAGE_BYTES = PT_SAMPLES * 128;                     // 32 KB: one 128-byte row per sample
constexpr uint32_t PT_REC_BYTES = KT * 32;                                // up to 1 KB of node records per stage
constexpr uint32_t PT_PARK_BYTES = kPushParkSlots * PT_SAMPLES * sizeof(double);   // 24 KB: [slot][sample]
constexpr uint32_t pt_smem_bytes(int spt)      // spt = samples per thread (tile = spt x 256 samples)
{
    return PT_STAGES * (spt * PT_STAGE_BYTES + PT_REC_BYTES) + spt * PT_PARK_BYTES + 1024 /* alignment */ + 64 /* barriers */;
}

// Per-job record of one node: the static plan (PushNode) joined with the job's centre value, 32 bytes.
struct PushRec {
    double w, c;               // weight, centre (both 0 for nodes this kernel does not write: the value and its error vanish)
    int32_t keep_off;          // byte offset of the park slot that keeps the prefix after this node; < 0: none
    int32_t take_off;          // byte offset of the parked prefix to subtract; -1: none (S = prefix); -2: leaf-like (S = x)
    int32_t store;             // -1 (all ones): this kernel writes the node's output; 0: somebody else does / padding
    int32_t cp_slot;           // checkpoint slot of the span kernel; < 0: none
};

__global__ void __launch_bounds__(256)
push_records_kernel(const PushNode *__restrict__ node, const double *__restrict__ center, int32_t n, int32_t n_padded,
                    int32_t slot_bytes, PushRec *__restrict__ out)
{
    const int32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_padded) return;
    const PushNode nd = node[k];
    const int32_t m = nd.meta, keep = (m >> 4) & 15, take = (m >> 8) & 15;
    const bool leaf = m & 1, skip = m & 2;
    PushRec r;
    r.w = skip ? 0.0 : nd.w;
    r.c = (center && k < n && !leaf && !skip) ? center[k] : 0.0;      // leaf-like nodes carry no centre
    r.keep_off = keep ? (keep - 1) * slot_bytes : -1;
    r.take_off = leaf ? -2 : (take ? (take - 1) * slot_bytes : -1);
    r.store = skip ? 0 : -1;
    r.cp_slot = (int32_t)((uint32_t)m >> 16) - 1;
    out[k] = r;
}

struct PushTmaParams {
    const PushRec *rec;        // [n_chunks * 32]
    float *PT;
    int64_t col0, N, ldt;
    uint32_t ldp;
    int32_t n, n_chunks, n_groups, n_items;
    double *T, *CP, *R;        // chunk totals, checkpoints (nullptr without spanning nodes), error partials (nullptr: none)
    double *GT;                // totals of whole groups of PT_GROUP chunks (nullptr without spanning nodes)
};

__device__ __forceinline__ uint32_t pu_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// try_wait with a suspend-time hint: a waiting warp sleeps in the barrier unit instead of spinning through the issue slots
// of the warps that have work (the producer waits most of the time); bounded, so a broken pipeline is a CUDA error
__device__ __forceinline__ void pu_mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity), "r"(20000u)
            : "memory");
        if (spin > (1u << 22)) __trap();
    }
}
// Parked prefixes are touched only through these two (volatile, so they keep their order among themselves; no memory
// clobber, so the compiler is free to move the record loads across them — the park array aliases nothing it reads).
__device__ __forceinline__ void park_store(uint32_t park, int32_t off, double r)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ge.s32 p, %1, 0;\n\t@p st.shared.f64 [%0], %2;\n\t}" ::"r"(park + (uint32_t)off), "r"(off), "d"(r));
}
__device__ __forceinline__ double park_load(uint32_t park, int32_t off)
{
    double v;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ge.s32 p, %2, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\t@p ld.shared.f64 %0, [%1];\n\t}"
                 : "=d"(v)
                 : "r"(park + (uint32_t)off), "r"(off));
    return v;
}

// the store of a checkpoint is rare (1.5 per chunk): kept out of line so that the node loop pays a compare and a branch
// for it, not a dozen predicated address instructions
__device__ __noinline__ void checkpoint_store(double *CP, int64_t ldt, int32_t slot, int64_t s, double r)
{
    CP[(int64_t)slot * ldt + s] = r;
}
__device__ __forceinline__ void store_if(float *q, float v, int32_t cond)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p st.global.f32 [%0], %1;\n\t}" ::"l"(q), "f"(v), "r"(cond) : "memory");
}

template <typename TX, int MODE, bool WANT_ERR, int SPT>
__global__ void __launch_bounds__(PT_SAMPLES + 32, SPT == 1 ? 2 : 1)
pushup_tma_kernel(const __grid_constant__ CUtensorMap tmap, const PushTmaParams p)
{
    // SPT samples per thread: the plan decoding, record loads and address arithmetic of a node are shared by the SPT
    // samples a thread owns (rows tid, tid + 256, ... of a tile of SPT x 256 samples); SPT = 2 halves that overhead per
    // entry and doubles the independent work per thread, at one CTA per SM instead of two.
    constexpr int NPS = 128 / (int)sizeof(TX);          // nodes per stage: 16 (float64) or 32 (float32)
    constexpr int SPC = KT / NPS;                        // stages per chunk: 2 or 1
    constexpr uint32_t STAGE = SPT * PT_STAGE_BYTES;
    constexpr int TILE_SAMPLES = SPT * PT_SAMPLES;
    extern __shared__ unsigned char pu_smem[];
    const uint32_t base = (pu_smem_u32(pu_smem) + 1023u) & ~1023u;      // SWIZZLE_128B atoms sit on 1 KB boundaries
    const uint32_t recs0 = base + PT_STAGES * STAGE;                    // per stage: the records of its nodes
    const uint32_t park0 = recs0 + PT_STAGES * PT_REC_BYTES;
    const uint32_t full0 = park0 + SPT * PT_PARK_BYTES, empty0 = full0 + 8 * PT_STAGES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < PT_STAGES; s++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full0 + 8 * s), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty0 + 8 * s), "r"(PT_SAMPLES / 32));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp == PT_SAMPLES / 32) {
        // ===================== producer: one lane issues every TMA load =====================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
            uint32_t it = 0;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                const int sb = item / p.n_groups, g = item % p.n_groups;
                const int st_end = min(p.n_chunks, (g + 1) * PT_GROUP) * SPC;
                for (int st = g * PT_GROUP * SPC; st < st_end; ++st, ++it) {
                    const uint32_t slot = it % PT_STAGES, phase = (it / PT_STAGES) & 1u;
                    pu_mbar_wait(empty0 + 8 * slot, phase ^ 1u);
                    const uint32_t full = full0 + 8 * slot;
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"(STAGE + NPS * 32) : "memory");
#pragma unroll
                    for (int q = 0; q < SPT; q++)      // one box of 256 samples x 128 bytes per owned sample
                        asm volatile(
                            "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                            ::"r"(base + slot * STAGE + q * PT_STAGE_BYTES), "l"(&tmap), "r"(full), "r"(st * NPS),
                            "r"(sb * TILE_SAMPLES + q * PT_SAMPLES)
                            : "memory");
                    // the stage's node records ride along (the table is padded to whole chunks)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(recs0 + slot * PT_REC_BYTES), "l"(p.rec + (int64_t)st * NPS), "r"(NPS * 32), "r"(full)
                                 : "memory");
                }
            }
        }
        return;
    }

    // ===================== consumers =====================
    unsigned char *gbase = pu_smem + (base - pu_smem_u32(pu_smem));       // generic pointer to the aligned ring
    const uint32_t sw = (uint32_t)(tid & 7) << 4;                         // XOR term of this row's 16-byte units
    const uint32_t row = (uint32_t)tid * 128u;
    const uint32_t park = park0 + tid * (uint32_t)sizeof(double);         // [sample of the thread][slot][thread]
    const uint32_t ldp = p.ldp;
    uint32_t it = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int sb = item / p.n_groups, g = item % p.n_groups;
        const int c_end = min(p.n_chunks, (g + 1) * PT_GROUP);
        const int64_t s0 = (int64_t)sb * TILE_SAMPLES + tid;
        int32_t live_mask[SPT];
        double err[SPT], group_total[SPT];
#pragma unroll
        for (int q = 0; q < SPT; q++) {
            live_mask[q] = (s0 + q * PT_SAMPLES < p.N) ? -1 : 0;
            err[q] = 0.0;
            group_total[q] = 0.0;
        }
        for (int c = g * PT_GROUP; c < c_end; ++c) {
            float *pt = p.PT + (int64_t)c * KT * ldp + p.col0 + s0;
            double r[SPT];
#pragma unroll
            for (int q = 0; q < SPT; q++) r[q] = 0.0;
#pragma unroll
            for (int h = 0; h < SPC; ++h, ++it) {
                const uint32_t slot = it % PT_STAGES, phase = (it / PT_STAGES) & 1u;
                const unsigned char *stage = gbase + slot * STAGE;
                const PushRec *rec = reinterpret_cast<const PushRec *>(gbase + PT_STAGES * STAGE + slot * PT_REC_BYTES);
                pu_mbar_wait(full0 + 8 * slot, phase);
                double x[SPT][NPS];
#pragma unroll
                for (int q = 0; q < SPT; q++) {
                    if (sizeof(TX) == 8) {
#pragma unroll
                        for (int u = 0; u < 8; u++) {
                            const double2 v = *reinterpret_cast<const double2 *>(stage + q * PT_STAGE_BYTES + row + (((uint32_t)u << 4) ^ sw));
                            x[q][2 * u] = v.x;
                            x[q][2 * u + 1] = v.y;
                        }
                    } else {
#pragma unroll
                        for (int u = 0; u < 8; u++) {
                            const float4 v = *reinterpret_cast<const float4 *>(stage + q * PT_STAGE_BYTES + row + (((uint32_t)u << 4) ^ sw));
                            x[q][4 * u] = (double)v.x, x[q][4 * u + 1] = (double)v.y, x[q][4 * u + 2] = (double)v.z, x[q][4 * u + 3] = (double)v.w;
                        }
                    }
                }
#pragma unroll
                for (int kk = 0; kk < NPS; kk++) {
                    // two 16-byte loads, the same address in every lane (broadcasts): weight + centre, plan
                    const double2 wc = *reinterpret_cast<const double2 *>(&rec[kk].w);
                    const int4 f = *reinterpret_cast<const int4 *>(&rec[kk].keep_off);
                    const bool leaf = f.y == -2;
                    float *row_ptr = pt + (size_t)(h * NPS + kk) * ldp;
#pragma unroll
                    for (int q = 0; q < SPT; q++) {
                        r[q] += x[q][kk];
                        park_store(park + q * PT_PARK_BYTES, f.x, r[q]);
                        const double parked = park_load(park + q * PT_PARK_BYTES, f.y);   // 0 unless an internal node has one
                        const double S = leaf ? x[q][kk] : r[q] - parked;
                        const double v = fma(wc.x, S, -wc.y);                // w S - c (leaf-like: c = 0, so v = fl(w S))
                        const float fl = (float)v;
                        store_if(row_ptr + q * PT_SAMPLES, fl, f.z & live_mask[q]);
                        if (WANT_ERR) {
                            const double e = v - (double)fl;                 // nodes written elsewhere: w = c = 0, so e = 0
                            err[q] += (MODE == FUF_MODE_L2) ? e * e : fabs(e);
                        }
                    }
                    if (f.w >= 0) {                                          // a checkpoint for the span kernel (rare)
#pragma unroll
                        for (int q = 0; q < SPT; q++)
                            if (live_mask[q]) checkpoint_store(p.CP, p.ldt, f.w, s0 + q * PT_SAMPLES, r[q]);
                    }
                }
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty0 + 8 * slot) : "memory");
            }
#pragma unroll
            for (int q = 0; q < SPT; q++) {
                if (p.T && live_mask[q]) p.T[(int64_t)c * p.ldt + s0 + q * PT_SAMPLES] = r[q];
                group_total[q] += r[q];
            }
        }
#pragma unroll
        for (int q = 0; q < SPT; q++) {
            if (p.GT && live_mask[q]) p.GT[(int64_t)g * p.ldt + s0 + q * PT_SAMPLES] = group_total[q];
            if (WANT_ERR && live_mask[q]) p.R[(int64_t)g * p.ldt + s0 + q * PT_SAMPLES] = err[q];
        }
    }
}

// err_stat[col0 + s] = sum of the per-chunk error statistics R[c][s] and of the span kernel's per-block partials
// R2[b][s].  32 samples x 8 chunk ranges per block, combined in fixed order: bit-reproducible.
__global__ void __launch_bounds__(256)
pushup_stat_kernel(const double *__restrict__ R, int32_t n_chunks, const double *__restrict__ R2, int32_t n_span_blocks,
                   int64_t N, int64_t ldt, double *__restrict__ err_stat, int64_t col0)
{
    __shared__ double part[8][33];
    const int sx = threadIdx.x & 31, py = threadIdx.x >> 5;
    const int64_t s = (int64_t)blockIdx.x * 32 + sx;
    double e[4] = {0.0, 0.0, 0.0, 0.0};
    if (s < N) {
        const int per = (n_chunks + 7) / 8;
        const int c0 = py * per, c1 = min(n_chunks, c0 + per);
        int c = c0;
        for (; c + 4 <= c1; c += 4) {
#pragma unroll
            for (int u = 0; u < 4; u++) e[u] += R[(int64_t)(c + u) * ldt + s];
        }
        for (; c < c1; c++) e[0] += R[(int64_t)c * ldt + s];
    }
    part[py][sx] = (e[0] + e[1]) + (e[2] + e[3]);
    __syncthreads();
    if (py == 0 && s < N) {
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < 8; q++) t += part[q][sx];
        for (int b = 0; b < n_span_blocks; b++) t += R2[(int64_t)b * ldt + s];
        err_stat[col0 + s] = t;
    }
}

// sum of A[c][s] for c in [lo, hi): eight independent loads in flight per round, four accumulators
__device__ __forceinline__ void span_sum(const double *__restrict__ A, int64_t ldt, int64_t s, int lo, int hi, double (&acc)[4])
{
    int c = lo;
    for (; c + 8 <= hi; c += 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = A[(int64_t)(c + u) * ldt + s];
#pragma unroll
        for (int u = 0; u < 8; u++) acc[u & 3] += v[u];
    }
    for (; c < hi; c++) acc[0] += A[(int64_t)c * ldt + s];
}

template <int MODE>
__global__ void pushup_span_kernel(const SpanNode *__restrict__ span, int32_t n_span, const double *__restrict__ w,
                                   const double *__restrict__ center, const double *__restrict__ T,
                                   const double *__restrict__ GT, const double *__restrict__ CP, int64_t ldt, int64_t N,
                                   float *__restrict__ PT, int64_t ldp, double *__restrict__ P64T, int64_t ldp64,
                                   int64_t col0, double *__restrict__ R2)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    double err = 0.0;
    for (int q = blockIdx.y; q < n_span; q += gridDim.y) {
        SpanNode sn = span[q];
        // subtree sum = tail of the first chunk + whole chunks in between + head of the last chunk: only
        // same-sign terms are added (no difference of long prefix sums), so the result is accurate relative
        // to the subtree's own mass.  With group totals GT (one per PT_GROUP consecutive chunks, written by the TMA
        // kernel) the chunks in between cost at most 2 (PT_GROUP - 1) + (chunks / PT_GROUP) loads: the root's 936 chunk
        // totals become 130 loads, and the longest dependent-latency chain of the launch shrinks accordingly.
        double head = T[(int64_t)sn.chunk_first * ldt + s];
        if (sn.slot_start >= 0) head -= CP[(int64_t)sn.slot_start * ldt + s];
        double mid[4] = {0.0, 0.0, 0.0, 0.0};
        const int lo = sn.chunk_first + 1, hi = sn.chunk_last;
        if (GT && hi - lo >= 2 * PT_GROUP) {
            const int g_lo = (lo + PT_GROUP - 1) / PT_GROUP, g_hi = hi / PT_GROUP;     // whole groups [g_lo, g_hi)
            span_sum(T, ldt, s, lo, g_lo * PT_GROUP, mid);
            span_sum(GT, ldt, s, g_lo, g_hi, mid);
            span_sum(T, ldt, s, g_hi * PT_GROUP, hi, mid);
        } else {
            span_sum(T, ldt, s, lo, hi, mid);
        }
        const double tail = CP[(int64_t)sn.slot_k * ldt + s];
        const double S = (head + tail) + ((mid[0] + mid[1]) + (mid[2] + mid[3]));
        const double v = w[sn.k] * S;
        if (P64T) P64T[(int64_t)sn.k * ldp64 + col0 + s] = v;
        if (PT) {
            const double c = v - (center ? center[sn.k] : 0.0);
            const float f = (float)c;
            PT[(int64_t)sn.k * ldp + col0 + s] = f;
            const double e = c - (double)f;
            err += (MODE == FUF_MODE_L2) ? e * e : fabs(e);
        }
    }
    if (R2) R2[(int64_t)blockIdx.y * ldt + s] = err;
}

// Centre support: column mean of the first m rows (double), and the leaf mask.
template <typename TX>
__global__ void column_mean_kernel(const TX *__restrict__ X, int64_t ldx, int64_t m, int32_t n, double *__restrict__ out)
{
    int32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double acc = 0.0;
    for (int64_t s = 0; s < m; s++) acc += (double)X[s * ldx + k];
    out[k] = acc / (double)m;
}

// centre of ALREADY pushed node-major profiles: mean over the first m columns, zero on the leaves
__global__ void center_of_pushed_kernel(const double *__restrict__ P64T, int64_t ldp64, int64_t m, int32_t n,
                                        const uint8_t *__restrict__ has_child, double *__restrict__ out)
{
    int32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double acc = 0.0;
    if (has_child[k]) {
        for (int64_t s = 0; s < m; s++) acc += P64T[(int64_t)k * ldp64 + s];
        acc /= (double)m;
    }
    out[k] = acc;
}

__global__ void mask_leaves_kernel(double *__restrict__ center, const uint8_t *__restrict__ has_child, int32_t n)
{
    int32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n && !has_child[k]) center[k] = 0.0;
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

// Non-DFS fallback: double node-major result -> centred float + error statistic (one thread per sample).
template <int MODE>
__global__ void f64_to_f32_kernel(const double *__restrict__ src, const double *__restrict__ center,
                                  float *__restrict__ dst, int64_t ld_src, int64_t ld_dst, int64_t col0, int64_t N,
                                  int32_t n, double *__restrict__ err_stat)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    double err = 0.0;
    for (int32_t k = 0; k < n; k++) {
        const double c = src[(int64_t)k * ld_src + s] - (center ? center[k] : 0.0);
        const float f = (float)c;
        dst[(int64_t)k * ld_dst + col0 + s] = f;
        const double e = c - (double)f;
        err += (MODE == FUF_MODE_L2) ? e * e : fabs(e);
    }
    if (err_stat) err_stat[col0 + s] = err;
}

__global__ void copy_cols_f64_kernel(const double *__restrict__ src, int64_t ld_src, double *__restrict__ dst,
                                     int64_t ld_dst, int64_t col0, int64_t N, int32_t n)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int32_t k = blockIdx.y;
    if (s < N && k < n) dst[(int64_t)k * ld_dst + col0 + s] = src[(int64_t)k * ld_src + s];
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

int push_up_f32(fuf_ctx *ctx, const void *X, int x_dtype, int64_t N, int64_t ldx, int mode, const double *center,
                float *PT, int64_t ldp, double *P64T, int64_t ldp64, double *err_stat, int64_t col0,
                cudaStream_t stream)
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
        if (PT) {
            const unsigned g = (unsigned)((N + 127) / 128);
            if (mode == FUF_MODE_L1)
                f64_to_f32_kernel<FUF_MODE_L1><<<g, 128, 0, stream>>>(tmp, center, PT, N, ldp, col0, N, n, err_stat);
            else
                f64_to_f32_kernel<FUF_MODE_L2><<<g, 128, 0, stream>>>(tmp, center, PT, N, ldp, col0, N, n, err_stat);
            FUF_CUDA(cudaGetLastError());
            ctx->launches += 1;
        }
        if (P64T) {
            dim3 g((unsigned)((N + 255) / 256), n);
            copy_cols_f64_kernel<<<g, 256, 0, stream>>>(tmp, N, P64T, ldp64, col0, N, n);
            FUF_CUDA(cudaGetLastError());
            ctx->launches += 1;
        }
        return FUF_OK;
    }
    const int64_t ldt = (N + 31) / 32 * 32;
    const bool want_err = (err_stat != nullptr && PT != nullptr);
    const int span_blocks = ctx->n_span > 0 ? min(ctx->n_span, 64) : 0;
    // The TMA kernel serves the production case (fp32 output only); its X tile is a tensor-map box, which needs a
    // 16-byte aligned base and row pitch.  Everything else (float64 output wanted, odd pitch) takes the chunk kernel.
    const size_t esz = x_dtype == FUF_F32 ? 4 : 8;
    const bool use_tma = PT != nullptr && P64T == nullptr && ctx->d_push_node[mode] != nullptr &&
                         ((uintptr_t)X & 15) == 0 && ((size_t)ldx * esz) % 16 == 0 && !getenv("FUF_PUSHUP_NO_TMA") &&
                         tensor_map_encoder() != nullptr;
    const int32_t n_groups = (ctx->n_chunks + PT_GROUP - 1) / PT_GROUP;
    double *T = nullptr, *CP = nullptr, *R = nullptr, *R2 = nullptr, *GT = nullptr;
    {
        size_t need = 0;
        if (ctx->n_span > 0)
            need += ((size_t)ctx->n_chunks + (size_t)ctx->n_cp + (use_tma ? (size_t)n_groups : 0)) * (size_t)ldt * sizeof(double);
        if (want_err) need += ((size_t)(use_tma ? n_groups : ctx->n_chunks) + (size_t)span_blocks) * (size_t)ldt * sizeof(double);
        if (need) {
            int rc = ctx->ws_push.reserve(need);
            if (rc) return rc;
            double *base = (double *)ctx->ws_push.ptr;
            if (ctx->n_span > 0) {
                T = base;
                CP = T + (size_t)ctx->n_chunks * ldt;
                base = CP + (size_t)ctx->n_cp * ldt;
                if (use_tma) {
                    GT = base;
                    base = GT + (size_t)n_groups * ldt;
                }
            }
            if (want_err) {
                R = base;
                R2 = R + (size_t)(use_tma ? n_groups : ctx->n_chunks) * ldt;
            }
        }
    }
    if (use_tma) {
        CUtensorMap tmap;
        const cuuint64_t gdim[2] = {(cuuint64_t)n, (cuuint64_t)N}, gstride[1] = {(cuuint64_t)ldx * esz};
        const cuuint32_t box[2] = {(cuuint32_t)(128 / esz), PT_SAMPLES}, estr[2] = {1, 1};   // one 128-byte row per sample
        CUresult cres = tensor_map_encoder()(&tmap, x_dtype == FUF_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64,
                                             2, const_cast<void *>(X), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cres != CUDA_SUCCESS) {
            set_error("fuf_push_up: cuTensorMapEncodeTiled failed with CUresult %d (N=%lld n=%d ldx=%lld)", (int)cres,
                      (long long)N, n, (long long)ldx);
            return FUF_ECUDA;
        }
        // the job's node records: static plan + centre (1 KB per chunk, built by a ~1 us kernel)
        const int32_t n_padded = ctx->n_chunks * KT;
        int rc2 = ctx->ws_rec.reserve((size_t)n_padded * sizeof(PushRec));
        if (rc2) return rc2;
        PushRec *recs = (PushRec *)ctx->ws_rec.ptr;
        push_records_kernel<<<(unsigned)((n_padded + 255) / 256), 256, 0, stream>>>(ctx->d_push_node[mode], center, n, n_padded,
                                                                                   PT_SAMPLES * (int)sizeof(double), recs);
        FUF_CUDA(cudaGetLastError());
        ctx->launches += 1;
        PushTmaParams p;
        p.rec = recs;
        p.PT = PT;
        FUF_REQUIRE(ldp < ((int64_t)1 << 31), "fuf_push_up: ldp too large");
        p.ldp = (uint32_t)ldp;
        p.col0 = col0;
        p.N = N;
        p.ldt = ldt;
        p.n = n;
        p.n_chunks = ctx->n_chunks;
        p.n_groups = n_groups;
        // One sample per thread, two CTAs per SM.  (FUF_PUSHUP_SPT=2: two samples per thread, one CTA per SM — measured
        // the same 0.65 ms at N = 10,000: the kernel is at the DRAM roof either way, and smaller tiles suit small N.)
        const char *spt_env = getenv("FUF_PUSHUP_SPT");
        const int spt = (spt_env && atoi(spt_env) == 2) ? 2 : 1;
        const int64_t n_sb = (N + spt * PT_SAMPLES - 1) / (spt * PT_SAMPLES);
        FUF_REQUIRE(n_sb * n_groups < ((int64_t)1 << 31), "fuf_push_up: N = %lld too large for one call", (long long)N);
        p.n_items = (int32_t)(n_sb * n_groups);
        p.T = T;
        p.CP = CP;
        p.R = want_err ? R : nullptr;
        p.GT = GT;
        void (*kern)(const CUtensorMap, const PushTmaParams) = nullptr;
#define FUF_TMA_KERN(TX, SPT_)                                                                                          \
    kern = mode == FUF_MODE_L1 ? (want_err ? pushup_tma_kernel<TX, 0, true, SPT_> : pushup_tma_kernel<TX, 0, false, SPT_>) \
                               : (want_err ? pushup_tma_kernel<TX, 1, true, SPT_> : pushup_tma_kernel<TX, 1, false, SPT_>)
        if (x_dtype == FUF_F32) {
            if (spt == 2) FUF_TMA_KERN(float, 2); else FUF_TMA_KERN(float, 1);
        } else {
            if (spt == 2) FUF_TMA_KERN(double, 2); else FUF_TMA_KERN(double, 1);
        }
#undef FUF_TMA_KERN
        const uint32_t smem_t = pt_smem_bytes(spt);
        FUF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
        const int grid_t = (int)std::min<int64_t>(p.n_items, (spt == 1 ? 2 : 1) * ctx->sm_count);
        kern<<<grid_t, PT_SAMPLES + 32, smem_t, stream>>>(tmap, p);
        FUF_CUDA(cudaGetLastError());
        ctx->launches += 1;
    } else {
    FUF_REQUIRE((N + TI - 1) / TI <= 65535, "fuf_push_up: N = %lld too large for one call (max %d)", (long long)N,
                65535 * TI);
    dim3 grid(ctx->n_chunks, (unsigned)((N + TI - 1) / TI));
    const size_t smem_f = sizeof(double) * KT * (TI + 1), smem_d = smem_f;
    // one instantiation per (input type, mode, which outputs exist): no pointer tests in the per-node loop
#define FUF_CHUNK(TX, MODE_, HP, H6)                                                                                  \
    do {                                                                                                              \
        auto kern = pushup_chunk_kernel<TX, MODE_, HP, H6>;                                                           \
        FUF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f));               \
        kern<<<grid, TI, smem_f, stream>>>((const TX *)X, ldx, N, n, ctx->d_w[mode], ctx->d_parent_local,             \
                                           ctx->d_cp_after, ctx->d_is_span, center, PT, ldp, P64T, ldp64, col0, T, CP, \
                                           R, ldt);                                                                   \
    } while (0)
#define FUF_CHUNK_OUT(TX, MODE_)                                                                                      \
    do {                                                                                                              \
        if (PT && P64T) FUF_CHUNK(TX, MODE_, true, true);                                                             \
        else if (PT) FUF_CHUNK(TX, MODE_, true, false);                                                               \
        else FUF_CHUNK(TX, MODE_, false, true);                                                                       \
    } while (0)
    if (x_dtype == FUF_F32) {
        if (mode == FUF_MODE_L1) FUF_CHUNK_OUT(float, 0); else FUF_CHUNK_OUT(float, 1);
    } else {
        if (mode == FUF_MODE_L1) FUF_CHUNK_OUT(double, 0); else FUF_CHUNK_OUT(double, 1);
    }
#undef FUF_CHUNK_OUT
#undef FUF_CHUNK
    FUF_CUDA(cudaGetLastError());
    ctx->launches += 1;
    }
    if (ctx->n_span > 0) {
        dim3 g2((unsigned)((N + 127) / 128), (unsigned)span_blocks);
        if (mode == FUF_MODE_L1)
            pushup_span_kernel<FUF_MODE_L1><<<g2, 128, 0, stream>>>(ctx->d_span, ctx->n_span, ctx->d_w[mode], center, T, GT, CP,
                                                                    ldt, N, PT, ldp, P64T, ldp64, col0, R2);
        else
            pushup_span_kernel<FUF_MODE_L2><<<g2, 128, 0, stream>>>(ctx->d_span, ctx->n_span, ctx->d_w[mode], center, T, GT, CP,
                                                                    ldt, N, PT, ldp, P64T, ldp64, col0, R2);
        FUF_CUDA(cudaGetLastError());
        ctx->launches += 1;
    }
    if (want_err) {
        pushup_stat_kernel<<<(unsigned)((N + 31) / 32), 256, 0, stream>>>(R, use_tma ? n_groups : ctx->n_chunks, R2,
                                                                         span_blocks, N, ldt, err_stat, col0);
        FUF_CUDA(cudaGetLastError());
        ctx->launches += 1;
    }
    return FUF_OK;
}

// centre = pushed column mean of the first m rows, zero on the leaves (see the header comment)
int push_up_center(fuf_ctx *ctx, const void *X, int x_dtype, int64_t m, int64_t ldx, int mode, double *center_out,
                   cudaStream_t stream)
{
    const int32_t n = ctx->n;
    int rc = ctx->ws_center.reserve((size_t)n * sizeof(double));
    if (rc) return rc;
    double *mean = (double *)ctx->ws_center.ptr;
    const unsigned g = (unsigned)((n + 255) / 256);
    if (x_dtype == FUF_F32)
        column_mean_kernel<float><<<g, 256, 0, stream>>>((const float *)X, ldx, m, n, mean);
    else
        column_mean_kernel<double><<<g, 256, 0, stream>>>((const double *)X, ldx, m, n, mean);
    FUF_CUDA(cudaGetLastError());
    ctx->launches += 1;
    rc = push_up_f32(ctx, mean, FUF_F64, 1, n, mode, nullptr, nullptr, 0, center_out, 1, nullptr, 0, stream);
    if (rc) return rc;
    mask_leaves_kernel<<<g, 256, 0, stream>>>(center_out, ctx->d_has_child, n);
    FUF_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return FUF_OK;
}

}  // namespace fuf

using namespace fuf;

extern "C" int fuf_push_up(fuf_ctx *ctx, const void *X, int x_dtype, int64_t N, int64_t ldx, int mode,
                           const double *center, float *PT, int64_t ldp, double *P64T, int64_t ldp64,
                           double *err_stat, int64_t col0, void *stream)
{
    FUF_REQUIRE(ctx && X && (PT || P64T), "fuf_push_up: null argument");
    FUF_REQUIRE(mode == FUF_MODE_L1 || mode == FUF_MODE_L2, "fuf_push_up: bad mode %d", mode);
    FUF_REQUIRE(x_dtype == FUF_F32 || x_dtype == FUF_F64, "fuf_push_up: bad x_dtype %d", x_dtype);
    FUF_REQUIRE(N >= 0 && ldx >= ctx->n && col0 >= 0 && (!PT || col0 + N <= ldp) && (!P64T || col0 + N <= ldp64),
                "fuf_push_up: bad shape N=%lld ldx=%lld ldp=%lld ldp64=%lld col0=%lld n=%d", (long long)N,
                (long long)ldx, (long long)ldp, (long long)ldp64, (long long)col0, ctx->n);
    FUF_CUDA(cudaSetDevice(ctx->device));
    return push_up_f32(ctx, X, x_dtype, N, ldx, mode, center, PT, ldp, P64T, ldp64, err_stat, col0,
                       (cudaStream_t)stream);
}

extern "C" int fuf_push_up_center(fuf_ctx *ctx, const void *X, int x_dtype, int64_t m, int64_t ldx, int mode,
                                  double *center_out, void *stream)
{
    FUF_REQUIRE(ctx && X && center_out, "fuf_push_up_center: null argument");
    FUF_REQUIRE(mode == FUF_MODE_L1 || mode == FUF_MODE_L2, "fuf_push_up_center: bad mode %d", mode);
    FUF_REQUIRE(x_dtype == FUF_F32 || x_dtype == FUF_F64, "fuf_push_up_center: bad x_dtype %d", x_dtype);
    FUF_REQUIRE(m >= 1 && ldx >= ctx->n, "fuf_push_up_center: bad shape m=%lld ldx=%lld", (long long)m, (long long)ldx);
    FUF_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->contiguous) {  // hand-built non-DFS order: no centring (the exact fallback handles it)
        FUF_CUDA(cudaMemsetAsync(center_out, 0, (size_t)ctx->n * sizeof(double), (cudaStream_t)stream));
        return FUF_OK;
    }
    return push_up_center(ctx, X, x_dtype, m, ldx, mode, center_out, (cudaStream_t)stream);
}

extern "C" int fuf_center_of_pushed(fuf_ctx *ctx, const double *P64T, int64_t m, int64_t ldp64, double *center_out,
                                    void *stream)
{
    FUF_REQUIRE(ctx && P64T && center_out, "fuf_center_of_pushed: null argument");
    FUF_REQUIRE(m >= 1 && ldp64 >= m, "fuf_center_of_pushed: bad shape m=%lld ldp64=%lld", (long long)m, (long long)ldp64);
    FUF_CUDA(cudaSetDevice(ctx->device));
    center_of_pushed_kernel<<<(unsigned)((ctx->n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        P64T, ldp64, m, ctx->n, ctx->d_has_child, center_out);
    FUF_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return FUF_OK;
}

extern "C" int fuf_round_pushed(fuf_ctx *ctx, const double *P64T, int64_t N, int64_t ldp64, int mode,
                                const double *center, float *PT, int64_t ldp, double *err_stat, int64_t col0,
                                void *stream)
{
    FUF_REQUIRE(ctx && P64T && PT, "fuf_round_pushed: null argument");
    FUF_REQUIRE(mode == FUF_MODE_L1 || mode == FUF_MODE_L2, "fuf_round_pushed: bad mode %d", mode);
    FUF_REQUIRE(N >= 0 && col0 >= 0 && col0 + N <= ldp && col0 + N <= ldp64, "fuf_round_pushed: bad shape");
    if (N == 0) return FUF_OK;
    FUF_CUDA(cudaSetDevice(ctx->device));
    const unsigned g = (unsigned)((N + 127) / 128);
    // source and destination use the same column offset
    if (mode == FUF_MODE_L1)
        f64_to_f32_kernel<FUF_MODE_L1><<<g, 128, 0, (cudaStream_t)stream>>>(P64T + col0, center, PT, ldp64, ldp, col0, N,
                                                                           ctx->n, err_stat);
    else
        f64_to_f32_kernel<FUF_MODE_L2><<<g, 128, 0, (cudaStream_t)stream>>>(P64T + col0, center, PT, ldp64, ldp, col0, N,
                                                                           ctx->n, err_stat);
    FUF_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return FUF_OK;
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
