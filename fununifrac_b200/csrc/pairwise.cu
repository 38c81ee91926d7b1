// pairwise.cu — all-pairs distance kernels over node-major pushed profiles.
//
// Replaces the N(N-1)/2 iteration Python loop of pairwise_computation
// (src/algorithms/emd_unifrac.py:42-54) and its per-pair numpy arithmetic get_L1 / get_L2
// (src/objects/profile_vector.py:9-11,17-19):
//     L1: D[i,j] = D[j,i] = sum_k |P_i[k] - P_j[k]|          L2: sum_k (P_i[k] - P_j[k])^2   (no sqrt)
// over ALL n basis entries including the unscaled root.
//
// K2 pairwise_tile_kernel (production, FP32-pipe bound: 2 FP32 lane-ops per (pair, basis entry)):
//   * persistent CTAs (one per SM) loop over their work items = (128x128 tile of D, range of basis entries): whole tiles
//     dealt round-robin for the full waves, then the tiles of the last, partial wave cut into equal contiguous spans
//     of the (tile, basis entry) space, one span per CTA (stream-K tail: no SM idles through a partial wave);
//   * warp-specialised: warps 0-7 (256 threads, 2 warpgroups) consume, the third warpgroup donates its
//     registers (setmaxnreg) and one of its lanes is the TMA producer;
//   * operand tiles A = PT[k0:k0+32][i0:i0+128], B = PT[k0:k0+32][j0:j0+128] are brought in by TMA
//     (cp.async.bulk.tensor.3d) into a 2-stage shared-memory ring (32 KB per stage) guarded by full/empty mbarriers;
//     out-of-range rows/columns are zero-filled by the TMA unit, so no edge cases in the inner loop;
//   * each consumer thread owns an 8x8 block of the tile: per basis entry 4 LDS.128 feed
//     32 FADD2 (d = a - b, packed pairs along j) + 32 FADD2 (acc += |d|)  [L2: 32 FFMA2 acc += d*d];
//   * accuracy: fp32 partial sums are folded into double accumulators (shared memory, one private
//     column per thread) every 256 basis entries, so each fp32 accumulator sees at most 256 adds;
//   * epilogue: the double tile is written through shared memory with fully coalesced rows, to
//     D[i,j] and (mirrored) D[j,i]; a tile shared by several CTAs goes to a partial buffer that a second kernel
//     reduces in fixed order (deterministic, no atomics).
// K5 pairwise_f64_kernel: validation mode, everything in double, no FMA contraction.
#include <cstring>

#include "common.cuh"

namespace fuf {

constexpr int TILE = FUF_TILE;    // 128
// Pipeline geometry, measured at N = 10,000 (tools/build_variant.sh, K2 ms per launch): 16 entries x 4 stages 91.35;
// 16 x 4 with the cross-boundary operand prefetch 91.85; 32 x 2 90.57; 32 x 3 90.81; 32 x 2 with the prefetch 89.98.
#ifndef FUF_KC
#define FUF_KC 32
#endif
#ifndef FUF_STAGES
#define FUF_STAGES 2
#endif
#ifndef FUF_PREFETCH
#define FUF_PREFETCH 1
#endif
constexpr int KC = FUF_KC;        // basis entries per pipeline stage (a multiple of 16: the inner loop is unrolled by 16)
constexpr int STAGES = FUF_STAGES;
constexpr int FLUSH_STAGES = 256 / KC;  // fold fp32 partials into double every 256 basis entries
static_assert(KC % 16 == 0 && 256 % KC == 0 && KC <= 256, "KC: a multiple of 16 that divides 256");
constexpr int NCONS = 256;        // consumer threads (8 warps, 16x16 grid of 8x8 register blocks)
constexpr int NTHREADS = NCONS + 128;  // 2 consumer warpgroups + 1 producer warpgroup (setmaxnreg works per warpgroup)
constexpr uint32_t STAGE_BYTES = 2u * KC * TILE * sizeof(float);                      // 16 KB
constexpr uint32_t ACC64_BYTES = 64u * NCONS * sizeof(double);                        // 128 KB
constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + ACC64_BYTES + 2 * STAGES * 8;  // + barriers

struct TileDesc {
    int32_t ti, tj;     // tile coordinates (tj >= ti)
    int32_t out_row0;   // first row of this tile row in the output buffer
    int32_t mirror;     // bit 0: also write the transposed tile to (tj, ti); bits 8-19: dependency + 1 (0 = none): the
                        // operands of this tile are complete once dep_flags[dependency] has reached the launch's epoch;
                        // bits 20-31: segment + 1 (0 = none): seg_done[segment] is incremented once the tile is stored
};
__device__ __host__ __forceinline__ int tile_dep(int32_t m) { return (m >> 8) & 0xfff; }
__device__ __host__ __forceinline__ int tile_seg(int32_t m) { return (m >> 20) & 0xfff; }

// One work item of a CTA: stages [s_begin, s_end) of a tile.  slot < 0: the item covers every stage and its result goes
// straight to D; otherwise the tile is shared with other CTAs and the item's sums go to partial slot `slot`.
struct WorkItem {
    int32_t tile, s_begin, s_end, slot;
};
// A tile assembled from partial slots [first_slot, first_slot + n_slots), in ascending order of the basis entries.
struct SharedTile {
    int32_t tile, first_slot, n_slots, pad;
};

struct PairParams {
    const TileDesc *tiles;
    const WorkItem *work;      // all CTAs' items; CTA c owns work[cta_first[c] .. cta_first[c + 1])
    const int32_t *cta_first;
    int32_t stages_total;      // ceil(n / KC)
    int32_t flush_stages;      // fp32 partial sums are folded into double every flush_stages stages
    int32_t shard;             // columns per shard of PT (>= padded N when there is a single block)
    int64_t N;
    double *D;
    int64_t ldd;
    double *partial;           // [n_slots][TILE*TILE]: sums of the items that cover only part of a tile
    // multi-GPU tiles mode (fuf_pairwise_tiles): operand shards arrive from peer GPUs while the kernel runs
    const uint32_t *dep_flags; // device flags, one per shard
    uint32_t epoch;            // value a flag takes once its shard has landed in this launch's step
    // refinement flags evaluated where the distances are produced (no separate scan of D): pairs i < j with
    // D < kappa (s_i + s_j) [L2: kappa (sqrt s_i + sqrt s_j)^2] are counted in flag_list[0] and the first list_cap of
    // them appended as (i, j) to flag_list[1..]
    const double *stat;
    double kappa;
    long long *flag_list;
    long long list_cap;
    // completion counters: a copy engine waiting on seg_done[s] (fuf_stream_wait_u32) moves a finished part of D while the
    // kernel is still working on the rest
    uint32_t *seg_done;
};

// one pair against the refinement criterion of refine.cu (rounding statistics of the fp32 operands)
template <int MODE>
__device__ __forceinline__ void flag_pair(const PairParams &p, int64_t gi, int64_t gj, double d, double si, double sj)
{
    const double b = (MODE == FUF_MODE_L2) ? sqrt(si) + sqrt(sj) : si + sj;
    const bool flag = (MODE == FUF_MODE_L2) ? !(d >= p.kappa * b * b) : !(d >= p.kappa * b);
    if (flag && b != 0.0) {
        const unsigned long long at = atomicAdd(reinterpret_cast<unsigned long long *>(p.flag_list), 1ull);
        if ((long long)at < p.list_cap) {
            p.flag_list[1 + 2 * at] = gi;
            p.flag_list[2 + 2 * at] = gj;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// PTX helpers (sm_100a)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a broken pipeline traps (CUDA error on the host) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (spin > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void consumer_bar()
{
    asm volatile("bar.sync 1, %0;" ::"n"(NCONS) : "memory");
}

// packed fp32x2 math (FADD2 / FFMA2 on sm_100a)
__device__ __forceinline__ uint64_t pk_sub_bcast(float a, uint64_t b)
{
    uint64_t aa, d;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("sub.f32x2 %0, %1, %2;" : "=l"(d) : "l"(aa), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t pk_add_abs(uint64_t acc, uint64_t d)
{
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(d));
    lo = fabsf(lo);
    hi = fabsf(hi);
    uint64_t ad;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ad) : "f"(lo), "f"(hi));
    asm("add.f32x2 %0, %0, %1;" : "+l"(acc) : "l"(ad));
    return acc;
}
__device__ __forceinline__ uint64_t pk_fma_sq(uint64_t acc, uint64_t d)
{
    asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(acc) : "l"(d));
    return acc;
}
__device__ __forceinline__ void pk_unpack(uint64_t v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

// ------------------------------------------------------------------------------------------------
// K2: production tile kernel
// ------------------------------------------------------------------------------------------------
template <int MODE, bool PACKED>
__global__ void __launch_bounds__(NTHREADS, 1)
pairwise_tile_kernel(const __grid_constant__ CUtensorMap tmap, const PairParams p)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    float *stage_base = reinterpret_cast<float *>(smem);
    double *acc64 = reinterpret_cast<double *>(smem + STAGES * STAGE_BYTES);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES + ACC64_BYTES);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(full0 + 8 * s, 1);           // one arrive.expect_tx by the producer + the TMA bytes
            mbar_init(empty0 + 8 * s, NCONS / 32);  // one arrive per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp >= NCONS / 32) {
        // ===================== producer warpgroup: gives its registers to the consumers =====================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == NCONS / 32 && lane == 0) {  // one elected lane issues all TMA loads
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
            uint32_t it = 0;
            for (int w = p.cta_first[blockIdx.x]; w < p.cta_first[blockIdx.x + 1]; ++w) {
                const WorkItem wi = p.work[w];
                const TileDesc td = p.tiles[wi.tile];
                const int s_begin = wi.s_begin, s_end = wi.s_end;
                const int ia = td.ti * TILE, ib = td.tj * TILE;
                if (tile_dep(td.mirror)) {
                    // this tile reads a shard that a copy engine is still bringing in from a peer GPU: wait for its
                    // flag (written in stream order behind the copy), then order the TMA reads behind the observation
                    const uint32_t *flag = p.dep_flags + (tile_dep(td.mirror) - 1);
                    for (uint32_t spin = 0;; ++spin) {
                        uint32_t seen;
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
                        if ((int32_t)(seen - p.epoch) >= 0) break;
                        if (spin > (1u << 24)) __trap();        // ~3 s: a lost copy is a CUDA error, not a hung GPU
                        __nanosleep(200);
                    }
                    asm volatile("fence.proxy.async.global;" ::: "memory");
                }
                const int a_col = ia % p.shard, a_sh = ia / p.shard;
                const int b_col = ib % p.shard, b_sh = ib / p.shard;
                for (int st = s_begin; st < s_end; ++st, ++it) {
                    const uint32_t slot = it % STAGES, phase = (it / STAGES) & 1u;
                    mbar_wait(empty0 + 8 * slot, phase ^ 1u);
                    const uint32_t full = full0 + 8 * slot;
                    mbar_expect_tx(full, STAGE_BYTES);
                    const uint32_t dst = smem_u32(stage_base) + slot * STAGE_BYTES;
                    tma_load_3d(dst, &tmap, full, a_col, st * KC, a_sh);
                    tma_load_3d(dst + STAGE_BYTES / 2, &tmap, full, b_col, st * KC, b_sh);
                }
            }
        }
        return;
    }

    // ===================== consumers: 256 threads, 8x8 register block each =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int tx = (warp & 1) * 8 + (lane & 7);   // 0..15 along i (rows of the tile)
    const int ty = (warp >> 1) * 4 + (lane >> 3); // 0..15 along j (columns of the tile)
    double *my64 = acc64 + tid;                   // slot (m*8+nn) lives at my64[(m*8+nn)*NCONS]

    uint32_t it = 0;
    for (int w = p.cta_first[blockIdx.x]; w < p.cta_first[blockIdx.x + 1]; ++w) {
        const WorkItem wi = p.work[w];
        const TileDesc td = p.tiles[wi.tile];
        const int s_begin = wi.s_begin, s_end = wi.s_end;

        float acc[8][8];
        uint64_t acc2[8][4];
#pragma unroll
        for (int m = 0; m < 8; m++) {
#pragma unroll
            for (int q = 0; q < 8; q++) acc[m][q] = 0.f;
#pragma unroll
            for (int q = 0; q < 4; q++) acc2[m][q] = 0ull;
        }
#pragma unroll
        for (int e = 0; e < 64; e++) my64[e * NCONS] = 0.0;

        auto flush = [&]() {
#pragma unroll
            for (int m = 0; m < 8; m++) {
                if (PACKED) {
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        float lo, hi;
                        pk_unpack(acc2[m][q], lo, hi);
                        my64[(m * 8 + 2 * q) * NCONS] += (double)lo;
                        my64[(m * 8 + 2 * q + 1) * NCONS] += (double)hi;
                        acc2[m][q] = 0ull;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        my64[(m * 8 + q) * NCONS] += (double)acc[m][q];
                        acc[m][q] = 0.f;
                    }
                }
            }
        };

#if FUF_PREFETCH
        if (PACKED) {
            // The operands of the next basis entry are fetched before the arithmetic of the current one — also across the
            // 16-entry unroll blocks and across stage boundaries (the next stage's barrier is awaited during the last
            // entry of the current stage), so the FMA pipe never waits for a shared-memory round trip at a boundary.
            float4 na0, na1;
            ulonglong2 nb0, nb1;
            {
                const uint32_t slot = it % STAGES, phase = (it / STAGES) & 1u;
                mbar_wait(full0 + 8 * slot, phase);
                const float *As = stage_base + slot * (STAGE_BYTES / 4);
                const float *Bs = As + KC * TILE;
                na0 = *reinterpret_cast<const float4 *>(As + tx * 4);
                na1 = *reinterpret_cast<const float4 *>(As + 64 + tx * 4);
                nb0 = *reinterpret_cast<const ulonglong2 *>(Bs + ty * 4);
                nb1 = *reinterpret_cast<const ulonglong2 *>(Bs + 64 + ty * 4);
            }
            for (int st = s_begin; st < s_end; ++st, ++it) {
                const uint32_t slot = it % STAGES;
                const float *As = stage_base + slot * (STAGE_BYTES / 4);
                const float *Bs = As + KC * TILE;
#pragma unroll 1
                for (int k0 = 0; k0 < KC; k0 += 16)
#pragma unroll
                    for (int k = k0; k < k0 + 16; k++) {
                        const float a[8] = {na0.x, na0.y, na0.z, na0.w, na1.x, na1.y, na1.z, na1.w};
                        const uint64_t b[4] = {nb0.x, nb0.y, nb1.x, nb1.y};
                        const float *An = As + (k + 1) * TILE, *Bn = Bs + (k + 1) * TILE;
                        bool fetch = true;
                        if (k + 1 == KC) {
                            fetch = st + 1 < s_end;
                            if (fetch) {
                                const uint32_t nslot = (it + 1) % STAGES, nphase = ((it + 1) / STAGES) & 1u;
                                mbar_wait(full0 + 8 * nslot, nphase);
                                An = stage_base + nslot * (STAGE_BYTES / 4);
                                Bn = An + KC * TILE;
                            }
                        }
                        if (fetch) {
                            na0 = *reinterpret_cast<const float4 *>(An + tx * 4);
                            na1 = *reinterpret_cast<const float4 *>(An + 64 + tx * 4);
                            nb0 = *reinterpret_cast<const ulonglong2 *>(Bn + ty * 4);
                            nb1 = *reinterpret_cast<const ulonglong2 *>(Bn + 64 + ty * 4);
                        }
#pragma unroll
                        for (int m = 0; m < 8; m++) {
#pragma unroll
                            for (int q = 0; q < 4; q++) {
                                const uint64_t d = pk_sub_bcast(a[m], b[q]);
                                acc2[m][q] = (MODE == FUF_MODE_L1) ? pk_add_abs(acc2[m][q], d) : pk_fma_sq(acc2[m][q], d);
                            }
                        }
                    }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty0 + 8 * slot);
                if ((st - s_begin + 1) % p.flush_stages == 0) flush();
            }
        } else
#endif
        for (int st = s_begin; st < s_end; ++st, ++it) {
            const uint32_t slot = it % STAGES, phase = (it / STAGES) & 1u;
            mbar_wait(full0 + 8 * slot, phase);
            const float *As = stage_base + slot * (STAGE_BYTES / 4);
            const float *Bs = As + KC * TILE;
#pragma unroll 1
            for (int k0 = 0; k0 < KC; k0 += 16)
#pragma unroll
            for (int k = k0; k < k0 + 16; k++) {
                const float4 a0 = *reinterpret_cast<const float4 *>(As + k * TILE + tx * 4);
                const float4 a1 = *reinterpret_cast<const float4 *>(As + k * TILE + 64 + tx * 4);
                const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                if (PACKED) {
                    const ulonglong2 b0 = *reinterpret_cast<const ulonglong2 *>(Bs + k * TILE + ty * 4);
                    const ulonglong2 b1 = *reinterpret_cast<const ulonglong2 *>(Bs + k * TILE + 64 + ty * 4);
                    const uint64_t b[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
                    for (int m = 0; m < 8; m++) {
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const uint64_t d = pk_sub_bcast(a[m], b[q]);
                            acc2[m][q] = (MODE == FUF_MODE_L1) ? pk_add_abs(acc2[m][q], d) : pk_fma_sq(acc2[m][q], d);
                        }
                    }
                } else {
                    const float4 b0 = *reinterpret_cast<const float4 *>(Bs + k * TILE + ty * 4);
                    const float4 b1 = *reinterpret_cast<const float4 *>(Bs + k * TILE + 64 + ty * 4);
                    const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                    for (int m = 0; m < 8; m++) {
#pragma unroll
                        for (int q = 0; q < 8; q++) {
                            const float d = a[m] - b[q];
                            if (MODE == FUF_MODE_L1)
                                acc[m][q] += fabsf(d);
                            else
                                acc[m][q] = fmaf(d, d, acc[m][q]);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * slot);
            if ((st - s_begin + 1) % p.flush_stages == 0) flush();
        }
        flush();
        consumer_bar();  // all double accumulators of the tile are in shared memory

        // ---- epilogue: coalesced rows through shared memory ---------------------------------
        // element (r, c) of the tile is owned by thread (tx = (r%64)/4, ty = (c%64)/4), register (m, nn)
        auto slot_of = [](int r, int c) {
            const int otx = (r & 63) >> 2, m = ((r >> 6) << 2) | (r & 3);
            const int oty = (c & 63) >> 2, nn = ((c >> 6) << 2) | (c & 3);
            const int ow = ((oty >> 2) << 1) | (otx >> 3), ol = ((oty & 3) << 3) | (otx & 7);
            return (m * 8 + nn) * NCONS + ow * 32 + ol;
        };
        const int half = tid >> 7, col = tid & 127;
        if (wi.slot >= 0) {
            double *dst = p.partial + (size_t)wi.slot * (TILE * TILE);
            for (int r = half; r < TILE; r += 2) dst[r * TILE + col] = acc64[slot_of(r, col)];
        } else {
            const int64_t gi0 = (int64_t)td.ti * TILE, gj0 = (int64_t)td.tj * TILE;
            if (gj0 + col < p.N) {
                const double sj = p.stat ? p.stat[gj0 + col] : 0.0;
                for (int r = half; r < TILE; r += 2)
                    if (gi0 + r < p.N) {
                        const double v = acc64[slot_of(r, col)];
                        p.D[((int64_t)td.out_row0 + r) * p.ldd + gj0 + col] = v;
                        if (p.stat && gi0 + r < gj0 + col) flag_pair<MODE>(p, gi0 + r, gj0 + col, v, p.stat[gi0 + r], sj);
                    }
            }
            if ((td.mirror & 1) && gi0 + col < p.N) {
                for (int c = half; c < TILE; c += 2)
                    if (gj0 + c < p.N) {
                        const double v = acc64[slot_of(col, c)];
                        p.D[(gj0 + c) * p.ldd + gi0 + col] = v;
                    }
            }
        }
        const bool count_seg = p.seg_done && wi.slot < 0 && tile_seg(td.mirror);
        if (count_seg) __threadfence();   // this thread's part of the tile is visible before the counter moves
        consumer_bar();  // the next item re-zeroes the accumulators
        if (count_seg && tid == 0) atomicAdd(p.seg_done + (tile_seg(td.mirror) - 1), 1u);
    }
}

// Sum the k-split partial tiles in fixed order and write D (+ mirror).
template <int MODE>
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const TileDesc *__restrict__ tiles, const SharedTile *__restrict__ shared, const double *__restrict__ partial, int64_t N,
                       double *__restrict__ D, int64_t ldd, const PairParams p)
{
    const SharedTile sh = shared[blockIdx.x];
    const TileDesc td = tiles[sh.tile];
    const int splits = sh.n_slots;
    const double *src = partial + (size_t)sh.first_slot * (TILE * TILE);
    const int64_t gi0 = (int64_t)td.ti * TILE, gj0 = (int64_t)td.tj * TILE;
    __shared__ double tile[32][33];
    // 4x4 sub-blocks of 32x32 so that both the direct and the mirrored write are coalesced
    for (int sb = 0; sb < 16; sb++) {
        const int r0 = (sb >> 2) * 32, c0 = (sb & 3) * 32;
        const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;  // 32 x 8
        for (int rr = ly; rr < 32; rr += 8) {
            double v = 0.0;
            for (int s = 0; s < splits; s++) v += src[(size_t)s * (TILE * TILE) + (r0 + rr) * TILE + c0 + lx];
            tile[rr][lx] = v;
            const int64_t gi = gi0 + r0 + rr, gj = gj0 + c0 + lx;
            if (gi < N && gj < N) {
                D[((int64_t)td.out_row0 + r0 + rr) * ldd + gj] = v;
                if (p.stat && gi < gj) flag_pair<MODE>(p, gi, gj, v, p.stat[gi], p.stat[gj]);
            }
        }
        __syncthreads();
        if (td.mirror & 1) {
            for (int cc = ly; cc < 32; cc += 8) {
                const int64_t gj = gj0 + c0 + cc, gi = gi0 + r0 + lx;
                if (gi < N && gj < N) D[gj * ldd + gi] = tile[lx][cc];
            }
        }
        __syncthreads();
    }
    if (p.seg_done && tile_seg(td.mirror)) {   // (a shared tile is finished here, not in the tile kernel)
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(p.seg_done + (tile_seg(td.mirror) - 1), 1u);
    }
}

// ------------------------------------------------------------------------------------------------
// K5: validation kernel, all double, 64x64 tiles, 4x4 per thread; separate mul/add (no FMA) so the
// squares match numpy's (P-Q)**2 followed by a sum.
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256)
pairwise_f64_kernel(const double *__restrict__ P, int64_t ldp, int64_t shard, int64_t shard_stride, int64_t N, int32_t n,
                    const int32_t *__restrict__ tile_rows, double *__restrict__ D, int64_t ldd)
{
    constexpr int T = 64, KB = 16;
    __shared__ double As[KB][T], Bs[KB][T];
    // full matrix: blockIdx.y = 64-row tile, output row = sample row, mirror written too;
    // tile rows: blockIdx.y = half of a compact 128-row block, output row = compact row, upper part only
    int64_t i0, out0;
    if (tile_rows) {
        const int32_t t = tile_rows[blockIdx.y >> 1];
        if (t < 0) return;
        i0 = (int64_t)t * TILE + (blockIdx.y & 1) * T;
        out0 = (int64_t)blockIdx.y * T;
    } else {
        i0 = out0 = (int64_t)blockIdx.y * T;
    }
    const int64_t j0 = (int64_t)blockIdx.x * T;
    if (j0 + T <= i0 || i0 >= N) return;   // entirely below the diagonal / beyond the samples
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    // node-major operand, single block (shard == 0: column s at P[k * ldp + s]) or all-gathered shards [G][n][shard]
    const int64_t kstride = shard ? shard : ldp;
    auto col = [&](int64_t s) { return shard ? (s / shard) * shard_stride + s % shard : s; };
    double acc[4][4] = {};
    for (int32_t k0 = 0; k0 < n; k0 += KB) {
        for (int e = tid; e < KB * T; e += 256) {
            const int kk = e / T, c = e % T;
            const int32_t k = k0 + kk;
            As[kk][c] = (k < n && i0 + c < N) ? P[(int64_t)k * kstride + col(i0 + c)] : 0.0;
            Bs[kk][c] = (k < n && j0 + c < N) ? P[(int64_t)k * kstride + col(j0 + c)] : 0.0;
        }
        __syncthreads();
#pragma unroll 4
        for (int kk = 0; kk < KB; kk++) {
            double a[4], b[4];
#pragma unroll
            for (int m = 0; m < 4; m++) a[m] = As[kk][tx * 4 + m];
#pragma unroll
            for (int q = 0; q < 4; q++) b[q] = Bs[kk][ty * 4 + q];
#pragma unroll
            for (int m = 0; m < 4; m++)
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const double d = fabs(__dsub_rn(a[m], b[q]));
                    acc[m][q] = __dadd_rn(acc[m][q], MODE == FUF_MODE_L1 ? d : __dmul_rn(d, d));
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int m = 0; m < 4; m++)
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int64_t r = tx * 4 + m, i = i0 + r, j = j0 + ty * 4 + q;
            if (i < N && j < N) {
                D[(out0 + r) * ldd + j] = acc[m][q];
                if (!tile_rows) D[j * ldd + i] = acc[m][q];
            }
        }
}

// Compact row blocks (one per tile row, upper part valid) -> full symmetric matrix.
__global__ void __launch_bounds__(256)
assemble_kernel(const double *__restrict__ blocks, const int32_t *__restrict__ tile_rows, int64_t N, int64_t ldb,
                double *__restrict__ D, int64_t ldd)
{
    const int c = blockIdx.y;            // compact block
    const int t = tile_rows[c];          // its tile row (-1: padding slot of a gathered buffer)
    if (t < 0) return;
    const int tj = t + blockIdx.x;       // tile column (only tj >= t carry data)
    const int64_t gi0 = (int64_t)t * TILE, gj0 = (int64_t)tj * TILE;
    if (gj0 >= N) return;
    __shared__ double tile[32][33];
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    for (int sb = 0; sb < 16; sb++) {
        const int r0 = (sb >> 2) * 32, c0 = (sb & 3) * 32;
        for (int rr = ly; rr < 32; rr += 8) {
            const int64_t gi = gi0 + r0 + rr, gj = gj0 + c0 + lx;
            double v = 0.0;
            if (gi < N && gj < N) {
                // diagonal tile: the lower triangle is the mirror of the upper one (fuf_refine corrects only i < j)
                if (tj == t && r0 + rr > c0 + lx)
                    v = blocks[((int64_t)c * TILE + c0 + lx) * ldb + gi0 + r0 + rr];
                else
                    v = blocks[((int64_t)c * TILE + r0 + rr) * ldb + gj];
                D[gi * ldd + gj] = v;
            }
            tile[rr][lx] = v;
        }
        __syncthreads();
        if (tj != t) {
            for (int cc = ly; cc < 32; cc += 8) {
                const int64_t gj = gj0 + c0 + cc, gi = gi0 + r0 + lx;
                if (gi < N && gj < N) D[gj * ldd + gi] = tile[lx][cc];
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// Work plan of one launch of the tile kernel.  Tiles cost the same, so the first floor(n_tiles / grid) * grid of them go
// round-robin as whole tiles — wave w runs tiles w * grid .. (w + 1) * grid - 1, which the rasterised order makes a compact
// block sharing its operand panels in L2 — and need no second pass.  The remaining R < grid tiles would leave most SMs
// idle for a whole tile time (N = 10,000: 3,160 tiles = 21.35 waves on 148 SMs; 8 GPUs: 405 = 2.74): their (tile, basis
// entry) space is cut into equal contiguous spans, one per CTA (stream-K).  Span boundaries fall on fp32 flush
// boundaries (256 basis entries), so the set of fp32 chunk sums of a pair never depends on the plan — only the order of
// the last few float64 additions does; a tile cut between CTAs is summed from its partial slots in ascending order by
// reduce_partials_kernel (fixed order, no atomics).
struct PairWorkPlan {
    int32_t grid = 0, n_slots = 0;
    std::vector<int32_t> cta_first;
    std::vector<WorkItem> work;
    std::vector<SharedTile> shared;
};
static void plan_pair_work(int32_t n_tiles, int32_t sm_count, int32_t stages_total, int32_t flush_stages, bool stream_k,
                           PairWorkPlan &plan)
{
    constexpr int32_t MIN_UNITS = 4;   // a span is at least 4 flush units (1,024 basis entries) long
    const int32_t units = (stages_total + flush_stages - 1) / flush_stages;   // flush units per tile
    const int32_t full_waves = n_tiles / sm_count, rest = n_tiles - full_waves * sm_count;
    const int64_t rest_units = (int64_t)rest * units;
    int32_t tail_ctas = 0;
    if (rest > 0)
        tail_ctas = stream_k ? (int32_t)std::max<int64_t>(rest, std::min<int64_t>(sm_count, rest_units / MIN_UNITS)) : rest;
    plan.grid = full_waves > 0 ? sm_count : tail_ctas;
    std::vector<std::vector<WorkItem>> per(plan.grid);
    for (int32_t w = 0; w < full_waves; w++)
        for (int32_t c = 0; c < sm_count; c++) per[c].push_back(WorkItem{w * sm_count + c, 0, stages_total, -1});
    for (int32_t c = 0; c < tail_ctas; c++) {
        int64_t lo = rest_units * c / tail_ctas;
        const int64_t hi = rest_units * (c + 1) / tail_ctas;
        while (lo < hi) {
            const int32_t r = (int32_t)(lo / units), u0 = (int32_t)(lo % units);
            const int32_t u1 = (int32_t)std::min<int64_t>(units, u0 + (hi - lo));
            const bool whole = u0 == 0 && u1 == units;
            per[c].push_back(WorkItem{full_waves * sm_count + r, u0 * flush_stages, std::min(stages_total, u1 * flush_stages),
                                      whole ? -1 : 0});
            lo += u1 - u0;
        }
    }
    // partial slots in (tile, basis entry) order: CTAs hold ascending spans, so walking them in order numbers the slots
    plan.cta_first.assign(1, 0);
    for (int32_t c = 0; c < plan.grid; c++) {
        for (WorkItem wi : per[c]) {
            if (wi.slot >= 0 && !(wi.s_begin == 0 && wi.s_end == stages_total)) {
                wi.slot = plan.n_slots++;
                if (plan.shared.empty() || plan.shared.back().tile != wi.tile) plan.shared.push_back(SharedTile{wi.tile, wi.slot, 0, 0});
                plan.shared.back().n_slots++;
            } else {
                wi.slot = -1;
            }
            plan.work.push_back(wi);
        }
        plan.cta_first.push_back((int32_t)plan.work.size());
    }
}

// Test hook, host only: the work plan for n_tiles tiles on sm_count SMs as (cta, tile, s_begin, s_end, slot) rows.
extern "C" int fuf_debug_pair_plan(int32_t n_tiles, int32_t sm_count, int32_t stages_total, int32_t flush_stages,
                                   int32_t stream_k, int32_t *rows_out, int32_t max_rows, int32_t *n_rows, int32_t *grid)
{
    FUF_REQUIRE(n_tiles >= 0 && sm_count > 0 && stages_total > 0 && flush_stages > 0 && n_rows && grid, "fuf_debug_pair_plan: bad argument");
    PairWorkPlan plan;
    if (n_tiles > 0) plan_pair_work(n_tiles, sm_count, stages_total, flush_stages, stream_k != 0, plan);
    *n_rows = (int32_t)plan.work.size();
    *grid = plan.grid;
    if (rows_out) {
        FUF_REQUIRE(max_rows >= *n_rows, "fuf_debug_pair_plan: %d rows needed", *n_rows);
        for (int32_t c = 0; c < plan.grid; c++)
            for (int32_t w = plan.cta_first[c]; w < plan.cta_first[c + 1]; w++) {
                const WorkItem &wi = plan.work[w];
                int32_t *o = rows_out + 5 * (size_t)w;
                o[0] = c, o[1] = wi.tile, o[2] = wi.s_begin, o[3] = wi.s_end, o[4] = wi.slot;
            }
    }
    return FUF_OK;
}

int pairwise_f32(fuf_ctx *ctx, const float *PT, int64_t N, int64_t ldp, int64_t shard, int64_t shard_stride, int mode,
                 const int32_t *tile_rows_h, int32_t n_tile_rows, double *D, int64_t ldd, int flags,
                 cudaStream_t stream, int32_t col_tile_begin, int32_t col_tile_end, int32_t n_rows_override,
                 int32_t flush_stages, const PairTileList *list)
{
    if (N == 0) return FUF_OK;
    if (flush_stages <= 0) flush_stages = FLUSH_STAGES;
    const int32_t n = n_rows_override > 0 ? n_rows_override : ctx->n;
    const int64_t tiles_per_dim = (N + TILE - 1) / TILE;
    FUF_REQUIRE(((uintptr_t)PT & 15) == 0, "fuf_pairwise: PT must be 16-byte aligned");
    FUF_REQUIRE(tiles_per_dim < (1 << 20), "fuf_pairwise: N too large");

    // ---- TMA descriptor over PT: dims (sample-in-shard, node, shard) -------------------------
    EncodeTiledFn encode = tensor_map_encoder();
    if (!encode) {
        set_error("fuf_pairwise: cuTensorMapEncodeTiled is not available from this driver");
        return FUF_ECUDA;
    }
    cuuint64_t gdim[3], gstride[2];
    int32_t shard_cols;
    if (shard == 0) {
        FUF_REQUIRE(ldp % 4 == 0 && ldp >= N, "fuf_pairwise: ldp=%lld must be a multiple of 4 and >= N=%lld",
                    (long long)ldp, (long long)N);
        gdim[0] = (cuuint64_t)N;  // columns beyond N are zero-filled by the TMA unit
        gdim[1] = (cuuint64_t)n;
        gdim[2] = 1;
        gstride[0] = (cuuint64_t)ldp * sizeof(float);
        gstride[1] = (cuuint64_t)ldp * sizeof(float) * (cuuint64_t)n;
        shard_cols = (int32_t)(tiles_per_dim * TILE);
    } else {
        FUF_REQUIRE(shard % TILE == 0 && shard > 0, "fuf_pairwise: shard=%lld must be a positive multiple of %d",
                    (long long)shard, TILE);
        FUF_REQUIRE(shard_stride % 4 == 0 && shard_stride >= shard * (int64_t)n,
                    "fuf_pairwise: shard_stride=%lld too small or not a multiple of 4", (long long)shard_stride);
        const int64_t n_shards = (N + shard - 1) / shard;
        gdim[0] = (cuuint64_t)shard;
        gdim[1] = (cuuint64_t)n;
        gdim[2] = (cuuint64_t)n_shards;
        gstride[0] = (cuuint64_t)shard * sizeof(float);
        gstride[1] = (cuuint64_t)shard_stride * sizeof(float);
        shard_cols = (int32_t)shard;
    }
    const cuuint32_t box[3] = {TILE, KC, 1}, estr[3] = {1, 1, 1};
    CUtensorMap tmap;
    CUresult cres = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)PT, gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cres != CUDA_SUCCESS) {
        set_error("fuf_pairwise: cuTensorMapEncodeTiled failed with CUresult %d (N=%lld n=%d ldp=%lld shard=%lld)",
                  (int)cres, (long long)N, n, (long long)ldp, (long long)shard);
        return FUF_ECUDA;
    }

    // ---- tile list: built once per (N, rows, column band), kept on the device (ctx->tables) ----------
    const bool full = (tile_rows_h == nullptr);
    const int32_t n_rows = full ? (int32_t)tiles_per_dim : n_tile_rows;
    const bool band = full && col_tile_end > 0 && !list;
    const int32_t cb = band ? std::max(0, col_tile_begin) : 0;
    const int32_t ce = band ? std::min<int32_t>(col_tile_end, (int32_t)tiles_per_dim) : 0;
    int32_t n_tiles = 0;
    if (list) {
        // explicit (ti, tj, dependency) triples, global addressing: D[i, j] and D[j, i] of the full matrix
        for (int32_t c = 0; c < list->n_tiles; c++) {
            const int32_t ti = list->tiles_h[3 * c], tj = list->tiles_h[3 * c + 1], dep = list->tiles_h[3 * c + 2];
            FUF_REQUIRE(ti >= 0 && ti <= tj && tj < tiles_per_dim, "fuf_pairwise_tiles: tile (%d, %d) outside the upper "
                        "triangle of %lld tile rows", ti, tj, (long long)tiles_per_dim);
            FUF_REQUIRE(dep >= -1 && dep < 4095 && (dep < 0 || list->dep_flags), "fuf_pairwise_tiles: bad dependency %d", dep);
            const int32_t seg = list->seg_h ? list->seg_h[c] : -1;
            FUF_REQUIRE(seg >= -1 && seg < 4095 && (seg < 0 || list->seg_done), "fuf_pairwise_tiles: bad segment %d", seg);
        }
        n_tiles = list->n_tiles;
    } else if (band) {
        for (int32_t tj = cb; tj < ce; tj++) n_tiles += tj + 1;
    } else {
        for (int32_t c = 0; c < n_rows; c++) {
            const int32_t t = full ? c : tile_rows_h[c];
            FUF_REQUIRE(t >= 0 && t < tiles_per_dim, "fuf_pairwise: tile row %d out of range [0,%lld)", t,
                        (long long)tiles_per_dim);
            n_tiles += (int32_t)tiles_per_dim - t;
        }
    }
    if (n_tiles == 0) return FUF_OK;
    std::string key(list ? "pt" : "pw");
    key_append(key, tiles_per_dim);
    key_append(key, cb);
    key_append(key, ce);
    key_append(key, n_rows);
    if (list) key.append((const char *)list->tiles_h, (size_t)n_tiles * 3 * sizeof(int32_t));
    if (list && list->seg_h) key.append((const char *)list->seg_h, (size_t)n_tiles * sizeof(int32_t));
    else if (!full) key.append((const char *)tile_rows_h, (size_t)n_rows * sizeof(int32_t));
    TileDesc *d_tiles = (TileDesc *)ctx->tables.find(key);
    if (!d_tiles) {
        std::vector<TileDesc> tiles;
        tiles.reserve(n_tiles);
        if (list) {
            for (int32_t c = 0; c < n_tiles; c++) {
                const int32_t ti = list->tiles_h[3 * c], tj = list->tiles_h[3 * c + 1], dep = list->tiles_h[3 * c + 2];
                const int32_t seg = list->seg_h ? list->seg_h[c] : -1;
                tiles.push_back(TileDesc{ti, tj, ti * TILE, (tj != ti ? 1 : 0) | ((dep + 1) << 8) | ((seg + 1) << 20)});
            }
        } else if (band) {
            // column band [cb, ce) of the upper triangle (the host-buffer path computes the matrix band by band
            // while later samples are still in flight over PCIe), rasterised like the full matrix below: strips of
            // about sm_count / width tile rows, so that one wave of CTAs reads ~width + sm_count / width panels
            const int32_t width = ce - cb;
            const int32_t S = std::max(1, (ctx->sm_count + width - 1) / width);
            for (int32_t r0 = 0; r0 < ce; r0 += S)
                for (int32_t tj = std::max(cb, r0); tj < ce; tj++)
                    for (int32_t t = r0; t < std::min(r0 + S, tj + 1); t++)
                        tiles.push_back(TileDesc{t, tj, t * TILE, tj != t ? 1 : 0});
        } else if (full) {
            // L2 rasterisation.  The CTAs of one wave (items w * grid .. (w + 1) * grid - 1, same cost each) stream their
            // operand panels [n, 128] front to back at the same pace, so a panel shared by several tiles of the wave
            // is fetched from DRAM once and found in L2 by the others.  Row-major order puts ~2 tile rows in a wave
            // (grid / 2 + 2 distinct panels); strips of S = floor(sqrt(grid)) tile rows walked column by column make
            // every window of `grid` consecutive tiles an S x S block: ~2 S panels (25 instead of 76 on 148 SMs).
            int32_t S = 1;
            while ((S + 1) * (S + 1) <= ctx->sm_count) S++;
            for (int32_t r0 = 0; r0 < tiles_per_dim; r0 += S)
                for (int32_t tj = r0; tj < tiles_per_dim; tj++)
                    for (int32_t t = r0; t < std::min<int32_t>(r0 + S, tj + 1); t++)
                        tiles.push_back(TileDesc{t, tj, t * TILE, tj != t ? 1 : 0});
        } else {
            // compact row blocks of the listed tile rows (collective multi-GPU form): slot c of the output holds row t
            for (int32_t c = 0; c < n_rows; c++) {
                const int32_t t = tile_rows_h[c];
                for (int32_t tj = t; tj < tiles_per_dim; tj++) tiles.push_back(TileDesc{t, tj, c * TILE, 0});
            }
        }
        int rc = ctx->tables.put(key, tiles.data(), tiles.size() * sizeof(TileDesc), (void **)&d_tiles);
        if (rc) return rc;
    }

    // ---- work plan: whole tiles for the full waves, equal spans of the (tile, basis entry) space for the last wave ----
    const int32_t stages_total = (n + KC - 1) / KC;
    PairWorkPlan plan;
    plan_pair_work(n_tiles, ctx->sm_count, stages_total, flush_stages, getenv("FUF_PW_NO_STREAMK") == nullptr, plan);
    std::string wkey("wk");
    key_append(wkey, n_tiles);
    key_append(wkey, ctx->sm_count);
    key_append(wkey, stages_total);
    key_append(wkey, flush_stages);
    key_append(wkey, plan.n_slots);
    const size_t first_bytes = ((size_t)(plan.grid + 1) * sizeof(int32_t) + 15) & ~(size_t)15;
    const size_t work_bytes = plan.work.size() * sizeof(WorkItem);
    char *d_plan = (char *)ctx->tables.find(wkey);
    if (!d_plan) {
        std::vector<char> blob(first_bytes + work_bytes + plan.shared.size() * sizeof(SharedTile));
        memcpy(blob.data(), plan.cta_first.data(), (size_t)(plan.grid + 1) * sizeof(int32_t));
        memcpy(blob.data() + first_bytes, plan.work.data(), work_bytes);
        if (!plan.shared.empty())
            memcpy(blob.data() + first_bytes + work_bytes, plan.shared.data(), plan.shared.size() * sizeof(SharedTile));
        int rc = ctx->tables.put(wkey, blob.data(), blob.size(), (void **)&d_plan);
        if (rc) return rc;
    }

    double *d_partial = nullptr;
    if (plan.n_slots > 0) {
        int rc = ctx->ws_pair.reserve((size_t)plan.n_slots * TILE * TILE * sizeof(double));
        if (rc) return rc;
        d_partial = (double *)ctx->ws_pair.ptr;
    }

    PairParams p;
    p.tiles = d_tiles;
    p.cta_first = (const int32_t *)d_plan;
    p.work = (const WorkItem *)(d_plan + first_bytes);
    p.stages_total = stages_total;
    p.flush_stages = flush_stages;
    p.shard = shard_cols;
    p.N = N;
    p.D = D;
    p.ldd = ldd;
    p.partial = d_partial;
    p.dep_flags = list ? list->dep_flags : nullptr;
    p.epoch = list ? list->epoch : 0;
    p.stat = list ? list->stat : nullptr;
    p.kappa = (list && list->kappa > 0) ? list->kappa : (mode == FUF_MODE_L2 ? 1.6e13 : 2e6);
    p.flag_list = list ? list->flag_list : nullptr;
    p.list_cap = list ? list->list_cap : 0;
    if (!p.flag_list) p.stat = nullptr;
    p.seg_done = list ? list->seg_done : nullptr;

    const bool packed = !(flags & FUF_PW_SCALAR_MATH);
    void (*kern)(const CUtensorMap, const PairParams) = nullptr;
    if (mode == FUF_MODE_L1)
        kern = packed ? pairwise_tile_kernel<FUF_MODE_L1, true> : pairwise_tile_kernel<FUF_MODE_L1, false>;
    else
        kern = packed ? pairwise_tile_kernel<FUF_MODE_L2, true> : pairwise_tile_kernel<FUF_MODE_L2, false>;
    FUF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    kern<<<plan.grid, NTHREADS, SMEM_BYTES, stream>>>(tmap, p);
    FUF_CUDA(cudaGetLastError());
    ctx->launches += 1;
    if (!plan.shared.empty()) {
        const SharedTile *d_shared = (const SharedTile *)(d_plan + first_bytes + work_bytes);
        const unsigned n_shared = (unsigned)plan.shared.size();
        if (mode == FUF_MODE_L1)
            reduce_partials_kernel<FUF_MODE_L1><<<n_shared, 256, 0, stream>>>(d_tiles, d_shared, d_partial, N, D, ldd, p);
        else
            reduce_partials_kernel<FUF_MODE_L2><<<n_shared, 256, 0, stream>>>(d_tiles, d_shared, d_partial, N, D, ldd, p);
        FUF_CUDA(cudaGetLastError());
        ctx->launches += 1;
    }
    return FUF_OK;
}

int pairwise_f64(fuf_ctx *ctx, const double *P64T, int64_t N, int64_t ldp, int64_t shard, int64_t shard_stride, int mode,
                 const int32_t *tile_rows_h, int32_t n_tile_rows, double *D, int64_t ldd, cudaStream_t stream)
{
    if (N == 0) return FUF_OK;
    const unsigned t = (unsigned)((N + 63) / 64);
    const int32_t *d_rows = nullptr;
    if (tile_rows_h) {
        if (n_tile_rows == 0) return FUF_OK;
        int rc = ctx->tables.get_by_content("rows", tile_rows_h, (size_t)n_tile_rows * sizeof(int32_t), (void **)&d_rows);
        if (rc) return rc;
    }
    dim3 grid(t, tile_rows_h ? 2u * (unsigned)n_tile_rows : t);
    if (mode == FUF_MODE_L1)
        pairwise_f64_kernel<FUF_MODE_L1><<<grid, 256, 0, stream>>>(P64T, ldp, shard, shard_stride, N, ctx->n, d_rows, D, ldd);
    else
        pairwise_f64_kernel<FUF_MODE_L2><<<grid, 256, 0, stream>>>(P64T, ldp, shard, shard_stride, N, ctx->n, d_rows, D, ldd);
    FUF_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return FUF_OK;
}

// Packed row blocks: slot c keeps only the columns its tile row owns, [128, N - 128 t] contiguous at offset[c] — what
// crosses NVLink to the rank that joins the matrix (half the volume of the [128, N] blocks).
struct PackedSlot {
    int32_t t, pad;
    int64_t offset;   // doubles
};

__global__ void __launch_bounds__(256)
pack_tile_rows_kernel(const double *__restrict__ blocks, const PackedSlot *__restrict__ slots, int64_t N, int64_t ldb,
                      double *__restrict__ packed)
{
    const int c = blockIdx.y;
    const PackedSlot slot = slots[c];
    if (slot.t < 0) return;
    const int64_t col0 = (int64_t)slot.t * TILE, gj0 = col0 + (int64_t)blockIdx.x * TILE, ld = N - col0;
    if (gj0 >= N) return;
    const int64_t rows = min((int64_t)TILE, N - col0);
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    for (int r = ly; r < rows; r += 8)
        for (int q = lx; q < TILE; q += 32) {
            const int64_t gj = gj0 + q;
            if (gj < N) packed[slot.offset + r * ld + (gj - col0)] = blocks[((int64_t)c * TILE + r) * ldb + gj];
        }
}

__global__ void __launch_bounds__(256)
assemble_packed_kernel(const double *__restrict__ packed, const PackedSlot *__restrict__ slots, int64_t N,
                       double *__restrict__ D, int64_t ldd)
{
    const int c = blockIdx.y;
    const PackedSlot slot = slots[c];
    const int t = slot.t;
    if (t < 0) return;
    const int tj = t + blockIdx.x;
    const int64_t gi0 = (int64_t)t * TILE, gj0 = (int64_t)tj * TILE, ld = N - gi0;
    if (gj0 >= N) return;
    const double *src = packed + slot.offset;          // src[r * ld + (gj - gi0)]
    __shared__ double tile[32][33];
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    for (int sb = 0; sb < 16; sb++) {
        const int r0 = (sb >> 2) * 32, c0 = (sb & 3) * 32;
        for (int rr = ly; rr < 32; rr += 8) {
            const int64_t gi = gi0 + r0 + rr, gj = gj0 + c0 + lx;
            double v = 0.0;
            if (gi < N && gj < N) {
                if (tj == t && r0 + rr > c0 + lx)       // diagonal tile: lower triangle = mirror of the upper one
                    v = src[(int64_t)(c0 + lx) * ld + r0 + rr];
                else
                    v = src[(int64_t)(r0 + rr) * ld + (gj - gi0)];
                D[gi * ldd + gj] = v;
            }
            tile[rr][lx] = v;
        }
        __syncthreads();
        if (tj != t) {
            for (int cc = ly; cc < 32; cc += 8) {
                const int64_t gj = gj0 + c0 + cc, gi = gi0 + r0 + lx;
                if (gi < N && gj < N) D[gj * ldd + gi] = tile[lx][cc];
            }
        }
        __syncthreads();
    }
}

// Prepares compact row blocks for the direct store into a host matrix shared by all ranks of one node
// (fuf_store_tile_rows_host): the diagonal tile is made symmetric in place, every other upper tile is written
// transposed into a [N - 128 (t + 1)] x 128 panel so that the mirror half leaves as one pitched DMA copy per tile row.
struct PanelSlot {
    int32_t t, pad;
    int64_t offset;   // of the slot's panel in the panel buffer, doubles
};

__global__ void __launch_bounds__(256)
mirror_panels_kernel(double *__restrict__ blocks, const PanelSlot *__restrict__ slots, int64_t N, int64_t ldb,
                     double *__restrict__ panels)
{
    const int c = blockIdx.y;
    const PanelSlot slot = slots[c];
    const int t = slot.t;
    if (t < 0) return;
    const int tj = t + blockIdx.x;
    const int64_t gi0 = (int64_t)t * TILE, gj0 = (int64_t)tj * TILE;
    if (gj0 >= N) return;
    __shared__ double tile[32][33];
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    double *brow = blocks + (int64_t)c * TILE * ldb;
    double *panel = panels + slot.offset;
    for (int sb = 0; sb < 16; sb++) {
        const int r0 = (sb >> 2) * 32, c0 = (sb & 3) * 32;
        if (tj == t && r0 > c0) continue;            // diagonal tile: only sub-tiles on or above the diagonal are sources
        for (int rr = ly; rr < 32; rr += 8) {
            const int64_t gi = gi0 + r0 + rr, gj = gj0 + c0 + lx;
            tile[rr][lx] = (gi < N && gj < N) ? brow[(int64_t)(r0 + rr) * ldb + gj] : 0.0;
        }
        __syncthreads();
        for (int cc = ly; cc < 32; cc += 8) {
            const int64_t gj = gj0 + c0 + cc, gi = gi0 + r0 + lx;   // destination (row gj, column gi)
            if (gi < N && gj < N) {
                if (tj == t) {
                    if (gj > gi) brow[(int64_t)(c0 + cc) * ldb + gi] = tile[lx][cc];
                } else {
                    panel[(gj - gi0 - TILE) * TILE + r0 + lx] = tile[lx][cc];
                }
            }
        }
        __syncthreads();
    }
}

int assemble(fuf_ctx *ctx, const double *blocks, const int32_t *tile_rows_h, int32_t n_blocks, int64_t N, int64_t ldb,
             double *D, int64_t ldd, cudaStream_t stream)
{
    if (n_blocks == 0 || N == 0) return FUF_OK;
    const int64_t tiles_per_dim = (N + TILE - 1) / TILE;
    for (int32_t c = 0; c < n_blocks; c++)
        FUF_REQUIRE(tile_rows_h[c] >= -1 && tile_rows_h[c] < tiles_per_dim, "fuf_assemble: tile row %d out of range",
                    tile_rows_h[c]);
    int32_t *d_rows = nullptr;
    int rc = ctx->tables.get_by_content("rows", tile_rows_h, (size_t)n_blocks * sizeof(int32_t), (void **)&d_rows);
    if (rc) return rc;
    dim3 grid((unsigned)tiles_per_dim, (unsigned)n_blocks);
    assemble_kernel<<<grid, 256, 0, stream>>>(blocks, d_rows, N, ldb, D, ldd);
    FUF_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return FUF_OK;
}

static int upload_packed_slots(fuf_ctx *ctx, const char *who, const int32_t *tile_rows_h, const int64_t *offsets_h,
                               int32_t n_blocks, int64_t N, PackedSlot **d_slots)
{
    const int64_t tiles_per_dim = (N + TILE - 1) / TILE;
    std::vector<PackedSlot> slots((size_t)n_blocks);
    for (int32_t c = 0; c < n_blocks; c++) {
        FUF_REQUIRE(tile_rows_h[c] >= -1 && tile_rows_h[c] < tiles_per_dim, "%s: tile row %d out of range", who,
                    tile_rows_h[c]);
        FUF_REQUIRE(tile_rows_h[c] < 0 || offsets_h[c] >= 0, "%s: negative offset", who);
        slots[c] = PackedSlot{tile_rows_h[c], 0, offsets_h[c]};
    }
    return ctx->tables.get_by_content("packed", slots.data(), slots.size() * sizeof(PackedSlot), (void **)d_slots);
}

int pack_tile_rows(fuf_ctx *ctx, const double *blocks, int64_t ldb, const int32_t *tile_rows_h, const int64_t *offsets_h,
                   int32_t n_blocks, int64_t N, double *packed, cudaStream_t stream)
{
    if (n_blocks == 0 || N == 0) return FUF_OK;
    PackedSlot *d_slots = nullptr;
    int rc = upload_packed_slots(ctx, "fuf_pack_tile_rows", tile_rows_h, offsets_h, n_blocks, N, &d_slots);
    if (rc) return rc;
    dim3 grid((unsigned)((N + TILE - 1) / TILE), (unsigned)n_blocks);
    pack_tile_rows_kernel<<<grid, 256, 0, stream>>>(blocks, d_slots, N, ldb, packed);
    FUF_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return FUF_OK;
}

int assemble_packed(fuf_ctx *ctx, const double *packed, const int32_t *tile_rows_h, const int64_t *offsets_h,
                    int32_t n_blocks, int64_t N, double *D, int64_t ldd, cudaStream_t stream)
{
    if (n_blocks == 0 || N == 0) return FUF_OK;
    PackedSlot *d_slots = nullptr;
    int rc = upload_packed_slots(ctx, "fuf_assemble_packed", tile_rows_h, offsets_h, n_blocks, N, &d_slots);
    if (rc) return rc;
    dim3 grid((unsigned)((N + TILE - 1) / TILE), (unsigned)n_blocks);
    assemble_packed_kernel<<<grid, 256, 0, stream>>>(packed, d_slots, N, D, ldd);
    FUF_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return FUF_OK;
}

int store_tile_rows_host(fuf_ctx *ctx, double *blocks, const int32_t *tile_rows_h, int32_t n_blocks, int64_t N,
                         int64_t ldb, double *D_h, int64_t ldd, cudaStream_t stream)
{
    if (n_blocks == 0 || N == 0) return FUF_OK;
    const int64_t tiles_per_dim = (N + TILE - 1) / TILE;
    std::vector<PanelSlot> slots((size_t)n_blocks);
    int64_t total = 0;
    for (int32_t c = 0; c < n_blocks; c++) {
        const int32_t t = tile_rows_h[c];
        FUF_REQUIRE(t >= -1 && t < tiles_per_dim, "fuf_store_tile_rows_host: tile row %d out of range", t);
        slots[c] = PanelSlot{t, 0, total};
        if (t >= 0) total += std::max<int64_t>(0, N - (int64_t)(t + 1) * TILE) * TILE;
    }
    PanelSlot *d_slots = nullptr;
    int rc = ctx->tables.get_by_content("panel", slots.data(), slots.size() * sizeof(PanelSlot), (void **)&d_slots);
    if (rc) return rc;
    rc = ctx->ws_panel.reserve((size_t)std::max<int64_t>(total, 1) * sizeof(double));
    if (rc) return rc;
    double *panels = (double *)ctx->ws_panel.ptr;
    dim3 grid((unsigned)tiles_per_dim, (unsigned)n_blocks);
    mirror_panels_kernel<<<grid, 256, 0, stream>>>(blocks, d_slots, N, ldb, panels);
    FUF_CUDA(cudaGetLastError());
    ctx->launches += 1;
    const char *part_env = getenv("FUF_STORE_PART");      // diagnosis only: 1 = upper part only, 2 = mirror only
    const int part = part_env ? atoi(part_env) : 0;
    for (int32_t c = 0; c < n_blocks; c++) {
        const int64_t t = slots[c].t;
        if (t < 0) continue;
        const int64_t row0 = t * TILE, rows = std::min<int64_t>(TILE, N - row0), below = N - row0 - rows;
        // rows [row0, row0 + rows), columns >= row0 (the diagonal tile is symmetric by now)
        if (part != 2)
        FUF_RC(copy2d_async(D_h + row0 * ldd + row0, (size_t)ldd * sizeof(double),
                                   blocks + (int64_t)c * TILE * ldb + row0, (size_t)ldb * sizeof(double),
                                   (size_t)(N - row0) * sizeof(double), (size_t)rows, stream));
        // the mirror: rows below the diagonal tile, columns [row0, row0 + rows)
        if (below > 0 && part != 1)
            FUF_RC(copy2d_async(D_h + (row0 + rows) * ldd + row0, (size_t)ldd * sizeof(double),
                                       panels + slots[c].offset, (size_t)TILE * sizeof(double),
                                       (size_t)rows * sizeof(double), (size_t)below, stream));
    }
    return FUF_OK;
}

}  // namespace fuf

using namespace fuf;

extern "C" int fuf_pairwise(fuf_ctx *ctx, const float *PT, int64_t N, int64_t ldp, int64_t shard, int64_t shard_stride,
                            int mode, const int32_t *tile_rows_h, int32_t n_tile_rows, double *D, int64_t ldd,
                            int flags, void *stream)
{
    FUF_REQUIRE(ctx && PT && D, "fuf_pairwise: null argument");
    FUF_REQUIRE(mode == FUF_MODE_L1 || mode == FUF_MODE_L2, "fuf_pairwise: bad mode %d", mode);
    FUF_REQUIRE(N >= 0 && ldd >= N, "fuf_pairwise: bad shape N=%lld ldd=%lld", (long long)N, (long long)ldd);
    FUF_REQUIRE(tile_rows_h != nullptr || n_tile_rows == 0, "fuf_pairwise: n_tile_rows without tile_rows_h");
    FUF_CUDA(cudaSetDevice(ctx->device));
    return pairwise_f32(ctx, PT, N, ldp, shard, shard_stride, mode, tile_rows_h, n_tile_rows, D, ldd, flags,
                        (cudaStream_t)stream, 0, 0, 0, 0, nullptr);
}

extern "C" int fuf_pairwise_tiles(fuf_ctx *ctx, const float *PT, int64_t N, int64_t shard, int64_t shard_stride, int mode,
                                  const int32_t *tiles_h, int32_t n_tiles, const uint32_t *dep_flags, uint32_t epoch,
                                  double *D, int64_t ldd, const double *err_stat, double kappa, int64_t *flag_list,
                                  int64_t list_cap, const int32_t *seg_h, uint32_t *seg_done, void *stream)
{
    FUF_REQUIRE(ctx && PT && D && (tiles_h || n_tiles == 0), "fuf_pairwise_tiles: null argument");
    FUF_REQUIRE(mode == FUF_MODE_L1 || mode == FUF_MODE_L2, "fuf_pairwise_tiles: bad mode %d", mode);
    FUF_REQUIRE(N >= 0 && ldd >= N && n_tiles >= 0, "fuf_pairwise_tiles: bad shape N=%lld ldd=%lld", (long long)N, (long long)ldd);
    FUF_REQUIRE(shard > 0, "fuf_pairwise_tiles: the operands are the sharded layout [G][n][shard]");
    FUF_REQUIRE(!flag_list || (err_stat && list_cap >= 0), "fuf_pairwise_tiles: flag_list needs err_stat");
    if (n_tiles == 0) return FUF_OK;
    FUF_CUDA(cudaSetDevice(ctx->device));
    PairTileList list;
    list.tiles_h = tiles_h;
    list.n_tiles = n_tiles;
    list.dep_flags = dep_flags;
    list.epoch = epoch;
    list.stat = err_stat;
    list.kappa = kappa;
    list.flag_list = (long long *)flag_list;
    list.list_cap = list_cap;
    list.seg_h = seg_done ? seg_h : nullptr;
    list.seg_done = seg_h ? seg_done : nullptr;
    return pairwise_f32(ctx, PT, N, 0, shard, shard_stride, mode, nullptr, 0, D, ldd, 0, (cudaStream_t)stream, 0, 0, 0, 0,
                        &list);
}

extern "C" int fuf_pairwise_f64(fuf_ctx *ctx, const double *P64T, int64_t N, int64_t ldp, int64_t shard,
                                int64_t shard_stride, int mode, const int32_t *tile_rows_h, int32_t n_tile_rows,
                                double *D, int64_t ldd, void *stream)
{
    FUF_REQUIRE(ctx && P64T && D, "fuf_pairwise_f64: null argument");
    FUF_REQUIRE(mode == FUF_MODE_L1 || mode == FUF_MODE_L2, "fuf_pairwise_f64: bad mode %d", mode);
    FUF_REQUIRE(N >= 0 && ldd >= N, "fuf_pairwise_f64: bad shape N=%lld ldd=%lld", (long long)N, (long long)ldd);
    if (shard == 0)
        FUF_REQUIRE(ldp >= N, "fuf_pairwise_f64: ldp=%lld < N=%lld", (long long)ldp, (long long)N);
    else
        FUF_REQUIRE(shard > 0 && shard_stride >= shard * (int64_t)ctx->n, "fuf_pairwise_f64: bad shard layout");
    FUF_REQUIRE(tile_rows_h != nullptr || n_tile_rows == 0, "fuf_pairwise_f64: n_tile_rows without tile_rows_h");
    const int64_t tiles_per_dim = (N + TILE - 1) / TILE;
    for (int32_t c = 0; c < n_tile_rows; c++)
        FUF_REQUIRE(tile_rows_h[c] >= -1 && tile_rows_h[c] < tiles_per_dim, "fuf_pairwise_f64: tile row %d out of range",
                    tile_rows_h[c]);
    FUF_CUDA(cudaSetDevice(ctx->device));
    return pairwise_f64(ctx, P64T, N, ldp, shard, shard_stride, mode, tile_rows_h, n_tile_rows, D, ldd,
                        (cudaStream_t)stream);
}

extern "C" int fuf_assemble(fuf_ctx *ctx, const double *blocks, const int32_t *tile_rows_h, int32_t n_blocks, int64_t N,
                            int64_t ldb, double *D, int64_t ldd, void *stream)
{
    FUF_REQUIRE(ctx && blocks && tile_rows_h && D, "fuf_assemble: null argument");
    FUF_REQUIRE(N >= 0 && ldd >= N && ldb >= N, "fuf_assemble: bad shape");
    FUF_CUDA(cudaSetDevice(ctx->device));
    return assemble(ctx, blocks, tile_rows_h, n_blocks, N, ldb, D, ldd, (cudaStream_t)stream);
}

extern "C" int fuf_store_tile_rows_host(fuf_ctx *ctx, double *blocks, const int32_t *tile_rows_h, int32_t n_blocks,
                                        int64_t N, int64_t ldb, double *D_h, int64_t ldd, void *stream)
{
    FUF_REQUIRE(ctx && blocks && tile_rows_h && D_h, "fuf_store_tile_rows_host: null argument");
    FUF_REQUIRE(N >= 0 && ldd >= N && ldb >= N, "fuf_store_tile_rows_host: bad shape");
    FUF_CUDA(cudaSetDevice(ctx->device));
    return store_tile_rows_host(ctx, blocks, tile_rows_h, n_blocks, N, ldb, D_h, ldd, (cudaStream_t)stream);
}

extern "C" int fuf_pack_tile_rows(fuf_ctx *ctx, const double *blocks, int64_t ldb, const int32_t *tile_rows_h,
                                  const int64_t *offsets_h, int32_t n_blocks, int64_t N, double *packed, void *stream)
{
    FUF_REQUIRE(ctx && blocks && tile_rows_h && offsets_h && packed, "fuf_pack_tile_rows: null argument");
    FUF_REQUIRE(N >= 0 && ldb >= N, "fuf_pack_tile_rows: bad shape");
    FUF_CUDA(cudaSetDevice(ctx->device));
    return pack_tile_rows(ctx, blocks, ldb, tile_rows_h, offsets_h, n_blocks, N, packed, (cudaStream_t)stream);
}

extern "C" int fuf_assemble_packed(fuf_ctx *ctx, const double *packed, const int32_t *tile_rows_h,
                                   const int64_t *offsets_h, int32_t n_blocks, int64_t N, double *D, int64_t ldd,
                                   void *stream)
{
    FUF_REQUIRE(ctx && packed && tile_rows_h && offsets_h && D, "fuf_assemble_packed: null argument");
    FUF_REQUIRE(N >= 0 && ldd >= N, "fuf_assemble_packed: bad shape");
    FUF_CUDA(cudaSetDevice(ctx->device));
    return assemble_packed(ctx, packed, tile_rows_h, offsets_h, n_blocks, N, D, ldd, (cudaStream_t)stream);
}
