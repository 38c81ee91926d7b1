// gram.cu — --L2 all-pairs distances as a length-weighted Gram matrix on the tcgen05 tensor cores.
//
// Replaces, for is_L2, the pair loop of pairwise_computation (src/algorithms/emd_unifrac.py:49-54) with
// get_L2 = sum_k (P_i[k] - P_j[k])^2 (src/objects/profile_vector.py:17-19) rewritten as
//     D[i,j] = q_i + q_j - 2 G[i,j],     G = P' P'^T,     q_i = sum_k P'_i[k]^2,
// where P' are the centred fp32 pushed profiles (sqrt(L) weights are already inside P, see pushup.cu; centring
// is what keeps q_i + q_j - 2G from cancelling: for centred profiles G is small against q).
//
// K3c gram_mean_kernel   column (node) means of the operand over the samples: the Gram path centres EVERY node,
//                        leaves included, so that the products a_k b_k have no common sign and the fp32 partial
//                        sums in tensor memory stay small against q (distances do not depend on the centre).
// K3p gram_prep_kernel   a = fl32(PT - mean); splits it into two TF32 numbers, a = hi + lo (hi = tf32(a),
//                        lo = tf32(a-hi)); accumulates q_i and the extra rounding statistic in double.  HBM-bound.
// K3  gram_tile_kernel   persistent, warp-specialised, one CTA per SM, 128x128 tiles of the upper triangle:
//       warp 0   TMA producer: per stage (32 nodes) 16 boxes of 32 samples x 32 nodes, SWIZZLE_128B_ATOM_32B, into a
//                3-stage 64 KB ring (A_hi, A_lo, B_hi, B_lo), full/empty mbarriers;
//       warp 1   MMA issuer: one elected lane issues tcgen05.mma.cta_group::1.kind::tf32, M = N = 128, K = 8,
//                MN-major shared-memory descriptors; three products per k-block (hi*hi + hi*lo + lo*hi =
//                "3xTF32", products exact in fp32, dropped lo*lo ~ 2^-22); accumulators in TMEM (2 x 128 columns,
//                double-buffered), tcgen05.commit releases smem stages and hands accumulators to the epilogue;
//       warps 4-11 epilogue (two warpgroups, setmaxnreg 232): every 256 nodes the fp32 accumulator tile is read
//                with tcgen05.ld (32x32b.x32), widened and added to float64 accumulators held in REGISTERS
//                (each thread owns one row x 64 columns = 64 doubles), so no fp32 sum ever sees more than 96
//                tensor-core additions; at the end of the tile the thread writes q_i + q_j - 2G (512 B contiguous).
// K3m mirror_kernel      lower triangle := transpose of the upper one, zero diagonal.
// Hybrid split.  The few nodes near the root carry large, strongly correlated entries: for them the products
// a_k b_k are of the order of q while their contribution to D is small, and the 2^-22 relative error of a 3xTF32
// product shows (measured: 1.3e-6 on 1.8 %% of the nodes, 1.2e-7 on the rest).  The host therefore picks the
// GRAM_DIRECT_ROWS node columns of largest energy sum_s a_k[s]^2; they are excluded from the Gram operands and go
// through the direct fp32 tile kernel (pairwise.cu, K2) on a gathered copy — 1.7 %% of its full cost — whose
// result the Gram epilogue adds.
// The pairs the Gram form cannot certify (D small against q_i + q_j: near-duplicates) are recomputed by the
// float64 refinement pass (refine.cu) with the linear criterion D < tau (q_i + q_j).
//
// Tensor-pipe accounting: executed MMA flops = 3 passes x 2 x 128 x 128 x n per tile; algorithmic flops = 2 n per pair.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace fuf {

namespace gram {

constexpr int TILE = FUF_TILE;              // 128 = UMMA M = UMMA N
constexpr int KC = 32;                      // nodes per pipeline stage
constexpr int KB = 8;                       // nodes per tcgen05.mma (kind::tf32: 32 bytes of K)
constexpr int STAGES = 3;
constexpr int FLUSH_STAGES = 8;             // fp32 accumulation length in TMEM: 8 stages = 256 nodes (96 MMA adds)
constexpr int BOX_W = 32;                   // samples per TMA box (128 bytes: the swizzle span)
constexpr uint32_t BOX_BYTES = BOX_W * KC * 4;              // 4 KB
constexpr uint32_t OPER_BYTES = (TILE / BOX_W) * BOX_BYTES;  // 16 KB: one operand tile (128 samples x 32 nodes)
constexpr uint32_t STAGE_BYTES = 4 * OPER_BYTES;             // A_hi, A_lo, B_hi, B_lo
constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers + tmem ptr */;
constexpr int NTHREADS = 384;               // warpgroup 0: warp 0 TMA, warp 1 MMA (2 idle); warpgroups 1-2: epilogue
constexpr int NEPI_WARPS = 8;
constexpr uint32_t TMEM_COLS = 256;         // two 128-column fp32 accumulators
constexpr int GRAM_DIRECT_ROWS = 512;       // node columns routed through the direct fp32 kernel

struct Tile {
    int32_t ti, tj;
    int32_t out_row0;  // first row of this tile row in the output buffer
    int32_t pad;
};

// Where sample s of a node-major operand lives: element (k, s) at base[(s / S) * SS + k * LD + s % S].
// Single block: S = padded sample count, SS = 0, LD = ldp.  All-gathered shards [G][n][S]: SS = shard stride, LD = S.
struct Lay {
    int64_t S, SS, LD;
    __host__ __device__ int64_t at(int64_t k, int64_t s) const { return (s / S) * SS + k * LD + s % S; }
};

struct Params {
    const Tile *tiles;
    int32_t n_tiles;
    int32_t stages_total;  // ceil(n / KC)
    int32_t flush_stages;  // fp32 accumulation length in tensor memory, in stages
    int32_t S;             // samples per operand shard (a multiple of 128)
    int64_t N;
    const double *q;
    double *D;
    int64_t ldd;
};

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
// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor).  MN-major 32-bit operands have exactly one legal
// layout: SWIZZLE_128B_BASE32B = rows of 128 B (32 samples) along MN, atoms of 4 node-rows (512 B), the 32-byte chunk
// index XORed with (row % 4) — what the TMA mode CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B writes.
//   [0,14) start >> 4 | [16,30) LBO >> 4 (between 32-sample blocks) | [32,46) SBO >> 4 (between 4-node atoms)
//   | [46,48) version = 1 | [61,64) layout = 1 (SWIZZLE_128B_BASE32B)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr)
{
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)(BOX_BYTES >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) |
           (1ull << 46) | (1ull << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c = F32, a = b = TF32, both MN-major, N = 128, M = 128.
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(TILE >> 3) << 17) |
                           ((uint32_t)(TILE >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(NTHREADS, 1)
gram_tile_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo, const Params p)
{
    extern __shared__ unsigned char smem_raw[];
    // swizzle atoms must sit on their own size boundaries; 1,024 bytes covers every mode
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + STAGES * STAGE_BYTES;
    const uint32_t full0 = bars, empty0 = bars + 8 * STAGES, tfull0 = bars + 16 * STAGES, tempty0 = tfull0 + 16;
    const uint32_t tmem_slot = tempty0 + 16;
    volatile uint32_t *tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(full0 + 8 * s, 1);   // producer's arrive.expect_tx (+ the TMA bytes)
            mbar_init(empty0 + 8 * s, 1);  // tcgen05.commit of the MMA warp
        }
        for (int b = 0; b < 2; b++) {
            mbar_init(tfull0 + 8 * b, 1);   // tcgen05.commit after the last MMA of a flush chunk
            mbar_init(tempty0 + 8 * b, NEPI_WARPS);  // one arrive per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {  // the MMA warp owns the tensor-memory allocation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int FS = p.flush_stages;
    const int n_chunks = (p.stages_total + FS - 1) / FS;

    if (warp < 4) {
        // warpgroup 0 gives its registers to the epilogue warpgroups (the register re-allocation has to sit inside
        // the role branches: ptxas budgets each branch by the setmaxnreg that dominates it)
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_hi) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_lo) : "memory");
            uint32_t it = 0;
            for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
                const Tile tl = p.tiles[t];
                const int ia = (tl.ti * TILE) % p.S, ga = (tl.ti * TILE) / p.S;
                const int ib = (tl.tj * TILE) % p.S, gb = (tl.tj * TILE) / p.S;
                for (int st = 0; st < p.stages_total; ++st, ++it) {
                    const uint32_t slot = it % STAGES, phase = (it / STAGES) & 1u;
                    mbar_wait(empty0 + 8 * slot, phase ^ 1u);
                    const uint32_t full = full0 + 8 * slot;
                    mbar_expect_tx(full, STAGE_BYTES);
                    const uint32_t dst = base + slot * STAGE_BYTES;
#pragma unroll
                    for (int b = 0; b < TILE / BOX_W; b++) {
                        tma_load_3d(dst + 0 * OPER_BYTES + b * BOX_BYTES, &map_hi, full, ia + b * BOX_W, st * KC, ga);
                        tma_load_3d(dst + 1 * OPER_BYTES + b * BOX_BYTES, &map_lo, full, ia + b * BOX_W, st * KC, ga);
                        tma_load_3d(dst + 2 * OPER_BYTES + b * BOX_BYTES, &map_hi, full, ib + b * BOX_W, st * KC, gb);
                        tma_load_3d(dst + 3 * OPER_BYTES + b * BOX_BYTES, &map_lo, full, ib + b * BOX_W, st * KC, gb);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one elected lane) =====================
        uint32_t it = 0, chunk_it = 0;
        for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
            for (int c = 0; c < n_chunks; ++c, ++chunk_it) {
                const uint32_t buf = chunk_it & 1u, bphase = (chunk_it >> 1) & 1u;
                mbar_wait(tempty0 + 8 * buf, bphase ^ 1u);  // the epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + buf * TILE;
                const int s_end = min(p.stages_total, (c + 1) * FS);
                for (int st = c * FS; st < s_end; ++st, ++it) {
                    const uint32_t slot = it % STAGES, phase = (it / STAGES) & 1u;
                    mbar_wait(full0 + 8 * slot, phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (lane == 0) {
                        const uint32_t a_hi = base + slot * STAGE_BYTES, a_lo = a_hi + OPER_BYTES;
                        const uint32_t b_hi = a_hi + 2 * OPER_BYTES, b_lo = a_hi + 3 * OPER_BYTES;
#pragma unroll
                        for (int kb = 0; kb < KC / KB; kb++) {
                            const uint32_t off = kb * 1024;  // 8 nodes x 128 B
                            const uint32_t first = (st == c * FS && kb == 0) ? 0u : 1u;
                            umma_tf32(d_tmem, umma_desc(a_lo + off), umma_desc(b_hi + off), first);
                            umma_tf32(d_tmem, umma_desc(a_hi + off), umma_desc(b_lo + off), 1u);
                            umma_tf32(d_tmem, umma_desc(a_hi + off), umma_desc(b_hi + off), 1u);
                        }
                        umma_commit(empty0 + 8 * slot);  // frees the smem stage when these MMAs have read it
                    }
                    __syncwarp();
                }
                if (lane == 0) umma_commit(tfull0 + 8 * buf);  // accumulator complete
                __syncwarp();
            }
        }
    }
    } else {
        // ===================== epilogue warps: TMEM -> float64 accumulators in registers =====================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        const int quarter = warp & 3;         // tcgen05.ld: a warp may touch TMEM lanes 32 (warp % 4) .. + 31
        const int half = (warp >> 2) - 1;     // columns [64 half, 64 half + 64)
        const int row = quarter * 32 + lane;
        uint32_t chunk_it = 0;
        for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
            const Tile tl = p.tiles[t];
            double acc[64];
#pragma unroll
            for (int e = 0; e < 64; e++) acc[e] = 0.0;
            for (int c = 0; c < n_chunks; ++c, ++chunk_it) {
                const uint32_t buf = chunk_it & 1u, bphase = (chunk_it >> 1) & 1u;
                mbar_wait(tfull0 + 8 * buf, bphase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int cb = 0; cb < 2; cb++) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * TILE + half * 64 + cb * 32, v);
#pragma unroll
                    for (int e = 0; e < 32; e++) acc[cb * 32 + e] += (double)__uint_as_float(v[e]);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
            }
            const int64_t gi = (int64_t)tl.ti * TILE + row, gj0 = (int64_t)tl.tj * TILE + half * 64;
            if (gi < p.N) {
                const double qi = p.q[gi];
                double *drow = p.D + ((int64_t)tl.out_row0 + row) * p.ldd + gj0;
#pragma unroll
                for (int e = 0; e < 64; e++) {
                    const int64_t gj = gj0 + e;
                    if (gj < p.N) drow[e] = (gi == gj) ? 0.0 : drow[e] + ((qi + p.q[gj]) - 2.0 * acc[e]);
                }
            }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// mean[k] = (1/N) sum_s PT[k][s] and energy[k] = sum_s (PT[k][s] - mean[k])^2, in double, one warp per node row.
__global__ void __launch_bounds__(256)
gram_mean_kernel(const float *__restrict__ PT, const Lay lay, int64_t N, int32_t n, float *__restrict__ mean,
                 float *__restrict__ energy)
{
    const int32_t k = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (k >= n) return;
    const int lane = threadIdx.x & 31;
    double acc = 0.0;
    for (int64_t s = lane; s < N; s += 32) acc += (double)PT[lay.at(k, s)];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    const float m = (float)(acc / (double)N);
    double e = 0.0;
    for (int64_t s = lane; s < N; s += 32) {
        const double d = (double)PT[lay.at(k, s)] - (double)m;
        e = fma(d, d, e);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if (lane == 0) {
        mean[k] = m;
        energy[k] = (float)e;
    }
}

// a = fl32(PT - mean) = hi + lo with hi, lo representable in TF32 (10 explicit mantissa bits);
// q[block][s] = sum_k a^2 (Gram columns only) and r2[block][s] = sum_k (rounding error of a)^2 over the block's nodes.
__global__ void __launch_bounds__(256)
gram_prep_kernel(const float *__restrict__ PT, const Lay lay, const Lay ws, const Lay sub, int64_t n_cols, int64_t N,
                 int32_t n, int32_t k_per_block, const float *__restrict__ mean, const int32_t *__restrict__ direct_slot,
                 float *__restrict__ HI, float *__restrict__ LO, float *__restrict__ SUB, double *__restrict__ q,
                 double *__restrict__ r2)
{
    // s runs over every sample slot of the workspace planes (padding slots are written as zeros)
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_cols) return;
    const int64_t ldp = n_cols;  // row length of the partial-statistics arrays
    const int32_t k0 = blockIdx.y * k_per_block, k1 = min(n, k0 + k_per_block);
    double acc = 0.0, err = 0.0;
    for (int32_t k = k0; k < k1; k++) {
        float a = 0.f;
        if (s < N) {
            const double c = (double)PT[lay.at(k, s)] - (double)mean[k];
            a = (float)c;
            const double e = c - (double)a;
            err = fma(e, e, err);
        }
        const int32_t slot = direct_slot[k];
        if (slot >= 0) {  // this node column goes through the direct fp32 kernel
            SUB[sub.at(slot, s)] = a;
            HI[ws.at(k, s)] = 0.f;
            LO[ws.at(k, s)] = 0.f;
            continue;
        }
        uint32_t hb, lb;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(a));
        const float hi = __uint_as_float(hb);
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(a - hi));
        HI[ws.at(k, s)] = hi;
        LO[ws.at(k, s)] = __uint_as_float(lb);
        acc = fma((double)a, (double)a, acc);
    }
    // per-(node block) partials, summed in fixed order by gram_stat_kernel: bit-reproducible norms
    q[(int64_t)blockIdx.y * ldp + s] = acc;
    r2[(int64_t)blockIdx.y * ldp + s] = err;
}

// q[s] = sum of the partial norms in block order; err_out[s] = (sqrt(err_in[s]) + sqrt(sum r2))^2: the rounding
// statistics of two successive roundings add in norm.
__global__ void gram_stat_kernel(const double *__restrict__ qpart, const double *__restrict__ r2part, int32_t n_parts,
                                 int64_t ldp, const double *__restrict__ err_in, int64_t N, double *__restrict__ q,
                                 double *__restrict__ err_out)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    double qs = 0.0, rs = 0.0;
    for (int32_t b = 0; b < n_parts; b++) {
        qs += qpart[(int64_t)b * ldp + s];
        rs += r2part[(int64_t)b * ldp + s];
    }
    q[s] = qs;
    if (err_out) {
        const double a = sqrt(err_in ? err_in[s] : 0.0) + sqrt(rs);
        err_out[s] = a * a;
    }
}

// lower triangle := transpose of the upper triangle (tiles tj > ti and the diagonal tiles), 32x32 sub-blocks.
__global__ void __launch_bounds__(256)
mirror_kernel(double *__restrict__ D, int64_t N, int64_t ldd)
{
    const int bi = blockIdx.y, bj = blockIdx.x;  // 32x32 block coordinates, bj >= bi
    if (bj < bi) return;
    __shared__ double tile[32][33];
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int64_t i0 = (int64_t)bi * 32, j0 = (int64_t)bj * 32;
    for (int r = ly; r < 32; r += 8) {
        const int64_t gi = i0 + r, gj = j0 + lx;
        tile[r][lx] = (gi < N && gj < N) ? D[gi * ldd + gj] : 0.0;
    }
    __syncthreads();
    for (int r = ly; r < 32; r += 8) {
        const int64_t gj = j0 + r, gi = i0 + lx;  // writes D[gj][gi] = upper[gi][gj]
        if (gi < N && gj < N && gj > gi) D[gj * ldd + gi] = tile[lx][r];
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn()
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

static int make_map(CUtensorMap *map, const float *ptr, int64_t N, int32_t n, const Lay &ws)
{
    EncodeTiledFn encode = get_encode_fn();
    if (!encode) {
        set_error("fuf_pairwise_gram: cuTensorMapEncodeTiled is not available from this driver");
        return FUF_ECUDA;
    }
    // dims (sample in shard, node, shard); samples beyond N / nodes beyond n read as zero (the planes are zero there)
    const int64_t n_shards = ws.SS ? (N + ws.S - 1) / ws.S : 1;
    const cuuint64_t gdim[3] = {(cuuint64_t)(ws.SS ? ws.S : N), (cuuint64_t)n, (cuuint64_t)n_shards};
    const cuuint64_t gstride[2] = {(cuuint64_t)ws.LD * sizeof(float),
                                   (cuuint64_t)(ws.SS ? ws.SS : ws.LD * (int64_t)n) * sizeof(float)};
    const cuuint32_t box[3] = {BOX_W, KC, 1}, estr[3] = {1, 1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)ptr, gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("fuf_pairwise_gram: cuTensorMapEncodeTiled failed with CUresult %d (N=%lld n=%d S=%lld)", (int)r,
                  (long long)N, n, (long long)ws.S);
        return FUF_ECUDA;
    }
    return FUF_OK;
}

}  // namespace gram

int pairwise_gram(fuf_ctx *ctx, const float *PT, int64_t N, int64_t ldp, int64_t shard, int64_t shard_stride,
                  const double *err_in, double *q, double *err_out, const int32_t *tile_rows_h, int32_t n_tile_rows,
                  double *D, int64_t ldd, cudaStream_t stream)
{
    using namespace gram;
    if (N == 0) return FUF_OK;
    const int32_t n = ctx->n;
    const int64_t tiles_per_dim = (N + TILE - 1) / TILE;
    FUF_REQUIRE(((uintptr_t)PT & 15) == 0, "fuf_pairwise_gram: PT must be 16-byte aligned");
    Lay lay, ws, sub;
    int32_t n_direct = GRAM_DIRECT_ROWS;
    if (getenv("FUF_GRAM_DIRECT")) n_direct = std::max(16, atoi(getenv("FUF_GRAM_DIRECT")));
    n_direct = std::min<int32_t>(n, n_direct);
    const int32_t n_direct_pad = (n_direct + 15) / 16 * 16;
    int64_t n_cols;  // sample slots of the workspace planes
    if (shard == 0) {
        FUF_REQUIRE(ldp % 4 == 0 && ldp >= N, "fuf_pairwise_gram: ldp=%lld must be a multiple of 4 and >= N=%lld",
                    (long long)ldp, (long long)N);
        n_cols = tiles_per_dim * TILE;
        lay = Lay{ldp, 0, ldp};
        ws = Lay{n_cols, 0, n_cols};
        sub = Lay{n_cols, 0, n_cols};
    } else {
        FUF_REQUIRE(shard % TILE == 0 && shard > 0 && shard_stride >= shard * (int64_t)n,
                    "fuf_pairwise_gram: bad shard layout (shard=%lld stride=%lld)", (long long)shard,
                    (long long)shard_stride);
        const int64_t n_shards = (N + shard - 1) / shard;
        n_cols = n_shards * shard;
        lay = Lay{shard, shard_stride, shard};
        ws = Lay{shard, (int64_t)n * shard, shard};
        sub = Lay{shard, (int64_t)n_direct_pad * shard, shard};
    }
    // workspace: HI | LO (n x n_cols floats each) | SUB (n_direct x n_cols) | mean, energy (n floats) | slot (n ints)
    //            | partial norms and rounding statistics | tile list
    auto up = [](size_t b) { return (b + 1023) & ~(size_t)1023; };
    const size_t plane = up((size_t)n * n_cols * sizeof(float)), sub_bytes = up((size_t)n_direct_pad * n_cols * sizeof(float));
    const int32_t k_per_block = 512, n_parts = (n + k_per_block - 1) / k_per_block;
    const size_t vec_bytes = up((size_t)n * sizeof(float)), part_bytes = up((size_t)n_parts * n_cols * sizeof(double));
    std::vector<Tile> tiles;
    const bool full = (tile_rows_h == nullptr);
    const int32_t n_rows = full ? (int32_t)tiles_per_dim : n_tile_rows;
    for (int32_t c = 0; c < n_rows; c++) {
        const int32_t t = full ? c : tile_rows_h[c];
        FUF_REQUIRE(t >= 0 && t < tiles_per_dim, "fuf_pairwise_gram: tile row %d out of range [0,%lld)", t,
                    (long long)tiles_per_dim);
        for (int32_t tj = t; tj < tiles_per_dim; tj++) tiles.push_back(Tile{t, tj, full ? t * TILE : c * TILE, 0});
    }
    if (tiles.empty()) return FUF_OK;
    const size_t tiles_bytes = tiles.size() * sizeof(Tile);
    int rc = ctx->ws_gram.reserve(2 * plane + sub_bytes + 3 * vec_bytes + 2 * part_bytes + tiles_bytes);
    if (rc) return rc;
    char *wsp = (char *)ctx->ws_gram.ptr;
    float *HI = (float *)wsp, *LO = (float *)(wsp + plane), *SUB = (float *)(wsp + 2 * plane);
    float *mean = (float *)(wsp + 2 * plane + sub_bytes), *energy = (float *)((char *)mean + vec_bytes);
    int32_t *d_slot = (int32_t *)((char *)mean + 2 * vec_bytes);
    double *qpart = (double *)((char *)mean + 3 * vec_bytes), *r2part = (double *)((char *)qpart + part_bytes);
    Tile *d_tiles = (Tile *)((char *)r2part + part_bytes);
    FUF_CUDA(cudaMemcpyAsync(d_tiles, tiles.data(), tiles_bytes, cudaMemcpyHostToDevice, stream));
    FUF_CUDA(cudaMemsetAsync(SUB, 0, sub_bytes, stream));
    gram_mean_kernel<<<(unsigned)((n + 7) / 8), 256, 0, stream>>>(PT, lay, N, n, mean, energy);
    FUF_CUDA(cudaGetLastError());
    // the n_direct node columns of largest energy -> direct kernel (host selection: n floats, once per call)
    std::vector<float> h_energy(n);
    FUF_CUDA(cudaMemcpyAsync(h_energy.data(), energy, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, stream));
    FUF_CUDA(cudaStreamSynchronize(stream));  // also: `tiles` is pageable and dies with this call
    std::vector<int32_t> order(n), h_slot(n, -1);
    for (int32_t k = 0; k < n; k++) order[k] = k;
    std::nth_element(order.begin(), order.begin() + (n_direct - 1), order.end(), [&](int32_t a, int32_t b) {
        return h_energy[a] > h_energy[b] || (h_energy[a] == h_energy[b] && a < b);
    });
    std::sort(order.begin(), order.begin() + n_direct);
    for (int32_t i = 0; i < n_direct; i++) h_slot[order[i]] = i;
    FUF_CUDA(cudaMemcpyAsync(d_slot, h_slot.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
    FUF_CUDA(cudaStreamSynchronize(stream));
    {
        dim3 g((unsigned)((n_cols + 255) / 256), (unsigned)n_parts);
        gram_prep_kernel<<<g, 256, 0, stream>>>(PT, lay, ws, sub, n_cols, N, n, k_per_block, mean, d_slot, HI, LO, SUB,
                                                qpart, r2part);
        FUF_CUDA(cudaGetLastError());
        gram_stat_kernel<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(qpart, r2part, n_parts, n_cols, err_in, N, q, err_out);
        FUF_CUDA(cudaGetLastError());
    }
    // direct part first: D = sum over the selected node columns of (a_i - a_j)^2
    // (fp32 partial sums over 16 entries only: these columns carry most of the magnitude)
    rc = pairwise_f32(ctx, SUB, N, n_cols, shard ? shard : 0, shard ? sub.SS : 0, FUF_MODE_L2, tile_rows_h, n_tile_rows, D,
                      ldd, 0, stream, nullptr, 0, 0, 0, n_direct_pad, 1);
    if (rc) return rc;
    ctx->launches += 3;
    CUtensorMap map_hi, map_lo;
    if ((rc = make_map(&map_hi, HI, N, n, ws)) || (rc = make_map(&map_lo, LO, N, n, ws))) return rc;
    Params p;
    p.tiles = d_tiles;
    p.n_tiles = (int32_t)tiles.size();
    p.stages_total = (n + KC - 1) / KC;
    p.flush_stages = FLUSH_STAGES;
    if (getenv("FUF_GRAM_FLUSH")) p.flush_stages = std::max(1, atoi(getenv("FUF_GRAM_FLUSH")));
    p.S = (int32_t)ws.S;
    p.N = N;
    p.q = q;
    p.D = D;
    p.ldd = ldd;
    FUF_CUDA(cudaFuncSetAttribute(gram_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    const int grid = std::min<int64_t>(p.n_tiles, ctx->sm_count);
    if (!ctx->ev_gram[0]) {
        FUF_CUDA(cudaEventCreate(&ctx->ev_gram[0]));
        FUF_CUDA(cudaEventCreate(&ctx->ev_gram[1]));
    }
    FUF_CUDA(cudaEventRecord(ctx->ev_gram[0], stream));
    gram_tile_kernel<<<grid, NTHREADS, SMEM_BYTES, stream>>>(map_hi, map_lo, p);
    FUF_CUDA(cudaGetLastError());
    FUF_CUDA(cudaEventRecord(ctx->ev_gram[1], stream));
    ctx->gram_flops = 3.0 * 2.0 * TILE * TILE * (double)p.stages_total * KC * (double)p.n_tiles;
    ctx->launches += 1;
    if (full) {
        const unsigned b = (unsigned)((N + 31) / 32);
        mirror_kernel<<<dim3(b, b), 256, 0, stream>>>(D, N, ldd);
        FUF_CUDA(cudaGetLastError());
        ctx->launches += 1;
    }
    return FUF_OK;
}

}  // namespace fuf

using namespace fuf;

extern "C" int fuf_gram_last_kernel(fuf_ctx *ctx, float *ms, double *executed_flops)
{
    FUF_REQUIRE(ctx && ms && executed_flops, "fuf_gram_last_kernel: null argument");
    *ms = 0.f;
    *executed_flops = ctx->gram_flops;
    if (!ctx->ev_gram[0]) return FUF_OK;
    FUF_CUDA(cudaSetDevice(ctx->device));
    FUF_CUDA(cudaEventSynchronize(ctx->ev_gram[1]));
    FUF_CUDA(cudaEventElapsedTime(ms, ctx->ev_gram[0], ctx->ev_gram[1]));
    return FUF_OK;
}

extern "C" int fuf_pairwise_gram(fuf_ctx *ctx, const float *PT, int64_t N, int64_t ldp, int64_t shard,
                                 int64_t shard_stride, int mode, const double *err_in, double *q, double *err_out,
                                 const int32_t *tile_rows_h, int32_t n_tile_rows, double *D, int64_t ldd, void *stream)
{
    FUF_REQUIRE(ctx && PT && q && D, "fuf_pairwise_gram: null argument");
    FUF_REQUIRE(mode == FUF_MODE_L2, "fuf_pairwise_gram: only FUF_MODE_L2 has a Gram form (got mode %d)", mode);
    FUF_REQUIRE(N >= 0 && ldd >= N, "fuf_pairwise_gram: bad shape N=%lld ldd=%lld", (long long)N, (long long)ldd);
    FUF_CUDA(cudaSetDevice(ctx->device));
    FUF_REQUIRE(tile_rows_h != nullptr || n_tile_rows == 0, "fuf_pairwise_gram: n_tile_rows without tile_rows_h");
    return pairwise_gram(ctx, PT, N, ldp, shard, shard_stride, err_in, q, err_out, tile_rows_h, n_tile_rows, D, ldd,
                         (cudaStream_t)stream);
}
