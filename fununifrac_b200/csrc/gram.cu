// gram.cu — --L2 all-pairs distances as an ERROR-FREE integer Gram matrix on the tcgen05 tensor cores.
//
// Replaces, for is_L2, the pair loop of pairwise_computation (src/algorithms/emd_unifrac.py:49-54) with
// get_L2 = sum_k (P_i[k] - P_j[k])^2 (src/objects/profile_vector.py:17-19) rewritten as
//     D[i,j] = q_i + q_j - 2 G[i,j],     G = A A^T,     q_i = sum_k A_i[k]^2.
// A Gram form cancels: for similar samples G ~ q and every relative error of G is amplified by q / D.  A first
// version of this file used 3xTF32 floating-point MMAs; it measured 1e-7 on independent samples and 1e-5 on
// correlated ones, because the tensor core's fp32 accumulation truncates (errors add up linearly with G).  The cure
// is to make the arithmetic exact:
//
//   * block fixed point along the basis.  Node columns are bucketed by the power of two of their largest entry and
//     permuted so that a bucket is a contiguous range of basis rows (D does not depend on the order of the basis);
//     within a bucket with scale S every entry is a / S in [-1, 1], written with four signed 8-bit digits
//         a / S = (d0 + d1 / 254 + d2 / 254^2 + d3 / 254^3) / 127,      |d_p| <= 127      (32.6 bits below S;
//     three digits were tried first: 24.6 bits below a scale that can exceed a typical entry by 16-64x left the
//     refinement criterion flagging every pair, and the truncated product expansion showed at 1e-6)
//   * integer MMAs.  tcgen05.mma kind::i8 (int8 x int8 -> int32 in tensor memory) computes the digit products by
//     order,  c0 = d0 d0',  c1 = d0 d1' + d1 d0',  c2 = d1 d1' + d0 d2' + d2 d0',
//             c3 = d0 d3' + d3 d0' + d1 d2' + d2 d1'      (10 MMAs per 32 basis rows)
//     in four int32 accumulators; sums of at most 32K x 4 products of magnitude < 2^14 cannot overflow, so the
//     accumulation is EXACT and needs no periodic flush; the dropped orders are below 3 x 254^-4 = 7e-10 of S^2.
//   * per bucket, the epilogue adds S^2 (c0 + c1/254 + c2/254^2 + c3/254^3) / 127^2 into float64 accumulators in
//     registers.
// q_i is the same truncated digit form evaluated on (i, i) with the same operations in the same order (integer
// self-products from gram_digits_kernel, combined in gram_stat_kernel), so D[i,j] = B(delta, delta) for the digit
// difference delta of the two samples: identical samples give exactly 0, nothing cancels, and
// |D - exact squared distance of the quantised profiles| is at most 2.9e-9 S^2 per column in which the two samples
// differ.  What is left is the quantisation of the inputs, which is the same kind of error as rounding them to fp32 and
// is covered by the same refinement criterion (refine.cu) through the per-sample statistic sum_k (a - a_quantised)^2
// this file emits.
//
// K3c gram_colmax_kernel  largest |entry|, non-zero count and energy of every node column (one warp per row, 16-byte
//                         loads).  The host turns them into the plan: scale buckets (tiny ones merged into the next
//                         larger scale under an h2 budget), the columns kept in float64, the permuted row map.
// K3p gram_digits_kernel  fp32 operand -> four int8 digit planes in permuted row order, four samples per thread; the
//                         quantisation / scale statistics and the integer self-products of every sample in the same
//                         pass.  HBM-bound.
// K3s gram_stat_kernel    per sample: error statistic, h2, and q_i with the epilogue's own floating-point operations.
// K3  gram_tile_kernel    persistent, warp-specialised, one CTA per SM, 128x128 tiles of the upper triangle:
//       warp 0   TMA producer: per stage (64 basis rows) 8 boxes of 128 samples x 64 rows (A and B, four digit
//                planes each; 1 byte per entry, SWIZZLE_128B), 3-stage 64 KB ring, full/empty mbarriers;
//       warp 1   MMA issuer: one lane, M = N = 128, K = 32, MN-major descriptors, ten MMAs per 32 rows, ~6 instructions
//                per MMA (descriptor low words = stage base + constant); tcgen05.commit frees smem stages and, at the
//                end of a bucket, publishes the accumulators;
//       warps 4-11 (two warpgroups, setmaxnreg 232) epilogue: tcgen05.ld of the four int32 accumulators; each one is
//                handed back to the MMA warp as soon as its raw values are in registers, then weighted into 64 float64
//                registers per thread (int32 -> double without I2F); the float64 bypass columns are added here too.
//       With the full matrix as output the epilogue also stores the transposed tile (no separate mirror pass).
// K3x2 gram_pair_kernel   the same on CTA pairs (cta_group::2): the DEFAULT (FUF_GRAM_2CTA=0 selects gram_tile_kernel);
//                         see its own comment.
//
// Tensor-pipe accounting: executed int8 MACs = 10 x 128 x 128 x (padded n) per tile; algorithmic flops = 2 n per pair.
#include <algorithm>
#include <cmath>
#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace fuf {

namespace gram {

constexpr int TILE = FUF_TILE;              // 128 = UMMA M = UMMA N
constexpr int KC = 64;                      // basis rows per pipeline stage
constexpr int KB = 32;                      // basis rows per tcgen05.mma (kind::i8: 32 bytes of K)
constexpr int STAGES = 3;
constexpr int NDIG = 4;                     // digits per entry = accumulators (product orders 0..3)
constexpr uint32_t PLANE_BYTES = TILE * KC;                 // 8 KB: one digit plane of one operand tile
constexpr uint32_t STAGE_BYTES = 2 * NDIG * PLANE_BYTES;    // A d0..d3 | B d0..d3
constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers + tmem ptr */;
constexpr int NTHREADS = 384;               // warpgroup 0: warp 0 TMA, warp 1 MMA (2 idle); warpgroups 1-2: epilogue
constexpr int NEPI_WARPS = 8;
constexpr uint32_t TMEM_COLS = 512;         // four 128-column int32 accumulators: all of tensor memory
constexpr double DIGIT0 = 127.0, DIGIT = 254.0;
constexpr int GRAM_DIRECT_ROWS = 256;       // node columns of largest scale computed in float64 instead

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

struct Group {
    int32_t stage_end;  // one past the last stage of this bucket
    int32_t pad;
    double w0;          // S^2 / 127^2: weight of the class-0 accumulator (class p: w0 / 254^p)
};

struct Params {
    const Tile *tiles;
    const Group *groups;
    int32_t n_tiles, n_groups;
    int32_t S;  // samples per operand shard (a multiple of 128)
    int64_t N;
    const double *q;
    double *D;
    int64_t ldd;
    int32_t mirror;      // 1: D is the full matrix (out_row0 = 128 ti): also store the transposed tile
    int32_t n_direct;    // node columns kept out of the integer Gram (largest scales): added here in float64
    const double *SUB;   // their values, [n_direct][samples] in layout `sub`
    Lay sub;
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
// Shared-memory matrix descriptors (cute::UMMA::SmemDescriptor), MN-major 8-bit operands.  SWIZZLE_128B: a row of the
// plane is 128 samples = 128 bytes = the whole M (or N) extent, 8 basis rows form a 1 KB swizzle atom; SWIZZLE_64B (the
// N-halves of the CTA-pair kernel): 64-byte rows, 512-byte atoms.
//   [0,14) start >> 4 | [16,30) LBO >> 4 (between blocks along M/N: unused, one block) | [32,46) SBO >> 4 (between
//   8-row atoms) | [46,48) version = 1 | [61,64) layout (2 = SWIZZLE_128B, 4 = SWIZZLE_64B)
// The low word is `desc_lo(address)`, the high word one of the DESC_HI_* constants below.
// Instruction descriptor (cute::UMMA::InstrDescriptor): c = S32 (2), a = b = signed INT8 (1), both MN-major,
// N = 128, M = 128.
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(TILE >> 3) << 17) |
                           ((uint32_t)(TILE >> 4) << 24);

// (Every tcgen05.mma / tcgen05.commit below is executed by the WHOLE MMA warp in converged code and predicated on
// elect.sync — the same leader every time for the same member mask: an `if (lane == 0)` around the instruction makes
// ptxas wrap each one in an ELECT / branch loop, in converged code it becomes one predicated UTCIMMA.)
// The issuing thread is the bottleneck of this kernel if it spends more than the ~68 cycles an MMA takes on building
// its operands: the descriptors of one stage differ only in the 14-bit start-address field, so the low words are
// `stage base + constant` (one add each) and the high word (SBO, version, layout) is a constant.
constexpr uint32_t DESC_HI_SW128 = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
constexpr uint32_t DESC_HI_SW64 = (uint32_t)(512 >> 4) | (1u << 14) | (4u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) { return (smem_addr >> 4) & 0x3FFFu; }
template <int CTA_GROUP, uint32_t HI_A, uint32_t HI_B, uint32_t IDESC_V, int ACCUMULATE>
__device__ __forceinline__ void umma_i8_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo)
{
    if (CTA_GROUP == 1)
        asm volatile(
            "{\n\t.reg .b64 da, db;\n\t.reg .pred p, pe;\n\t"
            "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %4};\n\t"
            "setp.ne.b32 p, %6, 0;\n\t"
            "elect.sync _|pe, 0xffffffff;\n\t"
            "@pe tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %5, p;\n\t}"
            ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "n"(HI_A), "n"(HI_B), "r"(IDESC_V), "n"(ACCUMULATE)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .b64 da, db;\n\t.reg .pred p, pe;\n\t"
            "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %4};\n\t"
            "setp.ne.b32 p, %6, 0;\n\t"
            "elect.sync _|pe, 0xffffffff;\n\t"
            "@pe tcgen05.mma.cta_group::2.kind::i8 [%0], da, db, %5, p;\n\t}"
            ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "n"(HI_A), "n"(HI_B), "r"(IDESC_V), "n"(ACCUMULATE)
            : "memory");
}
// All ten accumulating MMAs of one 32-row block in ONE asm statement: the compiler wraps every asm that issues a
// tcgen05 instruction from a single lane into an ELECT / branch sequence, once per statement.  Plane p of A (B) starts
// 512 descriptor units (8 KB) after plane p - 1; accumulator c is 128 TMEM columns after accumulator c - 1.
__device__ __forceinline__ void umma_block10(uint32_t tmem, uint32_t a_lo, uint32_t b_lo)
{
#define FUF_MMA(C, X, Y)                                                                          \
    "add.u32 la, %1, " #X "*512;\n\tadd.u32 lb, %2, " #Y "*512;\n\t"                              \
    "mov.b64 da, {la, %3};\n\tmov.b64 db, {lb, %3};\n\tadd.u32 td, %0, " #C "*128;\n\t"           \
    "@pe tcgen05.mma.cta_group::1.kind::i8 [td], da, db, %4, p;\n\t"
    asm volatile(
        "{\n\t.reg .b64 da, db;\n\t.reg .b32 la, lb, td;\n\t.reg .pred p, pe;\n\t"
        "setp.eq.b32 p, 0, 0;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        FUF_MMA(0, 0, 0)
        FUF_MMA(1, 0, 1) FUF_MMA(1, 1, 0)
        FUF_MMA(2, 0, 2) FUF_MMA(2, 1, 1) FUF_MMA(2, 2, 0)
        FUF_MMA(3, 0, 3) FUF_MMA(3, 1, 2) FUF_MMA(3, 2, 1) FUF_MMA(3, 3, 0)
        "}"
        ::"r"(tmem), "r"(a_lo), "r"(b_lo), "n"(DESC_HI_SW128), "r"(IDESC)
        : "memory");
#undef FUF_MMA
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile(
        "{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar)
        : "memory");
}
// int32 -> double without the quarter-rate I2F.F64: 2^52 + 2^31 + v is exactly representable, its bit pattern is
// {0x43300000, v ^ 0x80000000}; one LOP3 and one DADD (exact for every int32).
__device__ __forceinline__ double i32_to_f64(uint32_t v)
{
    return __hiloint2double(0x43300000, (int)(v ^ 0x80000000u)) - 4503601774854144.0;
}
// 64 consecutive TMEM columns of this warp's lane quarter in one go (two x32 loads in flight, one wait): the
// accumulator can be handed back to the MMA warp as soon as its raw int32 values sit in registers.
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&v)[64])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%64];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%65];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]),
          "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]),
          "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]),
          "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]),
          "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
        : "r"(taddr), "r"(taddr + 32)
        : "memory");
}

__global__ void __launch_bounds__(NTHREADS, 1)
gram_tile_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
                 const __grid_constant__ CUtensorMap map2, const __grid_constant__ CUtensorMap map3, const Params p)
{
    extern __shared__ unsigned char smem_raw[];
    // swizzle atoms must sit on their own size boundaries (1 KB)
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + STAGES * STAGE_BYTES;
    const uint32_t full0 = bars, empty0 = bars + 8 * STAGES, tfull = bars + 16 * STAGES, tempty0 = tfull + 8;
    const uint32_t tmem_slot = tempty0 + 8 * NDIG;
    volatile uint32_t *tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(full0 + 8 * s, 1);   // producer's arrive.expect_tx (+ the TMA bytes)
            mbar_init(empty0 + 8 * s, 1);  // tcgen05.commit of the MMA warp
        }
        mbar_init(tfull, 1);                                                  // tcgen05.commit after a bucket's last MMA
        for (int c = 0; c < NDIG; c++) mbar_init(tempty0 + 8 * c, NEPI_WARPS);  // accumulator c has been read
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
    const int stages_total = p.groups[p.n_groups - 1].stage_end;

    if (warp < 4) {
        // warpgroup 0 gives its registers to the epilogue warpgroups (the register re-allocation has to sit inside
        // the role branches: ptxas budgets each branch by the setmaxnreg that dominates it)
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == 0) {
            // ===================== TMA producer =====================
            if (lane == 0) {
                asm volatile("prefetch.tensormap [%0];" ::"l"(&map0) : "memory");
                asm volatile("prefetch.tensormap [%0];" ::"l"(&map1) : "memory");
                asm volatile("prefetch.tensormap [%0];" ::"l"(&map2) : "memory");
                asm volatile("prefetch.tensormap [%0];" ::"l"(&map3) : "memory");
                uint32_t it = 0;
                for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
                    const Tile tl = p.tiles[t];
                    const int ia = (tl.ti * TILE) % p.S, ga = (tl.ti * TILE) / p.S;
                    const int ib = (tl.tj * TILE) % p.S, gb = (tl.tj * TILE) / p.S;
                    for (int st = 0; st < stages_total; ++st, ++it) {
                        const uint32_t slot = it % STAGES, phase = (it / STAGES) & 1u;
                        mbar_wait(empty0 + 8 * slot, phase ^ 1u);
                        const uint32_t full = full0 + 8 * slot;
                        mbar_expect_tx(full, STAGE_BYTES);
                        const uint32_t dst = base + slot * STAGE_BYTES;
                        tma_load_3d(dst + 0 * PLANE_BYTES, &map0, full, ia, st * KC, ga);
                        tma_load_3d(dst + 1 * PLANE_BYTES, &map1, full, ia, st * KC, ga);
                        tma_load_3d(dst + 2 * PLANE_BYTES, &map2, full, ia, st * KC, ga);
                        tma_load_3d(dst + 3 * PLANE_BYTES, &map3, full, ia, st * KC, ga);
                        tma_load_3d(dst + 4 * PLANE_BYTES, &map0, full, ib, st * KC, gb);
                        tma_load_3d(dst + 5 * PLANE_BYTES, &map1, full, ib, st * KC, gb);
                        tma_load_3d(dst + 6 * PLANE_BYTES, &map2, full, ib, st * KC, gb);
                        tma_load_3d(dst + 7 * PLANE_BYTES, &map3, full, ib, st * KC, gb);
                    }
                }
            }
        } else if (warp == 1) {
            // ===================== MMA issuer (one elected lane) =====================
            uint32_t it = 0, grp_it = 0;
            for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
                int st = 0;
                for (int g = 0; g < p.n_groups; ++g, ++grp_it) {
                    const int s_end = p.groups[g].stage_end;
                    const uint32_t eph = (grp_it & 1u) ^ 1u;  // parity of "the previous bucket has been read"
                    bool fresh = true;                        // the first MMA of a class overwrites its accumulator
                    for (; st < s_end; ++st, ++it) {
                        const uint32_t slot = it % STAGES, phase = (it / STAGES) & 1u;
                        mbar_wait(full0 + 8 * slot, phase);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t a_lo = desc_lo(base + slot * STAGE_BYTES), b_lo = a_lo + ((NDIG * PLANE_BYTES) >> 4);
                        if (fresh) {
                            // first 32 rows of a bucket: accumulator c must have been drained by the epilogue, and the
                            // first product of every order overwrites it
#pragma unroll
                            for (int c = 0; c < NDIG; c++) {
                                mbar_wait(tempty0 + 8 * c, eph);
                                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                                umma_i8_lo<1, DESC_HI_SW128, DESC_HI_SW128, IDESC, 0>(
                                    tmem_base + c * TILE, a_lo, b_lo + ((c * PLANE_BYTES) >> 4));
#pragma unroll
                                for (int x = 1; x <= c; x++)
                                    umma_i8_lo<1, DESC_HI_SW128, DESC_HI_SW128, IDESC, 1>(
                                        tmem_base + c * TILE, a_lo + ((x * PLANE_BYTES) >> 4),
                                        b_lo + (((c - x) * PLANE_BYTES) >> 4));
                            }
                            fresh = false;
                        } else {
                            umma_block10(tmem_base, a_lo, b_lo);   // whole warp: one elected lane issues
                        }
                        {   // second 32 rows of the stage: order c = all products d_x d'_y with x + y = c
                            constexpr uint32_t off = (KB * TILE) >> 4;
                            umma_block10(tmem_base, a_lo + off, b_lo + off);
                        }
                        umma_commit(empty0 + 8 * slot);  // frees the smem stage when the MMAs have read it
                    }
                    umma_commit(tfull);  // the accumulators of this bucket are complete
                }
            }
        }
    } else {
        // ===================== epilogue warps: int32 TMEM accumulators -> float64 registers =====================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        const int quarter = warp & 3;         // tcgen05.ld: a warp may touch TMEM lanes 32 (warp % 4) .. + 31
        const int half = (warp >> 2) - 1;     // columns [64 half, 64 half + 64)
        const int row = quarter * 32 + lane;
        uint32_t grp_it = 0;
        for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
            const Tile tl = p.tiles[t];
            double acc[64];
#pragma unroll
            for (int e = 0; e < 64; e++) acc[e] = 0.0;
            for (int g = 0; g < p.n_groups; ++g, ++grp_it) {
                double w = p.groups[g].w0;
                mbar_wait(tfull, grp_it & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int c = 0; c < NDIG; c++) {
                    uint32_t v[64];
                    tmem_ld64(tmem_base + ((uint32_t)(quarter * 32) << 16) + c * TILE + half * 64, v);
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty0 + 8 * c);  // accumulator c may be overwritten: the raw values
#pragma unroll                                                    // are in registers, the arithmetic comes after
                    for (int e = 0; e < 64; e++) acc[e] = fma(i32_to_f64(v[e]), w, acc[e]);
                    w *= (1.0 / DIGIT);
                }
            }
            const int64_t gi = (int64_t)tl.ti * TILE + row, gj0 = (int64_t)tl.tj * TILE + half * 64;
            const bool mirror = p.mirror != 0 && tl.ti != tl.tj;
            // bypass columns in float64: D += sum_k (a_i - a_j)^2, folded into acc as -(a_i - a_j)^2 / 2 (exact scaling);
            // all lanes of a warp read the same a_j (one broadcast transaction), a_i is coalesced along the lanes
            if (p.n_direct > 0 && gi < p.N) {
                for (int32_t k = 0; k < p.n_direct; k++) {
                    const double ai = p.SUB[p.sub.at(k, gi)];
                    const double *bj = p.SUB + p.sub.at(k, gj0);
#pragma unroll
                    for (int e = 0; e < 64; e++) {
                        const double d = ai - bj[e];            // columns >= N hold zeros and are never stored
                        acc[e] = fma(-0.5 * d, d, acc[e]);
                    }
                }
            }
            if (gi < p.N) {
                const double qi = p.q[gi];
                double *drow = p.D + ((int64_t)tl.out_row0 + row) * p.ldd + gj0;
#pragma unroll
                for (int e = 0; e < 64; e++) {
                    const int64_t gj = gj0 + e;
                    if (gj < p.N) {
                        const double v = (gi == gj) ? 0.0 : (qi + p.q[gj]) - 2.0 * acc[e];
                        drow[e] = v;
                        if (mirror) p.D[gj * p.ldd + gi] = v;    // lanes = consecutive gi: one 256-byte store per warp
                    }
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

// ---------------------------------------------------------------------------------------------------------------------
// K3x2 gram_pair_kernel: the same computation on CTA PAIRS (thread-block clusters of 2, tcgen05 cta_group::2).
// With one CTA per tile every operand byte is fetched and read per CTA: an int8 MMA (M = N = 128, K = 32)
// reads 8 KB of operands per 68 cycles (120 B/clk) while TMA writes another 48 B/clk into the same banks.  A pair
// computes a 256 x 128 tile: M = 256 are the samples of TWO column tiles (2b, 2b+1), one per CTA, N = 128 the samples
// of row tile ti, split in halves of 64 between the two CTAs.  Per CTA and MMA that is 4 KB (A) + 2 KB (half of B):
// 90 B/clk of reads + 36 B/clk of TMA writes, and a stage shrinks from 64 KB to 48 KB (4 stages).
//   * both CTAs run a TMA producer for their own operands; every byte is accounted on CTA 0's `full` barrier
//     (cp.async.bulk.tensor ... .cta_group::2 with the barrier address of the even CTA);
//   * CTA 0's MMA warp issues tcgen05.mma.cta_group::2; its commits are multicast to the `empty` / `tfull` barriers
//     of both CTAs;
//   * each CTA's accumulators (TMEM lanes = its 128 column-tile samples, columns = the 128 row-tile samples) hold the
//     TRANSPOSED tile: the epilogue's stores into D[row][column] run along the lanes (coalesced);
//   * the epilogue warps of both CTAs release accumulator c on CTA 0's `tempty[c]` (remote mbarrier arrive).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int STAGES2 = 4;
constexpr uint32_t PLANE_B2 = 64 * KC;                                   // 4 KB: half of B, one digit plane
constexpr uint32_t STAGE2_BYTES = NDIG * PLANE_BYTES + NDIG * PLANE_B2;  // 48 KB per CTA
constexpr uint32_t SMEM2_BYTES = STAGES2 * STAGE2_BYTES + 1024 + 256;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;                              // shared::cluster address of the even CTA

struct PairTile {
    int32_t ti, b;       // row tile, column-tile pair (2b, 2b + 1)
    int32_t out_row0, pad;   // pad: bits 0-11 dependency + 1 (0 = none), bits 12-31 segment + 1 (0 = none)
};
constexpr int32_t GRAM_STRIP = 12;                    // tile rows per rasterisation strip (74 CTA pairs ~ 12 rows x 6 column pairs)
constexpr uint32_t GRAM_SEG_UNITS = 2 * NEPI_WARPS;   // a finished item advances its segment counter by this much
static_assert(GRAM_SEG_UNITS == FUF_GRAM_SEG_UNITS, "include/fununifrac.h documents the count per item");

struct Params2 {
    const PairTile *tiles;
    const Group *groups;
    int32_t n_tiles, n_groups;
    int32_t S;
    int32_t tiles_per_dim;
    int64_t N;
    const double *q;
    double *D;
    int64_t ldd;
    int32_t mirror, n_direct;
    const double *SUB;
    Lay sub;
    // sharded jobs (fuf_gram_tiles): digit planes of peer shards arrive while the kernel runs; refinement flags are
    // evaluated where the distances are produced
    const uint32_t *dep_flags;   // PairTile::pad = dependency + 1: wait until dep_flags[dependency] reaches `epoch`
    uint32_t epoch;
    const double *err;           // per-sample rounding statistic (push-up + quantisation): D < kappa_r (sqrt e_i + sqrt e_j)^2
    const double *h2;            // per-sample scale statistic: D < kappa_l (h2_i + h2_j)
    double kappa_r, kappa_l;
    long long *flag_list;        // [0] = count, then (i, j) pairs; nullptr: no flagging
    long long list_cap;
    // completion counters (one per segment of the item list): every epilogue warp of the pair adds 1 once its part of the
    // item is stored and fenced, so a copy engine waiting on the counter (fuf_stream_wait_u32) can move the finished part
    // of D to the host while the kernel works on the next segment
    uint32_t *seg_done;
};

__device__ __forceinline__ void gram_wait_shard(const uint32_t *flag, uint32_t epoch)
{
    for (uint32_t spin = 0;; ++spin) {
        uint32_t seen;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
        if ((int32_t)(seen - epoch) >= 0) break;
        if (spin > (1u << 24)) __trap();        // ~3 s: a lost copy is a CUDA error, not a hung GPU
        __nanosleep(200);
    }
    asm volatile("fence.proxy.async.global;" ::: "memory");     // the TMA reads come after the observation
}

__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap *map, uint32_t bar_cta0, int c0, int c1,
                                                int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
        "%5}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar_cta0), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta0(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & PEER_MASK) : "memory");
}
constexpr uint32_t IDESC2 = (2u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(TILE >> 3) << 17) |
                            ((uint32_t)((2 * TILE) >> 4) << 24);   // M = 256 over the pair, N = 128
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar)
{
    asm volatile(
        "{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
        "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
        ::"r"(bar), "h"((uint16_t)3)
        : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
gram_pair_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                 const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapA3,
                 const __grid_constant__ CUtensorMap mapB0, const __grid_constant__ CUtensorMap mapB1,
                 const __grid_constant__ CUtensorMap mapB2, const __grid_constant__ CUtensorMap mapB3, const Params2 p)
{
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // same offset in both CTAs of the pair
    const uint32_t bars = base + STAGES2 * STAGE2_BYTES;
    const uint32_t full0 = bars, empty0 = bars + 8 * STAGES2, tfull = bars + 16 * STAGES2, tempty0 = tfull + 8;
    const uint32_t tmem_slot = tempty0 + 8 * NDIG;
    volatile uint32_t *tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES2; s++) {
            mbar_init(full0 + 8 * s, 1);   // CTA 0: its producer's arrive.expect_tx (+ the TMA bytes of both CTAs)
            mbar_init(empty0 + 8 * s, 1);  // both: multicast tcgen05.commit of the MMA warp
        }
        mbar_init(tfull, 1);                                                        // both: multicast commit
        for (int c = 0; c < NDIG; c++) mbar_init(tempty0 + 8 * c, 2 * NEPI_WARPS);  // CTA 0: epilogue warps of the pair
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {  // warp 1 of BOTH CTAs: paired tensor-memory allocation
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();   // the peer's barriers are initialised before anything can signal them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot_ptr;
    const int stages_total = p.groups[p.n_groups - 1].stage_end;

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == 0) {
            // ===================== TMA producer (both CTAs, each for its own operands) =====================
            if (lane == 0) {
                uint32_t it = 0;
                for (int t = pair; t < p.n_tiles; t += n_pairs) {
                    const PairTile tl = p.tiles[t];
                    if (tl.pad & 0xfff) gram_wait_shard(p.dep_flags + ((tl.pad & 0xfff) - 1), p.epoch);   // a peer's shard still in flight
                    const int sa = (2 * tl.b + (int)rank) * TILE;          // first sample of this CTA's column tile
                    const int sb = tl.ti * TILE + 64 * (int)rank;          // its half of the row tile
                    const int ia = sa % p.S, ga = sa / p.S, ib = sb % p.S, gb = sb / p.S;
                    for (int st = 0; st < stages_total; ++st, ++it) {
                        const uint32_t slot = it % STAGES2, phase = (it / STAGES2) & 1u;
                        mbar_wait(empty0 + 8 * slot, phase ^ 1u);
                        const uint32_t full = (full0 + 8 * slot) & PEER_MASK;   // CTA 0's barrier
                        if (rank == 0) mbar_expect_tx(full0 + 8 * slot, 2 * STAGE2_BYTES);
                        const uint32_t dst = base + slot * STAGE2_BYTES, dstB = dst + NDIG * PLANE_BYTES;
                        tma_load_3d_2sm(dst + 0 * PLANE_BYTES, &mapA0, full, ia, st * KC, ga);
                        tma_load_3d_2sm(dst + 1 * PLANE_BYTES, &mapA1, full, ia, st * KC, ga);
                        tma_load_3d_2sm(dst + 2 * PLANE_BYTES, &mapA2, full, ia, st * KC, ga);
                        tma_load_3d_2sm(dst + 3 * PLANE_BYTES, &mapA3, full, ia, st * KC, ga);
                        tma_load_3d_2sm(dstB + 0 * PLANE_B2, &mapB0, full, ib, st * KC, gb);
                        tma_load_3d_2sm(dstB + 1 * PLANE_B2, &mapB1, full, ib, st * KC, gb);
                        tma_load_3d_2sm(dstB + 2 * PLANE_B2, &mapB2, full, ib, st * KC, gb);
                        tma_load_3d_2sm(dstB + 3 * PLANE_B2, &mapB3, full, ib, st * KC, gb);
                    }
                }
            }
        } else if (warp == 1 && rank == 0) {
            // ===================== MMA issuer: one elected lane of the even CTA, for the pair =====================
            uint32_t it = 0, grp_it = 0;
            for (int t = pair; t < p.n_tiles; t += n_pairs) {
                int st = 0;
                for (int g = 0; g < p.n_groups; ++g, ++grp_it) {
                    const int s_end = p.groups[g].stage_end;
                    const uint32_t eph = (grp_it & 1u) ^ 1u;
                    bool fresh = true;
                    for (; st < s_end; ++st, ++it) {
                        const uint32_t slot = it % STAGES2, phase = (it / STAGES2) & 1u;
                        mbar_wait(full0 + 8 * slot, phase);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t a_lo = desc_lo(base + slot * STAGE2_BYTES), b_lo = a_lo + ((NDIG * PLANE_BYTES) >> 4);
                        if (fresh) {
#pragma unroll
                            for (int c = 0; c < NDIG; c++) {
                                mbar_wait(tempty0 + 8 * c, eph);
                                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                                umma_i8_lo<2, DESC_HI_SW128, DESC_HI_SW64, IDESC2, 0>(
                                    tmem_base + c * TILE, a_lo, b_lo + ((c * PLANE_B2) >> 4));
#pragma unroll
                                for (int x = 1; x <= c; x++)
                                    umma_i8_lo<2, DESC_HI_SW128, DESC_HI_SW64, IDESC2, 1>(
                                        tmem_base + c * TILE, a_lo + ((x * PLANE_BYTES) >> 4),
                                        b_lo + (((c - x) * PLANE_B2) >> 4));
                            }
                            fresh = false;
                        } else {
#pragma unroll
                            for (int c = 0; c < NDIG; c++)
#pragma unroll
                                for (int x = 0; x <= c; x++)
                                    umma_i8_lo<2, DESC_HI_SW128, DESC_HI_SW64, IDESC2, 1>(
                                        tmem_base + c * TILE, a_lo + ((x * PLANE_BYTES) >> 4),
                                        b_lo + (((c - x) * PLANE_B2) >> 4));
                        }
                        {
                            constexpr uint32_t offA = (KB * TILE) >> 4, offB = (KB * 64) >> 4;
#pragma unroll
                            for (int c = 0; c < NDIG; c++)
#pragma unroll
                                for (int x = 0; x <= c; x++)
                                    umma_i8_lo<2, DESC_HI_SW128, DESC_HI_SW64, IDESC2, 1>(
                                        tmem_base + c * TILE, a_lo + ((x * PLANE_BYTES) >> 4) + offA,
                                        b_lo + (((c - x) * PLANE_B2) >> 4) + offB);
                        }
                        umma_commit_2sm(empty0 + 8 * slot);
                    }
                    umma_commit_2sm(tfull);
                }
            }
        }
    } else {
        // ===================== epilogue warps (both CTAs): transposed tile, lanes = column samples =====================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        const int quarter = warp & 3;
        const int half = (warp >> 2) - 1;     // row-tile samples [64 half, 64 half + 64) = TMEM columns
        uint32_t grp_it = 0;
        for (int t = pair; t < p.n_tiles; t += n_pairs) {
            const PairTile tl = p.tiles[t];
            double acc[64];
#pragma unroll
            for (int e = 0; e < 64; e++) acc[e] = 0.0;
            for (int g = 0; g < p.n_groups; ++g, ++grp_it) {
                double w = p.groups[g].w0;
                mbar_wait(tfull, grp_it & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int c = 0; c < NDIG; c++) {
                    uint32_t v[64];
                    tmem_ld64(tmem_base + ((uint32_t)(quarter * 32) << 16) + c * TILE + half * 64, v);
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cta0(tempty0 + 8 * c);
#pragma unroll
                    for (int e = 0; e < 64; e++) acc[e] = fma(i32_to_f64(v[e]), w, acc[e]);
                    w *= (1.0 / DIGIT);
                }
            }
            const int tj = 2 * tl.b + (int)rank;
            const bool live = tj >= tl.ti && tj < p.tiles_per_dim;   // else: the lower-triangle / out-of-range half of a pair
            const int64_t gj = (int64_t)tj * TILE + quarter * 32 + lane, gi0 = (int64_t)tl.ti * TILE + half * 64;
            const bool mirror = p.mirror != 0 && tl.ti != tj;
            if (live && p.n_direct > 0 && gj < p.N) {
                for (int32_t k = 0; k < p.n_direct; k++) {
                    const double aj = p.SUB[p.sub.at(k, gj)];
                    const double *bi = p.SUB + p.sub.at(k, gi0);
#pragma unroll
                    for (int e = 0; e < 64; e++) {
                        const double d = bi[e] - aj;
                        acc[e] = fma(-0.5 * d, d, acc[e]);
                    }
                }
            }
            if (live && gj < p.N) {
                const double qj = p.q[gj];
                const double rej = p.flag_list ? sqrt(p.err[gj]) : 0.0, hj = p.flag_list ? p.h2[gj] : 0.0;
                double *dcol = p.D + ((int64_t)tl.out_row0 + half * 64) * p.ldd + gj;   // walks down the rows
                double *mrow = p.D + gj * p.ldd + gi0;
#pragma unroll
                for (int e = 0; e < 64; e++) {
                    const int64_t gi = gi0 + e;
                    if (gi < p.N) {
                        const double v = (gi == gj) ? 0.0 : (p.q[gi] + qj) - 2.0 * acc[e];
                        acc[e] = v;
                        if (p.flag_list && gi < gj) {
                            // the two criteria of the Gram refinement (refine.cu): rounding of the operands, and the
                            // dropped digit-product orders measured by the scale statistic
                            const double b = sqrt(p.err[gi]) + rej;
                            bool flag = !(v >= p.kappa_r * b * b) && b != 0.0;
                            flag = flag || !(v >= p.kappa_l * (p.h2[gi] + hj));
                            if (flag) {
                                const unsigned long long at = atomicAdd(reinterpret_cast<unsigned long long *>(p.flag_list), 1ull);
                                if ((long long)at < p.list_cap) {
                                    p.flag_list[1 + 2 * at] = gi;
                                    p.flag_list[2 + 2 * at] = gj;
                                }
                            }
                        }
                        dcol[(int64_t)e * p.ldd] = v;         // lanes = consecutive gj: one 256-byte store per warp
                    }
                }
                if (mirror) {
                    // the transposed tile: every lane owns 64 consecutive entries of row gj.  32-byte stores (one full
                    // sector each) when the row segment is aligned — the matrix may live on a peer GPU, where partial
                    // sectors cost NVLink packets
                    if ((((uintptr_t)mrow) & 31) == 0 && gi0 + 64 <= p.N) {
#pragma unroll
                        for (int e = 0; e < 64; e += 4)
                            asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(mrow + e), "d"(acc[e]), "d"(acc[e + 1]),
                                         "d"(acc[e + 2]), "d"(acc[e + 3])
                                         : "memory");
                    } else {
#pragma unroll
                        for (int e = 0; e < 64; e++)
                            if (gi0 + e < p.N) mrow[e] = acc[e];
                    }
                }
            }
            if (p.seg_done && (tl.pad >> 12)) {
                __threadfence();      // every lane's stores, then the warp's one count
                __syncwarp();
                if (lane == 0) atomicAdd(p.seg_done + ((tl.pad >> 12) - 1), 1u);
            }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();   // both CTAs are done with both tensor memories
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Tensor-pipe peak for THIS instruction mix (bench.py's roofline denominator, measured instead of assumed): the MMA
// warp issues the same ten kind::i8 MMAs per 32 basis rows as the tile kernels, back to back, on operands that already
// sit in shared memory (pseudo-random digits; no TMA, no epilogue, accumulators never read).  CTA_GROUP = 1: M = N = 128
// per CTA; CTA_GROUP = 2: M = 256 over a CTA pair, N = 128.  Every SM runs one CTA.
// ---------------------------------------------------------------------------------------------------------------------
template <int CTA_GROUP>
__global__ void __launch_bounds__(128, 1)
i8_mma_peak_kernel(int32_t iters)
{
    extern __shared__ unsigned char smem_raw[];
    constexpr uint32_t B_PLANE = CTA_GROUP == 1 ? PLANE_BYTES : PLANE_B2;
    constexpr uint32_t OPER_BYTES = NDIG * PLANE_BYTES + NDIG * B_PLANE;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar = base + OPER_BYTES, tmem_slot = bar + 8;
    unsigned char *oper = smem_raw + (base - smem_u32(smem_raw));
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(oper + OPER_BYTES + 8);
    const int warp = threadIdx.x >> 5;
    uint32_t rank = 0;
    if (CTA_GROUP == 2) rank = cluster_ctarank();
    for (uint32_t i = threadIdx.x; i < OPER_BYTES / 4; i += blockDim.x) {
        uint32_t h = (i + 977u * blockIdx.x) * 2654435761u;
        h ^= h >> 15;
        reinterpret_cast<uint32_t *>(oper)[i] = h * 2246822519u;      // digits in [-128, 127]: any int8 is a valid operand
    }
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        if (CTA_GROUP == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CTA_GROUP == 2) cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot_ptr;
    if (warp == 0 && rank == 0) {
        const uint32_t a_lo = desc_lo(base), b_lo = a_lo + ((NDIG * PLANE_BYTES) >> 4);
        for (int32_t it = 0; it < iters; it++) {
            if (CTA_GROUP == 1) {
                umma_block10(tmem_base, a_lo, b_lo);
                umma_block10(tmem_base, a_lo + ((KB * TILE) >> 4), b_lo + ((KB * TILE) >> 4));
            } else {
#pragma unroll
                for (int h = 0; h < 2; h++)
#pragma unroll
                    for (int c = 0; c < NDIG; c++)
#pragma unroll
                        for (int x = 0; x <= c; x++)
                            umma_i8_lo<2, DESC_HI_SW128, DESC_HI_SW64, IDESC2, 1>(
                                tmem_base + c * TILE, a_lo + ((x * PLANE_BYTES) >> 4) + h * ((KB * TILE) >> 4),
                                b_lo + (((c - x) * PLANE_B2) >> 4) + h * ((KB * 64) >> 4));
            }
        }
        if (CTA_GROUP == 1) umma_commit(bar); else umma_commit_2sm(bar);
    }
    mbar_wait(bar, 0);     // every MMA above has completed
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CTA_GROUP == 2) cluster_sync_all();
    if (warp == 0) {
        if (CTA_GROUP == 1)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// Per node column (one warp per row): cmax[k] = max_s |PT[k][s]|, nnz[k] = number of non-zero entries,
// energy[k] = sum_s PT[k][s]^2 (fp32: it only steers the bypass heuristic).  Shard by shard, 16-byte loads, four of
// them in flight per lane.
__global__ void __launch_bounds__(256)
gram_colmax_kernel(const float *__restrict__ PT, const Lay lay, int64_t N, int32_t n, float *__restrict__ cmax,
                   float *__restrict__ nnz, float *__restrict__ energy)
{
    const int32_t k = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (k >= n) return;
    const int lane = threadIdx.x & 31;
    float m = 0.f, c = 0.f, e = 0.f;
    auto take = [&](float a) {
        m = fmaxf(m, fabsf(a));
        c += (a != 0.f) ? 1.f : 0.f;
        e = fmaf(a, a, e);
    };
    for (int64_t s0 = 0; s0 < N; s0 += lay.S) {
        const float *row = PT + lay.at(k, s0);               // 16-byte aligned: S, LD and SS are multiples of 4
        const int64_t len = min(lay.S, N - s0), len4 = len >> 2;
        const float4 *row4 = reinterpret_cast<const float4 *>(row);
        int64_t i = lane;
        for (; i + 96 < len4; i += 128) {
            const float4 v0 = row4[i], v1 = row4[i + 32], v2 = row4[i + 64], v3 = row4[i + 96];
            take(v0.x), take(v0.y), take(v0.z), take(v0.w);
            take(v1.x), take(v1.y), take(v1.z), take(v1.w);
            take(v2.x), take(v2.y), take(v2.z), take(v2.w);
            take(v3.x), take(v3.y), take(v3.z), take(v3.w);
        }
        for (; i < len4; i += 32) {
            const float4 v = row4[i];
            take(v.x), take(v.y), take(v.z), take(v.w);
        }
        for (int64_t j = (len4 << 2) + lane; j < len; j += 32) take(row[j]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        c += __shfl_xor_sync(0xffffffffu, c, o);
        e += __shfl_xor_sync(0xffffffffu, e, o);
    }
    if (lane == 0) {
        cmax[k] = m;
        nnz[k] = c;
        energy[k] = e;
    }
}

// One part = up to 512 permuted rows of ONE bucket.
struct Part {
    int32_t row0, row1;   // permuted rows [row0, row1)
    int32_t group, pad;
};

// Permuted row r (node row_map[r], bucket scale 1 / inv_scale[r]) of every sample -> four int8 digits, four samples per
// thread (one 16-byte load, one 4-byte store per digit plane).  Per (part, sample) it also emits
//   r2 = sum of (quantisation error)^2,  h2 = sum of S^2 over the non-zero entries,
//   c_p = sum over the rows of sum_{x+y=p} d_x d_y  (p = 0..3, exact integers: the sample's digit form with itself).
// Padding rows (row_map < 0) and padding samples (s >= N) get zero digits.
__global__ void __launch_bounds__(128)
gram_digits_kernel(const float *__restrict__ PT, const Lay lay, const Lay ws, int64_t n_cols, int64_t N,
                   const Part *__restrict__ parts, const int32_t *__restrict__ row_map,
                   const float *__restrict__ inv_scale, int8_t *__restrict__ D0, int8_t *__restrict__ D1,
                   int8_t *__restrict__ D2, int8_t *__restrict__ D3, double *__restrict__ r2, double *__restrict__ h2,
                   int32_t *__restrict__ cself)
{
    const int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (s >= n_cols) return;
    const Part part = parts[blockIdx.y];
    double err[4] = {0.0, 0.0, 0.0, 0.0}, hh[4] = {0.0, 0.0, 0.0, 0.0};
    int32_t c[4][4] = {};
    const bool any = s < N;
    for (int32_t r = part.row0; r < part.row1; r++) {
        const int32_t k = row_map[r];
        int d[4][4] = {};                 // [sample][digit]
        if (k >= 0 && any) {
            const float4 v = *reinterpret_cast<const float4 *>(PT + lay.at(k, s));
            const double is = (double)inv_scale[r];             // a power of two: a * is is exact and in [-1, 1]
            const double unit = 1.0 / (is * DIGIT0) * (1.0 / (DIGIT * DIGIT * DIGIT));   // value of one step of d3
            const double S2 = 1.0 / (is * is);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const float a = u == 0 ? v.x : u == 1 ? v.y : u == 2 ? v.z : v.w;
                if (a != 0.f && s + u < N) {
                    double t = (double)a * is * DIGIT0;
                    const double f0 = rint(t);
                    t = (t - f0) * DIGIT;
                    const double f1 = rint(t);
                    t = (t - f1) * DIGIT;
                    const double f2 = rint(t);
                    t = (t - f2) * DIGIT;
                    const double f3 = rint(t);
                    d[u][0] = (int)f0, d[u][1] = (int)f1, d[u][2] = (int)f2, d[u][3] = (int)f3;
                    const double e = (t - f3) * unit;
                    err[u] = fma(e, e, err[u]);
                    hh[u] += S2;                                // S^2 of every non-zero entry
                    c[u][0] += d[u][0] * d[u][0];
                    c[u][1] += 2 * d[u][0] * d[u][1];
                    c[u][2] += d[u][1] * d[u][1] + 2 * d[u][0] * d[u][2];
                    c[u][3] += 2 * d[u][0] * d[u][3] + 2 * d[u][1] * d[u][2];
                }
            }
        }
        const int64_t o = ws.at(r, s);
        auto pack = [&](int p) {
            return (uint32_t)(uint8_t)d[0][p] | ((uint32_t)(uint8_t)d[1][p] << 8) | ((uint32_t)(uint8_t)d[2][p] << 16) |
                   ((uint32_t)(uint8_t)d[3][p] << 24);
        };
        *reinterpret_cast<uint32_t *>(D0 + o) = pack(0);
        *reinterpret_cast<uint32_t *>(D1 + o) = pack(1);
        *reinterpret_cast<uint32_t *>(D2 + o) = pack(2);
        *reinterpret_cast<uint32_t *>(D3 + o) = pack(3);
    }
    const int64_t base = (int64_t)blockIdx.y * n_cols + s;
    *reinterpret_cast<double4 *>(r2 + base) = make_double4(err[0], err[1], err[2], err[3]);
    *reinterpret_cast<double4 *>(h2 + base) = make_double4(hh[0], hh[1], hh[2], hh[3]);
#pragma unroll
    for (int p = 0; p < 4; p++)
        *reinterpret_cast<int4 *>(cself + ((int64_t)blockIdx.y * 4 + p) * n_cols + s) =
            make_int4(c[0][p], c[1][p], c[2][p], c[3][p]);
}

// Gather of the node rows that bypass the integer Gram (largest scales): SUB[slot][s] = (double) PT[node][s].
__global__ void __launch_bounds__(256)
gram_gather_kernel(const float *__restrict__ PT, const Lay lay, const Lay sub, int64_t n_cols, int64_t N,
                   const int32_t *__restrict__ direct_nodes, double *__restrict__ SUB)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_cols) return;
    const int32_t k = direct_nodes[blockIdx.y];
    SUB[sub.at(blockIdx.y, s)] = (k >= 0 && s < N) ? (double)PT[lay.at(k, s)] : 0.0;
}

// Per sample, from the per-part outputs of gram_digits_kernel:
//   err_out[s] = (sqrt(err_in[s]) + sqrt(sum r2))^2: the statistics of two successive roundings (fp32 rounding in the
//     push-up, digit quantisation here) add in norm;
//   h2[s] = sum of S^2 over the sample's non-zero entries (bounds what the dropped product orders can contribute, see
//     fuf_pairwise_gram);
//   q[s] = the truncated digit form of sample s with itself: per bucket the integer sums c_p (exact), combined with
//     the SAME floating-point operations in the same order as the tile epilogue uses for G[i,j] — so that
//     q_i + q_j - 2 G[i,j] is exactly 0 when the digit rows of i and j coincide.
__global__ void gram_stat_kernel(const double *__restrict__ r2part, const double *__restrict__ h2part,
                                 const int32_t *__restrict__ cself, const Part *__restrict__ parts, int32_t n_parts,
                                 const Group *__restrict__ groups, int64_t ldp, const double *__restrict__ err_in,
                                 int64_t N, double *__restrict__ h2, double *__restrict__ err_out, double *__restrict__ q)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    double rs = 0.0, hs = 0.0, acc = 0.0;
    for (int32_t b = 0; b < n_parts;) {
        const int32_t g = parts[b].group;
        long long t[4] = {0, 0, 0, 0};
        for (; b < n_parts && parts[b].group == g; b++) {
            rs += r2part[(int64_t)b * ldp + s];
            hs += h2part[(int64_t)b * ldp + s];
#pragma unroll
            for (int p = 0; p < 4; p++) t[p] += cself[((int64_t)b * 4 + p) * ldp + s];
        }
        double w = groups[g].w0;
#pragma unroll
        for (int p = 0; p < 4; p++) {   // the epilogue's operations, in the epilogue's order
            acc = fma((double)(int32_t)t[p], w, acc);
            w *= (1.0 / DIGIT);
        }
    }
    q[s] = acc;
    if (h2) h2[s] = hs;
    if (err_out) {
        const double a = sqrt(err_in ? err_in[s] : 0.0) + sqrt(rs);
        err_out[s] = a * a;
    }
}

static int make_map(CUtensorMap *map, const int8_t *ptr, int64_t N, int32_t n_rows, const Lay &ws, uint32_t box_samples = TILE)
{
    EncodeTiledFn encode = tensor_map_encoder();
    if (!encode) {
        set_error("fuf_pairwise_gram: cuTensorMapEncodeTiled is not available from this driver");
        return FUF_ECUDA;
    }
    // dims (sample in shard, permuted basis row, shard), one byte per entry
    const int64_t n_shards = ws.SS ? (N + ws.S - 1) / ws.S : 1;
    const cuuint64_t gdim[3] = {(cuuint64_t)ws.S, (cuuint64_t)n_rows, (cuuint64_t)n_shards};
    const cuuint64_t gstride[2] = {(cuuint64_t)ws.LD, (cuuint64_t)(ws.SS ? ws.SS : ws.LD * (int64_t)n_rows)};
    const cuuint32_t box[3] = {box_samples, KC, 1}, estr[3] = {1, 1, 1};   // 128-byte rows: SWIZZLE_128B, 64-byte: 64B
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)ptr, gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        box_samples == TILE ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("fuf_pairwise_gram: cuTensorMapEncodeTiled failed with CUresult %d (N=%lld rows=%d S=%lld)", (int)r,
                  (long long)N, n_rows, (long long)ws.S);
        return FUF_ECUDA;
    }
    return FUF_OK;
}

// The host-side plan of a Gram job from the column statistics (cmax | nnz | energy, n floats each): scale buckets,
// float64 bypass columns, permuted row map.  Deterministic in its input, so ranks that all-reduce the statistics of
// their shards arrive at the same plan.
struct GramPlanHost {
    std::vector<int32_t> row_map, direct_nodes;
    std::vector<float> inv_scale;
    std::vector<Group> groups;
    std::vector<Part> parts;
    int32_t n_rows = 0, n_direct = 0, n_direct_pad = 16;
};

static int gram_make_plan(fuf_ctx *ctx, const std::vector<float> &cstat, int32_t n, GramPlanHost &plan)
{
    const float *cmax = cstat.data(), *nnz = cmax + n, *energy = nnz + n;
    std::vector<std::pair<int32_t, int32_t>> keyed, direct_keyed;  // (bucket exponent, node), all-zero columns dropped
    keyed.reserve(n);
    for (int32_t k = 0; k < n; k++) {
        FUF_REQUIRE(std::isfinite(cmax[k]), "fuf_pairwise_gram: profile column %d holds a non-finite value", k);
        if (cmax[k] > 0.f) {
            int e;
            std::frexp(cmax[k], &e);                    // cmax = m 2^e, m in [0.5, 1): cmax < 2^e = the bucket scale
            keyed.emplace_back(e, k);
        }
    }
    {   // order: decreasing exponent, then node index — a counting sort (float exponents span < 300 values, the nodes
        // arrive in increasing order); this runs on the host while the GPU waits for the plan
        int32_t e_min = INT32_MAX, e_max = INT32_MIN;
        for (const auto &kv : keyed) e_min = std::min(e_min, kv.first), e_max = std::max(e_max, kv.first);
        if (!keyed.empty()) {
            std::vector<size_t> start((size_t)(e_max - e_min) + 2, 0);
            for (const auto &kv : keyed) start[(size_t)(e_max - kv.first) + 1]++;
            for (size_t i = 1; i < start.size(); i++) start[i] += start[i - 1];
            std::vector<std::pair<int32_t, int32_t>> sorted(keyed.size());
            for (const auto &kv : keyed) sorted[start[(size_t)(e_max - kv.first)]++] = kv;
            keyed.swap(sorted);
        }
    }
    // Bypass.  What the dropped product orders can cost a pair is bounded by 1.44e-9 (h2_i + h2_j), h2_s = sum of S_k^2
    // over the sample's non-zero Gram columns; the refinement flags pairs with D < 2.9e-3 (h2_i + h2_j).  To keep
    // ordinary pairs (D ~ q_i + q_j) a factor ~10 clear of that line, the columns contributing most to sum_s h2_s
    // (S_k^2 nnz_k) are computed in float64 instead, until the mean h2 is below 30 x the mean squared norm, at most
    // GRAM_DIRECT_ROWS of them.  `keyed` is sorted by decreasing scale.
    {
        double sum_h = 0.0, sum_e = 0.0;
        std::vector<std::pair<double, size_t>> contrib;
        contrib.reserve(keyed.size());
        for (size_t i = 0; i < keyed.size(); i++) {
            const int32_t k = keyed[i].second;
            const double S = std::ldexp(1.0, keyed[i].first), h = S * S * (double)nnz[k];
            contrib.emplace_back(h, i);
            sum_h += h;
            sum_e += (double)energy[k];
        }
        // only the largest GRAM_DIRECT_ROWS (or the test hook's count) contributions can be selected
        const size_t top = std::min<size_t>(contrib.size(), (size_t)std::max(GRAM_DIRECT_ROWS,
                                                getenv("FUF_GRAM_DIRECT") ? atoi(getenv("FUF_GRAM_DIRECT")) : 0));
        std::partial_sort(contrib.begin(), contrib.begin() + top, contrib.end(),
                          [](const std::pair<double, size_t> &a, const std::pair<double, size_t> &b) {
                              return a.first > b.first || (a.first == b.first && a.second < b.second);
                          });
        int32_t max_direct = GRAM_DIRECT_ROWS;
        if (getenv("FUF_GRAM_DIRECT")) max_direct = std::max(0, atoi(getenv("FUF_GRAM_DIRECT")));
        std::vector<uint8_t> bypass(keyed.size(), 0);
        int32_t n_sel = 0;
        for (size_t u = 0; u < contrib.size() && n_sel < max_direct && sum_h > 30.0 * sum_e; u++) {
            bypass[contrib[u].second] = 1;   // (its energy stays in sum_e: D does not shrink when a column is bypassed)
            sum_h -= contrib[u].first;
            n_sel++;
        }
        if (getenv("FUF_GRAM_DIRECT") && max_direct > 0 && getenv("FUF_GRAM_DIRECT_FORCE"))  // test hook: exactly max_direct
            for (size_t u = 0; u < contrib.size() && n_sel < max_direct; u++)
                if (!bypass[contrib[u].second]) bypass[contrib[u].second] = 1, n_sel++;
        std::vector<std::pair<int32_t, int32_t>> rest;
        rest.reserve(keyed.size());
        for (size_t i = 0; i < keyed.size(); i++) (bypass[i] ? direct_keyed : rest).push_back(keyed[i]);
        keyed.swap(rest);
    }
    // Small buckets are expensive: every bucket ends with a drain of the four accumulators during which the tensor
    // pipe idles (~3 us, the time of 250 rows of MMAs), and a bucket is padded to whole 64-row stages.  A bucket with
    // fewer than 1,024 columns therefore joins the next LARGER scale (coarser digits for its columns, still covered by
    // the h2 bound, which is computed from the scale a column is actually quantised with) whenever that raises
    // sum_s h2_s by less than 1 % — in practice the tail of tiny-scale columns (zero-length edges), never the few
    // large upper-level columns.  `keyed` is sorted by decreasing exponent.
    if (!getenv("FUF_GRAM_NO_MERGE")) {
        double total_h = 0.0;
        for (const auto &kv : keyed) total_h += std::ldexp(1.0, 2 * kv.first) * (double)nnz[kv.second];
        const double allowance = 0.01 * total_h;
        double spent = 0.0;
        size_t end = keyed.size();
        while (end > 0) {
            size_t begin = end;
            const int32_t e = keyed[end - 1].first;
            while (begin > 0 && keyed[begin - 1].first == e) begin--;
            if (begin == 0) break;                                  // the largest scale has nowhere to go
            const int32_t e_up = keyed[begin - 1].first;            // next larger exponent present
            if (end - begin < 1024) {
                double cost = 0.0;
                for (size_t i = begin; i < end; i++)
                    cost += (std::ldexp(1.0, 2 * e_up) - std::ldexp(1.0, 2 * e)) * (double)nnz[keyed[i].second];
                if (spent + cost <= allowance) {
                    spent += cost;
                    for (size_t i = begin; i < end; i++) keyed[i].first = e_up;   // stays sorted: joins the bucket above
                    continue;                                       // re-examine the enlarged bucket [.., end)
                }
            }
            end = begin;
        }
    }
    const int32_t n_direct = (int32_t)direct_keyed.size();
    const int32_t n_direct_pad = std::max(16, (n_direct + 15) / 16 * 16);
    std::vector<int32_t> &direct_nodes = plan.direct_nodes;
    direct_nodes.assign(n_direct_pad, -1);
    for (int32_t i = 0; i < n_direct; i++) direct_nodes[i] = direct_keyed[i].second;
    ctx->gram_bypassed = n_direct;
    plan.n_direct = n_direct;
    plan.n_direct_pad = n_direct_pad;
    std::vector<int32_t> &row_map = plan.row_map;
    std::vector<float> &inv_scale = plan.inv_scale;
    std::vector<Group> &groups = plan.groups;
    row_map.clear();
    inv_scale.clear();
    groups.clear();
    for (size_t i = 0; i < keyed.size();) {
        size_t j = i;
        // one bucket; at most 32,768 rows per group so that 4 x 127^2 x rows stays below 2^31 in accumulator 3
        while (j < keyed.size() && keyed[j].first == keyed[i].first && j - i < 32768) j++;
        const double S = std::ldexp(1.0, keyed[i].first);
        for (size_t u = i; u < j; u++) {
            row_map.push_back(keyed[u].second);
            inv_scale.push_back((float)(1.0 / S));
        }
        while (row_map.size() % KC) {  // pad the bucket to whole stages with zero rows
            row_map.push_back(-1);
            inv_scale.push_back(0.f);
        }
        groups.push_back(Group{(int32_t)(row_map.size() / KC), 0, S * S / (DIGIT0 * DIGIT0)});
        i = j;
    }
    if (getenv("FUF_GRAM_DEBUG")) {
        int32_t prev = 0;
        for (size_t g = 0; g < groups.size(); g++) {
            fprintf(stderr, "gram bucket %zu: %d rows (w0 %.3g)\n", g, (groups[g].stage_end - prev) * KC, groups[g].w0);
            prev = groups[g].stage_end;
        }
        fprintf(stderr, "gram: %d float64 columns\n", n_direct);
    }
    if (groups.empty()) {  // every profile is identically zero: one all-zero stage
        row_map.assign(KC, -1);
        inv_scale.assign(KC, 0.f);
        groups.push_back(Group{1, 0, 0.0});
    }
    plan.n_rows = (int32_t)row_map.size();
    plan.parts.clear();          // <= 512 rows each, never across a bucket boundary
    {
        int32_t r0 = 0;
        for (size_t g = 0; g < groups.size(); g++) {
            const int32_t r_end = groups[g].stage_end * KC;
            for (int32_t r = r0; r < r_end; r += 512) plan.parts.push_back(Part{r, std::min(r_end, r + 512), (int32_t)g, 0});
            r0 = r_end;
        }
    }
    ctx->gram_groups = (int32_t)groups.size();
    return FUF_OK;
}

}  // namespace gram

// Items (row tile x column-tile pair) of one launch of the CTA-pair kernel, in processing order, and — when n_seg_wanted
// > 0 (full matrix only) — their completion segments: equally many tile rows each; `segs` receives the row boundary and
// the counter target of every segment.
//
// L2 rasterisation (as in pairwise_f32): the 74 CTA pairs in flight stream their digit-plane panels at the same pace, so
// a panel shared by several of their items is fetched from DRAM once.  Row-major order put one tile row and 74 column
// pairs into a wave — 149 panels of 17 MB, 1.15 TB of DRAM reads per launch at N = 50,000 (ncu: 4.5 TB/s, L2 hit 38 %) —;
// strips of ~12 tile rows walked column pair by column pair make it ~12 rows x 6 pairs = 24 panels.  With completion
// segments a strip is a whole number of segments.
static void plan_gram_items(int32_t tiles_per_dim, const int32_t *tile_rows_h, int32_t n_trows, int32_t n_seg_wanted,
                            std::vector<gram::PairTile> &ptiles, GramSegments *segs)
{
    using namespace gram;
    const bool full = tile_rows_h == nullptr;
    int32_t per = 0;   // tile rows per completion segment
    if (n_seg_wanted > 0 && full) {
        const int32_t n_seg = std::min(std::min(n_seg_wanted, 4000), n_trows);
        per = (n_trows + n_seg - 1) / n_seg;
    }
    if (full) {
        const int32_t S = per > 0 ? per * ((GRAM_STRIP + per - 1) / per) : GRAM_STRIP;
        for (int32_t r0 = 0; r0 < n_trows; r0 += S)
            for (int32_t b = r0 / 2; 2 * b < tiles_per_dim; b++)
                for (int32_t t = r0; t < std::min(n_trows, r0 + S); t++)
                    if (b >= t / 2) ptiles.push_back(PairTile{t, b, t * TILE, 0});
    } else {
        for (int32_t c = 0; c < n_trows; c++) {
            const int32_t t = tile_rows_h[c];
            for (int32_t b = t / 2; 2 * b < tiles_per_dim; b++) ptiles.push_back(PairTile{t, b, c * TILE, 0});
        }
    }
    if (per > 0 && segs) {
        // once segment s (and every earlier one) is counted complete, rows [row_end[s-1], row_end[s]) x 128 of D are final
        // in ALL columns — right of the diagonal by the segment's own tiles, left of it by the mirror stores of the earlier
        // segments — and leave as one contiguous copy
        for (int32_t t0 = 0; t0 < n_trows; t0 += per) {
            segs->row_end.push_back(std::min(n_trows, t0 + per));
            segs->expected.push_back(0);
        }
        for (PairTile &pt : ptiles) {
            const int32_t seg = pt.ti / per;
            pt.pad = (seg + 1) << 12;
            segs->expected[seg] += GRAM_SEG_UNITS;
        }
    }
}

// Test hook, host only: the items of a full-matrix launch for `tiles_per_dim` tile rows as (ti, b, segment) rows.
extern "C" int fuf_debug_gram_items(int32_t tiles_per_dim, int32_t n_seg_wanted, int32_t *rows_out, int32_t max_rows,
                                    int32_t *n_rows, int32_t *seg_row_end, int32_t *seg_expected, int32_t max_seg, int32_t *n_seg)
{
    FUF_REQUIRE(tiles_per_dim > 0 && n_rows && n_seg, "fuf_debug_gram_items: bad argument");
    std::vector<gram::PairTile> ptiles;
    GramSegments segs;
    plan_gram_items(tiles_per_dim, nullptr, tiles_per_dim, n_seg_wanted, ptiles, &segs);
    *n_rows = (int32_t)ptiles.size();
    *n_seg = (int32_t)segs.row_end.size();
    if (rows_out) {
        FUF_REQUIRE(max_rows >= *n_rows && max_seg >= *n_seg, "fuf_debug_gram_items: %d rows, %d segments needed", *n_rows, *n_seg);
        for (size_t i = 0; i < ptiles.size(); i++) {
            rows_out[3 * i] = ptiles[i].ti;
            rows_out[3 * i + 1] = ptiles[i].b;
            rows_out[3 * i + 2] = (ptiles[i].pad >> 12) - 1;
        }
        for (int32_t q = 0; q < *n_seg; q++) {
            if (seg_row_end) seg_row_end[q] = segs.row_end[q];
            if (seg_expected) seg_expected[q] = (int32_t)segs.expected[q];
        }
    }
    return FUF_OK;
}

int pairwise_gram(fuf_ctx *ctx, const float *PT, int64_t N, int64_t ldp, int64_t shard, int64_t shard_stride,
                  const double *err_in, double *h2, double *err_out, const int32_t *tile_rows_h, int32_t n_tile_rows,
                  double *D, int64_t ldd, cudaStream_t stream, GramSegments *segs)
{
    using namespace gram;
    if (segs) segs->row_end.clear(), segs->expected.clear();
    if (N == 0) return FUF_OK;
    const int32_t n = ctx->n;
    const int64_t tiles_per_dim = (N + TILE - 1) / TILE;
    FUF_REQUIRE(((uintptr_t)PT & 15) == 0, "fuf_pairwise_gram: PT must be 16-byte aligned");
    Lay lay;
    int64_t n_cols;  // sample slots of the digit planes
    if (shard == 0) {
        FUF_REQUIRE(ldp % 4 == 0 && ldp >= N, "fuf_pairwise_gram: ldp=%lld must be a multiple of 4 and >= N=%lld",
                    (long long)ldp, (long long)N);
        n_cols = tiles_per_dim * TILE;
        lay = Lay{ldp, 0, ldp};
    } else {
        FUF_REQUIRE(shard % TILE == 0 && shard > 0 && shard_stride >= shard * (int64_t)n && shard_stride % 4 == 0,
                    "fuf_pairwise_gram: bad shard layout (shard=%lld stride=%lld)", (long long)shard,
                    (long long)shard_stride);
        n_cols = (N + shard - 1) / shard * shard;
        lay = Lay{shard, shard_stride, shard};
    }
    std::vector<Tile> tiles;
    const bool full = (tile_rows_h == nullptr);
    const int32_t n_trows = full ? (int32_t)tiles_per_dim : n_tile_rows;
    for (int32_t c = 0; c < n_trows; c++) {
        const int32_t t = full ? c : tile_rows_h[c];
        FUF_REQUIRE(t >= 0 && t < tiles_per_dim, "fuf_pairwise_gram: tile row %d out of range [0,%lld)", t,
                    (long long)tiles_per_dim);
        for (int32_t tj = t; tj < tiles_per_dim; tj++) tiles.push_back(Tile{t, tj, full ? t * TILE : c * TILE, 0});
    }
    if (tiles.empty()) return FUF_OK;

    // ---- column maxima -> buckets (powers of four), permutation, stage table --------------------------------
    int rc = ctx->ws_gram_small.reserve(3 * (size_t)n * sizeof(float));
    if (rc) return rc;
    float *d_cmax = (float *)ctx->ws_gram_small.ptr;
    gram_colmax_kernel<<<(unsigned)((n + 7) / 8), 256, 0, stream>>>(PT, lay, N, n, d_cmax, d_cmax + n, d_cmax + 2 * n);
    FUF_CUDA(cudaGetLastError());
    std::vector<float> cstat(3 * (size_t)n);
    FUF_CUDA(cudaMemcpyAsync(cstat.data(), d_cmax, 3 * (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, stream));
    FUF_CUDA(cudaStreamSynchronize(stream));
    const auto t_plan0 = std::chrono::steady_clock::now();
    GramPlanHost plan;
    if ((rc = gram_make_plan(ctx, cstat, n, plan))) return rc;
    std::vector<int32_t> &row_map = plan.row_map, &direct_nodes = plan.direct_nodes;
    std::vector<float> &inv_scale = plan.inv_scale;
    std::vector<Group> &groups = plan.groups;
    const int32_t n_direct = plan.n_direct, n_direct_pad = plan.n_direct_pad;
    const int32_t n_rows = (int32_t)row_map.size();
    Lay ws = shard == 0 ? Lay{n_cols, 0, n_cols} : Lay{shard, (int64_t)n_rows * shard, shard};

    // ---- workspace: four digit planes | partial statistics | row map, scales, groups, tiles -----------------
    auto up = [](size_t b) { return (b + 1023) & ~(size_t)1023; };
    const size_t plane = up((size_t)n_rows * n_cols);
    std::vector<Part> &parts = plan.parts;
    const int32_t n_parts = (int32_t)parts.size();
    const size_t part_bytes = up((size_t)n_parts * n_cols * sizeof(double));
    const size_t cself_bytes = up((size_t)n_parts * 4 * n_cols * sizeof(int32_t));
    const size_t parts_bytes = up(parts.size() * sizeof(Part));
    const size_t map_bytes = up((size_t)n_rows * sizeof(int32_t)), scl_bytes = up((size_t)n_rows * sizeof(float));
    const size_t grp_bytes = up(groups.size() * sizeof(Group)), tiles_bytes = up(tiles.size() * sizeof(Tile));
    const size_t sub_bytes = up((size_t)n_direct_pad * n_cols * sizeof(double)), dn_bytes = up((size_t)n_direct_pad * sizeof(int32_t));
    const size_t q_bytes = up((size_t)n_cols * sizeof(double));
    if ((rc = ctx->ws_gram.reserve(NDIG * plane + 2 * part_bytes + map_bytes + scl_bytes + grp_bytes + tiles_bytes + sub_bytes +
                                   dn_bytes + q_bytes + cself_bytes + parts_bytes)))
        return rc;
    char *wsp = (char *)ctx->ws_gram.ptr;
    int8_t *D0 = (int8_t *)wsp, *D1 = D0 + plane, *D2 = D1 + plane, *D3 = D2 + plane;
    double *r2part = (double *)(wsp + NDIG * plane), *h2part = (double *)((char *)r2part + part_bytes);
    int32_t *d_map = (int32_t *)((char *)h2part + part_bytes);
    float *d_scl = (float *)((char *)d_map + map_bytes);
    Group *d_groups = (Group *)((char *)d_scl + scl_bytes);
    Tile *d_tiles = (Tile *)((char *)d_groups + grp_bytes);
    double *SUB = (double *)((char *)d_tiles + tiles_bytes);
    int32_t *d_dn = (int32_t *)((char *)SUB + sub_bytes);
    double *q = (double *)((char *)d_dn + dn_bytes);  // squared norms of the quantised Gram part (kernel-internal)
    int32_t *cself = (int32_t *)((char *)q + q_bytes);
    Part *d_parts = (Part *)((char *)cself + cself_bytes);
    FUF_CUDA(cudaMemcpyAsync(d_parts, parts.data(), parts.size() * sizeof(Part), cudaMemcpyHostToDevice, stream));
    FUF_CUDA(cudaMemcpyAsync(d_dn, direct_nodes.data(), (size_t)n_direct_pad * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
    FUF_CUDA(cudaMemcpyAsync(d_map, row_map.data(), (size_t)n_rows * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
    FUF_CUDA(cudaMemcpyAsync(d_scl, inv_scale.data(), (size_t)n_rows * sizeof(float), cudaMemcpyHostToDevice, stream));
    FUF_CUDA(cudaMemcpyAsync(d_groups, groups.data(), groups.size() * sizeof(Group), cudaMemcpyHostToDevice, stream));
    FUF_CUDA(cudaMemcpyAsync(d_tiles, tiles.data(), tiles.size() * sizeof(Tile), cudaMemcpyHostToDevice, stream));
    FUF_CUDA(cudaStreamSynchronize(stream));  // the host vectors are pageable and die with this call
    if (getenv("FUF_GRAM_DEBUG"))
        fprintf(stderr, "gram: host planning + table upload %.3f ms\n",
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_plan0).count());
    const Lay sub = shard == 0 ? Lay{n_cols, 0, n_cols} : Lay{shard, (int64_t)n_direct_pad * shard, shard};
    {
        dim3 g((unsigned)((n_cols / 4 + 127) / 128), (unsigned)n_parts);
        gram_digits_kernel<<<g, 128, 0, stream>>>(PT, lay, ws, n_cols, N, d_parts, d_map, d_scl, D0, D1, D2, D3, r2part,
                                                  h2part, cself);
        FUF_CUDA(cudaGetLastError());
        gram_stat_kernel<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(r2part, h2part, cself, d_parts, n_parts, d_groups,
                                                                          n_cols, err_in, N, h2, err_out, q);
        FUF_CUDA(cudaGetLastError());
        dim3 gg((unsigned)((n_cols + 255) / 256), (unsigned)n_direct_pad);
        if (n_direct > 0) {
            gram_gather_kernel<<<gg, 256, 0, stream>>>(PT, lay, sub, n_cols, N, d_dn, SUB);
            FUF_CUDA(cudaGetLastError());
            ctx->launches += 1;
        }
    }
    ctx->launches += 3;   // colmax, digits, stat
    CUtensorMap map0, map1, map2, map3;
    if ((rc = make_map(&map0, D0, N, n_rows, ws)) || (rc = make_map(&map1, D1, N, n_rows, ws)) ||
        (rc = make_map(&map2, D2, N, n_rows, ws)) || (rc = make_map(&map3, D3, N, n_rows, ws)))
        return rc;
    Params p;
    p.tiles = d_tiles;
    p.groups = d_groups;
    p.n_tiles = (int32_t)tiles.size();
    p.n_groups = (int32_t)groups.size();
    p.S = (int32_t)ws.S;
    p.N = N;
    p.q = q;
    p.D = D;
    p.ldd = ldd;
    p.mirror = full;
    p.n_direct = n_direct;
    p.SUB = SUB;
    p.sub = sub;
    if (!ctx->ev_gram[0]) {
        FUF_CUDA(cudaEventCreate(&ctx->ev_gram[0]));
        FUF_CUDA(cudaEventCreate(&ctx->ev_gram[1]));
    }
    const char *pair_env = getenv("FUF_GRAM_2CTA");
    const bool use_pairs = pair_env ? atoi(pair_env) != 0 : true;   // default: CTA pairs (FUF_GRAM_2CTA=0: one CTA per tile)
    if (use_pairs) {
        // CTA pairs: work item = row tile x two column tiles (2b, 2b+1); the half below the diagonal is computed and dropped
        std::vector<PairTile> ptiles;
        plan_gram_items((int32_t)tiles_per_dim, full ? nullptr : tile_rows_h, n_trows,
                        (segs && segs->seg_done && full) ? segs->n_wanted : 0, ptiles, segs);
        static_assert(PLANE_BYTES == 8192 && TILE == 128, "umma_block10 hard-codes the plane and accumulator strides");
        static_assert(KC == 2 * KB, "the issue loops are written for two 32-row MMA blocks per stage");
        static_assert(sizeof(PairTile) == sizeof(Tile), "the pair tiles reuse the tile table's room");
        FUF_REQUIRE(ptiles.size() <= tiles.size(), "fuf_pairwise_gram: pair-tile table larger than the tile table");
        FUF_CUDA(cudaMemcpyAsync(d_tiles, ptiles.data(), ptiles.size() * sizeof(PairTile), cudaMemcpyHostToDevice, stream));
        FUF_CUDA(cudaStreamSynchronize(stream));
        CUtensorMap mb0, mb1, mb2, mb3;
        if ((rc = make_map(&mb0, D0, N, n_rows, ws, 64)) || (rc = make_map(&mb1, D1, N, n_rows, ws, 64)) ||
            (rc = make_map(&mb2, D2, N, n_rows, ws, 64)) || (rc = make_map(&mb3, D3, N, n_rows, ws, 64)))
            return rc;
        Params2 p2;
        p2.tiles = (const PairTile *)d_tiles;
        p2.groups = d_groups;
        p2.n_tiles = (int32_t)ptiles.size();
        p2.n_groups = p.n_groups;
        p2.S = p.S;
        p2.tiles_per_dim = (int32_t)tiles_per_dim;
        p2.N = N;
        p2.q = q;
        p2.D = D;
        p2.ldd = ldd;
        p2.mirror = full;
        p2.n_direct = n_direct;
        p2.SUB = SUB;
        p2.sub = sub;
        p2.dep_flags = nullptr;
        p2.epoch = 0;
        p2.err = p2.h2 = nullptr;
        p2.kappa_r = p2.kappa_l = 0.0;
        p2.flag_list = nullptr;
        p2.list_cap = 0;
        p2.seg_done = (segs && !segs->row_end.empty()) ? segs->seg_done : nullptr;
        FUF_CUDA(cudaFuncSetAttribute(gram_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM2_BYTES));
        const int n_pairs = (int)std::min<int64_t>(p2.n_tiles, ctx->sm_count / 2);
        FUF_CUDA(cudaEventRecord(ctx->ev_gram[0], stream));
        gram_pair_kernel<<<2 * n_pairs, NTHREADS, SMEM2_BYTES, stream>>>(map0, map1, map2, map3, mb0, mb1, mb2, mb3, p2);
        FUF_CUDA(cudaGetLastError());
        FUF_CUDA(cudaEventRecord(ctx->ev_gram[1], stream));
        ctx->gram_flops = 10.0 * 2.0 * (2.0 * TILE) * TILE * (double)n_rows * (double)p2.n_tiles;
    } else {
        FUF_CUDA(cudaFuncSetAttribute(gram_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        const int grid = std::min<int64_t>(p.n_tiles, ctx->sm_count);
        FUF_CUDA(cudaEventRecord(ctx->ev_gram[0], stream));
        gram_tile_kernel<<<grid, NTHREADS, SMEM_BYTES, stream>>>(map0, map1, map2, map3, p);
        FUF_CUDA(cudaGetLastError());
        FUF_CUDA(cudaEventRecord(ctx->ev_gram[1], stream));
        ctx->gram_flops = 10.0 * 2.0 * TILE * TILE * (double)n_rows * (double)p.n_tiles;  // int8 multiply-adds x 2
    }
    ctx->gram_groups = p.n_groups;
    ctx->launches += 1;
    return FUF_OK;
}

}  // namespace fuf

using namespace fuf;

// ---------------------------------------------------------------------------------------------------------------------
// Sharded --L2 job (several GPUs of one node): the same computation as fuf_pairwise_gram in phases, so that every rank
// quantises only its own shard and pulls the digit planes of the shards its tiles need from its peers.
//   fuf_gram_colstats   column statistics of the own shard                        (all-reduce them: max | sum | sum)
//   fuf_gram_plan       the plan from the reduced statistics (identical on every rank); keeps its tables in the context
//   fuf_gram_digits     digit planes, bypass rows and per-sample statistics of the own shard
//   fuf_gram_tiles      the CTA-pair tile kernel on an explicit list of (row tile, column-tile pair, dependency) items,
//                       every tile stored in place into the full matrix, refinement flags from the epilogue
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int fuf_gram_colstats(fuf_ctx *ctx, const float *PT_shard, int64_t n_valid, int64_t ld, float *stats, void *stream)
{
    using namespace gram;
    FUF_REQUIRE(ctx && PT_shard && stats, "fuf_gram_colstats: null argument");
    FUF_REQUIRE(n_valid >= 0 && ld >= n_valid && ld % 4 == 0 && ((uintptr_t)PT_shard & 15) == 0, "fuf_gram_colstats: bad shape");
    FUF_CUDA(cudaSetDevice(ctx->device));
    const int32_t n = ctx->n;
    gram_colmax_kernel<<<(unsigned)((n + 7) / 8), 256, 0, (cudaStream_t)stream>>>(PT_shard, Lay{ld, 0, ld}, n_valid, n, stats,
                                                                                 stats + n, stats + 2 * n);
    FUF_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return FUF_OK;
}

extern "C" int fuf_gram_plan(fuf_ctx *ctx, const float *stats, int32_t *n_rows, int32_t *n_direct_pad, void *stream)
{
    using namespace gram;
    FUF_REQUIRE(ctx && stats && n_rows && n_direct_pad, "fuf_gram_plan: null argument");
    FUF_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int32_t n = ctx->n;
    std::vector<float> cstat(3 * (size_t)n);
    FUF_CUDA(cudaMemcpyAsync(cstat.data(), stats, cstat.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
    FUF_CUDA(cudaStreamSynchronize(st));
    GramPlanHost plan;
    int rc = gram_make_plan(ctx, cstat, n, plan);
    if (rc) return rc;
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t map_b = up(plan.row_map.size() * sizeof(int32_t)), scl_b = up(plan.inv_scale.size() * sizeof(float));
    const size_t grp_b = up(plan.groups.size() * sizeof(Group)), part_b = up(plan.parts.size() * sizeof(Part));
    const size_t dn_b = up(plan.direct_nodes.size() * sizeof(int32_t));
    if ((rc = ctx->ws_gram_tab.reserve(map_b + scl_b + grp_b + part_b + dn_b))) return rc;
    char *w = (char *)ctx->ws_gram_tab.ptr;
    GramJob &job = ctx->gram_job;
    job.d_map = (int32_t *)w;
    job.d_scl = (float *)(w + map_b);
    job.d_groups = w + map_b + scl_b;
    job.d_parts = w + map_b + scl_b + grp_b;
    job.d_dn = (int32_t *)(w + map_b + scl_b + grp_b + part_b);
    FUF_CUDA(cudaMemcpyAsync(job.d_map, plan.row_map.data(), plan.row_map.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    FUF_CUDA(cudaMemcpyAsync(job.d_scl, plan.inv_scale.data(), plan.inv_scale.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    FUF_CUDA(cudaMemcpyAsync(job.d_groups, plan.groups.data(), plan.groups.size() * sizeof(Group), cudaMemcpyHostToDevice, st));
    FUF_CUDA(cudaMemcpyAsync(job.d_parts, plan.parts.data(), plan.parts.size() * sizeof(Part), cudaMemcpyHostToDevice, st));
    FUF_CUDA(cudaMemcpyAsync(job.d_dn, plan.direct_nodes.data(), plan.direct_nodes.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    FUF_CUDA(cudaStreamSynchronize(st));   // the host vectors are pageable and die with this call
    job.n_rows = plan.n_rows;
    job.n_direct = plan.n_direct;
    job.n_direct_pad = plan.n_direct_pad;
    job.n_parts = (int32_t)plan.parts.size();
    job.n_groups = (int32_t)plan.groups.size();
    *n_rows = plan.n_rows;
    *n_direct_pad = plan.n_direct_pad;
    return FUF_OK;
}

extern "C" int fuf_gram_digits(fuf_ctx *ctx, const float *PT_shard, int64_t n_valid, int64_t shard, int8_t *planes,
                               int64_t plane_stride, int64_t shard_stride, int32_t shard_index, const double *err_in,
                               double *q, double *h2, double *err_out, double *SUB, int64_t sub_shard_stride, void *stream)
{
    using namespace gram;
    FUF_REQUIRE(ctx && PT_shard && planes && q && h2 && err_out && SUB, "fuf_gram_digits: null argument");
    const GramJob &job = ctx->gram_job;
    FUF_REQUIRE(job.n_rows > 0, "fuf_gram_digits: no plan (call fuf_gram_plan first)");
    FUF_REQUIRE(shard > 0 && shard % TILE == 0 && n_valid >= 0 && n_valid <= shard && shard_index >= 0, "fuf_gram_digits: bad shape");
    FUF_REQUIRE(shard_stride >= (int64_t)job.n_rows * shard && plane_stride >= shard_stride &&
                    sub_shard_stride >= (int64_t)job.n_direct_pad * shard,
                "fuf_gram_digits: plane / bypass buffers too small for the plan (%d rows, %d bypass rows)", job.n_rows, job.n_direct_pad);
    FUF_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    auto up = [](size_t b) { return (b + 1023) & ~(size_t)1023; };
    const size_t part_bytes = up((size_t)job.n_parts * shard * sizeof(double));
    const size_t cself_bytes = up((size_t)job.n_parts * 4 * shard * sizeof(int32_t));
    int rc = ctx->ws_gram.reserve(2 * part_bytes + cself_bytes);
    if (rc) return rc;
    double *r2part = (double *)ctx->ws_gram.ptr, *h2part = (double *)((char *)r2part + part_bytes);
    int32_t *cself = (int32_t *)((char *)h2part + part_bytes);
    const Lay one{shard, 0, shard};
    int8_t *D0 = planes + (int64_t)shard_index * shard_stride;
    dim3 g((unsigned)((shard / 4 + 127) / 128), (unsigned)job.n_parts);
    gram_digits_kernel<<<g, 128, 0, st>>>(PT_shard, one, one, shard, n_valid, (const Part *)job.d_parts, job.d_map, job.d_scl, D0,
                                          D0 + plane_stride, D0 + 2 * plane_stride, D0 + 3 * plane_stride, r2part, h2part, cself);
    FUF_CUDA(cudaGetLastError());
    if (n_valid > 0) {
        gram_stat_kernel<<<(unsigned)((n_valid + 255) / 256), 256, 0, st>>>(r2part, h2part, cself, (const Part *)job.d_parts, job.n_parts,
                                                                            (const Group *)job.d_groups, shard, err_in, n_valid, h2,
                                                                            err_out, q);
        FUF_CUDA(cudaGetLastError());
    }
    dim3 gg((unsigned)((shard + 255) / 256), (unsigned)job.n_direct_pad);
    gram_gather_kernel<<<gg, 256, 0, st>>>(PT_shard, one, one, shard, n_valid, job.d_dn, SUB + (int64_t)shard_index * sub_shard_stride);
    FUF_CUDA(cudaGetLastError());
    ctx->launches += 3;
    return FUF_OK;
}

extern "C" int fuf_gram_tiles(fuf_ctx *ctx, const int8_t *planes, int64_t plane_stride, int64_t shard_stride, int64_t N,
                              int64_t shard, const int32_t *tiles_h, int32_t n_tiles, const uint32_t *dep_flags, uint32_t epoch,
                              const double *q, const double *h2, const double *err, const double *SUB,
                              int64_t sub_shard_stride, double *D, int64_t ldd, int64_t *flag_list, int64_t list_cap,
                              const int32_t *seg_h, uint32_t *seg_done, void *stream)
{
    using namespace gram;
    FUF_REQUIRE(ctx && planes && q && D && (tiles_h || n_tiles == 0), "fuf_gram_tiles: null argument");
    const GramJob &job = ctx->gram_job;
    FUF_REQUIRE(job.n_rows > 0, "fuf_gram_tiles: no plan (call fuf_gram_plan first)");
    FUF_REQUIRE(N > 0 && ldd >= N && shard > 0 && shard % (2 * TILE) == 0, "fuf_gram_tiles: bad shape (shards are multiples of 256)");
    FUF_REQUIRE(!flag_list || (h2 && err), "fuf_gram_tiles: flag_list needs the statistics");
    if (n_tiles == 0) return FUF_OK;
    FUF_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t tiles_per_dim = (N + TILE - 1) / TILE;
    std::vector<PairTile> ptiles((size_t)n_tiles);
    for (int32_t c = 0; c < n_tiles; c++) {
        const int32_t ti = tiles_h[3 * c], b = tiles_h[3 * c + 1], dep = tiles_h[3 * c + 2];
        FUF_REQUIRE(ti >= 0 && ti < tiles_per_dim && b >= ti / 2 && 2 * b < tiles_per_dim, "fuf_gram_tiles: item (%d, %d) outside the upper triangle", ti, b);
        FUF_REQUIRE(dep >= -1 && dep < 4095 && (dep < 0 || dep_flags), "fuf_gram_tiles: bad dependency %d", dep);
        const int32_t seg = (seg_h && seg_done) ? seg_h[c] : -1;
        FUF_REQUIRE(seg >= -1 && seg < (1 << 19), "fuf_gram_tiles: bad segment %d", seg);
        ptiles[c] = PairTile{ti, b, ti * TILE, (dep + 1) | ((seg + 1) << 12)};
    }
    PairTile *d_tiles = nullptr;
    int rc = ctx->tables.get_by_content("gtiles", ptiles.data(), ptiles.size() * sizeof(PairTile), (void **)&d_tiles);
    if (rc) return rc;
    const Lay ws{shard, shard_stride, shard};
    CUtensorMap ma[4], mb[4];
    for (int pl = 0; pl < 4; pl++)
        if ((rc = make_map(&ma[pl], planes + pl * plane_stride, N, job.n_rows, ws)) ||
            (rc = make_map(&mb[pl], planes + pl * plane_stride, N, job.n_rows, ws, 64)))
            return rc;
    Params2 p2;
    p2.tiles = d_tiles;
    p2.groups = (const Group *)job.d_groups;
    p2.n_tiles = n_tiles;
    p2.n_groups = job.n_groups;
    p2.S = (int32_t)shard;
    p2.tiles_per_dim = (int32_t)tiles_per_dim;
    p2.N = N;
    p2.q = q;
    p2.D = D;
    p2.ldd = ldd;
    p2.mirror = 1;
    p2.n_direct = job.n_direct;
    p2.SUB = SUB;
    p2.sub = Lay{shard, sub_shard_stride, shard};
    p2.dep_flags = dep_flags;
    p2.epoch = epoch;
    p2.err = err;
    p2.h2 = h2;
    p2.kappa_r = 1.6e13;
    p2.kappa_l = 2.9e-3;
    p2.flag_list = (long long *)flag_list;
    p2.list_cap = list_cap;
    p2.seg_done = (seg_h && seg_done) ? seg_done : nullptr;
    if (!ctx->ev_gram[0]) {
        FUF_CUDA(cudaEventCreate(&ctx->ev_gram[0]));
        FUF_CUDA(cudaEventCreate(&ctx->ev_gram[1]));
    }
    FUF_CUDA(cudaFuncSetAttribute(gram_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM2_BYTES));
    const int n_pairs = (int)std::min<int64_t>(n_tiles, ctx->sm_count / 2);
    FUF_CUDA(cudaEventRecord(ctx->ev_gram[0], st));
    gram_pair_kernel<<<2 * n_pairs, NTHREADS, SMEM2_BYTES, st>>>(ma[0], ma[1], ma[2], ma[3], mb[0], mb[1], mb[2], mb[3], p2);
    FUF_CUDA(cudaGetLastError());
    FUF_CUDA(cudaEventRecord(ctx->ev_gram[1], st));
    ctx->gram_flops = 10.0 * 2.0 * (2.0 * TILE) * TILE * (double)job.n_rows * (double)n_tiles;
    ctx->launches += 1;
    return FUF_OK;
}

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

extern "C" int fuf_i8_mma_peak(fuf_ctx *ctx, int cta_group, int32_t iters, int32_t launches, float *ms,
                               double *executed_ops, void *stream)
{
    using namespace gram;
    FUF_REQUIRE(ctx && ms && executed_ops, "fuf_i8_mma_peak: null argument");
    FUF_REQUIRE((cta_group == 1 || cta_group == 2) && iters > 0 && launches > 0, "fuf_i8_mma_peak: bad argument");
    FUF_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (!ctx->ev_gram[0]) {
        FUF_CUDA(cudaEventCreate(&ctx->ev_gram[0]));
        FUF_CUDA(cudaEventCreate(&ctx->ev_gram[1]));
    }
    const int grid = cta_group == 1 ? ctx->sm_count : ctx->sm_count / 2 * 2;
    const uint32_t smem = 200u << 10;   // operands need 48-64 KB; asking for 200 KB keeps it at one CTA (512 TMEM columns) per SM
    if (cta_group == 1)
        FUF_CUDA(cudaFuncSetAttribute(i8_mma_peak_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else
        FUF_CUDA(cudaFuncSetAttribute(i8_mma_peak_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FUF_CUDA(cudaEventRecord(ctx->ev_gram[0], st));
    for (int32_t l = 0; l < launches; l++) {
        if (cta_group == 1) {
            i8_mma_peak_kernel<1><<<grid, 128, smem, st>>>(iters);
        } else {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(grid);
            cfg.blockDim = dim3(128);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = st;
            cudaLaunchAttribute attr;
            attr.id = cudaLaunchAttributeClusterDimension;
            attr.val.clusterDim.x = 2;
            attr.val.clusterDim.y = 1;
            attr.val.clusterDim.z = 1;
            cfg.attrs = &attr;
            cfg.numAttrs = 1;
            FUF_CUDA(cudaLaunchKernelEx(&cfg, i8_mma_peak_kernel<2>, iters));
        }
    }
    FUF_CUDA(cudaGetLastError());
    FUF_CUDA(cudaEventRecord(ctx->ev_gram[1], st));
    FUF_CUDA(cudaEventSynchronize(ctx->ev_gram[1]));
    FUF_CUDA(cudaEventElapsedTime(ms, ctx->ev_gram[0], ctx->ev_gram[1]));
    ctx->launches += launches;
    // 20 MMAs per iteration; one MMA = M x N x K multiply-adds = 2 M N K ops (M = 128 per CTA either way)
    *executed_ops = (double)launches * (double)iters * 20.0 * 2.0 * TILE * TILE * KB * (double)grid;
    return FUF_OK;
}

extern "C" int fuf_gram_last_info(const fuf_ctx *ctx, int32_t *n_buckets, int32_t *n_bypassed)
{
    FUF_REQUIRE(ctx && n_buckets && n_bypassed, "fuf_gram_last_info: null argument");
    *n_buckets = ctx->gram_groups;
    *n_bypassed = ctx->gram_bypassed;
    return FUF_OK;
}

extern "C" int fuf_pairwise_gram(fuf_ctx *ctx, const float *PT, int64_t N, int64_t ldp, int64_t shard,
                                 int64_t shard_stride, int mode, const double *err_in, double *h2, double *err_out,
                                 const int32_t *tile_rows_h, int32_t n_tile_rows, double *D, int64_t ldd, void *stream)
{
    FUF_REQUIRE(ctx && PT && h2 && D, "fuf_pairwise_gram: null argument");
    FUF_REQUIRE(mode == FUF_MODE_L2, "fuf_pairwise_gram: only FUF_MODE_L2 has a Gram form (got mode %d)", mode);
    FUF_REQUIRE(N >= 0 && ldd >= N, "fuf_pairwise_gram: bad shape N=%lld ldd=%lld", (long long)N, (long long)ldd);
    FUF_REQUIRE(tile_rows_h != nullptr || n_tile_rows == 0, "fuf_pairwise_gram: n_tile_rows without tile_rows_h");
    FUF_CUDA(cudaSetDevice(ctx->device));
    return pairwise_gram(ctx, PT, N, ldp, shard, shard_stride, err_in, h2, err_out, tile_rows_h, n_tile_rows, D, ldd,
                         (cudaStream_t)stream, nullptr);
}
