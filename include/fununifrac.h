/*
 * fununifrac.h — C ABI of the B200-native all-pairs FunUniFrac engine (libfununifrac.so).
 *
 * The reference (KoslickiLab/FunUniFrac) is pure Python and has no FFI of its own; these entry
 * points are what a binding for its hot path would call, one per reference function they replace
 * (paths relative to the reference repository):
 *
 *   fuf_tree_create      <- the arrays tree_to_EMDU_input builds        src/factory/make_emd_input.py:10-54
 *                           (Tint -> parent[], lint -> edge_length[], postorder, root last)
 *   fuf_ingest_gather_csv <- functional_profile_to_vector + ingest pool  src/factory/make_emd_input.py:58-122
 *   fuf_push_up          <- push_up_L1 / push_up_L2 over all samples    src/algorithms/emd_unifrac.py:20-39
 *   fuf_push_up_center   <- (no counterpart) the common vector subtracted from every pushed profile before it
 *                           is rounded to fp32; distances are invariant under it
 *   fuf_push_up_f64      <- same, float64, reference operation order    src/algorithms/emd_unifrac.py:20-39
 *   fuf_load_pushed      <- the dict of pushed vectors handed to        src/algorithms/emd_unifrac.py:42
 *                           pairwise_computation (layout change only)
 *   fuf_pairwise         <- pairwise_computation (no diffab) =          src/algorithms/emd_unifrac.py:42-54
 *                           get_L1 / get_L2 over all i<j                src/objects/profile_vector.py:9-19
 *   fuf_pairwise_gram    <- same for is_L2, as a Gram matrix on the tcgen05 tensor cores
 *   fuf_pairwise_f64     <- same, float64 validation mode
 *   fuf_refine           <- get_L1 / get_L2 in float64 for the pairs    src/objects/profile_vector.py:9-19
 *                           the fp32 operands cannot certify to 1e-6
 *   fuf_assemble, fuf_pack_tile_rows, fuf_assemble_packed
 *                        <- dists[i,j] = dists[j,i] = Z                 src/algorithms/emd_unifrac.py:51-54
 *                           (joins row-block shards computed on several GPUs)
 *   fuf_store_tile_rows_host <- same, into the host matrix np.save writes src/compute_all_pairwise_fununifrac.py:88-89
 *   fuf_diffab_block     <- get_L1_diffab / get_L2_diffab per pair in   src/objects/profile_vector.py:13-23
 *                           itertools.combinations order                src/algorithms/emd_unifrac.py:49,56-67
 *   fuf_diffab_gather    <- diffabs[np.ix_(files1, files2, nodes)] of   src/utility/differential_abundance.py:18-55
 *                           DiffabArrayIndexer, from the pushed matrix
 *   fuf_all_pairs_host   <- the compute section of the driver           src/compute_all_pairwise_fununifrac.py:64-73
 *                           (host float64 profiles in, host float64 distance matrix out)
 *
 * Conventions
 *   - plain C: pointers and sizes only; every function returns 0 on success or a negative FUF_E* code
 *     and never throws; fuf_last_error() returns a thread-local message for the last failure.
 *   - "device" pointers are CUDA device pointers on the context's device (any allocator; the Python
 *     host uses torch tensors purely as buffers).  "host" pointers are ordinary process memory.
 *   - stream is a cudaStream_t passed as void* (NULL = the legacy default stream).
 *   - one context per device and tree; contexts are independent; a context is not thread-safe.
 *
 * Layouts
 *   X   : un-pushed profiles, sample-major  X[s * ldx + k],  s in [0,N), k in [0,n)  (float or double)
 *   PT  : pushed profiles, NODE-major (transposed)  PT[k * ldp + s], float; ldp % 4 == 0 and ldp >= N.
 *         Padding columns (s >= N) are never read into a result.
 *   P64T: same layout in double (validation mode).
 *   D   : distance matrix, row-major D[i * ldd + j], double.
 */
#ifndef FUNUNIFRAC_H
#define FUNUNIFRAC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FUF_ABI_VERSION 13

/* error codes */
#define FUF_OK 0
#define FUF_EINVAL (-1)  /* bad argument (shape, alignment, mode, tree not in postorder ...) */
#define FUF_ECUDA (-2)   /* a CUDA runtime / driver call failed */
#define FUF_ENOMEM (-3)  /* device or host allocation failed */
#define FUF_ENODEV (-4)  /* no usable sm_100 device */

/* modes */
#define FUF_MODE_L1 0 /* weights L_e   ; distance sum |a-b|   */
#define FUF_MODE_L2 1 /* weights sqrt L; distance sum (a-b)^2 (squared, no root: reference quirk) */

/* dtypes for void* operands */
#define FUF_F32 0
#define FUF_F64 1

/* fuf_all_pairs_host flags */
#define FUF_AP_VALIDATE 1  /* float64 validation kernels instead of the fp32 production path */
#define FUF_AP_NO_REFINE 2 /* production path without the float64 refinement pass */

#define FUF_GRAM_MIN_SAMPLES 512 /* fuf_all_pairs_host uses the tensor-core Gram path for FUF_MODE_L2 from this N on */
#define FUF_AP_NO_GRAM 4   /* FUF_MODE_L2 through the direct (a-b)^2 kernel whatever N */

#define FUF_CENTER_SAMPLES 32 /* fuf_all_pairs_host centres on the pushed mean of the first min(N, 32) samples */

/* fuf_refine criteria */
#define FUF_CRIT_ROUNDING 0 /* err_stat = rounding statistics of fuf_push_up; L1: D < kappa (s_i+s_j), L2: D < kappa (sqrt s_i + sqrt s_j)^2 */
#define FUF_CRIT_LINEAR 1   /* D < kappa (s_i + s_j) on any per-sample statistic (the Gram path passes squared norms) */

/* fuf_pairwise flags */
#define FUF_PW_SCALAR_MATH 1 /* use scalar FADD instead of packed FADD2/FFMA2 (A/B switch for profiling) */

#define FUF_TILE 128 /* edge of a distance-matrix tile; shard boundaries must be multiples of it */

typedef struct fuf_ctx fuf_ctx;
typedef struct fuf_name_index fuf_name_index; /* node name -> basis index (host) */

typedef struct fuf_tree_info {
    int32_t n;           /* nodes incl. root */
    int32_t n_leaves;    /* nodes without children */
    int32_t depth;       /* edges on the longest root-to-leaf path */
    int32_t chunk;       /* push-up chunk length (nodes) */
    int32_t n_chunks;    /* ceil(n / chunk) */
    int32_t n_spanning;  /* nodes whose subtree crosses a chunk boundary */
    int32_t n_zero_len;  /* edges of length exactly 0 (replaced by DBL_EPSILON, emd_unifrac.py:24-25) */
    int32_t device;
} fuf_tree_info;

int fuf_version(void);
const char *fuf_last_error(void);

/* Tree context.  parent_h[i] > i for i < n-1, parent_h[n-1] == -1 (postorder, root last);
 * length_h[i] = length of edge (i, parent[i]) (length_h[n-1] ignored).  Zero lengths are replaced by
 * DBL_EPSILON exactly as push_up_L1/L2 do.  device < 0 = current device. */
int fuf_tree_create(fuf_ctx **out, int device, const int32_t *parent_h, const double *length_h, int32_t n);
int fuf_tree_destroy(fuf_ctx *ctx);
int fuf_tree_get_info(const fuf_ctx *ctx, fuf_tree_info *info);

/* Smallest legal ldp for N samples (N rounded up to a multiple of FUF_TILE). */
int64_t fuf_padded_samples(int64_t N);

/* Centre of a job: center_out (device, n doubles) = the pushed (mode weights) mean of the first m rows of X,
 * set to zero on the leaves.  Every fuf_push_up call of one job must be given the same centre (on several
 * GPUs: compute it on one rank and broadcast it).  Distances do not depend on it; accuracy does. */
int fuf_push_up_center(fuf_ctx *ctx, const void *X, int x_dtype, int64_t m, int64_t ldx, int mode, double *center_out,
                       void *stream);

/* Push-up, production path.  X (device, x_dtype, sample-major, leading dimension ldx >= n) holds N samples.
 * With S_k(s) = sum of X[s] over the subtree of k (accumulated in double), w = L (L1) or sqrt(L) (L2), w_root = 1:
 *   PT   (device float, node-major, ld ldp; may be NULL):    PT[k][col0+s]   = (float)(w_k S_k(s) - center[k])
 *   P64T (device double, node-major, ld ldp64; may be NULL): P64T[k][col0+s] = w_k S_k(s)
 *   err_stat (device double, may be NULL): err_stat[col0+s] = sum_k |rounding error of PT[k][col0+s]|   (L1)
 *                                                             sum_k (rounding error of PT[k][col0+s])^2 (L2)
 * center: device, n doubles, or NULL for no centring.
 * Two kernels serve this call.  P64T == NULL (the production case: the float64 profiles are only needed when a pair
 * is flagged for refinement, and are then requested by a second call with PT == NULL): a persistent TMA-fed kernel —
 * it needs X 16-byte aligned with ldx * sizeof(element) a multiple of 16 (pad ldx; fuf_all_pairs_host does).
 * Otherwise a chunk kernel whose P64T reproduces the reference's additions child by child. */
int fuf_push_up(fuf_ctx *ctx, const void *X, int x_dtype, int64_t N, int64_t ldx, int mode, const double *center,
                float *PT, int64_t ldp, double *P64T, int64_t ldp64, double *err_stat, int64_t col0, void *stream);

/* Push-up, validation path: float64 in/out, additions in the reference's order (bit-identical to
 * push_up_L1 / push_up_L2 for finite inputs).  P64T: device double, node-major, ld ldp. */
int fuf_push_up_f64(fuf_ctx *ctx, const double *X, int64_t N, int64_t ldx, int mode, double *P64T, int64_t ldp,
                    int64_t col0, void *stream);

/* Load ALREADY pushed profiles (what pairwise_computation receives, src/algorithms/emd_unifrac.py:42):
 * P_rows (device, p_dtype, sample-major N x n, ld >= n) -> node-major PT of dtype out_dtype, columns [col0, col0+N). */
int fuf_load_pushed(fuf_ctx *ctx, const void *P_rows, int p_dtype, int64_t N, int64_t ld, void *PT, int out_dtype,
                    int64_t ldp, int64_t col0, void *stream);

/* The same centring / rounding for profiles that are ALREADY pushed (P64T from fuf_load_pushed with out_dtype
 * FUF_F64): centre = mean of the first m columns on internal nodes, then PT = (float)(P64T - centre) and the
 * per-sample rounding-error statistic (columns [col0, col0+N) of both buffers). */
int fuf_center_of_pushed(fuf_ctx *ctx, const double *P64T, int64_t m, int64_t ldp64, double *center_out, void *stream);
int fuf_round_pushed(fuf_ctx *ctx, const double *P64T, int64_t N, int64_t ldp64, int mode, const double *center,
                     float *PT, int64_t ldp, double *err_stat, int64_t col0, void *stream);

/* All-pairs distances from pushed profiles.
 *   PT          device float, node-major.  The samples may be split in equal shards of `shard` columns
 *               laid out one after the other: sample s lives at PT[(s / shard) * shard_stride + k * shard + s % shard]
 *               (what an all-gather of per-rank [n][shard] blocks produces).  shard == 0: a single block with
 *               leading dimension ldp (shard_stride ignored).  shard must be a multiple of FUF_TILE.
 *   tile_rows_h host array of the 128-row tile rows to compute, NULL = all of them.
 *   D           device (or pinned, device-accessible host) double.
 *                 tile_rows_h == NULL : full N x N matrix, ld ldd, both triangles and the zero diagonal written.
 *                 tile_rows_h != NULL : compact row blocks; block c (rows c*128 .. c*128+127) holds the rows of
 *                                       tile row tile_rows_h[c]; only columns j >= tile_rows_h[c]*128 are written.
 *   Accumulation: fp32 partial sums over 256 basis entries, combined in double. */
int fuf_pairwise(fuf_ctx *ctx, const float *PT, int64_t N, int64_t ldp, int64_t shard, int64_t shard_stride, int mode,
                 const int32_t *tile_rows_h, int32_t n_tile_rows, double *D, int64_t ldd, int flags, void *stream);

/* --L2 on the tensor cores: D[i,j] = q_i + q_j - 2 (A A^T)[i,j] as an ERROR-FREE integer Gram matrix.
 * Node columns are bucketed by the power of two of their largest entry (bucket scale S); every entry becomes four
 * signed 8-bit digits, a / S = (d0 + d1/254 + d2/254^2 + d3/254^3) / 127; tcgen05.mma kind::i8 accumulates the digit
 * products of order 0..3 exactly in int32 tensor memory (10 MMAs per 32 basis rows); the columns that dominate the
 * scale statistic h2 (at most 256) are computed in float64 instead.  D is then the exact squared distance of the quantised profiles up to the dropped
 * product orders, |error| <= 1.44e-9 (h2_i + h2_j), and exactly 0 for identical samples.
 *   PT       device float, node-major, ld ldp (centred profiles quantise better)
 *   err_in   device, N doubles or NULL: the L2 rounding statistic of PT from fuf_push_up
 *   h2       device, N doubles, out: sum of S^2 over the sample's non-zero Gram entries (statistic of the
 *            linear refinement criterion, lin_kappa = 2.9e-3)
 *   err_out  device, N doubles or NULL, out: rounding statistic incl. the quantisation = (sqrt err_in + sqrt(own))^2
 *   shard, shard_stride, tile_rows_h, n_tile_rows: as in fuf_pairwise (all-gathered shards; compact row blocks)
 *   D        full symmetric N x N with zero diagonal, or compact upper row blocks when tile_rows_h != NULL.
 * Follow with fuf_refine(err_stat = err_out, lin_stat = h2, lin_kappa = 2.9e-3).  mode must be FUF_MODE_L2. */
int fuf_pairwise_gram(fuf_ctx *ctx, const float *PT, int64_t N, int64_t ldp, int64_t shard, int64_t shard_stride, int mode,
                      const double *err_in, double *h2, double *err_out, const int32_t *tile_rows_h, int32_t n_tile_rows,
                      double *D, int64_t ldd, void *stream);

/* Device time of the tcgen05 tile kernel of the last fuf_pairwise_gram call (CUDA events on its stream; waits for
 * it) and the int8 operations it executed (10 products x 2 x 128 x 128 x padded n per tile) — for the tensor-pipe
 * roofline. */
int fuf_gram_last_kernel(fuf_ctx *ctx, float *ms, double *executed_flops);
/* Scale buckets and float64-bypassed node columns of the last fuf_pairwise_gram call. */
int fuf_gram_last_info(const fuf_ctx *ctx, int32_t *n_buckets, int32_t *n_bypassed);

/* Measurement aid (no counterpart in the reference): the tensor-pipe rate of the Gram kernels' own instruction mix —
 * `launches` back-to-back launches in which every SM's MMA warp issues iters x 20 tcgen05.mma kind::i8 (M = 128 per
 * CTA, N = 128, K = 32; cta_group 1 or 2) on operands resident in shared memory: no TMA, no epilogue.  *ms = device
 * time of all launches, *executed_ops = 2 x multiply-adds issued.  bench.py uses it as the roofline denominator of
 * fuf_pairwise_gram (a short call for the burst rate, a seconds-long one for the power-capped sustained rate). */
int fuf_i8_mma_peak(fuf_ctx *ctx, int cta_group, int32_t iters, int32_t launches, float *ms, double *executed_ops,
                    void *stream);

/* Validation mode: everything in double, the reference's operation order per entry (subtract, abs / square, add; no
 * FMA contraction).  Same P64T layouts (single block: ldp; all-gathered shards: shard, shard_stride) and the same
 * D / tile_rows_h conventions as fuf_pairwise: tile_rows_h == NULL -> full symmetric matrix; otherwise compact
 * 128-row blocks, upper part only — what the multi-GPU path uses when too many pairs are flagged for per-pair
 * refinement. */
int fuf_pairwise_f64(fuf_ctx *ctx, const double *P64T, int64_t N, int64_t ldp, int64_t shard, int64_t shard_stride,
                     int mode, const int32_t *tile_rows_h, int32_t n_tile_rows, double *D, int64_t ldd, void *stream);

/* Float64 refinement of what fuf_pairwise wrote into D (same D / tile_rows_h / shard conventions; P64T laid out
 * like PT): every pair i<j with
 *     L1: D[i,j] < kappa (err_stat[i] + err_stat[j])          L2: D[i,j] < kappa (sqrt err_stat[i] + sqrt err_stat[j])^2
 * is recomputed from P64T in double and overwritten (with its mirror when tile_rows_h == NULL).
 * kappa <= 0 selects the default (2e6 / 1.6e13: rounding of the fp32 operands may cost at most 0.5e-6 relative;
 * 0.02 for FUF_CRIT_LINEAR).
 * lin_stat != NULL adds, in the same scan, the linear criterion D[i,j] < lin_kappa (lin_stat[i] + lin_stat[j])
 * (lin_kappa <= 0: 0.02) — what the Gram path needs on top of the rounding criterion.
 * max_pairs > 0: only the first max_pairs flagged pairs are recomputed, the rest are just counted — a per-pair
 * recomputation costs ~1 us, the float64 tile kernel (fuf_pairwise_f64) ~3 ns per pair, so when more than ~0.3 %
 * of the pairs are flagged the caller should recompute everything with fuf_pairwise_f64 (fuf_all_pairs_host does).
 * fuf_refine_count returns the number of FLAGGED pairs.
 * P64T == NULL: count-only mode (D is not modified) — lets several GPUs agree on whether the float64 profiles
 * have to be exchanged at all.
 * fuf_refine_count synchronises the stream and returns how many pairs the last fuf_refine recomputed. */
int fuf_refine(fuf_ctx *ctx, const double *P64T, int64_t N, int64_t ldp64, int64_t shard, int64_t shard_stride, int mode,
               const double *err_stat, double kappa, int criterion, const int32_t *tile_rows_h, int32_t n_tile_rows,
               double *D, int64_t ldd, const double *lin_stat, double lin_kappa, int64_t max_pairs, void *stream);
int fuf_refine_count(fuf_ctx *ctx, int64_t *refined, void *stream);
/* The same count left on the DEVICE (count_dev: device int64), stream-ordered and without a host synchronisation: several
 * GPUs can all-reduce it and go on queueing work while the host has not looked at it yet. */
int fuf_refine_count_device(fuf_ctx *ctx, int64_t *count_dev, void *stream);

/* Join compact row blocks (as written by fuf_pairwise with tile_rows_h, possibly gathered from several
 * devices and concatenated: blocks[c] for c in [0, n_blocks)) into the full symmetric N x N matrix.
 * tile_rows_h[c] == -1 marks a padding block that is skipped. */
int fuf_assemble(fuf_ctx *ctx, const double *blocks, const int32_t *tile_rows_h, int32_t n_blocks, int64_t N,
                 int64_t ldb, double *D, int64_t ldd, void *stream);

/* The same join with half the traffic between devices: fuf_pack_tile_rows copies the part of each row block its tile
 * row owns (rows x columns >= 128 t, row-major with leading dimension N - 128 t) to packed + offsets_h[c] (doubles);
 * the packed buffers of several devices, gathered and concatenated, are joined by fuf_assemble_packed, which reads
 * slot c at packed + offsets_h[c].  tile_rows_h[c] == -1: padding slot (its offset is ignored). */
int fuf_pack_tile_rows(fuf_ctx *ctx, const double *blocks, int64_t ldb, const int32_t *tile_rows_h,
                       const int64_t *offsets_h, int32_t n_blocks, int64_t N, double *packed, void *stream);
int fuf_assemble_packed(fuf_ctx *ctx, const double *packed, const int32_t *tile_rows_h, const int64_t *offsets_h,
                        int32_t n_blocks, int64_t N, double *D, int64_t ldd, void *stream);

/* Same join, but straight into a HOST matrix: each GPU of a node stores the tile rows it owns (and their mirror image)
 * into one N x N host matrix shared by all ranks (shared memory registered with cudaHostRegister in every process), so
 * the device-to-host traffic runs over every GPU's own PCIe link at once instead of through rank 0.
 *   blocks  device double [n_blocks * 128, ldb], as written by fuf_pairwise / fuf_pairwise_gram / fuf_refine with
 *           tile_rows_h; its diagonal tiles are made symmetric in place;
 *   D_h     host double N x N (ldd >= N); slot c writes D_h[128 t : 128 t + 128, 128 t :] and
 *           D_h[128 (t + 1) :, 128 t : 128 t + 128] for t = tile_rows_h[c]  (-1 = padding slot, skipped).
 * Asynchronous on `stream` when D_h is pinned / registered; the union over all tile rows covers D_h exactly once. */
int fuf_store_tile_rows_host(fuf_ctx *ctx, double *blocks, const int32_t *tile_rows_h, int32_t n_blocks, int64_t N,
                             int64_t ldb, double *D_h, int64_t ldd, void *stream);

/* Differential-abundance rows for pairs [pair_begin, pair_end) in itertools.combinations(range(N), 2) order:
 *   out[(p - pair_begin) * n + k] = P_i[k] - P_j[k]          (L1)
 *                                   (P_i[k] - P_j[k])^2      (L2)
 * PT: node-major profiles of dtype p_dtype; out: device or pinned host memory of dtype out_dtype. */
int fuf_diffab_block(fuf_ctx *ctx, const void *PT, int p_dtype, int64_t N, int64_t ldp, int mode, int64_t pair_begin,
                     int64_t pair_end, void *out, int out_dtype, void *stream);

/* Sub-block of the antisymmetric diffab tensor T (N x N x n) that pairwise_computation returns as COO minus its
 * transpose (src/algorithms/emd_unifrac.py:56-72) and that DiffabArrayIndexer.get_diffab / get_diffab_for_node slice
 * (src/utility/differential_abundance.py:18-55), computed from the pushed matrix instead of being stored:
 *   out[(a * nb + b) * nk + c] = T[rows_i[a]][rows_j[b]][nodes[c]]
 *   T[i][j][k] = P_i[k] - P_j[k]  (L1);   sign(j - i) * (P_i[k] - P_j[k])^2  (L2).
 * rows_i (na), rows_j (nb), nodes (nk) are DEVICE int32 arrays of sample / node indices, each in range (not checked on
 * the device); nodes == NULL selects all n nodes in order (nk must be n).  out: device or pinned host, dense. */
int fuf_diffab_gather(fuf_ctx *ctx, const void *PT, int p_dtype, int64_t N, int64_t ldp, int mode, const int32_t *rows_i,
                      int64_t na, const int32_t *rows_j, int64_t nb, const int32_t *nodes, int64_t nk, void *out,
                      int out_dtype, void *stream);

/* Whole path with HOST buffers (the call a reference-side binding makes):
 *   X_h  host, x_dtype, sample-major N x n (ldx >= n), un-pushed normalised profiles;
 *   D_h  host double N x N (ldd >= N).
 * Copies X to the device in row blocks, pushes up (centred, see fuf_push_up_center), computes all pairs,
 * refines (fuf_refine) and returns D in D_h.  Pinned (cudaHostAlloc / cudaHostRegister) buffers are used in
 * place; pageable ones are staged.  flags: FUF_AP_VALIDATE, FUF_AP_NO_REFINE. */
int fuf_all_pairs_host(fuf_ctx *ctx, const void *X_h, int x_dtype, int64_t N, int64_t ldx, int mode, double *D_h,
                       int64_t ldd, int flags);

/* Timing of the phases of the last fuf_all_pairs_host call on this context, milliseconds (CUDA events). */
int fuf_last_timing(const fuf_ctx *ctx, float *h2d_pushup_ms, float *pairwise_ms, float *total_ms);

/* Number of kernel launches issued by this library since the context was created (bench bookkeeping). */
int64_t fuf_launch_count(const fuf_ctx *ctx);

/* Native ingest of sourmash gather CSVs (host only; replaces pd.read_csv + functional_profile_to_vector per file and
 * the ingest pool, src/factory/make_emd_input.py:58-122).
 *   fuf_name_index_create: names[k] = name of basis node k (idx_to_node of tree_to_EMDU_input), n of them.
 *   fuf_ingest_gather_csv: file f -> row f of X (host double, row-major, ld ldx >= n): the un-pushed profile exactly as
 *     functional_profile_to_vector(pd.read_csv(path), input, abundance_key, normalize) returns it (then P[P>0] = 1
 *     when `unweighted`), parsed by n_threads threads (<= 0: all cores).
 *     status[f]: 0 ok, 1 abundance not a number (messages: the offending text; reference: ValueError), 2 empty profile
 *     (reference: Exception "Functional profile is empty!"), 3 unreadable / malformed file (messages: why), 4 name
 *     missing (reference: Exception "Could not parse the name"), 5 column missing (messages: its name; KeyError).
 *     messages: n_files slots of message_stride bytes (may be NULL); warnings: "<file index>\t<id>\n" for every id
 *     that is not in the tree, in file and row order (the reference prints a warning and skips the row), truncated
 *     at warnings_cap.  Returns 0, or 1 if any file failed (see status). */
int fuf_name_index_create(fuf_name_index **out, const char *const *names, int32_t n);
int fuf_name_index_destroy(fuf_name_index *index);
int fuf_ingest_gather_csv(const fuf_name_index *index, const char *const *paths, int32_t n_files, const char *abundance_key,
                          int normalize, int unweighted, int n_threads, double *X, int64_t ldx, int32_t *status,
                          char *messages, int64_t message_stride, char *warnings, int64_t warnings_cap);

/* ---- several GPUs of one node, one process each: tiles with arriving operands, results written in place ---------
 * Replaces, for a sharded job, the same reference loop as fuf_pairwise (src/algorithms/emd_unifrac.py:42-54); the
 * reference has no multi-device form, so there is no reference interface for the exchange itself.
 *
 * fuf_pairwise_tiles: an explicit list of upper-triangle tiles (ti, tj, dependency) on the all-gathered operand layout
 * [G][n][shard].  Every tile is written straight to its place in the FULL matrix D (D[i, j] and D[j, i]; ldd >= N) —
 * D may be this device's memory, a peer GPU's (fuf_peer_open) or mapped page-locked host memory: plain stores.
 * dependency >= 0: the tile reads shard `dependency`, which a copy engine may still be pulling from a peer when the
 * launch starts; the kernel waits (producer warp, bounded) until dep_flags[dependency] has reached `epoch`
 * (fuf_stream_write_u32 behind the copy).  List the tiles in the order their shards arrive.
 * err_stat + flag_list: the refinement criterion of fuf_refine is evaluated in the tile epilogues, where the distances
 * are produced: flag_list[0] counts the flagged pairs i < j, flag_list[1 + 2q], flag_list[2 + 2q] = (i, j) of the first
 * list_cap of them (zero flag_list[0] before the launch).  kappa <= 0: the defaults of fuf_refine.
 * seg_h + seg_done (both or neither): seg_h[c] >= 0 puts tile c into a completion segment; once the tile (and its
 * mirror) is stored and fenced, the device counter seg_done[seg_h[c]] is incremented by 1 (zero the counters before the
 * launch).  A copy stream that waits with fuf_stream_wait_u32 for a segment's tile count can move that part of D
 * (fuf_copy2d_async) while the launch is still computing the rest. */
int fuf_pairwise_tiles(fuf_ctx *ctx, const float *PT, int64_t N, int64_t shard, int64_t shard_stride, int mode,
                       const int32_t *tiles_h, int32_t n_tiles, const uint32_t *dep_flags, uint32_t epoch, double *D,
                       int64_t ldd, const double *err_stat, double kappa, int64_t *flag_list, int64_t list_cap,
                       const int32_t *seg_h, uint32_t *seg_done, void *stream);
/* The listed pairs recomputed in float64 from P64T (layouts as in fuf_refine) into D[i, j] and D[j, i] of the full matrix. */
int fuf_refine_list(fuf_ctx *ctx, const double *P64T, int64_t N, int64_t ldp64, int64_t shard, int64_t shard_stride,
                    int mode, const int64_t *flag_list, int64_t list_cap, double *D, int64_t ldd, void *stream);
/* The --L2 tensor-core path of a sharded job, in phases (the monolithic fuf_pairwise_gram does the same on one GPU):
 * every rank quantises only its own shard and pulls the digit planes of the shards its tiles need from its peers.
 *   fuf_gram_colstats: stats (device, 3 n floats: max |entry| | non-zero count | energy per node) of PT_shard [n][ld], the
 *                      first n_valid columns.  All-reduce across the ranks (max, sum, sum) before planning.
 *   fuf_gram_plan:     from the reduced statistics (device), the plan — scale buckets, float64 bypass columns, permuted
 *                      row map; the same on every rank — kept in the context.  Synchronises the stream.  Returns the
 *                      number of digit-plane rows and of (padded) bypass rows: the buffers below must hold at least that.
 *   fuf_gram_digits:   the own shard (index shard_index, n_valid samples) -> its slot of the four int8 digit planes
 *                      planes[p * plane_stride + shard_index * shard_stride + row * shard + s], of the bypass rows
 *                      SUB[shard_index * sub_shard_stride + row * shard + s] (double), and the per-sample q (norm of the
 *                      digit form), h2 (scale statistic), err_out (rounding statistic incl. err_in of the push-up):
 *                      `shard` doubles each, this shard's slice.
 *   fuf_gram_tiles:    the CTA-pair tile kernel on items (row tile ti, column-tile pair b = tiles 2b, 2b+1, dependency);
 *                      q / h2 / err / SUB / planes hold all shards the items touch.  Tiles are stored in place into
 *                      the full matrix D (and mirrored) as in fuf_pairwise_tiles, dependencies and the flag list work
 *                      the same way (criteria: D < 1.6e13 (sqrt err_i + sqrt err_j)^2 or D < 2.9e-3 (h2_i + h2_j)).
 *                      Completion segments as in fuf_pairwise_tiles, except that a finished item advances its counter
 *                      by FUF_GRAM_SEG_UNITS (one count per epilogue warp of the CTA pair). */
#define FUF_GRAM_SEG_UNITS 16
int fuf_gram_colstats(fuf_ctx *ctx, const float *PT_shard, int64_t n_valid, int64_t ld, float *stats, void *stream);
int fuf_gram_plan(fuf_ctx *ctx, const float *stats, int32_t *n_rows, int32_t *n_direct_pad, void *stream);
int fuf_gram_digits(fuf_ctx *ctx, const float *PT_shard, int64_t n_valid, int64_t shard, int8_t *planes, int64_t plane_stride,
                    int64_t shard_stride, int32_t shard_index, const double *err_in, double *q, double *h2, double *err_out,
                    double *SUB, int64_t sub_shard_stride, void *stream);
int fuf_gram_tiles(fuf_ctx *ctx, const int8_t *planes, int64_t plane_stride, int64_t shard_stride, int64_t N, int64_t shard,
                   const int32_t *tiles_h, int32_t n_tiles, const uint32_t *dep_flags, uint32_t epoch, const double *q,
                   const double *h2, const double *err, const double *SUB, int64_t sub_shard_stride, double *D, int64_t ldd,
                   int64_t *flag_list, int64_t list_cap, const int32_t *seg_h, uint32_t *seg_done, void *stream);

/* Device buffers shared between the processes of one node (CUDA IPC): the owner allocates (zero-filled) and hands the
 * 64-byte handle to its peers (any channel: the Python host uses torch.distributed), which map it.  A mapping is an
 * ordinary device pointer for kernels (loads, stores) and copy engines (fuf_copy_async). */
int fuf_peer_alloc(fuf_ctx *ctx, size_t bytes, void **dptr, unsigned char handle[64]);
int fuf_peer_open(fuf_ctx *ctx, const unsigned char handle[64], void **dptr);
int fuf_peer_close(fuf_ctx *ctx, void *dptr);
int fuf_peer_free(fuf_ctx *ctx, void *dptr);
int fuf_copy_async(fuf_ctx *ctx, void *dst, const void *src, size_t bytes, void *stream);
/* Page-lock host memory owned by another allocator (e.g. a shared-memory mapping every rank of a node writes its tiles
 * into) — in 1 GB pieces when one registration of the whole range is refused.  Keep *n_pieces for fuf_host_unregister. */
int fuf_host_register(void *ptr, size_t bytes, int32_t *n_pieces);
int fuf_host_unregister(void *ptr, size_t bytes, int32_t n_pieces);
/* Device-side address of page-locked host memory (cudaHostAlloc / cudaHostRegister): lets fuf_pairwise_tiles store its
 * tiles straight into a host matrix (e.g. the one the ranks of a node share and np.save writes). */
int fuf_host_device_pointer(fuf_ctx *ctx, void *host_ptr, void **dptr);
/* Stream-ordered 32-bit store that needs no SM (cuStreamWriteValue32): the arrival flag of fuf_pairwise_tiles. */
int fuf_stream_write_u32(fuf_ctx *ctx, uint32_t *dptr, uint32_t value, void *stream);
/* Stream-ordered wait (cuStreamWaitValue32): the work queued on `stream` after this call starts once *dptr >= value —
 * a completion counter of fuf_pairwise_tiles / fuf_gram_tiles that a running kernel advances. */
int fuf_stream_wait_u32(fuf_ctx *ctx, const uint32_t *dptr, uint32_t value, void *stream);
/* Pitched copy by a copy engine, `rows` rows of `width` bytes; either side may be device, peer or page-locked host
 * memory (the finished rectangles of D leaving for the caller's matrix). */
int fuf_copy2d_async(fuf_ctx *ctx, void *dst, size_t dst_pitch, const void *src, size_t src_pitch, size_t width, size_t rows,
                     void *stream);

/* Test hook, host only (needs no device): the plan of the chunked push-up for a postorder parent array.
 * Arrays of length n; span_out gets 5 ints per spanning node (k, chunk_first, chunk_last, slot_start, slot_k). */
int fuf_debug_tree_plan(const int32_t *parent_h, int32_t n, int32_t *parent_local, int32_t *cp_after,
                        uint8_t *is_span, int32_t *span_out, int32_t *n_span, int32_t *n_cp, int32_t *contiguous,
                        int32_t *chunk);

/* Size the persistent grids of this context for `sms` SMs (0 = the context's default, the whole device unless FUF_SM_LIMIT
 * is set; < 0 = that many fewer than the default).  For callers that make a running tile kernel wait
 * for a strided peer copy, which the driver may execute as an SM kernel (it needs an SM no persistent CTA is spinning on). */
int fuf_set_sm_limit(fuf_ctx *ctx, int32_t sms);

/* Test hook, host only (needs no device): the work plan of the all-pairs tile kernel for n_tiles tiles on sm_count SMs —
 * whole tiles round-robin for the full waves, equal spans of the (tile, stage) space for the last partial wave when
 * stream_k != 0.  rows_out (may be NULL to query *n_rows): 5 ints per work item (cta, tile, s_begin, s_end, slot);
 * slot < 0: the item covers the whole tile. */
int fuf_debug_pair_plan(int32_t n_tiles, int32_t sm_count, int32_t stages_total, int32_t flush_stages, int32_t stream_k,
                        int32_t *rows_out, int32_t max_rows, int32_t *n_rows, int32_t *grid);

/* Test hook, host only: the work items (row tile ti, column-tile pair b, completion segment or -1) of a full-matrix launch
 * of the CTA-pair Gram kernel for tiles_per_dim tile rows, in processing order; seg_row_end / seg_expected (may be NULL):
 * the tile-row boundary and the counter target of every segment.  rows_out may be NULL to query the counts. */
int fuf_debug_gram_items(int32_t tiles_per_dim, int32_t n_seg_wanted, int32_t *rows_out, int32_t max_rows, int32_t *n_rows,
                         int32_t *seg_row_end, int32_t *seg_expected, int32_t max_seg, int32_t *n_seg);

/* Test hook, host only: the per-node plan words of the TMA push-up kernel (see PushNode in csrc/common.cuh),
 * ceil(n / 32) * 32 entries; *usable = 0 (meta_out untouched) when the tree cannot use that kernel. */
int fuf_debug_push_meta(const int32_t *parent_h, int32_t n, int32_t *meta_out, int32_t *usable);

#ifdef __cplusplus
}
#endif
#endif /* FUNUNIFRAC_H */
