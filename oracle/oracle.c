/*
 * oracle.c — CPU restatement of the reference's all-pairs FunUniFrac arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker or as the timed CPU baseline.  The product path
 * (fununifrac_b200/) never imports, links or executes it.
 *
 * Parity status: PINNED — checked against vectors produced by the unmodified reference
 * (tests/golden/make_golden.py → toy_config1.npz, synth_mid.npz, real_tree_rows.npz) in
 * tests/test_oracle.py: push-up is bit-exact, distances are bit-exact (numpy's pairwise
 * summation is restated below) on those fixtures.
 *
 * Each function cites the reference lines it follows (paths relative to /root/reference).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_EPS 2.220446049250313e-16 /* sys.float_info.epsilon, src/algorithms/emd_unifrac.py:10 */

/* push_up_L1 / push_up_L2 — src/algorithms/emd_unifrac.py:20-28 (L2), :31-39 (L1).
 *   P_pushed = deepcopy(P)
 *   for i in range(n-1):
 *       if lint[i,Tint[i]] == 0: lint[i,Tint[i]] = epsilon
 *       P_pushed[Tint[i]] += P_pushed[i]
 *       P_pushed[i] *= lint[i,Tint[i]]         (L2: np.sqrt(lint[...]))
 * parent[i] = Tint[i]; length[i] = lint[i,Tint[i]]; the root (n-1) is never scaled. */
void oracle_push_up(const int32_t *parent, const double *length, int32_t n, const double *P, double *P_pushed,
                    int is_l2)
{
    memcpy(P_pushed, P, (size_t)n * sizeof(double));
    for (int32_t i = 0; i < n - 1; i++) {
        double l = length[i];
        if (l == 0.0) l = ORACLE_EPS;
        P_pushed[parent[i]] += P_pushed[i];
        P_pushed[i] *= is_l2 ? sqrt(l) : l;
    }
}

void oracle_push_up_batch(const int32_t *parent, const double *length, int32_t n, const double *X, double *P,
                          int64_t N, int is_l2)
{
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < N; s++) oracle_push_up(parent, length, n, X + s * n, P + s * n, is_l2);
}

/* numpy's float64 add-reduce over a contiguous vector (np.sum in get_L1/get_L2,
 * src/objects/profile_vector.py:9-11,17-19): pairwise summation with an 8-way unrolled
 * 128-element base case (numpy/_core/src/umath/loops_utils.h.src, DOUBLE_pairwise_sum). */
static double np_pairwise_sum(const double *a, ptrdiff_t n)
{
    if (n < 8) {
        double res = 0.;
        for (ptrdiff_t i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8], res;
        ptrdiff_t i;
        for (i = 0; i < 8; i++) r[i] = a[i];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int u = 0; u < 8; u++) r[u] += a[i + u];
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        ptrdiff_t n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
    }
}

double oracle_np_sum(const double *a, int64_t n) { return np_pairwise_sum(a, (ptrdiff_t)n); }

/* get_L1: np.sum(np.abs(P-Q)); get_L2: np.sum(np.abs(P-Q)**2) — src/objects/profile_vector.py:9-19.
 * scratch holds the elementwise temporary numpy would materialise. */
static double pair_dist(const double *p, const double *q, int32_t n, int is_l2, double *scratch)
{
    if (is_l2) {
        for (int32_t k = 0; k < n; k++) {
            double d = fabs(p[k] - q[k]);
            scratch[k] = d * d;
        }
    } else {
        for (int32_t k = 0; k < n; k++) scratch[k] = fabs(p[k] - q[k]);
    }
    return np_pairwise_sum(scratch, n);
}

/* pairwise_computation without diffab — src/algorithms/emd_unifrac.py:42-54:
 * dists = zeros((N,N)); for i<j: dists[i,j] = dists[j,i] = get_L1|get_L2(P_i, P_j). */
void oracle_pairwise(const double *P, int64_t N, int32_t n, int is_l2, double *D)
{
    memset(D, 0, (size_t)N * (size_t)N * sizeof(double));
#pragma omp parallel
    {
        double *scratch = (double *)malloc((size_t)n * sizeof(double));
#pragma omp for schedule(dynamic, 1)
        for (int64_t i = 0; i < N; i++)
            for (int64_t j = i + 1; j < N; j++) {
                double z = pair_dist(P + i * n, P + j * n, n, is_l2, scratch);
                D[i * N + j] = z;
                D[j * N + i] = z;
            }
        free(scratch);
    }
}

/* Rectangular block of the same matrix (rows [r0,r1) × all columns) — for sharding tests. */
void oracle_pairwise_rows(const double *P, int64_t N, int32_t n, int is_l2, int64_t r0, int64_t r1, double *D_rows)
{
#pragma omp parallel
    {
        double *scratch = (double *)malloc((size_t)n * sizeof(double));
#pragma omp for schedule(dynamic, 1)
        for (int64_t i = r0; i < r1; i++)
            for (int64_t j = 0; j < N; j++)
                D_rows[(i - r0) * N + j] = (i == j) ? 0.0 : pair_dist(P + i * n, P + j * n, n, is_l2, scratch);
        free(scratch);
    }
}

/* get_L1_diffab: P-Q; get_L2_diffab: (P-Q)**2 — src/objects/profile_vector.py:13-15,21-23;
 * assembled as in src/algorithms/emd_unifrac.py:56-72: entry [i,j,:] for i<j, and
 * [j,i,:] = -[i,j,:] (COO minus its (1,0,2)-transpose, also for L2). Dense N×N×n. */
void oracle_diffab_dense(const double *P, int64_t N, int32_t n, int is_l2, double *out)
{
    memset(out, 0, (size_t)N * (size_t)N * (size_t)n * sizeof(double));
    for (int64_t i = 0; i < N; i++)
        for (int64_t j = i + 1; j < N; j++)
            for (int32_t k = 0; k < n; k++) {
                double d = P[i * n + k] - P[j * n + k];
                if (is_l2) d = d * d;
                out[(i * N + j) * n + k] = d;
                out[(j * N + i) * n + k] = -d;
            }
}

/* solve (weighted) — src/algorithms/emd_unifrac.py:80-114 after the optional 0/1 thresholding:
 *   partial_sums = P - Q
 *   for i in range(n-1): val = partial_sums[i]; partial_sums[Tint[i]] += val;
 *       if val != 0: diffab[(i,Tint[i])] = lint*val;  Z += lint*abs(val)
 * NOTE: no 0→epsilon substitution on this legacy path. diffab_out[i] = lint*val (0 where val == 0). */
double oracle_solve(const int32_t *parent, const double *length, int32_t n, const double *P, const double *Q,
                    double *diffab_out)
{
    double *ps = (double *)malloc((size_t)n * sizeof(double));
    double Z = 0.0;
    for (int32_t k = 0; k < n; k++) ps[k] = P[k] - Q[k];
    for (int32_t i = 0; i < n - 1; i++) {
        double val = ps[i];
        ps[parent[i]] += val;
        diffab_out[i] = (val != 0.0) ? length[i] * val : 0.0;
        Z += length[i] * fabs(val);
    }
    diffab_out[n - 1] = 0.0;
    free(ps);
    return Z;
}

void oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
