/*
 * wilcoxon_oracle.c -- CPU restatement of illico's asymptotic Wilcoxon rank-sum hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.  The product
 * (illico_b200/) never calls it and fails loudly when its CUDA library is missing.
 *
 * Parity status: PINNED.  This file is checked against golden vectors produced by running the
 * unmodified reference (remydubois/illico v0.2.0, imported from /root/reference in the build
 * container by tests/golden/make_golden.py) and against scipy.stats.mannwhitneyu, the
 * reference's own oracle (reference tests/test_asymptotic_wilcoxon.py:63-108).
 *
 * Every function names the reference file:line whose algorithm it restates.  The reference is
 * Python + numba; this is plain C written from the algorithm description, kept in the same
 * operation order wherever floating point is involved:
 *   - rank sums are exact multiples of 0.5 (< 2^53), so their order of accumulation is free;
 *   - tie sums are accumulated as  f64 += (double)(int64)(t^3 - t)  in ascending value order,
 *     the zero block last for the sparse kernels (order matters once the sum passes 2^53);
 *   - the p-value follows utils/math.py:95-118 operation by operation (compile with
 *     -ffp-contract=off; no fused multiply-add).
 *
 * Build: see oracle/Makefile (gcc -O2 -pthread -ffp-contract=off -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define ORACLE_API __attribute__((visibility("default")))

enum { ALT_TWO_SIDED = 0, ALT_LESS = 1, ALT_GREATER = 2 };
enum { DT_F32 = 0, DT_F64 = 1 };
enum { FMT_DENSE = 0, FMT_CSC = 1, FMT_CSR = 2 };

/* ------------------------------------------------------------------------------------------ */
/* sorting helpers (stand-ins for numba's np.argsort / ndarray.sort; results are tie-invariant) */
/* ------------------------------------------------------------------------------------------ */

static void sort_f64(double* a, int64_t n) {
    /* iterative quicksort, median of three, insertion sort below 16 */
    int64_t stack[128];
    int sp = 0;
    int64_t lo = 0, hi = n - 1;
    for (;;) {
        while (hi - lo > 16) {
            int64_t mid = lo + ((hi - lo) >> 1);
            double t;
            if (a[mid] < a[lo]) { t = a[mid]; a[mid] = a[lo]; a[lo] = t; }
            if (a[hi] < a[lo]) { t = a[hi]; a[hi] = a[lo]; a[lo] = t; }
            if (a[hi] < a[mid]) { t = a[hi]; a[hi] = a[mid]; a[mid] = t; }
            double pivot = a[mid];
            int64_t i = lo, j = hi;
            for (;;) {
                while (a[i] < pivot) i++;
                while (pivot < a[j]) j--;
                if (i >= j) break;
                t = a[i]; a[i] = a[j]; a[j] = t;
                i++; j--;
            }
            /* [lo, j] and [j+1, hi]; push the larger, loop on the smaller */
            if (j - lo < hi - (j + 1)) {
                stack[sp++] = j + 1; stack[sp++] = hi; hi = j;
            } else {
                stack[sp++] = lo; stack[sp++] = j; lo = j + 1;
            }
        }
        for (int64_t i = lo + 1; i <= hi; i++) {
            double v = a[i];
            int64_t j = i - 1;
            while (j >= lo && v < a[j]) { a[j + 1] = a[j]; j--; }
            a[j + 1] = v;
        }
        if (sp == 0) break;
        hi = stack[--sp]; lo = stack[--sp];
    }
}

typedef struct { double v; int64_t i; } kv_t;

static void sort_kv(kv_t* a, int64_t n) {
    int64_t stack[128];
    int sp = 0;
    int64_t lo = 0, hi = n - 1;
    for (;;) {
        while (hi - lo > 16) {
            int64_t mid = lo + ((hi - lo) >> 1);
            kv_t t;
            if (a[mid].v < a[lo].v) { t = a[mid]; a[mid] = a[lo]; a[lo] = t; }
            if (a[hi].v < a[lo].v) { t = a[hi]; a[hi] = a[lo]; a[lo] = t; }
            if (a[hi].v < a[mid].v) { t = a[hi]; a[hi] = a[mid]; a[mid] = t; }
            double pivot = a[mid].v;
            int64_t i = lo, j = hi;
            for (;;) {
                while (a[i].v < pivot) i++;
                while (pivot < a[j].v) j--;
                if (i >= j) break;
                t = a[i]; a[i] = a[j]; a[j] = t;
                i++; j--;
            }
            if (j - lo < hi - (j + 1)) {
                stack[sp++] = j + 1; stack[sp++] = hi; hi = j;
            } else {
                stack[sp++] = lo; stack[sp++] = j; lo = j + 1;
            }
        }
        for (int64_t i = lo + 1; i <= hi; i++) {
            kv_t v = a[i];
            int64_t j = i - 1;
            while (j >= lo && v.v < a[j].v) { a[j + 1] = a[j]; j--; }
            a[j + 1] = v;
        }
        if (sp == 0) break;
        hi = stack[--sp]; lo = stack[--sp];
    }
}

/* np.argsort(arr) -> idx (reference ovr/dense_ovr.py:52, ovr/sparse_ovr.py:66) */
static void argsort_f64(const double* arr, int64_t n, int64_t* idx, kv_t* scratch) {
    for (int64_t i = 0; i < n; i++) { scratch[i].v = arr[i]; scratch[i].i = i; }
    sort_kv(scratch, n);
    for (int64_t i = 0; i < n; i++) idx[i] = scratch[i].i;
}

/* ------------------------------------------------------------------------------------------ */
/* primitives                                                                                 */
/* ------------------------------------------------------------------------------------------ */

/* reference illico/utils/ranking.py:7-49  (_accumulate_group_ranksums_from_argsort)
 * Walks the argsorted column, finds tie runs [i, j), gives each member avg_rank = 0.5*(i+1+j),
 * adds it to its group's rank sum and adds the exact integer t^3 - t to the f64 tie sum. */
ORACLE_API double oracle_accumulate_group_ranksums_from_argsort(const double* arr, const int64_t* idx,
                                                                const int64_t* groups, int64_t n,
                                                                double* ranksums) {
    int64_t i = 0;
    double tie_sum = 0.0;
    while (i < n) {
        int64_t j = i + 1;
        while (j < n && arr[idx[j]] == arr[idx[i]]) j++;
        double avg_rank = 0.5 * (double)(i + 1 + j);
        for (int64_t k = i; k < j; k++) ranksums[groups[idx[k]]] += avg_rank;
        int64_t t = j - i;
        tie_sum += (double)(t * t * t - t);
        i = j;
    }
    return tie_sum;
}

/* reference illico/utils/ranking.py:52-158  (rank_sum_and_ties_from_sorted)
 * Two-pointer sweep over the distinct values of sorted A (controls) and sorted B (perturbed):
 * avg_rank = k + 0.5*(t+1); rank sum of B; tie sum over the combined runs (only t > 1 added). */
ORACLE_API void oracle_rank_sum_and_ties_from_sorted(const double* A, int64_t nA, const double* B, int64_t nB,
                                                     double* out_ranksum, double* out_tiesum) {
    int64_t i = 0, j = 0, k = 0;
    double sum_ranks_B = 0.0, tie_sum = 0.0;
    while (i < nA && j < nB) {
        double v = (A[i] < B[j]) ? A[i] : B[j];
        int64_t tA = 0, ii = i;
        while (ii < nA && A[ii] == v) { tA++; ii++; }
        int64_t tB = 0, jj = j;
        while (jj < nB && B[jj] == v) { tB++; jj++; }
        int64_t t = tA + tB;
        double avg_rank = (double)k + 0.5 * (double)(t + 1);
        if (t > 1) tie_sum += (double)(t * t * t - t);
        sum_ranks_B += (double)tB * avg_rank;
        k += t; i = ii; j = jj;
    }
    while (i < nA) {
        double v = A[i];
        int64_t tA = 0, ii = i;
        while (ii < nA && A[ii] == v) { tA++; ii++; }
        if (tA > 1) tie_sum += (double)(tA * tA * tA - tA);
        k += tA; i = ii;
    }
    while (j < nB) {
        double v = B[j];
        int64_t tB = 0, jj = j;
        while (jj < nB && B[jj] == v) { tB++; jj++; }
        double avg_rank = (double)k + 0.5 * (double)(tB + 1);
        sum_ranks_B += (double)tB * avg_rank;
        if (tB > 1) tie_sum += (double)(tB * tB * tB - tB);
        k += tB; j = jj;
    }
    *out_ranksum = sum_ranks_B;
    *out_tiesum = tie_sum;
}

/* reference illico/utils/math.py:64-118  (compute_pval), strict f64, operation by operation */
ORACLE_API double oracle_compute_pval(int64_t n_ref, int64_t n_tgt, int64_t n, double tie_sum, double U, double mu,
                                      double contin_corr, int alternative) {
    double tie_corr = 1.0 - tie_sum / (double)(n * (n - 1) * (n + 1));
    if (tie_corr > 1.0e-9) {
        double sigma = sqrt((double)(n_ref * n_tgt * (n_ref + n_tgt + 1)) / 12.0 * tie_corr);
        if (alternative == ALT_TWO_SIDED) {
            double other = (double)(n_ref * n_tgt) - U;
            if (other < U) U = other; /* min(U, n_ref*n_tgt - U) */
            double delta = U - mu;
            double sgn = (delta > 0.0) ? 1.0 : ((delta < 0.0) ? -1.0 : 0.0);
            double z = (fabs(delta) + sgn * contin_corr) / sigma;
            return erfc(z / sqrt(2.0));
        } else if (alternative == ALT_GREATER) {
            double delta = U - mu;
            double z = (delta - contin_corr) / sigma;
            return 0.5 * erfc(z / sqrt(2.0));
        } else {
            double delta = U - mu;
            double z = (delta + contin_corr) / sigma;
            return 0.5 * erfc(-z / sqrt(2.0));
        }
    }
    return 1.0;
}

/* reference illico/utils/math.py:168-193  (fold_change_from_summed_expr)
 * agg: [G, b] summed expression, out: [G, b] */
static void fold_change_from_summed_expr(const double* agg, int64_t G, int64_t b, const int64_t* counts,
                                         int64_t ref_group, double* out) {
    int64_t total_count = 0;
    for (int64_t g = 0; g < G; g++) total_count += counts[g];
    double* colsum = (double*)calloc((size_t)b, sizeof(double));
    if (ref_group < 0)
        for (int64_t g = 0; g < G; g++)
            for (int64_t j = 0; j < b; j++) colsum[j] += agg[g * b + j];
    for (int64_t g = 0; g < G; g++) {
        for (int64_t j = 0; j < b; j++) {
            double mu_tgt = agg[g * b + j] / (double)counts[g];
            double mu_ref;
            if (ref_group < 0)
                mu_ref = (colsum[j] - agg[g * b + j]) / (double)(total_count - counts[g]);
            else
                mu_ref = agg[ref_group * b + j] / (double)counts[ref_group];
            out[g * b + j] = (mu_ref == 0.0) ? INFINITY : mu_tgt / mu_ref;
        }
    }
    free(colsum);
}

/* value of element as f64, and the fold-change transform evaluated in the INPUT dtype
 * (np.expm1 on a float32 array stays float32: reference utils/math.py:212, sparse/csc.py:207) */
static inline double load_val(const void* p, int dtype, int64_t i) {
    return dtype == DT_F32 ? (double)((const float*)p)[i] : ((const double*)p)[i];
}
static inline double fc_val(double v, int dtype, int is_log1p) {
    if (!is_log1p) return v;
    return dtype == DT_F32 ? (double)expm1f((float)v) : expm1(v);
}

/* ------------------------------------------------------------------------------------------ */
/* sparse batch containers (stand-ins for the CSCMatrix / CSRMatrix namedtuples)               */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    double* data;      /* values widened to f64 (exact) */
    int32_t* indices;
    int64_t* indptr;
    int64_t n_rows, n_cols;
} sp_t;

static void sp_free(sp_t* m) { free(m->data); free(m->indices); free(m->indptr); }

static int64_t lower_bound_i32(const int32_t* a, int64_t n, int64_t key) {
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
}

/* reference illico/utils/sparse/csc.py:99-136  (csc_get_cols, contiguous range) */
static sp_t csc_get_cols(const void* data, int dtype, const int32_t* indices, const int64_t* indptr,
                         int64_t n_rows, int64_t lb, int64_t ub) {
    sp_t out; out.n_rows = n_rows; out.n_cols = ub - lb;
    int64_t s = indptr[lb], e = indptr[ub], nnz = e - s;
    out.indptr = (int64_t*)malloc(sizeof(int64_t) * (size_t)(ub - lb + 1));
    out.data = (double*)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
    out.indices = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
    for (int64_t j = lb; j <= ub; j++) out.indptr[j - lb] = indptr[j] - s;
    for (int64_t k = 0; k < nnz; k++) { out.data[k] = load_val(data, dtype, s + k); out.indices[k] = indices[s + k]; }
    return out;
}

/* reference illico/utils/sparse/csr.py:199-257  (csr_get_contig_cols_into_csc)
 * binary search of [lb, ub) in every row's sorted indices, count per column, scatter */
static sp_t csr_get_contig_cols_into_csc(const void* data, int dtype, const int32_t* indices, const int64_t* indptr,
                                         int64_t n_rows, int64_t lb, int64_t ub) {
    sp_t out; out.n_rows = n_rows; out.n_cols = ub - lb;
    int64_t b = ub - lb;
    int64_t* cnt = (int64_t*)calloc((size_t)(b + 1), sizeof(int64_t));
    int64_t* bounds = (int64_t*)malloc(sizeof(int64_t) * 2 * (size_t)(n_rows > 0 ? n_rows : 1));
    for (int64_t i = 0; i < n_rows; i++) {
        int64_t s = indptr[i], e = indptr[i + 1];
        int64_t cb = lower_bound_i32(indices + s, e - s, lb), rb = lower_bound_i32(indices + s, e - s, ub);
        bounds[2 * i] = cb; bounds[2 * i + 1] = rb;
        for (int64_t k = s + cb; k < s + rb; k++) cnt[indices[k] - lb + 1]++;
    }
    for (int64_t j = 0; j < b; j++) cnt[j + 1] += cnt[j];
    int64_t nnz = cnt[b];
    out.indptr = (int64_t*)malloc(sizeof(int64_t) * (size_t)(b + 1));
    memcpy(out.indptr, cnt, sizeof(int64_t) * (size_t)(b + 1));
    out.data = (double*)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
    out.indices = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
    for (int64_t i = 0; i < n_rows; i++) {
        int64_t s = indptr[i];
        for (int64_t k = s + bounds[2 * i]; k < s + bounds[2 * i + 1]; k++) {
            int64_t c = indices[k] - lb, dst = cnt[c]++;
            out.data[dst] = load_val(data, dtype, k);
            out.indices[dst] = (int32_t)i;
        }
    }
    free(cnt); free(bounds);
    return out;
}

/* reference illico/utils/sparse/csr.py:144-196  (csr_get_contig_cols_into_csr) */
static sp_t csr_get_contig_cols_into_csr(const void* data, int dtype, const int32_t* indices, const int64_t* indptr,
                                         int64_t n_rows, int64_t lb, int64_t ub) {
    sp_t out; out.n_rows = n_rows; out.n_cols = ub - lb;
    out.indptr = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_rows + 1));
    int64_t* bounds = (int64_t*)malloc(sizeof(int64_t) * 2 * (size_t)(n_rows > 0 ? n_rows : 1));
    out.indptr[0] = 0;
    for (int64_t i = 0; i < n_rows; i++) {
        int64_t s = indptr[i], e = indptr[i + 1];
        int64_t cb = lower_bound_i32(indices + s, e - s, lb), rb = lower_bound_i32(indices + s, e - s, ub);
        bounds[2 * i] = cb; bounds[2 * i + 1] = rb;
        out.indptr[i + 1] = out.indptr[i] + (rb - cb);
    }
    int64_t nnz = out.indptr[n_rows];
    out.data = (double*)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
    out.indices = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
    int64_t c = 0;
    for (int64_t i = 0; i < n_rows; i++) {
        int64_t s = indptr[i];
        for (int64_t k = s + bounds[2 * i]; k < s + bounds[2 * i + 1]; k++) {
            out.data[c] = load_val(data, dtype, k);
            out.indices[c] = (int32_t)(indices[k] - lb);
            c++;
        }
    }
    free(bounds);
    return out;
}

/* reference illico/utils/sparse/csc.py:139-183  (csc_get_contig_cols_into_csr) */
static sp_t csc_get_contig_cols_into_csr(const void* data, int dtype, const int32_t* indices, const int64_t* indptr,
                                         int64_t n_rows, int64_t lb, int64_t ub) {
    sp_t out; out.n_rows = n_rows; out.n_cols = ub - lb;
    int64_t* ptr = (int64_t*)calloc((size_t)(n_rows + 1), sizeof(int64_t));
    for (int64_t k = indptr[lb]; k < indptr[ub]; k++) ptr[indices[k] + 1]++;
    for (int64_t i = 0; i < n_rows; i++) ptr[i + 1] += ptr[i];
    int64_t nnz = ptr[n_rows];
    out.indptr = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_rows + 1));
    memcpy(out.indptr, ptr, sizeof(int64_t) * (size_t)(n_rows + 1));
    out.data = (double*)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
    out.indices = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
    for (int64_t j = lb; j < ub; j++)
        for (int64_t k = indptr[j]; k < indptr[j + 1]; k++) {
            int64_t r = indices[k], dst = ptr[r]++;
            out.data[dst] = load_val(data, dtype, k);
            out.indices[dst] = (int32_t)(j - lb);
        }
    free(ptr);
    return out;
}

/* reference illico/utils/sparse/csr.py:103-141  (csr_get_rows_into_csc) */
static sp_t csr_get_rows_into_csc(const sp_t* X, const int64_t* rows, int64_t n_sel) {
    sp_t out; out.n_rows = n_sel; out.n_cols = X->n_cols;
    int64_t b = X->n_cols;
    int64_t* cnt = (int64_t*)calloc((size_t)(b + 1), sizeof(int64_t));
    for (int64_t r = 0; r < n_sel; r++)
        for (int64_t k = X->indptr[rows[r]]; k < X->indptr[rows[r] + 1]; k++) cnt[X->indices[k] + 1]++;
    for (int64_t j = 0; j < b; j++) cnt[j + 1] += cnt[j];
    int64_t nnz = cnt[b];
    out.indptr = (int64_t*)malloc(sizeof(int64_t) * (size_t)(b + 1));
    memcpy(out.indptr, cnt, sizeof(int64_t) * (size_t)(b + 1));
    out.data = (double*)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
    out.indices = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
    for (int64_t r = 0; r < n_sel; r++)
        for (int64_t k = X->indptr[rows[r]]; k < X->indptr[rows[r] + 1]; k++) {
            int64_t c = X->indices[k], dst = cnt[c]++;
            out.data[dst] = X->data[k];
            out.indices[dst] = (int32_t)r;
        }
    free(cnt);
    return out;
}

/* reference illico/utils/ranking.py:161-172  (_sort_csc_columns_inplace) */
static void sort_csc_columns_inplace(sp_t* m) {
    for (int64_t j = 0; j < m->n_cols; j++) sort_f64(m->data + m->indptr[j], m->indptr[j + 1] - m->indptr[j]);
}

/* ------------------------------------------------------------------------------------------ */
/* the six batch kernels; outputs p, U, fc are [G, b] row-major; tie_dbg optional              */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int64_t n_groups;
    const int64_t* encoded_groups; /* [n] */
    const int64_t* counts;         /* [G] */
    const int64_t* indices;        /* [n] cells sorted by group */
    const int64_t* indptr;         /* [G+1] */
    int64_t ref;                   /* -1 = OVR */
} grpc_t;

/* reference illico/ovr/dense_ovr.py:15-80 */
static void dense_ovr_batch(const void* X, int dtype, int64_t n, int64_t ld, int64_t lb, int64_t ub, const grpc_t* g,
                            int is_log1p, int use_continuity, int tie_correct, int alternative,
                            double* p, double* U, double* fc, double* tie_dbg) {
    int64_t b = ub - lb, G = g->n_groups;
    double* col = (double*)malloc(sizeof(double) * (size_t)n);
    int64_t* idx = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
    kv_t* scratch = (kv_t*)malloc(sizeof(kv_t) * (size_t)n);
    double* ranksums = (double*)malloc(sizeof(double) * (size_t)G);
    double* agg = (double*)calloc((size_t)(G * b), sizeof(double));
    double cc = use_continuity ? 0.5 : 0.0;
    for (int64_t j = 0; j < b; j++) {
        for (int64_t i = 0; i < n; i++) col[i] = load_val(X, dtype, i * ld + lb + j); /* chunk_and_fortranize, math.py:247-278 */
        argsort_f64(col, n, idx, scratch);
        memset(ranksums, 0, sizeof(double) * (size_t)G);
        double tie_sum = oracle_accumulate_group_ranksums_from_argsort(col, idx, g->encoded_groups, n, ranksums);
        if (tie_dbg) tie_dbg[j] = tie_sum;
        for (int64_t k = 0; k < G; k++) {
            int64_t n_tgt = g->counts[k], n_ref = n - n_tgt;
            double stat = ((double)(n_ref * n_tgt) + (double)(n_tgt * (n_tgt + 1)) / 2.0) - ranksums[k];
            double mu = (double)(n_ref * n_tgt) / 2.0;
            U[k * b + j] = stat;
            p[k * b + j] = oracle_compute_pval(n_ref, n_tgt, n, tie_correct ? tie_sum : 0.0, stat, mu, cc, alternative);
        }
        /* dense_fold_change, math.py:196-221: per cell ascending, agg[group] += f(x) */
        for (int64_t i = 0; i < n; i++) agg[g->encoded_groups[i] * b + j] += fc_val(col[i], dtype, is_log1p);
    }
    fold_change_from_summed_expr(agg, G, b, g->counts, -1, fc);
    free(col); free(idx); free(scratch); free(ranksums); free(agg);
}

/* reference illico/ovr/sparse_ovr.py:23-97 (kernel) on a CSC batch, + csc_fold_change csc.py:186-211 */
static void sparse_ovr_batch(const sp_t* X, int dtype, const grpc_t* g, int is_log1p, int use_continuity,
                             int tie_correct, int alternative, double* p, double* U, double* fc, double* tie_dbg) {
    int64_t b = X->n_cols, G = g->n_groups;
    int64_t n = 0;
    for (int64_t k = 0; k < G; k++) n += g->counts[k];
    int64_t maxnnz = 1;
    for (int64_t j = 0; j < b; j++) { int64_t c = X->indptr[j + 1] - X->indptr[j]; if (c > maxnnz) maxnnz = c; }
    int64_t* idx = (int64_t*)malloc(sizeof(int64_t) * (size_t)maxnnz);
    int64_t* grp = (int64_t*)malloc(sizeof(int64_t) * (size_t)maxnnz);
    kv_t* scratch = (kv_t*)malloc(sizeof(kv_t) * (size_t)maxnnz);
    double* R1_nz = (double*)malloc(sizeof(double) * (size_t)G);
    double* nnz_pg = (double*)malloc(sizeof(double) * (size_t)G);
    double* agg = (double*)calloc((size_t)(G * b), sizeof(double));
    double cc = use_continuity ? 0.5 : 0.0;
    for (int64_t j = 0; j < b; j++) {
        int64_t s = X->indptr[j], e = X->indptr[j + 1], m = e - s;
        const double* d = X->data + s;
        for (int64_t k = 0; k < m; k++) grp[k] = g->encoded_groups[X->indices[s + k]];
        argsort_f64(d, m, idx, scratch);
        memset(R1_nz, 0, sizeof(double) * (size_t)G);
        memset(nnz_pg, 0, sizeof(double) * (size_t)G);
        double tie_sum = oracle_accumulate_group_ranksums_from_argsort(d, idx, grp, m, R1_nz);
        double n0 = (double)(X->n_rows - m);
        for (int64_t k = 0; k < m; k++) nnz_pg[grp[k]] += 1.0;
        double n0c = n0 * n0 * n0;
        n0c = n0c - n0;           /* n0**3 - n0 with n0 a float64: sparse_ovr.py:49,83 */
        tie_sum += n0c;           /* zero block added LAST */
        if (tie_dbg) tie_dbg[j] = tie_sum;
        for (int64_t k = 0; k < G; k++) {
            int64_t n_tgt = g->counts[k], n_ref = n - n_tgt;
            double nz_pg = (double)g->counts[k] - nnz_pg[k];
            double r1nz = R1_nz[k] + n0 * nnz_pg[k];
            double R1 = r1nz + nz_pg * (n0 + 1.0) / 2.0;
            double stat = (double)(n_ref * n_tgt) + (double)(n_tgt * (n_tgt + 1)) / 2.0 - R1;
            double mu = (double)(n_ref * n_tgt) / 2.0;
            U[k * b + j] = stat;
            p[k * b + j] = oracle_compute_pval(n_ref, n_tgt, n, tie_correct ? tie_sum : 0.0, stat, mu, cc, alternative);
        }
        for (int64_t k = 0; k < m; k++) agg[grp[k] * b + j] += fc_val(d[k], dtype, is_log1p);
    }
    fold_change_from_summed_expr(agg, G, b, g->counts, -1, fc);
    free(idx); free(grp); free(scratch); free(R1_nz); free(nnz_pg); free(agg);
}

/* reference illico/ovo/dense_ovo.py:15-62 (per-group kernel) and :65-137 (wrapper) */
static void dense_ovo_batch(const void* X, int dtype, int64_t n, int64_t ld, int64_t lb, int64_t ub, const grpc_t* g,
                            int is_log1p, int use_continuity, int tie_correct, int alternative,
                            double* p, double* U, double* fc, double* tie_dbg) {
    int64_t b = ub - lb, G = g->n_groups, r = g->ref;
    int64_t n_ref = g->indptr[r + 1] - g->indptr[r];
    const int64_t* ref_rows = g->indices + g->indptr[r];
    double cc = use_continuity ? 0.5 : 0.0;
    /* ref_chunk: gather + fortranize + sort each column once (dense_ovo.py:111-114) */
    double* refc = (double*)malloc(sizeof(double) * (size_t)(n_ref * b > 0 ? n_ref * b : 1));
    for (int64_t j = 0; j < b; j++) {
        for (int64_t i = 0; i < n_ref; i++) refc[j * n_ref + i] = load_val(X, dtype, ref_rows[i] * ld + lb + j);
        sort_f64(refc + j * n_ref, n_ref);
    }
    int64_t max_g = 1;
    for (int64_t k = 0; k < G; k++) if (g->counts[k] > max_g) max_g = g->counts[k];
    double* tgt = (double*)malloc(sizeof(double) * (size_t)(max_g * b));
    for (int64_t k = 0; k < G; k++) {
        if (k == r) { /* reference leaves this row uninitialised; use the sparse kernels' convention */
            for (int64_t j = 0; j < b; j++) { p[k * b + j] = 1.0; U[k * b + j] = -1.0; if (tie_dbg) tie_dbg[k * b + j] = 0.0; }
            continue;
        }
        int64_t n_tgt = g->indptr[k + 1] - g->indptr[k];
        const int64_t* rows = g->indices + g->indptr[k];
        int64_t nn = n_ref + n_tgt;
        double mu = (double)(n_ref * n_tgt) / 2.0;
        for (int64_t j = 0; j < b; j++) {
            double* t = tgt + j * n_tgt;
            for (int64_t i = 0; i < n_tgt; i++) t[i] = load_val(X, dtype, rows[i] * ld + lb + j);
            sort_f64(t, n_tgt);
            double R1, tie_sum;
            oracle_rank_sum_and_ties_from_sorted(refc + j * n_ref, n_ref, t, n_tgt, &R1, &tie_sum);
            double U1 = (double)(n_ref * n_tgt) + (double)(n_tgt * (n_tgt + 1)) / 2.0 - R1;
            U[k * b + j] = U1;
            p[k * b + j] = oracle_compute_pval(n_ref, n_tgt, nn, tie_correct ? tie_sum : 0.0, U1, mu, cc, alternative);
            if (tie_dbg) tie_dbg[k * b + j] = tie_sum;
        }
    }
    /* dense_fold_change on the whole chunk (dense_ovo.py:135) */
    double* agg = (double*)calloc((size_t)(G * b), sizeof(double));
    for (int64_t i = 0; i < n; i++) {
        int64_t gi = g->encoded_groups[i];
        for (int64_t j = 0; j < b; j++) agg[gi * b + j] += fc_val(load_val(X, dtype, i * ld + lb + j), dtype, is_log1p);
    }
    fold_change_from_summed_expr(agg, G, b, g->counts, r, fc);
    free(agg); free(refc); free(tgt);
}

/* reference illico/ovo/sparse_ovo.py:22-100 (single group) and :103-158 (multi group) on a CSR batch,
 * + csr_fold_change (sparse/csr.py:261-286) */
static void sparse_ovo_batch(const sp_t* X, int dtype, const grpc_t* g, int is_log1p, int use_continuity,
                             int tie_correct, int alternative, double* p, double* U, double* fc, double* tie_dbg) {
    int64_t b = X->n_cols, G = g->n_groups, r = g->ref;
    double cc = use_continuity ? 0.5 : 0.0;
    int64_t n_ref = g->indptr[r + 1] - g->indptr[r];
    sp_t ref = csr_get_rows_into_csc(X, g->indices + g->indptr[r], n_ref);
    sort_csc_columns_inplace(&ref);
    for (int64_t k = 0; k < G; k++) {
        if (k == r) {
            for (int64_t j = 0; j < b; j++) { p[k * b + j] = 1.0; U[k * b + j] = -1.0; if (tie_dbg) tie_dbg[k * b + j] = 0.0; }
            continue;
        }
        int64_t n_tgt = g->indptr[k + 1] - g->indptr[k];
        sp_t tgt = csr_get_rows_into_csc(X, g->indices + g->indptr[k], n_tgt);
        sort_csc_columns_inplace(&tgt);
        int64_t nn = n_ref + n_tgt;
        double mu = (double)(n_ref * n_tgt) / 2.0;
        for (int64_t j = 0; j < b; j++) {
            int64_t lbt = tgt.indptr[j], ubt = tgt.indptr[j + 1], lbr = ref.indptr[j], ubr = ref.indptr[j + 1];
            int64_t z_t = n_tgt - (ubt - lbt), z_r = n_ref - (ubr - lbr), Z = z_r + z_t;
            double ranksum, tie_sum;
            oracle_rank_sum_and_ties_from_sorted(ref.data + lbr, ubr - lbr, tgt.data + lbt, ubt - lbt, &ranksum, &tie_sum);
            ranksum += (double)(Z * (ubt - lbt));
            double R1 = ranksum + (double)(z_t * (z_r + z_t + 1)) / 2.0;
            double U1 = (double)(n_ref * n_tgt) + (double)(n_tgt * (n_tgt + 1)) / 2.0 - R1;
            tie_sum += (double)(Z * Z * Z - Z); /* int64 term, zero block LAST: sparse_ovo.py:85 */
            U[k * b + j] = U1;
            p[k * b + j] = oracle_compute_pval(n_ref, n_tgt, nn, tie_correct ? tie_sum : 0.0, U1, mu, cc, alternative);
            if (tie_dbg) tie_dbg[k * b + j] = tie_sum;
        }
        sp_free(&tgt);
    }
    sp_free(&ref);
    double* agg = (double*)calloc((size_t)(G * b), sizeof(double));
    for (int64_t i = 0; i < X->n_rows; i++) {
        int64_t gi = g->encoded_groups[i];
        for (int64_t k = X->indptr[i]; k < X->indptr[i + 1]; k++)
            agg[gi * b + X->indices[k]] += fc_val(X->data[k], dtype, is_log1p);
    }
    fold_change_from_summed_expr(agg, G, b, g->counts, r, fc);
    free(agg);
}

/* ------------------------------------------------------------------------------------------ */
/* driver: reference illico/asymptotic_wilcoxon.py:212-249 (gene batches on a thread pool,      */
/* each batch scattered into results[G, N, 3]); pthreads stand in for joblib's thread pool     */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int fmt, dtype;
    const void* data; const int32_t* indices; const int64_t* indptr;
    int64_t n_rows, n_cols, ld, gene_lb, gene_ub, batch_size, n_batches;
    grpc_t g;
    int is_log1p, use_continuity, tie_correct, alternative;
    double* results; double* tie_dbg;
    volatile int64_t next_batch;
} job_t;

static void run_batch(job_t* J, int64_t bi) {
    const grpc_t* g = &J->g;
    int64_t G = g->n_groups;
    int ovo = g->ref >= 0;
    int64_t lb = J->gene_lb + bi * J->batch_size, ub = lb + J->batch_size;
    if (ub > J->gene_ub) ub = J->gene_ub;
    int64_t b = ub - lb;
    double* p = (double*)malloc(sizeof(double) * (size_t)(G * b) * 3);
    double* U = p + G * b;
    double* fc = U + G * b;
    double* td = J->tie_dbg ? (double*)malloc(sizeof(double) * (size_t)(ovo ? G * b : b)) : NULL;
    if (J->fmt == FMT_DENSE) {
        if (ovo) dense_ovo_batch(J->data, J->dtype, J->n_rows, J->ld, lb, ub, g, J->is_log1p, J->use_continuity, J->tie_correct, J->alternative, p, U, fc, td);
        else dense_ovr_batch(J->data, J->dtype, J->n_rows, J->ld, lb, ub, g, J->is_log1p, J->use_continuity, J->tie_correct, J->alternative, p, U, fc, td);
    } else if (!ovo) {
        /* ovr/sparse_ovr.py:100-155 (CSC) and :158-208 (CSR) */
        sp_t chunk = (J->fmt == FMT_CSC) ? csc_get_cols(J->data, J->dtype, J->indices, J->indptr, J->n_rows, lb, ub)
                                         : csr_get_contig_cols_into_csc(J->data, J->dtype, J->indices, J->indptr, J->n_rows, lb, ub);
        sparse_ovr_batch(&chunk, J->dtype, g, J->is_log1p, J->use_continuity, J->tie_correct, J->alternative, p, U, fc, td);
        sp_free(&chunk);
    } else {
        /* ovo/sparse_ovo.py:163-210 (CSC) and :214-260 (CSR) */
        sp_t chunk = (J->fmt == FMT_CSC) ? csc_get_contig_cols_into_csr(J->data, J->dtype, J->indices, J->indptr, J->n_rows, lb, ub)
                                         : csr_get_contig_cols_into_csr(J->data, J->dtype, J->indices, J->indptr, J->n_rows, lb, ub);
        sparse_ovo_batch(&chunk, J->dtype, g, J->is_log1p, J->use_continuity, J->tie_correct, J->alternative, p, U, fc, td);
        sp_free(&chunk);
    }
    for (int64_t k = 0; k < G; k++)
        for (int64_t j = 0; j < b; j++) {
            double* o = J->results + (k * J->n_cols + lb + j) * 3;
            o[0] = p[k * b + j]; o[1] = U[k * b + j]; o[2] = fc[k * b + j];
            if (td && ovo) J->tie_dbg[k * J->n_cols + lb + j] = td[k * b + j];
        }
    if (td && !ovo) for (int64_t j = 0; j < b; j++) J->tie_dbg[lb + j] = td[j];
    free(p); free(td);
}

static void* worker(void* arg) {
    job_t* J = (job_t*)arg;
    for (;;) {
        int64_t bi = __atomic_fetch_add(&J->next_batch, 1, __ATOMIC_RELAXED);
        if (bi >= J->n_batches) break;
        run_batch(J, bi);
    }
    return NULL;
}

ORACLE_API int oracle_asymptotic_wilcoxon(
    int fmt, int dtype, const void* data, const int32_t* indices, const int64_t* indptr, /* sparse: all three; dense: data only */
    int64_t n_rows, int64_t n_cols, int64_t ld,                                            /* dense leading dimension (elements) */
    int64_t gene_lb, int64_t gene_ub,                                                      /* genes to compute (sample for timing) */
    int64_t n_groups, const int64_t* encoded_groups, const int64_t* counts, const int64_t* grp_indices,
    const int64_t* grp_indptr, int64_t ref_group,
    int is_log1p, int use_continuity, int tie_correct, int alternative,
    int64_t batch_size, int n_threads,
    double* results,   /* [G, n_cols, 3]; only genes in [gene_lb, gene_ub) are written */
    double* tie_dbg)   /* optional: OVR [n_cols], OVO [G, n_cols] */
{
    job_t J;
    memset(&J, 0, sizeof(J));
    J.fmt = fmt; J.dtype = dtype; J.data = data; J.indices = indices; J.indptr = indptr;
    J.n_rows = n_rows; J.n_cols = n_cols; J.ld = ld; J.gene_lb = gene_lb; J.gene_ub = gene_ub;
    J.g.n_groups = n_groups; J.g.encoded_groups = encoded_groups; J.g.counts = counts;
    J.g.indices = grp_indices; J.g.indptr = grp_indptr; J.g.ref = ref_group;
    J.is_log1p = is_log1p; J.use_continuity = use_continuity; J.tie_correct = tie_correct; J.alternative = alternative;
    J.results = results; J.tie_dbg = tie_dbg;
    if (batch_size <= 0) batch_size = gene_ub - gene_lb;
    if (batch_size <= 0) return 0;
    J.batch_size = batch_size;
    J.n_batches = (gene_ub - gene_lb + batch_size - 1) / batch_size;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > J.n_batches) n_threads = (int)J.n_batches;
    if (n_threads == 1) { worker(&J); return 0; }
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
    int started = 0;
    for (int t = 0; t < n_threads; t++) if (pthread_create(&th[t], NULL, worker, &J) == 0) started++; else break;
    if (started == 0) worker(&J);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    free(th);
    return 0;
}

/* reference illico/utils/ranking.py:223-273  (check_indices_sorted_per_parcel) */
ORACLE_API int oracle_check_indices_sorted_per_parcel(const int32_t* indices, const int64_t* indptr, int64_t n_parcels) {
    for (int64_t k = 0; k < n_parcels; k++)
        for (int64_t i = indptr[k] + 1; i < indptr[k + 1]; i++)
            if (indices[i] < indices[i - 1]) return 0;
    return 1;
}

ORACLE_API int oracle_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
