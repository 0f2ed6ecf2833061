/*
 * oracle.c -- CPU restatement of GRAFIMO's motif-scanning hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under grafimo_b200/ may import, link or execute this
 * file; it is the checker used by tests/, by __graft_entry__.smoke() and by bench.py's
 * cpu_baseline / --impl reference legs.  The shipped product is the CUDA library.
 *
 * Parity status: PINNED.  Every function here is checked in tests/test_oracle_golden.py
 * against (a) the reference's own golden table tests/test_data/expected_results/
 * scoring_results.tsv (704 rows: score, p-value, q-value) and integer matrices, and
 * (b) vectors produced by running the unmodified reference in the dev container
 * (tests/golden/make_golden.py -> tests/golden/cases/<case>.npz).
 *
 * Each function cites the reference lines (paths relative to /root/reference) it restates.
 * The arithmetic ORDER is part of the contract (fp64, no FMA contraction, no reassociation):
 * build with -O2 -ffp-contract=off and never with -ffast-math (see oracle/Makefile).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef __FAST_MATH__
#error "oracle.c must not be built with -ffast-math: summation order is part of the contract"
#endif

#define ORC_RANGE 1000 /* src/grafimo/utils.py:26 */

/* ---------------------------------------------------------------------------------------
 * Integer PWM score of one k-mer.        src/grafimo/score_sequences.py:373-388
 *   score = sum_i score_matrix[nuc_i, i]; an 'N' anywhere makes the score the matrix
 *   minimum entry (min_val) and stops the scan; letters are case-insensitive.
 *   sm is int64[4][w], rows A,C,G,T.  Any other symbol is undefined in the reference
 *   (stale index); here it sets *bad and is scored like 'N'.
 * ------------------------------------------------------------------------------------- */
int64_t orc_score_kmer(const char *seq, int w, const int64_t *sm, int64_t min_val, int *bad)
{
    int64_t score = 0;
    for (int i = 0; i < w; ++i) {
        int idx;
        switch (seq[i]) {
        case 'N': return min_val;            /* :376-378 (only upper-case N short-circuits) */
        case 'A': case 'a': idx = 0; break;
        case 'C': case 'c': idx = 1; break;
        case 'G': case 'g': idx = 2; break;
        case 'T': case 't': idx = 3; break;
        default:
            if (bad) *bad = 1;
            return min_val;
        }
        score += sm[(int64_t)idx * w + i];
    }
    return score;
}

/* ---------------------------------------------------------------------------------------
 * Sequential ascending fp64 sum == numba's ndarray.sum() (numba/np/arraymath.py:163-170:
 * c = 0; for v in nditer(arr): c += v).  Used twice per row at score_sequences.py:390-391.
 * ------------------------------------------------------------------------------------- */
static double seq_sum(const double *a, int64_t lo, int64_t hi)
{
    double c = 0.0; /* strictly sequential: GCC never reassociates fp adds without -ffast-math */
    for (int64_t k = lo; k < hi; ++k) c = c + a[k];
    return c;
}

/* p-value of an integer score.           src/grafimo/score_sequences.py:390-391
 *   tot = pval_mat.sum(); p = pval_mat[score:].sum() / tot                                  */
double orc_pvalue(const double *pval_mat, int64_t L, int64_t score)
{
    double tot = seq_sum(pval_mat, 0, L);
    if (score < 0) score += L; /* python negative slice start; never reached with valid motifs */
    if (score < 0) score = 0;
    if (score > L) score = L;
    return seq_sum(pval_mat, score, L) / tot;
}

/* log-odds.                              src/grafimo/score_sequences.py:393
 *   logodds = (score / scale) + (width * offset)                                           */
double orc_logodds(int64_t score, int64_t scale, int w, double offset)
{
    return ((double)score / (double)scale) + ((double)w * offset);
}

/* ---------------------------------------------------------------------------------------
 * Staden score-distribution DP.          src/grafimo/motif_processing.pyx:588-602
 *   pv[0, sm[n,0]] += 1*bg[n]                       for n in A,C,G,T
 *   pv[pos, sm[n,pos]+idx] += pv[pos-1,idx]*bg[n]   for n in A,C,G,T, idx ascending over
 *                                                   pv[pos-1] > 0   (product rounded, then add)
 *   returns the last row, length RANGE*w+1.
 * sm: int64[4][w] rows A,C,G,T; bg: A,C,G,T.  Returns 0, or -1 on an out-of-range index.
 * ------------------------------------------------------------------------------------- */
int orc_pval_dp(const int64_t *sm, int w, const double *bg, double *out)
{
    int64_t L = (int64_t)ORC_RANGE * w + 1;
    double *prev = (double *)calloc((size_t)L, sizeof(double));
    double *cur = (double *)calloc((size_t)L, sizeof(double));
    if (!prev || !cur) { free(prev); free(cur); return -2; }
    for (int n = 0; n < 4; ++n) {
        int64_t t = sm[(int64_t)n * w + 0];
        if (t < 0 || t >= L) { free(prev); free(cur); return -1; }
        prev[t] += 1 * bg[n];
    }
    for (int pos = 1; pos < w; ++pos) {
        memset(cur, 0, (size_t)L * sizeof(double));
        for (int n = 0; n < 4; ++n) {
            int64_t s = sm[(int64_t)n * w + pos];
            double b = bg[n];
            for (int64_t idx = 0; idx < L; ++idx) {
                double source = prev[idx];
                if (source > 0) {
                    int64_t t = s + idx;
                    if (t < 0 || t >= L) { free(prev); free(cur); return -1; }
                    double prod = source * b; /* separately rounded product (-ffp-contract=off) */
                    cur[t] = cur[t] + prod;
                }
            }
        }
        double *tmp = prev; prev = cur; cur = tmp;
    }
    memcpy(out, prev, (size_t)L * sizeof(double));
    free(prev); free(cur);
    return 0;
}

/* ---------------------------------------------------------------------------------------
 * Benjamini-Hochberg q-values.           src/grafimo/score_sequences.py:425
 *   statsmodels.stats.multitest.multipletests(p, method="fdr_bh")[1] (third party, not under
 *   /root/reference; published algorithm): sort ascending; raw_k = p_k / (k/float(n));
 *   reverse running minimum; clip at 1; undo the sort.
 * ------------------------------------------------------------------------------------- */
typedef struct { double p; int64_t i; } orc_pi;
static int cmp_pi(const void *a, const void *b)
{
    const orc_pi *x = (const orc_pi *)a, *y = (const orc_pi *)b;
    if (x->p < y->p) return -1;
    if (x->p > y->p) return 1;
    return (x->i > y->i) - (x->i < y->i);
}
int orc_bh(const double *p, int64_t n, double *q)
{
    if (n <= 0) return 0;
    orc_pi *v = (orc_pi *)malloc((size_t)n * sizeof(orc_pi));
    if (!v) return -2;
    for (int64_t i = 0; i < n; ++i) { v[i].p = p[i]; v[i].i = i; }
    qsort(v, (size_t)n, sizeof(orc_pi), cmp_pi);
    double run = INFINITY;
    for (int64_t k = n - 1; k >= 0; --k) {
        double ecdf = (double)(k + 1) / (double)n;
        double raw = v[k].p / ecdf;
        if (raw < run) run = raw;
        q[v[k].i] = run > 1.0 ? 1.0 : run;
    }
    free(v);
    return 0;
}

/* ---------------------------------------------------------------------------------------
 * Row loop of one scoring worker.        src/grafimo/score_sequences.py:273-321 (numeric part)
 *   rows: n k-mers of w ASCII bytes each, row stride `stride` bytes.
 *   Outputs per row: integer score, log-odds score, p-value -- with the reference's per-row cost
 *   (two full sequential sums over pval_mat per row, :390-391).
 *   nthreads > 1 splits the rows into contiguous blocks, like the reference splits its input
 *   files over `--cores` processes (score_sequences.py:123-147).
 * ------------------------------------------------------------------------------------- */
typedef struct {
    const char *rows; int64_t lo, hi, stride; int w;
    const int64_t *sm; const double *pval_mat; int64_t L, min_val, scale; double offset;
    int64_t *iscore; double *logodds; double *pvalue; int bad;
} orc_job;

static void *orc_worker(void *arg)
{
    orc_job *j = (orc_job *)arg;
    for (int64_t r = j->lo; r < j->hi; ++r) {
        int bad = 0;
        int64_t s = orc_score_kmer(j->rows + r * j->stride, j->w, j->sm, j->min_val, &bad);
        if (bad) j->bad = 1;
        if (j->iscore) j->iscore[r] = s;
        if (j->pvalue) j->pvalue[r] = orc_pvalue(j->pval_mat, j->L, s);
        if (j->logodds) j->logodds[r] = orc_logodds(s, j->scale, j->w, j->offset);
    }
    return NULL;
}

int orc_score_rows(const char *rows, int64_t n, int w, int64_t stride,
                   const int64_t *sm, const double *pval_mat, int64_t min_val, int64_t scale,
                   double offset, int64_t *iscore, double *logodds, double *pvalue, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 1024) nthreads = 1024;
    if ((int64_t)nthreads > n) nthreads = n > 0 ? (int)n : 1;
    orc_job *jobs = (orc_job *)calloc((size_t)nthreads, sizeof(orc_job));
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    if (!jobs || !th) { free(jobs); free(th); return -2; }
    int64_t L = (int64_t)ORC_RANGE * w + 1;
    for (int t = 0; t < nthreads; ++t) {
        orc_job *j = &jobs[t];
        j->rows = rows; j->stride = stride; j->w = w; j->sm = sm; j->pval_mat = pval_mat; j->L = L;
        j->min_val = min_val; j->scale = scale; j->offset = offset;
        j->iscore = iscore; j->logodds = logodds; j->pvalue = pvalue;
        j->lo = n * t / nthreads; j->hi = n * (t + 1) / nthreads;
    }
    for (int t = 1; t < nthreads; ++t) pthread_create(&th[t], NULL, orc_worker, &jobs[t]);
    orc_worker(&jobs[0]);
    int bad = jobs[0].bad;
    for (int t = 1; t < nthreads; ++t) { pthread_join(th[t], NULL); bad |= jobs[t].bad; }
    free(jobs); free(th);
    return bad ? 1 : 0;
}

/* ---------------------------------------------------------------------------------------
 * Helpers used by the tests only (derived views of the same arithmetic).
 * ------------------------------------------------------------------------------------- */

/* p-value for every integer score in [0, L): p[s] = seqsum(pv[s:]) / seqsum(pv).  O(L^2/2). */
int orc_pvalue_table(const double *pval_mat, int64_t L, double *p_out)
{
    double tot = seq_sum(pval_mat, 0, L);
    for (int64_t s = 0; s < L; ++s) p_out[s] = seq_sum(pval_mat, s, L) / tot;
    return 0;
}

/* reverse-complement score of a forward k-mer = forward score of its reverse complement
 * (what `vg find -E` hands the reference as the '-' row). */
int64_t orc_score_kmer_rc(const char *seq, int w, const int64_t *sm, int64_t min_val, int *bad)
{
    char buf[256];
    if (w > 255) return min_val;
    for (int i = 0; i < w; ++i) {
        char c = seq[w - 1 - i];
        switch (c) {
        case 'A': c = 'T'; break; case 'a': c = 't'; break;
        case 'C': c = 'G'; break; case 'c': c = 'g'; break;
        case 'G': c = 'C'; break; case 'g': c = 'c'; break;
        case 'T': c = 'A'; break; case 't': c = 'a'; break;
        default: break;
        }
        buf[i] = c;
    }
    return orc_score_kmer(buf, w, sm, min_val, bad);
}
