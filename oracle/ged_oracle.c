/*
 * ged_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU oracle ("O-canon") for the GED lower-level solver path of
 * Thinklab-SJTU/PPO-BiHyb.  Nothing in the product path (ppo-bihyb_b200/) may link or call this file; it is
 * compiled by oracle/Makefile into oracle/_build/libged_oracle.so and used by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg as the checker.
 *
 * It restates, in plain C with a FIXED summation order and no FMA contraction (-ffp-contract=off):
 *   - scipy.optimize.linear_sum_assignment (rectangular_lsap.cpp, shortest augmenting path, Crouse 2016) as
 *     called from utils/ged_env.py:407 -- third-party, not vendored by the reference, no version pinned there;
 *     this container has scipy 1.18.1 and the restatement is checked against it (tests/test_oracle_lsap.py);
 *   - hungarian_lap        utils/ged_env.py:329-351  (LSAPE through the expanded (n1+n2)^2 LAP)
 *   - heuristic_prediction_hun / hungarian_ged  utils/ged_env.py:303-326 (brute force AND the degree closed form)
 *   - construct_k / node_metric  utils/ged_env.py:160-243 (never materialised: factor identity)
 *   - ipfp_ged             utils/ged_env.py:256-289
 *   - GEDenv.comp_ged      utils/ged_env.py:245-253 (for binary x)
 *   - GEDenv.step edge toggle  utils/ged_env.py:110-124
 *
 * Parity pinning: the reference has no tests of its own; the oracle is pinned against
 * outputs of the unmodified reference imported in the build container (oracle/gen_golden.py ->
 * tests/golden/ fixtures) and against scipy itself.
 *
 * Canonical arithmetic (what "bit-exact trajectory" means; mirrored instruction for instruction by the CUDA
 * kernels in ppo-bihyb_b200/csrc/):
 *   X, cost: fp32, dense row-major R x C (R=n1+1, C=n2+1), flat index e = i*C + a.
 *   Neighbour sums run over neighbours in ASCENDING index order starting from 0.0f.
 *   rs[j] = sum_a X[j,a]: 32 partials p[l] over a = l, l+32, ... then xor-butterfly (16,8,4,2,1), fp32;
 *   cs[b] = sum_i X[i,b] (ascending i), fp32.
 *   cost[i,a] = ((t1 + t2) - (T3 + T3)) + Cn[i,a]*X[i,a]       (each op rounded to fp32, no fma)
 *       t1 = a<n2 ? rsP[i] - P[i,a] : rsP[i]      rsP[i] = sum_{j in N1(i)} rs[j]   P[i,a] = sum_{j in N1(i)} X[j,a]
 *       t2 = i<n1 ? csQ[a] - Q[i,a] : csQ[a]      csQ[a] = sum_{b in N2(a)} cs[b]   Q[i,a] = sum_{b in N2(a)} X[i,b]
 *       T3 = sum_{j in N1(i)} ( sum_{b in N2(a)} X[j,b] )
 *   Scalars in fp64: 32 partial sums followed by an xor-butterfly (16,8,4,2,1) (canonical order v2):
 *       s_vv = <cost, X>: partial[a & 31] accumulates cost[i,a]*X[i,a] over rows i ascending, then a ascending;
 *       s_vb = <cost, b>: partial[c & 31] accumulates cost[t, c] for rows t ascending with c = the matched column
 *              (n2 for a deleted row), then partial[a & 31] accumulates cost[n1, a] for inserted columns a ascending;
 *       node-cost sums (ub): partial[t & 31] over t = 0..n1+n2-1 (rows, then inserted columns)
 *   alpha = s_vb - s_vv ; beta = (ub - (s_vb + s_vb)) + s_vv ; t0 = -alpha/beta ; t0f = (float)t0
 *   if (beta <= 0 || t0f >= 1) X = b else X[e] = X[e] + t0f*(b[e] - X[e])   (fp32 sub, mul, add; no fma)
 *   break if |s_vv - s_vb| / s_vv < 1e-3   (fp64; NaN compares false)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define GED_OK 0
#define GED_ERR_INFEASIBLE 1
#define GED_ERR_INVALID 2
#define GED_ERR_ALLOC 3

/* ------------------------------------------------------------------------------------------------------------
 * scipy rectangular_lsap, square / nr<=nc case.  cost is nr x nc doubles, row-major.
 * Traversal order, tie-break and dual-update association follow scipy exactly.
 * col4row[nr] out, row4col[nc] out.  *steps (optional) counts inner scan steps (rows visited).
 * ---------------------------------------------------------------------------------------------------------- */
/* diagnostics for kernel design (how often SciPy's tie-break rule actually decides a step) */
static int64_t g_stat_steps = 0, g_stat_tie_steps = 0, g_stat_tie_unassigned = 0;
void ged_oracle_lsap_stats(int64_t *out, int reset)
{
    out[0] = g_stat_steps; out[1] = g_stat_tie_steps; out[2] = g_stat_tie_unassigned;
    if (reset) g_stat_steps = g_stat_tie_steps = g_stat_tie_unassigned = 0;
}

int ged_oracle_lsap(int nr, int nc, const double *cost, int32_t *col4row, int32_t *row4col, int64_t *steps)
{
    if (nr > nc) return GED_ERR_INVALID;
    for (int64_t k = 0; k < (int64_t)nr * nc; ++k)
        if (cost[k] != cost[k] || cost[k] == -INFINITY) return GED_ERR_INVALID;

    double *u = calloc((size_t)nr, sizeof(double));
    double *v = calloc((size_t)nc, sizeof(double));
    double *spc = malloc((size_t)nc * sizeof(double));
    int32_t *path = malloc((size_t)nc * sizeof(int32_t));
    int32_t *remaining = malloc((size_t)nc * sizeof(int32_t));
    uint8_t *SR = malloc((size_t)nr), *SC = malloc((size_t)nc);
    if (!u || !v || !spc || !path || !remaining || !SR || !SC) return GED_ERR_ALLOC;
    for (int i = 0; i < nr; ++i) col4row[i] = -1;
    for (int j = 0; j < nc; ++j) { row4col[j] = -1; path[j] = -1; }
    int status = GED_OK;
    int64_t nsteps = 0;

    for (int cur = 0; cur < nr && status == GED_OK; ++cur) {
        double minVal = 0.0;
        int i = cur, sink = -1, num_remaining = nc;
        for (int it = 0; it < nc; ++it) remaining[it] = nc - it - 1;
        memset(SR, 0, (size_t)nr);
        memset(SC, 0, (size_t)nc);
        for (int j = 0; j < nc; ++j) spc[j] = INFINITY;

        while (sink == -1) {
            int index = -1;
            double lowest = INFINITY;
            SR[i] = 1;
            ++nsteps;
            for (int it = 0; it < num_remaining; ++it) {
                int j = remaining[it];
                double r = minVal + cost[(int64_t)i * nc + j];
                r = r - u[i];
                r = r - v[j];
                if (r < spc[j]) { path[j] = i; spc[j] = r; }
                if (spc[j] < lowest || (spc[j] == lowest && row4col[j] == -1)) { lowest = spc[j]; index = it; }
            }
            minVal = lowest;
            if (minVal == INFINITY) { status = GED_ERR_INFEASIBLE; break; }
            {
                int ties = 0, un = 0;
                for (int it = 0; it < num_remaining; ++it)
                    if (spc[remaining[it]] == lowest) { ++ties; if (row4col[remaining[it]] == -1) ++un; }
                ++g_stat_steps;
                if (ties > 1) { ++g_stat_tie_steps; if (un > 0) ++g_stat_tie_unassigned; }
            }
            int j = remaining[index];
            if (row4col[j] == -1) sink = j; else i = row4col[j];
            SC[j] = 1;
            remaining[index] = remaining[--num_remaining];
        }
        if (status != GED_OK) break;

        u[cur] += minVal;
        for (int r = 0; r < nr; ++r)
            if (SR[r] && r != cur) u[r] += minVal - spc[col4row[r]];
        for (int j = 0; j < nc; ++j)
            if (SC[j]) v[j] -= minVal - spc[j];

        int j = sink;
        for (;;) {
            int r = path[j];
            row4col[j] = r;
            int t = col4row[r]; col4row[r] = j; j = t;
            if (r == cur) break;
        }
    }
    if (steps) *steps = nsteps;
    free(u); free(v); free(spc); free(path); free(remaining); free(SR); free(SC);
    return status;
}

/* ------------------------------------------------------------------------------------------------------------
 * hungarian_lap (utils/ged_env.py:329-351): LSAPE on an R x C fp32 cost (last row = insertions, last column =
 * deletions) through the expanded N x N LAP, N = n1+n2:
 *      [[ C[:n1,:n2] , diag(C[:n1,n2]) else +inf ],
 *       [ diag(C[n1,:n2]) else +inf , 0          ]]
 * rowmatch[i] in [0..n2] (n2 = row i deleted), colmatch[a] in [0..n1] (n1 = column a inserted).
 * lb (optional) = sum(pred_x * C) accumulated in fp64 over rows then inserted columns.
 * ---------------------------------------------------------------------------------------------------------- */
int ged_oracle_lsape(int n1, int n2, const float *C, int32_t *rowmatch, int32_t *colmatch, double *lb, int64_t *steps)
{
    const int N = n1 + n2, Cc = n2 + 1;
    double *big = malloc((size_t)N * N * sizeof(double));
    int32_t *c4r = malloc((size_t)N * sizeof(int32_t)), *r4c = malloc((size_t)N * sizeof(int32_t));
    if (!big || !c4r || !r4c) return GED_ERR_ALLOC;
    for (int r = 0; r < N; ++r)
        for (int c = 0; c < N; ++c) {
            double val;
            if (r < n1) {
                if (c < n2) val = (double)C[r * Cc + c];
                else val = (c - n2 == r) ? (double)C[r * Cc + n2] : INFINITY;
            } else {
                if (c < n2) val = (r - n1 == c) ? (double)C[n1 * Cc + c] : INFINITY;
                else val = 0.0;
            }
            big[(int64_t)r * N + c] = val;
        }
    int st = ged_oracle_lsap(N, N, big, c4r, r4c, steps);
    if (st == GED_OK) {
        for (int i = 0; i < n1; ++i) rowmatch[i] = (c4r[i] < n2) ? c4r[i] : n2;
        for (int a = 0; a < n2; ++a) colmatch[a] = (r4c[a] < n1) ? r4c[a] : n1;
        if (lb) {
            double s = 0.0;
            for (int i = 0; i < n1; ++i) s += (double)C[i * Cc + rowmatch[i]];
            for (int a = 0; a < n2; ++a) if (colmatch[a] == n1) s += (double)C[n1 * Cc + a];
            *lb = s;
        }
    }
    free(big); free(c4r); free(r4c);
    return st;
}

/* ---- helpers on graphs ------------------------------------------------------------------------------------ */

/* GEDenv.step toggle (utils/ged_env.py:110-121) on a symmetric 0/1 adjacency: present -> both directions
 * removed, absent -> both directions added.  u==v only ever touches the diagonal, which K masks out. */
static void apply_edit(int n, const uint8_t *adj, int eu, int ev, uint8_t *out)
{
    memcpy(out, adj, (size_t)n * n);
    for (int i = 0; i < n; ++i) out[i * n + i] = 0;          /* mask M kills the diagonal (ged_env.py:209) */
    if (eu >= 0 && ev >= 0 && eu != ev && eu < n && ev < n) {
        int present = adj[eu * n + ev] || adj[ev * n + eu];
        out[eu * n + ev] = out[ev * n + eu] = present ? 0 : 1;
    }
}

/* node_metric for the AIDS family (utils/ged_env.py:230-234): substitution 0/1 on the atom label,
 * insertion/deletion 1, dummy-dummy 0. */
void ged_oracle_label_cost(int n1, int n2, const int32_t *lab1, const int32_t *lab2, float *Cn)
{
    const int Cc = n2 + 1;
    for (int i = 0; i <= n1; ++i)
        for (int a = 0; a <= n2; ++a) {
            float c;
            if (i < n1 && a < n2) c = (lab1[i] != lab2[a]) ? 1.0f : 0.0f;
            else if (i == n1 && a == n2) c = 0.0f;
            else c = 1.0f;
            Cn[i * Cc + a] = c;
        }
}

/* heuristic_prediction_hun's node_cost_mat (utils/ged_env.py:310-315) in closed form: the optimal LSAPE value of
 * row (i,a) of K depends on node degrees only. */
void ged_oracle_init_cost_closed_form(int n1, int n2, const uint8_t *adj1, const uint8_t *adj2, float *NC)
{
    const int Cc = n2 + 1;
    int *d1 = calloc((size_t)n1 + 1, sizeof(int)), *d2 = calloc((size_t)n2 + 1, sizeof(int));
    for (int i = 0; i < n1; ++i) for (int j = 0; j < n1; ++j) if (i != j && adj1[i * n1 + j]) d1[i]++;
    for (int a = 0; a < n2; ++a) for (int b = 0; b < n2; ++b) if (a != b && adj2[a * n2 + b]) d2[a]++;
    for (int i = 0; i <= n1; ++i)
        for (int a = 0; a <= n2; ++a) {
            int val;
            if (i < n1 && a < n2) { int d = abs(d1[i] - d2[a]) - 1; val = d > 0 ? d : 0; }
            else if (i < n1) val = d1[i];
            else if (a < n2) val = d2[a];
            else val = 0;
            NC[i * Cc + a] = (float)val;
        }
    free(d1); free(d2);
}

/* The same matrix by brute force, exactly as the reference does it: one LSAPE per row of K
 * (utils/ged_env.py:310-313), K rows rebuilt from the factors (utils/ged_env.py:213-226). */
int ged_oracle_init_cost_bruteforce(int n1, int n2, const uint8_t *adj1, const uint8_t *adj2, const float *Cn, float *NC)
{
    const int R = n1 + 1, Cc = n2 + 1;
    float *row = malloc((size_t)R * Cc * sizeof(float));
    int32_t *rm = malloc((size_t)R * sizeof(int32_t)), *cm = malloc((size_t)Cc * sizeof(int32_t));
    if (!row || !rm || !cm) return GED_ERR_ALLOC;
    int st = GED_OK;
    for (int i = 0; i < R && st == GED_OK; ++i)
        for (int a = 0; a < Cc && st == GED_OK; ++a) {
            for (int j = 0; j < R; ++j)
                for (int b = 0; b < Cc; ++b) {
                    float val;
                    if (j == i && b == a) val = Cn[i * Cc + a];
                    else {
                        int a1 = (i < n1 && j < n1 && i != j) ? (adj1[i * n1 + j] != 0) : 0;
                        int a2 = (a < n2 && b < n2 && a != b) ? (adj2[a * n2 + b] != 0) : 0;
                        int m1 = !(i == j && i < n1), m2 = !(a == b && a < n2);
                        val = (float)(abs(a1 - a2) * m1 * m2);
                    }
                    row[j * Cc + b] = val;
                }
            double lb;
            st = ged_oracle_lsape(n1, n2, row, rm, cm, &lb, NULL);
            NC[i * Cc + a] = (float)lb;
        }
    free(row); free(rm); free(cm);
    return st;
}

/* ---- canonical K*vec(X) ----------------------------------------------------------------------------------- */
typedef struct {
    int n1, n2, R, C;
    int *nb1_off, *nb1;      /* CSR neighbour lists of the (edited, diagonal-free) graph 1, ascending */
    int *nb2_off, *nb2;
    float *rs, *cs, *rsP, *csQ;
} kv_ctx;

static int build_csr(int n, const uint8_t *adj, int **off_out, int **idx_out)
{
    int *off = malloc(((size_t)n + 2) * sizeof(int));
    int nnz = 0;
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) if (i != j && adj[i * n + j]) ++nnz;
    int *idx = malloc(((size_t)nnz + 1) * sizeof(int));
    if (!off || !idx) return GED_ERR_ALLOC;
    int p = 0;
    for (int i = 0; i < n; ++i) {
        off[i] = p;
        for (int j = 0; j < n; ++j) if (i != j && adj[i * n + j]) idx[p++] = j;
    }
    off[n] = p; off[n + 1] = p;      /* dummy node has no neighbours */
    *off_out = off; *idx_out = idx;
    return GED_OK;
}

static void kv_apply(const kv_ctx *k, const float *Cn, const float *X, float *cost)
{
    const int n1 = k->n1, n2 = k->n2, R = k->R, C = k->C;
    for (int j = 0; j < R; ++j) {            /* 32 lane-strided partials, then the xor butterfly (fp32) */
        float p[32], q[32];
        for (int l = 0; l < 32; ++l) p[l] = 0.0f;
        for (int a = 0; a < C; ++a) p[a & 31] = p[a & 31] + X[j * C + a];
        for (int off = 16; off >= 1; off >>= 1) {
            for (int l = 0; l < 32; ++l) q[l] = p[l] + p[l ^ off];
            memcpy(p, q, sizeof(q));
        }
        k->rs[j] = p[0];
    }
    for (int b = 0; b < C; ++b) { float s = 0.0f; for (int i = 0; i < R; ++i) s = s + X[i * C + b]; k->cs[b] = s; }
    for (int i = 0; i < R; ++i) { float s = 0.0f; for (int p = k->nb1_off[i]; p < k->nb1_off[i + 1]; ++p) s = s + k->rs[k->nb1[p]]; k->rsP[i] = s; }
    for (int a = 0; a < C; ++a) { float s = 0.0f; for (int p = k->nb2_off[a]; p < k->nb2_off[a + 1]; ++p) s = s + k->cs[k->nb2[p]]; k->csQ[a] = s; }
    for (int i = 0; i < R; ++i)
        for (int a = 0; a < C; ++a) {
            float P = 0.0f, Q = 0.0f, T3 = 0.0f;
            for (int p = k->nb1_off[i]; p < k->nb1_off[i + 1]; ++p) P = P + X[k->nb1[p] * C + a];
            for (int q = k->nb2_off[a]; q < k->nb2_off[a + 1]; ++q) Q = Q + X[i * C + k->nb2[q]];
            for (int p = k->nb1_off[i]; p < k->nb1_off[i + 1]; ++p) {
                const float *xr = X + k->nb1[p] * C;
                float in = 0.0f;
                for (int q = k->nb2_off[a]; q < k->nb2_off[a + 1]; ++q) in = in + xr[k->nb2[q]];
                T3 = T3 + in;
            }
            float t1 = (a < n2) ? (k->rsP[i] - P) : k->rsP[i];
            float t2 = (i < n1) ? (k->csQ[a] - Q) : k->csQ[a];
            float s = t1 + t2;
            float tt = T3 + T3;
            s = s - tt;
            float d = Cn[i * C + a] * X[i * C + a];
            cost[i * C + a] = s + d;
        }
}

static double lane_butterfly(double *p)
{
    for (int off = 16; off >= 1; off >>= 1) {
        double q[32];
        for (int l = 0; l < 32; ++l) q[l] = p[l] + p[l ^ off];
        memcpy(p, q, sizeof(q));
    }
    return p[0];
}

/* <cost, X> in the canonical order: partial[a & 31], rows ascending then columns ascending */
static double dot_rows(const float *a, const float *b, int R, int C)
{
    double p[32];
    for (int l = 0; l < 32; ++l) p[l] = 0.0;
    for (int i = 0; i < R; ++i)
        for (int c = 0; c < C; ++c) p[c & 31] = p[c & 31] + (double)a[i * C + c] * (double)b[i * C + c];
    return lane_butterfly(p);
}

/* <cost, b> for a binary b in the canonical order: partial index = matched column & 31 */
static double matched_sum_cols(int n1, int n2, const float *M, const int32_t *rowmatch, const int32_t *colmatch)
{
    const int C = n2 + 1;
    double p[32];
    for (int l = 0; l < 32; ++l) p[l] = 0.0;
    for (int t = 0; t < n1; ++t) { const int c = rowmatch[t]; p[c & 31] = p[c & 31] + (double)M[t * C + c]; }
    for (int a = 0; a < n2; ++a) if (colmatch[a] == n1) p[a & 31] = p[a & 31] + (double)M[n1 * C + a];
    return lane_butterfly(p);
}

/* sum over the matched entries of an R x C fp32 matrix: t<n1 -> (t,rowmatch[t]); t>=n1 -> (n1,t-n1) if inserted */
static double matched_sum(int n1, int n2, const float *M, const int32_t *rowmatch, const int32_t *colmatch)
{
    const int C = n2 + 1;
    double p[32];
    for (int l = 0; l < 32; ++l) p[l] = 0.0;
    for (int t = 0; t < n1 + n2; ++t) {
        double term;
        if (t < n1) term = (double)M[t * C + rowmatch[t]];
        else { int a = t - n1; term = (colmatch[a] == n1) ? (double)M[n1 * C + a] : 0.0; }
        p[t & 31] = p[t & 31] + term;
    }
    return lane_butterfly(p);
}

/* comp_ged for a binary x (utils/ged_env.py:245-253) from the factors: node costs of the matched entries plus
 * one unit per ORDERED edge of either graph that is not preserved by the mapping. adj must be diagonal-free. */
static double score_assignment(int n1, int n2, const uint8_t *adj1, const uint8_t *adj2, const float *Cn,
                               const int32_t *rowmatch, const int32_t *colmatch)
{
    int64_t nnz1 = 0, nnz2 = 0, preserved = 0;
    for (int i = 0; i < n1; ++i) for (int j = 0; j < n1; ++j) if (i != j && adj1[i * n1 + j]) {
        ++nnz1;
        int a = rowmatch[i], b = rowmatch[j];
        if (a < n2 && b < n2 && a != b && adj2[a * n2 + b]) ++preserved;
    }
    for (int a = 0; a < n2; ++a) for (int b = 0; b < n2; ++b) if (a != b && adj2[a * n2 + b]) ++nnz2;
    double nodes = matched_sum(n1, n2, Cn, rowmatch, colmatch);
    return nodes + (double)(nnz1 + nnz2 - 2 * preserved);
}

double ged_oracle_score(int n1, int n2, const uint8_t *adj1, const uint8_t *adj2, const float *Cn,
                        const int32_t *rowmatch, const int32_t *colmatch)
{
    uint8_t *a1 = malloc((size_t)n1 * n1 + 1), *a2 = malloc((size_t)n2 * n2 + 1);
    apply_edit(n1, adj1, -1, -1, a1);
    apply_edit(n2, adj2, -1, -1, a2);
    double s = score_assignment(n1, n2, a1, a2, Cn, rowmatch, colmatch);
    free(a1); free(a2);
    return s;
}

static void dense_from_match(int n1, int n2, const int32_t *rowmatch, const int32_t *colmatch, float *X)
{
    const int C = n2 + 1;
    memset(X, 0, (size_t)(n1 + 1) * C * sizeof(float));
    for (int i = 0; i < n1; ++i) X[i * C + rowmatch[i]] = 1.0f;
    for (int a = 0; a < n2; ++a) if (colmatch[a] == n1) X[n1 * C + a] = 1.0f;
}

/* Stand-alone canonical K*vec(X) (cost = K X) for tests. */
int ged_oracle_kv(int n1, int n2, const uint8_t *adj1, const uint8_t *adj2, const float *Cn, const float *X, float *cost)
{
    kv_ctx k = {0};
    k.n1 = n1; k.n2 = n2; k.R = n1 + 1; k.C = n2 + 1;
    uint8_t *a1 = malloc((size_t)n1 * n1 + 1), *a2 = malloc((size_t)n2 * n2 + 1);
    apply_edit(n1, adj1, -1, -1, a1);
    apply_edit(n2, adj2, -1, -1, a2);
    if (build_csr(n1, a1, &k.nb1_off, &k.nb1) || build_csr(n2, a2, &k.nb2_off, &k.nb2)) return GED_ERR_ALLOC;
    k.rs = malloc(sizeof(float) * k.R); k.cs = malloc(sizeof(float) * k.C);
    k.rsP = malloc(sizeof(float) * k.R); k.csQ = malloc(sizeof(float) * k.C);
    kv_apply(&k, Cn, X, cost);
    free(a1); free(a2); free(k.nb1_off); free(k.nb1); free(k.nb2_off); free(k.nb2);
    free(k.rs); free(k.cs); free(k.rsP); free(k.csQ);
    return GED_OK;
}

/* Optional per-iteration trace (all pointers may be NULL). Arrays are sized for max_iter iterations. */
typedef struct {
    double *s_vv, *s_vb, *ub, *alpha, *beta, *t0;     /* [max_iter] */
    int32_t *b_rowmatch;                              /* [max_iter * n1] */
    int32_t *b_colmatch;                              /* [max_iter * n2] */
    float *cost;                                      /* [max_iter * R*C] cost = K v at iteration start */
    float *x_after;                                   /* [max_iter * R*C] v after the line-search update */
    int32_t *init_rowmatch, *init_colmatch;           /* Hungarian initialiser (v0) */
    float *init_cost;                                 /* [R*C] node_cost_mat of the initialiser */
    int64_t *lap_steps;                               /* total LAP scan steps (scalar) */
} ged_trace;

/* ipfp_ged (utils/ged_env.py:256-289) on the pair (adj1 with the optional edit (eu,ev) applied, adj2), scored on
 * the UNEDITED pair as GEDenv.step does (utils/ged_env.py:122,158).  solver: 0 = ipfp, 1 = hungarian only. */
int ged_oracle_solve_from(int n1, int n2, const uint8_t *adj1, const uint8_t *adj2, const float *Cn,
                          int eu, int ev, int max_iter, int solver, const float *x_init,
                          int32_t *out_rowmatch, int32_t *out_colmatch, float *out_ged, float *out_ged_edited,
                          int32_t *out_iters, ged_trace *tr);

int ged_oracle_solve(int n1, int n2, const uint8_t *adj1, const uint8_t *adj2, const float *Cn,
                     int eu, int ev, int max_iter, int solver,
                     int32_t *out_rowmatch, int32_t *out_colmatch, float *out_ged, float *out_ged_edited,
                     int32_t *out_iters, ged_trace *tr)
{
    return ged_oracle_solve_from(n1, n2, adj1, adj2, Cn, eu, ev, max_iter, solver, NULL, out_rowmatch, out_colmatch, out_ged,
                                 out_ged_edited, out_iters, tr);
}

/* The same with an optional start iterate x_init (R x C fp32; NULL = the Hungarian start): teacher forcing of single
 * IPFP iterations from the reference's own iterates (tests/test_oracle_trajectory.py), mirroring ged_solve_desc.x_init of the
 * C-ABI.  With x_init the "best" solution before the first iteration is "everything deleted / inserted". */
int ged_oracle_solve_from(int n1, int n2, const uint8_t *adj1, const uint8_t *adj2, const float *Cn,
                          int eu, int ev, int max_iter, int solver, const float *x_init,
                          int32_t *out_rowmatch, int32_t *out_colmatch, float *out_ged, float *out_ged_edited,
                          int32_t *out_iters, ged_trace *tr)
{
    const int R = n1 + 1, C = n2 + 1, RC = R * C;
    int st = GED_OK;
    uint8_t *a1 = malloc((size_t)n1 * n1 + 1), *a1o = malloc((size_t)n1 * n1 + 1), *a2 = malloc((size_t)n2 * n2 + 1);
    float *X = malloc(sizeof(float) * RC), *cost = malloc(sizeof(float) * RC), *B = malloc(sizeof(float) * RC);
    int32_t *rm = malloc(sizeof(int32_t) * R), *cm = malloc(sizeof(int32_t) * C);
    int32_t *brm = malloc(sizeof(int32_t) * R), *bcm = malloc(sizeof(int32_t) * C);
    kv_ctx k = {0};
    k.n1 = n1; k.n2 = n2; k.R = R; k.C = C;
    k.rs = malloc(sizeof(float) * R); k.cs = malloc(sizeof(float) * C);
    k.rsP = malloc(sizeof(float) * R); k.csQ = malloc(sizeof(float) * C);
    apply_edit(n1, adj1, eu, ev, a1);
    apply_edit(n1, adj1, -1, -1, a1o);
    apply_edit(n2, adj2, -1, -1, a2);
    if (build_csr(n1, a1, &k.nb1_off, &k.nb1) || build_csr(n2, a2, &k.nb2_off, &k.nb2)) return GED_ERR_ALLOC;
    int64_t total_steps = 0, steps = 0;
    int iters = 0;

    if (x_init) {
        for (int i = 0; i < n1; ++i) brm[i] = n2;
        for (int a = 0; a < n2; ++a) bcm[a] = n1;
    } else {
        /* hungarian_ged: v0 = LSAPE(node_cost_mat) */
        ged_oracle_init_cost_closed_form(n1, n2, a1, a2, cost);
        if (tr && tr->init_cost) memcpy(tr->init_cost, cost, sizeof(float) * RC);
        st = ged_oracle_lsape(n1, n2, cost, rm, cm, NULL, &steps);
        total_steps += steps;
        if (st != GED_OK) goto done;
        if (tr && tr->init_rowmatch) memcpy(tr->init_rowmatch, rm, sizeof(int32_t) * n1);
        if (tr && tr->init_colmatch) memcpy(tr->init_colmatch, cm, sizeof(int32_t) * n2);
        memcpy(brm, rm, sizeof(int32_t) * n1);
        memcpy(bcm, cm, sizeof(int32_t) * n2);
    }

    if (solver == 0) {
        if (x_init) memcpy(X, x_init, sizeof(float) * RC);
        else dense_from_match(n1, n2, rm, cm, X);
        double best_ub = INFINITY;
        for (int it = 0; it < max_iter; ++it) {
            iters = it + 1;
            kv_apply(&k, Cn, X, cost);
            double s_vv = dot_rows(cost, X, R, C);
            st = ged_oracle_lsape(n1, n2, cost, rm, cm, NULL, &steps);
            total_steps += steps;
            if (st != GED_OK) goto done;
            double ub = score_assignment(n1, n2, a1, a2, Cn, rm, cm);
            if (ub < best_ub) { best_ub = ub; memcpy(brm, rm, sizeof(int32_t) * n1); memcpy(bcm, cm, sizeof(int32_t) * n2); }
            double s_vb = matched_sum_cols(n1, n2, cost, rm, cm);
            double alpha = s_vb - s_vv;
            double beta = (ub - (s_vb + s_vb)) + s_vv;
            double t0 = -alpha / beta;
            float t0f = (float)t0;
            dense_from_match(n1, n2, rm, cm, B);
            if (beta <= 0.0 || t0f >= 1.0f) memcpy(X, B, sizeof(float) * RC);
            else for (int e = 0; e < RC; ++e) { float d = B[e] - X[e]; float m = t0f * d; X[e] = X[e] + m; }
            if (tr) {
                if (tr->s_vv) tr->s_vv[it] = s_vv;
                if (tr->s_vb) tr->s_vb[it] = s_vb;
                if (tr->ub) tr->ub[it] = ub;
                if (tr->alpha) tr->alpha[it] = alpha;
                if (tr->beta) tr->beta[it] = beta;
                if (tr->t0) tr->t0[it] = t0;
                if (tr->b_rowmatch) memcpy(tr->b_rowmatch + (size_t)it * n1, rm, sizeof(int32_t) * n1);
                if (tr->b_colmatch) memcpy(tr->b_colmatch + (size_t)it * n2, cm, sizeof(int32_t) * n2);
                if (tr->cost) memcpy(tr->cost + (size_t)it * RC, cost, sizeof(float) * RC);
                if (tr->x_after) memcpy(tr->x_after + (size_t)it * RC, X, sizeof(float) * RC);
            }
            if (fabs(s_vv - s_vb) / s_vv < 1e-3) break;
        }
    }
    memcpy(out_rowmatch, brm, sizeof(int32_t) * n1);
    memcpy(out_colmatch, bcm, sizeof(int32_t) * n2);
    if (out_ged) *out_ged = (float)score_assignment(n1, n2, a1o, a2, Cn, brm, bcm);
    if (out_ged_edited) *out_ged_edited = (float)score_assignment(n1, n2, a1, a2, Cn, brm, bcm);
    if (out_iters) *out_iters = iters;
    if (tr && tr->lap_steps) *tr->lap_steps = total_steps;
done:
    free(a1); free(a1o); free(a2); free(X); free(cost); free(B); free(rm); free(cm); free(brm); free(bcm);
    free(k.nb1_off); free(k.nb1); free(k.nb2_off); free(k.nb2); free(k.rs); free(k.cs); free(k.rsP); free(k.csQ);
    return st;
}

const char *ged_oracle_version(void) { return "ged_oracle 3 (canonical order v2; x_init)"; }
