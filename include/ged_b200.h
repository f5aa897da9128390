/*
 * ged_b200.h -- C-ABI of the B200-native GED lower-level solver (libged_b200.so).
 *
 * The reference (Thinklab-SJTU/PPO-BiHyb) has no FFI for this path: the "operator interface" is a set of plain
 * Python functions in utils/ged_env.py.  Each entry point below names the reference function(s) it replaces; the
 * Python host code in ppo-bihyb_b200/ged_env.py keeps the reference's names and signatures on top of these.
 * INTEGRATION.md shows the ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer inside the descriptors is a DEVICE pointer unless the field says "host";
 *   - all calls are asynchronous on `stream` (a cudaStream_t passed as void*), never synchronise, never allocate;
 *   - return value: 0 on success, a GED_E_* code otherwise; ged_last_error() gives the message (thread local);
 *   - per-solve failures (invalid cost entries, malformed graphs) are reported in `status[b]`, not by the call;
 *   - a "pair" is (graph 1, graph 2); R = n1+1, C = n2+1 (one dummy node each, utils/ged_env.py:188-201);
 *     an R x C matrix of a pair is stored dense row-major with row stride C (its own C, not the batch maximum).
 */
#ifndef GED_B200_H
#define GED_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GED_ABI_VERSION 9

/* call-level errors */
#define GED_OK 0
#define GED_E_BADARG 1       /* null pointer / negative size / inconsistent descriptor            */
#define GED_E_TOO_LARGE 2    /* pair does not fit the shared-memory size classes (see ged_limits) */
#define GED_E_CUDA 3         /* a CUDA runtime call failed (message has the CUDA error string)    */

/* per-solve status bits written to status[b] */
#define GED_ST_OK 0
#define GED_ST_INVALID_COST 1   /* NaN or -inf reached the LAP (scipy raises ValueError, ged_env.py:407)   */
#define GED_ST_INFEASIBLE 2     /* LAP infeasible (all +inf)                                              */
#define GED_ST_BAD_GRAPH 4      /* adjacency not symmetric 0/1 off the diagonal                           */

#define GED_SOLVER_IPFP 0       /* ipfp_ged       utils/ged_env.py:256-289 */
#define GED_SOLVER_HUNGARIAN 1  /* hungarian_ged  utils/ged_env.py:303-305 */

/* A set of base pairs in factor form (what construct_k, utils/ged_env.py:160-228, would expand into K).
 * Graph g's adjacency is n x n uint8, row-major, at adj + adj_off[g]; only "non-zero off the diagonal" matters
 * (the mask of ged_env.py:205-211 removes the diagonal).  Labels select the AIDS node metric of
 * ged_env.py:230-234 (substitution 0/1, insertion/deletion 1, dummy-dummy 0); if node_cost != NULL it is used
 * instead: pair p's R x C fp32 matrix (diag(K) reshaped) at node_cost + node_cost_off[p]. */
typedef struct ged_pairs {
    int32_t num_pairs;
    const int32_t *n1, *n2;                   /* [num_pairs] */
    const uint8_t *adj1; const int64_t *adj1_off;
    const uint8_t *adj2; const int64_t *adj2_off;
    const int32_t *lab1; const int64_t *lab1_off;   /* may be NULL when node_cost is given */
    const int32_t *lab2; const int64_t *lab2_off;
    const float *node_cost; const int64_t *node_cost_off;  /* optional */
    int32_t max_n1, max_n2;                   /* host-known upper bounds over the set (sizes shared memory) */
    /* Row stride (bytes) of every adjacency matrix of graph 1 / graph 2; 0 = compact (stride n1 / n2).  A non-zero stride
     * lets the solver read graphs where another consumer keeps them, e.g. a padded [P, N, N] uint8 batch that the
     * policy network also reads and that edits are applied to in place on the device (adj1_off[p] = p * N * N). */
    int32_t adj1_ld, adj2_ld;
} ged_pairs;

/* One batch of B candidate solves.  Candidate b works on base pair base_idx[b] (identity if NULL, then
 * B == num_pairs) with the undirected edge (edit_u[b], edit_v[b]) of graph 1 toggled first, exactly as
 * GEDenv.step does (utils/ged_env.py:110-121); edit_u == NULL or a negative entry means "no edit".
 * The solve runs on the edited pair; `ged` is scored on the UNEDITED base pair (ged_env.py:122,158). */
typedef struct ged_solve_desc {
    ged_pairs pairs;
    int32_t B;
    const int32_t *base_idx;
    const int32_t *edit_u, *edit_v;
    /* Optional scoring graph: GEDenv.step scores on `ori_k`, the K of the ORIGINAL pair, while graph 1 may already
     * carry earlier edits (utils/ged_env.py:122; ged_ppo_bihyb_train.py:286-298).  If score_adj1 != NULL candidate b is
     * scored with graph 1 := the n1 x n1 uint8 adjacency at score_adj1 + score_adj1_off[score_idx[b]] (same n1,
     * same labels, same graph 2); otherwise with the base pair's own unedited graph 1. */
    const uint8_t *score_adj1; const int64_t *score_adj1_off; const int32_t *score_idx;
    int32_t score_adj1_ld;   /* row stride of the scoring adjacencies, 0 = compact (n1) */
    int32_t max_iter;        /* ipfp_ged's max_iter (100 in the reference) */
    int32_t solver_kind;     /* GED_SOLVER_* */
    /* Optional teacher forcing: start IPFP from x_init instead of the Hungarian initialiser.  Candidate b's R x C
     * matrix is at x_init + b * mat_stride. */
    const float *x_init;
    int64_t mat_stride;      /* elements between consecutive candidates in x_init / x_out / workspace / trace;
                                must be >= (max_n1+1)*(max_n2+1) */
    /* Scratch for the fractional iterate X and X*A2^T of the solves in flight (touched only between LAPs, L2 resident):
     * `workspace_slots` rows of 2 (or workspace_mats) * workspace_stride floats.  A solve claims a free row for its lifetime with an atomic
     * compare-and-swap on workspace_locks[row] (int32, zero-initialised by the caller, zero again when the launch ends),
     * so the footprint follows the solves RESIDENT on the device (ged_solve_max_resident), not the batch size.
     * workspace_slots == 0: legacy layout, row b belongs to candidate b (B rows, no locks).
     * workspace_stride == 0 means mat_stride; it must be >= (max_n1+1)*(max_n2+1). */
    float *workspace;
    int32_t *workspace_locks;
    int32_t workspace_slots;
    int64_t workspace_stride;
    /* Optional sub-launch: solve only candidates order[0 .. order_count) (indices into the B candidates; results are
     * still written at the candidate's own index).  The host layer uses this to launch one size class at a time with
     * that class's max_n1/max_n2, so that shared memory per solve follows the pair's size, not the batch maximum. */
    const int32_t *order;
    int32_t order_count;
    /* Optional work queue: one int32 on the device, ZERO when the launch starts (the caller resets it per launch).  When
     * given, the launch is persistent: as many warps as fit the device, each pulling the next candidate of the
     * (order-)list with an atomic increment until the list is exhausted -- a warp never idles behind a slower CTA-mate
     * and the last wave is as short as the longest single solve.  NULL: one candidate per warp, assigned statically. */
    int32_t *work_counter;
    /* Launch shape hint.  0: packed (CTAs filled with candidates; right when the device is busy with this or other
     * launches).  1: spread -- one candidate per CTA until every CTA slot of the device is taken (candidate = blockIdx +
     * warp * gridDim; the work queue is not used): right for a latency-bound call of a few hundred candidates in total
     * (a rollout or beam-search step), where each solve then has an SM (almost) to itself.  2: team -- latency-bound call
     * whose candidates are each solved by ALL warps of a CTA (one warp per 32 columns of the LAP, rows of K.X dealt out to the
     * warps): the reference's per-step fan-out (<= 27 candidates per pair, ged_ppo_bihyb_eval.py:95-100; batch_size solves
     * per rollout step, ged_ppo_bihyb_train.py:294-298) then finishes in 1/2 to 2/3 of a one-warp solve's time; pairs of up
     * to 32 nodes keep a one-warp LAP with three helper warps for K.X and the update (1.15x; calls of <= 512 candidates).
     * Results are bit-identical under every hint. */
    int32_t launch_hint;
    /* Matrices per workspace row: 0 / 2 = the row holds X and X*A2^T (2 * workspace_stride floats), 3 = it holds a third
     * matrix of workspace_stride floats (ABI 9).  With the third matrix a throughput launch of 75-85-node pairs -- where
     * tensor memory and shared memory together hold the LAP cost matrices of only 12 solves per SM -- runs additional warps
     * whose cost matrix lives there (L2 resident): 16 solves per SM, +10 % at 80 nodes.  Results are bit-identical; a caller
     * that passes 0 / 2 gets the tensor + shared-memory plan. */
    int32_t workspace_mats;
} ged_solve_desc;

typedef struct ged_solve_out {
    float *ged;              /* [B]  x^T K_base x of the returned solution  (GEDenv.comp_ged, ged_env.py:245-253) */
    float *ged_edited;       /* [B]  optional: same on the edited pair's K                                       */
    int32_t *rowmatch;       /* [B * match_stride_1]  column of row i, n2 = deleted                               */
    int32_t *colmatch;       /* [B * match_stride_2]  row of column a, n1 = inserted                              */
    int32_t match_stride_1, match_stride_2;   /* >= max_n1 / max_n2 */
    int32_t *iters;          /* [B]  IPFP iterations executed                                                     */
    int32_t *status;         /* [B]  GED_ST_* bits                                                                */
    float *x_out;            /* optional dense 0/1 solution, R x C at x_out + b * mat_stride                       */
    /* Optional per-iteration trace (tests / diagnostics); any pointer may be NULL.  T = max_iter. */
    float *tr_cost;          /* [B][T][mat_stride]  cost = K v at iteration start                                 */
    float *tr_x_after;       /* [B][T][mat_stride]  v after the line-search update                                */
    double *tr_scalars;      /* [B][T][6]  s_vv, s_vb, ub, alpha, beta, t0                                        */
    int32_t *tr_rowmatch;    /* [B][T][match_stride_1]  LAP projection b of each iteration                        */
    int32_t *tr_colmatch;    /* [B][T][match_stride_2]                                                            */
    float *tr_init_cost;     /* [B][mat_stride]  node_cost_mat of heuristic_prediction_hun (ged_env.py:315)        */
    int32_t *tr_init_rowmatch, *tr_init_colmatch;   /* [B][match_stride_*] Hungarian initialiser                  */
    unsigned long long *tr_lap_steps;               /* [B] LAP scan steps executed (row visits)                    */
} ged_solve_out;

/* Replaces GEDenv.step / solve_feasible_ged / ipfp_ged / hungarian_ged for a whole candidate batch
 * (utils/ged_env.py:110-158, 256-289, 303-326).  Throughput launches: one warp per candidate, the LAP's cost matrix in tensor
 * memory (tcgen05.ld/st) for four warps of every CTA and in shared memory for the others.  Latency-bound launches
 * (launch_hint 2): one CTA per candidate, the pair staged in shared memory. */
int ged_solve_batch(const ged_solve_desc *desc, const ged_solve_out *out, void *stream);

/* Replaces hungarian_lap (utils/ged_env.py:329-351) incl. scipy's linear_sum_assignment with its exact traversal
 * and tie-break order.  cost: B matrices, (n1[b]+1) x (n2[b]+1) fp32 at cost + b * mat_stride.
 * lb[b] = sum(pred_x * cost) (fp64 accumulate, rounded to fp32). */
int ged_lsape_batch(int32_t B, const int32_t *n1, const int32_t *n2, int32_t max_n1, int32_t max_n2,
                    const float *cost, int64_t mat_stride,
                    int32_t *rowmatch, int32_t match_stride_1, int32_t *colmatch, int32_t match_stride_2,
                    float *lb, int32_t *status, void *stream);

/* Replaces torch.mm(k, v) of ipfp_ged (utils/ged_env.py:269) without K: cost_b = K_b vec(X_b) from the factors.
 * x, cost and scratch hold one R x C matrix per pair at b * mat_stride; scratch receives X * A2^T. */
int ged_kv_batch(const ged_pairs *pairs, const float *x, float *cost, float *scratch, int64_t mat_stride, void *stream);

/* Replaces Sinkhorn(max_iter=iters, tau)(s, nrows, ncols) as rrwm_ged / ga_ged call it (utils/ged_env.py:434,468;
 * utils/sinkhorn.py:143-239, log domain, non-batched branch) for B matrices at once: matrix b is rows[b] x cols[b] in the
 * top-left corner of an ld_r x ld_c fp32 block at s + b * ld_r * ld_c; log_s = s / tau[b]; the passes alternate row (even)
 * and column (odd) log-sum-exp normalisation over the whole matrix; out receives exp(log_s) inside the block, 0 outside. */
int ged_sinkhorn_batch(int32_t B, const int32_t *rows, const int32_t *cols, int32_t ld_r, int32_t ld_c, const float *s,
                       const float *tau, int32_t iters, float *out, void *stream);

/* Replaces GEDenv.comp_ged (utils/ged_env.py:245-253) for binary x given as assignments: the coalesced edit-cost
 * reduction.  Candidate b scores (rowmatch_b, colmatch_b) on base pair base_idx[b] (identity if NULL). */
int ged_score_batch(const ged_pairs *pairs, int32_t B, const int32_t *base_idx,
                    const int32_t *rowmatch, int32_t match_stride_1, const int32_t *colmatch, int32_t match_stride_2,
                    float *ged, void *stream);

/* comp_ged against a DENSE K (any x, any K; utils/ged_env.py:245-253): out[b] = x_b^T K_b x_b.  K_b is nk x nk fp32
 * at k + b*k_stride (k_stride == 0 shares one K), x_b has nk entries at x + b*nk.  HBM-bound float4 streaming
 * reduction; rows of K whose x entry is zero are skipped (binary x touches n1+n2 rows only).  partial_ws needs
 * ged_quadform_ws_elems(B, nk) floats. */
int64_t ged_quadform_ws_elems(int32_t B, int32_t nk);
int ged_quadform_dense_batch(int32_t B, int32_t nk, const float *k, int64_t k_stride, const float *x, float *out,
                             float *partial_ws, void *stream);

/* Size limits of this build: largest n1+n2 and largest (n1+1)*(n2+1) a single warp's shared-memory slice holds. */
void ged_limits(int32_t *max_nodes_per_graph, int32_t *max_matrix_elems);

/* Upper bound on the solves of one ged_solve_batch launch that can be resident on the current device at once for the
 * given size bounds (SM count x CTAs per SM x warps per CTA of the launch shape): enough workspace rows for a launch. */
int32_t ged_solve_max_resident(int32_t max_n1, int32_t max_n2);

/* Bytes of dynamic shared memory one solve uses for the given bounds (diagnostics / occupancy planning). */
int64_t ged_solve_smem_bytes(int32_t max_n1, int32_t max_n2);

int ged_abi_version(void);
const char *ged_version(void);
const char *ged_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* GED_B200_H */
