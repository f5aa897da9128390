// ged_kernels.cu -- hand-written sm_100a kernels + C-ABI for the GED lower-level solver path of PPO-BiHyb.
//
// Path replaced (reference utils/ged_env.py): GEDenv.step :110-124 -> solve_feasible_ged :140-158 ->
// [construct_k :160-228, never materialised] -> ipfp_ged :256-289 -> hungarian_ged :303-326 ->
// hungarian_lap :329-351 -> scipy linear_sum_assignment :407 -> comp_ged :245-253.
//
// Execution model of the throughput launches: ONE WARP PER CANDIDATE SOLVE (latency-bound calls of a few hundred candidates
// take the CTA-per-candidate kernels further down: ged_solve_team_kernel).  The sequential core of the path is the LAP projection (a shortest-
// augmenting-path search whose scan order and tie-breaks must equal SciPy's: ~2000 dependent scan steps per LAP at n=60,
// 101 LAPs per solve), so a candidate is owned by a single warp that keeps the LAP's per-column state (dual v, shortest-
// path cost, position in SciPy's `remaining` vector, predecessor) in REGISTERS, one column per (lane, slot).  Throughput
// is bounded by how many solves are in flight per SM (dependent-latency bound) and, at small sizes, by instruction
// issue; both are set by where the n1 x n2 cost matrix of the running LAP lives:
//   * SmemCost: in the warp's private slice of shared memory (15 KB at n=60);
//   * TmemCost: in TENSOR MEMORY (tcgen05.st / tcgen05.ld, 32x32b shape): TMEM is 128 lanes x 512 columns of 32 bits
//     per SM, each warp of a CTA owns one 32-lane quarter, element (i, a) sits at lane a & 31, column i * S + (a >> 5);
//     a row of the matrix is ONE tcgen05.ld.  It is used purely as a second 256 KB scratchpad (no MMA): a hybrid CTA
//     runs 4 TMEM-backed solves next to k shared-memory-backed ones (one instantiation for both, AnyCost);
//   * GmemCost: in the solve's scratch row in global memory (L2 resident), for the four extra warps of a tensor + global
//     hybrid CTA where tensor and shared memory together hold only 12 matrices per SM (75-85 nodes): 16 solves per SM.
// The fractional IPFP iterate X and Y = X A2^T live in global memory (touched only between LAPs, L2 resident).
// K is never built: K vec(X) is evaluated from the two bit-packed adjacency factors.  Warps never synchronise with each
// other (the only CTA-wide barriers bracket the TMEM allocation).
// Launches are persistent (as many CTAs as the device holds, warps pulling candidates from an atomic work queue) or, for
// latency-bound calls of a few hundred candidates, spread one candidate per CTA or solved by all warps of a CTA
// (ged_solve_desc.launch_hint 1 / 2).
//
// All floating-point arithmetic follows the canonical order documented in oracle/ged_oracle.c (no FMA contraction:
// explicit *_rn intrinsics, and the file is compiled with --fmad=false).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>

#include "../../include/ged_b200.h"

// The file is compiled several times (ppo-bihyb_b200/build.py, in parallel): once per slot count S with -DGED_TU_S=S, each
// of those translation units providing the kernels of that S behind three plain functions (launch_solve_sS, ...), and once
// without the macro for the size-independent kernels, the dispatch and the C-ABI.
#if !defined(GED_TU_S) && !defined(GED_TU_TEAM)
#define GED_TU_MAIN 1                        // the translation unit with the size-independent kernels, the dispatch and the C-ABI
#endif

namespace ged_impl {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kMaxS = 7;                       // slots per column group: graphs up to 32*7 = 224 nodes
constexpr int kMaxSmemBytes = 227 * 1024;
constexpr int kSmemPerSm = 228 * 1024;       // per SM, every resident CTA also reserves 1 KB
constexpr int kMaxDeg = 6;                     // neighbour lists held in registers up to this degree (molecules: <= 5)
constexpr int kTmemWarps = 4;                  // one warp per 32-lane quarter of tensor memory

extern thread_local char g_err[512];     // defined in the main translation unit

static int fail(int code, const char *msg)
{
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}
static int fail_cuda(cudaError_t e, const char *where)
{
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return GED_E_CUDA;
}

// ---------------------------------------------------------------------------------------------------------------
// Shared-memory slice of one warp (host and device agree through this one function).  `bytes` is the slice without
// the cost matrix, `cost_bytes` the n1 x n2 block a shared-memory-backed solve adds.
// ---------------------------------------------------------------------------------------------------------------
struct SliceLayout {
    int off_u, off_cdel, off_rowbuf, off_bits1, off_bits2, off_lab1, off_lab2;
    int off_row4col, off_col4row, off_path, off_rowmatch, off_colmatch, off_bestrow, off_bestcol;
    int bytes, cost_bytes;
};

__host__ __device__ inline int align_up(int x, int a) { return (x + a - 1) / a * a; }

__host__ __device__ inline SliceLayout make_layout(int max_n1, int max_n2)
{
    const int Rm = max_n1 + 1, Cm = max_n2 + 1, Nm = max_n1 + max_n2;
    const int W1 = (max_n1 + 31) / 32 > 0 ? (max_n1 + 31) / 32 : 1;
    const int W2 = (max_n2 + 31) / 32 > 0 ? (max_n2 + 31) / 32 : 1;
    SliceLayout L;
    int o = 0;
    L.off_u = o;        o += 8 * (Nm + 2);     // LAP row duals (fp64); K*X temporaries rs, rsP, cs alias it
    L.off_cdel = o;     o += 4 * (Nm + 2);     // indexed by the row r of the expanded matrix: cost[i][n2] (deletion cost of
                                               // row i) at r = i < n1, cost[n1][a] (insertion cost of column a) at r = n1 + a
    L.off_rowbuf = o;   o += 2 * 4 * align_up(Cm, 4);   // two rows of X (double buffered) for the Y = X A2^T gathers
    L.off_bits1 = o;    o += 4 * Rm * W1;
    L.off_bits2 = o;    o += 4 * Cm * W2;
    L.off_lab1 = o;     o += 4 * Rm;
    L.off_lab2 = o;     o += 4 * Cm;
    L.off_row4col = o;  o += 2 * Nm;
    L.off_col4row = o;  o += 2 * Nm;
    L.off_path = o;     o += 2 * Nm;
    L.off_rowmatch = o; o += 2 * Rm;
    L.off_colmatch = o; o += 2 * Cm;
    L.off_bestrow = o;  o += 2 * Rm;
    L.off_bestcol = o;  o += 2 * Cm;
    L.bytes = align_up(o, 16);
    L.cost_bytes = align_up(4 * (max_n1 * max_n2 + 32 * ((max_n2 + 31) / 32 + 1)), 16);   // + padding: lanes read past n2
    return L;
}

struct Slice {
    double *u;
    float *rs, *rsP, *cs;                    // alias u (live only inside kv_apply / init_cost)
    float *cdel, *rowbuf;                    // cdel[r]: deletion cost of row r < n1, insertion cost of column r - n1 above
    uint32_t *bits1, *bits2;
    int32_t *lab1, *lab2;
    int16_t *row4col, *col4row, *path, *rowmatch, *colmatch, *bestrow, *bestcol;
    float *cost;                             // shared-memory cost block (SmemCost only), nullptr otherwise
};

__device__ inline Slice carve(unsigned char *base, const SliceLayout &L, int max_n1, int max_n2, bool with_cost)
{
    Slice s;
    s.u = reinterpret_cast<double *>(base + L.off_u);
    s.rs = reinterpret_cast<float *>(base + L.off_u);
    s.rsP = s.rs + (max_n1 + 1);
    s.cs = s.rsP + (max_n1 + 1);
    s.cdel = reinterpret_cast<float *>(base + L.off_cdel);
    s.rowbuf = reinterpret_cast<float *>(base + L.off_rowbuf);
    s.bits1 = reinterpret_cast<uint32_t *>(base + L.off_bits1);
    s.bits2 = reinterpret_cast<uint32_t *>(base + L.off_bits2);
    s.lab1 = reinterpret_cast<int32_t *>(base + L.off_lab1);
    s.lab2 = reinterpret_cast<int32_t *>(base + L.off_lab2);
    s.row4col = reinterpret_cast<int16_t *>(base + L.off_row4col);
    s.col4row = reinterpret_cast<int16_t *>(base + L.off_col4row);
    s.path = reinterpret_cast<int16_t *>(base + L.off_path);
    s.rowmatch = reinterpret_cast<int16_t *>(base + L.off_rowmatch);
    s.colmatch = reinterpret_cast<int16_t *>(base + L.off_colmatch);
    s.bestrow = reinterpret_cast<int16_t *>(base + L.off_bestrow);
    s.bestcol = reinterpret_cast<int16_t *>(base + L.off_bestcol);
    s.cost = with_cost ? reinterpret_cast<float *>(base + L.bytes) : nullptr;
    return s;
}

// Per-pair view used by the math routines.
struct Pair {
    int n1, n2, R, C, N, W1, W2;
    int maxdeg;                 // max node degree over both graphs (after the candidate edit)
    const float *node_cost;     // global, R x C, or nullptr -> label metric
};

// ---------------------------------------------------------------------------------------------------------------
// Where the n1 x n2 block of the LAP's cost matrix lives.  Element (i, a) belongs to lane a & 31, slot a >> 5.
//   put(i, s, v)   : every lane stores its element (i, lane + 32 s)   (lanes past n2 store a don't-care value)
//   row(i, cs)     : every lane fetches its S elements of row i
//   flush()        : stores are visible to later row() calls of this warp
// ---------------------------------------------------------------------------------------------------------------
template <int S>
struct SmemCost {
    static constexpr bool kPrefetch = true;  // row() is a plain LDS: issue it a step ahead, the scoreboard does the rest
    float *base;
    int stride;                              // = n2
    __device__ __forceinline__ void put(int i, int s, float v) const
    {
        const int a = (threadIdx.x & 31) + 32 * s;
        if (a < stride) base[i * stride + a] = v;
    }
    __device__ __forceinline__ void row(int i, float (&cs)[S]) const
    {
        const float *p = base + i * stride + (threadIdx.x & 31);
#pragma unroll
        for (int s = 0; s < S; ++s) cs[s] = p[32 * s];
    }
    __device__ __forceinline__ void row_issue(int i, float (&cs)[S]) const { row(i, cs); }
    __device__ __forceinline__ void row_wait(int, float (&)[S]) const {}
    __device__ __forceinline__ void flush() const { __syncwarp(); }
    __device__ __forceinline__ void bind_global(float *) {}
};

// The same layout in GLOBAL memory (the third matrix of the solve's scratch row, L2 / L1 resident): the store of the extra
// warps of a hybrid CTA that neither tensor memory nor shared memory has room for.  A row costs an L2 round trip (issued a
// step ahead, like the shared-memory row), so such a warp runs at roughly 0.6 of a shared-memory warp's speed -- it pays as
// an ADDITIONAL warp on an SM whose issue slots are half idle (profiles/r02o_ab_gmem_experiment.md), never as a replacement.
template <int S>
struct GmemCost {
    static constexpr bool kPrefetch = true;
    float *base;
    int stride;                              // = n2
    __device__ __forceinline__ void put(int i, int s, float v) const
    {
        const int a = (threadIdx.x & 31) + 32 * s;
        if (a < stride) base[i * stride + a] = v;
    }
    __device__ __forceinline__ void row(int i, float (&cs)[S]) const
    {
        const float *p = base + i * stride + (threadIdx.x & 31);
#pragma unroll
        for (int s = 0; s < S; ++s) cs[s] = p[32 * s];       // lanes past n2 read don't-care values inside the scratch row
    }
    __device__ __forceinline__ void row_issue(int i, float (&cs)[S]) const { row(i, cs); }
    __device__ __forceinline__ void row_wait(int, float (&)[S]) const {}
    __device__ __forceinline__ void flush() const { __syncwarp(); }
    __device__ __forceinline__ void bind_global(float *p) { base = p; }
};

template <int S>
struct TmemCost {
    // Fetch a step ahead (row_issue ... row_wait) or where the row is consumed (row)?  Measured on B200: prefetching the
    // tensor-memory row is 4-5 % SLOWER at n = 60 / 80 (tcgen05.ld + wait is ~12 cycles; the split form costs registers
    // and a predicated issue), so the row is fetched in place.  The shared-memory store prefetches (LDS ~30 cycles).
    static constexpr bool kPrefetch = false;
    uint32_t taddr;                          // (first lane of this warp's quarter) << 16 | first column
    __device__ __forceinline__ void put(int i, int s, float v) const
    {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr + (uint32_t)(i * S + s)), "r"(__float_as_uint(v)) : "memory");
    }
    __device__ __forceinline__ void row(int i, float (&cs)[S]) const
    {
        const uint32_t a = taddr + (uint32_t)(i * S);
        uint32_t r[8];
        if constexpr (S == 1)
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n\ttcgen05.wait::ld.sync.aligned;" : "=r"(r[0]) : "r"(a) : "memory");
        else if constexpr (S == 2)
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];\n\ttcgen05.wait::ld.sync.aligned;" : "=r"(r[0]), "=r"(r[1]) : "r"(a) : "memory");
        else if constexpr (S == 3)
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%3];\n\ttcgen05.ld.sync.aligned.32x32b.x1.b32 {%2}, [%4];\n\t"
                         "tcgen05.wait::ld.sync.aligned;" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]) : "r"(a), "r"(a + 2u) : "memory");
        else if constexpr (S == 4)
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];\n\ttcgen05.wait::ld.sync.aligned;"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a) : "memory");
        else {
#pragma unroll
            for (int s = 0; s < S; ++s)
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r[s]) : "r"(a + (uint32_t)s) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
#pragma unroll
        for (int s = 0; s < S; ++s) cs[s] = __uint_as_float(r[s]);
    }
    // Split form for prefetching: the loads are issued here ...
    __device__ __forceinline__ void row_issue(int i, float (&cs)[S]) const
    {
        const uint32_t a = taddr + (uint32_t)(i * S);
        uint32_t r[8];
        if constexpr (S == 1)
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(a) : "memory");
        else if constexpr (S == 2)
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a) : "memory");
        else if constexpr (S == 3)
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%3];\n\ttcgen05.ld.sync.aligned.32x32b.x1.b32 {%2}, [%4];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]) : "r"(a), "r"(a + 2u) : "memory");
        else if constexpr (S == 4)
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a) : "memory");
        else {
#pragma unroll
            for (int s = 0; s < S; ++s)
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r[s]) : "r"(a + (uint32_t)s) : "memory");
        }
#pragma unroll
        for (int s = 0; s < S; ++s) cs[s] = __uint_as_float(r[s]);
    }
    // ... and completed here.  The destination registers are read-write operands of the wait, so that no use of them can
    // be scheduled above it (the data of a tcgen05.ld is only defined after tcgen05.wait::ld).
    __device__ __forceinline__ void row_wait(int, float (&cs)[S]) const
    {
        if constexpr (S == 1) asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(cs[0])::"memory");
        else if constexpr (S == 2) asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(cs[0]), "+f"(cs[1])::"memory");
        else if constexpr (S == 3) asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(cs[0]), "+f"(cs[1]), "+f"(cs[2])::"memory");
        else if constexpr (S == 4)
            asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(cs[0]), "+f"(cs[1]), "+f"(cs[2]), "+f"(cs[3])::"memory");
        else if constexpr (S == 5)
            asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(cs[0]), "+f"(cs[1]), "+f"(cs[2]), "+f"(cs[3]), "+f"(cs[4])::"memory");
        else if constexpr (S == 6)
            asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(cs[0]), "+f"(cs[1]), "+f"(cs[2]), "+f"(cs[3]), "+f"(cs[4]), "+f"(cs[5])::"memory");
        else
            asm volatile("tcgen05.wait::ld.sync.aligned;"
                         : "+f"(cs[0]), "+f"(cs[1]), "+f"(cs[2]), "+f"(cs[3]), "+f"(cs[4]), "+f"(cs[5]), "+f"(cs[6])::"memory");
    }
    __device__ __forceinline__ void flush() const { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
    __device__ __forceinline__ void bind_global(float *) {}
};

// Hybrid CTAs (tensor-memory warps next to warps of a second store M: shared memory, or global memory) run ONE instantiation
// of the solve for both kinds of warp: the store is chosen by a warp-uniform flag at each access.  Two instantiations would
// double the hot code (IPFP iteration body ~59 KB each) and the two kinds of warp then evict each other's loops from the 32 KB
// instruction cache (ncu: stall_no_instruction 0.12 -> 1.04 per issue, no gain from the extra warps); for the same reason a
// CTA mixes tensor memory with ONE other store, not both (a three-way store made every hybrid launch 5-11 % slower,
// profiles/r02p_ab_gmem_threeway.md).  Prefetch policy per kind is kept: the shared- / global-memory row is issued a step
// ahead (row_issue), the tensor-memory row is fetched where it is consumed (row_wait carries the row index for that).
template <int S, class M>
struct AnyCost {
    static constexpr bool kPrefetch = true;
    TmemCost<S> t;
    M m;
    bool tm;                                 // warp-uniform
    __device__ __forceinline__ void put(int i, int s, float v) const
    {
        if (tm) t.put(i, s, v);
        else m.put(i, s, v);
    }
    __device__ __forceinline__ void row(int i, float (&cs)[S]) const
    {
        if (tm) t.row(i, cs);
        else m.row(i, cs);
    }
    __device__ __forceinline__ void row_issue(int i, float (&cs)[S]) const
    {
        if (!tm) m.row(i, cs);
    }
    __device__ __forceinline__ void row_wait(int i, float (&cs)[S]) const
    {
        if (tm) t.row(i, cs);
    }
    __device__ __forceinline__ void flush() const
    {
        if (tm) t.flush();
        else m.flush();
    }
    __device__ __forceinline__ void bind_global(float *p) { m.bind_global(p); }
};

// ---------------------------------------------------------------------------------------------------------------
// Warp-level helpers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_butterfly(double x)
{
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) x = __dadd_rn(x, __shfl_xor_sync(kFull, x, off));
    return x;
}
__device__ __forceinline__ float warp_butterfly(float x)
{
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) x = __fadd_rn(x, __shfl_xor_sync(kFull, x, off));
    return x;
}

// node_metric of the AIDS family (ged_env.py:230-234) or the explicit matrix.
__device__ __forceinline__ float node_cost_at(const Pair &P, const Slice &s, int i, int a)
{
    if (P.node_cost) return P.node_cost[i * P.C + a];
    if (i < P.n1 && a < P.n2) return (s.lab1[i] != s.lab2[a]) ? 1.0f : 0.0f;
    return (i == P.n1 && a == P.n2) ? 0.0f : 1.0f;
}

// Pack one graph's uint8 adjacency into bit rows (diagonal removed, ged_env.py:209).  rows = n + 1 (dummy row zero).
// Returns a warp-uniform flag: true if some off-diagonal entry is > 1 or the matrix is not symmetric.
__device__ inline bool load_bits(const uint8_t *adj, int n, int ld, int W, uint32_t *bits, int lane)
{
    bool bad = false;
    if (ld <= 0) ld = n;                     // compact
    for (int i = 0; i < n; ++i)
        for (int w = 0; w < W; ++w) {
            const int j = w * 32 + lane;
            unsigned val = 0, tval = 0;
            if (j < n && j != i) { val = adj[(int64_t)i * ld + j]; tval = adj[(int64_t)j * ld + i]; }
            bad |= (val > 1u) || ((val != 0) != (tval != 0));
            const unsigned word = __ballot_sync(kFull, val != 0);
            if (lane == 0) bits[i * W + w] = word;
        }
    for (int w = lane; w < W; w += 32) bits[n * W + w] = 0u;
    return __any_sync(kFull, bad);
}

__device__ __forceinline__ bool bit_at(const uint32_t *bits, int W, int i, int j)
{
    return (bits[i * W + (j >> 5)] >> (j & 31)) & 1u;
}

// Order-preserving map double -> (hi, lo) unsigned pair (and back), so that a 64-bit floating-point minimum becomes two
// 32-bit integer warp reductions.  -0.0 cannot occur among the reduced costs (minVal starts at +0.0 and x - x = +0.0 in
// round-to-nearest), so the map agrees with IEEE comparison on every value it sees.
__device__ __forceinline__ void ord_from_double(double x, unsigned &ohi, unsigned &olo)
{
    const unsigned hi = (unsigned)__double2hiint(x), lo = (unsigned)__double2loint(x);
    const unsigned neg = (unsigned)((int)hi >> 31);
    ohi = hi ^ (neg | 0x80000000u);
    olo = lo ^ neg;
}
__device__ __forceinline__ double double_from_ord(unsigned ohi, unsigned olo)
{
    const unsigned pos = (unsigned)((int)ohi >> 31);          // all ones for originally non-negative values
    return __hiloint2double((int)(ohi ^ (~pos | 0x80000000u)), (int)(olo ^ ~pos));
}

// ---- per-slot building blocks of the LAP scan step, written as PTX so that each is exactly the intended handful of
// SASS instructions (LOP3 -> predicate, DSETP with the predicate folded in, selects) without moves or branches -----------

// Scan-step shortcut.  Every shortest-path cost of a column still in `remaining` is >= minVal (minVal was their minimum when
// it was set, and only a relaxation changes one afterwards), so after a row's relaxations the new minimum is minVal itself
// unless (a) a value written by this row is BELOW minVal (possible through rounding only, and at the first row of a search,
// where minVal = 0) or (b) no remaining column attains minVal any more.  minVal stays the same in 89 % of the scan steps of an
// AIDS-50+ solve, so a step first bids for the tie-break key at the OLD minVal -- a lane that saw (a) bids key 0 -- and only
// recomputes the 64-bit minimum (two more warp reductions, the order-preserving split, the per-slot minimum chain) when the
// winning bid is 0 (a) or ~0 (b).  (A/B against the plain form: profiles/r02a_ab_lap_shortcut.md.)

// if (live & MASK) && r < spc:  spc = r, pth = i, and keep = 0 if also r < minVal.  `keep` starts a scan step as ~0 and is the
// start value of the lane's tie-break bid (a minimum of keys), so a lane that wrote a value below minVal bids key 0 for free.
template <unsigned MASK>
__device__ __forceinline__ void lap_relax_below(double &spc, int &pth, unsigned &keep, double r, int i, unsigned live, double minVal)
{
    asm("{\n\t.reg .pred q, p, p2;\n\t.reg .b32 t;\n\t"
        "and.b32 t, %5, %6;\n\t"
        "setp.ne.u32 q, t, 0;\n\t"
        "setp.lt.and.f64 p, %3, %0, q;\n\t"
        "selp.f64 %0, %3, %0, p;\n\t"
        "selp.b32 %1, %4, %1, p;\n\t"
        "setp.lt.and.f64 p2, %3, %7, p;\n\t"
        "@p2 mov.b32 %2, 0;\n\t}"
        : "+d"(spc), "+r"(pth), "+r"(keep) : "d"(r), "r"(i), "r"(live), "n"(MASK), "d"(minVal));
}

// the same for the one entry of a row that is not part of a block (a real row's deletion entry, an insertion row's own
// column): the slot is relaxed only in the lane and slot that own column k of the group, id = the column of the group this
// lane's slot holds -- one integer compare chained into the predicate instead of a lane / slot mask built per step
template <unsigned MASK>
__device__ __forceinline__ void lap_relax_own_below(double &spc, int &pth, unsigned &keep, double r, int i, unsigned live, double minVal,
                                                    int k, int id)
{
    asm("{\n\t.reg .pred q, p, p2;\n\t.reg .b32 t;\n\t"
        "and.b32 t, %5, %6;\n\t"
        "setp.ne.u32 q, t, 0;\n\t"
        "setp.eq.and.s32 q, %8, %9, q;\n\t"
        "setp.lt.and.f64 p, %3, %0, q;\n\t"
        "selp.f64 %0, %3, %0, p;\n\t"
        "selp.b32 %1, %4, %1, p;\n\t"
        "setp.lt.and.f64 p2, %3, %7, p;\n\t"
        "@p2 mov.b32 %2, 0;\n\t}"
        : "+d"(spc), "+r"(pth), "+r"(keep) : "d"(r), "r"(i), "r"(live), "n"(MASK), "d"(minVal), "r"(k), "r"(id));
}

// if (live & MASK) && spc < lm:  lm = spc
template <unsigned MASK>
__device__ __forceinline__ void lap_min_value(double &lm, double spc, unsigned live)
{
    asm("{\n\t.reg .pred q, p;\n\t.reg .b32 t;\n\t"
        "and.b32 t, %2, %3;\n\t"
        "setp.ne.u32 q, t, 0;\n\t"
        "setp.lt.and.f64 p, %1, %0, q;\n\t"
        "selp.f64 %0, %1, %0, p;\n\t}"
        : "+d"(lm) : "d"(spc), "r"(live), "n"(MASK));
}
// if (live & MASK) && spc == minVal:  lk = min(lk, posk ^ kx)
template <unsigned MASK>
__device__ __forceinline__ void lap_min_key(unsigned &lk, double spc, double minVal, unsigned live, int posk, int kx)
{
    asm("{\n\t.reg .pred q, p;\n\t.reg .b32 t, k;\n\t"
        "and.b32 t, %3, %4;\n\t"
        "setp.ne.u32 q, t, 0;\n\t"
        "setp.eq.and.f64 p, %1, %2, q;\n\t"
        "xor.b32 k, %5, %6;\n\t"
        "selp.b32 k, k, 0xffffffff, p;\n\t"
        "min.u32 %0, %0, k;\n\t}"
        : "+r"(lk) : "d"(spc), "d"(minVal), "r"(live), "n"(MASK), "r"(posk), "r"(kx));
}

// position field of the winning key back as pos << 19:  (key ^ ~(key >>s 31)) & mask  (kk = pos for an assigned column,
// bit 31 set; pos ^ 511 for an unassigned one)
__device__ __forceinline__ int lap_winner_pos(unsigned mk, unsigned mask)
{
    int r;
    asm("{\n\t.reg .b32 t;\n\t"
        "shr.s32 t, %1, 31;\n\t"
        "lop3.b32 %0, %1, t, %2, 0x82;\n\t}"
        : "=r"(r) : "r"(mk), "r"(mask));
    return r;
}

template <int S, int I = 0>
struct LapSlots {
    // slots [LO, HI): relax with r = base - v[slot] (see lap_relax_below)
    template <int LO, int HI>
    static __device__ __forceinline__ void relax_base_below(double (&spc)[2 * S], int (&pth)[2 * S], unsigned &below,
                                                            const double (&v)[2 * S], double base, int i, unsigned live, double minVal)
    {
        if constexpr (I < 2 * S) {
            if constexpr (I >= LO && I < HI)
                lap_relax_below<(1u << I)>(spc[I], pth[I], below, __dsub_rn(base, v[I]), i, live, minVal);
            LapSlots<S, I + 1>::template relax_base_below<LO, HI>(spc, pth, below, v, base, i, live, minVal);
        }
    }
    // slots [LO, HI) hold columns lane, lane + 32, ... of one group: relax the one that is column k of the group
    template <int LO, int HI>
    static __device__ __forceinline__ void relax_own_below(double (&spc)[2 * S], int (&pth)[2 * S], unsigned &below,
                                                           const double (&v)[2 * S], double base, int i, unsigned live, int k, int lane,
                                                           double minVal)
    {
        if constexpr (I < 2 * S) {
            if constexpr (I >= LO && I < HI)
                lap_relax_own_below<(1u << I)>(spc[I], pth[I], below, __dsub_rn(base, v[I]), i, live, minVal, k, lane + 32 * (I - LO));
            LapSlots<S, I + 1>::template relax_own_below<LO, HI>(spc, pth, below, v, base, i, live, k, lane, minVal);
        }
    }
    static __device__ __forceinline__ void relax_row_below(double (&spc)[2 * S], int (&pth)[2 * S], unsigned &below,
                                                           const double (&v)[2 * S], const float (&cs)[S], double minVal, double ui,
                                                           int i, unsigned live)
    {
        if constexpr (I < S) {
            const double r = __dsub_rn(__dsub_rn(__dadd_rn(minVal, (double)cs[I]), ui), v[I]);
            lap_relax_below<(1u << I)>(spc[I], pth[I], below, r, i, live, minVal);
            LapSlots<S, I + 1>::relax_row_below(spc, pth, below, v, cs, minVal, ui, i, live);
        }
    }
    static __device__ __forceinline__ void min_value2(double &lmA, double &lmB, const double (&spc)[2 * S], unsigned live)
    {
        if constexpr (I < 2 * S) {
            lap_min_value<(1u << I)>((I & 1) ? lmB : lmA, spc[I], live);
            LapSlots<S, I + 1>::min_value2(lmA, lmB, spc, live);
        }
    }
    static __device__ __forceinline__ void min_key(unsigned &lk, const double (&spc)[2 * S], double minVal, unsigned live,
                                                   const int (&posk)[2 * S], const int (&kx)[2 * S])
    {
        if constexpr (I < 2 * S) {
            lap_min_key<(1u << I)>(lk, spc[I], minVal, live, posk[I], kx[I]);
            LapSlots<S, I + 1>::min_key(lk, spc, minVal, live, posk, kx);
        }
    }
};

// ---------------------------------------------------------------------------------------------------------------
// LSAPE through the expanded (n1+n2)^2 LAP in SciPy's order (hungarian_lap ged_env.py:329-351; scipy rectangular_lsap).
// The expanded matrix is never formed: its real block comes from the cost store, row i's deletion entry from
// sm.cdel[i], column a's insertion entry from sm.cdel[n1 + a], everything else is +inf or 0 by construction.
// Columns live in registers, one per (lane, slot): real column a -> lane a & 31, slot a >> 5; deletion column n2+k ->
// lane k & 31, slot S + (k >> 5).  Tie-break key of a column (the smallest key wins):
//     assigned << 31 | kk << 19 | lane << 14 | slot << 9 | row
// with pos = the column's index in SciPy's `remaining` vector, kk = pos ^ 511 for an unassigned column (SciPy: the LAST
// unassigned minimum wins) and kk = pos for an assigned one (else the FIRST minimum wins; bit 31 puts every assigned column
// behind every unassigned one), row = row4col of an assigned column and the sentinel N -- a valid index of the padded u / cdel
// arrays -- for an unassigned one.  The layout is chosen for the bookkeeping after the reduction: the next row is key & 511
// without a decrement or a clamp, the slot field is five bits wide so that a wrapping shift takes 1 << slot straight from
// key >> 9, and the winner's position is (key ^ ~(key >>s 31)) & 511 << 19.  pos is unique per column, so the bits below kk
// never decide a comparison.
// `invalid` = the producer of the matrix saw NaN / -inf (scipy: RECTANGULAR_LSAP_INVALID).
// ---------------------------------------------------------------------------------------------------------------
template <int S, class Store>
__device__ int lsape_warp(const Store &store, bool invalid, int n1, int n2, const Slice &sm, unsigned long long *steps_out)
{
    const int lane = threadIdx.x & 31;
    const int N = n1 + n2;
    double *u = sm.u;
    if (invalid) return GED_ST_INVALID_COST;

    double v[2 * S];
    unsigned exist = 0;                      // bit s set: this lane's slot s holds a column of the expanded matrix
#pragma unroll
    for (int s = 0; s < S; ++s) {
        v[s] = 0.0;
        v[S + s] = 0.0;
        if (lane + 32 * s < n2) exist |= 1u << s;
        if (lane + 32 * s < n1) exist |= 1u << (S + s);
    }
    for (int t = lane; t < N; t += 32) { u[t] = 0.0; sm.row4col[t] = -1; sm.col4row[t] = -1; }
    __syncwarp();

    constexpr unsigned kPosShift = 19, kPosMask = 511u << kPosShift;
    const unsigned one = 1u;
    unsigned long long steps = 0;
    for (int cur = 0; cur < N; ++cur) {
        double spc[2 * S];
        int pth[2 * S], posk[2 * S], kx[2 * S];
#pragma unroll
        for (int s = 0; s < 2 * S; ++s) {
            const int j = (s < S) ? (lane + 32 * s) : (n2 + lane + 32 * (s - S));
            spc[s] = INFINITY;
            pth[s] = -1;
            posk[s] = (N - 1 - j) << kPosShift;
            int r4 = -1;
            if ((exist >> s) & 1u) r4 = sm.row4col[j];
            kx[s] = (int)(((r4 >= 0) ? 0x80000000u : kPosMask) | (unsigned)(lane << 14) | (unsigned)(s << 9) | (unsigned)((r4 >= 0) ? r4 : N));
            asm volatile("" : "+r"(kx[s]));          // keep it in a register (no re-derivation inside the scan loop)
        }
        unsigned live = exist;               // columns still in `remaining`
        double minVal = 0.0, rzmin = INFINITY;
        int i = cur, lastk = N << kPosShift;
        unsigned mk;
        bool infeasible = false;

        // operands of the row about to be scanned: row dual, the row's deletion / insertion entry and (shared-memory store)
        // its real entries are loaded one step ahead, right after the winner is known, so that the latency overlaps the
        // bookkeeping of the previous step
        bool real = i < n1;
        double ui = u[i];
        float cx = sm.cdel[i];
        float cs[S];
#pragma unroll
        for (int s = 0; s < S; ++s) cs[s] = 0.0f;
        if (Store::kPrefetch && real) store.row_issue(i, cs);

        for (;;) {
            unsigned keep = 0xffffffffu;     // 0 once this lane wrote a shortest-path cost below minVal (lap_relax_below)
            unsigned lk;
            if (real) {                      // real row: n2 substitution entries + its own deletion entry
                if (Store::kPrefetch) store.row_wait(i, cs);       // pairs with the row_issue of the previous step / prologue
                else store.row(i, cs);
                LapSlots<S>::relax_row_below(spc, pth, keep, v, cs, minVal, ui, i, live);
                const double rd = __dsub_rn(__dadd_rn(minVal, (double)cx), ui);
                LapSlots<S>::template relax_own_below<S, 2 * S>(spc, pth, keep, v, rd, i, live, i, lane, minVal);
            } else {                         // insertion row of column a0: its own entry + the zero block
                const int a0 = i - n1;
                const double ra = __dsub_rn(__dadd_rn(minVal, (double)cx), ui);
                LapSlots<S>::template relax_own_below<0, S>(spc, pth, keep, v, ra, i, live, a0, lane, minVal);
                // zero block: r = rz - v[j].  Rounding is monotone in rz, and after a scan with value rz0 every remaining
                // deletion column has spc <= fl(rz0 - v[j]); a later insertion row with rz >= rz0 cannot lower any of them.
                const double rz = __dsub_rn(__dadd_rn(minVal, 0.0), ui);
                if (rz < rzmin) {
                    rzmin = rz;
                    LapSlots<S>::template relax_base_below<S, 2 * S>(spc, pth, keep, v, rz, i, live, minVal);
                }
            }
            // Tie-break among the columns that attain the OLD minVal (scan-step shortcut, see lap_relax_below): the bid starts
            // from `keep`, so a lane that wrote a value below minVal bids key 0, below every real key.
            lk = keep;
            LapSlots<S>::min_key(lk, spc, minVal, live, posk, kx);
            mk = __reduce_min_sync(kFull, lk);
            if (mk - 1u >= 0xfffffffeu) {    // 0: some lane went below minVal; ~0: no remaining column attains minVal any more
                double lmA = INFINITY, lmB = INFINITY;          // two independent chains
                LapSlots<S>::min_value2(lmA, lmB, spc, live);
                const double lm = (lmB < lmA) ? lmB : lmA;
                unsigned ohi, olo;
                ord_from_double(lm, ohi, olo);
                const unsigned mhi = __reduce_min_sync(kFull, ohi);
                const unsigned mlo = __reduce_min_sync(kFull, ohi == mhi ? olo : 0xffffffffu);
                minVal = double_from_ord(mhi, mlo);
                lk = 0xffffffffu;
                LapSlots<S>::min_key(lk, spc, minVal, live, posk, kx);
                mk = __reduce_min_sync(kFull, lk);
                const bool inf = (minVal == INFINITY);          // nothing reachable: leave the loop, report after it
                infeasible |= inf;
                mk = inf ? (unsigned)N : mk;                    // row field N = "the search ends here"
            }

            int in = (int)(mk & 511u);                          // row matched with the winning column; N: it is unassigned
            asm volatile("" : "+r"(in));                        // (addresses from `in`: one IMAD each instead of a second decode)
            {   // operands of the next row (u[N], cdel[N] are padding)
                real = in < n1;
                ui = u[in];
                cx = sm.cdel[in];
                if (Store::kPrefetch && real) store.row_issue(in, cs);   // never leaves a fetch in flight past the search: N >= n1
            }
            // the winning column leaves `remaining`; the column at its end takes the freed position
            lastk -= 1 << kPosShift;
            const int posstark = lap_winner_pos(mk, kPosMask);
            // 1 << slot of the lane's own bid (ready before the reduction returns) leaves `live` in the lane that won
            asm("{\n\t.reg .pred p;\n\t.reg .b32 g;\n\t"
                "shf.l.wrap.b32 g, 0, %3, %4;\n\t"
                "setp.eq.u32 p, %1, %2;\n\t"
                "@p lop3.b32 %0, %0, g, 0, 0x30;\n\t}"
                : "+r"(live) : "r"(lk), "r"(mk), "r"(one), "r"(lk >> 9));
#pragma unroll
            for (int s = 0; s < 2 * S; ++s) posk[s] = (posk[s] == lastk) ? posstark : posk[s];
            if (in == N) break;
            i = in;
        }
        if (infeasible) return GED_ST_INFEASIBLE;
        steps += __reduce_add_sync(kFull, (unsigned)__popc(exist & ~live));   // every scan step removes exactly one column
        const int wl = (int)((mk >> 14) & 31u), ws = (int)((mk >> 9) & 15u);
        const int sink = (ws < S) ? (wl + 32 * ws) : (n2 + wl + 32 * (ws - S));

        // dual update (scipy solve(): u[cur] += minVal; u[i] += minVal - spc[col4row[i]]; v[j] -= minVal - spc[j]);
        // a removed column's spc is frozen at the minimum of the step that removed it
#pragma unroll
        for (int s = 0; s < 2 * S; ++s) {
            const bool scanned = ((exist & ~live) >> s) & 1u;
            if (scanned) {
                const int j = (s < S) ? (lane + 32 * s) : (n2 + lane + 32 * (s - S));
                const double delta = __dsub_rn(minVal, spc[s]);
                v[s] = __dsub_rn(v[s], delta);
                const int r4 = kx[s] & 511;
                if (j != sink) u[r4] = __dadd_rn(u[r4], delta);
                sm.path[j] = (int16_t)pth[s];
            }
        }
        if (lane == 0) u[cur] = __dadd_rn(u[cur], minVal);
        __syncwarp();
        if (lane == 0) {                     // augment along the predecessor chain
            int j = sink;
            for (;;) {
                const int r = sm.path[j];
                sm.row4col[j] = (int16_t)r;
                const int t = sm.col4row[r];
                sm.col4row[r] = (int16_t)j;
                j = t;
                if (r == cur) break;
            }
        }
        __syncwarp();
    }

    // fold the permutation back (ged_env.py:344-347)
    for (int i = lane; i < n1; i += 32) { const int c = sm.col4row[i]; sm.rowmatch[i] = (int16_t)(c < n2 ? c : n2); }
    for (int a = lane; a < n2; a += 32) { const int r = sm.row4col[a]; sm.colmatch[a] = (int16_t)(r < n1 ? r : n1); }
    __syncwarp();
    if (steps_out) *steps_out += steps;
    return GED_ST_OK;
}

// <cost, b> for the binary projection b = (rowmatch, colmatch) in the canonical order (oracle: matched_sum_cols):
// partial[c & 31] accumulates cost[t, c] over rows t ascending (c = matched column, n2 = deleted), then partial[a & 31]
// accumulates cost[n1, a] over inserted columns a ascending; fp64, xor butterfly.
template <int S, class Store>
__device__ double matched_cost_sum(const Store &store, const Slice &sm, int n1, int n2)
{
    const int lane = threadIdx.x & 31;
    double acc = 0.0;
    for (int t = 0; t < n1; ++t) {
        float cs[S];
        store.row(t, cs);
        const int c = sm.rowmatch[t];
        float term = sm.cdel[t];
#pragma unroll
        for (int s = 0; s < S; ++s) term = (c < n2 && (c >> 5) == s) ? cs[s] : term;
        if (lane == (c & 31)) acc = __dadd_rn(acc, (double)term);
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const int a = lane + 32 * s;
        if (a < n2 && sm.colmatch[a] == n1) acc = __dadd_rn(acc, (double)sm.cdel[n1 + a]);
    }
    return warp_butterfly(acc);
}

// Sum of the node costs of an assignment (ub, scoring): t < n1 -> (t, rowmatch[t]);  t >= n1 -> (n1, t - n1) if that
// column is inserted; partial[t & 31], fp64, xor butterfly (oracle: matched_sum).
__device__ inline double matched_node_cost(const Pair &P, const Slice &sm, const int16_t *rowmatch, const int16_t *colmatch)
{
    const int lane = threadIdx.x & 31;
    double acc = 0.0;
    for (int t = lane; t < P.N; t += 32) {
        double term = 0.0;
        if (t < P.n1) term = (double)node_cost_at(P, sm, t, rowmatch[t]);
        else {
            const int a = t - P.n1;
            if (colmatch[a] == P.n1) term = (double)node_cost_at(P, sm, P.n1, a);
        }
        acc = __dadd_rn(acc, term);
    }
    return warp_butterfly(acc);
}

// comp_ged of a binary x (ged_env.py:245-253) from the factors: node costs of the matched entries + one unit per
// ORDERED edge of either graph that the mapping does not preserve.
__device__ inline double score_assignment(const Pair &P, const Slice &sm, int nnz1, int nnz2,
                                          const int16_t *rowmatch, const int16_t *colmatch)
{
    const int lane = threadIdx.x & 31;
    int preserved = 0;
    for (int i = lane; i < P.n1; i += 32) {
        const int a = rowmatch[i];
        if (a < P.n2)
            for (int w = 0; w < P.W1; ++w) {
                uint32_t word = sm.bits1[i * P.W1 + w];
                while (word) {
                    const int j = w * 32 + __ffs(word) - 1;
                    word &= word - 1;
                    const int b = rowmatch[j];
                    if (b < P.n2 && bit_at(sm.bits2, P.W2, a, b)) ++preserved;
                }
            }
    }
    preserved = __reduce_add_sync(kFull, preserved);
    const double nodes = matched_node_cost(P, sm, rowmatch, colmatch);
    return __dadd_rn(nodes, (double)(nnz1 + nnz2 - 2 * preserved));
}

// maximum row popcount over both bit-packed adjacencies (warp-uniform)
__device__ inline int max_degree(const uint32_t *bits1, int R, int W1, const uint32_t *bits2, int C, int W2)
{
    const int lane = threadIdx.x & 31;
    int m = 0;
    for (int i = lane; i < R; i += 32) { int d = 0; for (int w = 0; w < W1; ++w) d += __popc(bits1[i * W1 + w]); m = max(m, d); }
    for (int a = lane; a < C; a += 32) { int d = 0; for (int w = 0; w < W2; ++w) d += __popc(bits2[a * W2 + w]); m = max(m, d); }
    return __reduce_max_sync(kFull, m);
}

__device__ inline int count_bits(const uint32_t *bits, int words)
{
    const int lane = threadIdx.x & 31;
    int c = 0;
    for (int w = lane; w < words; w += 32) c += __popc(bits[w]);
    return __reduce_add_sync(kFull, c);
}

// Route one element of an R x C cost matrix to where the LAP reads it from.  Every lane calls this for (i, lane + 32 t).
template <int S, class Store>
__device__ __forceinline__ void emit_cost(const Store &store, const Slice &sm, int n1, int n2, int i, int t, float c)
{
    const int a = (threadIdx.x & 31) + 32 * t;
    if (i < n1) {
        if (t < S) store.put(i, t, c);           // tcgen05.st is warp-collective: no lane predicate around it
        if (a == n2) sm.cdel[i] = c;
    } else if (a <= n2) sm.cdel[n1 + a] = c;
}

// heuristic_prediction_hun's node_cost_mat (ged_env.py:310-315) in closed form from node degrees.
template <int S, class Store>
__device__ inline void init_cost_closed_form(const Pair &P, const Slice &sm, const Store &store, float *trace)
{
    constexpr int T = S + 1;
    const int lane = threadIdx.x & 31;
    for (int i = lane; i < P.R; i += 32) {       // degrees into rs / cs (as floats; small integers)
        int d = 0;
        for (int w = 0; w < P.W1; ++w) d += __popc(sm.bits1[i * P.W1 + w]);
        sm.rs[i] = (float)d;
    }
    for (int a = lane; a < P.C; a += 32) {
        int d = 0;
        for (int w = 0; w < P.W2; ++w) d += __popc(sm.bits2[a * P.W2 + w]);
        sm.cs[a] = (float)d;
    }
    __syncwarp();
    for (int i = 0; i < P.R; ++i) {
        const int d1 = (int)sm.rs[i];
#pragma unroll
        for (int t = 0; t < T; ++t) {
            if (32 * t >= P.C) break;
            const int a = lane + 32 * t;
            const int d2 = (a < P.C) ? (int)sm.cs[a] : 0;
            int val;
            if (i < P.n1 && a < P.n2) { const int d = abs(d1 - d2) - 1; val = d > 0 ? d : 0; }
            else if (i < P.n1) val = d1;
            else if (a < P.n2) val = d2;
            else val = 0;
            emit_cost<S>(store, sm, P.n1, P.n2, i, t, (float)val);
            if (trace && a < P.C) trace[i * P.C + a] = (float)val;
        }
    }
    store.flush();
    __syncwarp();
}

__device__ __forceinline__ float dense_b(const Pair &P, const int16_t *rowmatch, const int16_t *colmatch, int i, int a)
{
    if (i < P.n1) return (rowmatch[i] == a) ? 1.0f : 0.0f;
    return (a < P.n2 && colmatch[a] == P.n1) ? 1.0f : 0.0f;
}

// ---------------------------------------------------------------------------------------------------------------
// cost = K vec(X) from the factors (replaces torch.mm(k, v), ged_env.py:269):
//     K vec(X) = A1 X M2^T + M1 X A2^T - 2 A1 X A2^T + Cn o X      (identity checked in the tests)
// organised around Y = X A2^T (Y[j,a] = sum_{b in N2(a)} X[j,b], ascending b) so that every inner term is a ROW gather:
// P[i,a] = sum_{j in N1(i)} X[j,a],  T3[i,a] = sum_{j in N1(i)} Y[j,a],  Q[i,a] = Y[i,a].
//   pass A: each row of X goes through a small double-buffered shared-memory row buffer for the random column gathers
//           of Y (neighbour lists of the lane's columns held in registers); row / column sums on the way; Y -> global;
//   pass B: per result row, the neighbour rows of X and Y are read with coalesced, mutually independent loads (one L2
//           round trip per row) and the row of cost goes straight to the LAP's cost store.
// Fused into pass B: s_vv = <cost, X> (fp64, partial per lane = a & 31, rows ascending: canonical order v2), the
// NaN / -inf check SciPy does on its input, and the optional trace copy.  T = number of 32-wide column tiles of a row.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void decode_neighbours(const uint32_t *row, int W, int (&nb)[kMaxDeg])
{
    int w = 0;
    uint32_t word = row[0];
#pragma unroll
    for (int k = 0; k < kMaxDeg; ++k) {
        while (word == 0u && w + 1 < W) word = row[++w];
        nb[k] = word ? (w * 32 + __ffs(word) - 1) : -1;
        word &= word - 1u;
    }
}

template <int S, class Store>
__device__ void kv_apply(const Pair &P, const Slice &sm, const float *X, float *Y, const Store &store, float *trace,
                         double &s_vv_out, bool &invalid_out)
{
    constexpr int T = S + 1;
    const int lane = threadIdx.x & 31;
    const int R = P.R, C = P.C, n1 = P.n1, n2 = P.n2;
    const bool fast = P.maxdeg <= kMaxDeg;
    const int rbs = (C + 3) & ~3;                              // row-buffer stride

    // ---- pass A --------------------------------------------------------------------------------------------------
    int nb2[T][kMaxDeg];
#pragma unroll
    for (int t = 0; t < T; ++t) {
        const int a = lane + 32 * t;
#pragma unroll
        for (int k = 0; k < kMaxDeg; ++k) nb2[t][k] = -1;
        if (fast && a < C) decode_neighbours(sm.bits2 + a * P.W2, P.W2, nb2[t]);
    }
    float csacc[T];
#pragma unroll
    for (int t = 0; t < T; ++t) csacc[t] = 0.0f;
    for (int j = 0; j < R; ++j) {
        float *buf = sm.rowbuf + (j & 1) * rbs;
        float part = 0.0f;
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const int a = lane + 32 * t;
            if (a < C) {
                const float x = X[j * C + a];
                buf[a] = x;
                csacc[t] = __fadd_rn(csacc[t], x);
                part = __fadd_rn(part, x);
            }
        }
        part = warp_butterfly(part);
        if (lane == 0) sm.rs[j] = part;
        __syncwarp();                                          // row buffer written -> gathers below
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const int a = lane + 32 * t;
            if (a < C) {
                float in = 0.0f;
                if (fast) {
#pragma unroll
                    for (int k = 0; k < kMaxDeg; ++k) if (nb2[t][k] >= 0) in = __fadd_rn(in, buf[nb2[t][k]]);
                } else {
                    for (int w = 0; w < P.W2; ++w) {
                        uint32_t word = sm.bits2[a * P.W2 + w];
                        while (word) { const int b = w * 32 + __ffs(word) - 1; word &= word - 1; in = __fadd_rn(in, buf[b]); }
                    }
                }
                Y[j * C + a] = in;
            }
        }
    }
#pragma unroll
    for (int t = 0; t < T; ++t) { const int a = lane + 32 * t; if (a < C) sm.cs[a] = csacc[t]; }
    __syncwarp();
    for (int i = lane; i < R; i += 32) {                       // rsP[i] = sum_{j in N1(i)} rs[j]
        float s = 0.0f;
        for (int w = 0; w < P.W1; ++w) {
            uint32_t word = sm.bits1[i * P.W1 + w];
            while (word) { const int j = w * 32 + __ffs(word) - 1; word &= word - 1; s = __fadd_rn(s, sm.rs[j]); }
        }
        sm.rsP[i] = s;
    }
    float csq[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {                              // csQ[a] = sum_{b in N2(a)} cs[b]   (kept in registers)
        const int a = lane + 32 * t;
        float s = 0.0f;
        if (a < C)
            for (int w = 0; w < P.W2; ++w) {
                uint32_t word = sm.bits2[a * P.W2 + w];
                while (word) { const int b = w * 32 + __ffs(word) - 1; word &= word - 1; s = __fadd_rn(s, sm.cs[b]); }
            }
        csq[t] = s;
    }
    __syncwarp();

    // ---- pass B --------------------------------------------------------------------------------------------------
    double svv = 0.0;
    bool bad = false;
    for (int i = 0; i < R; ++i) {
        int nb[kMaxDeg];
#pragma unroll
        for (int k = 0; k < kMaxDeg; ++k) nb[k] = -1;
        if (fast) decode_neighbours(sm.bits1 + i * P.W1, P.W1, nb);      // warp-uniform
        const float rsPi = sm.rsP[i];
#pragma unroll
        for (int t = 0; t < T; ++t) {
            if (32 * t >= C) break;
            const int a = lane + 32 * t;
            const bool in_row = a < C;
            float Pv = 0.0f, T3 = 0.0f, xia = 0.0f, Qv = 0.0f;
            if (in_row) {
                xia = X[i * C + a];
                Qv = Y[i * C + a];
                if (fast) {
                    float xv[kMaxDeg], yv[kMaxDeg];
#pragma unroll
                    for (int k = 0; k < kMaxDeg; ++k) {          // all loads first: they are independent
                        xv[k] = 0.0f;
                        yv[k] = 0.0f;
                        if (nb[k] >= 0) { xv[k] = X[nb[k] * C + a]; yv[k] = Y[nb[k] * C + a]; }
                    }
#pragma unroll
                    for (int k = 0; k < kMaxDeg; ++k)
                        if (nb[k] >= 0) { Pv = __fadd_rn(Pv, xv[k]); T3 = __fadd_rn(T3, yv[k]); }
                } else {
                    for (int w = 0; w < P.W1; ++w) {
                        uint32_t word = sm.bits1[i * P.W1 + w];
                        while (word) {
                            const int j = w * 32 + __ffs(word) - 1;
                            word &= word - 1;
                            Pv = __fadd_rn(Pv, X[j * C + a]);
                            T3 = __fadd_rn(T3, Y[j * C + a]);
                        }
                    }
                }
            }
            const float t1 = (a < n2) ? __fsub_rn(rsPi, Pv) : rsPi;
            const float t2 = (i < n1) ? __fsub_rn(csq[t], Qv) : csq[t];
            float sres = __fadd_rn(t1, t2);
            sres = __fsub_rn(sres, __fadd_rn(T3, T3));
            const float d = in_row ? __fmul_rn(node_cost_at(P, sm, i, a), xia) : 0.0f;
            const float c = __fadd_rn(sres, d);
            emit_cost<S>(store, sm, n1, n2, i, t, c);
            if (in_row) {
                svv = __dadd_rn(svv, __dmul_rn((double)c, (double)xia));
                if (!(i == n1 && a == n2)) bad |= (c != c) || (c == -INFINITY);
                if (trace) trace[i * C + a] = c;
            }
        }
    }
    store.flush();
    __syncwarp();
    s_vv_out = warp_butterfly(svv);
    invalid_out = __any_sync(kFull, bad);
}

// ---------------------------------------------------------------------------------------------------------------
// Kernel parameters (device pointers copied out of the C-ABI descriptors)
// ---------------------------------------------------------------------------------------------------------------
struct SolveParams {
    ged_pairs pairs;
    int B;
    const int32_t *base_idx, *edit_u, *edit_v;
    const uint8_t *score_adj1;
    int score_ld;
    const int64_t *score_adj1_off;
    const int32_t *score_idx;
    int max_iter, solver_kind;
    const float *x_init;
    long long mat_stride;
    float *workspace;
    int32_t *ws_locks;
    int ws_slots;
    long long ws_stride;
    const int32_t *order;
    int count;
    int32_t *work_counter;   // device int32, zero at launch: persistent warps pull candidates from it (nullptr: static assignment)
    int spread;              // static assignment: 1 = candidate blockIdx + warp * gridDim (spread over CTAs), 0 = CTA-major
    int tmem_warps;          // 0 or 4: warps 0..3 of every CTA keep their cost matrix in tensor memory
    int tmem_cols;           // columns allocated per CTA (power of two >= 32)
    int two_ended;           // tensor + global hybrid launches: the work queue is taken from both ends (see ged_solve_kernel)
    int ws_mats;             // matrices per scratch row: 2 (X, X A2^T) or 3 (+ a cost matrix for a global-memory warp)
    ged_solve_out out;
    SliceLayout layout;
};

__device__ inline void load_pair(const ged_pairs &pp, int p, const Slice &sm, Pair &P, int &status)
{
    const int lane = threadIdx.x & 31;
    P.n1 = pp.n1[p];
    P.n2 = pp.n2[p];
    P.R = P.n1 + 1;
    P.C = P.n2 + 1;
    P.N = P.n1 + P.n2;
    P.W1 = (P.n1 + 31) / 32 > 0 ? (P.n1 + 31) / 32 : 1;
    P.W2 = (P.n2 + 31) / 32 > 0 ? (P.n2 + 31) / 32 : 1;
    P.node_cost = pp.node_cost ? pp.node_cost + pp.node_cost_off[p] : nullptr;
    const bool bad1 = load_bits(pp.adj1 + pp.adj1_off[p], P.n1, pp.adj1_ld, P.W1, sm.bits1, lane);
    const bool bad2 = load_bits(pp.adj2 + pp.adj2_off[p], P.n2, pp.adj2_ld, P.W2, sm.bits2, lane);
    if (bad1 || bad2) status |= GED_ST_BAD_GRAPH;
    if (!P.node_cost) {
        const int32_t *l1 = pp.lab1 + pp.lab1_off[p], *l2 = pp.lab2 + pp.lab2_off[p];
        for (int i = lane; i < P.n1; i += 32) sm.lab1[i] = l1[i];
        for (int a = lane; a < P.n2; a += 32) sm.lab2[a] = l2[a];
    }
    __syncwarp();
    P.maxdeg = max_degree(sm.bits1, P.R, P.W1, sm.bits2, P.C, P.W2);
}

// One candidate solve by one warp: GEDenv.step :110-124 -> solve_feasible_ged :140-158 -> ipfp_ged :256-289.
template <int S, class Store>
__device__ void solve_one(const SolveParams &prm, int b, int slot, const Slice &sm, const Store &store_in)
{
    const int lane = threadIdx.x & 31;
    const ged_solve_out &out = prm.out;
    const int p = prm.base_idx ? prm.base_idx[b] : b;
    Pair P;
    int status = 0;
    load_pair(prm.pairs, p, sm, P, status);
    const int n1 = P.n1, n2 = P.n2;

    // GEDenv.step edge toggle (ged_env.py:110-121)
    int eu = prm.edit_u ? prm.edit_u[b] : -1, ev = prm.edit_v ? prm.edit_v[b] : -1;
    const bool has_edit = eu >= 0 && ev >= 0 && eu != ev && eu < n1 && ev < n1;
    bool was_present = false;
    if (has_edit) {
        was_present = bit_at(sm.bits1, P.W1, eu, ev) || bit_at(sm.bits1, P.W1, ev, eu);
        __syncwarp();
        if (lane == 0) {
            const uint32_t m_uv = 1u << (ev & 31), m_vu = 1u << (eu & 31);
            uint32_t &w_uv = sm.bits1[eu * P.W1 + (ev >> 5)];
            w_uv = was_present ? (w_uv & ~m_uv) : (w_uv | m_uv);
            uint32_t &w_vu = sm.bits1[ev * P.W1 + (eu >> 5)];
            w_vu = was_present ? (w_vu & ~m_vu) : (w_vu | m_vu);
        }
        __syncwarp();
        P.maxdeg = max_degree(sm.bits1, P.R, P.W1, sm.bits2, P.C, P.W2);
    }
    const int nnz1 = count_bits(sm.bits1, P.R * P.W1), nnz2 = count_bits(sm.bits2, P.C * P.W2);

    // scratch row for the fractional iterate X and Y = X A2^T: claimed among the rows of the solves in flight
    int ws_row = b;
    if (prm.workspace && prm.ws_slots > 0) {
        if (lane == 0) {
            int h = slot % prm.ws_slots;
            while (atomicCAS(prm.ws_locks + h, 0, 1) != 0) h = (h + 1 == prm.ws_slots) ? 0 : h + 1;
            ws_row = h;
        }
        ws_row = __shfl_sync(kFull, ws_row, 0);
    }
    float *X = prm.workspace ? prm.workspace + (long long)ws_row * prm.ws_mats * prm.ws_stride : nullptr;
    float *Y = X ? X + prm.ws_stride : nullptr;
    Store store = store_in;                  // a global-memory warp's cost matrix is the third matrix of its scratch row
    store.bind_global(X && prm.ws_mats >= 3 ? X + 2 * prm.ws_stride : nullptr);
    unsigned long long lap_steps = 0;
    int iters = 0;
    double best_ub = INFINITY;
#ifdef GED_TIMING
    long long t_lap = 0, t_kv = 0;
    const long long t_begin = clock64();
#endif

    if (status == 0) {
        if (prm.x_init) {
            const float *x0 = prm.x_init + (long long)b * prm.mat_stride;
            for (int e = lane; e < P.R * P.C; e += 32) X[e] = x0[e];
            for (int i = lane; i < n1; i += 32) sm.bestrow[i] = (int16_t)n2;
            for (int a = lane; a < n2; a += 32) sm.bestcol[a] = (int16_t)n1;
            __syncwarp();
        } else {
            // hungarian_ged (ged_env.py:303-326): v0 = LSAPE(node_cost_mat)
            init_cost_closed_form<S>(P, sm, store, out.tr_init_cost ? out.tr_init_cost + (long long)b * prm.mat_stride : nullptr);
            status |= lsape_warp<S>(store, false, n1, n2, sm, &lap_steps);
            if (status == 0) {
                for (int i = lane; i < n1; i += 32) {
                    sm.bestrow[i] = sm.rowmatch[i];
                    if (out.tr_init_rowmatch) out.tr_init_rowmatch[(long long)b * out.match_stride_1 + i] = sm.rowmatch[i];
                }
                for (int a = lane; a < n2; a += 32) {
                    sm.bestcol[a] = sm.colmatch[a];
                    if (out.tr_init_colmatch) out.tr_init_colmatch[(long long)b * out.match_stride_2 + a] = sm.colmatch[a];
                }
                __syncwarp();
                if (prm.solver_kind == GED_SOLVER_IPFP)
                    for (int i = 0; i < P.R; ++i)
                        for (int a = lane; a < P.C; a += 32) X[i * P.C + a] = dense_b(P, sm.rowmatch, sm.colmatch, i, a);
                __syncwarp();
            }
        }
    }

    if (status == 0 && prm.solver_kind == GED_SOLVER_IPFP) {
        for (int it = 0; it < prm.max_iter; ++it) {
            iters = it + 1;
            double s_vv;                                                 // = last_v^T K last_v, ged_env.py:283
            bool invalid;
#ifdef GED_TIMING
            const long long tk0 = clock64();
#endif
            kv_apply<S>(P, sm, X, Y, store, out.tr_cost ? out.tr_cost + ((long long)b * prm.max_iter + it) * prm.mat_stride : nullptr,
                        s_vv, invalid);                                  // ged_env.py:269
#ifdef GED_TIMING
            t_kv += clock64() - tk0;
            const long long tl0 = clock64();
#endif
            status |= lsape_warp<S>(store, invalid, n1, n2, sm, &lap_steps);   // ged_env.py:270
#ifdef GED_TIMING
            t_lap += clock64() - tl0;
#endif
            if (status != 0) break;
            const double ub = score_assignment(P, sm, nnz1, nnz2, sm.rowmatch, sm.colmatch);   // :271
            if (ub < best_ub) {                                          // :272-274
                best_ub = ub;
                for (int i = lane; i < n1; i += 32) sm.bestrow[i] = sm.rowmatch[i];
                for (int a = lane; a < n2; a += 32) sm.bestcol[a] = sm.colmatch[a];
            }
            const double s_vb = matched_cost_sum<S>(store, sm, n1, n2);
            const double alpha = __dsub_rn(s_vb, s_vv);                  // :275 (K symmetric: v^T K = (K v)^T)
            const double beta = __dadd_rn(__dsub_rn(ub, __dadd_rn(s_vb, s_vb)), s_vv);   // :277
            const double t0 = __ddiv_rn(-alpha, beta);                   // :278
            const float t0f = (float)t0;
            const bool jump = (beta <= 0.0) || (t0f >= 1.0f);            // :279
            float *trx = out.tr_x_after ? out.tr_x_after + ((long long)b * prm.max_iter + it) * prm.mat_stride : nullptr;
            for (int i = 0; i < P.R; ++i) {
                const int rm = (i < n1) ? sm.rowmatch[i] : -1;
                for (int a = lane; a < P.C; a += 32) {
                    const int e = i * P.C + a;
                    const float bv = (i < n1) ? ((rm == a) ? 1.0f : 0.0f) : ((a < n2 && sm.colmatch[a] == n1) ? 1.0f : 0.0f);
                    float xn;
                    if (jump) xn = bv;                                   // :280
                    else { const float x = X[e]; xn = __fadd_rn(x, __fmul_rn(t0f, __fsub_rn(bv, x))); }   // :282
                    X[e] = xn;
                    if (trx) trx[e] = xn;
                }
            }
            if (out.tr_scalars && lane == 0) {
                double *d = out.tr_scalars + ((long long)b * prm.max_iter + it) * 6;
                d[0] = s_vv; d[1] = s_vb; d[2] = ub; d[3] = alpha; d[4] = beta; d[5] = t0;
            }
            if (out.tr_rowmatch)
                for (int i = lane; i < n1; i += 32) out.tr_rowmatch[((long long)b * prm.max_iter + it) * out.match_stride_1 + i] = sm.rowmatch[i];
            if (out.tr_colmatch)
                for (int a = lane; a < n2; a += 32) out.tr_colmatch[((long long)b * prm.max_iter + it) * out.match_stride_2 + a] = sm.colmatch[a];
            __syncwarp();
            if (fabs(__dsub_rn(s_vv, s_vb)) / s_vv < 1e-3) break;        // :284-285
        }
    }

    // score the returned solution on the edited pair and on the base pair (ged_env.py:122,158)
    float ged_edit = NAN, ged_base = NAN;
    if (status == 0) {
        ged_edit = (float)score_assignment(P, sm, nnz1, nnz2, sm.bestrow, sm.bestcol);
        ged_base = ged_edit;
        if (prm.score_adj1) {               // score on the original pair's graph 1 (ori_k)
            __syncwarp();
            const bool bad = load_bits(prm.score_adj1 + prm.score_adj1_off[prm.score_idx[b]], n1, prm.score_ld, P.W1, sm.bits1, lane);
            __syncwarp();
            if (bad) status |= GED_ST_BAD_GRAPH;
            const int nnz1o = count_bits(sm.bits1, P.R * P.W1);
            ged_base = bad ? NAN : (float)score_assignment(P, sm, nnz1o, nnz2, sm.bestrow, sm.bestcol);
        } else if (has_edit) {
            __syncwarp();
            if (lane == 0) {
                const uint32_t m_uv = 1u << (ev & 31), m_vu = 1u << (eu & 31);
                uint32_t &w_uv = sm.bits1[eu * P.W1 + (ev >> 5)];
                w_uv = was_present ? (w_uv | m_uv) : (w_uv & ~m_uv);
                uint32_t &w_vu = sm.bits1[ev * P.W1 + (eu >> 5)];
                w_vu = was_present ? (w_vu | m_vu) : (w_vu & ~m_vu);
            }
            __syncwarp();
            ged_base = (float)score_assignment(P, sm, nnz1 + (was_present ? 2 : -2), nnz2, sm.bestrow, sm.bestcol);
        }
    }
#ifdef GED_TIMING
    if (lane == 0) printf("solve %d (%s): total %lld cycles, lap %lld (%.1f%%), kv %lld (%.1f%%), lap steps %llu -> %.1f cycles/step\n", b,
                          Store::kPrefetch ? "smem" : "tmem", clock64() - t_begin, t_lap, 100.0 * t_lap / (clock64() - t_begin), t_kv,
                          100.0 * t_kv / (clock64() - t_begin), lap_steps, (double)t_lap / (double)lap_steps);
#endif
    if (lane == 0) {
        out.ged[b] = ged_base;
        if (out.ged_edited) out.ged_edited[b] = ged_edit;
        if (out.iters) out.iters[b] = iters;
        if (out.status) out.status[b] = status;
        if (out.tr_lap_steps) out.tr_lap_steps[b] = lap_steps;
    }
    if (prm.workspace && prm.ws_slots > 0) {     // release the scratch row (all of this warp's accesses to it are done)
        __syncwarp();
        if (lane == 0) { __threadfence(); atomicExch(prm.ws_locks + ws_row, 0); }
    }
    if (status == 0) {
        for (int i = lane; i < n1; i += 32) out.rowmatch[(long long)b * out.match_stride_1 + i] = sm.bestrow[i];
        for (int a = lane; a < n2; a += 32) out.colmatch[(long long)b * out.match_stride_2 + a] = sm.bestcol[a];
        if (out.x_out) {
            float *xo = out.x_out + (long long)b * prm.mat_stride;
            for (int i = 0; i < P.R; ++i)
                for (int a = lane; a < P.C; a += 32) xo[i * P.C + a] = dense_b(P, sm.bestrow, sm.bestcol, i, a);
        }
    }
}

// CTA = [tmem_warps TMEM-backed solves][remaining warps: shared-memory-backed solves].  Dynamic shared memory:
// tmem_warps slices without a cost block, then the others with one.
// MODE = kModeSmem: no tensor-memory instruction in the binary at all (a CTA of a kernel that may allocate tensor memory
// holds the SM's allocation permit until it relinquishes it, which keeps other CTAs off the SM).
constexpr int kMaxCtaWarps = 8;          // 4 tensor-memory warps + at most 4 shared-memory (or global-memory) warps

// Spread (latency-bound) launches give every CTA ONE candidate in the first round.  Warp w of a CTA runs on SM sub-partition
// w % 4 (tensor memory ties a warp to the lane quarter of its sub-partition), so if that candidate always went to warp 0,
// every active warp of an SM would share sub-partition 0 while the other three idle.  Each CTA therefore takes a ticket
// from a per-SM counter (returned on exit, so the array is all zero between launches) and rotates its warps by it.
#ifndef GED_SPREAD_ROTATE
#define GED_SPREAD_ROTATE 1
#endif
static __device__ int g_sm_tickets[256];
// cost store of the CTA's warps: all shared memory / all tensor memory / tensor + shared memory / tensor + global memory
constexpr int kModeSmem = 0, kModeTmem = 1, kModeHybrid = 2, kModeHybridG = 3;

// S <= 3 (pairs up to 96 nodes): two 8-warp CTAs, or four 4-warp CTAs, must fit the register file -> at most 128 registers
// per thread (registers are allocated per CTA in units of 4 warps: a 7-warp CTA at 136 registers is alone on its SM).
#ifndef GED_TMEM_CTAS_S2
#define GED_TMEM_CTAS_S2 0               // experiment: > 0 = launch bounds (128 threads, that many CTAs per SM) for the S <= 2 tensor-memory kernel
#endif
#define GED_LB_THREADS(S_, MODE_) ((GED_TMEM_CTAS_S2 > 0 && (S_) <= 2 && (MODE_) < kModeHybrid) ? 32 * kTmemWarps : 32 * kMaxCtaWarps)
#define GED_LB_CTAS(S_, MODE_) ((GED_TMEM_CTAS_S2 > 0 && (S_) <= 2 && (MODE_) < kModeHybrid) ? GED_TMEM_CTAS_S2 : ((S_) <= 3 ? 2 : 1))
template <int S, int MODE>
__global__ void __launch_bounds__(GED_LB_THREADS(S, MODE), GED_LB_CTAS(S, MODE)) ged_solve_kernel(SolveParams prm)
{
    constexpr bool TM = MODE != kModeSmem;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t tmem_base_slot;
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int nt = TM ? prm.tmem_warps : 0;
    uint32_t tmem_base = 0;
    if constexpr (TM) {                      // the only CTA-wide synchronisation: tensor-memory allocation
        if (warp == 0) {
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tmem_base_slot);
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"((uint32_t)prm.tmem_cols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem_base = tmem_base_slot;
    }
    // Work distribution: static (one candidate per warp) or, with a work counter, persistent warps pulling the next
    // candidate of the list until it is exhausted (the list is ordered longest-first by the host layer).
    const int gwarp = blockIdx.x * nwarps + warp;
    const int m1 = prm.pairs.max_n1, m2 = prm.pairs.max_n2;
    const bool tm = TM && warp < nt;         // this warp keeps its cost matrix in tensor memory (lane quarter = warp id)
    // the other warps: in shared memory behind their slice, or (kModeHybridG) in global memory -- slice only
    const Slice sm = (tm || MODE == kModeHybridG)
                         ? carve(smem_raw + (size_t)warp * prm.layout.bytes, prm.layout, m1, m2, false)
                         : carve(smem_raw + (size_t)nt * prm.layout.bytes + (size_t)(warp - nt) * (prm.layout.bytes + prm.layout.cost_bytes),
                                 prm.layout, m1, m2, true);
    uint32_t taddr = tmem_base + ((uint32_t)(32 * warp) << 16);
    asm volatile("" : "+r"(taddr));          // opaque: kept in a register instead of being re-derived from %tid per row fetch
    int vwarp = warp;                        // position of this warp in the static assignment
#if GED_SPREAD_ROTATE
    __shared__ int ticket_slot;
    unsigned smid = 0;
    if (prm.spread) {
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        if (threadIdx.x == 0) ticket_slot = atomicAdd(&g_sm_tickets[smid & 255u], 1);
        __syncthreads();
        const int rot = ticket_slot % (nt > 0 ? nt : nwarps);
        if (nt > 0) vwarp = warp < nt ? (warp + nt - rot) % nt : warp;          // rotate among the tensor-memory warps
        else vwarp = (warp + nwarps - rot) % nwarps;
    }
#endif
    for (int slot = prm.spread ? blockIdx.x + vwarp * gridDim.x : gwarp;;) {      // static assignment (see launch_solve_as)
        if (prm.work_counter) {
            if (MODE == kModeHybridG && prm.two_ended) {
                // Two-ended queue: the tensor-memory warps take the list from its front (heaviest candidates), the slower
                // global-memory warps from its back (lightest).  One int32 holds both cursors (front in the low half, back
                // in the high half): an atomic add returns a unique (front, back) state, the taker's index is its own
                // half, and it is valid while front + back < count -- of two takers that would meet at the same index,
                // the later one sees the earlier one's increment and finds the list exhausted.
                if ((threadIdx.x & 31) == 0) {
                    const unsigned old = (unsigned)atomicAdd(prm.work_counter, tm ? 1 : (1 << 16));
                    const int f = (int)(old & 0xffffu), bk = (int)(old >> 16);
                    slot = (f + bk < prm.count) ? (tm ? f : prm.count - 1 - bk) : prm.count;
                }
            } else if ((threadIdx.x & 31) == 0) slot = atomicAdd(prm.work_counter, 1);
            slot = __shfl_sync(kFull, slot, 0);
        }
        if (slot >= prm.count) break;
        const int b = prm.order ? prm.order[slot] : slot;
        const int n2 = prm.pairs.n2[prm.base_idx ? prm.base_idx[b] : b];
        if constexpr (MODE == kModeTmem) {
            TmemCost<S> store{taddr};
            solve_one<S>(prm, b, gwarp, sm, store);
        } else if constexpr (MODE == kModeHybrid) {
            AnyCost<S, SmemCost<S>> store{TmemCost<S>{taddr}, SmemCost<S>{sm.cost, n2}, tm};
            solve_one<S>(prm, b, gwarp, sm, store);
        } else if constexpr (MODE == kModeHybridG) {
            AnyCost<S, GmemCost<S>> store{TmemCost<S>{taddr}, GmemCost<S>{nullptr, n2}, tm};
            solve_one<S>(prm, b, gwarp, sm, store);
        } else {
            SmemCost<S> store{sm.cost, n2};
            solve_one<S>(prm, b, gwarp, sm, store);
        }
        if (!prm.work_counter) break;
        __syncwarp();
    }
    if constexpr (TM) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)prm.tmem_cols) : "memory");
    }
#if GED_SPREAD_ROTATE
    if (prm.spread) {
        __syncthreads();                     // the ticket is held until every warp of the CTA is done
        if (threadIdx.x == 0) atomicSub(&g_sm_tickets[smid & 255u], 1);
    }
#endif
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-PER-CANDIDATE ("team") solve for latency-bound calls: a rollout / beam-search step of a few hundred candidates
// (reference ged_ppo_bihyb_eval.py:95-100, ged_ppo_bihyb_train.py:294-298) lasts as long as its LONGEST solve, and a
// solve owned by one warp runs at the speed of one warp's dependent instruction stream while > 95 % of the device idles.
// Here W = S warps of one CTA share one candidate:
//   * LAP: warp w owns the real columns [32 w, 32 w + 32) and the deletion columns n2 + [32 w, 32 w + 32) -- one real and
//     one deletion slot per lane instead of S of each, i.e. the scan step of the S = 1 kernel whatever the pair's size.
//     Per scan step every warp bids its best tie-break key; the bids meet in shared memory (one STS, one CTA barrier, one
//     128-bit LDS) and every warp continues with the same winner.  A step that has to recompute the minimum exchanges the
//     order-preserving 64-bit image of its local minimum the same way, then the keys again.  Row duals, assignments and
//     the cost matrix are shared (shared memory); column duals and shortest-path costs stay in the owner's registers.
//   * K.X and the line-search update are split by ROWS over the warps (row j mod W); the reductions whose canonical order
//     runs across rows (column sums, <K v, v>) are re-done in that order from the stored rows.
//   * Everything that is a function of shared data only (objective, scalars, convergence test) is computed by every warp
//     redundantly, so that control flow is uniform without broadcasts.
// Results are bit-identical to the one-warp kernels: same arithmetic on the same operands in the same order, same scan
// order, same tie-break keys (tests/test_gpu_team.py).  The cost matrix lives in shared memory (no tensor memory: a
// latency-bound call has the SM's 227 KB almost to itself).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kTeamMaxW = 7;

struct TeamXch {                             // exchange area of a team (one per CTA), 16-byte aligned
    unsigned key[2][8];                      // tie-break bids of the warps, double buffered
    unsigned long long val[2][8];            // order-preserving images of the warps' local minima, double buffered
    int bad, status, ws_row, next;           // NaN / -inf flag of a K.X pass; per-candidate status; scratch row; work-queue slot
};

template <int W>
__device__ __forceinline__ unsigned team_min_key(unsigned x, TeamXch *xch, int &xp)
{
    if ((threadIdx.x & 31) == 0) xch->key[xp][threadIdx.x >> 5] = x;
    __syncthreads();
    const uint4 a = *reinterpret_cast<const uint4 *>(xch->key[xp]);       // entries >= W stay at 0xffffffff
    unsigned m = min(min(a.x, a.y), min(a.z, a.w));
    if constexpr (W > 4) {
        const uint4 b = *reinterpret_cast<const uint4 *>(xch->key[xp] + 4);
        m = min(m, min(min(b.x, b.y), min(b.z, b.w)));
    }
    xp ^= 1;                                 // consecutive exchanges never share a buffer: one barrier per exchange is enough
    return m;
}

template <int W>
__device__ __forceinline__ unsigned long long team_min_val(unsigned long long x, TeamXch *xch, int &xp)
{
    if ((threadIdx.x & 31) == 0) xch->val[xp][threadIdx.x >> 5] = x;
    __syncthreads();
    unsigned long long m = xch->val[xp][0];
#pragma unroll
    for (int k = 1; k < W; ++k) { const unsigned long long o = xch->val[xp][k]; m = o < m ? o : m; }
    xp ^= 1;
    return m;
}

// LSAPE of the cost matrix in shared memory (`cost`: n1 x n2 block, row stride n2; sm.cdel: the deletion / insertion
// entries) by the W warps of the CTA -- lsape_warp with the columns dealt out to the warps, see above.
template <int W>
__device__ int lsape_team(const float *cost, bool invalid, int n1, int n2, const Slice &sm, TeamXch *xch, int &xp,
                          unsigned long long *steps_out)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int N = n1 + n2;
    double *u = sm.u;
    if (invalid) return GED_ST_INVALID_COST;                  // CTA-uniform
    const int aR = lane + 32 * w;                             // this lane's real column; n2 + aR is its deletion column
    double v[2] = {0.0, 0.0};
    unsigned exist = 0;
    if (aR < n2) exist |= 1u;
    if (aR < n1) exist |= 2u;
    for (int t = threadIdx.x; t < N; t += 32 * W) { u[t] = 0.0; sm.row4col[t] = -1; sm.col4row[t] = -1; }
    __syncthreads();
    const float *ccol = cost + aR;                            // element (i, aR) at ccol[i * n2]

    // tie-break keys as in lsape_warp (assigned << 31 | kk << 19 | lane << 14 | slot << 9 | row), slot = s << 4 | the column
    // group's index among the 2 W groups of the team (s = 0: real columns of warp w, 1: its deletion columns)
    constexpr unsigned kPosShift = 19, kPosMask = 511u << kPosShift;
    unsigned long long steps = 0;
    for (int cur = 0; cur < N; ++cur) {
        double spc[2];
        int pth[2], posk[2], kx[2];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int j = s ? n2 + aR : aR;
            spc[s] = INFINITY;
            pth[s] = -1;
            posk[s] = (N - 1 - j) << kPosShift;
            int r4 = -1;
            if ((exist >> s) & 1u) r4 = sm.row4col[j];
            kx[s] = (int)(((r4 >= 0) ? 0x80000000u : kPosMask) | (unsigned)(lane << 14) | (unsigned)((s ? 16 + W + w : w) << 9) |
                          (unsigned)((r4 >= 0) ? r4 : N));
            asm volatile("" : "+r"(kx[s]));
        }
        unsigned live = exist;
        double minVal = 0.0, rzmin = INFINITY;
        int i = cur, lastk = N << kPosShift;
        unsigned mk;
        int nsteps = 0;
        bool infeasible = false;
        bool real = i < n1;
        double ui = u[i];
        float cx = sm.cdel[i];
        float cs[1] = {0.0f};
        if (real) cs[0] = ccol[i * n2];

        for (;;) {
            ++nsteps;
            unsigned keep = 0xffffffffu;     // 0 once this lane wrote a shortest-path cost below minVal (lap_relax_below)
            unsigned lk;
            if (real) {                      // real row: this warp's 32 substitution entries; the deletion entry at its owner
                LapSlots<1>::relax_row_below(spc, pth, keep, v, cs, minVal, ui, i, live);
                const double rd = __dsub_rn(__dadd_rn(minVal, (double)cx), ui);
                lap_relax_own_below<2u>(spc[1], pth[1], keep, __dsub_rn(rd, v[1]), i, live, minVal, i, aR);
            } else {                         // insertion row of column a0: its own entry at its owner + the zero block
                const int a0 = i - n1;
                const double ra = __dsub_rn(__dadd_rn(minVal, (double)cx), ui);
                lap_relax_own_below<1u>(spc[0], pth[0], keep, __dsub_rn(ra, v[0]), i, live, minVal, a0, aR);
                const double rz = __dsub_rn(__dadd_rn(minVal, 0.0), ui);
                if (rz < rzmin) {
                    rzmin = rz;
                    LapSlots<1>::template relax_base_below<1, 2>(spc, pth, keep, v, rz, i, live, minVal);
                }
            }
            lk = keep;                       // a lane that went below minVal bids key 0
            LapSlots<1>::min_key(lk, spc, minVal, live, posk, kx);
            mk = team_min_key<W>(__reduce_min_sync(kFull, lk), xch, xp);
            if (mk - 1u >= 0xfffffffeu) {    // some value fell below minVal (bid 0) or no column attains it (all bids ~0)
                double lmA = INFINITY, lmB = INFINITY;
                LapSlots<1>::min_value2(lmA, lmB, spc, live);
                const double lm = (lmB < lmA) ? lmB : lmA;
                unsigned ohi, olo;
                ord_from_double(lm, ohi, olo);
                const unsigned mhi = __reduce_min_sync(kFull, ohi);
                const unsigned mlo = __reduce_min_sync(kFull, ohi == mhi ? olo : 0xffffffffu);
                const unsigned long long g = team_min_val<W>(((unsigned long long)mhi << 32) | mlo, xch, xp);
                minVal = double_from_ord((unsigned)(g >> 32), (unsigned)g);
                lk = 0xffffffffu;
                LapSlots<1>::min_key(lk, spc, minVal, live, posk, kx);
                mk = team_min_key<W>(__reduce_min_sync(kFull, lk), xch, xp);
                const bool inf = (minVal == INFINITY);
                infeasible |= inf;
                mk = inf ? (unsigned)N : mk;                    // row field N = "the search ends here"
            }

            int in = (int)(mk & 511u);                          // row matched with the winning column; N: it is unassigned
            asm volatile("" : "+r"(in));                        // (addresses from `in`: one IMAD each instead of a second decode)
            {   // operands of the next row (u[N], cdel[N] are padding)
                real = in < n1;
                ui = u[in];
                cx = sm.cdel[in];
                if (real) cs[0] = ccol[in * n2];
            }
            lastk -= 1 << kPosShift;
            const int posstark = lap_winner_pos(mk, kPosMask);
            live &= ~((lk == mk) ? ((lk & (16u << 9)) ? 2u : 1u) : 0u);
#pragma unroll
            for (int s = 0; s < 2; ++s) posk[s] = (posk[s] == lastk) ? posstark : posk[s];
            if (in == N) break;
            i = in;
        }
        if (infeasible) return GED_ST_INFEASIBLE;
        steps += nsteps;
        const int wl = (int)((mk >> 14) & 31u), ws = (int)((mk >> 9) & 15u);
        const int sink = (ws < W) ? (wl + 32 * ws) : (n2 + wl + 32 * (ws - W));

#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const bool scanned = ((exist & ~live) >> s) & 1u;
            if (scanned) {
                const int j = s ? n2 + aR : aR;
                const double delta = __dsub_rn(minVal, spc[s]);
                v[s] = __dsub_rn(v[s], delta);
                const int r4 = kx[s] & 511;
                if (j != sink) u[r4] = __dadd_rn(u[r4], delta);          // distinct rows: every column has its own row4col
                sm.path[j] = (int16_t)pth[s];
            }
        }
        if (threadIdx.x == 0) u[cur] = __dadd_rn(u[cur], minVal);
        __syncthreads();
        if (threadIdx.x == 0) {              // augment along the predecessor chain
            int j = sink;
            for (;;) {
                const int r = sm.path[j];
                sm.row4col[j] = (int16_t)r;
                const int t = sm.col4row[r];
                sm.col4row[r] = (int16_t)j;
                j = t;
                if (r == cur) break;
            }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < n1; i += 32 * W) { const int c = sm.col4row[i]; sm.rowmatch[i] = (int16_t)(c < n2 ? c : n2); }
    for (int a = threadIdx.x; a < n2; a += 32 * W) { const int r = sm.row4col[a]; sm.colmatch[a] = (int16_t)(r < n1 ? r : n1); }
    __syncthreads();
    if (steps_out) *steps_out += steps;
    return GED_ST_OK;
}

// element (i, a) of the R x C cost matrix from where the LAP keeps it (emit_cost's routing, read back)
__device__ __forceinline__ float team_cost_at(const float *cost, const Slice &sm, int n1, int n2, int i, int a)
{
    if (i < n1) return (a < n2) ? cost[i * n2 + a] : sm.cdel[i];
    return sm.cdel[n1 + a];
}

// kv_apply by the W warps of a CTA: rows dealt out round robin; results and every rounding identical to kv_apply.
template <int S, int W>
__device__ void kv_team(const Pair &P, const Slice &sm, float *rowbuf, const float *X, float *Y, const SmemCost<S> &store,
                        float *trace, TeamXch *xch, double &s_vv_out, bool &invalid_out)
{
    constexpr int T = S + 1;                                   // 32-column tiles of a row; W = warps of the CTA
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int R = P.R, C = P.C, n1 = P.n1, n2 = P.n2;
    const bool fast = P.maxdeg <= kMaxDeg;
    const int rbs = (C + 3) & ~3;
    if (threadIdx.x == 0) xch->bad = 0;

    // ---- pass A: Y = X A2^T and the row sums, row j by warp j mod W -------------------------------------------------
    {
        int nb2[T][kMaxDeg];
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const int a = lane + 32 * t;
#pragma unroll
            for (int k = 0; k < kMaxDeg; ++k) nb2[t][k] = -1;
            if (fast && a < C) decode_neighbours(sm.bits2 + a * P.W2, P.W2, nb2[t]);
        }
        int par = 0;
        for (int j = w; j < R; j += W, par ^= 1) {
            float *buf = rowbuf + par * rbs;
            float part = 0.0f;
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const int a = lane + 32 * t;
                if (a < C) {
                    const float x = X[j * C + a];
                    buf[a] = x;
                    part = __fadd_rn(part, x);
                }
            }
            part = warp_butterfly(part);
            if (lane == 0) sm.rs[j] = part;
            __syncwarp();
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const int a = lane + 32 * t;
                if (a < C) {
                    float in = 0.0f;
                    if (fast) {
#pragma unroll
                        for (int k = 0; k < kMaxDeg; ++k) if (nb2[t][k] >= 0) in = __fadd_rn(in, buf[nb2[t][k]]);
                    } else {
                        for (int ww = 0; ww < P.W2; ++ww) {
                            uint32_t word = sm.bits2[a * P.W2 + ww];
                            while (word) { const int b = ww * 32 + __ffs(word) - 1; word &= word - 1; in = __fadd_rn(in, buf[b]); }
                        }
                    }
                    Y[j * C + a] = in;
                }
            }
        }
    }
    // column sums in the canonical order (rows ascending), 32-column tiles dealt out to the warps
    for (int t = w; t < T; t += W) {
        const int a = lane + 32 * t;
        if (a < C) {
            float s = 0.0f;
            for (int j = 0; j < R; ++j) s = __fadd_rn(s, X[j * C + a]);
            sm.cs[a] = s;
        }
    }
    __syncthreads();                                           // Y, rs, cs complete
    for (int i = threadIdx.x; i < R; i += 32 * W) {            // rsP[i] = sum_{j in N1(i)} rs[j]
        float s = 0.0f;
        for (int ww = 0; ww < P.W1; ++ww) {
            uint32_t word = sm.bits1[i * P.W1 + ww];
            while (word) { const int j = ww * 32 + __ffs(word) - 1; word &= word - 1; s = __fadd_rn(s, sm.rs[j]); }
        }
        sm.rsP[i] = s;
    }
    float csq[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {                              // csQ[a] = sum_{b in N2(a)} cs[b]: every warp, all tiles
        const int a = lane + 32 * t;
        float s = 0.0f;
        if (a < C)
            for (int ww = 0; ww < P.W2; ++ww) {
                uint32_t word = sm.bits2[a * P.W2 + ww];
                while (word) { const int b = ww * 32 + __ffs(word) - 1; word &= word - 1; s = __fadd_rn(s, sm.cs[b]); }
            }
        csq[t] = s;
    }
    __syncthreads();                                           // rsP complete

    // ---- pass B: row i of the cost matrix by warp i mod W -------------------------------------------------------------
    bool bad = false;
    for (int i = w; i < R; i += W) {
        int nb[kMaxDeg];
#pragma unroll
        for (int k = 0; k < kMaxDeg; ++k) nb[k] = -1;
        if (fast) decode_neighbours(sm.bits1 + i * P.W1, P.W1, nb);
        const float rsPi = sm.rsP[i];
#pragma unroll
        for (int t = 0; t < T; ++t) {
            if (32 * t >= C) break;
            const int a = lane + 32 * t;
            const bool in_row = a < C;
            float Pv = 0.0f, T3 = 0.0f, xia = 0.0f, Qv = 0.0f;
            if (in_row) {
                xia = X[i * C + a];
                Qv = Y[i * C + a];
                if (fast) {
                    float xv[kMaxDeg], yv[kMaxDeg];
#pragma unroll
                    for (int k = 0; k < kMaxDeg; ++k) {
                        xv[k] = 0.0f;
                        yv[k] = 0.0f;
                        if (nb[k] >= 0) { xv[k] = X[nb[k] * C + a]; yv[k] = Y[nb[k] * C + a]; }
                    }
#pragma unroll
                    for (int k = 0; k < kMaxDeg; ++k)
                        if (nb[k] >= 0) { Pv = __fadd_rn(Pv, xv[k]); T3 = __fadd_rn(T3, yv[k]); }
                } else {
                    for (int ww = 0; ww < P.W1; ++ww) {
                        uint32_t word = sm.bits1[i * P.W1 + ww];
                        while (word) {
                            const int j = ww * 32 + __ffs(word) - 1;
                            word &= word - 1;
                            Pv = __fadd_rn(Pv, X[j * C + a]);
                            T3 = __fadd_rn(T3, Y[j * C + a]);
                        }
                    }
                }
            }
            const float t1 = (a < n2) ? __fsub_rn(rsPi, Pv) : rsPi;
            const float t2 = (i < n1) ? __fsub_rn(csq[t], Qv) : csq[t];
            float sres = __fadd_rn(t1, t2);
            sres = __fsub_rn(sres, __fadd_rn(T3, T3));
            const float d = in_row ? __fmul_rn(node_cost_at(P, sm, i, a), xia) : 0.0f;
            const float c = __fadd_rn(sres, d);
            emit_cost<S>(store, sm, n1, n2, i, t, c);
            if (in_row) {
                if (!(i == n1 && a == n2)) bad |= (c != c) || (c == -INFINITY);
                if (trace) trace[i * C + a] = c;
            }
        }
    }
    if (__any_sync(kFull, bad) && lane == 0) atomicOr(&xch->bad, 1);
    __syncthreads();                                           // cost matrix complete
    invalid_out = xch->bad != 0;
    // s_vv = <cost, X> in the canonical order (per lane: rows ascending, tiles ascending; fp64), by every warp
    // (the row of X for step i + 1 is requested before row i is consumed: X is global, one L2 round trip per row otherwise)
    double svv = 0.0;
    float xn[T];
#pragma unroll
    for (int t = 0; t < T; ++t) { const int a = lane + 32 * t; xn[t] = (a < C) ? X[a] : 0.0f; }
    for (int i = 0; i < R; ++i) {
        float xc[T];
#pragma unroll
        for (int t = 0; t < T; ++t) {
            xc[t] = xn[t];
            const int a = lane + 32 * t;
            xn[t] = (i + 1 < R && a < C) ? X[(i + 1) * C + a] : 0.0f;
        }
#pragma unroll
        for (int t = 0; t < T; ++t) {
            if (32 * t >= C) break;
            const int a = lane + 32 * t;
            if (a < C) svv = __dadd_rn(svv, __dmul_rn((double)team_cost_at(store.base, sm, n1, n2, i, a), (double)xc[t]));
        }
    }
    s_vv_out = warp_butterfly(svv);
}

// One candidate solve by the W warps of a CTA (solve_one for a team).  S = 32-column slots of the pair: S == W is the team
// proper (LAP columns dealt out to the warps); S == 1 < W serves pairs of up to 32 nodes, whose LAP has nothing to deal out --
// warp 0 runs the one-warp LAP (lsape_warp<1>) while the K.X passes, the line-search update and the set-up (19-24 % of a lone
// warp's solve at n = 25) are still split over the W warps.
template <int W, int S>
__device__ int team_lsape(const SmemCost<S> &store, bool invalid, int n1, int n2, const Slice &sm, TeamXch *xch, int &xp,
                          unsigned long long *steps)
{
    if constexpr (S == W) return lsape_team<W>(store.base, invalid, n1, n2, sm, xch, xp, steps);
    else {
        if (threadIdx.x < 32) {
            const int st = lsape_warp<S>(store, invalid, n1, n2, sm, steps);
            if (threadIdx.x == 0) xch->status = st;
        }
        __syncthreads();
        return xch->status;
    }
}

template <int W, int S>
__device__ void solve_team(const SolveParams &prm, int b, const Slice &sm, float *rowbuf, TeamXch *xch, int &xp)
{
    static_assert(S == W || S == 1, "team shapes: one warp per 32-column slot, or helpers around a one-slot LAP");
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nthr = 32 * W;
    const ged_solve_out &out = prm.out;
    const int p = prm.base_idx ? prm.base_idx[b] : b;
    const SmemCost<S> store{sm.cost, prm.pairs.n2[p]};
    Pair P;
    if (threadIdx.x == 0) xch->status = 0;
    __syncthreads();
    if (w == 0) {
        int st = 0;
        load_pair(prm.pairs, p, sm, P, st);
        if (st && lane == 0) xch->status = st;
    }
    __syncthreads();
    {   // the scalars of the pair, by every warp (load_pair filled shared memory)
        const ged_pairs &pp = prm.pairs;
        P.n1 = pp.n1[p];
        P.n2 = pp.n2[p];
        P.R = P.n1 + 1;
        P.C = P.n2 + 1;
        P.N = P.n1 + P.n2;
        P.W1 = (P.n1 + 31) / 32 > 0 ? (P.n1 + 31) / 32 : 1;
        P.W2 = (P.n2 + 31) / 32 > 0 ? (P.n2 + 31) / 32 : 1;
        P.node_cost = pp.node_cost ? pp.node_cost + pp.node_cost_off[p] : nullptr;
    }
    int status = xch->status;
    const int n1 = P.n1, n2 = P.n2;

    // GEDenv.step edge toggle (ged_env.py:110-121)
    const int eu = prm.edit_u ? prm.edit_u[b] : -1, ev = prm.edit_v ? prm.edit_v[b] : -1;
    const bool has_edit = eu >= 0 && ev >= 0 && eu != ev && eu < n1 && ev < n1;
    bool was_present = false;
    if (has_edit) {
        was_present = bit_at(sm.bits1, P.W1, eu, ev) || bit_at(sm.bits1, P.W1, ev, eu);
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t m_uv = 1u << (ev & 31), m_vu = 1u << (eu & 31);
            uint32_t &w_uv = sm.bits1[eu * P.W1 + (ev >> 5)];
            w_uv = was_present ? (w_uv & ~m_uv) : (w_uv | m_uv);
            uint32_t &w_vu = sm.bits1[ev * P.W1 + (eu >> 5)];
            w_vu = was_present ? (w_vu & ~m_vu) : (w_vu | m_vu);
        }
        __syncthreads();
    }
    P.maxdeg = max_degree(sm.bits1, P.R, P.W1, sm.bits2, P.C, P.W2);
    const int nnz1 = count_bits(sm.bits1, P.R * P.W1), nnz2 = count_bits(sm.bits2, P.C * P.W2);

    // scratch row for X and Y = X A2^T (global, L2 resident): claimed by one thread
    if (threadIdx.x == 0) {
        int h = b;
        if (prm.workspace && prm.ws_slots > 0) {
            h = blockIdx.x % prm.ws_slots;
            while (atomicCAS(prm.ws_locks + h, 0, 1) != 0) h = (h + 1 == prm.ws_slots) ? 0 : h + 1;
        }
        xch->ws_row = h;
    }
    __syncthreads();
    const int ws_row = xch->ws_row;
    float *X = prm.workspace ? prm.workspace + (long long)ws_row * prm.ws_mats * prm.ws_stride : nullptr;
    float *Y = X ? X + prm.ws_stride : nullptr;
    unsigned long long lap_steps = 0;
    int iters = 0;
    double best_ub = INFINITY;

    if (status == 0) {
        if (prm.x_init) {
            const float *x0 = prm.x_init + (long long)b * prm.mat_stride;
            for (int e = threadIdx.x; e < P.R * P.C; e += nthr) X[e] = x0[e];
            for (int i = threadIdx.x; i < n1; i += nthr) sm.bestrow[i] = (int16_t)n2;
            for (int a = threadIdx.x; a < n2; a += nthr) sm.bestcol[a] = (int16_t)n1;
            __syncthreads();
        } else {
            // hungarian_ged (ged_env.py:303-326): v0 = LSAPE(node_cost_mat); the closed form is cheap: warp 0 alone
            if (w == 0) init_cost_closed_form<S>(P, sm, store, out.tr_init_cost ? out.tr_init_cost + (long long)b * prm.mat_stride : nullptr);
            __syncthreads();
            status |= team_lsape<W, S>(store, false, n1, n2, sm, xch, xp, &lap_steps);
            if (status == 0) {
                for (int i = threadIdx.x; i < n1; i += nthr) {
                    sm.bestrow[i] = sm.rowmatch[i];
                    if (out.tr_init_rowmatch) out.tr_init_rowmatch[(long long)b * out.match_stride_1 + i] = sm.rowmatch[i];
                }
                for (int a = threadIdx.x; a < n2; a += nthr) {
                    sm.bestcol[a] = sm.colmatch[a];
                    if (out.tr_init_colmatch) out.tr_init_colmatch[(long long)b * out.match_stride_2 + a] = sm.colmatch[a];
                }
                if (prm.solver_kind == GED_SOLVER_IPFP)
                    for (int i = w; i < P.R; i += W)
                        for (int a = lane; a < P.C; a += 32) X[i * P.C + a] = dense_b(P, sm.rowmatch, sm.colmatch, i, a);
                __syncthreads();
            }
        }
    }

    if (status == 0 && prm.solver_kind == GED_SOLVER_IPFP) {
        for (int it = 0; it < prm.max_iter; ++it) {
            iters = it + 1;
            double s_vv;
            bool invalid;
            kv_team<S, W>(P, sm, rowbuf, X, Y, store, out.tr_cost ? out.tr_cost + ((long long)b * prm.max_iter + it) * prm.mat_stride : nullptr,
                       xch, s_vv, invalid);                                                    // ged_env.py:269
            status |= team_lsape<W, S>(store, invalid, n1, n2, sm, xch, xp, &lap_steps);      // :270
            if (status != 0) break;
            // scalars: functions of shared data only -> every warp computes them, control flow stays uniform
            const double ub = score_assignment(P, sm, nnz1, nnz2, sm.rowmatch, sm.colmatch);   // :271
            const bool better = ub < best_ub;                                                 // :272-274
            if (better) best_ub = ub;
            const double s_vb = matched_cost_sum<S>(store, sm, n1, n2);
            const double alpha = __dsub_rn(s_vb, s_vv);                                        // :275
            const double beta = __dadd_rn(__dsub_rn(ub, __dadd_rn(s_vb, s_vb)), s_vv);         // :277
            const double t0 = __ddiv_rn(-alpha, beta);                                         // :278
            const float t0f = (float)t0;
            const bool jump = (beta <= 0.0) || (t0f >= 1.0f);                                  // :279
            if (better && w == 0) {
                for (int i = lane; i < n1; i += 32) sm.bestrow[i] = sm.rowmatch[i];
                for (int a = lane; a < n2; a += 32) sm.bestcol[a] = sm.colmatch[a];
            }
            float *trx = out.tr_x_after ? out.tr_x_after + ((long long)b * prm.max_iter + it) * prm.mat_stride : nullptr;
            for (int i = w; i < P.R; i += W) {
                const int rm = (i < n1) ? sm.rowmatch[i] : -1;
                for (int a = lane; a < P.C; a += 32) {
                    const int e = i * P.C + a;
                    const float bv = (i < n1) ? ((rm == a) ? 1.0f : 0.0f) : ((a < n2 && sm.colmatch[a] == n1) ? 1.0f : 0.0f);
                    float xn;
                    if (jump) xn = bv;                                                         // :280
                    else { const float x = X[e]; xn = __fadd_rn(x, __fmul_rn(t0f, __fsub_rn(bv, x))); }   // :282
                    X[e] = xn;
                    if (trx) trx[e] = xn;
                }
            }
            if (w == 0) {
                if (out.tr_scalars && lane == 0) {
                    double *d = out.tr_scalars + ((long long)b * prm.max_iter + it) * 6;
                    d[0] = s_vv; d[1] = s_vb; d[2] = ub; d[3] = alpha; d[4] = beta; d[5] = t0;
                }
                if (out.tr_rowmatch)
                    for (int i = lane; i < n1; i += 32) out.tr_rowmatch[((long long)b * prm.max_iter + it) * out.match_stride_1 + i] = sm.rowmatch[i];
                if (out.tr_colmatch)
                    for (int a = lane; a < n2; a += 32) out.tr_colmatch[((long long)b * prm.max_iter + it) * out.match_stride_2 + a] = sm.colmatch[a];
            }
            __syncthreads();                                     // X rows of the other warps, bestrow / bestcol
            if (fabs(__dsub_rn(s_vv, s_vb)) / s_vv < 1e-3) break;                              // :284-285
        }
    }
    __syncthreads();

    // scoring on the edited pair and on the base pair (ged_env.py:122,158) and the outputs: warp 0
    if (w == 0) {
        float ged_edit = NAN, ged_base = NAN;
        if (status == 0) {
            ged_edit = (float)score_assignment(P, sm, nnz1, nnz2, sm.bestrow, sm.bestcol);
            ged_base = ged_edit;
            if (prm.score_adj1) {
                __syncwarp();
                const bool bad = load_bits(prm.score_adj1 + prm.score_adj1_off[prm.score_idx[b]], n1, prm.score_ld, P.W1, sm.bits1, lane);
                __syncwarp();
                if (bad) status |= GED_ST_BAD_GRAPH;
                const int nnz1o = count_bits(sm.bits1, P.R * P.W1);
                ged_base = bad ? NAN : (float)score_assignment(P, sm, nnz1o, nnz2, sm.bestrow, sm.bestcol);
            } else if (has_edit) {
                __syncwarp();
                if (lane == 0) {
                    const uint32_t m_uv = 1u << (ev & 31), m_vu = 1u << (eu & 31);
                    uint32_t &w_uv = sm.bits1[eu * P.W1 + (ev >> 5)];
                    w_uv = was_present ? (w_uv | m_uv) : (w_uv & ~m_uv);
                    uint32_t &w_vu = sm.bits1[ev * P.W1 + (eu >> 5)];
                    w_vu = was_present ? (w_vu | m_vu) : (w_vu & ~m_vu);
                }
                __syncwarp();
                ged_base = (float)score_assignment(P, sm, nnz1 + (was_present ? 2 : -2), nnz2, sm.bestrow, sm.bestcol);
            }
        }
        if (lane == 0) {
            out.ged[b] = ged_base;
            if (out.ged_edited) out.ged_edited[b] = ged_edit;
            if (out.iters) out.iters[b] = iters;
            if (out.status) out.status[b] = status;
            if (out.tr_lap_steps) out.tr_lap_steps[b] = lap_steps;
        }
        if (prm.workspace && prm.ws_slots > 0) {
            __syncwarp();
            if (lane == 0) { __threadfence(); atomicExch(prm.ws_locks + ws_row, 0); }
        }
        if (status == 0) {
            for (int i = lane; i < n1; i += 32) out.rowmatch[(long long)b * out.match_stride_1 + i] = sm.bestrow[i];
            for (int a = lane; a < n2; a += 32) out.colmatch[(long long)b * out.match_stride_2 + a] = sm.bestcol[a];
            if (out.x_out) {
                float *xo = out.x_out + (long long)b * prm.mat_stride;
                for (int i = 0; i < P.R; ++i)
                    for (int a = lane; a < P.C; a += 32) xo[i * P.C + a] = dense_b(P, sm.bestrow, sm.bestcol, i, a);
            }
        }
    }
    __syncthreads();                                             // the slice is reused by the CTA's next candidate
}

// Dynamic shared memory of a team CTA: [slice][cost block][W double-buffered X rows][exchange area]
__host__ __device__ inline int team_rowbuf_bytes(int max_n2) { return 2 * 4 * align_up(max_n2 + 1, 4); }
__host__ __device__ inline int team_smem_bytes(const SliceLayout &L, int W, int max_n2)
{
    return L.bytes + L.cost_bytes + W * team_rowbuf_bytes(max_n2) + (int)sizeof(TeamXch);
}

template <int W, int S>
__global__ void __launch_bounds__(32 * W) ged_solve_team_kernel(SolveParams prm)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int w = threadIdx.x >> 5;
    const int m1 = prm.pairs.max_n1, m2 = prm.pairs.max_n2;
    const Slice sm = carve(smem_raw, prm.layout, m1, m2, true);
    float *rowbuf = reinterpret_cast<float *>(smem_raw + prm.layout.bytes + prm.layout.cost_bytes + w * team_rowbuf_bytes(m2));
    TeamXch *xch = reinterpret_cast<TeamXch *>(smem_raw + prm.layout.bytes + prm.layout.cost_bytes + W * team_rowbuf_bytes(m2));
    if (threadIdx.x < 16) xch->key[threadIdx.x >> 3][threadIdx.x & 7] = 0xffffffffu;
    __syncthreads();
    int xp = 0;
    for (int slot = blockIdx.x;; slot += gridDim.x) {
        if (prm.work_counter) {                  // persistent CTAs pulling candidates (longest first) from the work queue
            if (threadIdx.x == 0) xch->next = atomicAdd(prm.work_counter, 1);
            __syncthreads();
            slot = xch->next;
        }
        if (slot >= prm.count) break;
        const int b = prm.order ? prm.order[slot] : slot;
        solve_team<W, S>(prm, b, sm, rowbuf, xch, xp);
    }
}

// ---- stand-alone stages (shared-memory cost store) ----------------------------------------------------------------
struct LsapeParams {
    int B;
    const int32_t *n1, *n2;
    const float *cost;
    long long mat_stride;
    int32_t *rowmatch, *colmatch;
    int ms1, ms2;
    float *lb;
    int32_t *status;
    int max_n1, max_n2;
    SliceLayout layout;
};

template <int S>
__global__ void ged_lsape_kernel(LsapeParams prm)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int T = S + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x * (blockDim.x >> 5) + warp;
    if (b >= prm.B) return;
    const Slice sm = carve(smem_raw + (size_t)warp * (prm.layout.bytes + prm.layout.cost_bytes), prm.layout, prm.max_n1, prm.max_n2, true);
    const int n1 = prm.n1[b], n2 = prm.n2[b], C = n2 + 1;
    SmemCost<S> store{sm.cost, n2};
    const float *src = prm.cost + (long long)b * prm.mat_stride;
    bool bad = false;
    for (int i = 0; i <= n1; ++i)
#pragma unroll
        for (int t = 0; t < T; ++t) {
            if (32 * t >= C) break;
            const int a = lane + 32 * t;
            const float c = (a < C) ? src[i * C + a] : 0.0f;
            if (a < C && !(i == n1 && a == n2)) bad |= (c != c) || (c == -INFINITY);
            emit_cost<S>(store, sm, n1, n2, i, t, c);
        }
    store.flush();
    const bool invalid = __any_sync(kFull, bad);
    const int st = lsape_warp<S>(store, invalid, n1, n2, sm, nullptr);
    if (st == 0) {
        for (int i = lane; i < n1; i += 32) prm.rowmatch[(long long)b * prm.ms1 + i] = sm.rowmatch[i];
        for (int a = lane; a < n2; a += 32) prm.colmatch[(long long)b * prm.ms2 + a] = sm.colmatch[a];
        if (prm.lb) {
            const double s = matched_cost_sum<S>(store, sm, n1, n2);     // ged_env.py:349
            if (lane == 0) prm.lb[b] = (float)s;
        }
    } else if (prm.lb && lane == 0) prm.lb[b] = NAN;
    if (prm.status && lane == 0) prm.status[b] = st;
}

struct KvParams {
    ged_pairs pairs;
    const float *x;
    float *cost;
    float *scratch;          // Y = X A2^T scratch, same layout as x / cost
    long long mat_stride;
    SliceLayout layout;
};

template <int S>
__global__ void ged_kv_kernel(KvParams prm)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5;
    const int b = blockIdx.x * (blockDim.x >> 5) + warp;
    if (b >= prm.pairs.num_pairs) return;
    const Slice sm = carve(smem_raw + (size_t)warp * (prm.layout.bytes + prm.layout.cost_bytes), prm.layout, prm.pairs.max_n1,
                           prm.pairs.max_n2, true);
    Pair P;
    int status = 0;
    load_pair(prm.pairs, b, sm, P, status);
    SmemCost<S> store{sm.cost, P.n2};
    double s_vv;
    bool invalid;
    kv_apply<S>(P, sm, prm.x + (long long)b * prm.mat_stride, prm.scratch + (long long)b * prm.mat_stride, store,
                prm.cost + (long long)b * prm.mat_stride, s_vv, invalid);         // the result is taken from the trace hook
}

#ifdef GED_TU_MAIN
struct ScoreParams {
    ged_pairs pairs;
    int B;
    const int32_t *base_idx, *rowmatch, *colmatch;
    int ms1, ms2;
    float *ged;
    SliceLayout layout;
};

__global__ void ged_score_kernel(ScoreParams prm)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x * (blockDim.x >> 5) + warp;
    if (b >= prm.B) return;
    const Slice sm = carve(smem_raw + (size_t)warp * prm.layout.bytes, prm.layout, prm.pairs.max_n1, prm.pairs.max_n2, false);
    Pair P;
    int status = 0;
    load_pair(prm.pairs, prm.base_idx ? prm.base_idx[b] : b, sm, P, status);
    for (int i = lane; i < P.n1; i += 32) sm.rowmatch[i] = (int16_t)prm.rowmatch[(long long)b * prm.ms1 + i];
    for (int a = lane; a < P.n2; a += 32) sm.colmatch[a] = (int16_t)prm.colmatch[(long long)b * prm.ms2 + a];
    __syncwarp();
    const int nnz1 = count_bits(sm.bits1, P.R * P.W1), nnz2 = count_bits(sm.bits2, P.C * P.W2);
    const double s = score_assignment(P, sm, nnz1, nnz2, sm.rowmatch, sm.colmatch);
    if (lane == 0) prm.ged[b] = status ? NAN : (float)s;
}

// x^T K x against a dense K: the HBM-bound form of comp_ged (ged_env.py:245-253).  One CTA per `rows_per_cta`
// rows of K; each warp streams rows with float4 loads (x staged in shared memory), the per-row dot is weighted by
// x[row] and accumulated per CTA; a second tiny pass adds the CTA partials in a fixed order (deterministic).
constexpr int kQfThreads = 256;
__global__ void __launch_bounds__(kQfThreads) ged_quadform_partial_kernel(int nk, const float *__restrict__ k, long long k_stride,
                                                                         const float *__restrict__ x, float *__restrict__ partial,
                                                                         int rows_per_cta, int ctas_per_mat)
{
    extern __shared__ __align__(16) float xs[];
    const int mat = blockIdx.y, chunk = blockIdx.x;
    const float *K = k + (long long)mat * k_stride;
    const float *xv = x + (long long)mat * nk;
    for (int e = threadIdx.x; e < nk; e += blockDim.x) xs[e] = xv[e];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int row0 = chunk * rows_per_cta;
    const int row1 = min(nk, row0 + rows_per_cta);
    const bool vec_ok = (nk % 4 == 0) && ((reinterpret_cast<uintptr_t>(K) & 15) == 0);
    float acc = 0.0f;
    for (int r = row0 + warp; r < row1; r += nwarps) {
        const float xr = xs[r];
        if (xr == 0.0f) continue;                      // binary x: only the matched rows are streamed
        const float *row = K + (long long)r * nk;
        float d = 0.0f;
        if (vec_ok) {
            const float4 *row4 = reinterpret_cast<const float4 *>(row);
            const float4 *xs4 = reinterpret_cast<const float4 *>(xs);
            for (int c = lane; c < nk / 4; c += 32) {
                const float4 kv = __ldg(row4 + c);
                const float4 xx = xs4[c];
                d += kv.x * xx.x + kv.y * xx.y + kv.z * xx.z + kv.w * xx.w;
            }
        } else {
            for (int c = lane; c < nk; c += 32) d += __ldg(row + c) * xs[c];
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) d += __shfl_xor_sync(kFull, d, off);
        acc += xr * d;
    }
    __shared__ float wsum[kQfThreads / 32];
    if (lane == 0) wsum[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.0f;
        for (int w = 0; w < nwarps; ++w) s += wsum[w];
        partial[(long long)mat * ctas_per_mat + chunk] = s;
    }
}

__global__ void ged_quadform_final_kernel(const float *partial, int ctas_per_mat, float *out, int B)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float s = 0.0f;
    for (int c = 0; c < ctas_per_mat; ++c) s += partial[(long long)b * ctas_per_mat + c];
    out[b] = s;
}

// ---- log-domain Sinkhorn (the projection step of the RRWM / graduated-assignment baselines) -----------------------------
// Reference utils/sinkhorn.py:143-239 (Sinkhorn.forward_log, non-batched branch): log_s = s / tau, then `iters` alternating
// passes -- even: every row minus its log-sum-exp, odd: every column minus its log-sum-exp -- over the whole rows x cols
// matrix, and exp at the end.  One CTA per matrix, the matrix in shared memory for all passes (odd pitch: the column pass
// walks down a column with consecutive lanes), one warp per row / column, max-subtracted sums through warp shuffles.
constexpr int kSinkhornThreads = 256;

__device__ __forceinline__ float warp_max(float x)
{
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) x = fmaxf(x, __shfl_xor_sync(kFull, x, off));
    return x;
}
__device__ __forceinline__ float warp_sum(float x)
{
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) x = __fadd_rn(x, __shfl_xor_sync(kFull, x, off));
    return x;
}

__global__ void __launch_bounds__(kSinkhornThreads) ged_sinkhorn_kernel(const int32_t *rows, const int32_t *cols, int ld_r, int ld_c,
                                                                       const float *__restrict__ s, const float *__restrict__ tau,
                                                                       int iters, float *__restrict__ out, int pitch)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *m = reinterpret_cast<float *>(smem_raw);
    const int b = blockIdx.x, R = rows[b], C = cols[b];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const float *src = s + (long long)b * ld_r * ld_c;
    float *dst = out + (long long)b * ld_r * ld_c;
    const float t = tau[b];
    for (int e = threadIdx.x; e < R * C; e += blockDim.x) {
        const int r = e / C, c = e - r * C;
        m[r * pitch + c] = __fdiv_rn(src[r * ld_c + c], t);
    }
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
        const bool by_row = (it & 1) == 0;
        const int lines = by_row ? R : C, len = by_row ? C : R;
        const int line_stride = by_row ? pitch : 1, elem_stride = by_row ? 1 : pitch;
        for (int l = warp; l < lines; l += nwarps) {
            float *p = m + l * line_stride;
            float mx = -INFINITY;
            for (int k = lane; k < len; k += 32) mx = fmaxf(mx, p[k * elem_stride]);
            mx = warp_max(mx);
            if (mx == -INFINITY) continue;               // an all -inf line stays as it is
            float acc = 0.0f;
            for (int k = lane; k < len; k += 32) acc = __fadd_rn(acc, expf(__fsub_rn(p[k * elem_stride], mx)));
            const float lse = __fadd_rn(mx, logf(warp_sum(acc)));
            for (int k = lane; k < len; k += 32) p[k * elem_stride] = __fsub_rn(p[k * elem_stride], lse);
        }
        __syncthreads();
    }
    for (int e = threadIdx.x; e < ld_r * ld_c; e += blockDim.x) {
        const int r = e / ld_c, c = e - r * ld_c;
        dst[e] = (r < R && c < C) ? expf(m[r * pitch + c]) : 0.0f;
    }
}

#endif  // GED_TU_MAIN

// ---------------------------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------------------------
static int slots_for(int max_n1, int max_n2)
{
    const int m = max_n1 > max_n2 ? max_n1 : max_n2;
    const int S = (m + 31) / 32;
    return S < 1 ? 1 : S;
}

static int env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

// Stand-alone stage kernels: all warps of a CTA have identical slices.
template <typename KernelT, typename ParamsT>
int launch_uniform(KernelT kernel, const ParamsT &prm, int count, int slice_bytes, cudaStream_t stream, const char *name)
{
    int wpb = (slice_bytes + 1024) * 32 <= kMaxSmemBytes / 2 ? 2 : 1;      // 32 CTAs/SM is the hardware cap
    const int smem = wpb * slice_bytes;
    if (smem > kMaxSmemBytes) return fail(GED_E_TOO_LARGE, "pair too large for one warp's shared-memory slice");
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail_cuda(e, name);
    kernel<<<(count + wpb - 1) / wpb, wpb * 32, smem, stream>>>(prm);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda(e, name);
    return GED_OK;
}

// Shape of the hybrid CTA of the solve kernel (see the file header): 4 tensor-memory-backed warps + k shared-memory-
// backed warps, sized so that shared memory and tensor-memory columns run out together.
struct CtaShape { int tmem_warps, tmem_cols, smem_warps, smem_bytes, gmem_warps; };

// allow_gmem: the caller's scratch rows hold a third matrix (ged_solve_desc.workspace_mats >= 3), so the warp slots that
// registers and the launch bounds leave after the shared-memory warps can be filled with global-memory warps.
static CtaShape plan_cta(const SliceLayout &L, int max_n1, int S, int regs_per_thread, bool prefer_smem = false, bool allow_gmem = false)
{
    CtaShape c{0, 0, 1, 0, 0};
    const int big = L.bytes + L.cost_bytes;
    const int mode = env_int("GED_B200_COST_STORE", prefer_smem ? 0 : 2);   // 0: shared memory only, 1: tensor memory only, 2: hybrid
    int cols = 32;
    while (cols < max_n1 * S) cols <<= 1;
    const int max_warps_by_regs = 65536 / (32 * ((regs_per_thread + 7) / 8 * 8));
    if (mode != 0 && cols <= 512) {
        const int q = 512 / cols;                              // CTAs per SM that tensor memory can serve
        const int budget = kMaxSmemBytes / q - 1024;           // shared memory per CTA (static + reserved included)
        int k = (budget - kTmemWarps * L.bytes) / big;
        if (budget < kTmemWarps * L.bytes) k = -1;
        // measured on B200 (profiles/): the solve is close to issue-bound once ~8-16 warps per SM are in flight, so extra
        // shared-memory-backed warps only pay when tensor memory can serve a single CTA per SM (n1 * S > 256 columns)
        if ((mode == 1 || q >= 4) && k > 0) k = 0;
        int kmax = max_warps_by_regs / q / 4 * 4 - kTmemWarps;           // registers are allocated in units of 4 warps per CTA
        if (kmax > kMaxCtaWarps - kTmemWarps) kmax = kMaxCtaWarps - kTmemWarps;  // __launch_bounds__ of the solve kernel
        if (k > kmax) k = kmax;
        const int kforce = env_int("GED_B200_SMEM_WARPS", -1);
        if (kforce >= 0) { k = (budget - kTmemWarps * L.bytes) / big; if (k > kmax) k = kmax; if (kforce < k) k = kforce; }
        if (k >= 0) {
            // Global-memory warps INSTEAD of the shared-memory ones where that puts at least a third more warps on the SM:
            // measured (profiles/r02o_ab_gmem_experiment.md) 16 global-hybrid against 12 shared-hybrid warps per SM is +12 % at
            // 80 nodes, 16 against 14 is -1 % (70 nodes), 8 against 7 is -9 % (100 nodes) -- a global-memory warp runs at
            // roughly 0.6 of a shared-memory warp's speed, it pays as an additional warp only.
            int g = 0;
            if (allow_gmem && mode == 2 && q == 2 && kmax > 0 && env_int("GED_B200_GMEM_WARPS", 1)) {   // (q == 1: not measured)
                int kg = (budget - kTmemWarps * L.bytes) / L.bytes;
                if (kg > kmax) kg = kmax;
                if (3 * (kTmemWarps + kg) >= 4 * (kTmemWarps + k)) { g = kg; k = 0; }
            }
            c.tmem_warps = kTmemWarps;
            c.tmem_cols = cols;
            c.smem_warps = k;
            c.gmem_warps = g;
            c.smem_bytes = kTmemWarps * L.bytes + k * big + g * L.bytes;
            return c;
        }
    }
    // shared memory only: one warp per CTA keeps CTA granularity minimal; pack two when > 32 CTAs would fit
    c.smem_warps = (big + 1024) * 32 <= kMaxSmemBytes / 2 ? 2 : 1;
    {
        const int force = env_int("GED_B200_SMEM_CTA_WARPS", 0);      // experiments: shared-memory-only CTAs of this many warps
        if (force > 0 && force <= kTmemWarps && force * big <= kMaxSmemBytes) c.smem_warps = force;
    }
    c.smem_bytes = c.smem_warps * big;
    return c;
}

// CTAs of this shape one SM holds at once.  Tensor-memory kernels: by hand -- shared memory (1 KB reserved per CTA),
// tensor-memory columns, warps, registers (allocated per CTA in units of 4 warps: ncu reports a 7-warp CTA at 136
// registers as alone on its SM) -- because the occupancy API counts a tensor-memory kernel as one CTA per SM.
// Shared-memory-only kernels: the occupancy API.
static int resident_ctas(const CtaShape &c, int regs_per_thread)
{
    const int wpb = c.tmem_warps + c.smem_warps + c.gmem_warps;
    int ctas = 32;
    ctas = min(ctas, kSmemPerSm / (c.smem_bytes + 1024));
    if (c.tmem_cols > 0) ctas = min(ctas, 512 / c.tmem_cols);
    ctas = min(ctas, 64 / wpb);
    const int regs = (regs_per_thread + 7) / 8 * 8;
    if (regs > 0) ctas = min(ctas, 65536 / (regs * 32 * ((wpb + 3) / 4 * 4)));
    return ctas < 1 ? 1 : ctas;
}

template <int S, int MODE>
int resident_ctas_of(const CtaShape &c)
{
    const int wpb = c.tmem_warps + c.smem_warps + c.gmem_warps;
    if (cudaFuncSetAttribute(ged_solve_kernel<S, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, c.smem_bytes) != cudaSuccess) return -1;
    if constexpr (MODE == kModeSmem) {
        int ctas = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, ged_solve_kernel<S, MODE>, wpb * 32, c.smem_bytes) != cudaSuccess) return -1;
        return ctas < 1 ? 1 : ctas;
    } else {
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, ged_solve_kernel<S, MODE>) != cudaSuccess) return -1;
        return resident_ctas(c, fa.numRegs);
    }
}

template <int S, int MODE>
int launch_solve_as(SolveParams prm, const CtaShape &c, cudaStream_t stream)
{
    prm.tmem_warps = c.tmem_warps;
    prm.tmem_cols = c.tmem_cols;
    const int wpb = c.tmem_warps + c.smem_warps + c.gmem_warps;
    cudaError_t e = cudaFuncSetAttribute(ged_solve_kernel<S, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, c.smem_bytes);
    if (e != cudaSuccess) return fail_cuda(e, "ged_solve_batch");
    // Shared-memory carve-out.  Packed (throughput) launches all ask for the maximum: an SM runs CTAs of one L1/shared split
    // at a time, so size classes with different preferences (launched concurrently on side streams) would exclude each
    // other from an SM -- measured +13 % on the AIDS-30-50 mix (two classes), +3.6 % on the AIDS-50+ mix.  Spread
    // (latency-bound) launches leave the choice to the driver: with one warp per SM the L1 serves the K.X gathers, and
    // the 28 KB that the maximum carve-out leaves made a 32-pair AIDS-50+ step 32 % longer.
    // GED_B200_CARVEOUT overrides both (percent; -1 = driver's choice).
    {
        // Global-hybrid launches keep only slices in shared memory; alone they run 2.7 % faster with half the carve-out (the L1
        // then serves their cost-matrix rows), next to the other size classes of a mix 2 % slower -- the uniform split wins
        // (profiles/r02q_ab_gmem.md).  GED_B200_CARVEOUT_G overrides it for them.
        const int carve_packed = env_int("GED_B200_CARVEOUT", prm.spread ? -1 : 100);
        const int carve = MODE == kModeHybridG ? env_int("GED_B200_CARVEOUT_G", carve_packed) : carve_packed;
        e = cudaFuncSetAttribute(ged_solve_kernel<S, MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        if (e != cudaSuccess) return fail_cuda(e, "ged_solve_batch");
    }
    if (env_int("GED_B200_VERBOSE", 0))
        fprintf(stderr, "[ged_b200] solve<S=%d> count=%d max_n=(%d,%d) cta: %d tmem warps (%d cols) + %d smem warps + %d gmem warps, %d B dynamic smem\n",
                S, prm.count, prm.pairs.max_n1, prm.pairs.max_n2, c.tmem_warps, c.tmem_cols, c.smem_warps, c.gmem_warps, c.smem_bytes);
    // Launch shape (ged_solve_desc.launch_hint).  Packed: persistent CTAs (no more than the device holds at once) pulling
    // from the work queue.  Spread (a latency-bound call of a few hundred candidates): one candidate per CTA until every
    // CTA slot is taken, then a second one per CTA, ... -- measured 7-13 % shorter steps at 60-270 candidates of AIDS-50+
    // shape; with more candidates the idle warps of spread CTAs hold tensor memory and shared memory that others need.
    int grid = (prm.count + wpb - 1) / wpb;
    {
        int dev = 0, sms = 0;
        const int per_sm = resident_ctas_of<S, MODE>(c);
        if (per_sm < 0 || cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            return fail(GED_E_CUDA, "ged_solve_batch: device query failed");
        const int resident = sms * per_sm;
        if (prm.spread && prm.count < resident * wpb) {
            prm.work_counter = nullptr;
            grid = prm.count < resident ? prm.count : resident;
        } else {
            prm.spread = 0;
            if (prm.work_counter && grid > resident) grid = resident;
            // both cursors of the two-ended queue share the caller's one int32 (16 bits each; every warp adds once more
            // than it takes)
            prm.two_ended = MODE == kModeHybridG && prm.work_counter && prm.count + grid * wpb < 60000 && env_int("GED_B200_TWO_ENDED", 1);
        }
    }
    ged_solve_kernel<S, MODE><<<grid, wpb * 32, c.smem_bytes, stream>>>(prm);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda(e, "ged_solve_batch");
    return GED_OK;
}

// registers per thread the CTA shape is planned with: the larger of the tensor-memory and the hybrid instantiation
template <int S>
int solve_regs()
{
    cudaFuncAttributes a, b, g;
    if (cudaFuncGetAttributes(&a, ged_solve_kernel<S, kModeTmem>) != cudaSuccess) return -1;
    if (cudaFuncGetAttributes(&b, ged_solve_kernel<S, kModeHybrid>) != cudaSuccess) return -1;
    if (cudaFuncGetAttributes(&g, ged_solve_kernel<S, kModeHybridG>) != cudaSuccess) return -1;
    const int r = a.numRegs > b.numRegs ? a.numRegs : b.numRegs;
    return r > g.numRegs ? r : g.numRegs;
}

template <int S>
int launch_solve(const SolveParams &prm, cudaStream_t stream)
{
    cudaFuncAttributes fa;
    const int regs = solve_regs<S>();
    if (regs < 0) return fail(GED_E_CUDA, "ged_solve_batch: cudaFuncGetAttributes failed");
    // Latency-bound (spread) launches of the S <= 2 classes with more candidates than SMs keep the cost matrix in shared
    // memory: measured on B200 (profiles/r02c_ab_spread_cost_store.md) a 270 / 512-candidate AIDS-30-50 step is 1.41x / 1.56x
    // shorter than with the tensor-memory CTAs (whose four warps host one candidate each in this mode), equal at <= 1 per SM.
    bool prefer_smem = false;
    if (prm.spread && S <= 2) {
        int dev = 0, sms = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess)
            prefer_smem = prm.count > sms;
    }
    // global-memory warps need the third matrix of the scratch rows, and rows long enough for the lanes that read past n2
    const int mn = prm.pairs.max_n1 < prm.pairs.max_n2 ? prm.pairs.max_n1 : prm.pairs.max_n2;
    const bool allow_gmem = prm.workspace && prm.ws_mats >= 3 && mn >= 31 && !prm.spread;
    const CtaShape c = plan_cta(prm.layout, prm.pairs.max_n1, S, regs, prefer_smem, allow_gmem);
    if (c.smem_bytes > kMaxSmemBytes) return fail(GED_E_TOO_LARGE, "pair too large for one warp's shared-memory slice");
    if (c.tmem_warps > 0 && c.gmem_warps > 0) return launch_solve_as<S, kModeHybridG>(prm, c, stream);
    if (c.tmem_warps > 0 && c.smem_warps > 0) return launch_solve_as<S, kModeHybrid>(prm, c, stream);
    return c.tmem_warps > 0 ? launch_solve_as<S, kModeTmem>(prm, c, stream) : launch_solve_as<S, kModeSmem>(prm, c, stream);
}

// ---- team (CTA-per-candidate) launches: their own translation unit (-DGED_TU_TEAM) --------------------------------------
constexpr int kTeamDoesNotFit = -1;          // not an error: the pair's team CTA exceeds shared memory
constexpr int kTeamHelpers = 4;              // warps of a team CTA around a one-slot (<= 32 nodes) pair
int launch_solve_team(int W, const SolveParams &prm, cudaStream_t stream);

#ifdef GED_TU_TEAM
template <int W, int S = W>
int launch_team(SolveParams prm, cudaStream_t stream)
{
    const int smem = team_smem_bytes(prm.layout, W, prm.pairs.max_n2);
    if (smem > kMaxSmemBytes) return kTeamDoesNotFit;          // the caller falls back to the one-warp launch
    cudaError_t e = cudaFuncSetAttribute(ged_solve_team_kernel<W, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail_cuda(e, "ged_solve_batch (team)");
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return fail(GED_E_CUDA, "ged_solve_batch (team): device query failed");
    {   // Shared-memory carve-out.  An SM runs CTAs of ONE L1 / shared split at a time and the driver's own choice follows the
        // single CTA: measured on B200 (profiles/r02d_team.md) a 270-candidate beam step (two size classes launched side by side)
        // takes 108 ms with the driver's choice against 58 ms at 50 % (89 ms with one warp per candidate).  Half of the array is
        // the default -- the K.X gathers of X and Y (global, L2 resident) want the other half as L1: 50 % is 3-8 % faster than
        // the maximum -- unless the CTAs of the whole call (prm.B candidates over all its size-class launches) need more to be
        // resident together.
        const int per_sm_needed = (prm.B + sms - 1) / sms;
        const int dflt = per_sm_needed * (smem + 1024) <= kSmemPerSm / 2 ? 50 : 100;
        const int carve = env_int("GED_B200_TEAM_CARVEOUT", dflt);
        e = cudaFuncSetAttribute(ged_solve_team_kernel<W, S>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        if (e != cudaSuccess) return fail_cuda(e, "ged_solve_batch (team)");
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ged_solve_team_kernel<W, S>, 32 * W, smem) != cudaSuccess || per_sm < 1)
        return fail(GED_E_CUDA, "ged_solve_batch (team): occupancy query failed");
    int grid = prm.count < sms * per_sm ? prm.count : sms * per_sm;      // persistent CTAs beyond one wave
    if (prm.ws_slots > 0 && grid > prm.ws_slots) grid = prm.ws_slots;    // one scratch row per CTA in flight
    if (env_int("GED_B200_VERBOSE", 0))
        fprintf(stderr, "[ged_b200] team solve<W=%d,S=%d> count=%d max_n=(%d,%d) grid=%d (%d CTAs/SM), %d B dynamic smem\n", W, S, prm.count,
                prm.pairs.max_n1, prm.pairs.max_n2, grid, per_sm, smem);
    ged_solve_team_kernel<W, S><<<grid, 32 * W, smem, stream>>>(prm);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda(e, "ged_solve_batch (team)");
    return GED_OK;
}

int launch_solve_team(int W, const SolveParams &prm, cudaStream_t stream)
{
    switch (W) {
    case 1: return launch_team<kTeamHelpers, 1>(prm, stream);       // pairs of up to 32 nodes: one-warp LAP, the rest split
    case 2: return launch_team<2>(prm, stream);
    case 3: return launch_team<3>(prm, stream);
    case 4: return launch_team<4>(prm, stream);
    case 5: return launch_team<5>(prm, stream);
    case 6: return launch_team<6>(prm, stream);
    case 7: return launch_team<7>(prm, stream);
    default: return fail(GED_E_BADARG, "team launch: 1..7 slots");
    }
}
#endif  // GED_TU_TEAM

// ---- per-S entry points (one translation unit each) and their dispatch ---------------------------------------------------
#define GED_CAT_(a, b) a##b
#define GED_CAT(a, b) GED_CAT_(a, b)
#define GED_DECLARE_S(k)                                                  \
    int GED_CAT(resident_solve_s, k)(const SliceLayout &, int);           \
    int GED_CAT(launch_solve_s, k)(const SolveParams &, cudaStream_t);    \
    int GED_CAT(launch_lsape_s, k)(const LsapeParams &, cudaStream_t);    \
    int GED_CAT(launch_kv_s, k)(const KvParams &, cudaStream_t);
GED_DECLARE_S(1) GED_DECLARE_S(2) GED_DECLARE_S(3) GED_DECLARE_S(4) GED_DECLARE_S(5) GED_DECLARE_S(6) GED_DECLARE_S(7)

#ifdef GED_TU_S
int GED_CAT(launch_solve_s, GED_TU_S)(const SolveParams &prm, cudaStream_t stream) { return launch_solve<GED_TU_S>(prm, stream); }
// solves of one launch that fit one SM at once (CTAs per SM by tensor-memory columns / shared memory / registers x warps per CTA)
int GED_CAT(resident_solve_s, GED_TU_S)(const SliceLayout &L, int max_n1)
{
    const int regs = solve_regs<GED_TU_S>();
    if (regs < 0) return -1;
    const CtaShape c = plan_cta(L, max_n1, GED_TU_S, regs, false, true);       // upper bound: with global-memory warps
    const int wpb = c.tmem_warps + c.smem_warps + c.gmem_warps;
    if (c.tmem_warps > 0 && c.gmem_warps > 0) return resident_ctas_of<GED_TU_S, kModeHybridG>(c) * wpb;
    if (c.tmem_warps > 0 && c.smem_warps > 0) return resident_ctas_of<GED_TU_S, kModeHybrid>(c) * wpb;
    return (c.tmem_warps > 0 ? resident_ctas_of<GED_TU_S, kModeTmem>(c) : resident_ctas_of<GED_TU_S, kModeSmem>(c)) * wpb;
}
int GED_CAT(launch_lsape_s, GED_TU_S)(const LsapeParams &prm, cudaStream_t stream)
{
    return launch_uniform(ged_lsape_kernel<GED_TU_S>, prm, prm.B, prm.layout.bytes + prm.layout.cost_bytes, stream, "ged_lsape_batch");
}
int GED_CAT(launch_kv_s, GED_TU_S)(const KvParams &prm, cudaStream_t stream)
{
    return launch_uniform(ged_kv_kernel<GED_TU_S>, prm, prm.pairs.num_pairs, prm.layout.bytes + prm.layout.cost_bytes, stream, "ged_kv_batch");
}
#elif defined(GED_TU_MAIN)
thread_local char g_err[512] = "";

#define GED_DISPATCH_S(S_, NAME, ...)                                                                   \
    switch (S_) {                                                                                       \
    case 1: return GED_CAT(NAME, 1)(__VA_ARGS__);                                                       \
    case 2: return GED_CAT(NAME, 2)(__VA_ARGS__);                                                       \
    case 3: return GED_CAT(NAME, 3)(__VA_ARGS__);                                                       \
    case 4: return GED_CAT(NAME, 4)(__VA_ARGS__);                                                       \
    case 5: return GED_CAT(NAME, 5)(__VA_ARGS__);                                                       \
    case 6: return GED_CAT(NAME, 6)(__VA_ARGS__);                                                       \
    case 7: return GED_CAT(NAME, 7)(__VA_ARGS__);                                                       \
    default: return fail(GED_E_TOO_LARGE, "graphs above 224 nodes are not supported by this build");    \
    }

static int check_pairs(const ged_pairs &p)
{
    if (p.num_pairs < 0 || p.max_n1 < 1 || p.max_n2 < 1) return fail(GED_E_BADARG, "ged_pairs: bad sizes");
    if (p.num_pairs > 0 && (!p.n1 || !p.n2 || !p.adj1 || !p.adj2 || !p.adj1_off || !p.adj2_off))
        return fail(GED_E_BADARG, "ged_pairs: null graph arrays");
    if (p.num_pairs > 0 && !p.node_cost && (!p.lab1 || !p.lab2 || !p.lab1_off || !p.lab2_off))
        return fail(GED_E_BADARG, "ged_pairs: neither labels nor node_cost given");
    if (p.node_cost && !p.node_cost_off) return fail(GED_E_BADARG, "ged_pairs: node_cost without offsets");
    if (p.max_n1 + p.max_n2 > 448) return fail(GED_E_TOO_LARGE, "n1+n2 above 448 is not supported by this build");
    return GED_OK;
}

static int dispatch_solve(int S, const SolveParams &prm, cudaStream_t stream) { GED_DISPATCH_S(S, launch_solve_s, prm, stream) }
static int dispatch_lsape(int S, const LsapeParams &prm, cudaStream_t stream) { GED_DISPATCH_S(S, launch_lsape_s, prm, stream) }
static int dispatch_kv(int S, const KvParams &prm, cudaStream_t stream) { GED_DISPATCH_S(S, launch_kv_s, prm, stream) }
static int dispatch_resident(int S, const SliceLayout &L, int max_n1) { GED_DISPATCH_S(S, resident_solve_s, L, max_n1) }
#endif  // GED_TU_S / GED_TU_MAIN

}  // namespace ged_impl

#ifdef GED_TU_MAIN
using namespace ged_impl;

extern "C" {

int ged_solve_batch(const ged_solve_desc *d, const ged_solve_out *o, void *stream)
{
    if (!d || !o) return fail(GED_E_BADARG, "ged_solve_batch: null descriptor");
    if (int rc = check_pairs(d->pairs)) return rc;
    if (d->B < 0 || d->max_iter < 0) return fail(GED_E_BADARG, "ged_solve_batch: negative B or max_iter");
    if (d->B == 0) return GED_OK;
    if (!d->base_idx && d->B != d->pairs.num_pairs) return fail(GED_E_BADARG, "ged_solve_batch: B != num_pairs without base_idx");
    if ((d->edit_u == nullptr) != (d->edit_v == nullptr)) return fail(GED_E_BADARG, "ged_solve_batch: edit_u/edit_v must come together");
    if (d->score_adj1 && (!d->score_adj1_off || !d->score_idx)) return fail(GED_E_BADARG, "ged_solve_batch: score_adj1 needs score_adj1_off and score_idx");
    if (d->solver_kind != GED_SOLVER_IPFP && d->solver_kind != GED_SOLVER_HUNGARIAN) return fail(GED_E_BADARG, "ged_solve_batch: unknown solver_kind");
    if (!o->ged || !o->rowmatch || !o->colmatch) return fail(GED_E_BADARG, "ged_solve_batch: ged/rowmatch/colmatch outputs are required");
    if (o->match_stride_1 < d->pairs.max_n1 || o->match_stride_2 < d->pairs.max_n2) return fail(GED_E_BADARG, "ged_solve_batch: match strides too small");
    const long long need = (long long)(d->pairs.max_n1 + 1) * (d->pairs.max_n2 + 1);
    if (d->mat_stride < need) return fail(GED_E_BADARG, "ged_solve_batch: mat_stride < (max_n1+1)*(max_n2+1)");
    if (d->solver_kind == GED_SOLVER_IPFP && !d->workspace) return fail(GED_E_BADARG, "ged_solve_batch: workspace required for IPFP");
    if (d->x_init && !d->workspace) return fail(GED_E_BADARG, "ged_solve_batch: workspace required with x_init");
    SolveParams prm;
    prm.pairs = d->pairs;
    prm.B = d->B;
    prm.base_idx = d->base_idx;
    prm.edit_u = d->edit_u;
    prm.edit_v = d->edit_v;
    prm.score_adj1 = d->score_adj1;
    prm.score_ld = d->score_adj1_ld;
    prm.score_adj1_off = d->score_adj1_off;
    prm.score_idx = d->score_idx;
    prm.max_iter = d->max_iter;
    prm.solver_kind = d->solver_kind;
    prm.x_init = d->x_init;
    prm.mat_stride = d->mat_stride;
    prm.workspace = d->workspace;
    prm.ws_locks = d->workspace_locks;
    prm.ws_slots = d->workspace_slots;
    prm.ws_stride = d->workspace_stride ? d->workspace_stride : d->mat_stride;
    prm.ws_mats = d->workspace_mats >= 3 ? 3 : 2;
    prm.two_ended = 0;
    if (prm.ws_slots < 0 || (prm.ws_slots > 0 && !prm.ws_locks)) return fail(GED_E_BADARG, "ged_solve_batch: workspace_slots needs workspace_locks");
    if (prm.ws_stride < need) return fail(GED_E_BADARG, "ged_solve_batch: workspace_stride < (max_n1+1)*(max_n2+1)");
    prm.order = d->order;
    prm.count = d->order ? d->order_count : d->B;
    prm.work_counter = d->work_counter;
    prm.spread = d->launch_hint >= 1 && env_int("GED_B200_SPREAD", 1);
    if (prm.count < 0 || prm.count > d->B) return fail(GED_E_BADARG, "ged_solve_batch: order_count out of range");
    if (prm.count == 0) return GED_OK;
    prm.tmem_warps = 0;
    prm.tmem_cols = 0;
    prm.out = *o;
    prm.layout = make_layout(d->pairs.max_n1, d->pairs.max_n2);
    const int S = slots_for(d->pairs.max_n1, d->pairs.max_n2);
    // launch_hint 2: a latency-bound call -- the warps of a CTA share one candidate (team kernel).
    // (pairs of up to 32 nodes: one-warp LAP with three helper warps for everything else -- 1.15x at <= 300 candidates, a loss at 1024:
    //  up to GED_B200_TEAM_S1_MAX candidates only)
    if (d->launch_hint == 2 && S <= kTeamMaxW && env_int("GED_B200_TEAM", 1) && (S >= 2 || d->B <= env_int("GED_B200_TEAM_S1_MAX", 512))) {
        const int rc = launch_solve_team(S, prm, (cudaStream_t)stream);
        if (rc != kTeamDoesNotFit) return rc;
    }
    return dispatch_solve(S, prm, (cudaStream_t)stream);
}

int ged_lsape_batch(int32_t B, const int32_t *n1, const int32_t *n2, int32_t max_n1, int32_t max_n2,
                    const float *cost, int64_t mat_stride, int32_t *rowmatch, int32_t ms1, int32_t *colmatch, int32_t ms2,
                    float *lb, int32_t *status, void *stream)
{
    if (B < 0 || max_n1 < 1 || max_n2 < 1) return fail(GED_E_BADARG, "ged_lsape_batch: bad sizes");
    if (B == 0) return GED_OK;
    if (!n1 || !n2 || !cost || !rowmatch || !colmatch) return fail(GED_E_BADARG, "ged_lsape_batch: null pointer");
    if (ms1 < max_n1 || ms2 < max_n2 || mat_stride < (int64_t)(max_n1 + 1) * (max_n2 + 1)) return fail(GED_E_BADARG, "ged_lsape_batch: strides too small");
    if (max_n1 + max_n2 > 448) return fail(GED_E_TOO_LARGE, "n1+n2 above 448 is not supported by this build");
    LsapeParams prm{B, n1, n2, cost, (long long)mat_stride, rowmatch, colmatch, ms1, ms2, lb, status, max_n1, max_n2, make_layout(max_n1, max_n2)};
    return dispatch_lsape(slots_for(max_n1, max_n2), prm, (cudaStream_t)stream);
}

int ged_kv_batch(const ged_pairs *pairs, const float *x, float *cost, float *scratch, int64_t mat_stride, void *stream)
{
    if (!pairs || !x || !cost || !scratch) return fail(GED_E_BADARG, "ged_kv_batch: null pointer");
    if (int rc = check_pairs(*pairs)) return rc;
    if (pairs->num_pairs == 0) return GED_OK;
    if (mat_stride < (int64_t)(pairs->max_n1 + 1) * (pairs->max_n2 + 1)) return fail(GED_E_BADARG, "ged_kv_batch: mat_stride too small");
    KvParams prm{*pairs, x, cost, scratch, (long long)mat_stride, make_layout(pairs->max_n1, pairs->max_n2)};
    return dispatch_kv(slots_for(pairs->max_n1, pairs->max_n2), prm, (cudaStream_t)stream);
}

int ged_score_batch(const ged_pairs *pairs, int32_t B, const int32_t *base_idx, const int32_t *rowmatch, int32_t ms1,
                    const int32_t *colmatch, int32_t ms2, float *ged, void *stream)
{
    if (!pairs || !rowmatch || !colmatch || !ged) return fail(GED_E_BADARG, "ged_score_batch: null pointer");
    if (int rc = check_pairs(*pairs)) return rc;
    if (B < 0) return fail(GED_E_BADARG, "ged_score_batch: negative B");
    if (B == 0) return GED_OK;
    if (!base_idx && B != pairs->num_pairs) return fail(GED_E_BADARG, "ged_score_batch: B != num_pairs without base_idx");
    if (ms1 < pairs->max_n1 || ms2 < pairs->max_n2) return fail(GED_E_BADARG, "ged_score_batch: match strides too small");
    ScoreParams prm{*pairs, B, base_idx, rowmatch, colmatch, ms1, ms2, ged, make_layout(pairs->max_n1, pairs->max_n2)};
    return launch_uniform(ged_score_kernel, prm, B, prm.layout.bytes, (cudaStream_t)stream, "ged_score_batch");
}

static int quadform_rows_per_cta(int B, int nk)
{
    long long r = ((long long)nk * B + 295) / 296;      // aim for >= 2 CTAs per SM over the whole batch
    if (r < 8) r = 8;
    if (r > 64) r = 64;
    return (int)r;
}

int64_t ged_quadform_ws_elems(int32_t B, int32_t nk)
{
    if (B <= 0 || nk <= 0) return 0;
    const int rpc = quadform_rows_per_cta(B, nk);
    return (int64_t)B * ((nk + rpc - 1) / rpc);
}

int ged_quadform_dense_batch(int32_t B, int32_t nk, const float *k, int64_t k_stride, const float *x, float *out,
                             float *partial_ws, void *stream)
{
    if (B < 0 || nk <= 0) return fail(GED_E_BADARG, "ged_quadform_dense_batch: bad sizes");
    if (B == 0) return GED_OK;
    if (!k || !x || !out || !partial_ws) return fail(GED_E_BADARG, "ged_quadform_dense_batch: null pointer");
    if ((size_t)nk * sizeof(float) > 200 * 1024) return fail(GED_E_TOO_LARGE, "ged_quadform_dense_batch: x does not fit shared memory");
    const int rpc = quadform_rows_per_cta(B, nk);
    const int ctas = (nk + rpc - 1) / rpc;
    const int smem = nk * (int)sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(ged_quadform_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail_cuda(e, "ged_quadform_dense_batch");
    ged_quadform_partial_kernel<<<dim3(ctas, B), kQfThreads, smem, (cudaStream_t)stream>>>(nk, k, (long long)k_stride, x, partial_ws, rpc, ctas);
    ged_quadform_final_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(partial_ws, ctas, out, B);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda(e, "ged_quadform_dense_batch");
    return GED_OK;
}

int ged_sinkhorn_batch(int32_t B, const int32_t *rows, const int32_t *cols, int32_t ld_r, int32_t ld_c, const float *s,
                       const float *tau, int32_t iters, float *out, void *stream)
{
    if (B < 0 || ld_r < 1 || ld_c < 1 || iters < 0) return fail(GED_E_BADARG, "ged_sinkhorn_batch: bad sizes");
    if (B == 0) return GED_OK;
    if (!rows || !cols || !s || !tau || !out) return fail(GED_E_BADARG, "ged_sinkhorn_batch: null pointer");
    const int pitch = ld_c | 1;
    const size_t smem = (size_t)ld_r * pitch * sizeof(float);
    if (smem > (size_t)kMaxSmemBytes) return fail(GED_E_TOO_LARGE, "ged_sinkhorn_batch: matrix does not fit shared memory");
    cudaError_t e = cudaFuncSetAttribute(ged_sinkhorn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail_cuda(e, "ged_sinkhorn_batch");
    ged_sinkhorn_kernel<<<B, kSinkhornThreads, smem, (cudaStream_t)stream>>>(rows, cols, ld_r, ld_c, s, tau, iters, out, pitch);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda(e, "ged_sinkhorn_batch");
    return GED_OK;
}

void ged_limits(int32_t *max_nodes_per_graph, int32_t *max_matrix_elems)
{
    if (max_nodes_per_graph) *max_nodes_per_graph = 32 * kMaxS;
    if (max_matrix_elems) {
        int n = 1;                          // largest square pair whose shared-memory-backed slice fits
        for (;;) {
            const SliceLayout L = make_layout(n + 1, n + 1);
            if (L.bytes + L.cost_bytes > kMaxSmemBytes) break;
            ++n;
        }
        *max_matrix_elems = (n + 1) * (n + 1);
    }
}

int32_t ged_solve_max_resident(int32_t max_n1, int32_t max_n2)
{
    if (max_n1 < 1 || max_n2 < 1 || max_n1 + max_n2 > 448) return 0;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    const int per_sm = dispatch_resident(slots_for(max_n1, max_n2), make_layout(max_n1, max_n2), max_n1);
    return per_sm > 0 ? per_sm * sms : 0;
}

int64_t ged_solve_smem_bytes(int32_t max_n1, int32_t max_n2)
{
    const SliceLayout L = make_layout(max_n1, max_n2);
    return L.bytes + L.cost_bytes;
}

int ged_abi_version(void) { return GED_ABI_VERSION; }
const char *ged_version(void) { return "ged_b200 0.7.0 (sm_100a; ABI 9; canonical order v2; tensor-memory cost store; persistent work-queue launches; CTA-per-candidate team launches)"; }
const char *ged_last_error(void) { return g_err; }

}  // extern "C"
#endif  // GED_TU_MAIN
