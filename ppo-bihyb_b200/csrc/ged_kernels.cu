// ged_kernels.cu -- hand-written sm_100a kernels + C-ABI for the GED lower-level solver path of PPO-BiHyb.
//
// Path replaced (reference utils/ged_env.py): GEDenv.step :110-124 -> solve_feasible_ged :140-158 ->
// [construct_k :160-228, never materialised] -> ipfp_ged :256-289 -> hungarian_ged :303-326 ->
// hungarian_lap :329-351 -> scipy linear_sum_assignment :407 -> comp_ged :245-253.
//
// Execution model: ONE WARP PER CANDIDATE SOLVE.  The sequential core of the path is the LAP projection
// (a shortest-augmenting-path search whose scan order and tie-breaks must equal SciPy's), so each candidate is
// owned by a single warp that keeps the LAP's per-column state (dual v, shortest-path cost, position in SciPy's
// `remaining` vector, predecessor) in REGISTERS, one column per (lane, slot), the cost matrix and row duals in its
// private slice of shared memory, and the fractional IPFP iterate X in global memory (it is only touched between
// LAPs and stays L2 resident).  K is never built: K vec(X) is evaluated from the two bit-packed adjacency factors.
// No __syncthreads anywhere -- warps of a CTA are independent, so a CTA is just a packing unit for occupancy.
//
// All floating-point arithmetic follows the canonical order documented in oracle/ged_oracle.c (no FMA
// contraction: explicit *_rn intrinsics, and the file is compiled with --fmad=false).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>

#include "../../include/ged_b200.h"

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kMaxS = 7;                       // slots per column group: graphs up to 32*7 = 224 nodes
constexpr int kMaxSmemBytes = 227 * 1024;

thread_local char g_err[512] = "";

int fail(int code, const char *msg)
{
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}
int fail_cuda(cudaError_t e, const char *where)
{
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return GED_E_CUDA;
}

// ---------------------------------------------------------------------------------------------------------------
// Shared-memory slice of one warp (host and device agree through this one function).
// ---------------------------------------------------------------------------------------------------------------
struct SliceLayout {
    int off_u, off_cost, off_bits1, off_bits2, off_rs, off_rsP, off_cs, off_csQ, off_lab1, off_lab2;
    int off_row4col, off_col4row, off_path, off_rowmatch, off_colmatch, off_bestrow, off_bestcol;
    int bytes;
};

__host__ __device__ inline int align_up(int x, int a) { return (x + a - 1) / a * a; }

__host__ __device__ inline SliceLayout make_layout(int max_n1, int max_n2)
{
    const int Rm = max_n1 + 1, Cm = max_n2 + 1, Nm = max_n1 + max_n2;
    const int W1 = (max_n1 + 31) / 32 > 0 ? (max_n1 + 31) / 32 : 1;
    const int W2 = (max_n2 + 31) / 32 > 0 ? (max_n2 + 31) / 32 : 1;
    SliceLayout L;
    int o = 0;
    L.off_u = o;        o += 8 * Nm;
    L.off_cost = o;     o += 4 * (Rm * Cm + 32 * ((Cm + 31) / 32));   // + one padded row: lanes read past C
    L.off_bits1 = o;    o += 4 * Rm * W1;
    L.off_bits2 = o;    o += 4 * Cm * W2;
    L.off_rs = o;       o += 4 * Rm;
    L.off_rsP = o;      o += 4 * Rm;
    L.off_cs = o;       o += 4 * Cm;
    L.off_csQ = o;      o += 4 * Cm;
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
    return L;
}

struct Slice {
    double *u;
    float *cost;
    uint32_t *bits1, *bits2;
    float *rs, *rsP, *cs, *csQ;
    int32_t *lab1, *lab2;
    int16_t *row4col, *col4row, *path, *rowmatch, *colmatch, *bestrow, *bestcol;
};

__device__ inline Slice carve(unsigned char *base, const SliceLayout &L)
{
    Slice s;
    s.u = reinterpret_cast<double *>(base + L.off_u);
    s.cost = reinterpret_cast<float *>(base + L.off_cost);
    s.bits1 = reinterpret_cast<uint32_t *>(base + L.off_bits1);
    s.bits2 = reinterpret_cast<uint32_t *>(base + L.off_bits2);
    s.rs = reinterpret_cast<float *>(base + L.off_rs);
    s.rsP = reinterpret_cast<float *>(base + L.off_rsP);
    s.cs = reinterpret_cast<float *>(base + L.off_cs);
    s.csQ = reinterpret_cast<float *>(base + L.off_csQ);
    s.lab1 = reinterpret_cast<int32_t *>(base + L.off_lab1);
    s.lab2 = reinterpret_cast<int32_t *>(base + L.off_lab2);
    s.row4col = reinterpret_cast<int16_t *>(base + L.off_row4col);
    s.col4row = reinterpret_cast<int16_t *>(base + L.off_col4row);
    s.path = reinterpret_cast<int16_t *>(base + L.off_path);
    s.rowmatch = reinterpret_cast<int16_t *>(base + L.off_rowmatch);
    s.colmatch = reinterpret_cast<int16_t *>(base + L.off_colmatch);
    s.bestrow = reinterpret_cast<int16_t *>(base + L.off_bestrow);
    s.bestcol = reinterpret_cast<int16_t *>(base + L.off_bestcol);
    return s;
}

// Per-pair view used by the math routines.
struct Pair {
    int n1, n2, R, C, N, W1, W2;
    int maxdeg;                 // max node degree over both graphs (after the candidate edit)
    const float *node_cost;     // global, R x C, or nullptr -> label metric
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
__device__ inline bool load_bits(const uint8_t *adj, int n, int W, uint32_t *bits, int lane)
{
    bool bad = false;
    for (int i = 0; i < n; ++i)
        for (int w = 0; w < W; ++w) {
            const int j = w * 32 + lane;
            unsigned val = 0, tval = 0;
            if (j < n && j != i) { val = adj[(int64_t)i * n + j]; tval = adj[(int64_t)j * n + i]; }
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

// ---------------------------------------------------------------------------------------------------------------
// LSAPE through the expanded (n1+n2)^2 LAP, SciPy order (hungarian_lap ged_env.py:329-351; scipy
// rectangular_lsap augmenting_path + solve).  Columns of the expanded matrix live in registers:
//   real column a      -> lane a & 31, slot a >> 5            (slots 0 .. S-1)
//   deletion column k  -> lane k & 31, slot S + (k >> 5)      (expanded column index n2 + k)
// `pos` is the column's index in SciPy's `remaining` vector (initially N-1-j, swap-removed on selection); the
// selection rule "last minimum whose column is unassigned, else first minimum" becomes a lexicographic arg-min
// over (spc, key) with key = unassigned ? -1-pos : pos, done with three integer warp reductions.
// ---------------------------------------------------------------------------------------------------------------
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

// if (live & MASK) && r < spc:  spc = r, pth = i, improved = 1
template <unsigned MASK>
__device__ __forceinline__ void lap_relax(double &spc, int &pth, int &improved, double r, int i, unsigned live)
{
    asm("{\n\t.reg .pred q, p;\n\t.reg .b32 t;\n\t"
        "and.b32 t, %5, %6;\n\t"
        "setp.ne.u32 q, t, 0;\n\t"
        "setp.lt.and.f64 p, %3, %0, q;\n\t"
        "selp.f64 %0, %3, %0, p;\n\t"
        "selp.b32 %1, %4, %1, p;\n\t"
        "@p mov.b32 %2, 1;\n\t}"
        : "+d"(spc), "+r"(pth), "+r"(improved) : "d"(r), "r"(i), "r"(live), "n"(MASK));
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

template <int S, int I = 0>
struct LapSlots {
    static __device__ __forceinline__ void relax_range(double (&spc)[2 * S], int (&pth)[2 * S], int &improved,
                                                       const double (&v)[2 * S], double base, int i, unsigned live, int lo, int hi)
    {
        if constexpr (I < 2 * S) {
            if (I >= lo && I < hi) lap_relax<(1u << I)>(spc[I], pth[I], improved, __dsub_rn(base, v[I]), i, live);
            LapSlots<S, I + 1>::relax_range(spc, pth, improved, v, base, i, live, lo, hi);
        }
    }
    static __device__ __forceinline__ void min_value(double &lm, const double (&spc)[2 * S], unsigned live)
    {
        if constexpr (I < 2 * S) {
            lap_min_value<(1u << I)>(lm, spc[I], live);
            LapSlots<S, I + 1>::min_value(lm, spc, live);
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

// LSAPE through the expanded (n1+n2)^2 LAP in SciPy's order (hungarian_lap ged_env.py:329-351; scipy rectangular_lsap).
// Columns live in registers, one per (lane, slot): real column a -> lane a & 31, slot a >> 5; deletion column n2+k ->
// lane k & 31, slot S + (k >> 5).  Tie-break key of a column:  kk << 17 | slot << 14 | lane << 9 | (row4col + 1), with
// kk = pos ^ 511 for an unassigned column (SciPy: the LAST unassigned minimum wins) and pos ^ 512 = 512 + pos for an
// assigned one (else the FIRST minimum wins); pos = index in SciPy's `remaining` vector.  The smallest key wins.
template <int S>
__device__ int lsape_warp(const float *cost, int n1, int n2, const Slice &sm, unsigned long long *steps_out)
{
    const int lane = threadIdx.x & 31;
    const int N = n1 + n2, C = n2 + 1;
    double *u = sm.u;

    {   // scipy rejects NaN / -inf (RECTANGULAR_LSAP_INVALID); the dummy-dummy entry is not part of the LAP
        bool bad = false;
        for (int e = lane; e < (n1 + 1) * C - 1; e += 32) {
            const float c = cost[e];
            bad |= (c != c) || (c == -INFINITY);
        }
        if (__any_sync(kFull, bad)) return GED_ST_INVALID_COST;
    }

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

    unsigned long long steps = 0;
    for (int cur = 0; cur < N; ++cur) {
        double spc[2 * S];
        int pth[2 * S], posk[2 * S], kx[2 * S];
#pragma unroll
        for (int s = 0; s < 2 * S; ++s) {
            const int j = (s < S) ? (lane + 32 * s) : (n2 + lane + 32 * (s - S));
            spc[s] = INFINITY;
            pth[s] = -1;
            posk[s] = (N - 1 - j) << 17;
            int r4 = -1;
            if ((exist >> s) & 1u) r4 = sm.row4col[j];
            kx[s] = (((r4 >= 0) ? 512 : 511) << 17) | (s << 14) | (lane << 9) | (r4 + 1);
            asm volatile("" : "+r"(kx[s]));          // keep it in a register (no re-derivation inside the scan loop)
        }
        unsigned live = exist;               // columns still in `remaining`
        double minVal = 0.0, rzmin = INFINITY;
        int i = cur, lastk = N << 17;
        unsigned mk;
        int nsteps = 0;
        bool infeasible = false;

        // operands of the row about to be scanned; loaded one step ahead (right after the winner is known) so that the
        // shared-memory latency overlaps the bookkeeping of the previous step
        double ui = u[i];
        float cs[S], cx;                     // cs: the row's real entries (real rows only); cx: its deletion / insertion entry
        {
            const bool real = i < n1;
            cx = cost[real ? (i * C + n2) : (n1 * C + (i - n1))];
#pragma unroll
            for (int s = 0; s < S; ++s) cs[s] = real ? cost[i * C + lane + 32 * s] : 0.0f;
        }

        for (;;) {
            ++nsteps;
            int improved = 0;
            if (i < n1) {                    // real row: n2 substitution entries + its own deletion entry
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    const double r = __dsub_rn(__dsub_rn(__dadd_rn(minVal, (double)cs[s]), ui), v[s]);
                    if (s == 0) lap_relax<1u>(spc[0], pth[0], improved, r, i, live);
                    if (s == 1) lap_relax<2u>(spc[s], pth[s], improved, r, i, live);
                    if (s == 2) lap_relax<4u>(spc[s], pth[s], improved, r, i, live);
                    if (s == 3) lap_relax<8u>(spc[s], pth[s], improved, r, i, live);
                    if (s == 4) lap_relax<16u>(spc[s], pth[s], improved, r, i, live);
                    if (s == 5) lap_relax<32u>(spc[s], pth[s], improved, r, i, live);
                    if (s == 6) lap_relax<64u>(spc[s], pth[s], improved, r, i, live);
                }
                const double rd = __dsub_rn(__dadd_rn(minVal, (double)cx), ui);
                const unsigned hl = (lane == (i & 31)) ? (live & ((1u << S) << (i >> 5))) : 0u;
                LapSlots<S>::relax_range(spc, pth, improved, v, rd, i, hl, S, 2 * S);
            } else {                         // insertion row of column a0: its own entry + the zero block
                const int a0 = i - n1;
                const double ra = __dsub_rn(__dadd_rn(minVal, (double)cx), ui);
                const unsigned hl = (lane == (a0 & 31)) ? (live & (1u << (a0 >> 5))) : 0u;
                LapSlots<S>::relax_range(spc, pth, improved, v, ra, i, hl, 0, S);
                // zero block: r = rz - v[j].  Rounding is monotone in rz, and after a scan with value rz0 every remaining
                // deletion column has spc <= fl(rz0 - v[j]); a later insertion row with rz >= rz0 cannot lower any of them.
                const double rz = __dsub_rn(__dadd_rn(minVal, 0.0), ui);
                if (rz < rzmin) {
                    rzmin = rz;
                    LapSlots<S>::relax_range(spc, pth, improved, v, rz, i, live, S, 2 * S);
                }
            }

            // Tie-break among the columns that still attain minVal (pass 2).  If some shortest-path cost was lowered
            // (a lane then bids key 0, below every real key) or no remaining column attains minVal (all bids are ~0u),
            // the minimum has to be recomputed first (pass 1).
            unsigned lk = 0xffffffffu;
            LapSlots<S>::min_key(lk, spc, minVal, live, posk, kx);
            lk = improved ? 0u : lk;
            mk = __reduce_min_sync(kFull, lk);
            if (mk - 1u >= 0xfffffffeu) {
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
                mk = inf ? 0u : mk;
            }

            const int r4star = (int)(mk & 511u) - 1;
            {   // operands of the next row
                const int in = max(r4star, 0);
                const bool real = in < n1;
                ui = u[in];
                cx = cost[real ? (in * C + n2) : (n1 * C + (in - n1))];
#pragma unroll
                for (int s = 0; s < S; ++s) if (real) cs[s] = cost[in * C + lane + 32 * s];
            }
            // the winning column leaves `remaining`; the column at its end takes the freed position
            lastk -= 1 << 17;
            const int posstark = (int)((mk & 0xfffe0000u) ^ ((mk & (512u << 17)) ? (512u << 17) : (511u << 17)));
            if (lk == mk) live &= ~(1u << ((mk >> 14) & 7u));
#pragma unroll
            for (int s = 0; s < 2 * S; ++s) posk[s] = (posk[s] == lastk) ? posstark : posk[s];
            if (r4star < 0) break;
            i = r4star;
        }
        if (infeasible) return GED_ST_INFEASIBLE;
        steps += nsteps;
        const int wl = (int)((mk >> 9) & 31u), ws = (int)((mk >> 14) & 7u);
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
                const int r4 = (kx[s] & 511) - 1;
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

// ---------------------------------------------------------------------------------------------------------------
// cost = K vec(X) from the factors (replaces torch.mm(k, v), ged_env.py:269).  X in global memory, cost in smem.
//   K vec(X) = A1 X M2^T + M1 X A2^T - 2 A1 X A2^T + Cn o X      (identity checked in the tests)
// T = number of 32-wide column tiles of a row.
// ---------------------------------------------------------------------------------------------------------------
template <int S>
__device__ void kv_apply_generic(const Pair &P, const Slice &sm, const float *X, float *cost)
{
    constexpr int T = S + 1;
    const int lane = threadIdx.x & 31;
    const int R = P.R, C = P.C, n1 = P.n1, n2 = P.n2;

    float csacc[T];
#pragma unroll
    for (int t = 0; t < T; ++t) csacc[t] = 0.0f;
    for (int i = 0; i < R; ++i) {
        float part = 0.0f;
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const int a = lane + 32 * t;
            if (a < C) {
                const float x = X[i * C + a];
                csacc[t] = __fadd_rn(csacc[t], x);
                part = __fadd_rn(part, x);
            }
        }
        part = warp_butterfly(part);
        if (lane == 0) sm.rs[i] = part;
    }
#pragma unroll
    for (int t = 0; t < T; ++t) { const int a = lane + 32 * t; if (a < C) sm.cs[a] = csacc[t]; }
    __syncwarp();
    for (int i = lane; i < R; i += 32) {
        float s = 0.0f;
        for (int w = 0; w < P.W1; ++w) {
            uint32_t word = sm.bits1[i * P.W1 + w];
            while (word) { const int j = w * 32 + __ffs(word) - 1; word &= word - 1; s = __fadd_rn(s, sm.rs[j]); }
        }
        sm.rsP[i] = s;
    }
    for (int a = lane; a < C; a += 32) {
        float s = 0.0f;
        for (int w = 0; w < P.W2; ++w) {
            uint32_t word = sm.bits2[a * P.W2 + w];
            while (word) { const int b = w * 32 + __ffs(word) - 1; word &= word - 1; s = __fadd_rn(s, sm.cs[b]); }
        }
        sm.csQ[a] = s;
    }
    __syncwarp();

    for (int i = 0; i < R; ++i) {
        const uint32_t *row1 = sm.bits1 + i * P.W1;
        const float rsPi = sm.rsP[i];
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const int a = lane + 32 * t;
            if (a < C) {
                const uint32_t *row2 = sm.bits2 + a * P.W2;
                float Pv = 0.0f, Qv = 0.0f, T3 = 0.0f;
                for (int w2 = 0; w2 < P.W2; ++w2) {
                    uint32_t word2 = row2[w2];
                    while (word2) { const int b = w2 * 32 + __ffs(word2) - 1; word2 &= word2 - 1; Qv = __fadd_rn(Qv, X[i * C + b]); }
                }
                for (int w1 = 0; w1 < P.W1; ++w1) {
                    uint32_t word1 = row1[w1];
                    while (word1) {
                        const int j = w1 * 32 + __ffs(word1) - 1;
                        word1 &= word1 - 1;
                        const float *xr = X + j * C;
                        Pv = __fadd_rn(Pv, xr[a]);
                        float in = 0.0f;
                        for (int w2 = 0; w2 < P.W2; ++w2) {
                            uint32_t word2 = row2[w2];
                            while (word2) { const int b = w2 * 32 + __ffs(word2) - 1; word2 &= word2 - 1; in = __fadd_rn(in, xr[b]); }
                        }
                        T3 = __fadd_rn(T3, in);
                    }
                }
                const float t1 = (a < n2) ? __fsub_rn(rsPi, Pv) : rsPi;
                const float t2 = (i < n1) ? __fsub_rn(sm.csQ[a], Qv) : sm.csQ[a];
                float s = __fadd_rn(t1, t2);
                s = __fsub_rn(s, __fadd_rn(T3, T3));
                const float d = __fmul_rn(node_cost_at(P, sm, i, a), X[i * C + a]);
                cost[i * C + a] = __fadd_rn(s, d);
            }
        }
    }
    __syncwarp();
}

// Fast path of K vec(X) for graphs whose maximum degree is <= kMaxDeg (molecules: <= 4, +1 for a candidate edit).
// Same arithmetic, same summation order as kv_apply_generic (and oracle/ged_oracle.c), restructured around
//     Y = X A2^T   (Y[j,a] = sum_{b in N2(a)} X[j,b], ascending b)
// so that every inner term is a ROW gather:  P[i,a] = sum_{j in N1(i)} X[j,a],  T3[i,a] = sum_{j in N1(i)} Y[j,a],
// Q[i,a] = Y[i,a].  X is staged once in the (not yet written) cost region of shared memory for the random column
// gathers of Y; Y goes to global scratch; the final pass reads X and Y rows with coalesced, mutually independent loads
// (one L2 round trip per row of the result) and writes cost into shared memory.
constexpr int kMaxDeg = 6;

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

template <int S>
__device__ void kv_apply(const Pair &P, const Slice &sm, const float *X, float *Y, float *cost)
{
    if (P.maxdeg > kMaxDeg) { kv_apply_generic<S>(P, sm, X, cost); return; }
    constexpr int T = S + 1;
    const int lane = threadIdx.x & 31;
    const int R = P.R, C = P.C, n1 = P.n1, n2 = P.n2, RC = R * C;

    for (int e = lane; e < RC; e += 32) cost[e] = X[e];          // stage X
    __syncwarp();

    float csacc[T];
#pragma unroll
    for (int t = 0; t < T; ++t) csacc[t] = 0.0f;
    for (int i = 0; i < R; ++i) {
        float part = 0.0f;
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const int a = lane + 32 * t;
            if (a < C) {
                const float x = cost[i * C + a];
                csacc[t] = __fadd_rn(csacc[t], x);
                part = __fadd_rn(part, x);
            }
        }
        part = warp_butterfly(part);
        if (lane == 0) sm.rs[i] = part;
    }
#pragma unroll
    for (int t = 0; t < T; ++t) { const int a = lane + 32 * t; if (a < C) sm.cs[a] = csacc[t]; }
    __syncwarp();
    for (int i = lane; i < R; i += 32) {
        int nb[kMaxDeg];
        decode_neighbours(sm.bits1 + i * P.W1, P.W1, nb);
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < kMaxDeg; ++k) if (nb[k] >= 0) s = __fadd_rn(s, sm.rs[nb[k]]);
        sm.rsP[i] = s;
    }
#pragma unroll
    for (int t = 0; t < T; ++t) {
        if (32 * t >= C) break;
        const int a = lane + 32 * t;
        if (a < C) {
            int nb[kMaxDeg];
            decode_neighbours(sm.bits2 + a * P.W2, P.W2, nb);
            float s = 0.0f;
#pragma unroll
            for (int k = 0; k < kMaxDeg; ++k) if (nb[k] >= 0) s = __fadd_rn(s, sm.cs[nb[k]]);
            sm.csQ[a] = s;
            for (int j = 0; j < R; ++j) {                        // Y[j, a]
                const float *xr = cost + j * C;
                float in = 0.0f;
#pragma unroll
                for (int k = 0; k < kMaxDeg; ++k) if (nb[k] >= 0) in = __fadd_rn(in, xr[nb[k]]);
                Y[j * C + a] = in;
            }
        }
    }
    __syncwarp();

    for (int i = 0; i < R; ++i) {
        int nb[kMaxDeg];
        decode_neighbours(sm.bits1 + i * P.W1, P.W1, nb);        // warp-uniform
        const float rsPi = sm.rsP[i];
#pragma unroll
        for (int t = 0; t < T; ++t) {
            if (32 * t >= C) break;
            const int a = lane + 32 * t;
            if (a < C) {
                float xv[kMaxDeg], yv[kMaxDeg];
#pragma unroll
                for (int k = 0; k < kMaxDeg; ++k) {              // all loads first: they are independent
                    xv[k] = 0.0f;
                    yv[k] = 0.0f;
                    if (nb[k] >= 0) { xv[k] = X[nb[k] * C + a]; yv[k] = Y[nb[k] * C + a]; }
                }
                const float xia = X[i * C + a], Qv = Y[i * C + a];
                float Pv = 0.0f, T3 = 0.0f;
#pragma unroll
                for (int k = 0; k < kMaxDeg; ++k)
                    if (nb[k] >= 0) { Pv = __fadd_rn(Pv, xv[k]); T3 = __fadd_rn(T3, yv[k]); }
                const float t1 = (a < n2) ? __fsub_rn(rsPi, Pv) : rsPi;
                const float t2 = (i < n1) ? __fsub_rn(sm.csQ[a], Qv) : sm.csQ[a];
                float sres = __fadd_rn(t1, t2);
                sres = __fsub_rn(sres, __fadd_rn(T3, T3));
                const float d = __fmul_rn(node_cost_at(P, sm, i, a), xia);
                cost[i * C + a] = __fadd_rn(sres, d);
            }
        }
    }
    __syncwarp();
}

// <A, B> over the flat R*C index, fp64: lane-strided partials + butterfly.
__device__ inline double dot_flat(const float *a_smem, const float *b_glob, int n)
{
    const int lane = threadIdx.x & 31;
    double acc = 0.0;
    for (int e = lane; e < n; e += 32) acc = __dadd_rn(acc, __dmul_rn((double)a_smem[e], (double)b_glob[e]));
    return warp_butterfly(acc);
}

// Sum of an R x C matrix (shared or global, or the implicit node-cost matrix when M == nullptr) over the matched
// entries: t < n1 -> (t, rowmatch[t]);  t >= n1 -> (n1, t - n1) if that column is inserted.
__device__ inline double matched_sum(const Pair &P, const Slice &sm, const float *M, const int16_t *rowmatch, const int16_t *colmatch)
{
    const int lane = threadIdx.x & 31;
    double acc = 0.0;
    for (int t = lane; t < P.N; t += 32) {
        double term = 0.0;
        if (t < P.n1) {
            const int a = rowmatch[t];
            term = (double)(M ? M[t * P.C + a] : node_cost_at(P, sm, t, a));
        } else {
            const int a = t - P.n1;
            if (colmatch[a] == P.n1) term = (double)(M ? M[P.n1 * P.C + a] : node_cost_at(P, sm, P.n1, a));
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
    const double nodes = matched_sum(P, sm, nullptr, rowmatch, colmatch);
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

// heuristic_prediction_hun's node_cost_mat (ged_env.py:310-315) in closed form from node degrees.
__device__ inline void init_cost_closed_form(const Pair &P, const Slice &sm, float *cost)
{
    const int lane = threadIdx.x & 31;
    // degrees into rs / cs (as floats; small integers)
    for (int i = lane; i < P.R; i += 32) {
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
    for (int e = lane; e < P.R * P.C; e += 32) {
        const int i = e / P.C, a = e - i * P.C;
        const int d1 = (int)sm.rs[i], d2 = (int)sm.cs[a];
        int val;
        if (i < P.n1 && a < P.n2) { const int d = abs(d1 - d2) - 1; val = d > 0 ? d : 0; }
        else if (i < P.n1) val = d1;
        else if (a < P.n2) val = d2;
        else val = 0;
        cost[e] = (float)val;
    }
    __syncwarp();
}

__device__ __forceinline__ float dense_b(const Pair &P, const int16_t *rowmatch, const int16_t *colmatch, int i, int a)
{
    if (i < P.n1) return (rowmatch[i] == a) ? 1.0f : 0.0f;
    return (a < P.n2 && colmatch[a] == P.n1) ? 1.0f : 0.0f;
}

// ---------------------------------------------------------------------------------------------------------------
// Kernel parameters (device pointers copied out of the C-ABI descriptors)
// ---------------------------------------------------------------------------------------------------------------
struct SolveParams {
    ged_pairs pairs;
    int B;
    const int32_t *base_idx, *edit_u, *edit_v;
    const uint8_t *score_adj1;
    const int64_t *score_adj1_off;
    const int32_t *score_idx;
    int max_iter, solver_kind;
    const float *x_init;
    long long mat_stride;
    float *workspace;
    const int32_t *order;
    int count;
    ged_solve_out out;
    SliceLayout layout;
};

template <int S>
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
    const bool bad1 = load_bits(pp.adj1 + pp.adj1_off[p], P.n1, P.W1, sm.bits1, lane);
    const bool bad2 = load_bits(pp.adj2 + pp.adj2_off[p], P.n2, P.W2, sm.bits2, lane);
    if (bad1 || bad2) status |= GED_ST_BAD_GRAPH;
    if (!P.node_cost) {
        const int32_t *l1 = pp.lab1 + pp.lab1_off[p], *l2 = pp.lab2 + pp.lab2_off[p];
        for (int i = lane; i < P.n1; i += 32) sm.lab1[i] = l1[i];
        for (int a = lane; a < P.n2; a += 32) sm.lab2[a] = l2[a];
    }
    __syncwarp();
}

template <int S>
__global__ void ged_solve_kernel(SolveParams prm)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = blockIdx.x * (blockDim.x >> 5) + warp;
    if (slot >= prm.count) return;
    const int b = prm.order ? prm.order[slot] : slot;
    const Slice sm = carve(smem_raw + (size_t)warp * prm.layout.bytes, prm.layout);
    const ged_solve_out &out = prm.out;

    const int p = prm.base_idx ? prm.base_idx[b] : b;
    Pair P;
    int status = 0;
    load_pair<S>(prm.pairs, p, sm, P, status);
    const int n1 = P.n1, n2 = P.n2, RC = P.R * P.C;

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
    }
    const int nnz1 = count_bits(sm.bits1, P.R * P.W1), nnz2 = count_bits(sm.bits2, P.C * P.W2);
    P.maxdeg = max_degree(sm.bits1, P.R, P.W1, sm.bits2, P.C, P.W2);

    float *X = prm.workspace + (long long)b * 2 * prm.mat_stride;        // fractional iterate
    float *Y = X + prm.mat_stride;                                       // X A2^T scratch of kv_apply
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
            for (int e = lane; e < RC; e += 32) X[e] = x0[e];
            for (int i = lane; i < n1; i += 32) sm.bestrow[i] = (int16_t)n2;
            for (int a = lane; a < n2; a += 32) sm.bestcol[a] = (int16_t)n1;
            __syncwarp();
        } else {
            // hungarian_ged (ged_env.py:303-326): v0 = LSAPE(node_cost_mat)
            init_cost_closed_form(P, sm, sm.cost);
            if (out.tr_init_cost)
                for (int e = lane; e < RC; e += 32) out.tr_init_cost[(long long)b * prm.mat_stride + e] = sm.cost[e];
            status |= lsape_warp<S>(sm.cost, n1, n2, sm, &lap_steps);
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
#ifdef GED_TIMING
            const long long tk0 = clock64();
#endif
            kv_apply<S>(P, sm, X, Y, sm.cost);                           // ged_env.py:269
#ifdef GED_TIMING
            t_kv += clock64() - tk0;
#endif
            const double s_vv = dot_flat(sm.cost, X, RC);                // = last_v^T K last_v, ged_env.py:283
            if (out.tr_cost) {
                float *dst = out.tr_cost + ((long long)b * prm.max_iter + it) * prm.mat_stride;
                for (int e = lane; e < RC; e += 32) dst[e] = sm.cost[e];
            }
#ifdef GED_TIMING
            const long long tl0 = clock64();
#endif
            status |= lsape_warp<S>(sm.cost, n1, n2, sm, &lap_steps);   // ged_env.py:270
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
            const double s_vb = matched_sum(P, sm, sm.cost, sm.rowmatch, sm.colmatch);
            const double alpha = __dsub_rn(s_vb, s_vv);                  // :275 (K symmetric: v^T K = (K v)^T)
            const double beta = __dadd_rn(__dsub_rn(ub, __dadd_rn(s_vb, s_vb)), s_vv);   // :277
            const double t0 = __ddiv_rn(-alpha, beta);                   // :278
            const float t0f = (float)t0;
            const bool jump = (beta <= 0.0) || (t0f >= 1.0f);            // :279
            float *trx = out.tr_x_after ? out.tr_x_after + ((long long)b * prm.max_iter + it) * prm.mat_stride : nullptr;
            for (int i = 0; i < P.R; ++i)
                for (int a = lane; a < P.C; a += 32) {
                    const int e = i * P.C + a;
                    const float bv = dense_b(P, sm.rowmatch, sm.colmatch, i, a);
                    float xn;
                    if (jump) xn = bv;                                   // :280
                    else { const float x = X[e]; xn = __fadd_rn(x, __fmul_rn(t0f, __fsub_rn(bv, x))); }   // :282
                    X[e] = xn;
                    if (trx) trx[e] = xn;
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
            const bool bad = load_bits(prm.score_adj1 + prm.score_adj1_off[prm.score_idx[b]], n1, P.W1, sm.bits1, lane);
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
    if (lane == 0) printf("solve %d: total %lld cycles, lap %lld (%.1f%%), kv %lld (%.1f%%), lap steps %llu -> %.1f cycles/step\n", b,
                          clock64() - t_begin, t_lap, 100.0 * t_lap / (clock64() - t_begin), t_kv, 100.0 * t_kv / (clock64() - t_begin),
                          lap_steps, (double)t_lap / (double)lap_steps);
#endif
    if (lane == 0) {
        out.ged[b] = ged_base;
        if (out.ged_edited) out.ged_edited[b] = ged_edit;
        if (out.iters) out.iters[b] = iters;
        if (out.status) out.status[b] = status;
        if (out.tr_lap_steps) out.tr_lap_steps[b] = lap_steps;
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

// ---- stand-alone stages ---------------------------------------------------------------------------------------
struct LsapeParams {
    int B;
    const int32_t *n1, *n2;
    const float *cost;
    long long mat_stride;
    int32_t *rowmatch, *colmatch;
    int ms1, ms2;
    float *lb;
    int32_t *status;
    SliceLayout layout;
};

template <int S>
__global__ void ged_lsape_kernel(LsapeParams prm)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x * (blockDim.x >> 5) + warp;
    if (b >= prm.B) return;
    const Slice sm = carve(smem_raw + (size_t)warp * prm.layout.bytes, prm.layout);
    const int n1 = prm.n1[b], n2 = prm.n2[b], C = n2 + 1, RC = (n1 + 1) * C;
    const float *src = prm.cost + (long long)b * prm.mat_stride;
    for (int e = lane; e < RC; e += 32) sm.cost[e] = src[e];
    __syncwarp();
    const int st = lsape_warp<S>(sm.cost, n1, n2, sm, nullptr);
    if (st == 0) {
        for (int i = lane; i < n1; i += 32) prm.rowmatch[(long long)b * prm.ms1 + i] = sm.rowmatch[i];
        for (int a = lane; a < n2; a += 32) prm.colmatch[(long long)b * prm.ms2 + a] = sm.colmatch[a];
        if (prm.lb) {
            Pair P;
            P.n1 = n1; P.n2 = n2; P.R = n1 + 1; P.C = C; P.N = n1 + n2; P.W1 = P.W2 = 1; P.node_cost = nullptr;
            const double s = matched_sum(P, sm, sm.cost, sm.rowmatch, sm.colmatch);   // ged_env.py:349
            if (lane == 0) prm.lb[b] = (float)s;
        }
    } else if (prm.lb && lane == 0) prm.lb[b] = NAN;
    if (prm.status && lane == 0) prm.status[b] = st;
}

struct KvParams {
    ged_pairs pairs;
    const float *x;
    float *cost;
    long long mat_stride;
    SliceLayout layout;
};

template <int S>
__global__ void ged_kv_kernel(KvParams prm)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x * (blockDim.x >> 5) + warp;
    if (b >= prm.pairs.num_pairs) return;
    const Slice sm = carve(smem_raw + (size_t)warp * prm.layout.bytes, prm.layout);
    Pair P;
    int status = 0;
    load_pair<S>(prm.pairs, b, sm, P, status);
    P.maxdeg = max_degree(sm.bits1, P.R, P.W1, sm.bits2, P.C, P.W2);
    float *dst = prm.cost + (long long)b * prm.mat_stride;
    kv_apply<S>(P, sm, prm.x + (long long)b * prm.mat_stride, dst, sm.cost);      // the output buffer doubles as Y scratch
    __syncwarp();
    for (int e = lane; e < P.R * P.C; e += 32) dst[e] = sm.cost[e];
}

struct ScoreParams {
    ged_pairs pairs;
    int B;
    const int32_t *base_idx, *rowmatch, *colmatch;
    int ms1, ms2;
    float *ged;
    SliceLayout layout;
};

template <int S>
__global__ void ged_score_kernel(ScoreParams prm)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x * (blockDim.x >> 5) + warp;
    if (b >= prm.B) return;
    const Slice sm = carve(smem_raw + (size_t)warp * prm.layout.bytes, prm.layout);
    Pair P;
    int status = 0;
    load_pair<S>(prm.pairs, prm.base_idx ? prm.base_idx[b] : b, sm, P, status);
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

// ---------------------------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------------------------
int slots_for(int max_n1, int max_n2)
{
    const int m = max_n1 > max_n2 ? max_n1 : max_n2;
    const int S = (m + 31) / 32;
    return S < 1 ? 1 : S;
}

int warps_per_block_for(int slice_bytes)
{
    static int override_wpb = -1;
    if (override_wpb < 0) {
        const char *e = getenv("GED_B200_WARPS_PER_BLOCK");
        override_wpb = e ? atoi(e) : 0;
    }
    if (override_wpb > 0) return override_wpb;
    // one warp per CTA keeps CTA granularity minimal; 32 CTAs/SM is the hardware cap, so for slices small enough
    // that more than 32 would fit, pack two warps per CTA
    return (slice_bytes + 1024) * 32 <= kMaxSmemBytes / 2 ? 2 : 1;
}

template <typename KernelT, typename ParamsT>
int launch_sized(KernelT kernel, const ParamsT &prm, int count, int slice_bytes, cudaStream_t stream, const char *name)
{
    int wpb = warps_per_block_for(slice_bytes);
    while (wpb > 1 && (long long)wpb * slice_bytes > kMaxSmemBytes) --wpb;
    static int pad = -1;                    // occupancy experiments only: extra dynamic shared memory per CTA
    if (pad < 0) { const char *e = getenv("GED_B200_SMEM_PAD"); pad = e ? atoi(e) : 0; }
    const int smem = wpb * slice_bytes + pad;
    if (smem > kMaxSmemBytes) return fail(GED_E_TOO_LARGE, "pair too large for one warp's shared-memory slice");
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail_cuda(e, name);
    const int blocks = (count + wpb - 1) / wpb;
    kernel<<<blocks, wpb * 32, smem, stream>>>(prm);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda(e, name);
    return GED_OK;
}

#define GED_DISPATCH_S(S_, KERNEL, ...)                                                                \
    switch (S_) {                                                                                      \
    case 1: return launch_sized(KERNEL<1>, __VA_ARGS__);                                               \
    case 2: return launch_sized(KERNEL<2>, __VA_ARGS__);                                               \
    case 3: return launch_sized(KERNEL<3>, __VA_ARGS__);                                               \
    case 4: return launch_sized(KERNEL<4>, __VA_ARGS__);                                               \
    case 5: return launch_sized(KERNEL<5>, __VA_ARGS__);                                               \
    case 6: return launch_sized(KERNEL<6>, __VA_ARGS__);                                               \
    case 7: return launch_sized(KERNEL<7>, __VA_ARGS__);                                               \
    default: return fail(GED_E_TOO_LARGE, "graphs above 224 nodes are not supported by this build");   \
    }

int check_pairs(const ged_pairs &p)
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

}  // namespace

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
    SolveParams prm;
    prm.pairs = d->pairs;
    prm.B = d->B;
    prm.base_idx = d->base_idx;
    prm.edit_u = d->edit_u;
    prm.edit_v = d->edit_v;
    prm.score_adj1 = d->score_adj1;
    prm.score_adj1_off = d->score_adj1_off;
    prm.score_idx = d->score_idx;
    prm.max_iter = d->max_iter;
    prm.solver_kind = d->solver_kind;
    prm.x_init = d->x_init;
    prm.mat_stride = d->mat_stride;
    prm.workspace = d->workspace;
    prm.order = d->order;
    prm.count = d->order ? d->order_count : d->B;
    if (prm.count < 0 || prm.count > d->B) return fail(GED_E_BADARG, "ged_solve_batch: order_count out of range");
    if (prm.count == 0) return GED_OK;
    prm.out = *o;
    prm.layout = make_layout(d->pairs.max_n1, d->pairs.max_n2);
    const int S = slots_for(d->pairs.max_n1, d->pairs.max_n2);
    GED_DISPATCH_S(S, ged_solve_kernel, prm, prm.count, prm.layout.bytes, (cudaStream_t)stream, "ged_solve_batch")
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
    LsapeParams prm{B, n1, n2, cost, (long long)mat_stride, rowmatch, colmatch, ms1, ms2, lb, status, make_layout(max_n1, max_n2)};
    const int S = slots_for(max_n1, max_n2);
    GED_DISPATCH_S(S, ged_lsape_kernel, prm, B, prm.layout.bytes, (cudaStream_t)stream, "ged_lsape_batch")
}

int ged_kv_batch(const ged_pairs *pairs, const float *x, float *cost, int64_t mat_stride, void *stream)
{
    if (!pairs || !x || !cost) return fail(GED_E_BADARG, "ged_kv_batch: null pointer");
    if (int rc = check_pairs(*pairs)) return rc;
    if (pairs->num_pairs == 0) return GED_OK;
    if (mat_stride < (int64_t)(pairs->max_n1 + 1) * (pairs->max_n2 + 1)) return fail(GED_E_BADARG, "ged_kv_batch: mat_stride too small");
    KvParams prm{*pairs, x, cost, (long long)mat_stride, make_layout(pairs->max_n1, pairs->max_n2)};
    const int S = slots_for(pairs->max_n1, pairs->max_n2);
    GED_DISPATCH_S(S, ged_kv_kernel, prm, pairs->num_pairs, prm.layout.bytes, (cudaStream_t)stream, "ged_kv_batch")
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
    const int S = slots_for(pairs->max_n1, pairs->max_n2);
    GED_DISPATCH_S(S, ged_score_kernel, prm, B, prm.layout.bytes, (cudaStream_t)stream, "ged_score_batch")
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

void ged_limits(int32_t *max_nodes_per_graph, int32_t *max_matrix_elems)
{
    if (max_nodes_per_graph) *max_nodes_per_graph = 32 * kMaxS;
    if (max_matrix_elems) {
        // largest square pair whose slice fits
        int n = 1;
        while (make_layout(n + 1, n + 1).bytes <= kMaxSmemBytes) ++n;
        *max_matrix_elems = (n + 1) * (n + 1);
    }
}

int64_t ged_solve_smem_bytes(int32_t max_n1, int32_t max_n2) { return make_layout(max_n1, max_n2).bytes; }

int ged_abi_version(void) { return GED_ABI_VERSION; }
const char *ged_version(void) { return "ged_b200 0.1.0 (sm_100a; canonical order v1)"; }
const char *ged_last_error(void) { return g_err; }

}  // extern "C"
