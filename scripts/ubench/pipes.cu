// Micro-benchmark of the instruction latencies / issue intervals the LAP scan step depends on (development aid).
#include <cstdio>
#include <cuda_runtime.h>
#define REP 256
template <int MODE>
__global__ void k(double *out, const float *in, int n, long long *cyc)
{
    double a0 = in[0], a1 = in[1], a2 = in[2], a3 = in[3], a4 = in[4], a5 = in[5], a6 = in[6], a7 = in[7];
    const double c = in[8];
    unsigned x0 = threadIdx.x + in[9], x1 = x0 * 3, x2 = x0 * 5, x3 = x0 * 7;
    float f0 = in[10] + threadIdx.x, f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3;
    __shared__ float sm[1024];
    sm[threadIdx.x] = threadIdx.x;
    __syncthreads();
    int idx = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int r = 0; r < REP; ++r) {
            if (MODE == 0) a0 = __dadd_rn(a0, c);                                   // dependent DADD latency
            if (MODE == 1) { a0 = __dadd_rn(a0, c); a1 = __dadd_rn(a1, c); a2 = __dadd_rn(a2, c); a3 = __dadd_rn(a3, c);
                             a4 = __dadd_rn(a4, c); a5 = __dadd_rn(a5, c); a6 = __dadd_rn(a6, c); a7 = __dadd_rn(a7, c); }  // 8 independent
            if (MODE == 2) { x0 = __reduce_min_sync(0xffffffffu, x0) + 1; }        // dependent CREDUX
            if (MODE == 3) { a0 = (a0 < a1) ? a2 : a0; a1 = __dadd_rn(a1, c); }      // DSETP + FSEL chain
            if (MODE == 4) { a0 = __dadd_rn(a0, (double)f0); f0 = f0 + 1.0f; }      // F2F in chain
            if (MODE == 5) { idx = (int)sm[idx & 1023]; }                          // dependent LDS
            if (MODE == 6) { x0 = (x0 ^ x1) + 1; }                                  // dependent ALU
            if (MODE == 7) { x0 = (x0 ^ x1) + 1; x1 = (x1 ^ x2) + 3; x2 = (x2 ^ x3) + 5; x3 = (x3 ^ x0) + 7; }
            if (MODE == 8) { x0 = __any_sync(0xffffffffu, x0 > 5) ? x0 + 1 : x0 + 2; }   // vote chain
            if (MODE == 9) { a0 = __dadd_rn(a0, c); a1 = __dadd_rn(a1, c); }            // 2 independent DADD
            if (MODE == 10) { a0 = __dadd_rn(a0, c); a1 = __dadd_rn(a1, c); a2 = __dadd_rn(a2, c); a3 = __dadd_rn(a3, c); }  // 4 independent
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[threadIdx.x + blockIdx.x * blockDim.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + x0 + x1 + x2 + x3 + f0 + f1 + f2 + f3 + idx;
}
template <int MODE> void run(const char *name, int per, int warps)
{
    double *out; float *in; long long *cyc;
    cudaMalloc(&out, 8 * 4096); cudaMalloc(&in, 64); cudaMalloc(&cyc, 8 * 64);
    cudaMemset(in, 0, 64);
    const int n = 64;
    k<MODE><<<1, 32 * warps>>>(out, in, n, cyc);
    k<MODE><<<1, 32 * warps>>>(out, in, n, cyc);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s warps/SM=%2d  %.2f cycles per op (%.2f per iteration)\n", name, warps, (double)h / (n * REP * per), (double)h / (n * REP));
}
int main()
{
    for (int w : {1, 4, 8, 16}) {
        if (w == 1) {
            run<0>("DADD dependent", 1, 1); run<3>("DSETP+FSEL+DADD", 1, 1); run<4>("F2F+DADD chain", 1, 1);
            run<2>("CREDUX dependent(+IADD)", 1, 1); run<5>("LDS dependent (+cvt)", 1, 1); run<6>("ALU dependent (2 ops)", 1, 1);
            run<8>("VOTE chain", 1, 1);
        }
        if (w == 1) { run<1>("DADD 8 independent", 8, 1); run<9>("DADD 2 independent", 2, 1); run<10>("DADD 4 independent", 4, 1); run<7>("ALU 4 indep x2", 8, 1); }
        if (w == 4) { run<1>("DADD 8 independent", 8, 4); run<7>("ALU 4 indep x2", 8, 4); run<2>("CREDUX dependent", 1, 4); }
        if (w == 8) { run<1>("DADD 8 independent", 8, 8); run<2>("CREDUX dependent", 1, 8); }
        if (w == 16) { run<1>("DADD 8 independent", 8, 16); run<7>("ALU 4 indep x2", 8, 16); run<2>("CREDUX dependent", 1, 16);}
    }
    return 0;
}
