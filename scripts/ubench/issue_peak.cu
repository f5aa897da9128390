// Measured instruction-issue roofs of one B200 (sm_100a) for the roofline of the GED solve kernel, which is bound by
// instruction issue / the ALU pipe, not by HBM or tensor cores.
// Whole-GPU kernels of independent instruction streams (8 chains per thread, 32 warps per SM), timed with CUDA events:
//   fma   : FFMA only            -> issue limit of an SM sub-partition (1 warp instruction per clock)
//   alu   : LOP3 only            -> the ALU pipe (integer / logic / select / compare: 16 lanes per sub-partition)
//   fp64  : DADD only            -> the FP64 pipe
//   mix   : LOP3 : FFMA = 1 : 1  -> two pipes fed from one issue port
//   solve : the solve kernel's scan-loop mix (ALU : FP64 : FMA : other = 56 : 23 : 8 : 13 per 100, from its SASS)
// Prints one JSON object: warp-instructions per second for each, the SM clock seen, SM count.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o issue_peak issue_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kUnroll = 64;      // instruction groups per loop trip (8 instructions each)

template <int MODE>
__global__ void __launch_bounds__(1024, 2) stream_kernel(float *out, const float *in, int trips, long long *clk)
{
    float f[8];
    unsigned x[8];
    double d[8];
    const float fb = in[1], fc = in[2];
    const unsigned xm = (unsigned)in[3] | 0x5a5a5a5au;
    const double dc = in[4];
#pragma unroll
    for (int k = 0; k < 8; ++k) { f[k] = in[0] + threadIdx.x + k; x[k] = threadIdx.x * (2 * k + 3); d[k] = f[k]; }
    const long long t0 = clock64();
#pragma unroll 1
    for (int t = 0; t < trips; ++t) {
#pragma unroll
        for (int r = 0; r < kUnroll; ++r) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (MODE == 0) f[k] = fmaf(f[k], fb, fc);
                if (MODE == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(xm), "r"(x[(k + 1) & 7]));
                if (MODE == 2) d[k] = __dadd_rn(d[k], dc);
                if (MODE == 3) {
                    if (k & 1) f[k] = fmaf(f[k], fb, fc);
                    else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(xm), "r"(x[(k + 2) & 7]));
                }
                if (MODE == 4) {     // 16 instructions per two groups: 9 ALU, 4 FP64, 2 FMA-pipe, 1 conversion
                    const int slot = (r & 1) * 8 + k;
                    if (slot == 3 || slot == 7 || slot == 11 || slot == 15) d[k] = __dadd_rn(d[k], dc);
                    else if (slot == 5 || slot == 13) f[k] = fmaf(f[k], fb, fc);
                    else if (slot == 9) d[k] = __dadd_rn(d[k], (double)f[k]);
                    else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(xm), "r"(x[(k + 1) & 7]));
                }
            }
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += f[k] + (float)x[k] + (float)d[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

template <int MODE>
static double run(int sms, float *out, const float *in, long long *clk, int trips, double *sm_mhz)
{
    const int grid = sms, block = 1024;      // one CTA per SM: block 0's clock64 span is the kernel's (sm_mhz)
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    stream_kernel<MODE><<<grid, block>>>(out, in, trips / 4, clk);      // warm-up
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        stream_kernel<MODE><<<grid, block>>>(out, in, trips, clk);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        long long c = 0;
        cudaMemcpy(&c, clk, sizeof(c), cudaMemcpyDeviceToHost);
        const double winst = (double)grid * (block / 32) * (double)trips * kUnroll * 8;
        const double rate = winst / (ms * 1e-3);
        if (rate > best) { best = rate; if (sm_mhz) *sm_mhz = (double)c / (ms * 1e-3) / 1e6; }
    }
    return best;
}

int main()
{
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    float *out, *in;
    long long *clk;
    cudaMalloc(&out, sizeof(float) * sms * 1024);
    cudaMalloc(&in, 64);
    cudaMalloc(&clk, 64);
    cudaMemset(in, 0, 64);
    double mhz = 0, m2 = 0;
    const double fma = run<0>(sms, out, in, clk, 2000, &mhz);
    const double alu = run<1>(sms, out, in, clk, 1000, &m2);
    const double fp64 = run<2>(sms, out, in, clk, 500, &m2);
    const double mix = run<3>(sms, out, in, clk, 1500, &m2);
    const double solve = run<4>(sms, out, in, clk, 1000, &m2);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("{\"error\": \"cuda\"}\n"); return 1; }
    printf("{\"sms\": %d, \"sm_mhz\": %.0f, \"fma_winst_per_s\": %.4e, \"alu_winst_per_s\": %.4e, \"fp64_winst_per_s\": %.4e, "
           "\"mix_alu_fma_winst_per_s\": %.4e, \"solve_mix_winst_per_s\": %.4e, \"nominal_issue_winst_per_s\": %.4e}\n",
           sms, mhz, fma, alu, fp64, mix, solve, (double)sms * 4 * mhz * 1e6);
    return 0;
}
