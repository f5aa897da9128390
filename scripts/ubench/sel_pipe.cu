// Can the selects of the LAP scan step leave the (half-rate) ALU pipe?  Whole-GPU streams of independent instruction chains,
// 32 warps per SM, timed with CUDA events; prints warp-instructions/s per variant:
//   sel      : selp.b32 (SEL, ALU pipe)
//   pimad    : @p mad.lo.u32 x, a, one, 0  (predicated IMAD by an opaque 1: FMA pipe, same effect as  x = p ? a : x)
//   lop      : lop3 only (ALU pipe reference)
//   lop+sel  : 1 : 1
//   lop+pimad: 1 : 1
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o sel_pipe sel_pipe.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kUnroll = 32;

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(unsigned *out, const unsigned *in, int trips)
{
    unsigned x[8], a[8];
    unsigned one = in[1];              // 1, opaque to the compiler
    const unsigned xm = in[2] | 0x5a5a5a5au;
#pragma unroll
    for (int j = 0; j < 8; ++j) { x[j] = threadIdx.x * (2 * j + 3) + in[0]; a[j] = x[j] ^ 0x1234567u; }
    int pred = (threadIdx.x ^ in[3]) & 1;
#pragma unroll 1
    for (int t = 0; t < trips; ++t) {
#pragma unroll
        for (int r = 0; r < kUnroll; ++r) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const bool alt = (j & 1);
                if (MODE == 0 || (MODE == 3 && alt))
                    asm volatile("{.reg .pred p; setp.ne.u32 p, %2, 0; selp.b32 %0, %1, %0, p;}" : "+r"(x[j]) : "r"(x[(j + 3) & 7]), "r"(pred));
                else if (MODE == 1 || (MODE == 4 && alt))
                    asm volatile("{.reg .pred p; setp.ne.u32 p, %2, 0; @p mad.lo.u32 %0, %1, %3, 0;}" : "+r"(x[j]) : "r"(x[(j + 3) & 7]), "r"(pred), "r"(one));
                else
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(xm), "r"(x[(j + 1) & 7]));
            }
        }
        pred ^= (int)(x[0] & 1u) & (int)in[4];
    }
    unsigned s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += x[j] + a[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static double run(int sms, unsigned *out, const unsigned *in, int trips)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<sms, 1024>>>(out, in, trips / 4);
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<sms, 1024>>>(out, in, trips);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double rate = (double)sms * 32 * (double)trips * kUnroll * 8 / (ms * 1e-3);
        if (rate > best) best = rate;
    }
    return best;
}

int main()
{
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    unsigned *out, *in;
    cudaMalloc(&out, sizeof(unsigned) * sms * 1024);
    cudaMalloc(&in, 64);
    unsigned h[16] = {0, 1, 0, 0, 0};
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    const double sel = run<0>(sms, out, in, 1000), pimad = run<1>(sms, out, in, 1000), lop = run<2>(sms, out, in, 1000);
    const double lopsel = run<3>(sms, out, in, 1000), loppimad = run<4>(sms, out, in, 1000);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("{\"error\": \"cuda\"}\n"); return 1; }
    printf("{\"setp_sel_slots_per_s\": %.4e, \"setp_pimad_slots_per_s\": %.4e, "
           "\"lop_per_s\": %.4e, \"lop_and_setp_sel_slots_per_s\": %.4e, \"lop_and_setp_pimad_slots_per_s\": %.4e}\n", sel, pimad, lop, lopsel, loppimad);
    return 0;
}
