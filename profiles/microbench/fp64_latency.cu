// Micro-benchmark: FP64 dependent-issue latency and throughput vs resident warps on B200 (sm_100a).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false fp64_latency.cu -o fp64_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int OP>
__global__ void chain(double *out, long long *cyc, int iters)
{
    double a[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) a[j] = 1.0 + threadIdx.x * 1e-3 + j;
    const double m = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int j = 0; j < ILP; ++j) {
                if (OP == 0) a[j] = __dadd_rn(a[j], c);
                else if (OP == 1) a[j] = __dmul_rn(a[j], m);
                else a[j] = __fma_rn(a[j], m, c);
            }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += a[j];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int ILP, int OP>
void run(const char *name, int warps_per_sm, int sms)
{
    double *out; long long *cyc;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    // one block per SM with warps_per_sm warps
    chain<ILP, OP><<<sms, 32 * warps_per_sm>>>(out, cyc, iters);
    chain<ILP, OP><<<sms, 32 * warps_per_sm>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per_op = (double)h / (iters * 16.0 * ILP);
    double ops_per_clk_sm = 32.0 * warps_per_sm / per_op;   // lanes per clock per SM
    printf("%-5s ILP=%d warps/SM=%2d : %.2f cycles per op per warp (chain step %.1f cyc), %.1f lanes/clk/SM\n",
           name, ILP, warps_per_sm, per_op, per_op * ILP, ops_per_clk_sm);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("%s, %d SMs\n", p.name, sms);
    for (int w : {1, 4, 8, 16, 32}) run<1, 0>("DADD", w, sms);
    for (int w : {1, 4, 8, 16, 32}) run<1, 2>("DFMA", w, sms);
    for (int w : {1, 4, 8, 16, 32}) run<2, 0>("DADD", w, sms);
    for (int w : {1, 4, 8, 16, 32}) run<4, 0>("DADD", w, sms);
    for (int w : {4, 16, 32}) run<8, 2>("DFMA", w, sms);
    for (int w : {1, 4}) run<1, 1>("DMUL", w, sms);
    return 0;
}
