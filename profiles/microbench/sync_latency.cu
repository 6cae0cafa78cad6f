// Micro-benchmark: latency of SHFL, __syncthreads (9 warps), LDS pointer chase, and of a transposed
// 8-value warp reduction on B200.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_shfl(long long *cyc, double *out, int iters)
{
    double v = threadIdx.x * 1.0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) v = __shfl_xor_sync(0xffffffffu, v, 1 + (u & 3));
    }
    long long t1 = clock64();
    if (v == 1.2345) out[0] = v;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}
__global__ void k_shfl_add(long long *cyc, double *out, int iters)
{
    double v = threadIdx.x * 1.0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) v = v + __shfl_xor_sync(0xffffffffu, v, 1 + (u & 3));
    }
    long long t1 = clock64();
    if (v == 1.2345) out[0] = v;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}
__global__ void k_bar(long long *cyc, int iters)
{
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) __syncthreads();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}
__global__ void k_lds(long long *cyc, int *out, int iters)
{
    __shared__ int nxt[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) nxt[i] = (i * 17 + 1) & 255;
    __syncthreads();
    int p = threadIdx.x & 255;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) p = nxt[p];
    }
    long long t1 = clock64();
    if (p == 12345) out[0] = p;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}
// STS -> barrier -> LDS round trip (what one phase boundary of the solver costs)
__global__ void k_sts_bar_lds(long long *cyc, double *out, int iters)
{
    __shared__ double buf[512];
    double v = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            buf[threadIdx.x] = v;
            __syncthreads();
            v = buf[(threadIdx.x + 1) % blockDim.x] + 1.0;
            __syncthreads();
        }
    }
    long long t1 = clock64();
    if (v == 1.2345) out[0] = v;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}

int main()
{
    long long *cyc, h; double *out; int *outi;
    cudaMalloc(&cyc, 8); cudaMalloc(&out, 64); cudaMalloc(&outi, 64);
    const int iters = 2000;
    k_shfl<<<1, 32>>>(cyc, out, iters); k_shfl<<<1, 32>>>(cyc, out, iters);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("SHFL (64-bit = 2 SHFL) dependent : %.1f cycles\n", h / (iters * 8.0));
    k_shfl_add<<<1, 32>>>(cyc, out, iters); k_shfl_add<<<1, 32>>>(cyc, out, iters);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("SHFL + DADD reduction level      : %.1f cycles\n", h / (iters * 8.0));
    for (int nt : {32, 128, 288, 576}) {
        k_bar<<<1, nt>>>(cyc, iters); k_bar<<<1, nt>>>(cyc, iters);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("__syncthreads, %3d threads        : %.1f cycles\n", nt, h / (iters * 8.0));
    }
    k_lds<<<1, 32>>>(cyc, outi, iters); k_lds<<<1, 32>>>(cyc, outi, iters);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("LDS pointer chase                : %.1f cycles\n", h / (iters * 8.0));
    for (int nt : {32, 288}) {
        k_sts_bar_lds<<<1, nt>>>(cyc, out, iters); k_sts_bar_lds<<<1, nt>>>(cyc, out, iters);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("STS+BAR+LDS+DADD+BAR, %3d threads : %.1f cycles\n", nt, h / (iters * 4.0));
    }
    return 0;
}
