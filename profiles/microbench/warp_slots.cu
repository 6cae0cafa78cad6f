// Where do the warps of small persistent CTAs land?  Prints, per SM, the hardware warp slot (%warpid) of every warp of
// the CTAs resident on it (96-thread CTAs, 2 per SM by shared-memory size), to check "scheduler = warp slot mod 4".
//   nvcc -arch=sm_100a -o warp_slots warp_slots.cu && ./warp_slots [threads] [ctas_per_sm]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
extern __shared__ double sm[];
__global__ void k(int *out, int spin)
{
    unsigned smid, warpid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
    const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if ((threadIdx.x & 31) == 0) { out[(blockIdx.x * nw + w) * 2] = smid; out[(blockIdx.x * nw + w) * 2 + 1] = warpid; }
    long long t0 = clock64();
    while (clock64() - t0 < spin) { }     // keep every CTA resident until all have started
    sm[threadIdx.x] = 1.0;
}
int main(int argc, char **argv)
{
    const int nt = argc > 1 ? atoi(argv[1]) : 96, per = argc > 2 ? atoi(argv[2]) : 2;
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int grid = p.multiProcessorCount * per, nw = nt / 32;
    const int smem = (227 * 1024 / per) - 2048;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int *d; cudaMalloc(&d, grid * nw * 2 * sizeof(int));
    k<<<grid, nt, smem>>>(d, 2000000);
    cudaDeviceSynchronize();
    int *h = (int *)malloc(grid * nw * 2 * sizeof(int));
    cudaMemcpy(h, d, grid * nw * 2 * sizeof(int), cudaMemcpyDeviceToHost);
    for (int b = 0; b < grid && b < 12; ++b) {
        printf("cta %3d sm %3d warp slots:", b, h[b * nw * 2]);
        for (int w = 0; w < nw; ++w) printf(" %d", h[(b * nw + w) * 2 + 1]);
        printf("\n");
    }
    // histogram of slot patterns
    int pat[64] = {0};
    for (int b = 0; b < grid; ++b) pat[h[b * nw * 2 + 1] & 63]++;
    printf("first-warp slot histogram:");
    for (int i = 0; i < 64; ++i) if (pat[i]) printf(" %d:%d", i, pat[i]);
    printf("\n%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
