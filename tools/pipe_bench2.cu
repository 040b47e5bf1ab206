// pipe_bench2.cu -- issue cost of IMAD.WIDE next to ALU work on sm_100a: cycles per loop body, to be read with the
// SASS instruction counts of each body (cuobjdump -sass). nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(int *sink, int iters, int seed, long long *clk_out)
{
    long long w[8]; int a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { w[i] = threadIdx.x * 7 + i * 13 + seed; a[i] = (int)w[i] ^ 0x1234; b[i] = a[i] * 3; }
    const int m = (int)threadIdx.x | 1, c = seed + 3;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) w[i] = (long long)(int)w[i] * m + w[i];                                   // IMAD.WIDE chain
            if (MODE == 1) { w[i] = (long long)(int)w[i] * m + w[i]; b[i] = (b[i] ^ m) & (c + it); }     // + 1 LOP3
            if (MODE == 2) { w[i] = (long long)(int)w[i] * m + w[i]; b[i] = (b[i] ^ m) & (c + it); a[i] = (a[i] | m) ^ (c + it); }  // + 2 LOP3
            if (MODE == 3) { a[i] = a[i] * m + c; }                                                   // IMAD chain
            if (MODE == 4) { a[i] = a[i] * m + c; b[i] = (b[i] ^ m) & (c + it); }                        // IMAD + 1 LOP3
            if (MODE == 5) { b[i] = (b[i] ^ m) & (c + it); }                                              // LOP3 only
            if (MODE == 6) { w[i] = (long long)(int)w[i] * m + w[i]; a[i] = (int)(w[i] >> 15); b[i] ^= a[i]; }  // WIDE + funnel SHF + LOP
            if (MODE == 7) { int hi = __mulhi(a[i] << 1, m << 16); a[i] = hi + c; }                        // IMAD.HI variant
        }
    }
    const long long t1 = clock64();
    int r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r ^= a[i] ^ b[i] ^ (int)w[i] ^ (int)(w[i] >> 32);
    if (r == 0x7fffffff) sink[threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk_out = t1 - t0;
}

template <int MODE> static void run(const char *name, int warps_per_sm)
{
    int *sink; long long *clk; cudaMalloc(&sink, 1 << 16); cudaMalloc(&clk, 8);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int iters = 4096, blocks = p.multiProcessorCount * warps_per_sm / 8;
    long long h = 0;
    for (int rep = 0; rep < 3; rep++) { k<MODE><<<blocks, 256>>>(sink, iters, rep, clk); cudaDeviceSynchronize(); }
    cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    // cycles per (8-chain) body per warp, with warps_per_sm/4 warps sharing a scheduler
    printf("%-28s warps/SM %2d : %.2f clk per body per scheduler-warp-slot (= %.2f clk per chain step per SMSP)\n", name, warps_per_sm,
           (double)h / iters, (double)h / iters / 8.0 / (warps_per_sm / 4.0));
    cudaFree(sink); cudaFree(clk);
}
int main()
{
    for (int w : {16, 32}) {
        run<0>("WIDE", w); run<1>("WIDE+1LOP", w); run<2>("WIDE+2LOP", w); run<3>("IMAD", w); run<4>("IMAD+1LOP", w);
        run<5>("LOP", w); run<6>("WIDE+SHF64+LOP", w); run<7>("IMAD.HI", w);
    }
}
