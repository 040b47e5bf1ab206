// pipe_bench3.cu -- does the FP64 pipe of sm_100a run next to the FMA-heavy pipe? Cycles per chain step per SM sub-partition
// for IMAD.WIDE alone, DFMA alone and both interleaved (independent register-resident chains), plus the magic-number
// floor (DADD.RM) that an exact Q15 product in double precision would need.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_bench3 tools/pipe_bench3.cu ; run: tools/pipe_bench3
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(int *sink, int iters, int seed, long long *clk_out)
{
    long long w[8]; double d[8], e[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { w[i] = threadIdx.x * 7 + i * 13 + seed; d[i] = (double)w[i] * 1e-3; e[i] = d[i] + 1.0; }
    const int m = (int)threadIdx.x | 1;
    const double dm = 1.0 + 1e-9 * (double)m, magic = 6755399441055744.0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) w[i] = (long long)(int)w[i] * m + w[i];                                   // IMAD.WIDE chain
            if (MODE == 1) d[i] = fma(d[i], dm, d[i]);                                              // DFMA chain
            if (MODE == 2) { w[i] = (long long)(int)w[i] * m + w[i]; d[i] = fma(d[i], dm, d[i]); }      // both
            if (MODE == 3) { w[i] = (long long)(int)w[i] * m + w[i]; d[i] = fma(d[i], dm, d[i]); e[i] = __dadd_rd(e[i], magic) - magic; }  // + floor
            if (MODE == 4) { d[i] = fma(d[i], dm, d[i]); e[i] = fma(e[i], dm, d[i]); }                 // 2 DFMA
        }
    }
    const long long t1 = clock64();
    long long r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r ^= w[i] ^ __double_as_longlong(d[i]) ^ __double_as_longlong(e[i]);
    if (r == 0x7fffffff) sink[threadIdx.x] = (int)r;
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
    printf("%-32s warps/SM %2d : %.2f clk per chain step per SMSP\n", name, warps_per_sm, (double)h / iters / 8.0 / (warps_per_sm / 4.0));
    cudaFree(sink); cudaFree(clk);
}
int main()
{
    for (int w : {16, 32}) {
        run<0>("IMAD.WIDE", w); run<1>("DFMA", w); run<2>("IMAD.WIDE + DFMA", w); run<3>("IMAD.WIDE + DFMA + 2 DADD", w); run<4>("2 DFMA", w);
    }
}
