// pipe_bench.cu -- integer-pipe issue rates on sm_100a (warp-instructions per clock per SM), register-resident chains.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_bench tools/pipe_bench.cu ; run: tools/pipe_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(int *sink, int iters, int seed)
{
    int a[8], b[8];
    long long w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 7 + i * 13 + seed; b[i] = a[i] ^ 0x55; w[i] = a[i]; }
    const int m = (int)threadIdx.x | 1, c = seed + 3;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (MODE == 0) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(c));
                if (MODE == 1) asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(m));
                if (MODE == 2) asm volatile("add.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(m));
                if (MODE == 3) asm volatile("shf.r.wrap.b32 %0, %0, %1, 15;" : "+r"(a[i]) : "r"(b[i]));
                if (MODE == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(m), "r"(c));
                if (MODE == 5) asm volatile("mul.hi.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(m));
                if (MODE == 6) { asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(c)); asm volatile("add.s32 %0, %0, %1;" : "+r"(b[i]) : "r"(m)); }
                if (MODE == 7) { asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(m)); asm volatile("add.s32 %0, %0, %1;" : "+r"(b[i]) : "r"(m)); }
                if (MODE == 8) { asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(m)); asm volatile("add.s32 %0, %0, %1;" : "+r"(b[i]) : "r"(m)); asm volatile("shf.r.wrap.b32 %0, %0, %1, 15;" : "+r"(a[i]) : "r"(b[i])); }
                if (MODE == 9) asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(a[i]) : "r"(m));
                if (MODE == 10) asm volatile("max.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(m));
                if (MODE == 11) { asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(c)); asm volatile("add.s32 %0, %0, %1;" : "+r"(b[i]) : "r"(m)); asm volatile("add.s32 %0, %0, %1;" : "+r"(b[i]) : "r"(c)); }
            }
        }
    }
    int r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r ^= a[i] ^ b[i] ^ (int)w[i] ^ (int)(w[i] >> 32);
    if (r == 0x7fffffff) sink[threadIdx.x] = r;
}

template <int MODE>
static void run(const char *name, int per, int warps_per_sm)
{
    int *sink; cudaMalloc(&sink, 1 << 16);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int iters = 2048, threads = 256, blocks = p.multiProcessorCount * warps_per_sm / 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(sink, iters, rep); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    const double inst = (double)blocks * (threads / 32) * iters * 32.0 * per;      // warp instructions
    const double clocks = best * 1e-3 * clk * 1e3;
    printf("%-34s warps/SM %2d : %.2f warp-inst/clk/SM  (%.1f lane-ops/clk/SM)\n", name, warps_per_sm, inst / clocks / p.multiProcessorCount, 32 * inst / clocks / p.multiProcessorCount);
    cudaFree(sink);
}

int main()
{
    for (int w : {16, 32}) {
        run<0>("IMAD", 1, w); run<1>("IMAD.WIDE", 1, w); run<2>("IADD", 1, w); run<3>("SHF", 1, w); run<4>("LOP3", 1, w);
        run<5>("IMAD.HI", 1, w); run<9>("PRMT", 1, w); run<10>("IMNMX", 1, w);
        run<6>("IMAD + IADD", 2, w); run<11>("IMAD + 2 IADD", 3, w); run<7>("IMAD.WIDE + IADD", 2, w); run<8>("IMAD.WIDE + IADD + SHF", 3, w);
    }
    return 0;
}
