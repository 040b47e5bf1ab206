/* tools/umma_rate.cu -- what one tcgen05.mma kind::i8 instruction costs on B200 when N is small: cycles per instruction for
 * back-to-back M = 128, K = 32 MMAs from shared memory, accumulating into ONE accumulator (dependent chain) or alternating
 * between two / four, for N = 32, 64, 80, 128, 256. One CTA on one SM; clock64 around issue + tcgen05.commit + mbarrier wait.
 *   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_rate tools/umma_rate.cu */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t a, uint32_t lbo, uint32_t sbo) { return (uint64_t)((a & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46); }
__device__ __forceinline__ void mma(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                 :: "r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
template <int N_ACC>
__global__ void __launch_bounds__(128) rate_kernel(int n_cols, int n_mma, long long *cycles, int a_off, int b_off)
{
    extern __shared__ __align__(1024) unsigned char smem[];      /* A: 128 x 256 B (8 k-steps), B: 256 x 256 B, core-matrix layout */
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x;
    for (int i = tid; i < (128 * 256 + 256 * 256) / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u * (i & 3);
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tslot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tslot;
    if (tid == 0) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_cols >> 3) << 17) | (8u << 24);
        const uint32_t a = smem_u32(smem), b = smem_u32(smem + 128 * 256);
        uint32_t phase = 0;
        uint64_t da[8], db[8];
        for (int ks = 0; ks < 8; ks++) { da[ks] = desc(a + ks * 256 + a_off, 128, 2048); db[ks] = desc(b + ks * 256 + b_off, 128, 2048); }
        for (int rep = 0; rep < 3; rep++) {                      /* the last repetition is reported */
            const long long t0 = clock64();
            for (int i = 0; i < n_mma; i += 8) {                 /* descriptors and accumulator addresses are loop constants: nothing but the issue is timed */
#pragma unroll
                for (int ks = 0; ks < 8; ks++)
                    mma(tmem + (uint32_t)(ks % N_ACC) * (512 / N_ACC), da[ks], db[ks], idesc, (i + ks) >= N_ACC);
            }
            const long long t1 = clock64();
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
            phase ^= 1;
            const long long t2 = clock64();
            cycles[0] = t1 - t0; cycles[1] = t2 - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u) : "memory");
}
int main()
{
    long long *d, h[2];
    CK(cudaMalloc(&d, 16));
    const size_t smem = 128 * 256 + 256 * 256 + 1024;
    CK(cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(rate_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    printf("tcgen05.mma.cta_group::1.kind::i8, M = 128, K = 32, operands in shared memory (no swizzle); 64 back-to-back instructions, one issuing thread\n");
    printf("%6s %6s %18s %22s %14s\n", "N", "accs", "issue cycles/MMA", "issue+complete cyc/MMA", "int8 MAC/clk");
    const int ns[] = { 32, 64, 80, 128, 256 }, accs[] = { 1, 2, 4 };
    for (int n : ns)
        for (int a : accs) {
            if (a * n > 512) continue;
            if (a == 1) rate_kernel<1><<<1, 128, smem>>>(n, 64, d, 0, 0);
            else if (a == 2) rate_kernel<2><<<1, 128, smem>>>(n, 64, d, 0, 0);
            else rate_kernel<4><<<1, 128, smem>>>(n, 64, d, 0, 0);
            CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
            printf("%6d %6d %18.1f %22.1f %14.0f\n", n, a, h[0] / 64.0, h[1] / 64.0, 128.0 * n * 32 * 64 / h[1]);
        }
    printf("operand start moved off the 128-byte core-matrix boundary (the shifted formulation of tools/tc5_gemm_bench.cu), N = 80, 1 accumulator\n");
    const int offs[][2] = { { 0, 0 }, { 16, 0 }, { 64, 0 }, { 0, 16 }, { 16, 16 } };
    for (auto &o : offs) {
        rate_kernel<1><<<1, 128, smem>>>(80, 64, d, o[0], o[1]);
        CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
        printf("  A + %2d B, B + %2d B: %6.1f cycles/MMA\n", o[0], o[1], h[1] / 64.0);
    }
    return 0;
}
