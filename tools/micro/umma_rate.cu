// Microbenchmark: cost of small tcgen05.mma (M=128, K=16, bf16, SS, no-swizzle K-major) as a function of N and of
// how many independent accumulators consecutive MMAs rotate over.   nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n)
{ return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }

__global__ void __launch_bounds__(128) k(int N, int rot, int iters, int a_rot, int a_tmem, long long *out)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u ^ (i * 2654435761u & 0x00ff00ffu);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(&bar)), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t taddr = slot;
    if (warp == 0) {
        uint32_t leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(leader));
        const uint32_t a_units = smem_u32(smem) >> 4;                 // A: rows of 16 B, chunk plane = 132 units
        const uint32_t b_units = (smem_u32(smem) + 96 * 1024) >> 4;    // B: [2][N][16 B]
        constexpr uint32_t kHi = 8u | (1u << 14);
        const uint32_t idesc = idesc_bf16(128, N);
        long long t0 = clock64();
        if (rot == 1 && a_rot == 1 && !a_tmem) {
            // descriptors hoisted: the loop body is nothing but UTCHMMA -> measures the pure dispatch / execution rate
            const uint64_t ad = ((uint64_t)kHi << 32) | (uint64_t)(a_units + (132u << 16));
            const uint64_t bd = ((uint64_t)kHi << 32) | (uint64_t)(b_units + ((uint32_t)N << 16));
            if (leader) {
#pragma unroll 8
                for (int it = 0; it < iters; ++it)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                                 :: "r"(taddr), "l"(ad), "l"(bd), "r"(idesc), "r"(1u));
            }
        } else
        #pragma unroll 4
        for (int it = 0; it < iters; ++it) {
            const uint32_t acc = taddr + (uint32_t)((it & (rot - 1)) * N);
            const uint32_t aoff = (uint32_t)((it & (a_rot - 1)) * 264);
            const uint64_t ad = ((uint64_t)kHi << 32) | (uint64_t)(a_units + aoff + (132u << 16));
            const uint64_t bd = ((uint64_t)kHi << 32) | (uint64_t)(b_units + ((uint32_t)N << 16));
            if (leader) {
                if (a_tmem)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                                 :: "r"(acc), "r"(taddr + 256u), "l"(bd), "r"(idesc), "r"(1u));
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                                 :: "r"(acc), "l"(ad), "l"(bd), "r"(idesc), "r"(1u));
            }
        }
        long long t1 = clock64();
        if (leader)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(smem_u32(&bar)) : "memory");
        asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}\n"
                     :: "r"(smem_u32(&bar)), "r"(0u) : "memory");
        long long t2 = clock64();
        if (tid == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(taddr), "r"(512u));
}

int main()
{
    long long *d, h[2];
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    const int iters = 2048;
    printf("mode N rot a_rot issue_clk_per_mma total_clk_per_mma\n");
    for (int a_tmem = 0; a_tmem < 2; ++a_tmem)
    for (int N : {16, 32, 48, 64, 80, 96, 128, 144, 192, 256})
        for (int rot : {1, 2, 4})
            for (int a_rot : {1, 8}) {
                if (rot * N > 256) continue;
                for (int rep = 0; rep < 2; ++rep) {
                    k<<<1, 128, 160 * 1024>>>(N, rot, iters, a_rot, a_tmem, d);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
                }
                cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                printf("%s %d %d %d %.1f %.1f\n", a_tmem ? "TS" : "SS", N, rot, a_rot, (double)h[0] / iters, (double)h[1] / iters);
            }
    return 0;
}
