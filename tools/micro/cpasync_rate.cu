// Microbenchmark: issue cost of global->shared staging per thread: cp.async 16 B (LDGSTS, zfill form) vs LDG.128 + STS.128
// batches, for 4 / 8 / 16 warps per CTA, one CTA per SM, data larger than L2.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp16(void *dst, const void *src, uint32_t bytes)
{ asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(smem_u32(dst)), "l"(src), "r"(bytes) : "memory"); }

template <int MODE>
__global__ void k(const uint4 *x, size_t n_vec, int per_thread, int rounds, long long *out, uint4 *sink)
{
    extern __shared__ uint4 sm[];
    const int tid = threadIdx.x, nt = blockDim.x;
    const uint4 *src = x + ((size_t)blockIdx.x * 7919 * 4096) % (n_vec - (size_t)rounds * per_thread * nt - 4096);
    long long t0 = clock64();
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (int r = 0; r < rounds; ++r) {
        const uint4 *s = src + (size_t)r * per_thread * nt;
        if (MODE == 0) {
            for (int i = 0; i < per_thread; ++i) cp16(sm + (i * nt + tid) % 8192, s + i * nt + tid, (i + tid + r) >= 0 ? 16u : 0u);
            asm volatile("cp.async.commit_group;\n" ::: "memory");
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        } else {
            for (int i0 = 0; i0 < per_thread; i0 += 8) {
                uint4 v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = __ldg(s + (i0 + j) * nt + tid);
#pragma unroll
                for (int j = 0; j < 8; ++j) sm[((i0 + j) * nt + tid) % 8192] = v[j];
            }
        }
    }
    __syncthreads();
    long long t1 = clock64();
    acc = sm[tid];
    if (acc.x == 0x12345678u) sink[tid] = acc;
    if (tid == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

int main()
{
    const size_t n_vec = (size_t)1 << 27;   // 2 GiB
    uint4 *x, *sink; long long *d, h;
    cudaMalloc(&x, n_vec * 16); cudaMemset(x, 1, n_vec * 16); cudaMalloc(&sink, 1 << 16); cudaMalloc(&d, 8);
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    printf("mode warps per_thread clk_per_round clk_per_copy_per_thread GBps_total\n");
    for (int mode = 0; mode < 2; ++mode)
        for (int warps : {4, 8, 16})
            for (int per_thread : {8, 16, 32, 64}) {
                const int rounds = 64;
                for (int rep = 0; rep < 2; ++rep) {
                    if (mode == 0) k<0><<<148, warps * 32, 128 * 1024>>>(x, n_vec, per_thread, rounds, d, sink);
                    else k<1><<<148, warps * 32, 128 * 1024>>>(x, n_vec, per_thread, rounds, d, sink);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
                }
                cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
                const double clk_round = (double)h / rounds;
                const double bytes = 148.0 * rounds * per_thread * warps * 32 * 16;
                printf("%s %d %d %.0f %.1f %.0f\n", mode ? "ldg+sts" : "cp.async", warps, per_thread, clk_round, clk_round / per_thread,
                       bytes / ((double)h / 1.9e9) / 1e9);
            }
    return 0;
}
