// Microbenchmark: global -> shared staging rate of the T-merged conv producer's access pattern, in isolation.
// One CTA per SM; per "slab" LINES bulk copies (cp.async.bulk, 2112 B each) whose sources are PLANE bytes apart (rows of a
// [D][H][W][16 B] volume), advancing one H-row per slab; ring of RING slots with full/empty mbarriers; the consumer warp only
// waits for `full` and releases the slot.  Variants: lanes that issue (1 elected lane vs one lane per line), warps.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" :: "r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) { asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n" :: "r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}\n"
                 :: "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(160)
k(const uint4 *x, size_t plane_vec, size_t row_vec, int lines, int ring, int slabs, int mode, long long *out)
{
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t full[8], empty[8];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int LB = 2112;
    if (tid == 0) {
        for (int i = 0; i < ring; ++i) { mbar_init(full + i, mode == 0 ? 1 : 4); mbar_init(empty + i, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    const uint4 *base = x + (size_t)blockIdx.x * 132 % row_vec + (size_t)(blockIdx.x / 8) * 64 * row_vec;
    long long t0 = clock64();
    if (warp < 4) {
        for (int s = 0; s < slabs; ++s) {
            const int slot = s % ring, q = s / ring;
            if (q >= 1) mbar_wait(empty + slot, (q - 1) & 1);
            uint8_t *dst = sm + (size_t)slot * lines * LB;
            const uint4 *src = base + (size_t)s * row_vec;
            if (mode == 0) {           // one elected lane of warp 0 issues all lines
                if (warp == 0 && lane == 0) {
                    mbar_expect(full + slot, lines * LB);
                    for (int l = 0; l < lines; ++l) bulk(dst + l * LB, src + (size_t)l * plane_vec, LB, full + slot);
                }
            } else {                   // lane j of warp w issues line w + 4 j
                const int l = warp + 4 * lane;
                int n_my = 0;
                for (int j = warp; j < lines; j += 4) ++n_my;
                if (lane == 0) mbar_expect(full + slot, n_my * LB);
                __syncwarp();
                if (l < lines) bulk(dst + l * LB, src + (size_t)l * plane_vec, LB, full + slot);
            }
        }
    } else {
        for (int s = 0; s < slabs; ++s) {
            const int slot = s % ring;
            mbar_wait(full + slot, (s / ring) & 1);
            if (lane == 0) mbar_arrive(empty + slot);
            __syncwarp();
        }
    }
    __syncthreads();
    if (tid == 0 && blockIdx.x == 0) out[0] = clock64() - t0;
}

int main()
{
    const size_t W = 1600, H = 1184, D = 8;
    const size_t row_vec = W, plane_vec = H * W, n_vec = D * plane_vec;       // [D][H][W] of 16 B voxels = 243 MB
    uint4 *x; long long *d, h;
    cudaMalloc(&x, (n_vec + (1 << 20)) * 16); cudaMemset(x, 1, n_vec * 16); cudaMalloc(&d, 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    printf("mode lines ring clk_per_slab B_per_clk_per_SM TBps_at_1.9GHz\n");
    for (int mode = 0; mode < 2; ++mode)
        for (int lines : {6, 7, 24})
            for (int ring : {2, 4, 8}) {
                if ((size_t)lines * 2112 * ring > 200 * 1024) continue;
                const int slabs = 400;
                for (int rep = 0; rep < 2; ++rep) {
                    k<<<148, 160, (size_t)lines * 2112 * ring>>>(x, plane_vec, row_vec, lines <= 8 ? lines : 6, ring, slabs, mode, d);
                    if (lines > 8) k<<<148, 160, (size_t)lines * 2112 * ring>>>(x, plane_vec / 4, row_vec, lines, ring, slabs, mode, d);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
                }
                cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
                const double cps = (double)h / slabs, bpc = lines * 2112.0 / cps;
                printf("%s %d %d %.0f %.1f %.2f\n", mode ? "lane-per-line" : "one-lane", lines, ring, cps, bpc, bpc * 148 * 1.9e9 / 1e12);
            }
    return 0;
}
