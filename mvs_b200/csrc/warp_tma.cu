// Fused warp + variance builder, TMA-staged: the default fast path for fp16 ("C8H") feature maps.
//
// Same contract as warp_c8.cu (every feature map read once from HBM, the B x C x D x H x W variance volume written once,
// no warped volumes / grids / sum volumes: MVSNet/models/mvsnet.py:152-170, module.py:46-87), different data path:
//
//   * a CTA owns one 32 x 8 pixel tile, DCH consecutive depth hypotheses and ONE channel block.  Warp 0 bounds the source
//     footprint of the whole (tile x depth chunk) per view from 8 evaluations -- the 4 tile corners at the chunk's smallest
//     and largest hypothesis: for a fixed depth the sample position is a homography of the pixel (extrema at corners), for
//     a fixed pixel it moves monotonically along the epipolar line (extrema at the depth end points) -- and issues ONE
//     tensor-map TMA (cp.async.bulk.tensor.4d, SASS UTMALDG) per source view that lands the footprint box in shared
//     memory.  Two box shapes per view (64 x 18 for views whose epipolar lines run along x, 40 x 28 along y; one 18 KB
//     slot either way) are encoded on the host; the CTA picks per view.
//   * the TMA zero-fills everything outside the image, which IS grid_sample's zero padding: the bilinear taps become four
//     LDS.128 at (y0 - by) * BW + (x0 - bx) + {0, 1, BW, BW + 1} with the plain weights -- no clamps, no selects;
//   * shared-memory gathers cost 4 wavefronts per 512 B warp request whatever the alignment; the same request through L1
//     (warp_c8.cu) costs ~7 because a misaligned 128 B quarter-warp segment straddles two cache lines (ncu: r1f_warp_c8h_*);
//   * a tap block that is not inside the staged box (footprint larger than the box, degenerate cameras, non-monotonic
//     hypotheses) takes a per-lane fallback through global memory with explicit bounds tests: the box only ever decides
//     WHERE a tap is read from, never its value, so results do not depend on the bounding step.
//
// Tap arithmetic: tap_position<> of warp_c8.cu (the reference's exact op sequence, bit-exact floor(ix), floor(iy)); blend in
// packed fp16 (HFMA2), running sum / sum of squares over views and the variance in packed fp32 -- instruction for
// instruction the BLEND == 2 path of warp_variance_c8_kernel, so both kernels produce identical bits
// (tests/test_gpu_parity.py::test_tma_builder_equals_gather_builder).
#include <cuda.h>          // CUtensorMap + enums only; cuTensorMapEncodeTiled is resolved at run time through cudart

#include "warp_fast.cuh"

namespace mvs {

constexpr int TB_SLOT_PX = 1152;                       // 18 KB per source view
constexpr int TB_W0 = 64, TB_H0 = 18;                  // box shape 0 (1152 px)
constexpr int TB_W1 = 40, TB_H1 = 28;                  // box shape 1 (1120 px)
constexpr int TB_SLOT_BYTES = TB_SLOT_PX * 16;

struct TmaMaps {
    CUtensorMap m[MVS_MAX_SRC][2];
};

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_load_box(void *dst, const CUtensorMap *map, int c1, int c2, int c3, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n"
        :: "r"(smem_addr(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_addr(bar))
        : "memory");
}

__device__ __forceinline__ void mbar_init_(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_(uint64_t *bar, uint32_t bytes)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n"
                 :: "r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}\n"
        :: "r"(smem_addr(bar)), "r"(parity) : "memory");
}

template <int NSRC, bool PL, int DCH, int MINCTAS>
__global__ void __launch_bounds__(256, MINCTAS)
warp_variance_tma_kernel(const uint4 *__restrict__ ref, SrcPtrs srcs, const float *__restrict__ rot,
                         const float *__restrict__ trans, const float *__restrict__ depth, int depth_mode,
                         uint4 *__restrict__ out, int CB, int D, int H, int W, GeomC8 g, int ref_sum_squared,
                         const __grid_constant__ TmaMaps maps)
{
    extern __shared__ __align__(128) uint8_t s_boxes[];          // [NSRC][TB_SLOT_BYTES]
    __shared__ float s_cam[NSRC][12];
    __shared__ float s_lo[8], s_hi[8];
    __shared__ int4 s_box[NSRC];                                  // x = box origin x, y = origin y, z = box width, w = height - 2
    __shared__ __align__(8) uint64_t s_bar;

    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int tile_x0 = blockIdx.y * 32, tile_y0 = blockIdx.z * 8;
    const int x = tile_x0 + threadIdx.x, y = tile_y0 + threadIdx.y;
    const bool valid = x < W && y < H;
    const int dchunks = (D + DCH - 1) / DCH;
    int bi = blockIdx.x;
    const int cb = bi % CB; bi /= CB;
    const int d0 = (bi % dchunks) * DCH;
    const int b = bi / dchunks;
    const int nd = min(DCH, D - d0);

    if (tid < NSRC * 12) {
        const int v = tid / 12, k = tid % 12;
        s_cam[v][k] = k < 9 ? __ldg(rot + ((size_t)b * NSRC + v) * 9 + k) : __ldg(trans + ((size_t)b * NSRC + v) * 3 + (k - 9));
    }
    if (tid == 0) {
        mbar_init_(&s_bar, NSRC);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    const size_t plane = (size_t)H * W;
    const int pix = y * W + x;

    // ---- this thread's hypotheses (kept in registers for the whole chunk) and the chunk's depth range over the tile ----
    float dv[DCH];
    {
        float lo = __int_as_float(0x7f800000), hi = -lo;
#pragma unroll
        for (int k = 0; k < DCH; ++k) {
            dv[k] = 0.f;
            if (k < nd && (valid || depth_mode == MVS_DEPTH_PLANE)) {
                dv[k] = depth_mode == MVS_DEPTH_PLANE ? __ldg(depth + (size_t)b * D + d0 + k)
                                                      : __ldg(depth + ((size_t)b * D + d0 + k) * plane + pix);
                lo = fminf(lo, dv[k]); hi = fmaxf(hi, dv[k]);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (threadIdx.x == 0) { s_lo[threadIdx.y] = lo; s_hi[threadIdx.y] = hi; }
    }
    __syncthreads();

    // ---- warp 0: footprint bounding box per source view, box shape, one TMA per view ----
    if (threadIdx.y == 0) {
        const int lane = threadIdx.x;
        float lo = s_lo[lane & 7], hi = s_hi[lane & 7];
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        const int c = lane & 7;
        const float cx = (float)((c & 1) ? min(tile_x0 + 31, W - 1) : tile_x0);
        const float cy = (float)((c & 2) ? min(tile_y0 + 7, H - 1) : tile_y0);
        const float cd = (c & 4) ? hi : lo;
#pragma unroll
        for (int v0 = 0; v0 < NSRC; v0 += 4) {
            const int v = min(v0 + (lane >> 3), NSRC - 1);
            float q[3];
#pragma unroll
            for (int i = 0; i < 3; ++i)
                q[i] = __fmaf_rn(s_cam[v][i * 3 + 2], 1.0f, __fmaf_rn(s_cam[v][i * 3 + 1], cy, __fmul_rn(s_cam[v][i * 3 + 0], cx)));
            float ix, iy;
            tap_position<PL>(q, s_cam[v], g, cx, cy, cd, ix, iy);
            float mnx = ix, mxx = ix, mny = iy, mxy = iy;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
                mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
                mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o)); mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
            }
            if (c == 0 && v0 + (lane >> 3) < NSRC) {
                // taps of the footprint: columns floor(mnx) .. floor(mxx) + 1, rows floor(mny) .. floor(mxy) + 1
                const float lim = 1048576.0f;
                const int fx0 = __float2int_rd(fminf(fmaxf(mnx, -lim), lim)), fx1 = __float2int_rd(fminf(fmaxf(mxx, -lim), lim)) + 1;
                const int fy0 = __float2int_rd(fminf(fmaxf(mny, -lim), lim)), fy1 = __float2int_rd(fminf(fmaxf(mxy, -lim), lim)) + 1;
                const int wn = fx1 - fx0 + 1, hn = fy1 - fy0 + 1;
                // shape: the one that holds the footprint (with a pixel of slack when possible), else the larger overlap
                int shape;
                if (wn <= TB_W0 && hn <= TB_H0) shape = 0;
                else if (wn <= TB_W1 && hn <= TB_H1) shape = 1;
                else shape = (min(wn, TB_W0) * min(hn, TB_H0) >= min(wn, TB_W1) * min(hn, TB_H1)) ? 0 : 1;
                const int bw = shape ? TB_W1 : TB_W0, bh = shape ? TB_H1 : TB_H0;
                const int bx = fx0 - ((bw - wn) >> 1), by = fy0 - ((bh - hn) >> 1);       // centred (arithmetic shift: also when too large)
                s_box[v] = make_int4(bx, by, bw, bh - 2);
                mbar_expect_tx_(&s_bar, (uint32_t)(bw * bh * 16));
                tma_load_box(s_boxes + (size_t)v * TB_SLOT_BYTES, &maps.m[v][shape], bx, by, b * CB + cb, &s_bar);
            }
        }
    }
    __syncthreads();
    if (!valid) return;

    // ---- per-thread set-up that does not need the staged boxes (overlaps the TMA) ----
    const float fx = (float)x, fy = (float)y;
    float q[NSRC][3];
#pragma unroll
    for (int v = 0; v < NSRC; ++v)
#pragma unroll
        for (int i = 0; i < 3; ++i)
            q[v][i] = __fmaf_rn(s_cam[v][i * 3 + 2], 1.0f, __fmaf_rn(s_cam[v][i * 3 + 1], fy, __fmul_rn(s_cam[v][i * 3 + 0], fx)));
    float2 rf[4];
    {
        const uint4 r4 = __ldg(ref + ((size_t)b * CB + cb) * plane + pix);
        rf[0] = __half22float2(u32_h2(r4.x)); rf[1] = __half22float2(u32_h2(r4.y));
        rf[2] = __half22float2(u32_h2(r4.z)); rf[3] = __half22float2(u32_h2(r4.w));
    }
    const float inv_n = 1.0f / (float)(NSRC + 1);
    const float2 invn2 = make_float2(inv_n, inv_n);
    uint4 *outp = out + (((size_t)b * CB + cb) * D + d0) * plane + pix;

    mbar_wait_(&s_bar, 0);

#pragma unroll 1
    for (int k = 0; k < nd; ++k) {
        float dvk = dv[0];
#pragma unroll
        for (int j = 1; j < DCH; ++j) dvk = (k == j) ? dv[j] : dvk;
        float2 sum[4], sq[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            sq[c] = __fmul2_rn(rf[c], rf[c]);
            sum[c] = ref_sum_squared ? sq[c] : rf[c];
        }
        bool bad = false;
#pragma unroll
        for (int v = 0; v < NSRC; ++v) {
            float ix, iy;
            tap_position<PL>(q[v], s_cam[v], g, fx, fy, dvk, ix, iy);
            bad |= !(fabsf(ix) <= 3.0e38f) || !(fabsf(iy) <= 3.0e38f);
            const float fw = __fsub_rn(ix, floorf(ix)), fe = __fsub_rn(1.0f, fw);
            const float fn = __fsub_rn(iy, floorf(iy)), fs = __fsub_rn(1.0f, fn);
            const int x0 = __float2int_rd(ix), y0 = __float2int_rd(iy);       // saturating; NaN -> 0 (the voxel is `bad` then)
            const uint32_t w00 = h2_u32(__float2half2_rn(__fmul_rn(fs, fe))), w01 = h2_u32(__float2half2_rn(__fmul_rn(fs, fw)));
            const uint32_t w10 = h2_u32(__float2half2_rn(__fmul_rn(fn, fe))), w11 = h2_u32(__float2half2_rn(__fmul_rn(fn, fw)));
            const int4 bx = s_box[v];
            const unsigned lx = (unsigned)x0 - (unsigned)bx.x, ly = (unsigned)y0 - (unsigned)bx.y;
            uint4 t0, t1, t2, t3;
            if (lx <= (unsigned)(bx.z - 2) && ly <= (unsigned)bx.w) {
                const uint4 *p = reinterpret_cast<const uint4 *>(s_boxes + (size_t)v * TB_SLOT_BYTES) + ly * (unsigned)bx.z + lx;
                t0 = p[0]; t1 = p[1]; t2 = p[bx.z]; t3 = p[bx.z + 1];
            } else {
                // tap block outside the staged box: global loads with grid_sample's zero padding
                const uint4 *p = (const uint4 *)srcs.p[v] + ((size_t)b * CB + cb) * plane;
                const bool xi0 = (unsigned)x0 < (unsigned)W, xi1 = (unsigned)x0 + 1u < (unsigned)W;
                const bool yi0 = (unsigned)y0 < (unsigned)H, yi1 = (unsigned)y0 + 1u < (unsigned)H;
                const uint4 z = make_uint4(0, 0, 0, 0);
                const long long o00 = (long long)y0 * W + x0;
                t0 = (xi0 && yi0) ? __ldg(p + o00) : z;
                t1 = (xi1 && yi0) ? __ldg(p + o00 + 1) : z;
                t2 = (xi0 && yi1) ? __ldg(p + o00 + W) : z;
                t3 = (xi1 && yi1) ? __ldg(p + o00 + W + 1) : z;
            }
            const uint32_t *a = &t0.x, *bb = &t1.x, *cc = &t2.x, *e = &t3.x;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                __half2 oh = __hmul2(u32_h2(a[c]), u32_h2(w00));
                oh = __hfma2(u32_h2(bb[c]), u32_h2(w01), oh);
                oh = __hfma2(u32_h2(cc[c]), u32_h2(w10), oh);
                oh = __hfma2(u32_h2(e[c]), u32_h2(w11), oh);
                const float2 o = __half22float2(oh);
                sum[c] = __fadd2_rn(sum[c], o);
                sq[c] = __ffma2_rn(o, o, sq[c]);
            }
        }
        uint32_t o4[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float2 mean = __fmul2_rn(sum[c], invn2);
            const float2 nm2 = __fmul2_rn(make_float2(-mean.x, -mean.y), mean);
            const float2 var = __ffma2_rn(sq[c], invn2, nm2);
            o4[c] = bad ? 0x7fc07fc0u : pack2(var.x, var.y);
        }
        __stcs(outp + (size_t)k * plane, make_uint4(o4[0], o4[1], o4[2], o4[3]));
    }
}

// ---- host side ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled()
{
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        (void)cudaGetLastError();
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// C8H feature map [B][CB][H][W][8 halves] as a rank-4 tensor {8, W, H, B*CB}; box {8, bw, bh, 1}; out-of-image elements read 0
static bool encode_feature_map(CUtensorMap *m, const void *base, int BCB, int H, int W, int bw, int bh)
{
    EncodeTiledFn fn = encode_tiled();
    if (!fn) return false;
    const cuuint64_t dims[4] = {8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)BCB};
    const cuuint64_t strides[3] = {16, (cuuint64_t)W * 16, (cuuint64_t)H * W * 16};
    const cuuint32_t box[4] = {8, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int NSRC, bool PL, int DCH>
static int launch_tma_t(const void *ref, const SrcPtrs &s, const TmaMaps &maps, const float *rot, const float *trans,
                        const float *depth, int depth_mode, void *out, int B, int C, int D, int H, int W, const GeomC8 &g, int rss,
                        cudaStream_t st)
{
    constexpr int MINCTAS = NSRC <= 4 ? 3 : (NSRC <= 6 ? 2 : 1);
    auto kern = warp_variance_tma_kernel<NSRC, PL, DCH, MINCTAS>;
    const size_t smem = (size_t)NSRC * TB_SLOT_BYTES;
    static bool attr_set = false;           // per instantiation; setting it twice is harmless
    if (!attr_set) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return fail(MVS_ERR_CUDA, "mvs_warp_variance_c8_fwd: cannot reserve shared memory for the TMA builder");
        attr_set = true;
    }
    const int CB = C / 8;
    dim3 grid(B * cdiv(D, DCH) * CB, cdiv(W, 32), cdiv(H, 8)), block(32, 8);
    kern<<<grid, block, smem, st>>>((const uint4 *)ref, s, rot, trans, depth, depth_mode, (uint4 *)out, CB, D, H, W, g, rss, maps);
    return check_launch("mvs_warp_variance_c8_fwd(tma)");
}

template <int NSRC>
static int launch_tma_n(const void *ref, const SrcPtrs &s, const TmaMaps &maps, const float *rot, const float *trans,
                        const float *depth, int depth_mode, void *out, int B, int C, int D, int H, int W, int flags, cudaStream_t st)
{
    const GeomC8 g = make_geom_c8(H, W, flags);
    const int rss = (flags & MVS_REF_SUM_SQUARED) ? 1 : 0;
    const bool pl = (flags & MVS_PL_ORDER) != 0;
    if (D > 4) {
        return pl ? launch_tma_t<NSRC, true, 8>(ref, s, maps, rot, trans, depth, depth_mode, out, B, C, D, H, W, g, rss, st)
                  : launch_tma_t<NSRC, false, 8>(ref, s, maps, rot, trans, depth, depth_mode, out, B, C, D, H, W, g, rss, st);
    }
    return pl ? launch_tma_t<NSRC, true, 4>(ref, s, maps, rot, trans, depth, depth_mode, out, B, C, D, H, W, g, rss, st)
              : launch_tma_t<NSRC, false, 4>(ref, s, maps, rot, trans, depth, depth_mode, out, B, C, D, H, W, g, rss, st);
}

// Returns MVS_OK / an error, or 1 when the TMA path cannot be used (no driver entry point, extents the tensor map cannot
// describe): the caller then launches the L1-gather kernel of warp_c8.cu.
int warp_variance_tma(const void *ref, const SrcPtrs &s, int nsrc, const float *rot, const float *trans, const float *depth,
                      int depth_mode, void *out, int B, int C, int D, int H, int W, int flags, cudaStream_t st)
{
    const int CB = C / 8;
    if ((long long)B * cdiv(D, 4) * CB > 2147483647LL || (long long)B * CB > 2147483647LL) return 1;
    TmaMaps maps;
    for (int v = 0; v < nsrc; ++v) {
        if ((reinterpret_cast<uintptr_t>(s.p[v]) & 15) != 0) return 1;
        if (!encode_feature_map(&maps.m[v][0], s.p[v], B * CB, H, W, TB_W0, TB_H0) ||
            !encode_feature_map(&maps.m[v][1], s.p[v], B * CB, H, W, TB_W1, TB_H1))
            return 1;
    }
    for (int v = nsrc; v < MVS_MAX_SRC; ++v) { maps.m[v][0] = maps.m[0][0]; maps.m[v][1] = maps.m[0][1]; }
    switch (nsrc) {
#define CASE(N) case N: return launch_tma_n<N>(ref, s, maps, rot, trans, depth, depth_mode, out, B, C, D, H, W, flags, st);
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
    }
    return 1;
}

}  // namespace mvs
