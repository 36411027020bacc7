// Fused warp + variance builder, TMA-staged and persistent: the default fast path for fp16 ("C8H") feature maps.
//
// Same contract as warp_c8.cu (every feature map read once from HBM, the B x C x D x H x W variance volume written once,
// no warped volumes / grids / sum volumes: MVSNet/models/mvsnet.py:152-170, module.py:46-87), different data path:
//
//   * ONE persistent CTA of 16 warps per SM walks work items (batch, 32 x 16 pixel tile, DCH consecutive hypotheses); per
//     item it loops over the channel blocks.  For every (item, channel block) warp 0 issues one tensor-map TMA
//     (cp.async.bulk.tensor.4d, SASS UTMALDG) per source view that lands the view's source FOOTPRINT BOX in shared memory;
//     the boxes are double-buffered, so the TMA of the next channel block / next item flies while this one is gathered.
//   * the footprint of (tile x depth chunk) is bounded from 8 evaluations per view -- the 4 tile corners at the chunk's
//     smallest and largest hypothesis: for a fixed depth the sample position is a homography of the pixel (extrema at the
//     corners), for a fixed pixel it moves monotonically along the epipolar line (extrema at the depth end points).  The
//     min / max hypothesis of the NEXT item is reduced one item ahead (its loads are issued at the top of the current
//     item), so the bounding never stalls the pipeline.  Two box shapes per view (wide for views whose epipolar lines run
//     along x, tall along y; one slot size) are encoded on the host; the CTA picks per view and item.
//   * the TMA zero-fills everything outside the image, which IS grid_sample's zero padding: the bilinear taps become four
//     LDS.128 at box[(y0 - by) * BW + (x0 - bx) + {0, 1, BW, BW + 1}] with the plain weights -- no clamps, no selects;
//     shared-memory gathers cost 4 wavefronts per 512 B warp request whatever the alignment, the same request through L1
//     (warp_c8.cu) ~7 because a misaligned 128 B quarter-warp segment straddles two cache lines (ncu: r1f_warp_c8h_*);
//   * CACHE mode (more than one channel block): the tap state of every (depth, view) of the item -- box offset + four
//     fp16 weights = 3 registers -- is computed ONCE and reused for all channel blocks (the L1-gather kernel could do this
//     only by keeping every channel block's taps in flight; here the boxes stream through shared memory instead);
//   * a tap block that is not inside the staged box (footprint larger than the box, degenerate cameras, non-monotonic
//     hypotheses) takes a per-lane slow path through global memory with explicit bounds tests: the box only ever decides
//     WHERE a tap is read from, never its value, so results do not depend on the bounding step.
//
// Tap arithmetic: tap_position<> of warp_fast.cuh (the reference's exact op sequence, bit-exact floor(ix), floor(iy));
// weights fp32 products rounded to fp16, blend in packed fp16 (HFMA2), running sum / sum of squares over views and the
// variance in packed fp32 -- value for value the BLEND == 2 path of warp_variance_c8_kernel, so both kernels produce
// identical bits (tests/test_gpu_parity.py::test_tma_builder_equals_gather_builder).
#include <cuda.h>          // CUtensorMap + enums only; cuTensorMapEncodeTiled is resolved at run time through cudart

#include "warp_fast.cuh"

namespace mvs {

#ifndef MVS_TMA_CHAIN
#define MVS_TMA_CHAIN 0      // 1: chain every gather to the previous blend (one tap block in flight per warp)
#endif
constexpr int TMA_MAX_SRC = 6;        // source views the TMA path takes (smem: 2 buffers x NSRC slots); more -> gather kernel
constexpr int TMA_THREADS = 512;
constexpr int TILE_W = 32, TILE_H = 16;

struct TmaMaps {
    CUtensorMap m[TMA_MAX_SRC][2];
};

struct BoxGeom {
    int bw[2], bh[2];       // the two box shapes (pixels)
    int slot_bytes;         // shared-memory bytes reserved per (buffer, view)
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

struct Item {
    int b, x0, y0, d0, nd;
};

__device__ __forceinline__ Item decode_item(int it, int dchunks, int tiles_x, int tiles_y, int D, int DCH)
{
    Item r;
    const int dc = it % dchunks; int t = it / dchunks;
    const int tx = t % tiles_x; t /= tiles_x;
    r.y0 = (t % tiles_y) * TILE_H; r.b = t / tiles_y;
    r.x0 = tx * TILE_W; r.d0 = dc * DCH; r.nd = min(DCH, D - r.d0);
    return r;
}

__device__ __forceinline__ int4 lds_box(const int4 *p)      // one broadcast LDS.128 per use: the boxes stay out of the register file
{
    int4 r;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];\n" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(smem_addr(p)));
    return r;
}

struct TapState {
    uint32_t off;          // 16 B-vector index of the tap block's north-west pixel inside the view's box, ~0u: slow path
    uint32_t wtop, wbot;   // (w_nw, w_ne), (w_sw, w_se) as fp16 pairs
};

template <bool PL, bool AC>
__device__ __forceinline__ TapState make_state(const float q[3], const float *cam, const GeomC8 &g, float fx, float fy, float dv,
                                               const int4 &box, bool &bad)
{
    float ix, iy;
    tap_position<PL, AC ? 1 : 0>(q, cam, g, fx, fy, dv, ix, iy);
    bad |= !(fabsf(ix) <= 3.0e38f) || !(fabsf(iy) <= 3.0e38f);
    const float fw = __fsub_rn(ix, floorf(ix)), fe = __fsub_rn(1.0f, fw);
    const float fn = __fsub_rn(iy, floorf(iy)), fs = __fsub_rn(1.0f, fn);
    const int x0 = __float2int_rd(ix), y0 = __float2int_rd(iy);       // saturating; NaN -> 0 (the voxel is `bad` then)
    TapState st;
    st.wtop = h2_u32(__floats2half2_rn(__fmul_rn(fs, fe), __fmul_rn(fs, fw)));
    st.wbot = h2_u32(__floats2half2_rn(__fmul_rn(fn, fe), __fmul_rn(fn, fw)));
    const unsigned lx = (unsigned)x0 - (unsigned)box.x, ly = (unsigned)y0 - (unsigned)box.y;
    const unsigned bw = (unsigned)box.z & 0xffffu, bhm2 = (unsigned)box.w;
    st.off = (lx <= bw - 2u && ly <= bhm2) ? ly * bw + lx : 0xffffffffu;
    return st;
}

// Slow path of one tap block (never taken when the footprint fits the box): recompute the position from scratch and read
// the four taps from global memory with grid_sample's zero padding.
template <bool PL, bool AC>
__device__ __forceinline__ void slow_taps(uint4 (&t)[4], const float *cam, const GeomC8 &g, float fx, float fy, float dv,
                                       const uint4 *plane_ptr, int H, int W)
{
    float q[3], ix, iy;
#pragma unroll
    for (int i = 0; i < 3; ++i) q[i] = __fmaf_rn(cam[i * 3 + 2], 1.0f, __fmaf_rn(cam[i * 3 + 1], fy, __fmul_rn(cam[i * 3 + 0], fx)));
    tap_position<PL, AC ? 1 : 0>(q, cam, g, fx, fy, dv, ix, iy);
    const int x0 = __float2int_rd(ix), y0 = __float2int_rd(iy);
    const bool xi0 = (unsigned)x0 < (unsigned)W, xi1 = (unsigned)x0 + 1u < (unsigned)W;
    const bool yi0 = (unsigned)y0 < (unsigned)H, yi1 = (unsigned)y0 + 1u < (unsigned)H;
    const uint4 z = make_uint4(0, 0, 0, 0);
    const long long o00 = (long long)y0 * W + x0;
    t[0] = (xi0 && yi0) ? __ldg(plane_ptr + o00) : z;
    t[1] = (xi1 && yi0) ? __ldg(plane_ptr + o00 + 1) : z;
    t[2] = (xi0 && yi1) ? __ldg(plane_ptr + o00 + W) : z;
    t[3] = (xi1 && yi1) ? __ldg(plane_ptr + o00 + W + 1) : z;
}

__device__ __forceinline__ void blend_accumulate(const uint4 (&t)[4], const TapState &st, float2 (&sum)[4], float2 (&sq)[4])
{
    const __half2 wt = u32_h2(st.wtop), wb = u32_h2(st.wbot);
    const __half2 w00 = __low2half2(wt), w01 = __high2half2(wt), w10 = __low2half2(wb), w11 = __high2half2(wb);
    const uint32_t *a = &t[0].x, *b = &t[1].x, *c = &t[2].x, *e = &t[3].x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        __half2 oh = __hmul2(u32_h2(a[k]), w00);
        oh = __hfma2(u32_h2(b[k]), w01, oh);
        oh = __hfma2(u32_h2(c[k]), w10, oh);
        oh = __hfma2(u32_h2(e[k]), w11, oh);
        const float2 o = __half22float2(oh);
        sum[k] = __fadd2_rn(sum[k], o);
        sq[k] = __ffma2_rn(o, o, sq[k]);
    }
}

template <int NSRC, bool PL, bool AC, int DCH, bool CACHE>
__global__ void __launch_bounds__(TMA_THREADS, 1)
warp_variance_tma_kernel(const uint4 *__restrict__ ref, SrcPtrs srcs, const float *__restrict__ rot,
                         const float *__restrict__ trans, const float *__restrict__ depth, int depth_mode,
                         uint4 *__restrict__ out, int B, int CB, int D, int H, int W, GeomC8 g, int ref_sum_squared,
                         BoxGeom bg, int tiles_x, int tiles_y, const __grid_constant__ TmaMaps maps)
{
    extern __shared__ __align__(128) uint8_t s_boxes[];          // [2][NSRC][slot_bytes]
    __shared__ float s_cam[3][NSRC][12];                          // item i uses slot i % 3: written one item ahead while slow
                                                                  // threads may still read the previous item's cameras
    __shared__ float s_lo[2][16], s_hi[2][16];
    __shared__ int4 s_box[2][NSRC];                               // x, y = box origin; z = box width | shape << 16; w = height - 2
    __shared__ __align__(8) uint64_t s_full[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int dchunks = (D + DCH - 1) / DCH;
    const int n_items = B * tiles_y * tiles_x * dchunks;
    const size_t plane = (size_t)H * W;
    const float inv_n = 1.0f / (float)(NSRC + 1);
    const float2 invn2 = make_float2(inv_n, inv_n);

    if (tid == 0) {
        mbar_init_(&s_full[0], NSRC);
        mbar_init_(&s_full[1], NSRC);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }

    // ---- look-ahead: everything item `nx` needs before its first TMA can be issued ----
    float dvn[DCH];
    auto lookahead_loads = [&](const Item &nx, int cpar) {
        if (tid < NSRC * 12) {
            const int v = tid / 12, k = tid % 12;
            s_cam[cpar][v][k] = k < 9 ? __ldg(rot + ((size_t)nx.b * NSRC + v) * 9 + k) : __ldg(trans + ((size_t)nx.b * NSRC + v) * 3 + (k - 9));
        }
        const int x = nx.x0 + lane, y = nx.y0 + warp;
        const bool ok = x < W && y < H;
#pragma unroll
        for (int k = 0; k < DCH; ++k) {
            dvn[k] = 0.f;
            if (k < nx.nd) {
                if (depth_mode == MVS_DEPTH_PLANE) dvn[k] = __ldg(depth + (size_t)nx.b * D + nx.d0 + k);
                else if (ok) dvn[k] = __ldg(depth + ((size_t)nx.b * D + nx.d0 + k) * plane + (size_t)y * W + x);
            }
        }
    };
    auto lookahead_minmax = [&](const Item &nx, int par) {
        const int x = nx.x0 + lane, y = nx.y0 + warp;
        const bool ok = (x < W && y < H) || depth_mode == MVS_DEPTH_PLANE;
        float lo = __int_as_float(0x7f800000), hi = -lo;
#pragma unroll
        for (int k = 0; k < DCH; ++k)
            if (k < nx.nd && ok) { lo = fminf(lo, dvn[k]); hi = fmaxf(hi, dvn[k]); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (lane == 0) { s_lo[par][warp] = lo; s_hi[par][warp] = hi; }
    };
    // warp 0: bounding box + box shape per view of item `nx` (s_lo / s_hi / s_cam of parity `par` are visible)
    auto bound_boxes = [&](const Item &nx, int par, int cpar) {
        float lo = s_lo[par][lane & 15], hi = s_hi[par][lane & 15];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        const int c = lane & 7;
        const float cx = (float)((c & 1) ? min(nx.x0 + TILE_W - 1, W - 1) : nx.x0);
        const float cy = (float)((c & 2) ? min(nx.y0 + TILE_H - 1, H - 1) : nx.y0);
        const float cd = (c & 4) ? hi : lo;
#pragma unroll
        for (int v0 = 0; v0 < NSRC; v0 += 4) {
            const int v = min(v0 + (lane >> 3), NSRC - 1);
            const float *cam = s_cam[cpar][v];
            float q[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) q[i] = __fmaf_rn(cam[i * 3 + 2], 1.0f, __fmaf_rn(cam[i * 3 + 1], cy, __fmul_rn(cam[i * 3 + 0], cx)));
            float ix, iy;
            tap_position<PL, AC ? 1 : 0>(q, cam, g, cx, cy, cd, ix, iy);
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
                int shape;          // the shape that holds the footprint, else the one with the larger overlap
                if (wn <= bg.bw[0] && hn <= bg.bh[0]) shape = 0;
                else if (wn <= bg.bw[1] && hn <= bg.bh[1]) shape = 1;
                else shape = (min(wn, bg.bw[0]) * min(hn, bg.bh[0]) >= min(wn, bg.bw[1]) * min(hn, bg.bh[1])) ? 0 : 1;
                const int bw = bg.bw[shape], bh = bg.bh[shape];
                // centred on the footprint (arithmetic shift: also when the footprint is larger than the box)
                s_box[par][v] = make_int4(fx0 - ((bw - wn) >> 1), fy0 - ((bh - hn) >> 1), bw | (shape << 16), bh - 2);
            }
        }
        __syncwarp();
    };
    // warp 0, lanes < NSRC: one TMA per view for (item, channel block) into buffer `buf`
    auto issue_boxes = [&](const Item &nx, int par, int cb, int buf) {
        if (lane < NSRC) {
            const int4 bx = s_box[par][lane];
            const int shape = bx.z >> 16;
            mbar_expect_tx_(&s_full[buf], (uint32_t)(bg.bw[shape] * bg.bh[shape] * 16));
            tma_load_box(s_boxes + ((size_t)buf * NSRC + lane) * bg.slot_bytes, &maps.m[lane][shape], bx.x, bx.y, nx.b * CB + cb,
                         &s_full[buf]);
        }
    };

    int it = blockIdx.x;
    if (it >= n_items) return;
    Item cur = decode_item(it, dchunks, tiles_x, tiles_y, D, DCH);
    lookahead_loads(cur, 0);
    lookahead_minmax(cur, 0);
    __syncthreads();
    if (warp == 0) {
        bound_boxes(cur, 0, 0);
        issue_boxes(cur, 0, 0, 0);
    }
    int par = 0, cpar = 0;
    uint32_t sub = 0;           // sub-step counter: buffer = sub & 1, barrier phase parity = (sub >> 1) & 1

    for (; it < n_items; it += gridDim.x, par ^= 1, cpar = cpar == 2 ? 0 : cpar + 1) {
        const int cnext = cpar == 2 ? 0 : cpar + 1;
        const float(*cams)[12] = s_cam[cpar];
        const int x = cur.x0 + lane, y = cur.y0 + warp;
        const bool valid = x < W && y < H;
        const int pix = y * W + x;
        const float fx = (float)x, fy = (float)y;
        auto load_depth = [&](int d) -> float {
            return depth_mode == MVS_DEPTH_PLANE ? __ldg(depth + (size_t)cur.b * D + cur.d0 + d)
                                                 : __ldg(depth + ((size_t)cur.b * D + cur.d0 + d) * plane + pix);
        };
        float dv[CACHE ? DCH : 1];
        if (CACHE) {
#pragma unroll
            for (int k = 0; k < DCH; ++k) dv[CACHE ? k : 0] = dvn[k];
        } else {
            dv[0] = valid ? load_depth(0) : 0.f;      // plain mode walks the hypotheses with a one-ahead prefetch
        }
        const int it_next = it + gridDim.x;
        const bool has_next = it_next < n_items;
        Item nxt = cur;
        if (has_next) {
            nxt = decode_item(it_next, dchunks, tiles_x, tiles_y, D, DCH);
            lookahead_loads(nxt, cnext);
        }

        // the first box of this item has landed: the wait also publishes s_box[par] (written before the arrive)
        mbar_wait_(&s_full[sub & 1], (sub >> 1) & 1);
        // The mbarrier (arrive = release after the s_box write, wait = acquire) already orders s_box; the CTA barrier makes the
        // hand-over visible to compute-sanitizer's racecheck as well (it does not model mbarrier ordering of generic stores)
        __syncthreads();

        float q[NSRC][3];
        int bwv[NSRC];
#pragma unroll
        for (int v = 0; v < NSRC; ++v) {
            bwv[v] = lds_box(&s_box[par][v]).z & 0xffff;
#pragma unroll
            for (int i = 0; i < 3; ++i)
                q[v][i] = __fmaf_rn(cams[v][i * 3 + 2], 1.0f, __fmaf_rn(cams[v][i * 3 + 1], fy, __fmul_rn(cams[v][i * 3 + 0], fx)));
        }
        // CACHE: the tap state of every (depth, view) of the item is computed once and reused for all channel blocks: the
        // box offset stays in a register, the four fp16 weights go to a thread-private shared-memory slot (one conflict-free
        // LDS.64 per tap block next to its four LDS.128; 48 state registers would not fit beside the gather's working set)
        uint32_t offs[CACHE ? DCH : 1][NSRC];
        uint2 *const s_w = reinterpret_cast<uint2 *>(s_boxes + (size_t)2 * NSRC * bg.slot_bytes) + tid;     // [DCH * NSRC][512]
        uint32_t badmask = 0;
        if (CACHE && valid) {
#pragma unroll
            for (int d = 0; d < DCH; ++d) {
                if (d < cur.nd) {
                    bool bad = false;
#pragma unroll
                    for (int v = 0; v < NSRC; ++v) {
                        const TapState st = make_state<PL, AC>(q[v], cams[v], g, fx, fy, dv[CACHE ? d : 0], lds_box(&s_box[par][v]), bad);
                        offs[CACHE ? d : 0][v] = st.off;
                        s_w[(d * NSRC + v) * TMA_THREADS] = make_uint2(st.wtop, st.wbot);
                    }
                    badmask |= bad ? (1u << d) : 0u;
                }
            }
        }
        if (has_next) lookahead_minmax(nxt, par ^ 1);

        for (int cb = 0; cb < CB; ++cb, ++sub) {
            __syncthreads();          // everyone is done with the other buffer (previous sub-step); s_lo / s_hi / s_cam of the next item visible
            if (warp == 0) {
                if (cb + 1 < CB) issue_boxes(cur, par, cb + 1, (sub + 1) & 1);
                else if (has_next) { bound_boxes(nxt, par ^ 1, cnext); issue_boxes(nxt, par ^ 1, 0, (sub + 1) & 1); }
            }
            if (!valid) continue;
            const uint4 r4 = __ldg(ref + ((size_t)cur.b * CB + cb) * plane + pix);
            float2 rf[4];
            rf[0] = __half22float2(u32_h2(r4.x)); rf[1] = __half22float2(u32_h2(r4.y));
            rf[2] = __half22float2(u32_h2(r4.z)); rf[3] = __half22float2(u32_h2(r4.w));
            if (cb > 0) mbar_wait_(&s_full[sub & 1], (sub >> 1) & 1);
            const uint8_t *bufp = s_boxes + (size_t)(sub & 1) * NSRC * bg.slot_bytes;
            uint4 *outp = out + (((size_t)cur.b * CB + cb) * D + cur.d0) * plane + pix;
            const uint4 *gplane = nullptr;
            // `dep` is always 0 but opaque to the compiler: it chains every gather to the previous blend, which stops ptxas from
            // hoisting the LDS of all DCH x NSRC unrolled tap blocks to the top (1.7 KB of spills per thread otherwise); the
            // latency of one tap block at a time is covered by the other 15 warps of the SM
            uint32_t dep = 0;
            auto tap_block = [&](const TapState &st, int v, int d, float2 (&sum)[4], float2 (&sq)[4]) {
                uint4 t[4];
                if (st.off != 0xffffffffu) {
                    const uint4 *p = reinterpret_cast<const uint4 *>(bufp + (size_t)v * bg.slot_bytes) + (st.off + dep);
                    t[0] = p[0]; t[1] = p[1]; t[2] = p[bwv[v]]; t[3] = p[bwv[v] + 1];
                } else {
                    slow_taps<PL, AC>(t, cams[v], g, fx, fy, load_depth(d), (const uint4 *)srcs.p[v] + ((size_t)cur.b * CB + cb) * plane, H, W);
                }
                blend_accumulate(t, st, sum, sq);
#if MVS_TMA_CHAIN
                asm volatile("and.b32 %0, %1, 0;\n" : "=r"(dep) : "r"(__float_as_uint(sum[3].y)));
#endif
            };
            auto finish = [&](int d, const float2 (&sum)[4], const float2 (&sq)[4], bool bad) {
                uint32_t o4[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float2 mean = __fmul2_rn(sum[c], invn2);
                    const float2 nm2 = __fmul2_rn(make_float2(-mean.x, -mean.y), mean);
                    const float2 var = __ffma2_rn(sq[c], invn2, nm2);
                    o4[c] = bad ? 0x7fc07fc0u : pack2(var.x, var.y);
                }
                __stcs(outp + (size_t)d * plane, make_uint4(o4[0], o4[1], o4[2], o4[3]));
            };
            (void)gplane;
            if (CACHE) {
#pragma unroll
                for (int d = 0; d < DCH; ++d) {
                    if (d >= cur.nd) break;
                    float2 sum[4], sq[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        sq[c] = __fmul2_rn(rf[c], rf[c]);
                        sum[c] = ref_sum_squared ? sq[c] : rf[c];
                    }
#pragma unroll
                    for (int v = 0; v < NSRC; ++v) {
                        const uint2 w = s_w[(d * NSRC + v) * TMA_THREADS];
                        TapState st;
                        st.off = offs[CACHE ? d : 0][v]; st.wtop = w.x; st.wbot = w.y;
                        tap_block(st, v, d, sum, sq);
                    }
                    finish(d, sum, sq, ((badmask >> d) & 1u) != 0);
                }
            } else {
                float dnext = dv[0];
                if (cb > 0) dnext = load_depth(0);
#pragma unroll 1
                for (int d = 0; d < cur.nd; ++d) {
                    const float dvd = dnext;
                    if (d + 1 < cur.nd) dnext = load_depth(d + 1);
                    float2 sum[4], sq[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        sq[c] = __fmul2_rn(rf[c], rf[c]);
                        sum[c] = ref_sum_squared ? sq[c] : rf[c];
                    }
                    bool bad = false;
#pragma unroll
                    for (int v = 0; v < NSRC; ++v) {
                        const TapState st = make_state<PL, AC>(q[v], cams[v], g, fx, fy, dvd, lds_box(&s_box[par][v]), bad);
                        tap_block(st, v, d, sum, sq);
                    }
                    finish(d, sum, sq, bad);
                }
            }
        }
        cur = nxt;
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

// Box shapes for a 32 x 16 tile walked over DCH hypotheses (~1.7 px of epipolar motion per hypothesis on DTU-like rigs,
// more for the near planes of a plane sweep) within the per-view slot the double buffer leaves: 220 KB / (2 * NSRC).
static BoxGeom box_geom(int nsrc, int dch)
{
    BoxGeom b;
    const int slot_px = (220 * 1024 / (2 * nsrc) / 16) & ~7;
    const int wide_h = dch >= 8 ? 27 : (dch >= 4 ? 24 : 22), tall_w = dch >= 8 ? 44 : 40;
    int want_px = dch >= 8 ? 1760 : (dch >= 4 ? 1280 : 960);
    if (want_px > slot_px) want_px = slot_px;
    b.bw[0] = want_px / wide_h; if (b.bw[0] > 96) b.bw[0] = 96;
    b.bh[0] = wide_h;
    b.bw[1] = tall_w;
    b.bh[1] = want_px / tall_w; if (b.bh[1] > 64) b.bh[1] = 64;
    const int px = b.bw[0] * b.bh[0] > b.bw[1] * b.bh[1] ? b.bw[0] * b.bh[0] : b.bw[1] * b.bh[1];
    b.slot_bytes = ((px * 16) + 127) & ~127;
    return b;
}

template <int NSRC, bool PL, bool AC, int DCH, bool CACHE>
static int launch_tma_t(const void *ref, const SrcPtrs &s, const float *rot, const float *trans, const float *depth, int depth_mode,
                        void *out, int B, int C, int D, int H, int W, int flags, cudaStream_t st)
{
    const int CB = C / 8;
    const BoxGeom bg = box_geom(NSRC, DCH);
    if (bg.bw[0] < TILE_W + 4 || bg.bh[1] < TILE_H + 4) return 1;
    TmaMaps maps;
    for (int v = 0; v < TMA_MAX_SRC; ++v) {
        const int u = v < NSRC ? v : 0;
        if (!encode_feature_map(&maps.m[v][0], s.p[u], B * CB, H, W, bg.bw[0], bg.bh[0]) ||
            !encode_feature_map(&maps.m[v][1], s.p[u], B * CB, H, W, bg.bw[1], bg.bh[1]))
            return 1;
    }
    auto kern = warp_variance_tma_kernel<NSRC, PL, AC, DCH, CACHE>;
    const size_t smem = (size_t)2 * NSRC * bg.slot_bytes + (CACHE ? (size_t)DCH * NSRC * TMA_THREADS * 8 : 0);
    if (smem > 226 * 1024) return 1;
    static bool attr_set = false;           // per instantiation; setting it twice is harmless
    if (!attr_set) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            (void)cudaGetLastError();
            return 1;
        }
        attr_set = true;
    }
    const int tiles_x = cdiv(W, TILE_W), tiles_y = cdiv(H, TILE_H);
    const long long n_items = (long long)B * tiles_x * tiles_y * cdiv(D, DCH);
    if (n_items > 2147483647LL) return 1;
    const int grid = (int)(n_items < sm_count() ? n_items : sm_count());
    const GeomC8 g = make_geom_c8(H, W, flags);
    kern<<<grid, TMA_THREADS, smem, st>>>((const uint4 *)ref, s, rot, trans, depth, depth_mode, (uint4 *)out, B, CB, D, H, W, g,
                                          (flags & MVS_REF_SUM_SQUARED) ? 1 : 0, bg, tiles_x, tiles_y, maps);
    return check_launch("mvs_warp_variance_c8_fwd(tma)");
}

template <int NSRC, bool PL, bool AC>
static int launch_tma_n(const void *ref, const SrcPtrs &s, const float *rot, const float *trans, const float *depth, int depth_mode,
                        void *out, int B, int C, int D, int H, int W, int flags, cudaStream_t st)
{
    // tap states cached across channel blocks when there is more than one (up to 4 source views: 16 states per item)
    constexpr int DCH_PLAIN = NSRC <= 4 ? 8 : 4;
    if (NSRC <= 4 && C > 8)
        return launch_tma_t<(NSRC <= 4 ? NSRC : 1), PL, AC, 4, true>(ref, s, rot, trans, depth, depth_mode, out, B, C, D, H, W, flags, st);
    if (C > 8) return launch_tma_t<NSRC, PL, AC, 4, false>(ref, s, rot, trans, depth, depth_mode, out, B, C, D, H, W, flags, st);
    if (D > DCH_PLAIN / 2)
        return launch_tma_t<NSRC, PL, AC, DCH_PLAIN, false>(ref, s, rot, trans, depth, depth_mode, out, B, C, D, H, W, flags, st);
    return launch_tma_t<NSRC, PL, AC, DCH_PLAIN / 2, false>(ref, s, rot, trans, depth, depth_mode, out, B, C, D, H, W, flags, st);
}

// Returns MVS_OK / an error, or 1 when the TMA path cannot be used (no driver entry point, more than TMA_MAX_SRC views, a
// flag combination it is not instantiated for): the caller then launches the L1-gather kernel of warp_c8.cu.
int warp_variance_tma(const void *ref, const SrcPtrs &s, int nsrc, const float *rot, const float *trans, const float *depth,
                      int depth_mode, void *out, int B, int C, int D, int H, int W, int flags, cudaStream_t st)
{
    if (nsrc > TMA_MAX_SRC || (long long)B * (C / 8) > 2147483647LL) return 1;
    const bool pl = (flags & MVS_PL_ORDER) != 0, ac = (flags & MVS_ALIGN_CORNERS) != 0;
    if (pl != ac) return 1;               // instantiated: the reference's two real combinations (module.py:46 / MVSNet_pl modules.py:25)
    for (int v = 0; v < nsrc; ++v)
        if ((reinterpret_cast<uintptr_t>(s.p[v]) & 15) != 0) return 1;
    switch (nsrc) {
#define CASE(N)                                                                                                                       \
    case N:                                                                                                                           \
        return pl ? launch_tma_n<N, true, true>(ref, s, rot, trans, depth, depth_mode, out, B, C, D, H, W, flags, st)                 \
                  : launch_tma_n<N, false, false>(ref, s, rot, trans, depth, depth_mode, out, B, C, D, H, W, flags, st);
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6)
#undef CASE
    }
    return 1;
}

}  // namespace mvs
