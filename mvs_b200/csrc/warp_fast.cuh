// Tap arithmetic and small packed-math helpers shared by the two fast builders: warp_c8.cu (L1-gather kernel, bf16 or
// fp16 features) and warp_tma.cu (TMA-staged kernel, fp16 features).  ONE definition of the sample position, so both
// kernels -- and the tap probe mvs_warp_taps(MVS_FAST_COORDS) -- see identical floor(ix), floor(iy).
#pragma once
#include "warp_common.cuh"

namespace mvs {

struct GeomC8 {
    float r_hw, r_hh;           // correctly rounded reciprocals of half_wm1 / half_hm1
    float half_wm1, half_hm1, sx, sy;
    int align_corners, pl_order, W, H;
};

struct TapC8 {
    float w[4];           // nw, ne, sw, se (0 for taps outside the image; NaN when the position is non-finite)
    int off[2];           // 16 B-vector index of the (clamped) north and south rows' west tap
    int dx;               // +1 when the east column is a distinct valid column, else 0
    float ix, iy;         // probe only (dead in the builder): sample position, integer taps, in-bounds mask
    int x0, y0;
    unsigned mask;
};

// a / b for b != 0 in the normal range: rcp + one Newton step + residual correction == IEEE RN quotient
__device__ __forceinline__ float rcp_refined(float b)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
    const float e = __fmaf_rn(-b, r, 1.0f);
    return __fmaf_rn(r, e, r);
}
__device__ __forceinline__ float div_by_rcp(float a, float b, float r)
{
    const float q = __fmul_rn(a, r);
    const float rem = __fmaf_rn(-b, q, a);
    return __fmaf_rn(rem, r, q);
}

// Sample position (ix, iy) of ref pixel (x, y) at `depth` in the source image: the reference's exact op
// sequence (warp_common.cuh contract), shared by the builder and the tap probe.
// AC: -1 = align_corners read from g at run time, 0 / 1 = compile-time
template <bool PL, int AC = -1>
__device__ __forceinline__ void tap_position(const float q[3], const float rt[12], const GeomC8 &g, float x, float y,
                                             float depth, float &ix, float &iy)
{
    float P0, P1, P2;
    if (PL) {
        const float g0 = __fmul_rn(x, depth), g1 = __fmul_rn(y, depth), g2 = depth;
        P0 = __fadd_rn(__fmaf_rn(rt[2], g2, __fmaf_rn(rt[1], g1, __fmul_rn(rt[0], g0))), rt[9]);
        P1 = __fadd_rn(__fmaf_rn(rt[5], g2, __fmaf_rn(rt[4], g1, __fmul_rn(rt[3], g0))), rt[10]);
        P2 = __fadd_rn(__fmaf_rn(rt[8], g2, __fmaf_rn(rt[7], g1, __fmul_rn(rt[6], g0))), rt[11]);
    } else {
        P0 = __fadd_rn(__fmul_rn(q[0], depth), rt[9]);
        P1 = __fadd_rn(__fmul_rn(q[1], depth), rt[10]);
        P2 = __fadd_rn(__fmul_rn(q[2], depth), rt[11]);
    }
    const float r = rcp_refined(P2);
    const float u = div_by_rcp(P0, P2, r), v = div_by_rcp(P1, P2, r);
    const float gx = __fsub_rn(div_by_rcp(u, g.half_wm1, g.r_hw), 1.0f);
    const float gy = __fsub_rn(div_by_rcp(v, g.half_hm1, g.r_hh), 1.0f);
    if (AC < 0 ? (g.align_corners != 0) : (AC != 0)) {
        ix = __fmul_rn(__fadd_rn(gx, 1.0f), g.sx);
        iy = __fmul_rn(__fadd_rn(gy, 1.0f), g.sy);
    } else {
        ix = __fmaf_rn(__fadd_rn(gx, 1.0f), g.sx, -0.5f);
        iy = __fmaf_rn(__fadd_rn(gy, 1.0f), g.sy, -0.5f);
    }
}

// Builder-side tap: the 2x2 source block the thread LOADS is (yc, xc) .. (yc+1, xc+1) with
// xc = clamp(x0, 0, W-2), yc = clamp(y0, 0, H-2) -- always inside the image, east = west + 16 B and
// south = north + W*16 B, so one address serves four loads.  The bilinear weights are attached to the
// loaded pixels: when the clamp moved the block (x0 = -1 or W-1, likewise y) the one tap that is still
// inside the image lands on the other column / row of the block and every out-of-image tap gets weight 0,
// which is exactly grid_sample's zero padding.  Products are formed as (row weight) * (column weight), the
// same two factors as the reference's (y1-iy)*(x1-ix) etc., so the weights are bit-identical.
struct TapV {
    float w00, w01, w10, w11;   // (north,west) (north,east) (south,west) (south,east) of the loaded block
    int off;                    // 16 B-vector index of the block's north-west pixel
};

template <bool PL>
__device__ __forceinline__ bool make_tap_v(const float q[3], const float rt[12], const GeomC8 &g, float x, float y,
                                           float depth, TapV &t)
{
    float ix, iy;
    tap_position<PL>(q, rt, g, x, y, depth, ix, iy);
    const float x0f = floorf(ix), y0f = floorf(iy);
    const float fw = __fsub_rn(ix, x0f), fe = __fsub_rn(1.0f, fw);
    const float fn = __fsub_rn(iy, y0f), fs = __fsub_rn(1.0f, fn);
    // cvt.rmi.s32.f32 saturates (and maps NaN to 0); every comparison below is written so that saturated
    // values behave like "far outside"
    const int x0 = __float2int_rd(ix), y0 = __float2int_rd(iy);
    const int xc = min(max(x0, 0), g.W - 2), yc = min(max(y0, 0), g.H - 2);
    const float wl = x0 == xc ? fe : ((unsigned)x0 + 1u == (unsigned)xc ? fw : 0.f);     // second case: x0 = -1
    const float wr = x0 == xc ? fw : ((unsigned)x0 == (unsigned)xc + 1u ? fe : 0.f);     // second case: x0 = W-1
    const float wt = y0 == yc ? fs : ((unsigned)y0 + 1u == (unsigned)yc ? fn : 0.f);
    const float wb = y0 == yc ? fn : ((unsigned)y0 == (unsigned)yc + 1u ? fs : 0.f);
    t.w00 = __fmul_rn(wt, wl); t.w01 = __fmul_rn(wt, wr);
    t.w10 = __fmul_rn(wb, wl); t.w11 = __fmul_rn(wb, wr);
    t.off = yc * g.W + xc;
    // non-finite sample position: the reference's CPU path yields NaN (0 * NaN weights) -- reported to the caller
    return !(fabsf(ix) <= 3.0e38f) || !(fabsf(iy) <= 3.0e38f);
}

template <bool PL>
__device__ __forceinline__ TapC8 make_tap_c8(const float q[3], const float rt[12], const GeomC8 &g, float x, float y,
                                             float depth)
{
    float ix, iy;
    tap_position<PL>(q, rt, g, x, y, depth, ix, iy);
    const float x0f = floorf(ix), y0f = floorf(iy);
    const float fw = __fsub_rn(ix, x0f), fe = __fsub_rn(1.0f, fw);
    const float fn = __fsub_rn(iy, y0f), fs = __fsub_rn(1.0f, fn);
    // integer tap coordinates; positions far outside (or non-finite) saturate and fail every range test
    const int x0 = __float2int_rd(fminf(fmaxf(ix, -4.0f), (float)g.W + 2.0f));
    const int y0 = __float2int_rd(fminf(fmaxf(iy, -4.0f), (float)g.H + 2.0f));
    const bool finite = (fabsf(ix) <= 3.0e38f) && (fabsf(iy) <= 3.0e38f);
    const bool xin0 = (unsigned)x0 < (unsigned)g.W, xin1 = (unsigned)(x0 + 1) < (unsigned)g.W;
    const bool yin0 = (unsigned)y0 < (unsigned)g.H, yin1 = (unsigned)(y0 + 1) < (unsigned)g.H;
    TapC8 t;
    const float nanv = __int_as_float(0x7fc00000);
    t.w[0] = finite ? ((xin0 && yin0) ? __fmul_rn(fs, fe) : 0.f) : nanv;
    t.w[1] = finite ? ((xin1 && yin0) ? __fmul_rn(fs, fw) : 0.f) : nanv;
    t.w[2] = finite ? ((xin0 && yin1) ? __fmul_rn(fn, fe) : 0.f) : nanv;
    t.w[3] = finite ? ((xin1 && yin1) ? __fmul_rn(fn, fw) : 0.f) : nanv;
    // clamped addresses: a tap with weight 0 may read any valid pixel
    const int xc = min(max(x0, 0), g.W - 1);
    const int yn = min(max(y0, 0), g.H - 1), ys = min(max(y0 + 1, 0), g.H - 1);
    t.off[0] = yn * g.W + xc;
    t.off[1] = ys * g.W + xc;
    t.dx = (x0 >= 0 && x0 + 1 < g.W) ? 1 : 0;       // east tap = west + 1 when both columns are inside
    // (x0 = -1: the west tap is outside and the east tap is column 0 = xc, so dx = 0 is right)
    t.ix = ix; t.iy = iy; t.x0 = x0; t.y0 = y0;
    t.mask = finite ? ((unsigned)(xin0 && yin0) | ((unsigned)(xin1 && yin0) << 1) | ((unsigned)(xin0 && yin1) << 2) |
                       ((unsigned)(xin1 && yin1) << 3)) : 0u;
    return t;
}

__device__ __forceinline__ void unpack8(const uint4 &r, float f[8])
{
    f[0] = __uint_as_float(r.x << 16); f[1] = __uint_as_float(r.x & 0xffff0000u);
    f[2] = __uint_as_float(r.y << 16); f[3] = __uint_as_float(r.y & 0xffff0000u);
    f[4] = __uint_as_float(r.z << 16); f[5] = __uint_as_float(r.z & 0xffff0000u);
    f[6] = __uint_as_float(r.w << 16); f[7] = __uint_as_float(r.w & 0xffff0000u);
}

__device__ __forceinline__ uint32_t pack2(float a, float b)
{
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&v);
}

__device__ __forceinline__ __nv_bfloat162 as_bf2(uint32_t u) { return *reinterpret_cast<__nv_bfloat162 *>(&u); }
__device__ __forceinline__ uint32_t as_u32(__nv_bfloat162 v) { return *reinterpret_cast<uint32_t *>(&v); }

__device__ __forceinline__ __half2 u32_h2(uint32_t u) { return *reinterpret_cast<__half2 *>(&u); }
__device__ __forceinline__ uint32_t h2_u32(__half2 v) { return *reinterpret_cast<uint32_t *>(&v); }

__device__ __forceinline__ float2 bf2_to_f2(uint32_t u)
{
    return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

inline GeomC8 make_geom_c8(int H, int W, int flags)
{
    const WarpGeom w = make_geom(H, W, flags);
    GeomC8 g;
    g.half_wm1 = w.half_wm1; g.half_hm1 = w.half_hm1; g.sx = w.sx; g.sy = w.sy;
    g.r_hw = (float)(1.0 / (double)w.half_wm1);
    g.r_hh = (float)(1.0 / (double)w.half_hm1);
    g.align_corners = w.align_corners; g.pl_order = w.pl_order; g.W = W; g.H = H;
    return g;
}

}  // namespace mvs
