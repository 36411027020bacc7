// Bilinear tap set-up of the homography warp: ONE definition shared by the strict fp32 kernels,
// the fast C8/bf16 kernel and the tap-index probe, so that every path samples identical taps.
//
// Arithmetic contract (restated on the CPU in oracle/mvs_oracle.c mvso_tap):
// every operation is an explicit round-to-nearest intrinsic, so nvcc's -fmad contraction can never
// fuse `rot_xyz*d + t` (two roundings in the reference: MVSNet/models/module.py:74-75) and the
// divisions are IEEE (module.py:76-79).  FMAs appear exactly where the reference's CPU path has
// them: the K=3 matmul chain (module.py:73) and ATen's unnormalize (x+1)*(size/2)-0.5.
#pragma once
#include "common.cuh"

namespace mvs {

struct Cam {          // proj = src_proj @ inverse(ref_proj); r = proj[:3,:3] row-major, t = proj[:3,3]
    float r[9];
    float t[3];
};

struct Tap {
    float ix, iy;     // sample position in source pixels
    float x0, y0;     // floor(ix), floor(iy), kept in float like ATen does
    float w_nw, w_ne, w_sw, w_se;
    unsigned mask;    // bit0 nw, bit1 ne, bit2 sw, bit3 se: tap inside the source image
};

struct WarpGeom {     // per-launch constants of the normalise / unnormalise steps
    float half_wm1, half_hm1;   // fp32((W-1)/2), fp32((H-1)/2): divisors of module.py:78-79
    float sx, sy;               // align_corners ? (size-1)/2 : size/2
    float fw, fh;               // (float)W, (float)H
    int align_corners, pl_order;
};

inline WarpGeom make_geom(int H, int W, int flags)
{
    WarpGeom g;
    g.half_wm1 = (float)((double)(W - 1) / 2.0);
    g.half_hm1 = (float)((double)(H - 1) / 2.0);
    g.align_corners = (flags & MVS_ALIGN_CORNERS) ? 1 : 0;
    g.pl_order = (flags & MVS_PL_ORDER) ? 1 : 0;
    g.sx = g.align_corners ? (float)(W - 1) / 2.0f : (float)W / 2.0f;
    g.sy = g.align_corners ? (float)(H - 1) / 2.0f : (float)H / 2.0f;
    g.fw = (float)W;
    g.fh = (float)H;
    return g;
}

// q_i = rot_xyz for pixel (x, y): fma(r2, 1, fma(r1, y, r0*x))  -- depth independent, hoistable.
__device__ __forceinline__ void rot_pixel(const Cam &c, float x, float y, float q[3])
{
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float v = __fmul_rn(c.r[i * 3 + 0], x);
        v = __fmaf_rn(c.r[i * 3 + 1], y, v);
        q[i] = __fmaf_rn(c.r[i * 3 + 2], 1.0f, v);
    }
}

__device__ __forceinline__ Tap make_tap(const Cam &c, const WarpGeom &g, const float q[3], float x,
                                        float y, float depth)
{
    float P[3];
    if (g.pl_order) {
        // MVSNet_pl/models/modules.py:46-50: (x,y,1)*d first, then the K=3 FMA chain, then + T
        const float g0 = __fmul_rn(x, depth), g1 = __fmul_rn(y, depth), g2 = __fmul_rn(1.0f, depth);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            float v = __fmul_rn(c.r[i * 3 + 0], g0);
            v = __fmaf_rn(c.r[i * 3 + 1], g1, v);
            v = __fmaf_rn(c.r[i * 3 + 2], g2, v);
            P[i] = __fadd_rn(v, c.t[i]);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) P[i] = __fadd_rn(__fmul_rn(q[i], depth), c.t[i]);
    }
    const float u = __fdiv_rn(P[0], P[2]);
    const float v = __fdiv_rn(P[1], P[2]);
    const float gx = __fsub_rn(__fdiv_rn(u, g.half_wm1), 1.0f);
    const float gy = __fsub_rn(__fdiv_rn(v, g.half_hm1), 1.0f);
    Tap t;
    if (g.align_corners) {
        t.ix = __fmul_rn(__fadd_rn(gx, 1.0f), g.sx);
        t.iy = __fmul_rn(__fadd_rn(gy, 1.0f), g.sy);
    } else {
        t.ix = __fmaf_rn(__fadd_rn(gx, 1.0f), g.sx, -0.5f);
        t.iy = __fmaf_rn(__fadd_rn(gy, 1.0f), g.sy, -0.5f);
    }
    t.x0 = floorf(t.ix);
    t.y0 = floorf(t.iy);
    const float w = __fsub_rn(t.ix, t.x0), e = __fsub_rn(1.0f, w);
    const float n = __fsub_rn(t.iy, t.y0), s = __fsub_rn(1.0f, n);
    t.w_nw = __fmul_rn(s, e);
    t.w_ne = __fmul_rn(s, w);
    t.w_sw = __fmul_rn(n, e);
    t.w_se = __fmul_rn(n, w);
    const float x1 = __fadd_rn(t.x0, 1.0f), y1 = __fadd_rn(t.y0, 1.0f);
    const bool xin0 = (t.x0 > -1.0f) && (t.x0 < g.fw), xin1 = (x1 > -1.0f) && (x1 < g.fw);
    const bool yin0 = (t.y0 > -1.0f) && (t.y0 < g.fh), yin1 = (y1 > -1.0f) && (y1 < g.fh);
    t.mask = (unsigned)(xin0 && yin0) | ((unsigned)(xin1 && yin0) << 1) | ((unsigned)(xin0 && yin1) << 2) |
             ((unsigned)(xin1 && yin1) << 3);
    return t;
}

// A tap whose four neighbours are all outside contributes exactly 0 -- unless the coordinates are
// non-finite, in which case the reference's CPU path yields NaN (0 * NaN); keep that behaviour.
__device__ __forceinline__ bool tap_is_zero(const Tap &t)
{
    return t.mask == 0u && (fabsf(t.ix) <= 3.0e38f) && (fabsf(t.iy) <= 3.0e38f);
}

__device__ __forceinline__ void load_cam(Cam &c, const float *rot, const float *trans)
{
#pragma unroll
    for (int i = 0; i < 9; ++i) c.r[i] = __ldg(rot + i);
#pragma unroll
    for (int i = 0; i < 3; ++i) c.t[i] = __ldg(trans + i);
}

}  // namespace mvs
