// Strict fp32 warp kernels: reference layouts (NCHW in, NCDHW out), bit-exact with the reference's
// CPU path.  mvs_warp_fwd serves the drop-in homo_warping()/homo_warp() callables; the fused
// mvs_warp_variance_fwd builds the variance volume without materialising the N-1 warped volumes
// (MVSNet/models/mvsnet.py:152-170 et al.).  Thread <-> one (x, y, d) voxel, lanes along x so that
// every per-channel plane access is one coalesced 128 B row segment.
#include "warp_common.cuh"

namespace mvs {

__device__ __forceinline__ float depth_at(const float *depth, int mode, int b, int d, int y, int x, int D,
                                          int H, int W)
{
    return mode == MVS_DEPTH_PLANE ? __ldg(depth + (size_t)b * D + d)
                                   : __ldg(depth + (((size_t)b * D + d) * H + y) * W + x);
}

struct TapAddr {          // a Tap reduced to what the channel loop needs
    float w_nw, w_ne, w_sw, w_se;
    int off;              // y0*W + x0 (may be negative; only dereferenced under the mask)
    unsigned mask;
};

__device__ __forceinline__ TapAddr reduce_tap(const Tap &t, int W)
{
    TapAddr a;
    a.w_nw = t.w_nw; a.w_ne = t.w_ne; a.w_sw = t.w_sw; a.w_se = t.w_se;
    a.mask = t.mask;
    // with any tap in bounds x0 in [-1, W-1], y0 in [-1, H-1]: the int conversion is exact
    a.off = t.mask ? (int)t.y0 * W + (int)t.x0 : 0;
    return a;
}

// out = fma(v_se,w_se, fma(v_sw,w_sw, fma(v_ne,w_ne, v_nw*w_nw))); masked taps read as 0 but still
// multiply their weight (NaN propagation identical to ATen's CPU kernel).
__device__ __forceinline__ float blend(const float *__restrict__ plane, const TapAddr &a, int W)
{
    const float *p = plane + a.off;
    const float v_nw = (a.mask & 1u) ? __ldg(p) : 0.0f;
    const float v_ne = (a.mask & 2u) ? __ldg(p + 1) : 0.0f;
    const float v_sw = (a.mask & 4u) ? __ldg(p + W) : 0.0f;
    const float v_se = (a.mask & 8u) ? __ldg(p + W + 1) : 0.0f;
    float o = __fmul_rn(v_nw, a.w_nw);
    o = __fmaf_rn(v_ne, a.w_ne, o);
    o = __fmaf_rn(v_sw, a.w_sw, o);
    o = __fmaf_rn(v_se, a.w_se, o);
    return o;
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
warp_fwd_kernel(const float *__restrict__ src, const float *__restrict__ rot, const float *__restrict__ trans,
                const float *__restrict__ depth, int depth_mode, float *__restrict__ out, int B, int C, int D,
                int H, int W, WarpGeom g)
{
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int b = blockIdx.z / D, d = blockIdx.z % D;
    if (x >= W || y >= H) return;
    Cam cam;
    load_cam(cam, rot + b * 9, trans + b * 3);
    float q[3];
    rot_pixel(cam, (float)x, (float)y, q);
    const Tap t = make_tap(cam, g, q, (float)x, (float)y, depth_at(depth, depth_mode, b, d, y, x, D, H, W));
    const TapAddr a = reduce_tap(t, W);
    const size_t plane = (size_t)H * W;
    const float *sp = src + (size_t)b * C * plane;
    float *op = out + (((size_t)b * C * D + d) * H + y) * W + x;
#pragma unroll 4
    for (int c = 0; c < C; ++c) op[(size_t)c * D * plane] = blend(sp + c * plane, a, W);
}

__global__ void __launch_bounds__(256)
warp_taps_kernel(const float *__restrict__ rot, const float *__restrict__ trans, const float *__restrict__ depth,
                 int depth_mode, int32_t *__restrict__ x0, int32_t *__restrict__ y0, uint8_t *__restrict__ mask,
                 float *__restrict__ ixy, int B, int D, int H, int W, WarpGeom g)
{
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int b = blockIdx.z / D, d = blockIdx.z % D;
    if (x >= W || y >= H) return;
    Cam cam;
    load_cam(cam, rot + b * 9, trans + b * 3);
    float q[3];
    rot_pixel(cam, (float)x, (float)y, q);
    const Tap t = make_tap(cam, g, q, (float)x, (float)y, depth_at(depth, depth_mode, b, d, y, x, D, H, W));
    const size_t o = (((size_t)b * D + d) * H + y) * W + x;
    auto sat = [](float f) -> int32_t {
        if (!(fabsf(f) <= 3.0e38f)) return INT32_MIN;
        if (f > 1073741824.0f) return 1073741824;
        if (f < -1073741824.0f) return -1073741824;
        return (int32_t)f;
    };
    x0[o] = sat(t.x0);
    y0[o] = sat(t.y0);
    mask[o] = (uint8_t)t.mask;
    if (ixy) { ixy[2 * o] = t.ix; ixy[2 * o + 1] = t.iy; }
}

// Backward of warp_fwd w.r.t. src_fea: the transpose of the bilinear gather.
__global__ void __launch_bounds__(256)
warp_bwd_kernel(const float *__restrict__ gout, const float *__restrict__ rot, const float *__restrict__ trans,
                const float *__restrict__ depth, int depth_mode, float *__restrict__ gsrc, int B, int C, int D, int H,
                int W, WarpGeom g)
{
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int b = blockIdx.z / D, d = blockIdx.z % D;
    if (x >= W || y >= H) return;
    Cam cam;
    load_cam(cam, rot + b * 9, trans + b * 3);
    float q[3];
    rot_pixel(cam, (float)x, (float)y, q);
    const TapAddr a = reduce_tap(make_tap(cam, g, q, (float)x, (float)y, depth_at(depth, depth_mode, b, d, y, x, D, H, W)), W);
    if (a.mask == 0u) return;
    const size_t plane = (size_t)H * W;
    const float *gp = gout + (((size_t)b * C * D + d) * H + y) * W + x;
    for (int c = 0; c < C; ++c) {
        const float go = __ldg(gp + (size_t)c * D * plane);
        float *p = gsrc + ((size_t)b * C + c) * plane + a.off;
        if (a.mask & 1u) atomicAdd(p, go * a.w_nw);
        if (a.mask & 2u) atomicAdd(p + 1, go * a.w_ne);
        if (a.mask & 4u) atomicAdd(p + W, go * a.w_sw);
        if (a.mask & 8u) atomicAdd(p + W + 1, go * a.w_se);
    }
}

// Fused builder.  sum = ref (+= w_i), sq = ref*ref (+= w_i*w_i), var = sq/N - (sum/N)^2, each op
// individually rounded and views visited left to right exactly like the reference's in-place
// tensor ops (mvsnet.py:152-170) => bit-exact volume.
template <int NSRC>
__global__ void __launch_bounds__(256)
warp_variance_kernel(const float *__restrict__ ref, SrcPtrs srcs, const float *__restrict__ rot,
                     const float *__restrict__ trans, const float *__restrict__ depth, int depth_mode,
                     float *__restrict__ out, int B, int C, int D, int H, int W, WarpGeom g, int ref_sum_squared)
{
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int b = blockIdx.z / D, d = blockIdx.z % D;
    if (x >= W || y >= H) return;
    const float dv = depth_at(depth, depth_mode, b, d, y, x, D, H, W);
    TapAddr taps[NSRC];
#pragma unroll
    for (int v = 0; v < NSRC; ++v) {
        Cam cam;
        load_cam(cam, rot + ((size_t)b * NSRC + v) * 9, trans + ((size_t)b * NSRC + v) * 3);
        float q[3];
        rot_pixel(cam, (float)x, (float)y, q);
        taps[v] = reduce_tap(make_tap(cam, g, q, (float)x, (float)y, dv), W);
    }
    const size_t plane = (size_t)H * W;
    const float nviews = (float)(NSRC + 1);
    const float *rp = ref + (size_t)b * C * plane + (size_t)y * W + x;
    float *op = out + (((size_t)b * C * D + d) * H + y) * W + x;
#pragma unroll 2
    for (int c = 0; c < C; ++c) {
        const float r = __ldg(rp + c * plane);
        const float r2 = __fmul_rn(r, r);
        float sum = ref_sum_squared ? r2 : r;
        float sq = r2;
#pragma unroll
        for (int v = 0; v < NSRC; ++v) {
            const float wv = blend((const float *)srcs.p[v] + ((size_t)b * C + c) * plane, taps[v], W);
            sum = __fadd_rn(sum, wv);
            sq = __fadd_rn(sq, __fmul_rn(wv, wv));
        }
        const float mean = __fdiv_rn(sum, nviews);
        op[(size_t)c * D * plane] = __fsub_rn(__fdiv_rn(sq, nviews), __fmul_rn(mean, mean));
    }
}

// Backward of the builder w.r.t. the feature maps.  var = sq/N - (sum/N)^2 with
// sum = s0(ref) + sum_i V_i, sq = ref^2 + sum_i V_i^2:
//   dvar/dV_i = (2/N) (V_i - mean);  dvar/dref = (2/N)(ref - mean * ds0/dref'), see below.
// The warped values are recomputed (never stored); each tap scatters g*dvar/dV_i*w_tap with a
// red.global.add.  Grid is built under no_grad in the reference (module.py:62) => no gradient to
// cameras / hypotheses here.
template <int NSRC>
__global__ void __launch_bounds__(256)
warp_variance_bwd_kernel(const float *__restrict__ gout, const float *__restrict__ ref, SrcPtrs srcs,
                         const float *__restrict__ rot, const float *__restrict__ trans,
                         const float *__restrict__ depth, int depth_mode, float *__restrict__ gref, DstPtrs gsrcs,
                         int B, int C, int D, int H, int W, WarpGeom g, int ref_sum_squared)
{
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int b = blockIdx.z;
    if (x >= W || y >= H) return;
    const size_t plane = (size_t)H * W;
    const float inv_n = 1.0f / (float)(NSRC + 1);
    Cam cams[NSRC];
    float q[NSRC][3];
#pragma unroll
    for (int v = 0; v < NSRC; ++v) {
        load_cam(cams[v], rot + ((size_t)b * NSRC + v) * 9, trans + ((size_t)b * NSRC + v) * 3);
        rot_pixel(cams[v], (float)x, (float)y, q[v]);
    }
    for (int d = 0; d < D; ++d) {
        const float dv = depth_at(depth, depth_mode, b, d, y, x, D, H, W);
        TapAddr taps[NSRC];
#pragma unroll
        for (int v = 0; v < NSRC; ++v) taps[v] = reduce_tap(make_tap(cams[v], g, q[v], (float)x, (float)y, dv), W);
        for (int c = 0; c < C; ++c) {
            const float go = __ldg(gout + ((((size_t)b * C + c) * D + d) * H + y) * W + x);
            const float r = __ldg(ref + ((size_t)b * C + c) * plane + (size_t)y * W + x);
            float vals[NSRC];
            float sum = ref_sum_squared ? r * r : r;
#pragma unroll
            for (int v = 0; v < NSRC; ++v) {
                vals[v] = blend((const float *)srcs.p[v] + ((size_t)b * C + c) * plane, taps[v], W);
                sum += vals[v];
            }
            const float mean = sum * inv_n;
            // d sq / d ref = 2 ref ; d sum / d ref = 1 (or 2 ref under the CVP aliasing quirk)
            const float dsum_dref = ref_sum_squared ? 2.0f * r : 1.0f;
            const float gr = go * inv_n * (2.0f * r - 2.0f * mean * dsum_dref);
            if (gref) atomicAdd(gref + ((size_t)b * C + c) * plane + (size_t)y * W + x, gr);
#pragma unroll
            for (int v = 0; v < NSRC; ++v) {
                float *gp = (float *)gsrcs.p[v];
                if (!gp || taps[v].mask == 0u) continue;
                const float gv = go * 2.0f * inv_n * (vals[v] - mean);
                float *p = gp + ((size_t)b * C + c) * plane + taps[v].off;
                if (taps[v].mask & 1u) atomicAdd(p, gv * taps[v].w_nw);
                if (taps[v].mask & 2u) atomicAdd(p + 1, gv * taps[v].w_ne);
                if (taps[v].mask & 4u) atomicAdd(p + W, gv * taps[v].w_sw);
                if (taps[v].mask & 8u) atomicAdd(p + W + 1, gv * taps[v].w_se);
            }
        }
    }
}

int warp_taps_fast(const float *rot, const float *trans, const float *depth, int depth_mode, int32_t *x0, int32_t *y0,
                   uint8_t *mask, float *ixy, int B, int D, int H, int W, int flags, cudaStream_t st);   // warp_c8.cu

static int check_dims(int B, int C, int D, int H, int W)
{
    MVS_REQUIRE(B > 0 && C > 0 && D > 0 && H > 0 && W > 0, "extents must be positive");
    MVS_REQUIRE((long long)B * D <= 65535, "B*D exceeds the grid.z limit (65535)");
    MVS_REQUIRE((long long)H * W < (1ll << 30), "H*W too large");
    return MVS_OK;
}

}  // namespace mvs

using namespace mvs;

extern "C" int mvs_warp_fwd(const float *src_fea, const float *rot, const float *trans, const float *depth,
                            int depth_mode, float *out, int B, int C, int D, int H, int W, int flags, void *stream)
{
    if (B == 0 || C == 0 || D == 0 || H == 0 || W == 0) return MVS_OK;   // empty input: nothing to do
    if (int e = check_dims(B, C, D, H, W)) return e;
    MVS_REQUIRE(src_fea && rot && trans && depth && out, "null pointer");
    MVS_REQUIRE(depth_mode == MVS_DEPTH_PLANE || depth_mode == MVS_DEPTH_PIXEL, "bad depth_mode");
    dim3 grid(cdiv(W, 32), cdiv(H, 8), B * D), block(32, 8);
    warp_fwd_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(src_fea, rot, trans, depth, depth_mode, out, B, C, D, H,
                                                              W, make_geom(H, W, flags));
    return check_launch("mvs_warp_fwd");
}

extern "C" int mvs_warp_bwd(const float *grad_out, const float *rot, const float *trans, const float *depth,
                            int depth_mode, float *grad_src, int B, int C, int D, int H, int W, int flags, void *stream)
{
    if (B == 0 || C == 0 || D == 0 || H == 0 || W == 0) return MVS_OK;
    if (int e = check_dims(B, C, D, H, W)) return e;
    MVS_REQUIRE(grad_out && rot && trans && depth && grad_src, "null pointer");
    dim3 grid(cdiv(W, 32), cdiv(H, 8), B * D), block(32, 8);
    warp_bwd_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(grad_out, rot, trans, depth, depth_mode, grad_src, B, C,
                                                              D, H, W, make_geom(H, W, flags));
    return check_launch("mvs_warp_bwd");
}

extern "C" int mvs_warp_taps(const float *rot, const float *trans, const float *depth, int depth_mode, int32_t *x0,
                             int32_t *y0, uint8_t *mask, float *ixy, int B, int D, int H, int W, int flags,
                             void *stream)
{
    if (B == 0 || D == 0 || H == 0 || W == 0) return MVS_OK;
    if (int e = check_dims(B, 1, D, H, W)) return e;
    MVS_REQUIRE(rot && trans && depth && x0 && y0 && mask, "null pointer");
    if (flags & MVS_FAST_COORDS)
        return warp_taps_fast(rot, trans, depth, depth_mode, x0, y0, mask, ixy, B, D, H, W, flags, (cudaStream_t)stream);
    dim3 grid(cdiv(W, 32), cdiv(H, 8), B * D), block(32, 8);
    warp_taps_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(rot, trans, depth, depth_mode, x0, y0, mask, ixy, B, D,
                                                               H, W, make_geom(H, W, flags));
    return check_launch("mvs_warp_taps");
}

template <int NSRC>
static void launch_variance(const float *ref, const SrcPtrs &s, const float *rot, const float *trans,
                            const float *depth, int depth_mode, float *out, int B, int C, int D, int H, int W,
                            int flags, cudaStream_t st)
{
    dim3 grid(cdiv(W, 32), cdiv(H, 8), B * D), block(32, 8);
    warp_variance_kernel<NSRC><<<grid, block, 0, st>>>(ref, s, rot, trans, depth, depth_mode, out, B, C, D, H, W,
                                                       make_geom(H, W, flags), (flags & MVS_REF_SUM_SQUARED) ? 1 : 0);
}

extern "C" int mvs_warp_variance_fwd(const float *ref, const float *const *srcs_host, int nsrc, const float *rot,
                                     const float *trans, const float *depth, int depth_mode, float *out, int B, int C,
                                     int D, int H, int W, int flags, void *stream)
{
    if (B == 0 || C == 0 || D == 0 || H == 0 || W == 0) return MVS_OK;
    if (int e = check_dims(B, C, D, H, W)) return e;
    MVS_REQUIRE(nsrc >= 1 && nsrc <= MVS_MAX_SRC, "nsrc must be in [1, MVS_MAX_SRC]");
    MVS_REQUIRE(ref && srcs_host && rot && trans && depth && out, "null pointer");
    MVS_REQUIRE(depth_mode == MVS_DEPTH_PLANE || depth_mode == MVS_DEPTH_PIXEL, "bad depth_mode");
    SrcPtrs s{};
    for (int i = 0; i < nsrc; ++i) {
        MVS_REQUIRE(srcs_host[i], "null source pointer");
        s.p[i] = srcs_host[i];
    }
    cudaStream_t st = (cudaStream_t)stream;
    switch (nsrc) {
#define CASE(N) case N: launch_variance<N>(ref, s, rot, trans, depth, depth_mode, out, B, C, D, H, W, flags, st); break;
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
    }
    return check_launch("mvs_warp_variance_fwd");
}

template <int NSRC>
static void launch_variance_bwd(const float *gout, const float *ref, const SrcPtrs &s, const float *rot,
                                const float *trans, const float *depth, int depth_mode, float *gref,
                                const DstPtrs &gs, int B, int C, int D, int H, int W, int flags, cudaStream_t st)
{
    dim3 grid(cdiv(W, 32), cdiv(H, 8), B), block(32, 8);
    warp_variance_bwd_kernel<NSRC><<<grid, block, 0, st>>>(gout, ref, s, rot, trans, depth, depth_mode, gref, gs, B, C,
                                                           D, H, W, make_geom(H, W, flags),
                                                           (flags & MVS_REF_SUM_SQUARED) ? 1 : 0);
}

extern "C" int mvs_warp_variance_bwd(const float *grad_out, const float *ref, const float *const *srcs_host, int nsrc,
                                     const float *rot, const float *trans, const float *depth, int depth_mode,
                                     float *grad_ref, float *const *grad_srcs_host, int B, int C, int D, int H, int W,
                                     int flags, void *stream)
{
    if (B == 0 || C == 0 || D == 0 || H == 0 || W == 0) return MVS_OK;
    if (int e = check_dims(B, C, D, H, W)) return e;
    MVS_REQUIRE(nsrc >= 1 && nsrc <= MVS_MAX_SRC, "nsrc must be in [1, MVS_MAX_SRC]");
    MVS_REQUIRE(grad_out && ref && srcs_host && rot && trans && depth, "null pointer");
    SrcPtrs s{};
    DstPtrs gs{};
    for (int i = 0; i < nsrc; ++i) {
        MVS_REQUIRE(srcs_host[i], "null source pointer");
        s.p[i] = srcs_host[i];
        gs.p[i] = grad_srcs_host ? grad_srcs_host[i] : nullptr;
    }
    cudaStream_t st = (cudaStream_t)stream;
    switch (nsrc) {
#define CASE(N) case N: launch_variance_bwd<N>(grad_out, ref, s, rot, trans, depth, depth_mode, grad_ref, gs, B, C, D, H, W, flags, st); break;
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
    }
    return check_launch("mvs_warp_variance_bwd");
}
