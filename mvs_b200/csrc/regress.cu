// Fused softmax-over-D + soft-argmin depth + expected-index + 4-window photometric confidence.
// Replaces F.softmax / depth_regression (x2) / F.pad / avg_pool3d / gather
// (MVSNet/models/mvsnet.py:183-191; CasMVSNet/models/cas_mvsnet.py:51-64; CVP net.py:185-199):
// one read of the logits (plus a 4-element re-read that hits L1), no probability volume unless asked.
// Thread <-> pixel, lanes along x: every logits[b,d,:,:] access is a coalesced row segment.
#include "common.cuh"

namespace mvs {

__global__ void __launch_bounds__(256)
softargmin_conf_kernel(const float *__restrict__ logits, const float *__restrict__ depth, int depth_mode,
                       float *__restrict__ out_depth, float *__restrict__ out_conf, float *__restrict__ out_prob,
                       int32_t *__restrict__ out_index, int B, int D, long long plane, int clamp_index, int is_prob)
{
    // (lets a dependent conv layer launched with programmatic stream serialization start its prologue under our tail)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= plane) return;
    const float *l = logits + (size_t)b * D * plane + i;
    // pass 1: max over D (keeps exp() in range exactly like F.softmax)
    float m = -INFINITY;
    if (!is_prob)
        for (int d = 0; d < D; ++d) m = fmaxf(m, __ldg(l + (size_t)d * plane));
    // pass 2 (L1/L2-resident re-read): sum e, sum e*depth, sum e*k
    float s = 0.f, sd = 0.f, sk = 0.f;
    const float *dp = depth_mode == MVS_DEPTH_PLANE ? depth + (size_t)b * D : depth + (size_t)b * D * plane + i;
    const size_t dstride = depth_mode == MVS_DEPTH_PLANE ? 1 : (size_t)plane;
    for (int d = 0; d < D; ++d) {
        const float lv = __ldg(l + (size_t)d * plane);
        const float e = is_prob ? lv : expf(lv - m);
        s += e;
        sd = fmaf(e, __ldg(dp + d * dstride), sd);
        sk = fmaf(e, (float)d, sk);
    }
    const float inv = is_prob ? 1.0f : 1.0f / s;   // depth_regression(p, d): p is used as given
    const float idxf = sk * inv;
    int k = (int)idxf;                                   // .long(): truncate toward zero
    if (clamp_index) k = min(max(k, 0), D - 1);
    float c4 = 0.f;
#pragma unroll
    for (int j = -1; j <= 2; ++j) {
        const int kk = k + j;
        if (kk >= 0 && kk < D) {
            const float lv = __ldg(l + (size_t)kk * plane);
            c4 += (is_prob ? lv : expf(lv - m)) * inv;
        }
    }
    out_depth[(size_t)b * plane + i] = sd * inv;
    out_conf[(size_t)b * plane + i] = c4;
    if (out_index) out_index[(size_t)b * plane + i] = k;
    if (out_prob)
        for (int d = 0; d < D; ++d)
            out_prob[((size_t)b * D + d) * plane + i] =
                is_prob ? __ldg(l + (size_t)d * plane) : expf(__ldg(l + (size_t)d * plane) - m) * inv;
}

// get_depth_range_samples, per-pixel branch (CasMVSNet/models/module.py:485-504): each op rounded
// separately like the reference's tensor expression.
__global__ void __launch_bounds__(256)
depth_range_kernel(const float *__restrict__ cur, float half, int ndepth, float *__restrict__ out, long long plane)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= plane) return;
    const float c = __ldg(cur + (size_t)b * plane + i);
    const float lo = __fsub_rn(c, half), hi = __fadd_rn(c, half);
    const float step = __fdiv_rn(__fsub_rn(hi, lo), (float)(ndepth - 1));
    for (int k = 0; k < ndepth; ++k)
        out[((size_t)b * ndepth + k) * plane + i] = __fadd_rn(lo, __fmul_rn((float)k, step));
}

// ATen's area_pixel_compute_source_index (align_corners = False, non-cubic) + tap set-up of
// upsample_bilinear2d / upsample_trilinear3d: src = scale * (dst + 0.5) - 0.5, clamped at 0.
__device__ __forceinline__ void lin_tap(float scale, int dst, int in_size, int &i0, int &i1, float &l0, float &l1)
{
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    src = src < 0.f ? 0.f : src;
    i0 = (int)src;
    i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
    l1 = src - (float)i0;
    l0 = 1.0f - l1;
}

// Fused hypothesis generation of one cascade stage (CasMVSNet/models/cas_mvsnet.py:129-151,
// module.py:485-504): bilinear up-sampling of the previous depth to the image extent, +-ndepth/2
// samples around it, trilinear resampling (D unchanged) to the stage extent -- without
// materialising the [B,D,H,W] full-resolution sample volume.  Thread <-> output pixel, lanes along x.
__global__ void __launch_bounds__(256)
cas_hypotheses_kernel(const float *__restrict__ prev, int hp, int wp, int H, int W, int h, int w, int nd, float half,
                      float *__restrict__ out)
{
    // (lets a dependent conv layer launched with programmatic stream serialization start its prologue under our tail)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    if (x >= w) return;
    const float *pd = prev + (size_t)b * hp * wp;
    // trilinear taps of the (H,W) -> (h,w) resampling
    int Y0, Y1, X0, X1; float ly0, ly1, lx0, lx1;
    lin_tap((float)H / (float)h, y, H, Y0, Y1, ly0, ly1);
    lin_tap((float)W / (float)w, x, W, X0, X1, lx0, lx1);
    // up-sampled previous depth at the (up to) four full-resolution positions
    float lo[2][2], step[2][2];
    const float sh = (float)hp / (float)H, sw = (float)wp / (float)W;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int Y = a ? Y1 : Y0, X = c ? X1 : X0;
            int y0, y1, x0, x1; float my0, my1, mx0, mx1;
            lin_tap(sh, Y, hp, y0, y1, my0, my1);
            lin_tap(sw, X, wp, x0, x1, mx0, mx1);
            const float cur = my0 * (mx0 * __ldg(pd + y0 * wp + x0) + mx1 * __ldg(pd + y0 * wp + x1)) +
                              my1 * (mx0 * __ldg(pd + y1 * wp + x0) + mx1 * __ldg(pd + y1 * wp + x1));
            const float l = __fsub_rn(cur, half), hi = __fadd_rn(cur, half);
            lo[a][c] = l;
            step[a][c] = __fdiv_rn(__fsub_rn(hi, l), (float)(nd - 1));
        }
    const size_t plane = (size_t)h * w;
    float *op = out + (size_t)b * nd * plane + (size_t)y * w + x;
    for (int k = 0; k < nd; ++k) {
        float v[2][2];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int c = 0; c < 2; ++c) v[a][c] = __fadd_rn(lo[a][c], __fmul_rn((float)k, step[a][c]));
        // depth axis: same extent in and out => tap weight exactly (1, 0) on plane k
        const float val = ly0 * (lx0 * v[0][0] + lx1 * v[0][1]) + ly1 * (lx0 * v[1][0] + lx1 * v[1][1]);
        op[(size_t)k * plane] = val;
    }
}

}  // namespace mvs

using namespace mvs;

extern "C" int mvs_cas_hypotheses(const float *prev_depth, int hp, int wp, int H, int W, int h, int w, int ndepth,
                                  double interval, float *out, int B, void *stream)
{
    if (B == 0 || h == 0 || w == 0) return MVS_OK;
    MVS_REQUIRE(B > 0 && B <= 65535 && hp > 0 && wp > 0 && H > 0 && W > 0 && h > 0 && w > 0 && h <= 65535, "bad extents");
    MVS_REQUIRE(ndepth > 1, "ndepth must be > 1");
    MVS_REQUIRE(prev_depth && out, "null pointer");
    const float half = (float)((double)ndepth / 2.0 * interval);
    dim3 grid(cdiv(w, 128), h, B);
    cas_hypotheses_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(prev_depth, hp, wp, H, W, h, w, ndepth, half, out);
    return check_launch("mvs_cas_hypotheses");
}

extern "C" int mvs_softargmin_conf_fwd(const float *logits, const float *depth, int depth_mode, float *out_depth,
                                       float *out_conf, float *out_prob, int32_t *out_index, int B, int D, int H,
                                       int W, int flags, void *stream)
{
    if (B == 0 || H == 0 || W == 0) return MVS_OK;
    MVS_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && B <= 65535, "bad extents");
    MVS_REQUIRE(logits && depth && out_depth && out_conf, "null pointer");
    MVS_REQUIRE(depth_mode == MVS_DEPTH_PLANE || depth_mode == MVS_DEPTH_PIXEL, "bad depth_mode");
    const long long plane = (long long)H * W;
    dim3 grid(cdiv(plane, 256), B);
    softargmin_conf_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(logits, depth, depth_mode, out_depth, out_conf,
                                                                   out_prob, out_index, B, D, plane,
                                                                   (flags & MVS_CLAMP_INDEX) ? 1 : 0,
                                                                   (flags & MVS_INPUT_IS_PROB) ? 1 : 0);
    return check_launch("mvs_softargmin_conf_fwd");
}

extern "C" int mvs_depth_range_samples(const float *cur, double interval, int ndepth, float *out, int B, int H, int W,
                                       void *stream)
{
    if (B == 0 || H == 0 || W == 0) return MVS_OK;
    MVS_REQUIRE(B > 0 && ndepth > 1 && H > 0 && W > 0 && B <= 65535, "bad extents");
    MVS_REQUIRE(cur && out, "null pointer");
    const long long plane = (long long)H * W;
    // ndepth / 2 * interval evaluated like the Python expression (double), then applied to fp32 tensors
    const float half = (float)((double)ndepth / 2.0 * interval);
    dim3 grid(cdiv(plane, 256), B);
    depth_range_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(cur, half, ndepth, out, plane);
    return check_launch("mvs_depth_range_samples");
}

// ---- relative poses of a CasMVSNet projection block in ONE launch ---------------------------------------------------
// proj [n_sets][N][2][4][4] fp32 ([..,0] = extrinsic E, [..,1,:3,:3] = intrinsic K, CasMVSNet/datasets/general_eval.py);
// per set: fused_v = E_v with [:3,:4] = K_v[:3,:3] @ E_v[:3,:4] (cas_mvsnet.py:30-33), rel_v = fused_v @ inverse(fused_0)
// (module.py:257-259) -> rot [n_sets][N-1][9], trans [n_sets][N-1][3].  The reference does this with ~25 tiny ATen / cuSOLVER
// launches per stage in fp32; here one thread per (set, source view) in fp64, rounded once to fp32 (fast path only: the strict
// path keeps the torch ops so that both sides of a bit-exactness test consume identical rot / trans bits).
namespace mvs {
__device__ inline void fuse_proj(const float *p, double f[16])
{
    const float *E = p, *K = p + 16;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) {
            double a = 0.0;
            for (int k = 0; k < 3; ++k) a += (double)K[r * 4 + k] * (double)E[k * 4 + c];
            f[r * 4 + c] = a;
        }
    for (int c = 0; c < 4; ++c) f[12 + c] = (double)E[12 + c];
}

__global__ void cas_poses_kernel(const float *__restrict__ proj, float *__restrict__ rot, float *__restrict__ trans, int n_sets, int N)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_sets * (N - 1)) return;
    const int set = i / (N - 1), v = i % (N - 1) + 1;
    double a[16], s[16], inv[16];
    fuse_proj(proj + ((size_t)set * N) * 32, a);
    fuse_proj(proj + ((size_t)set * N + v) * 32, s);
    // Gauss-Jordan with partial pivoting on [a | I]
    double m[4][8];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 8; ++c) m[r][c] = c < 4 ? a[r * 4 + c] : (c - 4 == r ? 1.0 : 0.0);
    for (int col = 0; col < 4; ++col) {
        int piv = col;
        for (int r = col + 1; r < 4; ++r)
            if (fabs(m[r][col]) > fabs(m[piv][col])) piv = r;
        for (int c = 0; c < 8; ++c) { const double t = m[col][c]; m[col][c] = m[piv][c]; m[piv][c] = t; }
        const double d = 1.0 / m[col][col];
        for (int c = 0; c < 8; ++c) m[col][c] *= d;
        for (int r = 0; r < 4; ++r) {
            if (r == col) continue;
            const double f = m[r][col];
            for (int c = 0; c < 8; ++c) m[r][c] -= f * m[col][c];
        }
    }
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) inv[r * 4 + c] = m[r][c + 4];
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 4; ++c) {
            double acc = 0.0;
            for (int k = 0; k < 4; ++k) acc += s[r * 4 + k] * inv[k * 4 + c];
            if (c < 3) rot[(size_t)i * 9 + r * 3 + c] = (float)acc;
            else trans[(size_t)i * 3 + r] = (float)acc;
        }
    }
}
}  // namespace mvs

extern "C" int mvs_cas_poses(const float *proj, float *rot, float *trans, int n_sets, int N, void *stream)
{
    if (n_sets == 0 || N <= 1) return MVS_OK;
    MVS_REQUIRE(n_sets > 0 && N >= 2, "need at least one set and two views");
    MVS_REQUIRE(proj && rot && trans, "null pointer");
    const int total = n_sets * (N - 1);
    mvs::cas_poses_kernel<<<mvs::cdiv(total, 64), 64, 0, (cudaStream_t)stream>>>(proj, rot, trans, n_sets, N);
    return mvs::check_launch("mvs_cas_poses");
}
