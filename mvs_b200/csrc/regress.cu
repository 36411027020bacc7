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

}  // namespace mvs

using namespace mvs;

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
