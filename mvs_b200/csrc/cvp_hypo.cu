// CVP-MVSNet's statistical depth interval (test branch of calDepthHypo, CVP-MVSNet/models/modules.py:146-209) on the device:
// for every pixel of the up-sampled depth map, back-project at depth D and D + 1, project both into the first source view,
// step one pixel along that epipolar direction and solve the 2x2 system for the depth change that produces the step; the
// hypothesis spacing of the level is mean |delta_d| over the image -- ONE scalar per batch element.  The reference builds it
// from ~40 full-image float64 torch ops and a batched 2x2 inverse per batch element; here one pass, float64 throughout like
// the reference, per-CTA partial sums reduced with a double atomicAdd.
#include "common.cuh"

namespace mvs {

// cam per batch element (59 doubles): inv(K_ref)[9] | inv(E_ref)[16] | E_src[16] | K_src[9] | A = (K_ref R_ref) inv(K_src R_src) [9]
constexpr int CVP_CAM_DOUBLES = 59;

__device__ __forceinline__ void cvp_project(const double *cam, double x, double y, double d, double out[3], double &z)
{
    const double *Kri = cam, *Eri = cam + 9, *Es = cam + 25, *Ks = cam + 41;
    const double X[3] = {x * d, y * d, 1.0 * d};
    double ray[4], w[4], c[3];
    for (int r = 0; r < 3; ++r) ray[r] = Kri[r * 3] * X[0] + Kri[r * 3 + 1] * X[1] + Kri[r * 3 + 2] * X[2];
    ray[3] = 1.0;
    for (int r = 0; r < 4; ++r) w[r] = Eri[r * 4] * ray[0] + Eri[r * 4 + 1] * ray[1] + Eri[r * 4 + 2] * ray[2] + Eri[r * 4 + 3] * ray[3];
    for (int r = 0; r < 3; ++r) c[r] = Es[r * 4] * w[0] + Es[r * 4 + 1] * w[1] + Es[r * 4 + 2] * w[2] + Es[r * 4 + 3] * w[3];
    double img[3];
    for (int r = 0; r < 3; ++r) img[r] = Ks[r * 3] * c[0] + Ks[r * 3 + 1] * c[1] + Ks[r * 3 + 2] * c[2];
    z = img[2];
    out[0] = img[0] / z; out[1] = img[1] / z; out[2] = img[2] / z;
}

__global__ void __launch_bounds__(256)
cvp_interval_kernel(const float *__restrict__ depth, const double *__restrict__ cams, double *__restrict__ sum_abs, int H, int W,
                    double pixel_interval)
{
    __shared__ double s_cam[CVP_CAM_DOUBLES];
    __shared__ double s_part[8];
    const int b = blockIdx.y, tid = threadIdx.x;
    if (tid < CVP_CAM_DOUBLES) s_cam[tid] = cams[(size_t)b * CVP_CAM_DOUBLES + tid];
    __syncthreads();
    const double *A = s_cam + 50;
    double acc = 0.0;
    const long long n = (long long)H * W;
    for (long long i = (long long)blockIdx.x * 256 + tid; i < n; i += (long long)gridDim.x * 256) {
        const int y = (int)(i / W), x = (int)(i % W);
        const double d1 = (double)__ldg(depth + (size_t)b * n + i);
        double x1[3], x2[3], z1, z2;
        cvp_project(s_cam, (double)x, (double)y, d1, x1, z1);
        cvp_project(s_cam, (double)x, (double)y, d1 + 1.0, x2, z2);
        const double k = (x2[1] - x1[1]) / (x2[0] - x1[0]);
        const double theta = atan(k);
        const double x3[3] = {x1[0] + cos(theta) * pixel_interval, x1[1] + sin(theta) * pixel_interval, x1[2]};
        double t1[3], t2[3];
        for (int r = 0; r < 3; ++r) {
            t1[r] = z1 * (A[r * 3] * x1[0] + A[r * 3 + 1] * x1[1] + A[r * 3 + 2] * x1[2]);
            t2[r] = A[r * 3] * x3[0] + A[r * 3 + 1] * x3[1] + A[r * 3 + 2] * x3[2];
        }
        // rows y, z of [X | tmp2] ans = rows y, z of tmp1:   [[y, t2y], [1, t2z]] (dd, .) = (t1y, t1z)
        const double det = (double)y * t2[2] - t2[1];
        const double dd = (t2[2] * t1[1] - t2[1] * t1[2]) / det;
        acc += fabs(dd);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((tid & 31) == 0) s_part[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s_part[w];
        atomicAdd(sum_abs + b, t);
    }
}

}  // namespace mvs

// sum_abs [B] must be zero on entry; the caller divides by H * W.
extern "C" int mvs_cvp_depth_interval(const float *ref_depth, const double *cams, double *sum_abs, int B, int H, int W,
                                      double pixel_interval, void *stream)
{
    if (B == 0 || H == 0 || W == 0) return MVS_OK;
    MVS_REQUIRE(B > 0 && H > 0 && W > 0 && B <= 65535, "bad extents");
    MVS_REQUIRE(ref_depth && cams && sum_abs, "null pointer");
    const long long n = (long long)H * W;
    const int blocks = (int)((n + 255) / 256 < 2 * mvs::sm_count() ? (n + 255) / 256 : 2 * mvs::sm_count());
    mvs::cvp_interval_kernel<<<dim3(blocks, B), 256, 0, (cudaStream_t)stream>>>(ref_depth, cams, sum_abs, H, W, pixel_interval);
    return mvs::check_launch("mvs_cvp_depth_interval");
}
