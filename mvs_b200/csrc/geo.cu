// Geometric-consistency filter of estimated depth maps -- the step AFTER the hot path:
//   reproject_with_depth + check_geometric_consistency   MVSNet/eval.py:138-208 == CasMVSNet/test.py:237-294
//   the per-reference-view fusion loop of filter_depth    MVSNet/eval.py:240-263 (CasMVSNet/test.py:318-343)
// The reference runs this in NumPy (float64) + cv2.remap, one Python loop iteration per (ref, src) pair with ~20
// full-image temporaries; here one thread owns one reference pixel and walks all source views, nothing is materialised.
//
// Arithmetic contract (restated on the CPU in oracle/geo_oracle.py, pinned against the reference + real cv2):
//   * float64 for the projective chain, float32 exactly where the reference casts (.astype(np.float32));
//   * matmul rows as a k-ordered FMA chain (what the BLAS dgemm behind np.matmul does for K = 3 / 4);
//   * cv2.remap(INTER_LINEAR, BORDER_CONSTANT 0) with its fixed-point coordinates: sx = cvRound(x * 32), pixel sx >> 5,
//     fraction (sx & 31) / 32, float weights (1-fy)(1-fx), (1-fy)fx, fy(1-fx), fy fx, left-to-right float sum.
// HBM-bound by construction: per pixel and source view 16 B of taps (L2-resident gather) and, in the fused form, 13 B out.
#include "common.cuh"

namespace mvs {

constexpr int GEO_MAX_SRC = 16;

// per (ref, src) pair, derived on the host exactly as the reference derives them (np.linalg.inv / np.matmul):
//   Kri = inv(K_ref) [9] | Trs = (E_src @ inv(E_ref))[:3] [12] | Ks = K_src [9] | Ksi = inv(K_src) [9] |
//   Tsr = (E_ref @ inv(E_src))[:3] [12] | Kr = K_ref [9]                                            = 60 doubles
constexpr int GEO_CAM_DOUBLES = 60;

struct GeoSrcs {
    const float *depth[GEO_MAX_SRC];
};

__device__ __forceinline__ double dot3(const double *m, double a, double b, double c)
{
    return fma(m[2], c, fma(m[1], b, __dmul_rn(m[0], a)));
}
__device__ __forceinline__ double dot4h(const double *m, double a, double b, double c)     // (a, b, c, 1)
{
    return fma(m[3], 1.0, fma(m[2], c, fma(m[1], b, __dmul_rn(m[0], a))));
}

__device__ __forceinline__ float remap_bilinear(const float *__restrict__ src, int H, int W, float x, float y)
{
    const float xs = __fmul_rn(x, 32.0f), ys = __fmul_rn(y, 32.0f);
    // cvRound of a non-finite / out-of-int-range value is the x86 "integer indefinite" INT_MIN: far outside => 0
    if (!(fabsf(xs) < 2147483648.0f) || !(fabsf(ys) < 2147483648.0f)) return 0.0f;
    const int sx = __float2int_rn(xs), sy = __float2int_rn(ys);
    const int ix = sx >> 5, iy = sy >> 5;
    const float fx = (float)(sx & 31) * (1.0f / 32.0f), fy = (float)(sy & 31) * (1.0f / 32.0f);
    const float gx = __fsub_rn(1.0f, fx), gy = __fsub_rn(1.0f, fy);
    const float w0 = __fmul_rn(gy, gx), w1 = __fmul_rn(gy, fx);
    const float w2 = __fmul_rn(fy, gx), w3 = __fmul_rn(fy, fx);
    const bool x0 = (unsigned)ix < (unsigned)W, x1 = (unsigned)(ix + 1) < (unsigned)W;
    const bool y0 = (unsigned)iy < (unsigned)H, y1 = (unsigned)(iy + 1) < (unsigned)H;
    const float t0 = (x0 && y0) ? __ldg(src + (size_t)iy * W + ix) : 0.0f;
    const float t1 = (x1 && y0) ? __ldg(src + (size_t)iy * W + ix + 1) : 0.0f;
    const float t2 = (x0 && y1) ? __ldg(src + (size_t)(iy + 1) * W + ix) : 0.0f;
    const float t3 = (x1 && y1) ? __ldg(src + (size_t)(iy + 1) * W + ix + 1) : 0.0f;
    return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t0, w0), __fmul_rn(t1, w1)), __fmul_rn(t2, w2)), __fmul_rn(t3, w3));
}

struct GeoOut {
    float depth_reproj, x_rep, y_rep, x_src, y_src;
    bool mask;
};

__device__ __forceinline__ GeoOut geo_pixel(const double *__restrict__ cam, const float *__restrict__ depth_src, int H, int W,
                                            int x, int y, float d_ref, double dist_thresh, float rel_thresh)
{
    const double *Kri = cam, *Trs = cam + 9, *Ks = cam + 21, *Ksi = cam + 30, *Tsr = cam + 39, *Kr = cam + 51;
    const double d = (double)d_ref;
    const double px = __dmul_rn((double)x, d), py = __dmul_rn((double)y, d), pz = d;   // vstack((x, y, 1)) * depth
    const double rx = dot3(Kri, px, py, pz), ry = dot3(Kri + 3, px, py, pz), rz = dot3(Kri + 6, px, py, pz);
    const double sx = dot4h(Trs, rx, ry, rz), sy = dot4h(Trs + 4, rx, ry, rz), sz = dot4h(Trs + 8, rx, ry, rz);
    const double kx = dot3(Ks, sx, sy, sz), ky = dot3(Ks + 3, sx, sy, sz), kz = dot3(Ks + 6, sx, sy, sz);
    const double xs = __ddiv_rn(kx, kz), ys = __ddiv_rn(ky, kz);
    GeoOut o;
    o.x_src = (float)xs;
    o.y_src = (float)ys;
    const float sampled = remap_bilinear(depth_src, H, W, o.x_src, o.y_src);
    const double sd = (double)sampled;
    const double qx = __dmul_rn(xs, sd), qy = __dmul_rn(ys, sd), qz = sd;              // vstack((xy_src, 1)) * sampled
    const double ux = dot3(Ksi, qx, qy, qz), uy = dot3(Ksi + 3, qx, qy, qz), uz = dot3(Ksi + 6, qx, qy, qz);
    const double vx = dot4h(Tsr, ux, uy, uz), vy = dot4h(Tsr + 4, ux, uy, uz), vz = dot4h(Tsr + 8, ux, uy, uz);
    o.depth_reproj = (float)vz;
    const double wx = dot3(Kr, vx, vy, vz), wy = dot3(Kr + 3, vx, vy, vz), wz = dot3(Kr + 6, vx, vy, vz);
    o.x_rep = (float)__ddiv_rn(wx, wz);
    o.y_rep = (float)__ddiv_rn(wy, wz);
    // dist in float64 from the float32-cast reprojection (x2d_reprojected - x_ref promotes to float64)
    const double dx = (double)o.x_rep - (double)x, dy = (double)o.y_rep - (double)y;
    const double dist = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));    // NumPy: two rounded squares, one add (no FMA contraction)
    const float rel = __fdiv_rn(fabsf(__fsub_rn(o.depth_reproj, d_ref)), d_ref);      // float32 throughout
    o.mask = (dist < dist_thresh) && (rel < rel_thresh);
    return o;
}

__global__ void __launch_bounds__(256)
geo_pair_kernel(const float *__restrict__ depth_ref, const float *__restrict__ depth_src, const double *__restrict__ cam,
                uint8_t *__restrict__ mask, float *__restrict__ depth_reproj, float *__restrict__ x_src,
                float *__restrict__ y_src, float *__restrict__ x_rep, float *__restrict__ y_rep, int H, int W,
                double dist_thresh, float rel_thresh, int apply_mask)
{
    __shared__ double s_cam[GEO_CAM_DOUBLES];
    const int t = threadIdx.y * 32 + threadIdx.x;
    if (t < GEO_CAM_DOUBLES) s_cam[t] = cam[t];
    __syncthreads();
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= W || y >= H) return;
    const size_t i = (size_t)y * W + x;
    const GeoOut o = geo_pixel(s_cam, depth_src, H, W, x, y, __ldg(depth_ref + i), dist_thresh, rel_thresh);
    if (mask) mask[i] = o.mask ? 1 : 0;
    if (depth_reproj) depth_reproj[i] = (apply_mask && !o.mask) ? 0.0f : o.depth_reproj;
    if (x_src) x_src[i] = o.x_src;
    if (y_src) y_src[i] = o.y_src;
    if (x_rep) x_rep[i] = o.x_rep;
    if (y_rep) y_rep[i] = o.y_rep;
}

__global__ void __launch_bounds__(256)
geo_fuse_kernel(const float *__restrict__ depth_ref, const float *__restrict__ conf, GeoSrcs srcs, int nsrc,
                const double *__restrict__ cams, int32_t *__restrict__ geo_sum, double *__restrict__ depth_avg,
                uint8_t *__restrict__ final_mask, uint8_t *__restrict__ geo_masks, float *__restrict__ depth_reproj,
                int H, int W, double dist_thresh, float rel_thresh, float conf_thresh, int min_views)
{
    extern __shared__ double s_cams[];
    const int t = threadIdx.y * 32 + threadIdx.x;
    for (int k = t; k < nsrc * GEO_CAM_DOUBLES; k += 256) s_cams[k] = cams[k];
    __syncthreads();
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= W || y >= H) return;
    const size_t plane = (size_t)H * W, i = (size_t)y * W + x;
    const float d_ref = __ldg(depth_ref + i);
    int cnt = 0;
    float acc = 0.0f;                       // sum(all_srcview_depth_ests): float32 adds in source order, starting from 0
    for (int v = 0; v < nsrc; ++v) {
        const GeoOut o = geo_pixel(s_cams + v * GEO_CAM_DOUBLES, srcs.depth[v], H, W, x, y, d_ref, dist_thresh, rel_thresh);
        const float dm = o.mask ? o.depth_reproj : 0.0f;
        cnt += o.mask ? 1 : 0;
        acc = v == 0 ? dm : __fadd_rn(acc, dm);
        if (geo_masks) geo_masks[(size_t)v * plane + i] = o.mask ? 1 : 0;
        if (depth_reproj) depth_reproj[(size_t)v * plane + i] = dm;
    }
    acc = __fadd_rn(acc, d_ref);
    if (geo_sum) geo_sum[i] = cnt;
    if (depth_avg) depth_avg[i] = __ddiv_rn((double)acc, (double)(cnt + 1));           // float32 / int32 promotes to float64
    if (final_mask) final_mask[i] = (cnt >= min_views && __ldg(conf + i) > conf_thresh) ? 1 : 0;
}

// Back-projection of the fused depth map to world points: filter_depth, MVSNet/eval.py:297-300
//   xyz_ref = inv(K_ref) @ (x*d, y*d, d);  xyz_world = (inv(E_ref) @ (xyz_ref, 1))[:3]      (float64, stored as float32)
// cam: inv(K_ref)[9] | inv(E_ref)[:3][12].  Dense output [H,W,3]; pixels outside `mask` get NaN (callers compact with it).
__global__ void __launch_bounds__(256)
geo_backproject_kernel(const double *__restrict__ depth, const uint8_t *__restrict__ mask, const double *__restrict__ cam,
                       float *__restrict__ xyz, int H, int W)
{
    __shared__ double s_cam[21];
    const int t = threadIdx.y * 32 + threadIdx.x;
    if (t < 21) s_cam[t] = cam[t];
    __syncthreads();
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= W || y >= H) return;
    const size_t i = (size_t)y * W + x;
    float ox = __int_as_float(0x7fc00000), oy = ox, oz = ox;
    if (!mask || mask[i]) {
        const double d = depth[i];
        const double px = __dmul_rn((double)x, d), py = __dmul_rn((double)y, d), pz = d;
        const double rx = dot3(s_cam, px, py, pz), ry = dot3(s_cam + 3, px, py, pz), rz = dot3(s_cam + 6, px, py, pz);
        ox = (float)dot4h(s_cam + 9, rx, ry, rz);
        oy = (float)dot4h(s_cam + 13, rx, ry, rz);
        oz = (float)dot4h(s_cam + 17, rx, ry, rz);
    }
    xyz[3 * i] = ox; xyz[3 * i + 1] = oy; xyz[3 * i + 2] = oz;
}

}  // namespace mvs

using namespace mvs;

extern "C" int mvs_geo_backproject(const double *depth, const uint8_t *mask, const double *cam, float *xyz, int H, int W,
                                   void *stream)
{
    if (H == 0 || W == 0) return MVS_OK;
    MVS_REQUIRE(H > 0 && W > 0, "extents must be positive");
    MVS_REQUIRE(depth && cam && xyz, "null pointer");
    dim3 grid(cdiv(W, 32), cdiv(H, 8)), block(32, 8);
    geo_backproject_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(depth, mask, cam, xyz, H, W);
    return check_launch("mvs_geo_backproject");
}

extern "C" int mvs_geo_consistency(const float *depth_ref, const float *depth_src, const double *cam, uint8_t *mask,
                                   float *depth_reproj, float *x_src, float *y_src, float *x_rep, float *y_rep, int H, int W,
                                   double dist_thresh, float rel_thresh, int apply_mask, void *stream)
{
    if (H == 0 || W == 0) return MVS_OK;
    MVS_REQUIRE(H > 0 && W > 0, "extents must be positive");
    MVS_REQUIRE(depth_ref && depth_src && cam, "null pointer");
    dim3 grid(cdiv(W, 32), cdiv(H, 8)), block(32, 8);
    geo_pair_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(depth_ref, depth_src, cam, mask, depth_reproj, x_src, y_src, x_rep,
                                                              y_rep, H, W, dist_thresh, rel_thresh, apply_mask);
    return check_launch("mvs_geo_consistency");
}

extern "C" int mvs_geo_fuse(const float *depth_ref, const float *conf, const void *const *depth_srcs_host, int nsrc,
                            const double *cams, int32_t *geo_sum, double *depth_avg, uint8_t *final_mask, uint8_t *geo_masks,
                            float *depth_reproj, int H, int W, double dist_thresh, float rel_thresh, float conf_thresh,
                            int min_views, void *stream)
{
    if (H == 0 || W == 0) return MVS_OK;
    MVS_REQUIRE(H > 0 && W > 0, "extents must be positive");
    MVS_REQUIRE(nsrc >= 1 && nsrc <= GEO_MAX_SRC, "nsrc must be in [1, 16]");
    MVS_REQUIRE(depth_ref && conf && depth_srcs_host && cams, "null pointer");
    GeoSrcs s{};
    for (int v = 0; v < nsrc; ++v) {
        MVS_REQUIRE(depth_srcs_host[v], "null source depth pointer");
        s.depth[v] = (const float *)depth_srcs_host[v];
    }
    dim3 grid(cdiv(W, 32), cdiv(H, 8)), block(32, 8);
    geo_fuse_kernel<<<grid, block, (size_t)nsrc * GEO_CAM_DOUBLES * sizeof(double), (cudaStream_t)stream>>>(
        depth_ref, conf, s, nsrc, cams, geo_sum, depth_avg, final_mask, geo_masks, depth_reproj, H, W, dist_thresh, rel_thresh,
        conf_thresh, min_views);
    return check_launch("mvs_geo_fuse");
}
