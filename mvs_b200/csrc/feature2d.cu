// Caller-side kernels of the FeatureNet hand-off: what CasMVSNet's 2D FPN extractor
// (CasMVSNet/models/module.py:304-405) needs AROUND the 3x3 convolutions, which run on the tcgen05 kernel of
// conv3d_umma.cu with D = 1 and fp16 C8 activations (MVS_ACT_F16):
//   mvs_img_to_c8h      uint8 / fp32 images [N,3,H,W] -> fp16 C8 [N,1,H,W,8] (channels 3..7 zero); uint8 is scaled by 1/255
//                       exactly like the loader does on the host (CasMVSNet/datasets/general_eval.py:81-86)
//   mvs_s2d_c8          2x2 space-to-depth of a C8 map: the 5x5 / stride-2 / pad-2 convolutions (module.py:336,342) become
//                       3x3 / stride-1 / pad-1 convolutions over 4*Cin channels (host re-lays the weights: featurenet.py)
//   mvs_fpn_merge_c8h   the FPN lateral step (module.py:393-398) in one pass:
//                       out = nearest_up2(prev) + conv1x1(x) + bias   -- 32 channels out, never materialising the up-sampled
//                       map or the lateral map; weights sit in the constant bank (kernel parameter), fp32 math
// All HBM-bound, one 16-byte vector per lane per channel block, fully coalesced.
#include <cstring>

#include "common.cuh"

namespace mvs {

__global__ void __launch_bounds__(256)
img_to_c8h_kernel(const void *__restrict__ img, int is_u8, uint4 *__restrict__ dst, long long plane)
{
    // (lets a dependent conv layer launched with programmatic stream serialization start its prologue under our tail)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= plane) return;
    const int n = blockIdx.y;
    float c[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const size_t o = ((size_t)n * 3 + k) * plane + i;
        c[k] = is_u8 ? __fdiv_rn((float)__ldg((const uint8_t *)img + o), 255.0f) : __ldg((const float *)img + o);
    }
    const __half2 a = __floats2half2_rn(c[0], c[1]), b = __floats2half2_rn(c[2], 0.f);
    dst[(size_t)n * plane + i] = make_uint4(*reinterpret_cast<const uint32_t *>(&a), *reinterpret_cast<const uint32_t *>(&b), 0u, 0u);
}

// uint8 images whose planes are 4-byte aligned: four pixels per thread (one 32-bit load per colour plane) and the 256 possible
// quotients v / 255 -- the same correctly rounded fp32 division as above -- from a shared-memory table instead of 12 divisions
__global__ void __launch_bounds__(256)
img_u8x4_to_c8h_kernel(const uint32_t *__restrict__ img, uint4 *__restrict__ dst, long long plane4)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    __shared__ float lut[256];
    lut[threadIdx.x] = __fdiv_rn((float)threadIdx.x, 255.0f);
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= plane4) return;
    const int n = blockIdx.y;
    uint32_t v[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) v[k] = __ldg(img + ((size_t)n * 3 + k) * plane4 + i);
    uint4 *o = dst + ((size_t)n * plane4 + i) * 4;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const __half2 a = __floats2half2_rn(lut[(v[0] >> (8 * p)) & 255u], lut[(v[1] >> (8 * p)) & 255u]);
        const __half2 b = __floats2half2_rn(lut[(v[2] >> (8 * p)) & 255u], 0.f);
        o[p] = make_uint4(*reinterpret_cast<const uint32_t *>(&a), *reinterpret_cast<const uint32_t *>(&b), 0u, 0u);
    }
}

// Map layouts: "batch-major" [N][CB][H][W][8] (what the builder takes, one contiguous map per view) or "folded"
// [CB][N][H][W][8] -- the SAME bytes the tcgen05 convolution reads as ONE volume [B=1][CB][D=N][H][W][8], so that the N images
// of a reference view ride the kernel's row axis (3D weights populated in the kd = 1 slice only: no mixing between images)
// instead of being N one-row volumes: 2-3x fewer pipeline steps per pixel (profiles/r2c_launches.csv vs r2d).
__device__ __forceinline__ size_t map_plane(int n, int cb, int N, int CB, bool folded)
{
    return folded ? (size_t)cb * N + n : (size_t)n * CB + cb;
}

// dst[n][(py*2+px)*CB + cb][y][x] = src[n][cb][2y+py][2x+px]  (zero outside)
__global__ void __launch_bounds__(256)
s2d_c8_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst, int N, int CB, int H, int W, int Ho, int Wo, int flags)
{
    // (lets a dependent conv layer launched with programmatic stream serialization start its prologue under our tail)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= Wo || y >= Ho) return;
    const int n = blockIdx.z / CB, cb = blockIdx.z % CB;
    const uint4 *s = src + map_plane(n, cb, N, CB, flags & 1) * H * W;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int yy = 2 * y + (p >> 1), xx = 2 * x + (p & 1);
        const uint4 v = (yy < H && xx < W) ? __ldg(s + (size_t)yy * W + xx) : make_uint4(0, 0, 0, 0);
        dst[(map_plane(n, p * CB + cb, N, 4 * CB, flags & 2) * Ho + y) * Wo + x] = v;
    }
}

struct Lateral {           // 1x1 convolution Cin (8 | 16) -> 32 + bias, in the constant bank
    float w[32][16];
    float b[32];
};

__device__ __forceinline__ float2 h2f(uint32_t u) { return __half22float2(*reinterpret_cast<__half2 *>(&u)); }
__device__ __forceinline__ uint32_t f2h(float a, float b)
{
    __half2 v = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&v);
}

template <int CIN>
__global__ void __launch_bounds__(128)
fpn_merge_kernel(const uint4 *__restrict__ x, const uint4 *__restrict__ prev, uint4 *__restrict__ out,
                 const __grid_constant__ Lateral L, int H, int W, int Hp, int Wp, int flags)
{
    // (lets a dependent conv layer launched with programmatic stream serialization start its prologue under our tail)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int px = blockIdx.x * 128 + threadIdx.x, py = blockIdx.y, n = blockIdx.z, N = gridDim.z;
    if (px >= W) return;
    const size_t plane = (size_t)H * W, pix = (size_t)py * W + px;
    float xin[CIN];
#pragma unroll
    for (int cb = 0; cb < CIN / 8; ++cb) {
        const uint4 v = __ldg(x + map_plane(n, cb, N, CIN / 8, flags & 1) * plane + pix);
        const float2 a = h2f(v.x), b = h2f(v.y), c = h2f(v.z), d = h2f(v.w);
        xin[cb * 8 + 0] = a.x; xin[cb * 8 + 1] = a.y; xin[cb * 8 + 2] = b.x; xin[cb * 8 + 3] = b.y;
        xin[cb * 8 + 4] = c.x; xin[cb * 8 + 5] = c.y; xin[cb * 8 + 6] = d.x; xin[cb * 8 + 7] = d.y;
    }
    const size_t pplane = (size_t)Hp * Wp, ppix = (size_t)min(py >> 1, Hp - 1) * Wp + min(px >> 1, Wp - 1);
#pragma unroll
    for (int ob = 0; ob < 4; ++ob) {
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float a = L.b[ob * 8 + k];
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) a = fmaf(L.w[ob * 8 + k][ci], xin[ci], a);
            acc[k] = a;
        }
        if (prev) {                 // F.interpolate(scale_factor=2, mode="nearest"): src index = floor(dst / 2)
            const uint4 p = __ldg(prev + map_plane(n, ob, N, 4, flags & 4) * pplane + ppix);
            const float2 a = h2f(p.x), b = h2f(p.y), c = h2f(p.z), d = h2f(p.w);
            acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y; acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
        }
        out[map_plane(n, ob, N, 4, flags & 2) * plane + pix] = make_uint4(f2h(acc[0], acc[1]), f2h(acc[2], acc[3]), f2h(acc[4], acc[5]), f2h(acc[6], acc[7]));
    }
}

struct BorderCorr {        // per-channel additive terms of the 8 border classes (row class * 3 + column class; centre unused)
    float c[9][32];
};

// y[n, :, h, w] += corr[rc(h) * 3 + cc(w)] on the one-pixel border of fp16 C8 maps [N][CB][H][W][8] (rc: 0 top, 1 inner, 2 bottom)
__global__ void __launch_bounds__(128)
border_add_c8h_kernel(uint4 *__restrict__ y, const __grid_constant__ BorderCorr B, int CB, int H, int W)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int per = 2 * W + 2 * (H - 2);                       // perimeter pixels: top row, bottom row, left / right columns
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= per) return;
    int h, w;
    if (i < W) { h = 0; w = i; }
    else if (i < 2 * W) { h = H - 1; w = i - W; }
    else { const int k = i - 2 * W; h = 1 + (k >> 1); w = (k & 1) ? W - 1 : 0; }
    const int cls = (h == 0 ? 0 : (h == H - 1 ? 2 : 1)) * 3 + (w == 0 ? 0 : (w == W - 1 ? 2 : 1));
    const int cb = blockIdx.y, n = blockIdx.z;
    uint4 *p = y + (((size_t)n * CB + cb) * H + h) * W + w;
    const uint4 v = *p;
    const float2 a = h2f(v.x), b = h2f(v.y), c = h2f(v.z), d = h2f(v.w);
    const float *k = &B.c[cls][cb * 8];
    *p = make_uint4(f2h(a.x + k[0], a.y + k[1]), f2h(b.x + k[2], b.y + k[3]), f2h(c.x + k[4], c.y + k[5]), f2h(d.x + k[6], d.y + k[7]));
}

}  // namespace mvs

using namespace mvs;

extern "C" int mvs_border_add_c8h(void *y_c8h, const float *corr_host, int N, int C, int H, int W, void *stream)
{
    if (N == 0 || H == 0 || W == 0) return MVS_OK;
    MVS_REQUIRE(y_c8h && corr_host, "null pointer");
    MVS_REQUIRE(N > 0 && N <= 65535 && C > 0 && C <= 32 && C % 8 == 0 && H >= 2 && W >= 2, "bad extents (C <= 32, H, W >= 2)");
    BorderCorr B;
    memset(&B, 0, sizeof(B));
    for (int k = 0; k < 9; ++k)
        for (int c = 0; c < C; ++c) B.c[k][c] = corr_host[k * C + c];
    const int per = 2 * W + 2 * (H - 2);
    border_add_c8h_kernel<<<dim3(cdiv(per, 128), C / 8, N), 128, 0, (cudaStream_t)stream>>>((uint4 *)y_c8h, B, C / 8, H, W);
    return check_launch("mvs_border_add_c8h");
}

extern "C" int mvs_img_to_c8h(const void *img, int src_dtype, void *dst_c8h, int N, int H, int W, void *stream)
{
    if (N == 0 || H == 0 || W == 0) return MVS_OK;
    MVS_REQUIRE(N > 0 && H > 0 && W > 0 && N <= 65535, "bad extents");
    MVS_REQUIRE(img && dst_c8h, "null pointer");
    MVS_REQUIRE(src_dtype == MVS_F32 || src_dtype == MVS_U8, "images must be float32 or uint8");
    const long long plane = (long long)H * W;
    if (src_dtype == MVS_U8 && plane % 4 == 0 && ((uintptr_t)img & 3) == 0)
        img_u8x4_to_c8h_kernel<<<dim3(cdiv(plane / 4, 256), N), 256, 0, (cudaStream_t)stream>>>((const uint32_t *)img, (uint4 *)dst_c8h, plane / 4);
    else
        img_to_c8h_kernel<<<dim3(cdiv(plane, 256), N), 256, 0, (cudaStream_t)stream>>>(img, src_dtype == MVS_U8, (uint4 *)dst_c8h, plane);
    return check_launch("mvs_img_to_c8h");
}

extern "C" int mvs_s2d_c8(const void *src_c8, void *dst_c8, int N, int CB, int H, int W, int flags, void *stream)
{
    if (N == 0 || CB == 0 || H == 0 || W == 0) return MVS_OK;
    MVS_REQUIRE(N > 0 && CB > 0 && H > 0 && W > 0 && (long long)N * CB <= 65535, "bad extents");
    MVS_REQUIRE(src_c8 && dst_c8, "null pointer");
    const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    MVS_REQUIRE(cdiv(Ho, 4) <= 65535, "H too large");
    s2d_c8_kernel<<<dim3(cdiv(Wo, 64), cdiv(Ho, 4), N * CB), 256, 0, (cudaStream_t)stream>>>((const uint4 *)src_c8, (uint4 *)dst_c8, N, CB, H, W, Ho, Wo, flags);
    return check_launch("mvs_s2d_c8");
}

extern "C" int mvs_fpn_merge_c8h(const void *x_c8h, const float *w_host, const float *bias_host, const void *prev_c8h,
                                 void *out_c8h, int N, int Cin, int H, int W, int Hp, int Wp, int flags, void *stream)
{
    if (N == 0 || H == 0 || W == 0) return MVS_OK;
    MVS_REQUIRE(N > 0 && H > 0 && W > 0 && N <= 65535 && H <= 65535, "bad extents");
    MVS_REQUIRE(Cin == 8 || Cin == 16, "the lateral 1x1 convolutions of the FPN have 8 or 16 input channels");
    MVS_REQUIRE(x_c8h && w_host && out_c8h, "null pointer");
    MVS_REQUIRE(!prev_c8h || (Hp >= (H + 1) / 2 && Wp >= (W + 1) / 2), "prev must be the half-resolution map");
    Lateral L;
    memset(&L, 0, sizeof(L));
    for (int co = 0; co < 32; ++co) {
        for (int ci = 0; ci < Cin; ++ci) L.w[co][ci] = w_host[co * Cin + ci];
        L.b[co] = bias_host ? bias_host[co] : 0.f;
    }
    dim3 grid(cdiv(W, 128), H, N);
    if (Cin == 8)
        fpn_merge_kernel<8><<<grid, 128, 0, (cudaStream_t)stream>>>((const uint4 *)x_c8h, (const uint4 *)prev_c8h, (uint4 *)out_c8h, L, H, W, Hp, Wp, flags);
    else
        fpn_merge_kernel<16><<<grid, 128, 0, (cudaStream_t)stream>>>((const uint4 *)x_c8h, (const uint4 *)prev_c8h, (uint4 *)out_c8h, L, H, W, Hp, Wp, flags);
    return check_launch("mvs_fpn_merge_c8h");
}
