// Layout hand-off between the reference's NC(D)HW tensors and the C8 fast layout
// ([B][C/8][inner][8] bf16: one 16-byte vector per voxel per 8-channel block).  SURVEY.md §8(f) f3.
// Thread <-> (voxel, channel block); lanes along the voxel index, so the eight plane reads are each
// a coalesced row segment and the C8 side is one coalesced 16 B vector per lane.
#include "common.cuh"

namespace mvs {

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T>
__global__ void __launch_bounds__(256)
pack_c8_kernel(const T *__restrict__ src, uint4 *__restrict__ dst, int C, long long inner)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= inner) return;
    const int cb = blockIdx.y, b = blockIdx.z, CB = gridDim.y;
    const T *s = src + ((size_t)b * C + (size_t)cb * 8) * inner + i;
    __nv_bfloat162 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c0 = cb * 8 + 2 * k;
        const float a = c0 < C ? to_f32<T>(s[(size_t)(2 * k) * inner]) : 0.f;
        const float bb = c0 + 1 < C ? to_f32<T>(s[(size_t)(2 * k + 1) * inner]) : 0.f;
        v[k] = __floats2bfloat162_rn(a, bb);
    }
    dst[((size_t)b * CB + cb) * inner + i] = *reinterpret_cast<uint4 *>(v);
}

template <typename T>
__global__ void __launch_bounds__(256)
pack_c8h_kernel(const T *__restrict__ src, uint4 *__restrict__ dst, int C, long long inner)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= inner) return;
    const int cb = blockIdx.y, b = blockIdx.z, CB = gridDim.y;
    const T *s = src + ((size_t)b * C + (size_t)cb * 8) * inner + i;
    __half2 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c0 = cb * 8 + 2 * k;
        float a = c0 < C ? to_f32<T>(s[(size_t)(2 * k) * inner]) : 0.f;
        float bb = c0 + 1 < C ? to_f32<T>(s[(size_t)(2 * k + 1) * inner]) : 0.f;
        // saturate instead of overflowing to inf; fminf / fmaxf drop NaN, so NaN is passed through explicitly
        a = a != a ? a : fminf(fmaxf(a, -65504.f), 65504.f);
        bb = bb != bb ? bb : fminf(fmaxf(bb, -65504.f), 65504.f);
        v[k] = __floats2half2_rn(a, bb);
    }
    dst[((size_t)b * CB + cb) * inner + i] = *reinterpret_cast<uint4 *>(v);
}

template <typename T>
__global__ void __launch_bounds__(256)
unpack_c8_kernel(const uint4 *__restrict__ src, T *__restrict__ dst, int C, long long inner)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= inner) return;
    const int cb = blockIdx.y, b = blockIdx.z, CB = gridDim.y;
    uint4 raw = __ldg(src + ((size_t)b * CB + cb) * inner + i);
    const __nv_bfloat16 *v = reinterpret_cast<const __nv_bfloat16 *>(&raw);
    T *d = dst + ((size_t)b * C + (size_t)cb * 8) * inner + i;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (cb * 8 + k < C) d[(size_t)k * inner] = from_f32<T>(__bfloat162float(v[k]));
}

}  // namespace mvs

using namespace mvs;

extern "C" int mvs_pack_c8(const void *src, int src_dtype, void *dst_c8, int B, int C, int64_t inner, void *stream)
{
    if (B == 0 || C == 0 || inner == 0) return MVS_OK;
    MVS_REQUIRE(B > 0 && C > 0 && inner > 0 && B <= 65535, "bad extents");
    MVS_REQUIRE(src && dst_c8, "null pointer");
    MVS_REQUIRE(src_dtype == MVS_F32 || src_dtype == MVS_BF16, "bad dtype");
    dim3 grid(cdiv(inner, 256), cdiv(C, 8), B);
    if (src_dtype == MVS_F32)
        pack_c8_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float *)src, (uint4 *)dst_c8, C, inner);
    else
        pack_c8_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)src, (uint4 *)dst_c8, C, inner);
    return check_launch("mvs_pack_c8");
}

extern "C" int mvs_pack_c8h(const void *src, int src_dtype, void *dst_c8h, int B, int C, int64_t inner, void *stream)
{
    if (B == 0 || C == 0 || inner == 0) return MVS_OK;
    MVS_REQUIRE(B > 0 && C > 0 && inner > 0 && B <= 65535, "bad extents");
    MVS_REQUIRE(src && dst_c8h, "null pointer");
    MVS_REQUIRE(src_dtype == MVS_F32 || src_dtype == MVS_BF16, "bad dtype");
    dim3 grid(cdiv(inner, 256), cdiv(C, 8), B);
    if (src_dtype == MVS_F32)
        pack_c8h_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float *)src, (uint4 *)dst_c8h, C, inner);
    else
        pack_c8h_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)src, (uint4 *)dst_c8h, C, inner);
    return check_launch("mvs_pack_c8h");
}

extern "C" int mvs_unpack_c8(const void *src_c8, void *dst, int dst_dtype, int B, int C, int64_t inner, void *stream)
{
    if (B == 0 || C == 0 || inner == 0) return MVS_OK;
    MVS_REQUIRE(B > 0 && C > 0 && inner > 0 && B <= 65535, "bad extents");
    MVS_REQUIRE(src_c8 && dst, "null pointer");
    MVS_REQUIRE(dst_dtype == MVS_F32 || dst_dtype == MVS_BF16, "bad dtype");
    dim3 grid(cdiv(inner, 256), cdiv(C, 8), B);
    if (dst_dtype == MVS_F32)
        unpack_c8_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4 *)src_c8, (float *)dst, C, inner);
    else
        unpack_c8_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4 *)src_c8, (__nv_bfloat16 *)dst, C, inner);
    return check_launch("mvs_unpack_c8");
}
