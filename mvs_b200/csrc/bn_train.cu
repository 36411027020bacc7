// Train-mode BatchNorm3d (+ ReLU) for the training step (BASELINE configs[3]): batch statistics, two passes each way.
// Replaces F.batch_norm(training=True) + F.relu and their autograd under loss.backward()
// (MVSNet/models/module.py:26-33, CasMVSNet/models/module.py:139,182, CVP-MVSNet/models/net.py:52-89; CasMVSNet/train.py:165-170).
// ATen's batch-norm kernels for NC(DHW) tensors launch a few CTAs per CHANNEL: with 8-64 channels and 10^7 voxels per channel
// they run on a fraction of the GPU (12.7 ms per backward call at the CVP coarse level, 2.3 GB of traffic = 0.4 ms at HBM
// speed).  Here every pass is a full-grid streaming kernel over [B][C][S]:
//   forward   pass 1  mvs_bn_stats      per channel  sum x, sum x^2                      (fp32 per thread, double across threads / CTAs)
//             pass 2  mvs_bn_apply      y = [relu]((x - mean) * invstd * gamma + beta)
//   backward  pass 1  mvs_bn_bwd_stats  per channel  sum g, sum g * xhat     with g = dy * [y > 0], y recomputed from x
//             pass 2  mvs_bn_bwd_apply  dx = gamma * invstd * (g - sum_g / M - xhat * sum_gx / M)
// The host side (train.py) turns the sums into mean / invstd / running statistics exactly as torch does (biased variance for
// the normalisation, unbiased for running_var).
#include "common.cuh"

namespace mvs {

constexpr int BN_THREADS = 256;
constexpr int BN_CHUNK = 16384;          // elements of one (b, c) row a CTA reduces / transforms

__device__ __forceinline__ double block_sum(double v, double *sh)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (warp == 0) {
        t = lane < BN_THREADS / 32 ? sh[lane] : 0.0;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    return t;                              // valid in thread 0
}

__global__ void __launch_bounds__(BN_THREADS)
bn_stats_kernel(const float *__restrict__ x, double *__restrict__ sums, int C, long long S)
{
    __shared__ double sh[BN_THREADS / 32];
    const int c = blockIdx.y, b = blockIdx.z;
    const float *row = x + ((size_t)b * C + c) * S;
    const long long i0 = (long long)blockIdx.x * BN_CHUNK, i1 = min(i0 + BN_CHUNK, S);
    float s = 0.f, q = 0.f;
    for (long long i = i0 + threadIdx.x; i < i1; i += BN_THREADS) {
        const float v = __ldg(row + i);
        s += v; q = fmaf(v, v, q);
    }
    const double ts = block_sum((double)s, sh), tq = block_sum((double)q, sh);
    if (threadIdx.x == 0) { atomicAdd(sums + c, ts); atomicAdd(sums + C + c, tq); }
}

// y = [relu](x * a[c] + k[c])   with a = invstd * gamma, k = beta - mean * a (host-computed, fp32)
__global__ void __launch_bounds__(BN_THREADS)
bn_apply_kernel(const float *__restrict__ x, const float *__restrict__ a, const float *__restrict__ k, float *__restrict__ y,
                int C, long long S, int relu)
{
    const int c = blockIdx.y, b = blockIdx.z;
    const size_t base = ((size_t)b * C + c) * S;
    const float ac = __ldg(a + c), kc = __ldg(k + c);
    const long long i0 = (long long)blockIdx.x * BN_CHUNK, i1 = min(i0 + BN_CHUNK, S);
    for (long long i = i0 + threadIdx.x; i < i1; i += BN_THREADS) {
        float v = fmaf(__ldg(x + base + i), ac, kc);
        if (relu) v = fmaxf(v, 0.f);
        y[base + i] = v;
    }
}

// sums[c] += sum g, sums[C + c] += sum g * xhat,  g = dy * [x * a + k > 0] (relu) | dy,  xhat = (x - mean) * invstd
__global__ void __launch_bounds__(BN_THREADS)
bn_bwd_stats_kernel(const float *__restrict__ x, const float *__restrict__ dy, const float *__restrict__ a,
                    const float *__restrict__ k, const float *__restrict__ mean, const float *__restrict__ invstd,
                    double *__restrict__ sums, int C, long long S, int relu)
{
    __shared__ double sh[BN_THREADS / 32];
    const int c = blockIdx.y, b = blockIdx.z;
    const size_t base = ((size_t)b * C + c) * S;
    const float ac = __ldg(a + c), kc = __ldg(k + c), mc = __ldg(mean + c), ic = __ldg(invstd + c);
    const long long i0 = (long long)blockIdx.x * BN_CHUNK, i1 = min(i0 + BN_CHUNK, S);
    float s = 0.f, q = 0.f;
    for (long long i = i0 + threadIdx.x; i < i1; i += BN_THREADS) {
        const float xv = __ldg(x + base + i);
        float g = __ldg(dy + base + i);
        if (relu && !(fmaf(xv, ac, kc) > 0.f)) g = 0.f;
        s += g; q = fmaf(g, (xv - mc) * ic, q);
    }
    const double ts = block_sum((double)s, sh), tq = block_sum((double)q, sh);
    if (threadIdx.x == 0) { atomicAdd(sums + c, ts); atomicAdd(sums + C + c, tq); }
}

// dx = ga[c] * (g - mg[c] - xhat * mgx[c])   with ga = gamma * invstd, mg = sum_g / M, mgx = sum_gx / M (host-computed)
__global__ void __launch_bounds__(BN_THREADS)
bn_bwd_apply_kernel(const float *__restrict__ x, const float *__restrict__ dy, const float *__restrict__ a,
                    const float *__restrict__ k, const float *__restrict__ mean, const float *__restrict__ invstd,
                    const float *__restrict__ ga, const float *__restrict__ mg, const float *__restrict__ mgx,
                    float *__restrict__ dx, int C, long long S, int relu)
{
    const int c = blockIdx.y, b = blockIdx.z;
    const size_t base = ((size_t)b * C + c) * S;
    const float ac = __ldg(a + c), kc = __ldg(k + c), mc = __ldg(mean + c), ic = __ldg(invstd + c);
    const float gac = __ldg(ga + c), mgc = __ldg(mg + c), mgxc = __ldg(mgx + c);
    const long long i0 = (long long)blockIdx.x * BN_CHUNK, i1 = min(i0 + BN_CHUNK, S);
    for (long long i = i0 + threadIdx.x; i < i1; i += BN_THREADS) {
        const float xv = __ldg(x + base + i);
        float g = __ldg(dy + base + i);
        if (relu && !(fmaf(xv, ac, kc) > 0.f)) g = 0.f;
        dx[base + i] = gac * (g - mgc - (xv - mc) * ic * mgxc);
    }
}

static bool bn_grid(dim3 &grid, int B, int C, long long S)
{
    const long long chunks = (S + BN_CHUNK - 1) / BN_CHUNK;
    if (chunks > 2147483647LL || C > 65535 || B > 65535) return false;
    grid = dim3((unsigned)chunks, (unsigned)C, (unsigned)B);
    return true;
}

}  // namespace mvs

using namespace mvs;

extern "C" int mvs_bn_stats(const float *x, double *sums, int B, int C, int64_t S, void *stream)
{
    if (B == 0 || C == 0 || S == 0) return MVS_OK;
    MVS_REQUIRE(x && sums && B > 0 && C > 0 && S > 0, "bad arguments");
    dim3 grid;
    MVS_REQUIRE(bn_grid(grid, B, C, S), "extents exceed the grid limits");
    bn_stats_kernel<<<grid, BN_THREADS, 0, (cudaStream_t)stream>>>(x, sums, C, S);
    return check_launch("mvs_bn_stats");
}

extern "C" int mvs_bn_apply(const float *x, const float *a, const float *k, float *y, int B, int C, int64_t S, int relu, void *stream)
{
    if (B == 0 || C == 0 || S == 0) return MVS_OK;
    MVS_REQUIRE(x && a && k && y && B > 0 && C > 0 && S > 0, "bad arguments");
    dim3 grid;
    MVS_REQUIRE(bn_grid(grid, B, C, S), "extents exceed the grid limits");
    bn_apply_kernel<<<grid, BN_THREADS, 0, (cudaStream_t)stream>>>(x, a, k, y, C, S, relu);
    return check_launch("mvs_bn_apply");
}

extern "C" int mvs_bn_bwd_stats(const float *x, const float *dy, const float *a, const float *k, const float *mean,
                                const float *invstd, double *sums, int B, int C, int64_t S, int relu, void *stream)
{
    if (B == 0 || C == 0 || S == 0) return MVS_OK;
    MVS_REQUIRE(x && dy && a && k && mean && invstd && sums && B > 0 && C > 0 && S > 0, "bad arguments");
    dim3 grid;
    MVS_REQUIRE(bn_grid(grid, B, C, S), "extents exceed the grid limits");
    bn_bwd_stats_kernel<<<grid, BN_THREADS, 0, (cudaStream_t)stream>>>(x, dy, a, k, mean, invstd, sums, C, S, relu);
    return check_launch("mvs_bn_bwd_stats");
}

extern "C" int mvs_bn_bwd_apply(const float *x, const float *dy, const float *a, const float *k, const float *mean,
                                const float *invstd, const float *ga, const float *mg, const float *mgx, float *dx, int B, int C,
                                int64_t S, int relu, void *stream)
{
    if (B == 0 || C == 0 || S == 0) return MVS_OK;
    MVS_REQUIRE(x && dy && a && k && mean && invstd && ga && mg && mgx && dx && B > 0 && C > 0 && S > 0, "bad arguments");
    dim3 grid;
    MVS_REQUIRE(bn_grid(grid, B, C, S), "extents exceed the grid limits");
    bn_bwd_apply_kernel<<<grid, BN_THREADS, 0, (cudaStream_t)stream>>>(x, dy, a, k, mean, invstd, ga, mg, mgx, dx, C, S, relu);
    return check_launch("mvs_bn_bwd_apply");
}
