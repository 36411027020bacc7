// Strict fp32 3x3x3 convolution (+ folded BN affine, ReLU, skip) on the reference's NCDHW layout.
// Direct convolution on the fp32 FMA pipe: used when results must track the reference's fp32
// path (TF32/bf16 tensor cores cannot); the bf16 tcgen05 path lives in
// conv3d_umma.cu.  Replaces ConvBnReLU3D / Conv3d / Deconv3d (+ skip adds) of CostRegNet:
//   MVSNet/models/mvsnet.py:55-93, CasMVSNet/models/module.py:115-200,407-438, CVP net.py:52-89.
//
// Thread <-> TW output voxels (consecutive flattened (oh,ow) positions of one output depth slice,
// lanes along w => coalesced), CO_T output channels each.  Weights of the CTA's channel group are
// staged through shared memory in chunks of CI_T input channels and read as broadcast float4s.
#include <cstdlib>

#include "common.cuh"

namespace mvs {

constexpr int CONV_THREADS = 128;
constexpr int CI_T = 8;

enum ConvMode { CONV_S1 = 0, CONV_S2 = 1, DECONV_S1 = 2, DECONV_S2 = 3 };

template <int MODE>
__device__ __forceinline__ bool in_coord(int o, int k, int n, int &i)
{
    if (MODE == CONV_S1) { i = o - 1 + k; }
    else if (MODE == CONV_S2) { i = 2 * o - 1 + k; }
    else if (MODE == DECONV_S1) { i = o + 1 - k; }
    else {                       // o = 2 i - 1 + k
        const int nn = o + 1 - k;
        if (nn & 1) return false;
        i = nn >> 1;
    }
    return i >= 0 && i < n;
}

template <int CO_T, int TW, int MODE>
__global__ void __launch_bounds__(CONV_THREADS)
conv3d_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ scale,
              const float *__restrict__ shift, const float *__restrict__ skip, float *__restrict__ y, int Cin, int Cout,
              int D, int H, int W, int Do, int Ho, int Wo, int relu)
{
    __shared__ __align__(16) float ws[CI_T][27][CO_T];
    const int co_groups = (Cout + CO_T - 1) / CO_T;
    const int b = blockIdx.z / co_groups, cg = blockIdx.z % co_groups;
    const int od = blockIdx.y;
    const int hw_o = Ho * Wo;
    const size_t ivol = (size_t)D * H * W, ovol = (size_t)Do * hw_o;

    int oh[TW], ow[TW];
    bool live[TW];
#pragma unroll
    for (int t = 0; t < TW; ++t) {
        const int p = (blockIdx.x * TW + t) * CONV_THREADS + threadIdx.x;
        live[t] = p < hw_o;
        oh[t] = live[t] ? p / Wo : 0;
        ow[t] = live[t] ? p % Wo : 0;
    }
    float acc[TW][CO_T];
#pragma unroll
    for (int t = 0; t < TW; ++t)
#pragma unroll
        for (int c = 0; c < CO_T; ++c) acc[t][c] = 0.f;

    const float *xb = x + (size_t)b * Cin * ivol;
    constexpr bool transposed = (MODE == DECONV_S1 || MODE == DECONV_S2);

    for (int ci0 = 0; ci0 < Cin; ci0 += CI_T) {
        __syncthreads();
        for (int e = threadIdx.x; e < CI_T * 27 * CO_T; e += CONV_THREADS) {
            const int c = e % CO_T, tap = (e / CO_T) % 27, ci = e / (CO_T * 27);
            const int gci = ci0 + ci, gco = cg * CO_T + c;
            float v = 0.f;
            if (gci < Cin && gco < Cout)
                v = transposed ? __ldg(w + ((size_t)gci * Cout + gco) * 27 + tap)
                               : __ldg(w + ((size_t)gco * Cin + gci) * 27 + tap);
            ws[ci][tap][c] = v;
        }
        __syncthreads();
        const int ci_n = min(CI_T, Cin - ci0);
        for (int ci = 0; ci < ci_n; ++ci) {
            const float *xc = xb + (size_t)(ci0 + ci) * ivol;
#pragma unroll
            for (int kd = 0; kd < 3; ++kd) {
                int id;
                if (!in_coord<MODE>(od, kd, D, id)) continue;       // block-uniform
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        float xv[TW];
#pragma unroll
                        for (int t = 0; t < TW; ++t) {
                            int ih, iw;
                            const bool ok = live[t] & in_coord<MODE>(oh[t], kh, H, ih) & in_coord<MODE>(ow[t], kw, W, iw);
                            xv[t] = ok ? __ldg(xc + ((size_t)id * H + ih) * W + iw) : 0.f;
                        }
                        const float4 *wp = reinterpret_cast<const float4 *>(&ws[ci][(kd * 3 + kh) * 3 + kw][0]);
#pragma unroll
                        for (int c4 = 0; c4 < CO_T / 4; ++c4) {
                            const float4 wv = wp[c4];
#pragma unroll
                            for (int t = 0; t < TW; ++t) {
                                acc[t][c4 * 4 + 0] = fmaf(xv[t], wv.x, acc[t][c4 * 4 + 0]);
                                acc[t][c4 * 4 + 1] = fmaf(xv[t], wv.y, acc[t][c4 * 4 + 1]);
                                acc[t][c4 * 4 + 2] = fmaf(xv[t], wv.z, acc[t][c4 * 4 + 2]);
                                acc[t][c4 * 4 + 3] = fmaf(xv[t], wv.w, acc[t][c4 * 4 + 3]);
                            }
                        }
                    }
            }
        }
    }
#pragma unroll
    for (int t = 0; t < TW; ++t) {
        if (!live[t]) continue;
#pragma unroll
        for (int c = 0; c < CO_T; ++c) {
            const int gco = cg * CO_T + c;
            if (gco >= Cout) break;
            float v = acc[t][c];
            const float sc = scale ? __ldg(scale + gco) : 1.f, sh = shift ? __ldg(shift + gco) : 0.f;
            v = fmaf(v, sc, sh);
            if (relu) v = fmaxf(v, 0.f);
            const size_t o = ((size_t)b * Cout + gco) * ovol + (size_t)od * hw_o + (size_t)oh[t] * Wo + ow[t];
            if (skip) v = __ldg(skip + o) + v;
            y[o] = v;
        }
    }
}

// Stride-1 layers (conv and flipped-tap transposed conv): a thread owns TH consecutive output ROWS at one w (lanes along w:
// every load is a coalesced row segment) and CO_T output channels.  For a (ci, kd, kw) the TH + 2 input rows it needs are
// loaded once and feed all three kh taps of the TH outputs: (TH + 2) loads per 3 TH CO_T FMAs instead of 3 TH -- the generic
// kernel above is bound by its L1 requests (one load per CO_T FMAs), not by the FMA pipe.
template <int CO_T, int TH, bool FLIP>
__global__ void __launch_bounds__(CONV_THREADS)
conv3d_s1_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ scale,
                 const float *__restrict__ shift, const float *__restrict__ skip, float *__restrict__ y, int Cin, int Cout,
                 int D, int H, int W, int relu)
{
    __shared__ __align__(16) float ws[CI_T][27][CO_T];
    const int co_groups = (Cout + CO_T - 1) / CO_T;
    const int b = blockIdx.z / co_groups, cg = blockIdx.z % co_groups;
    const int od = blockIdx.y;
    // threads run over the flattened (row block, w) positions, so narrow maps (W < 128 at the low-resolution levels) still
    // fill their warps
    const long long p = (long long)blockIdx.x * CONV_THREADS + threadIdx.x;
    const int ow = (int)(p % W);
    const int oh0 = (int)(p / W) * TH;
    const size_t hw = (size_t)H * W, vol = (size_t)D * hw;
    const bool w_live = oh0 < H;

    float acc[TH][CO_T];
#pragma unroll
    for (int t = 0; t < TH; ++t)
#pragma unroll
        for (int c = 0; c < CO_T; ++c) acc[t][c] = 0.f;
    const float *xb = x + (size_t)b * Cin * vol;

    for (int ci0 = 0; ci0 < Cin; ci0 += CI_T) {
        __syncthreads();
        for (int e = threadIdx.x; e < CI_T * 27 * CO_T; e += CONV_THREADS) {
            const int c = e % CO_T, tap = (e / CO_T) % 27, ci = e / (CO_T * 27);
            const int gci = ci0 + ci, gco = cg * CO_T + c;
            float v = 0.f;
            if (gci < Cin && gco < Cout)
                v = FLIP ? __ldg(w + ((size_t)gci * Cout + gco) * 27 + tap) : __ldg(w + ((size_t)gco * Cin + gci) * 27 + tap);
            ws[ci][tap][c] = v;
        }
        __syncthreads();
        const int ci_n = min(CI_T, Cin - ci0);
        for (int ci = 0; ci < ci_n; ++ci) {
            const float *xc = xb + (size_t)(ci0 + ci) * vol;
#pragma unroll
            for (int kd = 0; kd < 3; ++kd) {
                const int id = FLIP ? od + 1 - kd : od - 1 + kd;
                if (id < 0 || id >= D) continue;                    // block-uniform
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int iw = FLIP ? ow + 1 - kw : ow - 1 + kw;
                    const bool col_ok = w_live && iw >= 0 && iw < W;
                    float xv[TH + 2];                                // input rows oh0 - 1 .. oh0 + TH at column iw
#pragma unroll
                    for (int j = 0; j < TH + 2; ++j) {
                        const int ih = oh0 - 1 + j;
                        xv[j] = (col_ok && ih >= 0 && ih < H) ? __ldg(xc + (size_t)id * hw + (size_t)ih * W + iw) : 0.f;
                    }
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) {
                        const float4 *wp = reinterpret_cast<const float4 *>(&ws[ci][(kd * 3 + kh) * 3 + kw][0]);
#pragma unroll
                        for (int c4 = 0; c4 < CO_T / 4; ++c4) {
                            const float4 wv = wp[c4];
#pragma unroll
                            for (int t = 0; t < TH; ++t) {
                                // conv: ih = oh - 1 + kh -> row slot t + kh; flipped taps: ih = oh + 1 - kh -> slot t + 2 - kh
                                const float xin = xv[FLIP ? t + 2 - kh : t + kh];
                                acc[t][c4 * 4 + 0] = fmaf(xin, wv.x, acc[t][c4 * 4 + 0]);
                                acc[t][c4 * 4 + 1] = fmaf(xin, wv.y, acc[t][c4 * 4 + 1]);
                                acc[t][c4 * 4 + 2] = fmaf(xin, wv.z, acc[t][c4 * 4 + 2]);
                                acc[t][c4 * 4 + 3] = fmaf(xin, wv.w, acc[t][c4 * 4 + 3]);
                            }
                        }
                    }
                }
            }
        }
    }
    if (!w_live) return;
#pragma unroll
    for (int t = 0; t < TH; ++t) {
        const int oh = oh0 + t;
        if (oh >= H) break;
#pragma unroll
        for (int c = 0; c < CO_T; ++c) {
            const int gco = cg * CO_T + c;
            if (gco >= Cout) break;
            float v = acc[t][c];
            const float sc = scale ? __ldg(scale + gco) : 1.f, sh = shift ? __ldg(shift + gco) : 0.f;
            v = fmaf(v, sc, sh);
            if (relu) v = fmaxf(v, 0.f);
            const size_t o = ((size_t)b * Cout + gco) * vol + (size_t)od * hw + (size_t)oh * W + ow;
            if (skip) v = __ldg(skip + o) + v;
            y[o] = v;
        }
    }
}

template <int CO_T, int TH>
static void launch_conv_s1(bool flip, cudaStream_t st, const float *x, const float *w, const float *scale, const float *shift,
                           const float *skip, float *y, int B, int Cin, int Cout, int D, int H, int W, int relu)
{
    dim3 grid((unsigned)cdiv((long long)W * cdiv(H, TH), CONV_THREADS), (unsigned)D, (unsigned)(B * cdiv(Cout, CO_T)));
    if (flip) conv3d_s1_kernel<CO_T, TH, true><<<grid, CONV_THREADS, 0, st>>>(x, w, scale, shift, skip, y, Cin, Cout, D, H, W, relu);
    else conv3d_s1_kernel<CO_T, TH, false><<<grid, CONV_THREADS, 0, st>>>(x, w, scale, shift, skip, y, Cin, Cout, D, H, W, relu);
}

template <int CO_T, int TW>
static void launch_conv(int mode, dim3 grid, cudaStream_t st, const float *x, const float *w, const float *scale,
                        const float *shift, const float *skip, float *y, int Cin, int Cout, int D, int H, int W, int Do,
                        int Ho, int Wo, int relu)
{
    switch (mode) {
    case CONV_S1: conv3d_kernel<CO_T, TW, CONV_S1><<<grid, CONV_THREADS, 0, st>>>(x, w, scale, shift, skip, y, Cin, Cout, D, H, W, Do, Ho, Wo, relu); break;
    case CONV_S2: conv3d_kernel<CO_T, TW, CONV_S2><<<grid, CONV_THREADS, 0, st>>>(x, w, scale, shift, skip, y, Cin, Cout, D, H, W, Do, Ho, Wo, relu); break;
    case DECONV_S1: conv3d_kernel<CO_T, TW, DECONV_S1><<<grid, CONV_THREADS, 0, st>>>(x, w, scale, shift, skip, y, Cin, Cout, D, H, W, Do, Ho, Wo, relu); break;
    default: conv3d_kernel<CO_T, TW, DECONV_S2><<<grid, CONV_THREADS, 0, st>>>(x, w, scale, shift, skip, y, Cin, Cout, D, H, W, Do, Ho, Wo, relu); break;
    }
}

}  // namespace mvs

using namespace mvs;

extern "C" int mvs_conv3d_fwd(const float *x, const float *w, const float *scale, const float *shift, const float *skip,
                              float *y, int B, int Cin, int Cout, int D, int H, int W, int stride, int transposed,
                              int flags, void *stream)
{
    if (B == 0 || D == 0 || H == 0 || W == 0) return MVS_OK;
    MVS_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && D > 0 && H > 0 && W > 0, "extents must be positive");
    MVS_REQUIRE(stride == 1 || stride == 2, "stride must be 1 or 2");
    MVS_REQUIRE(x && w && y, "null pointer");
    int Do, Ho, Wo;
    if (transposed) { Do = D * stride; Ho = H * stride; Wo = W * stride; }
    else { Do = (D - 1) / stride + 1; Ho = (H - 1) / stride + 1; Wo = (W - 1) / stride + 1; }
    const int mode = transposed ? (stride == 1 ? DECONV_S1 : DECONV_S2) : (stride == 1 ? CONV_S1 : CONV_S2);
    const int relu = (flags & MVS_RELU) ? 1 : 0;
    MVS_REQUIRE(Do <= 65535, "output depth exceeds grid.y");
    cudaStream_t st = (cudaStream_t)stream;
    const long long hw_o = (long long)Ho * Wo;
    static const bool s1_tiled = !(getenv("MVS_STRICT_S1") && atoi(getenv("MVS_STRICT_S1")) == 0);       // A/B knob
    if (stride == 1 && s1_tiled) {
        const bool flip = transposed != 0;
        MVS_REQUIRE((long long)B * cdiv(Cout, 4) <= 65535 && (long long)cdiv(W, CONV_THREADS) * cdiv(H, 4) < (1ll << 31),
                    "grid too large");
        if (Cout >= 16) launch_conv_s1<16, 4>(flip, st, x, w, scale, shift, skip, y, B, Cin, Cout, D, H, W, relu);
        else if (Cout > 4) launch_conv_s1<8, 4>(flip, st, x, w, scale, shift, skip, y, B, Cin, Cout, D, H, W, relu);
        else launch_conv_s1<4, 4>(flip, st, x, w, scale, shift, skip, y, B, Cin, Cout, D, H, W, relu);
        return check_launch("mvs_conv3d_fwd");
    }
    if (Cout >= 16) {
        constexpr int CO_T = 16, TW = 2;
        const long long gz = (long long)B * cdiv(Cout, CO_T);
        MVS_REQUIRE(gz <= 65535, "B * channel groups exceeds grid.z");
        dim3 grid(cdiv(hw_o, CONV_THREADS * TW), Do, (unsigned)gz);
        launch_conv<CO_T, TW>(mode, grid, st, x, w, scale, shift, skip, y, Cin, Cout, D, H, W, Do, Ho, Wo, relu);
    } else if (Cout > 4) {
        constexpr int CO_T = 8, TW = 2;
        const long long gz = (long long)B * cdiv(Cout, CO_T);
        MVS_REQUIRE(gz <= 65535, "B * channel groups exceeds grid.z");
        dim3 grid(cdiv(hw_o, CONV_THREADS * TW), Do, (unsigned)gz);
        launch_conv<CO_T, TW>(mode, grid, st, x, w, scale, shift, skip, y, Cin, Cout, D, H, W, Do, Ho, Wo, relu);
    } else {
        constexpr int CO_T = 4, TW = 2;
        const long long gz = (long long)B * cdiv(Cout, CO_T);
        MVS_REQUIRE(gz <= 65535, "B * channel groups exceeds grid.z");
        dim3 grid(cdiv(hw_o, CONV_THREADS * TW), Do, (unsigned)gz);
        launch_conv<CO_T, TW>(mode, grid, st, x, w, scale, shift, skip, y, Cin, Cout, D, H, W, Do, Ho, Wo, relu);
    }
    return check_launch("mvs_conv3d_fwd");
}
