// Weight gradient of the 3x3x3 convolutions (training step, BASELINE configs[3]): strict fp32.
//   conv      y[co,o] = sum_{ci,k} w[co,ci,k] x[ci, s*o + k - 1]        gw[co,ci,k] = sum_{n,o} gy[n,co,o] x[n,ci, s*o + k - 1]
//   transposed y[co, s*i + k - 1] += w[ci,co,k] x[ci,i]                   gw[ci,co,k] = sum_{n,i} x[n,ci,i] gy[n,co, s*i + k - 1]
// Both are G[a][b][k] = sum_{n,p} A[n,a,p] * T[n,b, s*p + k - 1] with (A, T) = (gy, x) or (x, gy): one kernel.
// Replaces the wgrad half of autograd through nn.Conv3d / nn.ConvTranspose3d (CasMVSNet/models/module.py:137,180,
// MVSNet/models/module.py:29, CVP-MVSNet/models/net.py:56-76) under loss.backward() (CasMVSNet/train.py:165-170).
//
// A CTA owns an (8 NA) x 8 block of (a, b) channel pairs (NA = 1, 2 or 4 a-channels per thread) and a contiguous chunk of A's (n, d, h) rows; it walks the rows in
// 64-position W segments, staging the A segment and the nine (kd, kh) T rows (with their kw halo) in shared memory.
// Thread = (pair, lane of 4): 27 accumulators in registers over positions w = lane, lane + 4, ...; at the end the four
// lanes are reduced by shuffles and the CTA adds its 64 x 27 partial sums to gw with atomicAdd (the accumulation order
// over CTAs is therefore not deterministic in the last bits, like cuDNN's default wgrad algorithms).
#include <cstdlib>
#include <string>

#include "common.cuh"

namespace mvs {

constexpr int WG_SEG = 64;

// NA = a-channels per thread (a_loc, a_loc + 8, ...): the CTA owns an (8 NA) x 8 block of pairs, and every staged T row
// segment / shared-memory T load feeds NA times the FMAs (the 8 x 8 version moved 21 KB through L2 per 110 k FMAs and was
// bound by that)
// ONE_A: Ca == 1 (the `prob` layers' grad_y): the eight a-slots of the pair block would all but one multiply zeros, so the
// slots 0..3 split the segment's sixteen 4-position chunks among themselves instead (one trip each, four times less work)
template <int S, int NA, bool ONE_A = false>
__global__ void __launch_bounds__(256)
conv3d_wgrad_kernel(const float *__restrict__ A, const float *__restrict__ T, float *__restrict__ G, int N, int Ca, int Cb,
                    int Da, int Ha, int Wa, int Dt, int Ht, int Wt, int rows_per_cta)
{
    constexpr int TW = S * WG_SEG + 2;                 // staged T positions per row: s*w + kw - 1 for w < 64, kw < 3
    constexpr int TP = S == 1 ? 72 : 136;              // padded row pitch (pitch mod 32 == 8: conflict-free for 8 b x 4 lanes)
    // two stages: the segment q + 1 is fetched with cp.async while segment q is multiplied (staged synchronously, the loads'
    // latency was exposed once per segment -- about half of the kernel's time)
    extern __shared__ __align__(16) float smem_wg[];
    float (*sA)[8 * NA][WG_SEG] = reinterpret_cast<float (*)[8 * NA][WG_SEG]>(smem_wg);
    float (*sT)[8][9][TP] = reinterpret_cast<float (*)[8][9][TP]>(smem_wg + 2 * 8 * NA * WG_SEG);
    const int tid = threadIdx.x, pair = tid >> 2, lane = tid & 3;
    const int a_loc = pair >> 3, b_loc = pair & 7;
    const int a0 = blockIdx.y * 8 * NA, b0 = blockIdx.z * 8;
    const long long total_rows = (long long)N * Da * Ha;
    const long long row_begin = (long long)blockIdx.x * rows_per_cta;
    const long long row_end = min(row_begin + rows_per_cta, total_rows);
    float acc[NA][27];
#pragma unroll
    for (int u = 0; u < NA; ++u)
#pragma unroll
        for (int k = 0; k < 27; ++k) acc[u][k] = 0.f;
    const size_t avol = (size_t)Da * Ha * Wa, tvol = (size_t)Dt * Ht * Wt;

    const int nseg = (Wa + WG_SEG - 1) / WG_SEG;
    const long long n_q = (row_end - row_begin) * nseg;        // (row, segment) work items of this CTA
    auto stage = [&](int buf, long long q) {
        const long long row = row_begin + q / nseg;
        const int w0 = (int)(q % nseg) * WG_SEG;
        const int h = (int)(row % Ha);
        const int d = (int)((row / Ha) % Da);
        const int n = (int)(row / ((long long)Ha * Da));
        // A segment: 8 NA channels x 64 positions (zero beyond the channels / the row end: 0-byte copies zero-fill)
        for (int i = tid; i < 8 * NA * WG_SEG; i += 256) {
            const int c = i / WG_SEG, w = i % WG_SEG;
            const int a = a0 + c;
            const bool ok = a < Ca && w0 + w < Wa;
            const float *src = ok ? A + ((size_t)n * Ca + a) * avol + ((size_t)d * Ha + h) * Wa + w0 + w : A;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n"
                         :: "r"((uint32_t)__cvta_generic_to_shared(&sA[buf][c][w])), "l"(src), "r"(ok ? 4u : 0u) : "memory");
        }
        // T rows: 8 channels x 9 (kd, kh) rows x (S*64 + 2) positions starting at S*w0 - 1
        for (int i = tid; i < 8 * 9 * TW; i += 256) {
            const int j = i % TW, r = (i / TW) % 9, c = i / (TW * 9);
            const int b = b0 + c;
            const int td = S * d + r / 3 - 1, th = S * h + r % 3 - 1, tw = S * w0 - 1 + j;
            const bool ok = b < Cb && td >= 0 && td < Dt && th >= 0 && th < Ht && tw >= 0 && tw < Wt;
            const float *src = ok ? T + ((size_t)n * Cb + b) * tvol + ((size_t)td * Ht + th) * Wt + tw : T;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n"
                         :: "r"((uint32_t)__cvta_generic_to_shared(&sT[buf][c][r][j])), "l"(src), "r"(ok ? 4u : 0u) : "memory");
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    if (n_q > 0) stage(0, 0);
    for (long long q = 0; q < n_q; ++q) {
        {
            const int buf = (int)(q & 1);
            if (q + 1 < n_q) {
                stage(buf ^ 1, q + 1);
                asm volatile("cp.async.wait_group 1;\n" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;\n" ::: "memory");
            }
            __syncthreads();
            // four consecutive positions per lane and trip: the A values and each T row segment come as 16-byte shared-memory
            // loads (19 / 28 loads per 108 FMAs; one scalar load per FMA made the kernel shared-memory bound).  Positions
            // beyond the row end carry A = 0 (staged above), so no bound is needed.
            for (int w = ONE_A ? (a_loc * 4 + lane) * 4 : lane * 4; w < WG_SEG; w += ONE_A ? WG_SEG : 16) {
                float av[NA][4];
#pragma unroll
                for (int u = 0; u < NA; ++u) {
                    const float4 a4 = *reinterpret_cast<const float4 *>(&sA[buf][ONE_A ? 0 : a_loc + 8 * u][w]);
                    av[u][0] = a4.x; av[u][1] = a4.y; av[u][2] = a4.z; av[u][3] = a4.w;
                }
#pragma unroll
                for (int r = 0; r < 9; ++r) {
                    constexpr int NV = (S * 3 + 3 + 3) / 4;                     // float4 loads covering offsets 0 .. S*3 + 2
                    float tt[NV * 4];
                    const float4 *t4 = reinterpret_cast<const float4 *>(&sT[buf][b_loc][r][S * w]);
#pragma unroll
                    for (int v = 0; v < NV; ++v) {
                        const float4 q4 = t4[v];
                        tt[v * 4 + 0] = q4.x; tt[v * 4 + 1] = q4.y; tt[v * 4 + 2] = q4.z; tt[v * 4 + 3] = q4.w;
                    }
#pragma unroll
                    for (int u = 0; u < NA; ++u)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            acc[u][r * 3 + 0] = fmaf(av[u][j], tt[S * j + 0], acc[u][r * 3 + 0]);
                            acc[u][r * 3 + 1] = fmaf(av[u][j], tt[S * j + 1], acc[u][r * 3 + 1]);
                            acc[u][r * 3 + 2] = fmaf(av[u][j], tt[S * j + 2], acc[u][r * 3 + 2]);
                        }
                }
            }
            __syncthreads();            // everyone is done with this stage before the next iteration refills the other one
        }
    }
#pragma unroll
    for (int u = 0; u < NA; ++u) {
#pragma unroll
        for (int k = 0; k < 27; ++k) {
            float v = acc[u][k];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            acc[u][k] = v;
        }
        const int a = ONE_A ? a0 : a0 + a_loc + 8 * u, b = b0 + b_loc;
        if (lane == 0 && a < Ca && b < Cb && (!ONE_A || a_loc < 4)) {
            float *g = G + ((size_t)a * Cb + b) * 27;
#pragma unroll
            for (int k = 0; k < 27; ++k) atomicAdd(g + k, acc[u][k]);
        }
    }
}

}  // namespace mvs

using namespace mvs;

// gw must be zero-initialised by the caller (the kernel accumulates with atomicAdd).
extern "C" int mvs_conv3d_wgrad(const float *x, const float *grad_y, float *gw, int B, int Cin, int Cout, int D, int H, int W,
                                int stride, int transposed, void *stream)
{
    if (B == 0 || D == 0 || H == 0 || W == 0) return MVS_OK;
    MVS_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && D > 0 && H > 0 && W > 0, "extents must be positive");
    MVS_REQUIRE(stride == 1 || stride == 2, "stride must be 1 or 2");
    MVS_REQUIRE(x && grad_y && gw, "null pointer");
    int Do, Ho, Wo;
    if (transposed) { Do = D * stride; Ho = H * stride; Wo = W * stride; }
    else { Do = (D - 1) / stride + 1; Ho = (H - 1) / stride + 1; Wo = (W - 1) / stride + 1; }
    // (A, T): conv: (grad_y [Cout, out extents], x [Cin, in extents]); transposed: (x [Cin, in], grad_y [Cout, out])
    const float *A = transposed ? x : grad_y, *T = transposed ? grad_y : x;
    const int Ca = transposed ? Cin : Cout, Cb = transposed ? Cout : Cin;
    const int Da = transposed ? D : Do, Ha = transposed ? H : Ho, Wa = transposed ? W : Wo;
    const int Dt = transposed ? Do : D, Ht = transposed ? Ho : H, Wt = transposed ? Wo : W;
    const long long rows = (long long)B * Da * Ha;
    static const int na_max = getenv("MVS_WGRAD_NA") ? atoi(getenv("MVS_WGRAD_NA")) : 4;      // tuning knob
    const int na = (Ca >= 32 && na_max >= 4) ? 4 : (Ca > 8 && na_max >= 2 ? 2 : 1);           // a-channels per thread
    const int pair_blocks = cdiv(Ca, 8 * na) * cdiv(Cb, 8);
    // ~16 CTAs per SM in total: enough parallelism, few enough CTAs that the final atomics stay cheap
    long long chunks = (16LL * sm_count() + pair_blocks - 1) / pair_blocks;
    if (chunks > rows) chunks = rows;
    if (chunks < 1) chunks = 1;
    const int rows_per_cta = (int)((rows + chunks - 1) / chunks);
    dim3 grid((unsigned)cdiv(rows, rows_per_cta), cdiv(Ca, 8 * na), cdiv(Cb, 8));
    MVS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "too many channel blocks");
    cudaStream_t st = (cudaStream_t)stream;
    auto launch = [&](auto kernel, int S_, int NA_) -> cudaError_t {
        const int tp = S_ == 1 ? 72 : 136;
        const size_t smem = (size_t)(2 * 8 * NA_ * WG_SEG + 2 * 8 * 9 * tp) * sizeof(float);
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kernel<<<grid, 256, smem, st>>>(A, T, gw, B, Ca, Cb, Da, Ha, Wa, Dt, Ht, Wt, rows_per_cta);
        return cudaSuccess;
    };
    cudaError_t e;
    if (Ca == 1) e = stride == 1 ? launch(conv3d_wgrad_kernel<1, 1, true>, 1, 1) : launch(conv3d_wgrad_kernel<2, 1, true>, 2, 1);
    else if (stride == 1) e = na == 4 ? launch(conv3d_wgrad_kernel<1, 4>, 1, 4) : (na == 2 ? launch(conv3d_wgrad_kernel<1, 2>, 1, 2) : launch(conv3d_wgrad_kernel<1, 1>, 1, 1));
    else e = na == 4 ? launch(conv3d_wgrad_kernel<2, 4>, 2, 4) : (na == 2 ? launch(conv3d_wgrad_kernel<2, 2>, 2, 2) : launch(conv3d_wgrad_kernel<2, 1>, 2, 1));
    if (e != cudaSuccess) return fail(MVS_ERR_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    return check_launch("mvs_conv3d_wgrad");
}
