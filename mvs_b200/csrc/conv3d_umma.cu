// bf16 3x3x3 convolution on the 5th-generation tensor cores: tcgen05.mma (UMMA) implicit GEMM with
// the accumulators in TMEM, operands read straight from shared memory in the C8 layout.
// Replaces the cuDNN conv3d / conv_transpose3d + BatchNorm3d + ReLU + skip-add launches of
// CostRegNet (MVSNet/models/mvsnet.py:55-93, CasMVSNet/models/module.py:115-200,407-438,
// CVP-MVSNet/models/net.py:52-89).
//
// Why C8 makes this an implicit GEMM without im2col: in [B][C/8][D][H][W][8] bf16, eight consecutive
// w-voxels x eight channels are 128 contiguous bytes -- exactly one UMMA "core matrix" of the
// K-major, no-swizzle canonical layout ((8,n),2):((16B,SBO),LBO).  So for a tile of 128 consecutive
// w-voxels (the MMA M dimension) staged once in shared memory with its halo, the A operand of tap
// (kd,kh,kw) is the SAME bytes at a different start address: start = row(kd,kh) + kw*16 B, SBO = 128 B,
// LBO = distance to the next 8-channel block.  The 27 taps x Cin/16 k-steps accumulate into one TMEM
// tile [128 x N] (N = Cout padded to 16).  Cin = 8 layers pair two taps per K=16 step (LBO = 16 B);
// stride-2 layers stage even / odd columns separately; transposed convs run as 8 output-parity
// classes over input-resolution rows.  All of that is expressed as a per-layer table of MMA ops
// built on the host (ConvPlan), so there is one kernel.
//
// Two kernels (template <TM, MC>), both warp-specialised and mbarrier-synchronised; a CTA owns one M tile of 128
// w-positions x ht rows x a chunk of the "step" axis (real H; the rows tile real D), weights staged once per CTA.
//
// (1) conv3d_umma_kernel<true, .>  -- "T-merged" mode, every stride-1 layer (conv0/2/4/6, prob; CVP's stride-1 layers):
//   Measured cost model (tools/micro/umma_rate.cu): an M128 x N x K16 MMA costs max(N/2, 32 + N/4) clk in the tensor pipe
//   and ~85 clk of one issuing warp, almost independent of N for N <= 80 -- and CostRegNet has Cout = 1 / 8 / 16.  So ONE
//   MMA per (staged input row, kw group, Cin pair) carries all nine (row tap, step tap) weight blocks along N: the row
//   taps accumulate inside the MMA (B sub-block 2 - (i - lo) lines the taps up with output rows lo..hi), the three step
//   taps land in column blocks [row][t][n] of a PER-SLAB TMEM buffer (ring of 4), and the epilogue adds the partials of
//   slabs s, s+1, s+2.  Every slab is read once.
//     warps 4-7,10,11 producers : one 1-D bulk copy (cp.async.bulk + mbarrier complete_tx) per staged line; halo parts
//                                 are zero-filled with ordinary stores on border CTAs only
//     warps 8-9       issuers   : alternate slabs (ring and buffer count even => one waiting warp per barrier); op table
//                                 in kernel-parameter space, descriptors built in the uniform datapath
//     warps 0-3,12-15 epilogue  : two groups split the rows; three tcgen05.ld per 8-channel block -> sum -> folded BN +
//                                 ReLU (+ skip) -> C8 bf16 store / fp32 logits
// (2) conv3d_umma_kernel<false, .> -- table-driven modes for stride-2 convs (even / odd staged arrays) and transposed
//   stride-2 convs (8 output-parity accumulators per row):
//     warps 4-7 producers (cp.async 16 B with zero-fill, slab-invariant index math in a shared-memory line table),
//     warps 8-11 issuers (each owns a slice of the accumulators), warps 0-3 epilogue (skip vectors prefetched before the
//     TMEM wait, four accumulators per batch); TMEM double-buffered per step.  <false, 2>: skip layers, 2 CTAs/SM;
//     <false, 3>: the others, 3 CTAs/SM.
//
// Round-2 additions (DESIGN.md 4.2 has the measurements behind each):
//   * `prob` (Cout = 1) packs one TMEM column per (row tap, step tap) -- N = 16 / 32 MMAs, 8 columns per output row, up to 15
//     rows per step; Cin = 8 layers pair input rows (three K = 16 steps per two rows, the shared step updates four rows);
//   * W-de-interleaved skip tensors (MVS_Y_DW / MVS_X_DW / MVS_SKIP_DW): conv0 / 2 / 4 write [even | odd] columns, the stride-2
//     layers stage contiguous runs, the transposed layers read one run per output parity and store column pairs as 32 bytes;
//   * 2D layers (the feature extractors): MVS_KD1 (images stacked on the depth axis: only the centre depth tap's MMAs) and the
//     flat 2D mode MVS_FLAT2D = conv3d_umma_kernel<true, 4 | 5> (a slab is a block of image rows, no step-tap partials, 15
//     rows per step; <true, 5> adds a pixel-shuffled half-resolution skip operand, MVS_SKIP_PS);
//   * programmatic dependent launch: everything before `griddepcontrol.wait` (TMEM allocation, barriers, weight staging, index
//     tables) runs under the tail of the previous kernel in the stream.
//
// Lessons that shaped the code (DESIGN.md 4.2 has the measurements): every role is ONE warp running dependent scalar
// code (~5 clk per instruction), so per-step bookkeeping -- runtime divisions, re-materialised 64-bit index arithmetic,
// dynamically indexed parameter reads -- is what the pipeline waits for, not bytes or FLOPs; issuing tcgen05.mma from
// inside `if (lane == 0)` makes the compiler wrap every UTCHMMA in an R2UR waterfall loop (use elect.sync + uniform
// operands); a UBLKCP costs its issuing warp 300-450 clk (spread the lines over warps).
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace mvs {

constexpr int UM_EPI_THREADS = 128;
constexpr int UM_PROD_THREADS = 128;
constexpr int UM_PROD_WARPS_TM = 6;      // T-merged kernel: warps 4-7 + the two issuer warps it does not use (10, 11)
constexpr int UM_MAX_ISSUERS = 4;
constexpr int UM_THREADS = UM_EPI_THREADS + UM_PROD_THREADS + 32 * UM_MAX_ISSUERS;
constexpr int UM_THREADS_TM = UM_THREADS + UM_EPI_THREADS;     // T-merged kernel: + a second epilogue group (warps 12-15)
constexpr int UM_COLS = 132;        // staged columns per row (128 + halo + pairing pad)
constexpr int UM_MAX_OPS = 224;
constexpr int UM_MAX_ACC = 16;
constexpr int UM_MAX_KSTEPS = 112;
constexpr int UM_MAX_RING = 8;
constexpr int UM_MAX_LINES = 288;   // staged lines per slab (rows x channel blocks x arrays)
#ifndef UM_MIN_CTAS
#define UM_MIN_CTAS 2
#endif
#ifndef UM_TBUFS_LOG2
#define UM_TBUFS_LOG2 2
#endif
constexpr int UM_TBUFS = 1 << UM_TBUFS_LOG2;    // per-slab accumulator buffers of the T-merged mode (a power of two)

enum UmMode { UM_CONV_S1 = 0, UM_CONV_S2 = 1, UM_DECONV_S2 = 2 };

struct MmaOp {                // offsets in 16-byte units
    uint16_t a_off;           // within a depth slab
    uint16_t b_off;           // within the packed weights of this Cout tile
    uint16_t a_lbo;           // A leading-byte-offset (distance between the two 8-channel K chunks)
    uint8_t acc;              // accumulator index
    uint8_t rd_first;         // depth slab of the step (0..2)
};

struct AccOut {               // where accumulator `acc` lands: od = od_mul*step + dd, oh = oh_mul*(h0+th) + dh
    int8_t th, dd, dh, wadd;  // ow = w_mul*m + wadd
};

struct KStepSrc {             // weight source of the two K chunks of a k-step: tap index (-1: zeros), cin chunk
    int8_t tap[2];
    int8_t chunk[2];
};

struct ConvPlan {
    int B, D, H, W, Do, Ho, Wo;
    int cin_chunks, cout, cout_chunks, n, cout_tiles;
    int mode, ht, rh, arr, rd, ring;
    int slab_units, weight_units, tmem_cols, acc_cols;
    // Internally the kernel walks a "step" axis (slabs, named d below) and tiles a "row" axis (named h).
    // swap = 1 maps step -> real H and row -> real D (every stage then has tens of steps per CTA even
    // when D is 1..8); D, H, Do, Ho above are the INTERNAL extents, Dr/Hr/Dor/Hor the real ones.
    int swap, Dr, Hr, Dor, Hor, row_blocks, steps_per_cta;
    int h_mul, h_base;        // h_in(r) = h_mul * h0 + h_base + r          (h0 = ht * row block)
    int d_mul, d_base;        // d_in(slab i) = d_base + i ; step s uses slabs d_mul*s + {0..rd-1}
    int w_step, w_base[2];    // w_in(arr, col) = w_step * (m0 + col) + w_base[arr]
    int od_mul, oh_mul, w_mul;
    int n_ops, n_acc, steps;
    int relu, out_f32, has_skip;
    int f16;                  // MVS_ACT_F16: activations / weights / output are fp16 instead of bf16 (same 16-bit C8 layout)
    // "DW" = W de-interleaved: column w of a row sits at (w & 1) * ceil(W / 2) + (w >> 1), i.e. [even columns | odd columns].
    // A stride-1 layer writes it (y_dw) so that the stride-2 layer reading it (x_dw) stages its even / odd arrays from
    // CONTIGUOUS bytes, and the transposed layer adding it as skip (skip_dw) reads one contiguous run per output parity.
    int x_dw, y_dw, skip_dw;
    int n_issuers, zero_units;                      // zero_units: 16 B units of the all-zero B block (merged mode)
    int merged;                                     // stride-1 kh-merged mode: 2 issuers alternate depth steps, epilogue frees slabs
    // T-merged mode (stride 1): ONE MMA per (input row, k-step) carries all nine (row tap, step tap) weights along N,
    // each slab is read once; accumulators live in a ring of 4 per-slab TMEM buffers laid out [row][step tap][n] and
    // the epilogue adds the three step-tap partials of an output step (slabs s, s+1, s+2).
    long long *trace;                               // profiling hook (mvs_conv3d_c8_set_trace): per-CTA role timers, or null
    int trace_ctas;
    uint32_t ring_magic;                            // (1 << 18) / ring + 1: x / ring == (x * ring_magic) >> 18 for x < 32768
    int tmerged, buf_cols;
    int row_cols;                                   // T-merged: TMEM columns per output row of a slab buffer (3 n; 8 for n = 1)
    // flat 2D mode (MVS_FLAT2D; real D = 1, one image per batch element): a slab is a block of ht + 2 consecutive IMAGE rows
    // and the step axis walks the row blocks of the image.  The kh taps ride the row-tap groups, there are NO step taps: an
    // output step reads one buffer, one partial (n TMEM columns per row instead of 3 n -> up to 15 rows per step for n = 8).
    int flat2d;
    int pair_store;                                 // transposed stride-2 layers with Cout <= 8: even / odd output columns of a row
                                                    // leave as one 32-byte store (y 32-byte aligned, Wo even)
    int op_begin[UM_MAX_ISSUERS][4];                // issuer j, depth slab r: ops [op_begin[j][r], op_begin[j][r+1])
    AccOut acc[UM_MAX_ACC];
    // issue-ready op table (16 B per MMA, read with one uniform constant load):
    //   x = a_off | a_lbo << 16      (+ slab base at issue time)      y = b_off | N << 16   (+ weights base)
    //   z = accumulator column                                        w = accumulate (bit 0) | (MMA N >> 3) << 8
    uint4 ops[UM_MAX_OPS];
};

struct PackPlan {
    int cin, cout, n, cout_tiles, n_ksteps, nblk, transposed_weights, flip;
    int pad_rows;                 // all-zero rows appended to every chunk of a B block (T-merged, n = 8)
    int grp_rows;                 // T-merged: rows per row-tap group (3 n real rows, the rest of the group zero); 0: no groups
    int n_grp;                    // T-merged: groups per block (3; 4 with paired input rows)
    int n_t;                      // T-merged: step taps per group (3; 1 in flat 2D mode)
    KStepSrc ks[UM_MAX_KSTEPS];   // [n_ksteps * nblk]: weight source of block `blk` of k-step `k` at [k * nblk + blk]
};

// ---- device helpers (inline PTX; sm_100a) -----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start[0,14) | LBO[16,30) | SBO[32,46) | version=1 [46,48)
// | base_offset[49,52)=0 | lbo_mode[52]=0 | layout_type[61,64)=0 (SWIZZLE_NONE / interleave); all in 16 B
// units.  The issue loop assembles it as (kDescHi << 32) | (start + LBO << 16).

// fp16 operands: a_format = b_format = 0
__host__ __device__ constexpr uint32_t umma_idesc_f16(int m, int n)
{
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n)
{
    // InstrDescriptor: c_format[4,6)=1 (F32) | a_format[7,10)=1 (BF16) | b_format[10,13)=1 | a_major[15]=0 (K)
    // | b_major[16]=0 (K) | n_dim[17,23)=N>>3 | m_dim[24,29)=M>>4
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate));
}

__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
                 :: "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" :: "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}\n"
        :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// x / ring and x % ring without a hardware-less integer division (~25 dependent instructions each): exact for x < 32768
__device__ __forceinline__ int div_ring(int x, uint32_t magic) { return (int)(((uint32_t)x * magic) >> 18); }
__device__ __forceinline__ int mod_ring(int x, int ring, uint32_t magic) { return x - div_ring(x, magic) * ring; }

// position of column w in a W-de-interleaved row (we = ceil(W / 2))
__device__ __forceinline__ int dw_col(int w, int we) { return (w & 1) * we + (w >> 1); }

struct RoleTimer {          // accumulates in registers (a global read-modify-write per lap would cost ~700 clk each)
    long long *p; long long t, a0, a1, a2;
    int k0;
    __device__ __forceinline__ void start(bool on, long long *base, int first_slot) { p = on ? base : nullptr; k0 = first_slot; a0 = a1 = a2 = 0; if (p) t = clock64(); }
    __device__ __forceinline__ void lap(int k) {
        if (p) { const long long n = clock64(); const long long d = n - t; t = n; const int j = k - k0; if (j == 0) a0 += d; else if (j == 1) a1 += d; else a2 += d; }
    }
    __device__ __forceinline__ void flush() { if (p) { p[k0] = a0; p[k0 + 1] = a1; if (k0 == 3) p[k0 + 2] = a2; } }
};

// The table-driven kernels run at 56 / 64 registers: their role timers exist only in -DUM_TRACE_OLD=1 diagnostic builds
// (tools/prof_conv_trace.py with MVS_B200_LIB pointing at such a build); in the shipped build they compile to nothing.
#ifndef UM_TRACE_OLD
#define UM_TRACE_OLD 0
#endif
#if UM_TRACE_OLD
using RoleTimerOld = RoleTimer;
#else
struct RoleTimerOld {
    __device__ __forceinline__ void start(bool, long long *, int) {}
    __device__ __forceinline__ void lap(int) {}
    __device__ __forceinline__ void flush() {}
};
#endif

__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}

__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(taddr), "r"(cols) : "memory");
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

__device__ __forceinline__ void cp_async16(void *dst, const void *src, uint32_t src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
// 1-D bulk copy global -> shared (TMA engine, no tensor map): completion is reported to `bar` as `bytes` of tx-count
__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n"
                 :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N) : "memory"); }
// TMEM loads WITHOUT the wait: issue several, then tmem_wait_ld() once.
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld1_nowait(uint32_t taddr, uint32_t &r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n" : "=r"(r) : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b)
{
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&v);
}
__device__ __forceinline__ uint32_t pack_f16x2(float a, float b)
{
    __half2 v = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&v);
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t u) { return __half22float2(*reinterpret_cast<__half2 *>(&u)); }

// ---- the kernel -----------------------------------------------------------------------------------
// MC = CTAs per SM the register allocator must allow (old modes): 3 for layers that live on occupancy, 2 for the
// skip layers whose epilogue keeps four rows (TMEM + skip loads) in flight
template <bool TM, int MC>
__global__ void __launch_bounds__(TM ? UM_THREADS_TM : (MC == 2 ? UM_THREADS + UM_EPI_THREADS : UM_THREADS), TM ? 1 : MC)
conv3d_umma_kernel(const __grid_constant__ ConvPlan P, const uint4 *__restrict__ x, const uint4 *__restrict__ wpk,
                   const float *__restrict__ scale, const float *__restrict__ shift, const uint4 *__restrict__ skip,
                   void *__restrict__ y)
{
    constexpr bool kSkip = MC == 2;          // the launcher picks the <., 2> instantiations exactly when a skip tensor is given
    // <true, 4> = the flat 2D mode (ConvPlan::flat2d): its own instantiation, so that the 3D T-merged kernels keep the code
    // and the register allocation they were tuned with (as run-time branches the extra paths cost them 5-15 %)
    constexpr bool kFlat = TM && (MC == 4 || MC == 5);
    // <true, 5>: flat 2D + a "pixel-shuffled" skip operand (MVS_SKIP_PS): the value added to output pixel (h, w) of channel
    // block cb sits in block ((h & 1) * 2 + (w & 1)) * cout_chunks + cb of a HALF-resolution map [N][4 * cout_chunks][H/2][W/2][8]
    // -- a convolution of a nearest-up-sampled map, computed at the low resolution as four parity classes (featurenet.py)
    constexpr bool kFlatSkip = TM && MC == 5;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint4 *sw = reinterpret_cast<uint4 *>(smem_raw);                 // packed weights of this Cout tile
    uint4 *sa = sw + P.weight_units + P.zero_units;                  // ring of depth slabs (after weights + zero block)
    uint64_t *bars = reinterpret_cast<uint64_t *>(sa + (size_t)P.ring * P.slab_units);
    uint64_t *full = bars;                       // [ring]  slab landed            (128 producer arrivals)
    uint64_t *empty = bars + UM_MAX_RING;        // [ring]  slab no longer read     (1 tcgen05.commit)
    uint64_t *tfull = bars + 2 * UM_MAX_RING;    // [4]     accumulators complete   (1 tcgen05.commit)
    uint64_t *tempty = tfull + UM_TBUFS;         // [4]     accumulators drained    (128 epilogue arrivals)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + UM_TBUFS);
    float *s_scale = reinterpret_cast<float *>(tmem_slot + 4);       // [n] folded-BN scale of this Cout tile (0 for padding)
    float *s_shift = s_scale + 32;                                   // [n] shift
    // [lines] per staged line: vector offset of (line, step 0, w 0) in x, or ~0 for a line outside the volume
    unsigned long long *s_line = reinterpret_cast<unsigned long long *>(s_shift + 32);
    // [n_acc] per accumulator: x = voxel offset of (its row, step 0, w 0), y = dd | wadd << 8 | row-in-range << 16
    ulonglong2 *s_acc = reinterpret_cast<ulonglong2 *>(s_line + UM_MAX_LINES);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);          // warp-uniform for the compiler (role dispatch, issuer id)
    const int m0 = blockIdx.x * 128;
    const int h0 = (int)(blockIdx.y % (unsigned)P.row_blocks) * P.ht;
    const int step_begin = (int)(blockIdx.y / (unsigned)P.row_blocks) * P.steps_per_cta;      // this CTA's chunk of the step axis
    const int nsteps = min(P.steps - step_begin, P.steps_per_cta);
    const int b = blockIdx.z / P.cout_tiles, ct = blockIdx.z % P.cout_tiles;

    const int cta_lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    long long *trace = (P.trace && cta_lin < P.trace_ctas) ? P.trace + (size_t)cta_lin * 16 : nullptr;
    const long long t_cta0 = trace ? clock64() : 0;
    if (warp == 0) tmem_alloc(tmem_slot, (uint32_t)P.tmem_cols);
    if (tid == 32) {
        const uint32_t n_commit = (P.merged || TM) ? 1u : (uint32_t)P.n_issuers;
        // T-merged: slabs arrive by bulk copies (tx bytes) + one arrival per producer warp
        for (int i = 0; i < P.ring; ++i) { mbar_init(full + i, TM ? UM_PROD_WARPS_TM : UM_PROD_THREADS); mbar_init(empty + i, n_commit); }
        for (int i = 0; i < UM_TBUFS; ++i) { mbar_init(tfull + i, n_commit); mbar_init(tempty + i, (TM || MC == 2) ? 2 * UM_EPI_THREADS : UM_EPI_THREADS); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    {   // weights: staged once per CTA
        const uint4 *src = wpk + (size_t)ct * P.weight_units;
        for (int i = tid; i < P.weight_units; i += blockDim.x) sw[i] = __ldg(src + i);
        for (int i = tid; i < P.zero_units; i += blockDim.x) sw[P.weight_units + i] = make_uint4(0, 0, 0, 0);
    }
    if (tid < P.n) {   // epilogue affine of this Cout tile; padded channels get (0, 0) so they store 0
        const int c = ct * P.n + tid;
        s_scale[tid] = c < P.cout ? (scale ? __ldg(scale + c) : 1.f) : 0.f;
        s_shift[tid] = c < P.cout ? (shift ? __ldg(shift + c) : 0.f) : 0.f;
    }
    if (!TM && tid < P.n_acc) {
        // epilogue accumulator table: the step-invariant part of out_pos (the per-step 64-bit index arithmetic and the
        // dynamically indexed reads of P.acc were ~25 dependent instructions per accumulator and step)
        const AccOut ao = P.acc[tid];
        const int oh = P.oh_mul * (h0 + ao.th) + ao.dh, od0 = P.od_mul * step_begin + ao.dd;
        const size_t plane_o = (size_t)P.Hor * P.Wo;
        const size_t off = P.swap ? (size_t)oh * plane_o + (size_t)od0 * P.Wo : (size_t)od0 * plane_o + (size_t)oh * P.Wo;
        // high half of x: the same voxel's offset in a W-de-interleaved skip tensor, less the thread's column / 2
        // (transposed layers only: ow = 2 (m0 + m) + wadd sits at (wadd & 1) * ceil(Wo / 2) + m0 + m)
        const unsigned long long off_dw = (unsigned long long)(uint32_t)(off + (size_t)((ao.wadd & 1) * ((P.Wo + 1) >> 1)));
        s_acc[tid] = make_ulonglong2((unsigned long long)(uint32_t)(off + (size_t)ao.wadd) | (off_dw << 32),
                                     (unsigned long long)((uint32_t)ao.dd | ((uint32_t)ao.wadd << 8) | ((oh < P.Ho ? 1u : 0u) << 16)));
    }
    if (!TM) {
        // producer line table: everything about a staged line that does not depend on the slab
        const int lines = P.rh * P.cin_chunks * P.arr;
        const size_t plane_in = (size_t)P.Hr * P.W;
        for (int ln = tid; ln < lines; ln += blockDim.x) {
            const int chunk = (ln / P.arr) % P.cin_chunks, r = ln / (P.arr * P.cin_chunks);
            const int h_in = P.h_mul * h0 + P.h_base + r;
            unsigned long long v = ~0ull;
            if (h_in >= 0 && h_in < P.H)
                v = (unsigned long long)((((size_t)b * P.cin_chunks + chunk) * P.Dr + (P.swap ? h_in : 0)) * plane_in +
                                         (P.swap ? (size_t)0 : (size_t)h_in * P.W));
            s_line[ln] = v;
        }
    }
    fence_async_smem();            // generic-proxy smem writes -> visible to the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // Programmatic dependent launch (the launcher sets the attribute): everything above -- TMEM allocation, barrier
    // initialisation, weight staging, index tables -- reads nothing a preceding kernel writes, so it may run under the tail of
    // the previous layer; from here on the activations are touched, so wait for the preceding grid to complete and flush.
    // The next layer's CTAs may start their own prologue as soon as every CTA of this grid is past this point.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint32_t taddr = *tmem_slot;
    if (trace && tid == 0) { trace[9] = clock64() - t_cta0; trace[10] = nsteps; }
    const int n_slabs = TM ? (kFlat ? nsteps : nsteps + 2) : P.d_mul * (nsteps - 1) + P.rd;   // slabs this CTA stages in total

    if (TM && (warp < 4 || warp >= 12)) {
        // =========================== T-merged epilogue: TMEM -> registers -> global =====================
        // Two groups of four warps (a warp may only read the TMEM lanes 32 (warp % 4) ..): group g takes the rows
        // g, g+2, ...  An output step adds the step-tap partials of slabs step, step+1, step+2.  Work items
        // (row, 8-channel block) are processed two at a time so that six TMEM loads are in flight per wait.
        const int eg = warp >= 12 ? 1 : 0, ew = warp & 3;
        const int m = ew * 32 + lane;                                 // row of the M tile owned by this thread
        const uint32_t lane_base = taddr + ((uint32_t)(ew * 32) << 16);
        const size_t vol_o = (size_t)P.Dor * P.Hor * P.Wo;
        const int n3 = P.row_cols, nb = P.n >> 3, nb_shift = nb == 4 ? 2 : (nb == 2 ? 1 : 0);
        const int my_rows = (P.ht - eg + 1) >> 1, n_items = my_rows * nb;
        const int ow = m0 + m;
        // output index of (row a, step): pos0 + a * row_stride + step * step_stride  (hoisted 64-bit arithmetic)
        const size_t plane_o = (size_t)P.Hor * P.Wo;
        const size_t row_stride = (P.swap && !kFlat) ? plane_o : (size_t)P.Wo;
        const size_t step_stride = kFlat ? (size_t)P.ht * P.Wo : (P.swap ? (size_t)P.Wo : plane_o);
        const size_t pos00 = (size_t)h0 * row_stride + (size_t)step_begin * step_stride;
        const size_t pos0 = pos00 + (size_t)ow;
        const bool w_ok = ow < P.Wo;
        const size_t chunk_base = ((size_t)b * P.cout_chunks + (size_t)ct * nb) * vol_o;
        const int we_o = (P.Wo + 1) >> 1;
        // 64-bit bases once per thread, 32-bit offsets in the loops (ptxas re-materialised the 64-bit index arithmetic per
        // row otherwise: ~50 of the ~140 instructions a row cost).  The launcher checks that the offsets fit 32 bits.
        uint4 *const ybase = reinterpret_cast<uint4 *>(y) + chunk_base + pos00 + (size_t)(P.y_dw ? dw_col(ow, we_o) : ow);
        const uint4 *const sbase = skip + chunk_base + pos00 + (size_t)(P.skip_dw ? dw_col(ow, we_o) : ow);
        float *const fbase = reinterpret_cast<float *>(y) + (size_t)b * vol_o + pos0;
        const uint32_t rs32 = (uint32_t)row_stride, ss32 = (uint32_t)step_stride, vo32 = (uint32_t)vol_o;
        // folded-BN affine of the first 8-channel block in registers (the only block when Cout <= 8)
        float2 sc2[4], sh2[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            sc2[e] = make_float2(s_scale[2 * e], s_scale[2 * e + 1]);
            sh2[e] = make_float2(s_shift[2 * e], s_shift[2 * e + 1]);
        }
        RoleTimer rt; rt.start(trace && tid == 0, trace, 6);
        for (int step = 0; step < nsteps; ++step) {
            if (kFlat) {
                // ---- flat 2D: one slab, one partial per output row block ----
                const int row_lim = P.Hr - (step_begin + step) * P.ht;            // image rows left in this block
                // the skip vectors do not depend on the accumulators: all of this step's are requested BEFORE waiting for the
                // MMAs (at most eight work items per thread: ht * n <= 120 columns per buffer, two epilogue groups)
                uint4 sk[8];
                if (kFlatSkip) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        sk[j] = make_uint4(0, 0, 0, 0);
                        if (j < n_items) {
                            const int a = eg + 2 * (j >> nb_shift), n0 = (j & (nb - 1)) * 8;
                            const int h = (step_begin + step) * P.ht + a, cb = ct * nb + (n0 >> 3);
                            if (w_ok && a < row_lim && cb < P.cout_chunks) {
                                const int hh = P.Hr >> 1, wh = P.W >> 1;
                                const size_t blk = (size_t)b * 4 * P.cout_chunks + (size_t)(((h & 1) * 2 + (ow & 1)) * P.cout_chunks + cb);
                                sk[j] = __ldg(skip + (blk * hh + (h >> 1)) * wh + (ow >> 1));
                            }
                        }
                    }
                }
                mbar_wait(tfull + (step & (UM_TBUFS - 1)), (uint32_t)(step >> UM_TBUFS_LOG2) & 1u);
                rt.lap(6);
                tc_fence_after();
                const uint32_t tc = lane_base + (uint32_t)((step & (UM_TBUFS - 1)) * P.buf_cols);
                const uint32_t so = (uint32_t)step * ss32;
#pragma unroll
                for (int it = 0; it < 8; it += 4) {
                    if (it >= n_items) break;
                    uint32_t q[4][8];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (it + j < n_items)
                            tmem_ld8_nowait(tc + (uint32_t)((eg + 2 * ((it + j) >> nb_shift)) * n3 + ((it + j) & (nb - 1)) * 8), q[j]);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (it + j >= n_items) break;
                        const int a = eg + 2 * ((it + j) >> nb_shift), n0 = ((it + j) & (nb - 1)) * 8;
                        if (!w_ok || a >= row_lim || ct * nb + (n0 >> 3) >= P.cout_chunks) continue;
                        uint32_t pk[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float2 v = __ffma2_rn(make_float2(__uint_as_float(q[j][2 * e]), __uint_as_float(q[j][2 * e + 1])),
                                                  make_float2(s_scale[n0 + 2 * e], s_scale[n0 + 2 * e + 1]),
                                                  make_float2(s_shift[n0 + 2 * e], s_shift[n0 + 2 * e + 1]));
                            if (P.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); }
                            if (kFlatSkip) {
                                const uint4 &s4 = sk[it + j];
                                const uint32_t sv = e == 0 ? s4.x : (e == 1 ? s4.y : (e == 2 ? s4.z : s4.w));
                                if (P.f16) { const float2 s2 = unpack_f16x2(sv); v.x += s2.x; v.y += s2.y; }
                                else { v.x += __uint_as_float(sv << 16); v.y += __uint_as_float(sv & 0xffff0000u); }
                            }
                            pk[e] = P.f16 ? pack_f16x2(v.x, v.y) : pack_bf16x2(v.x, v.y);
                        }
                        ybase[so + (uint32_t)a * rs32 + (uint32_t)(n0 >> 3) * vo32] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                }
                tc_fence_before();
                mbar_arrive(tempty + (step & (UM_TBUFS - 1)));
                rt.lap(7);
                continue;
            }
            // earlier slabs were waited for in earlier steps
            for (int sl = step == 0 ? 0 : step + 2; sl <= step + 2; ++sl)
                mbar_wait(tfull + (sl & (UM_TBUFS - 1)), (uint32_t)(sl >> UM_TBUFS_LOG2) & 1u);
            rt.lap(6);
            tc_fence_after();
            uint32_t tcol[3];
#pragma unroll
            for (int t = 0; t < 3; ++t)
                tcol[t] = lane_base + (uint32_t)(((step + t) & (UM_TBUFS - 1)) * P.buf_cols + t * P.n);
            const uint32_t so = (uint32_t)step * ss32;
            if (P.out_f32) {
                // `prob`: one real channel -> fp32 logits; up to four rows (12 single-column loads) per wait
                float *yo = fbase + so;
                for (int j0 = 0; j0 < my_rows; j0 += 4) {
                    uint32_t r[4][3];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (j0 + j < my_rows) {
                            const uint32_t c0 = (uint32_t)((eg + 2 * (j0 + j)) * n3);
#pragma unroll
                            for (int t = 0; t < 3; ++t) tmem_ld1_nowait(tcol[t] + c0, r[j][t]);
                        }
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (j0 + j >= my_rows) break;
                        const int a = eg + 2 * (j0 + j);
                        if (!w_ok || h0 + a >= P.Ho) continue;
                        const float acc = (__uint_as_float(r[j][0]) + __uint_as_float(r[j][1])) + __uint_as_float(r[j][2]);
                        float o = fmaf(acc, sc2[0].x, sh2[0].x);
                        if (P.relu) o = fmaxf(o, 0.f);
                        yo[(uint32_t)a * rs32] = o;
                    }
                }
            } else {
                auto finish_with = [&](const uint32_t (&r0)[8], const uint32_t (&r1)[8], const uint32_t (&r2)[8], int a, int n0,
                                       const float2 (&scl)[4], const float2 (&shl)[4]) {
                    const int cb = ct * nb + (n0 >> 3);                       // output channel block
                    if (!w_ok || h0 + a >= P.Ho || cb >= P.cout_chunks) return;
                    const uint32_t oidx = so + (uint32_t)a * rs32 + (uint32_t)(n0 >> 3) * vo32;
                    uint4 sk = make_uint4(0, 0, 0, 0);
                    if (kSkip) sk = __ldg(sbase + oidx);
                    const uint32_t sv[4] = {sk.x, sk.y, sk.z, sk.w};
                    uint32_t pk[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float2 v = __fadd2_rn(make_float2(__uint_as_float(r0[2 * e]), __uint_as_float(r0[2 * e + 1])),
                                              make_float2(__uint_as_float(r1[2 * e]), __uint_as_float(r1[2 * e + 1])));
                        v = __fadd2_rn(v, make_float2(__uint_as_float(r2[2 * e]), __uint_as_float(r2[2 * e + 1])));
                        v = __ffma2_rn(v, scl[e], shl[e]);
                        if (P.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); }
                        if (kSkip) {
                            if (P.f16) { const float2 s2 = unpack_f16x2(sv[e]); v.x += s2.x; v.y += s2.y; }
                            else { v.x += __uint_as_float(sv[e] << 16); v.y += __uint_as_float(sv[e] & 0xffff0000u); }
                        }
                        pk[e] = P.f16 ? pack_f16x2(v.x, v.y) : pack_bf16x2(v.x, v.y);
                    }
                    ybase[oidx] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                };
                // one channel block (Cout <= 8): the affine lives in registers; otherwise it is read per block
                auto finish = [&](const uint32_t (&r0)[8], const uint32_t (&r1)[8], const uint32_t (&r2)[8], int a, int n0) {
                    if (nb == 1) { finish_with(r0, r1, r2, a, n0, sc2, sh2); return; }
                    float2 scl[4], shl[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        scl[e] = make_float2(s_scale[n0 + 2 * e], s_scale[n0 + 2 * e + 1]);
                        shl[e] = make_float2(s_shift[n0 + 2 * e], s_shift[n0 + 2 * e + 1]);
                    }
                    finish_with(r0, r1, r2, a, n0, scl, shl);
                };
                for (int it = 0; it < n_items; it += 2) {
                    // nb = n / 8 is 1, 2 or 4: shifts instead of runtime divisions (a division is ~25 dependent
                    // instructions = ~125 clk of this warp)
                    const int a0 = eg + 2 * (it >> nb_shift), n00 = (it & (nb - 1)) * 8;
                    const bool two = it + 1 < n_items;
                    const int a1 = eg + 2 * ((it + 1) >> nb_shift), n01 = ((it + 1) & (nb - 1)) * 8;
                    uint32_t p0[8], p1[8], p2[8], q0[8], q1[8], q2[8];
                    const uint32_t c0 = (uint32_t)(a0 * n3 + n00), c1 = (uint32_t)(a1 * n3 + n01);
                    tmem_ld8_nowait(tcol[0] + c0, p0);
                    tmem_ld8_nowait(tcol[1] + c0, p1);
                    tmem_ld8_nowait(tcol[2] + c0, p2);
                    if (two) {
                        tmem_ld8_nowait(tcol[0] + c1, q0);
                        tmem_ld8_nowait(tcol[1] + c1, q1);
                        tmem_ld8_nowait(tcol[2] + c1, q2);
                    }
                    tmem_wait_ld();
                    finish(p0, p1, p2, a0, n00);
                    if (two) finish(q0, q1, q2, a1, n01);
                }
            }
            // (Round 2 experiment, dropped: reading the tap-0 partials first and handing the buffer back BEFORE the math and
            // the stores -- the issuer of slab step + 4 waits for exactly this arrival -- made every T-merged layer 10-25 %
            // SLOWER: epilogue work 569 -> 892 clk per step on `prob`, 897 -> 1303 on conv0 stage 3.  The MMAs then run
            // concurrently with this role's TMEM reads and stores; serialised, as here, both are faster.)
            tc_fence_before();                                     // this thread's TMEM reads are complete ...
            mbar_arrive(tempty + (step & (UM_TBUFS - 1)));         // ... slab `step`'s buffer may be overwritten
            rt.lap(7);
        }
        rt.flush();
    } else
    if ((warp >= 4 && warp < 8) || (TM && (warp == 10 || warp == 11))) {
        // =========================== producers: global -> shared (cp.async / bulk copies) =============
        // (T-merged: a UBLKCP costs its issuing WARP ~300-450 clk whichever lane issues it, so the lines of a slab are
        // spread over six warps: the four producer warps and the two issuer warps this mode leaves idle)
        const int pwarp = warp < 8 ? warp - 4 : warp - 6;
        constexpr int NPW = TM ? UM_PROD_WARPS_TM : UM_PROD_THREADS / 32;
        const int lines = P.rh * P.cin_chunks * P.arr;                // lines of UM_COLS 16-byte vectors per slab
        const size_t plane_in = (size_t)P.Hr * P.W;
        int pending = -1;              // slab staged (cp.async committed) but not yet published
        // per-thread column offsets of the (up to two) staged arrays and the per-slab stride, hoisted out of the loops
        int wofs[2][(UM_COLS + 31) / 32];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int cc = 0; cc < (UM_COLS + 31) / 32; ++cc) {
                const int w_in = P.w_step * (cc * 32 + lane) + P.w_step * m0 + P.w_base[a];
                wofs[a][cc] = (w_in >= 0 && w_in < P.W) ? (P.x_dw ? dw_col(w_in, (P.W + 1) >> 1) : w_in) : -1;
            }
        const size_t slab_stride = P.swap ? (size_t)P.W : plane_in;
        // (Stride-2 layers over a W-de-interleaved input, x_dw: their cp.async lanes read contiguous 16 B vectors, 4 L1
        // wavefronts per instruction instead of 8.  One bulk copy per staged line was tried on top of that and is not
        // faster -- conv1 unchanged, the multi-chunk conv3 / conv5 10-25 % slower: 2-3 UBLKCP per lane and slab.)
        if (kFlat) {
            // flat 2D: slab i = image rows (step_begin + i) * ht - 1 .. + ht of every channel block; one bulk copy per line,
            // rows above / below the image (first / last block) and the column halo are zero-filled
            RoleTimer rt; rt.start(trace && tid == 128, trace, 1);
            const int c_lo = m0 == 0 ? 1 : 0;                                  // column c holds w = m0 - 1 + c
            const int c_hi = min(UM_COLS, P.W - m0 + 1);
            const int ln = pwarp + NPW * lane;
            const int my_row = ln / P.cin_chunks, my_chunk = ln % P.cin_chunks;
            const bool has_line = ln < lines && c_hi > c_lo;
            const uint32_t my_bytes = has_line ? (uint32_t)(c_hi - c_lo) * 16u : 0u;
            uint32_t full_bytes = my_bytes;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) full_bytes += __shfl_xor_sync(0xffffffffu, full_bytes, o);
            const long long src0 = (((long long)b * P.cin_chunks + my_chunk) * P.Hr + (my_row - 1)) * (long long)P.W + (m0 - 1 + c_lo);
            const uint32_t my_dst_off = (uint32_t)ln * UM_COLS + (uint32_t)c_lo;
            const bool edge_cols = c_lo > 0 || c_hi < UM_COLS;
            int slot = 0, q = 0;
            for (int i = 0; i < n_slabs; ++i, slot = slot + 1 == P.ring ? 0 : slot + 1, q += slot == 0 ? 1 : 0) {
                if (q >= 1) mbar_wait(empty + slot, (uint32_t)(q - 1) & 1u);
                rt.lap(1);
                const int blk = step_begin + i;
                const int row0 = blk * P.ht - 1;                                // image row of staged row 0
                const bool clipped = row0 < 0 || row0 + P.rh > P.Hr;
                const bool my_ok = my_bytes != 0 && (unsigned)(row0 + my_row) < (unsigned)P.Hr;
                uint4 *slab = sa + (size_t)slot * P.slab_units;
                uint32_t warp_bytes = full_bytes;
                if (edge_cols || clipped) {
                    for (int l2 = pwarp; l2 < lines; l2 += NPW) {
                        const bool row_ok = (unsigned)(row0 + l2 / P.cin_chunks) < (unsigned)P.Hr;
                        uint4 *dst = slab + (size_t)l2 * UM_COLS;
                        if (!row_ok) {
                            for (int c = lane; c < UM_COLS; c += 32) dst[c] = make_uint4(0, 0, 0, 0);
                        } else {
                            if (lane < c_lo) dst[lane] = make_uint4(0, 0, 0, 0);
                            for (int c = c_hi + lane; c < UM_COLS; c += 32) dst[c] = make_uint4(0, 0, 0, 0);
                        }
                    }
                    fence_async_smem();
                    if (clipped) {
                        warp_bytes = my_ok ? my_bytes : 0u;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) warp_bytes += __shfl_xor_sync(0xffffffffu, warp_bytes, o);
                    }
                    __syncwarp();
                }
                if (lane == 0) mbar_arrive_expect_tx(full + slot, warp_bytes);
                __syncwarp();
                if (my_ok) bulk_copy_g2s(slab + my_dst_off, x + (src0 + (long long)blk * P.ht * P.W), my_bytes, full + slot);
                rt.lap(2);
            }
            rt.flush();
        } else
        if (TM) {
            // A staged line (one row of one channel block, 132 consecutive voxels) is contiguous in global memory, so it
            // moves as ONE bulk copy issued by one lane (lane l of producer warp p owns line p + 4 l); only the out-of-
            // volume parts (halo rows / columns, slabs beyond the step axis) are zero-filled with ordinary stores.
            // Everything but the slab index is hoisted, so an interior slab costs a barrier wait + one copy per lane.
            // (Measured: a UBLKCP costs ~450 clk of its warp whoever issues it; spreading the lines over lanes of four
            // warps beats one elected lane walking them.  One tensor-map TMA per slab is the next step.)
            RoleTimer rt; rt.start(trace && tid == 128, trace, 1);
            const int c_lo = m0 == 0 ? 1 : 0;                                  // column c holds w = m0 - 1 + c
            const int c_hi = min(UM_COLS, P.W - m0 + 1);
            const int ln = pwarp + NPW * lane;
            const int my_row = ln / P.cin_chunks, my_chunk = ln % P.cin_chunks;
            const int my_h = h0 - 1 + my_row;
            const bool my_ok = ln < lines && my_h >= 0 && my_h < P.H && c_hi > c_lo;
            const uint32_t my_bytes = my_ok ? (uint32_t)(c_hi - c_lo) * 16u : 0u;
            uint32_t warp_bytes = my_bytes;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) warp_bytes += __shfl_xor_sync(0xffffffffu, warp_bytes, o);
            // source of this lane's line in slab 0 (d_in = d_base + step_begin) and its stride per slab
            const size_t step_stride = P.swap ? (size_t)P.W : plane_in;
            const uint4 *my_src = x + (((size_t)b * P.cin_chunks + my_chunk) * P.Dr + (P.swap ? (my_ok ? my_h : 0) : 0)) * plane_in
                                  + (P.swap ? (size_t)0 : (size_t)(my_ok ? my_h : 0) * P.W) + (m0 - 1 + c_lo);
            const uint32_t my_dst_off = (uint32_t)ln * UM_COLS + (uint32_t)c_lo;
            // does any line of this warp need zero columns / zero rows in an in-range slab?
            const bool edge_cols = c_lo > 0 || c_hi < UM_COLS;
            const bool edge_rows = __any_sync(0xffffffffu, ln < lines && !(my_h >= 0 && my_h < P.H));
            int slot = 0, q = 0;                       // i = q * ring + slot, kept incrementally (no divisions per slab)
            for (int i = 0; i < n_slabs; ++i, slot = slot + 1 == P.ring ? 0 : slot + 1, q += slot == 0 ? 1 : 0) {
                if (q >= 1) mbar_wait(empty + slot, (uint32_t)(q - 1) & 1u);
                rt.lap(1);
                const int d_in = P.d_base + step_begin + i;
                const bool d_ok = d_in >= 0 && d_in < P.D;
                uint4 *slab = sa + (size_t)slot * P.slab_units;
                if (!d_ok || edge_cols || edge_rows) {
                    // zero the parts no copy will write (smem slots are recycled, so this is redone per slab)
                    for (int l2 = pwarp; l2 < lines; l2 += NPW) {
                        const int h_in = h0 - 1 + l2 / P.cin_chunks;
                        const bool row_ok = d_ok && h_in >= 0 && h_in < P.H;
                        uint4 *dst = slab + (size_t)l2 * UM_COLS;
                        if (!row_ok) {
                            for (int c = lane; c < UM_COLS; c += 32) dst[c] = make_uint4(0, 0, 0, 0);
                        } else {
                            if (lane < c_lo) dst[lane] = make_uint4(0, 0, 0, 0);
                            for (int c = c_hi + lane; c < UM_COLS; c += 32) dst[c] = make_uint4(0, 0, 0, 0);
                        }
                    }
                    fence_async_smem();
                    __syncwarp();
                }
                if (lane == 0) mbar_arrive_expect_tx(full + slot, d_ok ? warp_bytes : 0u);
                __syncwarp();
                if (d_ok && my_bytes) bulk_copy_g2s(slab + my_dst_off, my_src + (size_t)d_in * step_stride, my_bytes, full + slot);
                rt.lap(2);
            }
            rt.flush();
        } else {
        RoleTimerOld rto; rto.start(trace && tid == 128, trace, 1);       // old-mode producer: 1 wait empty | 2 stage + publish
        for (int i = 0; i < n_slabs; ++i) {
            const int q = div_ring(i, P.ring_magic), slot = i - q * P.ring;
            if (q >= 1) {
                // never block on `empty` while holding an unpublished slab: the issuer may need it to
                // retire the very slab we are waiting for (ring == rd leaves no slack)
                if (pending >= 0) {
                    cp_async_wait<0>();
                    fence_async_smem();
                    mbar_arrive(full + mod_ring(pending, P.ring, P.ring_magic));
                    pending = -1;
                }
                rto.lap(2);
                mbar_wait(empty + slot, (uint32_t)(q - 1) & 1u);
                rto.lap(1);
            }
            const int d_in = P.d_base + P.d_mul * step_begin + i;
            const bool d_ok = d_in >= 0 && d_in < P.D;
            uint4 *slab = sa + (size_t)slot * P.slab_units;
            const uint4 *xs = x + (size_t)(d_ok ? d_in : 0) * slab_stride;
            for (int ln = pwarp; ln < lines; ln += UM_PROD_THREADS / 32) {
                const unsigned long long lb = s_line[ln];
                const bool row_ok = d_ok && lb != ~0ull;
                const uint4 *src = xs + (row_ok ? (size_t)lb : 0);
                uint4 *dst = slab + (size_t)ln * UM_COLS;
                const int a = P.arr == 2 ? (ln & 1) : 0;
#pragma unroll
                for (int cc = 0; cc < (UM_COLS + 31) / 32; ++cc) {
                    const int c = cc * 32 + lane;
                    if (c < UM_COLS) {
                        const int w_in = a ? wofs[1][cc] : wofs[0][cc];     // -1: outside the row (static indices: registers)
                        const bool ok = row_ok && w_in >= 0;
                        cp_async16(dst + c, src + (ok ? w_in : 0), ok ? 16u : 0u);
                    }
                }
            }
            cp_async_commit();
            if (pending >= 0) {        // the previous slab has landed for this thread: publish it
                cp_async_wait<1>();
                fence_async_smem();
                mbar_arrive(full + mod_ring(pending, P.ring, P.ring_magic));
            }
            pending = i;
            rto.lap(2);
        }
        rto.flush();
        if (pending >= 0) {
            cp_async_wait<0>();
            fence_async_smem();
            mbar_arrive(full + mod_ring(pending, P.ring, P.ring_magic));
        }
        }
    } else if (warp >= 8 && warp < 12) {       // (T-merged: warps 10, 11 were taken by the producer branch above)
        // =========================== MMA issuers ======================================================
        // Up to UM_MAX_ISSUERS warps, each owning a disjoint set of accumulators (independent chains).
        // A whole warp walks its slice of the op table on warp-uniform values (kernel-parameter table,
        // loop counters), so descriptors are built in the uniform datapath; only the tcgen05 instructions
        // are predicated on one elected lane.  (Issuing from inside `if (lane == 0)` makes the compiler
        // wrap every UTCHMMA in an R2UR waterfall loop.)
        const int iss = warp - 8;
        if (TM) {
            // ONE issuer walks the slabs in order: every slab is read once, by the MMAs of its own op table, into the
            // slab's own TMEM buffer; the step-axis reduction happens in the epilogue.
            // Two issuers alternate slabs (one warp's descriptor arithmetic costs ~85 clk per MMA, the tensor pipe needs
            // ~45-70).  ring and UM_TBUFS are even, so a slab's smem slot and TMEM buffer always belong to the same
            // issuer: every barrier is waited on by exactly one warp, phase after phase.
            if (iss < 2) {
                uint32_t leader;
                asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(leader));
                const uint32_t sa_units = smem_u32(sa) >> 4, sw_units = smem_u32(sw) >> 4;
                const uint32_t idesc0 = P.f16 ? umma_idesc_f16(128, 0) : umma_idesc_bf16(128, 0);
                constexpr uint32_t kDescHi = 8u | (1u << 14);
                RoleTimer rt; rt.start(trace && lane == 0 && iss == 0, trace, 3);
                int slot = iss, q = 0;                     // sl = q * ring + slot (ring is even and >= 2, so slot = iss < ring)
                for (int sl = iss; sl < n_slabs; sl += 2) {
                    const int buf = sl & (UM_TBUFS - 1), use = sl >> UM_TBUFS_LOG2;
                    mbar_wait(full + slot, (uint32_t)q & 1u);
                    rt.lap(3);
                    if (use >= 1) mbar_wait(tempty + buf, (uint32_t)(use - 1) & 1u);
                    rt.lap(4);
                    tc_fence_after();
                    const uint32_t tbase = taddr + (uint32_t)(buf * P.buf_cols);
                    const uint32_t base = sa_units + (uint32_t)(slot * P.slab_units);
                    int i = 0;
                    for (; i + 4 <= P.n_ops; i += 4) {
                        uint64_t ad[4], bd[4];
                        uint32_t dc[4], ac[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint4 e = P.ops[i + j];
                            ad[j] = ((uint64_t)kDescHi << 32) | (uint64_t)(e.x + base);
                            bd[j] = ((uint64_t)kDescHi << 32) | (uint64_t)(e.y + sw_units);
                            dc[j] = tbase + e.z;
                            ac[j] = e.w;
                        }
                        if (leader) {
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                umma_bf16_ss(dc[j], ad[j], bd[j], idesc0 | ((ac[j] >> 8) << 17), ac[j] & 1u);
                        }
                    }
                    for (; i < P.n_ops; ++i) {
                        const uint4 e = P.ops[i];
                        const uint64_t ad = ((uint64_t)kDescHi << 32) | (uint64_t)(e.x + base);
                        const uint64_t bd = ((uint64_t)kDescHi << 32) | (uint64_t)(e.y + sw_units);
                        if (leader) umma_bf16_ss(tbase + e.z, ad, bd, idesc0 | ((e.w >> 8) << 17), e.w & 1u);
                    }
                    if (leader) {
                        umma_commit(empty + slot);     // the slab goes back to the producers once these MMAs retire
                        umma_commit(tfull + buf);
                    }
                    __syncwarp();
                    rt.lap(5);
                    slot += 2;
                    if (slot >= P.ring) { slot -= P.ring; ++q; }
                }
                rt.flush();
            }
        } else if (iss < P.n_issuers) {
            uint32_t leader;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(leader));
            const uint32_t sa_units = smem_u32(sa) >> 4, sw_units = smem_u32(sw) >> 4;
            const uint32_t idesc0 = P.f16 ? umma_idesc_f16(128, 0) : umma_idesc_bf16(128, 0);        // N comes from the op entry
            constexpr uint32_t kDescHi = 8u | (1u << 14);           // SBO = 8 units (128 B) | version = 1 (bit 46)
            // merged mode: the issuers alternate depth steps (issuer j owns TMEM buffer j and the whole op
            // table); otherwise every issuer works on every step with its own slice of accumulators
            const int step0 = P.merged ? iss : 0, dstep = P.merged ? P.n_issuers : 1;
            const int tbl = P.merged ? 0 : iss;
            int waited = 0;
            RoleTimerOld rti; rti.start(trace && lane == 0 && iss == 0, trace, 3);
            for (int step = step0; step < nsteps; step += dstep) {
                const int first = P.d_mul * step;
                // only slabs this step reads: an issuer that skips steps must not test the parity of a
                // barrier whose slot may already have been released and refilled (the phase would alias)
                if (waited < first) waited = first;
                while (waited < first + P.rd) {
                    const int wq = div_ring(waited, P.ring_magic);
                    mbar_wait(full + (waited - wq * P.ring), (uint32_t)wq & 1u);
                    ++waited;
                }
                rti.lap(3);
                const int buf = step & 1, use = step >> 1;
                if (use >= 1) mbar_wait(tempty + buf, (uint32_t)(use - 1) & 1u);
                rti.lap(4);
                tc_fence_after();
                const uint32_t tbase = taddr + (uint32_t)(buf * P.acc_cols);
                // ops are grouped by the depth slab they read, so the slab base is loop-invariant and an op
                // costs one 16 B constant load + three adds + the predicate in the uniform datapath
                for (int r = 0; r < P.rd; ++r) {
                    const uint32_t base = sa_units + (uint32_t)(mod_ring(first + r, P.ring, P.ring_magic) * P.slab_units);
                    const int op0 = P.op_begin[tbl][r], op1 = P.op_begin[tbl][r + 1];
                    int i = op0;
                    for (; i + 4 <= op1; i += 4) {
                        uint64_t ad[4], bd[4];
                        uint32_t dc[4], ac[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint4 e = P.ops[i + j];
                            ad[j] = ((uint64_t)kDescHi << 32) | (uint64_t)(e.x + base);
                            bd[j] = ((uint64_t)kDescHi << 32) | (uint64_t)(e.y + sw_units);
                            dc[j] = tbase + e.z;
                            ac[j] = e.w;
                        }
                        if (leader) {
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                umma_bf16_ss(dc[j], ad[j], bd[j], idesc0 | ((ac[j] >> 8) << 17), ac[j] & 1u);
                        }
                    }
                    for (; i < op1; ++i) {
                        const uint4 e = P.ops[i];
                        const uint64_t ad = ((uint64_t)kDescHi << 32) | (uint64_t)(e.x + base);
                        const uint64_t bd = ((uint64_t)kDescHi << 32) | (uint64_t)(e.y + sw_units);
                        if (leader) umma_bf16_ss(tbase + e.z, ad, bd, idesc0 | ((e.w >> 8) << 17), e.w & 1u);
                    }
                }
                if (leader) {
                    // slabs the next step no longer reads go back to the producers once these MMAs retire
                    // (merged mode: another issuer's step may still read them -> the epilogue frees them)
                    if (!P.merged)
                        for (int k = 0; k < P.d_mul; ++k) umma_commit(empty + mod_ring(first + k, P.ring, P.ring_magic));
                    umma_commit(tfull + buf);
                }
                __syncwarp();
                rti.lap(5);
            }
            rti.flush();
        }
    } else if (warp < 4 || (MC == 2 && warp >= 12)) {
        // =========================== epilogue: TMEM -> registers -> global ============================
        // (skip variant <false, 2>: a second group of four warps, 12-15, takes every other batch of accumulators -- the
        // transposed layers are epilogue-bound: 16 accumulators x (TMEM load, skip load, affine, store) per step)
        constexpr int NEG = MC == 2 ? 2 : 1;
        const int eg = warp >= 12 ? 1 : 0, ew = warp & 3;
        const int m = ew * 32 + lane;                                 // row of the M tile owned by this thread
        const uint32_t lane_base = taddr + ((uint32_t)(ew * 32) << 16);
        const int ow_thread = P.w_mul * (m0 + m);                     // + wadd = output w
        const uint32_t epi_step_stride = (uint32_t)(P.swap ? (size_t)P.Wo : (size_t)P.Hor * P.Wo) * (uint32_t)P.od_mul;
        // 64-bit bases once per thread, 32-bit voxel offsets in the loops (the launcher checks the volume fits 31 bits)
        const size_t vol_o = (size_t)P.Dor * P.Hor * P.Wo;
        const uint32_t vo32 = (uint32_t)vol_o;
        const size_t tile_base = ((size_t)b * P.cout_chunks + (size_t)((ct * P.n) >> 3)) * vol_o;
        uint4 *const ybase = reinterpret_cast<uint4 *>(y) + tile_base;
        const uint4 *const sbase = skip + tile_base;
        float *const fbase = reinterpret_cast<float *>(y) + (size_t)b * vol_o;
        // folded-BN affine of the first 8-channel block in registers (the only block when Cout <= 8)
        float2 sc2[4], sh2[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            sc2[e] = make_float2(s_scale[2 * e], s_scale[2 * e + 1]);
            sh2[e] = make_float2(s_shift[2 * e], s_shift[2 * e + 1]);
        }
        RoleTimerOld rte; rte.start(trace && tid == 0, trace, 6);      // old-mode epilogue: 6 wait tfull | 7 work
        for (int step = 0; step < nsteps; ++step) {
            const int buf = step & 1, use = step >> 1;
            mbar_wait(tfull + buf, (uint32_t)use & 1u);
            rte.lap(6);
            tc_fence_after();
            if (P.merged && tid == 0) {      // (merged mode never has MC == 2 groups competing: it is T-merged or skip-free)
                // steps complete in order as seen from here (we waited on every earlier tfull), so no MMA of
                // any step <= `step` still reads the slabs this step retires
                for (int k = 0; k < P.d_mul; ++k) mbar_arrive(empty + mod_ring(P.d_mul * step + k, P.ring, P.ring_magic));
            }
            const uint32_t tcol0 = lane_base + (uint32_t)(buf * P.acc_cols);
            // affine + ReLU + skip + store of one 8-channel block held in v[0..7]
            // (the skip operand is fetched by the caller BEFORE the TMEM wait: issued one at a time next to its use,
            // every skip load costs a full DRAM round trip of this warp.  Tried and dropped in round 2: prefetch.global.L1
            // of the step's skip vectors before the tfull wait -- epilogue work 3.6 k -> 2.9 k clk per step but the CTA
            // slower, 5.1 k -> 6.5 k: the prefetch pass itself sits on this role's critical path; holding the first batch in
            // registers across the wait spills at 64 registers.)
            auto pack_with = [&](const uint32_t (&v)[8], const uint4 &sk, const float2 (&scl)[4], const float2 (&shl)[4]) -> uint4 {
                const uint32_t sv[4] = {sk.x, sk.y, sk.z, sk.w};
                uint32_t pk[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float2 t = __ffma2_rn(make_float2(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1])), scl[e], shl[e]);
                    if (P.relu) { t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); }
                    if (kSkip) {
                        if (P.f16) { const float2 s2 = unpack_f16x2(sv[e]); t.x += s2.x; t.y += s2.y; }
                        else { t.x += __uint_as_float(sv[e] << 16); t.y += __uint_as_float(sv[e] & 0xffff0000u); }
                    }
                    pk[e] = P.f16 ? pack_f16x2(t.x, t.y) : pack_bf16x2(t.x, t.y);
                }
                return make_uint4(pk[0], pk[1], pk[2], pk[3]);
            };
            auto store_with = [&](const uint32_t (&v)[8], uint32_t oidx, const uint4 &sk, const float2 (&scl)[4],
                                  const float2 (&shl)[4]) { ybase[oidx] = pack_with(v, sk, scl, shl); };
            auto store_chunk = [&](const uint32_t (&v)[8], int nloc, uint32_t oidx, const uint4 &sk) {
                float2 scl[4], shl[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    scl[e] = make_float2(s_scale[nloc + 2 * e], s_scale[nloc + 2 * e + 1]);
                    shl[e] = make_float2(s_shift[nloc + 2 * e], s_shift[nloc + 2 * e + 1]);
                }
                store_with(v, oidx, sk, scl, shl);
            };
            const int od_step = P.od_mul * (step_begin + step);                   // + dd = output position on the step axis
            const uint32_t step_off = (uint32_t)step * epi_step_stride + (uint32_t)ow_thread;
            // voxel offset of accumulator a's row for this thread in y (return value) and in the skip tensor (spos)
            const uint32_t step_off_dw = (uint32_t)step * epi_step_stride + (uint32_t)(ow_thread >> 1);
            auto out_pos = [&](int a, bool &ok, uint32_t &spos) -> uint32_t {
                const ulonglong2 e = s_acc[a];
                const uint32_t f = (uint32_t)e.y;
                ok = (f >> 16) != 0 && od_step + (int)(f & 0xffu) < P.Do && ow_thread + (int)((f >> 8) & 0xffu) < P.Wo;
                const uint32_t pos = (uint32_t)e.x + step_off;
                spos = P.skip_dw ? (uint32_t)(e.x >> 32) + step_off_dw : pos;
                return pos;
            };
            if (P.out_f32) {
                // `prob` layer: one real channel -> fp32 logits; four rows' single-column loads per wait
                for (int a0 = 4 * eg; a0 < P.n_acc; a0 += 4 * NEG) {
                    uint32_t r[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (a0 + j < P.n_acc) tmem_ld1_nowait(tcol0 + (uint32_t)((a0 + j) * P.n), r[j]);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (a0 + j >= P.n_acc) break;
                        bool ok;
                        uint32_t spos;
                        const uint32_t pos = out_pos(a0 + j, ok, spos);
                        if (!ok) continue;
                        float o = fmaf(__uint_as_float(r[j]), sc2[0].x, sh2[0].x);
                        if (P.relu) o = fmaxf(o, 0.f);
                        fbase[pos] = o;
                    }
                }
            } else if (P.cout <= 8) {
                // one 8-channel block per row: x8 loads, two rows in flight per wait
                // four rows per batch: positions, skip loads and TMEM loads are issued together, one wait
                constexpr int EB = MC >= 3 ? 2 : 4;
                for (int a0 = EB * eg; a0 < P.n_acc; a0 += EB * NEG) {
                    uint32_t r[EB][8];
                    uint32_t pos[EB], spos[EB];
                    bool ok[EB];
                    uint4 sk[EB];
#pragma unroll
                    for (int j = 0; j < EB; ++j) {
                        ok[j] = false;
                        sk[j] = make_uint4(0, 0, 0, 0);
                        if (a0 + j < P.n_acc) {
                            pos[j] = out_pos(a0 + j, ok[j], spos[j]);
                            if (kSkip && ok[j]) sk[j] = __ldg(sbase + spos[j]);
                            tmem_ld8_nowait(tcol0 + (uint32_t)((a0 + j) * P.n), r[j]);
                        }
                    }
                    tmem_wait_ld();
                    if (P.pair_store) {
                        // transposed layers: accumulators 2k, 2k + 1 are the even / odd output columns of the same row --
                        // one 32-byte store of the two neighbours instead of two 16-byte stores at a 32-byte stride
#pragma unroll
                        for (int j = 0; j < EB; j += 2)
                            if (ok[j]) {
                                const uint4 v0 = pack_with(r[j], sk[j], sc2, sh2), v1 = pack_with(r[j + 1], sk[j + 1], sc2, sh2);
                                asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                                             :: "l"(ybase + pos[j]), "r"(v0.x), "r"(v0.y), "r"(v0.z), "r"(v0.w), "r"(v1.x), "r"(v1.y),
                                                "r"(v1.z), "r"(v1.w) : "memory");
                            }
                    } else {
#pragma unroll
                        for (int j = 0; j < EB; ++j)
                            if (ok[j]) store_with(r[j], pos[j], sk[j], sc2, sh2);
                    }
                }
            } else {
                for (int a = eg; a < P.n_acc; a += NEG) {
                    bool ok;
                    uint32_t spos;
                    const uint32_t pos = out_pos(a, ok, spos);
                    for (int n0 = 0; n0 < P.n; n0 += 16) {
                        uint32_t r[16];
                        const int c0 = ct * P.n + n0;
                        const bool live = ok && c0 < P.cout, has_hi = c0 + 8 < P.cout_chunks * 8;
                        const uint32_t base = (uint32_t)(n0 >> 3) * vo32 + pos;
                        uint4 sk_lo = make_uint4(0, 0, 0, 0), sk_hi = make_uint4(0, 0, 0, 0);
                        if (kSkip && live) {
                            const uint32_t sb = (uint32_t)(n0 >> 3) * vo32 + spos;
                            sk_lo = __ldg(sbase + sb);
                            if (has_hi) sk_hi = __ldg(sbase + sb + vo32);
                        }
                        tmem_ld16_nowait(tcol0 + (uint32_t)(a * P.n + n0), r);
                        tmem_wait_ld();
                        if (!live) continue;
                        uint32_t lo[8], hi[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) { lo[e] = r[e]; hi[e] = r[8 + e]; }
                        store_chunk(lo, n0, base, sk_lo);
                        if (has_hi) store_chunk(hi, n0 + 8, base + vo32, sk_hi);
                    }
                }
            }
            tc_fence_before();         // this thread's TMEM reads are complete (tcgen05.wait::ld) ...
            mbar_arrive(tempty + buf); // ... so the issuer may overwrite the buffer
            rte.lap(7);
        }
        rte.flush();
    }
    tc_fence_before();
    __syncthreads();
    if (trace && tid == 0) trace[0] = clock64() - t_cta0;
    if (warp == 0) tmem_dealloc(taddr, (uint32_t)P.tmem_cols);
}

// Packs fp32 weights into the per-k-step B blocks: [cout_tile][kstep][2 chunks][nblk][N rows][8] bf16
// (nblk = 3 for kh-merged stride-1 layers: blocks ordered kh = 2, 1, 0; else 1).
__global__ void __launch_bounds__(256)
pack_weights_kernel(const __grid_constant__ PackPlan P, const float *__restrict__ w, uint16_t *__restrict__ out, int f16)
{
    const int grp = P.grp_rows > 0 ? P.grp_rows : P.nblk * P.n;  // T-merged: three row-tap groups of grp rows, 3 n of them real
    const int n_grp = P.grp_rows > 0 ? P.n_grp : 1, real_pg = P.grp_rows > 0 ? P.n_t * P.n : P.nblk * P.n;
    const int rows_pc = n_grp * grp + P.pad_rows;                // rows per K chunk of a B block
    const long long total = (long long)P.cout_tiles * P.n_ksteps * 2 * rows_pc * 8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long t = i;
        const int e = (int)(t % 8); t /= 8;
        const int rr = (int)(t % rows_pc); t /= rows_pc;
        const int gi = rr / grp, within = rr - gi * grp;
        if (gi >= n_grp || within >= real_pg) { out[i] = 0; continue; }
        const int row = within % P.n, blk = gi * (real_pg / P.n) + within / P.n;
        const int j = (int)(t % 2); t /= 2;
        const int ks = (int)(t % P.n_ksteps);
        const int ct = (int)(t / P.n_ksteps);
        const KStepSrc src = P.ks[ks * P.nblk + blk];
        const int tap = src.tap[j];
        const int ci = src.chunk[j] * 8 + e, co = ct * P.n + row;
        float v = 0.f;
        if (tap >= 0 && ci < P.cin && co < P.cout) {
            const int tt = P.flip ? 26 - tap : tap;
            v = P.transposed_weights ? w[((size_t)ci * P.cout + co) * 27 + tt] : w[((size_t)co * P.cin + ci) * 27 + tt];
        }
        out[i] = f16 ? __half_as_ushort(__float2half_rn(v)) : __bfloat16_as_ushort(__float2bfloat16_rn(v));
    }
}

// ---- host: layer plan ------------------------------------------------------------------------------
struct KStep {
    int rd, rh, arr, col, chunk, lbo;    // A operand: depth slab, row offset (relative to S*th), array, column shift, chunk
    int cls;                             // deconv output parity class (pd*4 + ph*2 + pw), 0 otherwise
    KStepSrc src;
};

struct LayerGeom {
    int mode, cin_chunks, n, cout_tiles, arr;
    int nblk;                            // 3: stride-1 layers merge the three kh taps of an input row into one MMA
    int tmerged, pad_rows;               // T-merged: nblk = 9 (row tap 2,1,0 major, step tap 0,1,2 minor) [+ zero rows]
    int row_cols;                        // T-merged: B rows per row-tap group = TMEM columns per output row (3 n; 8 for n = 1)
    int n_grp;                           // T-merged: row-tap groups per B block: 3, or 4 when Cin = 8 pairs input rows (below)
    int n_t;                             // T-merged: step taps per group (3; 1 in flat 2D mode)
    std::vector<KStep> ks;
    std::vector<KStepSrc> srcs;          // [ks.size() * nblk]
};

static int round_up(int a, int b) { return (a + b - 1) / b * b; }

// The kernel's step axis is the real H axis and its row axis the real D axis (see ConvPlan::swap).
constexpr bool kStepAlongH = true;
// real tap index from the internal (step-axis k, row-axis k, kw)
static int tap_index(int k_step, int k_row, int kw)
{
    const int kd = kStepAlongH ? k_row : k_step, kh = kStepAlongH ? k_step : k_row;
    return (kd * 3 + kh) * 3 + kw;
}

static LayerGeom make_geom(int Cin, int Cout, int stride, int transposed, bool flat2d = false)
{
    LayerGeom g;
    g.mode = transposed ? (stride == 2 ? UM_DECONV_S2 : UM_CONV_S1) : (stride == 2 ? UM_CONV_S2 : UM_CONV_S1);
    g.cin_chunks = (Cin + 7) / 8;
    const int n_full = round_up(Cout, 16);
    // Cout tiles keep (packed weights + slab ring) within shared memory: <= 32 wide, 16 for Cin >= 64
    const int n_cap = g.cin_chunks >= 8 ? 16 : 32;
    g.n = n_full > n_cap ? n_cap : n_full;
    g.cout_tiles = (n_full + g.n - 1) / g.n;
    g.arr = g.mode == UM_CONV_S2 ? 2 : 1;
    const int CH = g.cin_chunks;
    const int plane = g.arr * UM_COLS;               // 16 B units between consecutive chunks of one staged row
    auto add = [&](int rd, int rh, int arr, int col, int chunk, int lbo, int cls, int tap0, int ch0, int tap1, int ch1) {
        KStep k{rd, rh, arr, col, chunk, lbo, cls, {{(int8_t)tap0, (int8_t)tap1}, {(int8_t)ch0, (int8_t)ch1}}};
        g.ks.push_back(k);
        g.srcs.push_back(k.src);
    };
    g.nblk = g.mode == UM_CONV_S1 ? 3 : 1;
    g.tmerged = 0; g.pad_rows = 0; g.row_cols = 0; g.n_grp = 0; g.n_t = 0;
    static const int no_tmerged = getenv("MVS_UMMA_NO_TMERGED") ? atoi(getenv("MVS_UMMA_NO_TMERGED")) : 0;   // A/B knob
    if (g.mode == UM_CONV_S1 && !no_tmerged) {
        // T-merged candidate: n = 8 for Cout <= 8 (N = rows*24 is padded to a multiple of 16 with 8 zero B rows), else
        // Cout tiles of n = 32 or 16 (Cout padded to 16).  Taken when the packed weights of one Cout tile + a 2-deep ring
        // of 3-row slabs fit shared memory (Cin = 64 -> 64: four tiles of 16).
        // Cout = 1 (`prob`, fp32 logits out): ONE column per (row tap, step tap) -- N = 16 per MMA (9 real columns + zero B
        // rows) instead of 72, three TMEM columns per output row instead of 24, so up to 16 rows fit a buffer
        int n = Cout == 1 ? 1 : (Cout <= 8 ? 8 : (n_full > 32 ? 32 : n_full));
        const int ksteps = CH == 1 ? 4 : 3 * ((CH + 1) / 2);
        const size_t slab3 = (size_t)3 * CH * UM_COLS * 16;
        // n = 1: a row-tap group is padded to 8 B rows / 8 TMEM columns -- the accumulator column of an MMA must stay 8-aligned
        // (three-column groups fault with "misaligned address")
        auto pad_of = [](int nn) { return nn <= 8 ? 8 : 0; };                     // zero B rows behind the tap groups
        const int n_t = flat2d ? 1 : 3;               // flat 2D: no step taps, a row-tap group is n rows / n TMEM columns
        auto cols_of = [&](int nn) { return nn == 1 ? 8 : n_t * nn; };
        auto wbytes_of = [&](int nn) { return (size_t)ksteps * 2 * ((CH == 1 ? 4 : 3) * cols_of(nn) + pad_of(nn)) * 16; };
        if (n == 32 && wbytes_of(32) + 2 * slab3 + 8192 > 226 * 1024) n = 16;
        const int pad = pad_of(n);
        const size_t wbytes = wbytes_of(n);
        if (wbytes + 2 * slab3 + 8192 <= 226 * 1024 && ksteps * (CH == 1 ? 4 : 3) * n_t <= UM_MAX_KSTEPS && !(flat2d && n == 1)) {
            g.tmerged = 1; g.pad_rows = pad; g.row_cols = cols_of(n); g.n = n; g.cout_tiles = Cout <= 8 ? 1 : (n_full + n - 1) / n;
            g.n_grp = CH == 1 ? 4 : 3;
            g.n_t = n_t;
            g.nblk = n_t * g.n_grp;
            // a k-step: K chunk 0 = tap kw0 of an input row, chunk 1 = tap kw1 of the same row (shift1 = 0) or of the NEXT
            // input row (shift1 = 1: its row-tap groups sit one group further along N); kw < 0 = zero weights
            auto add9 = [&](int col, int chunk, int lbo, int kw0, int ch0, int kw1, int ch1, int shift1 = 0) {
                KStep k{0, 0, 0, col, chunk, lbo, 0, {{0, 0}, {0, 0}}};
                g.ks.push_back(k);
                // flat 2D: the groups carry the kh taps of the centre depth slice (kd = 1), there is no step tap
                auto tap_of = [&](int t, int krow, int kw) { return flat2d ? tap_index(krow, 1, kw) : tap_index(t, krow, kw); };
                for (int grp = 0; grp < g.n_grp; ++grp)
                    for (int t = 0; t < n_t; ++t) {
                        const int krow0 = 2 - grp, krow1 = 2 - (grp - shift1);          // row taps 2, 1, 0 along the groups
                        const int t0 = (kw0 >= 0 && krow0 >= 0 && krow0 <= 2) ? tap_of(t, krow0, kw0) : -1;
                        const int t1 = (kw1 >= 0 && krow1 >= 0 && krow1 <= 2) ? tap_of(t, krow1, kw1) : -1;
                        KStepSrc sc{{(int8_t)t0, (int8_t)t1}, {(int8_t)ch0, (int8_t)ch1}};
                        g.srcs.push_back(sc);
                    }
            };
            if (CH == 1) {
                // Cin = 8: a tap is half a K = 16 step.  Two input rows i, i+1 share three steps instead of four:
                //   ks[0] (i: kw 0, 1)   ks[2] (i: kw 2 | i+1: kw 0, A chunk 1 one staged row further)   ks[3] (i+1: kw 1, 2)
                // and a last unpaired row takes ks[0] + ks[1] (kw 2 | zeros).  The shared step updates FOUR output rows.
                add9(0, 0, 1, 0, 0, 1, 0);                    // ks[0]: kw 0, 1 on adjacent staged columns (LBO = 16 B)
                add9(2, 0, 1, 2, 0, -1, 0);                   // ks[1]: kw 2 | zero weights
                add9(2, 0, UM_COLS - 2, 2, 0, 0, 0, 1);       // ks[2]: kw 2 of row i | kw 0 of row i + 1
                add9(1, 0, 1, 1, 0, 2, 0);                    // ks[3]: kw 1, 2
            } else {
                for (int kw = 0; kw < 3; ++kw)
                    for (int sidx = 0; sidx < CH; sidx += 2) {
                        const bool pair = sidx + 1 < CH;
                        add9(kw, sidx, pair ? plane : 1, kw, sidx, pair ? kw : -1, sidx + 1);
                    }
            }
            return g;
        }
    }
    if (g.mode == UM_CONV_S1) {
        // kh-merged: a k-step is (kd, kw-group, cin pair); its B block holds the kh = 2, 1, 0 taps side by side,
        // so ONE MMA on input row i feeds output rows i-2 .. i (the A operand is fetched once for three rows)
        for (int kd = 0; kd < 3; ++kd) {
            auto tap = [&](int kh, int kw) { return tap_index(kd, kh, kw); };
            auto add3 = [&](int col, int chunk, int lbo, int kw0, int ch0, int kw1, int ch1) {
                KStep k{kd, 0, 0, col, chunk, lbo, 0, {{0, 0}, {0, 0}}};
                g.ks.push_back(k);
                for (int kh = 2; kh >= 0; --kh) {
                    KStepSrc sc{{(int8_t)tap(kh, kw0), (int8_t)(kw1 >= 0 ? tap(kh, kw1) : -1)}, {(int8_t)ch0, (int8_t)ch1}};
                    g.srcs.push_back(sc);
                }
            };
            if (CH == 1) {
                add3(0, 0, 1, 0, 0, 1, 0);        // taps kw = 0, 1 on adjacent staged columns (LBO = 16 B)
                add3(2, 0, 1, 2, 0, -1, 0);       // tap kw = 2 paired with zero weights
            } else {
                for (int kw = 0; kw < 3; ++kw)
                    for (int sidx = 0; sidx < CH; sidx += 2) {
                        const bool pair = sidx + 1 < CH;
                        add3(kw, sidx, pair ? plane : 1, kw, sidx, pair ? kw : -1, sidx + 1);
                    }
            }
        }
    } else if (g.mode == UM_CONV_S2) {
        for (int kd = 0; kd < 3; ++kd)
            for (int kh = 0; kh < 3; ++kh) {
                auto tap = [&](int kw) { return tap_index(kd, kh, kw); };
                // (array, column shift) of tap kw: kw=0 -> O[c], 1 -> E[c], 2 -> O[c+1]
                const int arr_of[3] = {1, 0, 1};
                const int col_of[3] = {0, 0, 1};
                if (CH == 1) {
                    add(kd, kh, 1, 0, 0, 1, 0, tap(0), 0, tap(2), 0);      // O[c], O[c+1]
                    add(kd, kh, 0, 0, 0, 1, 0, tap(1), 0, -1, 0);          // E[c], (E[c+1] x 0)
                } else {
                    for (int kw = 0; kw < 3; ++kw)
                        for (int s = 0; s < CH; s += 2) {
                            const bool pair = s + 1 < CH;
                            add(kd, kh, arr_of[kw], col_of[kw], s, pair ? plane : 1, 0, tap(kw), s, pair ? tap(kw) : -1, s + 1);
                        }
                }
            }
    } else {
        // transposed stride 2: o = 2i - 1 + k.  parity 0: k=1 (i = o/2); parity 1: k=0 (i+1), k=2 (i).
        for (int pd = 0; pd < 2; ++pd)
            for (int ph = 0; ph < 2; ++ph)
                for (int pw = 0; pw < 2; ++pw)
                    for (int kd = 0; kd < 3; ++kd) {
                        if ((pd == 0) != (kd == 1)) continue;
                        for (int kh = 0; kh < 3; ++kh) {
                            if ((ph == 0) != (kh == 1)) continue;
                            for (int kw = 0; kw < 3; ++kw) {
                                if ((pw == 0) != (kw == 1)) continue;
                                const int t = tap_index(kd, kh, kw);
                                for (int s = 0; s < CH; s += 2) {
                                    const bool pair = s + 1 < CH;
                                    add(kd == 0 ? 1 : 0, kh == 0 ? 1 : 0, 0, kw == 0 ? 1 : 0, s, pair ? plane : 1,
                                        pd * 4 + ph * 2 + pw, t, s, pair ? t : -1, s + 1);
                                }
                            }
                        }
                    }
    }
    return g;
}

static size_t plan_smem_bytes(int weight_units, int ring, int slab_units)
{
    return ((size_t)weight_units + (size_t)ring * slab_units) * 16 + (2 * UM_MAX_RING + 2 * UM_TBUFS) * 8 + 16 + 64 * 4 + UM_MAX_LINES * 8 + UM_MAX_ACC * 16;
}

static bool build_plan(ConvPlan &P, const LayerGeom &g, int B, int Cin, int Cout, int D, int H, int W, int stride,
                       int transposed, int flags, bool out_f32, bool has_skip, size_t &smem_bytes)
{
    (void)Cin; (void)stride; (void)transposed;
    memset(&P, 0, sizeof(P));
    P.B = B; P.D = D; P.H = H; P.W = W;
    const bool deconv = g.mode == UM_DECONV_S2;
    const bool merged = g.nblk == 3;                 // stride-1 conv: kh taps merged along N
    if (g.tmerged) {
        P.Do = D; P.Ho = H; P.Wo = W;
        P.cin_chunks = g.cin_chunks; P.cout = Cout; P.cout_chunks = (Cout + 7) / 8; P.n = g.n; P.cout_tiles = g.cout_tiles;
        P.mode = g.mode; P.arr = 1; P.tmerged = 1;
        P.relu = (flags & MVS_RELU) ? 1 : 0; P.out_f32 = out_f32; P.has_skip = has_skip; P.f16 = (flags & MVS_ACT_F16) ? 1 : 0;
        P.rd = 3; P.d_mul = 1;
        const int rows_pc = g.n_grp * g.row_cols + g.pad_rows, n3 = g.row_cols;
        const bool pair_rows = g.n_grp == 4;
        P.row_cols = g.row_cols;
        const int packed_units = (int)g.ks.size() * 2 * rows_pc;
        // rows per CTA: minimise the staged (and multiplied) rows, row_blocks * (ht + 2); ring: as deep as fits
        int best = 0, best_ring = 0;
        long long best_cost = -1;
        const bool flat2d = (flags & MVS_FLAT2D) != 0;   // rows of a slab = image rows; D (the step extent) is the image height
        P.flat2d = flat2d ? 1 : 0;
        const int rows_ext = flat2d ? D : H;             // extent the ht-row blocks tile
        for (int ht = (g.n == 1 || flat2d) ? UM_MAX_ACC : 8; ht >= 1; --ht) {
            if (ht > rows_ext && ht > 1) continue;
            const int buf_cols = round_up(ht * n3 + g.pad_rows, 16);
            if (UM_TBUFS * buf_cols > 512 || buf_cols > 256) continue;
            if (flat2d && ((ht + 1) / 2) * (g.n >> 3) > 8) continue;     // the flat epilogue keeps at most eight work items per thread
            if ((pair_rows ? 2 : (int)g.ks.size()) * (ht + 2) + 1 > UM_MAX_OPS) continue;
            int ring = 0;
            for (int r = UM_MAX_RING; r >= 2 && !ring; r -= 2)          // even: see the issuer role
                if (plan_smem_bytes(packed_units + 2 * buf_cols, r, (ht + 2) * g.cin_chunks * UM_COLS) <= 226 * 1024) ring = r;
            if (!ring) continue;
            // a 2-deep ring leaves each issuer one slot: its next slab cannot load while the current one is multiplied
            // (flat 2D: every block is a pipeline step of its own -- charge the per-step overhead, ~4 rows' worth)
            const long long cost = (long long)((rows_ext + ht - 1) / ht) * (ht + 2 + (flat2d ? 4 : 0)) * (ring < 4 ? 10 : 8);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = ht; best_ring = ring; }
        }
        if (!best) return false;
        P.ht = best; P.ring = best_ring; P.rh = P.ht + 2;
        P.slab_units = P.rh * g.cin_chunks * UM_COLS;
        P.buf_cols = round_up(P.ht * n3 + g.pad_rows, 16);
        P.zero_units = 2 * P.buf_cols;
        P.weight_units = packed_units;
        smem_bytes = plan_smem_bytes(P.weight_units + P.zero_units, P.ring, P.slab_units);
        P.n_acc = P.ht; P.acc_cols = P.buf_cols;
        int cols = 32;
        while (cols < UM_TBUFS * P.buf_cols) cols *= 2;
        P.tmem_cols = cols;
        P.h_mul = 1; P.h_base = -1; P.d_base = -1; P.w_step = 1; P.w_base[0] = -1; P.w_base[1] = -1;
        P.od_mul = 1; P.oh_mul = 1; P.w_mul = 1; P.steps = flat2d ? (D + P.ht - 1) / P.ht : P.Do;
        P.n_issuers = 1;
        auto entry = [&](int a_off, int a_lbo, int b_off, int b_lbo, int col, int n_mma, int accumulate) {
            return make_uint4((uint32_t)a_off | ((uint32_t)a_lbo << 16), (uint32_t)b_off | ((uint32_t)b_lbo << 16), (uint32_t)col,
                              (uint32_t)accumulate | ((uint32_t)(n_mma >> 3) << 8));
        };
        int n_ops = 0;
        // zero-initialise the slab's whole buffer (A: any staged finite bytes, B: the zero block), then one MMA per
        // (staged input row i, k-step): rows lo..hi of the tile get the row taps i - row, all three step taps at once
        P.ops[n_ops++] = entry(0, 1, packed_units, P.buf_cols, 0, P.buf_cols, 0);
        // one MMA: k-step k on staged input row i, updating output rows first .. first + span - 1 (clipped to the tile) with
        // the row-tap groups that line up with them; group 0 of the block belongs to output row i - 2
        // (skip: leading groups / output rows left out -- MVS_KD1 layers only have the centre row tap, group 1)
        bool ok = true;
        auto emit = [&](int k, int i, int span, int skip = 0) {
            const int first = i - 2 + skip;
            const int lo = first < 0 ? 0 : first, hi = first + span - 1 > P.ht - 1 ? P.ht - 1 : first + span - 1;
            if (hi < lo) return;
            const KStep &ks = g.ks[(size_t)k];
            const int cnt = hi - lo + 1, grp0 = lo - first + skip;
            const int a_off = (i * g.cin_chunks + ks.chunk) * UM_COLS + ks.col;
            const int b_off = k * 2 * rows_pc + grp0 * n3;
            // N is rounded up to a multiple of 16 (n = 8: +8 columns; n = 1: 8 / 24 -> 16 / 32): the extra columns either
            // read the zero pad rows of the B block or land in the spare columns behind the last row of this buffer
            const int n_mma = round_up(cnt * n3, 16);
            if (ks.lbo > 0x3FFF || a_off > 0xFFFF || b_off > 0x3FFF || rows_pc > 0x3FFF || n_ops >= UM_MAX_OPS ||
                lo * n3 + n_mma > P.buf_cols || n_mma > 256) { ok = false; return; }
            P.ops[n_ops++] = entry(a_off, ks.lbo, b_off, rows_pc, lo * n3, n_mma, 1);
        };
        if (flags & MVS_KD1) {
            // weights are zero outside the centre row tap (a 2D convolution whose images ride the row axis): input row i only
            // feeds output row i - 1 -- N = one group (two for the shared step of a row pair), and the halo rows 0 and
            // rh - 1 are not multiplied at all
            if (pair_rows) {
                for (int i = 1; i + 1 < P.rh; i += 2) {
                    if (i + 2 < P.rh) { emit(0, i, 1, 1); emit(2, i, 2, 1); emit(3, i + 1, 1, 1); }
                    else { emit(0, i, 1, 1); emit(1, i, 1, 1); }
                }
            } else {
                for (int i = 1; i + 1 < P.rh; ++i)
                    for (size_t k = 0; k < g.ks.size(); ++k) emit((int)k, i, 1, 1);
            }
        } else if (pair_rows) {
            for (int i = 0; i < P.rh; i += 2) {
                if (i + 1 < P.rh) { emit(0, i, 3); emit(2, i, 4); emit(3, i + 1, 3); }
                else { emit(0, i, 3); emit(1, i, 3); }
            }
        } else {
            for (int i = 0; i < P.rh; ++i)
                for (size_t k = 0; k < g.ks.size(); ++k) emit((int)k, i, 3);
        }
        if (!ok) return false;
        P.n_ops = n_ops;
        return true;
    }
    if (deconv) { P.Do = 2 * D; P.Ho = 2 * H; P.Wo = 2 * W; }
    else if (g.mode == UM_CONV_S2) { P.Do = (D - 1) / 2 + 1; P.Ho = (H - 1) / 2 + 1; P.Wo = (W - 1) / 2 + 1; }
    else { P.Do = D; P.Ho = H; P.Wo = W; }
    P.cin_chunks = g.cin_chunks; P.cout = Cout; P.cout_chunks = (Cout + 7) / 8; P.n = g.n; P.cout_tiles = g.cout_tiles;
    P.mode = g.mode; P.arr = g.arr;
    P.relu = (flags & MVS_RELU) ? 1 : 0; P.out_f32 = out_f32; P.has_skip = has_skip; P.f16 = (flags & MVS_ACT_F16) ? 1 : 0;
    const int S = g.mode == UM_CONV_S2 ? 2 : 1;
    const int acc_per_row = deconv ? 8 : 1;
    const int packed_units = (int)g.ks.size() * 2 * g.nblk * g.n;     // what mvs_conv3d_c8_pack_weights wrote per Cout tile
    P.rd = deconv ? 2 : 3;
    P.d_mul = deconv ? 1 : S;
    const int rows_h = deconv ? H : P.Ho;
    auto n_ops_for = [&](int ht) { return merged ? (int)g.ks.size() * (ht + 2) + 1 : (int)g.ks.size() * ht; };
    // ht (rows per CTA) and ring depth: prefer one full step of prefetch and two CTAs per SM, then relax.
    int best = 0, best_ring = 0;
    static const int first_pass = getenv("MVS_UMMA_FIRST_PASS") ? atoi(getenv("MVS_UMMA_FIRST_PASS")) : 0;   // tuning knob
    for (int pass = first_pass; pass < 3 && !best; ++pass) {
        const size_t budget = pass == 0 ? 112 * 1024 : 226 * 1024;
        const int ring = pass < 2 ? P.rd + P.d_mul : P.rd;
        for (int ht = 8; ht >= 1; --ht) {
            if (ht > rows_h && ht > 1) continue;
            // TMEM: two accumulator buffers per CTA; with two CTAs per SM (pass 0) each may take half of the 512 columns
            if (n_ops_for(ht) > UM_MAX_OPS || ht * acc_per_row > UM_MAX_ACC ||
                2 * ht * acc_per_row * g.n > (pass == 0 ? 256 : 512)) continue;
            const int rh = deconv ? ht + 1 : S * (ht - 1) + 3;
            const int zero_units = merged ? 2 * ht * g.n : 0;
            if (plan_smem_bytes(packed_units + zero_units, ring, rh * g.cin_chunks * g.arr * UM_COLS) > budget) continue;
            best = ht; best_ring = ring;
            break;
        }
    }
    if (!best) return false;
    P.ht = best;
    P.ring = best_ring;
    P.rh = deconv ? P.ht + 1 : S * (P.ht - 1) + 3;
    P.slab_units = P.rh * g.cin_chunks * g.arr * UM_COLS;
    P.zero_units = merged ? 2 * P.ht * g.n : 0;
    P.weight_units = packed_units;
    smem_bytes = plan_smem_bytes(P.weight_units + P.zero_units, P.ring, P.slab_units);
    P.n_acc = P.ht * acc_per_row;
    P.acc_cols = P.n_acc * g.n;
    int cols = 32;
    while (cols < 2 * P.acc_cols) cols *= 2;
    P.tmem_cols = cols;
    if (deconv) {
        P.h_mul = 1; P.h_base = 0; P.d_base = 0; P.w_step = 1; P.w_base[0] = 0; P.w_base[1] = 0;
        P.od_mul = 2; P.oh_mul = 2; P.w_mul = 2; P.steps = D;
    } else {
        P.h_mul = S; P.h_base = -1; P.d_base = -1; P.w_step = S;
        P.w_base[0] = S == 1 ? -1 : 0; P.w_base[1] = -1;
        P.od_mul = 1; P.oh_mul = 1; P.w_mul = 1; P.steps = P.Do;
    }
    for (int th = 0; th < P.ht; ++th)
        for (int c = 0; c < acc_per_row; ++c) {
            AccOut &ao = P.acc[th * acc_per_row + c];
            ao.th = (int8_t)th;
            ao.dd = deconv ? (int8_t)(c >> 2) : 0;
            ao.dh = deconv ? (int8_t)((c >> 1) & 1) : 0;
            ao.wadd = deconv ? (int8_t)(c & 1) : 0;
        }
    auto entry = [&](int a_off, int a_lbo, int b_off, int b_lbo, int col, int n_mma, int accumulate) {
        // low 14 bits + smem base stay below 2^14 (227 KB / 16), so no masking at issue time
        return make_uint4((uint32_t)a_off | ((uint32_t)a_lbo << 16), (uint32_t)b_off | ((uint32_t)b_lbo << 16), (uint32_t)col,
                          (uint32_t)accumulate | ((uint32_t)(n_mma >> 3) << 8));
    };
    int n_ops = 0;
    if (merged) {
        // ONE issuer (the merged MMAs of neighbouring input rows overlap in the accumulators they update).
        // Step layout: [zero-initialise all ht*n columns] then, per depth slab r = kd and staged input row i,
        // one MMA per (kw group, cin pair) that updates output rows max(0,i-2) .. min(ht-1,i) with the
        // kh = i - row taps: B sub-block [2 - (i - lo)] .. of the (kh = 2,1,0)-ordered packed block.
        P.merged = 1;
        P.n_issuers = 2;
        for (int iss = 0; iss < UM_MAX_ISSUERS; ++iss)
            for (int r = 0; r < 4; ++r) P.op_begin[iss][r] = 0;
        if (P.ht * g.n > 256) return false;
        P.ops[n_ops++] = entry(0, 1, packed_units, P.ht * g.n, 0, P.ht * g.n, 0);   // A: any staged (finite) bytes; B: zeros
        const int ks_per_kd = (int)g.ks.size() / 3;
        for (int r = 0; r < 3; ++r) {
            P.op_begin[0][r] = r == 0 ? 0 : n_ops;
            for (int i = 0; i < P.rh; ++i) {
                const int lo = i - 2 < 0 ? 0 : i - 2, hi = i > P.ht - 1 ? P.ht - 1 : i;
                if (hi < lo) continue;
                const int cnt = hi - lo + 1, blk0 = 2 - (i - lo);
                for (int k = r * ks_per_kd; k < (r + 1) * ks_per_kd; ++k) {
                    const KStep &ks = g.ks[(size_t)k];
                    const int a_off = ((i * g.cin_chunks + ks.chunk) * g.arr + ks.arr) * UM_COLS + ks.col;
                    const int b_off = k * 2 * 3 * g.n + blk0 * g.n;
                    if (ks.lbo > 0x3FFF || a_off > 0xFFFF || b_off > 0x3FFF) return false;
                    if (n_ops >= UM_MAX_OPS) return false;
                    P.ops[n_ops++] = entry(a_off, ks.lbo, b_off, 3 * g.n, lo * g.n, cnt * g.n, 1);
                }
            }
        }
        P.op_begin[0][3] = n_ops;
        for (int iss = 1; iss < UM_MAX_ISSUERS; ++iss)
            for (int r = 0; r < 4; ++r) P.op_begin[iss][r] = n_ops;
    } else {
        // ops per accumulator; issuer j owns accumulators {a : a % n_issuers == j}; within an issuer the ops
        // are grouped by the depth slab they read (rd), round-robin over its accumulators inside a group
        // (independent chains pipeline in the tensor core).  The first op an accumulator sees overwrites it.
        std::vector<std::vector<MmaOp>> per_acc((size_t)P.n_acc);
        for (int th = 0; th < P.ht; ++th)
            for (size_t k = 0; k < g.ks.size(); ++k) {
                const KStep &ks = g.ks[k];
                MmaOp op;
                const int row = S * th + ks.rh;
                const int a_off = ((row * g.cin_chunks + ks.chunk) * g.arr + ks.arr) * UM_COLS + ks.col;
                const int acc = th * acc_per_row + ks.cls;
                op.a_off = (uint16_t)a_off;
                op.b_off = (uint16_t)(k * 2 * g.n);
                op.a_lbo = (uint16_t)ks.lbo;
                op.acc = (uint8_t)acc;
                op.rd_first = (uint8_t)ks.rd;
                if (ks.lbo > 0x3FFF || a_off > 0xFFFF || k * 2 * g.n > 0x3FFF) return false;
                per_acc[(size_t)acc].push_back(op);
            }
        P.n_issuers = P.n_acc < UM_MAX_ISSUERS ? P.n_acc : UM_MAX_ISSUERS;
        std::vector<char> started((size_t)P.n_acc, 0);
        for (int iss = 0; iss < UM_MAX_ISSUERS; ++iss) {
            for (int r = 0; r < 3; ++r) {
                P.op_begin[iss][r] = n_ops;
                if (iss >= P.n_issuers) continue;
                std::vector<std::vector<MmaOp>> sel;
                for (int a = iss; a < P.n_acc; a += P.n_issuers) {
                    std::vector<MmaOp> v;
                    for (const MmaOp &op : per_acc[(size_t)a]) if ((op.rd_first & 3) == r) v.push_back(op);
                    sel.push_back(v);
                }
                for (size_t j = 0;; ++j) {
                    bool any = false;
                    for (const auto &v : sel) {
                        if (j >= v.size()) continue;
                        const MmaOp &op = v[j];
                        const int accumulate = started[op.acc] ? 1 : 0;
                        started[op.acc] = 1;
                        P.ops[n_ops++] = entry(op.a_off, op.a_lbo, op.b_off, g.n, op.acc * g.n, g.n, accumulate);
                        any = true;
                    }
                    if (!any) break;
                }
            }
            P.op_begin[iss][3] = n_ops;
        }
    }
    P.n_ops = n_ops;
    return true;
}

}  // namespace mvs

using namespace mvs;

static thread_local long long *g_trace = nullptr;
static thread_local int g_trace_ctas = 0;

extern "C" int mvs_conv3d_c8_set_trace(void *dev_buf, int n_ctas)
{
    g_trace = (long long *)dev_buf;
    g_trace_ctas = dev_buf ? n_ctas : 0;
    return MVS_OK;
}

extern "C" int64_t mvs_conv3d_c8_packed_weight_bytes(int Cin, int Cout, int stride, int transposed)
{
    if (Cin <= 0 || Cout <= 0 || (stride != 1 && stride != 2)) return -1;
    const LayerGeom g = make_geom(Cin, Cout, stride, transposed);
    return (int64_t)g.cout_tiles * (int64_t)g.ks.size() * 2 * ((g.tmerged ? g.n_grp * g.row_cols : g.nblk * g.n) + g.pad_rows) * 16;
}

extern "C" int mvs_conv3d_c8_pack_weights(const float *w, void *packed, int Cin, int Cout, int stride, int transposed,
                                          void *stream)
{
    return mvs_conv3d_c8_pack_weights_ex(w, packed, Cin, Cout, stride, transposed, 0, stream);
}

extern "C" int mvs_conv3d_c8_pack_weights_ex(const float *w, void *packed, int Cin, int Cout, int stride, int transposed,
                                             int flags, void *stream)
{
    MVS_REQUIRE(w && packed, "null pointer");
    MVS_REQUIRE(Cin > 0 && Cout > 0 && (stride == 1 || stride == 2), "bad layer shape");
    const bool flat2d = (flags & MVS_FLAT2D) != 0;
    MVS_REQUIRE(!flat2d || (stride == 1 && Cout > 1), "MVS_FLAT2D: stride-1 layers with a C8 output only");
    const LayerGeom g = make_geom(Cin, Cout, stride, transposed, flat2d);
    MVS_REQUIRE(!flat2d || g.tmerged, "MVS_FLAT2D: this layer shape has no T-merged plan");
    MVS_REQUIRE((int)g.srcs.size() <= UM_MAX_KSTEPS, "too many k-steps for this layer (Cin too large)");
    PackPlan pp;
    memset(&pp, 0, sizeof(pp));
    pp.cin = Cin; pp.cout = Cout; pp.n = g.n; pp.cout_tiles = g.cout_tiles; pp.n_ksteps = (int)g.ks.size(); pp.nblk = g.nblk;
    pp.transposed_weights = transposed ? 1 : 0;
    pp.flip = (transposed && stride == 1) ? 1 : 0;       // ConvTranspose3d(stride 1, pad 1) == conv with flipped taps
    for (size_t k = 0; k < g.srcs.size(); ++k) pp.ks[k] = g.srcs[k];
    pp.pad_rows = g.pad_rows;
    pp.grp_rows = g.tmerged ? g.row_cols : 0;
    pp.n_grp = g.n_grp;
    pp.n_t = g.n_t;
    const long long total = (long long)g.cout_tiles * pp.n_ksteps * 2 * ((g.tmerged ? g.n_grp * g.row_cols : g.nblk * g.n) + g.pad_rows) * 8;
    pack_weights_kernel<<<cdiv(total, 256) > 1024 ? 1024 : cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
        pp, w, (uint16_t *)packed, (flags & MVS_ACT_F16) ? 1 : 0);
    return check_launch("mvs_conv3d_c8_pack_weights");
}

extern "C" int mvs_conv3d_c8_fwd(const void *x_c8, const void *w_packed, const float *scale, const float *shift,
                                 const void *skip_c8, void *y, int B, int Cin, int Cout, int D, int H, int W, int stride,
                                 int transposed, int flags, void *stream)
{
    if (B == 0 || D == 0 || H == 0 || W == 0) return MVS_OK;
    MVS_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && D > 0 && H > 0 && W > 0, "extents must be positive");
    MVS_REQUIRE(stride == 1 || stride == 2, "stride must be 1 or 2");
    MVS_REQUIRE(x_c8 && w_packed && y, "null pointer");
    const bool flat2d = (flags & MVS_FLAT2D) != 0;
    const bool skip_ps = (flags & MVS_SKIP_PS) != 0;
    MVS_REQUIRE(!flat2d || (D == 1 && stride == 1 && Cout > 1 && (!skip_c8 || skip_ps) && !(flags & (MVS_Y_DW | MVS_SKIP_DW))),
                "MVS_FLAT2D: D = 1, stride 1, C8 output, natural W order; a skip operand only as MVS_SKIP_PS");
    MVS_REQUIRE(!skip_ps || (flat2d && skip_c8 && H % 2 == 0 && W % 2 == 0),
                "MVS_SKIP_PS: flat 2D layers with a skip operand and even H, W only");
    const LayerGeom g = make_geom(Cin, Cout, stride, transposed, flat2d);
    MVS_REQUIRE(!flat2d || g.tmerged, "MVS_FLAT2D: this layer shape has no T-merged plan");
    MVS_REQUIRE((int)g.srcs.size() <= UM_MAX_KSTEPS, "too many k-steps for this layer (Cin too large)");
    static thread_local ConvPlan P;
    size_t smem = 0;
    const bool out_f32 = Cout == 1;
    // internal (step, row) axes = real (H, D): the step axis is always the long one, chunked per CTA below
    const int Di = kStepAlongH ? H : D, Hi = kStepAlongH ? D : H;
    if (!build_plan(P, g, B, Cin, Cout, Di, Hi, W, stride, transposed, flags, out_f32, skip_c8 != nullptr, smem))
        return fail(MVS_ERR_UNSUPPORTED, "mvs_conv3d_c8_fwd: no tile configuration fits shared memory / TMEM for this layer");
    MVS_REQUIRE(!(out_f32 && skip_c8), "skip is not supported on the fp32 (Cout == 1) output");
    P.x_dw = (flags & MVS_X_DW) ? 1 : 0; P.y_dw = (flags & MVS_Y_DW) ? 1 : 0; P.skip_dw = (flags & MVS_SKIP_DW) ? 1 : 0;
    MVS_REQUIRE(!P.x_dw || g.mode == UM_CONV_S2, "MVS_X_DW: only the stride-2 layers read a W-de-interleaved input");
    MVS_REQUIRE(!P.y_dw || (P.tmerged && !out_f32), "MVS_Y_DW: only the stride-1 C8 layers write a W-de-interleaved output");
    MVS_REQUIRE(!P.skip_dw || skip_c8, "MVS_SKIP_DW without a skip tensor");
    MVS_REQUIRE(!P.skip_dw || P.tmerged || g.mode == UM_DECONV_S2, "MVS_SKIP_DW: stride-1 and transposed stride-2 layers only");
    P.pair_store = (g.mode == UM_DECONV_S2 && Cout <= 8 && !out_f32 && ((uintptr_t)y & 31) == 0 && P.n_acc % 2 == 0) ? 1 : 0;
    P.trace = g_trace; P.trace_ctas = g_trace_ctas;
    P.ring_magic = (1u << 18) / (uint32_t)P.ring + 1u;
    MVS_REQUIRE((long long)P.Do * P.Ho * P.Wo * (P.n >> 3 > 0 ? P.n >> 3 : 1) < (1ll << 31), "output volume too large for 32-bit tile offsets");
    MVS_REQUIRE(2 * P.steps + 4 < 32768, "step axis too long for the slab-ring arithmetic");
    P.swap = kStepAlongH ? 1 : 0;
    P.Dr = D; P.Hr = H;
    P.Dor = kStepAlongH ? P.Ho : P.Do;
    P.Hor = kStepAlongH ? P.Do : P.Ho;
    const int rows_h = g.mode == UM_DECONV_S2 ? Hi : P.Ho;
    const int m_ext = g.mode == UM_DECONV_S2 ? W : P.Wo;
    P.row_blocks = cdiv(rows_h, P.ht);
    // chunk the step axis: aim for ~8 waves of CTAs (wave quantisation costs up to a whole wave otherwise),
    // but keep chunks >= 8 steps (>= 4 when the layer is too small to fill the GPU) so pipeline fill amortises
    {
        const long long base_ctas = (long long)cdiv(m_ext, 128) * P.row_blocks * B * P.cout_tiles;
        static const int waves_env = getenv("MVS_UMMA_WAVES") ? atoi(getenv("MVS_UMMA_WAVES")) : 0;   // tuning knob
        const int n_sm = sm_count();
        const long long target = (waves_env > 0 ? (long long)waves_env : 8LL * 2) * n_sm;
        long long chunks = (target + base_ctas - 1) / base_ctas;
        static const int min_steps_env = getenv("MVS_UMMA_MIN_STEPS") ? atoi(getenv("MVS_UMMA_MIN_STEPS")) : 0;
        // layers too small to fill the GPU even with 4-step chunks (the 1/8-resolution bottleneck) go down to 2 steps
        const int min_steps = min_steps_env > 0 ? min_steps_env
                              : (base_ctas * (P.steps / 4 > 0 ? P.steps / 4 : 1) < n_sm ? 2
                                 : (base_ctas * (P.steps / 8 > 0 ? P.steps / 8 : 1) < 2 * n_sm ? 4 : 8));
        const long long max_chunks = P.steps / min_steps > 1 ? P.steps / min_steps : 1;
        if (chunks > max_chunks) chunks = max_chunks;
        if (chunks < 1) chunks = 1;
        P.steps_per_cta = cdiv(P.steps, chunks);
        if (P.tmerged && waves_env == 0) {
            // every CTA stages steps + 2 slabs and pays a pipeline fill (~2 steps) before its first store: pick the
            // chunking that minimises waves x (steps per CTA + 4)
            const long long slots = (long long)n_sm * (UM_TBUFS * P.buf_cols <= 256 && smem <= 112 * 1024 ? 2 : 1);
            long long best_cost = -1;
            for (int c = 1; c <= P.steps; ++c) {
                const int spc = cdiv(P.steps, c);
                if (spc < 4 && c > 1) break;
                const long long ctas = base_ctas * cdiv(P.steps, spc);
                const long long cost = cdiv(ctas, slots) * (spc + 4);
                if (best_cost < 0 || cost < best_cost) { best_cost = cost; P.steps_per_cta = spc; }
            }
        }
    }
    const int step_chunks = cdiv(P.steps, P.steps_per_cta);
    dim3 grid(cdiv(m_ext, 128), (unsigned)(P.row_blocks * step_chunks), B * P.cout_tiles);
    MVS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "grid too large");
    auto launch = [&](auto kernel, int threads) -> cudaError_t {
        cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        // launched with programmatic stream serialization: the prologue of this layer overlaps the tail of the previous
        // kernel in the stream (see griddepcontrol.wait in the kernel); MVS_PDL=0 switches it off (A/B knob)
        static const bool pdl = !(getenv("MVS_PDL") && atoi(getenv("MVS_PDL")) == 0);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid; cfg.blockDim = dim3((unsigned)threads); cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
        return cudaLaunchKernelEx(&cfg, kernel, P, (const uint4 *)x_c8, (const uint4 *)w_packed, scale, shift,
                                  (const uint4 *)skip_c8, y);
    };
    cudaError_t e;
    if (P.tmerged && P.flat2d) e = skip_c8 ? launch(conv3d_umma_kernel<true, 5>, UM_THREADS_TM) : launch(conv3d_umma_kernel<true, 4>, UM_THREADS_TM);
    else if (P.tmerged) e = skip_c8 ? launch(conv3d_umma_kernel<true, 2>, UM_THREADS_TM) : launch(conv3d_umma_kernel<true, 1>, UM_THREADS_TM);
    else if (skip_c8) e = launch(conv3d_umma_kernel<false, 2>, UM_THREADS + UM_EPI_THREADS);
    else e = launch(conv3d_umma_kernel<false, 3>, UM_THREADS);
    if (e != cudaSuccess) return fail(MVS_ERR_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    return check_launch("mvs_conv3d_c8_fwd");
}
