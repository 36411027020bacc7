// bf16 C8 3x3x3 convolution on the 5th-generation tensor cores (tcgen05.mma, TMEM accumulator,
// TMA-staged bricks).  UNDER CONSTRUCTION in this revision: the entry points exist so that the ABI
// is complete and callers fail loudly (MVS_ERR_UNSUPPORTED) instead of silently taking another path.
#include "common.cuh"

using namespace mvs;

extern "C" int64_t mvs_conv3d_c8_packed_weight_bytes(int Cin, int Cout, int stride, int transposed)
{
    (void)stride; (void)transposed;
    const int64_t cin8 = (Cin + 7) / 8 * 8, cout16 = (Cout + 15) / 16 * 16;
    return cin8 * cout16 * 27 * 2;
}

extern "C" int mvs_conv3d_c8_pack_weights(const float *, void *, int, int, int, int, void *)
{
    return fail(MVS_ERR_UNSUPPORTED, "mvs_conv3d_c8_pack_weights: tcgen05 conv path not built in this revision");
}

extern "C" int mvs_conv3d_c8_fwd(const void *, const void *, const float *, const float *, const void *, void *, int, int,
                                 int, int, int, int, int, int, int, void *)
{
    return fail(MVS_ERR_UNSUPPORTED, "mvs_conv3d_c8_fwd: tcgen05 conv path not built in this revision");
}
