// Shared helpers for the sm_100a kernels behind include/mvs_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>

#include "../../include/mvs_b200.h"

namespace mvs {

std::string &last_error_ref();                 // thread-local storage lives in abi.cu
extern std::atomic<long long> g_launches;

inline int fail(int code, const std::string &msg)
{
    last_error_ref() = msg;
    return code;
}

inline int check_launch(const char *what)
{
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(MVS_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return MVS_OK;
}

#define MVS_REQUIRE(cond, msg)                                                                   \
    do {                                                                                         \
        if (!(cond)) return ::mvs::fail(MVS_ERR_INVALID, std::string(__func__) + ": " + (msg));  \
    } while (0)

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

int sm_count();        // multiProcessorCount of the current device, queried once per device (abi.cu); 148 when unknown

struct SrcPtrs {
    const void *p[MVS_MAX_SRC];
};
struct DstPtrs {
    void *p[MVS_MAX_SRC];
};

}  // namespace mvs
