// common.h -- error plumbing shared by the C-ABI translation units of libamps_b200.
#pragma once
#include <cuda_runtime.h>
#include "../../include/amps_b200.h"

namespace amps {
int set_error(int status, const char *msg);
int set_cuda_error(cudaError_t e, const char *what);
int select_device(int device);     // cudaSetDevice + "is this an sm_100 part" check
}

// NVTX range over one C-ABI call (header-only NVTX 3: no library to link; a no-op unless a profiler is attached)
#include <nvtx3/nvToolsExt.h>
namespace amps {
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};
}
#define AMPS_NVTX(name) amps::NvtxRange amps_nvtx_range_(name)

#define CK(call)                                                                  \
    do {                                                                          \
        cudaError_t ck_e_ = (call);                                               \
        if (ck_e_ != cudaSuccess) return amps::set_cuda_error(ck_e_, #call);      \
    } while (0)
#define CKL(call) CK(call)
