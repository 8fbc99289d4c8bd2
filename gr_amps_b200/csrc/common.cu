#include "common.h"
#include <cstdio>
#include <cstring>

namespace amps {

static thread_local char g_err[512] = "";

int set_error(int status, const char *msg) {
    std::snprintf(g_err, sizeof g_err, "%s", msg ? msg : "");
    return status;
}

int set_cuda_error(cudaError_t e, const char *what) {
    std::snprintf(g_err, sizeof g_err, "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    cudaGetLastError();   // clear the sticky-less error state
    return e == cudaErrorMemoryAllocation ? AMPS_E_NOMEM : AMPS_E_CUDA;
}

int select_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        std::fprintf(stderr, "libamps_b200: no CUDA device available (%s); this library has no CPU fallback\n",
                     e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
        cudaGetLastError();
        return set_error(AMPS_E_NODEVICE, "no CUDA device available; libamps_b200 has no CPU fallback");
    }
    if (device < 0 || device >= n) return set_error(AMPS_E_INVAL, "device ordinal out of range");
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaGetDeviceProperties");
    if (prop.major != 10) {
        std::fprintf(stderr, "libamps_b200: device %d is sm_%d%d; the kernels are built for sm_100a only\n", device,
                     prop.major, prop.minor);
        return set_error(AMPS_E_NODEVICE, "device is not an sm_100 (Blackwell B200) part");
    }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaSetDevice");
    return AMPS_OK;
}

}  // namespace amps

extern "C" int amps_b200_version(void) { return 100; }

extern "C" int amps_b200_abi_sizes(size_t *burst_bytes, size_t *words_bytes) {
    if (burst_bytes) *burst_bytes = sizeof(amps_burst);
    if (words_bytes) *words_bytes = sizeof(amps_recc_words);
    return AMPS_OK;
}

extern "C" const char *amps_b200_last_error(void) { return amps::g_err; }

extern "C" const char *amps_b200_strerror(int status) {
    switch (status) {
        case AMPS_OK: return "ok";
        case AMPS_E_INVAL: return "invalid argument";
        case AMPS_E_NODEVICE: return "no usable sm_100 CUDA device";
        case AMPS_E_CUDA: return "CUDA runtime error";
        case AMPS_E_NOMEM: return "out of memory";
        case AMPS_E_ALIGN: return "alignment requirement not met";
        case AMPS_E_OVERFLOW: return "capacity exceeded";
        case AMPS_E_STATE: return "call not valid in the handle's current state";
        default: return "unknown status";
    }
}

extern "C" int amps_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int i = 0; i < n; ++i) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, i) == cudaSuccess && prop.major == 10) ++ok;
    }
    return ok;
}
