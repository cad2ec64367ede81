// Library-level entry points: version, thread-local error text, device check.
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace factk {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return FACTK_ERR_CUDA;
    }
    return FACTK_OK;
}

}  // namespace factk

extern "C" int factk_version(void) { return 200; }   // 2.0: training step (backward kernels, tcgen05 weight gradient, GRU BPTT, loss gradients), tcgen05 attention

extern "C" const char* factk_last_error(void) { return factk::g_err; }

extern "C" int factk_device_check(void) {
    int dev = 0;
    cudaDeviceProp p;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
        factk::set_error("factk_device_check: no CUDA device");
        cudaGetLastError();
        return FACTK_ERR_DEVICE;
    }
    if (p.major != 10) {
        factk::set_error("factk_device_check: device is sm_%d%d, this library is built for sm_100a only", p.major, p.minor);
        return FACTK_ERR_DEVICE;
    }
    return FACTK_OK;
}
