// common.cu -- error reporting and device queries shared by every translation unit
#include "common.cuh"
#include <stdarg.h>
#include "../../include/sparsefusion_b200.h"

namespace sfb {

std::string& last_error() {
    static thread_local std::string msg;
    return msg;
}

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error() = buf;
    return code;
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(SFB_ERR_CUDA, "%s: launch failed: %s", what, cudaGetErrorString(e));
    return SFB_OK;
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace sfb

extern "C" {

const char* sfb_last_error(void) { return sfb::last_error().c_str(); }

int sfb_abi_version(void) { return SFB_ABI_VERSION; }

int sfb_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    SFB_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    SFB_CUDA(cudaGetDeviceProperties(&p, dev));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return SFB_OK;
}
}
