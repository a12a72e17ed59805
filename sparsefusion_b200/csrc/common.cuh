// common.cuh -- shared helpers for the sm_100a kernels behind the C-ABI (include/sparsefusion_b200.h)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include "../../include/sparsefusion_b200.h"  // SFB_OK / SFB_ERR_* codes

namespace sfb {

// last error message, per host thread; read through sfb_last_error()
std::string& last_error();
int fail(int code, const char* fmt, ...);


static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

template <typename T>
static inline __host__ __device__ T ceil_div(T a, T b) { return (a + b - 1) / b; }

int check_launch(const char* what);   // also counts one kernel launch (sfb_launch_count)
void count_launches(int n);            // extra launches made under a single check_launch

// tensor-core operand precision of the UNet GEMMs: 0 = TF32 single pass, 1 = 3xTF32 (error compensated)
int precision_mode();
void set_precision_mode(int m);

// Every kernel of the library asks for the maximum shared-memory carve-out, whether it needs it or not: the tcgen05 GEMMs use ~200 KB of
// smem, and alternating them with small-smem kernels would otherwise make the SMs re-partition L1/smem between consecutive launches.
void prefer_smem(const void* kernel);
template <typename... A>
static inline auto kernel_with_carveout(void (*k)(A...)) -> void (*)(A...) {
    prefer_smem((const void*)k);
    return k;
}
}  // namespace sfb
#define SFB_K(...) sfb::kernel_with_carveout(__VA_ARGS__)
namespace sfb {

// number of SMs of the current device (cached)
int sm_count();

}  // namespace sfb

#define SFB_REQUIRE(cond, ...)                                   \
    do {                                                         \
        if (!(cond)) return sfb::fail(SFB_ERR_ARG, __VA_ARGS__); \
    } while (0)

#define SFB_CUDA(call)                                                                                  \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess) return sfb::fail(SFB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
    } while (0)
