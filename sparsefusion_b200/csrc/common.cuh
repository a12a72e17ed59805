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

// Programmatic dependent launch (PDL).  The UNet evaluation is ~350 short kernels replayed from one CUDA graph; each of them
// (a) is launched with programmatic stream serialisation, so its CTAs may become resident while the previous kernel drains, and
// (b) starts with pdl_sync(): signal that the NEXT kernel may be scheduled, then wait until every kernel before this one has completed
// and flushed its memory.  All global reads and writes of a kernel come after that wait, so ordering is unchanged.
bool pdl_enabled();
// which single-launch fused variants are enabled (sfb_set_fusion): bit 0 GroupNorm cluster kernel, bit 1 global-context pooling cluster kernel
int fusion_mask();
static inline bool gn_fused_enabled() { return (fusion_mask() & 1) != 0; }
static inline bool gca_fused_enabled() { return (fusion_mask() & 2) != 0; }
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() { pdl_trigger(); pdl_wait(); }

// cluster_x > 1 launches thread-block clusters of (cluster_x, 1, 1); grid.x must be a multiple of it
template <typename... KA, typename... A>
static inline cudaError_t launch_pdl_cluster(void (*kernel)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, unsigned cluster_x,
                                             A&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.numAttrs = 1;
    if (cluster_x > 1) {
        attr[1].id = cudaLaunchAttributeClusterDimension;
        attr[1].val.clusterDim.x = cluster_x;
        attr[1].val.clusterDim.y = 1;
        attr[1].val.clusterDim.z = 1;
        cfg.numAttrs = 2;
    }
    cfg.attrs = attr;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KA>(args)...);
}
template <typename... KA, typename... A>
static inline cudaError_t launch_pdl(void (*kernel)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
    return launch_pdl_cluster(kernel, grid, block, smem, st, 1u, static_cast<A&&>(args)...);
}
#endif


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
