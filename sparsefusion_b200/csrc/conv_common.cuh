// conv_common.cuh -- definitions shared by the two generations of the tcgen05 implicit-GEMM convolution kernel
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sfb {

constexpr int kBM = 128;           // output pixels per CTA (UMMA M)
constexpr int kBK = 32;            // fp32 elements per k-step: one 128-byte swizzle row
constexpr int kABytes = kBM * 128; // 16 KB per stage
constexpr int kThreads = 192;

struct alignas(64) ConvGemmParams {
    CUtensorMap tmA[4];
    CUtensorMap tmB;
    CUtensorMap tmBlo;      // pre-split weights (non-swap v2 only): tmB = hi = w & 0xFFFFE000, tmBlo = lo = w - hi
    float* out;
    const float* bias;
    const float* residual;  // optional [pixel][ldr] tensor added to the result (by split 0)
    int64_t ldo;            // floats between consecutive output pixels
    int64_t ldr;
    int32_t NB, Ho, Wo;
    int32_t TW, TH, TN;
    int32_t tiles_w, tiles_h;
    int32_t Cout, cin_chunks;
    int32_t KH, KW, pad, stride;
    int32_t k_iters, splits;
    int32_t accumulate;
    int32_t presplit;       // tmB / tmBlo hold (hi, lo) of the weights: the converter leaves the N-side tile alone
    int32_t w_dynamic;      // the "weight" operand was produced by an earlier kernel of this stream (attention GEMMs): no prefetch before griddepcontrol.wait
    int32_t raw_hi;         // experiment (variant 3): feed the un-masked fp32 word as the "hi" tensor-core operand (is the hardware's tf32 read a truncation?)
};

// per-launch CUDA-event timing hooks (bench.py roofline line)
int conv_prof_begin(cudaStream_t st);
int conv_prof_end(cudaStream_t st);

// v2: M-side operand through tensor memory, optional swap-AB (conv_tcgen05_v2.cu)
int launch_conv_v2(const ConvGemmParams& p, int BN, bool swap, dim3 grid, cudaStream_t st);

}  // namespace sfb
