// conv_tcgen05.cu -- NHWC convolution / linear layers of the VLDM UNet as implicit GEMM on the
// 5th-generation tensor cores (tcgen05.mma kind::tf32, fp32 accumulation in TMEM), operands staged by
// TMA into 128B-swizzled shared memory.  Replaces the cuDNN/cuBLAS calls behind nn.Conv2d / nn.Linear in
// the reference's Unet (external/imagen_pytorch.py:641-662 Block.project, :708 res_conv, :608-610
// Downsample, :578-606 PixelShuffleUpsample conv, :1017-1042 CrossEmbedLayer, :953-961 ChanFeedForward,
// :1386 final_conv), which the reference GPU build runs in TF32 (torch 1.11 default).
//
// GEMM view:   D[pixel, cout] = sum_{tap, c} X[pixel shifted by tap, c] * W[cout, tap, c]
//   A operand = activations, M = 128 output pixels per CTA.  No im2col buffer exists: for every filter
//               tap the producer issues one 4-D TMA box load {32 ch, TW, TH, TN} from the NHWC tensor at
//               the tap's spatial offset; out-of-bounds rows/columns/channels are zero-filled by the
//               TMA unit, which implements the convolution padding and the K tail for free.  Stride-2
//               convolutions read four parity-plane views (separate tensor maps over the same buffer).
//   B operand = weights pre-packed as [Cout][tap][Cin rounded up to 32] (K-major, TF32-rounded), N = BN.
//   D         = 128 lanes x BN fp32 columns of TMEM.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected
// lane), warps 2-5 = epilogue (tcgen05.ld -> bias -> 16-byte stores, or red.global.add.v4.f32 when the
// K range is split across CTAs / the output accumulates several convolutions).
// Roofline: small-batch layers are bound by streaming the weights once from HBM (split-K spreads one
// layer's weight stream over all 148 SMs); large-batch layers by the tensor pipe.
#include "common.cuh"
#include "tcgen05.cuh"
#include "conv_common.cuh"
#include "../../include/sparsefusion_b200.h"
#include <mutex>
#include <unordered_map>
#include <vector>
#include <string.h>

namespace sfb {

// NP = 1: single-pass TF32 (operands are expected TF32-rounded).  NP = 3: error-compensated "3xTF32": the epilogue
// warps split every landed stage into hi (= what the tensor core keeps of the raw fp32 word) and lo = x - hi in shared
// memory, and the issuer runs hi*hi + lo*hi + hi*lo, which restores ~fp32 accuracy (rel. error ~2^-21 per product).
template <int BN, int NP>
struct ConvCfg {
    static constexpr int kBBytes = BN * 128;
    static constexpr int kRawBytes = kABytes + kBBytes;          // what TMA delivers per stage
    static constexpr int kStageBytes = kRawBytes * (NP == 3 ? 2 : 1);
    static constexpr int kStages = NP == 3 ? ((BN >= 256) ? 2 : (BN >= 128 ? 3 : (BN >= 64 ? 4 : 5)))
                                           : ((BN >= 256) ? 4 : (BN >= 128 ? 6 : 8));
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 512 /*barriers, bias*/ + BN * 4;
};

template <int BN, int NP>
__global__ void __launch_bounds__(kThreads, 1) conv_gemm_tf32_kernel(const __grid_constant__ ConvGemmParams p) {
    using Cfg = ConvCfg<BN, NP>;
    constexpr int S = Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by OFFSET from the __shared__ symbol, so the compiler keeps the shared address space (LDS/STS, not generic LD/ST)
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    // stage s: [A raw 16 KB | B raw BN*128 B | (NP == 3) A lo | B lo]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S * Cfg::kStageBytes);
    uint64_t* empty_bar = full_bar + S;
    uint64_t* split_bar = empty_bar + S;
    uint64_t* tmem_full_bar = split_bar + S;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
    float* bias_s = reinterpret_cast<float*>(smem + S * Cfg::kStageBytes + 512);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // tile coordinates
    const int tile = blockIdx.x;
    const int tw_i = tile % p.tiles_w, th_i = (tile / p.tiles_w) % p.tiles_h, tn_i = tile / (p.tiles_w * p.tiles_h);
    const int w0 = tw_i * p.TW, h0 = th_i * p.TH, n0 = tn_i * p.TN;
    const int cout0 = blockIdx.y * BN;
    const int kb = (int)(((int64_t)p.k_iters * blockIdx.z) / p.splits);
    const int ke = (int)(((int64_t)p.k_iters * (blockIdx.z + 1)) / p.splits);
    const int niter = ke - kb;

    if (threadIdx.x == 0) {
        for (int i = 0; i < S; ++i) { tc::mbar_init(&full_bar[i], 1); tc::mbar_init(&empty_bar[i], 1); tc::mbar_init(&split_bar[i], 4); }
        tc::mbar_init(tmem_full_bar, 1);
        tc::fence_mbar_init();
    }
    if (warp == 0 && lane == 0) {
        tc::prefetch_tensormap(&p.tmA[0]);
        tc::prefetch_tensormap(&p.tmB);
    }
    if (warp == 1) tc::tmem_alloc<(BN < 32 ? 32 : BN)>(tmem_holder);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (tc::elect_one()) {
            for (int it = 0; it < niter; ++it) {
                const int s = it % S;
                const uint32_t ph = (uint32_t)(it / S) & 1u;
                tc::mbar_wait(&empty_bar[s], ph ^ 1u);
                tc::mbar_expect_tx(&full_bar[s], Cfg::kRawBytes);
                const int i = kb + it;
                const int tap = i / p.cin_chunks, cc = i - tap * p.cin_chunks;
                const int ky = tap / p.KW, kx = tap - ky * p.KW;
                int oy = ky - p.pad, ox = kx - p.pad, map = 0;
                if (p.stride == 2) {
                    const int py = oy & 1, px = ox & 1;
                    map = py * 2 + px;
                    oy = (oy - py) >> 1;
                    ox = (ox - px) >> 1;
                }
                uint8_t* st = smem + s * Cfg::kStageBytes;
                tc::tma_load_4d(st, &p.tmA[map], &full_bar[s], cc * kBK, w0 + ox, h0 + oy, n0);
                tc::tma_load_2d(st + kABytes, &p.tmB, &full_bar[s], i * kBK, cout0);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (tc::elect_one()) {
            constexpr uint32_t idesc = tc::umma_idesc_tf32(kBM, BN);
            for (int it = 0; it < niter; ++it) {
                const int s = it % S;
                const uint32_t ph = (uint32_t)(it / S) & 1u;
                tc::mbar_wait(NP == 3 ? &split_bar[s] : &full_bar[s], ph);
                tc::fence_after_sync();
                const uint32_t a0 = tc::smem_u32(smem + s * Cfg::kStageBytes), b0 = a0 + kABytes;
#pragma unroll
                for (int k = 0; k < kBK / 8; ++k) {  // UMMA K = 8 tf32 = 32 bytes inside the 128-byte swizzle row
                    const uint64_t da = tc::umma_desc_k_sw128(a0 + k * 32), db = tc::umma_desc_k_sw128(b0 + k * 32);
                    tc::umma_tf32(tmem_base, da, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
                    if (NP == 3) {
                        const uint64_t dal = tc::umma_desc_k_sw128(a0 + Cfg::kRawBytes + k * 32);
                        const uint64_t dbl = tc::umma_desc_k_sw128(b0 + Cfg::kRawBytes + k * 32);
                        tc::umma_tf32(tmem_base, dal, db, idesc, 1u);
                        tc::umma_tf32(tmem_base, da, dbl, idesc, 1u);
                    }
                }
                tc::umma_commit(&empty_bar[s]);  // frees the smem slot when these MMAs have read it
            }
            tc::umma_commit(tmem_full_bar);
        }
    } else {
        // ------------------------------------------------------------ epilogue (warps 2..5)
        const int et = threadIdx.x - 64;  // 0..127
        for (int i = et; i < BN; i += 128) {
            const int c = cout0 + i;
            bias_s[i] = (p.bias != nullptr && blockIdx.z == 0 && c < p.Cout) ? p.bias[c] : 0.f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");

        if (NP == 3) {
            // hi/lo split of every landed stage (same byte offsets => the swizzled layout is preserved)
            constexpr int kVec = Cfg::kRawBytes / 16;
            for (int it = 0; it < niter; ++it) {
                const int s = it % S;
                const uint32_t ph = (uint32_t)(it / S) & 1u;
                tc::mbar_wait(&full_bar[s], ph);
                float4* raw = reinterpret_cast<float4*>(smem + s * Cfg::kStageBytes);
                float4* lo = reinterpret_cast<float4*>(smem + s * Cfg::kStageBytes + Cfg::kRawBytes);
#pragma unroll 4
                for (int i = et; i < kVec; i += 128) {
                    const float4 v = raw[i];
                    // hi is written back explicitly (exact TF32 bit patterns), so the result does not depend on how the
                    // tensor core would have reduced a full fp32 word (truncation vs rounding)
                    float4 h, l;
                    h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
                    h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
                    h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
                    h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
                    raw[i] = h;
                    lo[i] = l;
                }
                tc::fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&split_bar[s]);
            }
        }

        const int q = warp & 3;            // TMEM lane quarter this warp may access
        const int r = q * 32 + lane;       // row of the tile == output pixel
        const int tw = r % p.TW, th = (r / p.TW) % p.TH, tn = r / (p.TW * p.TH);
        const int n = n0 + tn, h = h0 + th, w = w0 + tw;
        const bool valid = (n < p.NB) && (h < p.Ho) && (w < p.Wo);
        const int64_t pix = (int64_t)((int64_t)n * p.Ho + h) * p.Wo + w;
        float* orow = p.out + pix * p.ldo + cout0;
        const float* rrow = (p.residual != nullptr && blockIdx.z == 0) ? p.residual + pix * p.ldr + cout0 : nullptr;
        const bool use_red = (p.splits > 1) || (p.accumulate != 0);

        tc::mbar_wait(tmem_full_bar, 0);
        tc::fence_after_sync();
#pragma unroll 1
        for (int j = 0; j < BN / 32; ++j) {
            uint32_t v[32];
            tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * 32), v);
            tc::tmem_ld_wait();
            if (valid) {
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const int c = j * 32 + g * 4;
                    if (cout0 + c < p.Cout) {
                        float x0 = __uint_as_float(v[g * 4 + 0]) + bias_s[c + 0];
                        float x1 = __uint_as_float(v[g * 4 + 1]) + bias_s[c + 1];
                        float x2 = __uint_as_float(v[g * 4 + 2]) + bias_s[c + 2];
                        float x3 = __uint_as_float(v[g * 4 + 3]) + bias_s[c + 3];
                        if (rrow) {
                            const float4 rv = __ldg(reinterpret_cast<const float4*>(rrow + c));
                            x0 += rv.x; x1 += rv.y; x2 += rv.z; x3 += rv.w;
                        }
                        if (use_red) {
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(orow + c), "f"(x0), "f"(x1), "f"(x2), "f"(x3)
                                         : "memory");
                        } else {
                            *reinterpret_cast<float4*>(orow + c) = make_float4(x0, x1, x2, x3);
                        }
                    }
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc::fence_after_sync();
        tc::tmem_dealloc<(BN < 32 ? 32 : BN)>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    });
    return fn;
}

static int g_tma_tf32_type = 0;
static int g_variant = 2;         // 3xTF32 kernel generation: 1 = all operands split in smem, 2 = M-side through TMEM (+ swap-AB)  // 0: FLOAT32 loads (hardware truncates to tf32); 1: TFLOAT32 tensor-map type

struct MapKey {
    uint64_t v[12];
    bool operator==(const MapKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        uint64_t h = 1469598103934665603ull;
        for (uint64_t x : k.v) { h ^= x; h *= 1099511628211ull; }
        return (size_t)h;
    }
};
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash>& map_cache() {
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> c;
    return c;
}
static std::mutex g_map_mutex;

// rank-4 fp32 tensor map: dims {d0..d3} (d0 innermost), byte strides for d1..d3, box {b0..b3}
static int make_map(CUtensorMap* out, const void* base, const uint64_t dims[4], const uint64_t strides[3], const uint32_t box[4], int rank) {
    MapKey key{};
    key.v[0] = reinterpret_cast<uint64_t>(base);
    for (int i = 0; i < 4; ++i) key.v[1 + i] = dims[i];
    for (int i = 0; i < 3; ++i) key.v[5 + i] = strides[i];
    key.v[8] = ((uint64_t)box[0] << 32) | box[1];
    key.v[9] = ((uint64_t)box[2] << 32) | box[3];
    key.v[10] = (uint64_t)rank;
    key.v[11] = (uint64_t)g_tma_tf32_type;
    std::lock_guard<std::mutex> lock(g_map_mutex);
    auto it = map_cache().find(key);
    if (it != map_cache().end()) { *out = it->second; return SFB_OK; }
    PFN_encodeTiled enc = get_encode();
    if (!enc) return fail(SFB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t gd[4] = {dims[0], dims[1], dims[2], dims[3]};
    cuuint64_t gs[3] = {strides[0], strides[1], strides[2]};
    cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(out, g_tma_tf32_type ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                     const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(SFB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rank=%d dims=[%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u] stride1=%llu", (int)r,
                    rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2], (unsigned long long)dims[3],
                    box[0], box[1], box[2], box[3], (unsigned long long)strides[0]);
    if (map_cache().size() > 65536) map_cache().clear();
    map_cache()[key] = *out;
    return SFB_OK;
}

// optional per-launch timing of the conv kernel (CUDA events on the launch stream) for bench.py's roofline line
static bool g_prof = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_events;
static double g_prof_bytes = 0.0, g_prof_flops = 0.0;

static cudaEvent_t g_prof_e0 = nullptr;
int conv_prof_begin(cudaStream_t st) {
    if (!g_prof) return SFB_OK;
    cudaEvent_t e1 = nullptr;
    SFB_CUDA(cudaEventCreate(&g_prof_e0));
    SFB_CUDA(cudaEventCreate(&e1));
    SFB_CUDA(cudaEventRecord(g_prof_e0, st));
    g_prof_events.emplace_back(g_prof_e0, e1);
    return SFB_OK;
}
int conv_prof_end(cudaStream_t st) {
    if (!g_prof) return SFB_OK;
    SFB_CUDA(cudaEventRecord(g_prof_events.back().second, st));
    return SFB_OK;
}

template <int BN, int NP>
static int launch_conv(const ConvGemmParams& p, dim3 grid, cudaStream_t st) {
    SFB_ONCE_PER_DEVICE(SFB_CUDA(cudaFuncSetAttribute(conv_gemm_tf32_kernel<BN, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg<BN, NP>::kSmemBytes)));
    conv_prof_begin(st);
    void (*kp)(const ConvGemmParams) = conv_gemm_tf32_kernel<BN, NP>;
    kp<<<grid, kThreads, ConvCfg<BN, NP>::kSmemBytes, st>>>(p);
    conv_prof_end(st);
    return check_launch("conv2d_nhwc_tf32");
}

static inline int pow2_floor(int v) {
    int r = 1;
    while (r * 2 <= v) r *= 2;
    return r;
}

}  // namespace sfb

using namespace sfb;

extern "C" {

int sfb_conv_set_tma_tf32(int enable) {
    g_tma_tf32_type = enable ? 1 : 0;
    return SFB_OK;
}

int sfb_conv_prof_enable(int on) {
    g_prof = on != 0;
    if (on) { g_prof_bytes = 0.0; g_prof_flops = 0.0; }
    return SFB_OK;
}

/* sums the recorded conv launches: kernel time (ms), launch count, algorithmic weight bytes and FLOPs (2*MAC); clears the record */
int sfb_conv_prof_collect(double* total_ms, int* launches, double* weight_bytes, double* flops) {
    double t = 0.0;
    for (auto& pr : g_prof_events) {
        SFB_CUDA(cudaEventSynchronize(pr.second));
        float ms = 0.f;
        SFB_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
        t += ms;
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    if (total_ms) *total_ms = t;
    if (launches) *launches = (int)g_prof_events.size();
    if (weight_bytes) *weight_bytes = g_prof_bytes;
    if (flops) *flops = g_prof_flops;
    g_prof_events.clear();
    g_prof_bytes = 0.0; g_prof_flops = 0.0;
    return SFB_OK;
}

int sfb_conv_set_variant(int v) {
    if (v < 1 || v > 3) return fail(SFB_ERR_ARG, "conv_set_variant: 1, 2 or 3 (3 = 2 with un-masked hi operands, experiment)");
    g_variant = v;
    return SFB_OK;
}

int sfb_conv_weight_k(int Cin, int KH, int KW) { return KH * KW * ((Cin + 31) / 32) * 32; }

int sfb_conv2d_nhwc_tf32(const float* x, int NB, int H, int W, int Cin, int64_t ldx, const float* w_packed, int Cout, int KH, int KW, int stride,
                         int pad, const float* bias, const float* residual, int64_t ldr, float* out, int64_t ldo, int accumulate, int splits, int bn,
                         void* stream) {
    return sfb_conv2d_nhwc_tf32_pad(x, NB, H, W, Cin, ldx, w_packed, Cout, KH, KW, stride, pad, pad, bias, residual, ldr, out, ldo, accumulate, splits, bn,
                                    stream);
}

int sfb_conv2d_nhwc_tf32_pad(const float* x, int NB, int H, int W, int Cin, int64_t ldx, const float* w_packed, int Cout, int KH, int KW, int stride,
                             int pad, int pad_after, const float* bias, const float* residual, int64_t ldr, float* out, int64_t ldo, int accumulate,
                             int splits, int bn, void* stream) {
    return sfb_conv2d_nhwc_tf32_ex(x, NB, H, W, Cin, ldx, w_packed, nullptr, Cout, KH, KW, stride, pad, pad_after, bias, residual, ldr, out, ldo, accumulate,
                                   splits, bn, stream);
}

static thread_local int t_w_dynamic = 0;
int sfb_conv2d_nhwc_tf32_dyn(const float* x, int NB, int H, int W, int Cin, int64_t ldx, const float* w, int Cout, int KH, int KW, int stride, int pad,
                             int pad_after, const float* bias, const float* residual, int64_t ldr, float* out, int64_t ldo, int accumulate, int splits,
                             int bn, void* stream) {
    t_w_dynamic = 1;
    const int rc = sfb_conv2d_nhwc_tf32_ex(x, NB, H, W, Cin, ldx, w, nullptr, Cout, KH, KW, stride, pad, pad_after, bias, residual, ldr, out, ldo, accumulate,
                                           splits, bn, stream);
    t_w_dynamic = 0;
    return rc;
}

int sfb_conv2d_nhwc_tf32_ex(const float* x, int NB, int H, int W, int Cin, int64_t ldx, const float* w_packed, const float* w_hi_lo, int Cout, int KH, int KW,
                            int stride, int pad, int pad_after, const float* bias, const float* residual, int64_t ldr, float* out, int64_t ldo,
                            int accumulate, int splits, int bn, void* stream) {
    SFB_REQUIRE(x && w_packed && out, "conv2d_nhwc_tf32: null pointer");
    SFB_REQUIRE(NB > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "conv2d_nhwc_tf32: empty shape");
    SFB_REQUIRE(stride == 1 || stride == 2, "conv2d_nhwc_tf32: stride must be 1 or 2");
    SFB_REQUIRE(Cin % 4 == 0 && ldx % 4 == 0 && ldx >= Cin, "conv2d_nhwc_tf32: Cin and ldx must be multiples of 4 floats (TMA 16-byte strides)");
    SFB_REQUIRE(Cout % 4 == 0 && ldo % 4 == 0, "conv2d_nhwc_tf32: Cout and ldo must be multiples of 4");
    SFB_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)w_packed & 15) == 0 && ((uintptr_t)out & 15) == 0, "conv2d_nhwc_tf32: pointers must be 16-byte aligned");
    SFB_REQUIRE(pad >= 0 && pad_after >= 0, "conv2d_nhwc_tf32: negative padding");
    // `pad` zeros before row/column 0, `pad_after` after the last one; both are TMA out-of-bounds zero fill, only the output extent differs
    const int Ho = (H + pad + pad_after - KH) / stride + 1, Wo = (W + pad + pad_after - KW) / stride + 1;
    SFB_REQUIRE(Ho > 0 && Wo > 0, "conv2d_nhwc_tf32: empty output");
    if (stride == 2) SFB_REQUIRE(H % 2 == 0 && W % 2 == 0, "conv2d_nhwc_tf32: stride 2 needs even H and W");

    ConvGemmParams p;
    memset(&p, 0, sizeof(p));
    // pixel tile.  Normal: TW x TH x TN = 128 output pixels on the M side.  Swap-AB (v2, <= 64 output pixels in total): the pixels
    // are the N side (16 / 32 / 64 of them) and 128 output channels the M side.
    const int planeW = (stride == 2) ? W / 2 : W, planeH = (stride == 2) ? H / 2 : H;
    const int64_t P_total = (int64_t)NB * Ho * Wo;
    const bool v2 = (precision_mode() == 1 && g_variant >= 2);
    const bool swap = v2 && P_total <= 64 && bn <= 0;
    int tile_pix = kBM;
    if (swap) { tile_pix = 16; while (tile_pix < P_total) tile_pix *= 2; }
    p.TW = Wo >= tile_pix ? tile_pix : pow2_floor(Wo);
    p.TH = tile_pix / p.TW;
    if (p.TH > Ho) p.TH = pow2_floor(Ho);
    p.TN = tile_pix / (p.TW * p.TH);
    p.tiles_w = ceil_div(Wo, p.TW);
    p.tiles_h = ceil_div(Ho, p.TH);
    const int tiles_n = ceil_div(NB, p.TN);
    const int tiles_m = p.tiles_w * p.tiles_h * tiles_n;

    const int cin_pad = ((Cin + 31) / 32) * 32;
    p.cin_chunks = cin_pad / 32;
    p.KH = KH; p.KW = KW; p.pad = pad; p.stride = stride;
    p.k_iters = KH * KW * p.cin_chunks;
    p.NB = NB; p.Ho = Ho; p.Wo = Wo; p.Cout = Cout;
    p.out = out; p.bias = bias; p.ldo = ldo; p.accumulate = accumulate;
    p.raw_hi = (g_variant == 3) ? 1 : 0;
    p.w_dynamic = t_w_dynamic;
    p.residual = residual; p.ldr = ldr;
    SFB_REQUIRE(residual == nullptr || (ldr % 4 == 0 && ((uintptr_t)residual & 15) == 0), "conv2d_nhwc_tf32: residual must be 16-byte aligned");

    // N tile
    int BN = bn;
    if (BN <= 0) {
        BN = Cout >= 256 ? 256 : (Cout >= 128 ? 128 : (Cout >= 64 ? 64 : 32));
        const int sms = sm_count();
        while (BN > 128 && tiles_m * ceil_div(Cout, BN) < sms) BN /= 2;  // prefer more CTAs when the grid is small
        if (v2 && BN > 128) BN = 128;
    }
    SFB_REQUIRE(BN == 32 || BN == 64 || BN == 128 || BN == 256, "conv2d_nhwc_tf32: bn must be 32, 64, 128 or 256");
    const int tiles_c = swap ? ceil_div(Cout, kBM) : ceil_div(Cout, BN);
    if (splits <= 0) {
        const int ctas = tiles_m * tiles_c;
        splits = 1;
        if (ctas < sm_count()) {
            splits = sm_count() / ctas;
            const int max_by_k = p.k_iters / 4 > 0 ? p.k_iters / 4 : 1;  // keep >= 4 k-steps per split
            if (splits > max_by_k) splits = max_by_k;
            if (splits < 1) splits = 1;
        }
    }
    if (precision_mode() == 1) {
        // the tensor core's fp32 accumulation is not exactly rounded: the error of one accumulator grows ~linearly with the number of
        // MMAs chained into it (measured 4.7e-5 rel. for 441 k-steps x 12 MMAs).  Bound the chain; partial sums meet in fp32 reds.
        const int min_splits = ceil_div(p.k_iters, 96);
        if (splits < min_splits) splits = min_splits;
    }
    if (splits > p.k_iters) splits = p.k_iters;
    p.splits = splits;

    // tensor maps: activations (1 map, or 4 parity planes for stride 2) and packed weights
    cudaStream_t st = as_stream(stream);
    const uint32_t boxA[4] = {32u, (uint32_t)p.TW, (uint32_t)p.TH, (uint32_t)p.TN};
    const int w_rows = swap ? kBM : BN;   // rows of the weight box
    if (stride == 1) {
        const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)NB};
        const uint64_t strides[3] = {(uint64_t)ldx * 4, (uint64_t)W * ldx * 4, (uint64_t)H * W * ldx * 4};
        if (int rc = make_map(&p.tmA[0], x, dims, strides, boxA, 4)) return rc;
        p.tmA[1] = p.tmA[2] = p.tmA[3] = p.tmA[0];
    } else {
        for (int py = 0; py < 2; ++py)
            for (int px = 0; px < 2; ++px) {
                const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)planeW, (uint64_t)planeH, (uint64_t)NB};
                const uint64_t strides[3] = {(uint64_t)2 * ldx * 4, (uint64_t)2 * W * ldx * 4, (uint64_t)H * W * ldx * 4};
                if (int rc = make_map(&p.tmA[py * 2 + px], x + ((int64_t)py * W + px) * ldx, dims, strides, boxA, 4)) return rc;
            }
    }
    {
        const uint64_t ktot = (uint64_t)KH * KW * cin_pad;
        const uint64_t dims[4] = {ktot, (uint64_t)Cout, 1, 1};
        const uint64_t strides[3] = {ktot * 4, 0, 0};
        const uint32_t boxB[4] = {32u, (uint32_t)w_rows, 1, 1};
        const bool presplit = (w_hi_lo != nullptr) && v2 && !swap && (swap || BN <= 128);
        p.presplit = presplit ? 1 : 0;
        if (presplit) {
            SFB_REQUIRE(((uintptr_t)w_hi_lo & 15) == 0, "conv2d_nhwc_tf32: w_hi_lo must be 16-byte aligned");
            if (int rc = make_map(&p.tmB, w_hi_lo, dims, strides, boxB, 2)) return rc;
            if (int rc = make_map(&p.tmBlo, w_hi_lo + (size_t)Cout * ktot, dims, strides, boxB, 2)) return rc;
        } else {
            if (int rc = make_map(&p.tmB, w_packed, dims, strides, boxB, 2)) return rc;
            p.tmBlo = p.tmB;
        }
    }

    if (splits > 1 && !accumulate) {
        // partial sums are reduced with red.global.add: the destination slice must start from zero
        SFB_CUDA(cudaMemset2DAsync(out, (size_t)ldo * 4, 0, (size_t)Cout * 4, (size_t)NB * Ho * Wo, st));
    }
    dim3 grid(tiles_m, tiles_c, splits);
    if (g_prof) {
        g_prof_bytes += (double)Cout * KH * KW * Cin * 4.0;
        g_prof_flops += 2.0 * NB * Ho * Wo * (double)Cout * KH * KW * Cin;
    }
    if (v2 && (swap || BN <= 128)) return launch_conv_v2(p, swap ? tile_pix : BN, swap, grid, st);
    if (precision_mode() == 1) {
        switch (BN) {
            case 32: return launch_conv<32, 3>(p, grid, st);
            case 64: return launch_conv<64, 3>(p, grid, st);
            case 128: return launch_conv<128, 3>(p, grid, st);
            default: return launch_conv<256, 3>(p, grid, st);
        }
    }
    switch (BN) {
        case 32: return launch_conv<32, 1>(p, grid, st);
        case 64: return launch_conv<64, 1>(p, grid, st);
        case 128: return launch_conv<128, 1>(p, grid, st);
        default: return launch_conv<256, 1>(p, grid, st);
    }
}
}
