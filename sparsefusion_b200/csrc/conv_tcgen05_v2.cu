// conv_tcgen05_v2.cu -- second-generation 3xTF32 implicit-GEMM convolution: the M-side operand goes through TENSOR MEMORY.
//
// Why: ncu on the first kernel of this round (conv_tcgen05.cu, "v1") shows the error-compensated path bound by SHARED-MEMORY bandwidth, not HBM: every
// stage was read by the splitter, written back twice (hi, lo) and then read six more times by the three MMAs per k-step
// (~190 KB of smem traffic per 32 KB of operands).  Here:
//   * the 128-row ("M-side") operand tile lands in smem by TMA once, is read ONCE by the converter warps (thread r owns row r,
//     swizzle-aware LDS.128), split into hi = tf32(x) and lo = x - hi in registers and written to TMEM with tcgen05.st; the MMAs
//     then take A from TMEM (tcgen05.mma ... [d], [a], b_desc) -- no further smem reads for that operand;
//   * only the narrow ("N-side") tile stays in smem (hi in place + lo copy);
//   * SWAP mode for layers with <= 64 output pixels per launch (the 4x4 / 8x8 stages at batch 1, which hold 1.39 of the UNet's
//     1.60 GB of weights): the WEIGHTS are the M-side operand (128 output channels per CTA, streamed HBM -> smem -> TMEM) and
//     the pixels the N-side (N = 16/32/64), so the tensor core does no work on padding rows and a stage is 20-32 KB of smem:
//     6-7 stages = ~100 KB of weights in flight per SM, enough to cover HBM latency at full bandwidth.
//   * the producer streams the first ring of WEIGHT tiles before griddepcontrol.wait (programmatic dependent launch) on their own arrival
//     barriers; in swap mode the converter turns them into TMEM operands while the previous kernel is still running;
//   * launches with more than 64 output pixels (tensor-bound) can take pre-split (hi, lo) weights by TMA (sfb_conv2d_nhwc_tf32_ex), so
//     the converter only touches the activation tile;
//   * the swap-mode epilogue transposes the accumulator through the idle stages so that global traffic is 16-byte vectors along channels.
// D = sum over (tap, channel chunk) of  A_hi*B_hi + A_lo*B_hi + A_hi*B_lo   (fp32 accumulation in TMEM), as in v1.
// Tiling, tap -> TMA coordinate mapping, split-K, bias / residual / accumulate epilogue semantics are v1's (conv_tcgen05.cu).
#include "common.cuh"
#include "tcgen05.cuh"
#include "conv_common.cuh"
#include "../../include/sparsefusion_b200.h"
#include <string.h>

namespace sfb {

template <int BN>
struct Conv2Cfg {
    static constexpr int kNBytes = BN * 128;                 // N-side tile (raw == hi after the split)
    static constexpr int kStageBytes = kABytes + 2 * kNBytes;  // M raw | N hi | N lo
    static constexpr int kDCols = BN < 32 ? 32 : BN;          // accumulator columns
    static constexpr int kABase = kDCols < 64 ? 64 : kDCols;  // first TMEM column of the A-operand ring
    static constexpr int kStagesSmem = (200 * 1024) / kStageBytes;
    static constexpr int kStagesTmem = (512 - kABase) / 64;   // 32 hi + 32 lo columns per stage
    static constexpr int kStages = kStagesSmem < kStagesTmem ? kStagesSmem : kStagesTmem;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 1024 + 512 + 512;   // stages | barriers 512 | bias 1024 | pixel index 512 | alignment slack
};

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
        "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]),
        "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// A from tensor memory, B from a shared-memory descriptor
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Phase tracer (sfb_conv_phase_trace): CTA (0,0,0) of every launch stamps %globaltimer at 7 points into buf[1 + 8*launch + phase]:
//   0 kernel entry, 1 prologue done (barriers, TMEM alloc), 2 griddepcontrol.wait returned, 3 first stage converted (MMA can start),
//   4 last MMA committed, 5 accumulator complete (epilogue starts), 6 epilogue done.
static __device__ unsigned long long* t_phase_buf = nullptr;
static __device__ unsigned int t_phase_cap = 0;
__device__ __forceinline__ void phase_mark(unsigned int slot, int phase) {
    if (slot != 0xffffffffu) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        t_phase_buf[1 + (size_t)slot * 8 + phase] = t;
    }
}

// SWAP == false: M-side = 128 output pixels (tmA), N-side = BN output channels (tmB).
// SWAP == true : M-side = 128 output channels (tmB), N-side = BN output pixels (tmA, box {32, TW, TH, TN} with TW*TH*TN == BN).
template <int BN, bool SWAP>
__global__ void __launch_bounds__(kThreads, 1) conv_gemm_v2_kernel(const __grid_constant__ ConvGemmParams p) {
    using Cfg = Conv2Cfg<BN>;
    constexpr int S = Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by OFFSET from the __shared__ symbol, so the compiler keeps the shared address space (LDS/STS, not generic LD/ST)
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    // two arrival barriers per stage -- weights and activations -- so that the converter can process the weight tiles of the first stages while
    // the activations do not exist yet (they are loaded after griddepcontrol.wait; the weights before)
    uint64_t* wfull_bar = reinterpret_cast<uint64_t*>(smem + S * Cfg::kStageBytes);
    uint64_t* afull_bar = wfull_bar + S;
    uint64_t* empty_bar = afull_bar + S;
    uint64_t* conv_bar = empty_bar + S;
    uint64_t* tmem_full_bar = conv_bar + S;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
    float* bias_s = reinterpret_cast<float*>(smem + S * Cfg::kStageBytes + 512);   // up to 256 floats
    int* pix_s = reinterpret_cast<int*>(smem + S * Cfg::kStageBytes + 512 + 1024);   // SWAP: linear output pixel of tile column j, or -1

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x;
    const int tw_i = tile % p.tiles_w, th_i = (tile / p.tiles_w) % p.tiles_h, tn_i = tile / (p.tiles_w * p.tiles_h);
    const int w0 = tw_i * p.TW, h0 = th_i * p.TH, n0 = tn_i * p.TN;
    const int cout0 = blockIdx.y * (SWAP ? kBM : BN);
    const int kb = (int)(((int64_t)p.k_iters * blockIdx.z) / p.splits);
    const int ke = (int)(((int64_t)p.k_iters * (blockIdx.z + 1)) / p.splits);
    const int niter = ke - kb;

    uint32_t* phase_slot = tmem_holder + 1;
    if (threadIdx.x == 0) {
        unsigned int slot = 0xffffffffu;
        if (t_phase_buf != nullptr && (blockIdx.x | blockIdx.y | blockIdx.z) == 0) {
            const unsigned long long sidx = atomicAdd(t_phase_buf, 1ull);
            if (sidx < t_phase_cap) slot = (unsigned int)sidx;
        }
        *phase_slot = slot;
        phase_mark(slot, 0);
        for (int i = 0; i < S; ++i) {
            tc::mbar_init(&wfull_bar[i], 1); tc::mbar_init(&afull_bar[i], 1); tc::mbar_init(&empty_bar[i], 1); tc::mbar_init(&conv_bar[i], 4);
        }
        tc::mbar_init(tmem_full_bar, 1);
        tc::fence_mbar_init();
    }
    if (warp == 0 && lane == 0) {
        tc::prefetch_tensormap(&p.tmA[0]);
        tc::prefetch_tensormap(&p.tmB);
    }
    if (warp == 1) tc::tmem_alloc<512>(tmem_holder);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_holder;
    const uint32_t pslot = *phase_slot;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (tc::elect_one()) {
            const bool presplit = !SWAP && p.presplit != 0;
            const uint32_t w_tx = SWAP ? (uint32_t)kABytes : (uint32_t)(Cfg::kNBytes * (presplit ? 2 : 1));
            const uint32_t a_tx = SWAP ? (uint32_t)Cfg::kNBytes : (uint32_t)kABytes;
            auto load_weights = [&](int it) {
                const int s = it % S;
                uint8_t* st = smem + s * Cfg::kStageBytes;
                tc::tma_load_2d(SWAP ? st : st + kABytes, &p.tmB, &wfull_bar[s], (kb + it) * kBK, cout0);
                if (presplit) tc::tma_load_2d(st + kABytes + Cfg::kNBytes, &p.tmBlo, &wfull_bar[s], (kb + it) * kBK, cout0);
            };
            auto load_acts = [&](int it) {
                const int s = it % S;
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
                tc::tma_load_4d(SWAP ? st + kABytes : st, &p.tmA[map], &afull_bar[s], cc * kBK, w0 + ox, h0 + oy, n0);
            };
            // PDL: the weights do not depend on the previous kernel -- the first S stages of weight tiles stream in while it drains;
            // the activation tiles of those stages follow once griddepcontrol.wait returns.
            const int pre = niter < S ? niter : S;
            phase_mark(pslot, 1);
            pdl_trigger();
            // ... unless the "weights" are themselves the output of an earlier kernel (q k^T, P v of the VAE's attention): prologues of a PDL chain run
            // arbitrarily far ahead of the bodies, so such an operand may not exist yet -- wait first (found as intermittent NaN in the decoded image)
            if (p.w_dynamic) pdl_wait();
            for (int it = 0; it < pre; ++it) {
                tc::mbar_expect_tx(&wfull_bar[it], w_tx);
                load_weights(it);
            }
            pdl_wait();
            phase_mark(pslot, 2);
            trace_mark();
            for (int it = 0; it < pre; ++it) {
                tc::mbar_expect_tx(&afull_bar[it], a_tx);
                load_acts(it);
            }
            for (int it = pre; it < niter; ++it) {
                const int s = it % S;
                const uint32_t ph = (uint32_t)(it / S) & 1u;
                tc::mbar_wait(&empty_bar[s], ph ^ 1u);
                tc::mbar_expect_tx(&afull_bar[s], a_tx);
                tc::mbar_expect_tx(&wfull_bar[s], w_tx);
                load_acts(it);
                load_weights(it);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (tc::elect_one()) {
            constexpr uint32_t idesc = tc::umma_idesc_tf32(kBM, BN);
            for (int it = 0; it < niter; ++it) {
                const int s = it % S;
                const uint32_t ph = (uint32_t)(it / S) & 1u;
                tc::mbar_wait(&conv_bar[s], ph);
                tc::fence_after_sync();
                if (it == 0) phase_mark(pslot, 3);
                const uint32_t b_hi = tc::smem_u32(smem + s * Cfg::kStageBytes + kABytes), b_lo = b_hi + Cfg::kNBytes;
                const uint32_t a_hi = tmem_base + (uint32_t)(Cfg::kABase + 64 * s), a_lo = a_hi + 32;
#pragma unroll
                for (int k = 0; k < kBK / 8; ++k) {
                    const uint64_t dbh = tc::umma_desc_k_sw128(b_hi + k * 32), dbl = tc::umma_desc_k_sw128(b_lo + k * 32);
                    umma_tf32_ts(tmem_base, a_hi + k * 8, dbh, idesc, (it > 0 || k > 0) ? 1u : 0u);
                    umma_tf32_ts(tmem_base, a_lo + k * 8, dbh, idesc, 1u);
                    umma_tf32_ts(tmem_base, a_hi + k * 8, dbl, idesc, 1u);
                }
                tc::umma_commit(&empty_bar[s]);
            }
            tc::umma_commit(tmem_full_bar);
            phase_mark(pslot, 4);
        }
    } else {
        // ------------------------------------------------------------ converter + epilogue (warps 2..5)
        const int et = threadIdx.x - 64;   // 0..127
        const int q = warp & 3;            // TMEM lane quarter of this warp
        const int r = q * 32 + lane;       // M-side row owned by this thread
        const int nb = SWAP ? kBM : BN;
        for (int i = et; i < nb; i += 128) {
            const int c = cout0 + i;
            bias_s[i] = (p.bias != nullptr && blockIdx.z == 0 && c < p.Cout) ? p.bias[c] : 0.f;
        }
        if (SWAP && et < BN) {
            const int tw = et % p.TW, th = (et / p.TW) % p.TH, tn = et / (p.TW * p.TH);
            const int n = n0 + tn, h = h0 + th, w = w0 + tw;
            pix_s[et] = (n < p.NB && h < p.Ho && w < p.Wo) ? (n * p.Ho + h) * p.Wo + w : -1;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");

        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        uint64_t* mside_bar = SWAP ? wfull_bar : afull_bar;   // arrival of the 128-row tile (weights in swap mode)
        uint64_t* nside_bar = SWAP ? afull_bar : wfull_bar;
        // (1) M-side row r of stage s: 8 swizzled 16-byte chunks -> hi / lo -> TMEM columns [kABase + 64 s, +32) / [+32, +64)
        auto convert_m = [&](int it) {
            const int s = it % S;
            tc::mbar_wait(&mside_bar[s], (uint32_t)(it / S) & 1u);
            const uint8_t* rowp = smem + s * Cfg::kStageBytes + r * 128;
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 v = *reinterpret_cast<const float4*>(rowp + ((c ^ (r & 7)) << 4));
                const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint32_t h = __float_as_uint(f[e]) & 0xFFFFE000u;
                    hi[c * 4 + e] = p.raw_hi ? __float_as_uint(f[e]) : h;
                    lo[c * 4 + e] = __float_as_uint(f[e] - __uint_as_float(h));
                }
            }
            const uint32_t ta = lane_addr + (uint32_t)(Cfg::kABase + 64 * s);
            tmem_st_32x32(ta, hi);
            tmem_st_32x32(ta + 32, lo);
        };
        // (2) N-side tile of stage s: hi in place, lo to the second buffer (same offsets => same swizzled layout); nothing to convert when the
        //     weights arrive pre-split (hi and lo tiles are TMA-loaded straight into the two buffers).  Then hand the stage to the MMA issuer.
        auto convert_n_and_release = [&](int it) {
            const int s = it % S;
            tc::mbar_wait(&nside_bar[s], (uint32_t)(it / S) & 1u);
            if (SWAP || !p.presplit) {
                uint8_t* st = smem + s * Cfg::kStageBytes;
                float4* nraw = reinterpret_cast<float4*>(st + kABytes);
                float4* nlo = reinterpret_cast<float4*>(st + kABytes + Cfg::kNBytes);
#pragma unroll 2
                for (int i = et; i < BN * 8; i += 128) {
                    const float4 v = nraw[i];
                    float4 h, l;
                    h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
                    h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
                    h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
                    h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
                    if (!p.raw_hi) nraw[i] = h;
                    nlo[i] = l;
                }
            }
            tmem_st_wait();
            tc::fence_proxy_async();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&conv_bar[s]);
        };
        // Swap-AB: the 128-row tiles are WEIGHTS, which the producer streams in before the dependency wait -- convert all of the first ring's
        // tiles into tensor memory up front, so that once the (small) activation tiles land only their split and the MMAs remain.
        const int pre = SWAP ? (niter < S ? niter : S) : 0;
        for (int it = 0; it < pre; ++it) convert_m(it);
        for (int it = 0; it < pre; ++it) convert_n_and_release(it);
        for (int it = pre; it < niter; ++it) {
            convert_m(it);
            convert_n_and_release(it);
        }

        // ---- epilogue
        const bool use_red = (p.splits > 1) || (p.accumulate != 0);
        pdl_wait();   // residual reads / output writes below must see the previous kernel's results (returns at once: the activations are in)
        tc::mbar_wait(tmem_full_bar, 0);
        tc::fence_after_sync();
        if (et == 0) phase_mark(pslot, 5);
        if (!SWAP) {
            const int tw = r % p.TW, th = (r / p.TW) % p.TH, tn = r / (p.TW * p.TH);
            const int n = n0 + tn, h = h0 + th, w = w0 + tw;
            const bool valid = (n < p.NB) && (h < p.Ho) && (w < p.Wo);
            const int64_t pix = (int64_t)((int64_t)n * p.Ho + h) * p.Wo + w;
            float* orow = p.out + pix * p.ldo + cout0;
            const float* rrow = (p.residual != nullptr && blockIdx.z == 0) ? p.residual + pix * p.ldr + cout0 : nullptr;
#pragma unroll 1
            for (int j = 0; j < Cfg::kDCols / 32; ++j) {
                uint32_t v[32];
                tc::tmem_ld_32x32(lane_addr + (uint32_t)(j * 32), v);
                tc::tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        const int c = j * 32 + g * 4;
                        if (c < BN && cout0 + c < p.Cout) {
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
        } else {
            // rows = output channels (this thread: cout0 + r), columns = the BN pixels of the tile.  The accumulator is transposed through
            // shared memory (the operand stages are idle by now) so that global traffic is 16-byte vectors along the channel dimension:
            // one warp = the 128 channels of one pixel = 512 contiguous bytes per red.v4 / st.v4 instruction (scalar reds cost 3.7 us for 64 pixels).
            constexpr int kLd = 132;                                    // floats per staged pixel row (16-byte aligned, bank-conflict free)
            float* stg = reinterpret_cast<float*>(smem);
#pragma unroll 1
            for (int j = 0; j < Cfg::kDCols / 32; ++j) {
                uint32_t v[32];
                tc::tmem_ld_32x32(lane_addr + (uint32_t)(j * 32), v);
                tc::tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 32; ++e)
                    if (j * 32 + e < BN) stg[(j * 32 + e) * kLd + r] = __uint_as_float(v[e]);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const bool add_res = (p.residual != nullptr && blockIdx.z == 0);
            for (int idx = et; idx < BN * 32; idx += 128) {
                const int pj = idx >> 5, c4 = idx & 31;
                const int pix = pix_s[pj];
                const int c = cout0 + c4 * 4;
                if (pix >= 0 && c < p.Cout) {
                    float4 a = *reinterpret_cast<const float4*>(stg + pj * kLd + c4 * 4);
                    const float4 bv = *reinterpret_cast<const float4*>(bias_s + c4 * 4);
                    a.x += bv.x; a.y += bv.y; a.z += bv.z; a.w += bv.w;
                    if (add_res) {
                        const float4 rv = __ldg(reinterpret_cast<const float4*>(p.residual + (int64_t)pix * p.ldr + c));
                        a.x += rv.x; a.y += rv.y; a.z += rv.z; a.w += rv.w;
                    }
                    float* dst = p.out + (int64_t)pix * p.ldo + c;
                    if (use_red) {
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w) : "memory");
                    } else {
                        *reinterpret_cast<float4*>(dst) = a;
                    }
                }
            }
        }
    }
    if (threadIdx.x == 64) phase_mark(pslot, 6);
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc::fence_after_sync();
        tc::tmem_dealloc<512>(tmem_base);
    }
}

template <int BN, bool SWAP>
static int launch_v2(const ConvGemmParams& p, dim3 grid, cudaStream_t st) {
    SFB_ONCE_PER_DEVICE(SFB_CUDA(cudaFuncSetAttribute(conv_gemm_v2_kernel<BN, SWAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv2Cfg<BN>::kSmemBytes)));
    conv_prof_begin(st);
    trace_name(SWAP ? (BN == 16 ? "conv_v2<16,swap>" : BN == 32 ? "conv_v2<32,swap>" : "conv_v2<64,swap>")
                    : (BN == 16 ? "conv_v2<16>" : BN == 32 ? "conv_v2<32>" : BN == 64 ? "conv_v2<64>" : BN == 128 ? "conv_v2<128>" : "conv_v2<256>"));
    launch_pdl(conv_gemm_v2_kernel<BN, SWAP>, grid, dim3(kThreads), Conv2Cfg<BN>::kSmemBytes, st, p);
    conv_prof_end(st);
    return check_launch("conv2d_nhwc_tf32(v2)");
}

void trace_bind_conv_v2(unsigned long long* buf, unsigned int cap) { trace_bind_this_tu(buf, cap); }
void phase_bind_conv_v2(unsigned long long* buf, unsigned int cap) {
    cudaMemcpyToSymbol(t_phase_buf, &buf, sizeof(buf));
    cudaMemcpyToSymbol(t_phase_cap, &cap, sizeof(cap));
}

int launch_conv_v2(const ConvGemmParams& p, int BN, bool swap, dim3 grid, cudaStream_t st) {
    if (swap) {
        switch (BN) {
            case 16: return launch_v2<16, true>(p, grid, st);
            case 32: return launch_v2<32, true>(p, grid, st);
            default: return launch_v2<64, true>(p, grid, st);
        }
    }
    switch (BN) {
        case 32: return launch_v2<32, false>(p, grid, st);
        case 64: return launch_v2<64, false>(p, grid, st);
        default: return launch_v2<128, false>(p, grid, st);
    }
}

}  // namespace sfb
