// unet_ops.cu -- the non-GEMM operators of the VLDM UNet forward, NHWC fp32, for sm_100a.
//
// They replace the eager PyTorch elementwise / normalisation / tiny-attention launches of the
// reference's Unet (external/imagen_pytorch.py): GroupNorm+FiLM+SiLU of `Block` (:654-662), gain-only
// LayerNorm / ChanLayerNorm (:301-329), GELU of ChanFeedForward (:953-961), the multi-query Attention
// (:511-566) and CrossAttention (:764-805) softmax cores (at most 1+2+H*W keys), GlobalContext pooling
// (:936-941), PixelShuffle(2)+SiLU (:588-592), skip concatenation (:1639), the learned sinusoidal time
// embedding (:634-639) and small-M Linear layers (time MLPs, token projections), which are weight
// streaming GEMVs and stay in full fp32.
// Tensors that feed a tcgen05 GEMM are rounded to TF32 (cvt.rna) when they are written here, so the
// tensor core's operand truncation never biases the result.
#include <cooperative_groups.h>
#include "common.cuh"
#include "tcgen05.cuh"
#include "../../include/sparsefusion_b200.h"

namespace sfb {

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + __expf(-x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ------------------------------------------------------------------------------------ layout
// NCHW [NB,C,H,W] -> channel slice [c_off, c_off+C) of NHWC [NB,H,W,ld]; 32x32 smem transpose per (n)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int HW, int64_t ld, int c_off, int round) {
    pdl_sync();
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, p = p0 + tx;
        tile[i][tx] = (c < C && p < HW) ? src[((int64_t)n * C + c) * HW + p] : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int p = p0 + i, c = c0 + tx;
        if (c < C && p < HW) {
            float v = tile[tx][i];
            if (round) v = tc::round_tf32(v);
            dst[((int64_t)n * HW + p) * ld + c_off + c] = v;
        }
    }
}
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int HW, int64_t ld) {
    pdl_sync();
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int i = ty; i < 32; i += 8) {
        const int p = p0 + i, c = c0 + tx;
        tile[i][tx] = (c < C && p < HW) ? src[((int64_t)n * HW + p) * ld + c] : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, p = p0 + tx;
        if (c < C && p < HW) dst[((int64_t)n * C + c) * HW + p] = tile[tx][i];
    }
}

// out[pix][0:C1] = a[pix][0:C1]; out[pix][C1:C1+C2] = b[pix][0:C2] * scale_b   (float4 granularity)
__global__ void concat2_kernel(const float4* __restrict__ a, int C1v, int64_t lda_v, const float4* __restrict__ b, int C2v, int64_t ldb_v,
                               float scale_b, float4* __restrict__ out, int64_t ldo_v, int64_t npix) {
    pdl_sync();
    const int Cv = C1v + C2v;
    const int64_t total = npix * Cv;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pix = i / Cv;
        const int c = (int)(i - pix * Cv);
        float4 v;
        if (c < C1v) v = __ldg(a + pix * lda_v + c);
        else {
            v = __ldg(b + pix * ldb_v + (c - C1v));
            v.x *= scale_b; v.y *= scale_b; v.z *= scale_b; v.w *= scale_b;
        }
        out[pix * ldo_v + c] = v;
    }
}

// PixelShuffle(2) of silu(y): y [NB,H,W,4*Co] -> out [NB,2H,2W, channel slice of width Co]; y channel = c*4 + i*2 + j
__global__ void pixel_shuffle_silu_kernel(const float* __restrict__ y, float* __restrict__ out, int H, int W, int Co, int64_t ldo, int64_t total) {
    pdl_sync();
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        // idx enumerates output elements: ((n*2H + oh)*2W + ow)*Co + c
        const int c = (int)(idx % Co);
        int64_t t = idx / Co;
        const int ow = (int)(t % (2 * W)); t /= (2 * W);
        const int oh = (int)(t % (2 * H));
        const int64_t n = t / (2 * H);
        const int h = oh >> 1, i = oh & 1, w = ow >> 1, j = ow & 1;
        const float v = __ldg(y + ((n * H + h) * W + w) * (int64_t)(4 * Co) + c * 4 + i * 2 + j);
        out[((n * 2 * H + oh) * (int64_t)(2 * W) + ow) * ldo + c] = silu_f(v);
    }
}

// ------------------------------------------------------------------------------------ GroupNorm
// CTAs (group, n, pixel-slab): partial fp64 (sum, sumsq) per slab; the last slab to finish folds the partials into (mean, rstd).
// `counters` is a persistent zero-initialised buffer (one uint per (n, group)); the finishing CTA resets its counter.
__global__ void __launch_bounds__(256) gn_stats_kernel(const float* __restrict__ x, int64_t ldx, int HW, int C, int G, float eps, int S,
                                                      double2* __restrict__ partial, unsigned int* __restrict__ counters,
                                                      float2* __restrict__ stats) {
    pdl_sync();
    const int g = blockIdx.x, n = blockIdx.y, sl = blockIdx.z;
    const int Cg = C / G;
    const int Cg4 = Cg >> 2;
    const int p0 = (int)(((int64_t)HW * sl) / S), p1 = (int)(((int64_t)HW * (sl + 1)) / S);
    const float* base = x + ((int64_t)n * HW + p0) * ldx + g * Cg;
    double s = 0.0, ss = 0.0;
    const int64_t total4 = (int64_t)(p1 - p0) * Cg4;
    for (int64_t i = threadIdx.x; i < total4; i += blockDim.x) {
        const int64_t pix = i / Cg4;
        const int c4 = (int)(i - pix * Cg4);
        const float4 v = __ldg(reinterpret_cast<const float4*>(base + pix * ldx) + c4);
        const float a = v.x + v.y + v.z + v.w;
        const float b = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        s += (double)a;
        ss += (double)b;
    }
    __shared__ double sh_s[8], sh_ss[8];
    __shared__ bool is_last;
    s = warp_sum_d(s);
    ss = warp_sum_d(ss);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sh_s[warp] = s; sh_ss[warp] = ss; }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = blockDim.x >> 5;
        s = 0.0; ss = 0.0;
        for (int w = 0; w < nw; ++w) { s += sh_s[w]; ss += sh_ss[w]; }
        const int slot = n * G + g;
        partial[(int64_t)slot * S + sl] = make_double2(s, ss);
        __threadfence();
        const unsigned int done = atomicAdd(&counters[slot], 1u);
        is_last = (done == (unsigned int)(S - 1));
        if (is_last) {
            __threadfence();
            double ts = 0.0, tss = 0.0;
            for (int k = 0; k < S; ++k) {
                const volatile double* pr = reinterpret_cast<volatile double*>(&partial[(int64_t)slot * S + k]);
                ts += pr[0]; tss += pr[1];
            }
            const double cnt = (double)HW * Cg;
            const double mean = ts / cnt;
            double var = tss / cnt - mean * mean;
            if (var < 0) var = 0;
            stats[slot] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
            counters[slot] = 0u;
        }
    }
}

// y = silu( ((x-mean)*rstd*gamma + beta) * (scale+1) + shift ), rounded to tf32.  film [NB, 2C] (scale | shift) or null
__global__ void gn_apply_kernel(const float* __restrict__ x, int64_t ldx, const float2* __restrict__ stats, const float* __restrict__ gamma,
                                const float* __restrict__ beta, const float* __restrict__ film, int64_t ldf, float* __restrict__ y, int64_t ldy, int HW, int C,
                                int G, int act, int round, int64_t total4) {
    pdl_sync();
    const int C4 = C >> 2, Cg = C / G;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pix = i / C4;
        const int c = (int)(i - pix * C4) * 4;
        const int n = (int)(pix / HW);
        const float2 st = __ldg(stats + n * G + c / Cg);
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + pix * ldx + c));
        const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c));
        const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c));
        float o[4] = {(v.x - st.x) * st.y * ga.x + be.x, (v.y - st.x) * st.y * ga.y + be.y, (v.z - st.x) * st.y * ga.z + be.z,
                      (v.w - st.x) * st.y * ga.w + be.w};
        if (film) {
            const float4 sc = __ldg(reinterpret_cast<const float4*>(film + (int64_t)n * ldf + c));
            const float4 sh = __ldg(reinterpret_cast<const float4*>(film + (int64_t)n * ldf + C + c));
            o[0] = o[0] * (sc.x + 1.f) + sh.x; o[1] = o[1] * (sc.y + 1.f) + sh.y;
            o[2] = o[2] * (sc.z + 1.f) + sh.z; o[3] = o[3] * (sc.w + 1.f) + sh.w;
        }
        if (act) {
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = silu_f(o[k]);
        }
        if (round) {
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = tc::round_tf32(o[k]);
        }
        *reinterpret_cast<float4*>(y + pix * ldy + c) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// Single-launch GroupNorm for the sizes of a B<=8 UNet evaluation: one thread-block CLUSTER per (n, group), CTA r of the cluster owns a
// slab of pixels, keeps its <= 8 float4 per thread in registers, publishes fp64 (sum, sumsq) in shared memory, reads the other CTAs'
// partials through distributed shared memory and applies affine / FiLM / SiLU straight from the registers -- x is read once.
constexpr int kGnFusedItems = 8;   // float4 per thread
__global__ void __launch_bounds__(256) gn_fused_kernel(const float* __restrict__ x, int64_t ldx, int HW, int C, int G, float eps,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       const float* __restrict__ film, int64_t ldf, float* __restrict__ y, int64_t ldy, int act,
                                                       int round) {
    pdl_sync();
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int R = (int)gridDim.x, sl = (int)blockIdx.x, g = blockIdx.y, n = blockIdx.z;
    const int Cg = C / G, Cg4 = Cg >> 2;
    const int p0 = (int)(((int64_t)HW * sl) / R), p1 = (int)(((int64_t)HW * (sl + 1)) / R);
    const int total4 = (p1 - p0) * Cg4;
    const float* base = x + ((int64_t)n * HW + p0) * ldx + g * Cg;
    float4 v[kGnFusedItems];
    double s = 0.0, ss = 0.0;
#pragma unroll
    for (int k = 0; k < kGnFusedItems; ++k) {
        const int i = threadIdx.x + k * 256;
        if (i < total4) {
            const int pix = i / Cg4, c4 = i - pix * Cg4;
            v[k] = __ldg(reinterpret_cast<const float4*>(base + (int64_t)pix * ldx) + c4);
            s += (double)(v[k].x + v[k].y + v[k].z + v[k].w);
            ss += (double)(v[k].x * v[k].x + v[k].y * v[k].y + v[k].z * v[k].z + v[k].w * v[k].w);
        }
    }
    __shared__ double sh_s[8], sh_ss[8];
    __shared__ double part[2];
    __shared__ float2 st_sh;
    s = warp_sum_d(s);
    ss = warp_sum_d(ss);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sh_s[warp] = s; sh_ss[warp] = ss; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < 8; ++w) { a += sh_s[w]; b += sh_ss[w]; }
        part[0] = a; part[1] = b;
    }
    cluster.sync();
    if (threadIdx.x == 0) {
        double ts = 0.0, tss = 0.0;
        for (int r = 0; r < R; ++r) {
            const double* rp = cluster.map_shared_rank(part, r);
            ts += rp[0]; tss += rp[1];
        }
        const double cnt = (double)HW * Cg;
        const double mean = ts / cnt;
        double var = tss / cnt - mean * mean;
        if (var < 0) var = 0;
        st_sh = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
    }
    cluster.sync();   // also keeps every CTA's `part` alive until all remote reads are done
    const float2 st = st_sh;
    float* ybase = y + ((int64_t)n * HW + p0) * ldy + g * Cg;
#pragma unroll
    for (int k = 0; k < kGnFusedItems; ++k) {
        const int i = threadIdx.x + k * 256;
        if (i < total4) {
            const int pix = i / Cg4, c4 = i - pix * Cg4;
            const int c = g * Cg + c4 * 4;
            const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c));
            const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c));
            float o[4] = {(v[k].x - st.x) * st.y * ga.x + be.x, (v[k].y - st.x) * st.y * ga.y + be.y, (v[k].z - st.x) * st.y * ga.z + be.z,
                          (v[k].w - st.x) * st.y * ga.w + be.w};
            if (film) {
                const float4 sc = __ldg(reinterpret_cast<const float4*>(film + (int64_t)n * ldf + c));
                const float4 sh = __ldg(reinterpret_cast<const float4*>(film + (int64_t)n * ldf + C + c));
                o[0] = o[0] * (sc.x + 1.f) + sh.x; o[1] = o[1] * (sc.y + 1.f) + sh.y;
                o[2] = o[2] * (sc.z + 1.f) + sh.z; o[3] = o[3] * (sc.w + 1.f) + sh.w;
            }
            if (act) {
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = silu_f(o[e]);
            }
            if (round) {
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = tc::round_tf32(o[e]);
            }
            *(reinterpret_cast<float4*>(ybase + (int64_t)pix * ldy) + c4) = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

// ------------------------------------------------------------------------------------ LayerNorm over the last dim
// rows [T, C] (row stride ldx); y = LN(pre(x)) * g (+ b) (+ res); pre: 0 none, 1 GELU.  One 128-thread CTA per row.  eps 1e-5, biased variance.
__device__ __forceinline__ float block_sum_128(float v, float* sh) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    const float t = sh[0] + sh[1] + sh[2] + sh[3];
    __syncthreads();
    return t;
}
__global__ void __launch_bounds__(128) layernorm_rows_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ g,
                                                            const float* __restrict__ b, const float* __restrict__ res, int64_t ldr,
                                                            float* __restrict__ y, int64_t ldy, int T, int C, int pre, int round) {
    pdl_sync();
    __shared__ float sh[4];
    const int row = blockIdx.x;
    const float* xr = x + (int64_t)row * ldx;
    float s = 0.f;
    for (int c = threadIdx.x; c < C; c += 128) {
        float v = xr[c];
        if (pre == 1) v = gelu_f(v);
        s += v;
    }
    const float mean = block_sum_128(s, sh) / C;
    float ss = 0.f;
    for (int c = threadIdx.x; c < C; c += 128) {
        float v = xr[c];
        if (pre == 1) v = gelu_f(v);
        const float d = v - mean;
        ss += d * d;
    }
    const float rstd = rsqrtf(block_sum_128(ss, sh) / C + 1e-5f);
    float* yr = y + (int64_t)row * ldy;
    for (int c = threadIdx.x; c < C; c += 128) {
        float v = xr[c];
        if (pre == 1) v = gelu_f(v);
        float o = (v - mean) * rstd * g[c] + (b ? b[c] : 0.f);
        if (res) o += res[(int64_t)row * ldr + c];
        if (round) o = tc::round_tf32(o);
        yr[c] = o;
    }
}

// ------------------------------------------------------------------------------------ small-M Linear (fp32 GEMV)
// y[m][o] = post( bias[o] + sum_k pre(x[m][k]) * W[o][k] ) (+ res[m][o]);  M <= 8 rows per pass, one warp per output feature.
// pre: 0 none, 1 SiLU ; post: 0 none, 1 SiLU, 2 sigmoid
template <int MT>
__global__ void __launch_bounds__(256) linear_small_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ Wt,
                                                          const float* __restrict__ bias, const float* __restrict__ res, int64_t ldr,
                                                          float* __restrict__ y, int64_t ldy, int M, int K, int O, int pre, int post, int round) {
    pdl_sync();
    extern __shared__ float xs[];  // [MT][K]
    const int m0 = blockIdx.y * MT;
    const int mrows = min(MT, M - m0);
    for (int i = threadIdx.x; i < MT * K; i += blockDim.x) {
        const int m = i / K, k = i - m * K;
        float v = (m < mrows) ? x[(int64_t)(m0 + m) * ldx + k] : 0.f;
        if (pre == 1) v = silu_f(v);
        xs[i] = v;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int o = blockIdx.x * (blockDim.x >> 5) + warp;
    if (o >= O) return;
    const float* wr = Wt + (int64_t)o * K;
    float acc[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) acc[m] = 0.f;
    if ((K & 3) == 0) {
        for (int k = lane * 4; k < K; k += 128) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(wr + k));
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                const float4 xv = *reinterpret_cast<const float4*>(xs + m * K + k);
                acc[m] += w.x * xv.x + w.y * xv.y + w.z * xv.z + w.w * xv.w;
            }
        }
    } else {
        for (int k = lane; k < K; k += 32) {
            const float w = __ldg(wr + k);
#pragma unroll
            for (int m = 0; m < MT; ++m) acc[m] += w * xs[m * K + k];
        }
    }
#pragma unroll
    for (int m = 0; m < MT; ++m) acc[m] = warp_sum(acc[m]);
    if (lane == 0) {
        const float bv = bias ? bias[o] : 0.f;
        for (int m = 0; m < mrows; ++m) {
            float v = acc[m] + bv;
            if (post == 1) v = silu_f(v);
            else if (post == 2) v = sigmoid_f(v);
            if (res) v += res[(int64_t)(m0 + m) * ldr + o];
            if (round) v = tc::round_tf32(v);
            y[(int64_t)(m0 + m) * ldy + o] = v;
        }
    }
}

// ------------------------------------------------------------------------------------ time embedding
// four[b] = [t, sin(2 pi t w_0..), cos(2 pi t w_0..)]   (imagen_pytorch.py:634-639)
__global__ void time_fourier_kernel(const float* __restrict__ t, const float* __restrict__ w, float* __restrict__ out, int B, int half) {
    pdl_sync();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int width = 2 * half + 1;
    if (i >= B * width) return;
    const int b = i / width, j = i - b * width;
    const float tv = t[b];
    float v;
    if (j == 0) v = tv;
    else {
        const int k = (j - 1) % half;
        const float f = tv * w[k] * 2.f * 3.14159265358979323846f;
        v = (j - 1 < half) ? sinf(f) : cosf(f);
    }
    out[i] = v;
}

// ------------------------------------------------------------------------------------ attention cores
// Multi-query self attention (imagen_pytorch.py:511-566): q [B,n,heads*dh] (already scaled? no: scaled here), kv [B,n,2*dh] shared by all
// heads, null_kv [2,dh], optional context kv ckv [B,nc,2*dh].  Key order: context, null, tokens (:523-532).  out [B,n,heads*dh].
// One warp per (b, head, query); dh <= 128.
__global__ void mq_attention_kernel(const float* __restrict__ q, const float* __restrict__ kv, const float* __restrict__ null_kv,
                                    const float* __restrict__ ckv, float* __restrict__ out, int B, int n, int heads, int dh, int nc, float scale, int round) {
    pdl_sync();
    extern __shared__ float sm[];  // per warp: scores[nk]
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gw = blockIdx.x * warps + warp;
    const int nk = nc + 1 + n;
    float* sc = sm + warp * nk;
    if (gw >= B * heads * n) return;
    const int i = gw % n, h = (gw / n) % heads, b = gw / (n * heads);
    const float* qr = q + ((int64_t)(b * n + i) * heads + h) * dh;
    auto key_ptr = [&](int j, int which) -> const float* {  // which 0 = k, 1 = v
        if (j < nc) return ckv + ((int64_t)(b * nc + j) * 2 + which) * dh;
        if (j == nc) return null_kv + which * dh;
        return kv + ((int64_t)(b * n + (j - nc - 1)) * 2 + which) * dh;
    };
    float mx = -INFINITY;
    for (int j = 0; j < nk; ++j) {
        const float* kr = key_ptr(j, 0);
        float d = 0.f;
        for (int e = lane; e < dh; e += 32) d += qr[e] * scale * kr[e];
        d = warp_sum(d);
        if (lane == 0) sc[j] = d;
        mx = fmaxf(mx, d);
    }
    __syncwarp();
    float den = 0.f;
    for (int j = 0; j < nk; ++j) den += __expf(sc[j] - mx);
    const float inv = 1.f / den;
    float* orow = out + ((int64_t)(b * n + i) * heads + h) * dh;
    for (int e = lane; e < dh; e += 32) {
        float acc = 0.f;
        for (int j = 0; j < nk; ++j) acc += __expf(sc[j] - mx) * inv * key_ptr(j, 1)[e];
        orow[e] = round ? tc::round_tf32(acc) : acc;
    }
}

// Cross attention (imagen_pytorch.py:764-805): q [B,n,heads*dh]; kvc [B,nc,2*heads*dh] = (k | v) per context token, per-head slices;
// null_kv [2,dh] shared by heads; keys: null, context.  One warp per (b, head, query).
__global__ void cross_attention_kernel(const float* __restrict__ q, const float* __restrict__ kvc, const float* __restrict__ null_kv,
                                       float* __restrict__ out, int B, int n, int heads, int dh, int nc, float scale, int round) {
    pdl_sync();
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gw = blockIdx.x * warps + warp;
    if (gw >= B * heads * n) return;
    const int i = gw % n, h = (gw / n) % heads, b = gw / (n * heads);
    const int inner = heads * dh;
    const float* qr = q + ((int64_t)(b * n + i)) * inner + h * dh;
    float scv[9];  // nc <= 8
    float mx = -INFINITY;
    for (int j = 0; j <= nc; ++j) {
        const float* kr = (j == 0) ? null_kv : kvc + ((int64_t)(b * nc + (j - 1))) * 2 * inner + h * dh;
        float d = 0.f;
        for (int e = lane; e < dh; e += 32) d += qr[e] * scale * kr[e];
        d = warp_sum(d);
        scv[j] = d;
        mx = fmaxf(mx, d);
    }
    float den = 0.f;
    for (int j = 0; j <= nc; ++j) { scv[j] = __expf(scv[j] - mx); den += scv[j]; }
    const float inv = 1.f / den;
    float* orow = out + ((int64_t)(b * n + i)) * inner + h * dh;
    for (int e = lane; e < dh; e += 32) {
        float acc = 0.f;
        for (int j = 0; j <= nc; ++j) {
            const float* vr = (j == 0) ? null_kv + dh : kvc + ((int64_t)(b * nc + (j - 1))) * 2 * inner + inner + h * dh;
            acc += scv[j] * inv * vr[e];
        }
        orow[e] = round ? tc::round_tf32(acc) : acc;
    }
}

// ------------------------------------------------------------------------------------ GlobalContext
// logits[n][p] = bias + sum_c x[n][p][c] * wk[c]      (to_k, imagen_pytorch.py:926,937) -- one warp per pixel
__global__ void gca_logits_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ wk, const float* __restrict__ bk,
                                  float* __restrict__ logits, int64_t npix, int C) {
    pdl_sync();
    const int64_t p = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= npix) return;
    const int lane = threadIdx.x & 31;
    const float* xr = x + p * ldx;
    float d = 0.f;
    for (int c = lane * 4; c < C; c += 128) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(xr + c));
        const float4 w = __ldg(reinterpret_cast<const float4*>(wk + c));
        d += v.x * w.x + v.y * w.y + v.z * w.z + v.w * w.w;
    }
    d = warp_sum(d);
    if (lane == 0) logits[p] = d + bk[0];
}
// softmax statistics of one image's logits -> stat[n] = (max, 1/sum exp); also zeroes pooled[n][:] for the accumulation pass
__global__ void __launch_bounds__(256) gca_stats_kernel(const float* __restrict__ logits, float2* __restrict__ stat, float* __restrict__ pooled,
                                                       int HW, int C) {
    pdl_sync();
    __shared__ float red[8];
    const int n = blockIdx.x;
    const float* lg = logits + (int64_t)n * HW;
    float mx = -INFINITY;
    for (int p = threadIdx.x; p < HW; p += 256) mx = fmaxf(mx, lg[p]);
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
    __syncthreads();
    float sm = 0.f;
    for (int p = threadIdx.x; p < HW; p += 256) sm += __expf(lg[p] - mx);
    sm = warp_sum(sm);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sm;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        stat[n] = make_float2(mx, 1.f / t);
    }
    for (int c = threadIdx.x; c < C; c += 256) pooled[(int64_t)n * C + c] = 0.f;
}
// pooled[n][c] += sum_{p in slab} softmax(logits)[p] * x[n][p][c]      grid (slabs of 16 pixels, NB), threads = channels
__global__ void __launch_bounds__(256) gca_pool_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ logits,
                                                      const float2* __restrict__ stat, float* __restrict__ pooled, int HW, int C) {
    pdl_sync();
    constexpr int SLAB = 16;
    __shared__ float wts[SLAB];
    const int n = blockIdx.y, p0 = blockIdx.x * SLAB;
    const int np = min(SLAB, HW - p0);
    const float2 st = stat[n];
    if (threadIdx.x < np) wts[threadIdx.x] = __expf(logits[(int64_t)n * HW + p0 + threadIdx.x] - st.x) * st.y;
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        const float* xb = x + ((int64_t)n * HW + p0) * ldx + c;
        float acc = 0.f;
#pragma unroll 4
        for (int p = 0; p < np; ++p) acc += wts[p] * __ldg(xb + (int64_t)p * ldx);
        atomicAdd(pooled + (int64_t)n * C + c, acc);
    }
}

// Single-launch GlobalContext pooling: one cluster of R CTAs per image, CTA r owns a slab of pixels.  Per CTA: logits of its pixels (warp per
// pixel), slab-local softmax statistics (m_r, l_r) and the slab-local weighted channel sums P_r[c] = sum_p exp(logit_p - m_r) x[p][c]; after a
// cluster barrier CTA r combines channel slice r of all slabs through distributed shared memory:
//   pooled[c] = sum_r e^{m_r - m} P_r[c] / sum_r e^{m_r - m} l_r,   m = max_r m_r            (the flash-attention merge; C <= 1024, slab <= 256 px)
constexpr int kGcaMaxC = 1024, kGcaMaxSlab = 256;
__global__ void __launch_bounds__(256) gca_fused_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ wk,
                                                       const float* __restrict__ bk, float* __restrict__ pooled, int HW, int C) {
    pdl_sync();
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ float wts[kGcaMaxSlab];
    __shared__ __align__(16) float part[1024];
    __shared__ __align__(16) float pp[kGcaMaxC];
    __shared__ float red[8];
    __shared__ float ml[2];
    const int R = (int)gridDim.x, sl = (int)blockIdx.x, n = blockIdx.y;
    const int p0 = (int)(((int64_t)HW * sl) / R), p1 = (int)(((int64_t)HW * (sl + 1)) / R);
    const int np = p1 - p0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* xb = x + ((int64_t)n * HW + p0) * ldx;
    const float b0 = __ldg(bk);
    // (1) logits
    for (int p = warp; p < np; p += 8) {
        const float* xr = xb + (int64_t)p * ldx;
        float d = 0.f;
        for (int c = lane * 4; c < C; c += 128) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(xr + c));
            const float4 w = __ldg(reinterpret_cast<const float4*>(wk + c));
            d += v.x * w.x + v.y * w.y + v.z * w.z + v.w * w.w;
        }
        d = warp_sum(d);
        if (lane == 0) wts[p] = d + b0;
    }
    __syncthreads();
    // (2) slab-local max and exp-sum; wts <- exp(logit - m_r)
    float mx = -INFINITY;
    for (int p = threadIdx.x; p < np; p += 256) mx = fmaxf(mx, wts[p]);
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
    __syncthreads();
    float sm = 0.f;
    for (int p = threadIdx.x; p < np; p += 256) {
        const float e = __expf(wts[p] - mx);
        wts[p] = e;
        sm += e;
    }
    sm = warp_sum(sm);
    if (lane == 0) red[warp] = sm;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        ml[0] = mx; ml[1] = t;
    }
    // (3) P_r[c]: thread = (pixel group, float4 column); groups split the slab's pixels, then fold through shared memory
    const int ncol = C >> 2;                       // <= 256
    const int ngrp = 256 / ncol;                   // >= 1
    const int grp = threadIdx.x / ncol, col = threadIdx.x - grp * ncol;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (grp < ngrp) {
#pragma unroll 4
        for (int p = grp; p < np; p += ngrp) {
            const float w = wts[p];
            const float4 v = __ldg(reinterpret_cast<const float4*>(xb + (int64_t)p * ldx) + col);
            acc.x += w * v.x; acc.y += w * v.y; acc.z += w * v.z; acc.w += w * v.w;
        }
        reinterpret_cast<float4*>(part)[grp * ncol + col] = acc;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        float t = 0.f;
        for (int g = 0; g < ngrp; ++g) t += part[g * C + c];
        pp[c] = t;
    }
    cluster.sync();
    // (4) merge channel slice `sl` across the cluster's slabs
    float m = -INFINITY;
    for (int r = 0; r < R; ++r) m = fmaxf(m, cluster.map_shared_rank(ml, r)[0]);
    float den = 0.f;
    for (int r = 0; r < R; ++r) {
        const float* rml = cluster.map_shared_rank(ml, r);
        den += __expf(rml[0] - m) * rml[1];
    }
    const int c0 = (int)(((int64_t)C * sl) / R), c1 = (int)(((int64_t)C * (sl + 1)) / R);
    for (int c = c0 + threadIdx.x; c < c1; c += 256) {
        float t = 0.f;
        for (int r = 0; r < R; ++r) t += __expf(cluster.map_shared_rank(ml, r)[0] - m) * cluster.map_shared_rank(pp, r)[c];
        pooled[(int64_t)n * C + c] = t / den;
    }
    cluster.sync();   // keep this CTA's shared memory alive until every remote read is done
}

// out = h * gate[n][c] + res      (ResnetBlock tail, imagen_pytorch.py:727-729); gate may be null (== 1)
__global__ void gate_residual_kernel(const float4* __restrict__ h, int64_t ldh_v, const float* __restrict__ gate, const float4* __restrict__ res,
                                     int64_t ldr_v, float4* __restrict__ out, int64_t ldo_v, int HW, int Cv, int64_t total) {
    pdl_sync();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pix = i / Cv;
        const int c = (int)(i - pix * Cv);
        float4 v = __ldg(h + pix * ldh_v + c);
        if (gate) {
            const int n = (int)(pix / HW);
            const float4 g = __ldg(reinterpret_cast<const float4*>(gate + (int64_t)n * Cv * 4) + c);
            v.x *= g.x; v.y *= g.y; v.z *= g.z; v.w *= g.w;
        }
        const float4 r = __ldg(res + pix * ldr_v + c);
        out[pix * ldo_v + c] = make_float4(v.x + r.x, v.y + r.y, v.z + r.z, v.w + r.w);
    }
}

static inline int ew_blocks(int64_t total, int threads = 256) {
    int64_t b = (total + threads - 1) / threads;
    const int64_t cap = (int64_t)sm_count() * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace sfb

using namespace sfb;

extern "C" {

int sfb_nchw_to_nhwc(const float* src, float* dst, int NB, int C, int H, int W, int64_t ld, int c_off, int round_tf32, void* stream) {
    SFB_REQUIRE(src && dst, "nchw_to_nhwc: null pointer");
    dim3 grid(ceil_div(H * W, 32), ceil_div(C, 32), NB), block(32, 8);
    launch_pdl(nchw_to_nhwc_kernel, grid, block, 0, as_stream(stream), src, dst, C, H * W, ld, c_off, round_tf32 && precision_mode() == 0);
    return check_launch("nchw_to_nhwc");
}

int sfb_nhwc_to_nchw(const float* src, float* dst, int NB, int C, int H, int W, int64_t ld, void* stream) {
    SFB_REQUIRE(src && dst, "nhwc_to_nchw: null pointer");
    dim3 grid(ceil_div(H * W, 32), ceil_div(C, 32), NB), block(32, 8);
    launch_pdl(nhwc_to_nchw_kernel, grid, block, 0, as_stream(stream), src, dst, C, H * W, ld);
    return check_launch("nhwc_to_nchw");
}

int sfb_concat2_nhwc(const float* a, int C1, int64_t lda, const float* b, int C2, int64_t ldb, float scale_b, float* out, int64_t ldo,
                     int64_t npix, void* stream) {
    SFB_REQUIRE(a && b && out, "concat2_nhwc: null pointer");
    SFB_REQUIRE(C1 % 4 == 0 && C2 % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0 && ldo % 4 == 0, "concat2_nhwc: channel counts must be multiples of 4");
    const int64_t total = npix * ((C1 + C2) / 4);
    launch_pdl(concat2_kernel, ew_blocks(total), 256, 0, as_stream(stream), reinterpret_cast<const float4*>(a), C1 / 4, lda / 4,
                                                                    reinterpret_cast<const float4*>(b), C2 / 4, ldb / 4, scale_b,
                                                                    reinterpret_cast<float4*>(out), ldo / 4, npix);
    return check_launch("concat2_nhwc");
}

int sfb_pixel_shuffle_silu_nhwc(const float* y, float* out, int NB, int H, int W, int Co, int64_t ldo, void* stream) {
    SFB_REQUIRE(y && out, "pixel_shuffle_silu: null pointer");
    const int64_t total = (int64_t)NB * 4 * H * W * Co;
    launch_pdl(pixel_shuffle_silu_kernel, ew_blocks(total), 256, 0, as_stream(stream), y, out, H, W, Co, ldo, total);
    return check_launch("pixel_shuffle_silu");
}

int sfb_groupnorm_ws_floats(int NB, int G) { return ((2 * NB * G + 3) / 4) * 4 + NB * G * 64 * 4; }

int sfb_groupnorm_nhwc(const float* x, int64_t ldx, int NB, int HW, int C, int G, const float* gamma, const float* beta, const float* film,
                       int64_t film_ld, int act_silu, float eps, float* stats_ws, unsigned int* counters, float* y, int64_t ldy, void* stream) {
    SFB_REQUIRE(x && gamma && beta && stats_ws && y, "groupnorm_nhwc: null pointer");
    SFB_REQUIRE(((uintptr_t)stats_ws & 15) == 0, "groupnorm_nhwc: workspace must be 16-byte aligned");
    SFB_REQUIRE(C % G == 0 && (C / G) % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0, "groupnorm_nhwc: channels per group must be a multiple of 4");
    cudaStream_t st = as_stream(stream);
    {   // single-launch cluster path: the (n, group) slab split over R <= 8 CTAs must fit 8 float4 per thread
        const int64_t group4 = (int64_t)HW * (C / G / 4);
        int R = 1;
        while (R < 8 && (group4 + R - 1) / R > 256 * kGnFusedItems) R *= 2;
        const bool fits = (((int64_t)HW + R - 1) / R) * (C / G / 4) <= 256 * kGnFusedItems && HW >= R && NB <= 65535;
        if (fits && gn_fused_enabled()) {
            launch_pdl_cluster(gn_fused_kernel, dim3(R, G, NB), dim3(256), 0, st, (unsigned)R, x, ldx, HW, C, G, eps, gamma, beta, film, film_ld, y, ldy,
                               act_silu, (int)(precision_mode() == 0));
            return check_launch("groupnorm_nhwc(fused)");
        }
    }
    // workspace: [0, 2*NB*G) floats = (mean, rstd); then 16-byte aligned fp64 partials [NB*G*S][2]
    int S = 1;
    while (S < 64 && G * NB * S * 2 <= sm_count() * 2 && HW / (S * 2) >= 8) S *= 2;
    float2* stats = reinterpret_cast<float2*>(stats_ws);
    double2* partial = reinterpret_cast<double2*>(stats_ws + (((size_t)2 * NB * G + 3) / 4) * 4);
    SFB_REQUIRE(counters != nullptr, "groupnorm_nhwc: counters workspace is null");
    launch_pdl(gn_stats_kernel, dim3(G, NB, S), 256, 0, st, x, ldx, HW, C, G, eps, S, partial, counters, stats);
    if (int rc = check_launch("groupnorm_nhwc(stats)")) return rc;
    const int64_t total4 = (int64_t)NB * HW * (C / 4);
    launch_pdl(gn_apply_kernel, ew_blocks(total4), 256, 0, st, x, ldx, reinterpret_cast<const float2*>(stats_ws), gamma, beta, film, film_ld, y, ldy, HW, C, G,
                                                      act_silu, precision_mode() == 0, total4);
    return check_launch("groupnorm_nhwc(apply)");
}

int sfb_layernorm_rows(const float* x, int64_t ldx, const float* g, const float* b, const float* res, int64_t ldr, float* y, int64_t ldy, int T,
                       int C, int pre_gelu, int round_tf32, void* stream) {
    SFB_REQUIRE(x && g && y, "layernorm_rows: null pointer");
    launch_pdl(layernorm_rows_kernel, T, 128, 0, as_stream(stream), x, ldx, g, b, res, ldr, y, ldy, T, C, pre_gelu, round_tf32 && precision_mode() == 0);
    return check_launch("layernorm_rows");
}

int sfb_linear_small(const float* x, int64_t ldx, const float* w, const float* bias, const float* res, int64_t ldr, float* y, int64_t ldy, int M,
                     int K, int O, int pre, int post, int round_tf32, void* stream) {
    SFB_REQUIRE(x && w && y, "linear_small: null pointer");
    SFB_REQUIRE(K <= 8192, "linear_small: K too large");
    cudaStream_t st = as_stream(stream);
    if (M <= 2) {
        const size_t sm = (size_t)2 * K * 4;
        static bool cfg = false;
        if (!cfg) { SFB_CUDA(cudaFuncSetAttribute(linear_small_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 8192 * 4)); cfg = true; }
        launch_pdl(linear_small_kernel<2>, dim3(ceil_div(O, 8), ceil_div(M, 2)), 256, sm, st, x, ldx, w, bias, res, ldr, y, ldy, M, K, O, pre, post, round_tf32 && precision_mode() == 0);
    } else {
        const size_t sm = (size_t)8 * K * 4;
        SFB_REQUIRE(sm <= 200 * 1024, "linear_small: K too large for 8-row tile");
        static bool cfg = false;
        if (!cfg) { SFB_CUDA(cudaFuncSetAttribute(linear_small_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); cfg = true; }
        launch_pdl(linear_small_kernel<8>, dim3(ceil_div(O, 8), ceil_div(M, 8)), 256, sm, st, x, ldx, w, bias, res, ldr, y, ldy, M, K, O, pre, post, round_tf32 && precision_mode() == 0);
    }
    return check_launch("linear_small");
}

int sfb_time_fourier(const float* t, const float* w, float* out, int B, int half, void* stream) {
    SFB_REQUIRE(t && w && out, "time_fourier: null pointer");
    const int total = B * (2 * half + 1);
    launch_pdl(time_fourier_kernel, ceil_div(total, 128), 128, 0, as_stream(stream), t, w, out, B, half);
    return check_launch("time_fourier");
}

int sfb_mq_attention(const float* q, const float* kv, const float* null_kv, const float* ckv, float* out, int B, int n, int heads, int dh, int nc,
                     float scale, void* stream) {
    SFB_REQUIRE(q && kv && null_kv && out && (nc == 0 || ckv), "mq_attention: null pointer");
    const int nk = nc + 1 + n;
    const int warps = 4;
    const size_t sm = (size_t)warps * nk * 4;
    SFB_REQUIRE(sm <= 48 * 1024, "mq_attention: too many keys for the single-pass kernel");
    launch_pdl(mq_attention_kernel, ceil_div(B * heads * n, warps), warps * 32, sm, as_stream(stream), q, kv, null_kv, ckv, out, B, n, heads, dh, nc, scale, precision_mode() == 0);
    return check_launch("mq_attention");
}

int sfb_cross_attention(const float* q, const float* kvc, const float* null_kv, float* out, int B, int n, int heads, int dh, int nc, float scale,
                        void* stream) {
    SFB_REQUIRE(q && kvc && null_kv && out, "cross_attention: null pointer");
    SFB_REQUIRE(nc <= 8, "cross_attention: at most 8 context tokens");
    launch_pdl(cross_attention_kernel, ceil_div(B * heads * n, 4), 128, 0, as_stream(stream), q, kvc, null_kv, out, B, n, heads, dh, nc, scale, precision_mode() == 0);
    return check_launch("cross_attention");
}

int sfb_gca_pool(const float* x, int64_t ldx, int NB, int HW, int C, const float* wk, const float* bk, float* logits_ws, float* pooled,
                 void* stream) {
    SFB_REQUIRE(x && wk && bk && logits_ws && pooled, "gca_pool: null pointer");
    SFB_REQUIRE(C % 4 == 0 && ldx % 4 == 0, "gca_pool: unsupported shape");
    cudaStream_t st = as_stream(stream);
    const int64_t npix = (int64_t)NB * HW;
    if (gca_fused_enabled() && C <= kGcaMaxC && 1024 % C == 0 && HW <= 8 * kGcaMaxSlab && NB <= 65535) {
        int R = 8;
        while (R > 1 && R > HW) R >>= 1;
        launch_pdl_cluster(gca_fused_kernel, dim3(R, NB), dim3(256), 0, st, (unsigned)R, x, ldx, wk, bk, pooled, HW, C);
        return check_launch("gca_pool(fused)");
    }
    launch_pdl(gca_logits_kernel, (unsigned)ceil_div(npix, (int64_t)8), 256, 0, st, x, ldx, wk, bk, logits_ws, npix, C);
    if (int rc = check_launch("gca_pool(logits)")) return rc;
    // logits_ws: NB*HW logits followed by NB (max, 1/sum) pairs
    float2* stat = reinterpret_cast<float2*>(logits_ws + ((npix + 1) / 2) * 2);
    launch_pdl(gca_stats_kernel, NB, 256, 0, st, logits_ws, stat, pooled, HW, C);
    if (int rc = check_launch("gca_pool(stats)")) return rc;
    launch_pdl(gca_pool_kernel, dim3(ceil_div(HW, 16), NB), 256, 0, st, x, ldx, logits_ws, stat, pooled, HW, C);
    return check_launch("gca_pool(pool)");
}

int sfb_gate_residual_nhwc(const float* h, int64_t ldh, const float* gate, const float* res, int64_t ldr, float* out, int64_t ldo, int NB, int HW,
                           int C, void* stream) {
    SFB_REQUIRE(h && res && out, "gate_residual: null pointer");
    SFB_REQUIRE(C % 4 == 0 && ldh % 4 == 0 && ldr % 4 == 0 && ldo % 4 == 0, "gate_residual: channel counts must be multiples of 4");
    const int64_t total = (int64_t)NB * HW * (C / 4);
    launch_pdl(gate_residual_kernel, ew_blocks(total), 256, 0, as_stream(stream), reinterpret_cast<const float4*>(h), ldh / 4, gate,
                                                                          reinterpret_cast<const float4*>(res), ldr / 4,
                                                                          reinterpret_cast<float4*>(out), ldo / 4, HW, C / 4, total);
    return check_launch("gate_residual");
}
}

// ------------------------------------------------------------------------------------ PLMS sampler update
// One fused pass for PLMSSampler.p_sample's tail (external/plms.py:143-154 + get_model_output :183-212):
//   e' = sum_i coef[i] * eps[i]                                   (Adams-Bashforth combination, :143-152)
//   x0 = clamp((x - sigma * e') / max(alpha, 1e-8), -clip, clip)  (predict_start_from_noise + clamp, :184,:205)
//   x_prev = alpha_next * (x * (1 - c) / alpha + c * x0) + noise_scale * noise     (q_posterior mean + sigma_t * z, :208-212)
// The per-step scalars (alpha, sigma, alpha_next, c, noise_scale = mask * exp(0.5 * log var)) are computed on the host from t, t_next.
namespace sfb {
__global__ void plms_update_kernel(const float* __restrict__ x, const float* __restrict__ e0, const float* __restrict__ e1,
                                   const float* __restrict__ e2, const float* __restrict__ e3, float c0, float c1, float c2, float c3,
                                   const float* __restrict__ noise, float alpha, float sigma, float alpha_next, float c, float noise_scale,
                                   float clip, float* __restrict__ x_prev, float* __restrict__ x0_out, float* __restrict__ e_out, int64_t n) {
    pdl_sync();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float e = c0 * e0[i];
        if (e1) e += c1 * e1[i];
        if (e2) e += c2 * e2[i];
        if (e3) e += c3 * e3[i];
        const float xv = x[i];
        float x0 = (xv - sigma * e) / fmaxf(alpha, 1e-8f);
        x0 = fminf(fmaxf(x0, -clip), clip);
        const float mean = alpha_next * (xv * (1.f - c) / alpha + c * x0);
        x_prev[i] = mean + noise_scale * noise[i];
        if (x0_out) x0_out[i] = x0;
        if (e_out) e_out[i] = e;
    }
}
}  // namespace sfb

extern "C" int sfb_plms_update(const float* x, const float* e0, const float* e1, const float* e2, const float* e3, float c0, float c1, float c2,
                               float c3, const float* noise, float alpha, float sigma, float alpha_next, float c, float noise_scale, float clip,
                               float* x_prev, float* x0_out, float* e_out, int64_t n, void* stream) {
    if (n == 0) return SFB_OK;
    SFB_REQUIRE(x && e0 && noise && x_prev, "plms_update: null pointer");
    int64_t blocks = (n + 255) / 256;
    if (blocks > 1024) blocks = 1024;
    sfb::launch_pdl(sfb::plms_update_kernel, (unsigned)blocks, 256, 0, sfb::as_stream(stream), x, e0, e1, e2, e3, c0, c1, c2, c3, noise, alpha, sigma, alpha_next, c,
                                                                                 noise_scale, clip, x_prev, x0_out, e_out, n);
    return sfb::check_launch("plms_update");
}
