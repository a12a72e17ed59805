// gridencoder.cuh -- device helpers shared by the grid-encoder operators and the fused NGP field kernels.
//
// The integer contract (which embedding row each interpolation corner reads) must be bit-identical to
// the reference (external/gridencoder/src/gridencoder.cu:54-72, :124-137), so the expressions below
// keep the reference's precision and operation order: scale = exp2f(level*S)*H - 1 in fp32 on the
// device, resolution = ceil(scale)+1 through a double ceil, cell = floorf(x*scale + 0.5).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sfb {

struct GridLevel {
    float scale;
    uint32_t resolution;
    uint32_t rows;    // rows of this level ("hashmap_size")
    uint32_t offset;  // first row of this level in the table
};

__device__ __forceinline__ float grid_level_scale(uint32_t level, float S, uint32_t H) {
    return exp2f(level * S) * H - 1.0f;
}

__device__ __forceinline__ GridLevel grid_level(uint32_t level, float S, uint32_t H, const int32_t* __restrict__ offsets) {
    GridLevel lv;
    lv.offset = (uint32_t)offsets[level];
    lv.rows = (uint32_t)offsets[level + 1] - lv.offset;
    lv.scale = grid_level_scale(level, S, H);
    lv.resolution = (uint32_t)ceil((double)lv.scale) + 1;
    return lv;
}

template <uint32_t D>
__device__ __forceinline__ bool grid_out_of_range(const float (&x)[D]) {
    bool oob = false;
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) oob |= (x[d] < 0 || x[d] > 1);
    return oob;
}

template <uint32_t D>
__device__ __forceinline__ void grid_locate(const float (&x)[D], float scale, bool align_corners, float (&frac)[D], uint32_t (&cell)[D]) {
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) {
        float p = x[d] * scale + (align_corners ? 0.0f : 0.5f);
        cell[d] = (uint32_t)floorf(p);
        frac[d] = p - (float)cell[d];
    }
}

template <uint32_t D>
__device__ __forceinline__ uint32_t grid_hash(const uint32_t (&p)[D]) {
    constexpr uint32_t primes[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};
    uint32_t r = 0;
#pragma unroll
    for (uint32_t i = 0; i < D; ++i) r ^= p[i] * primes[i];
    return r;
}

// row (not element) index of a grid vertex inside its level
template <uint32_t D>
__device__ __forceinline__ uint32_t grid_row(uint32_t gridtype, bool align_corners, uint32_t rows, uint32_t resolution,
                                             const uint32_t (&p)[D]) {
    uint32_t stride = 1, index = 0;
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) {
        if (stride <= rows) {  // once false it stays false: stride only grows
            index += p[d] * stride;
            stride *= align_corners ? resolution : (resolution + 1);
        }
    }
    if (gridtype == 0 && stride > rows) index = grid_hash<D>(p);
    return index % rows;
}

template <uint32_t C>
__device__ __forceinline__ void grid_load_row(const float* __restrict__ base, uint32_t row, float (&v)[C]) {
    if constexpr (C == 2) {
        const float2 t = __ldg(reinterpret_cast<const float2*>(base) + row);
        v[0] = t.x; v[1] = t.y;
    } else if constexpr (C == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(base) + row);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else if constexpr (C == 8) {
        const float4 t0 = __ldg(reinterpret_cast<const float4*>(base) + 2 * (size_t)row);
        const float4 t1 = __ldg(reinterpret_cast<const float4*>(base) + 2 * (size_t)row + 1);
        v[0] = t0.x; v[1] = t0.y; v[2] = t0.z; v[3] = t0.w; v[4] = t1.x; v[5] = t1.y; v[6] = t1.z; v[7] = t1.w;
    } else {
#pragma unroll
        for (uint32_t c = 0; c < C; ++c) v[c] = __ldg(base + (size_t)row * C + c);
    }
}

template <uint32_t C>
__device__ __forceinline__ void grid_store_row(float* __restrict__ dst, const float (&v)[C]) {
    if constexpr (C == 2) {
        *reinterpret_cast<float2*>(dst) = make_float2(v[0], v[1]);
    } else if constexpr (C == 4) {
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
        for (uint32_t c = 0; c < C; ++c) dst[c] = v[c];
    }
}

// fire-and-forget reduction into the gradient table: red.global.add(.v2).f32
template <uint32_t C>
__device__ __forceinline__ void grid_red_add_row(float* __restrict__ base, uint32_t row, const float (&v)[C]) {
    if constexpr (C % 2 == 0) {
#pragma unroll
        for (uint32_t c = 0; c < C; c += 2) {
            float* p = base + (size_t)row * C + c;
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v[c]), "f"(v[c + 1]) : "memory");
        }
    } else {
#pragma unroll
        for (uint32_t c = 0; c < C; ++c) atomicAdd(base + (size_t)row * C + c, v[c]);
    }
}

}  // namespace sfb
