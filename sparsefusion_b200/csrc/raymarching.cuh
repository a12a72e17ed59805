// raymarching.cuh -- device helpers shared by the raymarching operators and the fused render kernels.
// Arithmetic follows raymarching/src/raymarching.cu of the reference bit for bit where integers are
// derived from floats (mip level, occupancy cell, morton code); see the citations on each helper.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>

namespace sfb {

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }
__device__ __forceinline__ float signf(float x) { return copysignf(1.0f, x); }

// raymarching.cu:56-81
__host__ __device__ __forceinline__ uint32_t morton_expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__host__ __device__ __forceinline__ uint32_t morton3D_encode(uint32_t x, uint32_t y, uint32_t z) {
    return morton_expand_bits(x) | (morton_expand_bits(y) << 1) | (morton_expand_bits(z) << 2);
}
__host__ __device__ __forceinline__ uint32_t morton3D_decode(uint32_t x) {
    x = x & 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

// raymarching.cu:42-54
__device__ __forceinline__ int mip_from_pos(float x, float y, float z, float max_cascade) {
    const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    int exponent;
    frexpf(mx, &exponent);
    return fminf(max_cascade - 1, fmaxf(0, exponent));
}
__device__ __forceinline__ int mip_from_dt(float dt, float H, float max_cascade) {
    const float mx = dt * H * 0.5;
    int exponent;
    frexpf(mx, &exponent);
    return fminf(max_cascade - 1, fmaxf(0, exponent));
}

// slab test, raymarching.cu:108-144.  A miss reports near == far == FLT_MAX.
__device__ __forceinline__ void ray_aabb(const float* __restrict__ o, const float* __restrict__ d, const float* __restrict__ aabb,
                                         float min_near, float& near, float& far) {
    const float ox = o[0], oy = o[1], oz = o[2];
    const float rdx = 1 / d[0], rdy = 1 / d[1], rdz = 1 / d[2];
    near = (aabb[0] - ox) * rdx;
    far = (aabb[3] - ox) * rdx;
    if (near > far) { const float t = near; near = far; far = t; }
    float near_y = (aabb[1] - oy) * rdy, far_y = (aabb[4] - oy) * rdy;
    if (near_y > far_y) { const float t = near_y; near_y = far_y; far_y = t; }
    if (near > far_y || near_y > far) { near = far = FLT_MAX; return; }
    if (near_y > near) near = near_y;
    if (far_y < far) far = far_y;
    float near_z = (aabb[2] - oz) * rdz, far_z = (aabb[5] - oz) * rdz;
    if (near_z > far_z) { const float t = near_z; near_z = far_z; far_z = t; }
    if (near > far_z || near_z > far) { near = far = FLT_MAX; return; }
    if (near_z > near) near = near_z;
    if (far_z < far) far = far_z;
    if (near < min_near) near = min_near;
}

struct MarchSample {
    float x, y, z, dt;
    float mip_bound;
    int nx, ny, nz;
};

// DDA through the cascaded occupancy bitfield: raymarching.cu:359-400 / :427-479 / :750-804
struct Marcher {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz, rH, H3, bound, dt_gamma, dt_min, dt_max;
    uint32_t C, H;
    const uint8_t* __restrict__ grid;

    __device__ __forceinline__ Marcher(const float* __restrict__ o, const float* __restrict__ d, const uint8_t* __restrict__ grid_, float bound_,
                                       float dt_gamma_, uint32_t max_steps, uint32_t C_, uint32_t H_) {
        const float SQRT3 = 1.7320508075688772f;
        ox = o[0]; oy = o[1]; oz = o[2];
        dx = d[0]; dy = d[1]; dz = d[2];
        rdx = 1 / dx; rdy = 1 / dy; rdz = 1 / dz;
        rH = 1 / (float)H_;
        H3 = H_ * H_ * H_;
        bound = bound_; dt_gamma = dt_gamma_;
        dt_min = 2 * SQRT3 / max_steps;
        dt_max = 2 * SQRT3 * (1 << (C_ - 1)) / H_;
        C = C_; H = H_; grid = grid_;
    }

    // evaluate the sample at t; returns whether its occupancy cell is set
    __device__ __forceinline__ bool probe(float t, MarchSample& s) const {
        s.x = clampf(ox + t * dx, -bound, bound);
        s.y = clampf(oy + t * dy, -bound, bound);
        s.z = clampf(oz + t * dz, -bound, bound);
        s.dt = clampf(t * dt_gamma, dt_min, dt_max);
        const int level = max(mip_from_pos(s.x, s.y, s.z, C), mip_from_dt(s.dt, H, C));
        s.mip_bound = fminf(scalbnf(1.0f, level), bound);
        const float mip_rbound = 1 / s.mip_bound;
        // double-precision product, converted to float by clampf's parameter, truncated to int
        s.nx = clampf(0.5 * (s.x * mip_rbound + 1) * H, 0.0f, (float)(H - 1));
        s.ny = clampf(0.5 * (s.y * mip_rbound + 1) * H, 0.0f, (float)(H - 1));
        s.nz = clampf(0.5 * (s.z * mip_rbound + 1) * H, 0.0f, (float)(H - 1));
        const uint32_t index = level * H3 + morton3D_encode(s.nx, s.ny, s.nz);
        return (grid[index / 8] & (1 << (index % 8))) != 0;
    }

    // empty cell: advance t by whole steps until it leaves the cell
    __device__ __forceinline__ void skip(float& t, const MarchSample& s) const {
        const float tx = (((s.nx + 0.5f + 0.5f * signf(dx)) * rH * 2 - 1) * s.mip_bound - s.x) * rdx;
        const float ty = (((s.ny + 0.5f + 0.5f * signf(dy)) * rH * 2 - 1) * s.mip_bound - s.y) * rdy;
        const float tz = (((s.nz + 0.5f + 0.5f * signf(dz)) * rH * 2 - 1) * s.mip_bound - s.z) * rdz;
        const float tt = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
        do {
            t += clampf(t * dt_gamma, dt_min, dt_max);
        } while (t < tt);
    }
};

}  // namespace sfb
