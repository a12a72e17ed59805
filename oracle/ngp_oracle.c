/*
 * oracle/ngp_oracle.c -- CPU restatement of the reference's CUDA-only NGP operators.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under sparsefusion_b200/ may link, import or call this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, as the checker (never as the thing shipped).
 *
 * The reference has no CPU path for these operators (the pybind modules `_gridencoder` and
 * `_raymarching` are CUDA-only), so this file restates the arithmetic of each kernel in scalar
 * C99, one function per reference kernel, citing the file:line it follows.  Where nvcc's default
 * -fmad=true contracts a*b+c in the reference kernel the restatement calls fmaf() explicitly
 * and the file is compiled with -ffp-contract=off so that nothing else is fused.
 *
 * Parity status: the reference holds no tests or golden vectors for this path,
 * so the restatement is pinned two ways: (1) tests/golden/ngp_*.npz are produced by
 * oracle/gen_golden.py from THIS restatement and cross-checked on the GPU box against the
 * reference's own CUDA sources compiled into oracle/_ref/ (tests/test_ref_cuda_gpu.py);
 * (2) property tests (partition of unity, linearity in the embedding table, adjointness of
 * forward and backward).
 *
 * Device-libm caveat: the reference computes the per-level scale with the device's exp2f
 * (gridencoder.cu:125) which is not correctly rounded.  Callers may therefore pass an explicit
 * `level_scales` array (as computed on the device); when NULL, exp2f from the host libm is used.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------ */
/* grid encoder -- external/gridencoder/src/gridencoder.cu                                    */
/* ------------------------------------------------------------------------------------------ */

/* gridencoder.cu:35-51 (fast_hash) */
static uint32_t oracle_fast_hash(const uint32_t *pos, uint32_t D) {
    static const uint32_t primes[7] = {1u, 2654435761u, 805459861u, 3674653429u,
                                       2097192037u, 1434869437u, 2165219737u};
    uint32_t r = 0;
    for (uint32_t i = 0; i < D; ++i) r ^= pos[i] * primes[i];
    return r;
}

/* gridencoder.cu:54-72 (get_grid_index).  Returns the ROW index (the reference multiplies by C
 * and adds ch afterwards).  Note the stride loop stops as soon as stride > hashmap_size, which
 * drops the trailing coordinates for the tiled grid type. */
static uint32_t oracle_grid_row(uint32_t gridtype, int align_corners, uint32_t hashmap_size,
                                uint32_t resolution, const uint32_t *pos, uint32_t D) {
    uint32_t stride = 1, index = 0;
    for (uint32_t d = 0; d < D && stride <= hashmap_size; ++d) {
        index += pos[d] * stride;
        stride *= align_corners ? resolution : (resolution + 1);
    }
    if (gridtype == 0 && stride > hashmap_size) index = oracle_fast_hash(pos, D);
    return index % hashmap_size;
}

/* gridencoder.cu:124-126 */
float oracle_grid_level_scale(uint32_t level, float S, uint32_t H) {
    return fmaf(exp2f((float)level * S), (float)H, -1.0f);
}

typedef struct {
    float scale;
    uint32_t resolution;
    uint32_t hashmap_size;
} level_info_t;

static level_info_t oracle_level_info(uint32_t level, float S, uint32_t H, const int32_t *offsets,
                                      const float *level_scales) {
    level_info_t li;
    li.scale = level_scales ? level_scales[level] : oracle_grid_level_scale(level, S, H);
    li.resolution = (uint32_t)ceil((double)li.scale) + 1u; /* gridencoder.cu:126 (double ceil) */
    li.hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
    return li;
}

/* gridencoder.cu:129-137: pos = fma(x, scale, 0.5 or 0); pos_grid = floorf(pos); pos -= pos_grid */
static void oracle_locate(const float *x, uint32_t D, float scale, int align_corners,
                          float *frac, uint32_t *cell) {
    for (uint32_t d = 0; d < D; ++d) {
        float p = fmaf(x[d], scale, align_corners ? 0.0f : 0.5f);
        float fl = floorf(p);
        cell[d] = (uint32_t)fl;
        frac[d] = p - (float)cell[d];
    }
}

/* Forward: gridencoder.cu:75-223.  inputs [B,D] in [0,1]; embeddings [rows,C]; offsets [L+1];
 * outputs [L,B,C]; dy_dx [B,L,D,C] or NULL; corner_rows (debug, may be NULL) [L,B,2^D] receives
 * the absolute row index (offsets[level]+row) of every corner, -1 for out-of-range points. */
void oracle_grid_encode_forward(const float *inputs, const float *embeddings, const int32_t *offsets,
                                float *outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L,
                                float S, uint32_t H, float *dy_dx, uint32_t gridtype,
                                int align_corners, const float *level_scales, int32_t *corner_rows) {
    const uint32_t ncorner = 1u << D;
#pragma omp parallel for schedule(static)
    for (int64_t lb = 0; lb < (int64_t)L * B; ++lb) {
        const uint32_t level = (uint32_t)(lb / B), b = (uint32_t)(lb % B);
        const float *x = inputs + (size_t)b * D;
        float *out = outputs + ((size_t)level * B + b) * C;
        const float *grid = embeddings + (size_t)(uint32_t)offsets[level] * C;
        float *dy = dy_dx ? dy_dx + (size_t)b * D * L * C + (size_t)level * D * C : NULL;
        int oob = 0;
        for (uint32_t d = 0; d < D; ++d)
            if (x[d] < 0 || x[d] > 1) oob = 1;
        if (oob) { /* gridencoder.cu:98-122 */
            for (uint32_t c = 0; c < C; ++c) out[c] = 0;
            if (dy) for (uint32_t i = 0; i < D * C; ++i) dy[i] = 0;
            if (corner_rows)
                for (uint32_t i = 0; i < ncorner; ++i) corner_rows[((size_t)level * B + b) * ncorner + i] = -1;
            continue;
        }
        level_info_t li = oracle_level_info(level, S, H, offsets, level_scales);
        float frac[8];
        uint32_t cell[8], cl[8];
        oracle_locate(x, D, li.scale, align_corners, frac, cell);
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (uint32_t idx = 0; idx < ncorner; ++idx) { /* gridencoder.cu:145-169 */
            float w = 1;
            for (uint32_t d = 0; d < D; ++d) {
                if ((idx & (1u << d)) == 0) { w *= 1 - frac[d]; cl[d] = cell[d]; }
                else { w *= frac[d]; cl[d] = cell[d] + 1; }
            }
            uint32_t row = oracle_grid_row(gridtype, align_corners, li.hashmap_size, li.resolution, cl, D);
            if (corner_rows) corner_rows[((size_t)level * B + b) * ncorner + idx] = (int32_t)(offsets[level] + row);
            for (uint32_t c = 0; c < C; ++c) acc[c] = fmaf(w, grid[(size_t)row * C + c], acc[c]);
        }
        for (uint32_t c = 0; c < C; ++c) out[c] = acc[c];
        if (dy) { /* gridencoder.cu:179-222 */
            for (uint32_t gd = 0; gd < D; ++gd) {
                float g[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                for (uint32_t idx = 0; idx < (1u << (D - 1)); ++idx) {
                    float w = li.scale;
                    for (uint32_t nd = 0; nd < D - 1; ++nd) {
                        uint32_t d = (nd >= gd) ? nd + 1 : nd;
                        if ((idx & (1u << nd)) == 0) { w *= 1 - frac[d]; cl[d] = cell[d]; }
                        else { w *= frac[d]; cl[d] = cell[d] + 1; }
                    }
                    cl[gd] = cell[gd];
                    uint32_t rl = oracle_grid_row(gridtype, align_corners, li.hashmap_size, li.resolution, cl, D);
                    cl[gd] = cell[gd] + 1;
                    uint32_t rr = oracle_grid_row(gridtype, align_corners, li.hashmap_size, li.resolution, cl, D);
                    for (uint32_t c = 0; c < C; ++c)
                        g[c] = fmaf(w, grid[(size_t)rr * C + c] - grid[(size_t)rl * C + c], g[c]);
                }
                for (uint32_t c = 0; c < C; ++c) dy[gd * C + c] = g[c];
            }
        }
    }
}

/* Backward: gridencoder.cu:226-313 (scatter w*grad into grad_embeddings, which the caller has
 * zeroed -- grid.py:72) and gridencoder.cu:316-342 (grad_inputs from dy_dx).  The reference uses
 * float atomicAdd in undefined order; this restatement accumulates in double and rounds once,
 * which is the value every ordering converges to; comparisons use a tolerance. */
void oracle_grid_encode_backward(const float *grad, const float *inputs, const int32_t *offsets,
                                 float *grad_embeddings, uint32_t B, uint32_t D, uint32_t C,
                                 uint32_t L, float S, uint32_t H, const float *dy_dx,
                                 float *grad_inputs, uint32_t gridtype, int align_corners,
                                 const float *level_scales) {
    const uint32_t ncorner = 1u << D;
    const size_t rows = (size_t)offsets[L];
    double *acc = (double *)calloc(rows * C, sizeof(double));
    for (uint32_t level = 0; level < L; ++level) {
        level_info_t li = oracle_level_info(level, S, H, offsets, level_scales);
        double *g = acc + (size_t)(uint32_t)offsets[level] * C;
        for (uint32_t b = 0; b < B; ++b) {
            const float *x = inputs + (size_t)b * D;
            int oob = 0;
            for (uint32_t d = 0; d < D; ++d)
                if (x[d] < 0 || x[d] > 1) oob = 1;
            if (oob) continue;
            float frac[8];
            uint32_t cell[8], cl[8];
            oracle_locate(x, D, li.scale, align_corners, frac, cell);
            const float *gr = grad + ((size_t)level * B + b) * C;
            for (uint32_t idx = 0; idx < ncorner; ++idx) {
                float w = 1;
                for (uint32_t d = 0; d < D; ++d) {
                    if ((idx & (1u << d)) == 0) { w *= 1 - frac[d]; cl[d] = cell[d]; }
                    else { w *= frac[d]; cl[d] = cell[d] + 1; }
                }
                uint32_t row = oracle_grid_row(gridtype, align_corners, li.hashmap_size, li.resolution, cl, D);
                for (uint32_t c = 0; c < C; ++c) g[(size_t)row * C + c] += (double)(w * gr[c]);
            }
        }
    }
    for (size_t i = 0; i < rows * C; ++i) grad_embeddings[i] = (float)((double)grad_embeddings[i] + acc[i]);
    free(acc);
    if (dy_dx && grad_inputs) {
        for (uint32_t t = 0; t < B * D; ++t) {
            uint32_t b = t / D, d = t % D;
            const float *dy = dy_dx + (size_t)b * L * D * C;
            float r = 0;
            for (uint32_t l = 0; l < L; ++l)
                for (uint32_t c = 0; c < C; ++c)
                    r = fmaf(grad[((size_t)l * B + b) * C + c], dy[(size_t)l * D * C + d * C + c], r);
            grad_inputs[t] = r;
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* raymarching -- raymarching/src/raymarching.cu                                               */
/* ------------------------------------------------------------------------------------------ */

static inline float oracle_clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }
static inline float oracle_signf(float x) { return copysignf(1.0f, x); }

/* raymarching.cu:56-81 */
static inline uint32_t oracle_expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
static inline uint32_t oracle_morton3D_1(uint32_t x, uint32_t y, uint32_t z) {
    return oracle_expand_bits(x) | (oracle_expand_bits(y) << 1) | (oracle_expand_bits(z) << 2);
}
static inline uint32_t oracle_morton3D_invert_1(uint32_t x) {
    x = x & 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

/* raymarching.cu:42-54 */
static inline int oracle_mip_from_pos(float x, float y, float z, float max_cascade) {
    const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    int e;
    frexpf(mx, &e);
    return (int)fminf(max_cascade - 1, fmaxf(0, (float)e));
}
static inline int oracle_mip_from_dt(float dt, float H, float max_cascade) {
    const float mx = (float)((double)(dt * H) * 0.5);
    int e;
    frexpf(mx, &e);
    return (int)fminf(max_cascade - 1, fmaxf(0, (float)e));
}

/* raymarching.cu:91-145 */
void oracle_near_far_from_aabb(const float *rays_o, const float *rays_d, const float *aabb,
                               uint32_t N, float min_near, float *nears, float *fars) {
    for (uint32_t n = 0; n < N; ++n) {
        const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
        const float rdx = 1 / rays_d[n * 3], rdy = 1 / rays_d[n * 3 + 1], rdz = 1 / rays_d[n * 3 + 2];
        float near = (aabb[0] - ox) * rdx, far = (aabb[3] - ox) * rdx, t;
        if (near > far) { t = near; near = far; far = t; }
        float ny = (aabb[1] - oy) * rdy, fy = (aabb[4] - oy) * rdy;
        if (ny > fy) { t = ny; ny = fy; fy = t; }
        if (near > fy || ny > far) { nears[n] = fars[n] = FLT_MAX; continue; }
        if (ny > near) near = ny;
        if (fy < far) far = fy;
        float nz = (aabb[2] - oz) * rdz, fz = (aabb[5] - oz) * rdz;
        if (nz > fz) { t = nz; nz = fz; fz = t; }
        if (near > fz || nz > far) { nears[n] = fars[n] = FLT_MAX; continue; }
        if (nz > near) near = nz;
        if (fz < far) far = fz;
        if (near < min_near) near = min_near;
        nears[n] = near;
        fars[n] = far;
    }
}

/* raymarching.cu:162-198 */
void oracle_sph_from_ray(const float *rays_o, const float *rays_d, float radius, uint32_t N, float *coords) {
    const float RPI = 0.3183098861837907f;
    for (uint32_t n = 0; n < N; ++n) {
        const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
        const float dx = rays_d[n * 3], dy = rays_d[n * 3 + 1], dz = rays_d[n * 3 + 2];
        const float A = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        const float Bh = fmaf(oz, dz, fmaf(oy, dy, ox * dx));
        const float Cc = fmaf(oz, oz, fmaf(oy, oy, ox * ox)) - radius * radius;
        const float t = (-Bh + sqrtf(fmaf(Bh, Bh, -(A * Cc)))) / A;
        const float x = fmaf(t, dx, ox), y = fmaf(t, dy, oy), z = fmaf(t, dz, oz);
        const float theta = atan2f(sqrtf(fmaf(z, z, x * x)), y);
        const float phi = atan2f(z, x);
        coords[n * 2] = fmaf(2 * theta, RPI, -1.0f);
        coords[n * 2 + 1] = phi * RPI;
    }
}

/* raymarching.cu:214-260 */
void oracle_morton3D(const int32_t *coords, uint32_t N, int32_t *indices) {
    for (uint32_t n = 0; n < N; ++n)
        indices[n] = (int32_t)oracle_morton3D_1((uint32_t)coords[n * 3], (uint32_t)coords[n * 3 + 1], (uint32_t)coords[n * 3 + 2]);
}
void oracle_morton3D_invert(const int32_t *indices, uint32_t N, int32_t *coords) {
    for (uint32_t n = 0; n < N; ++n) {
        const int32_t ind = indices[n]; /* arithmetic shift of a signed int, as in the reference */
        coords[n * 3] = (int32_t)oracle_morton3D_invert_1((uint32_t)(ind >> 0));
        coords[n * 3 + 1] = (int32_t)oracle_morton3D_invert_1((uint32_t)(ind >> 1));
        coords[n * 3 + 2] = (int32_t)oracle_morton3D_invert_1((uint32_t)(ind >> 2));
    }
}

/* raymarching.cu:267-289 */
void oracle_packbits(const float *grid, uint32_t N, float density_thresh, uint8_t *bitfield) {
    for (uint32_t n = 0; n < N; ++n) {
        uint8_t bits = 0;
        for (int i = 0; i < 8; ++i) bits |= (grid[(size_t)n * 8 + i] > density_thresh) ? (uint8_t)(1u << i) : 0;
        bitfield[n] = bits;
    }
}

/* One DDA state update shared by the train and inference marchers (raymarching.cu:359-400,
 * 427-479, 750-804).  Returns 1 if the cell at t is occupied. */
typedef struct {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz, rH, H3, bound, dt_gamma, dt_min, dt_max;
    uint32_t C, H;
    const uint8_t *grid;
} march_ctx_t;

static void march_ctx_init(march_ctx_t *m, const float *o, const float *d, const uint8_t *grid, float bound,
                           float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H) {
    const float SQRT3 = 1.7320508075688772f;
    m->ox = o[0]; m->oy = o[1]; m->oz = o[2];
    m->dx = d[0]; m->dy = d[1]; m->dz = d[2];
    m->rdx = 1 / m->dx; m->rdy = 1 / m->dy; m->rdz = 1 / m->dz;
    m->rH = 1 / (float)H;
    m->H3 = (float)(H * H * H);
    m->bound = bound; m->dt_gamma = dt_gamma;
    m->dt_min = 2 * SQRT3 / (float)max_steps;
    m->dt_max = 2 * SQRT3 * (float)(1 << (C - 1)) / (float)H;
    m->C = C; m->H = H; m->grid = grid;
}

static int march_probe(const march_ctx_t *m, float t, float *px, float *py, float *pz, float *pdt, float *ptt) {
    const float x = oracle_clampf(fmaf(t, m->dx, m->ox), -m->bound, m->bound);
    const float y = oracle_clampf(fmaf(t, m->dy, m->oy), -m->bound, m->bound);
    const float z = oracle_clampf(fmaf(t, m->dz, m->oz), -m->bound, m->bound);
    const float dt = oracle_clampf(t * m->dt_gamma, m->dt_min, m->dt_max);
    const int la = oracle_mip_from_pos(x, y, z, (float)m->C), lb = oracle_mip_from_dt(dt, (float)m->H, (float)m->C);
    const int level = la > lb ? la : lb;
    const float mip_bound = fminf(scalbnf(1.0f, level), m->bound);
    const float mip_rbound = 1 / mip_bound;
    /* 0.5 * (x * mip_rbound + 1) * H evaluated in double, converted to float for clamp(), then
     * truncated to int (raymarching.cu:374-376) */
    const int nx = (int)oracle_clampf((float)(0.5 * (double)fmaf(x, mip_rbound, 1.0f) * (double)m->H), 0.0f, (float)(m->H - 1));
    const int ny = (int)oracle_clampf((float)(0.5 * (double)fmaf(y, mip_rbound, 1.0f) * (double)m->H), 0.0f, (float)(m->H - 1));
    const int nz = (int)oracle_clampf((float)(0.5 * (double)fmaf(z, mip_rbound, 1.0f) * (double)m->H), 0.0f, (float)(m->H - 1));
    const uint32_t index = (uint32_t)((float)level * m->H3 + (float)oracle_morton3D_1((uint32_t)nx, (uint32_t)ny, (uint32_t)nz));
    const int occ = (m->grid[index / 8] & (1u << (index % 8))) != 0;
    *px = x; *py = y; *pz = z; *pdt = dt;
    if (!occ) {
        const float tx = (fmaf(fmaf(0.5f, oracle_signf(m->dx), (float)nx + 0.5f) * m->rH, 2.0f, -1.0f) * mip_bound - x) * m->rdx;
        const float ty = (fmaf(fmaf(0.5f, oracle_signf(m->dy), (float)ny + 0.5f) * m->rH, 2.0f, -1.0f) * mip_bound - y) * m->rdy;
        const float tz = (fmaf(fmaf(0.5f, oracle_signf(m->dz), (float)nz + 0.5f) * m->rH, 2.0f, -1.0f) * mip_bound - z) * m->rdz;
        *ptt = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
    }
    return occ;
}

/* raymarching.cu:311-480.  The reference claims point ranges with atomicAdd in arbitrary order;
 * the restatement claims them in ray order, so `rays` is [n, offset, num_steps] with ray_index==n.
 * Compare per ray through the rays table, not positionally. */
void oracle_march_rays_train(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound,
                             float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                             uint32_t M, const float *nears, const float *fars, float *xyzs, float *dirs,
                             float *deltas, int32_t *rays, int32_t *counter, const float *noises) {
    for (uint32_t n = 0; n < N; ++n) {
        march_ctx_t m;
        march_ctx_init(&m, rays_o + n * 3, rays_d + n * 3, grid, bound, dt_gamma, max_steps, C, H);
        const float far = fars[n];
        float t0 = nears[n];
        t0 = fmaf(oracle_clampf(t0 * dt_gamma, m.dt_min, m.dt_max), noises[n], t0);
        float t = t0, x, y, z, dt, tt;
        uint32_t num_steps = 0;
        while (t < far && num_steps < max_steps) {
            if (march_probe(&m, t, &x, &y, &z, &dt, &tt)) { num_steps++; t += dt; }
            else do { t += oracle_clampf(t * dt_gamma, m.dt_min, m.dt_max); } while (t < tt);
        }
        const uint32_t point_index = (uint32_t)counter[0];
        counter[0] += (int32_t)num_steps;
        const uint32_t ray_index = (uint32_t)counter[1];
        counter[1] += 1;
        rays[ray_index * 3] = (int32_t)n;
        rays[ray_index * 3 + 1] = (int32_t)point_index;
        rays[ray_index * 3 + 2] = (int32_t)num_steps;
        if (num_steps == 0 || point_index + num_steps > M) continue;
        float *px = xyzs + (size_t)point_index * 3, *pd = dirs + (size_t)point_index * 3, *pl = deltas + (size_t)point_index * 2;
        t = t0;
        uint32_t step = 0;
        float last_t = t;
        while (t < far && step < num_steps) {
            if (march_probe(&m, t, &x, &y, &z, &dt, &tt)) {
                px[0] = x; px[1] = y; px[2] = z;
                pd[0] = m.dx; pd[1] = m.dy; pd[2] = m.dz;
                t += dt;
                pl[0] = dt; pl[1] = t - last_t;
                last_t = t;
                px += 3; pd += 3; pl += 2; step++;
            } else do { t += oracle_clampf(t * dt_gamma, m.dt_min, m.dt_max); } while (t < tt);
        }
    }
}

/* raymarching.cu:500-577.  `fast_exp` != 0 selects exp2f(x*log2e) to mimic __expf's formulation;
 * either way the device intrinsic differs in the last bits, so comparisons use a tolerance. */
void oracle_composite_rays_train_forward(const float *sigmas, const float *rgbs, const float *deltas,
                                         const int32_t *rays, uint32_t M, uint32_t N, float T_thresh,
                                         float *weights_sum, float *depth, float *image) {
    for (uint32_t n = 0; n < N; ++n) {
        const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1], num_steps = (uint32_t)rays[n * 3 + 2];
        if (num_steps == 0 || offset + num_steps > M) {
            weights_sum[index] = 0; depth[index] = 0;
            image[index * 3] = image[index * 3 + 1] = image[index * 3 + 2] = 0;
            continue;
        }
        const float *s = sigmas + offset, *c = rgbs + (size_t)offset * 3, *dl = deltas + (size_t)offset * 2;
        float T = 1.0f, r = 0, g = 0, b = 0, ws = 0, t = 0, d = 0;
        for (uint32_t step = 0; step < num_steps; ++step) {
            const float alpha = 1.0f - expf(-s[0] * dl[0]);
            const float w = alpha * T;
            r = fmaf(w, c[0], r); g = fmaf(w, c[1], g); b = fmaf(w, c[2], b);
            t += dl[1];
            d = fmaf(w, t, d);
            ws += w;
            T *= 1.0f - alpha;
            if (T < T_thresh) break;
            s++; c += 3; dl += 2;
        }
        weights_sum[index] = ws; depth[index] = d;
        image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
    }
}

/* raymarching.cu:601-682 */
void oracle_composite_rays_train_backward(const float *grad_weights_sum, const float *grad_image,
                                          const float *sigmas, const float *rgbs, const float *deltas,
                                          const int32_t *rays, const float *weights_sum, const float *image,
                                          uint32_t M, uint32_t N, float T_thresh, float *grad_sigmas,
                                          float *grad_rgbs) {
    for (uint32_t n = 0; n < N; ++n) {
        const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1], num_steps = (uint32_t)rays[n * 3 + 2];
        if (num_steps == 0 || offset + num_steps > M) continue;
        const float gws = grad_weights_sum[index];
        const float *gi = grad_image + (size_t)index * 3;
        const float rf = image[index * 3], gf = image[index * 3 + 1], bf = image[index * 3 + 2], wsf = weights_sum[index];
        const float *s = sigmas + offset, *c = rgbs + (size_t)offset * 3, *dl = deltas + (size_t)offset * 2;
        float *gs = grad_sigmas + offset, *gc = grad_rgbs + (size_t)offset * 3;
        float T = 1.0f, r = 0, g = 0, b = 0, ws = 0;
        for (uint32_t step = 0; step < num_steps; ++step) {
            const float alpha = 1.0f - expf(-s[0] * dl[0]);
            const float w = alpha * T;
            r = fmaf(w, c[0], r); g = fmaf(w, c[1], g); b = fmaf(w, c[2], b);
            ws += w;
            T *= 1.0f - alpha;
            gc[0] = gi[0] * w; gc[1] = gi[1] * w; gc[2] = gi[2] * w;
            gs[0] = dl[0] * (gi[0] * (T * c[0] - (rf - r)) + gi[1] * (T * c[1] - (gf - g)) +
                             gi[2] * (T * c[2] - (bf - b)) + gws * (1 - wsf));
            if (T < T_thresh) break;
            s++; c += 3; dl += 2; gs++; gc += 3;
        }
    }
}

/* raymarching.cu:700-805 */
void oracle_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t *rays_alive, const float *rays_t,
                       const float *rays_o, const float *rays_d, float bound, float dt_gamma,
                       uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t *grid, const float *nears,
                       const float *fars, float *xyzs, float *dirs, float *deltas, const float *noises) {
    for (uint32_t n = 0; n < n_alive; ++n) {
        const int32_t index = rays_alive[n];
        march_ctx_t m;
        march_ctx_init(&m, rays_o + (size_t)index * 3, rays_d + (size_t)index * 3, grid, bound, dt_gamma, max_steps, C, H);
        float *px = xyzs + (size_t)n * n_step * 3, *pd = dirs + (size_t)n * n_step * 3, *pl = deltas + (size_t)n * n_step * 2;
        float t = rays_t[index];
        const float far = fars[index];
        (void)nears;
        t = fmaf(oracle_clampf(t * dt_gamma, m.dt_min, m.dt_max), noises[n], t);
        float last_t = t, x, y, z, dt, tt;
        uint32_t step = 0;
        while (t < far && step < n_step) {
            if (march_probe(&m, t, &x, &y, &z, &dt, &tt)) {
                px[0] = x; px[1] = y; px[2] = z;
                pd[0] = m.dx; pd[1] = m.dy; pd[2] = m.dz;
                t += dt;
                pl[0] = dt; pl[1] = t - last_t;
                last_t = t;
                px += 3; pd += 3; pl += 2; step++;
            } else do { t += oracle_clampf(t * dt_gamma, m.dt_min, m.dt_max); } while (t < tt);
        }
    }
}

/* raymarching.cu:818-905 */
void oracle_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t *rays_alive,
                           float *rays_t, const float *sigmas, const float *rgbs, const float *deltas,
                           float *weights_sum, float *depth, float *image) {
    for (uint32_t n = 0; n < n_alive; ++n) {
        const int32_t index = rays_alive[n];
        const float *s = sigmas + (size_t)n * n_step, *c = rgbs + (size_t)n * n_step * 3, *dl = deltas + (size_t)n * n_step * 2;
        float t = rays_t[index], ws = weights_sum[index], d = depth[index];
        float r = image[index * 3], g = image[index * 3 + 1], b = image[index * 3 + 2];
        uint32_t step = 0;
        while (step < n_step) {
            if (dl[0] == 0) break;
            const float alpha = 1.0f - expf(-s[0] * dl[0]);
            const float T = 1 - ws;
            const float w = alpha * T;
            ws += w;
            t += dl[1];
            d = fmaf(w, t, d);
            r = fmaf(w, c[0], r); g = fmaf(w, c[1], g); b = fmaf(w, c[2], b);
            if (T < T_thresh) break;
            s++; c += 3; dl += 2; step++;
        }
        if (step < n_step) rays_alive[n] = -1; else rays_t[index] = t;
        weights_sum[index] = ws; depth[index] = d;
        image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
    }
}
