/*
 * oracle/raymarch_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, IEEE fp32, no FMA contraction) of the reference's
 * ray-march / compositing / bit-packing kernels.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * Parity status: the reference ships NO golden vectors for this path;
 * this restatement is pinned (a) by running the compiled reference .cu
 * (oracle/_ref, built by oracle/build_ref.py) against it on the GPU box, and
 * (b) by fixtures produced with the reference's own Python (tests/golden/make_golden.py).
 *
 * Each function cites the reference lines (relative to /root/reference) it follows.
 * One deliberate difference: sample compaction is ordered by ray id (exclusive prefix sum
 * of the per-ray counts) instead of by atomicAdd arrival order
 * (raymarching/src/raymarching.cu:405-406), which is nondeterministic in the reference.
 * Per-ray contents are identical.
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -fopenmp raymarch_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <float.h>

static inline float clampf(float v, float lo, float hi) { return fminf(hi, fmaxf(lo, v)); }
static inline float sgn1(float v) { return copysignf(1.0f, v); }

/* raymarching.cu:56-71 : 10-bit-per-axis Morton interleave */
static inline uint32_t spread3(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
static inline uint32_t morton_enc(uint32_t x, uint32_t y, uint32_t z) {
    return spread3(x) | (spread3(y) << 1) | (spread3(z) << 2);
}
/* raymarching.cu:73-81 */
static inline uint32_t compact3(uint32_t x) {
    x &= 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

/* raymarching.cu:42-54 : cascade level from position / step size (frexpf exponent, clamped) */
static inline int level_from_pos(float x, float y, float z, int C) {
    float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    int e; frexpf(mx, &e);
    return (int)fminf((float)(C - 1), fmaxf(0.0f, (float)e));
}
static inline int level_from_dt(float dt, float H, int C) {
    float mx = dt * H * 0.5f;
    int e; frexpf(mx, &e);
    return (int)fminf((float)(C - 1), fmaxf(0.0f, (float)e));
}

/* raymarching.cu:92-145 : slab test against aabb, near clamped to min_near, miss => FLT_MAX */
void orc_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb,
                            uint32_t N, float min_near, float* nears, float* fars) {
    #pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < (int64_t)N; n++) {
        const float* o = rays_o + 3 * n; const float* d = rays_d + 3 * n;
        float r0 = 1 / d[0], r1 = 1 / d[1], r2 = 1 / d[2];
        float tn = (aabb[0] - o[0]) * r0, tf = (aabb[3] - o[0]) * r0;
        if (tn > tf) { float s = tn; tn = tf; tf = s; }
        float yn = (aabb[1] - o[1]) * r1, yf = (aabb[4] - o[1]) * r1;
        if (yn > yf) { float s = yn; yn = yf; yf = s; }
        if (tn > yf || yn > tf) { nears[n] = fars[n] = FLT_MAX; continue; }
        if (yn > tn) tn = yn;
        if (yf < tf) tf = yf;
        float zn = (aabb[2] - o[2]) * r2, zf = (aabb[5] - o[2]) * r2;
        if (zn > zf) { float s = zn; zn = zf; zf = s; }
        if (tn > zf || zn > tf) { nears[n] = fars[n] = FLT_MAX; continue; }
        if (zn > tn) tn = zn;
        if (zf < tf) tf = zf;
        if (tn < min_near) tn = min_near;
        nears[n] = tn; fars[n] = tf;
    }
}

/* raymarching.cu:214-226 / :237-254 */
void orc_morton3D(const int* coords, uint32_t N, int* indices) {
    for (uint32_t n = 0; n < N; n++)
        indices[n] = (int)morton_enc((uint32_t)coords[3*n], (uint32_t)coords[3*n+1], (uint32_t)coords[3*n+2]);
}
void orc_morton3D_invert(const int* indices, uint32_t N, int* coords) {
    for (uint32_t n = 0; n < N; n++) {
        uint32_t v = (uint32_t)indices[n];
        coords[3*n] = (int)compact3(v); coords[3*n+1] = (int)compact3(v >> 1); coords[3*n+2] = (int)compact3(v >> 2);
    }
}

/* raymarching.cu:268-289 : bit i of byte n set iff grid[8n+i] > thresh */
void orc_packbits(const float* grid, uint32_t Nbytes, float thresh, uint8_t* bits) {
    for (uint32_t n = 0; n < Nbytes; n++) {
        uint8_t b = 0;
        for (int i = 0; i < 8; i++) if (grid[8*(size_t)n + i] > thresh) b |= (uint8_t)(1u << i);
        bits[n] = b;
    }
}

/* Shared marching state machine for raymarching.cu:351-399 (count pass), :422-475 (emit pass)
 * and :955-1010 (inference).  Walks one ray from t until `far` or until `max_emit` occupied
 * samples were produced; if out pointers are non-NULL the samples are written.
 * Returns the number of occupied samples; *t_io holds the final t. */
typedef struct {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz;
    float bound, dt_gamma, dt_min, dt_max, rH, H3f;
    uint32_t C, H;
    const uint8_t* grid;
} ray_ctx;

static uint32_t walk(const ray_ctx* r, float* t_io, float far, uint32_t max_emit,
                     float* xyzs, float* dirs, float* deltas) {
    float t = *t_io, last_t = t;
    uint32_t k = 0;
    const float Hf = (float)r->H;
    while (t < far && k < max_emit) {
        const float x = clampf(r->ox + t * r->dx, -r->bound, r->bound);
        const float y = clampf(r->oy + t * r->dy, -r->bound, r->bound);
        const float z = clampf(r->oz + t * r->dz, -r->bound, r->bound);
        const float dt = clampf(t * r->dt_gamma, r->dt_min, r->dt_max);
        int la = level_from_pos(x, y, z, (int)r->C), lb = level_from_dt(dt, Hf, (int)r->C);
        const int level = la > lb ? la : lb;
        const float mip_bound = fminf(scalbnf(1.0f, level), r->bound);
        const float mip_rbound = 1 / mip_bound;
        /* reference evaluates 0.5 * (..) * H in double then narrows; H<=2^10 and the 0.5 make
         * that a single rounding, identical to the fp32 evaluation below */
        const int nx = (int)clampf(0.5f * (x * mip_rbound + 1) * Hf, 0.0f, (float)(r->H - 1));
        const int ny = (int)clampf(0.5f * (y * mip_rbound + 1) * Hf, 0.0f, (float)(r->H - 1));
        const int nz = (int)clampf(0.5f * (z * mip_rbound + 1) * Hf, 0.0f, (float)(r->H - 1));
        /* level * H3 is a float product in the reference (H3 is float), then converted to uint32 */
        const uint32_t index = (uint32_t)((float)level * r->H3f + (float)morton_enc((uint32_t)nx, (uint32_t)ny, (uint32_t)nz));
        const int occ = r->grid[index >> 3] & (1u << (index & 7u));
        if (occ) {
            if (xyzs) {
                xyzs[3*k] = x; xyzs[3*k+1] = y; xyzs[3*k+2] = z;
                dirs[3*k] = r->dx; dirs[3*k+1] = r->dy; dirs[3*k+2] = r->dz;
            }
            t += dt;
            if (deltas) { deltas[2*k] = dt; deltas[2*k+1] = t - last_t; }
            last_t = t;
            k++;
        } else {
            const float tx = (((nx + 0.5f + 0.5f * sgn1(r->dx)) * r->rH * 2 - 1) * mip_bound - x) * r->rdx;
            const float ty = (((ny + 0.5f + 0.5f * sgn1(r->dy)) * r->rH * 2 - 1) * mip_bound - y) * r->rdy;
            const float tz = (((nz + 0.5f + 0.5f * sgn1(r->dz)) * r->rH * 2 - 1) * mip_bound - z) * r->rdz;
            const float tt = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
            do { t += clampf(t * r->dt_gamma, r->dt_min, r->dt_max); } while (t < tt);
        }
    }
    *t_io = t;
    return k;
}

static void ray_setup(ray_ctx* r, const float* o, const float* d, float bound, float dt_gamma,
                      uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* grid) {
    r->ox = o[0]; r->oy = o[1]; r->oz = o[2];
    r->dx = d[0]; r->dy = d[1]; r->dz = d[2];
    r->rdx = 1 / d[0]; r->rdy = 1 / d[1]; r->rdz = 1 / d[2];
    r->bound = bound; r->dt_gamma = dt_gamma;
    r->dt_min = 2 * 1.7320508075688772f / (float)max_steps;           /* raymarching.cu:345 */
    r->dt_max = 2 * 1.7320508075688772f * (float)(1u << (C - 1)) / (float)H;  /* :346 */
    r->rH = 1 / (float)H; r->H3f = (float)H * (float)H * (float)H;
    r->C = C; r->H = H; r->grid = grid;
}

/* raymarching.cu:312-480 with ray-id-ordered compaction.
 * rays[n] = (n, offset, count); counter[0] += total samples, counter[1] += N.
 * Rays whose samples would overflow M keep their rays[] entry but write nothing (:416). */
void orc_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid,
                          float bound, float dt_gamma, uint32_t max_steps,
                          uint32_t N, uint32_t C, uint32_t H, uint32_t M,
                          const float* nears, const float* fars,
                          float* xyzs, float* dirs, float* deltas, int* rays, int* counter,
                          const float* noises) {
    #pragma omp parallel for schedule(dynamic, 64)
    for (int64_t n = 0; n < (int64_t)N; n++) {
        ray_ctx r; ray_setup(&r, rays_o + 3*n, rays_d + 3*n, bound, dt_gamma, max_steps, C, H, grid);
        float t0 = nears[n];
        t0 += clampf(t0 * dt_gamma, r.dt_min, r.dt_max) * noises[n];   /* :351 */
        float t = t0;
        rays[3*n + 2] = (int)walk(&r, &t, fars[n], max_steps, 0, 0, 0);
    }
    uint32_t run = (uint32_t)counter[0];
    for (uint32_t n = 0; n < N; n++) { rays[3*n] = (int)n; rays[3*n+1] = (int)run; run += (uint32_t)rays[3*n+2]; }
    counter[0] = (int)run; counter[1] += (int)N;
    #pragma omp parallel for schedule(dynamic, 64)
    for (int64_t n = 0; n < (int64_t)N; n++) {
        uint32_t off = (uint32_t)rays[3*n+1], cnt = (uint32_t)rays[3*n+2];
        if (cnt == 0 || off + cnt > M) continue;
        ray_ctx r; ray_setup(&r, rays_o + 3*n, rays_d + 3*n, bound, dt_gamma, max_steps, C, H, grid);
        float t0 = nears[n];
        t0 += clampf(t0 * dt_gamma, r.dt_min, r.dt_max) * noises[n];
        float t = t0;
        walk(&r, &t, fars[n], cnt, xyzs + 3*(size_t)off, dirs + 3*(size_t)off, deltas + 2*(size_t)off);
    }
}

/* raymarching.cu:501-577.  use_fast_exp mirrors __expf (ex2.approx(x*log2e)); on the CPU we can
 * only offer expf, parity tests bound the difference. */
void orc_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* deltas,
                                      const int* rays, uint32_t M, uint32_t N, float T_thresh,
                                      float* weights_sum, float* depth, float* image) {
    #pragma omp parallel for schedule(dynamic, 64)
    for (int64_t n = 0; n < (int64_t)N; n++) {
        uint32_t index = (uint32_t)rays[3*n], off = (uint32_t)rays[3*n+1], cnt = (uint32_t)rays[3*n+2];
        float T = 1.0f, r = 0, g = 0, b = 0, ws = 0, t = 0, d = 0;
        if (cnt != 0 && off + cnt <= M) {
            for (uint32_t s = 0; s < cnt; s++) {
                const size_t i = (size_t)off + s;
                const float alpha = 1.0f - expf(-sigmas[i] * deltas[2*i]);
                const float w = alpha * T;
                r += w * rgbs[3*i]; g += w * rgbs[3*i+1]; b += w * rgbs[3*i+2];
                t += deltas[2*i+1];
                d += w * t;
                ws += w;
                T *= 1.0f - alpha;
                if (T < T_thresh) break;
            }
        }
        weights_sum[index] = ws; depth[index] = d;
        image[3*index] = r; image[3*index+1] = g; image[3*index+2] = b;
    }
}

/* raymarching.cu:602-682.  grad_sigmas / grad_rgbs must be pre-zeroed by the caller
 * (raymarching.py:295-296); samples after the early-out keep zero gradient. */
void orc_composite_rays_train_backward(const float* grad_ws, const float* grad_image,
                                       const float* sigmas, const float* rgbs, const float* deltas,
                                       const int* rays, const float* weights_sum, const float* image,
                                       uint32_t M, uint32_t N, float T_thresh,
                                       float* grad_sigmas, float* grad_rgbs) {
    #pragma omp parallel for schedule(dynamic, 64)
    for (int64_t n = 0; n < (int64_t)N; n++) {
        uint32_t index = (uint32_t)rays[3*n], off = (uint32_t)rays[3*n+1], cnt = (uint32_t)rays[3*n+2];
        if (cnt == 0 || off + cnt > M) continue;
        const float* gi = grad_image + 3*(size_t)index;
        const float gw = grad_ws[index];
        const float rf = image[3*index], gf = image[3*index+1], bf = image[3*index+2], wsf = weights_sum[index];
        float T = 1.0f, r = 0, g = 0, b = 0, ws = 0;
        for (uint32_t s = 0; s < cnt; s++) {
            const size_t i = (size_t)off + s;
            const float alpha = 1.0f - expf(-sigmas[i] * deltas[2*i]);
            const float w = alpha * T;
            r += w * rgbs[3*i]; g += w * rgbs[3*i+1]; b += w * rgbs[3*i+2];
            ws += w;
            T *= 1.0f - alpha;
            grad_rgbs[3*i] = gi[0] * w; grad_rgbs[3*i+1] = gi[1] * w; grad_rgbs[3*i+2] = gi[2] * w;
            grad_sigmas[i] = deltas[2*i] * (
                gi[0] * (T * rgbs[3*i]   - (rf - r)) +
                gi[1] * (T * rgbs[3*i+1] - (gf - g)) +
                gi[2] * (T * rgbs[3*i+2] - (bf - b)) +
                gw * (1 - wsf));
            if (T < T_thresh) break;
        }
    }
}

/* raymarching.cu:907-1011 : inference march, n_step slots per alive ray (unused slots stay zero). */
void orc_march_rays(uint32_t n_alive, uint32_t n_step, const int* rays_alive, const float* rays_t,
                    const float* rays_o, const float* rays_d, float bound, float dt_gamma,
                    uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* grid,
                    const float* nears, const float* fars,
                    float* xyzs, float* dirs, float* deltas, const float* noises) {
    (void)nears;
    #pragma omp parallel for schedule(dynamic, 64)
    for (int64_t n = 0; n < (int64_t)n_alive; n++) {
        const int id = rays_alive[n];
        ray_ctx r; ray_setup(&r, rays_o + 3*(size_t)id, rays_d + 3*(size_t)id, bound, dt_gamma, max_steps, C, H, grid);
        float t = rays_t[id];
        t += clampf(t * dt_gamma, r.dt_min, r.dt_max) * noises[n];     /* :955 */
        size_t base = (size_t)n * n_step;
        walk(&r, &t, fars[id], n_step, xyzs + 3*base, dirs + 3*base, deltas + 2*base);
    }
}

/* raymarching.cu:1024-1115 : in-place accumulation; T = 1 - weight_sum; dead rays marked -1. */
void orc_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int* rays_alive, float* rays_t,
                        const float* sigmas, const float* rgbs, const float* normals, const float* deltas,
                        float* weights_sum, float* depth, float* image, float* normal) {
    #pragma omp parallel for schedule(dynamic, 64)
    for (int64_t n = 0; n < (int64_t)n_alive; n++) {
        const int id = rays_alive[n];
        size_t base = (size_t)n * n_step;
        float t = rays_t[id], d = depth[id], ws = weights_sum[id];
        float r = image[3*id], g = image[3*id+1], b = image[3*id+2];
        float nx = normal[3*id], ny = normal[3*id+1], nz = normal[3*id+2];
        uint32_t s = 0;
        while (s < n_step) {
            const size_t i = base + s;
            if (deltas[2*i] == 0) break;
            const float alpha = 1.0f - expf(-sigmas[i] * deltas[2*i]);
            const float T = 1 - ws;
            const float w = alpha * T;
            ws += w;
            t += deltas[2*i+1];
            d += w * t;
            r += w * rgbs[3*i]; g += w * rgbs[3*i+1]; b += w * rgbs[3*i+2];
            nx += w * normals[3*i]; ny += w * normals[3*i+1]; nz += w * normals[3*i+2];
            if (T < T_thresh) break;
            s++;
        }
        if (s < n_step) rays_alive[n] = -1; else rays_t[id] = t;
        weights_sum[id] = ws; depth[id] = d;
        image[3*id] = r; image[3*id+1] = g; image[3*id+2] = b;
        normal[3*id] = nx; normal[3*id+1] = ny; normal[3*id+2] = nz;
    }
}

/* ------------------------------------------------------------------------------------------
 * tiny-cuda-nn HashGrid forward (third-party, NOT in /root/reference; un-pinned, README.md:43).
 * Restated from the published algorithm (tiny-cuda-nn include/tiny-cuda-nn/encodings/grid.h,
 * common_device.h: grid_scale / grid_resolution / grid_index / lcg primes) as called at
 * nerf/network_tcnn.py:54-65,107.  PARITY UNPINNED for this function (no tcnn install here).
 * x in [0,1]^3, table fp32 level-major, 2 features per entry, out [E, 2L].
 * ------------------------------------------------------------------------------------------ */
typedef struct { uint32_t offset, size, res; float scale; } hg_level;

/* Fills `lv[L]`, returns total entries. */
uint32_t orc_hashgrid_levels(uint32_t L, uint32_t base_res, double per_level_scale, uint32_t log2_T,
                             uint32_t* offsets, uint32_t* sizes, uint32_t* ress, float* scales) {
    uint32_t off = 0;
    for (uint32_t l = 0; l < L; l++) {
        /* grid_scale(): exp2f(level * log2_per_level_scale) * base_resolution - 1 */
        float scale = exp2f((float)l * log2f((float)per_level_scale)) * (float)base_res - 1.0f;
        uint32_t res = (uint32_t)ceilf(scale) + 1;
        uint64_t dense = (uint64_t)res * res * res;
        uint64_t cap = (uint64_t)1 << log2_T;
        uint32_t size = (uint32_t)(dense < cap ? dense : cap);
        size = (size + 7u) / 8u * 8u;             /* next_multiple(params_in_level, 8) */
        if (size > cap) size = (uint32_t)cap;
        offsets[l] = off; sizes[l] = size; ress[l] = res; scales[l] = scale;
        off += size;
    }
    return off;
}

static inline uint32_t hg_index(uint32_t cx, uint32_t cy, uint32_t cz, uint32_t res, uint32_t size) {
    /* grid_index(): dense stride walk while stride <= size, else the 3-prime xor hash; then % size */
    uint32_t stride = 1, idx = 0;
    uint32_t c[3] = { cx, cy, cz };
    for (int d = 0; d < 3 && stride <= size; d++) { idx += c[d] * stride; stride *= res; }
    if (size < stride) idx = cx ^ (cy * 2654435761u) ^ (cz * 805459861u);
    return idx % size;
}

void orc_hashgrid_forward(const float* x, uint32_t E, const float* table, uint32_t L,
                          const uint32_t* offsets, const uint32_t* sizes, const uint32_t* ress,
                          const float* scales, float* out) {
    #pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < (int64_t)E; e++) {
        for (uint32_t l = 0; l < L; l++) {
            float pos[3], w[3]; uint32_t cell[3];
            for (int d = 0; d < 3; d++) {
                float p = fmaf(scales[l], x[3*e + d], 0.5f);
                float f = floorf(p);
                cell[d] = (uint32_t)(int32_t)f; w[d] = p - f; pos[d] = p;
            }
            (void)pos;
            float f0 = 0, f1 = 0;
            for (uint32_t corner = 0; corner < 8; corner++) {
                float wt = 1; uint32_t c[3];
                for (int d = 0; d < 3; d++) {
                    if (corner & (1u << d)) { wt *= w[d]; c[d] = cell[d] + 1; }
                    else { wt *= 1 - w[d]; c[d] = cell[d]; }
                }
                uint32_t idx = hg_index(c[0], c[1], c[2], ress[l], sizes[l]);
                const float* ent = table + 2 * ((size_t)offsets[l] + idx);
                f0 += wt * ent[0]; f1 += wt * ent[1];
            }
            out[(size_t)e * 2 * L + 2*l] = f0; out[(size_t)e * 2 * L + 2*l + 1] = f1;
        }
    }
}
