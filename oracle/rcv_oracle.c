/*
 * rcv_oracle.c — CPU restatement of the recursive-control-variates hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (the package
 * recursive_control_variates_b200/ or librcv_b200.so) may import, call or link
 * this file; only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
 * reference legs use it, as the checker.
 *
 * PARITY UNPINNED: the reference (rgl-epfl/recursive_control_variates) is pure
 * Python on top of Mitsuba 3 (branch unbiased-volume-opt) + Dr.Jit, neither of
 * which is installable here, and it ships no test, golden vector or fixture.
 * This file follows the reference sources statement by statement (citations
 * `file:line` are relative to the reference repository) and restates the
 * upstream Mitsuba/Dr.Jit conventions it relies on from their published
 * behaviour (marked [UPSTREAM]); DESIGN.md §2 lists every such convention.
 * The only external known-answer test is the public PCG32 demo vector.
 *
 * Structure: each vectorised Dr.Jit statement of the reference becomes one
 * scalar statement per lane with the masks turned into ints, in the same
 * order, so that the random-number stream of every lane is consumed exactly
 * as in the reference (masked draws only advance active lanes).
 *
 * Floating point: everything is IEEE binary32 with a fixed evaluation order
 * (compile with -ffp-contract=off; fused multiply-adds are written as fmaf).
 * log and sincos are the Cephes single-precision polynomials (what Dr.Jit's
 * LLVM backend evaluates), spelled out below, so results are reproducible
 * bit for bit on any IEEE machine — the CUDA kernels implement the same
 * arithmetic and are compared lane by lane against this file.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/rcv_b200.h"
#include "rcv_oracle.h"

/* ------------------------------------------------------------------------- */
/* small vector helpers                                                      */
/* ------------------------------------------------------------------------- */
typedef struct { float x, y, z; } v3;

static inline v3 V3(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline float fminf_(float a, float b) { return (b < a) ? b : a; }
static inline float fmaxf_(float a, float b) { return (b > a) ? b : a; }
static inline float max3(v3 a) { return fmaxf_(fmaxf_(a.x, a.y), a.z); }
static inline int any_nz(v3 a) { return a.x != 0.f || a.y != 0.f || a.z != 0.f; }
static inline v3 vmul(v3 a, v3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 vscale(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline v3 vadd(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vmax(v3 a, v3 b) { return V3(fmaxf_(a.x, b.x), fmaxf_(a.y, b.y), fmaxf_(a.z, b.z)); }
static inline v3 vdiv(v3 a, v3 b) { return V3(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline float get3(const float *a, int i) { return a[i]; }
/* ray(t) = fmadd(d, t, o) [UPSTREAM Ray::operator()] */
static inline v3 ray_at(v3 o, v3 d, float t) {
    return V3(fmaf(d.x, t, o.x), fmaf(d.y, t, o.y), fmaf(d.z, t, o.z));
}

#define ORC_INF (__builtin_inff())
#define ORC_LARGEST 3.402823466e+38f               /* dr.largest(Float) */
#define ORC_RAY_EPS (1500.f * 5.9604644775390625e-8f) /* [UPSTREAM] RayEpsilon = 1500 * 2^-24 */
#define ORC_SHADOW_EPS (10.f * ORC_RAY_EPS)           /* [UPSTREAM] ShadowEpsilon */
#define ORC_INV_FOURPI 0.07957747154594766788f
#define ORC_TWO_PI 6.28318530717958647692f

/* ------------------------------------------------------------------------- */
/* RNG: TEA-4 scrambling + PCG32  [UPSTREAM independent sampler / drjit PCG32] */
/* ------------------------------------------------------------------------- */
#define PCG32_MULT 0x5851f42d4c957f2dULL

void orc_tea32(uint32_t v0, uint32_t v1, uint32_t *o0, uint32_t *o1) {
    /* mi.sample_tea_32(v0, v1, rounds=4) (volpathsimple.py:118, twostatepath.py:89) */
    uint32_t sum = 0;
    for (int i = 0; i < 4; ++i) {
        sum += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + sum) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + sum) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    *o0 = v0; *o1 = v1;
}

typedef struct { uint64_t state, inc; } pcg32;

static inline uint32_t pcg_next_u32(pcg32 *r) {
    uint64_t old = r->state;
    r->state = old * PCG32_MULT + r->inc;
    uint32_t xs = (uint32_t) (((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t) (old >> 59u);
    return (xs >> rot) | (xs << ((-rot) & 31));
}
static inline void pcg_seed(pcg32 *r, uint64_t initstate, uint64_t initseq) {
    r->state = 0;
    r->inc = (initseq << 1u) | 1u;
    pcg_next_u32(r);
    r->state += initstate;
    pcg_next_u32(r);
}
static inline float u32_to_f32(uint32_t v) {
    union { uint32_t u; float f; } x;
    x.u = (v >> 9) | 0x3f800000u;
    return x.f - 1.f;
}
/* sampler.next_1d(mask): masked lanes neither advance nor receive a usable
 * value; the returned value for a masked lane is what Dr.Jit would compute
 * from the current state (only needed for volpathsimple.py:117, see
 * orc_alt_seed). */
static inline float next_1d(pcg32 *r, int active) {
    if (active) return u32_to_f32(pcg_next_u32(r));
    pcg32 c = *r;
    return u32_to_f32(pcg_next_u32(&c));
}
/* Sampler.seed(seed, wavefront) for lane idx [UPSTREAM independent.cpp] */
static inline void sampler_seed(pcg32 *r, uint32_t seed, uint32_t idx) {
    uint32_t v0, v1;
    orc_tea32(seed, idx, &v0, &v1);
    pcg_seed(r, (uint64_t) v0, (uint64_t) v1);
}

void orc_pcg32_kat(uint64_t initstate, uint64_t initseq, uint32_t *out, int n) {
    pcg32 r; pcg_seed(&r, initstate, initseq);
    for (int i = 0; i < n; ++i) out[i] = pcg_next_u32(&r);
}
float orc_u32_to_f32(uint32_t v) { return u32_to_f32(v); }

/* ------------------------------------------------------------------------- */
/* elementary functions (Cephes single precision, fixed operation order)     */
/* ------------------------------------------------------------------------- */
float orc_logf(float x) {
    /* x is a positive normal number (1-u with u in [0,1) => [2^-24, 1]) */
    union { float f; uint32_t u; } b; b.f = x;
    int e = (int) (b.u >> 23) - 126;
    b.u = (b.u & 0x007fffffu) | 0x3f000000u;   /* mantissa in [0.5, 1) */
    float m = b.f;
    if (m < 0.707106781186547524f) { e -= 1; m = (m + m) - 1.f; }
    else                           { m = m - 1.f; }
    float z = m * m;
    float y = 7.0376836292e-2f;
    y = fmaf(y, m, -1.1514610310e-1f);
    y = fmaf(y, m, 1.1676998740e-1f);
    y = fmaf(y, m, -1.2420140846e-1f);
    y = fmaf(y, m, 1.4249322787e-1f);
    y = fmaf(y, m, -1.6668057665e-1f);
    y = fmaf(y, m, 2.0000714765e-1f);
    y = fmaf(y, m, -2.4999993993e-1f);
    y = fmaf(y, m, 3.3333331174e-1f);
    y = (y * m) * z;
    float fe = (float) e;
    y = fmaf(fe, -2.12194440e-4f, y);
    y = fmaf(-0.5f, z, y);
    float r = m + y;
    r = fmaf(fe, 0.693359375f, r);
    return r;
}

/* Cephes expf for x <= 0 of moderate size (reconstruction-filter weights), fixed operation order */
float orc_expf(float x) {
    float z = floorf(fmaf(1.44269504088896341f, x, 0.5f));      /* n = round(x / ln 2) */
    x = fmaf(z, -0.693359375f, x);
    x = fmaf(z, 2.12194440e-4f, x);
    int n = (int) z;
    float xx = x * x;
    float y = 1.9875691500e-4f;
    y = fmaf(y, x, 1.3981999507e-3f);
    y = fmaf(y, x, 8.3334519073e-3f);
    y = fmaf(y, x, 4.1665795894e-2f);
    y = fmaf(y, x, 1.6666665459e-1f);
    y = fmaf(y, x, 5.0000001201e-1f);
    y = fmaf(y, xx, x);
    y = y + 1.f;
    if (n < -126) return 0.f;
    union { float f; uint32_t u; } s2; s2.u = (uint32_t) (n + 127) << 23;   /* 2^n */
    return y * s2.f;
}

void orc_sincosf(float x, float *s_out, float *c_out) {
    /* x >= 0 (x = 2*pi*u).  Cephes sinf/cosf with 3-term Cody-Waite reduction. */
    int j = (int) (x * 1.27323954473516f);   /* x * 4/pi */
    j = j + (j & 1);
    float y = (float) j;
    float r = fmaf(y, -0.78515625f, x);
    r = fmaf(y, -2.4187564849853515625e-4f, r);
    r = fmaf(y, -3.77489497744594108e-8f, r);
    float z = r * r;
    float ps = -1.9515295891e-4f;
    ps = fmaf(ps, z, 8.3321608736e-3f);
    ps = fmaf(ps, z, -1.6666654611e-1f);
    ps = fmaf(ps * z, r, r);
    float pc = 2.443315711809948e-5f;
    pc = fmaf(pc, z, -1.388731625493765e-3f);
    pc = fmaf(pc, z, 4.166664568298827e-2f);
    pc = fmaf(pc * z, z, fmaf(-0.5f, z, 1.f));
    int q = (j >> 1) & 3;                    /* quadrant */
    float s = (q & 1) ? pc : ps;
    float c = (q & 1) ? ps : pc;
    if (q & 2) s = -s;
    if ((q + 1) & 2) c = -c;
    *s_out = s; *c_out = c;
}

/* warp.square_to_uniform_sphere [UPSTREAM warp.h] */
static inline v3 square_to_uniform_sphere(float sx, float sy) {
    float z = fmaf(-2.f, sy, 1.f);
    float r2 = fmaf(-z, z, 1.f);
    float r = sqrtf(fmaxf_(r2, 0.f));
    float s, c;
    orc_sincosf(ORC_TWO_PI * sx, &s, &c);
    return V3(r * c, r * s, z);
}

/* mi.ad.common.mis_weight [UPSTREAM]: power heuristic, non-finite -> 0 */
static inline float mis_weight(float pdf_a, float pdf_b) {
    float a2 = pdf_a * pdf_a, b2 = pdf_b * pdf_b;
    float w = a2 / (a2 + b2);
    return isfinite(w) ? w : 0.f;
}

/* Cephes atanf / atan2f, fixed operation order (envmap direction -> (u, v)) */
#define ORC_PI 3.14159265358979323846f
#define ORC_PIO2 1.57079632679489661923f
#define ORC_PIO4 0.78539816339744830962f
static inline float spec_atanf(float xx) {
    float x = fabsf(xx), y;
    if (x > 2.414213562373095f) { y = ORC_PIO2; x = -(1.f / x); }
    else if (x > 0.4142135623730950f) { y = ORC_PIO4; x = (x - 1.f) / (x + 1.f); }
    else y = 0.f;
    float z = x * x;
    float p = 8.05374449538e-2f;
    p = fmaf(p, z, -1.38776856032e-1f);
    p = fmaf(p, z, 1.99777106478e-1f);
    p = fmaf(p, z, -3.33329491539e-1f);
    y = y + fmaf(p * z, x, x);
    return (xx < 0.f) ? -y : y;
}
float orc_atan2f(float y, float x) {
    if (x == 0.f) return (y > 0.f) ? ORC_PIO2 : ((y < 0.f) ? -ORC_PIO2 : 0.f);
    float z = spec_atanf(y / x);
    if (x < 0.f) z = (y < 0.f) ? z - ORC_PI : z + ORC_PI;
    return z;
}

/* ------------------------------------------------------------------------- */
/* `envmap` emitter [UPSTREAM envmap.cpp, restated; DESIGN.md §2]: lat-long   */
/* texels at the nodes of a (H-1) x W cell grid (longitude wraps), bilinear    */
/* radiance, piecewise-constant sampling density proportional to the cell      */
/* mean of luminance * sin(theta).  Tables are built in float64, sequentially, */
/* and rounded once (the library builds the same tables independently).        */
/* ------------------------------------------------------------------------- */
typedef struct {
    const float *key; int w, h; float scale; uint64_t checksum;
    float *data;      /* (h, w + 1, 3) scaled texels, column w = column 0 */
    float *cdf_m;     /* h: cdf over the h - 1 cell rows */
    float *cdf_c;     /* (h - 1, w + 1) */
    float *pdf;       /* (h - 1, w): density over [0,1]^2 of (u, v) */
} env_tables;
static env_tables g_env = {0};

static void quads_build(const rcv_scene_desc *S);
static void env_build(const rcv_scene_desc *S) {
    quads_build(S);
    if (S->emitter_type != RCV_EMITTER_ENVMAP) return;
    /* the tables are cached across calls; the key includes a checksum of the texels (a freed bitmap's address may be reused) */
    uint64_t sum = 1469598103934665603ULL;
    {
        const uint32_t *tx = (const uint32_t *) S->envmap_rgb;
        size_t n = (size_t) 3 * S->env_width * S->env_height;
        for (size_t i = 0; i < n; ++i) sum = (sum ^ tx[i]) * 1099511628211ULL;
    }
    if (g_env.key == S->envmap_rgb && g_env.w == S->env_width && g_env.h == S->env_height && g_env.scale == S->env_scale &&
        g_env.checksum == sum) return;
    g_env.checksum = sum;
    free(g_env.data); free(g_env.cdf_m); free(g_env.cdf_c); free(g_env.pdf);
    int w = S->env_width, h = S->env_height;
    g_env.key = S->envmap_rgb; g_env.w = w; g_env.h = h; g_env.scale = S->env_scale;
    g_env.data = (float *) malloc(sizeof(float) * 3 * (size_t) h * (w + 1));
    g_env.cdf_m = (float *) malloc(sizeof(float) * h);
    g_env.cdf_c = (float *) malloc(sizeof(float) * (size_t) (h - 1) * (w + 1));
    g_env.pdf = (float *) malloc(sizeof(float) * (size_t) (h - 1) * w);
    double *g = (double *) malloc(sizeof(double) * (size_t) h * (w + 1));
    double *rowsum = (double *) malloc(sizeof(double) * (h - 1));
    for (int r = 0; r < h; ++r) {
        double st = sin(3.14159265358979323846 * (double) r / (double) (h - 1));
        for (int c = 0; c <= w; ++c) {
            const float *src = S->envmap_rgb + 3 * ((size_t) r * w + (c % w));
            float *dst = g_env.data + 3 * ((size_t) r * (w + 1) + c);
            for (int k = 0; k < 3; ++k) dst[k] = src[k] * S->env_scale;
            g[(size_t) r * (w + 1) + c] = (0.2126 * (double) dst[0] + 0.7152 * (double) dst[1] + 0.0722 * (double) dst[2]) * st;
        }
    }
    double total = 0.0;
    for (int r = 0; r < h - 1; ++r) {
        double s = 0.0;
        for (int c = 0; c < w; ++c) {
            size_t i = (size_t) r * (w + 1) + c;
            s += 0.25 * (g[i] + g[i + 1] + g[i + w + 1] + g[i + w + 2]);
        }
        rowsum[r] = s; total += s;
    }
    double acc = 0.0;
    g_env.cdf_m[0] = 0.f;
    for (int r = 0; r < h - 1; ++r) {
        acc += rowsum[r];
        g_env.cdf_m[r + 1] = (total > 0.0) ? (float) (acc / total) : (float) (r + 1) / (float) (h - 1);
        float *row = g_env.cdf_c + (size_t) r * (w + 1);
        double a2 = 0.0;
        row[0] = 0.f;
        for (int c = 0; c < w; ++c) {
            size_t i = (size_t) r * (w + 1) + c;
            double f = 0.25 * (g[i] + g[i + 1] + g[i + w + 1] + g[i + w + 2]);
            a2 += f;
            row[c + 1] = (rowsum[r] > 0.0) ? (float) (a2 / rowsum[r]) : (float) (c + 1) / (float) w;
            g_env.pdf[(size_t) r * w + c] = (total > 0.0) ? (float) (f / total * (double) w * (double) (h - 1)) : 1.f;
        }
        row[w] = 1.f;
    }
    g_env.cdf_m[h - 1] = 1.f;
    free(g); free(rowsum);
}

/* world direction -> cell (r, c), in-cell position (fu, fv) and sin(theta) */
static inline void env_locate(const rcv_scene_desc *S, v3 d, int *r, int *c, float *fu, float *fv, float *st_out) {
    const float *R = S->env_to_world;
    float lx = fmaf(R[0], d.x, fmaf(R[3], d.y, R[6] * d.z));       /* R^T d */
    float ly = fmaf(R[1], d.x, fmaf(R[4], d.y, R[7] * d.z));
    float lz = fmaf(R[2], d.x, fmaf(R[5], d.y, R[8] * d.z));
    float st = sqrtf(fmaf(lx, lx, lz * lz));
    float u = orc_atan2f(lx, -lz) * (0.5f / ORC_PI);
    u = u - floorf(u);
    if (!(u < 1.f)) u = 0.f;
    float v = orc_atan2f(st, ly) * (1.f / ORC_PI);
    float x = u * (float) g_env.w, y = v * (float) (g_env.h - 1);
    int ci = (int) floorf(x), ri = (int) floorf(y);
    ci = ci < g_env.w - 1 ? ci : g_env.w - 1; ri = ri < g_env.h - 2 ? ri : g_env.h - 2;
    ci = ci > 0 ? ci : 0; ri = ri > 0 ? ri : 0;
    *c = ci; *r = ri; *fu = x - (float) ci; *fv = y - (float) ri; *st_out = st;
}
static inline v3 env_eval(const rcv_scene_desc *S, v3 d) {
    int r, c; float fu, fv, st;
    env_locate(S, d, &r, &c, &fu, &fv, &st);
    const float *t00 = g_env.data + 3 * ((size_t) r * (g_env.w + 1) + c), *t01 = t00 + 3;
    const float *t10 = t00 + 3 * (g_env.w + 1), *t11 = t10 + 3;
    float gu = 1.f - fu, gv = 1.f - fv, o[3];
    for (int k = 0; k < 3; ++k) {
        float a = fmaf(gu, t00[k], fu * t01[k]), b = fmaf(gu, t10[k], fu * t11[k]);
        o[k] = fmaf(gv, a, fv * b);
    }
    return V3(o[0], o[1], o[2]);
}
static inline float env_pdf(const rcv_scene_desc *S, v3 d) {
    int r, c; float fu, fv, st;
    env_locate(S, d, &r, &c, &fu, &fv, &st);
    return g_env.pdf[(size_t) r * g_env.w + c] / ((2.f * ORC_PI * ORC_PI) * fmaxf_(st, 1e-6f));
}
/* largest i in [0, n - 2] with cdf[i] <= u */
static inline int cdf_find(const float *cdf, int n, float u) {
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (cdf[mid] <= u) lo = mid; else hi = mid; }
    return lo;
}
static inline v3 env_sample(const rcv_scene_desc *S, float u1, float u2) {
    int r = cdf_find(g_env.cdf_m, g_env.h, u2);
    float den = g_env.cdf_m[r + 1] - g_env.cdf_m[r];
    float dv = (den > 0.f) ? (u2 - g_env.cdf_m[r]) / den : 0.5f;
    const float *row = g_env.cdf_c + (size_t) r * (g_env.w + 1);
    int c = cdf_find(row, g_env.w + 1, u1);
    den = row[c + 1] - row[c];
    float du = (den > 0.f) ? (u1 - row[c]) / den : 0.5f;
    float u = ((float) c + du) / (float) g_env.w, v = ((float) r + dv) / (float) (g_env.h - 1);
    float sp, cp, st, ct;
    orc_sincosf(u * ORC_TWO_PI, &sp, &cp);
    orc_sincosf(v * ORC_PI, &st, &ct);
    float lx = sp * st, ly = ct, lz = -(cp * st);
    const float *R = S->env_to_world;
    return V3(fmaf(R[0], lx, fmaf(R[1], ly, R[2] * lz)), fmaf(R[3], lx, fmaf(R[4], ly, R[5] * lz)),
              fmaf(R[6], lx, fmaf(R[7], ly, R[8] * lz)));
}

/* The scene's environment emitter (`constant` [UPSTREAM constant.cpp] or `envmap`): radiance arriving along a ray of
 * direction d that leaves the scene, the solid-angle density of sampling d, and the direction sample.  For the envmap
 * the sample's density and value are evaluated FROM the sampled direction (env_pdf / env_eval), so emitter hits and
 * emitter samples use the same two functions. */
static inline v3 emitter_eval(const rcv_scene_desc *S, v3 d) {
    if (S->emitter_type == RCV_EMITTER_ENVMAP) return env_eval(S, d);
    if (S->emitter_type == RCV_EMITTER_AREA) return V3(0, 0, 0);
    return V3(S->env_radiance[0], S->env_radiance[1], S->env_radiance[2]);
}
static inline float emitter_pdf(const rcv_scene_desc *S, v3 d) {
    if (S->emitter_type == RCV_EMITTER_ENVMAP) return env_pdf(S, d);
    if (S->emitter_type == RCV_EMITTER_AREA) return 0.f;
    return ORC_INV_FOURPI;
}
/* d, pdf and radiance / pdf of an environment sample */
static inline void emitter_sample_env(const rcv_scene_desc *S, float s2x, float s2y, v3 *d, float *pdf, v3 *val) {
    if (S->emitter_type == RCV_EMITTER_ENVMAP) {
        *d = env_sample(S, s2x, s2y);
        *pdf = env_pdf(S, *d);
        v3 e = env_eval(S, *d);
        *val = (*pdf != 0.f) ? V3(e.x / *pdf, e.y / *pdf, e.z / *pdf) : V3(0, 0, 0);
        return;
    }
    *d = square_to_uniform_sphere(s2x, s2y);
    *pdf = ORC_INV_FOURPI;
    *val = V3(S->env_radiance[0] / *pdf, S->env_radiance[1] / *pdf, S->env_radiance[2] / *pdf);
}
/* test hooks */
void orc_env_eval(const rcv_scene_desc *S, const float d[3], float rgb[3], float *pdf) {
    env_build(S);
    v3 e = emitter_eval(S, V3(d[0], d[1], d[2]));
    rgb[0] = e.x; rgb[1] = e.y; rgb[2] = e.z; *pdf = emitter_pdf(S, V3(d[0], d[1], d[2]));
}
void orc_env_sample(const rcv_scene_desc *S, float u1, float u2, float d[3], float *pdf, float val[3]) {
    env_build(S);
    v3 dd, vv;
    emitter_sample_env(S, u1, u2, &dd, pdf, &vv);
    d[0] = dd.x; d[1] = dd.y; d[2] = dd.z; val[0] = vv.x; val[1] = vv.y; val[2] = vv.z;
}

/* ------------------------------------------------------------------------- */
/* scene geometry: one axis-aligned box with a null BSDF                     */
/* ------------------------------------------------------------------------- */
typedef struct {
    int valid;     /* si.is_valid() */
    float t;       /* inf when invalid */
    v3 p, n;
    int shape;     /* -1: the medium's box (null BSDF, medium transition); k >= 0: rectangle k (diffuse BSDF) */
} si_t;

typedef struct { v3 o, d; float maxt; } ray_t;

/* BoundingBox::ray_intersect [UPSTREAM bbox.h] */
static inline int aabb_intersect(const rcv_scene_desc *S, v3 o, v3 d, float *mint, float *maxt) {
    const float *bmin = S->bbox_min, *bmax = S->bbox_max;
    float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    int active = 1;
    float tn = -ORC_INF, tf = ORC_INF;
    for (int a = 0; a < 3; ++a) {
        active = active && (dd[a] != 0.f || oo[a] > bmin[a] || oo[a] < bmax[a]);
        float rcp = 1.f / dd[a];
        float t1 = (bmin[a] - oo[a]) * rcp, t2 = (bmax[a] - oo[a]) * rcp;
        float t1p = fminf_(t1, t2), t2p = fmaxf_(t1, t2);
        tn = fmaxf_(tn, t1p);
        tf = fminf_(tf, t2p);
    }
    active = active && (tf >= tn);
    *mint = tn; *maxt = tf;
    return active;
}

/* scene.ray_intersect for the box shape: nearest boundary hit with 0 < t <= maxt.
 * (The reference's scene is a cube mesh traced by OptiX/Embree; the analytic
 * slab test is this build's equivalent) */
static inline si_t scene_intersect(const rcv_scene_desc *S, const ray_t *ray) {
    si_t si; si.valid = 0; si.t = ORC_INF; si.p = V3(0, 0, 0); si.n = V3(0, 0, 0); si.shape = -1;
    const float *bmin = S->bbox_min, *bmax = S->bbox_max;
    float oo[3] = {ray->o.x, ray->o.y, ray->o.z}, dd[3] = {ray->d.x, ray->d.y, ray->d.z};
    float tn = -ORC_INF, tf = ORC_INF; int an = 0, af = 0;
    for (int a = 0; a < 3; ++a) {
        float rcp = 1.f / dd[a];
        float t1 = (bmin[a] - oo[a]) * rcp, t2 = (bmax[a] - oo[a]) * rcp;
        float t1p = fminf_(t1, t2), t2p = fmaxf_(t1, t2);
        if (t1p > tn) { tn = t1p; an = a; }
        if (t2p < tf) { tf = t2p; af = a; }
    }
    if (!(tf >= tn)) return si;
    float t; int axis;
    if (tn > 0.f) { t = tn; axis = an; }
    else if (tf > 0.f) { t = tf; axis = af; }
    else return si;
    if (!(t <= ray->maxt) || !isfinite(t)) return si;
    float pp[3] = {fmaf(dd[0], t, oo[0]), fmaf(dd[1], t, oo[1]), fmaf(dd[2], t, oo[2])};
    float nn[3] = {0.f, 0.f, 0.f};
    /* outward normal of the face that was hit; snap p onto the face */
    int entering = (tn > 0.f);
    int positive_face = entering ? (dd[axis] < 0.f) : (dd[axis] > 0.f);
    nn[axis] = positive_face ? 1.f : -1.f;
    pp[axis] = positive_face ? bmax[axis] : bmin[axis];
    si.valid = 1; si.t = t;
    si.p = V3(pp[0], pp[1], pp[2]); si.n = V3(nn[0], nn[1], nn[2]);
    return si;
}

/* ------------------------------------------------------------------------- */
/* rectangles (ABI 4): `rectangle` shapes with a one-sided `diffuse` BSDF and, on one of them, the `area` light
 * [UPSTREAM rectangle.cpp, diffuse.cpp, area.cpp, shape.cpp sample_direction / pdf_direction; restated].
 * Derived quantities are computed in binary32 with the operation order below (the library does the same). */
/* ------------------------------------------------------------------------- */
#define ORC_INV_PI 0.31830988618379067154f
typedef struct { v3 p0, eu, ev, n, s, t, refl, rad; float inv_eu2, inv_ev2, inv_area; } rquad_t;
static rquad_t g_quads[RCV_MAX_QUADS];
static inline float dot3(v3 a, v3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
static inline v3 cross3(v3 a, v3 b) {
    return V3(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
static void quads_build(const rcv_scene_desc *S) {
    for (uint32_t k = 0; k < S->n_quads && k < RCV_MAX_QUADS; ++k) {
        const rcv_quad *q = &S->quads[k]; rquad_t *g = &g_quads[k];
        g->p0 = V3(q->p0[0], q->p0[1], q->p0[2]); g->eu = V3(q->eu[0], q->eu[1], q->eu[2]); g->ev = V3(q->ev[0], q->ev[1], q->ev[2]);
        g->refl = V3(q->reflectance[0], q->reflectance[1], q->reflectance[2]);
        g->rad = V3(q->radiance[0], q->radiance[1], q->radiance[2]);
        v3 c = cross3(g->eu, g->ev);
        float len = sqrtf(dot3(c, c));
        g->n = V3(c.x / len, c.y / len, c.z / len);
        float lu2 = dot3(g->eu, g->eu), lv2 = dot3(g->ev, g->ev);
        g->inv_eu2 = 1.f / lu2; g->inv_ev2 = 1.f / lv2; g->inv_area = 1.f / len;
        float lu = sqrtf(lu2);
        g->s = V3(g->eu.x / lu, g->eu.y / lu, g->eu.z / lu);          /* shading frame (s, t, n) */
        g->t = cross3(g->n, g->s);
    }
}
/* scene.ray_intersect: the nearest of the medium's box and the rectangles, 0 < t <= maxt */
static inline si_t scene_intersect_all(const rcv_scene_desc *S, const ray_t *ray) {
    si_t si = scene_intersect(S, ray);
    for (uint32_t k = 0; k < S->n_quads; ++k) {
        const rquad_t *q = &g_quads[k];
        float denom = dot3(q->n, ray->d);
        if (denom == 0.f) continue;
        float t = dot3(q->n, vsub(q->p0, ray->o)) / denom;
        if (!(t > 0.f) || !(t <= ray->maxt) || !(t < si.t)) continue;
        v3 p = ray_at(ray->o, ray->d, t);
        v3 rel = vsub(p, q->p0);
        float a = dot3(rel, q->eu) * q->inv_eu2, b = dot3(rel, q->ev) * q->inv_ev2;
        if (!(a >= 0.f && a <= 1.f && b >= 0.f && b <= 1.f)) continue;
        si.valid = 1; si.t = t; si.p = p; si.n = q->n; si.shape = (int) k;
    }
    return si;
}
/* diffuse.cpp eval / pdf for incident ray direction d_in (wi = -d_in) and outgoing world direction wo */
static inline void diffuse_eval_pdf(const rquad_t *q, v3 d_in, v3 wo, v3 *val, float *pdf) {
    float ci = -dot3(q->n, d_in), co = dot3(q->n, wo);
    if (ci > 0.f && co > 0.f) {
        *pdf = ORC_INV_PI * co;
        *val = vscale(vscale(q->refl, ORC_INV_PI), co);
    } else { *pdf = 0.f; *val = V3(0, 0, 0); }
}
/* warp.square_to_cosine_hemisphere (concentric disk) [UPSTREAM warp.h] */
static inline v3 square_to_cosine_hemisphere(float sx, float sy) {
    float x = fmaf(2.f, sx, -1.f), y = fmaf(2.f, sy, -1.f);
    int is_zero = (x == 0.f) && (y == 0.f), q13 = fabsf(x) < fabsf(y);
    float r = q13 ? y : x, rp = q13 ? x : y;
    float phi = (0.25f * ORC_PI) * (rp / r);
    if (q13) phi = (0.5f * ORC_PI) - phi;
    if (is_zero) phi = 0.f;
    float sn, cs;
    orc_sincosf(fabsf(phi), &sn, &cs);
    if (phi < 0.f) sn = -sn;
    float px = r * cs, py = r * sn;
    float z = sqrtf(fmaxf_(1.f - fmaf(px, px, py * py), 0.f));
    return V3(px, py, z);
}
/* area light on rectangle S->area_quad: sample_direction / pdf_direction / eval */
static inline void area_sample(const rcv_scene_desc *S, v3 ref_p, float s2x, float s2y, v3 *d, float *dist, float *pdf, v3 *val) {
    const rquad_t *q = &g_quads[S->area_quad];
    v3 p = V3(fmaf(s2y, q->ev.x, fmaf(s2x, q->eu.x, q->p0.x)), fmaf(s2y, q->ev.y, fmaf(s2x, q->eu.y, q->p0.y)),
              fmaf(s2y, q->ev.z, fmaf(s2x, q->eu.z, q->p0.z)));
    v3 v = vsub(p, ref_p);
    float dist2 = dot3(v, v);
    *dist = sqrtf(dist2);
    *d = V3(v.x / *dist, v.y / *dist, v.z / *dist);
    float dp = dot3(*d, q->n);
    float x = dist2 / fabsf(dp);
    *pdf = isfinite(x) ? q->inv_area * x : 0.f;
    int ok = (dp < 0.f) && (*pdf != 0.f);
    *val = ok ? V3(q->rad.x / *pdf, q->rad.y / *pdf, q->rad.z / *pdf) : V3(0, 0, 0);
}
static inline float area_pdf(const rcv_scene_desc *S, v3 ref_p, v3 hit_p) {
    const rquad_t *q = &g_quads[S->area_quad];
    v3 v = vsub(hit_p, ref_p);
    float dist2 = dot3(v, v), dist = sqrtf(dist2);
    v3 d = V3(v.x / dist, v.y / dist, v.z / dist);
    float dp = dot3(d, q->n);
    float x = dist2 / fabsf(dp);
    return ((dp < 0.f) && isfinite(x)) ? q->inv_area * x : 0.f;
}

/* Interaction::spawn_ray / offset_p [UPSTREAM interaction.h] */
static inline ray_t spawn_ray(v3 p, v3 n, v3 d) {
    float mag = (1.f + fmaxf_(fmaxf_(fabsf(p.x), fabsf(p.y)), fabsf(p.z))) * ORC_RAY_EPS;
    float dn = fmaf(n.x, d.x, fmaf(n.y, d.y, n.z * d.z));
    if (dn < 0.f) mag = -mag;                   /* mulsign(mag, dot(n, d)) */
    ray_t r;
    r.o = V3(fmaf(mag, n.x, p.x), fmaf(mag, n.y, p.y), fmaf(mag, n.z, p.z));
    r.d = d; r.maxt = ORC_LARGEST;
    return r;
}

/* ------------------------------------------------------------------------- */
/* grid volume lookups  [UPSTREAM gridvolume.cpp / drjit texture.h, accel=False] */
/* ------------------------------------------------------------------------- */
typedef struct {
    int inside;
    int i0[3], i1[3];
    float w0[3], w1[3];
} tri_t;

static inline tri_t tri_setup(const rcv_scene_desc *S, const orc_medium *M, v3 p) {
    tri_t T;
    float pp[3] = {p.x, p.y, p.z};
    int res[3] = {M->nx, M->ny, M->nz};
    T.inside = 1;
    for (int a = 0; a < 3; ++a) {
        float ext_inv = 1.f / (S->bbox_max[a] - S->bbox_min[a]);
        float pl = (pp[a] - S->bbox_min[a]) * ext_inv;          /* to_local * p */
        T.inside = T.inside && (pl >= 0.f) && (pl <= 1.f);
        float pos = fmaf(pl, (float) res[a], -0.5f);
        float fl = floorf(pos);
        int i = (int) fl;
        T.w1[a] = pos - fl;
        T.w0[a] = 1.f - T.w1[a];
        int i0 = i, i1 = i + 1;
        if (i0 < 0) i0 = 0; if (i0 > res[a] - 1) i0 = res[a] - 1;
        if (i1 < 0) i1 = 0; if (i1 > res[a] - 1) i1 = res[a] - 1;
        T.i0[a] = i0; T.i1[a] = i1;
    }
    return T;
}
static inline float tri_lerp(const tri_t *T, const float v[8]) {
    /* v[k], k = 4*zbit + 2*ybit + xbit */
    float v0 = fmaf(T->w0[0], v[0], T->w1[0] * v[1]);
    float v1 = fmaf(T->w0[0], v[2], T->w1[0] * v[3]);
    float v2 = fmaf(T->w0[0], v[4], T->w1[0] * v[5]);
    float v3_ = fmaf(T->w0[0], v[6], T->w1[0] * v[7]);
    float v4 = fmaf(T->w0[1], v0, T->w1[1] * v1);
    float v5 = fmaf(T->w0[1], v2, T->w1[1] * v3_);
    return fmaf(T->w0[2], v4, T->w1[2] * v5);
}
static inline size_t tri_index(const orc_medium *M, const tri_t *T, int k) {
    int x = (k & 1) ? T->i1[0] : T->i0[0];
    int y = (k & 2) ? T->i1[1] : T->i0[1];
    int z = (k & 4) ? T->i1[2] : T->i0[2];
    return ((size_t) z * M->ny + y) * M->nx + x;
}
static inline float tri_weight(const tri_t *T, int k) {
    float wx = (k & 1) ? T->w1[0] : T->w0[0];
    float wy = (k & 2) ? T->w1[1] : T->w0[1];
    float wz = (k & 4) ? T->w1[2] : T->w0[2];
    return (wz * wy) * wx;
}
/* sigma_t(p) = scale * grid(p) (monochromatic), 0 outside the volume */
static inline float lookup_sigma_t(const rcv_scene_desc *S, const orc_medium *M, const tri_t *T) {
    if (!T->inside) return 0.f;
    float v[8];
    for (int k = 0; k < 8; ++k) v[k] = M->sigma_t[tri_index(M, T, k)];
    return S->sigma_t_scale * tri_lerp(T, v);
}
static inline v3 lookup_albedo(const orc_medium *M, const tri_t *T) {
    if (!T->inside) return V3(0, 0, 0);
    float out[3];
    for (int c = 0; c < 3; ++c) {
        float v[8];
        for (int k = 0; k < 8; ++k) v[k] = M->albedo[tri_index(M, T, k) * 3 + c];
        out[c] = tri_lerp(T, v);
    }
    return V3(out[0], out[1], out[2]);
}

float orc_majorant(const rcv_scene_desc *S, const float *sigma_t, size_t n) {
    /* heterogeneous: m_max_density = scale * sigma_t.max() [UPSTREAM] */
    float m = sigma_t[0];
    for (size_t i = 1; i < n; ++i) m = fmaxf_(m, sigma_t[i]);
    return S->sigma_t_scale * m;
}

/* ------------------------------------------------------------------------- */
/* event counters (roofline accounting + path-identity checks)               */
/* ------------------------------------------------------------------------- */
/* ------------------------------------------------------------------------- */
/* majorant supergrid  [UPSTREAM heterogeneous.cpp update_majorant_supergrid /  */
/* gridvolume.cpp local_majorants; sources absent from /root/reference, restated */
/* from the published Mitsuba 3 algorithm — [SPEC] marks the choices made here]  */
/* Each supergrid cell stores scale * max of the fine voxels whose trilinear      */
/* support can reach the cell: voxel indices [c*f - 1, c*f + f] per axis, clamped.*/
/* ------------------------------------------------------------------------- */
int orc_majorant_grid(const rcv_scene_desc *S, const float *sigma_t, int nz, int ny, int nx, int factor, float *out) {
    if (factor <= 0 || nz % factor || ny % factor || nx % factor) return -1;
    int mz = nz / factor, my = ny / factor, mx = nx / factor;
    for (int cz = 0; cz < mz; ++cz)
        for (int cy = 0; cy < my; ++cy)
            for (int cx = 0; cx < mx; ++cx) {
                float m = 0.f;                                   /* densities are >= 0 */
                for (int dz = -1; dz <= factor; ++dz)
                    for (int dy = -1; dy <= factor; ++dy)
                        for (int dx = -1; dx <= factor; ++dx) {
                            int z = cz * factor + dz, y = cy * factor + dy, x = cx * factor + dx;
                            z = z < 0 ? 0 : (z > nz - 1 ? nz - 1 : z);
                            y = y < 0 ? 0 : (y > ny - 1 ? ny - 1 : y);
                            x = x < 0 ? 0 : (x > nx - 1 ? nx - 1 : x);
                            m = fmaxf_(m, sigma_t[((size_t) z * ny + y) * nx + x]);
                        }
                out[((size_t) cz * my + cy) * mx + cx] = S->sigma_t_scale * m;
            }
    return 0;
}

/* nearest-neighbour lookup of the supergrid at a world position (majorant_grid.eval_1), clamped to the border cells */
static inline float majorant_at(const rcv_scene_desc *S, const orc_medium *M, v3 p) {
    float pp[3] = {p.x, p.y, p.z};
    int res[3] = {M->mx, M->my, M->mz}, idx[3];
    for (int a = 0; a < 3; ++a) {
        float ext_inv = 1.f / (S->bbox_max[a] - S->bbox_min[a]);
        float pl = (pp[a] - S->bbox_min[a]) * ext_inv;
        int i = (int) floorf(pl * (float) res[a]);               /* [SPEC] clamped, no inside test: see dda_sample */
        idx[a] = i < 0 ? 0 : (i > res[a] - 1 ? res[a] - 1 : i);
    }
    return M->majorant_grid[((size_t) idx[2] * M->my + idx[1]) * M->mx + idx[0]];
}
/* max over the one or two states that take part (Mo may be NULL) */
static inline float majorant2_at(const rcv_scene_desc *S, const orc_medium *Mo, const orc_medium *Mn, v3 p) {
    float m = majorant_at(S, Mn, p);
    return Mo ? fmaxf_(majorant_at(S, Mo, p), m) : m;
}

/* DDA free-flight sampling through the supergrid (twostatemedium.py:25-70; Medium::prepare_dda_traversal and the
 * identical single-state loop of Medium::sample_interaction are [UPSTREAM]).  Returns the sampled t (or inf) and, in
 * *maj_out, the majorant of the cell the sample fell into.
 * [SPEC] The traversal is the classic incremental one the reference's code was adapted from (fast voxel traversal):
 * the supercell index is carried along and advanced by +-1 on the axis whose boundary is crossed, instead of being
 * recovered from the midpoint of every interval (:53) — same cells except where float rounding puts a midpoint on the
 * other side of a boundary, where the one-voxel margin of the supergrid keeps either majorant valid — and the majorant
 * used for the collision weights is that of the stopping cell (:82-86 looks it up again at the sampled point). */
static float dda_sample(const rcv_scene_desc *S, const orc_medium *Mo, const orc_medium *Mn, const ray_t *ray,
                        float mint, float maxt, float desired_tau, float *maj_out) {
    float o[3] = {ray->o.x, ray->o.y, ray->o.z}, d[3] = {ray->d.x, ray->d.y, ray->d.z};
    int res[3] = {Mn->mx, Mn->my, Mn->mz};
    float tmax3[3], tdelta[3];
    int cell[3], step[3];
    for (int a = 0; a < 3; ++a) {
        float ext = S->bbox_max[a] - S->bbox_min[a];
        float ext_inv = 1.f / ext;
        float lo = (o[a] - S->bbox_min[a]) * ext_inv, ld = d[a] * ext_inv;
        float fres = (float) res[a], voxel = 1.f / fres;
        int cur = (int) floorf(fmaf(ld, mint, lo) * fres);
        cur = cur < 0 ? 0 : (cur > res[a] - 1 ? res[a] - 1 : cur);
        cell[a] = cur;
        step[a] = (ld >= 0.f) ? 1 : -1;
        float nb = (float) (cur + (ld >= 0.f ? 1 : 0)) * voxel;  /* the boundary ahead */
        if (ld != 0.f) {
            float rld = (1.f / d[a]) * ext;                      /* 1 / ld from the ray's cached reciprocal direction */
            tmax3[a] = (nb - lo) * rld;
            tdelta[a] = ((float) step[a] * voxel) * rld;
        } else {
            tmax3[a] = ORC_INF; tdelta[a] = ORC_INF;
        }
    }
    float dda_t = mint, tau_acc = 0.f, majorant = 0.f;
    int reached = 0;
    while (1) {                                                  /* :39-66 */
        float t_next = fminf_(fminf_(tmax3[0], tmax3[1]), tmax3[2]);
        int axis = (tmax3[0] == t_next) ? 0 : ((tmax3[1] == t_next) ? 1 : 2);   /* first axis that attains the minimum */
        size_t idx = ((size_t) cell[2] * Mn->my + cell[1]) * Mn->mx + cell[0];
        majorant = Mn->majorant_grid[idx];
        if (Mo) majorant = fmaxf_(Mo->majorant_grid[idx], majorant);
        float tau_next = tau_acc + (majorant * (t_next - dda_t));
        if ((majorant > 0.f) && (tau_next >= desired_tau)) {     /* :59-62; the division only where it is used */
            float t_precise = dda_t + ((desired_tau - tau_acc) / majorant);
            if (t_precise < maxt) { dda_t = t_precise; reached = 1; break; }
        }
        dda_t = t_next;
        if (!(t_next < maxt)) break;
        cell[axis] += step[axis];
        if (cell[axis] < 0 || cell[axis] >= res[axis]) break;    /* left the supergrid (rounding at the exit) */
        tmax3[axis] = tmax3[axis] + tdelta[axis];
        tau_acc = tau_next;
    }
    *maj_out = majorant;
    return reached ? dda_t : ORC_INF;                            /* :70 */
}

/* test hook: the DDA alone (single state when Mo == NULL) */
float orc_dda_sample(const rcv_scene_desc *S, const orc_medium *Mo, const orc_medium *Mn, const float o[3], const float d[3],
                     float mint, float maxt, float desired_tau) {
    ray_t r; r.o = V3(o[0], o[1], o[2]); r.d = V3(d[0], d[1], d[2]); r.maxt = maxt;
    float mj;
    return dda_sample(S, Mo, Mn, &r, mint, maxt, desired_tau, &mj);
}

typedef struct { uint64_t c[8]; } counters_t;
static counters_t g_counters;
void orc_counters(uint64_t out[8], int reset) {
    memcpy(out, g_counters.c, sizeof(g_counters.c));
    if (reset) memset(&g_counters, 0, sizeof(g_counters));
}
static void counters_merge(const counters_t *c) {
#pragma omp critical(orc_counters)
    for (int i = 0; i < 8; ++i) g_counters.c[i] += c->c[i];
}

/* ------------------------------------------------------------------------- */
/* sensor: ADIntegrator.sample_rays + perspective.sample_ray  [UPSTREAM]      */
/* ------------------------------------------------------------------------- */
typedef struct { pcg32 rng; ray_t ray; int32_t pix; float fx, fy; } lane_t;   /* (fx, fy): film position in pixels */

/* ImageBlock.put [UPSTREAM imageblock.cpp]: the pixels a sample at film position (fx, fy) is splatted to and the
 * reconstruction-filter weights.  Box: the pixel that contains the sample, weight 1.  Gaussian (gaussian.cpp, the
 * hdrfilm default): radius r = 4 stddev, per-axis weight max(0, exp(a x^2) - exp(a r^2)), a = -1 / (2 stddev^2), at the
 * pixel centres within the radius (2 ceil(r) <= 4 per axis); weights are not normalised, develop() divides. */
typedef struct { int n; int32_t pix[16]; float w[16]; } footprint_t;
static inline void film_footprint(const rcv_scene_desc *S, float fx, float fy, footprint_t *fp) {
    int W = (int) S->film_width, H = (int) S->film_height;
    fp->n = 0;
    if (S->rfilter_type != RCV_RFILTER_GAUSSIAN) {
        int ix = (int) floorf(fx), iy = (int) floorf(fy);
        if (ix >= 0 && ix < W && iy >= 0 && iy < H) { fp->pix[0] = iy * W + ix; fp->w[0] = 1.f; fp->n = 1; }
        return;
    }
    float sd = S->rfilter_stddev, r = 4.f * sd;
    float alpha = -1.f / ((2.f * sd) * sd);
    float bias = orc_expf(alpha * (r * r));
    int count = 2 * (int) ceilf(r);
    float px = fx - 0.5f, py = fy - 0.5f;                      /* pixel centres sit at integer + 0.5 */
    int lox = (int) ceilf(px - r), loy = (int) ceilf(py - r);
    float wx[4], wy[4];
    for (int i = 0; i < count; ++i) {
        float dx = (float) (lox + i) - px, dy = (float) (loy + i) - py;
        wx[i] = fmaxf_(0.f, orc_expf(alpha * (dx * dx)) - bias);
        wy[i] = fmaxf_(0.f, orc_expf(alpha * (dy * dy)) - bias);
    }
    for (int j = 0; j < count; ++j)
        for (int i = 0; i < count; ++i) {
            int x = lox + i, y = loy + j;
            if (x < 0 || x >= W || y < 0 || y >= H) continue;
            fp->pix[fp->n] = y * W + x; fp->w[fp->n] = wy[j] * wx[i]; fp->n++;
        }
}

static inline lane_t lane_setup(const rcv_scene_desc *S, uint32_t seed, uint32_t spp, uint64_t idx) {
    lane_t L;
    sampler_seed(&L.rng, seed, (uint32_t) idx);
    uint32_t W = S->film_width, H = S->film_height;
    uint32_t pixel = (uint32_t) (idx / spp);
    uint32_t py = pixel / W, px = pixel - W * py;
    float u1 = next_1d(&L.rng, 1), u2 = next_1d(&L.rng, 1);      /* sampler.next_2d() */
    float pfx = (float) px + u1, pfy = (float) py + u2;           /* pos_f */
    /* ImageBlock.put, box filter: pixel = floor(pos), dropped when outside */
    float fx = floorf(pfx), fy = floorf(pfy);
    int ix = (int) fx, iy = (int) fy;
    L.pix = (ix >= 0 && ix < (int) W && iy >= 0 && iy < (int) H) ? (int32_t) (iy * (int) W + ix) : -1;
    L.fx = pfx; L.fy = pfy;
    float sx = pfx * (1.f / (float) W), sy = pfy * (1.f / (float) H);   /* pos_adjusted */
    const float *A = S->cam_axes, *O = S->cam_origin;
    if (S->n_sensors > 1) {
        /* `batch` sensor [UPSTREAM batch.cpp sample_ray]: the view is chosen by the film column and sees
         * the sample at its own normalised position */
        float fs = sx * (float) S->n_sensors;
        uint32_t iu = (uint32_t) fs;
        uint32_t si = iu < S->n_sensors - 1 ? iu : S->n_sensors - 1;
        sx = fs - (float) iu;
        A = S->batch_axes[si]; O = S->batch_origin[si];
    }
    /* near-plane point in camera space (x to the left, y up, z forward) */
    float nx = fmaf(sx, -2.f * S->cam_tan_x, S->cam_tan_x);
    float ny = fmaf(sy, -2.f * S->cam_tan_y, S->cam_tan_y);
    float len2 = fmaf(nx, nx, fmaf(ny, ny, 1.f));
    float inv = 1.f / sqrtf(len2);
    float dx = nx * inv, dy = ny * inv, dz = inv;
    v3 d = V3(fmaf(A[0], dx, fmaf(A[1], dy, A[2] * dz)),
              fmaf(A[3], dx, fmaf(A[4], dy, A[5] * dz)),
              fmaf(A[6], dx, fmaf(A[7], dy, A[8] * dz)));
    float near_t = S->cam_near_clip / dz;
    L.ray.o = ray_at(V3(O[0], O[1], O[2]), d, near_t);
    L.ray.d = d;
    L.ray.maxt = ORC_LARGEST;
    return L;
}

/* phase function sample: returns wo and pdf  (isotropic / HG) [UPSTREAM] */
static inline v3 phase_sample(const rcv_scene_desc *S, v3 wi, float s1, float s2x, float s2y, float *pdf) {
    (void) s1; (void) wi;
    if (S->phase_type == RCV_PHASE_ISOTROPIC) {
        *pdf = ORC_INV_FOURPI;
        return square_to_uniform_sphere(s2x, s2y);
    }
    /* Henyey-Greenstein [UPSTREAM hg.cpp]: sample cos_theta, build frame around -wi */
    float g = S->phase_g;
    float sqr_term = (1.f - g * g) / (1.f - g + 2.f * g * s2x);
    float cos_theta = (1.f + g * g - sqr_term * sqr_term) / (2.f * g);
    if (fabsf(g) < 1.1920929e-7f) cos_theta = 1.f - 2.f * s2x;
    float sin_theta = sqrtf(fmaxf_(fmaf(-cos_theta, cos_theta, 1.f), 0.f));
    float sp, cp; orc_sincosf(ORC_TWO_PI * s2y, &sp, &cp);
    v3 local = V3(sin_theta * cp, sin_theta * sp, -cos_theta);
    /* coordinate_system(wi) (Duff et al.) */
    float sign = copysignf(1.f, wi.z);
    float a = -1.f / (sign + wi.z);
    float b = wi.x * wi.y * a;
    v3 s = V3(fmaf(wi.x * wi.x, a * sign, 1.f), b * sign, -wi.x * sign);
    v3 t = V3(b, fmaf(wi.y * wi.y, a, sign), -wi.y);
    v3 wo = V3(fmaf(s.x, local.x, fmaf(t.x, local.y, wi.x * local.z)),
               fmaf(s.y, local.x, fmaf(t.y, local.y, wi.y * local.z)),
               fmaf(s.z, local.x, fmaf(t.z, local.y, wi.z * local.z)));
    float temp = 1.f + g * g - 2.f * g * cos_theta;          /* eval_hg(-cos_theta) */
    *pdf = ORC_INV_FOURPI * (1.f - g * g) / (temp * sqrtf(temp));
    return wo;
}
static inline float phase_eval(const rcv_scene_desc *S, v3 wi, v3 wo) {
    if (S->phase_type == RCV_PHASE_ISOTROPIC) return ORC_INV_FOURPI;
    float g = S->phase_g;
    float cos_theta = fmaf(wi.x, wo.x, fmaf(wi.y, wo.y, wi.z * wo.z));
    float temp = 1.f + g * g + 2.f * g * cos_theta;
    return ORC_INV_FOURPI * (1.f - g * g) / (temp * sqrtf(temp));
}

/* ------------------------------------------------------------------------- */
/* TwoStateMedium.sample_interaction_twostates  (plugins/twostatemedium.py:15-107) */
/* ------------------------------------------------------------------------- */
typedef struct {
    float t;                 /* inf when invalid */
    v3 p;
    float mint;
    v3 sigma_s, sigma_n, sigma_t;
    float combined_extinction;
    v3 albedo;               /* of state `new` (sigma_s = sigma_t * albedo), for the prbvolpath adjoint */
} mei_t;

static inline int mei_valid(const mei_t *m) { return m->t < ORC_INF; } /* isfinite(t) */

/* Medium::prepare_interaction_sampling [UPSTREAM medium.cpp] (twostatemedium.py:22) */
static inline int prepare_interaction_sampling(const rcv_scene_desc *S, const ray_t *ray, int active,
                                               float *mint, float *maxt) {
    float a, b;
    int hit = aabb_intersect(S, ray->o, ray->d, &a, &b);
    hit = hit && (isfinite(a) || isfinite(b));
    active = active && hit;
    if (!active) { a = 0.f; b = ORC_INF; }
    *mint = fmaxf_(0.f, a);
    *maxt = fminf_(ray->maxt, b);
    return active;
}

static void sample_interaction_twostates(const rcv_scene_desc *S, const orc_medium *Mo, const orc_medium *Mn,
                                         const ray_t *ray, float sample, int active,
                                         mei_t *mei_old, mei_t *mei_new, counters_t *cnt) {
    float mint, maxt;
    active = prepare_interaction_sampling(S, ray, active, &mint, &maxt);       /* :22 */
    float desired_tau = -orc_logf(1.f - sample);                               /* :24 */
    float combined_extinction, sampled_t;
    if (Mn->majorant_grid) {
        /* spatially varying majorant: DDA through max(old, new) supergrids (:25-70) */
        float mj = 0.f;
        sampled_t = active ? dda_sample(S, Mo, Mn, ray, mint, maxt, desired_tau, &mj) : ORC_INF;
        combined_extinction = (active && (sampled_t <= maxt)) ? mj : 0.f;   /* the stopping cell's (:82-86) */
    } else {
        /* single majorant for the whole volume (:71-78) */
        combined_extinction = fmaxf_(Mo->majorant, Mn->majorant);              /* :73-75 */
        float m = combined_extinction;                                         /* :76, channel 0 */
        float inv_m = 1.f / m;                                                 /* [SPEC] x / majorant == x * (1 / majorant) */
        sampled_t = mint + (desired_tau * inv_m);                              /* :78 */
    }
    int valid_mi = active && (sampled_t <= maxt);                              /* :80 */
    mei_t mn;
    mn.t = valid_mi ? sampled_t : ORC_INF;                                     /* :88 */
    mn.p = ray_at(ray->o, ray->d, sampled_t);                                  /* :89 */
    mn.mint = mint;
    mn.albedo = V3(0, 0, 0);
    v3 ss_o = V3(0, 0, 0), st_o = V3(0, 0, 0), ss_n = V3(0, 0, 0), st_n = V3(0, 0, 0);
    if (valid_mi) {
        /* get_scattering_coefficients of both states at the same point (:93-94):
         * sigma_t = scale * grid, sigma_s = sigma_t * albedo [UPSTREAM heterogeneous.cpp] */
        tri_t T = tri_setup(S, Mn, mn.p);
        float so = lookup_sigma_t(S, Mo, &T), sn = lookup_sigma_t(S, Mn, &T);
        v3 ao = lookup_albedo(Mo, &T), an = lookup_albedo(Mn, &T);
        st_o = V3(so, so, so); st_n = V3(sn, sn, sn);
        ss_o = vmul(st_o, ao); ss_n = vmul(st_n, an);
        mn.albedo = an;
        cnt->c[0]++;
    }
    v3 ce = V3(combined_extinction, combined_extinction, combined_extinction);
    mn.sigma_s = ss_n; mn.sigma_t = st_n; mn.sigma_n = vsub(ce, st_n);         /* :96-97 */
    mn.combined_extinction = combined_extinction;
    *mei_old = mn;                                                             /* :99 */
    mei_old->sigma_s = ss_o; mei_old->sigma_t = st_o; mei_old->sigma_n = vsub(ce, st_o);
    *mei_new = mn;
}

/* ------------------------------------------------------------------------- */
/* gradient sink (the adjoints scatter d loss / d sigma_t.data, d albedo.data) */
/* ------------------------------------------------------------------------- */
typedef struct {
    double *g_sigma_t;   /* (nz,ny,nx,1) or NULL in primal mode */
    double *g_albedo;    /* (nz,ny,nx,3) */
    int atomic;          /* use omp atomic (multi-threaded) */
} grad_sink;

static inline void sink_add(double *a, double v, int atomic) {
    if (atomic) {
#pragma omp atomic
        *a += v;
    } else *a += v;
}
/* AD of the trilinear gather = scatter-add of w_corner * grad (SURVEY §2.3 gridvolume) */
static void scatter_sigma_t(const rcv_scene_desc *S, const orc_medium *M, grad_sink *G, v3 p, float g, counters_t *cnt) {
    tri_t T = tri_setup(S, M, p);
    if (!T.inside) return;
    for (int k = 0; k < 8; ++k)
        sink_add(&G->g_sigma_t[tri_index(M, &T, k)], (double) (tri_weight(&T, k) * (S->sigma_t_scale * g)), G->atomic);
    cnt->c[4]++;
}
static void scatter_albedo(const rcv_scene_desc *S, const orc_medium *M, grad_sink *G, v3 p, v3 g, counters_t *cnt) {
    tri_t T = tri_setup(S, M, p);
    if (!T.inside) return;
    for (int k = 0; k < 8; ++k) {
        float w = tri_weight(&T, k);
        size_t o = tri_index(M, &T, k) * 3;
        sink_add(&G->g_albedo[o], (double) (w * g.x), G->atomic);
        sink_add(&G->g_albedo[o + 1], (double) (w * g.y), G->atomic);
        sink_add(&G->g_albedo[o + 2], (double) (w * g.z), G->atomic);
    }
    cnt->c[5]++;
}

/* PRBVolpathIntegrator.sample in Backward mode [UPSTREAM-RECALLED mi.ad.integrators.prbvolpath; the adjoint the
 * two-state integrator inherits, plugins/twostatevolpath.py:11]: the same path as the primal pass (same stream), L
 * starts as the primal result and every contribution is subtracted when it is met, so that at each medium event L is
 * the radiance still to come; the event's weight w(sigma_t, albedo) is differentiated as dL * L * dw / w. */
typedef struct {
    v3 dL;          /* d loss / d radiance of this sample */
    v3 L_total;     /* result of the primal pass */
    grad_sink *G;
} prb_adj;

/* ------------------------------------------------------------------------- */
/* sample_emitter_twostates  (plugins/twostatevolpath.py:387-479)             */
/* `adjoint` != NULL: prbvolpath's sample_emitter in Backward mode — the same walk with the cloned sampler, scattering
 * adjoint * d log(sigma_n / majorant) / d sigma_t at every ratio-tracking collision [UPSTREAM-RECALLED]            */
/* ------------------------------------------------------------------------- */
static void sample_emitter_twostates(const rcv_scene_desc *S, const orc_medium *Mo, const orc_medium *Mn,
                                     pcg32 *rng, v3 ref_p, v3 ref_n, int medium, int active,
                                     v3 *emitted_old, v3 *emitted_new, v3 *ds_d, float *ds_pdf,
                                     const v3 *adjoint, grad_sink *G, counters_t *cnt) {
    if (!active) medium = 0;                                                   /* :391 */
    /* scene.sample_emitter_direction with the scene's one emitter: `constant` / `envmap` [UPSTREAM constant.cpp:
     * d uniform on the sphere, pdf 1/(4 pi), dist = 2 * bsphere radius, value = radiance / pdf] or the `area` light
     * (a point on the rectangle) */
    float s2x = next_1d(rng, active), s2y = next_1d(rng, active);              /* :393 next_2d(active) */
    v3 d = V3(0, 0, 0); float pdf = 0.f, dist = 0.f; v3 emitter_val = V3(0, 0, 0);
    if (active) {
        if (S->emitter_type == RCV_EMITTER_AREA) area_sample(S, ref_p, s2x, s2y, &d, &dist, &pdf, &emitter_val);
        else {
            emitter_sample_env(S, s2x, s2y, &d, &pdf, &emitter_val);           /* `constant` or `envmap` */
            dist = S->env_dist;
        }
    }
    int invalid = (pdf == 0.f);                                                /* :395 */
    if (invalid) emitter_val = V3(0, 0, 0);
    active = active && !invalid;                                               /* :397 */

    ray_t ray = spawn_ray(ref_p, ref_n, d);                                    /* :399 */
    float total_dist = 0.f;
    si_t si; si.valid = 0; si.t = 0.f; si.p = V3(0, 0, 0); si.n = V3(0, 0, 0); si.shape = -1; /* dr.zeros(SurfaceInteraction3f) */
    int needs_intersection = 1;
    v3 tr_old = V3(1, 1, 1), tr_new = V3(1, 1, 1);
    while (active) {                                                           /* :408 */
        float remaining_dist = dist * (1.f - ORC_SHADOW_EPS) - total_dist;     /* :409 */
        ray.maxt = remaining_dist;
        active = active && (remaining_dist > 0.f);
        needs_intersection = needs_intersection && active;                     /* :414 */
        if (needs_intersection) si = scene_intersect_all(S, &ray);
        needs_intersection = 0;
        int active_medium = active && (medium != 0);                           /* :418 */
        int active_surface = active && !active_medium;
        mei_t mo, mn;
        float u = next_1d(rng, active_medium);                                 /* :422 */
        sample_interaction_twostates(S, Mo, Mn, &ray, u, active_medium, &mo, &mn, cnt);
        if (active_medium && (si.t < mo.t)) mo.t = ORC_INF;                    /* :423 */
        if (active_medium && (si.t < mn.t)) mn.t = ORC_INF;
        v3 m_old = V3(1, 1, 1), m_new = V3(1, 1, 1);
        int escaped_medium = active_medium && !mei_valid(&mn);                 /* :441 */
        active_medium = active_medium && mei_valid(&mn);                       /* :444 */
        if (active_medium) {
            ray.o = mn.p;                                                      /* :445 */
            si.t = si.t - mn.t;                                                /* :446 */
            float ce = mn.combined_extinction;
            float inv_ce = 1.f / ce;                                           /* [SPEC] reciprocal majorant */
            m_old = (ce != 0.f) ? vmul(m_old, vscale(mo.sigma_n, inv_ce)) : vmul(m_old, mo.sigma_n);   /* :447-449 */
            m_new = (ce != 0.f) ? vmul(m_new, vscale(mn.sigma_n, inv_ce)) : vmul(m_new, mn.sigma_n);   /* :450-452 */
            if (adjoint && (m_new.x > 0.f)) {                                  /* prbvolpath sample_emitter, Backward */
                float k = (-inv_ce) / m_new.x, g = 0.f;
                g += adjoint->x * k; g += adjoint->y * k; g += adjoint->z * k;
                scatter_sigma_t(S, Mn, G, mn.p, g, cnt);
            }
        }
        active_surface = active_surface || escaped_medium;                     /* :455 */
        active_surface = active_surface && si.valid && !active_medium;
        if (active_surface) {                                                  /* :457-460 eval_null_transmission */
            float bsdf_val = (si.shape < 0) ? 1.f : 0.f;                       /* null BSDF: 1; diffuse: 0 */
            m_old = vscale(m_old, bsdf_val); m_new = vscale(m_new, bsdf_val);
        }
        tr_old = vmul(tr_old, m_old);                                          /* :462 */
        tr_new = vmul(tr_new, m_new);
        if (active_surface) {                                                  /* :466-467 */
            ray_t nr = spawn_ray(si.p, si.n, ray.d);
            ray = nr;
        }
        ray.maxt = remaining_dist;                                             /* :468 */
        needs_intersection = needs_intersection || active_surface;
        active = active && (active_medium || active_surface) && (any_nz(tr_new) || any_nz(tr_old)); /* :472 */
        if (active) total_dist += active_medium ? mn.t : si.t;                 /* :473 */
        if (active_surface && si.shape < 0) {                                  /* :476-477 si.is_medium_transition() */
            float dn = fmaf(si.n.x, ray.d.x, fmaf(si.n.y, ray.d.y, si.n.z * ray.d.z));
            medium = (dn > 0.f) ? 0 : 1;                                       /* target_medium(d): exterior / interior */
        }
    }
    *emitted_old = vmul(emitter_val, tr_old);                                  /* :479 */
    *emitted_new = vmul(emitter_val, tr_new);
    *ds_d = d; *ds_pdf = pdf;
}

/* ------------------------------------------------------------------------- */
/* sample_twostates  (plugins/twostatevolpath.py:141-385)                     */
/* adj != NULL: upstream prbvolpath's sample() in Backward mode on state `new` (call with Mo == Mn), see prb_adj      */
/* ------------------------------------------------------------------------- */
static void sample_twostates(const rcv_scene_desc *S, const orc_medium *Mo, const orc_medium *Mn,
                             pcg32 *rng, ray_t ray, float L_old_out[3], float L_new_out[3],
                             int *valid_out, counters_t *cnt, const prb_adj *adj) {
    int depth = 0;
    v3 L_old = V3(0, 0, 0), L_new = adj ? adj->L_total : V3(0, 0, 0);
    v3 thr_old = V3(1, 1, 1), thr_new = V3(1, 1, 1);
    float eta = 1.f;
    int active = 1;
    si_t si; si.valid = 0; si.t = 0.f; si.p = V3(0, 0, 0); si.n = V3(0, 0, 0); si.shape = -1;
    int needs_intersection = 1;
    v3 last_scatter_p = V3(0, 0, 0);
    float last_scatter_direction_pdf = 1.f;
    int medium = 0;                         /* MediumPtr: 0 = None, 1 = the TwoStateMedium */
    int valid_ray = 0, specular_chain = 1;
    /* :171-173 channel = min(3 * next_1d, 2); a float that index_spectrum never
     * matches against an integer, so channel 0 is always used */
    (void) next_1d(rng, active);

    while (active) {                                                           /* :180 */
        active = active && (any_nz(thr_new) || any_nz(thr_old));               /* :182 */
        float q = fminf_(max3(vmax(thr_new, thr_old)) * (eta * eta), 0.99f);   /* :183 */
        int perform_rr = depth > S->rr_depth;                                  /* :184 */
        float rr = next_1d(rng, active);                                       /* :185 */
        active = active && ((rr < q) || !perform_rr);
        if (perform_rr) {                                                      /* :187-188 */
            float rq = 1.f / q;
            thr_new = vscale(thr_new, rq); thr_old = vscale(thr_old, rq);
        }
        int active_medium = active && (medium != 0);                           /* :190 */
        int active_surface = active && !active_medium;

        float u = next_1d(rng, active_medium);                                 /* :194 */
        mei_t mo, mn;
        sample_interaction_twostates(S, Mo, Mn, &ray, u, active_medium, &mo, &mn, cnt);   /* :196 */
        int intersect = needs_intersection && active_medium;                   /* :200 */
        if (intersect) si = scene_intersect_all(S, &ray);
        needs_intersection = needs_intersection && !active_medium;             /* :204 */
        if (active_medium && (si.t < mn.t)) mn.t = ORC_INF;                    /* :205 */
        if (active_medium && (si.t < mo.t)) mo.t = ORC_INF;

        v3 weight_new = V3(1, 1, 1), weight_old = V3(1, 1, 1);
        int escaped_medium = active_medium && !mei_valid(&mn);                 /* :216 */
        active_medium = active_medium && mei_valid(&mn);
        /* handle_null_scattering branch (:219-231); heterogeneous media always take it (:99) */
        float majorant = mn.combined_extinction;
        float inv_maj = 1.f / majorant;                                        /* [SPEC] reciprocal majorant */
        float sp_new = mn.sigma_t.x * inv_maj, sp_old = mo.sigma_t.x * inv_maj;
        float scatter_prob = (majorant != 0.f) ? (sp_new + sp_old) * 0.5f : 0.f;   /* :225 */
        float xi = next_1d(rng, active_medium);                                /* :227 */
        int act_null_scatter = (xi >= scatter_prob) && active_medium;
        int act_medium_scatter = !act_null_scatter && active_medium;
        if (act_null_scatter) {                                                /* :230-231 */
            v3 den = vadd(mn.sigma_n, mo.sigma_n);
            v3 rden = vdiv(V3(1, 1, 1), den);                                  /* [SPEC] a / den == a * (1 / den) */
            weight_new = vmul(weight_new, vmul(vscale(mn.sigma_n, 2.f), rden));
            weight_old = vmul(weight_old, vmul(vscale(mo.sigma_n, 2.f), rden));
        }
        if (act_medium_scatter) { depth += 1; last_scatter_p = mn.p; }         /* :249-250 */
        active = active && (depth < S->max_depth);                             /* :253 */
        act_medium_scatter = act_medium_scatter && active;
        if (act_null_scatter) {                                                /* :256-257 */
            ray.o = mn.p;
            si.t = si.t - mn.t;
        }
        if (act_medium_scatter) {                                              /* :259-260 */
            v3 den = vadd(mn.sigma_t, mo.sigma_t);
            v3 rden = vdiv(V3(1, 1, 1), den);                                  /* [SPEC] reciprocal */
            weight_new = vmul(weight_new, vmul(vscale(mn.sigma_s, 2.f), rden));
            weight_old = vmul(weight_old, vmul(vscale(mo.sigma_s, 2.f), rden));
            cnt->c[1]++;
        }
        if (active_medium) {                                                   /* :265-266 */
            thr_new = vmul(thr_new, weight_new);
            thr_old = vmul(thr_old, weight_old);
        }
        if (adj && active_medium) {
            /* prbvolpath Backward: dr.backward(dL * weight * L / max(1e-8, weight)) with weight = sigma_n * c (null) or
             * sigma_t * albedo * c (real), c detached: dL * L * d log(weight) [UPSTREAM-RECALLED]; L = L_new is the
             * radiance still to come */
            float aw = (adj->dL.x * L_new.x + adj->dL.y * L_new.y) + adj->dL.z * L_new.z;
            if (act_null_scatter) {
                float sn = mn.sigma_n.x;
                if (sn > 0.f) scatter_sigma_t(S, Mn, adj->G, mn.p, -(aw * (1.f / sn)), cnt);
            } else if (act_medium_scatter) {
                float st = mn.sigma_t.x;
                v3 al = mn.albedo;
                v3 Lo = vdiv(L_new, vmax(V3(1e-8f, 1e-8f, 1e-8f), al));
                scatter_sigma_t(S, Mn, adj->G, mn.p, aw * (1.f / st), cnt);
                scatter_albedo(S, Mn, adj->G, mn.p, vmul(adj->dL, Lo), cnt);
            }
        }
        valid_ray = valid_ray || act_medium_scatter;                           /* :274 */
        float s1 = next_1d(rng, act_medium_scatter);                           /* :275 */
        float s2x = next_1d(rng, act_medium_scatter), s2y = next_1d(rng, act_medium_scatter);
        v3 wi = V3(-ray.d.x, -ray.d.y, -ray.d.z);
        v3 wo = V3(0, 0, 0); float phase_pdf = 0.f;
        if (act_medium_scatter) wo = phase_sample(S, wi, s1, s2x, s2y, &phase_pdf);
        act_medium_scatter = act_medium_scatter && (phase_pdf > 0.f);          /* :276 */
        v3 scatter_wi = wi;
        if (act_medium_scatter) {                                              /* :278-281 */
            ray = spawn_ray(mn.p, V3(0, 0, 0), wo);
            needs_intersection = 1;
            last_scatter_direction_pdf = phase_pdf;
        }

        /* ---- surface interactions (:283-286) */
        active_surface = active_surface || escaped_medium;
        intersect = active_surface && needs_intersection;
        if (intersect) si = scene_intersect_all(S, &ray);

        /* ---- intersection with emitters (:288-308): the `area` light's rectangle, or the environment for rays that
         * hit nothing [UPSTREAM si.emitter(scene)] */
        int ray_from_camera = active_surface && (depth == 0);
        int count_direct = ray_from_camera || specular_chain;
        int has_emitter = si.valid ? (S->emitter_type == RCV_EMITTER_AREA && si.shape == S->area_quad)
                                   : (S->emitter_type != RCV_EMITTER_AREA);
        int active_e = active_surface && has_emitter && !((depth == 0) && S->hide_emitters);
        if (active_e) {
            float emitter_pdf_; v3 emitted;
            if (si.valid) {                                                    /* area.cpp pdf_direction / eval */
                emitter_pdf_ = S->use_nee ? area_pdf(S, last_scatter_p, si.p) : 0.f;
                emitted = (-dot3(si.n, ray.d) > 0.f) ? g_quads[si.shape].rad : V3(0, 0, 0);
            } else {
                emitter_pdf_ = S->use_nee ? emitter_pdf(S, ray.d) : 0.f;       /* :297 pdf_emitter_direction */
                emitted = emitter_eval(S, ray.d);                              /* :300 */
            }
            float mis_bsdf = mis_weight(last_scatter_direction_pdf, emitter_pdf_);         /* :302 */
            v3 c_new = count_direct ? vmul(thr_new, emitted) : vmul(vscale(thr_new, mis_bsdf), emitted);   /* :304-308 */
            v3 c_old = count_direct ? vmul(thr_old, emitted) : vmul(vscale(thr_old, mis_bsdf), emitted);
            L_new = adj ? vsub(L_new, c_new) : vadd(L_new, c_new);
            L_old = vadd(L_old, c_old);
        }
        active_surface = active_surface && si.valid;                           /* :310 */

        /* ---- emitter sampling (:315-344) */
        if (S->use_nee) {
            int smooth = active_surface && (si.shape >= 0);                    /* diffuse: Smooth; null: not (:316) */
            int active_e_surface = smooth && (depth + 1 < S->max_depth);
            int sample_emitters = S->sample_emitters;                          /* :317 */
            specular_chain = specular_chain && !act_medium_scatter;            /* :318 */
            specular_chain = specular_chain || (act_medium_scatter && !sample_emitters);
            int active_e_medium = act_medium_scatter && sample_emitters;       /* :320 */
            int active_e2 = active_e_surface || active_e_medium;
            v3 ref_p = V3(0, 0, 0), ref_n = V3(0, 0, 0);                       /* :322-324 */
            if (act_medium_scatter) { ref_p = mn.p; }
            if (active_surface) { ref_p = si.p; ref_n = si.n; }
            v3 em_old, em_new, ds_d; float ds_pdf;
            pcg32 nee_sampler = *rng;                                          /* prbvolpath: sampler.clone() (Backward) */
            sample_emitter_twostates(S, Mo, Mn, rng, ref_p, ref_n, medium, active_e2,
                                     &em_old, &em_new, &ds_d, &ds_pdf, NULL, NULL, cnt);   /* :325 */
            v3 bsdf_val = V3(0, 0, 0); float bsdf_pdf = 0.f;                   /* :329-330 eval_pdf; eval_old == eval */
            if (active_e_surface) diffuse_eval_pdf(&g_quads[si.shape], ray.d, ds_d, &bsdf_val, &bsdf_pdf);
            float phase_val_new = active_e_medium ? phase_eval(S, scatter_wi, ds_d) : 0.f;   /* :333 */
            float phase_val_old = active_e_medium ? phase_eval(S, scatter_wi, ds_d) : 0.f;   /* :334 */
            v3 nee_weight_new = active_e_surface ? bsdf_val : V3(phase_val_new, phase_val_new, phase_val_new);   /* :336-337 */
            v3 nee_weight_old = active_e_surface ? bsdf_val : V3(phase_val_old, phase_val_old, phase_val_old);
            float nee_directional_pdf = active_e_surface ? bsdf_pdf : phase_val_new;       /* :339 (ds.delta is false) */
            float mis_em = mis_weight(ds_pdf, nee_directional_pdf);            /* :341 */
            if (active) {                                                      /* :343-344 */
                v3 c_new = vmul(vscale(vmul(thr_new, nee_weight_new), mis_em), em_new);
                v3 c_old = vmul(vscale(vmul(thr_old, nee_weight_old), mis_em), em_old);
                L_new = adj ? vsub(L_new, c_new) : vadd(L_new, c_new);
                L_old = vadd(L_old, c_old);
                if (adj) {                                                     /* the shadow ray again, with gradients */
                    v3 adjoint = vmul(adj->dL, c_new), e1, e2, d2; float p2;
                    sample_emitter_twostates(S, Mo, Mn, &nee_sampler, ref_p, ref_n, medium, active_e2, &e1, &e2, &d2, &p2,
                                             &adjoint, adj->G, cnt);
                }
            }
        }

        /* ---- BSDF sampling (:346-379).  `null`: wo = -wi, pdf = 1, weight = 1, sampled_type = Null (a Delta type)
         * [UPSTREAM null.cpp]; `diffuse`: cosine-weighted wo, weight = reflectance [UPSTREAM diffuse.cpp] */
        float b1 = next_1d(rng, active_surface);                               /* :347 */
        float b2x = next_1d(rng, active_surface), b2y = next_1d(rng, active_surface);   /* :348 next_2d */
        (void) b1;
        int is_null = active_surface && (si.shape < 0);
        v3 bs_wo = ray.d; float bs_pdf = 1.f;
        if (active_surface && !is_null) {
            const rquad_t *qd = &g_quads[si.shape];
            float ci = -dot3(qd->n, ray.d);
            v3 wl = square_to_cosine_hemisphere(b2x, b2y);
            float pdf = ORC_INV_PI * wl.z;
            int ok = (ci > 0.f) && (pdf > 0.f);
            bs_pdf = ok ? pdf : 0.f;
            active_surface = active_surface && (bs_pdf > 0.f);                 /* :349 */
            if (active_surface) {
                v3 val_old = vscale(vscale(qd->refl, ORC_INV_PI), wl.z);       /* :351 eval_old(bs.wo) */
                v3 w_old = V3(val_old.x / bs_pdf, val_old.y / bs_pdf, val_old.z / bs_pdf);   /* :354 */
                thr_old = vmul(thr_old, w_old);                                /* :355 (not delta) */
                thr_new = vmul(thr_new, qd->refl);                             /* :357 */
                bs_wo = V3(fmaf(qd->s.x, wl.x, fmaf(qd->t.x, wl.y, qd->n.x * wl.z)),
                           fmaf(qd->s.y, wl.x, fmaf(qd->t.y, wl.y, qd->n.y * wl.z)),
                           fmaf(qd->s.z, wl.x, fmaf(qd->t.z, wl.y, qd->n.z * wl.z)));   /* si.to_world(bs.wo) */
            }
        }
        /* null: bs.pdf = 1 > 0 (:349); prev_bsdf_delta = True -> throughput_old *= bsdf_weight = 1 (:355-357) */
        if (act_medium_scatter) {                                              /* :360-361 */
            float phase_val = phase_eval(S, scatter_wi, wo);
            float ratio = phase_val / phase_pdf;
            thr_old = vscale(thr_old, ratio);
        }
        if (active_surface) {                                                  /* :363-367 */
            ray = spawn_ray(si.p, si.n, bs_wo);   /* null: si.to_world(-si.wi) == ray.d */
            needs_intersection = 1;
            if (!is_null) {                                                    /* non_null_bsdf (:368-377) */
                depth += 1;
                last_scatter_p = si.p;
                last_scatter_direction_pdf = bs_pdf;
                valid_ray = 1;
                specular_chain = 0;                                            /* a Smooth, non-Delta lobe was sampled */
            } else {
                float dn = fmaf(si.n.x, ray.d.x, fmaf(si.n.y, ray.d.y, si.n.z * ray.d.z));
                medium = (dn > 0.f) ? 0 : 1;                                   /* :378-379 */
            }
        }
        active = active && (active_surface || active_medium);                  /* :380 */
    }
    L_old_out[0] = L_old.x; L_old_out[1] = L_old.y; L_old_out[2] = L_old.z;
    L_new_out[0] = L_new.x; L_new_out[1] = L_new.y; L_new_out[2] = L_new.z;
    *valid_out = valid_ray;
}

/* ------------------------------------------------------------------------- */
/* render_twostates  (plugins/twostatevolpath.py:91-138)                      */
/* ------------------------------------------------------------------------- */
int orc_twostates_lanes(const rcv_scene_desc *S, const orc_medium *Mo, const orc_medium *Mn,
                        uint32_t seed, uint32_t spp, uint64_t lane_begin, uint64_t lane_end,
                        float *L_old, float *L_new, int32_t *pix, int nthreads) {
    env_build(S);
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
    {
        counters_t cnt; memset(&cnt, 0, sizeof(cnt));
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = (int64_t) lane_begin; i < (int64_t) lane_end; ++i) {
            lane_t ln = lane_setup(S, seed, spp, (uint64_t) i);                /* :104-112 */
            int valid;
            size_t k = (size_t) (i - (int64_t) lane_begin);
            sample_twostates(S, Mo, Mn, &ln.rng, ln.ray, L_old + 3 * k, L_new + 3 * k, &valid, &cnt, NULL);   /* :115 */
            if (pix) pix[k] = ln.pix;
            cnt.c[6]++;
        }
        counters_merge(&cnt);
    }
    return 0;
}

/* film accumulation for the sample sub-range [spp_begin, spp_end) of every
 * pixel: ImageBlock.put with a box filter = += (L*weight, weight) at
 * floor(pos) (twostatevolpath.py:80-89, weight = 1 in RGB mode) */
int orc_render_twostates(const rcv_scene_desc *S, const orc_medium *Mo, const orc_medium *Mn,
                         uint32_t seed, uint32_t spp, uint32_t spp_begin, uint32_t spp_end,
                         float *accum_old, float *accum_new, int nthreads) {
    env_build(S);
    uint64_t npix = (uint64_t) S->film_width * S->film_height;
    uint32_t ns = spp_end - spp_begin;
    const uint64_t chunk = 1u << 16;
    float *Lo = (float *) malloc(chunk * ns * 3 * sizeof(float));
    float *Ln = (float *) malloc(chunk * ns * 3 * sizeof(float));
    float *px = (float *) malloc(chunk * ns * 2 * sizeof(float));     /* film positions */
    if (!Lo || !Ln || !px) return 1;
    if (nthreads < 1) nthreads = 1;
    for (uint64_t p0 = 0; p0 < npix; p0 += chunk) {
        uint64_t p1 = p0 + chunk < npix ? p0 + chunk : npix;
#pragma omp parallel num_threads(nthreads)
        {
            counters_t cnt; memset(&cnt, 0, sizeof(cnt));
#pragma omp for schedule(dynamic, 64)
            for (int64_t p = (int64_t) p0; p < (int64_t) p1; ++p) {
                for (uint32_t k = spp_begin; k < spp_end; ++k) {
                    uint64_t idx = (uint64_t) p * spp + k;
                    size_t o = ((size_t) (p - (int64_t) p0) * ns + (k - spp_begin));
                    lane_t ln = lane_setup(S, seed, spp, idx);
                    int valid;
                    sample_twostates(S, Mo, Mn, &ln.rng, ln.ray, Lo + 3 * o, Ln + 3 * o, &valid, &cnt, NULL);
                    px[2 * o] = ln.fx; px[2 * o + 1] = ln.fy;
                    cnt.c[6]++;
                }
            }
            counters_merge(&cnt);
        }
        size_t n = (size_t) (p1 - p0) * ns;
        for (size_t o = 0; o < n; ++o) {            /* serial splat: deterministic summation order */
            footprint_t fp; film_footprint(S, px[2 * o], px[2 * o + 1], &fp);
            for (int q = 0; q < fp.n; ++q) {
                float *ao = accum_old + 4 * (size_t) fp.pix[q], *an = accum_new + 4 * (size_t) fp.pix[q];
                float w = fp.w[q];
                ao[0] += Lo[3 * o] * w; ao[1] += Lo[3 * o + 1] * w; ao[2] += Lo[3 * o + 2] * w; ao[3] += w;
                an[0] += Ln[3 * o] * w; an[1] += Ln[3 * o + 1] * w; an[2] += Ln[3 * o + 2] * w; an[3] += w;
            }
        }
    }
    free(Lo); free(Ln); free(px);
    return 0;
}

/* hdrfilm.develop(): RGB / W, weight 0 -> divide by 1 [UPSTREAM hdrfilm.cpp] */
void orc_develop(const float *accum, float *img, uint64_t npix) {
    for (uint64_t p = 0; p < npix; ++p) {
        float w = accum[4 * p + 3];
        float d = (w == 0.f) ? 1.f : w;
        img[3 * p] = accum[4 * p] / d; img[3 * p + 1] = accum[4 * p + 1] / d; img[3 * p + 2] = accum[4 * p + 2] / d;
    }
}

/* per-pixel film weight (number of samples splatted into each pixel) for all
 * spp samples: needed by the adjoint to differentiate the weight division */
void orc_film_weights(const rcv_scene_desc *S, uint32_t seed, uint32_t spp, float *Wt) {
    uint64_t npix = (uint64_t) S->film_width * S->film_height;
    memset(Wt, 0, npix * sizeof(float));
    orc_film_weights_range(S, seed, spp, 0, spp, Wt);
}
/* the weights of samples [spp_begin, spp_end) of every pixel, ADDED to Wt (one rank's piece of an spp-sharded job) */
void orc_film_weights_range(const rcv_scene_desc *S, uint32_t seed, uint32_t spp, uint32_t spp_begin, uint32_t spp_end, float *Wt) {
    uint64_t npix = (uint64_t) S->film_width * S->film_height;
    for (uint64_t p = 0; p < npix; ++p)
        for (uint32_t k = spp_begin; k < spp_end; ++k) {
            lane_t ln = lane_setup(S, seed, spp, p * spp + k);
            footprint_t fp; film_footprint(S, ln.fx, ln.fy, &fp);
            for (int q = 0; q < fp.n; ++q) Wt[fp.pix[q]] += fp.w[q];
        }
}

/* ------------------------------------------------------------------------- */
/* CVIntegrator.render cv update (plugins/cv_integrator.py:58-77,             */
/* plugins/metaintegrator.py:93-113) + StatisticsEstimator (welford.py:38-78) */
/* ------------------------------------------------------------------------- */
/* bias-correction terms 1 - beta**n: Dr.Jit evaluates TensorXf ** TensorXi in
 * float32; this build fixes them as float32(1 - beta^n) computed in double
 * on the host and hands them to oracle and kernel alike. */
static float one_minus_pow(float beta, int n) {
    return (float) (1.0 - pow((double) beta, (double) n));
}

void orc_cv_update(const float *H, const float *F, float *img, float *v_n, float *w_s,
                   float *mean_x, float *mean_y, float *var_x, float *var_y, float *cov,
                   uint64_t n, int it, int warmup, int stats_n, float beta1, float beta2,
                   int pre_update) {
    int sn = stats_n;
    for (int phase = 0; phase < 3; ++phase) {
        int do_update = (it > 0) && ((phase == 0 && pre_update) || (phase == 2 && !pre_update));
        if (phase == 1) {
            /* cv_integrator.py:61-70 */
            float c2 = one_minus_pow(beta2, sn);
            for (uint64_t i = 0; i < n; ++i) {
                float ws = w_s[i];
                if (it > warmup) {
                    float v0 = var_x[i] / c2, v1 = var_y[i] / c2, cv = cov[i] / c2;    /* welford.py:77-78 */
                    float vn = ((ws * ws) * (v_n[i] + v0) + v1) - ((2.f * ws) * cv);   /* :63 */
                    v_n[i] = vn;
                    ws = cv / (v0 + vn);                                               /* :65 */
                    if (isnan(ws) || isinf(ws)) ws = 0.f;                              /* :67 */
                    ws = fminf_(fmaxf_(ws, 0.f), 1.f);                                 /* :68 */
                    w_s[i] = ws;
                }
                img[i] = ws * (img[i] - H[i]) + F[i];                                  /* :70 */
            }
        } else if (do_update) {
            /* welford.py:59-75, x = H, y = F */
            float c1a = one_minus_pow(beta1, sn), c1b = one_minus_pow(beta1, sn + 1);
            for (uint64_t i = 0; i < n; ++i) {
                float x = H[i], y = F[i];
                float dx1 = (sn > 0) ? x - mean_x[i] / c1a : x;                        /* :61 */
                float dy1 = (sn > 0) ? y - mean_y[i] / c1a : y;                        /* :62 */
                mean_y[i] = beta1 * mean_y[i] + (1.f - beta1) * y;                     /* :64 */
                mean_x[i] = beta1 * mean_x[i] + (1.f - beta1) * x;                     /* :65 */
                float dx2 = x - mean_x[i] / c1b;                                       /* :68 */
                float dy2 = y - mean_y[i] / c1b;                                       /* :69 */
                var_x[i] = beta2 * var_x[i] + ((1.f - beta2) * dx1) * dx2;             /* :71 */
                var_y[i] = beta2 * var_y[i] + ((1.f - beta2) * dy1) * dy2;             /* :72 */
                cov[i] = beta2 * cov[i] + ((1.f - beta2) * dx1) * dy2;                 /* :73 */
            }
            sn += 1;                                                                   /* :67 */
        }
    }
}

/* utils.py:21-25 and the gradient dr.backward(loss) seeds into the image */
double orc_loss_grad(int kind, const float *img, const float *ref, uint64_t n, float *grad) {
    double acc = 0.0;
    float inv_n = 1.f / (float) n;
    for (uint64_t i = 0; i < n; ++i) {
        float d = img[i] - ref[i];
        if (kind == RCV_LOSS_L1) {
            acc += fabs((double) d);
            if (grad) grad[i] = ((d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f)) * inv_n;
        } else {
            acc += (double) d * (double) d;
            if (grad) grad[i] = (2.f * d) * inv_n;
        }
    }
    return acc / (double) n;
}

/* ------------------------------------------------------------------------- */
/* VolpathSimpleIntegrator  (plugins/volpathsimple.py), use_drt = False        */
/* ------------------------------------------------------------------------- */
/* Medium::sample_interaction [UPSTREAM medium.cpp], single state (state `new`,
 * twostatemedium.py:135-136), global majorant */
static void sample_interaction(const rcv_scene_desc *S, const orc_medium *M, const ray_t *ray, float sample,
                               int active, mei_t *mei, counters_t *cnt) {
    float mint, maxt;
    active = prepare_interaction_sampling(S, ray, active, &mint, &maxt);
    float desired_tau = -orc_logf(1.f - sample);
    float m = M->majorant, sampled_t, m_cell = 0.f;
    if (M->majorant_grid) {
        sampled_t = active ? dda_sample(S, NULL, M, ray, mint, maxt, desired_tau, &m_cell) : ORC_INF;
    } else {
        float inv_m = 1.f / m;                       /* [SPEC] reciprocal majorant */
        sampled_t = mint + (desired_tau * inv_m);
    }
    int valid = active && (sampled_t <= maxt);
    mei->t = valid ? sampled_t : ORC_INF;
    mei->p = ray_at(ray->o, ray->d, sampled_t);
    mei->mint = mint;
    if (M->majorant_grid && valid) m = m_cell;                      /* the local majorant of the stopping cell */
    v3 st = V3(0, 0, 0), ss = V3(0, 0, 0);
    if (valid) {
        tri_t T = tri_setup(S, M, mei->p);
        float s = lookup_sigma_t(S, M, &T);
        v3 a = lookup_albedo(M, &T);
        st = V3(s, s, s); ss = vmul(st, a);
        cnt->c[2]++;
    }
    float ce = valid ? m : 0.f;                      /* get_majorant(mei, valid_mei) */
    mei->sigma_t = st; mei->sigma_s = ss;
    mei->sigma_n = valid ? vsub(V3(m, m, m), st) : V3(0, 0, 0);
    mei->combined_extinction = ce;
}

/* volpathsimple.py:336-390 */
static void sample_real_interaction(const rcv_scene_desc *S, const orc_medium *M, const ray_t *ray, pcg32 *rng,
                                    int _active, mei_t *mei, counters_t *cnt) {
    ray_t running_ray = *ray;                                                  /* :349 */
    float running_t = 0.f;                                                     /* :351 */
    int active = _active;
    memset(mei, 0, sizeof(*mei));
    mei->t = active ? __builtin_nanf("") : ORC_INF;                            /* :356 */
    while (active) {                                                           /* :360 */
        mei_t mei_next;
        float u = next_1d(rng, active);
        sample_interaction(S, M, &running_ray, u, active, &mei_next, cnt);     /* :361 */
        *mei = mei_next;                                                       /* :363 */
        mei->t = mei->t + running_t;                                           /* :364 */
        float majorant = mei_next.combined_extinction;                         /* :366 */
        float r = (majorant != 0.f) ? mei_next.sigma_t.x * (1.f / majorant) : 0.f;   /* :367 [SPEC] reciprocal */
        active = active && mei_valid(&mei_next);                               /* :371 */
        float xi = next_1d(rng, active);                                       /* :372 */
        int did_null_scatter = active && (xi >= r);
        active = active && did_null_scatter;                                   /* :374 */
        float next_t = mei_next.t;                                             /* :377 */
        if (active) {                                                          /* :378-380 */
            running_ray.o = vadd(running_ray.o, vscale(running_ray.d, next_t));
            running_ray.maxt = running_ray.maxt - next_t;
            running_t = running_t + next_t;
        }
    }
    int did_sample = _active && mei_valid(mei);                                /* :383 */
    if (did_sample) {
        mei->p = ray_at(ray->o, ray->d, mei->t);                               /* :384 */
        /* :386-388 coefficients re-fetched at the final point */
        tri_t T = tri_setup(S, M, mei->p);
        float s = lookup_sigma_t(S, M, &T);
        v3 a = lookup_albedo(M, &T);
        mei->sigma_t = V3(s, s, s); mei->sigma_s = vmul(mei->sigma_t, a);
        { float mj = M->majorant_grid ? majorant_at(S, M, mei->p) : M->majorant;
          mei->sigma_n = vsub(V3(mj, mj, mj), mei->sigma_t); }
        cnt->c[3]++;
    }
}

/* volpathsimple.py:449-517 (ratio tracking; adjoint != NULL replays with gradients) */
static v3 estimate_transmittance(const rcv_scene_desc *S, const orc_medium *M, const ray_t *ray_full, float tmin,
                                 float tmax, pcg32 *rng, int active, const v3 *adjoint, grad_sink *G,
                                 counters_t *cnt) {
    if (tmax < tmin) { float tmp = tmin; tmin = tmax; tmax = tmp; }            /* :459-462 */
    ray_t ray = *ray_full;
    ray.o = ray_at(ray_full->o, ray_full->d, tmin);                            /* :467 */
    tmax = tmax - tmin;
    ray.maxt = tmax;
    float one = active ? 1.f : 0.f;
    v3 transmittance = V3(one, one, one);                                      /* :472 */
    while (active) {                                                           /* :478 */
        mei_t mei;
        float u = next_1d(rng, active);
        sample_interaction(S, M, &ray, u, active, &mei, cnt);                  /* :482 */
        float ce = mei.combined_extinction;
        float inv_ce = 1.f / ce;                                               /* [SPEC] reciprocal majorant */
        v3 tr = (ce != 0.f) ? vscale(mei.sigma_n, inv_ce) : mei.sigma_n;       /* :486-489 */
        if (active && (mei.t > tmax)) mei.t = ORC_INF;                         /* :493 */
        active = active && mei_valid(&mei);
        if (adjoint && active) {                                               /* :496-505 */
            /* d/d sigma_t of adjoint * tr / detach(tr), tr = (mu - sigma_t)/mu, summed over channels */
            float g = 0.f;
            if (tr.x > 0.f) g += adjoint->x * ((-inv_ce) / tr.x);
            if (tr.y > 0.f) g += adjoint->y * ((-inv_ce) / tr.y);
            if (tr.z > 0.f) g += adjoint->z * ((-inv_ce) / tr.z);
            scatter_sigma_t(S, M, G, mei.p, g, cnt);
        }
        if (active) {
            transmittance = vmul(transmittance, tr);                           /* :508 */
            ray.o = mei.p;                                                     /* :510 */
            tmax = tmax - mei.t;
            ray.maxt = tmax;
        }
        active = active && any_nz(transmittance);                              /* :515 */
    }
    return transmittance;
}

/* volpathsimple.py:419-446 */
static v3 vps_sample_emitter(const rcv_scene_desc *S, const orc_medium *M, v3 ref_p, pcg32 *rng, int active,
                             const v3 *adjoint, grad_sink *G, v3 *ds_d, float *ds_pdf, counters_t *cnt) {
    float s2x = next_1d(rng, active), s2y = next_1d(rng, active);              /* :431 */
    v3 d = V3(0, 0, 0); float pdf = 0.f; v3 emitter_val = V3(0, 0, 0);
    if (active) emitter_sample_env(S, s2x, s2y, &d, &pdf, &emitter_val);       /* :432 `constant` or `envmap` */
    int sampling_worked = (pdf != 0.f);                                        /* :434-436 */
    if (!sampling_worked) emitter_val = V3(0, 0, 0);
    active = active && sampling_worked;
    ray_t ray = spawn_ray(ref_p, V3(0, 0, 0), d);                              /* :440 */
    si_t si; si.valid = 0; si.t = ORC_INF;
    if (active) si = scene_intersect(S, &ray);                                 /* :441 */
    ray.maxt = si.t;
    v3 tr = estimate_transmittance(S, M, &ray, 0.f, si.t, rng, active && si.valid, adjoint, G, cnt);   /* :443 */
    *ds_d = d; *ds_pdf = pdf;
    return vmul(emitter_val, tr);                                              /* :446 */
}

/* volpathsimple.py:393-416 */
static v3 sample_emitter_for_nee(const rcv_scene_desc *S, const orc_medium *M, v3 mei_p, v3 wi, pcg32 *rng,
                                 v3 throughput, int active, int primal, v3 dL, grad_sink *G, counters_t *cnt) {
    pcg32 nee_sampler = *rng;                                                  /* :396 sampler.clone() */
    v3 ds_d; float ds_pdf;
    v3 emitted = vps_sample_emitter(S, M, mei_p, rng, active, NULL, NULL, &ds_d, &ds_pdf, cnt);   /* :398 */
    float phase_val = active ? phase_eval(S, wi, ds_d) : 0.f;                  /* :401 */
    float phase_pdf = phase_val;
    v3 nee_contrib = vmul(vscale(vscale(throughput, phase_val), mis_weight(ds_pdf, phase_pdf)), emitted);  /* :404 */
    if (!primal) {                                                             /* :406-414 */
        v3 adjoint = vmul(dL, nee_contrib);
        v3 d2; float p2;
        (void) vps_sample_emitter(S, M, mei_p, &nee_sampler, active, &adjoint, G, &d2, &p2, cnt);
    }
    return nee_contrib;
}

/* volpathsimple.py:597-620 */
static void backpropagate_transmittance(const rcv_scene_desc *S, const orc_medium *M, pcg32 *alt, float si_t_,
                                        float mei_t_, const ray_t *ray, v3 adj_weight, int did_scatter,
                                        int did_escape, grad_sink *G, counters_t *cnt) {
    int active = did_scatter || did_escape;                                    /* :600 */
    float interval = did_escape ? si_t_ : mei_t_;                              /* :601 */
    float inv_pdf = interval / 4.f;                                            /* :617 */
    for (int j = 0; j < 4; ++j) {                                              /* :607 */
        float u = next_1d(alt, active);                                        /* :608 */
        if (!active) continue;
        float t = u * interval;
        v3 p = ray_at(ray->o, ray->d, t);                                      /* :609 */
        /* contribs -= sigma_t(p) (:611-613); backward of adj_weight * contribs * inv_pdf (:620) */
        float g = -((adj_weight.x * inv_pdf) + (adj_weight.y * inv_pdf) + (adj_weight.z * inv_pdf));
        scatter_sigma_t(S, M, G, p, g, cnt);
    }
}

/* ------------------------------------------------------------------------- */
/* DRT (differential ratio tracking) in-scattering term, volpathsimple.py:520-594,623-778.         */
/* ORACLE ONLY so far: the CUDA path implements use_drt = False.                                                      */
/* ------------------------------------------------------------------------- */
typedef struct { int use_drt, use_drt_subsampling, use_drt_mis; } drt_opts;

/* PathState / DRTPathState / DRTReservoir with n = 1 (:673-778) */
typedef struct {
    v3 wsum, current_weight;
    int depth;              /* -1: never set */
    si_t si; ray_t ray; int active;
} drt_reservoir;
static inline float mean3(v3 a) { return ((a.x + a.y) + a.z) / 3.f; }

static void reservoir_update(drt_reservoir *R, int depth, si_t si, ray_t ray, v3 weight, float sample, int active) {   /* :757-765 */
    if (!active) return;
    R->wsum = vadd(R->wsum, weight);
    float ratio = mean3(vdiv(weight, R->wsum));
    if (sample <= ratio) { R->current_weight = weight; R->depth = depth; R->si = si; R->ray = ray; R->active = active; }
}
static v3 reservoir_weight(const drt_reservoir *R) {                                                /* :768-772 */
    float d = mean3(R->current_weight);
    if (d != 0.f) return vscale(R->current_weight, mean3(R->wsum) / d);
    return V3(0, 0, 0);
}

/* Medium::sample_interaction_drt [UPSTREAM, absent]: restated from differential ratio tracking (Nimier-David et al. 2022):
 * tentative collisions t_i with the majorant; w_i = prod_{j<i} (1 - sigma_t(t_j) / mu) estimates T(t_i); one of them is kept
 * with probability proportional to w_i (reservoir), weight = (sum_i w_i) / mu.  [SPEC] two draws per tentative collision
 * (distance, reservoir), global majorant. */
static float sample_interaction_drt(const rcv_scene_desc *S, const orc_medium *M, const ray_t *ray, pcg32 *alt, int active,
                                    mei_t *mei, counters_t *cnt) {
    float mint, maxt;
    active = prepare_interaction_sampling(S, ray, active, &mint, &maxt);
    mei->t = ORC_INF; mei->p = V3(0, 0, 0); mei->mint = mint;
    float m = M->majorant, inv_m = 1.f / m;
    float wsum = 0.f, w = 1.f, t = mint;
    while (active) {
        float u = next_1d(alt, 1);
        t = t + ((-orc_logf(1.f - u)) * inv_m);
        if (!(t <= maxt)) break;
        wsum = wsum + w;
        float xi = next_1d(alt, 1);
        if (xi * wsum <= w) mei->t = t;
        tri_t T = tri_setup(S, M, ray_at(ray->o, ray->d, t));
        float st = lookup_sigma_t(S, M, &T);
        cnt->c[2]++;
        w = w * ((m - st) * inv_m);
        if (!(w > 0.f)) break;
    }
    if (mei->t < ORC_INF) mei->p = ray_at(ray->o, ray->d, mei->t);
    return wsum * inv_m;
}

static void vps_sample_core(const rcv_scene_desc *S, const orc_medium *M, int primal, pcg32 *rng, pcg32 *alt, ray_t ray, v3 dL,
                            v3 state_in, float result_out[3], grad_sink *G, counters_t *cnt, const drt_opts *drt,
                            int from_state, int ps_active, int ps_depth, si_t ps_si, float ps_last_pdf);

/* sample_recursive (:623-668): a detached primal path from mei, on the secondary sampler */
static v3 sample_recursive(const rcv_scene_desc *S, const orc_medium *M, pcg32 *alt, const ray_t *ray, v3 mei_p, int depth,
                           int active, counters_t *cnt) {
    v3 result = V3(0, 0, 0);
    v3 wi = V3(-ray->d.x, -ray->d.y, -ray->d.z);
    if (S->use_nee)                                                            /* :635-638 */
        result = vadd(result, sample_emitter_for_nee(S, M, mei_p, wi, alt, V3(1, 1, 1), active, 1, V3(0, 0, 0), NULL, cnt));
    float s1 = next_1d(alt, active), s2x = next_1d(alt, active), s2y = next_1d(alt, active);      /* :644-646 */
    ray_t rec_ray = *ray;
    float phase_pdf = 1.f;
    if (active) {
        v3 wo = phase_sample(S, wi, s1, s2x, s2y, &phase_pdf);
        rec_ray = spawn_ray(mei_p, V3(0, 0, 0), wo);                           /* :647-648 */
    }
    si_t si_next; si_next.valid = 0; si_next.t = ORC_INF; si_next.p = V3(0, 0, 0); si_next.n = V3(0, 0, 0);
    if (active) {
        si_next = scene_intersect(S, &rec_ray);                                /* :651 */
        rec_ray.maxt = isfinite(si_next.t) ? si_next.t : ORC_LARGEST;          /* :653-654 */
    }
    int next_depth = active ? depth + 1 : depth;
    float Li[3];
    vps_sample_core(S, M, 1, alt, NULL, rec_ray, V3(0, 0, 0), V3(0, 0, 0), Li, NULL, cnt, NULL, 1,
                    active && (next_depth < S->max_depth), next_depth, si_next, active ? phase_pdf : 1.f);   /* :656-665 */
    result = vadd(result, V3(Li[0], Li[1], Li[2]));
    return active ? result : V3(0, 0, 0);                                      /* :668 */
}

/* backpropagate_scattering_drt without a reservoir (:551-594) */
static void backpropagate_scattering_drt(const rcv_scene_desc *S, const orc_medium *M, pcg32 *alt, ray_t ray, si_t si, int depth,
                                         v3 adjoint, int active, const drt_opts *drt, grad_sink *G, counters_t *cnt) {
    ray_t sub_ray = ray;
    if (active) sub_ray.maxt = isfinite(si.t) ? si.t : ORC_LARGEST;            /* :553-555 */
    mei_t mei_sub;
    float drt_weight = sample_interaction_drt(S, M, &sub_ray, alt, active, &mei_sub, cnt);   /* :560-561 */
    active = active && mei_valid(&mei_sub);                                    /* :568 */
    float st = 0.f; v3 albedo = V3(0, 0, 0);
    if (active) {
        tri_t T = tri_setup(S, M, mei_sub.p);
        st = lookup_sigma_t(S, M, &T); albedo = lookup_albedo(M, &T);
    }
    v3 drt_Li = sample_recursive(S, M, alt, &sub_ray, mei_sub.p, depth, active, cnt);      /* :575-578 */
    if (!active) return;
    float mis = drt->use_drt_mis ? (1.f / (1.f + st * st)) : 1.f;              /* :581-585 */
    /* backward_from(mis * drt_weight * adjoint * (sigma_t * albedo) * drt_Li) (:587-591) */
    v3 c = vmul(vscale(adjoint, mis * drt_weight), drt_Li);
    float g_st = (c.x * albedo.x + c.y * albedo.y) + c.z * albedo.z;
    v3 g_al = vscale(c, st);
    scatter_sigma_t(S, M, G, mei_sub.p, g_st, cnt);
    scatter_albedo(S, M, G, mei_sub.p, g_al, cnt);
}

/* volpathsimple.py:51-303; reach_medium :305-332.  from_state: a recursive call that restores a PathState (:75-81) */
static void vps_sample_core(const rcv_scene_desc *S, const orc_medium *M, int primal, pcg32 *rng, pcg32 *alt, ray_t ray, v3 dL,
                            v3 state_in, float result_out[3], grad_sink *G, counters_t *cnt, const drt_opts *drt,
                            int from_state, int ps_active, int ps_depth, si_t ps_si, float ps_last_pdf) {
    v3 result = primal ? V3(0, 0, 0) : state_in;                               /* :68 */
    v3 throughput = V3(1, 1, 1);
    int active, depth, escaped, has_scattered;
    float last_scatter_direction_pdf;
    si_t si;
    if (from_state) {                                                          /* :75-81, :98-100 */
        active = ps_active; depth = ps_depth; si = ps_si; escaped = 0;
        has_scattered = active && !escaped;
        last_scatter_direction_pdf = ps_last_pdf;
    } else {
        active = 1;                                                            /* :82 */
        depth = 0;
        (void) next_1d(rng, active);                                           /* :84 */
        /* reach_medium (:305-332) */
        si = scene_intersect(S, &ray);                                         /* :311 */
        escaped = active && !si.valid;
        active = active && si.valid;
        if (active) ray = spawn_ray(si.p, si.n, ray.d);                        /* :319 */
        si_t si_new; si_new.valid = 0; si_new.t = ORC_INF; si_new.p = V3(0, 0, 0); si_new.n = V3(0, 0, 0);
        if (active) si_new = scene_intersect(S, &ray);                         /* :320 */
        active = active && si_new.valid;                                       /* :323 */
        if (active) {                                                          /* :329-330 */
            ray.maxt = isfinite(si_new.t) ? si_new.t : ORC_LARGEST;
            si = si_new;
        }
        has_scattered = 0;                                                     /* :101 */
        last_scatter_direction_pdf = 1.f;
    }
    const int use_drt = (!primal) && drt && drt->use_drt;
    drt_reservoir R; memset(&R, 0, sizeof(R)); R.depth = -1; R.active = active;   /* :107-109 */
    (void) next_1d(rng, active);                                               /* :112 alt_seed_rnd */

    while (active) {                                                           /* :127 */
        float q = fminf_(max3(throughput), 0.99f);                             /* :130 */
        int perform_rr = depth > S->rr_depth;
        float rr = next_1d(rng, active);                                       /* :133 */
        active = active && any_nz(throughput) && (!perform_rr || (rr < q));
        if (perform_rr) throughput = vscale(throughput, 1.f / q);              /* :134 */

        mei_t mei;
        sample_real_interaction(S, M, &ray, rng, active, &mei, cnt);           /* :139 */
        int did_escape = active && !mei_valid(&mei);                           /* :143 */
        int still_in_medium = active && mei_valid(&mei);
        int did_scatter = still_in_medium;                                     /* :147 */
        has_scattered = has_scattered || did_scatter;                          /* :150 */
        v3 albedo = V3(1, 1, 1);
        if (did_scatter) {                                                     /* :154 get_albedo */
            tri_t T = tri_setup(S, M, mei.p);
            albedo = lookup_albedo(M, &T);
        }
        if (!primal) {
            if (use_drt) {                                                     /* :156-163 */
                if (drt->use_drt_subsampling) {
                    float xi = next_1d(alt, active);                           /* :546 */
                    reservoir_update(&R, depth, si, ray, throughput, xi, active);
                } else {
                    backpropagate_scattering_drt(S, M, alt, ray, si, depth, vmul(dL, throughput), active, drt, G, cnt);
                }
            }
            if (did_scatter && (!use_drt || drt->use_drt_mis)) {
                /* scattering term (:165-185) */
                float mis = 1.f;
                if (use_drt && drt->use_drt_mis) { float s2 = mei.sigma_t.x * mei.sigma_t.x; mis = s2 / (1.f + s2); }   /* :169-170 */
                v3 inv_pdf = V3(1.f / mei.sigma_t.x, 1.f / mei.sigma_t.y, 1.f / mei.sigma_t.z);   /* :175 */
                v3 Li = vdiv(result, vmax(V3(1e-8f, 1e-8f, 1e-8f), albedo));                       /* :180 */
                /* backward_from(drt_mis_weight * dL * (sigma_t * albedo) * Li * inv_pdf) (:183) */
                v3 c = vmul(vmul(dL, Li), inv_pdf);
                if (use_drt && drt->use_drt_mis) c = vscale(c, mis);
                float g_st = (c.x * albedo.x + c.y * albedo.y) + c.z * albedo.z;
                v3 g_al = vmul(c, mei.sigma_t);
                scatter_sigma_t(S, M, G, mei.p, g_st, cnt);
                scatter_albedo(S, M, G, mei.p, g_al, cnt);
            }
            /* transmittance term (:194-202) */
            v3 adj_weight = vmul(dL, result);                                  /* :198 */
            backpropagate_transmittance(S, M, alt, si.t, mei.t, &ray, adj_weight, did_scatter, did_escape, G, cnt);
        }
        throughput = vmul(throughput, albedo);                                 /* :206 */
        if (did_scatter) depth = depth + 1;                                    /* :212 */
        active = active && still_in_medium && (depth < S->max_depth);          /* :213 */

        v3 wi = V3(-ray.d.x, -ray.d.y, -ray.d.z);
        if (S->use_nee) {                                                      /* :219-228 */
            int active_e = did_scatter && active;
            v3 nee = sample_emitter_for_nee(S, M, mei.p, wi, rng, throughput, active_e, primal, dL, G, cnt);
            if (active_e) result = primal ? vadd(result, nee) : vsub(result, nee);
        }
        float s1 = next_1d(rng, did_scatter);                                  /* :235 */
        float s2x = next_1d(rng, did_scatter), s2y = next_1d(rng, did_scatter);
        if (did_scatter) {
            float phase_pdf;
            v3 wo = phase_sample(S, wi, s1, s2x, s2y, &phase_pdf);
            ray = spawn_ray(mei.p, V3(0, 0, 0), wo);                           /* :236-237 */
            last_scatter_direction_pdf = phase_pdf;                            /* :243 */
        }
        int needs_update = did_scatter || did_escape;                          /* :246 */
        if (needs_update) {
            si = scene_intersect(S, &ray);                                     /* :247 */
            ray.maxt = isfinite(si.t) ? si.t : ORC_LARGEST;                    /* :248 */
        }
        int accidental_escape = did_scatter && !si.valid;                      /* :253 */
        active = active && !accidental_escape;
        if (did_escape) {                                                      /* :257-258 */
            ray = spawn_ray(si.p, si.n, ray.d);
            escaped = 1;
        }
    }
    if (use_drt && drt->use_drt_subsampling) {                                 /* :262-272 */
        v3 adjoint = vmul(reservoir_weight(&R), dL);
        backpropagate_scattering_drt(S, M, alt, R.ray, R.si, R.depth, adjoint, R.active && (R.depth != -1), drt, G, cnt);
    }
    if (primal) {                                                              /* :276-300 */
        int si_update_needed = escaped && si.valid;
        if (si_update_needed) si = scene_intersect(S, &ray);                   /* :278 */
        int has_emitter = !si.valid;                                           /* environment */
        int active_e = escaped && has_emitter && !((depth <= 0) && S->hide_emitters);
        if (active_e) {
            float hit_mis_weight = 1.f;
            if (S->use_nee) {
                float emitter_pdf_ = has_scattered ? emitter_pdf(S, ray.d) : 0.f;   /* :286-290 */
                hit_mis_weight = mis_weight(last_scatter_direction_pdf, emitter_pdf_);
            }
            v3 contrib = emitter_eval(S, ray.d);
            result = vadd(result, vmul(vscale(throughput, hit_mis_weight), contrib));   /* :298 */
        }
    }
    result_out[0] = result.x; result_out[1] = result.y; result_out[2] = result.z;
}

static void vps_sample(const rcv_scene_desc *S, const orc_medium *M, int primal, pcg32 *rng, pcg32 *alt,
                       ray_t ray, v3 dL, v3 state_in, float result_out[3], grad_sink *G, counters_t *cnt) {
    si_t none; memset(&none, 0, sizeof(none));
    vps_sample_core(S, M, primal, rng, alt, ray, dL, state_in, result_out, G, cnt, NULL, 0, 0, 0, none, 1.f);
}

/* the seed of the adjoint's secondary sampler (volpathsimple.py:112-120): the
 * float drawn by *lane 0* after the film and channel draws, reinterpreted as
 * an integer and scrambled.  A pure function of seed_grad. */
/* test hooks for the phase function (isotropic / Henyey-Greenstein [UPSTREAM hg.cpp]) */
float orc_phase_sample(const rcv_scene_desc *S, const float wi[3], float s2x, float s2y, float wo_out[3]) {
    float pdf;
    v3 wo = phase_sample(S, V3(wi[0], wi[1], wi[2]), 0.f, s2x, s2y, &pdf);
    wo_out[0] = wo.x; wo_out[1] = wo.y; wo_out[2] = wo.z;
    return pdf;
}
float orc_phase_eval(const rcv_scene_desc *S, const float wi[3], const float wo[3]) {
    return phase_eval(S, V3(wi[0], wi[1], wi[2]), V3(wo[0], wo[1], wo[2]));
}

uint32_t orc_alt_seed(uint32_t seed_grad) {
    pcg32 r; sampler_seed(&r, seed_grad, 0);
    (void) next_1d(&r, 1); (void) next_1d(&r, 1);       /* film jitter */
    (void) next_1d(&r, 1);                               /* :84 */
    float f = next_1d(&r, 0);                            /* :112 value of lane 0 (advance or not: same value) */
    union { float f; uint32_t u; } b; b.f = f;
    uint32_t v0, v1; orc_tea32(b.u, 1, &v0, &v1);        /* :117-118 */
    return v0;
}

int orc_vps_lanes(const rcv_scene_desc *S, const orc_medium *M, uint32_t seed, uint32_t spp,
                  uint64_t lane_begin, uint64_t lane_end, float *L, int32_t *pix, int nthreads) {
    env_build(S);
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
    {
        counters_t cnt; memset(&cnt, 0, sizeof(cnt));
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = (int64_t) lane_begin; i < (int64_t) lane_end; ++i) {
            lane_t ln = lane_setup(S, seed, spp, (uint64_t) i);
            size_t k = (size_t) (i - (int64_t) lane_begin);
            vps_sample(S, M, 1, &ln.rng, NULL, ln.ray, V3(0, 0, 0), V3(0, 0, 0), L + 3 * k, NULL, &cnt);
            if (pix) pix[k] = ln.pix;
            cnt.c[6]++;
        }
        counters_merge(&cnt);
    }
    return 0;
}

int orc_vps_render(const rcv_scene_desc *S, const orc_medium *M, uint32_t seed, uint32_t spp,
                   uint32_t spp_begin, uint32_t spp_end, float *accum, int nthreads) {
    env_build(S);
    uint64_t npix = (uint64_t) S->film_width * S->film_height;
    uint32_t ns = spp_end - spp_begin;
    float *L = (float *) malloc((size_t) ns * 3 * sizeof(float) * 4096);
    float *px = (float *) malloc((size_t) ns * 2 * sizeof(float) * 4096);
    if (!L || !px) return 1;
    if (nthreads < 1) nthreads = 1;
    for (uint64_t p0 = 0; p0 < npix; p0 += 4096) {
        uint64_t p1 = p0 + 4096 < npix ? p0 + 4096 : npix;
#pragma omp parallel num_threads(nthreads)
        {
            counters_t cnt; memset(&cnt, 0, sizeof(cnt));
#pragma omp for schedule(dynamic, 16)
            for (int64_t p = (int64_t) p0; p < (int64_t) p1; ++p)
                for (uint32_t k = spp_begin; k < spp_end; ++k) {
                    size_t o = (size_t) (p - (int64_t) p0) * ns + (k - spp_begin);
                    lane_t ln = lane_setup(S, seed, spp, (uint64_t) p * spp + k);
                    vps_sample(S, M, 1, &ln.rng, NULL, ln.ray, V3(0, 0, 0), V3(0, 0, 0), L + 3 * o, NULL, &cnt);
                    px[2 * o] = ln.fx; px[2 * o + 1] = ln.fy;
                    cnt.c[6]++;
                }
            counters_merge(&cnt);
        }
        size_t n = (size_t) (p1 - p0) * ns;
        for (size_t o = 0; o < n; ++o) {
            footprint_t fp; film_footprint(S, px[2 * o], px[2 * o + 1], &fp);
            for (int q = 0; q < fp.n; ++q) {
                float *a = accum + 4 * (size_t) fp.pix[q];
                float w = fp.w[q];
                a[0] += L[3 * o] * w; a[1] += L[3 * o + 1] * w; a[2] += L[3 * o + 2] * w; a[3] += w;
            }
        }
    }
    free(L); free(px);
    return 0;
}

/* RBIntegrator.render_backward [UPSTREAM common.py] driving volpathsimple:
 * pass 1 traces L per lane; dL = d(develop o put)/dL * grad_in = sum_p w_p grad_in[p] / W[p] over the
 * sample's filter footprint (one pixel, weight 1, for the box filter); pass 2 replays the identical stream and scatters gradients. */
static int vps_backward_impl(const rcv_scene_desc *S, const orc_medium *M, uint32_t seed_grad, uint32_t spp_grad,
                             uint32_t spp_begin, uint32_t spp_end, const float *grad_img, const float *film_w,
                             double *grad_sigma_t, double *grad_albedo, int nthreads, const drt_opts *drt) {
    env_build(S);
    uint64_t npix = (uint64_t) S->film_width * S->film_height;
    uint32_t alt_seed = orc_alt_seed(seed_grad);
    if (nthreads < 1) nthreads = 1;
    grad_sink G; G.g_sigma_t = grad_sigma_t; G.g_albedo = grad_albedo; G.atomic = nthreads > 1;
#pragma omp parallel num_threads(nthreads)
    {
        counters_t cnt; memset(&cnt, 0, sizeof(cnt));
#pragma omp for schedule(dynamic, 16)
        for (int64_t p = 0; p < (int64_t) npix; ++p)
            for (uint32_t k = spp_begin; k < spp_end; ++k) {
                uint64_t idx = (uint64_t) p * spp_grad + k;
                lane_t ln = lane_setup(S, seed_grad, spp_grad, idx);
                pcg32 rng1 = ln.rng;                       /* sampler.clone() for pass 1 */
                float L[3];
                vps_sample(S, M, 1, &rng1, NULL, ln.ray, V3(0, 0, 0), V3(0, 0, 0), L, NULL, &cnt);
                /* dL = sum over the footprint of w * grad_in[p] / W[p] */
                v3 dL = V3(0, 0, 0);
                footprint_t fp; film_footprint(S, ln.fx, ln.fy, &fp);
                for (int q = 0; q < fp.n; ++q) {
                    int pq = fp.pix[q];
                    float w = film_w[pq];
                    float d = (w == 0.f) ? 1.f : w;
                    dL.x += fp.w[q] * (grad_img[3 * pq] / d);
                    dL.y += fp.w[q] * (grad_img[3 * pq + 1] / d);
                    dL.z += fp.w[q] * (grad_img[3 * pq + 2] / d);
                }
                pcg32 alt; sampler_seed(&alt, alt_seed, (uint32_t) idx);       /* volpathsimple.py:119-120 */
                float L2[3];
                si_t none; memset(&none, 0, sizeof(none));
                vps_sample_core(S, M, 0, &ln.rng, &alt, ln.ray, dL, V3(L[0], L[1], L[2]), L2, &G, &cnt, drt, 0, 0, 0, none, 1.f);
                cnt.c[6] += 2;
            }
        counters_merge(&cnt);
    }
    return 0;
}
int orc_vps_backward(const rcv_scene_desc *S, const orc_medium *M, uint32_t seed_grad, uint32_t spp_grad,
                     uint32_t spp_begin, uint32_t spp_end, const float *grad_img, const float *film_w,
                     double *grad_sigma_t, double *grad_albedo, int nthreads) {
    return vps_backward_impl(S, M, seed_grad, spp_grad, spp_begin, spp_end, grad_img, film_w, grad_sigma_t, grad_albedo,
                             nthreads, NULL);
}
/* the same with use_drt = True (volpathsimple.py:33-36): ORACLE ONLY, see the DRT block above */
int orc_vps_backward_drt(const rcv_scene_desc *S, const orc_medium *M, uint32_t seed_grad, uint32_t spp_grad,
                         uint32_t spp_begin, uint32_t spp_end, const float *grad_img, const float *film_w,
                         double *grad_sigma_t, double *grad_albedo, int nthreads, int use_drt_subsampling, int use_drt_mis) {
    drt_opts o; o.use_drt = 1; o.use_drt_subsampling = use_drt_subsampling; o.use_drt_mis = use_drt_mis;
    return vps_backward_impl(S, M, seed_grad, spp_grad, spp_begin, spp_end, grad_img, film_w, grad_sigma_t, grad_albedo,
                             nthreads, &o);
}

/* RBIntegrator.render_backward [UPSTREAM] driving the adjoint the two-state integrator inherits — upstream
 * PRBVolpathIntegrator.sample (plugins/twostatevolpath.py:11; used by cv_integrator.py:87-94 / metaintegrator.py:140-147
 * when the scene gives no `adjoint_integrator`): pass 1 = the primal estimator of state `new` (sample_twostates with
 * old == new), pass 2 = the same stream again with gradients (prb_adj).  Handles surfaces and the area light. */
int orc_prb_backward(const rcv_scene_desc *S, const orc_medium *M, uint32_t seed_grad, uint32_t spp_grad,
                     uint32_t spp_begin, uint32_t spp_end, const float *grad_img, const float *film_w,
                     double *grad_sigma_t, double *grad_albedo, int nthreads) {
    env_build(S);
    uint64_t npix = (uint64_t) S->film_width * S->film_height;
    if (nthreads < 1) nthreads = 1;
    grad_sink G; G.g_sigma_t = grad_sigma_t; G.g_albedo = grad_albedo; G.atomic = nthreads > 1;
#pragma omp parallel num_threads(nthreads)
    {
        counters_t cnt; memset(&cnt, 0, sizeof(cnt));
#pragma omp for schedule(dynamic, 16)
        for (int64_t p = 0; p < (int64_t) npix; ++p)
            for (uint32_t k = spp_begin; k < spp_end; ++k) {
                uint64_t idx = (uint64_t) p * spp_grad + k;
                lane_t ln = lane_setup(S, seed_grad, spp_grad, idx);
                pcg32 rng1 = ln.rng;                       /* sampler.clone() for pass 1 */
                float Lo[3], L[3]; int valid;
                sample_twostates(S, M, M, &rng1, ln.ray, Lo, L, &valid, &cnt, NULL);
                v3 dL = V3(0, 0, 0);
                footprint_t fp; film_footprint(S, ln.fx, ln.fy, &fp);
                for (int q = 0; q < fp.n; ++q) {
                    int pq = fp.pix[q];
                    float w = film_w[pq];
                    float d = (w == 0.f) ? 1.f : w;
                    dL.x += fp.w[q] * (grad_img[3 * pq] / d);
                    dL.y += fp.w[q] * (grad_img[3 * pq + 1] / d);
                    dL.z += fp.w[q] * (grad_img[3 * pq + 2] / d);
                }
                prb_adj adj; adj.dL = dL; adj.L_total = V3(L[0], L[1], L[2]); adj.G = &G;
                float L2o[3], L2[3];
                sample_twostates(S, M, M, &ln.rng, ln.ray, L2o, L2, &valid, &cnt, &adj);
                cnt.c[6] += 2;
            }
        counters_merge(&cnt);
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* optimiser step, clamp, upsample, rotation                                  */
/* ------------------------------------------------------------------------- */
/* mi.ad.Adam.step [UPSTREAM]: lr_t = lr * sqrt(1-b2^t)/(1-b1^t) (host, double -> float) */
float orc_adam_lr_t(float lr, float beta1, float beta2, int t) {
    double s = sqrt(1.0 - pow((double) beta2, (double) t)) / (1.0 - pow((double) beta1, (double) t));
    return (float) ((double) lr * s);
}
void orc_adam(float *p, const float *g, float *m, float *v, uint64_t n, int t, float lr, float beta1,
              float beta2, float eps) {
    float lr_t = orc_adam_lr_t(lr, beta1, beta2, t);
    for (uint64_t i = 0; i < n; ++i) {
        float m_t = beta1 * m[i] + (1.f - beta1) * g[i];
        float v_t = beta2 * v[i] + ((1.f - beta2) * g[i]) * g[i];
        m[i] = m_t; v[i] = v_t;
        p[i] = fmaf(-lr_t / (sqrtf(v_t) + eps), m_t, p[i]);
    }
}
void orc_clamp(float *p, uint64_t n, float lo, float hi) {               /* utils.py:172-175 */
    for (uint64_t i = 0; i < n; ++i) p[i] = fminf_(fmaxf_(p[i], lo), hi);
}
void orc_upsample2x(const float *in, float *out, int nz, int ny, int nx, int c) {   /* utils.py:184 */
    for (int z = 0; z < 2 * nz; ++z)
        for (int y = 0; y < 2 * ny; ++y)
            for (int x = 0; x < 2 * nx; ++x)
                for (int k = 0; k < c; ++k)
                    out[(((size_t) z * 2 * ny + y) * 2 * nx + x) * c + k] =
                        in[(((size_t) (z / 2) * ny + (y / 2)) * nx + (x / 2)) * c + k];
}
