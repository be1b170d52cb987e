// rcv_device.cuh — device-side building blocks shared by the tracing kernels:
// RNG (TEA-4 + PCG32), the fixed-order float32 elementary functions, the box
// geometry, the packed two-state grid and its trilinear lookups.
//
// Numerical contract: every value that can influence a path
// decision is IEEE binary32 with a fixed evaluation order.  This translation
// unit is compiled with -fmad=false, so a*b+c is never contracted; where the
// specification wants a fused multiply-add it is written as __fmaf_rn.
// Divisions and square roots are the IEEE-rounded ones (-prec-div/-prec-sqrt
// defaults).  log and sincos are the Cephes single-precision polynomials
// spelled out here (not the libdevice versions), so that a lane's random walk
// is reproducible bit for bit against the CPU restatement used in the tests.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

// ----------------------------------------------------------------------------
// RCV_HOST_EMU: the per-lane device code of this directory also compiles with
// a plain host C++ compiler so that tests/ can single-step the lane state
// machines on the CPU and compare them with the oracle without a GPU
// (tests/host_emu).  It is a test harness only: the library (librcv_b200.so)
// is never built this way and has no CPU path.
// ----------------------------------------------------------------------------
#if defined(RCV_HOST_EMU)
#include <algorithm>
#include <cmath>
#include <cstring>
#ifndef __CUDACC__
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
template <typename T> static inline T __ldg(const T *p) { return *p; }
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
using std::isfinite; using std::isnan; using std::isinf; using std::max; using std::min;
#endif
#endif

namespace rcv {

// ----------------------------------------------------------------------------
// kernel-side scene / medium parameter blocks (passed by value)
// ----------------------------------------------------------------------------
// A rectangle with its derived quantities (rcv_scene.h computes them in binary32, fixed operation order)
struct QuadK {
    float p0[3], eu[3], ev[3], n[3], s[3], t[3], refl[3], rad[3];
    float inv_eu2, inv_ev2, inv_area;
};

struct SceneK {
    uint32_t W, H;
    float rcpW, rcpH;
    float cam_o[3];
    float cam_axes[9];
    float tan_x, tan_y, near_clip;
    float bmin[3], bmax[3], inv_ext[3];
    float scale;              // sigma_t scale
    int phase;                // RCV_PHASE_*
    float g;
    int sample_emitters;
    float env[3];
    float env_over_pdf[3];    // env[c] / (1 / 4 pi): constant.cpp sample_direction value (radiance / pdf)
    float env_dist;
    int max_depth, rr_depth, hide_emitters, use_nee;
    int n_sensors;            // `batch` sensor: views side by side on the film (<= 1: the single camera above)
    float batch_o[16 * 3], batch_axes[16 * 9];
    int rfilter;              // RCV_RFILTER_*
    float rf_stddev;
    // emitter: 0 `constant` environment, 1 `envmap`, 2 `area` light on a rectangle (no environment)
    int emitter;
    int env_w, env_h;         // envmap texels: nodes of a (env_h - 1) x env_w cell grid, longitude wraps
    const float4 *env_data;   // (env_h, env_w + 1) scaled texels (rgb, 0), column env_w = column 0
    const float *env_cdf_m;   // env_h: cdf over the cell rows
    const float *env_cdf_c;   // (env_h - 1, env_w + 1): cdf over the cells of a row
    const float *env_pdf;     // (env_h - 1, env_w): density of (u, v) on the unit square
    float env_R[9];           // emitter frame -> world (row-major)
    // rectangles outside the medium's box (diffuse BSDF; quads[area_quad] carries the `area` light when emitter == 2)
    int n_quads, area_quad;
    QuadK quads[8];
};

// Packed medium state in HBM:
//   sig2[v]      = (sigma_t_old, sigma_t_new) of voxel v         (8 B / voxel; canonical state)
//   alb2[2v + s] = (albedo rgb of state s, 0)                    (32 B / voxel)
// v = (z * ny + y) * nx + x, x fastest (the reference's (Z,Y,X,C) layout), and the
// gather bricks built from sig2 after every rotation: one brick per interpolation
// cell c = (cz * (ny+1) + cy) * (nx+1) + cx, cx = floor(pos.x) + 1 in [0, nx], holding
// the cell's 8 corner values (clamp addressing resolved at build time) contiguously,
//   blk2[4c .. 4c+3] = 8 x (sigma_t_old, sigma_t_new)            (64 B / cell: two sectors)
//   blk1[2c .. 2c+1] = 8 x  sigma_t_new                          (32 B / cell: one sector)
// corner order k = 4*zbit + 2*ybit + xbit.  A tentative collision therefore reads ONE
// contiguous brick instead of eight scattered voxels.
struct GridK {
    static constexpr bool kSupergrid = false;
    const float2 *sig2;
    const float4 *blk2;
    const float4 *blk1;
    const float4 *alb2;
    int nx, ny, nz;
    const float *majorants;   // device: [0] = old, [1] = new  (scale * max)
};
// ... plus the majorant supergrid (heterogeneous `majorant_resolution_factor` > 0): per supercell of f^3 voxels the
// (old, new) local majorants, scale baked in.  Kernels are instantiated for GridK (one majorant for the volume) and
// for GridKS (DDA free-flight sampling through the supergrid, twostatemedium.py:25-70).
struct GridKS : GridK {
    static constexpr bool kSupergrid = true;
    const float2 *mgrid;      // [(z * my + y) * mx + x]
    int mx, my, mz;
    float vox[3];             // 1 / (mx, my, mz)
};

// gradient grids: d sigma_t as one float per voxel — dense, so the
// sigma_t-only scatters (transmittance term, NEE replay: ~94 % of all scatter
// sites) hit an array that fits in L2 at 256^3 — and d albedo as float4 per
// voxel (rgb + one unused lane) so a scattering vertex needs one
// red.global.add.v4.f32 per corner.
// RCV_GRAD_BRICKS = 1 (default): the d sigma_t reductions of one interpolation site go to ONE 32-byte "gradient brick"
// of the site's cell (two red.v4, one sector) instead of eight scalar reds in four sectors of the voxel grid;
// launch_unbrick folds the bricks into gsig after the adjoint launch and leaves them zero.  0 = reds on the voxel grid.
#ifndef RCV_GRAD_BRICKS
#define RCV_GRAD_BRICKS 1
#endif
struct GradK {
    float *gsig;
    float4 *galb;
    float4 *gbrick;     // (nz+1)(ny+1)(nx+1) cells x 2 float4, zero between launches (RCV_GRAD_BRICKS), else null
};

struct Counters {
    unsigned long long c[8];
};

// path record queue (structure of arrays, one float4 triple per path)
//   r0 = (rng.state lo, rng.state hi, rng stream v1, film pixel)   [bit patterns]
//   r1 = (o.x, o.y, o.z, aux)        aux: boundary distance (volpathsimple)
//   r2 = (d.x, d.y, d.z, lane id)    lane id: global PCG32 lane index (bit pattern)
struct PathQueue {
    float4 *r0, *r1, *r2;
    float4 *r3;                  // adjoint only: dL of the lane (d loss / d radiance of its sample), written by ray generation
    unsigned long long *count;   // number of records appended by the ray-generation kernel
};

// ----------------------------------------------------------------------------
// float helpers with oracle-identical NaN behaviour
// ----------------------------------------------------------------------------
__device__ __forceinline__ float fmin_(float a, float b) { return (b < a) ? b : a; }
__device__ __forceinline__ float fmax_(float a, float b) { return (b > a) ? b : a; }
__device__ __forceinline__ float inf_() { return __int_as_float(0x7f800000); }
#define RCV_LARGEST 3.402823466e+38f
#define RCV_RAY_EPS (1500.f * 5.9604644775390625e-8f)
#define RCV_SHADOW_EPS (10.f * RCV_RAY_EPS)
#define RCV_INV_FOURPI 0.07957747154594766788f
#define RCV_TWO_PI 6.28318530717958647692f

struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 mk3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ V3 operator/(V3 a, V3 b) { return mk3(a.x / b.x, a.y / b.y, a.z / b.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 splat3(float s) { return mk3(s, s, s); }
__device__ __forceinline__ float max3(V3 a) { return fmax_(fmax_(a.x, a.y), a.z); }
__device__ __forceinline__ V3 vmax3(V3 a, V3 b) { return mk3(fmax_(a.x, b.x), fmax_(a.y, b.y), fmax_(a.z, b.z)); }
__device__ __forceinline__ bool any_nz(V3 a) { return a.x != 0.f || a.y != 0.f || a.z != 0.f; }
__device__ __forceinline__ V3 ray_at(V3 o, V3 d, float t) {
    return mk3(__fmaf_rn(d.x, t, o.x), __fmaf_rn(d.y, t, o.y), __fmaf_rn(d.z, t, o.z));
}
__device__ __forceinline__ float dot3(V3 a, V3 b) { return __fmaf_rn(a.x, b.x, __fmaf_rn(a.y, b.y, a.z * b.z)); }

// ----------------------------------------------------------------------------
// RNG: mi.sample_tea_32 (4 rounds) + drjit PCG32 (independent sampler)
// ----------------------------------------------------------------------------
__host__ __device__ __forceinline__ void tea4(uint32_t v0, uint32_t v1, uint32_t &o0, uint32_t &o1) {
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        sum += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + sum) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + sum) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    o0 = v0; o1 = v1;
}

struct Pcg {
    uint64_t state, inc;
};
__host__ __device__ __forceinline__ uint32_t pcg_u32(Pcg &r) {
    uint64_t old = r.state;
    r.state = old * 0x5851f42d4c957f2dULL + r.inc;
    uint32_t xs = (uint32_t) (((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t) (old >> 59u);
    return (xs >> rot) | (xs << ((0u - rot) & 31u));
}
__host__ __device__ __forceinline__ void pcg_seed(Pcg &r, uint64_t initstate, uint64_t initseq) {
    r.state = 0;
    r.inc = (initseq << 1u) | 1u;
    pcg_u32(r);
    r.state += initstate;
    pcg_u32(r);
}
__device__ __forceinline__ float pcg_f32(Pcg &r) {
    return __uint_as_float((pcg_u32(r) >> 9) | 0x3f800000u) - 1.f;
}
__host__ __device__ __forceinline__ void sampler_seed(Pcg &r, uint32_t seed, uint32_t idx) {
    uint32_t v0, v1;
    tea4(seed, idx, v0, v1);
    pcg_seed(r, (uint64_t) v0, (uint64_t) v1);
}

// ----------------------------------------------------------------------------
// elementary functions, fixed operation order (Cephes logf / sinf / cosf)
// ----------------------------------------------------------------------------
__device__ __forceinline__ float spec_log(float x) {
    uint32_t u = __float_as_uint(x);
    int e = (int) (u >> 23) - 126;
    float m = __uint_as_float((u & 0x007fffffu) | 0x3f000000u);
    if (m < 0.707106781186547524f) { e -= 1; m = (m + m) - 1.f; }
    else { m = m - 1.f; }
    float z = m * m;
    float y = 7.0376836292e-2f;
    y = __fmaf_rn(y, m, -1.1514610310e-1f);
    y = __fmaf_rn(y, m, 1.1676998740e-1f);
    y = __fmaf_rn(y, m, -1.2420140846e-1f);
    y = __fmaf_rn(y, m, 1.4249322787e-1f);
    y = __fmaf_rn(y, m, -1.6668057665e-1f);
    y = __fmaf_rn(y, m, 2.0000714765e-1f);
    y = __fmaf_rn(y, m, -2.4999993993e-1f);
    y = __fmaf_rn(y, m, 3.3333331174e-1f);
    y = (y * m) * z;
    float fe = (float) e;
    y = __fmaf_rn(fe, -2.12194440e-4f, y);
    y = __fmaf_rn(-0.5f, z, y);
    float r = m + y;
    r = __fmaf_rn(fe, 0.693359375f, r);
    return r;
}

// Cephes expf for x <= 0 of moderate size (reconstruction-filter weights)
__device__ __forceinline__ float spec_exp(float x) {
    float z = floorf(__fmaf_rn(1.44269504088896341f, x, 0.5f));
    x = __fmaf_rn(z, -0.693359375f, x);
    x = __fmaf_rn(z, 2.12194440e-4f, x);
    int n = (int) z;
    float xx = x * x;
    float y = 1.9875691500e-4f;
    y = __fmaf_rn(y, x, 1.3981999507e-3f);
    y = __fmaf_rn(y, x, 8.3334519073e-3f);
    y = __fmaf_rn(y, x, 4.1665795894e-2f);
    y = __fmaf_rn(y, x, 1.6666665459e-1f);
    y = __fmaf_rn(y, x, 5.0000001201e-1f);
    y = __fmaf_rn(y, xx, x);
    y = y + 1.f;
    if (n < -126) return 0.f;
    return y * __uint_as_float((uint32_t) (n + 127) << 23);
}

__device__ __forceinline__ void spec_sincos(float x, float &s_out, float &c_out) {
    int j = (int) (x * 1.27323954473516f);
    j = j + (j & 1);
    float y = (float) j;
    float r = __fmaf_rn(y, -0.78515625f, x);
    r = __fmaf_rn(y, -2.4187564849853515625e-4f, r);
    r = __fmaf_rn(y, -3.77489497744594108e-8f, r);
    float z = r * r;
    float ps = -1.9515295891e-4f;
    ps = __fmaf_rn(ps, z, 8.3321608736e-3f);
    ps = __fmaf_rn(ps, z, -1.6666654611e-1f);
    ps = __fmaf_rn(ps * z, r, r);
    float pc = 2.443315711809948e-5f;
    pc = __fmaf_rn(pc, z, -1.388731625493765e-3f);
    pc = __fmaf_rn(pc, z, 4.166664568298827e-2f);
    pc = __fmaf_rn(pc * z, z, __fmaf_rn(-0.5f, z, 1.f));
    int q = (j >> 1) & 3;
    float s = (q & 1) ? pc : ps;
    float c = (q & 1) ? ps : pc;
    if (q & 2) s = -s;
    if ((q + 1) & 2) c = -c;
    s_out = s; c_out = c;
}

// Cephes atanf / atan2f, fixed operation order (envmap: direction -> (u, v))
#define RCV_PI 3.14159265358979323846f
__device__ __forceinline__ float spec_atan(float xx) {
    float x = fabsf(xx), y;
    if (x > 2.414213562373095f) { y = 1.57079632679489661923f; x = -(1.f / x); }
    else if (x > 0.4142135623730950f) { y = 0.78539816339744830962f; x = (x - 1.f) / (x + 1.f); }
    else y = 0.f;
    float z = x * x;
    float p = 8.05374449538e-2f;
    p = __fmaf_rn(p, z, -1.38776856032e-1f);
    p = __fmaf_rn(p, z, 1.99777106478e-1f);
    p = __fmaf_rn(p, z, -3.33329491539e-1f);
    y = y + __fmaf_rn(p * z, x, x);
    return (xx < 0.f) ? -y : y;
}
__device__ __forceinline__ float spec_atan2(float y, float x) {
    if (x == 0.f) return (y > 0.f) ? 1.57079632679489661923f : ((y < 0.f) ? -1.57079632679489661923f : 0.f);
    float z = spec_atan(y / x);
    if (x < 0.f) z = (y < 0.f) ? z - RCV_PI : z + RCV_PI;
    return z;
}

__device__ __forceinline__ V3 uniform_sphere(float sx, float sy) {
    float z = __fmaf_rn(-2.f, sy, 1.f);
    float r2 = __fmaf_rn(-z, z, 1.f);
    float r = sqrtf(fmax_(r2, 0.f));
    float s, c;
    spec_sincos(RCV_TWO_PI * sx, s, c);
    return mk3(r * c, r * s, z);
}

__device__ __forceinline__ float mis_weight(float pdf_a, float pdf_b) {
    float a2 = pdf_a * pdf_a, b2 = pdf_b * pdf_b;
    float w = a2 / (a2 + b2);
    return isfinite(w) ? w : 0.f;
}

// ----------------------------------------------------------------------------
// environment emitter: `constant`, or `envmap` (upstream envmap.cpp restated): bilinear radiance over the
// lat-long nodes, piecewise-constant sampling density per cell (tables built by rcv_scene_set).  A sample's density and
// value are evaluated FROM the sampled direction, so emitter hits and emitter samples share env_pdf / env_eval.
// Out of line: called once per path / per scattering vertex, in the vertex and slow steps.
// ----------------------------------------------------------------------------
__device__ __forceinline__ void env_locate(const SceneK &S, V3 d, int &r, int &c, float &fu, float &fv, float &st) {
    const float *R = S.env_R;
    float lx = __fmaf_rn(R[0], d.x, __fmaf_rn(R[3], d.y, R[6] * d.z));
    float ly = __fmaf_rn(R[1], d.x, __fmaf_rn(R[4], d.y, R[7] * d.z));
    float lz = __fmaf_rn(R[2], d.x, __fmaf_rn(R[5], d.y, R[8] * d.z));
    st = sqrtf(__fmaf_rn(lx, lx, lz * lz));
    float u = spec_atan2(lx, -lz) * (0.5f / RCV_PI);
    u = u - floorf(u);
    if (!(u < 1.f)) u = 0.f;
    float v = spec_atan2(st, ly) * (1.f / RCV_PI);
    float x = u * (float) S.env_w, y = v * (float) (S.env_h - 1);
    int ci = (int) floorf(x), ri = (int) floorf(y);
    ci = min(ci, S.env_w - 1); ri = min(ri, S.env_h - 2);
    ci = max(ci, 0); ri = max(ri, 0);
    c = ci; r = ri; fu = x - (float) ci; fv = y - (float) ri;
}
static __device__ __noinline__ float4 env_eval_pdf(const SceneK &S, float dx, float dy, float dz) {
    int r, c; float fu, fv, st;
    env_locate(S, mk3(dx, dy, dz), r, c, fu, fv, st);
    const float4 *t = S.env_data + (size_t) r * (S.env_w + 1) + c;
    float4 t00 = __ldg(t), t01 = __ldg(t + 1), t10 = __ldg(t + S.env_w + 1), t11 = __ldg(t + S.env_w + 2);
    float gu = 1.f - fu, gv = 1.f - fv;
    float4 o;
    o.x = __fmaf_rn(gv, __fmaf_rn(gu, t00.x, fu * t01.x), fv * __fmaf_rn(gu, t10.x, fu * t11.x));
    o.y = __fmaf_rn(gv, __fmaf_rn(gu, t00.y, fu * t01.y), fv * __fmaf_rn(gu, t10.y, fu * t11.y));
    o.z = __fmaf_rn(gv, __fmaf_rn(gu, t00.z, fu * t01.z), fv * __fmaf_rn(gu, t10.z, fu * t11.z));
    o.w = __ldg(S.env_pdf + (size_t) r * S.env_w + c) / ((2.f * RCV_PI * RCV_PI) * fmax_(st, 1e-6f));
    return o;
}
__device__ __forceinline__ int cdf_find(const float *cdf, int n, float u) {
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (__ldg(cdf + mid) <= u) lo = mid; else hi = mid; }
    return lo;
}
static __device__ __noinline__ float4 env_sample_dir(const SceneK &S, float u1, float u2) {
    int r = cdf_find(S.env_cdf_m, S.env_h, u2);
    float c0 = __ldg(S.env_cdf_m + r), den = __ldg(S.env_cdf_m + r + 1) - c0;
    float dv = (den > 0.f) ? (u2 - c0) / den : 0.5f;
    const float *row = S.env_cdf_c + (size_t) r * (S.env_w + 1);
    int c = cdf_find(row, S.env_w + 1, u1);
    c0 = __ldg(row + c); den = __ldg(row + c + 1) - c0;
    float du = (den > 0.f) ? (u1 - c0) / den : 0.5f;
    float u = ((float) c + du) / (float) S.env_w, v = ((float) r + dv) / (float) (S.env_h - 1);
    float sp, cp, st, ct;
    spec_sincos(u * RCV_TWO_PI, sp, cp);
    spec_sincos(v * RCV_PI, st, ct);
    float lx = sp * st, ly = ct, lz = -(cp * st);
    const float *R = S.env_R;
    return make_float4(__fmaf_rn(R[0], lx, __fmaf_rn(R[1], ly, R[2] * lz)), __fmaf_rn(R[3], lx, __fmaf_rn(R[4], ly, R[5] * lz)),
                       __fmaf_rn(R[6], lx, __fmaf_rn(R[7], ly, R[8] * lz)), 0.f);
}
// radiance arriving along a ray of direction d that leaves the scene, and the solid-angle density of sampling d
__device__ __forceinline__ V3 emitter_eval(const SceneK &S, V3 d, float &pdf) {
    if (S.emitter == 1) { float4 e = env_eval_pdf(S, d.x, d.y, d.z); pdf = e.w; return mk3(e.x, e.y, e.z); }
    if (S.emitter == 2) { pdf = 0.f; return mk3(0, 0, 0); }
    pdf = RCV_INV_FOURPI;
    return mk3(S.env[0], S.env[1], S.env[2]);
}
// direction of an environment sample
__device__ __forceinline__ V3 emitter_sample_dir(const SceneK &S, float s2x, float s2y) {
    if (S.emitter == 1) { float4 d = env_sample_dir(S, s2x, s2y); return mk3(d.x, d.y, d.z); }
    return uniform_sphere(s2x, s2y);
}
// its density and radiance / density (0 density: the sample is invalid, value 0)
__device__ __forceinline__ V3 emitter_sample_value(const SceneK &S, V3 d, float &pdf) {
    if (S.emitter == 1) {
        float4 e = env_eval_pdf(S, d.x, d.y, d.z);
        pdf = e.w;
        return (pdf != 0.f) ? mk3(e.x / pdf, e.y / pdf, e.z / pdf) : mk3(0, 0, 0);
    }
    pdf = RCV_INV_FOURPI;
    return mk3(S.env_over_pdf[0], S.env_over_pdf[1], S.env_over_pdf[2]);
}

// ----------------------------------------------------------------------------
// phase functions
// ----------------------------------------------------------------------------
__device__ __forceinline__ float phase_eval(const SceneK &S, V3 wi, V3 wo) {
    if (S.phase == 0) return RCV_INV_FOURPI;
    float g = S.g;
    float cos_theta = dot3(wi, wo);
    float temp = 1.f + g * g + 2.f * g * cos_theta;
    return RCV_INV_FOURPI * (1.f - g * g) / (temp * sqrtf(temp));
}
__device__ __forceinline__ V3 phase_sample(const SceneK &S, V3 wi, float s2x, float s2y, float &pdf) {
    if (S.phase == 0) {
        pdf = RCV_INV_FOURPI;
        return uniform_sphere(s2x, s2y);
    }
    float g = S.g;
    float sqr_term = (1.f - g * g) / (1.f - g + 2.f * g * s2x);
    float cos_theta = (1.f + g * g - sqr_term * sqr_term) / (2.f * g);
    if (fabsf(g) < 1.1920929e-7f) cos_theta = 1.f - 2.f * s2x;
    float sin_theta = sqrtf(fmax_(__fmaf_rn(-cos_theta, cos_theta, 1.f), 0.f));
    float sp, cp;
    spec_sincos(RCV_TWO_PI * s2y, sp, cp);
    V3 l = mk3(sin_theta * cp, sin_theta * sp, -cos_theta);
    float sign = copysignf(1.f, wi.z);
    float a = -1.f / (sign + wi.z);
    float b = wi.x * wi.y * a;
    V3 s = mk3(__fmaf_rn(wi.x * wi.x, a * sign, 1.f), b * sign, -wi.x * sign);
    V3 t = mk3(b, __fmaf_rn(wi.y * wi.y, a, sign), -wi.y);
    V3 wo = mk3(__fmaf_rn(s.x, l.x, __fmaf_rn(t.x, l.y, wi.x * l.z)),
                __fmaf_rn(s.y, l.x, __fmaf_rn(t.y, l.y, wi.y * l.z)),
                __fmaf_rn(s.z, l.x, __fmaf_rn(t.z, l.y, wi.z * l.z)));
    float temp = 1.f + g * g - 2.f * g * cos_theta;
    pdf = RCV_INV_FOURPI * (1.f - g * g) / (temp * sqrtf(temp));
    return wo;
}

// ----------------------------------------------------------------------------
// geometry: the medium's box (null-BSDF boundary)
// ----------------------------------------------------------------------------
struct Hit {
    bool valid;
    float t;
    V3 p, n;
};

// reciprocal direction of a ray (IEEE divisions); computed once per ray and
// reused by every slab test along it — bit-identical to recomputing it
__device__ __forceinline__ V3 rcp3(V3 d) { return mk3(1.f / d.x, 1.f / d.y, 1.f / d.z); }

// slab intervals of all three axes; returns tn (entry), tf (exit) and the axes
__device__ __forceinline__ void slabs(const SceneK &S, V3 o, V3 id, float &tn, float &tf, int &an, int &af) {
    float oo[3] = {o.x, o.y, o.z}, ii[3] = {id.x, id.y, id.z};
    tn = -inf_(); tf = inf_(); an = 0; af = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float rcp = ii[a];
        float t1 = (S.bmin[a] - oo[a]) * rcp, t2 = (S.bmax[a] - oo[a]) * rcp;
        float t1p = fmin_(t1, t2), t2p = fmax_(t1, t2);
        if (t1p > tn) { tn = t1p; an = a; }
        if (t2p < tf) { tf = t2p; af = a; }
    }
}

// Medium::prepare_interaction_sampling: [mint, maxt] of the ray inside the box
__device__ __forceinline__ bool medium_interval(const SceneK &S, V3 o, V3 d, V3 id, float ray_maxt, float &mint, float &maxt) {
    float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z}, ii[3] = {id.x, id.y, id.z};
    bool active = true;
    float tn = -inf_(), tf = inf_();
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        active = active && (dd[a] != 0.f || oo[a] > S.bmin[a] || oo[a] < S.bmax[a]);
        float rcp = ii[a];
        float t1 = (S.bmin[a] - oo[a]) * rcp, t2 = (S.bmax[a] - oo[a]) * rcp;
        tn = fmax_(tn, fmin_(t1, t2));
        tf = fmin_(tf, fmax_(t1, t2));
    }
    active = active && (tf >= tn);
    active = active && (isfinite(tn) || isfinite(tf));
    if (!active) { tn = 0.f; tf = inf_(); }
    mint = fmax_(0.f, tn);
    maxt = fmin_(ray_maxt, tf);
    return active;
}

// nearest boundary hit with 0 < t <= maxt
__device__ __forceinline__ Hit box_hit(const SceneK &S, V3 o, V3 d, V3 id, float maxt) {
    Hit h; h.valid = false; h.t = inf_(); h.p = mk3(0, 0, 0); h.n = mk3(0, 0, 0);
    float tn, tf; int an, af;
    slabs(S, o, id, tn, tf, an, af);
    if (!(tf >= tn)) return h;
    float t; int axis;
    bool entering = tn > 0.f;
    if (entering) { t = tn; axis = an; }
    else if (tf > 0.f) { t = tf; axis = af; }
    else return h;
    if (!(t <= maxt) || !isfinite(t)) return h;
    float dd = (axis == 0) ? d.x : ((axis == 1) ? d.y : d.z);
    bool positive_face = entering ? (dd < 0.f) : (dd > 0.f);
    float nsign = positive_face ? 1.f : -1.f;
    V3 p = ray_at(o, d, t);
    float plane;
    if (axis == 0) { plane = positive_face ? S.bmax[0] : S.bmin[0]; p.x = plane; h.n = mk3(nsign, 0, 0); }
    else if (axis == 1) { plane = positive_face ? S.bmax[1] : S.bmin[1]; p.y = plane; h.n = mk3(0, nsign, 0); }
    else { plane = positive_face ? S.bmax[2] : S.bmin[2]; p.z = plane; h.n = mk3(0, 0, nsign); }
    h.valid = true; h.t = t; h.p = p;
    return h;
}

// Interaction::spawn_ray origin: p offset along n towards the side d points to
__device__ __forceinline__ V3 offset_p(V3 p, V3 n, V3 d) {
    float mag = (1.f + fmax_(fmax_(fabsf(p.x), fabsf(p.y)), fabsf(p.z))) * RCV_RAY_EPS;
    float dn = dot3(n, d);
    if (dn < 0.f) mag = -mag;
    return mk3(__fmaf_rn(mag, n.x, p.x), __fmaf_rn(mag, n.y, p.y), __fmaf_rn(mag, n.z, p.z));
}

// ----------------------------------------------------------------------------
// trilinear lookups on the packed grids
// ----------------------------------------------------------------------------
struct Tri {
    bool inside;
    int x0, x1, y0, y1, z0, z1;
    int cx, cy, cz;           // interpolation cell = floor(pos) + 1 per axis, in [0, res]
    float wx0, wx1, wy0, wy1, wz0, wz1;
};

__device__ __forceinline__ void tri_axis(float p, float bmin, float inv_ext, int res, bool &inside, int &i0, int &i1, int &c,
                                         float &w0, float &w1) {
    float pl = (p - bmin) * inv_ext;
    inside = inside && (pl >= 0.f) && (pl <= 1.f);
    float pos = __fmaf_rn(pl, (float) res, -0.5f);
    float fl = floorf(pos);
    int i = (int) fl;
    w1 = pos - fl;
    w0 = 1.f - w1;
    i0 = min(max(i, 0), res - 1);
    i1 = min(max(i + 1, 0), res - 1);
    c = min(max(i + 1, 0), res);
}
__device__ __forceinline__ Tri tri_setup(const SceneK &S, const GridK &G, V3 p) {
    Tri T; T.inside = true;
    tri_axis(p.x, S.bmin[0], S.inv_ext[0], G.nx, T.inside, T.x0, T.x1, T.cx, T.wx0, T.wx1);
    tri_axis(p.y, S.bmin[1], S.inv_ext[1], G.ny, T.inside, T.y0, T.y1, T.cy, T.wy0, T.wy1);
    tri_axis(p.z, S.bmin[2], S.inv_ext[2], G.nz, T.inside, T.z0, T.z1, T.cz, T.wz0, T.wz1);
    return T;
}
__device__ __forceinline__ uint32_t tri_cell(const GridK &G, const Tri &T) {
    return ((uint32_t) T.cz * (uint32_t) (G.ny + 1) + (uint32_t) T.cy) * (uint32_t) (G.nx + 1) + (uint32_t) T.cx;
}
// the 8 corner voxels of cell (cx,cy,cz) in brick order — used by the brick builder
__host__ __device__ __forceinline__ size_t cell_corner(int nx, int ny, int nz, int cx, int cy, int cz, int k) {
    int x = (k & 1) ? cx : cx - 1, y = (k & 2) ? cy : cy - 1, z = (k & 4) ? cz : cz - 1;
    x = x < 0 ? 0 : (x > nx - 1 ? nx - 1 : x);
    y = y < 0 ? 0 : (y > ny - 1 ? ny - 1 : y);
    z = z < 0 ? 0 : (z > nz - 1 ? nz - 1 : z);
    return ((size_t) z * ny + y) * nx + x;
}
// v[k], k = 4*zbit + 2*ybit + xbit
__device__ __forceinline__ float tri_lerp(const Tri &T, float v0_, float v1_, float v2_, float v3_, float v4_, float v5_,
                                          float v6_, float v7_) {
    float a0 = __fmaf_rn(T.wx0, v0_, T.wx1 * v1_);
    float a1 = __fmaf_rn(T.wx0, v2_, T.wx1 * v3_);
    float a2 = __fmaf_rn(T.wx0, v4_, T.wx1 * v5_);
    float a3 = __fmaf_rn(T.wx0, v6_, T.wx1 * v7_);
    float b0 = __fmaf_rn(T.wy0, a0, T.wy1 * a1);
    float b1 = __fmaf_rn(T.wy0, a2, T.wy1 * a3);
    return __fmaf_rn(T.wz0, b0, T.wz1 * b1);
}
__device__ __forceinline__ void tri_rows(const GridK &G, const Tri &T, uint32_t r[4]) {
    const uint32_t ny = (uint32_t) G.ny, nx = (uint32_t) G.nx;
    r[0] = ((uint32_t) T.z0 * ny + (uint32_t) T.y0) * nx;
    r[1] = ((uint32_t) T.z0 * ny + (uint32_t) T.y1) * nx;
    r[2] = ((uint32_t) T.z1 * ny + (uint32_t) T.y0) * nx;
    r[3] = ((uint32_t) T.z1 * ny + (uint32_t) T.y1) * nx;
}
__device__ __forceinline__ float tri_weight(const Tri &T, int k) {
    float wx = (k & 1) ? T.wx1 : T.wx0;
    float wy = (k & 2) ? T.wy1 : T.wy0;
    float wz = (k & 4) ? T.wz1 : T.wz0;
    return (wz * wy) * wx;
}
__device__ __forceinline__ uint32_t tri_voxel(const GridK &G, const Tri &T, int k) {
    int x = (k & 1) ? T.x1 : T.x0;
    int y = (k & 2) ? T.y1 : T.y0;
    int z = (k & 4) ? T.z1 : T.z0;
    return ((uint32_t) z * (uint32_t) G.ny + (uint32_t) y) * (uint32_t) G.nx + (uint32_t) x;
}

// brick loads: read-only, one 256-bit load per 32 bytes (sm_100 LDG.256: half the load instructions and half the L1
// wavefronts of 128-bit loads) with the L2::64B prefetch-size hint — without it every L2 miss of these random
// gathers fetches a 128-byte line from HBM (measured: scripts/microbench/gather_bw.cu, profiles/README.md).
// RCV_BRICK_LOAD: 0 = four/two 128-bit __ldg, 1 = 256-bit, 2 = 256-bit + L2::64B (default)
#ifndef RCV_BRICK_LOAD
#define RCV_BRICK_LOAD 2
#endif
__device__ __forceinline__ void ld_brick32(const float4 *p, float4 &a, float4 &b) {
#if RCV_BRICK_LOAD == 0 || !defined(__CUDA_ARCH__)
    a = __ldg(p); b = __ldg(p + 1);
#elif RCV_BRICK_LOAD == 1
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
#else
    asm("ld.global.nc.L2::64B.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
#endif
}

// both states' sigma_t (scale * grid) with one gather of 8 float2
__device__ __forceinline__ void lookup_sigma2(const SceneK &S, const GridK &G, const Tri &T, float &s_old, float &s_new) {
    if (!T.inside) { s_old = 0.f; s_new = 0.f; return; }
    const float4 *b = G.blk2 + 4 * (size_t) tri_cell(G, T);
    float4 q0, q1, q2, q3;
    ld_brick32(b, q0, q1); ld_brick32(b + 2, q2, q3);                                // one 64-byte brick
    s_old = S.scale * tri_lerp(T, q0.x, q0.z, q1.x, q1.z, q2.x, q2.z, q3.x, q3.z);
    s_new = S.scale * tri_lerp(T, q0.y, q0.w, q1.y, q1.w, q2.y, q2.w, q3.y, q3.w);
}
// state `new` only
__device__ __forceinline__ float lookup_sigma1(const SceneK &S, const GridK &G, const Tri &T) {
    if (!T.inside) return 0.f;
    const float4 *b = G.blk1 + 2 * (size_t) tri_cell(G, T);
    float4 q0, q1;
    ld_brick32(b, q0, q1);                                                           // one 32-byte brick
    return S.scale * tri_lerp(T, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w);
}
// Two-phase lookups (rcv_pool.cuh, RCV_POOL_ILP > 1): *_issue starts the brick load of a tentative collision and *_finish
// interpolates it, so that one thread can have the bricks of several lanes in flight before it consumes the first.
// The loads are volatile: they stay where they are written (a pure asm would be sunk to its first use).  Same values
// and the same arithmetic as lookup_sigma1.
struct Brick1 { float4 q0, q1; };
__device__ __forceinline__ void ld_brick32_issue(const float4 *p, float4 &a, float4 &b) {
#if RCV_BRICK_LOAD != 2 || !defined(__CUDA_ARCH__)
    a = __ldg(p); b = __ldg(p + 1);
#else
    asm volatile("ld.global.nc.L2::64B.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
#endif
}
__device__ __forceinline__ void sigma1_issue(const GridK &G, const Tri &T, Brick1 &b) {
    b.q0 = make_float4(0.f, 0.f, 0.f, 0.f); b.q1 = b.q0;
    if (T.inside) ld_brick32_issue(G.blk1 + 2 * (size_t) tri_cell(G, T), b.q0, b.q1);
}
__device__ __forceinline__ float sigma1_finish(const SceneK &S, const Tri &T, const Brick1 &b) {
    if (!T.inside) return 0.f;
    return S.scale * tri_lerp(T, b.q0.x, b.q0.y, b.q0.z, b.q0.w, b.q1.x, b.q1.y, b.q1.z, b.q1.w);
}
// 16-byte read-only load with the same L2::64B hint (albedo corners: scattered 32-byte voxel records)
__device__ __forceinline__ float4 ld_corner16(const float4 *p) {
#if RCV_BRICK_LOAD != 2 || !defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    float4 a;
    asm("ld.global.nc.L2::64B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(p));
    return a;
#endif
}
// albedo of state s (0 = old, 1 = new)
__device__ __forceinline__ V3 lookup_albedo(const GridK &G, const Tri &T, int s) {
    if (!T.inside) return mk3(0, 0, 0);
    uint32_t r[4];
    tri_rows(G, T, r);
    float4 v[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        v[2 * k] = ld_corner16(G.alb2 + 2 * (size_t) (r[k] + T.x0) + s);
        v[2 * k + 1] = ld_corner16(G.alb2 + 2 * (size_t) (r[k] + T.x1) + s);
    }
    return mk3(tri_lerp(T, v[0].x, v[1].x, v[2].x, v[3].x, v[4].x, v[5].x, v[6].x, v[7].x),
               tri_lerp(T, v[0].y, v[1].y, v[2].y, v[3].y, v[4].y, v[5].y, v[6].y, v[7].y),
               tri_lerp(T, v[0].z, v[1].z, v[2].z, v[3].z, v[4].z, v[5].z, v[6].z, v[7].z));
}
// both states' albedo: the two records of a voxel are adjacent, one 256-bit load per corner
__device__ __forceinline__ void lookup_albedo2(const GridK &G, const Tri &T, V3 &a_old, V3 &a_new) {
    if (!T.inside) { a_old = mk3(0, 0, 0); a_new = mk3(0, 0, 0); return; }
    uint32_t r[4];
    tri_rows(G, T, r);
    float4 o[8], n[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        ld_brick32(G.alb2 + 2 * (size_t) (r[k] + T.x0), o[2 * k], n[2 * k]);
        ld_brick32(G.alb2 + 2 * (size_t) (r[k] + T.x1), o[2 * k + 1], n[2 * k + 1]);
    }
    a_old = mk3(tri_lerp(T, o[0].x, o[1].x, o[2].x, o[3].x, o[4].x, o[5].x, o[6].x, o[7].x),
                tri_lerp(T, o[0].y, o[1].y, o[2].y, o[3].y, o[4].y, o[5].y, o[6].y, o[7].y),
                tri_lerp(T, o[0].z, o[1].z, o[2].z, o[3].z, o[4].z, o[5].z, o[6].z, o[7].z));
    a_new = mk3(tri_lerp(T, n[0].x, n[1].x, n[2].x, n[3].x, n[4].x, n[5].x, n[6].x, n[7].x),
                tri_lerp(T, n[0].y, n[1].y, n[2].y, n[3].y, n[4].y, n[5].y, n[6].y, n[7].y),
                tri_lerp(T, n[0].z, n[1].z, n[2].z, n[3].z, n[4].z, n[5].z, n[6].z, n[7].z));
}

// ----------------------------------------------------------------------------
// majorant supergrid: nearest lookup (clamped) and DDA free-flight sampling.
// The arithmetic and its order are part of the numerical specification: the CPU restatement used by
// the tests evaluates exactly these expressions.
// SEL: 0 = max(old, new), 1 = old, 2 = new.
// ----------------------------------------------------------------------------
template <int SEL>
__device__ __forceinline__ float majorant_at(const SceneK &S, const GridKS &G, V3 p) {
    int ix = (int) floorf(((p.x - S.bmin[0]) * S.inv_ext[0]) * (float) G.mx);
    int iy = (int) floorf(((p.y - S.bmin[1]) * S.inv_ext[1]) * (float) G.my);
    int iz = (int) floorf(((p.z - S.bmin[2]) * S.inv_ext[2]) * (float) G.mz);
    ix = min(max(ix, 0), G.mx - 1); iy = min(max(iy, 0), G.my - 1); iz = min(max(iz, 0), G.mz - 1);
    float2 m = __ldg(G.mgrid + ((size_t) iz * G.my + iy) * G.mx + ix);
    return (SEL == 1) ? m.x : ((SEL == 2) ? m.y : fmax_(m.x, m.y));
}

__device__ __forceinline__ void dda_axis(float o, float d, float id, float bmin, float bmax, float inv_ext, int res, float voxel,
                                         float mint, int &cell, int &step, float &tmax, float &tdelta) {
    float lo = (o - bmin) * inv_ext, ld = d * inv_ext;
    float fres = (float) res;
    int cur = (int) floorf(__fmaf_rn(ld, mint, lo) * fres);
    cur = min(max(cur, 0), res - 1);
    cell = cur;
    step = (ld >= 0.f) ? 1 : -1;
    float nb = (float) (cur + (ld >= 0.f ? 1 : 0)) * voxel;            // the boundary ahead
    if (ld != 0.f) {
        float rld = id * (bmax - bmin);                               // 1 / ld from the ray's cached reciprocal direction
        tmax = (nb - lo) * rld;
        tdelta = ((float) step * voxel) * rld;
    } else {
        tmax = inf_(); tdelta = inf_();
    }
}

// Sampled free-flight distance (or inf) for optical depth `tau` along o + t d, t in [mint, maxt], and the majorant of the
// cell the sample fell into.  Incremental traversal: the supercell index is carried and advanced by +-1 on the axis
// whose boundary is crossed (the numerical specification of DESIGN.md §2; id = 1 / d).
//
// RCV_DDA_AHEAD = K > 1 (build option, measured slower, profiles/README.md r2ag): which supercells the ray visits is
// geometry — only where the walk STOPS depends on the majorants read — so the walk can be taken K cells at a time: the
// crossing distances of the next K cells are computed and their majorants requested together, then the optical depth is
// accumulated over them in order with the tests of the plain loop.  Same operations on the same values in the same
// order; the majorants of cells beyond the stopping point are read and dropped.  The walk turned out to be bound by
// its instructions, not by the majorant loads (rays of a warp walk neighbouring supercells: L1 hits), so the default
// stays K = 1, the plain loop.
#ifndef RCV_DDA_AHEAD
#define RCV_DDA_AHEAD 1
#endif
// RCV_DDA_LIN_TWOSTATE / RCV_DDA_LIN_VPS: which machines walk with dda_sample_lin (r2ah: the two-state render gains
// 22 % at factor 8, the adjoint — at its register limit — does not).
#ifndef RCV_DDA_LIN_TWOSTATE
#define RCV_DDA_LIN_TWOSTATE 1
#endif
#ifndef RCV_DDA_LIN_VPS
#define RCV_DDA_LIN_VPS 0
#endif
// The same walk, leaner.  What the loop carries is chosen for the instruction count of one cell (a warp's collision round lasts
// as long as its longest walk — 31 cells for a mean of 9 at 256^3 / factor 8 — so the cell is the unit that matters): the
// supercell is a running LINEAR index with one signed stride per axis, the grid bounds are the number of steps left on
// each axis, and the step is taken with selects (every lane of the warp steps a different axis; three divergent
// branches cost three times the instructions).  Same values as stepping (cx, cy, cz): the cell sequence, the crossing
// distances and the stopping rule are those of the specification.
template <int SEL>
__device__ __forceinline__ float dda_sample_lin(const SceneK &S, const GridKS &G, V3 o, V3 d, V3 id, float mint, float maxt, float tau,
                                                float &maj_out) {
    float tx, ty, tz, dx, dy, dz;
    int cx, cy, cz, sx, sy, sz;
    dda_axis(o.x, d.x, id.x, S.bmin[0], S.bmax[0], S.inv_ext[0], G.mx, G.vox[0], mint, cx, sx, tx, dx);
    dda_axis(o.y, d.y, id.y, S.bmin[1], S.bmax[1], S.inv_ext[1], G.my, G.vox[1], mint, cy, sy, ty, dy);
    dda_axis(o.z, d.z, id.z, S.bmin[2], S.bmax[2], S.inv_ext[2], G.mz, G.vox[2], mint, cz, sz, tz, dz);
    int ci = (cz * G.my + cy) * G.mx + cx;
    const int stx = sx, sty = sy * G.mx, stz = sz * (G.mx * G.my);
    int rx = (sx > 0) ? (G.mx - 1 - cx) : cx;                         // steps left before the walk leaves the grid
    int ry = (sy > 0) ? (G.my - 1 - cy) : cy;
    int rz = (sz > 0) ? (G.mz - 1 - cz) : cz;
    float dda_t = mint, tau_acc = 0.f, majorant = 0.f;
    bool reached = false;
    while (true) {
        const float2 m = __ldg(G.mgrid + ci);
        const float t_next = fmin_(fmin_(tx, ty), tz);
        const bool a0 = (tx == t_next), a1 = !a0 && (ty == t_next);  // the axis whose boundary is crossed (else z)
        majorant = (SEL == 1) ? m.x : ((SEL == 2) ? m.y : fmax_(m.x, m.y));
        const float tau_next = tau_acc + (majorant * (t_next - dda_t));
        if ((majorant > 0.f) && (tau_next >= tau)) {                  // the division only where its result can be used
            float t_precise = dda_t + ((tau - tau_acc) / majorant);
            if (t_precise < maxt) { dda_t = t_precise; reached = true; break; }
        }
        dda_t = t_next;
        if (!(t_next < maxt)) break;
        const int left = a0 ? rx : (a1 ? ry : rz);
        if (left == 0) break;
        const bool a2 = !a0 && !a1;
        rx -= a0 ? 1 : 0; ry -= a1 ? 1 : 0; rz -= a2 ? 1 : 0;
        tx = a0 ? (tx + dx) : tx; ty = a1 ? (ty + dy) : ty; tz = a2 ? (tz + dz) : tz;
        ci += a0 ? stx : (a1 ? sty : stz);
        tau_acc = tau_next;
    }
    maj_out = majorant;
    return reached ? dda_t : inf_();
}

// the walk of the specification as written (K = 1), or K cells at a time
template <int SEL>
__device__ __forceinline__ float dda_sample(const SceneK &S, const GridKS &G, V3 o, V3 d, V3 id, float mint, float maxt, float tau,
                                            float &maj_out) {
    constexpr int K = RCV_DDA_AHEAD;
    float tx, ty, tz, dx, dy, dz;
    int cx, cy, cz, sx, sy, sz;
    dda_axis(o.x, d.x, id.x, S.bmin[0], S.bmax[0], S.inv_ext[0], G.mx, G.vox[0], mint, cx, sx, tx, dx);
    dda_axis(o.y, d.y, id.y, S.bmin[1], S.bmax[1], S.inv_ext[1], G.my, G.vox[1], mint, cy, sy, ty, dy);
    dda_axis(o.z, d.z, id.z, S.bmin[2], S.bmax[2], S.inv_ext[2], G.mz, G.vox[2], mint, cz, sz, tz, dz);
    float dda_t = mint, tau_acc = 0.f, majorant = 0.f;
    bool reached = false, done = false;
    while (!done) {
        float tn[K], mj[K];
        bool last[K];
        bool open = true;                                             // the walk has not left the grid / the interval yet
#pragma unroll
        for (int k = 0; k < K; ++k) {
            tn[k] = inf_(); mj[k] = 0.f; last[k] = true;
            if (open) {
                float t_next = fmin_(fmin_(tx, ty), tz);
                int axis = (tx == t_next) ? 0 : ((ty == t_next) ? 1 : 2);
                float2 m = __ldg(G.mgrid + ((size_t) cz * G.my + cy) * G.mx + cx);
                mj[k] = (SEL == 1) ? m.x : ((SEL == 2) ? m.y : fmax_(m.x, m.y));
                tn[k] = t_next;
                bool end = !(t_next < maxt);
                if (!end) {
                    if (axis == 0) { cx += sx; if (cx < 0 || cx >= G.mx) end = true; else tx = tx + dx; }
                    else if (axis == 1) { cy += sy; if (cy < 0 || cy >= G.my) end = true; else ty = ty + dy; }
                    else { cz += sz; if (cz < 0 || cz >= G.mz) end = true; else tz = tz + dz; }
                }
                last[k] = end;
                open = !end;
            }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (!done) {
                majorant = mj[k];
                float tau_next = tau_acc + (majorant * (tn[k] - dda_t));
                if ((majorant > 0.f) && (tau_next >= tau)) {          // the division only where its result can be used
                    float t_precise = dda_t + ((tau - tau_acc) / majorant);
                    if (t_precise < maxt) { dda_t = t_precise; reached = true; done = true; }
                }
                if (!done) {
                    dda_t = tn[k];
                    if (last[k]) done = true;
                    else tau_acc = tau_next;
                }
            }
        }
    }
    maj_out = majorant;
    return reached ? dda_t : inf_();
}

// ----------------------------------------------------------------------------
// speculative L2 prefetch of the NEXT tentative collision's gather brick.
// After a null collision the lane continues from p along d; its next free-flight
// distance depends only on its random stream (`skip` draws are consumed before the
// distance sample), not on the lookup that is still in flight.  The stream is
// advanced on a copy; nothing observable changes (results are bit-identical with
// or without the prefetch).
// ----------------------------------------------------------------------------
#ifndef RCV_PREFETCH
#define RCV_PREFETCH 0   // measured slower on B200 (profiles/README.md): the extra instructions cost more than the latency they hide
#endif
__device__ __forceinline__ void prefetch_l2(const void *p) {
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void) p;
#endif
}
template <bool TWO>
__device__ __forceinline__ void prefetch_next_brick(const SceneK &S, const GridK &G, float inv_mu, Pcg rng, int skip, V3 p, V3 d) {
#if RCV_PREFETCH
    for (int i = 0; i < skip; ++i) (void) pcg_u32(rng);
    float u = pcg_f32(rng);
    float t = (-spec_log(1.f - u)) * inv_mu;
    Tri T = tri_setup(S, G, ray_at(p, d, t));
    if (T.inside) {
        if (TWO) prefetch_l2(G.blk2 + 4 * (size_t) tri_cell(G, T));
        else prefetch_l2(G.blk1 + 2 * (size_t) tri_cell(G, T));
    }
#endif
}

// ----------------------------------------------------------------------------
// sensor: lane -> (rng, camera ray, film pixel)
// ----------------------------------------------------------------------------
struct Lane {
    Pcg rng;
    V3 o, d;
    int pix;   // -1: the sample fell outside the film
    float fx, fy;   // film position in pixels
};
__device__ __forceinline__ Lane lane_setup(const SceneK &S, uint32_t seed, uint32_t spp, uint64_t idx) {
    Lane L;
    sampler_seed(L.rng, seed, (uint32_t) idx);
    uint32_t pixel = (uint32_t) (idx / spp);
    uint32_t py = pixel / S.W, px = pixel - S.W * py;
    float u1 = pcg_f32(L.rng), u2 = pcg_f32(L.rng);
    float pfx = (float) px + u1, pfy = (float) py + u2;
    int ix = (int) floorf(pfx), iy = (int) floorf(pfy);
    L.pix = (ix >= 0 && ix < (int) S.W && iy >= 0 && iy < (int) S.H) ? iy * (int) S.W + ix : -1;
    L.fx = pfx; L.fy = pfy;
    float sx = pfx * S.rcpW, sy = pfy * S.rcpH;
    const float *A = S.cam_axes, *O = S.cam_o;
    if (S.n_sensors > 1) {                                 // batch.cpp sample_ray: view by film column
        float fs = sx * (float) S.n_sensors;
        uint32_t iu = (uint32_t) fs;
        int si = (int) (iu < (uint32_t) (S.n_sensors - 1) ? iu : (uint32_t) (S.n_sensors - 1));
        sx = fs - (float) iu;
        A = S.batch_axes + 9 * si; O = S.batch_o + 3 * si;
    }
    float nx = __fmaf_rn(sx, -2.f * S.tan_x, S.tan_x);
    float ny = __fmaf_rn(sy, -2.f * S.tan_y, S.tan_y);
    float len2 = __fmaf_rn(nx, nx, __fmaf_rn(ny, ny, 1.f));
    float inv = 1.f / sqrtf(len2);
    float dx = nx * inv, dy = ny * inv, dz = inv;
    L.d = mk3(__fmaf_rn(A[0], dx, __fmaf_rn(A[1], dy, A[2] * dz)),
              __fmaf_rn(A[3], dx, __fmaf_rn(A[4], dy, A[5] * dz)),
              __fmaf_rn(A[6], dx, __fmaf_rn(A[7], dy, A[8] * dz)));
    float near_t = S.near_clip / dz;
    L.o = ray_at(mk3(O[0], O[1], O[2]), L.d, near_t);
    return L;
}

// the film position of lane idx alone (the first two draws of its stream)
__device__ __forceinline__ void film_position(const SceneK &S, uint32_t seed, uint32_t spp, uint32_t idx, float &fx, float &fy) {
    Pcg rng;
    sampler_seed(rng, seed, idx);
    uint32_t pixel = idx / spp;
    uint32_t py = pixel / S.W, px = pixel - S.W * py;
    float u1 = pcg_f32(rng), u2 = pcg_f32(rng);
    fx = (float) px + u1; fy = (float) py + u2;
}

// ----------------------------------------------------------------------------
// ImageBlock.put: f(pixel, weight) for every pixel a sample at film position (fx, fy) is splatted to.
// Box: the pixel that contains it, weight 1.  Gaussian (hdrfilm default): radius 4 stddev, separable weights
// max(0, exp(a x^2) - exp(a r^2)) at the pixel centres within the radius (at most 4 x 4).
// ----------------------------------------------------------------------------
// the gaussian filter's separable weights of a sample at film position (fx, fy): wx[i] / wy[j] for pixel columns lox + i /
// rows loy + j, i, j < count (0 beyond)
__device__ __forceinline__ void film_gauss_weights(const SceneK &S, float fx, float fy, int &lox, int &loy, int &count, float (&wx)[4],
                                                   float (&wy)[4]) {
    const float sd = S.rf_stddev, r = 4.f * sd;
    const float alpha = -1.f / ((2.f * sd) * sd);
    const float bias = spec_exp(alpha * (r * r));
    count = 2 * (int) ceilf(r);
    const float px = fx - 0.5f, py = fy - 0.5f;
    lox = (int) ceilf(px - r); loy = (int) ceilf(py - r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float dx = (float) (lox + i) - px, dy = (float) (loy + i) - py;
        wx[i] = (i < count) ? fmax_(0.f, spec_exp(alpha * (dx * dx)) - bias) : 0.f;
        wy[i] = (i < count) ? fmax_(0.f, spec_exp(alpha * (dy * dy)) - bias) : 0.f;
    }
}
template <class F>
__device__ __forceinline__ void film_for_each(const SceneK &S, float fx, float fy, F f) {
    const int W = (int) S.W, H = (int) S.H;
    if (S.rfilter != 1) {
        int ix = (int) floorf(fx), iy = (int) floorf(fy);
        if (ix >= 0 && ix < W && iy >= 0 && iy < H) f(iy * W + ix, 1.f);
        return;
    }
    int lox, loy, count;
    float wx[4], wy[4];
    film_gauss_weights(S, fx, fy, lox, loy, count, wx, wy);
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int x = lox + i, y = loy + j;
            if (i < count && j < count && x >= 0 && x < W && y >= 0 && y < H) f(y * W + x, wy[j] * wx[i]);
        }
}

// ----------------------------------------------------------------------------
// vector reductions into global memory (sm_90+: red.global.add.v4.f32)
// ----------------------------------------------------------------------------
#if defined(RCV_HOST_EMU) && !defined(__CUDACC__)
static inline void red_add_v4(float4 *addr, float a, float b, float c, float d) {
    addr->x += a; addr->y += b; addr->z += c; addr->w += d;
}
static inline void red_add_f32(float *addr, float a) { *addr += a; }
#else
__device__ __forceinline__ void red_add_v4(float4 *addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_f32(float *addr, float a) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}

// warp-aggregated counter flush
__device__ __forceinline__ void counter_flush(Counters *C, int slot, unsigned v) {
    unsigned s = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(&C->c[slot], (unsigned long long) s);
}
#endif

// d(develop o put)/dL * grad_in for a sample at film position (fx, fy): sum over its footprint of w * grad_in[p] / W[p]
__device__ __forceinline__ V3 film_gather(const SceneK &S, const float *grad_img, const float *film_w, float fx, float fy) {
    float ax = 0.f, ay = 0.f, az = 0.f;
    film_for_each(S, fx, fy, [&](int p, float w) {
        float fw = film_w[p];
        float dv = (fw == 0.f) ? 1.f : fw;
        ax += w * (grad_img[3 * p] / dv);
        ay += w * (grad_img[3 * p + 1] / dv);
        az += w * (grad_img[3 * p + 2] / dv);
    });
    return mk3(ax, ay, az);
}

// Out-of-line forms for the gaussian filter, so that the (default) box path of the tracing kernels keeps its register
// allocation: they run once per path, in the slow step.  Values travel in registers (no reference parameters).
static __device__ __noinline__ void film_splat_filtered(const SceneK &S, float4 *acc, float fx, float fy, float lx, float ly, float lz) {
    film_for_each(S, fx, fy, [&](int p, float w) { red_add_v4(acc + p, lx * w, ly * w, lz * w, w); });
}
// the same for the lane whose film position is recomputed from its stream
static __device__ __noinline__ void film_splat_filtered_lane(const SceneK &S, float4 *acc, uint32_t seed, uint32_t spp, uint32_t idx,
                                                             float lx, float ly, float lz) {
    float fx, fy;
    film_position(S, seed, spp, idx, fx, fy);
    film_for_each(S, fx, fy, [&](int p, float w) { red_add_v4(acc + p, lx * w, ly * w, lz * w, w); });
}
}  // namespace rcv
