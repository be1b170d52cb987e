// rcv_pixel.cu — streaming (HBM-bound) kernels: medium rotation + majorant,
// film develop, the fused recursive-CV / EMA-Welford update, loss + image
// gradient, gradient unpack, Adam + clamp, nearest x2 upsample.
// All of them are grid-stride, coalesced, float4-vectorised where the layout
// allows; grids are sized to a multiple of the SM count.
#include "rcv_internal.h"

namespace rcv {

static constexpr int kBlock = 256;

static inline unsigned grid_for(const LaunchCfg &cfg, uint64_t n_items, int per_sm = 8) {
    uint64_t want = (n_items + kBlock - 1) / kBlock;
    uint64_t cap = (uint64_t) cfg.sm_count * per_sm;
    uint64_t g = want < cap ? want : cap;
    return (unsigned) (g < 1 ? 1 : g);
}

// ---------------------------------------------------------------------------
// TwoStateMedium.parameters_changed (plugins/twostatemedium.py:189-200):
// old <- new, new <- incoming, fused with the max-reduction for the new
// majorant (gridvolume max, recomputed by new_params.update()).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_rotate(const float *__restrict__ sigma_in, const float *__restrict__ albedo_in,
                                                   size_t nvox, float2 *__restrict__ sig2, float4 *__restrict__ alb2,
                                                   float *__restrict__ block_max) {
    float mx = -inf_();
    for (size_t v = (size_t) blockIdx.x * kBlock + threadIdx.x; v < nvox; v += (size_t) gridDim.x * kBlock) {
        float s_in = __ldg(sigma_in + v);
        float2 s = sig2[v];
        sig2[v] = make_float2(s.y, s_in);
        float4 a_new = alb2[2 * v + 1];
        alb2[2 * v] = a_new;
        alb2[2 * v + 1] = make_float4(__ldg(albedo_in + 3 * v), __ldg(albedo_in + 3 * v + 1), __ldg(albedo_in + 3 * v + 2), 0.f);
        mx = fmax_(mx, s_in);
    }
    __shared__ float red[kBlock / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax_(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = red[0];
        for (int i = 1; i < kBlock / 32; ++i) m = fmax_(m, red[i]);
        block_max[blockIdx.x] = m;
    }
}
// the same, four voxels per thread with 16-byte accesses only (needs 16-byte aligned inputs; the scalar kernel above
// takes the tail and unaligned callers).  block_max slots [gridDim.x, 2 gridDim.x) of the scalar launch are not used.
__global__ void __launch_bounds__(kBlock) k_rotate4(const float4 *__restrict__ sigma_in4, const float4 *__restrict__ albedo_in4,
                                                    size_t nquad, float4 *__restrict__ sig2q, float4 *__restrict__ alb2,
                                                    float *__restrict__ block_max) {
    float mx = -inf_();
    for (size_t q = (size_t) blockIdx.x * kBlock + threadIdx.x; q < nquad; q += (size_t) gridDim.x * kBlock) {
        const float4 s_in = __ldg(sigma_in4 + q);
        const float4 a0 = __ldg(albedo_in4 + 3 * q), a1 = __ldg(albedo_in4 + 3 * q + 1), a2 = __ldg(albedo_in4 + 3 * q + 2);
        const float4 s01 = sig2q[2 * q], s23 = sig2q[2 * q + 1];          // (old0, new0, old1, new1), (old2, new2, old3, new3)
        float4 n0 = alb2[8 * q + 1], n1 = alb2[8 * q + 3], n2 = alb2[8 * q + 5], n3 = alb2[8 * q + 7];
        sig2q[2 * q] = make_float4(s01.y, s_in.x, s01.w, s_in.y);
        sig2q[2 * q + 1] = make_float4(s23.y, s_in.z, s23.w, s_in.w);
        alb2[8 * q] = n0; alb2[8 * q + 1] = make_float4(a0.x, a0.y, a0.z, 0.f);
        alb2[8 * q + 2] = n1; alb2[8 * q + 3] = make_float4(a0.w, a1.x, a1.y, 0.f);
        alb2[8 * q + 4] = n2; alb2[8 * q + 5] = make_float4(a1.z, a1.w, a2.x, 0.f);
        alb2[8 * q + 6] = n3; alb2[8 * q + 7] = make_float4(a2.y, a2.z, a2.w, 0.f);
        mx = fmax_(mx, fmax_(fmax_(s_in.x, s_in.y), fmax_(s_in.z, s_in.w)));
    }
    __shared__ float red[kBlock / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax_(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = red[0];
        for (int i = 1; i < kBlock / 32; ++i) m = fmax_(m, red[i]);
        block_max[blockIdx.x] = m;
    }
}
__global__ void k_majorant_final(const float *block_max, int n_blocks, float scale, float *majorants) {
    // one warp; majorants[0] (old) <- majorants[1] (new); new <- scale * max
    float mx = -inf_();
    for (int i = threadIdx.x; i < n_blocks; i += 32) mx = fmax_(mx, block_max[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax_(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (threadIdx.x == 0) {
        majorants[0] = majorants[1];
        majorants[1] = scale * mx;
    }
}
__device__ __forceinline__ void st_global_v8(float4 *p, float a, float b, float c, float d, float e, float f, float g, float h) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d), "f"(e), "f"(f), "f"(g), "f"(h)
                 : "memory");
}
// gather bricks: the 8 corner values of every interpolation cell, contiguous (rcv_device.cuh, GridK)
__global__ void __launch_bounds__(kBlock) k_build_bricks(const float2 *__restrict__ sig2, int nx, int ny, int nz,
                                                         float4 *__restrict__ blk2, float4 *__restrict__ blk1) {
    const size_t ncell = (size_t) (nx + 1) * (ny + 1) * (nz + 1);
    for (size_t c = (size_t) blockIdx.x * kBlock + threadIdx.x; c < ncell; c += (size_t) gridDim.x * kBlock) {
        size_t r = c;
        int cx = (int) (r % (nx + 1)); r /= (nx + 1);
        int cy = (int) (r % (ny + 1)); r /= (ny + 1);
        int cz = (int) r;
        float2 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = __ldg(sig2 + cell_corner(nx, ny, nz, cx, cy, cz, k));
        // 256-bit stores (sm_100 STG.256): every store writes whole 32-byte sectors — three per cell instead of six halves
        st_global_v8(blk2 + 4 * c, v[0].x, v[0].y, v[1].x, v[1].y, v[2].x, v[2].y, v[3].x, v[3].y);
        st_global_v8(blk2 + 4 * c + 2, v[4].x, v[4].y, v[5].x, v[5].y, v[6].x, v[6].y, v[7].x, v[7].y);
        st_global_v8(blk1 + 2 * c, v[0].y, v[1].y, v[2].y, v[3].y, v[4].y, v[5].y, v[6].y, v[7].y);
    }
}

cudaError_t launch_rotate(const LaunchCfg &cfg, const float *sigma_in, const float *albedo_in, int nz, int ny, int nx, float scale,
                          float2 *sig2, float4 *blk2, float4 *blk1, float4 *alb2, float *block_max, int n_blocks,
                          float *majorants) {
    size_t nvox = (size_t) nz * ny * nx, ncell = (size_t) (nz + 1) * (ny + 1) * (nx + 1);
    const bool aligned = ((((uintptr_t) sigma_in) | ((uintptr_t) albedo_in)) & 15u) == 0 && nvox >= 4;
    if (aligned && n_blocks >= 2) {
        // four voxels per thread; the (< 4 voxel) tail goes through the scalar kernel with one block
        const size_t nquad = nvox / 4, tail = nvox - 4 * nquad;
        const int nb = n_blocks - 1;
        k_rotate4<<<nb, kBlock, 0, cfg.stream>>>((const float4 *) sigma_in, (const float4 *) albedo_in, nquad, (float4 *) sig2, alb2,
                                                 block_max);
        k_rotate<<<1, kBlock, 0, cfg.stream>>>(sigma_in + 4 * nquad, albedo_in + 12 * nquad, tail, sig2 + 4 * nquad, alb2 + 8 * nquad,
                                               block_max + nb);
        (*cfg.launches)++;
    } else {
        k_rotate<<<n_blocks, kBlock, 0, cfg.stream>>>(sigma_in, albedo_in, nvox, sig2, alb2, block_max);
    }
    k_majorant_final<<<1, 32, 0, cfg.stream>>>(block_max, n_blocks, scale, majorants);
    k_build_bricks<<<grid_for(cfg, ncell, 16), kBlock, 0, cfg.stream>>>(sig2, nx, ny, nz, blk2, blk1);
    (*cfg.launches) += 3;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// majorant supergrid (heterogeneous `majorant_resolution_factor`, upstream update_majorant_supergrid):
// cell (cx,cy,cz) of f^3 voxels stores scale * max over the voxels [c f - 1, c f + f] per axis (clamped) — every
// voxel a trilinear lookup inside the cell can touch — for both states.  One warp per cell.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_majorant_grid(const float2 *__restrict__ sig2, int nx, int ny, int nz, int f, float scale,
                                                          int mx, int my, int mz, float2 *__restrict__ mgrid) {
    const int lane = threadIdx.x & 31;
    const size_t ncell = (size_t) mx * my * mz, span = (size_t) (f + 2);
    for (size_t c = ((size_t) blockIdx.x * kBlock + threadIdx.x) >> 5; c < ncell; c += ((size_t) gridDim.x * kBlock) >> 5) {
        int cx = (int) (c % mx), cy = (int) ((c / mx) % my), cz = (int) (c / ((size_t) mx * my));
        float mo = 0.f, mn = 0.f;
        for (size_t k = lane; k < span * span * span; k += 32) {
            int x = cx * f - 1 + (int) (k % span), y = cy * f - 1 + (int) ((k / span) % span), z = cz * f - 1 + (int) (k / (span * span));
            x = min(max(x, 0), nx - 1); y = min(max(y, 0), ny - 1); z = min(max(z, 0), nz - 1);
            float2 s = __ldg(sig2 + ((size_t) z * ny + y) * nx + x);
            mo = fmaxf(mo, s.x); mn = fmaxf(mn, s.y);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mo = fmaxf(mo, __shfl_xor_sync(0xffffffffu, mo, o));
            mn = fmaxf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        }
        if (lane == 0) mgrid[c] = make_float2(scale * mo, scale * mn);
    }
}
cudaError_t launch_majorant_grid(const LaunchCfg &cfg, const float2 *sig2, int nz, int ny, int nx, int factor, float scale,
                                 float2 *mgrid) {
    int mx = nx / factor, my = ny / factor, mz = nz / factor;
    k_majorant_grid<<<grid_for(cfg, (size_t) mx * my * mz * 32), kBlock, 0, cfg.stream>>>(sig2, nx, ny, nz, factor, scale, mx, my, mz,
                                                                                          mgrid);
    (*cfg.launches)++;
    return cudaGetLastError();
}

__global__ void __launch_bounds__(kBlock) k_medium_read(const float2 *__restrict__ sig2, const float4 *__restrict__ alb2,
                                                        size_t nvox, int state, float *sigma_out, float *albedo_out) {
    for (size_t v = (size_t) blockIdx.x * kBlock + threadIdx.x; v < nvox; v += (size_t) gridDim.x * kBlock) {
        float2 s = sig2[v];
        float4 a = alb2[2 * v + state];
        if (sigma_out) sigma_out[v] = state ? s.y : s.x;
        if (albedo_out) { albedo_out[3 * v] = a.x; albedo_out[3 * v + 1] = a.y; albedo_out[3 * v + 2] = a.z; }
    }
}
cudaError_t launch_medium_read(const LaunchCfg &cfg, const float2 *sig2, const float4 *alb2, size_t nvox, int state,
                               float *sigma_out, float *albedo_out) {
    k_medium_read<<<grid_for(cfg, nvox), kBlock, 0, cfg.stream>>>(sig2, alb2, nvox, state, sigma_out, albedo_out);
    (*cfg.launches)++;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// hdrfilm.develop(): RGB / W
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_develop(const float4 *__restrict__ accum, float *__restrict__ img, size_t npix) {
    for (size_t p = (size_t) blockIdx.x * kBlock + threadIdx.x; p < npix; p += (size_t) gridDim.x * kBlock) {
        float4 a = __ldg(accum + p);
        float d = (a.w == 0.f) ? 1.f : a.w;
        img[3 * p] = a.x / d; img[3 * p + 1] = a.y / d; img[3 * p + 2] = a.z / d;
    }
}
cudaError_t launch_develop(const LaunchCfg &cfg, const float4 *accum, float *img, size_t npix) {
    k_develop<<<grid_for(cfg, npix), kBlock, 0, cfg.stream>>>(accum, img, npix);
    (*cfg.launches)++;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// recursive control variate + EMA-Welford statistics, fused
// (plugins/cv_integrator.py:58-77 / plugins/metaintegrator.py:93-113,
//  plugins/welford.py:59-78).  One pass over the ten state arrays.
// ---------------------------------------------------------------------------
struct CvElem {
    float H, F, img, v_n, w_s, mx, my, vx, vy, cov;
};
__device__ __forceinline__ void cv_stats_update(CvElem &e, const CvArgs &a) {
    float x = e.H, y = e.F;
    float dx1 = a.stats_n_is_zero ? x : x - e.mx / a.c1a;              // welford.py:61
    float dy1 = a.stats_n_is_zero ? y : y - e.my / a.c1a;              // :62
    e.my = a.beta1 * e.my + (1.f - a.beta1) * y;                       // :64
    e.mx = a.beta1 * e.mx + (1.f - a.beta1) * x;                       // :65
    float dx2 = x - e.mx / a.c1b;                                      // :68
    float dy2 = y - e.my / a.c1b;                                      // :69
    e.vx = a.beta2 * e.vx + ((1.f - a.beta2) * dx1) * dx2;             // :71
    e.vy = a.beta2 * e.vy + ((1.f - a.beta2) * dy1) * dy2;             // :72
    e.cov = a.beta2 * e.cov + ((1.f - a.beta2) * dx1) * dy2;           // :73
}
__device__ __forceinline__ void cv_elem(CvElem &e, const CvArgs &a) {
    if (a.do_update && a.pre_update) cv_stats_update(e, a);            // metaintegrator.py:95-96
    if (a.do_weights) {                                                // cv_integrator.py:61-68
        float v0 = e.vx / a.c2, v1 = e.vy / a.c2, cv = e.cov / a.c2;   // welford.py:77-78
        float ws = e.w_s;
        float vn = ((ws * ws) * (e.v_n + v0) + v1) - ((2.f * ws) * cv);
        e.v_n = vn;
        ws = cv / (v0 + vn);
        if (isnan(ws) || isinf(ws)) ws = 0.f;
        ws = fmin_(fmax_(ws, 0.f), 1.f);
        e.w_s = ws;
    }
    e.img = e.w_s * (e.img - e.H) + e.F;                               // cv_integrator.py:70
    if (a.do_update && !a.pre_update) cv_stats_update(e, a);           // :73-74
}
template <bool VEC4>
__global__ void __launch_bounds__(kBlock) k_cv_update(const __grid_constant__ CvArgs a) {
    if (VEC4) {
        size_t n4 = a.n / 4;
        for (size_t i = (size_t) blockIdx.x * kBlock + threadIdx.x; i < n4; i += (size_t) gridDim.x * kBlock) {
            float4 H = __ldg((const float4 *) a.H + i), F = __ldg((const float4 *) a.F + i);
            float4 img = ((float4 *) a.img)[i], v_n = ((float4 *) a.v_n)[i], w_s = ((float4 *) a.w_s)[i];
            float4 mx = ((float4 *) a.mean_x)[i], my = ((float4 *) a.mean_y)[i];
            float4 vx = ((float4 *) a.var_x)[i], vy = ((float4 *) a.var_y)[i], cv = ((float4 *) a.cov)[i];
            CvElem e[4] = {{H.x, F.x, img.x, v_n.x, w_s.x, mx.x, my.x, vx.x, vy.x, cv.x},
                           {H.y, F.y, img.y, v_n.y, w_s.y, mx.y, my.y, vx.y, vy.y, cv.y},
                           {H.z, F.z, img.z, v_n.z, w_s.z, mx.z, my.z, vx.z, vy.z, cv.z},
                           {H.w, F.w, img.w, v_n.w, w_s.w, mx.w, my.w, vx.w, vy.w, cv.w}};
#pragma unroll
            for (int k = 0; k < 4; ++k) cv_elem(e[k], a);
            ((float4 *) a.img)[i] = make_float4(e[0].img, e[1].img, e[2].img, e[3].img);
            if (a.do_weights) {
                ((float4 *) a.v_n)[i] = make_float4(e[0].v_n, e[1].v_n, e[2].v_n, e[3].v_n);
                ((float4 *) a.w_s)[i] = make_float4(e[0].w_s, e[1].w_s, e[2].w_s, e[3].w_s);
            }
            if (a.do_update) {
                ((float4 *) a.mean_x)[i] = make_float4(e[0].mx, e[1].mx, e[2].mx, e[3].mx);
                ((float4 *) a.mean_y)[i] = make_float4(e[0].my, e[1].my, e[2].my, e[3].my);
                ((float4 *) a.var_x)[i] = make_float4(e[0].vx, e[1].vx, e[2].vx, e[3].vx);
                ((float4 *) a.var_y)[i] = make_float4(e[0].vy, e[1].vy, e[2].vy, e[3].vy);
                ((float4 *) a.cov)[i] = make_float4(e[0].cov, e[1].cov, e[2].cov, e[3].cov);
            }
        }
    } else {
        for (size_t i = (size_t) blockIdx.x * kBlock + threadIdx.x; i < a.n; i += (size_t) gridDim.x * kBlock) {
            CvElem e = {a.H[i], a.F[i], a.img[i], a.v_n[i], a.w_s[i], a.mean_x[i], a.mean_y[i], a.var_x[i], a.var_y[i], a.cov[i]};
            cv_elem(e, a);
            a.img[i] = e.img;
            if (a.do_weights) { a.v_n[i] = e.v_n; a.w_s[i] = e.w_s; }
            if (a.do_update) { a.mean_x[i] = e.mx; a.mean_y[i] = e.my; a.var_x[i] = e.vx; a.var_y[i] = e.vy; a.cov[i] = e.cov; }
        }
    }
}
static bool aligned16(const void *p) { return ((uintptr_t) p & 15u) == 0; }
cudaError_t launch_cv_update(const LaunchCfg &cfg, const CvArgs &a) {
    bool vec = (a.n % 4 == 0) && aligned16(a.H) && aligned16(a.F) && aligned16(a.img) && aligned16(a.v_n) && aligned16(a.w_s) &&
               aligned16(a.mean_x) && aligned16(a.mean_y) && aligned16(a.var_x) && aligned16(a.var_y) && aligned16(a.cov);
    if (vec) k_cv_update<true><<<grid_for(cfg, a.n / 4), kBlock, 0, cfg.stream>>>(a);
    else k_cv_update<false><<<grid_for(cfg, a.n), kBlock, 0, cfg.stream>>>(a);
    (*cfg.launches)++;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// loss (utils.py:21-25) and d loss / d img; deterministic two-stage reduction
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_loss_partial(int kind, const float *__restrict__ img, const float *__restrict__ ref,
                                                         uint64_t n, float inv_n, double *partial, float *grad) {
    double acc = 0.0;
    for (uint64_t i = (uint64_t) blockIdx.x * kBlock + threadIdx.x; i < n; i += (uint64_t) gridDim.x * kBlock) {
        float d = img[i] - __ldg(ref + i);
        if (kind == 0) {
            acc += (double) fabsf(d);
            if (grad) grad[i] = ((d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f)) * inv_n;
        } else {
            acc += (double) d * (double) d;
            if (grad) grad[i] = (2.f * d) * inv_n;
        }
    }
    __shared__ double red[kBlock / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < kBlock / 32; ++i) s += red[i];
        partial[blockIdx.x] = s;
    }
}
__global__ void k_loss_final(const double *partial, int n_partial, double inv_n, float *loss_out) {
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < n_partial; ++i) s += partial[i];
        *loss_out = (float) (s * inv_n);
    }
}
cudaError_t launch_loss_grad(const LaunchCfg &cfg, int kind, const float *img, const float *ref, uint64_t n, double *partial,
                             int n_partial, float *loss_out, float *grad_img) {
    unsigned g = grid_for(cfg, n, 4);
    if ((int) g > n_partial) g = n_partial;
    k_loss_partial<<<g, kBlock, 0, cfg.stream>>>(kind, img, ref, n, 1.f / (float) n, partial, grad_img);
    k_loss_final<<<1, 32, 0, cfg.stream>>>(partial, (int) g, 1.0 / (double) n, loss_out);
    (*cfg.launches) += 2;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// gradient grid unpack, Adam (+ clamp), clamp, upsample
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_grad_unpack(const float *__restrict__ gsig, const float4 *__restrict__ galb, size_t nvox,
                                                        float *g_sigma, float *g_albedo) {
    for (size_t v = (size_t) blockIdx.x * kBlock + threadIdx.x; v < nvox; v += (size_t) gridDim.x * kBlock) {
        if (g_sigma) g_sigma[v] = __ldg(gsig + v);
        if (g_albedo) {
            float4 g = __ldg(galb + v);
            g_albedo[3 * v] = g.x; g_albedo[3 * v + 1] = g.y; g_albedo[3 * v + 2] = g.z;
        }
    }
}
cudaError_t launch_grad_unpack(const LaunchCfg &cfg, const GradK &D, size_t nvox, float *g_sigma, float *g_albedo) {
    k_grad_unpack<<<grid_for(cfg, nvox), kBlock, 0, cfg.stream>>>(D.gsig, D.galb, nvox, g_sigma, g_albedo);
    (*cfg.launches)++;
    return cudaGetLastError();
}

// RCV_GRAD_BRICKS: fold the per-cell gradient bricks of an adjoint launch into the voxel grid and leave them zero for
// the next launch.  One thread per cell; most cells are untouched (zero) and cost one 32-byte read.
__global__ void __launch_bounds__(kBlock) k_unbrick(float4 *__restrict__ gbrick, int nx, int ny, int nz, float *__restrict__ gsig) {
    const size_t ncell = (size_t) (nx + 1) * (ny + 1) * (nz + 1);
    for (size_t c = (size_t) blockIdx.x * kBlock + threadIdx.x; c < ncell; c += (size_t) gridDim.x * kBlock) {
        float4 a = gbrick[2 * c], b = gbrick[2 * c + 1];
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        bool any = false;
#pragma unroll
        for (int k = 0; k < 8; ++k) any = any || (v[k] != 0.f);
        if (!any) continue;
        size_t r = c;
        const int cx = (int) (r % (size_t) (nx + 1)); r /= (size_t) (nx + 1);
        const int cy = (int) (r % (size_t) (ny + 1)); r /= (size_t) (ny + 1);
        const int cz = (int) r;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (v[k] != 0.f) red_add_f32(gsig + cell_corner(nx, ny, nz, cx, cy, cz, k), v[k]);
        gbrick[2 * c] = make_float4(0.f, 0.f, 0.f, 0.f);
        gbrick[2 * c + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
cudaError_t launch_unbrick(const LaunchCfg &cfg, const GradK &D, int nx, int ny, int nz) {
    const size_t ncell = (size_t) (nx + 1) * (ny + 1) * (nz + 1);
    k_unbrick<<<grid_for(cfg, ncell), kBlock, 0, cfg.stream>>>(D.gbrick, nx, ny, nz, D.gsig);
    (*cfg.launches)++;
    return cudaGetLastError();
}

// mi.ad.Adam.step [UPSTREAM]: m,v EMA; p = fma(-lr_t / (sqrt(v) + eps), m, p)
__device__ __forceinline__ float adam_elem(float p, float g, float &m, float &v, float lr_t, float b1, float b2, float eps) {
    float m_t = b1 * m + (1.f - b1) * g;
    float v_t = b2 * v + ((1.f - b2) * g) * g;
    m = m_t; v = v_t;
    return __fmaf_rn(-lr_t / (sqrtf(v_t) + eps), m_t, p);
}
__global__ void __launch_bounds__(kBlock) k_adam_packed(const float *__restrict__ gsig, const float4 *__restrict__ galb, size_t nvox,
                                                        float *sigma, float *albedo,
                                                        float *m_s, float *v_s, float *m_a, float *v_a, float lr_t_s, float lr_t_a,
                                                        float b1, float b2, float eps, float slo, float shi, float alo, float ahi) {
    for (size_t v = (size_t) blockIdx.x * kBlock + threadIdx.x; v < nvox; v += (size_t) gridDim.x * kBlock) {
        float4 g = __ldg(galb + v);
        float ms = m_s[v], vs = v_s[v];
        float s = adam_elem(sigma[v], __ldg(gsig + v), ms, vs, lr_t_s, b1, b2, eps);
        m_s[v] = ms; v_s[v] = vs;
        sigma[v] = fmin_(fmax_(s, slo), shi);                          // utils.py:172-175
        float gg[3] = {g.x, g.y, g.z};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float ma = m_a[3 * v + c], va = v_a[3 * v + c];
            float a = adam_elem(albedo[3 * v + c], gg[c], ma, va, lr_t_a, b1, b2, eps);
            m_a[3 * v + c] = ma; v_a[3 * v + c] = va;
            albedo[3 * v + c] = fmin_(fmax_(a, alo), ahi);
        }
    }
}
// the same with 16-byte accesses only: sigma_t four voxels at a time, then albedo as a flat array of 3 nvox floats, four
// elements (of one or two voxels) at a time.  Needs nvox % 4 == 0 and 16-byte aligned arrays.
__global__ void __launch_bounds__(kBlock) k_adam_packed4(const float4 *__restrict__ gsig4, const float4 *__restrict__ galb, size_t nvox,
                                                         float4 *sigma4, float4 *albedo4, float4 *m_s4, float4 *v_s4, float4 *m_a4,
                                                         float4 *v_a4, float lr_t_s, float lr_t_a, float b1, float b2, float eps,
                                                         float slo, float shi, float alo, float ahi) {
    const size_t tid = (size_t) blockIdx.x * kBlock + threadIdx.x, stride = (size_t) gridDim.x * kBlock;
    for (size_t q = tid; q < nvox / 4; q += stride) {
        float4 p = sigma4[q], g = __ldg(gsig4 + q), m = m_s4[q], v = v_s4[q];
        p.x = fmin_(fmax_(adam_elem(p.x, g.x, m.x, v.x, lr_t_s, b1, b2, eps), slo), shi);
        p.y = fmin_(fmax_(adam_elem(p.y, g.y, m.y, v.y, lr_t_s, b1, b2, eps), slo), shi);
        p.z = fmin_(fmax_(adam_elem(p.z, g.z, m.z, v.z, lr_t_s, b1, b2, eps), slo), shi);
        p.w = fmin_(fmax_(adam_elem(p.w, g.w, m.w, v.w, lr_t_s, b1, b2, eps), slo), shi);
        sigma4[q] = p; m_s4[q] = m; v_s4[q] = v;
    }
    for (size_t q = tid; q < 3 * (nvox / 4); q += stride) {
        // elements 4q .. 4q+3 of the (nvox, 3) albedo array: voxel e / 3, channel e % 3
        const size_t e0 = 4 * q, v0 = e0 / 3;
        const int c0 = (int) (e0 - 3 * v0);
        const float4 ga = __ldg(galb + v0), gb = __ldg(galb + (v0 + 1 < nvox ? v0 + 1 : v0));
        float g0, g1, g2, g3;
        if (c0 == 0) { g0 = ga.x; g1 = ga.y; g2 = ga.z; g3 = gb.x; }
        else if (c0 == 1) { g0 = ga.y; g1 = ga.z; g2 = gb.x; g3 = gb.y; }
        else { g0 = ga.z; g1 = gb.x; g2 = gb.y; g3 = gb.z; }
        float4 p = albedo4[q], m = m_a4[q], v = v_a4[q];
        p.x = fmin_(fmax_(adam_elem(p.x, g0, m.x, v.x, lr_t_a, b1, b2, eps), alo), ahi);
        p.y = fmin_(fmax_(adam_elem(p.y, g1, m.y, v.y, lr_t_a, b1, b2, eps), alo), ahi);
        p.z = fmin_(fmax_(adam_elem(p.z, g2, m.z, v.z, lr_t_a, b1, b2, eps), alo), ahi);
        p.w = fmin_(fmax_(adam_elem(p.w, g3, m.w, v.w, lr_t_a, b1, b2, eps), alo), ahi);
        albedo4[q] = p; m_a4[q] = m; v_a4[q] = v;
    }
}
cudaError_t launch_adam_packed(const LaunchCfg &cfg, const GradK &D, size_t nvox, float *sigma, float *albedo, float *m_s,
                               float *v_s, float *m_a, float *v_a, float lr_t_s, float lr_t_a, float beta1, float beta2,
                               float eps, float slo, float shi, float alo, float ahi) {
    const uintptr_t bits = (uintptr_t) D.gsig | (uintptr_t) sigma | (uintptr_t) albedo | (uintptr_t) m_s | (uintptr_t) v_s |
                           (uintptr_t) m_a | (uintptr_t) v_a;
    if (nvox % 4 == 0 && (bits & 15u) == 0)
        k_adam_packed4<<<grid_for(cfg, nvox), kBlock, 0, cfg.stream>>>((const float4 *) D.gsig, D.galb, nvox, (float4 *) sigma,
                                                                      (float4 *) albedo, (float4 *) m_s, (float4 *) v_s, (float4 *) m_a,
                                                                      (float4 *) v_a, lr_t_s, lr_t_a, beta1, beta2, eps, slo, shi, alo, ahi);
    else
        k_adam_packed<<<grid_for(cfg, nvox), kBlock, 0, cfg.stream>>>(D.gsig, D.galb, nvox, sigma, albedo, m_s, v_s, m_a, v_a, lr_t_s,
                                                                       lr_t_a, beta1, beta2, eps, slo, shi, alo, ahi);
    (*cfg.launches)++;
    return cudaGetLastError();
}
__global__ void __launch_bounds__(kBlock) k_adam_flat(float *p, const float *__restrict__ g, float *m, float *v, uint64_t n,
                                                      float lr_t, float b1, float b2, float eps) {
    for (uint64_t i = (uint64_t) blockIdx.x * kBlock + threadIdx.x; i < n; i += (uint64_t) gridDim.x * kBlock) {
        float mm = m[i], vv = v[i];
        p[i] = adam_elem(p[i], __ldg(g + i), mm, vv, lr_t, b1, b2, eps);
        m[i] = mm; v[i] = vv;
    }
}
cudaError_t launch_adam_flat(const LaunchCfg &cfg, float *p, const float *g, float *m, float *v, uint64_t n, float lr_t,
                             float beta1, float beta2, float eps) {
    k_adam_flat<<<grid_for(cfg, n), kBlock, 0, cfg.stream>>>(p, g, m, v, n, lr_t, beta1, beta2, eps);
    (*cfg.launches)++;
    return cudaGetLastError();
}
__global__ void __launch_bounds__(kBlock) k_clamp(float *p, uint64_t n, float lo, float hi) {
    for (uint64_t i = (uint64_t) blockIdx.x * kBlock + threadIdx.x; i < n; i += (uint64_t) gridDim.x * kBlock)
        p[i] = fmin_(fmax_(p[i], lo), hi);
}
cudaError_t launch_clamp(const LaunchCfg &cfg, float *p, uint64_t n, float lo, float hi) {
    k_clamp<<<grid_for(cfg, n), kBlock, 0, cfg.stream>>>(p, n, lo, hi);
    (*cfg.launches)++;
    return cudaGetLastError();
}
// dr.upsample nearest x2 (utils.py:184)
__global__ void __launch_bounds__(kBlock) k_upsample2x(const float *__restrict__ in, float *out, int nz, int ny, int nx, int c) {
    size_t n_out = (size_t) 8 * nz * ny * nx * c;
    for (size_t i = (size_t) blockIdx.x * kBlock + threadIdx.x; i < n_out; i += (size_t) gridDim.x * kBlock) {
        size_t r = i;
        int k = (int) (r % c); r /= c;
        int x = (int) (r % (2 * nx)); r /= (2 * nx);
        int y = (int) (r % (2 * ny)); r /= (2 * ny);
        int z = (int) r;
        out[i] = __ldg(in + ((((size_t) (z >> 1) * ny + (y >> 1)) * nx + (x >> 1)) * c + k));
    }
}
cudaError_t launch_upsample2x(const LaunchCfg &cfg, const float *in, float *out, int nz, int ny, int nx, int c) {
    k_upsample2x<<<grid_for(cfg, (uint64_t) 8 * nz * ny * nx * c), kBlock, 0, cfg.stream>>>(in, out, nz, ny, nx, c);
    (*cfg.launches)++;
    return cudaGetLastError();
}

}  // namespace rcv
