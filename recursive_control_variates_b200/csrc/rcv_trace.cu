// rcv_trace.cu — tracing kernels: two-state primal (K1), volpathsimple primal
// (K5) and the fused pass-1 + pass-2 adjoint (K5+K6).
//
// Three schedules of the same per-lane estimators:
//   * WAVEFRONT (default)  a fully convergent ray-generation kernel writes one
//       48-byte record per path that enters the medium; persistent warps then
//       run the per-lane event machines (rcv_*_wf.cuh): every round a warp
//       refills its finished lanes from the record queue and executes ONE event
//       handler — the tentative-collision step (the only place with a grid
//       gather), the vertex step or the batched slow step — with the lanes
//       waiting for it.
//   * SIMPLE   one thread per lane, nested loops (A/B baseline).
//   * BATCH    persistent warps pulling 32-lane batches, nested loops (A/B baseline).
// Results are schedule-independent: each lane owns its PCG32 stream, indexed
// by the global lane id pixel*spp + k.
#include "rcv_internal.h"
#include "rcv_twostate.cuh"
#include "rcv_twostate_wf.cuh"
#include "rcv_volpathsimple.cuh"
#include "rcv_volpathsimple_wf.cuh"
#include "rcv_pool.cuh"
#include "rcv_sched.cuh"

namespace rcv {

static constexpr int kBlock = 256;

// ===========================================================================
// SIMPLE / BATCH schedules
// ===========================================================================
template <int MODE, class GRID>
__device__ __forceinline__ void twostates_lane(const SceneK &S, const GRID &G, uint32_t seed, uint32_t spp, uint64_t idx,
                                               float4 *acc_old, float4 *acc_new, unsigned &nc, unsigned &nsct) {
    Lane L = lane_setup(S, seed, spp, idx);
    V3 Lo, Ln;
    trace_twostates<MODE>(S, G, L.rng, L.o, L.d, Lo, Ln, nc, nsct);
    // ImageBlock.put: (L * weight, weight) added to the pixels of the filter footprint
    film_for_each(S, L.fx, L.fy, [&](int p, float w) {
        if (MODE != 2) red_add_v4(acc_old + p, Lo.x * w, Lo.y * w, Lo.z * w, w);
        if (MODE != 1) red_add_v4(acc_new + p, Ln.x * w, Ln.y * w, Ln.z * w, w);
    });
}

template <int MODE, class GRID>
__global__ void __launch_bounds__(kBlock) k_twostates_simple(const __grid_constant__ SceneK S, const __grid_constant__ GRID G,
                                                             uint32_t seed, uint32_t spp, uint32_t spp_begin, uint32_t ns,
                                                             uint64_t n_local, float4 *acc_old, float4 *acc_new, Counters *C,
                                                             unsigned long long *work, int persistent) {
    unsigned nc = 0, nsct = 0, np = 0;
    if (!persistent) {
        uint64_t j = (uint64_t) blockIdx.x * kBlock + threadIdx.x;
        if (j < n_local) {
            twostates_lane<MODE>(S, G, seed, spp, global_lane(j, spp, spp_begin, ns), acc_old, acc_new, nc, nsct);
            np = 1;
        }
    } else {
        const unsigned lane = threadIdx.x & 31;
        while (true) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(work, 32ull);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base >= n_local) break;
            uint64_t j = base + lane;
            if (j < n_local) {
                twostates_lane<MODE>(S, G, seed, spp, global_lane(j, spp, spp_begin, ns), acc_old, acc_new, nc, nsct);
                np++;
            }
            __syncwarp();
        }
    }
    counter_flush(C, MODE == 0 ? 0 : 2, nc);
    counter_flush(C, MODE == 0 ? 1 : 3, nsct);
    counter_flush(C, 6, np);
}

template <bool BACKWARD, class GRID>
__device__ __forceinline__ void vps_lane(const SceneK &S, const GRID &G, const GradK &D, uint32_t seed, uint32_t spp, uint64_t idx,
                                         uint32_t alt_seed, const float *grad_img, const float *film_w, float4 *accum,
                                         VpsCounts &cnt) {
    Lane L = lane_setup(S, seed, spp, idx);
    Pcg alt; alt.state = 0; alt.inc = 1;
    V3 res;
    Pcg rng1 = L.rng;                                   // RBIntegrator.render_backward: sampler.clone() for pass 1
    vps_trace<false>(S, G, D, rng1, alt, L.o, L.d, mk3(0, 0, 0), mk3(0, 0, 0), res, cnt);
    if (!BACKWARD) {
        film_for_each(S, L.fx, L.fy, [&](int p, float w) { red_add_v4(accum + p, res.x * w, res.y * w, res.z * w, w); });
        return;
    }
    // dL = d(develop o put)/dL * grad_in = sum over the filter footprint of w * grad_in[p] / W[p]
    V3 dL = mk3(0, 0, 0);
    film_for_each(S, L.fx, L.fy, [&](int p, float w) {
        float fw = __ldg(film_w + p);
        float dv = (fw == 0.f) ? 1.f : fw;
        dL.x += w * (__ldg(grad_img + 3 * p) / dv);
        dL.y += w * (__ldg(grad_img + 3 * p + 1) / dv);
        dL.z += w * (__ldg(grad_img + 3 * p + 2) / dv);
    });
    sampler_seed(alt, alt_seed, (uint32_t) idx);        // volpathsimple.py:119-120
    V3 res2;
    vps_trace<true>(S, G, D, L.rng, alt, L.o, L.d, dL, res, res2, cnt);
}

template <bool BACKWARD, class GRID>
__global__ void __launch_bounds__(kBlock) k_vps_simple(const __grid_constant__ SceneK S, const __grid_constant__ GRID G, GradK D,
                                                       uint32_t seed, uint32_t spp, uint32_t spp_begin, uint32_t ns, uint64_t n_local,
                                                       uint32_t alt_seed, const float *grad_img, const float *film_w, float4 *accum,
                                                       Counters *C, unsigned long long *work, int persistent) {
    VpsCounts cnt; cnt.coll = cnt.scat = cnt.g_sig = cnt.g_alb = 0;
    unsigned np = 0;
    if (!persistent) {
        uint64_t j = (uint64_t) blockIdx.x * kBlock + threadIdx.x;
        if (j < n_local) {
            vps_lane<BACKWARD>(S, G, D, seed, spp, global_lane(j, spp, spp_begin, ns), alt_seed, grad_img, film_w, accum, cnt);
            np = BACKWARD ? 2 : 1;
        }
    } else {
        const unsigned lane = threadIdx.x & 31;
        while (true) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(work, 32ull);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base >= n_local) break;
            uint64_t j = base + lane;
            if (j < n_local) {
                vps_lane<BACKWARD>(S, G, D, seed, spp, global_lane(j, spp, spp_begin, ns), alt_seed, grad_img, film_w, accum, cnt);
                np += BACKWARD ? 2 : 1;
            }
            __syncwarp();
        }
    }
    counter_flush(C, 2, cnt.coll);
    counter_flush(C, 3, cnt.scat);
    counter_flush(C, 4, cnt.g_sig);
    counter_flush(C, 5, cnt.g_alb);
    counter_flush(C, 6, np);
}

// use_drt = True (volpathsimple.py:33-47): the adjoint with the differential-ratio-tracking in-scattering term.  The DRT
// estimate is one more full path per lane, drawn from the secondary sampler, so it runs as a persistent one-thread-per-lane
// kernel under every tracer setting (the event machines above cover use_drt = False).
#ifndef RCV_DRT_BLOCK
#define RCV_DRT_BLOCK 128
#endif
#ifndef RCV_DRT_MINBLOCKS
#define RCV_DRT_MINBLOCKS 4
#endif
static constexpr int kDrtBlock = RCV_DRT_BLOCK;
template <class GRID>
__global__ void __launch_bounds__(kDrtBlock, RCV_DRT_MINBLOCKS) k_vps_drt(const __grid_constant__ SceneK S, const __grid_constant__ GRID G, GradK D,
                                                    uint32_t seed, uint32_t spp, uint32_t spp_begin, uint32_t ns, uint64_t n_local,
                                                    uint32_t alt_seed, const float *grad_img, const float *film_w, int drt_flags,
                                                    Counters *C, unsigned long long *work) {
    VpsCounts cnt; cnt.coll = cnt.scat = cnt.g_sig = cnt.g_alb = 0;
    unsigned np = 0;
    const unsigned lane = threadIdx.x & 31;
    while (true) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(work, 32ull);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n_local) break;
        uint64_t j = base + lane;
        if (j < n_local) {
            vps_backward_lane(S, G, D, seed, spp, global_lane(j, spp, spp_begin, ns), alt_seed, grad_img, film_w, drt_flags, cnt);
            np += 2;
        }
        __syncwarp();
    }
    counter_flush(C, 2, cnt.coll);
    counter_flush(C, 3, cnt.scat);
    counter_flush(C, 4, cnt.g_sig);
    counter_flush(C, 5, cnt.g_alb);
    counter_flush(C, 6, np);
}

// ===========================================================================
// WAVEFRONT schedule
// ===========================================================================
#ifndef RCV_WF_BLOCK
#define RCV_WF_BLOCK 128
#endif
#ifndef RCV_WF_MINBLOCKS
#define RCV_WF_MINBLOCKS 4
#endif
#ifndef RCV_TS_MINBLOCKS
#define RCV_TS_MINBLOCKS RCV_WF_MINBLOCKS
#endif
#ifndef RCV_VP_MINBLOCKS
#define RCV_VP_MINBLOCKS RCV_WF_MINBLOCKS
#endif
static constexpr int kWfBlock = RCV_WF_BLOCK;
static constexpr int kWfMinBlocks = RCV_WF_MINBLOCKS;


// ---------------------------------------------------------------------------
// Warp-aggregated film splats.  Lane indices are pixel-major (idx = pixel * spp + k), so when the number of samples a
// launch takes per pixel is a multiple of 32 a warp of the one-thread-per-lane kernels holds 32 samples of ONE pixel: their
// reconstruction-filter footprints (the pixel itself for the box filter, up to 4 x 4 pixels for the gaussian) fall into
// the 5 x 5 tile centred on that pixel, and the warp adds ONE value per tile cell instead of one per lane and footprint
// pixel (gaussian film weights: 25 reductions per warp instead of 512).
// ---------------------------------------------------------------------------
// v[k], k < 32, on every lane -> the sum over the warp of v[lane], on lane `lane` (31 shuffles)
__device__ __forceinline__ float warp_reduce_scatter32(float (&v)[32]) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int h = 16; h >= 1; h >>= 1) {
        const bool up = (lane & h) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
            const float send = up ? v[i] : v[i + h];
            const float keep = up ? v[i + h] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
        }
    }
    return v[0];
}
// this lane's filter weights on the 5 x 5 tile whose corner pixel is (tx, ty): w[5 * j + i] for pixel (tx + i, ty + j),
// the same products wy * wx film_for_each hands out.  false: the footprint does not fit the tile (never for samples of the
// tile's centre pixel; the caller then falls back to per-lane splats)
__device__ __forceinline__ bool film_tile_weights(const SceneK &S, float fx, float fy, int tx, int ty, float (&w)[32]) {
    float cx[5], cy[5];
    bool ok;
    if (S.rfilter != 1) {
        const int dx = (int) floorf(fx) - tx, dy = (int) floorf(fy) - ty;
        ok = dx >= 0 && dx < 5 && dy >= 0 && dy < 5;
#pragma unroll
        for (int i = 0; i < 5; ++i) { cx[i] = (i == dx) ? 1.f : 0.f; cy[i] = (i == dy) ? 1.f : 0.f; }
    } else {
        int lox, loy, count;
        float wx[4], wy[4];
        film_gauss_weights(S, fx, fy, lox, loy, count, wx, wy);
        const int dx = lox - tx, dy = loy - ty;
        ok = (dx == 0 || dx == 1) && (dy == 0 || dy == 1);
        cx[0] = dx == 0 ? wx[0] : 0.f; cy[0] = dy == 0 ? wy[0] : 0.f;
#pragma unroll
        for (int i = 1; i < 4; ++i) { cx[i] = dx == 0 ? wx[i] : wx[i - 1]; cy[i] = dy == 0 ? wy[i] : wy[i - 1]; }
        cx[4] = dx == 0 ? 0.f : wx[3]; cy[4] = dy == 0 ? 0.f : wy[3];
    }
#pragma unroll
    for (int c = 0; c < 32; ++c) w[c] = (c < 25) ? cy[c / 5] * cx[c % 5] : 0.f;
    return ok;
}
// true when the whole warp is live and its lanes are samples of one pixel (then *pixel is that pixel)
__device__ __forceinline__ bool warp_one_pixel(bool live, uint64_t j, uint32_t ns, uint32_t *pixel) {
    const uint32_t pix = (uint32_t) (j / ns);
    const uint32_t p0 = __shfl_sync(0xffffffffu, pix, 0);
    *pixel = p0;
    return __all_sync(0xffffffffu, live && pix == p0);
}

// --- ray generation: one thread per lane, no divergence to speak of -----------
template <int MODE>
__global__ void __launch_bounds__(kBlock) k_twostates_raygen(const __grid_constant__ SceneK S, uint32_t seed, uint32_t spp,
                                                             uint32_t spp_begin, uint32_t ns, uint64_t n_local, float4 *acc_old,
                                                             float4 *acc_new, PathQueue Q, Counters *C) {
    uint64_t j = (uint64_t) blockIdx.x * kBlock + threadIdx.x;
    const bool live = j < n_local;
    TsLane ln; ln.ev = EV_INIT; ln.fx = 0.f; ln.fy = 0.f; ln.pix = -1;
    ln.L_old = mk3(0, 0, 0); ln.L_new = mk3(0, 0, 0);
    TsSink sk; sk.acc_old = acc_old; sk.acc_new = acc_new; sk.n_coll = sk.n_scat = sk.n_paths = 0;
    sk.defer = true;                                      // paths that end here (they missed the medium) are splatted below
    if (live) ts_raygen<MODE>(S, ln, sk, seed, spp, global_lane(j, spp, spp_begin, ns));
    bool enter = live && ln.ev == EV_COLL;
    unsigned long long pos = queue_slot(Q.count, enter);
    if (enter) ts_store(ln, Q, pos);
    // film splats of the finished paths: one value per pixel of the 5 x 5 tile and warp when the warp's lanes are samples of
    // one pixel (see warp_one_pixel), else per lane.  Old and new radiance of such a path are the same number (unit
    // throughputs): one reduction serves both accumulators.
    const bool fin = live && ln.ev == EV_INIT;
    const V3 L = (MODE == 1) ? ln.L_old : ln.L_new;
    uint32_t pixel;
    bool one = (S.rfilter == 1) && warp_one_pixel(live, j, ns, &pixel);   // box filter: one splat per lane is cheaper than the tile
    const int py = (int) (pixel / S.W), px = (int) (pixel - (uint32_t) py * S.W);
    float w[32];
    if (one) one = __all_sync(0xffffffffu, film_tile_weights(S, ln.fx, ln.fy, px - 2, py - 2, w));
    if (one) {
        float v[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] = fin ? w[c] * L.x : 0.f;
        const float sx = warp_reduce_scatter32(v);
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] = fin ? w[c] * L.y : 0.f;
        const float sy = warp_reduce_scatter32(v);
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] = fin ? w[c] * L.z : 0.f;
        const float sz = warp_reduce_scatter32(v);
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] = fin ? w[c] : 0.f;
        const float sw = warp_reduce_scatter32(v);
        const int c = (int) (threadIdx.x & 31), x = px - 2 + c % 5, y = py - 2 + c / 5;
        if (c < 25 && sw != 0.f && x >= 0 && x < (int) S.W && y >= 0 && y < (int) S.H) {
            const int p = y * (int) S.W + x;
            if (MODE != 2) red_add_v4(acc_old + p, sx, sy, sz, sw);
            if (MODE != 1) red_add_v4(acc_new + p, sx, sy, sz, sw);
        }
    } else if (fin) {
        film_for_each(S, ln.fx, ln.fy, [&](int p, float wt) {
            if (MODE != 2) red_add_v4(acc_old + p, ln.L_old.x * wt, ln.L_old.y * wt, ln.L_old.z * wt, wt);
            if (MODE != 1) red_add_v4(acc_new + p, ln.L_new.x * wt, ln.L_new.y * wt, ln.L_new.z * wt, wt);
        });
    }
    counter_flush(C, 6, sk.n_paths);
}

template <bool BACKWARD>
__global__ void __launch_bounds__(kBlock) k_vps_raygen(const __grid_constant__ SceneK S, uint32_t seed, uint32_t spp, uint32_t spp_begin,
                                                       uint32_t ns, uint64_t n_local, float4 *accum, PathQueue Q, Counters *C,
                                                       const float *grad_img, const float *film_w) {
    uint64_t j = (uint64_t) blockIdx.x * kBlock + threadIdx.x;
    VpLane ln; ln.ev = EV_INIT; ln.flags = 0;
    VpSink sk; sk.accum = accum; sk.D.gsig = nullptr; sk.D.galb = nullptr; sk.grad_img = nullptr; sk.film_w = nullptr;
    sk.arena = nullptr; sk.vcap = 0; sk.flat = 0; sk.drt_mis = 0; sk.coop = 0; sk.coop_n = 0;
    sk.seed = seed; sk.spp = spp; sk.alt_seed = 0;
    sk.cnt.coll = sk.cnt.scat = sk.cnt.g_sig = sk.cnt.g_alb = 0; sk.n_paths = 0;
    if (j < n_local) vp_raygen<BACKWARD>(S, ln, sk, global_lane(j, spp, spp_begin, ns));
    bool enter = (j < n_local) && ln.ev == EV_COLL;
    unsigned long long pos = queue_slot(Q.count, enter);
    if (enter) {
        vp_store(ln, Q, pos);
        if (BACKWARD) {
            // dL = d(develop o put)/dL * grad_in of this sample (volpathsimple.py:62-70 via RBIntegrator.render_backward):
            // sum over the reconstruction filter's footprint of w * grad_in[p] / W[p]; the tracer reads it with the record
            float fx, fy;
            film_position(S, seed, spp, ln.idx, fx, fy);
            V3 dL = film_gather(S, grad_img, film_w, fx, fy);
            Q.r3[pos] = make_float4(dL.x, dL.y, dL.z, 0.f);
        }
    }
    counter_flush(C, 6, sk.n_paths);
}

template <int MODE, class GRID>
__global__ void __launch_bounds__(kWfBlock, RCV_TS_MINBLOCKS) k_twostates_wavefront(const __grid_constant__ SceneK S,
                                                                                 const __grid_constant__ GRID G, float4 *acc_old,
                                                                                 float4 *acc_new, PathQueue Q, Counters *C,
                                                                                 unsigned long long *work) {
    const float mu = combined_majorant<MODE>(G);
    const float inv_mu = 1.f / mu;
    const unsigned long long n_total = *Q.count;
    TsLane ln;
    ln.ev = EV_INIT;
    TsSink sk; sk.acc_old = acc_old; sk.acc_new = acc_new; sk.n_coll = sk.n_scat = sk.n_paths = 0;
    ChunkCursor cc; cc.next = 0; cc.end = 0; cc.more = true;
    while (true) {
        const int n_coll = __popc(__ballot_sync(0xffffffffu, ln.ev == EV_COLL));
        const int n_scat = __popc(__ballot_sync(0xffffffffu, ln.ev == EV_SCAT));
        // lanes waiting for the slow step: parked work, or empty lanes while records remain
        const int n_slow = __popc(__ballot_sync(0xffffffffu, ln.ev == EV_SLOW || (ln.ev == EV_INIT && cc.more)));
        if (n_coll == 0 && n_scat == 0 && n_slow == 0) break;
        const int sel = pick_event(n_coll, n_scat, n_slow);
        if (sel == EV_COLL) {
            if (ln.ev == EV_COLL) ts_collide<MODE>(S, G, mu, inv_mu, ln, sk);
        } else if (sel == EV_SCAT) {
            if (ln.ev == EV_SCAT) ts_scatter<MODE>(S, G, ln, sk);
        } else {
            if (ln.ev == EV_SLOW) ts_slow<MODE>(S, ln, sk);
            // refill finished lanes from the record queue
            unsigned m_init = __ballot_sync(0xffffffffu, ln.ev == EV_INIT);
            if (m_init && cc.more) {
                unsigned long long pos = take_positions(cc, work, n_total, m_init);
                if (pos != ~0ull) ts_load(ln, Q, pos);
            }
        }
        __syncwarp();
    }
    counter_flush(C, MODE == 0 ? 0 : 2, sk.n_coll);
    counter_flush(C, MODE == 0 ? 1 : 3, sk.n_scat);
    counter_flush(C, 6, sk.n_paths);
}

template <bool BACKWARD, class GRID>
__global__ void __launch_bounds__(kWfBlock, RCV_VP_MINBLOCKS) k_vps_wavefront(const __grid_constant__ SceneK S,
                                                                           const __grid_constant__ GRID G, GradK D, uint32_t seed,
                                                                           uint32_t spp, uint32_t alt_seed, const float *grad_img,
                                                                           const float *film_w, float4 *accum, PathQueue Q,
                                                                           Counters *C, unsigned long long *work, float4 *arena,
                                                                           uint32_t vcap) {
    const float mu = __ldg(G.majorants + 1);
    const float inv_mu = 1.f / mu;
    const unsigned long long n_total = *Q.count;
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;       // arena slot of this thread's lane
    VpLane ln;
    ln.ev = EV_INIT; ln.flags = 0;
    VpSink sk;
    sk.accum = accum; sk.D = D; sk.grad_img = grad_img; sk.film_w = film_w; sk.arena = arena; sk.vcap = vcap;
    sk.seed = seed; sk.spp = spp; sk.alt_seed = alt_seed; sk.flat = 0; sk.drt_mis = 0; sk.coop = 0; sk.coop_n = 0;
    sk.cnt.coll = sk.cnt.scat = sk.cnt.g_sig = sk.cnt.g_alb = 0; sk.n_paths = 0;
    ChunkCursor cc; cc.next = 0; cc.end = 0; cc.more = true;
    while (true) {
        const int n_coll = __popc(__ballot_sync(0xffffffffu, ln.ev == EV_COLL));
        const int n_scat = __popc(__ballot_sync(0xffffffffu, ln.ev == EV_SCAT));
        // lanes waiting for the slow step: parked work, or empty lanes while records remain
        const int n_slow = __popc(__ballot_sync(0xffffffffu, ln.ev == EV_SLOW || (ln.ev == EV_INIT && cc.more)));
        if (n_coll == 0 && n_scat == 0 && n_slow == 0) break;
        const int sel = pick_event(n_coll, n_scat, n_slow);
        if (sel == EV_COLL) {
            if (ln.ev == EV_COLL) vp_collide<BACKWARD>(S, G, mu, inv_mu, ln, sk);
        } else if (sel == EV_SCAT) {
            if (ln.ev == EV_SCAT) vp_vertex<BACKWARD>(S, G, ln, sk);
        } else {
            if (ln.ev == EV_SLOW) vp_slow<BACKWARD>(S, G, ln, sk);
            // refill the finished lanes from the record queue
            unsigned m_new = __ballot_sync(0xffffffffu, ln.ev == EV_INIT);
            if (m_new && cc.more) {
                unsigned long long pos = take_positions(cc, work, n_total, m_new);
                if (pos != ~0ull) vp_load(ln, sk, Q, pos, slot);
            }
        }
        __syncwarp();
    }
    counter_flush(C, 2, sk.cnt.coll);
    counter_flush(C, 3, sk.cnt.scat);
    counter_flush(C, 4, sk.cnt.g_sig);
    counter_flush(C, 5, sk.cnt.g_alb);
    counter_flush(C, 6, sk.n_paths);
}


// ===========================================================================
// POOL schedule: warp-private pools of 64 lane states in shared memory (rcv_pool.cuh)
// ===========================================================================
#ifndef RCV_POOL_BLOCK
#define RCV_POOL_BLOCK 128
#endif
// CTAs per SM.  Shared memory allows 4 pools of 4 warps; the adjoint's speculative collision step (two bricks in flight
// per lane) needs ~165 registers, so its kernels run 3 (r2c: 4 CTAs at 128 registers spill, 18.6 ms against 15.3 ms).
#ifndef RCV_POOL_MINBLOCKS
#define RCV_POOL_MINBLOCKS (RCV_SPEC_COLLIDE > 1 ? 3 : 4)
#endif
#ifndef RCV_TS_POOL_MINBLOCKS
#define RCV_TS_POOL_MINBLOCKS 4
#endif
#ifndef RCV_VP_POOL_BLOCK
#define RCV_VP_POOL_BLOCK RCV_POOL_BLOCK
#endif
// r2x (profiles/README.md): C1 2.98 ms against 3.12 ms, C2 unchanged — the cost of a tail step is the dependent chain
// inside the collision handler, not the pool's scheduling.  Off.
#ifndef RCV_STICKY_ROUNDS
#define RCV_STICKY_ROUNDS 0
#endif
#ifndef RCV_STICKY_MAX
#define RCV_STICKY_MAX 8
#endif
static constexpr int kStickyMax = RCV_STICKY_MAX;          // collision steps a sticky round takes at most before it reschedules
static constexpr int kPoolBlock = RCV_POOL_BLOCK;          // two-state kernels
static constexpr int kVpPoolBlock = RCV_VP_POOL_BLOCK;     // volpathsimple kernels

template <int WORDS>
__host__ __device__ constexpr size_t pool_warp_floats() { return (size_t) WORDS * kPoolSlots + 2 * kPoolSlots / 4 + (kPoolSlots + 8) / 2; }   // + cls, list (bytes), prefix sums of the flattened pass-2 rounds (16 bit)

// the class of every slot of the warp's pool as ballot masks (slots 32 j .. 32 j + 31 in word j)
struct PoolMasks {
    unsigned coll[kPoolMaskWords], scat[kPoolMaskWords], slow[kPoolMaskWords];
    int n_coll, n_scat, n_slow;
};
__device__ __forceinline__ void pool_masks(const unsigned char *cls, bool more, PoolMasks &M) {
    const unsigned lane = threadIdx.x & 31;
    M.n_coll = M.n_scat = M.n_slow = 0;
#pragma unroll
    for (int j = 0; j < kPoolMaskWords; ++j) {
        const int c = cls[lane + 32 * j];
        M.coll[j] = __ballot_sync(0xffffffffu, c == EV_COLL);
        M.scat[j] = __ballot_sync(0xffffffffu, c == EV_SCAT);
        // lanes waiting for the slow step: parked work, or empty slots while records remain
        M.slow[j] = __ballot_sync(0xffffffffu, c == EV_SLOW || (c == EV_INIT && more));
        M.n_coll += __popc(M.coll[j]); M.n_scat += __popc(M.scat[j]); M.n_slow += __popc(M.slow[j]);
    }
}

template <int MODE, class GRID>
__global__ void __launch_bounds__(kPoolBlock, RCV_TS_POOL_MINBLOCKS) k_twostates_pool(const __grid_constant__ SceneK S,
                                                                                   const __grid_constant__ GRID G, float4 *acc_old,
                                                                                   float4 *acc_new, PathQueue Q, Counters *C,
                                                                                   unsigned long long *work) {
    extern __shared__ float smem[];
    const unsigned lane = threadIdx.x & 31;
    float *pw = smem + (threadIdx.x >> 5) * pool_warp_floats<kTsWords>();
    unsigned char *cls = (unsigned char *) (pw + (size_t) kTsWords * kPoolSlots);
    unsigned char *list = cls + kPoolSlots;
    const float mu = combined_majorant<MODE>(G);
    const float inv_mu = 1.f / mu;
    const unsigned long long n_total = *Q.count;
    TsSink sk; sk.acc_old = acc_old; sk.acc_new = acc_new; sk.n_coll = sk.n_scat = sk.n_paths = 0;
    ChunkCursor cc; cc.next = 0; cc.end = 0; cc.more = true;
#pragma unroll
    for (int j = 0; j < kPoolMaskWords; ++j) cls[lane + 32 * j] = EV_INIT;
    __syncwarp();
    while (true) {
        PoolMasks M;
        pool_masks(cls, cc.more, M);
        if (M.n_coll == 0 && M.n_scat == 0 && M.n_slow == 0) break;
        const int sel = pool_pick(M.n_coll, M.n_scat, M.n_slow);
        if (sel == EV_COLL) {
            const int n = min(pool_list(M.coll, list), 32);
            const int s = pool_slot(list, n, 0);
            // Sticky round: when nothing else in this pool could run (no vertex, no slow step, no record to take) the same
            // lanes go straight into their next tentative collision from registers, for as long as all of them still want
            // one — the scheduling and the shared-memory round trip of a round are skipped.  This is the regime a warp
            // spends most of a small launch in (C1: one long lane per warp after the first 0.1 ms).
            const bool sticky = RCV_STICKY_ROUNDS && M.n_scat == 0 && M.n_slow == 0 && M.n_coll <= 32;
            TsLane ln;
            ln.ev = EV_INIT;
            if (s >= 0) { ts_pool_load_hot(pw, s, ln); ln.ev = EV_COLL; }
            int left = kStickyMax;
            bool again;
            do {
                if (s >= 0) ts_collide<MODE>(S, G, mu, inv_mu, ln, sk);
                again = sticky && (--left > 0) && __all_sync(0xffffffffu, s < 0 || ln.ev == EV_COLL);
            } while (again);
            if (s >= 0) {
                ts_pool_store_coll(pw, s, ln);
                cls[s] = (unsigned char) ln.ev;
            }
        } else if (sel == EV_SCAT) {
            const int s = pool_slot(list, min(pool_list(M.scat, list), 32), 0);
            if (s >= 0) {
                TsLane ln;
                ts_pool_load_hot(pw, s, ln); ts_pool_load_cold(pw, s, ln);
                ln.ev = EV_SCAT;
                ts_scatter<MODE>(S, G, ln, sk);
                ts_pool_store_all(pw, s, ln);
                cls[s] = (unsigned char) ln.ev;
            }
        } else {
            const int s = pool_slot(list, min(pool_list(M.slow, list), 32), 0);
            TsLane ln;
            ln.ev = EV_INIT;
            bool loaded = false;
            if (s >= 0 && cls[s] == EV_SLOW) {
                ts_pool_load_hot(pw, s, ln); ts_pool_load_cold(pw, s, ln);
                ln.ev = EV_SLOW;
                ts_slow<MODE>(S, ln, sk);
                loaded = true;
            }
            // refill the slots that are (now) empty from the record queue
            unsigned m_need = __ballot_sync(0xffffffffu, s >= 0 && ln.ev == EV_INIT);
            if (m_need && cc.more) {
                unsigned long long pos = take_positions(cc, work, n_total, m_need);
                if (pos != ~0ull) { ts_load(ln, Q, pos); loaded = true; }
            }
            if (s >= 0) {
                if (loaded && ln.ev != EV_INIT) ts_pool_store_all(pw, s, ln);
                cls[s] = (unsigned char) ln.ev;
            }
        }
        __syncwarp();
    }
    counter_flush(C, MODE == 0 ? 0 : 2, sk.n_coll);
    counter_flush(C, MODE == 0 ? 1 : 3, sk.n_scat);
    counter_flush(C, 6, sk.n_paths);
}

// DRT stages 1 and 2 (rcv_volpathsimple_wf.cuh): what the scheduler does with a lane that just completed (VF_JOB).
// Called by the whole warp; `done` is false for threads without such a lane.
template <int DRT, class GRID>
__device__ __forceinline__ void pool_drt_complete(const SceneK &S, const GRID &G, VpLane &ln, VpSink &sk, bool done) {
    if (DRT == 1) {
        float4 j0, j1, j2, j3;
        const bool want = done && vp_drt_job(ln, sk, j0, j1, j2, j3);
        const unsigned long long pos = queue_slot(sk.jobs.count, want);
        if (want) { sk.jobs.r0[pos] = j0; sk.jobs.r1[pos] = j1; sk.jobs.r2[pos] = j2; sk.jobs.r3[pos] = j3; }
    }
    if (DRT == 2) {
        if (done) {
            float4 a = sk.jobs.r0[ln.pix], b = sk.jobs.r1[ln.pix];
            vp_drt_scatter(S, G, sk.D, mk3(a.x, a.y, a.z), mk3(a.w, b.x, b.y), ln.result, sk.cnt);
        }
    }
    if (done) ln.flags = 0;
}

// RCV_COOP_NEE_RECORDS = 1: the recorded collisions of the shadow rays that ended in this slow step are scattered by
// the whole warp.  A lane walking its own list kept 7.6 of 32 threads busy over a fifth of the adjoint's warp
// instructions (r2w: the lists differ in length, and a slow step holds 15 lanes on average); here the lists of the
// step are laid end to end and every thread takes each 32nd record, whoever's it is.  Same values per record; only
// the order of the (unordered anyway) reductions into the gradient bricks changes.
#ifndef RCV_COOP_NEE_RECORDS
#define RCV_COOP_NEE_RECORDS 1
#endif
__device__ __forceinline__ void vp_nee_records_warp(const SceneK &S, const GridK &G, VpSink &sk) {
    const unsigned full = 0xffffffffu;
    const int lane = (int) (threadIdx.x & 31);
    const int n = sk.coop_n;
    sk.coop_n = 0;
    if (!__any_sync(full, n > 0)) return;
    int x = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(full, x, o); if (lane >= o) x += y; }
    const int total = __shfl_sync(full, x, 31);
    const int excl = x - n;
    const unsigned long long rec_bits = (unsigned long long) sk.coop_rec;
    for (int base = 0; base < total; base += 32) {
        const int j = base + lane;
        const int own = concat_owner(j, [&](int t) { return __shfl_sync(full, excl, t); });   // every thread takes part
        const int first = __shfl_sync(full, excl, own);
        const float4 *rec = (const float4 *) __shfl_sync(full, rec_bits, own);
        const V3 adj = mk3(__shfl_sync(full, sk.coop_adj.x, own), __shfl_sync(full, sk.coop_adj.y, own),
                           __shfl_sync(full, sk.coop_adj.z, own));
        if (j < total) vp_nee_record_scatter(S, G, sk, adj, rec[j - first]);
    }
}

template <bool BACKWARD, class GRID, int DRT = 0>
__global__ void __launch_bounds__(kVpPoolBlock, RCV_POOL_MINBLOCKS) k_vps_pool(const __grid_constant__ SceneK S,
                                                                             const __grid_constant__ GRID G, GradK D, uint32_t seed,
                                                                             uint32_t spp, uint32_t alt_seed, const float *grad_img,
                                                                             const float *film_w, float4 *accum, PathQueue Q,
                                                                             Counters *C, unsigned long long *work, float4 *arena,
                                                                             uint32_t vcap, PathQueue jobs, int drt_mis) {
    extern __shared__ float smem[];
    const unsigned lane = threadIdx.x & 31;
    float *pw = smem + (threadIdx.x >> 5) * pool_warp_floats<kVpWords>();
    unsigned char *cls = (unsigned char *) (pw + (size_t) kVpWords * kPoolSlots);
    unsigned char *list = cls + kPoolSlots;
    const float mu = __ldg(G.majorants + 1);
    const float inv_mu = 1.f / mu;
    const unsigned long long n_total = *Q.count;
    VpSink sk;
    sk.accum = accum; sk.D = D; sk.grad_img = grad_img; sk.film_w = film_w; sk.arena = arena; sk.vcap = vcap;
    sk.seed = seed; sk.spp = spp; sk.alt_seed = alt_seed;
    sk.cnt.coll = sk.cnt.scat = sk.cnt.g_sig = sk.cnt.g_alb = 0; sk.n_paths = 0;
    sk.jobs = jobs; sk.drt_mis = drt_mis;
    constexpr bool kFlat = RCV_FLAT_PASS2 && BACKWARD && DRT == 0;
    sk.flat = kFlat ? 1 : 0;
    constexpr bool kCoop = RCV_COOP_NEE_RECORDS && BACKWARD;
    sk.coop = kCoop ? 1 : 0; sk.coop_n = 0; sk.coop_rec = nullptr; sk.coop_adj = mk3(0, 0, 0);
    unsigned short *pre = (unsigned short *) (list + kPoolSlots);      // kPoolSlots + 1 exclusive prefix sums (flattened pass 2)
    const uint32_t slot0 = (blockIdx.x * (kVpPoolBlock / 32) + (threadIdx.x >> 5)) * kPoolSlots;   // first arena slot of this warp's pool
    ChunkCursor cc; cc.next = 0; cc.end = 0; cc.more = true;
#pragma unroll
    for (int j = 0; j < kPoolMaskWords; ++j) cls[lane + 32 * j] = EV_INIT;
    __syncwarp();
    while (true) {
        PoolMasks M;
        pool_masks(cls, cc.more, M);
        if (kFlat) {
            // Flattened pass 2: the vertices of the lanes that finished pass 1 (class EV_P2) are independent of each other
            // once the running result and the secondary sampler's position are rebuilt per vertex (vp_vertex2_flat), so a
            // round gives each thread ONE vertex of whichever waiting lane — 32 at a time, all in the same code — instead of
            // walking each lane's list with the few lanes that happen to wait for a vertex step.
            static_assert(kPoolMaskWords == 2, "two slots per thread");
            int c[2], tot[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int sl = (int) lane + 32 * j;
                c[j] = (cls[sl] == EV_P2) ? (int) pu_f(pw[48 * kPoolSlots + sl]) - (int) pu_f(pw[49 * kPoolSlots + sl]) : 0;   // nvert - vk
                int x = c[j];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if ((int) lane >= o) x += y; }
                tot[j] = __shfl_sync(0xffffffffu, x, 31);
                c[j] = x - c[j];                                       // exclusive prefix within this half
            }
            const int total = tot[0] + tot[1];
            const bool idle = (M.n_coll == 0 && M.n_scat == 0 && M.n_slow == 0);
            if (total == 0 && idle) break;
            if (total >= 32 || (total > 0 && idle)) {
                pre[lane] = (unsigned short) c[0]; pre[lane + 32] = (unsigned short) (tot[0] + c[1]);
                if (lane == 0) pre[kPoolSlots] = (unsigned short) total;
                __syncwarp();
                const int nb = min(total, 32), g = (int) lane;
                int s = -1, k0 = 0, k = 0;
                V3 res = mk3(0, 0, 0);
                if (g < nb) {
                    int lo = 0, hi = kPoolSlots;                       // the slot with pre[s] <= g < pre[s + 1]
                    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if ((int) pre[mid] <= g) lo = mid; else hi = mid; }
                    s = lo;
                    k0 = (int) pu_f(pw[49 * kPoolSlots + s]);
                    k = k0 + (g - (int) pre[s]);
                    const V3 res0 = mk3(pw[28 * kPoolSlots + s], pw[29 * kPoolSlots + s], pw[30 * kPoolSlots + s]);
                    const V3 dL = mk3(pw[31 * kPoolSlots + s], pw[32 * kPoolSlots + s], pw[33 * kPoolSlots + s]);
                    const uint32_t idx = pu_f(pw[2 * kPoolSlots + s]);
                    res = vp_vertex2_flat(S, G, sk, (slot0 + (uint32_t) s) * (vcap + kNeeCap / 4), k0, k, res0, dL, idx);
                }
                __syncwarp();
                // the thread that took the last vertex of a slot in this round writes the lane's progress back
                if (g < nb && (g == nb - 1 || g + 1 == (int) pre[s + 1])) {
                    const int nvert = (int) pu_f(pw[48 * kPoolSlots + s]);
                    pw[28 * kPoolSlots + s] = res.x; pw[29 * kPoolSlots + s] = res.y; pw[30 * kPoolSlots + s] = res.z;
                    pw[49 * kPoolSlots + s] = pf_u((uint32_t) (k + 1));
                    if (k + 1 == nvert) { cls[s] = EV_INIT; sk.n_paths++; }
                }
                __syncwarp();
                continue;
            }
        } else if (M.n_coll == 0 && M.n_scat == 0 && M.n_slow == 0) break;
        const int sel = pool_pick(M.n_coll, M.n_scat, M.n_slow);
        VpLane ln;
        if (sel == EV_COLL) {
            const int n = min(pool_list(M.coll, list), 32);
            const int s = pool_slot(list, n, 0);
            const bool sticky = RCV_STICKY_ROUNDS && M.n_scat == 0 && M.n_slow == 0 && M.n_coll <= 32;   // see k_twostates_pool
            ln.ev = EV_INIT;
            if (s >= 0) {
                vp_pool_load_hot(pw, s, ln);
                ln.ev = EV_COLL;
                ln.vbase = (slot0 + (uint32_t) s) * (vcap + kNeeCap / 4);    // a cold word; the collision records live behind it
            }
            int left = kStickyMax;
            bool again;
            do {
                if (s >= 0) {
#if RCV_SPEC_COLLIDE > 1
                    vp_collide_spec<BACKWARD>(S, G, mu, inv_mu, ln, sk);
#else
                    vp_collide<BACKWARD>(S, G, mu, inv_mu, ln, sk);
#endif
                }
                again = sticky && (--left > 0) && __all_sync(0xffffffffu, s < 0 || ln.ev == EV_COLL);
            } while (again);
            if (s >= 0) {
                vp_pool_store_coll(pw, s, ln);
                cls[s] = (unsigned char) ln.ev;
            }
        } else if (sel == EV_SCAT) {
            const int s = pool_slot(list, min(pool_list(M.scat, list), 32), 0);
            if (s >= 0) {
                vp_pool_load_hot(pw, s, ln); vp_pool_load_cold(pw, s, ln);
                ln.ev = EV_SCAT;
                vp_vertex<BACKWARD, DRT>(S, G, ln, sk);
                vp_pool_store_all(pw, s, ln);
                cls[s] = (unsigned char) ln.ev;
            }
            if (DRT != 0) pool_drt_complete<DRT>(S, G, ln, sk, s >= 0 && ln.ev == EV_INIT && (ln.flags & VF_JOB));
        } else {
            const int s = pool_slot(list, min(pool_list(M.slow, list), 32), 0);
            ln.ev = EV_INIT; ln.flags = 0;
            bool loaded = false;
            if (s >= 0 && cls[s] == EV_SLOW) {
                vp_pool_load_hot(pw, s, ln); vp_pool_load_cold(pw, s, ln);
                loaded = true;
                ln.ev = EV_SLOW;
                vp_slow<BACKWARD, DRT>(S, G, ln, sk);
            }
            if (DRT != 0) pool_drt_complete<DRT>(S, G, ln, sk, s >= 0 && ln.ev == EV_INIT && (ln.flags & VF_JOB));
            // empty slots take the next record
            unsigned m_need = __ballot_sync(0xffffffffu, s >= 0 && ln.ev == EV_INIT);
            if (m_need && cc.more) {
                unsigned long long pos = take_positions(cc, work, n_total, m_need);
                if (pos != ~0ull) { vp_load<DRT>(ln, sk, Q, pos, slot0 + (uint32_t) s); loaded = true; }
            }
            if (s >= 0) {
                if (loaded && ln.ev != EV_INIT) vp_pool_store_all(pw, s, ln);
                cls[s] = (unsigned char) ln.ev;
            }
            if (kCoop) vp_nee_records_warp(S, G, sk);
        }
        __syncwarp();
    }
    counter_flush(C, 2, sk.cnt.coll);
    counter_flush(C, 3, sk.cnt.scat);
    counter_flush(C, 4, sk.cnt.g_sig);
    counter_flush(C, 5, sk.cnt.g_alb);
    counter_flush(C, 6, sk.n_paths);
}

// DRT, between the stages: one thread per job record (vp_drt_sample).  Estimates that are complete are scattered
// here; the others become path records of the recursive radiance estimate (stage 2) in Q, with what to scatter at
// their end left in the job's own slot.
template <class GRID>
__global__ void __launch_bounds__(kBlock) k_drt_sample(const __grid_constant__ SceneK S, const __grid_constant__ GRID G, GradK D,
                                                       PathQueue JQ, PathQueue Q, int use_mis, Counters *C) {
    VpsCounts cnt; cnt.coll = cnt.scat = cnt.g_sig = cnt.g_alb = 0;
    const unsigned long long n = *JQ.count;
    const unsigned long long stride = (unsigned long long) gridDim.x * kBlock;
    for (unsigned long long base = (unsigned long long) blockIdx.x * kBlock; base < n; base += stride) {
        const unsigned long long i = base + threadIdx.x;
        int r = 0;
        float4 rec[4];
        V3 p, c, Li;
        if (i < n) {
            float4 job[4] = {__ldg(JQ.r0 + i), __ldg(JQ.r1 + i), __ldg(JQ.r2 + i), __ldg(JQ.r3 + i)};
            r = vp_drt_sample(S, G, D, job, use_mis != 0, rec, p, c, Li, cnt);
            if (r == 1) vp_drt_scatter(S, G, D, p, c, Li, cnt);
        }
        const unsigned long long pos = queue_slot(Q.count, r == 2);
        if (r == 2) {
            rec[0].w = __uint_as_float((uint32_t) i);
            Q.r0[pos] = rec[0]; Q.r1[pos] = rec[1]; Q.r2[pos] = rec[2]; Q.r3[pos] = rec[3];
            JQ.r0[i] = make_float4(p.x, p.y, p.z, c.x);
            JQ.r1[i] = make_float4(c.y, c.z, 0.f, 0.f);
        }
    }
    counter_flush(C, 2, cnt.coll);
    counter_flush(C, 4, cnt.g_sig);
    counter_flush(C, 5, cnt.g_alb);
}

// launch a pool kernel: dynamic shared memory opt-in, persistent grid
template <typename K, typename... Args>
static cudaError_t launch_pool(const LaunchCfg &cfg, K kernel, int block, size_t smem_bytes, uint64_t n_items, Args... args) {
    const int warps = block / 32;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_bytes);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem_bytes);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    uint64_t want = (n_items + (uint64_t) warps * kPoolSlots - 1) / ((uint64_t) warps * kPoolSlots);
    uint64_t b = (uint64_t) per_sm * cfg.sm_count;
    if (b > want) b = want;
    if (b < 1) b = 1;
    kernel<<<(unsigned) b, block, smem_bytes, cfg.stream>>>(args...);
    return cudaGetLastError();
}

// ===========================================================================
// launchers
// ===========================================================================
template <int MODE, class GRID>
static cudaError_t launch_twostates_mode(const LaunchCfg &cfg, const SceneK &S, const GRID &G, uint32_t seed, uint32_t spp,
                                         uint32_t spp_begin, uint32_t ns, uint64_t n_local, float4 *acc_old, float4 *acc_new) {
    cudaError_t e;
    if (cfg.tracer == 0 || cfg.tracer == 3) {
        if ((e = cudaMemsetAsync(cfg.work_counter, 0, sizeof(unsigned long long), cfg.stream)) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(cfg.queue.count, 0, sizeof(unsigned long long), cfg.stream)) != cudaSuccess) return e;
        k_twostates_raygen<MODE><<<(unsigned) ((n_local + kBlock - 1) / kBlock), kBlock, 0, cfg.stream>>>(
            S, seed, spp, spp_begin, ns, n_local, acc_old, acc_new, cfg.queue, cfg.counters);
        if (cfg.tracer == 3) {
            e = launch_pool(cfg, k_twostates_pool<MODE, GRID>, kPoolBlock, (kPoolBlock / 32) * pool_warp_floats<kTsWords>() * sizeof(float), n_local, S, G,
                            acc_old, acc_new, cfg.queue, cfg.counters, cfg.work_counter);
            if (e != cudaSuccess) return e;
        } else {
            unsigned blocks;
            if ((e = persistent_grid(cfg, k_twostates_wavefront<MODE, GRID>, kWfBlock, n_local, blocks)) != cudaSuccess) return e;
            k_twostates_wavefront<MODE, GRID><<<blocks, kWfBlock, 0, cfg.stream>>>(S, G, acc_old, acc_new, cfg.queue, cfg.counters,
                                                                             cfg.work_counter);
        }
        (*cfg.launches) += 2;
    } else {
        int persistent = cfg.tracer == 2;
        unsigned blocks = (unsigned) ((n_local + kBlock - 1) / kBlock);
        if (persistent) {
            if ((e = cudaMemsetAsync(cfg.work_counter, 0, sizeof(unsigned long long), cfg.stream)) != cudaSuccess) return e;
            if ((e = persistent_grid(cfg, k_twostates_simple<MODE, GRID>, kBlock, n_local, blocks)) != cudaSuccess) return e;
        }
        k_twostates_simple<MODE, GRID><<<blocks, kBlock, 0, cfg.stream>>>(S, G, seed, spp, spp_begin, ns, n_local, acc_old, acc_new,
                                                                    cfg.counters, cfg.work_counter, persistent);
        (*cfg.launches)++;
    }
    return cudaGetLastError();
}

template <class GRID>
static cudaError_t launch_twostates_grid(const LaunchCfg &cfg, const SceneK &S, const GRID &G, int mode, uint32_t seed, uint32_t spp,
                                         uint32_t spp_begin, uint32_t ns, uint64_t n_local, float4 *acc_old, float4 *acc_new) {
    switch (mode) {
        case 0: return launch_twostates_mode<0>(cfg, S, G, seed, spp, spp_begin, ns, n_local, acc_old, acc_new);
        case 1: return launch_twostates_mode<1>(cfg, S, G, seed, spp, spp_begin, ns, n_local, acc_old, acc_new);
        default: return launch_twostates_mode<2>(cfg, S, G, seed, spp, spp_begin, ns, n_local, acc_old, acc_new);
    }
}
// G.mgrid != nullptr selects the kernels that sample free flights through the majorant supergrid
cudaError_t launch_twostates(const LaunchCfg &cfg, const SceneK &S, const GridKS &G, int mode, uint32_t seed, uint32_t spp,
                             uint32_t spp_begin, uint32_t spp_end, float4 *acc_old, float4 *acc_new) {
    uint32_t ns = spp_end - spp_begin;
    uint64_t n_local = (uint64_t) S.W * S.H * ns;
    if (n_local == 0) return cudaSuccess;
    if (G.mgrid) return launch_twostates_grid<GridKS>(cfg, S, G, mode, seed, spp, spp_begin, ns, n_local, acc_old, acc_new);
    return launch_twostates_grid<GridK>(cfg, S, G, mode, seed, spp, spp_begin, ns, n_local, acc_old, acc_new);
}

template <bool BACKWARD, class GRID>
static cudaError_t launch_vps(const LaunchCfg &cfg, const SceneK &S, const GRID &G, const GradK &D, uint32_t seed, uint32_t spp,
                              uint32_t spp_begin, uint32_t spp_end, uint32_t alt_seed, const float *grad_img, const float *film_w,
                              float4 *accum) {
    uint32_t ns = spp_end - spp_begin;
    uint64_t n_local = (uint64_t) S.W * S.H * ns;
    if (n_local == 0) return cudaSuccess;
    cudaError_t e;
    if (BACKWARD && (cfg.drt_flags & 3) == 3 && cfg.tracer == 3) {
        // use_drt with sub-sampling on the pool machines: adjoint + reservoir -> jobs -> recursive radiance estimate
        const size_t smem = (kVpPoolBlock / 32) * pool_warp_floats<kVpWords>() * sizeof(float);
        const int mis = (cfg.drt_flags & 4) ? 1 : 0;
        if ((e = cudaMemsetAsync(cfg.work_counter, 0, sizeof(unsigned long long), cfg.stream)) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(cfg.queue.count, 0, sizeof(unsigned long long), cfg.stream)) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(cfg.jobs.count, 0, sizeof(unsigned long long), cfg.stream)) != cudaSuccess) return e;
        k_vps_raygen<true><<<(unsigned) ((n_local + kBlock - 1) / kBlock), kBlock, 0, cfg.stream>>>(
            S, seed, spp, spp_begin, ns, n_local, accum, cfg.queue, cfg.counters, grad_img, film_w);
        e = launch_pool(cfg, k_vps_pool<true, GRID, 1>, kVpPoolBlock, smem, n_local, S, G, D, seed, spp, alt_seed, grad_img, film_w, accum,
                        cfg.queue, cfg.counters, cfg.work_counter, cfg.arena, cfg.vcap, cfg.jobs, mis);
        if (e != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(cfg.queue.count, 0, sizeof(unsigned long long), cfg.stream)) != cudaSuccess) return e;
        unsigned blocks;
        if ((e = persistent_grid(cfg, k_drt_sample<GRID>, kBlock, n_local, blocks)) != cudaSuccess) return e;
        k_drt_sample<GRID><<<blocks, kBlock, 0, cfg.stream>>>(S, G, D, cfg.jobs, cfg.queue, mis, cfg.counters);
        if ((e = cudaMemsetAsync(cfg.work_counter, 0, sizeof(unsigned long long), cfg.stream)) != cudaSuccess) return e;
        e = launch_pool(cfg, k_vps_pool<false, GRID, 2>, kVpPoolBlock, smem, n_local, S, G, D, seed, spp, alt_seed, grad_img, film_w, accum,
                        cfg.queue, cfg.counters, cfg.work_counter, cfg.arena, cfg.vcap, cfg.jobs, mis);
        if (e != cudaSuccess) return e;
        (*cfg.launches) += 4;
        return cudaGetLastError();
    }
    if (BACKWARD && (cfg.drt_flags & 1)) {
        unsigned blocks;
        if ((e = cudaMemsetAsync(cfg.work_counter, 0, sizeof(unsigned long long), cfg.stream)) != cudaSuccess) return e;
        if ((e = persistent_grid(cfg, k_vps_drt<GRID>, kDrtBlock, n_local, blocks)) != cudaSuccess) return e;
        k_vps_drt<GRID><<<blocks, kDrtBlock, 0, cfg.stream>>>(S, G, D, seed, spp, spp_begin, ns, n_local, alt_seed, grad_img, film_w,
                                                           cfg.drt_flags, cfg.counters, cfg.work_counter);
        (*cfg.launches)++;
        return cudaGetLastError();
    }
    if (cfg.tracer == 0 || cfg.tracer == 3) {
        if ((e = cudaMemsetAsync(cfg.work_counter, 0, sizeof(unsigned long long), cfg.stream)) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(cfg.queue.count, 0, sizeof(unsigned long long), cfg.stream)) != cudaSuccess) return e;
        PathQueue qv = cfg.queue;
        if (!BACKWARD) qv.r3 = nullptr;              // the primal render carries no dL
        k_vps_raygen<BACKWARD><<<(unsigned) ((n_local + kBlock - 1) / kBlock), kBlock, 0, cfg.stream>>>(
            S, seed, spp, spp_begin, ns, n_local, accum, qv, cfg.counters, grad_img, film_w);
        if (cfg.tracer == 3) {
            e = launch_pool(cfg, k_vps_pool<BACKWARD, GRID>, kVpPoolBlock, (kVpPoolBlock / 32) * pool_warp_floats<kVpWords>() * sizeof(float), n_local, S, G, D, seed,
                            spp, alt_seed, grad_img, film_w, accum, qv, cfg.counters, cfg.work_counter, cfg.arena, cfg.vcap, cfg.jobs, 0);
            if (e != cudaSuccess) return e;
        } else {
            unsigned blocks;
            if ((e = persistent_grid(cfg, k_vps_wavefront<BACKWARD, GRID>, kWfBlock, n_local, blocks)) != cudaSuccess) return e;
            k_vps_wavefront<BACKWARD, GRID><<<blocks, kWfBlock, 0, cfg.stream>>>(S, G, D, seed, spp, alt_seed, grad_img, film_w, accum,
                                                                           qv, cfg.counters, cfg.work_counter, cfg.arena,
                                                                           cfg.vcap);
        }
        (*cfg.launches) += 2;
    } else {
        int persistent = cfg.tracer == 2;
        unsigned blocks = (unsigned) ((n_local + kBlock - 1) / kBlock);
        if (persistent) {
            if ((e = cudaMemsetAsync(cfg.work_counter, 0, sizeof(unsigned long long), cfg.stream)) != cudaSuccess) return e;
            if ((e = persistent_grid(cfg, k_vps_simple<BACKWARD, GRID>, kBlock, n_local, blocks)) != cudaSuccess) return e;
        }
        k_vps_simple<BACKWARD, GRID><<<blocks, kBlock, 0, cfg.stream>>>(S, G, D, seed, spp, spp_begin, ns, n_local, alt_seed, grad_img, film_w,
                                                                  accum, cfg.counters, cfg.work_counter, persistent);
        (*cfg.launches)++;
    }
    return cudaGetLastError();
}

// upper bound of the arena slots (lanes in flight) the adjoint launch of this schedule can use
size_t vps_backward_slots(const LaunchCfg &cfg) {
    if (cfg.tracer == 3) {
        int per_sm = 0;
        size_t smem = (kVpPoolBlock / 32) * pool_warp_floats<kVpWords>() * sizeof(float);
        cudaFuncSetAttribute(k_vps_pool<true, GridK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_vps_pool<true, GridK>, kVpPoolBlock, smem) != cudaSuccess || per_sm < 1)
            per_sm = 1;
        if (cfg.drt_flags & 1) {         // the DRT instance of the adjoint may fit a different number of blocks
            int per_sm_drt = 0;
            cudaFuncSetAttribute(k_vps_pool<true, GridK, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_drt, k_vps_pool<true, GridK, 1>, kVpPoolBlock, smem) == cudaSuccess &&
                per_sm_drt > per_sm)
                per_sm = per_sm_drt;
        }
        return (size_t) per_sm * cfg.sm_count * (kVpPoolBlock / 32) * kPoolSlots;
    }
    if (cfg.tracer == 0) {
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_vps_wavefront<true, GridK>, kWfBlock, 0) != cudaSuccess || per_sm < 1)
            per_sm = 1;
        return (size_t) per_sm * cfg.sm_count * kWfBlock;
    }
    return 0;
}

cudaError_t launch_vps_primal(const LaunchCfg &cfg, const SceneK &S, const GridKS &G, uint32_t seed, uint32_t spp,
                              uint32_t spp_begin, uint32_t spp_end, float4 *accum) {
    GradK D; D.gsig = nullptr; D.galb = nullptr; D.gbrick = nullptr;
    if (G.mgrid) return launch_vps<false, GridKS>(cfg, S, G, D, seed, spp, spp_begin, spp_end, 0u, nullptr, nullptr, accum);
    return launch_vps<false, GridK>(cfg, S, G, D, seed, spp, spp_begin, spp_end, 0u, nullptr, nullptr, accum);
}

cudaError_t launch_vps_backward(const LaunchCfg &cfg, const SceneK &S, const GridKS &G, const GradK &D, uint32_t seed_grad,
                                uint32_t spp_grad, uint32_t spp_begin, uint32_t spp_end, uint32_t alt_seed,
                                const float *grad_img, const float *film_w) {
    if (G.mgrid)
        return launch_vps<true, GridKS>(cfg, S, G, D, seed_grad, spp_grad, spp_begin, spp_end, alt_seed, grad_img, film_w, nullptr);
    return launch_vps<true, GridK>(cfg, S, G, D, seed_grad, spp_grad, spp_begin, spp_end, alt_seed, grad_img, film_w, nullptr);
}

// per-pixel film weight of ALL spp samples (a pure function of the first two
// draws of every lane): the denominator of hdrfilm.develop, needed by the
// adjoint to differentiate the weight division.
__global__ void __launch_bounds__(kBlock) k_film_weights(const __grid_constant__ SceneK S, uint32_t seed, uint32_t spp,
                                                         uint32_t spp_begin, uint32_t ns, uint64_t n_local, float *film_w) {
    uint64_t j = (uint64_t) blockIdx.x * kBlock + threadIdx.x;
    const bool live = j < n_local;
    float fx = 0.f, fy = 0.f;
    if (live) film_position(S, seed, spp, (uint32_t) global_lane(j, spp, spp_begin, ns), fx, fy);
    uint32_t pixel;
    bool one = warp_one_pixel(live, j, ns, &pixel);
    const int py = (int) (pixel / S.W), px = (int) (pixel - (uint32_t) py * S.W);
    float w[32];
    if (one) one = __all_sync(0xffffffffu, film_tile_weights(S, fx, fy, px - 2, py - 2, w));
    if (one) {
        const float sum = warp_reduce_scatter32(w);
        const int c = (int) (threadIdx.x & 31), x = px - 2 + c % 5, y = py - 2 + c / 5;
        if (c < 25 && sum != 0.f && x >= 0 && x < (int) S.W && y >= 0 && y < (int) S.H) red_add_f32(film_w + y * (int) S.W + x, sum);
        return;
    }
    if (live) film_for_each(S, fx, fy, [&](int p, float w) { red_add_f32(film_w + p, w); });
}

// adds the film weights of samples [spp_begin, spp_end) of every pixel to film_w (zeroed first when `zero`)
cudaError_t launch_film_weights(const LaunchCfg &cfg, const SceneK &S, uint32_t seed, uint32_t spp, uint32_t spp_begin,
                                uint32_t spp_end, float *film_w, bool zero) {
    uint32_t ns = spp_end - spp_begin;
    uint64_t n = (uint64_t) S.W * S.H * ns;
    if (zero) {
        cudaError_t e = cudaMemsetAsync(film_w, 0, (size_t) S.W * S.H * sizeof(float), cfg.stream);
        if (e != cudaSuccess) return e;
    }
    if (n == 0) return cudaSuccess;
    k_film_weights<<<(unsigned) ((n + kBlock - 1) / kBlock), kBlock, 0, cfg.stream>>>(S, seed, spp, spp_begin, ns, n, film_w);
    (*cfg.launches)++;
    return cudaGetLastError();
}

}  // namespace rcv
