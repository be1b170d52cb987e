// rcv_trace.cu — tracing kernels: two-state primal (K1), volpathsimple primal
// (K5) and the fused pass-1 + pass-2 adjoint (K5+K6).
//
// Two schedules of the same per-lane code (rcv_twostate.cuh, rcv_volpathsimple.cuh):
//   * SIMPLE     one thread per lane; a warp runs until its longest path ends.
//   * WAVEFRONT  persistent warps, one resident grid sized to the SM count;
//                every warp pulls 32-lane batches from a global work counter
//                until the lane range is exhausted (no tail of idle CTAs).
// Results are schedule-independent: each lane owns its PCG32 stream, indexed
// by the global lane id pixel*spp + k.
#include "rcv_internal.h"
#include "rcv_twostate.cuh"
#include "rcv_volpathsimple.cuh"

namespace rcv {

static constexpr int kBlock = 256;

__device__ __forceinline__ uint64_t global_lane(uint64_t j, uint32_t spp, uint32_t spp_begin, uint32_t ns) {
    uint64_t pixel = j / ns;
    uint32_t k = spp_begin + (uint32_t) (j - pixel * ns);
    return pixel * spp + k;
}

// ---------------------------------------------------------------------------
// K1: two-state primal
// ---------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ void twostates_lane(const SceneK &S, const GridK &G, uint32_t seed, uint32_t spp, uint64_t idx,
                                               float4 *acc_old, float4 *acc_new, unsigned &nc, unsigned &nsct) {
    Lane L = lane_setup(S, seed, spp, idx);
    V3 Lo, Ln;
    trace_twostates<MODE>(S, G, L.rng, L.o, L.d, Lo, Ln, nc, nsct);
    if (L.pix >= 0) {
        // ImageBlock.put with the box filter: (L * weight, weight) added at floor(pos)
        if (MODE != 2) red_add_v4(acc_old + L.pix, Lo.x, Lo.y, Lo.z, 1.f);
        if (MODE != 1) red_add_v4(acc_new + L.pix, Ln.x, Ln.y, Ln.z, 1.f);
    }
}

template <int MODE>
__global__ void __launch_bounds__(kBlock) k_twostates_simple(const __grid_constant__ SceneK S, const __grid_constant__ GridK G,
                                                             uint32_t seed, uint32_t spp, uint32_t spp_begin, uint32_t ns,
                                                             uint64_t n_local, float4 *acc_old, float4 *acc_new, Counters *C) {
    uint64_t j = (uint64_t) blockIdx.x * kBlock + threadIdx.x;
    unsigned nc = 0, nsct = 0, np = 0;
    if (j < n_local) {
        twostates_lane<MODE>(S, G, seed, spp, global_lane(j, spp, spp_begin, ns), acc_old, acc_new, nc, nsct);
        np = 1;
    }
    counter_flush(C, MODE == 0 ? 0 : 2, nc);
    counter_flush(C, MODE == 0 ? 1 : 3, nsct);
    counter_flush(C, 6, np);
}

template <int MODE>
__global__ void __launch_bounds__(kBlock) k_twostates_persistent(const __grid_constant__ SceneK S, const __grid_constant__ GridK G,
                                                                 uint32_t seed, uint32_t spp, uint32_t spp_begin, uint32_t ns,
                                                                 uint64_t n_local, float4 *acc_old, float4 *acc_new, Counters *C,
                                                                 unsigned long long *work) {
    unsigned nc = 0, nsct = 0, np = 0;
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
    counter_flush(C, MODE == 0 ? 0 : 2, nc);
    counter_flush(C, MODE == 0 ? 1 : 3, nsct);
    counter_flush(C, 6, np);
}

template <int MODE>
static cudaError_t launch_twostates_mode(const LaunchCfg &cfg, const SceneK &S, const GridK &G, uint32_t seed, uint32_t spp,
                                         uint32_t spp_begin, uint32_t ns, uint64_t n_local, float4 *acc_old, float4 *acc_new) {
    if (cfg.tracer == 1) {
        uint64_t blocks = (n_local + kBlock - 1) / kBlock;
        k_twostates_simple<MODE><<<(unsigned) blocks, kBlock, 0, cfg.stream>>>(S, G, seed, spp, spp_begin, ns, n_local, acc_old,
                                                                                acc_new, cfg.counters);
    } else {
        cudaError_t e = cudaMemsetAsync(cfg.work_counter, 0, sizeof(unsigned long long), cfg.stream);
        if (e != cudaSuccess) return e;
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_twostates_persistent<MODE>, kBlock, 0);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
        uint64_t want = (n_local + kBlock - 1) / kBlock;
        uint64_t blocks = (uint64_t) per_sm * cfg.sm_count;
        if (blocks > want) blocks = want;
        k_twostates_persistent<MODE><<<(unsigned) blocks, kBlock, 0, cfg.stream>>>(S, G, seed, spp, spp_begin, ns, n_local, acc_old,
                                                                                    acc_new, cfg.counters, cfg.work_counter);
    }
    (*cfg.launches)++;
    return cudaGetLastError();
}

cudaError_t launch_twostates(const LaunchCfg &cfg, const SceneK &S, const GridK &G, int mode, uint32_t seed, uint32_t spp,
                             uint32_t spp_begin, uint32_t spp_end, float4 *acc_old, float4 *acc_new) {
    uint32_t ns = spp_end - spp_begin;
    uint64_t n_local = (uint64_t) S.W * S.H * ns;
    if (n_local == 0) return cudaSuccess;
    switch (mode) {
        case 0: return launch_twostates_mode<0>(cfg, S, G, seed, spp, spp_begin, ns, n_local, acc_old, acc_new);
        case 1: return launch_twostates_mode<1>(cfg, S, G, seed, spp, spp_begin, ns, n_local, acc_old, acc_new);
        default: return launch_twostates_mode<2>(cfg, S, G, seed, spp, spp_begin, ns, n_local, acc_old, acc_new);
    }
}

// ---------------------------------------------------------------------------
// K5 / K6: volpathsimple primal render and the fused two-pass adjoint
// ---------------------------------------------------------------------------
template <bool BACKWARD>
__device__ __forceinline__ void vps_lane(const SceneK &S, const GridK &G, const GradK &D, uint32_t seed, uint32_t spp, uint64_t idx,
                                         uint32_t alt_seed, const float *grad_img, const float *film_w, float4 *accum,
                                         VpsCounts &cnt) {
    Lane L = lane_setup(S, seed, spp, idx);
    Pcg alt; alt.state = 0; alt.inc = 1;
    V3 res;
    Pcg rng1 = L.rng;                                   // RBIntegrator.render_backward: sampler.clone() for pass 1
    vps_trace<false>(S, G, D, rng1, alt, L.o, L.d, mk3(0, 0, 0), mk3(0, 0, 0), res, cnt);
    if (!BACKWARD) {
        if (L.pix >= 0) red_add_v4(accum + L.pix, res.x, res.y, res.z, 1.f);
        return;
    }
    // dL = d(develop o put)/dL * grad_in: box filter -> grad_in[pixel] / W[pixel]
    V3 dL = mk3(0, 0, 0);
    if (L.pix >= 0) {
        float w = __ldg(film_w + L.pix);
        float dv = (w == 0.f) ? 1.f : w;
        dL = mk3(__ldg(grad_img + 3 * L.pix) / dv, __ldg(grad_img + 3 * L.pix + 1) / dv, __ldg(grad_img + 3 * L.pix + 2) / dv);
    }
    sampler_seed(alt, alt_seed, (uint32_t) idx);        // volpathsimple.py:119-120
    V3 res2;
    vps_trace<true>(S, G, D, L.rng, alt, L.o, L.d, dL, res, res2, cnt);
}

template <bool BACKWARD>
__global__ void __launch_bounds__(kBlock) k_vps(const __grid_constant__ SceneK S, const __grid_constant__ GridK G, GradK D,
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

template <bool BACKWARD>
static cudaError_t launch_vps(const LaunchCfg &cfg, const SceneK &S, const GridK &G, const GradK &D, uint32_t seed, uint32_t spp,
                              uint32_t spp_begin, uint32_t spp_end, uint32_t alt_seed, const float *grad_img, const float *film_w,
                              float4 *accum) {
    uint32_t ns = spp_end - spp_begin;
    uint64_t n_local = (uint64_t) S.W * S.H * ns;
    if (n_local == 0) return cudaSuccess;
    uint64_t want = (n_local + kBlock - 1) / kBlock;
    uint64_t blocks = want;
    int persistent = cfg.tracer != 1;
    if (persistent) {
        cudaError_t e = cudaMemsetAsync(cfg.work_counter, 0, sizeof(unsigned long long), cfg.stream);
        if (e != cudaSuccess) return e;
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_vps<BACKWARD>, kBlock, 0);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
        blocks = (uint64_t) per_sm * cfg.sm_count;
        if (blocks > want) blocks = want;
    }
    k_vps<BACKWARD><<<(unsigned) blocks, kBlock, 0, cfg.stream>>>(S, G, D, seed, spp, spp_begin, ns, n_local, alt_seed, grad_img,
                                                                   film_w, accum, cfg.counters, cfg.work_counter, persistent);
    (*cfg.launches)++;
    return cudaGetLastError();
}

cudaError_t launch_vps_primal(const LaunchCfg &cfg, const SceneK &S, const GridK &G, uint32_t seed, uint32_t spp,
                              uint32_t spp_begin, uint32_t spp_end, float4 *accum) {
    GradK D; D.grad = nullptr;
    return launch_vps<false>(cfg, S, G, D, seed, spp, spp_begin, spp_end, 0u, nullptr, nullptr, accum);
}

cudaError_t launch_vps_backward(const LaunchCfg &cfg, const SceneK &S, const GridK &G, const GradK &D, uint32_t seed_grad,
                                uint32_t spp_grad, uint32_t spp_begin, uint32_t spp_end, uint32_t alt_seed,
                                const float *grad_img, const float *film_w) {
    return launch_vps<true>(cfg, S, G, D, seed_grad, spp_grad, spp_begin, spp_end, alt_seed, grad_img, film_w, nullptr);
}

// per-pixel film weight of ALL spp samples (a pure function of the first two
// draws of every lane): the denominator of hdrfilm.develop, needed by the
// adjoint to differentiate the weight division.
__global__ void __launch_bounds__(kBlock) k_film_weights(const __grid_constant__ SceneK S, uint32_t seed, uint32_t spp,
                                                         uint64_t n_lanes, float *film_w) {
    uint64_t idx = (uint64_t) blockIdx.x * kBlock + threadIdx.x;
    if (idx >= n_lanes) return;
    Lane L = lane_setup(S, seed, spp, idx);
    if (L.pix >= 0) red_add_f32(film_w + L.pix, 1.f);
}

cudaError_t launch_film_weights(const LaunchCfg &cfg, const SceneK &S, uint32_t seed, uint32_t spp, float *film_w) {
    uint64_t n = (uint64_t) S.W * S.H * spp;
    cudaError_t e = cudaMemsetAsync(film_w, 0, (size_t) S.W * S.H * sizeof(float), cfg.stream);
    if (e != cudaSuccess) return e;
    k_film_weights<<<(unsigned) ((n + kBlock - 1) / kBlock), kBlock, 0, cfg.stream>>>(S, seed, spp, n, film_w);
    (*cfg.launches)++;
    return cudaGetLastError();
}

}  // namespace rcv
