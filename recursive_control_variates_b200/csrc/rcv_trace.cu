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
#include "rcv_twostate_wf.cuh"
#include "rcv_volpathsimple.cuh"
#include "rcv_volpathsimple_wf.cuh"

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

// ---------------------------------------------------------------------------
// Warp scheduler of the event machines (rcv_twostate_wf.cuh).
// Every warp is persistent.  Each round it counts the lanes waiting for each
// event class and runs ONE handler: the tentative-collision step whenever at
// least kCollThreshold lanes are ready for it (it contains the grid gather, the
// only expensive memory access), otherwise the larger of the scatter / init
// queues.  Finished lanes are refilled from a warp-private chunk of lane
// indices taken from the global work counter (one atomic per kChunk paths).
// ---------------------------------------------------------------------------
#ifndef RCV_WF_BLOCK
#define RCV_WF_BLOCK 256
#endif
#ifndef RCV_WF_MINBLOCKS
#define RCV_WF_MINBLOCKS 2
#endif
#ifndef RCV_COLL_THRESHOLD
#define RCV_COLL_THRESHOLD 16
#endif
static constexpr int kWfBlock = RCV_WF_BLOCK;
static constexpr int kWfMinBlocks = RCV_WF_MINBLOCKS;
static constexpr int kCollThreshold = RCV_COLL_THRESHOLD;
static constexpr unsigned long long kChunk = 512;

template <int MODE>
__global__ void __launch_bounds__(kWfBlock, kWfMinBlocks) k_twostates_wavefront(const __grid_constant__ SceneK S, const __grid_constant__ GridK G,
                                                                      uint32_t seed, uint32_t spp, uint32_t spp_begin, uint32_t ns,
                                                                      uint64_t n_local, float4 *acc_old, float4 *acc_new, Counters *C,
                                                                      unsigned long long *work) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const float mu = combined_majorant<MODE>(G);
    TsLane ln;
    ln.ev = EV_INIT;
    TsSink sk; sk.acc_old = acc_old; sk.acc_new = acc_new; sk.n_coll = sk.n_scat = sk.n_paths = 0;
    unsigned long long chunk_next = 0, chunk_end = 0;
    bool more = true;
    while (true) {
        unsigned m_init = __ballot_sync(0xffffffffu, ln.ev == EV_INIT);
        unsigned m_coll = __ballot_sync(0xffffffffu, ln.ev == EV_COLL);
        unsigned m_scat = __ballot_sync(0xffffffffu, ln.ev == EV_SCAT);
        int n_init = more ? __popc(m_init) : 0, n_coll = __popc(m_coll), n_scat = __popc(m_scat);
        if (n_coll == 0 && n_scat == 0 && n_init == 0) break;
        int sel;
        if (n_coll >= kCollThreshold || (n_scat == 0 && n_init == 0)) sel = EV_COLL;
        else if (n_scat >= n_init) sel = EV_SCAT;
        else sel = EV_INIT;
        if (sel == EV_COLL) {
            if (ln.ev == EV_COLL) ts_collide<MODE>(S, G, mu, ln, sk);
        } else if (sel == EV_SCAT) {
            if (ln.ev == EV_SCAT) ts_scatter<MODE>(S, G, ln, sk);
        } else {
            if (chunk_next >= chunk_end) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(work, kChunk);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base >= n_local) { more = false; continue; }
                chunk_next = base;
                chunk_end = (base + kChunk < n_local) ? base + kChunk : n_local;
            }
            unsigned long long avail = chunk_end - chunk_next;
            unsigned rank = __popc(m_init & lt_mask);
            if (ln.ev == EV_INIT && rank < avail)
                ts_init<MODE>(S, ln, sk, seed, spp, global_lane(chunk_next + rank, spp, spp_begin, ns));
            unsigned long long take = (unsigned long long) n_init < avail ? (unsigned long long) n_init : avail;
            chunk_next += take;
        }
        __syncwarp();
    }
    counter_flush(C, MODE == 0 ? 0 : 2, sk.n_coll);
    counter_flush(C, MODE == 0 ? 1 : 3, sk.n_scat);
    counter_flush(C, 6, sk.n_paths);
}

template <int MODE>
static cudaError_t launch_twostates_mode(const LaunchCfg &cfg, const SceneK &S, const GridK &G, uint32_t seed, uint32_t spp,
                                         uint32_t spp_begin, uint32_t ns, uint64_t n_local, float4 *acc_old, float4 *acc_new) {
    if (cfg.tracer == 1) {
        uint64_t blocks = (n_local + kBlock - 1) / kBlock;
        k_twostates_simple<MODE><<<(unsigned) blocks, kBlock, 0, cfg.stream>>>(S, G, seed, spp, spp_begin, ns, n_local, acc_old,
                                                                                acc_new, cfg.counters);
    } else if (cfg.tracer == 0) {
        cudaError_t e = cudaMemsetAsync(cfg.work_counter, 0, sizeof(unsigned long long), cfg.stream);
        if (e != cudaSuccess) return e;
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_twostates_wavefront<MODE>, kWfBlock, 0);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
        uint64_t want = (n_local + kWfBlock - 1) / kWfBlock;
        uint64_t blocks = (uint64_t) per_sm * cfg.sm_count;
        if (blocks > want) blocks = want;
        k_twostates_wavefront<MODE><<<(unsigned) blocks, kWfBlock, 0, cfg.stream>>>(S, G, seed, spp, spp_begin, ns, n_local, acc_old,
                                                                                     acc_new, cfg.counters, cfg.work_counter);
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

// wavefront schedule of the volpathsimple event machine (rcv_volpathsimple_wf.cuh);
// same warp scheduler as k_twostates_wavefront, plus the "resume as pass 2" flavour of EV_INIT
template <bool BACKWARD>
__global__ void __launch_bounds__(kWfBlock, kWfMinBlocks) k_vps_wavefront(const __grid_constant__ SceneK S, const __grid_constant__ GridK G, GradK D,
                                                                uint32_t seed, uint32_t spp, uint32_t spp_begin, uint32_t ns,
                                                                uint64_t n_local, uint32_t alt_seed, const float *grad_img,
                                                                const float *film_w, float4 *accum, Counters *C,
                                                                unsigned long long *work) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const float mu = __ldg(G.majorants + 1);
    VpLane ln;
    ln.ev = EV_INIT; ln.flags = 0;
    VpSink sk;
    sk.accum = accum; sk.D = D; sk.grad_img = grad_img; sk.film_w = film_w;
    sk.seed = seed; sk.spp = spp; sk.alt_seed = alt_seed;
    sk.cnt.coll = sk.cnt.scat = sk.cnt.g_sig = sk.cnt.g_alb = 0; sk.n_paths = 0;
    unsigned long long chunk_next = 0, chunk_end = 0;
    bool more = true;
    while (true) {
        const bool is_init = ln.ev == EV_INIT;
        const bool resume = BACKWARD && is_init && (ln.flags & VF_PASS2);
        unsigned m_new = __ballot_sync(0xffffffffu, is_init && !resume);
        unsigned m_res = __ballot_sync(0xffffffffu, resume);
        unsigned m_coll = __ballot_sync(0xffffffffu, ln.ev == EV_COLL);
        unsigned m_scat = __ballot_sync(0xffffffffu, ln.ev == EV_SCAT);
        int n_new = more ? __popc(m_new) : 0;
        int n_init = n_new + __popc(m_res), n_coll = __popc(m_coll), n_scat = __popc(m_scat);
        if (n_coll == 0 && n_scat == 0 && n_init == 0) break;
        int sel;
        if (n_coll >= kCollThreshold || (n_scat == 0 && n_init == 0)) sel = EV_COLL;
        else if (n_scat >= n_init) sel = EV_SCAT;
        else sel = EV_INIT;
        if (sel == EV_COLL) {
            if (ln.ev == EV_COLL) vp_collide<BACKWARD>(S, G, mu, ln, sk);
        } else if (sel == EV_SCAT) {
            if (ln.ev == EV_SCAT) vp_vertex<BACKWARD>(S, G, mu, ln, sk);
        } else {
            bool have = true;
            if (n_new > 0 && chunk_next >= chunk_end) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(work, kChunk);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base >= n_local) { more = false; have = false; }
                else {
                    chunk_next = base;
                    chunk_end = (base + kChunk < n_local) ? base + kChunk : n_local;
                }
            }
            if (resume) {
                vp_resume<BACKWARD>(S, ln, sk);
            } else if (is_init && have && n_new > 0) {
                unsigned long long avail = chunk_end - chunk_next;
                unsigned rank = __popc(m_new & lt_mask);
                if (rank < avail) vp_begin<BACKWARD>(S, ln, sk, global_lane(chunk_next + rank, spp, spp_begin, ns));
            }
            if (have && n_new > 0) {
                unsigned long long avail = chunk_end - chunk_next;
                chunk_next += ((unsigned long long) n_new < avail) ? (unsigned long long) n_new : avail;
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

template <bool BACKWARD>
static cudaError_t launch_vps(const LaunchCfg &cfg, const SceneK &S, const GridK &G, const GradK &D, uint32_t seed, uint32_t spp,
                              uint32_t spp_begin, uint32_t spp_end, uint32_t alt_seed, const float *grad_img, const float *film_w,
                              float4 *accum) {
    uint32_t ns = spp_end - spp_begin;
    uint64_t n_local = (uint64_t) S.W * S.H * ns;
    if (n_local == 0) return cudaSuccess;
    if (cfg.tracer == 0) {
        cudaError_t e = cudaMemsetAsync(cfg.work_counter, 0, sizeof(unsigned long long), cfg.stream);
        if (e != cudaSuccess) return e;
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_vps_wavefront<BACKWARD>, kWfBlock, 0);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
        uint64_t want = (n_local + kWfBlock - 1) / kWfBlock;
        uint64_t blocks = (uint64_t) per_sm * cfg.sm_count;
        if (blocks > want) blocks = want;
        k_vps_wavefront<BACKWARD><<<(unsigned) blocks, kWfBlock, 0, cfg.stream>>>(S, G, D, seed, spp, spp_begin, ns, n_local, alt_seed,
                                                                                   grad_img, film_w, accum, cfg.counters,
                                                                                   cfg.work_counter);
    } else {
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
    }
    (*cfg.launches)++;
    return cudaGetLastError();
}

cudaError_t launch_vps_primal(const LaunchCfg &cfg, const SceneK &S, const GridK &G, uint32_t seed, uint32_t spp,
                              uint32_t spp_begin, uint32_t spp_end, float4 *accum) {
    GradK D; D.gsig = nullptr; D.galb = nullptr;
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
