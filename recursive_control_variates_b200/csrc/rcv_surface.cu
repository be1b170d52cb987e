// rcv_surface.cu — kernels for scenes with surfaces (rectangles, an area light): the two-state primal render and
// the adjoint it inherits (upstream prbvolpath), both on the lane event machine of rcv_surface_wf.cuh.
//
// Schedule: persistent warps, one lane state per thread in registers; every round the warp votes for the event class
// with the most waiting lanes (tentative collisions first), runs that handler, and refills finished lanes from a
// chunked work counter in the slow step.  There is no ray-generation pass: with walls around the medium almost every
// path has work to do, so a lane starts at the sensor inside the machine.
#include "rcv_internal.h"
#include "rcv_sched.cuh"
#include "rcv_surface_wf.cuh"

namespace rcv {

#ifndef RCV_SF_BLOCK
#define RCV_SF_BLOCK 128
#endif
#ifndef RCV_SF_MINBLOCKS
#define RCV_SF_MINBLOCKS 3
#endif
static constexpr int kSfBlock = RCV_SF_BLOCK;

template <int MODE, bool ADJ, class GRID>
__global__ void __launch_bounds__(kSfBlock, RCV_SF_MINBLOCKS) k_surface_wavefront(const __grid_constant__ SceneK S,
                                                                               const __grid_constant__ GRID G, GradK D,
                                                                               uint32_t seed, uint32_t spp, uint32_t spp_begin,
                                                                               uint32_t ns, unsigned long long n_local,
                                                                               const float *grad_img, const float *film_w,
                                                                               float4 *acc_old, float4 *acc_new, Counters *C,
                                                                               unsigned long long *work) {
    const float mu = combined_majorant<MODE>(G);
    const float inv_mu = 1.f / mu;
    SfLane ln;
    ln.ev = EV_INIT; ln.flags = 0;
    SfSink sk;
    sk.acc_old = acc_old; sk.acc_new = acc_new; sk.D = D; sk.grad_img = grad_img; sk.film_w = film_w;
    sk.seed = seed; sk.spp = spp;
    sk.cnt.coll = sk.cnt.scat = sk.cnt.g_sig = sk.cnt.g_alb = 0; sk.n_paths = 0;
    ChunkCursor cc; cc.next = 0; cc.end = 0; cc.more = true;
    while (true) {
        const int n_coll = __popc(__ballot_sync(0xffffffffu, ln.ev == EV_COLL));
        const int n_scat = __popc(__ballot_sync(0xffffffffu, ln.ev == EV_SCAT));
        const int n_slow = __popc(__ballot_sync(0xffffffffu, ln.ev == EV_SLOW || (ln.ev == EV_INIT && cc.more)));
        const int n_nee = __popc(__ballot_sync(0xffffffffu, ln.ev == EV_NEE));
        if (n_coll == 0 && n_scat == 0 && n_slow == 0 && n_nee == 0) break;
        // the fullest class runs (every lane is in exactly one: the fullest holds a quarter of the live lanes at least)
        int sel = EV_COLL, n_sel = n_coll;
        if (n_slow > n_sel) { sel = EV_SLOW; n_sel = n_slow; }
        if (n_nee > n_sel) { sel = EV_NEE; n_sel = n_nee; }
        if (n_scat > n_sel) { sel = EV_SCAT; n_sel = n_scat; }
        if (sel == EV_COLL) {
            if (ln.ev == EV_COLL) sf_collide<MODE, ADJ>(S, G, mu, inv_mu, ln, sk);
        } else if (sel == EV_SCAT) {
            if (ln.ev == EV_SCAT) sf_scatter<MODE, ADJ>(S, G, ln, sk);
        } else if (sel == EV_NEE) {
            if (ln.ev == EV_NEE) sf_nee<ADJ>(S, ln);
        } else {
            if (ln.ev == EV_SLOW) sf_slow<MODE, ADJ>(S, ln, sk);
            unsigned m_init = __ballot_sync(0xffffffffu, ln.ev == EV_INIT);
            if (m_init && cc.more) {
                unsigned long long pos = take_positions(cc, work, n_local, m_init);
                if (pos != ~0ull) sf_begin(S, ln, seed, spp, (uint32_t) global_lane(pos, spp, spp_begin, ns), 0u);
            }
        }
        __syncwarp();
    }
    counter_flush(C, MODE == 0 ? 0 : 2, sk.cnt.coll);
    counter_flush(C, MODE == 0 ? 1 : 3, sk.cnt.scat);
    counter_flush(C, 4, sk.cnt.g_sig);
    counter_flush(C, 5, sk.cnt.g_alb);
    counter_flush(C, 6, sk.n_paths);
}

template <int MODE, bool ADJ, class GRID>
static cudaError_t launch_surface_kernel(const LaunchCfg &cfg, const SceneK &S, const GRID &G, const GradK &D, uint32_t seed,
                                         uint32_t spp, uint32_t spp_begin, uint32_t ns, uint64_t n_local, const float *grad_img,
                                         const float *film_w, float4 *acc_old, float4 *acc_new) {
    cudaError_t e;
    if ((e = cudaMemsetAsync(cfg.work_counter, 0, sizeof(unsigned long long), cfg.stream)) != cudaSuccess) return e;
    unsigned blocks;
    if ((e = persistent_grid(cfg, k_surface_wavefront<MODE, ADJ, GRID>, kSfBlock, n_local, blocks)) != cudaSuccess) return e;
    k_surface_wavefront<MODE, ADJ, GRID><<<blocks, kSfBlock, 0, cfg.stream>>>(S, G, D, seed, spp, spp_begin, ns, n_local, grad_img,
                                                                          film_w, acc_old, acc_new, cfg.counters,
                                                                          cfg.work_counter);
    (*cfg.launches)++;
    return cudaGetLastError();
}

template <class GRID>
static cudaError_t launch_surface_grid(const LaunchCfg &cfg, const SceneK &S, const GRID &G, int mode, uint32_t seed, uint32_t spp,
                                       uint32_t spp_begin, uint32_t ns, uint64_t n_local, float4 *acc_old, float4 *acc_new) {
    GradK D; D.gsig = nullptr; D.galb = nullptr; D.gbrick = nullptr;
    switch (mode) {
        case 0: return launch_surface_kernel<0, false>(cfg, S, G, D, seed, spp, spp_begin, ns, n_local, nullptr, nullptr, acc_old, acc_new);
        case 1: return launch_surface_kernel<1, false>(cfg, S, G, D, seed, spp, spp_begin, ns, n_local, nullptr, nullptr, acc_old, acc_new);
        default: return launch_surface_kernel<2, false>(cfg, S, G, D, seed, spp, spp_begin, ns, n_local, nullptr, nullptr, acc_old, acc_new);
    }
}

// the two-state (mode 0) or single-state (mode 1 / 2) primal render of a scene with rectangles
cudaError_t launch_surface_primal(const LaunchCfg &cfg, const SceneK &S, const GridKS &G, int mode, uint32_t seed, uint32_t spp,
                                  uint32_t spp_begin, uint32_t spp_end, float4 *acc_old, float4 *acc_new) {
    uint32_t ns = spp_end - spp_begin;
    uint64_t n_local = (uint64_t) S.W * S.H * ns;
    if (n_local == 0) return cudaSuccess;
    if (G.mgrid) return launch_surface_grid<GridKS>(cfg, S, G, mode, seed, spp, spp_begin, ns, n_local, acc_old, acc_new);
    return launch_surface_grid<GridK>(cfg, S, G, mode, seed, spp, spp_begin, ns, n_local, acc_old, acc_new);
}

// RCV_ADJOINT_PRBVOLPATH: both walks of upstream prbvolpath's render_backward on state `new`; works with and without
// rectangles
cudaError_t launch_prb_backward(const LaunchCfg &cfg, const SceneK &S, const GridKS &G, const GradK &D, uint32_t seed_grad,
                                uint32_t spp_grad, uint32_t spp_begin, uint32_t spp_end, const float *grad_img, const float *film_w) {
    uint32_t ns = spp_end - spp_begin;
    uint64_t n_local = (uint64_t) S.W * S.H * ns;
    if (n_local == 0) return cudaSuccess;
    if (G.mgrid)
        return launch_surface_kernel<2, true, GridKS>(cfg, S, G, D, seed_grad, spp_grad, spp_begin, ns, n_local, grad_img, film_w,
                                                      nullptr, nullptr);
    return launch_surface_kernel<2, true, GridK>(cfg, S, G, D, seed_grad, spp_grad, spp_begin, ns, n_local, grad_img, film_w, nullptr,
                                                 nullptr);
}

}  // namespace rcv
