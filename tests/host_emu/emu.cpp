// emu.cpp — TEST HARNESS: compiles the per-lane event machines of
// recursive_control_variates_b200/csrc for the host (RCV_HOST_EMU) and
// single-steps them lane by lane, so that their logic and arithmetic can be
// compared with the oracle without a GPU.  Never part of the product.
#define RCV_HOST_EMU 1
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "../../recursive_control_variates_b200/csrc/rcv_scene.h"
#include "../../recursive_control_variates_b200/csrc/rcv_twostate_wf.cuh"
#include "../../recursive_control_variates_b200/csrc/rcv_volpathsimple_wf.cuh"
#include "../../recursive_control_variates_b200/csrc/rcv_surface_wf.cuh"

using namespace rcv;

// the collision handler the pool kernels run (rcv_trace.cu): the speculative two-step form when RCV_SPEC_COLLIDE > 1
#if RCV_SPEC_COLLIDE > 1
#define EMU_VP_COLLIDE vp_collide_spec
#else
#define EMU_VP_COLLIDE vp_collide
#endif

namespace {

int g_flat = 0;      // emu_set_flat: run pass 2 of the volpathsimple adjoint through vp_vertex2_flat (as the pool kernel does)

// SceneK of a description, with the `envmap` tables (host memory) when the scene has one
struct HostScene {
    SceneK S;
    EnvTables env;
    explicit HostScene(const rcv_scene_desc &d) {
        scene_from_desc(d, S);
        if (d.emitter_type == RCV_EMITTER_ENVMAP) {
            env_tables_build(d, env);
            S.env_data = env.data.data(); S.env_cdf_m = env.cdf_m.data(); S.env_cdf_c = env.cdf_c.data(); S.env_pdf = env.pdf.data();
        }
    }
};

struct Packed {
    std::vector<float2> sig2;
    std::vector<float4> blk2, blk1;
    std::vector<float4> alb2;
    float maj[2];
    std::vector<float2> mgrid;      // majorant supergrid (factor > 0)
    GridKS G;
};

void pack(const rcv_scene_desc &d, const float *so, const float *ao, const float *sn, const float *an, int nz, int ny, int nx,
          Packed &P, int factor = 0) {
    size_t n = (size_t) nz * ny * nx;
    P.sig2.resize(n); P.alb2.resize(2 * n);
    float mo = so[0], mn = sn[0];
    for (size_t v = 0; v < n; ++v) {
        P.sig2[v] = make_float2(so[v], sn[v]);
        P.alb2[2 * v] = make_float4(ao[3 * v], ao[3 * v + 1], ao[3 * v + 2], 0.f);
        P.alb2[2 * v + 1] = make_float4(an[3 * v], an[3 * v + 1], an[3 * v + 2], 0.f);
        mo = fmax_(mo, so[v]); mn = fmax_(mn, sn[v]);
    }
    P.maj[0] = d.sigma_t_scale * mo; P.maj[1] = d.sigma_t_scale * mn;
    // gather bricks, as k_build_bricks does on the device
    size_t ncell = (size_t) (nz + 1) * (ny + 1) * (nx + 1);
    P.blk2.resize(4 * ncell); P.blk1.resize(2 * ncell);
    for (size_t c = 0; c < ncell; ++c) {
        size_t r = c;
        int cx = (int) (r % (nx + 1)); r /= (nx + 1);
        int cy = (int) (r % (ny + 1)); r /= (ny + 1);
        int cz = (int) r;
        float2 v8[8];
        for (int k = 0; k < 8; ++k) v8[k] = P.sig2[cell_corner(nx, ny, nz, cx, cy, cz, k)];
        for (int j = 0; j < 4; ++j) P.blk2[4 * c + j] = make_float4(v8[2 * j].x, v8[2 * j].y, v8[2 * j + 1].x, v8[2 * j + 1].y);
        P.blk1[2 * c] = make_float4(v8[0].y, v8[1].y, v8[2].y, v8[3].y);
        P.blk1[2 * c + 1] = make_float4(v8[4].y, v8[5].y, v8[6].y, v8[7].y);
    }
    P.G.sig2 = P.sig2.data(); P.G.blk2 = P.blk2.data(); P.G.blk1 = P.blk1.data(); P.G.alb2 = P.alb2.data();
    P.G.nx = nx; P.G.ny = ny; P.G.nz = nz; P.G.majorants = P.maj;
    P.G.mgrid = nullptr; P.G.mx = P.G.my = P.G.mz = 0; P.G.vox[0] = P.G.vox[1] = P.G.vox[2] = 0.f;
    if (factor > 0) {
        // as k_majorant_grid does on the device: voxels [c f - 1, c f + f] per axis, clamped
        int mx = nx / factor, my = ny / factor, mz = nz / factor;
        P.mgrid.resize((size_t) mx * my * mz);
        for (int cz = 0; cz < mz; ++cz)
            for (int cy = 0; cy < my; ++cy)
                for (int cx = 0; cx < mx; ++cx) {
                    float mo2 = 0.f, mn2 = 0.f;
                    for (int dz = -1; dz <= factor; ++dz)
                        for (int dy = -1; dy <= factor; ++dy)
                            for (int dx = -1; dx <= factor; ++dx) {
                                int z = std::min(std::max(cz * factor + dz, 0), nz - 1), y = std::min(std::max(cy * factor + dy, 0), ny - 1),
                                    x = std::min(std::max(cx * factor + dx, 0), nx - 1);
                                float2 v = P.sig2[((size_t) z * ny + y) * nx + x];
                                mo2 = fmax_(mo2, v.x); mn2 = fmax_(mn2, v.y);
                            }
                    P.mgrid[((size_t) cz * my + cy) * mx + cx] = make_float2(d.sigma_t_scale * mo2, d.sigma_t_scale * mn2);
                }
        P.G.mgrid = P.mgrid.data(); P.G.mx = mx; P.G.my = my; P.G.mz = mz;
        P.G.vox[0] = 1.f / (float) mx; P.G.vox[1] = 1.f / (float) my; P.G.vox[2] = 1.f / (float) mz;
    }
}

// RCV_GRAD_BRICKS: host stand-in for the gradient bricks and launch_unbrick
struct HostBricks {
    std::vector<float4> b;
    int nx, ny, nz;
    HostBricks(int nx_, int ny_, int nz_, GradK &D) : b((size_t) 2 * (nx_ + 1) * (ny_ + 1) * (nz_ + 1), make_float4(0.f, 0.f, 0.f, 0.f)),
                                                     nx(nx_), ny(ny_), nz(nz_) {
        D.gbrick = RCV_GRAD_BRICKS ? b.data() : nullptr;
    }
    void fold(float *gsig) {
        if (!RCV_GRAD_BRICKS) return;
        size_t c = 0;
        for (int cz = 0; cz <= nz; ++cz)
            for (int cy = 0; cy <= ny; ++cy)
                for (int cx = 0; cx <= nx; ++cx, ++c) {
                    const float v[8] = {b[2 * c].x, b[2 * c].y, b[2 * c].z, b[2 * c].w, b[2 * c + 1].x, b[2 * c + 1].y, b[2 * c + 1].z, b[2 * c + 1].w};
                    for (int k = 0; k < 8; ++k) gsig[cell_corner(nx, ny, nz, cx, cy, cz, k)] += v[k];
                }
    }
};

struct HostQueue {
    std::vector<float4> r0, r1, r2, r3;
    unsigned long long count;
    PathQueue Q;
    explicit HostQueue(size_t n) : r0(n), r1(n), r2(n), r3(n), count(0) {
        Q.r0 = r0.data(); Q.r1 = r1.data(); Q.r2 = r2.data(); Q.r3 = nullptr; Q.count = &count;
    }
};

// raygen for all lanes (as the GPU's ray-generation kernel), then the event
// machines record by record
template <int MODE, class GRID>
void run_twostates(const SceneK &S, const GRID &G, uint32_t seed, uint32_t spp, float4 *acc_old, float4 *acc_new, uint64_t *cnt) {
    const float mu = combined_majorant<MODE>(G);
    const float inv_mu = 1.f / mu;
    TsSink sk; sk.acc_old = acc_old; sk.acc_new = acc_new; sk.n_coll = sk.n_scat = sk.n_paths = 0;
    uint64_t n = (uint64_t) S.W * S.H * spp;
    HostQueue hq(n);
    for (uint64_t idx = 0; idx < n; ++idx) {
        TsLane ln; ln.ev = EV_INIT;
        ts_raygen<MODE>(S, ln, sk, seed, spp, idx);
        if (ln.ev == EV_COLL) ts_store(ln, hq.Q, hq.count++);
    }
    for (unsigned long long pos = 0; pos < hq.count; ++pos) {
        TsLane ln;
        ts_load(ln, hq.Q, pos);
        while (ln.ev != EV_INIT) {
            if (ln.ev == EV_COLL) ts_collide<MODE>(S, G, mu, inv_mu, ln, sk);
            else if (ln.ev == EV_SCAT) ts_scatter<MODE>(S, G, ln, sk);
            else ts_slow<MODE>(S, ln, sk);
        }
    }
    cnt[0] += sk.n_coll; cnt[1] += sk.n_scat; cnt[6] += sk.n_paths;
}

template <bool BW, class GRID>
void run_vps(const SceneK &S, const GRID &G, VpSink &sk) {
    const float mu = G.majorants[1];
    const float inv_mu = 1.f / mu;
    uint64_t n = (uint64_t) S.W * S.H * sk.spp;
    HostQueue hq(n);
    std::vector<float4> arena((size_t) 4 * (S.max_depth + 2) + kNeeCap);       // one arena slot: the lanes run one after the other
    sk.arena = arena.data(); sk.vcap = (uint32_t) S.max_depth + 2u;
    for (uint64_t idx = 0; idx < n; ++idx) {
        VpLane ln; ln.ev = EV_INIT; ln.flags = 0;
        vp_raygen<BW>(S, ln, sk, idx);
        if (ln.ev == EV_COLL) {
            if (BW) {      // as k_vps_raygen: the record carries dL
                hq.Q.r3 = hq.r3.data();
                float fx, fy;
                film_position(S, sk.seed, sk.spp, ln.idx, fx, fy);
                V3 dL = film_gather(S, sk.grad_img, sk.film_w, fx, fy);
                hq.r3[hq.count] = make_float4(dL.x, dL.y, dL.z, 0.f);
            }
            vp_store(ln, hq.Q, hq.count++);
        }
    }
    for (unsigned long long pos = 0; pos < hq.count; ++pos) {
        VpLane ln; ln.flags = 0;
        vp_load(ln, sk, hq.Q, pos, 0u);
        while (ln.ev != EV_INIT) {
            if (ln.ev == EV_P2) {
                // the flattened pass 2: every vertex on its own, from the lane's pass-1 result (k0 = 0) or from the running
                // result a previous round left (k0 = k): both must give what the list walk gives
                V3 run = ln.result;
                for (int k = 0; k < ln.nvert; ++k)
                    run = (k & 1) ? vp_vertex2_flat(S, G, sk, ln.vbase, k, k, run, ln.dL, ln.idx)
                                  : vp_vertex2_flat(S, G, sk, ln.vbase, 0, k, ln.result, ln.dL, ln.idx);
                (void) run;
                sk.n_paths++;
                ln.ev = EV_INIT;
            } else if (ln.ev == EV_COLL) {
                EMU_VP_COLLIDE<BW>(S, G, mu, inv_mu, ln, sk);
            } else if (ln.ev == EV_SCAT) {
                vp_vertex<BW>(S, G, ln, sk);
            } else {
                vp_slow<BW>(S, G, ln, sk);
            }
        }
    }
}

// use_drt with sub-sampling on the event machines: stage 1 (adjoint + reservoir -> job records), the job step
// (vp_drt_sample) and stage 2 (the recursive radiance estimate), as k_vps_pool<true, GRID, 1>, k_drt_sample and
// k_vps_pool<false, GRID, 2> run them
template <class GRID>
void run_vps_drt(const SceneK &S, const GRID &G, VpSink &sk) {
    const float mu = G.majorants[1];
    const float inv_mu = 1.f / mu;
    uint64_t n = (uint64_t) S.W * S.H * sk.spp;
    HostQueue hq(n), jq(n);
    jq.Q.r3 = jq.r3.data();
    std::vector<float4> arena((size_t) 4 * (S.max_depth + 2) + kNeeCap);
    sk.arena = arena.data(); sk.vcap = (uint32_t) S.max_depth + 2u;
    sk.jobs = jq.Q;
    hq.Q.r3 = hq.r3.data();
    for (uint64_t idx = 0; idx < n; ++idx) {
        VpLane ln; ln.ev = EV_INIT; ln.flags = 0;
        vp_raygen<true>(S, ln, sk, idx);
        if (ln.ev == EV_COLL) {
            float fx, fy;
            film_position(S, sk.seed, sk.spp, ln.idx, fx, fy);
            V3 dL = film_gather(S, sk.grad_img, sk.film_w, fx, fy);
            hq.r3[hq.count] = make_float4(dL.x, dL.y, dL.z, 0.f);
            vp_store(ln, hq.Q, hq.count++);
        }
    }
    for (unsigned long long pos = 0; pos < hq.count; ++pos) {
        VpLane ln; ln.flags = 0;
        vp_load<1>(ln, sk, hq.Q, pos, 0u);
        while (ln.ev != EV_INIT) {
            if (ln.ev == EV_COLL) EMU_VP_COLLIDE<true>(S, G, mu, inv_mu, ln, sk);
            else if (ln.ev == EV_SCAT) vp_vertex<true, 1>(S, G, ln, sk);
            else vp_slow<true, 1>(S, G, ln, sk);
        }
        float4 j0, j1, j2, j3;
        if ((ln.flags & VF_JOB) && vp_drt_job(ln, sk, j0, j1, j2, j3)) {
            jq.r0[jq.count] = j0; jq.r1[jq.count] = j1; jq.r2[jq.count] = j2; jq.r3[jq.count] = j3;
            jq.count++;
        }
    }
    hq.count = 0;
    for (unsigned long long i = 0; i < jq.count; ++i) {
        float4 job[4] = {jq.r0[i], jq.r1[i], jq.r2[i], jq.r3[i]}, rec[4];
        V3 p, c, Li;
        int r = vp_drt_sample(S, G, sk.D, job, sk.drt_mis != 0, rec, p, c, Li, sk.cnt);
        if (r == 1) vp_drt_scatter(S, G, sk.D, p, c, Li, sk.cnt);
        if (r == 2) {
            rec[0].w = __uint_as_float((uint32_t) i);
            hq.r0[hq.count] = rec[0]; hq.r1[hq.count] = rec[1]; hq.r2[hq.count] = rec[2]; hq.r3[hq.count] = rec[3];
            hq.count++;
            jq.r0[i] = make_float4(p.x, p.y, p.z, c.x);
            jq.r1[i] = make_float4(c.y, c.z, 0.f, 0.f);
        }
    }
    for (unsigned long long pos = 0; pos < hq.count; ++pos) {
        VpLane ln; ln.flags = 0;
        vp_load<2>(ln, sk, hq.Q, pos, 0u);
        while (ln.ev != EV_INIT) {
            if (ln.ev == EV_COLL) EMU_VP_COLLIDE<false>(S, G, mu, inv_mu, ln, sk);
            else if (ln.ev == EV_SCAT) vp_vertex<false, 2>(S, G, ln, sk);
            else vp_slow<false, 2>(S, G, ln, sk);
        }
        if (ln.flags & VF_JOB) {
            float4 a = jq.r0[ln.pix], b = jq.r1[ln.pix];
            vp_drt_scatter(S, G, sk.D, mk3(a.x, a.y, a.z), mk3(a.w, b.x, b.y), ln.result, sk.cnt);
        }
    }
}

}  // namespace

// the surface machine (rcv_surface_wf.cuh), lane by lane from the sensor, as k_surface_wavefront runs it
template <int MODE, bool ADJ, class GRID>
void run_surface(const SceneK &S, const GRID &G, SfSink &sk) {
    const float mu = combined_majorant<MODE>(G);
    const float inv_mu = 1.f / mu;
    uint64_t n = (uint64_t) S.W * S.H * sk.spp;
    for (uint64_t idx = 0; idx < n; ++idx) {
        SfLane ln; ln.flags = 0;
        sf_begin(S, ln, sk.seed, sk.spp, (uint32_t) idx, 0u);
        while (ln.ev != EV_INIT) {
            if (ln.ev == EV_COLL) sf_collide<MODE, ADJ>(S, G, mu, inv_mu, ln, sk);
            else if (ln.ev == EV_SCAT) sf_scatter<MODE, ADJ>(S, G, ln, sk);
            else if (ln.ev == EV_NEE) sf_nee<ADJ>(S, ln);
            else sf_slow<MODE, ADJ>(S, ln, sk);
        }
    }
}

extern "C" {

void emu_set_flat(int on) { g_flat = on; }
// concat_owner (rcv_volpathsimple_wf.cuh) over an array of 32 exclusive prefix sums
int emu_concat_owner(const int *excl, int j) { return concat_owner(j, [&](int t) { return excl[t]; }); }

// scenes with rectangles: the two-state / single-state primal (adj == 0; mode as emu_twostates) or the inherited
// prbvolpath adjoint of state `new` (adj == 1)
int emu_surface(const rcv_scene_desc *d, const float *so, const float *ao, const float *sn, const float *an, int nz, int ny, int nx,
                uint32_t seed, uint32_t spp, int mode, int adj, const float *grad_img, const float *film_w, float *acc_old,
                float *acc_new, float *gsig, float *galb4, uint64_t *counters, int factor) {
    HostScene hs(*d); const SceneK &S = hs.S;
    Packed P; pack(*d, so, ao, sn, an, nz, ny, nx, P, factor);
    for (int i = 0; i < 8; ++i) counters[i] = 0;
    SfSink sk; sk.acc_old = (float4 *) acc_old; sk.acc_new = (float4 *) acc_new; sk.D.gsig = gsig; sk.D.galb = (float4 *) galb4;
    sk.D.gbrick = nullptr;
    sk.grad_img = grad_img; sk.film_w = film_w; sk.seed = seed; sk.spp = spp;
    sk.cnt.coll = sk.cnt.scat = sk.cnt.g_sig = sk.cnt.g_alb = 0; sk.n_paths = 0;
    HostBricks hb(nx, ny, nz, sk.D);
    const GridK &G = P.G;
    if (adj) {
        if (factor > 0) run_surface<2, true>(S, P.G, sk); else run_surface<2, true>(S, G, sk);
        hb.fold(gsig);
    } else if (factor > 0) {
        if (mode == 0) run_surface<0, false>(S, P.G, sk); else if (mode == 1) run_surface<1, false>(S, P.G, sk); else run_surface<2, false>(S, P.G, sk);
    } else {
        if (mode == 0) run_surface<0, false>(S, G, sk); else if (mode == 1) run_surface<1, false>(S, G, sk); else run_surface<2, false>(S, G, sk);
    }
    counters[mode == 0 && !adj ? 0 : 2] = sk.cnt.coll; counters[mode == 0 && !adj ? 1 : 3] = sk.cnt.scat;
    counters[4] = sk.cnt.g_sig; counters[5] = sk.cnt.g_alb; counters[6] = sk.n_paths;
    return 0;
}

// the three DRT stages of the pool tracer (use_drt + use_drt_subsampling; use_mis = use_drt_mis)
int emu_vps_drt_stages(const rcv_scene_desc *d, const float *sn, const float *an, int nz, int ny, int nx, uint32_t seed, uint32_t spp,
                       uint32_t alt_seed, const float *grad_img, const float *film_w, float *gsig, float *galb4,
                       uint64_t *counters, int factor, int use_mis) {
    HostScene hs(*d); const SceneK &S = hs.S;
    Packed P; pack(*d, sn, an, sn, an, nz, ny, nx, P, factor);
    for (int i = 0; i < 8; ++i) counters[i] = 0;
    VpSink sk; sk.accum = nullptr; sk.D.gsig = gsig; sk.D.galb = (float4 *) galb4; sk.grad_img = grad_img; sk.film_w = film_w;
    sk.seed = seed; sk.spp = spp; sk.alt_seed = alt_seed; sk.drt_mis = use_mis; sk.flat = 0; sk.coop = 0; sk.coop_n = 0;
    sk.cnt.coll = sk.cnt.scat = sk.cnt.g_sig = sk.cnt.g_alb = 0; sk.n_paths = 0;
    HostBricks hb(nx, ny, nz, sk.D);
    if (factor > 0) run_vps_drt(S, P.G, sk);
    else run_vps_drt(S, (const GridK &) P.G, sk);
    hb.fold(gsig);
    counters[2] = sk.cnt.coll; counters[3] = sk.cnt.scat; counters[4] = sk.cnt.g_sig; counters[5] = sk.cnt.g_alb;
    counters[6] = sk.n_paths;
    return 0;
}


// factor > 0: majorant supergrid + DDA (the GridKS instantiation of the handlers)
int emu_twostates_sg(const rcv_scene_desc *d, const float *so, const float *ao, const float *sn, const float *an, int nz, int ny,
                     int nx, uint32_t seed, uint32_t spp, int mode, int factor, float *acc_old, float *acc_new, uint64_t *counters) {
    HostScene hs(*d); const SceneK &S = hs.S;
    Packed P; pack(*d, so, ao, sn, an, nz, ny, nx, P, factor);
    for (int i = 0; i < 8; ++i) counters[i] = 0;
    float4 *ao4 = (float4 *) acc_old, *an4 = (float4 *) acc_new;
    if (factor > 0) {
        if (mode == 0) run_twostates<0>(S, P.G, seed, spp, ao4, an4, counters);
        else if (mode == 1) run_twostates<1>(S, P.G, seed, spp, ao4, an4, counters);
        else run_twostates<2>(S, P.G, seed, spp, ao4, an4, counters);
    } else {
        const GridK &G = P.G;
        if (mode == 0) run_twostates<0>(S, G, seed, spp, ao4, an4, counters);
        else if (mode == 1) run_twostates<1>(S, G, seed, spp, ao4, an4, counters);
        else run_twostates<2>(S, G, seed, spp, ao4, an4, counters);
    }
    return 0;
}
int emu_twostates(const rcv_scene_desc *d, const float *so, const float *ao, const float *sn, const float *an, int nz, int ny,
                  int nx, uint32_t seed, uint32_t spp, int mode, float *acc_old, float *acc_new, uint64_t *counters) {
    return emu_twostates_sg(d, so, ao, sn, an, nz, ny, nx, seed, spp, mode, 0, acc_old, acc_new, counters);
}

// volpathsimple: primal render (backward == 0) or the two-pass adjoint; the
// gradient grids use the library's layout (float + float4 per voxel)
int emu_vps_sg(const rcv_scene_desc *d, const float *sn, const float *an, int nz, int ny, int nx, uint32_t seed, uint32_t spp,
               int backward, uint32_t alt_seed, const float *grad_img, const float *film_w, float *accum, float *gsig,
               float *galb4, uint64_t *counters, int factor) {
    HostScene hs(*d); const SceneK &S = hs.S;
    Packed P; pack(*d, sn, an, sn, an, nz, ny, nx, P, factor);
    for (int i = 0; i < 8; ++i) counters[i] = 0;
    VpSink sk; sk.accum = (float4 *) accum; sk.D.gsig = gsig; sk.D.galb = (float4 *) galb4; sk.grad_img = grad_img; sk.film_w = film_w;
    sk.seed = seed; sk.spp = spp; sk.alt_seed = alt_seed; sk.drt_mis = 0; sk.flat = g_flat; sk.coop = 0; sk.coop_n = 0;
    sk.cnt.coll = sk.cnt.scat = sk.cnt.g_sig = sk.cnt.g_alb = 0; sk.n_paths = 0;
    HostBricks hb(nx, ny, nz, sk.D);
    if (factor > 0) {
        if (backward) run_vps<true>(S, P.G, sk); else run_vps<false>(S, P.G, sk);
    } else {
        const GridK &G = P.G;
        if (backward) run_vps<true>(S, G, sk); else run_vps<false>(S, G, sk);
    }
    if (backward) hb.fold(gsig);
    counters[2] = sk.cnt.coll; counters[3] = sk.cnt.scat; counters[4] = sk.cnt.g_sig; counters[5] = sk.cnt.g_alb;
    counters[6] = sk.n_paths;
    return 0;
}
// the one-thread-per-lane adjoint (vps_backward_lane): drt_flags = 0 is use_drt = False, else bit 0 use_drt,
// bit 1 use_drt_subsampling, bit 2 use_drt_mis
int emu_vps_backward_lane(const rcv_scene_desc *d, const float *sn, const float *an, int nz, int ny, int nx, uint32_t seed,
                          uint32_t spp, uint32_t alt_seed, const float *grad_img, const float *film_w, float *gsig, float *galb4,
                          uint64_t *counters, int factor, int drt_flags) {
    HostScene hs(*d); const SceneK &S = hs.S;
    Packed P; pack(*d, sn, an, sn, an, nz, ny, nx, P, factor);
    for (int i = 0; i < 8; ++i) counters[i] = 0;
    GradK D; D.gsig = gsig; D.galb = (float4 *) galb4;
    HostBricks hb(nx, ny, nz, D);
    VpsCounts cnt; cnt.coll = cnt.scat = cnt.g_sig = cnt.g_alb = 0;
    uint64_t n = (uint64_t) S.W * S.H * spp;
    for (uint64_t idx = 0; idx < n; ++idx) {
        if (factor > 0) vps_backward_lane(S, P.G, D, seed, spp, idx, alt_seed, grad_img, film_w, drt_flags, cnt);
        else vps_backward_lane(S, (const GridK &) P.G, D, seed, spp, idx, alt_seed, grad_img, film_w, drt_flags, cnt);
    }
    hb.fold(gsig);
    counters[2] = cnt.coll; counters[3] = cnt.scat; counters[4] = cnt.g_sig; counters[5] = cnt.g_alb; counters[6] = 2 * n;
    return 0;
}
int emu_vps(const rcv_scene_desc *d, const float *sn, const float *an, int nz, int ny, int nx, uint32_t seed, uint32_t spp,
            int backward, uint32_t alt_seed, const float *grad_img, const float *film_w, float *accum, float *gsig,
            float *galb4, uint64_t *counters) {
    return emu_vps_sg(d, sn, an, nz, ny, nx, seed, spp, backward, alt_seed, grad_img, film_w, accum, gsig, galb4, counters, 0);
}

}  // extern "C"
