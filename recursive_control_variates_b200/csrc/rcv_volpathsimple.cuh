// rcv_volpathsimple.cuh — per-lane restatement of VolpathSimpleIntegrator
// (plugins/volpathsimple.py) in Primal and Backward mode:
// sample (:51-303), reach_medium (:305-332), sample_real_interaction
// (:336-390), sample_emitter_for_nee / sample_emitter / estimate_transmittance
// (:393-517), backpropagate_transmittance (:597-620); with DRT = true also the
// differential-ratio-tracking in-scattering term: backpropagate_scattering_drt
// (:520-594), sample_recursive (:623-668), PathState / DRTReservoir (:673-778).
//
// Backward mode replays the identical main random stream and turns the AD of
// the trilinear gathers into reductions on the
// gradient grids (d sigma_t: float per voxel, d albedo: float4 per voxel): one
// red.global.add.f32 per corner for every sigma_t term, plus one
// red.global.add.v4.f32 per corner for the albedo at scattering vertices.
#pragma once
#include "rcv_device.cuh"

namespace rcv {

struct VpsCounts {
    unsigned coll, scat, g_sig, g_alb;
};

struct Mei1 {
    float t;      // inf when invalid
    V3 p;
    float st;     // sigma_t(p) (0 when invalid)
    bool valid;
    float mu, inv_mu;   // the majorant the collision was sampled with (volume or supergrid cell at p)
};

// Medium::sample_interaction of state `new`: one majorant for the volume, or DDA through the majorant supergrid
template <class GRID>
__device__ __forceinline__ void sample_interaction1(const SceneK &S, const GRID &G, float mu, float inv_mu, V3 o, V3 d, V3 id,
                                                    float ray_maxt, float u, Mei1 &m, Tri &T, VpsCounts &cnt) {
    float mint, maxt;
    bool active = medium_interval(S, o, d, id, ray_maxt, mint, maxt);
    float tau = -spec_log(1.f - u);
    float t;
    float m_cell = 0.f;
    if constexpr (GRID::kSupergrid) t = active ? ((RCV_DDA_LIN_VPS) ? dda_sample_lin<2>(S, G, o, d, id, mint, maxt, tau, m_cell) : dda_sample<2>(S, G, o, d, id, mint, maxt, tau, m_cell)) : inf_();
    else t = mint + (tau * inv_mu);
    m.valid = active && (t <= maxt);
    m.t = m.valid ? t : inf_();
    m.p = ray_at(o, d, t);
    m.st = 0.f;
    m.mu = mu; m.inv_mu = inv_mu;
    if (m.valid) {
        T = tri_setup(S, G, m.p);
        m.st = lookup_sigma1(S, G, T);
        cnt.coll++;
        if constexpr (GRID::kSupergrid) { m.mu = m_cell; m.inv_mu = 1.f / m.mu; }
    }
}
// the same in two phases (rcv_device.cuh "Two-phase lookups"): _issue samples the distance and starts the brick load,
// _finish interpolates sigma_t
template <class GRID>
__device__ __forceinline__ void sample_interaction1_issue(const SceneK &S, const GRID &G, float mu, float inv_mu, V3 o, V3 d, V3 id,
                                                          float ray_maxt, float u, Mei1 &m, Tri &T, Brick1 &b, VpsCounts &cnt) {
    float mint, maxt;
    bool active = medium_interval(S, o, d, id, ray_maxt, mint, maxt);
    float tau = -spec_log(1.f - u);
    float t;
    float m_cell = 0.f;
    if constexpr (GRID::kSupergrid) t = active ? ((RCV_DDA_LIN_VPS) ? dda_sample_lin<2>(S, G, o, d, id, mint, maxt, tau, m_cell) : dda_sample<2>(S, G, o, d, id, mint, maxt, tau, m_cell)) : inf_();
    else t = mint + (tau * inv_mu);
    m.valid = active && (t <= maxt);
    m.t = m.valid ? t : inf_();
    m.p = ray_at(o, d, t);
    m.st = 0.f;
    m.mu = mu; m.inv_mu = inv_mu;
    T.inside = false;
    if (m.valid) {
        T = tri_setup(S, G, m.p);
        sigma1_issue(G, T, b);
        cnt.coll++;
        if constexpr (GRID::kSupergrid) { m.mu = m_cell; m.inv_mu = 1.f / m.mu; }
    }
}
__device__ __forceinline__ void sample_interaction1_finish(const SceneK &S, Mei1 &m, const Tri &T, const Brick1 &b) {
    if (m.valid) m.st = sigma1_finish(S, T, b);
}
// convenience overload for callers that do not need the interpolation set-up
template <class GRID>
__device__ __forceinline__ void sample_interaction1(const SceneK &S, const GRID &G, float mu, float inv_mu, V3 o, V3 d, V3 id,
                                                    float ray_maxt, float u, Mei1 &m, VpsCounts &cnt) {
    Tri T;
    sample_interaction1(S, G, mu, inv_mu, o, d, id, ray_maxt, u, m, T, cnt);
}

// AD of sigma_t(p) = scale * trilinear(grid): scatter w_corner * (scale * g) on a known interpolation set-up
// the eight reductions of one site: corner weight x sg, k = 4 zbit + 2 ybit + xbit
__device__ __forceinline__ void red_sigma_site(const GridK &G, const GradK &D, const Tri &T, float sg) {
    const float wzy[4] = {T.wz0 * T.wy0, T.wz0 * T.wy1, T.wz1 * T.wy0, T.wz1 * T.wy1};
#if RCV_GRAD_BRICKS
    float4 *b = D.gbrick + 2 * (size_t) tri_cell(G, T);
    red_add_v4(b, (wzy[0] * T.wx0) * sg, (wzy[0] * T.wx1) * sg, (wzy[1] * T.wx0) * sg, (wzy[1] * T.wx1) * sg);
    red_add_v4(b + 1, (wzy[2] * T.wx0) * sg, (wzy[2] * T.wx1) * sg, (wzy[3] * T.wx0) * sg, (wzy[3] * T.wx1) * sg);
#else
    uint32_t r[4];
    tri_rows(G, T, r);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        red_add_f32(D.gsig + r[k] + T.x0, (wzy[k] * T.wx0) * sg);
        red_add_f32(D.gsig + r[k] + T.x1, (wzy[k] * T.wx1) * sg);
    }
#endif
}
__device__ __forceinline__ void scatter_sigma_t_T(const GridK &G, const GradK &D, const Tri &T, float sg, VpsCounts &cnt) {
    if (!T.inside) return;
    red_sigma_site(G, D, T, sg);
    cnt.g_sig++;
}
__device__ __forceinline__ void scatter_sigma_t(const SceneK &S, const GridK &G, const GradK &D, V3 p, float g, VpsCounts &cnt) {
    Tri T = tri_setup(S, G, p);
    scatter_sigma_t_T(G, D, T, S.scale * g, cnt);
}

// volpathsimple.py:449-517, ratio tracking from o along d over [0, tmax]
template <bool ADJ, class GRID>
__device__ __forceinline__ float estimate_transmittance(const SceneK &S, const GRID &G, const GradK &D, float mu, float inv_mu,
                                                        Pcg &rng, V3 o, V3 d, V3 id, float tmax, bool active, V3 adjoint,
                                                        VpsCounts &cnt) {
    o = ray_at(o, d, 0.f);                                            // :467 (tmin = 0)
    float tr_acc = active ? 1.f : 0.f;                                // :472
    while (active) {                                                  // :478
        float u = pcg_f32(rng);
        Mei1 m;
        sample_interaction1(S, G, mu, inv_mu, o, d, id, tmax, u, m, cnt); // :482
        float sn = m.valid ? (m.mu - m.st) : 0.f;
        float ce = m.valid ? m.mu : 0.f;
        float tr = (ce != 0.f) ? (sn * m.inv_mu) : sn;                // :486-489
        if (m.t > tmax) m.t = inf_();                                 // :493
        if (!(m.t < inf_())) break;                                   // :494
        if (ADJ) {                                                    // :496-505
            float g = 0.f;
            if (tr > 0.f) {
                float k = (-m.inv_mu) / tr;
                g += adjoint.x * k; g += adjoint.y * k; g += adjoint.z * k;
            }
            scatter_sigma_t(S, G, D, m.p, g, cnt);
        }
        tr_acc = tr_acc * tr;                                         // :508
        o = m.p;                                                      // :510
        tmax = tmax - m.t;
        if (!(tr_acc != 0.f)) break;                                  // :515
    }
    return tr_acc;
}

// volpathsimple.py:419-446; returns the scalar transmittance, the sampled direction in `dir`
template <bool ADJ, class GRID>
__device__ __forceinline__ float vps_sample_emitter(const SceneK &S, const GRID &G, const GradK &D, float mu, float inv_mu,
                                                    Pcg &rng, V3 ref_p, V3 adjoint, V3 &dir, VpsCounts &cnt) {
    float s2x = pcg_f32(rng), s2y = pcg_f32(rng);                     // :431
    V3 d = emitter_sample_dir(S, s2x, s2y);
    dir = d;
    if (S.emitter == 1) {                                             // :434-436 a sample of zero density is invalid
        float epdf;
        (void) emitter_sample_value(S, d, epdf);
        if (epdf == 0.f) return 0.f;
    }
    V3 id = rcp3(d);
    Hit si = box_hit(S, ref_p, d, id, RCV_LARGEST);                   // :440-441
    return estimate_transmittance<ADJ>(S, G, D, mu, inv_mu, rng, ref_p, d, id, si.t, si.valid, adjoint, cnt);   // :443
}

// PathState (volpathsimple.py:673-719): where a recursive call of `sample` resumes
struct VpsState {
    bool active;
    int depth;
    Hit si;
    float ray_maxt, last_pdf;
};
// use_drt_subsampling / use_drt_mis (:41-47)
struct DrtOpts {
    bool subsampling, mis;
};
// DRTReservoir with n = 1 (:743-778): one path depth, kept with probability proportional to the throughput there
struct DrtReservoir {
    V3 wsum, cur, o, d;
    float ray_maxt;
    int depth;            // -1: never set
    bool active;
};
__device__ __forceinline__ float mean3(V3 a) { return ((a.x + a.y) + a.z) / 3.f; }

template <class GRID>
__device__ __noinline__ void drt_backpropagate(const SceneK &S, const GRID &G, const GradK &D, float mu, float inv_mu, Pcg &alt,
                                               V3 o, V3 d, float ray_maxt, int depth, V3 adjoint, bool active, bool use_mis,
                                               VpsCounts &cnt);

// ADJ: Backward mode.  DRT: use_drt (Backward only).  FROM_STATE: resume at `ps` (the recursive call of :656-665)
// instead of starting at the sensor.
template <bool ADJ, bool DRT, bool FROM_STATE, class GRID>
__device__ __forceinline__ void vps_trace_core(const SceneK &S, const GRID &G, const GradK &D, Pcg &rng, Pcg &alt, V3 o, V3 d, V3 dL,
                                               V3 state_in, V3 &result, VpsCounts &cnt, const VpsState &ps, DrtOpts drt) {
    static_assert(!DRT || ADJ, "DRT is a term of the adjoint");
    const float mu = __ldg(G.majorants + 1);       // state `new` (twostatemedium.py:125-136)
    const float inv_mu = 1.f / mu;
    V3 id = rcp3(d);
    result = ADJ ? state_in : mk3(0, 0, 0);                           // :68
    V3 thr = mk3(1, 1, 1);
    int depth = 0;
    Hit si;
    bool escaped, active, has_scattered = false;
    float ray_maxt = RCV_LARGEST, last_pdf = 1.f;
    if constexpr (FROM_STATE) {                                       // :75-81, :98-100
        active = ps.active; depth = ps.depth; si = ps.si; ray_maxt = ps.ray_maxt;
        escaped = false;
        has_scattered = active;
        last_pdf = ps.last_pdf;
    } else {
        (void) pcg_f32(rng);                                          // :84
        // reach_medium (:305-332)
        si = box_hit(S, o, d, id, RCV_LARGEST);                       // :311
        escaped = !si.valid; active = si.valid;
        if (active) {
            o = offset_p(si.p, si.n, d);                              // :319
            Hit si2 = box_hit(S, o, d, id, RCV_LARGEST);              // :320
            active = si2.valid;                                       // :323
            if (active) { ray_maxt = isfinite(si2.t) ? si2.t : RCV_LARGEST; si = si2; }   // :329-330
        }
    }
    DrtReservoir R;                                                   // :107-109
    if constexpr (DRT) {
        R.wsum = R.cur = R.o = R.d = mk3(0, 0, 0);
        R.ray_maxt = 0.f; R.depth = -1; R.active = active;
    }
    if (active) (void) pcg_f32(rng);                                  // :112 alt_seed_rnd

    while (active) {                                                  // :127
        float q = fmin_(max3(thr), 0.99f);                            // :130
        bool perform_rr = depth > S.rr_depth;
        float rr = pcg_f32(rng);                                      // :133
        if (!(any_nz(thr) && (!perform_rr || (rr < q)))) break;
        if (perform_rr) thr = thr * (1.f / q);                        // :134

        // ---- sample_real_interaction (:336-390): delta tracking to the next real collision
        V3 ro = o; float rmaxt = ray_maxt, running_t = 0.f, mei_t;
        bool did_scatter;
        while (true) {                                                // :360
            float u = pcg_f32(rng);
            Mei1 m;
            sample_interaction1(S, G, mu, inv_mu, ro, d, id, rmaxt, u, m, cnt);   // :361
            mei_t = m.t + running_t;                                  // :364
            if (!m.valid) { did_scatter = false; break; }             // :371
            float r = (m.mu != 0.f) ? (m.st * m.inv_mu) : 0.f;        // :367
            float xi = pcg_f32(rng);                                  // :372
            if (!(xi >= r)) { did_scatter = true; break; }
            ro = ro + d * m.t;                                        // :378-380
            rmaxt = rmaxt - m.t;
            running_t = running_t + m.t;
        }
        bool did_escape = !did_scatter;
        has_scattered = has_scattered || did_scatter;                 // :150
        V3 mei_p = mk3(0, 0, 0), albedo = mk3(1, 1, 1);
        float st = 0.f;
        Tri T;
        if (did_scatter) {
            mei_p = ray_at(o, d, mei_t);                              // :384
            T = tri_setup(S, G, mei_p);
            if (ADJ) st = lookup_sigma1(S, G, T);                     // :386-388 (value only used by the adjoint)
            albedo = lookup_albedo(G, T, 1);                          // :154
            cnt.scat++;
        }
        if (ADJ) {
            if constexpr (DRT) {                                      // :156-163, :533-549
                if (drt.subsampling) {
                    float xi = pcg_f32(alt);                          // :546
                    R.wsum = R.wsum + thr;                            // :757-765
                    if (xi <= mean3(thr / R.wsum)) {
                        R.cur = thr; R.depth = depth; R.o = o; R.d = d; R.ray_maxt = ray_maxt; R.active = true;
                    }
                } else {
                    drt_backpropagate(S, G, D, mu, inv_mu, alt, o, d, ray_maxt, depth, dL * thr, true, drt.mis, cnt);
                }
            }
            if (did_scatter && (!DRT || drt.mis)) {
                // scattering term (:165-185): backward_from(drt_mis_weight * dL * (sigma_t * albedo) * Li * inv_pdf)
                V3 inv_pdf = splat3(1.f / st);                        // :175
                V3 Li = result / vmax3(splat3(1e-8f), albedo);        // :180
                V3 c = (dL * Li) * inv_pdf;
                if constexpr (DRT) {                                  // :167-170: power heuristic against the DRT estimate
                    float s2 = st * st;
                    c = c * (s2 / (1.f + s2));
                }
                float g_st = (c.x * albedo.x + c.y * albedo.y) + c.z * albedo.z;
                V3 g_al = c * splat3(st);
                if (T.inside) {
                    float sg = S.scale * g_st;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        float w = tri_weight(T, k);
                        uint32_t v = tri_voxel(G, T, k);
#if !RCV_GRAD_BRICKS
                        red_add_f32(D.gsig + v, w * sg);
#endif
                        red_add_v4(D.galb + v, w * g_al.x, w * g_al.y, w * g_al.z, 0.f);
                    }
#if RCV_GRAD_BRICKS
                    red_sigma_site(G, D, T, sg);
#endif
                    cnt.g_sig++; cnt.g_alb++;
                }
            }
            // transmittance term (:194-202, :597-620): 4 uniform points on the last segment
            V3 aw = dL * result;                                      // :198
            float interval = did_escape ? si.t : mei_t;               // :601
            float inv_pdf = interval / 4.f;                           // :617
            float g = -((aw.x * inv_pdf) + (aw.y * inv_pdf) + (aw.z * inv_pdf));
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                float t = pcg_f32(alt) * interval;                    // :608
                scatter_sigma_t(S, G, D, ray_at(o, d, t), g, cnt);    // :609-613
            }
        }
        thr = thr * albedo;                                           // :206
        if (did_scatter) depth += 1;                                  // :212
        active = did_scatter && (depth < S.max_depth);                // :213

        V3 wi = mk3(-d.x, -d.y, -d.z);
        if (S.use_nee && did_scatter && active) {                     // :219-228, :393-416
            Pcg nee_rng = rng;                                        // :396 sampler.clone()
            V3 dir;
            float tr = vps_sample_emitter<false>(S, G, D, mu, inv_mu, rng, mei_p, mk3(0, 0, 0), dir, cnt);
            float epdf;
            V3 emitted = emitter_sample_value(S, dir, epdf) * splat3(tr);
            float phase_val = phase_eval(S, wi, dir);                 // :401
            V3 nee = ((thr * phase_val) * mis_weight(epdf, phase_val)) * emitted;   // :404
            if (ADJ) {
                V3 dir2;
                (void) vps_sample_emitter<true>(S, G, D, mu, inv_mu, nee_rng, mei_p, dL * nee, dir2, cnt);   // :412-414
                result = result - nee;                                // :227
            } else {
                result = result + nee;                                // :225
            }
        }
        if (did_scatter) {                                            // :234-243
            (void) pcg_f32(rng);
            float s2x = pcg_f32(rng), s2y = pcg_f32(rng);
            float pdf;
            V3 wo = phase_sample(S, wi, s2x, s2y, pdf);
            o = mei_p; d = wo; id = rcp3(d); ray_maxt = RCV_LARGEST;
            last_pdf = pdf;
        }
        si = box_hit(S, o, d, id, ray_maxt);                          // :246-248 (needs_update is always true here)
        ray_maxt = isfinite(si.t) ? si.t : RCV_LARGEST;
        if (did_scatter && !si.valid) active = false;                 // :253-254
        if (did_escape) {                                             // :257-258
            o = offset_p(si.p, si.n, d);
            ray_maxt = RCV_LARGEST;
            escaped = true;
        }
    }
    if constexpr (DRT) {                                              // :262-272: the one delayed DRT estimate of the lane
        if (drt.subsampling) {
            float dn = mean3(R.cur);                                  // :768-772
            V3 wgt = (dn != 0.f) ? R.cur * (mean3(R.wsum) / dn) : mk3(0, 0, 0);
            drt_backpropagate(S, G, D, mu, inv_mu, alt, R.o, R.d, R.ray_maxt, R.depth, wgt * dL, R.active && (R.depth != -1),
                              drt.mis, cnt);
        }
    }
    if (!ADJ) {                                                       // :276-300
        if (escaped && si.valid) si = box_hit(S, o, d, id, ray_maxt); // :277-278
        if (escaped && !si.valid && !((depth <= 0) && S.hide_emitters)) {
            float w = 1.f, epdf;
            const V3 env = emitter_eval(S, d, epdf);
            if (S.use_nee) w = mis_weight(last_pdf, has_scattered ? epdf : 0.f);   // :286-291
            result = result + (thr * w) * env;                        // :298
        }
    }
}

template <bool ADJ, class GRID>
__device__ __forceinline__ void vps_trace(const SceneK &S, const GRID &G, const GradK &D, Pcg &rng, Pcg &alt, V3 o, V3 d, V3 dL,
                                          V3 state_in, V3 &result, VpsCounts &cnt) {
    VpsState ps; ps.active = false; ps.depth = 0; ps.si.valid = false; ps.ray_maxt = RCV_LARGEST; ps.last_pdf = 1.f;
    DrtOpts drt; drt.subsampling = drt.mis = false;
    vps_trace_core<ADJ, false, false>(S, G, D, rng, alt, o, d, dL, state_in, result, cnt, ps, drt);
}

// ---------------------------------------------------------------------------
// DRT in-scattering term (use_drt = True, volpathsimple.py:520-594, :623-668)
// ---------------------------------------------------------------------------
// Medium::sample_interaction_drt [UPSTREAM, source absent]: differential ratio tracking restated from the published
// algorithm.  Tentative collisions t_i are sampled with the volume's majorant; w_i = prod_{j<i} (1 - sigma_t(t_j) / mu)
// estimates the transmittance up to t_i; one t_i is kept with probability proportional to w_i and the estimate carries
// (sum_i w_i) / mu.  Two secondary-sampler draws per tentative collision (distance, selection).
//
// RCV_DRT_AHEAD = K > 1 (build option; no measurable difference on the reference-default line, profiles/README.md r2al,
// so the default is the plain loop): where the tentative collisions fall depends on the secondary sampler alone (one majorant for
// the volume), not on the densities read there, so the walk is taken K collisions at a time — K positions and selection
// draws from a copy of the sampler, K bricks requested together — and the weights are then accumulated in order with
// the plain loop's tests.  When the walk stops inside a group the sampler is set to where the plain loop would have
// left it (two draws per completed collision, one for the distance that overshoots).  Same values, same order; the
// interpolation set-up is computed again when a brick is consumed rather than kept in registers.
#ifndef RCV_DRT_AHEAD
#define RCV_DRT_AHEAD 1
#endif
template <class GRID>
__device__ __forceinline__ float sample_interaction_drt(const SceneK &S, const GRID &G, float mu, float inv_mu, Pcg &alt, V3 o, V3 d,
                                                        V3 id, float ray_maxt, bool active, float &sel_t, VpsCounts &cnt) {
    float mint, maxt;
    active = medium_interval(S, o, d, id, ray_maxt, mint, maxt) && active;
    sel_t = inf_();
    float wsum = 0.f, w = 1.f, t = mint;
#if RCV_DRT_AHEAD > 1
    constexpr int K = RCV_DRT_AHEAD;
    while (active) {
        Pcg r = alt;
        float tk[K], xik[K];
        Brick1 bk[K];
        int n = 0;                                                    // collisions of this group inside the interval
        bool over = false;                                            // the group ends with a distance beyond maxt
        float tt = t;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            tk[k] = 0.f; xik[k] = 0.f;
            if (!over) {
                float u = pcg_f32(r);
                tt = tt + ((-spec_log(1.f - u)) * inv_mu);
                if (!(tt <= maxt)) over = true;
                else {
                    xik[k] = pcg_f32(r);
                    tk[k] = tt;
                    Tri T = tri_setup(S, G, ray_at(o, d, tt));
                    sigma1_issue(G, T, bk[k]);
                    n = k + 1;
                }
            }
        }
        int used = 0;                                                 // collisions consumed before the walk stopped
        bool dead = false;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (k < n && !dead) {
                t = tk[k];
                wsum = wsum + w;
                if (xik[k] * wsum <= w) sel_t = t;
                Tri T = tri_setup(S, G, ray_at(o, d, t));
                float st = sigma1_finish(S, T, bk[k]);
                cnt.coll++;
                w = w * ((mu - st) * inv_mu);
                used = k + 1;
                if (!(w > 0.f)) dead = true;
            }
        }
        // the sampler after `used` collisions, plus the overshooting distance draw if the walk got that far
        const int draws = 2 * used + ((!dead && over) ? 1 : 0);
        for (int q = 0; q < draws; ++q) (void) pcg_f32(alt);
        if (dead || over) break;
    }
#else
    while (active) {
        float u = pcg_f32(alt);
        t = t + ((-spec_log(1.f - u)) * inv_mu);
        if (!(t <= maxt)) break;
        wsum = wsum + w;
        float xi = pcg_f32(alt);
        if (xi * wsum <= w) sel_t = t;
        Tri T = tri_setup(S, G, ray_at(o, d, t));
        float st = lookup_sigma1(S, G, T);
        cnt.coll++;
        w = w * ((mu - st) * inv_mu);
        if (!(w > 0.f)) break;
    }
#endif
    return wsum * inv_mu;
}

// sample_recursive (:623-668): a detached radiance estimate from `p`, drawn from the secondary sampler
template <class GRID>
__device__ __forceinline__ V3 sample_recursive(const SceneK &S, const GRID &G, const GradK &D, float mu, float inv_mu, Pcg &alt,
                                               V3 ray_d, V3 p, int depth, VpsCounts &cnt) {
    V3 result = mk3(0, 0, 0);
    V3 wi = mk3(-ray_d.x, -ray_d.y, -ray_d.z);
    if (S.use_nee) {                                                  // :635-638, throughput 1
        V3 dir;
        float tr = vps_sample_emitter<false>(S, G, D, mu, inv_mu, alt, p, mk3(0, 0, 0), dir, cnt);
        float epdf;
        V3 emitted = emitter_sample_value(S, dir, epdf) * splat3(tr);
        float phase_val = phase_eval(S, wi, dir);
        result = result + ((mk3(1, 1, 1) * phase_val) * mis_weight(epdf, phase_val)) * emitted;
    }
    (void) pcg_f32(alt);                                              // :644-646
    float s2x = pcg_f32(alt), s2y = pcg_f32(alt);
    float pdf;
    V3 wo = phase_sample(S, wi, s2x, s2y, pdf);
    V3 id = rcp3(wo);
    VpsState ps;
    ps.si = box_hit(S, p, wo, id, RCV_LARGEST);                       // :651
    ps.ray_maxt = isfinite(ps.si.t) ? ps.si.t : RCV_LARGEST;          // :653-654
    ps.depth = depth + 1;
    ps.active = ps.depth < S.max_depth;
    ps.last_pdf = pdf;
    DrtOpts none; none.subsampling = none.mis = false;
    V3 Li;
    vps_trace_core<false, false, true>(S, G, D, alt, alt, p, wo, mk3(0, 0, 0), mk3(0, 0, 0), Li, cnt, ps, none);   // :656-665
    return result + Li;
}

// backpropagate_scattering_drt without a reservoir (:551-594)
template <class GRID>
__device__ __noinline__ void drt_backpropagate(const SceneK &S, const GRID &G, const GradK &D, float mu, float inv_mu, Pcg &alt,
                                               V3 o, V3 d, float ray_maxt, int depth, V3 adjoint, bool active, bool use_mis,
                                               VpsCounts &cnt) {
    V3 id = rcp3(d);
    float sel_t;
    float drt_weight = sample_interaction_drt(S, G, mu, inv_mu, alt, o, d, id, ray_maxt, active, sel_t, cnt);   // :560-561
    if (!(active && (sel_t < inf_()))) return;                        // :568
    V3 p = ray_at(o, d, sel_t);
    Tri T = tri_setup(S, G, p);
    float st = lookup_sigma1(S, G, T);
    V3 albedo = lookup_albedo(G, T, 1);
    V3 Li = sample_recursive(S, G, D, mu, inv_mu, alt, d, p, depth, cnt);   // :575-578
    float mis = use_mis ? (1.f / (1.f + st * st)) : 1.f;              // :581-585
    // backward_from(mis * drt_weight * adjoint * (sigma_t * albedo) * drt_Li) (:587-591)
    V3 c = (adjoint * (mis * drt_weight)) * Li;
    float g_st = (c.x * albedo.x + c.y * albedo.y) + c.z * albedo.z;
    V3 g_al = c * splat3(st);
    if (T.inside) {
        float sg = S.scale * g_st;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float w = tri_weight(T, k);
            uint32_t v = tri_voxel(G, T, k);
#if !RCV_GRAD_BRICKS
            red_add_f32(D.gsig + v, w * sg);
#endif
            red_add_v4(D.galb + v, w * g_al.x, w * g_al.y, w * g_al.z, 0.f);
        }
#if RCV_GRAD_BRICKS
        red_sigma_site(G, D, T, sg);
#endif
        cnt.g_sig++; cnt.g_alb++;
    }
}

// RBIntegrator.render_backward for one lane: pass 1 traces L on a clone of the sampler, pass 2 replays the stream and
// scatters the gradients.  drt_flags: 0 = use_drt False; else bit 0 use_drt, bit 1 use_drt_subsampling, bit 2 use_drt_mis.
template <class GRID>
__device__ __forceinline__ void vps_backward_lane(const SceneK &S, const GRID &G, const GradK &D, uint32_t seed, uint32_t spp,
                                                  uint64_t idx, uint32_t alt_seed, const float *grad_img, const float *film_w,
                                                  int drt_flags, VpsCounts &cnt) {
    Lane L = lane_setup(S, seed, spp, idx);
    Pcg alt; alt.state = 0; alt.inc = 1;
    V3 res, res2;
    Pcg rng1 = L.rng;                                   // sampler.clone() for pass 1
    vps_trace<false>(S, G, D, rng1, alt, L.o, L.d, mk3(0, 0, 0), mk3(0, 0, 0), res, cnt);
    V3 dL = film_gather(S, grad_img, film_w, L.fx, L.fy);
    sampler_seed(alt, alt_seed, (uint32_t) idx);        // volpathsimple.py:119-120
    if (drt_flags & 1) {
        VpsState ps; ps.active = false; ps.depth = 0; ps.si.valid = false; ps.ray_maxt = RCV_LARGEST; ps.last_pdf = 1.f;
        DrtOpts drt; drt.subsampling = (drt_flags & 2) != 0; drt.mis = (drt_flags & 4) != 0;
        vps_trace_core<true, true, false>(S, G, D, L.rng, alt, L.o, L.d, dL, res, res2, cnt, ps, drt);
    } else {
        vps_trace<true>(S, G, D, L.rng, alt, L.o, L.d, dL, res, res2, cnt);
    }
}

}  // namespace rcv
