// rcv_twostate.cuh — per-lane two-state volumetric path tracer (device code),
// straight-line version used by the SIMPLE / BATCH schedules.
//
// Restates, for one lane and with the Dr.Jit masks resolved into control
// flow, TwoStatePRBVolpathIntegrator.sample_twostates
// (plugins/twostatevolpath.py:141-385), sample_emitter_twostates (:387-479)
// and TwoStateMedium.sample_interaction_twostates (plugins/twostatemedium.py:
// 15-107) for the supported scene class (one null-BSDF box with a
// heterogeneous medium, `constant` environment emitter).  The random-number
// stream of a lane is consumed in exactly the reference's order
// (SURVEY.md Appendix A).
//
// MODE 0: two states (sig2.x = theta_{i-1}, sig2.y = theta_i), shared stream.
// MODE 1/2: both "states" are the old/new one -> the estimator reduces to the
// ordinary single-state null-scattering path tracer (baseline `prbvolpath`).
//
// Arithmetic forms fixed by this build (mirrored by the oracle):
// divisions by the majorant are multiplications by its IEEE reciprocal, the
// two weights that share a denominator use one IEEE reciprocal, and 1/d of a
// ray is computed once per ray.
#pragma once
#include "rcv_device.cuh"

namespace rcv {

struct Mei2 {
    float t;        // inf when invalid
    V3 p;
    float so, sn;   // sigma_t of the old / new state at p
    float mu, inv_mu;   // the majorant the collision was sampled with: the volume's, or the supergrid cell's at p
};

template <int MODE>
__device__ __forceinline__ float combined_majorant(const GridK &G) {
    float mo = __ldg(G.majorants), mn = __ldg(G.majorants + 1);
    if (MODE == 0) return fmax_(mo, mn);          // twostatemedium.py:73-75
    return MODE == 1 ? mo : mn;
}

// twostatemedium.py:15-107 with a global majorant.  Only sigma_t is fetched
// here; sigma_s (albedo) is needed at real scatters only and fetched there —
// the results are identical to fetching it at every tentative collision.
template <int MODE, class GRID>
__device__ __forceinline__ void sample_interaction2(const SceneK &S, const GRID &G, float mu, float inv_mu, V3 o, V3 d, V3 id,
                                                    float ray_maxt, float u, Mei2 &m, Tri &T, unsigned &n_coll) {
    float mint, maxt;
    bool active = medium_interval(S, o, d, id, ray_maxt, mint, maxt);   // :22
    float tau = -spec_log(1.f - u);                                   // :24
    float t;
    float m_cell = 0.f;
    if constexpr (GRID::kSupergrid) t = active ? ((RCV_DDA_LIN_TWOSTATE) ? dda_sample_lin<MODE>(S, G, o, d, id, mint, maxt, tau, m_cell) : dda_sample<MODE>(S, G, o, d, id, mint, maxt, tau, m_cell)) : inf_();   // :25-70
    else t = mint + (tau * inv_mu);                                   // :78
    bool valid = active && (t <= maxt);                               // :80
    m.t = valid ? t : inf_();                                         // :88
    m.p = ray_at(o, d, t);                                            // :89
    m.so = 0.f; m.sn = 0.f;
    m.mu = mu; m.inv_mu = inv_mu;
    if (valid) {                                                      // :93-94
        T = tri_setup(S, G, m.p);
        lookup_sigma2(S, G, T, m.so, m.sn);
        n_coll++;
        if (MODE == 1) m.sn = m.so;
        if (MODE == 2) m.so = m.sn;
        if constexpr (GRID::kSupergrid) { m.mu = m_cell; m.inv_mu = 1.f / m.mu; }   // :82-86: the stopping cell's majorant
    }
}

// twostatevolpath.py:387-479: emitter direction sample + two-state ratio
// tracking towards it.  Transmittances are scalars: all three channels carry
// the same value because sigma_t is monochromatic.
template <int MODE, class GRID>
__device__ __forceinline__ void nee_twostates(const SceneK &S, const GRID &G, float mu, float inv_mu, Pcg &rng, V3 ref_p,
                                              float &tr_old, float &tr_new, V3 &dir, unsigned &n_coll) {
    float s2x = pcg_f32(rng), s2y = pcg_f32(rng);                     // :393
    V3 d = emitter_sample_dir(S, s2x, s2y);
    dir = d;
    tr_old = 1.f; tr_new = 1.f;
    if (S.emitter == 1) {                                             // :395-397 a sample of zero density is invalid
        float epdf;
        (void) emitter_sample_value(S, d, epdf);
        if (epdf == 0.f) return;
    }
    V3 id = rcp3(d);
    V3 o = ref_p;                                                     // :399 (n = 0: no offset)
    float total_dist = 0.f;
    Hit si; si.valid = false; si.t = 0.f; si.p = mk3(0, 0, 0); si.n = mk3(0, 0, 0);
    bool needs_isect = true, in_medium = true;
    tr_old = 1.f; tr_new = 1.f;
    while (true) {                                                    // :408
        float remaining = S.env_dist * (1.f - RCV_SHADOW_EPS) - total_dist;   // :409
        if (!(remaining > 0.f)) break;                                // :411
        if (needs_isect) { si = box_hit(S, o, d, id, remaining); needs_isect = false; }   // :414-416
        if (in_medium) {
            float u = pcg_f32(rng);                                   // :422
            Mei2 m; Tri T;
            sample_interaction2<MODE>(S, G, mu, inv_mu, o, d, id, remaining, u, m, T, n_coll);
            if (si.t < m.t) m.t = inf_();                             // :423-424
            if (m.t < inf_()) {                                       // :444
                o = m.p;                                              // :445
                si.t = si.t - m.t;                                    // :446
                float sn_o = m.mu - m.so, sn_n = m.mu - m.sn;
                tr_old = tr_old * ((m.mu != 0.f) ? (sn_o * m.inv_mu) : sn_o); // :447-452, :462-463
                tr_new = tr_new * ((m.mu != 0.f) ? (sn_n * m.inv_mu) : sn_n);
                if (!(tr_new != 0.f || tr_old != 0.f)) break;         // :472
                total_dist += m.t;                                    // :473
                continue;
            }
        }
        // surface: null-BSDF boundary (transmission 1) or nothing left to hit
        if (!si.valid) break;                                         // :456
        o = offset_p(si.p, si.n, d);                                  // :466-467
        needs_isect = true;
        if (!(tr_new != 0.f || tr_old != 0.f)) break;
        total_dist += si.t;
        in_medium = !(dot3(si.n, d) > 0.f);                           // :476-477
    }
    dir = d;
}

template <int MODE, class GRID>
__device__ __forceinline__ void trace_twostates(const SceneK &S, const GRID &G, Pcg &rng, V3 o, V3 d, V3 &L_old, V3 &L_new,
                                                unsigned &n_coll, unsigned &n_scat) {
    const float mu = combined_majorant<MODE>(G);
    const float inv_mu = 1.f / mu;
    int depth = 0;
    L_old = mk3(0, 0, 0); L_new = mk3(0, 0, 0);
    V3 thr_old = mk3(1, 1, 1), thr_new = mk3(1, 1, 1);
    Hit si; si.valid = false; si.t = 0.f; si.p = mk3(0, 0, 0); si.n = mk3(0, 0, 0);
    bool needs_isect = true, in_medium = false, specular_chain = true;
    float last_pdf = 1.f;
    V3 id = rcp3(d);
    (void) pcg_f32(rng);                                              // :173 channel draw

    while (true) {                                                    // :180
        if (!(any_nz(thr_new) || any_nz(thr_old))) break;             // :182
        float q = fmin_(max3(vmax3(thr_new, thr_old)), 0.99f);        // :183 (eta == 1)
        bool perform_rr = depth > S.rr_depth;                         // :184
        float rr = pcg_f32(rng);                                      // :185
        if (perform_rr) {
            if (!(rr < q)) break;
            float rq = 1.f / q;                                       // :187-188
            thr_new = thr_new * rq; thr_old = thr_old * rq;
        }
        if (in_medium) {
            float u = pcg_f32(rng);                                   // :194
            Mei2 m; Tri T;
            sample_interaction2<MODE>(S, G, mu, inv_mu, o, d, id, RCV_LARGEST, u, m, T, n_coll);   // :196
            if (needs_isect) { si = box_hit(S, o, d, id, RCV_LARGEST); needs_isect = false; }   // :200-204
            if (si.t < m.t) m.t = inf_();                             // :205-206
            if (m.t < inf_()) {                                       // :217
                float sp_new = m.sn * m.inv_mu, sp_old = m.so * m.inv_mu; // :223-224
                float prob = (m.mu != 0.f) ? (sp_new + sp_old) * 0.5f : 0.f;   // :225
                float xi = pcg_f32(rng);                              // :227
                if (xi >= prob) {
                    // null collision (:230-231, :256-257, :265-266)
                    float sn_n = m.mu - m.sn, sn_o = m.mu - m.so;
                    float rden = 1.f / (sn_n + sn_o);
                    float w_new = (sn_n * 2.f) * rden, w_old = (sn_o * 2.f) * rden;
                    o = m.p;
                    si.t = si.t - m.t;
                    thr_new = thr_new * w_new; thr_old = thr_old * w_old;
                    continue;
                }
                // real scatter
                depth += 1;                                           // :249
                if (!(depth < S.max_depth)) break;                    // :253
                n_scat++;
                V3 a_new = (MODE == 1) ? lookup_albedo(G, T, 0) : lookup_albedo(G, T, 1);
                V3 a_old = (MODE == 0) ? lookup_albedo(G, T, 0) : a_new;
                V3 ss_new = splat3(m.sn) * a_new, ss_old = splat3(m.so) * a_old;
                float rden = 1.f / (m.sn + m.so);                     // :259-260
                thr_new = thr_new * ((ss_new * 2.f) * rden);
                thr_old = thr_old * ((ss_old * 2.f) * rden);
                (void) pcg_f32(rng);                                  // :275 sample1
                float s2x = pcg_f32(rng), s2y = pcg_f32(rng);
                V3 wi = mk3(-d.x, -d.y, -d.z);
                float pdf;
                V3 wo = phase_sample(S, wi, s2x, s2y, pdf);
                bool scattered = pdf > 0.f;                           // :276
                if (scattered) {                                      // :278-281
                    o = m.p; d = wo; id = rcp3(d);
                    needs_isect = true;
                    last_pdf = pdf;
                }
                if (S.use_nee) {                                      // :315-344
                    specular_chain = specular_chain && !scattered;
                    specular_chain = specular_chain || (scattered && !S.sample_emitters);
                    if (scattered && S.sample_emitters) {
                        float tr_old, tr_new; V3 dir;
                        nee_twostates<MODE>(S, G, mu, inv_mu, rng, m.p, tr_old, tr_new, dir, n_coll);
                        float epdf;
                        V3 emitter_val = emitter_sample_value(S, dir, epdf);
                        float phase_val = phase_eval(S, wi, dir);     // :333-334
                        float mis_em = mis_weight(epdf, phase_val);   // :341
                        L_new = L_new + ((thr_new * phase_val) * mis_em) * (emitter_val * splat3(tr_new));   // :343
                        L_old = L_old + ((thr_old * phase_val) * mis_em) * (emitter_val * splat3(tr_old));   // :344
                    }
                }
                if (scattered) {                                      // :360-361
                    float ratio = phase_eval(S, wi, wo) / pdf;
                    thr_old = thr_old * ratio;
                }
                continue;
            }
            // escaped the medium: continue with the surface part using the stored si
        } else if (needs_isect) {                                     // :285-286
            si = box_hit(S, o, d, id, RCV_LARGEST);
            needs_isect = false;
        }
        // ---- emitter hit (:288-308): the environment is "hit" when nothing else is
        if (!si.valid) {
            if (!((depth == 0) && S.hide_emitters)) {
                bool count_direct = (depth == 0) || specular_chain;
                float epdf;
                const V3 env = emitter_eval(S, d, epdf);
                if (count_direct) {
                    L_new = L_new + thr_new * env;
                    L_old = L_old + thr_old * env;
                } else {
                    float mis_bsdf = mis_weight(last_pdf, S.use_nee ? epdf : 0.f);
                    L_new = L_new + (thr_new * mis_bsdf) * env;
                    L_old = L_old + (thr_old * mis_bsdf) * env;
                }
            }
            break;                                                    // :310, :380
        }
        // ---- null BSDF crossing (:346-379): three draws, weight 1, direction unchanged
        (void) pcg_f32(rng); (void) pcg_f32(rng); (void) pcg_f32(rng);
        o = offset_p(si.p, si.n, d);
        needs_isect = true;
        in_medium = !(dot3(si.n, d) > 0.f);
    }
}

}  // namespace rcv
