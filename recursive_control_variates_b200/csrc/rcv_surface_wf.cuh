// rcv_surface_wf.cuh — the two-state path tracer WITH SURFACES, and the adjoint it inherits, as a per-lane event
// machine.
//
// Scenes with rectangles (diffuse walls, an `area` light: BASELINE.json "cube in a box with an area light") take this
// machine instead of rcv_twostate_wf.cuh: a path no longer ends when it leaves the medium — it may hit a wall, gather
// the light there (the shadow ray crosses the medium's null boundary, twostatevolpath.py:455-477), bounce and
// re-enter.  Same estimator, same draws, same arithmetic as the oracle's sample_twostates / sample_emitter_twostates
// (plugins/twostatevolpath.py:141-479 statement by statement), cut into the three event classes of the other machines:
//
//   EV_COLL   one tentative collision inside the medium, on the main path (:180-266) or on a shadow ray (:408-452)
//   EV_SCAT   a real scattering vertex (:259-344): albedo gather, phase sampling, emitter sample
//   EV_SLOW   the main path outside the medium: the surface block of an iteration (:283-380: intersection, emitter hit,
//             emitter sampling at a diffuse surface, BSDF sampling, medium transition), the end of the path
//   EV_NEE    a shadow ray outside the medium: the surface part of its iterations (:455-477) and its end (the
//             contribution :336-344, then the BSDF sampling of the vertex it was cast from)
//
// ADJ = true (MODE 2 only) is upstream mi.ad.integrators.prbvolpath.PRBVolpathIntegrator.sample driven by
// RBIntegrator.render_backward — the adjoint TwoStatePRBVolpathIntegrator inherits (twostatevolpath.py:11) and every
// scene without an explicit `adjoint_integrator` uses (cv_integrator.py:87-94): the lane runs the path twice on the same
// stream.  Pass 1 is the primal estimate L; pass 2 starts from L, subtracts every contribution as it is met and, at each
// medium event, scatters dL * L * d log(weight): -1 / sigma_n at a null collision, 1 / sigma_t and 1 / albedo at a real
// one; a shadow ray is walked a second time from the saved sampler state (upstream's cloned sampler) to scatter
// dL * contribution * d log(sigma_n / majorant) at its collisions.  [upstream source is not on disk: restated, unpinned]
//
// The handlers contain no warp intrinsics and compile for the host (RCV_HOST_EMU).
#pragma once
#include "rcv_twostate_wf.cuh"
#include "rcv_volpathsimple.cuh"

namespace rcv {

#define RCV_INV_PI 0.31830988618379067154f
// a fourth event class: the shadow ray's work outside the medium (its surface iterations, its end).  Kept apart from
// EV_SLOW so that a warp runs the main path's surface blocks and the shadow rays' each with their own lanes.
enum : int { EV_NEE = 4 };
// consecutive pieces of outside-the-medium work one handler call chains (1 = one piece per scheduling round).  r2aa, c1box:
// two-state / adjoint 18.0 / 58.7 ms with 1, 20.5 / 59.2 with 2, 20.4 / 57.6 with 4 — lanes that run ahead of their
// warp only desynchronise it further.
#ifndef RCV_SF_CHAIN
#define RCV_SF_CHAIN 1
#endif

enum : uint32_t {
    SF_NEEDS_ISECT = 1u, SF_SPECULAR = 2u, SF_IN_NEE = 4u, SF_SI_VALID = 8u,
    SF_MEDIUM = 16u,         // the main path is inside the medium
    SF_ESC = 32u,            // the main path left the medium in this iteration: its surface block follows (si is known)
    SF_NEE_MEDIUM = 64u,     // the shadow ray is inside the medium
    SF_NEE_ESC = 128u,       // the shadow ray left the medium in this iteration: the surface part follows
    SF_NEE_SURF = 256u,      // the shadow ray started at a diffuse surface (else at a medium vertex)
    SF_NEE_FACING = 512u,    // area light: the sampled point emits towards the reference point
    SF_POST_NEE = 1024u,     // the shadow ray ended
    SF_FIN = 2048u,          // the path ended
    SF_PASS2 = 4096u,        // adjoint: this is the second walk
    SF_REPLAY = 8192u,       // adjoint: the shadow ray is being walked again, with gradients
    SF_BSDF = 16384u         // the BSDF sampling of the surface vertex is pending (after its shadow ray)
};

struct SfLane {
    int ev, pix;
    uint32_t idx;
    float fx, fy;
    Pcg rng;
    V3 o, d;                 // main ray; at a surface vertex o = the hit point and d = the incident direction
    V3 vn;                   // ... its normal and shape (the shadow ray cast there has intersections of its own)
    int vshape;
    V3 id;                   // 1 / direction of the CURRENT segment (main ray or shadow ray)
    V3 thr_old, thr_new, L_old, L_new;
    float si_t;              // si: distance, point, normal, shape (-1 the medium's box, k rectangle k)
    V3 sp, sn;
    int shape;
    float last_pdf;
    V3 last_p;               // last_scatter_event.p
    int depth;
    uint32_t flags;
    // shadow ray (valid while SF_IN_NEE); tr_old / tr_new double as the stash for (sigma_t old, new) between EV_COLL and EV_SCAT
    V3 no, nd;
    float total, tr_old, tr_new;
    float ndist, epdf;       // ds.dist, ds.pdf
    float nw;                // phase function value (medium vertex) or cos(theta_o) (diffuse surface, 0 when the BSDF is 0)
    float ratio;
    // adjoint
    V3 dL, adjv;             // d loss / d radiance of the sample; dL * contribution of the shadow ray being replayed
    uint64_t saved;          // sampler state before the shadow ray's draws
};

struct SfSink {
    float4 *acc_old, *acc_new;
    GradK D;
    const float *grad_img, *film_w;
    uint32_t seed, spp;
    VpsCounts cnt;           // coll, scat: the tracer's events; g_sig, g_alb: gradient sites
    unsigned n_paths;
};

struct SHit {
    bool valid;
    float t;
    V3 p, n;
    int shape;
};
__device__ __forceinline__ V3 v3of(const float *a) { return mk3(a[0], a[1], a[2]); }

// scene.ray_intersect: the nearest of the medium's box and the rectangles, 0 < t <= maxt.  Out of line: three call
// sites, and the kernels' code size matters (instruction-cache misses were the top stall of the first version).
static __device__ __noinline__ SHit scene_hit_all(const SceneK &S, V3 o, V3 d, V3 id, float maxt) {
    Hit h = box_hit(S, o, d, id, maxt);
    SHit r; r.valid = h.valid; r.t = h.t; r.p = h.p; r.n = h.n; r.shape = -1;
    for (int k = 0; k < S.n_quads; ++k) {
        const QuadK &q = S.quads[k];
        V3 n = v3of(q.n), p0 = v3of(q.p0);
        float denom = dot3(n, d);
        if (denom == 0.f) continue;
        float t = dot3(n, p0 - o) / denom;
        if (!(t > 0.f) || !(t <= maxt) || !(t < r.t)) continue;
        V3 p = ray_at(o, d, t);
        V3 rel = p - p0;
        float a = dot3(rel, v3of(q.eu)) * q.inv_eu2, b = dot3(rel, v3of(q.ev)) * q.inv_ev2;
        if (!(a >= 0.f && a <= 1.f && b >= 0.f && b <= 1.f)) continue;
        r.valid = true; r.t = t; r.p = p; r.n = n; r.shape = k;
    }
    return r;
}
__device__ __forceinline__ void sf_store_hit(SfLane &ln, const SHit &h) {
    ln.si_t = h.t; ln.sp = h.p; ln.sn = h.n; ln.shape = h.shape;
    ln.flags = (ln.flags & ~(SF_NEEDS_ISECT | SF_SI_VALID)) | (h.valid ? (uint32_t) SF_SI_VALID : 0u);
}

// warp.square_to_cosine_hemisphere (concentric disk)
__device__ __forceinline__ V3 cosine_hemisphere(float sx, float sy) {
    float x = __fmaf_rn(2.f, sx, -1.f), y = __fmaf_rn(2.f, sy, -1.f);
    bool is_zero = (x == 0.f) && (y == 0.f), q13 = fabsf(x) < fabsf(y);
    float r = q13 ? y : x, rp = q13 ? x : y;
    float phi = (0.25f * RCV_PI) * (rp / r);
    if (q13) phi = (0.5f * RCV_PI) - phi;
    if (is_zero) phi = 0.f;
    float sn, cs;
    spec_sincos(fabsf(phi), sn, cs);
    if (phi < 0.f) sn = -sn;
    float px = r * cs, py = r * sn;
    float z = sqrtf(fmax_(1.f - __fmaf_rn(px, px, py * py), 0.f));
    return mk3(px, py, z);
}

// area.cpp / shape.cpp: sample_direction of the rectangle light from ref_p
__device__ __forceinline__ void area_sample(const SceneK &S, V3 ref_p, float s2x, float s2y, V3 &d, float &dist, float &pdf, bool &facing) {
    const QuadK &q = S.quads[S.area_quad];
    V3 p = mk3(__fmaf_rn(s2y, q.ev[0], __fmaf_rn(s2x, q.eu[0], q.p0[0])), __fmaf_rn(s2y, q.ev[1], __fmaf_rn(s2x, q.eu[1], q.p0[1])),
               __fmaf_rn(s2y, q.ev[2], __fmaf_rn(s2x, q.eu[2], q.p0[2])));
    V3 v = p - ref_p;
    float dist2 = dot3(v, v);
    dist = sqrtf(dist2);
    d = mk3(v.x / dist, v.y / dist, v.z / dist);
    float dp = dot3(d, v3of(q.n));
    float x = dist2 / fabsf(dp);
    pdf = isfinite(x) ? q.inv_area * x : 0.f;
    facing = (dp < 0.f) && (pdf != 0.f);
}
// pdf_direction of the light for a ray from ref_p that hit it at hit_p
__device__ __forceinline__ float area_pdf(const SceneK &S, V3 ref_p, V3 hit_p) {
    const QuadK &q = S.quads[S.area_quad];
    V3 v = hit_p - ref_p;
    float dist2 = dot3(v, v), dist = sqrtf(dist2);
    V3 d = mk3(v.x / dist, v.y / dist, v.z / dist);
    float dp = dot3(d, v3of(q.n));
    float x = dist2 / fabsf(dp);
    return ((dp < 0.f) && isfinite(x)) ? q.inv_area * x : 0.f;
}

// loop head of the main path (:182-188); false = the path ends here
__device__ __forceinline__ bool sf_loop_head(const SceneK &S, SfLane &ln) {
    if (!(any_nz(ln.thr_new) || any_nz(ln.thr_old))) return false;
    float q = fmin_(max3(vmax3(ln.thr_new, ln.thr_old)), 0.99f);
    bool perform_rr = ln.depth > S.rr_depth;
    float rr = pcg_f32(ln.rng);
    if (perform_rr) {
        if (!(rr < q)) return false;
        float rq = 1.f / q;
        ln.thr_new = ln.thr_new * rq; ln.thr_old = ln.thr_old * rq;
    }
    return true;
}

// start of a path (:150-173): the first iteration runs outside the medium (medium = None)
__device__ __forceinline__ void sf_begin(const SceneK &S, SfLane &ln, uint32_t seed, uint32_t spp, uint32_t idx, uint32_t keep_flags) {
    Lane L = lane_setup(S, seed, spp, idx);
    ln.idx = idx;
    ln.rng = L.rng; ln.o = L.o; ln.d = L.d; ln.pix = L.pix; ln.fx = L.fx; ln.fy = L.fy;
    ln.id = rcp3(ln.d);
    ln.depth = 0;
    ln.L_old = mk3(0, 0, 0); ln.L_new = mk3(0, 0, 0);
    ln.thr_old = mk3(1, 1, 1); ln.thr_new = mk3(1, 1, 1);
    ln.si_t = 0.f; ln.sp = mk3(0, 0, 0); ln.sn = mk3(0, 0, 0); ln.shape = -1;
    ln.vn = mk3(0, 0, 0); ln.vshape = -1;
    ln.last_pdf = 1.f; ln.last_p = mk3(0, 0, 0);
    ln.flags = SF_NEEDS_ISECT | SF_SPECULAR | keep_flags;
    (void) pcg_f32(ln.rng);                                           // :173 channel draw
    ln.ev = EV_SLOW;
}

// (re)start the shadow ray of the current vertex (:399-407): from a medium vertex (ln.o, no offset) or a surface point
__device__ __forceinline__ void sf_nee_start(SfLane &ln) {
    if (ln.flags & SF_NEE_SURF) {
        ln.no = offset_p(ln.o, ln.vn, ln.nd);
        ln.flags = (ln.flags & ~(SF_NEE_MEDIUM | SF_NEE_ESC)) | SF_IN_NEE | SF_NEEDS_ISECT;
        ln.ev = EV_NEE;
    } else {
        ln.no = ln.o;
        ln.flags = (ln.flags & ~SF_NEE_ESC) | SF_IN_NEE | SF_NEE_MEDIUM | SF_NEEDS_ISECT;
        ln.ev = EV_COLL;
    }
    ln.id = rcp3(ln.nd);
    ln.total = 0.f; ln.tr_old = 1.f; ln.tr_new = 1.f;
}

// scene.sample_emitter_direction (:393-397) from ref_p; false: the sample is invalid (pdf == 0), no shadow ray
__device__ __forceinline__ bool sf_nee_sample(const SceneK &S, SfLane &ln, V3 ref_p) {
    float e2x = pcg_f32(ln.rng), e2y = pcg_f32(ln.rng);               // :393
    ln.flags &= ~SF_NEE_FACING;
    if (S.emitter == 2) {
        bool facing;
        area_sample(S, ref_p, e2x, e2y, ln.nd, ln.ndist, ln.epdf, facing);
        if (facing) ln.flags |= SF_NEE_FACING;
    } else {
        ln.nd = emitter_sample_dir(S, e2x, e2y);
        ln.ndist = S.env_dist;
        (void) emitter_sample_value(S, ln.nd, ln.epdf);
    }
    return ln.epdf != 0.f;
}

template <int MODE, bool ADJ>
__device__ __forceinline__ void sf_finish(const SceneK &S, SfLane &ln, SfSink &sk) {
    sk.n_paths++;
    if (ADJ && !(ln.flags & SF_PASS2)) {
        // RBIntegrator.render_backward: dL = d(develop o put)/dL * grad_in of this sample, then the second walk on the
        // same stream starting from the primal result
        V3 Ltot = ln.L_new;
        ln.dL = film_gather(S, sk.grad_img, sk.film_w, ln.fx, ln.fy);
        sf_begin(S, ln, sk.seed, sk.spp, ln.idx, SF_PASS2);
        ln.L_new = Ltot;
        return;
    }
    if (!ADJ) {
        if (S.rfilter == 0) {
            if (ln.pix >= 0) {
                if (MODE != 2) red_add_v4(sk.acc_old + ln.pix, ln.L_old.x, ln.L_old.y, ln.L_old.z, 1.f);
                if (MODE != 1) red_add_v4(sk.acc_new + ln.pix, ln.L_new.x, ln.L_new.y, ln.L_new.z, 1.f);
            }
        } else {
            if (MODE != 2) film_splat_filtered(S, sk.acc_old, ln.fx, ln.fy, ln.L_old.x, ln.L_old.y, ln.L_old.z);
            if (MODE != 1) film_splat_filtered(S, sk.acc_new, ln.fx, ln.fy, ln.L_new.x, ln.L_new.y, ln.L_new.z);
        }
    }
    ln.flags = 0;
    ln.ev = EV_INIT;
}

// a contribution to the radiance: added in the primal walk, subtracted from what is still to come in the adjoint's second
template <bool ADJ>
__device__ __forceinline__ void sf_add(SfLane &ln, V3 c_old, V3 c_new) {
    if (ADJ && (ln.flags & SF_PASS2)) ln.L_new = ln.L_new - c_new;
    else ln.L_new = ln.L_new + c_new;
    ln.L_old = ln.L_old + c_old;
}

// ---------------------------------------------------------------------------
// EV_COLL: one tentative collision inside the medium
// ---------------------------------------------------------------------------
template <int MODE, bool ADJ, class GRID>
__device__ __forceinline__ void sf_collide(const SceneK &S, const GRID &G, float mu_vol, float inv_mu_vol, SfLane &ln, SfSink &sk) {
    const bool nee = (ln.flags & SF_IN_NEE) != 0;
    V3 so, sd; float maxt;
    if (!nee) {
        if (!sf_loop_head(S, ln)) { ln.flags |= SF_FIN; ln.ev = EV_SLOW; return; }
        so = ln.o; sd = ln.d; maxt = RCV_LARGEST;
    } else {
        float remaining = ln.ndist * (1.f - RCV_SHADOW_EPS) - ln.total;   // :409
        if (!(remaining > 0.f)) { ln.flags |= SF_POST_NEE; ln.ev = EV_NEE; return; }   // :411
        so = ln.no; sd = ln.nd; maxt = remaining;
    }
    float u = pcg_f32(ln.rng);                                        // :194 / :422
    Mei2 m; Tri T;
    sample_interaction2<MODE>(S, G, mu_vol, inv_mu_vol, so, sd, ln.id, maxt, u, m, T, sk.cnt.coll);
    const float mu = m.mu, inv_mu = m.inv_mu;
    if (ln.flags & SF_NEEDS_ISECT) sf_store_hit(ln, scene_hit_all(S, so, sd, ln.id, maxt));   // :200-204 / :414-416
    if (ln.si_t < m.t) m.t = inf_();                                  // :205-206 / :423-424
    if (m.t < inf_()) {
        if (nee) {                                                    // :444-452, :472-473
            ln.no = m.p;
            ln.si_t = ln.si_t - m.t;
            float sn_o = mu - m.so, sn_n = mu - m.sn;
            float trm_o = (mu != 0.f) ? (sn_o * inv_mu) : sn_o, trm_n = (mu != 0.f) ? (sn_n * inv_mu) : sn_n;
            if (ADJ && (ln.flags & SF_REPLAY) && (trm_n > 0.f)) {     // prbvolpath sample_emitter, Backward
                float k = (-inv_mu) / trm_n, g = 0.f;
                g += ln.adjv.x * k; g += ln.adjv.y * k; g += ln.adjv.z * k;
                scatter_sigma_t_T(G, sk.D, T, S.scale * g, sk.cnt);
            }
            ln.tr_old = ln.tr_old * trm_o;
            ln.tr_new = ln.tr_new * trm_n;
            if (!(ln.tr_new != 0.f || ln.tr_old != 0.f)) { ln.flags |= SF_POST_NEE; ln.ev = EV_NEE; return; }
            ln.total += m.t;
            return;
        }
        float sp_new = m.sn * inv_mu, sp_old = m.so * inv_mu;         // :223-225
        float prob = (mu != 0.f) ? (sp_new + sp_old) * 0.5f : 0.f;
        float xi = pcg_f32(ln.rng);                                   // :227
        if (xi >= prob) {                                             // null collision
            float sn_n = mu - m.sn, sn_o = mu - m.so;
            float rden = 1.f / (sn_n + sn_o);
            float w_new = (sn_n * 2.f) * rden, w_old = (sn_o * 2.f) * rden;
            ln.o = m.p;
            ln.si_t = ln.si_t - m.t;
            ln.thr_new = ln.thr_new * w_new; ln.thr_old = ln.thr_old * w_old;
            if (ADJ && (ln.flags & SF_PASS2) && (sn_n > 0.f)) {       // dL * L * d log(sigma_n) / d sigma_t
                float aw = (ln.dL.x * ln.L_new.x + ln.dL.y * ln.L_new.y) + ln.dL.z * ln.L_new.z;
                scatter_sigma_t_T(G, sk.D, T, S.scale * (-(aw * (1.f / sn_n))), sk.cnt);
            }
            return;
        }
        ln.depth += 1;                                                // :249
        ln.last_p = m.p;                                              // :250
        if (!(ln.depth < S.max_depth)) { ln.flags |= SF_FIN; ln.ev = EV_SLOW; return; }   // :253
        ln.o = m.p;                                                   // the vertex; the new direction is chosen in EV_SCAT
        ln.tr_old = m.so; ln.tr_new = m.sn;
        ln.ev = EV_SCAT;
        return;
    }
    // the segment left the medium: the surface part of this iteration follows, with the si found above
    ln.flags |= nee ? SF_NEE_ESC : SF_ESC;
    ln.ev = nee ? EV_NEE : EV_SLOW;
}

// ---------------------------------------------------------------------------
// EV_SCAT: a real scattering vertex
// ---------------------------------------------------------------------------
template <int MODE, bool ADJ>
__device__ __forceinline__ void sf_scatter(const SceneK &S, const GridK &G, SfLane &ln, SfSink &sk) {
    sk.cnt.scat++;
    const float st_old = ln.tr_old, st_new = ln.tr_new;
    Tri T = tri_setup(S, G, ln.o);
    V3 a_new, a_old;
    if (MODE == 0) lookup_albedo2(G, T, a_old, a_new);
    else { a_new = lookup_albedo(G, T, MODE == 1 ? 0 : 1); a_old = a_new; }
    V3 ss_new = splat3(st_new) * a_new, ss_old = splat3(st_old) * a_old;
    float rden = 1.f / (st_new + st_old);                             // :259-260
    ln.thr_new = ln.thr_new * ((ss_new * 2.f) * rden);
    ln.thr_old = ln.thr_old * ((ss_old * 2.f) * rden);
    if (ADJ && (ln.flags & SF_PASS2) && T.inside) {                   // dL * L * d log(sigma_t * albedo)
        float aw = (ln.dL.x * ln.L_new.x + ln.dL.y * ln.L_new.y) + ln.dL.z * ln.L_new.z;
        V3 Lo = ln.L_new / vmax3(splat3(1e-8f), a_new);
        V3 g_al = ln.dL * Lo;
        float sg = S.scale * (aw * (1.f / st_new));
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float w = tri_weight(T, k);
            uint32_t v = tri_voxel(G, T, k);
#if !RCV_GRAD_BRICKS
            red_add_f32(sk.D.gsig + v, w * sg);
#endif
            red_add_v4(sk.D.galb + v, w * g_al.x, w * g_al.y, w * g_al.z, 0.f);
        }
#if RCV_GRAD_BRICKS
        red_sigma_site(G, sk.D, T, sg);
#endif
        sk.cnt.g_sig++; sk.cnt.g_alb++;
    }
    (void) pcg_f32(ln.rng);                                           // :275
    float s2x = pcg_f32(ln.rng), s2y = pcg_f32(ln.rng);
    V3 wi = mk3(-ln.d.x, -ln.d.y, -ln.d.z);
    float pdf;
    V3 wo = phase_sample(S, wi, s2x, s2y, pdf);
    ln.d = wo;                                                        // :278-281 (pdf > 0 always holds)
    ln.flags |= SF_NEEDS_ISECT;
    ln.last_pdf = pdf;
    float ratio = phase_eval(S, wi, wo) / pdf;                        // :360-361
    ln.ev = EV_COLL;
    if (S.use_nee) {                                                  // :315-325
        ln.flags &= ~SF_SPECULAR;
        if (!S.sample_emitters) ln.flags |= SF_SPECULAR;
        if (S.sample_emitters) {
            ln.saved = ln.rng.state;                                  // prbvolpath: sampler.clone()
            ln.flags &= ~SF_NEE_SURF;
            if (sf_nee_sample(S, ln, ln.o)) {
                ln.nw = phase_eval(S, wi, ln.nd);
                ln.ratio = ratio;
                sf_nee_start(ln);
                return;
            }
        }
    }
    ln.thr_old = ln.thr_old * ratio;
    ln.id = rcp3(ln.d);
}

// ---------------------------------------------------------------------------
// EV_SLOW pieces
// ---------------------------------------------------------------------------
// BSDF sampling and the end of the iteration (:346-380) at the surface point (sp, sn, shape), incident direction d
__device__ __forceinline__ void sf_bsdf(const SceneK &S, SfLane &ln) {
    ln.flags &= ~SF_BSDF;
    (void) pcg_f32(ln.rng);                                           // :347
    float b2x = pcg_f32(ln.rng), b2y = pcg_f32(ln.rng);               // :348
    if (ln.vshape < 0) {
        // `null`: wo = -wi, weight 1, a Delta lobe; the medium changes (:378-379)
        const bool leaving = dot3(ln.vn, ln.d) > 0.f;
        ln.o = offset_p(ln.o, ln.vn, ln.d);
        ln.flags |= SF_NEEDS_ISECT;
        if (leaving) { ln.flags &= ~SF_MEDIUM; ln.ev = EV_SLOW; }
        else { ln.flags |= SF_MEDIUM; ln.ev = EV_COLL; }
        return;
    }
    const QuadK &q = S.quads[ln.vshape];
    const V3 n = v3of(q.n), refl = v3of(q.refl);
    float ci = -dot3(n, ln.d);
    V3 wl = cosine_hemisphere(b2x, b2y);
    float pdf = RCV_INV_PI * wl.z;
    if (!((ci > 0.f) && (pdf > 0.f))) { ln.flags |= SF_FIN; ln.ev = EV_SLOW; return; }   // :349 bs.pdf == 0: the lane ends (:380)
    V3 val_old = (refl * RCV_INV_PI) * wl.z;                          // :351 eval_old(bs.wo)
    ln.thr_old = ln.thr_old * mk3(val_old.x / pdf, val_old.y / pdf, val_old.z / pdf);   // :354-355
    ln.thr_new = ln.thr_new * refl;                                   // :357
    V3 wo = mk3(__fmaf_rn(q.s[0], wl.x, __fmaf_rn(q.t[0], wl.y, q.n[0] * wl.z)),
                __fmaf_rn(q.s[1], wl.x, __fmaf_rn(q.t[1], wl.y, q.n[1] * wl.z)),
                __fmaf_rn(q.s[2], wl.x, __fmaf_rn(q.t[2], wl.y, q.n[2] * wl.z)));
    ln.last_p = ln.o;                                                 // :368-377
    ln.o = offset_p(ln.o, ln.vn, wo);
    ln.d = wo; ln.id = rcp3(wo);
    ln.depth += 1;
    ln.last_pdf = pdf;
    ln.flags = (ln.flags & ~SF_SPECULAR) | SF_NEEDS_ISECT;
    ln.ev = EV_SLOW;
}

// the surface block of a main-path iteration (:283-344); entered at the start of an iteration outside the medium, or
// with SF_ESC right after the medium part of the iteration
template <bool ADJ>
__device__ __forceinline__ void sf_surface(const SceneK &S, SfLane &ln) {
    bool last = false;
    if (ln.flags & SF_ESC) {
        ln.flags &= ~SF_ESC;
    } else {
        if (!sf_loop_head(S, ln)) { ln.flags |= SF_FIN; return; }
        last = !(ln.depth < S.max_depth);                             // :253 switches the lane off, the block below still runs
        sf_store_hit(ln, scene_hit_all(S, ln.o, ln.d, ln.id, RCV_LARGEST));   // :285-286
    }
    const bool si_valid = (ln.flags & SF_SI_VALID) != 0;
    // ---- intersection with emitters (:288-308)
    const bool has_emitter = si_valid ? (S.emitter == 2 && ln.shape == S.area_quad) : (S.emitter != 2);
    if (has_emitter && !((ln.depth == 0) && S.hide_emitters)) {
        float epdf; V3 emitted;
        if (si_valid) {
            epdf = area_pdf(S, ln.last_p, ln.sp);
            emitted = (-dot3(ln.sn, ln.d) > 0.f) ? v3of(S.quads[ln.shape].rad) : mk3(0, 0, 0);
        } else {
            emitted = emitter_eval(S, ln.d, epdf);
        }
        if (!S.use_nee) epdf = 0.f;                                   // :296-299
        bool count_direct = (ln.depth == 0) || (ln.flags & SF_SPECULAR);
        if (count_direct) sf_add<ADJ>(ln, ln.thr_old * emitted, ln.thr_new * emitted);
        else {
            float mis_bsdf = mis_weight(ln.last_pdf, epdf);
            sf_add<ADJ>(ln, (ln.thr_old * mis_bsdf) * emitted, (ln.thr_new * mis_bsdf) * emitted);
        }
    }
    if (!si_valid || last) { ln.flags |= SF_FIN; return; }            // :310, :380
    ln.o = ln.sp; ln.vn = ln.sn; ln.vshape = ln.shape;                // the surface vertex
    // ---- emitter sampling at a diffuse surface (:315-344)
    if (S.use_nee && ln.vshape >= 0 && (ln.depth + 1 < S.max_depth)) {
        ln.saved = ln.rng.state;
        ln.flags |= SF_NEE_SURF;
        if (sf_nee_sample(S, ln, ln.o)) {
            const V3 n = ln.vn;
            float ci = -dot3(n, ln.d), co = dot3(n, ln.nd);           // diffuse.cpp eval_pdf
            ln.nw = (ci > 0.f && co > 0.f) ? co : 0.f;
            sf_nee_start(ln);
            return;
        }
    }
    ln.flags |= SF_BSDF;
}

// the surface part of a shadow-ray iteration (:455-477) — at the start of an iteration outside the medium, or with
// SF_NEE_ESC right after its medium part
__device__ __forceinline__ void sf_nee_iter(const SceneK &S, SfLane &ln) {
    if (ln.flags & SF_NEE_ESC) {
        ln.flags &= ~SF_NEE_ESC;
    } else {
        float remaining = ln.ndist * (1.f - RCV_SHADOW_EPS) - ln.total;   // :409
        if (!(remaining > 0.f)) { ln.flags |= SF_POST_NEE; return; }
        sf_store_hit(ln, scene_hit_all(S, ln.no, ln.nd, ln.id, remaining));
    }
    if (!(ln.flags & SF_SI_VALID)) { ln.flags |= SF_POST_NEE; return; }   // nothing in the way
    if (ln.shape >= 0) {                                              // a wall: eval_null_transmission == 0 (:457-460)
        ln.tr_old = ln.tr_old * 0.f; ln.tr_new = ln.tr_new * 0.f;
        ln.flags |= SF_POST_NEE;
        return;
    }
    ln.no = offset_p(ln.sp, ln.sn, ln.nd);                            // the null boundary (:466-467)
    ln.flags |= SF_NEEDS_ISECT;
    ln.total += ln.si_t;                                              // :473
    if (dot3(ln.sn, ln.nd) > 0.f) ln.flags &= ~SF_NEE_MEDIUM;         // :476-477
    else { ln.flags |= SF_NEE_MEDIUM; ln.ev = EV_COLL; }
}

// the shadow ray ended: its contribution (:336-344); in the adjoint's second walk, walk it again with gradients; then the
// iteration continues (:346-380)
template <bool ADJ>
__device__ __forceinline__ void sf_nee_done(const SceneK &S, SfLane &ln) {
    ln.flags &= ~SF_POST_NEE;
    if (!(ADJ && (ln.flags & SF_REPLAY))) {
        V3 emitter_val;
        if (S.emitter == 2) {
            const V3 rad = v3of(S.quads[S.area_quad].rad);
            emitter_val = (ln.flags & SF_NEE_FACING) ? mk3(rad.x / ln.epdf, rad.y / ln.epdf, rad.z / ln.epdf) : mk3(0, 0, 0);
        } else {
            float e;
            emitter_val = emitter_sample_value(S, ln.nd, e);
        }
        V3 w3; float dir_pdf;
        if (ln.flags & SF_NEE_SURF) {
            w3 = (v3of(S.quads[ln.vshape].refl) * RCV_INV_PI) * ln.nw;
            dir_pdf = RCV_INV_PI * ln.nw;
        } else { w3 = splat3(ln.nw); dir_pdf = ln.nw; }
        float mis_em = mis_weight(ln.epdf, dir_pdf);
        V3 c_new = ((ln.thr_new * w3) * mis_em) * (emitter_val * splat3(ln.tr_new));
        V3 c_old = ((ln.thr_old * w3) * mis_em) * (emitter_val * splat3(ln.tr_old));
        sf_add<ADJ>(ln, c_old, c_new);
        if (ADJ && (ln.flags & SF_PASS2)) {
            ln.adjv = ln.dL * c_new;
            ln.rng.state = ln.saved;
            (void) pcg_u32(ln.rng); (void) pcg_u32(ln.rng);           // the direction sample, same values
            ln.flags |= SF_REPLAY;
            sf_nee_start(ln);
            return;
        }
    }
    ln.flags &= ~(SF_REPLAY | SF_IN_NEE | SF_NEE_MEDIUM);
    if (ln.flags & SF_NEE_SURF) { ln.flags |= SF_BSDF; ln.ev = EV_SLOW; return; }
    ln.thr_old = ln.thr_old * ln.ratio;                               // :360-361
    ln.flags |= SF_NEEDS_ISECT;
    ln.id = rcp3(ln.d);
    ln.ev = EV_COLL;
}

// EV_SLOW: the main path outside the medium
template <int MODE, bool ADJ>
__device__ __forceinline__ void sf_slow(const SceneK &S, SfLane &ln, SfSink &sk) {
    // one copy of each piece: a surface block flows into its BSDF sampling and, where the path ends, into the finish; a lane
    // that stays outside the medium goes on into the surface block of its next iteration in the same round
#pragma unroll 1
    for (int pass = 0; pass < RCV_SF_CHAIN; ++pass) {
        if (!(ln.flags & (SF_FIN | SF_BSDF))) sf_surface<ADJ>(S, ln);
        if (ln.flags & SF_BSDF) sf_bsdf(S, ln);
        if ((ln.flags & SF_FIN) || ln.ev != EV_SLOW) break;
    }
    if (ln.flags & SF_FIN) sf_finish<MODE, ADJ>(S, ln, sk);
}
// EV_NEE: the shadow ray outside the medium
template <bool ADJ>
__device__ __forceinline__ void sf_nee(const SceneK &S, SfLane &ln) {
#pragma unroll 1
    for (int pass = 0; pass < RCV_SF_CHAIN; ++pass) {
        if (!(ln.flags & SF_POST_NEE)) sf_nee_iter(S, ln);
        if (ln.flags & SF_POST_NEE) { sf_nee_done<ADJ>(S, ln); break; }
        if (ln.ev != EV_NEE) break;
    }
}

}  // namespace rcv
