// rcv_twostate_wf.cuh — the two-state path tracer as a per-lane event machine.
//
// Same estimator and same random-number stream as rcv_twostate.cuh (and as
// plugins/twostatevolpath.py:141-479), cut into pieces that a persistent warp
// can run with (nearly) all of its lanes:
//
//   ts_raygen  (its own, fully convergent kernel) start of every path:
//              sampler seed, camera ray, first loop head, boundary crossing.
//              Paths that miss the medium are finished here; the others are
//              appended as 48-byte path records to a queue in HBM.
//   ts_load    a finished lane picks up the next record (three float4 loads)
//   EV_COLL    ONE tentative collision on the lane's current segment — the
//              main path (delta tracking, :180-266) or the shadow ray of next
//              event estimation (ratio tracking, :408-477): loop head, free
//              flight, one two-state sigma_t gather, null/real/escape decision
//   EV_SCAT    a real scattering vertex: two-state albedo gather, throughput
//              weights, phase sampling, emitter direction sample (:259-344)
//   EV_SLOW    everything that happens once per path or per shadow ray and
//              would otherwise run with one or two lanes: end of the shadow
//              ray (:336-344, :360-361), boundary crossing + environment hit
//              of an escaping path (:346-379, :288-308), the film splat
//
// The handlers contain no warp intrinsics; they also compile for the host
// (RCV_HOST_EMU) where tests single-step them against the oracle.
#pragma once
#include "rcv_twostate.cuh"

namespace rcv {

enum : int { EV_INIT = 0, EV_COLL = 1, EV_SCAT = 2, EV_SLOW = 3 };
enum : uint32_t {
    F_NEEDS_ISECT = 1u, F_SPECULAR = 2u, F_IN_NEE = 4u, F_SI_VALID = 8u,
    // work parked for the batched EV_SLOW step
    F_POST_NEE = 16u,      // the shadow ray ended: add its contribution, resume the main path
    F_FIN = 32u,           // the path ended: splat it
    F_FIN_ESCAPE = 64u     // ... after leaving the medium through the boundary (:346-379, :288-308)
};

struct TsLane {
    int ev;
    int pix;                // film pixel of the path (-1: outside the film)
    float fx, fy;           // film position (the reconstruction filter's footprint is centred here)
    Pcg rng;
    V3 o, d;                // main-path ray
    V3 id;                  // 1/d of the CURRENT segment (main ray, or the shadow ray during NEE)
    V3 thr_old, thr_new, L_old, L_new;
    float si_t;             // distance to the boundary along the current segment
    float last_pdf;
    int depth;
    uint32_t flags;
    // shadow ray of next event estimation (valid while F_IN_NEE); tr_old/tr_new
    // double as the stash for (sigma_t old, new) between EV_COLL and EV_SCAT
    V3 no, nd;
    float total, tr_old, tr_new, phase_val, ratio;
};

struct TsSink {
    float4 *acc_old, *acc_new;
    unsigned n_coll, n_scat, n_paths;
    bool defer = false;     // ray generation: leave the film splat of a finished path to the kernel (warp-aggregated there)
};

template <int MODE>
__device__ __forceinline__ void ts_finish(const SceneK &S, TsLane &ln, TsSink &sk) {
    if (sk.defer) { sk.n_paths++; ln.ev = EV_INIT; return; }
    if (S.rfilter == 0) {                               // box filter: the pixel that contains the sample, weight 1
        if (ln.pix >= 0) {
            if (MODE != 2) red_add_v4(sk.acc_old + ln.pix, ln.L_old.x, ln.L_old.y, ln.L_old.z, 1.f);
            if (MODE != 1) red_add_v4(sk.acc_new + ln.pix, ln.L_new.x, ln.L_new.y, ln.L_new.z, 1.f);
        }
    } else {
        if (MODE != 2) film_splat_filtered(S, sk.acc_old, ln.fx, ln.fy, ln.L_old.x, ln.L_old.y, ln.L_old.z);
        if (MODE != 1) film_splat_filtered(S, sk.acc_new, ln.fx, ln.fy, ln.L_new.x, ln.L_new.y, ln.L_new.z);
    }
    sk.n_paths++;
    ln.ev = EV_INIT;
}

// environment "hit" of the main path (:288-308)
__device__ __forceinline__ void ts_env_hit(const SceneK &S, TsLane &ln) {
    if ((ln.depth == 0) && S.hide_emitters) return;
    float epdf;
    const V3 env = emitter_eval(S, ln.d, epdf);
    bool count_direct = (ln.depth == 0) || (ln.flags & F_SPECULAR);
    if (count_direct) {
        ln.L_new = ln.L_new + ln.thr_new * env;
        ln.L_old = ln.L_old + ln.thr_old * env;
    } else {
        float mis_bsdf = mis_weight(ln.last_pdf, S.use_nee ? epdf : 0.f);
        ln.L_new = ln.L_new + (ln.thr_new * mis_bsdf) * env;
        ln.L_old = ln.L_old + (ln.thr_old * mis_bsdf) * env;
    }
}

// loop head of the main path (:182-188); false = the path ends here
__device__ __forceinline__ bool ts_loop_head(const SceneK &S, TsLane &ln) {
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

// Start of a path (:150-173) and its surface iterations outside the medium
// (:283-380 with the null BSDF): the path ends in the environment (ev = EV_INIT,
// already splatted) or crosses the boundary into the medium (ev = EV_COLL).
template <int MODE>
__device__ __forceinline__ void ts_raygen(const SceneK &S, TsLane &ln, TsSink &sk, uint32_t seed, uint32_t spp, uint64_t idx) {
    Lane L = lane_setup(S, seed, spp, idx);
    ln.rng = L.rng; ln.o = L.o; ln.d = L.d; ln.pix = L.pix; ln.fx = L.fx; ln.fy = L.fy;
    ln.id = rcp3(ln.d);
    ln.depth = 0;
    ln.L_old = mk3(0, 0, 0); ln.L_new = mk3(0, 0, 0);
    ln.thr_old = mk3(1, 1, 1); ln.thr_new = mk3(1, 1, 1);
    ln.si_t = 0.f; ln.last_pdf = 1.f;
    ln.flags = F_NEEDS_ISECT | F_SPECULAR;
    (void) pcg_f32(ln.rng);                                           // :173 channel draw
    while (true) {
        if (!ts_loop_head(S, ln)) { ts_finish<MODE>(S, ln, sk); return; }
        Hit h = box_hit(S, ln.o, ln.d, ln.id, RCV_LARGEST);
        if (!h.valid) { ts_env_hit(S, ln); ts_finish<MODE>(S, ln, sk); return; }
        (void) pcg_f32(ln.rng); (void) pcg_f32(ln.rng); (void) pcg_f32(ln.rng);   // :347-348
        ln.o = offset_p(h.p, h.n, ln.d);
        if (!(dot3(h.n, ln.d) > 0.f)) { ln.ev = EV_COLL; return; }   // :378-379 entered the medium
    }
}

// path record <-> lane.  A record is written when the path enters the medium
// with depth 0, unit throughput and no radiance yet (rcv_scene_set rejects
// rr_depth < 0, the only way the first loop heads could rescale throughput).
__device__ __forceinline__ void ts_store(const TsLane &ln, const PathQueue &Q, unsigned long long pos) {
    Q.r0[pos] = make_float4(__uint_as_float((uint32_t) ln.rng.state), __uint_as_float((uint32_t) (ln.rng.state >> 32)),
                            __uint_as_float((uint32_t) (ln.rng.inc >> 1)), __int_as_float(ln.pix));
    Q.r1[pos] = make_float4(ln.o.x, ln.o.y, ln.o.z, ln.fx);
    Q.r2[pos] = make_float4(ln.d.x, ln.d.y, ln.d.z, ln.fy);
}
__device__ __forceinline__ void ts_load(TsLane &ln, const PathQueue &Q, unsigned long long pos) {
    float4 a = __ldg(Q.r0 + pos), b = __ldg(Q.r1 + pos), c = __ldg(Q.r2 + pos);
    ln.rng.state = (uint64_t) __float_as_uint(a.x) | ((uint64_t) __float_as_uint(a.y) << 32);
    ln.rng.inc = ((uint64_t) __float_as_uint(a.z) << 1) | 1u;
    ln.pix = (int) __float_as_uint(a.w);
    ln.o = mk3(b.x, b.y, b.z); ln.d = mk3(c.x, c.y, c.z);
    ln.fx = b.w; ln.fy = c.w;
    ln.id = rcp3(ln.d);
    ln.depth = 0;
    ln.L_old = mk3(0, 0, 0); ln.L_new = mk3(0, 0, 0);
    ln.thr_old = mk3(1, 1, 1); ln.thr_new = mk3(1, 1, 1);
    ln.si_t = 0.f; ln.last_pdf = 1.f;
    ln.flags = F_NEEDS_ISECT | F_SPECULAR;
    ln.ev = EV_COLL;
}

// end of the shadow ray: add the emitter-sampling contribution (:336-344),
// then the phase ratio of the old state (:360-361); the main path resumes
__device__ __forceinline__ void ts_nee_finish(const SceneK &S, TsLane &ln) {
    float epdf;
    V3 emitter_val = emitter_sample_value(S, ln.nd, epdf);
    float mis_em = mis_weight(epdf, ln.phase_val);
    ln.L_new = ln.L_new + ((ln.thr_new * ln.phase_val) * mis_em) * (emitter_val * splat3(ln.tr_new));
    ln.L_old = ln.L_old + ((ln.thr_old * ln.phase_val) * mis_em) * (emitter_val * splat3(ln.tr_old));
    ln.thr_old = ln.thr_old * ln.ratio;
    ln.flags = (ln.flags & ~F_IN_NEE) | F_NEEDS_ISECT;
    ln.id = rcp3(ln.d);
}

// the batched slow step (see the header comment)
template <int MODE>
__device__ __forceinline__ void ts_slow(const SceneK &S, TsLane &ln, TsSink &sk) {
    if (ln.flags & F_POST_NEE) {
        ln.flags &= ~F_POST_NEE;
        ts_nee_finish(S, ln);
        ln.ev = EV_COLL;
        return;
    }
    if (ln.flags & F_FIN_ESCAPE) {
        if (!(ln.flags & F_SI_VALID)) {
            ts_env_hit(S, ln);
        } else {
            // null-BSDF boundary crossing (:346-379), then one more loop head and the environment
            (void) pcg_f32(ln.rng); (void) pcg_f32(ln.rng); (void) pcg_f32(ln.rng);
            if (ts_loop_head(S, ln)) ts_env_hit(S, ln);
        }
    }
    ts_finish<MODE>(S, ln, sk);
}

template <int MODE, class GRID>
__device__ __forceinline__ void ts_collide(const SceneK &S, const GRID &G, float mu_vol, float inv_mu_vol, TsLane &ln, TsSink &sk) {
    const bool nee = (ln.flags & F_IN_NEE) != 0;
    V3 so, sd; float maxt;
    if (!nee) {
        if (!ts_loop_head(S, ln)) { ln.flags |= F_FIN; ln.ev = EV_SLOW; return; }
        so = ln.o; sd = ln.d; maxt = RCV_LARGEST;
    } else {
        float remaining = S.env_dist * (1.f - RCV_SHADOW_EPS) - ln.total;   // :409
        if (!(remaining > 0.f)) { ln.flags |= F_POST_NEE; ln.ev = EV_SLOW; return; }   // :411
        so = ln.no; sd = ln.nd; maxt = remaining;
    }
    float u = pcg_f32(ln.rng);                                        // :194 / :422
    Mei2 m; Tri T;
    sample_interaction2<MODE>(S, G, mu_vol, inv_mu_vol, so, sd, ln.id, maxt, u, m, T, sk.n_coll);
    const float mu = m.mu, inv_mu = m.inv_mu;                         // the volume's majorant, or the supergrid cell's at m.p
    if (ln.flags & F_NEEDS_ISECT) {                                   // :200-204 / :414-416
        Hit h = box_hit(S, so, sd, ln.id, maxt);
        ln.si_t = h.t;
        ln.flags = (ln.flags & ~(F_NEEDS_ISECT | F_SI_VALID)) | (h.valid ? F_SI_VALID : 0u);
    }
    if (ln.si_t < m.t) m.t = inf_();                                  // :205-206 / :423-424
    if (m.t < inf_()) {
        if (nee) {                                                    // :444-452, :472-473
            ln.no = m.p;
            ln.si_t = ln.si_t - m.t;
            float sn_o = mu - m.so, sn_n = mu - m.sn;
            ln.tr_old = ln.tr_old * ((mu != 0.f) ? (sn_o * inv_mu) : sn_o);
            ln.tr_new = ln.tr_new * ((mu != 0.f) ? (sn_n * inv_mu) : sn_n);
            if (!(ln.tr_new != 0.f || ln.tr_old != 0.f)) { ln.flags |= F_POST_NEE; ln.ev = EV_SLOW; return; }
            ln.total += m.t;
            prefetch_next_brick<true>(S, G, inv_mu, ln.rng, 0, ln.no, ln.nd);
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
            prefetch_next_brick<true>(S, G, inv_mu, ln.rng, 1, ln.o, ln.d);      // next step: loop-head draw, then the distance
            return;
        }
        ln.depth += 1;                                                // :249
        if (!(ln.depth < S.max_depth)) { ln.flags |= F_FIN; ln.ev = EV_SLOW; return; }   // :253
        ln.o = m.p;                                                   // the vertex; the new direction is chosen in EV_SCAT
        ln.tr_old = m.so; ln.tr_new = m.sn;
        ln.ev = EV_SCAT;
        return;
    }
    // the segment left the medium
    if (nee) { ln.flags |= F_POST_NEE; ln.ev = EV_SLOW; return; }     // nothing attenuates the shadow ray outside the box
    ln.flags |= F_FIN | F_FIN_ESCAPE;
    ln.ev = EV_SLOW;
}

template <int MODE>
__device__ __forceinline__ void ts_scatter(const SceneK &S, const GridK &G, TsLane &ln, TsSink &sk) {
    sk.n_scat++;
    const float st_old = ln.tr_old, st_new = ln.tr_new;
    Tri T = tri_setup(S, G, ln.o);
    V3 a_new, a_old;
    if (MODE == 0) lookup_albedo2(G, T, a_old, a_new);
    else { a_new = lookup_albedo(G, T, MODE == 1 ? 0 : 1); a_old = a_new; }
    V3 ss_new = splat3(st_new) * a_new, ss_old = splat3(st_old) * a_old;
    float rden = 1.f / (st_new + st_old);                             // :259-260
    ln.thr_new = ln.thr_new * ((ss_new * 2.f) * rden);
    ln.thr_old = ln.thr_old * ((ss_old * 2.f) * rden);
    (void) pcg_f32(ln.rng);                                           // :275
    float s2x = pcg_f32(ln.rng), s2y = pcg_f32(ln.rng);
    V3 wi = mk3(-ln.d.x, -ln.d.y, -ln.d.z);
    float pdf;
    V3 wo = phase_sample(S, wi, s2x, s2y, pdf);
    // pdf > 0 always holds for the isotropic and HG phase functions (:276)
    ln.d = wo;
    ln.flags |= F_NEEDS_ISECT;
    ln.last_pdf = pdf;
    float ratio = phase_eval(S, wi, wo) / pdf;                        // :360-361
    ln.ev = EV_COLL;
    if (S.use_nee) {                                                  // :315-325
        ln.flags &= ~F_SPECULAR;
        if (!S.sample_emitters) ln.flags |= F_SPECULAR;
        if (S.sample_emitters) {
            float e2x = pcg_f32(ln.rng), e2y = pcg_f32(ln.rng);       // :393
            ln.nd = emitter_sample_dir(S, e2x, e2y);
            if (S.emitter == 1) {                                     // :395-397 a sample of zero density is invalid
                float epdf;
                (void) emitter_sample_value(S, ln.nd, epdf);
                if (epdf == 0.f) { ln.thr_old = ln.thr_old * ratio; ln.id = rcp3(ln.d); return; }
            }
            ln.id = rcp3(ln.nd);
            ln.no = ln.o;
            ln.total = 0.f; ln.tr_old = 1.f; ln.tr_new = 1.f;
            ln.phase_val = phase_eval(S, wi, ln.nd);
            ln.ratio = ratio;
            ln.flags |= F_IN_NEE | F_NEEDS_ISECT;
            return;
        }
    }
    ln.thr_old = ln.thr_old * ratio;
    ln.id = rcp3(ln.d);
}

}  // namespace rcv
