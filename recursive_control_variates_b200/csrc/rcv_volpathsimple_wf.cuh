// rcv_volpathsimple_wf.cuh — VolpathSimpleIntegrator (plugins/volpathsimple.py;
// use_drt = True: see "DRT" below) as a per-lane event machine, for the persistent-warp
// schedulers.  Same estimator, same random streams and the same arithmetic as
// rcv_volpathsimple.cuh (the straight-line restatement kept as A/B baseline), but
// the adjoint is organised as a VERTEX LIST:
//
//   pass 1 (primal, RBIntegrator.render_backward's first sample() call) traces the
//     path and writes one 64-byte record per path vertex (real collision or
//     medium exit) into a small per-slot arena: ray, segment length, whether a
//     shadow ray was cast, the sampler state before it, its direction and its
//     contribution;
//   pass 2 (adjoint) does NOT re-trace the path: it walks the vertex list and
//     evaluates, per vertex, exactly the terms the reference evaluates while
//     replaying — scattering term (:165-185), transmittance term (:194-202,
//     :597-620) and the shadow ray's transmittance gradients (:406-414,
//     :496-505), the only part that needs tracing again (with the parked sampler
//     state, like the reference's cloned sampler).  The null collisions of the
//     main path and the primal shadow rays are not repeated: ~44 % fewer gathers
//     than replaying, identical gradient terms.
//
//   vp_raygen  (its own, fully convergent kernel) camera ray, dummy channel
//              draw (:84), reach_medium (:305-332), alt_seed_rnd draw (:112);
//              paths that never enter the medium end here, the others become
//              48-byte path records in HBM
//   EV_COLL    ONE tentative collision: a delta-tracking step of
//              sample_real_interaction (:360-380) or a ratio-tracking step of
//              estimate_transmittance (:478-515), in pass 2 with the sigma_t
//              gradient scatter of :496-505
//   EV_SCAT    pass 1: a path vertex — albedo gather, throughput/depth update,
//              vertex record, NEE set-up (:393-446) or phase sampling (:234-258);
//              pass 2: the next vertex of the list — coefficient gathers,
//              gradient scatters, start of the shadow-ray replay
//   EV_SLOW    once-per-shadow-ray / once-per-pass work that would otherwise run
//              with one or two lanes: end of a shadow ray (+ phase sampling) or of
//              its replay, end of a pass; the scheduler refills lanes in this step
//
// use_drt = True with sub-sampling (volpathsimple.py:39-47, :262-272) runs on the same machines in three stages
// (template parameter DRT of the handlers):
//   DRT = 1  the adjoint above, plus the DRTReservoir (:743-778) updated while pass 2 walks the vertex list: the
//            throughput at the head of every iteration is rebuilt from the albedos pass 2 gathers anyway and the
//            Russian-roulette factors, in lane words that pass 2 does not use otherwise (thr, depth, last_pdf,
//            saved); the MIS weight of the path-replay term (:165-170); at the end of pass 2 a JOB record:
//            selected ray, depth, adjoint weight, state of the secondary sampler
//   (k_drt_sample, rcv_trace.cu: distance sampling, shadow ray and phase sampling of the job: :560-561, :635-654)
//   DRT = 2  the recursive radiance estimate (:656-665): the primal machine resumed from the job's PathState on the
//            secondary sampler; its result is scattered into the gradient grids instead of the film
//
// The handlers contain no warp intrinsics and also compile for the host
// (RCV_HOST_EMU) where tests single-step them against the oracle.
#pragma once
#include "rcv_twostate_wf.cuh"
#include "rcv_volpathsimple.cuh"

namespace rcv {

enum : uint32_t {
    VF_ADJ = 1u,          // this is pass 2 (adjoint)
    VF_HAS_SCAT = 2u,
    VF_ESCAPED = 4u,
    VF_IN_NEE = 8u,
    VF_REPLAY = 16u,      // shadow-ray transmittance replay with gradients (pass 2)
    VF_HEAD = 32u,        // next main-path collision starts with the loop head (:129-134)
    VF_DID_SCATTER = 64u,
    VF_NEE_OVF = 128u,    // pass 1: the shadow ray had more tentative collisions than its record buffer holds
    // work parked for the batched EV_SLOW step
    VF_POSTNEE = 512u,    // the shadow ray (or its replay) ended
    VF_END = 1024u,       // the pass ended at a loop head (:132-133)
    VF_JOB = 2048u,       // DRT: the lane is complete and its job (DRT = 1) / its gradient (DRT = 2) is to be written
    VF_ANY_OVF = 4096u    // pass 1: some shadow ray of this path overflowed its record buffer (pass 2 must replay it)
};
// a fourth scheduling class of the pool kernel (RCV_FLAT_PASS2): pass 1 is complete and the lane's vertices wait for a
// flattened pass-2 round, in which every thread of the warp takes ONE vertex of whichever waiting lane (vp_vertex2_flat)
enum : int { EV_P2 = 4 };
// r2p (profiles/README.md): 15.37 ms against 15.52 ms for the list walk at C2 — within run-to-run noise, +6 registers and a
// 16-byte spill: the adjoint's second pass is bound by the read-modify-write traffic of its gradient scatters, not by the
// SIMD efficiency of its vertex rounds.  Off by default; the per-vertex function is covered by the host-emulation tests.
#ifndef RCV_FLAT_PASS2
#define RCV_FLAT_PASS2 0
#endif

// vertex record flags
enum : uint32_t { VR_SCATTER = 1u, VR_NEE = 2u, VR_NEE_DONE = 4u /* its transmittance gradients were scattered in pass 1 */,
                  VR_ALBEDO = 8u /* the record carries the vertex's albedo (RCV_VERTEX_ALBEDO) */,
                  VR_SIGMA = 16u /* ... and its sigma_t, in r[2].x (RCV_VERTEX_SIGMA) */ };
// RCV_VERTEX_SIGMA = 1: pass 1 also fetches sigma_t at the vertex (same interpolation set-up as its albedo gather, the
// loads go out together) and leaves it in the record: pass 2 then needs no grid gather at all — vertex records (L2 hits)
// in, gradient reductions out.  Same value, same arithmetic as the lookup pass 2 would make.
#ifndef RCV_VERTEX_SIGMA
#define RCV_VERTEX_SIGMA 1
#endif
// unroll factor of the four transmittance-term sites of a pass-2 vertex (:607-613)
#ifndef RCV_TRANS_UNROLL
#define RCV_TRANS_UNROLL 1
#endif
static constexpr int kTransUnroll = RCV_TRANS_UNROLL;
// RCV_VERTEX_ALBEDO = 1: pass 1 leaves the albedo it gathered at a scattering vertex in the vertex record (in the words
// that hold the shadow-ray direction, which only the replay fall-back needs), so pass 2 does not gather its eight
// corners again.  Same value, same arithmetic.  r2a (profiles/README.md): adjoint 17.8 -> 16.9 ms at C2; on by default.
#ifndef RCV_VERTEX_ALBEDO
#define RCV_VERTEX_ALBEDO 1
#endif
#if !RCV_VERTEX_ALBEDO           // the record words are laid out for both or neither
#undef RCV_VERTEX_SIGMA
#define RCV_VERTEX_SIGMA 0
#endif
// records (position, -1 / (majorant * tr)) of the tentative collisions of one shadow ray, kept per arena slot behind the
// vertex records; 16 vertex-record equivalents
static constexpr int kNeeCap = 64;
static constexpr int kVertexCapMax = 4096;   // largest supported max_depth + 1 for the vertex-list adjoint

struct VpLane {
    int ev, pix, depth;
    uint32_t idx, flags;
    uint32_t vbase;         // first record of this lane's arena slot
    int nvert, vk;          // vertices recorded in pass 1 / cursor of pass 2
    Pcg rng, alt;
    uint64_t saved;         // main sampler state before the shadow ray (:396 sampler.clone())
    V3 o, d;                // current ray (origin = last vertex)
    V3 id;                  // 1/direction of the current segment (ray, or shadow ray during NEE)
    float ray_maxt;         // distance to the boundary along the ray (== si.t)
    V3 thr, result, dL;
    float last_pdf;
    V3 ro;                  // running origin: delta tracking, or the shadow ray
    float rmaxt, running_t, mei_t;
    V3 nd;                  // shadow-ray direction
    float tr_acc;
    V3 nee;                 // NEE contribution of the current vertex
};

struct VpSink {
    float4 *accum;          // primal render target (H,W,4) or null
    GradK D;                // gradient grids (adjoint)
    const float *grad_img, *film_w;
    float4 *arena;          // vertex records, 4 float4 each
    uint32_t vcap;          // records per arena slot (>= max_depth + 1)
    uint32_t seed, spp, alt_seed;
    VpsCounts cnt;
    unsigned n_paths;
    int drt_mis;            // use_drt_mis (DRT != 0)
    int flat;               // the scheduler runs pass 2 vertex-parallel (EV_P2 / vp_vertex2_flat); else the lane walks its list
    PathQueue jobs;         // DRT = 1: job records written at the end of pass 2; DRT = 2: r0/r1 = where and what to scatter
    // coop != 0: a finished shadow ray leaves its recorded collisions to the scheduler (vp_nee_records_warp) instead of
    // walking them itself: how many, where, and the adjoint they are weighted with
    int coop, coop_n;
    const float4 *coop_rec;
    V3 coop_adj;
};

__device__ __forceinline__ float4 *vp_record(const VpSink &sk, const VpLane &ln, int k) {
    return sk.arena + ((size_t) ln.vbase + (size_t) k) * 4;
}
__device__ __forceinline__ float4 *vp_nee_record(const VpSink &sk, const VpLane &ln, int i) {
    return sk.arena + ((size_t) ln.vbase + (size_t) sk.vcap) * 4 + i;
}

// RCV_PREFETCH_VERTEX = 1: pass 2 asks for the NEXT vertex record (L1) when it leaves a vertex — the record is an L2
// hit whose latency the vertex step otherwise waits out at its first use (10 % of the adjoint's stall samples, r2a)
#ifndef RCV_PREFETCH_VERTEX
#define RCV_PREFETCH_VERTEX 0
#endif
__device__ __forceinline__ void prefetch_l1(const void *p) {
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void) p;
#endif
}
// end of pass 2 (or of a render-only path)
template <int DRT = 0>
__device__ __forceinline__ void vp_finish(VpLane &ln, VpSink &sk) {
    if (DRT != 2) sk.n_paths++;
    ln.flags = (DRT != 0) ? (uint32_t) VF_JOB : 0u;
    ln.ev = EV_INIT;
}

// DRTReservoir (:743-778) of the lane: the spare last record of its arena slot
__device__ __forceinline__ float4 *vp_reservoir(const VpSink &sk, const VpLane &ln) { return vp_record(sk, ln, (int) sk.vcap - 1); }

// end of pass 1: environment contribution (:276-300); then either splat (render) or start pass 2
template <bool BW, int DRT = 0>
__device__ __forceinline__ void vp_end_pass(const SceneK &S, VpLane &ln, VpSink &sk) {
    if ((ln.flags & VF_ESCAPED) && !((ln.depth <= 0) && S.hide_emitters)) {
        float w = 1.f, epdf;
        const V3 env = emitter_eval(S, ln.d, epdf);
        if (S.use_nee) w = mis_weight(ln.last_pdf, (ln.flags & VF_HAS_SCAT) ? epdf : 0.f);
        ln.result = ln.result + (ln.thr * w) * env;
    }
    if (!BW) {
        if (DRT == 2) { vp_finish<2>(ln, sk); return; }   // the scheduler scatters ln.result (vp_drt_scatter)
        if (S.rfilter == 0) {
            if (ln.pix >= 0) red_add_v4(sk.accum + ln.pix, ln.result.x, ln.result.y, ln.result.z, 1.f);
        } else {
            film_splat_filtered_lane(S, sk.accum, sk.seed, sk.spp, ln.idx, ln.result.x, ln.result.y, ln.result.z);
        }
        vp_finish(ln, sk);
        return;
    }
    sk.n_paths++;                                   // pass 1 complete
    // ln.dL = d(develop o put)/dL * grad_in of this sample came with the path record (computed by ray generation)
    // pass 2 starts from state_in = the pass-1 result (:68) with the secondary sampler (:119-120)
    if (DRT == 0 && sk.flat && !(ln.flags & VF_ANY_OVF)) {
        // pass 2 of this lane runs vertex-parallel: result = the pass-1 radiance, vk = 0 vertices done so far
        ln.flags = VF_ADJ;
        ln.vk = 0;
        if (ln.nvert > 0) ln.ev = EV_P2;
        else vp_finish<DRT>(ln, sk);
        return;
    }
    sampler_seed(ln.alt, sk.alt_seed, ln.idx);
    ln.flags = VF_ADJ;
    ln.vk = 0;
    if (DRT == 1) {
        // pass 2 rebuilds the throughput (thr) for the DRTReservoir (:107-109): wsum in (last_pdf, saved), the
        // selected vertex in depth, current_weight in the spare arena record
        ln.thr = mk3(1, 1, 1);
        ln.depth = -1;
        ln.last_pdf = 0.f;
        ln.saved = 0;
    }
    if (ln.nvert > 0) {
        ln.ev = EV_SCAT;
#if RCV_PREFETCH_VERTEX
        prefetch_l1(vp_record(sk, ln, 0));
#endif
    } else vp_finish<DRT>(ln, sk);
}

// Start of a path up to the first loop iteration: camera ray, dummy channel
// draw (:84), reach_medium (:305-332), alt_seed_rnd draw (:112).  Paths that do
// not enter the medium are complete: their primal value is splatted here when
// rendering (BW = false); they carry no gradient.  ev = EV_COLL: record it.
template <bool BW>
__device__ __forceinline__ void vp_raygen(const SceneK &S, VpLane &ln, VpSink &sk, uint64_t idx) {
    Lane L = lane_setup(S, sk.seed, sk.spp, idx);
    ln.idx = (uint32_t) idx;
    ln.rng = L.rng; ln.o = L.o; ln.d = L.d; ln.pix = L.pix;
    ln.id = rcp3(ln.d);
    ln.result = mk3(0, 0, 0);
    ln.thr = mk3(1, 1, 1);
    ln.depth = 0;
    ln.last_pdf = 1.f;
    ln.flags = 0;
    ln.nvert = 0; ln.vk = 0;
    (void) pcg_f32(ln.rng);                                           // :84
    Hit si = box_hit(S, ln.o, ln.d, ln.id, RCV_LARGEST);              // :311
    bool active = si.valid;
    if (!si.valid) ln.flags |= VF_ESCAPED;
    ln.ray_maxt = RCV_LARGEST;
    if (active) {
        ln.o = offset_p(si.p, si.n, ln.d);                            // :319
        Hit si2 = box_hit(S, ln.o, ln.d, ln.id, RCV_LARGEST);         // :320
        active = si2.valid;                                           // :323
        if (active) ln.ray_maxt = isfinite(si2.t) ? si2.t : RCV_LARGEST;   // :329
    }
    if (!active) {
        if (BW) {
            sk.n_paths += 2;                    // nothing to differentiate: both passes are complete
            ln.flags = 0; ln.ev = EV_INIT;
        } else {
            vp_end_pass<false>(S, ln, sk);      // environment contribution (:276-300) + splat
        }
        return;
    }
    (void) pcg_f32(ln.rng);                                           // :112 alt_seed_rnd
    ln.ev = EV_COLL;
}

__device__ __forceinline__ void vp_store(const VpLane &ln, const PathQueue &Q, unsigned long long pos) {
    Q.r0[pos] = make_float4(__uint_as_float((uint32_t) ln.rng.state), __uint_as_float((uint32_t) (ln.rng.state >> 32)),
                            __uint_as_float((uint32_t) (ln.rng.inc >> 1)), __int_as_float(ln.pix));
    Q.r1[pos] = make_float4(ln.o.x, ln.o.y, ln.o.z, ln.ray_maxt);
    Q.r2[pos] = make_float4(ln.d.x, ln.d.y, ln.d.z, __uint_as_float(ln.idx));
}
// begin pass 1 of the path recorded at pos in arena slot `slot` (Q.r3 != null: the adjoint, the record carries dL)
template <int DRT = 0>
__device__ __forceinline__ void vp_load(VpLane &ln, const VpSink &sk, const PathQueue &Q, unsigned long long pos, uint32_t slot) {
    float4 a = __ldg(Q.r0 + pos), b = __ldg(Q.r1 + pos), c = __ldg(Q.r2 + pos);
    ln.dL = mk3(0, 0, 0);
    if (DRT != 2 && Q.r3) { float4 g = __ldg(Q.r3 + pos); ln.dL = mk3(g.x, g.y, g.z); }
    ln.rng.state = (uint64_t) __float_as_uint(a.x) | ((uint64_t) __float_as_uint(a.y) << 32);
    ln.rng.inc = ((uint64_t) __float_as_uint(a.z) << 1) | 1u;
    ln.pix = (int) __float_as_uint(a.w);
    ln.o = mk3(b.x, b.y, b.z); ln.ray_maxt = b.w;
    ln.d = mk3(c.x, c.y, c.z); ln.idx = __float_as_uint(c.w);
    ln.id = rcp3(ln.d);
    ln.result = mk3(0, 0, 0);                                         // :68
    ln.thr = mk3(1, 1, 1);
    ln.depth = 0;
    ln.last_pdf = 1.f;
    ln.flags = VF_HEAD;
    ln.vbase = slot * (sk.vcap + kNeeCap / 4);
    ln.nvert = 0; ln.vk = 0;
    ln.ev = EV_COLL;
    if (DRT == 2) {
        // PathState of the recursive call (:75-81, :98-100, :656-665): a path record written by k_drt_sample.
        // pix = its job, idx word = depth, r3 = radiance so far (the NEE of :635-638) and the phase pdf
        float4 g = __ldg(Q.r3 + pos);
        ln.depth = (int) ln.idx; ln.idx = 0;
        ln.result = mk3(g.x, g.y, g.z);
        ln.last_pdf = g.w;
        ln.flags = VF_HEAD | VF_HAS_SCAT;
    }
}

// (re)start the shadow ray at the current vertex; false when it has no extent
__device__ __forceinline__ bool vp_nee_start(const SceneK &S, VpLane &ln) {
    V3 mei_p = ray_at(ln.o, ln.d, ln.mei_t);
    Hit si = box_hit(S, mei_p, ln.nd, ln.id, RCV_LARGEST);            // :440-441
    ln.ro = ray_at(mei_p, ln.nd, 0.f);                                // :467 (tmin = 0)
    ln.rmaxt = si.t;
    ln.tr_acc = si.valid ? 1.f : 0.f;                                 // :472
    return si.valid;
}

// pass 1, after the vertex (and its NEE): phase sampling, new ray, boundary distance (:234-258)
template <bool BW, int DRT = 0>
__device__ __forceinline__ void vp_post_vertex(const SceneK &S, VpLane &ln, VpSink &sk, bool active) {
    const bool did_scatter = (ln.flags & VF_DID_SCATTER) != 0;
    if (!did_scatter) {                                               // did_escape (:257-258)
        ln.flags |= VF_ESCAPED;
        vp_end_pass<BW, DRT>(S, ln, sk);
        return;
    }
    V3 mei_p = ray_at(ln.o, ln.d, ln.mei_t);
    V3 wi = mk3(-ln.d.x, -ln.d.y, -ln.d.z);
    (void) pcg_f32(ln.rng);                                           // :235
    float s2x = pcg_f32(ln.rng), s2y = pcg_f32(ln.rng);
    float pdf;
    V3 wo = phase_sample(S, wi, s2x, s2y, pdf);
    ln.o = mei_p; ln.d = wo; ln.id = rcp3(wo);
    ln.last_pdf = pdf;
    Hit si = box_hit(S, ln.o, ln.d, ln.id, RCV_LARGEST);              // :247
    ln.ray_maxt = isfinite(si.t) ? si.t : RCV_LARGEST;                // :248
    if (!si.valid) active = false;                                    // :253-254
    if (!active) { vp_end_pass<BW, DRT>(S, ln, sk); return; }
    ln.flags |= VF_HEAD;
    ln.ev = EV_COLL;
}

// Lists of n[0], ..., n[31] items laid end to end (vp_nee_records_warp in rcv_trace.cu: thread t holds the exclusive
// prefix sum excl[t]): the owner of item j is the LAST t with excl[t] <= j — empty lists share their start with the
// next one, so the last one is the list that holds j.  Five-step search; `excl_of(t)` reads thread t's prefix sum (a
// shuffle on the device, an array on the host where the tests run it).  j beyond the total gives 31.
template <class F>
__device__ __forceinline__ int concat_owner(int j, F excl_of) {
    int own = 0;
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
        const int cand = own + step;
        const int e = excl_of(cand & 31);
        if (cand < 32 && e <= j) own = cand;
    }
    return own;
}

// one recorded collision of a shadow ray: adjoint * d log(tr) / d sigma_t at its position (:496-505)
__device__ __forceinline__ void vp_nee_record_scatter(const SceneK &S, const GridK &G, VpSink &sk, V3 adjoint, float4 c) {
    float g = 0.f;
    g += adjoint.x * c.w; g += adjoint.y * c.w; g += adjoint.z * c.w;
    scatter_sigma_t(S, G, sk.D, mk3(c.x, c.y, c.z), g, sk.cnt);
}

// pass 1: the shadow ray ended — its contribution (:401-404, :225), the vertex record, then the path continues
template <bool BW, int DRT = 0>
__device__ __forceinline__ void vp_nee_done(const SceneK &S, const GridK &G, VpLane &ln, VpSink &sk) {
    V3 wi = mk3(-ln.d.x, -ln.d.y, -ln.d.z);
    float epdf;
    V3 emitted = emitter_sample_value(S, ln.nd, epdf) * splat3(ln.tr_acc);
    float phase_val = phase_eval(S, wi, ln.nd);                       // :401
    ln.nee = ((ln.thr * phase_val) * mis_weight(epdf, phase_val)) * emitted;   // :404
    ln.result = ln.result + ln.nee;                                   // :225
    if (BW) {
        float4 *r = vp_record(sk, ln, ln.nvert - 1);
        const bool ovf = (ln.flags & VF_NEE_OVF) != 0;
#if RCV_VERTEX_SIGMA
        if (ovf) r[2] = make_float4(__uint_as_float((uint32_t) ln.saved), __uint_as_float((uint32_t) (ln.saved >> 32)), ln.nee.x, ln.nee.y);
        else { r[2].z = ln.nee.x; r[2].w = ln.nee.y; }               // .x keeps sigma_t; the sampler state is needed by the replay only
#else
        r[2] = make_float4(__uint_as_float((uint32_t) ln.saved), __uint_as_float((uint32_t) (ln.saved >> 32)), ln.nee.x, ln.nee.y);
#endif
#if RCV_VERTEX_ALBEDO
        if (ovf) r[3] = make_float4(ln.nee.z, ln.nd.x, ln.nd.y, ln.nd.z);      // the replay needs the direction: the albedo gives way
        else r[3].x = ln.nee.z;
#else
        r[3] = make_float4(ln.nee.z, ln.nd.x, ln.nd.y, ln.nd.z);
#endif
        // a vertex that cast a shadow ray is a scattering vertex with VR_NEE set (vp_vertex1): its flag word is written
        // whole, no read-modify-write of the record
        r[1].w = __uint_as_float(VR_SCATTER | VR_NEE | (ovf ? 0u : (uint32_t) VR_NEE_DONE) |
                                 ((RCV_VERTEX_ALBEDO && !ovf) ? (uint32_t) VR_ALBEDO : 0u) |
                                 ((RCV_VERTEX_SIGMA && !ovf) ? (uint32_t) VR_SIGMA : 0u));
        if (!ovf) {
            // transmittance gradients of the shadow ray (:406-414, :496-505) from the recorded collisions: the same
            // values the replay in pass 2 would compute, scattered now.  The records are read three ahead of their use
            // (they are L2 hits a few hundred cycles away; one dependent load per iteration stalled the whole step).
            const V3 adjoint = ln.dL * ln.nee;                        // :412
            const int n = __float_as_int(ln.running_t);
            const float4 *rec = vp_nee_record(sk, ln, 0);
            if (sk.coop) { sk.coop_n = n; sk.coop_rec = rec; sk.coop_adj = adjoint; }
            else {
                const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
                float4 c0 = (0 < n) ? rec[0] : zero, c1 = (1 < n) ? rec[1] : zero, c2 = (2 < n) ? rec[2] : zero;
                for (int i = 0; i < n; ++i) {
                    const float4 c3 = (i + 3 < n) ? rec[i + 3] : zero;
                    vp_nee_record_scatter(S, G, sk, adjoint, c0);
                    c0 = c1; c1 = c2; c2 = c3;
                }
            }
        }
        ln.flags &= ~VF_NEE_OVF;
    }
    ln.flags &= ~VF_IN_NEE;
    vp_post_vertex<BW, DRT>(S, ln, sk, true);
}

// pass 2: the replay of a shadow ray ended (:227), next vertex
template <int DRT = 0>
__device__ __forceinline__ void vp_next_vertex(VpLane &ln, VpSink &sk) {
    ln.vk += 1;
    if (ln.vk < ln.nvert) {
        ln.ev = EV_SCAT;
#if RCV_PREFETCH_VERTEX
        prefetch_l1(vp_record(sk, ln, ln.vk));
#endif
    } else vp_finish<DRT>(ln, sk);
}
template <int DRT = 0>
__device__ __forceinline__ void vp_replay_done(VpLane &ln, VpSink &sk) {
    ln.result = ln.result - ln.nee;                                   // :227
    ln.flags &= ~(VF_IN_NEE | VF_REPLAY);
    vp_next_vertex<DRT>(ln, sk);
}

// One tentative collision in two phases, so that a scheduler can start the bricks of several lanes before it consumes
// the first (rcv_pool.cuh, RCV_POOL_ILP): _pre runs the loop head, samples the free flight and issues the brick load,
// _post interpolates sigma_t and takes the null / real / escape decision.  vp_collide is the two back to back.
struct VpColl {
    bool go;            // false: the lane left the collision class at its loop head, nothing is in flight
    Mei1 m; Tri T; Brick1 b;
};

template <bool BW, class GRID>
__device__ __forceinline__ void vp_collide_pre(const SceneK &S, const GRID &G, float mu_vol, float inv_mu_vol, VpLane &ln, VpSink &sk,
                                               VpColl &c) {
    c.go = false;
    const bool nee = (ln.flags & VF_IN_NEE) != 0;
    if (!nee && (ln.flags & VF_HEAD)) {                               // loop head (:129-134)
        float q = fmin_(max3(ln.thr), 0.99f);
        bool perform_rr = ln.depth > S.rr_depth;
        float rr = pcg_f32(ln.rng);
        if (!(any_nz(ln.thr) && (!perform_rr || (rr < q)))) { ln.flags |= VF_END; ln.ev = EV_SLOW; return; }
        if (perform_rr) ln.thr = ln.thr * (1.f / q);
        ln.ro = ln.o; ln.rmaxt = ln.ray_maxt; ln.running_t = 0.f;     // :349-351
        ln.flags &= ~VF_HEAD;
    }
    // one free-flight sample + sigma_t gather on the current segment, shared by
    // delta tracking (:361) and ratio tracking (:482) so that all lanes of the
    // warp execute the gather together
    const V3 sd = nee ? ln.nd : ln.d;
    float u = pcg_f32(ln.rng);
    sample_interaction1_issue(S, G, mu_vol, inv_mu_vol, ln.ro, sd, ln.id, ln.rmaxt, u, c.m, c.T, c.b, sk.cnt);
    c.go = true;
}

template <bool BW, class GRID>
__device__ __forceinline__ void vp_collide_post(const SceneK &S, const GRID &G, VpLane &ln, VpSink &sk, const VpColl &c) {
    if (!c.go) return;
    const bool nee = (ln.flags & VF_IN_NEE) != 0;
    Mei1 m = c.m;
    const Tri &T = c.T;
    sample_interaction1_finish(S, m, T, c.b);
    const float mu = m.mu, inv_mu = m.inv_mu;                         // the volume's majorant, or the supergrid cell's at m.p
    if (nee) {
        // estimate_transmittance step (:478-515)
        float sn = m.valid ? (mu - m.st) : 0.f;
        float ce = m.valid ? mu : 0.f;
        float tr = (ce != 0.f) ? (sn * inv_mu) : sn;                  // :486-489
        if (m.t > ln.rmaxt) m.t = inf_();                             // :493
        if (!(m.t < inf_())) { ln.flags |= VF_POSTNEE; ln.ev = EV_SLOW; return; }  // :494
        if (BW && !(ln.flags & VF_ADJ)) {
            // pass 1 of the adjoint: remember this tentative collision.  Its gradient, adjoint * d log(tr) / d sigma_t
            // (:496-505), needs the finished shadow ray's contribution; it is scattered when the ray ends (vp_nee_done)
            // instead of replaying the ray in pass 2.  running_t is idle during a shadow ray: it counts the records.
            int n = __float_as_int(ln.running_t);
            if (n < kNeeCap) *vp_nee_record(sk, ln, n) = make_float4(m.p.x, m.p.y, m.p.z, (tr > 0.f) ? ((-inv_mu) / tr) : 0.f);
            else ln.flags |= VF_NEE_OVF | VF_ANY_OVF;
            ln.running_t = __int_as_float(n + 1);
        }
        if (BW && (ln.flags & VF_REPLAY)) {                           // :496-505
            V3 adjoint = ln.dL * ln.nee;                              // :412
            float g = 0.f;
            if (tr > 0.f) {
                float k = (-inv_mu) / tr;
                g += adjoint.x * k; g += adjoint.y * k; g += adjoint.z * k;
            }
            scatter_sigma_t_T(G, sk.D, T, S.scale * g, sk.cnt);           // the gather's set-up is reused
        }
        ln.tr_acc = ln.tr_acc * tr;                                   // :508
        ln.ro = m.p;                                                  // :510
        ln.rmaxt = ln.rmaxt - m.t;
        if (!(ln.tr_acc != 0.f)) { ln.flags |= VF_POSTNEE; ln.ev = EV_SLOW; }   // :515
        return;
    }
    // sample_real_interaction step (:360-380)
    ln.mei_t = m.t + ln.running_t;                                    // :364
    if (!m.valid) { ln.flags &= ~VF_DID_SCATTER; ln.ev = EV_SCAT; return; }   // :371 escape
    float r = (mu != 0.f) ? (m.st * inv_mu) : 0.f;                    // :367
    float xi = pcg_f32(ln.rng);                                       // :372
    if (!(xi >= r)) { ln.flags |= VF_DID_SCATTER; ln.ev = EV_SCAT; return; }
    ln.ro = ln.ro + ln.d * m.t;                                       // :378-380
    ln.rmaxt = ln.rmaxt - m.t;
    ln.running_t = ln.running_t + m.t;
}

template <bool BW, class GRID>
__device__ __forceinline__ void vp_collide(const SceneK &S, const GRID &G, float mu_vol, float inv_mu_vol, VpLane &ln, VpSink &sk) {
    VpColl c;
    vp_collide_pre<BW>(S, G, mu_vol, inv_mu_vol, ln, sk, c);
    vp_collide_post<BW>(S, G, ln, sk, c);
}

// RCV_SPEC_COLLIDE = 2: TWO consecutive tentative collisions of a segment per call, both bricks in flight together.
// Where the next tentative collision falls does not depend on the lookup that is still in flight — only on the lane's
// random stream and on the current one being a null collision (delta tracking) or an ordinary ratio-tracking step —
// so its free flight is sampled and its brick requested speculatively, as if that were the case.  If the first
// collision turns out real (or the segment ends) the speculation is dropped: the lane's stream was advanced on a copy
// and nothing else was touched, so results are bit-identical with or without it.  A path then needs about half as
// many dependent memory round trips, which is what bounds these kernels (profiles/README.md).
#ifndef RCV_SPEC_COLLIDE
#define RCV_SPEC_COLLIDE 2
#endif
template <bool BW, class GRID>
__device__ __forceinline__ void vp_collide_spec(const SceneK &S, const GRID &G, float mu_vol, float inv_mu_vol, VpLane &ln, VpSink &sk) {
    VpColl a;
    vp_collide_pre<BW>(S, G, mu_vol, inv_mu_vol, ln, sk, a);
    if (!a.go) return;
    VpColl b; b.go = false;
    Pcg rng_b = ln.rng;                                               // the stream after this collision's distance draw
    VpsCounts spec; spec.coll = spec.scat = spec.g_sig = spec.g_alb = 0;
    if (a.m.valid) {
        const bool nee = (ln.flags & VF_IN_NEE) != 0;
        V3 ro_b;
        if (nee) { ro_b = a.m.p; }                                    // :510
        else { (void) pcg_f32(rng_b); ro_b = ln.ro + ln.d * a.m.t; }  // the null / real draw (:372), then :378-380
        const float rmaxt_b = ln.rmaxt - a.m.t;
        const float u = pcg_f32(rng_b);
        sample_interaction1_issue(S, G, mu_vol, inv_mu_vol, ro_b, nee ? ln.nd : ln.d, ln.id, rmaxt_b, u, b.m, b.T, b.b, spec);
        b.go = true;
    }
    vp_collide_post<BW>(S, G, ln, sk, a);
    if (b.go && ln.ev == EV_COLL) {
        // the first was a null collision (or a ratio-tracking step that continues): the lane is where the speculation
        // put it — same origin, same remaining distance — and its stream continues after the second distance draw
        ln.rng = rng_b;
        sk.cnt.coll += spec.coll;
        vp_collide_post<BW>(S, G, ln, sk, b);
    }
}

// the batched slow step (see the header comment)
template <bool BW, int DRT = 0>
__device__ __forceinline__ void vp_slow(const SceneK &S, const GridK &G, VpLane &ln, VpSink &sk) {
    if (ln.flags & VF_POSTNEE) {
        ln.flags &= ~VF_POSTNEE;
        ln.ev = EV_COLL;
        if (BW && (ln.flags & VF_ADJ)) vp_replay_done<DRT>(ln, sk);
        else vp_nee_done<BW, DRT>(S, G, ln, sk);
    } else {
        ln.flags &= ~VF_END;
        vp_end_pass<BW, DRT>(S, ln, sk);
    }
}

// pass 1 vertex: albedo gather, throughput/depth update (:150-213), vertex record, NEE set-up or continuation
template <bool BW, int DRT = 0>
__device__ __forceinline__ void vp_vertex1(const SceneK &S, const GridK &G, VpLane &ln, VpSink &sk) {
    const bool did_scatter = (ln.flags & VF_DID_SCATTER) != 0;
    if (did_scatter) ln.flags |= VF_HAS_SCAT;                         // :150
    V3 albedo = mk3(1, 1, 1);
    float st_vertex = 0.f;
    if (did_scatter) {
        V3 mei_p = ray_at(ln.o, ln.d, ln.mei_t);                      // :384
        Tri T = tri_setup(S, G, mei_p);
        albedo = lookup_albedo(G, T, 1);                              // :154
#if RCV_VERTEX_SIGMA
        if (BW) st_vertex = lookup_sigma1(S, G, T);                   // :386-388, for pass 2
#endif
        sk.cnt.scat++;
    }
    ln.thr = ln.thr * albedo;                                         // :206
    if (did_scatter) ln.depth += 1;                                   // :212
    bool active = did_scatter && (ln.depth < S.max_depth);            // :213
    const bool do_nee = S.use_nee && did_scatter && active;           // :219-220
    if (BW) {
        float4 *r = vp_record(sk, ln, ln.nvert);
        r[0] = make_float4(ln.o.x, ln.o.y, ln.o.z, did_scatter ? ln.mei_t : ln.ray_maxt);   // interval of :601
#if RCV_VERTEX_ALBEDO
        r[1] = make_float4(ln.d.x, ln.d.y, ln.d.z,
                           __uint_as_float((did_scatter ? (VR_SCATTER | VR_ALBEDO | (RCV_VERTEX_SIGMA ? (uint32_t) VR_SIGMA : 0u)) : 0u) |
                                           (do_nee ? VR_NEE : 0u)));
        if (did_scatter) r[3] = make_float4(0.f, albedo.x, albedo.y, albedo.z);
#if RCV_VERTEX_SIGMA
        if (did_scatter) r[2] = make_float4(st_vertex, 0.f, 0.f, 0.f);
#endif
#else
        r[1] = make_float4(ln.d.x, ln.d.y, ln.d.z, __uint_as_float((did_scatter ? VR_SCATTER : 0u) | (do_nee ? VR_NEE : 0u)));
#endif
        ln.nvert += 1;
    }
    if (do_nee) {                                                     // :393-446
        ln.saved = ln.rng.state;                                      // :396 sampler.clone()
        float s2x = pcg_f32(ln.rng), s2y = pcg_f32(ln.rng);           // :431
        ln.nd = emitter_sample_dir(S, s2x, s2y);
        ln.id = rcp3(ln.nd);
        ln.flags |= VF_IN_NEE;
        ln.running_t = __int_as_float(0);                             // shadow-ray collision records (vp_collide)
        ln.ev = EV_COLL;
        bool valid = true;
        if (S.emitter == 1) {                                         // :434-436 a sample of zero density is invalid
            float epdf;
            (void) emitter_sample_value(S, ln.nd, epdf);
            valid = epdf != 0.f;
            if (!valid) ln.tr_acc = 0.f;
        }
        if (!valid || !vp_nee_start(S, ln)) vp_nee_done<BW, DRT>(S, G, ln, sk);
        return;
    }
    vp_post_vertex<BW, DRT>(S, ln, sk, active);
}

// pass 2 vertex: the gradient terms of vertex vk, then the replay of its shadow ray
template <int DRT = 0>
__device__ __forceinline__ void vp_vertex2(const SceneK &S, const GridK &G, VpLane &ln, VpSink &sk) {
    if (DRT == 1) {
        // head of iteration vk as pass 1 ran it: Russian-roulette factor (:130-134; depth == vk), then
        // DRTReservoir.update with the throughput there (:546, :757-765)
        float xi = pcg_f32(ln.alt);
        if (ln.vk > S.rr_depth) {
            float q = fmin_(max3(ln.thr), 0.99f);
            ln.thr = ln.thr * (1.f / q);
        }
        V3 wsum = mk3(ln.last_pdf, __uint_as_float((uint32_t) ln.saved), __uint_as_float((uint32_t) (ln.saved >> 32))) + ln.thr;
        if (xi <= mean3(ln.thr / wsum)) {
            ln.depth = ln.vk;
            *vp_reservoir(sk, ln) = make_float4(ln.thr.x, ln.thr.y, ln.thr.z, 0.f);
        }
        ln.last_pdf = wsum.x;
        ln.saved = (uint64_t) __float_as_uint(wsum.y) | ((uint64_t) __float_as_uint(wsum.z) << 32);
    }
    const float4 *r = vp_record(sk, ln, ln.vk);
    float4 r0 = r[0], r1 = r[1];
    ln.o = mk3(r0.x, r0.y, r0.z); ln.d = mk3(r1.x, r1.y, r1.z);
    const float interval = r0.w;
    const uint32_t vf = __float_as_uint(r1.w);
    const bool did_scatter = (vf & VR_SCATTER) != 0;
    ln.mei_t = interval;
    if (did_scatter) {
        V3 mei_p = ray_at(ln.o, ln.d, interval);                      // :384
        Tri T = tri_setup(S, G, mei_p);
#if RCV_VERTEX_SIGMA
        float st = (vf & VR_SIGMA) ? r[2].x : lookup_sigma1(S, G, T);   // :386-388
#else
        float st = lookup_sigma1(S, G, T);                            // :386-388
#endif
#if RCV_VERTEX_ALBEDO
        V3 albedo;
        if (vf & VR_ALBEDO) { float4 r3 = r[3]; albedo = mk3(r3.y, r3.z, r3.w); }
        else albedo = lookup_albedo(G, T, 1);
#else
        V3 albedo = lookup_albedo(G, T, 1);                           // :154
#endif
        sk.cnt.scat++;
        if (DRT == 1) ln.thr = ln.thr * albedo;                       // :206, for the next iteration's reservoir update
        // scattering term (:165-185): backward_from(drt_mis_weight * dL * (sigma_t * albedo) * Li * inv_pdf)
        V3 inv_pdf = splat3(1.f / st);
        V3 Li = ln.result / vmax3(splat3(1e-8f), albedo);
        V3 c = (ln.dL * Li) * inv_pdf;
        if (DRT == 1) {                                               // :167-170 (the term is dropped without MIS, :165)
            float s2 = st * st;
            c = c * (s2 / (1.f + s2));
        }
        float g_st = (c.x * albedo.x + c.y * albedo.y) + c.z * albedo.z;
        V3 g_al = c * splat3(st);
        if (T.inside && (DRT != 1 || sk.drt_mis)) {
            float sg = S.scale * g_st;
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
    }
    // transmittance term (:194-202, :597-620): 4 uniform points on the segment
    {
        V3 aw = ln.dL * ln.result;                                    // :198
        float inv_pdf = interval / 4.f;                               // :617
        float g = -((aw.x * inv_pdf) + (aw.y * inv_pdf) + (aw.z * inv_pdf));
#pragma unroll kTransUnroll
        for (int j = 0; j < 4; ++j) {
            float t = pcg_f32(ln.alt) * interval;                     // :608
            scatter_sigma_t(S, G, sk.D, ray_at(ln.o, ln.d, t), g, sk.cnt);
        }
    }
    if (vf & VR_NEE_DONE) {                                           // the shadow ray's gradients were scattered in pass 1
        float4 r2 = r[2], r3 = r[3];
        ln.nee = mk3(r2.z, r2.w, r3.x);
        ln.result = ln.result - ln.nee;                               // :227
        vp_next_vertex<DRT>(ln, sk);
        return;
    }
    if (vf & VR_NEE) {                                                // :406-414: replay with the cloned sampler
        float4 r2 = r[2], r3 = r[3];
        ln.rng.state = (uint64_t) __float_as_uint(r2.x) | ((uint64_t) __float_as_uint(r2.y) << 32);
        (void) pcg_u32(ln.rng); (void) pcg_u32(ln.rng);               // the direction sample (:431), same values
        ln.nee = mk3(r2.z, r2.w, r3.x);
        ln.nd = mk3(r3.y, r3.z, r3.w);
        ln.id = rcp3(ln.nd);
        ln.flags |= VF_IN_NEE | VF_REPLAY;
        ln.ev = EV_COLL;
        if (!vp_nee_start(S, ln)) vp_replay_done<DRT>(ln, sk);
        return;
    }
    vp_next_vertex<DRT>(ln, sk);
}

// PCG32 skip-ahead (O'Neill's pcg_advance_lcg_64): the state after `delta` draws
__host__ __device__ __forceinline__ void pcg_advance(Pcg &r, uint64_t delta) {
    uint64_t cur_mult = 0x5851f42d4c957f2dULL, cur_plus = r.inc, acc_mult = 1u, acc_plus = 0u;
    while (delta > 0) {
        if (delta & 1u) { acc_mult *= cur_mult; acc_plus = acc_plus * cur_mult + cur_plus; }
        cur_plus = (cur_mult + 1u) * cur_plus;
        cur_mult *= cur_mult;
        delta >>= 1;
    }
    r.state = acc_mult * r.state + acc_plus;
}

// Pass 2 of ONE vertex, outside the lane machine (RCV_FLAT_PASS2, use_drt = False): vertex k of the lane whose records
// start at record `vbase`, global lane index idx, with res0 = the lane's running result before vertex k0 <= k.  Exactly
// what vp_vertex2 computes when the lane reaches vertex k: the running result is rebuilt by subtracting the shadow-ray
// contributions of vertices k0 .. k-1 in order (:227), the secondary sampler is skipped ahead to its 4 k-th draw (four
// draws per vertex, :608).  Returns the running result AFTER vertex k.  Every shadow ray of such a lane had its
// gradients scattered in pass 1 (VR_NEE_DONE): lanes with an overflowed ray walk their list in the machine instead.
__device__ __forceinline__ V3 vp_vertex2_flat(const SceneK &S, const GridK &G, VpSink &sk, uint32_t vbase, int k0, int k, V3 res0,
                                              V3 dL, uint32_t idx) {
    V3 result = res0;
    for (int j = k0; j < k; ++j) {
        const float4 *rj = sk.arena + ((size_t) vbase + (size_t) j) * 4;
        if (__float_as_uint(rj[1].w) & VR_NEE) {
            float4 a = rj[2], b = rj[3];
            result = result - mk3(a.z, a.w, b.x);
        }
    }
    Pcg alt;
    sampler_seed(alt, sk.alt_seed, idx);
    pcg_advance(alt, 4u * (uint64_t) k);
    const float4 *r = sk.arena + ((size_t) vbase + (size_t) k) * 4;
    float4 r0 = r[0], r1 = r[1];
    const V3 o = mk3(r0.x, r0.y, r0.z), d = mk3(r1.x, r1.y, r1.z);
    const float interval = r0.w;
    const uint32_t vf = __float_as_uint(r1.w);
    if (vf & VR_SCATTER) {
        V3 mei_p = ray_at(o, d, interval);                            // :384
        Tri T = tri_setup(S, G, mei_p);
#if RCV_VERTEX_SIGMA
        float st = (vf & VR_SIGMA) ? r[2].x : lookup_sigma1(S, G, T);   // :386-388
#else
        float st = lookup_sigma1(S, G, T);                            // :386-388
#endif
#if RCV_VERTEX_ALBEDO
        V3 albedo;
        if (vf & VR_ALBEDO) { float4 r3 = r[3]; albedo = mk3(r3.y, r3.z, r3.w); }
        else albedo = lookup_albedo(G, T, 1);
#else
        V3 albedo = lookup_albedo(G, T, 1);                           // :154
#endif
        sk.cnt.scat++;
        V3 inv_pdf = splat3(1.f / st);                                // scattering term (:165-185)
        V3 Li = result / vmax3(splat3(1e-8f), albedo);
        V3 c = (dL * Li) * inv_pdf;
        float g_st = (c.x * albedo.x + c.y * albedo.y) + c.z * albedo.z;
        V3 g_al = c * splat3(st);
        if (T.inside) {
            float sg = S.scale * g_st;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float w = tri_weight(T, q);
                uint32_t v = tri_voxel(G, T, q);
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
    }
    {                                                                 // transmittance term (:194-202, :597-620)
        V3 aw = dL * result;                                          // :198
        float inv_pdf = interval / 4.f;                               // :617
        float g = -((aw.x * inv_pdf) + (aw.y * inv_pdf) + (aw.z * inv_pdf));
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            float t = pcg_f32(alt) * interval;                        // :608
            scatter_sigma_t(S, G, sk.D, ray_at(o, d, t), g, sk.cnt);
        }
    }
    if (vf & VR_NEE) {                                                // :227 (its gradients were scattered in pass 1)
        float4 r2 = r[2], r3 = r[3];
        result = result - mk3(r2.z, r2.w, r3.x);
    }
    return result;
}

template <bool BW, int DRT = 0>
__device__ __forceinline__ void vp_vertex(const SceneK &S, const GridK &G, VpLane &ln, VpSink &sk) {
    if (BW && (ln.flags & VF_ADJ)) vp_vertex2<DRT>(S, G, ln, sk);
    else vp_vertex1<BW, DRT>(S, G, ln, sk);
}

// ---------------------------------------------------------------------------
// DRT stages (see the header comment).  Job record at index i of a PathQueue:
//   r0 = secondary sampler (state lo, state hi, inc >> 1), depth     r1 = ray origin, adjoint.x
//   r2 = ray direction, adjoint.y                                     r3 = adjoint.z
// ---------------------------------------------------------------------------
// DRT = 1, a lane whose pass 2 just ended (VF_JOB): DRTReservoir.get (:768-772) and the delayed estimate's inputs
// (:262-272).  false: the reservoir is empty, nothing to estimate.
__device__ __forceinline__ bool vp_drt_job(const VpLane &ln, const VpSink &sk, float4 &j0, float4 &j1, float4 &j2, float4 &j3) {
    const int sel = ln.depth;
    if (sel < 0) return false;
    float4 b = *vp_reservoir(sk, ln);
    V3 wsum = mk3(ln.last_pdf, __uint_as_float((uint32_t) ln.saved), __uint_as_float((uint32_t) (ln.saved >> 32)));
    V3 cur = mk3(b.x, b.y, b.z);
    float dn = mean3(cur);
    V3 wgt = (dn != 0.f) ? cur * (mean3(wsum) / dn) : mk3(0, 0, 0);
    V3 adjoint = wgt * ln.dL;
    const float4 *r = vp_record(sk, ln, sel);
    float4 r0 = r[0], r1 = r[1];
    j0 = make_float4(__uint_as_float((uint32_t) ln.alt.state), __uint_as_float((uint32_t) (ln.alt.state >> 32)),
                     __uint_as_float((uint32_t) (ln.alt.inc >> 1)), __int_as_float(sel));       // depth == vertex index
    j1 = make_float4(r0.x, r0.y, r0.z, adjoint.x);
    j2 = make_float4(r1.x, r1.y, r1.z, adjoint.y);
    j3 = make_float4(adjoint.z, 0.f, 0.f, 0.f);
    return true;
}

// backward_from(mis * drt_weight * adjoint * (sigma_t * albedo) * drt_Li) (:587-591) at p, c = mis * drt_weight * adjoint
template <class GRID>
__device__ __forceinline__ void vp_drt_scatter(const SceneK &S, const GRID &G, const GradK &D, V3 p, V3 c, V3 Li, VpsCounts &cnt) {
    Tri T = tri_setup(S, G, p);
    float st = lookup_sigma1(S, G, T);
    V3 albedo = lookup_albedo(G, T, 1);
    c = c * Li;
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

// The job between the stages (:551-578 up to the recursive call): distance sampling on the selected ray, the shadow
// ray and the phase sample of sample_recursive, all on the secondary sampler.  Returns 0 when there is nothing to
// scatter, 1 when the estimate is complete (Li holds it; the recursive path would start beyond max_depth), 2 when
// the recursive path is to be traced: rec[] is its path record (see vp_load<2>), p / c go with it.
template <class GRID>
__device__ __forceinline__ int vp_drt_sample(const SceneK &S, const GRID &G, const GradK &D, const float4 job[4], bool use_mis,
                                             float4 rec[4], V3 &p, V3 &c, V3 &Li, VpsCounts &cnt) {
    const float mu = __ldg(G.majorants + 1), inv_mu = 1.f / mu;
    Pcg alt;
    alt.state = (uint64_t) __float_as_uint(job[0].x) | ((uint64_t) __float_as_uint(job[0].y) << 32);
    alt.inc = ((uint64_t) __float_as_uint(job[0].z) << 1) | 1u;
    const int depth = __float_as_int(job[0].w);
    V3 o = mk3(job[1].x, job[1].y, job[1].z), d = mk3(job[2].x, job[2].y, job[2].z);
    V3 adjoint = mk3(job[1].w, job[2].w, job[3].x);
    V3 id = rcp3(d);
    Hit si = box_hit(S, o, d, id, RCV_LARGEST);                       // the ray's own boundary distance (:553-555)
    float ray_maxt = isfinite(si.t) ? si.t : RCV_LARGEST;
    float sel_t;
    float drt_weight = sample_interaction_drt(S, G, mu, inv_mu, alt, o, d, id, ray_maxt, true, sel_t, cnt);   // :560-561
    if (!(sel_t < inf_())) return 0;                                  // :568
    p = ray_at(o, d, sel_t);
    Tri T = tri_setup(S, G, p);
    float st = lookup_sigma1(S, G, T);
    float mis = use_mis ? (1.f / (1.f + st * st)) : 1.f;              // :581-585
    c = adjoint * (mis * drt_weight);
    // sample_recursive (:623-668) up to the recursive call
    V3 wi = mk3(-d.x, -d.y, -d.z);
    Li = mk3(0, 0, 0);
    if (S.use_nee) {                                                  // :635-638, throughput 1
        V3 dir;
        float tr = vps_sample_emitter<false>(S, G, D, mu, inv_mu, alt, p, mk3(0, 0, 0), dir, cnt);
        float epdf;
        V3 emitted = emitter_sample_value(S, dir, epdf) * splat3(tr);
        float phase_val = phase_eval(S, wi, dir);
        Li = Li + ((mk3(1, 1, 1) * phase_val) * mis_weight(epdf, phase_val)) * emitted;
    }
    (void) pcg_f32(alt);                                              // :644-646
    float s2x = pcg_f32(alt), s2y = pcg_f32(alt);
    float pdf;
    V3 wo = phase_sample(S, wi, s2x, s2y, pdf);
    if (!(depth + 1 < S.max_depth)) return 1;                         // :660: the recursive call starts inactive
    (void) pcg_f32(alt);                                              // :112 of the recursive call (its alt_seed_rnd draw)
    Hit sn = box_hit(S, p, wo, rcp3(wo), RCV_LARGEST);                // :651
    rec[0] = make_float4(__uint_as_float((uint32_t) alt.state), __uint_as_float((uint32_t) (alt.state >> 32)),
                         __uint_as_float((uint32_t) (alt.inc >> 1)), 0.f);      // .w: the job's index, set by the caller
    rec[1] = make_float4(p.x, p.y, p.z, isfinite(sn.t) ? sn.t : RCV_LARGEST);   // :653-654
    rec[2] = make_float4(wo.x, wo.y, wo.z, __uint_as_float((uint32_t) (depth + 1)));
    rec[3] = make_float4(Li.x, Li.y, Li.z, pdf);
    return 2;
}

}  // namespace rcv
