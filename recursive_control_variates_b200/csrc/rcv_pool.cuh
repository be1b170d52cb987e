// rcv_pool.cuh — warp-private lane pools in shared memory (POOL schedule).
//
// A warp owns kPoolSlots = 64 lane states, stored structure-of-arrays in shared
// memory (word w of slot s at pool[w * kPoolSlots + s]: consecutive slots hit consecutive
// banks).  Every round the warp picks ONE event class and up to 32 slots waiting
// for it, loads those lanes into registers, runs the handler of rcv_*_wf.cuh and
// stores them back.  With twice as many lanes as threads the tentative-collision
// step almost always finds 32 ready lanes, and the rarer vertex / slow steps wait
// until they too fill most of the warp — without any cross-warp synchronisation.
// The collision step moves only the words it reads / writes.
#pragma once
#include "rcv_twostate_wf.cuh"
#include "rcv_volpathsimple_wf.cuh"

namespace rcv {

static constexpr int kPoolSlots = 64;
static constexpr int kPoolMaskWords = kPoolSlots / 32;
static_assert(kPoolSlots <= 256, "slot indices are bytes");

__device__ __forceinline__ float pf_u(uint32_t v) { return __uint_as_float(v); }
__device__ __forceinline__ uint32_t pu_f(float v) { return __float_as_uint(v); }
#define PW(w) pw[(w) * kPoolSlots + s]

// ---------------------------------------------------------------------------
// TsLane <-> pool.  Word map:
//  0 pix  1-2 rng.state  3-4 rng.inc  5-7 o  8-10 d  11-13 id  14-16 thr_old  17-19 thr_new
//  20 si_t  21 depth  22 flags  23-25 no  26-28 nd  29 total  30 tr_old  31 tr_new
//  32-34 L_old  35-37 L_new  38 last_pdf  39 phase_val  40 ratio  41-42 film position
// ---------------------------------------------------------------------------
static constexpr int kTsWords = 43;

__device__ __forceinline__ void ts_pool_load_hot(const float *pw, int s, TsLane &ln) {
    ln.rng.state = (uint64_t) pu_f(PW(1)) | ((uint64_t) pu_f(PW(2)) << 32);
    ln.rng.inc = (uint64_t) pu_f(PW(3)) | ((uint64_t) pu_f(PW(4)) << 32);
    ln.o = mk3(PW(5), PW(6), PW(7)); ln.d = mk3(PW(8), PW(9), PW(10)); ln.id = mk3(PW(11), PW(12), PW(13));
    ln.thr_old = mk3(PW(14), PW(15), PW(16)); ln.thr_new = mk3(PW(17), PW(18), PW(19));
    ln.si_t = PW(20); ln.depth = (int) pu_f(PW(21)); ln.flags = pu_f(PW(22));
    ln.no = mk3(PW(23), PW(24), PW(25)); ln.nd = mk3(PW(26), PW(27), PW(28));
    ln.total = PW(29); ln.tr_old = PW(30); ln.tr_new = PW(31);
}
__device__ __forceinline__ void ts_pool_load_cold(const float *pw, int s, TsLane &ln) {
    ln.pix = (int) pu_f(PW(0));
    ln.L_old = mk3(PW(32), PW(33), PW(34)); ln.L_new = mk3(PW(35), PW(36), PW(37));
    ln.last_pdf = PW(38); ln.phase_val = PW(39); ln.ratio = PW(40);
    ln.fx = PW(41); ln.fy = PW(42);
}
// what ts_collide can modify
__device__ __forceinline__ void ts_pool_store_coll(float *pw, int s, const TsLane &ln) {
    PW(1) = pf_u((uint32_t) ln.rng.state); PW(2) = pf_u((uint32_t) (ln.rng.state >> 32));
    PW(5) = ln.o.x; PW(6) = ln.o.y; PW(7) = ln.o.z;
    PW(14) = ln.thr_old.x; PW(15) = ln.thr_old.y; PW(16) = ln.thr_old.z;
    PW(17) = ln.thr_new.x; PW(18) = ln.thr_new.y; PW(19) = ln.thr_new.z;
    PW(20) = ln.si_t; PW(21) = pf_u((uint32_t) ln.depth); PW(22) = pf_u(ln.flags);
    PW(23) = ln.no.x; PW(24) = ln.no.y; PW(25) = ln.no.z;
    PW(29) = ln.total; PW(30) = ln.tr_old; PW(31) = ln.tr_new;
}
__device__ __forceinline__ void ts_pool_store_all(float *pw, int s, const TsLane &ln) {
    ts_pool_store_coll(pw, s, ln);
    PW(0) = pf_u((uint32_t) ln.pix);
    PW(3) = pf_u((uint32_t) ln.rng.inc); PW(4) = pf_u((uint32_t) (ln.rng.inc >> 32));
    PW(8) = ln.d.x; PW(9) = ln.d.y; PW(10) = ln.d.z; PW(11) = ln.id.x; PW(12) = ln.id.y; PW(13) = ln.id.z;
    PW(26) = ln.nd.x; PW(27) = ln.nd.y; PW(28) = ln.nd.z;
    PW(32) = ln.L_old.x; PW(33) = ln.L_old.y; PW(34) = ln.L_old.z;
    PW(35) = ln.L_new.x; PW(36) = ln.L_new.y; PW(37) = ln.L_new.z;
    PW(38) = ln.last_pdf; PW(39) = ln.phase_val; PW(40) = ln.ratio;
    PW(41) = ln.fx; PW(42) = ln.fy;
}

// ---------------------------------------------------------------------------
// VpLane <-> pool.  Word map:
//  0 pix  1 depth  2 idx  3 flags  4 vbase  5-6 rng.state  7-8 rng.inc  9-10 alt.state  11-12 alt.inc
//  13-14 saved  15-17 o  18-20 d  21-23 id  24 ray_maxt  25-27 thr  28-30 result  31-33 dL  34 last_pdf
//  35-37 ro  38 rmaxt  39 running_t  40 mei_t  41-43 nd  44 tr_acc  45-47 nee  48 nvert  49 vk
// ---------------------------------------------------------------------------
static constexpr int kVpWords = 50;

__device__ __forceinline__ void vp_pool_load_hot(const float *pw, int s, VpLane &ln) {
    ln.depth = (int) pu_f(PW(1)); ln.flags = pu_f(PW(3));
    ln.rng.state = (uint64_t) pu_f(PW(5)) | ((uint64_t) pu_f(PW(6)) << 32);
    ln.rng.inc = (uint64_t) pu_f(PW(7)) | ((uint64_t) pu_f(PW(8)) << 32);
    ln.o = mk3(PW(15), PW(16), PW(17)); ln.d = mk3(PW(18), PW(19), PW(20)); ln.id = mk3(PW(21), PW(22), PW(23));
    ln.ray_maxt = PW(24);
    ln.thr = mk3(PW(25), PW(26), PW(27));
    ln.dL = mk3(PW(31), PW(32), PW(33));
    ln.ro = mk3(PW(35), PW(36), PW(37)); ln.rmaxt = PW(38); ln.running_t = PW(39); ln.mei_t = PW(40);
    ln.nd = mk3(PW(41), PW(42), PW(43)); ln.tr_acc = PW(44);
    ln.nee = mk3(PW(45), PW(46), PW(47));
}
__device__ __forceinline__ void vp_pool_load_cold(const float *pw, int s, VpLane &ln) {
    ln.pix = (int) pu_f(PW(0)); ln.idx = pu_f(PW(2)); ln.vbase = pu_f(PW(4));
    ln.nvert = (int) pu_f(PW(48)); ln.vk = (int) pu_f(PW(49));
    ln.alt.state = (uint64_t) pu_f(PW(9)) | ((uint64_t) pu_f(PW(10)) << 32);
    ln.alt.inc = (uint64_t) pu_f(PW(11)) | ((uint64_t) pu_f(PW(12)) << 32);
    ln.saved = (uint64_t) pu_f(PW(13)) | ((uint64_t) pu_f(PW(14)) << 32);
    ln.result = mk3(PW(28), PW(29), PW(30));
    ln.last_pdf = PW(34);
}
// what vp_collide can modify
__device__ __forceinline__ void vp_pool_store_coll(float *pw, int s, const VpLane &ln) {
    PW(3) = pf_u(ln.flags);
    PW(5) = pf_u((uint32_t) ln.rng.state); PW(6) = pf_u((uint32_t) (ln.rng.state >> 32));
    PW(25) = ln.thr.x; PW(26) = ln.thr.y; PW(27) = ln.thr.z;
    PW(35) = ln.ro.x; PW(36) = ln.ro.y; PW(37) = ln.ro.z; PW(38) = ln.rmaxt; PW(39) = ln.running_t; PW(40) = ln.mei_t;
    PW(44) = ln.tr_acc;
}
__device__ __forceinline__ void vp_pool_store_all(float *pw, int s, const VpLane &ln) {
    vp_pool_store_coll(pw, s, ln);
    PW(0) = pf_u((uint32_t) ln.pix); PW(1) = pf_u((uint32_t) ln.depth); PW(2) = pf_u(ln.idx); PW(4) = pf_u(ln.vbase);
    PW(48) = pf_u((uint32_t) ln.nvert); PW(49) = pf_u((uint32_t) ln.vk);
    PW(7) = pf_u((uint32_t) ln.rng.inc); PW(8) = pf_u((uint32_t) (ln.rng.inc >> 32));
    PW(9) = pf_u((uint32_t) ln.alt.state); PW(10) = pf_u((uint32_t) (ln.alt.state >> 32));
    PW(11) = pf_u((uint32_t) ln.alt.inc); PW(12) = pf_u((uint32_t) (ln.alt.inc >> 32));
    PW(13) = pf_u((uint32_t) ln.saved); PW(14) = pf_u((uint32_t) (ln.saved >> 32));
    PW(15) = ln.o.x; PW(16) = ln.o.y; PW(17) = ln.o.z; PW(18) = ln.d.x; PW(19) = ln.d.y; PW(20) = ln.d.z;
    PW(21) = ln.id.x; PW(22) = ln.id.y; PW(23) = ln.id.z; PW(24) = ln.ray_maxt;
    PW(28) = ln.result.x; PW(29) = ln.result.y; PW(30) = ln.result.z;
    PW(31) = ln.dL.x; PW(32) = ln.dL.y; PW(33) = ln.dL.z; PW(34) = ln.last_pdf;
    PW(41) = ln.nd.x; PW(42) = ln.nd.y; PW(43) = ln.nd.z;
    PW(45) = ln.nee.x; PW(46) = ln.nee.y; PW(47) = ln.nee.z;
}
#undef PW

// ---------------------------------------------------------------------------
// slot selection: list[r] = the r-th lowest slot whose class bit is set (m[j] covers slots 32 j .. 32 j + 31); returns
// their number.  `list` is a kPoolSlots-byte warp-private scratch; thread k then reads list[k], list[k + 32], ...
// ---------------------------------------------------------------------------
__device__ __forceinline__ int pool_list(const unsigned (&m)[kPoolMaskWords], unsigned char *list) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    int base = 0;
#pragma unroll
    for (int j = 0; j < kPoolMaskWords; ++j) {
        if ((m[j] >> lane) & 1u) list[base + __popc(m[j] & lt)] = (unsigned char) (lane + 32 * j);
        base += __popc(m[j]);
    }
    __syncwarp();
    return base;
}
// the slot of this thread in batch b of a round over the first n listed slots (or -1)
__device__ __forceinline__ int pool_slot(const unsigned char *list, int n, int b) {
    const int r = (int) (threadIdx.x & 31) + 32 * b;
    return (r < n) ? (int) list[r] : -1;
}

// which class to run: collisions when a full round of them is ready, otherwise the fullest queue
__device__ __forceinline__ int pool_pick(int n_coll, int n_scat, int n_slow) {
    if (n_coll >= 32) return EV_COLL;
    if (n_coll >= n_scat && n_coll >= n_slow && n_coll > 0) return EV_COLL;
    return (n_slow >= n_scat) ? EV_SLOW : EV_SCAT;
}

}  // namespace rcv
