// rcv_sched.cuh — scheduling helpers shared by the tracing translation units (rcv_trace.cu, rcv_surface.cu): lane
// indexing of an spp shard, warp-aggregated queue appends, the event vote of the WAVEFRONT schedule, warp-private
// chunks of work, and the grid size of a persistent kernel.
#pragma once
#include "rcv_internal.h"
#include "rcv_twostate_wf.cuh"

namespace rcv {

#ifndef RCV_EVENT_BATCH
#define RCV_EVENT_BATCH 12
#endif
static constexpr int kEventBatch = RCV_EVENT_BATCH;    // run the vertex / slow step once this many lanes wait for it
static constexpr unsigned long long kChunk = 256;      // records a warp takes from the queue per atomic

__device__ __forceinline__ uint64_t global_lane(uint64_t j, uint32_t spp, uint32_t spp_begin, uint32_t ns) {
    uint64_t pixel = j / ns;
    uint32_t k = spp_begin + (uint32_t) (j - pixel * ns);
    return pixel * spp + k;
}

// warp-aggregated append to the record queue
__device__ __forceinline__ unsigned long long queue_slot(unsigned long long *count, bool want) {
    unsigned m = __ballot_sync(0xffffffffu, want);
    if (m == 0) return 0;
    const unsigned lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if ((int) lane == leader) base = atomicAdd(count, (unsigned long long) __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + __popc(m & ((1u << lane) - 1u));
}

// Which handler a warp runs this round.  The collision step is the common case;
// the vertex step and the slow step run once kEventBatch lanes wait for them (or
// when no lane is ready to collide), so that they too execute with many lanes.
__device__ __forceinline__ int pick_event(int n_coll, int n_scat, int n_slow) {
    if (n_slow >= kEventBatch || n_scat >= kEventBatch) return (n_slow >= n_scat) ? EV_SLOW : EV_SCAT;
    if (n_coll > 0) return EV_COLL;
    return (n_slow >= n_scat) ? EV_SLOW : EV_SCAT;
}

// warp-private chunk of queue positions
struct ChunkCursor {
    unsigned long long next, end;
    bool more;
};
// all lanes call; the lanes of m_need receive distinct queue positions while work lasts, the others ~0
__device__ __forceinline__ unsigned long long take_positions(ChunkCursor &cc, unsigned long long *work, unsigned long long n_total,
                                                             unsigned m_need) {
    const unsigned lane = threadIdx.x & 31;
    unsigned long long mine = ~0ull;
    while (m_need && cc.more) {
        if (cc.next >= cc.end) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(work, kChunk);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base >= n_total) { cc.more = false; break; }
            cc.next = base;
            cc.end = (base + kChunk < n_total) ? base + kChunk : n_total;
        }
        unsigned long long avail = cc.end - cc.next;
        unsigned rank = __popc(m_need & ((1u << lane) - 1u));
        bool need = (m_need >> lane) & 1u;
        bool got = need && rank < avail;
        if (got) mine = cc.next + rank;
        unsigned n_need = __popc(m_need);
        cc.next += (n_need < avail) ? n_need : avail;
        m_need &= ~__ballot_sync(0xffffffffu, got);
    }
    return mine;
}

template <typename K>
static cudaError_t persistent_grid(const LaunchCfg &cfg, K kernel, int block, uint64_t n_items, unsigned &blocks) {
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    uint64_t want = (n_items + block - 1) / block;
    uint64_t b = (uint64_t) per_sm * cfg.sm_count;
    if (b > want) b = want;
    blocks = (unsigned) (b < 1 ? 1 : b);
    return cudaSuccess;
}

}  // namespace rcv
