// gather_bw.cu — random-granule read bandwidth of HBM on this GPU (the roofline of the brick gathers).
// Every thread reads ILP independent GRANULE-byte granules per iteration at pseudo-random granule indices of a
// `GiB`-sized array (default 1 GiB >> L2).
//   usage: gather_bw [GiB] [l2_fetch_granularity_limit_bytes (0 = leave)]
//   build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_bw gather_bw.cu
// Load kinds: 0 ld.global.nc (__ldg), 1 ld.global.cg, 2 ld.global.nc.L2::64B, 3 ld.global.nc.L2::128B (prefetch-size hints)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

template <int KIND>
__device__ __forceinline__ float4 ld16(const float4 *p) {
    float4 v;
    if (KIND == 0) v = __ldg(p);
    else if (KIND == 1) v = __ldcg(p);
    else if (KIND == 2) asm volatile("ld.global.nc.L2::64B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    else asm volatile("ld.global.nc.L2::128B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

template <int G16, int ILP, int KIND>   // G16 = granule size in 16-byte units
__global__ void k_gather(const float4 *__restrict__ a, uint64_t n_gran, int iters, float *sink) {
    uint64_t s = (uint64_t) (blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345u;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        float4 v[ILP][G16];
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            uint64_t g = (uint64_t) (((s >> 32) * n_gran) >> 32);
#pragma unroll
            for (int q = 0; q < G16; ++q) v[j][q] = ld16<KIND>(a + g * G16 + q);
        }
#pragma unroll
        for (int j = 0; j < ILP; ++j)
#pragma unroll
            for (int q = 0; q < G16; ++q) acc += v[j][q].x + v[j][q].w;
    }
    if (acc == 123.456f) *sink = acc;
}

template <int G16, int ILP, int KIND>
static void run(const float4 *a, size_t bytes, float *sink, int sms) {
    uint64_t n_gran = bytes / (16 * G16);
    int iters = 64;
    int blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_gather<G16, ILP, KIND><<<blocks, threads>>>(a, n_gran, 4, sink);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        k_gather<G16, ILP, KIND><<<blocks, threads>>>(a, n_gran, iters, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double gb = (double) blocks * threads * iters * ILP * G16 * 16 / 1e9;
    printf("{\"granule_bytes\": %d, \"ilp\": %d, \"load\": %d, \"array_GiB\": %.2f, \"ms\": %.3f, \"useful_GBps\": %.1f, \"Ggranules_per_s\": %.2f}\n",
           G16 * 16, ILP, KIND, bytes / 1073741824.0, best, gb / (best * 1e-3), gb / (best * 1e-3) / (G16 * 16));
}

// One 256-bit load per 32-byte granule (LDG.256 + L2::64B, what the tracing kernels issue for a brick), ILP granules in
// flight per thread.
template <int ILP>
__global__ void k_gather256(const float4 *__restrict__ a, uint64_t n_gran, int iters, float *sink) {
    uint64_t s = (uint64_t) (blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345u;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        float4 v[ILP][2];
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            uint64_t g = (uint64_t) (((s >> 32) * n_gran) >> 32);
            const float4 *p = a + g * 2;
            asm volatile("ld.global.nc.L2::64B.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=f"(v[j][0].x), "=f"(v[j][0].y), "=f"(v[j][0].z), "=f"(v[j][0].w), "=f"(v[j][1].x), "=f"(v[j][1].y),
                           "=f"(v[j][1].z), "=f"(v[j][1].w) : "l"(p));
        }
#pragma unroll
        for (int j = 0; j < ILP; ++j) acc += v[j][0].x + v[j][1].w;
    }
    if (acc == 123.456f) *sink = acc;
}

// random 32-byte granules of an array of `mib` MiB, read with LDG.256: L2-resident footprints give the L2 gather peak
// (the roofline of the L2-bound configuration C1), 1 GiB the HBM one
template <int ILP>
static double run256(const float4 *a, double mib, float *sink, int sms, int blocks_per_sm) {
    size_t bytes = (size_t) (mib * 1048576.0);
    uint64_t n_gran = bytes / 32;
    int iters = 256;
    int blocks = sms * blocks_per_sm, threads = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_gather256<ILP><<<blocks, threads>>>(a, n_gran, 64, sink);           // warm the L2
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        k_gather256<ILP><<<blocks, threads>>>(a, n_gran, iters, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double gb = (double) blocks * threads * iters * ILP * 32 / 1e9;
    double gbs = gb / (best * 1e-3);
    printf("{\"mode\": \"ldg256\", \"granule_bytes\": 32, \"ilp\": %d, \"blocks_per_sm\": %d, \"array_MiB\": %.0f, \"ms\": %.3f, "
           "\"useful_GBps\": %.1f, \"Ggranules_per_s\": %.2f}\n", ILP, blocks_per_sm, mib, best, gbs, gbs / 32);
    return gbs;
}

int main(int argc, char **argv) {
    if (argc > 1 && argv[1][0] == 'l') {        // gather_bw l2: the L2 / HBM gather peaks with the kernels' own load
        float4 *a; float *sink;
        size_t cap = (size_t) 1 << 30;
        cudaMalloc(&a, cap); cudaMemset(a, 0, cap); cudaMalloc(&sink, 4);
        cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
        int sms = p.multiProcessorCount;
        fprintf(stderr, "L2 %d MiB\n", p.l2CacheSize >> 20);
        double best_l2 = 0;
        for (double mib : {8.0, 16.0, 32.0, 48.0, 64.0, 96.0}) {
            double g1 = run256<4>(a, mib, sink, sms, 8), g2 = run256<8>(a, mib, sink, sms, 8), g3 = run256<16>(a, mib, sink, sms, 4);
            if (mib <= 48.0) { best_l2 = g1 > best_l2 ? g1 : best_l2; best_l2 = g2 > best_l2 ? g2 : best_l2; best_l2 = g3 > best_l2 ? g3 : best_l2; }
        }
        double h1 = run256<8>(a, 1024.0, sink, sms, 8), h2 = run256<16>(a, 1024.0, sink, sms, 4);
        printf("{\"l2_gather_gbs\": %.1f, \"hbm_gather_gbs\": %.1f, \"l2_MiB\": %d, \"source\": \"scripts/microbench/gather_bw.cu l2: "
               "random 32-byte granules, one LDG.256 (L2::64B) each, best of ILP 4/8/16, footprints <= 48 MiB (L2) and 1 GiB (HBM)\"}\n",
               best_l2, h1 > h2 ? h1 : h2, p.l2CacheSize >> 20);
        return 0;
    }
    size_t bytes = (argc > 1 ? atof(argv[1]) : 1.0) * 1073741824.0;
    if (argc > 2 && atoi(argv[2]) > 0) {
        cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t) atoi(argv[2]));
        size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
        fprintf(stderr, "L2 fetch granularity limit now %zu\n", g);
    }
    float4 *a; float *sink;
    cudaMalloc(&a, bytes); cudaMemset(a, 0, bytes); cudaMalloc(&sink, 4);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    run<2, 4, 0>(a, bytes, sink, sms);
    run<2, 4, 1>(a, bytes, sink, sms);
    run<2, 4, 2>(a, bytes, sink, sms);
    run<2, 4, 3>(a, bytes, sink, sms);
    run<4, 4, 0>(a, bytes, sink, sms);
    run<4, 4, 1>(a, bytes, sink, sms);
    run<4, 4, 2>(a, bytes, sink, sms);
    run<8, 4, 0>(a, bytes, sink, sms);
    run<16, 2, 0>(a, bytes, sink, sms);
    return 0;
}
