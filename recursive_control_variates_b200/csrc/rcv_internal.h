// rcv_internal.h — launch wrappers shared between the kernel translation units
// and the C-ABI (rcv_api.cu).  Not part of the public interface.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "rcv_device.cuh"

namespace rcv {

struct LaunchCfg {
    cudaStream_t stream;
    int sm_count;
    int tracer;        // RCV_TRACER_*
    int drt_flags;     // 0: use_drt = False; else bit 0 use_drt, bit 1 use_drt_subsampling, bit 2 use_drt_mis
    Counters *counters;   // device
    unsigned long long *work_counter;   // device, persistent-kernel work queue head
    PathQueue queue;                    // device, path records of the wavefront schedule
    PathQueue jobs;                     // device, DRT job records (use_drt on the pool tracer), else null
    float4 *arena;                      // device, vertex records of the adjoint (4 float4 each)
    uint32_t vcap;                      // records per arena slot
    uint64_t *launches;   // host-side launch counter
};

// ---- tracing (rcv_trace.cu)
// mode 0: two states -> acc_old & acc_new; mode 1 / 2: single state old / new -> acc_old / acc_new
cudaError_t launch_twostates(const LaunchCfg &cfg, const SceneK &S, const GridKS &G, int mode, uint32_t seed, uint32_t spp,
                             uint32_t spp_begin, uint32_t spp_end, float4 *acc_old, float4 *acc_new);
cudaError_t launch_vps_primal(const LaunchCfg &cfg, const SceneK &S, const GridKS &G, uint32_t seed, uint32_t spp,
                              uint32_t spp_begin, uint32_t spp_end, float4 *accum);
cudaError_t launch_vps_backward(const LaunchCfg &cfg, const SceneK &S, const GridKS &G, const GradK &D, uint32_t seed_grad,
                                uint32_t spp_grad, uint32_t spp_begin, uint32_t spp_end, uint32_t alt_seed,
                                const float *grad_img, const float *film_w);
size_t vps_backward_slots(const LaunchCfg &cfg);
// ---- scenes with rectangles, and the adjoint inherited from upstream prbvolpath (rcv_surface.cu)
cudaError_t launch_surface_primal(const LaunchCfg &cfg, const SceneK &S, const GridKS &G, int mode, uint32_t seed, uint32_t spp,
                                  uint32_t spp_begin, uint32_t spp_end, float4 *acc_old, float4 *acc_new);
cudaError_t launch_prb_backward(const LaunchCfg &cfg, const SceneK &S, const GridKS &G, const GradK &D, uint32_t seed_grad,
                                uint32_t spp_grad, uint32_t spp_begin, uint32_t spp_end, const float *grad_img, const float *film_w);
cudaError_t launch_film_weights(const LaunchCfg &cfg, const SceneK &S, uint32_t seed, uint32_t spp, uint32_t spp_begin,
                                uint32_t spp_end, float *film_w, bool zero);

// ---- medium state (rcv_pixel.cu)
cudaError_t launch_rotate(const LaunchCfg &cfg, const float *sigma_in, const float *albedo_in, int nz, int ny, int nx, float scale,
                          float2 *sig2, float4 *blk2, float4 *blk1, float4 *alb2, float *block_max, int n_blocks,
                          float *majorants);
cudaError_t launch_majorant_grid(const LaunchCfg &cfg, const float2 *sig2, int nz, int ny, int nx, int factor, float scale,
                                 float2 *mgrid);
cudaError_t launch_medium_read(const LaunchCfg &cfg, const float2 *sig2, const float4 *alb2, size_t nvox, int state,
                               float *sigma_out, float *albedo_out);

// ---- per-pixel / per-voxel elementwise work (rcv_pixel.cu)
cudaError_t launch_develop(const LaunchCfg &cfg, const float4 *accum, float *img, size_t npix);
struct CvArgs {
    const float *H, *F;
    float *img, *v_n, *w_s, *mean_x, *mean_y, *var_x, *var_y, *cov;
    uint64_t n;
    int do_weights;      // it > warmup
    int do_update;       // it > 0
    int pre_update;
    int stats_n_is_zero; // stats.n == 0 before the update
    float c1a, c1b;      // 1 - beta1^n, 1 - beta1^(n+1)
    float c2;            // 1 - beta2^(n at get())
    float beta1, beta2;
};
cudaError_t launch_cv_update(const LaunchCfg &cfg, const CvArgs &a);
cudaError_t launch_loss_grad(const LaunchCfg &cfg, int kind, const float *img, const float *ref, uint64_t n, double *partial,
                             int n_partial, float *loss_out, float *grad_img);
cudaError_t launch_unbrick(const LaunchCfg &cfg, const GradK &D, int nx, int ny, int nz);
cudaError_t launch_grad_unpack(const LaunchCfg &cfg, const GradK &D, size_t nvox, float *g_sigma, float *g_albedo);
cudaError_t launch_adam_packed(const LaunchCfg &cfg, const GradK &D, size_t nvox, float *sigma, float *albedo, float *m_s,
                               float *v_s, float *m_a, float *v_a, float lr_t_s, float lr_t_a, float beta1, float beta2,
                               float eps, float slo, float shi, float alo, float ahi);
cudaError_t launch_adam_flat(const LaunchCfg &cfg, float *p, const float *g, float *m, float *v, uint64_t n, float lr_t,
                             float beta1, float beta2, float eps);
cudaError_t launch_clamp(const LaunchCfg &cfg, float *p, uint64_t n, float lo, float hi);
cudaError_t launch_upsample2x(const LaunchCfg &cfg, const float *in, float *out, int nz, int ny, int nx, int c);

}  // namespace rcv
