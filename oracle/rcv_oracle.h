/*
 * rcv_oracle.h — API of the CPU restatement (oracle).  TEST INFRASTRUCTURE
 * ONLY; see the header of rcv_oracle.c.  PARITY UNPINNED (no reference golden
 * vectors exist; the reference cannot be run here).
 */
#ifndef RCV_ORACLE_H
#define RCV_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#include "../include/rcv_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* one medium state in the reference layout */
typedef struct orc_medium {
    const float *sigma_t;   /* (nz,ny,nx,1) */
    const float *albedo;    /* (nz,ny,nx,3) */
    int32_t nz, ny, nx;
    float majorant;         /* scale * max(sigma_t) */
    /* majorant supergrid (heterogeneous `majorant_resolution_factor` > 0), NULL = one majorant for the volume */
    const float *majorant_grid;   /* (mz,my,mx), scale baked in */
    int32_t mz, my, mx;
} orc_medium;

void orc_tea32(uint32_t v0, uint32_t v1, uint32_t *o0, uint32_t *o1);
void orc_pcg32_kat(uint64_t initstate, uint64_t initseq, uint32_t *out, int n);
float orc_u32_to_f32(uint32_t v);
float orc_logf(float x);
float orc_expf(float x);
void orc_sincosf(float x, float *s, float *c);
float orc_majorant(const rcv_scene_desc *S, const float *sigma_t, size_t n);
/* local majorants of a (nz,ny,nx) grid on a supergrid with cells of `factor` voxels; out: (nz/f, ny/f, nx/f); -1 if not divisible */
int orc_majorant_grid(const rcv_scene_desc *S, const float *sigma_t, int nz, int ny, int nx, int factor, float *out);
float orc_dda_sample(const rcv_scene_desc *S, const orc_medium *Mo, const orc_medium *Mn, const float o[3], const float d[3],
                     float mint, float maxt, float desired_tau);
uint32_t orc_alt_seed(uint32_t seed_grad);
float orc_atan2f(float y, float x);
/* environment emitter test hooks: radiance + solid-angle density for a world direction; direction sample */
void orc_env_eval(const rcv_scene_desc *S, const float d[3], float rgb[3], float *pdf);
void orc_env_sample(const rcv_scene_desc *S, float u1, float u2, float d[3], float *pdf, float val[3]);
void orc_counters(uint64_t out[8], int reset);

int orc_twostates_lanes(const rcv_scene_desc *S, const orc_medium *Mo, const orc_medium *Mn,
                        uint32_t seed, uint32_t spp, uint64_t lane_begin, uint64_t lane_end,
                        float *L_old, float *L_new, int32_t *pix, int nthreads);
int orc_render_twostates(const rcv_scene_desc *S, const orc_medium *Mo, const orc_medium *Mn,
                         uint32_t seed, uint32_t spp, uint32_t spp_begin, uint32_t spp_end,
                         float *accum_old, float *accum_new, int nthreads);
void orc_develop(const float *accum, float *img, uint64_t npix);
void orc_film_weights(const rcv_scene_desc *S, uint32_t seed, uint32_t spp, float *Wt);
void orc_film_weights_range(const rcv_scene_desc *S, uint32_t seed, uint32_t spp, uint32_t spp_begin, uint32_t spp_end, float *Wt);

void orc_cv_update(const float *H, const float *F, float *img, float *v_n, float *w_s,
                   float *mean_x, float *mean_y, float *var_x, float *var_y, float *cov,
                   uint64_t n, int it, int warmup, int stats_n, float beta1, float beta2,
                   int pre_update);
double orc_loss_grad(int kind, const float *img, const float *ref, uint64_t n, float *grad);

int orc_vps_lanes(const rcv_scene_desc *S, const orc_medium *M, uint32_t seed, uint32_t spp,
                  uint64_t lane_begin, uint64_t lane_end, float *L, int32_t *pix, int nthreads);
int orc_vps_render(const rcv_scene_desc *S, const orc_medium *M, uint32_t seed, uint32_t spp,
                   uint32_t spp_begin, uint32_t spp_end, float *accum, int nthreads);
int orc_vps_backward(const rcv_scene_desc *S, const orc_medium *M, uint32_t seed_grad, uint32_t spp_grad,
                     uint32_t spp_begin, uint32_t spp_end, const float *grad_img, const float *film_w,
                     double *grad_sigma_t, double *grad_albedo, int nthreads);
/* render_backward with upstream prbvolpath (the adjoint twostateprbvolpath inherits): surfaces + area light */
int orc_prb_backward(const rcv_scene_desc *S, const orc_medium *M, uint32_t seed_grad, uint32_t spp_grad,
                     uint32_t spp_begin, uint32_t spp_end, const float *grad_img, const float *film_w,
                     double *grad_sigma_t, double *grad_albedo, int nthreads);
/* use_drt = True (volpathsimple.py:33-36, :520-594, :623-778): oracle only, not part of the product ABI */
int orc_vps_backward_drt(const rcv_scene_desc *S, const orc_medium *M, uint32_t seed_grad, uint32_t spp_grad,
                         uint32_t spp_begin, uint32_t spp_end, const float *grad_img, const float *film_w,
                         double *grad_sigma_t, double *grad_albedo, int nthreads, int use_drt_subsampling, int use_drt_mis);

float orc_phase_sample(const rcv_scene_desc *S, const float wi[3], float s2x, float s2y, float wo_out[3]);
float orc_phase_eval(const rcv_scene_desc *S, const float wi[3], const float wo[3]);

float orc_adam_lr_t(float lr, float beta1, float beta2, int t);
void orc_adam(float *p, const float *g, float *m, float *v, uint64_t n, int t, float lr, float beta1,
              float beta2, float eps);
void orc_clamp(float *p, uint64_t n, float lo, float hi);
void orc_upsample2x(const float *in, float *out, int nz, int ny, int nx, int c);

#ifdef __cplusplus
}
#endif
#endif
