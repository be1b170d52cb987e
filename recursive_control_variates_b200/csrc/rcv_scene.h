// rcv_scene.h — rcv_scene_desc (public ABI) -> SceneK (kernel parameter block).
#pragma once
#include "../../include/rcv_b200.h"
#include "rcv_device.cuh"

namespace rcv {

inline void scene_from_desc(const rcv_scene_desc &d, SceneK &S) {
    S.W = d.film_width; S.H = d.film_height;
    S.rcpW = 1.f / (float) S.W; S.rcpH = 1.f / (float) S.H;
    for (int i = 0; i < 3; ++i) {
        S.cam_o[i] = d.cam_origin[i];
        S.bmin[i] = d.bbox_min[i]; S.bmax[i] = d.bbox_max[i];
        S.inv_ext[i] = 1.f / (d.bbox_max[i] - d.bbox_min[i]);
        S.env[i] = d.env_radiance[i];
        S.env_over_pdf[i] = d.env_radiance[i] / RCV_INV_FOURPI;
    }
    for (int i = 0; i < 9; ++i) S.cam_axes[i] = d.cam_axes[i];
    S.tan_x = d.cam_tan_x; S.tan_y = d.cam_tan_y; S.near_clip = d.cam_near_clip;
    S.scale = d.sigma_t_scale; S.phase = d.phase_type; S.g = d.phase_g;
    S.sample_emitters = d.sample_emitters; S.env_dist = d.env_dist;
    S.max_depth = d.max_depth; S.rr_depth = d.rr_depth; S.hide_emitters = d.hide_emitters; S.use_nee = d.use_nee;
    S.rfilter = d.rfilter_type; S.rf_stddev = d.rfilter_stddev;
    S.n_sensors = (int) d.n_sensors;
    for (int k = 0; k < RCV_MAX_SENSORS; ++k) {
        for (int i = 0; i < 3; ++i) S.batch_o[3 * k + i] = d.batch_origin[k][i];
        for (int i = 0; i < 9; ++i) S.batch_axes[9 * k + i] = d.batch_axes[k][i];
    }
}

}  // namespace rcv
