// rcv_scene.h — rcv_scene_desc (public ABI) -> SceneK (kernel parameter block).
#pragma once
#include <cmath>
#include <vector>
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
    S.emitter = d.emitter_type;
    S.env_w = d.env_width; S.env_h = d.env_height;
    for (int i = 0; i < 9; ++i) S.env_R[i] = d.env_to_world[i];
    S.env_data = nullptr; S.env_cdf_m = nullptr; S.env_cdf_c = nullptr; S.env_pdf = nullptr;   // set by the owner of the tables
    // rectangles: normal, shading frame and reciprocals in binary32 with this operation order (the oracle does the same)
    S.n_quads = (int) d.n_quads; S.area_quad = d.area_quad;
    auto dot = [](const float *a, const float *b) { return std::fmaf(a[0], b[0], std::fmaf(a[1], b[1], a[2] * b[2])); };
    auto cross = [](const float *a, const float *b, float *c) {
        c[0] = std::fmaf(a[1], b[2], -(a[2] * b[1])); c[1] = std::fmaf(a[2], b[0], -(a[0] * b[2])); c[2] = std::fmaf(a[0], b[1], -(a[1] * b[0]));
    };
    for (int k = 0; k < 8; ++k) {
        QuadK &q = S.quads[k];
        q = QuadK{};
        if (k >= S.n_quads) continue;
        const rcv_quad &r = d.quads[k];
        for (int i = 0; i < 3; ++i) { q.p0[i] = r.p0[i]; q.eu[i] = r.eu[i]; q.ev[i] = r.ev[i]; q.refl[i] = r.reflectance[i]; q.rad[i] = r.radiance[i]; }
        float c[3];
        cross(q.eu, q.ev, c);
        float len = std::sqrt(dot(c, c));
        for (int i = 0; i < 3; ++i) q.n[i] = c[i] / len;
        float lu2 = dot(q.eu, q.eu), lv2 = dot(q.ev, q.ev);
        q.inv_eu2 = 1.f / lu2; q.inv_ev2 = 1.f / lv2; q.inv_area = 1.f / len;
        float lu = std::sqrt(lu2);
        for (int i = 0; i < 3; ++i) q.s[i] = q.eu[i] / lu;
        cross(q.n, q.s, q.t);
    }
}

// `envmap` tables (host): scaled texels with the wrap column, marginal / conditional cdfs over the (h - 1) x w cells and
// the density of (u, v) per cell, proportional to the cell mean of luminance * sin(theta).  Float64, sequential sums,
// rounded once (the test oracle builds the same tables independently).
struct EnvTables {
    std::vector<float4> data;
    std::vector<float> cdf_m, cdf_c, pdf;
};
inline void env_tables_build(const rcv_scene_desc &d, EnvTables &T) {
    const int w = d.env_width, h = d.env_height;
    T.data.assign((size_t) h * (w + 1), make_float4(0.f, 0.f, 0.f, 0.f));
    T.cdf_m.assign(h, 0.f); T.cdf_c.assign((size_t) (h - 1) * (w + 1), 0.f); T.pdf.assign((size_t) (h - 1) * w, 0.f);
    std::vector<double> g((size_t) h * (w + 1)), rowsum(h - 1);
    for (int r = 0; r < h; ++r) {
        double st = std::sin(3.14159265358979323846 * (double) r / (double) (h - 1));
        for (int c = 0; c <= w; ++c) {
            const float *src = d.envmap_rgb + 3 * ((size_t) r * w + (c % w));
            float4 t = make_float4(src[0] * d.env_scale, src[1] * d.env_scale, src[2] * d.env_scale, 0.f);
            T.data[(size_t) r * (w + 1) + c] = t;
            g[(size_t) r * (w + 1) + c] = (0.2126 * (double) t.x + 0.7152 * (double) t.y + 0.0722 * (double) t.z) * st;
        }
    }
    double total = 0.0;
    for (int r = 0; r < h - 1; ++r) {
        double s = 0.0;
        for (int c = 0; c < w; ++c) {
            size_t i = (size_t) r * (w + 1) + c;
            s += 0.25 * (g[i] + g[i + 1] + g[i + w + 1] + g[i + w + 2]);
        }
        rowsum[r] = s; total += s;
    }
    double acc = 0.0;
    T.cdf_m[0] = 0.f;
    for (int r = 0; r < h - 1; ++r) {
        acc += rowsum[r];
        T.cdf_m[r + 1] = (total > 0.0) ? (float) (acc / total) : (float) (r + 1) / (float) (h - 1);
        float *row = T.cdf_c.data() + (size_t) r * (w + 1);
        double a2 = 0.0;
        row[0] = 0.f;
        for (int c = 0; c < w; ++c) {
            size_t i = (size_t) r * (w + 1) + c;
            double f = 0.25 * (g[i] + g[i + 1] + g[i + w + 1] + g[i + w + 2]);
            a2 += f;
            row[c + 1] = (rowsum[r] > 0.0) ? (float) (a2 / rowsum[r]) : (float) (c + 1) / (float) w;
            T.pdf[(size_t) r * w + c] = (total > 0.0) ? (float) (f / total * (double) w * (double) (h - 1)) : 1.f;
        }
        row[w] = 1.f;
    }
    T.cdf_m[h - 1] = 1.f;
}

}  // namespace rcv
