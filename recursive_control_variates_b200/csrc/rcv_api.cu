// rcv_api.cu — the C-ABI of librcv_b200.so (include/rcv_b200.h): context,
// packed medium state, and the entry points the reference's plugin classes
// call in place of their Dr.Jit code.  No torch types, no hidden host syncs
// except in the *_host functions.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rcv_b200.h"
#include "rcv_internal.h"
#include "rcv_scene.h"

using namespace rcv;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
int fail_cuda(cudaError_t e, const char *what) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return RCV_ERR_CUDA;
}
#define CK(expr)                                              \
    do {                                                      \
        cudaError_t _e = (expr);                              \
        if (_e != cudaSuccess) return fail_cuda(_e, #expr);   \
    } while (0)

}  // namespace

struct rcv_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    int tracer = RCV_TRACER_POOL;
    int drt_flags = 0;                 // volpathsimple's use_drt / use_drt_subsampling / use_drt_mis
    bool l2_persist = false;      // pin the d sigma_t grid in L2 while the adjoint runs (RCV_L2_PERSIST=1)
    size_t l2_persist_max = 0, l2_window_max = 0;
    bool has_scene = false;
    rcv_scene_desc desc{};
    SceneK S{};
    // packed medium state
    int nx = 0, ny = 0, nz = 0;
    size_t nvox = 0;
    float2 *sig2 = nullptr;
    float4 *blk2 = nullptr, *blk1 = nullptr;
    int maj_factor = 0;           // majorant supergrid: voxels per supercell side (0 = one majorant for the volume)
    float2 *mgrid = nullptr;      // (old, new) local majorants, (nz/f, ny/f, nx/f)
    size_t mgrid_cells = 0;
    bool mgrid_valid = false;     // mgrid matches the current grids and factor
    float4 *alb2 = nullptr;
    float *grad = nullptr;       // [pad4(nvox) floats d sigma_t][nvox float4 d albedo]
    float4 *gbrick = nullptr;    // RCV_GRAD_BRICKS: per-cell d sigma_t bricks of the adjoint launch in flight (zero otherwise)
    size_t gbrick_cells = 0;
    size_t grad_floats = 0;
    float *majorants = nullptr;
    float *block_max = nullptr;
    int n_rot_blocks = 0;
    // scratch
    Counters *counters = nullptr;
    unsigned long long *work = nullptr;
    uint64_t launches = 0;
    float *film_w = nullptr;
    size_t film_w_pix = 0;
    // `envmap` emitter tables (rcv_scene_set)
    float4 *env_data = nullptr;
    float *env_cdf_m = nullptr, *env_cdf_c = nullptr, *env_pdf = nullptr;
    PathQueue queue{nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t queue_cap = 0;
    PathQueue jobs{nullptr, nullptr, nullptr, nullptr, nullptr};     // DRT job records (use_drt on the pool tracer)
    size_t jobs_cap = 0;
    float4 *arena = nullptr;      // vertex records of the adjoint: slots x (max_depth + 1) x 64 B
    size_t arena_records = 0;
    uint32_t vcap = 0;
    double *loss_partial = nullptr;
    int n_partial = 0;
    // state of rcv_cv_iteration_host
    struct {
        size_t nvox = 0, npix = 0;
        float *sigma = nullptr, *albedo = nullptr, *ref = nullptr;
        float *acc = nullptr;      // 2 x (H,W,4)
        float *HF = nullptr;       // 2 x (H,W,3)
        float *cv = nullptr;       // 8 x (H,W,3): img v_n w_s mean_x mean_y var_x var_y cov
        float *grad_img = nullptr, *loss = nullptr;
        float *g_sigma = nullptr, *g_albedo = nullptr;
        int it = 0, stats_n = 0;
    } e2e;

    LaunchCfg cfg() {
        LaunchCfg c;
        c.stream = stream; c.sm_count = sm_count; c.tracer = tracer; c.drt_flags = drt_flags;
        c.counters = counters; c.work_counter = work; c.launches = &launches; c.queue = queue; c.jobs = jobs;
        c.arena = arena; c.vcap = vcap;
        return c;
    }
    GradK gradk() const {
        GradK g;
        g.gsig = grad; g.galb = (float4 *) (grad + ((nvox + 3) / 4) * 4);
        g.gbrick = gbrick;
        return g;
    }
    GridKS gridk() const {
        GridKS g;
        g.sig2 = sig2; g.blk2 = blk2; g.blk1 = blk1; g.alb2 = alb2; g.nx = nx; g.ny = ny; g.nz = nz; g.majorants = majorants;
        g.mgrid = nullptr; g.mx = g.my = g.mz = 0; g.vox[0] = g.vox[1] = g.vox[2] = 0.f;
        if (maj_factor > 0 && mgrid_valid) {
            g.mgrid = mgrid; g.mx = nx / maj_factor; g.my = ny / maj_factor; g.mz = nz / maj_factor;
            g.vox[0] = 1.f / (float) g.mx; g.vox[1] = 1.f / (float) g.my; g.vox[2] = 1.f / (float) g.mz;
        }
        return g;
    }
};

namespace {

int use_device(rcv_ctx *ctx) {
    if (!ctx) return fail(RCV_ERR_ARG, "null context");
    CK(cudaSetDevice(ctx->device));
    return RCV_OK;
}

void free_medium(rcv_ctx *c) {
    cudaFree(c->sig2); cudaFree(c->blk2); cudaFree(c->blk1); cudaFree(c->alb2); cudaFree(c->grad); cudaFree(c->block_max);
    c->sig2 = nullptr; c->blk2 = nullptr; c->blk1 = nullptr; c->alb2 = nullptr; c->grad = nullptr; c->block_max = nullptr;
    c->nvox = 0; c->nx = c->ny = c->nz = 0;
}

// nearest x2 upsample of the previous `new` state into the `new` slot of
// freshly allocated packed grids, so that the following rotation moves it to
// the `old` slot (the reference keeps theta_{i-1} at the coarse
// resolution for the single iteration after utils.py:177-192 upsample()).
__global__ void k_upsample_state(const float2 *sig2_lo, const float4 *alb2_lo, int nz, int ny, int nx, float2 *sig2_hi,
                                 float4 *alb2_hi) {
    size_t n_hi = (size_t) 8 * nz * ny * nx;
    for (size_t v = (size_t) blockIdx.x * blockDim.x + threadIdx.x; v < n_hi; v += (size_t) gridDim.x * blockDim.x) {
        size_t r = v;
        int x = (int) (r % (2 * nx)); r /= (2 * nx);
        int y = (int) (r % (2 * ny)); r /= (2 * ny);
        int z = (int) r;
        size_t lo = ((size_t) (z >> 1) * ny + (y >> 1)) * nx + (x >> 1);
        sig2_hi[v] = make_float2(0.f, sig2_lo[lo].y);
        alb2_hi[2 * v] = make_float4(0.f, 0.f, 0.f, 0.f);
        alb2_hi[2 * v + 1] = alb2_lo[2 * lo + 1];
    }
}

float one_minus_pow(float beta, int n) { return (float) (1.0 - std::pow((double) beta, (double) n)); }
float adam_lr_t(float lr, float beta1, float beta2, int t) {
    double s = std::sqrt(1.0 - std::pow((double) beta2, (double) t)) / (1.0 - std::pow((double) beta1, (double) t));
    return (float) ((double) lr * s);
}

// seed of volpathsimple's secondary sampler (plugins/volpathsimple.py:112-120):
// lane 0's fourth float reinterpreted as uint32, scrambled by TEA — a pure
// function of seed_grad, so no device->host read is needed.
uint32_t alt_seed_for(uint32_t seed_grad) {
    Pcg r;
    sampler_seed(r, seed_grad, 0u);
    pcg_u32(r); pcg_u32(r); pcg_u32(r);
    uint32_t bits = (pcg_u32(r) >> 9) | 0x3f800000u;
    float f;
    std::memcpy(&f, &bits, 4);
    f = f - 1.f;
    uint32_t fb;
    std::memcpy(&fb, &f, 4);
    uint32_t v0, v1;
    tea4(fb, 1u, v0, v1);
    return v0;
}

// record queue of the wavefront schedule: 48 B per lane of the largest launch so far
int ensure_queue(rcv_ctx *c, uint32_t spp_begin, uint32_t spp_end) {
    size_t n = (size_t) c->S.W * c->S.H * (spp_end - spp_begin);
    if (!c->queue.count) CK(cudaMalloc(&c->queue.count, sizeof(unsigned long long)));
    if (n > c->queue_cap) {
        cudaFree(c->queue.r0); cudaFree(c->queue.r1); cudaFree(c->queue.r2); cudaFree(c->queue.r3);
        c->queue.r0 = c->queue.r1 = c->queue.r2 = c->queue.r3 = nullptr; c->queue_cap = 0;
        CK(cudaMalloc(&c->queue.r0, n * sizeof(float4)));
        CK(cudaMalloc(&c->queue.r1, n * sizeof(float4)));
        CK(cudaMalloc(&c->queue.r2, n * sizeof(float4)));
        CK(cudaMalloc(&c->queue.r3, n * sizeof(float4)));
        c->queue_cap = n;
    }
    return RCV_OK;
}

// RCV_GRAD_BRICKS: the per-cell gradient bricks of the current grid, zero-filled once (k_unbrick leaves them zero)
int ensure_gbrick(rcv_ctx *c) {
#if RCV_GRAD_BRICKS
    size_t ncell = (size_t) (c->nz + 1) * (c->ny + 1) * (c->nx + 1);
    if (ncell != c->gbrick_cells) {
        cudaFree(c->gbrick); c->gbrick = nullptr; c->gbrick_cells = 0;
        CK(cudaMalloc(&c->gbrick, 2 * ncell * sizeof(float4)));
        CK(cudaMemset(c->gbrick, 0, 2 * ncell * sizeof(float4)));     // once per grid shape; synchronous: the stream may change
        c->gbrick_cells = ncell;
    }
#endif
    return RCV_OK;
}

// job records of the DRT stages: one per lane of the launch
int ensure_jobs(rcv_ctx *c, uint32_t spp_begin, uint32_t spp_end) {
    size_t n = (size_t) c->S.W * c->S.H * (spp_end - spp_begin);
    if (!c->jobs.count) CK(cudaMalloc(&c->jobs.count, sizeof(unsigned long long)));
    if (n > c->jobs_cap) {
        cudaFree(c->jobs.r0); cudaFree(c->jobs.r1); cudaFree(c->jobs.r2); cudaFree(c->jobs.r3);
        c->jobs.r0 = c->jobs.r1 = c->jobs.r2 = c->jobs.r3 = nullptr; c->jobs_cap = 0;
        CK(cudaMalloc(&c->jobs.r0, n * sizeof(float4)));
        CK(cudaMalloc(&c->jobs.r1, n * sizeof(float4)));
        CK(cudaMalloc(&c->jobs.r2, n * sizeof(float4)));
        CK(cudaMalloc(&c->jobs.r3, n * sizeof(float4)));
        c->jobs_cap = n;
    }
    return RCV_OK;
}

// vertex-record arena of the vertex-list adjoint: one slot of (max_depth + 1) vertex records per lane in flight, one
// spare record (the DRT reservoir) and the shadow-ray collision records
int ensure_arena(rcv_ctx *c) {
    c->vcap = (uint32_t) c->S.max_depth + 2u;
    size_t need = vps_backward_slots(c->cfg()) * ((size_t) c->vcap + 16);      // + the shadow-ray collision records (kNeeCap / 4)
    if (need > c->arena_records) {
        cudaFree(c->arena); c->arena = nullptr; c->arena_records = 0;
        CK(cudaMalloc(&c->arena, need * 4 * sizeof(float4)));
        c->arena_records = need;
    }
    return RCV_OK;
}

// (re)build the majorant supergrid from the current two-state grid.  No-op while the factor is 0 or no medium is set;
// a factor that does not divide the current grid leaves the supergrid invalid (renders are refused until the next
// rotation brings a grid it divides: utils.py:177-192 changes the factor before it pushes the upsampled grid).
int build_majorant_grid(rcv_ctx *c) {
    c->mgrid_valid = false;
    if (c->maj_factor <= 0 || !c->sig2) return RCV_OK;
    const int f = c->maj_factor;
    if (c->nz % f || c->ny % f || c->nx % f) return RCV_OK;
    size_t cells = (size_t) (c->nz / f) * (c->ny / f) * (c->nx / f);
    if (cells > c->mgrid_cells) {
        cudaFree(c->mgrid); c->mgrid = nullptr; c->mgrid_cells = 0;
        CK(cudaMalloc(&c->mgrid, cells * sizeof(float2)));
        c->mgrid_cells = cells;
    }
    CK(launch_majorant_grid(c->cfg(), c->sig2, c->nz, c->ny, c->nx, f, c->S.scale, c->mgrid));
    c->mgrid_valid = true;
    return RCV_OK;
}

int ensure_film_w(rcv_ctx *c) {
    size_t npix = (size_t) c->S.W * c->S.H;
    if (c->film_w_pix != npix) {
        cudaFree(c->film_w); c->film_w = nullptr;
        CK(cudaMalloc(&c->film_w, npix * sizeof(float)));
        c->film_w_pix = npix;
    }
    return RCV_OK;
}

}  // namespace

extern "C" {

const char *rcv_last_error(void) { return g_err.c_str(); }
int rcv_abi_version(void) { return RCV_ABI_VERSION; }
int rcv_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return -1; }
    return n;
}

int rcv_ctx_create(int device, void *stream, rcv_ctx **out) {
    if (!out) return fail(RCV_ERR_ARG, "out is null");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(RCV_ERR_CUDA, "no CUDA device available (librcv_b200 has no CPU fallback)");
    }
    if (device < 0 || device >= n) return fail(RCV_ERR_ARG, "invalid device index");
    CK(cudaSetDevice(device));
    rcv_ctx *c = new rcv_ctx();
    c->device = device;
    c->stream = (cudaStream_t) stream;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    c->l2_persist_max = (size_t) prop.persistingL2CacheMaxSize;
    c->l2_window_max = (size_t) prop.accessPolicyMaxWindowSize;
    if (const char *e = getenv("RCV_L2_PERSIST")) c->l2_persist = atoi(e) != 0;
    if (getenv("RCV_VERBOSE")) fprintf(stderr, "rcv: L2 %d MB, persisting max %zu MB, window max %zu MB\n", prop.l2CacheSize >> 20, c->l2_persist_max >> 20, c->l2_window_max >> 20);
    if (c->l2_persist && c->l2_persist_max > 0) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, c->l2_persist_max);
    CK(cudaMalloc(&c->counters, sizeof(Counters)));
    CK(cudaMemset(c->counters, 0, sizeof(Counters)));
    CK(cudaMalloc(&c->work, sizeof(unsigned long long)));
    CK(cudaMalloc(&c->majorants, 2 * sizeof(float)));
    CK(cudaMemset(c->majorants, 0, 2 * sizeof(float)));
    c->n_partial = c->sm_count * 4;
    CK(cudaMalloc(&c->loss_partial, c->n_partial * sizeof(double)));
    *out = c;
    return RCV_OK;
}

int rcv_ctx_destroy(rcv_ctx *c) {
    if (!c) return RCV_OK;
    cudaSetDevice(c->device);
    free_medium(c);
    cudaFree(c->counters); cudaFree(c->work); cudaFree(c->majorants); cudaFree(c->film_w); cudaFree(c->loss_partial);
    cudaFree(c->queue.r0); cudaFree(c->queue.r1); cudaFree(c->queue.r2); cudaFree(c->queue.r3); cudaFree(c->queue.count);
    cudaFree(c->jobs.r0); cudaFree(c->jobs.r1); cudaFree(c->jobs.r2); cudaFree(c->jobs.r3); cudaFree(c->jobs.count);
    cudaFree(c->gbrick); cudaFree(c->arena); cudaFree(c->mgrid);
    cudaFree(c->env_data); cudaFree(c->env_cdf_m); cudaFree(c->env_cdf_c); cudaFree(c->env_pdf);
    auto &E = c->e2e;
    cudaFree(E.sigma); cudaFree(E.albedo); cudaFree(E.ref); cudaFree(E.acc); cudaFree(E.HF); cudaFree(E.cv);
    cudaFree(E.grad_img); cudaFree(E.loss); cudaFree(E.g_sigma); cudaFree(E.g_albedo);
    delete c;
    return RCV_OK;
}

int rcv_ctx_set_stream(rcv_ctx *c, void *stream) {
    if (!c) return fail(RCV_ERR_ARG, "null context");
    c->stream = (cudaStream_t) stream;
    return RCV_OK;
}
int rcv_ctx_set_tracer(rcv_ctx *c, int kind) {
    if (!c) return fail(RCV_ERR_ARG, "null context");
    if (kind < 0 || kind > 3) return fail(RCV_ERR_ARG, "unknown tracer kind");
    c->tracer = kind;
    return RCV_OK;
}

int rcv_adjoint_set_drt(rcv_ctx *c, int use_drt, int use_drt_subsampling, int use_drt_mis) {
    if (!c) return fail(RCV_ERR_ARG, "null context");
    c->drt_flags = use_drt ? (1 | (use_drt_subsampling ? 2 : 0) | (use_drt_mis ? 4 : 0)) : 0;
    return RCV_OK;
}

int rcv_scene_set(rcv_ctx *c, const rcv_scene_desc *d) {
    if (!c || !d) return fail(RCV_ERR_ARG, "null argument");
    if (d->struct_size != sizeof(rcv_scene_desc)) return fail(RCV_ERR_ARG, "rcv_scene_desc size mismatch (ABI)");
    if (d->film_width == 0 || d->film_height == 0) return fail(RCV_ERR_ARG, "empty film");
    if (d->phase_type != RCV_PHASE_ISOTROPIC && d->phase_type != RCV_PHASE_HG) return fail(RCV_ERR_ARG, "unknown phase function");
    for (int a = 0; a < 3; ++a)
        if (!(d->bbox_max[a] > d->bbox_min[a])) return fail(RCV_ERR_ARG, "degenerate medium bounding box");
    if (d->rr_depth < 0 || d->max_depth < 1) return fail(RCV_ERR_ARG, "rr_depth must be >= 0 and max_depth >= 1");
    if (d->rfilter_type != RCV_RFILTER_BOX && d->rfilter_type != RCV_RFILTER_GAUSSIAN) return fail(RCV_ERR_ARG, "unknown reconstruction filter");
    if (d->rfilter_type == RCV_RFILTER_GAUSSIAN && !(d->rfilter_stddev > 0.f && d->rfilter_stddev <= 0.5f))
        return fail(RCV_ERR_ARG, "gaussian reconstruction filter: 0 < stddev <= 0.5");
    if (d->n_sensors > RCV_MAX_SENSORS) return fail(RCV_ERR_ARG, "too many sensors in the batch (RCV_MAX_SENSORS)");
    if (d->n_sensors > 1 && d->film_width % d->n_sensors) return fail(RCV_ERR_ARG, "the film width must be a multiple of the number of batched sensors");
    if (d->emitter_type != RCV_EMITTER_CONSTANT && d->emitter_type != RCV_EMITTER_ENVMAP && d->emitter_type != RCV_EMITTER_AREA)
        return fail(RCV_ERR_ARG, "unknown emitter type");
    if (d->emitter_type == RCV_EMITTER_ENVMAP && (!d->envmap_rgb || d->env_width < 1 || d->env_height < 2))
        return fail(RCV_ERR_ARG, "envmap: a host pointer to (env_height >= 2, env_width >= 1, 3) texels is required");
    if (d->n_quads > RCV_MAX_QUADS) return fail(RCV_ERR_ARG, "too many rectangles (RCV_MAX_QUADS)");
    for (uint32_t k = 0; k < d->n_quads; ++k) {
        const rcv_quad &q = d->quads[k];
        double uu = 0, vv = 0, uv = 0;
        for (int i = 0; i < 3; ++i) { uu += (double) q.eu[i] * q.eu[i]; vv += (double) q.ev[i] * q.ev[i]; uv += (double) q.eu[i] * q.ev[i]; }
        if (!(uu > 0) || !(vv > 0) || uv * uv > 1e-10 * uu * vv) return fail(RCV_ERR_ARG, "rectangle: the edges must be non-zero and perpendicular");
    }
    if (d->emitter_type == RCV_EMITTER_AREA && (d->area_quad < 0 || (uint32_t) d->area_quad >= d->n_quads))
        return fail(RCV_ERR_ARG, "area light: area_quad must index one of the rectangles");
    float4 *e_data = nullptr; float *e_m = nullptr, *e_c = nullptr, *e_p = nullptr;
    if (d->emitter_type == RCV_EMITTER_ENVMAP) {
        if (int r = use_device(c)) return r;
        EnvTables T;
        env_tables_build(*d, T);
        struct Fresh {
            void *p[4] = {nullptr, nullptr, nullptr, nullptr};
            ~Fresh() { for (void *q : p) cudaFree(q); }
        } fresh;
        CK(cudaMalloc(&fresh.p[0], T.data.size() * sizeof(float4)));
        CK(cudaMalloc(&fresh.p[1], T.cdf_m.size() * sizeof(float)));
        CK(cudaMalloc(&fresh.p[2], T.cdf_c.size() * sizeof(float)));
        CK(cudaMalloc(&fresh.p[3], T.pdf.size() * sizeof(float)));
        CK(cudaMemcpyAsync(fresh.p[0], T.data.data(), T.data.size() * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(fresh.p[1], T.cdf_m.data(), T.cdf_m.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(fresh.p[2], T.cdf_c.data(), T.cdf_c.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(fresh.p[3], T.pdf.data(), T.pdf.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));       // the host tables go out of scope; kernels of the old scene are done
        e_data = (float4 *) fresh.p[0]; e_m = (float *) fresh.p[1]; e_c = (float *) fresh.p[2]; e_p = (float *) fresh.p[3];
        for (void *&q : fresh.p) q = nullptr;
    }
    if (c->env_data) {
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        cudaFree(c->env_data); cudaFree(c->env_cdf_m); cudaFree(c->env_cdf_c); cudaFree(c->env_pdf);
    }
    c->env_data = e_data; c->env_cdf_m = e_m; c->env_cdf_c = e_c; c->env_pdf = e_p;
    c->desc = *d;
    c->desc.envmap_rgb = nullptr;                    // the caller's texels are not referenced after this call
    scene_from_desc(*d, c->S);
    c->S.env_data = c->env_data; c->S.env_cdf_m = c->env_cdf_m; c->S.env_cdf_c = c->env_cdf_c; c->S.env_pdf = c->env_pdf;
    c->has_scene = true;
    return RCV_OK;
}

int rcv_medium_rotate(rcv_ctx *c, const float *sigma_t_in, const float *albedo_in, int nz, int ny, int nx) {
    if (int r = use_device(c)) return r;
    if (!c->has_scene) return fail(RCV_ERR_STATE, "rcv_scene_set must be called before rcv_medium_rotate");
    if (!sigma_t_in || !albedo_in || nz <= 0 || ny <= 0 || nx <= 0) return fail(RCV_ERR_ARG, "invalid grid");
    if (c->maj_factor > 0 && (nz % c->maj_factor || ny % c->maj_factor || nx % c->maj_factor))
        return fail(RCV_ERR_ARG, "grid resolution must be divisible by the majorant resolution factor");
    size_t nvox = (size_t) nz * ny * nx;
    if ((size_t) (nz + 1) * (ny + 1) * (nx + 1) > 0xffffffffull)
        return fail(RCV_ERR_ARG, "grid too large: voxels and interpolation cells are indexed with 32 bits");
    LaunchCfg cfg = c->cfg();
    if (nz != c->nz || ny != c->ny || nx != c->nx) {
        // new buffers are owned by `fresh` until every allocation (and the state transfer) has succeeded: an
        // out-of-memory error half-way leaves the context as it was and leaks nothing
        struct Fresh {
            void *p[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
            ~Fresh() { for (void *q : p) cudaFree(q); }
            void release() { for (void *&q : p) q = nullptr; }
        } fresh;
        size_t ncell = (size_t) (nz + 1) * (ny + 1) * (nx + 1);
        size_t grad_floats = ((nvox + 3) / 4) * 4 + 4 * nvox;
        CK(cudaMalloc(&fresh.p[0], nvox * sizeof(float2)));
        CK(cudaMalloc(&fresh.p[1], 4 * ncell * sizeof(float4)));
        CK(cudaMalloc(&fresh.p[2], 2 * ncell * sizeof(float4)));
        CK(cudaMalloc(&fresh.p[3], 2 * nvox * sizeof(float4)));
        CK(cudaMalloc(&fresh.p[4], grad_floats * sizeof(float)));
        uint64_t want = (nvox + 255) / 256, cap = (uint64_t) c->sm_count * 8;
        int nb = (int) (want < cap ? want : cap);
        CK(cudaMalloc(&fresh.p[5], nb * sizeof(float)));
        float2 *sig2 = (float2 *) fresh.p[0];
        float4 *blk2 = (float4 *) fresh.p[1], *blk1 = (float4 *) fresh.p[2], *alb2 = (float4 *) fresh.p[3];
        float *grad = (float *) fresh.p[4], *bm = (float *) fresh.p[5];
        if (c->sig2 && nz == 2 * c->nz && ny == 2 * c->ny && nx == 2 * c->nx) {
            k_upsample_state<<<c->sm_count * 8, 256, 0, c->stream>>>(c->sig2, c->alb2, c->nz, c->ny, c->nx, sig2, alb2);
            c->launches++;
            CK(cudaGetLastError());
            CK(cudaStreamSynchronize(c->stream));
        } else {
            CK(cudaMemsetAsync(sig2, 0, nvox * sizeof(float2), c->stream));
            CK(cudaMemsetAsync(alb2, 0, 2 * nvox * sizeof(float4), c->stream));
            CK(cudaMemsetAsync(c->majorants, 0, 2 * sizeof(float), c->stream));
        }
        CK(cudaMemsetAsync(grad, 0, grad_floats * sizeof(float), c->stream));
        free_medium(c);
        fresh.release();
        c->sig2 = sig2; c->blk2 = blk2; c->blk1 = blk1; c->alb2 = alb2; c->grad = grad; c->block_max = bm; c->n_rot_blocks = nb;
        c->nz = nz; c->ny = ny; c->nx = nx; c->nvox = nvox; c->grad_floats = grad_floats;
    }
    CK(launch_rotate(cfg, sigma_t_in, albedo_in, nz, ny, nx, c->S.scale, c->sig2, c->blk2, c->blk1, c->alb2, c->block_max,
                     c->n_rot_blocks, c->majorants));
    return build_majorant_grid(c);
}

int rcv_medium_set_majorant_factor(rcv_ctx *c, int factor) {
    if (int r = use_device(c)) return r;
    if (factor < 0) return fail(RCV_ERR_ARG, "majorant resolution factor must be >= 0");
    if (factor == 1) factor = 0;     /* utils.py:36-37: a factor <= 1 disables the supergrid */
    c->maj_factor = factor;
    return build_majorant_grid(c);
}

int rcv_medium_majorant_grid_host(rcv_ctx *c, float *out_old, float *out_new, int dims[3]) {
    if (int r = use_device(c)) return r;
    if (!dims) return fail(RCV_ERR_ARG, "null argument");
    dims[0] = dims[1] = dims[2] = 0;
    if (c->maj_factor <= 0 || !c->mgrid_valid) return RCV_OK;
    int mz = c->nz / c->maj_factor, my = c->ny / c->maj_factor, mx = c->nx / c->maj_factor;
    dims[0] = mz; dims[1] = my; dims[2] = mx;
    if (!out_old && !out_new) return RCV_OK;
    size_t n = (size_t) mz * my * mx;
    std::vector<float2> h(n);
    CK(cudaMemcpyAsync(h.data(), c->mgrid, n * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < n; ++i) { if (out_old) out_old[i] = h[i].x; if (out_new) out_new[i] = h[i].y; }
    return RCV_OK;
}

int rcv_medium_majorants_host(rcv_ctx *c, float out[2]) {
    if (int r = use_device(c)) return r;
    CK(cudaMemcpyAsync(out, c->majorants, 2 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return RCV_OK;
}

int rcv_medium_read(rcv_ctx *c, int state, float *sigma_t_out, float *albedo_out) {
    if (int r = use_device(c)) return r;
    if (!c->sig2) return fail(RCV_ERR_STATE, "no medium state");
    if (state != 0 && state != 1) return fail(RCV_ERR_ARG, "state must be 0 (old) or 1 (new)");
    CK(launch_medium_read(c->cfg(), c->sig2, c->alb2, c->nvox, state, sigma_t_out, albedo_out));
    return RCV_OK;
}

static int check_render(rcv_ctx *c, uint32_t spp, uint32_t b, uint32_t e) {
    if (int r = use_device(c)) return r;
    if (!c->has_scene) return fail(RCV_ERR_STATE, "no scene set");
    if (!c->sig2) return fail(RCV_ERR_STATE, "no medium state (call rcv_medium_rotate)");
    if (spp == 0 || b > e || e > spp) return fail(RCV_ERR_ARG, "invalid sample range");
    if (c->maj_factor > 0 && !c->mgrid_valid)
        return fail(RCV_ERR_STATE, "the majorant resolution factor does not divide the current grid resolution");
    if ((uint64_t) c->S.W * c->S.H * spp > 0xffffffffull) return fail(RCV_ERR_ARG, "wavefront exceeds 2^32 lanes");
    if (c->tracer == RCV_TRACER_WAVEFRONT || c->tracer == RCV_TRACER_POOL)
        if (int r = ensure_queue(c, b, e)) return r;
    return RCV_OK;
}

int rcv_render_twostates(rcv_ctx *c, uint32_t seed, uint32_t spp, uint32_t spp_begin, uint32_t spp_end, float *accum_old,
                         float *accum_new) {
    if (int r = check_render(c, spp, spp_begin, spp_end)) return r;
    if (!accum_old || !accum_new) return fail(RCV_ERR_ARG, "null accumulator");
    if (c->S.n_quads > 0)      // surfaces: the machine of rcv_surface_wf.cuh
        CK(launch_surface_primal(c->cfg(), c->S, c->gridk(), 0, seed, spp, spp_begin, spp_end, (float4 *) accum_old, (float4 *) accum_new));
    else
        CK(launch_twostates(c->cfg(), c->S, c->gridk(), 0, seed, spp, spp_begin, spp_end, (float4 *) accum_old, (float4 *) accum_new));
    return RCV_OK;
}

int rcv_render_primal(rcv_ctx *c, int state, uint32_t seed, uint32_t spp, uint32_t spp_begin, uint32_t spp_end, float *accum) {
    if (int r = check_render(c, spp, spp_begin, spp_end)) return r;
    if (!accum) return fail(RCV_ERR_ARG, "null accumulator");
    if (state != 0 && state != 1) return fail(RCV_ERR_ARG, "state must be 0 (old) or 1 (new)");
    if (c->S.n_quads > 0)
        CK(launch_surface_primal(c->cfg(), c->S, c->gridk(), state == 0 ? 1 : 2, seed, spp, spp_begin, spp_end, (float4 *) accum,
                                 (float4 *) accum));
    else
        CK(launch_twostates(c->cfg(), c->S, c->gridk(), state == 0 ? 1 : 2, seed, spp, spp_begin, spp_end, (float4 *) accum,
                            (float4 *) accum));
    return RCV_OK;
}

int rcv_develop(rcv_ctx *c, const float *accum, float *img) {
    if (int r = use_device(c)) return r;
    if (!c->has_scene) return fail(RCV_ERR_STATE, "no scene set");
    if (!accum || !img) return fail(RCV_ERR_ARG, "null pointer");
    CK(launch_develop(c->cfg(), (const float4 *) accum, img, (size_t) c->S.W * c->S.H));
    return RCV_OK;
}

int rcv_cv_update(rcv_ctx *c, const float *H, const float *F, float *img, float *v_n, float *w_s, float *mean_x, float *mean_y,
                  float *var_x, float *var_y, float *cov, uint64_t n, int32_t it, int32_t warmup, int32_t stats_n, float beta1,
                  float beta2, int32_t pre_update) {
    if (int r = use_device(c)) return r;
    if (!H || !F || !img || !v_n || !w_s || !mean_x || !mean_y || !var_x || !var_y || !cov) return fail(RCV_ERR_ARG, "null pointer");
    if (n == 0) return RCV_OK;
    CvArgs a;
    a.H = H; a.F = F; a.img = img; a.v_n = v_n; a.w_s = w_s; a.mean_x = mean_x; a.mean_y = mean_y;
    a.var_x = var_x; a.var_y = var_y; a.cov = cov; a.n = n;
    a.do_weights = it > warmup; a.do_update = it > 0; a.pre_update = pre_update != 0;
    a.stats_n_is_zero = stats_n == 0;
    a.c1a = one_minus_pow(beta1, stats_n); a.c1b = one_minus_pow(beta1, stats_n + 1);
    int n_get = stats_n + ((a.do_update && a.pre_update) ? 1 : 0);
    a.c2 = one_minus_pow(beta2, n_get);
    a.beta1 = beta1; a.beta2 = beta2;
    CK(launch_cv_update(c->cfg(), a));
    return RCV_OK;
}

int rcv_loss_grad(rcv_ctx *c, int kind, const float *img, const float *ref, uint64_t n, float *loss_out, float *grad_img) {
    if (int r = use_device(c)) return r;
    if (kind != RCV_LOSS_L1 && kind != RCV_LOSS_L2) return fail(RCV_ERR_ARG, "unknown loss");
    if (!img || !ref || !loss_out || n == 0) return fail(RCV_ERR_ARG, "invalid argument");
    CK(launch_loss_grad(c->cfg(), kind, img, ref, n, c->loss_partial, c->n_partial, loss_out, grad_img));
    return RCV_OK;
}

int rcv_film_weights(rcv_ctx *c, uint32_t seed, uint32_t spp, uint32_t spp_begin, uint32_t spp_end, float *film_w) {
    if (int r = use_device(c)) return r;
    if (!c->has_scene) return fail(RCV_ERR_STATE, "no scene set");
    if (!film_w || spp == 0 || spp_begin > spp_end || spp_end > spp) return fail(RCV_ERR_ARG, "invalid argument");
    CK(launch_film_weights(c->cfg(), c->S, seed, spp, spp_begin, spp_end, film_w, false));
    return RCV_OK;
}

int rcv_render_backward(rcv_ctx *c, int adjoint_kind, uint32_t seed_grad, uint32_t spp_grad, uint32_t spp_begin, uint32_t spp_end,
                        const float *grad_img) {
    return rcv_render_backward_weighted(c, adjoint_kind, seed_grad, spp_grad, spp_begin, spp_end, grad_img, nullptr);
}

int rcv_render_backward_weighted(rcv_ctx *c, int adjoint_kind, uint32_t seed_grad, uint32_t spp_grad, uint32_t spp_begin,
                                 uint32_t spp_end, const float *grad_img, const float *film_w) {
    if (int r = check_render(c, spp_grad, spp_begin, spp_end)) return r;
    if (adjoint_kind != RCV_ADJOINT_VOLPATHSIMPLE && adjoint_kind != RCV_ADJOINT_PRBVOLPATH)
        return fail(RCV_ERR_ARG, "unknown adjoint integrator");
    if (!grad_img) return fail(RCV_ERR_ARG, "null grad_img");
    if (adjoint_kind == RCV_ADJOINT_VOLPATHSIMPLE && (c->S.n_quads > 0 || c->S.emitter == RCV_EMITTER_AREA))
        return fail(RCV_ERR_STATE, "volpathsimple handles neither surfaces nor area lights (volpathsimple.py:24-30): use RCV_ADJOINT_PRBVOLPATH");
    if (int r = ensure_film_w(c)) return r;
    if (adjoint_kind == RCV_ADJOINT_PRBVOLPATH) {
        // upstream prbvolpath (the adjoint twostateprbvolpath inherits): both walks on the surface machine
        LaunchCfg cfg = c->cfg();
        if (!film_w) {
            CK(launch_film_weights(cfg, c->S, seed_grad, spp_grad, 0, spp_grad, c->film_w, true));
            film_w = c->film_w;
        }
        if (int r = ensure_gbrick(c)) return r;
        GradK D = c->gradk();
        CK(launch_prb_backward(cfg, c->S, c->gridk(), D, seed_grad, spp_grad, spp_begin, spp_end, grad_img, film_w));
#if RCV_GRAD_BRICKS
        CK(launch_unbrick(cfg, D, c->nx, c->ny, c->nz));
#endif
        return RCV_OK;
    }
    // deep paths (max_depth = -1 ...) would need an unbounded vertex arena: use the replaying schedule for them
    const int saved_tracer = c->tracer;
    if ((c->tracer == RCV_TRACER_POOL || c->tracer == RCV_TRACER_WAVEFRONT) && c->S.max_depth >= 4096) c->tracer = RCV_TRACER_BATCH;
    // use_drt: sub-sampling on the pool tracer runs on the event machines (arena + job records), every other
    // combination on the one-thread-per-lane kernel, which needs neither
    const bool drt_pool = (c->drt_flags & 3) == 3 && c->tracer == RCV_TRACER_POOL;
    if ((c->tracer == RCV_TRACER_POOL || c->tracer == RCV_TRACER_WAVEFRONT) && (!(c->drt_flags & 1) || drt_pool))
        if (int r = ensure_arena(c)) { c->tracer = saved_tracer; return r; }
    if (drt_pool)
        if (int r = ensure_jobs(c, spp_begin, spp_end)) { c->tracer = saved_tracer; return r; }
    LaunchCfg cfg = c->cfg();
    c->tracer = saved_tracer;
    if (!film_w) {       // no weights supplied: those of ALL spp_grad samples are computed here
        CK(launch_film_weights(cfg, c->S, seed_grad, spp_grad, 0, spp_grad, c->film_w, true));
        film_w = c->film_w;
    }
    if (int r = ensure_gbrick(c)) return r;
    GradK D = c->gradk();
    const bool pin = c->l2_persist && c->l2_persist_max > 0 && c->l2_window_max > 0;
    // the access-policy window is dropped (and the persisting lines released) on every way out, errors included
    struct PinGuard {
        cudaStream_t st; bool on;
        ~PinGuard() {
            if (!on) return;
            cudaStreamAttrValue a{};
            a.accessPolicyWindow.num_bytes = 0;
            cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &a);
            cudaCtxResetPersistingL2Cache();       // give the set-aside lines back to the kernels that follow
        }
    } guard{c->stream, false};
    if (pin) {
        // L2 access-policy window over the d sigma_t grid: the adjoint's scattered reductions hit L2
        cudaStreamAttrValue a{};
        size_t bytes = c->nvox * sizeof(float);
        if (bytes > c->l2_window_max) bytes = c->l2_window_max;
        a.accessPolicyWindow.base_ptr = (void *) D.gsig;
        a.accessPolicyWindow.num_bytes = bytes;
        a.accessPolicyWindow.hitRatio = bytes <= c->l2_persist_max ? 1.f : (float) c->l2_persist_max / (float) bytes;
        a.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        a.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        CK(cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &a));
        guard.on = true;
    }
    CK(launch_vps_backward(cfg, c->S, c->gridk(), D, seed_grad, spp_grad, spp_begin, spp_end, alt_seed_for(seed_grad), grad_img,
                           film_w));
#if RCV_GRAD_BRICKS
    CK(launch_unbrick(cfg, D, c->nx, c->ny, c->nz));
#endif
    return RCV_OK;
}

int rcv_render_adjoint_primal(rcv_ctx *c, int adjoint_kind, uint32_t seed, uint32_t spp, uint32_t spp_begin, uint32_t spp_end,
                              float *accum) {
    if (int r = check_render(c, spp, spp_begin, spp_end)) return r;
    if (adjoint_kind == RCV_ADJOINT_PRBVOLPATH) return rcv_render_primal(c, 1, seed, spp, spp_begin, spp_end, accum);
    if (adjoint_kind != RCV_ADJOINT_VOLPATHSIMPLE) return fail(RCV_ERR_ARG, "unknown adjoint integrator");
    if (c->S.n_quads > 0 || c->S.emitter == RCV_EMITTER_AREA)
        return fail(RCV_ERR_STATE, "volpathsimple handles neither surfaces nor area lights (volpathsimple.py:24-30)");
    if (!accum) return fail(RCV_ERR_ARG, "null accumulator");
    CK(launch_vps_primal(c->cfg(), c->S, c->gridk(), seed, spp, spp_begin, spp_end, (float4 *) accum));
    return RCV_OK;
}

int rcv_grad_zero(rcv_ctx *c) {
    if (int r = use_device(c)) return r;
    if (!c->grad) return fail(RCV_ERR_STATE, "no medium state");
    CK(cudaMemsetAsync(c->grad, 0, c->grad_floats * sizeof(float), c->stream));
    return RCV_OK;
}
int rcv_grad_buffer(rcv_ctx *c, float **ptr, uint64_t *n_floats) {
    if (!c || !ptr || !n_floats) return fail(RCV_ERR_ARG, "null argument");
    if (!c->grad) return fail(RCV_ERR_STATE, "no medium state");
    *ptr = c->grad; *n_floats = (uint64_t) c->grad_floats;
    return RCV_OK;
}
int rcv_grad_read(rcv_ctx *c, float *grad_sigma_t, float *grad_albedo) {
    if (int r = use_device(c)) return r;
    if (!c->grad) return fail(RCV_ERR_STATE, "no medium state");
    CK(launch_grad_unpack(c->cfg(), c->gradk(), c->nvox, grad_sigma_t, grad_albedo));
    return RCV_OK;
}

int rcv_adam_step(rcv_ctx *c, float *sigma_t, float *albedo, float *m_sigma_t, float *v_sigma_t, float *m_albedo, float *v_albedo,
                  int32_t t, float lr_sigma_t, float lr_albedo, float beta1, float beta2, float eps, float slo, float shi,
                  float alo, float ahi) {
    if (int r = use_device(c)) return r;
    if (!c->grad) return fail(RCV_ERR_STATE, "no medium state");
    if (t < 1) return fail(RCV_ERR_ARG, "t must be >= 1");
    if (!sigma_t || !albedo || !m_sigma_t || !v_sigma_t || !m_albedo || !v_albedo) return fail(RCV_ERR_ARG, "null pointer");
    CK(launch_adam_packed(c->cfg(), c->gradk(), c->nvox, sigma_t, albedo, m_sigma_t, v_sigma_t, m_albedo, v_albedo,
                          adam_lr_t(lr_sigma_t, beta1, beta2, t), adam_lr_t(lr_albedo, beta1, beta2, t), beta1, beta2, eps, slo,
                          shi, alo, ahi));
    return RCV_OK;
}
int rcv_adam_flat(rcv_ctx *c, float *p, const float *g, float *m, float *v, uint64_t n, int32_t t, float lr, float beta1,
                  float beta2, float eps) {
    if (int r = use_device(c)) return r;
    if (t < 1) return fail(RCV_ERR_ARG, "t must be >= 1");
    if (!p || !g || !m || !v) return fail(RCV_ERR_ARG, "null pointer");
    if (n == 0) return RCV_OK;
    CK(launch_adam_flat(c->cfg(), p, g, m, v, n, adam_lr_t(lr, beta1, beta2, t), beta1, beta2, eps));
    return RCV_OK;
}
int rcv_clamp(rcv_ctx *c, float *p, uint64_t n, float lo, float hi) {
    if (int r = use_device(c)) return r;
    if (!p) return fail(RCV_ERR_ARG, "null pointer");
    if (n == 0) return RCV_OK;
    CK(launch_clamp(c->cfg(), p, n, lo, hi));
    return RCV_OK;
}
int rcv_upsample2x(rcv_ctx *c, const float *in, float *out, int nz, int ny, int nx, int ch) {
    if (int r = use_device(c)) return r;
    if (!in || !out || nz <= 0 || ny <= 0 || nx <= 0 || ch <= 0) return fail(RCV_ERR_ARG, "invalid argument");
    CK(launch_upsample2x(c->cfg(), in, out, nz, ny, nx, ch));
    return RCV_OK;
}

int rcv_counters_host(rcv_ctx *c, uint64_t out[8], int reset) {
    if (int r = use_device(c)) return r;
    Counters h;
    CK(cudaMemcpyAsync(&h, c->counters, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < 8; ++i) out[i] = h.c[i];
    out[7] = c->launches;
    if (reset) {
        CK(cudaMemsetAsync(c->counters, 0, sizeof(Counters), c->stream));
        c->launches = 0;
    }
    return RCV_OK;
}

int rcv_cv_iteration_host(rcv_ctx *c, const rcv_iter_host_args *a) {
    if (int r = use_device(c)) return r;
    if (!a || a->struct_size != sizeof(rcv_iter_host_args)) return fail(RCV_ERR_ARG, "rcv_iter_host_args size mismatch (ABI)");
    if (!c->has_scene) return fail(RCV_ERR_STATE, "no scene set");
    if (!a->sigma_t_in || !a->albedo_in || !a->ref || !a->img_out || !a->loss_out) return fail(RCV_ERR_ARG, "null host buffer");
    auto &E = c->e2e;
    size_t nvox = (size_t) a->nz * a->ny * a->nx, npix = (size_t) c->S.W * c->S.H;
    if (E.nvox != nvox) {
        cudaFree(E.sigma); cudaFree(E.albedo); cudaFree(E.g_sigma); cudaFree(E.g_albedo);
        CK(cudaMalloc(&E.sigma, nvox * sizeof(float)));
        CK(cudaMalloc(&E.albedo, 3 * nvox * sizeof(float)));
        CK(cudaMalloc(&E.g_sigma, nvox * sizeof(float)));
        CK(cudaMalloc(&E.g_albedo, 3 * nvox * sizeof(float)));
        E.nvox = nvox;
    }
    if (E.npix != npix || a->reset) {
        if (E.npix != npix) {
            cudaFree(E.ref); cudaFree(E.acc); cudaFree(E.HF); cudaFree(E.cv); cudaFree(E.grad_img); cudaFree(E.loss);
            CK(cudaMalloc(&E.ref, 3 * npix * sizeof(float)));
            CK(cudaMalloc(&E.acc, 8 * npix * sizeof(float)));
            CK(cudaMalloc(&E.HF, 6 * npix * sizeof(float)));
            CK(cudaMalloc(&E.cv, 24 * npix * sizeof(float)));
            CK(cudaMalloc(&E.grad_img, 3 * npix * sizeof(float)));
            CK(cudaMalloc(&E.loss, sizeof(float)));
            E.npix = npix;
        }
        CK(cudaMemsetAsync(E.cv, 0, 24 * npix * sizeof(float), c->stream));   // integrator.reset() (cv_integrator.py:29-43)
        E.it = 0; E.stats_n = 0;
    }
    cudaStream_t st = c->stream;
    const uint64_t n3 = 3 * npix;
    // host -> device: this iteration's parameters and the reference image
    CK(cudaMemcpyAsync(E.sigma, a->sigma_t_in, nvox * sizeof(float), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(E.albedo, a->albedo_in, 3 * nvox * sizeof(float), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(E.ref, a->ref, n3 * sizeof(float), cudaMemcpyHostToDevice, st));
    if (int r = rcv_medium_rotate(c, E.sigma, E.albedo, a->nz, a->ny, a->nx)) return r;      // optimize.py:107
    CK(cudaMemsetAsync(E.acc, 0, 8 * npix * sizeof(float), st));
    float *acc_old = E.acc, *acc_new = E.acc + 4 * npix, *H = E.HF, *F = E.HF + n3;
    if (int r = rcv_render_twostates(c, a->seed, a->spp, 0, a->spp, acc_old, acc_new)) return r;   // optimize.py:108
    if (int r = rcv_develop(c, acc_old, H)) return r;
    if (int r = rcv_develop(c, acc_new, F)) return r;
    float *cv = E.cv;
    if (int r = rcv_cv_update(c, H, F, cv, cv + n3, cv + 2 * n3, cv + 3 * n3, cv + 4 * n3, cv + 5 * n3, cv + 6 * n3, cv + 7 * n3,
                              n3, E.it, a->warmup, E.stats_n, a->beta1, a->beta2, 0)) return r;
    if (E.it > 0) E.stats_n += 1;
    E.it += 1;
    if (int r = rcv_loss_grad(c, a->loss_kind, cv, E.ref, n3, E.loss, E.grad_img)) return r;   // optimize.py:110-122
    if (int r = rcv_grad_zero(c)) return r;
    if (int r = rcv_render_backward(c, a->adjoint_kind, a->seed_grad, a->spp_grad, 0, a->spp_grad, E.grad_img)) return r;
    // device -> host: image, loss, gradients
    CK(cudaMemcpyAsync(a->img_out, cv, n3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(a->loss_out, E.loss, sizeof(float), cudaMemcpyDeviceToHost, st));
    if (a->grad_sigma_t_out || a->grad_albedo_out) {
        if (int r = rcv_grad_read(c, a->grad_sigma_t_out ? E.g_sigma : nullptr, a->grad_albedo_out ? E.g_albedo : nullptr)) return r;
        if (a->grad_sigma_t_out)
            CK(cudaMemcpyAsync(a->grad_sigma_t_out, E.g_sigma, nvox * sizeof(float), cudaMemcpyDeviceToHost, st));
        if (a->grad_albedo_out)
            CK(cudaMemcpyAsync(a->grad_albedo_out, E.g_albedo, 3 * nvox * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    return RCV_OK;
}

}  // extern "C"
