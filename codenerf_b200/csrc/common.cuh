// common.cuh -- shared host/device helpers of the CodeNeRF B200 render path.
//
// Reference semantics cited as file:line into yuliangguo/code-nerf.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>

#include "../../include/codenerf_b200.h"

#define CNB_MAX_BLOCKS 8          // shape_blocks / texture_blocks supported by the kernels
#define CNB_MAX_TENSORS (12 + 8 * CNB_MAX_BLOCKS)

extern std::atomic<long long> g_cnb_launches;
#define CNB_COUNT_LAUNCH() (g_cnb_launches.fetch_add(1, std::memory_order_relaxed))

#define CNB_CUDA_TRY(expr)                                  \
    do {                                                    \
        cudaError_t _e = (expr);                            \
        if (_e != cudaSuccess) return (int)_e;              \
    } while (0)
#define CNB_TRY(expr)                                       \
    do {                                                    \
        int _rc = (expr);                                   \
        if (_rc != 0) return _rc;                           \
    } while (0)
#define CNB_LAUNCH_CHECK()                                  \
    do {                                                    \
        CNB_COUNT_LAUNCH();                                 \
        cudaError_t _e = cudaGetLastError();                \
        if (_e != cudaSuccess) return (int)_e;              \
    } while (0)

// Tuning / experiment switches (cnb_set_option in the C ABI).  Lookup order: a value set through cnb_set_option,
// else the environment variable CNB_<NAME> (upper case), else `dflt`.  Read at launch time.
int64_t cnb_option(const char* name, int64_t dflt);

// Optional per-kernel CUDA-event timing (bench.py's roofline leg); no-ops unless enabled.
enum { CNB_K_FWD = 0, CNB_K_BWD = 1, CNB_K_WGRAD = 2, CNB_K_COUNT = 3 };
void cnb_prof_begin(int kernel_id, cudaStream_t st);
void cnb_prof_end(int kernel_id, cudaStream_t st);

// Offsets (in floats) of every parameter tensor inside the flat state_dict-ordered
// vector -- reference src/model.py:20-34.
struct CnbLayout {
    int64_t enc_xyz_w, enc_xyz_b;
    int64_t sl_w[CNB_MAX_BLOCKS], sl_b[CNB_MAX_BLOCKS], s_w[CNB_MAX_BLOCKS], s_b[CNB_MAX_BLOCKS];
    int64_t enc_shape_w, enc_shape_b, sigma_w, sigma_b, enc_vd_w, enc_vd_b;
    int64_t tl_w[CNB_MAX_BLOCKS], tl_b[CNB_MAX_BLOCKS], t_w[CNB_MAX_BLOCKS], t_b[CNB_MAX_BLOCKS];
    int64_t rgb0_w, rgb0_b, rgb2_w, rgb2_b;
    int64_t total;
    int d_xyz, d_dir, n_tensors;
    // tensor index (state_dict order) of each entry, for the pointer table
    int i_enc_xyz, i_sl[CNB_MAX_BLOCKS], i_s[CNB_MAX_BLOCKS], i_enc_shape, i_sigma, i_enc_vd;
    int i_tl[CNB_MAX_BLOCKS], i_t[CNB_MAX_BLOCKS], i_rgb0, i_rgb2;   // index of the weight; bias = +1
};

inline int cnb_validate_config(const cnb_net_config* c) {
    if (!c) return CNB_E_INVALID;
    if (c->shape_blocks < 1 || c->texture_blocks < 1 || c->shape_blocks > CNB_MAX_BLOCKS ||
        c->texture_blocks > CNB_MAX_BLOCKS)
        return CNB_E_UNSUPPORTED;
    if (c->W < 16 || (c->W % 16) != 0 || c->W > 1024) return CNB_E_UNSUPPORTED;
    if (c->num_xyz_freq < 0 || c->num_xyz_freq > 16 || c->num_dir_freq < 0 || c->num_dir_freq > 16)
        return CNB_E_UNSUPPORTED;
    if (c->latent_dim < 1 || c->latent_dim > 4096) return CNB_E_UNSUPPORTED;
    return CNB_OK;
}

inline void cnb_make_layout(const cnb_net_config* c, CnbLayout* L) {
    int64_t o = 0;
    int t = 0;
    const int W = c->W, LD = c->latent_dim;
    L->d_xyz = 3 + 6 * c->num_xyz_freq;
    L->d_dir = 3 + 6 * c->num_dir_freq;
    L->i_enc_xyz = t; t += 2;
    L->enc_xyz_w = o; o += (int64_t)W * L->d_xyz; L->enc_xyz_b = o; o += W;
    for (int j = 0; j < c->shape_blocks; ++j) {
        L->i_sl[j] = t; t += 2;
        L->sl_w[j] = o; o += (int64_t)W * LD; L->sl_b[j] = o; o += W;
        L->i_s[j] = t; t += 2;
        L->s_w[j] = o; o += (int64_t)W * W; L->s_b[j] = o; o += W;
    }
    L->i_enc_shape = t; t += 2;
    L->enc_shape_w = o; o += (int64_t)W * W; L->enc_shape_b = o; o += W;
    L->i_sigma = t; t += 2;
    L->sigma_w = o; o += W; L->sigma_b = o; o += 1;
    L->i_enc_vd = t; t += 2;
    L->enc_vd_w = o; o += (int64_t)W * (W + L->d_dir); L->enc_vd_b = o; o += W;
    for (int j = 0; j < c->texture_blocks; ++j) {
        L->i_tl[j] = t; t += 2;
        L->tl_w[j] = o; o += (int64_t)W * LD; L->tl_b[j] = o; o += W;
        L->i_t[j] = t; t += 2;
        L->t_w[j] = o; o += (int64_t)W * W; L->t_b[j] = o; o += W;
    }
    L->i_rgb0 = t; t += 2;
    L->rgb0_w = o; o += (int64_t)(W / 2) * W; L->rgb0_b = o; o += W / 2;
    L->i_rgb2 = t; t += 2;
    L->rgb2_w = o; o += (int64_t)3 * (W / 2); L->rgb2_b = o; o += 3;
    L->total = o;
    L->n_tensors = t;
}

#ifndef CNB_FAST_DDIV
#define CNB_FAST_DDIV 1       // fp64 pixel / focal quotients from the host's reciprocal (0: __ddiv_rn)
#endif
// Camera / segment description passed by value to kernels that generate rays.
struct CnbRaySource {
    const float* rays_o;      // explicit rays or nullptr
    const float* viewdirs;
    const float* c2w;         // [n_segments,16]
    const int32_t* pix_begin; // [n_segments] or nullptr
    double focal;
    double inv_focal;         // RN(1 / focal), computed on the host (see cnb_ddiv_by)
    float focal32, half_w, half_h;
    int focal_is_f64, H, W;
    int rays_per_segment;
    const float* z_vals;
    int z_per_segment;
    int N;
};

inline CnbRaySource cnb_make_ray_source(const cnb_ray_batch* r) {
    CnbRaySource s;
    s.rays_o = r->rays_o; s.viewdirs = r->viewdirs; s.c2w = r->c2w; s.pix_begin = r->pix_begin;
    s.focal = r->focal; s.focal32 = (float)r->focal; s.inv_focal = 1.0 / r->focal;
    s.half_w = (float)(r->W * 0.5); s.half_h = (float)(r->H * 0.5);
    s.focal_is_f64 = r->focal_is_f64; s.H = r->H; s.W = r->W;
    s.rays_per_segment = r->rays_per_segment;
    s.z_vals = r->z_vals; s.z_per_segment = r->z_per_segment; s.N = r->n_samples;
    return s;
}

inline int cnb_validate_rays(const cnb_ray_batch* r) {
    if (!r || r->n_rays <= 0 || r->rays_per_segment <= 0 || r->n_samples <= 0) return CNB_E_INVALID;
    if (r->n_rays % r->rays_per_segment != 0) return CNB_E_INVALID;
    if (!r->z_vals || !r->shape_codes || !r->texture_codes || r->n_codes < 1) return CNB_E_INVALID;
    if (r->rays_o == nullptr) {
        if (!r->c2w || r->H <= 0 || r->W <= 0 || !(r->focal > 0.0)) return CNB_E_INVALID;
    } else if (!r->viewdirs) return CNB_E_INVALID;
    const int64_t n_seg = r->n_rays / r->rays_per_segment;
    if (r->n_codes > 1) {
        if (r->segments_per_code < 1) return CNB_E_INVALID;
        if ((n_seg + r->segments_per_code - 1) / r->segments_per_code > r->n_codes) return CNB_E_INVALID;
    }
    return CNB_OK;
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------
// Bit-exact ray generation -- reference src/utils.py:10-19.  Every rounding is
// spelled out (_rn intrinsics never contract into FMAs); see SURVEY.md section 8a R1.
// RN(a / b) from the correctly rounded reciprocal rb = RN(1 / b): q = RN(a rb) is within one ulp, the remainder
// a - q b is exact in one FMA, and RN(q + rem rb) is then the correctly rounded quotient (Markstein's final division step).
// Three fp64 instructions instead of the ~15 of a full division -- the per-sample ray generation inside the chain kernels was
// bound by the fp64 pipe.  An exact q (rem == 0) is returned as is, which also keeps the sign of a zero quotient.
__device__ __forceinline__ double cnb_ddiv_by(double a, double b, double rb) {
    const double q = __dmul_rn(a, rb);
    const double rem = __fma_rn(-q, b, a);
    return rem == 0.0 ? q : __fma_rn(rem, rb, q);
}
__device__ __forceinline__ void cnb_pixel_ray(const CnbRaySource& s, const float* __restrict__ c2w, int pix,
                                              float o[3], float v[3]) {
    const int r = pix / s.W, c = pix - r * s.W;
    const float fi = __fsub_rn((float)c, s.half_w);   // utils.py:14 (i - W*.5)
    const float fj = __fsub_rn((float)r, s.half_h);
    float dx, dy;
    if (s.focal_is_f64) {                              // fp64 focal tensor: divide in fp64, round once (utils.py:14-15)
#if CNB_FAST_DDIV
        dx = __double2float_rn(cnb_ddiv_by((double)fi, s.focal, s.inv_focal));
        dy = __double2float_rn(cnb_ddiv_by(-(double)fj, s.focal, s.inv_focal));
#else
        dx = __double2float_rn(__ddiv_rn((double)fi, s.focal));
        dy = __double2float_rn(__ddiv_rn(-(double)fj, s.focal));
#endif
    } else {
        dx = __fdiv_rn(fi, s.focal32);
        dy = __fdiv_rn(-fj, s.focal32);
    }
    float d[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {                      // utils.py:15: (p0 + p1) + p2, products rounded separately
        const float p0 = __fmul_rn(dx, c2w[k * 4 + 0]);
        const float p1 = __fmul_rn(dy, c2w[k * 4 + 1]);
        const float p2 = __fmul_rn(-1.0f, c2w[k * 4 + 2]);
        d[k] = __fadd_rn(__fadd_rn(p0, p1), p2);
        o[k] = c2w[k * 4 + 3];                         // utils.py:17
    }
    const float n2 = __fmaf_rn(d[2], d[2], __fmaf_rn(d[1], d[1], __fmul_rn(d[0], d[0])));  // torch.norm fma chain
    const float nrm = __fsqrt_rn(n2);
#pragma unroll
    for (int k = 0; k < 3; ++k) v[k] = __fdiv_rn(d[k], nrm);   // utils.py:16
}

// Ray `ray` of the batch: origin + unit direction, from memory or from the segment's camera.
__device__ __forceinline__ void cnb_fetch_ray(const CnbRaySource& s, int64_t ray, float o[3], float v[3]) {
    if (s.rays_o != nullptr) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { o[k] = __ldg(s.rays_o + ray * 3 + k); v[k] = __ldg(s.viewdirs + ray * 3 + k); }
    } else {
        const int64_t seg = ray / s.rays_per_segment;
        const int within = (int)(ray - seg * s.rays_per_segment);
        const int pix = (s.pix_begin ? __ldg(s.pix_begin + seg) : 0) + within;
        float cam[16];
#pragma unroll
        for (int k = 0; k < 12; ++k) cam[k] = __ldg(s.c2w + seg * 16 + k);
        cnb_pixel_ray(s, cam, pix, o, v);
    }
}

// xyz = ro + vd * z with separate roundings -- src/utils.py:30.
__device__ __forceinline__ float cnb_sample_coord(float o, float v, float z) { return __fadd_rn(o, __fmul_rn(v, z)); }

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

__device__ __forceinline__ float cnb_softplus(float x) { return x > 20.f ? x : log1pf(expf(x)); }  // nn.Softplus(1, 20)
#endif  // __CUDACC__
