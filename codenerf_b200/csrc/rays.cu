// rays.cu -- get_rays, sample_from_rays and the standalone volume_rendering
// kernels (the unfused pieces of the reference API; the fused path lives in
// render_*.cu).  HBM-bound elementwise work: coalesced, one element per thread.
#include "common.cuh"

// ---------------------------------------------------------------------------
// get_rays -- reference src/utils.py:10-19 (bit-exact).
__global__ void k_get_rays(CnbRaySource s, int n_pix, float* __restrict__ rays_o, float* __restrict__ viewdirs) {
    __shared__ float cam[16];
    if (threadIdx.x < 16) cam[threadIdx.x] = s.c2w[threadIdx.x];
    __syncthreads();
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= n_pix) return;
    float o[3], v[3];
    cnb_pixel_ray(s, cam, pix, o, v);
#pragma unroll
    for (int k = 0; k < 3; ++k) { rays_o[(size_t)pix * 3 + k] = o[k]; viewdirs[(size_t)pix * 3 + k] = v[k]; }
}

extern "C" int cnb_get_rays(int H, int W, double focal, int focal_is_f64, const float* c2w,
                            float* rays_o, float* viewdirs, cnb_stream_t stream) {
    if (H <= 0 || W <= 0 || !(focal > 0.0) || !c2w || !rays_o || !viewdirs) return CNB_E_INVALID;
    if ((int64_t)H * W > (int64_t)1 << 30) return CNB_E_UNSUPPORTED;
    CnbRaySource s = {};
    s.c2w = c2w; s.focal = focal; s.inv_focal = 1.0 / focal; s.focal32 = (float)focal; s.focal_is_f64 = focal_is_f64;
    s.H = H; s.W = W; s.half_w = (float)(W * 0.5); s.half_h = (float)(H * 0.5);
    const int n = H * W;
    k_get_rays<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(s, n, rays_o, viewdirs);
    CNB_LAUNCH_CHECK();
    return CNB_OK;
}

// ---------------------------------------------------------------------------
// sample_from_rays -- reference src/utils.py:30-31 (bit-exact: mul then add).
// One thread per output float: consecutive threads write consecutive addresses.
__global__ void k_sample_from_rays(const float* __restrict__ ro, const float* __restrict__ vd,
                                   const float* __restrict__ z, int64_t total /* R*N*3 */, int N,
                                   float* __restrict__ xyz, float* __restrict__ vd_rep) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t sample = i / 3;
    const int k = (int)(i - sample * 3);
    const int64_t ray = sample / N;
    const int zi = (int)(sample - ray * N);
    const float v = __ldg(vd + ray * 3 + k);
    xyz[i] = cnb_sample_coord(__ldg(ro + ray * 3 + k), v, __ldg(z + zi));
    vd_rep[i] = v;
}

extern "C" int cnb_sample_from_rays(const float* rays_o, const float* viewdirs, const float* z_vals, int64_t R, int N,
                                    float* xyz, float* viewdir_rep, cnb_stream_t stream) {
    if (R < 0 || N <= 0 || !rays_o || !viewdirs || !z_vals || !xyz || !viewdir_rep) return CNB_E_INVALID;
    if (R == 0) return CNB_OK;
    const int64_t total = R * N * 3;
    const int64_t blocks = (total + 255) / 256;
    if (blocks > 0x7fffffff) return CNB_E_UNSUPPORTED;
    k_sample_from_rays<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(rays_o, viewdirs, z_vals, total, N, xyz,
                                                                           viewdir_rep);
    CNB_LAUNCH_CHECK();
    return CNB_OK;
}

// ---------------------------------------------------------------------------
// volume_rendering -- reference src/utils.py:34-47.  One warp per ray: each lane
// owns a contiguous run of samples, the exclusive transmittance product is a
// warp shuffle scan, the weighted sums a warp reduction.
#define VR_MAX_PER_LANE 16   // N <= 512

__device__ __forceinline__ float warp_excl_prod(float local, float& total) {
    // inclusive scan of per-lane products, then shift
    float incl = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        float up = __shfl_up_sync(0xffffffffu, incl, d);
        if ((threadIdx.x & 31) >= d) incl *= up;
    }
    total = __shfl_sync(0xffffffffu, incl, 31);
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    return (threadIdx.x & 31) == 0 ? 1.f : excl;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

__global__ void k_volume_rendering_fwd(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                       const float* __restrict__ z, int64_t B, int N, int white_bg,
                                       float* __restrict__ rgb, float* __restrict__ depth, float* __restrict__ acc) {
    const int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (ray >= B) return;
    const int lane = threadIdx.x & 31;
    const int per = (N + 31) / 32;
    const int i0 = lane * per;
    float alpha[VR_MAX_PER_LANE];
    float tl = 1.f;
#pragma unroll
    for (int j = 0; j < VR_MAX_PER_LANE; ++j) {
        const int i = i0 + j;
        if (j < per && i < N) {
            const float delta = (i + 1 < N) ? (z[i + 1] - z[i]) : 1e10f;       // utils.py:35-36
            alpha[j] = 1.f - expf(-sigmas[ray * N + i] * delta);                // utils.py:37
            tl *= (1.f - alpha[j] + 1e-10f);                                    // utils.py:38
        } else alpha[j] = 0.f;
    }
    float tot;
    float T = warp_excl_prod(tl, tot);                                          // utils.py:39-40
    float cr = 0.f, cg = 0.f, cb = 0.f, d = 0.f, ws = 0.f;
#pragma unroll
    for (int j = 0; j < VR_MAX_PER_LANE; ++j) {
        const int i = i0 + j;
        if (j < per && i < N) {
            const float w = alpha[j] * T;                                       // utils.py:41
            const float* c = rgbs + (ray * N + i) * 3;
            cr += w * c[0]; cg += w * c[1]; cb += w * c[2];                     // utils.py:42
            d += w * z[i];                                                      // utils.py:43
            ws += w;
            T *= (1.f - alpha[j] + 1e-10f);
        }
    }
    cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb); d = warp_sum(d); ws = warp_sum(ws);
    if (lane == 0) {
        if (white_bg) { cr = cr + 1.f - ws; cg = cg + 1.f - ws; cb = cb + 1.f - ws; }   // utils.py:44-46
        rgb[ray * 3 + 0] = cr; rgb[ray * 3 + 1] = cg; rgb[ray * 3 + 2] = cb;
        depth[ray] = d;
        if (acc) acc[ray] = ws;
    }
}

// Reverse mode of the above.  Adjoint of T_{i+1} = T_i t_i, division free:
//   aT_i = g_i alpha_i + aT_{i+1} t_i ;  d alpha_i = g_i T_i - aT_{i+1} T_i.
// aT is an affine recurrence run backwards; each lane composes its run into
// (A, Bc) with aT_in = A * aT_out + Bc, a warp suffix-scan composes the lanes.
__global__ void k_volume_rendering_bwd(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                       const float* __restrict__ z_base, int z_per_segment, int rays_per_segment,
                                       int64_t ray_first, int64_t B, int N, int white_bg,
                                       const float* __restrict__ d_rgb, const float* __restrict__ d_depth,
                                       float* __restrict__ d_sigmas, float* __restrict__ d_rgbs) {
    const int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (ray >= B) return;
    const int lane = threadIdx.x & 31;
    const int per = (N + 31) / 32;
    const int i0 = lane * per;
    const float* z = z_base + (z_per_segment ? ((ray_first + ray) / rays_per_segment) * N : 0);
    const float gr = d_rgb[ray * 3], gg = d_rgb[ray * 3 + 1], gb = d_rgb[ray * 3 + 2];
    const float gd = d_depth ? d_depth[ray] : 0.f;
    const float bg = white_bg ? 1.f : 0.f;
    float alpha[VR_MAX_PER_LANE], g[VR_MAX_PER_LANE], dl[VR_MAX_PER_LANE], ex[VR_MAX_PER_LANE];
    float tl = 1.f;
#pragma unroll
    for (int j = 0; j < VR_MAX_PER_LANE; ++j) {
        const int i = i0 + j;
        if (j < per && i < N) {
            dl[j] = (i + 1 < N) ? (z[i + 1] - z[i]) : 1e10f;
            ex[j] = expf(-sigmas[ray * N + i] * dl[j]);
            alpha[j] = 1.f - ex[j];
            tl *= (1.f - alpha[j] + 1e-10f);
            const float* c = rgbs + (ray * N + i) * 3;
            g[j] = gr * (c[0] - bg) + gg * (c[1] - bg) + gb * (c[2] - bg) + gd * z[i];
        } else { alpha[j] = 0.f; g[j] = 0.f; dl[j] = 0.f; ex[j] = 0.f; }
    }
    float tot;
    const float T0 = warp_excl_prod(tl, tot);
    // per-lane affine map of the backward recurrence: aT_in(lane) = A * aT_out + Bc
    float A = 1.f, Bc = 0.f;
#pragma unroll
    for (int j = VR_MAX_PER_LANE - 1; j >= 0; --j) {
        const int i = i0 + j;
        if (j < per && i < N) {
            const float t = 1.f - alpha[j] + 1e-10f;
            // aT_i = g alpha + t * aT_{i+1};  compose with (A,Bc) representing aT_{i+1} = A x + Bc
            Bc = g[j] * alpha[j] + t * Bc;
            A = t * A;
        }
    }
    // suffix composition across lanes: value entering lane L from above = f_{L+1}(f_{L+2}(...f_31(0)))
    float sA = A, sB = Bc;   // inclusive suffix map for lanes >= this lane
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const float oA = __shfl_down_sync(0xffffffffu, sA, d);
        const float oB = __shfl_down_sync(0xffffffffu, sB, d);
        if (lane + d < 32) { sB = sA * oB + sB; sA = sA * oA; }
    }
    float aT_next = __shfl_down_sync(0xffffffffu, sB, 1);   // aT entering this lane's last sample (x = 0 beyond lane 31)
    if (lane == 31) aT_next = 0.f;
    // this lane's T at each of its samples
    float Tj[VR_MAX_PER_LANE];
    float T = T0;
#pragma unroll
    for (int j = 0; j < VR_MAX_PER_LANE; ++j) {
        Tj[j] = T;
        const int i = i0 + j;
        if (j < per && i < N) T *= (1.f - alpha[j] + 1e-10f);
    }
#pragma unroll
    for (int j = VR_MAX_PER_LANE - 1; j >= 0; --j) {
        const int i = i0 + j;
        if (j < per && i < N) {
            const float w = alpha[j] * Tj[j];
            float* dc = d_rgbs + (ray * N + i) * 3;
            dc[0] = w * gr; dc[1] = w * gg; dc[2] = w * gb;
            const float t = 1.f - alpha[j] + 1e-10f;
            const float a_alpha = g[j] * Tj[j] - aT_next * Tj[j];
            d_sigmas[ray * N + i] = a_alpha * dl[j] * ex[j];
            aT_next = g[j] * alpha[j] + aT_next * t;
        }
    }
}

extern "C" int cnb_volume_rendering_forward(const float* sigmas, const float* rgbs, const float* z_vals, int64_t B,
                                            int N, int white_bg, float* rgb, float* depth, float* acc,
                                            cnb_stream_t stream) {
    if (B < 0 || N <= 0 || !sigmas || !rgbs || !z_vals || !rgb || !depth) return CNB_E_INVALID;
    if (N > 32 * VR_MAX_PER_LANE) return CNB_E_UNSUPPORTED;
    if (B == 0) return CNB_OK;
    const int64_t blocks = (B * 32 + 255) / 256;
    if (blocks > 0x7fffffff) return CNB_E_UNSUPPORTED;
    k_volume_rendering_fwd<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(sigmas, rgbs, z_vals, B, N, white_bg,
                                                                               rgb, depth, acc);
    CNB_LAUNCH_CHECK();
    return CNB_OK;
}

extern "C" int cnb_volume_rendering_backward(const float* sigmas, const float* rgbs, const float* z_vals, int64_t B,
                                             int N, int white_bg, const float* d_rgb, const float* d_depth,
                                             float* d_sigmas, float* d_rgbs, cnb_stream_t stream) {
    if (B < 0 || N <= 0 || !sigmas || !rgbs || !z_vals || !d_rgb || !d_sigmas || !d_rgbs) return CNB_E_INVALID;
    if (N > 32 * VR_MAX_PER_LANE) return CNB_E_UNSUPPORTED;
    if (B == 0) return CNB_OK;
    const int64_t blocks = (B * 32 + 255) / 256;
    if (blocks > 0x7fffffff) return CNB_E_UNSUPPORTED;
    k_volume_rendering_bwd<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(sigmas, rgbs, z_vals, 0, 1, 0, B, N,
                                                                               white_bg, d_rgb, d_depth, d_sigmas,
                                                                               d_rgbs);
    CNB_LAUNCH_CHECK();
    return CNB_OK;
}

// Same kernel for a slice [ray0, ray0 + n_rays) of a segmented batch (one z row per segment).
int cnb_vr_backward_segments(const float* sigmas, const float* rgbs, const float* z_vals, int z_per_segment,
                             int rays_per_segment, int64_t ray0, int64_t n_rays, int N, int white_bg,
                             const float* d_rgb, const float* d_depth, float* d_sigmas, float* d_rgbs, cudaStream_t st) {
    if (N > 32 * VR_MAX_PER_LANE) return CNB_E_UNSUPPORTED;
    if (n_rays <= 0) return CNB_OK;
    const int64_t blocks = (n_rays * 32 + 255) / 256;
    k_volume_rendering_bwd<<<(unsigned)blocks, 256, 0, st>>>(sigmas, rgbs, z_vals, z_per_segment, rays_per_segment, ray0,
                                                            n_rays, N, white_bg, d_rgb, d_depth, d_sigmas, d_rgbs);
    CNB_LAUNCH_CHECK();
    return CNB_OK;
}
