/*
 * codenerf_b200.h -- C ABI of the B200-native CodeNeRF render path.
 *
 * One shared library (codenerf_b200/libcodenerf_b200.so), plain pointers and
 * sizes, no torch / C++ types.  Every entry point is asynchronous on the given
 * CUDA stream, never allocates or frees caller-visible memory, and returns an
 * int status: 0 = ok, < 0 = cnb_status below, > 0 = a cudaError_t.  No
 * exceptions or aborts cross this boundary.  Process-global state is limited to
 * what this header names: the launch counter (cnb_launch_count), the optional
 * kernel-timing record (cnb_profile_*), per-device helper streams/events, and
 * experiment switches read from CNB_* environment variables at launch time
 * (listed in INTEGRATION.md; none is needed for correct results).  Compute
 * entry points may be called from several host threads as long as each call
 * gets its own workspace.
 *
 * The reference (yuliangguo/code-nerf) has no FFI: its boundary for this path
 * is four Python callables and one nn.Module.  Each entry point cites the
 * reference interface it replaces; the Python mirror in codenerf_b200/ (same
 * names, signatures and return arity as the reference) binds these with ctypes
 * -- see INTEGRATION.md for the stub a maintainer of the reference would add.
 *
 * All pointers are DEVICE pointers unless a parameter says "host".
 * All tensors are dense, row-major, fp32 unless stated.
 */
#ifndef CODENERF_B200_H
#define CODENERF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* cnb_stream_t; /* == cudaStream_t */

typedef enum cnb_status {
    CNB_OK = 0,
    CNB_E_INVALID = -1,      /* bad argument (NULL, negative size, ...) */
    CNB_E_UNSUPPORTED = -2,  /* network shape / N outside what the kernels implement */
    CNB_E_WORKSPACE = -3,    /* workspace missing or too small */
    CNB_E_ALIGNMENT = -4,    /* pointer not aligned as required */
    CNB_E_DEVICE = -5        /* not an sm_100 device */
} cnb_status;

typedef enum cnb_precision {
    CNB_PRECISION_BF16 = 0,  /* bf16 tensor-core operands, fp32 accumulate: tcgen05/TMEM path (rgb/depth/acc within 1e-2) */
    CNB_PRECISION_FP32 = 1   /* fp32 operands and accumulate, CUDA cores (within 1e-4 of the reference) */
} cnb_precision;

/* CodeNeRF.__init__ keyword arguments -- reference src/model.py:11-12,
 * jsonfiles/srncar.json:2-9. */
typedef struct cnb_net_config {
    int32_t shape_blocks;
    int32_t texture_blocks;
    int32_t W;
    int32_t num_xyz_freq;
    int32_t num_dir_freq;
    int32_t latent_dim;
} cnb_net_config;

/* A batch of rays, organised as equal segments.  A segment is a run of
 * `rays_per_segment` consecutive rays that share one camera, one z_vals row and
 * one (shape, texture) code pair -- one chunk of one view in the reference's
 * loop (src/trainer.py:65-73).  n_rays = n_segments * rays_per_segment.
 *
 * Rays come either from memory (rays_o / viewdirs, as returned by get_rays) or,
 * when rays_o == NULL, are generated in-kernel from the cameras with exactly
 * the arithmetic of cnb_get_rays. */
typedef struct cnb_ray_batch {
    int64_t n_rays;
    int32_t rays_per_segment;
    int32_t n_samples;            /* N: depth samples per ray */
    const float* rays_o;          /* [n_rays,3] or NULL (camera mode) */
    const float* viewdirs;        /* [n_rays,3] or NULL (camera mode) */
    const float* c2w;             /* camera mode: [n_segments,4,4] cam-to-world */
    const int32_t* pix_begin;     /* camera mode: [n_segments] first pixel (row-major in HxW) of the segment, or NULL = 0 */
    double focal;                 /* camera mode */
    int32_t focal_is_f64;         /* camera mode: 1 = focal was an fp64 tensor (divide in fp64, reference src/utils.py:14-15) */
    int32_t H, W;                 /* camera mode */
    const float* z_vals;          /* [n_segments or 1, N] depth values, strictly increasing */
    int32_t z_per_segment;        /* 1: one z row per segment; 0: a single row shared by all */
    int32_t segments_per_code;    /* consecutive segments sharing one code row (>=1); ignored when n_codes == 1 */
    const float* shape_codes;     /* [n_codes, latent_dim] */
    const float* texture_codes;   /* [n_codes, latent_dim] */
    int32_t n_codes;              /* 1 = broadcast (the [1,256] code of src/trainer.py:70) */
    int32_t white_bg;             /* volume_rendering(white_bg=...) src/utils.py:34 */
} cnb_ray_batch;

/* ---- library ----------------------------------------------------------- */
int cnb_version(void);
const char* cnb_strerror(int status);
/* 0 when the current device is sm_100 and the kernels can run there. */
int cnb_check_device(void);

/* ---- parameter layout: CodeNeRF.state_dict() order, src/model.py:20-34 -- */
int64_t cnb_param_count(const cnb_net_config* cfg);           /* 714,756 for the SRN configs; <0 on error */
int cnb_num_param_tensors(const cnb_net_config* cfg);         /* 28 for the SRN configs */
/* offsets/rows/cols (host arrays of cnb_num_param_tensors entries; bias tensors have cols == 1) */
int cnb_param_layout(const cnb_net_config* cfg, int64_t* offsets, int32_t* rows, int32_t* cols);

/* ---- get_rays(H, W, focal, c2w) -- src/utils.py:10-19 ------------------- */
/* bit-exact with the reference's CPU arithmetic; rays_o, viewdirs: [H*W,3] */
int cnb_get_rays(int H, int W, double focal, int focal_is_f64, const float* c2w,
                 float* rays_o, float* viewdirs, cnb_stream_t stream);

/* ---- sample_from_rays(ro, vd, ...) -- src/utils.py:30-31 ---------------- */
/* z_vals [N] is computed on the host by the Python mirror with the reference's
 * own arithmetic (N scalars from python floats and the CPU mt19937 stream,
 * utils.py:24-29); this call materialises xyz [R,N,3] and the repeated
 * viewdir [R,N,3], bit-exact (separate multiply and add roundings). */
int cnb_sample_from_rays(const float* rays_o, const float* viewdirs, const float* z_vals, int64_t R, int N,
                         float* xyz, float* viewdir_rep, cnb_stream_t stream);

/* ---- volume_rendering(sigmas, rgbs, z_vals, white_bg) -- src/utils.py:34-47 */
/* sigmas [B,N], rgbs [B,N,3], z_vals [N]; rgb [B,3], depth [B], acc [B] (acc = weights.sum, utils.py:45; may be NULL) */
int cnb_volume_rendering_forward(const float* sigmas, const float* rgbs, const float* z_vals, int64_t B, int N,
                                 int white_bg, float* rgb, float* depth, float* acc, cnb_stream_t stream);
/* autograd of the above: d_depth may be NULL */
int cnb_volume_rendering_backward(const float* sigmas, const float* rgbs, const float* z_vals, int64_t B, int N,
                                  int white_bg, const float* d_rgb, const float* d_depth,
                                  float* d_sigmas, float* d_rgbs, cnb_stream_t stream);

/* ---- packed weights ------------------------------------------------------ */
/* bf16 copies of the layer matrices in the tensor-core operand layouts (both
 * orientations, K padded 63->64 and 283->288).  Derived state: rebuild after
 * every optimizer step.  `params` is a HOST array of cnb_num_param_tensors
 * DEVICE pointers in state_dict order. */
size_t cnb_packed_weights_bytes(const cnb_net_config* cfg);
int cnb_pack_weights(const cnb_net_config* cfg, const float* const* params, void* packed, cnb_stream_t stream);

/* ---- CodeNeRF.forward(xyz, viewdir, shape_latent, texture_latent) -- src/model.py:36-53 */
/* xyz, viewdir [S,3]; sample s uses code row s / samples_per_code (n_codes == 1: row 0 for all).
 * sigmas [S], rgbs [S,3]. */
size_t cnb_mlp_workspace_bytes(const cnb_net_config* cfg, int64_t S, int n_codes, int precision, int backward);
int cnb_mlp_forward(const cnb_net_config* cfg, const float* const* params, const void* packed,
                    const float* xyz, const float* viewdir,
                    const float* shape_codes, const float* texture_codes, int n_codes, int64_t samples_per_code,
                    int64_t S, int precision, float* sigmas, float* rgbs,
                    void* workspace, size_t workspace_bytes, cnb_stream_t stream);
/* autograd of the above (recomputes the forward).  d_params: flat [cnb_param_count] in
 * state_dict order, ACCUMULATED into; NULL skips the weight gradient (latent-only fit,
 * src/optimizer.py:192-198 optimises the codes only).  d_shape_codes / d_texture_codes
 * [n_codes, latent_dim] are overwritten. */
int cnb_mlp_backward(const cnb_net_config* cfg, const float* const* params, const void* packed,
                     const float* xyz, const float* viewdir,
                     const float* shape_codes, const float* texture_codes, int n_codes, int64_t samples_per_code,
                     int64_t S, int precision, const float* d_sigmas, const float* d_rgbs,
                     float* d_params, float* d_shape_codes, float* d_texture_codes,
                     void* workspace, size_t workspace_bytes, cnb_stream_t stream);

/* ---- fused render: get_rays -> sample_from_rays -> CodeNeRF -> volume_rendering
 *      (the 4-call idiom of src/trainer.py:65-74, src/optimizer.py:75-83, :113-121) */
size_t cnb_render_workspace_bytes(const cnb_net_config* cfg, const cnb_ray_batch* rays, int precision, int backward);
/* rgb [n_rays,3], depth [n_rays], acc [n_rays]; sigmas/rgbs never reach HBM in the bf16 path */
int cnb_render_forward(const cnb_net_config* cfg, const float* const* params, const void* packed,
                       const cnb_ray_batch* rays, int precision,
                       float* rgb, float* depth, float* acc,
                       void* workspace, size_t workspace_bytes, cnb_stream_t stream);
/* autograd of cnb_render_forward (recompute).  d_depth may be NULL.  d_params as in cnb_mlp_backward. */
int cnb_render_backward(const cnb_net_config* cfg, const float* const* params, const void* packed,
                        const cnb_ray_batch* rays, int precision,
                        const float* d_rgb, const float* d_depth,
                        float* d_params, float* d_shape_codes, float* d_texture_codes,
                        void* workspace, size_t workspace_bytes, cnb_stream_t stream);
/* Training step body: forward + L2 loss against `target` [n_rays,3] + backward in one pass
 * (loss_l2 = mean((rgb - target)^2) over each segment, src/trainer.py:75; the seed
 * d_rgb = 2 (rgb - target) / (3 rays_per_segment) * loss_scale never leaves the chip).
 * sq_err_sum [n_segments] receives sum((rgb-target)^2); rgb/depth/acc may be NULL. */
int cnb_render_train_step(const cnb_net_config* cfg, const float* const* params, const void* packed,
                          const cnb_ray_batch* rays, int precision,
                          const float* target, float loss_scale,
                          float* rgb, float* depth, float* acc, float* sq_err_sum,
                          float* d_params, float* d_shape_codes, float* d_texture_codes,
                          void* workspace, size_t workspace_bytes, cnb_stream_t stream);

/* Timing aid for bench.py's roofline leg.  When enabled the library brackets its main kernels
 * with CUDA events on the launching stream (up to 512 launches per kernel between reads).
 * kernel_id: 0 = fused forward, 1 = fused backward, 2 = weight-gradient GEMM.
 * cnb_profile_read waits for the recorded launches, writes their durations in ms (launch
 * order) and returns how many were written (< 0 on error), then clears the record. */
int cnb_profile_enable(int on);
int cnb_profile_read(int kernel_id, float* ms, int cap);

/* A pipeline wait inside the tensor-core kernels that exceeds its watchdog (a protocol bug)
 * TRAPS: the launch fails and every later call on the context returns a cudaError_t status.
 * This debug aid synchronises the device and returns non-zero if that happened (or if the
 * context is in an error state for any other reason). */
int cnb_debug_pipeline_timeouts(void);

/* Tuning / experiment switches (process-global; read at launch time).  Lookup order: the value set here, else the
 * environment variable CNB_<NAME> in upper case, else the built-in default.  None is needed for correct results.
 *   sub_tiles      128-row tiles per backward sub-batch (default 8192; the training stash needs ~1 MB of workspace
 *                  per tile, larger is faster; changes cnb_*_workspace_bytes)
 *   bwd_pairs, cta_pairs, weight_mcast, epi_warps, fwd_kernel_ts, k3_overlap, k3_sms, k3_items_per_sm_x10,
 *   head_mma, keep_weights, share_fills, stash_early, early_pieces, stash_lanes, stash_copy, k3_grid
 *                  kernel variants kept for measurement
 *   stash_wrap, experiment
 *                  timing experiments that produce WRONG gradients on purpose (never set them outside a benchmark)
 * Returns CNB_E_INVALID for an unknown name. */
int cnb_set_option(const char* name, int64_t value);
int cnb_clear_option(const char* name);
int64_t cnb_get_option(const char* name, int64_t default_value);

/* Number of kernel launches this library has issued in this process (for bench.py's gpu_launches). */
int64_t cnb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* CODENERF_B200_H */
