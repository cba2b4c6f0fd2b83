// render_sm100.cuh -- host entry points of the bf16 tcgen05/TMEM path; see render_sm100.cu.
#pragma once
#include "common.cuh"

size_t cnb_sm100_packed_bytes(const cnb_net_config* cfg);
int cnb_sm100_pack_weights(const cnb_net_config* cfg, const float* const* params, void* packed, cudaStream_t st);
size_t cnb_sm100_mlp_workspace_bytes(const cnb_net_config* cfg, int64_t S, int n_codes, int backward);
size_t cnb_sm100_render_workspace_bytes(const cnb_net_config* cfg, const cnb_ray_batch* rays, int backward);
int cnb_sm100_mlp_forward(const cnb_net_config* cfg, const float* const* params, const void* packed,
                          const float* xyz, const float* viewdir, const float* shape_codes, const float* tex_codes,
                          int n_codes, int64_t samples_per_code, int64_t S, float* sigmas, float* rgbs, void* ws,
                          size_t ws_bytes, cudaStream_t st);
int cnb_sm100_mlp_backward(const cnb_net_config* cfg, const float* const* params, const void* packed,
                           const float* xyz, const float* viewdir, const float* shape_codes, const float* tex_codes,
                           int n_codes, int64_t samples_per_code, int64_t S, const float* d_sigmas,
                           const float* d_rgbs, float* d_params, float* d_shape, float* d_tex, void* ws,
                           size_t ws_bytes, cudaStream_t st);
// mode 0: forward; 1: backward from (d_rgb, d_depth); 2: train step (seed from target)
int cnb_sm100_render(const cnb_net_config* cfg, const float* const* params, const void* packed,
                     const cnb_ray_batch* rays, int mode, const float* d_rgb, const float* d_depth,
                     const float* target, float loss_scale, float* rgb, float* depth, float* acc, float* sq_err,
                     float* d_params, float* d_shape, float* d_tex, void* ws, size_t ws_bytes, cudaStream_t st);
// debug: non-zero if any mbarrier wait inside the tensor-core kernels ever timed out (synchronises the device)
int cnb_sm100_pipeline_timeouts(void);
// 1 once the tensor-core backward kernels are in place (interim: bf16 backward runs the fp32 kernels)
int cnb_sm100_has_backward(void);
