// mlp_fp32.cuh -- host entry points of the fp32 (CUDA-core) mode; see mlp_fp32.cu.
#pragma once
#include "common.cuh"

int64_t cnb_fp32_chunk_rows(int N);
size_t cnb_fp32_workspace_bytes(const cnb_net_config* c, int64_t S, int N, int n_codes, int backward, int fused);
int cnb_fp32_mlp_forward(const cnb_net_config* cfg, const float* const* params, const float* xyz,
                         const float* viewdir, const float* shape_codes, const float* tex_codes, int n_codes,
                         int64_t samples_per_code, int64_t S, float* sigmas, float* rgbs, void* ws, size_t ws_bytes,
                         cudaStream_t st);
int cnb_fp32_mlp_backward(const cnb_net_config* cfg, const float* const* params, const float* xyz,
                          const float* viewdir, const float* shape_codes, const float* tex_codes, int n_codes,
                          int64_t samples_per_code, int64_t S, const float* d_sigmas, const float* d_rgbs,
                          float* d_params, float* d_shape, float* d_tex, void* ws, size_t ws_bytes, cudaStream_t st);
int cnb_fp32_render(const cnb_net_config* cfg, const float* const* params, const cnb_ray_batch* rays, int mode,
                    const float* d_rgb, const float* d_depth, const float* target, float loss_scale,
                    float* rgb, float* depth, float* acc, float* sq_err,
                    float* d_params, float* d_shape, float* d_tex, void* ws, size_t ws_bytes, cudaStream_t st);

// z[code][n] = ReLU(Wl[n,:] . codes[code,:] + bl[n])  -- reference src/model.py:41, :49
int cnb_launch_latent_fwd(const float* Wl, const float* bl, const float* codes, int n_codes, int LD, int W,
                          float* z, int64_t ldz, cudaStream_t st);
// dpre = dz * (z > 0); dW += dpre (x) code; db += dpre; dcodes += Wl^T dpre   (dW / db may be null)
int cnb_launch_latent_bwd(const float* Wl, const float* codes, const float* z, const float* dz, int64_t ldz,
                          int n_codes, int LD, int W, float* dW, float* db, float* dcodes, cudaStream_t st);
