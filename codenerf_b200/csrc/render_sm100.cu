// TEMPORARY stub (replaced by the tcgen05 implementation).
#include "render_sm100.cuh"
size_t cnb_sm100_packed_bytes(const cnb_net_config*) { return 256; }
int cnb_sm100_pack_weights(const cnb_net_config*, const float* const*, void*, cudaStream_t) { return CNB_OK; }
size_t cnb_sm100_mlp_workspace_bytes(const cnb_net_config*, int64_t, int, int) { return 256; }
size_t cnb_sm100_render_workspace_bytes(const cnb_net_config*, const cnb_ray_batch*, int) { return 256; }
int cnb_sm100_mlp_forward(const cnb_net_config*, const float* const*, const void*, const float*, const float*, const float*, const float*, int, int64_t, int64_t, float*, float*, void*, size_t, cudaStream_t) { return CNB_E_UNSUPPORTED; }
int cnb_sm100_mlp_backward(const cnb_net_config*, const float* const*, const void*, const float*, const float*, const float*, const float*, int, int64_t, int64_t, const float*, const float*, float*, float*, float*, void*, size_t, cudaStream_t) { return CNB_E_UNSUPPORTED; }
int cnb_sm100_render(const cnb_net_config*, const float* const*, const void*, const cnb_ray_batch*, int, const float*, const float*, const float*, float, float*, float*, float*, float*, float*, float*, float*, void*, size_t, cudaStream_t) { return CNB_E_UNSUPPORTED; }
