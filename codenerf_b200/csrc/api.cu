// api.cu -- the extern "C" boundary declared in include/codenerf_b200.h: argument
// validation and dispatch to the tcgen05 (bf16) or CUDA-core (fp32) implementation.
#include "common.cuh"
#include "mlp_fp32.cuh"
#include "render_sm100.cuh"
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <mutex>

std::atomic<long long> g_cnb_launches{0};

extern "C" int cnb_version(void) { return 100; }

extern "C" int64_t cnb_launch_count(void) { return (int64_t)g_cnb_launches.load(); }

extern "C" const char* cnb_strerror(int status) {
    switch (status) {
        case CNB_OK: return "ok";
        case CNB_E_INVALID: return "codenerf_b200: invalid argument";
        case CNB_E_UNSUPPORTED: return "codenerf_b200: unsupported network shape or sample count";
        case CNB_E_WORKSPACE: return "codenerf_b200: workspace missing or too small";
        case CNB_E_ALIGNMENT: return "codenerf_b200: pointer alignment";
        case CNB_E_DEVICE: return "codenerf_b200: device is not sm_100 (B200)";
        default: break;
    }
    if (status > 0) return cudaGetErrorString((cudaError_t)status);
    return "codenerf_b200: unknown status";
}

// ---- options --------------------------------------------------------------------------------
namespace {
struct Opt { const char* name; bool set; int64_t value; };
Opt g_opts[] = {
    {"sub_tiles", false, 0},      // 128-row tiles per backward sub-batch (training stash: ~1 MB per tile); default 8192
    {"bwd_pairs", false, 0},      // K2 on CTA pairs (cta_group::2)
    {"cta_pairs", false, 0},      // K1 on CTA pairs
    {"weight_mcast", false, 0},   // 2 | 4: CTAs per cluster sharing one multicast weight stream
    {"epi_warps", false, 0},      // 8: eight epilogue warps per accumulator in K1
    {"fwd_kernel_ts", false, 0},  // 1: tensor-memory-operand forward kernel
    {"k3_overlap", false, 0},     // 1: K3 of sub-batch i on a second stream under K2 of sub-batch i + 1
    {"k3_sms", false, 0},         // SMs left to K3 when overlapped
    {"k3_items_per_sm_x10", false, 0},
    {"k3_grid", false, 0},        // CTAs of the weight-gradient kernel (0: one per SM)
    {"head_mma", false, 0},       // 0: rgb.2 weight gradient by the separate head kernel from the stash (default 1: inside K2)
    {"keep_weights", false, 0},   // 0: no L2 evict_last policy on the weight loads of the training kernel (default 1)
    {"stash_copy", false, 0},     // 1: the auxiliary warp copies the stash with ld.shared / st.global instead of TMA bulk stores (slower)
    {"stash_lanes", false, 0},    // bulk stores per stashed operand image (32 = 2 KB pieces; 1, 2, 4, 8, 16)
    {"share_fills", false, 0},    // 0: CTA pairs: every GEMM streams its own copy of the layer's weights (default 1: group Y reuses group X's fill)
    {"stash_early", false, 0},    // 0: one stash hand-off per operand image (after the whole epilogue) instead of one per 64-column block
    {"early_pieces", false, 0},   // bulk stores per signalled block of the early stash (1, 2, 4, 8)
    {"experiment", false, 0},     // timing experiments (bit mask, WRONG results): 1 no stash stores, 2 group Y reuses group X's weight fills
    {"stash_wrap", false, 0},     // timing experiment: stash tile t in slot t % value (stays in L2; gradients are WRONG)
};
std::mutex g_opt_mu;
}  // namespace

int64_t cnb_option(const char* name, int64_t dflt) {
    {
        std::lock_guard<std::mutex> lk(g_opt_mu);
        for (Opt& o : g_opts) if (!strcmp(o.name, name) && o.set) return o.value;
    }
    char env[64] = "CNB_";
    size_t n = 4;
    for (const char* c = name; *c && n + 1 < sizeof(env); ++c) env[n++] = (char)toupper((unsigned char)*c);
    env[n] = 0;
    const char* e = getenv(env);
    return e && *e ? (int64_t)atoll(e) : dflt;
}

extern "C" int cnb_set_option(const char* name, int64_t value) {
    if (!name) return CNB_E_INVALID;
    std::lock_guard<std::mutex> lk(g_opt_mu);
    for (Opt& o : g_opts) if (!strcmp(o.name, name)) { o.set = true; o.value = value; return CNB_OK; }
    return CNB_E_INVALID;
}

extern "C" int cnb_clear_option(const char* name) {
    if (!name) return CNB_E_INVALID;
    std::lock_guard<std::mutex> lk(g_opt_mu);
    for (Opt& o : g_opts) if (!strcmp(o.name, name)) { o.set = false; return CNB_OK; }
    return CNB_E_INVALID;
}

extern "C" int64_t cnb_get_option(const char* name, int64_t dflt) { return name ? cnb_option(name, dflt) : dflt; }

// ---- kernel timing aid --------------------------------------------------------------------
namespace {
constexpr int kProfCap = 512;
struct ProfState {
    bool on = false;
    cudaEvent_t ev[CNB_K_COUNT][kProfCap][2];
    bool made[CNB_K_COUNT][kProfCap] = {};
    int n[CNB_K_COUNT] = {};
} g_prof;
}  // namespace

void cnb_prof_begin(int k, cudaStream_t st) {
    if (!g_prof.on || g_prof.n[k] >= kProfCap) return;
    const int i = g_prof.n[k];
    if (!g_prof.made[k][i]) {
        cudaEventCreate(&g_prof.ev[k][i][0]); cudaEventCreate(&g_prof.ev[k][i][1]);
        g_prof.made[k][i] = true;
    }
    cudaEventRecord(g_prof.ev[k][i][0], st);
}
void cnb_prof_end(int k, cudaStream_t st) {
    if (!g_prof.on || g_prof.n[k] >= kProfCap) return;
    cudaEventRecord(g_prof.ev[k][g_prof.n[k]][1], st);
    g_prof.n[k]++;
}
extern "C" int cnb_profile_enable(int on) {
    g_prof.on = on != 0;
    for (int k = 0; k < CNB_K_COUNT; ++k) g_prof.n[k] = 0;
    return CNB_OK;
}
extern "C" int cnb_profile_read(int kernel_id, float* ms, int cap) {
    if (kernel_id < 0 || kernel_id >= CNB_K_COUNT || !ms || cap < 0) return CNB_E_INVALID;
    int n = g_prof.n[kernel_id] < cap ? g_prof.n[kernel_id] : cap;
    for (int i = 0; i < n; ++i) {
        if (cudaEventSynchronize(g_prof.ev[kernel_id][i][1]) != cudaSuccess) return -1;
        if (cudaEventElapsedTime(&ms[i], g_prof.ev[kernel_id][i][0], g_prof.ev[kernel_id][i][1]) != cudaSuccess) return -1;
    }
    g_prof.n[kernel_id] = 0;
    return n;
}

extern "C" int cnb_debug_pipeline_timeouts(void) { return cnb_sm100_pipeline_timeouts(); }

extern "C" int cnb_check_device(void) {
    int dev = 0;
    CNB_CUDA_TRY(cudaGetDevice(&dev));
    int major = 0;
    CNB_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    return major == 10 ? CNB_OK : CNB_E_DEVICE;
}

extern "C" int64_t cnb_param_count(const cnb_net_config* cfg) {
    if (cnb_validate_config(cfg) != CNB_OK) return -1;
    CnbLayout L; cnb_make_layout(cfg, &L);
    return L.total;
}

extern "C" int cnb_num_param_tensors(const cnb_net_config* cfg) {
    if (cnb_validate_config(cfg) != CNB_OK) return -1;
    CnbLayout L; cnb_make_layout(cfg, &L);
    return L.n_tensors;
}

extern "C" int cnb_param_layout(const cnb_net_config* cfg, int64_t* offsets, int32_t* rows, int32_t* cols) {
    CNB_TRY(cnb_validate_config(cfg));
    if (!offsets || !rows || !cols) return CNB_E_INVALID;
    CnbLayout L; cnb_make_layout(cfg, &L);
    const int W = cfg->W, LD = cfg->latent_dim;
    int t = 0;
    auto put = [&](int64_t woff, int64_t boff, int r, int c) {
        offsets[t] = woff; rows[t] = r; cols[t] = c; ++t;
        offsets[t] = boff; rows[t] = r; cols[t] = 1; ++t;
    };
    put(L.enc_xyz_w, L.enc_xyz_b, W, L.d_xyz);
    for (int j = 0; j < cfg->shape_blocks; ++j) { put(L.sl_w[j], L.sl_b[j], W, LD); put(L.s_w[j], L.s_b[j], W, W); }
    put(L.enc_shape_w, L.enc_shape_b, W, W);
    put(L.sigma_w, L.sigma_b, 1, W);
    put(L.enc_vd_w, L.enc_vd_b, W, W + L.d_dir);
    for (int j = 0; j < cfg->texture_blocks; ++j) { put(L.tl_w[j], L.tl_b[j], W, LD); put(L.t_w[j], L.t_b[j], W, W); }
    put(L.rgb0_w, L.rgb0_b, W / 2, W);
    put(L.rgb2_w, L.rgb2_b, 3, W / 2);
    return t == L.n_tensors ? CNB_OK : CNB_E_INVALID;
}

static int check_params(const cnb_net_config* cfg, const float* const* params) {
    CNB_TRY(cnb_validate_config(cfg));
    if (!params) return CNB_E_INVALID;
    CnbLayout L; cnb_make_layout(cfg, &L);
    for (int i = 0; i < L.n_tensors; ++i) if (!params[i]) return CNB_E_INVALID;
    return CNB_OK;
}

// ---------------------------------------------------------------------------
extern "C" size_t cnb_packed_weights_bytes(const cnb_net_config* cfg) {
    if (cnb_validate_config(cfg) != CNB_OK) return 0;
    return cnb_sm100_packed_bytes(cfg);
}

extern "C" int cnb_pack_weights(const cnb_net_config* cfg, const float* const* params, void* packed,
                                cnb_stream_t stream) {
    CNB_TRY(check_params(cfg, params));
    if (!packed) return CNB_E_INVALID;
    return cnb_sm100_pack_weights(cfg, params, packed, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------
extern "C" size_t cnb_mlp_workspace_bytes(const cnb_net_config* cfg, int64_t S, int n_codes, int precision,
                                          int backward) {
    if (cnb_validate_config(cfg) != CNB_OK || S <= 0 || n_codes < 1) return 0;
    if (precision == CNB_PRECISION_FP32) return cnb_fp32_workspace_bytes(cfg, S, 0, n_codes, backward, 0);
    size_t n = cnb_sm100_mlp_workspace_bytes(cfg, S, n_codes, backward);
    // codes that change inside a 128-row tile (e.g. one code per ray, [B,1,256]): the tensor-core backward needs one
    // code per tile and hands the call to the fp32 kernels (see cnb_mlp_backward)
    const bool ragged_codes = backward && n_codes > 1 && ((S / n_codes) % 128) != 0;
    if (backward && (!cnb_sm100_has_backward() || ragged_codes)) {
        const size_t m = cnb_fp32_workspace_bytes(cfg, S, 0, n_codes, backward, 0);
        if (m > n) n = m;
    }
    return n;
}

extern "C" int cnb_mlp_forward(const cnb_net_config* cfg, const float* const* params, const void* packed,
                               const float* xyz, const float* viewdir, const float* shape_codes,
                               const float* texture_codes, int n_codes, int64_t samples_per_code, int64_t S,
                               int precision, float* sigmas, float* rgbs, void* workspace, size_t workspace_bytes,
                               cnb_stream_t stream) {
    CNB_TRY(check_params(cfg, params));
    if (S < 0 || !xyz || !viewdir || !shape_codes || !texture_codes || n_codes < 1 || !sigmas || !rgbs)
        return CNB_E_INVALID;
    if (n_codes > 1 && samples_per_code < 1) return CNB_E_INVALID;
    if (S == 0) return CNB_OK;
    if (precision == CNB_PRECISION_FP32)
        return cnb_fp32_mlp_forward(cfg, params, xyz, viewdir, shape_codes, texture_codes, n_codes,
                                    n_codes > 1 ? samples_per_code : 0, S, sigmas, rgbs, workspace, workspace_bytes,
                                    (cudaStream_t)stream);
    if (precision != CNB_PRECISION_BF16) return CNB_E_INVALID;
    if (!packed) return CNB_E_INVALID;
    return cnb_sm100_mlp_forward(cfg, params, packed, xyz, viewdir, shape_codes, texture_codes, n_codes,
                                 n_codes > 1 ? samples_per_code : 0, S, sigmas, rgbs, workspace, workspace_bytes,
                                 (cudaStream_t)stream);
}

extern "C" int cnb_mlp_backward(const cnb_net_config* cfg, const float* const* params, const void* packed,
                                const float* xyz, const float* viewdir, const float* shape_codes,
                                const float* texture_codes, int n_codes, int64_t samples_per_code, int64_t S,
                                int precision, const float* d_sigmas, const float* d_rgbs, float* d_params,
                                float* d_shape_codes, float* d_texture_codes, void* workspace, size_t workspace_bytes,
                                cnb_stream_t stream) {
    CNB_TRY(check_params(cfg, params));
    if (S <= 0 || !xyz || !viewdir || !shape_codes || !texture_codes || n_codes < 1 || !d_sigmas || !d_rgbs ||
        !d_shape_codes || !d_texture_codes)
        return CNB_E_INVALID;
    if (n_codes > 1 && samples_per_code < 1) return CNB_E_INVALID;
    if (precision == CNB_PRECISION_FP32)
        return cnb_fp32_mlp_backward(cfg, params, xyz, viewdir, shape_codes, texture_codes, n_codes,
                                     n_codes > 1 ? samples_per_code : 0, S, d_sigmas, d_rgbs, d_params, d_shape_codes,
                                     d_texture_codes, workspace, workspace_bytes, (cudaStream_t)stream);
    if (precision != CNB_PRECISION_BF16) return CNB_E_INVALID;
    if (!packed) return CNB_E_INVALID;
    if (!cnb_sm100_has_backward())
        return cnb_fp32_mlp_backward(cfg, params, xyz, viewdir, shape_codes, texture_codes, n_codes,
                                     n_codes > 1 ? samples_per_code : 0, S, d_sigmas, d_rgbs, d_params, d_shape_codes,
                                     d_texture_codes, workspace, workspace_bytes, (cudaStream_t)stream);
    // The tensor-core backward folds the code-conditioned terms into one bias row per 128-row tile; when the codes change
    // inside a tile (samples_per_code not a multiple of 128: per-ray codes [B,1,256] with N = 64 or 96) the gradients are
    // computed by the fp32 kernels instead -- same mathematics, higher precision, not a throughput path (the
    // reference's own loops always broadcast one code over a whole chunk).
    if (n_codes > 1 && (samples_per_code % 128) != 0)
        return cnb_fp32_mlp_backward(cfg, params, xyz, viewdir, shape_codes, texture_codes, n_codes, samples_per_code, S,
                                     d_sigmas, d_rgbs, d_params, d_shape_codes, d_texture_codes, workspace, workspace_bytes,
                                     (cudaStream_t)stream);
    return cnb_sm100_mlp_backward(cfg, params, packed, xyz, viewdir, shape_codes, texture_codes, n_codes,
                                  n_codes > 1 ? samples_per_code : 0, S, d_sigmas, d_rgbs, d_params, d_shape_codes,
                                  d_texture_codes, workspace, workspace_bytes, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------
extern "C" size_t cnb_render_workspace_bytes(const cnb_net_config* cfg, const cnb_ray_batch* rays, int precision,
                                             int backward) {
    if (cnb_validate_config(cfg) != CNB_OK || cnb_validate_rays(rays) != CNB_OK) return 0;
    if (precision == CNB_PRECISION_FP32)
        return cnb_fp32_workspace_bytes(cfg, rays->n_rays * rays->n_samples, rays->n_samples, rays->n_codes, backward, 1);
    size_t n = cnb_sm100_render_workspace_bytes(cfg, rays, backward);
    const int64_t rows_per_code = (int64_t)rays->segments_per_code * rays->rays_per_segment * rays->n_samples;
    const bool ragged_codes = backward && rays->n_codes > 1 && (rows_per_code % 128) != 0;
    if (backward && (!cnb_sm100_has_backward() || ragged_codes)) {
        const size_t m = cnb_fp32_workspace_bytes(cfg, rays->n_rays * rays->n_samples, rays->n_samples, rays->n_codes, backward, 1);
        if (m > n) n = m;
    }
    return n;
}

static int render_dispatch(const cnb_net_config* cfg, const float* const* params, const void* packed,
                           const cnb_ray_batch* rays, int precision, int mode, const float* d_rgb,
                           const float* d_depth, const float* target, float loss_scale, float* rgb, float* depth,
                           float* acc, float* sq_err, float* d_params, float* d_shape, float* d_tex, void* ws,
                           size_t ws_bytes, cnb_stream_t stream) {
    CNB_TRY(check_params(cfg, params));
    CNB_TRY(cnb_validate_rays(rays));
    if (mode == 0 && (!rgb || !depth)) return CNB_E_INVALID;
    if (mode == 1 && (!d_rgb || !d_shape || !d_tex)) return CNB_E_INVALID;
    if (mode == 2 && (!target || !d_shape || !d_tex)) return CNB_E_INVALID;
    if (precision == CNB_PRECISION_FP32)
        return cnb_fp32_render(cfg, params, rays, mode, d_rgb, d_depth, target, loss_scale, rgb, depth, acc, sq_err,
                               d_params, d_shape, d_tex, ws, ws_bytes, (cudaStream_t)stream);
    if (precision != CNB_PRECISION_BF16) return CNB_E_INVALID;
    if (!packed) return CNB_E_INVALID;
    const int64_t rows_per_code = (int64_t)rays->segments_per_code * rays->rays_per_segment * rays->n_samples;
    if (mode != 0 && (!cnb_sm100_has_backward() || (rays->n_codes > 1 && (rows_per_code % 128) != 0)))   // see cnb_mlp_backward
        return cnb_fp32_render(cfg, params, rays, mode, d_rgb, d_depth, target, loss_scale, rgb, depth, acc, sq_err,
                               d_params, d_shape, d_tex, ws, ws_bytes, (cudaStream_t)stream);
    return cnb_sm100_render(cfg, params, packed, rays, mode, d_rgb, d_depth, target, loss_scale, rgb, depth, acc,
                            sq_err, d_params, d_shape, d_tex, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int cnb_render_forward(const cnb_net_config* cfg, const float* const* params, const void* packed,
                                  const cnb_ray_batch* rays, int precision, float* rgb, float* depth, float* acc,
                                  void* workspace, size_t workspace_bytes, cnb_stream_t stream) {
    return render_dispatch(cfg, params, packed, rays, precision, 0, nullptr, nullptr, nullptr, 1.f, rgb, depth, acc,
                           nullptr, nullptr, nullptr, nullptr, workspace, workspace_bytes, stream);
}

extern "C" int cnb_render_backward(const cnb_net_config* cfg, const float* const* params, const void* packed,
                                   const cnb_ray_batch* rays, int precision, const float* d_rgb, const float* d_depth,
                                   float* d_params, float* d_shape_codes, float* d_texture_codes, void* workspace,
                                   size_t workspace_bytes, cnb_stream_t stream) {
    return render_dispatch(cfg, params, packed, rays, precision, 1, d_rgb, d_depth, nullptr, 1.f, nullptr, nullptr,
                           nullptr, nullptr, d_params, d_shape_codes, d_texture_codes, workspace, workspace_bytes,
                           stream);
}

extern "C" int cnb_render_train_step(const cnb_net_config* cfg, const float* const* params, const void* packed,
                                     const cnb_ray_batch* rays, int precision, const float* target, float loss_scale,
                                     float* rgb, float* depth, float* acc, float* sq_err_sum, float* d_params,
                                     float* d_shape_codes, float* d_texture_codes, void* workspace,
                                     size_t workspace_bytes, cnb_stream_t stream) {
    return render_dispatch(cfg, params, packed, rays, precision, 2, nullptr, nullptr, target, loss_scale, rgb, depth,
                           acc, sq_err_sum, d_params, d_shape_codes, d_texture_codes, workspace, workspace_bytes,
                           stream);
}
