// sm100_common.cuh -- pieces shared by the tensor-core kernels (render_sm100.cu, backward_sm100.cu):
// tile constants, layer descriptors, the positional-encoding operand writer, the layer plan.
#pragma once
#include "render_sm100.cuh"
#include "mlp_fp32.cuh"
#include "umma.cuh"

namespace sm100 {


constexpr int kW = 256;                 // layer width the tensor-core path implements
constexpr int kLx = 10, kLd = 4;        // PE frequencies (63 / 27 channels)
constexpr int kTileRows = 128;          // UMMA M
constexpr int kSlot = 16384;            // one weight stage: [128 n x 64 k] bf16, 128-B swizzle
constexpr int kNumStages = 4;
constexpr int kABlock = 16384;          // activation K-block [128 rows x 64] bf16, 128-B swizzle
constexpr int kDirBlock = 8192;         // PE(viewdir) block [128 rows x 32] bf16, 64-B swizzle
constexpr int kATile = 4 * kABlock + kDirBlock;
constexpr int kThreads = 320;           // warp 0 producer, warp 1 MMA, warps 2-5 group X, 6-9 group Y
constexpr int kMaxLayers = 2 * CNB_MAX_BLOCKS + 4;

struct FwdLayer {
    uint32_t w_off;       // byte offset of the first stage slot inside the packed buffer
    uint8_t n_kchunks;    // 64-wide K chunks taken from activation blocks 0..n-1
    uint8_t has_dir;      // one more K=32 chunk from the PE(viewdir) block
    uint8_t n_halves;     // n_out / 128
    uint8_t relu;
    uint8_t kind;         // 0 hidden, 1 encoding_shape (+ sigma head), 2 rgb.0 (+ rgb head)
    int8_t folded;        // >= 0: bias row of the per-code folded table; < 0: shared bias
    uint16_t pad;
    const float* bias;    // shared bias [n_out] (fp32 parameter tensor)
};


__device__ __forceinline__ void st_shared_v4(uint8_t* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(umma::smem_u32(p)), "r"(a), "r"(b), "r"(c), "r"(d)
                 : "memory");
}

// PE of one row into the operand blocks -- reference src/model.py:4-7 (x, sines, cosines).
// sin/cos(2^i x) by exact doubling from an accurate sincosf(x): error < 2^i * 1e-7, far below bf16.
__device__ __forceinline__ void encode_row(const float p[3], const float v[3], bool valid, uint8_t* blk0,
                                           uint8_t* dirblk, int row) {
    float e[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) e[i] = 0.f;
    if (valid) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            e[k] = p[k];
            float s, c;
            sincosf(p[k], &s, &c);
#pragma unroll
            for (int i = 0; i < kLx; ++i) {
                e[3 + 3 * i + k] = s;
                e[3 + 3 * kLx + 3 * i + k] = c;
                const float s2 = 2.f * s * c;
                c = fmaf(-2.f * s, s, 1.f);
                s = s2;
            }
        }
    }
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
        st_shared_v4(blk0 + row * 128 + ((ch ^ (row & 7)) << 4), umma::pack_bf16(e[ch * 8 + 0], e[ch * 8 + 1]),
                     umma::pack_bf16(e[ch * 8 + 2], e[ch * 8 + 3]), umma::pack_bf16(e[ch * 8 + 4], e[ch * 8 + 5]),
                     umma::pack_bf16(e[ch * 8 + 6], e[ch * 8 + 7]));
    }
    float d[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) d[i] = 0.f;
    if (valid) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            d[k] = v[k];
            float s, c;
            sincosf(v[k], &s, &c);
#pragma unroll
            for (int i = 0; i < kLd; ++i) {
                d[3 + 3 * i + k] = s;
                d[3 + 3 * kLd + 3 * i + k] = c;
                const float s2 = 2.f * s * c;
                c = fmaf(-2.f * s, s, 1.f);
                s = s2;
            }
        }
    }
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
        st_shared_v4(dirblk + row * 64 + ((ch ^ ((row >> 1) & 3)) << 4), umma::pack_bf16(d[ch * 8 + 0], d[ch * 8 + 1]),
                     umma::pack_bf16(d[ch * 8 + 2], d[ch * 8 + 3]), umma::pack_bf16(d[ch * 8 + 4], d[ch * 8 + 5]),
                     umma::pack_bf16(d[ch * 8 + 6], d[ch * 8 + 7]));
    }
}


// ---------------------------------------------------------------------------
struct Plan {
    int n_layers, n_folded;
    FwdLayer fwd[kMaxLayers];
    int w_index[kMaxLayers];          // parameter-tensor index (state_dict order) of each layer's weight
    uint32_t bwd_w_off[kMaxLayers];   // W^T stage images (dgrad operand) of fwd layer l >= 1
    size_t fwd_bytes, total_bytes;
};

inline int make_plan(const cnb_net_config* c, const float* const* P, Plan* pl) {
    if (c->W != kW || c->num_xyz_freq != kLx || c->num_dir_freq != kLd) return CNB_E_UNSUPPORTED;
    CnbLayout L; cnb_make_layout(c, &L);
    int n = 0, nf = 0; size_t off = 0;
    auto add = [&](int wi, int kch, int dir, int halves, int relu, int kind, int folded) {
        FwdLayer f = {};
        f.w_off = (uint32_t)off; f.n_kchunks = (uint8_t)kch; f.has_dir = (uint8_t)dir; f.n_halves = (uint8_t)halves;
        f.relu = (uint8_t)relu; f.kind = (uint8_t)kind; f.folded = (int8_t)folded;
        f.bias = P ? P[wi + 1] : nullptr;
        off += (size_t)(kch + dir) * halves * kSlot;
        pl->w_index[n] = wi;
        pl->fwd[n++] = f;
    };
    add(L.i_enc_xyz, 1, 0, 2, 1, 0, -1);
    for (int j = 0; j < c->shape_blocks; ++j) add(L.i_s[j], 4, 0, 2, 1, 0, nf++);
    add(L.i_enc_shape, 4, 0, 2, 0, 1, -1);
    add(L.i_enc_vd, 4, 1, 2, 1, 0, -1);
    for (int j = 0; j < c->texture_blocks; ++j) add(L.i_t[j], 4, 0, 2, 1, 0, nf++);
    add(L.i_rgb0, 4, 0, 1, 1, 2, -1);
    pl->n_layers = n; pl->n_folded = nf; pl->fwd_bytes = off;
    pl->bwd_w_off[0] = 0;
    for (int l = 1; l < n; ++l) {     // dgrad operand of layer l: B[n = k_in (256)][k = n_out]
        pl->bwd_w_off[l] = (uint32_t)off;
        off += (size_t)(pl->fwd[l].n_halves * 2) * 2 * kSlot;
    }
    pl->total_bytes = off;
    return CNB_OK;
}

struct FwdWorkspace { float *z, *folded; float4* samples; size_t bytes; };

inline size_t carve_fwd(const cnb_net_config* c, int n_codes, int64_t spill_samples, char* base, FwdWorkspace* w) {
    size_t off = 0;
    auto take = [&](size_t bytes) -> char* { char* p = base ? base + off : nullptr; off += (bytes + 255) & ~(size_t)255; return p; };
    const int nf = c->shape_blocks + c->texture_blocks;
    FwdWorkspace x = {};
    x.z = (float*)take(sizeof(float) * (size_t)n_codes * nf * kW);
    x.folded = (float*)take(sizeof(float) * (size_t)n_codes * nf * kW);
    x.samples = (float4*)take(sizeof(float4) * (size_t)spill_samples);
    x.bytes = off;
    if (w) *w = x;
    return off;
}

// z_j = ReLU(latent layer_j(code)) and the folded per-code biases (defined in render_sm100.cu).
int latent_and_fold(const cnb_net_config* c, const float* const* P, const float* shape_codes, const float* tex_codes,
                    int n_codes, FwdWorkspace& w, cudaStream_t st);

// K1 over a ray sub-range (defined in render_sm100.cu).
int launch_render_rays(const cnb_net_config* cfg, const float* const* P, const void* packed, const Plan& pl,
                       const cnb_ray_batch* rays, int64_t ray_begin, int64_t ray_count, const float* folded,
                       float* spill_sig, float* spill_rgb, float* rgb, float* depth, float* acc, cudaStream_t st);
// backward (defined in backward_sm100.cu)
size_t bwd_workspace_bytes(const cnb_net_config* cfg, int64_t S, int64_t n_rays, int N, int n_codes, int fused);
int render_backward(const cnb_net_config* cfg, const float* const* P, const void* packed, const cnb_ray_batch* rays,
                    int mode, const float* d_rgb, const float* d_depth, const float* target, float loss_scale,
                    float* rgb, float* depth, float* acc, float* sq_err, float* d_params, float* d_shape, float* d_tex,
                    void* ws, size_t ws_bytes, cudaStream_t st);
}  // namespace sm100

int cnb_sm100_pipeline_timeouts_bwd(void);
// compositing backward with one z row per segment (rays.cu)
int cnb_vr_backward_segments(const float* sigmas, const float* rgbs, const float* z_vals, int z_per_segment,
                             int rays_per_segment, int64_t ray0, int64_t n_rays, int N, int white_bg,
                             const float* d_rgb, const float* d_depth, float* d_sigmas, float* d_rgbs, cudaStream_t st);
