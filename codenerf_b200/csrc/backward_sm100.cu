// backward_sm100.cu -- CNB_PRECISION_BF16 backward of the CodeNeRF MLP on Blackwell tensor cores.
//
// K2  k_mlp_bwd   : per 128-row tile, recompute the forward chain (activations only ever in
//      shared memory / TMEM, ReLU masks as bit words), then run the input-gradient chain
//      dY_l -> dY_{l-1} = (dY_l W_l) * relu'(.) through tcgen05.mma with the transposed weight
//      images.  Auxiliary warps reduce per-code column sums of every dY (bias and latent-code
//      gradients) and, for training, stream the bf16 operand tiles (layer inputs A_l and
//      pre-activation gradients dY_l) to HBM with TMA bulk stores.
// K3  k_wgrad     : dW_l = dY_l^T A_l over all rows, a split-K tcgen05 GEMM whose operands are
//      the stashed tiles read back as MN-major UMMA operands (same bytes, no transpose).
//
// Why the weight gradient is a second kernel: one layer's dW (256 x 256 fp32) fills all of TMEM,
// so it cannot be accumulated on chip next to the per-tile chain; see DESIGN.md section 4.
//
// Reference semantics: autograd of src/model.py:36-53 (what loss.backward() does in
// src/trainer.py:82 and src/optimizer.py:92).
#include <cstdlib>
#include "sm100_common.cuh"

using namespace sm100;

// Where K2 reads the rows every thread needs at the same address (see sm100_common.cuh, ConstRows): 1 = constant
// memory (uniform loads), 0 = shared memory (LDS broadcasts; the launch-wide rows are then staged like the folded ones).
// Measured (round 2, A/B in one session): constant memory is SLOWER for K2 -- 9.55 ms all shared, 9.9 heads constant,
// 10.0 biases constant, 10.5 both: the uniform loads (LDCU.128) are served at a lower rate than LDS broadcasts.
#ifndef CNB_K2_BIAS_CONST
#define CNB_K2_BIAS_CONST 0
#endif
#ifndef CNB_K2_HEADS_CONST
#define CNB_K2_HEADS_CONST 0
#endif
constexpr int kHeadSrc = CNB_K2_HEADS_CONST ? 2 : 1;
// 1: the epilogues of the 256-wide layers keep four tcgen05.ld in flight (software pipeline, see fwd_epilogue_layer).
// Measured on the CTA-pair kernel (round 2, same-session A/B): training +-0, latent fit 3 % SLOWER (a few spills with
// the prefetched encodings live); the TMEM latency is not what the epilogues wait for.  Off.
#ifndef CNB_K2_DEEP_LD
#define CNB_K2_DEEP_LD 0
#endif
constexpr bool kDeep = CNB_K2_DEEP_LD != 0;

#ifndef CNB_EARLY_NAMED
#define CNB_EARLY_NAMED 0      // 1: early-stash hand-offs on hardware named barriers instead of mbarriers (measured: no faster)
#endif
#ifdef CNB_TRACE
extern "C" int cnb_debug_events_bwd(unsigned long long* out, unsigned int* counts, int reset) {
    if (out && cudaMemcpyFromSymbol(out, sm100::g_events, sizeof(unsigned long long) * 4 * 16384) != cudaSuccess) return -1;
    if (counts && cudaMemcpyFromSymbol(counts, sm100::g_event_count, sizeof(unsigned int) * 4) != cudaSuccess) return -1;
    if (reset) { unsigned int z[4] = {}; if (cudaMemcpyToSymbol(sm100::g_event_count, z, sizeof(z)) != cudaSuccess) return -1; }
    return 0;
}
extern "C" int cnb_debug_trace_bwd(unsigned long long* out32, int reset) {
    if (out32 && cudaMemcpyFromSymbol(out32, sm100::g_trace, sizeof(unsigned long long) * 32) != cudaSuccess) return -1;
    if (reset) { unsigned long long z[32] = {}; if (cudaMemcpyToSymbol(sm100::g_trace, z, sizeof(z)) != cudaSuccess) return -1; }
    return 0;
}
#endif

namespace {

constexpr int kBwdThreads = 384;   // warp 0 producer, 1 MMA, 2/3 aux X/Y, 4-7 group X, 8-11 group Y

struct BwdStep {
    uint32_t w_off;       // W^T stage images of the layer this step back-propagates through (step >= 1)
    uint8_t n_kchunks;    // 64-wide K chunks of the incoming dY operand (2 or 4)
    int8_t mask_layer;    // fwd layer whose ReLU mask gates the result (-1: none)
    uint8_t add_sigma;    // result += dsigma_pre * w_sigma (gradient entering through the sigma head)
    uint8_t out_layer;    // fwd layer whose pre-activation gradient (dY) this step produces
    uint8_t out_blocks;   // 2 (128 wide) or 4 (256 wide)
    uint16_t pad;
};

struct BwdParams {
    WeightMaps maps;                // tensor maps of the packed weights (CTA-pair kernel)
    int n_layers, n_steps;
    FwdLayer layers[kMaxLayers];
    BwdStep steps[kMaxLayers];
    const uint8_t* packed;
    const float* folded;
    int n_folded, n_codes;
    int64_t rows_per_code;
    const float *w_sigma, *b_sigma, *w_rgb2;
    int mode;                       // 0: rows are samples of rays (rs), 1: xyz / viewdir arrays
    CnbRaySource rs;
    const float *xyz, *viewdir;
    int64_t S, row_offset;          // rows of this launch; global row of its first row (codes, rays)
    const float *d_sigmas, *d_rgbs; // seeds, launch-relative rows: [S], [S,3]
    uint32_t* mask_scratch;         // [grid][2 groups][1 + lookahead][n_layers][8][128]
    float* colsum;                  // [n_codes][n_layers][256] += column sums of dY_l
    int stash;                      // 1: stream operand tiles + dspre to HBM (training)
    // rgb.2 weight gradient inside K2 (training): d W_rgb2 = sum_rows d_rgb (x) r1 as one small MMA per tile -- A = the
    // rgb.2 input r1 still in the operand buffer (MN-major: its 128 columns are M), B = the tile's per-sample d_rgb (bf16
    // hi + lo) written into the idle PE(viewdir) block (MN-major, N = 32) -- read back from TMEM into three registers per
    // thread.  r1 is then neither stashed (32 KB per tile) nor re-read by a head kernel (0.47 ms per step).
    int head_mma;
    float *d_wrgb2, *d_brgb2;
    int stash_lanes;                // lanes of the auxiliary warp that issue the bulk stores of an operand image (32: 2 KB pieces)
    int keep_weights;               // 1: weight loads carry an L2 evict_last policy while the stash streams through L2
    int share_fills;                // CTA pairs: group Y's GEMM of a layer whose weights fill the ring exactly once (4 chunks) reuses the
                                    // fill of group X's GEMM of the same layer, issued just before it -- half the weight stream
    int stash_early;                // 1: every 64-column block of an operand image is handed to the auxiliary warp's bulk store as soon as
                                    // the epilogue has written it (4 signals per image instead of 1): the store drains during the epilogue
    int early_pieces;               // (stash_early) bulk stores per signalled part, one lane each
    int stash_copy;                 // 1: the auxiliary warp copies the image itself (ld.shared + st.global), off the TMA queue
    int experiment;                 // timing experiments (option experiment, WRONG results): bit 0 no stash stores, bit 1 no weight fills for group 1
    int64_t stash_wrap;             // timing experiment (option stash_wrap): tile t is stashed in slot t % stash_wrap (WRONG gradients)
    uint32_t colsum_layers;         // bit l: the aux warps reduce column sums of dY_l (training: none, K3 does it)
    uint8_t *stashA, *stashD;
    float* dspre;                   // [S] d(loss)/d(sigma pre-activation)
    // Work is handed out in UNITS of `unit_tiles` consecutive tiles that hold whole rays (unit_tiles * 128 is a
    // multiple of N).  With rays straddling tiles (unit_tiles > 1, e.g. N = 96: 4 rays = 3 tiles) a tile's backward
    // chain starts one tile late (`lookahead` = 1): F(0) F(1) B(0) F(2) B(1) ... B(U-1), so every ray touching the
    // tile has been composited; ReLU masks and the sample ring are double-buffered by tile parity.
    int unit_tiles, lookahead;
    // fused compositing (+ loss): the seeds are made in-kernel
    int fuse_comp;                  // 0: seeds from d_sigmas / d_rgbs; 1: from d_rgb_rays / d_depth_rays; 2: from target (L2)
    int white_bg;
    int64_t n_rays_total;
    const float *b_rgb2, *d_rgb_rays, *d_depth_rays, *target;   // per-ray arrays indexed by global ray
    float loss_scale;
    float *out_rgb, *out_depth, *out_acc, *sq_err;
    float* drgb_out;                // [S,3] per-sample d rgb written for the head weight gradient (fused + stash)
    uint32_t a_slot[kMaxLayers + 1];// byte offset in a tile's A stash of the input of layer l; [n_layers] = rgb.2 input
    uint32_t dir_slot;
    uint32_t d_slot[kMaxLayers];    // byte offset in a tile's dY stash of dY_l
    uint32_t a_tile_bytes, d_tile_bytes;
};

__device__ __forceinline__ uint4 ld_shared_v4(const uint8_t* p) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"(umma::smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t add_bf16x2(uint32_t a, uint32_t b) {
    uint32_t d; asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d;
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// Forward compositing, loss seed and reverse-mode compositing of ONE ray by one warp -- reference
// src/utils.py:34-47 and its autograd, the L2 seed of src/trainer.py:75.  The ray's per-sample (sigma, r, g, b) sit
// in the group's 256-entry sample ring (two tiles) at ring[(start + k) & 255], k < N <= 128, so a ray may straddle
// two tiles; the per-sample seeds (d sigma, d r, d g, d b) overwrite them in place.  Any N in [1, 128]: lane l owns
// samples [l * per, l * per + per) with per = ceil(N / 32); samples past N are identities of both scans.
template <int PER>
__device__ __noinline__ void composite_fwd_bwd_t(const BwdParams& p, float4* ring, int start, int64_t gray, int lane) {
    const int N = p.rs.N;
    constexpr int per = PER;
    const int i0 = lane * per;
    const bool live = gray < p.n_rays_total;
    const int64_t seg = (live ? gray : 0) / p.rs.rays_per_segment;
    const float* z = p.rs.z_vals + (p.rs.z_per_segment ? seg * N : 0);
    float alpha[PER], tt[PER], dl[PER], ex[PER], zz[PER];
    float4 s[PER];
    float tl = 1.f;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        alpha[j] = 0.f; tt[j] = 1.f; dl[j] = 0.f; ex[j] = 0.f; zz[j] = 0.f; s[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int i = i0 + j;
        if (j < per && i < N) {
            s[j] = ring[(start + i) & 255];
            zz[j] = __ldg(z + i);
            dl[j] = (i + 1 < N) ? (__ldg(z + i + 1) - zz[j]) : 1e10f;
            ex[j] = expf(-s[j].x * dl[j]);
            alpha[j] = 1.f - ex[j];
            tt[j] = 1.f - alpha[j] + 1e-10f;
            tl *= tt[j];
        }
    }
    // exclusive transmittance product across lanes
    float incl = tl;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const float up = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl *= up; }
    float T0 = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) T0 = 1.f;
    float Tj[PER], cr = 0.f, cg = 0.f, cb = 0.f, dep = 0.f, ws = 0.f;
    {
        float T = T0;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            Tj[j] = T;
            const float w = alpha[j] * T;
            cr += w * s[j].y; cg += w * s[j].z; cb += w * s[j].w; dep += w * zz[j]; ws += w;
            T *= tt[j];
        }
    }
    cr = warp_sum_f(cr); cg = warp_sum_f(cg); cb = warp_sum_f(cb); dep = warp_sum_f(dep); ws = warp_sum_f(ws);
    if (p.white_bg) { cr = cr + 1.f - ws; cg = cg + 1.f - ws; cb = cb + 1.f - ws; }
    // seed of the ray
    float gr = 0.f, gg = 0.f, gb = 0.f, gd = 0.f;
    if (live) {
        if (p.fuse_comp == 2) {
            const float er = cr - __ldg(p.target + gray * 3), eg = cg - __ldg(p.target + gray * 3 + 1),
                        eb = cb - __ldg(p.target + gray * 3 + 2);
            const float k = 2.f / (3.f * (float)p.rs.rays_per_segment) * p.loss_scale;
            gr = er * k; gg = eg * k; gb = eb * k;
            if (lane == 0 && p.sq_err) atomicAdd(p.sq_err + seg, er * er + eg * eg + eb * eb);
        } else {
            gr = __ldg(p.d_rgb_rays + gray * 3); gg = __ldg(p.d_rgb_rays + gray * 3 + 1); gb = __ldg(p.d_rgb_rays + gray * 3 + 2);
            if (p.d_depth_rays) gd = __ldg(p.d_depth_rays + gray);
        }
        if (lane == 0) {
            if (p.out_rgb) { p.out_rgb[gray * 3] = cr; p.out_rgb[gray * 3 + 1] = cg; p.out_rgb[gray * 3 + 2] = cb; }
            if (p.out_depth) p.out_depth[gray] = dep;
            if (p.out_acc) p.out_acc[gray] = ws;
        }
    }
    // reverse mode: aT_i = g_i alpha_i + t_i aT_{i+1}; d alpha_i = (g_i - aT_{i+1}) T_i   (division free)
    const float bg = p.white_bg ? 1.f : 0.f;
    float gsm[PER];
    float A = 1.f, Bc = 0.f;
#pragma unroll
    for (int j = PER - 1; j >= 0; --j) {
        gsm[j] = gr * (s[j].y - bg) + gg * (s[j].z - bg) + gb * (s[j].w - bg) + gd * zz[j];
        Bc = gsm[j] * alpha[j] + tt[j] * Bc;
        A = tt[j] * A;
    }
    float sA = A, sB = Bc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const float oA = __shfl_down_sync(0xffffffffu, sA, d);
        const float oB = __shfl_down_sync(0xffffffffu, sB, d);
        if (lane + d < 32) { sB = sA * oB + sB; sA = sA * oA; }
    }
    float aT_next = __shfl_down_sync(0xffffffffu, sB, 1);
    if (lane == 31) aT_next = 0.f;
#pragma unroll
    for (int j = PER - 1; j >= 0; --j) {
        const float w = alpha[j] * Tj[j];
        const float a_alpha = (gsm[j] - aT_next) * Tj[j];
        if (j < per && i0 + j < N) ring[(start + i0 + j) & 255] = make_float4(a_alpha * dl[j] * ex[j], w * gr, w * gg, w * gb);
        aT_next = gsm[j] * alpha[j] + aT_next * tt[j];
    }
}

// samples per lane = ceil(N / 32) as a compile-time constant: the per-lane loops carry no dead (predicated-off) iterations
__device__ __forceinline__ void composite_fwd_bwd(const BwdParams& p, float4* ring, int start, int64_t gray, int lane) {
    switch ((p.rs.N + 31) >> 5) {
        case 1: composite_fwd_bwd_t<1>(p, ring, start, gray, lane); break;
        case 2: composite_fwd_bwd_t<2>(p, ring, start, gray, lane); break;
        case 3: composite_fwd_bwd_t<3>(p, ring, start, gray, lane); break;
        default: composite_fwd_bwd_t<4>(p, ring, start, gray, lane); break;
    }
}

// Input-gradient epilogue of 32 accumulator columns: [+ dsigma_pre * w_sigma] [* ReLU mask bits] -> bf16 operand.
template <int CC, bool HAS_MASK, bool ADD_SIGMA>
__device__ __forceinline__ void bwd_epilogue32(const uint32_t (&rr)[32], const uint32_t (&a8)[8], uint32_t mw,
                                               uint64_t dsp2, const float* __restrict__ w_sigma) {
#pragma unroll
    for (int j8 = 0; j8 < 4; ++j8) {
        const int col = CC * 32 + j8 * 8;
        uint64_t v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = pk2(rr[j8 * 8 + 2 * i], rr[j8 * 8 + 2 * i + 1]);
        if (ADD_SIGMA) {
            const float4 w0 = ld_vec4<kHeadSrc>(w_sigma + col);  // sigma-head weights: constant memory (crow_ptr) or shared
            const float4 w1 = ld_vec4<kHeadSrc>(w_sigma + col + 4);
            v[0] = ffma2(dsp2, pk2f(w0.x, w0.y), v[0]); v[1] = ffma2(dsp2, pk2f(w0.z, w0.w), v[1]);
            v[2] = ffma2(dsp2, pk2f(w1.x, w1.y), v[2]); v[3] = ffma2(dsp2, pk2f(w1.z, w1.w), v[3]);
        }
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] = cvt_bf16x2<false>(v[i]);
#ifndef CNB_EXPERIMENT_NOMASK     // (timing experiment only, wrong gradients)
        if (HAS_MASK) {         // columns 8 j8 + 2 i, + 1: bits of word << (2 j8 + (i >> 1)), byte pair i & 1 (relu_mask_pair)
            const uint32_t x0 = mw << (2 * j8), x1 = mw << (2 * j8 + 1);
            w[0] &= relu_mask_pair(x0, 0); w[1] &= relu_mask_pair(x0, 1);
            w[2] &= relu_mask_pair(x1, 0); w[3] &= relu_mask_pair(x1, 1);
        }
#endif
        constexpr int blk = CC >> 1;
        const int chunk = ((CC & 1) << 2) + j8;
        st_shared_v4_off<blk * kABlock>(a8[chunk], w[0], w[1], w[2], w[3]);
    }
}
template <bool HAS_MASK, bool ADD_SIGMA, class Hook>
__device__ __forceinline__ void bwd_epilogue_layer(uint32_t taddr, const uint32_t (&a8)[8], const uint32_t* mscr,
                                                   uint64_t dsp2, const float* __restrict__ w_sigma, uint64_t pol, Hook hook) {
    if constexpr (kDeep) {
        // software pipeline: the next two tcgen05.ld (and their ReLU bit words) are issued before the current two chunks
        // are processed; only the first wait exposes the TMEM latency
        uint32_t ra[32], rb[32], rc[32], rd[32];
        uint32_t m[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) m[c] = HAS_MASK ? umma::ld_global_hint(mscr + (size_t)c * kTileRows, pol) : 0xffffffffu;
        umma::tmem_ld32(taddr + 0, ra);
        umma::tmem_ld32(taddr + 32, rb);
        umma::tmem_ld_wait();
        tmem_regs_ready(ra); tmem_regs_ready(rb);
        umma::tmem_ld32(taddr + 64, rc);
        umma::tmem_ld32(taddr + 96, rd);
        bwd_epilogue32<0, HAS_MASK, ADD_SIGMA>(ra, a8, m[0], dsp2, w_sigma);
        bwd_epilogue32<1, HAS_MASK, ADD_SIGMA>(rb, a8, m[1], dsp2, w_sigma);
        hook(0);
        umma::tmem_ld_wait();
        tmem_regs_ready(rc); tmem_regs_ready(rd);
        umma::tmem_ld32(taddr + 128, ra);
        umma::tmem_ld32(taddr + 160, rb);
        bwd_epilogue32<2, HAS_MASK, ADD_SIGMA>(rc, a8, m[2], dsp2, w_sigma);
        bwd_epilogue32<3, HAS_MASK, ADD_SIGMA>(rd, a8, m[3], dsp2, w_sigma);
        hook(1);
        umma::tmem_ld_wait();
        tmem_regs_ready(ra); tmem_regs_ready(rb);
        umma::tmem_ld32(taddr + 192, rc);
        umma::tmem_ld32(taddr + 224, rd);
        bwd_epilogue32<4, HAS_MASK, ADD_SIGMA>(ra, a8, m[4], dsp2, w_sigma);
        bwd_epilogue32<5, HAS_MASK, ADD_SIGMA>(rb, a8, m[5], dsp2, w_sigma);
        hook(2);
        umma::tmem_ld_wait();
        tmem_regs_ready(rc); tmem_regs_ready(rd);
        bwd_epilogue32<6, HAS_MASK, ADD_SIGMA>(rc, a8, m[6], dsp2, w_sigma);
        bwd_epilogue32<7, HAS_MASK, ADD_SIGMA>(rd, a8, m[7], dsp2, w_sigma);
        return;
    }
    auto pair = [&](auto cc_tag) {
        constexpr int CC = decltype(cc_tag)::value;
        uint32_t ra[32], rb[32];
        umma::tmem_ld32(taddr + CC * 32, ra);
        umma::tmem_ld32(taddr + CC * 32 + 32, rb);
        const uint32_t m0 = HAS_MASK ? umma::ld_global_hint(mscr + (size_t)CC * kTileRows, pol) : 0xffffffffu;
        const uint32_t m1 = HAS_MASK ? umma::ld_global_hint(mscr + (size_t)(CC + 1) * kTileRows, pol) : 0xffffffffu;
        if constexpr (CC > 0) hook((CC >> 1) - 1);      // the previous block is handed over while this pair's loads travel
        umma::tmem_ld_wait();
        bwd_epilogue32<CC, HAS_MASK, ADD_SIGMA>(ra, a8, m0, dsp2, w_sigma);
        bwd_epilogue32<CC + 1, HAS_MASK, ADD_SIGMA>(rb, a8, m1, dsp2, w_sigma);
        // (block 3 is handed over by the caller together with the operand's publication)
    };
    pair(std::integral_constant<int, 0>{});
    pair(std::integral_constant<int, 2>{});
    pair(std::integral_constant<int, 4>{});
    pair(std::integral_constant<int, 6>{});
}

// MC > 1: clusters of MC CTAs share one multicast weight stream (see produce_stages); every CTA of a cluster runs
// the same number of unit slots, the surplus ones as phantom units (all rows invalid, nothing written).
// CG = 2: CTA pairs (tcgen05 cta_group::2, M = 256 = one tile of each CTA): every CTA streams half of each weight
// chunk and reads half of the B operand; the leader CTA issues the MMAs, the epilogue warps of both CTAs arrive on
// its a_ready barriers, accumulator / ring-slot commits are multicast to both.
//
// Schedule of one compute group over one unit of U tiles (LA = lookahead, 0 or 1):
//     for i in [0, U + LA):   if (i < U) F(i);   if (i >= LA) B(i - LA);
// F(t) = positional encodings + the nl forward GEMMs of tile t (+ compositing of every ray whose last sample lies
// in t), B(t) = step 0 + the ns - 1 input-gradient GEMMs of tile t.  The producer, the MMA issuer and the auxiliary
// warps walk the same sequence; the two groups' GEMMs alternate op by op.
template <int MC, int CG = 1>
__global__ void __launch_bounds__(kBwdThreads, 1) k_mlp_bwd(const __grid_constant__ BwdParams p) {
    static_assert(CG == 1 || MC == 1, "pairs and multicast clusters are alternatives");
    constexpr int kCluster = CG * MC;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA0 = smem;
    uint8_t* sW = smem + 2 * kATile;
    uint64_t* bars = (uint64_t*)(sW + kNumStages * kSlot);
    uint64_t* w_full = bars;
    uint64_t* w_empty = bars + kNumStages;
    uint64_t* a_ready = bars + 2 * kNumStages;   // [2] operand of the next MMA is in shared memory
    uint64_t* acc_full = a_ready + 2;            // [2] accumulator complete
    uint64_t* aux_ready = acc_full + 2;          // [2] a tile operand was (re)written (every phase)
    uint64_t* buf_free = aux_ready + 2;          // [2] aux warp finished reading the operand buffer
    uint64_t* aux_blk = buf_free + 2;            // [2][3] (stash_early) block 1..3 of the operand image was written (block 0: aux_ready)
    uint32_t* tmem_slot = (uint32_t*)(aux_blk + 6);
    float4* sRing = (float4*)(tmem_slot + 4);      // [2 groups][256] per-sample (sigma, r, g, b), then seeds: two tiles per group
    float* sBias = (float*)(sRing + 4 * kTileRows);  // [2 groups][2 buffers][256] per-code (folded) bias row of the layer being drained
    float* sWsig = sBias + 4 * kW;                  // (CNB_K2_HEADS_CONST == 0) [256] sigma-head weights, [3][128] rgb.2 weights
    float* sWrgb = sWsig + kW;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nl = p.n_layers, ns = p.n_steps;
    const int U = p.unit_tiles, LA = p.lookahead;

    const int64_t tiles = (p.S + kTileRows - 1) / kTileRows;
    const int64_t units = (tiles + U - 1) / U;
    const int64_t ubase = units / gridDim.x, urem = units % gridDim.x;
    const int64_t unit0 = (int64_t)blockIdx.x * ubase + min((int64_t)blockIdx.x, urem);
    const int UN_own = (int)(ubase + ((int64_t)blockIdx.x < urem ? 1 : 0));
    const uint32_t rank = kCluster > 1 ? umma::cluster_ctarank() : 0u;
    const int UN = kCluster > 1 ? (int)(ubase + ((int64_t)(blockIdx.x - rank) < urem ? 1 : 0)) : UN_own;   // cluster-uniform
    const int rounds = (UN + 1) >> 1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kNumStages; ++i) { umma::mbar_init(&w_full[i], 1); umma::mbar_init(&w_empty[i], MC); }
        for (int g = 0; g < 2; ++g) {
            umma::mbar_init(&a_ready[g], 4 * CG); umma::mbar_init(&acc_full[g], 1);
            umma::mbar_init(&aux_ready[g], 4); umma::mbar_init(&buf_free[g], 1);
            for (int b = 0; b < 3; ++b) umma::mbar_init(&aux_blk[g * 3 + b], 4);
        }
        umma::fence_mbar_init();
    }
    if (warp == 1) { if (CG == 2) umma::tmem_alloc2(tmem_slot, 512); else umma::tmem_alloc(tmem_slot, 512); }
    if (!CNB_K2_HEADS_CONST) {
        for (int i = threadIdx.x; i < kW; i += kBwdThreads) sWsig[i] = __ldg(p.w_sigma + i);
        for (int i = threadIdx.x; i < 3 * (kW / 2); i += kBwdThreads) sWrgb[i] = __ldg(p.w_rgb2 + i);
    }
    umma::tc_fence_before();
    if (kCluster > 1) umma::cluster_sync_all(); else __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // a_ready lives in the leader CTA of a pair: the partner's epilogue warps arrive on it remotely
    const uint32_t a_ready_addr0 = CG == 2 ? umma::mapa(umma::smem_u32(&a_ready[0]), 0) : 0u;

    if (warp < 4) {
    umma::setmaxnreg_dec<kRegsAux>();
    if (warp == 0) {
        // ===== weight producer =====
        int stage = 0; uint32_t ph = 0;
        const uint64_t pol_w = (p.stash && p.keep_weights) ? umma::l2_policy_evict_last() : 0ull;   // the stash stream must not evict the weights
        for (int r = 0; r < rounds; ++r)
            for (int i = 0; i < U + LA; ++i) {
                if (i < U)
                    for (int op = 0; op < nl; ++op)
                        for (int g = 0; g < 2; ++g) {
                            if (2 * r + g >= UN) continue;
                            const FwdLayer& L = p.layers[op];
                            const int n_dir = L.has_dir ? L.n_halves : 0;
                            if (CG == 2) produce_stages_2cta(&p.maps, L.w_off, L.n_kchunks, L.n_halves, L.has_dir, rank, sW, w_full, w_empty, stage, ph, pol_w,
                                                             g == 1 && ((p.experiment & 2) || (p.share_fills && L.n_kchunks + (L.has_dir ? 1 : 0) == kNumStages)));
                            else produce_stages<MC>(p.packed + L.w_off, L.n_kchunks * L.n_halves + n_dir, n_dir, sW, w_full, w_empty, stage, ph, rank, nullptr, pol_w);
                        }
                if (i >= LA)
                    for (int s = 1; s < ns; ++s)
                        for (int g = 0; g < 2; ++g) {
                            if (2 * r + g >= UN) continue;
                            const BwdStep& B = p.steps[s];
                            if (CG == 2) produce_stages_2cta(&p.maps, B.w_off, B.n_kchunks, 2, 0, rank, sW, w_full, w_empty, stage, ph, pol_w,
                                                             g == 1 && ((p.experiment & 2) || (p.share_fills && B.n_kchunks == kNumStages)));
                            else produce_stages<MC>(p.packed + B.w_off, B.n_kchunks * 2, 0, sW, w_full, w_empty, stage, ph, rank, nullptr, pol_w);
                        }
            }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        int stage = 0; uint32_t ph = 0;
        uint32_t cnt[2] = {0u, 0u};               // operands consumed so far per group (a_ready parity)
        CNB_TR_DECL(tr_wa); CNB_TR_DECL(tr_ww); CNB_TR_DECL(tr_tot);
        const long long tr_t0 = CNB_TR_NOW();
        auto issue = [&](int g, int n_kchunks, int n_halves, int has_dir, int ev_op, uint32_t a_shift = 0u) {
            const uint32_t par = cnt[g] & 1u; ++cnt[g];
            if (CG == 2 && rank != 0) return;            // the partner CTA issues no MMAs
            if (CG == 2) CNB_TR(tr_wa, umma::mbar_wait_cluster(&a_ready[g], par));
            else CNB_TR(tr_wa, umma::mbar_wait(&a_ready[g], par));
            CNB_EV(lane, 0, (1 << 12) | (g << 8) | ev_op);          // operands ready, issue starts
            umma::tc_fence_after();
            if (CG == 2)
                issue_gemm_2cta(umma::smem_u32(sA0 + g * kATile) + a_shift, tmem + (uint32_t)g * 256u, sW, w_full, w_empty, n_kchunks,
                                n_halves, has_dir, stage, ph, &acc_full[g], &tr_ww);
            else
                issue_gemm<MC>(umma::smem_u32(sA0 + g * kATile) + a_shift, tmem + (uint32_t)g * 256u, sW, w_full, w_empty, n_kchunks,
                               n_halves, has_dir, stage, ph, &acc_full[g], &tr_ww);
            CNB_EV(lane, 0, (2 << 12) | (g << 8) | ev_op);          // all MMAs of the op issued
        };
        // rgb.2 weight gradient of one tile: D[m = r1 column][n] = sum_rows r1[row][m] * seeds[row][n], contraction over the
        // tile's 128 rows (8 K steps of 16 rows).  Both operands are MN-major views of shared-memory images that exist
        // anyway (the same views K3 takes of the stash).  On CTA pairs (M = 256, N = 64) each CTA supplies its own r1 as its
        // half of M and its own seeds as its half of N, and reads only its own half of the N columns of D: the other half
        // pairs its r1 with the partner's seeds and is ignored.
        auto issue_head = [&](int g) {
            const uint32_t par = cnt[g] & 1u; ++cnt[g];
            if (CG == 2 && rank != 0) return;
            if (CG == 2) CNB_TR(tr_wa, umma::mbar_wait_cluster(&a_ready[g], par));
            else CNB_TR(tr_wa, umma::mbar_wait(&a_ready[g], par));
            umma::tc_fence_after();
            if (umma::elect_one()) {
                const uint32_t a_base = umma::smem_u32(sA0 + g * kATile);
                const uint64_t dA = umma::make_sdesc(a_base, kABlock, 1024, umma::SWZ_128B);              // two 64-column spans, 16 KB apart
                const uint64_t dB = umma::make_sdesc(a_base + 4 * kABlock, kDirBlock / 2, 512, umma::SWZ_64B);
                const uint32_t idesc = CG == 2 ? umma::make_idesc(256, 64, 1, 1) : umma::make_idesc(128, 32, 1, 1);
                const uint32_t d = tmem + (uint32_t)g * 256u;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    if (CG == 2) umma::mma_bf16_2cta(d, dA + (uint64_t)(ks * 128), dB + (uint64_t)(ks * 64), idesc, ks > 0 ? 1u : 0u);
                    else umma::mma_bf16(d, dA + (uint64_t)(ks * 128), dB + (uint64_t)(ks * 64), idesc, ks > 0 ? 1u : 0u);
                }
                if (CG == 2) umma::mma_commit_2cta(&acc_full[g], 3); else umma::mma_commit(&acc_full[g]);
            }
            __syncwarp();
        };
        for (int r = 0; r < rounds; ++r)
            for (int i = 0; i < U + LA; ++i) {
                if (i < U)
                    for (int op = 0; op < nl; ++op)
                        for (int g = 0; g < 2; ++g)
                            if (2 * r + g < UN) issue(g, p.layers[op].n_kchunks, p.layers[op].n_halves, p.layers[op].has_dir, op);
                if (i >= LA) {
                    if (p.stash && p.head_mma)
                        for (int g = 0; g < 2; ++g)
                            if (2 * r + g < UN) issue_head(g);
                    for (int s = 1; s < ns; ++s)
                        for (int g = 0; g < 2; ++g)
                            if (2 * r + g < UN) issue(g, p.steps[s].n_kchunks, 2, 0, nl + s - 1, (s == 1 && p.stash && p.head_mma) ? 2u * kABlock : 0u);
                }
            }
        tr_tot = (unsigned long long)(CNB_TR_NOW() - tr_t0);
        CNB_TR_FLUSH(0, tr_wa); CNB_TR_FLUSH(1, tr_ww); CNB_TR_FLUSH(2, tr_tot);
    } else {
        // ===== auxiliary warps: operand stash (TMA bulk stores) and column sums of every dY =====
        const int g = warp - 2;
        const uint8_t* sA = sA0 + g * kATile;
        uint32_t ap = 0;
        const uint64_t pol_stream = (p.experiment & 8) ? umma::l2_policy_evict_normal() : (p.experiment & 16) ? umma::l2_policy_evict_last()
                                                                                                   : umma::l2_policy_evict_first();   // the stash is written once and read by a later kernel
        CNB_TR_DECL(tr_wx); CNB_TR_DECL(tr_tot);
        const long long tr_t0 = CNB_TR_NOW();
        // one operand-buffer phase: wait for the write, stash it and / or reduce its column sums, release the buffer
        const bool early = p.stash && p.stash_early && !p.stash_copy;
        uint32_t apb[3] = {0u, 0u, 0u};       // phases consumed of aux_blk[g][0..2]
        auto phase = [&](int64_t tile, bool live, int phs) {
            // (stash_early: the blocks are waited for one by one below, block 0 not necessarily first)
            if (!early) { CNB_TR(tr_wx, umma::mbar_wait(&aux_ready[g], ap & 1u)); ++ap; }
            const bool stash = p.stash && live;
            if (p.stash_wrap > 0) tile %= p.stash_wrap;
            // with the in-kernel rgb.2 gradient step 0 writes blocks 2-3 (blocks 0-1 still hold r1 for that MMA)
            const uint8_t* src = sA + ((phs == nl + 1 && p.stash && p.head_mma) ? 2 * kABlock : 0);
            int blocks = 0, out_layer = -1;
            uint8_t* dst = nullptr;
            if (phs == 0) { blocks = 1; dst = p.stashA + (size_t)tile * p.a_tile_bytes + p.a_slot[0]; }
            else if (phs < nl) { blocks = 4; dst = p.stashA + (size_t)tile * p.a_tile_bytes + p.a_slot[phs]; }
            else if (phs == nl) { blocks = 2; dst = p.stashA + (size_t)tile * p.a_tile_bytes + p.a_slot[nl]; }
            else {
                const BwdStep& B = p.steps[phs - nl - 1];
                blocks = B.out_blocks; out_layer = B.out_layer;
                dst = p.stashD + (size_t)tile * p.d_tile_bytes + p.d_slot[B.out_layer];
            }
            if (stash && p.stash_copy) {
                // EXPERIMENT (option stash_copy, off): the warp copies the image itself, 512 coalesced bytes per instruction, 8
                // loads in flight.  Idea: bulk stores share the SM's TMA queue with the weight refills (in training the MMA warp
                // waits 3.5x longer for weights than in the latent fit), plain stores do not.  Measured: K2 9.9 vs 8.6 ms --
                // one warp cannot move 64 KB per tile-op through the LSU fast enough.
                const uint4* s4 = reinterpret_cast<const uint4*>(src) + lane;
                uint4* d4 = reinterpret_cast<uint4*>(dst) + lane;
                const int n512 = blocks * (kABlock / 512);
                for (int i = 0; i < n512; i += 8) {
                    uint4 v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) v[u] = ld_shared_v4(reinterpret_cast<const uint8_t*>(s4 + (size_t)(i + u) * 32));
#pragma unroll
                    for (int u = 0; u < 8; ++u) umma::st_global_v4_hint(d4 + (size_t)(i + u) * 32, v[u], pol_stream);
                }
                if (phs == 0) {
                    const uint4* sd = reinterpret_cast<const uint4*>(sA + 4 * kABlock) + lane;
                    uint4* dd = reinterpret_cast<uint4*>(p.stashA + (size_t)tile * p.a_tile_bytes + p.dir_slot) + lane;
#pragma unroll
                    for (int u = 0; u < kDirBlock / 512; ++u)
                        umma::st_global_v4_hint(dd + (size_t)u * 32, ld_shared_v4(reinterpret_cast<const uint8_t*>(sd + (size_t)u * 32)), pol_stream);
                }
            } else if (early) {
                // block by block, as the epilogue signals them (block 0 arrived with aux_ready): the image drains while the
                // rest of it is still being written, so the buffer is free again by the time the next epilogue needs it
                const int n_sig = (p.stash_early == 2 && blocks > 1) ? blocks / 2 : blocks;      // (stash_early 2: one signal per two blocks)
                const uint32_t part = (uint32_t)(blocks / n_sig) * kABlock;
                for (int b = 0; b < n_sig; ++b) {
#if CNB_EARLY_NAMED
                    // hardware named barriers 4 + 4 g + b: 128 epilogue threads arrive, this warp syncs (measured: no faster)
                    CNB_TR(tr_wx, umma::named_bar_sync(4 + 4 * g + b, 160));
#else
                    if (b > 0) { CNB_TR(tr_wx, umma::mbar_wait(&aux_blk[g * 3 + b - 1], apb[b - 1] & 1u)); ++apb[b - 1]; }
                    else { CNB_TR(tr_wx, umma::mbar_wait(&aux_ready[g], ap & 1u)); ++ap; }
#endif
                    if (stash && !(p.experiment & 1) && lane < p.early_pieces) {
                        const uint32_t piece = part / (uint32_t)p.early_pieces;
                        umma::bulk_s2g_hint(dst + (size_t)b * part + (size_t)lane * piece, src + (size_t)b * part + (size_t)lane * piece, piece, pol_stream);
                        if (phs == 0 && lane == 0) umma::bulk_s2g_hint(p.stashA + (size_t)tile * p.a_tile_bytes + p.dir_slot, sA + 4 * kABlock, kDirBlock, pol_stream);
                        umma::bulk_commit();
                    }
                    __syncwarp();
                }
            } else if (stash && !(p.experiment & 1)) {
                // every lane stores 1/32 of the image: short bulk stores let the weight loads that share this
                // SM's copy engine slip in between (one 64 KB store ahead of a refill stalls the MMA ring)
                const uint32_t piece = (uint32_t)blocks * (kABlock / p.stash_lanes);
                if (lane < p.stash_lanes) {
                    umma::bulk_s2g_hint(dst + (size_t)lane * piece, src + (size_t)lane * piece, piece, pol_stream);
                    if (phs == 0)
                        umma::bulk_s2g_hint(p.stashA + (size_t)tile * p.a_tile_bytes + p.dir_slot + lane * (kDirBlock / p.stash_lanes),
                                            sA + 4 * kABlock + lane * (kDirBlock / p.stash_lanes), kDirBlock / p.stash_lanes, pol_stream);
                    umma::bulk_commit();
                }
            }
            if (live && out_layer >= 0 && ((p.colsum_layers >> out_layer) & 1u) && lane < blocks * 8) {
                int64_t code = 0;
                if (p.n_codes > 1) { code = (p.row_offset + tile * kTileRows) / p.rows_per_code; if (code >= p.n_codes) code = p.n_codes - 1; }
                const int blk = lane >> 3, chunk = lane & 7;
                float acc[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = 0.f;
                const uint8_t* base = src + blk * kABlock;
                // four rows are summed as packed bf16 pairs first (two roundings of the size the operands already carry),
                // then widened: 7 instead of 17 instructions per row -- this warp's column sums were what the epilogue of
                // the next step waited for in the latent fit (buf_free)
#pragma unroll 2
                for (int rr = 0; rr < kTileRows; rr += 4) {
                    const uint4 w0 = ld_shared_v4(base + (rr + 0) * 128 + ((chunk ^ ((rr + 0) & 7)) << 4));
                    const uint4 w1 = ld_shared_v4(base + (rr + 1) * 128 + ((chunk ^ ((rr + 1) & 7)) << 4));
                    const uint4 w2 = ld_shared_v4(base + (rr + 2) * 128 + ((chunk ^ ((rr + 2) & 7)) << 4));
                    const uint4 w3 = ld_shared_v4(base + (rr + 3) * 128 + ((chunk ^ ((rr + 3) & 7)) << 4));
                    const uint32_t sx = add_bf16x2(add_bf16x2(w0.x, w1.x), add_bf16x2(w2.x, w3.x));
                    const uint32_t sy = add_bf16x2(add_bf16x2(w0.y, w1.y), add_bf16x2(w2.y, w3.y));
                    const uint32_t sz = add_bf16x2(add_bf16x2(w0.z, w1.z), add_bf16x2(w2.z, w3.z));
                    const uint32_t sw = add_bf16x2(add_bf16x2(w0.w, w1.w), add_bf16x2(w2.w, w3.w));
                    acc[0] += bf_lo(sx); acc[1] += bf_hi(sx); acc[2] += bf_lo(sy); acc[3] += bf_hi(sy);
                    acc[4] += bf_lo(sz); acc[5] += bf_hi(sz); acc[6] += bf_lo(sw); acc[7] += bf_hi(sw);
                }
                float* out = p.colsum + ((size_t)code * nl + out_layer) * kW + blk * 64 + chunk * 8;
#pragma unroll
                for (int i = 0; i < 8; ++i) atomicAdd(out + i, acc[i]);
            }
            if (stash && !p.stash_copy && !(p.experiment & 5)) umma::bulk_wait_read_all();      // (bit 2: release the buffer without waiting)
            __syncwarp();
            if (lane == 0) umma::mbar_arrive(&buf_free[g]);
        };
        for (int r = 0; r < rounds; ++r) {
            const int u = 2 * r + g;
            if (u >= UN) break;
            const int64_t utile0 = (unit0 + u) * U;
            for (int i = 0; i < U + LA; ++i) {
                if (i < U) {
                    const int64_t tile = utile0 + i;
                    const bool live = u < UN_own && tile < tiles;
                    for (int phs = 0; phs <= nl; ++phs) {
                        if (phs == nl && (!p.stash || p.head_mma)) continue;      // the rgb.2 input is only stashed for the head kernel
                        phase(tile, live, phs);
                    }
                }
                if (i >= LA) {
                    const int64_t tile = utile0 + i - LA;
                    const bool live = u < UN_own && tile < tiles;
                    for (int s = 0; s < ns; ++s) phase(tile, live, nl + 1 + s);
                }
            }
        }
        if (p.stash && !p.stash_copy) umma::bulk_wait_all();
        tr_tot = (unsigned long long)(CNB_TR_NOW() - tr_t0);
        if (g == 0) { CNB_TR_FLUSH(3, tr_wx); CNB_TR_FLUSH(4, tr_tot); }
    }
    } else {
        umma::setmaxnreg_inc<kRegsCompute>();
        // ===== compute groups =====
        const int g = (warp - 4) >> 2;
        const int q = warp & 3;
        const int row = q * 32 + lane;
        uint8_t* sA = sA0 + g * kATile;
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)g * 256u;
        const int N = p.rs.N;
        uint32_t a8[8];       // shared address of each 16-byte chunk of this row inside K-block 0
#pragma unroll
        for (int c = 0; c < 8; ++c) a8[c] = umma::smem_u32(sA + row * 128 + ((c ^ (row & 7)) << 4));
        const uint64_t pol_keep = umma::l2_policy_evict_last();   // ReLU masks: written now, re-read ~50 us later
        uint32_t wp = 0;      // operand-buffer write phases so far (buf_free bookkeeping)
        uint32_t opc = 0;     // accumulator phases consumed
        const bool head_mma = p.stash && p.head_mma;
        float hw[3] = {0.f, 0.f, 0.f};      // d rgb.2.weight[k][row]: this thread's TMEM lane is the rgb.2 input column `row`
        float hb[3] = {0.f, 0.f, 0.f};      // d rgb.2.bias[k], this thread's rows
        const size_t mask_set = (size_t)nl * 8 * kTileRows;      // words of one tile's ReLU masks
        uint32_t* mscr = p.mask_scratch + (size_t)(blockIdx.x * 2 + g) * (1 + LA) * mask_set + row;
        const int tg = (warp & 3) * 32 + lane;           // thread index inside the group
        float* sB = sBias + g * 2 * kW;
        float4* ring = sRing + g * 2 * kTileRows;
        uint32_t bsel = 0;                               // staging buffer of the next layer (alternates)

        CNB_TR_DECL(tr_wbuf); CNB_TR_DECL(tr_wacc_f); CNB_TR_DECL(tr_epi_f); CNB_TR_DECL(tr_mid); CNB_TR_DECL(tr_mid_f); CNB_TR_DECL(tr_headw); CNB_TR_DECL(tr_wacc_b);
        CNB_TR_DECL(tr_epi_b); CNB_TR_DECL(tr_enc); CNB_TR_DECL(tr_tot);
        const long long tr_t0 = CNB_TR_NOW();
        auto wait_buf_free = [&]() { if (wp > 0) CNB_TR(tr_wbuf, umma::mbar_wait(&buf_free[g], (wp - 1) & 1u)); };
        const bool early = p.stash && p.stash_early && !p.stash_copy;
        // (stash_early) block b of the operand image being written is complete: hand it to the auxiliary warp
        auto block_done = [&](int b) {
            if (!(p.experiment & 32)) umma::fence_proxy_async_smem();      // (bit 5: timing experiment without the proxy fence)
#if CNB_EARLY_NAMED
            umma::named_bar_arrive(4 + 4 * g + b, 160);
#else
            __syncwarp();
            if (lane == 0) umma::mbar_arrive(b == 0 ? &aux_ready[g] : &aux_blk[g * 3 + b - 1]);
#endif
        };
        const bool by2 = p.stash_early == 2;      // one signal per two blocks (every phase the hooks serve has 2 or 4 blocks)
        auto hook_on = [&](int b) { if (early) { if (!by2) block_done(b); else if (b & 1) block_done(b >> 1); } };
        auto hook_last = [&](int b) { if (early && !head_mma) { if (!by2) block_done(b); else if (b & 1) block_done(b >> 1); } };      // rgb.0's output is only stashed for the head kernel
        // `nblk`: 64-column blocks of the image just written; with the early stash its LAST block is handed to the auxiliary
        // warp here (one fence and one elected-lane branch for both signals), the others by the hooks
        auto publish = [&](bool to_mma, int nblk) {
            umma::tc_fence_before();
            umma::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                if (to_mma) { if (CG == 2) umma::mbar_arrive_cluster(a_ready_addr0 + g * 8); else umma::mbar_arrive(&a_ready[g]); }
                if (!early) umma::mbar_arrive(&aux_ready[g]);
                else {
                    const int sig = by2 ? (nblk > 1 ? nblk / 2 : 1) - 1 : nblk - 1;      // index of the last signal of this image
                    umma::mbar_arrive(sig == 0 ? &aux_ready[g] : &aux_blk[g * 3 + sig - 1]);
                }
            }
            ++wp;
        };

        // rays / samples of one tile (tile < 0: none), then their positional encodings as packed bf16 rows in registers.
        // Two steps so that each fits inside ONE GEMM wait of the backward chain (together they took ~4.5 K cycles against a
        // ~2.8 K wait: the trace showed the accumulator of that step picked up 2-4 K cycles late).
        struct RaySample { float pos[3], dir[3]; bool ok; };
        auto fetch_tile = [&](int64_t tile, RaySample& rsm) {
            const int64_t lr = tile * kTileRows + row;
            rsm.ok = tile >= 0 && lr < p.S;
#pragma unroll
            for (int k = 0; k < 3; ++k) { rsm.pos[k] = 0.f; rsm.dir[k] = 0.f; }
            if (rsm.ok) {
                if (p.mode == 0) {
                    const int64_t gr = p.row_offset + lr;
                    const int64_t ray = gr / N;
                    const int zi = (int)(gr - ray * N);
                    float o[3];
                    cnb_fetch_ray(p.rs, ray, o, rsm.dir);
                    const int64_t seg = ray / p.rs.rays_per_segment;
                    const float z = __ldg(p.rs.z_vals + (p.rs.z_per_segment ? seg * N : 0) + zi);
#pragma unroll
                    for (int k = 0; k < 3; ++k) rsm.pos[k] = cnb_sample_coord(o[k], rsm.dir[k], z);
                } else {
#pragma unroll
                    for (int k = 0; k < 3; ++k) { rsm.pos[k] = __ldg(p.xyz + lr * 3 + k); rsm.dir[k] = __ldg(p.viewdir + lr * 3 + k); }
                }
            }
        };
        auto encode_tile = [&](const RaySample& rsm, PeRow& pe) {
            pe_compute_xyz(rsm.pos, rsm.ok, pe.x);
            pe_compute_dir(rsm.dir, rsm.ok, pe.d);
        };
        auto prepare_tile = [&](int64_t tile, PeRow& pe) { RaySample rsm; fetch_tile(tile, rsm); encode_tile(rsm, pe); };
        // first tile of unit slot u of this group (-1: none; phantom units encode all-zero rows)
        auto unit_tile0 = [&](int u) -> int64_t { return (u < UN && u < UN_own) ? (unit0 + u) * U : (int64_t)-1; };
        PeRow pe;
        RaySample next_rays;
#ifdef CNB_PE_ONE_STEP      // (A/B: both halves inside the same wait, as before)
        const int s_fetch = ns > 3 ? 3 : 1, s_encode = s_fetch;
#else
        const int s_fetch = ns > 3 ? 3 : 1, s_encode = ns > 5 ? 5 : s_fetch;
#endif      // backward steps whose GEMM waits hide the two halves
        bool pe_ready = false;
        if (g < UN) { prepare_tile(unit_tile0(g), pe); pe_ready = true; }

        for (int r = 0; r < rounds; ++r) {
            const int u = 2 * r + g;
            if (u >= UN) break;
            const bool phantom = u >= UN_own;
            const int64_t utile0 = (unit0 + u) * U;
            float x_even = 0.f, x_odd = 0.f;       // sigma pre-activation of this row in the unit's even / odd tiles
            for (int i = 0; i < U + LA; ++i) {
                uint32_t mlast4[4];
                if (i < U) {
                    // =================== F(i): forward chain of tile utile0 + i ===================
                    const int64_t tile = utile0 + i;
                    const long long tr_e0 = CNB_TR_NOW();
                    int64_t code = 0;
                    if (p.n_codes > 1) { code = (p.row_offset + tile * kTileRows) / p.rows_per_code; if (code >= p.n_codes) code = p.n_codes - 1; }
                    if (!pe_ready) prepare_tile(phantom ? (int64_t)-1 : tile, pe);
                    pe_ready = false;
                    // ---- phase F0: positional encodings (normally computed during the previous tile's backward chain) ----
                    wait_buf_free();
                    pe_store_xyz(pe.x, sA, row);
                    pe_store_dir(pe.d, sA + 4 * kABlock, row);
                    publish(true, 1);
                    if ((warp & 3) == 0) CNB_EV(lane, 1 + g, (6 << 12) | (g << 8));               // encodings stored: tile starts
                    tr_enc += (unsigned long long)(CNB_TR_NOW() - tr_e0);

                    uint32_t* mset = mscr + (size_t)(LA ? (i & 1) : 0) * mask_set;
                    HeadAcc hacc = {0ull, 0ull, 0ull, 0ull, pol_keep};
                    for (int l = 0; l < nl; ++l) {
                        const FwdLayer& L = p.layers[l];
                        // code-conditioned layers: the tile's folded bias row is staged in shared memory (2 floats per thread,
                        // in flight during the wait); every other layer reads its row from constant memory
                        const bool staged = L.folded >= 0 || !CNB_K2_BIAS_CONST;
                        const float* bias_g = L.folded >= 0 ? p.folded + ((size_t)code * p.n_folded + L.folded) * kW : L.bias;
                        float2 bias2 = make_float2(0.f, 0.f);
                        if (staged && 2 * tg < L.n_halves * 128) bias2 = __ldg(reinterpret_cast<const float2*>(bias_g) + tg);
                        CNB_TR(tr_wacc_f, umma::mbar_wait(&acc_full[g], opc & 1u)); ++opc;
                        if ((warp & 3) == 0) CNB_EV(lane, 1 + g, (3 << 12) | (g << 8) | l);      // accumulator visible: epilogue starts
                        const long long tr_p0 = CNB_TR_NOW();
                        umma::tc_fence_after();
                        const bool last = (l + 1 == nl);
                        const bool store = !last || p.stash;
                        if (store) wait_buf_free();
                        uint32_t* ml = mset + (size_t)l * 8 * kTileRows;
                        // head weights: constant memory, or the shared-memory copy (a per-layer token keeps the loads inside the layer)
                        const uint32_t htok = CNB_K2_HEADS_CONST ? 0u : order_token();
                        const float* wsig_c = CNB_K2_HEADS_CONST ? crow_ptr(kCrowWsig) : smem_fptr(sWsig, htok);
                        const float* wrgb_c = CNB_K2_HEADS_CONST ? crow_ptr(kCrowWrgb) : smem_fptr(sWrgb, htok);
                        if (staged) {
                            float* sb = sB + bsel * kW; bsel ^= 1u;
                            if (2 * tg < L.n_halves * 128) *reinterpret_cast<float2*>(sb + 2 * tg) = bias2;
                            const uint32_t tok = bar_sync_token(1 + g, 128);
                            const float* bias = smem_fptr(sb, tok);
                            if (L.kind == 1) fwd_epilogue_layer<8, 1, true, true, 1, kHeadSrc, kDeep>(taddr, bias, a8, wsig_c, wrgb_c, hacc, ml, hook_on);
                            else if (L.n_halves == 2) fwd_epilogue_layer<8, 0, true, true, 1, kHeadSrc, kDeep>(taddr, bias, a8, wsig_c, wrgb_c, hacc, ml, hook_on);
                            else if (p.fuse_comp) {
                                if (store) fwd_epilogue_layer<4, 2, true, true, 1, kHeadSrc>(taddr, bias, a8, wsig_c, wrgb_c, hacc, ml, hook_last);
                                else fwd_epilogue_layer<4, 2, false, true, 1, kHeadSrc>(taddr, bias, a8, wsig_c, wrgb_c, hacc, ml);
                            }
                            else if (store) fwd_epilogue_layer<4, 0, true, true, 1, kHeadSrc>(taddr, bias, a8, wsig_c, wrgb_c, hacc, ml, hook_last);
                            else fwd_epilogue_layer<4, 0, false, true, 1, kHeadSrc>(taddr, bias, a8, wsig_c, wrgb_c, hacc, ml);
                        } else {
                            // fixed constant-memory slots (crow_slot): immediate addresses, uniform loads
                            if (L.kind == 1) fwd_epilogue_layer<8, 1, true, true, 2, kHeadSrc>(taddr, crow_ptr(kCrowBias + 1 * kW), a8, wsig_c, wrgb_c, hacc, ml, hook_on);
                            else if (L.n_halves == 2) {
                                if (l == 0) fwd_epilogue_layer<8, 0, true, true, 2, kHeadSrc>(taddr, crow_ptr(kCrowBias + 0 * kW), a8, wsig_c, wrgb_c, hacc, ml, hook_on);
                                else fwd_epilogue_layer<8, 0, true, true, 2, kHeadSrc>(taddr, crow_ptr(kCrowBias + 2 * kW), a8, wsig_c, wrgb_c, hacc, ml, hook_on);
                            } else {
                                const float* bias = crow_ptr(kCrowBias + 3 * kW);
                                if (p.fuse_comp) {
                                    if (store) fwd_epilogue_layer<4, 2, true, true, 2, kHeadSrc>(taddr, bias, a8, wsig_c, wrgb_c, hacc, ml, hook_last);
                                    else fwd_epilogue_layer<4, 2, false, true, 2, kHeadSrc>(taddr, bias, a8, wsig_c, wrgb_c, hacc, ml);
                                }
                                else if (store) fwd_epilogue_layer<4, 0, true, true, 2, kHeadSrc>(taddr, bias, a8, wsig_c, wrgb_c, hacc, ml, hook_last);
                                else fwd_epilogue_layer<4, 0, false, true, 2, kHeadSrc>(taddr, bias, a8, wsig_c, wrgb_c, hacc, ml);
                            }
                        }
                        if (store && !(last && head_mma)) publish(!last, L.n_halves * 2);      // (r1 for the in-kernel rgb.2 gradient is published with the seeds)
                        if ((warp & 3) == 0) CNB_EV(lane, 1 + g, (4 << 12) | (g << 8) | l);      // epilogue done, operand published
                        tr_epi_f += (unsigned long long)(CNB_TR_NOW() - tr_p0);
                    }
                    const long long tr_m0 = CNB_TR_NOW();
                    // ReLU bit words of rgb.0 of the tile whose backward chain runs next: fetched now, used after the
                    // compositing (L2 latency overlapped)
                    if (i >= LA) {
                        const uint32_t* mb = mscr + (size_t)(LA ? ((i - LA) & 1) : 0) * mask_set;
#pragma unroll
                        for (int c = 0; c < 4; ++c) mlast4[c] = umma::ld_global_hint(mb + ((size_t)(nl - 1) * 8 + c) * kTileRows, pol_keep);
                    }
                    float sig_pre;
                    { float a0, a1; unpk2(hacc.sig2, a0, a1); sig_pre = a0 + a1; }
                    const float x = sig_pre + (CNB_K2_HEADS_CONST ? crow(kCrowBsig) : __ldg(p.b_sigma));
                    if (i & 1) x_odd = x; else x_even = x;
                    if (p.fuse_comp) {
                        // per-ray compositing, loss seed and compositing backward of every ray that ends in this tile
                        float a0, a1, cr, cg, cb;
                        unpk2(hacc.r2, a0, a1); cr = a0 + a1 + (CNB_K2_HEADS_CONST ? crow(kCrowBsig + 1) : __ldg(p.b_rgb2 + 0));
                        unpk2(hacc.g2, a0, a1); cg = a0 + a1 + (CNB_K2_HEADS_CONST ? crow(kCrowBsig + 2) : __ldg(p.b_rgb2 + 1));
                        unpk2(hacc.b2, a0, a1); cb = a0 + a1 + (CNB_K2_HEADS_CONST ? crow(kCrowBsig + 3) : __ldg(p.b_rgb2 + 2));
                        ring[((i & 1) << 7) + row] = make_float4(cnb_softplus(x), cr, cg, cb);
                        umma::named_bar_sync(1 + g, 128);
                        const int j_lo = (i * kTileRows) / N, j_hi = ((i + 1) * kTileRows) / N;      // rays of the unit ending in tile i
                        const int64_t ray_unit0 = phantom ? p.n_rays_total : (p.row_offset + utile0 * kTileRows) / N;
                        for (int j = j_lo + (warp & 3); j < j_hi; j += 4)
                            composite_fwd_bwd(p, ring, (j * N) & 255, phantom ? p.n_rays_total : ray_unit0 + j, lane);
                        umma::named_bar_sync(1 + g, 128);
                    }
                    tr_mid += (unsigned long long)(CNB_TR_NOW() - tr_m0); tr_mid_f += (unsigned long long)(CNB_TR_NOW() - tr_m0);
                } else {
                    const uint32_t* mb = mscr + (size_t)((i - LA) & 1) * mask_set;       // i == U: only with lookahead
#pragma unroll
                    for (int c = 0; c < 4; ++c) mlast4[c] = umma::ld_global_hint(mb + ((size_t)(nl - 1) * 8 + c) * kTileRows, pol_keep);
                }
                if (i >= LA) {
                    // =================== B(i - LA): input-gradient chain of tile utile0 + i - LA ===================
                    const int ib = i - LA;
                    const int64_t tile = utile0 + ib;
                    const long long tr_m0 = CNB_TR_NOW();
                    const int64_t lrow = tile * kTileRows + row;      // launch-relative row
                    const bool valid = !phantom && lrow < p.S;
                    float ds = 0.f, dcr = 0.f, dcg = 0.f, dcb = 0.f;
                    if (p.fuse_comp) {
                        const float4 sd = ring[((ib & 1) << 7) + row];
                        ds = sd.x; dcr = sd.y; dcg = sd.z; dcb = sd.w;
                    } else if (valid) {
                        ds = __ldg(p.d_sigmas + lrow);
                        dcr = __ldg(p.d_rgbs + lrow * 3 + 0); dcg = __ldg(p.d_rgbs + lrow * 3 + 1); dcb = __ldg(p.d_rgbs + lrow * 3 + 2);
                    }
                    const float x = (ib & 1) ? x_odd : x_even;
                    const float ex = expf(x);
                    const float dspre = x > 20.f ? ds : ds * ex / (ex + 1.f);      // softplus backward (ATen form)
                    if (p.stash && valid) {
                        p.dspre[lrow] = dspre;
                        if (p.fuse_comp && !head_mma) { p.drgb_out[lrow * 3] = dcr; p.drgb_out[lrow * 3 + 1] = dcg; p.drgb_out[lrow * 3 + 2] = dcb; }
                    }
                    const uint32_t* mset = mscr + (size_t)(LA ? (ib & 1) : 0) * mask_set;
                    if (head_mma) {
                        // this row's d_rgb as bf16 hi + lo -> chunk 0 of its row of the (idle) PE(viewdir) block: the B operand
                        // of the rgb.2 weight-gradient MMA; r1, its A operand, is still in blocks 0-1
                        const float s3[3] = {valid ? dcr : 0.f, valid ? dcg : 0.f, valid ? dcb : 0.f};
                        float hi[3], lo[3];
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            hi[k] = __bfloat162float(__float2bfloat16_rn(s3[k]));
                            lo[k] = s3[k] - hi[k];
                            hb[k] += s3[k];
                        }
                        st_shared_v4(sA + 4 * kABlock + row * 64 + ((0 ^ ((row >> 1) & 3)) << 4), umma::pack_bf16(hi[0], hi[1]),
                                     umma::pack_bf16(hi[2], lo[0]), umma::pack_bf16(lo[1], lo[2]), 0u);
                        umma::tc_fence_before();
                        umma::fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) { if (CG == 2) umma::mbar_arrive_cluster(a_ready_addr0 + g * 8); else umma::mbar_arrive(&a_ready[g]); }
                    }

                    // ---- step 0: gradient of the rgb.0 pre-activation = (d_rgb . W_rgb2) * relu' ----
                    if (!head_mma) wait_buf_free();      // (with the in-kernel head gradient: blocks 2-3, free since the rgb.0 GEMM)
                    {
                        const float* wrgb_s = CNB_K2_HEADS_CONST ? crow_ptr(kCrowWrgb) : smem_fptr(sWrgb, order_token());
                        const uint64_t r2 = pk2f(dcr, dcr), g2 = pk2f(dcg, dcg), b2 = pk2f(dcb, dcb);
                        // the rgb.2 rows of column group c8 + 1 are loaded before group c8 is stored (see fwd_epilogue32)
                        float4 wn[6];
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            wn[hh * 3 + 0] = ld_vec4<kHeadSrc>(wrgb_s + hh * 4);
                            wn[hh * 3 + 1] = ld_vec4<kHeadSrc>(wrgb_s + (kW / 2) + hh * 4);
                            wn[hh * 3 + 2] = ld_vec4<kHeadSrc>(wrgb_s + kW + hh * 4);
                        }
#pragma unroll
                        for (int c8 = 0; c8 < 16; ++c8) {
                            const uint32_t mlast = mlast4[c8 >> 2];
                            const int col = c8 * 8;
                            uint32_t w[4];
                            float4 wc[6];
#pragma unroll
                            for (int k = 0; k < 6; ++k) wc[k] = wn[k];
                            if (c8 + 1 < 16) {
#pragma unroll
                                for (int hh = 0; hh < 2; ++hh) {
                                    wn[hh * 3 + 0] = ld_vec4<kHeadSrc>(wrgb_s + col + 8 + hh * 4);
                                    wn[hh * 3 + 1] = ld_vec4<kHeadSrc>(wrgb_s + (kW / 2) + col + 8 + hh * 4);
                                    wn[hh * 3 + 2] = ld_vec4<kHeadSrc>(wrgb_s + kW + col + 8 + hh * 4);
                                }
                            }
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh) {
                                const float4 w0 = wc[hh * 3 + 0], w1 = wc[hh * 3 + 1], w2 = wc[hh * 3 + 2];
                                uint64_t v0 = ffma2(r2, pk2f(w0.x, w0.y), ffma2(g2, pk2f(w1.x, w1.y), ffma2(b2, pk2f(w2.x, w2.y), 0ull)));
                                uint64_t v1 = ffma2(r2, pk2f(w0.z, w0.w), ffma2(g2, pk2f(w1.z, w1.w), ffma2(b2, pk2f(w2.z, w2.w), 0ull)));
                                // columns col + 4 hh .. + 3 of the chunk: word << ((col & 31) / 4 + hh), both byte pairs
                                const uint32_t x = mlast << (((col & 31) >> 2) + hh);
                                w[hh * 2 + 0] = cvt_bf16x2<false>(v0) & relu_mask_pair(x, 0);
                                w[hh * 2 + 1] = cvt_bf16x2<false>(v1) & relu_mask_pair(x, 1);
                            }
                            if (head_mma) {
                                if (c8 < 8) st_shared_v4_off<2 * kABlock>(a8[c8 & 7], w[0], w[1], w[2], w[3]);
                                else st_shared_v4_off<3 * kABlock>(a8[c8 & 7], w[0], w[1], w[2], w[3]);
                            } else {
                                if (c8 < 8) st_shared_v4_off<0>(a8[c8 & 7], w[0], w[1], w[2], w[3]);
                                else st_shared_v4_off<kABlock>(a8[c8 & 7], w[0], w[1], w[2], w[3]);
                            }
                            if (c8 == 7) hook_on(0);
                        }
                    }
                    if (head_mma) {
                        // the small MMA has had the whole of step 0 to finish: D[m = row][0..5] = sum over the tile's rows
                        CNB_TR(tr_headw, umma::mbar_wait(&acc_full[g], opc & 1u)); ++opc;
                        umma::tc_fence_after();
                        uint32_t d8[8];
                        umma::tmem_ld8(taddr + (CG == 2 ? rank * 32u : 0u), d8);
                        umma::tmem_ld_wait();
#pragma unroll
                        for (int k = 0; k < 3; ++k) hw[k] += __uint_as_float(d8[k]) + __uint_as_float(d8[k + 3]);
                    }
                    publish(ns > 1, 2);
                    if ((warp & 3) == 0) CNB_EV(lane, 1 + g, (5 << 12) | (g << 8) | nl);         // compositing + step 0 done
                    tr_mid += (unsigned long long)(CNB_TR_NOW() - tr_m0);

                    // the tile whose forward chain this group runs next: its encodings are computed inside a long MMA wait
                    int64_t next_tile = -2;                      // -2: nothing follows
                    if (i + 1 < U) next_tile = phantom ? (int64_t)-1 : utile0 + i + 1;
                    else if (i + 1 == U + LA && u + 2 < UN) next_tile = unit_tile0(u + 2);

                    // ---- input-gradient chain ----
                    const uint64_t dsp2 = pk2f(dspre, dspre);
                    for (int s = 1; s < ns; ++s) {
                        const BwdStep& B = p.steps[s];
                        if (next_tile != -2) {
                            if (s == s_fetch) fetch_tile(next_tile, next_rays);
                            if (s == s_encode) { encode_tile(next_rays, pe); pe_ready = true; }
                        }
                        CNB_TR(tr_wacc_b, umma::mbar_wait(&acc_full[g], opc & 1u)); ++opc;
                        if ((warp & 3) == 0) CNB_EV(lane, 1 + g, (3 << 12) | (g << 8) | (nl + s - 1));
                        const long long tr_b0 = CNB_TR_NOW();
                        umma::tc_fence_after();
                        wait_buf_free();
                        const uint32_t* ml = mset + (size_t)(B.mask_layer >= 0 ? B.mask_layer : 0) * 8 * kTileRows;
                        const float* wsig_s = CNB_K2_HEADS_CONST ? crow_ptr(kCrowWsig) : smem_fptr(sWsig, order_token());
                        if (B.add_sigma) bwd_epilogue_layer<false, true>(taddr, a8, ml, dsp2, wsig_s, pol_keep, hook_on);
                        else if (B.mask_layer >= 0) bwd_epilogue_layer<true, false>(taddr, a8, ml, dsp2, wsig_s, pol_keep, hook_on);
                        else bwd_epilogue_layer<false, false>(taddr, a8, ml, dsp2, wsig_s, pol_keep, hook_on);
                        publish(s + 1 < ns, 4);
                        if ((warp & 3) == 0) CNB_EV(lane, 1 + g, (4 << 12) | (g << 8) | (nl + s - 1));
                        tr_epi_b += (unsigned long long)(CNB_TR_NOW() - tr_b0);
                    }
                }
            }
        }
        if (head_mma) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                atomicAdd(p.d_wrgb2 + k * (kW / 2) + row, hw[k]);
                const float b = warp_sum_f(hb[k]);
                if (lane == 0) atomicAdd(p.d_brgb2 + k, b);
            }
        }
        tr_tot = (unsigned long long)(CNB_TR_NOW() - tr_t0);
        if (warp == 4) {
            CNB_TR_FLUSH(5, tr_wbuf); CNB_TR_FLUSH(6, tr_wacc_f); CNB_TR_FLUSH(7, tr_epi_f); CNB_TR_FLUSH(8, tr_mid);
            CNB_TR_FLUSH(9, tr_wacc_b); CNB_TR_FLUSH(10, tr_epi_b); CNB_TR_FLUSH(11, tr_enc); CNB_TR_FLUSH(12, tr_tot);
            CNB_TR_FLUSH(13, tr_mid_f); CNB_TR_FLUSH(14, tr_headw);
        }
    }
    umma::tc_fence_before();
    if (kCluster > 1) umma::cluster_sync_all(); else __syncthreads();
    if (warp == 1) { if (CG == 2) umma::tmem_dealloc2(tmem, 512); else umma::tmem_dealloc(tmem, 512); }
}

// ===========================================================================
// K3: weight gradients from the stashed operand tiles.
struct WgProblem {
    uint32_t d_off;       // byte offset of dY_l inside a tile's dY stash
    uint32_t a_off;       // byte offset of the layer input inside a tile's A stash
    int64_t w_off;        // float offset of dW inside the flat parameter-gradient vector
    int32_t ld;           // fan_in: row stride of dW
    int32_t col0;         // first dW column this problem writes
    int32_t n_valid;      // valid columns
    uint8_t m_blocks;     // dY width / 64 (2 or 4)
    uint8_t n_blocks;     // input width / 64 (1 or 4)
    uint8_t is_dir;       // input is the [128 x 32] PE(viewdir) block (64-B swizzle)
    uint8_t layer;        // fwd layer of dY_l: slot of its column sums (bias / latent-code gradients); 0xff: none
    uint8_t sigma;        // the staged input is f (encoding_viewdir's input): the flush warps also reduce the sigma-head
                          // weight gradient sum_rows dspre[row] * f[row][:] from it (the head kernel need not re-read f)
    int32_t splits;       // work items this problem is cut into (along rows)
    int32_t item0;        // index of its first work item
};
struct WgParams {
    int n_problems, n_items;
    WgProblem prob[kMaxLayers + 1];
    const uint8_t *stashA, *stashD;
    uint32_t a_tile_bytes, d_tile_bytes;
    int64_t n_tiles;
    float* dP;
    float* colsum;                  // [n_codes][n_layers][256] += column sums of dY_l (reduced here from the staged tiles)
    int n_layers, n_codes;
    int64_t rows_per_code, row_offset;
    const float* dspre;             // [S] d(loss)/d(sigma pre-activation), launch-relative rows
    int64_t S;
    float *d_wsigma, *d_bsigma;     // sigma.0.weight / bias gradients (accumulated)
};
constexpr int kWgStage = 65536;      // 64 rows: dY half-blocks [0, 32 KB) + input half-blocks [32 KB, 64 KB)
constexpr int kWgStages = 3;
constexpr int kWgThreads = 192;      // warp 0 producer, warp 1 MMA, warps 2-5 flush

__device__ __forceinline__ void wg_item(const WgParams& p, int item, int& pi, int64_t& h_begin, int64_t& h_end) {
    pi = 0;
    while (pi + 1 < p.n_problems && item >= p.prob[pi + 1].item0) ++pi;
    const int s = item - p.prob[pi].item0;
    const int64_t H = p.n_tiles * 2;
    const int64_t sp = p.prob[pi].splits;
    h_begin = H * s / sp; h_end = H * (s + 1) / sp;
}

__global__ void __launch_bounds__(kWgThreads, 1) k_wgrad(const __grid_constant__ WgParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(smem + kWgStages * kWgStage);
    uint64_t* full = bars;
    uint64_t* empty = bars + kWgStages;
    uint64_t* acc_done = bars + 2 * kWgStages;
    uint64_t* acc_free = acc_done + 1;
    uint32_t* tmem_slot = (uint32_t*)(acc_free + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kWgStages; ++i) { umma::mbar_init(&full[i], 1); umma::mbar_init(&empty[i], 1 + 4); }   // MMA commit + 4 reducer warps
        umma::mbar_init(acc_done, 1); umma::mbar_init(acc_free, 4);
        umma::fence_mbar_init();
    }
    if (warp == 1) umma::tmem_alloc(tmem_slot, 512);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ===== producer: half-tile operand slices of the stash, TMA bulk copies =====
        int stage = 0; uint32_t ph = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            int pi; int64_t hb, he;
            wg_item(p, item, pi, hb, he);
            const WgProblem& P = p.prob[pi];
            const uint32_t bytes = P.m_blocks * 8192u + (P.is_dir ? 4096u : P.n_blocks * 8192u);
            for (int64_t h = hb; h < he; ++h) {
                const int64_t tile = h >> 1; const uint32_t half = (uint32_t)(h & 1);
                umma::mbar_wait(&empty[stage], ph ^ 1);
                if (umma::elect_one()) {
                    umma::mbar_arrive_expect_tx(&full[stage], bytes);
                    uint8_t* dst = smem + stage * kWgStage;
                    const uint8_t* srcD = p.stashD + (size_t)tile * p.d_tile_bytes + P.d_off + half * 8192u;
                    for (int b = 0; b < P.m_blocks; ++b) umma::bulk_g2s(dst + b * 8192, srcD + (size_t)b * kABlock, 8192, &full[stage]);
                    if (P.is_dir) {
                        umma::bulk_g2s(dst + 32768, p.stashA + (size_t)tile * p.a_tile_bytes + P.a_off + half * 4096u, 4096, &full[stage]);
                    } else {
                        const uint8_t* srcA = p.stashA + (size_t)tile * p.a_tile_bytes + P.a_off + half * 8192u;
                        for (int b = 0; b < P.n_blocks; ++b) umma::bulk_g2s(dst + 32768 + b * 8192, srcA + (size_t)b * kABlock, 8192, &full[stage]);
                    }
                }
                __syncwarp();
                if (++stage == kWgStages) { stage = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (whole warp runs the loop, one elected lane issues) =====
        int stage = 0; uint32_t ph = 0; uint32_t it = 0;
        const uint64_t dA0 = umma::make_sdesc(umma::smem_u32(smem), 8192, 1024, umma::SWZ_128B);
        const uint64_t dB0 = umma::make_sdesc(umma::smem_u32(smem) + 32768, 8192, 1024, umma::SWZ_128B);
        const uint64_t dD0 = umma::make_sdesc(umma::smem_u32(smem) + 32768, 4096, 512, umma::SWZ_64B);
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
            int pi; int64_t hb, he;
            wg_item(p, item, pi, hb, he);
            const WgProblem& P = p.prob[pi];
            const uint32_t idesc = umma::make_idesc(128, P.is_dir ? 32 : P.n_blocks * 64, 1, 1);
            const int mhalves = P.m_blocks / 2;
            if (it > 0) { umma::mbar_wait(acc_free, (it - 1) & 1u); umma::tc_fence_after(); }
            for (int64_t h = hb; h < he; ++h) {
                umma::mbar_wait(&full[stage], ph);
                umma::tc_fence_after();
                if (umma::elect_one()) {
                    const uint64_t so = (uint64_t)((stage * kWgStage) >> 4);
                    const uint32_t first = (h > hb) ? 1u : 0u;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t db = P.is_dir ? dD0 + so + (uint64_t)(ks * 64) : dB0 + so + (uint64_t)(ks * 128);
                        for (int mh = 0; mh < mhalves; ++mh)
                            umma::mma_bf16(tmem + mh * 256, dA0 + so + (uint64_t)(mh * 1024 + ks * 128), db, idesc, first | (ks > 0 ? 1u : 0u));
                    }
                    umma::mma_commit(&empty[stage]);
                }
                __syncwarp();
                if (++stage == kWgStages) { stage = 0; ph ^= 1; }
            }
            if (umma::elect_one()) umma::mma_commit(acc_done);
            __syncwarp();
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int wq = warp - 2;                 // reducer index 0..3: owns dY half-block wq of every stage
        const int chunk = lane & 7, rg = lane >> 3;
        uint32_t it = 0;
        int stage = 0; uint32_t ph = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
            int pi; int64_t hb, he;
            wg_item(p, item, pi, hb, he);
            const WgProblem& P = p.prob[pi];
            // ---- column sums of dY_l from the staged half-tiles (bias and latent-code gradients) ----
            const bool reduce = P.layer != 0xff && wq < P.m_blocks;
            const bool sigma = P.sigma != 0;           // input blocks are f: 4 blocks, one per flush warp
            float cs[8], sg[8], sgb = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) { cs[i] = 0.f; sg[i] = 0.f; }
            int64_t cur_code = -1;
            auto flush_cs = [&]() {
                if (cur_code < 0) return;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    cs[i] += __shfl_xor_sync(0xffffffffu, cs[i], 8);
                    cs[i] += __shfl_xor_sync(0xffffffffu, cs[i], 16);
                }
                if (rg == 0) {
                    float* out = p.colsum + ((size_t)cur_code * p.n_layers + P.layer) * kW + wq * 64 + chunk * 8;
#pragma unroll
                    for (int i = 0; i < 8; ++i) atomicAdd(out + i, cs[i]);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) cs[i] = 0.f;
            };
            for (int64_t h = hb; h < he; ++h) {
                umma::mbar_wait(&full[stage], ph);
                if (reduce || sigma) {
                    int64_t code = 0;
                    if (p.n_codes > 1) { code = (p.row_offset + (h >> 1) * kTileRows) / p.rows_per_code; if (code >= p.n_codes) code = p.n_codes - 1; }
                    if (reduce && code != cur_code) { flush_cs(); cur_code = code; }
                    const uint8_t* base = smem + stage * kWgStage + wq * 8192;
                    uint4 w[16], wf[16];          // load first, release the stage, then add
                    float dsp[16];
                    if (reduce) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int rr = rg * 16 + i;
                            w[i] = ld_shared_v4(base + rr * 128 + ((chunk ^ (rr & 7)) << 4));
                        }
                    }
                    if (sigma) {
                        const int64_t r0 = h * 64 + rg * 16;       // launch-relative row of this lane's first row in the half-tile
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int rr = rg * 16 + i;
                            wf[i] = ld_shared_v4(base + 32768 + rr * 128 + ((chunk ^ (rr & 7)) << 4));
                            dsp[i] = (r0 + i < p.S) ? __ldg(p.dspre + r0 + i) : 0.f;
                        }
                    }
                    __syncwarp();
                    if (lane == 0) umma::mbar_arrive(&empty[stage]);
                    if (reduce) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            cs[0] += bf_lo(w[i].x); cs[1] += bf_hi(w[i].x); cs[2] += bf_lo(w[i].y); cs[3] += bf_hi(w[i].y);
                            cs[4] += bf_lo(w[i].z); cs[5] += bf_hi(w[i].z); cs[6] += bf_lo(w[i].w); cs[7] += bf_hi(w[i].w);
                        }
                    }
                    if (sigma) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            sg[0] = fmaf(dsp[i], bf_lo(wf[i].x), sg[0]); sg[1] = fmaf(dsp[i], bf_hi(wf[i].x), sg[1]);
                            sg[2] = fmaf(dsp[i], bf_lo(wf[i].y), sg[2]); sg[3] = fmaf(dsp[i], bf_hi(wf[i].y), sg[3]);
                            sg[4] = fmaf(dsp[i], bf_lo(wf[i].z), sg[4]); sg[5] = fmaf(dsp[i], bf_hi(wf[i].z), sg[5]);
                            sg[6] = fmaf(dsp[i], bf_lo(wf[i].w), sg[6]); sg[7] = fmaf(dsp[i], bf_hi(wf[i].w), sg[7]);
                            if (wq == 0 && chunk == 0) sgb += dsp[i];
                        }
                    }
                } else {
                    __syncwarp();
                    if (lane == 0) umma::mbar_arrive(&empty[stage]);
                }
                if (++stage == kWgStages) { stage = 0; ph ^= 1; }
            }
            if (sigma) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    sg[i] += __shfl_xor_sync(0xffffffffu, sg[i], 8);
                    sg[i] += __shfl_xor_sync(0xffffffffu, sg[i], 16);
                }
                sgb += __shfl_xor_sync(0xffffffffu, sgb, 8);
                sgb += __shfl_xor_sync(0xffffffffu, sgb, 16);
                if (rg == 0) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) atomicAdd(p.d_wsigma + wq * 64 + chunk * 8 + i, sg[i]);
                    if (wq == 0 && chunk == 0) atomicAdd(p.d_bsigma, sgb);
                }
            }
            if (reduce) flush_cs();
            umma::mbar_wait(acc_done, it & 1u);
            umma::tc_fence_after();
            const int ncols = P.is_dir ? 32 : P.n_blocks * 64;
            for (int mh = 0; mh < P.m_blocks / 2; ++mh) {
                float* out = p.dP + P.w_off + (size_t)(mh * 128 + row) * P.ld + P.col0;
                for (int cc = 0; cc < ncols; cc += 32) {
                    uint32_t rr[32];
                    umma::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + mh * 256 + cc, rr);
                    umma::tmem_ld_wait();
                    if (he > hb) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (cc + j < P.n_valid) atomicAdd(out + cc + j, __uint_as_float(rr[j]));
                    }
                }
            }
            umma::tc_fence_before();
            __syncwarp();
            if (lane == 0) umma::mbar_arrive(acc_free);
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 1) umma::tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------
// rgb.2 weight / bias gradients from the stash: d(rgb.2.weight) = sum d_rgb (x) r1.  (The sigma head's gradient is
// reduced inside K3, whose encoding_viewdir problem stages the sigma head's input anyway.)  A pure HBM stream
// (32 KB per tile): every warp takes 16 KB block images of r1 (the rgb.2 input, 128 wide: two blocks per tile) and
// reads them with fully coalesced 512-byte loads, eight in flight.  Load i of a block gives lane l the 16-byte chunk
// at position l & 7 of row 4 i + (l >> 3); with the 128-byte swizzle that is logical chunk (l & 7) ^ (row & 7), which
// for a fixed lane alternates between two values with the parity of i: two accumulator sets per lane, reduced across
// CTAs with atomics at the end.
constexpr int kHeadWarps = 8;
__global__ void __launch_bounds__(kHeadWarps * 32) k_head_wgrad(const uint8_t* __restrict__ stashA, uint32_t a_tile_bytes, uint32_t r1_off,
                             const float* __restrict__ d_rgbs, int64_t S, int64_t n_tiles,
                             float* __restrict__ d_wrgb2, float* __restrict__ d_brgb2) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int blk = warp & 1;                   // which of the tile's two r1 blocks this warp reads
    const int rsub = lane >> 3, pos = lane & 7;
    float acc[2][3][8];            // [row parity class][r, g, b][8 columns of the chunk]
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[a][k][j] = 0.f;
    float bsum[3] = {0.f, 0.f, 0.f};            // bias gradient: block-0 warps, lanes with pos == 0
    const int64_t tstride = (int64_t)gridDim.x * (kHeadWarps / 2);
    for (int64_t tile = (int64_t)blockIdx.x * (kHeadWarps / 2) + (warp >> 1); tile < n_tiles; tile += tstride) {
        const uint4* src = reinterpret_cast<const uint4*>(stashA + (size_t)tile * a_tile_bytes + r1_off + (size_t)blk * kABlock) + lane;
        const int64_t row0 = tile * kTileRows;
#pragma unroll
        for (int i0 = 0; i0 < 32; i0 += 8) {
            uint4 w[8]; float s0[8], s1[8], s2[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int64_t r = row0 + 4 * (i0 + u) + rsub;
                const bool ok = r < S;
                w[u] = __ldcs(src + (size_t)(i0 + u) * 32);          // streaming: every byte is read once
                s0[u] = ok ? __ldg(d_rgbs + r * 3) : 0.f; s1[u] = ok ? __ldg(d_rgbs + r * 3 + 1) : 0.f; s2[u] = ok ? __ldg(d_rgbs + r * 3 + 2) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float h[8] = {bf_lo(w[u].x), bf_hi(w[u].x), bf_lo(w[u].y), bf_hi(w[u].y),
                                    bf_lo(w[u].z), bf_hi(w[u].z), bf_lo(w[u].w), bf_hi(w[u].w)};
                const int a = u & 1;                      // parity of the load index = parity class of the row
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    acc[a][0][j] = fmaf(s0[u], h[j], acc[a][0][j]); acc[a][1][j] = fmaf(s1[u], h[j], acc[a][1][j]);
                    acc[a][2][j] = fmaf(s2[u], h[j], acc[a][2][j]);
                }
                if (pos == 0 && blk == 0) { bsum[0] += s0[u]; bsum[1] += s1[u]; bsum[2] += s2[u]; }
            }
        }
    }
    // CTA-level reduction in shared memory first: same-address atomics are served one at a time by the L2 (straight global
    // atomics from every lane, 3.6 M on 384 addresses, cost 1 ms); then 387 global atomics per CTA
    __shared__ float red[3 * (kW / 2) + 4];
    for (int i = threadIdx.x; i < 3 * (kW / 2) + 4; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    // logical chunk of parity class a: pos ^ ((4 a + rsub) & 7)
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const int chunk = pos ^ ((4 * a + rsub) & 7);
        const int col = blk * 64 + chunk * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            atomicAdd(red + 0 * (kW / 2) + col + j, acc[a][0][j]);
            atomicAdd(red + 1 * (kW / 2) + col + j, acc[a][1][j]);
            atomicAdd(red + 2 * (kW / 2) + col + j, acc[a][2][j]);
        }
    }
    if (pos == 0 && blk == 0) { atomicAdd(red + 3 * (kW / 2), bsum[0]); atomicAdd(red + 3 * (kW / 2) + 1, bsum[1]); atomicAdd(red + 3 * (kW / 2) + 2, bsum[2]); }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * (kW / 2); i += blockDim.x) atomicAdd(d_wrgb2 + i, red[i]);
    if (threadIdx.x < 3) atomicAdd(d_brgb2 + threadIdx.x, red[3 * (kW / 2) + threadIdx.x]);
}

// Bias gradients of every layer in one launch: db_l[n] += sum_codes colsum[code][l][n]  (blockIdx.y = layer).
struct BiasOut { float* db[kMaxLayers]; int n_out[kMaxLayers]; };
__global__ void k_bias_from_colsum(const float* __restrict__ colsum, int n_codes, int nl, const BiasOut o) {
    const int l = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= o.n_out[l]) return;
    float a = 0.f;
    for (int c = 0; c < n_codes; ++c) a += colsum[((size_t)c * nl + l) * kW + n];
    o.db[l][n] += a;
}

// Folded layers: dz[code][k] = sum_n cs[code][n] W[n][k]  (gradient reaching the latent branch output; blocks
// [0, n_codes)), and the rank-1 weight term the fold moved out of the GEMM, summed over the codes without atomics:
// dW[n][k] += sum_code cs[code][n] * z[code][k]  (blocks [n_codes, n_codes + 256): one dW row each; nothing else
// writes dW at that point of the stream).
__global__ void k_fold_bwd(const float* __restrict__ Wj, const float* __restrict__ cs /*[n_codes] stride cs_ld*/,
                           int64_t cs_ld, const float* __restrict__ z, int64_t z_ld, int n_codes, float* __restrict__ dz,
                           float* __restrict__ dW) {
    const int k = threadIdx.x;      // 256 threads
    if ((int)blockIdx.x < n_codes) {
        const int code = blockIdx.x;
        const float* c = cs + (size_t)code * cs_ld;
        float a = 0.f;
#pragma unroll 8
        for (int n = 0; n < kW; ++n) a = fmaf(c[n], __ldg(Wj + (size_t)n * kW + k), a);
        dz[(size_t)code * z_ld + k] = a;
    } else if (dW) {
        const int n = blockIdx.x - n_codes;
        float a = 0.f;
        for (int code = 0; code < n_codes; ++code) a = fmaf(cs[(size_t)code * cs_ld + n], z[(size_t)code * z_ld + k], a);
        dW[(size_t)n * kW + k] += a;
    }
}

// d_rgb seed of the fused training step (mean L2 over each segment, src/trainer.py:75).
__global__ void k_l2_seed_sm100(const float* __restrict__ rgb, const float* __restrict__ target, int64_t n_rays,
                                int64_t ray0, int rays_per_segment, float scale, float* __restrict__ d_rgb,
                                float* __restrict__ sq_err) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    float e2 = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float e = rgb[(ray0 + r) * 3 + k] - target[(ray0 + r) * 3 + k];
        d_rgb[r * 3 + k] = 2.f * e / (3.f * (float)rays_per_segment) * scale;
        e2 += e * e;
    }
    if (sq_err) atomicAdd(sq_err + (ray0 + r) / rays_per_segment, e2);
}

// ===========================================================================
// Host side

// rows per sub-batch = 128 x this: 8192 tiles = 1 Mi rows (training stash ~ 8 GB).  Larger sub-batches are faster
// (fewer split-K flushes of the weight gradient: 16.3 -> 15.65 ms per 32768-tile step at 32768) and need ~1 MB of
// workspace per tile; the Python layer raises the option `sub_tiles` when the device has the memory.
static int64_t max_sub_tiles() { const int64_t v = cnb_option("sub_tiles", 8192); return v >= 256 ? v : 8192; }
#define kMaxSubTiles (max_sub_tiles())

struct StashLayout {
    uint32_t a_slot[kMaxLayers + 1], dir_slot, d_slot[kMaxLayers], a_tile_bytes, d_tile_bytes;
};
StashLayout make_stash_layout(const Plan& pl) {
    StashLayout s = {};
    uint32_t off = 0;
    s.a_slot[0] = off; off += kABlock;                                   // PE(xyz)
    for (int l = 1; l < pl.n_layers; ++l) { s.a_slot[l] = off; off += 4 * kABlock; }
    s.a_slot[pl.n_layers] = off; off += 2 * kABlock;                     // rgb.2 input (128 wide)
    s.dir_slot = off; off += kABlock;                                    // PE(viewdir), 8 KB used
    s.a_tile_bytes = off;
    off = 0;
    for (int l = 0; l < pl.n_layers; ++l) { s.d_slot[l] = off; off += (uint32_t)pl.fwd[l].n_halves * 2 * kABlock; }
    s.d_tile_bytes = off;
    return s;
}

struct BwdWorkspace {
    FwdWorkspace fw;
    float *colsum, *dz;
    uint32_t* masks;
    float *spill_sig, *spill_rgb, *dsig, *drgb, *dspre, *ray_drgb, *ray_rgb, *ray_depth, *ray_acc;
    uint8_t *stashA, *stashD;
    // second copy of everything K2 hands to K3 (stash, dspre, per-sample d rgb): with more than one sub-batch K3 of
    // sub-batch i runs on a second stream while K2 of sub-batch i + 1 fills the other copy
    uint8_t *stashA2, *stashD2;
    float *dspre2, *drgb2;
    int nbuf;
    size_t bytes;
};

size_t carve_bwd(const cnb_net_config* c, const Plan& pl, int n_codes, int64_t sub_rows, int64_t sub_rays, int fused,
                 int stash, int grid, char* base, BwdWorkspace* out, int nbuf = 1) {
    size_t off = 0;
    auto take = [&](size_t bytes) -> char* { char* p = base ? base + off : nullptr; off += (bytes + 1023) & ~(size_t)1023; return p; };
    BwdWorkspace w = {};
    const int nf = c->shape_blocks + c->texture_blocks, nl = pl.n_layers;
    w.fw.z = (float*)take(sizeof(float) * (size_t)n_codes * nf * kW);
    w.fw.folded = (float*)take(sizeof(float) * (size_t)n_codes * nf * kW);
    w.fw.rows = (float*)take(sizeof(ConstRows));
    w.colsum = (float*)take(sizeof(float) * (size_t)n_codes * nl * kW);
    w.dz = (float*)take(sizeof(float) * (size_t)n_codes * nf * kW);
    w.masks = (uint32_t*)take(sizeof(uint32_t) * (size_t)grid * 2 * 2 * nl * 8 * kTileRows);   // [grid][2 groups][2 tile parities]
    w.dspre = (float*)take(sizeof(float) * (size_t)sub_rows);
    if (fused) {
        w.spill_sig = (float*)take(sizeof(float) * (size_t)sub_rows);
        w.spill_rgb = (float*)take(sizeof(float) * (size_t)sub_rows * 3);
        w.dsig = (float*)take(sizeof(float) * (size_t)sub_rows);
        w.drgb = (float*)take(sizeof(float) * (size_t)sub_rows * 3);
        w.ray_drgb = (float*)take(sizeof(float) * (size_t)sub_rays * 3);
        w.ray_rgb = (float*)take(sizeof(float) * (size_t)sub_rays * 3);
        w.ray_depth = (float*)take(sizeof(float) * (size_t)sub_rays);
        w.ray_acc = (float*)take(sizeof(float) * (size_t)sub_rays);
    }
    if (stash) {
        const StashLayout sl = make_stash_layout(pl);
        const int64_t tiles = (sub_rows + kTileRows - 1) / kTileRows;
        w.stashA = (uint8_t*)take((size_t)tiles * sl.a_tile_bytes);
        w.stashD = (uint8_t*)take((size_t)tiles * sl.d_tile_bytes);
        if (nbuf > 1) {
            w.stashA2 = (uint8_t*)take((size_t)tiles * sl.a_tile_bytes);
            w.stashD2 = (uint8_t*)take((size_t)tiles * sl.d_tile_bytes);
            w.dspre2 = (float*)take(sizeof(float) * (size_t)sub_rows);
            w.drgb2 = (float*)take(sizeof(float) * (size_t)sub_rows * 3);
        }
    }
    w.nbuf = nbuf;
    w.bytes = off;
    if (out) *out = w;
    return off;
}

int num_sms() {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

// K2 -> K3 software pipeline across sub-batches.  K2 is bound by on-chip latencies and K3 by HBM, so they overlap
// well: K3 (+ the head-gradient kernel) of sub-batch i runs on a helper stream on `k3_sms` SMs while K2 of
// sub-batch i + 1 runs on the remaining SMs (neither kernel can share an SM with the other: shared memory).
struct Pipeline {
    bool on = false;
    cudaStream_t st3 = nullptr;
    cudaEvent_t k2_done[2] = {nullptr, nullptr}, k3_done[2] = {nullptr, nullptr};
    bool pending[2] = {false, false};      // a K3 of this buffer is in flight on st3
    int k3_sms = 0;
};
struct PipelineResources { cudaStream_t st3; cudaEvent_t ev[4]; bool ok; };
PipelineResources* pipeline_resources() {
    static thread_local PipelineResources res[16] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    PipelineResources& r = res[dev];
    if (!r.ok) {
        if (cudaStreamCreateWithFlags(&r.st3, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        for (int i = 0; i < 4; ++i)
            if (cudaEventCreateWithFlags(&r.ev[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        r.ok = true;
    }
    return &r;
}
bool pipeline_wanted() {
    return cnb_option("k3_overlap", 0) != 0;      // measured: no gain
}
int pipeline_begin(Pipeline& pp, int64_t n_sub, bool training) {
    const bool want = pipeline_wanted();
    const int k3_sms = (int)cnb_option("k3_sms", 40);
    pp = Pipeline();
    if (!want || !training || n_sub < 2) return CNB_OK;
    PipelineResources* r = pipeline_resources();
    if (!r) return CNB_E_DEVICE;
    pp.on = true; pp.st3 = r->st3;
    pp.k2_done[0] = r->ev[0]; pp.k2_done[1] = r->ev[1]; pp.k3_done[0] = r->ev[2]; pp.k3_done[1] = r->ev[3];
    pp.k3_sms = k3_sms;
    return CNB_OK;
}
// join: everything queued on the helper stream is ordered before what follows on `st`
int pipeline_end(Pipeline& pp, cudaStream_t st) {
    if (!pp.on) return CNB_OK;
    for (int b = 0; b < 2; ++b)
        if (pp.pending[b]) { CNB_CUDA_TRY(cudaStreamWaitEvent(st, pp.k3_done[b], 0)); pp.pending[b] = false; }
    return CNB_OK;
}

// K2 (+ K3 and head gradients when d_params != null) over launch-relative rows [0, S).
struct FuseArgs {     // compositing (+ loss) fused into K2: per-ray arrays are indexed by global ray
    int kind;         // 1: seeds from d_rgb / d_depth, 2: L2 loss against target
    int unit_tiles;   // tiles per unit of whole rays: N / gcd(N, 128)
    int white_bg; int64_t n_rays_total;
    const float *d_rgb, *d_depth, *target; float loss_scale;
    float *rgb, *depth, *acc, *sq_err;
};

int run_mlp_bwd(const cnb_net_config* c, const float* const* P, const void* packed, const Plan& pl, BwdWorkspace& w,
                int mode, const CnbRaySource* rs, const float* xyz, const float* viewdir, int64_t S, int64_t row_offset,
                int n_codes, int64_t rows_per_code, const float* d_sigmas, const float* d_rgbs, float* d_params,
                cudaStream_t st, const FuseArgs* fuse = nullptr, Pipeline* pp = nullptr, int sub_index = 0,
                bool last_sub = true) {
    // which copy of the K2 -> K3 hand-over buffers this sub-batch uses
    const int buf = (pp && pp->on && w.nbuf > 1) ? (sub_index & 1) : 0;
    uint8_t* stashA = buf ? w.stashA2 : w.stashA;
    uint8_t* stashD = buf ? w.stashD2 : w.stashD;
    float* dspre_buf = buf ? w.dspre2 : w.dspre;
    float* drgb_buf = buf ? w.drgb2 : w.drgb;
    const bool piped = pp && pp->on && w.nbuf > 1;
    CnbLayout L; cnb_make_layout(c, &L);
    const StashLayout sl = make_stash_layout(pl);
    BwdParams bp = {};
    bp.n_layers = pl.n_layers;
    for (int i = 0; i < pl.n_layers; ++i) bp.layers[i] = pl.fwd[i];
    // step s >= 1 back-propagates through fwd layer nl - s and produces dY of layer nl - s - 1
    const int nl = pl.n_layers;
    bp.n_steps = nl;
    bp.steps[0] = BwdStep{0, 0, (int8_t)(nl - 1), 0, (uint8_t)(nl - 1), 2, 0};
    for (int s = 1; s < nl; ++s) {
        const int lay = nl - s, out = lay - 1;
        BwdStep b = {};
        b.w_off = pl.bwd_w_off[lay];
        b.n_kchunks = (uint8_t)(pl.fwd[lay].n_halves * 2);
        b.mask_layer = pl.fwd[out].relu ? (int8_t)out : (int8_t)-1;
        b.add_sigma = pl.fwd[out].kind == 1 ? 1 : 0;
        b.out_layer = (uint8_t)out; b.out_blocks = 4;
        bp.steps[s] = b;
    }
    bp.packed = (const uint8_t*)packed; bp.folded = w.fw.folded; bp.n_folded = pl.n_folded; bp.n_codes = n_codes;
    bp.rows_per_code = rows_per_code;
    bp.w_sigma = P[L.i_sigma]; bp.b_sigma = P[L.i_sigma + 1]; bp.w_rgb2 = P[L.i_rgb2];
    bp.mode = mode; if (rs) bp.rs = *rs; else bp.rs.N = 1;
    bp.xyz = xyz; bp.viewdir = viewdir; bp.S = S; bp.row_offset = row_offset;
    bp.d_sigmas = d_sigmas; bp.d_rgbs = d_rgbs;
    bp.mask_scratch = w.masks; bp.colsum = w.colsum;
    bp.stash_wrap = cnb_option("stash_wrap", 0);
    bp.experiment = (int)cnb_option("experiment", 0);
    bp.head_mma = 0;      // set below, once the unit shape is known
    bp.d_wrgb2 = d_params ? d_params + L.rgb2_w : nullptr; bp.d_brgb2 = d_params ? d_params + L.rgb2_b : nullptr;
    bp.stash_lanes = 0;      // decided below: 2 on CTA pairs, 32 otherwise
    bp.keep_weights = cnb_option("keep_weights", 1) != 0 ? 1 : 0;
    { const int64_t se = cnb_option("stash_early", 1); bp.stash_early = se == 2 ? 2 : (se != 0 ? 1 : 0); }
    { const int64_t ep = cnb_option("early_pieces", 1); bp.early_pieces = (ep == 2 || ep == 4 || ep == 8) ? (int)ep : 1; }
    bp.stash_copy = cnb_option("stash_copy", 0) != 0 ? 1 : 0;      // measured slower (K2 9.9 vs 8.6 ms): one warp's LSU rate
    bp.stash = d_params ? 1 : 0; bp.stashA = stashA; bp.stashD = stashD; bp.dspre = dspre_buf;
    // column sums of dY: with a weight-gradient pass K3 reduces them from the stash for free; otherwise the aux
    // warps of K2 do it, and only for the folded layers (the latent-code gradients need nothing else)
    bp.share_fills = cnb_option("share_fills", 1) != 0 ? 1 : 0;
    bp.colsum_layers = 0u;
    if (!d_params)
        for (int l = 0; l < nl; ++l) if (pl.fwd[l].folded >= 0) bp.colsum_layers |= 1u << l;
    bp.b_rgb2 = P[L.i_rgb2 + 1];
    bp.unit_tiles = 1; bp.lookahead = 0;
    if (fuse) {
        bp.unit_tiles = fuse->unit_tiles; bp.lookahead = fuse->unit_tiles > 1 ? 1 : 0;
        bp.fuse_comp = fuse->kind; bp.white_bg = fuse->white_bg; bp.n_rays_total = fuse->n_rays_total;
        bp.d_rgb_rays = fuse->d_rgb; bp.d_depth_rays = fuse->d_depth; bp.target = fuse->target; bp.loss_scale = fuse->loss_scale;
        bp.out_rgb = fuse->rgb; bp.out_depth = fuse->depth; bp.out_acc = fuse->acc; bp.sq_err = fuse->sq_err;
        bp.drgb_out = drgb_buf;
        d_rgbs = drgb_buf;        // the head weight gradient reads the per-sample seeds K2 writes
    }
    // the in-kernel rgb.2 gradient needs the tile's r1 still in the operand buffer when its backward chain starts: not with
    // the one-tile lookahead of straddling rays (the next tile's forward chain has overwritten it) -- those use the head kernel
    bp.head_mma = (d_params && bp.lookahead == 0 && cnb_option("head_mma", 1) != 0) ? 1 : 0;
    for (int l = 0; l <= nl; ++l) bp.a_slot[l] = sl.a_slot[l];
    for (int l = 0; l < nl; ++l) bp.d_slot[l] = sl.d_slot[l];
    bp.dir_slot = sl.dir_slot; bp.a_tile_bytes = sl.a_tile_bytes; bp.d_tile_bytes = sl.d_tile_bytes;

    const int sms = num_sms();
    const int64_t tiles = (S + kTileRows - 1) / kTileRows;
    const int64_t units = ((tiles + bp.unit_tiles - 1) / bp.unit_tiles + 1) / 2;      // pairs of units: one CTA runs two at a time
    int grid = (int)(units < sms ? (units < 1 ? 1 : units) : sms);
    // pipelined: leave SMs to the K3 of the previous sub-batch (the first K2 has the GPU to itself)
    if (piped && sub_index > 0 && grid > sms - pp->k3_sms && sms > pp->k3_sms) grid = sms - pp->k3_sms;
    if (piped && pp->pending[buf]) {      // the K3 that read this copy two sub-batches ago must be done
        CNB_CUDA_TRY(cudaStreamWaitEvent(st, pp->k3_done[buf], 0));
        pp->pending[buf] = false;
    }
    const size_t smem = 1024 + 2 * (size_t)kATile + (size_t)kNumStages * kSlot + 256 + 4 * kTileRows * sizeof(float4) +
                        sizeof(float) * (4 * kW + (CNB_K2_HEADS_CONST ? 0 : kW + 3 * (kW / 2)));
    // CNB_WEIGHT_MCAST=2: clusters of 2 share one multicast weight stream (+2.5 % in K2 in a back-to-back run, within
    // the box-to-box noise of the bench: left opt-in)
    // CTA pairs by default (round 2): every CTA streams half of each weight chunk, so the 4-slot ring covers a whole
    // layer and the MMA phase is no longer refill-bound (K2 9.35 -> 8.45 ms per 65,536-ray step, latent fit +5 %).
    // Needed two fixes to pay off: remote arrives without cluster-scope release (umma.cuh) and the stash written
    // as a few large bulk stores (below).  bwd_pairs = 0 selects the single-CTA kernel.
    const int use_pairs = (int)cnb_option("bwd_pairs", 1);
    const bool pairs = use_pairs && grid == sms;
    const int mc = pairs ? 2 : (grid == sms ? weight_multicast() : 1);      // cluster size
    if (pairs) CNB_TRY(make_weight_maps(packed, pl.total_bytes, &bp.maps));
    {   // Bulk stores per stashed operand image.  Single CTA: 32 pieces of 2 KB, so that the weight refills sharing the SM's
        // copy engine slip in between (one 64 KB store ahead of a refill stalls the ring: 10.9 -> 9.35 ms in round 1).
        // CTA pairs: the ring is not the limit any more and the epilogue waits for the store instead; two 32 KB
        // pieces drain the buffer fastest (measured: 32 pieces 10.4 ms, 8: 8.57, 2: 8.44, 1: 8.49).
        const int64_t sl = cnb_option("stash_lanes", pairs ? 2 : 32);
        bp.stash_lanes = (sl == 1 || sl == 2 || sl == 4 || sl == 8 || sl == 16) ? (int)sl : 32;
    }
    void (*kern)(const BwdParams) = pairs ? k_mlp_bwd<1, 2> : mc == 4 ? k_mlp_bwd<4> : mc == 2 ? k_mlp_bwd<2> : k_mlp_bwd<1>;
    CNB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kBwdThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = mc; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = mc > 1 ? 1 : 0;
    if (mc > 1) CNB_TRY(cluster_grid(kern, &cfg, mc, &grid));
    cfg.gridDim = dim3(grid);
    cnb_prof_begin(CNB_K_BWD, st);
    const cudaError_t le = cudaLaunchKernelEx(&cfg, kern, bp);
    cnb_prof_end(CNB_K_BWD, st);
    if (le != cudaSuccess) return (int)le;
    CNB_LAUNCH_CHECK();
    if (!d_params) return CNB_OK;

    // ---- K3: weight gradients (on the helper stream when pipelined) ----
    cudaStream_t st2 = st;
    if (piped) {
        CNB_CUDA_TRY(cudaEventRecord(pp->k2_done[buf], st));
        CNB_CUDA_TRY(cudaStreamWaitEvent(pp->st3, pp->k2_done[buf], 0));
        st2 = pp->st3;
    }
    WgParams wp = {};
    int np = 0;
    auto add = [&](int lay, int64_t w_off, int ld, int col0, int n_valid, int n_blocks, uint32_t a_off, int is_dir) {
        WgProblem q = {};
        q.d_off = sl.d_slot[lay]; q.a_off = a_off; q.w_off = w_off; q.ld = ld; q.col0 = col0; q.n_valid = n_valid;
        q.m_blocks = (uint8_t)(pl.fwd[lay].n_halves * 2); q.n_blocks = (uint8_t)n_blocks; q.is_dir = (uint8_t)is_dir;
        q.layer = is_dir ? (uint8_t)0xff : (uint8_t)lay;      // each dY_l is reduced exactly once
        q.sigma = (!is_dir && pl.fwd[lay].has_dir) ? 1 : 0;   // encoding_viewdir: its input is f, the sigma head's input
        wp.prob[np++] = q;
    };
    {
        int l = 0;
        add(l, L.enc_xyz_w, L.d_xyz, 0, L.d_xyz, 1, sl.a_slot[l], 0); ++l;
        for (int j = 0; j < c->shape_blocks; ++j, ++l) add(l, L.s_w[j], kW, 0, kW, 4, sl.a_slot[l], 0);
        add(l, L.enc_shape_w, kW, 0, kW, 4, sl.a_slot[l], 0); ++l;
        add(l, L.enc_vd_w, kW + L.d_dir, 0, kW, 4, sl.a_slot[l], 0);
        add(l, L.enc_vd_w, kW + L.d_dir, kW, L.d_dir, 0, sl.dir_slot, 1); ++l;
        for (int j = 0; j < c->texture_blocks; ++j, ++l) add(l, L.t_w[j], kW, 0, kW, 4, sl.a_slot[l], 0);
        add(l, L.rgb0_w, kW, 0, kW, 4, sl.a_slot[l], 0);
    }
    // split rows into ~k3_items_per_sm work items per SM (each ends with an atomic flush of its dW tile), >= 8 half-tiles each
    const double k3_items_per_sm = (double)cnb_option("k3_items_per_sm_x10", 20) / 10.0;
    double total_cost = 0;
    for (int i = 0; i < np; ++i) total_cost += (double)wp.prob[i].m_blocks + (wp.prob[i].is_dir ? 0.5 : wp.prob[i].n_blocks);   // HBM bytes per half-tile
    const int64_t H = tiles * 2;
    int items = 0;
    for (int i = 0; i < np; ++i) {
        const double cost = (double)wp.prob[i].m_blocks + (wp.prob[i].is_dir ? 0.5 : wp.prob[i].n_blocks);
        int64_t sp = (int64_t)(k3_items_per_sm * sms * cost / total_cost + 0.5);
        if (sp > H / 8) sp = H / 8;
        if (sp < 1) sp = 1;
        wp.prob[i].splits = (int32_t)sp; wp.prob[i].item0 = items; items += (int)sp;
    }
    wp.n_problems = np; wp.n_items = items;
    wp.stashA = stashA; wp.stashD = stashD; wp.a_tile_bytes = sl.a_tile_bytes; wp.d_tile_bytes = sl.d_tile_bytes;
    wp.n_tiles = tiles; wp.dP = d_params;
    wp.colsum = w.colsum; wp.n_layers = nl; wp.n_codes = n_codes; wp.rows_per_code = rows_per_code; wp.row_offset = row_offset;
    wp.dspre = dspre_buf; wp.S = S; wp.d_wsigma = d_params + L.sigma_w; wp.d_bsigma = d_params + L.sigma_b;
    const size_t wsmem = 1024 + (size_t)kWgStages * kWgStage + 256;
    CNB_CUDA_TRY(cudaFuncSetAttribute(k_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem));
    int wgrid = items < sms ? items : sms;
    { const int64_t kg = cnb_option("k3_grid", 0); if (kg > 0 && kg < wgrid) wgrid = (int)kg; }      // (measurement: K3 on fewer SMs)
    if (piped && !last_sub && wgrid > pp->k3_sms) wgrid = pp->k3_sms;      // the last K3 overlaps nothing: whole GPU
    cnb_prof_begin(CNB_K_WGRAD, st2);
    k_wgrad<<<wgrid, kWgThreads, wsmem, st2>>>(wp);
    cnb_prof_end(CNB_K_WGRAD, st2);
    CNB_LAUNCH_CHECK();
    // rgb.2 head: reduced inside K2 by default (head_mma); otherwise from the stashed rgb.2 input r1 = the rgb.0 hidden
    // (the sigma head is reduced inside K3 either way)
    if (!bp.head_mma) {
        const int64_t hwant = (tiles + kHeadWarps / 2 - 1) / (kHeadWarps / 2);
        const int hgrid = (int)(hwant < 2 * sms ? hwant : 2 * sms);
        k_head_wgrad<<<hgrid, kHeadWarps * 32, 0, st2>>>(stashA, sl.a_tile_bytes, sl.a_slot[nl], d_rgbs, S, tiles,
                                                        d_params + L.rgb2_w, d_params + L.rgb2_b);
        CNB_LAUNCH_CHECK();
    }
    if (piped) {
        CNB_CUDA_TRY(cudaEventRecord(pp->k3_done[buf], pp->st3));
        pp->pending[buf] = true;
    }
    return CNB_OK;
}

// After all rows: bias gradients, the per-code folded-layer terms and the latent branches.
int finish_bwd(const cnb_net_config* c, const float* const* P, const Plan& pl, BwdWorkspace& w,
               const float* shape_codes, const float* tex_codes, int n_codes, float* d_params, float* d_shape,
               float* d_tex, cudaStream_t st) {
    CnbLayout L; cnb_make_layout(c, &L);
    const int nl = pl.n_layers, nf = pl.n_folded;
    if (d_params) {
        int64_t boff[kMaxLayers]; int l = 0;
        boff[l++] = L.enc_xyz_b;
        for (int j = 0; j < c->shape_blocks; ++j) boff[l++] = L.s_b[j];
        boff[l++] = L.enc_shape_b; boff[l++] = L.enc_vd_b;
        for (int j = 0; j < c->texture_blocks; ++j) boff[l++] = L.t_b[j];
        boff[l++] = L.rgb0_b;
        BiasOut bo = {};
        for (int i = 0; i < nl; ++i) { bo.db[i] = d_params + boff[i]; bo.n_out[i] = pl.fwd[i].n_halves * 128; }
        k_bias_from_colsum<<<dim3(2, nl), 128, 0, st>>>(w.colsum, n_codes, nl, bo);
        CNB_LAUNCH_CHECK();
    }
    CNB_CUDA_TRY(cudaMemsetAsync(d_shape, 0, sizeof(float) * (size_t)n_codes * c->latent_dim, st));
    CNB_CUDA_TRY(cudaMemsetAsync(d_tex, 0, sizeof(float) * (size_t)n_codes * c->latent_dim, st));
    const int64_t zld = (int64_t)nf * kW;
    for (int l = 0; l < nl; ++l) {
        const int j = pl.fwd[l].folded;
        if (j < 0) continue;
        const bool shape = j < c->shape_blocks;
        const int jj = shape ? j : j - c->shape_blocks;
        const int iw = shape ? L.i_s[jj] : L.i_t[jj];
        const int il = shape ? L.i_sl[jj] : L.i_tl[jj];
        const int64_t w_off = shape ? L.s_w[jj] : L.t_w[jj];
        const int64_t lw_off = shape ? L.sl_w[jj] : L.tl_w[jj];
        const int64_t lb_off = shape ? L.sl_b[jj] : L.tl_b[jj];
        k_fold_bwd<<<n_codes + (d_params ? kW : 0), kW, 0, st>>>(P[iw], w.colsum + (size_t)l * kW, (int64_t)nl * kW, w.fw.z + (size_t)j * kW,
                                                                zld, n_codes, w.dz + (size_t)j * kW, d_params ? d_params + w_off : nullptr);
        CNB_LAUNCH_CHECK();
        CNB_TRY(cnb_launch_latent_bwd(P[il], shape ? shape_codes : tex_codes, w.fw.z + (size_t)j * kW, w.dz + (size_t)j * kW,
                                      zld, n_codes, c->latent_dim, kW, d_params ? d_params + lw_off : nullptr,
                                      d_params ? d_params + lb_off : nullptr, shape ? d_shape : d_tex, st));
    }
    return CNB_OK;
}

}  // namespace

// ===========================================================================
namespace sm100 {

// Units of whole rays in whole tiles: N / gcd(N, 128) tiles hold 128 / gcd(N, 128) rays.
struct UnitShape { int tiles, rays; };
static UnitShape unit_shape(int N) {
    int a = N, b = kTileRows;
    while (b) { const int t = a % b; a = b; b = t; }
    return UnitShape{N / a, kTileRows / a};
}
constexpr int kMaxUnitTiles = 32;      // longer units (odd N) take the unfused path
static bool can_fuse_compositing(int N) { return N >= 1 && N <= kTileRows && unit_shape(N).tiles <= kMaxUnitTiles; }
// Rays per sub-batch: at most kMaxSubTiles tiles, whole units (so sub-batches start on tile boundaries and a tile
// never straddles rays of two sub-batches or, with rows_per_code % 128 == 0, two codes).  (Rounding the size down to
// whole waves of 2 x #SMs units was measured: the extra launch costs more than the fuller last round gains.)
static int64_t sub_batch_rays(int N, int64_t n_rays) {
    const int64_t max_rows = kMaxSubTiles * kTileRows;
    if (n_rays * N <= max_rows) return n_rays < 1 ? 1 : n_rays;
    const UnitShape us = unit_shape(N);
    int64_t units = kMaxSubTiles / us.tiles;
    if (units < 1) units = 1;
    return units * us.rays;
}

size_t bwd_workspace_bytes(const cnb_net_config* cfg, int64_t S, int64_t n_rays, int N, int n_codes, int fused) {
    Plan pl;
    if (make_plan(cfg, nullptr, &pl) != CNB_OK) return 256;
    int64_t sub_rows = S < kMaxSubTiles * kTileRows ? S : kMaxSubTiles * kTileRows;
    int64_t sub_rays = 0;
    if (fused) { sub_rays = sub_batch_rays(N, n_rays); sub_rows = sub_rays * N; }
    return carve_bwd(cfg, pl, n_codes, sub_rows, sub_rays, fused, 1, num_sms(), nullptr, nullptr, (pipeline_wanted() && S > sub_rows) ? 2 : 1) + 1024;
}

// Fused render backward (mode 1: seeds given; mode 2: L2 loss against target).
int render_backward(const cnb_net_config* cfg, const float* const* P, const void* packed, const cnb_ray_batch* rays,
                    int mode, const float* d_rgb, const float* d_depth, const float* target, float loss_scale,
                    float* rgb, float* depth, float* acc, float* sq_err, float* d_params, float* d_shape, float* d_tex,
                    void* ws, size_t ws_bytes, cudaStream_t st) {
    Plan pl;
    CNB_TRY(make_plan(cfg, P, &pl));
    const int N = rays->n_samples;
    if (N > 512) return CNB_E_UNSUPPORTED;
    const int64_t S = rays->n_rays * N;
    const int64_t rows_per_code = (int64_t)rays->segments_per_code * rays->rays_per_segment * N;
    if (rays->n_codes > 1 && (rows_per_code % kTileRows) != 0) return CNB_E_UNSUPPORTED;
    const int64_t sub_rays = sub_batch_rays(N, rays->n_rays);
    const int64_t sub_rows = sub_rays * N;
    BwdWorkspace w;
    const int64_t n_sub = (rays->n_rays + sub_rays - 1) / sub_rays;
    int nbuf = (pipeline_wanted() && n_sub > 1) ? 2 : 1;
    const int sms = num_sms();
    size_t need = carve_bwd(cfg, pl, rays->n_codes, sub_rows, sub_rays, 1, 1, sms, nullptr, nullptr, nbuf);
    if (nbuf > 1 && ws && ws_bytes < need) {     // a caller-sized workspace without the second hand-over copy: no overlap
        nbuf = 1;
        need = carve_bwd(cfg, pl, rays->n_codes, sub_rows, sub_rays, 1, 1, sms, nullptr, nullptr, nbuf);
    }
    if (!ws || ws_bytes < need) return CNB_E_WORKSPACE;
    if (((uintptr_t)ws & 255) != 0) return CNB_E_ALIGNMENT;
    carve_bwd(cfg, pl, rays->n_codes, sub_rows, sub_rays, 1, 1, sms, (char*)ws, &w, nbuf);
    Pipeline pipe;
    CNB_TRY(pipeline_begin(pipe, n_sub, d_params != nullptr));
    CNB_TRY(latent_and_fold(cfg, P, rays->shape_codes, rays->texture_codes, rays->n_codes, w.fw, st));
#if CNB_K2_BIAS_CONST || CNB_K2_HEADS_CONST
    ConstRowsScope crs;                    // launch-wide bias rows and head weights -> constant memory, for K2
    CNB_TRY(crs.begin(cfg, P, pl, w.fw.rows, st));
#endif
    CNB_CUDA_TRY(cudaMemsetAsync(w.colsum, 0, sizeof(float) * (size_t)rays->n_codes * pl.n_layers * kW, st));
    if (mode == 2 && sq_err)
        CNB_CUDA_TRY(cudaMemsetAsync(sq_err, 0, sizeof(float) * (size_t)(rays->n_rays / rays->rays_per_segment), st));
    CnbRaySource rs = cnb_make_ray_source(rays);
    const bool fuse = can_fuse_compositing(N);
    int sub = 0;
    for (int64_t r0 = 0; r0 < rays->n_rays; r0 += sub_rays, ++sub) {
        const int64_t nr = rays->n_rays - r0 < sub_rays ? rays->n_rays - r0 : sub_rays;
        const bool last_sub = r0 + sub_rays >= rays->n_rays;
        if (fuse) {
            // compositing, the loss seed and its backward run inside K2 (rays may straddle the tiles of a unit)
            FuseArgs fa = {};
            fa.kind = mode; fa.unit_tiles = unit_shape(N).tiles; fa.white_bg = rays->white_bg; fa.n_rays_total = rays->n_rays;
            fa.d_rgb = d_rgb; fa.d_depth = d_depth; fa.target = target; fa.loss_scale = loss_scale;
            fa.rgb = rgb; fa.depth = depth; fa.acc = acc; fa.sq_err = mode == 2 ? sq_err : nullptr;
            CNB_TRY(run_mlp_bwd(cfg, P, packed, pl, w, 0, &rs, nullptr, nullptr, nr * N, r0 * N, rays->n_codes,
                                rays->n_codes > 1 ? rows_per_code : S, nullptr, nullptr, d_params, st, &fa, &pipe, sub, last_sub));
            continue;
        }
        // general N: forward kernel (spilling per-sample sigma / rgb), compositing backward, then K2
        float* o_rgb = rgb ? rgb : w.ray_rgb - r0 * 3;
        float* o_depth = depth ? depth : w.ray_depth - r0;
        float* o_acc = acc ? acc : w.ray_acc - r0;
        CNB_TRY(launch_render_rays(cfg, P, packed, pl, rays, r0, nr, w.fw.folded, w.fw.rows, w.spill_sig, w.spill_rgb, o_rgb, o_depth,
                                   o_acc, st));
        const float* seed_rgb; const float* seed_depth = nullptr;
        if (mode == 2) {
            k_l2_seed_sm100<<<(unsigned)((nr + 255) / 256), 256, 0, st>>>(o_rgb, target, nr, r0, rays->rays_per_segment,
                                                                         loss_scale, w.ray_drgb, sq_err);
            CNB_LAUNCH_CHECK();
            seed_rgb = w.ray_drgb;
        } else {
            seed_rgb = d_rgb + r0 * 3;
            seed_depth = d_depth ? d_depth + r0 : nullptr;
        }
        // the per-sample d rgb is also read by the head-gradient kernel that runs with K3: it lives in the hand-over copy
        float* drgb_sub = (pipe.on && w.nbuf > 1 && (sub & 1)) ? w.drgb2 : w.drgb;
        if (pipe.on && w.nbuf > 1 && pipe.pending[sub & 1]) {
            CNB_CUDA_TRY(cudaStreamWaitEvent(st, pipe.k3_done[sub & 1], 0));
            pipe.pending[sub & 1] = false;
        }
        CNB_TRY(cnb_vr_backward_segments(w.spill_sig, w.spill_rgb, rays->z_vals, rays->z_per_segment,
                                         rays->rays_per_segment, r0, nr, N, rays->white_bg, seed_rgb, seed_depth, w.dsig,
                                         drgb_sub, st));
        CNB_TRY(run_mlp_bwd(cfg, P, packed, pl, w, 0, &rs, nullptr, nullptr, nr * N, r0 * N, rays->n_codes,
                            rays->n_codes > 1 ? rows_per_code : S, w.dsig, drgb_sub, d_params, st, nullptr, &pipe, sub, last_sub));
    }
    CNB_TRY(pipeline_end(pipe, st));
    return finish_bwd(cfg, P, pl, w, rays->shape_codes, rays->texture_codes, rays->n_codes, d_params, d_shape, d_tex, st);
}

}  // namespace sm100

int cnb_sm100_mlp_backward(const cnb_net_config* cfg, const float* const* P, const void* packed, const float* xyz,
                           const float* viewdir, const float* shape_codes, const float* tex_codes, int n_codes,
                           int64_t samples_per_code, int64_t S, const float* d_sigmas, const float* d_rgbs,
                           float* d_params, float* d_shape, float* d_tex, void* ws, size_t ws_bytes, cudaStream_t st) {
    Plan pl;
    CNB_TRY(make_plan(cfg, P, &pl));
    if (n_codes > 1 && (samples_per_code % kTileRows) != 0) return CNB_E_UNSUPPORTED;
    const int64_t sub_rows = sm100::sub_batch_rays(1, S);      // rows are "rays of one sample": whole waves of tiles
    BwdWorkspace w;
    const int64_t n_sub = sub_rows > 0 ? (S + sub_rows - 1) / sub_rows : 1;
    int nbuf = (pipeline_wanted() && n_sub > 1) ? 2 : 1;
    size_t need = carve_bwd(cfg, pl, n_codes, sub_rows, 0, 0, 1, num_sms(), nullptr, nullptr, nbuf);
    if (nbuf > 1 && ws && ws_bytes < need) {
        nbuf = 1;
        need = carve_bwd(cfg, pl, n_codes, sub_rows, 0, 0, 1, num_sms(), nullptr, nullptr, nbuf);
    }
    if (!ws || ws_bytes < need) return CNB_E_WORKSPACE;
    if (((uintptr_t)ws & 255) != 0) return CNB_E_ALIGNMENT;
    carve_bwd(cfg, pl, n_codes, sub_rows, 0, 0, 1, num_sms(), (char*)ws, &w, nbuf);
    Pipeline pipe;
    CNB_TRY(pipeline_begin(pipe, n_sub, d_params != nullptr));
    CNB_TRY(latent_and_fold(cfg, P, shape_codes, tex_codes, n_codes, w.fw, st));
#if CNB_K2_BIAS_CONST || CNB_K2_HEADS_CONST
    ConstRowsScope crs;
    CNB_TRY(crs.begin(cfg, P, pl, w.fw.rows, st));
#endif
    CNB_CUDA_TRY(cudaMemsetAsync(w.colsum, 0, sizeof(float) * (size_t)n_codes * pl.n_layers * kW, st));
    int sub = 0;
    for (int64_t r0 = 0; r0 < S; r0 += sub_rows, ++sub) {
        const int64_t m = S - r0 < sub_rows ? S - r0 : sub_rows;
        CNB_TRY(run_mlp_bwd(cfg, P, packed, pl, w, 1, nullptr, xyz + r0 * 3, viewdir + r0 * 3, m, r0, n_codes,
                            n_codes > 1 ? samples_per_code : S, d_sigmas + r0, d_rgbs + r0 * 3, d_params, st, nullptr, &pipe,
                            sub, r0 + sub_rows >= S));
    }
    CNB_TRY(pipeline_end(pipe, st));
    return finish_bwd(cfg, P, pl, w, shape_codes, tex_codes, n_codes, d_params, d_shape, d_tex, st);
}

size_t cnb_sm100_mlp_bwd_workspace_bytes(const cnb_net_config* cfg, int64_t S, int n_codes) {
    return sm100::bwd_workspace_bytes(cfg, S, 0, 1, n_codes, 0);
}

int cnb_sm100_has_backward(void) { return 1; }

int cnb_sm100_pipeline_timeouts_bwd(void) {
    unsigned int v = 0;
    if (cudaMemcpyFromSymbol(&v, umma::g_umma_timeout, sizeof(unsigned int)) != cudaSuccess) return -1;
    return (int)v;
}
