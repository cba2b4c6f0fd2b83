// render_sm100.cu -- CNB_PRECISION_BF16: the fused CodeNeRF render path on Blackwell tensor
// cores.  Hand-written sm_100a: tcgen05.mma (bf16 x bf16 -> fp32 in TMEM), weights streamed
// through shared memory by the TMA engine (1-D bulk copies of pre-swizzled stage images),
// mbarrier pipelines, warp-specialised roles.
//
// K1  k_render_fwd : [ray gen + sampling +] positional encoding -> 8 chained GEMM layers whose
//      activations live only in shared memory / TMEM -> sigma & rgb heads in the epilogue ->
//      per-ray transmittance scan + alpha compositing.  sigmas / rgbs never reach HBM
//      (unless the caller asks for them: unfused CodeNeRF.forward, or the training spill).
//
// Reference semantics: src/model.py:36-53 (network), src/utils.py:10-47 (rays, sampling,
// compositing).  See DESIGN.md for the tile / pipeline design and the roofline arithmetic.
#include <cstdlib>
#include "sm100_common.cuh"

using namespace sm100;

#ifdef CNB_TRACE
extern "C" int cnb_debug_trace(unsigned long long* out32, int reset) {
    if (out32 && cudaMemcpyFromSymbol(out32, sm100::g_trace, sizeof(unsigned long long) * 32) != cudaSuccess) return -1;
    if (reset) { unsigned long long z[32] = {}; if (cudaMemcpyToSymbol(sm100::g_trace, z, sizeof(z)) != cudaSuccess) return -1; }
    return 0;
}
#endif

// 1: K1 reads the launch-wide bias rows and the head weights from constant memory (sm100_common.cuh, ConstRows);
// 0: from shared memory like the per-code rows.  Measured (round 2, A/B in one session, 30 steps): constant memory is
// slower, 17.0-18.0 vs 18.3-18.8 M rays/s -- the uniform loads are served at a lower rate than LDS broadcasts.
#ifndef CNB_K1_CONST
#define CNB_K1_CONST 0
#endif

namespace {

struct FwdParams {
    WeightMaps maps;              // tensor maps of the packed weights (CTA-pair kernels)
    int n_layers;
    FwdLayer layers[kMaxLayers];
    const uint8_t* packed;
    const float* folded;          // [n_codes][n_folded][256]: b_j + W_j z_j  (z_j = ReLU(latent layer))
    int n_folded, n_codes;
    int64_t rows_per_code;
    const float *w_sigma, *b_sigma, *w_rgb2, *b_rgb2;
    int mode;                     // 0: rays -> composite; 1: xyz / viewdir arrays -> sigmas, rgbs
    CnbRaySource rs;
    const float *xyz, *viewdir;
    int64_t n_rays, S;            // rays / rows of THIS launch
    int64_t ray_offset;           // mode 0: global index of this launch's first ray (sub-batching)
    int white_bg, ring_cap;
    int share_fills;                // CTA pairs: group Y's GEMM reuses group X's weight fill of a layer that fills the ring exactly once
    int stage_bias;               // 1: shared memory has room for the bias / head-weight staging area
    float *rgb, *depth, *acc;
    float *spill_sig, *spill_rgb; // optional per-sample spill (launch-relative rows) for the training backward
    float *sigmas, *rgbs;
};

__device__ __forceinline__ float warp_excl_prod_f(float local, int lane) {
    float incl = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const float up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl *= up;
    }
    const float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    return lane == 0 ? 1.f : excl;
}

// Alpha compositing of one ray by one warp from the shared-memory sample ring --
// reference src/utils.py:34-47 (exclusive transmittance product, white background).
__device__ __forceinline__ void composite_ray(const FwdParams& p, const float4* ring, int ring_cap, int64_t lrow0,
                                              int64_t gray, int lane) {
    const int N = p.rs.N;
    const int64_t seg = gray / p.rs.rays_per_segment;
    const float* z = p.rs.z_vals + (p.rs.z_per_segment ? seg * N : 0);
    const int per = (N + 31) >> 5;
    const int i0 = lane * per;
    float T_local = 1.f;
    for (int j = 0; j < per; ++j) {
        const int i = i0 + j;
        if (i < N) {
            const float4 s = ring[(lrow0 + i) % ring_cap];
            const float delta = (i + 1 < N) ? (__ldg(z + i + 1) - __ldg(z + i)) : 1e10f;
            const float alpha = 1.f - expf(-s.x * delta);
            T_local *= (1.f - alpha + 1e-10f);
        }
    }
    float T = warp_excl_prod_f(T_local, lane);
    float cr = 0.f, cg = 0.f, cb = 0.f, d = 0.f, ws = 0.f;
    for (int j = 0; j < per; ++j) {
        const int i = i0 + j;
        if (i < N) {
            const float4 s = ring[(lrow0 + i) % ring_cap];
            const float zi = __ldg(z + i);
            const float delta = (i + 1 < N) ? (__ldg(z + i + 1) - zi) : 1e10f;
            const float alpha = 1.f - expf(-s.x * delta);
            const float w = alpha * T;
            cr += w * s.y; cg += w * s.z; cb += w * s.w;
            d += w * zi; ws += w;
            T *= (1.f - alpha + 1e-10f);
        }
    }
    cr = warp_sum_f(cr); cg = warp_sum_f(cg); cb = warp_sum_f(cb); d = warp_sum_f(d); ws = warp_sum_f(ws);
    if (lane == 0) {
        if (p.white_bg) { cr = cr + 1.f - ws; cg = cg + 1.f - ws; cb = cb + 1.f - ws; }
        p.rgb[gray * 3 + 0] = cr; p.rgb[gray * 3 + 1] = cg; p.rgb[gray * 3 + 2] = cb;
        p.depth[gray] = d;
        if (p.acc) p.acc[gray] = ws;
    }
}

// CG = 1: one CTA per SM, M = 128 MMAs.  CG = 2: CTA pairs (cluster of 2), tcgen05 cta_group::2 with M = 256:
// each CTA streams half of every weight chunk, halving L2 and shared-memory operand traffic per row.
// MC > 1 (with CG = 1): clusters of MC CTAs with independent M = 128 MMAs whose weight stream is multicast --
// every stage leaves L2 once per cluster (the weight stream from L2 is what bounds the CG = 1, MC = 1 kernel).
// EW = 4: each tile has its own four epilogue warps (X / Y groups).  EW = 8: all eight warps drain every accumulator
// (two per TMEM lane quarter, half of the columns each): half the epilogue latency per tile -- what the CTA-pair
// kernel needs, whose per-tile chain (epilogue -> remote arrive -> MMA -> multicast commit) is the limit.
template <int CG, int MC, int EW = 4>
__global__ void __launch_bounds__(kThreads, 1) k_render_fwd(const __grid_constant__ FwdParams p) {
    static_assert(CG == 1 || MC == 1, "pairs and multicast clusters are alternatives");
    constexpr int kCluster = CG * MC;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA0 = smem;
    uint8_t* sW = smem + 2 * kATile;
    float4* sRing = (float4*)(sW + kNumStages * kSlot);
    uint64_t* bars = (uint64_t*)((uint8_t*)sRing + (size_t)p.ring_cap * 16);
    uint64_t* w_full = bars;
    uint64_t* w_empty = bars + kNumStages;
    uint64_t* a_ready = bars + 2 * kNumStages;
    uint64_t* acc_full = a_ready + 2;
    uint64_t* w_full_peer = acc_full + 2;        // [kNumStages] (leader of a CTA pair): the peer's half has landed
    uint32_t* tmem_slot = (uint32_t*)(w_full_peer + kNumStages);
    volatile int* final_count = (volatile int*)(tmem_slot + 2);
    float* sBias = (float*)(tmem_slot + 8);      // [2 groups][2 buffers][256], then sigma-head [256] and rgb.2 [3][128] weights
    float* sWsig = sBias + 4 * kW;
    float* sWrgb = sWsig + kW;
    const uint32_t rank = kCluster > 1 ? umma::cluster_ctarank() : 0u;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nl = p.n_layers;

    // ---- this CTA's contiguous share of the work -------------------------------------------
    int64_t row0 = 0, nrows = 0, ray0 = 0;
    int T = 0;                                   // tile slots of this CTA (uniform over its cluster)
    for (int who = 0; who < kCluster; ++who) {   // only the tile count of the other CTAs of the cluster matters
        const int64_t b = (int64_t)blockIdx.x - rank + who;
        int64_t r0_, n_, ray0_ = 0;
        if (p.mode == 0) {
            const int64_t base = p.n_rays / gridDim.x, rem = p.n_rays % gridDim.x;
            ray0_ = b * base + min(b, rem);
            const int64_t nr = base + (b < rem ? 1 : 0);
            r0_ = ray0_ * p.rs.N; n_ = nr * p.rs.N;
        } else {
            const int64_t tiles = (p.S + kTileRows - 1) / kTileRows;
            const int64_t base = tiles / gridDim.x, rem = tiles % gridDim.x;
            const int64_t t0 = b * base + min(b, rem);
            const int64_t nt = base + (b < rem ? 1 : 0);
            r0_ = t0 * kTileRows;
            n_ = min(p.S - r0_, nt * kTileRows);
            if (n_ < 0) n_ = 0;
        }
        if ((uint32_t)who == rank) { row0 = r0_; nrows = n_; ray0 = ray0_; }
        T = max(T, (int)((n_ + kTileRows - 1) / kTileRows));
    }
    const int rounds = (T + 1) >> 1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kNumStages; ++i) { umma::mbar_init(&w_full[i], 1); umma::mbar_init(&w_empty[i], MC); }
        for (int i = 0; i < kNumStages; ++i) umma::mbar_init(&w_full_peer[i], 1);
        for (int g = 0; g < 2; ++g) { umma::mbar_init(&a_ready[g], EW * CG); umma::mbar_init(&acc_full[g], 1); }
        final_count[0] = 0; final_count[1] = 0;
        umma::fence_mbar_init();
    }
    if (warp == 1) { if (CG == 2) umma::tmem_alloc2(tmem_slot, 512); else umma::tmem_alloc(tmem_slot, 512); }
    if (p.stage_bias) {
        for (int i = threadIdx.x; i < kW; i += kThreads) sWsig[i] = __ldg(p.w_sigma + i);
        for (int i = threadIdx.x; i < 3 * (kW / 2); i += kThreads) sWrgb[i] = __ldg(p.w_rgb2 + i);
    }
    umma::tc_fence_before();
    if (kCluster > 1) umma::cluster_sync_all(); else __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // a_ready lives in the leader CTA: the partner's compute warps arrive on it remotely
    const uint32_t a_ready_addr0 = CG == 2 ? umma::mapa(umma::smem_u32(&a_ready[0]), 0) : 0u;

    if (warp < 4) {
    umma::setmaxnreg_dec<kRegsAux>();
    if (warp == 0) {
        // ===== weight producer: TMA-engine bulk copies of stage images into the ring =====
        int stage = 0; uint32_t ph = 0;
        CNB_TR_DECL(tr_we); CNB_TR_DECL(tr_tot);
        const long long tr_t0 = CNB_TR_NOW();
        for (int r = 0; r < rounds; ++r)
            for (int l = 0; l < nl; ++l)
                for (int g = 0; g < 2; ++g) {
                    if (2 * r + g >= T) continue;
                    const FwdLayer& L = p.layers[l];
                    const int n_dir = L.has_dir ? L.n_halves : 0;
                    if (CG == 2) produce_stages_2cta(&p.maps, L.w_off, L.n_kchunks, L.n_halves, L.has_dir, rank, sW, w_full, w_empty, stage, ph, 0ull,
                                                     g == 1 && p.share_fills && L.n_kchunks + (L.has_dir ? 1 : 0) == kNumStages);
                    else produce_stages<MC>(p.packed + L.w_off, L.n_kchunks * L.n_halves + n_dir, n_dir, sW, w_full, w_empty, stage, ph, rank, &tr_we);
                }
        tr_tot = (unsigned long long)(CNB_TR_NOW() - tr_t0);
        CNB_TR_FLUSH(0, tr_we); CNB_TR_FLUSH(1, tr_tot);
    } else if (warp == 1) {
        // ===== MMA issuer: one elected thread drives the tensor core for both tiles =====
        int stage = 0; uint32_t ph = 0;
        CNB_TR_DECL(tr_wa); CNB_TR_DECL(tr_ww); CNB_TR_DECL(tr_tot);
        const long long tr_t0 = CNB_TR_NOW();
        for (int r = 0; r < rounds; ++r)
            for (int l = 0; l < nl; ++l)
                for (int g = 0; g < 2; ++g) {
                    if (2 * r + g >= T) continue;
                    const FwdLayer& L = p.layers[l];
                    if (CG == 2 && rank != 0) continue;      // the partner CTA issues no MMAs
                    if (CG == 2) CNB_TR(tr_wa, umma::mbar_wait_cluster(&a_ready[g], (uint32_t)(r * nl + l) & 1u));
                    else CNB_TR(tr_wa, umma::mbar_wait(&a_ready[g], (uint32_t)(r * nl + l) & 1u));
                    umma::tc_fence_after();
                    if (CG == 2)
                        issue_gemm_2cta(umma::smem_u32(sA0 + g * kATile), tmem + (uint32_t)g * 256u, sW, w_full,
                                        w_empty, L.n_kchunks, L.n_halves, L.has_dir, stage, ph, &acc_full[g], &tr_ww);
                    else
                        issue_gemm<MC>(umma::smem_u32(sA0 + g * kATile), tmem + (uint32_t)g * 256u, sW, w_full, w_empty,
                                       L.n_kchunks, L.n_halves, L.has_dir, stage, ph, &acc_full[g], &tr_ww);
                }
        tr_tot = (unsigned long long)(CNB_TR_NOW() - tr_t0);
        CNB_TR_FLUSH(2, tr_wa); CNB_TR_FLUSH(3, tr_ww); CNB_TR_FLUSH(4, tr_tot);
    }
    } else if constexpr (EW == 8) {
        umma::setmaxnreg_inc<kRegsCompute>();
        // ===== eight epilogue warps serving tiles X and Y alternately, in the order their accumulators complete =====
        const int q = warp & 3;                      // TMEM lane quarter
        const int ch = (warp - 4) >> 2;              // column half of a 256-wide layer (64-column half of rgb.0)
        const int w8 = warp - 4;
        const int row = q * 32 + lane;
        const int tid8 = w8 * 32 + lane;             // 0 .. 255
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int N = p.rs.N;
        uint32_t a8x[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) a8x[c] = umma::smem_u32(sA0 + row * 128 + ((c ^ (row & 7)) << 4));
        float4* sHead = (float4*)(sBias + 2 * kW);   // [128] partial heads of the second column-half warp (second half of the staging area)
        uint64_t sig0 = 0ull, sig1 = 0ull;           // sigma-head partial sums, per tile slot
        uint32_t bsel = 0;
        CNB_TR_DECL(tr_wacc); CNB_TR_DECL(tr_epi); CNB_TR_DECL(tr_enc); CNB_TR_DECL(tr_tot);
        const long long tr_t0 = CNB_TR_NOW();

        // this thread's part of a tile's encodings: xyz (32 words, ch == 0) or view direction (16 words, ch == 1)
        auto prepare_tile = [&](int t, uint32_t (&w)[32]) {
            const int64_t lr = (int64_t)t * kTileRows + row;
            const bool ok = t < T && lr < nrows;
            const int64_t gr = row0 + lr;
            float pos[3] = {0.f, 0.f, 0.f}, dir[3] = {0.f, 0.f, 0.f};
            if (ok) {
                if (p.mode == 0) {
                    const int64_t lray = gr / N;
                    const int zi = (int)(gr - lray * N);
                    const int64_t ray = p.ray_offset + lray;
                    float o[3];
                    cnb_fetch_ray(p.rs, ray, o, dir);
                    const int64_t seg = ray / p.rs.rays_per_segment;
                    const float z = __ldg(p.rs.z_vals + (p.rs.z_per_segment ? seg * N : 0) + zi);
#pragma unroll
                    for (int k = 0; k < 3; ++k) pos[k] = cnb_sample_coord(o[k], dir[k], z);
                } else {
#pragma unroll
                    for (int k = 0; k < 3; ++k) { pos[k] = __ldg(p.xyz + gr * 3 + k); dir[k] = __ldg(p.viewdir + gr * 3 + k); }
                }
            }
            if (ch == 0) pe_compute_xyz(pos, ok, w);
            else {
                uint32_t d[16];
                pe_compute_dir(dir, ok, d);
#pragma unroll
                for (int i = 0; i < 16; ++i) w[i] = d[i];
            }
        };
        uint32_t pe0[32], pe1[32];
        prepare_tile(0, pe0);
        prepare_tile(1, pe1);

        auto tile_op = [&](const int g, int r, int op) {
            const int t = 2 * r + g;
            const int64_t lrow = (int64_t)t * kTileRows + row;
            const bool valid = lrow < nrows;
            const int64_t grow = row0 + lrow;
            uint8_t* sA = sA0 + g * kATile;
            const uint32_t taddr = tmem + lane_off + (uint32_t)g * 256u;
            auto publish = [&]() {
                umma::tc_fence_before();
                umma::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) { if (CG == 2) umma::mbar_arrive_cluster(a_ready_addr0 + g * 8); else umma::mbar_arrive(&a_ready[g]); }
            };
            if (op == 0) {
                const long long tr_e0 = CNB_TR_NOW();
                uint32_t (&pe)[32] = g ? pe1 : pe0;
                if (ch == 0) pe_store_xyz(pe, sA, row);
                else {
                    uint32_t d[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) d[i] = pe[i];
                    pe_store_dir(d, sA + 4 * kABlock, row);
                }
                publish();
                tr_enc += (unsigned long long)(CNB_TR_NOW() - tr_e0);
                return;
            }
            const int l = op - 1;
            const FwdLayer& L = p.layers[l];
            const int n_out = L.n_halves * 128;
            int64_t code = 0;
            if (p.n_codes > 1) {
                code = (p.ray_offset * N + (valid ? grow : row0)) / p.rows_per_code;
                if (code >= p.n_codes) code = p.n_codes - 1;
            }
            bool staged = p.stage_bias != 0;
            int64_t code_b = code;
            if (staged && p.n_codes > 1) {
                const int64_t last_l = min((int64_t)t * kTileRows + kTileRows, nrows) - 1;
                const int64_t c_lo = min((p.ray_offset * N + row0 + (int64_t)t * kTileRows) / p.rows_per_code, (int64_t)p.n_codes - 1);
                const int64_t c_hi = min((p.ray_offset * N + row0 + last_l) / p.rows_per_code, (int64_t)p.n_codes - 1);
                staged = c_lo == c_hi;
                if (staged) code_b = c_lo;
            }
            const float* bias_g = L.folded >= 0 ? p.folded + ((size_t)code_b * p.n_folded + L.folded) * kW : L.bias;
            float bias1 = 0.f;
            if (staged && tid8 < n_out) bias1 = __ldg(bias_g + tid8);      // in flight during the wait
            // the next round's encodings are computed inside the waits of the last layers
            if (l == nl - 3 && g == 0) prepare_tile(t + 2, pe0);
            if (l == nl - 2 && g == 1) prepare_tile(t + 2, pe1);
            CNB_TR(tr_wacc, umma::mbar_wait(&acc_full[g], (uint32_t)(r * nl + l) & 1u));
            const long long tr_p0 = CNB_TR_NOW();
            umma::tc_fence_after();
            HeadAcc hacc = {l == 0 ? 0ull : (g ? sig1 : sig0), 0ull, 0ull, 0ull, 0ull};
            uint32_t a8h[8];
            const int kblk = L.n_halves == 2 ? ch * 2 : ch;
#pragma unroll
            for (int c = 0; c < 8; ++c) a8h[c] = a8x[c] + (uint32_t)(g * kATile + kblk * kABlock);
            const int c0 = L.n_halves == 2 ? ch * 128 : ch * 64;       // first column of this thread
            const uint32_t th = taddr + (uint32_t)c0;
            if (staged) {
                float* sb = sBias + bsel * kW; bsel ^= 1u;
                if (tid8 < n_out) sb[tid8] = bias1;
                const uint32_t tok = bar_sync_token(1, 256);
                const float* bs = smem_fptr(sb, tok) + c0;
                const float* ws = smem_fptr(sWsig, tok) + c0;
                const float* wr = smem_fptr(sWrgb, tok) + c0;
                if (L.kind == 0) fwd_epilogue_chunks<4, 0, true, false, true>(th, bs, a8h, ws, wr, hacc, nullptr);
                else if (L.kind == 1) fwd_epilogue_chunks<4, 1, true, false, true>(th, bs, a8h, ws, wr, hacc, nullptr);
                else fwd_epilogue_chunks<2, 2, false, false, true>(th, bs, a8h, ws, wr, hacc, nullptr);
            } else {
                const float* bias_t = (L.folded >= 0 ? p.folded + ((size_t)code * p.n_folded + L.folded) * kW : L.bias) + c0;
                if (L.kind == 0) fwd_epilogue_chunks<4, 0, true, false, false>(th, bias_t, a8h, p.w_sigma + c0, p.w_rgb2 + c0, hacc, nullptr);
                else if (L.kind == 1) fwd_epilogue_chunks<4, 1, true, false, false>(th, bias_t, a8h, p.w_sigma + c0, p.w_rgb2 + c0, hacc, nullptr);
                else fwd_epilogue_chunks<2, 2, false, false, false>(th, bias_t, a8h, p.w_sigma + c0, p.w_rgb2 + c0, hacc, nullptr);
            }
            if (g) sig1 = hacc.sig2; else sig0 = hacc.sig2;
            if (l + 1 < nl) { publish(); tr_epi += (unsigned long long)(CNB_TR_NOW() - tr_p0); return; }
            tr_epi += (unsigned long long)(CNB_TR_NOW() - tr_p0);

            // ---- heads: the two column halves of a row exchange their partial sums, then one of them finishes ----
            float sig_p, cr, cg, cb;
            { float a0, a1; unpk2(hacc.sig2, a0, a1); sig_p = a0 + a1; unpk2(hacc.r2, a0, a1); cr = a0 + a1;
              unpk2(hacc.g2, a0, a1); cg = a0 + a1; unpk2(hacc.b2, a0, a1); cb = a0 + a1; }
            if (ch == 1) sHead[row] = make_float4(sig_p, cr, cg, cb);
            umma::named_bar_sync(1, 256);
            if (ch == 0) {
                const float4 o = sHead[row];
                const float sigma = cnb_softplus(sig_p + o.x + __ldg(p.b_sigma));
                cr += o.y + __ldg(p.b_rgb2 + 0); cg += o.z + __ldg(p.b_rgb2 + 1); cb += o.w + __ldg(p.b_rgb2 + 2);
                if (p.mode == 1) {
                    if (valid) {
                        p.sigmas[grow] = sigma;
                        p.rgbs[grow * 3 + 0] = cr; p.rgbs[grow * 3 + 1] = cg; p.rgbs[grow * 3 + 2] = cb;
                    }
                } else if (valid) {
                    sRing[lrow % p.ring_cap] = make_float4(sigma, cr, cg, cb);
                    if (p.spill_sig) {
                        p.spill_sig[grow] = sigma;
                        p.spill_rgb[grow * 3 + 0] = cr; p.spill_rgb[grow * 3 + 1] = cg; p.spill_rgb[grow * 3 + 2] = cb;
                    }
                }
            }
            umma::named_bar_sync(1, 256);          // samples of the tile are in the ring; sHead may be rewritten
            if (p.mode == 0) {
                const int64_t tile_lo = (int64_t)t * kTileRows, tile_hi = min(tile_lo + kTileRows, nrows);
                const int64_t q_first = tile_lo / N;        // first ray whose last row lies in this tile
                const int64_t q_last = tile_hi / N - 1;     // last ray completed by the end of this tile
                for (int64_t qr = q_first + w8; qr <= q_last; qr += 8)
                    composite_ray(p, sRing, p.ring_cap, qr * N, p.ray_offset + ray0 + qr, lane);
            }
        };
        for (int r = 0; r < rounds; ++r) {
            const bool two = 2 * r + 1 < T;
            for (int op = 0; op <= nl; ++op)
                for (int gi = 0; gi < (two ? 2 : 1); ++gi) tile_op(gi, r, op);      // one call site: the lambda is inlined
        }
        tr_tot = (unsigned long long)(CNB_TR_NOW() - tr_t0);
        if (warp == 4) { CNB_TR_FLUSH(5, tr_wacc); CNB_TR_FLUSH(6, tr_epi); CNB_TR_FLUSH(7, tr_enc); CNB_TR_FLUSH(8, tr_tot); }
    } else {
        umma::setmaxnreg_inc<kRegsCompute>();
        // ===== compute groups: PE, per-layer epilogues (TMEM -> bias/ReLU -> bf16 operand), heads, compositing =====
        const int g = (warp - 4) >> 2;
        const int q = warp & 3;                  // TMEM lane quarter this warp may access
        const int wi = q;
        const int row = q * 32 + lane;
        uint8_t* sA = sA0 + g * kATile;
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)g * 256u;
        const int N = p.rs.N;
        uint32_t a8[8];     // shared address of each 16-byte chunk of this row inside K-block 0 (128-B swizzle)
#pragma unroll
        for (int c = 0; c < 8; ++c) a8[c] = umma::smem_u32(sA + row * 128 + ((c ^ (row & 7)) << 4));
        CNB_TR_DECL(tr_wacc); CNB_TR_DECL(tr_epi); CNB_TR_DECL(tr_enc); CNB_TR_DECL(tr_tot);
        const long long tr_t0 = CNB_TR_NOW();
        const int tg = wi * 32 + lane;               // thread index inside the group
        uint32_t bsel = 0;                           // bias staging buffer of the next layer (alternates)
        // rays / samples / positional encodings of tile slot t of this group, as packed bf16 rows in registers
        auto prepare_tile = [&](int t, PeRow& pe) {
            const int64_t lr = (int64_t)t * kTileRows + row;
            const bool ok = t < T && lr < nrows;
            const int64_t gr = row0 + lr;
            float pos[3] = {0.f, 0.f, 0.f}, dir[3] = {0.f, 0.f, 0.f};
            if (ok) {
                if (p.mode == 0) {
                    const int64_t lray = gr / N;
                    const int zi = (int)(gr - lray * N);
                    const int64_t ray = p.ray_offset + lray;
                    float o[3];
                    cnb_fetch_ray(p.rs, ray, o, dir);
                    const int64_t seg = ray / p.rs.rays_per_segment;
                    const float z = __ldg(p.rs.z_vals + (p.rs.z_per_segment ? seg * N : 0) + zi);
#pragma unroll
                    for (int k = 0; k < 3; ++k) pos[k] = cnb_sample_coord(o[k], dir[k], z);
                } else {
#pragma unroll
                    for (int k = 0; k < 3; ++k) { pos[k] = __ldg(p.xyz + gr * 3 + k); dir[k] = __ldg(p.viewdir + gr * 3 + k); }
                }
            }
            pe_compute_xyz(pos, ok, pe.x);
            pe_compute_dir(dir, ok, pe.d);
        };
        PeRow pe;
        if (g < T) prepare_tile(g, pe);
        for (int r = 0; r < rounds; ++r) {
            const int t = 2 * r + g;
            if (t >= T) break;
            const long long tr_e0 = CNB_TR_NOW();
            const int64_t lrow = (int64_t)t * kTileRows + row;
            const bool valid = lrow < nrows;
            const int64_t grow = row0 + lrow;
            int64_t code = 0;
            if (p.n_codes > 1) {
                code = (p.ray_offset * N + (valid ? grow : row0)) / p.rows_per_code;
                if (code >= p.n_codes) code = p.n_codes - 1;
            }
            // ---- positional encodings (computed during the previous tile's last layers) -> operand blocks ----
            pe_store_xyz(pe.x, sA, row);
            pe_store_dir(pe.d, sA + 4 * kABlock, row);
            umma::tc_fence_before();
            umma::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) { if (CG == 2) umma::mbar_arrive_cluster(a_ready_addr0 + g * 8); else umma::mbar_arrive(&a_ready[g]); }
            tr_enc += (unsigned long long)(CNB_TR_NOW() - tr_e0);

            HeadAcc hacc = {0ull, 0ull, 0ull, 0ull, 0ull};
            // bias rows go through shared memory when the whole tile uses one code (always, unless a tile straddles objects)
            bool staged = p.stage_bias != 0;
            int64_t code_b = code;                   // code whose bias rows this thread reads (staged: the tile's code)
            if (staged && p.n_codes > 1) {
                const int64_t last_l = min((int64_t)t * kTileRows + kTileRows, nrows) - 1;
                const int64_t c_lo = min((p.ray_offset * N + row0 + (int64_t)t * kTileRows) / p.rows_per_code, (int64_t)p.n_codes - 1);
                const int64_t c_hi = min((p.ray_offset * N + row0 + last_l) / p.rows_per_code, (int64_t)p.n_codes - 1);
                staged = c_lo == c_hi;
                if (staged) code_b = c_lo;
            }
            for (int l = 0; l < nl; ++l) {
                const FwdLayer& L = p.layers[l];
                const float* bias = L.folded >= 0 ? p.folded + ((size_t)code_b * p.n_folded + L.folded) * kW : L.bias;
                float2 bias2 = make_float2(0.f, 0.f);
                const bool mine = 2 * tg < L.n_halves * 128;
                if (staged && mine && (L.folded >= 0 || !CNB_K1_CONST)) bias2 = __ldg(reinterpret_cast<const float2*>(bias) + tg);    // in flight during the wait
                if (l == nl - 2) prepare_tile(t + 2, pe);      // next tile of this group: its PE is computed inside this wait
                CNB_TR(tr_wacc, umma::mbar_wait(&acc_full[g], (uint32_t)(r * nl + l) & 1u));
                const long long tr_p0 = CNB_TR_NOW();
                umma::tc_fence_after();
                if (staged && !CNB_K1_CONST) {
                    float* sb = sBias + (g * 2 + bsel) * kW; bsel ^= 1u;
                    if (mine) *reinterpret_cast<float2*>(sb + 2 * tg) = bias2;
                    const uint32_t tok = bar_sync_token(1 + g, 128);
                    const float* bs = smem_fptr(sb, tok);
                    const float* ws = smem_fptr(sWsig, tok);
                    const float* wr = smem_fptr(sWrgb, tok);
                    if (L.kind == 0) fwd_epilogue_layer<8, 0, true, false, 1>(taddr, bs, a8, ws, wr, hacc, nullptr);
                    else if (L.kind == 1) fwd_epilogue_layer<8, 1, true, false, 1>(taddr, bs, a8, ws, wr, hacc, nullptr);
                    else fwd_epilogue_layer<4, 2, false, false, 1>(taddr, bs, a8, ws, wr, hacc, nullptr);
                } else if (staged) {
                    // head weights and the bias rows of the layers that are not code-conditioned: constant memory
                    // (uniform loads, no shared-memory wavefronts); a folded layer's per-code row is staged in shared memory
                    const float* ws = crow_ptr(kCrowWsig);
                    const float* wr = crow_ptr(kCrowWrgb);
                    if (L.folded >= 0) {
                        float* sb = sBias + (g * 2 + bsel) * kW; bsel ^= 1u;
                        if (mine) *reinterpret_cast<float2*>(sb + 2 * tg) = bias2;
                        const uint32_t tok = bar_sync_token(1 + g, 128);
                        fwd_epilogue_layer<8, 0, true, false, 1, 2>(taddr, smem_fptr(sb, tok), a8, ws, wr, hacc, nullptr);
                    } else {
                        // fixed constant-memory slots (crow_slot): immediate addresses, uniform loads
                        if (L.kind == 0) {
                            if (l == 0) fwd_epilogue_layer<8, 0, true, false, 2>(taddr, crow_ptr(kCrowBias + 0 * kW), a8, ws, wr, hacc, nullptr);
                            else fwd_epilogue_layer<8, 0, true, false, 2>(taddr, crow_ptr(kCrowBias + 2 * kW), a8, ws, wr, hacc, nullptr);
                        }
                        else if (L.kind == 1) fwd_epilogue_layer<8, 1, true, false, 2>(taddr, crow_ptr(kCrowBias + 1 * kW), a8, ws, wr, hacc, nullptr);
                        else fwd_epilogue_layer<4, 2, false, false, 2>(taddr, crow_ptr(kCrowBias + 3 * kW), a8, ws, wr, hacc, nullptr);
                    }
                } else {
                    if (L.kind == 0) fwd_epilogue_layer<8, 0, true, false>(taddr, bias, a8, p.w_sigma, p.w_rgb2, hacc, nullptr);
                    else if (L.kind == 1) fwd_epilogue_layer<8, 1, true, false>(taddr, bias, a8, p.w_sigma, p.w_rgb2, hacc, nullptr);
                    else fwd_epilogue_layer<4, 2, false, false>(taddr, bias, a8, p.w_sigma, p.w_rgb2, hacc, nullptr);
                }
                if (l + 1 < nl) {
                    umma::tc_fence_before();
                    umma::fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) { if (CG == 2) umma::mbar_arrive_cluster(a_ready_addr0 + g * 8); else umma::mbar_arrive(&a_ready[g]); }
                }
                tr_epi += (unsigned long long)(CNB_TR_NOW() - tr_p0);
            }
            float sig_pre, cr, cg, cb;
            { float a0, a1; unpk2(hacc.sig2, a0, a1); sig_pre = a0 + a1; unpk2(hacc.r2, a0, a1); cr = a0 + a1;
              unpk2(hacc.g2, a0, a1); cg = a0 + a1; unpk2(hacc.b2, a0, a1); cb = a0 + a1; }
            // ---- heads -> outputs ----
            const float sigma = cnb_softplus(sig_pre + __ldg(p.b_sigma));
            cr += __ldg(p.b_rgb2 + 0); cg += __ldg(p.b_rgb2 + 1); cb += __ldg(p.b_rgb2 + 2);
            if (p.mode == 1) {
                if (valid) {
                    p.sigmas[grow] = sigma;
                    p.rgbs[grow * 3 + 0] = cr; p.rgbs[grow * 3 + 1] = cg; p.rgbs[grow * 3 + 2] = cb;
                }
                continue;
            }
            const float4 smp = make_float4(sigma, cr, cg, cb);
            if (valid) {
                sRing[lrow % p.ring_cap] = smp;
                if (p.spill_sig) {
                    p.spill_sig[grow] = sigma;
                    p.spill_rgb[grow * 3 + 0] = cr; p.spill_rgb[grow * 3 + 1] = cg; p.spill_rgb[grow * 3 + 2] = cb;
                }
            }
            umma::named_bar_sync(1 + g, 128);
            if (row == 0) { __threadfence_block(); final_count[g] = r + 1; }
            {   // rays ending in this tile may have begun in the other group's previous tile
                const int need = (g == 1) ? r + 1 : r;
                const long long t0 = clock64();
                while (final_count[g ^ 1] < need) {
                    if (clock64() - t0 > umma::kWatchdogCycles) umma::mbar_watchdog_trip(2u);
                }
                __threadfence_block();
            }
            const int64_t tile_lo = (int64_t)t * kTileRows, tile_hi = min(tile_lo + kTileRows, nrows);
            const int64_t q_first = tile_lo / N;        // first ray whose last row lies in this tile
            const int64_t q_last = tile_hi / N - 1;     // last ray completed by the end of this tile
            for (int64_t qr = q_first + wi; qr <= q_last; qr += 4)
                composite_ray(p, sRing, p.ring_cap, qr * N, p.ray_offset + ray0 + qr, lane);
        }
        tr_tot = (unsigned long long)(CNB_TR_NOW() - tr_t0);
        if (q == 0) { CNB_TR_FLUSH(5 + 4 * g, tr_wacc); CNB_TR_FLUSH(6 + 4 * g, tr_epi); CNB_TR_FLUSH(7 + 4 * g, tr_enc); CNB_TR_FLUSH(8 + 4 * g, tr_tot); }
    }
    umma::tc_fence_before();
    if (kCluster > 1) umma::cluster_sync_all(); else __syncthreads();
    if (warp == 1) { if (CG == 2) umma::tmem_dealloc2(tmem, 512); else umma::tmem_dealloc(tmem, 512); }
}

// ===========================================================================
// K1, tensor-memory operand form (default).  What bounds the kernel above is the weight stream: with both
// operand tiles in shared memory only a 4-slot (64 KB) ring fits, and a slot's round trip (MMA done -> refill
// from L2 -> MMA) is ~1500 cycles (tests/probe_ts.cu: 3058 cycles per 128x256x256 layer with 4 slots, 2360 with
// 8).  Here the activations never touch shared memory:
//   TMEM  [0,128) D0, [128,256) D1 : fp32 accumulators of the two 128-column halves of a layer
//         [256,384) A_X, [384,512) A_Y : the bf16 operand (K = 256) of tiles X and Y (tcgen05.st, "TS" MMAs)
//   SMEM  a 9/10-slot weight ring; every stage is used by tile X and, one layer-time later, by tile Y before it
//         is released (half the L2 -> SM traffic, 2300 cycles of slack per refill); PE blocks of the next tiles.
// Tensor pipe order per layer: X.h0 -> D0, X.h1 -> D1, Y.h0 -> D0, Y.h1 -> D1.  The eight epilogue warps (two per
// TMEM lane quarter, 64 columns each) drain a half while the next one is computed: + bias, ReLU, bf16, and the
// packed result goes back to A_tile once the layer's MMAs have all read it.  Four more warps generate rays /
// samples / positional encodings two tiles ahead and composite finished rays, so tile boundaries cost nothing.
constexpr int kTsThreads = 512;      // warp 0 producer, 1 MMA, 2-3 idle, 4-11 epilogue, 12-15 rays + PE + compositing
constexpr int kTsMaxSlots = 10;
constexpr uint32_t kColA = 256;

struct TsBars {
    uint64_t full[kTsMaxSlots], empty[kTsMaxSlots];
    uint64_t pe_ready[2], xyz_free[2], dir_free[2];   // positional encodings of the tile in slot X / Y
    uint64_t a_ready[2];                              // A_tile rewritten for the next layer
    uint64_t d_full[2], d_free[2];                    // accumulator halves
    uint64_t samp_full[2], samp_free[2];              // per-sample (sigma, rgb) of a finished tile in the sample ring
    uint32_t tmem_slot, pad;
};

// 32 accumulator columns -> + bias, heads, [ReLU] bf16 pairs in out[OFF .. OFF + 16).  Pointers are pre-offset to
// the first of the 32 columns.  KIND: 0 hidden, 1 encoding_shape (+ sigma head, no ReLU), 2 rgb.0 (+ rgb head, no operand).
template <int KIND, int OFF>
__device__ __forceinline__ void ts_epilogue32(const uint32_t (&rr)[32], const float* __restrict__ bias,
                                              const float* __restrict__ w_sigma, const float* __restrict__ w_rgb2,
                                              HeadAcc& acc, uint32_t (&out)[32]) {
    constexpr bool RELU = (KIND != 1);
#pragma unroll
    for (int j8 = 0; j8 < 4; ++j8) {
        const int col = j8 * 8;
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + col));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + col + 4));
        uint64_t v[4];
        v[0] = fadd2(pk2(rr[col + 0], rr[col + 1]), pk2f(b0.x, b0.y));
        v[1] = fadd2(pk2(rr[col + 2], rr[col + 3]), pk2f(b0.z, b0.w));
        v[2] = fadd2(pk2(rr[col + 4], rr[col + 5]), pk2f(b1.x, b1.y));
        v[3] = fadd2(pk2(rr[col + 6], rr[col + 7]), pk2f(b1.z, b1.w));
        if (KIND == 1) {        // sigma head on the fp32 feature (reference src/model.py:45)
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(w_sigma + col));
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(w_sigma + col + 4));
            acc.sig2 = ffma2(v[0], pk2f(w0.x, w0.y), acc.sig2); acc.sig2 = ffma2(v[1], pk2f(w0.z, w0.w), acc.sig2);
            acc.sig2 = ffma2(v[2], pk2f(w1.x, w1.y), acc.sig2); acc.sig2 = ffma2(v[3], pk2f(w1.z, w1.w), acc.sig2);
        } else if (KIND == 2) { // rgb.2 on the fp32 hidden (reference src/model.py:52)
            uint64_t h[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { float lo, hi; unpk2(v[i], lo, hi); h[i] = pk2f(fmaxf(lo, 0.f), fmaxf(hi, 0.f)); }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float4 w0 = __ldg(reinterpret_cast<const float4*>(w_rgb2 + k * (kW / 2) + col));
                const float4 w1 = __ldg(reinterpret_cast<const float4*>(w_rgb2 + k * (kW / 2) + col + 4));
                uint64_t& a = (k == 0) ? acc.r2 : (k == 1 ? acc.g2 : acc.b2);
                a = ffma2(h[0], pk2f(w0.x, w0.y), a); a = ffma2(h[1], pk2f(w0.z, w0.w), a);
                a = ffma2(h[2], pk2f(w1.x, w1.y), a); a = ffma2(h[3], pk2f(w1.z, w1.w), a);
            }
        }
        if (KIND != 2) {
#pragma unroll
            for (int i = 0; i < 4; ++i) out[OFF + j8 * 4 + i] = cvt_bf16x2<RELU>(v[i]);
        }
    }
}
// One thread's 64 columns of an accumulator half: columns [c0, c0 + 64) of the layer.
template <int KIND>
__device__ __forceinline__ void ts_half(uint32_t taddr, int c0, const float* __restrict__ bias,
                                        const float* __restrict__ w_sigma, const float* __restrict__ w_rgb2, HeadAcc& acc,
                                        uint32_t (&P)[32]) {
    uint32_t ra[32], rb[32];
    umma::tmem_ld32(taddr, ra);
    umma::tmem_ld32(taddr + 32, rb);
    umma::tmem_ld_wait();
    ts_epilogue32<KIND, 0>(ra, bias + c0, w_sigma + c0, w_rgb2 + c0, acc, P);
    ts_epilogue32<KIND, 16>(rb, bias + c0 + 32, w_sigma + c0 + 32, w_rgb2 + c0 + 32, acc, P);
}
__device__ __forceinline__ void ts_store_operand(uint32_t taddr, const uint32_t (&P)[32]) {
    uint32_t lo[16], hi[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { lo[i] = P[i]; hi[i] = P[16 + i]; }
    umma::tmem_st16(taddr, lo);
    umma::tmem_st16(taddr + 16, hi);
}

__global__ void __launch_bounds__(kTsThreads, 1) k_render_fwd_ts(const __grid_constant__ FwdParams p, const int n_slots) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sW = smem;                                        // [n_slots][16 KB]
    uint8_t* sXyz = sW + (size_t)n_slots * kSlot;              // [2][16 KB]
    uint8_t* sDir = sXyz + 2 * kABlock;                        // [2][8 KB]
    float4* sHead = (float4*)(sDir + 2 * kDirBlock);           // [2][128] partial heads of the second column-half warp
    float4* sRing = sHead + 2 * kTileRows;                     // [ring_cap] per-sample (sigma, r, g, b)
    TsBars* B = (TsBars*)((uint8_t*)sRing + (size_t)p.ring_cap * 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nl = p.n_layers;

    // ---- this CTA's contiguous share of the work (same split as k_render_fwd) ----
    int64_t row0 = 0, nrows = 0, ray0 = 0;
    {
        const int64_t b = (int64_t)blockIdx.x;
        if (p.mode == 0) {
            const int64_t base = p.n_rays / gridDim.x, rem = p.n_rays % gridDim.x;
            ray0 = b * base + min(b, rem);
            const int64_t nr = base + (b < rem ? 1 : 0);
            row0 = ray0 * p.rs.N; nrows = nr * p.rs.N;
        } else {
            const int64_t tiles = (p.S + kTileRows - 1) / kTileRows;
            const int64_t base = tiles / gridDim.x, rem = tiles % gridDim.x;
            const int64_t t0 = b * base + min(b, rem);
            const int64_t nt = base + (b < rem ? 1 : 0);
            row0 = t0 * kTileRows;
            nrows = min(p.S - row0, nt * kTileRows);
            if (nrows < 0) nrows = 0;
        }
    }
    const int T = (int)((nrows + kTileRows - 1) / kTileRows);
    const int rounds = (T + 1) >> 1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kTsMaxSlots; ++i) { umma::mbar_init(&B->full[i], 1); umma::mbar_init(&B->empty[i], 1); }
        for (int g = 0; g < 2; ++g) {
            umma::mbar_init(&B->pe_ready[g], 4); umma::mbar_init(&B->xyz_free[g], 1); umma::mbar_init(&B->dir_free[g], 1);
            umma::mbar_init(&B->a_ready[g], 8);
            umma::mbar_init(&B->d_full[g], 1); umma::mbar_init(&B->d_free[g], 8);
            umma::mbar_init(&B->samp_full[g], 4); umma::mbar_init(&B->samp_free[g], 4);
        }
        umma::fence_mbar_init();
    }
    if (warp == 1) umma::tmem_alloc(&B->tmem_slot, 512);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = B->tmem_slot;

    if (warp == 0) {
        // ===== weight producer: each stage of a layer once per tile PAIR, half-major (h, chunk) =====
        int slot = 0; uint32_t ph = 0;
        for (int r = 0; r < rounds; ++r)
            for (int l = 0; l < nl; ++l) {
                const FwdLayer& L = p.layers[l];
                const uint8_t* src = p.packed + L.w_off;
                for (int h = 0; h < L.n_halves; ++h)
                    for (int c = 0; c < L.n_kchunks + L.has_dir; ++c) {
                        const bool dir = c == L.n_kchunks;
                        const uint32_t bytes = dir ? kSlot / 2 : kSlot;
                        const size_t img = dir ? (size_t)(L.n_kchunks * L.n_halves + h) : (size_t)(c * L.n_halves + h);
                        umma::mbar_wait(&B->empty[slot], ph ^ 1);
                        if (umma::elect_one()) {
                            umma::mbar_arrive_expect_tx(&B->full[slot], bytes);
                            umma::bulk_g2s(sW + slot * kSlot, src + img * kSlot, bytes, &B->full[slot]);
                        }
                        __syncwarp();
                        if (++slot == n_slots) { slot = 0; ph ^= 1; }
                    }
            }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        int slot = 0; uint32_t ph = 0;
        uint32_t kd0 = 0, kd1 = 0;       // uses of D0 / D1 so far
        uint32_t ua0 = 0, ua1 = 0;       // a_ready phases consumed per tile slot
        CNB_TR_DECL(tr_wa); CNB_TR_DECL(tr_wd); CNB_TR_DECL(tr_ww); CNB_TR_DECL(tr_tot);
        const long long tr_t0 = CNB_TR_NOW();
        const uint32_t id128 = umma::make_idesc(128, 128, 0, 0);
        const uint64_t dW = umma::make_sdesc(umma::smem_u32(sW), 16, 1024, umma::SWZ_128B);
        const uint64_t dWd = umma::make_sdesc(umma::smem_u32(sW), 16, 512, umma::SWZ_64B);
        for (int r = 0; r < rounds; ++r) {
            const int ntile = min(2, T - 2 * r);
            for (int l = 0; l < nl; ++l) {
                const FwdLayer& L = p.layers[l];
                const int slot0 = slot; const uint32_t ph0 = ph;
                for (int g = 0; g < ntile; ++g) {
                    slot = slot0; ph = ph0;                       // tile Y walks the same stages again
                    if (l == 0) CNB_TR(tr_wa, umma::mbar_wait(&B->pe_ready[g], (uint32_t)r & 1u));
                    else if (g == 0) { CNB_TR(tr_wa, umma::mbar_wait(&B->a_ready[0], ua0 & 1u)); ++ua0; }
                    else { CNB_TR(tr_wa, umma::mbar_wait(&B->a_ready[1], ua1 & 1u)); ++ua1; }
                    const uint64_t dX = umma::make_sdesc(umma::smem_u32(sXyz + g * kABlock), 16, 1024, umma::SWZ_128B);
                    const uint64_t dD = umma::make_sdesc(umma::smem_u32(sDir + g * kDirBlock), 16, 512, umma::SWZ_64B);
                    const uint32_t aT = tmem + kColA + (uint32_t)g * 128u;
                    const bool first = g == 0, last = g == ntile - 1;
                    for (int h = 0; h < L.n_halves; ++h) {
                        const uint32_t kd = h ? kd1 : kd0;
                        CNB_TR(tr_wd, umma::mbar_wait(&B->d_free[h], (kd & 1u) ^ 1u));
                        umma::tc_fence_after();
                        const uint32_t dT = tmem + (uint32_t)h * 128u;
                        for (int c = 0; c < L.n_kchunks; ++c) {
                            if (first) { CNB_TR(tr_ww, umma::mbar_wait(&B->full[slot], ph)); umma::tc_fence_after(); }
                            if (umma::elect_one()) {
                                const uint64_t db = dW + (uint64_t)((slot * kSlot) >> 4);
                                if (l == 0) {
#pragma unroll
                                    for (int ks = 0; ks < 4; ++ks) umma::mma_bf16(dT, dX + ks * 2, db + ks * 2, id128, (c | ks) ? 1u : 0u);
                                } else {
#pragma unroll
                                    for (int ks = 0; ks < 4; ++ks) umma::mma_bf16_ts(dT, aT + c * 32 + ks * 8, db + ks * 2, id128, (c | ks) ? 1u : 0u);
                                }
                                if (last) umma::mma_commit(&B->empty[slot]);
                            }
                            __syncwarp();
                            if (++slot == n_slots) { slot = 0; ph ^= 1; }
                        }
                        if (L.has_dir) {
                            if (first) { umma::mbar_wait(&B->full[slot], ph); umma::tc_fence_after(); }
                            if (umma::elect_one()) {
                                const uint64_t db = dWd + (uint64_t)((slot * kSlot) >> 4);
#pragma unroll
                                for (int ks = 0; ks < 2; ++ks) umma::mma_bf16(dT, dD + ks * 2, db + ks * 2, id128, 1u);
                                if (last) umma::mma_commit(&B->empty[slot]);
                            }
                            __syncwarp();
                            if (++slot == n_slots) { slot = 0; ph ^= 1; }
                        }
                        if (umma::elect_one()) umma::mma_commit(&B->d_full[h]);
                        __syncwarp();
                        if (h) ++kd1; else ++kd0;
                    }
                    if (umma::elect_one()) {
                        if (l == 0) umma::mma_commit(&B->xyz_free[g]);
                        if (L.has_dir) umma::mma_commit(&B->dir_free[g]);
                    }
                    __syncwarp();
                }
            }
        }
        tr_tot = (unsigned long long)(CNB_TR_NOW() - tr_t0);
        CNB_TR_FLUSH(2, tr_wa); CNB_TR_FLUSH(3, tr_ww); CNB_TR_FLUSH(4, tr_tot); CNB_TR_FLUSH(13, tr_wd);
    } else if (warp >= 4 && warp < 12) {
        // ===== epilogue warps =====
        const int q = warp & 3;                        // TMEM lane quarter
        const int ch = (warp - 4) >> 2;                // which 64 of a half's 128 columns
        const int row = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int N = p.rs.N;
        uint32_t kd0 = 0, kd1 = 0;
        uint64_t sigX = 0ull, sigY = 0ull;             // sigma-head partial of this thread's columns, per tile slot
        CNB_TR_DECL(tr_w0); CNB_TR_DECL(tr_w1); CNB_TR_DECL(tr_fin); CNB_TR_DECL(tr_tot);
        const long long tr_t0 = CNB_TR_NOW();
        for (int r = 0; r < rounds; ++r) {
            const int ntile = min(2, T - 2 * r);
            for (int l = 0; l < nl; ++l) {
                const FwdLayer& L = p.layers[l];
                for (int g = 0; g < ntile; ++g) {
                    const int t = 2 * r + g;
                    const int64_t lrow = (int64_t)t * kTileRows + row;
                    const bool valid = lrow < nrows;
                    const int64_t grow = row0 + lrow;
                    int64_t code = 0;
                    if (p.n_codes > 1) {
                        code = (p.ray_offset * N + (valid ? grow : row0)) / p.rows_per_code;
                        if (code >= p.n_codes) code = p.n_codes - 1;
                    }
                    const float* bias = L.folded >= 0 ? p.folded + ((size_t)code * p.n_folded + L.folded) * kW : L.bias;
                    HeadAcc hacc = {g ? sigY : sigX, 0ull, 0ull, 0ull, 0ull};
                    if (l == 0) hacc.sig2 = 0ull;
                    uint32_t P[32];
                    const uint32_t tD = tmem + lane_off + (uint32_t)ch * 64u;
                    const uint32_t tA = tmem + lane_off + kColA + (uint32_t)g * 128u + (uint32_t)ch * 32u;
                    // ---- half 0 ----
                    CNB_TR(tr_w0, umma::mbar_wait(&B->d_full[0], kd0 & 1u)); ++kd0;
                    umma::tc_fence_after();
                    if (L.kind == 0) ts_half<0>(tD, ch * 64, bias, p.w_sigma, p.w_rgb2, hacc, P);
                    else if (L.kind == 1) ts_half<1>(tD, ch * 64, bias, p.w_sigma, p.w_rgb2, hacc, P);
                    else ts_half<2>(tD, ch * 64, bias, p.w_sigma, p.w_rgb2, hacc, P);
                    umma::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) umma::mbar_arrive(&B->d_free[0]);
                    if (L.n_halves == 2) {
                        // ---- half 1: all MMAs of the layer have read A_tile, the operand may be rewritten ----
                        CNB_TR(tr_w1, umma::mbar_wait(&B->d_full[1], kd1 & 1u)); ++kd1;
                        umma::tc_fence_after();
                        ts_store_operand(tA, P);
                        if (L.kind == 0) ts_half<0>(tD + 128u, 128 + ch * 64, bias, p.w_sigma, p.w_rgb2, hacc, P);
                        else ts_half<1>(tD + 128u, 128 + ch * 64, bias, p.w_sigma, p.w_rgb2, hacc, P);
                        ts_store_operand(tA + 64u, P);
                        umma::tmem_st_wait();
                        umma::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) { umma::mbar_arrive(&B->d_free[1]); umma::mbar_arrive(&B->a_ready[g]); }
                        if (g) sigY = hacc.sig2; else sigX = hacc.sig2;
                        continue;
                    }
                    // ---- last layer (rgb.0, one half): heads -> per-sample outputs ----
                    const long long tr_f0 = CNB_TR_NOW();
                    float sig_p, cr, cg, cb;
                    { float a0, a1; unpk2(hacc.sig2, a0, a1); sig_p = a0 + a1; unpk2(hacc.r2, a0, a1); cr = a0 + a1;
                      unpk2(hacc.g2, a0, a1); cg = a0 + a1; unpk2(hacc.b2, a0, a1); cb = a0 + a1; }
                    if (ch == 1) sHead[g * kTileRows + row] = make_float4(sig_p, cr, cg, cb);
                    umma::named_bar_sync(1, 256);
                    if (ch == 0) {
                        const float4 o = sHead[g * kTileRows + row];
                        const float sigma = cnb_softplus(sig_p + o.x + __ldg(p.b_sigma));
                        cr += o.y + __ldg(p.b_rgb2 + 0); cg += o.z + __ldg(p.b_rgb2 + 1); cb += o.w + __ldg(p.b_rgb2 + 2);
                        if (p.mode == 1) {
                            if (valid) {
                                p.sigmas[grow] = sigma;
                                p.rgbs[grow * 3 + 0] = cr; p.rgbs[grow * 3 + 1] = cg; p.rgbs[grow * 3 + 2] = cb;
                            }
                        } else {
                            umma::mbar_wait(&B->samp_free[g], ((uint32_t)r & 1u) ^ 1u);   // tile t - 2 has been composited
                            if (valid) {
                                sRing[lrow % p.ring_cap] = make_float4(sigma, cr, cg, cb);
                                if (p.spill_sig) {
                                    p.spill_sig[grow] = sigma;
                                    p.spill_rgb[grow * 3 + 0] = cr; p.spill_rgb[grow * 3 + 1] = cg; p.spill_rgb[grow * 3 + 2] = cb;
                                }
                            }
                            __syncwarp();
                            if (lane == 0) umma::mbar_arrive(&B->samp_full[g]);
                        }
                    }
                    umma::named_bar_sync(1, 256);      // sHead[g] may be rewritten
                    tr_fin += (unsigned long long)(CNB_TR_NOW() - tr_f0);
                }
            }
        }
        tr_tot = (unsigned long long)(CNB_TR_NOW() - tr_t0);
        if (warp == 4) { CNB_TR_FLUSH(5, tr_w0); CNB_TR_FLUSH(6, tr_w1); CNB_TR_FLUSH(7, tr_fin); CNB_TR_FLUSH(8, tr_tot); }
    } else if (warp >= 12) {
        // ===== ray / sample generation + positional encoding (two tiles ahead), compositing of finished tiles =====
        const int iw = warp - 12;
        const int row = iw * 32 + lane;
        const int N = p.rs.N;
        CNB_TR_DECL(tr_wx); CNB_TR_DECL(tr_ws); CNB_TR_DECL(tr_tot);
        const long long tr_t0 = CNB_TR_NOW();
        for (int i = 0; i < T + 2; ++i) {
            if (i < T) {
                const int g = i & 1; const uint32_t u = (uint32_t)(i >> 1);
                const int64_t lrow = (int64_t)i * kTileRows + row;
                const bool valid = lrow < nrows;
                const int64_t grow = row0 + lrow;
                float pos[3] = {0.f, 0.f, 0.f}, dir[3] = {0.f, 0.f, 0.f};
                if (valid) {
                    if (p.mode == 0) {
                        const int64_t lray = grow / N;
                        const int zi = (int)(grow - lray * N);
                        const int64_t ray = p.ray_offset + lray;
                        float o[3];
                        cnb_fetch_ray(p.rs, ray, o, dir);
                        const int64_t seg = ray / p.rs.rays_per_segment;
                        const float z = __ldg(p.rs.z_vals + (p.rs.z_per_segment ? seg * N : 0) + zi);
#pragma unroll
                        for (int k = 0; k < 3; ++k) pos[k] = cnb_sample_coord(o[k], dir[k], z);
                    } else {
#pragma unroll
                        for (int k = 0; k < 3; ++k) { pos[k] = __ldg(p.xyz + grow * 3 + k); dir[k] = __ldg(p.viewdir + grow * 3 + k); }
                    }
                }
                CNB_TR(tr_wx, umma::mbar_wait(&B->xyz_free[g], (u & 1u) ^ 1u));       // layer 0 of the previous tile in this slot is done
                encode_xyz_row(pos, valid, sXyz + g * kABlock, row);
                CNB_TR(tr_wx, umma::mbar_wait(&B->dir_free[g], (u & 1u) ^ 1u));       // ... and its PE(viewdir) layer
                encode_dir_row(dir, valid, sDir + g * kDirBlock, row);
                umma::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(&B->pe_ready[g]);
            }
            if (p.mode == 0 && i >= 2) {
                const int t = i - 2, g = t & 1; const uint32_t u = (uint32_t)(t >> 1);
                CNB_TR(tr_ws, umma::mbar_wait(&B->samp_full[g], u & 1u));
                const int64_t tile_lo = (int64_t)t * kTileRows, tile_hi = min(tile_lo + kTileRows, nrows);
                const int64_t q_first = tile_lo / N;        // first ray whose last row lies in this tile
                const int64_t q_last = tile_hi / N - 1;     // last ray completed by the end of this tile
                for (int64_t qr = q_first + iw; qr <= q_last; qr += 4)
                    composite_ray(p, sRing, p.ring_cap, qr * N, p.ray_offset + ray0 + qr, lane);
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(&B->samp_free[g]);
            }
        }
        tr_tot = (unsigned long long)(CNB_TR_NOW() - tr_t0);
        if (warp == 12) { CNB_TR_FLUSH(9, tr_wx); CNB_TR_FLUSH(10, tr_ws); CNB_TR_FLUSH(11, tr_tot); }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 1) umma::tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------
// Weight packing: fp32 parameter matrices -> bf16 stage images in the exact shared-memory
// byte layout the UMMA descriptors expect (so a stage is one linear bulk copy).
struct PackArgs {
    const float* W; int ld;          // source matrix, row stride
    int n_total, k_total;            // valid extents of the operand's N (rows) and K (cols)
    int transpose;                   // 0: B[n][k] = W[n*ld + k];  1: B[n][k] = W[k*ld + n]  (dgrad)
    int n_kchunks, n_halves, has_dir;
    int dir_k0, dir_width;           // K range of the PE(viewdir) columns inside W
    uint8_t* dst;
};

// All layers in ONE launch: the weights are repacked after every optimiser step, and a training iteration on a single
// view is launch bound (17 launches here were ~5 % of it).  Block b packs stage (b - first[layer]) of its layer.
struct PackTable { int n; int first[2 * kMaxLayers + 1]; PackArgs a[2 * kMaxLayers]; };

__global__ void k_pack_layer(const __grid_constant__ PackTable t) {
    int li = 0;
    while (li + 1 < t.n && (int)blockIdx.x >= t.first[li + 1]) ++li;
    const PackArgs& a = t.a[li];
    const int s = (int)blockIdx.x - t.first[li];
    const int n_main = a.n_kchunks * a.n_halves;
    uint8_t* out = a.dst + (size_t)s * kSlot;
    if (s < n_main) {
        const int c = s / a.n_halves, h = s % a.n_halves;
        for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {
            const int rown = i >> 3, ch = i & 7;
            const int n = h * 128 + rown;
            uint32_t w[4];
#pragma unroll
            for (int e2 = 0; e2 < 4; ++e2) {
                float f[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int k = c * 64 + ch * 8 + e2 * 2 + e;
                    f[e] = (n < a.n_total && k < a.k_total)
                               ? (a.transpose ? a.W[(size_t)k * a.ld + n] : a.W[(size_t)n * a.ld + k]) : 0.f;
                }
                w[e2] = umma::pack_bf16(f[0], f[1]);
            }
            *reinterpret_cast<uint4*>(out + rown * 128 + ((ch ^ (rown & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    } else {
        const int h = s - n_main;
        for (int i = threadIdx.x; i < 128 * 4; i += blockDim.x) {
            const int rown = i >> 2, ch = i & 3;
            const int n = h * 128 + rown;
            uint32_t w[4];
#pragma unroll
            for (int e2 = 0; e2 < 4; ++e2) {
                float f[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int j = ch * 8 + e2 * 2 + e;
                    f[e] = (n < a.n_total && j < a.dir_width) ? a.W[(size_t)n * a.ld + a.dir_k0 + j] : 0.f;
                }
                w[e2] = umma::pack_bf16(f[0], f[1]);
            }
            *reinterpret_cast<uint4*>(out + rown * 64 + ((ch ^ ((rown >> 1) & 3)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

// folded[code][j][n] = b_j[n] + sum_k W_j[n][k] z_j[code][k]   (the y + z add of src/model.py:42,50
// moved across the linear layer: (y + z) W^T + b = y W^T + (W z + b)).  One warp per output.
__global__ void k_fold_bias(const float* __restrict__ Wj, const float* __restrict__ bj, const float* __restrict__ z,
                            int64_t ldz, int n_codes, float* __restrict__ out, int64_t ldo) {
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= n_codes * kW) return;
    const int o = gw / kW, n = gw - o * kW;
    float acc = 0.f;
    for (int i = lane; i < kW; i += 32) acc = fmaf(__ldg(Wj + (size_t)n * kW + i), z[(size_t)o * ldz + i], acc);
    acc = warp_sum_f(acc);
    if (lane == 0) out[(size_t)o * ldo + n] = acc + __ldg(bj + n);
}

// ---------------------------------------------------------------------------
}  // namespace

namespace sm100 {
// z_j = ReLU(latent layer_j(code)) and the folded per-code biases.
int latent_and_fold(const cnb_net_config* c, const float* const* P, const float* shape_codes, const float* tex_codes,
                    int n_codes, FwdWorkspace& w, cudaStream_t st) {
    CnbLayout L; cnb_make_layout(c, &L);
    const int nf = c->shape_blocks + c->texture_blocks;
    const int64_t ld = (int64_t)nf * kW;
    const int threads = 256, blocks = (n_codes * kW * 32 + threads - 1) / threads;
    for (int j = 0; j < nf; ++j) {
        const bool shape = j < c->shape_blocks;
        const int jj = shape ? j : j - c->shape_blocks;
        const int il = shape ? L.i_sl[jj] : L.i_tl[jj];
        const int iw = shape ? L.i_s[jj] : L.i_t[jj];
        CNB_TRY(cnb_launch_latent_fwd(P[il], P[il + 1], shape ? shape_codes : tex_codes, n_codes, c->latent_dim, kW,
                                      w.z + (size_t)j * kW, ld, st));
        k_fold_bias<<<blocks, threads, 0, st>>>(P[iw], P[iw + 1], w.z + (size_t)j * kW, ld, n_codes,
                                               w.folded + (size_t)j * kW, ld);
        CNB_LAUNCH_CHECK();
    }
    return CNB_OK;
}


}  // namespace sm100

namespace {

int launch_fwd(const cnb_net_config* c, const float* const* P, const void* packed, const Plan& pl, FwdParams& fp,
               int64_t total_rows, float* rows_staging, cudaStream_t st) {
    CnbLayout L; cnb_make_layout(c, &L);
#if CNB_K1_CONST
    ConstRowsScope crs;                    // launch-wide bias rows and head weights -> constant memory (until the launch below)
    CNB_TRY(crs.begin(c, P, pl, rows_staging, st));
#else
    (void)rows_staging;
#endif
    fp.n_layers = pl.n_layers;
    for (int i = 0; i < pl.n_layers; ++i) fp.layers[i] = pl.fwd[i];
    fp.packed = (const uint8_t*)packed;
    fp.n_folded = pl.n_folded;
    fp.w_sigma = P[L.i_sigma]; fp.b_sigma = P[L.i_sigma + 1];
    fp.w_rgb2 = P[L.i_rgb2]; fp.b_rgb2 = P[L.i_rgb2 + 1];
    const int N = fp.mode == 0 ? fp.rs.N : 1;
    fp.ring_cap = fp.mode == 0 ? N + 384 : 16;
    size_t smem = 1024 + 2 * (size_t)kATile + (size_t)kNumStages * kSlot + (size_t)fp.ring_cap * 16 + 256;
    if (smem > 232448) return CNB_E_UNSUPPORTED;
    const size_t staging = sizeof(float) * (4 * kW + kW + 3 * (kW / 2));
    fp.stage_bias = smem + staging <= 232448 ? 1 : 0;       // very long rays (N > ~300) leave no room: global bias loads
    if (fp.stage_bias) smem += staging;
    int dev = 0, sms = 0;
    CNB_CUDA_TRY(cudaGetDevice(&dev));
    CNB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int64_t units = (total_rows + 2 * kTileRows - 1) / (2 * kTileRows);
    if (fp.mode == 0 && units > fp.n_rays) units = fp.n_rays;
    int grid = (int)(units < sms ? (units < 1 ? 1 : units) : sms);
    // CTA pairs by default (round 2, with the light-weight remote arrives of umma.cuh): +8-13 % in a same-session A/B
    const int use_pairs = (int)cnb_option("cta_pairs", 1);
    fp.share_fills = cnb_option("share_fills", 1) != 0 ? 1 : 0;
    const char* form = getenv("CNB_FWD_KERNEL");       // "ts": the tensor-memory operand experiment (slower)
    const bool use_ts = cnb_option("fwd_kernel_ts", 0) != 0 || (form && form[0] == 't');
    if (!use_pairs && use_ts) {
        int n_slots = kTsMaxSlots;
        auto need = [&](int slots) {
            return 1024 + (size_t)slots * kSlot + 2 * (size_t)kABlock + 2 * (size_t)kDirBlock + 2 * kTileRows * sizeof(float4) +
                   (size_t)fp.ring_cap * 16 + sizeof(TsBars) + 64;
        };
        while (n_slots > 8 && need(n_slots) > 232448) --n_slots;
        if (need(n_slots) > 232448) return CNB_E_UNSUPPORTED;
        const size_t tsmem = need(n_slots);
        CNB_CUDA_TRY(cudaFuncSetAttribute(k_render_fwd_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
        cnb_prof_begin(CNB_K_FWD, st);
        k_render_fwd_ts<<<grid, kTsThreads, tsmem, st>>>(fp, n_slots);
        cnb_prof_end(CNB_K_FWD, st);
        CNB_LAUNCH_CHECK();
        return CNB_OK;
    }
    const int mc = (grid == sms && !use_pairs) ? weight_multicast() : 1;
    const bool pairs = use_pairs && grid >= 2;
    if (pairs) { grid &= ~1; CNB_TRY(make_weight_maps(packed, pl.total_bytes, &fp.maps)); }
    const int epi8 = cnb_option("epi_warps", 4) == 8 ? 1 : 0;
    const bool ew8 = epi8 && fp.stage_bias;       // the head exchange lives in the staging area
    void (*kern)(const FwdParams) = pairs ? (ew8 ? k_render_fwd<2, 1, 8> : k_render_fwd<2, 1, 4>)
                                    : mc == 4 ? k_render_fwd<1, 4> : mc == 2 ? k_render_fwd<1, 2>
                                    : (ew8 ? k_render_fwd<1, 1, 8> : k_render_fwd<1, 1, 4>);
    const int csize = pairs ? 2 : mc;
    CNB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = csize > 1 ? 1 : 0;
    if (mc > 1) CNB_TRY(cluster_grid(kern, &cfg, mc, &grid));
    cfg.gridDim = dim3(grid);
    cnb_prof_begin(CNB_K_FWD, st);
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, fp);
    cnb_prof_end(CNB_K_FWD, st);
    if (e != cudaSuccess) return (int)e;
    CNB_LAUNCH_CHECK();
    return CNB_OK;
}

}  // namespace

// ===========================================================================
size_t cnb_sm100_packed_bytes(const cnb_net_config* cfg) {
    Plan pl;
    if (make_plan(cfg, nullptr, &pl) != CNB_OK) return 256;   // fp32-only network shapes need no packed copy
    return pl.total_bytes + 1024;
}

int cnb_sm100_pack_weights(const cnb_net_config* cfg, const float* const* P, void* packed, cudaStream_t st) {
    Plan pl;
    const int rc = make_plan(cfg, P, &pl);
    if (rc == CNB_E_UNSUPPORTED) return CNB_OK;      // nothing to pack; the bf16 entry points will refuse
    CNB_TRY(rc);
    if (((uintptr_t)packed & 127) != 0) return CNB_E_ALIGNMENT;
    CnbLayout L; cnb_make_layout(cfg, &L);
    PackTable tab = {};
    int total = 0;
    auto push = [&](const PackArgs& a, int n_stages) {
        tab.first[tab.n] = total; tab.a[tab.n] = a; ++tab.n; total += n_stages; tab.first[tab.n] = total;
    };
    int li = 0;
    auto pack = [&](int wi, int ld, int n_total, int k_total, int dir_k0, int dir_w) {
        const FwdLayer& f = pl.fwd[li++];
        PackArgs a = {};
        a.W = P[wi]; a.ld = ld; a.n_total = n_total; a.k_total = k_total; a.transpose = 0;
        a.n_kchunks = f.n_kchunks; a.n_halves = f.n_halves; a.has_dir = f.has_dir; a.dir_k0 = dir_k0; a.dir_width = dir_w;
        a.dst = (uint8_t*)packed + f.w_off;
        push(a, (f.n_kchunks + f.has_dir) * f.n_halves);
    };
    pack(L.i_enc_xyz, L.d_xyz, kW, L.d_xyz, 0, 0);
    for (int j = 0; j < cfg->shape_blocks; ++j) pack(L.i_s[j], kW, kW, kW, 0, 0);
    pack(L.i_enc_shape, kW, kW, kW, 0, 0);
    pack(L.i_enc_vd, kW + L.d_dir, kW, kW, kW, L.d_dir);
    for (int j = 0; j < cfg->texture_blocks; ++j) pack(L.i_t[j], kW, kW, kW, 0, 0);
    pack(L.i_rgb0, kW, kW / 2, kW, 0, 0);
    // dgrad operands: B[n = k_in][k = n_out] = W[k][n], K chunks over n_out, two N halves over the first 256 inputs
    for (int l = 1; l < pl.n_layers; ++l) {
        const FwdLayer& f = pl.fwd[l];
        PackArgs a = {};
        a.W = P[pl.w_index[l]];
        a.ld = f.has_dir ? kW + L.d_dir : kW;
        a.n_total = kW; a.k_total = f.n_halves * 128; a.transpose = 1;
        a.n_kchunks = f.n_halves * 2; a.n_halves = 2; a.has_dir = 0;
        a.dst = (uint8_t*)packed + pl.bwd_w_off[l];
        push(a, a.n_kchunks * 2);
    }
    k_pack_layer<<<total, 256, 0, st>>>(tab);
    CNB_LAUNCH_CHECK();
    return CNB_OK;
}

size_t cnb_sm100_mlp_workspace_bytes(const cnb_net_config* cfg, int64_t S, int n_codes, int backward) {
    if (backward) return bwd_workspace_bytes(cfg, S, 0, 1, n_codes, 0);
    return carve_fwd(cfg, n_codes, 0, nullptr, nullptr) + 256;
}

size_t cnb_sm100_render_workspace_bytes(const cnb_net_config* cfg, const cnb_ray_batch* rays, int backward) {
    if (backward) return bwd_workspace_bytes(cfg, rays->n_rays * rays->n_samples, rays->n_rays, rays->n_samples, rays->n_codes, 1);
    return carve_fwd(cfg, rays->n_codes, 0, nullptr, nullptr) + 256;
}

namespace sm100 {
int make_weight_maps(const void* packed, size_t bytes, WeightMaps* out) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = [] {
        void* f = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) f = nullptr;
        return (EncodeFn)f;
    }();
    if (!fn) return CNB_E_DEVICE;
    const cuuint64_t dims[2] = {64, (cuuint64_t)(bytes / 128)};     // rows of 64 bf16 = 128 bytes
    const cuuint64_t strides[1] = {128};
    const cuuint32_t estr[2] = {1, 1};
    const cuuint32_t box16[2] = {64, 128}, box8[2] = {64, 64};
    if (fn(&out->m16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(packed), dims, strides, box16, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return CNB_E_INVALID;
    if (fn(&out->m8, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(packed), dims, strides, box8, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return CNB_E_INVALID;
    return CNB_OK;
}
// K1 over rays [ray_begin, ray_begin + ray_count) of the batch; outputs indexed by global ray,
// the optional per-sample spill by launch-relative row.
int launch_render_rays(const cnb_net_config* cfg, const float* const* P, const void* packed, const Plan& pl,
                       const cnb_ray_batch* rays, int64_t ray_begin, int64_t ray_count, const float* folded,
                       float* rows_staging, float* spill_sig, float* spill_rgb, float* rgb, float* depth, float* acc,
                       cudaStream_t st) {
    const int N = rays->n_samples;
    const int64_t rows_per_code = (int64_t)rays->segments_per_code * rays->rays_per_segment * N;
    FwdParams fp = {};
    fp.mode = 0; fp.rs = cnb_make_ray_source(rays); fp.n_rays = ray_count; fp.S = ray_count * N; fp.ray_offset = ray_begin;
    fp.folded = folded; fp.n_codes = rays->n_codes;
    fp.rows_per_code = rays->n_codes > 1 ? rows_per_code : rays->n_rays * N;
    fp.white_bg = rays->white_bg;
    fp.rgb = rgb; fp.depth = depth; fp.acc = acc; fp.spill_sig = spill_sig; fp.spill_rgb = spill_rgb;
    return launch_fwd(cfg, P, packed, pl, fp, fp.S, rows_staging, st);
}
}  // namespace sm100

int cnb_sm100_mlp_forward(const cnb_net_config* cfg, const float* const* P, const void* packed, const float* xyz,
                          const float* viewdir, const float* shape_codes, const float* tex_codes, int n_codes,
                          int64_t samples_per_code, int64_t S, float* sigmas, float* rgbs, void* ws, size_t ws_bytes,
                          cudaStream_t st) {
    Plan pl;
    CNB_TRY(make_plan(cfg, P, &pl));
    if (n_codes > 1 && (samples_per_code % 32) != 0) return CNB_E_UNSUPPORTED;
    FwdWorkspace w;
    if (!ws || ws_bytes < carve_fwd(cfg, n_codes, 0, nullptr, nullptr)) return CNB_E_WORKSPACE;
    if (((uintptr_t)ws & 255) != 0) return CNB_E_ALIGNMENT;
    carve_fwd(cfg, n_codes, 0, (char*)ws, &w);
    CNB_TRY(latent_and_fold(cfg, P, shape_codes, tex_codes, n_codes, w, st));
    FwdParams fp = {};
    fp.mode = 1; fp.xyz = xyz; fp.viewdir = viewdir; fp.S = S; fp.n_rays = 0;
    fp.folded = w.folded; fp.n_codes = n_codes; fp.rows_per_code = samples_per_code > 0 ? samples_per_code : S;
    fp.sigmas = sigmas; fp.rgbs = rgbs;
    fp.rs.N = 1;
    return launch_fwd(cfg, P, packed, pl, fp, S, w.rows, st);
}

int cnb_sm100_render(const cnb_net_config* cfg, const float* const* P, const void* packed, const cnb_ray_batch* rays,
                     int mode, const float* d_rgb, const float* d_depth, const float* target, float loss_scale,
                     float* rgb, float* depth, float* acc, float* sq_err, float* d_params, float* d_shape,
                     float* d_tex, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (mode != 0)
        return render_backward(cfg, P, packed, rays, mode, d_rgb, d_depth, target, loss_scale, rgb, depth, acc, sq_err,
                               d_params, d_shape, d_tex, ws, ws_bytes, st);
    Plan pl;
    CNB_TRY(make_plan(cfg, P, &pl));
    const int N = rays->n_samples;
    if (N > 512) return CNB_E_UNSUPPORTED;
    const int64_t rows_per_code = (int64_t)rays->segments_per_code * rays->rays_per_segment * N;
    if (rays->n_codes > 1 && (rows_per_code % 32) != 0) return CNB_E_UNSUPPORTED;
    FwdWorkspace w;
    if (!ws || ws_bytes < carve_fwd(cfg, rays->n_codes, 0, nullptr, nullptr)) return CNB_E_WORKSPACE;
    if (((uintptr_t)ws & 255) != 0) return CNB_E_ALIGNMENT;
    carve_fwd(cfg, rays->n_codes, 0, (char*)ws, &w);
    CNB_TRY(latent_and_fold(cfg, P, rays->shape_codes, rays->texture_codes, rays->n_codes, w, st));
    return launch_render_rays(cfg, P, packed, pl, rays, 0, rays->n_rays, w.folded, w.rows, nullptr, nullptr, rgb, depth, acc, st);
}

int cnb_sm100_pipeline_timeouts(void) {
    unsigned int v = 0;
    if (cudaMemcpyFromSymbol(&v, umma::g_umma_timeout, sizeof(unsigned int)) != cudaSuccess) return -1;
    const int b = cnb_sm100_pipeline_timeouts_bwd();
    return b < 0 ? b : (int)v | b;
}
