// probe_ts.cu -- A operand in tensor memory ("TS" tcgen05.mma) vs. A in shared memory ("SS").
//   check : one CTA computes D = A[128 x 256] . W[256 x 256]^T both ways and against a CPU sum; the TMEM image of A is
//           written with tcgen05.st (row m -> lane m, elements 2c, 2c+1 of the row packed in 32-bit column c).
//   perf  : 148 persistent CTAs stream weight stages through a shared-memory ring (as the render kernels do) and issue
//           one layer's MMAs per 8 stages; cycles per layer for SS (N = 256, A re-read from shared memory) and TS
//           (N = 128 halves, A from tensor memory), ring depth 4 or 12.  No epilogue: this isolates what the
//           shared-memory port (operand reads + TMA fills) lets the tensor pipe reach.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <string>
#include <cuda_runtime.h>
#include "../codenerf_b200/csrc/sm100_common.cuh"

namespace {
constexpr int kSlot = 16384;

__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__host__ __device__ inline float a_val(int m, int k) { return (float)(((m * 37 + k * 11) % 17) - 8) * 0.125f; }
__host__ __device__ inline float w_val(int n, int k) { return (float)(((n * 13 + k * 7) % 19) - 9) * 0.0625f; }

// weights as 8 stage images in the order (chunk c, half h): [128 n x 64 k] bf16, 128-byte swizzle
__global__ void k_make_weights(uint8_t* dst) {
    const int s = blockIdx.x, c = s >> 1, h = s & 1;
    for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) {
        const int r = i >> 6, col = i & 63;
        const __nv_bfloat16 v = __float2bfloat16(w_val(h * 128 + r, c * 64 + col));
        *reinterpret_cast<__nv_bfloat16*>(dst + (size_t)s * kSlot + umma::sw128_offset(r, col)) = v;
    }
}

// ---- correctness ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(160, 1) k_check(const uint8_t* wimg, float* out_ss, float* out_ts) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                    // 4 K-blocks [128 x 64]
    uint8_t* sW = smem + 4 * kSlot;        // 8 stages
    __shared__ uint64_t bar_w, bar_d;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { umma::mbar_init(&bar_w, 1); umma::mbar_init(&bar_d, 1); umma::fence_mbar_init(); }
    if (warp == 4) umma::tmem_alloc(&tmem_slot, 512);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t colA = 256;             // TMEM columns [256, 384): A as packed bf16 pairs
    if (warp < 4) {
        const int row = warp * 32 + lane;
        for (int k = 0; k < 256; ++k)
            *reinterpret_cast<__nv_bfloat16*>(sA + (k >> 6) * kSlot + umma::sw128_offset(row, k & 63)) = __float2bfloat16(a_val(row, k));
        for (int cc = 0; cc < 4; ++cc) {       // 32 columns = 64 elements per store
            uint32_t r[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = umma::pack_bf16(a_val(row, cc * 64 + 2 * j), a_val(row, cc * 64 + 2 * j + 1));
            tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + colA + cc * 32, r);
        }
        tmem_st_wait();
        umma::tc_fence_before();
        umma::fence_proxy_async_smem();
    }
    __syncthreads();
    umma::tc_fence_after();
    if (warp == 4) {
        if (umma::elect_one()) {
            umma::mbar_arrive_expect_tx(&bar_w, 8 * kSlot);
            for (int s = 0; s < 8; ++s) umma::bulk_g2s(sW + s * kSlot, wimg + (size_t)s * kSlot, kSlot, &bar_w);
        }
        __syncwarp();
        umma::mbar_wait(&bar_w, 0);
        umma::tc_fence_after();
        if (umma::elect_one()) {
            const uint64_t dA = umma::make_sdesc(umma::smem_u32(sA), 16, 1024, umma::SWZ_128B);
            const uint64_t dB = umma::make_sdesc(umma::smem_u32(sW), 16, 1024, umma::SWZ_128B);
            const uint32_t idesc = umma::make_idesc(128, 256, 0, 0);
            for (int c = 0; c < 4; ++c)
                for (int ks = 0; ks < 4; ++ks)
                    umma::mma_bf16(tmem, dA + ((c * kSlot) >> 4) + ks * 2, dB + ((c * 2 * kSlot) >> 4) + ks * 2, idesc, (c | ks) ? 1u : 0u);
            umma::mma_commit(&bar_d);
        }
        __syncwarp();
    }
    umma::mbar_wait(&bar_d, 0);
    umma::tc_fence_after();
    if (warp < 4) {
        const int row = warp * 32 + lane;
        for (int cc = 0; cc < 8; ++cc) {
            uint32_t r[32];
            umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + cc * 32, r);
            umma::tmem_ld_wait();
            for (int j = 0; j < 32; ++j) out_ss[row * 256 + cc * 32 + j] = __uint_as_float(r[j]);
        }
        umma::tc_fence_before();
    }
    __syncthreads();
    umma::tc_fence_after();
    if (warp == 4) {
        if (umma::elect_one()) {
            const uint64_t dB = umma::make_sdesc(umma::smem_u32(sW), 16, 1024, umma::SWZ_128B);
            const uint32_t idesc = umma::make_idesc(128, 128, 0, 0);
            for (int h = 0; h < 2; ++h)
                for (int c = 0; c < 4; ++c)
                    for (int ks = 0; ks < 4; ++ks)
                        mma_bf16_ts(tmem + h * 128, tmem + colA + c * 32 + ks * 8, dB + (((c * 2 + h) * kSlot) >> 4) + ks * 2, idesc,
                                    (c | ks) ? 1u : 0u);
            umma::mma_commit(&bar_d);
        }
        __syncwarp();
    }
    umma::mbar_wait(&bar_d, 1);
    umma::tc_fence_after();
    if (warp < 4) {
        const int row = warp * 32 + lane;
        for (int cc = 0; cc < 8; ++cc) {
            uint32_t r[32];
            umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + cc * 32, r);
            umma::tmem_ld_wait();
            for (int j = 0; j < 32; ++j) out_ts[row * 256 + cc * 32 + j] = __uint_as_float(r[j]);
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 4) umma::tmem_dealloc(tmem, 512);
}

// ---- throughput -----------------------------------------------------------------------------------
// MODE 0: SS, N = 256 (a pair of stages per MMA group), A [128 x 256] in shared memory.  MODE 1: TS, N = 128 per stage.
// FILL 0: the ring is filled once and never refilled (operand reads only).  FILL 1: every stage is streamed from L2.
template <int MODE, int FILL>
__global__ void __launch_bounds__(128, 1) k_perf(const uint8_t* wimg, int n_slots, int layers, unsigned long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                                 // SS only: 64 KB (contents irrelevant)
    uint8_t* sW = smem + (MODE == 0 ? 4 * kSlot : 0);
    __shared__ uint64_t full[16], empty[16], done;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; ++i) { umma::mbar_init(&full[i], 1); umma::mbar_init(&empty[i], 1); }
        umma::mbar_init(&done, 1);
        umma::fence_mbar_init();
    }
    if (warp == 1) umma::tmem_alloc(&tmem_slot, 512);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    {   // non-trivial operand data (tensor-core power depends on it)
        const int row = threadIdx.x;
        if (MODE == 0) {
            for (int k = 0; k < 256; ++k)
                *reinterpret_cast<__nv_bfloat16*>(sA + (k >> 6) * kSlot + umma::sw128_offset(row, k & 63)) = __float2bfloat16(a_val(row, k));
            umma::fence_proxy_async_smem();
        } else {
            for (int cc = 0; cc < 4; ++cc) {
                uint32_t r[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = umma::pack_bf16(a_val(row, cc * 64 + 2 * j), a_val(row, cc * 64 + 2 * j + 1));
                tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + 256 + cc * 32, r);
            }
            tmem_st_wait();
        }
        umma::tc_fence_before();
        __syncthreads();
        umma::tc_fence_after();
    }
    const long long t0 = clock64();
    const int total = layers * 8;
    if (warp == 0) {
        int stage = 0; uint32_t ph = 0;
        const int n_fill = FILL ? total : n_slots;
        for (int s = 0; s < n_fill; ++s) {
            umma::mbar_wait(&empty[stage], ph ^ 1);
            if (umma::elect_one()) {
                umma::mbar_arrive_expect_tx(&full[stage], kSlot);
                umma::bulk_g2s(sW + stage * kSlot, wimg + (size_t)(s & 7) * kSlot, kSlot, &full[stage]);
            }
            __syncwarp();
            if (++stage == n_slots) { stage = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        int stage = 0; uint32_t ph = 0;
        const uint64_t dA = umma::make_sdesc(umma::smem_u32(sA), 16, 1024, umma::SWZ_128B);
        const uint64_t dB = umma::make_sdesc(umma::smem_u32(sW), 16, 1024, umma::SWZ_128B);
        const uint32_t id256 = umma::make_idesc(128, 256, 0, 0), id128 = umma::make_idesc(128, 128, 0, 0);
        for (int s = 0; s < total; s += (MODE == 0 ? 2 : 1)) {
            const int c = (s >> 1) & 3, h = s & 1;
            if (FILL || s < n_slots) {
                umma::mbar_wait(&full[stage], ph);
                if (MODE == 0) umma::mbar_wait(&full[stage + 1], ph);
            }
            umma::tc_fence_after();
            if (umma::elect_one()) {
                const uint64_t db = dB + (uint64_t)((stage * kSlot) >> 4);
                if (MODE == 0) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) umma::mma_bf16(tmem, dA + ((c * kSlot) >> 4) + ks * 2, db + ks * 2, id256, (c | ks) ? 1u : 0u);
                    if (FILL) { umma::mma_commit(&empty[stage]); umma::mma_commit(&empty[stage + 1]); }
                } else {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) mma_bf16_ts(tmem + h * 128, tmem + 256 + c * 32 + ks * 8, db + ks * 2, id128, (c | ks) ? 1u : 0u);
                    if (FILL) umma::mma_commit(&empty[stage]);
                }
            }
            __syncwarp();
            stage += (MODE == 0 ? 2 : 1);
            if (stage >= n_slots) { stage = 0; ph ^= 1; }
        }
        if (umma::elect_one()) umma::mma_commit(&done);
        __syncwarp();
        umma::mbar_wait(&done, 0);
        if (threadIdx.x == 32) atomicAdd(cycles, (unsigned long long)(clock64() - t0));
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 1) umma::tmem_dealloc(tmem, 512);
}

// The proposed schedule: tiles X and Y keep their A operands in tensor memory; each layer's 8 stages are consumed by
// X (h-major: 4 stages -> D0, 4 stages -> D1) and again by Y one layer-time later, then released.  TSMODE 1: A from
// TMEM (N = 128); TSMODE 0: A from shared memory (N = 128), to separate the TS cost from the N = 128 cost.
template <int TSMODE>
__global__ void __launch_bounds__(128, 1) k_perf_shared(const uint8_t* wimg, int n_slots, int layers, unsigned long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                                 // TSMODE 0 only
    uint8_t* sW = smem + (TSMODE == 0 ? 4 * kSlot : 0);
    __shared__ uint64_t full[16], empty[16], done;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; ++i) { umma::mbar_init(&full[i], 1); umma::mbar_init(&empty[i], 1); }
        umma::mbar_init(&done, 1);
        umma::fence_mbar_init();
    }
    if (warp == 1) umma::tmem_alloc(&tmem_slot, 512);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    {
        const int row = threadIdx.x;
        if (TSMODE == 0) {
            for (int k = 0; k < 256; ++k)
                *reinterpret_cast<__nv_bfloat16*>(sA + (k >> 6) * kSlot + umma::sw128_offset(row, k & 63)) = __float2bfloat16(a_val(row, k));
            umma::fence_proxy_async_smem();
        } else {
            for (int t = 0; t < 2; ++t)
                for (int cc = 0; cc < 4; ++cc) {
                    uint32_t r[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = umma::pack_bf16(a_val(row + t, cc * 64 + 2 * j), a_val(row + t, cc * 64 + 2 * j + 1));
                    tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + 256 + t * 128 + cc * 32, r);
                }
            tmem_st_wait();
        }
        umma::tc_fence_before();
        __syncthreads();
        umma::tc_fence_after();
    }
    const long long t0 = clock64();
    if (warp == 0) {
        int stage = 0; uint32_t ph = 0;
        for (int s = 0; s < layers * 8; ++s) {
            umma::mbar_wait(&empty[stage], ph ^ 1);
            if (umma::elect_one()) {
                umma::mbar_arrive_expect_tx(&full[stage], kSlot);
                umma::bulk_g2s(sW + stage * kSlot, wimg + (size_t)(s & 7) * kSlot, kSlot, &full[stage]);
            }
            __syncwarp();
            if (++stage == n_slots) { stage = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        const uint64_t dA = umma::make_sdesc(umma::smem_u32(sA), 16, 1024, umma::SWZ_128B);
        const uint64_t dB = umma::make_sdesc(umma::smem_u32(sW), 16, 1024, umma::SWZ_128B);
        const uint32_t id128 = umma::make_idesc(128, 128, 0, 0);
        int base = 0; uint32_t bph = 0;             // ring position / phase of the layer's first stage
        for (int l = 0; l < layers; ++l) {
            for (int t = 0; t < 2; ++t)
                for (int h = 0; h < 2; ++h) {
                    // the 4 stages of this half: ring slots base + h*4 + c (mod n_slots)
                    if (t == 0) {
                        for (int c = 0; c < 4; ++c) {
                            int sl = base + h * 4 + c; uint32_t p = bph;
                            if (sl >= n_slots) { sl -= n_slots; p ^= 1; }
                            umma::mbar_wait(&full[sl], p);
                        }
                        umma::tc_fence_after();
                    }
                    if (umma::elect_one()) {
                        for (int c = 0; c < 4; ++c) {
                            int sl = base + h * 4 + c;
                            if (sl >= n_slots) sl -= n_slots;
                            const uint64_t db = dB + (uint64_t)((sl * kSlot) >> 4);
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                if (TSMODE) mma_bf16_ts(tmem + h * 128, tmem + 256 + t * 128 + c * 32 + ks * 8, db + ks * 2, id128, (c | ks) ? 1u : 0u);
                                else umma::mma_bf16(tmem + h * 128, dA + ((c * kSlot) >> 4) + ks * 2, db + ks * 2, id128, (c | ks) ? 1u : 0u);
                            }
                            if (t == 1) umma::mma_commit(&empty[sl]);
                        }
                    }
                    __syncwarp();
                }
            base += 8;
            if (base >= n_slots) { base -= n_slots; bph ^= 1; }
        }
        if (umma::elect_one()) umma::mma_commit(&done);
        __syncwarp();
        umma::mbar_wait(&done, 0);
        if (threadIdx.x == 32) atomicAdd(cycles, (unsigned long long)(clock64() - t0));
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 1) umma::tmem_dealloc(tmem, 512);
}

template <int TSMODE>
void run_shared(const uint8_t* wimg, int n_slots, int layers, unsigned long long* d_cyc, const char* name) {
    const size_t smem = 1024 + (TSMODE == 0 ? 4 * kSlot : 0) + (size_t)n_slots * kSlot;
    cudaFuncSetAttribute(k_perf_shared<TSMODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; ++rep) {
        cudaMemset(d_cyc, 0, 8);
        k_perf_shared<TSMODE><<<148, 128, smem>>>(wimg, n_slots, layers, d_cyc);
        cudaError_t e = cudaDeviceSynchronize();
        unsigned long long c = 0; cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
        unsigned int to = 0; cudaMemcpyFromSymbol(&to, umma::g_umma_timeout, sizeof(to));
        if (rep == 1)
            printf("TS_PROBE perf %-28s slots=%2d: %8.1f cycles/tile-layer (two tiles per weight pass)  err=%s timeout=%u\n", name,
                   n_slots, (double)c / 148.0 / layers / 2.0, cudaGetErrorName(e), to);
    }
}

// tensor-memory -> register read throughput: `nwarps` warps per CTA stream tcgen05.ld.32x32b.x32 over their lane quarter
__global__ void __launch_bounds__(512, 1) k_ldtm(int iters, int per_wait, unsigned long long* cycles, uint32_t* sink) {
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) umma::tmem_alloc(&tmem_slot, 512);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64 % 512);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += per_wait) {
        uint32_t r[32];
        for (int j = 0; j < per_wait; ++j) {
            umma::tmem_ld32(base + (uint32_t)(((i + j) * 32) % 448), r);
            if (j + 1 == per_wait) umma::tmem_ld_wait();
            acc ^= r[(i + j) & 31];
        }
    }
    umma::tmem_ld_wait();
    const long long dt = clock64() - t0;
    if ((threadIdx.x & 31) == 0) atomicMax(cycles, (unsigned long long)dt);
    if (acc == 0x12345678u) sink[0] = acc;
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

// The forward epilogue of one 128 x 256 layer in isolation (no MMAs, no weight stream): `nwarps` = 4 (one tile) or 8
// (two tiles at once).  WITH_MMA = 1 adds a warp that keeps the tensor pipe busy with SS MMAs on resident operands.
template <int WITH_MMA>
__global__ void __launch_bounds__(384, 1) k_epi(int layers, int nwarps, unsigned long long* cycles, const float* gbias) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                      // 2 operand tiles
    uint8_t* sW = smem + 2 * sm100::kATile;  // 4 stages (garbage is fine)
    float* sBias = (float*)(sW + 4 * kSlot);
    __shared__ uint32_t tmem_slot;
    __shared__ uint64_t done;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { umma::mbar_init(&done, 1); umma::fence_mbar_init(); }
    if (warp == 1) umma::tmem_alloc(&tmem_slot, 512);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sBias[i] = gbias[i];
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp >= 4 && warp < 4 + nwarps) {
        const int g = (warp - 4) >> 2, q = warp & 3, row = q * 32 + lane;
        uint8_t* sAg = sA + g * sm100::kATile;
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)g * 256u;
        uint32_t a8[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) a8[c] = umma::smem_u32(sAg + row * 128 + ((c ^ (row & 7)) << 4));
        sm100::HeadAcc acc = {0ull, 0ull, 0ull, 0ull, 0ull};
        const float* bs = sm100::smem_fptr(sBias, sm100::order_token());
        const long long t0 = clock64();
        for (int l = 0; l < layers; ++l) {
            sm100::fwd_epilogue_layer<8, 0, true, false, true>(taddr, bs, a8, bs, bs, acc, nullptr);
            umma::tc_fence_before();
            umma::fence_proxy_async_smem();
            __syncwarp();
        }
        const long long dt = clock64() - t0;
        if (lane == 0) atomicMax(cycles, (unsigned long long)dt);
    } else if (WITH_MMA && warp == 1) {
        const uint64_t dA = umma::make_sdesc(umma::smem_u32(sA), 16, 1024, umma::SWZ_128B);
        const uint64_t dB = umma::make_sdesc(umma::smem_u32(sW), 16, 1024, umma::SWZ_128B);
        const uint32_t id256 = umma::make_idesc(128, 256, 0, 0);
        // about as many layers' worth of MMAs as the epilogue loop will take (2048 cycles each)
        for (int l = 0; l < layers; ++l) {
            if (umma::elect_one()) {
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma::mma_bf16(tmem + 256, dA + ((c * kSlot) >> 4) + ks * 2, dB + (((c & 1) * 2 * kSlot) >> 4) + ks * 2, id256, 1u);
            }
            __syncwarp();
        }
        if (umma::elect_one()) umma::mma_commit(&done);
        __syncwarp();
        umma::mbar_wait(&done, 0);
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 1) umma::tmem_dealloc(tmem, 512);
}

// Half-K weight stages: [128 n x 32 k] bf16, 64-byte swizzle, 8 KB; a pair of adjacent stages is one N = 256, K = 32
// operand (two MMAs), released right after them.  Same 64 KB of ring, twice as many, shorter-lived slots.
__global__ void __launch_bounds__(128, 1) k_perf_halfk(const uint8_t* wimg, int n_slots, int layers, unsigned long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sW = smem + 4 * kSlot;
    __shared__ uint64_t full[32], empty[32], done;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 32; ++i) { umma::mbar_init(&full[i], 1); umma::mbar_init(&empty[i], 1); }
        umma::mbar_init(&done, 1);
        umma::fence_mbar_init();
    }
    if (warp == 1) umma::tmem_alloc(&tmem_slot, 512);
    {
        const int row = threadIdx.x;
        for (int k = 0; k < 256; ++k)
            *reinterpret_cast<__nv_bfloat16*>(sA + (k >> 6) * kSlot + umma::sw128_offset(row, k & 63)) = __float2bfloat16(a_val(row, k));
        umma::fence_proxy_async_smem();
    }
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const long long t0 = clock64();
    const int total = layers * 16;            // 8 KB stages per layer
    if (warp == 0) {
        int stage = 0; uint32_t ph = 0;
        for (int s = 0; s < total; ++s) {
            umma::mbar_wait(&empty[stage], ph ^ 1);
            if (umma::elect_one()) {
                umma::mbar_arrive_expect_tx(&full[stage], 8192);
                umma::bulk_g2s(sW + stage * 8192, wimg + (size_t)(s & 15) * 8192, 8192, &full[stage]);
            }
            __syncwarp();
            if (++stage == n_slots) { stage = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        int stage = 0; uint32_t ph = 0;
        const uint64_t dA = umma::make_sdesc(umma::smem_u32(sA), 16, 1024, umma::SWZ_128B);
        const uint64_t dB = umma::make_sdesc(umma::smem_u32(sW), 16, 512, umma::SWZ_64B);
        const uint32_t id256 = umma::make_idesc(128, 256, 0, 0);
        for (int s = 0; s < total; s += 2) {
            const int c = (s >> 2) & 3, kh = (s >> 1) & 1;
            umma::mbar_wait(&full[stage], ph);
            umma::mbar_wait(&full[stage + 1], ph);
            umma::tc_fence_after();
            if (umma::elect_one()) {
                const uint64_t db = dB + (uint64_t)((stage * 8192) >> 4);
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
                    umma::mma_bf16(tmem, dA + ((c * kSlot) >> 4) + kh * 4 + ks * 2, db + ks * 2, id256, (c | kh | ks) ? 1u : 0u);
                umma::mma_commit(&empty[stage]); umma::mma_commit(&empty[stage + 1]);
            }
            __syncwarp();
            stage += 2;
            if (stage >= n_slots) { stage = 0; ph ^= 1; }
        }
        if (umma::elect_one()) umma::mma_commit(&done);
        __syncwarp();
        umma::mbar_wait(&done, 0);
        if (threadIdx.x == 32) atomicAdd(cycles, (unsigned long long)(clock64() - t0));
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 1) umma::tmem_dealloc(tmem, 512);
}

// SM-to-SM bulk copy bandwidth: every CTA of a cluster sends `bytes` from its shared memory to the next CTA's shared
// memory (a ring), `iters` times, with cp.async.bulk.shared::cluster.shared::cta completing on the receiver's mbarrier.
__global__ void __launch_bounds__(64, 1) k_dsmem(int csize, uint32_t bytes, int iters, unsigned long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* src = smem;                 // what this CTA sends
    uint8_t* dst = smem + 65536;         // where the previous CTA of the ring writes
    __shared__ uint64_t rx_bar;          // completes when the upstream neighbour's data has landed here
    __shared__ uint64_t tx_credit;       // the downstream neighbour has consumed my previous message and re-armed
    const uint32_t rank = umma::cluster_ctarank();
    if (threadIdx.x == 0) {
        umma::mbar_init(&rx_bar, 1); umma::mbar_init(&tx_credit, 1); umma::fence_mbar_init();
        umma::mbar_arrive_expect_tx(&rx_bar, bytes);              // armed for round 0 before anybody may send
    }
    for (uint32_t i = threadIdx.x; i < bytes / 4; i += blockDim.x) ((uint32_t*)src)[i] = i * 2654435761u + rank;
    umma::fence_proxy_async_smem();
    umma::cluster_sync_all();
    const uint32_t next = (rank + 1) % (uint32_t)csize, prev = (rank + (uint32_t)csize - 1) % (uint32_t)csize;
    const uint32_t dst_remote = umma::mapa(umma::smem_u32(dst), next);
    const uint32_t bar_remote = umma::mapa(umma::smem_u32(&rx_bar), next);
    const uint32_t credit_remote = umma::mapa(umma::smem_u32(&tx_credit), prev);
    const long long t0 = clock64();
    if (threadIdx.x == 0) {
        for (int it = 0; it < iters; ++it) {
            if (it > 0) umma::mbar_wait_cluster(&tx_credit, (uint32_t)(it - 1) & 1u);
            asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst_remote), "r"(umma::smem_u32(src)), "r"(bytes), "r"(bar_remote) : "memory");
            umma::mbar_wait_cluster(&rx_bar, (uint32_t)it & 1u);  // my inbox is full
            if (it + 1 < iters) umma::mbar_arrive_expect_tx(&rx_bar, bytes);
            umma::mbar_arrive_cluster(credit_remote);             // upstream may send the next round
        }
    }
    __syncthreads();
    const long long dt = clock64() - t0;
    umma::cluster_sync_all();
    if (threadIdx.x == 0) atomicMax(cycles, (unsigned long long)dt);
}

template <int MODE, int FILL>
void run_perf(const uint8_t* wimg, int n_slots, int layers, unsigned long long* d_cyc, const char* name) {
    const size_t smem = 1024 + (MODE == 0 ? 4 * kSlot : 0) + (size_t)n_slots * kSlot;
    cudaFuncSetAttribute(k_perf<MODE, FILL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; ++rep) {
        cudaMemset(d_cyc, 0, 8);
        k_perf<MODE, FILL><<<148, 128, smem>>>(wimg, n_slots, layers, d_cyc);
        cudaError_t e = cudaDeviceSynchronize();
        unsigned long long c = 0; cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
        unsigned int to = 0; cudaMemcpyFromSymbol(&to, umma::g_umma_timeout, sizeof(to));
        if (rep == 1)
            printf("TS_PROBE perf %-28s slots=%2d: %8.1f cycles/layer (128x256x256; tensor floor 2048)  err=%s timeout=%u\n", name,
                   n_slots, (double)c / 148.0 / layers, cudaGetErrorName(e), to);
    }
}
}  // namespace

int main(int argc, char** argv) {
    const bool check_only = argc > 1 && std::string(argv[1]) == "--check-only";
    const bool dsmem_only = argc > 1 && std::string(argv[1]) == "--dsmem-only";
    uint8_t* wimg; cudaMalloc(&wimg, 16 * kSlot);
    if (dsmem_only) {
        unsigned long long* dc; cudaMalloc(&dc, 8); cudaMemset(dc, 0, 8);
        const size_t dsm = 1024 + 2 * 65536;
        cudaFuncSetAttribute(k_dsmem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = dsm;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t e1 = cudaLaunchKernelEx(&cfg, k_dsmem, 2, 4096u, 2, dc);
        cudaError_t e2 = cudaDeviceSynchronize();
        printf("dsmem-only launch=%s sync=%s\n", cudaGetErrorName(e1), cudaGetErrorName(e2));
        return 0;
    }
    k_make_weights<<<8, 256>>>(wimg);
    float *o_ss, *o_ts; cudaMalloc(&o_ss, 128 * 256 * 4); cudaMalloc(&o_ts, 128 * 256 * 4);
    cudaMemset(o_ss, 0, 128 * 256 * 4); cudaMemset(o_ts, 0, 128 * 256 * 4);
    const size_t smem = 1024 + 12 * kSlot;
    cudaFuncSetAttribute(k_check, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_check<<<1, 160, smem>>>(wimg, o_ss, o_ts);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> hs(128 * 256), ht(128 * 256);
    cudaMemcpy(hs.data(), o_ss, hs.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(ht.data(), o_ts, ht.size() * 4, cudaMemcpyDeviceToHost);
    double max_ss = 0, max_ts = 0; int diff = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 256; ++n) {
            double ref = 0;
            for (int k = 0; k < 256; ++k) ref += (double)a_val(m, k) * (double)w_val(n, k);   // values are exact in bf16
            max_ss = fmax(max_ss, fabs(hs[m * 256 + n] - ref));
            max_ts = fmax(max_ts, fabs(ht[m * 256 + n] - ref));
            diff += hs[m * 256 + n] != ht[m * 256 + n];
        }
    unsigned int to = 0; cudaMemcpyFromSymbol(&to, umma::g_umma_timeout, sizeof(to));
    const bool ok = e == cudaSuccess && !to && max_ss < 1e-3 && max_ts < 1e-3;
    printf("TS_PROBE check err=%s timeout=%u max|ss-ref|=%.3g max|ts-ref|=%.3g ss!=ts:%d %s\n", cudaGetErrorName(e), to, max_ss, max_ts,
           diff, ok ? "OK" : "FAIL");
    if (check_only) { printf("TS_PROBE %s\n", ok ? "PASS" : "FAIL"); return ok ? 0 : 1; }
    unsigned long long* d_cyc; cudaMalloc(&d_cyc, 8);
    const int layers = 4000;
    run_perf<0, 0>(wimg, 4, layers, d_cyc, "SS N=256, no refill");
    run_perf<1, 0>(wimg, 4, layers, d_cyc, "TS N=128, no refill");
    run_perf<0, 1>(wimg, 4, layers, d_cyc, "SS N=256, streamed");
    run_perf<0, 1>(wimg, 8, layers, d_cyc, "SS N=256, streamed");
    for (int ns : {8, 12}) {
        const size_t hsm = 1024 + 4 * kSlot + (size_t)ns * 8192;
        cudaFuncSetAttribute(k_perf_halfk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsm);
        for (int rep = 0; rep < 2; ++rep) {
            cudaMemset(d_cyc, 0, 8);
            k_perf_halfk<<<148, 128, hsm>>>(wimg, ns, layers, d_cyc);
            cudaError_t e2 = cudaDeviceSynchronize();
            unsigned long long c = 0; cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
            if (rep) printf("TS_PROBE perf %-28s slots=%2d: %8.1f cycles/layer (8 KB half-K stages, %d KB ring)  err=%s\n", "SS N=256, streamed", ns,
                            (double)c / 148.0 / layers, ns * 8, cudaGetErrorName(e2));
        }
    }
    run_perf<1, 1>(wimg, 4, layers, d_cyc, "TS N=128, streamed");
    run_perf<1, 1>(wimg, 8, layers, d_cyc, "TS N=128, streamed");
    run_perf<1, 1>(wimg, 12, layers, d_cyc, "TS N=128, streamed");
    {
        uint32_t* sink; cudaMalloc(&sink, 4);
        for (int nw : {4, 8})
            for (int pw : {1, 2, 4}) {
                const int iters = 4096;
                cudaMemset(d_cyc, 0, 8);
                k_ldtm<<<148, nw * 32, 0>>>(iters, pw, d_cyc, sink);
                cudaError_t e2 = cudaDeviceSynchronize();
                unsigned long long c = 0; cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
                printf("TS_PROBE ldtm warps=%2d loads/wait=%d: %6.1f B/clk/SM (%5.1f clk per 4 KB load per warp)  err=%s\n", nw, pw,
                       (double)nw * iters * 4096.0 / (double)c, (double)c / iters, cudaGetErrorName(e2));
            }
    }
    {
        float* gb; cudaMalloc(&gb, 1024); cudaMemset(gb, 0, 1024);
        const size_t esm = 1024 + 2 * sm100::kATile + 4 * kSlot + 1024;
        cudaFuncSetAttribute(k_epi<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)esm);
        cudaFuncSetAttribute(k_epi<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)esm);
        for (int with_mma = 0; with_mma < 2; ++with_mma)
            for (int nw : {4, 8}) {
                if (with_mma && nw == 8) continue;       // the MMA warp accumulates into tile 1's columns
                for (int rep = 0; rep < 2; ++rep) {
                    cudaMemset(d_cyc, 0, 8);
                    if (with_mma) k_epi<1><<<148, 384, esm>>>(2000, nw, d_cyc, gb); else k_epi<0><<<148, 384, esm>>>(2000, nw, d_cyc, gb);
                    cudaError_t e2 = cudaDeviceSynchronize();
                    unsigned long long c = 0; cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
                    if (rep) printf("TS_PROBE epilogue warps=%d mma=%d: %7.1f cycles per 128x256 layer per tile  err=%s\n", nw, with_mma, (double)c / 2000.0, cudaGetErrorName(e2));
                }
            }
    }
    for (int cs : {2, 4, 8}) {
        const uint32_t bytes = 65536; const int iters = 200;
        const size_t dsm = 1024 + 2 * 65536;
        cudaFuncSetAttribute(k_dsmem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm);
        cudaFuncSetAttribute(k_dsmem, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((148 / cs) * cs); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = dsm;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int ncl = 0; cudaOccupancyMaxActiveClusters(&ncl, k_dsmem, &cfg);
        if (ncl > 0 && ncl * cs < (int)cfg.gridDim.x) cfg.gridDim = dim3(ncl * cs);
        cudaMemset(d_cyc, 0, 8);
        cudaError_t e1 = cudaLaunchKernelEx(&cfg, k_dsmem, cs, bytes, iters, d_cyc);
        cudaError_t e2 = cudaDeviceSynchronize();
        unsigned long long c = 0; cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
        unsigned int to = 0; cudaMemcpyFromSymbol(&to, umma::g_umma_timeout, sizeof(to));
        printf("TS_PROBE dsmem ring of %d CTAs (%u clusters): %7.1f cycles per 64 KB hop = %5.1f B/clk per SM  launch=%s sync=%s timeout=%u\n",
               cs, cfg.gridDim.x / cs, (double)c / iters, 65536.0 * iters / (double)c, cudaGetErrorName(e1), cudaGetErrorName(e2), to);
    }
    run_shared<1>(wimg, 10, layers / 2, d_cyc, "TS N=128, X+Y share a pass");
    run_shared<1>(wimg, 12, layers / 2, d_cyc, "TS N=128, X+Y share a pass");
    run_shared<0>(wimg, 9, layers / 2, d_cyc, "SS N=128, X+Y share a pass");
    printf("TS_PROBE %s\n", ok ? "PASS" : "FAIL");
    return ok ? 0 : 1;
}
