// sm100_common.cuh -- pieces shared by the tensor-core kernels (render_sm100.cu, backward_sm100.cu):
// tile constants, layer descriptors, the positional-encoding operand writer, the layer plan.
#pragma once
#include "render_sm100.cuh"
#include "mlp_fp32.cuh"
#include <cuda.h>
#include <cstdlib>
#include "umma.cuh"
#include <type_traits>
#include <mutex>

namespace sm100 {


constexpr int kW = 256;                 // layer width the tensor-core path implements
constexpr int kLx = 10, kLd = 4;        // PE frequencies (63 / 27 channels)
constexpr int kTileRows = 128;          // UMMA M
constexpr int kSlot = 16384;            // one weight stage: [128 n x 64 k] bf16, 128-B swizzle
constexpr int kNumStages = 4;
constexpr int kABlock = 16384;          // activation K-block [128 rows x 64] bf16, 128-B swizzle
constexpr int kDirBlock = 8192;         // PE(viewdir) block [128 rows x 32] bf16, 64-B swizzle
constexpr int kATile = 4 * kABlock + kDirBlock;
constexpr int kThreads = 384;           // warpgroup 0: producer, MMA issuer, 2 auxiliary warps; warps 4-7 group X, 8-11 group Y
#ifndef CNB_EPI_PREFETCH
#define CNB_EPI_PREFETCH 1      // epilogues load the next column group's bias / head rows before storing the current group
#endif
constexpr int kRegsAux = 64;            // setmaxnreg budgets: the pipeline warps give registers to the epilogue warps,
constexpr int kRegsCompute = 216;       // which keep four 32-column TMEM loads in flight (384*168 >= 128*64 + 256*216)
constexpr int kMaxLayers = 2 * CNB_MAX_BLOCKS + 4;

// ---- optional cycle accounting of the pipeline roles (build with -DCNB_TRACE; debugging only) ----------
#ifdef CNB_TRACE
static __device__ unsigned long long g_trace[32];   // one copy per translation unit
#define CNB_TR_NOW() clock64()
#define CNB_TR_DECL(name) unsigned long long name = 0ull
#define CNB_TR(acc, ...) do { const long long t_ = clock64(); __VA_ARGS__; (acc) += (unsigned long long)(clock64() - t_); } while (0)
#define CNB_TR_PTR(ptr, ...) do { const long long t_ = clock64(); __VA_ARGS__; if (ptr) *(ptr) += (unsigned long long)(clock64() - t_); } while (0)
#define CNB_TR_FLUSH(slot, acc) do { if ((threadIdx.x & 31) == 0) atomicAdd(&g_trace[slot], (acc)); } while (0)
// event log of CTA 0 (timeline of the pipeline): lane 0 of a warp appends (clock, code)
static __device__ unsigned long long g_events[4][16384];
static __device__ unsigned int g_event_count[4];
#define CNB_EV(lane_, slot_, code_) do { if (blockIdx.x == 0 && (lane_) == 0) { const unsigned int i_ = g_event_count[slot_]++; \
    if (i_ < 16384u) g_events[slot_][i_] = ((unsigned long long)(clock64() & 0xffffffffffffull) << 16) | (unsigned long long)((code_) & 0xffff); } } while (0)
#else
#define CNB_EV(lane_, slot_, code_) do { } while (0)
#define CNB_TR_NOW() 0ll
#define CNB_TR_DECL(name) unsigned long long name = 0ull; (void)name
#define CNB_TR(acc, ...) do { __VA_ARGS__; } while (0)
#define CNB_TR_FLUSH(slot, acc) do { } while (0)
#endif

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

// ---- packed fp32x2 / bf16x2 helpers (sm_100: FADD2 / FFMA2, F2FP with fused ReLU) -----------------
__device__ __forceinline__ uint64_t pk2(uint32_t lo, uint32_t hi) {
    uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r;
}
__device__ __forceinline__ uint64_t pk2f(float lo, float hi) { return pk2(__float_as_uint(lo), __float_as_uint(hi)); }
__device__ __forceinline__ void unpk2(uint64_t v, float& lo, float& hi) {
    uint32_t a, b; asm("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(v)); lo = __uint_as_float(a); hi = __uint_as_float(b);
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
    uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
// two fp32 -> packed bf16x2 (first value in the low half), optionally with ReLU fused into the conversion
template <bool kRelu>
__device__ __forceinline__ uint32_t cvt_bf16x2(uint64_t v) {
    float lo, hi; unpk2(v, lo, hi);
    uint32_t d;
    if (kRelu) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
template <int kOff>
__device__ __forceinline__ void st_shared_v4_off(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    // volatile keeps it ordered against the fences / barrier arrivals (all volatile); no memory clobber, so ordinary
    // loads (bias rows, head weights) can be scheduled across the operand stores
    asm volatile("st.shared.v4.b32 [%0 + %5], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d), "n"(kOff));
}

// ---- rows every thread reads at the same address: bias rows, head weights --------------------------------
// A warp-wide LDS.128 of ONE address still costs four shared-memory wavefronts (128-bit accesses are served per
// quarter warp), so a 256-float bias row costs each epilogue warp 256 wavefronts -- bias and head-weight
// broadcasts were 60 % of the LSU wavefronts of K2 and ~17 % of its shared-memory pipe time (round-2 profile), and
// that pipe looked like what bounds the kernels.  EXPERIMENT (opt-in at compile time, CNB_K1_CONST / CNB_K2_*_CONST;
// measured slower, kept as a record): rows that are the same for the whole launch (biases of the layers that are
// not code-conditioned, the sigma / rgb.2 head weights) live in CONSTANT memory: the loads become LDCU.128 into
// uniform registers and FADD2 / FFMA2 take them as uniform operands, no shared-memory traffic at all.  Result:
// K2 9.55 -> 10.5 ms, K1 18.5 -> 17.5 M rays/s; and skipping the bias adds altogether (CNB_EXPERIMENT_NOBIAS) leaves
// K2 unchanged: the broadcasts are not on K2's critical path (its period is set by the weight ring).
constexpr int kCrowSlots = 4;        // launch-wide bias rows: 0 encoding_xyz, 1 encoding_shape, 2 encoding_viewdir, 3 rgb.0
struct ConstRows {
    float bias[kCrowSlots][kW];      // fixed slots, so that every load has an immediate address (LDCU; a run-time
                                     // layer index in a vector register turns them into per-thread LDC.64)
    float w_sigma[kW];               // sigma.0.weight
    float w_rgb2[3 * (kW / 2)];      // rgb.2.weight
    float b_sigma, b_rgb2[3];
};
static __constant__ ConstRows c_rows;      // one copy per translation unit; uploaded by ConstRowsScope before the launches
constexpr int kCrowBias = 0;
constexpr int kCrowWsig = kCrowSlots * kW;
constexpr int kCrowWrgb = kCrowWsig + kW;
constexpr int kCrowBsig = kCrowWrgb + 3 * (kW / 2);
__device__ __forceinline__ float crow(int idx) { return reinterpret_cast<const float*>(&c_rows)[idx]; }
// "pointer" to float `idx` of c_rows for ld_vec4<2>: an index in disguise, never dereferenced
__device__ __forceinline__ const float* crow_ptr(int idx) { return reinterpret_cast<const float*>((uintptr_t)idx << 2); }

// Vector load of 4 floats that every lane reads from the same address.  SRC = 0: read-only global load; 1: `p`
// carries a shared-window address (see smem_fptr), LDS broadcast; 2: `p` is crow_ptr(index), constant memory.
// The global form misses the small L1 left beside 220 KB of shared memory, and 64 such loads per layer with the
// few registers available to prefetch them were the dominant stall of the epilogues (ncu source view, round 1).
template <int SRC>
__device__ __forceinline__ float4 ld_vec4(const float* p) {
    if (SRC == 1) {
        // not volatile, no memory clobber: the scheduler may hoist these well ahead of their use (the address carries
        // a data dependence on the barrier that published the staging buffer, see bar_sync_token)
        float4 v;
        asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
            : "r"((uint32_t)(uintptr_t)p));
        return v;
    }
    if (SRC == 2) {
        const int idx = (int)((uintptr_t)p >> 2);
        return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(&c_rows) + idx);
    }
    return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ const float* smem_fptr(const void* shared_ptr, uint32_t token = 0u) {
    return reinterpret_cast<const float*>((uintptr_t)(umma::smem_u32(shared_ptr) + token));
}
// 0 in a register the compiler cannot see through, defined after every earlier volatile asm (barriers included).
__device__ __forceinline__ uint32_t order_token() { uint32_t t; asm volatile("mov.u32 %0, 0;" : "=r"(t) :: "memory"); return t; }
// Named barrier that also returns 0 in a register defined by the barrier instruction itself: adding it to a
// shared-memory address orders (non-volatile) loads from that address after the barrier.
__device__ __forceinline__ uint32_t bar_sync_token(uint32_t id, uint32_t threads) {
    uint32_t tok;
    asm volatile("bar.sync %1, %2;\n\tmov.u32 %0, 0;" : "=r"(tok) : "r"(id), "r"(threads) : "memory");
    return tok;
}

// ReLU bit words (one per row and 32-column chunk): column c of the chunk is bit 8 (c & 3) + 7 - (c >> 2), i.e. byte k
// holds columns k, k + 4, ..., k + 28 from its bit 7 down.  With that order ONE shift (x = word << j) puts the bits of the
// four columns 4 j .. 4 j + 3 into the four byte sign positions, and a byte permute with sign replication turns two of
// them into the 0xffff / 0x0000 halves that mask a packed bf16 pair: 1.25 instructions per column in the backward
// epilogues instead of a bit test and a select each (the masks were ~1 ms of the 8.5 ms of K2 per step).
__device__ __forceinline__ uint32_t relu_mask_pair(uint32_t x, int h) {      // halves mask of columns 4 j + 2 h, 4 j + 2 h + 1
    uint32_t d;
    if (h == 0) asm("prmt.b32 %0, %1, %1, 0x9988;" : "=r"(d) : "r"(x));
    else asm("prmt.b32 %0, %1, %1, 0xbbaa;" : "=r"(d) : "r"(x));
    return d;
}

// Accumulators of the narrow heads carried across a layer's epilogue as packed pairs.
struct HeadAcc { uint64_t sig2, r2, g2, b2; uint64_t mask_policy; };

// Forward epilogue of 32 accumulator columns (one tcgen05.ld) of this thread's row:
// + bias, [ReLU], [sigma / rgb head partial dot products in fp32], bf16 pack, swizzled store of
// the next operand, [ReLU sign bits for the backward].   KIND: 0 hidden, 1 encoding_shape
// (+ sigma head, no ReLU), 2 rgb.0 (+ rgb head).   a8[c] = shared address of 16-byte chunk c of
// this row inside K-block 0.
template <int CC, int KIND, bool STORE, bool MASK, int SRC = 0, int HSRC = SRC, bool PRE_IN = false, bool PRE_OUT = false>
__device__ __forceinline__ void fwd_epilogue32(const uint32_t (&rr)[32], const float* __restrict__ bias,
                                               const uint32_t (&a8)[8], const float* __restrict__ w_sigma,
                                               const float* __restrict__ w_rgb2, HeadAcc& acc, uint32_t* mscr,
                                               bool do_store = true, float4* pre = nullptr) {
    constexpr bool RELU = (KIND != 1);
#if CNB_EPI_PREFETCH
    // The bias row (and head weights) of column group j8 + 1 are loaded BEFORE group j8 is stored: ptxas cannot move a
    // ld.shared above an earlier st.shared (possible alias), so loads placed next to their use sat two instructions
    // ahead of it and every group paid the ~35-cycle shared-memory latency (ncu source view: 4 x 38 of a chunk's ~410
    // cycles).
    // PRE_IN / PRE_OUT: the first group of a chunk is loaded by the previous chunk (pre[0..1]).
    float4 nb0, nb1;
    if (PRE_IN) { nb0 = pre[0]; nb1 = pre[1]; }
    else { nb0 = ld_vec4<SRC>(bias + CC * 32); nb1 = ld_vec4<SRC>(bias + CC * 32 + 4); }
#endif
    // ReLU bit word of the chunk (see relu_mask_pair): byte k collects the sign bits of columns k, k + 4, ..., k + 28,
    // first column in bit 7 -- four independent funnel-shift chains
    uint32_t sg0 = 0u, sg1 = 0u, sg2 = 0u, sg3 = 0u;
#pragma unroll
    for (int j8 = 0; j8 < 4; ++j8) {
        constexpr int dummy = 0; (void)dummy;
        const int col = CC * 32 + j8 * 8;
        uint64_t v[4];
#ifdef CNB_EXPERIMENT_NOBIAS     // timing experiment only (wrong results): what the bias broadcasts cost
        v[0] = pk2(rr[j8 * 8 + 0], rr[j8 * 8 + 1]); v[1] = pk2(rr[j8 * 8 + 2], rr[j8 * 8 + 3]);
        v[2] = pk2(rr[j8 * 8 + 4], rr[j8 * 8 + 5]); v[3] = pk2(rr[j8 * 8 + 6], rr[j8 * 8 + 7]);
#else
#if CNB_EPI_PREFETCH
        const float4 b0 = nb0, b1 = nb1;
        if (j8 < 3) { nb0 = ld_vec4<SRC>(bias + col + 8); nb1 = ld_vec4<SRC>(bias + col + 12); }
        else if (PRE_OUT) { pre[0] = ld_vec4<SRC>(bias + col + 8); pre[1] = ld_vec4<SRC>(bias + col + 12); }
#else
        const float4 b0 = ld_vec4<SRC>(bias + col);
        const float4 b1 = ld_vec4<SRC>(bias + col + 4);
#endif
        v[0] = fadd2(pk2(rr[j8 * 8 + 0], rr[j8 * 8 + 1]), pk2f(b0.x, b0.y));
        v[1] = fadd2(pk2(rr[j8 * 8 + 2], rr[j8 * 8 + 3]), pk2f(b0.z, b0.w));
        v[2] = fadd2(pk2(rr[j8 * 8 + 4], rr[j8 * 8 + 5]), pk2f(b1.x, b1.y));
        v[3] = fadd2(pk2(rr[j8 * 8 + 6], rr[j8 * 8 + 7]), pk2f(b1.z, b1.w));
#endif
#ifndef CNB_EXPERIMENT_NOMASK     // (timing experiment only, wrong gradients: what the ReLU bit masks cost)
        if (MASK && RELU) {     // columns 8 j8 + e, e = 0..7: e and e + 4 go to byte e & 3, in this order
            float e0, e1, e2, e3, e4, e5, e6, e7;
            unpk2(v[0], e0, e1); unpk2(v[1], e2, e3); unpk2(v[2], e4, e5); unpk2(v[3], e6, e7);
            sg0 = __funnelshift_l(__float_as_uint(e0), sg0, 1); sg1 = __funnelshift_l(__float_as_uint(e1), sg1, 1);
            sg2 = __funnelshift_l(__float_as_uint(e2), sg2, 1); sg3 = __funnelshift_l(__float_as_uint(e3), sg3, 1);
            sg0 = __funnelshift_l(__float_as_uint(e4), sg0, 1); sg1 = __funnelshift_l(__float_as_uint(e5), sg1, 1);
            sg2 = __funnelshift_l(__float_as_uint(e6), sg2, 1); sg3 = __funnelshift_l(__float_as_uint(e7), sg3, 1);
        }
#endif
        if (KIND == 1) {        // sigma head on the fp32 feature (reference src/model.py:45)
            const float4 w0 = ld_vec4<HSRC>(w_sigma + col);
            const float4 w1 = ld_vec4<HSRC>(w_sigma + col + 4);
            acc.sig2 = ffma2(v[0], pk2f(w0.x, w0.y), acc.sig2); acc.sig2 = ffma2(v[1], pk2f(w0.z, w0.w), acc.sig2);
            acc.sig2 = ffma2(v[2], pk2f(w1.x, w1.y), acc.sig2); acc.sig2 = ffma2(v[3], pk2f(w1.z, w1.w), acc.sig2);
        } else if (KIND == 2) { // rgb.2 on the fp32 hidden (reference src/model.py:52)
            uint64_t h[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { float lo, hi; unpk2(v[i], lo, hi); h[i] = pk2f(fmaxf(lo, 0.f), fmaxf(hi, 0.f)); }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float4 w0 = ld_vec4<HSRC>(w_rgb2 + k * (kW / 2) + col);
                const float4 w1 = ld_vec4<HSRC>(w_rgb2 + k * (kW / 2) + col + 4);
                uint64_t& a = (k == 0) ? acc.r2 : (k == 1 ? acc.g2 : acc.b2);
                a = ffma2(h[0], pk2f(w0.x, w0.y), a); a = ffma2(h[1], pk2f(w0.z, w0.w), a);
                a = ffma2(h[2], pk2f(w1.x, w1.y), a); a = ffma2(h[3], pk2f(w1.z, w1.w), a);
            }
        }
        if (STORE && do_store) {
            constexpr int blk = CC >> 1;
            const int chunk = ((CC & 1) << 2) + j8;
            st_shared_v4_off<blk * kABlock>(a8[chunk], cvt_bf16x2<RELU>(v[0]), cvt_bf16x2<RELU>(v[1]),
                                            cvt_bf16x2<RELU>(v[2]), cvt_bf16x2<RELU>(v[3]));
        }
    }
#ifndef CNB_EXPERIMENT_NOMASK
    if (MASK && RELU)       // bit set <=> pre-activation >= +0
        umma::st_global_hint(mscr + (size_t)CC * kTileRows, ~(sg0 | (sg1 << 8) | (sg2 << 16) | (sg3 << 24)), acc.mask_policy);
#endif
}

// Ties the registers of a tcgen05.ld result to program order after the tcgen05.wait::ld that precedes this call (an
// empty volatile asm per register: no instruction, but nothing that uses r[] may be scheduled above it).
__device__ __forceinline__ void tmem_regs_ready(uint32_t (&r)[32]) {
#pragma unroll
    for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(r[i]));
}

// A whole layer: NCC x 32 columns.  DEEP = false: two tcgen05.ld in flight per wait.  DEEP = true (NCC = 8): software
// pipeline, the next two loads are issued before the current two chunks are processed, so only the first wait of a
// layer exposes the TMEM latency (~300 cycles each; with CTA pairs the kernels' period is MMA + epilogue, so the
// epilogue's stalls count -- in round 1, when the weight ring set the period, this measured +-0).
struct NoBlockHook { __device__ __forceinline__ void operator()(int) const {} };
// `hook(b)` is called as soon as 64-column block b of the next operand is complete in shared memory (K2 hands it to
// the auxiliary warp's bulk store then, instead of after the whole layer) -- for every block but the last, which the
// caller hands over when it publishes the operand.
template <int NCC, int KIND, bool STORE, bool MASK, int SRC = 0, int HSRC = SRC, bool DEEP = false, class Hook = NoBlockHook>
__device__ __forceinline__ void fwd_epilogue_layer(uint32_t taddr, const float* __restrict__ bias,
                                                   const uint32_t (&a8)[8], const float* __restrict__ w_sigma,
                                                   const float* __restrict__ w_rgb2, HeadAcc& acc, uint32_t* mscr,
                                                   Hook hook = Hook()) {
    if constexpr (DEEP && NCC == 8) {
        uint32_t ra[32], rb[32], rc[32], rd[32];
        umma::tmem_ld32(taddr + 0, ra);
        umma::tmem_ld32(taddr + 32, rb);
        umma::tmem_ld_wait();
        tmem_regs_ready(ra); tmem_regs_ready(rb);
        umma::tmem_ld32(taddr + 64, rc);
        umma::tmem_ld32(taddr + 96, rd);
        fwd_epilogue32<0, KIND, STORE, MASK, SRC, HSRC>(ra, bias, a8, w_sigma, w_rgb2, acc, mscr);
        fwd_epilogue32<1, KIND, STORE, MASK, SRC, HSRC>(rb, bias, a8, w_sigma, w_rgb2, acc, mscr);
        hook(0);
        umma::tmem_ld_wait();
        tmem_regs_ready(rc); tmem_regs_ready(rd);
        umma::tmem_ld32(taddr + 128, ra);
        umma::tmem_ld32(taddr + 160, rb);
        fwd_epilogue32<2, KIND, STORE, MASK, SRC, HSRC>(rc, bias, a8, w_sigma, w_rgb2, acc, mscr);
        fwd_epilogue32<3, KIND, STORE, MASK, SRC, HSRC>(rd, bias, a8, w_sigma, w_rgb2, acc, mscr);
        hook(1);
        umma::tmem_ld_wait();
        tmem_regs_ready(ra); tmem_regs_ready(rb);
        umma::tmem_ld32(taddr + 192, rc);
        umma::tmem_ld32(taddr + 224, rd);
        fwd_epilogue32<4, KIND, STORE, MASK, SRC, HSRC>(ra, bias, a8, w_sigma, w_rgb2, acc, mscr);
        fwd_epilogue32<5, KIND, STORE, MASK, SRC, HSRC>(rb, bias, a8, w_sigma, w_rgb2, acc, mscr);
        hook(2);
        umma::tmem_ld_wait();
        tmem_regs_ready(rc); tmem_regs_ready(rd);
        fwd_epilogue32<6, KIND, STORE, MASK, SRC, HSRC>(rc, bias, a8, w_sigma, w_rgb2, acc, mscr);
        fwd_epilogue32<7, KIND, STORE, MASK, SRC, HSRC>(rd, bias, a8, w_sigma, w_rgb2, acc, mscr);
        return;
    }
#if CNB_EPI_PREFETCH
    float4 pre[2] = {ld_vec4<SRC>(bias), ld_vec4<SRC>(bias + 4)};
#else
    float4* pre = nullptr;
#endif
    auto pair = [&](auto cc_tag) {
        constexpr int CC = decltype(cc_tag)::value;
        constexpr bool P = CNB_EPI_PREFETCH != 0;
        uint32_t ra[32], rb[32];
        umma::tmem_ld32(taddr + CC * 32, ra);
        umma::tmem_ld32(taddr + CC * 32 + 32, rb);
        if constexpr (CC > 0) hook((CC >> 1) - 1);      // the previous block is handed over while this pair's TMEM loads travel
        umma::tmem_ld_wait();
        fwd_epilogue32<CC, KIND, STORE, MASK, SRC, HSRC, P, P>(ra, bias, a8, w_sigma, w_rgb2, acc, mscr, true, pre);
        fwd_epilogue32<CC + 1, KIND, STORE, MASK, SRC, HSRC, P, P && (CC + 2 < NCC)>(rb, bias, a8, w_sigma, w_rgb2, acc, mscr, true, pre);
        // (the LAST block is handed over by the caller together with the operand's publication)
    };
    pair(std::integral_constant<int, 0>{});
    pair(std::integral_constant<int, 2>{});
    if constexpr (NCC == 8) {
        pair(std::integral_constant<int, 4>{});
        pair(std::integral_constant<int, 6>{});
    }
}

// NC (2 or 4) consecutive 32-column chunks of a layer: one thread's share when two warps split the columns of a row.
// Every pointer / address argument is pre-offset to the thread's first chunk (so the column half is a run-time
// value and the code exists once); a8 must point at the K-block that receives the first chunk.
template <int NC, int KIND, bool STORE, bool MASK, int SM>
__device__ __forceinline__ void fwd_epilogue_chunks(uint32_t taddr, const float* __restrict__ bias, const uint32_t (&a8)[8],
                                                    const float* __restrict__ w_sigma, const float* __restrict__ w_rgb2,
                                                    HeadAcc& acc, uint32_t* mscr) {
    uint32_t ra[32], rb[32];
    umma::tmem_ld32(taddr + 0, ra);
    umma::tmem_ld32(taddr + 32, rb);
    umma::tmem_ld_wait();
    if constexpr (NC == 4) {       // the other two loads travel while the first two chunks are consumed
        uint32_t rc[32], rd[32];
        umma::tmem_ld32(taddr + 64, rc);
        umma::tmem_ld32(taddr + 96, rd);
        fwd_epilogue32<0, KIND, STORE, MASK, SM>(ra, bias, a8, w_sigma, w_rgb2, acc, mscr);
        fwd_epilogue32<1, KIND, STORE, MASK, SM>(rb, bias, a8, w_sigma, w_rgb2, acc, mscr);
        umma::tmem_ld_wait();
        fwd_epilogue32<2, KIND, STORE, MASK, SM>(rc, bias, a8, w_sigma, w_rgb2, acc, mscr);
        fwd_epilogue32<3, KIND, STORE, MASK, SM>(rd, bias, a8, w_sigma, w_rgb2, acc, mscr);
    } else {
        fwd_epilogue32<0, KIND, STORE, MASK, SM>(ra, bias, a8, w_sigma, w_rgb2, acc, mscr);
        fwd_epilogue32<1, KIND, STORE, MASK, SM>(rb, bias, a8, w_sigma, w_rgb2, acc, mscr);
    }
}

// ---- warp-uniform pipeline roles ------------------------------------------------------------------
// Both helpers are executed by a WHOLE warp (so addresses / descriptors live in uniform registers);
// a single elected lane issues the asynchronous instructions.

// Stream `n_stages` consecutive weight stage images (the last `n_small` of them half-size) into the ring.
// MC > 1: the CTAs of a cluster of MC consume the SAME stage sequence; ring slot i is filled for all of them by
// CTA (i mod MC) with one multicast bulk copy (w_empty then counts the MMA commits of all MC CTAs), so every
// weight byte leaves L2 once per cluster instead of once per CTA.
template <int MC = 1>
__device__ __forceinline__ void produce_stages(const uint8_t* __restrict__ src, int n_stages, int n_small, uint8_t* sW,
                                               uint64_t* w_full, uint64_t* w_empty, int& stage, uint32_t& ph,
                                               uint32_t rank = 0, unsigned long long* tr_wait = nullptr, uint64_t l2_policy = 0ull) {
    static_assert(MC == 1 || kNumStages % MC == 0, "ring slots map to issuing CTAs by slot index");
    for (int s = 0; s < n_stages; ++s) {
#ifdef CNB_TRACE
        CNB_TR_PTR(tr_wait, umma::mbar_wait(&w_empty[stage], ph ^ 1));
#else
        umma::mbar_wait(&w_empty[stage], ph ^ 1);
#endif
        if (umma::elect_one()) {
            const uint32_t bytes = s < n_stages - n_small ? kSlot : kSlot / 2;
            umma::mbar_arrive_expect_tx(&w_full[stage], bytes);
            if (MC == 1) {
                if (l2_policy) umma::bulk_g2s_hint(sW + stage * kSlot, src + (size_t)s * kSlot, bytes, &w_full[stage], l2_policy);
                else umma::bulk_g2s(sW + stage * kSlot, src + (size_t)s * kSlot, bytes, &w_full[stage]);
            } else if ((uint32_t)(stage & (MC - 1)) == rank) {
                if (l2_policy) umma::bulk_g2s_mcast_hint(sW + stage * kSlot, src + (size_t)s * kSlot, bytes, &w_full[stage], (uint16_t)((1u << MC) - 1u), l2_policy);
                else umma::bulk_g2s_mcast(sW + stage * kSlot, src + (size_t)s * kSlot, bytes, &w_full[stage], (uint16_t)((1u << MC) - 1u));
            }
        }
        __syncwarp();
        if (++stage == kNumStages) { stage = 0; ph ^= 1; }
    }
}

// Issue one GEMM of one tile: D[128 x (128*n_halves)] = A[128 x 64*n_kchunks (+32)] . B^T.
// A = K-blocks 0.. of the tile's operand buffer (+ the PE(viewdir) block); B = ring stages in the
// order (chunk, half).  With two halves the pair of adjacent stages is one N = 256 operand.
template <int MC = 1>
__device__ __forceinline__ void release_slot(uint64_t* bar) {
    if (MC == 1) umma::mma_commit(bar); else umma::mma_commit_mcast(bar, (uint16_t)((1u << MC) - 1u));
}
template <int MC = 1>
__device__ __forceinline__ void issue_gemm(uint32_t a_base, uint32_t d_base, uint8_t* sW, uint64_t* w_full,
                                           uint64_t* w_empty, int n_kchunks, int n_halves, int has_dir, int& stage,
                                           uint32_t& ph, uint64_t* done_bar, unsigned long long* tr_wait = nullptr) {
    const uint64_t dA = umma::make_sdesc(a_base, 16, 1024, umma::SWZ_128B);
    const uint64_t dB = umma::make_sdesc(umma::smem_u32(sW), 16, 1024, umma::SWZ_128B);
    if (n_halves == 2) {
        const uint32_t idesc = umma::make_idesc(128, 256, 0, 0);
        for (int c = 0; c < n_kchunks; ++c) {
#ifdef CNB_TRACE
            CNB_TR_PTR(tr_wait, umma::mbar_wait(&w_full[stage], ph); umma::mbar_wait(&w_full[stage + 1], ph));
#else
            umma::mbar_wait(&w_full[stage], ph);
            umma::mbar_wait(&w_full[stage + 1], ph);
#endif
            umma::tc_fence_after();
            if (umma::elect_one()) {
                const uint64_t da = dA + (uint64_t)((c * kABlock) >> 4);
                const uint64_t db = dB + (uint64_t)((stage * kSlot) >> 4);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma::mma_bf16(d_base, da + ks * 2, db + ks * 2, idesc, (c | ks) ? 1u : 0u);
                release_slot<MC>(&w_empty[stage]);
                release_slot<MC>(&w_empty[stage + 1]);
            }
            __syncwarp();
            stage += 2;
            if (stage == kNumStages) { stage = 0; ph ^= 1; }
        }
    } else {
        const uint32_t idesc = umma::make_idesc(128, 128, 0, 0);
        for (int c = 0; c < n_kchunks; ++c) {
            umma::mbar_wait(&w_full[stage], ph);
            umma::tc_fence_after();
            if (umma::elect_one()) {
                const uint64_t da = dA + (uint64_t)((c * kABlock) >> 4);
                const uint64_t db = dB + (uint64_t)((stage * kSlot) >> 4);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma::mma_bf16(d_base, da + ks * 2, db + ks * 2, idesc, (c | ks) ? 1u : 0u);
                release_slot<MC>(&w_empty[stage]);
            }
            __syncwarp();
            if (++stage == kNumStages) { stage = 0; ph ^= 1; }
        }
    }
    if (has_dir) {
        const uint32_t idesc = umma::make_idesc(128, 128, 0, 0);
        const uint64_t dAd = umma::make_sdesc(a_base + 4 * kABlock, 16, 512, umma::SWZ_64B);
        const uint64_t dBd = umma::make_sdesc(umma::smem_u32(sW), 16, 512, umma::SWZ_64B);
        for (int h = 0; h < n_halves; ++h) {
            umma::mbar_wait(&w_full[stage], ph);
            umma::tc_fence_after();
            if (umma::elect_one()) {
                const uint64_t db = dBd + (uint64_t)((stage * kSlot) >> 4);
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) umma::mma_bf16(d_base + h * 128, dAd + ks * 2, db + ks * 2, idesc, 1u);
                release_slot<MC>(&w_empty[stage]);
            }
            __syncwarp();
            if (++stage == kNumStages) { stage = 0; ph ^= 1; }
        }
    }
    if (umma::elect_one()) umma::mma_commit(done_bar);
    __syncwarp();
}

// ---- CTA-pair (cta_group::2) versions: each CTA streams HALF of every weight chunk (its 128 of the 256
// output columns, or 64 of 128), the leader issues M = 256 MMAs covering one tile of each CTA ------------
// Tensor maps over the packed weight buffer viewed as rows of 128 bytes: boxes of 128 rows (16 KB) / 64 rows (8 KB).
struct alignas(64) WeightMaps { CUtensorMap m16, m8; };

// stages of one GEMM for CTA `rank`: chunk c -> image (c * n_halves + rank) [n_halves == 2], or the
// rank-th 8 KB of image c [n_halves == 1]; the PE(viewdir) chunk -> image (n_main + rank), 8 KB.
// Every load completes on the LEADER's barrier; the leader arms it with the bytes of both CTAs.
__device__ __forceinline__ void produce_stages_2cta(const WeightMaps* maps, uint32_t w_off, int n_kchunks, int n_halves,
                                                    int has_dir, uint32_t rank, uint8_t* sW, uint64_t* w_full,
                                                    uint64_t* w_empty, int& stage, uint32_t& ph, uint64_t l2_policy = 0ull,
                                                    bool skip_fill = false) {
    const int n = n_kchunks + (has_dir ? 1 : 0);
    const uint32_t full0 = umma::mapa(umma::smem_u32(&w_full[0]), 0);
    for (int c = 0; c < n; ++c) {
        umma::mbar_wait(&w_empty[stage], ph ^ 1);
        if (skip_fill) {        // timing experiment (option experiment & 2): the slot keeps whatever it holds
            if (rank == 0 && umma::elect_one()) umma::mbar_arrive(&w_full[stage]);
        } else if (umma::elect_one()) {
            const bool dir = c >= n_kchunks;
            const bool small = dir || n_halves == 1;
            const uint32_t off = dir ? w_off + (uint32_t)(n_kchunks * n_halves + rank) * kSlot
                                     : (n_halves == 2 ? w_off + (uint32_t)(c * 2 + rank) * kSlot : w_off + (uint32_t)c * kSlot + rank * (kSlot / 2));
            if (rank == 0) umma::mbar_arrive_expect_tx(&w_full[stage], small ? kSlot : 2 * kSlot);
            if (l2_policy)
                umma::tma_load_2d_pair_hint(sW + stage * kSlot, small ? (const void*)&maps->m8 : (const void*)&maps->m16, 0, (int)(off >> 7),
                                            full0 + stage * 8, l2_policy);
            else
                umma::tma_load_2d_pair(sW + stage * kSlot, small ? (const void*)&maps->m8 : (const void*)&maps->m16, 0, (int)(off >> 7),
                                       full0 + stage * 8);
        }
        __syncwarp();
        if (++stage == kNumStages) { stage = 0; ph ^= 1; }
    }
}
// leader CTA: issue one GEMM for the pair's two tiles (M = 256)
__device__ __forceinline__ void issue_gemm_2cta(uint32_t a_base, uint32_t d_base, uint8_t* sW, uint64_t* w_full,
                                                uint64_t* w_empty, int n_kchunks, int n_halves,
                                                int has_dir, int& stage, uint32_t& ph, uint64_t* done_bar,
                                                unsigned long long* tr_wait = nullptr) {
    const uint64_t dA = umma::make_sdesc(a_base, 16, 1024, umma::SWZ_128B);
    const uint64_t dB = umma::make_sdesc(umma::smem_u32(sW), 16, 1024, umma::SWZ_128B);
    const uint32_t idesc = umma::make_idesc(256, n_halves * 128, 0, 0);
    for (int c = 0; c < n_kchunks; ++c) {
#ifdef CNB_TRACE
        CNB_TR_PTR(tr_wait, umma::mbar_wait_cluster(&w_full[stage], ph));
#else
        umma::mbar_wait_cluster(&w_full[stage], ph);
#endif
        umma::tc_fence_after();
        if (umma::elect_one()) {
            const uint64_t da = dA + (uint64_t)((c * kABlock) >> 4);
            const uint64_t db = dB + (uint64_t)((stage * kSlot) >> 4);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma::mma_bf16_2cta(d_base, da + ks * 2, db + ks * 2, idesc, (c | ks) ? 1u : 0u);
            umma::mma_commit_2cta(&w_empty[stage], 3);
        }
        __syncwarp();
        if (++stage == kNumStages) { stage = 0; ph ^= 1; }
    }
    if (has_dir) {
        const uint64_t dAd = umma::make_sdesc(a_base + 4 * kABlock, 16, 512, umma::SWZ_64B);
        const uint64_t dBd = umma::make_sdesc(umma::smem_u32(sW), 16, 512, umma::SWZ_64B);
        umma::mbar_wait_cluster(&w_full[stage], ph);
        umma::tc_fence_after();
        if (umma::elect_one()) {
            const uint64_t db = dBd + (uint64_t)((stage * kSlot) >> 4);
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) umma::mma_bf16_2cta(d_base, dAd + ks * 2, db + ks * 2, idesc, 1u);
            umma::mma_commit_2cta(&w_empty[stage], 3);
        }
        __syncwarp();
        if (++stage == kNumStages) { stage = 0; ph ^= 1; }
    }
    if (umma::elect_one()) umma::mma_commit_2cta(done_bar, 3);
    __syncwarp();
}

// PE of one row -- reference src/model.py:4-7 (x, sines, cosines), as packed bf16 pairs in registers.
// sin/cos(2^i x) by exact doubling from an accurate sincosf(x): error < 2^i * 1e-7, far below bf16.
// Computing (registers) and storing (shared memory) are separate so that a tile's encodings can be prepared while
// the previous tile still owns the operand buffer.
struct PeRow { uint32_t x[32]; uint32_t d[16]; };    // xyz: 63 channels (+1 zero); view direction: 27 (+5 zeros)
__device__ __forceinline__ void pe_compute_xyz(const float p[3], bool valid, uint32_t (&out)[32]) {
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
    for (int i = 0; i < 32; ++i) out[i] = umma::pack_bf16(e[2 * i], e[2 * i + 1]);
}
__device__ __forceinline__ void pe_compute_dir(const float v[3], bool valid, uint32_t (&out)[16]) {
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
    for (int i = 0; i < 16; ++i) out[i] = umma::pack_bf16(d[2 * i], d[2 * i + 1]);
}
// row `row` of a [128 x 64] bf16 block, 128-byte swizzle
__device__ __forceinline__ void pe_store_xyz(const uint32_t (&w)[32], uint8_t* blk0, int row) {
#pragma unroll
    for (int ch = 0; ch < 8; ++ch)
        st_shared_v4(blk0 + row * 128 + ((ch ^ (row & 7)) << 4), w[ch * 4 + 0], w[ch * 4 + 1], w[ch * 4 + 2], w[ch * 4 + 3]);
}
// row `row` of a [128 x 32] bf16 block, 64-byte swizzle
__device__ __forceinline__ void pe_store_dir(const uint32_t (&w)[16], uint8_t* dirblk, int row) {
#pragma unroll
    for (int ch = 0; ch < 4; ++ch)
        st_shared_v4(dirblk + row * 64 + ((ch ^ ((row >> 1) & 3)) << 4), w[ch * 4 + 0], w[ch * 4 + 1], w[ch * 4 + 2], w[ch * 4 + 3]);
}
__device__ __forceinline__ void encode_xyz_row(const float p[3], bool valid, uint8_t* blk0, int row) {
    uint32_t w[32];
    pe_compute_xyz(p, valid, w);
    pe_store_xyz(w, blk0, row);
}
__device__ __forceinline__ void encode_dir_row(const float v[3], bool valid, uint8_t* dirblk, int row) {
    uint32_t w[16];
    pe_compute_dir(v, valid, w);
    pe_store_dir(w, dirblk, row);
}
__device__ __forceinline__ void encode_row(const float p[3], const float v[3], bool valid, uint8_t* blk0,
                                           uint8_t* dirblk, int row) {
    encode_xyz_row(p, valid, blk0, row);
    encode_dir_row(v, valid, dirblk, row);
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

// CNB_WEIGHT_MCAST = 1 | 2 | 4: CTAs per cluster sharing one multicast weight stream (full-grid launches only).
inline int weight_multicast(int dflt = 1) {
    const int m = (int)cnb_option("weight_mcast", dflt);      // read per launch: tests switch it inside one process
    return (m == 2 || m == 4) ? m : 1;
}
// Largest grid (a multiple of the cluster size) whose clusters are all co-resident: the kernels are persistent
// and split the work by gridDim, so a second wave of clusters would double the run time.
template <class K>
inline int cluster_grid(K kern, cudaLaunchConfig_t* cfg, int csize, int* grid) {
    int n = 0;
    cfg->gridDim = dim3((unsigned)((*grid / csize) * csize));
    CNB_CUDA_TRY(cudaOccupancyMaxActiveClusters(&n, kern, cfg));
    if (n < 1) return CNB_E_DEVICE;
    int g = *grid / csize;
    if (g > n) g = n;
    *grid = g * csize;
    return CNB_OK;
}

struct FwdWorkspace { float *z, *folded, *rows; float4* samples; size_t bytes; };

inline size_t carve_fwd(const cnb_net_config* c, int n_codes, int64_t spill_samples, char* base, FwdWorkspace* w) {
    size_t off = 0;
    auto take = [&](size_t bytes) -> char* { char* p = base ? base + off : nullptr; off += (bytes + 255) & ~(size_t)255; return p; };
    const int nf = c->shape_blocks + c->texture_blocks;
    FwdWorkspace x = {};
    x.z = (float*)take(sizeof(float) * (size_t)n_codes * nf * kW);
    x.folded = (float*)take(sizeof(float) * (size_t)n_codes * nf * kW);
    x.rows = (float*)take(sizeof(ConstRows));              // staging image of c_rows (ConstRowsScope)
    x.samples = (float4*)take(sizeof(float4) * (size_t)spill_samples);
    x.bytes = off;
    if (w) *w = x;
    return off;
}

// ---- upload of c_rows -------------------------------------------------------------------------------------
struct RowPtrs {
    const float* bias[kCrowSlots];
    const float *w_sigma, *w_rgb2, *b_sigma, *b_rgb2;
    int n_out[kCrowSlots];
};
// constant-memory slot of a layer with a launch-wide bias
__host__ __device__ __forceinline__ int crow_slot(const FwdLayer& L, int l) { return L.kind == 1 ? 1 : L.kind == 2 ? 3 : (l == 0 ? 0 : 2); }
static __global__ void k_gather_const_rows(const RowPtrs r, float* __restrict__ out) {
    const int n = (int)(sizeof(ConstRows) / sizeof(float));
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float v = 0.f;
        if (i < kCrowWsig) { const int l = i / kW, c = i % kW; if (r.bias[l] && c < r.n_out[l]) v = r.bias[l][c]; }
        else if (i < kCrowWrgb) v = r.w_sigma[i - kCrowWsig];
        else if (i < kCrowBsig) v = r.w_rgb2[i - kCrowWrgb];
        else if (i == kCrowBsig) v = r.b_sigma[0];
        else if (i < kCrowBsig + 4) v = r.b_rgb2[i - kCrowBsig - 1];
        out[i] = v;
    }
}
// The constant rows are per-module state shared by every launch of this translation unit's kernels.  A scope
// (one per API call) serialises their users: it takes a host lock, makes the stream wait for the last kernel that read
// the previous contents (possibly on another stream), gathers the rows from the parameter tensors into `staging`
// (workspace) and copies them into c_rows on the stream; its destructor records the "last use" event after the
// call's launches and releases the lock.  Kernels of different streams therefore never see each other's rows.
struct ConstRowsShared { std::mutex mu; cudaEvent_t ev[16] = {}; bool made[16] = {}; };
static ConstRowsShared g_crow;
class ConstRowsScope {
  public:
    ConstRowsScope() = default;
    ConstRowsScope(const ConstRowsScope&) = delete;
    int begin(const cnb_net_config* c, const float* const* P, const Plan& pl, float* staging, cudaStream_t st) {
        g_crow.mu.lock(); locked_ = true; st_ = st;
        if (cudaGetDevice(&dev_) != cudaSuccess || dev_ < 0 || dev_ >= 16) return CNB_E_DEVICE;
        if (!g_crow.made[dev_]) {
            CNB_CUDA_TRY(cudaEventCreateWithFlags(&g_crow.ev[dev_], cudaEventDisableTiming));
            g_crow.made[dev_] = true;
        } else {
            CNB_CUDA_TRY(cudaStreamWaitEvent(st, g_crow.ev[dev_], 0));
        }
        CnbLayout L; cnb_make_layout(c, &L);
        RowPtrs r = {};
        for (int l = 0; l < pl.n_layers; ++l) {
            if (pl.fwd[l].folded >= 0) continue;
            const int slot = crow_slot(pl.fwd[l], l);
            r.bias[slot] = pl.fwd[l].bias;
            r.n_out[slot] = pl.fwd[l].n_halves * 128;
        }
        r.w_sigma = P[L.i_sigma]; r.b_sigma = P[L.i_sigma + 1]; r.w_rgb2 = P[L.i_rgb2]; r.b_rgb2 = P[L.i_rgb2 + 1];
        k_gather_const_rows<<<8, 256, 0, st>>>(r, staging);
        CNB_LAUNCH_CHECK();
        CNB_CUDA_TRY(cudaMemcpyToSymbolAsync(c_rows, staging, sizeof(ConstRows), 0, cudaMemcpyDeviceToDevice, st));
        armed_ = true;
        return CNB_OK;
    }
    ~ConstRowsScope() {
        if (armed_) cudaEventRecord(g_crow.ev[dev_], st_);
        if (locked_) g_crow.mu.unlock();
    }
  private:
    bool locked_ = false, armed_ = false;
    int dev_ = 0;
    cudaStream_t st_ = nullptr;
};

// z_j = ReLU(latent layer_j(code)) and the folded per-code biases (defined in render_sm100.cu).
int latent_and_fold(const cnb_net_config* c, const float* const* P, const float* shape_codes, const float* tex_codes,
                    int n_codes, FwdWorkspace& w, cudaStream_t st);

// Tensor maps of the packed weight buffer (defined in render_sm100.cu); CNB_OK or a status.
int make_weight_maps(const void* packed, size_t bytes, WeightMaps* out);

// K1 over a ray sub-range (defined in render_sm100.cu).
int launch_render_rays(const cnb_net_config* cfg, const float* const* P, const void* packed, const Plan& pl,
                       const cnb_ray_batch* rays, int64_t ray_begin, int64_t ray_count, const float* folded,
                       float* rows_staging, float* spill_sig, float* spill_rgb, float* rgb, float* depth, float* acc,
                       cudaStream_t st);
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
