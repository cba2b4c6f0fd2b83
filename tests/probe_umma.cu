// probe_umma.cu -- standalone B200 probe for the hand-written tcgen05 building blocks in
// codenerf_b200/csrc/umma.cuh: instruction / shared-memory descriptors (K-major SW128 and
// SW64, MN-major SW128 with LBO), K advance inside a swizzle span, N split, TMEM load
// mapping, tcgen05.commit -> mbarrier, 1-D bulk copies in both directions.
// Build + run by tests/test_gpu_probe.py (GPU box only).  Prints one line per case.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../codenerf_b200/csrc/umma.cuh"

struct MmaOp {
    uint32_t a_off, a_lbo, a_sbo, a_swz;
    uint32_t b_off, b_lbo, b_sbo, b_swz;
    uint32_t idesc, d_col, accumulate;
};
#define MAX_OPS 64
struct ProbeArgs {
    const uint8_t* image; uint32_t image_bytes;
    int n_ops; MmaOp ops[MAX_OPS];
    float* out; int n_cols;          // out [128][n_cols]
    uint8_t* echo;                   // image copied back with bulk_s2g (may be null)
};

__global__ void __launch_bounds__(128) k_probe(ProbeArgs p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar_load, bar_mma;
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { umma::mbar_init(&bar_load, 1); umma::mbar_init(&bar_mma, 1); umma::fence_mbar_init(); }
    if (warp == 0) umma::tmem_alloc(&tmem_base, 512);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tbase = tmem_base;
    if (threadIdx.x == 0) {
        umma::mbar_arrive_expect_tx(&bar_load, p.image_bytes);
        umma::bulk_g2s(smem, p.image, p.image_bytes, &bar_load);
        umma::mbar_wait(&bar_load, 0);
        umma::tc_fence_after();
        const uint32_t sbase = umma::smem_u32(smem);
        for (int i = 0; i < p.n_ops; ++i) {
            const MmaOp& o = p.ops[i];
            const uint64_t da = umma::make_sdesc(sbase + o.a_off, o.a_lbo, o.a_sbo, o.a_swz);
            const uint64_t db = umma::make_sdesc(sbase + o.b_off, o.b_lbo, o.b_sbo, o.b_swz);
            umma::mma_bf16(tbase + o.d_col, da, db, o.idesc, o.accumulate);
        }
        umma::mma_commit(&bar_mma);
        if (p.echo) { umma::bulk_s2g(p.echo, smem, p.image_bytes); umma::bulk_commit(); umma::bulk_wait_all(); }
    }
    __syncwarp();
    umma::mbar_wait(&bar_mma, 0);
    umma::tc_fence_after();
    for (int c0 = 0; c0 < p.n_cols; c0 += 32) {
        uint32_t r[32];
        umma::tmem_ld32(tbase + ((uint32_t)(warp * 32) << 16) + c0, r);
        umma::tmem_ld_wait();
        const int row = warp * 32 + (threadIdx.x & 31);
        for (int j = 0; j < 32; ++j) p.out[(size_t)row * p.n_cols + c0 + j] = __uint_as_float(r[j]);
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tbase, 512);
}

static uint16_t f2bf(float f) {
    uint32_t u; memcpy(&u, &f, 4);
    uint32_t r = u + 0x7FFF + ((u >> 16) & 1);
    return (uint16_t)(r >> 16);
}
static float bf2f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }
static float rnd(uint32_t& s) { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xFFFF) / 65536.0f - 0.5f; }

static int run_case(const char* name, const std::vector<uint8_t>& image, const std::vector<MmaOp>& ops, int n_cols,
                    const std::vector<float>& ref, bool echo) {
    ProbeArgs p = {};
    uint8_t* d_img; float* d_out; uint8_t* d_echo = nullptr;
    cudaMalloc(&d_img, image.size()); cudaMemcpy(d_img, image.data(), image.size(), cudaMemcpyHostToDevice);
    cudaMalloc(&d_out, sizeof(float) * 128 * n_cols); cudaMemset(d_out, 0xFF, sizeof(float) * 128 * n_cols);
    if (echo) { cudaMalloc(&d_echo, image.size()); cudaMemset(d_echo, 0, image.size()); }
    p.image = d_img; p.image_bytes = (uint32_t)image.size(); p.n_ops = (int)ops.size();
    for (size_t i = 0; i < ops.size(); ++i) p.ops[i] = ops[i];
    p.out = d_out; p.n_cols = n_cols; p.echo = d_echo;
    const size_t smem = image.size() + 1024;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_probe<<<1, 128, smem>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned int timeout = 0;
    cudaMemcpyFromSymbol(&timeout, umma::g_umma_timeout, sizeof(timeout));
    std::vector<float> out(128 * n_cols);
    cudaMemcpy(out.data(), d_out, sizeof(float) * out.size(), cudaMemcpyDeviceToHost);
    double max_err = 0, max_ref = 0;
    for (size_t i = 0; i < out.size(); ++i) {
        double d = fabs((double)out[i] - ref[i]);
        if (!(d <= max_err)) max_err = d;     // NaN-propagating
        if (fabs(ref[i]) > max_ref) max_ref = fabs(ref[i]);
    }
    int echo_ok = 1;
    if (echo) {
        std::vector<uint8_t> back(image.size());
        cudaMemcpy(back.data(), d_echo, back.size(), cudaMemcpyDeviceToHost);
        echo_ok = memcmp(back.data(), image.data(), back.size()) == 0;
    }
    const int ok = (e == cudaSuccess) && !timeout && (max_err <= 1e-3 * (max_ref + 1)) && echo_ok;
    printf("PROBE %-28s %s  cuda=%s timeout=%u max_err=%.3e max_ref=%.3e echo=%d\n", name, ok ? "PASS" : "FAIL",
           cudaGetErrorName(e), timeout, max_err, max_ref, echo_ok);
    cudaFree(d_img); cudaFree(d_out); if (d_echo) cudaFree(d_echo);
    return ok;
}

int main() {
    int all = 1;
    uint32_t seed = 12345;
    // ---- case 1/2: K-major SW128, K = 256 in four [rows x 64] blocks ------------------------------
    {
        const int M = 128, N = 256, K = 256;
        std::vector<float> A(M * K), B(N * K);
        for (auto& v : A) v = bf2f(f2bf(rnd(seed)));
        for (auto& v : B) v = bf2f(f2bf(rnd(seed)));
        std::vector<float> ref(M * N, 0.f);
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
            double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k];
            ref[m * N + n] = (float)s;
        }
        // image: A blocks (4 x 16 KB) then B blocks (4 x 32 KB)
        const uint32_t a_base = 0, b_base = 4 * 16384;
        std::vector<uint8_t> img(4 * 16384 + 4 * 32768, 0);
        for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) {
            uint16_t h = f2bf(A[m * K + k]);
            memcpy(&img[a_base + (k / 64) * 16384 + umma::sw128_offset(m, k % 64)], &h, 2);
        }
        for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) {
            uint16_t h = f2bf(B[n * K + k]);
            memcpy(&img[b_base + (k / 64) * 32768 + umma::sw128_offset(n, k % 64)], &h, 2);
        }
        std::vector<MmaOp> ops;
        for (int kb = 0; kb < 4; ++kb) for (int ks = 0; ks < 4; ++ks) {
            MmaOp o = {};
            o.a_off = a_base + kb * 16384 + ks * 32; o.a_lbo = 16; o.a_sbo = 1024; o.a_swz = umma::SWZ_128B;
            o.b_off = b_base + kb * 32768 + ks * 32; o.b_lbo = 16; o.b_sbo = 1024; o.b_swz = umma::SWZ_128B;
            o.idesc = umma::make_idesc(128, 256, 0, 0); o.d_col = 0; o.accumulate = (kb | ks) ? 1 : 0;
            ops.push_back(o);
        }
        all &= run_case("kmajor_sw128_n256", img, ops, N, ref, true);
        // N split: two N=128 MMAs per k-step into column halves (B rows 0..127 / 128..255 of each block)
        ops.clear();
        for (int kb = 0; kb < 4; ++kb) for (int ks = 0; ks < 4; ++ks) for (int h = 0; h < 2; ++h) {
            MmaOp o = {};
            o.a_off = a_base + kb * 16384 + ks * 32; o.a_lbo = 16; o.a_sbo = 1024; o.a_swz = umma::SWZ_128B;
            o.b_off = b_base + kb * 32768 + h * 16384 + ks * 32; o.b_lbo = 16; o.b_sbo = 1024; o.b_swz = umma::SWZ_128B;
            o.idesc = umma::make_idesc(128, 128, 0, 0); o.d_col = h * 128; o.accumulate = (kb | ks) ? 1 : 0;
            ops.push_back(o);
        }
        all &= run_case("kmajor_sw128_nsplit128", img, ops, N, ref, false);
    }
    // ---- case 3: K-major SW64, K = 32 --------------------------------------------------------------
    {
        const int M = 128, N = 256, K = 32;
        std::vector<float> A(M * K), B(N * K);
        for (auto& v : A) v = bf2f(f2bf(rnd(seed)));
        for (auto& v : B) v = bf2f(f2bf(rnd(seed)));
        std::vector<float> ref(M * N, 0.f);
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
            double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k];
            ref[m * N + n] = (float)s;
        }
        const uint32_t a_base = 0, b_base = 8192;
        std::vector<uint8_t> img(8192 + 16384, 0);
        for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) { uint16_t h = f2bf(A[m * K + k]); memcpy(&img[a_base + umma::sw64_offset(m, k)], &h, 2); }
        for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) { uint16_t h = f2bf(B[n * K + k]); memcpy(&img[b_base + umma::sw64_offset(n, k)], &h, 2); }
        std::vector<MmaOp> ops;
        for (int ks = 0; ks < 2; ++ks) {
            MmaOp o = {};
            o.a_off = a_base + ks * 32; o.a_lbo = 16; o.a_sbo = 512; o.a_swz = umma::SWZ_64B;
            o.b_off = b_base + ks * 32; o.b_lbo = 16; o.b_sbo = 512; o.b_swz = umma::SWZ_64B;
            o.idesc = umma::make_idesc(128, 256, 0, 0); o.d_col = 0; o.accumulate = ks ? 1 : 0;
            ops.push_back(o);
        }
        all &= run_case("kmajor_sw64_k32", img, ops, N, ref, false);
    }
    // ---- case 4: MN-major SW128 both operands (the wgrad shape): D[m][n] = sum_r At[r][m] Bt[r][n] ----
    {
        const int M = 128, N = 256, R = 128;   // R = contraction (rows of the stashed tiles)
        std::vector<float> At(R * M), Bt(R * N);
        for (auto& v : At) v = bf2f(f2bf(rnd(seed)));
        for (auto& v : Bt) v = bf2f(f2bf(rnd(seed)));
        std::vector<float> ref(M * N, 0.f);
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
            double s = 0; for (int r = 0; r < R; ++r) s += (double)At[r * M + m] * Bt[r * N + n];
            ref[m * N + n] = (float)s;
        }
        // blocks of [128 rows x 64 cols]: At -> 2 blocks, Bt -> 4 blocks, each 16 KB, same byte image
        // as a K-major activation block (row = r, col = feature)
        const uint32_t a_base = 0, b_base = 2 * 16384;
        std::vector<uint8_t> img(6 * 16384, 0);
        for (int r = 0; r < R; ++r) for (int m = 0; m < M; ++m) { uint16_t h = f2bf(At[r * M + m]); memcpy(&img[a_base + (m / 64) * 16384 + umma::sw128_offset(r, m % 64)], &h, 2); }
        for (int r = 0; r < R; ++r) for (int n = 0; n < N; ++n) { uint16_t h = f2bf(Bt[r * N + n]); memcpy(&img[b_base + (n / 64) * 16384 + umma::sw128_offset(r, n % 64)], &h, 2); }
        std::vector<MmaOp> ops;
        for (int ks = 0; ks < R / 16; ++ks) {
            MmaOp o = {};
            o.a_off = a_base + ks * 2048; o.a_lbo = 16384; o.a_sbo = 1024; o.a_swz = umma::SWZ_128B;
            o.b_off = b_base + ks * 2048; o.b_lbo = 16384; o.b_sbo = 1024; o.b_swz = umma::SWZ_128B;
            o.idesc = umma::make_idesc(128, 256, 1, 1); o.d_col = 0; o.accumulate = ks ? 1 : 0;
            ops.push_back(o);
        }
        all &= run_case("mnmajor_sw128_wgrad", img, ops, N, ref, false);
        // same with N = 64 (a single B block; the enc_xyz weight gradient)
        std::vector<float> ref64(M * 64);
        for (int m = 0; m < M; ++m) for (int n = 0; n < 64; ++n) ref64[m * 64 + n] = ref[m * N + n];
        ops.clear();
        for (int ks = 0; ks < R / 16; ++ks) {
            MmaOp o = {};
            o.a_off = a_base + ks * 2048; o.a_lbo = 16384; o.a_sbo = 1024; o.a_swz = umma::SWZ_128B;
            o.b_off = b_base + ks * 2048; o.b_lbo = 16384; o.b_sbo = 1024; o.b_swz = umma::SWZ_128B;
            o.idesc = umma::make_idesc(128, 64, 1, 1); o.d_col = 0; o.accumulate = ks ? 1 : 0;
            ops.push_back(o);
        }
        all &= run_case("mnmajor_sw128_n64", img, ops, 64, ref64, false);
    }
    // ---- case 5: the wgrad stage layout: 64-row halves of [128 x 64] blocks (LBO = 8192) and an
    //      MN-major 64-B-swizzle operand (the PE(viewdir) block, N = 32) ------------------------------
    {
        const int M = 128, N = 256, R = 64, ND = 32;
        std::vector<float> At(R * M), Bt(R * N), Dt(R * ND);
        for (auto& v : At) v = bf2f(f2bf(rnd(seed)));
        for (auto& v : Bt) v = bf2f(f2bf(rnd(seed)));
        for (auto& v : Dt) v = bf2f(f2bf(rnd(seed)));
        std::vector<float> ref(M * N, 0.f), refd(M * ND, 0.f);
        for (int m = 0; m < M; ++m) {
            for (int n = 0; n < N; ++n) { double s = 0; for (int r = 0; r < R; ++r) s += (double)At[r * M + m] * Bt[r * N + n]; ref[m * N + n] = (float)s; }
            for (int n = 0; n < ND; ++n) { double s = 0; for (int r = 0; r < R; ++r) s += (double)At[r * M + m] * Dt[r * ND + n]; refd[m * ND + n] = (float)s; }
        }
        // 8-KB half blocks: At -> 2, Bt -> 4, then the [64 x 32] SW64 block (4 KB)
        const uint32_t a_base = 0, b_base = 2 * 8192, d_base = 6 * 8192;
        std::vector<uint8_t> img(6 * 8192 + 4096, 0);
        for (int r = 0; r < R; ++r) for (int m = 0; m < M; ++m) { uint16_t h = f2bf(At[r * M + m]); memcpy(&img[a_base + (m / 64) * 8192 + umma::sw128_offset(r, m % 64)], &h, 2); }
        for (int r = 0; r < R; ++r) for (int n = 0; n < N; ++n) { uint16_t h = f2bf(Bt[r * N + n]); memcpy(&img[b_base + (n / 64) * 8192 + umma::sw128_offset(r, n % 64)], &h, 2); }
        for (int r = 0; r < R; ++r) for (int n = 0; n < ND; ++n) { uint16_t h = f2bf(Dt[r * ND + n]); memcpy(&img[d_base + umma::sw64_offset(r, n)], &h, 2); }
        std::vector<MmaOp> ops;
        for (int ks = 0; ks < R / 16; ++ks) {
            MmaOp o = {};
            o.a_off = a_base + ks * 2048; o.a_lbo = 8192; o.a_sbo = 1024; o.a_swz = umma::SWZ_128B;
            o.b_off = b_base + ks * 2048; o.b_lbo = 8192; o.b_sbo = 1024; o.b_swz = umma::SWZ_128B;
            o.idesc = umma::make_idesc(128, 256, 1, 1); o.d_col = 0; o.accumulate = ks ? 1 : 0;
            ops.push_back(o);
        }
        all &= run_case("mnmajor_halfblocks_lbo8k", img, ops, N, ref, false);
        ops.clear();
        for (int ks = 0; ks < R / 16; ++ks) {
            MmaOp o = {};
            o.a_off = a_base + ks * 2048; o.a_lbo = 8192; o.a_sbo = 1024; o.a_swz = umma::SWZ_128B;
            o.b_off = d_base + ks * 1024; o.b_lbo = 4096; o.b_sbo = 512; o.b_swz = umma::SWZ_64B;
            o.idesc = umma::make_idesc(128, 32, 1, 1); o.d_col = 0; o.accumulate = ks ? 1 : 0;
            ops.push_back(o);
        }
        all &= run_case("mnmajor_sw64_n32", img, ops, ND, refd, false);
    }
    printf("PROBE_SUMMARY %s\n", all ? "ALL_PASS" : "SOME_FAIL");
    return all ? 0 : 1;
}
