// probe_store.cu -- how fast can the SMs WRITE a stream of 64 KB shared-memory images to global memory?
//
// K2 (k_mlp_bwd) stashes one 64 KB operand image per tile and GEMM (7.6 KB per sample, 32 GB per 65,536-ray step)
// with TMA bulk stores; its pipeline trace says the drain of those stores is what the epilogues wait for.  This probe
// isolates the write path: 148 persistent CTAs, two "groups" per CTA each owning a 64 KB shared-memory buffer, write
// `images` images per group to distinct addresses (HBM stream) or wrapped into a small window (L2 only):
//   mode 0: cp.async.bulk shared -> global, `pieces` bulk stores per image issued by `pieces` lanes (the stash: 2)
//   mode 1: st.global.v4 by `warps` warps per group, coalesced (512 B per instruction), data from registers
//   mode 2: ld.shared.v4 + st.global.v4 by `warps` warps per group (a warp-copied stash)
// Usage: probe_store <mode> <pieces|warps> <images per group> <wrap MB, 0 = none> [evict_first 0|1]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../codenerf_b200/csrc/umma.cuh"

namespace {
constexpr int kImage = 65536;

__global__ void __launch_bounds__(576, 1) k_store(uint8_t* dst, int mode, int n, long long images, long long wrap_images, int hint) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 2 * kImage / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i * 2654435761u;
    umma::fence_proxy_async_smem();
    __syncthreads();
    const uint64_t pol = umma::l2_policy_evict_first();
    const int wpg = mode == 0 ? 1 : n;              // warps per group
    const int g = warp / wpg, wg = warp % wpg;
    if (g >= 2) return;
    const uint8_t* src = smem + g * kImage;
    const long long base = ((long long)blockIdx.x * 2 + g) * images;
    for (long long i = 0; i < images; ++i) {
        long long img = base + i;
        if (wrap_images > 0) img %= wrap_images;
        uint8_t* d = dst + img * kImage;
        if (mode == 0) {
            const uint32_t piece = kImage / n;
            if (lane < n) {
                if (hint) umma::bulk_s2g_hint(d + (size_t)lane * piece, src + (size_t)lane * piece, piece, pol);
                else umma::bulk_s2g(d + (size_t)lane * piece, src + (size_t)lane * piece, piece);
                umma::bulk_commit();
                umma::bulk_wait_read_all();
            }
            __syncwarp();
        } else {
            const int per_warp = kImage / 512 / wpg;        // 512-byte rows of this warp
            uint4 v = make_uint4(lane, (uint32_t)i, 3u, 4u);
            for (int r = 0; r < per_warp; r += 8) {
                uint4 x[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const size_t off = ((size_t)(wg * per_warp + r + u) * 32 + lane) * 16;
                    if (mode == 2) {
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x[u].x), "=r"(x[u].y), "=r"(x[u].z), "=r"(x[u].w)
                                     : "r"(umma::smem_u32(src + off)));
                    } else x[u] = v;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const size_t off = ((size_t)(wg * per_warp + r + u) * 32 + lane) * 16;
                    if (hint) umma::st_global_v4_hint(d + off, x[u], pol);
                    else *reinterpret_cast<uint4*>(d + off) = x[u];
                }
            }
        }
    }
    if (mode == 0 && lane < n) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
}  // namespace

int main(int argc, char** argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    const int n = argc > 2 ? atoi(argv[2]) : 2;
    const long long images = argc > 3 ? atoll(argv[3]) : 800;
    const long long wrap_mb = argc > 4 ? atoll(argv[4]) : 0;
    const int hint = argc > 5 ? atoi(argv[5]) : 1;
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const long long total_images = (long long)sms * 2 * images;
    const long long wrap_images = wrap_mb > 0 ? wrap_mb * 1024 * 1024 / kImage : 0;
    const size_t bytes = (size_t)(wrap_images > 0 ? wrap_images : total_images) * kImage;
    uint8_t* dst = nullptr;
    if (cudaMalloc(&dst, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    const size_t smem = 2 * kImage + 1024;
    cudaFuncSetAttribute(k_store, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int threads = (mode == 0 ? 2 : 2 * n) * 32;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k_store<<<sms, threads, smem>>>(dst, mode, n, images, wrap_images, hint);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    const double gb = (double)total_images * kImage / 1e9;
    printf("mode %d n %d images/group %lld wrap %lld MB hint %d: %.1f GB in %.3f ms = %.0f GB/s (%.1f B/clk/SM at 1.9 GHz)\n", mode, n, images,
           wrap_mb, hint, gb, best, gb / best * 1e3, gb * 1e9 / (best * 1e-3) / sms / 1.9e9);
    cudaFree(dst);
    return 0;
}
