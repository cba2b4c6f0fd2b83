// probe_pair.cu -- can a CTA's TMA bulk copy (into its OWN shared memory) signal an mbarrier that lives in
// the OTHER CTA of a 2-CTA cluster?  (Needed to feed cta_group::2 MMAs without a forwarding hop.)
#include <cstdio>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../codenerf_b200/csrc/umma.cuh"

__global__ void __launch_bounds__(128) k_pair(const uint8_t* src, uint32_t bytes, int* result) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    const uint32_t rank = umma::cluster_ctarank();
    if (threadIdx.x == 0) { umma::mbar_init(&bar, 1); umma::fence_mbar_init(); }
    umma::cluster_sync_all();
    if (threadIdx.x == 0) {
        if (rank == 0) {
            umma::mbar_arrive_expect_tx(&bar, 2 * bytes);
            umma::bulk_g2s(smem, src, bytes, &bar);
        } else {
            const uint32_t remote_bar = umma::mapa(umma::smem_u32(&bar), 0);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(umma::smem_u32(smem)), "l"(src + bytes), "r"(bytes), "r"(remote_bar) : "memory");
        }
    }
    if (rank == 0) umma::mbar_wait_cluster(&bar, 0);
    umma::cluster_sync_all();
    // verify own shared memory
    int bad = 0;
    for (uint32_t i = threadIdx.x; i < bytes; i += blockDim.x) bad |= smem[i] != src[rank * bytes + i];
    if (bad) atomicExch(&result[rank], 1);
    if (threadIdx.x == 0) atomicAdd(&result[2 + rank], 1);
}

int main() {
    const uint32_t bytes = 16384;
    std::vector<uint8_t> h(2 * bytes);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 7 + 3);
    uint8_t* d; int* r;
    cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    cudaMalloc(&r, 16); cudaMemset(r, 0, 16);
    cudaFuncSetAttribute(k_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes + 1024);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = bytes + 1024;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_pair, (const uint8_t*)d, bytes, r);
    cudaError_t e2 = cudaDeviceSynchronize();
    int hr[4] = {-1, -1, -1, -1};
    cudaMemcpy(hr, r, 16, cudaMemcpyDeviceToHost);
    unsigned int to = 0; cudaMemcpyFromSymbol(&to, umma::g_umma_timeout, sizeof(to));
    printf("PAIR_PROBE launch=%s sync=%s timeout=%u bad0=%d bad1=%d done0=%d done1=%d -> %s\n", cudaGetErrorName(e), cudaGetErrorName(e2),
           to, hr[0], hr[1], hr[2], hr[3], (e == cudaSuccess && e2 == cudaSuccess && !to && !hr[0] && !hr[1] && hr[2] == 1 && hr[3] == 1) ? "REMOTE_SIGNAL_OK" : "REMOTE_SIGNAL_FAIL");
    return 0;
}
