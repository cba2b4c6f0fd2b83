// probe_pair.cu -- CTA-pair weight feed: can each CTA's TMA load (into its OWN shared memory) complete on an
// mbarrier that lives in the LEADER CTA of a 2-CTA cluster?
//   case A: plain 1-D bulk copy with a remote (mapa) mbarrier address            -> expected NOT to work
//   case B: tensor-map TMA with .cta_group::2 and the leader's mbarrier address  -> the documented 2-SM form
#include <cstdio>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../codenerf_b200/csrc/umma.cuh"

__global__ void __launch_bounds__(128) k_pair(const uint8_t* src, const __grid_constant__ CUtensorMap tmap, int use_tensor,
                                              uint32_t bytes, int* result) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    const uint32_t rank = umma::cluster_ctarank();
    if (threadIdx.x == 0) { umma::mbar_init(&bar, 1); umma::fence_mbar_init(); }
    umma::cluster_sync_all();
    if (threadIdx.x == 0) {
        const uint32_t leader_bar = umma::mapa(umma::smem_u32(&bar), 0);
        if (rank == 0) umma::mbar_arrive_expect_tx(&bar, 2 * bytes);
        if (use_tensor) {
            const int row0 = (int)(rank * (bytes / 128));
            asm volatile(
                "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                ::"r"(umma::smem_u32(smem)), "l"(&tmap), "r"(leader_bar), "r"(0), "r"(row0) : "memory");
        } else {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(umma::smem_u32(smem)), "l"(src + rank * bytes), "r"(bytes), "r"(leader_bar) : "memory");
        }
    }
    if (rank == 0) umma::mbar_wait_cluster(&bar, 0);
    umma::cluster_sync_all();
    int bad = 0;
    for (uint32_t i = threadIdx.x; i < bytes; i += blockDim.x) bad |= smem[i] != src[rank * bytes + i];
    if (bad) atomicExch(&result[rank], 1);
    if (threadIdx.x == 0) atomicAdd(&result[2 + rank], 1);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const uint32_t bytes = 16384;
    std::vector<uint8_t> h(2 * bytes);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 7 + 3);
    uint8_t* d; int* r;
    cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    cudaMalloc(&r, 16);
    void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
    cudaError_t ge = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    CUtensorMap tmap; memset(&tmap, 0, sizeof(tmap));
    CUresult cr = CUDA_ERROR_UNKNOWN;
    if (ge == cudaSuccess && fn) {
        const cuuint64_t dims[2] = {64, (cuuint64_t)(h.size() / 128)};      // [rows][64 bf16]: one row = 128 B
        const cuuint64_t strides[1] = {128};
        const cuuint32_t box[2] = {64, bytes / 128};
        const cuuint32_t estr[2] = {1, 1};
        cr = ((EncodeFn)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    printf("PAIR_PROBE entrypoint=%s encode=%d\n", cudaGetErrorName(ge), (int)cr);
    cudaFuncSetAttribute(k_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes + 1024);
    for (int use_tensor = 0; use_tensor < 2; ++use_tensor) {
        if (use_tensor && cr != CUDA_SUCCESS) break;
        cudaMemset(r, 0, 16);
        const unsigned int zero = 0; cudaMemcpyToSymbol(umma::g_umma_timeout, &zero, sizeof(zero));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = bytes + 1024;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, k_pair, (const uint8_t*)d, tmap, use_tensor, bytes, r);
        cudaError_t e2 = cudaDeviceSynchronize();
        int hr[4] = {-1, -1, -1, -1};
        cudaMemcpy(hr, r, 16, cudaMemcpyDeviceToHost);
        unsigned int to = 0; cudaMemcpyFromSymbol(&to, umma::g_umma_timeout, sizeof(to));
        const bool ok = e == cudaSuccess && e2 == cudaSuccess && !to && !hr[0] && !hr[1] && hr[2] == 1 && hr[3] == 1;
        printf("PAIR_PROBE %s launch=%s sync=%s timeout=%u bad0=%d bad1=%d -> %s\n", use_tensor ? "tensor_cta_group2" : "bulk_remote_mbar",
               cudaGetErrorName(e), cudaGetErrorName(e2), to, hr[0], hr[1], ok ? "REMOTE_SIGNAL_OK" : "REMOTE_SIGNAL_FAIL");
        if (e2 != cudaSuccess) break;
    }
    return 0;
}
