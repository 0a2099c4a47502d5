// tools/mma_bench2.cu -- issue-cost microbenchmark, part 2: tcgen05.mma issued from an elect_one() region with the
// descriptors precomputed once and the K loop unrolled (per-step descriptor = base + compile-time constant).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I safe-interactive-crowdnav_b200/csrc -o tools/mma_bench2 tools/mma_bench2.cu
#include <cstdio>
#include <cstdlib>

#include "tc_utils.cuh"

constexpr int TILE = 32768;

template <int N, int TS, int STYLE>   // STYLE 0: `threadIdx.x == 32` | 1: warp-uniform branch + elect_one()
__global__ void __launch_bounds__(128, 1) mma_bench2_kernel(int n_iter, long long *out)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    for (int i = threadIdx.x; i < 4 * TILE / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u + i;
    if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
    if (warp == 0) tc::tmem_alloc<512>(&tmem_slot);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const bool issuer = STYLE == 0 ? threadIdx.x == 32 : (warp == 1 && tc::elect_one());
    if (issuer) {
        const uint64_t da0 = tc::make_smem_desc_sw128(tc::smem_u32(smem), 16, 1024);
        const uint64_t db0 = TS ? tc::make_smem_desc_sw128(tc::smem_u32(smem + 2 * TILE), TILE / 2, 1024)
                                : tc::make_smem_desc_sw128(tc::smem_u32(smem + 2 * TILE), 16, 1024);
        constexpr uint32_t idesc = tc::make_idesc_bf16(128, N, 0, TS ? 1 : 0);
        for (int rep = 0; rep < 3; ++rep) {
            const long long t0 = clock64();
#pragma unroll 1
            for (int it = 0; it < n_iter; ++it) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (TS) tc::umma_ts(tmem, tmem + 256 + k * 8, db0 + (uint64_t)((k * 2048) >> 4), idesc, 1u);
                    else tc::umma_ss(tmem, da0 + (uint64_t)(((k >> 2) * (TILE / 2) + (k & 3) * 32) >> 4),
                                     db0 + (uint64_t)(((k >> 2) * (TILE / 2) + (k & 3) * 32) >> 4), idesc, 1u);
                }
            }
            const long long t1 = clock64();
            tc::umma_commit(&bar);
            tc::mbar_wait(&bar, rep & 1);
            const long long t2 = clock64();
            if (rep == 2 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<512>(tmem);
}

template <int N, int TS, int STYLE> void run(long long *d_out)
{
    const int smem = 4 * TILE + 1024, n_iter = 32;
    long long h[2];
    cudaFuncSetAttribute(mma_bench2_kernel<N, TS, STYLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    mma_bench2_kernel<N, TS, STYLE><<<148, 128, smem>>>(n_iter, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N %d TS %d: %s\n", N, TS, cudaGetErrorString(e)); exit(1); }
    cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
    printf("style %d %s M=128 N=%3d K=16: issue %6.1f clk/MMA, complete %6.1f clk/MMA (ideal %d)\n", STYLE, TS ? "TS(MN-major B)" : "SS            ", N,
           (double)h[0] / (8 * n_iter), (double)h[1] / (8 * n_iter), N / 2);
    fflush(stdout);
}

int main()
{
    long long *d_out;
    cudaMalloc(&d_out, 16);
    run<64, 0, 0>(d_out); run<128, 0, 0>(d_out); run<256, 0, 0>(d_out); run<128, 1, 0>(d_out);
    run<64, 0, 1>(d_out); run<128, 0, 1>(d_out); run<256, 0, 1>(d_out); run<64, 1, 1>(d_out); run<128, 1, 1>(d_out);
    return 0;
}
