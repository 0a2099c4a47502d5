// tools/mma_bench.cu -- microbenchmark: how many SM clocks does one tcgen05.mma (kind::f16, bf16 in, fp32 out) take for the
// shapes the attention kernel issues?  One CTA per SM, one thread issues `n` MMAs back to back, clock64() around
// issue + completion (tcgen05.commit -> mbarrier).  Operands are whatever is in shared memory (timing only).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I safe-interactive-crowdnav_b200/csrc -o tools/mma_bench tools/mma_bench.cu
#include <cstdio>
#include <cstdlib>

#include "tc_utils.cuh"

constexpr int TILE = 32768;

// mode: 0 SS same D | 1 SS two alternating D | 2 TS (A in TMEM), B MN-major, same D | 3 TS, B K-major, same D
//       4 TS MN-major alternating D | 5 SS alternating operands A/B tiles too (no operand reuse)
// ISSUE: 0 = `if (threadIdx.x == 32)` around the loop | 1 = `if (warp == 1 && elect_one())` around the loop | 2 = whole warp 1 runs the
// loop, elect_one() per MMA
template <int ISSUE>
__global__ void __launch_bounds__(128, 1) mma_bench_kernel(int mode, int N, int n_mma, long long *out, int alt)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 6 * TILE / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u + i;
    if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
    if (warp == 0) tc::tmem_alloc<512>(&tmem_slot);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    bool issuer;
    if (ISSUE == 0) issuer = threadIdx.x == 32;
    else if (ISSUE == 1) issuer = warp == 1 && tc::elect_one();
    else issuer = __shfl_sync(0xffffffffu, warp, 0) == 1;
    if (issuer) {
        const uint32_t a_addr = tc::smem_u32(smem), b_addr = tc::smem_u32(smem + 2 * TILE);
        const uint32_t idesc_ss = tc::make_idesc_bf16(128, (uint32_t)N, 0, 0);
        const uint32_t idesc_mn = tc::make_idesc_bf16(128, (uint32_t)N, 0, 1);
        for (int rep = 0; rep < 3; ++rep) {
            const long long t0 = clock64();
            for (int i = 0; i < n_mma; ++i) {
                const int k = i & 7;
                const uint32_t off = (k >> 2) * (TILE / 2) + (k & 3) * 32;
                const uint32_t d = tmem + (((mode == 1 || mode == 4 || mode == 5) && (i & alt)) ? 256u : 0u);
                const uint32_t tile_sel = (mode == 5 && (i & 8)) ? TILE : 0;
                if (ISSUE == 2) __syncwarp();
                if (ISSUE == 2 && !tc::elect_one()) { } else
                if (mode == 0 || mode == 1 || mode == 5) {
                    tc::umma_ss(d, tc::make_smem_desc_sw128(a_addr + tile_sel + off, 16, 1024),
                                tc::make_smem_desc_sw128(b_addr + tile_sel + off, 16, 1024), idesc_ss, 1u);
                } else if (mode == 2 || mode == 4) {
                    tc::umma_ts(d, tmem + 128 + k * 8, tc::make_smem_desc_sw128(b_addr + k * 2048, TILE / 2, 1024), idesc_mn, 1u);
                } else {
                    tc::umma_ts(d, tmem + 128 + k * 8, tc::make_smem_desc_sw128(b_addr + off, 16, 1024), idesc_ss, 1u);
                }
            }
            const long long t1 = clock64();
            if (ISSUE != 2 || tc::elect_one()) tc::umma_commit(&bar);
            if (ISSUE == 2) __syncwarp();
            tc::mbar_wait(&bar, rep & 1);
            const long long t2 = clock64();
            if (rep == 2 && blockIdx.x == 0 && (threadIdx.x & 31) == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<512>(tmem);
}

int main(int argc, char **argv)
{
    const int only = argc > 1 ? atoi(argv[1]) : -1;
    long long *d_out, h[2];
    cudaMalloc(&d_out, 16);
    const int smem = 6 * TILE + 1024;
    const char *names[] = {"SS same D", "SS alternating D", "TS B MN-major same D", "TS B K-major same D", "TS B MN-major alternating D",
                           "SS alternating D + operands"};
    cudaFuncSetAttribute(mma_bench_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(mma_bench_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(mma_bench_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int issue = 0; issue < 3; ++issue) if (only < 0 || only == issue)
    for (int mode : {0, 1, 4})
        for (int N : {64, 128, 256})
            for (int alt : {1, 2, 8}) {
                const int grid = 148;
                if (mode == 0 && alt != 1) continue;
                const int n = 256;
                printf("issue-style %d alt-every %d ", issue, alt); fflush(stdout);
                if (issue == 0) mma_bench_kernel<0><<<grid, 128, smem>>>(mode, N, n, d_out, alt);
                else if (issue == 1) mma_bench_kernel<1><<<grid, 128, smem>>>(mode, N, n, d_out, alt);
                else mma_bench_kernel<2><<<grid, 128, smem>>>(mode, N, n, d_out, alt);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("mode %d N %d: %s\n", mode, N, cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
                printf("%-30s M=128 N=%3d K=16 grid=%3d: issue %6.1f clk/MMA, complete %6.1f clk/MMA (ideal %d)\n", names[mode], N, grid,
                       (double)h[0] / n, (double)h[1] / n, N / 2); fflush(stdout);
            }
    return 0;
}
