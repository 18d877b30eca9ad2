// Micro-benchmark: execution rate of tcgen05.mma (kind::f16, M = 128, K = 16) as a function of N, operand majorness /
// swizzle, A in TMEM, and the number of INDEPENDENT accumulators the issue order interleaves -- in isolation
// (one CTA per SM, one issuing thread, operands = whatever bytes sit in shared memory; results are not checked).
// Everything is a compile-time constant and the 16-MMA body is fully unrolled, so the issue loop is as tight as in the kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -o tools/mma_bench tools/mma_bench.cu && tools/mma_bench
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include "../global-flow-local-attention_b200/csrc/tc_common.cuh"
using namespace gfla::tc;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

// a_mode: 0 K-major SW128, 1 K-major SW64 (forward weight slab), 3 MN-major SW64 (backward weight slab), 5 TMEM
// b_mode: 0 K-major SW128 (source rows / grad_out as K-major), 2 MN-major SW128 (source rows / grad_out as MN-major)
template <int A_MODE>
__device__ __forceinline__ uint64_t a_desc(uint32_t base, int k) {
    if (A_MODE == 0) return make_smem_desc(base + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024, kSwizzle128);
    if (A_MODE == 1) return make_smem_desc(base + (k >> 1) * 8192 + (k & 1) * 32, 16, 512, kSwizzle64);
    return make_smem_desc(base + (k & 7) * 1024, 8192, 512, kSwizzle64);
}
template <int B_MODE, int N>
__device__ __forceinline__ uint64_t b_desc(uint32_t base, int k) {
    if (B_MODE == 0) return make_smem_desc(base + (k >> 2) * (N * 128) + (k & 3) * 32, 16, 1024, kSwizzle128);
    return make_smem_desc(base + (k & 7) * 2048, 16384, 1024, kSwizzle128);
}

template <int N, int A_MODE, int B_MODE, int ILP, bool SHARE_B>
__global__ void __launch_bounds__(128, 1) k_mma(int rounds, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i & 0xff);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tslot, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tslot;
    if (warp == 1) {
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 64 * 1024);
        constexpr bool a_mn = A_MODE == 3, b_mn = B_MODE == 2;
        constexpr uint32_t idesc = make_idesc_f16(128, N, true, a_mn, b_mn);
        if (A_MODE == 5 && lane == 0)
            for (int i = 0; i < 16; ++i) tmem_cp_128x256b(tb + 384 + i * 8, a_desc<0>(a0, i));
        __syncwarp();
        const long long t0 = clock64();
        if (elect_one()) {
#pragma unroll 1
            for (int r = 0; r < rounds; ++r) {
#pragma unroll
                for (int k = 0; k < 16; ++k)
#pragma unroll
                    for (int u = 0; u < ILP; ++u) {
                        const uint32_t d = tb + u * N;
                        const uint32_t bb = SHARE_B ? b0 : b0 + u * 8192;
                        if (A_MODE == 5) umma_f16_ts(d, tb + 384 + k * 8, b_desc<B_MODE, N>(bb, k), idesc, k ? 1u : 0u);
                        else umma_f16(d, a_desc<A_MODE>(a0 + u * 32768, k), b_desc<B_MODE, N>(bb, k), idesc, k ? 1u : 0u);
                    }
            }
            tc_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        if (lane == 0 && blockIdx.x == 0) out[0] = t2 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

template <int N, int A_MODE, int B_MODE, int ILP, bool SHARE_B>
static int run(const char* name, long long* d) {
    auto kern = k_mma<N, A_MODE, B_MODE, ILP, SHARE_B>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const int rounds = 64;
    for (int rep = 0; rep < 2; ++rep) {
        kern<<<148, 128, 200 * 1024>>>(rounds, d);
        CK(cudaDeviceSynchronize());
    }
    long long h;
    CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
    const double per = (double)h / (rounds * 16 * ILP);
    printf("%-62s N=%3d ilp=%d%s  %6.1f clk/MMA  %5.2f clk per N column  (math floor %d)\n", name, N, ILP, SHARE_B ? " (shared B)" : "", per, per / N, N / 2);
    return 0;
}

int main() {
    long long* d;
    CK(cudaMalloc(&d, 16));
    // Q stage of the backward: grad_out (K-major) x source rows (K-major)
    run<64, 0, 0, 1, false>("A K-major SW128 x B K-major SW128 [Q stage]", d);
    run<64, 0, 0, 2, false>("A K-major SW128 x B K-major SW128 [Q stage]", d);
    run<64, 0, 0, 4, false>("A K-major SW128 x B K-major SW128 [Q stage]", d);
    run<128, 0, 0, 1, false>("A K-major SW128 x B K-major SW128", d);
    run<128, 0, 0, 2, false>("A K-major SW128 x B K-major SW128", d);
    run<256, 0, 0, 1, false>("A K-major SW128 x B K-major SW128", d);
    // the same with A in TMEM
    run<64, 5, 0, 1, false>("A in TMEM x B K-major SW128", d);
    run<64, 5, 0, 2, false>("A in TMEM x B K-major SW128", d);
    run<64, 5, 0, 4, false>("A in TMEM x B K-major SW128", d);
    run<128, 5, 0, 1, false>("A in TMEM x B K-major SW128", d);
    run<128, 5, 0, 2, false>("A in TMEM x B K-major SW128", d);
    run<256, 5, 0, 1, false>("A in TMEM x B K-major SW128", d);
    // grad_source block of the backward: weight slabs^T (MN-major SW64) x grad_out (MN-major SW128)
    run<64, 3, 2, 1, false>("A MN-major SW64 x B MN-major SW128 [gs block]", d);
    run<128, 3, 2, 1, false>("A MN-major SW64 x B MN-major SW128 [gs block]", d);
    run<128, 3, 2, 2, false>("A MN-major SW64 x B MN-major SW128 [gs block]", d);
    run<256, 3, 2, 1, false>("A MN-major SW64 x B MN-major SW128 [gs block]", d);
    // forward: weight slab (K-major SW64) x source rows (MN-major SW128)
    run<128, 1, 2, 1, false>("A K-major SW64 x B MN-major SW128 [forward step]", d);
    run<256, 1, 2, 1, false>("A K-major SW64 x B MN-major SW128 [forward step]", d);
    run<128, 1, 2, 2, true>("A K-major SW64 x B MN-major SW128 [forward shared step]", d);
    // mixed
    run<128, 0, 2, 1, false>("A K-major SW128 x B MN-major SW128", d);
    run<256, 0, 2, 1, false>("A K-major SW128 x B MN-major SW128", d);
    run<128, 3, 0, 1, false>("A MN-major SW64 x B K-major SW128", d);
    run<256, 3, 0, 1, false>("A MN-major SW64 x B K-major SW128", d);
    return 0;
}
