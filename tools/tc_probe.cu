// GPU micro-probes for the tile-kernel building blocks (run on the B200 box):
//   tc_probe tma <rank> <swizzle:0|32|64|128>   one TMA box load, checks the smem image
//   tc_probe mma <N>                            one K=32 (2 x K16) UMMA from SW64 K-major operands, checks D
// Each invocation is one process (a faulting probe must not poison the next).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_bf16.h>
#include "../global-flow-local-attention_b200/csrc/tc_common.cuh"
using namespace gfla::tc;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

__global__ void k_tma(const __grid_constant__ CUtensorMap tm, int rank, int x, int y, int c, int b, uint8_t* dump, int bytes) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar, bytes);
        if (rank == 4) tma_load_4d(smem, &tm, &bar, x, y, c, b);
        else if (rank == 2)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(smem_u32(smem)), "l"((uint64_t)&tm), "r"(smem_u32(&bar)), "r"(x), "r"(c) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(smem_u32(smem)), "l"((uint64_t)&tm), "r"(smem_u32(&bar)), "r"(x), "r"(y), "r"(c) : "memory");
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) dump[i] = smem[i];
}

static int probe_tma(int rank, int swz, int x, int y) {
    const int Ws = 64, Hs = 16, C = 64, B = 2, BWp = (swz == 32 ? 16 : swz == 128 ? 64 : 32), CN = 64;
    std::vector<__nv_bfloat16> h((size_t)B * C * Hs * Ws);
    for (size_t i = 0; i < h.size(); ++i) h[i] = __float2bfloat16((float)(i % 2039));
    __nv_bfloat16* d; CK(cudaMalloc(&d, h.size() * 2)); CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
    PFN_tmapEncodeTiled enc = tmap_encoder();
    if (!enc) { printf("no encoder\n"); return 2; }
    CUtensorMap tm;
    CUtensorMapSwizzle sw = swz == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE : swz == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
    CUresult r;
    const int b = 1, c = 0;
    const __nv_bfloat16* base = d;
    if (rank == 4) {
        cuuint64_t gd[4] = {(cuuint64_t)Ws, (cuuint64_t)Hs, (cuuint64_t)C, (cuuint64_t)B}, gs[3] = {(cuuint64_t)Ws * 2, (cuuint64_t)Hs * Ws * 2, (cuuint64_t)C * Hs * Ws * 2};
        cuuint32_t bx[4] = {(cuuint32_t)BWp, 1, CN, 1}, es[4] = {1, 1, 1, 1};
        r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else if (rank == 3) {
        base = d + (size_t)b * C * Hs * Ws;
        cuuint64_t gd[3] = {(cuuint64_t)Ws, (cuuint64_t)Hs, (cuuint64_t)C}, gs[2] = {(cuuint64_t)Ws * 2, (cuuint64_t)Hs * Ws * 2};
        cuuint32_t bx[3] = {(cuuint32_t)BWp, 1, CN}, es[3] = {1, 1, 1};
        r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {  // 2D view of row y of sample b: (x, c) with stride Hs*Ws between channels
        base = d + (size_t)b * C * Hs * Ws + (size_t)y * Ws;
        cuuint64_t gd[2] = {(cuuint64_t)Ws, (cuuint64_t)C}, gs[1] = {(cuuint64_t)Hs * Ws * 2};
        cuuint32_t bx[2] = {(cuuint32_t)BWp, CN}, es[2] = {1, 1};
        r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    printf("encode rank=%d swizzle=%d -> %d\n", rank, swz, (int)r);
    if (r != CUDA_SUCCESS) return 1;
    const int bytes = BWp * CN * 2;
    uint8_t* dump; CK(cudaMalloc(&dump, bytes));
    k_tma<<<1, 128, bytes + 1024>>>(tm, rank, x, y, c, b, dump, bytes);
    CK(cudaDeviceSynchronize());
    std::vector<uint8_t> img(bytes); CK(cudaMemcpy(img.data(), dump, bytes, cudaMemcpyDeviceToHost));
    const int rowB = BWp * 2, nchunk = rowB / 16;
    const int xorbits = swz == 0 ? 0 : swz == 32 ? 1 : swz == 64 ? 3 : 7;
    int bad = 0;
    for (int cc = 0; cc < CN; ++cc)
        for (int xx = 0; xx < BWp; ++xx) {
            const bool oob = (x + xx < 0) || (x + xx >= Ws) || y < 0 || y >= Hs;
            const size_t gi = oob ? 0 : (((size_t)b * C + c + cc) * Hs + y) * Ws + x + xx;
            const int byte_lin = cc * rowB + xx * 2;
            // hardware swizzle = XOR of the 16B-chunk index with address bits [7,10) (masked to the span)
            const int chunk = (byte_lin >> 4), sw_chunk = chunk ^ ((byte_lin >> 7) & xorbits);
            (void)nchunk;
            const int off = (sw_chunk << 4) | (byte_lin & 15);
            __nv_bfloat16 got; memcpy(&got, &img[off], 2);
            const float want = oob ? 0.f : __bfloat162float(h[gi]);
            if (__bfloat162float(got) != want) { if (bad < 4) printf("  mismatch c=%d x=%d got %g want %g\n", cc, xx, __bfloat162float(got), __bfloat162float(h[gi])); ++bad; }
        }
    printf("tma rank=%d swizzle=%d x=%d y=%d: %s (%d mismatches)\n", rank, swz, x, y, bad ? "LAYOUT MISMATCH" : "OK", bad);
    return bad ? 1 : 0;
}

// ---- one UMMA: D[128 x N] = A[128 x 32] * B[N x 32]^T, operands K-major with 64B rows + 64B swizzle
__global__ void k_mma(const __nv_bfloat16* A, const __nv_bfloat16* Bm, float* D, int N) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = smem;                 // [128][64B]
    uint8_t* sb = smem + 8192;          // [N][64B]
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tslot, 256);
    // fill operands with the canonical swizzle: byte = row*64 + ((chunk ^ ((row>>1)&3))<<4) + within
    for (int i = threadIdx.x; i < 128 * 32; i += blockDim.x) {
        const int row = i / 32, e = i % 32;
        *(__nv_bfloat16*)(sa + row * 64 + ((((e >> 3) ^ ((row >> 1) & 3))) << 4) + (e & 7) * 2) = A[i];
    }
    for (int i = threadIdx.x; i < N * 32; i += blockDim.x) {
        const int row = i / 32, e = i % 32;
        *(__nv_bfloat16*)(sb + row * 64 + ((((e >> 3) ^ ((row >> 1) & 3))) << 4) + (e & 7) * 2) = Bm[i];
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tslot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_f16(128, N, true, false, false);
        for (int h = 0; h < 2; ++h) {
            const uint64_t ad = make_smem_desc(smem_u32(sa) + h * 32, 16, 512, kSwizzle64);
            const uint64_t bd = make_smem_desc(smem_u32(sb) + h * 32, 16, 512, kSwizzle64);
            umma_f16(tbase, ad, bd, idesc, h);
        }
        tc_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    if (warp < 4) {
        for (int cc = 0; cc < N / 32; ++cc) {
            uint32_t v[32];
            tmem_ld_32x32(tbase + ((uint32_t)(warp * 32) << 16) + cc * 32, v);
            tmem_ld_wait();
            for (int i = 0; i < 32; ++i) D[(size_t)(warp * 32 + (threadIdx.x & 31)) * N + cc * 32 + i] = __uint_as_float(v[i]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 256);
}

// ---- same with 32B rows + 32B swizzle, one K=16 UMMA (the layout the forward tile kernel uses)
__global__ void k_mma32(const __nv_bfloat16* A, const __nv_bfloat16* Bm, float* D, int N) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = smem;                 // [128][32B]
    uint8_t* sb = smem + 4096;          // [N][32B]
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tslot, 256);
    for (int i = threadIdx.x; i < 128 * 16; i += blockDim.x) {
        const int row = i / 16, e = i % 16;
        *(__nv_bfloat16*)(sa + row * 32 + ((((e >> 3) ^ ((row >> 2) & 1))) << 4) + (e & 7) * 2) = A[row * 32 + e];
    }
    for (int i = threadIdx.x; i < N * 16; i += blockDim.x) {
        const int row = i / 16, e = i % 16;
        *(__nv_bfloat16*)(sb + row * 32 + ((((e >> 3) ^ ((row >> 2) & 1))) << 4) + (e & 7) * 2) = Bm[row * 32 + e];
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tslot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_f16(128, N, true, false, false);
        umma_f16(tbase, make_smem_desc(smem_u32(sa), 16, 256, kSwizzle32), make_smem_desc(smem_u32(sb), 16, 256, kSwizzle32), idesc, 0);
        tc_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    if (warp < 4) {
        for (int cc = 0; cc < N / 32; ++cc) {
            uint32_t v[32];
            tmem_ld_32x32(tbase + ((uint32_t)(warp * 32) << 16) + cc * 32, v);
            tmem_ld_wait();
            for (int i = 0; i < 32; ++i) D[(size_t)(warp * 32 + (threadIdx.x & 31)) * N + cc * 32 + i] = __uint_as_float(v[i]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 256);
}

static int probe_mma(int N, int kdim = 32) {
    std::vector<__nv_bfloat16> A(128 * 32), Bm((size_t)N * 32);
    std::vector<float> Af(A.size()), Bf(Bm.size());
    srand(1);
    for (size_t i = 0; i < A.size(); ++i) { Af[i] = (float)((rand() % 17) - 8) / 8.f; A[i] = __float2bfloat16(Af[i]); }
    for (size_t i = 0; i < Bm.size(); ++i) { Bf[i] = (float)((rand() % 13) - 6) / 4.f; Bm[i] = __float2bfloat16(Bf[i]); }
    __nv_bfloat16 *dA, *dB; float* dD;
    CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, Bm.size() * 2)); CK(cudaMalloc(&dD, (size_t)128 * N * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, Bm.data(), Bm.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(k_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 + N * 64 + 1024));
    CK(cudaFuncSetAttribute(k_mma32, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 + N * 64 + 1024));
    if (kdim == 16) k_mma32<<<1, 128, 8192 + N * 64 + 1024>>>(dA, dB, dD, N);
    else k_mma<<<1, 128, 8192 + N * 64 + 1024>>>(dA, dB, dD, N);
    CK(cudaDeviceSynchronize());
    std::vector<float> D((size_t)128 * N); CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0; double maxerr = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            float ref = 0; for (int k = 0; k < kdim; ++k) ref += Af[m * 32 + k] * Bf[n * 32 + k];
            const double e = fabs(ref - D[(size_t)m * N + n]); if (e > maxerr) maxerr = e;
            if (e > 1e-3) { if (bad < 4) printf("  D[%d][%d] = %g want %g\n", m, n, D[(size_t)m * N + n], ref); ++bad; }
        }
    printf("mma K=%d N=%d: %s (%d bad, max err %g)\n", kdim, N, bad ? "MISMATCH" : "OK", bad, maxerr);
    return bad ? 1 : 0;
}

int main(int argc, char** argv) {
    if (argc >= 6 && !strcmp(argv[1], "tma")) return probe_tma(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]));
    if (argc >= 3 && !strcmp(argv[1], "mma")) return probe_mma(atoi(argv[2]));
    if (argc >= 3 && !strcmp(argv[1], "mma32")) return probe_mma(atoi(argv[2]), 16);
    printf("usage: tc_probe tma <rank> <swizzle> | mma <N>\n");
    return 2;
}
