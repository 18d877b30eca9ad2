// NCHW <-> NHWC (channels-last) re-layout of a [B,C,H,W] feature tensor: per sample a [C][H*W] <-> [H*W][C]
// matrix transpose through a padded shared-memory tile, coalesced on both sides.  Used by the Python layer
// when a planar (NCHW) bf16 caller is routed through the channels-last tile kernels.
#include "common.cuh"

namespace gfla {

template <typename T>
__global__ void __launch_bounds__(256)
k_transpose(const T* __restrict__ in, T* __restrict__ out, int rows, int cols) {
    // in: [batch][rows][cols] -> out: [batch][cols][rows]; 64x64 tiles, 256 threads (64 x 4)
    __shared__ T tile[64][64 + 2];
    const long long base = (long long)blockIdx.z * rows * cols;
    const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
#pragma unroll
    for (int i = ty; i < 64; i += 4) {
        const int r = r0 + i, c = c0 + tx;
        if (r < rows && c < cols) tile[i][tx] = in[base + (long long)r * cols + c];
    }
    __syncthreads();
#pragma unroll
    for (int i = ty; i < 64; i += 4) {
        const int c = c0 + i, r = r0 + tx;
        if (r < rows && c < cols) out[base + (long long)c * rows + r] = tile[tx][i];
    }
}

int relayout(const void* src, void* dst, int B, int C, int H, int W, int dtype, int to_nhwc, cudaStream_t st_) {
    const int rows = to_nhwc ? C : H * W, cols = to_nhwc ? H * W : C;
    dim3 grid((unsigned)((cols + 63) / 64), (unsigned)((rows + 63) / 64), (unsigned)B);
    switch (elem_size(dtype)) {
        case 2: k_transpose<unsigned short><<<grid, 256, 0, st_>>>((const unsigned short*)src, (unsigned short*)dst, rows, cols); break;
        case 4: k_transpose<unsigned int><<<grid, 256, 0, st_>>>((const unsigned int*)src, (unsigned int*)dst, rows, cols); break;
        case 8: k_transpose<unsigned long long><<<grid, 256, 0, st_>>>((const unsigned long long*)src, (unsigned long long*)dst, rows, cols); break;
        default: return GFLA_E_DTYPE;
    }
    return launch_status();
}

}  // namespace gfla
