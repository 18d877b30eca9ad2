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

// 16-bit elements, rows % 8 == 0 and cols % 8 == 0, 16-byte aligned pointers: 64 x 64 tiles moved with 128-bit global
// accesses on both sides.  Load: thread <-> (row, 8-column chunk), 4 conflict-free 32-bit shared stores into a 33-word
// pitch tile.  Store: thread <-> (column pair, 8-row chunk): 8 conflict-free 32-bit shared loads (one per row), the two
// 16-bit halves are split with byte permutes into two 16-byte output vectors (one per column).  ~2x the scalar kernel.
__global__ void __launch_bounds__(256)
k_transpose16_vec(const uint4* __restrict__ in, uint4* __restrict__ out, int rows, int cols) {
    __shared__ uint32_t tile[64][33];                      // [row][column pair]
    const long long base = (long long)blockIdx.z * rows * cols;   // in elements
    const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
    const int t = threadIdx.x;
#pragma unroll
    for (int i = 0; i < 2; ++i) {                          // 512 16-byte loads per tile
        const int item = t + i * 256, r = item >> 3, q = item & 7;
        if (r0 + r < rows && c0 + q * 8 < cols) {
            const uint4 v = in[(base + (long long)(r0 + r) * cols + c0 + q * 8) >> 3];
            const int sh = (r >> 5) << 2;                  // rows 32..63 are rotated by 4 words: keeps the column reads below conflict-free
            tile[r][(4 * q + sh) & 31] = v.x; tile[r][(4 * q + 1 + sh) & 31] = v.y;
            tile[r][(4 * q + 2 + sh) & 31] = v.z; tile[r][(4 * q + 3 + sh) & 31] = v.w;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 1; ++i) {                          // 256 items: (row chunk pc of 8, column pair w of 32)
        const int pc = t & 7, w = t >> 3;
        const int c = c0 + 2 * w, r = r0 + pc * 8;
        if (c < cols && r < rows) {
            uint32_t u[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) u[k] = tile[pc * 8 + k][(w + ((pc >> 2) << 2)) & 31];
            uint4 lo, hi;                                  // column c (low halves), column c + 1 (high halves) of rows r .. r+7
            lo.x = __byte_perm(u[0], u[1], 0x5410); lo.y = __byte_perm(u[2], u[3], 0x5410);
            lo.z = __byte_perm(u[4], u[5], 0x5410); lo.w = __byte_perm(u[6], u[7], 0x5410);
            hi.x = __byte_perm(u[0], u[1], 0x7632); hi.y = __byte_perm(u[2], u[3], 0x7632);
            hi.z = __byte_perm(u[4], u[5], 0x7632); hi.w = __byte_perm(u[6], u[7], 0x7632);
            out[(base + (long long)c * rows + r) >> 3] = lo;
            out[(base + (long long)(c + 1) * rows + r) >> 3] = hi;
        }
    }
}

int relayout(const void* src, void* dst, int B, int C, int H, int W, int dtype, int to_nhwc, cudaStream_t st_) {
    const int rows = to_nhwc ? C : H * W, cols = to_nhwc ? H * W : C;
    dim3 grid((unsigned)((cols + 63) / 64), (unsigned)((rows + 63) / 64), (unsigned)B);
    if (elem_size(dtype) == 2 && rows % 8 == 0 && cols % 8 == 0 && aligned(src, 16) && aligned(dst, 16)) {
        k_transpose16_vec<<<grid, 256, 0, st_>>>((const uint4*)src, (uint4*)dst, rows, cols);
        return launch_status();
    }
    switch (elem_size(dtype)) {
        case 2: k_transpose<unsigned short><<<grid, 256, 0, st_>>>((const unsigned short*)src, (unsigned short*)dst, rows, cols); break;
        case 4: k_transpose<unsigned int><<<grid, 256, 0, st_>>>((const unsigned int*)src, (unsigned int*)dst, rows, cols); break;
        case 8: k_transpose<unsigned long long><<<grid, 256, 0, st_>>>((const unsigned long long*)src, (unsigned long long*)dst, rows, cols); break;
        default: return GFLA_E_DTYPE;
    }
    return launch_status();
}

}  // namespace gfla
