/* TEST INFRASTRUCTURE ONLY: reference local_attn_reshape kernels (local_attn_reshape_kernel.cu:5-108) built by nvcc. */
#include "ref_cuda_common.cuh"
#include "_ref/local_attn_reshape_body.inc"

template <typename T>
static int fwd(const T* in, T* out, int B, int H, int W, int k, void* stream) {
    const long n = (long)B * k * H * k * W;
    REF_FITS_INT(n);
    kernel_local_attn_reshape_update_output<T><<<REF_GRID(n)>>>((int)n, in, make_long4(B, k * k, H, W), contig_stride(k * k, H, W), out,
        make_long4(B, 1, k * H, k * W), contig_stride(1, k * H, k * W), k);
    return ref_cuda_status();
}
template <typename T>
static int bwd(const T* in, const T* go, T* gi, int B, int H, int W, int k, void* stream) {
    const long n = (long)B * k * H * k * W;
    REF_FITS_INT(n);
    kernel_local_attn_reshape_backward<T><<<REF_GRID(n)>>>((int)n, in, make_long4(B, k * k, H, W), contig_stride(k * k, H, W), go,
        make_long4(B, 1, k * H, k * W), contig_stride(1, k * H, k * W), gi, make_long4(B, k * k, H, W), contig_stride(k * k, H, W), k);
    return ref_cuda_status();
}
extern "C" {
int refcuda_attn_reshape_fwd_f32(const float* in, float* out, int B, int H, int W, int k, void* st) { return fwd(in, out, B, H, W, k, st); }
int refcuda_attn_reshape_fwd_f64(const double* in, double* out, int B, int H, int W, int k, void* st) { return fwd(in, out, B, H, W, k, st); }
int refcuda_attn_reshape_bwd_f32(const float* in, const float* go, float* gi, int B, int H, int W, int k, void* st) { return bwd(in, go, gi, B, H, W, k, st); }
int refcuda_attn_reshape_bwd_f64(const double* in, const double* go, double* gi, int B, int H, int W, int k, void* st) { return bwd(in, go, gi, B, H, W, k, st); }
}
