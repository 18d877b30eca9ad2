/* TEST INFRASTRUCTURE ONLY: reference resample2d kernels (resample2d_kernel.cu:5-330) built by nvcc. */
#include "ref_cuda_common.cuh"
#include "_ref/resample2d_body.inc"

template <typename T>
static int fwd(const T* a, const T* f, T* o, int B, int C, int Hi, int Wi, int H, int W, int ks, int dil, void* stream) {
    const long n = (long)B * C * H * W;
    REF_FITS_INT(n);
    kernel_resample2d_update_output<T><<<REF_GRID(n)>>>((int)n, a, make_long4(B, C, Hi, Wi), contig_stride(C, Hi, Wi), f,
        make_long4(B, 3, H, W), contig_stride(3, H, W), o, make_long4(B, C, H, W), contig_stride(C, H, W), ks, dil);
    return ref_cuda_status();
}
template <typename T>
static int bwd1(const T* a, const T* f, const T* go, T* g1, int B, int C, int Hi, int Wi, int H, int W, int ks, int dil, void* stream) {
    const long n = (long)B * C * H * W;
    REF_FITS_INT(n);
    kernel_resample2d_backward_input1<T><<<REF_GRID(n)>>>((int)n, a, make_long4(B, C, Hi, Wi), contig_stride(C, Hi, Wi), f,
        make_long4(B, 3, H, W), contig_stride(3, H, W), go, make_long4(B, C, H, W), contig_stride(C, H, W), g1,
        make_long4(B, C, Hi, Wi), contig_stride(C, Hi, Wi), ks, dil);
    return ref_cuda_status();
}
template <typename T>
static int bwd2(const T* a, const T* f, const T* go, T* g2, int B, int C, int Hi, int Wi, int H, int W, int ks, int dil, void* stream) {
    const long n = (long)B * 3 * H * W;
    REF_FITS_INT(n);
    kernel_resample2d_backward_input2<T><<<REF_GRID(n)>>>((int)n, a, make_long4(B, C, Hi, Wi), contig_stride(C, Hi, Wi), f,
        make_long4(B, 3, H, W), contig_stride(3, H, W), go, make_long4(B, C, H, W), contig_stride(C, H, W), g2,
        make_long4(B, 3, H, W), contig_stride(3, H, W), ks, dil);
    return ref_cuda_status();
}
extern "C" {
int refcuda_resample2d_fwd_f32(const float* a, const float* f, float* o, int B, int C, int Hi, int Wi, int H, int W, int ks, int dil, void* st) { return fwd(a, f, o, B, C, Hi, Wi, H, W, ks, dil, st); }
int refcuda_resample2d_fwd_f64(const double* a, const double* f, double* o, int B, int C, int Hi, int Wi, int H, int W, int ks, int dil, void* st) { return fwd(a, f, o, B, C, Hi, Wi, H, W, ks, dil, st); }
int refcuda_resample2d_bwd_input1_f32(const float* a, const float* f, const float* go, float* g1, int B, int C, int Hi, int Wi, int H, int W, int ks, int dil, void* st) { return bwd1(a, f, go, g1, B, C, Hi, Wi, H, W, ks, dil, st); }
int refcuda_resample2d_bwd_input1_f64(const double* a, const double* f, const double* go, double* g1, int B, int C, int Hi, int Wi, int H, int W, int ks, int dil, void* st) { return bwd1(a, f, go, g1, B, C, Hi, Wi, H, W, ks, dil, st); }
int refcuda_resample2d_bwd_input2_f32(const float* a, const float* f, const float* go, float* g2, int B, int C, int Hi, int Wi, int H, int W, int ks, int dil, void* st) { return bwd2(a, f, go, g2, B, C, Hi, Wi, H, W, ks, dil, st); }
int refcuda_resample2d_bwd_input2_f64(const double* a, const double* f, const double* go, double* g2, int B, int C, int Hi, int Wi, int H, int W, int ks, int dil, void* st) { return bwd2(a, f, go, g2, B, C, Hi, Wi, H, W, ks, dil, st); }
}
