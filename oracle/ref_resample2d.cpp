/* TEST INFRASTRUCTURE ONLY: host driver for the reference resample2d kernel
 * bodies (resample2d_package/resample2d_kernel.cu:5-330).  Launch geometry
 * follows :335-375 (fwd) and :378-454 (bwd: input1 over B*C*H*W threads, then
 * input2 over B*3*H*W threads). */
#include "ref_shim.h"
#include "_ref/resample2d_body.inc"

template <typename T>
static void fwd(const T* a, const T* f, T* o, int B, int C, int Hi, int Wi, int H, int W, int ks, int dil) {
    long4 as = make_long4(B, C, Hi, Wi), fs = make_long4(B, 3, H, W), os = make_long4(B, C, H, W);
    long n = (long)B * C * H * W;
    ref_launch(n, [&] {
        kernel_resample2d_update_output<T>((int)n, a, as, contig_stride(B, C, Hi, Wi), f, fs,
                                           contig_stride(B, 3, H, W), o, os, contig_stride(B, C, H, W), ks, dil);
    });
}
template <typename T>
static void bwd1(const T* a, const T* f, const T* go, T* g1, int B, int C, int Hi, int Wi, int H, int W, int ks, int dil) {
    long4 as = make_long4(B, C, Hi, Wi), fs = make_long4(B, 3, H, W), os = make_long4(B, C, H, W);
    long n = (long)B * C * H * W;
    ref_launch(n, [&] {
        kernel_resample2d_backward_input1<T>((int)n, a, as, contig_stride(B, C, Hi, Wi), f, fs,
                                             contig_stride(B, 3, H, W), go, os, contig_stride(B, C, H, W), g1, as,
                                             contig_stride(B, C, Hi, Wi), ks, dil);
    });
}
template <typename T>
static void bwd2(const T* a, const T* f, const T* go, T* g2, int B, int C, int Hi, int Wi, int H, int W, int ks, int dil) {
    long4 as = make_long4(B, C, Hi, Wi), fs = make_long4(B, 3, H, W), os = make_long4(B, C, H, W);
    long n = (long)B * 3 * H * W;
    ref_launch(n, [&] {
        kernel_resample2d_backward_input2<T>((int)n, a, as, contig_stride(B, C, Hi, Wi), f, fs,
                                             contig_stride(B, 3, H, W), go, os, contig_stride(B, C, H, W), g2, fs,
                                             contig_stride(B, 3, H, W), ks, dil);
    });
}
extern "C" {
void ref_resample2d_fwd_f32(const float* a, const float* f, float* o, int B, int C, int Hi, int Wi, int H, int W, int ks, int dil) { fwd(a, f, o, B, C, Hi, Wi, H, W, ks, dil); }
void ref_resample2d_fwd_f64(const double* a, const double* f, double* o, int B, int C, int Hi, int Wi, int H, int W, int ks, int dil) { fwd(a, f, o, B, C, Hi, Wi, H, W, ks, dil); }
void ref_resample2d_bwd_input1_f32(const float* a, const float* f, const float* go, float* g1, int B, int C, int Hi, int Wi, int H, int W, int ks, int dil) { bwd1(a, f, go, g1, B, C, Hi, Wi, H, W, ks, dil); }
void ref_resample2d_bwd_input1_f64(const double* a, const double* f, const double* go, double* g1, int B, int C, int Hi, int Wi, int H, int W, int ks, int dil) { bwd1(a, f, go, g1, B, C, Hi, Wi, H, W, ks, dil); }
void ref_resample2d_bwd_input2_f32(const float* a, const float* f, const float* go, float* g2, int B, int C, int Hi, int Wi, int H, int W, int ks, int dil) { bwd2(a, f, go, g2, B, C, Hi, Wi, H, W, ks, dil); }
void ref_resample2d_bwd_input2_f64(const double* a, const double* f, const double* go, double* g2, int B, int C, int Hi, int Wi, int H, int W, int ks, int dil) { bwd2(a, f, go, g2, B, C, Hi, Wi, H, W, ks, dil); }
}
