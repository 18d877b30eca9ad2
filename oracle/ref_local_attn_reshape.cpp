/* TEST INFRASTRUCTURE ONLY: host driver for the reference local_attn_reshape
 * kernel bodies (local_attn_reshape/local_attn_reshape_kernel.cu:5-108).
 * Launch geometry follows :110-148 / :153-195. */
#include "ref_shim.h"
#include "_ref/local_attn_reshape_body.inc"

template <typename T>
static void fwd(const T* in, T* out, int B, int H, int W, int k) {
    long4 is = make_long4(B, k * k, H, W), os = make_long4(B, 1, k * H, k * W);
    long n = (long)B * k * H * k * W;
    ref_launch(n, [&] {
        kernel_local_attn_reshape_update_output<T>((int)n, in, is, contig_stride(B, k * k, H, W), out, os,
                                                   contig_stride(B, 1, k * H, k * W), k);
    });
}
template <typename T>
static void bwd(const T* in, const T* go, T* gi, int B, int H, int W, int k) {
    long4 is = make_long4(B, k * k, H, W), os = make_long4(B, 1, k * H, k * W);
    long n = (long)B * k * H * k * W;
    ref_launch(n, [&] {
        kernel_local_attn_reshape_backward<T>((int)n, in, is, contig_stride(B, k * k, H, W), go, os,
                                              contig_stride(B, 1, k * H, k * W), gi, is,
                                              contig_stride(B, k * k, H, W), k);
    });
}
extern "C" {
void ref_attn_reshape_fwd_f32(const float* in, float* out, int B, int H, int W, int k) { fwd(in, out, B, H, W, k); }
void ref_attn_reshape_fwd_f64(const double* in, double* out, int B, int H, int W, int k) { fwd(in, out, B, H, W, k); }
void ref_attn_reshape_bwd_f32(const float* in, const float* go, float* gi, int B, int H, int W, int k) { bwd(in, go, gi, B, H, W, k); }
void ref_attn_reshape_bwd_f64(const double* in, const double* go, double* gi, int B, int H, int W, int k) { bwd(in, go, gi, B, H, W, k); }
}
