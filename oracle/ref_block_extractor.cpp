/* TEST INFRASTRUCTURE ONLY: host driver for the reference block_extractor
 * kernel bodies (block_extractor/block_extractor_kernel.cu:5-170, pulled in
 * from the generated include).  Launch geometry follows :172-217 / :222-278. */
#include "ref_shim.h"
#include <omp.h>
#include "_ref/block_extractor_body.inc"

template <typename T>
static void fwd(const T* s, const T* f, T* o, int B, int C, int Hs, int Ws, int Hf, int Wf, int k) {
    long4 ss = make_long4(B, C, Hs, Ws), fs = make_long4(B, 2, Hf, Wf), os = make_long4(B, C, k * Hf, k * Wf);
    long n = (long)B * C * k * Hf * k * Wf;
    ref_launch(n, [&] {
        kernel_block_extractor_update_output<T>((int)n, s, ss, contig_stride(B, C, Hs, Ws), f, fs,
                                                contig_stride(B, 2, Hf, Wf), o, os,
                                                contig_stride(B, C, k * Hf, k * Wf), k);
    });
}
template <typename T>
static void bwd(const T* s, const T* f, const T* go, T* gs, T* gf, int B, int C, int Hs, int Ws, int Hf, int Wf, int k) {
    long4 ss = make_long4(B, C, Hs, Ws), fs = make_long4(B, 2, Hf, Wf), os = make_long4(B, C, k * Hf, k * Wf);
    long n = (long)B * C * k * Hf * k * Wf;
    ref_launch(n, [&] {
        kernel_block_extractor_backward<T>((int)n, s, ss, contig_stride(B, C, Hs, Ws), f, fs,
                                           contig_stride(B, 2, Hf, Wf), go, os,
                                           contig_stride(B, C, k * Hf, k * Wf), gs, ss,
                                           contig_stride(B, C, Hs, Ws), gf, fs, contig_stride(B, 2, Hf, Wf), k);
    });
}
extern "C" {
/* threads used by every ref_* entry point of this library (one libgomp per process) */
void ref_set_threads(int n) { omp_set_num_threads(n); }
int ref_max_threads(void) { return omp_get_max_threads(); }
void ref_block_extract_fwd_f32(const float* s, const float* f, float* o, int B, int C, int Hs, int Ws, int Hf, int Wf, int k) { fwd(s, f, o, B, C, Hs, Ws, Hf, Wf, k); }
void ref_block_extract_fwd_f64(const double* s, const double* f, double* o, int B, int C, int Hs, int Ws, int Hf, int Wf, int k) { fwd(s, f, o, B, C, Hs, Ws, Hf, Wf, k); }
void ref_block_extract_bwd_f32(const float* s, const float* f, const float* go, float* gs, float* gf, int B, int C, int Hs, int Ws, int Hf, int Wf, int k) { bwd(s, f, go, gs, gf, B, C, Hs, Ws, Hf, Wf, k); }
void ref_block_extract_bwd_f64(const double* s, const double* f, const double* go, double* gs, double* gf, int B, int C, int Hs, int Ws, int Hf, int Wf, int k) { bwd(s, f, go, gs, gf, B, C, Hs, Ws, Hf, Wf, k); }
}
