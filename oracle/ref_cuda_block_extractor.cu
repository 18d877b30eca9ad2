/* TEST INFRASTRUCTURE ONLY: reference block_extractor kernels (block_extractor_kernel.cu:5-170) built by nvcc. */
#include "ref_cuda_common.cuh"
#include "_ref/block_extractor_body.inc"

template <typename T>
static int fwd(const T* s, const T* f, T* o, int B, int C, int Hs, int Ws, int Hf, int Wf, int k, void* stream) {
    const long n = (long)B * C * k * Hf * k * Wf;
    REF_FITS_INT(n);
    kernel_block_extractor_update_output<T><<<REF_GRID(n)>>>((int)n, s, make_long4(B, C, Hs, Ws), contig_stride(C, Hs, Ws), f,
        make_long4(B, 2, Hf, Wf), contig_stride(2, Hf, Wf), o, make_long4(B, C, k * Hf, k * Wf), contig_stride(C, k * Hf, k * Wf), k);
    return ref_cuda_status();
}
template <typename T>
static int bwd(const T* s, const T* f, const T* go, T* gs, T* gf, int B, int C, int Hs, int Ws, int Hf, int Wf, int k, void* stream) {
    const long n = (long)B * C * k * Hf * k * Wf;
    REF_FITS_INT(n);
    kernel_block_extractor_backward<T><<<REF_GRID(n)>>>((int)n, s, make_long4(B, C, Hs, Ws), contig_stride(C, Hs, Ws), f,
        make_long4(B, 2, Hf, Wf), contig_stride(2, Hf, Wf), go, make_long4(B, C, k * Hf, k * Wf), contig_stride(C, k * Hf, k * Wf),
        gs, make_long4(B, C, Hs, Ws), contig_stride(C, Hs, Ws), gf, make_long4(B, 2, Hf, Wf), contig_stride(2, Hf, Wf), k);
    return ref_cuda_status();
}
extern "C" {
int refcuda_block_extract_fwd_f32(const float* s, const float* f, float* o, int B, int C, int Hs, int Ws, int Hf, int Wf, int k, void* st) { return fwd(s, f, o, B, C, Hs, Ws, Hf, Wf, k, st); }
int refcuda_block_extract_fwd_f64(const double* s, const double* f, double* o, int B, int C, int Hs, int Ws, int Hf, int Wf, int k, void* st) { return fwd(s, f, o, B, C, Hs, Ws, Hf, Wf, k, st); }
int refcuda_block_extract_bwd_f32(const float* s, const float* f, const float* go, float* gs, float* gf, int B, int C, int Hs, int Ws, int Hf, int Wf, int k, void* st) { return bwd(s, f, go, gs, gf, B, C, Hs, Ws, Hf, Wf, k, st); }
int refcuda_block_extract_bwd_f64(const double* s, const double* f, const double* go, double* gs, double* gf, int B, int C, int Hs, int Ws, int Hf, int Wf, int k, void* st) { return bwd(s, f, go, gs, gf, B, C, Hs, Ws, Hf, Wf, k, st); }
}
