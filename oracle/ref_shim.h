/*
 * TEST INFRASTRUCTURE ONLY.
 * Host shim that lets the reference's CUDA kernel *bodies* (extracted at build
 * time from /root/reference into the git-ignored oracle/_ref/, never
 * committed) compile with plain g++ so they can serve as the ground truth the
 * CPU restatement (gfla_oracle_impl.h) and the CUDA library are checked
 * against.  Everything in this file is ours; no reference text lives here.
 */
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>

#define __global__
#define __restrict__
#define __device__
#define __host__

struct long4 { long x, y, z, w; };
static inline long4 make_long4(long x, long y, long z, long w) { return long4{x, y, z, w}; }

struct ref_dim3 { int x = 0, y = 0, z = 0; };
static thread_local ref_dim3 blockIdx, threadIdx;
static thread_local ref_dim3 blockDim;

using std::exp;
using std::floor;
using std::max;
using std::min;

/* CUDA resolves atomicAdd(float*, <double expr>) to the float overload, i.e.
 * the addend is narrowed to T first; omp atomic keeps multi-threaded timing
 * runs race-free. */
template <typename T, typename U>
static inline void atomicAdd(T* p, U v) {
    T add = static_cast<T>(v);
#pragma omp atomic
    *p += add;
}

/* One "launch": run `body(index)` for every global thread index < n with the
 * reference's <<<ceil(n/256), 256>>> decomposition. */
template <typename F>
static inline void ref_launch(long n, F body) {
#pragma omp parallel for schedule(static)
    for (long index = 0; index < n; ++index) {
        blockDim.x = 256;
        blockIdx.x = (int)(index / 256);
        threadIdx.x = (int)(index % 256);
        body();
    }
}

static inline long4 contig_stride(long a, long b, long c, long d) {
    (void)a;
    return make_long4(b * c * d, c * d, d, 1);
}
