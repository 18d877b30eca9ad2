// resample2d (Gaussian-weighted ks x ks warp, FlowNet2-derived) for sm_100a.
//
// Arithmetic contract: resample2d_package/resample2d_kernel.cu:20-95 (forward),
// :98-202 (grad input1), :204-330 (grad input2 = d/d(dx, dy, sigma)).
// Reference quirks that are reproduced on purpose:
//   * SAFE_DIV(a,b) = b==0 ? a/1e-8 : a/b with a *double* 1e-8 (:14-15): for
//     float tensors the quotient is formed in float, widened, and exp() runs in
//     double before being narrowed -- done here the same way, but ONCE PER
//     PIXEL instead of once per (pixel, channel);
//   * grad_input1 takes its Gaussian weights from alpha = xf - int(xf)
//     (truncation, :137-138) while the tap indices use floor(xf);
//   * all three planes of grad_input2 (including d/dsigma) are produced.
// Layout on the machine (not the reference's):
//   * one thread per output PIXEL (b,y,x); the 4*(ks/2)^2 weights and clamped
//     tap offsets live in registers and are reused for every channel;
//     a CTA owns a 32 x 4 (grad_input1: 32 x 8) pixel tile: a warp is a 128-byte
//     row segment, and the ks rows a pixel row reads are shared with the rows
//     above / below through L1 (one-row CTAs pulled every source row ks times
//     over the L2 -> L1 fabric: 6.4x the tensor, profiles/r2_resample2d.md);
//     channel slices in grid.y for small images;
//   * grad_input1 (fp32): the CTA accumulates the scatter of its tile in a shared-memory
//     box around the tile's footprint and flushes the box with one coalesced red.global
//     per element -- (2*ks/2)^2 global atomics per (pixel, channel) become ~3;
//   * grad_input2: the reference runs 3*H*W threads that each stride twice
//     through all C channel planes; here one thread per pixel accumulates the
//     4*(ks/2)^2 corner dot products sum_c g[c]*v[c,corner] in ONE pass and
//     derives the three gradients from them.
// Compiled with -fmad=false: the fp32/fp64 forward is bit-identical to the
// (uncontracted) CPU oracle up to the last-ulp behaviour of exp().
#include "common.cuh"

namespace gfla {

template <typename A>
__device__ __forceinline__ double safe_div(A a, A b) {  // the reference macro, same typing
    return (b == static_cast<A>(0)) ? (static_cast<double>(a) / 1e-8) : static_cast<double>(a / b);
}

// per-pixel quantities shared by the three kernels
template <typename A, int NT>
struct RsTaps {
    int off[NT * NT * 4];   // clamped tap offsets y*Wi+x, order per (fy,fx): TL, TR, BL, BR
    A xL_[NT], xR_[NT], yT_[NT], yB_[NT];          // distances
    A xL_P[NT], xR_P[NT], yT_P[NT], yB_P[NT];      // Gaussian factors (depend on fx resp. fy only)
    A sigma;
    int flx, fly;           // floor(x + dx), floor(y + dy)
};

// CTA = 32 x TH pixel tile of one sample; blockIdx.x enumerates (sample, tile row, tile column)
struct RsPixel { int x, y, b; bool active; };
template <int TH>
__device__ __forceinline__ RsPixel rs_pixel(int H, int W) {
    const int tiles_x = (W + 31) >> 5, tiles_y = (H + TH - 1) / TH;
    unsigned t = blockIdx.x;
    const int tx = (int)(t % (unsigned)tiles_x); t /= (unsigned)tiles_x;
    const int ty = (int)(t % (unsigned)tiles_y);
    RsPixel p;
    p.b = (int)(t / (unsigned)tiles_y);
    p.x = tx * 32 + (int)(threadIdx.x & 31);
    p.y = ty * TH + (int)(threadIdx.x >> 5);
    p.active = p.x < W && p.y < H;
    return p;
}
template <int TH>
static inline long long rs_tiles(int B, int H, int W) { return (long long)B * ((H + TH - 1) / TH) * ((W + 31) >> 5); }

template <typename A, int NT>
__device__ __forceinline__ void rs_setup(RsTaps<A, NT>& t, const A* __restrict__ in2, int b, int y, int x, int H, int W,
                                         int Hi, int Wi, int dil, bool trunc_frac) {
    const long long hw = (long long)H * W;
    const A* p = in2 + (long long)b * 3 * hw + (long long)y * W + x;
    const A dx = p[0], dy = p[hw];
    t.sigma = p[2 * hw];
    const A xf = static_cast<A>(x) + dx, yf = static_cast<A>(y) + dy;
    const A alpha = trunc_frac ? xf - static_cast<A>(static_cast<int>(xf)) : xf - flr(xf);
    const A beta = trunc_frac ? yf - static_cast<A>(static_cast<int>(yf)) : yf - flr(yf);
    t.flx = static_cast<int>(flr(xf));
    t.fly = static_cast<int>(flr(yf));
    const A two_s2 = 2 * t.sigma * t.sigma;
#pragma unroll
    for (int f = 0; f < NT; ++f) {
        t.xL_[f] = static_cast<A>(f * dil) + alpha;
        t.xR_[f] = static_cast<A>((1. + f) * dil) - alpha;
        t.yT_[f] = static_cast<A>(f * dil) + beta;
        t.yB_[f] = static_cast<A>((1. + f) * dil) - beta;
        t.xL_P[f] = static_cast<A>(exp(safe_div<A>(-t.xL_[f] * t.xL_[f], two_s2)));
        t.xR_P[f] = static_cast<A>(exp(safe_div<A>(-t.xR_[f] * t.xR_[f], two_s2)));
        t.yT_P[f] = static_cast<A>(exp(safe_div<A>(-t.yT_[f] * t.yT_[f], two_s2)));
        t.yB_P[f] = static_cast<A>(exp(safe_div<A>(-t.yB_[f] * t.yB_[f], two_s2)));
    }
#pragma unroll
    for (int fy = 0; fy < NT; ++fy) {
        const int yT = clampi(static_cast<int>(flr(yf) - fy * dil), Hi - 1);
        const int yB = clampi(static_cast<int>(flr(yf) + (fy + 1) * dil), Hi - 1);
#pragma unroll
        for (int fx = 0; fx < NT; ++fx) {
            const int xL = clampi(static_cast<int>(flr(xf) - fx * dil), Wi - 1);
            const int xR = clampi(static_cast<int>(flr(xf) + (fx + 1) * dil), Wi - 1);
            int* o = t.off + (fy * NT + fx) * 4;
            o[0] = yT * Wi + xL; o[1] = yT * Wi + xR; o[2] = yB * Wi + xL; o[3] = yB * Wi + xR;
        }
    }
}

// sum of the 4*NT*NT weights in the reference's order (:80-92)
template <typename A, int NT>
__device__ __forceinline__ A rs_weight_sum(const RsTaps<A, NT>& t) {
    A sum = static_cast<A>(0);
#pragma unroll
    for (int fy = 0; fy < NT; ++fy)
#pragma unroll
        for (int fx = 0; fx < NT; ++fx)
            sum += (t.yT_P[fy] * t.xL_P[fx] + t.yT_P[fy] * t.xR_P[fx] + t.yB_P[fy] * t.xL_P[fx] + t.yB_P[fy] * t.xR_P[fx]);
    return sum;
}

template <typename A, int NT>
__global__ void __launch_bounds__(128)
k_resample2d_fwd(const A* __restrict__ in1, const A* __restrict__ in2, A* __restrict__ out, int B, int C, int Hi, int Wi,
                 int H, int W, int dil, int c_per_slice) {
    const RsPixel px = rs_pixel<4>(H, W);
    if (!px.active) return;
    const int x = px.x, y = px.y, b = px.b;
    RsTaps<A, NT> t;
    rs_setup<A, NT>(t, in2, b, y, x, H, W, Hi, Wi, dil, false);
    A w[NT * NT * 4];
#pragma unroll
    for (int fy = 0; fy < NT; ++fy)
#pragma unroll
        for (int fx = 0; fx < NT; ++fx) {
            A* q = w + (fy * NT + fx) * 4;
            q[0] = t.yT_P[fy] * t.xL_P[fx]; q[1] = t.yT_P[fy] * t.xR_P[fx];
            q[2] = t.yB_P[fy] * t.xL_P[fx]; q[3] = t.yB_P[fy] * t.xR_P[fx];
        }
    const A sum = rs_weight_sum<A, NT>(t);
    const long long ipl = (long long)Hi * Wi, opl = (long long)H * W;
    const int c0 = blockIdx.y * c_per_slice, c1 = min(C, c0 + c_per_slice);
    const A* s = in1 + ((long long)b * C + c0) * ipl;
    A* o = out + ((long long)b * C + c0) * opl + (long long)y * W + x;
#pragma unroll 4
    for (int c = c0; c < c1; ++c, s += ipl, o += opl) {   // unrolled: 4 channels x taps of independent loads in flight
        A val = static_cast<A>(0);
#pragma unroll
        for (int q = 0; q < NT * NT * 4; ++q) val += w[q] * s[t.off[q]];
        *o = static_cast<A>(safe_div<A>(val, sum));
    }
}

// grad_input1: a scatter of 4*(ks/2)^2 weighted copies of grad_out per (pixel, channel) (:180-199).  A CTA owns a
// 32 x 8 pixel tile.  Its taps fall into the box [min floor - (NT-1)*dil, max floor + NT*dil] of the tile's flow (clamped to
// the image like the taps themselves); when that box fits RS_BOX_W x RS_BOX_H (any flow that does not tear the tile apart) the
// scatter of RS_CH channels goes into shared memory and the box is flushed with ONE red.global per element, a warp per box row.
// Otherwise (and for double) every tap is a global atomic, as in the reference.
constexpr int RS_TH1 = 8, RS_BOX_W = 64, RS_BOX_H = 24, RS_CH = 4;

template <typename A, int NT>
__global__ void __launch_bounds__(32 * RS_TH1)
k_resample2d_bwd_in1(const A* __restrict__ in2, const A* __restrict__ gout, A* __restrict__ gin1, int B, int C, int Hi,
                     int Wi, int H, int W, int dil, int c_per_slice) {
    constexpr bool kBox = sizeof(A) == 4;
    __shared__ float box[kBox ? RS_CH * RS_BOX_H * RS_BOX_W : 1];
    __shared__ int ext[4][RS_TH1];
    const RsPixel px = rs_pixel<RS_TH1>(H, W);
    const bool active = px.active;
    const int x = min(px.x, W - 1), y = min(px.y, H - 1), b = px.b;   // inactive lanes stay alive for the CTA barriers
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    RsTaps<A, NT> t;
    rs_setup<A, NT>(t, in2, b, y, x, H, W, Hi, Wi, dil, true);  // truncating fraction for the weights
    const A sum = rs_weight_sum<A, NT>(t);
    double wn[NT * NT * 4];  // SAFE_DIV(w, sum), kept in double like the reference expression (:195-198)
#pragma unroll
    for (int fy = 0; fy < NT; ++fy)
#pragma unroll
        for (int fx = 0; fx < NT; ++fx) {
            double* q = wn + (fy * NT + fx) * 4;
            q[0] = safe_div<A>(t.yT_P[fy] * t.xL_P[fx], sum); q[1] = safe_div<A>(t.yT_P[fy] * t.xR_P[fx], sum);
            q[2] = safe_div<A>(t.yB_P[fy] * t.xL_P[fx], sum); q[3] = safe_div<A>(t.yB_P[fy] * t.xR_P[fx], sum);
        }
    const long long ipl = (long long)Hi * Wi, opl = (long long)H * W;
    const int c0 = blockIdx.y * c_per_slice, c1 = min(C, c0 + c_per_slice);
    A* gi = gin1 + ((long long)b * C + c0) * ipl;
    const A* go = gout + ((long long)b * C + c0) * opl + (long long)y * W + x;

    bool use_box = false;
    int x_lo = 0, y_lo = 0, bw = 0, bh = 0;
    const int big = 1 << 28;   // far outside any image: the taps clamp to the border either way, and +- (NT*dil) cannot overflow
    const int fx_ = max(-big, min(big, t.flx)), fy_ = max(-big, min(big, t.fly));
    if (kBox) {
        const unsigned full = 0xffffffffu;
        // extent of floor(x + dx), floor(y + dy) over the tile
        const int mnx = __reduce_min_sync(full, active ? fx_ : big), mxx = __reduce_max_sync(full, active ? fx_ : -big);
        const int mny = __reduce_min_sync(full, active ? fy_ : big), mxy = __reduce_max_sync(full, active ? fy_ : -big);
        if (lane == 0) { ext[0][warp] = mnx; ext[1][warp] = mxx; ext[2][warp] = mny; ext[3][warp] = mxy; }
        for (int i = threadIdx.x; i < RS_CH * RS_BOX_H * RS_BOX_W; i += 32 * RS_TH1) box[i] = 0.f;
        __syncthreads();
        int e0 = big, e1 = -big, e2 = big, e3 = -big;
#pragma unroll
        for (int wv = 0; wv < RS_TH1; ++wv) {
            e0 = min(e0, ext[0][wv]); e1 = max(e1, ext[1][wv]); e2 = min(e2, ext[2][wv]); e3 = max(e3, ext[3][wv]);
        }
        if (e0 <= e1) {   // at least one active pixel
            x_lo = clampi(e0 - (NT - 1) * dil, Wi - 1);
            y_lo = clampi(e2 - (NT - 1) * dil, Hi - 1);
            bw = clampi(e1 + NT * dil, Wi - 1) - x_lo + 1;
            bh = clampi(e3 + NT * dil, Hi - 1) - y_lo + 1;
            use_box = bw <= RS_BOX_W && bh <= RS_BOX_H;
        }
    }
    if (use_box) {
        int toff[NT * NT * 4];   // the clamped taps of rs_setup, as offsets into the box
#pragma unroll
        for (int fy = 0; fy < NT; ++fy) {
            const int yT = clampi(fy_ - fy * dil, Hi - 1) - y_lo, yB = clampi(fy_ + (fy + 1) * dil, Hi - 1) - y_lo;
#pragma unroll
            for (int fx = 0; fx < NT; ++fx) {
                const int xL = clampi(fx_ - fx * dil, Wi - 1) - x_lo, xR = clampi(fx_ + (fx + 1) * dil, Wi - 1) - x_lo;
                int* o = toff + (fy * NT + fx) * 4;
                o[0] = yT * RS_BOX_W + xL; o[1] = yT * RS_BOX_W + xR; o[2] = yB * RS_BOX_W + xL; o[3] = yB * RS_BOX_W + xR;
            }
        }
        A* gbox = gi + (long long)y_lo * Wi + x_lo;
        for (int c = c0; c < c1; c += RS_CH, go += RS_CH * opl, gbox += RS_CH * ipl) {
            const int nch = min(RS_CH, c1 - c);
            if (active) {
#pragma unroll
                for (int j = 0; j < RS_CH; ++j) {
                    if (j < nch) {
                        const double g = static_cast<double>(go[j * opl]);
                        float* bj = box + j * (RS_BOX_H * RS_BOX_W);
#pragma unroll
                        for (int q = 0; q < NT * NT * 4; ++q) atomicAdd(bj + toff[q], static_cast<float>(wn[q] * g));
                    }
                }
            }
            __syncthreads();
            for (int j = 0; j < nch; ++j)
                for (int r = warp; r < bh; r += RS_TH1) {
                    float* row = box + j * (RS_BOX_H * RS_BOX_W) + r * RS_BOX_W;
                    for (int col = lane; col < bw; col += 32) {
                        const float v = row[col];
                        if (v != 0.f) {
                            row[col] = 0.f;
                            atomicAdd(reinterpret_cast<float*>(gbox) + j * ipl + (long long)r * Wi + col, v);
                        }
                    }
                }
            __syncthreads();
        }
        return;
    }
    if (!active) return;
#pragma unroll 4
    for (int c = c0; c < c1; ++c, gi += ipl, go += opl) {
        const double g = static_cast<double>(*go);
#pragma unroll
        for (int q = 0; q < NT * NT * 4; ++q) atomicAdd(gi + t.off[q], static_cast<A>(wn[q] * g));
    }
}

template <typename A, int NT>
__global__ void __launch_bounds__(128)
k_resample2d_bwd_in2(const A* __restrict__ in1, const A* __restrict__ in2, const A* __restrict__ gout,
                     A* __restrict__ gin2, int B, int C, int Hi, int Wi, int H, int W, int dil, int accumulate) {
    const RsPixel px = rs_pixel<4>(H, W);
    if (!px.active) return;
    const int x = px.x, y = px.y, b = px.b;
    RsTaps<A, NT> t;
    rs_setup<A, NT>(t, in2, b, y, x, H, W, Hi, Wi, dil, false);
    const A sum = rs_weight_sum<A, NT>(t);
    // corner dot products over the channels: D[q] = sum_c g[c] * in1[c, tap q]
    A D[NT * NT * 4];
#pragma unroll
    for (int q = 0; q < NT * NT * 4; ++q) D[q] = static_cast<A>(0);
    const long long ipl = (long long)Hi * Wi, opl = (long long)H * W;
    const A* s = in1 + (long long)b * C * ipl;
    const A* go = gout + (long long)b * C * opl + (long long)y * W + x;
#pragma unroll 4
    for (int c = 0; c < C; ++c, s += ipl, go += opl) {
        const A g = *go;
#pragma unroll
        for (int q = 0; q < NT * NT * 4; ++q) D[q] += g * s[t.off[q]];
    }
    // combine (per pixel, in double): reference :271-296 (grad1, sumgrad), :304-326 (grad2), :328
    const double sg = static_cast<double>(t.sigma);
    const bool s0 = (t.sigma == static_cast<A>(0));
    const double den_xy = s0 ? 1e-8 : -(sg * sg);          // SAFE_DIV(., -sigma*sigma)
    const double den_s = s0 ? 1e-8 : sg * sg * sg;         // SAFE_DIV(., sigma^3)
    double g1[3] = {0, 0, 0}, sgrad[3] = {0, 0, 0}, wd = 0;
#pragma unroll
    for (int fy = 0; fy < NT; ++fy)
#pragma unroll
        for (int fx = 0; fx < NT; ++fx) {
            const A* d = D + (fy * NT + fx) * 4;
            const double xL = t.xL_[fx], xR = t.xR_[fx], yT = t.yT_[fy], yB = t.yB_[fy];
            const double wTL = (double)t.yT_P[fy] * t.xL_P[fx], wTR = (double)t.yT_P[fy] * t.xR_P[fx];
            const double wBL = (double)t.yB_P[fy] * t.xL_P[fx], wBR = (double)t.yB_P[fy] * t.xR_P[fx];
            g1[0] += (xL * wTL * d[0] - xR * wTR * d[1] + xL * wBL * d[2] - xR * wBR * d[3]) / den_xy;
            sgrad[0] += (xL * wTL - xR * wTR + xL * wBL - xR * wBR) / den_xy;
            g1[1] += (yT * wTL * d[0] + yT * wTR * d[1] - yB * wBL * d[2] - yB * wBR * d[3]) / den_xy;
            sgrad[1] += (yT * wTL + yT * wTR - yB * wBL - yB * wBR) / den_xy;
            const double rTL = yT * yT + xL * xL, rTR = yT * yT + xR * xR, rBL = yB * yB + xL * xL, rBR = yB * yB + xR * xR;
            g1[2] += (rTL * wTL * d[0] + rTR * wTR * d[1] + rBL * wBL * d[2] + rBR * wBR * d[3]) / den_s;
            sgrad[2] += (rTL * wTL + rTR * wTR + rBL * wBL + rBR * wBR) / den_s;
            wd += wTL * d[0] + wTR * d[1] + wBL * d[2] + wBR * d[3];
        }
    const double S = static_cast<double>(sum);
    const double inv1 = (sum == static_cast<A>(0)) ? 1e8 : 1.0 / S;
    const double S2 = static_cast<double>(sum * sum);
    const double inv2 = (sum * sum == static_cast<A>(0)) ? 1e8 : 1.0 / S2;
    A* gp = gin2 + (long long)b * 3 * opl + (long long)y * W + x;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const A v = static_cast<A>(g1[c] * inv1 - (sgrad[c] * wd) * inv2);
        gp[c * opl] = accumulate ? gp[c * opl] + v : v;
    }
}

template <typename A, int NT>
static int rs_launch_fwd(const void* in1, const void* in2, void* out, int B, int C, int Hi, int Wi, int H, int W, int dil,
                         cudaStream_t st_) {
    const long long total = (long long)B * H * W;
    const int threads = 128, slices0 = channel_splits(total, C, threads), cps = (C + slices0 - 1) / slices0;
    dim3 grid((unsigned)rs_tiles<4>(B, H, W), (unsigned)((C + cps - 1) / cps));
    k_resample2d_fwd<A, NT><<<grid, threads, 0, st_>>>((const A*)in1, (const A*)in2, (A*)out, B, C, Hi, Wi, H, W, dil, cps);
    return launch_status();
}

template <typename A, int NT>
static int rs_launch_bwd(const void* in1, const void* in2, const void* gout, void* gin1, void* gin2, int B, int C, int Hi,
                         int Wi, int H, int W, int dil, int accumulate, cudaStream_t st_) {
    const long long total = (long long)B * H * W;
    const int threads = 32 * RS_TH1, slices0 = channel_splits(total, C, threads);
    const int cps = ((C + slices0 - 1) / slices0 + RS_CH - 1) / RS_CH * RS_CH;      // whole shared-memory channel groups per slice
    dim3 grid((unsigned)rs_tiles<RS_TH1>(B, H, W), (unsigned)((C + cps - 1) / cps));
    k_resample2d_bwd_in1<A, NT><<<grid, threads, 0, st_>>>((const A*)in2, (const A*)gout, (A*)gin1, B, C, Hi, Wi, H, W, dil, cps);
    int e = launch_status();
    if (e) return e;
    k_resample2d_bwd_in2<A, NT><<<(unsigned)rs_tiles<4>(B, H, W), 128, 0, st_>>>(
        (const A*)in1, (const A*)in2, (const A*)gout, (A*)gin2, B, C, Hi, Wi, H, W, dil, accumulate);
    return launch_status();
}

#define GFLA_RS_DISPATCH(A_, fn, ...)                          \
    switch (ks / 2) {                                          \
        case 1: return fn<A_, 1>(__VA_ARGS__);                 \
        case 2: return fn<A_, 2>(__VA_ARGS__);                 \
        case 3: return fn<A_, 3>(__VA_ARGS__);                 \
        case 4: return fn<A_, 4>(__VA_ARGS__);                 \
        default: return GFLA_E_SHAPE;                          \
    }

int resample2d_fwd(const void* in1, const void* in2, void* out, int B, int C, int Hi, int Wi, int H, int W, int ks,
                   int dil, int dtype, cudaStream_t st_) {
    if (dtype == GFLA_F32) { GFLA_RS_DISPATCH(float, rs_launch_fwd, in1, in2, out, B, C, Hi, Wi, H, W, dil, st_) }
    if (dtype == GFLA_F64) { GFLA_RS_DISPATCH(double, rs_launch_fwd, in1, in2, out, B, C, Hi, Wi, H, W, dil, st_) }
    return GFLA_E_DTYPE;
}

int resample2d_bwd(const void* in1, const void* in2, const void* gout, void* gin1, void* gin2, int B, int C, int Hi,
                   int Wi, int H, int W, int ks, int dil, int dtype, int accumulate, cudaStream_t st_) {
    if (!accumulate) {
        const int e = zero_async(gin1, (size_t)B * C * Hi * Wi * elem_size(dtype), st_);
        if (e != GFLA_OK) return e;
    }
    if (dtype == GFLA_F32) { GFLA_RS_DISPATCH(float, rs_launch_bwd, in1, in2, gout, gin1, gin2, B, C, Hi, Wi, H, W, dil, accumulate, st_) }
    if (dtype == GFLA_F64) { GFLA_RS_DISPATCH(double, rs_launch_bwd, in1, in2, gout, gin1, gin2, B, C, Hi, Wi, H, W, dil, accumulate, st_) }
    return GFLA_E_DTYPE;
}

}  // namespace gfla
