/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the GFLA warping hot path.
 *
 * This file is a *restatement* (not a copy) of the arithmetic performed by the
 * reference CUDA extensions, written as plain sequential C loops so that it
 * can be checked against (a) the reference kernel bodies compiled for the
 * host (oracle/_ref, see oracle/Makefile) and (b) the committed golden
 * vectors in tests/golden/.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * path (the CUDA library behind include/gfla_warp.h) never links or calls it.
 *
 * It is included twice by gfla_oracle.c with
 *     REAL = float  / SFX(x) = x##_f32
 *     REAL = double / SFX(x) = x##_f64
 * mirroring the reference's AT_DISPATCH_FLOATING_TYPES (float and double only;
 * block_extractor_kernel.cu:196, resample2d_kernel.cu:354).
 *
 * Reference citations are relative to /root/reference/model/networks/.
 *
 * Iteration order: every function walks its output in the reference's thread
 * index order (index = ((b*C + c)*H + y)*W + x), so that with one host thread
 * the floating-point accumulation order of the atomics in the reference's
 * backward kernels is reproduced exactly.
 */

#ifndef REAL
#error "include from gfla_oracle.c"
#endif

/* ------------------------------------------------------------------------- */
/* shared tap arithmetic for block_extractor                                  */
/* block_extractor/block_extractor_kernel.cu:57-76 (fwd) and :127-146 (bwd)    */
/* ------------------------------------------------------------------------- */
typedef struct {
    int xL, xR, yT, yB;       /* clamped integer taps                         */
    REAL wxL, wxR, wyT, wyB;  /* weights from the UNCLAMPED fractional part   */
} SFX(be_tap);

static inline int SFX(clampi)(int v, int hi) { /* max(min(v, hi), 0) */
    if (v > hi) v = hi;
    if (v < 0) v = 0;
    return v;
}

static inline SFX(be_tap) SFX(be_make_tap)(const REAL *flow, int64_t fbase_x,
                                           int64_t fbase_y, int yf, int xf,
                                           int oy, int ox, int Hs, int Ws) {
    SFX(be_tap) t;
    /* :62-63  flow channel 1 is y, channel 0 is x; offset added first */
    REAL flow_y = flow[fbase_y] + oy;
    REAL flow_x = flow[fbase_x] + ox;
    /* :65-66  then the integer pixel coordinate */
    REAL dy = flow_y + (REAL)yf;
    REAL dx = flow_x + (REAL)xf;
    REAL fx = FLOOR(dx), fy = FLOOR(dy);
    /* :69-72  clamp AFTER the int conversion (replicate border) */
    t.xL = SFX(clampi)((int)(fx), Ws - 1);
    t.xR = SFX(clampi)((int)(fx + 1), Ws - 1);
    t.yT = SFX(clampi)((int)(fy), Hs - 1);
    t.yB = SFX(clampi)((int)(fy + 1), Hs - 1);
    /* :73-76 */
    t.wxL = 1 - (dx - fx);
    t.wxR = dx - fx;
    t.wyT = 1 - (dy - fy);
    t.wyB = dy - fy;
    return t;
}

/* block_extractor forward.
 * source [B,C,Hs,Ws], flow [B,2,Hf,Wf] -> out [B,C,k*Hf,k*Wf]
 * block_extractor_kernel.cu:20-85, shapes from block_extractor.py:13-21 */
void SFX(oracle_block_extract_fwd)(const REAL *source, const REAL *flow,
                                   REAL *out, int B, int C, int Hs, int Ws,
                                   int Hf, int Wf, int k) {
    const int Ho = k * Hf, Wo = k * Wf;
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c) {
            const REAL *src = source + ((int64_t)b * C + c) * Hs * Ws;
            for (int y = 0; y < Ho; ++y)
                for (int x = 0; x < Wo; ++x) {
                    int yf = y / k, xf = x / k;
                    int oy = y % k - k / 2, ox = x % k - k / 2;
                    int64_t fx_i = (((int64_t)b * 2 + 0) * Hf + yf) * Wf + xf;
                    int64_t fy_i = (((int64_t)b * 2 + 1) * Hf + yf) * Wf + xf;
                    SFX(be_tap) t = SFX(be_make_tap)(flow, fx_i, fy_i, yf, xf,
                                                     oy, ox, Hs, Ws);
                    REAL s = 0;
                    s += (t.wxL * t.wyT * src[(int64_t)t.yT * Ws + t.xL]);
                    s += (t.wxR * t.wyT * src[(int64_t)t.yT * Ws + t.xR]);
                    s += (t.wxL * t.wyB * src[(int64_t)t.yB * Ws + t.xL]);
                    s += (t.wxR * t.wyB * src[(int64_t)t.yB * Ws + t.xR]);
                    out[(((int64_t)b * C + c) * Ho + y) * Wo + x] = s;
                }
        }
}

/* block_extractor backward: ACCUMULATES into grad_source / grad_flow (the
 * reference kernels atomicAdd into caller-zeroed buffers,
 * block_extractor_kernel.cu:158-168, block_extractor.py:35-36). */
void SFX(oracle_block_extract_bwd)(const REAL *source, const REAL *flow,
                                   const REAL *grad_out, REAL *grad_source,
                                   REAL *grad_flow, int B, int C, int Hs,
                                   int Ws, int Hf, int Wf, int k) {
    const int Ho = k * Hf, Wo = k * Wf;
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c) {
            const REAL *src = source + ((int64_t)b * C + c) * Hs * Ws;
            REAL *gs = grad_source + ((int64_t)b * C + c) * Hs * Ws;
            for (int y = 0; y < Ho; ++y)
                for (int x = 0; x < Wo; ++x) {
                    int yf = y / k, xf = x / k;
                    int oy = y % k - k / 2, ox = x % k - k / 2;
                    int64_t fx_i = (((int64_t)b * 2 + 0) * Hf + yf) * Wf + xf;
                    int64_t fy_i = (((int64_t)b * 2 + 1) * Hf + yf) * Wf + xf;
                    SFX(be_tap) t = SFX(be_make_tap)(flow, fx_i, fy_i, yf, xf,
                                                     oy, ox, Hs, Ws);
                    REAL vLT = src[(int64_t)t.yT * Ws + t.xL];
                    REAL vRT = src[(int64_t)t.yT * Ws + t.xR];
                    REAL vLB = src[(int64_t)t.yB * Ws + t.xL];
                    REAL vRB = src[(int64_t)t.yB * Ws + t.xR];
                    REAL g = grad_out[(((int64_t)b * C + c) * Ho + y) * Wo + x];
                    /* :158-161 */
                    gs[(int64_t)t.yT * Ws + t.xL] += g * t.wxL * t.wyT;
                    gs[(int64_t)t.yT * Ws + t.xR] += g * t.wxR * t.wyT;
                    gs[(int64_t)t.yB * Ws + t.xL] += g * t.wxL * t.wyB;
                    gs[(int64_t)t.yB * Ws + t.xR] += g * t.wxR * t.wyB;
                    /* :163-164 */
                    REAL gy = g * (-t.wxL * vLT - t.wxR * vRT + t.wxL * vLB + t.wxR * vRB);
                    REAL gx = g * (-t.wyT * vLT - t.wyB * vLB + t.wyT * vRT + t.wyB * vRB);
                    /* :167-168 */
                    grad_flow[fy_i] += gy;
                    grad_flow[fx_i] += gx;
                }
        }
}

/* ------------------------------------------------------------------------- */
/* local_attn_reshape: [B,k*k,H,W] -> [B,1,k*H,k*W]                            */
/* local_attn_reshape/local_attn_reshape_kernel.cu:20-61, :65-108              */
/* ------------------------------------------------------------------------- */
void SFX(oracle_attn_reshape_fwd)(const REAL *in, REAL *out, int B, int H,
                                  int W, int k) {
    const int Ho = k * H, Wo = k * W;
    for (int b = 0; b < B; ++b)
        for (int y = 0; y < Ho; ++y)
            for (int x = 0; x < Wo; ++x) {
                int cs = (y % k) * k + (x % k); /* :52-56 */
                out[((int64_t)b * Ho + y) * Wo + x] =
                    in[(((int64_t)b * k * k + cs) * H + y / k) * W + x / k];
            }
}

/* accumulates (reference: atomicAdd, :106) */
void SFX(oracle_attn_reshape_bwd)(const REAL *grad_out, REAL *grad_in, int B,
                                  int H, int W, int k) {
    const int Ho = k * H, Wo = k * W;
    for (int b = 0; b < B; ++b)
        for (int y = 0; y < Ho; ++y)
            for (int x = 0; x < Wo; ++x) {
                int cs = (y % k) * k + (x % k);
                grad_in[(((int64_t)b * k * k + cs) * H + y / k) * W + x / k] +=
                    grad_out[((int64_t)b * Ho + y) * Wo + x];
            }
}

/* ------------------------------------------------------------------------- */
/* resample2d (Gaussian-weighted ks x ks warp)                                 */
/* resample2d_package/resample2d_kernel.cu:20-95 / 98-202 / 204-330            */
/* ------------------------------------------------------------------------- */

/* :14-15.  EPS is a *double* literal, so for REAL=float the conditional
 * expression has type double: the quotient a/b is formed in REAL and then
 * widened, a/EPS is formed in double.  Everything that consumes the result
 * (exp, the += accumulations) therefore runs in double before being narrowed
 * back to REAL on assignment.  The macro below keeps exactly that typing. */
#define ORACLE_EPS 1e-8
#define ORACLE_SAFE_DIV(a, b) (((b) == 0) ? ((a) / (ORACLE_EPS)) : ((a) / (b)))

typedef struct {
    REAL xf, yf, alpha, beta, sigma;
} SFX(rs_pix);

/* in2 = [B,3,H,W]: channel 0 dx, 1 dy, 2 sigma (:47-49) */
static inline SFX(rs_pix) SFX(rs_make_pix)(const REAL *in2, int b, int y, int x,
                                           int H, int W, int trunc_frac) {
    SFX(rs_pix) p;
    int64_t hw = (int64_t)H * W;
    int64_t o = (int64_t)b * 3 * hw + (int64_t)y * W + x;
    REAL dx = in2[o], dy = in2[o + hw];
    p.sigma = in2[o + 2 * hw];
    p.xf = (REAL)x + dx; /* :52-53 */
    p.yf = (REAL)y + dy;
    if (trunc_frac) { /* backward_input1 quirk, :137-138: int() not floor() */
        p.alpha = p.xf - (int)(p.xf);
        p.beta = p.yf - (int)(p.yf);
    } else { /* :54-55 */
        p.alpha = p.xf - FLOOR(p.xf);
        p.beta = p.yf - FLOOR(p.yf);
    }
    return p;
}

typedef struct {
    REAL xL_, xR_, yT_, yB_;     /* distances  (:70-73) */
    REAL xL_P, xR_P, yT_P, yB_P; /* Gaussian weights (:75-78) */
} SFX(rs_w);

static inline SFX(rs_w) SFX(rs_make_w)(const SFX(rs_pix) * p, int fx, int fy,
                                       int dil) {
    SFX(rs_w) w;
    REAL sigma = p->sigma;
    w.xL_ = ((REAL)(fx * dil) + p->alpha);
    w.xR_ = ((REAL)((1. + fx) * dil) - p->alpha);
    w.yT_ = ((REAL)(fy * dil) + p->beta);
    w.yB_ = ((REAL)((1. + fy) * dil) - p->beta);
    w.xL_P = exp(ORACLE_SAFE_DIV(-w.xL_ * w.xL_, 2 * sigma * sigma));
    w.xR_P = exp(ORACLE_SAFE_DIV(-w.xR_ * w.xR_, 2 * sigma * sigma));
    w.yT_P = exp(ORACLE_SAFE_DIV(-w.yT_ * w.yT_, 2 * sigma * sigma));
    w.yB_P = exp(ORACLE_SAFE_DIV(-w.yB_ * w.yB_, 2 * sigma * sigma));
    return w;
}

/* in1 [B,C,Hi,Wi], in2 [B,3,H,W] -> out [B,C,H,W]   (:20-95) */
void SFX(oracle_resample2d_fwd)(const REAL *in1, const REAL *in2, REAL *out,
                                int B, int C, int Hi, int Wi, int H, int W,
                                int ks, int dil) {
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c) {
            const REAL *src = in1 + ((int64_t)b * C + c) * Hi * Wi;
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    SFX(rs_pix) p = SFX(rs_make_pix)(in2, b, y, x, H, W, 0);
                    REAL val = 0, sum = 0;
                    for (int fy = 0; fy < ks / 2; ++fy) {
                        int yT = SFX(clampi)((int)(FLOOR(p.yf) - fy * dil), Hi - 1);
                        int yB = SFX(clampi)((int)(FLOOR(p.yf) + (fy + 1) * dil), Hi - 1);
                        for (int fx = 0; fx < ks / 2; ++fx) {
                            int xL = SFX(clampi)((int)(FLOOR(p.xf) - fx * dil), Wi - 1);
                            int xR = SFX(clampi)((int)(FLOOR(p.xf) + (fx + 1) * dil), Wi - 1);
                            SFX(rs_w) w = SFX(rs_make_w)(&p, fx, fy, dil);
                            val += (REAL)(w.yT_P * w.xL_P * src[(int64_t)yT * Wi + xL]);
                            val += (REAL)(w.yT_P * w.xR_P * src[(int64_t)yT * Wi + xR]);
                            val += (REAL)(w.yB_P * w.xL_P * src[(int64_t)yB * Wi + xL]);
                            val += (REAL)(w.yB_P * w.xR_P * src[(int64_t)yB * Wi + xR]);
                            sum += (w.yT_P * w.xL_P + w.yT_P * w.xR_P +
                                    w.yB_P * w.xL_P + w.yB_P * w.xR_P);
                        }
                    }
                    out[(((int64_t)b * C + c) * H + y) * W + x] = ORACLE_SAFE_DIV(val, sum);
                }
        }
}

/* gradient w.r.t. in1: ACCUMULATES (atomicAdd, :195-198).  (:98-202) */
void SFX(oracle_resample2d_bwd_input1)(const REAL *in1, const REAL *in2,
                                       const REAL *grad_out, REAL *grad_in1,
                                       int B, int C, int Hi, int Wi, int H,
                                       int W, int ks, int dil) {
    (void)in1;
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c) {
            REAL *gi = grad_in1 + ((int64_t)b * C + c) * Hi * Wi;
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    /* weights use the truncating fraction (:137-138) ...   */
                    SFX(rs_pix) p = SFX(rs_make_pix)(in2, b, y, x, H, W, 1);
                    REAL sum = 0;
                    for (int fy = 0; fy < ks / 2; ++fy)
                        for (int fx = 0; fx < ks / 2; ++fx) {
                            SFX(rs_w) w = SFX(rs_make_w)(&p, fx, fy, dil);
                            sum += (w.yT_P * w.xL_P + w.yT_P * w.xR_P +
                                    w.yB_P * w.xL_P + w.yB_P * w.xR_P);
                        }
                    REAL g = grad_out[(((int64_t)b * C + c) * H + y) * W + x];
                    /* atomicAdd(float*, <double expr>) narrows the addend to
                     * REAL before adding -- hence the casts below. */
                    for (int fy = 0; fy < ks / 2; ++fy) {
                        /* ... while the tap indices use floor (:170-171) */
                        int yT = SFX(clampi)((int)(FLOOR(p.yf) - fy * dil), Hi - 1);
                        int yB = SFX(clampi)((int)(FLOOR(p.yf) + (fy + 1) * dil), Hi - 1);
                        for (int fx = 0; fx < ks / 2; ++fx) {
                            int xL = SFX(clampi)((int)(FLOOR(p.xf) - fx * dil), Wi - 1);
                            int xR = SFX(clampi)((int)(FLOOR(p.xf) + (fx + 1) * dil), Wi - 1);
                            SFX(rs_w) w = SFX(rs_make_w)(&p, fx, fy, dil);
                            gi[(int64_t)yT * Wi + xL] += (REAL)(ORACLE_SAFE_DIV(w.yT_P * w.xL_P, sum) * g);
                            gi[(int64_t)yT * Wi + xR] += (REAL)(ORACLE_SAFE_DIV(w.yT_P * w.xR_P, sum) * g);
                            gi[(int64_t)yB * Wi + xL] += (REAL)(ORACLE_SAFE_DIV(w.yB_P * w.xL_P, sum) * g);
                            gi[(int64_t)yB * Wi + xR] += (REAL)(ORACLE_SAFE_DIV(w.yB_P * w.xR_P, sum) * g);
                        }
                    }
                }
        }
}

/* gradient w.r.t. in2 = (dx, dy, sigma): OVERWRITES grad_in2 [B,3,H,W]
 * (plain store at :328).  (:204-330)
 * Note the reference accumulates `sumgrad` once PER CHANNEL inside the ch loop
 * (:281,:288,:295) and later divides it by the channel count (:318): the
 * restatement keeps that (it is not the same rounding as computing it once). */
void SFX(oracle_resample2d_bwd_input2)(const REAL *in1, const REAL *in2,
                                       const REAL *grad_out, REAL *grad_in2,
                                       int B, int C, int Hi, int Wi, int H,
                                       int W, int ks, int dil) {
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < 3; ++c)
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    SFX(rs_pix) p = SFX(rs_make_pix)(in2, b, y, x, H, W, 0);
                    REAL sigma = p.sigma;
                    REAL grad1 = 0.0, grad2 = 0.0, sum = 0.0, sumgrad = 0.0;
                    for (int fy = 0; fy < ks / 2; ++fy) {
                        int yT = SFX(clampi)((int)(FLOOR(p.yf) - fy * dil), Hi - 1);
                        int yB = SFX(clampi)((int)(FLOOR(p.yf) + (fy + 1) * dil), Hi - 1);
                        for (int fx = 0; fx < ks / 2; ++fx) {
                            int xL = SFX(clampi)((int)(FLOOR(p.xf) - fx * dil), Wi - 1);
                            int xR = SFX(clampi)((int)(FLOOR(p.xf) + (fx + 1) * dil), Wi - 1);
                            SFX(rs_w) w = SFX(rs_make_w)(&p, fx, fy, dil);
                            sum += (w.yT_P * w.xL_P + w.yT_P * w.xR_P +
                                    w.yB_P * w.xL_P + w.yB_P * w.xR_P);
                            for (int ch = 0; ch < C; ++ch) {
                                const REAL *src = in1 + ((int64_t)b * C + ch) * Hi * Wi;
                                REAL g = grad_out[(((int64_t)b * C + ch) * H + y) * W + x];
                                REAL vLT = src[(int64_t)yT * Wi + xL];
                                REAL vRT = src[(int64_t)yT * Wi + xR];
                                REAL vLB = src[(int64_t)yB * Wi + xL];
                                REAL vRB = src[(int64_t)yB * Wi + xR];
                                if (c == 0) { /* d/dx :276-281 */
                                    grad1 += ORACLE_SAFE_DIV(w.xL_ * w.yT_P * w.xL_P * g * vLT, -sigma * sigma);
                                    grad1 -= ORACLE_SAFE_DIV(w.xR_ * w.yT_P * w.xR_P * g * vRT, -sigma * sigma);
                                    grad1 += ORACLE_SAFE_DIV(w.xL_ * w.yB_P * w.xL_P * g * vLB, -sigma * sigma);
                                    grad1 -= ORACLE_SAFE_DIV(w.xR_ * w.yB_P * w.xR_P * g * vRB, -sigma * sigma);
                                    sumgrad += ORACLE_SAFE_DIV((w.xL_ * w.yT_P * w.xL_P - w.xR_ * w.yT_P * w.xR_P +
                                                                w.xL_ * w.yB_P * w.xL_P - w.xR_ * w.yB_P * w.xR_P),
                                                               -sigma * sigma);
                                } else if (c == 1) { /* d/dy :283-288 */
                                    grad1 += ORACLE_SAFE_DIV(w.yT_ * w.yT_P * w.xL_P * g * vLT, -sigma * sigma);
                                    grad1 += ORACLE_SAFE_DIV(w.yT_ * w.yT_P * w.xR_P * g * vRT, -sigma * sigma);
                                    grad1 -= ORACLE_SAFE_DIV(w.yB_ * w.yB_P * w.xL_P * g * vLB, -sigma * sigma);
                                    grad1 -= ORACLE_SAFE_DIV(w.yB_ * w.yB_P * w.xR_P * g * vRB, -sigma * sigma);
                                    sumgrad += ORACLE_SAFE_DIV((w.yT_ * w.yT_P * w.xL_P + w.yT_ * w.yT_P * w.xR_P -
                                                                w.yB_ * w.yB_P * w.xL_P - w.yB_ * w.yB_P * w.xR_P),
                                                               -sigma * sigma);
                                } else { /* d/dsigma :290-296 */
                                    grad1 += ORACLE_SAFE_DIV((w.yT_ * w.yT_ + w.xL_ * w.xL_) * w.yT_P * w.xL_P * g * vLT, sigma * sigma * sigma);
                                    grad1 += ORACLE_SAFE_DIV((w.yT_ * w.yT_ + w.xR_ * w.xR_) * w.yT_P * w.xR_P * g * vRT, sigma * sigma * sigma);
                                    grad1 += ORACLE_SAFE_DIV((w.yB_ * w.yB_ + w.xL_ * w.xL_) * w.yB_P * w.xL_P * g * vLB, sigma * sigma * sigma);
                                    grad1 += ORACLE_SAFE_DIV((w.yB_ * w.yB_ + w.xR_ * w.xR_) * w.yB_P * w.xR_P * g * vRB, sigma * sigma * sigma);
                                    sumgrad += ORACLE_SAFE_DIV(((w.yT_ * w.yT_ + w.xL_ * w.xL_) * w.yT_P * w.xL_P +
                                                                (w.yT_ * w.yT_ + w.xR_ * w.xR_) * w.yT_P * w.xR_P +
                                                                (w.yB_ * w.yB_ + w.xL_ * w.xL_) * w.yB_P * w.xL_P +
                                                                (w.yB_ * w.yB_ + w.xR_ * w.xR_) * w.yB_P * w.xR_P),
                                                               sigma * sigma * sigma);
                                }
                            }
                        }
                    }
                    /* second sweep :304-326 */
                    for (int fy = 0; fy < ks / 2; ++fy) {
                        int yT = SFX(clampi)((int)(FLOOR(p.yf) - fy * dil), Hi - 1);
                        int yB = SFX(clampi)((int)(FLOOR(p.yf) + (fy + 1) * dil), Hi - 1);
                        for (int fx = 0; fx < ks / 2; ++fx) {
                            int xL = SFX(clampi)((int)(FLOOR(p.xf) - fx * dil), Wi - 1);
                            int xR = SFX(clampi)((int)(FLOOR(p.xf) + (fx + 1) * dil), Wi - 1);
                            SFX(rs_w) w = SFX(rs_make_w)(&p, fx, fy, dil);
                            for (int ch = 0; ch < C; ++ch) {
                                const REAL *src = in1 + ((int64_t)b * C + ch) * Hi * Wi;
                                REAL g = grad_out[(((int64_t)b * C + ch) * H + y) * W + x];
                                grad2 += sumgrad / C * w.yT_P * w.xL_P * g * src[(int64_t)yT * Wi + xL];
                                grad2 += sumgrad / C * w.yT_P * w.xR_P * g * src[(int64_t)yT * Wi + xR];
                                grad2 += sumgrad / C * w.yB_P * w.xL_P * g * src[(int64_t)yB * Wi + xL];
                                grad2 += sumgrad / C * w.yB_P * w.xR_P * g * src[(int64_t)yB * Wi + xR];
                            }
                        }
                    }
                    grad_in2[(((int64_t)b * 3 + c) * H + y) * W + x] =
                        ORACLE_SAFE_DIV(grad1, sum) - ORACLE_SAFE_DIV(grad2, sum * sum);
                }
}

/* ------------------------------------------------------------------------- */
/* fused local attention = the tail of ExtractorAttn.forward                   */
/* base_function.py:804-810, with softmax=True (generator.py:112):             */
/*   block  = BlockExtractor(k)(source, flow)            [B,C,kH,kW]           */
/*   p      = Softmax(dim=1)(logits)                     [B,k*k,H,W]           */
/*   attn   = LocalAttnReshape()(p, k)                   [B,1,kH,kW]           */
/*   out    = avg_pool2d(attn * block, k, k)             [B,C,H,W]             */
/* The oracle composes its own pieces literally (materialising `block`), so it */
/* is only meant for small shapes.                                             */
/* ------------------------------------------------------------------------- */
static void SFX(softmax_dim1)(const REAL *logits, REAL *p, int B, int K, int H, int W) {
    int64_t hw = (int64_t)H * W;
    for (int b = 0; b < B; ++b)
        for (int64_t i = 0; i < hw; ++i) {
            const REAL *l = logits + (int64_t)b * K * hw + i;
            REAL *q = p + (int64_t)b * K * hw + i;
            REAL m = l[0];
            for (int t = 1; t < K; ++t)
                if (l[t * hw] > m) m = l[t * hw];
            REAL s = 0;
            for (int t = 0; t < K; ++t) {
                q[t * hw] = EXP(l[t * hw] - m);
                s += q[t * hw];
            }
            for (int t = 0; t < K; ++t) q[t * hw] = q[t * hw] / s;
        }
}

/* source [B,C,Hs,Ws], flow [B,2,H,W], logits [B,k*k,H,W] -> out [B,C,H,W].
 * probs_out (optional, [B,k*k,H,W]) receives the softmax, which is what
 * ExtractorAttn.hook_attn_param returns (base_function.py:812-818). */
void SFX(oracle_local_attn_fwd)(const REAL *source, const REAL *flow,
                                const REAL *logits, REAL *out, REAL *probs_out,
                                int B, int C, int Hs, int Ws, int H, int W, int k) {
    const int K = k * k, Ho = k * H, Wo = k * W;
    REAL *block = (REAL *)malloc(sizeof(REAL) * (size_t)B * C * Ho * Wo);
    REAL *p = (REAL *)malloc(sizeof(REAL) * (size_t)B * K * H * W);
    REAL *attn = (REAL *)malloc(sizeof(REAL) * (size_t)B * Ho * Wo);
    SFX(oracle_block_extract_fwd)(source, flow, block, B, C, Hs, Ws, H, W, k);
    SFX(softmax_dim1)(logits, p, B, K, H, W);
    SFX(oracle_attn_reshape_fwd)(p, attn, B, H, W, k);
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    REAL acc = 0;
                    for (int i = 0; i < k; ++i)
                        for (int j = 0; j < k; ++j) {
                            int64_t o = ((int64_t)(y * k + i)) * Wo + (x * k + j);
                            acc += attn[(int64_t)b * Ho * Wo + o] *
                                   block[((int64_t)b * C + c) * Ho * Wo + o];
                        }
                    out[(((int64_t)b * C + c) * H + y) * W + x] = acc / (REAL)K;
                }
    if (probs_out) memcpy(probs_out, p, sizeof(REAL) * (size_t)B * K * H * W);
    free(block); free(p); free(attn);
}

/* backward of the composition above.  grad_source / grad_flow ACCUMULATE (they
 * come from block_extract_bwd), grad_logits is overwritten. */
void SFX(oracle_local_attn_bwd)(const REAL *source, const REAL *flow,
                                const REAL *logits, const REAL *grad_out,
                                REAL *grad_source, REAL *grad_flow,
                                REAL *grad_logits, int B, int C, int Hs, int Ws,
                                int H, int W, int k) {
    const int K = k * k, Ho = k * H, Wo = k * W;
    int64_t hw = (int64_t)H * W;
    REAL *block = (REAL *)malloc(sizeof(REAL) * (size_t)B * C * Ho * Wo);
    REAL *gblock = (REAL *)malloc(sizeof(REAL) * (size_t)B * C * Ho * Wo);
    REAL *p = (REAL *)malloc(sizeof(REAL) * (size_t)B * K * H * W);
    REAL *attn = (REAL *)malloc(sizeof(REAL) * (size_t)B * Ho * Wo);
    REAL *gattn = (REAL *)calloc((size_t)B * Ho * Wo, sizeof(REAL));
    REAL *gp = (REAL *)calloc((size_t)B * K * H * W, sizeof(REAL));
    SFX(oracle_block_extract_fwd)(source, flow, block, B, C, Hs, Ws, H, W, k);
    SFX(softmax_dim1)(logits, p, B, K, H, W);
    SFX(oracle_attn_reshape_fwd)(p, attn, B, H, W, k);
    /* avg_pool2d backward spreads g/K over the k x k cell; then the product rule */
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int y = 0; y < Ho; ++y)
                for (int x = 0; x < Wo; ++x) {
                    REAL g = grad_out[(((int64_t)b * C + c) * H + y / k) * W + x / k] / (REAL)K;
                    int64_t o = (int64_t)y * Wo + x;
                    int64_t bo = ((int64_t)b * C + c) * Ho * Wo + o;
                    gblock[bo] = g * attn[(int64_t)b * Ho * Wo + o];
                    gattn[(int64_t)b * Ho * Wo + o] += g * block[bo];
                }
    SFX(oracle_attn_reshape_bwd)(gattn, gp, B, H, W, k);
    /* softmax backward: dl_t = p_t * (dp_t - sum_u p_u dp_u) */
    for (int b = 0; b < B; ++b)
        for (int64_t i = 0; i < hw; ++i) {
            const REAL *q = p + (int64_t)b * K * hw + i;
            const REAL *dq = gp + (int64_t)b * K * hw + i;
            REAL dot = 0;
            for (int t = 0; t < K; ++t) dot += q[t * hw] * dq[t * hw];
            for (int t = 0; t < K; ++t)
                grad_logits[(int64_t)b * K * hw + t * hw + i] = q[t * hw] * (dq[t * hw] - dot);
        }
    SFX(oracle_block_extract_bwd)(source, flow, gblock, grad_source, grad_flow,
                                  B, C, Hs, Ws, H, W, k);
    free(block); free(gblock); free(p); free(attn); free(gattn); free(gp);
}

#undef ORACLE_EPS
#undef ORACLE_SAFE_DIV
