// Fused local-attention FORWARD on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
//
//   out[b,c,p] = sum_t  Wfull[p,t] * source[b,c,t]
//
// Per output pixel p the reference's 4*k*k bilinear taps times softmax
// probabilities (base_function.py:804-810, block_extractor_kernel.cu:57-82)
// collapse to a (k+1)x(k+1) window of weights W_p.  Doing that weighted sum on
// the CUDA cores costs >= (k+1)^2 = 36 multiply-adds AND shared-memory reads
// per (pixel, channel): at cfg2 that is ~4x over the LSU/FMA budget of a
// 70 %-of-HBM-roofline kernel.  So the sum is embedded in a dense GEMM:
//
//   * a group of 16x8 = 128 output pixels is the M dimension;
//   * the channels (CN <= 256) are the N dimension;
//   * K runs over the source positions of the group's tap footprint, one
//     16-wide row segment at a time (K = 16 = one tcgen05.mma per segment);
//   * B = the source rows themselves, TMA-loaded straight from NCHW as boxes
//     [CN channels][16 x] with the 32-byte swizzle: that IS a K-major UMMA
//     operand, no transposition;
//   * A = the sparse weight matrix [128 pixels][16 positions] per segment, built
//     by 128 "builder" threads (one per pixel) from flow + logits;
//   * D accumulates in TMEM ([128 lanes = pixels] x [CN fp32 columns]),
//     double-buffered so the epilogue of group g overlaps the MMAs of g+1.
//
// The weight matrix is ~90 % zeros, i.e. the tensor pipe does ~12x redundant
// work -- and is still several times faster than the CUDA cores would be,
// which is what makes the kernel memory-bound again.
//
// Warp roles (10 warps, persistent CTA, static round-robin over groups):
//   warp 0      producer: per group computes the tap bounding box from the flow,
//               publishes it, issues the TMA row loads;
//   warp 1      MMA issuer (one elected lane);
//   warps 2-5   builders: softmax, tap/clamp arithmetic (bit-identical to the
//               reference), window collapse, per-stage weight slabs;
//   warps 6-9   epilogue: TMEM -> registers -> bf16 -> coalesced NCHW stores;
//               pixels whose taps are not consecutive integers (fp32 rounding
//               straddling an integer -- measure-zero) are recomputed here with
//               the literal 4-tap path so indexing stays bit-identical.
#include "tile_window.cuh"

namespace gfla {

namespace tc {

constexpr int RCH = 2;                  // source rows per pipeline stage
constexpr int NSTAGE = 4;                // stages of FBW x RCH positions
constexpr int NINFO = 8;                 // >= NSTAGE + 3 (producer run-ahead + 2 accumulators in flight)
constexpr int NTHREADS = 320;

// Row-segment width: channels-last boxes start at any x, so one 32-wide segment usually spans the whole tap
// footprint of a 16-pixel-wide group (half as many stages and slab rows as two 16-wide ones).  Planar (NCHW)
// boxes must start on an 8-pixel boundary, where 16-wide segments waste less.
template <bool NHWC> struct SegW { static constexpr int value = NHWC ? 32 : 16; };

template <int CN, int FBW>
struct Smem {
    static constexpr int S_SLAB = CN * FBW * 2;              // one source row segment: [CN channels] x [FBW x] bf16
    static constexpr int FA_SLAB = 128 * FBW * 2;            // weight slab: [128 pixels][FBW positions] bf16
    static constexpr int S_STAGE = RCH * S_SLAB;
    static constexpr int A_STAGE = RCH * FA_SLAB;
    static constexpr int OFF_S = 0;
    static constexpr int OFF_A = OFF_S + NSTAGE * S_STAGE;
    static constexpr int OFF_W = OFF_A + NSTAGE * A_STAGE;   // [36][128] bf16 collapsed windows
    static constexpr int OFF_INFO = OFF_W + 36 * 128 * 2;
    static constexpr int OFF_BAR = OFF_INFO + NINFO * 16;
    static constexpr int NBAR = 3 * NSTAGE + 4 + NINFO;
    static constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
    static constexpr int TOTAL = OFF_TMEM + 16;
    static constexpr int ALLOC = TOTAL + 1024;               // slack to align the base to 1024 B
};

// NHWC = channels-last storage of source / out (logical shapes stay [B,C,H,W]): every source position is
// 2*C contiguous bytes, so the TMA boxes [64 channels][16 x] are made of 128-byte runs (the NCHW variant has
// to fetch 32-byte runs, one per channel, and needs its box origin aligned to 8 pixels), the box origin is
// unconstrained in x, the smem tile is an MN-major (channel-contiguous) UMMA B operand, and the epilogue
// stores 64 contiguous bytes per thread.
// `knobs` (environment GFLA_TC_KNOBS, default 0 = production): bits 0-7 = 1/2 L2 prefetch of the next group's source box
// (tensor / bulk; measured: no gain), bit 8 = skip output stores, bit 9 = skip window construction, bit 10 = skip the
// slab scatter -- timing experiments only, results are wrong when bits 8-10 are set.
template <int K, int CN, bool NHWC>
__global__ void __launch_bounds__(NTHREADS, 1)
k_local_attn_fwd_tc(const __grid_constant__ CUtensorMap tmap_src, const __nv_bfloat16* __restrict__ src,
                    const float* __restrict__ flow, const __nv_bfloat16* __restrict__ logits,
                    __nv_bfloat16* __restrict__ out, __nv_bfloat16* __restrict__ probs,
                    const __nv_bfloat16* __restrict__ prev, const __nv_bfloat16* __restrict__ mask, int B, int C, int Hs,
                    int Ws, int H, int W, int knobs_arg) {
    const int knobs = GFLA_KNOBS(knobs_arg);      // 0 in the shipped build: every `knobs & x` test folds away
    constexpr int FBW = SegW<NHWC>::value;
    using SM = Smem<CN, FBW>;
    constexpr int K1 = K + 1, KK = K * K;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::OFF_BAR);
    uint64_t* full_s = bars;                      // [NSTAGE] TMA bytes landed
    uint64_t* full_a = bars + NSTAGE;             // [NSTAGE] 128 builder arrivals
    uint64_t* empty = bars + 2 * NSTAGE;          // [NSTAGE] MMAs of the stage retired
    uint64_t* acc_full = bars + 3 * NSTAGE;       // [2]
    uint64_t* acc_empty = bars + 3 * NSTAGE + 2;  // [2]
    uint64_t* info_full = bars + 3 * NSTAGE + 4;  // [NINFO]
    GroupInfo* infos = reinterpret_cast<GroupInfo*>(smem + SM::OFF_INFO);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM::OFF_TMEM);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t_start = tc_profile_clock();
    const int gxn = (W + GW - 1) / GW, gyn = (H + GH - 1) / GH;
    const int ngroups = B * gyn * gxn;
    const int c0 = blockIdx.y * CN;
    const long long hw = (long long)H * W;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(&full_s[i], 1); mbar_init(&full_a[i], 128); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
        for (int i = 0; i < NINFO; ++i) mbar_init(&info_full[i], 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmap_src);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * CN >= 32 ? 2 * CN : 32);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================================================= producer
        auto bbox_of = [&](int g, int& x0, int& y0, int& x1, int& y1) {
            group_bbox<K>(flow, g / (gxn * gyn), (g % gxn) * GW, ((g / gxn) % gyn) * GH, H, W, Hs, Ws, lane, !NHWC, x0, y0, x1, y1);
        };
        uint32_t it = 0;  // global stage counter
        int gi = 0;
        int nx0 = 0, ny0 = 0, nx1 = 0, ny1 = 0;
        if ((int)blockIdx.x < ngroups) bbox_of(blockIdx.x, nx0, ny0, nx1, ny1);
        for (int g = blockIdx.x; g < ngroups; g += gridDim.x, ++gi) {
            const int b = g / (gxn * gyn);
            const int xmin = nx0, ymin = ny0, xmax = nx1, ymax = ny1;
            // the next group of this CTA lies ~gridDim.x groups ahead in memory: nobody has touched its source
            // rows yet, so pull them into L2 now -- a whole group of time before the TMA loads need them
            const int gn = g + gridDim.x;
            if (gn < ngroups) {
                const long long tb0 = tc_profile_clock();
                bbox_of(gn, nx0, ny0, nx1, ny1);
                tc_profile_add(0, 6, tc_profile_clock() - tb0);      // next group's bounding box
                const int bn = gn / (gxn * gyn), prow = ny1 - ny0 + 1;
                if ((knobs & 255) == 2 && NHWC && CN == C) {
                    // channels-last, all channels in this CTA: a row segment of the box is one contiguous range
                    const uint32_t bytes = static_cast<uint32_t>(nx1 - nx0 + 1) * C * 2;
                    for (int r = lane; r < prow; r += 32)
                        prefetch_l2_bulk(src + (((long long)bn * Hs + ny0 + r) * Ws + nx0) * C, bytes);
                } else if ((knobs & 255) == 1) {
                    const int pcb = (nx1 - nx0 + FBW) / FBW;
                    for (int idx = lane; idx < pcb * prow; idx += 32) {
                        const int cb = idx / prow, yy = ny0 + idx % prow;
                        if (NHWC) {
#pragma unroll
                            for (int cg = 0; cg < CN / 64; ++cg) tma_prefetch_4d(&tmap_src, c0 + cg * 64, nx0 + cb * FBW, yy, bn);
                        } else {
                            tma_prefetch_4d(&tmap_src, nx0 + cb * FBW, yy, c0, bn);
                        }
                    }
                }
            }
            const int ncb = (xmax - xmin + FBW) / FBW, nrc = (ymax - ymin + RCH) / RCH;
            if (lane == 0) {
                infos[gi % NINFO] = GroupInfo{xmin, ymin, ncb, nrc};
                mbar_arrive(&info_full[gi % NINFO]);
            }
            for (int cb = 0; cb < ncb; ++cb)
                for (int rc = 0; rc < nrc; ++rc, ++it) {
                    const int slot = it % NSTAGE;
                    mbar_wait(&empty[slot], ((it / NSTAGE) & 1) ^ 1, 0x000200 | slot, it);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&full_s[slot], SM::S_STAGE);
#pragma unroll
                        for (int rr = 0; rr < RCH; ++rr) {
                            uint8_t* dst = smem + SM::OFF_S + slot * SM::S_STAGE + rr * SM::S_SLAB;
                            if (NHWC) {  // [CN/64 channel groups][FBW x][64 channels]: one box per channel group
#pragma unroll
                                for (int cg = 0; cg < CN / 64; ++cg)
                                    tma_load_4d(dst + cg * (FBW * 128), &tmap_src, &full_s[slot], c0 + cg * 64, xmin + cb * FBW,
                                                ymin + rc * RCH + rr, b);
                            } else {     // [CN channels][FBW x]
                                tma_load_4d(dst, &tmap_src, &full_s[slot], xmin + cb * FBW, ymin + rc * RCH + rr, c0, b);
                            }
                        }
                    }
                    __syncwarp();
                }
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer
        constexpr uint32_t idesc = make_idesc_f16(128, CN, true, false, NHWC);
        uint32_t it = 0;
        int gi = 0;
        for (int g = blockIdx.x; g < ngroups; g += gridDim.x, ++gi) {
            mbar_wait(&info_full[gi % NINFO], (gi / NINFO) & 1, 0x010500, gi);
            const GroupInfo inf = infos[gi % NINFO];
            const int nst = inf.ncb * inf.nrc, buf = gi & 1;
            mbar_wait(&acc_empty[buf], ((gi >> 1) & 1) ^ 1, 0x010400 | buf, gi);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * CN;
            for (int st = 0; st < nst; ++st, ++it) {
                const int slot = it % NSTAGE;
                const uint32_t par = (it / NSTAGE) & 1;
                mbar_wait(&full_s[slot], par, 0x010000 | slot, it);
                mbar_wait(&full_a[slot], par, 0x010100 | slot, it);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a0 = smem_u32(smem + SM::OFF_A + slot * SM::A_STAGE);
                    const uint32_t b0 = smem_u32(smem + SM::OFF_S + slot * SM::S_STAGE);
#pragma unroll
                    for (int rr = 0; rr < RCH; ++rr)
#pragma unroll
                        for (int h = 0; h < FBW / 16; ++h) {  // K = 16 positions per MMA
                            // A = [128 px][FBW pos], K-major, rows of FBW*2 bytes with the matching swizzle; K-advance = +32 B
                            const uint64_t ad = FBW == 32 ? make_smem_desc(a0 + rr * SM::FA_SLAB + h * 32, 16, 512, kSwizzle64)
                                                          : make_smem_desc(a0 + rr * SM::FA_SLAB, 16, 256, kSwizzle32);
                            // NCHW: B = [CN rows][16 x], K-major, 32B swizzle.  NHWC: B = [FBW x][64 ch] per channel group,
                            // MN-major, 128B swizzle: LBO = next channel group, SBO = next 8 positions (1 KB); K-advance = 2 KB
                            const uint64_t bd = NHWC ? make_smem_desc(b0 + rr * SM::S_SLAB + h * 2048, FBW * 128, 1024, kSwizzle128)
                                                     : make_smem_desc(b0 + rr * SM::S_SLAB, 16, 256, kSwizzle32);
                            umma_f16(d_tmem, ad, bd, idesc, (st | rr | h) != 0 ? 1u : 0u);
                        }
                    tc_commit(&empty[slot]);
                    if (st == nst - 1) tc_commit(&acc_full[buf]);
                }
                __syncwarp();
            }
        }
    } else if (warp < 6) {
        // ================================================================= builders
        const int q = warp & 3, m = q * 32 + lane;  // pixel index inside the group
        const float inv_kk = 1.0f / static_cast<float>(KK);
        const uint32_t wsm_a = smem_u32(smem + SM::OFF_W) + m * 4;      // [18 words][128 pixels] packed bf16x2 rows
        const uint32_t a_base = smem_u32(smem + SM::OFF_A) + m * (FBW * 2);   // this pixel's row in slab 0
        // swizzle XOR of the 16B chunks of this row: 32B rows -> bit 2 of the row index, 64B rows -> bits 1-2
        const uint32_t swz = FBW == 32 ? (((m >> 1) & 3) << 4) : (((m >> 2) & 1) << 4);
        uint32_t it = 0, dirty = 0xffffffffu;   // slab rows start with unknown contents: treat them as dirty
        int gi = 0;
        for (int g = blockIdx.x; g < ngroups; g += gridDim.x, ++gi) {
            const int gx0 = (g % gxn) * GW, gy0 = ((g / gxn) % gyn) * GH, b = g / (gxn * gyn);
            const int px = gx0 + (m & 15), py = gy0 + (m >> 4);
            const bool valid = px < W && py < H;
            int X0 = 0, Y0 = 0;
            bool live = false;
            const long long tw0 = tc_profile_clock();
            if (valid && !(knobs & 512)) {   // bit 9: timing experiment, skip the per-pixel window construction
                const long long pofs = (long long)py * W + px;
                float p[KK];
                pixel_softmax_f32<KK>(logits + (long long)b * KK * hw + pofs, hw, p);
                if (probs != nullptr && blockIdx.y == 0) {
                    __nv_bfloat16* pr = probs + (long long)b * KK * hw + pofs;
#pragma unroll
                    for (int t = 0; t < KK; ++t) pr[t * hw] = __float2bfloat16_rn(p[t]);
                }
                const float fx = flow[(long long)b * 2 * hw + pofs], fy = flow[(long long)b * 2 * hw + hw + pofs];
                AxisTap<float> tx[K], ty[K];
                live = taps_regular<K>(fx, fy, px, py, Hs, Ws, tx, ty);
                if (live) {
                    float w[K1 * K1];
                    build_window<K>(p, tx, ty, Hs, Ws, inv_kk, w, X0, Y0);
                    store_window_words<K>(wsm_a, w);
                }
            }
            tc_profile_add(2, 6, tc_profile_clock() - tw0);          // window construction
            mbar_wait(&info_full[gi % NINFO], (gi / NINFO) & 1, 0x020500, gi);
            const GroupInfo inf = infos[gi % NINFO];
            long long fill_cycles = 0;
            for (int cb = 0; cb < inf.ncb; ++cb) {
                const int e0 = X0 - (inf.x0 + cb * FBW);           // box position of window column 0
                const bool cols_hit = live && e0 > -K1 && e0 < FBW && !(knobs & 1024);   // bit 10: timing experiment
                for (int rc = 0; rc < inf.nrc; ++rc, ++it) {
                    const int slot = it % NSTAGE;
                    mbar_wait(&empty[slot], ((it / NSTAGE) & 1) ^ 1, 0x020200 | slot, it);
                    const long long tf0 = tc_profile_clock();
                    const uint32_t a_stage = a_base + slot * SM::A_STAGE;
                    const int R0 = inf.y0 + rc * RCH;
                    bool wrote = false;
#pragma unroll
                    for (int rr = 0; rr < RCH; ++rr)
                        wrote |= fill_slab_row<K, FBW>(a_stage + rr * SM::FA_SLAB, swz, wsm_a, cols_hit, (R0 + rr) - Y0, e0, dirty,
                                                  1u << (slot * RCH + rr));
                    if (wrote) fence_proxy_async_smem();
                    mbar_arrive(&full_a[slot]);
                    fill_cycles += tc_profile_clock() - tf0;
                }
            }
            tc_profile_add(2, 7, fill_cycles);              // slab fills
        }
    } else {
        // ================================================================= epilogue
        const int q = warp & 3, m = q * 32 + lane;
        int gi = 0;
        for (int g = blockIdx.x; g < ngroups; g += gridDim.x, ++gi) {
            const int gx0 = (g % gxn) * GW, gy0 = ((g / gxn) % gyn) * GH, b = g / (gxn * gyn);
            const int px = gx0 + (m & 15), py = gy0 + (m >> 4);
            const bool valid = px < W && py < H;
            const long long pofs = (long long)py * W + px;
            bool regular = false;
            float fx = 0.f, fy = 0.f;
            if (valid) {
                fx = flow[(long long)b * 2 * hw + pofs];
                fy = flow[(long long)b * 2 * hw + hw + pofs];
                AxisTap<float> tx[K], ty[K];
                regular = taps_regular<K>(fx, fy, px, py, Hs, Ws, tx, ty);
            }
            const int buf = gi & 1;
            mbar_wait(&acc_full[buf], (gi >> 1) & 1, 0x030300 | buf, gi);
            tc_fence_after();
            const long long te0 = tc_profile_clock();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * CN;
            __nv_bfloat16* o = NHWC ? out + ((long long)b * hw + pofs) * C + c0 : out + ((long long)b * C + c0) * hw + pofs;
            // optional fused mask blend (generator.py:130): out = prev * (1 - mask) + attention * mask
            const __nv_bfloat16* pv = prev == nullptr ? nullptr
                                      : (NHWC ? prev + ((long long)b * hw + pofs) * C + c0 : prev + ((long long)b * C + c0) * hw + pofs);
            const float mk = (prev != nullptr && valid) ? __bfloat162float(mask[(long long)b * hw + pofs]) : 1.f;
#pragma unroll 1
            for (int cc = 0; cc < CN / 32; ++cc) {
                uint32_t v[32];
                tmem_ld_32x32(taddr + cc * 32, v);
                tmem_ld_wait();
                if (valid && regular && !(knobs & 256)) {   // bit 8: debug knob, skip the stores
                    if (pv != nullptr) {   // blend in fp32 before the single rounding to bf16
                        if (NHWC) {
                            const uint4* p4 = reinterpret_cast<const uint4*>(pv + cc * 32);
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const uint4 pq = p4[i];
                                const uint32_t pw[4] = {pq.x, pq.y, pq.z, pq.w};
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const float2 pf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pw[j]));
                                    v[8 * i + 2 * j] = __float_as_uint(pf.x * (1.f - mk) + __uint_as_float(v[8 * i + 2 * j]) * mk);
                                    v[8 * i + 2 * j + 1] = __float_as_uint(pf.y * (1.f - mk) + __uint_as_float(v[8 * i + 2 * j + 1]) * mk);
                                }
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                v[i] = __float_as_uint(__bfloat162float(pv[(long long)(cc * 32 + i) * hw]) * (1.f - mk) + __uint_as_float(v[i]) * mk);
                        }
                    }
                    if (NHWC) {
                        uint4* o4 = reinterpret_cast<uint4*>(o + cc * 32);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            uint4 pk;
                            __nv_bfloat162 t0 = __floats2bfloat162_rn(__uint_as_float(v[8 * i + 0]), __uint_as_float(v[8 * i + 1]));
                            __nv_bfloat162 t1 = __floats2bfloat162_rn(__uint_as_float(v[8 * i + 2]), __uint_as_float(v[8 * i + 3]));
                            __nv_bfloat162 t2 = __floats2bfloat162_rn(__uint_as_float(v[8 * i + 4]), __uint_as_float(v[8 * i + 5]));
                            __nv_bfloat162 t3 = __floats2bfloat162_rn(__uint_as_float(v[8 * i + 6]), __uint_as_float(v[8 * i + 7]));
                            pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
                            pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
                            __stcs(o4 + i, pk);   // streaming store: written once, never re-read by this kernel
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[(long long)(cc * 32 + i) * hw] = __float2bfloat16_rn(__uint_as_float(v[i]));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
            tc_profile_add(3, 6, tc_profile_clock() - te0);          // TMEM -> registers -> global
            // irregular pixels keep the reference's literal 4-tap arithmetic; the warp shares each one (tile_window.cuh)
            unsigned todo = __ballot_sync(0xffffffffu, valid && !regular);
            while (todo) {
                const int src_lane = __ffs(todo) - 1;
                todo &= todo - 1;
                const int qx = __shfl_sync(0xffffffffu, px, src_lane), qy = __shfl_sync(0xffffffffu, py, src_lane);
                const float qfx = __shfl_sync(0xffffffffu, fx, src_lane), qfy = __shfl_sync(0xffffffffu, fy, src_lane);
                irregular_pixel<K, NHWC>(src, logits, out, prev, mask, b, C, c0, CN, Hs, Ws, H, W, qx, qy, qfx, qfy, lane);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 2 * CN >= 32 ? 2 * CN : 32);
    tc_profile_total(t_start);
}

template <int K, int CN, bool NHWC>
static int launch_tc(const void* src, const void* flow, const void* logits, void* out, void* probs, const void* prev,
                     const void* mask, int B, int C, int Hs, int Ws, int H, int W, cudaStream_t st_) {
    static const PFN_tmapEncodeTiled enc = tmap_encoder();
    if (enc == nullptr) return GFLA_E_NOTSUP;
    CUtensorMap tmap;
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r;
    if (NHWC) {  // (c, x, y, b), box [64 c][16 x]: 128-byte runs, 128B swizzle
        const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)Ws, (cuuint64_t)Hs, (cuuint64_t)B};
        const cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)Ws * C * 2, (cuuint64_t)Hs * Ws * C * 2};
        const cuuint32_t box[4] = {64, SegW<true>::value, 1, 1};
        r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(src), gdim, gstr, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {     // (x, y, c, b), box [CN c][16 x]: 32-byte runs, 32B swizzle
        const cuuint64_t gdim[4] = {(cuuint64_t)Ws, (cuuint64_t)Hs, (cuuint64_t)C, (cuuint64_t)B};
        const cuuint64_t gstr[3] = {(cuuint64_t)Ws * 2, (cuuint64_t)Hs * Ws * 2, (cuuint64_t)C * Hs * Ws * 2};
        const cuuint32_t box[4] = {SegW<false>::value, 1, CN, 1};
        r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(src), gdim, gstr, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) return GFLA_E_NOTSUP;
    auto kern = k_local_attn_fwd_tc<K, CN, NHWC>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<CN, SegW<NHWC>::value>::ALLOC);
    if (e != cudaSuccess) return static_cast<int>(e);
    const int ngroups = B * ((H + GH - 1) / GH) * ((W + GW - 1) / GW);
    dim3 grid((unsigned)min(ngroups, sm_count()), (unsigned)(C / CN));
    kern<<<grid, NTHREADS, Smem<CN, SegW<NHWC>::value>::ALLOC, st_>>>(tmap, (const __nv_bfloat16*)src, (const float*)flow,
                                                   (const __nv_bfloat16*)logits, (__nv_bfloat16*)out,
                                                   (__nv_bfloat16*)probs, (const __nv_bfloat16*)prev,
                                                   (const __nv_bfloat16*)mask, B, C, Hs, Ws, H, W,
                                                   tune_knob("GFLA_TC_KNOBS", 0));
    return launch_status();
}

}  // namespace tc

int tc_wait_profile_fwd(int enable, unsigned long long* out32) { return tc::tc_wait_profile(enable, out32); }

int tc_debug_set_buffer(void* host_mapped) {
    unsigned long long* p = static_cast<unsigned long long*>(host_mapped);
    return static_cast<int>(cudaMemcpyToSymbol(tc::g_tc_dbg, &p, sizeof(p)));
}

static int pick_cn(int C) {
    if (C % 256 == 0) return 256;
    if (C == 128 || C == 64) return C;
    return 0;
}

bool local_attn_fwd_tc_supported(int B, int C, int Hs, int Ws, int H, int W, int k, int dtype, int flow_dtype,
                                 int layout, const void* src, const void* out) {
    (void)B; (void)H; (void)W; (void)Hs;
    if (!(dtype == GFLA_BF16 && flow_dtype == GFLA_F32 && (k == 3 || k == 5) && pick_cn(C) != 0 && aligned(src, 16)))
        return false;
    // NCHW: row pitch must be a multiple of 16 B for the tensor map; NHWC: pixel pitch 2*C always is, the
    // epilogue's 16-byte stores need `out` aligned
    return layout == GFLA_NHWC ? aligned(out, 16) : (Ws % 8) == 0;
}

int local_attn_fwd_strip_tc(const void*, const void*, const void*, void*, void*, const void*, const void*, int, int, int, int,
                            int, int, int, int, cudaStream_t);   // local_attn_strip_tc.cu

// Channels-last inputs run the strip schedule (local_attn_strip_tc.cu).  GFLA_TC_STRIP: -1 = per-tile kernel of this
// file instead, 0 = strip length chosen per launch, n > 0 = n tiles per strip.
constexpr int kStripDefault = 0;

int local_attn_fwd_tc(const void* src, const void* flow, const void* logits, void* out, void* probs, const void* prev,
                      const void* mask, int B, int C, int Hs, int Ws, int H, int W, int k, int dtype, int flow_dtype,
                      int layout, cudaStream_t st_) {
    if (!local_attn_fwd_tc_supported(B, C, Hs, Ws, H, W, k, dtype, flow_dtype, layout, src, out)) return GFLA_E_NOTSUP;
    const int cn = pick_cn(C);
    const bool nhwc = layout == GFLA_NHWC;
    const int strip = tc::tune_knob("GFLA_TC_STRIP", kStripDefault);
    if (nhwc && strip >= 0 && aligned(out, 32))   // the strip kernel's epilogue stores 32 bytes per lane
        return local_attn_fwd_strip_tc(src, flow, logits, out, probs, prev, mask, B, C, Hs, Ws, H, W, k, strip, st_);
#define GFLA_TC_CASE(K_, CN_)                                                                                        \
    if (k == K_ && cn == CN_)                                                                                        \
        return nhwc ? tc::launch_tc<K_, CN_, true>(src, flow, logits, out, probs, prev, mask, B, C, Hs, Ws, H, W, st_)  \
                    : tc::launch_tc<K_, CN_, false>(src, flow, logits, out, probs, prev, mask, B, C, Hs, Ws, H, W, st_);
    GFLA_TC_CASE(5, 256) GFLA_TC_CASE(5, 128) GFLA_TC_CASE(5, 64)
    GFLA_TC_CASE(3, 256) GFLA_TC_CASE(3, 128) GFLA_TC_CASE(3, 64)
#undef GFLA_TC_CASE
    return GFLA_E_NOTSUP;
}

}  // namespace gfla
