// Fused local-attention BACKWARD, part 2: grad_flow and grad_logits on the tensor cores
// (channels-last, bf16, C <= 256).
//
// Both gradients are functions of the per-pixel dot products
//
//   Q[p, t] = sum_c grad_out[b, p, c] * source[b, t, c]        t in the (k+1)x(k+1) window of pixel p
//
// (reference: block_extractor_kernel.cu:163-168 for d/dflow; the d/dlogits path is autograd through
// avg_pool2d * / LocalAttnReshape / Softmax, base_function.py:803-809).  On the CUDA cores that is
// (k+1)^2 gathered loads and FMAs per (pixel, channel).  Here it is the dense GEMM
//
//   D[128 pixels][64 positions] = G[128 pixels][C] * S[64 positions][C]^T
//
// per pipeline stage (2 source rows x 32 columns of the group's tap footprint): both operands are K-major
// (channel-contiguous) TMA boxes -- grad_out tile [128 px][64 ch] per channel group, source rows
// [64 positions][64 ch] per channel group -- so there is no operand construction at all.  Each epilogue
// thread (one per pixel) then picks its window entries out of the 64 accumulator columns and, after the
// last stage, turns them into d/dlogits (softmax backward) and d/dflow.
//
// Warp roles (6 warps): warp 0 producer (bbox + TMA), warp 1 MMA issuer, warps 2-5 epilogue.
#include "tile_window.cuh"

namespace gfla {
namespace tc {

constexpr int Q_ROWS = 2;          // source rows per stage: N = 2 x 32 = 64 positions (rows rounded to 2, one 32-wide
constexpr int Q_BW = 32;           // segment usually spans the whole footprint of a 16-pixel-wide group)
constexpr int Q_NS = 4;            // source-row stages (64 KB grad_out tile + 4 x 32 KB)
constexpr int Q_NACC = 4;          // accumulator buffers of 64 TMEM columns
constexpr int Q_NINFO = 8;
constexpr int Q_NTHREADS = 192;

template <int CN>
struct SmemQ {
    static constexpr int G_CG = 128 * 128;                 // [128 pixels][64 channels]
    static constexpr int G_BYTES = (CN / 64) * G_CG;
    static constexpr int S_CG = Q_ROWS * Q_BW * 128;       // [64 positions][64 channels] = 8 KB
    static constexpr int S_STAGE = (CN / 64) * S_CG;
    static constexpr int OFF_G = 0;
    static constexpr int OFF_S = OFF_G + G_BYTES;
    static constexpr int OFF_INFO = OFF_S + Q_NS * S_STAGE;
    static constexpr int OFF_BAR = OFF_INFO + Q_NINFO * 16;
    static constexpr int NBAR = 2 + 2 * Q_NS + 2 * Q_NACC + Q_NINFO;
    static constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
    static constexpr int ALLOC = OFF_TMEM + 16 + 1024;
};

template <int K, int CN>
__global__ void __launch_bounds__(Q_NTHREADS, 1)
k_local_attn_bwd_q_tc(const __grid_constant__ CUtensorMap tmap_g, const __grid_constant__ CUtensorMap tmap_s,
                      const __nv_bfloat16* __restrict__ src, const float* __restrict__ flow,
                      const __nv_bfloat16* __restrict__ logits, const __nv_bfloat16* __restrict__ gout,
                      float* __restrict__ gflow, __nv_bfloat16* __restrict__ glogits, int B, int C, int Hs, int Ws, int H,
                      int W, int accumulate) {
    using SM = SmemQ<CN>;
    constexpr int K1 = K + 1, KK = K * K;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::OFF_BAR);
    uint64_t* g_full = bars;
    uint64_t* g_empty = bars + 1;
    uint64_t* s_full = bars + 2;                       // [Q_NS]
    uint64_t* s_empty = s_full + Q_NS;                 // [Q_NS]
    uint64_t* acc_full = s_empty + Q_NS;               // [Q_NACC]
    uint64_t* acc_empty = acc_full + Q_NACC;           // [Q_NACC]
    uint64_t* info_full = acc_empty + Q_NACC;          // [Q_NINFO]
    GroupInfo* infos = reinterpret_cast<GroupInfo*>(smem + SM::OFF_INFO);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM::OFF_TMEM);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gxn = (W + GW - 1) / GW, gyn = (H + GH - 1) / GH;
    const int ngroups = B * gyn * gxn;
    const long long hw = (long long)H * W;

    if (threadIdx.x == 0) {
        mbar_init(g_full, 1);
        mbar_init(g_empty, 1);
        for (int i = 0; i < Q_NS; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 1); }
        for (int i = 0; i < Q_NACC; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
        for (int i = 0; i < Q_NINFO; ++i) mbar_init(&info_full[i], 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmap_g);
        tma_prefetch_desc(&tmap_s);
    }
    if (warp == 1) tmem_alloc(tmem_slot, Q_NACC * 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================================================= producer
        uint32_t it = 0;
        int gi = 0;
        for (int g = blockIdx.x; g < ngroups; g += gridDim.x, ++gi) {
            const int gx0 = (g % gxn) * GW, gy0 = ((g / gxn) % gyn) * GH, b = g / (gxn * gyn);
            int x0, y0, x1, y1;
            group_bbox<K>(flow, b, gx0, gy0, H, W, Hs, Ws, lane, false, x0, y0, x1, y1);
            const int ncb = (x1 - x0 + Q_BW) / Q_BW, nrc = (y1 - y0 + Q_ROWS) / Q_ROWS;
            if (lane == 0) {
                infos[gi % Q_NINFO] = GroupInfo{x0, y0, ncb, nrc};
                mbar_arrive(&info_full[gi % Q_NINFO]);
            }
            mbar_wait(g_empty, (gi & 1) ^ 1, 0x000600, gi);
            if (elect_one()) {
                mbar_arrive_expect_tx(g_full, SM::G_BYTES);
#pragma unroll
                for (int cg = 0; cg < CN / 64; ++cg)
                    tma_load_4d(smem + SM::OFF_G + cg * SM::G_CG, &tmap_g, g_full, cg * 64, gx0, gy0, b);
            }
            __syncwarp();
            for (int cb = 0; cb < ncb; ++cb)
                for (int rc = 0; rc < nrc; ++rc, ++it) {
                    const int slot = it % Q_NS;
                    mbar_wait(&s_empty[slot], ((it / Q_NS) & 1) ^ 1, 0x000200 | slot, it);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&s_full[slot], SM::S_STAGE);
#pragma unroll
                        for (int cg = 0; cg < CN / 64; ++cg)
                            tma_load_4d(smem + SM::OFF_S + slot * SM::S_STAGE + cg * SM::S_CG, &tmap_s, &s_full[slot], cg * 64,
                                        x0 + cb * Q_BW, y0 + rc * Q_ROWS, b);
                    }
                    __syncwarp();
                }
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer
        constexpr uint32_t idesc = make_idesc_f16(128, Q_ROWS * Q_BW, true, false, false);  // both K-major
        uint32_t it = 0;
        int gi = 0;
        for (int g = blockIdx.x; g < ngroups; g += gridDim.x, ++gi) {
            mbar_wait(&info_full[gi % Q_NINFO], (gi / Q_NINFO) & 1, 0x010500, gi);
            const GroupInfo inf = infos[gi % Q_NINFO];
            const int nst = inf.ncb * inf.nrc;
            mbar_wait(g_full, gi & 1, 0x010700, gi);
            for (int st = 0; st < nst; ++st, ++it) {
                const int slot = it % Q_NS, buf = it % Q_NACC;
                mbar_wait(&s_full[slot], (it / Q_NS) & 1, 0x010000 | slot, it);
                mbar_wait(&acc_empty[buf], ((it / Q_NACC) & 1) ^ 1, 0x010400 | buf, it);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a0 = smem_u32(smem + SM::OFF_G);
                    const uint32_t b0 = smem_u32(smem + SM::OFF_S + slot * SM::S_STAGE);
                    const uint32_t d_tmem = tmem_base + buf * 64;
#pragma unroll
                    for (int cg = 0; cg < CN / 64; ++cg)
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {  // 16 channels = 32 bytes inside the 128-byte swizzled row
                            const uint64_t ad = make_smem_desc(a0 + cg * SM::G_CG + kk * 32, 16, 1024, kSwizzle128);
                            const uint64_t bd = make_smem_desc(b0 + cg * SM::S_CG + kk * 32, 16, 1024, kSwizzle128);
                            umma_f16(d_tmem, ad, bd, idesc, (cg | kk) != 0 ? 1u : 0u);
                        }
                    tc_commit(&s_empty[slot]);
                    tc_commit(&acc_full[buf]);
                    if (st == nst - 1) tc_commit(g_empty);
                }
                __syncwarp();
            }
        }
    } else {
        // ================================================================= epilogue (thread <-> pixel)
        const int q = warp & 3, m = q * 32 + lane;
        const float inv_kk = 1.0f / static_cast<float>(KK);
        uint32_t it = 0;
        int gi = 0;
        for (int g = blockIdx.x; g < ngroups; g += gridDim.x, ++gi) {
            const int gx0 = (g % gxn) * GW, gy0 = ((g / gxn) % gyn) * GH, b = g / (gxn * gyn);
            const int px = gx0 + (m & 15), py = gy0 + (m >> 4);
            const bool valid = px < W && py < H;
            const long long pofs = (long long)py * W + px;
            float p[KK];
            AxisTap<float> tx[K], ty[K];
            float fx = 0.f, fy = 0.f;
            bool regular = false;
            if (valid) {
                pixel_softmax_f32<KK>(logits + (long long)b * KK * hw + pofs, hw, p);
                fx = flow[(long long)b * 2 * hw + pofs];
                fy = flow[(long long)b * 2 * hw + hw + pofs];
                regular = taps_regular<K>(fx, fy, px, py, Hs, Ws, tx, ty);
            }
            const bool live = valid && regular;
            const int X0 = live ? tx[0].fl : 0, Y0 = live ? ty[0].fl : 0;
            float Qw[K1 * K1];  // Q at the (clamped) window positions
#pragma unroll
            for (int i = 0; i < K1 * K1; ++i) Qw[i] = 0.f;

            mbar_wait(&info_full[gi % Q_NINFO], (gi / Q_NINFO) & 1, 0x030500, gi);
            const GroupInfo inf = infos[gi % Q_NINFO];
            for (int cb = 0; cb < inf.ncb; ++cb) {
                const int C0 = inf.x0 + cb * Q_BW;
                for (int rc = 0; rc < inf.nrc; ++rc, ++it) {
                    const int buf = it % Q_NACC, R0 = inf.y0 + rc * Q_ROWS;
                    mbar_wait(&acc_full[buf], (it / Q_NACC) & 1, 0x030300 | buf, it);
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * 64;
                    float v[64];
                    {
                        uint32_t u0[32], u1[32];
                        tmem_ld_32x32(taddr, u0);
                        tmem_ld_32x32(taddr + 32, u1);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) { v[i] = __uint_as_float(u0[i]); v[32 + i] = __uint_as_float(u1[i]); }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[buf]);
                    if (live) {
#pragma unroll
                        for (int r = 0; r < K1; ++r) {
                            const int rr = clampi(Y0 + r, Hs - 1) - R0;
                            if (rr >= 0 && rr < Q_ROWS) {
#pragma unroll
                                for (int c = 0; c < K1; ++c) {
                                    const int e = clampi(X0 + c, Ws - 1) - C0;
                                    if (e >= 0 && e < Q_BW) Qw[r * K1 + c] = v[rr * Q_BW + e];   // dynamic index: v lives in local memory
                                }
                            }
                        }
                    }
                }
            }
            // ---- finalize this pixel
            float dp[KK];
            float gfx = 0.f, gfy = 0.f;
            if (live) {
#pragma unroll
                for (int i = 0; i < K; ++i)
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        const float qLT = Qw[i * K1 + j], qRT = Qw[i * K1 + j + 1], qLB = Qw[(i + 1) * K1 + j], qRB = Qw[(i + 1) * K1 + j + 1];
                        dp[i * K + j] = inv_kk * (ty[i].wlo * (tx[j].wlo * qLT + tx[j].whi * qRT) + ty[i].whi * (tx[j].wlo * qLB + tx[j].whi * qRB));
                        const float pij = p[i * K + j] * inv_kk;
                        gfy += pij * (-tx[j].wlo * qLT - tx[j].whi * qRT + tx[j].wlo * qLB + tx[j].whi * qRB);
                        gfx += pij * (-ty[i].wlo * qLT - ty[i].whi * qLB + ty[i].wlo * qRT + ty[i].whi * qRB);
                    }
            }
            // irregular pixels: literal 4-tap dot products, the warp shares the channels of one pixel at a time
            unsigned todo = __ballot_sync(0xffffffffu, valid && !regular);
            while (todo) {
                const int sl = __ffs(todo) - 1;
                todo &= todo - 1;
                const int qx = __shfl_sync(0xffffffffu, px, sl), qy = __shfl_sync(0xffffffffu, py, sl);
                const float qfx = __shfl_sync(0xffffffffu, fx, sl), qfy = __shfl_sync(0xffffffffu, fy, sl);
                const long long qofs = (long long)qy * W + qx;
                const __nv_bfloat16* go = gout + ((long long)b * hw + qofs) * C;
                const __nv_bfloat16* sb = src + (long long)b * Hs * Ws * C;
                float gx_acc = 0.f, gy_acc = 0.f;
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    const AxisTap<float> ayy = axis_tap<float>(qfy, i - K / 2, qy, Hs);
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        const AxisTap<float> axx = axis_tap<float>(qfx, j - K / 2, qx, Ws);
                        float qLT = 0.f, qRT = 0.f, qLB = 0.f, qRB = 0.f;
                        for (int c = lane; c < C; c += 32) {
                            const float gv = __bfloat162float(go[c]);
                            qLT += gv * __bfloat162float(sb[((long long)ayy.lo * Ws + axx.lo) * C + c]);
                            qRT += gv * __bfloat162float(sb[((long long)ayy.lo * Ws + axx.hi) * C + c]);
                            qLB += gv * __bfloat162float(sb[((long long)ayy.hi * Ws + axx.lo) * C + c]);
                            qRB += gv * __bfloat162float(sb[((long long)ayy.hi * Ws + axx.hi) * C + c]);
                        }
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            qLT += __shfl_xor_sync(0xffffffffu, qLT, o); qRT += __shfl_xor_sync(0xffffffffu, qRT, o);
                            qLB += __shfl_xor_sync(0xffffffffu, qLB, o); qRB += __shfl_xor_sync(0xffffffffu, qRB, o);
                        }
                        if (lane == sl) {  // the owner keeps the results (its p[] is the right softmax)
                            dp[i * K + j] = inv_kk * (ayy.wlo * (axx.wlo * qLT + axx.whi * qRT) + ayy.whi * (axx.wlo * qLB + axx.whi * qRB));
                            const float pij = p[i * K + j] * inv_kk;
                            gy_acc += pij * (-axx.wlo * qLT - axx.whi * qRT + axx.wlo * qLB + axx.whi * qRB);
                            gx_acc += pij * (-ayy.wlo * qLT - ayy.whi * qLB + ayy.wlo * qRT + ayy.whi * qRB);
                        }
                    }
                }
                if (lane == sl) { gfx = gx_acc; gfy = gy_acc; }
            }
            if (valid) {
                float dot = 0.f;
#pragma unroll
                for (int t = 0; t < KK; ++t) dot += p[t] * dp[t];
                __nv_bfloat16* gl = glogits + (long long)b * KK * hw + pofs;
#pragma unroll
                for (int t = 0; t < KK; ++t) {
                    const float val = p[t] * (dp[t] - dot);
                    gl[t * hw] = __float2bfloat16_rn(accumulate ? __bfloat162float(gl[t * hw]) + val : val);
                }
                float* gf = gflow + (long long)b * 2 * hw + pofs;
                gf[0] = accumulate ? gf[0] + gfx : gfx;
                gf[hw] = accumulate ? gf[hw] + gfy : gfy;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, Q_NACC * 64);
}

template <int K, int CN>
static int launch_q(const void* src, const void* flow, const void* logits, const void* gout, void* gflow, void* glogits,
                    int B, int C, int Hs, int Ws, int H, int W, int accumulate, cudaStream_t st_) {
    static const PFN_tmapEncodeTiled enc = tmap_encoder();
    if (enc == nullptr) return GFLA_E_NOTSUP;
    CUtensorMap tg, ts;
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    {
        const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        const cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
        const cuuint32_t box[4] = {64, GW, GH, 1};
        if (enc(&tg, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(gout), gdim, gstr, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return GFLA_E_NOTSUP;
    }
    {
        const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)Ws, (cuuint64_t)Hs, (cuuint64_t)B};
        const cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)Ws * C * 2, (cuuint64_t)Hs * Ws * C * 2};
        const cuuint32_t box[4] = {64, Q_BW, Q_ROWS, 1};
        if (enc(&ts, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(src), gdim, gstr, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return GFLA_E_NOTSUP;
    }
    auto kern = k_local_attn_bwd_q_tc<K, CN>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemQ<CN>::ALLOC);
    if (e != cudaSuccess) return static_cast<int>(e);
    const int ngroups = B * ((H + GH - 1) / GH) * ((W + GW - 1) / GW);
    kern<<<(unsigned)min(ngroups, sm_count()), Q_NTHREADS, SmemQ<CN>::ALLOC, st_>>>(
        tg, ts, (const __nv_bfloat16*)src, (const float*)flow, (const __nv_bfloat16*)logits, (const __nv_bfloat16*)gout,
        (float*)gflow, (__nv_bfloat16*)glogits, B, C, Hs, Ws, H, W, accumulate);
    return launch_status();
}

}  // namespace tc

bool local_attn_bwd_q_tc_supported(int C, int k) { return (C == 64 || C == 128 || C == 256) && (k == 3 || k == 5); }

// grad_flow / grad_logits from tensor-core dot products (channels-last bf16 source / grad_out, fp32 flow, C <= 256)
int local_attn_bwd_q_tc(const void* src, const void* flow, const void* logits, const void* gout, void* gflow, void* glogits,
                        int B, int C, int Hs, int Ws, int H, int W, int k, int accumulate, cudaStream_t st_) {
#define GFLA_Q_CASE(K_, CN_) \
    if (k == K_ && C == CN_) return tc::launch_q<K_, CN_>(src, flow, logits, gout, gflow, glogits, B, C, Hs, Ws, H, W, accumulate, st_);
    GFLA_Q_CASE(5, 256) GFLA_Q_CASE(5, 128) GFLA_Q_CASE(5, 64)
    GFLA_Q_CASE(3, 256) GFLA_Q_CASE(3, 128) GFLA_Q_CASE(3, 64)
#undef GFLA_Q_CASE
    return GFLA_E_NOTSUP;
}

}  // namespace gfla
