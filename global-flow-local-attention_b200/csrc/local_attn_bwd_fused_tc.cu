// Fused local-attention BACKWARD in ONE kernel (channels-last bf16, C in {64, 128, 256}, k in {3, 5}):
// grad_source, grad_flow and grad_logits of
//
//   out[b,p,c] = (1/k^2) sum_ij softmax(logits[b,:,p])[ij] * bilinear(source[b,:,c], p + flow[b,p] + (i,j) - k/2)
//
// (reference: block_extractor_kernel.cu:89-170 for d/dsource and d/dflow; the d/dlogits path is autograd through
// avg_pool2d * / LocalAttnReshape / Softmax, base_function.py:803-809).
//
// Per 16x8 pixel group the grad_out tile G[128 px][C] is brought into shared memory ONCE (TMA) and feeds two
// tensor-core contractions that walk the group's tap footprint together:
//
//   Q stage   (1 source row x 32 columns):   Dq[128 px][32 pos]  = G[128 px][C] * S[32 pos][C]^T
//             G is copied once per group from shared memory into TENSOR memory (tcgen05.cp) and is the A operand of every
//             stage from there; B = the source row, a K-major TMA box.  The per-pixel dot products Q[p,t] at the (k+1)^2
//             window positions are all that grad_flow / grad_logits need (local_attn_bwd_q_tc.cu);
//   gs block  (4 source rows x 32 columns):  Dgs[128 pos][C]     = Wfull^T[128 pos][128 px] * G[128 px][C]
//             A = the forward kernel's weight slabs read MN-major, B = the same G tile read MN-major; the tile leaves
//             as a TMA REDUCE-ADD box into grad_source (local_attn_bwd_tc.cu).  The accumulator is split into two
//             channel halves so the epilogue drains one half while the tensor pipe fills the other.
//
// Warp roles (4 warpgroups, persistent CTA, static round-robin over pixel groups; setmaxnreg moves registers from the
// control warpgroup to the builders):
//   warp 0        producer: tap bounding box from the flow, TMA of the grad_out tile and of the source-row stages;
//   warp 1        MMA issuer of the Q stages, warp 2 MMA issuer of the gs blocks (2 channel halves each);   (warp 3 idle)
//   warps 4-7     Q extraction (thread = pixel): per Q stage TMEM -> thread-private shared-memory row -> picks its window
//                 entries with DYNAMIC SHARED addresses (no dynamically indexed registers, hence no local-memory stack);
//                 after the group's last stage the 36 values stay in the row for the builder thread of the same pixel;
//   warps 8-11    builders (thread = pixel): softmax, taps, collapsed window, weight slabs of every gs block -- and, one group
//                 later, softmax-backward and the d/dflow formula from the 36 window dot products -> grad_logits, grad_flow;
//   warps 12-15   gs epilogue (thread = source position): TMEM -> bf16 -> swizzled staging -> TMA reduce-add;
//                 irregular pixels (non-consecutive taps) are scattered here with vector reductions.
// What the time is made of (tools/ablate_bwd.py, profiles/r2_bwd_ablation.md): with every load, MMA, fill and store switched
// off, the barrier rings and schedule alone take a third of the kernel, and the per-pixel arithmetic (softmax, taps, window,
// softmax backward: ~2500 dependent instructions per pixel and group on one or two warps per scheduler) another third; the
// data movement and the tensor work hide behind them.  A fifth warpgroup for the finalize was tried and was slower (its register
// budget has to come out of the other four: setmaxnreg only redistributes the launch allocation).
// TMEM: 4 x 32 columns of Q accumulators, C/2 columns holding G as an MMA operand, 2 x C/2 (C = 64: 2 x 64) columns of
// grad_source accumulators.  Shared-memory bandwidth (128 B/clk: MMA operand fetch + TMA + staging) is the resource this kernel
// runs out of first (ncu: l1tex__data_pipe_{tc,lsu}_wavefronts_mem_shared); keeping G in TMEM removes the largest single reader.
#include "tile_window.cuh"

namespace gfla {
namespace tc {

constexpr int FB_BW = 32;            // source positions per row segment
constexpr int FB_QROWS = 1;          // source rows per Q stage (N = 32)
constexpr int FB_GROWS = 4;          // source rows per gs block (M = 128)
constexpr int FB_SLAB = 128 * FB_BW * 2;   // [128 pixels][32 positions] bf16, 64-byte rows, 64B swizzle
constexpr int FB_NINFO = 8;
constexpr int FB_THREADS = 512;       // 4 warpgroups: {producer, MMA, 2 idle}, pixel team, slab builders, gs epilogue
constexpr int FB_QS_STRIDE = 144;    // bytes per thread row of the Q staging (32 fp32 + 16: 16-byte stores of 8 lanes tile all banks)

// schedule of one pixel group (32-byte slots): tap bounding box origin, column blocks, source rows, and the width (24 / 28 / 32
// positions) of the LAST column block's reduce-add box -- the TMA reduce-adds are the slow end of the grad_source chain
// (L2 reduction throughput), so boxes are clipped to the footprint: the last column block to its width, the last row block to 2
// rows when the footprint ends in its first half.  (A grad_source block covers 4 rows; Q stages walk the rows one by one.)
struct FbInfo { int x0, y0, ncb, nrows, wlast, pad0, pad1, pad2; };
struct FbReduceMaps { CUtensorMap m[6]; };   // [width 24, 28, 32][rows 2, 4]

template <int CN>
struct SmemFB {
    static constexpr int NQ = 4;                               // Q accumulator buffers of 32 TMEM columns
    static constexpr int GA_COL0 = 128;                        // TMEM: grad_out tile as the A operand of the Q stages, CN/2 columns
    static constexpr int NS = 4;                               // source-row stages
    static constexpr int NA = CN == 256 ? 1 : 2;               // weight-slab stages (one gs block each)
    static constexpr int NH = CN >= 128 ? 2 : 1;               // channel halves of a gs block
    static constexpr int HN = CN / NH;                         // channels per half (multiple of 64)
    static constexpr int G_CG = 128 * 128;                     // [128 pixels][64 channels] bf16
    static constexpr int G_BYTES = (CN / 64) * G_CG;
    static constexpr int S_CG = FB_QROWS * FB_BW * 128;        // [32 positions][64 channels] = 4 KB
    static constexpr int S_STAGE = (CN / 64) * S_CG;
    static constexpr int A_STAGE = FB_GROWS * FB_SLAB;         // 4 slabs
    static constexpr int O_BUF = 128 * 128;                    // staging: [128 positions][64 channels] bf16
    static constexpr int OFF_G = 0;
    static constexpr int OFF_S = OFF_G + G_BYTES;
    static constexpr int OFF_A = OFF_S + NS * S_STAGE;
    static constexpr int OFF_O = OFF_A + NA * A_STAGE;
    static constexpr int OFF_W = OFF_O + 2 * O_BUF;
    static constexpr int OFF_QS = OFF_W + 36 * 128 * 2;
    static constexpr int OFF_INFO = OFF_QS + 128 * FB_QS_STRIDE;
    static constexpr int OFF_BAR = OFF_INFO + FB_NINFO * 32;
    static constexpr int NBAR = 2 + 2 * NS + 2 * NQ + 2 * NA + 4 + 2 + FB_NINFO;
    static constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
    static constexpr int ALLOC = OFF_TMEM + 16 + 1024;
    static constexpr int GS_COL0 = 256;                        // TMEM column of the first grad_source accumulator
};
static_assert(SmemFB<256>::ALLOC <= 232448, "shared memory budget");
static_assert(SmemFB<256>::OFF_O % 1024 == 0 && SmemFB<128>::OFF_O % 1024 == 0 && SmemFB<64>::OFF_O % 1024 == 0, "staging alignment");
static_assert(SmemFB<256>::OFF_A % 1024 == 0 && SmemFB<128>::OFF_A % 1024 == 0 && SmemFB<64>::OFF_A % 1024 == 0, "slab alignment");

__device__ __forceinline__ void fb_named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void fb_red_add_bf16x2(void* gptr, uint32_t v) {
    asm volatile("red.global.add.noftz.bf16x2 [%0], %1;" ::"l"(gptr), "r"(v) : "memory");
}
// register re-distribution between the warpgroups (all 4 warps of a warpgroup execute it; the CTA holds 512 x 128 registers)
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
constexpr int FB_REG_CTRL = 64, FB_REG_PIX = 112, FB_REG_FILL = 208, FB_REG_EPI = 128;   // sum = 512
static_assert(FB_REG_CTRL + FB_REG_PIX + FB_REG_FILL + FB_REG_EPI <= 512, "register budget");

__device__ __forceinline__ float lds_f32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}

// The four channel dot products <grad_out[p,:], source[tap,:]> of one tap of an irregular pixel, warp-cooperative.  Out of line:
// the k*k taps of the literal path are unrolled (their results index register arrays) and this body, inlined 25 times, was
// two thirds of the kernel's 34 k instructions -- for a path that ~1e-5 of the pixels take.
static __device__ __noinline__ float4 fb_tap_dots(const __nv_bfloat16* __restrict__ go, const __nv_bfloat16* __restrict__ sLT,
                                                  const __nv_bfloat16* __restrict__ sRT, const __nv_bfloat16* __restrict__ sLB,
                                                  const __nv_bfloat16* __restrict__ sRB, int C, int lane) {
    float qLT = 0.f, qRT = 0.f, qLB = 0.f, qRB = 0.f;
    for (int c = lane; c < C; c += 32) {
        const float gv = __bfloat162float(go[c]);
        qLT += gv * __bfloat162float(sLT[c]);
        qRT += gv * __bfloat162float(sRT[c]);
        qLB += gv * __bfloat162float(sLB[c]);
        qRB += gv * __bfloat162float(sRB[c]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        qLT += __shfl_xor_sync(0xffffffffu, qLT, o); qRT += __shfl_xor_sync(0xffffffffu, qRT, o);
        qLB += __shfl_xor_sync(0xffffffffu, qLB, o); qRB += __shfl_xor_sync(0xffffffffu, qRB, o);
    }
    return make_float4(qLT, qRT, qLB, qRB);
}

template <int K, int CN>
__global__ void __launch_bounds__(FB_THREADS, 1)
k_local_attn_bwd_fused(const __grid_constant__ CUtensorMap tmap_g, const __grid_constant__ CUtensorMap tmap_s,
                       const __grid_constant__ FbReduceMaps tmaps_gs, const __nv_bfloat16* __restrict__ src,
                       const float* __restrict__ flow, const __nv_bfloat16* __restrict__ logits,
                       const __nv_bfloat16* __restrict__ gout, __nv_bfloat16* __restrict__ gsrc, float* __restrict__ gflow,
                       __nv_bfloat16* __restrict__ glogits, int B, int C, int Hs, int Ws, int H, int W, int accumulate, int knobs_arg,
                       unsigned int* __restrict__ zero_flags) {
    const int knobs = GFLA_KNOBS(knobs_arg);      // 0 in the shipped build: every `knobs & x` test folds away
    // zero_flags != nullptr: grad_source arrives UNINITIALISED and is zero-filled here, sample by sample, by the otherwise idle
    // warp 3 of every CTA, at most two samples ahead of the CTA's own progress (so the zeros are still in L2 when the
    // reduce-adds land on them and reach HBM once); zero_flags[b] counts the CTAs that have finished their slice of sample b
    // and the epilogue waits for all of them before its first add into a sample.  All CTAs are co-resident (grid <= SM count,
    // one CTA per SM), and a CTA's zeroing never waits for another CTA, so the wait cannot deadlock.
    // `knobs` (environment GFLA_BWD_KNOBS, default 0 = production): bit 0 = also pull the next group's source rows into L2 ahead of
    // time.  Timing experiments (results are wrong when set): bit 1 pixel team skips the TMEM read / window picks, bit 2 gs epilogue
    // skips staging + reduce-add, bit 3 builders skip the slab fills, bit 4 no Q MMAs, bit 5 no grad_source MMAs, bit 6 no source-row loads,
    // bit 7 pixel team: no softmax / finalize / stores, bit 8 builders: no softmax / window, bit 9 gs epilogue: no irregular-pixel check and
    // no TMEM reads, bit 10 no grad_out tile load / TMEM copy.
    using SM = SmemFB<CN>;
    constexpr int K1 = K + 1, KK = K * K, NS = SM::NS, NA = SM::NA, NH = SM::NH, HN = SM::HN, FB_NQ = SM::NQ;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::OFF_BAR);
    uint64_t* g_full = bars;                         // grad_out tile landed
    uint64_t* g_empty = bars + 1;                    // all MMAs of the group retired
    uint64_t* s_full = bars + 2;                     // [NS] source-row stage landed
    uint64_t* s_empty = s_full + NS;                 // [NS]
    uint64_t* q_full = s_empty + NS;                 // [FB_NQ] Q accumulator complete
    uint64_t* q_empty = q_full + FB_NQ;              // [FB_NQ] 4 pixel-team warps drained it
    uint64_t* a_full = q_empty + FB_NQ;              // [NA] 128 builder arrivals
    uint64_t* a_empty = a_full + NA;                 // [NA]
    uint64_t* gs_full = a_empty + NA;                // [2] grad_source accumulator half complete
    uint64_t* gs_empty = gs_full + 2;                // [2] 4 epilogue warps drained it
    uint64_t* qw_full = gs_empty + 2;                // the pixel team's 36 window dot products of a group sit in its staging rows
    uint64_t* qw_empty = qw_full + 1;                // ... and have been read by the builders
    uint64_t* info_full = qw_empty + 1;              // [FB_NINFO]
    FbInfo* infos = reinterpret_cast<FbInfo*>(smem + SM::OFF_INFO);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM::OFF_TMEM);
    // progress of this CTA (sample its producer is working on), read by its zero-fill warp: a word of the workspace in GLOBAL
    // memory (a shared-memory word would do, but an unsynchronised shared word is a racecheck hazard by construction)
    volatile unsigned int* cur_sample = zero_flags != nullptr ? zero_flags + B + blockIdx.x : nullptr;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t_start = tc_profile_clock();
    const int gxn = (W + GW - 1) / GW, gyn = (H + GH - 1) / GH;
    const int ngroups = B * gyn * gxn;
    FastDiv fdx, fdy;
    fdx.init((uint32_t)gxn);
    fdy.init((uint32_t)gyn);
    auto split = [&](int g, int& gx, int& gy, int& b) {      // group index -> (column, row, sample) of the group
        uint32_t q, r, q2, r2;
        fdx.divmod((uint32_t)g, q, r);
        fdy.divmod(q, q2, r2);
        gx = (int)r; gy = (int)r2; b = (int)q2;
    };
    const long long hw = (long long)H * W;

    if (threadIdx.x == 0) {
        mbar_init(g_full, 1);
        mbar_init(g_empty, 2);      // one commit from each MMA-issuing warp
        for (int i = 0; i < NS; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 1); }
        for (int i = 0; i < FB_NQ; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 4); }
        for (int i = 0; i < NA; ++i) { mbar_init(&a_full[i], 128); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&gs_full[i], 1); mbar_init(&gs_empty[i], 4); }
        mbar_init(qw_full, 128);
        mbar_init(qw_empty, 128);
        for (int i = 0; i < FB_NINFO; ++i) mbar_init(&info_full[i], 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmap_g);
        tma_prefetch_desc(&tmap_s);
        for (int i = 0; i < 6; ++i) tma_prefetch_desc(&tmaps_gs.m[i]);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
      reg_dec<FB_REG_CTRL>();
      if (warp == 0) {
        // ================================================================= producer
        // Per group: publish the schedule, queue the first source-row stages (their slots free up while the previous
        // group's tail is still in the tensor pipe), then -- once every MMA of the previous group has retired -- the
        // grad_out tile, then the remaining stages.  The flow of the NEXT group is loaded before the stage loop and reduced
        // to its bounding box after it, so that latency never sits between two groups; its grad_out tile and source rows
        // are then pulled into L2 (with only NS stages in flight, a stage that had to come from HBM would expose the whole
        // DRAM latency once per stage).
        uint32_t it = 0;
        int gi = 0;
        int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
        if (blockIdx.x < ngroups) {
            int gx, gy, b;
            split((int)blockIdx.x, gx, gy, b);
            group_bbox<K>(flow, b, gx * GW, gy * GH, H, W, Hs, Ws, lane, false, x0, y0, x1, y1);
        }
        for (int g = blockIdx.x; g < ngroups; g += gridDim.x, ++gi) {
            int gxi, gyi, b;
            split(g, gxi, gyi, b);
            const int gx0 = gxi * GW, gy0 = gyi * GH;
            const int ncb = (x1 - x0 + FB_BW) / FB_BW, nrows = y1 - y0 + 1, nst = ncb * nrows;
            if (lane == 0) {
                if (cur_sample != nullptr) *cur_sample = (unsigned int)b;
                const int wl = x1 - (x0 + FB_BW * (ncb - 1)) + 1;
                infos[gi % FB_NINFO] = FbInfo{x0, y0, ncb, nrows, wl <= 24 ? 24 : (wl <= 28 ? 28 : 32), 0, 0, 0};
                mbar_arrive(&info_full[gi % FB_NINFO]);
            }
            const int gn = g + gridDim.x;
            const bool has_next = gn < ngroups;
            int ngxi = 0, ngyi = 0, nb = 0;
            if (has_next) split(gn, ngxi, ngyi, nb);
            const int ngx0 = ngxi * GW, ngy0 = ngyi * GH;
            TileFlow nf;
            if (has_next) tile_flow_load(flow, nb, ngx0, ngy0, H, W, lane, nf);
            const int cx0 = x0, cy0 = y0;
            int st_cb = 0, st_rc = 0;                           // column block / row of the next stage (stages run down the rows of a block)
            auto load_stage = [&](int) {
                const int cb = st_cb, rc = st_rc, slot = it % NS;
                if (++st_rc == nrows) { st_rc = 0; ++st_cb; }
                mbar_wait(&s_empty[slot], ((it / NS) & 1) ^ 1, 0x000200 | slot, it);
                if ((knobs & 64) && lane == 0) mbar_arrive(&s_full[slot]);
                if (!(knobs & 64) && elect_one()) {
                    mbar_arrive_expect_tx(&s_full[slot], SM::S_STAGE);
#pragma unroll
                    for (int cg = 0; cg < CN / 64; ++cg)
                        tma_load_4d(smem + SM::OFF_S + slot * SM::S_STAGE + cg * SM::S_CG, &tmap_s, &s_full[slot], cg * 64,
                                    cx0 + cb * FB_BW, cy0 + rc * FB_QROWS, b);
                }
                __syncwarp();
                ++it;
            };
            int s = 0;
            for (; s < nst && s < NS; ++s) load_stage(s);
            mbar_wait(g_empty, (gi & 1) ^ 1, 0x000600, gi);
            if ((knobs & 1024) && lane == 0) mbar_arrive(g_full);
            if (!(knobs & 1024) && elect_one()) {
                mbar_arrive_expect_tx(g_full, SM::G_BYTES);
#pragma unroll
                for (int cg = 0; cg < CN / 64; ++cg)
                    tma_load_4d(smem + SM::OFF_G + cg * SM::G_CG, &tmap_g, g_full, cg * 64, gx0, gy0, b);
                if (has_next) {
#pragma unroll
                    for (int cg = 0; cg < CN / 64; ++cg) tma_prefetch_4d(&tmap_g, cg * 64, ngx0, ngy0, nb);
                }
            }
            __syncwarp();
            for (; s < nst; ++s) load_stage(s);
            if (has_next) {
                tile_bbox_reduce<K>(nf, ngx0, ngy0, H, W, Hs, Ws, lane, false, x0, y0, x1, y1);
                const int pcb = (x1 - x0 + FB_BW) / FB_BW, prc = (knobs & 1) ? y1 - y0 + 1 : 0;
                for (int i = lane; i < pcb * prc * (CN / 64); i += 32) {
                    const int cg = i % (CN / 64), st = i / (CN / 64), cb = st / max(prc, 1), rc = st - cb * prc;
                    tma_prefetch_4d(&tmap_s, cg * 64, x0 + cb * FB_BW, y0 + rc * FB_QROWS, nb);
                }
            }
        }
        // no groups left: release the zero-fill warp for the remaining samples (other CTAs wait for this CTA's slice of them)
        if (lane == 0 && cur_sample != nullptr) *cur_sample = (unsigned int)B;
      } else if (warp == 1) {
        // ================================================================= MMA issuer 1: Q stages
        // Two issuing warps, one per contraction: each blocks only on its own chain's barriers, so a grad_source block
        // waiting for its epilogue (the TMA reduce-adds are the slow end of that chain) never holds back the Q stages.
        constexpr uint32_t idesc_q = make_idesc_f16(128, FB_QROWS * FB_BW, true, false, false);   // A from TMEM, B K-major
        uint32_t it = 0;
        int gi = 0;
        for (int g = blockIdx.x; g < ngroups; g += gridDim.x, ++gi) {
            mbar_wait(&info_full[gi % FB_NINFO], (gi / FB_NINFO) & 1, 0x010500, gi);
            const FbInfo inf = infos[gi % FB_NINFO];
            const int nst = inf.ncb * inf.nrows;
            mbar_wait(g_full, gi & 1, 0x010700, gi);
            const uint32_t g0 = smem_u32(smem + SM::OFF_G);
            tc_fence_after();
            if (elect_one()) {
                // G[128 px][CN] -> TMEM, 16 channels (8 columns) per copy; ordered before the MMAs below, after those of the last group
                if (!(knobs & 1024))
#pragma unroll
                for (int cg = 0; cg < CN / 64; ++cg)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        tmem_cp_128x256b(tmem_base + SM::GA_COL0 + (cg * 4 + kk) * 8,
                                         make_smem_desc(g0 + cg * SM::G_CG + kk * 32, 16, 1024, kSwizzle128));
                tc_commit(g_empty);        // the Q chain is done with the shared-memory copy of the tile
            }
            __syncwarp();
            for (int s = 0; s < nst; ++s, ++it) {
                const int slot = it % NS, buf = it % FB_NQ;
                mbar_wait(&s_full[slot], (it / NS) & 1, 0x010000 | slot, it);
                mbar_wait(&q_empty[buf], ((it / FB_NQ) & 1) ^ 1, 0x010400 | buf, it);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t b0 = smem_u32(smem + SM::OFF_S + slot * SM::S_STAGE);
                    const uint32_t d_tmem = tmem_base + buf * 32;
                    if (!(knobs & 16))
#pragma unroll
                    for (int cg = 0; cg < CN / 64; ++cg)
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)     // 16 channels = 32 bytes inside the 128-byte swizzled row
                            umma_f16_ts(d_tmem, tmem_base + SM::GA_COL0 + (cg * 4 + kk) * 8,
                                        make_smem_desc(b0 + cg * SM::S_CG + kk * 32, 16, 1024, kSwizzle128), idesc_q, (cg | kk) != 0 ? 1u : 0u);
                    tc_commit(&s_empty[slot]);
                    tc_commit(&q_full[buf]);
                }
                __syncwarp();
            }
        }
      } else if (warp == 2) {
        // ================================================================= MMA issuer 2: grad_source blocks (Wfull^T * G, one accumulator per channel half)
        constexpr uint32_t idesc_gs = make_idesc_f16(128, HN, true, true, true);                  // both MN-major
        uint32_t blk = 0, u = 0;
        int gi = 0;
        for (int g = blockIdx.x; g < ngroups; g += gridDim.x, ++gi) {
            mbar_wait(&info_full[gi % FB_NINFO], (gi / FB_NINFO) & 1, 0x050500, gi);
            const FbInfo inf = infos[gi % FB_NINFO];
            const int nblk = inf.ncb * ((inf.nrows + FB_GROWS - 1) / FB_GROWS);
            mbar_wait(g_full, gi & 1, 0x050700, gi);
            const uint32_t g0 = smem_u32(smem + SM::OFF_G);
            for (int bi = 0; bi < nblk; ++bi, ++blk) {
                const int st = blk % NA;
                mbar_wait(&a_full[st], (blk / NA) & 1, 0x050100 | st, blk);
#pragma unroll
                for (int hf = 0; hf < NH; ++hf, ++u) {
                    const int buf = u & 1;
                    mbar_wait(&gs_empty[buf], ((u >> 1) & 1) ^ 1, 0x050600 | buf, u);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t a0 = smem_u32(smem + SM::OFF_A + st * SM::A_STAGE);
                        const uint32_t d_tmem = tmem_base + SM::GS_COL0 + buf * HN;
                        if (!(knobs & 32))
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks) {   // 16 pixels per MMA
                            // A^T: M = positions (32 per slab, LBO = next slab), K = pixels (8 per 512-byte atom)
                            const uint64_t ad = make_smem_desc(a0 + ks * 1024, FB_SLAB, 512, kSwizzle64);
                            // B: N = channels of this half (64 per 128-byte row, LBO = next channel group), K = pixels
                            const uint64_t bd = make_smem_desc(g0 + hf * (HN / 64) * SM::G_CG + ks * 2048, SM::G_CG, 1024, kSwizzle128);
                            umma_f16(d_tmem, ad, bd, idesc_gs, ks != 0 ? 1u : 0u);
                        }
                        tc_commit(&gs_full[buf]);
                        if (hf == NH - 1) {
                            tc_commit(&a_empty[st]);
                            if (bi == nblk - 1) tc_commit(g_empty);
                        }
                    }
                    __syncwarp();
                }
            }
        }
      } else if (warp == 3 && zero_flags != nullptr) {
        // ================================================================= zero-fill of grad_source, one slice per CTA and sample
        const long long per_sample = (long long)Hs * Ws * C * 2;                    // bytes (multiple of 128: C is a multiple of 64)
        const long long slice = ((per_sample / 16 + gridDim.x - 1) / gridDim.x) * 16;
        const long long lo = min(per_sample, slice * blockIdx.x), hi = min(per_sample, lo + slice);
        const int b_first = (int)(blockIdx.x / (gxn * gyn));                        // sample of this CTA's first group
        for (int zb = 0; zb < B; ++zb) {
            while (zb > max((int)*cur_sample, b_first) + 2) __nanosleep(256);       // stay at most two samples ahead
            char* base = reinterpret_cast<char*>(gsrc) + (long long)zb * per_sample;
            for (long long o = lo + lane * 16; o < hi; o += 512)
                asm volatile("st.global.v4.b32 [%0], {%1, %1, %1, %1};" ::"l"(base + o), "r"(0u) : "memory");
            __threadfence();
            __syncwarp();
            if (lane == 0) atomicAdd(&zero_flags[zb], 1u);
        }
      }   // (warp 3 otherwise idle: it pads the control warpgroup so that setmaxnreg can hand its registers on)
    } else if (warp < 8) {
        // ================================================================= pixel team: the (k+1)^2 window dot products Q of every pixel
        // Extraction only: per Q stage TMEM -> thread-private staging row -> picks with dynamic shared addresses.  After the
        // group's last stage the 36 values are left in the staging row (it is exactly 36 floats wide) for the builder thread
        // of the same pixel, which owns the softmax and turns them into grad_logits / grad_flow: the per-pixel softmax,
        // tap and softmax-backward arithmetic used to sit on this team's -- i.e. the Q chain's -- critical path (0.31 ms of 1.06).
        reg_dec<FB_REG_PIX>();
        const int q = warp & 3, m = q * 32 + lane;
        const uint32_t qs_row = smem_u32(smem + SM::OFF_QS) + m * FB_QS_STRIDE;     // thread-private staging row
        uint32_t it = 0;
        int gi = 0;
        float pfx = 0.f, pfy = 0.f;          // flow of this thread's pixel, loaded one group ahead
        auto load_pixel = [&](int g) {
            int gxi, gyi, b;
            split(g, gxi, gyi, b);
            const int px = gxi * GW + (m & 15), py = gyi * GH + (m >> 4);
            if (px < W && py < H) {
                const long long pofs = (long long)py * W + px;
                pfx = flow[(long long)b * 2 * hw + pofs];
                pfy = flow[(long long)b * 2 * hw + hw + pofs];
            }
        };
        if (blockIdx.x < ngroups) load_pixel(blockIdx.x);
        for (int g = blockIdx.x; g < ngroups; g += gridDim.x, ++gi) {
            int gxi, gyi, b_unused;
            split(g, gxi, gyi, b_unused);
            const int gx0 = gxi * GW, gy0 = gyi * GH;
            const int px = gx0 + (m & 15), py = gy0 + (m >> 4);
            const bool live = px < W && py < H;      // irregular pixels are extracted too (their values are ignored by the builders)
            // window origin = unclamped floor of the first tap (block_extractor_kernel.cu:62-66)
            const int X0 = live ? axis_tap<float>(pfx, -(K / 2), px, Ws).fl : 0, Y0 = live ? axis_tap<float>(pfy, -(K / 2), py, Hs).fl : 0;
            if (g + (int)gridDim.x < ngroups) load_pixel(g + gridDim.x);
            float Qw[K1 * K1];  // Q at the (clamped) window positions
#pragma unroll
            for (int i = 0; i < K1 * K1; ++i) Qw[i] = 0.f;

            mbar_wait(&info_full[gi % FB_NINFO], (gi / FB_NINFO) & 1, 0x020500, gi);
            const FbInfo inf = infos[gi % FB_NINFO];
            mbar_wait(qw_empty, (gi & 1) ^ 1, 0x020600, gi);      // the builders have read the previous group's values out of the staging rows
            for (int cb = 0; cb < inf.ncb; ++cb) {
                const int C0 = inf.x0 + cb * FB_BW;
                for (int rc = 0; rc < inf.nrows; ++rc, ++it) {
                    const int buf = it % FB_NQ, R0 = inf.y0 + rc;
                    // does any pixel of this warp need this row?  (warp-uniform: a row nobody needs is never read)
                    bool need = false;
                    if (live) {
#pragma unroll
                        for (int r = 0; r < K1; ++r) need = need || (clampi(Y0 + r, Hs - 1) == R0);
                        need = need && (clampi(X0 + K, Ws - 1) >= C0) && (clampi(X0, Ws - 1) < C0 + FB_BW);
                    }
                    const unsigned any = __ballot_sync(0xffffffffu, need);
                    mbar_wait(&q_full[buf], (it / FB_NQ) & 1, 0x020300 | buf, it);
                    if (any != 0u && !(knobs & 2)) {
                        tc_fence_after();
                        uint32_t v[32];
                        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 8; ++i) sts128(qs_row + i * 16, v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                        if (need) {
#pragma unroll
                            for (int r = 0; r < K1; ++r) {
                                if (clampi(Y0 + r, Hs - 1) == R0) {
#pragma unroll
                                    for (int c = 0; c < K1; ++c) {
                                        const int e = clampi(X0 + c, Ws - 1) - C0;
                                        if (e >= 0 && e < FB_BW) Qw[r * K1 + c] = lds_f32(qs_row + e * 4);
                                    }
                                }
                            }
                        }
                        tc_fence_before();
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&q_empty[buf]);
                }
            }
            // hand the window over: 36 floats = the whole staging row
#pragma unroll
            for (int i = 0; i < K1 * K1 / 4; ++i)
                sts128(qs_row + i * 16, __float_as_uint(Qw[4 * i]), __float_as_uint(Qw[4 * i + 1]), __float_as_uint(Qw[4 * i + 2]), __float_as_uint(Qw[4 * i + 3]));
            mbar_arrive(qw_full);
        }
    } else if (warp < 12) {
        // ================================================================= builders (thread = pixel): softmax, taps, collapsed window,
        // the weight slabs of every grad_source block -- and, one group later, the pixel's grad_logits / grad_flow from the window dot
        // products the pixel team left in its staging row.  Order per group g: window(g) -> finalize(g-1) -> slabs(g): the window does
        // not depend on the Q chain, so only the finalize sits between "Q chain of g-1 done" and "first slab of g".
        reg_inc<FB_REG_FILL>();
        const int q = warp & 3, m = q * 32 + lane;
        const float inv_kk = 1.0f / static_cast<float>(KK);
        const uint32_t wsm_a = smem_u32(smem + SM::OFF_W) + m * 4;
        const uint32_t a_base = smem_u32(smem + SM::OFF_A) + m * (FB_BW * 2);
        const uint32_t qs_row = smem_u32(smem + SM::OFF_QS) + m * FB_QS_STRIDE;     // the pixel team's staging row of this pixel
        const uint32_t swz = ((m >> 1) & 3) << 4;   // 64B swizzle: 16B chunk ^= bits 1-2 of the row
        uint32_t blk = 0, dirty = 0xffffffffu;
        int gi = 0;
        __nv_bfloat16 lg[KK];
        float pfx = 0.f, pfy = 0.f;
        auto load_pixel = [&](int g) {
            int gxi, gyi, b;
            split(g, gxi, gyi, b);
            const int px = gxi * GW + (m & 15), py = gyi * GH + (m >> 4);
            if (px < W && py < H) {
                const long long pofs = (long long)py * W + px;
                const __nv_bfloat16* lp = logits + (long long)b * KK * hw + pofs;
#pragma unroll
                for (int t = 0; t < KK; ++t) lg[t] = lp[t * hw];
                pfx = flow[(long long)b * 2 * hw + pofs];
                pfy = flow[(long long)b * 2 * hw + hw + pofs];
            }
        };
        // state of the previous group's pixel, kept for its finalize
        float pp[KK];
        float fx_p = 0.f, fy_p = 0.f;
        int px_p = 0, py_p = 0, b_p = 0;
        bool valid_p = false, regular_p = false, have_p = false;
        // grad_logits / grad_flow of the previous group's pixel (softmax backward; d/dflow as block_extractor_kernel.cu:163-168)
        auto finalize = [&](int gip) {
            const long long tq0 = tc_profile_clock();
            mbar_wait(qw_full, gip & 1, 0x030700, gip);
            float Qw[K1 * K1];
#pragma unroll
            for (int i = 0; i < K1 * K1 / 4; ++i) lds128(qs_row + i * 16, Qw[4 * i], Qw[4 * i + 1], Qw[4 * i + 2], Qw[4 * i + 3]);
            mbar_arrive(qw_empty);
            if (knobs & 128) return;
            const long long pofs = (long long)py_p * W + px_p;
            float dp[KK];
            float gfx = 0.f, gfy = 0.f;
            if (valid_p && regular_p) {
                AxisTap<float> tx[K], ty[K];
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    tx[j] = axis_tap<float>(fx_p, j - K / 2, px_p, Ws);
                    ty[j] = axis_tap<float>(fy_p, j - K / 2, py_p, Hs);
                }
#pragma unroll
                for (int i = 0; i < K; ++i)
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        const float qLT = Qw[i * K1 + j], qRT = Qw[i * K1 + j + 1], qLB = Qw[(i + 1) * K1 + j], qRB = Qw[(i + 1) * K1 + j + 1];
                        dp[i * K + j] = inv_kk * (ty[i].wlo * (tx[j].wlo * qLT + tx[j].whi * qRT) + ty[i].whi * (tx[j].wlo * qLB + tx[j].whi * qRB));
                        const float pij = pp[i * K + j] * inv_kk;
                        gfy += pij * (-tx[j].wlo * qLT - tx[j].whi * qRT + tx[j].wlo * qLB + tx[j].whi * qRB);
                        gfx += pij * (-ty[i].wlo * qLT - ty[i].whi * qLB + ty[i].wlo * qRT + ty[i].whi * qRB);
                    }
            }
            // irregular pixels: literal 4-tap dot products, the warp shares the channels of one pixel at a time
            unsigned todo = __ballot_sync(0xffffffffu, valid_p && !regular_p);
            while (todo) {
                const int sl = __ffs(todo) - 1;
                todo &= todo - 1;
                const int qx = __shfl_sync(0xffffffffu, px_p, sl), qy = __shfl_sync(0xffffffffu, py_p, sl);
                const float qfx = __shfl_sync(0xffffffffu, fx_p, sl), qfy = __shfl_sync(0xffffffffu, fy_p, sl);
                const long long qofs = (long long)qy * W + qx;
                const __nv_bfloat16* go = gout + ((long long)b_p * hw + qofs) * C;
                const __nv_bfloat16* sb = src + (long long)b_p * Hs * Ws * C;
                float gx_acc = 0.f, gy_acc = 0.f;
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    const AxisTap<float> ayy = axis_tap<float>(qfy, i - K / 2, qy, Hs);
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        const AxisTap<float> axx = axis_tap<float>(qfx, j - K / 2, qx, Ws);
                        const float4 qd = fb_tap_dots(go, sb + ((long long)ayy.lo * Ws + axx.lo) * C, sb + ((long long)ayy.lo * Ws + axx.hi) * C,
                                                      sb + ((long long)ayy.hi * Ws + axx.lo) * C, sb + ((long long)ayy.hi * Ws + axx.hi) * C, C, lane);
                        const float qLT = qd.x, qRT = qd.y, qLB = qd.z, qRB = qd.w;
                        if (lane == sl) {  // the owner keeps the results (its pp[] is the right softmax)
                            dp[i * K + j] = inv_kk * (ayy.wlo * (axx.wlo * qLT + axx.whi * qRT) + ayy.whi * (axx.wlo * qLB + axx.whi * qRB));
                            const float pij = pp[i * K + j] * inv_kk;
                            gy_acc += pij * (-axx.wlo * qLT - axx.whi * qRT + axx.wlo * qLB + axx.whi * qRB);
                            gx_acc += pij * (-ayy.wlo * qLT - ayy.whi * qLB + ayy.wlo * qRT + ayy.whi * qRB);
                        }
                    }
                }
                if (lane == sl) { gfx = gx_acc; gfy = gy_acc; }
            }
            if (valid_p) {
                float dot = 0.f;
#pragma unroll
                for (int t = 0; t < KK; ++t) dot += pp[t] * dp[t];
                __nv_bfloat16* gl = glogits + (long long)b_p * KK * hw + pofs;
#pragma unroll
                for (int t = 0; t < KK; ++t) {
                    const float val = pp[t] * (dp[t] - dot);
                    gl[t * hw] = __float2bfloat16_rn(accumulate ? __bfloat162float(gl[t * hw]) + val : val);
                }
                float* gf = gflow + (long long)b_p * 2 * hw + pofs;
                gf[0] = accumulate ? gf[0] + gfx : gfx;
                gf[hw] = accumulate ? gf[hw] + gfy : gfy;
            }
            tc_profile_add(3, 7, tc_profile_clock() - tq0);          // wait for Q, softmax backward, d/dflow, stores
        };
        if (blockIdx.x < ngroups) load_pixel(blockIdx.x);
        for (int g = blockIdx.x; g < ngroups; g += gridDim.x, ++gi) {
            int gxi, gyi, b;
            split(g, gxi, gyi, b);
            const int gx0 = gxi * GW, gy0 = gyi * GH;
            const int px = gx0 + (m & 15), py = gy0 + (m >> 4);
            const bool valid = px < W && py < H;
            int X0 = 0, Y0 = 0;
            bool live = false;
            const long long tw0 = tc_profile_clock();
            float p[KK];
            if (valid && !(knobs & 256)) {
#pragma unroll
                for (int t = 0; t < KK; ++t) p[t] = __bfloat162float(lg[t]);
                softmax_inplace_f32<KK>(p);
                AxisTap<float> tx[K], ty[K];
                live = taps_regular<K>(pfx, pfy, px, py, Hs, Ws, tx, ty);
                if (live) {
                    float w[K1 * K1];
                    build_window<K>(p, tx, ty, Hs, Ws, inv_kk, w, X0, Y0);
                    store_window_words<K>(wsm_a, w);
                }
            }
            const float fx_c = pfx, fy_c = pfy;
            if (g + (int)gridDim.x < ngroups && !(knobs & 256)) load_pixel(g + gridDim.x);
            tc_profile_add(3, 6, tc_profile_clock() - tw0);          // window of this group, raw loads of the next
            if (have_p) finalize(gi - 1);
            mbar_wait(&info_full[gi % FB_NINFO], (gi / FB_NINFO) & 1, 0x030500, gi);
            const FbInfo inf = infos[gi % FB_NINFO];
            const int nb4 = (inf.nrows + FB_GROWS - 1) / FB_GROWS;
            for (int cb = 0; cb < inf.ncb; ++cb) {
                const int e0 = X0 - (inf.x0 + cb * FB_BW);
                const bool cols_hit = live && e0 > -K1 && e0 < FB_BW;
                for (int rb = 0; rb < nb4; ++rb, ++blk) {
                    const int st = blk % NA;
                    mbar_wait(&a_empty[st], ((blk / NA) & 1) ^ 1, 0x030200 | st, blk);
                    const uint32_t a_stage = a_base + st * SM::A_STAGE;
                    const int R0 = inf.y0 + rb * FB_GROWS;
                    bool wrote = false;
                    if (!(knobs & 8)) {
#pragma unroll
                    for (int seg = 0; seg < FB_GROWS; ++seg)
                        wrote |= fill_slab_row<K, FB_BW>(a_stage + seg * FB_SLAB, swz, wsm_a, cols_hit, (R0 + seg) - Y0, e0, dirty,
                                                         1u << (st * FB_GROWS + seg));
                    }
                    if (wrote) fence_proxy_async_smem();
                    mbar_arrive(&a_full[st]);
                }
            }
            // this group becomes the "previous" one
#pragma unroll
            for (int t = 0; t < KK; ++t) pp[t] = p[t];
            fx_p = fx_c; fy_p = fy_c; px_p = px; py_p = py; b_p = b;
            valid_p = valid; regular_p = live; have_p = true;
        }
        if (have_p) finalize(gi - 1);
    } else {
        // ================================================================= grad_source epilogue (thread = position of the block)
        const int q = warp & 3, t = q * 32 + lane;          // block row t/32, column t%32 (as a PIXEL index for the irregular pass: 16 wide)
        const uint32_t o_base = smem_u32(smem + SM::OFF_O);
        uint32_t u = 0, oi = 0;   // oi: running index of the staging tile (alternates between the two buffers)
        int gi = 0, zeroed_b = -1;
        for (int g = blockIdx.x; g < ngroups; g += gridDim.x, ++gi) {
            int gxi, gyi, b;
            split(g, gxi, gyi, b);
            const int gx0 = gxi * GW, gy0 = gyi * GH;
            if (zero_flags != nullptr && b != zeroed_b) {      // first adds into this sample: its zero-fill must be complete
                if (lane == 0) {
                    unsigned int seen;
                    do {
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(zero_flags + b) : "memory");
                        if (seen < gridDim.x) __nanosleep(128);
                    } while (seen < gridDim.x);
                }
                __syncwarp();
                asm volatile("fence.proxy.async;" ::: "memory");   // the zeros were written through the generic proxy, the reduce-adds go through the async one
                zeroed_b = b;
            }
            // ---- irregular pixels of this group (thread <-> pixel t): literal scatter, warp-cooperative
            if (!(knobs & 512)) {
                const int px = gx0 + (t & 15), py = gy0 + (t >> 4);
                const bool valid = px < W && py < H;
                bool regular = true;
                float fx = 0.f, fy = 0.f;
                if (valid) {
                    const long long pofs = (long long)py * W + px;
                    fx = flow[(long long)b * 2 * hw + pofs];
                    fy = flow[(long long)b * 2 * hw + hw + pofs];
                    AxisTap<float> tx[K], ty[K];
                    regular = taps_regular<K>(fx, fy, px, py, Hs, Ws, tx, ty);
                }
                unsigned todo = __ballot_sync(0xffffffffu, valid && !regular);
                while (todo) {
                    const int sl = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const int qx = __shfl_sync(0xffffffffu, px, sl), qy = __shfl_sync(0xffffffffu, py, sl);
                    const float qfx = __shfl_sync(0xffffffffu, fx, sl), qfy = __shfl_sync(0xffffffffu, fy, sl);
                    const long long qofs = (long long)qy * W + qx;
                    float p[KK];
                    pixel_softmax_f32<KK>(logits + (long long)b * KK * hw + qofs, hw, p);
                    const __nv_bfloat162* go2 = reinterpret_cast<const __nv_bfloat162*>(gout + ((long long)b * hw + qofs) * C);
                    __nv_bfloat16* gs = gsrc + (long long)b * Hs * Ws * C;
                    for (int c2 = lane; c2 < CN / 2; c2 += 32) {
                        const float2 gv = __bfloat1622float2(go2[c2]);
                        const float g0 = gv.x * (1.0f / static_cast<float>(KK)), g1 = gv.y * (1.0f / static_cast<float>(KK));
#pragma unroll 1
                        for (int i = 0; i < K; ++i) {      // rolled on purpose (rare path): p[] is a small local array here
                            const AxisTap<float> ty = axis_tap<float>(qfy, i - K / 2, qy, Hs);
#pragma unroll 1
                            for (int j = 0; j < K; ++j) {
                                const AxisTap<float> tx = axis_tap<float>(qfx, j - K / 2, qx, Ws);
                                const float pij = p[i * K + j];
                                const float w4[4] = {tx.wlo * ty.wlo, tx.whi * ty.wlo, tx.wlo * ty.whi, tx.whi * ty.whi};
                                const long long o4[4] = {(long long)ty.lo * Ws + tx.lo, (long long)ty.lo * Ws + tx.hi,
                                                         (long long)ty.hi * Ws + tx.lo, (long long)ty.hi * Ws + tx.hi};
#pragma unroll
                                for (int q4 = 0; q4 < 4; ++q4) {
                                    const __nv_bfloat162 v2 = __floats2bfloat162_rn(g0 * pij * w4[q4], g1 * pij * w4[q4]);
                                    fb_red_add_bf16x2(gs + o4[q4] * C + 2 * c2, *reinterpret_cast<const uint32_t*>(&v2));
                                }
                            }
                        }
                    }
                }
            }
            mbar_wait(&info_full[gi % FB_NINFO], (gi / FB_NINFO) & 1, 0x000500 | 0x40000, gi);
            const FbInfo inf = infos[gi % FB_NINFO];
            const int nb4 = (inf.nrows + FB_GROWS - 1) / FB_GROWS;
            for (int cb = 0; cb < inf.ncb; ++cb)
                for (int rb = 0; rb < nb4; ++rb) {
                    // reduce-add box of this block: [bw positions][brows rows]; staging rows are packed in the same order
                    const int bw = (cb == inf.ncb - 1) ? inf.wlast : FB_BW;
                    const int brows = (rb == nb4 - 1 && inf.nrows - FB_GROWS * rb <= 2) ? 2 : FB_GROWS;
                    const CUtensorMap* rmap = &tmaps_gs.m[((bw - 24) >> 2) * 2 + (brows >> 2)];
                    const int lin = (t >> 5) * bw + (t & 31);                 // this thread's row of the staging tile
                    const bool in_box = (t & 31) < bw && (t >> 5) < brows;
#pragma unroll 1
                    for (int hf = 0; hf < NH; ++hf, ++u) {
                        const int buf = u & 1;
                        mbar_wait(&gs_full[buf], (u >> 1) & 1, 0x000300 | 0x40000 | buf, u);
                        tc_fence_after();
                        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + SM::GS_COL0 + buf * HN;
#pragma unroll 1
                        for (int cg = 0; cg < HN / 64; ++cg, ++oi) {
                            uint32_t v0[32], v1[32];
                            if (!(knobs & 512)) {
                            tmem_ld_32x32(taddr + cg * 64, v0);
                            tmem_ld_32x32(taddr + cg * 64 + 32, v1);
                            tmem_ld_wait();
                            }
                            if (cg == HN / 64 - 1) {  // accumulator half fully read: hand it back to the MMA warp
                                tc_fence_before();
                                __syncwarp();
                                if (lane == 0) mbar_arrive(&gs_empty[buf]);
                            }
                            if (knobs & 4) continue;
                            const uint32_t ob = o_base + (oi & 1) * SM::O_BUF + lin * 128;
                            fb_named_bar_sync(1, 128);        // staging buffer (oi & 1) is free (issuer waited on its reader)
                            if (in_box) {
#pragma unroll
                            for (int ch = 0; ch < 8; ++ch) {  // 8 x 16 bytes = 64 channels, 128B swizzle (chunk ^= row & 7)
                                const uint32_t* v = ch < 4 ? v0 + 8 * ch : v1 + 8 * (ch - 4);
                                uint32_t pk[4];
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
                                    pk[i] = *reinterpret_cast<const uint32_t*>(&h2);
                                }
                                sts128(ob + ((ch ^ (lin & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
                            }
                            }
                            fence_proxy_async_smem();
                            fb_named_bar_sync(2, 128);        // tile complete
                            if (warp == 12 && elect_one()) {
                                tma_reduce_add_4d(rmap, o_base + (oi & 1) * SM::O_BUF, hf * HN + cg * 64, inf.x0 + cb * FB_BW,
                                                  inf.y0 + rb * FB_GROWS, b);
                                bulk_commit();
                                bulk_wait_read<1>();          // the OTHER buffer's reduce has finished reading smem
                            }
                        }
                    }
                }
        }
        if (warp == 12 && elect_one()) bulk_wait<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
    tc_profile_total(t_start);
}

template <int K, int CN>
static int launch_fused(const void* src, const void* flow, const void* logits, const void* gout, void* gsrc, void* gflow,
                        void* glogits, int B, int C, int Hs, int Ws, int H, int W, int accumulate, void* workspace,
                        long long workspace_bytes, cudaStream_t st_) {
    static const PFN_tmapEncodeTiled enc = tmap_encoder();
    if (enc == nullptr) return GFLA_E_NOTSUP;
    CUtensorMap tg, ts;
    FbReduceMaps tgs;
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const cuuint64_t odim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t ostr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    const cuuint64_t sdim[4] = {(cuuint64_t)C, (cuuint64_t)Ws, (cuuint64_t)Hs, (cuuint64_t)B};
    const cuuint64_t sstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)Ws * C * 2, (cuuint64_t)Hs * Ws * C * 2};
    const cuuint32_t gbox[4] = {64, GW, GH, 1};                 // grad_out tile: 16 x 8 pixels
    const cuuint32_t sbox[4] = {64, FB_BW, FB_QROWS, 1};        // the source row of one Q stage

    // all three maps exist before anything is written (a failure here leaves the caller's buffers untouched)
    if (enc(&tg, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(gout), odim, ostr, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
        enc(&ts, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(src), sdim, sstr, sbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return GFLA_E_NOTSUP;
    for (int wi = 0; wi < 3; ++wi)
        for (int ri = 0; ri < 2; ++ri) {     // reduce-add tiles: (24 | 28 | 32) x (2 | 4) source positions
            const cuuint32_t rbox[4] = {64, (cuuint32_t)(24 + 4 * wi), (cuuint32_t)(ri == 0 ? 2 : FB_GROWS), 1};
            if (enc(&tgs.m[wi * 2 + ri], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, gsrc, sdim, sstr, rbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return GFLA_E_NOTSUP;
        }
    auto kern = k_local_attn_bwd_fused<K, CN>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemFB<CN>::ALLOC);
    if (e != cudaSuccess) return static_cast<int>(e);
    // the reduce-adds need a zero-filled grad_source (nothing was written before this point): with a workspace (gfla_local_attn_bwd_workspace_bytes)
    // the kernel zero-fills it itself, just ahead of its own adds (no separate pass, the zeros never travel to HBM and back);
    // without one, a memset in front of the launch
    unsigned int* zero_flags = nullptr;
    if (!accumulate) {
        const long long need = 4LL * B + 4LL * sm_count();     // per-sample counters + one progress word per CTA
        const bool in_kernel = workspace != nullptr && workspace_bytes >= need && aligned(workspace, 4) && tune_knob("GFLA_BWD_ZERO_IN_KERNEL", 1) != 0;
        const int z = in_kernel ? zero_async(workspace, (size_t)need, st_) : zero_async(gsrc, (size_t)B * C * Hs * Ws * 2, st_);
        if (z != GFLA_OK) return z;
        if (in_kernel) zero_flags = static_cast<unsigned int*>(workspace);
    }
    const int ngroups = B * ((H + GH - 1) / GH) * ((W + GW - 1) / GW);
    // With the in-kernel zero fill the CTAs depend on each other (a CTA's first add into a sample waits for every CTA's slice of
    // that sample's zeros): the grid must be co-resident, which only a COOPERATIVE launch guarantees -- a plain launch that shares the
    // GPU with another grid (a second stream's backward, an overlapping NCCL all-reduce under DDP) may start with part of its CTAs and
    // leave the rest waiting for SMs held by a grid in the same position.  One CTA per SM fits by construction (grid <= SM count).
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)min(ngroups, sm_count()));
    cfg.blockDim = dim3(FB_THREADS);
    cfg.dynamicSmemBytes = SmemFB<CN>::ALLOC;
    cfg.stream = st_;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    const int knobs_env = (int)tune_knob("GFLA_BWD_KNOBS", 0);
    auto launch = [&](unsigned int* zf, bool coop) -> cudaError_t {
        cfg.numAttrs = coop ? 1 : 0;
        return cudaLaunchKernelEx(&cfg, kern, tg, ts, tgs, (const __nv_bfloat16*)src, (const float*)flow, (const __nv_bfloat16*)logits,
                                  (const __nv_bfloat16*)gout, (__nv_bfloat16*)gsrc, (float*)gflow, (__nv_bfloat16*)glogits, B, C, Hs, Ws, H, W,
                                  accumulate, knobs_env, zf);
    };
    // GFLA_BWD_COOP: 1 (default) cooperative; 0 plain launch of the dependent grid (A/B timing on an otherwise idle GPU only);
    // 2 behave as if the cooperative launch had been refused (exercises the path below)
    const int coop_mode = zero_flags != nullptr ? (int)tune_knob("GFLA_BWD_COOP", 1) : 0;
    e = coop_mode == 2 ? cudaErrorCooperativeLaunchTooLarge : launch(zero_flags, coop_mode == 1);
    if (zero_flags != nullptr && (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported)) {
        // The device cannot hold the whole grid at once (an MPS client with a reduced SM share, a partitioned GPU) or has no cooperative
        // launch: then the CTAs must not depend on each other -- zero grad_source with a memset and run the same kernel without the
        // in-kernel zero fill (independent CTAs, any number resident).
        cudaGetLastError();
        const int z = zero_async(gsrc, (size_t)B * C * Hs * Ws * 2, st_);
        if (z != GFLA_OK) return z;
        e = launch(nullptr, false);
    }
    if (e != cudaSuccess) { cudaGetLastError(); return static_cast<int>(e); }
    return launch_status();
}

}  // namespace tc

int tc_wait_profile_bwd_fused(int enable, unsigned long long* out64) { return tc::tc_wait_profile(enable, out64); }

bool local_attn_bwd_fused_supported(int C, int k, const void* src) {
    return (C == 64 || C == 128 || C == 256) && (k == 3 || k == 5) && aligned(src, 16);
}

// accumulate = 0: grad_source is zero-filled here (after the tensor maps exist, i.e. after the last point of failure
// other than the launch itself) and all three gradients are overwritten; 1: everything is added into the caller's buffers.
int local_attn_bwd_fused_tc(const void* src, const void* flow, const void* logits, const void* gout, void* gsrc, void* gflow,
                            void* glogits, int B, int C, int Hs, int Ws, int H, int W, int k, int accumulate, void* workspace,
                            long long workspace_bytes, cudaStream_t st_) {
#define GFLA_FB_CASE(K_, CN_) \
    if (k == K_ && C == CN_) return tc::launch_fused<K_, CN_>(src, flow, logits, gout, gsrc, gflow, glogits, B, C, Hs, Ws, H, W, accumulate, workspace, workspace_bytes, st_);
    GFLA_FB_CASE(5, 256) GFLA_FB_CASE(5, 128) GFLA_FB_CASE(5, 64)
    GFLA_FB_CASE(3, 256) GFLA_FB_CASE(3, 128) GFLA_FB_CASE(3, 64)
#undef GFLA_FB_CASE
    return GFLA_E_NOTSUP;
}

}  // namespace gfla
