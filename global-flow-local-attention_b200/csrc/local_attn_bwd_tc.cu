// Fused local-attention BACKWARD, part 1: grad_source on the tensor cores (channels-last, bf16).
//
//   grad_source[b, t, c] += sum_p  Wfull[p, t] * grad_out[b, p, c]
//
// Wfull is the same sparse (pixels x source positions) weight matrix as in the forward tile kernel
// (local_attn_tc.cu): softmax probability x bilinear tap weight / k^2, with the reference's tap / clamp
// arithmetic (block_extractor_kernel.cu:127-161 defines the scatter this replaces).  The reference -- and
// our CUDA-core kernel -- issue one atomicAdd per (pixel, channel, tap corner): 4*k*k*C*H*W*B = 2.7e10
// scalar atomics at cfg2.  Here the scatter is a GEMM per pixel group:
//
//   * K = the 128 pixels of a 16x8 group;  A = Wfull^T, i.e. the [128 pixels][32 positions] weight slabs of
//     the forward kernel read as an MN-major operand (M = source positions: 4 row segments of 32 = 128);
//   * B = the grad_out tile [128 pixels][CN channels], one TMA box per 64 channels (128-byte rows,
//     MN-major = channel-contiguous);
//   * D[128 positions][CN channels] accumulates in TMEM (fp32), double-buffered;
//   * the epilogue converts D to bf16 into a swizzled staging tile and hands it to the TMA unit as a
//     REDUCE-ADD box store (cp.reduce.async.bulk.tensor .add): 128-byte vector atomics in L2, clipped
//     at the image border by the hardware -- ~(footprint / 128 pixels) ~ 6 box-adds per group and channel
//     block instead of 100 scalar atomics per (pixel, channel).
//
// grad_flow / grad_logits need the per-pixel dot products Q = sum_c grad_out * source (local_attn.cu) and
// are produced by a separate kernel; this one does not read `source` at all.
//
// Warp roles (10 warps, persistent CTA, static round-robin over pixel groups):
//   warp 0      producer: tap bounding box of the group from the flow, TMA load of the grad_out tile;
//   warp 1      MMA issuer;
//   warps 2-5   builders (one thread per pixel): softmax, taps, collapsed window, weight slabs per block;
//   warps 6-9   epilogue (one thread per source position of the block): TMEM -> bf16 -> staging -> TMA
//               reduce-add; irregular pixels (non-consecutive taps) are scattered here with scalar atomics.
#include "tile_window.cuh"

namespace gfla {
namespace tc {

constexpr int GS_ROWS = 4;        // source rows per block: M = 4 segments x 32 positions = 128 (blocks of 32 x 4 waste
                                  // far less of the footprint than 16 x 8 ones: rows are needed in multiples of 4, not 8)
constexpr int GS_BW = 32;         // positions per row segment
constexpr int GS_SLAB = 128 * GS_BW * 2;   // [128 pixels][32 positions] bf16, 64-byte rows, 64B swizzle
constexpr int GS_NA = 2;          // weight-slab stages
constexpr int GS_NINFO = 8;
constexpr int GS_NTHREADS = 320;

template <int CN>
struct SmemGS {
    static constexpr int G_CG = 128 * 128;                     // [128 pixels][64 channels] bf16
    static constexpr int G_BYTES = (CN / 64) * G_CG;
    static constexpr int A_STAGE = GS_ROWS * GS_SLAB;          // 4 slabs of [128 pixels][32 positions]
    static constexpr int O_BUF = 128 * 128;                    // staging: [128 positions][64 channels] bf16
    static constexpr int OFF_G = 0;
    static constexpr int OFF_A = OFF_G + G_BYTES;
    static constexpr int OFF_O = OFF_A + GS_NA * A_STAGE;
    static constexpr int OFF_W = OFF_O + 2 * O_BUF;
    static constexpr int OFF_INFO = OFF_W + 36 * 128 * 2;
    static constexpr int OFF_BAR = OFF_INFO + GS_NINFO * 16;
    static constexpr int NBAR = 2 + 2 * GS_NA + 4 + GS_NINFO;
    static constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
    static constexpr int ALLOC = OFF_TMEM + 16 + 1024;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void red_add_bf16x2(void* gptr, uint32_t v) {
    asm volatile("red.global.add.noftz.bf16x2 [%0], %1;" ::"l"(gptr), "r"(v) : "memory");
}
template <int K, int CN>
__global__ void __launch_bounds__(GS_NTHREADS, 1)
k_local_attn_bwd_gs_tc(const __grid_constant__ CUtensorMap tmap_g, const __grid_constant__ CUtensorMap tmap_gs,
                       const float* __restrict__ flow, const __nv_bfloat16* __restrict__ logits,
                       const __nv_bfloat16* __restrict__ gout, __nv_bfloat16* __restrict__ gsrc, int B, int C, int Hs,
                       int Ws, int H, int W) {
    using SM = SmemGS<CN>;
    constexpr int K1 = K + 1, KK = K * K;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::OFF_BAR);
    uint64_t* g_full = bars;                        // grad_out tile landed
    uint64_t* g_empty = bars + 1;                   // all MMAs of the group retired
    uint64_t* a_full = bars + 2;                    // [GS_NA] 128 builder arrivals
    uint64_t* a_empty = bars + 2 + GS_NA;           // [GS_NA]
    uint64_t* acc_full = bars + 2 + 2 * GS_NA;      // [2]
    uint64_t* acc_empty = acc_full + 2;             // [2]
    uint64_t* info_full = acc_empty + 2;            // [GS_NINFO]
    GroupInfo* infos = reinterpret_cast<GroupInfo*>(smem + SM::OFF_INFO);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM::OFF_TMEM);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gxn = (W + GW - 1) / GW, gyn = (H + GH - 1) / GH;
    const int ngroups = B * gyn * gxn;
    const int c0 = blockIdx.y * CN;
    const long long hw = (long long)H * W;

    if (threadIdx.x == 0) {
        mbar_init(g_full, 1);
        mbar_init(g_empty, 1);
        for (int i = 0; i < GS_NA; ++i) { mbar_init(&a_full[i], 128); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
        for (int i = 0; i < GS_NINFO; ++i) mbar_init(&info_full[i], 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmap_g);
        tma_prefetch_desc(&tmap_gs);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * CN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================================================= producer
        int gi = 0;
        for (int g = blockIdx.x; g < ngroups; g += gridDim.x, ++gi) {
            const int gx0 = (g % gxn) * GW, gy0 = ((g / gxn) % gyn) * GH, b = g / (gxn * gyn);
            int x0, y0, x1, y1;
            group_bbox<K>(flow, b, gx0, gy0, H, W, Hs, Ws, lane, false, x0, y0, x1, y1);
            if (lane == 0) {
                infos[gi % GS_NINFO] = GroupInfo{x0, y0, (x1 - x0 + GS_BW) / GS_BW, (y1 - y0 + GS_ROWS) / GS_ROWS};
                mbar_arrive(&info_full[gi % GS_NINFO]);
            }
            mbar_wait(g_empty, (gi & 1) ^ 1, 0x000600, gi);
            if (elect_one()) {
                mbar_arrive_expect_tx(g_full, SM::G_BYTES);
#pragma unroll
                for (int cg = 0; cg < CN / 64; ++cg)
                    tma_load_4d(smem + SM::OFF_G + cg * SM::G_CG, &tmap_g, g_full, c0 + cg * 64, gx0, gy0, b);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer
        constexpr uint32_t idesc = make_idesc_f16(128, CN, true, true, true);  // A and B both MN-major
        uint32_t blk = 0;
        int gi = 0;
        for (int g = blockIdx.x; g < ngroups; g += gridDim.x, ++gi) {
            mbar_wait(&info_full[gi % GS_NINFO], (gi / GS_NINFO) & 1, 0x010500, gi);
            const GroupInfo inf = infos[gi % GS_NINFO];
            const int nblk = inf.ncb * inf.nrc;
            mbar_wait(g_full, gi & 1, 0x010700, gi);
            for (int bi = 0; bi < nblk; ++bi, ++blk) {
                const int st = blk % GS_NA, buf = blk & 1;
                mbar_wait(&a_full[st], (blk / GS_NA) & 1, 0x010100 | st, blk);
                mbar_wait(&acc_empty[buf], ((blk >> 1) & 1) ^ 1, 0x010400 | buf, blk);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a0 = smem_u32(smem + SM::OFF_A + st * SM::A_STAGE);
                    const uint32_t b0 = smem_u32(smem + SM::OFF_G);
                    const uint32_t d_tmem = tmem_base + buf * CN;
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {  // 16 pixels per MMA
                        // A^T: M = positions (32 per slab, LBO = next slab), K = pixels (8 per 512-byte atom)
                        const uint64_t ad = make_smem_desc(a0 + ks * 1024, GS_SLAB, 512, kSwizzle64);
                        // B: N = channels (64 per 128-byte row, LBO = next channel group), K = pixels (8 per 1 KB atom)
                        const uint64_t bd = make_smem_desc(b0 + ks * 2048, SM::G_CG, 1024, kSwizzle128);
                        umma_f16(d_tmem, ad, bd, idesc, ks != 0 ? 1u : 0u);
                    }
                    tc_commit(&a_empty[st]);
                    tc_commit(&acc_full[buf]);
                    if (bi == nblk - 1) tc_commit(g_empty);
                }
                __syncwarp();
            }
        }
    } else if (warp < 6) {
        // ================================================================= builders
        const int q = warp & 3, m = q * 32 + lane;
        const float inv_kk = 1.0f / static_cast<float>(KK);
        const uint32_t wsm_a = smem_u32(smem + SM::OFF_W) + m * 4;
        const uint32_t a_base = smem_u32(smem + SM::OFF_A) + m * (GS_BW * 2);
        const uint32_t swz = ((m >> 1) & 3) << 4;   // 64B swizzle: 16B chunk ^= bits 1-2 of the row
        uint32_t blk = 0, dirty = 0xffffffffu;
        int gi = 0;
        for (int g = blockIdx.x; g < ngroups; g += gridDim.x, ++gi) {
            const int gx0 = (g % gxn) * GW, gy0 = ((g / gxn) % gyn) * GH, b = g / (gxn * gyn);
            const int px = gx0 + (m & 15), py = gy0 + (m >> 4);
            const bool valid = px < W && py < H;
            int X0 = 0, Y0 = 0;
            bool live = false;
            if (valid) {
                const long long pofs = (long long)py * W + px;
                float p[KK];
                pixel_softmax_f32<KK>(logits + (long long)b * KK * hw + pofs, hw, p);
                const float fx = flow[(long long)b * 2 * hw + pofs], fy = flow[(long long)b * 2 * hw + hw + pofs];
                AxisTap<float> tx[K], ty[K];
                live = taps_regular<K>(fx, fy, px, py, Hs, Ws, tx, ty);
                if (live) {
                    float w[K1 * K1];
                    build_window<K>(p, tx, ty, Hs, Ws, inv_kk, w, X0, Y0);
                    store_window_words<K>(wsm_a, w);
                }
            }
            mbar_wait(&info_full[gi % GS_NINFO], (gi / GS_NINFO) & 1, 0x020500, gi);
            const GroupInfo inf = infos[gi % GS_NINFO];
            for (int cb = 0; cb < inf.ncb; ++cb) {
                const int e0 = X0 - (inf.x0 + cb * GS_BW);
                const bool cols_hit = live && e0 > -K1 && e0 < GS_BW;
                for (int rb = 0; rb < inf.nrc; ++rb, ++blk) {
                    const int st = blk % GS_NA;
                    mbar_wait(&a_empty[st], ((blk / GS_NA) & 1) ^ 1, 0x020200 | st, blk);
                    const uint32_t a_stage = a_base + st * SM::A_STAGE;
                    const int R0 = inf.y0 + rb * GS_ROWS;
                    bool wrote = false;
#pragma unroll
                    for (int seg = 0; seg < GS_ROWS; ++seg)
                        wrote |= fill_slab_row<K, GS_BW>(a_stage + seg * GS_SLAB, swz, wsm_a, cols_hit, (R0 + seg) - Y0, e0, dirty,
                                                  1u << (st * GS_ROWS + seg));
                    if (wrote) fence_proxy_async_smem();
                    mbar_arrive(&a_full[st]);
                }
            }
        }
    } else {
        // ================================================================= epilogue
        const int q = warp & 3, t = q * 32 + lane;          // source position of the block: row t/32, column t%32
                                                            // (as a PIXEL index for the irregular-tap pass below: 16 wide)
        const uint32_t o_base = smem_u32(smem + SM::OFF_O);
        const bool issuer = (warp == 6 && lane == 0);
        uint32_t blk = 0, oi = 0;   // oi: running index of the staging tile (alternates between the two buffers)
        int gi = 0;
        for (int g = blockIdx.x; g < ngroups; g += gridDim.x, ++gi) {
            const int gx0 = (g % gxn) * GW, gy0 = ((g / gxn) % gyn) * GH, b = g / (gxn * gyn);
            // ---- irregular pixels of this group (thread <-> pixel t): literal scalar scatter, warp-cooperative
            {
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
                    // fire-and-forget vector reductions (red.global.add.noftz.bf16x2), lanes <-> channel pairs:
                    // scalar bf16 atomics with a return value made one such pixel cost ~1 ms (a straggler CTA)
                    const __nv_bfloat162* go2 = reinterpret_cast<const __nv_bfloat162*>(gout + ((long long)b * hw + qofs) * C + c0);
                    __nv_bfloat16* gs = gsrc + (long long)b * Hs * Ws * C + c0;
                    for (int c2 = lane; c2 < CN / 2; c2 += 32) {
                        const float2 gv = __bfloat1622float2(go2[c2]);
                        const float g0 = gv.x * (1.0f / static_cast<float>(KK)), g1 = gv.y * (1.0f / static_cast<float>(KK));
                        for (int i = 0; i < K; ++i) {
                            const AxisTap<float> ty = axis_tap<float>(qfy, i - K / 2, qy, Hs);
                            for (int j = 0; j < K; ++j) {
                                const AxisTap<float> tx = axis_tap<float>(qfx, j - K / 2, qx, Ws);
                                const float pij = p[i * K + j];
                                const float w4[4] = {tx.wlo * ty.wlo, tx.whi * ty.wlo, tx.wlo * ty.whi, tx.whi * ty.whi};
                                const long long o4[4] = {(long long)ty.lo * Ws + tx.lo, (long long)ty.lo * Ws + tx.hi,
                                                         (long long)ty.hi * Ws + tx.lo, (long long)ty.hi * Ws + tx.hi};
#pragma unroll
                                for (int q4 = 0; q4 < 4; ++q4) {
                                    const __nv_bfloat162 v2 = __floats2bfloat162_rn(g0 * pij * w4[q4], g1 * pij * w4[q4]);
                                    red_add_bf16x2(gs + o4[q4] * C + 2 * c2, *reinterpret_cast<const uint32_t*>(&v2));
                                }
                            }
                        }
                    }
                }
            }
            mbar_wait(&info_full[gi % GS_NINFO], (gi / GS_NINFO) & 1, 0x030500, gi);
            const GroupInfo inf = infos[gi % GS_NINFO];
            for (int cb = 0; cb < inf.ncb; ++cb)
                for (int rb = 0; rb < inf.nrc; ++rb, ++blk) {
                    const int buf = blk & 1;
                    mbar_wait(&acc_full[buf], (blk >> 1) & 1, 0x030300 | buf, blk);
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * CN;
#pragma unroll 1
                    for (int cg = 0; cg < CN / 64; ++cg, ++oi) {
                        uint32_t v0[32], v1[32];
                        tmem_ld_32x32(taddr + cg * 64, v0);
                        tmem_ld_32x32(taddr + cg * 64 + 32, v1);
                        tmem_ld_wait();
                        if (cg == CN / 64 - 1) {  // accumulator fully read: hand it back to the MMA warp
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&acc_empty[buf]);
                        }
                        const uint32_t ob = o_base + (oi & 1) * SM::O_BUF + t * 128;
                        named_bar_sync(1, 128);           // staging buffer (oi & 1) is free (issuer waited on its reader)
#pragma unroll
                        for (int ch = 0; ch < 8; ++ch) {  // 8 x 16 bytes = 64 channels, 128B swizzle (chunk ^= row & 7)
                            const uint32_t* v = ch < 4 ? v0 + 8 * ch : v1 + 8 * (ch - 4);
                            uint32_t pk[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
                                pk[i] = *reinterpret_cast<const uint32_t*>(&h2);
                            }
                            sts128(ob + ((ch ^ (t & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
                        }
                        fence_proxy_async_smem();
                        named_bar_sync(2, 128);           // tile complete
                        if (issuer) {
                            tma_reduce_add_4d(&tmap_gs, o_base + (oi & 1) * SM::O_BUF, c0 + cg * 64, inf.x0 + cb * GS_BW,
                                              inf.y0 + rb * GS_ROWS, b);
                            bulk_commit();
                            bulk_wait_read<1>();          // the OTHER buffer's reduce has finished reading smem
                        }
                    }
                }
        }
        if (issuer) bulk_wait<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 2 * CN);
}

template <int K, int CN>
static int launch_gs(const void* flow, const void* logits, const void* gout, void* gsrc, int B, int C, int Hs, int Ws,
                     int H, int W, cudaStream_t st_) {
    static const PFN_tmapEncodeTiled enc = tmap_encoder();
    if (enc == nullptr) return GFLA_E_NOTSUP;
    CUtensorMap tg, tgs;
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const cuuint32_t box[4] = {64, GW, GH, 1};             // grad_out tile: 16 x 8 pixels
    const cuuint32_t rbox[4] = {64, GS_BW, GS_ROWS, 1};    // reduce-add tile: 32 x 4 source positions
    {
        const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        const cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
        if (enc(&tg, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(gout), gdim, gstr, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return GFLA_E_NOTSUP;
    }
    {
        const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)Ws, (cuuint64_t)Hs, (cuuint64_t)B};
        const cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)Ws * C * 2, (cuuint64_t)Hs * Ws * C * 2};
        if (enc(&tgs, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, gsrc, gdim, gstr, rbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) !=
            CUDA_SUCCESS)
            return GFLA_E_NOTSUP;
    }
    auto kern = k_local_attn_bwd_gs_tc<K, CN>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemGS<CN>::ALLOC);
    if (e != cudaSuccess) return static_cast<int>(e);
    const int ngroups = B * ((H + GH - 1) / GH) * ((W + GW - 1) / GW);
    dim3 grid((unsigned)min(ngroups, sm_count()), (unsigned)(C / CN));
    kern<<<grid, GS_NTHREADS, SmemGS<CN>::ALLOC, st_>>>(tg, tgs, (const float*)flow, (const __nv_bfloat16*)logits,
                                                        (const __nv_bfloat16*)gout, (__nv_bfloat16*)gsrc, B, C, Hs, Ws, H, W);
    return launch_status();
}

}  // namespace tc

int tc_debug_set_buffer_bwd(void* host_mapped) {
    unsigned long long* p = static_cast<unsigned long long*>(host_mapped);
    return static_cast<int>(cudaMemcpyToSymbol(tc::g_tc_dbg, &p, sizeof(p)));
}

static int pick_cn_bwd(int C) {
    if (C % 256 == 0) return 256;
    if (C == 128 || C == 64) return C;
    return 0;
}

bool local_attn_bwd_tc_supported(int C, int k, int dtype, int flow_dtype, int layout, const void* gout, const void* gsrc) {
    return dtype == GFLA_BF16 && flow_dtype == GFLA_F32 && layout == GFLA_NHWC && (k == 3 || k == 5) &&
           pick_cn_bwd(C) != 0 && aligned(gout, 16) && aligned(gsrc, 16);
}

// grad_source += Wfull^T * grad_out  (channels-last bf16); the caller has zero-filled grad_source if it must not accumulate
int local_attn_bwd_gs_tc(const void* flow, const void* logits, const void* gout, void* gsrc, int B, int C, int Hs, int Ws,
                         int H, int W, int k, cudaStream_t st_) {
    const int cn = pick_cn_bwd(C);
#define GFLA_GS_CASE(K_, CN_) \
    if (k == K_ && cn == CN_) return tc::launch_gs<K_, CN_>(flow, logits, gout, gsrc, B, C, Hs, Ws, H, W, st_);
    GFLA_GS_CASE(5, 256) GFLA_GS_CASE(5, 128) GFLA_GS_CASE(5, 64)
    GFLA_GS_CASE(3, 256) GFLA_GS_CASE(3, 128) GFLA_GS_CASE(3, 64)
#undef GFLA_GS_CASE
    return GFLA_E_NOTSUP;
}

}  // namespace gfla
