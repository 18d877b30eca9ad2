// Fused local-attention FORWARD, channels-last, "strip" schedule (tcgen05 + TMEM + TMA).
//
// Same GEMM embedding as local_attn_tc.cu (16x8 pixel tile = M, channels = N, footprint positions = K,
// A = per-pixel weight slabs built on the fly, B = source rows by TMA, D in TMEM), different traversal:
// local_attn_tc.cu treats every tile on its own and pulls its whole (k + flow variation)-row footprint
// through the L2->SM fabric, ~4.3 source positions per output pixel -- and that fabric, not HBM, is what
// bounds the kernel.  Here a CTA owns a vertical run of tiles; a source row chunk
// that the current tile loads and the next tile also needs is multiplied into BOTH accumulators while it
// sits in shared memory (two weight slabs, one B operand), so in steady state a tile fetches only the
// ~8 rows its predecessor did not already see.  The schedule is in strip_plan.h.
//
// Warp roles (4 warpgroups = 16 warps, persistent CTA, static round-robin over strips; setmaxnreg moves the control
// warpgroup's registers to the builders and the epilogue):
//   warp 0        producer: tile bounding boxes, step schedule, TMA row loads;
//   warp 1        MMA issuer: per step one MMA set per consuming tile (accumulators alternate between the
//                 two TMEM halves; the next tile's half is claimed at its first shared step);   (warps 2, 3 idle)
//   warps 4-7     builder team 0, warps 8-11 builder team 1 (thread = pixel): team T owns the tiles of parity T --
//                 softmax + tap arithmetic + window collapse once per tile, then that tile's weight slab (sub-slab T
//                 of the stage) in every step that feeds it: the shared steps of the previous tile's pass and the
//                 steps of its own pass.  While one team fills, the other builds its next window: the per-pixel
//                 work, which bounded the 4-builder-warp version, runs two tiles deep;
//   warps 12-15   epilogue: TMEM -> bf16 -> swizzled staging -> TMA tensor store (optional mask blend), literal
//                 4-tap recomputation of the irregular pixels (bit-identical indexing).
#include "strip_plan.h"
#include "tile_window.cuh"

namespace gfla {

namespace tc {

constexpr int ST_RCH = 2;        // source rows per step (fixed by strip_plan.h: chunk j = rows 2j, 2j+1)
constexpr int ST_FBW = 32;       // source positions per row segment
constexpr int ST_NINFO = 8;      // >= stages + 3: every tile pass has at least one step (strip_plan), so the
                                 // producer is never more than `stages` tiles ahead of the builders / MMA warp
constexpr int ST_THREADS = 512;
constexpr int ST_REG_CTRL = 72, ST_REG_BUILD = 152, ST_REG_EPI = 136;   // per-thread registers by warpgroup; sum over the 4 groups = 512
static_assert(ST_REG_CTRL + 2 * ST_REG_BUILD + ST_REG_EPI <= 512, "register budget");
template <int N> __device__ __forceinline__ void st_reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void st_reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

template <int CN>
struct StripSmem {
    static constexpr int NSTAGE = CN == 256 ? 3 : 4;
    static constexpr int S_SLAB = CN * ST_FBW * 2;             // one source row segment: [CN/64][32 x][64 ch] bf16
    static constexpr int FA_SLAB = 128 * ST_FBW * 2;           // weight slab: [128 pixels][32 positions] bf16
    static constexpr int S_STAGE = ST_RCH * S_SLAB;
    static constexpr int A_TILE = ST_RCH * FA_SLAB;            // the slabs of one consuming tile
    static constexpr int A_STAGE = 2 * A_TILE;                 // sub-slab 0: tiles of even parity (team 0), sub-slab 1: odd (team 1)
    static constexpr int W_TILE = 36 * 128 * 2;                // collapsed windows of one tile, [18 words][128 pixels]
    static constexpr int OFF_S = 0;
    static constexpr int OFF_A = OFF_S + NSTAGE * S_STAGE;
    static constexpr int OFF_W = OFF_A + NSTAGE * A_STAGE;
    static constexpr int O_WARP = 32 * 64;                     // output staging of one epilogue warp: [32 pixels][32 channels] bf16
    static constexpr int OFF_O = OFF_W + 2 * W_TILE;           // 64B-swizzled: 512-byte aligned
    static constexpr int OFF_INFO = OFF_O + 4 * O_WARP;
    static constexpr int OFF_BAR = OFF_INFO + ST_NINFO * 32;
    static constexpr int NBAR = 6 * NSTAGE + 4 + ST_NINFO;
    static constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
    static constexpr int TOTAL = OFF_TMEM + 16;
    static constexpr int ALLOC = TOTAL + 1024;                 // slack to align the base to 1024 B
};
static_assert(sizeof(StripTile) == 32, "StripTile is stored in 32-byte info slots");
static_assert(StripSmem<256>::ALLOC <= 232448, "shared memory budget");
static_assert(StripSmem<256>::OFF_O % 512 == 0 && StripSmem<128>::OFF_O % 512 == 0 && StripSmem<64>::OFF_O % 512 == 0, "staging alignment");

// strips: units of `ts` vertically adjacent tiles; unit -> (b, gx, first tile row, end tile row)
struct StripGeom {
    int gxn, gyn, nseg, ts, units;
    __device__ __forceinline__ void decode(int unit, int& b, int& gx, int& ty0, int& ty1) const {
        gx = unit % gxn;
        const int seg = (unit / gxn) % nseg;
        b = unit / (gxn * nseg);
        ty0 = seg * ts;
        ty1 = min(gyn, ty0 + ts);
    }
};

// the tiles of one CTA in processing order: strips blockIdx.x, blockIdx.x + gridDim.x, ..., top to bottom inside a strip
struct TileIt {
    int unit, b, gx, ty, ty1;
    __device__ __forceinline__ bool ok(const StripGeom& g) const { return unit < g.units; }
    __device__ __forceinline__ void enter(const StripGeom& g) {
        if (unit < g.units) {
            int ty0;
            g.decode(unit, b, gx, ty0, ty1);
            ty = ty0;
        }
    }
    __device__ __forceinline__ void start(const StripGeom& g, int first_unit) {
        unit = first_unit; b = 0; gx = 0; ty = 0; ty1 = 0;
        enter(g);
    }
    __device__ __forceinline__ void next(const StripGeom& g, int stride) {
        if (++ty >= ty1) {
            unit += stride;
            enter(g);
        }
    }
};

template <int K, int CN>
__global__ void __launch_bounds__(ST_THREADS, 1)
k_local_attn_fwd_strip(const __grid_constant__ CUtensorMap tmap_src, const __grid_constant__ CUtensorMap tmap_out,
                       const __nv_bfloat16* __restrict__ src,
                       const float* __restrict__ flow, const __nv_bfloat16* __restrict__ logits,
                       __nv_bfloat16* __restrict__ out, __nv_bfloat16* __restrict__ probs,
                       const __nv_bfloat16* __restrict__ prev, const __nv_bfloat16* __restrict__ mask, int B, int C, int Hs,
                       int Ws, int H, int W, int ts, int knobs_arg) {
    const int knobs = GFLA_KNOBS(knobs_arg);      // 0 in the shipped build: every `knobs & x` test folds away
    // `knobs` (environment GFLA_TC_KNOBS, default 0 = production) switch parts of the pipeline off for timing experiments
    // (results are wrong when set): bit 8 no output stores, bit 10 no weight scatter (slabs only zeroed), bit 11 no slab
    // writes at all, bit 12 no MMAs, bit 13 no TMA loads, bit 14 L2 prefetch of upcoming rows ON (measured slower: 0.54 vs 0.49 ms at cfg2; off by default).
    using SM = StripSmem<CN>;
    constexpr int NSTAGE = SM::NSTAGE, FBW = ST_FBW, RCH = ST_RCH;
    constexpr int K1 = K + 1, KK = K * K;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::OFF_BAR);
    uint64_t* full_s = bars;                      // [NSTAGE] TMA bytes landed
    uint64_t* full_a = bars + NSTAGE;             // [2][NSTAGE] 128 arrivals of builder team T: sub-slab T of the stage is written
    uint64_t* empty = bars + 3 * NSTAGE;          // [NSTAGE] MMAs of the stage retired: the source rows may be overwritten
    uint64_t* empty_a = bars + 4 * NSTAGE;        // [2][NSTAGE] MMAs that read sub-slab T of the stage retired (only steps that fed team T's tile)
    uint64_t* acc_full = bars + 6 * NSTAGE;       // [2]
    uint64_t* acc_empty = bars + 6 * NSTAGE + 2;  // [2]
    uint64_t* info_full = bars + 6 * NSTAGE + 4;  // [ST_NINFO]
    StripTile* infos = reinterpret_cast<StripTile*>(smem + SM::OFF_INFO);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM::OFF_TMEM);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t_start = tc_profile_clock();
    StripGeom geo;
    geo.gxn = (W + GW - 1) / GW;
    geo.gyn = (H + GH - 1) / GH;
    geo.ts = ts;
    geo.nseg = (geo.gyn + ts - 1) / ts;
    geo.units = B * geo.gxn * geo.nseg;
    const int c0 = blockIdx.y * CN;
    const long long hw = (long long)H * W;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; ++i) {
            mbar_init(&full_s[i], 1); mbar_init(&full_a[i], 128); mbar_init(&full_a[NSTAGE + i], 128); mbar_init(&empty[i], 1);
            mbar_init(&empty_a[i], 1); mbar_init(&empty_a[NSTAGE + i], 1);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
        for (int i = 0; i < ST_NINFO; ++i) mbar_init(&info_full[i], 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmap_src);
        tma_prefetch_desc(&tmap_out);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * CN >= 32 ? 2 * CN : 32);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    TileIt first;
    first.start(geo, blockIdx.x);
    const int ustride = gridDim.x;

    if (warp < 4) {
      st_reg_dec<ST_REG_CTRL>();
      if (warp == 0) {
        // ================================================================= producer
        // The flow of a tile is loaded a whole pass before its bounding box is needed (TileFlow registers in flight
        // across the TMA loop): a box costs ~4 us of load latency, which otherwise stalls the row loads once per tile.
        TileIt cur = first, ld = first, pend_it = first;
        TileFlow fr;
        TileBox box0 = TileBox{0, 0, 0, 0}, box1 = box0;
        int unit1 = -1;
        bool pend = false;
        if (ld.ok(geo)) {
            tile_flow_load(flow, ld.b, ld.gx * GW, ld.ty * GH, H, W, lane, fr);
            tile_bbox_reduce<K>(fr, ld.gx * GW, ld.ty * GH, H, W, Hs, Ws, lane, false, box0.x0, box0.y0, box0.x1, box0.y1);
            ld.next(geo, ustride);
        }
        if (ld.ok(geo)) {
            tile_flow_load(flow, ld.b, ld.gx * GW, ld.ty * GH, H, W, lane, fr);
            tile_bbox_reduce<K>(fr, ld.gx * GW, ld.ty * GH, H, W, Hs, Ws, lane, false, box1.x0, box1.y0, box1.x1, box1.y1);
            unit1 = ld.unit;
            ld.next(geo, ustride);
        }
        if (ld.ok(geo)) {
            tile_flow_load(flow, ld.b, ld.gx * GW, ld.ty * GH, H, W, lane, fr);
            pend_it = ld;
            pend = true;
            ld.next(geo, ustride);
        }
        uint32_t it = 0;  // global step counter
        int ti = 0;       // global tile counter of this CTA
        int k0 = 1, k1 = 0;
        for (; cur.ok(geo); cur.next(geo, ustride), ++ti) {
            const bool has_next = unit1 == cur.unit;   // the next tile of this CTA continues the same strip
            const StripTile t = strip_plan(box0, has_next, box1, k0, k1, FBW);
            if (lane == 0) {
                infos[ti % ST_NINFO] = t;
                mbar_arrive(&info_full[ti % ST_NINFO]);
            }
            for (int cb = 0; cb < t.ncb; ++cb)
                for (int j = t.j0; j <= t.j1; ++j) {
                    if (strip_skipped(t, j)) continue;
                    const int slot = it % NSTAGE;
                    mbar_wait(&empty[slot], ((it / NSTAGE) & 1) ^ 1, 0x000200 | slot, it);
                    if (lane == 0 && (knobs & 8192)) mbar_arrive(&full_s[slot]);
                    if (!(knobs & 8192) && elect_one()) {
                        mbar_arrive_expect_tx(&full_s[slot], SM::S_STAGE);
#pragma unroll
                        for (int rr = 0; rr < RCH; ++rr) {
                            uint8_t* dst = smem + SM::OFF_S + slot * SM::S_STAGE + rr * SM::S_SLAB;
#pragma unroll
                            for (int cg = 0; cg < CN / 64; ++cg)   // [CN/64 channel groups][FBW x][64 channels]
                                tma_load_4d(dst + cg * (FBW * 128), &tmap_src, &full_s[slot], c0 + cg * 64, t.xs + cb * FBW,
                                            2 * j + rr, cur.b);
                        }
                    }
                    __syncwarp();
                    ++it;
                }
            k0 = t.s0; k1 = t.s1;
            const long long tb0 = tc_profile_clock();
            box0 = box1;
            if (pend) {
                tile_bbox_reduce<K>(fr, pend_it.gx * GW, pend_it.ty * GH, H, W, Hs, Ws, lane, false, box1.x0, box1.y0, box1.x1, box1.y1);
                unit1 = pend_it.unit;
                // pull the source rows of the tile after next into L2 (two tile passes ahead of their TMA loads): with NSTAGE
                // stages in flight a row chunk that must come from HBM exposes the DRAM latency once per stage
                if (knobs & 16384) {
                    const int pj0 = box1.y0 >> 1, pj1 = box1.y1 >> 1, pcb = (box1.x1 - box1.x0 + FBW) / FBW;
                    const int per = (pj1 - pj0 + 1) * RCH * (CN / 64);
                    for (int i = lane; i < per * pcb; i += 32) {
                        const int cb = i / per, r = i - cb * per, cg = r % (CN / 64), row = 2 * pj0 + r / (CN / 64);
                        tma_prefetch_4d(&tmap_src, c0 + cg * 64, box1.x0 + cb * FBW, row, pend_it.b);
                    }
                }
            } else {
                unit1 = -1;
            }
            pend = ld.ok(geo);
            if (pend) {
                tile_flow_load(flow, ld.b, ld.gx * GW, ld.ty * GH, H, W, lane, fr);
                pend_it = ld;
                ld.next(geo, ustride);
            }
            tc_profile_add(0, 6, tc_profile_clock() - tb0);          // box of tile n+2, flow loads of tile n+3
        }
      } else if (warp == 1) {
        // ================================================================= MMA issuer
        constexpr uint32_t idesc = make_idesc_f16(128, CN, true, false, true);
        uint32_t it = 0;
        uint32_t par_a[2] = {0u, 0u};   // per builder team: phase parity of full_a[team][slot], one bit per slot
        int ti = 0;
        bool started = false;   // this tile's accumulator already holds the previous pass's shared steps
        for (TileIt cur = first; cur.ok(geo); cur.next(geo, ustride), ++ti) {
            mbar_wait(&info_full[ti % ST_NINFO], (ti / ST_NINFO) & 1, 0x010500, ti);
            const StripTile t = infos[ti % ST_NINFO];
            const int buf = ti & 1;
            bool next_started = false;
            for (int cb = 0; cb < t.ncb; ++cb)
                for (int j = t.j0; j <= t.j1; ++j) {
                    if (strip_skipped(t, j)) continue;
                    const bool shared = strip_shared(t, j);
                    if (!started) {   // first touch of this tile's TMEM half: its previous user must be drained
                        mbar_wait(&acc_empty[buf], ((ti >> 1) & 1) ^ 1, 0x010400 | buf, ti);
                        tc_fence_after();
                    }
                    if (shared && !next_started) {
                        mbar_wait(&acc_empty[buf ^ 1], (((ti + 1) >> 1) & 1) ^ 1, 0x010600 | (buf ^ 1), ti);
                        tc_fence_after();
                    }
                    const int slot = it % NSTAGE;
                    const uint32_t par = (it / NSTAGE) & 1;
                    mbar_wait(&full_s[slot], par, 0x010000 | slot, it);
                    mbar_wait(&full_a[buf * NSTAGE + slot], (par_a[buf] >> slot) & 1u, 0x010100 | slot, it);   // this tile's slab
                    par_a[buf] ^= 1u << slot;
                    if (shared) {                                                                              // the next tile's slab
                        mbar_wait(&full_a[(buf ^ 1) * NSTAGE + slot], (par_a[buf ^ 1] >> slot) & 1u, 0x010700 | slot, it);
                        par_a[buf ^ 1] ^= 1u << slot;
                    }
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t a0 = smem_u32(smem + SM::OFF_A + slot * SM::A_STAGE);
                        const uint32_t b0 = smem_u32(smem + SM::OFF_S + slot * SM::S_STAGE);
#pragma unroll
                        for (int sub = 0; sub < 2; ++sub) {
                            if ((sub == 1 && !shared) || (knobs & 4096)) break;
                            const int par_t = sub == 0 ? buf : (buf ^ 1);      // parity of the consuming tile = its TMEM half = its sub-slab
                            const uint32_t d_tmem = tmem_base + par_t * CN;
                            const bool fresh = sub == 0 ? !started : !next_started;
#pragma unroll
                            for (int rr = 0; rr < RCH; ++rr)
#pragma unroll
                                for (int h = 0; h < FBW / 16; ++h) {  // K = 16 positions per MMA
                                    // A = [128 px][32 pos], K-major, 64B rows, 64B swizzle; K-advance = +32 B
                                    const uint64_t ad = make_smem_desc(a0 + par_t * SM::A_TILE + rr * SM::FA_SLAB + h * 32, 16, 512, kSwizzle64);
                                    // B = [32 x][64 ch] per channel group, MN-major, 128B swizzle: LBO = next channel group,
                                    // SBO = next 8 positions (1 KB); K-advance = 2 KB
                                    const uint64_t bd = make_smem_desc(b0 + rr * SM::S_SLAB + h * 2048, FBW * 128, 1024, kSwizzle128);
                                    umma_f16(d_tmem, ad, bd, idesc, (fresh && rr == 0 && h == 0) ? 0u : 1u);
                                }
                        }
                        tc_commit(&empty[slot]);
                        tc_commit(&empty_a[buf * NSTAGE + slot]);
                        if (shared) tc_commit(&empty_a[(buf ^ 1) * NSTAGE + slot]);
                    }
                    __syncwarp();
                    started = true;
                    if (shared) next_started = true;
                    ++it;
                }
            if (elect_one()) tc_commit(&acc_full[buf]);
            __syncwarp();
            started = next_started;   // false after the last tile of a strip (it never shares)
        }
      }   // warps 2, 3: no role (they pad the control warpgroup so that setmaxnreg can hand their registers on)
    } else if (warp < 12) {
        // ================================================================= builders, team T = tiles of parity T
        st_reg_inc<ST_REG_BUILD>();
        const int T = (warp - 4) >> 2;
        const int q = warp & 3, m = q * 32 + lane;  // pixel index inside a tile
        const float inv_kk = 1.0f / static_cast<float>(KK);
        const uint32_t wsm_a = smem_u32(smem + SM::OFF_W) + T * SM::W_TILE + m * 4;                 // this team's window words
        const uint32_t a_base = smem_u32(smem + SM::OFF_A) + T * SM::A_TILE + m * (FBW * 2);        // this pixel's row in sub-slab T of stage 0
        const uint32_t swz = ((m >> 1) & 3) << 4;                               // 64B-swizzle XOR of this row's 16B chunks
        uint64_t* my_full = full_a + T * NSTAGE;
        uint64_t* my_empty = empty_a + T * NSTAGE;
        uint32_t fill_par = 0u;                 // per slot: parity of the number of fills this team has made into it
        uint32_t it = 0, dirty = 0xffffffffu;   // slab rows start with unknown contents: treat them as dirty
        int ti = 0;
        // Raw inputs of this thread's pixel of the team's NEXT tile: loaded a whole tile pass before the window is built
        // from them, so their latency hides behind the fills.
        __nv_bfloat16 lg[KK];
        float pfx = 0.f, pfy = 0.f;
        auto load_pixel = [&](const TileIt& tl) {
            const int px = tl.gx * GW + (m & 15), py = tl.ty * GH + (m >> 4);
            if (px < W && py < H) {
                const long long pofs = (long long)py * W + px;
                const __nv_bfloat16* lp = logits + (long long)tl.b * KK * hw + pofs;
#pragma unroll
                for (int t = 0; t < KK; ++t) lg[t] = lp[t * hw];
                pfx = flow[(long long)tl.b * 2 * hw + pofs];
                pfy = flow[(long long)tl.b * 2 * hw + hw + pofs];
            }
        };
        // collapsed window of that pixel -> shared memory; returns whether the pixel exists and is regular
        auto make_window = [&](const TileIt& tl, int& X0, int& Y0) -> bool {
            const int px = tl.gx * GW + (m & 15), py = tl.ty * GH + (m >> 4);
            X0 = 0; Y0 = 0;
            if (!(px < W && py < H)) return false;
            const long long pofs = (long long)py * W + px;
            float p[KK];
#pragma unroll
            for (int t = 0; t < KK; ++t) p[t] = __bfloat162float(lg[t]);
            softmax_inplace_f32<KK>(p);
            if (probs != nullptr && blockIdx.y == 0) {
                __nv_bfloat16* pr = probs + (long long)tl.b * KK * hw + pofs;
#pragma unroll
                for (int t = 0; t < KK; ++t) pr[t * hw] = __float2bfloat16_rn(p[t]);
            }
            AxisTap<float> tx[K], ty_[K];
            if (!taps_regular<K>(pfx, pfy, px, py, Hs, Ws, tx, ty_)) return false;
            float w[K1 * K1];
            build_window<K>(p, tx, ty_, Hs, Ws, inv_kk, w, X0, Y0);
            store_window_words<K>(wsm_a, w);
            return true;
        };
        // `mine` walks this team's tiles (parity T of the CTA's tile sequence), `ld` one own tile further (raw-input prefetch)
        TileIt mine = first, ld = first;
        if (T == 1 && mine.ok(geo)) mine.next(geo, ustride);
        ld = mine;
        int X0 = 0, Y0 = 0;
        bool live = false, have = false;          // window of tile `mine` is in shared memory / registers
        if (ld.ok(geo)) {
            load_pixel(ld);
            ld.next(geo, ustride);
            if (ld.ok(geo)) ld.next(geo, ustride);
        }
        if (T == 0 && mine.ok(geo)) {             // the CTA's first tile has no previous pass to hide its window behind
            live = make_window(mine, X0, Y0);
            have = true;
            if (ld.ok(geo)) {
                load_pixel(ld);
                ld.next(geo, ustride);
                if (ld.ok(geo)) ld.next(geo, ustride);
            }
        }
        for (TileIt cur = first; cur.ok(geo); cur.next(geo, ustride), ++ti) {
            const bool own = (ti & 1) == T;
            if (!own && !have && mine.ok(geo)) {
                // the other team's pass: this team's tile is the NEXT one -- build its window now (the other team is filling)
                const long long tw0 = tc_profile_clock();
                live = make_window(mine, X0, Y0);
                have = true;
                if (ld.ok(geo)) {
                    load_pixel(ld);
                    ld.next(geo, ustride);
                    if (ld.ok(geo)) ld.next(geo, ustride);
                }
                tc_profile_add(2, 6, tc_profile_clock() - tw0);          // window of the team's next tile, loads of the one after
            }
            mbar_wait(&info_full[ti % ST_NINFO], (ti / ST_NINFO) & 1, 0x020500, ti);
            const StripTile t = infos[ti % ST_NINFO];
            long long fill_cycles = 0;
            for (int cb = 0; cb < t.ncb; ++cb) {
                const int e0 = X0 - (t.xs + cb * FBW);                                   // box position of window column 0
                const bool cols_hit = have && live && e0 > -K1 && e0 < FBW && !(knobs & 1024);
                for (int j = t.j0; j <= t.j1; ++j) {
                    if (strip_skipped(t, j)) continue;
                    // Only the steps that feed this team's tile are waited for, on the team's OWN "sub-slab free" barrier (its phases
                    // advance with this team's fills alone): the stage's `empty` barrier flips once per step, so a team that sat out
                    // a few steps -- or was busy building its window -- would meet it two phases late and misread the parity.
                    const int slot = it % NSTAGE;
                    if (own || strip_shared(t, j)) {        // as the pass owner, or as the next tile of a shared step
                        mbar_wait(&my_empty[slot], ((fill_par >> slot) & 1u) ^ 1u, 0x020200 | slot, it);
                        fill_par ^= 1u << slot;
                        const long long tf0 = tc_profile_clock();
                        const uint32_t a_stage = a_base + slot * SM::A_STAGE;
                        const int R0 = 2 * j;
                        bool wrote = false;
                        if (!(knobs & 2048)) {
#pragma unroll
                            for (int rr = 0; rr < RCH; ++rr)
                                wrote |= fill_slab_row<K, FBW>(a_stage + rr * SM::FA_SLAB, swz, wsm_a, cols_hit, (R0 + rr) - Y0, e0, dirty,
                                                               1u << (slot * RCH + rr));
                        }
                        if (wrote) fence_proxy_async_smem();
                        mbar_arrive(&my_full[slot]);
                        fill_cycles += tc_profile_clock() - tf0;
                    }
                    ++it;
                }
            }
            tc_profile_add(2, 7, fill_cycles);              // slab fills
            if (own) {                                      // this team's tile is finished: move on to its next one
                have = false;
                mine.next(geo, ustride);
                if (mine.ok(geo)) mine.next(geo, ustride);
            }
        }
    } else {
        // ================================================================= epilogue
        st_reg_inc<ST_REG_EPI>();
        const int q = warp & 3, m = q * 32 + lane;
        int ti = 0;
        {
            // the flow of this thread's pixel is loaded one tile ahead (it only decides "regular or not"): taken on the spot it
            // cost a full global-load latency at the head of every tile's epilogue
            float nfx = 0.f, nfy = 0.f;
            auto load_flow = [&](const TileIt& tl) {
                const int px = tl.gx * GW + (m & 15), py = tl.ty * GH + (m >> 4);
                if (px < W && py < H) {
                    const long long o = (long long)tl.b * 2 * hw + (long long)py * W + px;
                    nfx = flow[o];
                    nfy = flow[o + hw];
                }
            };
            TileIt nxt = first;
            if (nxt.ok(geo)) { load_flow(nxt); nxt.next(geo, ustride); }
            for (TileIt cur = first; cur.ok(geo); cur.next(geo, ustride), ++ti) {
                const int b = cur.b, gx = cur.gx, ty = cur.ty;
                const int px = gx * GW + (m & 15), py = ty * GH + (m >> 4);
                const bool valid = px < W && py < H;
                const long long pofs = (long long)py * W + px;
                bool regular = false;
                const float fx = nfx, fy = nfy;
                if (nxt.ok(geo)) { load_flow(nxt); nxt.next(geo, ustride); }
                if (valid) {
                    AxisTap<float> tx[K], ty_[K];
                    regular = taps_regular<K>(fx, fy, px, py, Hs, Ws, tx, ty_);
                }
                const int buf = ti & 1;
                mbar_wait(&acc_full[buf], (ti >> 1) & 1, 0x030300 | buf, ti);
                tc_fence_after();
                const long long te0 = tc_profile_clock();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * CN;
                // optional fused mask blend (generator.py:130): out = prev * (1 - mask) + attention * mask
                const __nv_bfloat16* pv = (prev == nullptr || !valid) ? nullptr : prev + ((long long)b * hw + pofs) * C + c0;
                const float mk = pv != nullptr ? __bfloat162float(mask[(long long)b * hw + pofs]) : 1.f;
                // The warp's 32 pixels (2 rows x 16) go out as ONE tensor tile per 32 channels: registers -> 64B-swizzled
                // staging -> TMA store.  Per-lane global stores would touch 32 different lines per instruction, and that
                // store path, not HBM, was the largest single cost of the kernel.  The TMA clips the tile at the image edge;
                // irregular pixels (zero weights -> zeros here) are overwritten below.
                const uint32_t ob = smem_u32(smem + SM::OFF_O) + q * SM::O_WARP;
                const uint32_t orow = ob + lane * 64, oswz = (lane >> 1) & 3;
                // one 32-channel slice: (blend) -> bf16 -> swizzled staging -> TMA store
                auto emit = [&](uint32_t (&v)[32], int cc) {
                    if (pv != nullptr) {   // blend in fp32 before the single rounding to bf16
                        const uint4* p4 = reinterpret_cast<const uint4*>(pv + cc * 32);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint4 pq = p4[i];
                            const uint32_t pw[4] = {pq.x, pq.y, pq.z, pq.w};
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                const float2 pf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pw[jj]));
                                v[8 * i + 2 * jj] = __float_as_uint(pf.x * (1.f - mk) + __uint_as_float(v[8 * i + 2 * jj]) * mk);
                                v[8 * i + 2 * jj + 1] = __float_as_uint(pf.y * (1.f - mk) + __uint_as_float(v[8 * i + 2 * jj + 1]) * mk);
                            }
                        }
                    }
                    if (elect_one()) bulk_wait_read<0>();   // the previous tile store (same elected lane) has finished reading the staging buffer
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint32_t pk[4];
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const __nv_bfloat162 t2 = __floats2bfloat162_rn(__uint_as_float(v[8 * i + 2 * jj]), __uint_as_float(v[8 * i + 2 * jj + 1]));
                            pk[jj] = *reinterpret_cast<const uint32_t*>(&t2);
                        }
                        sts128(orow + ((i ^ oswz) << 4), pk[0], pk[1], pk[2], pk[3]);
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (!(knobs & 256) && elect_one()) {
                        tma_store_4d(&tmap_out, ob, c0 + cc * 32, gx * GW, ty * GH + 2 * q, b);
                        bulk_commit();
                    }
                };
                // TMEM loads run one slice ahead of the conversion / store of the previous one (two register sets)
                auto release_acc = [&]() {   // accumulator fully read: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[buf]);
                };
                {
                    constexpr int NCC = CN / 32;
                    static_assert(NCC % 2 == 0, "channel slices are processed in pairs");
                    uint32_t va[32], vb[32];
                    tmem_ld_32x32(taddr, va);
#pragma unroll 1
                    for (int cc = 0; cc < NCC; cc += 2) {
                        tmem_ld_wait();
                        tmem_ld_32x32(taddr + (cc + 1) * 32, vb);
                        emit(va, cc);
                        tmem_ld_wait();
                        if (cc + 2 < NCC) tmem_ld_32x32(taddr + (cc + 2) * 32, va);
                        else release_acc();
                        emit(vb, cc + 1);
                    }
                }
                tc_profile_add(3, 6, tc_profile_clock() - te0);          // TMEM -> registers -> staging -> TMA store
                // irregular pixels (fp32 rounding of (flow+offset)+coord straddling an integer, ~1e-5 of all pixels): the
                // reference's literal 4-taps-per-(i,j) arithmetic, the warp shares one pixel (lanes split the channels)
                unsigned todo = __ballot_sync(0xffffffffu, valid && !regular);
                if (todo) {   // the tile stores above must have landed before these pixels are rewritten
                    if (elect_one()) bulk_wait<0>();
                    __syncwarp();
                }
                while (todo) {
                    const int src_lane = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const int qx = __shfl_sync(0xffffffffu, px, src_lane), qy = __shfl_sync(0xffffffffu, py, src_lane);
                    const float qfx = __shfl_sync(0xffffffffu, fx, src_lane), qfy = __shfl_sync(0xffffffffu, fy, src_lane);
                    irregular_pixel<K, true>(src, logits, out, prev, mask, b, C, c0, CN, Hs, Ws, H, W, qx, qy, qfx, qfy, lane);
                }
            }
        }
    }
    if (warp >= 12 && elect_one()) bulk_wait_read<0>();   // staging buffers stay valid until their last store has read them
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 2 * CN >= 32 ? 2 * CN : 32);
    tc_profile_total(t_start);
}

template <int K, int CN>
static int launch_strip(const void* src, const void* flow, const void* logits, void* out, void* probs, const void* prev,
                        const void* mask, int B, int C, int Hs, int Ws, int H, int W, int ts, cudaStream_t st_) {
    static const PFN_tmapEncodeTiled enc = tmap_encoder();
    if (enc == nullptr) return GFLA_E_NOTSUP;
    CUtensorMap tmap;
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    // (c, x, y, b), box [64 c][32 x]: 128-byte runs, 128B swizzle
    const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)Ws, (cuuint64_t)Hs, (cuuint64_t)B};
    const cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)Ws * C * 2, (cuuint64_t)Hs * Ws * C * 2};
    const cuuint32_t box[4] = {64, ST_FBW, 1, 1};
    if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(src), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return GFLA_E_NOTSUP;
    // out (c, x, y, b), box [32 c][16 x][2 y]: what one epilogue warp stores per step; 64-byte rows, 64B swizzle
    CUtensorMap tmap_o;
    const cuuint64_t odim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t ostr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    const cuuint32_t obox[4] = {32, GW, 2, 1};
    if (enc(&tmap_o, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, out, odim, ostr, obox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return GFLA_E_NOTSUP;
    auto kern = k_local_attn_fwd_strip<K, CN>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, StripSmem<CN>::ALLOC);
    if (e != cudaSuccess) return static_cast<int>(e);
    const int gxn = (W + GW - 1) / GW, gyn = (H + GH - 1) / GH;
    if (ts <= 0) {   // longest strips that still leave every SM several units to balance the tail
        ts = 8;
        while (ts > 1 && (long long)B * gxn * ((gyn + ts - 1) / ts) < 4LL * sm_count()) ts >>= 1;
    }
    const int units = B * gxn * ((gyn + ts - 1) / ts);
    dim3 grid((unsigned)min(units, sm_count()), (unsigned)(C / CN));
    kern<<<grid, ST_THREADS, StripSmem<CN>::ALLOC, st_>>>(tmap, tmap_o, (const __nv_bfloat16*)src, (const float*)flow,
                                                          (const __nv_bfloat16*)logits, (__nv_bfloat16*)out, (__nv_bfloat16*)probs,
                                                          (const __nv_bfloat16*)prev, (const __nv_bfloat16*)mask, B, C, Hs, Ws, H, W, ts,
                                                          tune_knob("GFLA_TC_KNOBS", 0));
    return launch_status();
}

}  // namespace tc

int tc_wait_profile_strip(int enable, unsigned long long* out32) { return tc::tc_wait_profile(enable, out32); }

// channels-last only; same eligibility as local_attn_fwd_tc (checked by the caller).  ts = tiles per strip, 0 = automatic.
int local_attn_fwd_strip_tc(const void* src, const void* flow, const void* logits, void* out, void* probs, const void* prev,
                            const void* mask, int B, int C, int Hs, int Ws, int H, int W, int k, int ts, cudaStream_t st_) {
    const int cn = (C % 256 == 0) ? 256 : C;
#define GFLA_STRIP_CASE(K_, CN_) \
    if (k == K_ && cn == CN_) return tc::launch_strip<K_, CN_>(src, flow, logits, out, probs, prev, mask, B, C, Hs, Ws, H, W, ts, st_);
    GFLA_STRIP_CASE(5, 256) GFLA_STRIP_CASE(5, 128) GFLA_STRIP_CASE(5, 64)
    GFLA_STRIP_CASE(3, 256) GFLA_STRIP_CASE(3, 128) GFLA_STRIP_CASE(3, 64)
#undef GFLA_STRIP_CASE
    return GFLA_E_NOTSUP;
}

}  // namespace gfla
