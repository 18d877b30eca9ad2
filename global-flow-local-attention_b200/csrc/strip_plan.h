// Step schedule of the strip forward kernel (local_attn_strip_tc.cu): plain integer logic, host-testable
// (tests/test_strip_plan.py compiles this header with g++ and checks the invariants on random boxes).
//
// A CTA walks down a column of vertically adjacent 16x8 pixel tiles.  The tap footprints of neighbouring
// tiles overlap by (k + flow variation) source rows; a source row chunk that tile t loads for itself and
// that tile t+1 needs as well is fed to BOTH accumulators while it sits in shared memory ("shared" steps),
// and tile t+1 then skips it in its own pass.  Rows are handled in chunks of 2 (absolute chunk index
// j = row >> 1) so that the chunk grids of all tiles coincide.
#pragma once

#ifdef __CUDACC__
#define GFLA_HD __host__ __device__ __forceinline__
#else
#define GFLA_HD inline
#endif

namespace gfla {
namespace tc {

struct TileBox { int x0, y0, x1, y1; };   // clamped tap bounding box of one tile (inclusive)

// One tile's pass: for cb in [0, ncb), for j in [j0, j1] except [k0, k1]: load rows 2j, 2j+1 of the box
// columns [xs + cb*fbw, +fbw) and accumulate them into this tile; chunks in [s0, s1] (disjoint from the
// skipped ones) are also accumulated into the NEXT tile, whose pass then skips exactly [s0, s1].
// Empty intervals are encoded as (1, 0).
struct StripTile { int xs, ncb, j0, j1, k0, k1, s0, s1; };

GFLA_HD bool strip_skipped(const StripTile& t, int j) { return j >= t.k0 && j <= t.k1; }
GFLA_HD bool strip_shared(const StripTile& t, int j) { return j >= t.s0 && j <= t.s1; }

// (k0, k1) = the chunks the previous tile's pass already accumulated into this tile (its [s0, s1]).
GFLA_HD StripTile strip_plan(const TileBox& cur, bool has_next, const TileBox& nxt, int k0, int k1, int fbw) {
    StripTile t;
    t.ncb = (cur.x1 - cur.x0 + fbw) / fbw;
    t.xs = cur.x0;
    t.j0 = cur.y0 >> 1;
    t.j1 = cur.y1 >> 1;
    t.k0 = k0; t.k1 = k1;
    t.s0 = 1; t.s1 = 0;
    if (has_next) {
        const int ux0 = cur.x0 < nxt.x0 ? cur.x0 : nxt.x0, ux1 = cur.x1 > nxt.x1 ? cur.x1 : nxt.x1;
        if ((ux1 - ux0 + fbw) / fbw <= t.ncb) {   // both tiles' columns fit the boxes this pass loads anyway
            const int nj0 = nxt.y0 >> 1, nj1 = nxt.y1 >> 1;
            int s0 = t.j0 > nj0 ? t.j0 : nj0, s1 = t.j1 < nj1 ? t.j1 : nj1;
            if (k0 <= k1 && s0 <= k1 && k0 <= s1) s0 = k1 + 1;    // chunks this pass skips cannot be shared
            if (s0 <= nj0 && s1 >= nj1) s1 = nj1 - 1;             // leave the next pass at least one own step
            if (s0 <= s1) { t.xs = ux0; t.s0 = s0; t.s1 = s1; }
        }
    }
    return t;
}

// number of (chunk) steps per column block of the pass
GFLA_HD int strip_steps(const StripTile& t) {
    int n = t.j1 - t.j0 + 1;
    if (t.k0 <= t.k1) n -= (t.k1 - t.k0 + 1);
    return n;
}

}  // namespace tc
}  // namespace gfla
