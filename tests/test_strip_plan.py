"""Host-side check of the strip forward kernel's step schedule (csrc/strip_plan.h, plain integer logic).

The header is compiled with g++ into a throw-away shared object; random tile boxes are run through
`strip_plan` the way the producer warp does, and the invariants the MMA / builder warps rely on are asserted:
every source row chunk of every tile's footprint is accumulated exactly once (own pass or the previous
tile's shared steps), shared steps cover the next tile's columns, every pass has at least one step.
"""
import ctypes
import os
import random
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "global-flow-local-attention_b200", "csrc")

SHIM = r"""
#include "strip_plan.h"
using namespace gfla::tc;
extern "C" void plan(const int* cur, int has_next, const int* nxt, int k0, int k1, int fbw, int* out) {
    TileBox c{cur[0], cur[1], cur[2], cur[3]}, n{nxt[0], nxt[1], nxt[2], nxt[3]};
    StripTile t = strip_plan(c, has_next != 0, n, k0, k1, fbw);
    out[0] = t.xs; out[1] = t.ncb; out[2] = t.j0; out[3] = t.j1; out[4] = t.k0; out[5] = t.k1; out[6] = t.s0; out[7] = t.s1;
    out[8] = strip_steps(t);
    int n_exec = 0;
    for (int j = t.j0; j <= t.j1; ++j) n_exec += strip_skipped(t, j) ? 0 : 1;
    out[9] = n_exec;
}
"""


@pytest.fixture(scope="module")
def plan_lib():
    d = tempfile.mkdtemp(prefix="strip_plan_")
    src = os.path.join(d, "shim.cpp")
    with open(src, "w") as f:
        f.write(SHIM)
    so = os.path.join(d, "shim.so")
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I", CSRC, src, "-o", so], check=True)
    lib = ctypes.CDLL(so)
    lib.plan.argtypes = [ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_int,
                         ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    return lib


def _plan(lib, cur, has_next, nxt, k0, k1, fbw=32):
    out = (ctypes.c_int * 10)()
    lib.plan((ctypes.c_int * 4)(*cur), int(has_next), (ctypes.c_int * 4)(*nxt), k0, k1, fbw, out)
    keys = ("xs", "ncb", "j0", "j1", "k0", "k1", "s0", "s1", "steps", "n_exec")
    return dict(zip(keys, list(out)))


def _random_strip(rng, n_tiles, hs, ws, wild):
    """boxes the way group_bbox produces them: clamped to the image, 16x8 pixels + k taps + flow variation"""
    boxes = []
    y = rng.randint(0, 8)
    x = rng.randint(0, ws - 1)
    for _ in range(n_tiles):
        dx, dy = (rng.randint(-40, 40), rng.randint(-20, 20)) if wild else (rng.randint(-3, 3), rng.randint(-3, 3))
        x0 = min(max(x + dx, 0), ws - 1)
        y0 = min(max(y + dy, 0), hs - 1)
        x1 = min(x0 + rng.randint(15, 70 if wild else 30), ws - 1)
        y1 = min(y0 + rng.randint(7, 40 if wild else 18), hs - 1)
        boxes.append((x0, y0, x1, y1))
        y += 8
    return boxes


@pytest.mark.parametrize("wild", [False, True])
def test_strip_plan_invariants(plan_lib, wild):
    rng = random.Random(1234 + wild)
    n_shared = 0
    for _ in range(400):
        n = rng.randint(1, 9)
        boxes = _random_strip(rng, n, hs=rng.choice([8, 37, 256]), ws=rng.choice([16, 50, 256]), wild=wild)
        acc = [dict() for _ in range(n)]            # tile -> chunk -> times accumulated
        k0, k1 = 1, 0
        for t, cur in enumerate(boxes):
            has_next = t + 1 < n
            nxt = boxes[t + 1] if has_next else (0, 0, 0, 0)
            p = _plan(plan_lib, cur, has_next, nxt, k0, k1)
            assert p["ncb"] >= 1 and p["xs"] <= cur[0] and p["xs"] + 32 * p["ncb"] > cur[2]
            assert (p["j0"], p["j1"]) == (cur[1] >> 1, cur[3] >> 1)
            assert (p["k0"], p["k1"]) == (k0, k1)
            if k0 <= k1:
                assert p["j0"] <= k0 and k1 <= p["j1"]
            assert p["steps"] == p["n_exec"] >= 1
            shared = p["s0"] <= p["s1"]
            if shared:
                n_shared += 1
                assert has_next
                assert p["xs"] <= nxt[0] and p["xs"] + 32 * p["ncb"] > nxt[2]       # next tile's columns are loaded
                assert p["j0"] <= p["s0"] and p["s1"] <= p["j1"]
                assert (nxt[1] >> 1) <= p["s0"] and p["s1"] <= (nxt[3] >> 1)
                assert k0 > k1 or k1 < p["s0"] or p["s1"] < k0                        # disjoint from the skipped chunks
            else:
                assert p["xs"] == cur[0]
            for j in range(p["j0"], p["j1"] + 1):
                if k0 <= j <= k1:
                    continue
                acc[t][j] = acc[t].get(j, 0) + 1
                if shared and p["s0"] <= j <= p["s1"]:
                    acc[t + 1][j] = acc[t + 1].get(j, 0) + 1
            k0, k1 = p["s0"], p["s1"]
        for t, cur in enumerate(boxes):
            want = set(range(cur[1] >> 1, (cur[3] >> 1) + 1))
            assert set(acc[t]) == want and all(v == 1 for v in acc[t].values()), (t, boxes)
    if not wild:
        assert n_shared > 200        # smooth strips actually share
