#!/usr/bin/env python
"""Generate tests/golden/ref_leaves.npz from the REFERENCE's own leaf sources.

    python tests/golden/make_golden.py        (needs /root/reference; run in the build container)

oracle/_ref/libaqsis_refleaf.so is the reference's libs/math/random.cpp, libs/core/multijitter.cpp,
grid.cpp, filters.cpp, bound.cpp and bilinear.h compiled IN PLACE (oracle/Makefile, target `ref`).
The vectors written here are its outputs; tests/test_oracle_leaves.py pins the oracle restatement
and the product's host-side code (RNG replay, sampler tables, filter table) against them
bit-for-bit, on the GPU box too, where /root/reference does not exist.
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import orc  # noqa: E402

SAMPLER_CASES = [(1, 1), (2, 2), (4, 4), (2, 3), (3, 2), (5, 5), (8, 8), (16, 16)]
SAMPLER_ROUNDS = 12

# The (u,v) fixture and the eleven micropolygon shapes of libs/core/bilinear_test.cpp:107-261.
E = 1e-6
UV_FIXTURE = [(0.5, 0.5), (0.12345, 0.67891), (0.42042, 0.42042), (0.3141592, 0.2718281),
              (0, 0), (1, 0), (0, 1), (1, 1), (E, E), (1 - E, E), (E, 1 - E), (1 - E, 1 - E),
              (0.5, 0), (0.5, 1), (0, 0.5), (1, 0.5), (0.5, E), (0.5, 1 - E), (E, 0.5), (1 - E, 0.5)]
# name, A, B, C, D, uExclude, vExclude, relTol, absTol
BILINEAR_CASES = [
    ("convex_irregular", (0.1, 0.1), (1.1, 0), (-0.1, 1.5), (1, 1), -1, -1, 1e-3, 1e-3),
    ("convex_irregular_rot90", (-0.1, 1.5), (0.1, 0.1), (1, 1), (1.1, 0), -1, -1, 1e-3, 1e-3),
    ("convex_irregular_rot180", (1, 1), (-0.1, 1.5), (1.1, 0), (0.1, 0.1), -1, -1, 1e-3, 1e-3),
    ("convex_irregular_rot270", (1.1, 0), (1, 1), (0.1, 0.1), (-0.1, 1.5), -1, -1, 1e-3, 1e-3),
    ("large_offset", (1000, 2000), (1002, 2000), (1000, 2001), (1002, 2001), -1, -1, 1e-3, 1e-3),
    ("exactly_rectangular", (0, 0), (2, 0), (0, 1), (2, 1), -1, -1, 1e-3, 1e-3),
    ("almost_rectangular", (0.0001, 0.000005), (2, 0), (0, 1), (2, 1), -1, -1, 2e-4, 1e-4),
    ("degenerate_u_verts", (0, 0), (0, 0), (0, 1), (1, 1.5), -1, 0, 1e-3, 1e-3),
    ("degenerate_u_verts2", (0, 0), (1.1, 0), (0, 1), (0, 1), -1, 1, 1e-3, 1e-3),
    ("degenerate_v_verts", (0, 0), (1, 0), (0, 0), (1, 1.5), 0, -1, 1e-3, 1e-3),
    ("parallel_adjacent_edges_a", (0, 0), (1, 0), (1, 1.5), (2, 0.01), 1, 0, 0.03, 0.03),
    ("parallel_adjacent_edges_b", (0, 0), (1, 0), (1, 1.5), (2, 0.01), -1, -1, 0.2, 0.2),
]


def uv_list(uex, vex):
    out = []
    for u, v in UV_FIXTURE:
        u32, v32 = np.float32(u), np.float32(v)
        if (uex == -1 or abs(np.float32(uex) - u32) > 1e-2) and (vex == -1 or abs(np.float32(vex) - v32) > 1e-2):
            out.append((u32, v32))
    return np.array(out, dtype=np.float32)


def main():
    R = orc.ref()
    if R is None:
        raise SystemExit("oracle/_ref/libaqsis_refleaf.so is absent and /root/reference is not present")
    R.ref_bilerp2.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_void_p]
    R.ref_bilerp2.restype = None
    R.ref_lfloorf.argtypes = [C.c_float]
    R.ref_lfloorf.restype = C.c_long
    R.ref_lceilf.argtypes = [C.c_float]
    R.ref_lceilf.restype = C.c_long
    out = {}
    # ---- CqRandom
    for seed in (545, 19, 5489, 0, 0xffffffff):
        R.ref_random_reseed(seed)
        out[f"rng_uint_{seed}"] = np.array([R.ref_random_uint() for _ in range(1300)], dtype=np.uint32)
        R.ref_random_reseed(seed)
        out[f"rng_float_{seed}"] = np.array([R.ref_random_float() for _ in range(64)], dtype=np.float32)
        R.ref_random_reseed(seed)
        out[f"rng_int250_{seed}"] = np.array([R.ref_random_int(250) for _ in range(64)], dtype=np.uint32)
    # ---- samplers: Reseed(545) -> ctor -> SAMPLER_ROUNDS x setSamples draw order
    for jitter in (1, 0):
        for xs, ys in SAMPLER_CASES:
            n = xs * ys
            R.ref_random_reseed(545)
            # RenderImage always builds the jittered sampler first (imagebuffer.cpp:694-695)
            sj = R.ref_sampler_create(xs, ys, 1)
            s = sj if jitter else R.ref_sampler_create(xs, ys, 0)
            sh = np.zeros((SAMPLER_ROUNDS, n), np.int32)
            pos = np.zeros((SAMPLER_ROUNDS, n, 2), np.float32)
            dof = np.zeros((SAMPLER_ROUNDS, n, 2), np.float32)
            tm = np.zeros((SAMPLER_ROUNDS, n), np.float32)
            lod = np.zeros((SAMPLER_ROUNDS, n), np.float32)
            for r in range(SAMPLER_ROUNDS):
                R.ref_sampler_draw(s, sh[r].ctypes.data, pos[r].ctypes.data, dof[r].ctypes.data,
                                   tm[r].ctypes.data, lod[r].ctypes.data)
            tag = f"samp_j{jitter}_{xs}x{ys}"
            out[tag + "_shuf"], out[tag + "_pos"], out[tag + "_dof"], out[tag + "_time"], out[tag + "_lod"] = sh, pos, dof, tm, lod
            # the stream position after the draws pins how many numbers each getter consumed
            out[tag + "_next"] = np.array([R.ref_random_uint() for _ in range(4)], dtype=np.uint32)
            if not jitter:
                R.ref_sampler_destroy(s)
            R.ref_sampler_destroy(sj)
    # ---- pixel filters
    rng = np.random.default_rng(7)
    pts = np.concatenate([rng.uniform(-3.5, 3.5, (400, 2)),
                          np.array([[0, 0], [0.5, 0.5], [-0.5, 0.5], [1, 0], [0, 1], [1, 1], [2, 0], [1.5, -1.5],
                                    [0.25, 0.0], [3, 3], [-3, 3], [0.999999, 0], [1.000001, 0]])]).astype(np.float32)
    widths = np.array([[1, 1], [2, 2], [3, 3], [4, 4], [5, 5], [6, 6], [2, 3], [2.5, 2.5], [7, 7]], dtype=np.float32)
    fv = np.zeros((8, len(widths), len(pts)), np.float32)
    for which in range(8):
        for wi, (xw, yw) in enumerate(widths):
            for pi, (x, y) in enumerate(pts):
                fv[which, wi, pi] = R.ref_filter(which, float(x), float(y), float(xw), float(yw))
    out["filter_pts"], out["filter_widths"], out["filter_values"] = pts, widths, fv
    # ---- inverse bilinear: the reference test's cases + random quads
    for name, A, B, Cc, D, uex, vex, _, _ in BILINEAR_CASES:
        verts = np.array([A, B, Cc, D], dtype=np.float32).ravel()
        uvs = uv_list(uex, vex)
        P = np.zeros_like(uvs)
        uvo = np.zeros_like(uvs)
        for i, (u, v) in enumerate(uvs):
            R.ref_bilerp2(verts.ctypes.data, float(u), float(v), P[i].ctypes.data)
            R.ref_invbilinear(verts.ctypes.data, float(P[i, 0]), float(P[i, 1]), uvo[i].ctypes.data)
        out[f"bil_{name}_verts"], out[f"bil_{name}_uvin"], out[f"bil_{name}_P"], out[f"bil_{name}_uvout"] = verts, uvs, P, uvo
    q = (rng.uniform(0, 1, (300, 4, 2)) * 0.6 + np.array([[0, 0], [1, 0], [0, 1], [1, 1]])).astype(np.float32)
    q[:150] += rng.uniform(0, 1900, (150, 1, 2)).astype(np.float32)           # raster-sized offsets
    q[100:150] = (q[100:150] - q[100:150, :1]) * np.float32(0.01) + q[100:150, :1]  # sub-pixel micropolygons
    qp = (q.mean(axis=1) + rng.uniform(-0.3, 0.3, (300, 2)) * (q[:, 3] - q[:, 0])).astype(np.float32)
    quv = np.zeros((300, 2), np.float32)
    qz = np.zeros(300, np.float32)
    zc = rng.uniform(1, 100, (300, 4)).astype(np.float32)
    for i in range(300):
        v = np.ascontiguousarray(q[i].ravel())
        R.ref_invbilinear(v.ctypes.data, float(qp[i, 0]), float(qp[i, 1]), quv[i].ctypes.data)
        qz[i] = R.ref_bilerp(float(zc[i, 0]), float(zc[i, 1]), float(zc[i, 2]), float(zc[i, 3]), float(quv[i, 0]), float(quv[i, 1]))
    out["bilq_verts"], out["bilq_P"], out["bilq_uv"], out["bilq_zc"], out["bilq_z"] = q.reshape(300, 8), qp, quv, zc, qz
    # ---- rounding helpers (math.h:47-70)
    xs_ = np.concatenate([rng.uniform(-1000, 1000, 200), np.arange(-5, 6), np.arange(-5, 6) + 0.5,
                          [254.5, 255.49999, 255.5, -0.0, 1e-9, -1e-9]]).astype(np.float64)
    out["round_x"] = xs_
    out["round_lfloor"] = np.array([R.ref_lfloor(float(x)) for x in xs_], np.int64)
    out["round_lceil"] = np.array([R.ref_lceil(float(x)) for x in xs_], np.int64)
    out["round_lround"] = np.array([R.ref_lround(float(x)) for x in xs_], np.int64)
    # ---- CqBound::Contains2D / Intersects (bound.h:144-160)
    b6 = np.array([1.0, 2.0, 3.0, 4.0, 6.0, 9.0], np.float32)
    bp = np.array([[1, 2], [4, 6], [1, 6], [4, 2], [0.999999, 3], [4.000001, 3], [2, 1.999999], [2, 6.000001], [2.5, 4]], np.float32)
    out["bound_b6"], out["bound_pts"] = b6, bp
    out["bound_contains"] = np.array([R.ref_bound_contains2d(b6.ctypes.data, float(x), float(y)) for x, y in bp], np.int32)
    np.savez_compressed(os.path.join(HERE, "ref_leaves.npz"), **out)
    print("wrote", os.path.join(HERE, "ref_leaves.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
