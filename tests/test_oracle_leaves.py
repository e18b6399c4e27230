"""Pin the leaves of the path against the REFERENCE's own code (CPU only).

tests/golden/ref_leaves.npz holds outputs of the reference's leaf sources compiled in place
(oracle/_ref, see tests/golden/make_golden.py).  Both the oracle restatement (oracle/) and the
product's host-side code (the RNG replay, sampler tables and filter table inside
libaqsis_b200_hider.so, reached through the C ABI) must reproduce them BIT FOR BIT.
The inverse-bilinear known-answer cases and tolerances are those of the reference's own
unit test, libs/core/bilinear_test.cpp:107-261.
"""
import ctypes as C
import os

import numpy as np
import pytest

import orc
from golden.make_golden import BILINEAR_CASES, SAMPLER_CASES, SAMPLER_ROUNDS

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_leaves.npz"))


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


# ---------------------------------------------------------------- CqRandom (random.cpp:97-224)
@pytest.mark.parametrize("seed", [545, 19, 5489, 0, 0xffffffff])
def test_rng_oracle(seed):
    L = orc.lib()
    L.orc_random_reseed(seed)
    assert np.array_equal(np.array([L.orc_random_uint() for _ in range(1300)], np.uint32), G[f"rng_uint_{seed}"])
    L.orc_random_reseed(seed)
    assert np.array_equal(bits([L.orc_random_float() for _ in range(64)]), bits(G[f"rng_float_{seed}"]))
    L.orc_random_reseed(seed)
    assert np.array_equal(np.array([L.orc_random_int(250) for _ in range(64)], np.uint32), G[f"rng_int250_{seed}"])


@pytest.mark.parametrize("seed", [545, 19, 5489, 0, 0xffffffff])
def test_rng_product(native_lib, seed):
    L = native_lib
    r = L.aqh_random_create(seed)
    try:
        assert np.array_equal(np.array([L.aqh_random_uint(r) for _ in range(1300)], np.uint32), G[f"rng_uint_{seed}"])
        L.aqh_random_reseed(r, seed)
        assert np.array_equal(bits([L.aqh_random_float(r) for _ in range(64)]), bits(G[f"rng_float_{seed}"]))
        L.aqh_random_reseed(r, seed)
        assert np.array_equal(np.array([L.aqh_random_int(r, 250) for _ in range(64)], np.uint32), G[f"rng_int250_{seed}"])
    finally:
        L.aqh_random_destroy(r)


def test_rng_survey_known_answers():
    """SURVEY.md appendix B values obtained from the reference at survey time."""
    assert G["rng_uint_19"][0] == 418903645
    assert list(G["rng_int250_19"][:10]) == [24, 107, 190, 103, 61, 178, 34, 64, 82, 168]
    assert np.allclose(G["rng_float_545"][:2], [0.0863071531, 0.485246003], rtol=0, atol=1e-9)
    assert list(G["samp_j1_4x4_shuf"][0][:4]) == [12, 8, 15, 7]
    assert np.allclose(G["samp_j1_4x4_pos"][0][:2], [[0.0362413637, 0.225688934], [0.487524688, 0.0274800472]], atol=1e-9)


# ------------------------------------------- samplers (multijitter.cpp:83-222, grid.cpp:37-63)
def _check_sampler(tables_fn, next_uint, rand_int, xs, ys, jitter):
    """tables_fn() builds the tables from an RNG reseeded with 545; then the per-pixel draw
    order of CqImagePixel::setSamples (imagepixel.cpp:338-347) is replayed."""
    n = xs * ys
    pos, v1d, shuf, ncache = tables_fn()
    assert ncache == (250 if jitter else 1)
    tag = f"samp_j{jitter}_{xs}x{ys}"
    for r in range(SAMPLER_ROUNDS):
        # the grid sampler hands out its single pattern without touching the stream
        k = [rand_int(250) if jitter else 0 for _ in range(5)]
        assert np.array_equal(shuf[k[0]], G[tag + "_shuf"][r]), (tag, r)
        assert np.array_equal(bits(pos[k[1]]), bits(G[tag + "_pos"][r])), (tag, r)
        assert np.array_equal(bits(pos[k[2]]), bits(G[tag + "_dof"][r])), (tag, r)
        assert np.array_equal(bits(v1d[k[3]]), bits(G[tag + "_time"][r])), (tag, r)
        assert np.array_equal(bits(v1d[k[4]]), bits(G[tag + "_lod"][r])), (tag, r)
    assert [next_uint() for _ in range(4)] == list(G[tag + "_next"])


@pytest.mark.parametrize("jitter", [1, 0])
@pytest.mark.parametrize("xs,ys", SAMPLER_CASES)
def test_sampler_oracle(xs, ys, jitter):
    L = orc.lib()
    n = xs * ys

    def tables():
        L.orc_random_reseed(545)
        nc = 250
        pos = np.zeros((nc, n, 2), np.float32)
        v1d = np.zeros((nc, n), np.float32)
        shuf = np.zeros((nc, n), np.int32)
        # RenderImage always constructs the jittered sampler (consumes the stream, reseeds 19)
        got = L.orc_sampler_tables(xs, ys, 1, pos.ctypes.data, v1d.ctypes.data, shuf.ctypes.data)
        if not jitter:
            got = L.orc_sampler_tables(xs, ys, 0, pos.ctypes.data, v1d.ctypes.data, shuf.ctypes.data)
        return pos, v1d, shuf, got

    _check_sampler(tables, L.orc_random_uint, L.orc_random_int, xs, ys, jitter)


@pytest.mark.parametrize("jitter", [1, 0])
@pytest.mark.parametrize("xs,ys", SAMPLER_CASES)
def test_sampler_product(native_lib, xs, ys, jitter):
    L = native_lib
    n = xs * ys
    r = L.aqh_random_create(545)

    def tables():
        nc = C.c_int()
        pos = np.zeros((250, n, 2), np.float32)
        v1d = np.zeros((250, n), np.float32)
        shuf = np.zeros((250, n), np.int32)
        assert L.aqh_sampler_tables(r, xs, ys, 1, pos.ctypes.data, v1d.ctypes.data, shuf.ctypes.data, C.byref(nc)) == 0
        if not jitter:
            assert L.aqh_sampler_tables(r, xs, ys, 0, pos.ctypes.data, v1d.ctypes.data, shuf.ctypes.data, C.byref(nc)) == 0
        return pos, v1d, shuf, nc.value

    try:
        _check_sampler(tables, lambda: L.aqh_random_uint(r), lambda k: L.aqh_random_int(r, k), xs, ys, jitter)
    finally:
        L.aqh_random_destroy(r)


def test_grid_sampler_times_are_zero():
    """CqGridSampler's `dt = 1/nSamples` is an integer division: all times 0 (grid.cpp:52)."""
    assert np.all(G["samp_j0_4x4_time"] == 0) and np.all(G["samp_j0_8x8_lod"] == 0)


# ------------------------------------------------------------ pixel filters (filters.cpp:71-348)
FILTERS = ["box", "triangle", "gaussian", "catmullrom", "sinc", "mitchell", "disk", "bessel"]


@pytest.mark.parametrize("which", range(8))
def test_filters_oracle_and_product(native_lib, which):
    pts, widths, want = G["filter_pts"], G["filter_widths"], G["filter_values"][which]
    O = orc.lib()
    fn = getattr(native_lib, f"aqh_{FILTERS[which]}_filter")
    for wi, (xw, yw) in enumerate(widths):
        o = np.array([O.orc_filter(which, float(x), float(y), float(xw), float(yw)) for x, y in pts], np.float32)
        p = np.array([fn(float(x), float(y), float(xw), float(yw)) for x, y in pts], np.float32)
        assert np.array_equal(bits(o), bits(want[wi])), (FILTERS[which], xw, yw)
        assert np.array_equal(bits(p), bits(want[wi])), (FILTERS[which], xw, yw)


def test_filter_survey_known_answers(native_lib):
    assert abs(native_lib.aqh_gaussian_filter(.3, .2, 2, 2) - 0.771051586) < 1e-7
    assert abs(native_lib.aqh_catmullrom_filter(.3, .2, 3, 3) - 1.49061644) < 1e-6
    assert abs(native_lib.aqh_sinc_filter(.3, .2, 4, 4) - 0.794993699) < 1e-7


# ----------------------------------------------- inverse bilinear (bilinear.h:99-310) + bilerp
@pytest.mark.parametrize("case", BILINEAR_CASES, ids=[c[0] for c in BILINEAR_CASES])
def test_invbilinear_reference_unit_test(case):
    """libs/core/bilinear_test.cpp: forward bilerp, invert, IsCloseRelAbs(relTol, absTol)."""
    name, *_, rel, ab = case
    L = orc.lib()
    verts, uvin, P, uvref = (G[f"bil_{name}_{k}"] for k in ("verts", "uvin", "P", "uvout"))
    for i in range(len(uvin)):
        uv = np.zeros(2, np.float32)
        L.orc_invbilinear(verts.ctypes.data, float(P[i, 0]), float(P[i, 1]), uv.ctypes.data)
        # bit-identical to the reference's CqInvBilinear ...
        assert np.array_equal(bits(uv), bits(uvref[i])), (name, i)
        # ... and inside the reference test's own tolerance
        for a, b in zip(uvin[i], uv):
            d = abs(np.float32(a) - np.float32(b))
            assert d < ab or d < rel * abs(a) or d < rel * abs(b), (name, i, uvin[i], uv)
        assert -ab <= uv[0] <= 1 + ab


def test_invbilinear_random_quads_bit_exact():
    L = orc.lib()
    V, P, UV, ZC, Z = G["bilq_verts"], G["bilq_P"], G["bilq_uv"], G["bilq_zc"], G["bilq_z"]
    for i in range(len(V)):
        uv = np.zeros(2, np.float32)
        v = np.ascontiguousarray(V[i])
        L.orc_invbilinear(v.ctypes.data, float(P[i, 0]), float(P[i, 1]), uv.ctypes.data)
        assert np.array_equal(bits(uv), bits(UV[i])), i
        z = L.orc_bilerp(*[float(x) for x in ZC[i]], float(uv[0]), float(uv[1]))
        assert bits([z])[0] == bits([Z[i]])[0], i


# ----------------------------------------------------- frame tables: oracle == product
def _params(native_lib, **kw):
    from aqsis_b200 import default_params
    return default_params(**kw)


@pytest.mark.parametrize("name,w", [("box", 1.0), ("triangle", 2.0), ("gaussian", 2.0), ("catmull-rom", 3.0),
                                    ("sinc", 4.0), ("gaussian", 5.0), ("sinc", 6.0), ("mitchell", 2.5)])
@pytest.mark.parametrize("samples", [(1, 1), (4, 4), (3, 5)])
def test_filter_table_product_matches_oracle(native_lib, name, w, samples):
    p = _params(native_lib, resolution=(64, 48), samples=samples, filter=(name, w, w))
    n = C.c_int()
    assert native_lib.aqh_filter_table(C.byref(p), None, C.byref(n)) == 0
    shift = int(np.floor(w / 2))
    assert n.value == (2 * shift + 1) ** 2 * samples[0] * samples[1]
    a = np.zeros(n.value, np.float32)
    b = np.zeros(n.value, np.float32)
    native_lib.aqh_filter_table(C.byref(p), a.ctypes.data, C.byref(n))
    assert orc.lib().orc_filter_table(C.byref(p), b.ctypes.data) == n.value
    assert np.array_equal(bits(a), bits(b))


@pytest.mark.parametrize("res,crop,bucket,fw", [((40, 24), None, (16, 16), 2.0), ((50, 37), (5, 44, 3, 30), (16, 16), 3.0),
                                                ((33, 33), None, (8, 12), 4.0), ((20, 20), None, (16, 16), 1.0),
                                                ((64, 32), (16, 48, 0, 32), (16, 16), 6.0)])
@pytest.mark.parametrize("jitter", [1, 0])
def test_rng_replay_product_matches_oracle(native_lib, res, crop, bucket, fw, jitter):
    """Appendix B of SURVEY.md: bucket-order replay of the global stream into per-pixel pattern
    planes + dither planes; the first sampled pixel must use the survey's known patterns."""
    kw = dict(resolution=res, samples=(2, 2), filter=("gaussian", fw, fw), jitter=jitter,
              bucket_xsize=bucket[0], bucket_ysize=bucket[1],
              displays=[("rgba", 1, 255.0, 0.0, 255.0, 0.5), ("z", 0, 0.0, 0.0, 0.0, 0.0)])
    if crop:
        kw["crop"] = crop
    p = _params(native_lib, **kw)
    geo = [C.c_int() for _ in range(4)]
    assert native_lib.aqh_replay_frame_rng(C.byref(p), None, None, *[C.byref(g) for g in geo]) == 0
    sx0, sy0, sw, sh = [g.value for g in geo]
    shift = int(np.floor(fw / 2))
    assert (sx0, sy0) == (p.crop_xmin - shift, p.crop_ymin - shift)
    assert (sw, sh) == (p.crop_xmax - p.crop_xmin + 2 * shift, p.crop_ymax - p.crop_ymin + 2 * shift)
    pa = np.zeros(5 * sw * sh, np.uint8)
    pb = np.zeros_like(pa)
    da = np.zeros(2 * res[0] * res[1], np.float32)
    db = np.zeros_like(da)
    assert native_lib.aqh_replay_frame_rng(C.byref(p), pa.ctypes.data, da.ctypes.data, *[C.byref(g) for g in geo]) == 0
    assert orc.lib().orc_replay(C.byref(p), pb.ctypes.data, db.ctypes.data, *[C.byref(g) for g in geo]) == 0
    assert np.array_equal(pa, pb)
    assert np.array_equal(bits(da), bits(db))
    if jitter:
        # first pixel of the first bucket: patterns 24, 107, 190, 103, 61
        assert list(pa.reshape(5, sh, sw)[:, 0, 0]) == [24, 107, 190, 103, 61]
        assert list(pa.reshape(5, sh, sw)[:, 0, 1]) == [178, 34, 64, 82, 168]
    else:
        assert not pa.any()


def test_rounding_helpers():
    """lfloor / lceil / lround of include/aqsis/math/math.h:47-70 as restated in oracle and kernels."""
    x = G["round_x"]
    lf = np.array([int(v) - (1 if (v < 0 and v != int(v)) else 0) for v in x])
    lc = np.array([int(v) + (1 if (v > 0 and v != int(v)) else 0) for v in x])
    assert np.array_equal(lf, G["round_lfloor"]) and np.array_equal(lc, G["round_lceil"])
    xm = x - 0.5
    lr = np.array([int(v) - (1 if (v < 0 and v != int(v)) else 0) + 1 for v in xm])
    assert np.array_equal(lr, G["round_lround"])


def test_bound_contains2d_inclusive():
    b, pts = G["bound_b6"], G["bound_pts"]
    mine = [int(b[0] <= x <= b[3] and b[1] <= y <= b[4]) for x, y in pts]
    assert mine == list(G["bound_contains"])


@pytest.mark.skipif(orc.ref() is None, reason="oracle/_ref not built (reference tree absent)")
def test_golden_matches_live_reference_build():
    """When the reference leaf library is present, spot-check that the committed fixtures are its output."""
    R = orc.ref()
    R.ref_random_reseed(545)
    assert [R.ref_random_uint() for _ in range(16)] == list(G["rng_uint_545"][:16])
    pts, widths = G["filter_pts"], G["filter_widths"]
    for which in range(8):
        v = np.array([R.ref_filter(which, float(x), float(y), float(widths[1][0]), float(widths[1][1])) for x, y in pts[:50]], np.float32)
        assert np.array_equal(bits(v), bits(G["filter_values"][which, 1, :50]))
