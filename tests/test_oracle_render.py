"""Frame-level checks of the CPU oracle (CPU only): known answers that follow from the reference's
semantics, size-independent properties, and a checksum pin of small renders.

The reference has no golden images or bucket-level tests for this path, so
the frame-level stages are pinned here by cases whose answer can be derived by hand from the
reference source: regular-grid sampling (`Hider "jitter" 0`, grid.cpp:37-63), the edge rule of
CqMicroPolygon::fContains (micropolygon.cpp:1306-1335), the over operator of CqImagePixel::Combine
(imagepixel.cpp:217-222) and the normalised gather of FilterBucket (bucketprocessor.cpp:584-707).
"""
import hashlib
import json
import os

import numpy as np
import pytest

import orc
from aqsis_b200 import abi, default_params, scenes, GridArrays

FLT_MAX = np.float32(3.4028234663852886e38)
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "oracle_frames.json")


def one_grid(xs, ys, z=5.0, ci=(0.25, 0.5, 0.75), oi=(1, 1, 1), flags=0):
    """A (len(xs)-1) x (len(ys)-1) grid with vertices on the lattice xs x ys."""
    xs, ys = np.asarray(xs, np.float32), np.asarray(ys, np.float32)
    X, Y = np.meshgrid(xs, ys)
    nv = X.size
    P = np.stack([X.ravel(), Y.ravel(), np.full(nv, z, np.float32)], axis=1).astype(np.float32)
    return GridArrays(cu=np.array([len(xs) - 1], np.int32), cv=np.array([len(ys) - 1], np.int32),
                      flags=np.array([flags], np.uint32), P=P,
                      Ci=np.tile(np.asarray(ci, np.float32), (nv, 1)), Oi=np.tile(np.asarray(oi, np.float32), (nv, 1)))


def params_1spp(res=(10, 8), **kw):
    kw.setdefault("samples", (1, 1))
    kw.setdefault("filter", ("box", 1.0, 1.0))
    kw.setdefault("jitter", 0)
    kw.setdefault("displays", [("rgba", 1, 255.0, 0.0, 255.0, 0.0)])
    return default_params(resolution=res, **kw)


def test_empty_frame():
    p = default_params(resolution=(40, 24), samples=(2, 2), displays=[("rgba", 1, 255.0, 0.0, 255.0, 0.5)])
    g = GridArrays(cu=np.zeros(0, np.int32), cv=np.zeros(0, np.int32), flags=np.zeros(0, np.uint32), P=np.zeros((0, 3), np.float32))
    ch, disp, _ = orc.render(p, g, 1)
    assert np.all(ch[..., :7] == 0) and np.all(ch[..., abi.CH_COVERAGE] == 0) and np.all(ch[..., abi.CH_Z] == FLT_MAX)
    assert disp[0].max() <= 1     # lround(0 + 0.5*dither): dither noise only (ddmanager.cpp:1065-1070)


def test_regular_grid_rectangle_coverage():
    """jitter 0, 1 sample at the pixel centre, box 1x1: a pixel is covered iff its centre is inside."""
    p = params_1spp()
    # constant shading (corner 0, micropolygon.cpp:1518-1521): the colour comes through untouched
    ch, disp, _ = orc.render(p, one_grid([2.25, 6.75], [1.25, 4.75], z=7.0, flags=0), 1)
    want = np.zeros((8, 10), bool)
    want[1:5, 2:7] = True
    assert np.array_equal(ch[..., abi.CH_COVERAGE] == 1.0, want)
    assert np.all(ch[want][:, :3] == np.float32([0.25, 0.5, 0.75])) and np.all(ch[want][:, 3:6] == 1.0)
    assert np.all(ch[want][:, abi.CH_ALPHA] == 1.0) and np.allclose(ch[want][:, abi.CH_Z], 7.0, rtol=3e-7)
    assert np.all(ch[~want][:, :7] == 0) and np.all(ch[~want][:, abi.CH_Z] == FLT_MAX)
    # Quantize 255 0 255 0: lround(255*v) (ddmanager.cpp:1065-1070)
    assert np.all(disp[0][want] == [64, 128, 191, 255]) and np.all(disp[0][~want] == 0)


def test_shared_edge_is_hit_exactly_once():
    """Sample points exactly on the edge shared by two micropolygons belong to exactly one of them
    (edges 0,1 fail on <= 0, edges 2,3 on < 0; micropolygon.cpp:1313-1330).  With Oi = 0.5 a double
    hit would composite to 0.75 and a crack would leave 0."""
    p = params_1spp()
    for xs, ys in [([2.5, 4.5, 6.5], [1.25, 4.75]), ([2.25, 6.75], [1.5, 3.5, 5.5]), ([2.5, 4.5, 6.5], [1.5, 3.5, 5.5])]:
        g = one_grid(xs, ys, ci=(0.5, 0.5, 0.5), oi=(0.5, 0.5, 0.5))
        ch, _, st = orc.render(p, g, 1)
        o = ch[..., abi.CH_OI_R]
        assert set(np.unique(o)) <= {0.0, 0.5}, np.unique(o)
        inner = o[int(np.ceil(ys[0])):int(np.floor(ys[-1])), int(np.ceil(xs[0])):int(np.floor(xs[-1]))]
        assert np.all(inner == 0.5)
        # the outer boundary is half open: of two opposite boundary lines through sample points exactly one is kept
        if xs[0] == 2.5:
            col_l, col_r = o[2:4, 2], o[2:4, 6]
            assert (np.all(col_l == 0.5)) != (np.all(col_r == 0.5))
        if ys[0] == 1.5:
            row_t, row_b = o[1, 3:6], o[5, 3:6]
            assert (np.all(row_t == 0.5)) != (np.all(row_b == 0.5))


def test_opaque_nearest_wins_and_ties_keep_first():
    p = params_1spp()
    near = one_grid([1.25, 8.75], [1.25, 6.75], z=3.0, ci=(1, 0, 0))
    far = one_grid([1.25, 8.75], [1.25, 6.75], z=9.0, ci=(0, 1, 0))
    for order in ([near, far], [far, near]):
        ch, _, _ = orc.render(p, scenes.concat(order), 1)
        assert np.all(ch[2:6, 2:8, :3] == np.float32([1, 0, 0])) and np.allclose(ch[2:6, 2:8, abi.CH_Z], 3.0, rtol=3e-7)
    # equal depth: occlZ <= D drops the later hit (bucketprocessor.cpp:1475)
    a = one_grid([1.25, 8.75], [1.25, 6.75], z=4.0, ci=(0, 0, 1))
    b = one_grid([1.25, 8.75], [1.25, 6.75], z=4.0, ci=(1, 1, 0))
    ch, _, _ = orc.render(p, scenes.concat([a, b]), 1)
    assert np.all(ch[2:6, 2:8, :3] == np.float32([0, 0, 1]))


def test_transparent_layers_composite_back_to_front():
    """C = C*(1-clamp(O)) + Ci ; A = (1-A)*O + A, farthest first (imagepixel.cpp:217-222)."""
    p = params_1spp()
    back = one_grid([1.25, 8.75], [1.25, 6.75], z=9.0, ci=(0.0, 0.5, 1.0))                       # opaque
    mid = one_grid([1.25, 8.75], [1.25, 6.75], z=6.0, ci=(0.125, 0.125, 0.0), oi=(0.25, 0.25, 0.25))
    front = one_grid([1.25, 8.75], [1.25, 6.75], z=3.0, ci=(0.5, 0.0, 0.25), oi=(0.5, 0.5, 0.5))
    f32 = np.float32
    c = np.zeros(3, f32)
    a = np.zeros(3, f32)
    for ci, oi in [((0.0, 0.5, 1.0), (1, 1, 1)), ((0.125, 0.125, 0.0), (0.25,) * 3), ((0.5, 0.0, 0.25), (0.5,) * 3)]:
        ci, oi = np.asarray(ci, f32), np.asarray(oi, f32)
        c = c * (f32(1) - np.clip(oi, 0, 1)) + ci
        a = (f32(1) - a) * oi + a
    for order in ([back, mid, front], [front, back, mid], [mid, front, back]):
        ch, _, _ = orc.render(p, scenes.concat(order), 1)
        assert np.all(ch[2:6, 2:8, :3] == c) and np.all(ch[2:6, 2:8, 3:6] == a)
        assert np.allclose(ch[2:6, 2:8, abi.CH_Z], 9.0, rtol=3e-7)           # depth of the nearest surface with Oi >= zthreshold
        assert np.all(ch[2:6, 2:8, abi.CH_COVERAGE] == 1.0)


def test_constant_colour_survives_any_filter():
    """A full-screen constant surface: the normalised gather acc/gTot returns the colour
    (to rounding) for every filter and width, coverage 1 everywhere."""
    for name, w in [("box", 1.0), ("triangle", 2.0), ("gaussian", 2.0), ("catmull-rom", 3.0), ("sinc", 4.0), ("gaussian", 6.0)]:
        p = default_params(resolution=(24, 20), samples=(3, 3), filter=(name, w, w), displays=[("rgba", 1, 255.0, 0.0, 255.0, 0.5)])
        ch, disp, _ = orc.render(p, one_grid([-8, 32], [-8, 28], ci=(0.2, 0.4, 0.8)), 2)
        assert np.allclose(ch[..., :3], [0.2, 0.4, 0.8], rtol=2e-5, atol=0)
        assert np.all(ch[..., abi.CH_COVERAGE] == 1.0) and np.allclose(ch[..., abi.CH_ALPHA], 1.0, rtol=2e-6)
        assert np.all(np.abs(disp[0].astype(int) - [51, 102, 204, 255]) <= 1)


def test_power_of_two_linearity_and_thread_invariance():
    p, g = scenes.config1(scale=0.15)
    ch1, d1, s1 = orc.render(p, g, 1)
    ch4, d4, s4 = orc.render(p, g, 4)
    assert np.array_equal(ch1.view(np.uint32), ch4.view(np.uint32)) and np.array_equal(d1[0], d4[0])
    assert s1["spl_hits"] == s4["spl_hits"] and s1["n_micropolygons"] == s4["n_micropolygons"]
    g2 = GridArrays(cu=g.cu, cv=g.cv, flags=g.flags, P=g.P, Ci=g.Ci * np.float32(0.5), Oi=g.Oi)
    ch2, _, _ = orc.render(p, g2, 4)
    assert np.array_equal((ch2[..., :3] * np.float32(2)).view(np.uint32), ch1[..., :3].view(np.uint32))
    assert np.array_equal(ch2[..., 3:].view(np.uint32), ch1[..., 3:].view(np.uint32))


def test_culled_and_trimmed_micropolygons_are_skipped():
    p = params_1spp()
    g = one_grid([1.25, 3.25, 5.25, 7.25], [1.25, 6.75], ci=(1, 1, 1))
    culled = np.zeros(8, np.uint8)
    culled[1] = 1                      # micropolygon starting at vertex 1 (x in [3.25,5.25])
    g.culled = culled
    ch, _, _ = orc.render(p, g, 1)
    cov = ch[3, :, abi.CH_COVERAGE]
    assert list(cov) == [0, 1, 1, 0, 0, 1, 1, 0, 0, 0]


def test_exposure_and_quantize():
    """ExposeBucket: pow(Ci*gain, 1/gamma) (bucketprocessor.cpp:766-806); quantise in double with
    lround(x) = lfloor(x-0.5)+1 then clamp (ddmanager.cpp:1065-1070)."""
    p = params_1spp(exposure=(2.0, 2.0), displays=[("rgba", 1, 255.0, 0.0, 255.0, 0.0), ("rgb", 1, 1000.0, 10.0, 600.0, 0.0)])
    ch, disp, _ = orc.render(p, one_grid([-2, 12], [-2, 10], ci=(0.02, 0.125, 0.5)), 1)
    f32 = np.float32
    want = np.array([f32(np.float64(f32(v) * f32(2)) ** np.float64(f32(1) / f32(2))) for v in (0.02, 0.125, 0.5)], f32)
    assert np.allclose(ch[4, 4, :3], want, rtol=2e-7)
    q = [int(np.floor(255.0 * np.float64(v) - 0.5)) + 1 for v in ch[4, 4, :3]]
    assert list(disp[0][4, 4]) == q + [255]
    assert disp[1].dtype == np.uint16
    q2 = [min(600, max(10, int(np.floor(1000.0 * np.float64(v) - 0.5)) + 1)) for v in ch[4, 4, :3]]
    assert list(disp[1][4, 4]) == q2


def _digest(ch, disp):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(ch).tobytes())
    for d in disp:
        h.update(np.ascontiguousarray(d).tobytes())
    return h.hexdigest()


FRAMES = {
    "config1_s0.15": lambda: scenes.config1(scale=0.15),
    "config2_s0.05": lambda: scenes.config2(scale=0.05),
    "config3_s0.04": lambda: scenes.config3(scale=0.04, motion_px=6.0),
    "config4_s0.015": lambda: scenes.config4(scale=0.015),
    "sweep_sinc5_s0.04": lambda: scenes.config2(scale=0.04, filter=("sinc", 5.0, 5.0), samples=(4, 4)),
    "depthfilter_midpoint_rgbaz_deep": lambda: _with(scenes.config4(scale=0.015), depth_filter=abi.DEPTHFILTER_MIDPOINT, display_mode=7),
    "depthfilter_max_rgbaz_deep": lambda: _with(scenes.config4(scale=0.015), depth_filter=abi.DEPTHFILTER_MAX, display_mode=7),
    "depthfilter_average_rgbaz_mbdof": lambda: _with(scenes.config3(scale=0.04, motion_px=6.0), depth_filter=abi.DEPTHFILTER_AVERAGE, display_mode=7),
    "motion_6keys_dof": lambda: scenes.multikey(scale=0.04, nkeys=6, dof=True),
    "camera_space_static": lambda: scenes.to_camera_space(*scenes.config1(scale=0.15)),
    "deep_stack_150": lambda: scenes.deep_stack(),
}


def _with(scene, **kw):
    p, g = scene
    for k, v in kw.items():
        setattr(p, k, v)
    return p, g


@pytest.mark.parametrize("name", sorted(FRAMES))
def test_oracle_frame_checksums(name):
    """Regression pin of the oracle itself (written by `python tests/test_oracle_render.py`, which refuses to write a
    frame the reference's own hider does not reproduce bit for bit): the oracle is the yardstick of every GPU parity
    test, so it must not drift silently -- also on machines where /root/reference and oracle/_ref are absent."""
    want = json.load(open(GOLDEN))
    p, g = FRAMES[name]()
    ch, disp, st = orc.render(p, g, 4)
    assert _digest(ch, disp) == want[name]["sha256"], (name, st)
    assert st["spl_hits"] == want[name]["spl_hits"]


if __name__ == "__main__":
    out = {}
    for name, fn in sorted(FRAMES.items()):
        p, g = fn()
        ch, disp, st = orc.render(p, g, 4)
        # golden = output of the reference itself: only frames aqsis' own hider reproduces exactly are written
        ch_r, disp_r, _ = orc.render_reference(p, g)
        ys, xs = slice(p.crop_ymin, p.crop_ymax), slice(p.crop_xmin, p.crop_xmax)
        assert np.array_equal(ch[ys, xs].view(np.uint32), ch_r[ys, xs].view(np.uint32)), name
        assert all(np.array_equal(a[ys, xs], b[ys, xs]) for a, b in zip(disp, disp_r)), name
        out[name] = {"sha256": _digest(ch, disp), "spl_hits": int(st["spl_hits"]), "xres": p.xres, "yres": p.yres,
                     "micropolygons": g.n_micropolygons, "equals_reference_hider": True}
    json.dump(out, open(GOLDEN, "w"), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1))
