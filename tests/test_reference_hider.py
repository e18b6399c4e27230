"""The oracle restatement against the REFERENCE'S OWN HIDER, bit for bit (CPU only).

oracle/_ref/libaqsis_refhider.so is aqsis' libs/core/{imagebuffer,bucketprocessor,imagepixel,micropolygon,
occlusion,bucket,bound,optioncache,options,...}.cpp compiled in place and driven by oracle/ref_hider.cpp
(what that driver supplies itself is listed in its header).  Every frame below must come out IDENTICAL
from the reference code and from oracle/oracle_hider.cpp: float channel buffer bit for bit, quantised
display bytes exactly -- inside the crop window (outside it the reference's bucket pixels are filtered
from stale pixel state of earlier buckets, bucketprocessor.cpp:168-178 only clears the sample region).
This is what pins the oracle: the GPU parity tests then compare the CUDA path with the oracle
(everywhere) and with this library directly (where it travelled to the GPU box).
"""
import ctypes as C

import numpy as np
import pytest

import orc
from aqsis_b200 import abi, default_params, lib, scenes

pytestmark = pytest.mark.skipif(orc.refhider() is None, reason="oracle/_ref/libaqsis_refhider.so not built (reference tree absent)")


def _dof(p):
    lib().aqh_frame_params_set_dof(C.byref(p), 2.8, 0.05, 20.0, 60.0, 60.0)
    return p


def _crop(p):
    p.crop_xmin, p.crop_xmax, p.crop_ymin, p.crop_ymax = 7, p.xres - 5, 3, p.yres - 9
    return p


def _set(p, **kw):
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def _mod(scene, fn):
    p, g = scene
    return fn(p), g


CASES = {
    "config1": lambda: scenes.config1(scale=0.3),
    "config2": lambda: scenes.config2(scale=0.08),
    "config3-mb-dof": lambda: scenes.config3(scale=0.05, motion_px=6.0),
    "motion-only": lambda: _mod(scenes.config3(scale=0.05, motion_px=8.0), lambda p: _set(p, use_dof=0)),
    "motion-long": lambda: _mod(scenes.config3(scale=0.04, motion_px=40.0), lambda p: _set(p, use_dof=0)),
    "dof-only": lambda: _mod(scenes.config2(scale=0.05), _dof),
    "config4-deep": lambda: scenes.config4(scale=0.02),
    "crop": lambda: _mod(scenes.config1(scale=0.15), _crop),
    "crop-mbdof": lambda: _mod(scenes.config3(scale=0.04, motion_px=6.0), _crop),
    "jitter0": lambda: _mod(scenes.config1(scale=0.15), lambda p: _set(p, jitter=0)),
    "jitter0-mbdof": lambda: _mod(scenes.config3(scale=0.04, motion_px=6.0), lambda p: _set(p, jitter=0)),
    "exposure": lambda: _mod(scenes.config1(scale=0.15), lambda p: _set(p, exposure_gain=1.5, exposure_gamma=2.2)),
    "bucket-8x12": lambda: _mod(scenes.config1(scale=0.15), lambda p: _set(p, bucket_xsize=8, bucket_ysize=12)),
    "predraws": lambda: _mod(scenes.config1(scale=0.12), lambda p: _set(p, rng_predraws=1234, rng_seed=77)),
    "shutter-0.2-0.7": lambda: _mod(scenes.config3(scale=0.04, motion_px=6.0), lambda p: _set(p, shutter_open=0.2, shutter_close=0.7, use_dof=0)),
}
# depth filters (imagepixel.cpp:264-300, 319-327; StoreSample's midpoint rule bucketprocessor.cpp:1502-1529) with and
# without a z display (DMode_Z turns max/average into "nothing is cullable", bucketprocessor.cpp:1074-1079)
for _df, _dn in [(abi.DEPTHFILTER_MIDPOINT, "midpoint"), (abi.DEPTHFILTER_MAX, "max"), (abi.DEPTHFILTER_AVERAGE, "average")]:
    for _dm, _mn in [(abi.DMODE_RGB | abi.DMODE_A, "rgba"), (abi.DMODE_RGB | abi.DMODE_A | abi.DMODE_Z, "rgbaz")]:
        CASES[f"depthfilter-{_dn}-{_mn}-static"] = (lambda df=_df, dm=_dm: _mod(scenes.config1(scale=0.15), lambda p: _set(p, depth_filter=df, display_mode=dm)))
        CASES[f"depthfilter-{_dn}-{_mn}-deep"] = (lambda df=_df, dm=_dm: _mod(scenes.config4(scale=0.015), lambda p: _set(p, depth_filter=df, display_mode=dm)))
    CASES[f"depthfilter-{_dn}-rgbaz-mbdof"] = (lambda df=_df: _mod(scenes.config3(scale=0.04, motion_px=6.0),
                                                                   lambda p: _set(p, depth_filter=df, display_mode=abi.DMODE_RGB | abi.DMODE_A | abi.DMODE_Z)))
# more than two motion keys, non-uniform key times, curved paths (AppendKey / BuildBoundList over several keys)
for _nk in (3, 6):
    CASES[f"motion-{_nk}keys-dof"] = (lambda nk=_nk: scenes.multikey(scale=0.04, nkeys=nk, dof=True))
    CASES[f"motion-{_nk}keys"] = (lambda nk=_nk: scenes.multikey(scale=0.04, nkeys=nk, dof=False))
CASES["motion-5keys-subshutter"] = lambda: scenes.multikey(scale=0.04, nkeys=5, dof=False, shutter=(0.25, 0.75))
# Project_points on the hider's side of the seam (camera-space grids + matCameraToRaster, micropolygon.cpp:723-731;
# the reference wrapper multiplies with aqsis' own CqMatrix), and bins longer than one sorted run
CASES["camera-space-static"] = lambda: scenes.to_camera_space(*scenes.config1(scale=0.15))
CASES["camera-space-mbdof"] = lambda: scenes.to_camera_space(*scenes.config3(scale=0.04, motion_px=6.0))
CASES["deep-stack-150"] = lambda: scenes.deep_stack()
for _name, _w in [("box", 1.0), ("triangle", 2.0), ("gaussian", 3.0), ("catmull-rom", 4.0), ("sinc", 5.0), ("sinc", 6.0),
                  ("gaussian", 2.5), ("mitchell", 4.0), ("disk", 3.0), ("bessel", 4.0)]:
    CASES[f"filter-{_name}-{_w}"] = (lambda n=_name, w=_w: scenes.config2(scale=0.04, filter=(n, w, w), samples=(4, 4)))
for _s in [(1, 1), (2, 3), (5, 5), (16, 16)]:
    CASES[f"samples-{_s[0]}x{_s[1]}"] = (lambda s=_s: scenes.config2(scale=0.03, samples=s, filter=("gaussian", 2.0, 2.0)))


# arbitrary output variables through aqsis' own StoreExtraData / FilterBucket (bucketprocessor.cpp:1573-1643, 620-653)
CASES["aov-static"] = lambda: scenes.with_aovs(*scenes.config1(scale=0.2))
CASES["aov-deep"] = lambda: scenes.with_aovs(*scenes.config4(scale=0.02))
CASES["aov-mbdof"] = lambda: scenes.with_aovs(*scenes.config3(scale=0.04, motion_px=6.0))
CASES["aov-matrix"] = lambda: scenes.with_aovs(*scenes.config1(scale=0.12), aovs=(("N", 3), ("M", 16), ("s", 1), ("t", 1)))
# CSG solids through aqsis' own CqCSGTreeNode objects (csgtree.cpp:144-351)
for _op in ("difference", "union", "intersection"):
    for _nested in (False, True):
        CASES[f"csg-{_op}{'-nested' if _nested else ''}"] = (lambda op=_op, nested=_nested: scenes.csg_scene(op=op, nested=nested))
CASES["csg-midpoint-rgbaz"] = lambda: _mod(scenes.csg_scene(op="difference", nested=True),
                                            lambda p: _set(p, depth_filter=abi.DEPTHFILTER_MIDPOINT, display_mode=abi.DMODE_RGB | abi.DMODE_A | abi.DMODE_Z))
CASES["csg-average-rgbaz"] = lambda: _mod(scenes.csg_scene(op="difference", nested=True),
                                           lambda p: _set(p, depth_filter=abi.DEPTHFILTER_AVERAGE, display_mode=abi.DMODE_RGB | abi.DMODE_A | abi.DMODE_Z))
# the backface / transparency culls of CqMicroPolyGrid::Shade (restated with aqsis' vector and colour classes in the wrapper)
CASES["culls"] = lambda: scenes.cull_scene()
# motion blur (and depth of field) over semi-transparent layers: RenderMPG_MBOrDof feeding StoreSample's transparent branch
CASES["motion-transparent"] = lambda: scenes.layered_motion(dof=False)
CASES["motion-dof-transparent"] = lambda: scenes.layered_motion(dof=True)


def assert_identical(p, ch_r, d_r, ch_o, d_o, what):
    ys, xs = slice(p.crop_ymin, p.crop_ymax), slice(p.crop_xmin, p.crop_xmax)
    a, b = ch_r[ys, xs].view(np.uint32), ch_o[ys, xs].view(np.uint32)
    assert np.array_equal(a, b), (what, float((a == b).mean()))
    for x, y in zip(d_r, d_o):
        assert np.array_equal(x[ys, xs], y[ys, xs]), what


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_equals_reference_hider(name):
    p, g = CASES[name]()
    ch_r, d_r, st_r = orc.render_reference(p, g)
    ch_o, d_o, st_o = orc.render(p, g, 1)
    assert_identical(p, ch_r, d_r, ch_o, d_o, name)
    # the multi-threaded oracle (the CPU baseline of bench.py) is the same image again
    ch_m, d_m, _ = orc.render(p, g, 4)
    assert_identical(p, ch_r, d_r, ch_m, d_m, name + " (4 threads)")
    assert st_r["n_micropolygons"] >= st_o["n_micropolygons"] > 0


def test_reference_hider_known_answers():
    """The hand-derived scenes of test_oracle_render.py, on the reference code itself."""
    from test_oracle_render import one_grid, params_1spp
    p = params_1spp()
    ch, disp, _ = orc.render_reference(p, one_grid([2.25, 6.75], [1.25, 4.75], z=7.0))
    want = np.zeros((8, 10), bool)
    want[1:5, 2:7] = True
    assert np.array_equal(ch[..., abi.CH_COVERAGE] == 1.0, want)
    assert np.all(ch[want][:, :3] == np.float32([0.25, 0.5, 0.75])) and np.all(disp[0][want] == [64, 128, 191, 255])
    for xs, ys in [([2.5, 4.5, 6.5], [1.25, 4.75]), ([2.25, 6.75], [1.5, 3.5, 5.5]), ([2.5, 4.5, 6.5], [1.5, 3.5, 5.5])]:
        g = one_grid(xs, ys, ci=(0.5, 0.5, 0.5), oi=(0.5, 0.5, 0.5))
        ch_r, d_r, _ = orc.render_reference(p, g)
        ch_o, d_o, _ = orc.render(p, g, 1)
        assert_identical(p, ch_r, d_r, ch_o, d_o, "shared edge")
        assert set(np.unique(ch_r[..., abi.CH_OI_R])) <= {0.0, 0.5}
    back = one_grid([1.25, 8.75], [1.25, 6.75], z=9.0, ci=(0.0, 0.5, 1.0))
    mid = one_grid([1.25, 8.75], [1.25, 6.75], z=6.0, ci=(0.125, 0.125, 0.0), oi=(0.25, 0.25, 0.25))
    front = one_grid([1.25, 8.75], [1.25, 6.75], z=3.0, ci=(0.5, 0.0, 0.25), oi=(0.5, 0.5, 0.5))
    g = scenes.concat([mid, front, back])
    ch_r, d_r, _ = orc.render_reference(p, g)
    ch_o, d_o, _ = orc.render(p, g, 1)
    assert_identical(p, ch_r, d_r, ch_o, d_o, "layers")


def test_special_grids_match_reference():
    """Triangular grids (split line), matte objects, level-of-detail windows, culled micropolygons, constant shading."""
    for flags, lod in [(abi.GRID_SMOOTH | abi.GRID_TRIANGULAR, None), (abi.GRID_SMOOTH | abi.GRID_MATTE, None),
                       (abi.GRID_SMOOTH | abi.GRID_MATTE_ALPHA, None), (0, None), (abi.GRID_SMOOTH, (0.25, 0.75))]:
        for make in (lambda: scenes.config1(scale=0.15), lambda: scenes.config3(scale=0.04, motion_px=5.0), lambda: scenes.config4(scale=0.015)):
            p, g = make()
            g.flags = g.flags.copy()
            g.flags[::2] = flags
            if lod:
                lb = -np.ones((g.n_grids, 2), np.float32)
                lb[::3] = lod
                g.lod_bounds = lb.ravel()
            rng = np.random.default_rng(5)
            g.culled = (rng.uniform(size=g.n_verts) < 0.05).astype(np.uint8)
            ch_r, d_r, _ = orc.render_reference(p, g)
            ch_o, d_o, _ = orc.render(p, g, 2)
            assert_identical(p, ch_r, d_r, ch_o, d_o, (flags, lod))


@pytest.mark.parametrize("kw", [dict(), dict(motion=True), dict(dof=True), dict(motion=True, dof=True), dict(scale=0.25, outside_every=2)],
                         ids=["static", "motion", "dof", "motion+dof", "larger"])
def test_trim_curves_oracle_equals_reference(kw):
    """Trimmed surfaces: micropolygons trimmed away entirely are dropped while busting, the hits of the ones a trim curve
    crosses are tested against the curves (micropolygon.cpp:784-835, 1594-1628; moving micropolygons are not tested per hit,
    :1877-1884).  In the reference run the surface queries are CqTrimLoopArray::TrimPoint / LineIntersects of the
    reference's own geometry/trimcurve.cpp, asked by the reference's own CqMicroPolygon::Sample."""
    p, g = scenes.trim_scene(**kw)
    ch_r, d_r, _ = orc.render_reference(p, g)
    ch_o, d_o, st = orc.render(p, g, 2)
    assert_identical(p, ch_r, d_r, ch_o, d_o, ("trim", kw))
    # and the trimming does something: the untrimmed frame differs
    g.trim_set, p._trim = None, None
    ch_u, _, st_u = orc.render(p, g, 2)
    assert st_u["n_micropolygons"] > st["n_micropolygons"] and not np.array_equal(ch_u, ch_o)


def test_trim_leaves_match_reference():
    """CqTrimLoopArray::TrimPoint / LineIntersects (geometry/trimcurve.cpp:145-242): the oracle's restatement against the
    reference's own functions on random points and segments, including points exactly on loop vertices' y."""
    p, g = scenes.trim_scene()
    n, a, b, c = orc._trim_arrays(p)
    L, R = orc.lib(), orc.refhider()
    L.orc_set_trim_loops(n, a.ctypes.data, b.ctypes.data, c.ctypes.data)
    R.ref_set_trim_loops(n, a.ctypes.data, b.ctypes.data, c.ctypes.data)
    rng = np.random.default_rng(7)
    pts = rng.uniform(-0.1, 1.1, (4000, 2)).astype(np.float32)
    pts[::7, 1] = c.reshape(-1, 2)[rng.integers(0, len(c) // 2, len(pts[::7])), 1]      # on a vertex's y
    inside = 0
    for s in range(1, n + 1):
        for x, y in pts:
            o, r = L.orc_trim_point(s, float(x), float(y)), R.ref_trim_point(s, float(x), float(y))
            assert o == r, (s, x, y)
            inside += 1 - o
        for (x1, y1), (x2, y2) in zip(pts[:600], pts[600:1200] * 0.2 + pts[:600] * 0.8):
            assert L.orc_trim_line(s, float(x1), float(y1), float(x2), float(y2)) == R.ref_trim_line(s, float(x1), float(y1), float(x2), float(y2))
    assert inside > 1000
    L.orc_set_trim_loops(0, None, None, None)
    R.ref_set_trim_loops(0, None, None, None)


def test_pdiff_reports_zero_differences_between_oracle_and_reference():
    """BASELINE.json's third criterion, with the reference's own tool (thirdparty/pdiff compiled in place): the
    quantised images are binary identical; and the tool does see a real difference (one bucket of pixels shifted)."""
    if orc.pdiff_lib() is None:
        pytest.skip("oracle/_ref/libaqsis_pdiff.so not built")
    p, g = scenes.config2(scale=0.08)
    _, d_r, _ = orc.render_reference(p, g)
    _, d_o, _ = orc.render(p, g, 4)
    ok, failed, same = orc.pdiff(d_r[0], d_o[0])
    assert ok and failed == 0 and same
    broken = d_o[0].copy()
    broken[16:48, 16:48] = np.roll(broken[16:48, 16:48], 5, axis=1)
    broken[60:70, 60:70] = 255 - broken[60:70, 60:70]
    ok, failed, same = orc.pdiff(d_r[0], broken)
    assert not ok and failed > 0 and not same


@pytest.mark.parametrize("name,width", scenes.config5_filters())
def test_config5_filter_sweep_oracle_equals_reference(name, width):
    """BASELINE.json config 5: box, triangle, gaussian, catmull-rom, sinc at widths 1-6 -- every combination of the
    sweep, on a small copy of the 1080p scene, bit for bit against aqsis' own FilterBucket."""
    p, g = scenes.config2(scale=0.03, filter=(name, width, width), samples=(4, 4))
    ch_r, d_r, _ = orc.render_reference(p, g)
    ch_o, d_o, _ = orc.render(p, g, 2)
    assert_identical(p, ch_r, d_r, ch_o, d_o, (name, width))
