"""GPU parity of the "next" rows of SURVEY.md 8(f): arbitrary output variables, CSG solids, point micropolygons, the
backface / transparency culls, the imager hook, scan-line display order and the incremental occlusion flush.

The oracle is pinned to aqsis' own hider for AOVs (real StoreExtraData + FilterBucket), CSG (real CqCSGTreeNode objects)
and the culls in tests/test_reference_hider.py; here the CUDA path must equal the oracle bit for bit, and the reference
hider itself where it travelled."""
import ctypes as C

import numpy as np
import pytest

import orc
import parity_util as pu
from aqsis_b200 import abi, scenes, lib

pytestmark = pytest.mark.gpu


def same_bits(p, ch_g, d_g, ch_r, d_r, what):
    ys, xs = slice(p.crop_ymin, p.crop_ymax), slice(p.crop_xmin, p.crop_xmax)
    a, b = ch_g[ys, xs].view(np.uint32), ch_r[ys, xs].view(np.uint32)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert np.array_equal(a, b), (what, "float channel buffer differs", float((a == b).mean()),
                                  [float((a[..., k] == b[..., k]).mean()) for k in range(a.shape[-1])])
    for x, y in zip(d_g, d_r):
        assert np.array_equal(x[ys, xs].view(np.uint8), y[ys, xs].view(np.uint8)), (what, "display bytes differ")


def check(h, p, g, what, reference=True):
    ch_g, d_g, st = pu.run_product(h, p, g)
    ch_o, d_o, _ = orc.render(p, g, 8)
    same_bits(p, ch_g, d_g, ch_o, d_o, what + " vs oracle")
    if reference and orc.refhider() is not None:
        ch_r, d_r, _ = orc.render_reference(p, g)
        same_bits(p, ch_g, d_g, ch_r, d_r, what + " vs aqsis' hider")
    return ch_g, d_g, st


@pytest.mark.parametrize("kind", ["static", "deep", "mbdof", "banded"])
def test_arbitrary_output_variables(gpu_hider, kind):
    """AOV floats travel with the nearest hit (StoreExtraData, bucketprocessor.cpp:1573-1643; Combine keeps the nearest
    entry's extras, imagepixel.cpp:249-251), are filtered like colour (:620-653) and shown by a float display."""
    make = {"static": lambda: scenes.config1(scale=0.2), "deep": lambda: scenes.config4(scale=0.02),
            "mbdof": lambda: scenes.config3(scale=0.04, motion_px=6.0), "banded": lambda: scenes.config4(scale=0.04)}[kind]
    aovs = (("N", 3), ("_depthcue", 1), ("_albedo", 3)) if kind != "banded" else (("N", 3), ("M", 16), ("s", 1), ("t", 1))
    p, g = scenes.with_aovs(*make(), aovs=aovs)
    if kind == "banded":
        p.plane_budget_mb = 24
    ch, d, st = check(gpu_hider, p, g, f"AOV {kind}")
    assert ch.shape[-1] == 9 + sum(n for _, n in aovs) and np.abs(ch[..., 9:]).max() > 0.1
    assert d[1].dtype == np.float32
    if kind == "banded":
        assert st["n_bands"] > 1
    # a frame that declares AOVs but whose grids carry none: the extra channels read zero
    g.aov = None
    ch0, _, _ = pu.run_product(gpu_hider, p, g)
    assert np.all(ch0[..., 9:] == 0) and np.array_equal(ch0[..., :9].view(np.uint32), ch[..., :9].view(np.uint32))


@pytest.mark.parametrize("kw", [dict(), dict(motion=True), dict(dof=True), dict(motion=True, dof=True), dict(scale=0.25, outside_every=2)],
                         ids=["static", "motion", "dof", "motion+dof", "larger"])
def test_trim_curves(gpu_hider, kw):
    """Trimmed NURBS surfaces (aqh_set_trim_loops, per-grid trim set + surface parameters): k_project drops the
    micropolygons that are trimmed away and marks the ones a curve crosses, k_hide tests their hits
    (micropolygon.cpp:784-835, 1594-1628) -- against the oracle and the reference's own hider with its own trimcurve.cpp."""
    p, g = scenes.trim_scene(**kw)
    ch, d, st = check(gpu_hider, p, g, f"trim {kw}")
    g.trim_set, p._trim = None, None
    ch_u, _, st_u = pu.run_product(gpu_hider, p, g)
    assert st_u["n_micropolygons"] > st["n_micropolygons"] and not np.array_equal(ch_u, ch)


def test_transparent_hits_behind_a_later_opaque_surface_stay_within_tolerance(gpu_hider):
    """KNOWN DEVIATION: the reference keeps a transparent hit that lies BEHIND an opaque surface when the
    opaque surface is submitted later (StoreSample only culls against the occlusion depth of the moment,
    bucketprocessor.cpp:1475-1480); back-to-front compositing then covers it with the opaque hit, which changes the last bit
    of the result whenever the occluder's interpolated opacity is not exactly 1.  The device hides all opaque surfaces
    first and never stores such hits.  Mixed, randomly ordered transparent and opaque grids: equal to the oracle within
    a few ulp (far inside the north star's 1e-4), but not bit for bit."""
    rng = np.random.default_rng(5)
    p = scenes.default_params(resolution=(76, 57), samples=(4, 4), filter=("gaussian", 2.0, 2.0), displays=[scenes._RGBA8])
    G = 40
    centers = np.stack([rng.uniform(0, 76, G), rng.uniform(0, 57, G)], axis=1).astype(np.float32)
    P, Ci, Oi = scenes._grids(rng, centers, 18.0, 12, 12, 5.0, 60.0, opacity=np.where(rng.uniform(size=G) < 0.3, 0.6, 1.0).astype(np.float32))
    g = scenes._pack(P, Ci, Oi, 12, 12)
    ch_g, d_g, st = pu.run_product(gpu_hider, p, g)
    ch_o, d_o, so = orc.render(p, g, 4)
    assert so["n_deep_hits"] > st["n_deep_hits"] > 0            # the oracle (like the reference) keeps the hidden hits
    fin = np.isfinite(ch_o) & (np.abs(ch_o) < 1e30)
    rel = np.abs(ch_g[fin] - ch_o[fin]) / np.maximum(np.abs(ch_o[fin]), 1e-3)
    assert rel.max() < 1e-6
    assert np.abs(d_g[0].astype(int) - d_o[0].astype(int)).max() <= 1


def test_trim_through_add_grid(gpu_hider):
    p, g = scenes.trim_scene()
    ch_a, d_a, _ = pu.run_product(gpu_hider, p, g, use_block=True)
    h = gpu_hider
    h.begin_frame(p)
    nv = 13 * 13
    for i in range(g.n_grids):
        h.add_grid(g.P[i * nv:(i + 1) * nv], 12, 12, Ci=g.Ci[i * nv:(i + 1) * nv], Oi=g.Oi[i * nv:(i + 1) * nv],
                   flags=int(g.flags[i]), trim_set=int(g.trim_set[i]), trim_uv=g.trim_uv[i * nv:(i + 1) * nv])
    ch_b, d_b = h.end_frame()
    assert np.array_equal(ch_a.view(np.uint32), ch_b.view(np.uint32)) and np.array_equal(d_a[0], d_b[0])
    # a grid that names a trim set the frame does not have is rejected, and the frame goes on
    h.begin_frame(p)
    with pytest.raises(Exception):
        h.add_grid(g.P[:nv], 12, 12, Ci=g.Ci[:nv], Oi=g.Oi[:nv], trim_set=99, trim_uv=g.trim_uv[:nv])
    h.end_frame()


def test_aov_through_add_grid(gpu_hider):
    p, g = scenes.with_aovs(*scenes.config1(scale=0.12))
    ch_a, d_a, _ = pu.run_product(gpu_hider, p, g, use_block=True)
    h = gpu_hider
    h.begin_frame(p)
    nv, A = 81, p.aov_floats
    for i in range(g.n_grids):
        h.add_grid(g.P[i * nv:(i + 1) * nv], 8, 8, Ci=g.Ci[i * nv:(i + 1) * nv], Oi=g.Oi[i * nv:(i + 1) * nv],
                   flags=int(g.flags[i]), aov=g.aov[i * nv:(i + 1) * nv])
    ch_b, d_b = h.end_frame()
    assert np.array_equal(ch_a.view(np.uint32), ch_b.view(np.uint32)) and np.array_equal(d_a[1], d_b[1])


@pytest.mark.parametrize("op", ["difference", "union", "intersection"])
@pytest.mark.parametrize("nested", [False, True])
def test_csg_solids(gpu_hider, op, nested):
    """CSG sample resolve (imagepixel.cpp:166-189 -> CqCSGTreeNode::ProcessTree, csgtree.cpp:144-351)."""
    p, g = scenes.csg_scene(op=op, nested=nested)
    ch, d, st = check(gpu_hider, p, g, f"CSG {op} nested={nested}")
    assert st["n_deep_hits"] > 0
    # the tree matters: the same grids without CSG give another image
    g2 = scenes.csg_scene(op=op, nested=nested)[1]
    g2.flags = (g2.flags & ~np.uint32(abi.GRID_USES_CSG)).astype(np.uint32)
    p2 = scenes.csg_scene(op=op, nested=nested)[0]
    p2._csg = None
    ch2, _, _ = pu.run_product(gpu_hider, p2, g2)
    assert not np.array_equal(ch, ch2)


def test_csg_with_depth_filters_and_dof(gpu_hider):
    for df in (abi.DEPTHFILTER_MIDPOINT, abi.DEPTHFILTER_AVERAGE):
        p, g = scenes.csg_scene(op="difference", nested=True)
        p.depth_filter, p.display_mode = df, abi.DMODE_RGB | abi.DMODE_A | abi.DMODE_Z
        check(gpu_hider, p, g, f"CSG depth filter {df}")
    p, g = scenes.csg_scene(op="difference")
    s_ = 0.5 * p.yres / np.tan(np.radians(20.0))
    lib().aqh_frame_params_set_dof(C.byref(p), 2.8, 0.05, 14.0, s_, s_)
    check(gpu_hider, p, g, "CSG under depth of field")


@pytest.mark.parametrize("dof", [False, True])
def test_point_micropolygons(gpu_hider, dof):
    """CqMicroPolygonPoints (geometry/points.cpp:653-700): discs with constant shading, opaque and transparent, with
    and without depth of field.  The reference side is the oracle's restatement only (points.cpp needs the whole
    CqSurface family to compile): parity unpinned by reference code for this row."""
    p, g = scenes.points_scene(dof=dof)
    ch, d, st = check(gpu_hider, p, g, f"points dof={dof}", reference=False)
    assert st["n_micropolygons"] > 4000
    # points in a frame with moving grids but no depth of field take the static path
    if not dof:
        pm, gm = scenes.config3(scale=0.04, motion_px=6.0)
        pm.use_dof = 0
        pp, gp = scenes.points_scene(n_points=1500, res=(pm.xres, pm.yres))
        from aqsis_b200.hider import GridArrays
        nb = 12                                                   # the backdrop grids of points_scene (8x8: 81 vertices each)
        pts = GridArrays(cu=gp.cu[nb:], cv=gp.cv[nb:], flags=gp.flags[nb:], P=gp.P[nb * 81:], Ci=gp.Ci[nb * 81:], Oi=gp.Oi[nb * 81:],
                         nkeys=np.ones(gp.n_grids - nb, np.int32), key_times=np.zeros(gp.n_grids - nb, np.float32))
        both = scenes.concat([gm, pts])
        r = np.zeros(both.P.shape[0], np.float32)
        r[gm.P.shape[0]:] = gp.radius[nb * 81:]
        both.radius = r
        check(gpu_hider, pm, both, "points among moving grids", reference=False)


def test_backface_and_transparency_culls(gpu_hider):
    """The culls CqMicroPolyGrid::Shade applies before busting, on the device (micropolygon.cpp:431-474, 493-522).

    Tolerance instead of bit identity for the FLOAT channels of this frame: its partly transparent micropolygons are
    submitted in random depth order, so the reference keeps transparent hits that were stored BEFORE a nearer opaque hit
    arrived (bucketprocessor.cpp:1475-1480 only drops what comes after); they composite to nothing visible (C*(1-O)+Ci with
    O = 1) but perturb the last bit when the interpolated opacity of the opaque hit is 1 - 1 ulp.
    The device resolves after all opaque hits are known and never keeps them.  Stated tolerance: 1e-6 relative on floats
    (north_star allows 1e-4), quantised bytes identical."""
    p, g = scenes.cull_scene()
    ch_g, d_g, st = pu.run_product(gpu_hider, p, g)
    ch_o, d_o, ost = orc.render(p, g, 8)
    res = pu.compare(ch_g, d_g, ch_o, d_o, float_rtol=1e-6, quant_atol=0)
    assert res["float_bit_exact_frac"] > 0.999
    # the same micropolygons survive the culls on both sides (the oracle is pinned to aqsis' own arithmetic for them)
    assert abs(st["n_micropolygons"] - ost["n_micropolygons"]) <= 0.02 * ost["n_micropolygons"]
    assert 0.2 * g.n_micropolygons < st["n_micropolygons"] < 0.7 * g.n_micropolygons
    # an all-opaque variant of the frame (backface cull only) is bit-exact
    g1 = scenes.cull_scene()[1]
    g1.Oi = np.ones_like(g1.Oi)
    check(gpu_hider, p, g1, "backface cull")
    # without the flags nothing is culled
    g2 = scenes.cull_scene()[1]
    g2.flags = (g2.flags & ~np.uint32(abi.GRID_CULL_BACKFACING | abi.GRID_CULL_TRANSPARENT)).astype(np.uint32)
    _, _, st2 = pu.run_product(gpu_hider, p, g2)
    assert st2["n_micropolygons"] > st["n_micropolygons"]


def test_incremental_flush(gpu_hider):
    """aqh_flush hides only what was submitted since the previous flush against occlusion keys kept in HBM, the final
    frame starts from those keys and skips the flushed opaque micropolygons: N flushes cost O(total), the image is that
    of a single-shot frame bit for bit, aqh_can_cull is monotone and equals a single full flush."""
    from aqsis_b200.hider import GridArrays
    h = gpu_hider

    def parts(g, n):
        nv = (g.cu.astype(np.int64) + 1) * (g.cv + 1)
        nk = g.nkeys.astype(np.int64) if g.nkeys is not None else np.ones_like(nv)
        ps = np.concatenate([[0], np.cumsum(nv * nk)])
        vs = np.concatenate([[0], np.cumsum(nv)])
        ks = np.concatenate([[0], np.cumsum(nk)])
        cuts = [g.n_grids * i // n for i in range(n + 1)]
        out = []
        for a, b in zip(cuts[:-1], cuts[1:]):
            out.append(GridArrays(cu=g.cu[a:b], cv=g.cv[a:b], flags=g.flags[a:b], P=g.P[ps[a]:ps[b]], Ci=g.Ci[vs[a]:vs[b]], Oi=g.Oi[vs[a]:vs[b]],
                                  nkeys=None if g.nkeys is None else g.nkeys[a:b],
                                  key_times=None if g.key_times is None else g.key_times[ks[a]:ks[b]]))
        return out

    rng = np.random.default_rng(9)
    for make, nparts in ((lambda: scenes.config2(scale=0.1), 5), (lambda: scenes.config4(scale=0.03), 3),
                         (lambda: scenes.config3(scale=0.05, motion_px=6.0), 3)):
        p, g = make()
        ch_a, d_a, st_a = pu.run_product(h, p, g)                 # single shot
        n = 300
        c = np.stack([rng.uniform(0, p.xres, n), rng.uniform(0, p.yres, n)], 1)
        sz = rng.uniform(0.5, 10, (n, 2))
        z = rng.uniform(2, 110, n)
        bounds = np.concatenate([c - sz / 2, z[:, None], c + sz / 2, (z + 1)[:, None]], 1).astype(np.float32)
        h.begin_frame(p)
        h.add_grid_block(g)
        h.flush()
        full = np.array([h.can_cull(b) for b in bounds])
        h.end_frame()
        h.begin_frame(p)
        prev = np.zeros(n, bool)
        entries = []
        for part in parts(g, nparts):
            h.add_grid_block(part)
            h.flush()
            entries.append(h.stats()["n_bin_entries"])
            now = np.array([h.can_cull(b) for b in bounds])
            assert not np.any(prev & ~now)                         # monotone: what was culled stays culled
            prev = now
        assert np.array_equal(prev, full)                          # exact per-sample keys: same answers as one full flush
        ch_b, d_b = h.end_frame()
        st_b = h.stats()
        assert np.array_equal(ch_a.view(np.uint32), ch_b.view(np.uint32)) and np.array_equal(d_a[0], d_b[0])
        # O(total): every micropolygon is binned by exactly one flush (the opaque ones) or by the final frame (the others)
        assert sum(entries) + st_b["n_bin_entries"] <= st_a["n_bin_entries"]
        # mixed: some grids flushed, some only submitted before the end of the frame; staged one by one
        h.begin_frame(p)
        ps = parts(g, 3)
        h.add_grid_block(ps[0])
        h.flush()
        pu.add_grids_one_by_one(h, ps[1])
        h.flush()
        pu.add_grids_one_by_one(h, ps[2])
        ch_c, d_c = h.end_frame()
        assert np.array_equal(ch_a.view(np.uint32), ch_c.view(np.uint32)) and np.array_equal(d_a[0], d_c[0])


def test_imager_hook_and_exposure_order(gpu_hider):
    """on_imager runs per bucket after filtering and BEFORE exposure and quantisation (bucketprocessor.cpp:712-743, then
    ExposeBucket :766-806): expected = oracle channels without exposure -> imager -> exposure -> quantise, in numpy."""
    p, g = scenes.config1(scale=0.2)
    p.exposure_gain, p.exposure_gamma = 1.3, 2.2
    seen = []

    def imager(x0, x1, y, row):
        seen.append((x0, x1, y))
        ci = row[:, 0:3].copy()
        row[:, 0:3] = ci[:, ::-1] * np.float32(0.5) + np.float32(0.125)        # new Ci
        row[:, 3:6] = np.float32(1.0)                                          # opaque background (the "background" imager)
        row[:, 6] = np.float32(1.0)
    h = gpu_hider
    h.begin_frame(p)
    h.add_grid_block(g)
    ch_g, d_g = h.end_frame(on_imager=imager)
    assert len(seen) == p.yres * ((p.xres + 15) // 16)
    q = type(p).from_buffer_copy(p)
    q.filter_func = p.filter_func
    q.exposure_gain = q.exposure_gamma = 1.0
    ch_o, _, _ = orc.render(q, g, 4)
    want = ch_o.copy()
    want[..., 0:3] = ch_o[..., 2::-1][..., 0:3] * np.float32(0.5) + np.float32(0.125)
    want[..., 3:6] = 1.0
    want[..., 6] = 1.0
    ci = (want[..., 0:3] * np.float32(1.3)).astype(np.float32)
    inv = np.float32(1.0) / np.float32(2.2)
    want[..., 0:3] = np.power(ci.astype(np.float64), np.float64(inv)).astype(np.float32)
    assert np.array_equal(ch_g.view(np.uint32), want.view(np.uint32))
    # quantised bytes from the float image and the replayed dither (ddmanager.cpp:1046-1113)
    sw, sh = C.c_int(), C.c_int()
    lib().aqh_replay_frame_rng(C.byref(p), None, None, None, None, C.byref(sw), C.byref(sh))
    planes = np.zeros(5 * sw.value * sh.value, np.uint8)
    dither = np.zeros((p.yres, p.xres), np.float32)
    lib().aqh_replay_frame_rng(C.byref(p), planes.ctypes.data, dither.ctypes.data, None, None, None, None)
    rgba = want[..., [0, 1, 2, 6]].astype(np.float64)
    v = 0.0 + rgba * 255.0 + 0.5 * dither.astype(np.float64)[..., None]
    xm = v - 0.5
    li = np.trunc(xm)
    li = li - ((xm < 0) & (xm != li))
    qv = np.clip(li + 1, 0, 255).astype(np.uint8)
    assert np.array_equal(d_g[0], qv)


def test_scanline_order_display(gpu_hider):
    """AQH_DISPLAY_SCANLINE_ORDER: whole rows, one call each, as soon as a row of buckets is complete
    (CollapseBucketsToScanlines / SendToDisplay, ddmanager.cpp:1129-1175)."""
    p, g = scenes.config1(scale=0.15)
    p.display[0].flags = abi.DISPLAY_SCANLINE_ORDER
    ch, d, _ = pu.run_product(gpu_hider, p, g)
    calls, img = [], np.zeros_like(d[0])
    h = gpu_hider
    h.begin_frame(p)
    h.add_grid_block(g)

    def on_data(disp, x0, x1, y0, y1, es, data):
        calls.append((x0, x1, y0, y1))
        img[y0:y1, x0:x1] = data.reshape(y1 - y0, x1 - x0, es)
    h.end_frame(on_data=on_data)
    assert calls == [(0, p.xres, y, y + 1) for y in range(p.yres)]
    assert np.array_equal(img, d[0])


@pytest.mark.parametrize("dof", [False, True])
def test_transparent_motion_blur_and_both_motion_kernels(gpu_hider, dof, monkeypatch):
    """Motion blur (and depth of field) over semi-transparent layers: RenderMPG_MBOrDof feeding StoreSample's transparent
    branch and Combine (bucketprocessor.cpp:1221-1469, 1471-1569; imagepixel.cpp:144-262).  Such a frame runs the short
    motion kernel (k_hide<.., PLAIN>: no discs / level of detail / trim / triangular grids / more than two keys); forcing the
    general kernel on the same frame must give the same bits."""
    p, g = scenes.layered_motion(dof=dof)
    ch, d, st = check(gpu_hider, p, g, "transparent motion blur" + (" + depth of field" if dof else ""))
    assert st["n_deep_hits"] > 0
    monkeypatch.setenv("AQH_TUNE", "0,0,0,0,1")
    ch2, d2, st2 = pu.run_product(gpu_hider, p, g)
    monkeypatch.delenv("AQH_TUNE")
    same_bits(p, ch2, d2, ch, d, "general motion kernel vs short motion kernel")
    assert st2["n_deep_hits"] == st["n_deep_hits"]


def test_both_motion_kernels_on_config3(gpu_hider, monkeypatch):
    """The short and the general motion kernel on an opaque config-3 frame (bit for bit, and against the oracle)."""
    p, g = scenes.config3(scale=0.08)
    ch, d, _ = check(gpu_hider, p, g, "config 3 (short motion kernel)", reference=False)
    monkeypatch.setenv("AQH_TUNE", "0,0,0,0,1")
    ch2, d2, _ = pu.run_product(gpu_hider, p, g)
    monkeypatch.delenv("AQH_TUNE")
    same_bits(p, ch2, d2, ch, d, "general motion kernel vs short motion kernel")


@pytest.mark.parametrize("which", ["config1", "config2", "config4"])
def test_short_and_general_static_kernels_agree(gpu_hider, which, monkeypatch):
    """Frames without uncommon features run k_hide<.., PLAIN> (no out-of-line rare-record / CSG calls, no AOV / flush code);
    AQH_TUNE=0,0,0,0,1 forces the general instantiation on the same frame: same bits, and both equal to the oracle."""
    p, g = {"config1": lambda: scenes.config1(scale=0.3), "config2": lambda: scenes.config2(scale=0.08),
            "config4": lambda: scenes.config4(scale=0.03)}[which]()
    ch, d, st = check(gpu_hider, p, g, which + " (short static kernel)", reference=False)
    monkeypatch.setenv("AQH_TUNE", "0,0,0,0,1")
    ch2, d2, st2 = pu.run_product(gpu_hider, p, g)
    monkeypatch.delenv("AQH_TUNE")
    same_bits(p, ch2, d2, ch, d, "general static kernel vs short static kernel")
    assert st2["n_deep_hits"] == st["n_deep_hits"]
