"""Synthetic shaded grids of the shapes BASELINE.json names.

The renderer front end (RIB, dicing, aqsl shaders) is outside the hot path and not buildable
here, so the inputs of the hider -- shaded micropolygon grids (P, Ci, Oi) -- are generated
directly: P already in the hybrid raster-x,y / camera-z space CqMicroPolyGrid::Split produces,
"plastic" emulated by a smooth Lambert-like colour field.  Seed 20261017.
"""
import math

import numpy as np

from . import _abi as abi
from .hider import GridArrays, default_params

SEED = 20261017


def _grids(rng, centers, size_px, cu, cv, zmin, zmax, warp=0.12, noise=0.02, opacity=None, rot=True):
    """G grids of (cu+1)x(cv+1) vertices around `centers`, each about `size_px` pixels wide."""
    G = centers.shape[0]
    u = np.linspace(0.0, 1.0, cu + 1, dtype=np.float32)
    v = np.linspace(0.0, 1.0, cv + 1, dtype=np.float32)
    uu, vv = np.meshgrid(u, v)                       # (cv+1, cu+1): vertex index = iv*(cu+1)+iu
    uu = uu.reshape(1, -1)
    vv = vv.reshape(1, -1)
    sx = (size_px * rng.uniform(0.85, 1.15, (G, 1))).astype(np.float32)
    sy = (size_px * rng.uniform(0.85, 1.15, (G, 1))).astype(np.float32)
    lx = (uu - 0.5) * sx
    ly = (vv - 0.5) * sy
    # bilinear corner warp: makes the micropolygons non-rectangular (second Newton step path)
    w = (rng.uniform(-warp, warp, (G, 4, 2)) * size_px).astype(np.float32)
    b00, b10, b01, b11 = (1 - uu) * (1 - vv), uu * (1 - vv), (1 - uu) * vv, uu * vv
    lx = lx + b00 * w[:, 0:1, 0] + b10 * w[:, 1:2, 0] + b01 * w[:, 2:3, 0] + b11 * w[:, 3:4, 0]
    ly = ly + b00 * w[:, 0:1, 1] + b10 * w[:, 1:2, 1] + b01 * w[:, 2:3, 1] + b11 * w[:, 3:4, 1]
    if rot:
        th = rng.uniform(0, 2 * math.pi, (G, 1)).astype(np.float32)
        c, s = np.cos(th), np.sin(th)
        lx, ly = c * lx - s * ly, s * lx + c * ly
    nv = (cu + 1) * (cv + 1)
    P = np.empty((G, nv, 3), dtype=np.float32)
    P[:, :, 0] = centers[:, 0:1] + lx + rng.normal(0, noise, (G, nv)).astype(np.float32)
    P[:, :, 1] = centers[:, 1:2] + ly + rng.normal(0, noise, (G, nv)).astype(np.float32)
    z0 = rng.uniform(zmin, zmax, (G, 1)).astype(np.float32)
    slope = (rng.uniform(-0.05, 0.05, (G, 2)) * z0).astype(np.float32)
    P[:, :, 2] = z0 + slope[:, 0:1] * uu + slope[:, 1:2] * vv
    # "plastic": base colour times a smooth diffuse term plus a small highlight
    base = rng.uniform(0.15, 1.0, (G, 1, 3)).astype(np.float32)
    ph = rng.uniform(0, 2 * math.pi, (G, 2)).astype(np.float32)
    diff = 0.35 + 0.65 * (0.5 + 0.5 * np.sin(4.0 * uu + ph[:, 0:1]) * np.cos(3.0 * vv + ph[:, 1:2]))
    spec = 0.25 * np.clip(1.0 - 6.0 * ((uu - 0.4) ** 2 + (vv - 0.6) ** 2), 0, 1) ** 4
    Ci = (base * diff[:, :, None] + spec[:, :, None]).astype(np.float32)
    if opacity is None:
        Oi = np.ones((G, nv, 3), dtype=np.float32)
    else:
        Oi = np.broadcast_to(np.asarray(opacity, dtype=np.float32).reshape(-1, 1, 1), (G, nv, 3)).copy()
        Ci = Ci * Oi                                   # premultiplied, as shaders hand Ci over
    return P, Ci, Oi


def _pack(P, Ci, Oi, cu, cv, flags=abi.GRID_SMOOTH, P2=None, key_times=None):
    G = P.shape[0]
    if P2 is not None:
        Pk = np.stack([P, P2], axis=1).reshape(-1, 3)     # per grid: key-major
        nkeys = np.full(G, 2, dtype=np.int32)
        kt = np.tile(np.asarray(key_times, dtype=np.float32), G)
    else:
        Pk = P.reshape(-1, 3)
        nkeys = None
        kt = None
    return GridArrays(cu=np.full(G, cu, dtype=np.int32), cv=np.full(G, cv, dtype=np.int32),
                      flags=np.full(G, flags, dtype=np.uint32), P=np.ascontiguousarray(Pk),
                      Ci=np.ascontiguousarray(Ci.reshape(-1, 3)), Oi=np.ascontiguousarray(Oi.reshape(-1, 3)),
                      nkeys=nkeys, key_times=kt)


def concat(blocks):
    """Concatenate GridArrays (all static or all with nkeys given)."""
    any_keys = any(b.nkeys is not None for b in blocks)
    nk, kt = None, None
    if any_keys:
        nk = np.concatenate([b.nkeys if b.nkeys is not None else np.ones(b.n_grids, np.int32) for b in blocks])
        kt = np.concatenate([b.key_times if b.key_times is not None else np.zeros(b.n_grids, np.float32) for b in blocks])
    return GridArrays(cu=np.concatenate([b.cu for b in blocks]), cv=np.concatenate([b.cv for b in blocks]),
                      flags=np.concatenate([b.flags for b in blocks]), P=np.concatenate([b.P for b in blocks]),
                      Ci=np.concatenate([b.Ci for b in blocks]), Oi=np.concatenate([b.Oi for b in blocks]),
                      nkeys=nk, key_times=kt)


_RGBA8 = ("rgba", 1, 255.0, 0.0, 255.0, 0.5)     # Quantize "rgba" 255 0 255 0.5, file-driver channel order


def config1(scale=1.0, seed=SEED):
    """640x480, PixelSamples 4 4, gaussian 2x2, 10k bilinear patches diced 8x8 (~8x8 px each)."""
    xres, yres = max(16, int(640 * scale)), max(16, int(480 * scale))
    rng = np.random.default_rng(seed)
    G = max(4, int(10000 * scale * scale))
    centers = np.stack([rng.uniform(-4, xres + 4, G), rng.uniform(-4, yres + 4, G)], axis=1).astype(np.float32)
    P, Ci, Oi = _grids(rng, centers, 8.0, 8, 8, 5.0, 50.0)
    params = default_params(resolution=(xres, yres), samples=(4, 4), filter=("gaussian", 2.0, 2.0), displays=[_RGBA8])
    return params, _pack(P, Ci, Oi, 8, 8)


def config2(scale=1.0, seed=SEED + 1, filter=("catmull-rom", 3.0, 3.0), samples=(8, 8)):
    """1920x1080, PixelSamples 8 8, ShadingRate 1, ~20M micropolygons, opaque, catmull-rom 3x3."""
    xres, yres = max(16, int(1920 * scale)), max(16, int(1080 * scale))
    rng = np.random.default_rng(seed)
    G = max(4, int(78125 * scale * scale))
    centers = np.stack([rng.uniform(-8, xres + 8, G), rng.uniform(-8, yres + 8, G)], axis=1).astype(np.float32)
    P, Ci, Oi = _grids(rng, centers, 16.0, 16, 16, 2.0, 100.0)
    params = default_params(resolution=(xres, yres), samples=samples, filter=filter, displays=[_RGBA8])
    return params, _pack(P, Ci, Oi, 16, 16)


def config3(scale=1.0, seed=SEED + 1, motion_px=16.0, fstop=2.8, focallength=0.05, focaldistance=20.0):
    """config2 geometry + Shutter 0 1 with 2 keys per grid (raster motion ~U(0,16) px) + depth of field."""
    xres, yres = max(16, int(1920 * scale)), max(16, int(1080 * scale))
    rng = np.random.default_rng(seed)
    G = max(4, int(78125 * scale * scale))
    centers = np.stack([rng.uniform(-8, xres + 8, G), rng.uniform(-8, yres + 8, G)], axis=1).astype(np.float32)
    P, Ci, Oi = _grids(rng, centers, 16.0, 16, 16, 2.0, 100.0)
    ang = rng.uniform(0, 2 * math.pi, (G, 1))
    mag = rng.uniform(0, motion_px, (G, 1))
    P2 = P.copy()
    P2[:, :, 0] += (mag * np.cos(ang)).astype(np.float32)
    P2[:, :, 1] += (mag * np.sin(ang)).astype(np.float32)
    P2[:, :, 2] *= rng.uniform(0.97, 1.03, (G, 1)).astype(np.float32)
    # DoF scale: raster pixels per camera unit at z=1 for fov 40 degrees (options.cpp:162-171)
    s = 0.5 * yres / math.tan(math.radians(20.0))
    params = default_params(resolution=(xres, yres), samples=(8, 8), filter=("catmull-rom", 3.0, 3.0),
                            shutter=(0.0, 1.0), dof=(fstop, focallength, focaldistance, s, s), displays=[_RGBA8])
    return params, _pack(P, Ci, Oi, 16, 16, P2=P2, key_times=(0.0, 1.0))


def layered_motion(dof=True, scale=0.05, motion_px=6.0):
    """config-3 style frame (two motion keys per grid) whose nearest grids are semi-transparent and lie in front of every
    opaque one (no grid in the depth gap between them, so that the known deviation of DESIGN.md section 2 cannot occur):
    motion blur (and depth of field) feeding the transparent branch of StoreSample and Combine."""
    p, g = config3(scale=scale, motion_px=motion_px)
    if not dof:
        p.use_dof = 0
    G, nv = g.n_grids, 17 * 17
    P = np.asarray(g.P).reshape(G, 2, nv, 3)
    zmin, zmax = P[..., 2].min(axis=(1, 2)), P[..., 2].max(axis=(1, 2))
    near, far = zmax < 30.0, zmin > 36.0
    sel = near | far
    assert near.sum() > 5 and far.sum() > 5
    Oi = np.asarray(g.Oi).reshape(G, nv, 3).copy()
    Oi[near] = np.float32([0.5, 0.25, 0.75])
    g2 = GridArrays(cu=g.cu[sel], cv=g.cv[sel], flags=g.flags[sel], P=np.ascontiguousarray(P[sel].reshape(-1, 3)),
                    Ci=np.ascontiguousarray(np.asarray(g.Ci).reshape(G, nv, 3)[sel].reshape(-1, 3)),
                    Oi=np.ascontiguousarray(Oi[sel].reshape(-1, 3)), nkeys=g.nkeys[sel],
                    key_times=np.ascontiguousarray(np.asarray(g.key_times).reshape(G, 2)[sel].reshape(-1)))
    return p, g2


def to_camera_space(params, grids, fov_deg=40.0):
    """Re-express raster-space grids in camera space and set params.cam_to_raster to the perspective matrix that maps
    them back (row-vector convention of CqMatrix, include/aqsis/math/matrix.h:717-750): the device / the reference then
    do the Project_points step themselves (micropolygon.cpp:723-731).  Camera depth is kept as z."""
    xres, yres = params.xres, params.yres
    fy = 0.5 * yres / math.tan(math.radians(fov_deg / 2.0))
    fx, cx, cy = fy, 0.5 * xres, 0.5 * yres
    m = np.zeros(16, np.float32)
    m[0], m[5] = fx, fy            # x -> x', y -> y'
    m[8], m[9] = cx, cy            # z row: x' += cx*z, y' += cy*z
    m[11] = 1.0                    # h = z
    for i in range(16):
        params.cam_to_raster[i] = float(m[i])
    P = np.asarray(grids.P, np.float32).copy()
    z = P[:, 2].astype(np.float64)
    P[:, 0] = ((P[:, 0].astype(np.float64) - cx) * z / fx).astype(np.float32)
    P[:, 1] = ((P[:, 1].astype(np.float64) - cy) * z / fy).astype(np.float32)
    grids.P = np.ascontiguousarray(P)
    grids.flags = (grids.flags | np.uint32(abi.GRID_CAMERA_SPACE)).astype(np.uint32)
    return params, grids


def deep_stack(n_grids=150, seed=SEED + 11):
    """A small frame under a tall stack of full-frame opaque grids: every tile's bin holds more entries than one
    sorted run (8192), depth complexity = n_grids."""
    xres, yres = 16, 16
    rng = np.random.default_rng(seed)
    centers = np.tile(np.float32([[8.0, 8.0]]), (n_grids, 1)) + rng.uniform(-1, 1, (n_grids, 2)).astype(np.float32)
    P, Ci, Oi = _grids(rng, centers, 20.0, 16, 16, 2.0, 100.0)
    params = default_params(resolution=(xres, yres), samples=(8, 8), filter=("catmull-rom", 3.0, 3.0), displays=[_RGBA8])
    return params, _pack(P, Ci, Oi, 16, 16)


def multikey(scale=0.05, nkeys=3, seed=SEED + 7, motion_px=10.0, dof=True, shutter=(0.0, 1.0)):
    """config-3 style frame whose grids carry `nkeys` motion keys at non-uniform times on a curved path
    (CqMicroPolygonMotion::AppendKey once per key, micropolygon.cpp:1952-1967; BuildBoundList walks the keys, :1689-1756)."""
    xres, yres = max(16, int(1920 * scale)), max(16, int(1080 * scale))
    rng = np.random.default_rng(seed)
    G = max(4, int(78125 * scale * scale))
    centers = np.stack([rng.uniform(-8, xres + 8, G), rng.uniform(-8, yres + 8, G)], axis=1).astype(np.float32)
    P, Ci, Oi = _grids(rng, centers, 16.0, 16, 16, 2.0, 100.0)
    ang = rng.uniform(0, 2 * math.pi, (G, 1))
    mag = rng.uniform(0, motion_px, (G, 1))
    t = np.linspace(0.0, 1.0, nkeys) ** 1.5                       # non-uniform key times in [0, 1]
    times = (shutter[0] + (shutter[1] - shutter[0]) * t).astype(np.float32)
    keys = []
    for k in range(nkeys):
        Pk = P.copy()
        bend = math.sin(math.pi * t[k]) * 0.35                      # sideways bulge: the path is not a straight line
        Pk[:, :, 0] += (mag * (t[k] * np.cos(ang) - bend * np.sin(ang))).astype(np.float32)
        Pk[:, :, 1] += (mag * (t[k] * np.sin(ang) + bend * np.cos(ang))).astype(np.float32)
        Pk[:, :, 2] *= np.float32(1.0 + 0.02 * t[k])
        keys.append(Pk)
    Pall = np.stack(keys, axis=1).reshape(-1, 3)                   # per grid: key-major
    s_ = 0.5 * yres / math.tan(math.radians(20.0))
    kw = {"dof": (2.8, 0.05, 20.0, s_, s_)} if dof else {}
    params = default_params(resolution=(xres, yres), samples=(4, 4), filter=("gaussian", 2.0, 2.0), shutter=shutter,
                            displays=[_RGBA8], **kw)
    g = GridArrays(cu=np.full(G, 16, np.int32), cv=np.full(G, 16, np.int32), flags=np.full(G, abi.GRID_SMOOTH, np.uint32),
                   P=np.ascontiguousarray(Pall.astype(np.float32)), Ci=np.ascontiguousarray(Ci.reshape(-1, 3)),
                   Oi=np.ascontiguousarray(Oi.reshape(-1, 3)), nkeys=np.full(G, nkeys, np.int32), key_times=np.tile(times, G))
    return params, g


def config4(scale=1.0, seed=SEED + 3, layers=4, shard=None):
    """3840x2160, PixelSamples 16 16, ShadingRate 0.25, semi-transparent layered surfaces.

    shard: optional callable(params, block) -> block applied to every layer as it is generated (multi-GPU bench:
    a rank keeps only the grids that touch its strips, so that N ranks on one box never hold N full 5.4 GB scenes).
    The returned GridArrays then carries the totals of the WHOLE scene in .total_micropolygons / .total_vbytes."""
    xres, yres = max(16, int(3840 * scale)), max(16, int(2160 * scale))
    rng = np.random.default_rng(seed)
    blocks = []
    total_mps = total_vbytes = 0
    params = default_params(resolution=(xres, yres), samples=(16, 16), filter=("gaussian", 2.0, 2.0), displays=[_RGBA8])
    opac = [0.25, 0.5, 0.75]
    gx, gy = (xres + 7) // 8 + 1, (yres + 7) // 8 + 1       # 16x16-MP grids of 8x8 px => MP area 0.25 px^2
    for layer in range(layers):
        cx, cy = np.meshgrid(np.arange(gx, dtype=np.float32) * 8.0, np.arange(gy, dtype=np.float32) * 8.0)
        centers = np.stack([cx.ravel(), cy.ravel()], axis=1) + rng.uniform(-1.5, 1.5, (gx * gy, 2)).astype(np.float32)
        z0 = 10.0 + 10.0 * layer
        o = None if layer == layers - 1 else np.full(gx * gy, opac[layer % 3], dtype=np.float32)
        P, Ci, Oi = _grids(rng, centers, 9.0, 16, 16, z0, z0 + 5.0, warp=0.05, noise=0.01, opacity=o, rot=False)
        b = _pack(P, Ci, Oi, 16, 16)
        total_mps += b.n_micropolygons
        total_vbytes += int(b.n_verts) * 36
        blocks.append(shard(params, b) if shard else b)
    out = concat(blocks)
    out.total_micropolygons, out.total_vbytes = total_mps, total_vbytes
    return params, out


def with_aovs(params, grids, aovs=(("N", 3), ("_depthcue", 1), ("_albedo", 3)), seed=SEED + 21, partial=False):
    """Attach arbitrary output variables (RiDisplay "+name" -> CqRenderer::RegisterOutputData, renderer.cpp:1520-1546) to a
    frame: per-vertex values as the shaders would leave them in the grid (StoreExtraData copies the value at the
    micropolygon's index, bucketprocessor.cpp:1573-1643), one extra float32 display showing them."""
    rng = np.random.default_rng(seed)
    params.n_aovs = len(aovs)
    nf = 0
    for i, (name, n) in enumerate(aovs):
        params.aov[i].name = name.encode()
        params.aov[i].n_floats = n
        nf += n
    nv = grids.n_verts
    P = np.asarray(grids.P)
    a = rng.uniform(-1.0, 1.0, (nv, nf)).astype(np.float32)
    a[:, 0] = (np.asarray(grids.Ci)[:, 0] * 2.0 - 1.0)           # something image-like in the first slot
    grids.aov = np.ascontiguousarray(a)
    # a float display of the first AOVs next to the rgba8 one (quantize one = 0: unquantised, ddmanager.cpp:1065)
    d = params.display[params.n_displays]
    d.n_channels = min(nf, 4)
    for c in range(d.n_channels):
        d.channel[c] = abi.NUM_CHANNELS + c
    d.type = 0
    d.quantize_zero = d.quantize_one = d.quantize_min = d.quantize_max = 0.0
    d.quantize_dither = 0.0
    params.n_displays += 1
    return params, grids


def csg_scene(scale=0.1, seed=SEED + 31, op="difference", nested=False):
    """Two (nested: three) solids built from grid "shells" and combined by a CSG tree (RiSolidBegin "difference" ...):
    every solid is a front and a back sheet (entry and exit wall) at different depths, overlapping on screen, plus plain
    opaque and transparent grids around them.  The tree rides on params._csg = (types, parents)."""
    xres, yres = max(32, int(640 * scale)), max(32, int(480 * scale))
    rng = np.random.default_rng(seed)
    params = default_params(resolution=(xres, yres), samples=(3, 3), filter=("gaussian", 2.0, 2.0), displays=[_RGBA8])
    opn = {"union": abi.CSG_UNION, "intersection": abi.CSG_INTERSECTION, "difference": abi.CSG_DIFFERENCE}[op]
    if nested:
        # node 0: op(node 1 = union(prim 2, prim 3), prim 4)
        types = [opn, abi.CSG_UNION, abi.CSG_PRIMITIVE, abi.CSG_PRIMITIVE, abi.CSG_PRIMITIVE]
        parents = [-1, 0, 1, 1, 0]
        prims = [2, 3, 4]
    else:
        types = [opn, abi.CSG_PRIMITIVE, abi.CSG_PRIMITIVE]
        parents = [-1, 0, 0]
        prims = [1, 2]
    blocks, csg = [], []
    cx, cy = 0.5 * xres, 0.5 * yres
    # a small opaque patch in front of everything, submitted FIRST: what is stored behind an opaque hit that already
    # exists is dropped at store time by the reference (bucketprocessor.cpp:1475-1480) unless it belongs to a CSG solid
    P, Ci, Oi = _grids(rng, np.float32([[cx - 0.3 * xres, cy - 0.2 * yres]]), 0.25 * min(xres, yres), 8, 8, 5.0, 5.2)
    blocks.append(_pack(P, Ci, Oi, 8, 8)); csg.append(-1)
    for k, node in enumerate(prims):
        ox = (k - (len(prims) - 1) / 2.0) * 0.22 * xres
        size = 0.55 * min(xres, yres)
        for wall, z in enumerate((10.0 + 2.0 * k, 20.0 + 3.0 * k)):          # entry wall, exit wall
            centers = np.float32([[cx + ox, cy + (k % 2) * 0.08 * yres]])
            P, Ci, Oi = _grids(rng, centers, size, 16, 16, z, z + 0.5, warp=0.03, noise=0.0, rot=False)
            blocks.append(_pack(P, Ci, Oi, 16, 16, flags=abi.GRID_SMOOTH | abi.GRID_USES_CSG))
            csg.append(node)
    # a plain opaque backdrop behind and a transparent sheet in the middle of the solids
    P, Ci, Oi = _grids(rng, np.float32([[cx, cy]]), 1.3 * max(xres, yres), 16, 16, 40.0, 41.0, warp=0.0, noise=0.0, rot=False)
    blocks.append(_pack(P, Ci, Oi, 16, 16)); csg.append(-1)
    P, Ci, Oi = _grids(rng, np.float32([[cx, cy]]), 0.7 * max(xres, yres), 16, 16, 15.0, 15.5, warp=0.02, noise=0.0, opacity=[0.4])
    blocks.append(_pack(P, Ci, Oi, 16, 16)); csg.append(-1)
    g = concat(blocks)
    g.csg_node = np.asarray(csg, np.int32)
    params._csg = (np.asarray(types, np.int32), np.asarray(parents, np.int32))
    return params, g


def points_scene(scale=0.15, seed=SEED + 41, n_points=4000, dof=False, transparent=True, res=None):
    """RiPoints as the hider sees them (CqMicroPolyGridPoints, geometry/points.cpp): grids of up to 256 points, every point a
    disc of its own raster radius with constant colour / opacity; over a backdrop of ordinary grids."""
    xres, yres = res if res else (max(32, int(640 * scale)), max(32, int(480 * scale)))
    rng = np.random.default_rng(seed)
    kw = {}
    if dof:
        s_ = 0.5 * yres / math.tan(math.radians(20.0))
        kw["dof"] = (2.8, 0.05, 20.0, s_, s_)
    params = default_params(resolution=(xres, yres), samples=(4, 4), filter=("gaussian", 2.0, 2.0), displays=[_RGBA8], **kw)
    G = 12
    centers = np.stack([rng.uniform(0, xres, G), rng.uniform(0, yres, G)], axis=1).astype(np.float32)
    P, Ci, Oi = _grids(rng, centers, 24.0, 8, 8, 30.0, 60.0)
    back = _pack(P, Ci, Oi, 8, 8)
    per = 250
    ng = (n_points + per - 1) // per
    cu = np.full(ng, per - 1, np.int32)
    npts = ng * per
    Pp = np.stack([rng.uniform(-2, xres + 2, npts), rng.uniform(-2, yres + 2, npts), rng.uniform(5.0, 70.0, npts)], axis=1).astype(np.float32)
    rad = rng.uniform(0.15, 2.5, npts).astype(np.float32)
    Cip = rng.uniform(0.1, 1.0, (npts, 3)).astype(np.float32)
    Oip = np.ones((npts, 3), np.float32)
    if transparent:
        t = rng.uniform(size=npts) < 0.4
        Oip[t] = rng.uniform(0.2, 0.9, (int(t.sum()), 1)).astype(np.float32)
        Cip[t] *= Oip[t]
    pts = GridArrays(cu=cu, cv=np.zeros(ng, np.int32), flags=np.full(ng, abi.GRID_POINTS, np.uint32), P=Pp, Ci=Cip, Oi=Oip)
    g = concat([back, pts])
    r = np.zeros(g.P.shape[0], np.float32)
    r[back.P.shape[0]:] = rad
    g.radius = r
    return params, g


def cull_scene(scale=0.12, seed=SEED + 51):
    """Camera-space grids with geometric normals for the culls CqMicroPolyGrid::Shade applies before busting
    (micropolygon.cpp:431-474 backfacing, :493-522 fully transparent): half of the grids face away, some end in a run of
    Oi = 0 vertices."""
    p, g = to_camera_space(*config1(scale=scale, seed=seed))
    rng = np.random.default_rng(seed + 1)
    nv = 81
    G = g.n_grids
    Pc = np.asarray(g.P).reshape(G, nv, 3)
    # normals: towards the camera (-P direction) for even grids, away for odd ones, jittered so that the sign varies inside a grid
    Ng = -Pc / np.linalg.norm(Pc, axis=2, keepdims=True)
    Ng = Ng + rng.normal(0, 0.8, Ng.shape)
    Ng[1::2] *= -1.0
    g.Ng = np.ascontiguousarray(Ng.reshape(-1, 3).astype(np.float32))
    N = Ng.copy()
    N[::3] *= -1.0                                                  # a user normal on the other side flips Ng's facing
    g.N = np.ascontiguousarray(N.reshape(-1, 3).astype(np.float32))
    Oi = np.asarray(g.Oi).reshape(G, nv, 3).copy()
    Ci = np.asarray(g.Ci).reshape(G, nv, 3).copy()
    tail = rng.integers(0, nv, G)
    for i in range(0, G, 4):
        Oi[i, tail[i]:] = 0.0                                       # trailing run of fully transparent vertices
        Ci[i, tail[i]:] = 0.0
        if tail[i] > 10:
            Oi[i, 3] = 0.0                                          # an isolated one in the middle must survive (the loop breaks)
    g.Oi = np.ascontiguousarray(Oi.reshape(-1, 3))
    g.Ci = np.ascontiguousarray(Ci.reshape(-1, 3))
    g.flags = (g.flags | np.uint32(abi.GRID_CULL_BACKFACING | abi.GRID_CULL_TRANSPARENT)).astype(np.uint32)
    return p, g


def trim_scene(scale=0.12, seed=SEED + 61, outside_every=3, motion=False, dof=False):
    """Trimmed NURBS patches as the hider sees them (RiTrimCurve): grids with surface parameters (u, v) per vertex and, per
    surface, closed trim loops tessellated to polylines (CqTrimLoop::Prepare).  Every third trimmed surface has
    Attribute "trimcurve" "sense" "outside"; some grids are untrimmed; one set has no loops at all; a loop with a hole
    (two nested loops) exercises the crossing parity.  The loops ride on params._trim."""
    xres, yres = max(32, int(640 * scale)), max(32, int(480 * scale))
    rng = np.random.default_rng(seed)
    kw = {}
    if dof:
        s_ = 0.5 * yres / math.tan(math.radians(20.0))
        kw["dof"] = (2.8, 0.05, 20.0, s_, s_)
    if motion:
        kw["shutter"] = (0.0, 1.0)
    params = default_params(resolution=(xres, yres), samples=(4, 4), filter=("gaussian", 2.0, 2.0), displays=[_RGBA8], **kw)
    G = 40
    cu = cv = 12
    nv = (cu + 1) * (cv + 1)
    centers = np.stack([rng.uniform(0, xres, G), rng.uniform(0, yres, G)], axis=1).astype(np.float32)
    opac = np.where(rng.uniform(size=G) < 0.3, 0.6, 1.0).astype(np.float32)
    P, Ci, Oi = _grids(rng, centers, 18.0, cu, cv, 25.0, 60.0, opacity=opac)
    # the transparent grids lie in front of every opaque one (a transparent hit BEHIND an opaque surface that is submitted
    # later stays in the reference's sample list and changes the rounding of the composite: see DESIGN.md, known deviations)
    P[opac < 1.0, :, 2] = np.float32(5.0) + (P[opac < 1.0, :, 2] - np.float32(25.0)) * np.float32(0.4)
    P2 = None
    if motion:
        P2 = P.copy()
        P2[:, :, 0] += rng.uniform(-5, 5, (G, 1)).astype(np.float32)
        P2[:, :, 1] += rng.uniform(-5, 5, (G, 1)).astype(np.float32)
    g = _pack(P, Ci, Oi, cu, cv, P2=P2, key_times=(0.0, 1.0) if motion else None)
    # surface parameters: every grid is a sub-rectangle [u0,u1] x [v0,v1] of its patch's unit square
    u = np.linspace(0.0, 1.0, cu + 1, dtype=np.float32)
    v = np.linspace(0.0, 1.0, cv + 1, dtype=np.float32)
    uu, vv = np.meshgrid(u, v)
    lo = rng.uniform(0.0, 0.3, (G, 2)).astype(np.float32)
    hi = rng.uniform(0.7, 1.0, (G, 2)).astype(np.float32)
    uv = np.empty((G, nv, 2), np.float32)
    uv[:, :, 0] = lo[:, 0:1] + (hi[:, 0:1] - lo[:, 0:1]) * uu.reshape(1, -1)
    uv[:, :, 1] = lo[:, 1:2] + (hi[:, 1:2] - lo[:, 1:2]) * vv.reshape(1, -1)
    g.trim_uv = np.ascontiguousarray(uv.reshape(-1, 2))
    # trim sets: 0 a disc, 1 a disc with a hole, 2 a star-like loop and a separate small loop, 3 no loops at all
    def circle(cx, cy, r, n, wobble=0.0):
        t = np.linspace(0, 2 * math.pi, n, endpoint=False)
        rr = r * (1.0 + wobble * np.sin(5 * t))
        return np.stack([cx + rr * np.cos(t), cy + rr * np.sin(t)], axis=1).astype(np.float32)
    sets = [[circle(0.5, 0.5, 0.33, 24)],
            [circle(0.5, 0.5, 0.4, 32), circle(0.55, 0.45, 0.15, 16)],
            [circle(0.45, 0.5, 0.3, 40, wobble=0.35), circle(0.85, 0.85, 0.08, 8)],
            []]
    set_first, loop_first, pts = [0], [0], []
    for loops in sets:
        for lp in loops:
            pts.append(lp)
            loop_first.append(loop_first[-1] + len(lp))
        set_first.append(set_first[-1] + len(loops))
    params._trim = (np.asarray(set_first, np.int32), np.asarray(loop_first, np.int32), np.concatenate(pts).astype(np.float32))
    ts = (np.arange(G) % 5).astype(np.int32)              # 0 = untrimmed, 1..4 = the sets
    g.trim_set = ts
    fl = g.flags.copy()
    fl[(np.arange(G) % outside_every == 1) & (ts != 0)] |= np.uint32(abi.GRID_TRIM_OUTSIDE)
    g.flags = fl.astype(np.uint32)
    return params, g


def config5_filters():
    """The PixelFilter sweep of config 5: (name, width) pairs run on the config-2 scene."""
    return [(name, float(w)) for name in ("box", "triangle", "gaussian", "catmull-rom", "sinc") for w in range(1, 7)]


def algorithmic_bytes(params, grids: GridArrays):
    """B_alg of SURVEY.md 8(d): V*(12K + 24) + W*H*(36 + E)."""
    nv = (grids.cu.astype(np.int64) + 1) * (grids.cv.astype(np.int64) + 1)
    nk = grids.nkeys.astype(np.int64) if grids.nkeys is not None else np.ones_like(nv)
    vbytes = int((nv * (12 * nk + 24)).sum())
    w = params.crop_xmax - params.crop_xmin
    h = params.crop_ymax - params.crop_ymin
    e = 0
    from .hider import display_info
    for d in range(params.n_displays):
        e += display_info(params, d)[2]
    return vbytes + w * h * (36 + e)
