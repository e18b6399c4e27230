"""aqsis "mpdump" files: reader, writer and replay through the hider.

An aqsis built with AQSIS_ENABLE_MPDUMP writes `mpdump.mp` while rendering (libs/core/mpdump.cpp:41-217;
analysed by tools/scripts/mpanalyse.py).  The file is a wire format for exactly what crosses the hider's
inbound seam -- every busted micropolygon -- plus the sample positions the reference used, so a render made
elsewhere can be replayed through the B200 hider and the sample pattern replay can be checked against it.

Layout (native endianness, no padding; mpdump.cpp:44-56):
    int32   sizeof(TqFloat)                       -- 4
    records, each starting with an int16 id:
      1  micropolygon  (mpdump.cpp:147-180): 4 x (x, y, z) float vertices in the CIRCULAR order P0 P1 P3 P2 of
         CqMicroPolygon::GetVertices, then Ci (r, g, b) and Oi (r, g, b) -- the colour / opacity at the
         micropolygon's index, i.e. constant shading; (0.9, 0.9, 1) when the grid has no Ci / Oi
      2  pixel sample  (mpdump.cpp:117-144): int32 x, int32 y, int32 index, float pos.x, float pos.y
      3  image info    (mpdump.cpp:72-90):   int32 width, int32 height
"""
from dataclasses import dataclass, field

import numpy as np

from . import _abi as abi
from .hider import GridArrays

_MP = np.dtype([("id", "<i2"), ("v", "<f4", (4, 3)), ("ci", "<f4", 3), ("oi", "<f4", 3)])         # 2 + 72 bytes
_SAMPLE = np.dtype([("id", "<i2"), ("x", "<i4"), ("y", "<i4"), ("idx", "<i4"), ("px", "<f4"), ("py", "<f4")])
_IMAGE = np.dtype([("id", "<i2"), ("w", "<i4"), ("h", "<i4")])
assert _MP.itemsize == 74 and _SAMPLE.itemsize == 22 and _IMAGE.itemsize == 10
_CIRCULAR = [0, 1, 3, 2]      # file order -> natural (bilinear patch) order and back: the permutation is its own inverse


@dataclass
class MPDump:
    width: int = 0
    height: int = 0
    P: np.ndarray = field(default_factory=lambda: np.zeros((0, 4, 3), np.float32))   # natural order P0 P1 P2 P3
    Ci: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.float32))
    Oi: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.float32))
    samples: np.ndarray = field(default_factory=lambda: np.zeros(0, _SAMPLE))         # fields x, y, idx, px, py

    @property
    def n_micropolygons(self):
        return int(self.P.shape[0])


def read(path) -> MPDump:
    """Parse an mpdump file.  Records of one kind usually come in long runs (all samples of a bucket, all
    micropolygons of a grid), so runs are decoded with numpy instead of record by record."""
    raw = np.fromfile(path, dtype=np.uint8)
    if raw.size < 4 or int(raw[:4].view("<i4")[0]) != 4:
        raise ValueError("not an mpdump file written with 4-byte floats")
    out = MPDump()
    mps, smp = [], []
    pos, n = 4, raw.size
    sizes = {1: _MP.itemsize, 2: _SAMPLE.itemsize, 3: _IMAGE.itemsize}
    types = {1: _MP, 2: _SAMPLE, 3: _IMAGE}
    while pos < n:
        if pos + 2 > n:
            raise ValueError("truncated mpdump file")
        rid = int(raw[pos:pos + 2].view("<i2")[0])
        if rid not in sizes:
            raise ValueError(f"unknown mpdump record id {rid} at byte {pos}")
        sz = sizes[rid]
        # length of the run of records with this id
        cap = (n - pos) // sz
        if cap == 0:
            raise ValueError("truncated mpdump record")
        ids = raw[pos:pos + cap * sz].reshape(cap, sz)[:, :2].copy().view("<i2")[:, 0]
        stop = np.nonzero(ids != rid)[0]
        run = int(stop[0]) if stop.size else cap
        recs = raw[pos:pos + run * sz].view(types[rid])
        if rid == 1:
            mps.append(recs)
        elif rid == 2:
            smp.append(recs)
        else:
            out.width, out.height = int(recs["w"][-1]), int(recs["h"][-1])
        pos += run * sz
    if mps:
        m = np.concatenate(mps)
        out.P = np.ascontiguousarray(m["v"][:, _CIRCULAR, :])
        out.Ci = np.ascontiguousarray(m["ci"])
        out.Oi = np.ascontiguousarray(m["oi"])
    if smp:
        out.samples = np.concatenate(smp)
    return out


def write(path, dump: MPDump):
    """Write an mpdump file the way CqMPDump does: header, image info, samples, micropolygons."""
    with open(path, "wb") as f:
        np.array([4], "<i4").tofile(f)
        if dump.width or dump.height:
            np.array([(3, dump.width, dump.height)], _IMAGE).tofile(f)
        if len(dump.samples):
            s = np.zeros(len(dump.samples), _SAMPLE)
            for k in ("x", "y", "idx", "px", "py"):
                s[k] = dump.samples[k]
            s["id"] = 2
            s.tofile(f)
        n = dump.n_micropolygons
        if n:
            m = np.zeros(n, _MP)
            m["id"] = 1
            m["v"] = np.asarray(dump.P, np.float32)[:, _CIRCULAR, :]
            m["ci"], m["oi"] = dump.Ci, dump.Oi
            m.tofile(f)


def to_grids(dump: MPDump) -> GridArrays:
    """Every dumped micropolygon as a 1x1 grid with constant shading (the dump holds ONE colour per
    micropolygon: the value at its index) in dump order, i.e. the reference's submission order."""
    n = dump.n_micropolygons
    P = np.ascontiguousarray(np.asarray(dump.P, np.float32).reshape(n * 4, 3))
    # vertices index, index+1, index+cu+1, index+cu+2 of a 1x1 grid are its four corners in natural order;
    # constant shading reads corner 0 (micropolygon.cpp:1518-1521)
    Ci = np.repeat(np.asarray(dump.Ci, np.float32), 4, axis=0)
    Oi = np.repeat(np.asarray(dump.Oi, np.float32), 4, axis=0)
    return GridArrays(cu=np.ones(n, np.int32), cv=np.ones(n, np.int32), flags=np.zeros(n, np.uint32),
                      P=P, Ci=np.ascontiguousarray(Ci), Oi=np.ascontiguousarray(Oi))


def from_grids(grids: GridArrays, width=0, height=0) -> MPDump:
    """Bust (static, raster-space) grids the way CqMicroPolyGrid::Split does (micropolygon.cpp:770-856: iv outer,
    iu inner, culled micropolygons skipped) and record what CqMPDump::dump would: key 0 vertices, colour and
    opacity at the micropolygon's index."""
    Pall = np.asarray(grids.P, np.float32)
    Ci = None if grids.Ci is None else np.asarray(grids.Ci, np.float32)
    Oi = None if grids.Oi is None else np.asarray(grids.Oi, np.float32)
    culled = None if grids.culled is None else np.asarray(grids.culled)
    Ps, Cs, Os = [], [], []
    po = vo = 0
    for g in range(grids.n_grids):
        cu, cv = int(grids.cu[g]), int(grids.cv[g])
        nk = int(grids.nkeys[g]) if grids.nkeys is not None else 1
        nv = (cu + 1) * (cv + 1)
        if int(grids.flags[g]) & abi.GRID_CAMERA_SPACE:
            raise ValueError("from_grids expects raster-space grids")
        iv, iu = np.meshgrid(np.arange(cv), np.arange(cu), indexing="ij")
        idx = (iv * (cu + 1) + iu).ravel()
        if culled is not None:
            idx = idx[culled[vo + idx] == 0]
        Pg = Pall[po:po + nv]
        Ps.append(np.stack([Pg[idx], Pg[idx + 1], Pg[idx + cu + 1], Pg[idx + cu + 2]], axis=1))
        default = np.float32([0.9, 0.9, 1.0])
        Cs.append(Ci[vo + idx] if Ci is not None else np.tile(default, (len(idx), 1)))
        Os.append(Oi[vo + idx] if Oi is not None else np.tile(default, (len(idx), 1)))
        po += nv * nk
        vo += nv
    cat = lambda xs, shape: np.concatenate(xs) if xs else np.zeros(shape, np.float32)
    return MPDump(width=width, height=height, P=cat(Ps, (0, 4, 3)), Ci=cat(Cs, (0, 3)), Oi=cat(Os, (0, 3)))


def expected_samples(params) -> np.ndarray:
    """The sample records (x, y, idx, pos) the reference would dump for this frame, from the library's replay of the
    global random stream (aqh_replay_frame_rng + aqh_sampler_tables): one record per sample of every pixel of the
    sample region.  Comparing these with the id-2 records of a real dump checks the RNG replay end to end."""
    import ctypes as C
    from .hider import lib
    L = lib()
    n = params.xsamples * params.ysamples
    sx0, sy0, sw, sh = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    # first call sizes the planes
    rc = L.aqh_replay_frame_rng(C.byref(params), None, None, C.byref(sx0), C.byref(sy0), C.byref(sw), C.byref(sh))
    if rc:
        raise RuntimeError(f"aqh_replay_frame_rng failed ({rc})")
    planes = np.zeros(5 * sw.value * sh.value, np.uint8)
    dither = np.zeros(max(1, params.n_displays) * params.xres * params.yres, np.float32)
    rc = L.aqh_replay_frame_rng(C.byref(params), planes.ctypes.data, dither.ctypes.data,
                                C.byref(sx0), C.byref(sy0), C.byref(sw), C.byref(sh))
    if rc:
        raise RuntimeError(f"aqh_replay_frame_rng failed ({rc})")
    r = L.aqh_random_create(int(params.rng_seed))
    try:
        for _ in range(int(params.rng_predraws)):
            L.aqh_random_uint(r)
        ncache = C.c_int()
        pos = np.zeros(250 * n * 2, np.float32)
        v1d = np.zeros(250 * n, np.float32)
        shuf = np.zeros(250 * n, np.int32)
        rc = L.aqh_sampler_tables(r, params.xsamples, params.ysamples, int(params.jitter),
                                  pos.ctypes.data, v1d.ctypes.data, shuf.ctypes.data, C.byref(ncache))
        if rc:
            raise RuntimeError(f"aqh_sampler_tables failed ({rc})")
    finally:
        L.aqh_random_destroy(r)
    pos = pos[:ncache.value * n * 2].reshape(ncache.value, n, 2)
    pat_pos = planes.reshape(5, sh.value, sw.value)[1]
    ys, xs = np.mgrid[0:sh.value, 0:sw.value]
    X = (xs + sx0.value).astype(np.int32)
    Y = (ys + sy0.value).astype(np.int32)
    out = np.zeros(sh.value * sw.value * n, _SAMPLE)
    out["id"] = 2
    out["idx"] = np.tile(np.arange(n, dtype=np.int32), sh.value * sw.value)
    p = pos[pat_pos.ravel()]                                   # (pixels, n, 2)
    # position = (x, y) + pos[i] in float (imagepixel.cpp:349); the record's x, y are lfloor() of THAT
    # (mpdump.cpp:110), which is the next pixel when the sum rounds up to it
    out["px"] = (np.repeat(X.ravel(), n).astype(np.float32) + p[..., 0].ravel()).astype(np.float32)
    out["py"] = (np.repeat(Y.ravel(), n).astype(np.float32) + p[..., 1].ravel()).astype(np.float32)
    out["x"] = np.floor(out["px"]).astype(np.int32)
    out["y"] = np.floor(out["py"]).astype(np.int32)
    return out
