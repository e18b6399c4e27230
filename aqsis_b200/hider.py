"""Python face of the hider's C ABI (include/aqsis_b200_hider.h) via ctypes.

This module is plumbing for tests, the benchmark and multi-GPU orchestration; the product
is the native library.  It never falls back to a CPU implementation: if the library is
missing, or no sm_100 device is present, construction of `Hider` raises.
"""
import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _abi as abi
from ._abi import FrameParams, DisplayDesc, GridBlock, GridDesc, Callbacks, FrameStats

_LIB = None
_LIB_PATH = os.environ.get("AQSIS_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib", "libaqsis_b200_hider.so")


class HiderError(RuntimeError):
    def __init__(self, status, message=""):
        self.status = status
        super().__init__(f"{abi.STATUS_NAMES.get(status, status)}: {message}")


def lib():
    """Load the native library (building is the job of aqsis_b200.build / __graft_entry__.build)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(_LIB_PATH):
        raise HiderError(abi.AQH_ERR_NO_DEVICE,
                         f"{_LIB_PATH} is missing: run `python -m aqsis_b200.build` (there is no CPU fallback)")
    L = C.CDLL(_LIB_PATH)
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    L.aqh_create.argtypes = [C.POINTER(vp), ci]
    L.aqh_destroy.argtypes = [vp]
    L.aqh_last_error.argtypes = [vp]
    L.aqh_last_error.restype = C.c_char_p
    L.aqh_set_stream.argtypes = [vp, vp]
    L.aqh_frame_params_default.argtypes = [C.POINTER(FrameParams)]
    L.aqh_frame_params_set_dof.argtypes = [C.POINTER(FrameParams), cf, cf, cf, cf, cf]
    L.aqh_display_from_mode.argtypes = [C.POINTER(DisplayDesc), C.c_char_p, ci, cf, cf, cf, cf]
    L.aqh_begin_frame.argtypes = [vp, C.POINTER(FrameParams)]
    L.aqh_add_grid.argtypes = [vp, C.POINTER(GridDesc)]
    L.aqh_add_grid_block.argtypes = [vp, C.POINTER(GridBlock)]
    L.aqh_end_frame.argtypes = [vp, C.POINTER(Callbacks)]
    L.aqh_render_device.argtypes = [vp]
    L.aqh_flush.argtypes = [vp]
    L.aqh_can_cull.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_int)]
    L.aqh_frame_stats.argtypes = [vp, C.POINTER(FrameStats)]
    L.aqh_image_channels.argtypes = [vp, C.POINTER(C.POINTER(C.c_float)), C.POINTER(ci), C.POINTER(ci)]
    L.aqh_image_display.argtypes = [vp, ci, C.POINTER(C.POINTER(C.c_ubyte)), C.POINTER(ci), C.POINTER(ci)]
    L.aqh_device_channels.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.aqh_device_display.argtypes = [vp, ci, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.aqh_strip_layout.argtypes = [C.POINTER(FrameParams), ci, C.POINTER(ci), vp, vp, ci]
    L.aqh_num_strips.argtypes = [vp, C.POINTER(ci)]
    L.aqh_strip.argtypes = [vp, ci, C.POINTER(ci), C.POINTER(ci)]
    for name in ("box", "triangle", "gaussian", "catmullrom", "sinc", "mitchell", "disk", "bessel"):
        fn = getattr(L, f"aqh_{name}_filter")
        fn.argtypes = [cf, cf, cf, cf]
        fn.restype = cf
    L.aqh_filter_by_name.argtypes = [C.c_char_p]
    L.aqh_filter_by_name.restype = vp
    L.aqh_random_create.argtypes = [C.c_uint32]
    L.aqh_random_create.restype = vp
    L.aqh_random_destroy.argtypes = [vp]
    L.aqh_random_destroy.restype = None
    L.aqh_random_reseed.argtypes = [vp, C.c_uint32]
    L.aqh_random_reseed.restype = None
    L.aqh_random_uint.argtypes = [vp]
    L.aqh_random_uint.restype = C.c_uint32
    L.aqh_random_float.argtypes = [vp]
    L.aqh_random_float.restype = cf
    L.aqh_random_int.argtypes = [vp, C.c_uint32]
    L.aqh_random_int.restype = C.c_uint32
    L.aqh_sampler_tables.argtypes = [vp, ci, ci, ci, vp, vp, vp, C.POINTER(ci)]
    L.aqh_replay_frame_rng.argtypes = [C.POINTER(FrameParams), vp, vp, C.POINTER(ci), C.POINTER(ci),
                                       C.POINTER(ci), C.POINTER(ci)]
    L.aqh_filter_table.argtypes = [C.POINTER(FrameParams), vp, C.POINTER(ci)]
    L.aqh_set_csg_tree.argtypes = [vp, ci, vp, vp]
    L.aqh_set_trim_loops.argtypes = [vp, ci, vp, vp, vp]
    L.aqh_channel_count.argtypes = [vp, C.POINTER(ci)]
    L.aqh_grid_rank_masks.argtypes = [C.POINTER(FrameParams), C.POINTER(GridBlock), vp]
    L.aqh_grid_row_cost.argtypes = [C.POINTER(FrameParams), C.POINTER(GridBlock), vp]
    L.aqh_balance_strips.argtypes = [C.POINTER(FrameParams), vp]
    L.aqh_comm_unique_id.argtypes = [vp]
    L.aqh_comm_init.argtypes = [vp, vp, ci, ci]
    L.aqh_comm_destroy.argtypes = [vp]
    L.aqh_gather.argtypes = [vp, ci]
    L.aqh_clear_caches.argtypes = [vp]
    _LIB = L
    return L


FILTER_NAMES = ("box", "triangle", "gaussian", "catmull-rom", "sinc", "mitchell", "disk", "bessel")


# Pure-python mode: nothing in this module touches the native library (bench.py --impl reference must time the
# reference's CPU hider in a process that never maps the product).  The pixel filter is then carried by NAME in
# p._filter_name (filter_func stays NULL); tests/orc.py hands it to the reference / the oracle.
PURE = False


def _display_from_mode(d: DisplayDesc, mode, driver_order, one, mn, mx, dither):
    """aqh_display_from_mode in python (core order a,r,g,b,z, ddmanager.cpp:455-480; file driver order r,g,b,a)."""
    rgb, a, z = "rgb" in mode, "a" in mode, "z" in mode
    if not (rgb or a or z):
        raise HiderError(abi.AQH_ERR_BAD_PARAMS, f"display mode {mode!r}")
    ch = []
    if not driver_order and a:
        ch.append(abi.CH_ALPHA)
    if rgb:
        ch += [abi.CH_CI_R, abi.CH_CI_G, abi.CH_CI_B]
    if driver_order and a:
        ch.append(abi.CH_ALPHA)
    if z:
        ch.append(abi.CH_Z)
    C.memset(C.byref(d), 0, C.sizeof(d))
    d.n_channels = len(ch)
    for i, c in enumerate(ch):
        d.channel[i] = c
    d.quantize_zero, d.quantize_one, d.quantize_min, d.quantize_max, d.quantize_dither = 0.0, one, mn, mx, dither


def _defaults_pure(p: FrameParams):
    """aqh_frame_params_default in python (CqOptions defaults, libs/core/options.cpp:273-305); tests/test_abi.py
    checks that both fill the struct identically."""
    C.memset(C.byref(p), 0, C.sizeof(p))
    p.abi_version = abi.AQH_ABI_VERSION
    p.xres, p.yres = 640, 480
    p.crop_xmin, p.crop_xmax, p.crop_ymin, p.crop_ymax = 0, 640, 0, 480
    p.xsamples = p.ysamples = 2
    p.filter_xwidth = p.filter_ywidth = 2.0
    p.bucket_xsize = p.bucket_ysize = 16
    p.clip_near, p.clip_far = float(np.finfo(np.float32).eps), float(np.finfo(np.float32).max)
    p.depth_filter = abi.DEPTHFILTER_MIN
    p.zthreshold[0] = p.zthreshold[1] = p.zthreshold[2] = 1.0
    p.display_mode = abi.DMODE_RGB | abi.DMODE_A
    p.exposure_gain = p.exposure_gamma = 1.0
    p.jitter = 1
    for i in range(4):
        p.cam_to_raster[i * 4 + i] = 1.0
    p.rng_seed = 545
    p.world_size = 1
    p.filter_mode = abi.FILTER_REFERENCE_ORDER


def _set_dof_pure(p: FrameParams, fstop, focallength, focaldistance, sx, sy):
    """CqRenderer::SetDepthOfFieldData (renderer.h:368-377) with the same float / double staging."""
    f32 = np.float32
    p.use_dof = 1 if f32(fstop) < np.finfo(np.float32).max else 0
    if p.use_dof:
        lens = f32(focallength) / f32(fstop)
        fd = float(f32(focaldistance))
        p.dof_multiplier = float(f32(0.5 * float(lens) * fd / (fd + float(lens))))
        p.dof_one_over_focal_distance = float(f32(1.0 / fd))
    p.dof_scale_x, p.dof_scale_y = sx, sy


def default_params(**kw) -> FrameParams:
    """AqhFrameParams with the reference's option defaults, then overrides.

    Convenience keys: resolution=(x,y) also resets the crop window; samples=(xs,ys);
    filter=("name", xw, yw); displays=[("rgba", driver_order, one, min, max, dither), ...];
    aovs=[("name", n_floats), ...]."""
    p = FrameParams()
    if PURE:
        _defaults_pure(p)
    else:
        lib().aqh_frame_params_default(C.byref(p))
    p._filter_name = "gaussian"
    if "resolution" in kw:
        x, y = kw.pop("resolution")
        p.xres, p.yres = x, y
        p.crop_xmin, p.crop_xmax, p.crop_ymin, p.crop_ymax = 0, x, 0, y
    if "crop" in kw:
        p.crop_xmin, p.crop_xmax, p.crop_ymin, p.crop_ymax = kw.pop("crop")
    if "samples" in kw:
        p.xsamples, p.ysamples = kw.pop("samples")
    if "filter" in kw:
        name, xw, yw = kw.pop("filter")
        if name not in FILTER_NAMES:
            raise ValueError(f"unknown pixel filter {name!r}")
        p._filter_name = name
        if not PURE:
            p.filter_func = lib().aqh_filter_by_name(name.encode())
        p.filter_xwidth, p.filter_ywidth = xw, yw
    if "dof" in kw:
        fstop, fl, fd, sx, sy = kw.pop("dof")
        if PURE:
            _set_dof_pure(p, fstop, fl, fd, sx, sy)
        else:
            lib().aqh_frame_params_set_dof(C.byref(p), fstop, fl, fd, sx, sy)
    if "shutter" in kw:
        p.shutter_open, p.shutter_close = kw.pop("shutter")
    if "exposure" in kw:
        p.exposure_gain, p.exposure_gamma = kw.pop("exposure")
    if "displays" in kw:
        ds = kw.pop("displays")
        p.n_displays = len(ds)
        for i, d in enumerate(ds):
            mode, driver_order, one, mn, mx, dither = d
            _display_from_mode(p.display[i], mode, int(driver_order), one, mn, mx, dither)
    if "aovs" in kw:
        av = kw.pop("aovs")
        p.n_aovs = len(av)
        for i, (name, nf) in enumerate(av):
            p.aov[i].name = name.encode()
            p.aov[i].n_floats = nf
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(f"AqhFrameParams has no field {k!r}")
        setattr(p, k, v)
    return p


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    rc = lib().aqh_comm_unique_id(buf)
    if rc:
        raise HiderError(rc, "aqh_comm_unique_id (is libnccl.so.2 loadable?)")
    return buf.raw


def display_info(p: FrameParams, d: int):
    """(numpy dtype, channels, entrysize) of display d, following selectDataFormat."""
    dd = p.display[d]
    t = dd.type
    if t == 0:
        one, mn, mx = dd.quantize_one, dd.quantize_min, dd.quantize_max
        if one == 0:
            t = abi.FLOAT32
        elif mn >= 0:
            t = abi.UNSIGNED8 if mx <= 255 else (abi.UNSIGNED16 if mx <= 65535 else abi.UNSIGNED32)
        else:
            t = abi.SIGNED8 if (mn >= -128 and mx <= 127) else (abi.SIGNED16 if (mn >= -32768 and mx <= 32767) else abi.SIGNED32)
    return np.dtype(abi.TYPE_NUMPY[t]), dd.n_channels, abi.TYPE_SIZES[t] * dd.n_channels


@dataclass
class GridArrays:
    """A run of grids in the packed layout of AqhGridBlock (host numpy arrays or CUDA tensors)."""
    cu: np.ndarray
    cv: np.ndarray
    flags: np.ndarray
    P: object                      # (sum nkeys*nverts, 3) float32
    Ci: Optional[object] = None    # (sum nverts, 3) float32
    Oi: Optional[object] = None
    nkeys: Optional[np.ndarray] = None
    key_times: Optional[np.ndarray] = None
    lod_bounds: Optional[np.ndarray] = None
    culled: Optional[object] = None
    aov: Optional[object] = None       # (sum nverts, aov floats) float32
    Ng: Optional[object] = None        # (sum nverts, 3)
    N: Optional[object] = None
    radius: Optional[object] = None    # (sum nkeys*nverts,) float32, read for GRID_POINTS grids
    csg_node: Optional[np.ndarray] = None
    trim_set: Optional[np.ndarray] = None   # per grid: 1 + index of the surface's trim loops, 0 = untrimmed
    trim_uv: Optional[object] = None         # (sum nverts, 2) float32 surface parameters
    _keep: list = field(default_factory=list, repr=False)

    @property
    def n_grids(self):
        return int(len(self.cu))

    @property
    def n_verts(self):
        return int(((self.cu.astype(np.int64) + 1) * (self.cv.astype(np.int64) + 1)).sum())

    @property
    def n_micropolygons(self):
        pts = (self.flags.astype(np.int64) & abi.GRID_POINTS) != 0
        cu, cv = self.cu.astype(np.int64), self.cv.astype(np.int64)
        return int(np.where(pts, cu + 1, cu * cv).sum())

    def is_device(self):
        return hasattr(self.P, "data_ptr")

    def as_struct(self) -> GridBlock:
        b = GridBlock()
        self._keep = []

        def host(a, dtype):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=dtype)
            self._keep.append(a)
            return a.ctypes.data

        def bulk(a, dtype):
            if a is None:
                return None
            if hasattr(a, "data_ptr"):          # torch tensor (device or pinned host)
                assert a.is_contiguous()
                self._keep.append(a)
                return a.data_ptr()
            return host(a, dtype)

        b.n_grids = self.n_grids
        b.cu = host(self.cu, np.int32)
        b.cv = host(self.cv, np.int32)
        b.nkeys = host(self.nkeys, np.int32)
        b.flags = host(self.flags, np.uint32)
        b.lod_bounds = host(self.lod_bounds, np.float32)
        b.key_times = host(self.key_times, np.float32)
        b.P = bulk(self.P, np.float32)
        b.Ci = bulk(self.Ci, np.float32)
        b.Oi = bulk(self.Oi, np.float32)
        b.culled = bulk(self.culled, np.uint8)
        b.aov = bulk(self.aov, np.float32)
        b.Ng = bulk(self.Ng, np.float32)
        b.N = bulk(self.N, np.float32)
        b.radius = bulk(self.radius, np.float32)
        b.csg_node = host(self.csg_node, np.int32)
        b.trim_set = host(self.trim_set, np.int32)
        b.trim_uv = bulk(self.trim_uv, np.float32)
        b.memory_space = 1 if (hasattr(self.P, "is_cuda") and self.P.is_cuda) else 0
        return b

    def to_torch(self, device=None, pin=False):
        """Copy the bulk arrays into torch tensors: CUDA (device given) or pinned host memory."""
        import torch

        def conv(a, dtype):
            if a is None:
                return None
            t = torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)) if not hasattr(a, "data_ptr") else a
            if device is not None:
                return t.to(device, non_blocking=False).contiguous()
            return t.pin_memory() if pin else t

        return GridArrays(cu=self.cu, cv=self.cv, flags=self.flags, P=conv(self.P, np.float32),
                          Ci=conv(self.Ci, np.float32), Oi=conv(self.Oi, np.float32), nkeys=self.nkeys,
                          key_times=self.key_times, lod_bounds=self.lod_bounds, culled=conv(self.culled, np.uint8),
                          aov=conv(self.aov, np.float32), Ng=conv(self.Ng, np.float32), N=conv(self.N, np.float32),
                          radius=conv(self.radius, np.float32), csg_node=self.csg_node,
                          trim_set=self.trim_set, trim_uv=conv(self.trim_uv, np.float32))


class Hider:
    """One hider bound to one CUDA device (AqhHider*)."""

    def __init__(self, device=0, stream=None):
        self._L = lib()
        self._h = C.c_void_p()
        rc = self._L.aqh_create(C.byref(self._h), int(device))
        if rc:
            raise HiderError(rc, "aqh_create failed: the hider needs an sm_100 CUDA device (no CPU fallback)")
        self.params = None
        if stream is not None:
            self.set_stream(stream)

    def close(self):
        if self._h:
            self._L.aqh_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise HiderError(rc, (self._L.aqh_last_error(self._h) or b"").decode(errors="replace"))

    def set_stream(self, cuda_stream_ptr):
        self._check(self._L.aqh_set_stream(self._h, C.c_void_p(int(cuda_stream_ptr))))

    def begin_frame(self, params: FrameParams):
        self.params = params
        self._check(self._L.aqh_begin_frame(self._h, C.byref(params)))
        csg = getattr(params, "_csg", None)          # (types, parents) attached by the scene generators
        if csg:
            self.set_csg_tree(csg[0], csg[1])
        trim = getattr(params, "_trim", None)        # (set_first_loop, loop_first_point, points) attached by the scene generators
        if trim:
            self.set_trim_loops(*trim)

    def add_grid(self, P, cu, cv, Ci=None, Oi=None, flags=abi.GRID_SMOOTH, key_times=None, culled=None, lod_bounds=None,
                 aov=None, Ng=None, N=None, radius=None, csg_node=-1, trim_set=0, trim_uv=None):
        """P: (nkeys, nverts, 3) or (nverts, 3) float32."""
        P = np.ascontiguousarray(P, dtype=np.float32)
        if P.ndim == 2:
            P = P[None]
        nkeys = P.shape[0]
        g = GridDesc()
        g.cu, g.cv, g.nkeys, g.flags = cu, cv, nkeys, flags
        keep = [P]
        ptrs = (C.POINTER(C.c_float) * nkeys)(*[P[k].ctypes.data_as(C.POINTER(C.c_float)) for k in range(nkeys)])
        g.P = ptrs
        if key_times is not None:
            kt = np.ascontiguousarray(key_times, dtype=np.float32)
            keep.append(kt)
            g.key_times = kt.ctypes.data_as(C.POINTER(C.c_float))
        g.csg_node = int(csg_node)
        g.trim_set = int(trim_set)
        if trim_uv is not None:
            tuv = np.ascontiguousarray(trim_uv, dtype=np.float32)
            keep.append(tuv)
            g.trim_uv = tuv.ctypes.data_as(C.POINTER(C.c_float))
        for name, arr, dt in (("Ci", Ci, np.float32), ("Oi", Oi, np.float32), ("culled", culled, np.uint8),
                              ("aov", aov, np.float32), ("Ng", Ng, np.float32), ("N", N, np.float32), ("radius", radius, np.float32)):
            if arr is not None:
                a = np.ascontiguousarray(arr, dtype=dt)
                keep.append(a)
                setattr(g, name, a.ctypes.data_as(C.POINTER(C.c_float if dt == np.float32 else C.c_uint8)))
        if lod_bounds is not None:
            g.lod_bounds[0], g.lod_bounds[1] = lod_bounds
        else:
            g.lod_bounds[0], g.lod_bounds[1] = -1.0, -1.0
        self._check(self._L.aqh_add_grid(self._h, C.byref(g)))

    def add_grid_block(self, grids: GridArrays):
        b = grids.as_struct()
        self._block_keep = getattr(self, "_block_keep", [])
        self._block_keep.append(grids)
        self._check(self._L.aqh_add_grid_block(self._h, C.byref(b)))

    def flush(self):
        """aqh_flush: hide what has been submitted so far and refresh the occlusion image (the frame stays open)."""
        self._check(self._L.aqh_flush(self._h))

    def can_cull(self, bound) -> bool:
        """aqh_can_cull for a raster bound (xmin, ymin, zmin, xmax, ymax, zmax)."""
        b = (C.c_float * 6)(*[float(x) for x in bound])
        out = C.c_int()
        self._check(self._L.aqh_can_cull(self._h, b, C.byref(out)))
        return bool(out.value)

    def render_device(self):
        self._check(self._L.aqh_render_device(self._h))

    def end_frame(self, on_bucket=None, on_data=None, on_progress=None, fetch=True, on_imager=None):
        """aqh_end_frame.  fetch=True returns numpy COPIES of the images (convenience for tests);
        fetch=False leaves them in the library's pinned host buffers (see images(copy=False))."""
        cb = Callbacks()
        keep = []
        if on_bucket:
            def _b(user, x0, x1, y0, y1, ch, stride, nch):
                n = (y1 - y0 - 1) * stride + (x1 - x0) * nch
                a = np.ctypeslib.as_array(ch, shape=(n,))
                rows = [a[r * stride: r * stride + (x1 - x0) * nch].reshape(x1 - x0, nch) for r in range(y1 - y0)]
                return int(on_bucket(x0, x1, y0, y1, np.stack(rows)) or 0)
            cb.on_bucket = abi.BucketFunc(_b)
            keep.append(cb.on_bucket)
        if on_data:
            def _d(user, d, x0, x1, y0, y1, es, data):
                a = np.ctypeslib.as_array(data, shape=((y1 - y0) * (x1 - x0) * es,)).copy()
                return int(on_data(d, x0, x1, y0, y1, es, a) or 0)
            cb.on_data = abi.DataFunc(_d)
            keep.append(cb.on_data)
        if on_progress:
            cb.on_progress = abi.ProgressFunc(lambda user, pct: on_progress(pct))
            keep.append(cb.on_progress)
        if on_imager:
            def _i(user, x0, x1, y0, y1, ch, stride, nch):
                n = (y1 - y0 - 1) * stride + (x1 - x0) * nch
                a = np.ctypeslib.as_array(ch, shape=(n,))
                for r in range(y1 - y0):
                    row = a[r * stride: r * stride + (x1 - x0) * nch].reshape(x1 - x0, nch)     # a view: edits land in place
                    on_imager(x0, x1, y0 + r, row)
                return 0
            cb.on_imager = abi.ImagerFunc(_i)
            keep.append(cb.on_imager)
        self._check(self._L.aqh_end_frame(self._h, C.byref(cb) if keep else None))
        self._block_keep = []
        return self.images() if fetch else None

    def images(self, copy=True):
        """(channels[yres,xres,9] float32, [display arrays]) of the last aqh_end_frame: copies, or with
        copy=False views of the library's pinned host buffers (valid until the next frame)."""
        ptr = C.POINTER(C.c_float)()
        w, h = C.c_int(), C.c_int()
        self._check(self._L.aqh_image_channels(self._h, C.byref(ptr), C.byref(w), C.byref(h)))
        channels = np.ctypeslib.as_array(ptr, shape=(h.value, w.value, self.channel_count()))
        if copy:
            channels = channels.copy()
        displays = []
        for d in range(self.params.n_displays):
            bp = C.POINTER(C.c_ubyte)()
            es, ty = C.c_int(), C.c_int()
            self._check(self._L.aqh_image_display(self._h, d, C.byref(bp), C.byref(es), C.byref(ty)))
            raw = np.ctypeslib.as_array(bp, shape=(h.value * w.value * es.value,))
            if copy:
                raw = raw.copy()
            dt = np.dtype(abi.TYPE_NUMPY[ty.value])
            displays.append(raw.view(dt).reshape(h.value, w.value, es.value // dt.itemsize))
        return channels, displays

    def clear_caches(self):
        self._check(self._L.aqh_clear_caches(self._h))

    def end_frame_capture(self, channels=None, displays=()):
        """aqh_end_frame with the library's own capture display as callbacks (aqh_capture_on_bucket / _on_data): the
        buckets are copied, in reference bucket order, into the given full-frame numpy arrays -- what a framebuffer
        display driver does with DspyImageData.  Returns the number of buckets delivered."""
        cap = abi.Capture()
        cap.xres, cap.yres, cap.n_channels = self.params.xres, self.params.yres, self.channel_count()
        if channels is not None:
            cap.channels = channels.ctypes.data
        for d, a in enumerate(displays):
            if a is not None:
                cap.display[d] = a.ctypes.data
        cb = Callbacks()
        cb.user = C.cast(C.pointer(cap), C.c_void_p)
        cb.on_bucket = C.cast(self._L.aqh_capture_on_bucket, abi.BucketFunc)
        cb.on_data = C.cast(self._L.aqh_capture_on_data, abi.DataFunc)
        self._check(self._L.aqh_end_frame(self._h, C.byref(cb)))
        self._block_keep = []
        return int(cap.buckets)

    def channel_count(self) -> int:
        n = C.c_int()
        self._check(self._L.aqh_channel_count(self._h, C.byref(n)))
        return n.value

    def set_trim_loops(self, set_first_loop, loop_first_point, points):
        a = np.ascontiguousarray(set_first_loop, np.int32)
        b = np.ascontiguousarray(loop_first_point, np.int32)
        c = np.ascontiguousarray(points, np.float32)
        self._check(self._L.aqh_set_trim_loops(self._h, len(a) - 1, a.ctypes.data, b.ctypes.data, c.ctypes.data))

    def set_csg_tree(self, types, parents):
        t = np.ascontiguousarray(types, np.int32)
        pa = np.ascontiguousarray(parents, np.int32)
        self._check(self._L.aqh_set_csg_tree(self._h, len(t), t.ctypes.data, pa.ctypes.data))

    def comm_init(self, unique_id: bytes, rank, world):
        buf = C.create_string_buffer(unique_id, 128)
        self._check(self._L.aqh_comm_init(self._h, buf, int(rank), int(world)))

    def gather(self, root=0):
        self._check(self._L.aqh_gather(self._h, int(root)))

    def stats(self) -> dict:
        s = FrameStats()
        self._check(self._L.aqh_frame_stats(self._h, C.byref(s)))
        return s.as_dict()

    def device_channels(self):
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self._L.aqh_device_channels(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def device_display(self, d):
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self._L.aqh_device_display(self._h, d, C.byref(p), C.byref(n)))
        return p.value, n.value

    def strips(self):
        n = C.c_int()
        self._check(self._L.aqh_num_strips(self._h, C.byref(n)))
        out = []
        for i in range(n.value):
            y0, y1 = C.c_int(), C.c_int()
            self._check(self._L.aqh_strip(self._h, i, C.byref(y0), C.byref(y1)))
            out.append((y0.value, y1.value))
        return out
