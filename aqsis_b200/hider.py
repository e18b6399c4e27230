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
    _LIB = L
    return L


FILTER_NAMES = ("box", "triangle", "gaussian", "catmull-rom", "sinc", "mitchell", "disk", "bessel")


def default_params(**kw) -> FrameParams:
    """AqhFrameParams with the reference's option defaults, then overrides.

    Convenience keys: resolution=(x,y) also resets the crop window; samples=(xs,ys);
    filter=("name", xw, yw); displays=[("rgba", driver_order, one, min, max, dither), ...].
    """
    p = FrameParams()
    lib().aqh_frame_params_default(C.byref(p))
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
        fn = lib().aqh_filter_by_name(name.encode())
        if not fn:
            raise ValueError(f"unknown pixel filter {name!r}")
        p.filter_func = fn
        p.filter_xwidth, p.filter_ywidth = xw, yw
    if "dof" in kw:
        fstop, fl, fd, sx, sy = kw.pop("dof")
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
            rc = lib().aqh_display_from_mode(C.byref(p.display[i]), mode.encode(), int(driver_order), one, mn, mx, dither)
            if rc:
                raise HiderError(rc, f"display mode {mode!r}")
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(f"AqhFrameParams has no field {k!r}")
        setattr(p, k, v)
    return p


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
    _keep: list = field(default_factory=list, repr=False)

    @property
    def n_grids(self):
        return int(len(self.cu))

    @property
    def n_verts(self):
        return int(((self.cu.astype(np.int64) + 1) * (self.cv.astype(np.int64) + 1)).sum())

    @property
    def n_micropolygons(self):
        return int((self.cu.astype(np.int64) * self.cv.astype(np.int64)).sum())

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
                          key_times=self.key_times, lod_bounds=self.lod_bounds, culled=conv(self.culled, np.uint8))


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

    def add_grid(self, P, cu, cv, Ci=None, Oi=None, flags=abi.GRID_SMOOTH, key_times=None, culled=None, lod_bounds=None):
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
        for name, arr, dt in (("Ci", Ci, np.float32), ("Oi", Oi, np.float32), ("culled", culled, np.uint8)):
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

    def end_frame(self, on_bucket=None, on_data=None, on_progress=None, fetch=True):
        """aqh_end_frame.  fetch=True returns numpy COPIES of the images (convenience for tests);
        fetch=False leaves them in the library's pinned host buffers (see images(copy=False))."""
        cb = Callbacks()
        keep = []
        if on_bucket:
            def _b(user, x0, x1, y0, y1, ch, stride):
                n = (y1 - y0 - 1) * stride + (x1 - x0) * 9
                a = np.ctypeslib.as_array(ch, shape=(n,))
                rows = [a[r * stride: r * stride + (x1 - x0) * 9].reshape(x1 - x0, 9) for r in range(y1 - y0)]
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
        self._check(self._L.aqh_end_frame(self._h, C.byref(cb) if keep else None))
        self._block_keep = []
        return self.images() if fetch else None

    def images(self, copy=True):
        """(channels[yres,xres,9] float32, [display arrays]) of the last aqh_end_frame: copies, or with
        copy=False views of the library's pinned host buffers (valid until the next frame)."""
        ptr = C.POINTER(C.c_float)()
        w, h = C.c_int(), C.c_int()
        self._check(self._L.aqh_image_channels(self._h, C.byref(ptr), C.byref(w), C.byref(h)))
        channels = np.ctypeslib.as_array(ptr, shape=(h.value, w.value, 9))
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
