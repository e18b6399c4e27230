"""ctypes face of the CPU oracle (oracle/oracle_hider.h).  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (aqsis_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from aqsis_b200 import _abi as abi
from aqsis_b200._abi import FrameParams, GridBlock, DisplayDesc

FILTER_INDEX = {"box": 0, "triangle": 1, "gaussian": 2, "catmull-rom": 3, "sinc": 4, "mitchell": 5, "disk": 6, "bessel": 7}


def _csg_arrays(params):
    """(n, types, parents) of the frame's CSG tree: carried on the python parameter object as params._csg = (types, parents)
    (the C struct has no room for it: the product takes it through aqh_set_csg_tree)."""
    csg = getattr(params, "_csg", None)
    if not csg:
        return 0, None, None
    t = np.ascontiguousarray(csg[0], np.int32)
    pa = np.ascontiguousarray(csg[1], np.int32)
    return len(t), t, pa


def _trim_arrays(params):
    """(n_sets, set_first_loop, loop_first_point, points) of the frame's trim loops: carried on the python parameter object
    as params._trim (the product takes them through aqh_set_trim_loops)."""
    trim = getattr(params, "_trim", None)
    if not trim:
        return 0, None, None, None
    a = np.ascontiguousarray(trim[0], np.int32)
    b = np.ascontiguousarray(trim[1], np.int32)
    c = np.ascontiguousarray(trim[2], np.float32)
    return len(a) - 1, a, b, c


def _filter_name_of(params):
    """The pixel filter to select by NAME: only when the frame carries no function pointer (pure-python parameter
    blocks of bench.py --impl reference); otherwise the checkers recognise the function behind the pointer."""
    if params.filter_func:
        return None
    return getattr(params, "_filter_name", None) or "gaussian"

class OrcStats(C.Structure):
    """OrcStats of oracle/oracle_hider.h."""
    _fields_ = [
        ("prepare_s", C.c_double), ("bust_s", C.c_double), ("render_s", C.c_double), ("combine_s", C.c_double),
        ("filter_s", C.c_double), ("display_s", C.c_double), ("total_s", C.c_double),
        ("n_micropolygons", C.c_int64), ("n_bucket_entries", C.c_int64), ("n_samples", C.c_int64),
        ("spl_count", C.c_int64), ("spl_bound_hits", C.c_int64), ("spl_hits", C.c_int64),
        ("n_deep_hits", C.c_int64), ("threads", C.c_int32),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "_build", "liboracle_hider.so")
REF_LIB = os.path.join(ORACLE_DIR, "_ref", "libaqsis_refleaf.so")
REFHIDER_LIB = os.path.join(ORACLE_DIR, "_ref", "libaqsis_refhider.so")
_lib = None
_ref = None
_refhider = None


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "oracle"], check=True, capture_output=True)
    return ORACLE_LIB


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(ORACLE_DIR, "oracle_hider.cpp")
        if not os.path.exists(ORACLE_LIB) or os.path.getmtime(ORACLE_LIB) < os.path.getmtime(src):
            build_oracle()
        L = C.CDLL(ORACLE_LIB)
        vp, ci, cf = C.c_void_p, C.c_int, C.c_float
        L.orc_render.argtypes = [C.POINTER(FrameParams), C.POINTER(GridBlock), vp, C.POINTER(vp), ci, C.POINTER(OrcStats)]
        L.orc_display_entrysize.argtypes = [C.POINTER(DisplayDesc), C.POINTER(ci)]
        L.orc_random_reseed.argtypes = [C.c_uint32]
        L.orc_random_reseed.restype = None
        L.orc_random_uint.restype = C.c_uint32
        L.orc_random_float.restype = cf
        L.orc_random_int.argtypes = [C.c_uint32]
        L.orc_random_int.restype = C.c_uint32
        L.orc_sampler_tables.argtypes = [ci, ci, ci, vp, vp, vp]
        L.orc_filter.argtypes = [ci, cf, cf, cf, cf]
        L.orc_filter.restype = cf
        L.orc_invbilinear.argtypes = [vp, cf, cf, vp]
        L.orc_invbilinear.restype = None
        L.orc_bilerp.argtypes = [cf] * 6
        L.orc_bilerp.restype = cf
        L.orc_filter_table.argtypes = [C.POINTER(FrameParams), vp]
        L.orc_replay.argtypes = [C.POINTER(FrameParams), vp, vp, C.POINTER(ci), C.POINTER(ci), C.POINTER(ci), C.POINTER(ci)]
        L.orc_dof_bounds.argtypes = [ci, ci, vp]
        L.orc_dof_bounds.restype = None
        L.orc_set_filter.argtypes = [ci]
        L.orc_set_filter.restype = None
        L.orc_set_csg_tree.argtypes = [ci, vp, vp]
        L.orc_set_trim_loops.argtypes = [ci, vp, vp, vp]
        L.orc_trim_point.argtypes = [ci, C.c_float, C.c_float]
        L.orc_trim_line.argtypes = [ci, C.c_float, C.c_float, C.c_float, C.c_float]
        _lib = L
    return _lib


def ref():
    """The reference's own leaf sources compiled in place (None when oracle/_ref is absent)."""
    global _ref
    if _ref is None:
        if not os.path.exists(REF_LIB):
            if os.path.isdir("/root/reference/libs/core"):
                subprocess.run(["make", "-s", "-C", ORACLE_DIR, "ref"], check=True, capture_output=True)
            if not os.path.exists(REF_LIB):
                return None
        L = C.CDLL(REF_LIB)
        vp, ci, cf = C.c_void_p, C.c_int, C.c_float
        L.ref_random_reseed.argtypes = [C.c_uint]
        L.ref_random_reseed.restype = None
        L.ref_random_uint.restype = C.c_uint
        L.ref_random_float.restype = cf
        L.ref_random_int.argtypes = [C.c_uint]
        L.ref_random_int.restype = C.c_uint
        L.ref_sampler_create.argtypes = [ci, ci, ci]
        L.ref_sampler_create.restype = vp
        L.ref_sampler_destroy.argtypes = [vp]
        L.ref_sampler_destroy.restype = None
        L.ref_sampler_draw.argtypes = [vp] * 6
        L.ref_sampler_draw.restype = None
        L.ref_filter.argtypes = [ci, cf, cf, cf, cf]
        L.ref_filter.restype = cf
        L.ref_invbilinear.argtypes = [vp, cf, cf, vp]
        L.ref_invbilinear.restype = None
        L.ref_bilerp.argtypes = [cf] * 6
        L.ref_bilerp.restype = cf
        L.ref_lfloor.argtypes = [C.c_double]
        L.ref_lfloor.restype = C.c_long
        L.ref_lceil.argtypes = [C.c_double]
        L.ref_lceil.restype = C.c_long
        L.ref_lround.argtypes = [C.c_double]
        L.ref_lround.restype = C.c_long
        L.ref_bound_contains2d.argtypes = [vp, cf, cf]
        L.ref_bound_intersects.argtypes = [vp, cf, cf, cf, cf]
        _ref = L
    return _ref


def render(params: FrameParams, grids, nthreads=1):
    """Run the oracle. Returns (channels[yres,xres,9], [display arrays], stats dict)."""
    from aqsis_b200.hider import display_info
    L = lib()
    b = grids.as_struct()
    assert b.memory_space == 0
    ch = np.zeros((params.yres, params.xres, 9 + params.aov_floats), dtype=np.float32)
    outs, ptrs = [], (C.c_void_p * max(1, params.n_displays))()
    for d in range(params.n_displays):
        dt, nch, es = display_info(params, d)
        a = np.zeros((params.yres, params.xres, nch), dtype=dt)
        outs.append(a)
        ptrs[d] = a.ctypes.data
    st = OrcStats()
    name = _filter_name_of(params)
    L.orc_set_filter(FILTER_INDEX[name] if name else -1)
    n, t, pa = _csg_arrays(params)
    L.orc_set_csg_tree(n, t.ctypes.data if n else None, pa.ctypes.data if n else None)
    nt, ta, tb, tc = _trim_arrays(params)
    L.orc_set_trim_loops(nt, ta.ctypes.data if nt else None, tb.ctypes.data if nt else None, tc.ctypes.data if nt else None)
    rc = L.orc_render(C.byref(params), C.byref(b), ch.ctypes.data, ptrs, int(nthreads), C.byref(st))
    L.orc_set_filter(-1)
    L.orc_set_csg_tree(0, None, None)
    L.orc_set_trim_loops(0, None, None, None)
    if rc:
        raise RuntimeError(f"orc_render failed: {abi.STATUS_NAMES.get(rc, rc)}")
    return ch, outs, st.as_dict()


def refhider():
    """The reference's own hider compiled in place (oracle/ref_hider.cpp); None when oracle/_ref lacks it."""
    global _refhider
    if _refhider is None:
        if os.path.isdir("/root/reference/libs/core"):     # build container: (re)build in place when stale
            subprocess.run(["make", "-s", "-C", ORACLE_DIR, "refhider"], check=True, capture_output=True)
        if not os.path.exists(REFHIDER_LIB):
            return None
        L = C.CDLL(REFHIDER_LIB)
        L.ref_render.argtypes = [C.POINTER(FrameParams), C.POINTER(GridBlock), C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(OrcStats)]
        L.ref_set_filter.argtypes = [C.c_char_p]
        L.ref_set_csg_tree.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.ref_can_cull.argtypes = [C.POINTER(FrameParams), C.POINTER(GridBlock), C.c_int, C.c_void_p, C.c_void_p]
        L.ref_set_trim_loops.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_trim_point.argtypes = [C.c_int, C.c_float, C.c_float, ]
        L.ref_trim_line.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float, C.c_float]
        _refhider = L
    return _refhider


def render_reference(params: FrameParams, grids):
    """Run the reference's own hider (single-threaded, like aqsis).  Same returns as render()."""
    from aqsis_b200.hider import display_info
    L = refhider()
    if L is None:
        raise RuntimeError("oracle/_ref/libaqsis_refhider.so is not available")
    b = grids.as_struct()
    ch = np.zeros((params.yres, params.xres, 9 + params.aov_floats), dtype=np.float32)
    outs, ptrs = [], (C.c_void_p * max(1, params.n_displays))()
    for d in range(params.n_displays):
        dt, nch, es = display_info(params, d)
        a = np.zeros((params.yres, params.xres, nch), dtype=dt)
        outs.append(a)
        ptrs[d] = a.ctypes.data
    st = OrcStats()
    name = _filter_name_of(params)
    L.ref_set_filter(name.encode() if name else None)
    n, t, pa = _csg_arrays(params)
    L.ref_set_csg_tree(n, t.ctypes.data if n else None, pa.ctypes.data if n else None)
    nt, ta, tb, tc = _trim_arrays(params)
    L.ref_set_trim_loops(nt, ta.ctypes.data if nt else None, tb.ctypes.data if nt else None, tc.ctypes.data if nt else None)
    rc = L.ref_render(C.byref(params), C.byref(b), ch.ctypes.data, ptrs, C.byref(st))
    L.ref_set_filter(None)
    L.ref_set_csg_tree(0, None, None)
    L.ref_set_trim_loops(0, None, None, None)
    if rc:
        raise RuntimeError(f"ref_render failed: {abi.STATUS_NAMES.get(rc, rc)}")
    return ch, outs, st.as_dict()


def reference_can_cull(params: FrameParams, grids, bounds):
    """CqOcclusionTree::canCull of the reference's own hider for raster bounds (n, 6) = xmin ymin zmin xmax ymax zmax,
    asked once all of `grids` is rendered: True where every bucket's tree culls the bound (ref_hider.cpp: ref_can_cull)."""
    L = refhider()
    if L is None:
        raise RuntimeError("oracle/_ref/libaqsis_refhider.so is not available")
    b = grids.as_struct()
    bb = np.ascontiguousarray(bounds, np.float32).reshape(-1, 6)
    out = np.zeros(len(bb), np.uint8)
    name = _filter_name_of(params)
    L.ref_set_filter(name.encode() if name else None)
    rc = L.ref_can_cull(C.byref(params), C.byref(b), len(bb), bb.ctypes.data, out.ctypes.data)
    L.ref_set_filter(None)
    if rc:
        raise RuntimeError(f"ref_can_cull failed: {abi.STATUS_NAMES.get(rc, rc)}")
    return out.astype(bool)


_pdiff = None


def pdiff_lib():
    """The reference's perceptual diff (thirdparty/pdiff) compiled in place (oracle/ref_pdiff.cpp); None when absent."""
    global _pdiff
    if _pdiff is None:
        if os.path.isdir("/root/reference/thirdparty/pdiff"):
            subprocess.run(["make", "-s", "-C", ORACLE_DIR, "pdiff"], check=True, capture_output=True)
        path = os.path.join(ORACLE_DIR, "_ref", "libaqsis_pdiff.so")
        if not os.path.exists(path):
            _pdiff = False
        else:
            L = C.CDLL(path)
            L.ref_pdiff.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                    C.c_uint, C.POINTER(C.c_int), C.POINTER(C.c_int)]
            L.ref_pdiff.restype = C.c_int
            _pdiff = L
    return _pdiff or None


def pdiff(rgba_a, rgba_b, threshold_pixels=1, fov=45.0, gamma=2.2, luminance=100.0):
    """Yee's perceptual metric as aqsis' regression tool runs it (defaults of CompareArgs.cpp), with the pixel
    threshold at 1 = "zero pdiff-detected differences".  Returns (passed, failing pixels, binary identical)."""
    L = pdiff_lib()
    if L is None:
        raise RuntimeError("oracle/_ref/libaqsis_pdiff.so is not available")
    a = np.ascontiguousarray(rgba_a, dtype=np.uint8)
    b = np.ascontiguousarray(rgba_b, dtype=np.uint8)
    assert a.shape == b.shape and a.ndim == 3 and a.shape[2] == 4
    failed, same = C.c_int(), C.c_int()
    rc = L.ref_pdiff(a.ctypes.data, b.ctypes.data, a.shape[1], a.shape[0], fov, gamma, luminance, int(threshold_pixels),
                     C.byref(failed), C.byref(same))
    if rc < 0:
        raise ValueError("ref_pdiff: bad arguments")
    return bool(rc), int(failed.value), bool(same.value)
