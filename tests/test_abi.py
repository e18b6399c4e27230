"""The C-ABI boundary (include/aqsis_b200_hider.h): symbols, struct layout, error behaviour.  CPU only."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from aqsis_b200 import abi, default_params
from aqsis_b200 import build as libbuild

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "aqsis_b200_hider.h")


def header_functions():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"AQH_EXPORT\s+[\w\s\*]+?\b(aqh_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol(native_lib):
    names = header_functions()
    assert len(names) >= 45, names
    out = subprocess.run(["nm", "-D", "--defined-only", libbuild.LIB], capture_output=True, text=True, check=True).stdout
    exported = set(l.split()[-1] for l in out.splitlines() if " T " in l)
    missing = [n for n in names if n not in exported]
    assert not missing, missing
    for n in names:
        assert hasattr(native_lib, n)


def test_library_has_sm100a_kernels_only():
    out = subprocess.run(["cuobjdump", "-lelf", libbuild.LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, out


def test_struct_layout_matches_header(native_lib):
    """Compile a probe against the header with gcc and compare sizeof/offsetof with the ctypes mirror."""
    probes = {
        "AqhDisplayDesc": (abi.DisplayDesc, ["n_channels", "channel", "type", "quantize_zero", "quantize_dither", "flags"]),
        "AqhAovDesc": (abi.AovDesc, ["name", "n_floats"]),
        "AqhFrameParams": (abi.FrameParams, ["abi_version", "xres", "crop_ymax", "filter_func", "bucket_xsize", "shutter_close",
                                             "use_dof", "dof_scale_y", "depth_filter", "zthreshold", "exposure_gamma", "jitter",
                                             "cam_to_raster", "rng_seed", "rng_predraws", "n_displays", "display", "n_aovs", "aov",
                                             "rank", "world_size", "strip_rows", "strip_bounds", "deep_hits_per_sample",
                                             "filter_mode", "plane_budget_mb", "reserved"]),
        "AqhGridDesc": (abi.GridDesc, ["cu", "nkeys", "key_times", "P", "Ci", "Oi", "culled", "flags", "lod_bounds", "aov", "Ng", "N",
                                       "radius", "csg_node", "trim_set", "trim_uv"]),
        "AqhGridBlock": (abi.GridBlock, ["n_grids", "cu", "cv", "nkeys", "flags", "lod_bounds", "key_times", "P", "Ci", "Oi",
                                         "culled", "memory_space", "aov", "Ng", "N", "radius", "csg_node", "trim_set", "trim_uv"]),
        "AqhCallbacks": (abi.Callbacks, ["user", "on_bucket", "on_data", "on_progress", "on_imager"]),
        "AqhFrameStats": (abi.FrameStats, ["prepare_ms", "device_total_ms", "n_grids", "n_deep_hits", "gpu_launches", "d2h_bytes",
                                           "device_bytes", "n_bands", "gather_ms"]),
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for cname, (_, fields) in probes.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for f in fields:
            lines.append(f'printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
    lines.append("return 0;}")
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "p.c"), "w").write("\n".join(lines))
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-o", os.path.join(d, "p"), os.path.join(d, "p.c")], check=True)
        out = subprocess.run([os.path.join(d, "p")], capture_output=True, text=True, check=True).stdout
    got = dict(l.split() for l in out.splitlines())
    for cname, (ct, fields) in probes.items():
        assert int(got[cname]) == C.sizeof(ct), cname
        for f in fields:
            assert int(got[f"{cname}.{f}"]) == getattr(ct, f).offset, (cname, f)


def test_header_is_plain_c():
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(f'#include "{HEADER}"\nint main(void){{return AQH_ABI_VERSION-1;}}\n')
        subprocess.run(["gcc", "-std=c89", "-pedantic", "-Wall", "-c", "-o", os.path.join(d, "t.o"), os.path.join(d, "t.c")], check=True)


def test_defaults_follow_reference_options(native_lib):
    """aqh_frame_params_default mirrors CqOptions' defaults (libs/core/options.cpp:273-305)."""
    p = default_params()
    assert native_lib.aqh_abi_version() == abi.AQH_ABI_VERSION == p.abi_version
    assert (p.xres, p.yres) == (640, 480) and (p.xsamples, p.ysamples) == (2, 2)
    assert (p.filter_xwidth, p.filter_ywidth) == (2.0, 2.0) and (p.bucket_xsize, p.bucket_ysize) == (16, 16)
    assert (p.exposure_gain, p.exposure_gamma) == (1.0, 1.0) and p.rng_seed == 545 and p.jitter == 1
    assert list(p.zthreshold) == [1.0, 1.0, 1.0] and p.depth_filter == abi.DEPTHFILTER_MIN
    assert p.use_dof == 0 and p.shutter_open == 0.0 and p.shutter_close == 0.0


def test_pure_python_parameter_blocks_equal_the_librarys(native_lib):
    """bench.py --impl reference builds its parameter blocks without mapping the product library (hider.PURE): the
    python mirrors of aqh_frame_params_default / aqh_frame_params_set_dof / aqh_display_from_mode must fill the same bytes."""
    from aqsis_b200 import hider, scenes
    made = {}
    for pure in (False, True):
        hider.PURE = pure
        try:
            made[pure] = [scenes.config1(scale=0.1)[0], scenes.config3(scale=0.05)[0], scenes.config4(scale=0.02)[0], default_params()]
        finally:
            hider.PURE = False
    for a, b in zip(made[False], made[True]):
        assert b.filter_func is None and a.filter_func
        a.filter_func = None
        assert bytes(a) == bytes(b)
    d = abi.DisplayDesc()
    for mode, order in [("rgba", 1), ("rgbaz", 0), ("z", 0), ("a", 1), ("rgb", 0)]:
        assert native_lib.aqh_display_from_mode(C.byref(d), mode.encode(), order, 255.0, 0.0, 255.0, 0.5) == 0
        e = abi.DisplayDesc()
        hider._display_from_mode(e, mode, order, 255.0, 0.0, 255.0, 0.5)
        assert bytes(d) == bytes(e)


def test_set_dof_matches_reference_formula(native_lib):
    """CqRenderer::SetDepthOfFieldData, renderer.h:368-377."""
    p = default_params()
    native_lib.aqh_frame_params_set_dof(C.byref(p), 2.8, 0.05, 20.0, 100.0, 120.0)
    lens = np.float32(0.05) / np.float32(2.8)
    want = np.float32(0.5 * np.float64(lens) * 20.0 / (20.0 + np.float64(lens)))
    assert p.use_dof == 1 and np.float32(p.dof_multiplier) == want
    assert np.float32(p.dof_one_over_focal_distance) == np.float32(1.0 / 20.0)
    assert (p.dof_scale_x, p.dof_scale_y) == (100.0, 120.0)
    native_lib.aqh_frame_params_set_dof(C.byref(p), 3.4028234663852886e38, 0.05, 20.0, 1.0, 1.0)
    assert p.use_dof == 0


def test_display_channel_orders(native_lib):
    """Core request order a,r,g,b,z (ddmanager.cpp:455-480) vs the file driver's r,g,b,a (display.cpp:454-490)."""
    d = abi.DisplayDesc()
    assert native_lib.aqh_display_from_mode(C.byref(d), b"rgbaz", 0, 255.0, 0.0, 255.0, 0.5) == 0
    assert list(d.channel[:d.n_channels]) == [abi.CH_ALPHA, abi.CH_CI_R, abi.CH_CI_G, abi.CH_CI_B, abi.CH_Z]
    assert native_lib.aqh_display_from_mode(C.byref(d), b"rgba", 1, 255.0, 0.0, 255.0, 0.5) == 0
    assert list(d.channel[:d.n_channels]) == [abi.CH_CI_R, abi.CH_CI_G, abi.CH_CI_B, abi.CH_ALPHA]
    assert native_lib.aqh_display_from_mode(C.byref(d), b"q", 1, 255.0, 0.0, 255.0, 0.5) == abi.AQH_ERR_BAD_PARAMS


def test_select_data_format():
    """selectDataFormat, ddmanager.cpp:249-283 (python mirror used by the tests == the library's)."""
    from aqsis_b200 import display_info
    cases = [((0, 0, 0), np.float32), ((255, 0, 255), np.uint8), ((65535, 0, 65535), np.uint16), ((1e9, 0, 4e9), np.uint32),
             ((127, -128, 127), np.int8), ((1000, -32768, 32767), np.int16), ((1e6, -1e6, 1e6), np.int32)]
    for (one, mn, mx), dt in cases:
        p = default_params(displays=[("rgb", 1, one, mn, mx, 0.0)])
        assert display_info(p, 0)[0] == np.dtype(dt)


def test_no_device_means_no_result(native_lib):
    """There is no CPU fallback: without a CUDA device the hider cannot even be created."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    assert native_lib.aqh_create(C.byref(h), 0) == abi.AQH_ERR_NO_DEVICE
    assert not h.value
    from aqsis_b200 import Hider, HiderError
    with pytest.raises(HiderError):
        Hider(0)


def test_product_never_references_the_oracle():
    """Nothing under aqsis_b200/ or include/ may import, link or load oracle/."""
    bad = []
    for base in ("aqsis_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith((".py", ".cpp", ".cu", ".h")):
                    txt = open(os.path.join(dp, f), errors="replace").read()
                    if re.search(r"oracle_hider|liboracle|import orc\b|from orc\b|oracle/", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
    out = subprocess.run(["ldd", libbuild.LIB], capture_output=True, text=True).stdout
    assert "oracle" not in out
