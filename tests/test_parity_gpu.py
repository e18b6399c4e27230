"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

from aqsis_b200 import scenes, default_params, abi
import parity_util as pu

pytestmark = pytest.mark.gpu

EXACT = abi.FILTER_REFERENCE_ORDER     # bit-identical float sums (reference summation order)
TILED = abi.FILTER_TILE_PARTIALS       # default: same samples and weights, partial sums per (pixel, tap)
MODES = [pytest.param(EXACT, id="reference-order"), pytest.param(TILED, id="tile-partials")]


def check(h, p, g, mode, exact_frac=0.999, **kw):
    """Parity at the stated tolerance; in reference-order mode additionally (almost) bit-exact."""
    p.filter_mode = mode
    if mode == TILED:
        # different association of the filter sums: rounding noise only, but negative-lobed
        # filters amplify it on dark pixels (documented in include/aqsis_b200_hider.h)
        kw.setdefault("float_rtol", 5e-4)
        kw.setdefault("strict_special", False)
    r = pu.parity(h, p, g, **kw)
    if mode == EXACT:
        assert r["float_bit_exact_frac"] > exact_frac, r
    return r


@pytest.mark.parametrize("mode", MODES)
def test_config1_small_static(gpu_hider, mode):
    p, g = scenes.config1(scale=0.2)
    r = check(gpu_hider, p, g, mode)
    assert r["gpu_stats"]["gpu_launches"] >= 6


@pytest.mark.parametrize("mode", MODES)
def test_config2_small_static(gpu_hider, mode):
    p, g = scenes.config2(scale=0.08)
    check(gpu_hider, p, g, mode)


def test_add_grid_matches_block(gpu_hider):
    p, g = scenes.config1(scale=0.12)
    ch_a, disp_a, _ = pu.run_product(gpu_hider, p, g, use_block=True)
    ch_b, disp_b, _ = pu.run_product(gpu_hider, p, g, use_block=False)
    assert np.array_equal(ch_a, ch_b) and np.array_equal(disp_a[0], disp_b[0])


@pytest.mark.parametrize("mode", MODES)
def test_config3_small_mb_dof(gpu_hider, mode):
    p, g = scenes.config3(scale=0.05, motion_px=6.0)
    check(gpu_hider, p, g, mode, exact_frac=0.99)


@pytest.mark.parametrize("mode", MODES)
def test_motion_only(gpu_hider, mode):
    p, g = scenes.config3(scale=0.05, motion_px=8.0)
    p.use_dof = 0
    check(gpu_hider, p, g, mode, exact_frac=0.99)


@pytest.mark.parametrize("mode", MODES)
def test_dof_only(gpu_hider, mode):
    p, g = scenes.config2(scale=0.05)
    import ctypes as C
    from aqsis_b200 import lib
    lib().aqh_frame_params_set_dof(C.byref(p), 2.8, 0.05, 20.0, 60.0, 60.0)
    check(gpu_hider, p, g, mode, exact_frac=0.99)


@pytest.mark.parametrize("mode", MODES)
def test_config4_small_transparent(gpu_hider, mode):
    p, g = scenes.config4(scale=0.02)
    r = check(gpu_hider, p, g, mode, exact_frac=0.99)
    assert r["gpu_stats"]["n_deep_hits"] > 0


@pytest.mark.parametrize("name,width", [("box", 1.0), ("triangle", 2.0), ("gaussian", 3.0), ("catmull-rom", 4.0),
                                        ("sinc", 5.0), ("sinc", 6.0), ("gaussian", 2.5), ("mitchell", 4.0)])
@pytest.mark.parametrize("mode", MODES)
def test_filter_sweep_small(gpu_hider, name, width, mode):
    p, g = scenes.config2(scale=0.04, filter=(name, width, width), samples=(4, 4))
    check(gpu_hider, p, g, mode)


@pytest.mark.parametrize("mode", MODES)
def test_crop_window_and_odd_resolution(gpu_hider, mode):
    p, g = scenes.config1(scale=0.15)
    p.crop_xmin, p.crop_xmax, p.crop_ymin, p.crop_ymax = 7, p.xres - 5, 3, p.yres - 9
    check(gpu_hider, p, g, mode)


@pytest.mark.parametrize("samples", [(1, 1), (2, 3), (5, 5), (16, 16)])
def test_odd_sample_counts(gpu_hider, samples):
    p, g = scenes.config2(scale=0.03, samples=samples, filter=("gaussian", 2.0, 2.0))
    check(gpu_hider, p, g, EXACT)
    check(gpu_hider, p, g, TILED)


def test_empty_frame(gpu_hider):
    p = default_params(resolution=(40, 24), samples=(2, 2), displays=[("rgba", 1, 255.0, 0.0, 255.0, 0.5)])
    gpu_hider.begin_frame(p)
    ch, disp = gpu_hider.end_frame()
    assert np.all(ch[..., :7] == 0) and np.all(ch[..., abi.CH_Z] == np.float32(3.4028234663852886e38))
    assert disp[0].max() <= 1      # dither only
