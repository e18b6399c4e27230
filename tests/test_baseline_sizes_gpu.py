"""GPU parity AT THE SIZES BASELINE.json NAMES (through the C ABI, `-m gpu`).

The small-frame parity tests (tests/test_parity_gpu.py) cannot reach what only exists at size: more than 64 k tiles,
tens of millions of bin entries, 32-bit position packing, the chunked n > 64 sample-plane layout at 4K widths, bands
of tile rows.  Here:

  * config 1 at FULL size against the reference's own hider (oracle/_ref/libaqsis_refhider.so, when it travelled);
  * config 2 at FULL size (1920x1080, 8x8, 20 M micropolygons) against the multi-threaded oracle;
  * configs 3 and 4 at the largest scale the oracle finishes in about a minute on the box's host cores;
  * every one of the 30 PixelFilter combinations of config 5 at 8x8 samples per pixel against the oracle.

All bit for bit (floats and quantised bytes): the default filter mode sums in the reference's order.
"""
import os
import time

import numpy as np
import pytest

import orc
import parity_util as pu
from aqsis_b200 import abi, scenes

pytestmark = pytest.mark.gpu
THREADS = os.cpu_count() or 8


def assert_same_bits(p, ch_g, d_g, ch_r, d_r, what):
    ys, xs = slice(p.crop_ymin, p.crop_ymax), slice(p.crop_xmin, p.crop_xmax)
    a, b = ch_g[ys, xs].view(np.uint32), ch_r[ys, xs].view(np.uint32)
    assert np.array_equal(a, b), (what, "float channel buffer differs", float((a == b).mean()))
    for x, y in zip(d_g, d_r):
        assert np.array_equal(x[ys, xs], y[ys, xs]), (what, "quantised bytes differ")


def test_config1_full_size_vs_reference_hider(gpu_hider):
    p, g = scenes.config1()                                     # 640x480, 4x4, gaussian 2x2, 640 k micropolygons
    ch_g, d_g, st = pu.run_product(gpu_hider, p, g)
    if orc.refhider() is not None:
        ch_r, d_r, _ = orc.render_reference(p, g)               # aqsis' own hider, about 2 s
        assert_same_bits(p, ch_g, d_g, ch_r, d_r, "config 1 vs aqsis' hider")
    ch_o, d_o, _ = orc.render(p, g, THREADS)
    assert_same_bits(p, ch_g, d_g, ch_o, d_o, "config 1 vs oracle")
    assert st["n_micropolygons"] > 600000


def test_config2_full_size_vs_oracle(gpu_hider):
    p, g = scenes.config2()                                     # the bench workload itself
    ch_g, d_g, st = pu.run_product(gpu_hider, p, g)
    t0 = time.time()
    ch_o, d_o, _ = orc.render(p, g, THREADS)
    print(f"oracle: full config 2 on {THREADS} threads in {time.time() - t0:.1f} s")
    assert_same_bits(p, ch_g, d_g, ch_o, d_o, "config 2 full size")
    assert st["n_micropolygons"] > 19_000_000 and st["n_bin_entries"] > 25_000_000


def test_config3_half_scale_vs_oracle(gpu_hider):
    p, g = scenes.config3(scale=0.5)                            # 960x540, 8x8, motion blur + depth of field, 5 M micropolygons
    ch_g, d_g, st = pu.run_product(gpu_hider, p, g)
    t0 = time.time()
    ch_o, d_o, _ = orc.render(p, g, THREADS)
    print(f"oracle: config 3 at scale 0.5 on {THREADS} threads in {time.time() - t0:.1f} s")
    assert_same_bits(p, ch_g, d_g, ch_o, d_o, "config 3 scale 0.5")


def test_config4_large_scale_vs_oracle(gpu_hider):
    p, g = scenes.config4(scale=0.3)                            # 1152x648, 16x16, 4 layers: 12 M micropolygons, 191 M samples
    ch_g, d_g, st = pu.run_product(gpu_hider, p, g)
    t0 = time.time()
    ch_o, d_o, _ = orc.render(p, g, THREADS)
    print(f"oracle: config 4 at scale 0.3 on {THREADS} threads in {time.time() - t0:.1f} s")
    assert_same_bits(p, ch_g, d_g, ch_o, d_o, "config 4 scale 0.3")
    assert st["n_deep_hits"] > 0


def test_config4_banded_is_the_same_image(gpu_hider):
    """AqhFrameParams::plane_budget_mb: the frame is hidden and filtered in bands of tile rows that fit the budget;
    the image does not depend on the band height."""
    p, g = scenes.config4(scale=0.1)
    ch_a, d_a, st_a = pu.run_product(gpu_hider, p, g)
    for mb in (64, 17):
        p.plane_budget_mb = mb
        ch_b, d_b, st_b = pu.run_product(gpu_hider, p, g)
        assert st_b["n_bands"] > st_a["n_bands"]
        assert np.array_equal(ch_a.view(np.uint32), ch_b.view(np.uint32)) and np.array_equal(d_a[0], d_b[0])


@pytest.mark.parametrize("name,width", scenes.config5_filters())
def test_config5_every_filter_at_8x8(gpu_hider, name, width):
    """The PixelFilter sweep of config 5 (box, triangle, gaussian, catmull-rom, sinc at widths 1-6) at the config's own
    8x8 samples per pixel: 64 .. 3136 taps per pixel through the span-staged filter kernel."""
    p, g = scenes.config2(scale=0.1, filter=(name, width, width), samples=(8, 8))
    ch_g, d_g, _ = pu.run_product(gpu_hider, p, g)
    ch_o, d_o, _ = orc.render(p, g, THREADS)
    assert_same_bits(p, ch_g, d_g, ch_o, d_o, f"{name} {width}")
