"""aqsis mpdump files (libs/core/mpdump.cpp): reader / writer, replay of dumped micropolygons through the hider
seam, and the sample records against the library's replay of the global random stream."""
import struct

import numpy as np
import pytest

import orc
from aqsis_b200 import abi, mpdump, scenes


def _constant(g):
    """Constant shading (the dump keeps ONE colour per micropolygon: the value at its index)."""
    g.flags = np.zeros_like(g.flags)
    return g


def test_record_layout_matches_the_reference_writer(tmp_path):
    # hand-packed exactly as CqMPDump writes (mpdump.cpp:44-56, 72-90, 117-180)
    P = [(1, 2, 3), (4, 5, 6), (10, 11, 12), (7, 8, 9)]            # natural order P0 P1 P2 P3
    raw = struct.pack("<i", 4)
    raw += struct.pack("<hii", 3, 640, 480)
    raw += struct.pack("<hiiiff", 2, 5, 7, 3, 5.25, 7.75)
    raw += struct.pack("<h", 1)
    for v in (P[0], P[1], P[3], P[2]):                             # the circular order of the file
        raw += struct.pack("<fff", *v)
    raw += struct.pack("<fff", 0.1, 0.2, 0.3) + struct.pack("<fff", 1.0, 0.5, 0.25)
    f = tmp_path / "mpdump.mp"
    f.write_bytes(raw)
    d = mpdump.read(f)
    assert (d.width, d.height) == (640, 480) and d.n_micropolygons == 1
    assert np.array_equal(d.P[0], np.float32(P)) and np.allclose(d.Ci[0], [0.1, 0.2, 0.3]) and np.array_equal(d.Oi[0], np.float32([1.0, 0.5, 0.25]))
    s = d.samples[0]
    assert (s["x"], s["y"], s["idx"], s["px"], s["py"]) == (5, 7, 3, 5.25, 7.75)
    g = tmp_path / "copy.mp"
    mpdump.write(g, d)
    assert g.read_bytes() == raw
    with pytest.raises(ValueError):
        (tmp_path / "bad.mp").write_bytes(struct.pack("<i", 8))
        mpdump.read(tmp_path / "bad.mp")


def test_dump_replay_renders_the_same_image(tmp_path):
    """Grids -> busted micropolygons -> file -> 1x1 grids in dump order: the same frame bit for bit (same samples,
    same submission order, constant shading), on the oracle and on the reference's own hider."""
    p, g = scenes.config1(scale=0.12)
    g = _constant(g)
    rng = np.random.default_rng(3)
    g.culled = (rng.uniform(size=g.n_verts) < 0.04).astype(np.uint8)
    d = mpdump.from_grids(g, p.xres, p.yres)
    assert 0 < d.n_micropolygons < g.n_micropolygons
    mpdump.write(tmp_path / "mpdump.mp", d)
    d2 = mpdump.read(tmp_path / "mpdump.mp")
    assert np.array_equal(d2.P, d.P) and np.array_equal(d2.Ci, d.Ci) and (d2.width, d2.height) == (p.xres, p.yres)
    ch_a, disp_a, st_a = orc.render(p, g, 2)
    ch_b, disp_b, st = orc.render(p, mpdump.to_grids(d2), 2)
    assert st["n_micropolygons"] == st_a["n_micropolygons"] <= d.n_micropolygons      # the count after the crop-window reject
    assert np.array_equal(ch_a.view(np.uint32), ch_b.view(np.uint32)) and np.array_equal(disp_a[0], disp_b[0])
    if orc.refhider() is not None:
        ch_r, disp_r, _ = orc.render_reference(p, mpdump.to_grids(d2))
        assert np.array_equal(ch_a.view(np.uint32), ch_r.view(np.uint32)) and np.array_equal(disp_a[0], disp_r[0])


def test_expected_sample_records_follow_the_replayed_stream():
    """The id-2 records the reference would dump, from aqh_replay_frame_rng + aqh_sampler_tables: the known-answer
    pattern choices of SURVEY.md appendix B (first pixel: position table 107) and positions inside their pixels."""
    p, _ = scenes.config1(scale=0.1)
    s = mpdump.expected_samples(p)
    n = p.xsamples * p.ysamples
    assert len(s) % n == 0 and np.array_equal(s["idx"][:n], np.arange(n))
    assert np.all(np.floor(s["px"]) == s["x"]) and np.all(np.floor(s["py"]) == s["y"])
    assert np.all(np.diff(s["y"][::n]) >= -1)           # row-major over the sample region
    first = s[:n]
    # Reseed(545), CqMultiJitteredSampler(4,4): table 107 starts (0.0362413637, 0.225688934), (0.487524688, 0.0274800472)
    assert np.allclose(first["px"][:2] - first["x"][:2], [0.0362413637, 0.487524688], atol=2e-7)
    assert np.allclose(first["py"][:2] - first["y"][:2], [0.225688934, 0.0274800472], atol=2e-7)


@pytest.mark.gpu
def test_dump_replay_on_the_gpu(gpu_hider, tmp_path):
    import parity_util as pu
    p, g = scenes.config1(scale=0.12)
    g = _constant(g)
    d = mpdump.from_grids(g, p.xres, p.yres)
    mpdump.write(tmp_path / "mpdump.mp", d)
    replay = mpdump.to_grids(mpdump.read(tmp_path / "mpdump.mp"))
    ch_g, disp_g, st = pu.run_product(gpu_hider, p, replay)
    ch_o, disp_o, _ = orc.render(p, g, 4)
    assert 0 < st["n_micropolygons"] <= d.n_micropolygons
    assert np.array_equal(ch_g.view(np.uint32), ch_o.view(np.uint32)) and np.array_equal(disp_g[0], disp_o[0])
