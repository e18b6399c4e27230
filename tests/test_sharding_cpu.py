"""The N>1 path on CPU: strip dealing, grid replication and the final gather over gloo (world_size 2 and 3).

Each rank runs the ORACLE on only the grids sharding.split_grids_for_rank hands it and keeps only the
pixel rows it owns; the rows gathered on rank 0 must equal a single-process oracle render of the whole
frame bit for bit.  That proves (i) straddling grids are replicated to every rank that needs them,
halo and depth-of-field growth included, (ii) the strips partition the image, (iii) the gather puts
every row in place.  On the GPU box the same sharding module drives the CUDA hider over NCCL.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _scene(kind):
    from aqsis_b200 import scenes
    if kind == "static":
        p, g = scenes.config1(scale=0.2)
    elif kind == "mbdof":
        p, g = scenes.config3(scale=0.05, motion_px=6.0)
    else:
        p, g = scenes.config4(scale=0.02)
    p.strip_rows = 16
    return p, g


def _worker(rank, world, port, kind, outdir):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import orc
    from aqsis_b200 import sharding
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p, g = _scene(kind)
        p.rank, p.world_size = rank, world
        mine = sharding.split_grids_for_rank(p, g, rank, world)
        ch, disp, _ = orc.render(p, mine, 2)
        rows = sharding.rows_for_rank(p, rank)
        keep = np.zeros(p.yres, bool)
        keep[rows] = True
        ch[~keep] = 0
        disp[0][~keep] = 0
        gather = sharding.ImageGather(p, rank, world, torch.device("cpu"), dist)
        res = gather([torch.from_numpy(ch.reshape(p.yres, -1)), torch.from_numpy(disp[0].reshape(p.yres, -1))])
        frac = torch.tensor([mine.n_grids / g.n_grids])
        fr = [torch.zeros(1) for _ in range(world)]
        dist.all_gather(fr, frac)
        if rank == 0:
            np.savez(os.path.join(outdir, "gathered.npz"), ch=res[0].numpy(), disp=res[1].numpy(),
                     frac=np.array([float(f) for f in fr]))
        else:
            assert res is None
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("kind,world", [("static", 2), ("mbdof", 2), ("deep", 3)])
def test_sharded_render_equals_single_process(tmp_path, kind, world):
    import torch.multiprocessing as mp
    import orc
    orc.lib()                      # build the oracle once, before the ranks race for it
    mp.spawn(_worker, args=(world, _free_port(), kind, str(tmp_path)), nprocs=world, join=True)
    got = np.load(os.path.join(tmp_path, "gathered.npz"))
    p, g = _scene(kind)
    ch, disp, _ = orc.render(p, g, 4)
    assert np.array_equal(got["ch"].view(np.uint32), ch.reshape(p.yres, -1).view(np.uint32))
    assert np.array_equal(got["disp"], disp[0].reshape(p.yres, -1))
    # grids are replicated only where they straddle: the shares add up to a bit more than one frame
    assert 1.0 <= got["frac"].sum() < 1.0 + 0.9 * world / 2, got["frac"]


def test_strips_partition_the_crop_window(native_lib):
    from aqsis_b200 import default_params, sharding
    for yres, crop, strip, world in [(200, None, 0, 1), (200, None, 32, 3), (1080, None, 64, 8), (97, (0, 50, 7, 91), 16, 4),
                                     (64, None, 40, 2), (1080, None, 0, 8), (2160, None, 0, 8),
                                     (1080, None, 0, 2), (100, (0, 50, 3, 90), 0, 4), (5, None, 0, 8)]:
        kw = dict(resolution=(50, yres))
        if crop:
            kw["crop"] = crop
        p = default_params(**kw)
        p.world_size, p.strip_rows = world, strip
        owner = -np.ones(yres, int)
        for r in range(world):
            for y0, y1 in sharding.strips_for_rank(p, r):
                assert np.all(owner[y0:y1] == -1)
                owner[y0:y1] = r
                if strip > 0:
                    assert (y0 - p.crop_ymin) % 16 == 0
        assert np.all(owner[p.crop_ymin:p.crop_ymax] >= 0)
        assert np.all(owner[:p.crop_ymin] == -1) and np.all(owner[p.crop_ymax:] == -1)
        if world > 1 and (p.crop_ymax - p.crop_ymin) >= 16 * world * max(1, strip // 16):
            assert len(set(owner[p.crop_ymin:p.crop_ymax])) == world
        if strip == 0 and world > 1 and (p.crop_ymax - p.crop_ymin) >= world:
            # balanced mode: every rank owns the same number of strips and row counts differ by < one strip's rounding
            counts = [int((owner == r).sum()) for r in range(world)]
            nstr = [len(sharding.strips_for_rank(p, r)) for r in range(world)]
            assert len(set(nstr)) == 1 and max(counts) - min(counts) <= nstr[0]


def test_split_keeps_grid_payload_intact(native_lib):
    from aqsis_b200 import scenes, sharding
    p, g = scenes.config3(scale=0.05)
    p.world_size, p.strip_rows = 2, 16
    parts = [sharding.split_grids_for_rank(p, g, r, 2) for r in range(2)]
    for part in parts:
        assert part.P.shape[0] == int(((part.cu + 1) * (part.cv + 1) * part.nkeys).sum())
        assert part.Ci.shape[0] == part.n_verts and len(part.key_times) == int(part.nkeys.sum())
    # a grid far from every strip of a rank must not be sent to it: recompute the library's rule (aqh_grid_rank_masks) in numpy
    nv, nk = (g.cu.astype(np.int64) + 1) * (g.cv + 1), g.nkeys.astype(np.int64)
    pstart = np.concatenate([[0], np.cumsum(nv * nk)])[:-1]
    y, z = np.asarray(g.P)[:, 1], np.asarray(g.P)[:, 2].astype(np.float64)
    coc = lambda zz: p.dof_multiplier * np.abs(1.0 / zz - p.dof_one_over_focal_distance) * p.dof_scale_y
    pad = np.floor(p.filter_ywidth / 2.0) + 1.0 + np.maximum(coc(np.minimum.reduceat(z, pstart)), coc(np.maximum.reduceat(z, pstart))) * 1.001 + 1e-3
    lo = np.floor(np.minimum.reduceat(y, pstart).astype(np.float64) - pad)
    hi = np.ceil(np.maximum.reduceat(y, pstart).astype(np.float64) + pad)
    masks = sharding.grid_rank_masks(p, g)
    for r, part in enumerate(parts):
        touched = np.zeros(g.n_grids, bool)
        for y0, y1 in sharding.strips_for_rank(p, r):
            touched |= (hi >= y0) & (lo < y1)
        assert part.n_grids == int(touched.sum())
        assert np.array_equal(((masks >> np.uint64(r)) & np.uint64(1)).astype(bool), touched)


def test_camera_space_grids_are_sharded_by_their_projection(native_lib):
    """Ownership of AQH_GRID_CAMERA_SPACE grids is decided from the rows they project to (the same formula as k_project),
    not from their camera-space y: the raster-space and the camera-space version of a frame shard identically."""
    from aqsis_b200 import scenes, sharding
    p, g = scenes.config3(scale=0.06)
    p.world_size, p.strip_rows = 3, 16
    want = sharding.grid_rank_masks(p, g)
    pc, gc = scenes.to_camera_space(*scenes.config3(scale=0.06))
    pc.world_size, pc.strip_rows = 3, 16
    got = sharding.grid_rank_masks(pc, gc)
    # the round trip raster -> camera -> raster moves a vertex by rounding noise only: at most a handful of borderline grids differ
    assert (want != got).mean() < 0.01
    assert np.all((got & ~want) == 0) or (want != got).sum() <= 5


def test_contiguous_and_balanced_strips(native_lib):
    """strip_rows -1: one contiguous strip per rank; -2: boundaries from aqh_balance_strips (equal estimated work)."""
    from aqsis_b200 import scenes, sharding
    p, g = scenes.config2(scale=0.25)
    for world in (2, 4, 8):
        p.world_size, p.strip_rows = world, -1
        strips = [sharding.strips_for_rank(p, r) for r in range(world)]
        assert all(len(s) == 1 for s in strips)
        assert strips[0][0][0] == 0 and strips[-1][0][1] == p.yres
        assert all(strips[r][0][1] == strips[r + 1][0][0] for r in range(world - 1))
        sharding.balance_strips(p, [g])
        assert p.strip_rows == -2
        strips = [sharding.strips_for_rank(p, r) for r in range(world)]
        assert strips[0][0][0] == 0 and strips[-1][0][1] == p.yres
        assert all(strips[r][0][1] == strips[r + 1][0][0] for r in range(world - 1))
        # the estimated work per rank is even for a statistically uniform scene
        masks = sharding.grid_rank_masks(p, g)
        share = np.array([int(((masks >> np.uint64(r)) & np.uint64(1)).sum()) for r in range(world)], float)
        assert share.max() / share.min() < 1.35, share
    # a scene whose grids sit in the top third: the balanced cuts crowd there
    rows_used = np.asarray(g.P)[:, 1].reshape(g.n_grids, -1).mean(1) < p.yres / 3
    from aqsis_b200 import GridArrays
    idx = np.nonzero(rows_used)[0]
    nv = 17 * 17
    sel = (idx[:, None] * nv + np.arange(nv)[None, :]).ravel()
    top = GridArrays(cu=g.cu[idx], cv=g.cv[idx], flags=g.flags[idx], P=np.asarray(g.P)[sel], Ci=np.asarray(g.Ci)[sel], Oi=np.asarray(g.Oi)[sel])
    p.world_size = 4
    sharding.balance_strips(p, [top])
    assert p.strip_bounds[3] <= p.yres // 3 + 32


def test_config4_streaming_shard_equals_post_hoc_shard():
    """bench.py shards the 5.4 GB config-4 scene layer by layer while generating it (N ranks never hold N full
    copies): same grids, in the same order, as sharding the finished scene."""
    from aqsis_b200 import scenes, sharding
    for rank in (0, 1, 2):
        def shard(p, block, rank=rank):
            p.rank, p.world_size = rank, 3
            return sharding.split_grids_for_rank(p, block, rank, 3)
        p_s, g_s = scenes.config4(scale=0.04, shard=shard)
        p_f, g_f = scenes.config4(scale=0.04)
        p_f.rank, p_f.world_size = rank, 3
        m = sharding.split_grids_for_rank(p_f, g_f, rank, 3)
        assert g_s.total_micropolygons == g_f.n_micropolygons
        assert np.array_equal(np.asarray(g_s.P), np.asarray(m.P)) and np.array_equal(np.asarray(g_s.Oi), np.asarray(m.Oi))
        assert np.array_equal(g_s.cu, m.cu) and g_s.n_micropolygons == m.n_micropolygons
