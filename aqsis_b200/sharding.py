"""Multi-GPU sharding of one frame: image strips per rank, grid replication, final gather.

SURVEY.md 8(e): image buckets are independent once the `shift`-pixel halo of samples is recomputed
locally, so strips of pixel rows are dealt round-robin to the ranks, every grid is sent to each
rank whose strips its (motion / DoF / filter-expanded) bound touches -- straddling grids are
REPLICATED -- and the only collective is the gather of the finished strips to rank 0.
Works with any torch.distributed backend: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
import ctypes as C

import numpy as np

from .hider import GridArrays, lib


def strips_for_rank(params, rank):
    """[(y0, y1), ...] pixel-row strips owned by `rank` (aqh_strip_layout: the library's own dealing)."""
    L = lib()
    n = C.c_int()
    cap = params.yres // 16 + 2
    y0 = np.zeros(cap, np.int32)
    y1 = np.zeros(cap, np.int32)
    rc = L.aqh_strip_layout(C.byref(params), int(rank), C.byref(n), y0.ctypes.data, y1.ctypes.data, cap)
    if rc:
        raise ValueError(f"aqh_strip_layout failed ({rc})")
    return [(int(a), int(b)) for a, b in zip(y0[:n.value], y1[:n.value])]


def rows_for_rank(params, rank):
    s = strips_for_rank(params, rank)
    return np.concatenate([np.arange(a, b) for a, b in s]).astype(np.int64) if s else np.zeros(0, np.int64)


def grid_row_ranges(params, grids: GridArrays):
    """Conservative [lo, hi] pixel-row range each grid can contribute to (imagebuffer.cpp:519-554:
    union of the key bounds, grown by the largest circle of confusion and the filter half-width)."""
    nv = (grids.cu.astype(np.int64) + 1) * (grids.cv.astype(np.int64) + 1)
    nk = grids.nkeys.astype(np.int64) if grids.nkeys is not None else np.ones_like(nv)
    pstart = np.concatenate([[0], np.cumsum(nv * nk)])
    P = np.asarray(grids.P)
    y, z = P[:, 1], P[:, 2]
    ymin = np.minimum.reduceat(y, pstart[:-1])
    ymax = np.maximum.reduceat(y, pstart[:-1])
    pad = np.floor(params.filter_ywidth / 2.0) + 1.0
    if params.use_dof:
        zmin = np.minimum.reduceat(z, pstart[:-1]).astype(np.float64)
        zmax = np.maximum.reduceat(z, pstart[:-1]).astype(np.float64)

        def coc(zz):
            return params.dof_multiplier * np.abs(1.0 / zz - params.dof_one_over_focal_distance) * params.dof_scale_y
        # |1/z - 1/fd| is convex in 1/z: its maximum over the bound is at an end point
        pad = pad + np.maximum(coc(zmin), coc(zmax)) * 1.001 + 1e-3
    lo = np.floor(ymin - pad).astype(np.int64)
    hi = np.ceil(ymax + pad).astype(np.int64)
    return lo, hi


def split_grids_for_rank(params, grids: GridArrays, rank, world):
    """The grids `rank` must receive: those whose row range touches one of its strips."""
    if world == 1:
        return grids
    nv = (grids.cu.astype(np.int64) + 1) * (grids.cv.astype(np.int64) + 1)
    nk = grids.nkeys.astype(np.int64) if grids.nkeys is not None else np.ones_like(nv)
    pstart = np.concatenate([[0], np.cumsum(nv * nk)])
    vstart = np.concatenate([[0], np.cumsum(nv)])
    lo, hi = grid_row_ranges(params, grids)
    keep = np.zeros(grids.n_grids, dtype=bool)
    for y0, y1 in strips_for_rank(params, rank):
        keep |= (hi >= y0) & (lo < y1)
    idx = np.nonzero(keep)[0]

    def take(starts):
        if not len(idx):
            return np.zeros(0, np.int64)
        lens = starts[idx + 1] - starts[idx]
        base = np.repeat(starts[idx] - np.concatenate([[0], np.cumsum(lens)[:-1]]), lens)
        return base + np.arange(lens.sum())

    pos_idx, vert_idx = take(pstart), take(vstart)
    kt = None
    if grids.key_times is not None:
        kstart = np.concatenate([[0], np.cumsum(nk)])
        kt = np.asarray(grids.key_times)[take(kstart)]
    return GridArrays(cu=grids.cu[idx], cv=grids.cv[idx], flags=grids.flags[idx], P=np.asarray(grids.P)[pos_idx],
                      Ci=None if grids.Ci is None else np.asarray(grids.Ci)[vert_idx],
                      Oi=None if grids.Oi is None else np.asarray(grids.Oi)[vert_idx],
                      nkeys=None if grids.nkeys is None else grids.nkeys[idx], key_times=kt,
                      lod_bounds=None if grids.lod_bounds is None else np.asarray(grids.lod_bounds).reshape(-1, 2)[idx].ravel(),
                      culled=None if grids.culled is None else np.asarray(grids.culled)[vert_idx])


class ImageGather:
    """Gather of the finished strips to rank 0 -- the only collective of the path.

    Every rank holds full-size images in which only its own rows are valid.  Rows are packed,
    padded to the largest per-rank row count and gathered with dist.gather; rank 0 scatters them
    into place.  Index tensors are built once per frame layout, not per step."""

    def __init__(self, params, rank, world, device, dist=None):
        import torch
        self.rank, self.world, self.dist = rank, world, dist
        if world == 1:
            return
        self.rows = [torch.from_numpy(rows_for_rank(params, r)).to(device) for r in range(world)]
        self.max_rows = max(int(r.numel()) for r in self.rows)
        pad = torch.zeros(self.max_rows, dtype=torch.long, device=device)
        pad[:self.rows[rank].numel()] = self.rows[rank]
        self.send_idx = pad

    def __call__(self, images):
        """images: list of (yres, rowlen) tensors.  Returns the assembled list on rank 0, None elsewhere."""
        if self.world == 1:
            return images
        out = []
        for img in images:
            send = img.index_select(0, self.send_idx)
            if self.rank == 0:
                recv = [send.new_empty(send.shape) for _ in range(self.world)]
                self.dist.gather(send, recv, dst=0)
                full = img.new_zeros(img.shape)
                for r in range(self.world):
                    k = int(self.rows[r].numel())
                    full.index_copy_(0, self.rows[r], recv[r][:k])
                out.append(full)
            else:
                self.dist.gather(send, None, dst=0)
        return out if self.rank == 0 else None
