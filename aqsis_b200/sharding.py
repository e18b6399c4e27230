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


def grid_rank_masks(params, grids: GridArrays):
    """uint64 per grid: bit r set when rank r must receive the grid (aqh_grid_rank_masks: the library's own rule --
    row range of the grid over all its keys, camera-space grids projected, grown by the largest circle of confusion and
    the filter half-width; straddlers carry several bits)."""
    b = grids.as_struct()
    masks = np.zeros(grids.n_grids, np.uint64)
    rc = lib().aqh_grid_rank_masks(C.byref(params), C.byref(b), masks.ctypes.data)
    if rc:
        raise ValueError(f"aqh_grid_rank_masks failed ({rc})")
    return masks


def balance_strips(params, blocks):
    """One contiguous strip per rank with equal estimated work (aqh_grid_row_cost + aqh_balance_strips)."""
    cost = np.zeros(params.yres, np.float64)
    for g in blocks:
        b = g.as_struct()
        rc = lib().aqh_grid_row_cost(C.byref(params), C.byref(b), cost.ctypes.data)
        if rc:
            raise ValueError(f"aqh_grid_row_cost failed ({rc})")
    rc = lib().aqh_balance_strips(C.byref(params), cost.ctypes.data)
    if rc:
        raise ValueError(f"aqh_balance_strips failed ({rc})")
    return params


def split_grids_for_rank(params, grids: GridArrays, rank, world):
    """The grids `rank` must receive: those whose row range touches one of its strips."""
    if world == 1:
        return grids
    nv = (grids.cu.astype(np.int64) + 1) * (grids.cv.astype(np.int64) + 1)
    nk = grids.nkeys.astype(np.int64) if grids.nkeys is not None else np.ones_like(nv)
    pstart = np.concatenate([[0], np.cumsum(nv * nk)])
    vstart = np.concatenate([[0], np.cumsum(nv)])
    keep = (grid_rank_masks(params, grids) >> np.uint64(rank)) & np.uint64(1)
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

    def verts(a):
        return None if a is None else np.asarray(a)[vert_idx]

    return GridArrays(cu=grids.cu[idx], cv=grids.cv[idx], flags=grids.flags[idx], P=np.asarray(grids.P)[pos_idx],
                      Ci=verts(grids.Ci), Oi=verts(grids.Oi),
                      nkeys=None if grids.nkeys is None else grids.nkeys[idx], key_times=kt,
                      lod_bounds=None if grids.lod_bounds is None else np.asarray(grids.lod_bounds).reshape(-1, 2)[idx].ravel(),
                      culled=verts(grids.culled), aov=verts(grids.aov), Ng=verts(grids.Ng), N=verts(grids.N),
                      radius=None if grids.radius is None else np.asarray(grids.radius)[pos_idx],
                      csg_node=None if grids.csg_node is None else np.asarray(grids.csg_node)[idx],
                      trim_set=None if grids.trim_set is None else np.asarray(grids.trim_set)[idx], trim_uv=verts(grids.trim_uv))


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
