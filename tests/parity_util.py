"""Shared helpers of the parity tests: run the product (through the C ABI) and the oracle on the
same inputs and compare with the tolerances BASELINE.json's north_star states."""
import numpy as np

import orc

FLOAT_RTOL = 1e-4      # "relative error <= 1e-4 on float output"
QUANT_ATOL = 1         # "per-channel max absolute error <= 1/255 on 8-bit output"


def run_product(h, params, grids, use_block=True):
    h.begin_frame(params)
    if use_block:
        h.add_grid_block(grids)
    else:
        add_grids_one_by_one(h, grids)
    ch, disp = h.end_frame()
    return ch, disp, h.stats()


def add_grids_one_by_one(h, grids):
    import numpy as np
    po = vo = ko = 0
    for g in range(grids.n_grids):
        cu, cv = int(grids.cu[g]), int(grids.cv[g])
        nv = (cu + 1) * (cv + 1)
        nk = int(grids.nkeys[g]) if grids.nkeys is not None else 1
        P = np.asarray(grids.P[po:po + nv * nk]).reshape(nk, nv, 3)
        kt = grids.key_times[ko:ko + nk] if (grids.key_times is not None and nk > 1) else None
        h.add_grid(P, cu, cv,
                   Ci=None if grids.Ci is None else grids.Ci[vo:vo + nv],
                   Oi=None if grids.Oi is None else grids.Oi[vo:vo + nv],
                   flags=int(grids.flags[g]), key_times=kt,
                   culled=None if grids.culled is None else grids.culled[vo:vo + nv],
                   lod_bounds=None if grids.lod_bounds is None else grids.lod_bounds[2 * g:2 * g + 2])
        po += nv * nk
        vo += nv
        ko += nk


def compare(ch_gpu, disp_gpu, ch_ref, disp_ref, float_rtol=FLOAT_RTOL, quant_atol=QUANT_ATOL, strict_special=True):
    """Returns a dict of error measures; raises AssertionError beyond tolerance.

    strict_special=False (tile-partials filter mode): where the reference's filtered depth is not
    a finite ordinary number (sums involving FLT_MAX depths overflow, and how depends on the order
    of the additions) the z channel is not compared."""
    a = ch_gpu.astype(np.float64)
    b = ch_ref.astype(np.float64)
    if not strict_special:
        bad = ~(np.isfinite(b[..., 7]) & (np.abs(b[..., 7]) < 1e30))
        a = a.copy()
        a[..., 7] = np.where(bad, b[..., 7], a[..., 7])
    finite = np.isfinite(a) & np.isfinite(b) & (np.abs(b) < 1e30)
    # relative error against max(|ref|, small floor) so that exact zeros compare absolutely
    denom = np.maximum(np.abs(b), 1e-3)
    with np.errstate(invalid="ignore", over="ignore"):          # inf - inf / FLT_MAX - x where the reference holds specials
        rel = np.where(finite, np.abs(a - b) / denom, 0.0)
    # where the reference holds FLT_MAX / inf both must agree exactly
    special_equal = np.array_equal(np.where(finite, 0, a), np.where(finite, 0, b), equal_nan=True)
    out = {
        "float_max_rel": float(rel.max()) if rel.size else 0.0,
        "float_bit_exact_frac": float((ch_gpu.view(np.uint32) == ch_ref.view(np.uint32)).mean()),
        "special_equal": bool(special_equal),
        "quant_max_abs": 0,
        "quant_exact_frac": 1.0,
    }
    for dg, dr in zip(disp_gpu, disp_ref):
        assert dg.shape == dr.shape and dg.dtype == dr.dtype
        if dg.dtype.kind == "f":
            continue
        d = np.abs(dg.astype(np.int64) - dr.astype(np.int64))
        out["quant_max_abs"] = max(out["quant_max_abs"], int(d.max()) if d.size else 0)
        out["quant_exact_frac"] = min(out["quant_exact_frac"], float((d == 0).mean()) if d.size else 1.0)
    assert out["special_equal"], f"FLT_MAX/inf pattern differs: {out}"
    assert out["float_max_rel"] <= float_rtol, f"float channels differ: {out}"
    assert out["quant_max_abs"] <= quant_atol, f"quantised output differs: {out}"
    return out


def parity(h, params, grids, nthreads=8, use_block=True, **tol):
    ch_g, disp_g, st = run_product(h, params, grids, use_block=use_block)
    ch_r, disp_r, ost = orc.render(params, grids, nthreads)
    res = compare(ch_g, disp_g, ch_r, disp_r, **tol)
    res["gpu_stats"] = st
    res["oracle_stats"] = ost
    return res
