"""config 5 of BASELINE.json: PixelFilter sweep (box, triangle, gaussian, catmull-rom, sinc at widths 1-6) on the
1080p / 8x8 spp scene of config 2.  Prints one line per filter: frame, hide and filter stage times (resident grids),
filter-only Mpixels/s and Gtaps/s (taps = (2*shift+1)^2 * n samples per output pixel)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aqsis_b200 import Hider, scenes

dev = torch.device("cuda", 0)
h = Hider(0, stream=torch.cuda.current_stream().cuda_stream)
rows = []
grids_dev = None
for name, w in scenes.config5_filters():
    p, g = scenes.config2(filter=(name, w, w))
    if grids_dev is None:
        grids_dev = g.to_torch(device=dev)
    h.begin_frame(p)
    h.add_grid_block(grids_dev)
    for _ in range(2):
        h.render_device()
    acc = {"project_bust_ms": 0.0, "render_mpgs_ms": 0.0, "filter_ms": 0.0}
    reps = 3
    for _ in range(reps):
        h.render_device()
        s = h.stats()
        for k in acc:
            acc[k] += s[k] / reps
    shift = int(w // 2)
    taps = (2 * shift + 1) ** 2 * p.xsamples * p.ysamples
    px = p.xres * p.yres
    row = {"filter": name, "width": w, "shift": shift, "taps_per_pixel": taps, **{k: round(v, 3) for k, v in acc.items()},
           "frame_ms": round(sum(acc.values()), 3), "filter_mpixels_per_s": round(px / acc["filter_ms"] / 1e3, 1),
           "filter_gtaps_per_s": round(px * taps * 8 / acc["filter_ms"] / 1e6, 1)}
    rows.append(row)
    print(f"{name:12s} w={w:.0f} taps/px={taps:5d}  project+bin {acc['project_bust_ms']:6.2f}  hide {acc['render_mpgs_ms']:6.2f}  "
          f"filter {acc['filter_ms']:6.2f} ms  {row['filter_mpixels_per_s']:8.1f} Mpx/s  {row['filter_gtaps_per_s']:7.1f} G tap-channels/s", flush=True)
json.dump(rows, open(os.path.join(os.environ.get("GRAFT_REPO_ROOT", "."), "gpurun_out", "filter_sweep.json"), "w"), indent=1)
