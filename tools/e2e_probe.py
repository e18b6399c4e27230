"""Where does an e2e step spend its wall time?  (host timers around the three public calls; AQH_TRACE=1 adds the
library's own checkpoints of the last frame on stderr)

    python tools/e2e_probe.py <config> [capture]      capture: deliver buckets through the capture display callbacks
"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from aqsis_b200 import Hider, scenes
from aqsis_b200.hider import display_info
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
capture = len(sys.argv) > 2
params, grids = {1: scenes.config1, 2: scenes.config2, 3: scenes.config3, 4: scenes.config4}[cfg]()
h = Hider(0, stream=torch.cuda.current_stream().cuda_stream)
pin = grids.to_torch(pin=True)
es = display_info(params, 0)[2]
cap_ch = np.zeros((params.yres, params.xres, 9), np.float32)
cap_d = [np.zeros((params.yres, params.xres, es), np.uint8)]
for it in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); h.begin_frame(params)
    t1 = time.perf_counter(); h.add_grid_block(pin)
    t2 = time.perf_counter()
    if capture:
        h.end_frame_capture(cap_ch, cap_d)
    else:
        h.end_frame()
    t3 = time.perf_counter()
    s = h.stats()
    print(f"begin {1e3*(t1-t0):.2f}  add_block {1e3*(t2-t1):.2f}  end_frame {1e3*(t3-t2):.2f}  "
          f"[upload {s['upload_ms']:.2f} device {s['device_total_ms']:.2f} download {s['download_ms']:.2f} bands {s.get('n_bands')}]", flush=True)
