"""Where does an e2e step spend its wall time?  (host timers around the three public calls)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aqsis_b200 import Hider, scenes
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
params, grids = {2: scenes.config2, 3: scenes.config3, 4: scenes.config4}[cfg]()
h = Hider(0, stream=torch.cuda.current_stream().cuda_stream)
pin = grids.to_torch(pin=True)
for it in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); h.begin_frame(params)
    t1 = time.perf_counter(); h.add_grid_block(pin)
    t2 = time.perf_counter(); h.end_frame()
    t3 = time.perf_counter()
    s = h.stats()
    print(f"begin {1e3*(t1-t0):.2f}  add_block {1e3*(t2-t1):.2f}  end_frame {1e3*(t3-t2):.2f}  "
          f"[upload {s['upload_ms']:.2f} device {s['device_total_ms']:.2f} download {s['download_ms']:.2f}]")
print("--- raw ctypes calls")
import ctypes as C
for it in range(3):
    h.begin_frame(params)
    b = pin.as_struct()
    t1 = time.perf_counter(); rc = h._L.aqh_add_grid_block(h._h, C.byref(b))
    t2 = time.perf_counter(); rc = h._L.aqh_end_frame(h._h, None)
    t3 = time.perf_counter()
    print(f"raw add_block {1e3*(t2-t1):.2f}  raw end_frame {1e3*(t3-t2):.2f}")
