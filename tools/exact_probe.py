import sys, os
sys.path[:0]=['/root/repo','/root/repo/tests']
import numpy as np, ctypes as C
from aqsis_b200 import scenes, Hider, lib, abi
import orc, parity_util as pu
h=Hider(0)
def dof(p):
    lib().aqh_frame_params_set_dof(C.byref(p), 2.8, 0.05, 20.0, 60.0, 60.0); return p
cases={"config1":scenes.config1(scale=0.3),"config2":scenes.config2(scale=0.1),"config3":scenes.config3(scale=0.06, motion_px=6.0),
 "config4":scenes.config4(scale=0.02),"sinc5":scenes.config2(scale=0.05, filter=("sinc",5.0,5.0), samples=(4,4))}
p,g=scenes.config3(scale=0.06, motion_px=8.0); p.use_dof=0; cases["motion"]=(p,g)
p,g=scenes.config2(scale=0.06); cases["dof"]=(dof(p),g)
for name,(p,g) in cases.items():
    p.filter_mode=0
    ch_g,d_g,_=pu.run_product(h,p,g)
    ch_o,d_o,_=orc.render(p,g,8)
    ne=(ch_g.view(np.uint32)!=ch_o.view(np.uint32))
    print(name, "exact frac", 1-ne.mean(), "pixels differing", ne.any(axis=2).sum(), "of", ne.shape[0]*ne.shape[1], "max abs", np.nanmax(np.abs(np.where(np.abs(ch_o)<1e30, ch_g-ch_o, 0))), "quant diff", np.abs(d_g[0].astype(int)-d_o[0].astype(int)).max())
    if ne.any():
        ys,xs,cs=np.nonzero(ne); 
        print("   channels", np.bincount(cs, minlength=9), "first", ys[0],xs[0], ch_g[ys[0],xs[0]], ch_o[ys[0],xs[0]])
