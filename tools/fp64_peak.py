import sys, ctypes as C
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import helpers as H
from pd_mg_pin_corrosion_b200 import lib as L_, solver as S
dim,cfg,_=H.load_cfg('3d_small')
g=S.Grid(3); g.build(cfg)
L=L_.load(); a=C.c_double(); b=C.c_double()
L_.check(L.pdgpu_fp64_peak(g.ctx,C.byref(a))); L_.check(L.pdgpu_fp64_peak3(g.ctx,C.byref(b)))
print('fp64 peak (1 reg src)',a.value,'TF; 3 reg src',b.value,'TF')
