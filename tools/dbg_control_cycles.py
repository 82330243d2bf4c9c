import sys, ctypes, tempfile, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cityflow_b200 import scenario
from cityflow_b200.capi import CEngine
d=tempfile.mkdtemp()
cfg=scenario.make_grid_scenario(d,30,30,dense=dict(frac=0.5,interval=10.0,seed=1),name='dbg')
e=CEngine(cfg)
e.lib.cfb_debug_arrays.restype=ctypes.c_int64
e.lib.cfb_debug_arrays.argtypes=[ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
P=int(e.lib.cfb_debug_arrays(e.h,None,None,0))
cyc=np.zeros(P,np.uint32); path=np.zeros(P,np.uint32)
e.next_step(1200)
e.lib.cfb_debug_arrays(e.h,cyc.ctypes.data,path.ctypes.data,P)
e.next_step(1)
e.lib.cfb_debug_arrays(e.h,cyc.ctypes.data,path.ctypes.data,P)
m=cyc>0
c=cyc[m]; p=path[m]
print('vehicles',m.sum(),'cycles: mean',c.mean(),'median',np.median(c),'p90',np.percentile(c,90),'p99',np.percentile(c,99),'max',c.max())
bits=p&0xff; cf=(p>>8)<<6
print('carfollow-phase cycles: mean',cf.mean(),'p99',np.percentile(cf,99),'max',cf.max())
for b in range(16):
    s=bits==b
    if s.sum(): print('path',format(b,'04b'),'n',s.sum(),'mean',c[s].mean(),'p99',np.percentile(c[s],99),'max',c[s].max())
