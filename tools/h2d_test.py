import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from iv_slam_b200 import api
n=256; H,W=376,1241
pin=api.PinnedArray((n,H,W),np.uint8); pin.array[...]=7
ex=api.ORBextractor(2000,1.2,8,20,7); ex.reserve(W,H,n)
for _ in range(2): ex.upload(pin.array); ex.sync()
t=time.perf_counter()
for _ in range(5): ex.upload(pin.array); ex.sync()
dt=(time.perf_counter()-t)/5
print('pitched 3D upload: %.2f ms for %d images -> %.1f GB/s'%(dt*1e3,n,pin.array.nbytes/dt/1e9))
tp=torch.empty((n,H,W),dtype=torch.uint8).pin_memory(); td=torch.empty((n,H,W),dtype=torch.uint8,device='cuda')
for _ in range(2): td.copy_(tp,non_blocking=True); torch.cuda.synchronize()
t=time.perf_counter()
for _ in range(5): td.copy_(tp,non_blocking=True); torch.cuda.synchronize()
dt=(time.perf_counter()-t)/5
print('contiguous torch copy: %.2f ms -> %.1f GB/s'%(dt*1e3,tp.numel()/dt/1e9))
th=torch.empty((n,H,W),dtype=torch.uint8).pin_memory()
t=time.perf_counter()
for _ in range(5): th.copy_(td,non_blocking=True); torch.cuda.synchronize()
dt=(time.perf_counter()-t)/5
print('contiguous D2H: %.2f ms -> %.1f GB/s'%(dt*1e3,tp.numel()/dt/1e9))
