import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import bench
from gym_novel_gridworlds_b200.runtime import BatchHandle
desc, compiled, envs, rule, kw = bench.build_workload('C2')
h = BatchHandle(compiled, envs); h.reset()
d = torch.zeros(envs*63, dtype=torch.int32, device='cuda'); p = torch.zeros(envs*63, dtype=torch.int32).pin_memory()
for n in (1, 2, 3):
    torch.cuda.synchronize(); ts=[]
    for i in range(50):
        t0=time.perf_counter(); p.copy_(d, non_blocking=True); torch.cuda.synchronize(); ts.append(time.perf_counter()-t0)
    ts=np.array(ts)*1e3; print('D2H 16.5MB ms: min %.3f med %.3f max %.3f -> %.1f GB/s' % (ts.min(), np.median(ts), ts.max(), 16.5/np.median(ts)))
a = np.random.randint(0, 10, size=envs).astype(np.int32)
for rep in range(3):
    ts=[]
    for i in range(100):
        t0=time.perf_counter(); h.step_host(a); ts.append(time.perf_counter()-t0)
    ts=np.array(ts)*1e3; print('step_host ms: min %.3f med %.3f p90 %.3f max %.3f -> %.1f M steps/s (med)' % (ts.min(), np.median(ts), np.percentile(ts,90), ts.max(), envs/np.median(ts)/1e3))
