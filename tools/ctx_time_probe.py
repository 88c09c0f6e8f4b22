import sys, time, os
sys.path.insert(0, '/root/repo')
import lamsa_b200
t0=time.perf_counter(); c0=lamsa_b200.Context(0); t1=time.perf_counter()
ts=[]
cs=[]
for i in range(12):
    a=time.perf_counter(); cs.append(lamsa_b200.Context(0)); ts.append(time.perf_counter()-a)
print("MAXCONN", os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"), "first ctx %.3f s; next 12: min %.4f max %.4f mean %.4f" % (t1-t0, min(ts), max(ts), sum(ts)/len(ts)))
import torch
a=time.perf_counter(); ss=[torch.cuda.Stream() for _ in range(24)]; torch.cuda.synchronize(); print("24 torch streams %.4f s" % (time.perf_counter()-a))
