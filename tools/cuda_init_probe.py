import os, sys, time, subprocess
code = r'''
import time, ctypes, os, sys
t0=time.perf_counter()
sys.path.insert(0, ".")
import lamsa_b200
lib = lamsa_b200.load_library()
t1=time.perf_counter()
h=ctypes.c_void_p()
rc=lib.lb2_ctx_create(0, ctypes.byref(h))
t2=time.perf_counter()
h2=ctypes.c_void_p()
lib.lb2_ctx_create(0, ctypes.byref(h2))
t3=time.perf_counter()
print("load %.3f first ctx %.3f second ctx %.3f rc %d" % (t1-t0, t2-t1, t3-t2, rc))
'''
for env in ({}, {"CUDA_VISIBLE_DEVICES": "0"}, {"CUDA_VISIBLE_DEVICES": "0", "CUDA_MODULE_LOADING": "LAZY"}, {"CUDA_VISIBLE_DEVICES": "0", "CUDA_DEVICE_MAX_CONNECTIONS": "4"}):
    for rep in range(2):
        e = dict(os.environ); e.update(env)
        t=time.perf_counter()
        out = subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True).stdout.strip()
        print(env, out, "process %.3f" % (time.perf_counter()-t), flush=True)
subprocess.run("nvidia-smi -q | grep -i -E 'persistence|Product Name' | head -4; nvidia-smi -L | wc -l", shell=True)
