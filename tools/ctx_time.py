import time, sys, os
sys.path.insert(0, '/root/repo')
t0=time.perf_counter()
from approx_counter_b200 import ApproxCounter, load
load(); t1=time.perf_counter()
c=ApproxCounter(0); t2=time.perf_counter()
d=ApproxCounter(0); t3=time.perf_counter()
print(f"dlopen {t1-t0:.3f}s first ctx {t2-t1:.3f}s second ctx {t3-t2:.3f}s  CUDA_MODULE_LOADING={os.environ.get('CUDA_MODULE_LOADING')}")
