import sys, time; sys.path.insert(0,'/root/repo')
import torch, bench
dev=torch.device('cuda',0)
env=bench.make_env(dev,0)
acts=bench.synth_actions(4, bench.ENVS_PER_GPU, 1, device=dev)
for i in range(50): env.step(acts[i%4], auto_reset=True)
torch.cuda.synchronize()
def t(f,n=150):
    torch.cuda.synchronize(); t0=time.perf_counter()
    for i in range(n): f(i)
    dt=(time.perf_counter()-t0)/n*1e6; torch.cuda.synchronize(); return dt
import ctypes as C
print("step only   host us:", t(lambda i: env.step(acts[i%4])))
print("reset only  host us:", t(lambda i: env._reset_launch(env.t["done"],64,True,None,None)))
print("launch()    host us:", t(lambda i: env._launch(1,0)))
L=env._launch(1,0); lib=env._lib
print("raw mdg_step ctypes us:", t(lambda i: lib.mdg_step(C.byref(env.P),C.byref(env.R),C.byref(env._S),C.byref(env._IO),C.byref(L))))
print("current_stream us:", t(lambda i: torch.cuda.current_stream(dev).cuda_stream))
