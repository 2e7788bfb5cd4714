"""C4 (65,536 x composite16, margin .1, PPC n=5 per-asset rewards): step kernel / refill / window timings (events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from madigan_b200.environments import Env
dev = torch.device("cuda", 0)
wl = bench.WORKLOADS["c4"]
n, nA = 65536, 16
env = Env(wl["ds"][0], 1e6, {"data_source_config": wl["ds"][1]}, n_envs=n, window=64, seed=5, device=dev, reward=wl["reward"])
env.setRequiredMargin(.1); env.setMaintenanceMargin(.25); env.setTransactionCost(.001, 0.)
env.reset(fill_history=True)
g = torch.Generator().manual_seed(1)
acts = [(torch.randint(-1, 2, (n, nA), generator=g).double() * 5000.).to(dev) for _ in range(4)]
for i in range(100): env.step(acts[i % 4], auto_reset=True)
torch.cuda.synchronize()
K = 30
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
for i in range(K):
    ev[i][0].record(); env.step(acts[i % 4]); ev[i][1].record()
    env._reset_launch(env.t["done"], 64, True, None, None); ev[i][2].record()
torch.cuda.synchronize()
st = sorted(a.elapsed_time(b) for a, b, _ in ev); rs = sorted(b.elapsed_time(c) for _, b, c in ev)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(K): env.step(acts[i % 4], auto_reset=True)
b.record(); torch.cuda.synchronize()
print(f"C4 fused step+refill {a.elapsed_time(b) / K * 1e3:.1f} us per step")
print(f"C4 step kernel median {st[K//2]*1e3:.1f} us, masked reset {rs[K//2]*1e3:.1f} us, done rate {float(env.t['done'].float().mean()):.4f}")
