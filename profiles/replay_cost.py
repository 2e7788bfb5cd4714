"""Device time of the replay ingest (window materialisation into the observation slots + transition append) and of
sampling, at the bench shape (65,536 envs x 16 assets, window 64, fp32 normalised windows)."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from madigan_b200.utils.replay import DeviceReplay
dev = torch.device("cuda", 0)
env = bench.make_env(dev, 0)
acts = bench.synth_actions(4, bench.ENVS_PER_GPU, 1, device=dev)
rp = DeviceReplay(env, depth=16, norm_type="standard_normal", dtype=torch.float32)
rp.observe_start()
def timed(f, n):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for i in range(n): f(i)
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n * 1e3
def step_only(i): env.step(acts[i % 4], auto_reset=True)
def step_add(i):
    env.step(acts[i % 4], auto_reset=True); rp.add(acts[i % 4])
for i in range(30): step_add(i)
t0 = timed(step_only, 100); t1 = timed(step_add, 100)
print(f"step+reset {t0:.1f} us, step+reset+ingest {t1:.1f} us -> ingest {t1 - t0:.1f} us per 65,536-env step "
      f"({65536 * 64 * 16 * 4 / 1e6:.0f} MB of fp32 windows + records), {len(rp)} transitions stored")
import time
for B in (256, 4096, 32768):
    rp.sample(B); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(50): rp.sample(B)
    torch.cuda.synchronize()
    print(f"sample({B}): {(time.perf_counter() - t0) / 50 * 1e6:.1f} us per call, host + device "
          f"({B * 2 * 64 * 16 * 4 / 1e6:.1f} MB gathered)")
