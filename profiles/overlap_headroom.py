"""How much of the bench's whole-step time is the list-driven refill?  8 slabs on 8 streams, 2,048 setup steps, then
400 steps (a) with auto-reset (step + refill, the bench's timed region), (b) the step kernel only (finished envs are
left finished), (c) step only on ONE stream."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

dev = torch.device("cuda", 0)
n, slabs = 65536, 8
envs = [bench.make_env(dev, s * n, n_envs=n) for s in range(slabs)]
acts = bench.synth_actions(8, n, 1, device=dev)
streams = [torch.cuda.Stream(dev) for _ in range(slabs)]
for e, s in zip(envs, streams):
    e.bind_stream(s)
for i in range(int(os.environ.get("SETUP", "2048")) * slabs):
    envs[i % slabs].step(acts[i % 8], auto_reset=True)
torch.cuda.synchronize()


def run(auto, K=400):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream()
    a.record(cur)
    for s in streams:
        s.wait_event(a)
    for i in range(K):
        envs[i % slabs].step(acts[i % 8], auto_reset=auto)
    for s in streams:
        ev = torch.cuda.Event(); ev.record(s); cur.wait_event(ev)
    b.record(cur)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / K * 1e3


for rep in range(2):
    t_auto = run(True)
    t_step = run(False)
    # restore the population for the next repeat
    for i in range(8 * slabs):
        envs[i % slabs].step(acts[i % 8], auto_reset=True)
    print(f"8 streams: step + refill {t_auto:.1f} us/step, step kernel only {t_step:.1f} us/step")
for e in envs:
    e.bind_stream(None)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(400):
    envs[i % slabs].step(acts[i % 8])
b.record(); torch.cuda.synchronize()
print(f"1 stream: step kernel only {a.elapsed_time(b) / 400 * 1e3:.1f} us/step")
