"""Steady-state device-time breakdown of one bench step on ONE stream (CUDA events, no profiler):
step kernel alone, then + reset scan, + rng, + recurrence.  The reset phases are cut with the profiling knob
MDG_RESET_PHASES (read once per process), so each configuration runs in its own process:
    for p in 0 1 2 3; do python profiles/breakdown.py $p; done"""
import os, sys
phase = int(sys.argv[1]) if len(sys.argv) > 1 else 3
os.environ["MDG_RESET_PHASES"] = str(max(phase, 1))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
dev = torch.device("cuda", 0)
slabs = [bench.make_env(dev, i) for i in range(8)]
acts = bench.synth_actions(4, bench.ENVS_PER_GPU, 1, device=dev)
def one(i):
    env = slabs[i % 8]
    env.step(acts[i % 4], auto_reset=(phase > 0))
for i in range(160): one(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 400
e0.record()
for i in range(n): one(i)
e1.record(); torch.cuda.synchronize()
done_frac = sum(float(s.t["done"].float().mean()) for s in slabs) / 8
print(f"phases={phase}: {e0.elapsed_time(e1)/n*1e3:7.2f} us per step (single stream, 8 rotating slabs), done fraction {done_frac:.4f}")
