"""Timeline of a short burst of steps (what `bench.py --steps 20` times): CUDA events on every stream after each
step kernel and each refill kernel, printed relative to the start of the burst.
usage: python profiles/burst_timeline.py [K=20] [streams=8]"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from madigan_b200.environments import env as E

K = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n_str = int(sys.argv[2]) if len(sys.argv) > 2 else 8
slabs = 8
setup = int(sys.argv[3]) if len(sys.argv) > 3 else 64
dev = torch.device("cuda", 0)
envs = [bench.make_env(dev, s * 65536) for s in range(slabs)]
acts = bench.synth_actions(8, 65536, 1234, device=dev)
streams = [torch.cuda.Stream(dev) for _ in range(n_str)]
for i, e in enumerate(envs):
    e.bind_stream(streams[i % n_str])
for i in range(setup * slabs + 3):
    envs[i % slabs].step(acts[i % 8], auto_reset=True)
torch.cuda.synchronize()
main = torch.cuda.current_stream(dev)
for rep in range(3):
    ev0 = torch.cuda.Event(enable_timing=True)
    ev0.record(main)
    for st in streams:
        st.wait_event(ev0)
    marks = []
    for i in range(K):
        env = envs[i % slabs]
        st = streams[(i % slabs) % n_str]
        # split call: step kernel (appending to the list), then the refill
        env.step(acts[i % 8], auto_reset=True)
        e2 = torch.cuda.Event(enable_timing=True); e2.record(st)
        marks.append((i, e2))
    torch.cuda.synchronize()
    ts = [(i, ev0.elapsed_time(e) * 1e3) for i, e in marks]
    print(f"rep {rep}: end of (step+refill) i on its stream, us after the start:")
    print("  " + "  ".join(f"{i}:{t:.0f}" for i, t in ts))
    print(f"  total {max(t for _, t in ts):.0f} us for {K} steps = {max(t for _, t in ts)/K:.1f} us/step; done rates "
          + " ".join(f"{float(e.t['done'].float().mean()):.4f}" for e in envs))
