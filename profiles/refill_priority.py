"""Does the refill belong on a lower-priority stream?  8 slabs; every slab steps on its own stream and refills its
finished envs (list scan + refill kernel, the unfused calls) (A) on the same stream, (B) on a second stream of equal
priority, (C) step stream high priority / refill stream low priority.  us per step over 400 steps."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

dev = torch.device("cuda", 0)
n, slabs = 65536, 8
envs = [bench.make_env(dev, s * n, n_envs=n) for s in range(slabs)]
acts = bench.synth_actions(8, n, 1, device=dev)
lo_p, hi_p = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
for i in range(2048 * slabs):
    envs[i % slabs].step(acts[i % 8], auto_reset=True)
torch.cuda.synchronize()


def run(mode, K=400):
    if mode == "A":
        st = [torch.cuda.Stream(dev) for _ in range(slabs)]; rs = st
    elif mode == "B":
        st = [torch.cuda.Stream(dev) for _ in range(slabs)]; rs = [torch.cuda.Stream(dev) for _ in range(slabs)]
    else:
        st = [torch.cuda.Stream(dev, priority=-1) for _ in range(slabs)]; rs = [torch.cuda.Stream(dev, priority=0) for _ in range(slabs)]
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream()
    for warm in (True, False):
        if not warm:
            torch.cuda.synchronize()
            a.record(cur)
            for s in st + (rs if rs is not st else []):
                s.wait_event(a)
        for i in range(8 * slabs if warm else K):
            s_ = i % slabs
            env = envs[s_]
            env.bind_stream(st[s_])
            env.step(acts[i % 8])
            if rs is not st:
                e1 = torch.cuda.Event(); e1.record(st[s_]); rs[s_].wait_event(e1)
                env.bind_stream(rs[s_])
            env._reset_launch(env.t["done"], bench.WINDOW, True, None, None)
            if rs is not st:
                e2 = torch.cuda.Event(); e2.record(rs[s_]); st[s_].wait_event(e2)
    for s in st + (rs if rs is not st else []):
        ev = torch.cuda.Event(); ev.record(s); cur.wait_event(ev)
    b.record(cur)
    torch.cuda.synchronize()
    for e in envs:
        e.bind_stream(None)
    return a.elapsed_time(b) / K * 1e3


for rep in range(2):
    print("  ".join(f"{m}: {run(m):.1f} us/step" for m in ("A", "B", "C")))
