"""Device time of the step kernel and of step + auto-reset at the bench shape, L2-cold (8 rotating slabs of 65,536
envs), one stream, CUDA events around every launch.  Used to compare build variants inside ONE gpurun session:
    MDG_LIB_VARIANT=x MDG_EXTRA_NVCC_FLAGS="-DMDG_MINB_PAIRS=3" python -m madigan_b200.build   (here, no GPU needed)
    MDG_LIB_VARIANT=x python profiles/step_time.py                                             (on the GPU box)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

dev = torch.device("cuda", 0)
tag = os.environ.get("MDG_LIB_VARIANT", "") or "default"
sizes = [int(x) for x in (sys.argv[1:] or ["65536"])]
for n in sizes:
    slabs = max(1, min(8, (8 * 65_536) // n))
    envs = [bench.make_env(dev, s * n, n_envs=n) for s in range(slabs)]
    acts = bench.synth_actions(4, n, 1, device=dev)
    for i in range(40 * slabs):  # leave the synchronised start (waves of simultaneous resets)
        envs[i % slabs].step(acts[i % 4], auto_reset=True)
    torch.cuda.synchronize()
    K = 25 * slabs
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    for i in range(K):
        env = envs[i % slabs]
        ev[i][0].record()
        env.step(acts[i % 4])
        ev[i][1].record()
        env._reset_launch(env.t["done"], bench.WINDOW, True, None, None)
        ev[i][2].record()
    torch.cuda.synchronize()
    st = sorted(a.elapsed_time(b) for a, b, _ in ev)
    rs = sorted(b.elapsed_time(c) for _, b, c in ev)
    # fused call (step kernel appends to the list + refill), whole sequence back to back
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(K):
        envs[i % slabs].step(acts[i % 4], auto_reset=True)
    b.record()
    torch.cuda.synchronize()
    fused = a.elapsed_time(b) / K
    B = bench.bytes_per_env_step() * n
    med = st[len(st) // 2]
    print(f"[{tag}] N={n:8d} slabs={slabs}  step kernel median {med*1e3:7.1f} us (min {st[0]*1e3:.1f})  "
          f"= {B/(med*1e-3)/1e9:7.1f} GB/s = {B/(med*1e-3)/1e9/6549.4:.3f} of measured peak | masked reset (scan+refill) "
          f"median {rs[len(rs)//2]*1e3:6.1f} us | fused step+autoreset {fused*1e3:7.1f} us/step = {n/(fused*1e-3):.3e} env-steps/s "
          f"| done rate {float(envs[0].t['done'].float().mean()):.4f}", flush=True)
    del envs, acts
    torch.cuda.empty_cache()
