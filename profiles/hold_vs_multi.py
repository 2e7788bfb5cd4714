"""Where the one-wave step kernel's time goes: the same launch with no orders (Env.step() = hold), with orders, at
65,536 and 16,384 envs.  CUDA events around each launch, 8 rotating slabs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
dev = torch.device("cuda", 0)
for n in (65536, 16384):
    slabs = 8
    envs = [bench.make_env(dev, s * n, n_envs=n) for s in range(slabs)]
    acts = bench.synth_actions(4, n, 1, device=dev)
    for i in range(40 * slabs):
        envs[i % slabs].step(acts[i % 4], auto_reset=True)
    torch.cuda.synchronize()
    for mode in ("multi", "hold"):
        K = 25 * slabs
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(K)]
        for i in range(K):
            env = envs[i % slabs]
            ev[i][0].record()
            if mode == "multi":
                env.step(acts[i % 4])
            else:
                env.step()
            ev[i][1].record()
            env._reset_launch(env.t["done"], bench.WINDOW, True, None, None)
        torch.cuda.synchronize()
        st = sorted(a.elapsed_time(b) for a, b in ev)
        print(f"N={n:6d} {mode:5s}: median {st[len(st)//2]*1e3:6.1f} us  min {st[0]*1e3:6.1f} us", flush=True)
    del envs
    torch.cuda.empty_cache()
